"""Drop-in `deploy.tester.Tester` (reference: src/deploy/tester.py): checkpoint load, batch size 1, walk the
test dataset pair by pair, collect the relative transforms per (dataset, sequence).

`test_dataset(dataloader)` follows the reference loop (:38-107); with `inference_only` every pair goes through
`Deployer.step`'s inference branch (projection + encoder + quaternion->T, no loss), otherwise the fused loss
kernels run as in training without the optimiser.  `poses(index_of_dataset, index_of_sequence)` chains the
collected transforms (utility.poses.compute_poses).  The reference's MLflow metric upload and the matplotlib
map plots (`log_map`, :111-162) are logging plumbing outside the hot path and are not reproduced.
For frame-by-frame streaming (one projection per frame, CUDA graph) see `deploy.stream.OdometryStream`.
"""
import torch

from . import deployer
from ..utility import poses as poses_module


class Tester(deployer.Deployer):

    def __init__(self, config):
        super().__init__(config=config)
        self.training = False
        self.training_bool = False
        if self.config.get("checkpoint"):
            checkpoint = torch.load(self.config["checkpoint"], map_location=self.device, weights_only=False)
            self.model.load_state_dict(checkpoint["model_state_dict"])
            print("Model weights loaded from " + str(self.config["checkpoint"]))
        else:
            raise Exception("No checkpoint specified.")
        print("Batch size set to 1 for the testing.")
        self.batch_size = 1
        self.model.eval()
        self.computed_transformations_datasets = [
            [[] for _ in self.config[dataset]["data_identifiers"]] for dataset in self.config["datasets"]]

    @staticmethod
    def new_epoch_losses():
        return {"loss_epoch": 0.0, "loss_point_cloud_epoch": 0.0, "loss_field_of_view_epoch": 0.0,
                "loss_po2po_epoch": 0.0, "loss_po2pl_epoch": 0.0, "loss_pl2pl_epoch": 0.0,
                "visible_pixels_epoch": 0.0}

    def test_dataset(self, dataloader):
        epoch_losses = self.new_epoch_losses()
        for index, preprocessed_dicts in enumerate(dataloader):
            for d in preprocessed_dicts:
                for key in d:
                    if hasattr(d[key], "to"):
                        d[key] = d[key].to(self.device)
            if not self.config["inference_only"]:
                epoch_losses, computed_transformation = self.step(preprocessed_dicts=preprocessed_dicts,
                                                                  epoch_losses=epoch_losses)
            else:
                with torch.no_grad():
                    computed_transformation = self.step(preprocessed_dicts=preprocessed_dicts,
                                                        epoch_losses=epoch_losses)
            for d in preprocessed_dicts:
                self.computed_transformations_datasets[d["index_dataset"]][d["index_sequence"]].append(
                    computed_transformation.detach().cpu().numpy())
            if not index % 10:
                print("Index: " + str(index) + " / " + str(len(dataloader)))
        return epoch_losses

    def poses(self, index_of_dataset=0, index_of_sequence=0):
        return poses_module.compute_poses(
            self.computed_transformations_datasets[index_of_dataset][index_of_sequence])

    def test(self):
        dataloader = torch.utils.data.DataLoader(dataset=self.dataset, batch_size=self.batch_size, shuffle=False,
                                                 collate_fn=Tester.list_collate,
                                                 num_workers=self.config["num_dataloader_workers"])
        epoch_losses = self.test_dataset(dataloader=dataloader)
        if not self.config["inference_only"]:
            for k in ("loss_epoch", "loss_point_cloud_epoch", "loss_po2po_epoch", "loss_po2pl_epoch",
                      "loss_pl2pl_epoch", "visible_pixels_epoch"):
                epoch_losses[k] /= self.steps_per_epoch
        return epoch_losses
