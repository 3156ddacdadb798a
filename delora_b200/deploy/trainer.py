"""Drop-in `deploy.trainer.Trainer` (reference: src/deploy/trainer.py): Adam(lr), epoch loop,
identity-pretraining switch (loss < 1e-2 -> unsupervised, :184-186), checkpoints with the
reference's keys (`epoch`, `model_state_dict`, `optimizer_state_dict`, `loss`, `parameters`,
:155-173).  MLflow / qqdm are used when importable and skipped otherwise.

Data parallel: when launched under torchrun (WORLD_SIZE > 1) every rank trains on its own shard
of the scan pairs and ALL parameter gradients (11.88 M) are averaged over NCCL, bucket by bucket and overlapped
with the backward (parallel_grad.BucketedGradAllReduce); the loss kernels stay rank-local."""
import os

import numpy as np
import torch

from . import deployer
from ..data import batching
from ..parallel_grad import make_grad_sync

try:
    import mlflow
except ImportError:                                       # logging is optional plumbing
    mlflow = None


class Trainer(deployer.Deployer):

    def __init__(self, config):
        super().__init__(config=config)
        self.training_bool = True
        # the reference's optimizer (src/deploy/trainer.py:23-24); on CUDA torch's fused implementation of the same update
        self.optimizer = torch.optim.Adam(params=self.model.parameters(), lr=self.config["learning_rate"],
                                          fused=str(self.device).startswith("cuda"))
        if self.config.get("checkpoint"):
            checkpoint = torch.load(self.config["checkpoint"], map_location=self.device, weights_only=False)
            self.model.load_state_dict(checkpoint["model_state_dict"])
            self.optimizer.load_state_dict(checkpoint["optimizer_state_dict"])
            self.config["unsupervised_at_start"] = True   # pretrained -> directly unsupervised (:35-36)
        if self.config["inference_only"]:
            print("Config error: Inference only does not make sense during training. Changing to inference_only=False.")
            self.config["inference_only"] = False
        # flat gradient buffer + per-bucket all-reduce overlapped with the backward (parallel_grad.py); a no-op
        # wrapper on one process without the tensor-core encoder
        self.grad_sync = make_grad_sync(self.model, self.config.get("grad_sync", "bucketed"))
        inner_step = self.optimizer.step

        def synced_step(*a, **k):
            self.grad_sync.finish()
            return inner_step(*a, **k)
        self.optimizer.step = synced_step

    @staticmethod
    def new_epoch_losses():
        return {"loss_epoch": 0.0, "loss_point_cloud_epoch": 0.0, "loss_field_of_view_epoch": 0.0,
                "loss_po2po_epoch": 0.0, "loss_po2pl_epoch": 0.0, "loss_pl2pl_epoch": 0.0,
                "visible_pixels_epoch": 0.0, "loss_yaw_pitch_roll_epoch": np.zeros(3), "loss_true_trafo_epoch": 0.0}

    def train_epoch(self, epoch, dataloader):
        epoch_losses = self.new_epoch_losses()
        for counter, preprocessed_dicts in enumerate(dataloader):
            if not isinstance(preprocessed_dicts, batching.PaddedBatch):          # reference collate: list of dicts
                for d in preprocessed_dicts:
                    for key in d:
                        if hasattr(d[key], "to"):
                            d[key] = d[key].to(self.device)
            self.optimizer.zero_grad()
            epoch_losses, _ = self.step(preprocessed_dicts=preprocessed_dicts, epoch_losses=epoch_losses,
                                        log_images_bool=False)
        return epoch_losses

    def _sampler(self):
        if self.grad_sync.world > 1:
            return torch.utils.data.distributed.DistributedSampler(self.dataset, num_replicas=self.grad_sync.world,
                                                                   rank=self.grad_sync.rank, shuffle=True)
        return None

    def train(self, max_epochs=10000):
        sampler = self._sampler()
        on_gpu = str(self.config["device"]).startswith("cuda")
        # config["device_batching"] (default on): workers pack each batch into one pinned staging buffer and the
        # copy of batch i+1 overlaps step i (data/batching.py); off = the reference's list-of-dicts collate
        device_batching = bool(self.config.get("device_batching", True))
        dataloader = torch.utils.data.DataLoader(dataset=self.dataset, batch_size=self.batch_size,
                                                 shuffle=sampler is None, sampler=sampler,
                                                 collate_fn=batching.padded_collate if device_batching else Trainer.list_collate,
                                                 num_workers=self.config["num_dataloader_workers"],
                                                 pin_memory=on_gpu)
        if device_batching:
            dataloader = batching.PrefetchLoader(dataloader, self.config["device"])
        run = None
        if mlflow is not None and self.grad_sync.rank == 0:
            mlflow.set_experiment(self.config["experiment"])
            run = mlflow.start_run(run_name="Training: " + self.config["training_run_name"])
        history = []
        for epoch in range(max_epochs):
            if sampler is not None:
                sampler.set_epoch(epoch)
            epoch_losses = self.train_epoch(epoch=epoch, dataloader=dataloader)
            steps = max(1, len(dataloader))
            for k in ("loss_epoch", "loss_point_cloud_epoch", "loss_po2po_epoch", "loss_po2pl_epoch",
                      "loss_pl2pl_epoch", "visible_pixels_epoch"):
                epoch_losses[k] = epoch_losses[k] / steps
            history.append(float(np.asarray(epoch_losses["loss_epoch"]).reshape(-1)[0]))
            if self.grad_sync.world > 1:
                # every rank must take the identity -> unsupervised switch in the same epoch: decide on the mean loss
                t = torch.tensor([history[-1]], dtype=torch.float64, device=self.device)
                torch.distributed.all_reduce(t)
                history[-1] = float(t[0]) / self.grad_sync.world
                peer = getattr(self.grad_sync, "peer", None)
                if peer is not None:
                    peer.check()        # a rank that missed a peer-memory collective this epoch is an error, not noise
            if self.grad_sync.rank == 0:
                print("Epoch Summary: " + format(epoch, "05d") + ", loss: " + str(epoch_losses["loss_epoch"]) +
                      ", unsupervised: " + str(self.config["unsupervised_at_start"]))
                if run is not None:
                    mlflow.log_metric("loss", history[-1], step=epoch)
                ckpt = {"epoch": epoch, "model_state_dict": self.model.state_dict(),
                        "optimizer_state_dict": self.optimizer.state_dict(), "loss": history[-1],
                        "parameters": self.config}
                path = os.path.join(self.config.get("checkpoint_dir", "/tmp"),
                                    self.config["training_run_name"] + "_latest_checkpoint.pth")
                torch.save(ckpt, path)
                if not epoch % 5:
                    torch.save(ckpt, path.replace("_latest_checkpoint", "_checkpoint_epoch_" + str(epoch)))
            if not self.config["unsupervised_at_start"] and history[-1] < 1e-2:     # :184-186
                self.config["unsupervised_at_start"] = True
                if self.grad_sync.rank == 0:
                    print("Loss has decreased sufficiently. Switching to unsupervised mode.")
        if run is not None:
            mlflow.end_run()
        return history
