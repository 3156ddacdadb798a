#!/usr/bin/env python3
"""CLI of the testing / inference stage with the reference's options (reference: bin/run_testing.py): merges the
three YAML files of `config/`, takes the network parameters from the checkpoint when it carries them, switches
the datasets to their testing sequences, disables dropout and runs `deploy.tester.Tester.test()`.
`--poses_dir` additionally writes one KITTI pose file per tested sequence (utility.poses)."""
import os
import sys

import click
import numpy as np
import torch
import yaml

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from delora_b200.deploy import tester as tester_module  # noqa: E402
from delora_b200.utility import poses as poses_module  # noqa: E402


def build_config(testing_run_name, experiment_name="testing", checkpoint="", config_dir="config"):
    config = {}
    for name in ("config_datasets.yaml", "deployment_options.yaml", "hyperparameters.yaml"):
        with open(os.path.join(config_dir, name)) as f:
            config.update(yaml.load(f, Loader=yaml.FullLoader))
    loaded = torch.load(checkpoint, map_location="cpu", weights_only=False) if checkpoint else {}
    if "parameters" in loaded:                                   # the checkpoint's network config wins (:31-52)
        saved = loaded["parameters"]
        saved["device"] = torch.device(config["device"])
        saved["datasets"] = config["datasets"]
        for ds in saved["datasets"]:
            saved[ds]["testing_identifiers"] = config[ds]["testing_identifiers"]
            saved[ds]["data_identifiers"] = saved[ds]["testing_identifiers"]
        saved["inference_only"] = config["inference_only"]
        saved["store_dataset_in_RAM"] = config["store_dataset_in_RAM"]
        config = saved
    else:
        config["device"] = torch.device(config["device"])
        for ds in config["datasets"]:
            config[ds]["data_identifiers"] = config[ds]["testing_identifiers"]
            config[ds]["vertical_field_of_view"] = [a * (np.pi / 180.0) for a in config[ds]["vertical_field_of_view"]]
        config["horizontal_field_of_view"] = [a * (np.pi / 180.0) for a in config["horizontal_field_of_view"]]
    if config.get("use_dropout"):
        config["use_dropout"] = False
        print("Deactivating dropout for this mode.")
    config["run_name"] = str(testing_run_name)
    config["checkpoint"] = str(checkpoint)
    if experiment_name:
        config["experiment"] = experiment_name
    config["mode"] = "testing"
    config["unsupervised_at_start"] = True
    return config


@click.command()
@click.option("--testing_run_name", prompt="MLFlow name of the run",
              help="The name under which the run can be found afterwards.")
@click.option("--experiment_name", help="High-level testing sequence name for clustering in MLFlow.", default="testing")
@click.option("--checkpoint", prompt="Path to the saved checkpoint of the model you want to test")
@click.option("--poses_dir", default="", help="Write <dataset>_<sequence>.txt KITTI pose files here.")
def main(testing_run_name, experiment_name, checkpoint, poses_dir):
    config = build_config(testing_run_name, experiment_name, checkpoint)
    tester = tester_module.Tester(config=config)
    tester.test()
    if poses_dir:
        os.makedirs(poses_dir, exist_ok=True)
        for di, dataset in enumerate(config["datasets"]):
            for si, ident in enumerate(config[dataset]["data_identifiers"]):
                poses_module.write_poses_to_text_file(
                    os.path.join(poses_dir, dataset + "_" + format(ident, "02d") + ".txt"), tester.poses(di, si))


if __name__ == "__main__":
    main()
