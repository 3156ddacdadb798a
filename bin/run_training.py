#!/usr/bin/env python3
"""CLI with the reference's options (reference: bin/run_training.py): merges the three YAML
files of `config/`, converts the fields of view to radians, builds `deploy.trainer.Trainer`.
Run from a directory that holds `config/*.yaml` (the reference's own files work unchanged).
Under torchrun every rank binds its own GPU and the gradients are all-reduced over NCCL."""
import os
import sys

import click
import numpy as np
import torch
import yaml

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from delora_b200.deploy import trainer as trainer_module  # noqa: E402


def build_config(training_run_name, experiment_name="", checkpoint="", config_dir="config"):
    config = {}
    for name in ("config_datasets.yaml", "deployment_options.yaml", "hyperparameters.yaml"):
        with open(os.path.join(config_dir, name)) as f:
            config.update(yaml.load(f, Loader=yaml.FullLoader))
    loaded = torch.load(checkpoint, map_location="cpu", weights_only=False) if checkpoint else None
    if loaded is not None and "parameters" in loaded:            # the checkpoint's config wins (reference :36-55)
        saved = loaded["parameters"]
        saved["device"] = torch.device(config["device"])
        saved["datasets"] = config["datasets"]
        for ds in saved["datasets"]:
            saved[ds]["training_identifiers"] = config[ds]["training_identifiers"]
            saved[ds]["data_identifiers"] = saved[ds]["training_identifiers"]
        config = saved
    else:
        config["device"] = torch.device(config["device"])
        for ds in config["datasets"]:
            config[ds]["data_identifiers"] = config[ds]["training_identifiers"]
            config[ds]["vertical_field_of_view"] = [a * (np.pi / 180.0) for a in config[ds]["vertical_field_of_view"]]
        config["horizontal_field_of_view"] = [a * (np.pi / 180.0) for a in config["horizontal_field_of_view"]]
    config["checkpoint"] = str(checkpoint) if checkpoint else None
    config["training_run_name"] = config["run_name"] = str(training_run_name)
    if experiment_name:
        config["experiment"] = experiment_name
    config["mode"] = "training"
    return config


@click.command()
@click.option("--training_run_name", prompt="MLFlow name of the run",
              help="The name under which the run can be found afterwards.")
@click.option("--experiment_name", help="High-level training sequence name for clustering in MLFlow.", default="")
@click.option("--checkpoint", help="Path to the saved checkpoint. Leave empty if none.", default="")
def main(training_run_name, experiment_name, checkpoint):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    config = build_config(training_run_name, experiment_name, checkpoint)
    if world > 1:
        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        config["device"] = torch.device("cuda", local)
        torch.distributed.init_process_group("nccl", device_id=config["device"])
    trainer_module.Trainer(config=config).train()


if __name__ == "__main__":
    main()
