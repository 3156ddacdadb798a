#!/usr/bin/env python3
"""CLI of the offline preprocessing stage (reference: bin/preprocess_data.py): merges
`config/config_datasets.yaml` and `config/deployment_options.yaml`, switches every dataset to its
preprocessing width, converts the fields of view to radians and runs `Preprocesser.preprocess_data()`
on the GPU.  `--yes` skips the reference's interactive path confirmation (batch jobs)."""
import os
import sys

import click
import numpy as np
import torch
import yaml

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from delora_b200.preprocessing import preprocesser as preprocesser_module  # noqa: E402


def build_config(config_dir="config"):
    config = {}
    for name in ("config_datasets.yaml", "deployment_options.yaml"):
        with open(os.path.join(config_dir, name)) as f:
            config.update(yaml.load(f, Loader=yaml.FullLoader))
    config["device"] = torch.device(config["device"])
    for dataset in config["datasets"]:
        config[dataset]["horizontal_cells"] = config[dataset]["horizontal_cells_preprocessing"]
        config[dataset]["data_identifiers"] = config[dataset]["training_identifiers"] + config[dataset]["testing_identifiers"]
        config[dataset]["vertical_field_of_view"] = [a * (np.pi / 180.0) for a in config[dataset]["vertical_field_of_view"]]
        if not os.path.exists(config[dataset]["data_path"]):
            raise Exception("Path " + config[dataset]["data_path"] + " does not exist. Exiting.")
    config["horizontal_field_of_view"] = [a * (np.pi / 180.0) for a in config["horizontal_field_of_view"]]
    return config


@click.command()
@click.option("--yes", is_flag=True, help="Do not ask for confirmation of the data paths.")
@click.option("--config_dir", default="config", help="Directory holding the reference's YAML files.")
def main(yes, config_dir):
    config = build_config(config_dir)
    print("Run for the datasets: " + str(config["datasets"]))
    for dataset in config["datasets"]:
        print(config[dataset]["data_path"] + " -> " + config[dataset]["preprocessed_path"])
    if not yes and not click.confirm("Continue?"):
        print("Okay, then program will be stopped.")
        return
    preprocesser_module.Preprocesser(config=config).preprocess_data()


if __name__ == "__main__":
    main()
