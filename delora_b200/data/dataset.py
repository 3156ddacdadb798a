"""Drop-in `data.dataset.PreprocessedPointCloudDataset` (reference: src/data/dataset.py:19-157).

Indexes `<preprocessed_path>/<seq:02d>/{scans,normals}/NNNNNN.npy` ([P,3] fp32 each, the format
`Preprocesser.apply_preprocessing_step` writes: src/preprocessing/preprocesser.py:64-68) and
returns, for index i, the dict of scans t and t+1 with the reference's keys (:143-153).  Host-side
loader; the tensors are [1,3,P] CPU tensors exactly like the reference's (`.to(device)` happens
in the trainer)."""
import glob
import os

import numpy as np
import torch


class PreprocessedPointCloudDataset(torch.utils.data.dataset.Dataset):

    def __init__(self, config):
        super().__init__()
        self.config = config
        self.store_dataset_in_RAM = self.config["store_dataset_in_RAM"]
        self.normals_files_in_datasets, self.scans_files_in_datasets = [], []
        self.index = []                                   # (dataset, sequence, scan) per pair
        for di, dataset in enumerate(self.config["datasets"]):
            normals_seq, scans_seq = [], []
            for si, ident in enumerate(self.config[dataset]["data_identifiers"]):
                seq_dir = os.path.join(self.config[dataset]["preprocessed_path"], format(ident, "02d") + "/")
                if not os.path.exists(seq_dir):
                    raise Exception("The specified path and dataset " + seq_dir + "does not exist.")
                normals = sorted(glob.glob(os.path.join(seq_dir, "normals/", "*.npy")))
                scans = sorted(glob.glob(os.path.join(seq_dir, "scans/", "*.npy")))
                normals_seq.append(normals)
                scans_seq.append(scans)
                self.index += [(di, si, k) for k in range(len(normals) - 1)]     # consecutive pairs (t, t+1)
            self.normals_files_in_datasets.append(normals_seq)
            self.scans_files_in_datasets.append(scans_seq)
        self.num_scans_overall = len(self.index)
        self.indices_dataset = np.array([i[0] for i in self.index], dtype=int)
        self.indices_sequence = np.array([i[1] for i in self.index], dtype=int)
        self.indices_scan = np.array([i[2] for i in self.index], dtype=int)
        self._ram = {}
        if self.store_dataset_in_RAM:
            for di, seqs in enumerate(self.scans_files_in_datasets):
                for si, files in enumerate(seqs):
                    for k in range(len(files)):
                        self._ram[(di, si, k)] = self.load_files_from_disk(di, si, k)

    def load_files_from_disk(self, index_dataset, index_sequence, index_scan):
        def load(path):
            return torch.from_numpy(np.load(path)).to(torch.device("cpu")).permute(1, 0).view(1, 3, -1)
        return (load(self.normals_files_in_datasets[index_dataset][index_sequence][index_scan]),
                load(self.scans_files_in_datasets[index_dataset][index_sequence][index_scan]))

    def _get(self, di, si, k):
        if self.store_dataset_in_RAM:
            return self._ram[(di, si, k)]
        return self.load_files_from_disk(di, si, k)

    def __getitem__(self, index):
        di, si, k = self.index[index]
        normal_list_1, scan_1 = self._get(di, si, k)
        normal_list_2, scan_2 = self._get(di, si, k + 1)
        return {"index": index, "index_dataset": di, "index_sequence": si, "index_scan": k,
                "dataset": self.config["datasets"][di], "normal_list_1": normal_list_1,
                "normal_list_2": normal_list_2, "scan_1": scan_1, "scan_2": scan_2}

    def __len__(self):
        return self.num_scans_overall
