"""Drop-in `data.kitti_scans` (reference: src/data/kitti_scans.py).

`KITTIPointCloudDataset` lists `<base_dir>/<seq:02d>/velodyne/*.bin` and returns each scan as a
[4, N] float32 tensor (x, y, z, reflectance), like `get_velo_torch` (:46-50).  The reference reads the
files through `pykitti.utils.load_velo_scan`, which is `np.fromfile(file, dtype=np.float32).reshape(-1, 4)`
(pykitti 0.3.1, utils.py) -- restated here so that the third-party package is not needed.
`KITTIDatasetPreprocessor.preprocess()` walks the sequence like the reference (:25-32) but hands the
preprocessing function `batch_size` scans at a time when it accepts batches (`preprocessing_fct_batch`).
"""
import glob
import os

import numpy as np
import torch


def load_velo_scan(file):
    scan = np.fromfile(file, dtype=np.float32)
    return scan.reshape((-1, 4))


class KITTIPointCloudDataset(torch.utils.data.dataset.Dataset):
    def __init__(self, base_dir, identifier="00", device=torch.device("cuda")):
        super().__init__()
        self.base_dir = base_dir
        self.identifier = identifier
        self.device = device
        ident = format(self.identifier, '02d') if isinstance(self.identifier, int) else str(self.identifier)
        self.velo_file_list = sorted(glob.glob(os.path.join(self.base_dir, ident, "velodyne", '*.bin')))
        self.num_elements = len(self.velo_file_list)

    def get_velo(self, idx):
        return load_velo_scan(self.velo_file_list[idx])

    def get_velo_torch(self, idx):
        return torch.from_numpy(self.get_velo(idx)).to(torch.device("cpu")).transpose(0, 1)

    def __getitem__(self, index):
        return self.get_velo_torch(idx=index)

    def __len__(self):
        return self.num_elements


class KITTIDatasetPreprocessor():
    def __init__(self, config, dataset_name, preprocessing_fct, preprocessing_fct_batch=None, batch_size=16):
        self.config = config
        self.identifier = self.config[dataset_name]["data_identifier"]
        self.point_cloud_dataset = KITTIPointCloudDataset(base_dir=self.config[dataset_name]["data_path"],
                                                          identifier=self.identifier,
                                                          device=self.config["device"])
        self.preprocessing_fct = preprocessing_fct
        self.preprocessing_fct_batch = preprocessing_fct_batch
        self.batch_size = int(batch_size)

    def preprocess(self):
        n = self.point_cloud_dataset.num_elements
        ident = format(self.identifier, '02d') if isinstance(self.identifier, int) else str(self.identifier)
        if self.preprocessing_fct_batch is None:
            for index in range(n):
                if not index % 10:
                    print("Preprocessing scan " + str(index) + "/" + str(n) + " from sequence " + ident + ".")
                scan = self.point_cloud_dataset.get_velo_torch(index).unsqueeze(0)
                self.preprocessing_fct(scan=scan, index=index)
            return
        for start in range(0, n, self.batch_size):
            idx = list(range(start, min(n, start + self.batch_size)))
            print("Preprocessing scans " + str(idx[0]) + "-" + str(idx[-1]) + "/" + str(n) + " from sequence " + ident + ".")
            self.preprocessing_fct_batch(scans=[self.point_cloud_dataset.get_velo_torch(i) for i in idx], indices=idx)
