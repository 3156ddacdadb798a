"""Device-side batching for the training / testing step.

The reference collates a batch as a Python list of per-sample dicts (`Deployer.list_collate`,
src/deploy/deployer.py:68-71), moves every tensor to the device one by one (src/deploy/trainer.py:62-66) and
loops over the samples inside `step` (:245-268).  Here a DataLoader worker packs the whole batch into ONE
staging buffer -- points [2B,3,Nmax] (scan_1 of every sample, then scan_2), normals [2B,3,Nmax], counts [2B] --
pinned when CUDA is present, and `PrefetchLoader` ships batch i+1 with a single asynchronous copy on its own
stream while batch i trains.  `Deployer.step` takes the resulting `PaddedBatch` directly.
"""
import torch


def _layout(b, n_max):
    def up(x):
        return (x + 255) // 256 * 256
    pts = 2 * b * 3 * n_max * 4
    off_n = up(pts)
    off_c = up(off_n + pts)
    return {"normals": off_n, "counts": off_c, "bytes": up(off_c + 2 * b * 4)}


class PaddedBatch:
    """points / normals [2B,3,Nmax] fp32, counts [2B] int32 as views of one flat uint8 buffer + per-sample metadata."""

    def __init__(self, flat, b, n_max, meta, counts_host=None):
        self.flat, self.B, self.n_max, self.meta = flat, int(b), int(n_max), meta
        self.counts_host = counts_host          # python ints (no device sync needed to slice a scan)
        lay = _layout(self.B, self.n_max)
        nb = 2 * self.B * 3 * self.n_max * 4
        self.points = flat[:nb].view(torch.float32).view(2 * self.B, 3, self.n_max)
        self.normals = flat[lay["normals"]:lay["normals"] + nb].view(torch.float32).view(2 * self.B, 3, self.n_max)
        self.counts = flat[lay["counts"]:lay["counts"] + 2 * self.B * 4].view(torch.int32)

    @property
    def dataset(self):
        return self.meta[0]["dataset"]

    def __len__(self):
        return self.B

    def to(self, device, non_blocking=False):
        return PaddedBatch(self.flat.to(device, non_blocking=non_blocking), self.B, self.n_max, self.meta, self.counts_host)

    def pin_memory(self):                       # DataLoader(pin_memory=True) hook
        if self.flat.is_pinned() or not torch.cuda.is_available():
            return self
        return PaddedBatch(self.flat.pin_memory(), self.B, self.n_max, self.meta, self.counts_host)


def padded_collate(batch_dicts):
    """collate_fn: list of dataset items (src/data/dataset.py:143-153) -> PaddedBatch on the host."""
    b = len(batch_dicts)
    n_max = max(max(d["scan_1"].shape[2], d["scan_2"].shape[2]) for d in batch_dicts)
    flat = torch.zeros((_layout(b, n_max)["bytes"],), dtype=torch.uint8)
    meta = [{k: v for k, v in d.items() if k not in ("scan_1", "scan_2", "normal_list_1", "normal_list_2")}
            for d in batch_dicts]
    out = PaddedBatch(flat, b, n_max, meta, [0] * (2 * b))
    for i, d in enumerate(batch_dicts):
        for half, key_s, key_n in ((0, "scan_1", "normal_list_1"), (1, "scan_2", "normal_list_2")):
            n = d[key_s].shape[2]
            out.points[half * b + i, :, :n] = d[key_s][0]
            out.normals[half * b + i, :, :n] = d[key_n][0]
            out.counts[half * b + i] = n
            out.counts_host[half * b + i] = int(n)
    return out


class PrefetchLoader:
    """Wraps a DataLoader that yields PaddedBatch objects: the host->device copy of the next batch runs on a
    side stream while the current one is consumed."""

    def __init__(self, dataloader, device):
        self.dataloader, self.device = dataloader, torch.device(device)
        self.stream = torch.cuda.Stream(device=self.device) if self.device.type == "cuda" else None

    def __len__(self):
        return len(self.dataloader)

    def _ship(self, batch):
        if self.stream is None:
            return batch.to(self.device), None
        with torch.cuda.stream(self.stream):
            dev = batch.pin_memory().to(self.device, non_blocking=True)
            ready = torch.cuda.Event()
            ready.record(self.stream)
        return dev, ready

    def __iter__(self):
        it = iter(self.dataloader)
        try:
            nxt = self._ship(next(it))
        except StopIteration:
            return
        while nxt is not None:
            cur, ready = nxt
            try:
                nxt = self._ship(next(it))
            except StopIteration:
                nxt = None
            if ready is not None:
                torch.cuda.current_stream(self.device).wait_event(ready)
                cur.flat.record_stream(torch.cuda.current_stream(self.device))
            yield cur
