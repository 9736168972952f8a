"""On-device input path of the spiking datasets (SHD / SSC): the dense (B, nb_steps, nb_units) batch the reference
builds on the host, example by example (sparch/dataloaders/spiking_datasets.py:41-86: ``np.digitize`` of the firing
times, a sparse tensor of ones, ``to_dense()``, ``pad_sequence``), built on the B200 from the raw event lists in one
launch.  The batch travels host -> device as events (8 bytes each) instead of as dense fp32 (280 KB per example).

    batcher = SpikingBatcher(nb_steps=100, device="cuda")          # nb_units = 700, max_time = 1.4 as the reference
    x, xlens, y = batcher(times_list, units_list, labels)           # what DataLoader(..., collate_fn=generateBatch) yields
    out, rates = net(x)

CUDA only (no CPU path: the reference's own loader is the CPU path)."""
import numpy as np
import torch

from ._lib import call, ptr


class SpikingBatcher:
    def __init__(self, nb_steps=100, nb_units=700, max_time=1.4, device="cuda"):
        self.nb_steps, self.nb_units, self.max_time = int(nb_steps), int(nb_units), float(max_time)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("sparch_b200.data.SpikingBatcher builds batches on a CUDA device only")
        self.time_bins = np.linspace(0, self.max_time, num=self.nb_steps)          # spiking_datasets.py:54
        self._bins = torch.from_numpy(self.time_bins).to(self.device)

    def pack(self, times_list, units_list, labels=None):
        """Host side: concatenate the examples' event arrays into pinned buffers (times fp32 -- the files hold fp16,
        the upcast is exact --, units int32, offsets int64)."""
        n = [len(t) for t in times_list]
        if [len(u) for u in units_list] != n:
            raise ValueError("times and units of an example must have the same length")
        off = np.zeros(len(n) + 1, np.int64)
        np.cumsum(n, out=off[1:])
        tot = int(off[-1])
        times = torch.empty(tot, dtype=torch.float32).pin_memory()
        units = torch.empty(tot, dtype=torch.int32).pin_memory()
        if tot:
            times.numpy()[:] = np.concatenate([np.asarray(t, np.float32) for t in times_list])
            units.numpy()[:] = np.concatenate([np.asarray(u, np.int64) for u in units_list])
        offsets = torch.from_numpy(off).pin_memory()
        y = None if labels is None else torch.as_tensor(np.asarray(labels), dtype=torch.int64).pin_memory()
        return times, units, offsets, y

    def to_dense(self, times, units, offsets, out=None, check=True):
        """Packed events (host or device tensors) -> dense (B, nb_steps, nb_units) fp32 counts on the device."""
        B = offsets.numel() - 1
        nev = int(offsets[-1]) if not offsets.is_cuda else None
        t_d = times.to(self.device, non_blocking=True)
        u_d = units.to(self.device, non_blocking=True)
        o_d = offsets.to(self.device, non_blocking=True)
        if nev is None:
            nev = int(o_d[-1].item())
        if out is None:
            out = torch.empty(B, self.nb_steps, self.nb_units, device=self.device, dtype=torch.float32)
        bad = torch.empty(1, device=self.device, dtype=torch.int32)
        with torch.cuda.device(self.device):
            call("sparch_events_to_dense", ptr(t_d), ptr(u_d), ptr(o_d), ptr(self._bins), B, self.nb_steps, self.nb_units,
                 nev, ptr(out), ptr(bad), torch.cuda.current_stream().cuda_stream)
        if check and int(bad.item()):
            # the reference raises from torch.sparse.FloatTensor (index out of range) in this case
            raise ValueError("an event lies outside [0, max_time) x [0, nb_units)")
        return out

    def __call__(self, times_list, units_list, labels):
        """(x, xlens, y) as ``SpikingDataset.generateBatch`` returns them (spiking_datasets.py:80-86), x and y on the
        device.  Every example has nb_steps rows, so ``pad_sequence`` is the identity and xlens is constant."""
        times, units, offsets, y = self.pack(times_list, units_list, labels)
        x = self.to_dense(times, units, offsets)
        xlens = torch.full((len(times_list),), self.nb_steps, dtype=torch.int64)
        return x, xlens, y.to(self.device, non_blocking=True)
