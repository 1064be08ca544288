"""Input side of the step (SURVEY §8 f-N1): voxelisation and index tables of the NEXT example while the current one
trains.

The reference does its voxelisation ahead of time in DataLoader worker processes (`rslo/data/preprocess.py:493` called
from `train_hdf5.py:44-89`) and ships ~11 MB of padded voxels per frame to the GPU.  Here the raw scan goes to the
device and `net.prepare(example)` runs the fused voxeliser + every sparse index table on the network's preparation
stream; this class moves that call - ~300 small kernel launches and the one device->host copy of the row counts, which
blocks its caller until the preparation stream has drained - onto a worker thread, so the training thread neither pays
the launch overhead nor waits for the counts.  ctypes and torch release the GIL inside the driver calls.
"""
import concurrent.futures as cf

import torch


class PreparedPrefetcher:
    """prefetcher = PreparedPrefetcher(net); fut = prefetcher.submit(example); ...; prepared = fut.result(); net(prepared)"""

    def __init__(self, net, device=None):
        self.net = net
        self.device = device if device is not None else next(net.parameters()).device
        self._pool = cf.ThreadPoolExecutor(max_workers=1, thread_name_prefix="rslo-prepare", initializer=self._init)

    def _init(self):
        if self.device.type == "cuda":
            torch.cuda.set_device(self.device)

    def _run(self, example):
        with torch.no_grad():
            return self.net.prepare(example)

    def submit(self, example):
        """-> concurrent.futures.Future of the prepared example (raw-scan input form, `example["points"]`)."""
        return self._pool.submit(self._run, example)

    def shutdown(self):
        self._pool.shutdown(wait=True)
