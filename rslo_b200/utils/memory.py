"""Device-memory plumbing of the training loop (torch's caching allocator is the memory manager here).

Why this exists: torch's allocator keeps one pool PER STREAM, and the step uses three (main, `net.prepare`'s
preparation stream, the gradient communication stream).  When the host runs ahead of the device, blocks handed from
the preparation stream to the main stream (`record_stream`) are not reusable until the main stream has passed them,
so the pools keep growing by a `cudaMalloc` every few steps for hundreds of steps.  A `cudaMalloc` is harmless on an
idle driver but serialises with NVML polling (`nvidia-smi -lms`, DCGM): measured on B200, steps that contained one
stalled for 10-100 ms while a clock sampler was running (`scripts/gpu/spike_diag.py`).  Two remedies, both used by
`bench.py` and recommended for any training loop:

* `presize_stream_pools`: grow each stream's pools once, up front, so steady-state steps never reach the driver;
* `InflightLimiter`: keep the host at most `depth` steps ahead of the device, which bounds the pools' high-water mark.
"""
import collections

import torch


def presize_stream_pools(streams, large_bytes=1 << 30, small_bytes=128 << 20):
    """Reserve `large_bytes` of splittable large-pool space and `small_bytes` of small-pool space (allocations
    <= 1 MiB) in the caching allocator for every stream in `streams`, then hand them back to the cache."""
    for st in streams:
        if st is None:
            continue
        with torch.cuda.stream(st):
            big = torch.empty(large_bytes, dtype=torch.uint8, device=st.device)
            small = [torch.empty(1 << 20, dtype=torch.uint8, device=st.device) for _ in range(small_bytes >> 20)]
            del big, small


class InflightLimiter:
    """`tick()` at the end of every step: records an event on the current stream and blocks the host until the step
    `depth` steps back has finished on the device."""

    def __init__(self, depth=2):
        self.depth = depth
        self.events = collections.deque()

    def tick(self):
        ev = torch.cuda.Event()
        ev.record()
        self.events.append(ev)
        while len(self.events) > self.depth:
            self.events.popleft().synchronize()

    def reset(self):
        self.events.clear()
