"""Training-loop memory helpers (rslo_b200/utils/memory.py): pre-sized per-stream allocator pools and the in-flight
limiter that `bench.py` uses to keep `cudaMalloc` out of the timed steps (profiles/r02_step_outliers.md)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_presized_pools_serve_later_allocations_without_cudamalloc(cuda):
    from rslo_b200.utils.memory import presize_stream_pools
    side = torch.cuda.Stream()
    presize_stream_pools([torch.cuda.current_stream(), side, None], large_bytes=64 << 20, small_bytes=8 << 20)
    n0 = torch.cuda.memory_stats()["num_device_alloc"]
    keep = []
    for st in (torch.cuda.current_stream(), side):
        with torch.cuda.stream(st):
            keep.append(torch.empty(24 << 20, dtype=torch.uint8, device="cuda"))      # large pool, split from the block
            keep.append(torch.empty(3 << 20, dtype=torch.uint8, device="cuda"))
            keep += [torch.empty(256 << 10, dtype=torch.uint8, device="cuda") for _ in range(8)]   # small pool
    assert torch.cuda.memory_stats()["num_device_alloc"] == n0


def test_inflight_limiter_bounds_the_host_lead(cuda):
    from rslo_b200.utils.memory import InflightLimiter
    lim = InflightLimiter(depth=2)
    x = torch.zeros(1 << 20, device="cuda")
    done = []
    for i in range(6):
        for _ in range(20):
            x.add_(1.0)
        ev = torch.cuda.Event()
        ev.record()
        done.append(ev)
        lim.tick()
        assert len(lim.events) <= 2
        if i >= 2:
            assert done[i - 2].query()          # the step two back has finished before the host moves on
    lim.reset()
    assert not lim.events
