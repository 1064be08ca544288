"""Per-step diagnosis of step-time outliers in the bench loop: device time of every step next to the host time of
each phase, the number of cudaMalloc calls and Python GC passes that fell inside it.
    python scripts/gpu/spike_diag.py [steps] [variant...]      variants: sampler, nogc, inflight2, host, presize
"""
import gc
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    import torch
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    variants = set(sys.argv[2:])
    from_host = "host" in variants
    cfg = bench.workload_config("train")
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    run = bench.Runner(cfg, dev, 0, 1)
    if "reuse" in variants:          # upper bound of what removing the preparation-stream work could buy
        cache = {}
        orig = run.prepared_step
        def cached(i, fh):
            k = i % run.pool_n
            if k not in cache:
                cache[k] = orig(i, fh)
            return cache[k]
        run.prepared_step = cached
    for i in range(20):
        run.step(i, from_host)
    if "presize" in variants:
        run.presize_pools()
    if "inflight2" not in variants:
        run.limiter.depth = 1 << 30
    torch.cuda.synchronize()
    if "nogc" in variants:
        gc.collect()
        gc.freeze()
        gc.disable()
    sampler = bench.ClockSampler(0)
    if "sampler" in variants:
        sampler.start()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    rows = []
    evs[0].record()
    for i in range(steps):
        g0 = [s["collections"] for s in gc.get_stats()]
        m0 = torch.cuda.memory_stats()["num_device_alloc"]
        ph0 = dict(run.host_phase)
        t0 = time.perf_counter()
        run.flush.zero_()
        run.step(20 + i, from_host)
        evs[i + 1].record()
        t1 = time.perf_counter()
        g1 = [s["collections"] for s in gc.get_stats()]
        rows.append({"host_ms": 1e3 * (t1 - t0), "gc": [b - a for a, b in zip(g0, g1)],
                     "mallocs": torch.cuda.memory_stats()["num_device_alloc"] - m0,
                     "phases": {k: round(1e3 * (run.host_phase[k] - ph0[k]), 2) for k in ph0 if k != "steps"}})
    torch.cuda.synchronize()
    if "sampler" in variants:
        sampler.stop()
    per = [evs[i].elapsed_time(evs[i + 1]) for i in range(steps)]
    tot = evs[0].elapsed_time(evs[-1])
    med = sorted(per)[len(per) // 2]
    hmed = sorted(r["host_ms"] for r in rows)[len(rows) // 2]
    print(json.dumps({"variants": sorted(variants), "steps": steps, "mean_ms": tot / steps, "median_ms": med,
                      "host_median_ms": hmed, "reserved_GB": torch.cuda.memory_reserved() / 2**30}))
    for i, (d, r) in enumerate(zip(per, rows)):
        if d > 1.3 * med or r["host_ms"] > 1.5 * hmed or r["mallocs"] or r["gc"][2]:
            print(i, f"dev {d:.1f} ms", r)


if __name__ == "__main__":
    main()
