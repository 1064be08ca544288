#!/usr/bin/env python
"""bench.py — frame-pairs/s of RSLO's per-frame-pair hot path on B200 (contract: see DESIGN.md §Measurement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload train|eval|stress] [--impl reference]

A step = one pass of the hot path over one batch of synthetic KITTI-shaped scans:
  train (default; BASELINE.json configs[2]/[3]): 2 frame pairs per GPU, 120k-pt scans, 0.1 m voxels,
        voxelise -> sparse encoder -> head -> loss, forward + backward (+ one flat gradient
        all-reduce over NCCL when N > 1), global step > 1500 (icp_iter 2);
  eval  (configs[1]): 1 pair per GPU, forward only.
`value`   : pairs/s with the raw scans already resident in HBM.
`e2e`     : the same through net(example) with HOST (pinned) scans: H2D of the points and D2H of the
            loss / pose inside the timed region.
`roofline`: the dominant kernel (sparse-conv gather-GEMM) — algorithmic bytes / CUDA-event time,
            measured on a profiled replica of the timed steps (events bracket every C-ABI call).
`cpu_baseline`: the CPU oracle port (oracle/net.py) on this box's host cores, one pair.
--impl reference: times that CPU path alone (rank 0 only), same metric/config keys.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "frame-pairs/sec (fwd+bwd) on KITTI-shaped synthetic scans"
WEIGHT_SEED = 11
STEP_AFTER_WARMUP = 2000          # global step > 1500: predicted pose, icp_iter = 2 (voxel_odom_net.py:692-695)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="train", choices=["train", "eval", "stress"])
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--pairs-per-gpu", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-profile", action="store_true")
    ap.add_argument("--cpu-budget-s", type=float, default=150.0)
    return ap.parse_args()


def workload_config(args):
    if args.workload == "eval":
        ppg = args.pairs_per_gpu or 1
        return {"workload": "C2 eval fwd: 120k-pt pair (64 beams x 1875 az), voxel 0.1x0.1x0.2 m, <=40000 voxels/frame, "
                            "kitti_eval_ours model", "mode": "eval", "pairs_per_gpu": ppg, "beams": 64, "n_az": 1875}
    if args.workload == "stress":
        ppg = args.pairs_per_gpu or 1
        return {"workload": "C5 dense-scan stress: 128 beams x 2344 az (300k rays), voxel 0.05x0.05x0.2 m "
                            "(grid 2816x1536x40, BEV 192x352), <=250000 voxels/frame, fwd+bwd",
                "mode": "train", "pairs_per_gpu": ppg, "beams": 128, "n_az": 2344,
                "config_path": os.path.join(ROOT, "rslo_b200", "config", "stress_005.prototxt")}
    ppg = args.pairs_per_gpu or 2
    return {"workload": "C3/C4 train fwd+bwd: 2 pairs per GPU, 120k-pt scans (64 beams x 1875 az), voxel 0.1x0.1x0.2 m, "
                        "<=40000 voxels/frame, kitti_train_ours model section, Chamfer+covariance loss, step>1500",
            "mode": "train", "pairs_per_gpu": ppg, "beams": 64, "n_az": 1875}


def make_pairs(n, beams, n_az, first_seed):
    from rslo_b200.data import synthetic
    out = []
    for s in range(first_seed, first_seed + n):
        a, b, _ = synthetic.make_pair(s, n_beams=beams, n_az=n_az)
        out.append((a, b))
    return out


# ------------------------------------------------------------------------------------------------
# CPU arm (oracle port): cpu_baseline and --impl reference
# ------------------------------------------------------------------------------------------------
def cpu_run(cfg, steps, warmup, budget_s):
    """Times the CPU restatement of the reference path (oracle/net.py; the reference's own Python
    needs its un-vendored spconv/kornia/apex deps and cannot run on this box) on all host cores.
    Each step = ONE pair of the workload (bounded sample)."""
    import torch
    from oracle import net as onet
    import rslo_b200
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    net, _ = rslo_b200.build_network(testing=False, seed=7)
    onet.fill_weights(net, WEIGHT_SEED)
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    train = cfg["mode"] == "train"
    keys = [k for k, p in net.named_parameters() if p.requires_grad] if train else ()
    pairs = make_pairs(2, cfg["beams"], cfg["n_az"], 100)
    t_all0 = time.time()
    for i in range(min(warmup, 1)):
        onet.pair_forward(sd, list(pairs[i % 2]), training=train, step=STEP_AFTER_WARMUP, grads_for=keys)
    done, t0 = 0, time.time()
    while done < steps:
        onet.pair_forward(sd, list(pairs[done % 2]), training=train, step=STEP_AFTER_WARMUP, grads_for=keys)
        done += 1
        if time.time() - t_all0 > budget_s:
            break
    dt = time.time() - t0
    return {"value": done / dt, "unit": "pairs/s", "cores": cores, "kind": "port",
            "sample": f"{done} step(s) x 1 pair ({'fwd+bwd' if train else 'fwd'}) of the same workload, "
                      f"torch CPU {cores} threads + C oracle (single-thread voxeliser/NN)"}, done, dt


def reference_arm(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb, done, dt = cpu_run(cfg, args.steps, args.warmup, args.cpu_budget_s)
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "pairs/s", "n_gpus": args.gpus,
            "steps": done, "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * dt / max(done, 1),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {k: v for k, v in cfg.items() if k in ("workload", "mode", "pairs_per_gpu")},
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.rows.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for nm, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_traffic(kernel):
    """DRAM bytes per launch of `kernel` from the committed ncu --set full capture (profiles/), or None."""
    best = None
    pdir = os.path.join(ROOT, "profiles")
    if os.path.isdir(pdir):
        for fn in sorted(os.listdir(pdir)):
            if fn.endswith("_traffic.json"):
                try:
                    d = json.load(open(os.path.join(pdir, fn)))
                    if kernel in d:
                        best = d[kernel]["dram_bytes_per_launch_avg"]
                except Exception:
                    pass
    return best


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            if "hbm_gbs" in d:
                return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------
def main():
    args = parse_args()
    cfg = workload_config(args)
    if args.impl == "reference":
        return reference_arm(args, cfg)

    import torch
    import torch.distributed as dist
    import rslo_b200
    from rslo_b200.utils.weights import deterministic_fill
    from rslo_b200 import kernels as K
    from rslo_b200.utils.distributed import FlatGradAllReducer, init_from_env

    rank, local, world = init_from_env()
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    train = cfg["mode"] == "train"
    ppg = cfg["pairs_per_gpu"]

    net, vg = rslo_b200.build_network(cfg.get("config_path"), testing=False, seed=7)
    deterministic_fill(net, WEIGHT_SEED)
    net = net.to(dev)
    net.global_step.fill_(STEP_AFTER_WARMUP)
    net._step_host = None
    net.train(train)
    reducer = FlatGradAllReducer(net) if train else None

    # inputs: a pool of distinct pairs per rank (rotated so no step re-reads the previous step's scans)
    pool_n = max(2 * ppg, 4)
    pairs = make_pairs(pool_n, cfg["beams"], cfg["n_az"], first_seed=1000 * rank)
    host = [(torch.from_numpy(a).pin_memory(), torch.from_numpy(b).pin_memory()) for a, b in pairs]
    resident = [(a.to(dev), b.to(dev)) for a, b in host]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)       # > 126 MB L2

    h2d_bytes = [0]
    d2h_bytes = [0]

    image_caches = [m._images for m in net.modules() if hasattr(m, "_images")]
    PREFETCH_DEPTH = 1                       # examples prepared ahead (2 measured no faster, and slower from host)
    prefetch = {}                            # (pair index, from_host) -> prepared example

    def prepared_step(i, from_host):
        """net.prepare(): voxelisation + index tables of ALL samples of step i on a side stream (the reference does
        its voxelisation ahead of time in DataLoader workers).  From host: the pinned scans are copied inside prepare()."""
        src = host if from_host else resident
        pts = []
        for j in range(ppg):
            a, b = src[(i * ppg + j) % pool_n]
            pts += [a, b]
            if from_host:
                h2d_bytes[0] += a.numel() * 4 + b.numel() * 4
        return net.prepare({"points": pts, "n_samples": ppg, "host_outputs": False})

    def step(i, from_host):
        """one step = the ppg samples of the batch through ONE net(example) call (example["n_samples"] = ppg):
        loss = mean over the samples; one backward"""
        if train:
            reducer.zero_()
            for c in image_caches:          # a real training step changes the weights: rebuild the split-TF32
                c._c.clear()                # weight images once per step, as an optimizer step would force
        ex = prefetch.pop((i, from_host), None)
        if ex is None:
            ex = prepared_step(i, from_host)
        if train:
            ret = net(ex)
            ret["loss"].sum().backward()
            res = ret["loss"].detach().reshape(-1)
        else:
            with torch.no_grad():
                ret = net(ex)
            res = torch.cat([ret["translation_preds"], ret["rotation_preds"]], -1).reshape(-1)
        # the next step's samples are prepared while this one's kernels run (the preparation's only host wait
        # then happens with a whole step still queued on the main stream)
        for stale in [k for k in prefetch if k[1] != from_host or k[0] <= i]:
            del prefetch[stale]
        for ahead in range(1, PREFETCH_DEPTH + 1):
            if (i + ahead, from_host) not in prefetch:
                prefetch[(i + ahead, from_host)] = prepared_step(i + ahead, from_host)
        if train:
            reducer.all_reduce()                    # pack into the flat buffer (+ NCCL all-reduce when N > 1)
        if from_host:
            r = res.cpu()                                                  # D2H of the step's result
            d2h_bytes[0] += r.numel() * 4
        return res

    def timed(nsteps, from_host, first):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(nsteps):
            flush.zero_()                                                  # L2 flush between steps
            step(first + i, from_host)
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    # warm-up (>= 3 steps), then the timed region
    W = max(args.warmup, 3)
    for i in range(W):
        step(i, False)
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = K.kernel_launch_count()
    cuda_prof = os.environ.get("RSLO_BENCH_CUDA_PROFILER") == "1"      # ncu --profile-from-start off
    if cuda_prof:
        torch.cuda.profiler.start()
    ms = timed(args.steps, False, W)
    if cuda_prof:
        torch.cuda.profiler.stop()
    launches = K.kernel_launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    value = world * ppg * args.steps / (ms / 1e3)

    # e2e: host buffers -> net(example) -> host result
    for i in range(2):
        step(i, True)
    h2d_bytes[0] = d2h_bytes[0] = 0
    ms_e2e = timed(args.steps, True, W)
    e2e = {"value": world * ppg * args.steps / (ms_e2e / 1e3), "unit": "pairs/s",
           "h2d_bytes_per_step": h2d_bytes[0] // (args.steps + PREFETCH_DEPTH), "d2h_bytes_per_step": d2h_bytes[0] // args.steps,
           "ms_per_step": ms_e2e / args.steps}

    # roofline of the dominant kernel: profiled replica of the timed steps
    roofline, breakdown = None, None
    if not args.no_profile:                  # every rank runs the replica (its steps contain the collective)
        K.PROFILE = []
        nprof = min(args.steps, 5)
        step(W, False)                       # fills the rulebook-size caches outside the measured calls
        torch.cuda.synchronize()
        K.PROFILE = []
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(nprof):
            flush.zero_()
            step(W + i, False)
        e1.record()
        torch.cuda.synchronize()
        step_ms = e0.elapsed_time(e1) / nprof
        agg = {}
        for name, s, e, nbytes, flops in K.PROFILE:
            a = agg.setdefault(name, [0, 0.0, 0, 0])
            a[0] += 1
            a[1] += s.elapsed_time(e)
            a[2] += nbytes
            a[3] += flops
        K.PROFILE = None
    if not args.no_profile and rank == 0:
        breakdown = {n: {"calls_per_step": a[0] / nprof, "ms_per_step": a[1] / nprof,
                         "share_of_step": a[1] / nprof / step_ms,
                         "algorithmic_GBps": a[2] / (a[1] * 1e-3) / 1e9 if a[1] > 0 else None}
                     for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1])}
        breakdown["_step_ms_profiled"] = step_ms
        dom = max(agg.items(), key=lambda kv: kv[1][1])
        peak, peak_src = measured_peaks()
        nm, a = dom
        achieved = a[2] / (a[1] * 1e-3) / 1e9
        roofline = {"kernel": nm, "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": measured_traffic(nm), "peak_source": peak_src,
                    "launches_per_step": a[0] / nprof, "avg_launch_ms": a[1] / a[0],
                    "algorithmic_bytes_per_launch": a[2] / a[0], "gflops_per_launch": a[3] / a[0] / 1e9,
                    "share_of_step": a[1] / nprof / step_ms,
                    # the same launches seen as a contraction: useful FLOP/s (2*R*Cin*Cout); the split-TF32 scheme
                    # executes 3x that on the tensor pipe (ncu: 32 % tensor-pipe active, smem-bandwidth bound)
                    "useful_tflops": (a[3] / (a[1] * 1e-3) / 1e12) if a[1] > 0 else None,
                    "note": "dominant OWN kernel by summed CUDA-event time; the cuDNN FP32 head is the larger share "
                            "of the step (profiles/r01_launches_bench.md)"}
    if world > 1:
        dist.barrier()

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_baseline, _, _ = cpu_run(cfg, 2, 0, 40.0)

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps, "warmup": W,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": cfg["workload"], "mode": cfg["mode"], "pairs_per_gpu": ppg,
                           "global_pairs_per_step": ppg * world, "parallelism": f"dp{world}",
                           "l2": "256 MB memset between steps (inside the timed region); inputs rotate over "
                                 f"{pool_n} distinct pairs per rank",
                           "grad_allreduce_bytes": reducer.nbytes if (train and world > 1) else 0},
                "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline,
                "cpu_baseline": cpu_baseline, "kernel_breakdown": breakdown}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
