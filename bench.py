#!/usr/bin/env python
"""bench.py — frame-pairs/s of RSLO's per-frame-pair hot path on B200 (contract: see DESIGN.md §Measurement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload train|eval|warm|stress] [--impl reference]

A step = one pass of the hot path over one batch of synthetic KITTI-shaped scans:
  train (default; BASELINE.json configs[2]/[3]): 2 frame pairs per GPU through ONE net(example) call
        (example["n_samples"] = 2), 120k-pt scans, 0.1 m voxels, voxelise -> sparse encoder -> head -> loss,
        forward + backward (+ one flat gradient all-reduce over NCCL when N > 1), global step > 1500 (icp_iter 2);
  eval  (configs[1]): 1 pair per GPU, forward only;
  warm  : train at global step <= 1500 (identity pose, icp_iter 5: `voxel_odom_net.py:677-695`);
  stress (configs[4]): 300k-ray scans, 0.05 x 0.05 x 0.1 m voxels (grid 2816 x 1536 x 80), fwd+bwd.
`value`   : pairs/s with the raw scans already resident in HBM (total time of exactly K steps).
`e2e`     : the same through net.prepare / net(example) with HOST (pinned) scans: H2D of the points and D2H of the
            loss / pose inside the timed region.
`roofline`: the kernel with the largest summed time in a step (profiled replica with CUDA events around every
            C-ABI call, head un-graphed for that replica) against the roofline that bounds it; `roofline_hbm`:
            the largest HBM-bound own kernel.
`cpu_baseline` / --impl reference: the CPU oracle port (oracle/net.py) on this box's host cores.
`extra`   : (N = 1 only) the other BASELINE configs as pairs/s: eval, warm-up mode, dense-scan stress with a sweep of
            the voxel cap, and the exact-NN kernel next to the reference's own CUDA kernel (oracle/_ref/cd_ref.so).
"""
import argparse
import importlib.util
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
STEP_WARM_MODE = 100              # global step <= 1500: identity pose, icp_iter = 5


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--workload", default="train", choices=["train", "eval", "warm", "stress"])
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--pairs-per-gpu", type=int, default=None)
    ap.add_argument("--max-voxels", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-profile", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--cpu-budget-s", type=float, default=240.0)
    return ap.parse_args()


# optimizer of kitti_train_ours.prototxt:146-158 at the first step of its one-cycle schedule (lr_max 0.8e-3 / div_factor 10,
# moms[0] 0.95, betas[1] 0.99 with fixed_weight_decay, weight_decay 1e-5) + train_hdf5.py:671 clip at 10
OPT = {"lr": 0.8e-4, "mom": 0.95, "beta": 0.99, "eps": 1e-8, "wd": 1e-5, "max_norm": 10.0}


def optimizer_note(train):
    return (f"every train step ends with clip_grad_norm {OPT['max_norm']:g} + true weight decay + Adam (lr {OPT['lr']}, "
            f"betas ({OPT['mom']}, {OPT['beta']}), wd {OPT['wd']})") if train else None


def workload_config(workload, pairs_per_gpu=None, max_voxels=None):
    if workload == "eval":
        return {"workload": "C2 eval fwd: 120k-pt pair (64 beams x 1875 az), voxel 0.1x0.1x0.2 m, <=40000 voxels/frame, "
                            "kitti_eval_ours model", "mode": "eval", "pairs_per_gpu": pairs_per_gpu or 1, "beams": 64,
                "n_az": 1875, "global_step": STEP_AFTER_WARMUP}
    if workload == "stress":
        return {"workload": "C5 dense-scan stress: 128 beams x 2344 az (300k rays), voxel 0.05x0.05x0.1 m "
                            f"(grid 2816x1536x80, BEV 192x352x256ch), <={max_voxels or 250000} voxels/frame, fwd+bwd",
                "mode": "train", "pairs_per_gpu": pairs_per_gpu or 1, "beams": 128, "n_az": 2344,
                "global_step": STEP_AFTER_WARMUP, "max_voxels": max_voxels or 250000,
                "config_path": os.path.join(ROOT, "rslo_b200", "config", "stress_005.prototxt")}
    if workload == "warm":
        return {"workload": "C3 warm-up mode: train fwd+bwd at global step <= 1500 (identity pose, icp_iter 5), 2 pairs per "
                            "GPU, 120k-pt scans, voxel 0.1x0.1x0.2 m", "mode": "train", "pairs_per_gpu": pairs_per_gpu or 2,
                "beams": 64, "n_az": 1875, "global_step": STEP_WARM_MODE}
    return {"workload": "C3/C4 train fwd+bwd: 2 pairs per GPU, 120k-pt scans (64 beams x 1875 az), voxel 0.1x0.1x0.2 m, "
                        "<=40000 voxels/frame, kitti_train_ours model section, Chamfer+covariance loss, step>1500",
            "mode": "train", "pairs_per_gpu": pairs_per_gpu or 2, "beams": 64, "n_az": 1875,
            "global_step": STEP_AFTER_WARMUP}


def public_config(cfg, world=1, extra=None):
    out = {"workload": cfg["workload"], "mode": cfg["mode"], "pairs_per_gpu": cfg["pairs_per_gpu"],
           "global_pairs_per_step": cfg["pairs_per_gpu"] * world, "parallelism": f"dp{world}",
           "global_step": cfg["global_step"]}
    out.update(extra or {})
    return out


def b200_config(cfg, world, ppg, reducer_bytes, train):
    """`config` of the B200 arm's line; the reference arm carries the same keys (tests/test_cpu_host.py)"""
    return public_config(cfg, world, {
        "l2": "256 MB memset between steps (inside the timed region); inputs rotate over "
              f"{max(2 * ppg, 4)} distinct pairs per rank",
        "steps_in_flight": f"host at most {Runner.MAX_STEPS_IN_FLIGHT} steps ahead of the device (event wait)",
        "grad_allreduce_bytes": reducer_bytes,
        "optimizer": optimizer_note(train),
        "note": "one process per GPU; value = pairs of all ranks / max-over-ranks device time of the timed steps"})


def reference_config(cfg, cpu_budget_s):
    return public_config(cfg, 1, {
        "l2": "n/a (CPU arm)", "steps_in_flight": "n/a (CPU arm)", "grad_allreduce_bytes": 0,
        "optimizer": optimizer_note(cfg["mode"] == "train"),
        "note": f"CPU arm: rank 0 only, one replica's batch per step; steps/warm-up bounded by --cpu-budget-s {cpu_budget_s:.0f}"})


def make_pairs(n, beams, n_az, first_seed):
    from rslo_b200.data import synthetic
    out = []
    for s in range(first_seed, first_seed + n):
        a, b, _ = synthetic.make_pair(s, n_beams=beams, n_az=n_az)
        out.append((a, b))
    return out


# ------------------------------------------------------------------------------------------------
# CPU arm (oracle port): cpu_baseline and --impl reference.  Nothing here imports the product's kernels:
# the weights come from the committed shape table + the per-key deterministic fill.
# ------------------------------------------------------------------------------------------------
def reference_state_dict(seed):
    import torch
    spec = importlib.util.spec_from_file_location("_rslo_weights", os.path.join(ROOT, "rslo_b200", "utils", "weights.py"))
    wmod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(wmod)                       # plain torch/numpy module; does not load librslo_b200.so
    shapes = json.load(open(os.path.join(ROOT, "tests", "golden", "state_dict_shapes.json")))
    sd = {}
    for k, shp in shapes.items():
        if k == "global_step" or k.endswith("num_batches_tracked"):
            sd[k] = torch.zeros(shp, dtype=torch.long)
        else:
            sd[k] = torch.zeros(shp, dtype=torch.float32)
    # values the deterministic fill leaves alone: the shipped config's initial values
    sd["_rotation_loss.alpha"].fill_(-2.5)
    sd["_pyramid_rotation_loss.alpha"].fill_(-2.5)
    for k in sd:
        if "dynamic_sigma" in k:
            sd[k].fill_(0.1)
    if "_consistency_loss.svd.reflect" in sd:
        sd["_consistency_loss.svd.reflect"].copy_(torch.diag(torch.tensor([1.0, 1.0, -1.0])))
    wmod.deterministic_fill(sd, seed)
    return sd


def cpu_run(cfg, steps, warmup, budget_s):
    """Times the CPU restatement of the reference path (oracle/net.py; the reference's own Python needs its
    un-vendored spconv/kornia/apex deps and cannot run on this box) on all host cores.  A step = the same batch as the
    B200 arm's step (pairs_per_gpu pairs, processed one after the other as the reference would)."""
    import torch
    from oracle import net as onet
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    sd = reference_state_dict(WEIGHT_SEED)
    train = cfg["mode"] == "train"
    keys = [k for k, v in sd.items() if v.is_floating_point() and not k.endswith(("running_mean", "running_var", "reflect"))
            and k != "_consistency_loss.alpha"] if train else ()
    ppg = cfg["pairs_per_gpu"]
    pairs = make_pairs(max(2, ppg), cfg["beams"], cfg["n_az"], 100)

    from oracle import optim as oopt
    opt_state = oopt.new_state([sd[k] for k in keys]) if train else None

    def one_step(i):
        acc = None
        for j in range(ppg):
            out = onet.pair_forward(sd, list(pairs[(i * ppg + j) % len(pairs)]), training=train, step=cfg["global_step"],
                                    grads_for=keys)
            if train:
                gs = [None if out["grads"][k] is None else torch.from_numpy(out["grads"][k]) / ppg for k in keys]
                acc = gs if acc is None else [b if a is None else (a if b is None else a + b) for a, b in zip(acc, gs)]
        if train:           # clip_grad_norm_(10) + OptimWrapper.step() (true weight decay + Adam), as train_hdf5.py:671-672
            _, clipped = oopt.clip_grad_norm(acc, OPT["max_norm"])
            with torch.no_grad():
                oopt.adam_step([sd[k] for k in keys], clipped, opt_state, OPT["lr"], OPT["mom"], OPT["beta"], OPT["eps"],
                               wd=OPT["wd"], true_wd=True)

    t_all0 = time.time()
    w_done = 0
    for i in range(warmup):
        one_step(i)
        w_done += 1
        if time.time() - t_all0 > budget_s / 3:
            break
    done, t0 = 0, time.time()
    while done < steps:
        one_step(done)
        done += 1
        if time.time() - t_all0 > budget_s:
            break
    dt = time.time() - t0
    return {"value": done * ppg / dt, "unit": "pairs/s", "cores": cores, "kind": "port",
            "sample": f"{done} step(s) x {ppg} pair(s) ({'fwd+bwd' if train else 'fwd'}) of the same workload after "
                      f"{w_done} warm-up step(s), torch CPU {cores} threads + C oracle (single-thread voxeliser/NN)"}, done, w_done, dt


def reference_arm(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # bounded sample: the same step (same pairs per step, same mode) as the B200 arm, as many as fit the budget
    cb, done, w_done, dt = cpu_run(cfg, args.steps, args.warmup, args.cpu_budget_s)
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "pairs/s", "n_gpus": args.gpus,
            "steps": done, "warmup": w_done, "ms_per_step": 1e3 * dt / max(done, 1),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": reference_config(cfg, args.cpu_budget_s),
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
                                          "--format=csv,noheader,nounits", "-lms", "200"],
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
    """DRAM bytes per launch of `kernel` from the committed ncu --set full captures (profiles/*_traffic.json): the capture
    that averaged over the most launches of a step (ties: the latest), or None."""
    best, best_n = None, -1
    pdir = os.path.join(ROOT, "profiles")
    if os.path.isdir(pdir):
        for fn in sorted(os.listdir(pdir)):
            if fn.endswith("_traffic.json"):
                try:
                    d = json.load(open(os.path.join(pdir, fn)))
                    if kernel in d and int(d[kernel].get("launches_captured", 0)) >= best_n:
                        best, best_n = d[kernel]["dram_bytes_per_launch_avg"], int(d[kernel].get("launches_captured", 0))
                except Exception:
                    pass
    return best


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    d = {}
    if os.path.exists(p):
        try:
            d = json.load(open(p))
        except Exception:
            d = {}
    hbm = (float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)") if "hbm_gbs" in d else \
        (6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)")
    if "bf16_tflops_sustained" in d:
        tf32 = (float(d["bf16_tflops_sustained"]) / 2, "half of MEASURED_PEAKS.json bf16_tflops_sustained (kind::tf32 runs "
                                                        "at half the bf16 rate; kernel timed inside a long step)")
    else:
        tf32 = (2250.0 / 2, "fallback: half of the nominal 2.25 PFLOP/s dense bf16")
    return hbm, tf32


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------
class Runner:
    """One workload on this rank: network, a pool of synthetic pairs (device-resident and pinned-host copies), the
    step function through the public API (net.prepare on a side stream one step ahead, net(example), backward,
    flat gradient all-reduce) and the timed loop."""
    PREFETCH_DEPTH = 1
    MAX_STEPS_IN_FLIGHT = 2           # the host enqueues at most this many steps ahead of the device

    def __init__(self, cfg, dev, rank, world):
        import torch
        import rslo_b200
        from rslo_b200.utils.distributed import FlatGradAllReducer
        from rslo_b200.utils.weights import deterministic_fill
        self.torch, self.cfg, self.dev, self.world = torch, cfg, dev, world
        self.train = cfg["mode"] == "train"
        self.ppg = cfg["pairs_per_gpu"]
        net, vg = rslo_b200.build_network(cfg.get("config_path"), testing=False, seed=7)
        if cfg.get("max_voxels"):
            vg.max_voxels_per_call = int(cfg["max_voxels"])
        deterministic_fill(net, WEIGHT_SEED)
        net = net.to(dev)
        net.global_step.fill_(cfg["global_step"])
        net._step_host = None
        net.train(self.train)
        self.net, self.vg = net, vg
        self.reducer = self.opt = None
        if self.train:
            from rslo_b200.torchplus.train import FusedAdamClip
            self.reducer = FlatGradAllReducer(net, early_module=net.odom_predictor)
            self.opt = FusedAdamClip(self.reducer, lr=OPT["lr"], wd=OPT["wd"], true_wd=True, betas=(OPT["mom"], OPT["beta"]),
                                     eps=OPT["eps"], max_norm=OPT["max_norm"])
        self.pool_n = max(2 * self.ppg, 4)
        pairs = make_pairs(self.pool_n, cfg["beams"], cfg["n_az"], first_seed=1000 * rank)
        self.host = [(torch.from_numpy(a).pin_memory(), torch.from_numpy(b).pin_memory()) for a, b in pairs]
        self.resident = [(a.to(dev), b.to(dev)) for a, b in self.host]
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)       # > 126 MB L2
        self.h2d_bytes = self.d2h_bytes = 0
        self.prefetch = {}
        self.prefetcher = None
        if os.environ.get("RSLO_BENCH_PREFETCH_THREAD", "0") != "0":     # measured: GIL contention outweighs the overlap
            from rslo_b200.data.prefetch import PreparedPrefetcher
            self.prefetcher = PreparedPrefetcher(net, dev)
        from rslo_b200.utils.memory import InflightLimiter
        self.limiter = InflightLimiter(self.MAX_STEPS_IN_FLIGHT)
        self.result_slots = [[None, None], [None, None]]     # pinned result buffer + copy-done event, double-buffered
        self.last_result = None
        self.host_s = 0.0
        self.host_phase = {"forward": 0.0, "backward": 0.0, "prepare": 0.0, "reduce+optimizer": 0.0, "steps": 0}

    def prepared_step(self, i, from_host):
        """net.prepare(): voxelisation + index tables of ALL samples of step i on a side stream (the reference does
        its voxelisation ahead of time in DataLoader workers).  From host: the pinned scans are copied inside prepare()."""
        src = self.host if from_host else self.resident
        pts = []
        for j in range(self.ppg):
            a, b = src[(i * self.ppg + j) % self.pool_n]
            pts += [a, b]
            if from_host:
                self.h2d_bytes += a.numel() * 4 + b.numel() * 4
        ex = {"points": pts, "n_samples": self.ppg, "host_outputs": False}
        if self.prefetcher is not None:
            return self.prefetcher.submit(ex)            # worker thread (rslo_b200/data/prefetch.py)
        return self.net.prepare(ex)

    def step(self, i, from_host):
        """one step = the ppg samples of the batch through ONE net(example) call (example["n_samples"] = ppg):
        loss = mean over the samples; one backward"""
        torch, net = self.torch, self.net
        if self.train:
            self.reducer.zero_()
        ex = self.prefetch.pop((i, from_host), None)
        if ex is None:
            ex = self.prepared_step(i, from_host)
        if hasattr(ex, "result"):
            ex = ex.result()
        ph = self.host_phase
        t0 = time.perf_counter()
        if self.train:
            ret = net(ex)
            t1 = time.perf_counter()
            ret["loss"].sum().backward()
            t2 = time.perf_counter()
            ph["forward"] += t1 - t0
            ph["backward"] += t2 - t1
            res = ret["loss"].detach().reshape(-1)
        else:
            with torch.no_grad():
                ret = net(ex)
            res = torch.cat([ret["translation_preds"], ret["rotation_preds"]], -1).reshape(-1)
        # the next step's samples are prepared while this one's kernels run
        for stale in [k for k in self.prefetch if k[1] != from_host or k[0] <= i]:
            del self.prefetch[stale]
        t3 = time.perf_counter()
        for ahead in range(1, self.PREFETCH_DEPTH + 1):
            if (i + ahead, from_host) not in self.prefetch:
                self.prefetch[(i + ahead, from_host)] = self.prepared_step(i + ahead, from_host)
        t4 = time.perf_counter()
        ph["prepare"] += t4 - t3
        if self.train:
            self.reducer.all_reduce(average=False)  # pack into the flat buffer (+ NCCL all-reduce when N > 1)
            self.opt.step()                         # global-norm clip + 1/world + weight decay + Adam: 2 launches
            ph["reduce+optimizer"] += time.perf_counter() - t4
        ph["steps"] += 1
        if from_host:
            # D2H of the step's result, every step: asynchronous copy into pinned memory, read by the host one step
            # later (while the next step runs), as the scans are copied one step early
            slot = self.result_slots[i % 2]
            if slot[0] is None or slot[0].numel() != res.numel():
                slot[0] = torch.empty(res.numel(), dtype=res.dtype).pin_memory()
            self.read_result(1 - i % 2)
            slot[0].copy_(res, non_blocking=True)
            slot[1] = torch.cuda.Event()
            slot[1].record()
            self.d2h_bytes += res.numel() * 4
        self.limiter.tick()
        return res

    def read_result(self, k):
        """host read of the result copied by an earlier step (waits for its copy)"""
        buf, ev = self.result_slots[k]
        if ev is not None:
            ev.synchronize()
            self.last_result = buf.tolist()
            self.result_slots[k][1] = None

    def presize_pools(self):
        """after the warm-up steps: give every stream's allocator pool its head-room once, so no timed step calls
        cudaMalloc (rslo_b200/utils/memory.py: a cudaMalloc next to a polling nvidia-smi stalls the step 10-100 ms)"""
        from rslo_b200.utils.memory import presize_stream_pools
        torch = self.torch
        presize_stream_pools([torch.cuda.current_stream(self.dev), self.net.__dict__.get("_prep_stream"),
                              getattr(self.reducer, "_comm", None)])

    def timed(self, nsteps, from_host, first):
        """-> (total ms of exactly nsteps steps, max over ranks; per-step ms on this rank)"""
        torch = self.torch
        import torch.distributed as dist
        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(nsteps + 1)]
        evs[0].record()
        for k in self.host_phase:
            self.host_phase[k] = 0
        t_host = time.time()
        for i in range(nsteps):
            self.flush.zero_()                                             # L2 flush between steps
            self.step(first + i, from_host)
            evs[i + 1].record()
        self.host_s = time.time() - t_host                                 # time the training thread needed to ENQUEUE
        for k in (0, 1):                                                   # the last steps' results reach the host
            self.read_result(k)
        torch.cuda.synchronize()
        if self.world > 1:
            dist.barrier()
        ms = evs[0].elapsed_time(evs[-1])
        per = [evs[i].elapsed_time(evs[i + 1]) for i in range(nsteps)]
        if self.world > 1:
            t = torch.tensor([ms], device=self.dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, per

    def close(self):
        self.prefetch.clear()
        if self.prefetcher is not None:
            self.prefetcher.shutdown()
            self.prefetcher = None

    def quick(self, steps, warmup):
        """pairs/s of this workload, device-resident inputs (used for the extras)"""
        for i in range(warmup):
            self.step(i, False)
        self.presize_pools()
        ms, per = self.timed(steps, False, warmup)
        self.close()
        per.sort()
        return {"value": self.world * self.ppg * steps / (ms / 1e3), "unit": "pairs/s", "ms_per_step": ms / steps,
                "ms_per_step_median": per[len(per) // 2], "steps": steps, "warmup": warmup}


def median(xs):
    xs = sorted(xs)
    return xs[len(xs) // 2] if xs else None


def like_for_like_nn(dev):
    """exact NN, 40000 x 40000 voxel means of a synthetic pair: own kernel (csrc/nn.cu) next to the reference's own
    CUDA kernel compiled unmodified for sm_100a (oracle/_ref/cd_ref.so), same box, same inputs, results compared."""
    import torch
    from rslo_b200 import kernels as K
    from rslo_b200.data import synthetic
    import rslo_b200
    so = os.path.join(ROOT, "oracle", "_ref", "cd_ref.so")
    _, vg = rslo_b200.build_network(testing=False, seed=7)
    a, b, _ = synthetic.make_pair(0)
    pts = []
    for s in (a, b):
        out = K.voxelize(torch.from_numpy(s).to(dev), vg.voxel_size, vg.point_cloud_range, vg.grid_size, materialize=False)
        n = int(out["n_dev"].item())
        pts.append(out["mean"][:n, :3].contiguous())
    n = min(p.shape[0] for p in pts)
    q, t = pts[0][:n].contiguous(), pts[1][:n].contiguous()

    def timeit(fn, reps):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps * 1e3

    own_us = timeit(lambda: K.nn_exact(q, t), 50)
    d_own, i_own = K.nn_exact(q, t)
    out = {"n": int(n), "own_us": own_us, "reference_cuda_us": None, "bit_identical": None}
    if os.path.exists(so):
        try:
            spec = importlib.util.spec_from_file_location("cd_ref", so)
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            d = torch.zeros(1, n, device=dev)
            i = torch.zeros(1, n, dtype=torch.int32, device=dev)
            q1, t1 = q[None].contiguous(), t[None].contiguous()
            out["reference_cuda_us"] = timeit(lambda: mod.forward_cuda_one_direction(q1, t1, d, i), 5)
            out["bit_identical"] = bool(torch.equal(d[0], d_own) and torch.equal(i[0], i_own))
            out["speedup"] = out["reference_cuda_us"] / own_us
        except Exception as e:                      # the checker is optional on the box
            out["reference_cuda_error"] = str(e)[:200]
    else:
        out["reference_cuda_error"] = "oracle/_ref/cd_ref.so not present"
    return out


def profile_replica(run, W, nprof):
    """CUDA events around every C-ABI call of `nprof` steps (head un-graphed so its kernels are visible)."""
    torch = run.torch
    from rslo_b200 import kernels as K
    from rslo_b200.models import odom_pred
    cls = odom_pred.UNRResNetOdomPredEncDecSVDTempMask
    saved = cls.use_cuda_graph
    cls.use_cuda_graph = False
    def plain_step(i):
        # one stream only (no side-stream preparation): an event pair then brackets exactly the kernels of its call
        if run.train:
            run.reducer.zero_()
        pts = []
        for j in range(run.ppg):
            pts += list(run.resident[(i * run.ppg + j) % run.pool_n])
        ex = {"points": pts, "n_samples": run.ppg, "host_outputs": False}
        if run.train:
            run.net(ex)["loss"].sum().backward()
            run.reducer.all_reduce(average=False)
            run.opt.step()
        else:
            with torch.no_grad():
                run.net(ex)

    try:
        plain_step(W)                        # fills the rulebook-size caches outside the measured calls
        torch.cuda.synchronize()
        K.PROFILE = []
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(nprof):
            run.flush.zero_()
            plain_step(W + 1 + i)
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
    finally:
        K.PROFILE = None
        cls.use_cuda_graph = saved
        run.prefetch.clear()
    return agg, step_ms


TENSOR_KERNELS = ("conv2d_tc", "conv2d_tc_wgrad")


def roofline_of(name, a, nprof, ksum):
    (hbm, hbm_src), (tf32, tf32_src) = measured_peaks()
    common = {"kernel": name, "launches_per_step": a[0] / nprof, "avg_launch_ms": a[1] / a[0],
              "algorithmic_bytes_per_launch": a[2] / a[0], "gflops_per_launch": a[3] / a[0] / 1e9,
              "share_of_profiled_kernel_time": a[1] / nprof / ksum, "traffic": measured_traffic(name)}
    if name in TENSOR_KERNELS:
        ach = a[3] / (a[1] * 1e-3) / 1e12
        common.update({"bound": "tensor", "achieved": ach, "peak": tf32, "unit": "TFLOP/s", "frac": ach / tf32,
                       "peak_source": tf32_src, "executed_tflops": 3 * ach, "executed_frac": 3 * ach / tf32,
                       "note": "achieved = USEFUL FLOPs (2*pixels*taps*Cin*Cout) / time, averaged over all ~100 launches "
                               "of a step (3x3, 1x1, stride 2, data gradients; 10-40 us each); the split-TF32 scheme "
                               "executes 3x that on the tensor pipe (executed_*).  Per-stage trace of the dominant "
                               "layer shape (profiles/r02_conv2d_trace.md): inside a CTA's pipeline the tensor pipe "
                               "retires one M128xN64xK8 MMA every 73-77 cycles against 55 at the measured peak"})
    else:
        ach = a[2] / (a[1] * 1e-3) / 1e9
        common.update({"bound": "hbm", "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm,
                       "peak_source": hbm_src,
                       "useful_tflops": (a[3] / (a[1] * 1e-3) / 1e12) if a[3] else None})
    return common


def main():
    args = parse_args()
    cfg = workload_config(args.workload, args.pairs_per_gpu, args.max_voxels)
    if args.impl == "reference":
        return reference_arm(args, cfg)

    import torch
    import torch.distributed as dist
    from rslo_b200 import kernels as K
    from rslo_b200.utils.distributed import init_from_env

    rank, local, world = init_from_env()
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    run = Runner(cfg, dev, rank, world)
    ppg = run.ppg
    reducer_bytes = run.reducer.nbytes if (run.train and world > 1) else 0

    # warm-up (>= 3 steps), then the timed region
    W = max(args.warmup, 3)
    for i in range(W):
        run.step(i, False)
    run.presize_pools()
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = K.kernel_launch_count()
    cuda_prof = os.environ.get("RSLO_BENCH_CUDA_PROFILER") == "1"      # ncu --profile-from-start off
    if cuda_prof:
        torch.cuda.profiler.start()
    ms, per = run.timed(args.steps, False, W)
    if cuda_prof:
        torch.cuda.profiler.stop()
    launches = K.kernel_launch_count() - l0
    run_train = run.train
    host_ms = 1e3 * run.host_s / args.steps
    host_phase = {k: 1e3 * v / max(run.host_phase["steps"], 1) for k, v in run.host_phase.items() if k != "steps"}
    clocks = sampler.stop() if rank == 0 else None
    value = world * ppg * args.steps / (ms / 1e3)

    # e2e: host buffers -> net.prepare / net(example) -> host result
    for i in range(max(3, W // 2)):
        run.step(i, True)
    run.h2d_bytes = run.d2h_bytes = 0
    ms_e2e, per_e2e = run.timed(args.steps, True, W)
    e2e = {"value": world * ppg * args.steps / (ms_e2e / 1e3), "unit": "pairs/s",
           "h2d_bytes_per_step": run.h2d_bytes // (args.steps + run.PREFETCH_DEPTH),
           "d2h_bytes_per_step": run.d2h_bytes // args.steps,
           "ms_per_step": ms_e2e / args.steps, "ms_per_step_median": median(per_e2e),
           "result_read": "loss copied D2H into pinned memory every step, read by the host one step later; all K "
                          "results are on the host when the timed region ends",
           "host_phase_ms_per_step": {k: 1e3 * v / max(run.host_phase["steps"], 1) for k, v in run.host_phase.items()
                                      if k != "steps"}}

    # rooflines: profiled replica of the timed steps (every rank runs it: its steps contain the collective)
    roofline = roofline_hbm = breakdown = None
    if not args.no_profile:
        nprof = min(args.steps, 5)
        agg, _ = profile_replica(run, W, nprof)
        if rank == 0 and agg:
            ksum = sum(a[1] for a in agg.values()) / nprof
            breakdown = {n: {"calls_per_step": a[0] / nprof, "ms_per_step": a[1] / nprof,
                             "share_of_profiled_kernel_time": a[1] / nprof / ksum,
                             "algorithmic_GBps": a[2] / (a[1] * 1e-3) / 1e9 if a[1] > 0 else None,
                             "useful_TFLOPs": a[3] / (a[1] * 1e-3) / 1e12 if (a[1] > 0 and a[3]) else None}
                         for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1])}
            breakdown["_profiled_kernel_ms_per_step"] = ksum
            breakdown["_note"] = ("CUDA events around every C-ABI call of a replica of the timed steps with the head "
                                  "un-graphed; the fused elementwise kernels of the head (BN/ReLU, tails) are not bracketed")
            nm, a = max(agg.items(), key=lambda kv: kv[1][1])
            roofline = roofline_of(nm, a, nprof, ksum)
            hb = [(n, a) for n, a in agg.items() if n not in TENSOR_KERNELS]
            if hb:
                nm2, a2 = max(hb, key=lambda kv: kv[1][1])
                roofline_hbm = roofline_of(nm2, a2, nprof, ksum)
    if world > 1:
        dist.barrier()

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_baseline, _, _, _ = cpu_run(cfg, 2, 0, 40.0)

    extra = None
    if rank == 0 and world == 1 and not args.no_extras and args.workload == "train":
        extra = {}
        run.close()
        del run
        torch.cuda.empty_cache()
        try:
            extra["eval"] = dict(Runner(workload_config("eval"), dev, rank, world).quick(50, 10),
                                 config=public_config(workload_config("eval")))
            extra["train_warm"] = dict(Runner(workload_config("warm"), dev, rank, world).quick(30, 8),
                                       config=public_config(workload_config("warm")))
            torch.cuda.empty_cache()
            sweep = {}
            for nv in (40000, 80000, 160000, 250000):
                c = workload_config("stress", max_voxels=nv)
                r = Runner(c, dev, rank, world)
                sweep[str(nv)] = r.quick(8, 4)
                prep = r.net.prepare({"points": list(r.resident[0])})
                sweep[str(nv)]["voxels_per_frame"] = [int(v.shape[0]) for v in prep["_prepared"]["finish"]()["features"]]
                del r, prep
                torch.cuda.empty_cache()
            extra["stress"] = {"config": public_config(workload_config("stress")), "max_voxels_sweep": sweep}
            extra["like_for_like_nn"] = like_for_like_nn(dev)
        except Exception as e:          # the extras never invalidate the headline line
            extra["error"] = f"{type(e).__name__}: {e}"[:400]

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps, "warmup": W,
                "ms_per_step": ms / args.steps, "ms_per_step_median": median(per),
                "ms_per_step_p90_max": [sorted(per)[int(0.9 * (len(per) - 1))], max(per)],
                "host_enqueue_ms_per_step": host_ms, "host_phase_ms_per_step": host_phase, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": b200_config(cfg, world, ppg, reducer_bytes, run_train),
                "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline,
                "roofline_hbm": roofline_hbm, "cpu_baseline": cpu_baseline, "kernel_breakdown": breakdown, "extra": extra}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
