#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/pytest_all.log; tail -2 gpurun_out/pytest_all.log
timeout 900 python bench.py > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err; cut -c1-600 gpurun_out/bench_train.json
timeout 600 python bench.py --workload eval --no-cpu-baseline > gpurun_out/bench_eval.json 2> gpurun_out/bench_eval.err; cut -c1-300 gpurun_out/bench_eval.json
timeout 300 python - <<'PY'
import importlib.util, os, torch, sys
sys.path.insert(0, ".")
from rslo_b200 import kernels as K
from rslo_b200.data import synthetic
so = os.path.join("oracle", "_ref", "cd_ref.so")
vs, rg, grid = [0.1, 0.1, 0.2], [-70.4, -38.4, -3, 70.4, 38.4, 5], [1408, 768, 40]
a, b, _ = synthetic.make_pair(0)
ma = K.voxelize(torch.from_numpy(a).cuda(), vs, rg, grid, materialize=False)["mean"][:, :3].contiguous()
mb = K.voxelize(torch.from_numpy(b).cuda(), vs, rg, grid, materialize=False)["mean"][:, :3].contiguous()
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
print("ours nn_exact 40000x40000: %.1f us" % timeit(lambda: K.nn_exact(ma, mb)))
if os.path.exists(so):
    spec = importlib.util.spec_from_file_location("cd_ref", so); mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
    d = torch.zeros(1, ma.shape[0], device="cuda"); i = torch.zeros(1, ma.shape[0], dtype=torch.int32, device="cuda")
    q, t = ma[None].contiguous(), mb[None].contiguous()
    print("reference ChamferDistanceKernel (unmodified, sm_100a) 40000x40000: %.1f us" % timeit(lambda: mod.forward_cuda_one_direction(q, t, d, i), 5))
PY
