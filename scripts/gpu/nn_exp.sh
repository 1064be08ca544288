#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "nn_bit_exact or stress_grid" 2>&1 | tail -3
timeout 300 python - <<'PY'
import torch, sys
sys.path.insert(0, ".")
from rslo_b200 import kernels as K
from rslo_b200.data import synthetic
vs, rg, grid = [0.1, 0.1, 0.2], [-70.4, -38.4, -3, 70.4, 38.4, 5], [1408, 768, 40]
a, b, _ = synthetic.make_pair(0)
ma = K.voxelize(torch.from_numpy(a).cuda(), vs, rg, grid, materialize=False)["mean"][:, :3].contiguous()
mb = K.voxelize(torch.from_numpy(b).cuda(), vs, rg, grid, materialize=False)["mean"][:, :3].contiguous()
for _ in range(3): K.nn_exact(ma, mb)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): K.nn_exact(ma, mb)
e1.record(); torch.cuda.synchronize()
print("nn_exact 40000x40000 us:", e0.elapsed_time(e1) / 20 * 1e3)
PY
