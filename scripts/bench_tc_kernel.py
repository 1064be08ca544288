"""Time the tensor-core sparse-conv kernel alone on a realistic level (two frames' 64-channel level 2)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rslo_b200 import kernels as K
from rslo_b200.data import synthetic
from rslo_b200.models import middle
vs, rg, grid = [0.1, 0.1, 0.2], [-70.4, -38.4, -3, 70.4, 38.4, 5], [1408, 768, 40]
frames = []
for s in (0, 1):
    pts = torch.from_numpy(synthetic.make_pair(s)[0]).cuda()
    out = K.voxelize(pts, vs, rg, grid, materialize=False, with_table=True)
    frames.append((out["coordinates"], 40000, out["table"], out["n_dev"]))
entries, meta = middle.build_tables_batched(frames, [41, 768, 1408])
for key, cin, cout in (("subm2", 64, 64), ("subm1", 32, 32), ("subm3", 64, 64)):
    e = entries[key]
    feat = torch.randn(e.n_in, cin, device="cuda")
    w = torch.randn(27, cin, cout, device="cuda") * 0.1
    img = K.spconv_tc_prepare(w)
    for _ in range(5):
        K.spconv_tc_forward(feat, e.nbr, e.n_out, img, cin, cout)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(50):
        K.spconv_tc_forward(feat, e.nbr, e.n_out, img, cin, cout)
    ev[1].record(); torch.cuda.synchronize()
    R = int((e.nbr[:e.n_out] >= 0).sum())
    print(f"DIAG={os.environ.get('TC_DIAG','0')} {key} rows={e.n_out} R={R} {cin}->{cout}: {ev[0].elapsed_time(ev[1]) / 50 * 1e3:.1f} us")
    g = torch.randn(e.n_out, cout, device="cuda")
    for _ in range(3):
        K.spconv_tc_backward_weight(feat, g, e.nbr, e.n_out, (27, cin, cout))
    torch.cuda.synchronize()
    ev[0].record()
    for _ in range(50):
        K.spconv_tc_backward_weight(feat, g, e.nbr, e.n_out, (27, cin, cout))
    ev[1].record(); torch.cuda.synchronize()
    print(f"   wgrad {key} {cin}->{cout}: {ev[0].elapsed_time(ev[1]) / 50 * 1e3:.1f} us")
