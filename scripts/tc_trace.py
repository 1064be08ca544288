"""Per-CTA phase timeline of k_spconv_tc (build with `make -C rslo_b200/csrc EXTRA=-DTC_TRACE`): where a CTA's time goes
- prologue (neighbour tile, barriers, TMEM), the pipeline steps, the epilogue - and how the CTAs of a launch overlap."""
import ctypes, os, sys, torch
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rslo_b200 import kernels as K
from rslo_b200 import _lib
from rslo_b200.data import synthetic
from rslo_b200.models import middle
vs, rg, grid = [0.1, 0.1, 0.2], [-70.4, -38.4, -3, 70.4, 38.4, 5], [1408, 768, 40]
frames = []
for s in (0, 1):
    pts = torch.from_numpy(synthetic.make_pair(s)[0]).cuda()
    out = K.voxelize(pts, vs, rg, grid, materialize=False, with_table=True)
    frames.append((out["coordinates"], 40000, out["table"], out["n_dev"]))
entries, meta = middle.build_tables_batched(frames, [41, 768, 1408])
lib = _lib.lib
lib.rslo_debug_tc_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
lib.rslo_debug_tc_steps.argtypes = [ctypes.c_void_p]
for key, cin, cout in (("subm2", 64, 64), ("subm1", 32, 32)):
    e = entries[key]
    feat = torch.randn(e.n_in, cin, device="cuda")
    w = torch.randn(27, cin, cout, device="cuda") * 0.1
    img = K.spconv_tc_prepare(w)
    for _ in range(3):
        K.spconv_tc_forward(feat, e.nbr, e.n_out, img, cin, cout)
    torch.cuda.synchronize()
    buf = np.zeros(8 * 4096, dtype=np.uint64)
    lib.rslo_debug_tc_trace(buf.ctypes.data, 0)          # no-op copy (symbol warm)
    K.spconv_tc_forward(feat, e.nbr, e.n_out, img, cin, cout)
    torch.cuda.synchronize()
    lib.rslo_debug_tc_trace(buf.ctypes.data, buf.size)
    t = buf.reshape(-1, 8).astype(np.int64)
    t = t[t[:, 0] > 0]
    base = t[:, 0].min()
    steps = t[:, 6]
    def us(x): return np.round(np.percentile(x / 1e3, [10, 50, 90]), 1)
    print(f"== {key} {cin}->{cout}: {len(t)} CTAs, span {(t[:, 5].max() - base) / 1e3:.1f} us, steps/CTA p10/50/90 {np.percentile(steps, [10, 50, 90])}")
    print("   CTA start after kernel start (us) p10/50/90", us(t[:, 0] - base), " max", round((t[:, 0].max() - base) / 1e3, 1))
    print("   prologue", us(t[:, 1] - t[:, 0]), " producers' loop", us(t[:, 2] - t[:, 1]), " per step (ns)",
          np.round(np.percentile((t[:, 2] - t[:, 1]) / np.maximum(steps, 1), [10, 50, 90])))
    print("   MMA loop end - producers end", us(t[:, 3] - t[:, 2]), " drain end - MMA end", us(t[:, 4] - t[:, 3]),
          " epilogue", us(t[:, 5] - t[:, 4]), " CTA total", us(t[:, 5] - t[:, 0]))
    sb = np.zeros(6 * 64, dtype=np.uint64)
    lib.rslo_debug_tc_steps(sb.ctypes.data)
    sb = sb.reshape(6, 64).astype(np.int64)
    n = int(steps[7]) if len(steps) > 7 else 0
    t0 = sb[0, 0]
    print("   CTA 7, ns since its first step: st | loads issued, stage free, published | MMA saw stage, MMAs issued | offset drained")
    for st in range(min(n, 24)):
        dr = sb[5, st // (cin // 32)] - t0 if st % (cin // 32) == cin // 32 - 1 else -1
        print(f"   {st:3d} | {sb[0, st] - t0:6d} {sb[1, st] - t0:6d} {sb[2, st] - t0:6d} | {sb[3, st] - t0:6d} {sb[4, st] - t0:6d} | {dr:6d}")
