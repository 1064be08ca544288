"""Per-CTA / per-stage timeline of k_conv2d_tc (build with `make -C rslo_b200/csrc EXTRA=-DTC_TRACE`) on the head's
dominant layer shape: 3x3 stride 1, 128 -> 128 channels, 2 samples of 48 x 88 pixels."""
import ctypes, os, sys, torch
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rslo_b200 import kernels as K
from rslo_b200 import _lib
lib = _lib.lib
lib.rslo_debug_cv_trace.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
for (B, h, w, cin, cout, ks, st) in ((2, 48, 88, 128, 128, 3, 1), (2, 96, 176, 64, 64, 3, 1), (2, 24, 44, 256, 256, 3, 1)):
    x = torch.randn(B, h, w, cin, device="cuda")
    wt = torch.randn(cout, cin, ks, ks, device="cuda") * 0.05
    xs = K.conv2d_split(x)
    img = K.conv2d_tc_prepare(wt, 0, cout_padded=cout)
    for _ in range(4):                       # includes the one-off variant measurement
        y = K.conv2d_tc_forward(xs, img, cout, ks, st)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(30):
        K.conv2d_tc_forward(xs, img, cout, ks, st)
    ev[1].record(); torch.cuda.synchronize()
    lib.rslo_debug_cv_clear()
    K.conv2d_tc_forward(xs, img, cout, ks, st)
    torch.cuda.synchronize()
    ct = np.zeros(8 * 2048, dtype=np.uint64); sp = np.zeros(6 * 64, dtype=np.uint64)
    lib.rslo_debug_cv_trace(ct.ctypes.data, sp.ctypes.data)
    t = ct.reshape(-1, 8).astype(np.int64); t = t[t[:, 0] > 0]
    sp = sp.reshape(6, 64).astype(np.int64)
    base = t[:, 0].min()
    def us(x): return np.round(np.percentile(x / 1e3, [10, 50, 90]), 1)
    print(f"== conv {cin}->{cout} {ks}x{ks}/{st} on {B}x{h}x{w}: {ev[0].elapsed_time(ev[1]) / 30 * 1e3:.1f} us/launch; {len(t)} CTAs, "
          f"span {(t[:, 5].max() - base) / 1e3:.1f} us, stages/CTA {np.percentile(t[:, 6], [10, 50, 90])}")
    print("   CTA start (us)", us(t[:, 0] - base), " prologue", us(t[:, 1] - t[:, 0]), " pipeline", us(t[:, 4] - t[:, 1]),
          " per stage (ns)", np.round(np.percentile((t[:, 4] - t[:, 1]) / np.maximum(t[:, 6], 1), [10, 50, 90])),
          " epilogue", us(t[:, 5] - t[:, 4]), " CTA total", us(t[:, 5] - t[:, 0]))
    n = int(t[min(5, len(t) - 1), 6])
    t0 = sp[0, 0]
    print("   CTA 5, ns: stage | stage free (TMA issue) | acc buffer free, operands landed, MMAs issued | MMAs retired, drained")
    for i in range(min(n, 14)):
        print(f"   {i:3d} | {sp[0, i] - t0:6d} | {sp[1, i] - t0:6d} {sp[2, i] - t0:6d} {sp[3, i] - t0:6d} | {sp[4, i] - t0:6d} {sp[5, i] - t0:6d}")
