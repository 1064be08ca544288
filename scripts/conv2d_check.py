"""Standalone check + timing of the dense tensor-core convolutions (csrc/conv2d_tc.cu) against torch float64.
usage: python scripts/conv2d_check.py fwd|dgrad|wgrad [--time]"""
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from rslo_b200 import kernels as K

SHAPES = [  # B, Cin, Cout, H, W, ks, stride
    (1, 64, 64, 16, 32, 3, 1),
    (2, 256, 128, 96, 176, 3, 2), (2, 256, 128, 96, 176, 1, 2),
    (2, 128, 128, 48, 88, 3, 1), (2, 128, 128, 48, 88, 3, 2), (2, 128, 128, 48, 88, 1, 2),
    (2, 128, 128, 24, 44, 3, 1), (2, 128, 256, 24, 44, 3, 2), (2, 128, 256, 24, 44, 1, 2),
    (2, 256, 256, 12, 22, 3, 1), (2, 512, 128, 24, 44, 3, 1), (2, 256, 64, 48, 88, 3, 1),
    (2, 192, 64, 96, 176, 3, 1), (2, 64, 64, 96, 176, 3, 1), (2, 64, 32, 96, 176, 3, 1),
    (2, 128, 64, 24, 44, 3, 1), (2, 64, 64, 24, 44, 3, 1), (2, 64, 32, 48, 88, 3, 1), (2, 32, 64, 48, 88, 3, 1),
    (1, 256, 256, 12, 22, 3, 1), (3, 128, 128, 24, 44, 3, 1), (1, 192, 64, 96, 176, 3, 1),
]


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def report(name, got, ref):
    ref = ref.double()
    err = (got.double() - ref).abs()
    scale = ref.abs().max().item() + 1e-30
    rel = err.max().item() / scale
    bad = (err > 1e-4 * scale)
    msg = f"{name}: max_err/max_ref = {rel:.3e}  bad={int(bad.sum())}/{bad.numel()}"
    if bad.any():
        idx = bad.nonzero()[:6].tolist()
        msg += f" first_bad={idx} got={[float(got[tuple(i)]) for i in idx[:3]]} ref={[float(ref[tuple(i)]) for i in idx[:3]]}"
    print(msg, flush=True)
    return rel


def main():
    which = sys.argv[1]
    do_time = "--time" in sys.argv
    torch.manual_seed(0)
    dev = torch.device("cuda")
    worst = 0.0
    for (B, Cin, Cout, H, W, ks, st) in SHAPES:
        pad = ks // 2
        x = torch.randn(B, H, W, Cin, device=dev)
        w = torch.randn(Cout, Cin, ks, ks, device=dev) / (Cin * ks * ks) ** 0.5
        bias = torch.randn(Cout, device=dev)
        xs = K.conv2d_split(x)
        Ho, Wo = (H + 2 * pad - ks) // st + 1, (W + 2 * pad - ks) // st + 1
        g = torch.randn(B, Ho, Wo, Cout, device=dev)
        gs = K.conv2d_split(g)
        tag = f"[{which}] B{B} {Cin}->{Cout} {H}x{W} k{ks} s{st}"
        x64 = x.permute(0, 3, 1, 2).double().requires_grad_(True)
        w64 = w.double().requires_grad_(True)
        y64 = F.conv2d(x64, w64, bias.double(), stride=st, padding=pad)
        if which == "fwd":
            img = K.conv2d_tc_prepare(w, 0)
            y = K.conv2d_tc_forward(xs, img, Cout, ks, st, bias=bias)
            worst = max(worst, report(tag, y, y64.permute(0, 2, 3, 1)))
            if do_time:
                us = timeit(lambda: K.conv2d_tc_forward(xs, img, Cout, ks, st, bias=bias))
                with torch.backends.cudnn.flags(enabled=True, allow_tf32=False, benchmark=True):
                    xc = x.permute(0, 3, 1, 2)
                    us_ref = timeit(lambda: F.conv2d(xc, w, bias, stride=st, padding=pad))
                fl = 2 * B * Ho * Wo * ks * ks * Cin * Cout
                print(f"    time {us:.1f} us ({fl / us / 1e6:.1f} TFLOP/s useful)  cudnn-fp32 {us_ref:.1f} us", flush=True)
        else:
            y64.backward(g.permute(0, 3, 1, 2).double())
            if which == "dgrad":
                img_t = K.conv2d_tc_prepare(w, 1)
                dx = K.conv2d_tc_backward_data(gs, img_t, (B, H, W, Cin), ks, st)
                worst = max(worst, report(tag, dx, x64.grad.permute(0, 2, 3, 1)))
                if do_time:
                    us = timeit(lambda: K.conv2d_tc_backward_data(gs, img_t, (B, H, W, Cin), ks, st))
                    print(f"    time {us:.1f} us", flush=True)
            else:
                dw = K.conv2d_tc_backward_weight(xs, gs, ks, st)
                worst = max(worst, report(tag, dw, w64.grad))
                if do_time:
                    us = timeit(lambda: K.conv2d_tc_backward_weight(xs, gs, ks, st))
                    print(f"    time {us:.1f} us", flush=True)
    print(f"[{which}] worst {worst:.3e}", flush=True)


if __name__ == "__main__":
    main()
