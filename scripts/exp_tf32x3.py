"""Experiment: accuracy and speed of a conv2d computed as 3 TF32 cuDNN convolutions on pre-split operands."""
import time, torch, torch.nn.functional as F
torch.manual_seed(0)
dev = "cuda"
def rn_tf32(x):
    return ((x.view(torch.int32) + 0x1000) & ~0x1fff).view(torch.float32)
def split(x):
    hi = rn_tf32(x); return hi, x - hi
def conv_x3(x, w, stride=1, padding=1):
    xh, xl = split(x); wh, wl = split(w)
    with torch.backends.cudnn.flags(enabled=True, allow_tf32=True, benchmark=True):
        return F.conv2d(xl, wh, None, stride, padding) + F.conv2d(xh, wl, None, stride, padding) + F.conv2d(xh, wh, None, stride, padding)
def bench(fn, n=30):
    for _ in range(5): fn()
    torch.cuda.synchronize(); t = time.time()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.time() - t) / n * 1e6
for (cin, cout, h, w_, s) in [(256, 128, 96, 176, 2), (128, 128, 48, 88, 1), (256, 256, 12, 22, 1), (512, 128, 24, 44, 1), (64, 64, 96, 176, 1)]:
    x = torch.randn(1, cin, h, w_, device=dev); w = torch.randn(cout, cin, 3, 3, device=dev) * (2.0 / (cin * 9)) ** 0.5
    ref = F.conv2d(x.double(), w.double(), None, s, 1)
    with torch.backends.cudnn.flags(enabled=True, allow_tf32=False, benchmark=True):
        y32 = F.conv2d(x, w, None, s, 1); t32 = bench(lambda: F.conv2d(x, w, None, s, 1))
    with torch.backends.cudnn.flags(enabled=True, allow_tf32=True, benchmark=True):
        ytf = F.conv2d(x, w, None, s, 1); ttf = bench(lambda: F.conv2d(x, w, None, s, 1))
    y3 = conv_x3(x, w, s, 1); t3 = bench(lambda: conv_x3(x, w, s, 1))
    sc = ref.abs().max().item()
    e = lambda y: ((y.double() - ref).abs().max().item() / sc, ((y.double() - ref).mean().item()) / sc)
    print(f"cin={cin} cout={cout} {h}x{w_} s={s}: fp32 err {e(y32)[0]:.2e} ({t32:.0f}us) | tf32 err {e(ytf)[0]:.2e} ({ttf:.0f}us) | 3xtf32 err {e(y3)[0]:.2e} bias {e(y3)[1]:.1e} ({t3:.0f}us incl. split)")
