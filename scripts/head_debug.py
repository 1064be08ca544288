"""debug: per-parameter gradient error of the own trunk vs float64 torch, in forward order"""
import sys, os, copy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import test_gpu_head as T
def l2(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))
T._rel = l2

S = int(sys.argv[1]) if len(sys.argv) > 1 else 2
mode = sys.argv[2] if len(sys.argv) > 2 else "train"
import rslo_b200
from rslo_b200.utils.weights import deterministic_fill
net, _ = rslo_b200.build_network(testing=False, seed=7)
deterministic_fill(net, 11)
head = net.odom_predictor.cuda()
head.train(mode != "eval")
if mode == "frozen":
    for m in head.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.eval()
x1, x2 = T._inputs(S, 5 + S)
st0 = copy.deepcopy(head.state_dict())
h64, ref, ref_mask, ref_x = T._ref_trunk(head, x1, x2, S)
head.load_state_dict(st0)
outs, mask, (a, b) = T._own(head, x1, x2, 1)
for n, o, r in zip(["tq", "tl", "rl", "py0", "py1"], outs, ref):
    print("out", n, T._rel(o, r))
which = sys.argv[3] if len(sys.argv) > 3 else "all"
sel = {"all": [0, 1, 2, 3, 4], "tq": [0], "tl": [1], "py0": [3], "py1": [4]}[which]
T._loss([outs[i] for i in sel], 1).backward()
T._loss([ref[i] for i in sel], 1).backward()
gx1 = torch.cat([p[0].grad for p in ref_x]); gx2 = torch.cat([p[1].grad for p in ref_x])
print("gx1", T._rel(a.grad, gx1), "gx2", T._rel(b.grad, gx2))
p64 = dict(h64.named_parameters())
# the same through torch / cuDNN in FP32 (what the reference runs): its error vs float64 is the yardstick
stash = {k: head.__dict__.pop(k) for k in ("_trunk_engine", "_graphed") if k in head.__dict__}
h32 = copy.deepcopy(head)
head.__dict__.update(stash)
h32.load_state_dict(st0)
h32.zero_grad()
o32 = []
x32 = []
with torch.backends.cudnn.flags(enabled=True, allow_tf32=False, benchmark=False):
    for s_ in range(S):
        aa = x1[s_:s_ + 1].detach().requires_grad_(True); bb = x2[s_:s_ + 1].detach().requires_grad_(True)
        x32.append((aa, bb))
        tq, tl, rl, py, _m = h32._trunk_torch(aa, bb)
        o32.append([tq, tl, rl] + py)
    o32 = [torch.cat(v) for v in zip(*o32)]
    T._loss([o32[i] for i in sel], 1).backward()
print("torch-fp32 gx1", T._rel(torch.cat([p[0].grad for p in x32]), gx1))
p32 = dict(h32.named_parameters())
for k, p in head.named_parameters():
    r = p64[k].grad
    if r is None or p.grad is None:
        continue
    e = T._rel(p.grad, r)
    e32 = T._rel(p32[k].grad, r) if p32[k].grad is not None else -1
    flag = "  <<<<" if e > 3 * e32 and e > 1e-5 else ""
    print(f"{k:60s} own {e:.3e} torch32 {e32:.3e} max {float(r.abs().max()):.3e}{flag}")
