import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rslo_b200 import kernels as K
torch.set_printoptions(linewidth=200)
def probe(o, ci, co, k=0, Kk=2, cin=64, cout=64, n=64):
    nbr = torch.full((n, Kk), -1, dtype=torch.int32)
    nbr[:, k] = torch.arange(n, dtype=torch.int32)
    feat = torch.zeros(n, cin); feat[o, ci] = 1.0
    g = torch.zeros(n, cout); g[o, co] = 1.0
    gw = K.spconv_tc_backward_weight(feat.cuda(), g.cuda(), nbr.cuda(), n, (Kk, cin, cout)).cpu()
    nz = gw.nonzero().tolist()
    print(f"o={o} ci={ci} co={co} k={k} -> nonzeros {nz[:8]} vals {[round(float(gw[tuple(i)]),3) for i in nz[:8]]}  sum={float(gw.sum()):.3f}")
probe(0, 0, 0)
probe(5, 3, 7)
probe(13, 3, 7)
probe(5, 40, 7)
probe(5, 3, 40)
probe(5, 3, 7, k=1)
probe(63, 63, 63, k=1)
# dense ones
n=64; nbr = torch.arange(n, dtype=torch.int32)[:, None].repeat(1, 2)
feat = torch.ones(n, 64); g = torch.ones(n, 64)
gw = K.spconv_tc_backward_weight(feat.cuda(), g.cuda(), nbr.cuda(), n, (2, 64, 64)).cpu()
print("all ones: min/max", float(gw.min()), float(gw.max()), "expect 64")
