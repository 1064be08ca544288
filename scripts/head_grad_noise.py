"""Diagnostic: per-parameter gradient error of the head trunk (own kernels) vs float64, next to torch-FP32's error.
usage (GPU box): python scripts/head_grad_noise.py [train|frozen] [S]"""
import copy
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_gpu_head as T  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "frozen"
S = int(sys.argv[2]) if len(sys.argv) > 2 else 2
head = T.head.__wrapped__(None) if hasattr(T.head, "__wrapped__") else None
if head is None:
    import rslo_b200
    from rslo_b200.utils.weights import deterministic_fill
    net, _ = rslo_b200.build_network(testing=False, seed=7)
    deterministic_fill(net, 11)
    head = net.odom_predictor.cuda()
    g = torch.Generator(device="cpu").manual_seed(3)
    for m in head.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.weight.data = (0.5 + torch.rand(m.num_features, generator=g)).cuda()
            m.bias.data = (0.2 * torch.randn(m.num_features, generator=g)).cuda()
            m.running_mean.data = (0.1 * torch.randn(m.num_features, generator=g)).cuda()
            m.running_var.data = (0.5 + torch.rand(m.num_features, generator=g)).cuda()
head.train(True)
if mode == "frozen":
    for m in head.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.eval()
x1, x2 = T._inputs(S, 5 + S)
state0 = copy.deepcopy(head.state_dict())
h64, ref, ref_mask, ref_x = T._ref_trunk(head, x1, x2, S)
head.load_state_dict(state0)
outs, mask, (a, b) = T._own(head, x1, x2, 1)
for n, o, r in zip(["tq_map", "t_logit", "r_logit", "py0", "py1"], outs, ref):
    print(f"fwd {n}: rel {T._rel(o, r):.3e}")
T._loss(outs, 1).backward()
T._loss(ref, 1).backward()
h32, g32x1, g32x2 = T._torch32_grads(head, state0, x1, x2, S, 1)
gx1 = torch.cat([p[0].grad for p in ref_x])
print(f"dx1: own {T._l2(a.grad, gx1):.3e} torch32 {T._l2(g32x1, gx1):.3e}")
p64, p32 = dict(h64.named_parameters()), dict(h32.named_parameters())
rows = []
for k, p in head.named_parameters():
    r = p64[k].grad
    if r is None or p.grad is None or float(r.abs().max()) < 1e-9:
        continue
    rows.append((T._l2(p.grad, r), T._l2(p32[k].grad, r), k))
for e, e32, k in rows:
    flag = " <<<" if e > max(8 * e32, 5e-5) else ""
    print(f"{k:44s} own {e:.3e}  torch32 {e32:.3e}  ratio {e / max(e32, 1e-30):7.1f}{flag}")

# ---- "seq" mode: reproduce the test order (train-2, train-1, frozen-2 on ONE head) and isolate each convolution's
# weight-gradient kernel error from the upstream error (float64 wgrad of the kernel's own operands)
if len(sys.argv) > 3 and sys.argv[3] == "seq":
    from rslo_b200 import kernels as K
    import torch.nn.functional as F

    def one(mode, S, check):
        head.train(True)
        if mode == "frozen":
            for m in head.modules():
                if isinstance(m, torch.nn.BatchNorm2d):
                    m.eval()
        x1, x2 = T._inputs(S, 5 + S)
        state0 = copy.deepcopy(head.state_dict())
        y64s, hooks, gy64, gy_own = [], [], {}, {}
        names = {m.weight.data_ptr(): n for n, m in head.named_modules() if isinstance(m, torch.nn.BatchNorm2d)}
        if check:
            stash = {k: head.__dict__.pop(k) for k in ("_trunk_engine", "_graphed") if k in head.__dict__}
            for m in head.modules():
                if isinstance(m, torch.nn.BatchNorm2d):
                    def hk(mod, inp, out, ptr_=m.weight.data_ptr()):
                        t = inp[0][0] if isinstance(inp[0], (list, tuple)) else inp[0]
                        if t.dtype == torch.float64:
                            y64s.append(t.detach())
                            if t.requires_grad:
                                slot = gy64.setdefault(ptr_, [])
                                slot.append(None)
                                t.register_hook(lambda g_, slot=slot, i=len(slot) - 1: slot.__setitem__(i, g_.detach()))
                    hooks.append(m.register_forward_hook(hk))
            head.__dict__.update(stash)
        h64, ref, _, ref_x = T._ref_trunk(head, x1, x2, S)      # deepcopy carries the hooks along
        for h_ in hooks:
            h_.remove()
        head.load_state_dict(state0)
        calls = []
        ys_own = []
        orig_bf = K.bn_act_forward

        def spy_bf(y, *a, **k):
            ys_own.append(y)
            return orig_bf(y, *a, **k)
        import rslo_b200.layers.head_tc as HT0
        HT0.K.bn_act_forward = spy_bf
        orig = K.conv2d_tc_backward_weight

        def spy(x_split, g_split, ksize, stride, **kw):
            calls.append((x_split, g_split, ksize, stride, kw.get("scratch"), kw.get("cout_real")))
            return orig(x_split, g_split, ksize, stride, **kw)
        import rslo_b200.layers.head_tc as HT
        HT.K.conv2d_tc_backward_weight = spy
        orig_bn, orig_dg = K.bn_act_backward, K.conv2d_tc_backward_data
        bn_i = [0]

        def spy_bn(dz, z, y, ipg, mean_rstd, gamma, relu, batch_stats, sums, g_split, dres, dres_acc, dgamma, dbeta, dbias):
            orig_bn(dz, z, y, ipg, mean_rstd, gamma, relu, batch_stats, sums, g_split, dres, dres_acc, dgamma, dbeta, dbias)
            gy_own[gamma.data_ptr()] = g_split
            if check:
                d = dz.double() * ((z > 0).double() if relu else 1.0)
                Bq, Hq, Wq, Cq = y.shape
                G = Bq // ipg
                mr = mean_rstd.double()
                mean = mr[:, :, 0].repeat_interleave(ipg, 0).view(Bq, 1, 1, Cq)
                rstd = mr[:, :, 1].repeat_interleave(ipg, 0).view(Bq, 1, 1, Cq)
                xhat = (y.double() - mean) * rstd
                db, dg_ = d.sum((0, 1, 2)), (d * xhat).sum((0, 1, 2))
                gy = gamma.double().view(1, 1, 1, Cq) * rstd * d if not batch_stats else None
                l2 = lambda a, b: float((a.double() - b).norm() / b.norm().clamp_min(1e-300))
                msg = f"[{mode}-{S}] bn_bwd {bn_i[0]:2d} C{Cq} {Hq}x{Wq}: dbeta {l2(dbeta, db):.2e} dgamma {l2(dgamma, dg_):.2e}"
                if gy is not None:
                    msg += f" gy {l2(g_split[0].double() + g_split[1].double(), gy):.2e}"
                msg += f" |dz|max {float(dz.abs().max()):.2e} mean(d)/mean|d| {float(d.mean() / d.abs().mean().clamp_min(1e-300)):.2e}"
                print(msg)
            bn_i[0] += 1

        def spy_dg(g_split, image_t, in_shape, ksize, stride, out=None, accumulate=False):
            base = out.double().clone() if (accumulate and check) else None
            r = orig_dg(g_split, image_t, in_shape, ksize, stride, out=out, accumulate=accumulate)
            if check:
                Bq, Hq, Wq, cin = in_shape
                coutp = g_split.shape[-1]
                img = image_t.view(2, ksize * ksize, cin, coutp).double()
                w = (img[0] + img[1]).permute(2, 1, 0).reshape(coutp, cin, ksize, ksize)     # [co][ci][taps]
                g = (g_split[0].double() + g_split[1].double()).permute(0, 3, 1, 2)
                ref = torch.nn.grad.conv2d_input((Bq, cin, Hq, Wq), w, g, stride=stride, padding=ksize // 2).permute(0, 2, 3, 1)
                got = r.double() - (base if base is not None else 0)
                print(f"[{mode}-{S}] dgrad {coutp}->{cin} k{ksize} s{stride} {Hq}x{Wq} acc={int(accumulate)}: rel-L2 "
                      f"{float((got - ref).norm() / ref.norm().clamp_min(1e-300)):.2e}  mean err/mean|ref| "
                      f"{float((got - ref).mean() / ref.abs().mean().clamp_min(1e-300)):.2e}")
            return r
        HT.K.bn_act_backward = spy_bn
        HT.K.conv2d_tc_backward_data = spy_dg
        outs, mask, (a, b) = T._own(head, x1, x2, 1)
        T._loss(outs, 1).backward()
        HT.K.conv2d_tc_backward_weight = orig
        HT.K.bn_act_backward, HT.K.conv2d_tc_backward_data = orig_bn, orig_dg
        HT.K.bn_act_forward = orig_bf
        if check:
            # h64 ran sample by sample (S calls per BN); stitch its BN inputs back together
            per = len(y64s) // S
            for i, yo in enumerate(ys_own[:per]):
                yr = torch.cat([y64s[s_ * per + i] for s_ in range(S)]).permute(0, 2, 3, 1)
                err = (yo.double() - yr).abs()
                sd = yr.std(dim=(0, 1, 2)).clamp_min(1e-300)
                print(f"[{mode}-{S}] fwd conv->bn {i:2d} C{yo.shape[-1]} {yo.shape[1]}x{yo.shape[2]}: max err/max|y| {float(err.max() / yr.abs().max()):.2e}  "
                      f"worst channel max err/std {float((err.amax(dim=(0, 1, 2)) / sd).max()):.2e}  |y|max {float(yr.abs().max()):.2e} min std {float(sd.min()):.2e}")
        T._loss(ref, 1).backward()
        if check:
            for ptr_, gs_ in gy_own.items():
                if ptr_ in gy64 and all(t is not None for t in gy64[ptr_]):
                    r = torch.cat(gy64[ptr_]).permute(0, 2, 3, 1)
                    o = gs_[0].double() + gs_[1].double()
                    print(f"[{mode}-{S}] gy vs float64 net {names[ptr_]:36s}: rel-L2 {float((o - r).norm() / r.norm().clamp_min(1e-300)):.2e}  "
                          f"mean err/mean|ref| {float((o - r).mean() / r.abs().mean().clamp_min(1e-300)):.2e}")
            h32, _, _ = T._torch32_grads(head, state0, x1, x2, S, 1)
            p64, p32 = dict(h64.named_parameters()), dict(h32.named_parameters())
            for k, p in head.named_parameters():
                r = p64[k].grad
                if r is None or p.grad is None or float(r.abs().max()) < 1e-9:
                    continue
                e, e32 = T._l2(p.grad, r), T._l2(p32[k].grad, r)
                if e > max(8 * e32, 5e-5):
                    print(f"[{mode}-{S}] {k:44s} own {e:.3e} torch32 {e32:.3e} <<<")
            for i, (xs_, gs_, ks, st, scr, cr) in enumerate(calls):
                x = (xs_[0].double() + xs_[1].double()).permute(0, 3, 1, 2)
                g = (gs_[0].double() + gs_[1].double()).permute(0, 3, 1, 2)
                cin, coutp = x.shape[1], g.shape[1]
                wref = torch.nn.grad.conv2d_weight(x, (coutp, cin, ks, ks), g, stride=st, padding=ks // 2)
                got = scr.view(ks * ks, cin, coutp).permute(2, 1, 0).reshape(coutp, cin, ks, ks).double()
                err = float((got - wref).norm() / wref.norm().clamp_min(1e-300))
                amp = float((x.abs().mean() * g.abs().mean() * x.shape[0] * x.shape[2] * x.shape[3]) / wref.abs().mean().clamp_min(1e-300))
                print(f"[{mode}-{S}] wgrad call {i:2d} {cin:3d}->{coutp:3d} k{ks} s{st} {tuple(x.shape[2:])}: kernel rel-L2 err {err:.3e}  "
                      f"cancellation {amp:.1f}  |x|max {float(x.abs().max()):.2e} |g|max {float(g.abs().max()):.2e}")
        head.zero_grad()

    print("==== sequence ====")
    one("train", 2, False)
    one("train", 1, False)
    one("frozen", 2, True)
