"""f-N2: the fused optimizer step (csrc/optim.cu through rslo_b200/torchplus/train/fused_optim.py) against the CPU
oracle of the reference's `clip_grad_norm_` + `OptimWrapper.step()` + Adam + OneCycle (oracle/optim.py, pinned live
against the reference classes in tests/test_cpu_oracle.py).  Tolerance: 2e-6 relative to each parameter's largest
magnitude after several steps (the kernel fuses multiplies into FMAs and multiplies by reciprocal bias corrections)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _params(seed, shapes):
    g = torch.Generator().manual_seed(seed)
    return [torch.randn(s, generator=g) for s in shapes]


@pytest.mark.parametrize("true_wd,scale_by_world", [(True, False), (False, False), (True, True)])
def test_fused_adam_clip_matches_oracle(cuda, true_wd, scale_by_world):
    from oracle import optim as oopt
    from rslo_b200.torchplus.train import FusedAdamClip, OneCycle
    from rslo_b200.utils.distributed import FlatGradAllReducer
    shapes = [(1,), (7,), (3, 3, 3, 16, 32), (8193,), (64, 128, 3, 3), (100001,), (5,)]
    host = _params(1, shapes)
    mod = torch.nn.ParameterList([torch.nn.Parameter(p.clone().cuda()) for p in host])
    red = FlatGradAllReducer(mod)
    opt = FusedAdamClip(red, wd=1e-2, true_wd=true_wd, betas=(0.9, 0.99), eps=1e-8, max_norm=10.0, write_clipped_grad=True)
    sched = OneCycle(opt, 40, 0.8e-3, [0.95, 0.85], 10.0, 0.05)
    world = 4 if scale_by_world else 1
    if scale_by_world:
        red.world = world
    params = [p.clone() for p in host]
    state = oopt.new_state(params)
    order = {id(p): i for i, p in enumerate(mod)}
    g = torch.Generator().manual_seed(9)
    for step in range(4):
        mag = 40.0 if step != 2 else 1e-3               # steps 0,1,3 clip; step 2 stays below max_norm
        grads = [torch.randn(s, generator=g) * mag for s in shapes]
        if step >= 1:
            grads[4] = None                             # a parameter without gradient from step 1 on
        opt.zero_grad()
        for p, gr in zip(mod, grads):
            p.grad = None if gr is None else (gr * world).cuda()       # what the all-reduce SUM would hold
        sched.step(step)
        lr, mom = oopt.one_cycle(step, 40, 0.8e-3, [0.95, 0.85], 10.0, 0.05)
        assert abs(opt.lr - lr) < 1e-12 and abs(opt.mom - mom) < 1e-12
        red.pack()
        red.pending_average = scale_by_world
        norm = opt.clip_grad_norm_(10.0)
        opt.step()
        total, clipped = oopt.clip_grad_norm(grads, 10.0)
        assert abs(float(norm) - float(total)) <= 2e-6 * float(total)
        oopt.adam_step(params, clipped, state, lr, mom, 0.99, 1e-8, wd=1e-2, true_wd=true_wd)
        for i, (p, ref) in enumerate(zip(mod, params)):
            err = float((p.detach().cpu() - ref).abs().max() / ref.abs().max())
            assert err < 2e-6, (step, i, err)
            if grads[i] is not None:                    # clip_grad_norm_ leaves the clipped gradient behind
                gerr = float((p.grad.cpu() - clipped[i]).abs().max() / clipped[i].abs().max())
                assert gerr < 1e-6, (step, i, gerr)
    # moments of the parameter that stopped receiving gradients are frozen at their step-0 values
    off = sum(int(np.prod(s)) for s in shapes[:4])
    n4 = int(np.prod(shapes[4]))
    assert torch.allclose(opt.exp_avg[off:off + n4].cpu(), state["m"][4].reshape(-1), rtol=1e-5, atol=1e-7)


def test_grad_norm_is_bit_reproducible_and_exact(cuda):
    from rslo_b200 import kernels as K
    g = torch.Generator().manual_seed(3)
    flat = (torch.randn(12_000_003, generator=g) * 3).cuda()
    ws = torch.zeros(K.grad_norm_workspace_bytes(), dtype=torch.uint8, device="cuda")
    outs = []
    for _ in range(3):
        out = torch.zeros(1, dtype=torch.float64, device="cuda")
        K.grad_sumsq(flat, out, ws)
        outs.append(float(out))
    assert outs[0] == outs[1] == outs[2]
    ref = float((flat.double() ** 2).sum())
    assert abs(outs[0] - ref) <= 1e-12 * ref
    assert int(ws[:4].view(torch.int32)) == 0           # counter left zeroed


def test_optimizer_step_changes_sparse_conv_outputs(cuda):
    """the kernel writes weights through raw pointers (no version bump) - cached split-TF32 weight images
    of the sparse layers must not survive an optimizer step."""
    import rslo_b200
    from rslo_b200.data import synthetic
    from rslo_b200.torchplus.train import FusedAdamClip
    from rslo_b200.utils.distributed import FlatGradAllReducer
    from rslo_b200.utils.weights import deterministic_fill
    net, _ = rslo_b200.build_network(testing=False, seed=7)
    deterministic_fill(net, 11)
    net = net.cuda()
    net.global_step.fill_(2000)
    net._step_host = None
    a, b, _ = synthetic.make_pair(0, n_beams=16, n_az=600)
    ex = {"points": [torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()], "host_outputs": False}
    red = FlatGradAllReducer(net)
    opt = FusedAdamClip(red, lr=1e-3, wd=1e-5, max_norm=10.0)
    net.train()
    red.zero_()
    r0 = net(dict(ex))
    r0["loss"].sum().backward()
    red.all_reduce()
    conv = next(m for m in net.modules() if hasattr(m, "_images") and m.weight.shape[-2] >= 32)   # a tensor-core layer
    w0 = conv.weight.detach().clone()
    opt.step()
    assert float((conv.weight - w0).abs().max()) > 0
    net.eval()                                           # eval forward: no per-step invalidation, relies on step()'s
    with torch.no_grad():
        e1 = net(dict(ex))
        t1 = torch.cat([e1["translation_preds"], e1["rotation_preds"]], -1).clone()
        from rslo_b200.layers.sparse3d import invalidate_weight_images
        invalidate_weight_images()
        e2 = net(dict(ex))
        t2 = torch.cat([e2["translation_preds"], e2["rotation_preds"]], -1)
    assert torch.equal(t1, t2)
