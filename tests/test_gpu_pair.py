"""End-to-end parity of the CUDA hot path (net(example) through the C-ABI kernels) against
  (1) the golden fixtures produced by the REFERENCE network (tests/golden/make_golden.py), and
  (2) the CPU oracle (oracle/net.py) on the same seeded inputs.
Tolerance (north_star): integer outputs bit-exact; pose and loss 1e-4 relative fp32.  Gradients are
checked at 2e-3 of the tensor's max magnitude (fp32 accumulation order over ~1e4..1e5 terms differs
between the atomics-free CUDA reduction and torch's CPU index_add)."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import net as onet

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import make_golden as mg  # noqa: E402

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def net(cuda):
    import rslo_b200
    n, vg = rslo_b200.build_network(testing=True, seed=7)
    return n.cuda(), vg


def _prep(net, name):
    g = np.load(os.path.join(GOLDEN, f"pair_{name}.npz"))
    seed, beams, n_az, T, step, wseed = (int(v) for v in g["meta"])
    onet.fill_weights(net, wseed)
    net.global_step.fill_(step)
    net._step_host = None
    frames = mg.make_frames(seed, beams, n_az, T)
    return g, frames, step


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-12)


@pytest.mark.parametrize("name", ["small_eval", "full_eval"])
def test_eval_matches_reference_golden(net, name):
    net, vg = net
    g, frames, _ = _prep(net, name)
    net.eval()
    with torch.no_grad():
        out = net({"points": [torch.from_numpy(f).cuda() for f in frames]})
    assert [int(v.shape[0]) for v in out["voxel_features"]] == g["n_voxels"].tolist()       # bit-exact counts
    pose = torch.cat([out["translation_preds"], out["rotation_preds"]], -1).cpu().numpy()
    np.testing.assert_allclose(pose, g["pose"], rtol=1e-4, atol=1e-5)
    assert _rel(out["tq_map_g"][:, :, ::8, ::8].cpu().numpy(), g["tq_map_g_sample"]) < 1e-4
    assert _rel(out["t_conf"][:, :, ::8, ::8].cpu().numpy(), g["t_conf_sample"]) < 1e-4
    assert _rel(out["middle_conf_preds"][0][::97].cpu().numpy(), g["cov0_sample"]) < 1e-4
    assert _rel(out["tq_map_g"].abs().sum(dim=(1, 2, 3)).cpu().numpy(), g["tq_map_g_abs_sum"]) < 1e-4


@pytest.mark.parametrize("name", ["small_train", "small_train_warm", "small_train_t3", "full_train"])
def test_train_matches_reference_golden(net, name):
    net, vg = net
    g, frames, _ = _prep(net, name)
    net.train()
    net.zero_grad()
    ret = net({"points": [torch.from_numpy(f).cuda() for f in frames], "host_outputs": False})
    ret["loss"].sum().backward()
    pose = torch.cat([ret["translation_preds"], ret["rotation_preds"]], -1).cpu().numpy()
    np.testing.assert_allclose(pose, g["pose"], rtol=1e-4, atol=1e-5)
    for k in ("translation_loss", "rotation_loss", "pyramid_loss", "C_loss", "loss"):
        np.testing.assert_allclose(ret[k].detach().cpu().numpy().reshape(-1), g[k], rtol=1e-4, atol=1e-5, err_msg=k)
    params = dict(net.named_parameters())
    for k in mg.GRAD_KEYS:
        ref = g["grad:" + k]
        got = params[k].grad
        if ref.size == 0:
            assert got is None
            continue
        assert _rel(mg.grad_sample(got.cpu().numpy()), ref) < 2e-3, k


def test_reference_style_example_matches_points_path(net):
    """The reference's input contract (voxels / num_points / coordinates made by the voxel
    generator) and the fused raw-points path give identical outputs."""
    net, vg = net
    g, frames, _ = _prep(net, "small_eval")
    net.eval()
    ex = {"voxels": [], "num_points": [], "coordinates": [], "num_voxels": []}
    for f in frames:
        r = vg.generate(f, 40000)
        n = len(r["coordinates"])
        assert r["coordinates"].shape == (n, 3) and r["voxels"].shape == (n, 10, 7)
        ex["voxels"].append(torch.from_numpy(r["voxels"]).cuda())
        ex["num_points"].append(torch.from_numpy(r["num_points_per_voxel"]).cuda())
        ex["coordinates"].append(torch.from_numpy(np.concatenate([np.zeros((n, 1), np.int32), r["coordinates"]], 1)).cuda())
        ex["num_voxels"].append(torch.tensor([[n]], dtype=torch.int64))
    with torch.no_grad():
        a = net(ex)
        b = net({"points": [torch.from_numpy(f).cuda() for f in frames]})
    assert torch.equal(a["translation_preds"], b["translation_preds"])
    assert torch.equal(a["rotation_preds"], b["rotation_preds"])
    pose = torch.cat([a["translation_preds"], a["rotation_preds"]], -1).cpu().numpy()
    np.testing.assert_allclose(pose, g["pose"], rtol=1e-4, atol=1e-5)


def test_pair_matches_oracle_other_seed(net):
    """A pair no fixture covers: CUDA path vs the CPU oracle run on the box."""
    net, vg = net
    from rslo_b200.data import synthetic
    a, b, _ = synthetic.make_pair(5, n_beams=24, n_az=500)
    onet.fill_weights(net, 21)
    net.global_step.fill_(5000)
    net._step_host = None
    sd = {k: v.detach().cpu() for k, v in net.state_dict().items()}
    ref = onet.pair_forward(sd, [a, b], training=True, step=5000)
    net.train()
    ret = net({"points": [torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()], "host_outputs": False})
    pose = torch.cat([ret["translation_preds"], ret["rotation_preds"]], -1).cpu().numpy()
    np.testing.assert_allclose(pose, ref["pose"], rtol=1e-4, atol=1e-5)
    for k in ("translation_loss", "rotation_loss", "pyramid_loss", "C_loss", "loss"):
        np.testing.assert_allclose(ret[k].detach().cpu().numpy().reshape(-1), ref[k].reshape(-1), rtol=1e-4,
                                   atol=1e-5, err_msg=k)


def test_loss_stack_on_oracle_inputs(net):
    """Loss stack alone (a10-a14) fed with the oracle's own intermediate tensors, so the NN
    association sees bit-identical inputs: every term within 1e-4 relative, residual pose 1e-5."""
    net, vg = net
    from rslo_b200.data import synthetic
    a, b, _ = synthetic.make_pair(3, n_beams=32, n_az=900)
    onet.fill_weights(net, 31)
    sd = {k: v.detach().cpu() for k, v in net.state_dict().items()}
    ref = onet.pair_forward(sd, [a, b], training=True, step=3000)
    feats, covs = ref["voxel_features"], ref["cov"]
    n = min(f.shape[0] for f in feats)
    from rslo_b200.utils import pose_utils
    q = torch.from_numpy(ref["pose"][:, 3:]).cuda()
    t = torch.from_numpy(ref["pose"][:, :3]).cuda()
    R = pose_utils.quaternion_to_rotation_matrix(torch.roll(q, -1, -1))
    p0, p1 = feats[0][:n].cuda(), feats[1][:n].cuda()
    tgt = p1[None, :, :3] @ R.transpose(1, 2) + t[:, None, :]
    l, res_r, res_t = net._consistency_loss(p0[None, :, :3], tgt, cov_pred=covs[0][:n][None].cuda(),
                                            cov_target=covs[1][:n][None].cuda(), R_pred=R, t_pred=t,
                                            normal_pred=p0[None, :, [4, 5, 6]], normal_target=None, icp_iter=2)
    np.testing.assert_allclose(l.detach().cpu().numpy().reshape(-1), ref["C_loss"].reshape(-1), rtol=1e-4)
    np.testing.assert_allclose(res_r.cpu().numpy(), ref["res_r"], atol=1e-5)
    np.testing.assert_allclose(res_t.cpu().numpy(), ref["res_t"], atol=1e-5)


def test_prepared_example_matches_direct_call(net):
    """net.prepare() (voxelisation + tables on a side stream, ahead of time) changes nothing in the results."""
    net, vg = net
    g, frames, _ = _prep(net, "small_eval")
    net.eval()
    pts = [torch.from_numpy(f).cuda() for f in frames]
    with torch.no_grad():
        a = net({"points": pts})
        ex = net.prepare({"points": pts})
        ex2 = net.prepare({"points": [torch.from_numpy(f).pin_memory() for f in frames]})   # host scans
        b = net(ex)
        c = net(ex2)
    for k in ("translation_preds", "rotation_preds", "tq_map_g"):
        assert torch.equal(a[k], b[k]) and torch.equal(a[k], c[k]), k



def test_two_samples_in_one_call_equal_two_calls(net):
    """example["n_samples"] = 2 (one encoder pass over 4 frames, one head pass over 2 pairs with per-sample
    BatchNorm statistics, one backward) == two single-sample calls with gradient accumulation of loss / 2."""
    net, vg = net
    onet.fill_weights(net, 11)
    net.global_step.fill_(2000)
    net._step_host = None
    net.train()
    fa = mg.make_frames(3, 16, 600, 2)
    fb = mg.make_frames(4, 16, 600, 2)
    sd0 = {k: v.clone() for k, v in net.state_dict().items()}
    net.zero_grad()
    losses, poses = [], []
    for fr in (fa, fb):
        ret = net({"points": [torch.from_numpy(f).cuda() for f in fr], "host_outputs": False})
        (ret["loss"].sum() / 2).backward()
        losses.append(float(ret["loss"].detach().sum()))
        poses.append(torch.cat([ret["translation_preds"], ret["rotation_preds"]], -1))
        # gradients handed out by a CUDA-graph replay alias the graph's static buffers (torch steals them into
        # p.grad when it was None); take ownership before the same graph is replayed again
        for p in net.parameters():
            if p.grad is not None:
                p.grad = p.grad.clone()
    g_ref = {k: p.grad.clone() for k, p in net.named_parameters() if p.grad is not None}
    sd_seq = {k: v.clone() for k, v in net.state_dict().items()}
    net.load_state_dict(sd0)
    net.zero_grad()
    ret = net({"points": [torch.from_numpy(f).cuda() for f in fa + fb], "n_samples": 2, "host_outputs": False})
    ret["loss"].sum().backward()
    np.testing.assert_allclose(float(ret["loss"].sum()), 0.5 * (losses[0] + losses[1]), rtol=2e-5)
    pose = torch.cat([ret["translation_preds"], ret["rotation_preds"]], -1)
    assert _rel(pose.cpu().numpy(), torch.cat(poses).cpu().numpy()) < 1e-5
    worst = []
    for k, p in net.named_parameters():
        if k in g_ref and float(g_ref[k].abs().max()) > 0:
            assert p.grad is not None, k
            a, b = p.grad.double(), g_ref[k].double()
            wk = k[:-5] + ".weight"
            if k.endswith(".bias") and wk in g_ref and float(b.abs().max()) < 1e-4 * float(g_ref[wk].abs().max()):
                continue                       # bias ahead of a batch-statistics BatchNorm: true gradient 0, pure noise
            worst.append((float((a - b).norm() / b.norm().clamp_min(1e-30)), k))
    worst.sort(reverse=True)
    # same arithmetic, different launch geometry (tile / split-K choices depend on the batch): the difference is
    # rounding noise, amplified by the ~35 BatchNorm+ReLU layers of the head (see tests/test_gpu_head.py)
    assert worst[0][0] < 2e-2, worst[:5]
    assert worst[len(worst) // 2][0] < 2e-3, worst[len(worst) // 2]
    # BatchNorm running statistics advanced sample after sample, exactly as in the two calls
    for k, v in net.state_dict().items():
        if "running_" in k or "num_batches_tracked" in k:
            assert _rel(v.double().cpu().numpy(), sd_seq[k].double().cpu().numpy()) < 1e-5, k


def test_stress_grid_end_to_end_matches_oracle(cuda):
    """C5 geometry end to end (0.05 x 0.05 x 0.1 m voxels, grid 2816 x 1536 x 80 -> 4 z-slices -> 256-channel BEV maps,
    512-channel head input, 192 x 352 maps, raised voxel cap) against the CPU oracle: pose and every loss term at 1e-4.
    The scan is reduced (32 beams x 1200 az) so the oracle's single-thread voxeliser / NN finish in seconds; full-size
    C5 inputs are covered by the kernel-level bit-exact tests (test_gpu_kernels.py stress cases)."""
    import rslo_b200
    from rslo_b200.data import synthetic
    cfg = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "rslo_b200", "config", "stress_005.prototxt")
    net, vg = rslo_b200.build_network(cfg, testing=False, seed=7)
    assert vg.grid_size.tolist() == [2816, 1536, 80]
    onet.fill_weights(net, 13)
    net = net.cuda()
    net.global_step.fill_(2000)
    net._step_host = None
    a, b, _ = synthetic.make_pair(21, n_beams=32, n_az=1200)
    sd = {k: v.detach().cpu().clone() for k, v in net.state_dict().items()}
    ref = onet.pair_forward(sd, [a, b], training=True, step=2000, voxel_size=[0.05, 0.05, 0.1], max_voxels=250000)
    net.train()
    ret = net({"points": [torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()], "host_outputs": False})
    pose = torch.cat([ret["translation_preds"], ret["rotation_preds"]], -1).cpu().numpy()
    np.testing.assert_allclose(pose, ref["pose"], rtol=1e-4, atol=1e-5)
    for k in ("translation_loss", "rotation_loss", "pyramid_loss", "C_loss", "loss"):
        np.testing.assert_allclose(ret[k].detach().cpu().numpy().reshape(-1), np.asarray(ref[k]).reshape(-1), rtol=1e-4,
                                   atol=1e-5, err_msg=k)
    ret["loss"].sum().backward()
    assert torch.isfinite(net.odom_predictor.blocks[0][0].conv1.conv1.weight.grad).all()


def test_display_outputs_on_demand_equal_the_eager_ones(net):
    """`feature_mask` / `middle_feature` (display maps of the training log, `voxel_odom_net.py:449-462`) are produced
    eagerly with the reference's host_outputs and on first access with host_outputs=False: same values, same keys."""
    net, vg = net
    onet.fill_weights(net, 11)
    net.global_step.fill_(2000)
    net._step_host = None
    net.train()
    sd0 = {k: v.clone() for k, v in net.state_dict().items()}
    pts = [torch.from_numpy(f).cuda() for f in mg.make_frames(5, 16, 600, 2)]
    a = net({"points": pts})                                    # host_outputs defaults to the reference's behaviour
    net.load_state_dict(sd0)
    b = net({"points": pts, "host_outputs": False})
    assert "feature_mask" in a and not dict.__contains__(b, "feature_mask")
    assert torch.equal(a["feature_mask"], b["feature_mask"].cpu())
    assert dict.__contains__(b, "feature_mask") and b.get("middle_feature") is not None
    for x, y in zip(a["middle_feature"], b["middle_feature"]):
        assert torch.equal(x, y.cpu())
    assert torch.equal(a["loss"].detach().cpu(), b["loss"].detach().cpu())
    net.zero_grad()


def test_encoder_engine_matches_layerwise_path(net):
    """layers/encoder_engine.py (one autograd node over the 25 sparse layers) against the layer-by-layer modules of
    layers/sparse3d.py on the same step: same kernels in the same order -> identical outputs; gradients equal up to
    the summation order of the weight-gradient atomics."""
    from rslo_b200.layers import encoder_engine
    net, vg = net
    onet.fill_weights(net, 11)
    net.global_step.fill_(2000)
    net._step_host = None
    net.train()
    sd0 = {k: v.clone() for k, v in net.state_dict().items()}
    pts = [torch.from_numpy(f).cuda() for f in mg.make_frames(6, 16, 600, 2)]
    res = []
    for use in (False, True):
        encoder_engine.USE_ENGINE = use
        try:
            net.load_state_dict(sd0)
            net.zero_grad()
            ret = net({"points": pts, "host_outputs": False})
            ret["loss"].sum().backward()
        finally:
            encoder_engine.USE_ENGINE = True
        res.append((ret["loss"].detach().clone(), {k: p.grad.clone() for k, p in net.named_parameters() if p.grad is not None},
                    {k: v.clone() for k, v in net.state_dict().items() if "running_" in k or "num_batches" in k}))
    (la, ga, ba), (lb, gb, bb) = res
    assert torch.equal(la, lb)
    assert ga.keys() == gb.keys()
    for k in ga:
        scale = float(ga[k].abs().max())       # (biases ahead of a batch-statistics BatchNorm: true gradient 0, 1e-10 noise)
        assert float((ga[k] - gb[k]).abs().max()) <= 2e-5 * scale + 1e-8, k
    for k in ba:
        assert torch.equal(ba[k], bb[k]), k
    net.zero_grad()


def test_pass_through_generator_layout_is_voxelised_in_forward(net):
    """SURVEY §8b / f-N1: with `voxel_generator.pass_through` the worker-side generate() only packs the raw scan into
    the reference's voxels / coordinates / num_points slots (no CUDA in forked DataLoader workers); forward recognises
    the layout, voxelises on the device and gives the same outputs as the up-front voxelisation, bit for bit."""
    net, vg = net
    g, frames, _ = _prep(net, "small_eval")
    net.eval()
    ex_ref = {"voxels": [], "num_points": [], "coordinates": [], "num_voxels": []}
    ex_pt = {"voxels": [], "num_points": [], "coordinates": [], "num_voxels": []}
    for f in frames:
        for ex, pt in ((ex_ref, False), (ex_pt, True)):
            vg.pass_through = pt
            try:
                r = vg.generate(f, 40000)
            finally:
                vg.pass_through = False
            n = len(r["coordinates"])
            ex["voxels"].append(torch.from_numpy(r["voxels"]).cuda())
            ex["num_points"].append(torch.from_numpy(r["num_points_per_voxel"]).cuda())
            ex["coordinates"].append(torch.from_numpy(np.concatenate([np.zeros((n, 1), np.int32), r["coordinates"]], 1)).cuda())
            ex["num_voxels"].append(torch.tensor([[n]], dtype=torch.int64))
    assert ex_pt["voxels"][0].shape[1:] == (1, 7) and int(ex_pt["coordinates"][0][:, 1:].max()) == -1
    with torch.no_grad():
        a = net(ex_ref)
        b = net(ex_pt)
    for k in ("translation_preds", "rotation_preds", "tq_map_g"):
        assert torch.equal(a[k], b[k]), k
