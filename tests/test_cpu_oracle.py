"""CPU suite (no GPU): the oracle against the golden fixtures made from the REFERENCE network, the
oracle's integer parts against brute-force numpy restatements, and — when /root/reference is mounted
(build container only) — the oracle against the reference's own Python, live."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import native as onat
from oracle import net as onet
from oracle import quat, ref_shim
from rslo_b200.data import synthetic

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import make_golden as mg  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
VS, RG = [0.1, 0.1, 0.2], [-70.4, -38.4, -3, 70.4, 38.4, 5]


@pytest.fixture(scope="module")
def our_sd():
    import rslo_b200
    net, _ = rslo_b200.build_network(testing=True, seed=7)
    return net


def _case(net, name):
    g = np.load(os.path.join(GOLDEN, f"pair_{name}.npz"))
    seed, beams, n_az, T, step, wseed = (int(v) for v in g["meta"])
    onet.fill_weights(net, wseed)
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    return g, sd, mg.make_frames(seed, beams, n_az, T), step


def test_oracle_eval_matches_reference_golden(our_sd):
    g, sd, frames, step = _case(our_sd, "small_eval")
    out = onet.pair_forward(sd, frames, training=False)
    assert out["n_voxels"] == g["n_voxels"].tolist()
    np.testing.assert_allclose(out["pose"], g["pose"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(out["head"]["tq_map_g"][:, :, ::8, ::8].numpy(), g["tq_map_g_sample"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(out["cov"][0][::97].numpy(), g["cov0_sample"], rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("name", ["small_train", "small_train_warm", "small_train_t3", "full_train"])
def test_oracle_train_matches_reference_golden(our_sd, name):
    g, sd, frames, step = _case(our_sd, name)
    out = onet.pair_forward(sd, frames, training=True, step=step, grads_for=mg.GRAD_KEYS)
    np.testing.assert_allclose(out["pose"], g["pose"], rtol=1e-5, atol=1e-6)
    for k in ("loss", "translation_loss", "rotation_loss", "pyramid_loss", "C_loss"):
        np.testing.assert_allclose(out[k].reshape(-1), g[k], rtol=1e-5, atol=1e-6, err_msg=k)
    for k in mg.GRAD_KEYS:
        ref = g["grad:" + k]
        if ref.size == 0:
            assert out["grads"][k] is None
            continue
        got = mg.grad_sample(out["grads"][k])
        assert np.abs(got - ref).max() <= 1e-4 * max(np.abs(ref).max(), 1e-12), k


def test_voxeliser_oracle_vs_numpy_restatement():
    """oracle.c's sequential scan against an independent numpy statement of the same published
    algorithm (first-come ids, first max_points points, max_voxels cap)."""
    pts = synthetic.make_pair(4, n_beams=8, n_az=300)[0]
    for max_voxels in (100000, 500):
        r = onat.voxelize(pts, VS, RG, 10, max_voxels, 1, 8, -1.0)
        c = np.floor((pts[:, :3] - np.asarray(RG[:3], np.float32)) / np.asarray(VS, np.float32)).astype(np.int64)
        ok = ((c >= 0) & (c < np.array([1408, 768, 40]))).all(1)
        keys = (c[:, 2] * 768 + c[:, 1]) * 1408 + c[:, 0]
        seen, coords, counts, firsts = {}, [], [], []
        for i in np.nonzero(ok)[0]:
            k = int(keys[i])
            if k not in seen:
                if len(seen) >= max_voxels:
                    continue
                seen[k] = len(seen)
                coords.append(c[i, ::-1])
                counts.append(0)
                firsts.append(i)
            v = seen[k]
            if counts[v] < 10:
                counts[v] += 1
        assert np.array_equal(r["coordinates"], np.array(coords, np.int32))
        assert np.array_equal(r["num_points_per_voxel"], np.array(counts, np.int32))
        assert np.array_equal(r["voxels"][:, 0], pts[np.array(firsts)])


def test_voxeliser_edge_cases():
    empty = onat.voxelize(np.zeros((0, 7), np.float32), VS, RG)
    assert empty["coordinates"].shape == (0, 3)
    far = onat.voxelize(np.full((10, 7), 1e3, np.float32), VS, RG)
    assert far["voxels"].shape[0] == 0
    # boundary: a point exactly on the upper range edge is outside (floor gives grid size)
    edge = np.zeros((2, 7), np.float32)
    edge[0, :3] = [70.4, 0, 0]
    edge[1, :3] = [-70.4, -38.4, -3]
    r = onat.voxelize(edge, VS, RG)
    assert r["coordinates"].tolist() == [[0, 0, 0]]


def test_rulebook_oracle_vs_bruteforce():
    rng = np.random.default_rng(0)
    shape = [9, 20, 24]
    cells = rng.choice(np.prod(shape), 400, replace=False)
    co = np.stack([cells // (20 * 24), (cells // 24) % 20, cells % 24], 1).astype(np.int32)
    lut = {tuple(c): i for i, c in enumerate(co)}
    nbr = onat.subm_table(co, shape)
    for o in range(0, len(co), 7):
        for k in range(27):
            d = np.array([k // 9 - 1, (k // 3) % 3 - 1, k % 3 - 1])
            assert nbr[o, k] == lut.get(tuple(co[o] + d), -1)
    oc, oshape, tab, inv = onat.strided_table(co, shape, (3, 3, 3), (2, 2, 2), (1, 1, 1))
    assert oshape == [5, 10, 12]
    lin = (oc[:, 0] * 10 + oc[:, 1]) * 12 + oc[:, 2]
    assert (np.diff(lin) > 0).all()                                  # sorted-unique output order
    olut = {tuple(c): i for i, c in enumerate(oc)}
    for i in range(len(co)):
        for k in range(27):
            kz, ky, kx = k // 9, (k // 3) % 3, k % 3
            z, y, x = co[i] + 1 - np.array([kz, ky, kx])
            o = -1
            if min(z, y, x) >= 0 and z % 2 == 0 and y % 2 == 0 and x % 2 == 0:
                o = olut.get((z // 2, y // 2, x // 2), -1)
            assert inv[i, k] == o
            if o >= 0:
                assert tab[o, k] == i
    assert (tab >= 0).sum() == (inv >= 0).sum()


def test_nn_oracle_vs_numpy_and_ties():
    rng = np.random.default_rng(1)
    q = rng.integers(-4, 4, (300, 3)).astype(np.float32)
    t = rng.integers(-4, 4, (200, 3)).astype(np.float32)
    d, i = onat.nn(q, t)
    dd = ((q[:, None, :] - t[None]) ** 2).sum(-1)
    assert np.array_equal(i, dd.argmin(1).astype(np.int32))        # argmin returns the lowest index on ties
    assert np.array_equal(d, dd.min(1))


def test_nn_oracle_vs_reference_extension():
    """Pin: the reference's own chamfer extension (CPU `forward`, chamfer_distance.cpp:116-144),
    built by oracle/build.py into oracle/_ref/ where /root/reference is mounted."""
    so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "cd_ref.so")
    if not os.path.exists(so):
        pytest.skip("oracle/_ref/cd_ref.so not built (needs /root/reference)")
    import importlib.util
    spec = importlib.util.spec_from_file_location("cd_ref", so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    rng = np.random.default_rng(2)
    q = rng.uniform(-50, 50, (1, 700, 3)).astype(np.float32)
    t = rng.uniform(-50, 50, (1, 900, 3)).astype(np.float32)
    d1, d2 = torch.zeros(1, 700), torch.zeros(1, 900)
    i1, i2 = torch.zeros(1, 700, dtype=torch.int32), torch.zeros(1, 900, dtype=torch.int32)
    mod.forward(torch.from_numpy(q), torch.from_numpy(t), d1, d2, i1, i2)
    d, i = onat.nn(q[0], t[0], fused=False)            # the CPU twin is compiled without fma contraction
    assert np.array_equal(i, i1[0].numpy())
    np.testing.assert_allclose(d, d1[0].numpy(), rtol=1e-6)


def test_quaternion_roundtrip():
    g = torch.Generator().manual_seed(0)
    q = torch.nn.functional.normalize(torch.randn(64, 4, generator=g), dim=-1)
    q = q * torch.sign(q[:, 3:4])
    R = quat.quaternion_to_rotation_matrix(q)
    assert torch.allclose(R @ R.transpose(1, 2), torch.eye(3).expand(64, 3, 3), atol=1e-5)
    assert torch.allclose(torch.det(R), torch.ones(64), atol=1e-5)
    q2 = quat.rotation_matrix_to_quaternion(R)
    q2 = q2 * torch.sign(q2[:, 3:4])
    assert torch.allclose(q, q2, atol=1e-5)


def test_tq_map_local_global_roundtrip():
    """dataset.py:52-208: generate(local from global) then from_pointwise(local->global) returns the
    global pose in every cell."""
    tq = torch.tensor([0.8, -0.1, 0.05, 0.9995, 0.01, -0.02, 0.015])
    tq[3:] = tq[3:] / tq[3:].norm()
    local = onet.generate_pointwise_local_transformation(tq, 24, 44, np.asarray(RG, np.float32))
    glob = onet.from_pointwise_local_transformation(local[None], np.asarray(RG, np.float32))
    assert torch.allclose(glob[0, :3].reshape(3, -1).t(), tq[:3].expand(24 * 44, 3), atol=1e-4)
    assert torch.allclose(glob[0, 3:].reshape(4, -1).t(), tq[3:].expand(24 * 44, 4), atol=1e-5)


@pytest.mark.skipif(not ref_shim.available(), reason="reference tree not mounted (build container only)")
def test_oracle_matches_reference_live():
    """Live pin (build container): the REFERENCE network vs oracle/net.py on a pair no fixture holds."""
    rnet, vg = ref_shim.build_reference_net(testing=True, seed=7)
    onet.fill_weights(rnet, 17)
    rnet.global_step.fill_(4000)
    frames = mg.make_frames(9, 12, 300, 2)
    rnet.train()
    ret = rnet(mg.example_of(frames, vg))
    sd = {k: v.detach().clone() for k, v in rnet.state_dict().items()}
    out = onet.pair_forward(sd, frames, training=True, step=4000)
    np.testing.assert_allclose(out["loss"].reshape(-1), ret["loss"].detach().numpy().reshape(-1), rtol=1e-5)
    np.testing.assert_allclose(out["pose"], torch.cat([ret["translation_preds"], ret["rotation_preds"]], -1).numpy(),
                               rtol=1e-5, atol=1e-6)


@pytest.mark.skipif(not ref_shim.available(), reason="reference tree not mounted (build container only)")
def test_our_builders_accept_the_reference_protobuf_message():
    """Drop-in boundary: rslo_b200's voxel_builder / second_builder take the reference's REAL protobuf message
    (`pipeline_pb2.TrainEvalPipelineConfig().model.second` parsed by google.protobuf from the reference's own prototxt)
    exactly as `train_hdf5.py:92-101` passes it, and build a net with the reference's state_dict keys and shapes."""
    ref_shim.install()
    from google.protobuf import text_format
    from rslo.protos import pipeline_pb2                      # the reference's generated module
    from rslo_b200.builder import second_builder, voxel_builder
    cfg = pipeline_pb2.TrainEvalPipelineConfig()
    with open("/root/reference/config/kitti_train_ours.prototxt") as f:
        text_format.Merge(f.read(), cfg)
    vg = voxel_builder.build(cfg.model.second.voxel_generator)
    assert vg.grid_size.tolist() == [1408, 768, 40]
    net = second_builder.build(cfg.model.second, vg, measure_time=False, testing=False)
    rnet, _ = ref_shim.build_reference_net("/root/reference/config/kitti_train_ours.prototxt", testing=False, seed=7)
    ours = {k: tuple(v.shape) for k, v in net.state_dict().items()}
    theirs = {k: tuple(v.shape) for k, v in rnet.state_dict().items()}
    assert ours == theirs
    assert net.icp_iter == rnet.icp_iter and net._pyloss_exp_w_base == rnet._pyloss_exp_w_base
    assert float(net._rotation_loss.alpha) == float(rnet._rotation_loss.alpha)


def test_quaternion_conversions_vs_scipy():
    """kornia 0.4.0 is not vendored (parity unpinned): cross-check the restatement against an independent
    implementation (scipy Rotation, same x,y,z,w convention), incl. the four branches of the matrix->quaternion
    conversion."""
    from scipy.spatial.transform import Rotation
    rng = np.random.default_rng(0)
    q = rng.normal(size=(256, 4))
    q[:8] = np.array([[1, 0, 0, 1e-4], [0, 1, 0, 1e-4], [0, 0, 1, 1e-4], [0, 0, 0, 1],
                      [0.7, 0.7, 0, 0.1], [0, 0.7, 0.7, 0.1], [0.7, 0, 0.7, 0.1], [1, 1, 1, 1]])   # all trace branches
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    R_ref = Rotation.from_quat(q).as_matrix()
    R = quat.quaternion_to_rotation_matrix(torch.from_numpy(q)).numpy()
    np.testing.assert_allclose(R, R_ref, atol=1e-12)
    q_back = quat.rotation_matrix_to_quaternion(torch.from_numpy(R_ref)).numpy()
    q_ref = Rotation.from_matrix(R_ref).as_quat()
    sgn = np.sign(np.sum(q_back * q_ref, axis=1, keepdims=True))
    np.testing.assert_allclose(q_back * sgn, q_ref, atol=1e-6)


@pytest.mark.skipif(not ref_shim.available(), reason="reference tree not mounted (build container only)")
def test_optimizer_oracle_matches_reference_live():
    """f-N2 pin: oracle/optim.py (clip + true weight decay + Adam + one-cycle) against the reference's own
    OptimWrapper / OneCycle classes driving torch.optim.Adam, three steps, one parameter without gradient."""
    from functools import partial
    from oracle import optim as oopt
    ref_shim.install()
    from rslo.torchplus.train.fastai_optim import OptimWrapper
    from rslo.torchplus.train.learning_schedules_fastai import OneCycle
    g = torch.Generator().manual_seed(5)
    mods = torch.nn.ModuleList([torch.nn.Linear(7, 5), torch.nn.BatchNorm1d(5), torch.nn.Linear(5, 3), torch.nn.Linear(3, 2)])
    for p in mods.parameters():
        p.data = torch.randn(p.shape, generator=g)
    opt = OptimWrapper.create(partial(torch.optim.Adam, betas=(0.9, 0.99), amsgrad=False), 3e-3,
                              [[torch.nn.ModuleList(list(mods)[:2])], [torch.nn.ModuleList(list(mods)[2:])]],   # as
                              wd=1e-2, true_wd=True, bn_wd=True)     # get_voxeLO_net_layer_groups: a list of [ModuleList]
    sched = OneCycle(opt, 40, 0.8e-3, [0.95, 0.85], 10.0, 0.05)
    # the wrapper orders parameters by (non-BN, BN) group: follow torch's own order for the comparison
    params_ref = list(mods.parameters())
    params = [p.detach().clone() for p in params_ref]
    state = oopt.new_state(params)
    for step in range(3):
        grads = [torch.randn(p.shape, generator=g) * 30 for p in params_ref]
        grads[-1] = None                                    # last bias: no gradient this step
        for p, gr in zip(params_ref, grads):
            p.grad = None if gr is None else gr.clone()
        sched.step(step)
        lr, mom = oopt.one_cycle(step, 40, 0.8e-3, [0.95, 0.85], 10.0, 0.05)
        assert abs(opt.lr - lr) < 1e-12 and abs(opt.mom - mom) < 1e-12
        torch.nn.utils.clip_grad_norm_(params_ref, 10.0)
        opt.step()
        _, clipped = oopt.clip_grad_norm(grads, 10.0)
        oopt.adam_step(params, clipped, state, lr, mom, 0.99, 1e-8, wd=1e-2, true_wd=True)
        for a, b in zip(params, params_ref):
            assert torch.allclose(a, b.detach(), rtol=1e-6, atol=1e-7)


def _random_odoms(n, seed):
    rng = np.random.default_rng(seed)
    t = np.stack([rng.normal(1.0, 0.2, n), rng.normal(0, 0.03, n), rng.normal(0, 0.01, n)], 1)
    ang = rng.normal(0, 0.02, (n, 3))
    q = np.concatenate([np.ones((n, 1)), 0.5 * ang], 1)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    return np.concatenate([t, q], 1)


def test_kitti_eval_quaternion_matrix_matches_scipy():
    from scipy.spatial.transform import Rotation
    from oracle import kitti_eval as oke
    rng = np.random.default_rng(1)
    for _ in range(20):
        q = rng.normal(size=4) * rng.uniform(0.5, 2.0)
        ref = Rotation.from_quat([q[1], q[2], q[3], q[0]]).as_matrix()
        np.testing.assert_allclose(oke.quat_to_matrix(q), ref, atol=1e-12)


@pytest.mark.skipif(not ref_shim.available(), reason="reference tree not mounted (build container only)")
def test_kitti_eval_oracle_matches_reference_live():
    """f-N4 pin: oracle/kitti_eval.py against the reference's own odom_to_abs_pose and kittiOdomEval (its tq_to_RT needs
    the absent numpy-quaternion package and is handed the oracle's restatement of that one conversion)."""
    from oracle import kitti_eval as oke
    ref_shim.install()
    from rslo.utils import geometric, kitti_evaluation
    gts, preds = _random_odoms(1500, 3), _random_odoms(1500, 3)
    preds[:, :3] += np.random.default_rng(4).normal(0, 0.02, (1500, 3))
    a_ref, a_orc = geometric.odom_to_abs_pose(gts), oke.odom_to_abs_pose(gts)
    np.testing.assert_allclose(a_orc, a_ref, rtol=0, atol=1e-9)
    kitti_evaluation.tq_to_RT = oke.tq_to_RT
    ev = kitti_evaluation.kittiOdomEval()
    p_ref, p_orc = geometric.odom_to_abs_pose(preds), oke.odom_to_abs_pose(preds)
    e_ref = ev.calcSequenceErrors(p_ref, a_ref)
    e_orc = oke.calc_sequence_errors(p_orc, a_orc)
    assert len(e_ref) == len(e_orc) > 300
    np.testing.assert_allclose(np.asarray(e_orc), np.asarray(e_ref), rtol=1e-7, atol=1e-12)
    s_ref = ev.computeSegmentErr(e_ref)
    s_orc = oke.segment_errors(e_orc)
    assert s_ref.keys() == s_orc.keys()
    np.testing.assert_allclose(oke.segment_avg(s_orc), ev.computeSegmentAvgErr(s_ref), rtol=1e-9)
