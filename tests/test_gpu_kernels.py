"""Parity of the CUDA kernels (through the C ABI) against the CPU oracle.  Integer / index outputs
are compared bit-exactly; floating-point outputs with the tolerance written at each check."""
import importlib.util
import os

import numpy as np
import pytest
import torch

from oracle import native as onat
from oracle import sparse as osp
from rslo_b200.data import synthetic

pytestmark = pytest.mark.gpu

VS = [0.1, 0.1, 0.2]
RG = [-70.4, -38.4, -3, 70.4, 38.4, 5]
GRID = [1408, 768, 40]


@pytest.fixture(scope="module")
def K(cuda):
    from rslo_b200 import kernels
    return kernels


@pytest.fixture(scope="module")
def scan_pair():
    return synthetic.make_pair(0)


def _vox_gpu(K, pts, **kw):
    out = K.voxelize(torch.from_numpy(pts).cuda(), VS, RG, GRID, **kw)
    n = int(out["n_dev"].item())
    return out, n


@pytest.mark.parametrize("max_voxels,thr", [(40000, -1.0), (20000, -1.0), (40000, 0.2), (3000, 0.5)])
def test_voxelize_bit_exact(K, scan_pair, max_voxels, thr):
    pts = scan_pair[0]
    ref = onat.voxelize(pts, VS, RG, 10, max_voxels, 1, 8, thr)
    out, n = _vox_gpu(K, pts, max_voxels=max_voxels, height_threshold=thr)
    assert n == ref["voxels"].shape[0]
    assert np.array_equal(out["coordinates"][:n, 1:].cpu().numpy(), ref["coordinates"])
    assert (out["coordinates"][:n, 0] == 0).all()
    assert np.array_equal(out["num_points_per_voxel"][:n].cpu().numpy(), ref["num_points_per_voxel"])
    assert np.array_equal(out["voxels"][:n].cpu().numpy(), ref["voxels"])     # bit-exact copies
    mean_ref = osp.vfe_mean(ref["voxels"], ref["num_points_per_voxel"]).numpy()
    # fp32 sum-order tolerance
    np.testing.assert_allclose(out["mean"][:n].cpu().numpy(), mean_ref, rtol=1e-5, atol=1e-6)
    m2 = K.vfe_mean(out["voxels"][:n], out["num_points_per_voxel"][:n]).cpu().numpy()
    np.testing.assert_allclose(m2, mean_ref, rtol=1e-5, atol=1e-6)


def test_voxelize_edge_cases(K):
    # all points outside the range -> zero voxels; one voxel with more points than max_points;
    # shuffled (non scan-ordered) ragged input
    far = np.full((100, 7), 1000.0, np.float32)
    out, n = _vox_gpu(K, far)
    assert n == 0
    one = np.zeros((37, 7), np.float32)
    one[:, :3] = [1.01, 2.02, 0.03]
    one[:, 3] = np.arange(37)
    ref = onat.voxelize(one, VS, RG)
    out, n = _vox_gpu(K, one)
    assert n == 1 and int(out["num_points_per_voxel"][0]) == 10
    assert np.array_equal(out["voxels"][:1].cpu().numpy(), ref["voxels"])
    rng = np.random.default_rng(3)
    shuf = synthetic.make_pair(1)[0]
    shuf = shuf[rng.permutation(len(shuf))][:50001]
    ref = onat.voxelize(shuf, VS, RG)
    out, n = _vox_gpu(K, shuf)
    assert n == len(ref["coordinates"])
    assert np.array_equal(out["coordinates"][:n, 1:].cpu().numpy(), ref["coordinates"])
    assert np.array_equal(out["voxels"][:n].cpu().numpy(), ref["voxels"])


def _tables_gpu(K, coors4, n, shape):
    t = {}
    tab0 = K.site_table_build(coors4, n, shape)
    t["subm0"] = K.subm_table(coors4, n, tab0)
    tab1, c1, n1d, t["conv3d2"], t["conv3d2_inv"] = K.strided_table(coors4, n, shape, (3, 3, 3), (2, 2, 2), (1, 1, 1))
    n1 = int(n1d[0].item())
    t["L1"] = (c1, n1, tab1)
    t["subm1"] = K.subm_table(c1, n1, tab1)
    tab2, c2, n2d, t["conv3d3"], t["conv3d3_inv"] = K.strided_table(c1, n1, tab1.shape, (3, 3, 3), (2, 2, 2), (1, 1, 1))
    n2 = int(n2d[0].item())
    t["L2"] = (c2, n2, tab2)
    t["subm2"] = K.subm_table(c2, n2, tab2)
    tab3, c3, n3d, t["conv3d4"], _ = K.strided_table(c2, n2, tab2.shape, (3, 3, 3), (2, 2, 2), (0, 1, 1))
    n3 = int(n3d[0].item())
    t["L3"] = (c3, n3, tab3)
    t["subm3"] = K.subm_table(c3, n3, tab3)
    tab4, c4, n4d, t["conv3d5"], _ = K.strided_table(c3, n3, tab3.shape, (3, 1, 1), (2, 1, 1), (0, 0, 0))
    n4 = int(n4d[0].item())
    t["L4"] = (c4, n4, tab4)
    return t


def test_rulebook_bit_exact(K, scan_pair):
    pts = scan_pair[1]
    vox = onat.voxelize(pts, VS, RG)
    co = vox["coordinates"]
    n = len(co)
    shape = [41, 768, 1408]
    ref = osp.build_tables(co, shape)
    coors4 = torch.from_numpy(np.concatenate([np.zeros((n, 1), np.int32), co], 1)).cuda()
    t = _tables_gpu(K, coors4, n, shape)
    for lvl in ["L1", "L2", "L3", "L4"]:
        c, nn_, tab = t[lvl]
        assert nn_ == len(ref[lvl].coors), lvl
        assert list(tab.shape) == list(ref[lvl].shape), lvl
        assert np.array_equal(c[:nn_, 1:].cpu().numpy(), ref[lvl].coors), lvl
    sizes = {"subm0": n, "conv3d2": t["L1"][1], "conv3d2_inv": n, "subm1": t["L1"][1], "conv3d3": t["L2"][1],
             "conv3d3_inv": t["L1"][1], "subm2": t["L2"][1], "conv3d4": t["L3"][1], "subm3": t["L3"][1],
             "conv3d5": t["L4"][1]}
    for key, rows in sizes.items():
        assert np.array_equal(t[key][:rows].cpu().numpy(), ref[key]), key


@pytest.mark.parametrize("cin,cout,K_", [(7, 16, 27), (16, 16, 27), (16, 32, 27), (32, 32, 27), (32, 64, 27),
                                         (64, 64, 27), (64, 64, 3), (64, 32, 27), (32, 16, 27), (16, 7, 27)])
def test_spconv_forward_and_wgrad(K, cin, cout, K_):
    g = torch.Generator().manual_seed(cin * 100 + cout)
    n_in, n_out = 3000, 2500
    nbr = torch.randint(0, n_in, (n_out, K_), generator=g, dtype=torch.int32)
    nbr[torch.rand((n_out, K_), generator=g) < 0.7] = -1
    feat = torch.randn((n_in, cin), generator=g)
    w = torch.randn((K_, cin, cout), generator=g) * 0.1
    b = torch.randn(cout, generator=g)
    ref = torch.nn.functional.leaky_relu(osp.gather_conv(feat, nbr, w, b), 0.01)
    out = K.spconv_forward(feat.cuda(), nbr.cuda(), n_out, w.cuda(), b.cuda(), act=1, slope=0.01)
    # fp32 with a different summation order
    np.testing.assert_allclose(out.cpu().numpy(), ref.numpy(), rtol=1e-4, atol=1e-5)
    featr = feat.clone().requires_grad_(True)
    wr = w.clone().requires_grad_(True)
    br = b.clone().requires_grad_(True)
    o = osp.gather_conv(featr, nbr, wr, br)
    go = torch.randn(o.shape, generator=g)
    o.backward(go)
    gw, gb = K.spconv_backward_weight(feat.cuda(), go.cuda(), nbr.cuda(), n_out, (K_, cin, cout))
    np.testing.assert_allclose(gw.cpu().numpy(), wr.grad.numpy(), rtol=1e-3, atol=1e-3)
    np.testing.assert_allclose(gb.cpu().numpy(), br.grad.numpy(), rtol=1e-3, atol=1e-3)


def test_spconv_backward_data_on_real_tables(K, scan_pair):
    pts = scan_pair[0][:60000]
    vox = onat.voxelize(pts, VS, RG)
    co = vox["coordinates"]
    n = len(co)
    shape = [41, 768, 1408]
    nbr = onat.subm_table(co, shape)
    c1, s1, nb_s, nb_inv = onat.strided_table(co, shape, (3, 3, 3), (2, 2, 2), (1, 1, 1))
    g = torch.Generator().manual_seed(5)
    # (forward table, transposed table, n_in, mirror): submanifold, strided, inverse
    for (table, table_t, n_in, mirror) in [(nbr, nbr, n, True), (nb_s, nb_inv, n, False),
                                           (nb_inv, nb_s, len(c1), False)]:
        feat = torch.randn((n_in, 16), generator=g, requires_grad=True)
        w = (torch.randn((27, 16, 32), generator=g) * 0.1).requires_grad_(True)
        o = osp.gather_conv(feat, table, w)
        go = torch.randn(o.shape, generator=g)
        o.backward(go)
        gi = K.spconv_backward_data(go.cuda(), torch.from_numpy(table_t).cuda(), n_in, w.detach().cuda(), mirror)
        np.testing.assert_allclose(gi.cpu().numpy(), feat.grad.numpy(), rtol=1e-4, atol=1e-4)


def test_dense(K):
    rng = np.random.default_rng(0)
    D, H, W = 2, 96, 176
    cells = rng.choice(D * H * W, 9000, replace=False)
    cells.sort()
    co = np.stack([np.zeros_like(cells), cells // (H * W), (cells // W) % H, cells % W], 1).astype(np.int32)
    coors = torch.from_numpy(co).cuda()
    tab = K.site_table_build(coors, len(co), (D, H, W), need_perm=False)
    feat = torch.randn(len(co), 64)
    dense = K.dense_from_sites(feat.cuda(), tab).cpu()
    ref = torch.zeros(64, D, H, W)
    ref[:, co[:, 1], co[:, 2], co[:, 3]] = feat.t()
    assert torch.equal(dense, ref.view(64 * D, H, W))
    gd = torch.randn(64 * D, H, W)
    gf = K.dense_backward(gd.cuda(), coors, len(co), (D, H, W), 64).cpu()
    assert torch.equal(gf, gd.view(64, D, H, W)[:, co[:, 1], co[:, 2], co[:, 3]].t())


def _ref_cuda_nn(q, t):
    """The reference's own CUDA kernel, when oracle/_ref was built (it travels with the repo)."""
    so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "cd_ref.so")
    if not os.path.exists(so):
        return None
    spec = importlib.util.spec_from_file_location("cd_ref", so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    d = torch.zeros(1, q.shape[0], device="cuda")
    i = torch.zeros(1, q.shape[0], dtype=torch.int32, device="cuda")
    mod.forward_cuda_one_direction(q[None].contiguous(), t[None].contiguous(), d, i)
    torch.cuda.synchronize()
    return d[0], i[0]


@pytest.mark.parametrize("n,m,kind", [(5000, 5003, "uniform"), (3000, 1500, "uniform"), (40000, 40000, "lidar"),
                                      (20000, 777, "lidar"), (4096, 4096, "ties")])
def test_nn_bit_exact(K, scan_pair, n, m, kind):
    rng = np.random.default_rng(n + m)
    if kind == "uniform":
        q = rng.uniform(-50, 50, (n, 3)).astype(np.float32)
        t = rng.uniform(-50, 50, (m, 3)).astype(np.float32)
    elif kind == "ties":
        # integer lattice: many exactly equal distances -> the lowest index must win
        q = rng.integers(-8, 8, (n, 3)).astype(np.float32)
        t = rng.integers(-8, 8, (m, 3)).astype(np.float32)
    else:
        a = onat.voxelize(scan_pair[0], VS, RG)
        b = onat.voxelize(scan_pair[1], VS, RG)
        q = osp.vfe_mean(a["voxels"], a["num_points_per_voxel"]).numpy()[:n, :3].copy()
        t = osp.vfe_mean(b["voxels"], b["num_points_per_voxel"]).numpy()[:m, :3].copy()
    d_ref, i_ref = onat.nn(q, t, fused=True)
    qc, tc = torch.from_numpy(q).cuda(), torch.from_numpy(t).cuda()
    for brute in (False, True):
        d, i = K.nn_exact(qc, tc, brute=brute)
        assert np.array_equal(i.cpu().numpy(), i_ref), f"idx mismatch brute={brute}"
        assert np.array_equal(d.cpu().numpy(), d_ref), f"dist mismatch brute={brute}"
    r = _ref_cuda_nn(qc, tc)
    if r is not None:   # pin oracle AND kernel against the reference's own CUDA kernel
        assert np.array_equal(r[1].cpu().numpy(), i_ref)
        assert np.array_equal(r[0].cpu().numpy(), d_ref)


@pytest.mark.parametrize("cin,cout,K_,n_out", [(64, 64, 27, 2577), (32, 32, 27, 2577), (32, 64, 27, 2577),
                                               (64, 32, 27, 2577), (64, 64, 3, 2577), (64, 64, 27, 19594),
                                               (32, 32, 27, 40001), (16, 16, 27, 2577), (16, 32, 27, 2577),
                                               (32, 16, 27, 2577), (16, 16, 27, 80001)])
def test_spconv_tensor_core_forward(K, cin, cout, K_, n_out):
    """tcgen05 split-TF32 implicit-GEMM kernel vs the oracle's gather-conv: FP32-level agreement.
    Row counts cover the offset-split paths (4-way, 2-way) and the unsplit one."""
    g = torch.Generator().manual_seed(cin * 100 + cout + K_)
    n_in = 3000                              # n_out is not a multiple of the 128-row tile
    nbr = torch.randint(0, n_in, (n_out, K_), generator=g, dtype=torch.int32)
    nbr[torch.rand((n_out, K_), generator=g) < 0.6] = -1
    nbr[128:256] = -1                        # a whole tile without any neighbour -> bias only
    nbr[300:428, 1:] = -1                    # a tile that uses a single offset
    feat = torch.randn((n_in, cin), generator=g)
    w = torch.randn((K_, cin, cout), generator=g) * 0.1
    b = torch.randn(cout, generator=g)
    ref = torch.nn.functional.leaky_relu(osp.gather_conv(feat, nbr, w, b), 0.01)
    img = K.spconv_tc_prepare(w.cuda())
    out = K.spconv_tc_forward(feat.cuda(), nbr.cuda(), n_out, img, cin, cout, b.cuda(), act=1, slope=0.01)
    torch.cuda.synchronize()
    err = (out.cpu() - ref).abs().max().item()
    print(f"tc max abs err {err:.3e} (|ref| max {ref.abs().max().item():.2f})")
    np.testing.assert_allclose(out.cpu().numpy(), ref.numpy(), rtol=1e-5, atol=2e-5)
    # data gradient through the transposed image == FFMA data-gradient kernel == autograd
    go = torch.randn((n_out, cout), generator=g)
    n_t = 2000
    nbr_t = torch.randint(0, n_out, (n_t, K_), generator=g, dtype=torch.int32)
    nbr_t[torch.rand((n_t, K_), generator=g) < 0.6] = -1
    for mirror in (False, True):
        img_t = K.spconv_tc_prepare(w.cuda(), transpose=True, mirror=mirror)
        gi = K.spconv_tc_forward(go.cuda(), nbr_t.cuda(), n_t, img_t, cout, cin)
        gi_ref = K.spconv_backward_data(go.cuda(), nbr_t.cuda(), n_t, w.cuda(), mirror)
        np.testing.assert_allclose(gi.cpu().numpy(), gi_ref.cpu().numpy(), rtol=1e-5, atol=2e-5)


def _span_cov2_torch(p):
    """losses.py:348-363 with torch ops (test-side restatement, runs on the GPU through autograd)."""
    from oracle import quat
    l1 = p[:, 0:1]
    l2 = l1 + p[:, 1:2]
    l3 = l2 + p[:, 2:3]
    q = p[:, 3:] / (torch.norm(p[:, 3:], dim=-1, keepdim=True) + 1e-9)
    V = quat.quaternion_to_rotation_matrix(q)
    lam = torch.cat([l1, l2, l3], dim=1)
    return (V * lam[:, None, :]) @ V.transpose(-1, -2)


def test_cov_residual_matches_torch_autograd(K):
    """Fused covariance-residual kernel (forward + backward) vs the same math composed from torch ops
    (torch.inverse / torch.det, boolean-mask ROI) differentiated by autograd."""
    g = torch.Generator().manual_seed(3)
    n, m = 5000, 4700
    pred = (torch.randn(n, 3, generator=g) * 10).cuda().requires_grad_(True)
    tgt = (torch.randn(m, 3, generator=g) * 10).cuda().requires_grad_(True)
    idx = torch.randint(0, m, (n,), generator=g, dtype=torch.int32).cuda()
    cp = torch.randn(n, 7, generator=g)
    ct = torch.randn(m, 7, generator=g)
    cp[:, :3] = torch.rand(n, 3, generator=g) * 0.5 + 0.05
    ct[:, :3] = torch.rand(m, 3, generator=g) * 0.5 + 0.05
    cp, ct = cp.cuda().requires_grad_(True), ct.cuda().requires_grad_(True)
    ang = torch.tensor(0.3)
    R = torch.tensor([[torch.cos(ang), -torch.sin(ang), 0], [torch.sin(ang), torch.cos(ang), 0], [0, 0, 1.0]]).cuda()
    dist = torch.rand(n, generator=g).cuda() * 3
    thr = torch.tensor([2.0]).cuda()
    loss = K.cov_residual(pred, tgt, cp, ct, R, idx, dist, thr, 0.005)
    (loss.sum() * 1.7).backward()
    got = [loss.detach().clone(), pred.grad.clone(), tgt.grad.clone(), cp.grad.clone(), ct.grad.clone()]
    for t in (pred, tgt, cp, ct):
        t.grad = None
    roi = dist < thr
    il = idx.long()
    sigma = _span_cov2_torch(cp)[roi] + R @ _span_cov2_torch(ct)[il][roi] @ R.t()
    d = (pred - tgt[il])[roi]
    ref = ((d[:, None, :] @ torch.inverse(sigma) @ d[:, :, None]).reshape(-1).mean()
           + 0.005 * (0.5 * torch.log(torch.det(sigma))).mean())
    (ref * 1.7).backward()
    want = [ref.detach().reshape(1), pred.grad, tgt.grad, cp.grad, ct.grad]
    for a, b, name in zip(got, want, ["loss", "g_pred", "g_target", "g_cov_pred", "g_cov_target"]):
        scale = float(b.abs().max())
        assert float((a - b).abs().max()) <= 2e-4 * scale + 1e-7, name


def test_kabsch_fused_gather_and_weights(K):
    """kabsch with the association gather + normal-cosine weights fused == explicit torch preparation."""
    g = torch.Generator().manual_seed(9)
    n, m = 4000, 3500
    src = torch.randn(n, 3, generator=g).cuda() * 5
    tgt = torch.randn(m, 3, generator=g).cuda() * 5
    idx = torch.randint(0, m, (n,), generator=g, dtype=torch.int32).cuda()
    nrm = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1).cuda()
    nrm[::7] = 0                                                   # zeroed normals, as the dataset makes them
    dist = torch.rand(n, generator=g).cuda()
    thr = torch.tensor([0.8]).cuda()
    R1, t1 = K.kabsch(src, tgt, tgt_idx=idx, normal=nrm, dist=dist, dist_threshold=thr)
    assoc = tgt[idx.long()].contiguous()
    w = torch.nn.functional.cosine_similarity(nrm, assoc - src, dim=-1).abs()
    R2, t2 = K.kabsch(src, assoc, weight=(w * w).contiguous(), dist=dist, dist_threshold=thr)
    np.testing.assert_allclose(R1.cpu().numpy(), R2.cpu().numpy(), atol=1e-5)
    np.testing.assert_allclose(t1.cpu().numpy(), t2.cpu().numpy(), atol=1e-4)


@pytest.mark.parametrize("n", [1, 7, 1000, 40000, 123457])
def test_kth_threshold_matches_torch(K, n):
    g = torch.Generator().manual_seed(n)
    d = (torch.rand(n, generator=g) * 4).pow(2).cuda()
    d[::3] = d[0].clone()                                # ties
    for ratio in (0.97, 0.5, 0.0):
        k = min(n, 1 + int(n * ratio))
        want = torch.max(torch.kthvalue(d, k).values, torch.ones((), device="cuda"))
        got = K.kth_threshold(d, k, 1.0)
        assert float(got) == float(want)
    assert float(K.kth_threshold(d, n, 0.0)) == float(d.max())


@pytest.mark.parametrize("cin,cout,K_,n_out", [(64, 64, 27, 2577), (32, 32, 27, 2577), (32, 64, 27, 2577),
                                               (64, 32, 27, 2577), (64, 64, 3, 2577), (64, 64, 27, 39188),
                                               (64, 64, 27, 50), (16, 16, 27, 80001), (16, 32, 27, 2577),
                                               (32, 16, 27, 2577)])
def test_spconv_tensor_core_wgrad(K, cin, cout, K_, n_out):
    """tcgen05 weight-gradient kernel (MN-major operands, offsets stacked along M) vs autograd of the oracle."""
    g = torch.Generator().manual_seed(cin * 7 + cout + K_ + n_out)
    n_in = 3000
    nbr = torch.randint(0, n_in, (n_out, K_), generator=g, dtype=torch.int32)
    nbr[torch.rand((n_out, K_), generator=g) < 0.6] = -1
    feat = torch.randn((n_in, cin), generator=g)
    w = (torch.randn((K_, cin, cout), generator=g) * 0.1).requires_grad_(True)
    o = osp.gather_conv(feat, nbr, w)
    go = torch.randn(o.shape, generator=g)
    o.backward(go)
    gw = K.spconv_tc_backward_weight(feat.cuda(), go.cuda(), nbr.cuda(), n_out, (K_, cin, cout))
    torch.cuda.synchronize()
    ref = w.grad
    err = (gw.cpu() - ref).abs().max().item()
    print(f"tc wgrad max abs err {err:.3e} (|ref| max {ref.abs().max().item():.2f})")
    # fp32 sums over up to ~4e4 rows in a different order (TMEM accumulation + atomics across CTAs)
    assert err <= 2e-5 * ref.abs().max().item() + 1e-5


def test_stress_grid_voxelize_and_rulebook_bit_exact(K):
    """BASELINE config 5: 300k-ray scan, 0.05 x 0.05 x 0.1 m voxels (grid 2816 x 1536 x 80, 346 M cells),
    up to 250000 voxels: voxel coordinates / counts and the first two levels of rulebooks, bit-exact."""
    vs, grid = [0.05, 0.05, 0.1], [2816, 1536, 80]
    pts = synthetic.make_pair(7, n_beams=128, n_az=2344)[0]
    ref = onat.voxelize(pts, vs, RG, 10, 250000, 1, 8, -1.0)
    out = K.voxelize(torch.from_numpy(pts).cuda(), vs, RG, grid, max_voxels=250000, materialize=False, with_table=True)
    n = int(out["n_dev"].item())
    assert n == len(ref["coordinates"]) and n > 100000
    assert np.array_equal(out["coordinates"][:n, 1:].cpu().numpy(), ref["coordinates"])
    assert np.array_equal(out["num_points_per_voxel"][:n].cpu().numpy(), ref["num_points_per_voxel"])
    mean_ref = osp.vfe_mean(ref["voxels"], ref["num_points_per_voxel"]).numpy()
    np.testing.assert_allclose(out["mean"][:n].cpu().numpy(), mean_ref, rtol=1e-5, atol=1e-6)
    shape = [81, 1536, 2816]
    co = ref["coordinates"]
    coors4 = out["coordinates"][:n].contiguous()
    nbr = K.subm_table(coors4, n, out["table"])
    assert np.array_equal(nbr[:n].cpu().numpy(), onat.subm_table(co, shape))
    tab1, c1, n1d, nb, nb_inv = K.strided_table(coors4, n, shape, (3, 3, 3), (2, 2, 2), (1, 1, 1))
    oc, oshape, rnb, rinv = onat.strided_table(co, shape, (3, 3, 3), (2, 2, 2), (1, 1, 1))
    n1 = int(n1d[0].item())
    assert n1 == len(oc) and list(tab1.shape) == oshape
    assert np.array_equal(c1[:n1, 1:].cpu().numpy(), oc)
    assert np.array_equal(nb[:n1].cpu().numpy(), rnb) and np.array_equal(nb_inv[:n].cpu().numpy(), rinv)
    # exact NN at this size: grid search == tiled brute force (both bit-identical to the reference formula)
    q = out["mean"][:n, :3].contiguous()
    t = (q + torch.tensor([0.9, 0.02, 0.0], device="cuda")).contiguous()
    d1, i1 = K.nn_exact(q, t)
    d2, i2 = K.nn_exact(q, t, brute=True)
    assert torch.equal(i1, i2) and torch.equal(d1, d2)


@pytest.mark.parametrize("cin,cout,h,w,ks,st,B", [(256, 128, 24, 44, 3, 2, 1), (128, 128, 24, 22, 3, 1, 2),
                                                  (512, 128, 12, 22, 3, 1, 1), (192, 64, 48, 44, 3, 1, 1),
                                                  (64, 32, 48, 88, 3, 1, 1), (256, 128, 24, 44, 1, 2, 1),
                                                  (256, 256, 12, 22, 3, 1, 3), (32, 7, 24, 44, 1, 1, 2),
                                                  (128, 256, 24, 44, 3, 2, 2), (64, 1, 16, 24, 1, 1, 1)])
def test_conv2d_tensor_core_matches_fp64(cuda, cin, cout, h, w, ks, st, B):
    """Head convolutions (csrc/conv2d_tc.cu: TMA-staged split-TF32 tcgen05 implicit GEMM) through the C ABI:
    forward (+bias, + fused BatchNorm statistics), data gradient (store and accumulate) and weight gradient
    against float64 autograd.  Narrow heads (7 / 1 output channels) run zero-padded to 32."""
    from rslo_b200 import kernels as K
    g = torch.Generator().manual_seed(cin + cout + h)
    wt = (torch.randn(cout, cin, ks, ks, generator=g) * (2.0 / (cin * ks * ks)) ** 0.5).cuda()
    bias = (torch.randn(cout, generator=g) * 0.1).cuda()
    x = torch.randn(B, h, w, cin, generator=g).cuda()
    coutp = (cout + 31) // 32 * 32
    bias_p = torch.zeros(coutp, device="cuda")
    bias_p[:cout] = bias
    xs = K.conv2d_split(x)
    stats = torch.zeros(B, coutp, 2, dtype=torch.float64, device="cuda")
    y = K.conv2d_tc_forward(xs, K.conv2d_tc_prepare(wt, 0, cout_padded=coutp), coutp, ks, st, bias=bias_p, stats=stats,
                            imgs_per_group=1)
    xd = x.permute(0, 3, 1, 2).double().requires_grad_(True)
    wd = wt.double().requires_grad_(True)
    yd = torch.nn.functional.conv2d(xd, wd, bias.double(), st, ks // 2)
    go = torch.zeros(y.shape, device="cuda")
    go[..., :cout] = torch.randn(y.shape[:3] + (cout,), generator=g).cuda()
    yd.backward(go[..., :cout].permute(0, 3, 1, 2).double())

    def rel(a, b):
        return float((a.double() - b).abs().max() / b.abs().max())
    assert rel(y[..., :cout], yd.permute(0, 2, 3, 1)) < 5e-6
    if coutp != cout:
        assert float(y[..., cout:].abs().max()) == 0.0
    ref_stats = torch.stack([yd.sum(dim=(2, 3)), (yd * yd).sum(dim=(2, 3))], dim=-1)
    assert rel(stats[:, :cout], ref_stats) < 1e-5
    gs = K.conv2d_split(go)
    img_t = K.conv2d_tc_prepare(wt, 1, cout_padded=coutp)
    dx = K.conv2d_tc_backward_data(gs, img_t, (B, h, w, cin), ks, st)
    assert rel(dx, xd.grad.permute(0, 2, 3, 1)) < 5e-6
    base = torch.randn(dx.shape, generator=g).cuda()
    acc = base.clone()
    K.conv2d_tc_backward_data(gs, img_t, (B, h, w, cin), ks, st, out=acc, accumulate=True)
    assert rel(acc - base, xd.grad.permute(0, 2, 3, 1)) < 5e-6
    dw = K.conv2d_tc_backward_weight(xs, gs, ks, st, cout_real=cout)
    assert dw.shape == wt.shape and rel(dw, wd.grad) < 5e-6


@pytest.mark.parametrize("C,seg,train,slope", [(16, (1000, 2500, 37, 4000), True, 0.01), (32, (777, 1), True, 0.01),
                                               (16, (5000,), True, -1.0), (32, (300, 0, 900), False, 0.01),
                                               (64, (40000, 39000, 25000, 26000), True, 0.01)])
def test_bn1d_over_stacked_frames_matches_float64(cuda, C, seg, train, slope):
    """csrc/bn1d_seg.cu: BatchNorm1d (+ LeakyReLU) with per-frame statistics over stacked frames against float64
    torch.nn.functional.batch_norm applied frame by frame (`rslo/models/middle.py:181-213` run once per frame):
    outputs, running statistics (updated frame after frame), num_batches_tracked, dx, dgamma, dbeta."""
    from rslo_b200 import kernels as K
    g = torch.Generator().manual_seed(C + len(seg))
    n = sum(seg)
    x = (torch.randn(n, C, generator=g) * 1.7 + 0.4).cuda()
    gamma = (0.5 + torch.rand(C, generator=g)).cuda()
    beta = (0.3 * torch.randn(C, generator=g)).cuda()
    rm = (0.1 * torch.randn(C, generator=g)).cuda()
    rv = (0.5 + torch.rand(C, generator=g)).cuda()
    dz = torch.randn(n, C, generator=g).cuda()
    eps, mom = 1e-5, 0.1
    xd = x.double().requires_grad_(True)
    gd, bd = gamma.double().requires_grad_(True), beta.double().requires_grad_(True)
    rm_ref, rv_ref = rm.double().clone(), rv.double().clone()
    outs, frames = [], 0
    for f in torch.split(xd, list(seg)):
        if f.shape[0] == 0:
            continue
        if train and f.shape[0] == 1:           # torch refuses one value per channel; the statistics are still defined
            mu = f[0]
            o = (f - mu) * torch.rsqrt(torch.zeros_like(mu) + eps) * gd + bd
            rm_ref = (1 - mom) * rm_ref + mom * mu.detach()
            rv_ref = (1 - mom) * rv_ref
        else:
            o = torch.nn.functional.batch_norm(f, rm_ref, rv_ref, gd, bd, training=train, momentum=mom, eps=eps)
        frames += 1
        outs.append(torch.nn.functional.leaky_relu(o, slope) if slope >= 0 else o)
    zd = torch.cat(outs)
    zd.backward(dz.double())

    nbt = torch.zeros((), dtype=torch.long, device="cuda")
    rm_k, rv_k = rm.clone(), rv.clone()
    z, mr = K.bn1d_seg_forward(x, seg, gamma, beta, rm_k, rv_k, nbt, eps, mom, train, slope)
    dx, dgamma, dbeta = K.bn1d_seg_backward(dz, x, seg, mr, gamma, beta, slope, train)

    def rel(a, b):
        return float((a.double() - b.double()).abs().max() / b.double().abs().max())
    assert rel(z, zd) < 1e-5
    if train:
        assert rel(rm_k, rm_ref) < 1e-6 and rel(rv_k, rv_ref) < 1e-5 and int(nbt) == frames
    else:
        assert torch.equal(rm_k, rm) and torch.equal(rv_k, rv) and int(nbt) == 0
    assert rel(dx, xd.grad) < 2e-5
    assert rel(dgamma, gd.grad) < 1e-5 and rel(dbeta, bd.grad) < 1e-5


@pytest.mark.parametrize("n,identity", [(40000, False), (37, False), (5000, True)])
def test_pair_transform_matches_torch_float64(cuda, n, identity):
    """csrc/pair_transform.cu: y = x @ R(q)^T + t with the kornia-0.4.0 conversion restated in rslo_b200/utils/pose_utils.py
    (the torch chain it replaces, `voxel_odom_net.py:671-690`), forward, R, and gradients w.r.t. q (w,x,y,z) and t;
    x is the xyz part of 7-column rows (row-strided view)."""
    from rslo_b200 import kernels as K
    from rslo_b200.utils import pose_utils
    g = torch.Generator().manual_seed(n)
    feats = (torch.randn(n, 7, generator=g) * 20).cuda()
    q = torch.tensor([[0.9, 0.05, -0.1, 0.3]]) * 1.7                      # not normalised on purpose
    t = torch.tensor([[0.8, -0.2, 0.05]])
    gy = torch.randn(n, 3, generator=g).cuda()
    qk, tk = q.cuda().requires_grad_(True), t.cuda().requires_grad_(True)
    y, R = K.pair_transform(feats, qk[0], tk[0], identity=identity)
    qd, td = q.double().cuda().requires_grad_(True), t.double().cuda().requires_grad_(True)
    if identity:
        Rd = torch.eye(3, dtype=torch.float64, device="cuda")[None]
        yd = feats[:, :3].double()[None] @ Rd.transpose(1, 2)
    else:
        Rd = pose_utils.quaternion_to_rotation_matrix(torch.roll(qd, -1, -1))
        yd = feats[:, :3].double()[None] @ Rd.transpose(1, 2) + td[:, None, :]

    def rel(a, b):
        return float((a.double() - b).abs().max() / b.abs().max().clamp_min(1e-30))
    assert rel(y, yd[0]) < 1e-6 and rel(R, Rd[0]) < 1e-6
    (y * gy).sum().backward()
    if identity:
        assert qk.grad is None or float(qk.grad.abs().max()) == 0
        return
    (yd[0] * gy.double()).sum().backward()
    assert rel(qk.grad, qd.grad) < 1e-5 and rel(tk.grad, td.grad) < 1e-5
