"""f-N3: device normal estimation (csrc/normals.cu through rslo_b200/data/normals.py, the mirror of `estimate_normal` in
the reference's `script/create_hdf5.py:130-147`) against the numpy brute-force oracle (oracle/normals.py; open3d itself
is absent: parity unpinned).  Compared where the normal is well defined - smallest covariance eigenvalue separated from
the middle one - as |cos| of the angle between the two (1e-5) and the sign (orientation towards the sensor)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_normals_match_bruteforce_oracle(cuda):
    from oracle import normals as onrm
    from rslo_b200.data import normals as dn
    from rslo_b200.data import synthetic
    a, _, _ = synthetic.make_pair(3, n_beams=16, n_az=300)            # ~4k points of a structured scene
    pts = a[:, :4].copy()
    ref, lam = onrm.estimate_normals(pts, 0.6, 30)
    got = dn.estimate_normal(pts).cpu().numpy().astype(np.float64)
    assert got.shape == ref.shape
    np.testing.assert_allclose(np.linalg.norm(got, axis=1), 1.0, atol=1e-5)
    fallback = (lam == 0).all(1)                                       # < 3 neighbours: (0,0,+-1) on both sides
    assert np.array_equal(got[fallback], ref[fallback])
    well = (~fallback) & (lam[:, 1] - lam[:, 0] > 1e-3 * np.maximum(lam[:, 2], 1e-12))
    assert well.sum() > 0.5 * len(pts)
    cos = (got[well] * ref[well]).sum(1)
    assert (cos > 1 - 1e-5).all(), float(cos.min())
    # orientation: towards the sensor at the origin
    assert ((got * (0 - pts[:, :3])).sum(1) >= -1e-6).all()


def test_planar_patch_normals_and_input_rows(cuda):
    from rslo_b200.data import normals as dn
    g = torch.Generator().manual_seed(0)
    xy = torch.rand(20000, 2, generator=g) * 20 - 10
    n_true = torch.tensor([0.3, -0.2, 0.93])
    n_true = n_true / n_true.norm()
    z = -(xy[:, 0] * n_true[0] + xy[:, 1] * n_true[1]) / n_true[2] - 2.0          # plane below the sensor
    pts = torch.cat([xy, z[:, None], torch.rand(20000, 1, generator=g)], 1)
    nrm = dn.estimate_normal(pts.cuda()).cpu()
    assert float((nrm @ n_true).abs().min()) > 1 - 1e-4
    rows = dn.points_with_normals(pts.numpy())
    assert rows.shape == (20000, 7) and torch.equal(rows[:, :4].cpu(), pts)
    # an isolated point gets the (0,0,1) fallback, oriented, and the dataset rule zeroes it
    far = torch.tensor([[500.0, 500.0, 30.0, 0.5]])
    rows = dn.points_with_normals(torch.cat([pts, far]))
    assert float(rows[-1, 4:].abs().max()) == 0.0
