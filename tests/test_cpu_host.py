"""CPU suite (no GPU): the host-side mirror of the reference interface — prototxt reader, builders,
registries, state_dict contract, C-ABI exports, gradient all-reduce over gloo (world size 2)."""
import ctypes
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cabi_exports_every_declared_symbol():
    """The C-ABI library loads without a GPU and exports every function include/rslo_b200.h declares."""
    hdr = open(os.path.join(ROOT, "include", "rslo_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(rslo_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 20
    lib = ctypes.CDLL(os.path.join(ROOT, "rslo_b200", "_C", "librslo_b200.so"))
    for n in sorted(names):
        assert hasattr(lib, n), f"missing export {n}"
    lib.rslo_abi_version.restype = ctypes.c_int
    assert lib.rslo_abi_version() >= 1
    from rslo_b200 import _lib
    assert set(_lib.SIGNATURES) == names, "ctypes table and header disagree"
    # ... and nothing else: the trace / diagnostics builds (-DTC_TRACE: rslo_debug_*) must not be what ships
    import subprocess
    out = subprocess.run(["nm", "-D", "--defined-only", os.path.join(ROOT, "rslo_b200", "_C", "librslo_b200.so")],
                         capture_output=True, text=True).stdout
    exported = {ln.split()[-1] for ln in out.splitlines() if " T rslo_" in ln}
    assert exported == names, f"undeclared exports: {sorted(exported - names)}"


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "rslo_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{f} imports the oracle"


def test_prototxt_reader_matches_reference_values():
    from rslo_b200.builder import config
    cfg = config.load(os.path.join(ROOT, "rslo_b200", "config", "kitti_ours.prototxt"))
    m = cfg.model.second
    assert m.network_class_name == "UnVoxelOdomNetICP3"
    assert m.icp_iter == 2
    assert list(m.voxel_generator.voxel_size) == [float(np.float32(v)) for v in (0.1, 0.1, 0.2)]
    assert m.voxel_generator.height_threshold == -1.0
    assert m.odom_predictor.layer_nums == [3, 5, 5]
    assert m.odom_predictor.dropout > 0 and m.odom_predictor.dropout < 1e-20
    assert m.loss.consistency_loss.penalize_ratio == float(np.float32(0.97))
    assert m.loss.pyramid_rotation_loss.loss_type == ""            # proto3 default
    assert m.middle_feature_extractor.bn_type == "None"
    # comments that run on after values, brackets, nested messages without colon
    msg = config.parse_text('a: 1 # trailing comment\nb { c: [1, 2,3] d: "x#y" }', root="Any")
    assert msg.a == 1 and msg.b.c == [1, 2, 3] and msg.b.d == "x#y"


def test_prototxt_reader_reads_reference_files_when_mounted():
    ref = "/root/reference/config/kitti_train_ours.prototxt"
    if not os.path.exists(ref):
        pytest.skip("reference tree not mounted")
    from rslo_b200.builder import config
    ours = config.load(os.path.join(ROOT, "rslo_b200", "config", "kitti_ours.prototxt")).model.second
    theirs = config.load(ref).model.second
    for sect in ("voxel_generator", "voxel_feature_extractor", "middle_feature_extractor", "odom_predictor"):
        assert getattr(ours, sect)._v == getattr(theirs, sect)._v, sect
    assert ours.loss.consistency_loss._v == theirs.loss.consistency_loss._v


def test_builders_registries_and_state_dict_contract():
    import rslo_b200
    from rslo_b200.models import middle, odom_pred, voxel_encoder, voxel_odom_net
    assert voxel_odom_net.get_voxelnet_class("UnVoxelOdomNetICP3")
    assert voxel_encoder.get_vfe_class("SimpleVoxel_XYZINormalC")
    assert middle.get_middle_class("SpMiddleFHDWithCov2_3")
    assert odom_pred.get_odom_class("UNRResNetOdomPredEncDecSVDTempMask")
    net, vg = rslo_b200.build_network(testing=True, seed=7)
    assert vg.grid_size.tolist() == [1408, 768, 40]
    assert list(net.middle_feature_extractor.sparse_shape) == [41, 768, 1408]
    assert net.name == "voxel_odom_net" and net.voxel_generator is vg
    golden = json.load(open(os.path.join(ROOT, "tests", "golden", "state_dict_shapes.json")))
    sd = net.state_dict()
    assert {k: list(v.shape) for k, v in sd.items()} == golden          # the reference's 485 keys / shapes
    assert sum(p.numel() for p in net.parameters()) == 12004079
    assert sum(p.numel() for p in net.parameters() if p.requires_grad) == 12004069
    # optimizer constraint: parameters live in leaf modules only (optimizer_builder.py:37-45)
    for m in net.modules():
        if len(list(m.children())) > 0 and not hasattr(m, "alpha"):     # loss alphas are picked up by name (:48-65)
            assert len(list(m.parameters(recurse=False))) == 0, type(m)
    # losses: pyramid losses ARE the main loss modules when not configured (losses_builder.py:40-50)
    assert net._pyramid_rotation_loss is net._rotation_loss
    assert float(net._rotation_loss.alpha) == -2.5 and float(net._translation_loss.alpha) == 0.0
    assert not net._consistency_loss.alpha.requires_grad
    net.update_global_step()
    assert net.get_global_step() == 1
    net.clear_global_step()
    assert net.get_global_step() == 0


def test_kernels_fail_loudly_without_gpu():
    """No CPU fallback: the operator layer refuses CPU tensors instead of computing on the host."""
    from rslo_b200 import kernels as K
    with pytest.raises(AssertionError):
        K.nn_exact(torch.zeros(4, 3), torch.zeros(4, 3))
    from rslo_b200.thirdparty.chamfer_distance.chamfer_distance import OneDirectionChamferDistanceWithIdx
    with pytest.raises(NotImplementedError):
        OneDirectionChamferDistanceWithIdx()(torch.zeros(1, 4, 3), torch.zeros(1, 4, 3))


def test_adaptive_l2_loss_matches_formula():
    from rslo_b200.core.losses import AdaptiveWeightedL2Loss
    l = AdaptiveWeightedL2Loss(init_alpha=-2.5, learn_alpha=True, loss_weight=1)
    p, t = torch.randn(3, 4, 5, 6), torch.randn(3, 4, 5, 6)
    m = (torch.rand(3, 1, 5, 6) > 0.5).float()
    got = l(p, t, mask=m)
    me = m.expand_as(t)
    lb = ((p - t) ** 2 * me).sum(dim=(1, 2, 3)) / (me.sum(dim=(1, 2, 3)) + 1e-12)
    want = (lb * np.exp(2.5) / 3).sum() - 2.5
    assert torch.allclose(got, want, rtol=1e-5)


WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, {root!r})
from rslo_b200.utils.distributed import FlatGradAllReducer, init_from_env
rank, local, world = init_from_env("gloo")
torch.manual_seed(0)
net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 2), torch.nn.Linear(3, 3))
red = FlatGradAllReducer(net)
red.broadcast_params()
red.zero_()
g = torch.Generator().manual_seed(100 + rank)          # each rank: its own shard of frame pairs
x = torch.randn(4, 6, generator=g)
net[2](net[1](net[0](x))).pow(2).sum().backward()      # net[3] gets no gradient (unused head parts)
local_flat = red.pack().clone()
red.all_reduce()
gathered = [torch.zeros_like(local_flat) for _ in range(world)]
dist.all_gather(gathered, local_flat)
want = sum(gathered) / world
assert torch.allclose(red.flat, want, atol=1e-6), (red.flat - want).abs().max()
assert all(p.grad.data_ptr() >= red.flat.data_ptr() for p in net.parameters())
assert float(net[3].weight.grad.abs().sum()) == 0.0
red.zero_()
assert all(p.grad is None for p in net.parameters())
if rank == 0:
    print("OK", float(red.flat.abs().sum()))
dist.destroy_process_group()
"""


def test_flat_grad_allreduce_gloo_world2(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(WORKER.format(root=ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29631", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r), LOCAL_RANK=str(r)),
                              stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True) for r in range(2)]
    outs = [p.communicate(timeout=120) for p in procs]
    for p, (o, e) in zip(procs, outs):
        assert p.returncode == 0, e[-2000:]
    assert "OK" in outs[0][0]


def test_bench_reference_arm_rank_nonzero_exits_quietly():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       env=dict(os.environ, RANK="1", WORLD_SIZE="2"), capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_pass_through_generator_packs_the_raw_scan_without_cuda():
    """f-N1: generate() in pass-through mode (what a forked DataLoader worker calls) touches no GPU API"""
    import numpy as np
    import rslo_b200
    _, vg = rslo_b200.build_network(testing=True, seed=7)
    vg.pass_through = True
    pts = np.random.default_rng(0).normal(size=(500, 7)).astype(np.float32)
    r = vg.generate(pts, 40000)
    assert r["voxels"].shape == (500, 1, 7) and np.array_equal(r["voxels"][:, 0], pts)
    assert r["coordinates"].shape == (500, 3) and r["coordinates"].dtype == np.int32 and (r["coordinates"] == -1).all()
    assert r["num_points_per_voxel"].tolist() == [1] * 500


def test_bench_arms_describe_the_same_config():
    """Both arms of bench.py must name the same workload with the same keys (the driver compares their `config`)."""
    import importlib
    bench = importlib.import_module("bench")
    for wl in ("train", "eval", "warm", "stress"):
        cfg = bench.workload_config(wl)
        ours = bench.b200_config(cfg, 1, cfg["pairs_per_gpu"], 0, cfg["mode"] == "train")
        ref = bench.reference_config(cfg, 240.0)
        assert set(ours) == set(ref)
        for k in ("workload", "mode", "pairs_per_gpu", "global_pairs_per_step", "parallelism", "global_step", "optimizer"):
            assert ours[k] == ref[k], k
    assert bench.parse_args.__defaults__ is None        # defaults live in argparse: steps 100 / warm-up 20
    import sys
    argv, sys.argv = sys.argv, ["bench.py"]
    try:
        a = bench.parse_args()
    finally:
        sys.argv = argv
    assert a.gpus == 1 and a.steps >= 100 and a.warmup >= 20
