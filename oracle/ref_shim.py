"""Import the REFERENCE's own `rslo` package in this container (only here: /root/reference does not
exist on the GPU box) to pin the oracle and to generate tests/golden fixtures.

TEST INFRASTRUCTURE ONLY.  Recipe from SURVEY.md Appendix A: stub modules for the dependencies that
are not installed (apex, kornia, spconv, the chamfer JIT module, plotting / IO packages),
`collections.Iterable` back-compat and the pure-python protobuf implementation.  The stubs carry
the oracle's restatements where arithmetic is needed (kornia conversions, spconv layers, NN).
"""
import collections
import collections.abc
import importlib.abc
import importlib.machinery
import os
import sys
import types

REF = "/root/reference"


def available():
    return os.path.isdir(os.path.join(REF, "rslo"))


class _Permissive(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        m = _Permissive(self.__name__ + "." + name)
        setattr(self, name, m)
        return m

    def __call__(self, *a, **k):
        return None


class _Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    ROOTS = ("matplotlib", "seaborn", "mpl_toolkits", "open3d", "h5py", "fire", "quaternion", "skimage",
             "tensorboardX", "transforms3d", "petrel_client", "cv2", "numba", "pypcd", "shapely", "nuscenes")

    def find_spec(self, name, path, target=None):
        if name.split(".")[0] in self.ROOTS:
            return importlib.machinery.ModuleSpec(name, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _Permissive(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        pass


_installed = False


def install():
    """Make `import rslo...` resolve to the reference tree with working stubs."""
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError("reference tree not mounted")
    import torch
    from torch import nn

    from . import quat, sparse_modules

    os.environ.setdefault("PROTOCOL_BUFFERS_PYTHON_IMPLEMENTATION", "python")
    collections.Iterable = collections.abc.Iterable
    collections.Mapping = collections.abc.Mapping
    collections.Sequence = collections.abc.Sequence
    import numpy as np
    for alias, real in (("int", int), ("float", float), ("bool", bool), ("object", object)):
        if not hasattr(np, alias):
            setattr(np, alias, real)

    # apex: identity amp decorators, SyncBatchNorm == BatchNorm at world size 1
    apex = types.ModuleType("apex")
    amp = types.ModuleType("apex.amp")
    amp.float_function = lambda f: f
    amp.half_function = lambda f: f
    parallel = types.ModuleType("apex.parallel")

    class SyncBatchNorm(nn.BatchNorm2d):
        def __init__(self, num_features, eps=1e-5, momentum=0.1, affine=True, track_running_stats=True,
                     process_group=None, channel_last=False, fuse_relu=False):
            super().__init__(num_features, eps, momentum, affine, track_running_stats)

        def _check_input_dim(self, input):
            pass

    parallel.SyncBatchNorm = SyncBatchNorm
    parallel.ReduceOp = torch.distributed.ReduceOp if hasattr(torch.distributed, 'ReduceOp') else object
    parallel.DistributedDataParallel = nn.parallel.DistributedDataParallel
    sbk = types.ModuleType("apex.parallel.sync_batchnorm_kernel")
    sbk.SyncBatchnormFunction = object
    parallel.sync_batchnorm_kernel = sbk
    apex.amp, apex.parallel = amp, parallel
    sys.modules.update({"apex": apex, "apex.amp": amp, "apex.parallel": parallel,
                        "apex.parallel.sync_batchnorm_kernel": sbk})

    kornia = types.ModuleType("kornia")
    kornia.quaternion_to_rotation_matrix = quat.quaternion_to_rotation_matrix
    kornia.rotation_matrix_to_quaternion = quat.rotation_matrix_to_quaternion
    sys.modules["kornia"] = kornia

    sp = types.ModuleType("spconv")
    for name in ("SparseConvTensor", "SubMConv3d", "SparseConv3d", "SparseInverseConv3d", "SparseSequential"):
        setattr(sp, name, getattr(sparse_modules, name))
    sp_utils = types.ModuleType("spconv.utils")
    sp_utils.VoxelGenerator = sparse_modules.VoxelGenerator
    sp.utils = sp_utils
    sys.modules.update({"spconv": sp, "spconv.utils": sp_utils})

    cdm = types.ModuleType("thirdparty.chamfer_distance.chamfer_distance")
    cdm.OneDirectionChamferDistanceWithIdx = sparse_modules.OneDirectionChamferDistanceWithIdx
    cdm.ChamferDistanceWithIdx = sparse_modules.OneDirectionChamferDistanceWithIdx
    cdm.ChamferDistance = sparse_modules.OneDirectionChamferDistanceWithIdx
    tp = types.ModuleType("thirdparty")
    tp.__path__ = []
    tpc = types.ModuleType("thirdparty.chamfer_distance")
    tpc.__path__ = []
    sys.modules.update({"thirdparty": tp, "thirdparty.chamfer_distance": tpc,
                        "thirdparty.chamfer_distance.chamfer_distance": cdm})

    sys.meta_path.append(_Finder())
    for p in (REF, os.path.join(REF, "rslo")):
        if p not in sys.path:
            sys.path.insert(0, p)
    # the reference calls .cuda() unconditionally in the loss path (voxel_odom_net.py:340,625)
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
    _installed = True


def _patch_inplace_elu():
    """middle.py:237 writes `features[:, :3] = F.elu(features[:, :3]) + ...` in place; under current
    torch autograd rejects that (elu saved a view of the tensor being overwritten; torch 1.2 did not
    check).  Give elu a private copy of its input inside that module only — same values, same grads."""
    import types as _t

    import torch.nn.functional as _F
    from rslo.models import middle as _m
    if getattr(_m.F, "_rslo_patched", False):
        return
    proxy = _t.SimpleNamespace(**{k: getattr(_F, k) for k in dir(_F) if not k.startswith("__")})
    proxy.elu = lambda x, *a, **k: _F.elu(x.clone(), *a, **k)
    proxy._rslo_patched = True
    _m.F = proxy


def build_reference_net(prototxt=None, testing=True, seed=7):
    """The reference's own builders on the reference's own prototxt -> (net, voxel_generator)."""
    install()
    import torch
    from google.protobuf import text_format
    from rslo.builder import second_builder, voxel_builder
    from rslo.protos import pipeline_pb2
    prototxt = prototxt or os.path.join(REF, "config", "kitti_eval_ours.prototxt")
    cfg = pipeline_pb2.TrainEvalPipelineConfig()
    with open(prototxt) as f:
        text_format.Merge(f.read(), cfg)
    vg = voxel_builder.build(cfg.model.second.voxel_generator)
    _patch_inplace_elu()
    torch.manual_seed(seed)
    net = second_builder.build(cfg.model.second, vg, measure_time=False, testing=testing)
    return net, vg
