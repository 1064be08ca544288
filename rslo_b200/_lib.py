"""ctypes binding of the C ABI in include/rslo_b200.h (rslo_b200/_C/librslo_b200.so).

There is no CPU fallback: if the shared library is missing, importing this module raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_C", "librslo_b200.so")

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"rslo_b200: CUDA library not built ({LIB_PATH} missing). Run `python -c 'import "
        f"__graft_entry__ as g; g.build()'` or `make -C rslo_b200/csrc`. There is no CPU fallback.")

lib = C.CDLL(LIB_PATH)

_vp, _i, _f, _sz = C.c_void_p, C.c_int, C.c_float, C.c_size_t


class TqGeom(C.Structure):               # rslo_tq_geom_t
    _fields_ = [("H", C.c_int), ("W", C.c_int), ("ox", C.c_float), ("oy", C.c_float), ("oz", C.c_float),
                ("vsx", C.c_float), ("vsy", C.c_float), ("vsz", C.c_float)]


# name -> (restype, argtypes); mirrors include/rslo_b200.h one to one
SIGNATURES = {
    "rslo_abi_version": (_i, []),
    "rslo_last_error": (C.c_char_p, []),
    "rslo_kernel_launch_count": (C.c_ulonglong, []),
    "rslo_set_graph_capture_hint": (None, [_i]),
    "rslo_nn_workspace_bytes": (_sz, [_i, _i]),
    "rslo_nn_exact": (_i, [_vp, _i, _vp, _i, _vp, _vp, _vp, _sz, _vp]),
    "rslo_nn_brute": (_i, [_vp, _i, _vp, _i, _vp, _vp, _vp]),
    "rslo_voxelize_workspace_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "rslo_voxelize": (_i, [_vp, _i, _i, C.POINTER(_f), C.POINTER(_f), _i, _i, _i, _i, _i, _i, _i, _f, _i,
                           _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _i, _vp, _sz, _vp]),
    "rslo_vfe_mean": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp]),
    "rslo_site_table_workspace_bytes": (_sz, [_i, _i, _i]),
    "rslo_site_table_build": (_i, [_vp, _i, _i, _vp, _i, _i, _i, _vp, _vp, _vp, _sz, _vp]),
    "rslo_subm_table": (_i, [_vp, _i, _i, _vp, _i, _i, _i, _vp, _vp, _i, _i, _i, _vp, _vp]),
    "rslo_strided_workspace_bytes": (_sz, [_i, _i, _i]),
    "rslo_strided_table": (_i, [_vp, _i, _i, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i,
                                _vp, _vp, _i, _vp, _vp, _vp, _vp, _sz, _vp]),
    "rslo_table_concat": (_i, [_vp, C.c_longlong, _i, _vp, _vp]),
    "rslo_table_concat_multi": (_i, [_vp, _i, _vp]),
    "rslo_spconv_forward": (_i, [_vp, _vp, _i, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _i, _f, _vp, _vp]),
    "rslo_spconv_transpose_weight": (_i, [_vp, _i, _i, _i, _i, _vp, _vp]),
    "rslo_spconv_backward_data": (_i, [_vp, _vp, _i, _vp, _i, _i, _i, _vp, _vp, _vp]),
    "rslo_spconv_backward_weight": (_i, [_vp, _vp, _vp, _i, _vp, _i, _i, _i, _vp, _vp, _vp]),
    "rslo_act_backward": (_i, [_vp, _vp, _i, _vp, _i, _i, _f, _vp, _vp, _vp]),
    "rslo_spconv_tc_supported": (_i, [_i, _i, _i]),
    "rslo_spconv_tc_image_bytes": (_sz, [_i, _i, _i]),
    "rslo_spconv_tc_prepare": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _vp]),
    "rslo_spconv_tc_workspace_bytes": (_sz, [_i, _i]),
    "rslo_spconv_tc_forward": (_i, [_vp, _vp, _i, _vp, _i, _i, _i, _vp, _vp, _i, _f, _vp, _vp, _sz, _vp]),
    "rslo_spconv_tc_wgrad_supported": (_i, [_i, _i]),
    "rslo_spconv_tc_backward_weight": (_i, [_vp, _vp, _vp, _i, _vp, _i, _i, _i, _vp, _vp]),
    "rslo_conv2d_tc_supported": (_i, [_i, _i, _i, _i]),
    "rslo_conv2d_split": (_i, [_vp, _sz, _vp, _vp]),
    "rslo_conv2d_tc_prepare": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _vp]),
    "rslo_conv2d_tc_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "rslo_conv2d_tc_forward": (_i, [_vp, _i, _i, _i, _i, _vp, _i, _i, _i, _vp, _i, _vp, _vp, _i, _vp, _sz, _vp]),
    "rslo_conv2d_tc_backward_data": (_i, [_vp, _i, _i, _i, _i, _vp, _i, _i, _i, _vp, _i, _vp, _sz, _vp]),
    "rslo_conv2d_tc_wgrad_scratch_bytes": (_sz, [_i, _i, _i]),
    "rslo_conv2d_tc_backward_weight": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _i, _vp, _vp]),
    "rslo_conv2d_multi_prepare": (_i, [_vp, _i, _vp]),
    "rslo_conv2d_multi_wgrad_finish": (_i, [_vp, _i, _vp]),
    "rslo_head_pack_input": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _vp]),
    "rslo_head_unpack_grad": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp]),
    "rslo_bn_act_forward": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _f, _f, _i, _vp, _i, _vp, _vp, _vp, _vp]),
    "rslo_bn_act_backward": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _i, _i, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp]),
    "rslo_upcat_split": (_i, [_vp, _i, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    "rslo_upcat_backward": (_i, [_vp, _i, _i, _i, _i, _i, _i, _i, _vp, _i, _vp]),
    "rslo_bias_grad": (_i, [_vp, _i, _i, _i, _vp, _vp]),
    "rslo_head_tail_forward": (_i, [_vp] * 6 + [_i, TqGeom] + [_vp] * 14 + [_vp]),
    "rslo_head_tail_backward": (_i, [_vp] * 7 + [_i, TqGeom] + [_vp] * 13 + [_vp]),
    "rslo_loss_tail_forward": (_i, [_vp] * 10 + [_i, _i, TqGeom] + [_vp] * 4 + [_f] * 4 + [_vp] * 4 + [_vp]),
    "rslo_loss_tail_backward": (_i, [_vp] * 8 + [_i, TqGeom] + [_vp] * 4 + [_f] * 4 + [_vp] * 8 + [_vp]),
    "rslo_dense_from_sites": (_i, [_vp, _i, _vp, _vp, _i, _i, _i, _vp, _vp]),
    "rslo_kabsch_workspace_bytes": (_sz, []),
    "rslo_kabsch": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "rslo_kth_threshold": (_i, [_vp, _i, _i, _f, _vp, _vp]),
    "rslo_cov_residual_forward": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _f, _vp, _vp, _vp]),
    "rslo_cov_residual_backward": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _f, _vp, _vp, _vp, _vp, _vp,
                                        _vp, _vp]),
    "rslo_dense_backward": (_i, [_vp, _i, _vp, _i, _i, _vp, _i, _i, _i, _vp, _vp]),
    "rslo_bn1d_seg_forward": (_i, [_vp, _i, C.POINTER(_i), _i, _vp, _vp, _vp, _vp, _vp, _f, _f, _i, _f, _vp, _vp, _vp, _vp]),
    "rslo_bn1d_seg_backward": (_i, [_vp, _vp, _i, C.POINTER(_i), _i, _vp, _vp, _vp, _f, _i, _vp, _vp, _vp, _vp, _vp]),
    "rslo_pair_transform_workspace_bytes": (_sz, []),
    "rslo_pair_transform_forward": (_i, [_vp, _i, _i, _vp, _vp, _i, _vp, _vp, _vp]),
    "rslo_pair_transform_backward": (_i, [_vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _sz, _vp]),
    "rslo_odom_to_abs_pose": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _vp]),
    "rslo_kitti_sequence_errors": (_i, [_vp, _i, _vp, _i, _vp, _i, _vp, _vp, _vp]),
    "rslo_estimate_normals_workspace_bytes": (_sz, [_i]),
    "rslo_estimate_normals": (_i, [_vp, _i, _i, _f, _i, C.POINTER(_f), _vp, _vp, _sz, _vp]),
    "rslo_grad_norm_workspace_bytes": (_sz, []),
    "rslo_grad_sumsq": (_i, [_vp, _sz, _vp, _vp, _sz, _vp]),
    "rslo_adam_step": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _f, _f, _f, _f, _f, _f, _f, _i, _i, _i, _vp]),
}


def _bind():
    missing = []
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError:
            missing.append(name)
            continue
        fn.restype = res
        fn.argtypes = args
    if missing:
        raise ImportError(f"rslo_b200: {LIB_PATH} lacks symbols {missing}; rebuild it")


_bind()


class ConcatSeg(C.Structure):            # rslo_concat_seg_t
    _fields_ = [("src", C.c_void_p), ("dst", C.c_void_p), ("count", C.c_longlong), ("add", C.c_int), ("reserved", C.c_int)]


class AdamChunk(C.Structure):            # rslo_adam_chunk_t
    _fields_ = [("p", C.c_void_p), ("off", C.c_uint), ("n", C.c_uint), ("flags", C.c_uint), ("reserved", C.c_uint)]


class RsloError(RuntimeError):
    pass


def check(rc, what=""):
    if rc != 0:
        msg = lib.rslo_last_error().decode("utf-8", "replace")
        raise RsloError(f"{what} failed (cudaError {rc}): {msg}")


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def stream():
    """current CUDA stream of the current device as a raw handle (torch.cuda.current_stream() costs ~15 us of host
    time per call; this is on the path of every kernel launch)"""
    import torch
    return C.c_void_p(torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice()))


def raw_stream():
    import torch
    dev = torch._C._cuda_getDevice()
    return dev, torch._C._cuda_getCurrentRawStream(dev)


_WS = {}


def workspace(nbytes, slot="default"):
    """Growable per-(device, stream, slot) scratch buffer: reuse is ordered by the stream it belongs to, so
    work on different streams (preparation stream, concurrent examples) never shares scratch memory."""
    import torch
    dev, st = raw_stream()
    key = (dev, st, slot)
    buf = _WS.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(int(nbytes * 1.25) + 1024, dtype=torch.uint8, device=f"cuda:{dev}")
        _WS[key] = buf
    return buf
