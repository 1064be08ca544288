"""The sparse encoder + covariance decoder of `SpMiddleFHDWithCov2_3` (`rslo/models/middle.py:119-245`) as ONE autograd
node.

The layer-by-layer path (layers/sparse3d.py: a Module call, a SparseConvTensor and an autograd Function per layer) costs
~5 ms of host time per training step for ~150 kernel launches - more than the kernels take - and the step is host-bound.
This engine walks the same layer list (built from the same `SparseSequential` containers, so parameters, state_dict
keys and per-layer kernel choices are unchanged) in two tight loops, forward and backward, straight over the C-ABI
wrappers: no per-layer Module/Function/SparseConvTensor objects and one autograd node for the 25 layers.
Numerics are those of the per-layer path by construction (same kernels, same order); `RSLO_ENCODER_ENGINE=0` switches
back to it (A/B and the reference for tests/test_gpu_pair.py::test_encoder_engine_matches_layerwise_path).
"""
import os

import torch
from torch import nn

from .. import kernels as K
from . import sparse3d as S

USE_ENGINE = os.environ.get("RSLO_ENCODER_ENGINE", "1") != "0"


class _Conv:
    __slots__ = ("mod", "act", "slope", "cin", "cout", "kvol", "tc", "tc_wgrad", "pw", "pb")


class _BN:
    __slots__ = ("mod", "slope", "pw", "pb")


class SparseEncoderEngine:
    def __init__(self, middle):
        self.__dict__["middle"] = middle
        self.params = []
        self.trunk = self._collect(middle.middle_conv)              # L0 -> L2 (ret0)
        self.tail = self._collect(middle.middle_conv_tail)          # L2 -> L4 (dense BEV)
        self.cov = self._collect(middle.middle_cov_deconv)          # L2 -> L0 (covariance parameters)

    def _param(self, p):
        if p is None:
            return -1
        self.params.append(p)
        return len(self.params) - 1

    def _collect(self, seq):
        if seq._plan is None:
            seq._plan = seq._build_plan()
        out = []
        for m, act, slope in seq._plan:
            if act == "bn":
                r = _BN()
                r.mod, r.slope = m, slope
                r.pw, r.pb = self._param(m.weight), self._param(m.bias)
            elif isinstance(m, S.SparseConvolution):
                assert m.indice_key is not None, "the engine runs on prepared tables (keyed layers)"
                r = _Conv()
                r.mod, r.act, r.slope = m, act, slope
                r.cin, r.cout = m.in_channels, m.out_channels
                r.kvol = m.kernel_size[0] * m.kernel_size[1] * m.kernel_size[2]
                r.tc = S.USE_TC and min(r.cin, r.cout) >= 32 and K.spconv_tc_supported(r.cin, r.cout, r.kvol)
                r.tc_wgrad = S.USE_TC and K.spconv_tc_wgrad_supported(r.cin, r.cout)
                r.pw, r.pb = self._param(m.weight), self._param(m.bias)
            else:
                raise NotImplementedError(f"encoder engine: unexpected module {type(m).__name__}")
            out.append(r)
        return out

    # ---- forward ---------------------------------------------------------------------------------------------------
    def _run(self, layers, x, seg, entries, tape):
        for r in layers:
            if type(r) is _BN:
                m = r.mod
                z, mean_rstd = K.bn1d_seg_forward(x, seg, m.weight, m.bias, m.running_mean, m.running_var,
                                                  m.num_batches_tracked, m.eps, m.momentum, m.training, r.slope)
                if tape is not None:
                    tape.append((r, seg, m.training, x, mean_rstd))
                x = z
                continue
            m = r.mod
            e = entries[m.indice_key]
            inverse = m.inverse
            nbr = e.nbr_t if inverse else e.nbr
            n_out = e.n_in if inverse else e.n_out
            w = m.weight.view(-1, r.cin, r.cout)
            if r.tc:
                out = K.spconv_tc_forward(x, nbr, n_out, m._images.get(w, False, False), r.cin, r.cout, m.bias,
                                          act=r.act, slope=r.slope)
            else:
                out = K.spconv_forward(x, nbr, n_out, w, m.bias, act=r.act, slope=r.slope)
            if tape is not None:
                tape.append((r, e, inverse, x, out))
            x = out
            if inverse:
                seg = e.seg_in
            elif not m.subm:
                seg = e.seg_out
        return x, seg

    def forward(self, feats, entries, seg0, record):
        """feats [N0, Cin] rows of the stacked frames -> (tail features at L4, raw covariance parameters [N0, 7], tape)"""
        tape = ([], [], []) if record else (None, None, None)
        ret0, seg2 = self._run(self.trunk, feats.contiguous(), seg0, entries, tape[0])
        tail, _ = self._run(self.tail, ret0, seg2, entries, tape[1])
        cov, _ = self._run(self.cov, ret0, seg2, entries, tape[2])
        return tail, cov, tape

    # ---- backward --------------------------------------------------------------------------------------------------
    def _back(self, tape, g, pg, need_input_grad):
        """reverse walk of one layer list; g = gradient of its output -> gradient of its input (or None)"""
        for i in range(len(tape) - 1, -1, -1):
            rec = tape[i]
            r = rec[0]
            want_dx = need_input_grad or i > 0
            if type(r) is _BN:
                _, seg, training, x, mean_rstd = rec
                m = r.mod
                dx, dgamma, dbeta = K.bn1d_seg_backward(g.contiguous(), x, seg, mean_rstd, m.weight, m.bias, r.slope, training)
                pg[r.pw], pg[r.pb] = dgamma, dbeta
                g = dx
                continue
            _, e, inverse, x, out = rec
            m = r.mod
            g, gb = K.act_backward(g, out, r.act, r.slope, need_bias=r.pb >= 0)
            nbr = e.nbr_t if inverse else e.nbr
            nbr_t = e.nbr if inverse else e.nbr_t
            n_out = e.n_in if inverse else e.n_out
            n_in = e.n_out if inverse else e.n_in
            w = m.weight.view(-1, r.cin, r.cout)
            if r.tc_wgrad:
                gw = K.spconv_tc_backward_weight(x, g, nbr, n_out, w.shape)
            else:
                gw, _ = K.spconv_backward_weight(x, g, nbr, n_out, w.shape, need_bias=False)
            pg[r.pw] = gw.view_as(m.weight)
            if r.pb >= 0:
                pg[r.pb] = gb
            if want_dx:
                if r.tc:
                    g = K.spconv_tc_forward(g, nbr_t, n_in, m._images.get(w, True, e.mirror), r.cout, r.cin)
                else:
                    g = K.spconv_backward_data(g, nbr_t, n_in, w, e.mirror)
            else:
                g = None
        return g

    def backward(self, tape, g_tail, g_cov):
        pg = [None] * len(self.params)
        g_a = self._back(tape[1], g_tail, pg, True) if g_tail is not None else None
        g_b = self._back(tape[2], g_cov, pg, True) if g_cov is not None else None
        if g_a is None and g_b is None:
            return pg
        g0 = g_a if g_b is None else (g_b if g_a is None else g_a.add_(g_b))
        self._back(tape[0], g0, pg, False)
        return pg


class _EncoderFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, engine, entries, seg0, feats, *params):
        record = any(ctx.needs_input_grad[4:])
        tail, cov, tape = engine.forward(feats, entries, seg0, record)
        ctx.engine = engine
        if record:
            # the two outputs sit at the end of their layer lists' tapes: keep them through save_for_backward (a plain
            # attribute would be a reference cycle output -> grad_fn -> ctx -> output)
            t1, t2 = tape[1], tape[2]
            last1, last2 = t1[-1], t2[-1]
            t1[-1], t2[-1] = last1[:-1] + (None,), last2[:-1] + (None,)
            ctx.tape = tape
            ctx.save_for_backward(tail, cov)
        return tail, cov

    @staticmethod
    def backward(ctx, g_tail, g_cov):
        tail, cov = ctx.saved_tensors
        tape = ctx.tape
        t1, t2 = list(tape[1]), list(tape[2])
        t1[-1] = t1[-1][:-1] + (tail,)
        t2[-1] = t2[-1][:-1] + (cov,)
        pg = ctx.engine.backward((tape[0], t1, t2), g_tail, g_cov)
        return (None, None, None, None, *pg)


def encode(engine, feats, entries, seg0):
    """-> (tail features [N4, 64] at the coarsest level, raw covariance parameters [N0, 7])"""
    return _EncoderFn.apply(engine, entries, tuple(int(v) for v in seg0), feats, *engine.params)
