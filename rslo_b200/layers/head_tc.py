"""The convolutional trunk of the odometry head on the repo's own sm_100a kernels.

Replaces the cuDNN / eager-torch execution of `rslo/models/odom_pred.py:152-260` (`odom_pred_base.py:155-276`
layer lists, `custom_resnet_spc.py:224-298` BasicBlock, `layers/MaskConv.py:53-63`, `layers/SparseConv.py:96-132`
SyncBN/ReLU): every convolution is the TMA-staged split-TF32 tcgen05 implicit GEMM of csrc/conv2d_tc.cu (forward,
data gradient, weight gradient), everything between two convolutions is one fused streaming kernel of
csrc/head_ops.cu.  Activations stay NHWC and are handed from layer to layer as the split pair the convolutions read
(hi = RN_tf32(x), lo = RN_tf32(x - hi)), so there is no layout conversion, no separate BatchNorm-statistics pass (the
convolution epilogue accumulates them) and no separate operand-split pass anywhere in the trunk.

The engine owns no parameters: it reads the head module's `nn.Conv2d` / `nn.BatchNorm2d` parameters and buffers
(names and OIHW shapes of the reference's state_dict) and returns their gradients through one autograd Function.

Batching: `imgs_per_group` images share BatchNorm batch statistics (= the pairs of one sample, as in the
reference where one forward call sees one sample); several samples of a step go through the trunk in one pass
with per-sample statistics, which is numerically the reference's sample-by-sample execution.
"""
import ctypes as C

import torch

from .. import kernels as K


class _PrepEntry(C.Structure):           # rslo_conv_prep_t
    _fields_ = [("w", C.c_void_p), ("img_fwd", C.c_void_p), ("img_bwd", C.c_void_p), ("bias", C.c_void_p),
                ("bias_pad", C.c_void_p), ("Cout", C.c_int), ("CoutP", C.c_int), ("Cin", C.c_int), ("taps", C.c_int)]


class _FinishEntry(C.Structure):         # rslo_wgrad_finish_t
    _fields_ = [("dW", C.c_void_p), ("gw", C.c_void_p), ("Cout", C.c_int), ("CoutP", C.c_int), ("Cin", C.c_int),
                ("taps", C.c_int)]


def _to_device_table(entries, device, keepalive):
    """ctypes struct array -> device uint8 tensor.  While a CUDA graph is being captured the copy must come from
    pinned memory that outlives the graph (the memcpy node re-reads it on every replay)."""
    raw = bytes(entries)
    host = torch.frombuffer(bytearray(raw), dtype=torch.uint8)
    if torch.cuda.is_current_stream_capturing():
        host = host.pin_memory()
        keepalive.append(host)
        return host.to(device, non_blocking=True)
    return host.to(device)


class _Act:
    """One activation of the trunk: fp32 values `z` (None for pure operand buffers), split pair, gradient."""
    __slots__ = ("z", "split", "dz", "shape")

    def __init__(self, shape, z=None, split=None):
        self.shape, self.z, self.split, self.dz = tuple(shape), z, split, None


class _ConvRec:
    __slots__ = ("mod", "ks", "stride", "cin", "cout", "coutp", "taps", "img_off", "dw_off", "bias_off", "pw", "pb",
                 "img_fwd", "img_bwd", "bias_pad")


class _BNRec:
    __slots__ = ("mod", "c", "off", "pw", "pb", "repeat")


class _Run:
    """State of one forward pass that the backward pass needs."""
    __slots__ = ("tape", "ipg", "G", "B", "outputs", "xin", "x_shape")


class HeadTrunkEngine:
    def __init__(self, head):
        self.__dict__["head"] = head
        self.convs, self.bns, self.params = [], [], []
        self._conv_of, self._bn_of = {}, {}
        self._dev_state = {}
        self._keepalive = []
        self._collect()

    # ---- static structure ---------------------------------------------------------------------------
    def _param(self, p):
        if p is None:
            return -1
        self.params.append(p)
        return len(self.params) - 1

    def _add_conv(self, mod):
        assert mod.groups == 1 and mod.dilation == (1, 1) and mod.kernel_size[0] == mod.kernel_size[1]
        assert mod.stride[0] == mod.stride[1] and mod.padding[0] == mod.padding[1] == mod.kernel_size[0] // 2
        r = _ConvRec()
        r.mod, r.ks, r.stride = mod, mod.kernel_size[0], mod.stride[0]
        r.cin, r.cout = mod.in_channels, mod.out_channels
        r.coutp = (r.cout + 31) // 32 * 32
        r.taps = r.ks * r.ks
        assert r.cin % 32 == 0, "head convolutions need input channels in multiples of 32"
        r.pw, r.pb = self._param(mod.weight), self._param(mod.bias)
        self._conv_of[id(mod)] = r
        self.convs.append(r)
        return r

    def _add_bn(self, mod, repeat=1):
        assert mod.track_running_stats and mod.affine and mod.momentum is not None
        r = _BNRec()
        r.mod, r.c, r.repeat = mod, mod.num_features, repeat
        r.pw, r.pb = self._param(mod.weight), self._param(mod.bias)
        self._bn_of[id(mod)] = r
        self.bns.append(r)
        return r

    def _collect(self):
        h = self.head
        n_de = len(h.deblocks)
        for stage in h.blocks:
            for blk in stage:
                self._add_conv(blk.conv1.conv1); self._add_bn(blk.bn1)
                self._add_conv(blk.conv2.conv1); self._add_bn(blk.bn2)
                if blk.downsample is not None:
                    self._add_conv(blk.downsample[0].conv1); self._add_bn(blk.downsample[1])
        for sk in h.skip_blocks:
            self._add_conv(sk[0]); self._add_bn(sk[1])
        for de in h.deblocks:
            assert de[0].scale_factor == 2 and de[0].mode == "nearest"
            self._add_conv(de[1]); self._add_bn(de[2])
        stacks = [(s, 1) for s in list(h.pyramid_motion_blocks)[:n_de - 1]]
        stacks += [(h.tq_map_conv, 1), (h.t_map_conf.conf_model, 2), (h.q_map_conf.conf_model, 2)]
        for s, rep in stacks:
            # the reference scores the confidence stacks a second time at temperature 20 on the same (detached)
            # input in training mode (`odom_pred.py:242-258`): same values, but the BatchNorm running statistics
            # move twice per forward -> update_repeat 2
            self._add_conv(s[0]); self._add_bn(s[1], rep)
            self._add_conv(s[3]); self._add_bn(s[4], rep)
            self._add_conv(s[6])
        off = 0
        for r in self.convs:
            r.img_off = off
            off += 4 * r.taps * r.cin * r.coutp + (r.coutp if r.coutp != r.cout else 0)
        self._img_floats = off
        off = 0
        for r in self.convs:
            r.dw_off = off
            off += r.taps * r.cin * r.coutp
        self._dw_floats = off
        off = 0
        for b in self.bns:
            b.off = off
            off += b.c
        self._bn_channels = off

    # ---- per-device buffers: weight images + preparation tables -----------------------------------------
    def _state(self, device):
        sig = tuple(p.data_ptr() for p in self.params)
        st = self._dev_state.get(device)
        if st is not None and st["sig"] == sig:
            return st
        img = torch.empty(self._img_floats, dtype=torch.float32, device=device)
        ent_train = (_PrepEntry * len(self.convs))()
        ent_eval = (_PrepEntry * len(self.convs))()
        for i, r in enumerate(self.convs):
            n = 2 * r.taps * r.cin * r.coutp
            r.img_fwd = img[r.img_off:r.img_off + n]
            r.img_bwd = img[r.img_off + n:r.img_off + 2 * n]
            r.bias_pad = img[r.img_off + 2 * n:r.img_off + 2 * n + r.coutp] if r.coutp != r.cout else None
            for ent, bwd in ((ent_train, True), (ent_eval, False)):
                e = ent[i]
                e.w = r.mod.weight.data_ptr()
                e.img_fwd = r.img_fwd.data_ptr()
                e.img_bwd = r.img_bwd.data_ptr() if bwd else None
                e.bias = r.mod.bias.data_ptr() if r.mod.bias is not None else None
                e.bias_pad = r.bias_pad.data_ptr() if r.bias_pad is not None else None
                e.Cout, e.CoutP, e.Cin, e.taps = r.cout, r.coutp, r.cin, r.taps
        st = {"sig": sig, "img": img,
              "tab_train": _to_device_table(ent_train, device, self._keepalive),
              "tab_eval": _to_device_table(ent_eval, device, self._keepalive)}
        self._dev_state[device] = st
        return st

    def clear(self):
        self._dev_state.clear()

    # ---- forward ---------------------------------------------------------------------------------------
    def _conv_bn(self, run, stats_all, conv_mod, bn_mod, x, relu, residual=None, record=True):
        cv, bn = self._conv_of[id(conv_mod)], self._bn_of[id(bn_mod)]
        B, H, W, cin = x.shape
        assert cin == cv.cin
        Ho, Wo = (H - 1) // cv.stride + 1, (W - 1) // cv.stride + 1
        dev = x.split.device
        y = torch.empty((B, Ho, Wo, cv.cout), dtype=torch.float32, device=dev)
        m = bn.mod
        batch_stats = m.training
        stats = stats_all[run.G * 2 * bn.off:run.G * 2 * (bn.off + bn.c)] if batch_stats else None
        K.conv2d_tc_forward(x.split, cv.img_fwd, cv.cout, cv.ks, cv.stride, bias=cv.mod.bias, out=y, stats=stats,
                            imgs_per_group=run.ipg)
        out = _Act((B, Ho, Wo, cv.cout), z=torch.empty_like(y),
                   split=torch.empty((2, B, Ho, Wo, cv.cout), dtype=torch.float32, device=dev))
        mean_rstd = torch.empty((run.G, bn.c, 2), dtype=torch.float32, device=dev) if record else None
        K.bn_act_forward(y, run.ipg, stats, m.weight, m.bias, m.running_mean, m.running_var, m.num_batches_tracked,
                         m.eps, m.momentum, bn.repeat if batch_stats else 0, residual.z if residual is not None else None,
                         relu, out.z, out.split, mean_rstd)
        if record:
            run.tape.append(("convbn", cv, bn, x, y, mean_rstd, out, residual, relu, batch_stats))
        return out

    def _conv_out(self, run, conv_mod, x, slot, record):
        cv = self._conv_of[id(conv_mod)]
        B, H, W, _ = x.shape
        y = torch.empty((B, H, W, cv.coutp), dtype=torch.float32, device=x.split.device)
        K.conv2d_tc_forward(x.split, cv.img_fwd, cv.coutp, cv.ks, cv.stride,
                            bias=cv.bias_pad if cv.bias_pad is not None else cv.mod.bias, out=y)
        if record:
            run.tape.append(("convout", cv, x, slot))
        return y

    def _stack(self, run, stats_all, seq, x, slot, record):
        """conv3x3-BN-ReLU, conv3x3-BN-ReLU, conv1x1(+bias): `odom_pred_base.py:239-248` and the conf/pyramid stacks"""
        x = self._conv_bn(run, stats_all, seq[0], seq[1], x, True, record=record)
        x = self._conv_bn(run, stats_all, seq[3], seq[4], x, True, record=record)
        return self._conv_out(run, seq[6], x, slot, record)

    def forward(self, x1, x2, ipg, record):
        """x1, x2 [B,C,H,W] NCHW BEV maps of the B pairs -> _Run; outputs (NHWC, channel-padded to 32):
        tq [B,H,W,32] (7 used), t_logit, r_logit [B,H,W,32] (1 used), pyramid preds (7 used), input mask [B,H,W]."""
        h = self.head
        B, Cc, H, W = x1.shape
        assert B % ipg == 0 and (H % 8 == 0) and (W % 8 == 0)
        dev = x1.device
        st = self._state(dev)
        training_any = any(b.mod.training for b in self.bns)
        K.conv2d_multi_prepare(st["tab_train"] if record else st["tab_eval"], len(self.convs))
        run = _Run()
        run.tape, run.ipg, run.G, run.B = [], ipg, B // ipg, B
        run.x_shape = (B, Cc, H, W)
        stats_all = torch.zeros(run.G * 2 * self._bn_channels, dtype=torch.float64, device=dev) if training_any else None

        xin = _Act((B, H, W, 2 * Cc), split=torch.empty((2, B, H, W, 2 * Cc), dtype=torch.float32, device=dev))
        mask = torch.empty((B, H, W), dtype=torch.float32, device=dev)
        K.head_pack_input(x1, x2, xin.split, mask)
        run.xin = xin

        x = xin
        ups = []
        for i, stage in enumerate(h.blocks):
            for blk in stage:
                a = self._conv_bn(run, stats_all, blk.conv1.conv1, blk.bn1, x, True, record=record)
                if blk.downsample is not None:
                    res = self._conv_bn(run, stats_all, blk.downsample[0].conv1, blk.downsample[1], x, False, record=record)
                else:
                    res = x
                x = self._conv_bn(run, stats_all, blk.conv2.conv1, blk.bn2, a, True, residual=res, record=record)
            sk = h.skip_blocks[i]
            ups.append(self._conv_bn(run, stats_all, sk[0], sk[1], x, True, record=record))

        outs_py = []
        n_de = len(h.deblocks)
        for i, de in enumerate(h.deblocks):
            srcs = [x, ups[-(i + 1)]]
            Bc, Hc, Wc, _ = x.shape
            ld = sum(s.shape[3] for s in srcs)
            cat = _Act((Bc, 2 * Hc, 2 * Wc, ld),
                       split=torch.empty((2, Bc, 2 * Hc, 2 * Wc, ld), dtype=torch.float32, device=dev))
            off = 0
            for s in srcs:
                K.upcat_split(s.z, 2, ld, off, cat.split)
                off += s.shape[3]
            if record:
                run.tape.append(("upcat", srcs, cat, 2))
            x = self._conv_bn(run, stats_all, de[1], de[2], cat, True, record=record)
            if h.pred_pyramid_motion and i < n_de - 1:
                outs_py.append(self._stack(run, stats_all, h.pyramid_motion_blocks[i], x, 3 + i, record))
        tq = self._stack(run, stats_all, h.tq_map_conv, x, 0, record)
        tl = self._stack(run, stats_all, h.t_map_conf.conf_model, x, 1, record)
        rl = self._stack(run, stats_all, h.q_map_conf.conf_model, x, 2, record)
        run.outputs = [tq, tl, rl] + outs_py + [mask]
        return run

    # ---- backward ----------------------------------------------------------------------------------------
    @staticmethod
    def _grad_buf(act, dev):
        if act.dz is None:
            act.dz = torch.empty(act.shape, dtype=torch.float32, device=dev)
            return act.dz, False
        return act.dz, True

    def backward(self, run, grads):
        """grads: gradients of run.outputs[:-1] (NHWC, 32 channels; None = no gradient).
        -> (gx1, gx2 [B,C,H,W], [gradient per engine parameter])"""
        dev = run.xin.split.device
        for rec in run.tape:                         # a second backward over the same forward starts clean
            if rec[0] == "convbn":
                rec[3].dz = None
                rec[6].dz = None
                if rec[7] is not None:
                    rec[7].dz = None
            elif rec[0] == "convout":
                rec[2].dz = None
            else:
                rec[2].dz = None
        sums_all = torch.zeros(run.G * 2 * self._bn_channels, dtype=torch.float64, device=dev)
        dW_all = torch.zeros(self._dw_floats, dtype=torch.float32, device=dev)
        # parameter gradients: one flat buffer, weights first (fully written by the finish kernel), then the rest
        w_total = sum(self.params[r.pw].numel() for r in self.convs)
        o_total = sum(p.numel() for i, p in enumerate(self.params)) - w_total
        flat = torch.empty(w_total + o_total, dtype=torch.float32, device=dev)
        flat[w_total:].zero_()
        pg = [None] * len(self.params)
        off = 0
        for r in self.convs:
            n = self.params[r.pw].numel()
            pg[r.pw] = flat[off:off + n].view_as(self.params[r.pw])
            off += n
        for i, p in enumerate(self.params):
            if pg[i] is None:
                pg[i] = flat[off:off + p.numel()].view_as(p)
                off += p.numel()

        for rec in reversed(run.tape):
            kind = rec[0]
            if kind == "convout":
                _, cv, x, slot = rec
                g = grads[slot]
                if g is None:
                    continue
                g = g.contiguous()
                gs = K.conv2d_split(g)
                dx, acc = self._grad_buf(x, dev)
                K.conv2d_tc_backward_data(gs, cv.img_bwd, x.shape, cv.ks, cv.stride, out=dx, accumulate=acc)
                K.conv2d_tc_backward_weight(x.split, gs, cv.ks, cv.stride, cout_real=cv.cout,
                                            scratch=dW_all[cv.dw_off:cv.dw_off + cv.taps * cv.cin * cv.coutp])
                if cv.pb >= 0:
                    B, H, W, _ = x.shape
                    K.bias_grad(g, B * H * W, cv.coutp, cv.cout, pg[cv.pb])
            elif kind == "convbn":
                _, cv, bn, x, y, mean_rstd, out, residual, relu, batch_stats = rec
                if out.dz is None:
                    continue
                gs = torch.empty((2,) + tuple(y.shape), dtype=torch.float32, device=dev)
                dres, dres_acc = (None, False) if residual is None else self._grad_buf(residual, dev)
                sums = sums_all[run.G * 2 * bn.off:run.G * 2 * (bn.off + bn.c)]
                K.bn_act_backward(out.dz, out.z, y, run.ipg, mean_rstd, bn.mod.weight, relu, batch_stats, sums, gs, dres,
                                  dres_acc, pg[bn.pw], pg[bn.pb], pg[cv.pb] if cv.pb >= 0 else None)
                out.dz = None
                dx, acc = self._grad_buf(x, dev)
                K.conv2d_tc_backward_data(gs, cv.img_bwd, x.shape, cv.ks, cv.stride, out=dx, accumulate=acc)
                K.conv2d_tc_backward_weight(x.split, gs, cv.ks, cv.stride, cout_real=cv.cout,
                                            scratch=dW_all[cv.dw_off:cv.dw_off + cv.taps * cv.cin * cv.coutp])
            else:
                _, srcs, cat, up = rec
                if cat.dz is None:
                    continue
                off_c = 0
                for s in srcs:
                    dz, acc = self._grad_buf(s, dev)
                    K.upcat_backward(cat.dz, s.shape, up, cat.shape[3], off_c, dz, acc)
                    off_c += s.shape[3]
                cat.dz = None

        ent = (_FinishEntry * len(self.convs))()
        for i, r in enumerate(self.convs):
            e = ent[i]
            e.dW = dW_all.data_ptr() + 4 * r.dw_off
            e.gw = pg[r.pw].data_ptr()
            e.Cout, e.CoutP, e.Cin, e.taps = r.cout, r.coutp, r.cin, r.taps
        K.conv2d_multi_wgrad_finish(_to_device_table(ent, dev, self._keepalive), len(self.convs))

        B, Cc, H, W = run.x_shape
        g1 = torch.empty((B, Cc, H, W), dtype=torch.float32, device=dev)
        g2 = torch.empty_like(g1)
        if run.xin.dz is None:
            g1.zero_(); g2.zero_()
        else:
            K.head_unpack_grad(run.xin.dz, g1, g2)
            run.xin.dz = None
        return g1, g2, pg


class _HeadTrunkFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, engine, ipg, x1, x2, *params):
        record = any(ctx.needs_input_grad)
        run = engine.forward(x1.contiguous(), x2.contiguous(), ipg, record)
        ctx.engine, ctx.run = engine, run
        outs = tuple(run.outputs)
        # the returned tensors get this node as grad_fn: keeping them on ctx.run would be a reference cycle that only
        # the cycle collector frees, and a graph that lingers keeps the parameters' AccumulateGrad nodes (and the
        # stream they were created on) alive into the next capture
        run.outputs = None
        ctx.mark_non_differentiable(outs[-1])
        if not record:
            run.tape = None
        return outs

    @staticmethod
    def backward(ctx, *grads):
        g1, g2, pg = ctx.engine.backward(ctx.run, list(grads[:-1]))
        return (None, None, g1, g2, *pg)


def head_trunk(engine, x1, x2, imgs_per_group):
    """Differentiable call: (tq, t_logit, r_logit, pyramid..., mask) NHWC, see HeadTrunkEngine.forward."""
    return _HeadTrunkFn.apply(engine, imgs_per_group, x1, x2, *engine.params)
