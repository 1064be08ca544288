"""Sparse 3-D encoder with covariance decoder (`rslo/models/middle.py:37-245`)."""
import numpy as np
import torch
from torch import nn
from torch.nn import functional as F

from .. import kernels as K
from ..layers import encoder_engine
from ..layers import sparse3d as spconv
from ..layers.sparse3d import IndexEntry
from ..torchplus import Empty, change_default_args

REGISTERED_MIDDLE_CLASSES = {}


def register_middle(cls, name=None):
    name = cls.__name__ if name is None else name
    assert name not in REGISTERED_MIDDLE_CLASSES, f"exist class: {REGISTERED_MIDDLE_CLASSES}"
    REGISTERED_MIDDLE_CLASSES[name] = cls
    return cls


def get_middle_class(name):
    assert name in REGISTERED_MIDDLE_CLASSES, f"available class: {REGISTERED_MIDDLE_CLASSES}"
    return REGISTERED_MIDDLE_CLASSES[name]


_GEOMS = [("conv3d2", (3, 3, 3), (2, 2, 2), (1, 1, 1)), ("conv3d3", (3, 3, 3), (2, 2, 2), (1, 1, 1)),
          ("conv3d4", (3, 3, 3), (2, 2, 2), (0, 1, 1)), ("conv3d5", (3, 1, 1), (2, 1, 1), (0, 0, 0))]
_SUBM_KEYS = ["subm0", "subm1", "subm2", "subm3"]


def _enqueue_frame_tables(indices, n, sparse_shape, table0, caps, n_dev=None):
    """Enqueue every table kernel of one frame (no host round trip: the row counts of the deeper levels stay
    on the device while the next level is built).  -> (levels, raw tables)."""
    tab = table0 if table0 is not None else K.site_table_build(indices, n, sparse_shape)
    lv = [dict(idx=indices, cap=n, ndev=n_dev, shape=list(sparse_shape), tab=tab)]
    raw = {}
    for li, (key, ks, st, pd) in enumerate(_GEOMS):
        cur = lv[-1]
        raw[_SUBM_KEYS[li]] = K.subm_table(cur["idx"], cur["cap"], cur["tab"], (3, 3, 3), n_dev=cur["ndev"])
        cap = None if caps is None else caps[li]
        if cap is None:
            # typical LiDAR levels shrink; overflow is detected by the caller and retried exactly
            cells = int(np.prod(K.out_shape_of(cur["shape"], ks, st, pd)))
            cap = max(1, min(cur["cap"], cells))
        otab, oc, ndev2, nbr, nbr_inv = K.strided_table(cur["idx"], cur["cap"], cur["shape"], ks, st, pd,
                                                        n_dev=cur["ndev"], out_cap=cap)
        raw[key] = (nbr, nbr_inv)
        lv.append(dict(idx=oc, cap=cap, ndev=ndev2[:1], shape=list(otab.shape), tab=otab, ndev2=ndev2))
    return lv, raw


class PendingTables:
    """Index tables of T frames that have been ENQUEUED (capacity-sized, row counts still on the device) together with
    the asynchronous device->host copy of every count.  `finish()` waits for that copy (a no-op wait when it is called a
    step later), retries the rare frame whose level outgrew its capacity, and returns per frame
    (levels, raw tables, row counts per level).  Splitting the two lets `net.prepare()` return without blocking the
    training thread on the preparation stream."""

    def __init__(self, frames, sparse_shape):
        self.frames, self.sparse_shape = frames, sparse_shape
        self.pend = [_enqueue_frame_tables(idx, n, sparse_shape, tab, None, nd) for idx, n, tab, nd in frames]
        self.dev = frames[0][0].device
        self.counts_dev = torch.stack([self._counts_of(lv, fr[3], fr[1]) for (lv, _), fr in zip(self.pend, frames)])
        self.counts_host = torch.empty(self.counts_dev.shape, dtype=torch.int32, pin_memory=self.dev.type == "cuda")
        self.counts_host.copy_(self.counts_dev, non_blocking=True)
        self.event = None
        if self.dev.type == "cuda":
            self.event = torch.cuda.Event()
            self.event.record(torch.cuda.current_stream(self.dev))

    def _counts_of(self, lv, nd, n):
        n0 = nd.reshape(1).to(torch.int32) if nd is not None else torch.tensor([n], dtype=torch.int32, device=self.dev)
        return torch.cat([torch.stack([n0[0], n0[0]])[None], torch.stack([l["ndev2"] for l in lv[1:]])])

    def device_tensors(self):
        """every tensor the enqueue step allocated (for record_stream when it ran on a side stream)"""
        out = [self.counts_dev]
        for lv, raw in self.pend:
            for l in lv:
                out += [l["idx"], l.get("ndev"), l.get("ndev2"), l["tab"].cells, l["tab"].perm]
            for v in raw.values():
                out += list(v) if isinstance(v, tuple) else [v]
        return [t for t in out if t is not None]

    def finish(self):
        if self.event is not None:
            self.event.synchronize()
        counts = self.counts_host.numpy()
        out = []
        for f, (lv, raw) in enumerate(self.pend):
            c = counts[f]
            fr = self.frames[f]
            while not (c[1:, 1] <= [l["cap"] for l in lv[1:]]).all():       # rare: a level grew; redo this frame
                caps = [max(int(v), 1) * 2 for v in c[1:, 1]]
                lv, raw = _enqueue_frame_tables(fr[0], fr[1], self.sparse_shape, fr[2], caps, fr[3])
                c = self._counts_of(lv, fr[3], fr[1]).cpu().numpy()
            out.append((lv, raw, [int(v) for v in c[:, 0]]))
        return out


def _frames_tables(frames, sparse_shape):
    """frames: list of (indices, n, table0, n_dev): n rows are live, or - when n_dev (device int) is given -
    n is only the capacity and the live count is still on the device (voxeliser output that has not been
    synchronised).  All frames are enqueued first and ONE device->host copy brings back every count.
    -> per frame (levels, raw tables, row counts per level)."""
    return PendingTables(frames, sparse_shape).finish()


def build_frame_tables(indices, n, sparse_shape, table0=None):
    """All index tables of SpMiddleFHDWithCov2_3 for one frame.  Returns {indice_key: IndexEntry}."""
    return build_tables_batched([(indices, n, table0, None)], sparse_shape)[0]


def build_tables_batched(frames, sparse_shape, pending=None):
    """Index tables for T frames that share one pass through the encoder: per-frame tables are appended
    row-wise (row indices of frame f shifted by the rows before it).  -> ({indice_key: IndexEntry}, meta)."""
    per = pending.finish() if pending is not None else _frames_tables(frames, sparse_shape)
    T = len(per)
    ns = [[p[2][l] for p in per] for l in range(5)]                    # rows per level per frame
    offs = [[int(sum(ns[l][:f])) for f in range(T)] for l in range(5)]
    tot = [int(sum(ns[l])) for l in range(5)]

    def level_frames(l):
        return [(offs[l][f], ns[l][f], per[f][0][l]["tab"], per[f][0][l]["idx"]) for f in range(T)]

    # all 12 appended tables of the step in one launch
    jobs = [([p[1][_SUBM_KEYS[li]] for p in per], ns[li], offs[li]) for li in range(4)]
    for li, (key, ks, st, pd) in enumerate(_GEOMS):
        jobs.append(([p[1][key][0] for p in per], ns[li + 1], offs[li]))         # out rows -> in rows
        jobs.append(([p[1][key][1] for p in per], ns[li], offs[li + 1]))         # in rows -> out rows
    cats = [j[0][0] for j in jobs] if T == 1 else K.table_concat_many(jobs)
    cats = iter(cats)
    subm_tabs = [next(cats) for _ in range(4)]
    strided_tabs = [(next(cats), next(cats)) for _ in _GEOMS]

    entries = {}
    for li in range(4):
        nbr = subm_tabs[li]
        e = IndexEntry("subm", nbr, nbr, tot[li], tot[li], per[0][0][li]["idx"] if T == 1 else None,
                       per[0][0][li]["shape"], per[0][0][li]["tab"] if T == 1 else None, True)
        e.seg_in = e.seg_out = ns[li]
        entries[_SUBM_KEYS[li]] = e
    for li, (key, ks, st, pd) in enumerate(_GEOMS):
        nbr, nbr_inv = strided_tabs[li]
        o0, i0 = per[0][0][li + 1], per[0][0][li]
        e = IndexEntry("strided", nbr, nbr_inv, tot[li], tot[li + 1], o0["idx"] if T == 1 else None, o0["shape"],
                       o0["tab"] if T == 1 else None, False)
        e.in_indices, e.in_shape, e.in_table = (i0["idx"] if T == 1 else None), i0["shape"], (i0["tab"] if T == 1 else None)
        e.seg_in, e.seg_out = ns[li], ns[li + 1]
        e.in_frames, e.out_frames = level_frames(li), level_frames(li + 1)
        entries[key] = e
    # decoder keys are new names for site sets that already have tables (middle.py:181-213)
    entries["dsubm3"] = entries["subm1"]
    entries["dsubm2"] = entries["subm0"]
    entries["dsubm1"] = entries["subm0"]
    return entries, {"rows": ns, "offsets": offs, "frames0": level_frames(0)}


@register_middle
class SpMiddleFHDWithCov2_3(nn.Module):
    def __init__(self, output_shape, use_GN=False, sync_bn=False, bn_type="None", use_leakyReLU=False,
                 relu_type="ReLU", num_input_features=128, num_filters_down1=(64,), num_filters_down2=(64, 64),
                 name="SpMiddleFHDWithConf"):
        super().__init__()
        assert bn_type in ["None", "BN", "SyncBN", "SemiGlobalSyncBN", "MaskSyncBN"]
        assert relu_type in ["", "ReLU", "LeakyReLU", "PReLU"]
        self.name = name
        if bn_type != "None":
            # the shipped configs (bn_type "None") are the scope of this build
            raise NotImplementedError("SpMiddleFHDWithCov2_3: only bn_type 'None' is built")
        BatchNorm1d = Empty
        SpConv3d = change_default_args(bias=True)(spconv.SparseConv3d)
        SubMConv3d = change_default_args(bias=True)(spconv.SubMConv3d)
        ConvTranspose3d = change_default_args(bias=True)(spconv.SparseInverseConv3d)
        if use_leakyReLU or relu_type == "LeakyReLU":
            self.relu = nn.LeakyReLU
        else:
            raise NotImplementedError("SpMiddleFHDWithCov2_3: only LeakyReLU is built (shipped configs)")

        sparse_shape = np.array(output_shape[1:4]) + [1, 0, 0]
        self.sparse_shape = sparse_shape
        self.voxel_output_shape = output_shape

        self.middle_conv = spconv.SparseSequential(
            SubMConv3d(num_input_features, 16, 3, indice_key="subm0"), BatchNorm1d(16), self.relu(),
            SubMConv3d(16, 16, 3, indice_key="subm0"), BatchNorm1d(16), self.relu(),
            SpConv3d(16, 32, 3, 2, padding=1, indice_key="conv3d2"), BatchNorm1d(32), self.relu(),
            SubMConv3d(32, 32, 3, indice_key="subm1"), BatchNorm1d(32), self.relu(),
            SubMConv3d(32, 32, 3, indice_key="subm1"), BatchNorm1d(32), self.relu(),
            SpConv3d(32, 64, 3, 2, padding=1, indice_key="conv3d3"), BatchNorm1d(64), self.relu(),
        )
        self.middle_conv_tail = spconv.SparseSequential(
            SubMConv3d(64, 64, 3, indice_key="subm2"), BatchNorm1d(64), self.relu(),
            SubMConv3d(64, 64, 3, indice_key="subm2"), BatchNorm1d(64), self.relu(),
            SubMConv3d(64, 64, 3, indice_key="subm2"), BatchNorm1d(64), self.relu(),
            SpConv3d(64, 64, 3, 2, padding=[0, 1, 1], indice_key="conv3d4"), BatchNorm1d(64), self.relu(),
            SubMConv3d(64, 64, 3, indice_key="subm3"), BatchNorm1d(64), self.relu(),
            SubMConv3d(64, 64, 3, indice_key="subm3"), BatchNorm1d(64), self.relu(),
            SubMConv3d(64, 64, 3, indice_key="subm3"), BatchNorm1d(64), self.relu(),
            SpConv3d(64, 64, (3, 1, 1), (2, 1, 1), indice_key="conv3d5"), BatchNorm1d(64), self.relu(),
        )
        self.middle_cov_deconv = spconv.SparseSequential(
            ConvTranspose3d(64, 32, 3, indice_key="conv3d3"), nn.BatchNorm1d(32), self.relu(),
            SubMConv3d(32, 32, 3, indice_key="dsubm3"), nn.BatchNorm1d(32), self.relu(),
            ConvTranspose3d(32, 16, 3, indice_key="conv3d2"), nn.BatchNorm1d(16), self.relu(),
            SubMConv3d(16, 16, 3, indice_key="dsubm2"), nn.BatchNorm1d(16), self.relu(),
            SubMConv3d(16, 16, 3, indice_key="dsubm2"), nn.BatchNorm1d(16), self.relu(),
            SubMConv3d(16, 7, 3, indice_key="dsubm1"),
        )
        self.max_batch_size = 6

    def forward(self, voxel_features, coors, batch_size, table0=None):
        rets, covs, _, _ = self.forward_frames([voxel_features], [coors], batch_size, [table0])
        return rets[0], covs[0]

    def prepare_frames(self, voxel_features, coors, tables=None, n_devs=None):
        """Index tables of all frames (+ the single device->host copy of the row counts) and the inputs
        trimmed to their live rows.  Pure index work: can run ahead of time / on another stream."""
        T = len(voxel_features)
        tables = tables if tables is not None else [None] * T
        n_devs = n_devs if n_devs is not None else [None] * T
        return self.prepare_frames_begin(voxel_features, coors, tables, n_devs)()

    def prepare_frames_begin(self, voxel_features, coors, tables=None, n_devs=None):
        """Enqueue the table kernels and the asynchronous copy of the row counts; -> a callable that finishes the job
        (waits for the counts, appends the frames' tables, trims the inputs) and returns what prepare_frames returns.
        The callable carries `.pending` (PendingTables)."""
        T = len(voxel_features)
        tables = tables if tables is not None else [None] * T
        n_devs = n_devs if n_devs is not None else [None] * T
        coors = [c.int().contiguous() for c in coors]
        shape = [int(s) for s in self.sparse_shape]
        frames = [(coors[t], int(voxel_features[t].shape[0]), tables[t], n_devs[t]) for t in range(T)]
        pending = PendingTables(frames, shape)

        def finish():
            entries, meta = build_tables_batched(frames, shape, pending)
            feats = [voxel_features[t][:meta["rows"][0][t]] for t in range(T)]
            cs = [coors[t][:meta["rows"][0][t]] for t in range(T)]
            return {"entries": entries, "meta": meta, "features": feats, "coors": cs}
        finish.pending = pending
        return finish

    def forward_frames(self, voxel_features, coors, batch_size, tables=None, n_devs=None, prepared=None):
        """The reference calls the encoder once per frame (`voxel_odom_net.py:423-428`); here the T frames
        of an example share ONE pass: their rows are concatenated, every sparse convolution runs once on
        T times the rows (frames never mix: each frame's tables only reference its own rows), batch
        statistics of the covariance decoder's BatchNorm1d stay per frame.
        `n_devs[t]` (device int) marks frame t's inputs as capacity-sized with the live row count still on the
        device; the counts of all frames and levels come back in one copy.
        -> ([bev_t], [cov_t], [features_t], [coors_t]) with the inputs trimmed to their live rows."""
        assert batch_size == 1, "Only support batch_size=1 for now"
        if prepared is None:
            prepared = self.prepare_frames(voxel_features, coors, tables, n_devs)
        entries, meta = prepared["entries"], prepared["meta"]
        voxel_features, coors = prepared["features"], prepared["coors"]
        T = len(voxel_features)
        feats = voxel_features[0] if T == 1 else torch.cat(voxel_features, dim=0)
        ret = spconv.SparseConvTensor(feats, coors[0] if T == 1 else None, self.sparse_shape, batch_size,
                                      table=entries["subm0"].out_table)
        ret.indice_dict = entries
        ret.seg, ret.frames = meta["rows"][0], meta["frames0"]
        if encoder_engine.USE_ENGINE and feats.is_cuda:
            # one autograd node for the 25 layers (layers/encoder_engine.py): same kernels, a fraction of the host time
            eng = self.__dict__.get("_engine")
            if eng is None:
                eng = self.__dict__["_engine"] = encoder_engine.SparseEncoderEngine(self)
            tail, cov = encoder_engine.encode(eng, feats, entries, meta["rows"][0])
            e5 = entries["conv3d5"]
            ret = spconv.SparseConvTensor(tail, None, e5.out_shape, batch_size)
            ret.n, ret.seg, ret.frames = e5.n_out, e5.seg_out, e5.out_frames
        else:
            ret0 = self.middle_conv(ret)
            ret = self.middle_conv_tail(ret0)
            cov = self.middle_cov_deconv(ret0).features
        cov = torch.cat([F.elu(cov[:, :3]) + 1 + 1e-6, cov[:, 3:]], dim=1)     # middle.py:237
        covs = [cov] if T == 1 else list(torch.split(cov, meta["rows"][0]))
        return ret.dense_frames(), covs, voxel_features, coors
