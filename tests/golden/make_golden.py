"""Generate tests/golden/pair_*.npz by running the REFERENCE's own network (imported from
/root/reference through oracle/ref_shim.py) on seeded synthetic frame pairs.  Runs only in the build
container (the reference tree is not on the GPU box); the fixtures it writes are committed.

    python tests/golden/make_golden.py

What is reference code here: builders, VFE, middle.py layer list, the whole head, tq-map geometry,
create_loss, the consistency loss, SVDHead, the adaptive L2 losses.  What is the oracle's
restatement underneath (un-vendored dependencies): spconv voxeliser / rulebook /
gather-conv, kornia quaternion conversions, CPU nearest neighbour (pinned separately against the
reference's chamfer extension).
Weights: oracle.net.fill_weights(seed) — a per-key seeded fill so every implementation gets the
same values regardless of construction order.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import net as onet  # noqa: E402
from oracle import ref_shim  # noqa: E402
from rslo_b200.data import synthetic  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
GRAD_KEYS = ["middle_feature_extractor.middle_conv.0.weight", "middle_feature_extractor.middle_conv_tail.21.bias",
             "middle_feature_extractor.middle_cov_deconv.15.weight", "odom_predictor.blocks.0.0.conv1.conv1.weight",
             "odom_predictor.tq_map_conv.6.bias", "odom_predictor.t_map_conf.conf_model.6.weight",
             "_rotation_loss.alpha"]

CASES = {
    # name: (pair seed, beams, azimuth steps, n frames, global step, weight seed)
    "small_eval": (0, 16, 600, 2, 2000, 11),
    "small_train": (0, 16, 600, 2, 2000, 11),
    "small_train_warm": (1, 16, 600, 2, 100, 12),       # step <= 1500: identity pose, 5 ICP iterations
    "small_train_t3": (2, 16, 400, 3, 2000, 13),        # seq_length 3 -> 3 pairs (train prototxt)
    "full_eval": (0, 64, 1875, 2, 2000, 11),            # BASELINE config C1/C2: 120k-pt pair
    "full_train": (3, 64, 1875, 2, 2000, 14),           # BASELINE config C3: 120k-pt pair, fwd + bwd
}


def grad_sample(g):
    """Large gradients are stored as a strided sample (<= ~4096 values) to keep fixtures small."""
    g = np.asarray(g).ravel()
    return g[::max(1, g.size // 4096)]


def make_frames(seed, beams, n_az, n_frames):
    a, b, _ = synthetic.make_pair(seed, n_beams=beams, n_az=n_az)
    frames = [a, b]
    if n_frames == 3:
        c = synthetic.make_pair(seed, n_beams=beams, n_az=n_az, delta_t=(2.0, 0.05, 0.0),
                                delta_ypr_deg=(2.0, 0.1, 0.2))[1]
        frames.append(c)
    return frames


def example_of(frames, vg):
    ex = {"voxels": [], "num_points": [], "coordinates": [], "num_voxels": []}
    for pts in frames:
        r = vg.generate(pts, 40000)
        n = len(r["coordinates"])
        ex["voxels"].append(torch.from_numpy(r["voxels"]))
        ex["num_points"].append(torch.from_numpy(r["num_points_per_voxel"]))
        ex["coordinates"].append(torch.from_numpy(np.concatenate([np.zeros((n, 1), np.int32), r["coordinates"]], 1)))
        ex["num_voxels"].append(torch.tensor([[n]], dtype=torch.int64))
    B = len(frames) * (len(frames) - 1) // 2
    ex["tq_maps"] = [None]
    ex["icp_odometry"] = torch.zeros(B, 7)
    return ex


def run_case(name, rnet, vg):
    seed, beams, n_az, T, step, wseed = CASES[name]
    frames = make_frames(seed, beams, n_az, T)
    onet.fill_weights(rnet, wseed)
    rnet.global_step.fill_(step)
    training = "train" in name
    out = {}
    if training:
        rnet.train()
        rnet.zero_grad()
        ret = rnet(example_of(frames, vg))
        ret["loss"].sum().backward()
        for k in ("loss", "translation_loss", "rotation_loss", "pyramid_loss", "C_loss"):
            out[k] = ret[k].detach().numpy().reshape(-1)
        out["pose"] = torch.cat([ret["translation_preds"], ret["rotation_preds"]], -1).numpy()
        params = dict(rnet.named_parameters())
        for k in GRAD_KEYS:
            g = params[k].grad
            out["grad:" + k] = grad_sample(g.numpy()) if g is not None else np.zeros(0, np.float32)
        out["t_conf_sum"] = ret["t_conf"].sum(dim=(1, 2, 3)).numpy()
        out["tq_map_g_abs_sum"] = ret["tq_map_g"].abs().sum(dim=(1, 2, 3)).numpy()
    else:
        rnet.eval()
        with torch.no_grad():
            ret = rnet(example_of(frames, vg))
        out["pose"] = torch.cat([ret["translation_preds"], ret["rotation_preds"]], -1).numpy()
        out["tq_map_g_abs_sum"] = ret["tq_map_g"].abs().sum(dim=(1, 2, 3)).numpy()
        out["tq_map_g_sample"] = ret["tq_map_g"][:, :, ::8, ::8].numpy()
        out["t_conf_sample"] = ret["t_conf"][:, :, ::8, ::8].numpy()
        out["cov0_sample"] = ret["middle_conf_preds"][0][::97].numpy()
        out["n_voxels"] = np.array([v.shape[0] for v in ret["voxel_features"]])
    out["meta"] = np.array([seed, beams, n_az, T, step, wseed])
    np.savez_compressed(os.path.join(HERE, f"pair_{name}.npz"), **out)
    print(name, {k: (v.shape if v.size > 8 else v.ravel()) for k, v in out.items() if not k.startswith("grad:")})


def main():
    torch.set_num_threads(os.cpu_count())
    rnet, vg = ref_shim.build_reference_net(testing=True, seed=7)
    names = sys.argv[1:] or list(CASES)
    for name in names:
        run_case(name, rnet, vg)


if __name__ == "__main__":
    main()
