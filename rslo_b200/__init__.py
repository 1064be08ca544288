"""rslo_b200 — B200-native implementation of RSLO's per-frame-pair hot path.

Host code mirrors the reference's `rslo.models` / `rslo.layers` / builder surface; compute runs in
hand-written sm_100a CUDA kernels behind the C ABI in include/rslo_b200.h (rslo_b200/_C).
"""
__version__ = "0.1.0"
