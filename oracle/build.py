"""Build the oracle's C restatement (oracle/c/oracle.c -> oracle/_build/liboracle.so) and, when the
reference tree is mounted (this container only), the reference's own chamfer extension into
oracle/_ref/ (sources compiled where they lie under /root/reference; nothing is copied).

TEST INFRASTRUCTURE ONLY.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def build_oracle(force=False):
    src = os.path.join(HERE, "c", "oracle.c")
    out_dir = os.path.join(HERE, "_build")
    os.makedirs(out_dir, exist_ok=True)
    out = os.path.join(out_dir, "liboracle.so")
    if not force and os.path.exists(out) and os.path.getmtime(out) >= os.path.getmtime(src):
        return out
    # -ffp-contract=off: the oracle states every fma explicitly (the NN distance).
    cmd = ["gcc", "-O2", "-std=c99", "-fPIC", "-shared", "-ffp-contract=off", "-mfma", "-o", out, src, "-lm"]
    subprocess.check_call(cmd)
    return out


def build_ref(force=False):
    """JIT-build the reference's chamfer extension, unmodified, into oracle/_ref/ (needs the
    reference tree).  Returns the .so path or None when /root/reference is absent."""
    out_dir = os.path.join(HERE, "_ref")
    so = os.path.join(out_dir, "cd_ref.so")
    src_dir = os.path.join(REF, "thirdparty", "chamfer_distance")
    if not os.path.isdir(src_dir):
        return so if os.path.exists(so) else None
    if os.path.exists(so) and not force:
        return so
    os.makedirs(out_dir, exist_ok=True)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    from torch.utils.cpp_extension import load
    load(name="cd_ref",
         sources=[os.path.join(src_dir, "chamfer_distance.cpp"),
                  os.path.join(src_dir, "chamfer_distance.cu")],
         build_directory=out_dir, verbose=False, with_cuda=True)
    return so if os.path.exists(so) else None


if __name__ == "__main__":
    print(build_oracle(force="--force" in sys.argv))
    if "--ref" in sys.argv:
        print(build_ref(force="--force" in sys.argv))
