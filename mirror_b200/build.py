"""Compile mirror_b200/csrc/*.cu for sm_100a into the in-tree C-ABI library
``mirror_b200/libmirror_b200.so`` (nvcc cross-compiles without a GPU)."""
import glob
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
VARIANT = os.environ.get("MIRROR_B200_VARIANT", "")  # e.g. "transpose": A/B builds of kernel options, never the default
OUT = os.path.join(HERE, f"libmirror_b200{'_' + VARIANT if VARIANT else ''}.so")
OBJ = os.path.join(HERE, "csrc", "_obj" + ("_" + VARIANT if VARIANT else ""))
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC", "--use_fast_math", "-Xptxas", "-v"]
# --use_fast_math would change erff/expf accuracy in the loss math; keep IEEE-ish defaults instead.
FLAGS.remove("--use_fast_math")
FLAGS += os.environ.get("MIRROR_B200_EXTRA_NVCC_FLAGS", "").split()  # e.g. -DMIRROR_FLASH_TRACE (tools/flash_trace.py); part of the digest


def _digest(paths):
    h = hashlib.sha256()
    for p in sorted(paths):
        with open(p, "rb") as f:
            h.update(os.path.basename(p).encode())  # location-independent: the tree is copied to other paths (GPU boxes)
            h.update(f.read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    vdir = os.path.join(HERE, "csrc_variants", VARIANT)
    if VARIANT and os.path.isdir(vdir):  # experiment builds: same-named files override the product sources
        over = {os.path.basename(f): f for f in glob.glob(os.path.join(vdir, "*.cu"))}
        srcs = [over.get(os.path.basename(f), f) for f in srcs]
    hdrs = sorted(glob.glob(os.path.join(CSRC, "*.cuh"))) + [os.path.join(HERE, "..", "include", "mirror_b200.h")]
    os.makedirs(OBJ, exist_ok=True)
    stamp = os.path.join(OBJ, "stamp")
    dig = _digest(srcs + hdrs)
    if not force and os.path.exists(OUT) and os.path.exists(stamp) and open(stamp).read() == dig:
        return OUT

    def cc(src):
        obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
        r = subprocess.run([NVCC, *FLAGS, "-I", CSRC, "-c", src, "-o", obj], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(cc, srcs))
    r = subprocess.run([NVCC, "-shared", "-o", OUT, *objs, "-lcudart"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(stamp, "w") as f:
        f.write(dig)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
