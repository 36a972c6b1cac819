"""Build the CUDA library in-tree: neuralmelting_b200/libnm_b200.so (sm_100a only).

nvcc cross-compiles without a GPU. The .so is git-ignored but travels to the GPU box.
"""
import glob
import hashlib
import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
LIB = os.path.join(_HERE, "libnm_b200.so")
_STAMP = os.path.join(_HERE, ".libnm_b200.stamp")

EXTRA = os.environ.get("NM_NVCC_EXTRA", "").split()
NVCC_FLAGS = EXTRA + [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-O2", "-diag-suppress=177,550",      # unused helpers / variables of the other translation units
]


def sources():
    return sorted(glob.glob(os.path.join(_HERE, "csrc", "*.cu")) + glob.glob(os.path.join(_HERE, "csrc", "*.cpp")))


def _digest():
    h = hashlib.sha256()
    files = sources() + sorted(glob.glob(os.path.join(_HERE, "csrc", "*.cuh"))) + [os.path.join(_ROOT, "include", "nm_b200.h")]
    for f in files:
        h.update(f.encode())
        with open(f, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    """compile every .cu / .cpp under csrc/ into one shared library; returns its path"""
    dig = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(_STAMP) and open(_STAMP).read().strip() == dig:
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    objs = []
    procs = []
    bdir = os.path.join(_HERE, "build")
    os.makedirs(bdir, exist_ok=True)
    units = []
    for src in sources():
        if os.path.basename(src) == "nm_engine.cu":      # NM_TU: one unit per thread count + the host unit, built in parallel
            units += [(src, ["-DNM_TU=%d" % t], ".tu%d" % t) for t in (1025, 1024, 512, 256, 0)]     # 1025: the helper-capable 1024-thread cycle kernel
        else:
            units.append((src, [], ""))
    for src, defs, tag in units:
        obj = os.path.join(bdir, os.path.basename(src) + tag + ".o")
        cmd = [nvcc] + NVCC_FLAGS + defs + ["-I", os.path.join(_ROOT, "include"), "-I", os.path.join(_HERE, "csrc"),
                                             "-c", src, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), file=sys.stderr)
        procs.append((src + tag, subprocess.Popen(cmd)))
        objs.append(obj)
    for src, p in procs:
        if p.wait() != 0:
            raise RuntimeError("nvcc failed on %s" % src)
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs
    subprocess.check_call(link)
    with open(_STAMP, "w") as fh:
        fh.write(dig)
    return LIB


if __name__ == "__main__":
    print(build(force="-f" in sys.argv, verbose="-v" in sys.argv))
