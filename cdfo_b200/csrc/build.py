"""Builds cdfo_b200/libcdfo_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
OUT = os.path.join(PKG, "libcdfo_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC", "-shared", "-cudart", "static",
]


def sources():
    return sorted(glob.glob(os.path.join(HERE, "*.cu")))


def stale():
    if not os.path.isfile(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = sources() + glob.glob(os.path.join(HERE, "*.cuh")) + glob.glob(
        os.path.join(os.path.dirname(PKG), "include", "*.h")) + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not stale():
        return OUT
    objs = []
    odir = os.path.join(HERE, "_obj")
    os.makedirs(odir, exist_ok=True)
    procs = []
    for src in sources():
        obj = os.path.join(odir, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        if (not force and os.path.isfile(obj) and os.path.getmtime(obj) > os.path.getmtime(src)
                and all(os.path.getmtime(obj) > os.path.getmtime(h) for h in glob.glob(os.path.join(HERE, "*.cuh")))):
            continue
        cmd = [NVCC] + [f for f in FLAGS if f != "-shared"] + ["-c", src, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    bad = False
    for src, p in procs:
        out, _ = p.communicate()
        if out.strip():
            print(out)
        if p.returncode != 0:
            print("nvcc failed on", src, file=sys.stderr)
            bad = True
    if bad:
        raise RuntimeError("nvcc compilation failed")
    subprocess.check_call([NVCC] + FLAGS + objs + ["-o", OUT])
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
