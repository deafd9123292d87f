"""Builds the reference's OWN CUDA deformable-convolution extension (ops/dcn/src/*, the same-box GPU baseline of SURVEY 8c /
BASELINE.md step 5) into baseline/_ref/ -- git-ignored, not gpurun-ignored, so the .so travels to the GPU box.

Nothing of the reference is copied into the repo: the two source files are copied to a scratch directory under /tmp (the reference
tree is read-only), patched there with the 6-replacement `.type()` -> `.scalar_type()` fix torch >= 2.x needs
(deform_conv_cuda_kernel.cu:258,352,450,780,812,845), and compiled for sm_100a with the reference's own
setup.py.  Run in the build container:  python baseline/build_ref_dcn.py     (a few minutes; skipped when the .so is already there)

The product never loads this module; tools/bench_dcn.py (--ref) and tests/test_ref_extension_gpu.py do, as the baseline / checker.
"""
import glob
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "baseline", "_ref")
REF = os.environ.get("CDFO_REFERENCE_ROOT", "/root/reference")


def built():
    return sorted(glob.glob(os.path.join(OUT, "deform_conv_cuda*.so")))


def main(force=False):
    if built() and not force:
        print("already built:", built()[0])
        return 0
    src = os.path.join(REF, "ops", "dcn")
    if not os.path.isdir(src):
        print("reference tree not present at %s: nothing to build" % REF)
        return 0
    os.makedirs(OUT, exist_ok=True)
    work = tempfile.mkdtemp(prefix="cdfo_refdcn_")
    try:
        dst = os.path.join(work, "dcn")
        shutil.copytree(src, dst, ignore=shutil.ignore_patterns("__pycache__"))
        ku = os.path.join(dst, "src", "deform_conv_cuda_kernel.cu")
        text = open(ku).read()
        n = text.count('.type(), "')
        open(ku, "w").write(text.replace('.type(), "', '.scalar_type(), "'))
        print("patched %d x .type() -> .scalar_type() in the scratch copy" % n)
        env = dict(os.environ, TORCH_CUDA_ARCH_LIST="10.0a", MAX_JOBS="4")
        rc = subprocess.call([sys.executable, "setup.py", "build_ext", "--inplace"], cwd=dst, env=env)
        if rc != 0:
            print("build failed (rc %d)" % rc)
            return rc
        for so in glob.glob(os.path.join(dst, "deform_conv_cuda*.so")):
            shutil.copy2(so, OUT)
            print("->", os.path.join(OUT, os.path.basename(so)), os.path.getsize(so), "bytes")
    finally:
        shutil.rmtree(work, ignore_errors=True)
    return 0


if __name__ == "__main__":
    sys.exit(main(force="--force" in sys.argv))
