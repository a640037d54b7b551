#!/usr/bin/env python3
"""Build the UNMODIFIED reference CPU path into binaries under oracle/_ref/.

TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is imported by the product
package (dorknet_b200/); only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may load it, and only as the checker /
the baseline that is timed next to the GPU number.

What it does (no reference source is copied into the repo; the sources are
compiled where they lie under /root/reference and only .so files are written,
into the git-ignored oracle/_ref/):

  * the four Cython extension modules the reference's setup.py declares
    (/root/reference/setup.py:6-23): im2col, pooling_cy, relu_cy,
    batch_norm_stats_cy -- same flags (-fopenmp -O3 -ffast-math), built as
    TOP-LEVEL modules exactly as the reference imports them
    (/root/reference/layers/convolution.py:3, batch_norm.py:5, ...).
  * the reference's pure-Python layer / network / optimiser / regulariser
    modules, compiled by Cython to extension modules with their original
    qualified names (layers.convolution, network.feed_forward_network, ...)
    so that the whole reference CPU path can run on the GPU box (which has no
    /root/reference) without a single reference source file travelling.

The three packages the reference imports but this image lacks (cupy, numexpr,
h5py) are satisfied by the tiny stubs in oracle/stubs/ (ours, committed).

Usage:  python oracle/build_ref.py [--ref /root/reference] [--force]
"""
import argparse
import os
import shutil
import subprocess
import sys
import sysconfig
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")

PYX = ["im2col", "pooling_cy", "relu_cy", "batch_norm_stats_cy"]
PYMODS = {
    "layers": ["layer", "convolution", "depthwise_convolution", "pointwise_convolution",
               "batch_norm", "pooling", "activations", "dense_layer", "residual_block",
               "losses"],
    "network": ["feed_forward_network"],
    "optimisers": ["SGD", "SGDMomentum", "RMSProp"],
    "regularisers": ["l2"],
}


def _run(cmd):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + "\n")
        raise SystemExit("oracle/build_ref.py: command failed")
    return r.stdout


def build(ref="/root/reference", force=False, quiet=False):
    if not os.path.isdir(ref):
        raise SystemExit("reference tree %s not present (expected on the GPU box: use the prebuilt oracle/_ref)" % ref)
    import numpy
    ext = sysconfig.get_config_var("EXT_SUFFIX")
    stamp = os.path.join(OUT, ".built")
    if os.path.exists(stamp) and not force:
        return OUT
    os.makedirs(OUT, exist_ok=True)
    gcc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"  # $CC wrapper cannot link -fopenmp here
    inc = ["-I" + sysconfig.get_paths()["include"], "-I" + numpy.get_include()]
    common = ["-shared", "-fPIC", "-w", "-DNPY_NO_DEPRECATED_API=NPY_1_7_API_VERSION"]
    with tempfile.TemporaryDirectory(prefix="dk_ref_build_") as tmp:
        for name in PYX:
            src = os.path.join(ref, "layers", name + ".pyx")
            c = os.path.join(tmp, name + ".c")
            _run([sys.executable, "-m", "cython", "-3", "-o", c, src])
            _run([gcc] + common + ["-fopenmp", "-O3", "-ffast-math"] + inc + [c, "-o", os.path.join(OUT, name + ext)])
            if not quiet:
                print("built", name + ext)
        for pkg, mods in PYMODS.items():
            os.makedirs(os.path.join(OUT, pkg), exist_ok=True)
            for m in mods:
                src = os.path.join(ref, pkg, m + ".py")
                c = os.path.join(tmp, "%s_%s.c" % (pkg, m))
                _run([sys.executable, "-m", "cython", "-3", "--module-name", "%s.%s" % (pkg, m), "-o", c, src])
                _run([gcc] + common + ["-O1"] + inc + [c, "-o", os.path.join(OUT, pkg, m + ext)])
                if not quiet:
                    print("built", "%s/%s%s" % (pkg, m, ext))
    with open(stamp, "w") as f:
        f.write("built from %s\n" % ref)
    return OUT


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--force", action="store_true")
    a = ap.parse_args()
    if a.force and os.path.isdir(OUT):
        shutil.rmtree(OUT)
    print(build(a.ref, a.force))
