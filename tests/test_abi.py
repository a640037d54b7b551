"""CPU: the C-ABI library builds/loads without a GPU and exports every symbol the header declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "dorknet_b200.h")


def header_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dk_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def cdll():
    from dorknet_b200 import build
    return ctypes.CDLL(build.build())


def test_header_declares_the_families_the_survey_lists():
    names = set(header_functions())
    for fam in ["dk_conv2d_fwd", "dk_conv2d_dgrad", "dk_conv2d_wgrad", "dk_pwconv_fwd", "dk_pwconv_dgrad",
                "dk_pwconv_wgrad", "dk_dwconv_fwd", "dk_dwconv_bwd", "dk_bn_stats", "dk_bn_fwd_train",
                "dk_bn_fwd_infer", "dk_bn_bwd", "dk_relu_fwd", "dk_relu_bwd", "dk_gap_fwd", "dk_gap_bwd",
                "dk_maxpool_fwd", "dk_maxpool_fwd_train", "dk_maxpool_bwd", "dk_dense_fwd", "dk_dense_bwd",
                "dk_add_relu_fwd", "dk_sumsq", "dk_opt_sgd_multi", "dk_opt_sgdm_multi", "dk_opt_rmsprop_multi",
                "dk_softmax_xent_fwd", "dk_softmax_xent_bwd", "dk_bias_grad", "dk_im2col_materialise",
                "dk_init", "dk_destroy", "dk_last_error", "dk_version"]:
        assert fam in names, fam


def test_library_exports_every_declared_symbol(cdll):
    missing = [n for n in header_functions() if not hasattr(cdll, n)]
    assert not missing, missing


def test_python_prototypes_cover_the_header():
    from dorknet_b200 import _lib
    names = set(header_functions())
    assert names - set(_lib.PROTOS) == set()
    assert set(_lib.PROTOS) - names == set()


def test_no_compute_without_gpu_fails_loudly(cdll):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    cdll.dk_last_error.restype = ctypes.c_char_p
    assert cdll.dk_version() >= 100
    assert cdll.dk_init(0) != 0
    assert b"cuda" in cdll.dk_last_error().lower()
    import numpy as np
    from dorknet_b200.layers.activations import ReLu
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ReLu("r").forward(np.zeros((1, 1, 2, 2), np.float32))


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: no product source may import, load or execute it."""
    pkg = os.path.join(ROOT, "dorknet_b200")
    pat = re.compile(r"^\s*(import|from)\s+oracle\b|oracle[/.](oracle|refload|dk_oracle|_ref|_build)", re.M)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                assert not pat.search(open(os.path.join(dirpath, f)).read()), f


def test_struct_layouts_match_the_header(tmp_path):
    """The ctypes mirrors of the structs that cross the C ABI have the header's size and field offsets (checked with
    gcc on include/dorknet_b200.h: no GPU, no CUDA toolkit needed)."""
    import shutil
    import subprocess
    from dorknet_b200 import _lib
    cc = shutil.which("gcc") or shutil.which("cc")
    if cc is None:
        pytest.skip("no C compiler")
    fields = {"dk_opt_tensor": (_lib.OptTensor, ["param", "grad", "state", "n"]),
              "dk_p2p_ctx": (_lib.P2PCtx, ["world", "rank", "grad_delta", "ready", "done", "epoch", "reduced", "grad_base", "nfloats",
                                          "slice"])}
    src = ["#include <stdio.h>", "#include <stddef.h>", '#include "dorknet_b200.h"', "int main(void) {"]
    for name, (_, fl) in fields.items():
        src.append('printf("%s %%zu", sizeof(%s));' % (name, name))
        for f in fl:
            src.append('printf(" %%zu", offsetof(%s, %s));' % (name, f))
        src.append('printf("\\n");')
    src += ["return 0;", "}"]
    c = tmp_path / "layout.c"
    c.write_text("\n".join(src))
    exe = tmp_path / "layout"
    subprocess.run([cc, "-I", os.path.join(ROOT, "include"), str(c), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split("\n")
    for line in out:
        if not line.strip():
            continue
        name, size, *offs = line.split()
        cls, fl = fields[name]
        assert ctypes.sizeof(cls) == int(size), name
        assert [getattr(cls, f).offset for f in fl] == [int(o) for o in offs], name
