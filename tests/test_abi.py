"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol include/pafuse_b200.h declares.
No compute entry point is called without a GPU."""
import ctypes
import os
import re
import subprocess

import pytest

from pafuse_b200 import _native, build
from pafuse_testlib import ROOT

HEADER = os.path.join(ROOT, "include", "pafuse_b200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pafuse_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    build.build()
    return ctypes.CDLL(build.LIB_PATH)


def test_header_and_binding_tables_agree():
    assert declared_functions() == sorted(_native.EXPORTED_SYMBOLS)


def test_library_exports_every_declared_symbol(lib):
    for name in declared_functions():
        assert hasattr(lib, name), f"{name} is declared in the header but not exported"


def test_version_and_error_strings(lib):
    lib.pafuse_version.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.pafuse_version()
    lib.pafuse_last_error.restype = ctypes.c_char_p
    assert isinstance(lib.pafuse_last_error(), bytes)


def test_create_rejects_bad_arguments_without_touching_the_gpu(lib):
    lib.pafuse_create.restype = ctypes.c_int32
    assert lib.pafuse_create(None, None) == -1                       # PAFUSE_E_ARG
    lib.pafuse_last_error.restype = ctypes.c_char_p
    assert b"null" in lib.pafuse_last_error()


def test_sass_uses_tcgen05_and_tma():
    """Evidence that the GEMM is Blackwell-native: UTC*MMA (tcgen05.mma), LDTM (tcgen05.ld), UTMALDG (TMA)."""
    cuobjdump = "/usr/local/cuda/bin/cuobjdump"
    if not os.path.isfile(cuobjdump):
        pytest.skip("cuobjdump not available")
    build.build()
    sass = subprocess.run([cuobjdump, "-sass", build.LIB_PATH], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass and "LDTM" in sass and "UTMALDG" in sass
    assert "HMMA." not in sass.replace("UTCHMMA", "")                # no legacy mma.sync path


def test_product_fails_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    import pafuse_b200
    from pafuse_b200 import synthetic
    from pafuse_b200.h3wb import H3WBSkeleton
    sk = H3WBSkeleton()
    m = pafuse_b200.D3DP(synthetic.default_args(depth=1), sk.joints_left, sk.joints_right, sk, is_train=False)
    x2d, x2df = synthetic.synthetic_inputs(1)
    with pytest.raises(_native.PafuseError):
        m.eval()(x2d, None, input_2d_flip=x2df)
    with pytest.raises(_native.PafuseError):
        pafuse_b200.wb_pose_from_parts(torch.zeros(1, 134, 3), H3WBSkeleton())
    with pytest.raises(_native.PafuseError):
        pafuse_b200.project_to_2d(torch.zeros(1, 134, 3), torch.zeros(1, 9))
