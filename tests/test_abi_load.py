"""CPU-side checks of the drop-in boundary: the shared library loads, exports every symbol that
include/pwc_b200.h declares, validates arguments before touching CUDA, and fails loudly without a GPU."""
import ctypes
import os
import re

import pytest

from tfoptflow_b200 import _ffi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(_ffi.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return _ffi.load()


def header_symbols():
    src = open(os.path.join(ROOT, "include", "pwc_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pwc_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree(lib):
    syms = header_symbols()
    assert syms, "no prototypes found in include/pwc_b200.h"
    assert sorted(_ffi.SIGNATURES) == syms
    for s in syms:
        assert hasattr(lib, s), f"libpwc_b200.so does not export {s}"


def test_abi_version(lib):
    assert lib.pwc_abi_version() == 1


def test_argument_validation_without_gpu(lib):
    # these are rejected before any CUDA call, so they behave the same on a CPU-only box
    assert lib.pwc_cost_volume_fwd(None, None, None, 1, 4, 4, 4, 4, 7, 0, None) == -3  # unknown dtype
    assert b"dtype" in lib.pwc_last_error()
    assert lib.pwc_cost_volume_fwd(None, None, None, 1, 4, 4, 4, -1, 0, 0, None) == -2  # r < 0
    assert lib.pwc_warp_cost_volume_fwd(None, None, 8, 1.0, None, 1, 1, 4, 4, 4, 0, 0, 0, None) == -2  # H < 2
    assert b"height" in lib.pwc_last_error()
    assert lib.pwc_cost_volume_fwd(None, None, None, 1, 4, 4, 4, 4, 0, 0, None) == -1  # null pointers
    assert lib.pwc_set_option(b"no_such_option", 1) == -1
    assert lib.pwc_set_option(b"fwd_impl", 9) == -1
    # shapes only the unfused kernels take (r != 4): the re-materialised warp + the fp32 dwarp accumulator
    assert lib.pwc_bwd_workspace_bytes(2, 8, 8, 16, 3, 0, 1) == 2 * (2 * 8 * 8 * 16 * 4)
    assert lib.pwc_bwd_workspace_bytes(2, 8, 8, 16, 3, 1, 1) == 2 * 8 * 8 * 16 * (2 + 4 + 4)
    assert lib.pwc_bwd_workspace_bytes(2, 8, 8, 16, 3, 0, 0) == 0
    # shapes the fused kernels cover (r == 4, C % 4 == 0): the g' tensor [B * H][81][W] fp32, plus an fp32 dc2
    # accumulator for 16-bit tensors
    gp = 2 * 8 * 81 * 8 * 4
    assert lib.pwc_bwd_workspace_bytes(2, 8, 8, 16, 4, 0, 1) == gp
    assert lib.pwc_bwd_workspace_bytes(2, 8, 8, 16, 4, 0, 0) == gp
    assert lib.pwc_bwd_workspace_bytes(2, 8, 8, 16, 4, 1, 1) == gp + 2 * 8 * 8 * 16 * 4


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    buf = (ctypes.c_float * 4096)()
    p = ctypes.addressof(buf)
    rc = lib.pwc_cost_volume_fwd(p, p, p, 1, 4, 4, 4, 4, 0, 0, None)
    assert rc in (-5, -6), "without a B200 the call must fail loudly, not compute on the CPU"
    import tfoptflow_b200
    with pytest.raises(RuntimeError):
        tfoptflow_b200.cost_volume(torch.zeros(1, 4, 4, 4), torch.zeros(1, 4, 4, 4), 4, "corr")


def test_missing_library_fails_loudly(tmp_path):
    with pytest.raises(RuntimeError):
        _ffi.load(str(tmp_path / "libpwc_b200.so"))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "tfoptflow_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f


def test_every_option_key_is_documented_and_readable(lib):
    """every key pwc_set_option accepts is documented in include/pwc_b200.h, listed in bench.py's OPTION_DEFAULTS (so a
    non-default value is echoed in the bench line) and readable through pwc_get_option (both work without a GPU)"""
    import re
    src = open(os.path.join(ROOT, "tfoptflow_b200", "csrc", "pwc_abi.cu")).read()
    body = src[src.index("int pwc_set_option("):src.index("int pwc_get_option(")]
    keys = sorted(set(re.findall(r'strcmp\(key, "([a-z0-9_]+)"\)', body)))
    assert len(keys) >= 12
    header = open(os.path.join(ROOT, "include", "pwc_b200.h")).read()
    bench = open(os.path.join(ROOT, "bench.py")).read()
    for k in keys:
        assert f'"{k}"' in header, f"option {k} is not documented in include/pwc_b200.h"
        assert f'"{k}"' in bench, f"option {k} is missing from bench.py OPTION_DEFAULTS"
        v = ctypes.c_int(-12345)
        assert lib.pwc_get_option(k.encode(), ctypes.byref(v)) == 0, k
        assert v.value != -12345
