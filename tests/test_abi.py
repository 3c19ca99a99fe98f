"""CPU checks of the drop-in boundary: libtdb200.so loads and exports exactly what include/tdb200.h declares."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "tdb200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tdb200_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def libpath():
    from turbodiffusion_b200 import _build
    return _build.build()


def test_header_declares_the_hot_path():
    names = _declared()
    for required in ("tdb200_quant_int8_block128", "tdb200_gemm_w8a8", "tdb200_rms_norm", "tdb200_layer_norm",
                     "tdb200_layer_norm_modulate_quant", "tdb200_gate_residual", "tdb200_rope_interleaved",
                     "tdb200_sla_block_map", "tdb200_sla_quant_qk", "tdb200_sla_linear_moments", "tdb200_sla_attn_fwd"):
        assert required in names


def test_library_exports_every_declared_symbol(libpath):
    out = subprocess.run(["nm", "-D", "--defined-only", libpath], capture_output=True, text=True, check=True).stdout
    exported = sorted(set(re.findall(r" T (tdb200_[a-z0-9_]+)", out)))
    assert exported == _declared()


def test_ctypes_binding_matches_header(libpath):
    from turbodiffusion_b200 import _lib
    assert sorted(_lib.exported_names()) == _declared()
    lib = _lib.lib()  # raises if any prototype names a missing symbol
    assert lib.tdb200_abi_version() == 1
    assert lib.tdb200_last_error() is not None


def test_no_torch_types_in_the_abi():
    src = open(HEADER).read()
    assert "torch" not in src.replace("PyTorch", "") or "at::" not in src
    assert "at::Tensor" not in src and "torch::" not in src


def test_errors_are_codes_not_exits(libpath):
    """Argument validation happens before any CUDA call, so it can be exercised without a GPU."""
    from turbodiffusion_b200 import _lib
    lib = _lib.lib()
    rc = lib.tdb200_quant_int8_block128(None, 0, 4, 8, None, None, None)
    assert rc == -1 and b"null" in lib.tdb200_last_error()
    rc = lib.tdb200_gemm_w8a8(ctypes.c_void_p(16), ctypes.c_void_p(16), ctypes.c_void_p(16), ctypes.c_void_p(16), None,
                              ctypes.c_void_p(16), 0, 128, 128, 100, None)
    assert rc == -2 and b"multiple of 128" in lib.tdb200_last_error()


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "turbodiffusion_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "td_oracle" not in text and "from oracle" not in text and "import oracle" not in text, f
