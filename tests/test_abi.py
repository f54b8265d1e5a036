"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads, exports every symbol the headers under
include/ declare, and refuses to work without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import pytest

import libecc_b200
from common import ROOT


def declared_symbols(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return {m.group(1) for m in re.finditer(r"\b(eccb200_\w+|prj_pt_mul\w*|ec_verify\w*|is_verify_batch_mode_supported)\s*\(", txt)}


def test_library_exports_every_declared_symbol():
    if not os.path.exists(libecc_b200.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(libecc_b200.LIB_PATH)
    decl = declared_symbols("libecc_b200.h")
    assert set(libecc_b200.ABI_SYMBOLS) == decl
    for name in sorted(decl):
        assert hasattr(lib, name), f"{name} declared in include/libecc_b200.h but not exported"
    drop = ctypes.CDLL(os.path.join(os.path.dirname(libecc_b200.LIB_PATH), "libecc_b200_dropin.so"))
    decl2 = declared_symbols("libecc_b200_dropin.h")
    assert {"prj_pt_mul", "prj_pt_mul_blind", "ec_verify", "ec_verify_batch", "is_verify_batch_mode_supported",
            "eccb200_dropin_ecdsa_verify_batch",
            "eccb200_dropin_ecfsdsa_verify_batch"} <= decl2
    for name in sorted(decl2):
        assert hasattr(drop, name), f"{name} declared in include/libecc_b200_dropin.h but not exported"
    assert set(os.listdir(os.path.join(ROOT, "include"))) == {"libecc_b200.h", "libecc_b200_dropin.h"}


def test_curve_metadata_without_gpu():
    assert libecc_b200.curve_sizes("SECP256R1") == (32, 32)
    assert libecc_b200.curve_sizes("FRP256V1") == (32, 32)
    assert libecc_b200.curve_sizes("SECP384R1") == (48, 48)
    lib = libecc_b200.load_library()
    assert lib.eccb200_curve_name(4) == b"SECP256R1" and lib.eccb200_curve_name(99) is None


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(libecc_b200.EccB200Error, match="no CUDA device"):
        libecc_b200.Engine("SECP256R1")


def test_product_does_not_reference_the_oracle():
    """The product sources must not include, link or load anything under oracle/ or tests/hostsim."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "libecc_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".c", ".h", ".inc")) or f == "Makefile":
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                code = re.sub(r"/\*.*?\*/|#.*?$|//.*?$|\"\"\".*?\"\"\"", "", txt, flags=re.S | re.M)
                assert "ecc_oracle" not in code and "libecc_ref" not in code and "hostsim" not in code, f
