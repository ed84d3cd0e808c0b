"""The C-ABI library loads on a CPU-only box and exports every symbol include/dvpari.h declares;
compute entry points refuse to run without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

import dvpari

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "dvpari.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dvp_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = dvpari.lib()
    names = declared_symbols()
    assert len(names) >= 35
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert lib.dvp_abi_version() == 1
    assert lib.dvp_strerror(0) == b"ok" and lib.dvp_strerror(9) == b"no CUDA device"


def test_no_cpu_fallback_without_a_device():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(dvpari.DvpError) as e:
        dvpari.Context(0)
    assert e.value.code == 9  # DVP_ERR_NO_DEVICE


def test_product_does_not_import_the_oracle():
    """Only tests/, smoke() and bench.py's CPU-baseline legs may touch oracle/."""
    pkg = os.path.join(ROOT, "dv-pari_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h", "Makefile")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle/" not in text.replace("the oracle", "") or "NOT the oracle" in text, f
                assert "import oracle" not in text and "from oracle" not in text, f
    out = os.popen(f"ldd {dvpari.LIB_PATH}").read()
    assert "liboracle" not in out
