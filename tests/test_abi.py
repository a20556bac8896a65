"""The C-ABI library loads, exports every symbol include/phb200.h declares,
and the product path fails loudly (no CPU fallback) when no GPU is present."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "phb200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(phb200_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from phasta_b200 import lib
    L = lib.load()
    syms = header_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(L, s), "libphb200.so does not export %s" % s
    assert sorted(lib.SYMBOLS) == syms, "lib.SYMBOLS out of sync with include/phb200.h"


def test_struct_sizes_match_header():
    from phasta_b200 import lib
    L = lib.load()
    assert L.phb200_sizeof_common() == C.sizeof(lib.PhbCommon)
    assert L.phb200_sizeof_step() == C.sizeof(lib.PhbStep)


def test_sm100a_code_is_in_the_library():
    import subprocess
    from phasta_b200 import lib
    try:
        out = subprocess.run(["cuobjdump", "-lelf", lib.LIB_PATH], capture_output=True, text=True).stdout
    except FileNotFoundError:
        pytest.skip("cuobjdump not on PATH")
    assert "sm_100a" in out


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from common import make_case
    from phasta_b200.solver import PhastaGPU, PhastaError
    params, tables, parts, states = make_case(2, 2, 2, bc="none", periodic_z=False)
    with pytest.raises(PhastaError):
        PhastaGPU(parts[0], params, tables)


def test_product_never_imports_oracle():
    """The product path must not import, link or execute anything under oracle/."""
    pkg = os.path.join(ROOT, "phasta_b200")
    bad = re.compile(r"(from\s+oracle|import\s+oracle|oracle_py|libphasta_oracle|#include\s*[<\"].*oracle)")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".h", ".cpp", "Makefile")):
                txt = open(os.path.join(dp, f)).read()
                assert not bad.search(txt), "%s references the oracle" % f
