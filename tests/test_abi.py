"""CPU suite: the C-ABI library loads here (no GPU) and exports every symbol include/chalte.h declares."""
import ctypes as C
import os
import re

import pytest

from tests.helpers import ROOT


def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "chalte.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(cha_[a-z_0-9]+)\s*\(", txt)))


def test_library_builds_and_exports_every_declared_symbol():
    from cha1_mcmc_b200.build import build_library
    from cha1_mcmc_b200.engine import SIGNATURES
    lib = C.CDLL(build_library())
    syms = _header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/chalte.h but not exported"
        assert s in SIGNATURES, f"{s} has no ctypes signature in engine.py"
    assert set(SIGNATURES) == set(syms)


def test_library_is_sm100a_only_and_has_tma_and_mufu():
    """cuobjdump evidence that the product .so carries sm_100a SASS with TMA bulk copies and MUFU.EX2."""
    import shutil
    import subprocess
    from cha1_mcmc_b200.build import LIB, build_library
    build_library()
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe):
        pytest.skip("cuobjdump not available")
    elf = subprocess.run([exe, "-lelf", LIB], capture_output=True, text=True).stdout
    assert "sm_100a" in elf and "sm_90" not in elf
    sass = subprocess.run([exe, "-sass", LIB], capture_output=True, text=True).stdout
    assert "UBLKCP" in sass, "TMA bulk copy (cp.async.bulk) missing from SASS"
    assert "MUFU.EX2" in sass
    assert "SYNCS" in sass      # mbarrier


def test_no_cpu_fallback_without_gpu():
    """Without a device the engine must refuse loudly (on the GPU box this test is a no-op)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from cha1_mcmc_b200 import EngineError, LTEEngine
    with pytest.raises(EngineError, match="no CPU fallback"):
        LTEEngine(device=0)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "cha1_mcmc_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("no oracle", ""), f"{f} references the oracle"
