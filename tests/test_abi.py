"""-m "not gpu": the C-ABI library builds for sm_100a, loads, and exports every symbol that
include/delora_b200.h declares (no compute calls here)."""
import ctypes
import os
import re

from helpers import ROOT


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "delora_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(delora_[a-z0-9_A-Z]+)\s*\(", text)))


def test_library_builds_and_exports_header_symbols():
    from delora_b200 import build, _lib
    path = build.build()
    assert os.path.exists(path)
    handle = ctypes.CDLL(path)
    declared = _declared_symbols()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(handle, name), f"{name} declared in the header but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature in _lib.py"
    assert set(_lib.SIGNATURES) == set(declared)
    assert handle.delora_abi_version() == _lib.ABI_VERSION


def test_argument_errors_are_reported_without_a_gpu():
    from delora_b200 import _lib
    L = _lib.lib()
    rc = L.delora_normals_fwd(None, 1, 3, 4, 4, 7, 11, 0.5, 10, None, None, None, None)
    assert rc != 0
    assert b"null pointer" in L.delora_last_error()
    assert L.delora_scan_blocks(64 * 2048) == 128
    assert L.delora_icp_partial_rows(1000) == 32


def test_sm100a_sass_present():
    import subprocess
    from delora_b200 import build
    out = subprocess.run(["cuobjdump", "-lelf", build.build()], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_tensor_core_and_tma_instructions_in_sass():
    """The convolution kernels really are tcgen05 + TMA code (SASS mnemonics, /opt/skills/guides/B200_PROFILING.md):
    UTCHMMA = tcgen05.mma kind::f16, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, UTMALDG = cp.async.bulk.tensor;
    the normals kernel uses the packed fp32x2 pipe (FFMA2 / FADD2)."""
    import re
    import subprocess
    from delora_b200 import build
    sass = subprocess.run(["cuobjdump", "-sass", build.build()], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "LDTM", "UTCBAR", "UTMALDG", "FFMA2", "FADD2"):
        assert re.search(r"\b" + mnemonic + r"\b", sass), mnemonic
    assert "HMMA.16816" not in sass and "WGMMA" not in sass        # no mma.sync / wgmma fallbacks
