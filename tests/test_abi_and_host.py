"""CPU tests of the product's host side: the C-ABI library loads and exports every symbol include/compute_war.h
declares, the pure size functions agree with the reference's macros, and without a GPU the engine fails loudly
instead of falling back to a CPU path.  (No compute calls here; parity proper is -m gpu.)"""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, "include", "compute_war.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(cw_[a-z0-9_]+)\s*\(", txt)) - {"cw_complete_fn"})


def test_library_exports_every_declared_symbol():
    import compute_war_b200 as cw
    path = cw.library_path()
    if not os.path.exists(path):
        pytest.fail("libcomputewar.so not built: run __graft_entry__.build()")
    L = C.CDLL(path)
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/compute_war.h but not exported"


def test_size_functions_match_reference_macros():
    import compute_war_b200 as cw
    L = cw.load_library()
    for chunk in (13, 4096, 8192, 65536):
        assert L.cw_compress_bound(cw.COMP_LZ4, chunk) == chunk + chunk // 255 + 16  # LZ4_COMPRESSBOUND, lz4.h:156
        assert L.cw_compress_bound(cw.COMP_LZF, chunk) == chunk
        assert cw.compress_bound(cw.COMP_LZ4, chunk) == L.cw_compress_bound(cw.COMP_LZ4, chunk)
    assert [L.cw_digest_bytes(a) for a in (0, 1, 2)] == [64, 16, 32]
    assert L.cw_compress_bound(cw.COMP_LZ4, 4096) <= 2 * 4096  # the reference's 2*l scratch (HashAndCompress.cpp:353)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import compute_war_b200 as cw
    assert cw.load_library().cw_device_count() == 0
    with pytest.raises(cw.EngineError) as ei:
        cw.Engine(0)
    assert ei.value.status == -2


def test_product_does_not_reference_the_oracle():
    """the product path must not import, link or execute anything under oracle/"""
    pkg = os.path.join(ROOT, "compute_war_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", "Makefile")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "liboracle" not in txt and "cw_oracle" not in txt and "oracle_lib" not in txt, (dirpath, f)
    out = subprocess.run(["ldd", os.path.join(pkg, "libcomputewar.so")], capture_output=True, text=True).stdout
    assert "oracle" not in out and "skref" not in out


def test_host_tool_cli_rejects_bad_selector():
    exe = os.path.join(ROOT, "compute_war_b200", "hashandcompress")
    if not os.path.exists(exe):
        pytest.skip("host tool not built")
    r = subprocess.run([exe, "--comp-alg", "zip", "/dev/null"], capture_output=True, text=True)
    assert r.returncode == 2 and "invalid compression algorithm" in r.stderr
