"""CPU-only: the C-ABI library builds, loads, and exports every symbol include/jolt_atlas_b200.h declares
(no compute calls — there is no GPU here); without a device ja_init must fail loudly (no CPU fallback)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "jolt_atlas_b200.h")


def header_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ja_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib_path():
    from jolt_atlas_b200 import build
    return build.build()


def test_header_declares_something():
    syms = header_symbols()
    assert "ja_init" in syms and "ja_bind" in syms and "ja_round_eval" in syms
    assert len(syms) >= 20


def test_so_exports_every_header_symbol(lib_path):
    out = subprocess.run(["nm", "-D", "--defined-only", lib_path], capture_output=True, text=True, check=True).stdout
    exported = set(line.split()[-1] for line in out.splitlines() if line.strip())
    missing = [s for s in header_symbols() if s not in exported]
    assert not missing, f"symbols declared in the header but not exported: {missing}"


def test_ctypes_table_matches_header(lib_path):
    from jolt_atlas_b200 import _lib
    assert sorted(_lib.SIGNATURES) == header_symbols()
    _lib.load()   # resolves every symbol


def test_no_cpu_fallback_without_gpu(lib_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from jolt_atlas_b200 import Context, JoltAtlasError
    with pytest.raises(JoltAtlasError) as e:
        Context(0)
    assert e.value.code == -4   # JA_ERR_NO_DEVICE


def test_sass_is_sm100a_and_uses_wide_imad(lib_path):
    out = subprocess.run(["cuobjdump", "-lelf", lib_path], capture_output=True, text=True).stdout
    assert "sm_100a" in out
