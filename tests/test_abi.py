"""The C-ABI library builds for sm_100a, loads, and exports every symbol include/gens_b200.h
declares.  No compute calls (CPU only)."""
import ctypes
import os
import re

from gens_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "gens_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gens_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported():
    so = build.build()
    handle = ctypes.CDLL(so)
    names = _declared()
    assert len(names) >= 5
    for n in names:
        assert hasattr(handle, n), f"{n} declared in gens_b200.h but not exported"
    assert sorted(_lib.exported_symbols()) == names, "python binding table out of sync with the header"


def test_abi_version_and_error_strings():
    h = _lib.lib()
    assert h.gens_abi_version() == _lib.ABI_VERSION
    assert b"bad argument" in h.gens_error_string(-1)
    assert h.gens_error_string(0) == b"ok"


def test_bad_arguments_are_rejected_without_a_gpu():
    h = _lib.lib()
    assert h.gens_pack_feature_maps(None, None, 1, 1, 1, None) == -1
    assert h.gens_volume_agg_fwd(None, 3, 4, 4, None, None, 1.0, None, 8, 0, 8, 0, 512, 1, 0, None, None, None) == -1


def test_no_oracle_import_in_product():
    """The product package must never reach into oracle/ (or any CPU fallback)."""
    pkg = os.path.join(ROOT, "gens_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f
                assert "libgens_oracle" not in src, f
