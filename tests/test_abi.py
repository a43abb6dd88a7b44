"""The C-ABI library loads and exports every symbol include/c2a.h declares; without a GPU the device entry
points fail loudly (no CPU fallback).  No compute calls here."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "c2a.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(c2a_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(c2a):
    so = ctypes.CDLL(os.path.join(ROOT, "circom-2-arithc_b200", "libc2a.so"))
    syms = declared_symbols()
    assert len(syms) >= 40
    missing = [s for s in syms if not hasattr(so, s)]
    assert not missing, f"declared in include/c2a.h but not exported: {missing}"


def test_python_binding_covers_header(c2a):
    from circom_2_arithc_b200 import _lib
    assert set(declared_symbols()) <= set(_lib.EXPORTED)


def test_abi_version_and_names(c2a):
    assert c2a.lib.c2a_abi_version() == 1
    names = [c2a.lib.c2a_gate_type_name(i).decode() for i in range(20)]
    assert names == [t.name for t in c2a.AGateType]  # src/a_gate_type.rs:7-28 order
    assert c2a.lib.c2a_gate_type_name(20) is None
    assert c2a.lib.c2a_gate_type_from_name(b"AShiftR") == 15 and c2a.lib.c2a_gate_type_from_name(b"nope") == -1
    assert c2a.lib.c2a_status_string(3) == b"Signal already declared"


def test_no_device_fails_loudly(c2a):
    if c2a.have_device():
        pytest.skip("a CUDA device is present")
    with pytest.raises(c2a.C2AError):
        c2a.DeviceContext(0)
    c = c2a.Compiler()
    c.add_signal(0, "0.a", None)
    with pytest.raises(c2a.C2AError):
        c.build_circuit()


def test_product_never_loads_the_oracle():
    """The package must not import / dlopen anything under oracle/."""
    pkg = os.path.join(ROOT, "circom-2-arithc_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                for needle in ("libc2a_oracle", "oracle/", "oracle_lib", "orc_"):
                    assert needle not in text, f"{f} references the oracle ({needle})"
