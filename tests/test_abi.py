"""CPU-side checks of the boundary: the C-ABI library loads, exports every symbol
include/b200moc.h declares, and refuses to run without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT, load_case
from openmoc_b200 import capi


def header_symbols():
    text = open(os.path.join(ROOT, "include", "b200moc.h")).read()
    return sorted(set(re.findall(r"\b(b200_[A-Za-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    assert header_symbols() == capi.EXPORTS


def test_library_exports_every_declared_symbol():
    lib = capi.load()
    for name in header_symbols():
        assert hasattr(lib, name), name
    assert lib.b200_version() >= 100


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from openmoc_b200.solver import B200Solver
    ft, _ = load_case("pin_cell")
    with pytest.raises(capi.B200Error, match="no CPU fallback|CUDA"):
        B200Solver(ft)


def test_null_handle_is_an_error_not_a_crash():
    lib = capi.load()
    assert lib.b200_finalize(None) != 0
    assert b"null solver handle" in lib.b200_last_error()


def test_product_never_imports_oracle():
    """The product path must not route through the oracle (or the reference)."""
    pkg = os.path.join(ROOT, "openmoc_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert "moc_oracle" not in src and "oracle_py" not in src and "oracle/" not in src, f


def test_ctypes_structs_match_the_header(tmp_path):
    """capi.Config / CmfdConfig / CmfdStats mirror b200_config / b200_cmfd_config / b200_cmfd_stats field for field:
    sizes and offsets from a C compiler against ctypes."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no C compiler")
    fields = {"b200_config": (capi.Config, None), "b200_cmfd_config": (capi.CmfdConfig, None), "b200_cmfd_stats": (capi.CmfdStats, None)}
    src = ["#include <stdio.h>", "#include <stddef.h>", '#include "b200moc.h"', "int main(void) {"]
    for cname, (cls, _) in fields.items():
        src.append('printf("%s size %%zu\\n", sizeof(%s));' % (cname, cname))
        for fname, _ in cls._fields_:
            src.append('printf("%s %s %%zu\\n", offsetof(%s, %s));' % (cname, fname, cname, fname))
    src.append("return 0; }")
    c = os.path.join(tmp_path, "layout.c")
    open(c, "w").write("\n".join(src))
    exe = os.path.join(tmp_path, "layout")
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-o", exe, c], check=True)
    out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout.split("\n")
    for line in filter(None, out):
        cname, what, value = line.split()
        cls = fields[cname][0]
        if what == "size":
            assert C.sizeof(cls) == int(value), cname
        else:
            assert getattr(cls, what).offset == int(value), (cname, what)
