"""CPU: libtclip_b200.so loads, exports every symbol include/tclip_b200.h declares, the ctypes table matches the header,
and the argument / device checks fail cleanly without a GPU (no compute calls here)."""
from __future__ import annotations

import ctypes
import os
import re

import pytest

from tclip_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "tclip_b200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tclip_[a-z0-9_]+)\s*\(", src)))


def test_library_is_built():
    assert os.path.isfile(_lib.LIB_PATH), "run `python __graft_entry__.py` (build()) first"


def test_every_declared_symbol_is_exported_and_bound():
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    assert set(names) == set(_lib.SIGNATURES), set(names) ^ set(_lib.SIGNATURES)


def test_header_cites_the_reference():
    src = open(HEADER).read()
    for cite in ("zero_shot/em_dirichlet.py:153-177", "few_shot/em_dirichlet.py:196-200", "src/utils.py:380-399",
                 "zero_shot/hard_em_dirichlet.py:256-258"):
        assert cite in src


@pytest.mark.parametrize("struct,mirror", [("tclip_dirichlet_problem", "DirichletProblem"), ("tclip_kmeans_problem", "KMeansProblem")])
def test_struct_layout_matches_header(struct, mirror):
    src = open(HEADER).read()
    body = src[src.index("typedef struct " + struct):src.index("} " + struct + ";")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = []
    for decl in body.split("{", 1)[1].split(";"):
        decl = decl.strip()
        if not decl:
            continue
        names = decl.split(",")
        first = names[0].split()[-1].lstrip("*")
        fields.append(first)
        fields += [n.strip().lstrip("*") for n in names[1:]]
    assert fields == [f[0] for f in getattr(_lib, mirror)._fields_]


def test_plain_calls_without_gpu():
    lib = _lib.load()
    assert lib.tclip_version() == 102
    assert lib.tclip_mm_max_dim() == 1024
    assert lib.tclip_launch_count() >= 0
    assert lib.tclip_dirichlet_mm_workspace_bytes(75000) > 0
    assert lib.tclip_dirichlet_mm_workspace_bytes(0) == 0


def test_bad_arguments_fail_before_touching_the_device():
    lib = _lib.load()
    assert lib.tclip_log_features(None, None, 10, None) == -1
    assert b"bad arguments" in lib.tclip_last_error()
    assert lib.tclip_dirichlet_mm(None, None, None, 10, 10, 10, 50, 1e-11, None, None, 0, None) == -1
    # D beyond the register-resident limit is rejected, not silently truncated
    one = ctypes.c_void_p(8)
    assert lib.tclip_dirichlet_mm(one, one, one, 10, 2000, 10, 50, 1e-11, None, one, 1 << 20, None) == -1
    assert b"D=2000" in lib.tclip_last_error()
    p = _lib.DirichletProblem(n_task=1, n_query=75, n_class=10, dim=4096, iters=1, iter_mm=10, check_every=50)
    assert lib.tclip_dirichlet_em_workspace_bytes(ctypes.byref(p)) == 0
    p.dim = 10
    assert lib.tclip_dirichlet_em_workspace_bytes(ctypes.byref(p)) > 0
    assert lib.tclip_dirichlet_em_run(ctypes.byref(p), None, 0, None) == -1      # outputs missing


def test_no_device_is_an_error_not_a_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lib = _lib.load()
    assert lib.tclip_device_check(0) == -3
    with pytest.raises(_lib.TclipError):
        _lib.check(lib.tclip_device_check(0))


def test_header_is_plain_c99(tmp_path):
    """include/tclip_b200.h is the drop-in boundary: it must compile as C (no C++ types, no torch types) and link against
    the library from a C translation unit."""
    import subprocess
    src = tmp_path / "abi.c"
    src.write_text('#include "tclip_b200.h"\n#include <stddef.h>\n'
                   'int main(void) { tclip_dirichlet_problem p; (void)p;\n'
                   '  return (tclip_version() >= 100 && tclip_dirichlet_em_workspace_bytes(NULL) == 0) ? 0 : 1; }\n')
    exe = tmp_path / "abi"
    inc = os.path.join(ROOT, "include")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", inc, str(src), "-o", str(exe),
                    _lib.LIB_PATH, "-Wl,-rpath," + os.path.dirname(_lib.LIB_PATH)], check=True)
    assert subprocess.run([str(exe)]).returncode == 0
