"""The C-ABI library loads, exports every symbol include/treedet.h declares, and the
ctypes table agrees with the header.  No compute calls (no GPU needed)."""
import os
import re
import subprocess

import pytest

from treedetection_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def header_decls():
    text = open(os.path.join(ROOT, "include", "treedet.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    decls = {}
    for m in re.finditer(r"\b(int|long long|const char\*)\s+(td_\w+)\s*\(([^;]*?)\)\s*;", text, flags=re.S):
        args = m.group(3).strip()
        n = 0 if args in ("", "void") else len([a for a in args.split(",") if a.strip()])
        decls[m.group(2)] = n
    return decls


def test_library_builds_and_exports_header_symbols(header_decls):
    path = build.build()
    assert os.path.exists(path)
    out = subprocess.run(["nm", "-D", "--defined-only", path], capture_output=True, text=True, check=True).stdout
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    missing = sorted(set(header_decls) - exported)
    assert not missing, f"declared in treedet.h but not exported: {missing}"
    undeclared = sorted(n for n in exported if n.startswith("td_") and n not in header_decls)
    assert not undeclared, f"exported but not declared in treedet.h: {undeclared}"


def test_ctypes_table_matches_header(header_decls):
    assert set(_lib.SIGNATURES) == set(header_decls)
    for name, (_, args) in _lib.SIGNATURES.items():
        assert len(args) == header_decls[name], name
    assert sorted(_lib.exported_symbols()) == sorted(_lib.SIGNATURES)


def test_version_and_error_string_without_gpu():
    h = _lib.lib()
    assert h.td_version() == 100
    assert isinstance(h.td_last_error(), bytes)


def test_sm100a_code_is_present():
    r = subprocess.run(["cuobjdump", "--list-elf", _lib.LIB_PATH], capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in r.stdout


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.TreedetError):
        _lib.lib()


def test_cpu_tensors_are_rejected():
    import torch
    from treedetection_b200 import ops
    with pytest.raises(_lib.TreedetError):
        ops.containment(torch.zeros((3, 4), dtype=torch.float32), 0.5)
