"""CPU: the C-ABI library loads without a GPU and exports every symbol that
include/gpc_b200.h declares; the Python binding table matches the header."""
import os
import re

import pytest

from graph_peak_caller_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    text = open(os.path.join(ROOT, "include", "gpc_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gpc_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = header_functions()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), "libgpc_b200.so lacks %s" % name


def test_binding_table_matches_header():
    assert sorted(_lib.exported_symbols()) == header_functions()


def test_version_and_error_strings():
    lib = _lib.load()
    assert b"sm_100a" in lib.gpc_version_string()
    assert isinstance(lib.gpc_last_error_string(), bytes)


def test_no_cpu_fallback_without_a_device():
    """Without a CUDA device the product path raises instead of computing."""
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from graph_peak_caller_b200.custom_exceptions import DeviceLibraryMissing
    from graph_peak_caller_b200.intervals import Intervals
    from graph_peak_caller_b200.sample import get_fragment_pileup
    from graph_peak_caller_b200.synthetic import make_graph, make_reads
    from graph_peak_caller_b200 import Configuration
    g = make_graph(1000, 0.02, seed=1)
    r = make_reads(g, 10, 10, 30, seed=2)
    c = Configuration()
    c.read_length, c.fragment_length = 10, 30
    with pytest.raises(DeviceLibraryMissing):
        get_fragment_pileup(g, Intervals(r), c)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "graph_peak_caller_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f


def test_ctypes_structs_match_the_header(tmp_path):
    """sizeof / offsetof of gpc_graph_t and gpc_reads_t as the C compiler sees
    them against the ctypes mirrors the host side passes."""
    import ctypes
    import shutil
    import subprocess
    from graph_peak_caller_b200 import _lib
    if shutil.which("gcc") is None:
        pytest.skip("no C compiler")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "layout.c"
    fields = {"gpc_graph_t": [f[0] for f in _lib.GpcGraph._fields_],
              "gpc_reads_t": [f[0] for f in _lib.GpcReads._fields_]}
    body = ['#include <stdio.h>', '#include <stddef.h>', '#include "gpc_b200.h"', 'int main(void) {']
    for name, fs in fields.items():
        body.append('printf("%s %%zu\\n", sizeof(%s));' % (name, name))
        for f in fs:
            body.append('printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (name, f, name, f))
    body.append("return 0; }")
    src.write_text("\n".join(body))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(root, "include"), "-o", str(exe), str(src)], check=True)
    out = dict(line.split() for line in subprocess.run([str(exe)], capture_output=True, text=True,
                                                       check=True).stdout.splitlines())
    for name, cls in (("gpc_graph_t", _lib.GpcGraph), ("gpc_reads_t", _lib.GpcReads)):
        assert int(out[name]) == ctypes.sizeof(cls), name
        for f in fields[name]:
            assert int(out["%s.%s" % (name, f)]) == getattr(cls, f).offset, (name, f)
