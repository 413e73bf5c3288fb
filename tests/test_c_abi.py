"""CPU tests of the drop-in boundary: libcgoptim.so loads, exports every function that
include/cgoptim.h declares (parsed from the header, not from a hand-kept list), the ctypes table of
the Python host covers the same set, every entry point is `extern "C"` with plain-C types, and
without a GPU the product fails loudly instead of falling back to anything."""
import ctypes
import os
import re
import subprocess

import pytest

import cgoptim_b200 as cg
from cgoptim_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "cgoptim.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cgo_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_and_library_exports_the_same_functions():
    names = _declared()
    assert len(names) >= 50 and "cgo_eval_trial" in names and "cgo_batched_minimize_rosenbrock" in names
    lib = ctypes.CDLL(cg.LIB_PATH)
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in include/cgoptim.h but not exported: {missing}"
    out = subprocess.run(["nm", "-D", "--defined-only", cg.LIB_PATH], capture_output=True, text=True).stdout
    exported = sorted(set(re.findall(r"\sT\s+(cgo_[a-z0-9_]+)$", out, flags=re.M)))
    undeclared = [n for n in exported if n not in names]
    assert not undeclared, f"exported with C linkage but missing from include/cgoptim.h: {undeclared}"


def test_python_binding_table_matches_the_header():
    assert sorted(cg.EXPORTED_SYMBOLS) == _declared()


def test_header_is_plain_c():
    """no torch / C++ types in any signature: the header compiles as C"""
    n = ctypes.sizeof(_capi.BatchedConfig)                    # the ctypes mirror must have the C layout
    src = ("#include \"cgoptim.h\"\n_Static_assert(sizeof(cgo_batched_config) == %d, \"layout\");\n"
           "int main(void) { return 0; }\n" % n)
    p = subprocess.run(["gcc", "-std=c11", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-x", "c", "-",
                        "-o", "/dev/null"], input=src, capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    assert n == 104


def test_no_cpu_fallback_without_a_gpu():
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    with pytest.raises(cg.CgoError) as e:
        cg.Context(0)
    assert "no CPU fallback" in str(e.value)
    # the product package never reaches the oracle
    pkg = os.path.join(ROOT, "conjugategradientoptim.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, os.path.join(dirpath, f)


def test_shard_range_is_host_only_and_tiles_the_index_space():
    for n, world in ((200_000_000, 8), (2002, 3), (10, 4)):
        edges = [cg.shard_range(n, world, r, 2) for r in range(world)]
        assert edges[0][0] == 0 and edges[-1][1] == n
        assert all(edges[r][1] == edges[r + 1][0] for r in range(world - 1))
        assert all(lo % 2 == 0 for lo, _ in edges)
