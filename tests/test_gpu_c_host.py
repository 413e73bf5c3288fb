"""A non-Python host of the C ABI: tests/c_host/host.c (plain C11, dlopen + include/cgoptim.h only) runs
Hager-Zhang CG + StrongWolfeBisection on extended Rosenbrock n = 10,000 with exactly the call sequence the Julia
binding makes (julia/B200CGOptim/src/optim.jl:4-52), and its trace must match the committed
golden trace (tests/golden/traces.json, case rosenbrock_n10000 / cgo / fused) and a live oracle run bit for bit.
Julia itself is not installed in this image: this is the executable stand-in for the `ccall` host."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
SRC = os.path.join(ROOT, "tests", "c_host", "host.c")


def _build(tmp_path):
    exe = str(tmp_path / "c_host")
    p = subprocess.run(["gcc", "-std=c11", "-O2", "-Wall", "-Wextra", "-Werror", "-o", exe, SRC, "-ldl", "-lm"],
                       capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    return exe


def test_c_host_compiles_as_plain_c(tmp_path):
    """(CPU) the host builds against include/cgoptim.h with a C compiler and fails loudly without a GPU / library"""
    exe = _build(tmp_path)
    p = subprocess.run([exe, "/nonexistent/libcgoptim.so", "100", "1"], capture_output=True, text=True)
    assert p.returncode == 2 and "dlopen" in p.stderr


@pytest.mark.gpu
def test_c_host_matches_golden_and_oracle(tmp_path):
    import cgoptim_b200 as cg
    from oracle import oracle as O
    exe = _build(tmp_path)
    n = 10_000
    case = [c for c in json.load(open(os.path.join(ROOT, "tests", "golden", "traces.json")))["cases"]
            if c["name"] == "rosenbrock_n10000" and c["sum_mode"] == "cgo" and c["beta_form"] == "fused"][0]
    k = len(case["trace_objective"])
    exp = tmp_path / "expected.txt"
    with open(exp, "w") as f:
        for i in range(k):       # same format as host.c prints: k f ‖g‖ a* evals, C99 %a hex floats
            vals = [float.fromhex(case[key][i]) for key in ("trace_objective", "trace_grad_norm", "trace_step_size")]
            f.write("%d %s %s %s %d\n" % (i + 1, *[_c99_hex(v) for v in vals], case["trace_objective_evals"][i]))
    p = subprocess.run([exe, cg.LIB_PATH, str(n), str(k), str(exp)], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout[-1500:] + p.stderr[-1500:]
    assert f"compared {k} trace lines, 0 mismatches" in p.stderr
    # the whole run against a live oracle run (all iterations, final status and minimiser)
    p = subprocess.run([exe, cg.LIB_PATH, str(n), "1000"], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr[-1500:]
    lines = p.stdout.strip().splitlines()
    ora = O.minimize(O.Objective.rosenbrock(n), O.rosenbrock_x0(n, 24, 0.0),
                     O.make_config("HagerZhang", "StrongWolfeBisection", eps=1e-5, max_iters=1000, sum_mode="cgo", beta_form="fused"))
    last = lines[-1].split()
    assert last[0] == "status" and last[1] == ora.status and int(last[3]) == ora.iters_ran == case["iters_ran"]
    assert float.fromhex(last[5]) == ora.objective and float.fromhex(last[7]) == ora.minimizer[0]
    tr = np.array([[float.fromhex(t) for t in ln.split()[1:4]] for ln in lines[:-1]])
    assert len(tr) == ora.iters_ran
    assert np.array_equal(tr[:, 0], ora.trace_objective) and np.array_equal(tr[:, 1], ora.trace_grad_norm)
    assert np.array_equal(tr[:, 2], ora.trace_step_size)
    assert [int(ln.split()[4]) for ln in lines[:-1]] == [int(v) for v in ora.trace_objective_evals]


@pytest.mark.gpu
@pytest.mark.parametrize("coh", [0, 30])
def test_c_host_runs_the_csr_least_squares_path(tmp_path, coh):
    """the same C host on BASELINE.json's headline objective (banded-random CSR least squares; coh 0: sliced layout +
    k_spmv_direct, coh 30: the TMA-streamed k_csr_rows): whole run bit for bit against the oracle."""
    import cgoptim_b200 as cg
    from oracle import oracle as O
    exe = _build(tmp_path)
    n, W = 50_000, 4096
    p = subprocess.run([exe, cg.LIB_PATH, str(n), "1000", "-", "0", "sparse_ls", str(W), str(coh)],
                       capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout[-1500:] + p.stderr[-1500:]
    lines = p.stdout.strip().splitlines()
    ora = O.minimize(O.Objective.sparse_ls(n, 10, W, 24, coh), np.zeros(n),
                     O.make_config("HagerZhang", "StrongWolfeBisection", eps=1e-5, max_iters=1000, sum_mode="cgo", beta_form="fused"))
    last = lines[-1].split()
    assert last[1] == ora.status == "success" and int(last[3]) == ora.iters_ran
    assert float.fromhex(last[5]) == ora.objective and float.fromhex(last[7]) == ora.minimizer[0]
    tr = np.array([[float.fromhex(t) for t in ln.split()[1:4]] for ln in lines[:-1]])
    assert np.array_equal(tr[:, 0], ora.trace_objective) and np.array_equal(tr[:, 1], ora.trace_grad_norm)
    assert np.array_equal(tr[:, 2], ora.trace_step_size)
    assert [int(ln.split()[4]) for ln in lines[:-1]] == [int(v) for v in ora.trace_objective_evals]


def _c99_hex(v):
    """printf("%a") as glibc prints it (Python's float.hex pads the mantissa: 0x1.8000000000000p+1 vs 0x1.8p+1)"""
    s = float(v).hex()
    if "p" not in s:
        return s
    mant, ex = s.split("p")
    if "." in mant:
        mant = mant.rstrip("0").rstrip(".")
    return f"{mant}p{ex}"
