"""ctypes wrapper of the CPU oracle (TEST INFRASTRUCTURE, NOT PRODUCT CODE).

Only tests/, ``__graft_entry__.smoke()`` and bench.py's cpu_baseline / ``--impl reference`` leg
may import this module.  PARITY UNPINNED: see oracle/cgo_oracle.h.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

STATUS = [
    "incomplete", "success", "increasing_objective", "max_iters_reached",
    "non_finite_objective_or_gradient_proposed", "non_descent_search_direction",
    "linesearch_a_max_overflow", "linesearch_max_iters_reached", "zoom_max_iters_reached",
    "accepted_non_finite_iterate", "cannot_find_initial_feasible_step", "max_step_length_reached",
    "cannot_find_feasible_step", "non_finite_step_proposed", "proposed_step_same_as_current_step",
    "step_bracket_precision_issue", "linesearch_failed", "infeasible_start", "centering_step_issue",
]
SUM_NAMES = {0: "seq", 1: "pairwise", 2: "comp", 3: "cgo"}
FLAVOURS = {"HagerZhang": 0, "YuanWangSheng": 1, "SallehAlhawarat": 2, "LiuStorrey": 3, "LBFGS": 4}
LS_KINDS = {"StrongWolfeBisection": 0, "Wolfe": 1, "YuanWeiLuWolfe": 2, "Backtracking": 3}
SUM_MODES = {"seq": 0, "pairwise": 1, "comp": 2, "cgo": 3}
BETA_FORMS = {"literal": 0, "fused": 1}


class OrcConfig(C.Structure):
    _fields_ = [
        ("eps", C.c_double), ("max_iters", C.c_int64), ("flavour", C.c_int32),
        ("lbfgs_m", C.c_int32), ("mu", C.c_double), ("ls_kind", C.c_int32), ("_pad", C.c_int32),
        ("c1", C.c_double), ("c2", C.c_double), ("delta1", C.c_double), ("growth", C.c_double),
        ("ls_max_iters", C.c_int64), ("zoom_max_iters", C.c_int64), ("max_step_size", C.c_double),
        ("feas_max_iters", C.c_int64), ("discount", C.c_double), ("sum_mode", C.c_int32),
        ("threads", C.c_int32), ("beta_form", C.c_int32), ("_pad2", C.c_int32),
    ]


class OrcResult(C.Structure):
    _fields_ = [
        ("objective", C.c_double), ("iters_ran", C.c_int64), ("status", C.c_int32),
        ("_pad", C.c_int32), ("trace_len", C.c_int64), ("fdf_evals_total", C.c_int64),
    ]


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libcgo_oracle.so")
    src = [os.path.join(_HERE, f) for f in ("cgo_oracle.c", "cgo_oracle.h")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libcgo_oracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    L = C.CDLL(build())
    dp = C.POINTER(C.c_double)
    L.orc_obj_booth.restype = C.c_void_p
    L.orc_obj_rosenbrock.restype = C.c_void_p
    L.orc_obj_rosenbrock.argtypes = [C.c_int64]
    L.orc_obj_rosenbrock_chained.restype = C.c_void_p
    L.orc_obj_rosenbrock_chained.argtypes = [C.c_int64]
    L.orc_obj_quartic_barrier.restype = C.c_void_p
    L.orc_obj_quartic_barrier.argtypes = [C.c_int64]
    L.orc_obj_sparse_ls_synth.restype = C.c_void_p
    L.orc_obj_sparse_ls_synth.argtypes = [C.c_int64, C.c_int32, C.c_int64, C.c_uint64, C.c_int32, C.c_int32]
    L.orc_obj_sparse_ls_csr.restype = C.c_void_p
    L.orc_obj_sparse_ls_csr.argtypes = [C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]
    L.orc_obj_logreg_synth.restype = C.c_void_p
    L.orc_obj_logreg_synth.argtypes = [C.c_int64, C.c_int64, C.c_int32, C.c_uint64, C.c_double, C.c_int32]
    L.orc_obj_box_barrier.restype = C.c_void_p
    L.orc_obj_box_barrier.argtypes = [C.c_void_p, dp, dp, C.c_double]
    L.orc_obj_barrier_set_t.argtypes = [C.c_void_p, C.c_double]
    L.orc_obj_destroy.argtypes = [C.c_void_p]
    L.orc_obj_dim.restype = C.c_int64
    L.orc_obj_dim.argtypes = [C.c_void_p]
    for name, rt in [("orc_csr_nnz", C.c_int64), ("orc_csr_nrows", C.c_int64),
                     ("orc_csr_rowptr", C.c_void_p), ("orc_csr_col", C.c_void_p),
                     ("orc_csr_val", C.c_void_p), ("orc_csr_b", C.c_void_p),
                     ("orc_csrT_rowptr", C.c_void_p), ("orc_csrT_col", C.c_void_p),
                     ("orc_csrT_val", C.c_void_p)]:
        getattr(L, name).restype = rt
        getattr(L, name).argtypes = [C.c_void_p]
    L.orc_sparse_ls_xtrue.argtypes = [C.c_int64, C.c_uint64, dp]
    L.orc_fdf.restype = C.c_double
    L.orc_fdf.argtypes = [C.c_void_p, dp, dp]
    L.orc_dot.restype = C.c_double
    L.orc_dot.argtypes = [dp, dp, C.c_int64, C.c_int, C.c_int]
    L.orc_sum.restype = C.c_double
    L.orc_sum.argtypes = [dp, C.c_int64, C.c_int, C.c_int]
    L.orc_sum_cgo.restype = C.c_double
    L.orc_sum_cgo.argtypes = [dp, C.c_int64, C.c_int, C.c_int64]
    L.orc_set_cgo_order.argtypes = [C.c_int, C.c_int]
    L.orc_obj_set_sum_mode.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.orc_obj_trial_site.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.orc_obj_set_trial_site.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.orc_set_site.argtypes = [C.c_int, C.c_int]
    L.orc_set_cgo_lanes.argtypes = [C.c_int]
    L.orc_set_cgo_lanes.restype = None
    L.orc_set_cgo_batched.argtypes = [C.c_int]
    L.orc_set_cgo_batched.restype = None
    L.orc_spmv.argtypes = [C.c_void_p, C.c_int, dp, dp]
    L.orc_hash_u01.restype = C.c_double
    L.orc_hash_u01.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64]
    L.orc_rosenbrock_x0.argtypes = [C.c_int64, C.c_uint64, C.c_double, dp]
    L.orc_minimize.restype = C.c_int
    L.orc_minimize.argtypes = [C.c_void_p, dp, C.POINTER(OrcConfig), dp, dp, C.POINTER(OrcResult),
                               dp, dp, dp, C.POINTER(C.c_int64)]
    L.orc_minimize_rerun.restype = C.c_int
    L.orc_minimize_rerun.argtypes = [C.c_void_p, dp, C.POINTER(OrcConfig), C.c_int, dp, dp,
                                     C.POINTER(OrcResult), C.c_int64, dp, dp, dp, C.POINTER(C.c_int64)]
    _LIB = L
    return L


def _dp(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def make_config(flavour="HagerZhang", linesearch="StrongWolfeBisection", eps=1e-5, max_iters=1000,
                mu=0.1, lbfgs_m=10, c1=1e-5, c2=0.8, delta1=1e-6, growth=2.0, ls_max_iters=1000,
                zoom_max_iters=100, max_step_size=1e12, feas_max_iters=50, discount=0.9,
                sum_mode="seq", threads=1, beta_form="literal") -> OrcConfig:
    """Defaults are the reference's canonical run, examples/min.jl:16-35."""
    c = OrcConfig()
    c.eps, c.max_iters, c.flavour, c.lbfgs_m, c.mu = eps, max_iters, FLAVOURS[flavour], lbfgs_m, mu
    c.ls_kind, c.c1, c.c2, c.delta1, c.growth = LS_KINDS[linesearch], c1, c2, delta1, growth
    c.ls_max_iters, c.zoom_max_iters, c.max_step_size = ls_max_iters, zoom_max_iters, max_step_size
    c.feas_max_iters, c.discount = feas_max_iters, discount
    c.sum_mode, c.threads, c.beta_form = SUM_MODES[sum_mode], threads, BETA_FORMS[beta_form]
    return c


@dataclass
class OracleResult:
    objective: float
    minimizer: np.ndarray
    gradient: np.ndarray
    iters_ran: int
    status: str
    trace_objective: np.ndarray = field(default_factory=lambda: np.zeros(0))
    trace_grad_norm: np.ndarray = field(default_factory=lambda: np.zeros(0))
    trace_step_size: np.ndarray = field(default_factory=lambda: np.zeros(0))
    trace_objective_evals: np.ndarray = field(default_factory=lambda: np.zeros(0, dtype=np.int64))
    fdf_evals_total: int = 0


class Objective:
    """Owns an ``orc_objective*``."""

    def __init__(self, handle):
        if not handle:
            raise ValueError("oracle objective constructor rejected its arguments")
        self.h = C.c_void_p(handle)
        self.n = lib().orc_obj_dim(self.h)

    def __del__(self):
        try:
            if self.h:
                lib().orc_obj_destroy(self.h)
                self.h = None
        except Exception:
            pass

    # constructors ------------------------------------------------------------------
    @staticmethod
    def booth():
        return Objective(lib().orc_obj_booth())

    @staticmethod
    def rosenbrock(n):
        return Objective(lib().orc_obj_rosenbrock(n))

    @staticmethod
    def rosenbrock_chained(n):
        return Objective(lib().orc_obj_rosenbrock_chained(n))

    @staticmethod
    def barrier(n):
        return Objective(lib().orc_obj_quartic_barrier(n))

    @staticmethod
    def box_barrier(inner: "Objective", lbs, ubs, t=1.0):
        """t·f0 − Σ log(ubs − x) − Σ log(x − lbs): evalbarrier! (primal_barrier.jl:112-133) with the
        box constraints of examples/constrained.jl:17-47"""
        lbs = np.ascontiguousarray(lbs, dtype=np.float64)
        ubs = np.ascontiguousarray(ubs, dtype=np.float64)
        assert lbs.size == inner.n == ubs.size
        o = Objective(lib().orc_obj_box_barrier(inner.h, _dp(lbs), _dp(ubs), float(t)))
        o.inner, o.lbs, o.ubs, o.t = inner, lbs, ubs, float(t)      # keeps the inner objective alive
        return o

    def set_t(self, t):
        lib().orc_obj_barrier_set_t(self.h, float(t))
        self.t = float(t)

    @staticmethod
    def sparse_ls(n, nnz_per_row=10, W=None, seed=24, coh_log2=0, threads=1):
        if W is None:
            W = min(1 << 20, (n - 1) // 2)
        return Objective(lib().orc_obj_sparse_ls_synth(n, nnz_per_row, W, seed, coh_log2, threads))

    @staticmethod
    def sparse_ls_csr(nrows, ncols, rowptr, col, val, b, threads=1):
        rowptr = np.ascontiguousarray(rowptr, dtype=np.int64)
        col = np.ascontiguousarray(col, dtype=np.int32)
        val = np.ascontiguousarray(val, dtype=np.float64)
        b = np.ascontiguousarray(b, dtype=np.float64)
        return Objective(lib().orc_obj_sparse_ls_csr(nrows, ncols, rowptr.ctypes.data, col.ctypes.data,
                                                     val.ctypes.data, b.ctypes.data, threads))

    @staticmethod
    def logreg(nsamples, nfeat, nnz_per_row=20, seed=24, lam=1e-6, threads=1):
        return Objective(lib().orc_obj_logreg_synth(nsamples, nfeat, nnz_per_row, seed, lam, threads))

    # primitives --------------------------------------------------------------------
    def set_sum_mode(self, sum_mode="seq", threads=0):
        lib().orc_obj_set_sum_mode(self.h, SUM_MODES[sum_mode], threads)

    def trial_site(self):
        v, u = C.c_int(), C.c_int()
        lib().orc_obj_trial_site(self.h, C.byref(v), C.byref(u))
        return v.value, u.value

    def set_trial_site(self, V, U):
        """canonical-order mapping of the kernels that reduce this objective's trial dots (the device
        objective reports its own: DeviceObjective.trial_site)"""
        lib().orc_obj_set_trial_site(self.h, int(V), int(U))

    def fdf(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        g = np.empty(self.n)
        f = lib().orc_fdf(self.h, _dp(g), _dp(x))
        return f, g

    def csr(self, transposed=False):
        L = lib()
        nnz = L.orc_csr_nnz(self.h)
        nrows = self.n if transposed else L.orc_csr_nrows(self.h)
        fr, fc, fv = ((L.orc_csrT_rowptr, L.orc_csrT_col, L.orc_csrT_val) if transposed
                      else (L.orc_csr_rowptr, L.orc_csr_col, L.orc_csr_val))
        rp = np.ctypeslib.as_array(C.cast(fr(self.h), C.POINTER(C.c_int64)), (nrows + 1,)).copy()
        if nnz == 0:
            return rp, np.zeros(0, dtype=np.int32), np.zeros(0)
        ci = np.ctypeslib.as_array(C.cast(fc(self.h), C.POINTER(C.c_int32)), (nnz,)).copy()
        va = np.ctypeslib.as_array(C.cast(fv(self.h), C.POINTER(C.c_double)), (nnz,)).copy()
        return rp, ci, va

    def rhs(self):
        L = lib()
        nrows = L.orc_csr_nrows(self.h)
        return np.ctypeslib.as_array(C.cast(L.orc_csr_b(self.h), C.POINTER(C.c_double)), (nrows,)).copy()

    def spmv(self, x, transposed=False):
        L = lib()
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.empty(self.n if transposed else L.orc_csr_nrows(self.h))
        L.orc_spmv(self.h, int(transposed), _dp(x), _dp(y))
        return y


def dot(a, b, sum_mode="seq", threads=1):
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    return lib().orc_dot(_dp(a), _dp(b), a.size, SUM_MODES[sum_mode], threads)


def set_cgo_lanes(B=256):
    """lanes per virtual CTA of the canonical order (default; also undoes set_cgo_batched)"""
    lib().orc_set_cgo_lanes(int(B))


def set_cgo_batched(B):
    """reduction order of the batched on-device solver: B lanes, every item in one tile"""
    lib().orc_set_cgo_batched(int(B))


def set_cgo_order(G=296, shards=1):
    """Parameters of the canonical reduction order (include/cgoptim.h): virtual CTAs, shards."""
    lib().orc_set_cgo_order(G, shards)


def sum_cgo(a, U=4, align=1):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return lib().orc_sum_cgo(_dp(a), a.size, U, align)


def set_site(V=2, U=4):
    lib().orc_set_site(V, U)


def hash_u01(seed, i, k):
    return lib().orc_hash_u01(seed, i, k)


def rosenbrock_x0(n, seed=24, perturb=0.0):
    x0 = np.empty(n)
    lib().orc_rosenbrock_x0(n, seed, perturb, _dp(x0))
    return x0


def sparse_ls_xtrue(n, seed=24):
    x = np.empty(n)
    lib().orc_sparse_ls_xtrue(n, seed, _dp(x))
    return x


def minimize(obj: Objective, x0, cfg: OrcConfig, trace=True) -> OracleResult:
    """minimizeobjective, src/engine/optim.jl:6-171."""
    L = lib()
    x0 = np.ascontiguousarray(x0, dtype=np.float64)
    n = obj.n
    assert x0.size == n
    xo, go = np.empty(n), np.empty(n)
    res = OrcResult()
    mi = max(int(cfg.max_iters), 1)
    tf, tg, ta = np.zeros(mi), np.zeros(mi), np.zeros(mi)
    te = np.zeros(mi, dtype=np.int64)
    rc = L.orc_minimize(obj.h, _dp(x0), C.byref(cfg), _dp(xo), _dp(go), C.byref(res),
                        _dp(tf) if trace else None, _dp(tg) if trace else None,
                        _dp(ta) if trace else None,
                        te.ctypes.data_as(C.POINTER(C.c_int64)) if trace else None)
    if rc != 0:
        raise AssertionError(f"oracle config assertion failed (rc={rc})")
    k = res.trace_len if trace else 0
    return OracleResult(res.objective, xo, go, res.iters_ran, STATUS[res.status], tf[:k].copy(),
                        tg[:k].copy(), ta[:k].copy(), te[:k].copy(), res.fdf_evals_total)


class OrcSolveSysLS(C.Structure):
    """LinesearchSolveSys, src/engine/solve_system.jl:7-12"""
    _fields_ = [("rho", C.c_double), ("sigma", C.c_double), ("s", C.c_double), ("max_iters", C.c_int64),
                ("fix_stale_iterate", C.c_int32), ("_pad", C.c_int32)]


def solvesys_ls(s, sigma=0.5, rho=0.95, max_iters=None, fix_stale_iterate=False) -> OrcSolveSysLS:
    """setupLinesearchSolveSys (solve_system.jl:14-26); default max_iters = round(log(ρ, 1e-6))"""
    import math
    if max_iters is None:
        max_iters = round(math.log(1e-6) / math.log(rho))
    return OrcSolveSysLS(rho, sigma, s, int(max_iters), int(bool(fix_stale_iterate)), 0)


def solvesystem(obj: Objective, x0, cfg: OrcConfig, ls: OrcSolveSysLS, trace=True) -> OracleResult:
    """solvesystem, src/engine/solve_system.jl:64-239."""
    L = lib()
    x0 = np.ascontiguousarray(x0, dtype=np.float64)
    n = obj.n
    assert x0.size == n
    xo, go = np.empty(n), np.empty(n)
    res = OrcResult()
    mi = max(int(cfg.max_iters), 1)
    tf, tg, ta = np.zeros(mi), np.zeros(mi), np.zeros(mi)
    te = np.zeros(mi, dtype=np.int64)
    L.orc_solvesystem.restype = C.c_int
    L.orc_solvesystem.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(OrcConfig), C.POINTER(OrcSolveSysLS),
                                  C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(OrcResult),
                                  C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double),
                                  C.POINTER(C.c_int64)]
    rc = L.orc_solvesystem(obj.h, _dp(x0), C.byref(cfg), C.byref(ls), _dp(xo), _dp(go), C.byref(res),
                           _dp(tf) if trace else None, _dp(tg) if trace else None, _dp(ta) if trace else None,
                           te.ctypes.data_as(C.POINTER(C.c_int64)) if trace else None)
    if rc != 0:
        raise AssertionError(f"oracle config assertion failed (rc={rc})")
    k = res.trace_len if trace else 0
    return OracleResult(res.objective, xo, go, res.iters_ran, STATUS[res.status], tf[:k].copy(),
                        tg[:k].copy(), ta[:k].copy(), te[:k].copy(), res.fdf_evals_total)


def minimize_rerun(obj: Objective, x0, cfgs) -> list[OracleResult]:
    """minimizeobjectivererun, src/engine/optim.jl:173-208; cfgs[0] primary, cfgs[1:] backups."""
    L = lib()
    x0 = np.ascontiguousarray(x0, dtype=np.float64)
    n, k = obj.n, len(cfgs)
    arr = (OrcConfig * k)(*cfgs)
    stride = max(max(int(c.max_iters) for c in cfgs), 1)
    xo, go = np.empty((k, n)), np.empty((k, n))
    res = (OrcResult * k)()
    tf, tg, ta = np.zeros((k, stride)), np.zeros((k, stride)), np.zeros((k, stride))
    te = np.zeros((k, stride), dtype=np.int64)
    nret = L.orc_minimize_rerun(obj.h, _dp(x0), arr, k, _dp(xo), _dp(go), res, stride, _dp(tf),
                                _dp(tg), _dp(ta), te.ctypes.data_as(C.POINTER(C.c_int64)))
    if nret < 0:
        raise AssertionError(f"oracle config assertion failed (rc={nret})")
    out = []
    for i in range(nret):
        t = res[i].trace_len
        out.append(OracleResult(res[i].objective, xo[i].copy(), go[i].copy(), res[i].iters_ran,
                                STATUS[res[i].status], tf[i, :t].copy(), tg[i, :t].copy(),
                                ta[i, :t].copy(), te[i, :t].copy(), res[i].fdf_evals_total))
    return out


# ---------------------------------------------------------------------------------- primal barrier
@dataclass
class OraclePrimalBarrierResults:
    """PrimalBarrierResults, src/engine/primal_barrier.jl:1-7"""
    centering_results: list
    status: str
    iters_ran: int
    t_final: float
    total_objective_evals: int


def primalbarrier(f0: Objective, lbs, ubs, x0, cfgs, barrier_tol, barrier_growth_factor, max_iters,
                  t_initial=float("nan"), update_iterate=False) -> OraclePrimalBarrierResults:
    """primalbarriermethod! (src/engine/primal_barrier.jl:158-255, Alg. 11.1 of Boyd 2004) for box
    constraints; cfgs[0] is the centering config, cfgs[1:] the rerun backups.  As written, every
    centering step starts from x_initial (`x` is never updated inside the loop, :215-247); update_iterate=True
    (not in the reference) continues from the previous centre as Alg. 11.1 does."""
    lbs, ubs = np.asarray(lbs, dtype=np.float64), np.asarray(ubs, dtype=np.float64)
    x = np.array(x0, dtype=np.float64)
    n_constraints = 2 * x.size                                                  # :176
    rets = []

    def assemble(status, it, t):                                                # :9-34
        total = sum(int(r.trace_objective_evals.sum()) for step in rets[:it] for r in step)
        return OraclePrimalBarrierResults(rets[:it], status, it, t, total)

    if np.any(x - ubs >= 0.0) or np.any(lbs - x >= 0.0):                        # :187-198
        return assemble("infeasible_start", 0, t_initial)
    t = float(t_initial)                                                        # :200, verifyt0 :259-277
    if not np.isfinite(t) or t < 0.0:
        f0.set_sum_mode(SUM_NAMES[cfgs[0].sum_mode])
        t = (f0.fdf(x)[0] - 0.0) * barrier_growth_factor
    bar = Objective.box_barrier(f0, lbs, ubs, t)
    for i in range(1, max_iters + 1):                                           # :215
        bar.set_t(t)
        rets.append(minimize_rerun(bar, x, cfgs))                               # :217-223
        if rets[-1][-1].status != "success":                                    # :224-232
            return assemble("centering_step_issue", i, t)
        if n_constraints / t < barrier_tol:                                     # :235-243
            return assemble("success", i, t)
        if update_iterate:
            x = rets[-1][-1].minimizer.copy()
        t = barrier_growth_factor * t                                           # :246
    return assemble("max_iters_reached", max_iters, t)                          # :249-254
