/*
 * cgo_oracle.h — CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * A line-by-line C restatement of the hot path of ConjugateGradientOptim.jl
 * (reference at /root/reference, citations are file:line relative to it).
 *
 * PARITY UNPINNED: Julia is not installed in the build image nor on the GPU box, and the
 * reference's only test (test/runtests.jl:7-44) pins nothing on the solver path, so this
 * restatement cannot be executed against the reference nor against reference-held golden
 * vectors.  Known-answer anchors that do exist (Booth minimiser [1,3], f*=0,
 * test/runtests.jl:18-21; Rosenbrock minimiser ones(d), examples/helpers/test_funcs.jl:48)
 * are checked in tests/test_oracle.py.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
 * load this library.  The product (conjugategradientoptim.jl_b200/) never does.
 */
#ifndef CGO_ORACLE_H
#define CGO_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* status Symbols of the reference, same numbering as include/cgoptim.h */
enum {
    ORC_INCOMPLETE = 0,                 /* src/engine/optim.jl:39 */
    ORC_SUCCESS = 1,                    /* optim.jl:63 and all line searches */
    ORC_INCREASING_OBJECTIVE = 2,       /* optim.jl:76 */
    ORC_MAX_ITERS_REACHED = 3,          /* optim.jl:168 */
    ORC_NON_FINITE_PROPOSED = 4,        /* optim.jl:118 */
    ORC_NON_DESCENT = 5,                /* nocedal.jl:62, wolfe.jl:42, geometric.jl:45 */
    ORC_A_MAX_OVERFLOW = 6,             /* nocedal.jl:148 */
    ORC_LS_MAX_ITERS = 7,               /* nocedal.jl:157, wolfe.jl:164, geometric.jl:151 */
    ORC_ZOOM_MAX_ITERS = 8,             /* nocedal.jl:208 */
    ORC_ACCEPTED_NON_FINITE = 9,        /* wolfe.jl:37, geometric.jl:40 */
    ORC_NO_INITIAL_FEASIBLE = 10,       /* wolfe.jl:64, geometric.jl:74 */
    ORC_MAX_STEP_LENGTH = 11,           /* wolfe.jl:111 */
    ORC_NO_FEASIBLE_STEP = 12,          /* wolfe.jl:157 */
    ORC_NON_FINITE_STEP = 13,           /* geometric.jl:129 */
    ORC_SAME_STEP = 14,                 /* geometric.jl:133 */
    ORC_BRACKET_PRECISION = 15,         /* wolfe.jl:131 (unreachable: missing `return`) */
    ORC_LINESEARCH_FAILED = 16,         /* solve_system.jl:140 */
    /* primalbarriermethod!, primal_barrier.jl:189, 223, 250 (plus ORC_SUCCESS, ORC_MAX_ITERS_REACHED) */
    ORC_INFEASIBLE_START = 17,
    ORC_CENTERING_STEP_ISSUE = 18
};

enum { ORC_HZ = 0, ORC_YWS = 1, ORC_SA = 2, ORC_LS = 3, ORC_LBFGS = 4 };
enum { ORC_LS_STRONGWOLFE = 0, ORC_LS_WOLFE = 1, ORC_LS_YWL = 2, ORC_LS_BACKTRACK = 3 };
enum { ORC_SUM_SEQ = 0, ORC_SUM_PAIRWISE = 1, ORC_SUM_COMP = 2, ORC_SUM_CGO = 3 };
enum { ORC_BETA_LITERAL = 0, ORC_BETA_FUSED = 1 };

typedef struct {
    /* CGConfig, src/types.jl:156-168 */
    double eps;
    int64_t max_iters;
    int32_t flavour;        /* ORC_HZ ... */
    int32_t lbfgs_m;
    double mu;              /* YuanWangSheng.μ, cg_flavours.jl:46-48 */
    /* line search */
    int32_t ls_kind;
    int32_t _pad;
    double c1, c2;
    double delta1;          /* YuanWeiLuWolfe.δ1 wolfe.jl:213-217 */
    double growth;          /* a_max_growth_factor nocedal.jl:9 */
    int64_t ls_max_iters;
    int64_t zoom_max_iters;
    double max_step_size;   /* wolfe.jl:9 */
    int64_t feas_max_iters; /* wolfe.jl:10, geometric.jl:19 */
    double discount;        /* geometric.jl:17 */
    /* oracle knobs (not in the reference) */
    int32_t sum_mode;       /* reduction order used for dot / norm */
    int32_t threads;        /* OpenMP threads for objective and reductions (1 = scalar) */
    int32_t beta_form;      /* ORC_BETA_LITERAL: cg_flavours.jl as written (elementwise tmp1·tmp2);
                               ORC_BETA_FUSED: β = (y·g⁺ − m u·g⁺)/R from the fused dot pack */
    int32_t _pad2;
} orc_config;

typedef struct {
    double objective;
    int64_t iters_ran;
    int32_t status;
    int32_t _pad;
    int64_t trace_len;
    int64_t fdf_evals_total;
} orc_result;

typedef struct orc_objective orc_objective;

/* objectives */
orc_objective *orc_obj_booth(void);
orc_objective *orc_obj_rosenbrock(int64_t n);                  /* extended (pairs), MGH #21 */
orc_objective *orc_obj_rosenbrock_chained(int64_t n);          /* test_funcs.jl:50-57 + gradient */
orc_objective *orc_obj_quartic_barrier(int64_t n);             /* returns non-finite outside a box: status coverage */
orc_objective *orc_obj_sparse_ls_synth(int64_t n, int32_t nnz_per_row, int64_t W, uint64_t seed,
                                       int32_t coh_log2, int32_t threads);
orc_objective *orc_obj_sparse_ls_csr(int64_t nrows, int64_t ncols, const int64_t *rowptr,
                                     const int32_t *col, const double *val, const double *b,
                                     int32_t threads);
orc_objective *orc_obj_logreg_synth(int64_t nsamples, int64_t nfeat, int32_t nnz_per_row,
                                    uint64_t seed, double lambda, int32_t threads);
/* box-constraint log barrier t·f0(x) − Σ log(ubs − x) − Σ log(x − lbs) around `inner`
 * (evalbarrier!, src/engine/primal_barrier.jl:112-133 with examples/constrained.jl:17-47's box) */
orc_objective *orc_obj_box_barrier(orc_objective *inner, const double *lbs, const double *ubs, double t);
void orc_obj_barrier_set_t(orc_objective *, double t);
void orc_obj_destroy(orc_objective *);
int64_t orc_obj_dim(const orc_objective *);
void orc_obj_set_sum_mode(orc_objective *, int sum_mode, int threads);
void orc_obj_trial_site(const orc_objective *, int *V, int *U);
void orc_obj_set_trial_site(orc_objective *, int V, int U);
void orc_set_site(int V, int U);   /* mapping used by orc_dot in ORC_SUM_CGO mode (default 2,4) */

/* CSR access for cross checks (pointers owned by the objective) */
int64_t orc_csr_nnz(const orc_objective *);
int64_t orc_csr_nrows(const orc_objective *);
const int64_t *orc_csr_rowptr(const orc_objective *);
const int32_t *orc_csr_col(const orc_objective *);
const double *orc_csr_val(const orc_objective *);
const double *orc_csr_b(const orc_objective *);        /* rhs (LS) or labels (logreg) */
const int64_t *orc_csrT_rowptr(const orc_objective *);
const int32_t *orc_csrT_col(const orc_objective *);
const double *orc_csrT_val(const orc_objective *);
void orc_sparse_ls_xtrue(int64_t n, uint64_t seed, double *out);

/* primitives */
double orc_fdf(orc_objective *, double *g, const double *x);
double orc_dot(const double *a, const double *b, int64_t n, int sum_mode, int threads);
double orc_sum(const double *a, int64_t n, int sum_mode, int threads);
double orc_sum_cgo(const double *a, int64_t n, int U, int64_t align);
void orc_set_cgo_order(int G, int shards);   /* canonical-order parameters (global) */
void orc_set_cgo_lanes(int B);               /* lanes per virtual CTA (default 256), BLAS-1 tiles of 4 double2 per lane */
void orc_set_cgo_batched(int B);             /* order of the batched solver: B = 32..256 lanes, all items in one tile */
void orc_spmv(const orc_objective *, int transposed, const double *x, double *y);
double orc_hash_u01(uint64_t seed, uint64_t i, uint64_t k);
void orc_rosenbrock_x0(int64_t n, uint64_t seed, double perturb, double *x0);

/* engine: minimizeobjective (src/engine/optim.jl:6-171).  Trace arrays must hold max_iters. */
int orc_minimize(orc_objective *obj, const double *x0, const orc_config *cfg, double *x_out,
                 double *g_out, orc_result *res, double *tr_f, double *tr_gnorm, double *tr_step,
                 int64_t *tr_evals);

/* LinesearchSolveSys, src/engine/solve_system.jl:7-12 */
typedef struct {
    double rho, sigma, s;
    int64_t max_iters;
    int32_t fix_stale_iterate;  /* 0: as written (solve_system.jl:172,251); 1: Alg. 3.1 of Yuan et al. as published */
    int32_t _pad;
} orc_solvesys_ls;
/* engine: solvesystem (src/engine/solve_system.jl:64-239) */
int orc_solvesystem(orc_objective *obj, const double *x0, const orc_config *cfg,
                    const orc_solvesys_ls *ls, double *x_out, double *g_out, orc_result *res,
                    double *tr_f, double *tr_gnorm, double *tr_step, int64_t *tr_evals);

/* engine: minimizeobjectivererun (optim.jl:173-208).  cfgs[0] is the primary config, cfgs[1..]
 * the backups.  res/x_out/g_out/trace arrays are laid out attempt-major with stride
 * n (vectors) and max over cfgs of max_iters (traces, `tr_stride`).  Returns #attempts. */
int orc_minimize_rerun(orc_objective *obj, const double *x0, const orc_config *cfgs, int ncfg,
                       double *x_out, double *g_out, orc_result *res, int64_t tr_stride,
                       double *tr_f, double *tr_gnorm, double *tr_step, int64_t *tr_evals);

#ifdef __cplusplus
}
#endif
#endif
