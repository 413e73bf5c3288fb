/*
 * host.c — a plain-C host of libcgoptim.so: no Python, no torch, nothing but dlopen + the entry points
 * declared in include/cgoptim.h.
 *
 * It is the executable stand-in for the Julia binding of INTEGRATION.md (Julia is not installed in this
 * image): the same call sequence julia/B200CGOptim/src/optim.jl:4-52 makes — cgo_state_create,
 * cgo_reset_direction, then per iteration the strong-Wolfe line search over cgo_eval_trial[_fused_dir],
 * Hager-Zhang β from the dot pack, cgo_accept, and the direction update deferred into the next trial.
 * Host logic restated from the reference: minimizeobjective src/engine/optim.jl:6-171, linesearch!/zoom!
 * src/linesearch/nocedal.jl:33-209, getβ(HagerZhang) src/cg_flavours.jl:87-108, updatetrace!
 * src/types.jl:56-71.
 *
 *   host <libcgoptim.so> <n> <max_iters> [expected.txt | -] [x0 perturbation, default 0] [rosenbrock | sparse_ls] [W] [coh_log2]
 *
 * prints one line per recorded iteration:  k  f  ‖g‖  a*  evals   (hex floats), then "status <sym> iters <n>".
 * With expected.txt (same line format, written by tests/test_gpu_c_host.py from tests/golden/traces.json)
 * every printed line must match bit for bit; exit code 0 = identical, 3 = mismatch.
 */
#define _POSIX_C_SOURCE 200809L
#include <dlfcn.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "../../include/cgoptim.h"

#define SYM(ret, name, args) static ret(*p_##name) args
SYM(const char *, cgo_last_error, (void));
SYM(int, cgo_ctx_create, (int, void *, cgo_ctx **));
SYM(int, cgo_ctx_destroy, (cgo_ctx *));
SYM(int, cgo_ctx_kernel_launches, (cgo_ctx *, int64_t *));
SYM(int, cgo_obj_rosenbrock_create, (cgo_ctx *, int64_t, cgo_obj **));
SYM(int, cgo_obj_sparse_ls_create_synthetic, (cgo_ctx *, int64_t, int32_t, int64_t, uint64_t, int32_t, cgo_obj **));
SYM(int, cgo_obj_destroy, (cgo_obj *));
SYM(int, cgo_obj_default_x0, (cgo_obj *, uint64_t, double, double *));
SYM(int, cgo_state_create, (cgo_ctx *, cgo_obj *, const double *, int32_t, cgo_state **, double *));
SYM(int, cgo_state_destroy, (cgo_state *));
SYM(int, cgo_reset_direction, (cgo_state *, double *));
SYM(int, cgo_eval_trial, (cgo_state *, double, double *));
SYM(int, cgo_eval_trial_fused_dir, (cgo_state *, double, double, double *));
SYM(int, cgo_accept, (cgo_state *));
SYM(int, cgo_update_dir, (cgo_state *, double, double *));
SYM(int, cgo_download, (cgo_state *, double *, double *));

static void *g_lib;
static void load(const char *path) {
    g_lib = dlopen(path, RTLD_NOW);
    if (!g_lib) { fprintf(stderr, "dlopen(%s): %s\n", path, dlerror()); exit(2); }
#define L(name) do { *(void **)(&p_##name) = dlsym(g_lib, #name); \
        if (!p_##name) { fprintf(stderr, "dlsym(%s) failed\n", #name); exit(2); } } while (0)
    L(cgo_last_error); L(cgo_ctx_create); L(cgo_ctx_destroy); L(cgo_ctx_kernel_launches);
    L(cgo_obj_rosenbrock_create); L(cgo_obj_sparse_ls_create_synthetic); L(cgo_obj_destroy); L(cgo_obj_default_x0);
    L(cgo_state_create); L(cgo_state_destroy); L(cgo_reset_direction); L(cgo_eval_trial);
    L(cgo_eval_trial_fused_dir); L(cgo_accept); L(cgo_update_dir); L(cgo_download);
}
#define CHECK(call) do { int rc__ = (call); if (rc__) { fprintf(stderr, "%s -> %d: %s\n", #call, rc__, p_cgo_last_error()); exit(2); } } while (0)

/* ---- device workspace: the LineSearchContainer of types.jl:84-100, behind a cgo_state ---- */
typedef struct {
    cgo_state *st;
    double pack[CGO_PACK_LEN];   /* last trial pack */
    double dgu, duu;             /* g·u, u·u of the current direction */
    int pending;                 /* updatedir! deferred into the next trial */
    double pending_beta;
    int cached;                  /* a fused first trial was evaluated ahead of the line search */
    double cached_a;
    long evals;
} workspace;

/* dot(df_x, u) (nocedal.jl:56): if updatedir! is still pending, run it fused with the first trial */
static double dphi0(workspace *w, double a_first) {
    if (w->pending) {
        w->pending = 0;
        if (isfinite(a_first)) {
            CHECK(p_cgo_eval_trial_fused_dir(w->st, w->pending_beta, a_first, w->pack));
            w->cached = 1; w->cached_a = a_first;
            w->dgu = w->pack[CGO_P_DIR_GU]; w->duu = w->pack[CGO_P_DIR_UU];
        } else {
            double d[CGO_PACK_LEN];
            CHECK(p_cgo_update_dir(w->st, w->pending_beta, d));
            w->dgu = d[CGO_D_GU]; w->duu = d[CGO_D_UU];
        }
    }
    return w->dgu;
}
/* evalϕdϕ! (cg_utils.jl:3-22) */
static void eval_trial(workspace *w, double a, double *phi, double *dphi) {
    if (w->cached && w->cached_a == a) {
        w->cached = 0;
    } else {
        w->cached = 0;
        CHECK(p_cgo_eval_trial(w->st, a, w->pack));
    }
    w->evals++;
    *phi = w->pack[CGO_P_PHI];
    *dphi = w->pack[CGO_P_DPHI];
}

typedef struct { double c1, c2, growth; long max_iters, zoom_max_iters; } strong_wolfe;
typedef struct { double f_xp, a_star; long evals; const char *status; } ls_result;

/* zoom! (nocedal.jl:162-209): bisection on [a_lb, a_ub] */
static ls_result zoom(workspace *w, const strong_wolfe *c, double a_lb, double a_ub, double phi_lb, double phi0,
                      double dphi_0, long evals) {
    double a = 0.0, phi = 0.0, dphi = 0.0;
    for (long it = 0; it < c->zoom_max_iters; ++it) {
        a = (a_lb + a_ub) / 2;
        eval_trial(w, a, &phi, &dphi);
        evals++;
        if (phi > phi0 + c->c1 * a * dphi_0 || phi >= phi_lb) {
            a_ub = a;
        } else {
            if (fabs(dphi) <= -c->c2 * dphi_0) return (ls_result){phi, a, evals, "success"};
            if (dphi * (a_ub - a_lb) >= 0) a_ub = a_lb;
            a_lb = a;
            phi_lb = phi;
        }
    }
    return (ls_result){phi, a, evals, "zoom_max_iters_reached"};
}
/* linesearch! (nocedal.jl:33-158) */
static ls_result linesearch(workspace *w, const strong_wolfe *c, double f_x, double a_initial) {
    if (!(0.0 < a_initial && isfinite(a_initial))) a_initial = 1.0;
    const double phi0 = f_x;
    const double dphi_0 = dphi0(w, a_initial);
    if (dphi_0 > 0.0) return (ls_result){phi0, 0.0, 0, "non_descent_search_direction"};
    double a_prev = 0.0, phi_prev = phi0, a = a_initial, phi = phi0, dphi = dphi_0;
    double a_max = a * c->growth;
    long evals = 0;
    int non_initial = 0;
    for (long it = 0; it < c->max_iters; ++it) {
        eval_trial(w, a, &phi, &dphi);
        evals++;
        const int chk1 = phi > phi0 + c->c1 * a * dphi_0;
        const int chk2 = phi >= phi_prev;
        if (chk1 || (chk2 && non_initial)) return zoom(w, c, a_prev, a, phi_prev, phi0, dphi_0, evals);
        if (fabs(dphi) <= -c->c2 * dphi_0) return (ls_result){phi, a, evals, "success"};
        if (dphi >= 0) return zoom(w, c, a, a_prev, phi, phi0, dphi_0, evals);
        a_prev = a;
        phi_prev = phi;
        non_initial = 1;
        a_max = a * c->growth;
        if (a > a_max) return (ls_result){phi, a, evals, "linesearch_a_max_overflow"};
        a = (a_max + a) / 2;
    }
    return (ls_result){phi, a, evals, "linesearch_max_iters_reached"};
}

/* getβ(HagerZhang) (cg_flavours.jl:87-108) on the dot pack: Σ (y_i − m u_i)(g⁺_i / R) = (y·g⁺ − m u·g⁺)/R */
static double beta_hager_zhang(const double *P) {
    const double R = P[CGO_P_UY];
    const double m = 2 * P[CGO_P_YY] / R;
    return (P[CGO_P_YGP] - m * P[CGO_P_DPHI]) / R;
}

int main(int argc, char **argv) {
    if (argc < 4) {
        fprintf(stderr, "usage: host <libcgoptim.so> <n> <max_iters> [expected.txt | -] [perturb] [rosenbrock | sparse_ls] [W] [coh_log2]\n");
        return 2;
    }
    load(argv[1]);
    const int64_t n = atoll(argv[2]);
    const long max_iters = atol(argv[3]);
    const int want_exp = argc > 4 && strcmp(argv[4], "-") != 0;
    FILE *expf = want_exp ? fopen(argv[4], "r") : NULL;
    if (want_exp && !expf) { fprintf(stderr, "cannot open %s\n", argv[4]); return 2; }
    const double perturb = argc > 5 ? atof(argv[5]) : 0.0;      /* SURVEY.md §8d cfg 1: x0 = (−1.2, 1, −1.2, 1, …) */
    const double eps = 1e-5;                                   /* examples/min.jl:16-35 */
    const strong_wolfe lsc = {1e-5, 0.8, 2.0, 1000, 100};
    cgo_ctx *ctx; cgo_obj *obj;
    CHECK(p_cgo_ctx_create(0, NULL, &ctx));
    /* the objective is the only thing that differs between the configurations: cfg 1 / 2 extended Rosenbrock, cfg 3
     * the banded-random CSR least squares (10 entries per row, seed 24, x0 = 0: SURVEY.md §8d) */
    const int sparse_ls = argc > 6 && strcmp(argv[6], "sparse_ls") == 0;
    if (sparse_ls) {
        const int64_t W = argc > 7 ? atoll(argv[7]) : ((int64_t)1 << 20);
        const int32_t coh = argc > 8 ? atoi(argv[8]) : 0;
        CHECK(p_cgo_obj_sparse_ls_create_synthetic(ctx, n, 10, W, 24, coh, &obj));
    } else {
        CHECK(p_cgo_obj_rosenbrock_create(ctx, n, &obj));
    }
    double *x0 = malloc(sizeof(double) * (size_t)n), *xm = malloc(sizeof(double) * (size_t)n), *gm = malloc(sizeof(double) * (size_t)n);
    if (sparse_ls) memset(x0, 0, sizeof(double) * (size_t)n);
    else CHECK(p_cgo_obj_default_x0(obj, 24, perturb, x0));
    workspace w;
    memset(&w, 0, sizeof(w));
    CHECK(p_cgo_state_create(ctx, obj, x0, 0, &w.st, w.pack));              /* optim.jl:20-26 */
    double f_x = w.pack[CGO_P_PHI], norm_df_x = sqrt(w.pack[CGO_P_GPGP]);
    const double f_x0 = f_x;
    {
        double d[CGO_PACK_LEN];
        CHECK(p_cgo_reset_direction(w.st, d));                                /* optim.jl:46 */
        w.dgu = d[CGO_D_GU]; w.duu = d[CGO_D_UU];
    }
    double a_initial = NAN;
    struct timespec t_begin, t_end;
    clock_gettime(CLOCK_MONOTONIC, &t_begin);
    const char *status = "max_iters_reached";
    long iters_ran = max_iters, mismatches = 0, lines = 0;
    char got[256], want[256];
    for (long it = 1; it <= max_iters; ++it) {                                 /* optim.jl:50 */
        if (isfinite(f_x) && isfinite(norm_df_x) && norm_df_x < eps) {      /* :53-80 */
            status = f_x <= f_x0 ? "success" : "increasing_objective";
            iters_ran = it - 1;
            break;
        }
        ls_result r = linesearch(&w, &lsc, f_x, a_initial);                   /* :83 */
        a_initial = r.a_star;                                                 /* :92 */
        if (strcmp(r.status, "success") != 0) { status = r.status; iters_ran = it - 1; break; }
        const double norm_df_xp = sqrt(w.pack[CGO_P_GPGP]);                   /* :107 */
        if (!isfinite(r.f_xp) || !isfinite(norm_df_xp)) {                     /* :108-121 */
            status = "non_finite_objective_or_gradient_proposed"; iters_ran = it - 1; break;
        }
        const double beta = beta_hager_zhang(w.pack);                          /* :130 */
        CHECK(p_cgo_accept(w.st));                                            /* :136-140 */
        f_x = r.f_xp; norm_df_x = norm_df_xp;
        w.pending = 1; w.pending_beta = beta; w.cached = 0;                   /* :145, deferred */
        snprintf(got, sizeof(got), "%ld %a %a %a %ld", it, f_x, norm_df_x, r.a_star, r.evals);   /* :152-159 */
        puts(got);
        if (expf && fgets(want, sizeof(want), expf)) {
            want[strcspn(want, "\n")] = 0;
            lines++;
            if (strcmp(got, want) != 0) { mismatches++; fprintf(stderr, "MISMATCH at iteration %ld:\n  got  %s\n  want %s\n", it, got, want); }
        }
    }
    clock_gettime(CLOCK_MONOTONIC, &t_end);       /* every call returned synchronised: the loop's wall clock is device time + host */
    const double loop_s = (double)(t_end.tv_sec - t_begin.tv_sec) + 1e-9 * (double)(t_end.tv_nsec - t_begin.tv_nsec);
    fprintf(stderr, "loop: %ld iterations in %.4f s = %.2f iterations/s\n", iters_ran, loop_s, loop_s > 0 ? (double)iters_ran / loop_s : 0.0);
    CHECK(p_cgo_download(w.st, xm, gm));
    int64_t launches = 0;
    CHECK(p_cgo_ctx_kernel_launches(ctx, &launches));
    printf("status %s iters %ld f %a x[0] %a launches %lld\n", status, iters_ran, f_x, xm[0], (long long)launches);
    CHECK(p_cgo_state_destroy(w.st));
    CHECK(p_cgo_obj_destroy(obj));
    CHECK(p_cgo_ctx_destroy(ctx));
    free(x0); free(xm); free(gm);
    if (expf) {
        fclose(expf);
        fprintf(stderr, "compared %ld trace lines, %ld mismatches\n", lines, mismatches);
        if (mismatches || lines == 0) return 3;
    }
    return 0;
}
