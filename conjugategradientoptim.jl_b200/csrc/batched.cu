// batched.cu — many independent small problems, the whole solver on the device
// (BASELINE.json configs[4]; SURVEY.md §8 cfg 5).
//
// One CTA per problem.  The host cannot drive hundreds of thousands of divergent line-search
// state machines, so this file restates on the device, scalar for scalar:
//   minimizeobjective          src/engine/optim.jl:6-171
//   linesearch! / zoom!        src/linesearch/nocedal.jl:33-209   (StrongWolfeBisection)
//   linesearch!, findfeasiblestepsize!, evalwolfeconditions ×2   src/linesearch/wolfe.jl:13-294
//   linesearch!, geometricsearch!, evalbacktrackcondition        src/linesearch/geometric.jl:22-186
//   getβ                       src/cg_flavours.jl:51-79 (YuanWangSheng), :87-108 (HagerZhang),
//                              :133-151 (SallehAlhawarat), :157-170 (LiuStorrey)
//   updatedir!, initializeLineSearchContainer!   src/cg_flavours.jl:2-35
//   evalϕdϕ!                   src/cg_utils.jl:3-22, objective = extended Rosenbrock
// Every lane runs the same scalar state machine on identical reduced scalars (uniform control
// flow, no divergence inside a problem).  The five n-vectors x, g, u, xp, g⁺ live in REGISTERS
// (up to 8 element pairs per lane); the CTA is as small as that allows — ONE warp for n <= 512 —
// so a dot product is the lane's sequential sum followed by one five-step shuffle butterfly: no
// shared memory and no barrier on the BASELINE configuration.  The reduction order is the
// canonical order of include/cgoptim.h with B = 32·nwarp lanes and one tile; the oracle
// reproduces it (orc_set_cgo_lanes), so every problem is compared bit for bit.  No HBM traffic
// between reading x0 and writing the result: the roofline that binds is the FP64 pipe, not HBM.
#include <math.h>

#include "internal.cuh"

namespace {

struct Pack9 { double v[9]; };

__device__ __forceinline__ double jl_max(double a, double b) {     // Julia max: NaN-propagating
    if (a != a) return a;
    if (b != b) return b;
    return a > b ? a : b;
}
__device__ __forceinline__ double jl_min(double a, double b) {     // Julia min: NaN-propagating
    if (a != a) return a;
    if (b != b) return b;
    return a < b ? a : b;
}

// One problem = one CTA of BT = 32·NWARP lanes; lane t owns the element pairs {q = t + BT·j,
// j < NPT} and adds its terms in that order; lanes combine by the xor-butterfly, warps in warp
// order: the canonical order of include/cgoptim.h with B = BT lanes and a single tile.  The
// launcher picks the fewest warps that keep NPT <= 8, so that n <= 512 (the BASELINE config)
// runs on ONE warp: every reduction is five shuffles, no shared memory, no barrier.
// DOTS: bit k set = pack entry k is reduced.  Every flavour needs ϕ, dϕ and ‖g⁺‖²; the other six
// dots only where its getβ reads them (an entry that is not reduced stays 0 and is never read).
constexpr int DOTS_ALL = 0x1FF;
constexpr int dots_of(int flavour) {
    return flavour == 0 || flavour == 3 ? 0x03F                               // HagerZhang, LiuStorrey: + y·y, u·y, y·g⁺
         : flavour == 1 ? 0x13F                                               // YuanWangSheng: + u·u
         : 0x0C7;                                                             // SallehAlhawarat: g⁺·g, u·g
}

template <int NPT, int NWARP, int DOTS = DOTS_ALL>
struct Problem {
    static constexpr int BT = 32 * NWARP;
    double2 x[NPT], g[NPT], u[NPT];      // xp and g⁺ are not kept: accept() recomputes them from a_last
    double a_last;
    int n;                              // problem dimension (even)
    double *scratch;                    // NWARP > 1: 9·NWARP + 2·9 doubles of shared memory
    int phase;
    int64_t evals;

    __device__ __forceinline__ bool owns(int j) const { return 2 * ((int)threadIdx.x + j * BT) < n; }

    // canonical CTA combine of the entries selected by MASK, result in every lane
    template <int K, int MASK = DOTS_ALL>
    __device__ __forceinline__ void allreduce(double (&acc)[K]) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
            if (!((MASK >> k) & 1)) continue;
            double v = acc[k];
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) v = v + __shfl_xor_sync(0xffffffffu, v, off);
            acc[k] = v;
        }
        if (NWARP == 1) return;
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        double *sm = scratch;
        double *res = scratch + 9 * NWARP + phase * 9;   // double-buffered results: no third barrier
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < K; ++k) sm[k * NWARP + warp] = acc[k];
        }
        __syncthreads();
        if ((int)threadIdx.x < K) {
            const int k = threadIdx.x;
            double t = sm[k * NWARP];
#pragma unroll
            for (int w = 1; w < NWARP; ++w) t = t + sm[k * NWARP + w];
            res[k] = t;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < K; ++k) acc[k] = res[k];
        phase ^= 1;
    }

    // f, g at x (optim.jl:25): returns f and ‖g‖²
    __device__ __forceinline__ void eval_initial(double &f, double &gg) {
        double acc[2] = {0.0, 0.0};
#pragma unroll
        for (int j = 0; j < NPT; ++j) {
            if (owns(j)) {
                const double t = x[j].y - x[j].x * x[j].x;
                const double om = 1.0 - x[j].x;
                g[j].x = (-400.0 * x[j].x) * t - 2.0 * om;
                g[j].y = 200.0 * t;
                acc[0] = acc[0] + ((100.0 * t) * t + om * om);
                acc[1] = acc[1] + g[j].x * g[j].x;
                acc[1] = acc[1] + g[j].y * g[j].y;
            }
        }
        allreduce<2>(acc);
        f = acc[0]; gg = acc[1];
        evals++;
    }
    // u = −g + βu (reset: u = −g); g·u and u·u  (cg_flavours.jl:10-12, :28; nocedal.jl:56; wolfe.jl:240)
    __device__ __forceinline__ void update_dir(double beta, bool reset, double &gu, double &uu) {
        double acc[2] = {0.0, 0.0};
#pragma unroll
        for (int j = 0; j < NPT; ++j) {
            if (owns(j)) {
                if (reset) { u[j].x = -g[j].x; u[j].y = -g[j].y; }
                else { u[j].x = -g[j].x + beta * u[j].x; u[j].y = -g[j].y + beta * u[j].y; }
                acc[0] = acc[0] + g[j].x * u[j].x;
                acc[1] = acc[1] + u[j].x * u[j].x;
                acc[0] = acc[0] + g[j].y * u[j].y;
                acc[1] = acc[1] + u[j].y * u[j].y;
            }
        }
        allreduce<2>(acc);
        gu = acc[0]; uu = acc[1];
    }
    // ‖u + df_x‖² (wolfe.jl:123)
    __device__ __forceinline__ double norm_sq_u_plus_g() {
        double acc[1] = {0.0};
#pragma unroll
        for (int j = 0; j < NPT; ++j) {
            if (owns(j)) {
                const double t1 = u[j].x + g[j].x, t2 = u[j].y + g[j].y;
                acc[0] = acc[0] + t1 * t1;
                acc[0] = acc[0] + t2 * t2;
            }
        }
        allreduce<1>(acc);
        return acc[0];
    }
    // evalϕdϕ! (cg_utils.jl:3-22) + the dot pack of the trial kernel (same terms, same order)
    __device__ __forceinline__ void eval_trial(double a, Pack9 &P) {
        double acc[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) acc[k] = 0.0;
#pragma unroll
        for (int j = 0; j < NPT; ++j) {
            if (owns(j)) {
                double2 p;
                p.x = x[j].x + a * u[j].x;
                p.y = x[j].y + a * u[j].y;
                const double t = p.y - p.x * p.x;
                const double om = 1.0 - p.x;
                const double f = (100.0 * t) * t + om * om;
                double2 gn;
                gn.x = (-400.0 * p.x) * t - 2.0 * om;
                gn.y = 200.0 * t;
                const double y1 = gn.x - g[j].x, y2 = gn.y - g[j].y;
                acc[CGO_P_PHI] = acc[CGO_P_PHI] + f;
                acc[CGO_P_DPHI] = acc[CGO_P_DPHI] + gn.x * u[j].x;   acc[CGO_P_DPHI] = acc[CGO_P_DPHI] + gn.y * u[j].y;
                acc[CGO_P_GPGP] = acc[CGO_P_GPGP] + gn.x * gn.x;     acc[CGO_P_GPGP] = acc[CGO_P_GPGP] + gn.y * gn.y;
                if (DOTS & (1 << CGO_P_YY)) { acc[CGO_P_YY] = acc[CGO_P_YY] + y1 * y1; acc[CGO_P_YY] = acc[CGO_P_YY] + y2 * y2; }
                if (DOTS & (1 << CGO_P_UY)) { acc[CGO_P_UY] = acc[CGO_P_UY] + u[j].x * y1; acc[CGO_P_UY] = acc[CGO_P_UY] + u[j].y * y2; }
                if (DOTS & (1 << CGO_P_YGP)) { acc[CGO_P_YGP] = acc[CGO_P_YGP] + y1 * gn.x; acc[CGO_P_YGP] = acc[CGO_P_YGP] + y2 * gn.y; }
                if (DOTS & (1 << CGO_P_GPG)) { acc[CGO_P_GPG] = acc[CGO_P_GPG] + gn.x * g[j].x; acc[CGO_P_GPG] = acc[CGO_P_GPG] + gn.y * g[j].y; }
                if (DOTS & (1 << CGO_P_UG)) { acc[CGO_P_UG] = acc[CGO_P_UG] + u[j].x * g[j].x; acc[CGO_P_UG] = acc[CGO_P_UG] + u[j].y * g[j].y; }
                if (DOTS & (1 << CGO_P_UU)) { acc[CGO_P_UU] = acc[CGO_P_UU] + u[j].x * u[j].x; acc[CGO_P_UU] = acc[CGO_P_UU] + u[j].y * u[j].y; }
            }
        }
        allreduce<9, DOTS>(acc);
#pragma unroll
        for (int k = 0; k < 9; ++k) P.v[k] = acc[k];
        evals++;
        a_last = a;
    }
    // x[:] = info.xp; df_x[:] = info.df_xp (optim.jl:136-140): info.xp / info.df_xp are those of the
    // LAST evaluated trial (which Backtracking's :success need not be), re-formed here with the very
    // expressions eval_trial used — bit-identical, and two n-vectors fewer in registers
    __device__ __forceinline__ void accept() {
#pragma unroll
        for (int j = 0; j < NPT; ++j) {
            if (owns(j)) {
                double2 p;
                p.x = x[j].x + a_last * u[j].x;
                p.y = x[j].y + a_last * u[j].y;
                const double t = p.y - p.x * p.x;
                const double om = 1.0 - p.x;
                x[j] = p;
                g[j].x = (-400.0 * p.x) * t - 2.0 * om;
                g[j].y = 200.0 * t;
            }
        }
    }
};

// ---------------------------------------------------------------- StrongWolfeBisection
// zoom! (nocedal.jl:162-209)
template <class PB>
__device__ int zoom(PB &S, Pack9 &P, const cgo_batched_config &c, double a_lb, double a_ub,
                    double phi_lb, double phi0, double dphi0, int64_t evals, double &phi_out, double &a_out,
                    int64_t &evals_out) {
    double a = 0.0, phi_a = 0.0, dphi_a = 0.0;
    for (int64_t it = 0; it < c.zoom_max_iters; ++it) {
        a = (a_lb + a_ub) / 2;                                              // :187
        S.eval_trial(a, P);                                                 // :190
        phi_a = P.v[CGO_P_PHI]; dphi_a = P.v[CGO_P_DPHI];
        evals += 1;
        if ((phi_a > phi0 + c.c1 * a * dphi0) || (phi_a >= phi_lb)) {       // :193
            a_ub = a;
        } else {
            if (fabs(dphi_a) <= -c.c2 * dphi0) {                            // :196
                phi_out = phi_a; a_out = a; evals_out = evals;
                return CGO_ST_SUCCESS;
            }
            if (dphi_a * (a_ub - a_lb) >= 0) a_ub = a_lb;                   // :200
            a_lb = a;
            phi_lb = phi_a;
        }
    }
    phi_out = phi_a; a_out = a; evals_out = evals;
    return CGO_ST_ZOOM_MAX_ITERS;                                           // :208
}

// linesearch! (nocedal.jl:33-158); dphi0 = dot(df_x, u) was reduced by update_dir
template <class PB>
__device__ int ls_strong_wolfe(PB &S, Pack9 &P, const cgo_batched_config &c, double f_x, double dphi0,
                               double a_initial, double &phi_out, double &a_out, int64_t &evals_out) {
    if (!(0.0 < a_initial && isfinite(a_initial))) a_initial = 1.0;        // :49-52
    const double phi0 = f_x;
    if (dphi0 > 0.0) {                                                      // :57-63
        phi_out = phi0; a_out = 0.0; evals_out = 0;
        return CGO_ST_NON_DESCENT;
    }
    double a_prev = 0.0, phi_prev = phi0;
    double a = a_initial, phi_a = phi0, dphi_a = dphi0;
    double a_max = a * c.growth;
    int64_t evals = 0;
    bool non_initial = false;
    for (int64_t it = 0; it < c.ls_max_iters; ++it) {                       // :76
        S.eval_trial(a, P);                                                 // :78
        phi_a = P.v[CGO_P_PHI]; dphi_a = P.v[CGO_P_DPHI];
        evals += 1;
        const bool chk1 = phi_a > phi0 + c.c1 * a * dphi0;                  // :81
        const bool chk2 = phi_a >= phi_prev;                                // :82
        if (chk1 || (chk2 && non_initial))                                  // :83-105
            return zoom(S, P, c, a_prev, a, phi_prev, phi0, dphi0, evals, phi_out, a_out, evals_out);
        if (fabs(dphi_a) <= -c.c2 * dphi0) {                                // :107-110
            phi_out = phi_a; a_out = a; evals_out = evals;
            return CGO_ST_SUCCESS;
        }
        if (dphi_a >= 0)                                                    // :112-134
            return zoom(S, P, c, a, a_prev, phi_a, phi0, dphi0, evals, phi_out, a_out, evals_out);
        a_prev = a; phi_prev = phi_a; non_initial = true;                   // :137-139
        a_max = a * c.growth;                                               // :143
        if (a > a_max) {                                                    // :144-149
            phi_out = phi_a; a_out = a; evals_out = evals;
            return CGO_ST_A_MAX_OVERFLOW;
        }
        a = (a_max + a) / 2;                                                // :150
    }
    phi_out = phi_a; a_out = a; evals_out = evals;
    return CGO_ST_LS_MAX_ITERS;                                             // :157
}

// ---------------------------------------------------------------- WolfeBisection
// findfeasiblestepsize! (wolfe.jl:171-207): 0 feasible, 1 infeasible, 2 lower bound above the step
template <class PB>
__device__ int find_feasible(PB &S, Pack9 &P, int64_t &evals, double &a, double reduction, double lb,
                             int64_t max_iters, double &phi_a, double &dphi_a) {
    if (lb > a) { phi_a = 0.0; dphi_a = 0.0; return 2; }                    // :186-188
    S.eval_trial(a, P);                                                     // :191
    phi_a = P.v[CGO_P_PHI]; dphi_a = P.v[CGO_P_DPHI];
    evals += 1;
    int64_t iter = 1;
    while (a > lb && iter < max_iters) {                                    // :195
        if (isfinite(phi_a) && isfinite(dphi_a)) return 0;
        a = a * reduction;                                                  // :200
        S.eval_trial(a, P);
        phi_a = P.v[CGO_P_PHI]; dphi_a = P.v[CGO_P_DPHI];
        evals += 1;
        iter += 1;
    }
    return 1;                                                               // :206
}
// evalwolfeconditions (wolfe.jl:219-251 YuanWeiLuWolfe, :264-294 Wolfe); uu = dot(u, u) (:240)
__device__ __forceinline__ void eval_wolfe(const cgo_batched_config &c, double phi_a, double dphi_a, double a,
                                           double uu, double phi0, double dphi0, bool &valid_large, bool &valid_small) {
    if (c.linesearch == 2) {
        const double t1 = -c.delta1 * dphi0, t2 = c.c1 * a * uu / 2;
        const double rhs1 = phi0 + c.c1 * a * dphi0 + a * jl_min(t1, t2);   // :243
        valid_large = phi_a <= rhs1;
        const double t3 = c.c1 * a * uu;
        const double rhs2 = c.c2 * dphi0 + jl_min(t1, t3);                  // :247
        valid_small = dphi_a >= rhs2;
    } else {
        valid_large = phi_a <= phi0 + c.c1 * a * dphi0;                     // :285-286
        valid_small = dphi_a >= c.c2 * dphi0;                               // :289-290
    }
}
// linesearch! (wolfe.jl:13-165), quirks included: the reset u ← −df_x at :123-129 keeps the old
// dϕ_0, and the tuple built at :131 is not returned
template <class PB>
__device__ int ls_wolfe_bisection(PB &S, Pack9 &P, const cgo_batched_config &c, double f_x, double dphi0, double uu,
                                  double a_initial, double &phi_out, double &a_out, int64_t &evals_out) {
    const double reduction = 0.5, growth = 2.0;                             // :23-24
    const double max_step = c.max_step_size;
    if (!(max_step > a_initial && a_initial > 0.0)) a_initial = fmin(1.0, max_step / 2);   // :30-32
    const double phi0 = f_x;
    phi_out = phi0; a_out = 0.0; evals_out = 0;
    if (!isfinite(phi0)) return CGO_ST_ACCEPTED_NON_FINITE;                 // :36-38
    if (dphi0 > 0.0) return CGO_ST_NON_DESCENT;                             // :40-47
    double a = a_initial;
    int64_t evals = 0;
    double lb = 0.0, ub = INFINITY;
    double phi_a, dphi_a;
    int st = find_feasible(S, P, evals, a, reduction, 0.0, c.feas_max_iters, phi_a, dphi_a);   // :51-62
    if (st != 0) return CGO_ST_NO_INITIAL_FEASIBLE;
    for (int64_t it = 0; it < c.ls_max_iters; ++it) {                       // :67
        bool vl, vs;
        eval_wolfe(c, phi_a, dphi_a, a, uu, phi0, dphi0, vl, vs);           // :70-78
        if (!vl || !vs) {
            if (!vl) {
                ub = a;                                                     // :86
                a = (lb + ub) / 2;                                          // :95
            } else {
                lb = a;                                                     // :98
                if (!isfinite(ub)) {
                    a = growth * a;                                         // :102
                    if (a > max_step) return CGO_ST_MAX_STEP_LENGTH;        // :104-112
                } else {
                    a = (lb + ub) / 2;                                      // :114
                }
            }
            if (!(lb < a && a < ub)) {                                      // :122
                const double nrm = sqrt(S.norm_sq_u_plus_g());              // :123
                if (!(nrm == 0.0)) {
                    lb = 0.0; ub = INFINITY;
                    a = a_initial;
                    double gu_unused;
                    S.update_dir(0.0, true, gu_unused, uu);                 // :129  u[:] = −df_x (dϕ_0 kept)
                }
            }
            st = find_feasible(S, P, evals, a, reduction, lb, c.feas_max_iters, phi_a, dphi_a);   // :141-152
            if (st != 0) return CGO_ST_NO_FEASIBLE_STEP;                    // :157
        } else {
            phi_out = phi_a; a_out = a; evals_out = evals;                  // :160
            return CGO_ST_SUCCESS;
        }
    }
    phi_out = phi_a; a_out = a; evals_out = evals;
    return CGO_ST_LS_MAX_ITERS;                                             // :164
}

// ---------------------------------------------------------------- Backtracking (Armijo)
__device__ __forceinline__ bool eval_armijo(double c1, double phi_a, double a, double phi0, double dphi0) {
    if (!isfinite(phi0) || !isfinite(phi_a) || !isfinite(a)) return false;  // geometric.jl:177-179
    return (phi0 - phi_a) >= -c1 * a * dphi0;                               // :182-183
}
// linesearch! + geometricsearch! (geometric.jl:22-152), bug-for-bug: :success returns the PREVIOUS
// (ϕ, a) while xp / g⁺ hold the rejected trial, which the engine then adopts (optim.jl:136-139)
template <class PB>
__device__ int ls_backtracking(PB &S, Pack9 &P, const cgo_batched_config &c, double f_x, double dphi0, double uu,
                               double a_initial, double &phi_out, double &a_out, int64_t &evals_out) {
    const double phi0 = f_x;
    phi_out = phi0; a_out = 0.0; evals_out = 0;
    if (!isfinite(phi0)) return CGO_ST_ACCEPTED_NON_FINITE;                 // :39-41
    if (dphi0 > 0.0) return CGO_ST_NON_DESCENT;                             // :43-48
    int64_t evals = 0;
    double a = a_initial;
    if (!isfinite(a)) a = fabs(phi0) / uu;                                  // :50-53
    if (!isfinite(a)) a = 1.0;                                              // :54-57
    double phi_a, dphi_a;
    int st = find_feasible(S, P, evals, a, 0.5, 0.0, c.feas_max_iters, phi_a, dphi_a);   // :60-71
    if (st != 0) return CGO_ST_NO_INITIAL_FEASIBLE;                         // :74
    S.eval_trial(a, P);                                                     // :78 (redundant re-evaluation)
    phi_a = P.v[CGO_P_PHI];
    evals += 1;
    const bool divide = eval_armijo(c.c1, phi_a, a, phi0, dphi0);           // :81-97
    double a_prev = a, phi_prev = phi_a;
    for (int64_t it = 0; it < c.ls_max_iters; ++it) {                       // geometricsearch! :102-152
        a = divide ? a / c.discount : a * c.discount;                       // :127
        if (!isfinite(a)) { phi_out = phi_prev; a_out = a_prev; evals_out = evals; return CGO_ST_NON_FINITE_STEP; }
        if (a == a_prev) { phi_out = phi_prev; a_out = a_prev; evals_out = evals; return CGO_ST_SAME_STEP; }
        S.eval_trial(a, P);                                                 // :137
        phi_a = P.v[CGO_P_PHI];
        evals += 1;
        if (!eval_armijo(c.c1, phi_a, a, phi0, dphi0)) {                    // :140-144
            phi_out = phi_prev; a_out = a_prev; evals_out = evals;
            return CGO_ST_SUCCESS;
        }
        a_prev = a; phi_prev = phi_a;
    }
    phi_out = phi_a; a_out = a; evals_out = evals;
    return CGO_ST_LS_MAX_ITERS;                                             // :151
}

template <class PB>
__device__ __forceinline__ int linesearch(PB &S, Pack9 &P, const cgo_batched_config &c, double f_x, double dphi0,
                                          double uu, double a_initial, double &phi_out, double &a_out,
                                          int64_t &evals_out) {
    if (c.linesearch == 0) return ls_strong_wolfe(S, P, c, f_x, dphi0, a_initial, phi_out, a_out, evals_out);
    if (c.linesearch == 3) return ls_backtracking(S, P, c, f_x, dphi0, uu, a_initial, phi_out, a_out, evals_out);
    return ls_wolfe_bisection(S, P, c, f_x, dphi0, uu, a_initial, phi_out, a_out, evals_out);
}

// getβ from the dot pack (the single-pass forms of conjugategradientoptim.jl_b200/cg_flavours.py)
__device__ __forceinline__ double get_beta(const cgo_batched_config &c, const Pack9 &P) {
    const double *v = P.v;
    switch (c.flavour) {
    case 1: {                                                               // YuanWangSheng :51-79
        const double R1 = c.mu * sqrt(v[CGO_P_UU]) * sqrt(v[CGO_P_YY]);     // :65
        const double R2 = v[CGO_P_UY];                                      // :66
        const double R3 = 2 * v[CGO_P_YY] * v[CGO_P_DPHI] / v[CGO_P_YGP];   // :67
        const double R = jl_max(jl_max(R1, R2), R3);                        // :68
        const double m = 2 * v[CGO_P_YY] / R;                               // :73
        return (v[CGO_P_YGP] - m * v[CGO_P_DPHI]) / R;
    }
    case 2: {                                                               // SallehAlhawarat :133-151
        const double nrm = sqrt(v[CGO_P_GPGP]);
        const double norm_sq = nrm * nrm;                                   // :140
        const double tmp = v[CGO_P_GPG];                                    // :141
        if (norm_sq > tmp) return (norm_sq - tmp) / (v[CGO_P_DPHI] - v[CGO_P_UG]);   // :145
        return 0.0;
    }
    case 3:                                                                 // LiuStorrey :157-170
        return v[CGO_P_YGP] / (-v[CGO_P_UY]);
    default: {                                                              // HagerZhang :87-108
        const double R = v[CGO_P_UY];                                       // :98
        const double m = 2 * v[CGO_P_YY] / R;                               // :102
        return (v[CGO_P_YGP] - m * v[CGO_P_DPHI]) / R;
    }
    }
}

struct BatchedOut {
    double *objective, *minimizer, *grad_norm;
    int64_t *iters_ran, *fdf_evals;
    int32_t *status;
};

constexpr int batched_occ(int npt, int nwarp) {       // CTAs per SM the register budget allows
    return nwarp == 1 ? (npt <= 4 ? 16 : 12) : (nwarp == 2 ? 6 : (nwarp == 4 ? 3 : 1));
}

// Problems are handed out through an atomic counter (their iteration counts differ by two orders
// of magnitude: a static split leaves most of the grid idle behind the slowest CTAs).
template <int NPT, int NWARP, int DOTS>
__global__ void __launch_bounds__(32 * NWARP, batched_occ(NPT, NWARP))
k_batched_rosenbrock(int64_t nprob, int n, const double *__restrict__ x0, cgo_batched_config c, BatchedOut out,
                     unsigned long long *next_problem) {
    constexpr int BT = 32 * NWARP;
    __shared__ double scratch[NWARP > 1 ? 9 * NWARP + 2 * 9 : 1];
    __shared__ unsigned long long s_prob;
    for (;;) {
        unsigned long long pr = 0;
        if (NWARP == 1) {
            if (threadIdx.x == 0) pr = atomicAdd(next_problem, 1ULL);
            pr = __shfl_sync(0xffffffffu, pr, 0);
        } else {
            if (threadIdx.x == 0) s_prob = atomicAdd(next_problem, 1ULL);
            __syncthreads();
            pr = s_prob;
        }
        const int64_t prob = (int64_t)pr;
        if (prob >= nprob) break;
        Problem<NPT, NWARP, DOTS> S;
        S.n = n; S.scratch = scratch; S.phase = 0; S.evals = 0;
        const double2 *xin = reinterpret_cast<const double2 *>(x0 + prob * (int64_t)n);
#pragma unroll
        for (int j = 0; j < NPT; ++j) {
            if (S.owns(j)) S.x[j] = xin[threadIdx.x + j * BT];              // optim.jl:21 x = copy(x_initial)
        }
        double f_x, gg;
        S.eval_initial(f_x, gg);                                            // :25
        double norm_g = sqrt(gg);                                           // :26
        const double f_x0 = f_x;                                            // :31
        double dphi0, uu;
        S.update_dir(0.0, true, dphi0, uu);                                 // :46  u = −g
        double a_initial = nan("");                                         // :47
        int status = CGO_ST_MAX_ITERS_REACHED;
        int64_t iters_ran = c.max_iters;
        Pack9 P;
        for (int64_t it = 1; it <= c.max_iters; ++it) {                     // :50
            if (isfinite(f_x) && isfinite(norm_g) && norm_g < c.eps) {      // :53-80
                status = (f_x <= f_x0) ? CGO_ST_SUCCESS : CGO_ST_INCREASING_OBJECTIVE;
                iters_ran = it - 1;
                break;
            }
            double f_xp, a_star;
            int64_t evals;
            const int st = linesearch(S, P, c, f_x, dphi0, uu, a_initial, f_xp, a_star, evals);   // :83
            a_initial = a_star;                                             // :92
            if (st != CGO_ST_SUCCESS) { status = st; iters_ran = it - 1; break; }            // :93-104
            const double norm_gp = sqrt(P.v[CGO_P_GPGP]);                   // :107
            if (!isfinite(f_xp) || !isfinite(norm_gp)) {                    // :108-121
                status = CGO_ST_NON_FINITE_PROPOSED; iters_ran = it - 1; break;
            }
            const double beta = get_beta(c, P);                             // :130-135
            S.accept();                                                     // :136-140
            f_x = f_xp; norm_g = norm_gp;                                   // :138, :141
            S.update_dir(beta, false, dphi0, uu);                           // :145
        }
        // Results (types.jl:107-151)
        double2 *xout = out.minimizer ? reinterpret_cast<double2 *>(out.minimizer + prob * (int64_t)n) : nullptr;
        if (xout) {
#pragma unroll
            for (int j = 0; j < NPT; ++j) {
                if (S.owns(j)) xout[threadIdx.x + j * BT] = S.x[j];
            }
        }
        if (threadIdx.x == 0) {
            if (out.objective) out.objective[prob] = f_x;
            if (out.grad_norm) out.grad_norm[prob] = norm_g;
            if (out.iters_ran) out.iters_ran[prob] = iters_ran;
            if (out.status) out.status[prob] = status;
            if (out.fdf_evals) out.fdf_evals[prob] = S.evals;
        }
        if (NWARP > 1) __syncthreads();
    }
}

template <int NPT, int NWARP, int DOTS>
static void launch_batched_dots(cgo_ctx *ctx, int64_t nprob, int n, const double *d_x0, const cgo_batched_config &cfg,
                                const BatchedOut &o, unsigned long long *counter) {
    const int64_t cap = (int64_t)ctx->sms * batched_occ(NPT, NWARP);      // every resident slot, once
    const int grid = (int)(nprob < cap ? nprob : cap);
    k_batched_rosenbrock<NPT, NWARP, DOTS><<<grid, 32 * NWARP, 0, ctx->stream>>>(nprob, n, d_x0, cfg, o, counter);
}
template <int NPT, int NWARP>
static void launch_batched(cgo_ctx *ctx, int64_t nprob, int n, const double *d_x0, const cgo_batched_config &cfg,
                           const BatchedOut &o, unsigned long long *counter) {
    if constexpr (NPT == 8 && NWARP == 1) {         // the BASELINE layout (n <= 512): only the dots its flavour reads
        switch (cfg.flavour) {
        case 1: return launch_batched_dots<NPT, NWARP, dots_of(1)>(ctx, nprob, n, d_x0, cfg, o, counter);
        case 2: return launch_batched_dots<NPT, NWARP, dots_of(2)>(ctx, nprob, n, d_x0, cfg, o, counter);
        default: return launch_batched_dots<NPT, NWARP, dots_of(0)>(ctx, nprob, n, d_x0, cfg, o, counter);
        }
    } else {
        launch_batched_dots<NPT, NWARP, DOTS_ALL>(ctx, nprob, n, d_x0, cfg, o, counter);
    }
}

template <class T>
int dev_alloc(T **p, size_t count) {
    CGO_CUDA(cudaMalloc(p, sizeof(T) * (count > 0 ? count : 1)));
    return 0;
}

}  // namespace

// lanes per problem: the fewest warps (1, 2, 4, 8) that keep at most 8 element pairs per lane
extern "C" int cgo_batched_layout(int32_t n, int32_t *nwarp, int32_t *pairs_per_lane) {
    CGO_CHECK(nwarp && pairs_per_lane && n >= 2, "bad arguments");
    const int pairs = (n + 1) / 2;
    int w = 1;
    while (w < 8 && pairs > 32 * w * 8) w *= 2;
    *nwarp = w;
    *pairs_per_lane = (pairs + 32 * w - 1) / (32 * w);
    return 0;
}

extern "C" int cgo_batched_minimize_rosenbrock(cgo_ctx *ctx, int64_t nprob, int32_t n, const double *x0,
                                               const cgo_batched_config *cfg, double *objective,
                                               int64_t *iters_ran, int32_t *status, int64_t *fdf_evals,
                                               double *minimizer, double *grad_norm) {
    CGO_CHECK(ctx && x0 && cfg, "NULL argument");
    CGO_CHECK(nprob >= 0, "nprob < 0");
    CGO_CHECK(n >= 2 && n % 2 == 0 && n <= 4096, "n must be even and in [2, 4096]");
    CGO_CHECK(0.0 < cfg->eps && cfg->eps < 1.0, "need 0 < eps < 1 (types.jl:187)");
    CGO_CHECK(cfg->linesearch >= 0 && cfg->linesearch <= 3, "linesearch %d out of [0,3]", cfg->linesearch);
    if (cfg->linesearch == 0)
        CGO_CHECK(0.0 < cfg->c1 && cfg->c1 < cfg->c2 && cfg->c2 < 1.0 && cfg->growth > 1.0 && cfg->ls_max_iters >= 0 &&
                      cfg->zoom_max_iters >= 0, "StrongWolfeBisection config asserts failed (nocedal.jl:22-26)");
    else if (cfg->linesearch == 1)
        CGO_CHECK(0.0 < cfg->c1 && cfg->c1 < cfg->c2 && cfg->c2 < 1.0, "Wolfe needs 0 < c1 < c2 < 1 (wolfe.jl:278)");
    else if (cfg->linesearch == 2)
        CGO_CHECK(0.0 < cfg->delta1 && cfg->delta1 < cfg->c1 && cfg->c1 < cfg->c2 && cfg->c2 < 1.0,
                  "YuanWeiLuWolfe needs 0 < delta1 < c1 < c2 < 1 (wolfe.jl:233)");
    else
        CGO_CHECK(0.0 < cfg->c1 && cfg->c1 < 1.0 && 0.0 < cfg->discount && cfg->discount < 1.0,
                  "Armijo needs 0 < c1 < 1 (geometric.jl:174) and a discount factor in (0,1)");
    CGO_CHECK(cfg->flavour >= 0 && cfg->flavour <= 3, "flavour %d out of [0,3]", cfg->flavour);
    if (nprob == 0) return 0;
    CGO_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    const size_t nx = (size_t)nprob * (size_t)n;
    double *d_x0 = nullptr;
    BatchedOut o{};
    int rc = 0;
    auto body = [&]() -> int {
        CGO_TRY(dev_alloc(&d_x0, nx));
        CGO_CUDA(cudaMemcpyAsync(d_x0, x0, sizeof(double) * nx, cudaMemcpyHostToDevice, s));
        if (objective) CGO_TRY(dev_alloc(&o.objective, (size_t)nprob));
        if (grad_norm) CGO_TRY(dev_alloc(&o.grad_norm, (size_t)nprob));
        if (iters_ran) CGO_TRY(dev_alloc(&o.iters_ran, (size_t)nprob));
        if (fdf_evals) CGO_TRY(dev_alloc(&o.fdf_evals, (size_t)nprob));
        if (status) CGO_TRY(dev_alloc(&o.status, (size_t)nprob));
        if (minimizer) CGO_TRY(dev_alloc(&o.minimizer, nx));
        int nwarp, npt;
        cgo_batched_layout(n, &nwarp, &npt);
        unsigned long long *counter = reinterpret_cast<unsigned long long *>(ctx->d_ticket + 2);
        CGO_CUDA(cudaMemsetAsync(counter, 0, sizeof(unsigned long long), s));
        cgo_timer_begin(ctx, CGO_T_BATCHED);
        if (nwarp == 1 && npt <= 1) launch_batched<1, 1>(ctx, nprob, n, d_x0, *cfg, o, counter);
        else if (nwarp == 1 && npt <= 2) launch_batched<2, 1>(ctx, nprob, n, d_x0, *cfg, o, counter);
        else if (nwarp == 1 && npt <= 4) launch_batched<4, 1>(ctx, nprob, n, d_x0, *cfg, o, counter);
        else if (nwarp == 1) launch_batched<8, 1>(ctx, nprob, n, d_x0, *cfg, o, counter);
        else if (nwarp == 2) launch_batched<8, 2>(ctx, nprob, n, d_x0, *cfg, o, counter);
        else if (nwarp == 4) launch_batched<8, 4>(ctx, nprob, n, d_x0, *cfg, o, counter);
        else launch_batched<8, 8>(ctx, nprob, n, d_x0, *cfg, o, counter);
        cgo_timer_end(ctx);
        ctx->launches++;
        CGO_CUDA(cudaGetLastError());
        if (objective) CGO_CUDA(cudaMemcpyAsync(objective, o.objective, sizeof(double) * nprob, cudaMemcpyDeviceToHost, s));
        if (grad_norm) CGO_CUDA(cudaMemcpyAsync(grad_norm, o.grad_norm, sizeof(double) * nprob, cudaMemcpyDeviceToHost, s));
        if (iters_ran) CGO_CUDA(cudaMemcpyAsync(iters_ran, o.iters_ran, sizeof(int64_t) * nprob, cudaMemcpyDeviceToHost, s));
        if (fdf_evals) CGO_CUDA(cudaMemcpyAsync(fdf_evals, o.fdf_evals, sizeof(int64_t) * nprob, cudaMemcpyDeviceToHost, s));
        if (status) CGO_CUDA(cudaMemcpyAsync(status, o.status, sizeof(int32_t) * nprob, cudaMemcpyDeviceToHost, s));
        if (minimizer) CGO_CUDA(cudaMemcpyAsync(minimizer, o.minimizer, sizeof(double) * nx, cudaMemcpyDeviceToHost, s));
        CGO_CUDA(cudaStreamSynchronize(s));
        if (ctx->timing) cgo_timer_collect(ctx);
        return 0;
    };
    rc = body();
    cudaFree(d_x0); cudaFree(o.objective); cudaFree(o.grad_norm); cudaFree(o.iters_ran);
    cudaFree(o.fdf_evals); cudaFree(o.status); cudaFree(o.minimizer);
    return rc;
}
