// reduce.cuh — the canonical, deterministic reduction (spec: include/cgoptim.h).
//
// Replaces every LinearAlgebra.dot / norm call site of the reference (SURVEY.md §2 last row:
// src/cg_utils.jl:20; src/engine/optim.jl:26,107; src/cg_flavours.jl:65-67,73,76,98,102,105,
// 140-141,145,166-167; src/linesearch/nocedal.jl:56; wolfe.jl:40,123,240; geometric.jl:43,52).
// Warp-shuffle butterfly + fixed-order CTA combine + last-block finish: no atomics on data,
// run-to-run bitwise reproducible, independent of the physical grid.
#pragma once
#include "internal.cuh"

// CTA-wide barrier over the CGO_B reducing lanes.  BarAll: the whole CTA reduces.  BarLanes:
// only threads [0, CGO_B) do (warp-specialised kernels keep their producer warp out of it).
struct BarAll { __device__ __forceinline__ void operator()() const { __syncthreads(); } };
struct BarLanes {
    __device__ __forceinline__ void operator()() const { asm volatile("bar.sync 1, %0;" ::"n"(CGO_B) : "memory"); }
};

__device__ __forceinline__ void cgo_st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long cgo_ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// Combine the CGO_B lanes of a CTA.  Result valid in thread 0 (acc[] of thread 0).
template <int K, class Bar = BarAll>
__device__ __forceinline__ void cgo_cta_combine(double (&acc)[K], double *sm /* K*CGO_NW */, Bar bar = Bar()) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        double v = acc[k];
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) v = v + __shfl_xor_sync(0xffffffffu, v, off);
        if (lane == 0) sm[k * CGO_NW + warp] = v;
    }
    bar();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
            double s = sm[k * CGO_NW];
#pragma unroll
            for (int w = 1; w < CGO_NW; ++w) s = s + sm[k * CGO_NW + w];
            acc[k] = s;
        }
    }
    bar();
}

// Thread 0 of virtual CTA `vcta` publishes its K partials.
template <int K>
__device__ __forceinline__ void cgo_publish(const RedArgs &red, int vcta, const double (&acc)[K]) {
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) __stcg(&red.partial[(size_t)k * red.G + vcta], acc[k]);
    }
}

// Called by every physical CTA once all its virtual CTAs are published.  The last CTA to arrive
// combines the `nact` partials in canonical order and writes red.out[0..K).
template <int K, class Bar = BarAll>
__device__ __forceinline__ void cgo_grid_finish(const RedArgs &red, int nact, double *sm, Bar bar = Bar()) {
    __shared__ bool is_last;
    if (red.sig0 != nullptr || red.flags_all != nullptr) __threadfence_system();   // this CTA's peer-memory stores
    else __threadfence();
    bar();
    if (threadIdx.x == 0) {
        unsigned int t = atomicAdd(red.ticket, 1u);
        is_last = (t == gridDim.x - 1);
    }
    bar();
    if (!is_last) return;
    __threadfence();
    double acc[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        double s = 0.0;
        for (int c = threadIdx.x; c < nact; c += CGO_B) s = s + __ldcg(&red.partial[(size_t)k * red.G + c]);
        acc[k] = s;
    }
    cgo_cta_combine<K, Bar>(acc, sm, bar);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) red.out[k] = acc[k];
        *red.ticket = 0u;
        __threadfence_system();
        if (red.sig0 != nullptr) {      // every CTA fenced its peer stores before its ticket
            cgo_st_release_sys(red.sig0, red.sig_val);
            if (red.sig1 != nullptr) cgo_st_release_sys(red.sig1, red.sig_val);
        }
        if (red.flags_all != nullptr)
            for (int r = 0; r < red.nranks; ++r)
                cgo_st_release_sys((unsigned long long *)red.flags_all[r] + red.sig_all_slot + red.me, red.sig_val);
    }
}

// spin until *flag >= val.  Ranks run in lockstep, so a wait lasts micro- to milliseconds; a peer that
// died (host exception, killed process) must not leave this GPU spinning for ever: after ≈30 s the
// kernel traps and the next CUDA call of this rank reports the failure.
__device__ __forceinline__ void cgo_spin_until(const unsigned long long *flag, unsigned long long val) {
    unsigned int polls = 0;
    while (cgo_ld_acquire_sys(flag) < val) {
        __nanosleep(128);
        if (++polls > (1u << 28)) asm volatile("trap;");
    }
}
// consumer side of the hand-off: thread 0 spins on the local flags, `bar` releases the others
template <class Bar>
__device__ __forceinline__ void cgo_wait_flags(const RedArgs &red, Bar bar) {
    if (red.wait0 == nullptr && red.wait_all == nullptr) return;
    if (threadIdx.x == 0 && red.wait0 != nullptr) {
        cgo_spin_until(red.wait0, red.wait_val);
        if (red.wait1 != nullptr) cgo_spin_until(red.wait1, red.wait_val);
    }
    if (red.wait_all != nullptr && (int)threadIdx.x < red.nranks) cgo_spin_until(red.wait_all + threadIdx.x, red.wait_val);
    bar();
}

// 128-bit streaming loads / stores (inputs are read once per kernel: keep them out of L1)
__device__ __forceinline__ double2 cgo_ld2(const double2 *p) {
    double2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ void cgo_st2(double2 *p, double2 v) {
    asm volatile("st.global.L1::no_allocate.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
}
