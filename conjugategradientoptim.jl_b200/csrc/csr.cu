// csr.cu — CSR objectives: sparse least squares ½‖Ax − b‖² and logistic regression.
//
// These are device-resident replacements of the user callback `fdf!(g, x) -> f` that the
// reference calls at src/engine/optim.jl:25 and src/cg_utils.jl:18 (the reference ships no
// large objective; the definitions are restated in oracle/cgo_oracle.c
// sparse_ls_fdf / logreg_fdf).  One trial  evalϕdϕ! (src/cg_utils.jl:3-22)  is three kernels:
//   K_a  xp = x + a u  [after u = −g + βu, cg_flavours.jl:10-12]           BLAS-1, blas1.cu
//   K_b  r  = A xp − b,  Σ r²                (logreg: margins, loss, c)     k_csr_rows
//   K_c  g⁺ = Aᵀ r  + every dot of getβ / norm(df_xp) / dϕ                  k_csr_rows on Aᵀ
// Aᵀ is an explicit CSR whose rows are sorted by source entry, so the gradient is a gather
// (no atomics on data) and reproduces the sequential scatter order of the oracle bit for bit.
//
// k_csr_rows is a "CSR-stream" kernel: a CTA owns 256 consecutive rows at a time; their
// nonzeros are one contiguous range of val/col, which a producer warp streams into a circular
// shared-memory buffer with 1-D TMA bulk copies (cp.async.bulk + mbarrier, L2 evict-first) while
// the 256 row lanes gather the vector entries and add up their row's products in storage
// order.  Rows and tiles of any length work.
#include <cub/device/device_scan.cuh>

#include "internal.cuh"
#include "reduce.cuh"


constexpr int CSR_PAD = 16;                 // slack entries behind every CSR array (see csr_alloc)

__device__ __forceinline__ double ld_stream_f64(const double *p) {
    double r;
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ int32_t ld_stream_s32(const int32_t *p) {
    int32_t r;
    asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream_f64(double *p, double v) {
    asm volatile("st.global.L1::no_allocate.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}

// ------------------------------------------------------------------ mbarrier / TMA (1-D bulk) PTX
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
// global -> shared bulk copy (TMA engine), completion counted in bytes on `bar`; 16-byte
// aligned addresses and size
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar, uint64_t pol) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
        ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol) : "memory");
}

// ------------------------------------------------------------------ the CSR row kernel
// Canonical reduction site: V = 1, U = 1 (item = row, lane t of a tile owns row t;
// include/cgoptim.h).  Warp-specialised: warp 8 is the producer (one lane drives the TMA
// engine), warps 0-7 are the 256 row lanes.  Everything that streams — matrix values, column
// indices, row pointers and the epilogue's per-row operands (b; u, g; w) — reaches shared memory
// only through bulk copies with an evict-first L2 policy; the lanes' only global loads are the
// gathers of the vector, issued with an evict-last policy so that the gathered window stays in
// L2 while the matrix streams through it.
//
// Shared-memory layout: a circular buffer of RING entries (values + column indices) that holds
// variable-length chunks back to back, and TS_ND chunk descriptors (full/empty mbarrier pair,
// the tile's row-pointer slice and operand slices).  A chunk is the entry range [p0 & ~3, p1)
// of one tile (tiles of Aᵀ vary in length), or a CHMAX piece of it when the tile is longer than
// half the ring.  The producer runs ahead as far as ring space and descriptors allow (about
// three tiles); it reclaims space in FIFO order as the lanes release chunks.
constexpr int TS_ND = 3;                    // chunk descriptors
constexpr int TS_THREADS = CGO_B + 32;
constexpr int TS_OCC = 2;
constexpr int TS_UN = 10;                   // gathers in flight per lane
constexpr int TS_SMEM_BUDGET = 115200;      // two CTAs per SM (228 KB - 2 x 1 KB reserved)
// OCC CTAs per SM: 2 for the streaming-bound kernels (a ring of ~3 tiles of 10 entries per row);
// 3 for the short rows of a column-block pass, which are gather-latency bound and want more warps
template <int NOPS, int OCC = TS_OCC>
struct TsLayout {
    static constexpr int BUDGET = OCC == 2 ? TS_SMEM_BUDGET : (228 * 1024) / OCC - 1024 - 256;
    static constexpr int RP_BYTES = (CGO_B + 2) * 8;
    static constexpr int OP_BYTES = CGO_B * 8;
    static constexpr int DESC_BYTES = RP_BYTES + NOPS * OP_BYTES;
    static constexpr int TAIL_BYTES = 2 * TS_ND * 8 + TS_ND * 4 + CGO_MAXK * CGO_NW * 8;
    static constexpr int RING = ((BUDGET - TS_ND * DESC_BYTES - TAIL_BYTES - 64) / 12) & ~3;
    static constexpr int CHMAX = (RING / 2) & ~3;
    // byte offsets (all multiples of 16)
    static constexpr int OFF_VAL = 0;
    static constexpr int OFF_COL = OFF_VAL + RING * 8;
    static constexpr int OFF_DESC = (OFF_COL + RING * 4 + 15) & ~15;
    static constexpr int OFF_BAR = OFF_DESC + TS_ND * DESC_BYTES;
    static constexpr int OFF_LEN = OFF_BAR + 2 * TS_ND * 8;
    static constexpr int OFF_RED = (OFF_LEN + TS_ND * 4 + 15) & ~15;
    static constexpr int BYTES = OFF_RED + CGO_MAXK * CGO_NW * 8;
    static_assert(BYTES <= BUDGET && RP_BYTES % 16 == 0 && DESC_BYTES % 16 == 0, "smem layout");
};

__device__ __forceinline__ bool ts_next_tile(int &v, int64_t &tile, int nact, int64_t ntiles, int G) {
    tile += G;
    if (tile < ntiles) return true;
    v += gridDim.x;
    if (v >= nact) return false;
    tile = v;
    return true;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ double ld_gather_f64(const double *p, uint64_t pol) {
    double r;
    asm volatile("ld.global.nc.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(r) : "l"(p), "l"(pol));
    return r;
}
// predicated gather: a lane whose row has ended issues no request at all.  (A gather instruction
// costs one L1TEX wavefront per distinct 128-byte line; re-reading a dummy address on the idle
// lanes — the first version — spent half of the gather bandwidth of the 5-entry rows of a
// column-block pass on nothing.)
__device__ __forceinline__ double ld_gather_f64_if(const double *p, uint64_t pol, bool on) {
    double r = 0.0;
    asm volatile(
        "{\n"
        ".reg .pred q;\n"
        "setp.ne.s32 q, %3, 0;\n"
        "@q ld.global.nc.L2::cache_hint.f64 %0, [%1], %2;\n"
        "}\n" : "+d"(r) : "l"(p), "l"(pol), "r"((int)on));
    return r;
}

// ---------------- producer: one lane streams tiles into the ring, about three tiles ahead
template <class L, class Epi>
__device__ __forceinline__ void ts_produce(const CsrMat &A, const Epi &epi, const RedArgs &red,
                                           double *s_val, int32_t *s_col, unsigned char *s_desc, uint64_t *s_full,
                                           uint64_t *s_empty, uint32_t *s_len, int nact, int64_t ntiles) {
    constexpr int NOPS = Epi::NOPS;
    constexpr int RING = L::RING, CHMAX = L::CHMAX;
    const int64_t nrows = A.nrows;
    const uint64_t pol = l2_policy_evict_first();
    uint32_t c = 0, tail = 0;         // chunks issued / reclaimed
    uint32_t head = 0, nfree = RING;
    int v = blockIdx.x;
    int64_t tile = v;
    bool have = true;
    int nvalid_n = (int)(nrows - tile * CGO_B < CGO_B ? nrows - tile * CGO_B : CGO_B);
    int64_t p0n = __ldg(A.rowptr + tile * CGO_B), p1n = __ldg(A.rowptr + tile * CGO_B + nvalid_n);
    while (have) {
        const int64_t r0 = tile * CGO_B, p0 = p0n, p1 = p1n;
        const int nvalid = nvalid_n;
        have = ts_next_tile(v, tile, nact, ntiles, red.G);
        if (have) {                   // row pointers of the next tile: in flight during this one
            nvalid_n = (int)(nrows - tile * CGO_B < CGO_B ? nrows - tile * CGO_B : CGO_B);
            p0n = __ldg(A.rowptr + tile * CGO_B);
            p1n = __ldg(A.rowptr + tile * CGO_B + nvalid_n);
        }
        bool first = true;
        for (int64_t cs = p0 & ~(int64_t)3; first || cs < p1; cs += CHMAX) {
            const int64_t ce = cs + CHMAX < p1 ? cs + CHMAX : p1;
            const uint32_t cnt4 = p1 > p0 ? (uint32_t)((ce - cs + 3) & ~(int64_t)3) : 0u;
            // a free descriptor and cnt4 free ring entries: reclaim released chunks, oldest first
            while (tail + TS_ND <= c || nfree < cnt4) {
                mbar_wait(&s_empty[tail % TS_ND], (tail / TS_ND) & 1);
                nfree += s_len[tail % TS_ND];
                ++tail;
            }
            const int d = c % TS_ND;
            s_len[d] = cnt4;
            nfree -= cnt4;
            unsigned char *desc = s_desc + d * L::DESC_BYTES;
            const uint32_t rpb = first ? (uint32_t)((((nvalid + 1) * 8) + 15) & ~15) : 0u;
            const uint32_t opb = first ? (uint32_t)(((nvalid * 8) + 15) & ~15) : 0u;
            mbar_expect_tx(&s_full[d], cnt4 * 12u + rpb + NOPS * opb);
            if (first) {
                tma_load_1d(desc, A.rowptr + r0, rpb, &s_full[d], pol);
#pragma unroll
                for (int o = 0; o < NOPS; ++o)
                    tma_load_1d(desc + L::RP_BYTES + o * L::OP_BYTES, epi.operand(o) + r0, opb, &s_full[d], pol);
            }
            if (cnt4) {
                const uint32_t n1 = cnt4 < RING - head ? cnt4 : RING - head;   // up to the ring's end
                tma_load_1d(s_val + head, A.val + cs, n1 * 8u, &s_full[d], pol);
                tma_load_1d(s_col + head, A.col + cs, n1 * 4u, &s_full[d], pol);
                if (n1 < cnt4) {                                               // wrapped remainder
                    tma_load_1d(s_val, A.val + cs + n1, (cnt4 - n1) * 8u, &s_full[d], pol);
                    tma_load_1d(s_col, A.col + cs + n1, (cnt4 - n1) * 4u, &s_full[d], pol);
                }
                head += cnt4;
                if (head >= RING) head -= RING;
            }
            first = false;
            ++c;
        }
    }
}

template <class Epi, int OCC>
__global__ void __launch_bounds__(TS_THREADS, OCC)
k_csr_rows(CsrMat A, const double *__restrict__ xg, Epi epi, RedArgs red) {
    constexpr int K = Epi::K;
    constexpr int NOPS = Epi::NOPS;
    using L = TsLayout<NOPS, OCC>;
    constexpr int RING = L::RING, CHMAX = L::CHMAX;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *s_val = reinterpret_cast<double *>(smem_raw + L::OFF_VAL);
    int32_t *s_col = reinterpret_cast<int32_t *>(smem_raw + L::OFF_COL);
    unsigned char *s_desc = smem_raw + L::OFF_DESC;
    uint64_t *s_full = reinterpret_cast<uint64_t *>(smem_raw + L::OFF_BAR);
    uint64_t *s_empty = s_full + TS_ND;
    uint32_t *s_len = reinterpret_cast<uint32_t *>(smem_raw + L::OFF_LEN);   // producer's notes
    double *s_red = reinterpret_cast<double *>(smem_raw + L::OFF_RED);
    const int tid = threadIdx.x;
    const int64_t nrows = A.nrows;
    const int64_t ntiles = (nrows + CGO_B - 1) / CGO_B;
    const int nact = (int)(ntiles < (int64_t)red.G ? ntiles : (int64_t)red.G);
    if (tid == 0) {
        for (int s = 0; s < TS_ND; ++s) { mbar_init(&s_full[s], 1); mbar_init(&s_empty[s], CGO_NW); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (tid >= CGO_B) {                       // ---------------- producer warp
        if (tid == CGO_B && (int)blockIdx.x < nact)
            ts_produce<L, Epi>(A, epi, red, s_val, s_col, s_desc, s_full, s_empty, s_len, nact, ntiles);
        return;
    }

    // ---------------- 256 row lanes
    const BarLanes bar;
    cgo_wait_flags(red, bar);                 // sharded: the neighbours' halo pushes have landed
    const uint64_t gpol = l2_policy_evict_last();
    uint32_t c = 0, head = 0;                 // chunks consumed, ring offset of the next chunk
    for (int v = blockIdx.x; v < nact; v += gridDim.x) {
        double acc[K];
#pragma unroll
        for (int k = 0; k < K; ++k) acc[k] = 0.0;
        for (int64_t tile = v; tile < ntiles; tile += red.G) {
            const int64_t r0 = tile * CGO_B;
            const int nvalid = (int)(nrows - r0 < CGO_B ? nrows - r0 : CGO_B);
            const bool valid = tid < nvalid;
            typename Epi::Pre pre;
            int64_t rs = 0, re = 0, p0 = 0, p1 = 0, cs = 0;
            double sum = 0.0;
            bool first = true;
            do {
                const int d = c % TS_ND;
                mbar_wait(&s_full[d], (c / TS_ND) & 1);
                if (first) {
                    const unsigned char *desc = s_desc + d * L::DESC_BYTES;
                    const int64_t *rp = reinterpret_cast<const int64_t *>(desc);
                    if (valid) {
                        rs = rp[tid]; re = rp[tid + 1];
                        pre = epi.load(reinterpret_cast<const double *>(desc + L::RP_BYTES), tid);
                        if (Epi::HAS_INIT) sum = epi.init(pre);     // column-blocked pass: continue the row sum
                    }
                    p0 = rp[0];
                    p1 = rp[nvalid];
                    cs = p0 & ~(int64_t)3;
                    first = false;
                }
                const int64_t ce = cs + CHMAX < p1 ? cs + CHMAX : p1;
                const uint32_t cnt4 = p1 > p0 ? (uint32_t)((ce - cs + 3) & ~(int64_t)3) : 0u;
                const int lo = (int)((rs > cs ? rs : cs) - cs);
                const int hi = (int)((re < ce ? re : ce) - cs);
                for (int k0 = lo; k0 < hi; k0 += TS_UN) {
                    // branch-free batch: TS_UN gathers are issued back to back (entries past the
                    // row's end are predicated off and not added)
                    double vj[TS_UN], xj[TS_UN];
#pragma unroll
                    for (int j = 0; j < TS_UN; ++j) {
                        const int kk = k0 + j < hi ? k0 + j : hi - 1;
                        uint32_t idx = head + (uint32_t)kk;
                        if (idx >= RING) idx -= RING;
                        vj[j] = s_val[idx];
                        xj[j] = ld_gather_f64_if(xg + s_col[idx], gpol, k0 + j < hi);
                    }
#pragma unroll
                    for (int j = 0; j < TS_UN; ++j) {
                        const double t = sum + vj[j] * xj[j];
                        sum = k0 + j < hi ? t : sum;
                    }
                }
                __syncwarp();
                if ((tid & 31) == 0) mbar_arrive(&s_empty[d]);
                head += cnt4;
                if (head >= RING) head -= RING;
                ++c;
                cs += CHMAX;
            } while (cs < p1);
            if (valid) epi.row(r0 + tid, sum, pre, acc);
        }
        cgo_cta_combine<K, BarLanes>(acc, s_red, bar);
        cgo_publish<K>(red, v, acc);
    }
    cgo_grid_finish<K, BarLanes>(red, nact, s_red, bar);
}

template <class Epi, int OCC>
static int launch_csr_occ(cgo_ctx *c, const CsrMat &A, const double *xg, const Epi &epi, const RedArgs &red, int tclass) {
    const int64_t ntiles = (A.nrows + CGO_B - 1) / CGO_B;
    int64_t nact = ntiles < red.G ? ntiles : red.G;
    int64_t phys = (int64_t)c->sms * OCC;
    int grid = (int)(nact < phys ? nact : phys);
    if (grid < 1) grid = 1;
    CGO_CHECK(!A.sliced, "internal: k_csr_rows on a matrix in the sliced layout");
    const size_t smem = TsLayout<Epi::NOPS, OCC>::BYTES;
    CGO_CUDA(cudaFuncSetAttribute(k_csr_rows<Epi, OCC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cgo_timer_begin(c, tclass);
    k_csr_rows<Epi, OCC><<<grid, TS_THREADS, smem, c->stream>>>(A, xg, epi, red);
    cgo_timer_end(c);
    c->launches++;
    CGO_CUDA(cudaGetLastError());
    return 0;
}
template <class Epi>
static int launch_csr(cgo_ctx *c, const CsrMat &A, const double *xg, const Epi &epi, const RedArgs &red, int tclass) {
    return launch_csr_occ<Epi, TS_OCC>(c, A, xg, epi, red, tclass);
}
// a column-block pass: short rows, gather-latency bound
template <class Epi>
static int launch_csr_pass(cgo_ctx *c, const CsrMat &A, const double *xg, const Epi &epi, const RedArgs &red, int tclass) {
    if (c->csr_pass_occ == 3) return launch_csr_occ<Epi, 3>(c, A, xg, epi, red, tclass);
    return launch_csr_occ<Epi, TS_OCC>(c, A, xg, epi, red, tclass);
}

// ------------------------------------------------------------------ row epilogues
// operand(o): per-row input vectors of the epilogue; the producer stages the tile's slice of
// each in shared memory next to the row pointers (they must be 16-byte aligned and padded by
// one element).  load() picks lane t's values out of that slice; row() consumes them once the
// row sum is known.
struct EpiStore {                       // y = A x
    static constexpr int K = 1, NOPS = 0;
    static constexpr bool HAS_INIT = false;
    struct Pre {};
    __device__ __forceinline__ double init(const Pre &) const { return 0.0; }
    double *y;
    __device__ __forceinline__ const double *operand(int) const { return nullptr; }
    __device__ __forceinline__ Pre load(const double *, int) const { return Pre(); }
    __device__ __forceinline__ void row(int64_t i, double sum, const Pre &, double (&)[K]) const { y[i] = sum; }
};
// r = A xp − b ; Σ r².  PUSH: the first / last `halo` residuals also go straight into the ring
// neighbours' halos of r (peer memory), where their SpMVᵀ gathers them.
template <bool PUSH>
struct EpiResidualT {
    static constexpr int K = 1, NOPS = 1;
    static constexpr bool HAS_INIT = false;
    struct Pre { double b; };
    __device__ __forceinline__ double init(const Pre &) const { return 0.0; }
    const double *b;
    double *r;
    double *prev_right, *next_left;      // PUSH
    int64_t halo, nrows;                 // PUSH
    __device__ __forceinline__ const double *operand(int) const { return b; }
    __device__ __forceinline__ Pre load(const double *ops, int t) const { return Pre{ops[t]}; }
    __device__ __forceinline__ void row(int64_t i, double sum, const Pre &p, double (&acc)[K]) const {
        const double rr = sum - p.b;
        r[i] = rr;                      // re-read (gathered) by K_c: keep it cacheable
        if (PUSH) {
            if (i < halo) prev_right[i] = rr;
            if (i >= nrows - halo) next_left[i - (nrows - halo)] = rr;
        }
        acc[0] = acc[0] + rr * rr;
    }
};
// g⁺ = Aᵀ r [· 1/N + λ w], fused with norm(df_xp)² (optim.jl:107), dϕ = g⁺·u (cg_utils.jl:20) and
// the getβ dots (cg_flavours.jl:63-76, 96-105, 140-145, 164-167); fills pack slots 1..8.
template <bool LOGREG>
struct EpiGrad {
    static constexpr int K = 8, NOPS = LOGREG ? 3 : 2;
    static constexpr bool HAS_INIT = false;
    struct Pre { double u, g, w; };
    __device__ __forceinline__ double init(const Pre &) const { return 0.0; }
    double *gp;
    const double *g, *u, *w;
    double invN, lambda;
    __device__ __forceinline__ const double *operand(int o) const { return o == 0 ? u : (o == 1 ? g : w); }
    __device__ __forceinline__ Pre load(const double *ops, int t) const {
        Pre p;
        p.u = ops[t]; p.g = ops[CGO_B + t];
        p.w = LOGREG ? ops[2 * CGO_B + t] : 0.0;
        return p;
    }
    __device__ __forceinline__ void row(int64_t j, double sum, const Pre &p, double (&acc)[K]) const {
        double gn = sum;
        if (LOGREG) gn = sum * invN + lambda * p.w;
        st_stream_f64(gp + j, gn);
        const double uu = p.u, gg = p.g;
        const double y = gn - gg;
        acc[CGO_P_DPHI - 1] = acc[CGO_P_DPHI - 1] + gn * uu;
        acc[CGO_P_GPGP - 1] = acc[CGO_P_GPGP - 1] + gn * gn;
        acc[CGO_P_YY - 1] = acc[CGO_P_YY - 1] + y * y;
        acc[CGO_P_UY - 1] = acc[CGO_P_UY - 1] + uu * y;
        acc[CGO_P_YGP - 1] = acc[CGO_P_YGP - 1] + y * gn;
        acc[CGO_P_GPG - 1] = acc[CGO_P_GPG - 1] + gn * gg;
        acc[CGO_P_UG - 1] = acc[CGO_P_UG - 1] + uu * gg;
        acc[CGO_P_UU - 1] = acc[CGO_P_UU - 1] + uu * uu;
    }
};
// logistic loss of sample i with margin z = a_i·w (oracle logreg_fdf):
//   t = −y z; e = exp(−|t|); ℓ = max(t,0) + log1p(e); σ = t ≥ 0 ? 1/(1+e) : e/(1+e); c = −y σ
struct EpiLogit {
    static constexpr int K = 1, NOPS = 1;
    static constexpr bool HAS_INIT = false;
    struct Pre { double y; };
    __device__ __forceinline__ double init(const Pre &) const { return 0.0; }
    const double *label;
    double *c;
    __device__ __forceinline__ const double *operand(int) const { return label; }
    __device__ __forceinline__ Pre load(const double *ops, int t) const { return Pre{ops[t]}; }
    __device__ __forceinline__ void row(int64_t i, double z, const Pre &p, double (&acc)[K]) const {
        const double y = p.y;
        const double t = -y * z;
        const double e = exp(-fabs(t));
        const double l = (t > 0.0 ? t : 0.0) + log1p(e);
        const double sg = t >= 0.0 ? 1.0 / (1.0 + e) : e / (1.0 + e);
        c[i] = -y * sg;
        acc[0] = acc[0] + l;
    }
};

// Hessian-vector product of least squares, ∇²f u = Aᵀ(A u): v = A u with Σ v² (= u·Hu), then
// hv = Aᵀ v with u·hv and hv·hv
struct EpiStoreSq {
    static constexpr int K = 1, NOPS = 0;
    static constexpr bool HAS_INIT = false;
    struct Pre {};
    __device__ __forceinline__ double init(const Pre &) const { return 0.0; }
    double *y;
    __device__ __forceinline__ const double *operand(int) const { return nullptr; }
    __device__ __forceinline__ Pre load(const double *, int) const { return Pre(); }
    __device__ __forceinline__ void row(int64_t i, double sum, const Pre &, double (&acc)[K]) const {
        y[i] = sum;
        acc[0] = acc[0] + sum * sum;
    }
};
struct EpiHv {
    static constexpr int K = 2, NOPS = 1;
    static constexpr bool HAS_INIT = false;
    struct Pre { double u; };
    __device__ __forceinline__ double init(const Pre &) const { return 0.0; }
    double *hv;
    const double *u;
    __device__ __forceinline__ const double *operand(int) const { return u; }
    __device__ __forceinline__ Pre load(const double *ops, int t) const { return Pre{ops[t]}; }
    __device__ __forceinline__ void row(int64_t j, double sum, const Pre &p, double (&acc)[K]) const {
        hv[j] = sum;
        acc[0] = acc[0] + p.u * sum;
        acc[1] = acc[1] + sum * sum;
    }
};

// quadratic-aware line search: v = A u with r·v, v·v and r·r (r = residual at x)
struct EpiQuadV {
    static constexpr int K = 3, NOPS = 1;
    static constexpr bool HAS_INIT = false;
    struct Pre { double r; };
    __device__ __forceinline__ double init(const Pre &) const { return 0.0; }
    double *v;
    const double *r;
    __device__ __forceinline__ const double *operand(int) const { return r; }
    __device__ __forceinline__ Pre load(const double *ops, int t) const { return Pre{ops[t]}; }
    __device__ __forceinline__ void row(int64_t i, double sum, const Pre &p, double (&acc)[K]) const {
        v[i] = sum;
        acc[0] = acc[0] + p.r * sum;
        acc[1] = acc[1] + sum * sum;
        acc[2] = acc[2] + p.r * p.r;
    }
};

// Column-blocked matrices (CsrBlocked below): pass j > 0 of a row continues the sum pass j − 1
// left in `partial` (one more staged operand), so the products of a row are still added one by one
// in storage order — bit-identical to the single-pass kernel.
template <class Base>
struct WithInit : Base {
    static constexpr int K = Base::K, NOPS = Base::NOPS + 1;
    static constexpr bool HAS_INIT = true;
    struct Pre { typename Base::Pre base; double init; };
    const double *partial;
    __device__ __forceinline__ const double *operand(int o) const { return o < Base::NOPS ? Base::operand(o) : partial; }
    __device__ __forceinline__ Pre load(const double *ops, int t) const {
        Pre p;
        p.base = Base::load(ops, t);
        p.init = ops[Base::NOPS * CGO_B + t];
        return p;
    }
    __device__ __forceinline__ double init(const Pre &p) const { return p.init; }
    __device__ __forceinline__ void row(int64_t i, double sum, const Pre &p, double (&acc)[K]) const {
        Base::row(i, sum, p.base, acc);
    }
};
template <class Base>
static WithInit<Base> with_init(const Base &b, const double *partial) {
    WithInit<Base> w;
    static_cast<Base &>(w) = b;
    w.partial = partial;
    return w;
}

// ------------------------------------------------------------------ the gather-bound SpMV kernel
// Matrices whose gathers do not coalesce — every lane of a warp-level gather touches its own 128-byte line: the
// per-row-random cfg-3 matrix (coh_log2 < 4), the column-block passes of logistic regression — are bound by
// L1TEX, one line per clock and SM: ≈ 245 G gathers/s on a B200 whatever the load path (scratch/gather_bench.cu:
// LDG, LDG with L2 hints, TEX, LDGSTS, TMA gather4 all measured).  What reaches that ceiling is many warps with
// a short instruction stream, each keeping a batch of gathers in flight: the prototype in gather_bench.cu gets
// 244 G gathers/s with 32 warps per SM, against ≈ 105 G/s for k_csr_rows, whose canonical reduction order caps
// the grid at G = 296 CTAs (16 warps per SM) and whose ring, barriers and epilogue leave those warps without a
// gather in flight half of the time.  So for these matrices the reductions leave the SpMV:
//   k_spmv_direct  y = A x [− b | /N + λw | + partial]: no reduction ⇒ no canonical order to respect ⇒ any
//                  grid, any schedule; 24-32 warps per SM, values and column indices read straight from HBM by
//                  coalesced loads (sliced layout: the k-th entries of a slice's 32 rows are consecutive), no
//                  shared memory, slices handed out dynamically — which also keeps all warps inside one band
//                  of the gathered vector, so it stays L2-resident (the static tile order of k_csr_rows lets its
//                  CTAs drift apart by more rows than the band is wide on a 2e8-row matrix: then every gather is
//                  a DRAM sector read, 86 GB of DRAM traffic against 30 GB algorithmic, 36 ms instead of 15);
//   k_blas1        the dots, in the BLAS-1 canonical order (V = 2, U = 4): Σ r² after K_b, the eight getβ dots
//                  after K_c — 16n more bytes per evaluation than the fused epilogues (+5 %).
// Each row's products are still added one by one in storage order: row sums are bit-identical to k_csr_rows'.
constexpr int DIR_GRAB = 4;                 // slices (of 32 rows) a warp takes from the queue at a time
struct DirArgs {
    unsigned long long *next;               // slice queue head (reset by the last CTA)
    unsigned int *ticket;
    // cross-GPU hand-off, as in RedArgs
    unsigned long long *sig0, *sig1;
    const unsigned long long *wait0, *wait1;
    unsigned long long sig_val, wait_val;
    void *const *flags_all;
    const unsigned long long *wait_all;
    int sig_all_slot, nranks, me;
};
__device__ __forceinline__ int32_t ld_stream_s32_if(const int32_t *p, bool on) {
    int32_t r = 0;
    asm volatile("{\n.reg .pred q;\nsetp.ne.s32 q, %2, 0;\n@q ld.global.nc.L1::no_allocate.s32 %0, [%1];\n}\n" : "+r"(r) : "l"(p), "r"((int)on));
    return r;
}
__device__ __forceinline__ double ld_stream_f64_if(const double *p, bool on) {
    double r = 0.0;
    asm volatile("{\n.reg .pred q;\nsetp.ne.s32 q, %2, 0;\n@q ld.global.nc.L1::no_allocate.f64 %0, [%1];\n}\n" : "+d"(r) : "l"(p), "r"((int)on));
    return r;
}

// one batch of UN levels of a slice: the lane's entries, gathered and added in order
template <int UN, bool UNIFORM>
__device__ __forceinline__ double dir_batch(const CsrMat &A, const double *__restrict__ xg, uint64_t gpol, int64_t sb,
                                            int64_t &off, int len, int k0, int lane, uint32_t lt, double sum) {
    int32_t cj[UN];
    double vj[UN], xj[UN];
#pragma unroll
    for (int j = 0; j < UN; ++j) {
        const bool on = len > k0 + j;
        int64_t idx;
        if (UNIFORM) { idx = sb + off + lane; off += 32; }
        else {
            const uint32_t m = __ballot_sync(0xffffffffu, on);
            idx = sb + off + __popc(m & lt);
            off += __popc(m);
        }
        cj[j] = ld_stream_s32_if(A.col + idx, on);
        vj[j] = ld_stream_f64_if(A.val + idx, on);
    }
#pragma unroll
    for (int j = 0; j < UN; ++j) xj[j] = ld_gather_f64_if(xg + cj[j], gpol, len > k0 + j);
#pragma unroll
    for (int j = 0; j < UN; ++j) {
        const double p = vj[j] * xj[j];
        if (len > k0 + j) sum = sum + p;
    }
    return sum;
}

template <class F, int UN, int OCC>
__global__ void __launch_bounds__(CGO_B, OCC)
k_spmv_direct(CsrMat A, const double *__restrict__ xg, F f, DirArgs da) {
    const int tid = threadIdx.x, lane = tid & 31;
    const uint32_t lt = (1u << lane) - 1u;
    const int64_t nrows = A.nrows, nslices = (nrows + 31) / 32;
    if (da.wait0 != nullptr || da.wait_all != nullptr) {        // sharded: the peers' pushes have landed
        if (tid == 0 && da.wait0 != nullptr) {
            cgo_spin_until(da.wait0, da.wait_val);
            if (da.wait1 != nullptr) cgo_spin_until(da.wait1, da.wait_val);
        }
        if (da.wait_all != nullptr && tid < da.nranks) cgo_spin_until(da.wait_all + tid, da.wait_val);
        __syncthreads();
    }
    const uint64_t gpol = l2_policy_evict_last();
    for (;;) {
        unsigned long long s0 = 0;
        if (lane == 0) s0 = atomicAdd(da.next, (unsigned long long)DIR_GRAB);
        s0 = __shfl_sync(0xffffffffu, s0, 0);
        if ((int64_t)s0 >= nslices) break;
#pragma unroll 1
        for (int g = 0; g < DIR_GRAB; ++g) {
            const int64_t s = (int64_t)s0 + g;
            if (s >= nslices) break;
            const int64_t row = s * 32 + lane;
            const bool valid = row < nrows;
            int64_t rp0 = 0, rp1 = 0;
            if (valid) { rp0 = __ldg(A.rowptr + row); rp1 = __ldg(A.rowptr + row + 1); }
            const int len = valid ? (int)(rp1 - rp0) : -1;
            const int64_t sb = __shfl_sync(0xffffffffu, rp0, 0);
            const int maxlen = __reduce_max_sync(0xffffffffu, len);
            const bool uniform = __all_sync(0xffffffffu, len == maxlen);
            double sum = valid ? f.init(row) : 0.0;
            int64_t off = 0;
            if (uniform) { for (int k0 = 0; k0 < maxlen; k0 += UN) sum = dir_batch<UN, true>(A, xg, gpol, sb, off, len, k0, lane, lt, sum); }
            else { for (int k0 = 0; k0 < maxlen; k0 += UN) sum = dir_batch<UN, false>(A, xg, gpol, sb, off, len, k0, lane, lt, sum); }
            if (valid) f.row(row, sum);
        }
    }
    // the last CTA re-arms the queue and hands the pushed halos to the peers
    __shared__ bool is_last;
    if (da.sig0 != nullptr || da.flags_all != nullptr) __threadfence_system();
    else __threadfence();
    __syncthreads();
    if (tid == 0) is_last = atomicAdd(da.ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (is_last && tid == 0) {
        *da.ticket = 0u;
        *da.next = 0ULL;
        __threadfence_system();
        if (da.sig0 != nullptr) {
            cgo_st_release_sys(da.sig0, da.sig_val);
            if (da.sig1 != nullptr) cgo_st_release_sys(da.sig1, da.sig_val);
        }
        if (da.flags_all != nullptr)
            for (int r = 0; r < da.nranks; ++r)
                cgo_st_release_sys((unsigned long long *)da.flags_all[r] + da.sig_all_slot + da.me, da.sig_val);
    }
}

// row functors of k_spmv_direct: init(i) = what the row sum starts from, row(i, sum) = what to do with it
struct DirStore {                       // y = A x
    double *y;
    __device__ __forceinline__ double init(int64_t) const { return 0.0; }
    __device__ __forceinline__ void row(int64_t i, double sum) const { y[i] = sum; }
};
struct DirStoreInit {                   // column-block pass: y = partial + A_j x
    double *y;
    const double *partial;
    __device__ __forceinline__ double init(int64_t i) const { return ld_stream_f64(partial + i); }
    __device__ __forceinline__ void row(int64_t i, double sum) const { y[i] = sum; }
};
template <bool PUSH>
struct DirResidual {                    // r = A xp − b  (+ halo push, as EpiResidualT)
    const double *b;
    double *r;
    double *prev_right, *next_left;
    int64_t halo, nrows;
    __device__ __forceinline__ double init(int64_t) const { return 0.0; }
    __device__ __forceinline__ void row(int64_t i, double sum) const {
        const double rr = sum - ld_stream_f64(b + i);
        r[i] = rr;
        if (PUSH) {
            if (i < halo) prev_right[i] = rr;
            if (i >= nrows - halo) next_left[i - (nrows - halo)] = rr;
        }
    }
};
struct DirGradLS {                      // g⁺ = Aᵀ r
    double *gp;
    __device__ __forceinline__ double init(int64_t) const { return 0.0; }
    __device__ __forceinline__ void row(int64_t j, double sum) const { st_stream_f64(gp + j, sum); }
};

static DirArgs dir_args(cgo_ctx *c, const RedArgs *red = nullptr) {
    DirArgs d;
    d.next = c->d_progress + 4;
    d.ticket = c->d_ticket + 1;
    d.sig0 = d.sig1 = nullptr; d.wait0 = d.wait1 = nullptr; d.sig_val = d.wait_val = 0;
    d.flags_all = nullptr; d.wait_all = nullptr; d.sig_all_slot = -1; d.nranks = 1; d.me = 0;
    if (red) {
        d.sig0 = red->sig0; d.sig1 = red->sig1; d.wait0 = red->wait0; d.wait1 = red->wait1;
        d.sig_val = red->sig_val; d.wait_val = red->wait_val;
        d.flags_all = red->flags_all; d.wait_all = red->wait_all; d.sig_all_slot = red->sig_all_slot;
        d.nranks = red->nranks; d.me = red->me;
    }
    return d;
}
template <class F, int UN, int OCC>
static int launch_direct_k(cgo_ctx *c, const CsrMat &A, const double *xg, const F &f, const DirArgs &da, int tclass) {
    const int64_t nslices = (A.nrows + 31) / 32;
    int64_t want = (nslices + 8 * DIR_GRAB - 1) / (8 * DIR_GRAB), phys = (int64_t)c->sms * OCC;
    int grid = (int)(want < phys ? want : phys);
    if (grid < 1) grid = 1;
    cgo_timer_begin(c, tclass);
    k_spmv_direct<F, UN, OCC><<<grid, CGO_B, 0, c->stream>>>(A, xg, f, da);
    cgo_timer_end(c);
    c->launches++;
    CGO_CUDA(cudaGetLastError());
    return 0;
}
template <class F>
static int launch_direct(cgo_ctx *c, const CsrMat &A, const double *xg, const F &f, const DirArgs &da, int tclass) {
    CGO_CHECK(A.sliced, "internal: k_spmv_direct needs the sliced layout");
    // gathers per batch × CTAs per SM, measured at n = 2e8, coh_log2 = 0 (K_b = ten entries in every row / K_c =
    // ragged rows of Aᵀ, ms; profiles/r2_exp_coh0_n2e8_k_spmv_direct_configs.txt): 10 × 4: 7.5–7.9 / 9.3–10.5;
    // 8 × 4: 8.0 / 9.1; 6 × 5: 7.7–8.1 / 9.1; 5 × 4: 7.7 / 9.3; 8 × 5 (spills): 8.4 / 9.5; 10 × 3: 7.9 / 10.5.
    // Ragged slices waste the tail of a wide batch, equal-length rows like the widest batch that does not spill.
    const int cfg = c->direct_cfg ? c->direct_cfg : (A.ragged ? 3 : 2);
    if (cfg == 1) return launch_direct_k<F, 8, 4>(c, A, xg, f, da, tclass);
    if (cfg == 3) return launch_direct_k<F, 6, 5>(c, A, xg, f, da, tclass);
    return launch_direct_k<F, 10, 4>(c, A, xg, f, da, tclass);
}

// ------------------------------------------------------------------ counter-based hash
// (generator spec: oracle/cgo_oracle.c hash3 / u01; restated, identical arithmetic)
__host__ __device__ __forceinline__ uint64_t g_mix64(uint64_t z) {
    z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ULL;
    z ^= z >> 27; z *= 0x94D049BB133111EBULL;
    z ^= z >> 31;
    return z;
}
__host__ __device__ __forceinline__ uint64_t g_hash3(uint64_t seed, uint64_t i, uint64_t k) {
    uint64_t h = g_mix64(seed + 0x9E3779B97F4A7C15ULL);
    h = g_mix64(h ^ (i + 0x9E3779B97F4A7C15ULL));
    h = g_mix64(h ^ (k + 0x632BE59BD9B4E019ULL));
    return h;
}
__host__ __device__ __forceinline__ double g_u01(uint64_t seed, uint64_t i, uint64_t k) {
    return (double)(g_hash3(seed, i, k) >> 11) * (1.0 / 9007199254740992.0);
}

// banded-random least-squares generator (spec text: oracle/cgo_oracle.c orc_obj_sparse_ls_synth)
struct LsSpec {
    int64_t n;          // global dimension
    int32_t K;          // entries per row
    int32_t coh;        // log2 of the number of consecutive rows sharing their offsets
    int64_t W, w;       // half band width, stratum width 2W/(K−1)
    uint64_t seed;
    // signed column offset and value of entry k of global row i
    __host__ __device__ __forceinline__ void entry(int64_t i, int k, int64_t &d, double &v) const {
        if (k == 0) { d = 0; v = 4.0 + g_u01(seed, (uint64_t)i, 0); return; }
        const int64_t lo = -W + (int64_t)(k - 1) * w;
        d = lo + (int64_t)(g_hash3(seed ^ 0xA5A5A5A5A5A5A5A5ULL, (uint64_t)(i >> coh), (uint64_t)k) % (uint64_t)w);
        if (d == 0) d = (lo + w > 1) ? 1 : -1;
        v = 0.3 * (2.0 * g_u01(seed, (uint64_t)i, (uint64_t)k) - 1.0);
    }
    __host__ __device__ __forceinline__ int64_t wrap(int64_t c) const {
        if (c < 0) c += n;
        if (c >= n) c -= n;
        return c;
    }
};
// logistic-regression generator (spec text: oracle/cgo_oracle.c orc_obj_logreg_synth)
struct LrSpec {
    int64_t N, d;
    int32_t K;
    int64_t w;          // stratum width d / K
    uint64_t seed;
    __host__ __device__ __forceinline__ void entry(int64_t i, int k, int64_t &c, double &v) const {
        c = (int64_t)k * w + (int64_t)(g_hash3(seed, (uint64_t)i, (uint64_t)k) % (uint64_t)w);
        v = 2.0 * g_u01(seed + 7, (uint64_t)i, (uint64_t)k) - 1.0;
    }
};

static inline int grid_for(int64_t items, int sms) {
    int64_t b = (items + 255) / 256;
    int64_t cap = (int64_t)sms * 32;
    return (int)(b < 1 ? 1 : (b < cap ? b : cap));
}
#define GRID_STRIDE(idx, total) \
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < (total); idx += (int64_t)gridDim.x * blockDim.x)

// rows [lo, lo + nloc) of A.  unwrapped: column index relative to lo, in [−W, nloc + W) (halo
// layout, multi-rank); else the wrapped global column (single rank).
__global__ void k_ls_fill_A(LsSpec s, int64_t lo, int64_t nloc, bool unwrapped, CsrMat A) {
    const int64_t total = nloc * s.K;
    GRID_STRIDE(p, total) {
        const int64_t il = p / s.K;
        const int k = (int)(p - il * s.K);
        int64_t d; double v;
        s.entry(lo + il, k, d, v);
        A.col[p] = (int32_t)(unwrapped ? il + d : s.wrap(lo + il + d));
        A.val[p] = v;
        if (k == 0) A.rowptr[il] = p;
        if (p == total - 1) A.rowptr[nloc] = total;
    }
}
// x_true over [lo − halo, lo + nloc + halo), wrapped
__global__ void k_ls_xtrue(LsSpec s, int64_t lo, int64_t nloc, int64_t halo, double *xt /* local origin */) {
    GRID_STRIDE(e, nloc + 2 * halo) {
        int64_t i = s.wrap(lo - halo + e);
        xt[e - halo] = 2.0 * g_u01(s.seed + 1, (uint64_t)i, 0) - 1.0;
    }
}
// samples [i0, i0 + nloc) of the generator; column = global feature index
__global__ void k_lr_fill_A(LrSpec s, int64_t i0, int64_t nloc, CsrMat A, double *label) {
    GRID_STRIDE(il, nloc) {
        const int64_t i = i0 + il;
        const int64_t p = il * s.K;
        A.rowptr[il] = p;
        double acc = 0.0;
        for (int k = 0; k < s.K; ++k) {
            int64_t c; double v;
            s.entry(i, k, c, v);
            A.col[p + k] = (int32_t)c;
            A.val[p + k] = v;
            const double wt = 2.0 * g_u01(s.seed + 2, (uint64_t)c, 0) - 1.0;
            acc += v * wt;
        }
        const double noise = 2.0 * g_u01(s.seed + 3, (uint64_t)i, 0) - 1.0;
        label[il] = (acc + 0.1 * noise >= 0.0) ? 1.0 : -1.0;
        if (il == nloc - 1) A.rowptr[nloc] = nloc * s.K;
    }
}

// ------------------------------------------------------------------ explicit transpose (setup)
// Entry sources enumerate (target row of Aᵀ, sort key) pairs; keys order the entries of one Aᵀ
// row exactly like the oracle's stable counting sort (ascending source entry).
struct CsrSrc {                 // a materialised CSR on this device
    const int32_t *col;
    int64_t total;
    __device__ __forceinline__ bool get(int64_t e, int64_t &trow, int64_t &key) const {
        trow = col[e]; key = e;
        return true;
    }
};
struct LsExtSrc {               // generator rows [lo − W, hi + W): entries landing in columns [lo, hi)
    LsSpec s;
    int64_t lo, hi, total;
    __device__ __forceinline__ bool get(int64_t e, int64_t &trow, int64_t &key) const {
        const int64_t ie = e / s.K;
        const int k = (int)(e - ie * s.K);
        const int64_t i = s.wrap(lo - s.W + ie);
        int64_t d; double v;
        s.entry(i, k, d, v);
        const int64_t c = s.wrap(i + d);
        if (c < lo || c >= hi) return false;
        trow = c - lo; key = i * s.K + k;
        return true;
    }
};
template <class Src>
__global__ void k_tr_count(Src src, unsigned long long *counts) {
    GRID_STRIDE(e, src.total) {
        int64_t trow, key;
        if (src.get(e, trow, key)) atomicAdd(counts + trow, 1ULL);
    }
}
template <class Src>
__global__ void k_tr_fill(Src src, unsigned long long *cursor, int64_t *perm) {
    GRID_STRIDE(e, src.total) {
        int64_t trow, key;
        if (src.get(e, trow, key)) perm[atomicAdd(cursor + trow, 1ULL)] = key;
    }
}
// each Aᵀ row's keys ascending (insertion sort for short rows, heapsort beyond)
__global__ void k_tr_sort(const int64_t *rowptrT, int64_t nT, int64_t *perm) {
    GRID_STRIDE(j, nT) {
        int64_t *a = perm + rowptrT[j];
        const int64_t L = rowptrT[j + 1] - rowptrT[j];
        if (L <= 48) {
            for (int64_t i = 1; i < L; ++i) {
                const int64_t key = a[i];
                int64_t q = i - 1;
                while (q >= 0 && a[q] > key) { a[q + 1] = a[q]; --q; }
                a[q + 1] = key;
            }
        } else {
            for (int64_t start = L / 2 - 1; start >= 0; --start) {       // heapify
                int64_t root = start;
                for (;;) {
                    int64_t ch = 2 * root + 1;
                    if (ch >= L) break;
                    if (ch + 1 < L && a[ch] < a[ch + 1]) ++ch;
                    if (a[root] >= a[ch]) break;
                    int64_t t = a[root]; a[root] = a[ch]; a[ch] = t;
                    root = ch;
                }
            }
            for (int64_t end = L - 1; end > 0; --end) {
                int64_t t = a[0]; a[0] = a[end]; a[end] = t;
                int64_t root = 0;
                for (;;) {
                    int64_t ch = 2 * root + 1;
                    if (ch >= end) break;
                    if (ch + 1 < end && a[ch] < a[ch + 1]) ++ch;
                    if (a[root] >= a[ch]) break;
                    int64_t t2 = a[root]; a[root] = a[ch]; a[ch] = t2;
                    root = ch;
                }
            }
        }
    }
}
struct CsrFin {                 // key = entry index of a materialised CSR
    const int64_t *rowptr;
    int64_t nrows;
    const double *val;
    __device__ __forceinline__ void get(int64_t key, int32_t &srow, double &v) const {
        int64_t lo = 0, hi = nrows;              // last row with rowptr[row] <= key
        while (hi - lo > 1) {
            int64_t mid = (lo + hi) >> 1;
            if (rowptr[mid] <= key) lo = mid; else hi = mid;
        }
        srow = (int32_t)lo; v = val[key];
    }
};
struct FixedKFin {              // materialised CSR with exactly K entries per row
    int32_t K;
    const double *val;
    __device__ __forceinline__ void get(int64_t key, int32_t &srow, double &v) const {
        srow = (int32_t)(key / K); v = val[key];
    }
};
struct LsExtFin {               // key = global entry index of the generator
    LsSpec s;
    int64_t lo;
    __device__ __forceinline__ void get(int64_t key, int32_t &srow, double &v) const {
        const int64_t i = key / s.K;
        const int k = (int)(key - i * s.K);
        int64_t d;
        s.entry(i, k, d, v);
        int64_t rel = i - lo;                    // unwrapped position relative to this shard
        if (rel < -s.W) rel += s.n;
        if (rel >= s.n - s.W) rel -= s.n;
        srow = (int32_t)rel;
    }
};
template <class Fin>
__global__ void k_tr_finalize(Fin fin, int64_t nnzT, const int64_t *perm, CsrMat AT) {
    GRID_STRIDE(q, nnzT) {
        int32_t srow; double v;
        fin.get(perm[q], srow, v);
        AT.col[q] = srow;
        AT.val[q] = v;
    }
}

static void csr_free(CsrMat &M) {
    cudaFree(M.rowptr); cudaFree(M.col); cudaFree(M.val);
    M = CsrMat();
}
static int csr_alloc(CsrMat &M, int64_t nrows, int64_t nnz) {
    M.nrows = nrows; M.nnz = nnz;
    // CSR_PAD: bulk copies are rounded up to 16 bytes and may read past the last entry
    CGO_CUDA(cudaMalloc(&M.rowptr, sizeof(int64_t) * (size_t)(nrows + 1 + CSR_PAD)));
    CGO_CUDA(cudaMalloc(&M.col, sizeof(int32_t) * (size_t)(nnz + CSR_PAD)));
    CGO_CUDA(cudaMalloc(&M.val, sizeof(double) * (size_t)(nnz + CSR_PAD)));
    return 0;
}

// ------------------------------------------------------------------ sliced layout (setup)
// Re-order the entries of every 32-row slice level-major (CsrMat::sliced; k_csr_rows ts_issue).  Row pointers
// stay CSR row pointers: a slice still owns the range [rowptr[32s], rowptr[32s+32)), only the order inside it
// changes, and each row's own entries keep their relative order.
// one warp per slice; inverse = back to row-major; *n_ragged (optional) += slices whose rows differ in length
__global__ void k_slice_permute(CsrMat M, const int32_t *col_in, const double *val_in, int32_t *col_out, double *val_out,
                                bool inverse, unsigned long long *n_ragged = nullptr) {
    const int lane = threadIdx.x & 31;
    const uint32_t lt = (1u << lane) - 1u;
    const int64_t nslices = (M.nrows + 31) / 32;
    const int64_t wstride = (int64_t)gridDim.x * (blockDim.x / 32);
    for (int64_t s = (int64_t)blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5); s < nslices; s += wstride) {
        const int64_t row = s * 32 + lane;
        int64_t rs = 0, len = 0;
        if (row < M.nrows) { rs = M.rowptr[row]; len = M.rowptr[row + 1] - rs; }
        const int64_t sb = M.rowptr[s * 32];
        if (n_ragged != nullptr) {
            const int64_t len0 = __shfl_sync(0xffffffffu, len, 0);
            const bool same = __all_sync(0xffffffffu, row >= M.nrows || len == len0);
            if (!same && lane == 0) atomicAdd(n_ragged, 1ULL);
        }
        int64_t off = 0;
        for (int64_t k = 0;; ++k) {
            const uint32_t m = __ballot_sync(0xffffffffu, len > k);
            if (m == 0) break;
            if (len > k) {
                const int64_t lvl = sb + off + __popc(m & lt), rm = rs + k;
                const int64_t src = inverse ? lvl : rm, dst = inverse ? rm : lvl;
                col_out[dst] = col_in[src];
                val_out[dst] = val_in[src];
            }
            off += __popc(m);
        }
    }
}
// a 1-in-64 sample of the 32-row slices of a row-major matrix: stats[0] += warp-level gathers (the k-th entries
// of a slice's rows), stats[1] += distinct 128-byte lines of the gathered vector they touch
__global__ void k_gather_lines(CsrMat M, unsigned long long *stats) {
    const int lane = threadIdx.x & 31;
    const uint32_t lt = (1u << lane) - 1u;
    const int64_t nslices = (M.nrows + 31) / 32;
    const int64_t wstride = (int64_t)gridDim.x * (blockDim.x / 32);
    for (int64_t s = ((int64_t)blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5)) * 64; s < nslices; s += wstride * 64) {
        const int64_t row = s * 32 + lane;
        int64_t rs = 0, len = 0;
        if (row < M.nrows) { rs = M.rowptr[row]; len = M.rowptr[row + 1] - rs; }
        unsigned int nlev = 0, nlines = 0;
        for (int64_t k = 0;; ++k) {
            const uint32_t m = __ballot_sync(0xffffffffu, len > k);
            if (m == 0) break;
            if (len > k) {
                const uint32_t same = __match_any_sync(m, M.col[rs + k] >> 4);
                if ((same & lt) == 0) ++nlines;          // first lane of its line
            }
            ++nlev;
        }
        nlines = __reduce_add_sync(0xffffffffu, nlines);
        if (lane == 0) { atomicAdd(stats, (unsigned long long)nlev); atomicAdd(stats + 1, (unsigned long long)nlines); }
    }
}
// distinct 128-byte lines of the gathered vector one warp-level gather touches (row-major matrix, sampled)
static int csr_gather_lines(cgo_ctx *c, const CsrMat &M, float *lines) {
    unsigned long long *d_stats = c->d_progress + 2, h_stats[2] = {0, 0};
    *lines = 1.0f;
    if (M.nnz == 0) return 0;
    CGO_CUDA(cudaMemsetAsync(d_stats, 0, sizeof(h_stats), c->stream));
    k_gather_lines<<<grid_for((M.nrows + 31) / 32 * 32, c->sms), 256, 0, c->stream>>>(M, d_stats);
    CGO_CUDA(cudaGetLastError());
    CGO_CUDA(cudaMemcpyAsync(h_stats, d_stats, sizeof(h_stats), cudaMemcpyDeviceToHost, c->stream));
    CGO_CUDA(cudaStreamSynchronize(c->stream));
    if (h_stats[0]) *lines = (float)((double)h_stats[1] / (double)h_stats[0]);
    return 0;
}
static int csr_make_sliced(cgo_ctx *c, CsrMat &M) {
    if (M.sliced || M.nnz == 0 || !M.col || !M.val) return 0;
    int32_t *col2 = nullptr; double *val2 = nullptr;
    unsigned long long n_ragged = 0;
    auto body = [&]() -> int {
        CGO_CUDA(cudaMalloc(&col2, sizeof(int32_t) * (size_t)(M.nnz + CSR_PAD)));
        CGO_CUDA(cudaMalloc(&val2, sizeof(double) * (size_t)(M.nnz + CSR_PAD)));
        CGO_CUDA(cudaMemsetAsync(col2 + M.nnz, 0, sizeof(int32_t) * CSR_PAD, c->stream));
        CGO_CUDA(cudaMemsetAsync(val2 + M.nnz, 0, sizeof(double) * CSR_PAD, c->stream));
        unsigned long long *d_ragged = c->d_progress + 1;
        CGO_CUDA(cudaMemsetAsync(d_ragged, 0, sizeof(unsigned long long), c->stream));
        k_slice_permute<<<grid_for((M.nrows + 31) / 32 * 32, c->sms), 256, 0, c->stream>>>(M, M.col, M.val, col2, val2, false,
                                                                                          d_ragged);
        CGO_CUDA(cudaGetLastError());
        CGO_CUDA(cudaMemcpyAsync(&n_ragged, d_ragged, sizeof(n_ragged), cudaMemcpyDeviceToHost, c->stream));
        CGO_CUDA(cudaStreamSynchronize(c->stream));
        return 0;
    };
    int rc = body();
    if (rc) { cudaFree(col2); cudaFree(val2); return rc; }
    cudaFree(M.col); cudaFree(M.val);
    M.col = col2; M.val = val2;
    M.sliced = 1;
    M.ragged = 16 * n_ragged > (unsigned long long)((M.nrows + 31) / 32);   // more than 1 slice in 16
    return 0;
}

template <class Src, class Fin>
static int build_transpose(cgo_ctx *c, const Src &src, const Fin &fin, int64_t nT, CsrMat &AT) {
    cudaStream_t s = c->stream;
    unsigned long long *counts = nullptr, *cursor = nullptr;
    int64_t *perm = nullptr;
    void *tmp = nullptr;
    int rc = 0;
    auto body = [&]() -> int {
        CGO_CUDA(cudaMalloc(&counts, sizeof(unsigned long long) * (size_t)(nT + 1)));
        CGO_CUDA(cudaMemsetAsync(counts, 0, sizeof(unsigned long long) * (size_t)(nT + 1), s));
        if (src.total > 0) k_tr_count<<<grid_for(src.total, c->sms), 256, 0, s>>>(src, counts);
        CGO_CUDA(cudaGetLastError());
        AT.nrows = nT;
        CGO_CUDA(cudaMalloc(&AT.rowptr, sizeof(int64_t) * (size_t)(nT + 1 + CSR_PAD)));
        size_t tmp_bytes = 0;
        CGO_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, (const int64_t *)counts, AT.rowptr, nT + 1, s));
        CGO_CUDA(cudaMalloc(&tmp, tmp_bytes > 0 ? tmp_bytes : 1));
        CGO_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, (const int64_t *)counts, AT.rowptr, nT + 1, s));
        int64_t nnzT = 0;
        CGO_CUDA(cudaMemcpyAsync(&nnzT, AT.rowptr + nT, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
        CGO_CUDA(cudaStreamSynchronize(s));
        AT.nnz = nnzT;
        cudaFree(counts); counts = nullptr;
        cudaFree(tmp); tmp = nullptr;
        CGO_CUDA(cudaMalloc(&cursor, sizeof(unsigned long long) * (size_t)(nT + 1)));
        CGO_CUDA(cudaMemcpyAsync(cursor, AT.rowptr, sizeof(int64_t) * (size_t)(nT + 1), cudaMemcpyDeviceToDevice, s));
        CGO_CUDA(cudaMalloc(&perm, sizeof(int64_t) * (size_t)(nnzT > 0 ? nnzT : 1)));
        if (src.total > 0) k_tr_fill<<<grid_for(src.total, c->sms), 256, 0, s>>>(src, cursor, perm);
        CGO_CUDA(cudaGetLastError());
        if (nT > 0) k_tr_sort<<<grid_for(nT, c->sms), 256, 0, s>>>(AT.rowptr, nT, perm);
        CGO_CUDA(cudaGetLastError());
        CGO_CUDA(cudaMalloc(&AT.col, sizeof(int32_t) * (size_t)(nnzT + CSR_PAD)));
        CGO_CUDA(cudaMalloc(&AT.val, sizeof(double) * (size_t)(nnzT + CSR_PAD)));
        if (nnzT > 0) k_tr_finalize<<<grid_for(nnzT, c->sms), 256, 0, s>>>(fin, nnzT, perm, AT);
        CGO_CUDA(cudaGetLastError());
        CGO_CUDA(cudaStreamSynchronize(s));
        return 0;
    };
    rc = body();
    cudaFree(counts); cudaFree(cursor); cudaFree(perm); cudaFree(tmp);
    return rc;
}

// ------------------------------------------------------------------ column blocking (setup)
// A gather that ranges over more than about half of the 126 MB L2 (measured cliff: 48 MB still
// hits, 60 MB starts missing) turns every 8-byte gather into a 32-byte DRAM sector read.  For
// such matrices the columns are cut into nb blocks of at most ctx->gather_block_bytes and the
// matrix is stored block-major: blk[j] holds the entries of every row whose column lies in block
// j (rows' columns must ascend, which both logistic-regression matrices do by construction).  One
// pass per block keeps its window of the gathered vector L2-resident; passes chain the row sums
// (WithInit), so the arithmetic order is unchanged.
struct CsrBlocked {
    std::vector<CsrMat> blk;
    void free_all() { for (auto &m : blk) csr_free(m); blk.clear(); }
};
__global__ void k_blk_check_sorted(CsrMat A, int *bad) {
    GRID_STRIDE(i, A.nrows) {
        for (int64_t p = A.rowptr[i] + 1; p < A.rowptr[i + 1]; ++p)
            if (A.col[p] < A.col[p - 1]) { *bad = 1; break; }
    }
}
// entries of row i with c0 <= col < c1: first[i] = their start in A, cnt[i] = how many
__global__ void k_blk_count(CsrMat A, int32_t c0, int32_t c1, int64_t *first, int64_t *cnt) {
    GRID_STRIDE(i, A.nrows + 1) {
        if (i == A.nrows) { cnt[i] = 0; continue; }
        const int64_t rs = A.rowptr[i], re = A.rowptr[i + 1];
        int64_t lo = rs, hi = re;                // lower_bound(c0)
        while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (A.col[mid] < c0) lo = mid + 1; else hi = mid; }
        const int64_t b0 = lo;
        hi = re;                                 // lower_bound(c1)
        while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (A.col[mid] < c1) lo = mid + 1; else hi = mid; }
        first[i] = b0;
        cnt[i] = lo - b0;
    }
}
__global__ void k_blk_copy(CsrMat A, const int64_t *first, CsrMat B) {
    GRID_STRIDE(i, A.nrows) {
        const int64_t d0 = B.rowptr[i], n = B.rowptr[i + 1] - d0, s0 = first[i];
        for (int64_t e = 0; e < n; ++e) { B.col[d0 + e] = A.col[s0 + e]; B.val[d0 + e] = A.val[s0 + e]; }
    }
}
static int build_blocked(cgo_ctx *c, const CsrMat &A, int64_t ncols, CsrBlocked &out) {
    out.free_all();
    const int64_t cap = (int64_t)(c->gather_block_bytes / sizeof(double));
    if (cap <= 0 || ncols <= cap || A.nnz == 0) return 0;
    const int64_t nb = (ncols + cap - 1) / cap;
    const int64_t bc = (((ncols + nb - 1) / nb) + 1) & ~(int64_t)1;
    cudaStream_t s = c->stream;
    int *bad = nullptr;
    int64_t *first = nullptr, *cnt = nullptr;
    void *tmp = nullptr;
    auto body = [&]() -> int {
        CGO_CUDA(cudaMalloc(&bad, sizeof(int)));
        CGO_CUDA(cudaMemsetAsync(bad, 0, sizeof(int), s));
        k_blk_check_sorted<<<grid_for(A.nrows, c->sms), 256, 0, s>>>(A, bad);
        int hbad = 0;
        CGO_CUDA(cudaMemcpyAsync(&hbad, bad, sizeof(int), cudaMemcpyDeviceToHost, s));
        CGO_CUDA(cudaStreamSynchronize(s));
        if (hbad) return 0;                      // columns not ascending: stay single-pass
        CGO_CUDA(cudaMalloc(&first, sizeof(int64_t) * (size_t)(A.nrows + 1)));
        CGO_CUDA(cudaMalloc(&cnt, sizeof(int64_t) * (size_t)(A.nrows + 1)));
        size_t tmp_bytes = 0;
        CGO_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, cnt, cnt, A.nrows + 1, s));
        CGO_CUDA(cudaMalloc(&tmp, tmp_bytes > 0 ? tmp_bytes : 1));
        out.blk.resize((size_t)nb);
        for (int64_t j = 0; j < nb; ++j) {
            CsrMat &B = out.blk[(size_t)j];
            const int64_t c0 = j * bc, c1 = (j + 1) * bc < ncols ? (j + 1) * bc : ncols;
            k_blk_count<<<grid_for(A.nrows + 1, c->sms), 256, 0, s>>>(A, (int32_t)c0, (int32_t)c1, first, cnt);
            CGO_CUDA(cudaGetLastError());
            B.nrows = A.nrows;
            CGO_CUDA(cudaMalloc(&B.rowptr, sizeof(int64_t) * (size_t)(A.nrows + 1 + CSR_PAD)));
            CGO_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, cnt, B.rowptr, A.nrows + 1, s));
            int64_t nnzb = 0;
            CGO_CUDA(cudaMemcpyAsync(&nnzb, B.rowptr + A.nrows, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
            CGO_CUDA(cudaStreamSynchronize(s));
            B.nnz = nnzb;
            CGO_CUDA(cudaMalloc(&B.col, sizeof(int32_t) * (size_t)(nnzb + CSR_PAD)));
            CGO_CUDA(cudaMalloc(&B.val, sizeof(double) * (size_t)(nnzb + CSR_PAD)));
            k_blk_copy<<<grid_for(A.nrows, c->sms), 256, 0, s>>>(A, first, B);
            CGO_CUDA(cudaGetLastError());
        }
        CGO_CUDA(cudaStreamSynchronize(s));
        return 0;
    };
    int rc = body();
    cudaFree(bad); cudaFree(first); cudaFree(cnt); cudaFree(tmp);
    if (rc) out.free_all();
    return rc;
}
// one pass per column block; the last pass runs the real epilogue
// (`wait_first`: red's flag wait guards the gathered vector, so the FIRST pass has to honour it)
template <class Epi>
static int launch_csr_blocked(cgo_ctx *c, const CsrMat &A, const CsrBlocked &B, const double *xg, const Epi &epi,
                              double *partial, const RedArgs &red, int tclass, bool wait_first = false) {
    if (B.blk.empty()) return launch_csr(c, A, xg, epi, red, tclass);
    const size_t nb = B.blk.size();
    RedArgs scratch = cgo_red_args(c, CGO_PACK_LEN - 1);
    if (c->csr_pass_occ == 3) scratch.G = c->sms * 3;     // nothing is reduced in these passes: any sweep order will do
    if (wait_first) {
        scratch.wait0 = red.wait0; scratch.wait1 = red.wait1; scratch.wait_all = red.wait_all;
        scratch.wait_val = red.wait_val; scratch.nranks = red.nranks;
    }
    for (size_t j = 0; j + 1 < nb; ++j) {
        EpiStore es{partial};
        if (j == 0) CGO_TRY(launch_csr_pass(c, B.blk[j], xg, es, scratch, tclass));
        else CGO_TRY(launch_csr_pass(c, B.blk[j], xg, with_init(es, partial), scratch, tclass));
    }
    return launch_csr_pass(c, B.blk[nb - 1], xg, with_init(epi, partial), red, tclass);
}

// the same with k_spmv_direct (sliced blocks): every pass just stores the running row sums, the epilogue and its
// reductions are a BLAS-1 kernel of the caller's.  `wait`: flag hand-off that guards the gathered vector (first pass)
static int spmv_direct_passes(cgo_ctx *c, const CsrMat &M, const CsrBlocked &B, const double *xg, double *out,
                              const RedArgs *wait, int tclass) {
    if (B.blk.empty()) return launch_direct(c, M, xg, DirStore{out}, dir_args(c, wait), tclass);
    for (size_t j = 0; j < B.blk.size(); ++j) {
        if (j == 0) CGO_TRY(launch_direct(c, B.blk[j], xg, DirStore{out}, dir_args(c, wait), tclass));
        else CGO_TRY(launch_direct(c, B.blk[j], xg, DirStoreInit{out, out}, dir_args(c), tclass));
    }
    return 0;
}

// ------------------------------------------------------------------ the objective
struct CsrObj : cgo_obj {
    bool logreg = false;
    CsrMat A, AT;
    CsrBlocked Ab, ATb;                // column-blocked copies (logreg, large gather ranges)
    bool have_unblocked = true;        // A / AT arrays still allocated (test hooks need them)
    int64_t nrows = 0;                 // local rows of A (residuals / samples)
    double *b = nullptr;               // rhs (LS) or labels (logreg), nrows
    double *r_base = nullptr, *r = nullptr;   // residual / c vector with halo
    double *qv = nullptr;                     // v = A u of the quadratic-aware line search (nrows)
    // whose residual `r` holds: every trial leaves r(xp) of its state; cgo_accept makes that r(x); the
    // Hessian-vector product and trials of OTHER states on the same objective overwrite it
    const cgo_state *r_state = nullptr;
    bool r_at_xp = false, r_at_x = false;
    int quad_steps = 0;                       // accepted steps since r was last formed as A x − b
    std::vector<void *> rpeers;               // peer mappings of r_base (sharded LS with peer memory)
    bool r_is_peer = false;
    double lambda = 0.0;
    int64_t nsamples = 0;
    // sharded logistic regression (nranks > 1): samples AND features are sharded.  A holds this rank's samples ×
    // all features and gathers the all-gathered trial point; AT holds this rank's FEATURES × all samples (built
    // straight from the generator) and gathers the all-gathered c, so every row of Aᵀc is added over all samples in
    // the single-GPU order.  The state vectors are feature shards [flo[rank], flo[rank+1]).
    bool lr_sharded = false;
    bool lr_direct = false;            // k_spmv_direct passes + BLAS-1 epilogues (sliced matrices)
    std::vector<int64_t> flo, slo;     // feature / sample shard boundaries, nranks + 1
    double *xp_full = nullptr;         // all-gathered trial point, n_global
    double *c_full = nullptr;          // all-gathered c = −y σ, nsamples
    // peer-memory variant: xp_full and c_full are mapped by every rank
    bool lr_peer = false;
    std::vector<void *> xpf_peers, cf_peers;
    void **d_xpf_dst = nullptr;        // rank r's xp_full + flo[me]
    void **d_cf_dst = nullptr;         // rank r's c_full + slo[me]
    ~CsrObj() override {
        if (ctx && cgo_ctx_alive(ctx)) cudaSetDevice(ctx->device);
        csr_free(A); csr_free(AT);
        Ab.free_all(); ATb.free_all();
        if (r_is_peer) { cgo_peer_free(ctx, r_base, rpeers, true); r_base = nullptr; }
        if (lr_peer) {
            cgo_peer_free(ctx, xp_full, xpf_peers, true); xp_full = nullptr;
            cgo_peer_free(ctx, c_full, cf_peers, true); c_full = nullptr;
        }
        cudaFree(d_xpf_dst); cudaFree(d_cf_dst);
        cudaFree(b); cudaFree(r_base); cudaFree(qv);
        cudaFree(xp_full); cudaFree(c_full);
    }
    int alloc_r() {
        size_t bytes = sizeof(double) * (size_t)(nrows + 2 * halo + 4);
        if (ctx->nranks > 1 && ctx->peer_ok && halo > 0) {
            bytes = sizeof(double) * (size_t)((n_alloc > nrows ? n_alloc : nrows) + 2 * halo + 4);   // rank-independent
            void *p = nullptr;
            r_is_peer = true;
            CGO_TRY(cgo_peer_alloc(ctx, bytes, &p, rpeers));
            r_base = (double *)p;
        } else {
            CGO_CUDA(cudaMalloc(&r_base, bytes));
            CGO_CUDA(cudaMemsetAsync(r_base, 0, bytes, ctx->stream));
        }
        r = r_base + halo;
        return 0;
    }
    // local length of rank q's shard (the vector shards and the row shards coincide for LS)
    int64_t shard_len_of(int q) const {
        int64_t lo, hi;
        cgo_shard_range(n_global, ctx->nranks, q, 2, &lo, &hi);
        return hi - lo;
    }
    // sharded least squares over peer memory: K_a and K_b push their boundary elements into the
    // neighbours' halos and hand off with flags; no separate exchange step
    int eval_trial_ls_peer(cgo_state *st, double a, bool fused, double beta, double *out) {
        const int R = ctx->nranks, me = ctx->rank, prev = (me + R - 1) % R, next = (me + 1) % R;
        const unsigned long long e = ++ctx->epoch;
        unsigned long long *fprev = (unsigned long long *)ctx->flags_peer[(size_t)prev];
        unsigned long long *fnext = (unsigned long long *)ctx->flags_peer[(size_t)next];
        const int64_t nprev = shard_len_of(prev);
        HaloPush hp;
        double *xp_prev = (double *)st->xpeers[st->xp_alloc][(size_t)prev] + halo;   // origin of prev's xp
        double *xp_next = (double *)st->xpeers[st->xp_alloc][(size_t)next] + halo;
        hp.prev_right = xp_prev + nprev;
        hp.next_left = xp_next - halo;
        hp.sig_prev = fprev + CGO_F_XP_FROM_NEXT;       // I am prev's next
        hp.sig_next = fnext + CGO_F_XP_FROM_PREV;
        hp.epoch = e;
        CGO_TRY(cgo_blas1_axpy_dir(st, a, fused, beta, &hp));                          // K_a + halo push
        EpiResidualT<true> e1;
        e1.b = b; e1.r = r; e1.halo = halo; e1.nrows = nrows;
        e1.prev_right = (double *)rpeers[(size_t)prev] + halo + nprev;
        e1.next_left = (double *)rpeers[(size_t)next];
        RedArgs rb = cgo_red_args(ctx, CGO_P_PHI);
        rb.wait0 = ctx->flags_local + CGO_F_XP_FROM_PREV; rb.wait1 = ctx->flags_local + CGO_F_XP_FROM_NEXT; rb.wait_val = e;
        rb.sig0 = fprev + CGO_F_R_FROM_NEXT; rb.sig1 = fnext + CGO_F_R_FROM_PREV; rb.sig_val = e;
        CGO_TRY(launch_csr(ctx, A, st->xp, e1, rb, CGO_T_SPMV));                       // K_b + halo push
        EpiGrad<false> e2{st->gp, st->g, st->u, nullptr, 0.0, 0.0};
        RedArgs rc = cgo_red_args(ctx, CGO_P_DPHI);
        rc.wait0 = ctx->flags_local + CGO_F_R_FROM_PREV; rc.wait1 = ctx->flags_local + CGO_F_R_FROM_NEXT; rc.wait_val = e;
        CGO_TRY(launch_csr(ctx, AT, r, e2, rc, CGO_T_SPMVT));                          // K_c
        CGO_TRY(cgo_finish_pack(ctx, 12, out));
        out[CGO_P_PHI] = 0.5 * out[CGO_P_PHI];
        return 0;
    }
    // fill v[−halo, 0) and v[nloc, nloc + halo) from the ring neighbours
    int exchange(double *v, int64_t nloc) {
        if (halo == 0) return 0;
        return cgo_sendrecv_ring(ctx, v, v + nloc, v + nloc - halo, v - halo, halo);
    }
    // ∇²f u = Aᵀ(A u) along the current direction (north_star's Hessian-vector product; the reference
    // engine never calls one, src/engine/optim.jl:83-145).  `r` is scratch between trials.
    // ---- gather-bound matrices (sliced layout): k_spmv_direct + the dots as BLAS-1 passes
    bool direct() const { return A.sliced != 0; }
    void reduction_site(int32_t *V, int32_t *U) const override {
        if (direct() || lr_direct) { *V = 2; *U = CGO_U_VEC; } else { *V = 1; *U = 1; }
    }
    // logistic regression through k_spmv_direct.  One rank: K_a | passes of A → margins | logit + loss (BLAS-1) |
    // passes of Aᵀ | /N + λw + dots (BLAS-1).  R ranks: K_a also stores its shard of xp into every rank's
    // all-gathered copy; the logit kernel stores its shard of c into every rank's all-gathered c; flags hand the
    // data to the consuming passes (without peer memory: two NCCL all-gathers).
    int eval_trial_lr_direct(cgo_state *st, double a, bool fused, double beta, double *out) {
        const int R = ctx->nranks;
        const double invN = 1.0 / (double)nsamples;
        RedArgs wx, wc;                           // flag waits of the first K_b / K_c pass
        const RedArgs *pwx = nullptr, *pwc = nullptr;
        unsigned long long e = 0;
        if (R > 1 && lr_peer) {
            e = ++ctx->epoch;
            HaloPush hp;
            hp.dst_all = d_xpf_dst; hp.epoch = e;
            CGO_TRY(cgo_blas1_axpy_dir(st, a, fused, beta, &hp));                     // K_a + all-gather of xp
            wx.wait_all = ctx->flags_local + CGO_F_XPALL; wx.wait_val = e; wx.nranks = R; pwx = &wx;
            wc.wait_all = ctx->flags_local + CGO_F_GPART; wc.wait_val = e; wc.nranks = R; pwc = &wc;
        } else {
            CGO_TRY(cgo_blas1_axpy_dir(st, a, fused, beta));                          // K_a
            if (R > 1) CGO_TRY(cgo_allgatherv_f64(ctx, st->xp, xp_full, flo.data()));
        }
        const double *w_all = R > 1 ? xp_full : st->xp;
        CGO_TRY(spmv_direct_passes(ctx, A, Ab, w_all, r, pwx, CGO_T_SPMV));            // K_b: margins of my samples
        CGO_TRY(cgo_blas1_logit(ctx, r, b, nrows, CGO_P_PHI, (R > 1 && lr_peer) ? d_cf_dst : nullptr, e));   // c, Σ loss [+ all-gather of c]
        if (R > 1 && !lr_peer) CGO_TRY(cgo_allgatherv_f64(ctx, r, c_full, slo.data()));
        const double *c_all = R > 1 ? c_full : r;
        CGO_TRY(spmv_direct_passes(ctx, AT, ATb, c_all, st->gp, pwc, CGO_T_SPMVT));    // K_c: Aᵀc over my features
        CGO_TRY(cgo_blas1_grad_combine(st, st->gp, 1, 0, invN, lambda));               // /N + λw, dots
        CGO_TRY(cgo_finish_pack(ctx, 12, out));
        out[CGO_P_PHI] = out[CGO_P_PHI] / (double)nsamples + (0.5 * lambda) * out[CGO_P_XPXP];
        return 0;
    }
    int eval_trial_ls_direct(cgo_state *st, double a, bool fused, double beta, double *out) {
        const bool peer = r_is_peer && st->peer_x;
        RedArgs fb, fc;                                                                   // flag hand-offs only
        if (peer) {
            const int R = ctx->nranks, me = ctx->rank, prev = (me + R - 1) % R, next = (me + 1) % R;
            const unsigned long long e = ++ctx->epoch;
            unsigned long long *fprev = (unsigned long long *)ctx->flags_peer[(size_t)prev];
            unsigned long long *fnext = (unsigned long long *)ctx->flags_peer[(size_t)next];
            const int64_t nprev = shard_len_of(prev);
            HaloPush hp;
            hp.prev_right = (double *)st->xpeers[st->xp_alloc][(size_t)prev] + halo + nprev;
            hp.next_left = (double *)st->xpeers[st->xp_alloc][(size_t)next];
            hp.sig_prev = fprev + CGO_F_XP_FROM_NEXT;
            hp.sig_next = fnext + CGO_F_XP_FROM_PREV;
            hp.epoch = e;
            CGO_TRY(cgo_blas1_axpy_dir(st, a, fused, beta, &hp));                         // K_a + halo push
            DirResidual<true> f1{b, r, (double *)rpeers[(size_t)prev] + halo + nprev, (double *)rpeers[(size_t)next], halo, nrows};
            fb.wait0 = ctx->flags_local + CGO_F_XP_FROM_PREV; fb.wait1 = ctx->flags_local + CGO_F_XP_FROM_NEXT; fb.wait_val = e;
            fb.sig0 = fprev + CGO_F_R_FROM_NEXT; fb.sig1 = fnext + CGO_F_R_FROM_PREV; fb.sig_val = e;
            CGO_TRY(launch_direct(ctx, A, st->xp, f1, dir_args(ctx, &fb), CGO_T_SPMV));    // K_b + halo push
            fc.wait0 = ctx->flags_local + CGO_F_R_FROM_PREV; fc.wait1 = ctx->flags_local + CGO_F_R_FROM_NEXT; fc.wait_val = e;
        } else {
            CGO_TRY(cgo_blas1_axpy_dir(st, a, fused, beta));                              // K_a
            CGO_TRY(exchange(st->xp, st->n));
            DirResidual<false> f1{b, r, nullptr, nullptr, 0, nrows};
            CGO_TRY(launch_direct(ctx, A, st->xp, f1, dir_args(ctx), CGO_T_SPMV));         // K_b
        }
        CGO_TRY(cgo_blas1_sumsq(ctx, r, nrows, CGO_P_PHI));                               // Σ r²
        if (!peer) CGO_TRY(exchange(r, nrows));
        DirGradLS f2{st->gp};
        CGO_TRY(launch_direct(ctx, AT, r, f2, peer ? dir_args(ctx, &fc) : dir_args(ctx), CGO_T_SPMVT));   // K_c
        CGO_TRY(cgo_blas1_grad_dots(st));                                                 // dϕ, ‖g⁺‖², getβ dots
        CGO_TRY(cgo_finish_pack(ctx, 12, out));
        out[CGO_P_PHI] = 0.5 * out[CGO_P_PHI];
        return 0;
    }
    int hessvec_dir(cgo_state *st, double *out) override {
        CGO_CHECK(!logreg, "the Hessian-vector product is implemented for least squares");
        r_state = nullptr; r_at_x = r_at_xp = false;                                      // r is scratch here
        CGO_TRY(cgo_sendrecv_ring(ctx, st->u, st->u + st->n, st->u + st->n - halo, st->u - halo, halo));
        if (direct()) {
            CGO_TRY(launch_direct(ctx, A, st->u, DirStore{r}, dir_args(ctx), CGO_T_SPMV));
            CGO_TRY(cgo_blas1_sumsq(ctx, r, nrows, 0));                                   // u·Hu = ‖Au‖²
            CGO_TRY(exchange(r, nrows));
            CGO_TRY(launch_direct(ctx, AT, r, DirStore{st->hv}, dir_args(ctx), CGO_T_SPMVT));
            CGO_TRY(cgo_blas1_dots3(ctx, st->u, st->hv, st->n, 1));                       // u·hv, hv·hv
            return cgo_finish_pack(ctx, 3, out);
        }
        EpiStoreSq e1{r};
        CGO_TRY(launch_csr(ctx, A, st->u, e1, cgo_red_args(ctx, 0), CGO_T_SPMV));
        CGO_TRY(exchange(r, nrows));
        EpiHv e2{st->hv, st->u};
        CGO_TRY(launch_csr(ctx, AT, r, e2, cgo_red_args(ctx, 1), CGO_T_SPMVT));
        return cgo_finish_pack(ctx, 3, out);
    }
    // SURVEY.md §8f N1.  `r` must hold the residual at x: true after state creation and after every
    // accepted step whose last trial was the accepted one (strong-Wolfe / Wolfe searches).
    int quad_begin(cgo_state *st, double *out) override {
        CGO_CHECK(!logreg, "the quadratic-aware line search is for least squares");
        if (!qv) {
            CGO_CUDA(cudaMalloc(&qv, sizeof(double) * (size_t)(nrows + CSR_PAD)));
            CGO_CUDA(cudaMemsetAsync(qv, 0, sizeof(double) * (size_t)(nrows + CSR_PAD), ctx->stream));
        }
        // r must be the residual at THIS state's x.  It is after a trial of this state was accepted; it is not after
        // a Hessian-vector product, after another state used the objective, or — the recursive update r += a v
        // drifts by ≈ ε‖r₀‖ per step — after 64 accepted steps: then r = A x − b is formed again
        if (!(r_state == st && r_at_x) || quad_steps >= 64) {
            CGO_TRY(cgo_sendrecv_ring(ctx, st->x, st->x + st->n, st->x + st->n - halo, st->x - halo, halo));
            if (direct()) {
                DirResidual<false> f1{b, r, nullptr, nullptr, 0, nrows};
                CGO_TRY(launch_direct(ctx, A, st->x, f1, dir_args(ctx), CGO_T_SPMV));
            } else {
                EpiResidualT<false> e0;
                e0.b = b; e0.r = r; e0.prev_right = e0.next_left = nullptr; e0.halo = 0; e0.nrows = nrows;
                CGO_TRY(launch_csr(ctx, A, st->x, e0, cgo_red_args(ctx, CGO_PACK_LEN - 1), CGO_T_SPMV));
            }
            r_state = st; r_at_x = true; r_at_xp = false; quad_steps = 0;
        }
        CGO_TRY(cgo_sendrecv_ring(ctx, st->u, st->u + st->n, st->u + st->n - halo, st->u - halo, halo));
        if (direct()) {
            CGO_TRY(launch_direct(ctx, A, st->u, DirStore{qv}, dir_args(ctx), CGO_T_SPMV));
            CGO_TRY(cgo_blas1_dots3(ctx, r, qv, nrows, 0));                               // r·v, v·v, r·r
            return cgo_finish_pack(ctx, 3, out);
        }
        EpiQuadV e1{qv, r};
        CGO_TRY(launch_csr(ctx, A, st->u, e1, cgo_red_args(ctx, 0), CGO_T_SPMV));
        return cgo_finish_pack(ctx, 3, out);
    }
    int quad_accept(cgo_state *st, double a, double *out) override {
        CGO_CHECK(!logreg && qv != nullptr, "cgo_quad_accept before cgo_quad_begin");
        CGO_CHECK(r_state == st && r_at_x, "cgo_quad_accept: the residual of this state was overwritten since cgo_quad_begin");
        r_at_x = false; r_at_xp = true; ++quad_steps;                                     // r becomes r(xp) below
        CGO_TRY(cgo_blas1_axpy_dir(st, a, false, 0.0));                                   // xp = x + a u
        CGO_TRY(cgo_blas1_residual_axpy(ctx, r, qv, a, nrows, CGO_P_PHI));                // r += a v, Σ r²
        CGO_TRY(exchange(r, nrows));
        if (direct()) {
            CGO_TRY(launch_direct(ctx, AT, r, DirGradLS{st->gp}, dir_args(ctx), CGO_T_SPMVT));
            CGO_TRY(cgo_blas1_grad_dots(st));
            CGO_TRY(cgo_finish_pack(ctx, 12, out));
            out[CGO_P_PHI] = 0.5 * out[CGO_P_PHI];
            return 0;
        }
        EpiGrad<false> e2{st->gp, st->g, st->u, nullptr, 0.0, 0.0};
        CGO_TRY(launch_csr(ctx, AT, r, e2, cgo_red_args(ctx, CGO_P_DPHI), CGO_T_SPMVT));  // K_c
        CGO_TRY(cgo_finish_pack(ctx, 12, out));
        out[CGO_P_PHI] = 0.5 * out[CGO_P_PHI];
        return 0;
    }
    void on_accept(cgo_state *st) override {
        r_at_x = (r_state == st) && r_at_xp;
        r_at_xp = false;
        if (r_state != st) r_state = nullptr;
    }
    int eval_trial(cgo_state *st, double a, bool fused, double beta, double *out) override {
        r_state = st; r_at_xp = true; r_at_x = false; quad_steps = 0;        // r = A xp − b (least squares) / c (logreg)
        if (!logreg && direct()) return eval_trial_ls_direct(st, a, fused, beta, out);
        if (!logreg && r_is_peer && st->peer_x) return eval_trial_ls_peer(st, a, fused, beta, out);
        if (logreg && lr_direct) return eval_trial_lr_direct(st, a, fused, beta, out);
        CGO_TRY(cgo_blas1_axpy_dir(st, a, fused, beta));                               // K_a
        CGO_TRY(exchange(st->xp, st->n));
        if (logreg) {
            EpiLogit e1{b, r};
            CGO_TRY(launch_csr_blocked(ctx, A, Ab, st->xp, e1, r, cgo_red_args(ctx, CGO_P_PHI), CGO_T_SPMV));   // K_b
            EpiGrad<true> e2{st->gp, st->g, st->u, st->xp, 1.0 / (double)nsamples, lambda};
            CGO_TRY(launch_csr_blocked(ctx, AT, ATb, r, e2, st->gp, cgo_red_args(ctx, CGO_P_DPHI), CGO_T_SPMVT)); // K_c
            CGO_TRY(cgo_finish_pack(ctx, 12, out));
            out[CGO_P_PHI] = out[CGO_P_PHI] / (double)nsamples + (0.5 * lambda) * out[CGO_P_XPXP];
        } else {
            EpiResidualT<false> e1;
            e1.b = b; e1.r = r; e1.prev_right = e1.next_left = nullptr; e1.halo = 0; e1.nrows = nrows;
            CGO_TRY(launch_csr(ctx, A, st->xp, e1, cgo_red_args(ctx, CGO_P_PHI), CGO_T_SPMV));   // K_b
            CGO_TRY(exchange(r, nrows));
            EpiGrad<false> e2{st->gp, st->g, st->u, nullptr, 0.0, 0.0};
            CGO_TRY(launch_csr(ctx, AT, r, e2, cgo_red_args(ctx, CGO_P_DPHI), CGO_T_SPMVT));      // K_c
            CGO_TRY(cgo_finish_pack(ctx, 12, out));
            out[CGO_P_PHI] = 0.5 * out[CGO_P_PHI];
        }
        return 0;
    }
    // A and Aᵀ streamed once (8 B value + 4 B index per entry + row pointers)
    // plus the vector passes R x,u W xp | gather xp, R b, W r | gather r, W g⁺, R u (,g)
    double bytes_per_eval() const override {
        if (lr_direct)      // K_a 24d_loc | A_r + gather + W z | logit R z, y W c | Aᵀ_f + gather + W q | R q, u, g, w W g⁺
            return 12.0 * (double)(A.nnz + AT.nnz) + 8.0 * (double)(A.nrows + 1 + AT.nrows + 1) +
                   8.0 * (5.0 * (double)nrows + (ctx->nranks > 1 ? (double)n_global + (double)nsamples : 0.0) + 11.0 * (double)n_local);
        // (gather-bound matrices: the dots are separate BLAS-1 passes, R r | R g⁺, g, u instead of the staged u, g)
        return 12.0 * (double)(A.nnz + AT.nnz) + 8.0 * (double)(A.nrows + 1 + AT.nrows + 1) +
               8.0 * (3.0 * (double)nrows + 6.0 * (double)n_local) + (direct() ? 8.0 * ((double)nrows + (double)n_local) : 0.0);
    }
    int default_x0(uint64_t, double, double *x0) override {
        for (int64_t i = 0; i < n_local; ++i) x0[i] = 0.0;
        return 0;
    }
};

static int check_i32(int64_t v, const char *what) {
    CGO_CHECK(v < 2147483647LL, "%s = %lld does not fit the int32 column index", what, (long long)v);
    return 0;
}

extern "C" int cgo_obj_sparse_ls_create_synthetic(cgo_ctx *ctx, int64_t n, int32_t K, int64_t W,
                                                  uint64_t seed, int32_t coh_log2, cgo_obj **out) {
    CGO_CHECK(ctx && out, "NULL argument");
    const int64_t S = K - 1;
    CGO_CHECK(K >= 1 && n >= 2 && n % 2 == 0, "sparse_ls: need nnz_per_row >= 1 and an even n >= 2");
    CGO_CHECK(S == 0 || (W >= S && n > 2 * W), "sparse_ls: need W >= nnz_per_row-1 and n > 2W (n=%lld, W=%lld)", (long long)n, (long long)W);
    CGO_CHECK(coh_log2 >= 0 && coh_log2 < 31, "sparse_ls: coh_log2 out of range");
    CGO_CUDA(cudaSetDevice(ctx->device));
    int64_t lo, hi, n_alloc = 0;
    CGO_TRY(cgo_shard_range(n, ctx->nranks, ctx->rank, 2, &lo, &hi));
    for (int q = 0; q < ctx->nranks; ++q) {
        int64_t a, b;
        CGO_TRY(cgo_shard_range(n, ctx->nranks, q, 2, &a, &b));
        if (b - a > n_alloc) n_alloc = b - a;
    }
    const bool multi = ctx->nranks > 1;
    if (multi) {
        CGO_CHECK(hi - lo >= W && hi - lo + 2 * W <= n, "sparse_ls: shard of %lld rows needs W <= rows and rows + 2W <= n (W=%lld, n=%lld)",
                  (long long)(hi - lo), (long long)W, (long long)n);
        CGO_CHECK(W % 2 == 0, "sparse_ls: W must be even with more than one rank");
    }
    CsrObj *o = new CsrObj();          // (nothing below returns without deleting it)
    o->ctx = ctx; o->n_global = n;
    o->offset = lo; o->n_local = hi - lo; o->nrows = hi - lo; o->n_alloc = n_alloc;
    o->halo = multi ? W : 0;
    LsSpec sp;
    sp.n = n; sp.K = K; sp.coh = coh_log2; sp.W = W; sp.w = S > 0 ? (2 * W) / S : 0; sp.seed = seed;
    const int64_t nloc = o->n_local, nnz = nloc * K;
    double *xt_base = nullptr;
    auto body = [&]() -> int {
        CGO_TRY(check_i32(nloc + 2 * o->halo, "local dimension + halo"));
        CGO_TRY(csr_alloc(o->A, nloc, nnz));
        k_ls_fill_A<<<grid_for(nnz, ctx->sms), 256, 0, ctx->stream>>>(sp, lo, nloc, multi, o->A);
        CGO_CUDA(cudaGetLastError());
        // b = A x_true
        CGO_CUDA(cudaMalloc(&xt_base, sizeof(double) * (size_t)(nloc + 2 * o->halo)));
        k_ls_xtrue<<<grid_for(nloc + 2 * o->halo, ctx->sms), 256, 0, ctx->stream>>>(sp, lo, nloc, o->halo, xt_base + o->halo);
        CGO_CUDA(cudaGetLastError());
        CGO_CUDA(cudaMalloc(&o->b, sizeof(double) * (size_t)(nloc + CSR_PAD)));
        EpiStore es{o->b};
        CGO_TRY(launch_csr(ctx, o->A, xt_base + o->halo, es, cgo_red_args(ctx, CGO_PACK_LEN - 1), CGO_T_OTHER));
        CGO_CUDA(cudaStreamSynchronize(ctx->stream));
        cudaFree(xt_base); xt_base = nullptr;
        CGO_TRY(o->alloc_r());
        if (multi) {
            LsExtSrc src{sp, lo, hi, (nloc + 2 * W) * K};
            LsExtFin fin{sp, lo};
            CGO_TRY(build_transpose(ctx, src, fin, nloc, o->AT));
        } else {
            CsrSrc src{o->A.col, nnz};
            FixedKFin fin{K, o->A.val};
            CGO_TRY(build_transpose(ctx, src, fin, nloc, o->AT));
        }
        // offsets redrawn at least every 8 rows: no two lanes of a warp-level gather share a line ⇒ gather-bound
        // (the rule is part of the canonical order: include/cgoptim.h; CGO_CSR_MODE = 1 / 2 forces never / always)
        if (ctx->csr_mode == 2 || (ctx->csr_mode == 0 && coh_log2 < 4 && K > 1)) {
            CGO_TRY(csr_make_sliced(ctx, o->A));
            CGO_TRY(csr_make_sliced(ctx, o->AT));
        }
        return 0;
    };
    int rc = body();
    cudaFree(xt_base);
    if (rc) { delete o; return rc; }
    *out = o;
    return 0;
}

extern "C" int cgo_obj_sparse_ls_create_csr(cgo_ctx *ctx, int64_t nrows, int64_t ncols, const int64_t *rowptr,
                                            const int32_t *col, const double *val, const double *b,
                                            cgo_obj **out) {
    CGO_CHECK(ctx && rowptr && b && out, "NULL argument");
    CGO_CHECK(ctx->nranks == 1, "cgo_obj_sparse_ls_create_csr is single-GPU");
    CGO_CHECK(nrows >= 1 && ncols >= 1, "empty matrix");
    const int64_t nnz = rowptr[nrows];
    CGO_CHECK(nnz == 0 || (col && val), "NULL col/val");
    for (int64_t i = 0; i < nrows; ++i) CGO_CHECK(rowptr[i] <= rowptr[i + 1], "rowptr not monotone at row %lld", (long long)i);
    for (int64_t p = 0; p < nnz; ++p) CGO_CHECK(col[p] >= 0 && col[p] < ncols, "column index out of range at entry %lld", (long long)p);
    CGO_CUDA(cudaSetDevice(ctx->device));
    CsrObj *o = new CsrObj();
    o->ctx = ctx; o->n_global = ncols; o->n_local = ncols; o->offset = 0; o->nrows = nrows; o->halo = 0;
    auto body = [&]() -> int {
        CGO_TRY(check_i32(nrows > ncols ? nrows : ncols, "matrix dimension"));
        CGO_TRY(csr_alloc(o->A, nrows, nnz));
        cudaStream_t s = ctx->stream;
        CGO_CUDA(cudaMemcpyAsync(o->A.rowptr, rowptr, sizeof(int64_t) * (size_t)(nrows + 1), cudaMemcpyHostToDevice, s));
        if (nnz) {
            CGO_CUDA(cudaMemcpyAsync(o->A.col, col, sizeof(int32_t) * (size_t)nnz, cudaMemcpyHostToDevice, s));
            CGO_CUDA(cudaMemcpyAsync(o->A.val, val, sizeof(double) * (size_t)nnz, cudaMemcpyHostToDevice, s));
        }
        CGO_CUDA(cudaMalloc(&o->b, sizeof(double) * (size_t)(nrows + CSR_PAD)));
        CGO_CUDA(cudaMemcpyAsync(o->b, b, sizeof(double) * (size_t)nrows, cudaMemcpyHostToDevice, s));
        CGO_TRY(o->alloc_r());
        CsrSrc src{o->A.col, nnz};
        CsrFin fin{o->A.rowptr, nrows, o->A.val};
        CGO_TRY(build_transpose(ctx, src, fin, ncols, o->AT));
        float la = 1.f, lt = 1.f;
        CGO_TRY(csr_gather_lines(ctx, o->A, &la));
        CGO_TRY(csr_gather_lines(ctx, o->AT, &lt));
        o->A.lines_per_gather = la; o->AT.lines_per_gather = lt;
        if (ctx->csr_mode == 2 || (ctx->csr_mode == 0 && la > 4.0f && lt > 4.0f && nrows >= 4096)) {
            CGO_TRY(csr_make_sliced(ctx, o->A));
            CGO_TRY(csr_make_sliced(ctx, o->AT));
        }
        return 0;
    };
    int rc = body();
    if (rc) { delete o; return rc; }
    *out = o;
    return 0;
}

// rows [flo, fhi) of Aᵀ over ALL samples, straight from the generator (sharded logistic regression): feature c
// lies in stratum c / w, so only the entries k of the strata that overlap the shard have to be looked at
struct LrFeatSrc {
    LrSpec s;
    int64_t flo, fhi, total;        // total = N * nk candidates
    int32_t k0, nk;
    __device__ __forceinline__ bool get(int64_t e, int64_t &trow, int64_t &key) const {
        const int64_t i = e / nk;
        const int k = k0 + (int)(e - i * nk);
        int64_t c; double v;
        s.entry(i, k, c, v);
        if (c < flo || c >= fhi) return false;
        trow = c - flo; key = i * s.K + k;
        return true;
    }
};
struct LrFeatFin {                  // key = global entry index of the generator; column = global sample index
    LrSpec s;
    __device__ __forceinline__ void get(int64_t key, int32_t &srow, double &v) const {
        const int64_t i = key / s.K;
        int64_t c;
        s.entry(i, (int)(key - i * s.K), c, v);
        srow = (int32_t)i;
    }
};

extern "C" int cgo_obj_logreg_create_synthetic(cgo_ctx *ctx, int64_t N, int64_t d, int32_t K, uint64_t seed,
                                               double lambda, cgo_obj **out) {
    CGO_CHECK(ctx && out, "NULL argument");
    CGO_CHECK(K >= 1 && N >= 1 && d >= K, "logreg: need nnz_per_row >= 1, nsamples >= 1, nfeat >= nnz_per_row");
    const int R = ctx->nranks, me = ctx->rank;
    CGO_CHECK(R == 1 || (N >= 2 * R && d >= 2 * R), "logreg: %d ranks need nsamples >= %d and nfeat >= %d", R, 2 * R, 2 * R);
    CGO_CHECK(R == 1 || ctx->csr_mode != 1, "logreg: the sharded objective runs on k_spmv_direct (CGO_CSR_MODE=1 forbids it)");
    CGO_CUDA(cudaSetDevice(ctx->device));
    CsrObj *o = new CsrObj();
    o->ctx = ctx; o->logreg = true; o->lambda = lambda; o->nsamples = N;
    o->n_global = d; o->halo = 0;
    LrSpec sp;
    sp.N = N; sp.d = d; sp.K = K; sp.w = d / K; sp.seed = seed;
    auto body = [&]() -> int {
        // samples (rows of A) and features (state vectors, rows of Aᵀ) are both sharded contiguously
        o->flo.assign(R + 1, 0); o->slo.assign(R + 1, 0);
        for (int r = 0; r < R; ++r) {
            int64_t lo, hi;
            CGO_TRY(cgo_shard_range(d, R, r, 2, &lo, &hi));
            o->flo[r] = lo; o->flo[r + 1] = hi;
            CGO_TRY(cgo_shard_range(N, R, r, 2, &lo, &hi));
            o->slo[r] = lo; o->slo[r + 1] = hi;
        }
        const int64_t slo = o->slo[me], nloc = o->slo[me + 1] - o->slo[me];
        o->offset = o->flo[me]; o->n_local = o->flo[me + 1] - o->flo[me];
        o->nrows = nloc;
        o->lr_sharded = R > 1;
        CGO_TRY(check_i32(N > d ? N : d, "matrix dimension"));
        CGO_TRY(csr_alloc(o->A, nloc, nloc * K));
        CGO_CUDA(cudaMalloc(&o->b, sizeof(double) * (size_t)(nloc + CSR_PAD)));
        CGO_CUDA(cudaMemsetAsync(o->b, 0, sizeof(double) * (size_t)(nloc + CSR_PAD), ctx->stream));
        k_lr_fill_A<<<grid_for(nloc, ctx->sms), 256, 0, ctx->stream>>>(sp, slo, nloc, o->A, o->b);
        CGO_CUDA(cudaGetLastError());
        CGO_TRY(o->alloc_r());
        if (R == 1) {
            CsrSrc src{o->A.col, nloc * K};
            FixedKFin fin{K, o->A.val};
            CGO_TRY(build_transpose(ctx, src, fin, d, o->AT));
        } else {
            const int64_t flo = o->flo[me], fhi = o->flo[me + 1];
            int64_t k0 = flo / sp.w, k1 = (fhi - 1) / sp.w;
            if (k0 > K - 1) k0 = K - 1;
            if (k1 > K - 1) k1 = K - 1;
            LrFeatSrc src{sp, flo, fhi, N * (k1 - k0 + 1), (int32_t)k0, (int32_t)(k1 - k0 + 1)};
            LrFeatFin fin{sp};
            CGO_TRY(build_transpose(ctx, src, fin, fhi - flo, o->AT));
        }
        // K_b gathers over all d features, K_c over all N samples
        CGO_TRY(build_blocked(ctx, o->A, d, o->Ab));
        CGO_TRY(build_blocked(ctx, o->AT, N, o->ATb));
        const bool blocked = !o->Ab.blk.empty() || !o->ATb.blk.empty();
        if (blocked && o->A.nnz > (int64_t)200000000 / R) {
            // production sizes: drop the single-pass copies (only the CSR test hooks read them)
            if (!o->Ab.blk.empty()) { cudaFree(o->A.col); cudaFree(o->A.val); o->A.col = nullptr; o->A.val = nullptr; }
            if (!o->ATb.blk.empty()) { cudaFree(o->AT.col); cudaFree(o->AT.val); o->AT.col = nullptr; o->AT.val = nullptr; }
            o->have_unblocked = false;
        }
        // uniformly random columns: gather-bound (k_spmv_direct + BLAS-1 epilogues) whenever the gathers are what
        // costs — column-blocked sizes, and always when sharded
        o->lr_direct = ctx->csr_mode != 1 && (R > 1 || blocked || ctx->csr_mode == 2);
        if (o->lr_direct) {
            if (o->Ab.blk.empty()) CGO_TRY(csr_make_sliced(ctx, o->A));
            if (o->ATb.blk.empty()) CGO_TRY(csr_make_sliced(ctx, o->AT));
            for (auto &m : o->Ab.blk) CGO_TRY(csr_make_sliced(ctx, m));
            for (auto &m : o->ATb.blk) CGO_TRY(csr_make_sliced(ctx, m));
        }
        if (o->lr_sharded) {
            const size_t xb = sizeof(double) * (size_t)(d + CSR_PAD), cb = sizeof(double) * (size_t)(N + CSR_PAD);
            if (ctx->peer_ok) {
                void *p = nullptr;
                o->lr_peer = true;
                CGO_TRY(cgo_peer_alloc(ctx, xb, &p, o->xpf_peers));
                o->xp_full = (double *)p;
                CGO_TRY(cgo_peer_alloc(ctx, cb, &p, o->cf_peers));
                o->c_full = (double *)p;
                std::vector<void *> xd((size_t)R), cd((size_t)R);
                for (int r = 0; r < R; ++r) {
                    xd[(size_t)r] = (double *)o->xpf_peers[(size_t)r] + o->flo[(size_t)me];
                    cd[(size_t)r] = (double *)o->cf_peers[(size_t)r] + o->slo[(size_t)me];
                }
                CGO_CUDA(cudaMalloc(&o->d_xpf_dst, sizeof(void *) * (size_t)R));
                CGO_CUDA(cudaMalloc(&o->d_cf_dst, sizeof(void *) * (size_t)R));
                CGO_CUDA(cudaMemcpy(o->d_xpf_dst, xd.data(), sizeof(void *) * (size_t)R, cudaMemcpyHostToDevice));
                CGO_CUDA(cudaMemcpy(o->d_cf_dst, cd.data(), sizeof(void *) * (size_t)R, cudaMemcpyHostToDevice));
            } else {
                CGO_CUDA(cudaMalloc(&o->xp_full, xb));
                CGO_CUDA(cudaMalloc(&o->c_full, cb));
                CGO_CUDA(cudaMemsetAsync(o->xp_full, 0, xb, ctx->stream));
                CGO_CUDA(cudaMemsetAsync(o->c_full, 0, cb, ctx->stream));
            }
            CGO_CUDA(cudaStreamSynchronize(ctx->stream));
        }
        return 0;
    };
    int rc = body();
    if (rc) { delete o; return rc; }
    *out = o;
    return 0;
}

// ------------------------------------------------------------------ test hooks
static CsrObj *as_csr(cgo_obj *o) { return dynamic_cast<CsrObj *>(o); }

extern "C" int cgo_obj_csr_blocks(cgo_obj *obj, int transposed, int32_t *nblocks) {
    CsrObj *o = as_csr(obj);
    CGO_CHECK(o && nblocks, "not a CSR objective / NULL argument");
    const CsrBlocked &B = transposed ? o->ATb : o->Ab;
    *nblocks = B.blk.empty() ? 1 : (int32_t)B.blk.size();
    return 0;
}
extern "C" int cgo_obj_csr_nnz(cgo_obj *obj, int transposed, int64_t *nrows, int64_t *nnz) {
    CsrObj *o = as_csr(obj);
    CGO_CHECK(o != nullptr, "not a CSR objective");
    const CsrMat &M = transposed ? o->AT : o->A;
    if (nrows) *nrows = M.nrows;
    if (nnz) *nnz = M.nnz;
    return 0;
}
extern "C" int cgo_obj_csr_download(cgo_obj *obj, int transposed, int64_t *rowptr, int32_t *col, double *val, double *b) {
    CsrObj *o = as_csr(obj);
    CGO_CHECK(o != nullptr, "not a CSR objective");
    CGO_CHECK(o->have_unblocked, "the single-pass CSR copy was dropped at this size (column-blocked storage only)");
    const CsrMat &M = transposed ? o->AT : o->A;
    cudaStream_t s = o->ctx->stream;
    CGO_CUDA(cudaSetDevice(o->ctx->device));
    if (rowptr) CGO_CUDA(cudaMemcpyAsync(rowptr, M.rowptr, sizeof(int64_t) * (size_t)(M.nrows + 1), cudaMemcpyDeviceToHost, s));
    int32_t *ctmp = nullptr; double *vtmp = nullptr;          // row-major copies of a sliced matrix
    auto body = [&]() -> int {
        const int32_t *csrc = M.col; const double *vsrc = M.val;
        if (M.sliced && M.nnz && (col || val)) {
            CGO_CUDA(cudaMalloc(&ctmp, sizeof(int32_t) * (size_t)M.nnz));
            CGO_CUDA(cudaMalloc(&vtmp, sizeof(double) * (size_t)M.nnz));
            k_slice_permute<<<grid_for((M.nrows + 31) / 32 * 32, o->ctx->sms), 256, 0, s>>>(M, M.col, M.val, ctmp, vtmp, true);
            CGO_CUDA(cudaGetLastError());
            csrc = ctmp; vsrc = vtmp;
        }
        if (col && M.nnz) CGO_CUDA(cudaMemcpyAsync(col, csrc, sizeof(int32_t) * (size_t)M.nnz, cudaMemcpyDeviceToHost, s));
        if (val && M.nnz) CGO_CUDA(cudaMemcpyAsync(val, vsrc, sizeof(double) * (size_t)M.nnz, cudaMemcpyDeviceToHost, s));
        if (b) CGO_CUDA(cudaMemcpyAsync(b, o->b, sizeof(double) * (size_t)o->nrows, cudaMemcpyDeviceToHost, s));
        CGO_CUDA(cudaStreamSynchronize(s));
        return 0;
    };
    int rc = body();
    cudaFree(ctmp); cudaFree(vtmp);
    return rc;
}
// y = A x (transposed: Aᵀ x) through the production kernel; single GPU (no halo)
extern "C" int cgo_obj_spmv(cgo_obj *obj, int transposed, const double *x_host, double *y_host) {
    CsrObj *o = as_csr(obj);
    CGO_CHECK(o && x_host && y_host, "not a CSR objective / NULL argument");
    CGO_CHECK(o->halo == 0 && !o->lr_sharded, "cgo_obj_spmv is a single-GPU test hook");
    CGO_CHECK(o->have_unblocked, "the single-pass CSR copy was dropped at this size (column-blocked storage only)");
    cgo_ctx *c = o->ctx;
    CGO_CUDA(cudaSetDevice(c->device));
    const CsrMat &M = transposed ? o->AT : o->A;
    const int64_t nin = transposed ? o->A.nrows : o->AT.nrows, nout = M.nrows;
    double *dx = nullptr, *dy = nullptr;
    auto body = [&]() -> int {
        CGO_CUDA(cudaMalloc(&dx, sizeof(double) * (size_t)nin));
        CGO_CUDA(cudaMalloc(&dy, sizeof(double) * (size_t)nout));
        CGO_CUDA(cudaMemcpyAsync(dx, x_host, sizeof(double) * (size_t)nin, cudaMemcpyHostToDevice, c->stream));
        if (M.sliced) {
            CGO_TRY(launch_direct(c, M, dx, DirStore{dy}, dir_args(c), transposed ? CGO_T_SPMVT : CGO_T_SPMV));
        } else {
            EpiStore es{dy};
            CGO_TRY(launch_csr(c, M, dx, es, cgo_red_args(c, CGO_PACK_LEN - 1), transposed ? CGO_T_SPMVT : CGO_T_SPMV));
        }
        CGO_CUDA(cudaMemcpyAsync(y_host, dy, sizeof(double) * (size_t)nout, cudaMemcpyDeviceToHost, c->stream));
        CGO_CUDA(cudaStreamSynchronize(c->stream));
        return 0;
    };
    int rc = body();
    cudaFree(dx); cudaFree(dy);
    return rc;
}
