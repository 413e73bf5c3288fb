// gather_bench.cu — what is the ceiling for 8-byte random gathers out of an L2-resident window on B200?
//
// The per-row-random variant of the cfg-3 matrix (coh_log2 = 0) gathers 9 doubles per row, each from a
// different 1.86 MB stratum of a ±2^20 band: no two lanes of a warp share a 128-byte line.  This
// microbenchmark measures the gather rate of that access pattern WITHOUT the matrix stream, through every
// load path sm_100a offers, so that k_csr_rows can be judged against the path's own ceiling:
//   mode 0  ld.global.nc.f64                         (LDG.E.64.CONSTANT, 32 distinct lines per instruction)
//   mode 1  ld.global.nc.L1::no_allocate.f64
//   mode 2  ld.global.cg.f64                         (L2 only)
//   mode 3  ld.global.nc.L2::cache_hint evict_last   (what k_csr_rows issues)
//   mode 4  predicated: 2 active lanes per instruction (16 instructions per gather slot)
//   mode 5  predicated: 1 active lane per instruction
//   mode 6  predicated: 4 active lanes per instruction
//   mode 7  predicated: 8 active lanes per instruction
//   mode 8  TMA tile::gather4 (4 rows of 16 bytes per instruction) into shared memory
//   mode 9  cp.async (LDGSTS) 8 bytes per lane into shared memory
//   mode 10 coalesced reference (col = row + k): the harness's own ceiling
//   mode 12 SpMV prototype: mode 0 plus a coalesced 8-byte value stream per entry and y[row] = sum val*x[col] written per row
//   mode 11 tex1Dfetch<int2> through a linear texture object (TEX pipe instead of LSU)
// Output: G gathers/s and gathers per clock per SM (clock64 of block 0).
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o scratch/gather_bench scratch/gather_bench.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

constexpr int KG = 8;          // gathers per row (8 so that gather4 divides it)
constexpr int B = 256;

__host__ __device__ inline uint64_t mix64(uint64_t z) {
    z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ULL;
    z ^= z >> 27; z *= 0x94D049BB133111EBULL;
    z ^= z >> 31;
    return z;
}

// col[(tile * KG + k) * B + t]: coalesced index loads
__global__ void k_fill(int32_t *col, int64_t nrows, int64_t n, int64_t W, int coalesced) {
    const int64_t total = nrows * KG;
    const int64_t w = 2 * W / KG;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < total; p += (int64_t)gridDim.x * blockDim.x) {
        const int64_t tile = p / (KG * B);
        const int k = (int)((p / B) % KG), t = (int)(p % B);
        const int64_t i = tile * B + t;
        int64_t c;
        if (coalesced) c = i + k * 16;
        else c = i - W + (int64_t)k * w + (int64_t)(mix64(mix64(i + 0x9E3779B97F4A7C15ULL) ^ (uint64_t)k) % (uint64_t)w);
        if (c < 0) c += n;
        if (c >= n) c -= n;
        col[p] = (int32_t)c;
    }
}
__global__ void k_fillx(double *x, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        x[i] = (double)(mix64(i) >> 11) * (1.0 / 9007199254740992.0);
}

__device__ const double *g_val;     // mode 12: value stream, same layout as col
__device__ double *g_y;             // mode 12: per-row output
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int MODE>
__device__ __forceinline__ double gld(const double *p, uint64_t pol) {
    double r;
    if (MODE == 0) asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(r) : "l"(p));
    else if (MODE == 1) asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(r) : "l"(p));
    else if (MODE == 2) asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(r) : "l"(p));
    else asm volatile("ld.global.nc.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(r) : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ double gld_if(const double *p, bool on) {
    double r = 0.0;
    asm volatile("{\n.reg .pred q;\nsetp.ne.s32 q, %2, 0;\n@q ld.global.nc.f64 %0, [%1];\n}\n" : "+d"(r) : "l"(p), "r"((int)on));
    return r;
}

// UN rows per lane per iteration (UN * KG gathers in flight per lane)
template <int MODE, int UN>
__global__ void __launch_bounds__(B) k_gather(const int32_t *__restrict__ col, const double *__restrict__ x, double *out,
                                                int64_t ntiles, long long *cycles, const CUtensorMap *tmap_ptr, cudaTextureObject_t tex) {
    extern __shared__ __align__(128) double s_buf[];     // mode 8: 128 B per gather4; mode 9: 8 B per gather
    __shared__ uint64_t s_bar[8];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    uint64_t pol = 0;
    if (MODE == 3) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    if (MODE == 8) {
        if (lane == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&s_bar[warp])), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        __syncwarp();
    }
    uint32_t phase = 0;
    double acc = 0.0;
    const long long c0 = clock64();
    for (int64_t tile = (int64_t)blockIdx.x * UN; tile < ntiles; tile += (int64_t)gridDim.x * UN) {
        int32_t ci[UN][KG];
#pragma unroll
        for (int u = 0; u < UN; ++u)
#pragma unroll
            for (int k = 0; k < KG; ++k) ci[u][k] = (tile + u < ntiles) ? __ldg(col + ((tile + u) * KG + k) * B + t) : 0;
        double v[UN][KG];
        if (MODE == 12) {
            double a[UN][KG];
#pragma unroll
            for (int u = 0; u < UN; ++u)
#pragma unroll
                for (int k = 0; k < KG; ++k) {
                    const double *vp = g_val + ((tile + u) * KG + k) * B + t;
                    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(a[u][k]) : "l"(tile + u < ntiles ? vp : g_val));
                }
#pragma unroll
            for (int u = 0; u < UN; ++u)
#pragma unroll
                for (int k = 0; k < KG; ++k) v[u][k] = gld<0>(x + ci[u][k], pol);
#pragma unroll
            for (int u = 0; u < UN; ++u) {
                double sacc = 0.0;
#pragma unroll
                for (int k = 0; k < KG; ++k) { sacc = sacc + a[u][k] * v[u][k]; v[u][k] = 0.0; }
                if (tile + u < ntiles) g_y[(tile + u) * B + t] = sacc;
                v[u][0] = sacc;
            }
        } else if (MODE <= 3 || MODE == 10) {
#pragma unroll
            for (int u = 0; u < UN; ++u)
#pragma unroll
                for (int k = 0; k < KG; ++k) v[u][k] = gld<MODE == 10 ? 0 : MODE>(x + ci[u][k], pol);
        } else if (MODE >= 4 && MODE <= 7) {
            constexpr int AL = MODE == 4 ? 2 : (MODE == 5 ? 1 : (MODE == 6 ? 4 : 8));   // active lanes per instruction
#pragma unroll
            for (int u = 0; u < UN; ++u)
#pragma unroll
                for (int k = 0; k < KG; ++k) {
                    double r = 0.0;
#pragma unroll
                    for (int g = 0; g < 32 / AL; ++g) {
                        const double q = gld_if(x + ci[u][k], (lane / AL) == g);
                        r = (lane / AL) == g ? q : r;
                    }
                    v[u][k] = r;
                }
        } else if (MODE == 8) {
            // every lane issues its own gather4s into its private 16-byte slots; one mbarrier per warp
            double *mine = s_buf + (size_t)(warp * 32 + lane) * (KG * UN * 4);   // 128-byte aligned slots
            if (lane == 0)
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&s_bar[warp])),
                             "r"(32u * KG * UN * 16u) : "memory");
            __syncwarp();
#pragma unroll
            for (int u = 0; u < UN; ++u)
#pragma unroll
                for (int k = 0; k < KG; k += 4) {
                    asm volatile(
                        "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes"
                        " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                        ::"r"(smem_u32(mine + (u * KG + k) * 4)), "l"(tmap_ptr), "r"(smem_u32(&s_bar[warp])),
                          "r"(0), "r"(ci[u][k] >> 1), "r"(ci[u][k + 1] >> 1), "r"(ci[u][k + 2] >> 1), "r"(ci[u][k + 3] >> 1)
                        : "memory");
                }
            asm volatile(
                "{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}\n"
                ::"r"(smem_u32(&s_bar[warp])), "r"(phase) : "memory");
            phase ^= 1;
#pragma unroll
            for (int u = 0; u < UN; ++u)
#pragma unroll
                for (int k = 0; k < KG; ++k) v[u][k] = mine[(u * KG + (k & ~3)) * 4 + (k & 3) * 2 + (ci[u][k] & 1)];
            __syncwarp();
        } else if (MODE == 11) {
#pragma unroll
            for (int u = 0; u < UN; ++u)
#pragma unroll
                for (int k = 0; k < KG; ++k) {
                    const int2 q = tex1Dfetch<int2>(tex, ci[u][k]);
                    v[u][k] = __hiloint2double(q.y, q.x);
                }
        } else if (MODE == 9) {
            double *mine = s_buf + (size_t)t;       // [slot][B] layout: conflict-free
#pragma unroll
            for (int u = 0; u < UN; ++u)
#pragma unroll
                for (int k = 0; k < KG; ++k)
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(mine + (u * KG + k) * B)), "l"(x + ci[u][k]) : "memory");
            asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
#pragma unroll
            for (int u = 0; u < UN; ++u)
#pragma unroll
                for (int k = 0; k < KG; ++k) v[u][k] = mine[(u * KG + k) * B];
        }
#pragma unroll
        for (int u = 0; u < UN; ++u)
#pragma unroll
            for (int k = 0; k < KG; ++k) acc += v[u][k];
    }
    const long long c1 = clock64();
    out[(int64_t)blockIdx.x * B + t] = acc;
    if (blockIdx.x == 0 && t == 0) *cycles = c1 - c0;
}

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                             const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int MODE, int UN>
static void run(const char *name, int occ, const int32_t *col, const double *x, double *out, int64_t ntiles, long long *d_cyc,
                const CUtensorMap *d_tmap, int sms, double *checksum_ref, cudaTextureObject_t tex = 0) {
    const int grid = sms * occ;
    const size_t smem = MODE == 8 ? (size_t)B * KG * UN * 32 : (MODE == 9 ? (size_t)B * KG * UN * 8 : 0);
    CK(cudaFuncSetAttribute(k_gather<MODE, UN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 2; ++i) k_gather<MODE, UN><<<grid, B, smem>>>(col, x, out, ntiles, d_cyc, d_tmap, tex);
    CK(cudaDeviceSynchronize());
    const int reps = 5;
    CK(cudaEventRecord(e0));
    for (int i = 0; i < reps; ++i) k_gather<MODE, UN><<<grid, B, smem>>>(col, x, out, ntiles, d_cyc, d_tmap, tex);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    ms /= reps;
    long long cyc = 0;
    CK(cudaMemcpy(&cyc, d_cyc, sizeof(cyc), cudaMemcpyDeviceToHost));
    std::vector<double> h((size_t)grid * B);
    CK(cudaMemcpy(h.data(), out, sizeof(double) * h.size(), cudaMemcpyDeviceToHost));
    double s = 0;
    for (double v : h) s += v;
    const double gathers = (double)ntiles * B * KG;
    printf("%-44s UN=%d occ=%d  %8.3f ms  %7.1f G gathers/s  %6.3f gathers/clk/SM  (block0 %lld clk, %.0f MHz eff)  sum=%.6e%s\n",
           name, UN, occ, ms, gathers / ms * 1e-6, gathers / ((double)cyc * sms), cyc, (double)cyc / ms * 1e-3, s,
           (*checksum_ref != 0.0 && MODE != 10 && MODE != 12 && fabs(s - *checksum_ref) > 1e-6 * fabs(*checksum_ref)) ? "  CHECKSUM MISMATCH" : "");
    if (*checksum_ref == 0.0 && MODE != 10) *checksum_ref = s;
    fflush(stdout);
}

int main(int argc, char **argv) {
    const int64_t nrows = (int64_t)1 << (argc > 2 ? atoi(argv[2]) : 24), n = nrows, W = (int64_t)1 << 20;
    const int64_t ntiles = nrows / B;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    printf("device %s, %d SMs; %lld rows x %d gathers, band half-width %lld (window %.1f MB)\n", prop.name, sms, (long long)nrows, KG,
           (long long)W, 2.0 * W * 8 / 1e6);
    int32_t *col; double *x, *out; long long *d_cyc;
    CK(cudaMalloc(&col, sizeof(int32_t) * nrows * KG));
    CK(cudaMalloc(&x, sizeof(double) * (n + 16)));
    CK(cudaMalloc(&out, sizeof(double) * (size_t)sms * 8 * B));
    CK(cudaMalloc(&d_cyc, sizeof(long long)));
    k_fillx<<<sms * 8, 256>>>(x, n + 16);
    // TMA descriptor: x as [n/2 rows][2 doubles]
    CUtensorMap tmap, *d_tmap = nullptr;
    bool have_tma = false;
    {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess && fn) {
            cuuint64_t dims[2] = {2, (cuuint64_t)(n / 2)};
            cuuint64_t strides[1] = {16};
            cuuint32_t box[2] = {2, 1};
            cuuint32_t estr[2] = {1, 1};
            CUresult r = ((EncodeFn)fn)(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, x, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r == CUDA_SUCCESS) {
                CK(cudaMalloc(&d_tmap, sizeof(tmap)));
                CK(cudaMemcpy(d_tmap, &tmap, sizeof(tmap), cudaMemcpyHostToDevice));
                have_tma = true;
            } else printf("cuTensorMapEncodeTiled failed: %d\n", (int)r);
        }
    }
    double *val = nullptr, *y = nullptr;
    CK(cudaMalloc(&val, sizeof(double) * nrows * KG));
    CK(cudaMalloc(&y, sizeof(double) * nrows));
    CK(cudaMemset(val, 0, sizeof(double) * nrows * KG));
    k_fillx<<<sms * 8, 256>>>(val, nrows * KG);
    CK(cudaMemcpyToSymbol(g_val, &val, sizeof(val)));
    CK(cudaMemcpyToSymbol(g_y, &y, sizeof(y)));
    double ref = 0.0, refc = 1.0, ref12 = 0.0;
    k_fill<<<sms * 8, 256>>>(col, nrows, n, W, 0);
    CK(cudaDeviceSynchronize());
    const int only = argc > 1 ? atoi(argv[1]) : -1;
    cudaTextureObject_t tex = 0;
    {
        cudaResourceDesc rd = {};
        rd.resType = cudaResourceTypeLinear;
        rd.res.linear.devPtr = x;
        rd.res.linear.desc = cudaCreateChannelDesc<int2>();
        rd.res.linear.sizeInBytes = sizeof(double) * (size_t)n;
        cudaTextureDesc td = {};
        td.readMode = cudaReadModeElementType;
        cudaError_t e = cudaCreateTextureObject(&tex, &rd, &td, nullptr);
        if (e != cudaSuccess) { printf("texture object: %s\n", cudaGetErrorString(e)); tex = 0; cudaGetLastError(); }
        int maxw = 0;
        cudaDeviceGetAttribute(&maxw, cudaDevAttrMaxTexture1DLinearWidth, 0);
        printf("max 1D linear texture width: %d texels\n", maxw);
    }
#define RUN(M, U, O, NAME) if (only < 0 || only == M) run<M, U>(NAME, O, col, x, out, ntiles, d_cyc, d_tmap, sms, &ref, tex)
    RUN(0, 1, 4, "ld.global.nc");
    RUN(0, 2, 4, "ld.global.nc");
    RUN(0, 2, 8, "ld.global.nc");
    RUN(0, 4, 4, "ld.global.nc");
    RUN(1, 2, 4, "ld.global.nc.L1::no_allocate");
    RUN(2, 2, 4, "ld.global.cg");
    RUN(3, 2, 4, "ld.global.nc.L2::cache_hint evict_last");
    RUN(3, 2, 8, "ld.global.nc.L2::cache_hint evict_last");
    RUN(4, 1, 4, "predicated, 2 lanes / instruction");
    RUN(4, 2, 8, "predicated, 2 lanes / instruction");
    RUN(5, 1, 4, "predicated, 1 lane / instruction");
    RUN(6, 1, 4, "predicated, 4 lanes / instruction");
    RUN(6, 2, 8, "predicated, 4 lanes / instruction");
    RUN(7, 1, 4, "predicated, 8 lanes / instruction");
    RUN(7, 2, 8, "predicated, 8 lanes / instruction");
    if (have_tma) {
        RUN(8, 1, 2, "TMA tile::gather4 (16-byte rows)");
        RUN(8, 1, 3, "TMA tile::gather4 (16-byte rows)");
    }
    if (tex) {
        RUN(11, 1, 4, "tex1Dfetch<int2> (TEX pipe)");
        RUN(11, 2, 4, "tex1Dfetch<int2> (TEX pipe)");
        RUN(11, 2, 8, "tex1Dfetch<int2> (TEX pipe)");
    }
#define RUN12(U, O) if (only < 0 || only == 12) run<12, U>("SpMV prototype: sliced-ELL, direct coalesced val/col + gathers", O, col, x, out, ntiles, d_cyc, d_tmap, sms, &ref12, tex)
    RUN12(2, 2); RUN12(4, 2); RUN12(1, 2); RUN12(1, 4); RUN12(1, 6); RUN12(1, 8); RUN12(2, 3); RUN12(2, 4); RUN12(2, 6);
    RUN(9, 1, 4, "cp.async 8 B (LDGSTS)");
    RUN(9, 2, 2, "cp.async 8 B (LDGSTS)");
    // harness ceiling: the same kernel with coalesced columns
    k_fill<<<sms * 8, 256>>>(col, nrows, n, W, 1);
    CK(cudaDeviceSynchronize());
    if (only == 13) {      // SpMV prototype on coalesced columns (the ten-diagonal matrix): is the direct-load design HBM-bound there?
        double r13 = 0.0;
        const double bytes = (double)nrows * (KG * 12.0 + 8.0 + 8.0);
        printf("algorithmic bytes per launch %.3f GB (val+col %d entries, y, x once)\n", bytes * 1e-9, KG);
        run<12, 2>("SpMV prototype, coalesced columns", 2, col, x, out, ntiles, d_cyc, d_tmap, sms, &r13, tex);
        run<12, 4>("SpMV prototype, coalesced columns", 2, col, x, out, ntiles, d_cyc, d_tmap, sms, &r13, tex);
        run<12, 1>("SpMV prototype, coalesced columns", 4, col, x, out, ntiles, d_cyc, d_tmap, sms, &r13, tex);
        run<12, 1>("SpMV prototype, coalesced columns", 6, col, x, out, ntiles, d_cyc, d_tmap, sms, &r13, tex);
        run<12, 1>("SpMV prototype, coalesced columns", 8, col, x, out, ntiles, d_cyc, d_tmap, sms, &r13, tex);
        run<12, 2>("SpMV prototype, coalesced columns", 4, col, x, out, ntiles, d_cyc, d_tmap, sms, &r13, tex);
        run<12, 2>("SpMV prototype, coalesced columns", 6, col, x, out, ntiles, d_cyc, d_tmap, sms, &r13, tex);
        run<12, 4>("SpMV prototype, coalesced columns", 3, col, x, out, ntiles, d_cyc, d_tmap, sms, &r13, tex);
    }
    if (only < 0 || only == 10) run<10, 2>("coalesced columns (harness ceiling)", 4, col, x, out, ntiles, d_cyc, d_tmap, sms, &refc, tex);
    return 0;
}
