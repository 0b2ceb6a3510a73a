// kernels_gemm.cuh -- the Legendre stage as a batched FP64 tensor-core contraction.
//
// Replaces the triple loops of shtransforms.f90:152-221 (synthesis) and :692-726, :797-862 (analysis).
// For every azimuthal order m and equatorial parity the Legendre sums are a dense product
//        C[M x N] = A[M x K] * B[K x N]
// with A a slice of the precomputed P/D table, B the (level x field x re/im) batch.  The FP64 pipe is the
// binding roofline; B200 executes FP64 MMA as DMMA.8x8x4 (measured 37.0 TFLOP/s vs 33.8 for
// DFMA, profiles/fp64_peak_r01.json), there is no tcgen05 kind for f64, so the tile engine is
// mma.sync.m8n8k4.f64 fed from shared memory by a multi-stage cp.async pipeline whose stages are handed over with
// mbarriers (full: the cp.async of all threads have landed; empty: all warps have consumed the stage).  K-tile kt+STAGES-2
// is loaded at iteration kt into the stage k-tile kt-2 used, so a warp may run up to two k-tiles ahead of the slowest one
// instead of meeting all others at a CTA barrier every k-tile (measured at l_max=1023, 16 levels: synthesis 16.18 -> 15.10 ms,
// analysis 12.04 -> 11.76 ms; bit-identical results).
//
//   CTA tile 128 x 64, 8 warps as 4(M) x 2(N), warp tile 32 x 32 = 4x4 DMMA tiles, k-tile 16.
//   A_KCONTIG=false (synthesis): A tile stored As[k][m] (m = colatitude contiguous in the table).
//   A_KCONTIG=true  (analysis) : A tile stored As[m][k] (m = degree row, k = colatitude contiguous).
//   The [k][m] / [k][n] tiles have leading dimensions == 4 (mod 16) doubles, the [m][k] tile an XOR swizzle (below), so the
//   16 lanes of each half-warp of an LDS.64 fragment load hit 16 distinct 8-byte banks.
//   Each K segment of a CTA tile starts at the first k-tile that holds a non-negligible table entry for any of its rows
//   (triangular polar skipping, GemmProb::ks0/ks1).
#pragma once
#include "common.cuh"

namespace magic {

constexpr int G_THREADS = 256;
constexpr int LDA_M = GEMM_BM + 4;  // As[k][m]
// analysis A tile As[m][k]: padded rows (BK+4) or, with MAGIC_GEMM_SWZ, dense rows whose 4-double groups are XOR-swizzled
// by (m & 3): the same conflict-free LDS.64 fragment loads in 20 % less shared memory, which buys a 4th pipeline stage
// (measured 11.68 -> 11.55 ms per 16-level chunk at l_max=1023).  Not worth it / measured worse and removed again: 64x32 warp
// tiles with 4 warps, interleaved fragment ownership for ragged tiles, register double-buffering of the fragments.
#ifndef MAGIC_GEMM_SWZ
#define MAGIC_GEMM_SWZ 1
#endif
constexpr int LDA_K = MAGIC_GEMM_SWZ ? BK : BK + 4;  // As[m][k]
constexpr int LDB_S = GEMM_BN + 4;  // Bs[k][n]
constexpr int A_TILE_M = BK * LDA_M;
constexpr int A_TILE_K = GEMM_BM * LDA_K;
constexpr int B_TILE = BK * LDB_S;
constexpr int STAGES_M = 4;
constexpr int STAGES_K = MAGIC_GEMM_SWZ ? 4 : 3;

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}

// ---- mbarrier helpers ------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.shared::cta.b64 st, [%0];\n}\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
// arrival that fires when all prior cp.async of this thread have landed (does not raise the pending count)
__device__ __forceinline__ void mbar_cp_async_arrive(uint64_t *bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, int parity) {
    const unsigned addr = (unsigned)__cvta_generic_to_shared(bar);
    unsigned ok = 0;
    while (!ok) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(addr), "r"(parity)
            : "memory");
    }
}

__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

template <bool A_KCONTIG>
__global__ void __launch_bounds__(G_THREADS, 2)
legendre_gemm_kernel(const GemmProb *__restrict__ probs, const int2 *__restrict__ tiles, int lda) {
    extern __shared__ __align__(16) double smem[];
    constexpr int STAGES = A_KCONTIG ? STAGES_K : STAGES_M;
    constexpr int A_TILE = A_KCONTIG ? A_TILE_K : A_TILE_M;
    constexpr int STAGE = A_TILE + B_TILE;

    const int2 tile = tiles[blockIdx.x];
    const GemmProb pr = probs[tile.x];
    const int m0 = (tile.y >> 16) * GEMM_BM, n0 = (tile.y & 0xffff) * GEMM_BN;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int wm = (warp >> 1) * 32, wn = (warp & 1) * 32;
    // valid 8-row / 8-column fragments of this warp (ragged M of the analysis, padded N): the rest is skipped
    const int mfr = min(4, max(0, (pr.M - m0 - wm + 7) >> 3)), nfr = min(4, max(0, (pr.Nvalid - n0 - wn + 7) >> 3));
    const int mlo = min(4, max(0, (pr.Mlo - m0 - wm) >> 3));  // polar skipping: fragments [0, mlo) are all-negligible rows
    const bool active = mfr > mlo && nfr > 0;

    if (!A_KCONTIG && m0 + GEMM_BM <= pr.Mlo) {
        // whole tile lies in the negligible polar cap: its rows of F are exact zeros
        for (int idx = tid; idx < GEMM_BM * (GEMM_BN / 2); idx += G_THREADS) {
            int row = m0 + idx / (GEMM_BN / 2), c2 = idx % (GEMM_BN / 2);
            if (row < pr.M) *reinterpret_cast<double2 *>(pr.C + (size_t)row * pr.ldc + n0 + 2 * c2) = make_double2(0.0, 0.0);
        }
        return;
    }

    // ---- k range.  Segment s covers absolute k-tiles [st_s, pr.kt_s): st_s = first tile that holds a non-negligible table
    //      entry for any of the 16 fragments of this CTA tile (the table of order m is negligible polewards of the turning
    //      point sin(theta) ~ m/l: a triangle in the (degree, colatitude) plane, of which pr.Mlo / pr.klo only remove the
    //      rectangle common to all degrees).
    int st0 = pr.klo, st1 = pr.klo;
    if (pr.ks0 != nullptr) {
        auto min16 = [](const unsigned char *p) {
            const uint4 q = *reinterpret_cast<const uint4 *>(p);
            const unsigned mn = __vminu4(__vminu4(q.x, q.y), __vminu4(q.z, q.w));
            return (int)min(min(mn & 255u, (mn >> 8) & 255u), min((mn >> 16) & 255u, mn >> 24));
        };
        st0 = max(st0, min(min16(pr.ks0 + (m0 >> 3)), pr.kt0));
        if (pr.ks1 != nullptr) st1 = max(st1, min16(pr.ks1 + (m0 >> 3)));
    }
    st1 = min(st1, pr.kt1);
    const int n0t = pr.kt0 - st0, KT = n0t + (pr.kt1 - st1);
    const int kb1 = pr.kt0;  // B row tile of segment 1, tile 0 (the segments are stacked in B)

    auto load_stage = [&](int kt, int st) {
        double *As = smem + st * STAGE, *Bs = As + A_TILE;
        const bool seg1 = kt >= n0t;
        const int ka = seg1 ? st1 + (kt - n0t) : st0 + kt;  // absolute k-tile inside the segment
        const double *Ab = (seg1 ? pr.A1 : pr.A0) + (size_t)ka * BK * (A_KCONTIG ? 1 : lda);
        if (!A_KCONTIG) {
#pragma unroll
            for (int c = 0; c < 4; c++) {  // 16 rows x 64 chunks of 2 doubles
                int idx = tid + c * G_THREADS, k = idx >> 6, mc = idx & 63;
                cp_async16(As + k * LDA_M + mc * 2, Ab + (size_t)k * lda + m0 + mc * 2);
            }
        } else {
#pragma unroll
            for (int c = 0; c < 4; c++) {  // 128 rows x 8 chunks
                int idx = tid + c * G_THREADS, m = idx >> 3, kc = idx & 7;
                cp_async16(As + m * LDA_K + (MAGIC_GEMM_SWZ ? kc ^ ((m & 3) << 1) : kc) * 2, Ab + (size_t)(m0 + m) * lda + kc * 2);
            }
        }
        const int kb = seg1 ? kb1 + ka : ka;  // B row tile
        const double *Bb = pr.B + (size_t)kb * BK * pr.ldb + n0;
#pragma unroll
        for (int c = 0; c < 2; c++) {  // 16 rows x 32 chunks
            int idx = tid + c * G_THREADS, k = idx >> 5, nc = idx & 31;
            cp_async16(Bs + k * LDB_S + nc * 2, Bb + (size_t)k * pr.ldb + nc * 2);
        }
    };

    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

    constexpr int DIST = STAGES - 2;  // prefetch distance; the remaining stage is the slack between fastest and slowest warp
    __shared__ uint64_t bar_full[STAGES], bar_empty[STAGES];
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; s++) {
            mbar_init(&bar_full[s], G_THREADS);        // one cp.async arrival per thread
            mbar_init(&bar_empty[s], G_THREADS / 32);  // one arrival per warp
        }
    }
    __syncthreads();
#pragma unroll
    for (int s = 0; s < DIST; s++)
        if (s < KT) {
            load_stage(s, s);
            mbar_cp_async_arrive(&bar_full[s]);
        }
    for (int kt = 0; kt < KT; kt++) {
        {
            const int nk = kt + DIST;
            if (nk < KT) {
                const int st = nk % STAGES, use = nk / STAGES;
                if (use > 0) mbar_wait(&bar_empty[st], (use - 1) & 1);  // all warps are done with k-tile nk - STAGES
                load_stage(nk, st);
                mbar_cp_async_arrive(&bar_full[st]);
            }
        }
        mbar_wait(&bar_full[kt % STAGES], (kt / STAGES) & 1);
        const double *As = smem + (kt % STAGES) * STAGE, *Bs = As + A_TILE;
        if (active) {
            auto load_frags = [&](int kk, double *a, double *b) {
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    int row = wm + i * 8 + g;
                    a[i] = A_KCONTIG ? As[row * LDA_K + (MAGIC_GEMM_SWZ ? ((kk * 4 + t) ^ ((g & 3) << 2)) : kk * 4 + t)]
                                     : As[(kk * 4 + t) * LDA_M + row];
                }
#pragma unroll
                for (int j = 0; j < 4; j++) b[j] = Bs[(kk * 4 + t) * LDB_S + wn + j * 8 + g];
            };
            auto mma_frags = [&](const double *a, const double *b) {
                if (mfr == 4 && nfr == 4 && mlo == 0) {
#pragma unroll
                    for (int i = 0; i < 4; i++)
#pragma unroll
                        for (int j = 0; j < 4; j++) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
                } else {
#pragma unroll
                    for (int i = 0; i < 4; i++)
#pragma unroll
                        for (int j = 0; j < 4; j++)
                            if (i >= mlo && i < mfr && j < nfr) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
                }
            };
#pragma unroll
            for (int kk = 0; kk < BK / 4; kk++) {
                double a[4], b[4];
                load_frags(kk, a, b);
                mma_frags(a, b);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_empty[kt % STAGES]);
    }

#pragma unroll
    for (int i = 0; i < 4; i++) {
        int row = m0 + wm + i * 8 + g;
        if (row < pr.M) {
            double *crow = pr.C + (size_t)row * pr.ldc + n0 + wn + 2 * t;
#pragma unroll
            for (int j = 0; j < 4; j++) *reinterpret_cast<double2 *>(crow + j * 8) = make_double2(acc[i][j][0], acc[i][j][1]);
        }
    }
}

inline size_t gemm_smem_bytes(bool a_kcontig) {
    return sizeof(double) * (a_kcontig ? STAGES_K * (A_TILE_K + B_TILE) : STAGES_M * (A_TILE_M + B_TILE));
}

inline cudaError_t gemm_setup_attributes() {
    cudaError_t e = cudaFuncSetAttribute(legendre_gemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)gemm_smem_bytes(false));
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(legendre_gemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)gemm_smem_bytes(true));
}

inline void launch_legendre_gemm(bool a_kcontig, const GemmProb *probs, const int2 *tiles, int ntiles, int lda,
                                 cudaStream_t st) {
    if (ntiles <= 0) return;
    if (a_kcontig)
        legendre_gemm_kernel<true><<<ntiles, G_THREADS, gemm_smem_bytes(true), st>>>(probs, tiles, lda);
    else
        legendre_gemm_kernel<false><<<ntiles, G_THREADS, gemm_smem_bytes(false), st>>>(probs, tiles, lda);
}

}  // namespace magic
