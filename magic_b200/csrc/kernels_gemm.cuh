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
//   CTA tile 128 x 64, 8 warps as 4(M) x 2(N) (synthesis) or 64 x 128, 2(M) x 4(N) (analysis: ragged M, common.cuh), warp tile
//   32 x 32 = 4x4 DMMA tiles, k-tile 16.
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
constexpr int LDB_S = GEMM_BN + 4;       // Bs[k][n], 64-column tile
constexpr int LDB_W = GEMM_BN_WIDE + 4;  // Bs[k][n], 128-column tile of the wide analysis form (both == 4 mod 16)
constexpr int A_TILE_M = BK * LDA_M;
constexpr int A_TILE_K = GEMM_BM * LDA_K;
constexpr int A_TILE_KW = GEMM_BM_WIDE * LDA_K;
constexpr int B_TILE = BK * LDB_S;
constexpr int B_TILE_W = BK * LDB_W;
constexpr int STAGES_M = 4;
constexpr int STAGES_K = MAGIC_GEMM_SWZ ? 4 : 3;

__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

// A CTA works through tiles blockIdx.x, blockIdx.x + gridDim.x, ...; the k-tiles of all its tiles form ONE stream through the
// stage ring, so with a persistent grid (two CTAs per SM, MAGIC_GEMM_PERSIST=1) the first k-tiles of the next tile are loaded
// while the last ones of the current tile are multiplied.  MEASURED at l_max=1023, 32-level chunks: persistent 38.1 + 28.1 ms per
// 64 levels, one tile per CTA (gridDim.x = ntiles, the default) 36.1 + 26.7 ms, bit-identical: the hardware block scheduler
// balances the very unequal tiles (polar skipping, ragged edges) better than a static round-robin, and the descriptor chain of
// the next tile (tile -> problem -> skip table) stalls all warps of the persistent CTA at every tile change.
template <bool A_KCONTIG, bool WIDE = false>
__global__ void __launch_bounds__(G_THREADS, 2)
legendre_gemm_kernel(const GemmProb *__restrict__ probs, const int2 *__restrict__ tiles, int ntiles, int lda) {
    extern __shared__ __align__(16) double smem[];
    constexpr int STAGES = A_KCONTIG ? STAGES_K : STAGES_M;
    static_assert(A_KCONTIG || !WIDE, "the wide tile exists for the analysis form only");
    constexpr int A_TILE = A_KCONTIG ? (WIDE ? A_TILE_KW : A_TILE_K) : A_TILE_M;
    constexpr int BM = WIDE ? GEMM_BM_WIDE : GEMM_BM, BN = WIDE ? GEMM_BN_WIDE : GEMM_BN;  // CTA tile
    constexpr int LDB = WIDE ? LDB_W : LDB_S;
    constexpr int WN = BN / 32;                                                                  // warps along N (8 / WN along M)
    constexpr int STAGE = A_TILE + BK * LDB;
#ifndef MAGIC_GEMM_ARRIVE_INC
#define MAGIC_GEMM_ARRIVE_INC 0
#endif
#ifndef MAGIC_GEMM_SYNC_AFTER_FULL
#define MAGIC_GEMM_SYNC_AFTER_FULL 0
#endif
#ifndef MAGIC_GEMM_SLACK
#define MAGIC_GEMM_SLACK 2
#endif
    constexpr int DIST = STAGES - MAGIC_GEMM_SLACK;  // prefetch distance; with 2 the remaining stage is the slack between fastest and slowest warp

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int wm = (warp / WN) * 32, wn = (warp % WN) * 32;

    __shared__ uint64_t bar_full[STAGES], bar_empty[STAGES];
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; s++) {
            mbar_init(&bar_full[s], G_THREADS);        // one cp.async arrival per thread
            mbar_init(&bar_empty[s], G_THREADS / 32);  // one arrival per warp
        }
    }
    __syncthreads();

    // first k-tile of a CTA tile: the first one that holds a non-negligible table entry for any of its 16 row fragments (the
    // table of order m is negligible polewards of the turning point sin(theta) ~ m/l: a triangle in the (degree, colatitude)
    // plane, of which Mlo / klo only remove the rectangle common to all degrees); -1: the whole tile lies in the polar cap
    auto first_ktile = [&](const GemmProb &pr, int m0) -> int {
        if (!A_KCONTIG && m0 + BM <= pr.Mlo) return -1;
        int st0 = pr.klo;
        if (pr.ks0 != nullptr) {
            unsigned mn;  // byte-wise minimum over the BM / 8 row fragments of the tile
            if constexpr (BM == 128) {
                const uint4 q = *reinterpret_cast<const uint4 *>(pr.ks0 + (m0 >> 3));
                mn = __vminu4(__vminu4(q.x, q.y), __vminu4(q.z, q.w));
            } else {
                const uint2 q = *reinterpret_cast<const uint2 *>(pr.ks0 + (m0 >> 3));
                mn = __vminu4(q.x, q.y);
            }
            const int f = (int)min(min(mn & 255u, (mn >> 8) & 255u), min((mn >> 16) & 255u, mn >> 24));
            st0 = max(st0, min(f, pr.kt0));
        }
        return st0;
    };

    // ---- load cursor: k-tile lkt of lKT of tile lt goes into stage slot gl % STAGES
    int lt = blockIdx.x - gridDim.x, lkt = 0, lKT = 0, lldb = 0, gl = 0;
    const double *lA = nullptr, *lB = nullptr;
    auto load_next = [&]() {
        while (lkt >= lKT) {  // next tile with work
            lt += gridDim.x;
            if (lt >= ntiles) return;
            const int2 tile = tiles[lt];
            const GemmProb &pr = probs[tile.x];
            const int m0 = (tile.y >> 16) * BM, n0 = (tile.y & 0xffff) * BN;
            const int st0 = first_ktile(pr, m0);
            lkt = 0;
            lKT = st0 < 0 ? 0 : pr.kt0 - st0;
            if (lKT > 0) {
                lA = pr.A0 + (A_KCONTIG ? (size_t)m0 * lda + (size_t)st0 * BK : (size_t)st0 * BK * lda + m0);
                lB = pr.B + (size_t)st0 * BK * pr.ldb + n0;
                lldb = pr.ldb;
            }
        }
        const int st = gl % STAGES, use = gl / STAGES;
        if (use > 0) mbar_wait(&bar_empty[st], (use - 1) & 1);  // all warps are done with the k-tile that used this slot
        double *As = smem + st * STAGE, *Bs = As + A_TILE;
        if (!A_KCONTIG) {
            const double *Ab = lA + (size_t)lkt * BK * lda;
#pragma unroll
            for (int c = 0; c < 4; c++) {  // 16 rows x 64 chunks of 2 doubles
                int idx = tid + c * G_THREADS, k = idx >> 6, mc = idx & 63;
                cp_async16(As + k * LDA_M + mc * 2, Ab + (size_t)k * lda + mc * 2);
            }
        } else {
            const double *Ab = lA + (size_t)lkt * BK;
#pragma unroll
            for (int c = 0; c < BM * 8 / G_THREADS; c++) {  // BM rows x 8 chunks
                int idx = tid + c * G_THREADS, m = idx >> 3, kc = idx & 7;
                cp_async16(As + m * LDA_K + (MAGIC_GEMM_SWZ ? kc ^ ((m & 3) << 1) : kc) * 2, Ab + (size_t)m * lda + kc * 2);
            }
        }
        const double *Bb = lB + (size_t)lkt * BK * lldb;
#pragma unroll
        for (int c = 0; c < BK * (BN / 2) / G_THREADS; c++) {  // 16 rows x BN/2 chunks
            int idx = tid + c * G_THREADS, k = idx / (BN / 2), nc = idx % (BN / 2);
            cp_async16(Bs + k * LDB + nc * 2, Bb + (size_t)k * lldb + nc * 2);
        }
#if MAGIC_GEMM_ARRIVE_INC == 1
        mbar_cp_async_arrive_inc(&bar_full[st]);  // tracked copies hold the phase open (+1 / -1), the thread's own arrival counts
        mbar_arrive(&bar_full[st]);
#elif MAGIC_GEMM_ARRIVE_INC == 2   // diagnostic only (serialises the pipeline): copies completed before a plain arrival
        cp_async_commit();
        cp_async_wait_all();
        mbar_arrive(&bar_full[st]);
#else
        mbar_cp_async_arrive(&bar_full[st]);
#endif
        lkt++;
        gl++;
    };

#pragma unroll
    for (int s = 0; s < DIST; s++) load_next();

    int gc = 0;  // k-tiles consumed so far
    for (int ti = blockIdx.x; ti < ntiles; ti += gridDim.x) {
        const int2 tile = tiles[ti];
        const GemmProb &pr = probs[tile.x];
        const int m0 = (tile.y >> 16) * BM, n0 = (tile.y & 0xffff) * BN;
        const int M = pr.M, ldc = pr.ldc, Nstore = pr.Nstore;
        double *const C = pr.C;
        const int st0 = first_ktile(pr, m0);
        if (st0 < 0) {
            // whole tile lies in the negligible polar cap: its rows of F are exact zeros
            for (int idx = tid; idx < BM * (BN / 2); idx += G_THREADS) {
                int row = m0 + idx / (BN / 2), c2 = idx % (BN / 2);
                if (row < M && n0 + 2 * c2 < Nstore) *reinterpret_cast<double2 *>(C + (size_t)row * ldc + n0 + 2 * c2) = make_double2(0.0, 0.0);
            }
            continue;
        }
        const int KT = pr.kt0 - st0;
        // valid 8-row / 8-column fragments of this warp (ragged M of the analysis, padded N): the rest is skipped
        const int mfr = min(4, max(0, (M - m0 - wm + 7) >> 3)), nfr = min(4, max(0, (pr.Nvalid - n0 - wn + 7) >> 3));
        const int mlo = min(4, max(0, (pr.Mlo - m0 - wm) >> 3));  // polar skipping: fragments [0, mlo) are all-negligible rows
        const bool active = mfr > mlo && nfr > 0;
        const bool full = mfr == 4 && nfr == 4 && mlo == 0;

        double acc[4][4][2];
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

        for (int kt = 0; kt < KT; kt++, gc++) {
            load_next();  // k-tile gc + DIST of this CTA's stream (it may belong to the next tile)
            const int st = gc % STAGES;
            mbar_wait(&bar_full[st], (gc / STAGES) & 1);
#if MAGIC_GEMM_SYNC_AFTER_FULL   // diagnostic only
            __syncthreads();
#endif
            const double *As = smem + st * STAGE, *Bs = As + A_TILE;
            if (active) {
#pragma unroll
                for (int kk = 0; kk < BK / 4; kk++) {
                    double a[4], b[4];
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        int row = wm + i * 8 + g;
                        a[i] = A_KCONTIG ? As[row * LDA_K + (MAGIC_GEMM_SWZ ? ((kk * 4 + t) ^ ((g & 3) << 2)) : kk * 4 + t)]
                                         : As[(kk * 4 + t) * LDA_M + row];
                    }
#pragma unroll
                    for (int j = 0; j < 4; j++) b[j] = Bs[(kk * 4 + t) * LDB + wn + j * 8 + g];
                    if (full) {
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
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_empty[st]);
        }

#pragma unroll
        for (int i = 0; i < 4; i++) {
            int row = m0 + wm + i * 8 + g;
            if (row < M) {
                double *crow = C + (size_t)row * ldc + n0 + wn + 2 * t;
#pragma unroll
                for (int j = 0; j < 4; j++)
                    if (n0 + wn + 2 * t + j * 8 < Nstore) *reinterpret_cast<double2 *>(crow + j * 8) = make_double2(acc[i][j][0], acc[i][j][1]);
            }
        }
    }
}

inline size_t gemm_smem_bytes(bool a_kcontig, bool wide = false) {
    if (a_kcontig && wide) return sizeof(double) * STAGES_K * (A_TILE_KW + B_TILE_W);
    return sizeof(double) * (a_kcontig ? STAGES_K * (A_TILE_K + B_TILE) : STAGES_M * (A_TILE_M + B_TILE));
}

inline cudaError_t gemm_setup_attributes() {
    cudaError_t e = cudaFuncSetAttribute(legendre_gemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)gemm_smem_bytes(false));
    if (e != cudaSuccess) return e;
    cudaFuncSetAttribute(legendre_gemm_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    cudaFuncSetAttribute(legendre_gemm_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    cudaFuncSetAttribute(legendre_gemm_kernel<true, true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    e = cudaFuncSetAttribute(legendre_gemm_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)gemm_smem_bytes(true, true));
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(legendre_gemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)gemm_smem_bytes(true));
}

// grid: one CTA per tile (MAGIC_GEMM_PERSIST=1: two persistent CTAs per SM, see the kernel's header comment)
inline void launch_legendre_gemm(bool a_kcontig, const GemmProb *probs, const int2 *tiles, int ntiles, int lda,
                                 cudaStream_t st, bool wide = false) {
    if (ntiles <= 0) return;
    static int ctas = 0;
    if (ctas == 0) {
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const char *e = getenv("MAGIC_GEMM_PERSIST");
        ctas = (e && atoi(e) == 1) ? 2 * sms : (1 << 30);
    }
    const int grid = ntiles < ctas ? ntiles : ctas;
    if (a_kcontig && wide)
        legendre_gemm_kernel<true, true><<<grid, G_THREADS, gemm_smem_bytes(true, true), st>>>(probs, tiles, ntiles, lda);
    else if (a_kcontig)
        legendre_gemm_kernel<true><<<grid, G_THREADS, gemm_smem_bytes(true), st>>>(probs, tiles, ntiles, lda);
    else
        legendre_gemm_kernel<false><<<grid, G_THREADS, gemm_smem_bytes(false), st>>>(probs, tiles, ntiles, lda);
}

}  // namespace magic
