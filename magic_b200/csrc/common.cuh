// common.cuh -- shared structures of magic_b200 (device + host).  Layouts are described in DESIGN.md.
//
//   table   per order mc two row blocks [P_even][P_odd]; one row per degree l = m .. l_max+1, NHP doubles
//           per row (northern colatitudes k=0..nh-1, zero padded to a multiple of 16).  P = Plm of
//           shtransforms.f90:38-91 / plms.f90:14-189.  The reference's dPlm = sin(theta) dP/dtheta is by construction
//           l c(l+1) P(l+1) - (l+1) c(l) P(l-1) (plms.f90:117-187), so every sum against dPlm is a sum against P with
//           3-point-combined coefficients: the vector transforms need no D table and cost 2 (synthesis) or 2 (analysis)
//           scalar-equivalent passes instead of 4.
//   B       Legendre-GEMM right operand, [K_pad][N] row-major per problem (mc, parity), N a multiple of 64,
//           column n = (col*n_lev + lev)*2 + reim.
//   F       (theta,m)-space, per problem (mc, s) a [nh][N] matrix with the same columns as B.
//   grid    g[field][lev][s][k][phi]: phi fastest; s=0 holds the equatorially symmetric part E, s=1 the
//           antisymmetric part O of the pair (north row k, south row k): north = E+O, south = E-O.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define MAGIC_MAX_SRC 16

namespace magic {

constexpr int BK = 16;  // Legendre GEMM k-tile; all K extents are padded to multiples of BK
constexpr int GEMM_BM = 128;
constexpr int GEMM_BN = 64;
// CTA tile of the analysis (A_KCONTIG) form of the Legendre GEMM: 128 x 64 like the synthesis, or "wide" 64 x 128.  Its M dimension
// is the ragged number of degrees of one (order, parity) problem: with 128-row tiles only 80 % of the warp rows of all tiles hold
// degrees at l_max = 1023 (89 % with 64-row tiles), fewer at smaller truncations, and a warp without rows idles.  Measured
// (analysis GEMM, ms): l_max = 255 1.4 -> 1.3, l_max = 511 (32 levels) 7.2 -> 6.4, l_max = 1023 (32 levels) 13.35 -> 13.56: the wide
// tile is used below l_max = MAGIC_GEMM_WIDE_BELOW (the handle decides, engine.cu).
#ifndef MAGIC_GEMM_WIDE_BELOW
#define MAGIC_GEMM_WIDE_BELOW 768
#endif
constexpr int GEMM_BM_WIDE = 64, GEMM_BN_WIDE = 128;
inline int gemm_bm_an(bool wide) { return wide ? GEMM_BM_WIDE : GEMM_BM; }
inline int gemm_bn_an(bool wide) { return wide ? GEMM_BN_WIDE : GEMM_BN; }

struct GemmProb {
    const double *A0;  // the table operand (P block of one order and parity)
    const double *B;
    double *C;
    int kt0;           // k-tiles
    int M;             // valid rows of C
    int ldb, ldc;      // leading dimensions of B and C (doubles)
    int Nvalid;        // real (unpadded) column count: DMMA fragments beyond it are skipped
    int Nstore;        // columns >= Nstore are not stored (= ldc when C is padded to whole tiles)
    int Mlo;           // synthesis: rows (colatitudes) below Mlo hold only negligible table entries and are skipped
    int klo;           // analysis: leading k-tiles skipped for the same reason
    // Triangular polar skipping: ks0[f] = first k-tile that holds a non-negligible table entry for the 8-row fragment f of the
    // M dimension (255 = none).  Null = no skipping.  A CTA tile starts at the minimum over its 16 fragments.  (Skipping the
    // DMMAs per fragment as well was measured: the predicate state costs the 128-register kernel as much as it saves.)
    const unsigned char *ks0;
};

// factor applied to a spectral source when assembling synthesis operands (sht_native.f90 wrappers)
enum FType : int { F_NONE = 0, F_ONE = 1, F_DLH = 2, F_OR2DLH = 3, F_NEG = 4, F_IM = 5 };
struct Term { int src; int ftype; };
// which local levels a column is computed on (rIter.f90:466-622)
enum LMask : int { LM_ALL = 0, LM_VEL = 1, LM_VELBULK = 2, LM_DERIV = 3 };
struct ScalCol { Term t[2]; int lmask; int pad; };
struct VecPair { Term S[2]; Term T[2]; int lmask; int pad; };

struct LevelInfo {  // per local level, device resident
    int nR, lcut, nBc, lDeriv, nl_on, l_bound, cour_on, center;  // center: full-sphere r=0 level (v_center_sphere)
    int torque, pad2;  // torque: 1 = Lorentz torque on the inner core, 2 = on the mantle is taken at this level (rIter.f90:279-292)
    double r, or1, or2, or4, orho1, orho2, beta, rho0, otemp1, temp0, visc, lambda, epscProf, delxr2, delxh2;
};

struct FftPlan {
    int N, H, nfac;
    int fac[16];
    const double2 *tw;  // device: exp(+2 pi i k/N), k=0..N-1
};

// r2c destination: the FFT of a grid row (field,lev,s,k) is scaled by the quadrature weight (R_W) or by weight / sin^2 theta
// (R_WS, vector components: shtransforms.f90:787-794) and written into row k, column `col` of the analysis operand of the
// parity problem s (the symmetric part E = N+S meets the even-parity P rows, O = N-S the odd ones).
enum RType : int { R_NONE = 0, R_W = 1, R_WS = 2 };
struct R2cField { int col; int rtype; };


// ---- asynchronous-copy and mbarrier helpers (Legendre GEMM pipeline, FFT row prefetch) ---------------------------------------
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
// the incrementing form: pending count +1 now, -1 when the prior cp.async of this thread have landed (pair it with mbar_arrive)
__device__ __forceinline__ void mbar_cp_async_arrive_inc(uint64_t *bar) {
    asm volatile("cp.async.mbarrier.arrive.shared::cta.b64 [%0];\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
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

__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }
// makes the mbarrier initialisation visible to the asynchronous proxy (bulk copies complete on it)
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes) {
    asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n}\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes)
                 : "memory");
}
// TMA 1-D bulk copy global -> shared (bytes: multiple of 16, both addresses 16-byte aligned); completes on `bar`
__device__ __forceinline__ void bulk_g2s(void *smem, const void *gmem, unsigned bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                     (unsigned)__cvta_generic_to_shared(smem)),
                 "l"(gmem), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar))
                 : "memory");
}

}  // namespace magic
