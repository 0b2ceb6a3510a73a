// kernels_fft.cuh -- batched real FFT along phi, hand written (no cuFFT on the path).
//
// Replaces fft.f90:164-252 (fft_many / ifft_many, Temperton FFT99) with the same transform definition
// (fft.f90:262-268): c2r  x_j = sum_k c_k e^{+2 pi i jk/n}, c_{n-k}=conj(c_k) (unnormalised),
//                    r2c  c_k = (1/n) sum_j x_j e^{-2 pi i jk/n}.
// A real transform of length N is done as a complex transform of length H=N/2 (Cooley-Lewis-Welch packing, as
// fft991 does) with a mixed-radix {8,4,2,3,5} Stockham autosort in shared memory.  One CTA owns R rows.
//   * the radix plan is a compile-time function of H, so all index arithmetic is strength reduced;
//   * each pass is done IN PLACE: a thread pulls all its butterflies into registers, the CTA synchronises, the
//     results go back to the autosort positions -- one buffer per row instead of two;
//   * rows are stored with one pad element every 8 (index i -> i + i/8) so the stride-RADIX writes of the first
//     passes do not serialise on shared-memory banks.
// The c2r kernel gathers its input from the (theta,m)-space matrices written by the Legendre GEMM (zero padding of
// orders mc >= n_m_max, shtransforms.f90:224-232, is implicit); the r2c kernel scatters its output, already weighted
// for the quadrature, into the analysis GEMM operands.  Sizes without a compiled plan use the generic kernels below.
#pragma once
#include "common.cuh"

namespace magic {

__device__ __forceinline__ double2 cmul(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 cconj(double2 a) { return make_double2(a.x, -a.y); }
__device__ __forceinline__ double2 cscale(double2 a, double s) { return make_double2(a.x * s, a.y * s); }
// multiply by sg*i
__device__ __forceinline__ double2 cmuli(double2 a, double sg) { return make_double2(-sg * a.y, sg * a.x); }
// twiddle e^{sg * 2 pi i idx / N}
// Cache hints (MAGIC_FFT_HINTS=0 turns them off for A/B runs): the twiddle table (H entries in use, 24 KB at n_phi = 3072) is
// the only data of these kernels that is re-read, but it shares L1 with the streaming rows -- ncu on the unhinted kernels:
// 27-37 % of the twiddle sectors missed L1 and every miss is an L2 round trip in the middle of a barrier-separated pass
// (long-scoreboard stalls on the DMULs that consume them: a quarter of all samples).  Twiddle loads ask for evict_last, the
// streaming loads and stores do not allocate in L1.
#ifndef MAGIC_FFT_HINTS
#define MAGIC_FFT_HINTS 1
#endif
#ifndef MAGIC_FFT_TWHOIST
#define MAGIC_FFT_TWHOIST 0
#endif
// measured at l_max = 1023 (ms per 16-level chunk): cache hints 4.51 -> 3.95 (c2r), 3.04 -> 2.60 (r2c); twiddle load hoisted above the
// pass barrier: 3.95 vs 3.94 (off); table copy in shared memory for c2r: 3.92 vs 3.91 (off); r2c post-processing twiddles requested
// before the last butterflies: 2.56 vs 2.60 (off: changes FMA contraction, i.e. result bits, for 1.5 %)
#ifndef MAGIC_FFT_C2R_STW
#define MAGIC_FFT_C2R_STW 0
#endif
#ifndef MAGIC_FFT_R2C_TWHOIST
#define MAGIC_FFT_R2C_TWHOIST 0
#endif
__device__ __forceinline__ double2 ld_keep(const double2 *p) {
#if MAGIC_FFT_HINTS
    double2 v;
    asm("ld.global.nc.L1::evict_last.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
#else
    return __ldg(p);
#endif
}
__device__ __forceinline__ double2 ld_stream(const double *p) {
#if MAGIC_FFT_HINTS
    double2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
#else
    return *reinterpret_cast<const double2 *>(p);
#endif
}
__device__ __forceinline__ void st_stream(double *p, double2 v) {
#if MAGIC_FFT_HINTS
    asm volatile("st.global.L1::no_allocate.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
#else
    *reinterpret_cast<double2 *>(p) = v;
#endif
}
// TWS: `tw` is a copy of the table in shared memory (the c2r prefetch kernel keeps one: its L1 is too small to hold the table
// next to two CTAs' row and staging buffers -- ncu: 53 % twiddle hit rate even with evict_last)
template <bool TWS = false>
__device__ __forceinline__ double2 twid(const double2 *__restrict__ tw, int idx, double sg) {
    double2 w = TWS ? tw[idx] : ld_keep(tw + idx);
    w.y *= sg;
    return w;
}

// ---- radix butterflies: o_k = sum_j a_j e^{sg 2 pi i jk/R} ------------------------------------------------------
template <int R>
__device__ __forceinline__ void butterfly(const double2 *a, double2 *o, double sg);

template <>
__device__ __forceinline__ void butterfly<2>(const double2 *a, double2 *o, double) {
    o[0] = cadd(a[0], a[1]);
    o[1] = csub(a[0], a[1]);
}
template <>
__device__ __forceinline__ void butterfly<4>(const double2 *a, double2 *o, double sg) {
    double2 t0 = cadd(a[0], a[2]), t1 = csub(a[0], a[2]), t2 = cadd(a[1], a[3]), t3 = cmuli(csub(a[1], a[3]), sg);
    o[0] = cadd(t0, t2);
    o[1] = cadd(t1, t3);
    o[2] = csub(t0, t2);
    o[3] = csub(t1, t3);
}
template <>
__device__ __forceinline__ void butterfly<3>(const double2 *a, double2 *o, double sg) {
    const double s3 = 0.86602540378443864676372317075294;
    double2 t = cadd(a[1], a[2]);
    double2 u = cmuli(cscale(csub(a[1], a[2]), s3), sg);
    double2 c = make_double2(a[0].x - 0.5 * t.x, a[0].y - 0.5 * t.y);
    o[0] = cadd(a[0], t);
    o[1] = cadd(c, u);
    o[2] = csub(c, u);
}
template <>
__device__ __forceinline__ void butterfly<5>(const double2 *a, double2 *o, double sg) {
    const double c1 = 0.30901699437494742410229341718282, c2 = -0.80901699437494742410229341718282;
    const double s1 = 0.95105651629515357211643933337938, s2 = 0.58778525229247312916870595463907;
    double2 t1 = cadd(a[1], a[4]), t2 = cadd(a[2], a[3]), t3 = csub(a[1], a[4]), t4 = csub(a[2], a[3]);
    o[0] = cadd(a[0], cadd(t1, t2));
    double2 m1 = make_double2(a[0].x + c1 * t1.x + c2 * t2.x, a[0].y + c1 * t1.y + c2 * t2.y);
    double2 m2 = make_double2(a[0].x + c2 * t1.x + c1 * t2.x, a[0].y + c2 * t1.y + c1 * t2.y);
    double2 n1 = cmuli(make_double2(s1 * t3.x + s2 * t4.x, s1 * t3.y + s2 * t4.y), sg);
    double2 n2 = cmuli(make_double2(s2 * t3.x - s1 * t4.x, s2 * t3.y - s1 * t4.y), sg);
    o[1] = cadd(m1, n1);
    o[4] = csub(m1, n1);
    o[2] = cadd(m2, n2);
    o[3] = csub(m2, n2);
}
template <>
__device__ __forceinline__ void butterfly<8>(const double2 *a, double2 *o, double sg) {
    const double h = 0.70710678118654752440084436210485;
    double2 t0 = cadd(a[0], a[4]), t4 = csub(a[0], a[4]), t1 = cadd(a[1], a[5]), t5 = csub(a[1], a[5]);
    double2 t2 = cadd(a[2], a[6]), t6 = csub(a[2], a[6]), t3 = cadd(a[3], a[7]), t7 = csub(a[3], a[7]);
    // odd branch inputs multiplied by w8^j, w8 = (1 + sg i)/sqrt2
    double2 u1 = make_double2(h * (t5.x - sg * t5.y), h * (t5.y + sg * t5.x));
    double2 u2 = cmuli(t6, sg);
    double2 u3 = make_double2(h * (-t7.x - sg * t7.y), h * (-t7.y + sg * t7.x));
    double2 e[4] = {t0, t1, t2, t3}, d[4] = {t4, u1, u2, u3}, ye[4], yd[4];
    butterfly<4>(e, ye, sg);
    butterfly<4>(d, yd, sg);
    o[0] = ye[0]; o[2] = ye[1]; o[4] = ye[2]; o[6] = ye[3];
    o[1] = yd[0]; o[3] = yd[1]; o[5] = yd[2]; o[7] = yd[3];
}

__host__ __device__ constexpr int fft_pick_radix(int len) {
    return len % 8 == 0 ? 8 : len % 4 == 0 ? 4 : len % 2 == 0 ? 2 : len % 3 == 0 ? 3 : len % 5 == 0 ? 5 : 1;
}
__host__ __device__ constexpr int fft_pad(int i) { return i + (i >> 3); }

// One in-place Stockham pass (compile-time radix RADIX, sub-length LEN, stride S) over R rows of length H.
template <int H, int R, int NT, int RADIX, int LEN, int S, int ROWLEN = fft_pad(H), bool TWS = false>
__device__ __forceinline__ void fft_pass(double2 *buf, const double2 *__restrict__ tw, double sg) {
    constexpr int M = LEN / RADIX, NB = H / RADIX, TOTAL = R * NB, PER = (TOTAL + NT - 1) / NT;
    constexpr int TWSTEP = 2 * H / LEN;
    double2 reg[PER][RADIX];
    double2 w1[PER];
#pragma unroll
    for (int u = 0; u < PER; u++) {
        int idx = threadIdx.x + u * NT;
        if (TOTAL % NT == 0 || idx < TOTAL) {
            int row = idx / NB, b = idx - row * NB;
            const double2 *x = buf + row * ROWLEN;
            if constexpr (NB % 8 == 0) {  // padded index is affine in j: pad(b + NB j) = pad(b) + j (NB + NB/8)
                const double2 *xb = x + fft_pad(b);
#pragma unroll
                for (int j = 0; j < RADIX; j++) reg[u][j] = xb[j * (NB + NB / 8)];
            } else {
#pragma unroll
                for (int j = 0; j < RADIX; j++) reg[u][j] = x[fft_pad(b + NB * j)];
            }
            // the pass's one twiddle load is issued before the barrier, so its latency hides behind the shared-memory loads
            // and the barrier wait instead of heading the dependent chain of the butterfly outputs
#if MAGIC_FFT_TWHOIST
            if (M > 1) w1[u] = twid<TWS>(tw, (b / S) * TWSTEP, sg);
#endif
        }
    }
    __syncthreads();
#pragma unroll
    for (int u = 0; u < PER; u++) {
        int idx = threadIdx.x + u * NT;
        if (TOTAL % NT == 0 || idx < TOTAL) {
            int row = idx / NB, b = idx - row * NB;
            int p = b / S, q = b - p * S;
            double2 o[RADIX];
            butterfly<RADIX>(reg[u], o, sg);
            double2 *y = buf + row * ROWLEN;
            // output position of o[k]: pad(q + S (RADIX p + k)) = y0 + k * ystride for the strides that occur
            int y0, ystride;
            if constexpr (S % 8 == 0) { y0 = fft_pad(q + S * RADIX * p); ystride = S + S / 8; }
            else if constexpr (S == 1 && RADIX == 8) { y0 = 9 * p; ystride = 1; }
            else if constexpr (S == 1 && RADIX == 4) { y0 = 4 * p + (p >> 1); ystride = 1; }
            else { y0 = 0; ystride = 0; }
            constexpr bool AFFINE = (S % 8 == 0) || (S == 1 && (RADIX == 8 || RADIX == 4));
            if (M > 1) {
                // twiddles w^k, w = e^{sg 2 pi i p/LEN}: one table load, the powers by a depth-3 product tree (the
                // kernel is bound by L1/shared wavefronts, not by the FP64 pipe; each product costs ~1 ulp)
                double2 w[RADIX];
#if MAGIC_FFT_TWHOIST
                w[1] = w1[u];
#else
                w[1] = twid<TWS>(tw, p * TWSTEP, sg);
#endif
                if (RADIX > 2) w[2] = cmul(w[1], w[1]);
                if (RADIX > 3) w[3] = cmul(w[2], w[1]);
                if (RADIX > 4) w[4] = cmul(w[2], w[2]);
                if (RADIX > 5) { w[5] = cmul(w[4], w[1]); w[6] = cmul(w[3], w[3]); w[7 < RADIX ? 7 : 0] = cmul(w[4], w[3]); }
#pragma unroll
                for (int k = 1; k < RADIX; k++) o[k] = cmul(o[k], w[k]);
            }
#pragma unroll
            for (int k = 0; k < RADIX; k++) {
                if constexpr (AFFINE) y[y0 + k * ystride] = o[k];
                else y[fft_pad(q + S * (RADIX * p + k))] = o[k];
            }
        }
    }
    __syncthreads();
}

// ------------------------------------------------------------------------------------------------------
// r2c: one CTA transforms R grid rows (consecutive levels of one field, fixed s,k) and scatters the weighted
// coefficients of orders mc < n_m into the analysis operands.  blockIdx.x = level chunk, .y = s*nh + k, .z = field.
struct R2cArgs {
    const double *grid;
    int n_lev, nh, n_m, NHP;
    const double *wgauss;  // 2 pi * gauss(k) / n_phi  (quadrature weight and forward-FFT normalisation)
    const double *osin2;   // 1/sin^2(theta_k)
    const R2cField *fields;
    double *B;             // analysis operands: per parity problem (mc, s) a [NHP][ldB] matrix
    int ldB;
    int minc;
};

__device__ __forceinline__ void r2c_scatter(const R2cArgs &a, const R2cField fd, int s, int k, int mc, int lev, double2 zk, double2 zmc,
                                            double2 w8, double w, double ws) {
    // X_k = E_k + e^{-2 pi i k/N} O_k, E=(Z_k+conj Z_{H-k})/2, O=-i (Z_k-conj Z_{H-k})/2
    if (fd.rtype == R_NONE) return;
    double2 zm = cconj(zmc);
    double2 e = cscale(cadd(zk, zm), 0.5), o = cmuli(cscale(csub(zk, zm), 0.5), -1.0);
    double2 x = cadd(e, cmul(w8, o));
    const double2 v = cscale(x, fd.rtype == R_W ? w : ws);
    const size_t off = ((size_t)(mc * 2 + s) * a.NHP + k) * a.ldB + 2 * ((size_t)fd.col * a.n_lev + lev);
    st_stream(a.B + off, v);
}

// ------------------------------------------------------------------------------------------------------
// Planned kernels.  The first Stockham pass runs on the registers of the global gather and the last pass feeds the
// global store, so a row crosses shared memory 2 (passes - 1) times instead of 2 passes + 2 (H = 1536 = 8.8.8.3: 6
// transfers per element instead of 10; measured per 16-level chunk at l_max=1023: c2r 5.03 -> 4.54 ms, r2c 4.31 -> 4.08 ms
// against the first-generation kernels that staged the gather in shared memory first).
//   c2r: the Hermitian pre-processing pairs order kk with H-kk; first-pass butterfly b (inputs kk = b + NB j) pairs with
//        butterfly NB-b, so one thread gathers the 2 R1 orders of both and produces both butterflies.
//   r2c: the same pairing at the other end: last-pass butterfly q (outputs mc = q + S k) pairs with butterfly S-q.
__host__ __device__ constexpr int fft_last_radix(int H) {
    int len = H;
    while (len / fft_pick_radix(len) > 1) len /= fft_pick_radix(len);
    return len;
}
// Rows per CTA: as many as give about MAGIC_FFT_CTA_THREADS threads (one thread per paired first-pass item), at most 16.
// Measured (l_max = 511, H = 768: 2 rows = 96 threads 6.1 + 5.6 ms per step, 4 rows = 192 threads 4.7 + 4.2 ms, 8 rows 8.4 + 5.6 ms;
// l_max = 255, H = 384: 4 / 8 / 16 rows 2.1 + 1.6 / 2.0 + 1.3 / 2.8 + 1.6 ms; l_max = 1023, H = 1536: 1 / 2 / 4 rows 6.7 + 7.5 /
// 3.9 + 2.6 / 4.7 + 2.8 ms per 16-level chunk): six warps per CTA are the sweet spot at every length from H = 256 up.
#ifndef MAGIC_FFT_CTA_THREADS
#define MAGIC_FFT_CTA_THREADS 192
#endif
__host__ __device__ constexpr int fft2_rows(int H) {
    if (H < 256) return H >= 96 ? 8 : 16;  // short rows: many small CTAs per SM do better (H = 144: 8 rows 0.3 ms, 16 rows 0.5 ms per step)
    const int ni1 = (H / fft_pick_radix(H) + 1) / 2;
    const int r = MAGIC_FFT_CTA_THREADS / (ni1 > 0 ? ni1 : 1);
    return r < 1 ? 1 : r > 16 ? 16 : r;
}
__host__ __device__ constexpr int fft2_rowlen(int H) { return fft_pad(H) + ((fft_pad(H) % 8) == 4 ? 0 : (12 - fft_pad(H) % 8) % 8); }  // == 4 (mod 8): rows r, r+1 on complementary banks
__host__ __device__ constexpr int fft2_threads(int H) {
    // one thread per paired first-pass item: R * ceil(NB1 / 2), rounded to warps, at most 256
    int t = ((fft2_rows(H) * ((H / fft_pick_radix(H) + 1) / 2) + 31) / 32) * 32;
    return t > 256 ? 256 : t;
}

// middle passes: everything between the first (LEN = H) and the last (LEN = fft_last_radix(H)) pass
template <int H, int R, int NT, int ROWLEN_, int LEN, int S, bool TWS = false>
struct FftMid {
    static constexpr int RADIX = fft_pick_radix(LEN);
    static __device__ __forceinline__ void run(double2 *buf, const double2 *__restrict__ tw, double sg) {
        if constexpr (LEN / RADIX > 1) {
            fft_pass<H, R, NT, RADIX, LEN, S, ROWLEN_, TWS>(buf, tw, sg);
            FftMid<H, R, NT, ROWLEN_, LEN / RADIX, S * RADIX, TWS>::run(buf, tw, sg);
        }
    }
};

// twiddle powers w^1..w^{RADIX-1} by the same product tree as fft_pass
template <int RADIX>
__device__ __forceinline__ void twiddle_powers(double2 *w) {
    if (RADIX > 2) w[2] = cmul(w[1], w[1]);
    if (RADIX > 3) w[3] = cmul(w[2], w[1]);
    if (RADIX > 4) w[4] = cmul(w[2], w[2]);
    if (RADIX > 5) { w[5] = cmul(w[4], w[1]); w[6] = cmul(w[3], w[3]); w[7 < RADIX ? 7 : 0] = cmul(w[4], w[3]); }
}

// experiment knobs: minimum resident CTAs per SM handed to __launch_bounds__ (undefined = leave the register choice to ptxas)
#ifdef MAGIC_FFT_MAXREG
#define MAGIC_FFT_LB_C(H) __maxnreg__(MAGIC_FFT_MAXREG)
#elif defined(MAGIC_FFT_MINB_C)
#define MAGIC_FFT_LB_C(H) __launch_bounds__(fft2_threads(H), MAGIC_FFT_MINB_C)
#else
#define MAGIC_FFT_LB_C(H) __launch_bounds__(fft2_threads(H))
#endif
#ifdef MAGIC_FFT_MAXREG
#define MAGIC_FFT_LB_R(H) __maxnreg__(MAGIC_FFT_MAXREG)
#elif defined(MAGIC_FFT_MINB_R)
#define MAGIC_FFT_LB_R(H) __launch_bounds__(fft2_threads(H), MAGIC_FFT_MINB_R)
#else
#define MAGIC_FFT_LB_R(H) __launch_bounds__(fft2_threads(H))
#endif
// first-pass butterfly b of a row (S = 1: p = b, outputs at 8b..8b+7 for radix 8) written to shared memory
template <int H, int R1, bool TWS = false>
__device__ __forceinline__ void first_pass_out(double2 *row, const double2 *__restrict__ tw, int b, const double2 *in, double sg) {
    double2 o[R1];
    butterfly<R1>(in, o, sg);
    if (H / R1 > 1) {
        double2 w[R1];
        w[1] = twid<TWS>(tw, b * 2, sg);  // TWSTEP = 2H / LEN = 2
        twiddle_powers<R1>(w);
#pragma unroll
        for (int k = 1; k < R1; k++) o[k] = cmul(o[k], w[k]);
    }
#pragma unroll
    for (int k = 0; k < R1; k++) row[fft_pad(R1 * b + k)] = o[k];
}

template <int H>
__global__ void MAGIC_FFT_LB_C(H) fft_c2r_plan_kernel(const double2 *__restrict__ tw, const double *__restrict__ F, int ld,
                                                                    int n_m, int nh, int ncols, const int *__restrict__ colrow,
                                                                    double *__restrict__ grid) {
    constexpr int R = fft2_rows(H), NT = fft2_threads(H), N = 2 * H, ROWLEN = fft2_rowlen(H);
    constexpr int R1 = fft_pick_radix(H), NB1 = H / R1, NI1 = (NB1 + 1) / 2;
    constexpr int RL = fft_last_radix(H), SL = H / RL;
    extern __shared__ __align__(16) double2 fsm[];
    const int cc0 = blockIdx.x * R;
    const int sk = blockIdx.y, s = sk / nh, k = sk - s * nh;
    const int rows = min(R, ncols - cc0);
    const double *Fb = F + ((size_t)s * nh + k) * ld + 2 * cc0;
    const size_t mstride = (size_t)2 * nh * ld;
    // destination rows of this tile, fetched now (the store phase used to start with a dependent global load: 5 % of all stall
    // samples sat on its first use)
    __shared__ int s_row[R];
    if (threadIdx.x < R) s_row[threadIdx.x] = threadIdx.x < rows ? colrow[cc0 + threadIdx.x] : -1;
    // ---- gather + Hermitian pre-processing + first pass.  Item t of a row: butterflies (t, NB1 - t); t = 0: butterfly 0 and,
    //      for even NB1, the self-paired butterfly NB1/2.
    for (int item = threadIdx.x; item < R * NI1; item += NT) {
        const int t = item / R, r = item - t * R;
        const bool has_v = (t > 0) || (NB1 % 2 == 0);
        const int u = t, v = (t > 0) ? NB1 - t : NB1 / 2;
        double2 A[R1], B[R1];
#pragma unroll
        for (int j = 0; j < R1; j++) {
            A[j] = make_double2(0.0, 0.0);
            B[j] = A[j];
            const int ka = u + NB1 * j, kb = v + NB1 * j;
            if (r < rows) {
                if (ka < n_m) A[j] = ld_stream(Fb + (size_t)ka * mstride + 2 * r);
                if (has_v && kb < n_m) B[j] = ld_stream(Fb + (size_t)kb * mstride + 2 * r);
            }
        }
        double2 *row = fsm + r * ROWLEN;
        double2 Yu[R1], Yv[R1];
        if (t > 0) {
            // order ka = u + NB1 j pairs with H - ka = v + NB1 (R1-1-j)
#pragma unroll
            for (int j = 0; j < R1; j++) {
                const double2 a = A[j], b = B[R1 - 1 - j];
                const double2 w = twid(tw, u + NB1 * j, 1.0);
                const double2 cb = cconj(b), ca = cconj(a);
                Yu[j] = cadd(cadd(a, cb), cmuli(cmul(w, csub(a, cb)), 1.0));
                const double2 w2 = make_double2(-w.x, w.y);  // e^{2 pi i (H-ka)/N} = -conj(w)
                Yv[R1 - 1 - j] = cadd(cadd(b, ca), cmuli(cmul(w2, csub(b, ca)), 1.0));
            }
        } else {
            // butterfly 0: order NB1 j pairs with NB1 (R1 - j) (order H is absent: c_H = 0, Im c_0 ignored, fft.f90:262-268)
#pragma unroll
            for (int j = 0; j < R1; j++) {
                double2 a = A[j];
                double2 b = (j == 0) ? make_double2(0.0, 0.0) : A[R1 - j];
                if (j == 0) a.y = 0.0;
                const double2 w = twid(tw, NB1 * j, 1.0);
                const double2 cb = cconj(b);
                Yu[j] = cadd(cadd(a, cb), cmuli(cmul(w, csub(a, cb)), 1.0));
            }
            // butterfly NB1/2: order v + NB1 j pairs with v + NB1 (R1-1-j)
#pragma unroll
            for (int j = 0; j < R1; j++) {
                const double2 a = B[j], b = B[R1 - 1 - j];
                const double2 w = twid(tw, v + NB1 * j, 1.0);
                const double2 cb = cconj(b);
                Yv[j] = cadd(cadd(a, cb), cmuli(cmul(w, csub(a, cb)), 1.0));
            }
        }
        first_pass_out<H, R1>(row, tw, u, Yu, 1.0);
        if (has_v) first_pass_out<H, R1>(row, tw, v, Yv, 1.0);
    }
    __syncthreads();
    FftMid<H, R, NT, ROWLEN, H / R1, R1>::run(fsm, tw, 1.0);
    // ---- last pass (LEN = RL, p = 0: no twiddles) straight to the grid row: z_j = x_{2j} + i x_{2j+1}
    for (int idx = threadIdx.x; idx < R * SL; idx += NT) {
        const int r = idx / SL, q = idx - r * SL;
        const double2 *x = fsm + r * ROWLEN;
        double2 in[RL], o[RL];
#pragma unroll
        for (int j = 0; j < RL; j++) in[j] = x[fft_pad(q + SL * j)];
        butterfly<RL>(in, o, 1.0);
        {
            const int row = s_row[r];
            if (row >= 0) {
                double *g = grid + (((size_t)row * 2 + s) * nh + k) * N;
#pragma unroll
                for (int kq = 0; kq < RL; kq++) st_stream(g + 2 * (q + SL * kq), o[kq]);
            }
        }
    }
}

template <int H>
__global__ void MAGIC_FFT_LB_R(H) fft_r2c_plan_kernel(const double2 *__restrict__ tw, R2cArgs a) {
    constexpr int R = fft2_rows(H), NT = fft2_threads(H), N = 2 * H, ROWLEN = fft2_rowlen(H);
    constexpr int R1 = fft_pick_radix(H), NB1 = H / R1;
    constexpr int RL = fft_last_radix(H), SL = H / RL, NIL = (SL + 1) / 2;
    extern __shared__ __align__(16) double2 fsm[];
    const int lev0 = blockIdx.x * R;
    const int sk = blockIdx.y, s = sk / a.nh, k = sk - s * a.nh;
    const int field = blockIdx.z;
    const int rows = min(R, a.n_lev - lev0);
    // ---- first pass on the registers of the (coalesced) row loads
    {
        constexpr int PER = (R * NB1 + NT - 1) / NT;
        double2 reg[PER][R1];
#pragma unroll
        for (int u = 0; u < PER; u++) {
            const int idx = threadIdx.x + u * NT, r = idx / NB1, b = idx - r * NB1;
#pragma unroll
            for (int j = 0; j < R1; j++) {
                reg[u][j] = make_double2(0.0, 0.0);
                if (idx < R * NB1 && r < rows)
                    reg[u][j] = ld_stream(a.grid + ((((size_t)field * a.n_lev + lev0 + r) * 2 + s) * a.nh + k) * N + 2 * (b + NB1 * j));
            }
        }
#pragma unroll
        for (int u = 0; u < PER; u++) {
            const int idx = threadIdx.x + u * NT, r = idx / NB1, b = idx - r * NB1;
            if (idx < R * NB1) first_pass_out<H, R1>(fsm + r * ROWLEN, tw, b, reg[u], -1.0);
        }
    }
    __syncthreads();
    FftMid<H, R, NT, ROWLEN, H / R1, R1>::run(fsm, tw, -1.0);
    // ---- last pass + post-processing + scatter.  Item t of a row: butterflies (t, SL - t); t = 0: butterfly 0 and, for even
    //      SL, the self-paired butterfly SL/2.  Butterfly q yields orders mc = q + SL kq, whose partners H - mc belong to SL - q.
    const double w = a.wgauss[k], ws = w * a.osin2[k];
    const R2cField dests = a.fields[field];
    for (int item = threadIdx.x; item < R * NIL; item += NT) {
        const int t = item / R, r = item - t * R;
        if (r >= rows) continue;
        const bool has_v = (t > 0) || (SL % 2 == 0);
        const int u = t, v = (t > 0) ? SL - t : SL / 2;
        const double2 *x = fsm + r * ROWLEN;
        double2 in[RL], Zu[RL], Zv[RL];
#pragma unroll
        for (int j = 0; j < RL; j++) in[j] = x[fft_pad(u + SL * j)];
        butterfly<RL>(in, Zu, -1.0);
        if (has_v) {
#pragma unroll
            for (int j = 0; j < RL; j++) in[j] = x[fft_pad(v + SL * j)];
            butterfly<RL>(in, Zv, -1.0);
        }
        const int lev = lev0 + r;
        if (t > 0) {
#pragma unroll
            for (int kq = 0; kq < RL; kq++) {
                const int mu = u + SL * kq, mv = v + SL * (RL - 1 - kq);  // mu + mv = H
                if (mu < a.n_m) r2c_scatter(a, dests, s, k, mu, lev, Zu[kq], Zv[RL - 1 - kq], twid(tw, mu, -1.0), w, ws);
                if (mv < a.n_m) r2c_scatter(a, dests, s, k, mv, lev, Zv[RL - 1 - kq], Zu[kq], twid(tw, mv, -1.0), w, ws);
            }
        } else {
#pragma unroll
            for (int kq = 0; kq < RL; kq++) {
                const int mu = SL * kq;  // partner H - mu = SL (RL - kq); mu = 0 pairs with itself
                if (mu < a.n_m) r2c_scatter(a, dests, s, k, mu, lev, Zu[kq], kq == 0 ? Zu[0] : Zu[RL - kq], twid(tw, mu, -1.0), w, ws);
            }
            if (has_v) {
#pragma unroll
                for (int kq = 0; kq < RL; kq++) {
                    const int mv = v + SL * kq;  // partner v + SL (RL-1-kq)
                    if (mv < a.n_m) r2c_scatter(a, dests, s, k, mv, lev, Zv[kq], Zv[RL - 1 - kq], twid(tw, mv, -1.0), w, ws);
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// Prefetching variants of the planned kernels.  A CTA works through `tpc` consecutive row tiles; while tile t runs its
// barrier-separated Stockham passes, the input of tile t+1 is already on its way into a staging buffer in shared memory:
//   c2r: the gather of the (theta,m)-space rows (32-byte pieces, one per order) with cp.async (LDGSTS);
//   r2c: the grid rows (contiguous, 8 n_phi bytes each) with one TMA bulk copy per row (cp.async.bulk, completion on an
//        mbarrier).
// The first pass then reads its operands from the staging buffer instead of from global memory, so no warp ever waits a DRAM
// round trip inside the transform (ncu on the non-prefetching kernels: 40 % of all warp stalls were long-scoreboard stalls at
// 2 CTAs = 12 warps per SM; the register file, not shared memory, limits the CTA count, so the staging buffer costs no
// occupancy).  Arithmetic and results are identical to the kernels above.
// (128 registers: the register file is split over the four SM sub-partitions, a CTA of 6 warps puts 2 warps on two of them,
//  and two CTAs per SM need 4 warps x 32 lanes x 128 registers = one sub-partition's 16 K registers)
template <int H>
__global__ void __maxnreg__(128) fft_c2r_pf_kernel(const double2 *__restrict__ tw_g, const double *__restrict__ F, int ld, int n_m,
                                                                  int nh, int ncols, const int *__restrict__ colrow,
                                                                  double *__restrict__ grid, int tpc) {
    constexpr int R = fft2_rows(H), NT = fft2_threads(H), N = 2 * H, ROWLEN = fft2_rowlen(H);
    constexpr int R1 = fft_pick_radix(H), NB1 = H / R1, NI1 = (NB1 + 1) / 2;
    constexpr int RL = fft_last_radix(H), SL = H / RL;
    extern __shared__ __align__(16) double2 fsm[];
    double2 *stage = fsm + R * ROWLEN;  // [n_m][R]
    __shared__ int s_row[R];
#if MAGIC_FFT_C2R_STW
    constexpr bool TWS = true;
    double2 *stw = stage + n_m * R;     // tw[0 .. H): every index these passes and the Hermitian pre-processing use
    for (int i = threadIdx.x; i < H; i += NT) stw[i] = __ldg(tw_g + i);
    const double2 *tw = stw;            // visible after the first barrier of the tile loop
#else
    constexpr bool TWS = false;
    const double2 *tw = tw_g;
#endif
    const int sk = blockIdx.y, s = sk / nh, k = sk - s * nh;
    const int ntile = (ncols + R - 1) / R;
    const int t0 = blockIdx.x * tpc, t1 = min(t0 + tpc, ntile);
    const double *Frow = F + ((size_t)s * nh + k) * ld;
    const size_t mstride = (size_t)2 * nh * ld;
    auto gather = [&](int t) {
        const int cc0 = t * R, rows = min(R, ncols - cc0);
        for (int idx = threadIdx.x; idx < n_m * R; idx += NT) {
            const int mc = idx / R, r = idx - mc * R;
            if (r < rows) cp_async16(stage + idx, Frow + (size_t)mc * mstride + 2 * (cc0 + r));
            else stage[idx] = make_double2(0.0, 0.0);
        }
        cp_async_commit();
    };
    if (t0 < t1) gather(t0);
    for (int t = t0; t < t1; t++) {
        const int cc0 = t * R, rows = min(R, ncols - cc0);
        cp_async_wait_all();
        __syncthreads();  // staging buffer complete; every thread is past the last pass of the previous tile
        if (threadIdx.x < R) s_row[threadIdx.x] = threadIdx.x < rows ? colrow[cc0 + threadIdx.x] : -1;  // read after the next barriers
        // ---- Hermitian pre-processing + first pass (see fft_c2r_plan_kernel)
        for (int item = threadIdx.x; item < R * NI1; item += NT) {
            const int tt = item / R, r = item - tt * R;
            const bool has_v = (tt > 0) || (NB1 % 2 == 0);
            const int u = tt, v = (tt > 0) ? NB1 - tt : NB1 / 2;
            double2 A[R1], B[R1];
#pragma unroll
            for (int j = 0; j < R1; j++) {
                A[j] = make_double2(0.0, 0.0);
                B[j] = A[j];
                const int ka = u + NB1 * j, kb = v + NB1 * j;
                if (ka < n_m) A[j] = stage[ka * R + r];
                if (has_v && kb < n_m) B[j] = stage[kb * R + r];
            }
            double2 *row = fsm + r * ROWLEN;
            double2 Yu[R1], Yv[R1];
            if (tt > 0) {
#pragma unroll
                for (int j = 0; j < R1; j++) {
                    const double2 a = A[j], b = B[R1 - 1 - j];
                    const double2 w = twid<TWS>(tw, u + NB1 * j, 1.0);
                    const double2 cb = cconj(b), ca = cconj(a);
                    Yu[j] = cadd(cadd(a, cb), cmuli(cmul(w, csub(a, cb)), 1.0));
                    const double2 w2 = make_double2(-w.x, w.y);
                    Yv[R1 - 1 - j] = cadd(cadd(b, ca), cmuli(cmul(w2, csub(b, ca)), 1.0));
                }
            } else {
#pragma unroll
                for (int j = 0; j < R1; j++) {
                    double2 a = A[j];
                    double2 b = (j == 0) ? make_double2(0.0, 0.0) : A[R1 - j];
                    if (j == 0) a.y = 0.0;
                    const double2 w = twid<TWS>(tw, NB1 * j, 1.0);
                    const double2 cb = cconj(b);
                    Yu[j] = cadd(cadd(a, cb), cmuli(cmul(w, csub(a, cb)), 1.0));
                }
#pragma unroll
                for (int j = 0; j < R1; j++) {
                    const double2 a = B[j], b = B[R1 - 1 - j];
                    const double2 w = twid<TWS>(tw, v + NB1 * j, 1.0);
                    const double2 cb = cconj(b);
                    Yv[j] = cadd(cadd(a, cb), cmuli(cmul(w, csub(a, cb)), 1.0));
                }
            }
            first_pass_out<H, R1, TWS>(row, tw, u, Yu, 1.0);
            if (has_v) first_pass_out<H, R1, TWS>(row, tw, v, Yv, 1.0);
        }
        __syncthreads();  // the staging buffer has been consumed
        if (t + 1 < t1) gather(t + 1);
        FftMid<H, R, NT, ROWLEN, H / R1, R1, TWS>::run(fsm, tw, 1.0);
        for (int idx = threadIdx.x; idx < R * SL; idx += NT) {
            const int r = idx / SL, q = idx - r * SL;
            const double2 *x = fsm + r * ROWLEN;
            double2 in[RL], o[RL];
#pragma unroll
            for (int j = 0; j < RL; j++) in[j] = x[fft_pad(q + SL * j)];
            butterfly<RL>(in, o, 1.0);
            {
                const int row = s_row[r];
                if (row >= 0) {
                    double *g = grid + (((size_t)row * 2 + s) * nh + k) * N;
#pragma unroll
                    for (int kq = 0; kq < RL; kq++) st_stream(g + 2 * (q + SL * kq), o[kq]);
                }
            }
        }
    }
}

template <int H>
__global__ void __maxnreg__(128) fft_r2c_pf_kernel(const double2 *__restrict__ tw, R2cArgs a, int tpc) {
    constexpr int R = fft2_rows(H), NT = fft2_threads(H), N = 2 * H, ROWLEN = fft2_rowlen(H);
    constexpr int R1 = fft_pick_radix(H), NB1 = H / R1;
    constexpr int RL = fft_last_radix(H), SL = H / RL, NIL = (SL + 1) / 2;
    extern __shared__ __align__(16) double2 fsm[];
    double2 *stage = fsm + R * ROWLEN;  // [R][H]
    __shared__ uint64_t bar;
    const int sk = blockIdx.y, s = sk / a.nh, k = sk - s * a.nh;
    const int field = blockIdx.z;
    const int ntile = (a.n_lev + R - 1) / R;
    const int t0 = blockIdx.x * tpc, t1 = min(t0 + tpc, ntile);
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        mbar_init_fence();
    }
    __syncthreads();
    auto fetch = [&](int t) {  // one thread: arm the barrier with the byte count, then one bulk copy per grid row
        const int lev0 = t * R, rows = min(R, a.n_lev - lev0);
        mbar_expect_tx(&bar, (unsigned)(rows * N * sizeof(double)));
        for (int r = 0; r < rows; r++)
            bulk_g2s(stage + r * H, a.grid + ((((size_t)field * a.n_lev + lev0 + r) * 2 + s) * a.nh + k) * N, (unsigned)(N * sizeof(double)), &bar);
    };
    if (threadIdx.x == 0 && t0 < t1) fetch(t0);
    const double w = a.wgauss[k], ws = w * a.osin2[k];
    const R2cField dests = a.fields[field];
    for (int t = t0; t < t1; t++) {
        const int lev0 = t * R, rows = min(R, a.n_lev - lev0);
        mbar_wait(&bar, (t - t0) & 1);
        // ---- first pass on the staged rows
        {
            constexpr int PER = (R * NB1 + NT - 1) / NT;
#pragma unroll
            for (int u = 0; u < PER; u++) {
                const int idx = threadIdx.x + u * NT, r = idx / NB1, b = idx - r * NB1;
                if (idx < R * NB1) {
                    double2 reg[R1];
#pragma unroll
                    for (int j = 0; j < R1; j++) reg[j] = r < rows ? stage[r * H + b + NB1 * j] : make_double2(0.0, 0.0);
                    first_pass_out<H, R1>(fsm + r * ROWLEN, tw, b, reg, -1.0);
                }
            }
        }
        __syncthreads();  // the staged rows have been consumed, the first-pass results are in place
        if (threadIdx.x == 0 && t + 1 < t1) fetch(t + 1);
        FftMid<H, R, NT, ROWLEN, H / R1, R1>::run(fsm, tw, -1.0);
        // ---- last pass + post-processing + scatter (see fft_r2c_plan_kernel)
        for (int item = threadIdx.x; item < R * NIL; item += NT) {
            const int tt = item / R, r = item - tt * R;
            if (r >= rows) continue;
            const bool has_v = (tt > 0) || (SL % 2 == 0);
            const int u = tt, v = (tt > 0) ? SL - tt : SL / 2;
            const double2 *x = fsm + r * ROWLEN;
            double2 in[RL], Zu[RL], Zv[RL];
            // post-processing twiddles of the paired orders, requested before the butterflies (all indices are < H, so the loads
            // need no predicate): their L1 latency used to sit between the butterfly and the scatter (ncu: the DMULs consuming
            // them held 20 % of the stall samples)
            constexpr bool HOIST = MAGIC_FFT_R2C_TWHOIST && RL <= 4;
            double2 wu[HOIST ? RL : 1], wv[HOIST ? RL : 1];
            if constexpr (HOIST) {
                if (tt > 0) {
#pragma unroll
                    for (int kq = 0; kq < RL; kq++) {
                        wu[kq] = twid(tw, u + SL * kq, -1.0);
                        wv[kq] = twid(tw, v + SL * (RL - 1 - kq), -1.0);
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < RL; j++) in[j] = x[fft_pad(u + SL * j)];
            butterfly<RL>(in, Zu, -1.0);
            if (has_v) {
#pragma unroll
                for (int j = 0; j < RL; j++) in[j] = x[fft_pad(v + SL * j)];
                butterfly<RL>(in, Zv, -1.0);
            }
            const int lev = lev0 + r;
            if (tt > 0) {
#pragma unroll
                for (int kq = 0; kq < RL; kq++) {
                    const int mu = u + SL * kq, mv = v + SL * (RL - 1 - kq);  // mu + mv = H
                    if (mu < a.n_m) r2c_scatter(a, dests, s, k, mu, lev, Zu[kq], Zv[RL - 1 - kq], HOIST ? wu[HOIST ? kq : 0] : twid(tw, mu, -1.0), w, ws);
                    if (mv < a.n_m) r2c_scatter(a, dests, s, k, mv, lev, Zv[RL - 1 - kq], Zu[kq], HOIST ? wv[HOIST ? kq : 0] : twid(tw, mv, -1.0), w, ws);
                }
            } else {
#pragma unroll
                for (int kq = 0; kq < RL; kq++) {
                    const int mu = SL * kq;
                    if (mu < a.n_m) r2c_scatter(a, dests, s, k, mu, lev, Zu[kq], kq == 0 ? Zu[0] : Zu[RL - kq], twid(tw, mu, -1.0), w, ws);
                }
                if (has_v) {
#pragma unroll
                    for (int kq = 0; kq < RL; kq++) {
                        const int mv = v + SL * kq;
                        if (mv < a.n_m) r2c_scatter(a, dests, s, k, mv, lev, Zv[kq], Zv[RL - 1 - kq], twid(tw, mv, -1.0), w, ws);
                    }
                }
            }
        }
        __syncthreads();  // the last pass has read fsm: the next tile's first pass may overwrite it
    }
}

// ------------------------------------------------------------------------------------------------------
// Generic (run-time plan) kernels for lengths without a compiled plan: ping-pong Stockham, radices {4,2,3,5}.
template <int RX>
__device__ __forceinline__ void stockham_pass(const double2 *__restrict__ src, double2 *__restrict__ dst, int rows, int H, int len, int s,
                                              const double2 *__restrict__ tw, int N, double sg) {
    const int m = len / RX, nb = H / RX, twstep = N / len;
    for (int idx = threadIdx.x; idx < rows * nb; idx += blockDim.x) {
        int row = idx / nb, b = idx - row * nb;
        int p = b / s, q = b - p * s;
        const double2 *x = src + row * H;
        double2 *y = dst + row * H;
        double2 a[RX], o[RX];
#pragma unroll
        for (int j = 0; j < RX; j++) a[j] = x[q + s * (p + m * j)];
        butterfly<RX>(a, o, sg);
        y[q + s * (RX * p)] = o[0];
#pragma unroll
        for (int k = 1; k < RX; k++) y[q + s * (RX * p + k)] = (p == 0) ? o[k] : cmul(o[k], twid(tw, p * k * twstep, sg));
    }
}

__device__ __forceinline__ double2 *stockham_fft(double2 *buf0, double2 *buf1, int rows, const FftPlan &pl, double sg) {
    int len = pl.H, s = 1;
    double2 *src = buf0, *dst = buf1;
    for (int f = 0; f < pl.nfac; f++) {
        int r = pl.fac[f];
        if (r == 4) stockham_pass<4>(src, dst, rows, pl.H, len, s, pl.tw, pl.N, sg);
        else if (r == 2) stockham_pass<2>(src, dst, rows, pl.H, len, s, pl.tw, pl.N, sg);
        else if (r == 3) stockham_pass<3>(src, dst, rows, pl.H, len, s, pl.tw, pl.N, sg);
        else stockham_pass<5>(src, dst, rows, pl.H, len, s, pl.tw, pl.N, sg);
        __syncthreads();
        len /= r;
        s *= r;
        double2 *tmp = src; src = dst; dst = tmp;
    }
    return src;
}

constexpr int FFT_THREADS = 256;

__global__ void __launch_bounds__(FFT_THREADS) fft_c2r_kernel(FftPlan pl, const double *__restrict__ F, int ld, int n_m, int nh, int ncols,
                                                            const int *__restrict__ colrow, double *__restrict__ grid, int R) {
    extern __shared__ __align__(16) double2 fsm[];
    const int H = pl.H, N = pl.N;
    double2 *buf0 = fsm, *buf1 = fsm + (size_t)R * H;
    const int cc0 = blockIdx.x * R;
    const int sk = blockIdx.y, s = sk / nh, k = sk - s * nh;
    const int rows = min(R, ncols - cc0);
    for (int idx = threadIdx.x; idx < R * H; idx += blockDim.x) {
        int mc = idx / R, r = idx - mc * R;
        double2 v = make_double2(0.0, 0.0);
        if (mc < n_m && r < rows) {
            v = *reinterpret_cast<const double2 *>(F + ((size_t)(mc * 2 + s) * nh + k) * ld + 2 * (cc0 + r));
            if (mc == 0) v.y = 0.0;
        }
        buf1[r * H + mc] = v;
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < R * H; idx += blockDim.x) {
        int r = idx / H, kk = idx - r * H;
        const double2 *c = buf1 + r * H;
        double2 ck = c[kk];
        double2 cm = (kk == 0) ? make_double2(0.0, 0.0) : cconj(c[H - kk]);
        buf0[r * H + kk] = cadd(cadd(ck, cm), cmuli(cmul(twid(pl.tw, kk, 1.0), csub(ck, cm)), 1.0));
    }
    __syncthreads();
    double2 *res = stockham_fft(buf0, buf1, R, pl, 1.0);
    for (int idx = threadIdx.x; idx < R * H; idx += blockDim.x) {
        int r = idx / H, j = idx - r * H;
        if (r < rows) {
            int row = colrow[cc0 + r];
            if (row >= 0) *reinterpret_cast<double2 *>(grid + (((size_t)row * 2 + s) * nh + k) * N + 2 * j) = res[r * H + j];
        }
    }
}

__global__ void __launch_bounds__(FFT_THREADS) fft_r2c_kernel(FftPlan pl, R2cArgs a, int R) {
    extern __shared__ __align__(16) double2 fsm[];
    const int H = pl.H, N = pl.N;
    double2 *buf0 = fsm, *buf1 = fsm + (size_t)R * H;
    const int lev0 = blockIdx.x * R;
    const int sk = blockIdx.y, s = sk / a.nh, k = sk - s * a.nh;
    const int field = blockIdx.z;
    const int rows = min(R, a.n_lev - lev0);
    for (int idx = threadIdx.x; idx < R * H; idx += blockDim.x) {
        int r = idx / H, j = idx - r * H;
        double2 v = make_double2(0.0, 0.0);
        if (r < rows) v = *reinterpret_cast<const double2 *>(a.grid + ((((size_t)field * a.n_lev + lev0 + r) * 2 + s) * a.nh + k) * N + 2 * j);
        buf0[idx] = v;
    }
    __syncthreads();
    double2 *Z = stockham_fft(buf0, buf1, R, pl, -1.0);
    const double w = a.wgauss[k], ws = w * a.osin2[k];
    const R2cField dests = a.fields[field];
    for (int idx = threadIdx.x; idx < rows * a.n_m; idx += blockDim.x) {
        int mc = idx / rows, r = idx - mc * rows;
        const double2 *z = Z + r * H;
        r2c_scatter(a, dests, s, k, mc, lev0 + r, z[mc], z[mc == 0 ? 0 : H - mc], twid(pl.tw, mc, -1.0), w, ws);
    }
}

inline int fft_rows_per_cta(int H, int want) {
    int r = want;  // two ping-pong buffers of R*H complex doubles; keep below ~96 KB so two CTAs fit per SM
    while (r > 1 && (size_t)2 * r * H * sizeof(double2) > 96 * 1024) r >>= 1;
    return r;
}

// ---- dispatch ---------------------------------------------------------------------------------------------
#define MAGIC_FFT_PLANS(X) X(16) X(24) X(32) X(48) X(64) X(72) X(96) X(128) X(144) X(192) X(256) X(384) X(512) X(768) X(1024) X(1536)

inline bool fft_has_plan(int H) {
#define X(h) if (H == h) return true;
    MAGIC_FFT_PLANS(X)
#undef X
    return false;
}
inline size_t fft_plan_smem(int H) { return (size_t)fft2_rows(H) * fft2_rowlen(H) * sizeof(double2); }
inline size_t fft_pf_smem_c2r(int H, int n_m) { return fft_plan_smem(H) + (size_t)n_m * fft2_rows(H) * sizeof(double2) + (MAGIC_FFT_C2R_STW ? (size_t)H * sizeof(double2) : 0); }
inline size_t fft_pf_smem_r2c(int H) { return fft_plan_smem(H) + (size_t)H * fft2_rows(H) * sizeof(double2); }
// MAGIC_FFT_PF=0 selects the non-prefetching planned kernels (A/B measurements); tiles per CTA: MAGIC_FFT_TPC (default 8)
inline bool fft_use_pf() { static int v = -1; if (v < 0) { const char *e = getenv("MAGIC_FFT_PF"); v = (e && atoi(e) == 0) ? 0 : 1; } return v == 1; }
inline int fft_tpc() { static int v = -1; if (v < 0) { const char *e = getenv("MAGIC_FFT_TPC"); v = e ? max(1, atoi(e)) : 8; } return v; }

// percentage of the 228 KB shared-memory maximum that `ctas` resident CTAs of `bytes` dynamic shared memory (+ 1 KB each) need
#ifndef MAGIC_FFT_PF_CTAS
#define MAGIC_FFT_PF_CTAS 2
#endif
#ifndef MAGIC_FFT_PLAN_CTAS
#define MAGIC_FFT_PLAN_CTAS 2
#endif
inline int fft_carveout(size_t bytes, int ctas) {
    const double need = (double)ctas * (double)(bytes + 1024) / (228.0 * 1024.0) * 100.0;
    int pct = (int)need + 2;
    return pct > 100 ? 100 : pct;
}

inline cudaError_t fft_setup_attributes(int H) {
    cudaError_t e = cudaFuncSetAttribute(fft_c2r_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(fft_r2c_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return e;
#define X(h)                                                                                                              \
    if (H == h) {                                                                                                         \
        e = cudaFuncSetAttribute(fft_c2r_plan_kernel<h>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fft_plan_smem(h)); \
        if (e != cudaSuccess) return e;                                                                                   \
        e = cudaFuncSetAttribute(fft_r2c_plan_kernel<h>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fft_plan_smem(h)); \
        if (e != cudaSuccess) return e;                                                                                   \
        e = cudaFuncSetAttribute(fft_c2r_pf_kernel<h>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fft_pf_smem_c2r(h, h)); \
        if (e != cudaSuccess) return e;                                                                                   \
        e = cudaFuncSetAttribute(fft_r2c_pf_kernel<h>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fft_pf_smem_r2c(h)); \
        if (e != cudaSuccess) return e;                                                                                   \
        /* shared-memory carve-out: exactly what the resident CTAs need -- the rest of the 256 KB array stays L1, which    \
           holds the twiddle table (measured: with the whole array carved out as shared memory the c2r kernel loses 50 %) */ \
        cudaFuncSetAttribute(fft_c2r_pf_kernel<h>, cudaFuncAttributePreferredSharedMemoryCarveout, fft_carveout(fft_pf_smem_c2r(h, h), MAGIC_FFT_PF_CTAS)); \
        cudaFuncSetAttribute(fft_r2c_pf_kernel<h>, cudaFuncAttributePreferredSharedMemoryCarveout, fft_carveout(fft_pf_smem_r2c(h), MAGIC_FFT_PF_CTAS)); \
        cudaFuncSetAttribute(fft_c2r_plan_kernel<h>, cudaFuncAttributePreferredSharedMemoryCarveout, fft_carveout(fft_plan_smem(h), MAGIC_FFT_PLAN_CTAS)); \
        cudaFuncSetAttribute(fft_r2c_plan_kernel<h>, cudaFuncAttributePreferredSharedMemoryCarveout, fft_carveout(fft_plan_smem(h), MAGIC_FFT_PLAN_CTAS)); \
    }
    MAGIC_FFT_PLANS(X)
#undef X
    return cudaSuccess;
}

inline void launch_fft_c2r(const FftPlan &pl, const double *F, int ld, int n_m, int nh, int ncols, const int *colrow, double *grid,
                           cudaStream_t st) {
    const int H = pl.H;
#define X(h)                                                                                                                      \
    if (H == h) {                                                                                                                 \
        const int ntile = (ncols + fft2_rows(h) - 1) / fft2_rows(h);                                                              \
        if (fft_use_pf() && fft_pf_smem_c2r(h, n_m) <= (size_t)(227 * 1024 / MAGIC_FFT_PF_CTAS - 1024)) {                                                              \
            const int tpc = fft_tpc();                                                                                            \
            dim3 g((ntile + tpc - 1) / tpc, 2 * nh);                                                                              \
            fft_c2r_pf_kernel<h><<<g, fft2_threads(h), fft_pf_smem_c2r(h, n_m), st>>>(pl.tw, F, ld, n_m, nh, ncols, colrow, grid, tpc); \
            return;                                                                                                               \
        }                                                                                                                         \
        dim3 g(ntile, 2 * nh);                                                                                                    \
        fft_c2r_plan_kernel<h><<<g, fft2_threads(h), fft_plan_smem(h), st>>>(pl.tw, F, ld, n_m, nh, ncols, colrow, grid);          \
        return;                                                                                                                   \
    }
    MAGIC_FFT_PLANS(X)
#undef X
    const int R = fft_rows_per_cta(H, 8);
    dim3 g((ncols + R - 1) / R, 2 * nh);
    fft_c2r_kernel<<<g, FFT_THREADS, (size_t)2 * R * H * sizeof(double2), st>>>(pl, F, ld, n_m, nh, ncols, colrow, grid, R);
}

inline void launch_fft_r2c(const FftPlan &pl, const R2cArgs &a, int nfields, cudaStream_t st) {
    const int H = pl.H;
#define X(h)                                                                                      \
    if (H == h) {                                                                                 \
        const int ntile = (a.n_lev + fft2_rows(h) - 1) / fft2_rows(h);                            \
        if (fft_use_pf() && fft_pf_smem_r2c(h) <= (size_t)(227 * 1024 / MAGIC_FFT_PF_CTAS - 1024)) {                                   \
            const int tpc = fft_tpc();                                                            \
            dim3 g((ntile + tpc - 1) / tpc, 2 * a.nh, nfields);                                   \
            fft_r2c_pf_kernel<h><<<g, fft2_threads(h), fft_pf_smem_r2c(h), st>>>(pl.tw, a, tpc);   \
            return;                                                                               \
        }                                                                                         \
        dim3 g(ntile, 2 * a.nh, nfields);                                                         \
        fft_r2c_plan_kernel<h><<<g, fft2_threads(h), fft_plan_smem(h), st>>>(pl.tw, a);            \
        return;                                                                                   \
    }
    MAGIC_FFT_PLANS(X)
#undef X
    const int R = fft_rows_per_cta(H, 8);
    dim3 g((a.n_lev + R - 1) / R, 2 * a.nh, nfields);
    fft_r2c_kernel<<<g, FFT_THREADS, (size_t)2 * R * H * sizeof(double2), st>>>(pl, a, R);
}

}  // namespace magic
