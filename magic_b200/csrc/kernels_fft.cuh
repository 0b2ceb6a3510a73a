// kernels_fft.cuh -- batched real FFT along phi, hand written (no cuFFT on the path).
//
// Replaces fft.f90:164-252 (fft_many / ifft_many, Temperton FFT99) with the same transform definition
// (fft.f90:262-268): c2r  x_j = sum_k c_k e^{+2 pi i jk/n}, c_{n-k}=conj(c_k) (unnormalised),
//                    r2c  c_k = (1/n) sum_j x_j e^{-2 pi i jk/n}.
// A real transform of length N is done as a complex transform of length H=N/2 (Cooley-Lewis-Welch
// packing, as fft991 does) with a mixed-radix {4,2,3,5} Stockham autosort held in shared memory; one
// CTA owns R rows.  The c2r kernel gathers its input from the (theta,m)-space matrices written by the
// Legendre GEMM (zero padding of orders mc >= n_m_max, shtransforms.f90:224-232, is implicit); the r2c
// kernel scatters its output, already weighted for the quadrature, into the analysis GEMM operands.
#pragma once
#include "common.cuh"

namespace magic {

__device__ __forceinline__ double2 cmul(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 cconj(double2 a) { return make_double2(a.x, -a.y); }
__device__ __forceinline__ double2 cscale(double2 a, double s) { return make_double2(a.x * s, a.y * s); }
// multiply by sign*i
__device__ __forceinline__ double2 cmuli(double2 a, double sg) { return make_double2(-sg * a.y, sg * a.x); }

// twiddle e^{sg * 2 pi i idx / N}
__device__ __forceinline__ double2 twid(const double2 *__restrict__ tw, int idx, double sg) {
    double2 w = __ldg(tw + idx);
    w.y *= sg;
    return w;
}

// One Stockham pass of radix R over `rows` sequences of length H living in src[row*H ..], writing dst.
// len = current sub-transform length, s = stride (product of radices done).  See oracle fft_stockham.
template <int R>
__device__ __forceinline__ void stockham_pass(const double2 *__restrict__ src, double2 *__restrict__ dst, int rows, int H,
                                              int len, int s, const double2 *__restrict__ tw, int N, double sg) {
    const int m = len / R;
    const int nb = H / R;  // butterflies per row = m*s
    const int twstep = N / len;
    for (int idx = threadIdx.x; idx < rows * nb; idx += blockDim.x) {
        int row = idx / nb, b = idx - row * nb;
        int p = b / s, q = b - p * s;
        const double2 *x = src + row * H;
        double2 *y = dst + row * H;
        double2 a[R];
#pragma unroll
        for (int j = 0; j < R; j++) a[j] = x[q + s * (p + m * j)];
        double2 o[R];
        if (R == 2) {
            o[0] = cadd(a[0], a[1]);
            o[1] = csub(a[0], a[1]);
        } else if (R == 4) {
            double2 t0 = cadd(a[0], a[2]), t1 = csub(a[0], a[2]), t2 = cadd(a[1], a[3]), t3 = cmuli(csub(a[1], a[3]), sg);
            o[0] = cadd(t0, t2);
            o[1] = cadd(t1, t3);
            o[2] = csub(t0, t2);
            o[3] = csub(t1, t3);
        } else if (R == 3) {
            const double s3 = 0.86602540378443864676372317075294;
            double2 t = cadd(a[1], a[2]);
            double2 u = cmuli(cscale(csub(a[1], a[2]), s3), sg);
            double2 c = make_double2(a[0].x - 0.5 * t.x, a[0].y - 0.5 * t.y);
            o[0] = cadd(a[0], t);
            o[1] = cadd(c, u);
            o[2] = csub(c, u);
        } else {  // R == 5
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
        y[q + s * (R * p)] = o[0];
#pragma unroll
        for (int k = 1; k < R; k++) y[q + s * (R * p + k)] = (p == 0) ? o[k] : cmul(o[k], twid(tw, p * k * twstep, sg));
    }
}

// Complex FFT of length H for `rows` rows: data in buf0, scratch buf1; returns pointer holding the result.
__device__ __forceinline__ double2 *stockham_fft(double2 *buf0, double2 *buf1, int rows, const FftPlan &pl, double sg) {
    int len = pl.H, s = 1;
    double2 *src = buf0, *dst = buf1;
    for (int f = 0; f < pl.nfac; f++) {
        int r = pl.fac[f];
        // twiddles of the length-`len` sub-transform are tw[(p*k) * (N/len)]; N = 2H so N/len is integral
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

// ------------------------------------------------------------------------------------------------------
// c2r: grid row (row index r = colrow[cc]) of parity s and colatitude k from column cc of F.
//   F element (mc, s, k, col cc) at F[((mc*2+s)*nh + k)*ld + 2*cc].
//   blockIdx.x = column chunk, blockIdx.y = s*nh + k.
__global__ void __launch_bounds__(FFT_THREADS) fft_c2r_kernel(FftPlan pl, const double *__restrict__ F, int ld, int n_m,
                                                            int nh, int ncols, const int *__restrict__ colrow,
                                                            double *__restrict__ grid, int R) {
    extern __shared__ __align__(16) double2 fsm[];
    const int H = pl.H, N = pl.N;
    double2 *buf0 = fsm, *buf1 = fsm + (size_t)R * H;
    const int cc0 = blockIdx.x * R;
    const int sk = blockIdx.y;  // s*nh + k
    const int s = sk / nh, k = sk - s * nh;
    const int rows = min(R, ncols - cc0);
    // gather c_mc (zero for mc >= n_m); thread order: column fastest so a warp reads contiguous 16B pieces
    for (int idx = threadIdx.x; idx < R * H; idx += blockDim.x) {
        int mc = idx / R, r = idx - mc * R;
        double2 v = make_double2(0.0, 0.0);
        if (mc < n_m && r < rows) {
            v = *reinterpret_cast<const double2 *>(F + ((size_t)(mc * 2 + s) * nh + k) * ld + 2 * (cc0 + r));
            if (mc == 0) v.y = 0.0;  // c2r ignores Im c_0 (fft.f90:262-268)
        }
        buf1[r * H + mc] = v;
    }
    __syncthreads();
    // Y_k = (c_k + conj c_{H-k}) + i e^{2 pi i k/N} (c_k - conj c_{H-k}),  c_H = 0 here (n_m <= H)
    for (int idx = threadIdx.x; idx < R * H; idx += blockDim.x) {
        int r = idx / H, kk = idx - r * H;
        const double2 *c = buf1 + r * H;
        double2 ck = c[kk];
        double2 cm = (kk == 0) ? make_double2(0.0, 0.0) : cconj(c[H - kk]);
        double2 e = cadd(ck, cm), o = cmuli(cmul(twid(pl.tw, kk, 1.0), csub(ck, cm)), 1.0);
        buf0[r * H + kk] = cadd(e, o);
    }
    __syncthreads();
    double2 *res = stockham_fft(buf0, buf1, R, pl, 1.0);
    // z_j = x_{2j} + i x_{2j+1}: the row is the interleaved complex array itself
    for (int idx = threadIdx.x; idx < R * H; idx += blockDim.x) {
        int r = idx / H, j = idx - r * H;
        if (r < rows) {
            int row = colrow[cc0 + r];
            if (row >= 0)
                *reinterpret_cast<double2 *>(grid + (((size_t)row * 2 + s) * nh + k) * N + 2 * j) = res[r * H + j];
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// r2c: one CTA transforms R grid rows (consecutive levels of one field, fixed s,k) and scatters the
// weighted coefficients of orders mc < n_m into the analysis operands.
//   blockIdx.x = level chunk, blockIdx.y = s*nh + k, blockIdx.z = field.
struct R2cArgs {
    const double *grid;
    int n_lev, nh, n_m, NHP;
    const double *wgauss;  // 2 pi * gauss(k) / n_phi  (quadrature weight and forward-FFT normalisation)
    const double *osin2;   // 1/sin^2(theta_k)
    const R2cField *fields;
    double *B[2];          // scalar-class / vector-class analysis operands
    int ldB[2];
    int minc;
};

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
        if (r < rows)
            v = *reinterpret_cast<const double2 *>(a.grid + ((((size_t)field * a.n_lev + lev0 + r) * 2 + s) * a.nh + k) * N + 2 * j);
        buf0[idx] = v;
    }
    __syncthreads();
    double2 *Z = stockham_fft(buf0, buf1, R, pl, -1.0);
    const double w = a.wgauss[k], ws = w * a.osin2[k];
    const R2cField fd = a.fields[field];
    // X_k = E_k + e^{-2 pi i k/N} O_k, E=(Z_k+conj Z_{H-k})/2, O=-i (Z_k-conj Z_{H-k})/2 ; only k < n_m kept.
    for (int idx = threadIdx.x; idx < rows * a.n_m; idx += blockDim.x) {
        int mc = idx / rows, r = idx - mc * rows;
        const double2 *z = Z + r * H;
        double2 zk = z[mc], zm = cconj(z[mc == 0 ? 0 : H - mc]);
        double2 e = cscale(cadd(zk, zm), 0.5), o = cmuli(cscale(csub(zk, zm), 0.5), -1.0);
        double2 x = cadd(e, cmul(twid(pl.tw, mc, -1.0), o));
        const double dm = (double)(mc * a.minc);
#pragma unroll
        for (int d = 0; d < 2; d++) {
            const R2cDest ds = fd.d[s][d];
            if (ds.rtype == R_NONE) continue;
            double2 v;
            if (ds.rtype == R_W) v = cscale(x, w);
            else if (ds.rtype == R_WS) v = cscale(x, ws);
            else if (ds.rtype == R_NEG_WS) v = cscale(x, -ws);
            else v = make_double2(dm * ws * x.y, -dm * ws * x.x);  // -i m ws x
            const int rowsB = (ds.cls == 0) ? a.NHP : 2 * a.NHP;
            size_t off = ((size_t)(mc * 2 + ds.p) * rowsB + ds.seg * a.NHP + k) * a.ldB[ds.cls] +
                         2 * ((size_t)ds.col * a.n_lev + lev0 + r);
            *reinterpret_cast<double2 *>(a.B[ds.cls] + off) = v;
        }
    }
}

inline int fft_rows_per_cta(int H, int want) {
    // two ping-pong buffers of R*H complex doubles; keep below ~96 KB so two CTAs fit per SM
    int r = want;
    while (r > 1 && (size_t)2 * r * H * sizeof(double2) > 96 * 1024) r >>= 1;
    return r;
}

}  // namespace magic
