// engine.cuh -- host-side engine: the SHT handle (tables, maps, FFT plan) and the level-batched transform
// pipeline (`Batch`) that both the per-call `module sht` API and the radial loop are built on.
#pragma once
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "common.cuh"
#include "kernels_fft.cuh"
#include "kernels_gemm.cuh"
#include "kernels_grid.cuh"
#include "kernels_spec.cuh"

namespace magic {

extern thread_local std::string g_last_error;

#define MCHECK(call)                                                                                          \
    do {                                                                                                      \
        cudaError_t e__ = (call);                                                                             \
        if (e__ != cudaSuccess) {                                                                             \
            char b__[512];                                                                                    \
            snprintf(b__, sizeof(b__), "%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            magic::g_last_error = b__;                                                                        \
            return 1;                                                                                         \
        }                                                                                                     \
    } while (0)

#define MFAIL(msg)                       \
    do {                                 \
        magic::g_last_error = (msg);     \
        return 1;                        \
    } while (0)

template <typename T>
inline int dev_upload_vec(T **dptr, const std::vector<T> &v) {
    size_t bytes = sizeof(T) * (v.empty() ? 1 : v.size());
    MCHECK(cudaMalloc((void **)dptr, bytes));
    if (!v.empty()) MCHECK(cudaMemcpy(*dptr, v.data(), sizeof(T) * v.size(), cudaMemcpyHostToDevice));
    return 0;
}

inline int pad_up(int x, int m) { return (x + m - 1) / m * m; }

}  // namespace magic

struct magic_sht {
    int l_max, m_max, minc, n_theta, n_phi, nlat_padded, n_m, lm_max, nh, NHP, dev;
    std::vector<double> theta_ord, gauss;          // gauleg output (monotone north->south)
    std::vector<int> lm2l, lm2m, lstart, ne1, no1; // st_map; per-mc first lm and even/odd-parity degree counts of m .. l_max+1
    std::vector<long long> off;                    // [n_m][2] table block offsets: P_even, P_odd
    double *d_clm = nullptr;                       // c(l) = sqrt((l+m)(l-m)/((2l-1)(2l+1))), l = m .. l_max+2, at lstart[mc] + 2 mc + (l-m)
    std::vector<int> kmin;                         // per mc: first colatitude with non-negligible table entries
    double polar_eps = 1e-40;
    unsigned char *d_fskip_syn = nullptr, *d_fskip_an = nullptr;  // fragment-level skip tables (table_fskip_kernel), or null
    int FS = 0, FA = 0;                                           // their strides per table block
    double *d_tab = nullptr;
    long long *d_off = nullptr;
    double *d_sinth = nullptr, *d_costh = nullptr, *d_wgauss = nullptr, *d_osin2 = nullptr;
    int *d_lm2l = nullptr, *d_lm2m = nullptr, *d_lstart = nullptr;
    double2 *d_tw = nullptr;
    magic::FftPlan fft;
    bool an_wide = false;  // analysis GEMM with the 64 x 128 tile (common.cuh)
    cudaStream_t stream = nullptr;
    struct CallCtx *call = nullptr;  // lazily built single-level pipeline for the per-call API
    long long launches = 0;
};

namespace magic {

// A level-batched transform pipeline with a fixed column program.
struct BatchSpec {
    std::vector<ScalCol> scal;      // synthesis scalar-class columns
    std::vector<VecPair> vec;       // synthesis vector-class pairs
    std::vector<int> field_s;       // grid field index written by scalar column i
    std::vector<int> field_v;       // grid field index written by vector column i (2 per pair: theta, phi)
    int nfield_in = 0;              // number of synthesised grid fields
    int nfield_out = 0;             // number of product grid fields (analysis inputs)
    std::vector<int> afield_s;      // product field analysed as scalar column i
    std::vector<int> afield_vt, afield_vp;  // product fields (theta-type, phi-type) of analysis pair i
};

struct Buffers {  // big device arrays, sized for the largest chunk
    double *B = nullptr, *F = nullptr, *gin = nullptr, *gout = nullptr;   // synthesis operand, (theta,m) space, grids
    double *Ba = nullptr, *Ca = nullptr, *nl_s = nullptr, *nl_v = nullptr;  // analysis operand / result, per-call spectra
    unsigned long long *courmax = nullptr;
    size_t bytes = 0;
};

struct Layout {  // descriptors for one chunk size
    int n_lev = 0;
    int ncol_s = 0, npair_v = 0, ncol = 0, N = 0;    // synthesis: complex columns = ncol_s + 2 npair_v, N = pad64(2 ncol n_lev)
    int nf_s = 0, npair_a = 0, nfa = 0, Na = 0;      // analysis:  complex columns = nf_s + 2 npair_a
    std::vector<long long> offB, offC;
    long long szB = 0, szF = 0, szBa = 0, szCa = 0;
    int ntiles_syn = 0, ntiles_an = 0;
    long long *d_offB = nullptr, *d_offC = nullptr;
    int2 *d_prep_blks = nullptr;
    int n_prep_blks = 0;
    GemmProb *d_probs_syn = nullptr, *d_probs_an = nullptr;
    int2 *d_tiles_syn = nullptr, *d_tiles_an = nullptr;
    int *d_colrow = nullptr;
    ScalCol *d_scal = nullptr;
    VecPair *d_vec = nullptr;
    R2cField *d_r2c = nullptr;
    bool an_wide = false;                // tile shape of the analysis GEMM this layout was sized for
    int nsrc = 0, src_slot[MAGIC_MAX_SRC] = {0};  // sources the synthesis columns read: dense index -> caller's slot
    double flops_syn = 0, flops_an = 0;  // executed (padded) flops, for diagnostics
};

int layout_build(magic_sht *h, const BatchSpec &spec, int n_lev, Layout &L);
void layout_sizes(const magic_sht *h, const BatchSpec &spec, int n_lev, Layout &L);
int layout_bind(magic_sht *h, const BatchSpec &spec, Layout &L, const Buffers &buf);
void layout_free(Layout &L);
int buffers_alloc(magic_sht *h, const BatchSpec &spec, const Layout &Lmax, Buffers &buf);
void buffers_free(Buffers &buf);
size_t buffers_bytes(const magic_sht *h, const BatchSpec &spec, const Layout &L);                     // arena form: bytes needed ...
void buffers_carve(const magic_sht *h, const BatchSpec &spec, const Layout &L, char *base, Buffers &buf);  // ... and the carving

// pipeline stages (all asynchronous on h->stream)
// ev (optional): 4 events for synthesis (start, after prep, after Legendre, after FFT); 3 for analysis
// (start, after FFT, after Legendre).
int run_synthesis(magic_sht *h, const BatchSpec &spec, const Layout &L, const Buffers &buf, const double *const src[MAGIC_MAX_SRC],
                  const LevelInfo *d_lev, cudaEvent_t *ev);
int run_analysis(magic_sht *h, const BatchSpec &spec, const Layout &L, const Buffers &buf, const LevelInfo *d_lev, cudaEvent_t *ev,
                 bool extract = true);
// get_br_v_bcs (nonlinear_bcs.f90:24-74) for one boundary level, device pointers throughout: b, dw, z are the level's spectra
// (complex [lm_max]); writes br_vt_lm, br_vp_lm (complex [lm_max]).  Runs on h->stream with the per-call single-level pipeline.
int br_v_bcs_dev(magic_sht *h, const double *b, const double *dw, const double *z, int lcut, double fac, double omega,
                 double *br_vt_lm, double *br_vp_lm);
ExtractArgs make_extract_args(const magic_sht *h, const Layout &L, const Buffers &buf, const LevelInfo *d_lev);
int sht_init(magic_sht *h);
void sht_free(magic_sht *h);

}  // namespace magic
