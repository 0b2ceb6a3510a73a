// api_transp.cu -- C ABI: r <-> LM redistribution (a 5th `type_mpitransp`, mpi_transpose.f90:18-54) with
// the alltoallv semantics of type_mpiatoav: pack (:320-333 / :490-506), exchange, permuting unpack
// (:341-357 / :515-528).  The exchange is a grouped ncclSend/ncclRecv all-to-all over NVLink that reads / writes the
// LM-distributed array in place (its blocks are contiguous per peer and field), so only the R side packs; NCCL is
// resolved at run time (dlopen) so the library loads on machines without it and single-rank use needs none.
#include <dlfcn.h>
#include <nccl.h>

#include "../../include/magic_sht.h"
#include "engine.cuh"

using namespace magic;

namespace {

struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi g_nccl;

int nccl_load() {
    if (g_nccl.lib) return 0;
    void *lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) MFAIL(std::string("cannot load NCCL: ") + dlerror());
    g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))dlsym(lib, "ncclGetUniqueId");
    g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))dlsym(lib, "ncclCommInitRank");
    g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))dlsym(lib, "ncclCommDestroy");
    g_nccl.Send = (decltype(g_nccl.Send))dlsym(lib, "ncclSend");
    g_nccl.Recv = (decltype(g_nccl.Recv))dlsym(lib, "ncclRecv");
    g_nccl.GroupStart = (decltype(g_nccl.GroupStart))dlsym(lib, "ncclGroupStart");
    g_nccl.GroupEnd = (decltype(g_nccl.GroupEnd))dlsym(lib, "ncclGroupEnd");
    g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))dlsym(lib, "ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.Send || !g_nccl.Recv || !g_nccl.GroupStart || !g_nccl.GroupEnd)
        MFAIL("NCCL library lacks required symbols");
    g_nccl.lib = lib;
    return 0;
}

#define NCHECK(call)                                                                                      \
    do {                                                                                                  \
        ncclResult_t r__ = (call);                                                                        \
        if (r__ != ncclSuccess) {                                                                         \
            magic::g_last_error = std::string(#call) + " -> " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r__) : "nccl error"); \
            return 1;                                                                                     \
        }                                                                                                 \
    } while (0)

// getBlocks, parallel.f90:75-92 (1-based inclusive)
void get_blocks(int n_points, int n_procs, std::vector<int> &start, std::vector<int> &stop) {
    start.assign(n_procs, 0);
    stop.assign(n_procs, 0);
    int n_loc = n_points / n_procs, rem = n_points - n_loc * n_procs;
    for (int p = 0; p < n_procs; p++) {
        start[p] = n_loc * p + std::max(p + rem - n_procs, 0) + 1;
        stop[p] = n_loc * (p + 1) + std::max(p + rem + 1 - n_procs, 0);
        if (p != 0) start[p] = stop[p - 1] + 1;
    }
}

// lo_map: snake ordering for n_procs <= l_max/2 (blocking.f90:387-544), else l-major (blocking.f90:339-385).
// Returns lo2st (0-based st index of each lo position) and the 1-based inclusive lm range of every rank.
void build_lo_map(const magic_sht *h, int n_procs, std::vector<int> &lo2st, std::vector<int> &lm_start, std::vector<int> &lm_stop) {
    const int l_max = h->l_max, m_max = h->m_max, minc = h->minc;
    auto st_index = [&](int l, int m) { return h->lstart[m / minc] + (l - m); };
    lo2st.clear();
    if (n_procs <= l_max / 2) {
        // deal degrees l_max..0 to ranks 0,1,..,n-1,n-1,..,0,0,1,.. ("snake")
        std::vector<std::vector<int>> lists(n_procs);
        int proc = 0, owner_of_l0 = 0;
        bool up = true;
        for (int l = l_max; l >= 0; l--) {
            lists[proc].push_back(l);
            if (l == 0) owner_of_l0 = proc;
            if (up) { if (proc < n_procs - 1) proc++; else up = false; }
            else { if (proc > 0) proc--; else up = true; }
        }
        // rotate so that the owner of l=0 becomes rank 0: follow the cycle pc <- (owner+pc) mod n until it closes
        if (owner_of_l0 != 0) {
            std::vector<int> saved = lists[0];
            int pc = 0;
            for (;;) {
                int src = (owner_of_l0 + pc) % n_procs;
                if (src != 0) lists[pc] = lists[src];
                else { lists[pc] = saved; break; }
                pc = src;
            }
        }
        for (size_t i = 0; i < lists[0].size(); i++)
            if (lists[0][i] == 0) { std::swap(lists[0][0], lists[0][i]); break; }
        lm_start.assign(n_procs, 0);
        lm_stop.assign(n_procs, 0);
        for (int p = 0; p < n_procs; p++) {
            lm_start[p] = (int)lo2st.size() + 1;
            for (int l : lists[p])
                for (int m = 0; m <= std::min(m_max, l); m += minc) lo2st.push_back(st_index(l, m));
            lm_stop[p] = (int)lo2st.size();
        }
    } else {
        get_blocks(h->lm_max, n_procs, lm_start, lm_stop);
        for (int l = 0; l <= l_max; l++)
            for (int m = 0; m <= std::min(m_max, l); m += minc) lo2st.push_back(st_index(l, m));
    }
}

// ---- kernels: one thread per complex element of the packed buffer --------------------------------------
struct PackArgs {
    int n_procs, n_fields, n_r_max, lm_max;
    int llm, nlm;          // my lm slab (0-based start in lo order, count)
    int r0, nr;            // my radial slab (0-based start, count); for a part: my levels of this part
    int r_off, nr_arr;     // arr_Rloc holds nr_arr levels per field, of which this object moves [r_off, r_off + nr)
    const int *rstart;     // [n_procs] 0-based first level of each rank
    const int *rcount;     // [n_procs]
    const int *lstart;     // [n_procs] 0-based first lo index of each rank
    const int *lcount;     // [n_procs]
    const long long *disp; // [n_procs+1] element displacement of each peer segment, PER FIELD
    const int *lo2st;
    int direct;            // single rank: the "buffer" of the R-side kernels is arr_LMloc itself ([f][n_r_max][lm_lo]); a part
                           // then addresses row (f, r) at f * n_r_max + r0 + r instead of f * nr + r
};

__device__ __forceinline__ int find_seg(const long long *disp, int n, long long idx, int nf) {
    int lo = 0, hi = n;  // nf*disp[lo] <= idx < nf*disp[hi]
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (disp[mid] * nf <= idx) lo = mid; else hi = mid;
    }
    return lo;
}

// segment q of the LM-side buffer: [f][n_r in block q][lm in my slab]
__global__ void lmside_kernel(PackArgs a, const double2 *__restrict__ arr_LM_in, double2 *__restrict__ arr_LM_out,
                              const double2 *__restrict__ buf_in, double2 *__restrict__ buf_out, long long total) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    int q = find_seg(a.disp, a.n_procs, idx, a.n_fields);
    long long rel = idx - a.disp[q] * a.n_fields;
    int lm = (int)(rel % a.nlm);
    long long t = rel / a.nlm;
    int r = (int)(t % a.rcount[q]), f = (int)(t / a.rcount[q]);
    size_t pos = (size_t)lm + (size_t)a.nlm * ((size_t)(a.rstart[q] + r) + (size_t)a.n_r_max * f);
    if (buf_out) buf_out[idx] = arr_LM_in[pos];   // pack   (mpi_transpose.f90:320-333)
    else arr_LM_out[pos] = buf_in[idx];           // unpack (mpi_transpose.f90:515-528)
}

// segment p of the R-side buffer: [f][n_r in my slab][lm in rank p's lo slab], permuted to st order
__global__ void rside_kernel(PackArgs a, const double2 *__restrict__ arr_R_in, double2 *__restrict__ arr_R_out,
                             const double2 *__restrict__ buf_in, double2 *__restrict__ buf_out, long long total) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    int p = find_seg(a.disp, a.n_procs, idx, a.n_fields);
    long long rel = idx - a.disp[p] * a.n_fields;
    int lm = (int)(rel % a.lcount[p]);
    long long t = rel / a.lcount[p];
    int r = (int)(t % a.nr), f = (int)(t / a.nr);
    int lm_st = a.lo2st[a.lstart[p] + lm];
    size_t pos = (size_t)lm_st + (size_t)a.lm_max * ((size_t)(a.r_off + r) + (size_t)a.nr_arr * f);
    if (buf_out) buf_out[idx] = arr_R_in[pos];    // pack   (mpi_transpose.f90:490-506)
    else arr_R_out[pos] = buf_in[idx];            // unpack (mpi_transpose.f90:341-357)
}

// R-side pack/unpack as a tiled (l,m) <-> (m,l) transpose.  The packed buffers are in lo order (for one degree l the
// orders m are contiguous), the R-distributed arrays in st order (for one order m the degrees l are contiguous), so a
// CTA moves a tile of 32 degrees x 8 orders through shared memory: buffer accesses are 128-byte runs along m, array
// accesses 512-byte runs along l.  (The element-wise kernel above scatters 16-byte writes: 2.8 ms vs the copy-speed
// 0.6 ms for a 1.35 GB container at 8 ranks.)
constexpr int RT_L = 32, RT_M = 8, RT_ROWS = 16;
template <bool UNPACK>
__global__ void __launch_bounds__(256) rside_tiled_kernel(PackArgs a, const int *__restrict__ st2lo, const int *__restrict__ mstart,
                                                        const double2 *__restrict__ in, double2 *__restrict__ out, int rows_total,
                                                        int minc, int l_max, int n_m) {
    __shared__ double2 tile[4][RT_L][RT_M + 1];
    const int l0 = blockIdx.x * RT_L, mc0 = blockIdx.y * RT_M, row0 = blockIdx.z * RT_ROWS;
    if (mc0 * minc > l0 + RT_L - 1) return;  // tile above the diagonal m <= l
    const int tid = threadIdx.x;
    // buffer-side element of this thread (m fastest)
    const int il = tid / RT_M, im = tid % RT_M;
    const int lA = l0 + il, mcA = mc0 + im, mA = mcA * minc;
    const bool vA = lA <= l_max && mcA < n_m && mA <= lA;
    long long bufA = 0;
    int lcA = 0;
    if (vA) {
        int lo = st2lo[mstart[mcA] + lA - mA];
        int p = 0, hi = a.n_procs;
        while (hi - p > 1) { int mid = (p + hi) >> 1; if (a.lstart[mid] <= lo) p = mid; else hi = mid; }
        lcA = a.lcount[p];
        bufA = a.disp[p] * a.n_fields + (lo - a.lstart[p]);
    }
    // array-side element of this thread (l fastest)
    const int jm = tid / RT_L, jl = tid % RT_L;
    const int lB = l0 + jl, mcB = mc0 + jm, mB = mcB * minc;
    const bool vB = lB <= l_max && mcB < n_m && mB <= lB;
    const long long arrB = vB ? (long long)mstart[mcB] + lB - mB : 0;
    // buffer row (f, r) = f * nr + r; array row = f * nr_arr + r_off + r (a part moves a sub-range of the levels of each field)
    auto arr_row = [&](int row) { return (long long)(row / a.nr) * a.nr_arr + a.r_off + row % a.nr; };
    auto buf_row = [&](int row) { return a.direct ? (long long)(row / a.nr) * a.n_r_max + a.r0 + row % a.nr : (long long)row; };
    for (int rb = row0; rb < min(row0 + RT_ROWS, rows_total); rb += 4) {
        double2 v[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int row = rb + q;
            v[q] = make_double2(0.0, 0.0);
            if (row < rows_total) {
                if (UNPACK) { if (vA) v[q] = in[bufA + buf_row(row) * lcA]; }
                else { if (vB) v[q] = in[arrB + arr_row(row) * a.lm_max]; }
            }
        }
#pragma unroll
        for (int q = 0; q < 4; q++) {
            if (UNPACK) tile[q][il][im] = v[q];
            else tile[q][jl][jm] = v[q];
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int row = rb + q;
            if (row < rows_total) {
                if (UNPACK) { if (vB) out[arrB + arr_row(row) * a.lm_max] = tile[q][jl][jm]; }
                else { if (vA) out[bufA + buf_row(row) * lcA] = tile[q][il][im]; }
            }
        }
        __syncthreads();
    }
}

// single rank: the exchange is the identity, so lm2r / r2lm reduce to the lo<->st permutation
__global__ void permute_kernel(const double2 *__restrict__ in, double2 *__restrict__ out, const int *__restrict__ lo2st, int lm_max,
                               long long rows, int to_st) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows * lm_max) return;
    long long row = idx / lm_max;
    int lm = (int)(idx - row * lm_max);
    size_t st = (size_t)row * lm_max + lo2st[lm];
    if (to_st) out[st] = in[idx];
    else out[idx] = in[st];
}

// ---- copy-engine exchange: cross-process flags in IPC-shared device memory -----------------------------------------------
// slot[q] points at MY entry of peer q's flag array (mapped through CUDA IPC); the store is ordered after everything this
// stream did before (the peer copies) and made visible system-wide
__global__ void ce_signal_kernel(unsigned long long *const *slot, int n, int me, unsigned long long seq) {
    const int q = threadIdx.x;
    if (q < n && q != me) {
        __threadfence_system();
        *reinterpret_cast<volatile unsigned long long *>(slot[q]) = seq;
    }
}
// spins until every peer's flag has reached seq; traps instead of hanging when a peer never arrives (20 s)
__global__ void ce_wait_kernel(const volatile unsigned long long *flags, int n, int me, unsigned long long seq) {
    const int p = threadIdx.x;
    if (p < n && p != me) {
        const long long t0 = clock64();
        while (flags[p] < seq) {
            if (clock64() - t0 > 40000000000LL) __trap();
        }
        __threadfence_system();
    }
}

}  // namespace

struct magic_transp {
    magic_sht *h = nullptr;
    int rank = 0, n_procs = 1, n_r_max = 0, n_fields = 0;  // n_fields = widest container this object serves
    std::vector<int> rs, re, ls, le, lo2st;
    std::vector<long long> lm1, lmd1, r1, rd1;  // per-field counts / displacements (complex elements)
    int *d_rstart = nullptr, *d_rcount = nullptr, *d_lstart = nullptr, *d_lcount = nullptr, *d_lo2st = nullptr, *d_st2lo = nullptr;
    long long *d_lmdisp = nullptr, *d_rdisp = nullptr;
    double *sendbuf = nullptr, *recvbuf = nullptr, *stage_lm = nullptr, *stage_r = nullptr;
    ncclComm_t comm = nullptr;
    // copy-engine exchange (MAGIC_TRANSP_CE=1; measured equal to the NCCL exchange at N = 2, so NCCL stays the default): peers'
    // receive buffers mapped into this process;
    // the payload moves with cudaMemcpy2DAsync (DMA engines over NVLink, no SM), arrival / release are flagged in IPC memory
    bool ce = false;
    size_t half = 0;                                   // doubles per half of the double-buffered receive buffer
    std::vector<double *> peer_base;                   // [n_procs] peer q's receive buffer in my address space (own: recvbuf)
    unsigned long long *my_flags = nullptr;            // [2][n_procs] in my receive allocation: arrival flags, release flags
    unsigned long long **d_slot_data = nullptr, **d_slot_ack = nullptr;  // my entries in the peers' flag arrays
    unsigned long long *seq = nullptr;                 // host counter of exchanges, shared by a parent and its parts
    PackArgs args;
    cudaStream_t stream = nullptr;   // stream of the pack / exchange / unpack work (default: the handle's stream)
    magic_transp *parent = nullptr;  // a part shares maps, buffers and communicator with its parent
};

extern "C" int magic_get_blocks(int n_points, int n_procs, int *start, int *stop) {
    if (n_procs < 1 || n_points < n_procs || !start || !stop) MFAIL("magic_get_blocks: bad arguments");
    std::vector<int> s, e;
    get_blocks(n_points, n_procs, s, e);
    for (int p = 0; p < n_procs; p++) { start[p] = s[p]; stop[p] = e[p]; }
    return 0;
}

extern "C" int magic_lo_map(int l_max, int m_max, int minc, int n_procs, int *lo2st, int *lm_start, int *lm_stop) {
    if (l_max < 1 || minc < 1 || m_max < 0 || m_max > l_max || m_max % minc != 0 || n_procs < 1) MFAIL("magic_lo_map: bad arguments");
    magic_sht h;  // host-only shell carrying the truncation
    h.l_max = l_max; h.m_max = m_max; h.minc = minc; h.n_m = m_max / minc + 1;
    h.lstart.assign(h.n_m, 0);
    int lm = 0;
    for (int mc = 0; mc < h.n_m; mc++) { h.lstart[mc] = lm; lm += l_max - mc * minc + 1; }
    h.lm_max = lm;
    if (n_procs > lm) MFAIL("magic_lo_map: more ranks than modes");
    std::vector<int> map, s, e;
    build_lo_map(&h, n_procs, map, s, e);
    for (int i = 0; i < lm; i++) lo2st[i] = map[i];
    for (int p = 0; p < n_procs; p++) { lm_start[p] = s[p]; lm_stop[p] = e[p]; }
    return 0;
}

extern "C" int magic_transp_unique_id(char id[128]) {
    if (nccl_load()) return 1;
    ncclUniqueId uid;
    NCHECK(g_nccl.GetUniqueId(&uid));
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
    memcpy(id, &uid, 128);
    return 0;
}

extern "C" int magic_transp_destroy(magic_transp *t) {
    if (!t) return 0;
    cudaSetDevice(t->h->dev);
    if (t->parent) {  // a part owns only its level tables
        cudaFree(t->d_rstart); cudaFree(t->d_rcount); cudaFree(t->d_lmdisp); cudaFree(t->d_rdisp);
        delete t;
        return 0;
    }
    if (t->ce) {
        for (int p = 0; p < t->n_procs; p++)
            if (p != t->rank && t->peer_base[p]) cudaIpcCloseMemHandle(t->peer_base[p]);
        cudaFree(t->d_slot_data); cudaFree(t->d_slot_ack);
    }
    delete t->seq;
    if (t->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(t->comm);
    cudaFree(t->d_rstart); cudaFree(t->d_rcount); cudaFree(t->d_lstart); cudaFree(t->d_lcount); cudaFree(t->d_lo2st); cudaFree(t->d_st2lo);
    cudaFree(t->d_lmdisp); cudaFree(t->d_rdisp); cudaFree(t->sendbuf); cudaFree(t->recvbuf); cudaFree(t->stage_lm); cudaFree(t->stage_r);
    delete t;
    return 0;
}

// Maps every peer's receive buffer into this process: the IPC handles travel through the NCCL communicator (one 64-byte
// message per peer pair).  Leaves t->ce false (NCCL send/recv stays the exchange) when the devices cannot do peer access.
static int ce_setup(magic_transp *t) {
    const int n = t->n_procs, me = t->rank;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t size");
    cudaIpcMemHandle_t mine;
    if (cudaIpcGetMemHandle(&mine, t->recvbuf) != cudaSuccess) { cudaGetLastError(); return 0; }
    if (getenv("MAGIC_TRANSP_DEBUG")) {  // a recognisable word behind the flag arrays
        unsigned long long tag = 1000 + me;
        cudaMemcpy(reinterpret_cast<unsigned long long *>(t->recvbuf + 2 * t->half) + 2 * n, &tag, 8, cudaMemcpyHostToDevice);
    }
    std::vector<cudaIpcMemHandle_t> all(n);
    double *d_h = nullptr;
    MCHECK(cudaMalloc((void **)&d_h, 64 * (size_t)n));
    MCHECK(cudaMemcpyAsync(d_h + 8 * me, &mine, 64, cudaMemcpyHostToDevice, t->stream));
    NCHECK(g_nccl.GroupStart());
    for (int p = 0; p < n; p++) {
        if (p == me) continue;
        NCHECK(g_nccl.Send(d_h + 8 * me, 8, ncclDouble, p, t->comm, t->stream));
        NCHECK(g_nccl.Recv(d_h + 8 * p, 8, ncclDouble, p, t->comm, t->stream));
    }
    NCHECK(g_nccl.GroupEnd());
    MCHECK(cudaMemcpyAsync(all.data(), d_h, 64 * (size_t)n, cudaMemcpyDeviceToHost, t->stream));
    MCHECK(cudaStreamSynchronize(t->stream));
    cudaFree(d_h);
    t->peer_base.assign(n, nullptr);
    t->peer_base[me] = t->recvbuf;
    bool ok = true;
    for (int p = 0; p < n && ok; p++) {
        if (p == me) continue;
        void *ptr = nullptr;
        if (cudaIpcOpenMemHandle(&ptr, all[p], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = false; break; }
        t->peer_base[p] = (double *)ptr;
    }
    // every rank must take the same path: agree through one more tiny exchange (1.0 = mapped all peers)
    {
        double *d_ok = nullptr;
        MCHECK(cudaMalloc((void **)&d_ok, sizeof(double) * (size_t)n));
        std::vector<double> oks(n, 0.0);
        oks[me] = ok ? 1.0 : 0.0;
        MCHECK(cudaMemcpyAsync(d_ok, oks.data(), sizeof(double) * n, cudaMemcpyHostToDevice, t->stream));
        NCHECK(g_nccl.GroupStart());
        for (int p = 0; p < n; p++) {
            if (p == me) continue;
            NCHECK(g_nccl.Send(d_ok + me, 1, ncclDouble, p, t->comm, t->stream));
            NCHECK(g_nccl.Recv(d_ok + p, 1, ncclDouble, p, t->comm, t->stream));
        }
        NCHECK(g_nccl.GroupEnd());
        MCHECK(cudaMemcpyAsync(oks.data(), d_ok, sizeof(double) * n, cudaMemcpyDeviceToHost, t->stream));
        MCHECK(cudaStreamSynchronize(t->stream));
        cudaFree(d_ok);
        for (int p = 0; p < n; p++) ok = ok && oks[p] == 1.0;
    }
    if (!ok) {
        for (int p = 0; p < n; p++)
            if (p != me && t->peer_base[p]) cudaIpcCloseMemHandle(t->peer_base[p]);
        t->peer_base.clear();
        return 0;
    }
    if (getenv("MAGIC_TRANSP_DEBUG")) {
        for (int p = 0; p < n; p++) {
            cudaPointerAttributes at{};
            cudaError_t e = cudaPointerGetAttributes(&at, t->peer_base[p]);
            int can = -1;
            if (at.device >= 0 && at.device != t->h->dev) cudaDeviceCanAccessPeer(&can, t->h->dev, at.device);
            unsigned long long probe = 0, hh[2];
            memcpy(hh, &all[p], 16);
            cudaError_t e2 = cudaMemcpy(&probe, reinterpret_cast<unsigned long long *>(t->peer_base[p] + 2 * t->half) + 2 * n, 8, cudaMemcpyDefault);
            fprintf(stderr, "[magic_transp rank %d dev %d] peer %d base %p attr rc=%d type=%d device=%d canAccessPeer=%d half=%zu handle=%016llx%016llx probe rc=%d val=%llu\n",
                    me, t->h->dev, p, (void *)t->peer_base[p], (int)e, (int)at.type, at.device, can, t->half, hh[0], hh[1], (int)e2, probe);
        }
        cudaGetLastError();
    }
    std::vector<unsigned long long *> sd(n, nullptr), sa(n, nullptr);
    for (int p = 0; p < n; p++) {
        unsigned long long *pf = reinterpret_cast<unsigned long long *>(t->peer_base[p] + 2 * t->half);
        sd[p] = pf + me;        // arrival flag "from me" at peer p
        sa[p] = pf + n + me;    // release flag "by me" at peer p
    }
    if (dev_upload_vec(&t->d_slot_data, sd) || dev_upload_vec(&t->d_slot_ack, sa)) return 1;
    t->ce = true;
    return 0;
}

extern "C" int magic_transp_create(magic_sht *h, const char id[128], int rank, int n_procs, int n_r_max, int n_fields,
                                   magic_transp **out) {
    if (!h || !out) MFAIL("magic_transp_create: null argument");
    *out = nullptr;
    if (n_procs < 1 || rank < 0 || rank >= n_procs || n_r_max < n_procs || n_fields < 1) MFAIL("magic_transp_create: bad arguments");
    if (n_procs > h->lm_max) MFAIL("magic_transp_create: more ranks than (l,m) modes");
    MCHECK(cudaSetDevice(h->dev));
    magic_transp *t = new magic_transp();
    t->h = h; t->rank = rank; t->n_procs = n_procs; t->n_r_max = n_r_max; t->n_fields = n_fields;
    get_blocks(n_r_max, n_procs, t->rs, t->re);
    build_lo_map(h, n_procs, t->lo2st, t->ls, t->le);
    const int nlm = t->le[rank] - t->ls[rank] + 1, nr = t->re[rank] - t->rs[rank] + 1;
    std::vector<int> rstart(n_procs), rcount(n_procs), lstart(n_procs), lcount(n_procs);
    t->lm1.assign(n_procs, 0); t->lmd1.assign(n_procs + 1, 0);
    t->r1.assign(n_procs, 0); t->rd1.assign(n_procs + 1, 0);
    for (int p = 0; p < n_procs; p++) {
        rstart[p] = t->rs[p] - 1; rcount[p] = t->re[p] - t->rs[p] + 1;
        lstart[p] = t->ls[p] - 1; lcount[p] = t->le[p] - t->ls[p] + 1;
        // create_comm_alltoallv, mpi_transpose.f90:134-139 (per field)
        t->lm1[p] = (long long)rcount[p] * nlm;
        t->r1[p] = (long long)nr * lcount[p];
        t->lmd1[p + 1] = t->lmd1[p] + t->lm1[p];
        t->rd1[p + 1] = t->rd1[p] + t->r1[p];
    }
    std::vector<int> st2lo(h->lm_max);
    for (int i = 0; i < h->lm_max; i++) st2lo[t->lo2st[i]] = i;
    if (dev_upload_vec(&t->d_st2lo, st2lo) || dev_upload_vec(&t->d_rstart, rstart) || dev_upload_vec(&t->d_rcount, rcount) || dev_upload_vec(&t->d_lstart, lstart) ||
        dev_upload_vec(&t->d_lcount, lcount) || dev_upload_vec(&t->d_lo2st, t->lo2st) || dev_upload_vec(&t->d_lmdisp, t->lmd1) ||
        dev_upload_vec(&t->d_rdisp, t->rd1)) {
        magic_transp_destroy(t);
        return 1;
    }
    PackArgs &a = t->args;
    a.n_procs = n_procs; a.n_fields = n_fields; a.n_r_max = n_r_max; a.lm_max = h->lm_max;
    a.llm = t->ls[rank] - 1; a.nlm = nlm; a.r0 = t->rs[rank] - 1; a.nr = nr;
    a.r_off = 0; a.nr_arr = nr;
    t->stream = h->stream;
    a.rstart = t->d_rstart; a.rcount = t->d_rcount; a.lstart = t->d_lstart; a.lcount = t->d_lcount; a.lo2st = t->d_lo2st;
    a.disp = nullptr;
    a.direct = n_procs == 1 ? 1 : 0;
    if (n_procs > 1) {
        size_t maxel = (size_t)std::max(t->lmd1[n_procs], t->rd1[n_procs]) * n_fields;
        MCHECK(cudaMalloc((void **)&t->sendbuf, sizeof(double) * 2 * maxel));
        // receive buffer: two halves (exchange k lands in half k & 1) + the flag words of the copy-engine exchange.  The half
        // size is the same on every rank (the largest need of any rank): peers address my halves and flags with their own copy
        size_t gmax = 0;
        for (int p = 0; p < n_procs; p++)
            gmax = std::max(gmax, std::max((size_t)n_r_max * (size_t)lcount[p], (size_t)rcount[p] * (size_t)h->lm_max));
        t->half = (2 * gmax * n_fields + 31) / 32 * 32;
        MCHECK(cudaMalloc((void **)&t->recvbuf, sizeof(double) * 2 * t->half + 4096));
        MCHECK(cudaMemset(t->recvbuf + 2 * t->half, 0, 4096));
        t->my_flags = reinterpret_cast<unsigned long long *>(t->recvbuf + 2 * t->half);
        t->seq = new unsigned long long(0);
        // id == NULL: no communicator (pack/unpack halves only -- used to test the permutation kernels in one process)
        if (id) {
            if (nccl_load()) { magic_transp_destroy(t); return 1; }
            ncclUniqueId uid;
            memcpy(&uid, id, 128);
            ncclResult_t r = g_nccl.CommInitRank(&t->comm, n_procs, uid, rank);
            if (r != ncclSuccess) {
                g_last_error = std::string("ncclCommInitRank -> ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "error");
                t->comm = nullptr;
                magic_transp_destroy(t);
                return 1;
            }
            const char *e = getenv("MAGIC_TRANSP_CE");
            if (e && atoi(e) == 1 && 2 * n_procs * sizeof(unsigned long long) <= 4096) {
                if (ce_setup(t)) { magic_transp_destroy(t); return 1; }
            }
        }
    }
    *out = t;
    return 0;
}

// A part moves, for every rank q, the level sub-range [lev_off[q], lev_off[q] + lev_cnt[q]) of q's radial slab -- e.g. one
// level chunk of the radial loop -- with its own (smaller) all-to-all, so that the transposes of one chunk can overlap the
// compute of another.  It shares the maps, the staging buffers and the communicator of its parent: use the parts of one
// parent on ONE stream (they serialise on the shared buffers) and issue them in the same order on every rank.
extern "C" int magic_transp_create_part(magic_transp *parent, const int *lev_off, const int *lev_cnt, magic_transp **out) {
    if (!parent || !lev_off || !lev_cnt || !out) MFAIL("magic_transp_create_part: null argument");
    if (parent->parent) MFAIL("magic_transp_create_part: parent is itself a part");
    *out = nullptr;
    MCHECK(cudaSetDevice(parent->h->dev));
    const int n_procs = parent->n_procs, rank = parent->rank;
    magic_transp *t = new magic_transp(*parent);  // shares every pointer; the level tables are replaced below
    t->parent = parent;
    t->d_rstart = t->d_rcount = nullptr;
    t->d_lmdisp = t->d_rdisp = nullptr;
    t->stage_lm = t->stage_r = nullptr;
    const int nlm = t->le[rank] - t->ls[rank] + 1, nr_full = parent->re[rank] - parent->rs[rank] + 1;
    std::vector<int> rstart(n_procs), rcount(n_procs);
    for (int q = 0; q < n_procs; q++) {
        const int nq = parent->re[q] - parent->rs[q] + 1;
        if (lev_off[q] < 0 || lev_cnt[q] < 0 || lev_off[q] + lev_cnt[q] > nq) { delete t; MFAIL("magic_transp_create_part: level range outside a rank's slab"); }
        rstart[q] = parent->rs[q] - 1 + lev_off[q];
        rcount[q] = lev_cnt[q];
        t->rs[q] = rstart[q] + 1;
        t->re[q] = rstart[q] + rcount[q];
    }
    const int nr = rcount[rank];
    for (int p = 0; p < n_procs; p++) {
        const int lcount = t->le[p] - t->ls[p] + 1;
        t->lm1[p] = (long long)rcount[p] * nlm;
        t->r1[p] = (long long)nr * lcount;
        t->lmd1[p + 1] = t->lmd1[p] + t->lm1[p];
        t->rd1[p + 1] = t->rd1[p] + t->r1[p];
    }
    if (dev_upload_vec(&t->d_rstart, rstart) || dev_upload_vec(&t->d_rcount, rcount) || dev_upload_vec(&t->d_lmdisp, t->lmd1) ||
        dev_upload_vec(&t->d_rdisp, t->rd1)) {
        magic_transp_destroy(t);
        return 1;
    }
    PackArgs &a = t->args;
    a.rstart = t->d_rstart; a.rcount = t->d_rcount;
    a.r0 = rstart[rank]; a.nr = nr; a.r_off = lev_off[rank]; a.nr_arr = nr_full;
    *out = t;
    return 0;
}

extern "C" int magic_transp_info(const magic_transp *t, int *rank, int *n_procs, int *n_r_max, int *n_fields) {
    if (!t) MFAIL("null transposer");
    if (rank) *rank = t->rank;
    if (n_procs) *n_procs = t->n_procs;
    if (n_r_max) *n_r_max = t->n_r_max;
    if (n_fields) *n_fields = t->n_fields;
    return 0;
}

extern "C" int magic_transp_set_stream(magic_transp *t, void *stream) {
    if (!t) MFAIL("null transposer");
    t->stream = stream ? (cudaStream_t)stream : t->h->stream;
    return 0;
}

extern "C" int magic_transp_extents(const magic_transp *t, int *llm, int *ulm, int *nRstart, int *nRstop) {
    if (!t) MFAIL("null transposer");
    *llm = t->ls[t->rank]; *ulm = t->le[t->rank]; *nRstart = t->rs[t->rank]; *nRstop = t->re[t->rank];
    return 0;
}

extern "C" int magic_transp_counts(const magic_transp *t, int dir, long long *scounts, long long *sdisp, long long *rcounts,
                                   long long *rdisp) {
    if (!t) MFAIL("null transposer");
    const auto &sc = dir == 0 ? t->lm1 : t->r1, &sd = dir == 0 ? t->lmd1 : t->rd1;
    const auto &rc = dir == 0 ? t->r1 : t->lm1, &rd = dir == 0 ? t->rd1 : t->lmd1;
    const long long nf = t->n_fields;
    for (int p = 0; p < t->n_procs; p++) { scounts[p] = nf * sc[p]; sdisp[p] = nf * sd[p]; rcounts[p] = nf * rc[p]; rdisp[p] = nf * rd[p]; }
    return 0;
}

static int side_launch(magic_transp *t, int nf, bool lmside, const double *arr_in, double *arr_out, const double *buf_in, double *buf_out) {
    PackArgs a = t->args;
    a.n_fields = nf;
    a.disp = lmside ? t->d_lmdisp : t->d_rdisp;
    long long total = (lmside ? t->lmd1[t->n_procs] : t->rd1[t->n_procs]) * nf;
    if (total == 0) return 0;
    int blocks = (int)((total + 255) / 256);
    if (lmside)
        lmside_kernel<<<blocks, 256, 0, t->stream>>>(a, (const double2 *)arr_in, (double2 *)arr_out, (const double2 *)buf_in, (double2 *)buf_out, total);
    else {
        const magic_sht *h = t->h;
        const int rows = nf * a.nr;
        if (rows == 0) return 0;
        dim3 grid((h->l_max + RT_L) / RT_L, (h->n_m + RT_M - 1) / RT_M, (rows + RT_ROWS - 1) / RT_ROWS);
        if (buf_out)
            rside_tiled_kernel<false><<<grid, 256, 0, t->stream>>>(a, t->d_st2lo, h->d_lstart, (const double2 *)arr_in, (double2 *)buf_out, rows,
                                                                  h->minc, h->l_max, h->n_m);
        else
            rside_tiled_kernel<true><<<grid, 256, 0, t->stream>>>(a, t->d_st2lo, h->d_lstart, (const double2 *)buf_in, (double2 *)arr_out, rows,
                                                                 h->minc, h->l_max, h->n_m);
        (void)blocks;
    }
    t->h->launches++;
    MCHECK(cudaGetLastError());
    return 0;
}

#define TCHK(t, nf)                                                                  \
    if (!(t)) MFAIL("null transposer");                                              \
    if ((nf) < 1 || (nf) > (t)->n_fields) MFAIL("transposer: n_fields out of range"); \
    MCHECK(cudaSetDevice((t)->h->dev));

extern "C" int magic_transp_pack_lm2r_dev(magic_transp *t, const double *arr_LMloc, double *sendbuf) {
    TCHK(t, t->n_fields);
    return side_launch(t, t->n_fields, true, arr_LMloc, nullptr, nullptr, sendbuf);
}
extern "C" int magic_transp_unpack_lm2r_dev(magic_transp *t, const double *recvbuf, double *arr_Rloc) {
    TCHK(t, t->n_fields);
    return side_launch(t, t->n_fields, false, nullptr, arr_Rloc, recvbuf, nullptr);
}
extern "C" int magic_transp_pack_r2lm_dev(magic_transp *t, const double *arr_Rloc, double *sendbuf) {
    TCHK(t, t->n_fields);
    return side_launch(t, t->n_fields, false, arr_Rloc, nullptr, nullptr, sendbuf);
}
extern "C" int magic_transp_unpack_r2lm_dev(magic_transp *t, const double *recvbuf, double *arr_LMloc) {
    TCHK(t, t->n_fields);
    return side_launch(t, t->n_fields, true, nullptr, arr_LMloc, recvbuf, nullptr);
}

// single rank: the packed buffer of the one segment IS arr_LMloc ([f][n_r][lm_lo]), so lm2r / r2lm are one tiled
// lo<->st permutation without staging
static int permute_launch(magic_transp *t, int nf, const double *in, double *out, int to_st) {
    if (to_st) return side_launch(t, nf, false, nullptr, out, in, nullptr);   // unpack: buffer(lo) -> arr_R(st)
    return side_launch(t, nf, false, in, nullptr, nullptr, out);              // pack:   arr_R(st) -> buffer(lo)
}

// The LM side needs no packing: in arr_LMloc ([f][n_r_max][nlm]) the levels of rank q's slab (or of q's part of it) are
// contiguous rows of every field, and the R-side buffers keep a rank's segment field-major ([f][levels][modes]).  So every
// (peer, field) block is sent from / received into arr_LMloc directly -- the packing copy of mpi_transpose.f90:320-333 and
// the unpacking copy of :515-528 (2 x the container through HBM per transpose) do not exist here.
// Copy-engine form of the exchange.  Exchange number k of this communicator lands in half k & 1 of the receivers' buffers:
//   wait until every peer has released what exchange k-2 put there -> peer copies (2-D DMA, one per peer) -> raise my arrival
//   flag at every peer -> wait for all arrival flags -> [caller unpacks] -> ce_release raises my release flag at every peer.
// Returns the half that holds the received data.
static int ce_exchange(magic_transp *t, int nf, const double *arr_LM_send, bool r2lm, double **recv_half) {
    cudaStream_t st = t->stream;
    const int me = t->rank, n = t->n_procs;
    const unsigned long long seq = ++(*t->seq);
    const size_t hoff = (seq & 1ULL) * t->half;
    *recv_half = t->recvbuf + hoff;
    const size_t nlm = (size_t)(t->le[me] - t->ls[me] + 1), frow = nlm * t->n_r_max;
    if (seq > 2) ce_wait_kernel<<<1, n, 0, st>>>(t->my_flags + n, n, me, seq - 2);
    const size_t c16 = sizeof(double) * 2;
    if (!r2lm) {
        // my modes of q's levels -> q's buffer, segment "from me": [f][levels of q][my modes], contiguous
        for (int q = 0; q < n; q++) {
            if (t->lm1[q] == 0) continue;
            const size_t nr_q = (size_t)(t->re[q] - t->rs[q] + 1);
            const size_t dst = 2 * (size_t)nf * nr_q * (size_t)(t->ls[me] - t->ls[0]);  // nf * rd1 as rank q computes it
            MCHECK(cudaMemcpy2DAsync(t->peer_base[q] + hoff + dst, c16 * t->lm1[q], arr_LM_send + 2 * (size_t)(t->rs[q] - 1) * nlm, c16 * frow,
                                     c16 * t->lm1[q], nf, cudaMemcpyDefault, st));
        }
    } else {
        // my levels of p's modes (packed in sendbuf, segment p: [f][my levels][modes of p]) -> p's buffer, segment "from me"
        size_t lev_before = 0;  // levels of the ranks before me (of this part)
        for (int q = 0; q < me; q++) lev_before += (size_t)(t->re[q] - t->rs[q] + 1);
        for (int p = 0; p < n; p++) {
            if (t->r1[p] == 0) continue;
            const size_t nlm_p = (size_t)(t->le[p] - t->ls[p] + 1);
            const size_t dst = 2 * (size_t)nf * lev_before * nlm_p;  // nf * lmd1 as rank p computes it
            MCHECK(cudaMemcpyAsync(t->peer_base[p] + hoff + dst, t->sendbuf + 2 * (size_t)nf * t->rd1[p], c16 * (size_t)nf * t->r1[p],
                                   cudaMemcpyDefault, st));
        }
    }
    ce_signal_kernel<<<1, n, 0, st>>>(t->d_slot_data, n, me, seq);
    ce_wait_kernel<<<1, n, 0, st>>>(t->my_flags, n, me, seq);
    MCHECK(cudaGetLastError());
    return 0;
}
static int ce_release(magic_transp *t) {
    ce_signal_kernel<<<1, t->n_procs, 0, t->stream>>>(t->d_slot_ack, t->n_procs, t->rank, *t->seq);
    MCHECK(cudaGetLastError());
    return 0;
}

static int exchange_lm_direct(magic_transp *t, int nf, const double *arr_LM_send, double *arr_LM_recv) {
    cudaStream_t st = t->stream;
    const int me = t->rank, n = t->n_procs;
    if (!t->comm) MFAIL("transposer was created without an NCCL id: only the pack/unpack halves are available");
    const size_t nlm = (size_t)(t->le[me] - t->ls[me] + 1), frow = nlm * t->n_r_max;  // complex elements per field
    auto lm_block = [&](int q, int f) { return 2 * ((size_t)f * frow + (size_t)(t->rs[q] - 1) * nlm); };  // doubles
    // own segment: a strided device copy
    if (t->lm1[me] > 0) {
        if (arr_LM_send)
            MCHECK(cudaMemcpy2DAsync(t->recvbuf + 2 * (size_t)nf * t->rd1[me], sizeof(double) * 2 * t->r1[me], arr_LM_send + lm_block(me, 0),
                                     sizeof(double) * 2 * frow, sizeof(double) * 2 * t->lm1[me], nf, cudaMemcpyDeviceToDevice, st));
        else
            MCHECK(cudaMemcpy2DAsync(arr_LM_recv + lm_block(me, 0), sizeof(double) * 2 * frow, t->sendbuf + 2 * (size_t)nf * t->rd1[me],
                                     sizeof(double) * 2 * t->r1[me], sizeof(double) * 2 * t->lm1[me], nf, cudaMemcpyDeviceToDevice, st));
    }
    NCHECK(g_nccl.GroupStart());
    for (int p = 0; p < n; p++) {
        if (p == me) continue;
        for (int f = 0; f < nf; f++) {
            if (arr_LM_send) {  // lm2r: my modes of p's levels go out, p's modes of my levels come in
                if (t->lm1[p] > 0) NCHECK(g_nccl.Send(arr_LM_send + lm_block(p, f), (size_t)(2 * t->lm1[p]), ncclDouble, p, t->comm, st));
                if (t->r1[p] > 0) NCHECK(g_nccl.Recv(t->recvbuf + 2 * ((size_t)nf * t->rd1[p] + (size_t)f * t->r1[p]), (size_t)(2 * t->r1[p]), ncclDouble, p, t->comm, st));
            } else {            // r2lm: the mirror image
                if (t->r1[p] > 0) NCHECK(g_nccl.Send(t->sendbuf + 2 * ((size_t)nf * t->rd1[p] + (size_t)f * t->r1[p]), (size_t)(2 * t->r1[p]), ncclDouble, p, t->comm, st));
                if (t->lm1[p] > 0) NCHECK(g_nccl.Recv(arr_LM_recv + lm_block(p, f), (size_t)(2 * t->lm1[p]), ncclDouble, p, t->comm, st));
            }
        }
    }
    NCHECK(g_nccl.GroupEnd());
    return 0;
}

extern "C" int magic_transp_lm2r_dev_n(magic_transp *t, int nf, const double *arr_LMloc, double *arr_Rloc) {
    TCHK(t, nf);
    if (t->n_procs == 1) return permute_launch(t, nf, arr_LMloc, arr_Rloc, 1);
    if (t->ce) {
        double *rh = nullptr;
        if (ce_exchange(t, nf, arr_LMloc, false, &rh)) return 1;
        if (side_launch(t, nf, false, nullptr, arr_Rloc, rh, nullptr)) return 1;
        return ce_release(t);
    }
    if (exchange_lm_direct(t, nf, arr_LMloc, nullptr)) return 1;
    return side_launch(t, nf, false, nullptr, arr_Rloc, t->recvbuf, nullptr);
}

extern "C" int magic_transp_r2lm_dev_n(magic_transp *t, int nf, const double *arr_Rloc, double *arr_LMloc) {
    TCHK(t, nf);
    if (t->n_procs == 1) return permute_launch(t, nf, arr_Rloc, arr_LMloc, 0);
    if (side_launch(t, nf, false, arr_Rloc, nullptr, nullptr, t->sendbuf)) return 1;
    if (t->ce) {
        double *rh = nullptr;
        if (ce_exchange(t, nf, nullptr, true, &rh)) return 1;
        // LM side: segment q of the buffer ([f][levels of q][my modes]) goes to the rows of q's levels of every field
        const int me = t->rank;
        const size_t nlm = (size_t)(t->le[me] - t->ls[me] + 1), frow = nlm * t->n_r_max, c16 = sizeof(double) * 2;
        for (int q = 0; q < t->n_procs; q++) {
            if (t->lm1[q] == 0) continue;
            MCHECK(cudaMemcpy2DAsync(arr_LMloc + 2 * (size_t)(t->rs[q] - 1) * nlm, c16 * frow, rh + 2 * (size_t)nf * t->lmd1[q], c16 * t->lm1[q],
                                     c16 * t->lm1[q], nf, cudaMemcpyDeviceToDevice, t->stream));
        }
        return ce_release(t);
    }
    return exchange_lm_direct(t, nf, nullptr, arr_LMloc);
}

extern "C" int magic_transp_lm2r_dev(magic_transp *t, const double *arr_LMloc, double *arr_Rloc) {
    if (!t) MFAIL("null transposer");
    return magic_transp_lm2r_dev_n(t, t->n_fields, arr_LMloc, arr_Rloc);
}
extern "C" int magic_transp_r2lm_dev(magic_transp *t, const double *arr_Rloc, double *arr_LMloc) {
    if (!t) MFAIL("null transposer");
    return magic_transp_r2lm_dev_n(t, t->n_fields, arr_Rloc, arr_LMloc);
}

static int ensure_stage(magic_transp *t) {
    if (t->parent) MFAIL("the host-pointer transposes are not available on a part (use the parent, or the device-pointer calls)");
    if (t->stage_lm) return 0;
    MCHECK(cudaMalloc((void **)&t->stage_lm, sizeof(double) * 2 * (size_t)t->lmd1[t->n_procs] * t->n_fields));
    MCHECK(cudaMalloc((void **)&t->stage_r, sizeof(double) * 2 * (size_t)t->rd1[t->n_procs] * t->n_fields));
    return 0;
}

extern "C" int magic_transp_lm2r(magic_transp *t, const double *arr_LMloc, double *arr_Rloc) {
    TCHK(t, t->n_fields);
    if (ensure_stage(t)) return 1;
    size_t blm = sizeof(double) * 2 * (size_t)t->lmd1[t->n_procs] * t->n_fields, br = sizeof(double) * 2 * (size_t)t->rd1[t->n_procs] * t->n_fields;
    MCHECK(cudaMemcpyAsync(t->stage_lm, arr_LMloc, blm, cudaMemcpyHostToDevice, t->stream));  // same stream as pack / exchange / unpack
    if (magic_transp_lm2r_dev(t, t->stage_lm, t->stage_r)) return 1;
    MCHECK(cudaMemcpyAsync(arr_Rloc, t->stage_r, br, cudaMemcpyDeviceToHost, t->stream));
    MCHECK(cudaStreamSynchronize(t->stream));
    return 0;
}

extern "C" int magic_transp_r2lm(magic_transp *t, const double *arr_Rloc, double *arr_LMloc) {
    TCHK(t, t->n_fields);
    if (ensure_stage(t)) return 1;
    size_t blm = sizeof(double) * 2 * (size_t)t->lmd1[t->n_procs] * t->n_fields, br = sizeof(double) * 2 * (size_t)t->rd1[t->n_procs] * t->n_fields;
    MCHECK(cudaMemcpyAsync(t->stage_r, arr_Rloc, br, cudaMemcpyHostToDevice, t->stream));
    if (magic_transp_r2lm_dev(t, t->stage_r, t->stage_lm)) return 1;
    MCHECK(cudaMemcpyAsync(arr_LMloc, t->stage_lm, blm, cudaMemcpyDeviceToHost, t->stream));
    MCHECK(cudaStreamSynchronize(t->stream));
    return 0;
}
