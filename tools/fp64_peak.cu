// FP64 peak probe for B200 (sm_100a): the Legendre stage of the radial loop is FP64-pipe bound, and
// MEASURED_PEAKS.json carries no FP64 figure. This measures the denominators DESIGN.md quotes:
//   (1) DFMA register loop, (2) DMMA mma.sync m8n8k4, (3) DMMA m16n8k16 (sm_90+ shape).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_peak tools/fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

template <int NACC>
__global__ void dfma_kernel(double* out, int iters, double a, double b) {
    double acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; i++) acc[i] = threadIdx.x * 1e-9 + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NT>
__global__ void dmma884_kernel(double* out, int iters, double a, double b) {
    double c[NT][2];
#pragma unroll
    for (int i = 0; i < NT; i++) { c[i][0] = threadIdx.x * 1e-9; c[i][1] = i; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NT; i++) {
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NT; i++) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NT>
__global__ void dmma16816_kernel(double* out, int iters, double a, double b) {
    double c[NT][4];
#pragma unroll
    for (int i = 0; i < NT; i++) { c[i][0] = threadIdx.x * 1e-9; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NT; i++) {
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                         : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                         : "d"(a), "d"(a), "d"(a), "d"(a), "d"(a), "d"(a), "d"(a), "d"(a), "d"(b), "d"(b), "d"(b), "d"(b));
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NT; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static float time_it(F launch, int reps) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; i++) launch();
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < reps; r++) {
        cudaEventRecord(e0);
        launch();
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int nsm = p.multiProcessorCount;
    printf("{\"device\": \"%s\", \"sms\": %d, \"cc\": \"%d.%d\"", p.name, nsm, p.major, p.minor);
    double* out; CK(cudaMalloc(&out, sizeof(double) * nsm * 8 * 1024));
    const int iters = 20000;
    for (int tpb : {256, 512, 1024}) {
        for (int bps : {1, 2, 4}) {
            if (tpb * bps > 2048) continue;
            int grid = nsm * bps;
            float ms = time_it([&] { dfma_kernel<16><<<grid, tpb>>>(out, iters, 1.0000001, 1e-9); }, 5);
            double fl = 2.0 * 16 * iters * (double)grid * tpb;
            printf(", \"dfma_t%d_b%d_tflops\": %.2f", tpb, bps, fl / ms * 1e-9);
            ms = time_it([&] { dmma884_kernel<8><<<grid, tpb>>>(out, iters, 1.0000001, 1e-9); }, 5);
            fl = 2.0 * 8 * 8 * 4 * 8 * iters * (double)grid * (tpb / 32);
            printf(", \"dmma884_t%d_b%d_tflops\": %.2f", tpb, bps, fl / ms * 1e-9);
            ms = time_it([&] { dmma16816_kernel<4><<<grid, tpb>>>(out, iters / 4, 1.0000001, 1e-9); }, 5);
            fl = 2.0 * 16 * 8 * 16 * 4 * (iters / 4) * (double)grid * (tpb / 32);
            printf(", \"dmma16816_t%d_b%d_tflops\": %.2f", tpb, bps, fl / ms * 1e-9);
        }
    }
    CK(cudaGetLastError());
    printf("}\n");
    return 0;
}
