// Microbenchmarks that fix the design points of the fused lift+Gram engine on B200 (sm_100a):
//  (1) DMMA.8x8x4 register-only issue rate vs warps/SM  -> FP64 tensor peak actually reachable
//  (2) DFMA rate (same pipe?)                            -> is DMMA any faster than plain FMA
//  (3) red.global.add.f64 / ld+add+st throughput on 128 KB tiles -> cost of the Gram epilogue
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o dmma_peak dmma_peak.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} }while(0)

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int NACC>
__global__ void k_dmma(double *out, int iters, double a0, double b0) {
    double c[NACC][2];
#pragma unroll
    for (int i = 0; i < NACC; i++) { c[i][0] = 0; c[i][1] = 0; }
    double a = a0 + threadIdx.x * 1e-9, b = b0;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) dmma(c[i][0], c[i][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void k_dfma(double *out, int iters, double a0, double b0) {
    double c[NACC];
#pragma unroll
    for (int i = 0; i < NACC; i++) c[i] = i;
    double a = a0 + threadIdx.x * 1e-9, b = b0;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) c[i] = fma(c[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// epilogue candidates: every CTA adds a 128 KB register tile (64 doubles/thread, 256 threads) into global memory
__global__ void k_red(double *g, int tiles_per_cta, int layout) {
    double v = threadIdx.x * 1e-3;
    for (int t = 0; t < tiles_per_cta; t++) {
        double *base = g + ((size_t)(blockIdx.x * tiles_per_cta + t)) * 16384;
        int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
        for (int j = 0; j < 32; j++) {
            double *p = base + (warp * 32 + j) * 64;
            if (layout == 0) { // [tile][lane][2]
                asm volatile("red.global.add.f64 [%0], %1;" ::"l"(p + lane * 2), "d"(v));
                asm volatile("red.global.add.f64 [%0], %1;" ::"l"(p + lane * 2 + 1), "d"(v));
            } else {           // [tile][2][lane]
                asm volatile("red.global.add.f64 [%0], %1;" ::"l"(p + lane), "d"(v));
                asm volatile("red.global.add.f64 [%0], %1;" ::"l"(p + 32 + lane), "d"(v));
            }
        }
    }
}
__global__ void k_rmw(double *g, int tiles_per_cta) {
    double v = threadIdx.x * 1e-3;
    for (int t = 0; t < tiles_per_cta; t++) {
        double2 *base = (double2 *)(g + ((size_t)(blockIdx.x * tiles_per_cta + t)) * 16384);
        int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
        for (int jb = 0; jb < 32; jb += 8) {
            double2 r[8];
#pragma unroll
            for (int j = 0; j < 8; j++) r[j] = base[(warp * 32 + jb + j) * 32 + lane];
#pragma unroll
            for (int j = 0; j < 8; j++) { r[j].x += v; r[j].y += v; base[(warp * 32 + jb + j) * 32 + lane] = r[j]; }
        }
    }
}

template <typename F>
float timeit(F f, int rep = 3) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < rep; r++) {
        cudaEventRecord(e0); f(); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    int sms = prop.multiProcessorCount;
    printf("device %s sms %d clock %d kHz\n", prop.name, sms, prop.clockRate);
    double *out; CK(cudaMalloc(&out, sizeof(double) * sms * 8 * 1024));
    int iters = 20000;
    for (int warps : {4, 8, 16, 32}) {
        for (int ctas : {1, 2}) {
            if (warps * ctas > 32 && warps * ctas != 64) {}
            float ms = timeit([&] { k_dmma<32><<<sms * ctas, warps * 32>>>(out, iters, 1.0, 1e-9); });
            double flops = 2.0 * 256 * 32 * (double)iters * warps * ctas * sms;
            printf("DMMA  nacc=32 warps/CTA=%2d ctas/SM=%d : %.3f ms  %.2f TFLOP/s  (%.2f FMA/clk/SM @1.965GHz)\n", warps, ctas, ms,
                   flops / ms * 1e-9, flops / 2 / (ms * 1e-3) / sms / 1.965e9);
        }
    }
    for (int warps : {4, 8}) {
        float ms = timeit([&] { k_dmma<8><<<sms, warps * 32>>>(out, iters, 1.0, 1e-9); });
        double flops = 2.0 * 256 * 8 * (double)iters * warps * sms;
        printf("DMMA  nacc=8  warps/CTA=%2d : %.3f ms  %.2f TFLOP/s\n", warps, ms, flops / ms * 1e-9);
        ms = timeit([&] { k_dmma<2><<<sms, warps * 32>>>(out, iters, 1.0, 1e-9); });
        flops = 2.0 * 256 * 2 * (double)iters * warps * sms;
        printf("DMMA  nacc=2  warps/CTA=%2d : %.3f ms  %.2f TFLOP/s  (latency probe: %.1f clk per dependent pair @1.965GHz)\n", warps, ms,
               flops / ms * 1e-9, ms * 1e-3 * 1.965e9 / iters);
    }
    for (int warps : {8, 16, 32}) {
        float ms = timeit([&] { k_dfma<32><<<sms, warps * 32>>>(out, iters, 1.0000001, 1e-9); });
        double flops = 2.0 * 32 * 32 * (double)iters * warps * sms;
        printf("DFMA  nacc=32 warps/CTA=%2d : %.3f ms  %.2f TFLOP/s\n", warps, ms, flops / ms * 1e-9);
    }
    // long sustained run (power/clock settle): ~3 s of DMMA
    {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        int reps = 0;
        for (; reps < 40; reps++) k_dmma<32><<<sms * 2, 256>>>(out, iters * 4, 1.0, 1e-9);
        cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double flops = 2.0 * 256 * 32 * (double)iters * 4 * 8 * 2 * sms * reps;
        printf("DMMA sustained %.1f ms : %.2f TFLOP/s\n", ms, flops / ms * 1e-9);
    }
    // epilogue probes: 148 CTAs x 14 tiles x 128 KB = 272 MB (one chunk's worth of Gram tiles)
    double *g; size_t gbytes = (size_t)sms * 14 * 16384 * 8; CK(cudaMalloc(&g, gbytes)); CK(cudaMemset(g, 0, gbytes));
    for (int layout = 0; layout < 2; layout++) {
        float ms = timeit([&] { k_red<<<sms, 256>>>(g, 14, layout); });
        printf("RED.f64 layout=%d : %.3f ms for %.1f MB -> %.1f GB/s of accumulator bytes\n", layout, ms, gbytes / 1e6, gbytes / ms * 1e-6);
    }
    {
        float ms = timeit([&] { k_rmw<<<sms, 256>>>(g, 14); });
        printf("LD+ADD+ST v2   : %.3f ms for %.1f MB -> %.1f GB/s of accumulator bytes\n", ms, gbytes / 1e6, gbytes / ms * 1e-6);
    }
    printf("done\n");
    return 0;
}
