// Microbenchmark of the 64x32 warp-tile DMMA main loop fed from shared memory (no TMA, no producer):
// how much of the 37.1 TFLOP/s register-only DMMA rate survives the LDS.128 fragment traffic and the per-slab barriers?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mainloop_probe mainloop_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} }while(0)
__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ double2 lds_v2(uint32_t addr) { double2 v; asm("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr)); return v; }
__device__ __forceinline__ uint32_t mbar_test(uint32_t bar, uint32_t parity) {
    uint32_t ok; asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory"); return ok;
}

template <int MODE>   // 0: registers only, 1: LDS-fed i-major, 2: LDS-fed + a (pre-completed) mbarrier test per slab, 3: LDS-fed, all x then all y
__global__ void __launch_bounds__(256, 1) k_main(double *out, int nslabs) {
    extern __shared__ __align__(1024) unsigned char sm[];
    __shared__ uint64_t bar;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, wr = warp >> 2, wc = warp & 3;
    double *smd = (double *)sm;
    for (int i = tid; i < 3 * 4096; i += 256) smd[i] = 1e-3 * (i % 97);
    if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&bar))); }
    __syncthreads();
    if (tid == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((uint32_t)__cvta_generic_to_shared(&bar)) : "memory");
    __syncthreads();
    const uint32_t sm0 = (uint32_t)__cvta_generic_to_shared(sm);
    const uint32_t a_off = sm0 + (wr * 16) * 512 + lane * 16, b_off = sm0 + 16384 + (wc * 8) * 512 + lane * 16;
    const uint32_t barA = (uint32_t)__cvta_generic_to_shared(&bar);
    double acc[8][4][2];
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) { acc[i][j][0] = 0; acc[i][j][1] = 0; }
    double2 a[8], b[4];
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = lds_v2(a_off + i * 1024);
#pragma unroll
    for (int j = 0; j < 4; j++) b[j] = lds_v2(b_off + j * 1024);
    int stage = 0;
    for (int s = 0; s < nslabs; s++) {
        const uint32_t so = stage * 32768;
        if (MODE == 2) { while (!mbar_test(barA, 0)) {} }
#pragma unroll
        for (int q = 0; q < 2; q++) {
            if (MODE != 0) {
#pragma unroll
                for (int i = 0; i < 8; i++) a[i] = lds_v2(a_off + so + q * 512 + i * 1024);
#pragma unroll
                for (int j = 0; j < 4; j++) b[j] = lds_v2(b_off + so + q * 512 + j * 1024);
            }
            if (MODE == 3) {
#pragma unroll
                for (int i = 0; i < 8; i++)
#pragma unroll
                    for (int j = 0; j < 4; j++) dmma(acc[i][j][0], acc[i][j][1], a[i].x, b[j].x);
#pragma unroll
                for (int i = 0; i < 8; i++)
#pragma unroll
                    for (int j = 0; j < 4; j++) dmma(acc[i][j][0], acc[i][j][1], a[i].y, b[j].y);
            } else {
#pragma unroll
                for (int i = 0; i < 8; i++) {
#pragma unroll
                    for (int j = 0; j < 4; j++) dmma(acc[i][j][0], acc[i][j][1], a[i].x, b[j].x);
#pragma unroll
                    for (int j = 0; j < 4; j++) dmma(acc[i][j][0], acc[i][j][1], a[i].y, b[j].y);
                }
            }
        }
        if (MODE == 2) { __syncwarp(); }
        if (++stage == 3) stage = 0;
    }
    double sum = 0;
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) sum += acc[i][j][0] + acc[i][j][1];
    out[blockIdx.x * 256 + tid] = sum;
}

template <int MODE> void run(double *out, int sms, const char *name) {
    CK(cudaFuncSetAttribute(k_main<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * 32768));
    const int nslabs = 20000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_main<MODE><<<sms, 256, 3 * 32768>>>(out, nslabs); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 3; r++) {
        cudaEventRecord(e0); k_main<MODE><<<sms, 256, 3 * 32768>>>(out, nslabs); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    double flops = 2.0 * 128 * 128 * 16 * (double)nslabs * sms;
    printf("%-58s %.3f ms  %.2f TFLOP/s\n", name, best, flops / best * 1e-9);
}
int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    int sms = prop.multiProcessorCount;
    double *out; CK(cudaMalloc(&out, sizeof(double) * sms * 256));
    run<0>(out, sms, "registers only (8 warps, 64x32 warp tile)");
    run<1>(out, sms, "LDS.128-fed, per-row x/y order");
    run<3>(out, sms, "LDS.128-fed, all-x then all-y order");
    run<2>(out, sms, "LDS.128-fed + mbarrier try_wait per slab");
    return 0;
}
