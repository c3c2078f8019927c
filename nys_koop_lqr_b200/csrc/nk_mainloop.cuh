// nk_mainloop.cuh -- the DMMA main loop shared by the fused lift+Gram engine (nk_gram.cu) and the packed-operand
// persistent GEMM (nk_pgemm.cu): raw shared-address mbarrier / LDS helpers and one k8 step of the 64x32 warp tile.
#pragma once
#include "nk_common.cuh"

namespace nk {

// ------------------------------------------------------------------------------------------------
// raw 32-bit shared-address helpers (addresses are computed once; the generic->shared conversion of a pointer
// costs an S2R + LEA each time the compiler rematerialises it inside the slab loop)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_arrive_a(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint32_t mbar_test_a(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok;
}
__device__ __forceinline__ void mbar_wait_a(uint32_t bar, uint32_t parity) {
    while (!mbar_test_a(bar, parity)) {}
}
__device__ __forceinline__ double2 lds_v2(uint32_t addr) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_v2(uint32_t addr, double a, double b) {
    asm volatile("st.shared.v2.f64 [%0], {%1,%2};" ::"r"(addr), "d"(a), "d"(b) : "memory");
}
__device__ __forceinline__ uint32_t opaque(uint32_t v) { asm volatile("" : "+r"(v)); return v; }

// one k8 step (two DMMA k4 steps) of the 64x32 warp tile; a0 and b[] were prefetched, a[1..7] are loaded here
__device__ __forceinline__ void k8_step(double (&acc)[8][4][2], uint32_t aq, double2 a0, const double2 (&b)[4]) {
    double2 a[8];
    a[0] = a0;
#pragma unroll
    for (int i = 1; i < 8; i++) a[i] = lds_v2(aq + i * 1024);
#pragma unroll
    for (int i = 0; i < 8; i++) {
#pragma unroll
        for (int j = 0; j < 4; j++) dmma884(acc[i][j][0], acc[i][j][1], a[i].x, b[j].x);
#pragma unroll
        for (int j = 0; j < 4; j++) dmma884(acc[i][j][0], acc[i][j][1], a[i].y, b[j].y);
    }
}


}  // namespace nk
