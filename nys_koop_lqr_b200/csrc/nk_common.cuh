// nk_common.cuh -- sm_100a building blocks shared by every kernel of the Nystrom-Koopman path.
//
// FP64 on Blackwell: there is no FP64 tcgen05/TMEM path; the FP64 tensor instruction is the warp-level
// mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4).  Measured on this pool's B200 (tools/microbench, profiles/):
// 64 FMA/clk/SM = 37.1 TFLOP/s at 1965 MHz, one DMMA per 16 clk per SM sub-partition, identical to the
// DFMA rate.  Operand staging uses the Blackwell async-copy engine (cp.async.bulk -> SASS UBLKCP) with
// mbarrier transaction counts; accumulators live in registers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace nk {

constexpr int kTile = 128;        // CTA output tile (rows and cols)
constexpr int kSlabK = 16;        // contraction depth of one staged slab
constexpr int kPanel = 8;         // rows per packed panel (= DMMA m / n)
constexpr int kBlk = 64;          // doubles per packed 8x8 block
constexpr int kSlabTileDoubles = (kTile / kPanel) * (kSlabK / 8) * kBlk;  // 2048 doubles = 16 KB
constexpr int kConsumerWarps = 8; // 2 (rows) x 4 (cols), warp tile 64 x 32
constexpr int kThreads = (kConsumerWarps + 4) * 32;  // + 1 producer warpgroup (only its first warp works)

// ---- packed operand layout -------------------------------------------------------------------------
// A matrix operand M (R rows x K depth) that is contracted over K is stored as 8x8 row-major blocks,
// ordered [k16 slab][row panel][k8 block]:   off(r,k) = ((k/16)*RP + r/8)*128 + ((k%16)/8)*64 + (r%8)*8 + k%8
// (RP = number of 8-row panels, a multiple of 16).  A 128-row CTA tile of one slab is then one contiguous
// 16 KB run (a single bulk copy), and lane (g,t) of a warp reads its DMMA fragments for two consecutive k4
// steps with ONE conflict-free 16-byte shared load at block + lane*16:  .x -> k = 2t, .y -> k = 2t+1.
// The assignment of contraction indices to (step, t) slots is the same for both operands, which is all
// the contraction needs.  It also equals the C-fragment layout (row g, cols 2t,2t+1), so a kernel lift
// tile is written back to the packed feature buffer with one coalesced 16-byte store per lane.
__host__ __device__ inline size_t packed_off(int r, int k, int row_panels) {
    return ((size_t)(k >> 4) * row_panels + (r >> 3)) * 128 + ((k & 15) >> 3) * 64 + (r & 7) * 8 + (k & 7);
}

// Kernel attributes (opt-in dynamic shared memory) are per DEVICE: remember per device ordinal which kernels were configured,
// so that one process driving several GPUs (one handle each) configures each of them.
inline bool first_use_on_device(unsigned long long &mask) {
    int dev = 0;
    cudaGetDevice(&dev);
    const unsigned long long bit = 1ull << (dev & 63);
    if (mask & bit) return false;
    mask |= bit;
    return true;
}

// ---- PTX wrappers -----------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {}
}
// 1-D bulk async copy global -> shared (TMA engine; SASS UBLKCP), completion counted on an mbarrier.
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar, uint64_t policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy) : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_last() {
    uint64_t p; asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p)); return p;
}
__device__ __forceinline__ void st_v2_hint(double *p, double a, double b, uint64_t policy) {
    asm volatile("st.global.L2::cache_hint.v2.f64 [%0], {%1,%2}, %3;" ::"l"(p), "d"(a), "d"(b), "l"(policy) : "memory");
}
__device__ __forceinline__ int ld_acquire(const int *p) {
    int v; asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}
// register re-partitioning between warpgroups (sm_90a+): the producer warpgroup hands its registers to the consumers
template <int N> __device__ __forceinline__ void setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// kernel function of one (scaled) squared distance.  `e` is the accumulated exponent  -r^2/2.
// RBF (sklearn RBF.__call__): exp(-0.5 r^2);  Matern nu=2.5 (sklearn Matern.__call__): (1+a+a^2/3)exp(-a), a=sqrt(5) r.
enum KernelKind : int { kRBF = 0, kMatern52 = 1 };

// exp(x) for x <= 0, BRANCH-FREE: Cody-Waite reduction (n = rint(x log2 e) through the 2^52+2^51 magic add, r = x - n ln2 in two
// pieces), degree-13 Taylor polynomial in Horner form, scaling by 2^n as two exact power-of-two factors so that results in the
// subnormal range (x < -708) are rounded once by the hardware and x <= -745.2 gives exactly 0 -- no special-case branch.
// Within 1 ulp of the correctly rounded value (tests/test_gpu_kernels.py::test_kernel_function_accuracy gates 2 ulp against
// numpy over the whole range).  What it buys is the control flow: the library exp carries a slow-path branch, which kept the compiler
// from interleaving independent evaluations -- the lift epilogue of the fused kernel (64 per lane) ran them one after another,
// a dependent chain of 16 FP64 operations each (measured 320 clk per value; profiles/r02_gram_timing.md).
__device__ __forceinline__ double exp_nonpos(double x) {
    x = fmax(x, -750.0);
    const double t = fma(x, 1.4426950408889634074, 6755399441055744.0);
    const int n = __double2loint(t);
    const double nf = t - 6755399441055744.0;
    double r = fma(nf, -6.93147180559945286227e-01, x);
    r = fma(nf, -2.31904681384629955842e-17, r);
    // Taylor polynomial of degree 13 on |r| <= ln2/2 (truncation 4e-18 relative), Horner form
    double p = fma(r, 1.0 / 6227020800.0, 1.0 / 479001600.0);
    p = fma(r, p, 1.0 / 39916800.0);
    p = fma(r, p, 1.0 / 3628800.0);
    p = fma(r, p, 1.0 / 362880.0);
    p = fma(r, p, 1.0 / 40320.0);
    p = fma(r, p, 1.0 / 5040.0);
    p = fma(r, p, 1.0 / 720.0);
    p = fma(r, p, 1.0 / 120.0);
    p = fma(r, p, 1.0 / 24.0);
    p = fma(r, p, 1.0 / 6.0);
    p = fma(r, p, 0.5);
    p = fma(r, p, 1.0);
    p = fma(r, p, 1.0);
    const int n1 = n >> 1, n2 = n - n1;
    const double s1 = __hiloint2double((n1 + 1023) << 20, 0), s2 = __hiloint2double((n2 + 1023) << 20, 0);
    return (p * s1) * s2;
}

template <int KIND> __device__ __forceinline__ double kernel_from_exponent_t(double e) {
    if (KIND == kRBF) return exp_nonpos(fmin(e, 0.0));
    const double r2 = fmax(-2.0 * e, 0.0);
    const double a = sqrt(r2) * 2.23606797749978969641;  // dists * sqrt(5)
    return (1.0 + a + a * a / 3.0) * exp_nonpos(-a);
}
__device__ __forceinline__ double kernel_from_exponent(double e, int kind) {
    return kind == kRBF ? kernel_from_exponent_t<kRBF>(e) : kernel_from_exponent_t<kMatern52>(e);
}

}  // namespace nk
