// nk_dense.cu -- the n-independent dense FP64 stage on DMMA: GEMM, blocked Cholesky, triangular solves,
// symmetric square root, and the Grams -> (A, B, C, W) solve of regressors.py:139-140,147-169.
//
// Everything is phrased as "NT" products  C = A B^T  with both operands contraction-contiguous, so one staging
// path (16-byte cp.async into the 8x8-block shared layout of nk_common.cuh) feeds the same 64x32 DMMA warp tile
// as the fused Gram engine.  Triangular solves keep the unknown TRANSPOSED (X^T rows = right-hand sides), which
// turns every block update of forward/backward substitution into an NT product; 128x128 diagonal blocks are
// factored and inverted by a single-CTA kernel.
//
// Symmetric square root (scipy.linalg.sqrtm at regressors.py:140,163,175): K = L L^T, polar decomposition
// L^T = Q H by Newton-Schulz with the minimax cubic on [l,1] (l tracked analytically from the eigenvalue bound),
// S = H = Q^T L^T, S^-1 = L^-T Q.  Measured against eigh / sqrtm in tools/proto/polar_sqrt.py: same floor.
#include <cmath>
#include <cstdio>
#include <vector>
#include <cooperative_groups.h>
#include "nk_dense.cuh"

namespace nk {

constexpr int kGemmStages = 4;
constexpr size_t kGemmSmem = (size_t)kGemmStages * 2 * kSlabTileDoubles * 8;

struct GemmArgs {
    int M, N, K;
    double alpha, beta, diag;
    const double *A; long long lda;
    const double *B; long long ldb;
    double *C; long long ldc;
    double *Ct; long long ldct;
    int flags, epi_kind, tiles_n;
    long long sA, sB, sC, sCt;   // batch strides (elements); blockIdx.y = batch index
};

__device__ __forceinline__ void cp_async16(void *dst, const void *src, int bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async8(void *dst, const void *src, int bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// stage one 128 x 16 operand slab (rows row0.., depth k0..) into the 8x8-block layout
template <int VEC>
__device__ __forceinline__ void stage_slab(double *dst, const double *src, long long ld, int row0, int nrows, int k0, int K, int tid) {
    if (VEC == 2) {
#pragma unroll
        for (int e = 0; e < 4; e++) {
            const int idx = tid + e * 256;
            const int row = idx >> 3, kk = (idx & 7) * 2;
            const int gr = row0 + row, gk = k0 + kk;
            int bytes = 0;
            const double *s = src;
            if (gr < nrows && gk < K) { bytes = (K - gk >= 2) ? 16 : 8; s = src + (long long)gr * ld + gk; }
            cp_async16(dst + ((row >> 3) * 2 + (kk >> 3)) * kBlk + (row & 7) * 8 + (kk & 7), s, bytes);
        }
    } else {
#pragma unroll
        for (int e = 0; e < 8; e++) {
            const int idx = tid + e * 256;
            const int row = idx >> 4, kk = idx & 15;
            const int gr = row0 + row, gk = k0 + kk;
            int bytes = 0;
            const double *s = src;
            if (gr < nrows && gk < K) { bytes = 8; s = src + (long long)gr * ld + gk; }
            cp_async8(dst + ((row >> 3) * 2 + (kk >> 3)) * kBlk + (row & 7) * 8 + (kk & 7), s, bytes);
        }
    }
}

__device__ __forceinline__ void warp_mma_slab_d(double (&acc)[8][4][2], const double *As, const double *Bs, int wr, int wc, int lane) {
#pragma unroll
    for (int q = 0; q < 2; q++) {
        double2 a[8], b[4];
#pragma unroll
        for (int i = 0; i < 8; i++) a[i] = *reinterpret_cast<const double2 *>(As + ((wr * 8 + i) * 2 + q) * kBlk + lane * 2);
#pragma unroll
        for (int j = 0; j < 4; j++) b[j] = *reinterpret_cast<const double2 *>(Bs + ((wc * 4 + j) * 2 + q) * kBlk + lane * 2);
#pragma unroll
        for (int i = 0; i < 8; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) dmma884(acc[i][j][0], acc[i][j][1], a[i].x, b[j].x);
#pragma unroll
        for (int i = 0; i < 8; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) dmma884(acc[i][j][0], acc[i][j][1], a[i].y, b[j].y);
    }
}

template <int VEC>
__global__ void __launch_bounds__(256, 1) gemm_nt_kernel(GemmArgs g) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    double *sm = reinterpret_cast<double *>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    {
        const long long bz = blockIdx.y;
        g.A += bz * g.sA; g.B += bz * g.sB;
        if (g.C) g.C += bz * g.sC;
        if (g.Ct) g.Ct += bz * g.sCt;
    }
    int I, J;
    if (g.flags & kGemmLowerOnly) {
        // blockIdx.x enumerates the lower triangle row by row
        const int t = blockIdx.x;
        int r = (int)((sqrt(8.0 * t + 1.0) - 1.0) * 0.5);
        while ((r + 1) * (r + 2) / 2 <= t) r++;
        while (r * (r + 1) / 2 > t) r--;
        I = r; J = t - r * (r + 1) / 2;
    } else {
        I = blockIdx.x / g.tiles_n; J = blockIdx.x % g.tiles_n;
    }
    const int nslabs = (g.K + kSlabK - 1) / kSlabK;
    const int wr = warp >> 2, wc = warp & 3;
    double acc[8][4][2];
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }

#pragma unroll
    for (int s = 0; s < kGemmStages - 1; s++) {
        if (s < nslabs) {
            double *st = sm + (size_t)s * 2 * kSlabTileDoubles;
            stage_slab<VEC>(st, g.A, g.lda, I * kTile, g.M, s * kSlabK, g.K, tid);
            stage_slab<VEC>(st + kSlabTileDoubles, g.B, g.ldb, J * kTile, g.N, s * kSlabK, g.K, tid);
        }
        cp_async_commit();
    }
    for (int kt = 0; kt < nslabs; kt++) {
        cp_async_wait<kGemmStages - 2>();
        __syncthreads();
        const int nxt = kt + kGemmStages - 1;
        if (nxt < nslabs) {
            double *st = sm + (size_t)(nxt % kGemmStages) * 2 * kSlabTileDoubles;
            stage_slab<VEC>(st, g.A, g.lda, I * kTile, g.M, nxt * kSlabK, g.K, tid);
            stage_slab<VEC>(st + kSlabTileDoubles, g.B, g.ldb, J * kTile, g.N, nxt * kSlabK, g.K, tid);
        }
        cp_async_commit();
        const double *st = sm + (size_t)(kt % kGemmStages) * 2 * kSlabTileDoubles;
        warp_mma_slab_d(acc, st, st + kSlabTileDoubles, wr, wc, lane);
    }
    cp_async_wait<0>();

    const int gq = lane >> 2, t = lane & 3;
    // plain alpha/beta update of a row-major, 16-byte aligned C: one 16-byte load and store per fragment pair
    if (g.epi_kind < 0 && (g.flags & (kGemmMirror | kGemmStoreT | kGemmPackedOut)) == 0 && g.C != nullptr && (g.ldc % 2 == 0) &&
        (g.sC % 2 == 0) && (((uintptr_t)g.C) % 16 == 0)) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int r = I * kTile + wr * 64 + i * 8 + gq;
            if (r >= g.M) continue;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int c0 = J * kTile + wc * 32 + j * 8 + 2 * t;
                if (c0 >= g.N) continue;
                double *o = g.C + (long long)r * g.ldc + c0;
                double v0 = g.alpha * acc[i][j][0], v1 = g.alpha * acc[i][j][1];
                if (c0 + 1 < g.N) {
                    if (g.beta != 0.0) { const double2 old = *reinterpret_cast<const double2 *>(o); v0 += g.beta * old.x; v1 += g.beta * old.y; }
                    if (r == c0) v0 += g.diag;
                    if (r == c0 + 1) v1 += g.diag;
                    *reinterpret_cast<double2 *>(o) = make_double2(v0, v1);
                } else {
                    if (g.beta != 0.0) v0 += g.beta * o[0];
                    if (r == c0) v0 += g.diag;
                    o[0] = v0;
                }
            }
        }
        return;
    }
    const bool mirror = (g.flags & kGemmMirror) != 0;   // symmetric result: element (r,c), r >= c, is also written to (c,r)
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const int r = I * kTile + wr * 64 + i * 8 + gq;
        if (r >= g.M) continue;
#pragma unroll
        for (int j = 0; j < 4; j++) {
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const int c = J * kTile + wc * 32 + j * 8 + 2 * t + e;
                if (c >= g.N) continue;
                if (mirror && c > r) continue;   // diagonal tile: keep one accumulation per symmetric pair
                double v;
                if (g.epi_kind >= 0) {
                    v = kernel_from_exponent(acc[i][j][e], g.epi_kind);
                    if ((g.flags & kGemmUnitDiag) && r == c) v = 1.0;
                } else {
                    v = g.alpha * acc[i][j][e];
                    if (g.beta != 0.0) v += g.beta * g.C[(long long)r * g.ldc + c];
                    if (r == c) v += g.diag;
                }
                if (g.flags & kGemmPackedOut) { g.C[packed_off(r, c, (int)g.ldc)] = v; continue; }
                if (g.C) g.C[(long long)r * g.ldc + c] = v;
                if (mirror && c != r) g.C[(long long)c * g.ldc + r] = v;
                if (g.flags & kGemmStoreT) g.Ct[(long long)c * g.ldct + r] = v;
            }
        }
    }
}

void gemm_nt(nk_handle *h, int M, int N, int K, double alpha, const double *A, long long lda, const double *B, long long ldb,
             double beta, double *C, long long ldc, double diag, int flags, double *Ct, long long ldct, cudaStream_t stream,
             int epi_kind) {
    gemm_nt_batched(h, 1, M, N, K, alpha, A, lda, 0, B, ldb, 0, beta, C, ldc, 0, diag, flags, Ct, ldct, 0, stream, epi_kind);
}

void gemm_nt_batched(nk_handle *h, int batch, int M, int N, int K, double alpha, const double *A, long long lda, long long sA,
                     const double *B, long long ldb, long long sB, double beta, double *C, long long ldc, long long sC, double diag,
                     int flags, double *Ct, long long ldct, long long sCt, cudaStream_t stream, int epi_kind) {
    if (M <= 0 || N <= 0 || batch <= 0) return;
    if (tgemm_try(h, batch, M, N, K, alpha, A, lda, sA, B, ldb, sB, beta, C, ldc, sC, diag, flags, Ct, ldct, sCt, stream, epi_kind)) return;
    static unsigned long long configured = 0;
    if (first_use_on_device(configured)) {
        cudaFuncSetAttribute(gemm_nt_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGemmSmem);
        cudaFuncSetAttribute(gemm_nt_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGemmSmem);
    }
    GemmArgs g;
    g.M = M; g.N = N; g.K = K; g.alpha = alpha; g.beta = beta; g.diag = diag;
    g.A = A; g.lda = lda; g.B = B; g.ldb = ldb; g.C = C; g.ldc = ldc; g.Ct = Ct; g.ldct = ldct;
    g.flags = flags; g.epi_kind = epi_kind;
    g.sA = sA; g.sB = sB; g.sC = sC; g.sCt = sCt;
    const int tm = (M + kTile - 1) / kTile, tn = (N + kTile - 1) / kTile;
    g.tiles_n = tn;
    const int grid = (flags & kGemmLowerOnly) ? tm * (tm + 1) / 2 : tm * tn;
    const bool vec2 = ((lda | ldb | sA | sB) % 2 == 0) && (((uintptr_t)A | (uintptr_t)B) % 16 == 0);
    const dim3 grid2(grid, batch);
    if (vec2) gemm_nt_kernel<2><<<grid2, 256, kGemmSmem, stream>>>(g);
    else gemm_nt_kernel<1><<<grid2, 256, kGemmSmem, stream>>>(g);
    h->launches++;
}

// ---------------------------------------------------------------------------------------------------
// small kernels
// ---------------------------------------------------------------------------------------------------
__global__ void transpose_kernel(int rows, int cols, const double *src, long long lds, double *dst, long long ldd) {
    __shared__ double tile[32][33];
    const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += 8) {
        const int r = by + j, c = bx + threadIdx.x;
        if (r < rows && c < cols) tile[j][threadIdx.x] = src[(long long)r * lds + c];
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += 8) {
        const int c = bx + j, r = by + threadIdx.x;
        if (r < rows && c < cols) dst[(long long)c * ldd + r] = tile[threadIdx.x][j];
    }
}
void transpose(nk_handle *h, int rows, int cols, const double *src, long long lds, double *dst, long long ldd, cudaStream_t stream) {
    if (rows <= 0 || cols <= 0) return;
    dim3 grid((cols + 31) / 32, (rows + 31) / 32), block(32, 8);
    transpose_kernel<<<grid, block, 0, stream>>>(rows, cols, src, lds, dst, ldd);
    h->launches++;
}

// out = 0.5 (in + in^T)
__global__ void symmetrize_kernel(int n, const double *in, long long ldi, double *out, long long ldo) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
    if (c < n) out[(long long)r * ldo + c] = 0.5 * (in[(long long)r * ldi + c] + in[(long long)c * ldi + r]);
}
// dst = scale * src (rows x cols)
__global__ void scale_copy_kernel(int rows, int cols, double scale, const double *src, long long lds, double *dst, long long ldd) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
    if (c < cols) dst[(long long)r * ldd + c] = scale * src[(long long)r * lds + c];
}
__global__ void zero_upper_kernel(int n, double *A, long long lda, long long sA) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
    A += (long long)blockIdx.z * sA;
    if (c < n && c > r) A[(long long)r * lda + c] = 0.0;
}
__global__ void rowsum_max_kernel(int n, const double *A, long long lda, double *out) {
    // out[0] = max_r sum_c |A(r,c)|   (out must be zeroed; doubles are non-negative so the int64 compare order is monotone)
    const int r = blockIdx.x;
    double s = 0.0;
    for (int c = threadIdx.x; c < n; c += blockDim.x) s += fabs(A[(long long)r * lda + c]);
    __shared__ double red[256];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) { if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o]; __syncthreads(); }
    if (threadIdx.x == 0) atomicMax(reinterpret_cast<unsigned long long *>(out), (unsigned long long)__double_as_longlong(red[0]));
}
static void launch2d(int rows, int cols, dim3 &grid, dim3 &block) { block = dim3(128); grid = dim3((cols + 127) / 128, rows); }

// ---- helpers of the symmetric square root's spectrum estimate and convergence check ----
// rows of +-1 from a fixed hash: the start vectors of the inverse iteration (deterministic)
__global__ void probe_init_kernel(int rows, int n, double *V, long long ld) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
    if (c >= n) return;
    unsigned x = (unsigned)(c * 8 + r) * 2654435761u;
    x ^= x >> 15; x *= 2246822519u; x ^= x >> 13;
    V[(long long)r * ld + c] = (x & 1u) ? 1.0 : -1.0;
}
// out[r] = || V[r, :] ||_2, one CTA per row, fixed-order tree (deterministic)
__global__ void row_norm_kernel(int n, const double *V, long long ld, double *out) {
    const int r = blockIdx.x;
    double s = 0.0;
    for (int c = threadIdx.x; c < n; c += blockDim.x) { const double v = V[(long long)r * ld + c]; s += v * v; }
    __shared__ double red[256];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) { if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o]; __syncthreads(); }
    if (threadIdx.x == 0) out[r] = sqrt(red[0]);
}
// out[0] = max_{r,c} | T(r,c) - [r == c] |   (out zeroed by the caller; non-negative doubles order like their bit patterns)
__global__ void max_dev_identity_kernel(int n, const double *T, long long ld, double *out) {
    const int r = blockIdx.x;
    double s = 0.0;
    for (int c = threadIdx.x; c < n; c += blockDim.x) s = fmax(s, fabs(T[(long long)r * ld + c] - (r == c ? 1.0 : 0.0)));
    __shared__ double red[256];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) { if (threadIdx.x < o) red[threadIdx.x] = fmax(red[threadIdx.x], red[threadIdx.x + o]); __syncthreads(); }
    if (threadIdx.x == 0) atomicMax(reinterpret_cast<unsigned long long *>(out), (unsigned long long)__double_as_longlong(red[0]));
}

// centre = column means of the landmarks (any shift is valid: distances are shift invariant; the mean minimises the norms
// entering the expansion).  32 columns x 32 row lanes per CTA, fixed-order tree over the row lanes -> deterministic.
__global__ void landmark_center_kernel2(const double *Z, long long ldz, int m, int d, double *center) {
    __shared__ double part[32][33];
    const int k = blockIdx.x * 32 + threadIdx.x;
    double s = 0.0;
    if (k < d) for (int r = threadIdx.y; r < m; r += 32) s += Z[(long long)r * ldz + k];
    part[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0 && k < d) {
        double t = 0.0;
        for (int j = 0; j < 32; j++) t += part[j][threadIdx.x];
        center[k] = t / m;
    }
}
void landmark_center(const double *Z, long long ldz, int m, int d, double *center, cudaStream_t stream) {
    landmark_center_kernel2<<<(d + 31) / 32, dim3(32, 32), 0, stream>>>(Z, ldz, m, d, center);
}

__global__ void augment_rows_kernel(const double *src, long long lds, long long rows, int d, const double *inv_ls, const double *center,
                                    int landmark_form, double *out, int KA) {
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    double nrm = 0.0;
    for (int k = lane; k < KA; k += 32) {
        double v = 0.0;
        if (k < d) { v = (src[row * lds + k] - center[k]) * inv_ls[k]; nrm += v * v; }
        if (k < d || k >= d + 2) out[row * KA + k] = v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nrm += __shfl_xor_sync(0xffffffffu, nrm, o);
    if (lane == 0) {
        out[row * KA + d] = landmark_form ? 1.0 : -0.5 * nrm;
        out[row * KA + d + 1] = landmark_form ? -0.5 * nrm : 1.0;
    }
}
void augment_rows(nk_handle *h, const double *src, long long lds, long long rows, int d, const double *inv_ls, const double *center,
                  int landmark_form, double *out, int KA, cudaStream_t stream) {
    if (rows <= 0) return;
    const int warps = 8;
    augment_rows_kernel<<<(unsigned)((rows + warps - 1) / warps), warps * 32, 0, stream>>>(src, lds, rows, d, inv_ls, center, landmark_form, out, KA);
    h->launches++;
}

// ---------------------------------------------------------------------------------------------------
// 128 x 128 diagonal block: Cholesky factor and its inverse, one CTA of 256 threads per matrix of the batch.
//
// Register-tiled and cyclic: thread (ty, tx) = (tid / 16, tid % 16) owns the 64 elements (ty + 16 i, tx + 16 j).
// Pass 1 (right-looking Cholesky): per pivot k the owners of column k publish it through a double-buffered
// shared vector (ONE barrier per pivot); every thread scales it with rsqrt(a_kk) and applies the rank-1 update to
// the lower-triangular part of its registers.  Pass 2 (Gauss-Jordan on the identity): L^-1 by forward elimination
// with the columns of L kept in shared memory.  Loops are unrolled over the 16-wide register block index so that
// every register access is static.  Blocks smaller than 128 are padded with the identity.
// ---------------------------------------------------------------------------------------------------
constexpr int kDB = 128;
constexpr size_t kDiagSmem = (size_t)kDB * kDB * 8 + 2 * kDB * 8 + 2 * kDB * 8 + kDB * 8;

__global__ void __launch_bounds__(256, 1) potrf_diag_kernel(double *A_, long long lda, long long strideA, int nb, int j0, double *dinv_,
                                                            double *dinvT_, long long strideD, int *info_, int do_factor) {
    extern __shared__ __align__(16) double dsm[];
    double *Lc = dsm;                       // Lc[k * 128 + r] = L(r, k), r > k
    double *colb = Lc + kDB * kDB;          // 2 x 128: published pivot column / row
    double *invd = colb + 2 * kDB;          // 1 / L(k, k)
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    double *A = A_ + (long long)blockIdx.x * strideA;
    double *dinv = dinv_ + (long long)blockIdx.x * strideD;
    double *dinvT = dinvT_ ? dinvT_ + (long long)blockIdx.x * strideD : nullptr;
    int *info = info_ + blockIdx.x;

    double a[8][8];
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const int r = ty + 16 * i, c = tx + 16 * j;
            double v = (r == c) ? 1.0 : 0.0;
            if (r < nb && c < nb) v = (c <= r) ? A[(long long)r * lda + c] : 0.0;
            a[i][j] = v;
        }
    if (do_factor) {
#pragma unroll
        for (int kq = 0; kq < 8; kq++) {
            for (int kr = 0; kr < 16; kr++) {
                const int k = kq * 16 + kr;
                if (k >= nb) break;
                double *buf = colb + (k & 1) * kDB;
                if (tx == kr) {
#pragma unroll
                    for (int i = kq; i < 8; i++) buf[ty + 16 * i] = a[i][kq];
                }
                __syncthreads();
                double akk = buf[k];
                if (!(akk > 0.0)) { if (tid == 0) atomicCAS(info, 0, j0 + k + 1); akk = 1.0; }
                const double inv = rsqrt(akk);
                double lr[8], lc[8];
#pragma unroll
                for (int i = kq; i < 8; i++) lr[i] = (ty + 16 * i > k) ? buf[ty + 16 * i] * inv : 0.0;
#pragma unroll
                for (int j = kq; j < 8; j++) lc[j] = (tx + 16 * j > k) ? buf[tx + 16 * j] * inv : 0.0;
#pragma unroll
                for (int i = kq; i < 8; i++)
#pragma unroll
                    for (int j = kq; j <= i; j++) a[i][j] -= lr[i] * lc[j];
                if (tx == kr) {
#pragma unroll
                    for (int i = kq; i < 8; i++) {
                        const int r = ty + 16 * i;
                        if (r > k) { a[i][kq] = lr[i]; Lc[k * kDB + r] = lr[i]; }
                        else if (r == k) { a[i][kq] = akk * inv; invd[k] = inv; }
                    }
                }
            }
        }
        // identity padding of a short block
        for (int k = nb + tid; k < kDB; k += 256) invd[k] = 1.0;
        for (int idx = tid; idx < (kDB - nb) * kDB; idx += 256) Lc[(nb + idx / kDB) * kDB + idx % kDB] = 0.0;
#pragma unroll
        for (int i = 0; i < 8; i++)
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const int r = ty + 16 * i, c = tx + 16 * j;
                if (r < nb && c < nb) A[(long long)r * lda + c] = (c <= r) ? a[i][j] : 0.0;
            }
    } else {
        // the block already holds a factor: publish its columns for pass 2
#pragma unroll
        for (int i = 0; i < 8; i++)
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const int r = ty + 16 * i, c = tx + 16 * j;
                if (c < r) Lc[c * kDB + r] = a[i][j];
                else if (c == r) invd[r] = 1.0 / a[i][j];
            }
    }
    __syncthreads();

    // ---- pass 2: X = L^-1 by forward elimination of the identity ----
    double x[8][8];
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 8; j++) x[i][j] = (ty + 16 * i == tx + 16 * j) ? 1.0 : 0.0;
#pragma unroll
    for (int kq = 0; kq < 8; kq++) {
        for (int kr = 0; kr < 16; kr++) {
            const int k = kq * 16 + kr;
            double *buf = colb + (k & 1) * kDB;
            if (ty == kr) {
#pragma unroll
                for (int j = 0; j <= kq; j++) buf[tx + 16 * j] = x[kq][j];
            }
            __syncthreads();
            const double dk = invd[k];
            double xr[8], lk[8];
#pragma unroll
            for (int j = 0; j <= kq; j++) xr[j] = (tx + 16 * j <= k) ? buf[tx + 16 * j] * dk : 0.0;
#pragma unroll
            for (int i = kq; i < 8; i++) lk[i] = (ty + 16 * i > k) ? Lc[k * kDB + ty + 16 * i] : 0.0;
#pragma unroll
            for (int i = kq; i < 8; i++)
#pragma unroll
                for (int j = 0; j <= kq; j++) x[i][j] -= lk[i] * xr[j];
            if (ty == kr) {
#pragma unroll
                for (int j = 0; j <= kq; j++) x[kq][j] = xr[j];
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const int r = ty + 16 * i, c = tx + 16 * j;
            const double v = (c <= r) ? x[i][j] : 0.0;
            dinv[r * kDB + c] = v;
        }
    if (dinvT) {
        // transposed copy with coalesced stores: thread owns (r, c) -> write element (c, r) of L^-1 ... via shared staging
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 8; i++)
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const int r = ty + 16 * i, c = tx + 16 * j;
                Lc[c * kDB + r] = (c <= r) ? x[i][j] : 0.0;      // Lc[c][r] = X(r, c) = X^T(c, r)
            }
        __syncthreads();
        for (int idx = tid; idx < kDB * kDB; idx += 256) dinvT[idx] = Lc[idx];
    }
}

static void launch_diag(nk_handle *h, int batch, double *A, long long lda, long long sA, int nb, int j0, double *dinv, double *dinvT,
                        long long sD, int *dinfo, int do_factor, cudaStream_t stream) {
    static unsigned long long configured = 0;
    if (first_use_on_device(configured)) cudaFuncSetAttribute(potrf_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kDiagSmem);
    potrf_diag_kernel<<<batch, 256, kDiagSmem, stream>>>(A, lda, sA, nb, j0, dinv, dinvT, sD, dinfo, do_factor);
    h->launches++;
}

__global__ void transpose_batched_kernel(int rows, int cols, const double *src, long long lds, long long ss, double *dst, long long ldd, long long sd) {
    __shared__ double tile[32][33];
    src += (long long)blockIdx.z * ss; dst += (long long)blockIdx.z * sd;
    const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += 8) {
        const int r = by + j, c = bx + threadIdx.x;
        if (r < rows && c < cols) tile[j][threadIdx.x] = src[(long long)r * lds + c];
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += 8) {
        const int c = bx + j, r = by + threadIdx.x;
        if (r < rows && c < cols) dst[(long long)c * ldd + r] = tile[threadIdx.x][j];
    }
}
void transpose_batched(nk_handle *h, int batch, int rows, int cols, const double *src, long long lds, long long ss, double *dst,
                       long long ldd, long long sd, cudaStream_t stream) {
    if (rows <= 0 || cols <= 0 || batch <= 0) return;
    dim3 grid((cols + 31) / 32, (rows + 31) / 32, batch), block(32, 8);
    transpose_batched_kernel<<<grid, block, 0, stream>>>(rows, cols, src, lds, ss, dst, ldd, sd);
    h->launches++;
}

// Batched blocked Cholesky (right-looking, 128-wide panels): every matrix of the batch advances through the same
// launches (grid.y / grid.x = batch), so the serial diagonal-block kernel of one matrix overlaps with the others'.
// dinfo: `batch` device ints (0 or the 1-based index of the first non-positive pivot).
int potrf_batched(nk_handle *h, int batch, int n, double *A, long long lda, long long sA, double *Lt, long long ldlt, long long sLt,
                  double *dinv, double *dinvT, long long sD, int *dinfo, cudaStream_t stream, bool clean_upper) {
    cudaMemsetAsync(dinfo, 0, sizeof(int) * batch, stream);
    // Two-level blocking.  Inside a 512-wide panel the 128-wide steps only update the panel's own remaining columns
    // (short-K products, little work); everything to the right of the panel is updated ONCE per panel with K = 512,
    // so the bulk of the n^3/3 flops runs in long-K GEMMs and the trailing matrix is read-modify-written n/512 times
    // instead of n/128 times.
    constexpr int kOuter = 4 * kDB;
    for (int p0 = 0; p0 < n; p0 += kOuter) {
        const int pend = (p0 + kOuter < n) ? p0 + kOuter : n;
        for (int j0 = p0; j0 < pend; j0 += kDB) {
            const int kb = j0 / kDB, nb = (pend - j0 < kDB) ? pend - j0 : kDB, j1 = j0 + nb, rem = n - j1, pc = pend - j1;
            double *Akk = A + (long long)j0 * lda + j0;
            double *di = dinv + (size_t)kb * kDB * kDB;
            launch_diag(h, batch, Akk, lda, sA, nb, j0, di, dinvT ? dinvT + (size_t)kb * kDB * kDB : nullptr, sD, dinfo, 1, stream);
            if (rem > 0) {
                double *A21 = A + (long long)j1 * lda + j0;          // rows below the block, its nb columns
                gemm_nt_batched(h, batch, rem, nb, nb, 1.0, A21, lda, sA, di, kDB, sD, 0.0, A21, lda, sA, 0.0, 0, nullptr, 0, 0, stream);
                if (pc > 0)                                           // remaining columns of this panel (rectangular, K = nb)
                    gemm_nt_batched(h, batch, rem, pc, nb, -1.0, A21, lda, sA, A21, lda, sA, 1.0, A + (long long)j1 * lda + j1, lda, sA, 0.0, 0,
                                    nullptr, 0, 0, stream);
            }
        }
        const int rem2 = n - pend;
        if (rem2 > 0) {                                               // everything right of the panel, K = panel width
            double *P = A + (long long)pend * lda + p0;
            gemm_nt_batched(h, batch, rem2, rem2, pend - p0, -1.0, P, lda, sA, P, lda, sA, 1.0, A + (long long)pend * lda + pend, lda, sA, 0.0,
                            kGemmLowerOnly, nullptr, 0, 0, stream);
        }
    }
    // the strict upper triangle holds scratch values; the solves of this file never read it (they use the sub-diagonal
    // blocks of L / the super-diagonal blocks of L^T and the inverted diagonal blocks), so it is only cleaned on request
    if (clean_upper) {
        dim3 block(128), grid((n + 127) / 128, n, batch);
        zero_upper_kernel<<<grid, block, 0, stream>>>(n, A, lda, sA);
        h->launches++;
    }
    if (Lt) transpose_batched(h, batch, n, n, A, lda, sA, Lt, ldlt, sLt, stream);
    return NK_OK;
}

int potrf_blocked(nk_handle *h, int n, double *A, long long lda, double *Lt, long long ldlt, double *dinv, double *dinvT,
                  int *dinfo, cudaStream_t stream) {
    return potrf_batched(h, 1, n, A, lda, 0, Lt, ldlt, 0, dinv, dinvT, 0, dinfo, stream, true);
}

// ---- triangular solves in transposed storage (Xt rows = right-hand sides) ----
// Right-looking, batched: once block i of the unknown is known, all later (forward) / earlier (backward) blocks are
// updated by one wide product -> (r / 128) x (n / 128) CTAs per launch.  (The first version was left-looking -- block i as
// one product over all earlier blocks -- which launches only r / 128 CTAs at a time: 32 of 148 SMs busy at m = 4096, and
// 55 of the 88 ms of a fit's solve stage.)  sL / sD / sX = 0 shares the factor / right-hand sides across the batch.
void trsm_fwd_t_rl(nk_handle *h, int batch, int n, int r, const double *L, long long ldl, long long sL, const double *dinv, long long sD,
                   double *Xt, long long ldx, long long sX, cudaStream_t stream) {
    constexpr int kOuter = 4 * kDB;      // same two-level blocking as potrf_batched: the bulk of the update runs with K = 512
    for (int p0 = 0; p0 < n; p0 += kOuter) {
        const int pend = (p0 + kOuter < n) ? p0 + kOuter : n;
        for (int j0 = p0; j0 < pend; j0 += kDB) {
            const int nb = (pend - j0 < kDB) ? pend - j0 : kDB, j1 = j0 + nb, pc = pend - j1;
            gemm_nt_batched(h, batch, r, nb, nb, 1.0, Xt + j0, ldx, sX, dinv + (size_t)(j0 / kDB) * kDB * kDB, kDB, sD, 0.0, Xt + j0, ldx, sX, 0.0, 0,
                            nullptr, 0, 0, stream);
            if (pc > 0) gemm_nt_batched(h, batch, r, pc, nb, -1.0, Xt + j0, ldx, sX, L + (long long)j1 * ldl + j0, ldl, sL, 1.0, Xt + j1, ldx, sX,
                                        0.0, 0, nullptr, 0, 0, stream);
        }
        if (pend < n) gemm_nt_batched(h, batch, r, n - pend, pend - p0, -1.0, Xt + p0, ldx, sX, L + (long long)pend * ldl + p0, ldl, sL, 1.0,
                                      Xt + pend, ldx, sX, 0.0, 0, nullptr, 0, 0, stream);
    }
}
void trsm_bwd_t_rl(nk_handle *h, int batch, int n, int r, const double *Lt, long long ldlt, long long sLt, const double *dinvT, long long sD,
                   double *Xt, long long ldx, long long sX, cudaStream_t stream) {
    constexpr int kOuter = 4 * kDB;
    const int npan = (n + kOuter - 1) / kOuter;
    for (int pi = npan - 1; pi >= 0; pi--) {
        const int p0 = pi * kOuter, pend = (p0 + kOuter < n) ? p0 + kOuter : n;
        const int nsub = (pend - p0 + kDB - 1) / kDB;
        for (int si = nsub - 1; si >= 0; si--) {
            const int j0 = p0 + si * kDB, nb = (pend - j0 < kDB) ? pend - j0 : kDB, pc = j0 - p0;
            gemm_nt_batched(h, batch, r, nb, nb, 1.0, Xt + j0, ldx, sX, dinvT + (size_t)(j0 / kDB) * kDB * kDB, kDB, sD, 0.0, Xt + j0, ldx, sX, 0.0, 0,
                            nullptr, 0, 0, stream);
            if (pc > 0) gemm_nt_batched(h, batch, r, pc, nb, -1.0, Xt + j0, ldx, sX, Lt + (long long)p0 * ldlt + j0, ldlt, sLt, 1.0, Xt + p0, ldx, sX,
                                        0.0, 0, nullptr, 0, 0, stream);
        }
        if (p0 > 0) gemm_nt_batched(h, batch, r, p0, pend - p0, -1.0, Xt + p0, ldx, sX, Lt + p0, ldlt, sLt, 1.0, Xt, ldx, sX, 0.0, 0, nullptr, 0, 0,
                                    stream);
    }
}

void trsm_fwd_t(nk_handle *h, int n, int r, const double *L, long long ldl, const double *dinv, double *Xt, long long ldx, cudaStream_t stream) {
    trsm_fwd_t_rl(h, 1, n, r, L, ldl, 0, dinv, 0, Xt, ldx, 0, stream);
}
void trsm_bwd_t(nk_handle *h, int n, int r, const double *Lt, long long ldlt, const double *dinvT, double *Xt, long long ldx, cudaStream_t stream) {
    trsm_bwd_t_rl(h, 1, n, r, Lt, ldlt, 0, dinvT, 0, Xt, ldx, 0, stream);
}

double *dense_scratch(nk_handle *h, int slot, size_t doubles, int *rc) {
    *rc = ensure(h, h->dense[slot], doubles * 8);
    return (double *)h->dense[slot].ptr;
}

// minimax odd cubic p(x) = a x - b x^3 on [l, 1], rescaled so that max p = 1; returns the new lower bound
static double opt_cubic(double l, double *a, double *b) {
    const double s = 1.0 + l + l * l, xs = std::sqrt(s / 3.0);
    double bb = 2.0 / ((2.0 * s / 3.0) * xs + (l + l * l));
    double aa = bb * s;
    const double pmax = (2.0 * aa / 3.0) * xs, pmin = aa - bb;
    *a = aa / pmax; *b = bb / pmax;
    return pmin / pmax;
}

static inline int even(int x) { return (x + 1) & ~1; }

// ---------------------------------------------------------------------------------------------------
// Symmetric square root of a SMALL matrix (n <= 128: the script configurations, m = 10 ... 100) in ONE cooperative launch.
// The general path above is ~100 dependent small launches with a host synchronisation in the middle (1.4 ms at m = 100: half of a
// whole script-sized fit).  Here 16 CTAs each own a 32 x 32 tile of every n x n product (operand rows staged in shared memory,
// plain FP64 FMAs: the matrices are 128 KB and live in L2), separated by grid-wide barriers: the same polar Newton-Schulz
// iteration with the same minimax-cubic schedule, computed redundantly by every thread from the same numbers.
// Inputs: Kc (n x n copy of K), L / Linv / LinvT (factor of K and the inverse of the factor, from potrf_diag_kernel).
// ---------------------------------------------------------------------------------------------------
constexpr int kSN = 128;                                   // leading dimension of every scratch matrix
constexpr int kSmallPad = 129;
constexpr size_t kSmallSmem = 2 * 32 * kSmallPad * 8;      // two 32 x 128 operand strips

__device__ __forceinline__ double dev_opt_cubic(double l, double &a, double &b) {
    const double s = 1.0 + l + l * l, xs = sqrt(s / 3.0);
    const double bb = 2.0 / ((2.0 * s / 3.0) * xs + (l + l * l));
    const double aa = bb * s;
    const double pmax = (2.0 * aa / 3.0) * xs, pmin = aa - bb;
    a = aa / pmax; b = bb / pmax;
    return pmin / pmax;
}

// acc(2x2 per thread) = sum_k A[32 bi + r, k] B[32 bj + c, k]  for the CTA's tile; r in {ty, ty+16}, c in {tx, tx+16}
__device__ __forceinline__ void small_nt_tile(const double *__restrict__ A, const double *__restrict__ B, int n, int bi, int bj, double (&acc)[2][2],
                                              double *sA, double *sB) {
    const int tid = threadIdx.x;
    for (int idx = tid; idx < 32 * kSN; idx += 256) {
        const int r = idx >> 7, k = idx & (kSN - 1);
        const int ga = bi * 32 + r, gb = bj * 32 + r;
        sA[r * kSmallPad + k] = (ga < n && k < n) ? A[ga * kSN + k] : 0.0;
        sB[r * kSmallPad + k] = (gb < n && k < n) ? B[gb * kSN + k] : 0.0;
    }
    __syncthreads();
    const int ty = tid >> 4, tx = tid & 15;
    acc[0][0] = acc[0][1] = acc[1][0] = acc[1][1] = 0.0;
    for (int k = 0; k < n; k++) {
        const double a0 = sA[ty * kSmallPad + k], a1 = sA[(ty + 16) * kSmallPad + k];
        const double b0 = sB[tx * kSmallPad + k], b1 = sB[(tx + 16) * kSmallPad + k];
        acc[0][0] = fma(a0, b0, acc[0][0]); acc[0][1] = fma(a0, b1, acc[0][1]);
        acc[1][0] = fma(a1, b0, acc[1][0]); acc[1][1] = fma(a1, b1, acc[1][1]);
    }
    __syncthreads();
}

__global__ void __launch_bounds__(256, 1) small_polar_kernel(int n, double lambda_min_bound, const double *Kc, const double *L, const double *LinvT,
                                                             double *X, double *Xt, double *X2, double *Xt2, double *T, double *S, long long lds,
                                                             double *Sinv, long long ldsi, int *iters_out) {
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    extern __shared__ __align__(16) double ssm[];
    double *sA = ssm, *sB = ssm + 32 * kSmallPad;
    __shared__ double red[256];
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const int bi = blockIdx.x >> 2, bj = blockIdx.x & 3;
    const int r0 = bi * 32 + ty, r1 = r0 + 16, c0 = bj * 32 + tx, c1 = c0 + 16;
    // ||K||_inf (every CTA, redundantly): an upper bound on ||L^T||_2^2
    double rs = 0.0;
    if (tid < n) for (int c = 0; c < n; c++) rs += fabs(Kc[c * kSN + tid]);      // K is symmetric: column sums, coalesced
    red[tid] = rs;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) { if (tid < o) red[tid] = fmax(red[tid], red[tid + o]); __syncthreads(); }
    const double nrm = sqrt(red[0]);
    __syncthreads();
    // X0 = L^T / nrm, Xt0 = L / nrm  (this CTA's tile)
    {
        const int rr[2] = {r0, r1}, cc[2] = {c0, c1};
#pragma unroll
        for (int i = 0; i < 2; i++)
#pragma unroll
            for (int j = 0; j < 2; j++)
                if (rr[i] < n && cc[j] < n) {
                    const double lv = L[rr[i] * kSN + cc[j]] / nrm;      // L is lower triangular with an explicit zero upper part
                    Xt[rr[i] * kSN + cc[j]] = lv;
                    X[cc[j] * kSN + rr[i]] = lv;
                }
    }
    grid.sync();
    double l = fmin(1.0, 0.9 * sqrt(lambda_min_bound) / nrm);
    int it = 0, plain = 0;
    double acc[2][2];
    while (it < 100) {
        double a, b;
        if (l < 0.999) l = dev_opt_cubic(l, a, b);
        else { a = 1.5; b = 0.5; plain++; }
        // T = a I - b X X^T
        small_nt_tile(X, X, n, bi, bj, acc, sA, sB);
        {
            const int rr[2] = {r0, r1}, cc[2] = {c0, c1};
#pragma unroll
            for (int i = 0; i < 2; i++)
#pragma unroll
                for (int j = 0; j < 2; j++)
                    if (rr[i] < n && cc[j] < n) T[rr[i] * kSN + cc[j]] = -b * acc[i][j] + (rr[i] == cc[j] ? a : 0.0);
        }
        grid.sync();
        // X <- T X  (NT with the transposed copy), both orientations stored
        small_nt_tile(T, Xt, n, bi, bj, acc, sA, sB);
        {
            const int rr[2] = {r0, r1}, cc[2] = {c0, c1};
#pragma unroll
            for (int i = 0; i < 2; i++)
#pragma unroll
                for (int j = 0; j < 2; j++)
                    if (rr[i] < n && cc[j] < n) { X2[rr[i] * kSN + cc[j]] = acc[i][j]; Xt2[cc[j] * kSN + rr[i]] = acc[i][j]; }
        }
        grid.sync();
        double *t1 = X; X = X2; X2 = t1;
        double *t2 = Xt; Xt = Xt2; Xt2 = t2;
        it++;
        if (plain >= 3) break;
    }
    // M1 = Q^T L^T  (-> T),  M2 = L^-T Q  (-> X2);  Q = X
    small_nt_tile(Xt, L, n, bi, bj, acc, sA, sB);
    {
        const int rr[2] = {r0, r1}, cc[2] = {c0, c1};
#pragma unroll
        for (int i = 0; i < 2; i++)
#pragma unroll
            for (int j = 0; j < 2; j++)
                if (rr[i] < n && cc[j] < n) T[rr[i] * kSN + cc[j]] = acc[i][j];
    }
    small_nt_tile(LinvT, Xt, n, bi, bj, acc, sA, sB);
    {
        const int rr[2] = {r0, r1}, cc[2] = {c0, c1};
#pragma unroll
        for (int i = 0; i < 2; i++)
#pragma unroll
            for (int j = 0; j < 2; j++)
                if (rr[i] < n && cc[j] < n) X2[rr[i] * kSN + cc[j]] = acc[i][j];
    }
    grid.sync();
    {
        const int rr[2] = {r0, r1}, cc[2] = {c0, c1};
#pragma unroll
        for (int i = 0; i < 2; i++)
#pragma unroll
            for (int j = 0; j < 2; j++)
                if (rr[i] < n && cc[j] < n) {
                    S[(long long)rr[i] * lds + cc[j]] = 0.5 * (T[rr[i] * kSN + cc[j]] + T[cc[j] * kSN + rr[i]]);
                    Sinv[(long long)rr[i] * ldsi + cc[j]] = 0.5 * (X2[rr[i] * kSN + cc[j]] + X2[cc[j] * kSN + rr[i]]);
                }
    }
    if (blockIdx.x == 0 && tid == 0) *iters_out = it;
}

}  // namespace nk

using namespace nk;

// =====================================================================================================
// C ABI
// =====================================================================================================
extern "C" {

int nk_gemm(nk_handle *h, int transa, int transb, int M, int N, int K, double alpha, const double *A, long long lda,
            const double *B, long long ldb, double beta, double *C, long long ldc, void *stream_) {
    if (!h) return NK_E_INVALID;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (M < 0 || N < 0 || K < 0 || !A || !B || !C) return set_err(h, NK_E_INVALID, "nk_gemm: bad argument");
    NK_ON_DEVICE(h);
    int rc = NK_OK;
    const double *Ak = A; long long ldak = lda;      // (M,K) k-contiguous
    const double *Bk = B; long long ldbk = ldb;      // (N,K) k-contiguous
    if (transa) {  // stored (K,M): transpose into scratch
        double *t = dense_scratch(h, 10, (size_t)M * even(K), &rc); if (rc) return rc;
        transpose(h, K, M, A, lda, t, even(K), stream); Ak = t; ldak = even(K);
    }
    if (!transb) { // stored (K,N): transpose into scratch
        double *t = dense_scratch(h, 11, (size_t)N * even(K), &rc); if (rc) return rc;
        transpose(h, K, N, B, ldb, t, even(K), stream); Bk = t; ldbk = even(K);
    }
    gemm_nt(h, M, N, K, alpha, Ak, ldak, Bk, ldbk, beta, C, ldc, 0.0, 0, nullptr, 0, stream);
    NK_CUDA(h, cudaGetLastError());
    return NK_OK;
}

int nk_potrf(nk_handle *h, int n, double *A, long long lda, int *info, void *stream_) {
    if (!h) return NK_E_INVALID;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n < 1 || !A || lda < n) return set_err(h, NK_E_INVALID, "nk_potrf: bad argument");
    NK_ON_DEVICE(h);
    int rc;
    const int nblk = (n + kDB - 1) / kDB;
    double *dinv = dense_scratch(h, 8, (size_t)nblk * kDB * kDB, &rc); if (rc) return rc;
    double *dinvT = dense_scratch(h, 9, (size_t)nblk * kDB * kDB, &rc); if (rc) return rc;
    if ((rc = ensure(h, h->dinfo, 64)) != NK_OK) return rc;
    potrf_blocked(h, n, A, lda, nullptr, 0, dinv, dinvT, (int *)h->dinfo.ptr, stream);
    int hinfo = 0;
    NK_CUDA(h, cudaMemcpyAsync(&hinfo, h->dinfo.ptr, sizeof(int), cudaMemcpyDeviceToHost, stream));
    NK_CUDA(h, cudaStreamSynchronize(stream));
    if (info) *info = hinfo;
    if (hinfo != 0) return set_err(h, NK_E_NOT_SPD, "nk_potrf: matrix is not positive definite (pivot " + std::to_string(hinfo) + ")");
    return NK_OK;
}

int nk_trsm_lower(nk_handle *h, int trans, int n, int nrhs, const double *L, long long ldl, double *B, long long ldb, void *stream_) {
    if (!h) return NK_E_INVALID;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n < 1 || nrhs < 1 || !L || !B) return set_err(h, NK_E_INVALID, "nk_trsm_lower: bad argument");
    NK_ON_DEVICE(h);
    // generic entry: rebuilds the diagonal-block inverses from L (the fused paths reuse the ones potrf produced)
    int rc;
    const int nblk = (n + kDB - 1) / kDB, ldn = even(n);
    double *Lc = dense_scratch(h, 6, (size_t)n * ldn, &rc); if (rc) return rc;
    double *Lt = dense_scratch(h, 7, (size_t)n * ldn, &rc); if (rc) return rc;
    double *dinv = dense_scratch(h, 8, (size_t)nblk * kDB * kDB, &rc); if (rc) return rc;
    double *dinvT = dense_scratch(h, 9, (size_t)nblk * kDB * kDB, &rc); if (rc) return rc;
    double *Xt = dense_scratch(h, 10, (size_t)nrhs * ldn, &rc); if (rc) return rc;
    if ((rc = ensure(h, h->dinfo, 64)) != NK_OK) return rc;
    NK_CUDA(h, cudaMemcpy2DAsync(Lc, ldn * 8, L, ldl * 8, (size_t)n * 8, n, cudaMemcpyDeviceToDevice, stream));
    // diagonal-block inverses straight from the given factor (do_factor = 0)
    for (int i = 0; i < nblk; i++) {
        const int j0 = i * kDB, nb = (n - j0 < kDB) ? n - j0 : kDB;
        launch_diag(h, 1, Lc + (long long)j0 * ldn + j0, ldn, 0, nb, j0, dinv + (size_t)i * kDB * kDB, dinvT + (size_t)i * kDB * kDB, 0,
                    (int *)h->dinfo.ptr, 0, stream);
    }
    dim3 grid, block; launch2d(n, n, grid, block);
    zero_upper_kernel<<<grid, block, 0, stream>>>(n, Lc, ldn, 0);
    transpose(h, n, n, Lc, ldn, Lt, ldn, stream);
    transpose(h, n, nrhs, B, ldb, Xt, ldn, stream);
    if (!trans) trsm_fwd_t(h, n, nrhs, Lc, ldn, dinv, Xt, ldn, stream);
    else trsm_bwd_t(h, n, nrhs, Lt, ldn, dinvT, Xt, ldn, stream);
    transpose(h, nrhs, n, Xt, ldn, B, ldb, stream);
    NK_CUDA(h, cudaGetLastError());
    return NK_OK;
}

int nk_sym_sqrt(nk_handle *h, int n, const double *K, long long ldk, double lambda_min_bound, double *S, long long lds,
                double *Sinv, long long ldsi, int *iters, void *stream_) {
    if (!h) return NK_E_INVALID;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n < 1 || !K || !S || !Sinv || !(lambda_min_bound > 0.0)) return set_err(h, NK_E_INVALID, "nk_sym_sqrt: bad argument");
    NK_ON_DEVICE(h);
    int rc;
    if (n <= kSN && ldk >= n && lds >= n && ldsi >= n) {
        // small matrices: factor + inverse of the factor (one CTA), then the whole polar iteration in one cooperative launch
        const size_t mm = (size_t)kSN * kSN;
        double *buf = dense_scratch(h, 0, 9 * mm, &rc); if (rc) return rc;
        double *Kc = buf, *Lc = buf + mm, *Li = buf + 2 * mm, *LiT = buf + 3 * mm, *X = buf + 4 * mm, *Xt = buf + 5 * mm, *X2 = buf + 6 * mm,
               *Xt2 = buf + 7 * mm, *T = buf + 8 * mm;
        if ((rc = ensure(h, h->dinfo, 256)) != NK_OK) return rc;
        int *dinfo = (int *)h->dinfo.ptr;
        NK_CUDA(h, cudaMemsetAsync(dinfo, 0, 16, stream));
        NK_CUDA(h, cudaMemcpy2DAsync(Kc, kSN * 8, K, ldk * 8, (size_t)n * 8, n, cudaMemcpyDeviceToDevice, stream));
        NK_CUDA(h, cudaMemcpy2DAsync(Lc, kSN * 8, K, ldk * 8, (size_t)n * 8, n, cudaMemcpyDeviceToDevice, stream));
        launch_diag(h, 1, Lc, kSN, 0, n, 0, Li, LiT, 0, dinfo, 1, stream);
        static unsigned long long configured = 0;
        if (first_use_on_device(configured)) cudaFuncSetAttribute(small_polar_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmallSmem);
        int nn = n;
        double lmb = lambda_min_bound;
        const double *cKc = Kc, *cL = Lc, *cLiT = LiT;
        long long lds_ = lds, ldsi_ = ldsi;
        int *it_out = dinfo + 1;
        void *args[] = {&nn, &lmb, &cKc, &cL, &cLiT, &X, &Xt, &X2, &Xt2, &T, &S, &lds_, &Sinv, &ldsi_, &it_out};
        NK_CUDA(h, cudaLaunchCooperativeKernel((const void *)small_polar_kernel, dim3(16), dim3(256), args, kSmallSmem, stream));
        h->launches++;
        int hres[2] = {0, 0};
        NK_CUDA(h, cudaMemcpyAsync(hres, dinfo, 8, cudaMemcpyDeviceToHost, stream));
        NK_CUDA(h, cudaStreamSynchronize(stream));
        if (hres[0] != 0) return set_err(h, NK_E_NOT_SPD, "nk_sym_sqrt: matrix is not positive definite (pivot " + std::to_string(hres[0]) + ")");
        if (iters) *iters = hres[1];
        return NK_OK;
    }
    const int ldn = even(n), nblk = (n + kDB - 1) / kDB;
    const size_t nn = (size_t)n * ldn;
    double *L = dense_scratch(h, 0, nn, &rc); if (rc) return rc;
    double *Lt = dense_scratch(h, 1, nn, &rc); if (rc) return rc;
    double *X = dense_scratch(h, 2, nn, &rc); if (rc) return rc;
    double *Xt = dense_scratch(h, 3, nn, &rc); if (rc) return rc;
    double *X2 = dense_scratch(h, 4, nn, &rc); if (rc) return rc;
    double *Xt2 = dense_scratch(h, 5, nn, &rc); if (rc) return rc;
    double *T = dense_scratch(h, 6, nn, &rc); if (rc) return rc;
    double *dinv = dense_scratch(h, 8, (size_t)nblk * kDB * kDB, &rc); if (rc) return rc;
    double *dinvT = dense_scratch(h, 9, (size_t)nblk * kDB * kDB, &rc); if (rc) return rc;
    if ((rc = ensure(h, h->dinfo, 256)) != NK_OK) return rc;
    int *dinfo = (int *)h->dinfo.ptr;
    double *dnorm = (double *)((char *)h->dinfo.ptr + 16);

    NK_CUDA(h, cudaMemcpy2DAsync(L, ldn * 8, K, ldk * 8, (size_t)n * 8, n, cudaMemcpyDeviceToDevice, stream));
    NK_CUDA(h, cudaMemsetAsync(dnorm, 0, 8, stream));
    rowsum_max_kernel<<<n, 256, 0, stream>>>(n, L, ldn, dnorm);
    potrf_blocked(h, n, L, ldn, Lt, ldn, dinv, dinvT, dinfo, stream);
    // Spectrum estimate (large matrices only: it costs ~12 sweeps of small launches).  The Newton-Schulz schedule below starts from
    // a lower bound l on sigma_min(L^T) / ||L^T||; the only GUARANTEED one is sqrt(lambda_min_bound) -- the jitter, 1e-6 -- which at
    // m = 4096 on the benchmark's kernel matrix (cond 2e3) is 1000 x too pessimistic and costs 6-7 of 17 iterations.  A few steps of
    // inverse iteration with the factor just computed (8 fixed +-1 start vectors, v <- K^-1 v through the transposed-storage
    // triangular sweeps) give lambda_est = min_r |v_3| / |v_4| >= lambda_min, within a few percent after 4 steps; the schedule uses a
    // quarter of it, and the residual check after the scheduled iterations catches an estimate that was still too optimistic.
    constexpr int kProbeRows = 8, kProbeSteps = 4;
    const bool estimate = n >= 1024;
    double *dprobe = dnorm + 1;     // [1..8]: |v_5|, [9..16]: |v_6|, [17]: residual of the convergence check
    if (estimate) {
        double *V = X;              // (8, n) in the scratch the iteration will overwrite afterwards
        dim3 g2((n + 127) / 128, kProbeRows);
        probe_init_kernel<<<g2, 128, 0, stream>>>(kProbeRows, n, V, ldn);
        for (int st = 0; st < kProbeSteps; st++) {
            if (st == kProbeSteps - 1) row_norm_kernel<<<kProbeRows, 256, 0, stream>>>(n, V, ldn, dprobe);
            trsm_fwd_t(h, n, kProbeRows, L, ldn, dinv, V, ldn, stream);
            trsm_bwd_t(h, n, kProbeRows, Lt, ldn, dinvT, V, ldn, stream);
        }
        row_norm_kernel<<<kProbeRows, 256, 0, stream>>>(n, V, ldn, dprobe + kProbeRows);
        h->launches += 3;
    }
    struct { int info; int pad[3]; double nrm2; double probe[2 * kProbeRows]; } host;
    NK_CUDA(h, cudaMemcpyAsync(&host, h->dinfo.ptr, sizeof(host), cudaMemcpyDeviceToHost, stream));
    NK_CUDA(h, cudaStreamSynchronize(stream));
    if (host.info != 0) return set_err(h, NK_E_NOT_SPD, "nk_sym_sqrt: matrix is not positive definite (pivot " + std::to_string(host.info) + ")");
    const double nrm = std::sqrt(host.nrm2);   // >= ||L^T||_2
    double lam_lo = lambda_min_bound;
    if (estimate) {
        double lam_est = 1e300;
        for (int r = 0; r < kProbeRows; r++)
            if (host.probe[kProbeRows + r] > 0.0 && std::isfinite(host.probe[kProbeRows + r])) lam_est = std::min(lam_est, host.probe[r] / host.probe[kProbeRows + r]);
        if (lam_est < 1e300 && 0.25 * lam_est > lam_lo) lam_lo = 0.25 * lam_est;
    }
    dim3 grid, block; launch2d(n, n, grid, block);
    scale_copy_kernel<<<grid, block, 0, stream>>>(n, n, 1.0 / nrm, Lt, ldn, X, ldn);
    scale_copy_kernel<<<grid, block, 0, stream>>>(n, n, 1.0 / nrm, L, ldn, Xt, ldn);
    h->launches += 3;

    // Newton-Schulz polar iteration  X <- (a I - b X X^T) X  with the minimax cubic while the singular-value lower
    // bound l < 1, then plain (1.5, 0.5) steps (quadratic convergence) until l reaches 1 to rounding.
    double l = 0.9 * std::sqrt(lam_lo) / nrm;
    if (l > 1.0) l = 1.0;
    int it = 0, plain = 0;
    while (it < 100) {
        double a, b;
        if (l < 0.999) l = opt_cubic(l, &a, &b);
        else { a = 1.5; b = 0.5; plain++; }
        gemm_nt(h, n, n, n, -b, X, ldn, X, ldn, 0.0, T, ldn, a, kGemmLowerOnly | kGemmMirror, nullptr, 0, stream);
        gemm_nt(h, n, n, n, 1.0, T, ldn, Xt, ldn, 0.0, X2, ldn, 0.0, kGemmStoreT, Xt2, ldn, stream);
        std::swap(X, X2); std::swap(Xt, Xt2);
        it++;
        if (plain >= 3) {
            if (!estimate) break;
            // residual check: the T of this (plain) step was 1.5 I - 0.5 X X^T of the PREVIOUS iterate, so max|T - I| = r/2 with r the
            // orthogonality residual before the step, and the step squares it: r <= 2e-8 means converged to rounding.
            NK_CUDA(h, cudaMemsetAsync(dprobe + 2 * kProbeRows, 0, 8, stream));
            max_dev_identity_kernel<<<n, 256, 0, stream>>>(n, T, ldn, dprobe + 2 * kProbeRows);
            h->launches++;
            double resid = 0.0;
            NK_CUDA(h, cudaMemcpyAsync(&resid, dprobe + 2 * kProbeRows, 8, cudaMemcpyDeviceToHost, stream));
            NK_CUDA(h, cudaStreamSynchronize(stream));
            if (resid <= 1e-8 || it >= 60) break;
            plain = 2;          // not there yet (the estimate was too optimistic): one more plain step, check again
        }
    }
    if (iters) *iters = it;
    // S = Q^T L^T  ->  S(i,j) = sum_k Xt(i,k) L(j,k); symmetrised
    gemm_nt(h, n, n, n, 1.0, Xt, ldn, L, ldn, 0.0, T, ldn, 0.0, 0, nullptr, 0, stream);
    symmetrize_kernel<<<grid, block, 0, stream>>>(n, T, ldn, S, lds);
    // S^-1 = L^-T Q  ->  (S^-1)^T = Q^T L^-1 : backward substitution in transposed storage
    trsm_bwd_t(h, n, n, Lt, ldn, dinvT, Xt, ldn, stream);
    symmetrize_kernel<<<grid, block, 0, stream>>>(n, Xt, ldn, Sinv, ldsi);
    h->launches += 2;
    NK_CUDA(h, cudaGetLastError());
    return NK_OK;
}

// ---- assemble kernels for the two regularised systems ----
__global__ void assemble_inner_kernel(int m, int p, double gn, double jitter, const double *Gxx, long long ldxx, const double *Gxu,
                                      long long ldxu, const double *Guu, long long lduu, const double *Kin, long long ldk, double *inner,
                                      long long ld) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
    const int N1 = m + p;
    if (c >= N1) return;
    double v;
    if (r < m && c < m) v = Gxx[(long long)r * ldxx + c] + gn * (Kin[(long long)r * ldk + c] + (r == c ? jitter : 0.0));
    else if (r < m) v = Gxu[(long long)r * ldxu + (c - m)];
    else if (c < m) v = Gxu[(long long)c * ldxu + (r - m)];
    else v = Guu[(long long)(r - m) * lduu + (c - m)] + (r == c ? gn : 0.0);
    inner[(long long)r * ld + c] = v;
}
__global__ void assemble_rec_kernel(int m, double gn, double jitter, const double *Gyy, long long ldyy, const double *Kzz, long long ldk,
                                    double *out, long long ld) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
    if (c >= m) return;
    out[(long long)r * ld + c] = gn * (Kzz[(long long)r * ldk + c] + (r == c ? jitter : 0.0)) + Gyy[(long long)r * ldyy + c];
}
// rows [row0, row0 + rows) of right^T = blkdiag(S^-1 Kzz, I_p) when the input landmarks are the output landmarks:
// S^-1 Kzz = S^-1 (S^2 - jitter I) = S - jitter S^-1  (K_mm = Kzz + jitter I = S^2) -- elementwise, and more accurate than the product
__global__ void right_rows_kernel(int m, int p, int row0, int rows, double jitter, const double *S, long long lds, const double *Sinv,
                                  long long ldsi, double *Rt, long long ld) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x, rr = blockIdx.y;
    if (c >= m + p || rr >= rows) return;
    const int r = row0 + rr;
    double v = 0.0;
    if (r < m && c < m) v = S[(long long)r * lds + c] - jitter * Sinv[(long long)r * ldsi + c];
    else if (r >= m && c == r) v = 1.0;
    Rt[(long long)rr * ld + c] = v;
}
// rows of right^T below / right of the landmark block when that block came from a product (distinct input landmarks)
__global__ void right_pad_kernel(int m, int p, int row0, int rows, double *Rt, long long ld) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x, rr = blockIdx.y;
    if (c >= m + p || rr >= rows) return;
    const int r = row0 + rr;
    if (r < m && c < m) return;
    Rt[(long long)rr * ld + c] = (r >= m && c == r) ? 1.0 : 0.0;
}

static int check_solve_args(nk_handle *h, int m, int p, int d, const nk_grams *G, const nk_landmarks *L, const char *who) {
    if (m < 1 || p < 0 || d < 1 || !G || !L || !G->Gxx || !G->Gyx || !G->Gyy || !G->GYy || !L->Kzz || !L->S || !L->Sinv ||
        (p && (!G->Gxu || !G->Gyu || !G->Guu)) || ((L->Kzz_in == nullptr) != (L->Kio == nullptr)))
        return set_err(h, NK_E_INVALID, std::string(who) + ": bad argument");
    if (G->ld_gxx < m || G->ld_gyx < m || G->ld_gyy < m || G->ld_gYy < m || L->ld_kzz < m || L->ld_s < m || L->ld_sinv < m ||
        (p && (G->ld_gxu < p || G->ld_gyu < p || G->ld_guu < p)) || (L->Kzz_in && (L->ld_kzz_in < m || L->ld_kio < m)))
        return set_err(h, NK_E_INVALID, std::string(who) + ": a leading dimension is smaller than the row length");
    return NK_OK;
}

int nk_solve_abc_part(nk_handle *h, int m, int p, int d, double gamma_n, double jitter, const nk_grams *G, const nk_landmarks *L,
                      int g_row0, int g_rows, double *GT, long long ld_gt, int c_row0, int c_rows, double *CT, long long ld_ct,
                      int *info, void *stream_) {
    if (!h) return NK_E_INVALID;
    cudaStream_t stream = (cudaStream_t)stream_;
    int rc = check_solve_args(h, m, p, d, G, L, "nk_solve_abc_part");
    if (rc) return rc;
    const int N1 = m + p;
    if (g_row0 < 0 || g_rows < 0 || g_row0 + g_rows > N1 || c_row0 < 0 || c_rows < 0 || c_row0 + c_rows > m || (g_rows && (!GT || ld_gt < m)) ||
        (c_rows && (!CT || ld_ct < d)))
        return set_err(h, NK_E_INVALID, "nk_solve_abc_part: bad row range or output");
    NK_ON_DEVICE(h);
    if (info) *info = 0;
    const int ld1 = even(N1), ldm = even(m), nblk1 = (N1 + kDB - 1) / kDB;
    double *inner = dense_scratch(h, 0, (size_t)N1 * ld1, &rc); if (rc) return rc;
    double *Lt = dense_scratch(h, 1, (size_t)N1 * ld1, &rc); if (rc) return rc;
    double *Rt = dense_scratch(h, 2, (size_t)(g_rows > c_rows ? g_rows : c_rows) * ld1 + 2, &rc); if (rc) return rc;
    double *crossT = dense_scratch(h, 3, (size_t)N1 * ldm, &rc); if (rc) return rc;
    double *left = dense_scratch(h, 4, (size_t)m * ld1, &rc); if (rc) return rc;
    double *dinv = dense_scratch(h, 8, (size_t)nblk1 * kDB * kDB, &rc); if (rc) return rc;
    double *dinvT = dense_scratch(h, 9, (size_t)nblk1 * kDB * kDB, &rc); if (rc) return rc;
    if ((rc = ensure(h, h->dinfo, 64)) != NK_OK) return rc;
    int *dinfo = (int *)h->dinfo.ptr;      // [0]: inner_term, [1]: inner_term_rec
    dim3 grid, block;
    const bool same_landmarks = (L->Kzz_in == nullptr);
    const double *Kin = same_landmarks ? L->Kzz : L->Kzz_in;
    const long long ldkin = same_landmarks ? L->ld_kzz : L->ld_kzz_in;
    NK_CUDA(h, cudaMemsetAsync(dinfo, 0, 2 * sizeof(int), stream));

    // ---- reconstruction: C = GYy (gn Kmm + Gyy)^-1 S   (regressors.py:162-166), rows of C^T = columns of C ----
    // Independent of the dynamics solve, and both are latency-bound (a chain of 128x128 diagonal-block kernels): when both are
    // asked for, this one is queued FIRST, on the handle's side stream with its own scratch (forked from the caller's stream here,
    // joined after the dynamics block), so that the two chains interleave on the device.
    if (c_rows > 0) {
        const bool concurrent = g_rows > 0;
        cudaStream_t st2 = stream;
        double *rec, *Lt2, *Pt, *dinv2, *dinvT2;
        if (concurrent) {
            if (!h->side_stream) {
                NK_CUDA(h, cudaStreamCreateWithFlags(&h->side_stream, cudaStreamNonBlocking));
                NK_CUDA(h, cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
                NK_CUDA(h, cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
            }
            st2 = h->side_stream;
            rec = dense_scratch(h, 16, (size_t)m * ldm, &rc); if (rc) return rc;
            Lt2 = dense_scratch(h, 17, (size_t)m * ldm, &rc); if (rc) return rc;
            Pt = dense_scratch(h, 18, (size_t)c_rows * ldm + 2, &rc); if (rc) return rc;
            dinv2 = dense_scratch(h, 19, (size_t)nblk1 * kDB * kDB, &rc); if (rc) return rc;
            dinvT2 = dense_scratch(h, 20, (size_t)nblk1 * kDB * kDB, &rc); if (rc) return rc;
        } else {
            rec = inner; Lt2 = Lt; Pt = Rt; dinv2 = dinv; dinvT2 = dinvT;
        }
        if (concurrent) {      // fork: everything already queued on the caller's stream (the Grams!) precedes the side stream's work
            NK_CUDA(h, cudaEventRecord(h->ev_fork, stream));
            NK_CUDA(h, cudaStreamWaitEvent(st2, h->ev_fork, 0));
        }
        launch2d(m, m, grid, block);
        assemble_rec_kernel<<<grid, block, 0, st2>>>(m, gamma_n, jitter, G->Gyy, G->ld_gyy, L->Kzz, L->ld_kzz, rec, ldm);
        h->launches++;
        potrf_blocked(h, m, rec, ldm, Lt2, ldm, dinv2, dinvT2, dinfo + 1, st2);
        NK_CUDA(h, cudaMemcpy2DAsync(Pt, (size_t)ldm * 8, L->S + (long long)c_row0 * L->ld_s, (size_t)L->ld_s * 8, (size_t)m * 8, c_rows,
                                     cudaMemcpyDeviceToDevice, st2));                            // rows of S^T = S
        trsm_fwd_t(h, m, c_rows, rec, ldm, dinv2, Pt, ldm, st2);
        trsm_bwd_t(h, m, c_rows, Lt2, ldm, dinvT2, Pt, ldm, st2);                                // rows of (inner_rec^-1 S)^T
        gemm_nt(h, c_rows, d, m, 1.0, Pt, ldm, G->GYy, G->ld_gYy, 0.0, CT, ld_ct, 0.0, 0, nullptr, 0, st2);   // C^T rows
        if (concurrent) NK_CUDA(h, cudaEventRecord(h->ev_join, st2));
    }
    // ---- dynamics: G = S^-1 [Gyx|Gyu] inner^-1 blkdiag(K_io S^-1, I)   (regressors.py:147-159), rows of G^T = columns of G ----
    if (g_rows > 0) {
        launch2d(N1, N1, grid, block);
        assemble_inner_kernel<<<grid, block, 0, stream>>>(m, p, gamma_n, jitter, G->Gxx, G->ld_gxx, G->Gxu, G->ld_gxu, G->Guu, G->ld_guu, Kin,
                                                          ldkin, inner, ld1);
        h->launches++;
        potrf_blocked(h, N1, inner, ld1, Lt, ld1, dinv, dinvT, dinfo, stream);
        // rows of right^T = blkdiag(S^-1 K_io^T, I)
        launch2d(g_rows, N1, grid, block);
        if (same_landmarks) {
            right_rows_kernel<<<grid, block, 0, stream>>>(m, p, g_row0, g_rows, jitter, L->S, L->ld_s, L->Sinv, L->ld_sinv, Rt, ld1);
        } else {
            const int lm_rows = (g_row0 < m) ? ((g_row0 + g_rows < m ? g_row0 + g_rows : m) - g_row0) : 0;
            if (lm_rows > 0)
                gemm_nt(h, lm_rows, m, m, 1.0, L->Sinv + (long long)g_row0 * L->ld_sinv, L->ld_sinv, L->Kio, L->ld_kio, 0.0, Rt, ld1, 0.0, 0, nullptr, 0, stream);
            right_pad_kernel<<<grid, block, 0, stream>>>(m, p, g_row0, g_rows, Rt, ld1);
        }
        h->launches++;
        trsm_fwd_t(h, N1, g_rows, inner, ld1, dinv, Rt, ld1, stream);                         // sol^T rows = right^T rows inner^-1
        trsm_bwd_t(h, N1, g_rows, Lt, ld1, dinvT, Rt, ld1, stream);
        // left = S^-1 [Gyx | Gyu]  (regressors.py:153), formed BEFORE the product with sol, as the reference does: cross * sol
        // cancels by a factor cond(inner_term), and applying S^-1 to that product afterwards would amplify its rounding errors by
        // ||S^-1|| (measured on the hjb configuration, cond(K_mm) 7e7: 350 x further from a high-precision solve).  S^-1 cross
        // itself is benign: the columns of cross lie in the range of the kernel matrix.  Every device forms all of `left`
        // (2 m^2 (m+p) flop, 4.5 ms at m = 4096); only the product with its own columns of sol is sharded.
        transpose(h, m, m, G->Gyx, G->ld_gyx, crossT, ldm, stream);                               // cross^T = [Gyx | Gyu]^T
        if (p) transpose(h, m, p, G->Gyu, G->ld_gyu, crossT + (long long)m * ldm, ldm, stream);
        gemm_nt(h, m, N1, m, 1.0, L->Sinv, L->ld_sinv, crossT, ldm, 0.0, left, ld1, 0.0, 0, nullptr, 0, stream);   // left = S^-1 cross
        // G^T rows = sol^T rows * left^T      [G = left sol]
        gemm_nt(h, g_rows, m, N1, 1.0, Rt, ld1, left, ld1, 0.0, GT, ld_gt, 0.0, 0, nullptr, 0, stream);
    }
    if (g_rows > 0 && c_rows > 0) NK_CUDA(h, cudaStreamWaitEvent(stream, h->ev_join, 0));     // join the reconstruction solve

    int hinfo[2] = {0, 0};
    NK_CUDA(h, cudaMemcpyAsync(hinfo, dinfo, 2 * sizeof(int), cudaMemcpyDeviceToHost, stream));
    NK_CUDA(h, cudaStreamSynchronize(stream));      // the one synchronisation of this call: the two Cholesky verdicts
    NK_CUDA(h, cudaGetLastError());
    if ((rc = gram_watchdog_verdict(h)) != NK_OK) return rc;
    if (hinfo[0] != 0) { if (info) *info = 1; return set_err(h, NK_E_NOT_SPD, "nk_solve_abc: inner_term is not positive definite (pivot " + std::to_string(hinfo[0]) + ")"); }
    if (hinfo[1] != 0) { if (info) *info = 2; return set_err(h, NK_E_NOT_SPD, "nk_solve_abc: inner_term_rec is not positive definite (pivot " + std::to_string(hinfo[1]) + ")"); }
    return NK_OK;
}

int nk_solve_abc_finish(nk_handle *h, int m, int p, int d, const double *GT, long long ld_gt, const double *CT, long long ld_ct,
                        double *A, long long lda, double *B, long long ldb, double *C, long long ldc, double *W, long long ldw, void *stream_) {
    if (!h) return NK_E_INVALID;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (m < 1 || p < 0 || d < 1 || !GT || !CT || ld_gt < m || ld_ct < d || (A && lda < m) || (p && B && ldb < p) || (C && ldc < m) || (W && (ldw < m + p || !C)))
        return set_err(h, NK_E_INVALID, "nk_solve_abc_finish: bad argument");
    NK_ON_DEVICE(h);
    if (A) transpose(h, m, m, GT, ld_gt, A, lda, stream);                                        // A = G[:, :m]
    if (p && B) transpose(h, p, m, GT + (long long)m * ld_gt, ld_gt, B, ldb, stream);             // B = G[:, m:]
    if (C) transpose(h, m, d, CT, ld_ct, C, ldc, stream);
    if (W) gemm_nt(h, d, m + p, m, 1.0, C, ldc, GT, ld_gt, 0.0, W, ldw, 0.0, 0, nullptr, 0, stream);   // weights = C G (regressors.py:167)
    NK_CUDA(h, cudaGetLastError());
    return NK_OK;
}

int nk_solve_abc(nk_handle *h, int m, int p, int d, double gamma_n, double jitter, const nk_grams *G, const nk_landmarks *L,
                 double *A, long long lda, double *B, long long ldb, double *C, long long ldc, double *W, long long ldw, int *info,
                 void *stream_) {
    if (!h) return NK_E_INVALID;
    if (!A || !C || !W || (p && !B)) return set_err(h, NK_E_INVALID, "nk_solve_abc: bad argument");
    int rc;
    const int N1 = m + p, ldm = even(m), ldd = even(d);
    double *GT, *CT;
    {
        NK_ON_DEVICE(h);
        GT = dense_scratch(h, 5, (size_t)N1 * ldm, &rc); if (rc) return rc;
        CT = dense_scratch(h, 6, (size_t)m * ldd, &rc); if (rc) return rc;
    }
    if ((rc = nk_solve_abc_part(h, m, p, d, gamma_n, jitter, G, L, 0, N1, GT, ldm, 0, m, CT, ldd, info, stream_)) != NK_OK) return rc;
    return nk_solve_abc_finish(h, m, p, d, GT, ldm, CT, ldd, A, lda, B, ldb, C, ldc, W, ldw, stream_);
}

int nk_kernel_cross(nk_handle *h, const double *Z, long long ldz, int m, int d, const double *inv_ls, int kind,
                    const double *X, long long ldx, long long N, double *K, long long ldk, void *stream_) {
    if (!h) return NK_E_INVALID;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!Z || !X || !K || !inv_ls || m < 1 || d < 1 || N < 1 || N > 2000000000LL) return set_err(h, NK_E_INVALID, "nk_kernel_cross: bad argument");
    if (kind != NK_KERNEL_RBF && kind != NK_KERNEL_MATERN52) return set_err(h, NK_E_INVALID, "nk_kernel_cross: unsupported kernel kind");
    NK_ON_DEVICE(h);
    int rc;
    const int KA = even(d + 2);
    double *Za = dense_scratch(h, 0, (size_t)m * KA, &rc); if (rc) return rc;
    double *Xa = dense_scratch(h, 1, (size_t)N * KA, &rc); if (rc) return rc;
    double *ctr = dense_scratch(h, 7, (size_t)d, &rc); if (rc) return rc;
    landmark_center(Z, ldz, m, d, ctr, stream);
    augment_rows(h, Z, ldz, m, d, inv_ls, ctr, 1, Za, KA, stream);
    augment_rows(h, X, ldx, N, d, inv_ls, ctr, 0, Xa, KA, stream);
    const bool same = (Z == X && m == N);
    gemm_nt(h, m, (int)N, KA, 1.0, Za, KA, Xa, KA, 0.0, K, ldk, 0.0, same ? (kGemmUnitDiag | kGemmLowerOnly | kGemmMirror) : 0, nullptr, 0, stream, kind);
    NK_CUDA(h, cudaGetLastError());
    return NK_OK;
}

int nk_kzz(nk_handle *h, const double *Z, long long ldz, int m, int d, const double *inv_ls, int kind, double *Kzz, long long ldk,
           void *stream_) {
    return nk_kernel_cross(h, Z, ldz, m, d, inv_ls, kind, Z, ldz, m, Kzz, ldk, stream_);
}

}  // extern "C"
