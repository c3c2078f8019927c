// nk_pgemm.cuh -- packed-operand persistent NT GEMM (the rollout / scoring workhorse), see nk_pgemm.cu
#pragma once
#include "nk_handle.cuh"

namespace nk {

// result(r, c) = alpha * sum_k A(r, k) B(c, k),  r < M (rows of A), c < N (rows of B), both operands in the packed 8x8-block
// layout of nk_common.cuh with zero padding (rows to a multiple of 128, depth to a multiple of 16).
// Result columns c < cp_cols go to a PACKED destination whose contraction index is the result column (so the result is the
// next product's A operand without a repack); columns c >= c_col0 go to a row-major destination (+ beta * old value).
struct PGemmParams {
    int M, N, KS;                       // KS = contraction slabs of 16
    const double *Ap; int a_rp;         // a_rp / b_rp: 8-row panels of the operand (= padded rows / 8)
    const double *Bp; int b_rp;
    double alpha, beta;
    double *C; long long ldc; int c_col0;
    double *Cp; int c_rp; int cp_cols;
    int tiles_m, tiles_n;
};

void launch_pgemm(nk_handle *h, const PGemmParams &P, cudaStream_t stream);

// dst[packed_off(row0 + r, k0 + c)] = src[r * ld + c]   (r < rows, c < cols); the destination's padding is left untouched
void pack_rows(nk_handle *h, const double *src, long long ld, long long rows, int cols, double *dst, int rp, long long row0, int k0,
               cudaStream_t stream);
// dst[r * ld + c] = src[packed_off(row0 + r, k0 + c)]
void unpack_rows(nk_handle *h, const double *src, int rp, long long row0, int k0, long long rows, int cols, double *dst, long long ld,
                 cudaStream_t stream);

inline long long pad_to(long long x, int q) { return (x + q - 1) / q * q; }

}  // namespace nk
