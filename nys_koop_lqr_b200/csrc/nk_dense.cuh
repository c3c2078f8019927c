// nk_dense.cuh -- internal interface of the dense FP64 stage (DMMA GEMM, Cholesky, triangular solves, symmetric sqrt)
#pragma once
#include "nk_handle.cuh"

namespace nk {

enum GemmFlags : int {
    kGemmLowerOnly = 1,   // compute only tiles with row-block >= col-block (C symmetric, e.g. A A^T)
    kGemmStoreT = 2,      // store C^T (into Ct, ldct) in addition to C (C may be NULL)
    kGemmMirror = 4,      // with LowerOnly: also write the mirrored element so the full symmetric C is stored
    kGemmUnitDiag = 8,
    kGemmPackedOut = 16,  // C is a PACKED operand buffer (nk_common.cuh layout, contraction index = result column); ldc = its row panels
};

// C (M,N) = alpha * A (M,K) * B(N,K)^T + beta * C  [+ diag * I],  all row-major ("NT": both operands k-contiguous)
// epi_kind >= 0: C = kernel_from_exponent(acc, epi_kind) (kernel lift epilogue; alpha/beta ignored); flag kGemmUnitDiag
// forces C(i,i) = 1 (k(z,z) = 1 exactly, as the reference's direct-difference cdist gives).
void gemm_nt(nk_handle *h, int M, int N, int K, double alpha, const double *A, long long lda, const double *B, long long ldb,
             double beta, double *C, long long ldc, double diag, int flags, double *Ct, long long ldct, cudaStream_t stream,
             int epi_kind = -1);

void gemm_nt_batched(nk_handle *h, int batch, int M, int N, int K, double alpha, const double *A, long long lda, long long sA,
                     const double *B, long long ldb, long long sB, double beta, double *C, long long ldc, long long sC, double diag,
                     int flags, double *Ct, long long ldct, long long sCt, cudaStream_t stream, int epi_kind = -1);

// TMA-fed persistent variant (nk_tgemm.cu): launches the product and returns true when the operands qualify (16-byte aligned,
// even strides, enough tiles to fill the GPU); gemm_nt_batched tries it first.
bool tgemm_try(nk_handle *h, int batch, int M, int N, int K, double alpha, const double *A, long long lda, long long sA, const double *B,
               long long ldb, long long sB, double beta, double *C, long long ldc, long long sC, double diag, int flags, double *Ct,
               long long ldct, long long sCt, cudaStream_t stream, int epi_kind);

void transpose_batched(nk_handle *h, int batch, int rows, int cols, const double *src, long long lds, long long ss, double *dst,
                       long long ldd, long long sd, cudaStream_t stream);
void transpose(nk_handle *h, int rows, int cols, const double *src, long long lds, double *dst, long long ldd, cudaStream_t stream);

// Cholesky of the n x n SPD matrix in A (lower), in place; Lt (n,n) receives L^T (upper, row-major); inverses of the
// 128x128 diagonal blocks go to dinv (nblk x 128 x 128, row-major L_ii^-1) and dinvT (their transposes).  dinfo: device int.
int potrf_blocked(nk_handle *h, int n, double *A, long long lda, double *Lt, long long ldlt, double *dinv, double *dinvT,
                  int *dinfo, cudaStream_t stream);
// batched form: matrix b lives at A + b*sA (Lt + b*sLt, dinv + b*sD, ...); dinfo holds `batch` device ints
int potrf_batched(nk_handle *h, int batch, int n, double *A, long long lda, long long sA, double *Lt, long long ldlt, long long sLt,
                  double *dinv, double *dinvT, long long sD, int *dinfo, cudaStream_t stream, bool clean_upper = false);
// right-looking batched triangular solves for few right-hand sides (stride 0 = shared across the batch)
void trsm_fwd_t_rl(nk_handle *h, int batch, int n, int r, const double *L, long long ldl, long long sL, const double *dinv, long long sD,
                   double *Xt, long long ldx, long long sX, cudaStream_t stream);
void trsm_bwd_t_rl(nk_handle *h, int batch, int n, int r, const double *Lt, long long ldlt, long long sLt, const double *dinvT, long long sD,
                   double *Xt, long long ldx, long long sX, cudaStream_t stream);
// transposed-storage triangular solves with the factor above.  Xt (r, n) row-major holds B^T on entry.
//   forward : Xt <- Xt L^-T   (i.e. X = L^-1 B)        backward: Xt <- Xt L^-1   (i.e. X = L^-T B)
void trsm_fwd_t(nk_handle *h, int n, int r, const double *L, long long ldl, const double *dinv, double *Xt, long long ldx, cudaStream_t stream);
void trsm_bwd_t(nk_handle *h, int n, int r, const double *Lt, long long ldlt, const double *dinvT, double *Xt, long long ldx, cudaStream_t stream);
// scaled / centred / augmented rows for the kernel-lift GEMM: out (rows, KA): landmark form [z', 1, -|z'|^2/2, 0..] or
// sample form [x', -|x'|^2/2, 1, 0..];  z' = (z - center) * inv_ls.  KA = even(d + 2).
void augment_rows(nk_handle *h, const double *src, long long lds, long long rows, int d, const double *inv_ls, const double *center,
                  int landmark_form, double *out, int KA, cudaStream_t stream);
void landmark_center(const double *Z, long long ldz, int m, int d, double *center, cudaStream_t stream);

double *dense_scratch(nk_handle *h, int slot, size_t doubles, int *rc);

}  // namespace nk
