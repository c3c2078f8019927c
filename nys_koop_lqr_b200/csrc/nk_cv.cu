// nk_cv.cu -- the cross-validation sweep over the (kernel lengthscale, regularisation) grid
// (benchmark_lqr_hjb.py:47-71, benchmark_lqr_classic.py:44-64, benchmark_lqr_cloth.py:39-66: sklearn GridSearchCV
//  cloning KoopmanNystromRegressor per candidate and fold, scoring regressors.py:48-55 `predict` by RMSE).
//
// What is shared and what is not.  For one kernel and one training fold every regularisation value gamma sees the same
// seven Grams; gamma only enters as gamma*n*K_mm (regressors.py:151,162).  So per (kernel, fold) the nlam pairs of
// regularised systems are factored as ONE batch (batched blocked Cholesky of nk_dense.cu: the serial 128x128
// diagonal-block kernel of one matrix overlaps with the panel/trailing GEMMs of the others) and solved with only
// d right-hand sides, because scoring needs the one-step prediction weights, not A, B, C themselves:
//
//   weights [S^-1 k(Z,x); u] = GYy (gn Kmm + Gyy)^-1 [Gyx|Gyu] inner^-1 [Kzz Kmm^-1 k(Z,x); u]        (S cancels)
//
// i.e. Wk = [V_phi Kzz Kmm^-1 | V_u] with V (d, m+p) = GYy inner_rec^-1 [Gyx|Gyu] inner^-1: four triangular sweeps with
// d rows instead of m+p, and no matrix square root anywhere in the sweep (it is only needed for the final refit).
// Scoring multiplies the validation fold's kernel rows by the stacked weights of all nlam values at once.
#include <vector>
#include "nk_dense.cuh"
#include "nk_pgemm.cuh"

namespace nk {

static inline int even_c(int x) { return (x + 1) & ~1; }
constexpr int kDBc = 128;

__global__ void axpy_kernel(long long n, double alpha, const double *x, double *y) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) y[i] += alpha * x[i];
}

// out(r,c) = Kzz(r,c) + jitter*[r==c]
__global__ void assemble_kmm_kernel(int m, double jitter, const double *Kzz, long long ldk, double *out, long long ld) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
    if (c < m) out[(long long)r * ld + c] = Kzz[(long long)r * ldk + c] + (r == c ? jitter : 0.0);
}
// batched over blockIdx.z: rec_b = gn_b (Kzz + jitter I) + Gyy
__global__ void assemble_rec_batched_kernel(int m, const double *gn, double jitter, const double *Gyy, long long ldyy, const double *Kzz,
                                            long long ldk, double *out, long long ld, long long stride) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
    if (c >= m) return;
    const double g = gn[blockIdx.z];
    out[(long long)blockIdx.z * stride + (long long)r * ld + c] = g * (Kzz[(long long)r * ldk + c] + (r == c ? jitter : 0.0)) + Gyy[(long long)r * ldyy + c];
}
// batched: inner_b = [[Gxx + gn_b Kmm, Gxu],[Gxu^T, Guu + gn_b I]]
__global__ void assemble_inner_batched_kernel(int m, int p, const double *gn, double jitter, const double *Gxx, long long ldxx,
                                              const double *Gxu, long long ldxu, const double *Guu, long long lduu, const double *Kzz,
                                              long long ldk, double *inner, long long ld, long long stride) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
    const int N1 = m + p;
    if (c >= N1) return;
    const double g = gn[blockIdx.z];
    double v;
    if (r < m && c < m) v = Gxx[(long long)r * ldxx + c] + g * (Kzz[(long long)r * ldk + c] + (r == c ? jitter : 0.0));
    else if (r < m) v = Gxu[(long long)r * ldxu + (c - m)];
    else if (c < m) v = Gxu[(long long)c * ldxu + (r - m)];
    else v = Guu[(long long)(r - m) * lduu + (c - m)] + (r == c ? g : 0.0);
    inner[(long long)blockIdx.z * stride + (long long)r * ld + c] = v;
}
// dst_b (rows, cols; ld, stride) = src (rows, cols; lds) for every b
__global__ void broadcast_rows_kernel(int rows, int cols, const double *src, long long lds, double *dst, long long ldd, long long stride) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
    if (c < cols && r < rows) dst[(long long)blockIdx.z * stride + (long long)r * ldd + c] = src[(long long)r * lds + c];
}

// per-column squared error of one block of predictions:  sse[c] += sum_s (Yhat[s,c] - Y[s, c % d])^2.
// 32 columns x 32 row lanes per CTA, fixed-order reduction over the lanes, one writer per column -> deterministic.
__global__ void sse_columns_kernel(long long rows, int R, int d, const double *Yhat, long long ldh, const double *Y, long long ldy, double *sse) {
    __shared__ double part[32][33];
    const int c = blockIdx.x * 32 + threadIdx.x;
    double s = 0.0;
    if (c < R) {
        const int j = c % d;
        for (long long r = threadIdx.y; r < rows; r += 32) {
            const double e = Yhat[r * ldh + c] - Y[r * ldy + j];
            s += e * e;
        }
    }
    part[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0 && c < R) {
        double t = 0.0;
        for (int k = 0; k < 32; k++) t += part[k][threadIdx.x];
        sse[c] += t;
    }
}

__global__ void kernel_function_kernel(long long n, int kind, const double *e, double *out) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) out[i] = kernel_from_exponent(e[i], kind);
}

int kernel_cross_t(nk_handle *h, const double *Z, long long ldz, int m, int d, const double *inv_ls, int kind,
                   const double *X, long long ldx, long long N, double *Kt, long long ldkt, cudaStream_t stream, int packed_rp);
void copy_cols(nk_handle *h, long long rows, int cols, const double *src, long long lds, double *dst, long long ldd, cudaStream_t stream);

}  // namespace nk

using namespace nk;

extern "C" {

int nk_kernel_function(nk_handle *h, int kind, long long count, const double *exponent, double *out, void *stream_) {
    if (!h) return NK_E_INVALID;
    if (count < 0 || !exponent || !out || (kind != NK_KERNEL_RBF && kind != NK_KERNEL_MATERN52)) return set_err(h, NK_E_INVALID, "nk_kernel_function: bad argument");
    if (count == 0) return NK_OK;
    NK_ON_DEVICE(h);
    const long long blocks = (count + 255) / 256;
    kernel_function_kernel<<<(unsigned)(blocks < 8192 ? blocks : 8192), 256, 0, (cudaStream_t)stream_>>>(count, kind, exponent, out);
    h->launches++;
    NK_CUDA(h, cudaGetLastError());
    return NK_OK;
}

int nk_axpy(nk_handle *h, long long count, double alpha, const double *x, double *y, void *stream_) {
    if (!h) return NK_E_INVALID;
    if (count < 0 || !x || !y) return set_err(h, NK_E_INVALID, "nk_axpy: bad argument");
    if (count == 0) return NK_OK;
    NK_ON_DEVICE(h);
    const long long blocks = (count + 255) / 256;
    axpy_kernel<<<(unsigned)(blocks < 8192 ? blocks : 8192), 256, 0, (cudaStream_t)stream_>>>(count, alpha, x, y);
    h->launches++;
    NK_CUDA(h, cudaGetLastError());
    return NK_OK;
}

int nk_cv_weights(nk_handle *h, int m, int p, int d, int nlam, const double *gamma_n, double jitter, const nk_grams *G,
                  const double *Kzz, long long ld_kzz, double *Wk, long long ld_wk, int *info, void *stream_) {
    if (!h) return NK_E_INVALID;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (m < 1 || p < 0 || d < 1 || nlam < 1 || nlam > 4096 || !gamma_n || !G || !G->Gxx || !G->Gyx || !G->Gyy || !G->GYy || !Kzz || !Wk ||
        (p && (!G->Gxu || !G->Gyu || !G->Guu)))
        return set_err(h, NK_E_INVALID, "nk_cv_weights: bad argument");
    if (G->ld_gxx < m || G->ld_gyx < m || G->ld_gyy < m || G->ld_gYy < m || ld_kzz < m || ld_wk < m + p || (p && (G->ld_gxu < p || G->ld_gyu < p || G->ld_guu < p)))
        return set_err(h, NK_E_INVALID, "nk_cv_weights: a leading dimension is smaller than the row length");
    const double *Gxx = G->Gxx, *Gyx = G->Gyx, *Gyy = G->Gyy, *Gxu = G->Gxu, *Gyu = G->Gyu, *Guu = G->Guu, *GYy = G->GYy;
    NK_ON_DEVICE(h);
    int rc;
    const int N1 = m + p, ld1 = even_c(N1), ldm = even_c(m);
    const int nblk1 = (N1 + kDBc - 1) / kDBc, nblkm = (m + kDBc - 1) / kDBc;
    const long long sM = (long long)N1 * ld1;             // one system matrix of the batch
    const long long sD = (long long)nblk1 * kDBc * kDBc;  // its diagonal-block inverses
    double *Lb = dense_scratch(h, 0, (size_t)nlam * sM, &rc); if (rc) return rc;
    double *Ltb = dense_scratch(h, 1, (size_t)nlam * sM, &rc); if (rc) return rc;
    double *Ta = dense_scratch(h, 2, (size_t)nlam * d * ldm, &rc); if (rc) return rc;
    double *crossT = dense_scratch(h, 3, (size_t)N1 * ldm, &rc); if (rc) return rc;
    double *Lk = dense_scratch(h, 4, (size_t)m * ldm, &rc); if (rc) return rc;
    double *Lkt = dense_scratch(h, 5, (size_t)m * ldm, &rc); if (rc) return rc;
    double *Tb = dense_scratch(h, 6, (size_t)nlam * d * ld1, &rc); if (rc) return rc;
    double *dgn = dense_scratch(h, 7, (size_t)nlam + d, &rc); if (rc) return rc;
    double *dinv = dense_scratch(h, 8, (size_t)nlam * sD, &rc); if (rc) return rc;
    double *dinvT = dense_scratch(h, 9, (size_t)nlam * sD, &rc); if (rc) return rc;
    double *dinvk = dense_scratch(h, 12, (size_t)nblkm * kDBc * kDBc, &rc); if (rc) return rc;
    double *dinvkT = dense_scratch(h, 13, (size_t)nblkm * kDBc * kDBc, &rc); if (rc) return rc;
    double *Tc = dense_scratch(h, 14, (size_t)nlam * d * ldm, &rc); if (rc) return rc;
    if ((rc = ensure(h, h->dinfo, sizeof(int) * (size_t)(2 * nlam + 16))) != NK_OK) return rc;
    int *dinfo = (int *)h->dinfo.ptr;
    NK_CUDA(h, cudaMemcpyAsync(dgn, gamma_n, sizeof(double) * nlam, cudaMemcpyHostToDevice, stream));

    const dim3 block(128);
    // (1) Kmm = Kzz + jitter I = Lk Lk^T (shared by the whole batch)
    assemble_kmm_kernel<<<dim3((m + 127) / 128, m), block, 0, stream>>>(m, jitter, Kzz, ld_kzz, Lk, ldm);
    h->launches++;
    potrf_batched(h, 1, m, Lk, ldm, 0, Lkt, ldm, 0, dinvk, dinvkT, 0, dinfo + 2 * nlam, stream);
    // (2) cross^T = [Gyx | Gyu]^T  ((m+p) x m)
    transpose(h, m, m, Gyx, G->ld_gyx, crossT, ldm, stream);
    if (p) transpose(h, m, p, Gyu, G->ld_gyu, crossT + (long long)m * ldm, ldm, stream);
    // (3) reconstruction systems  rec_b = gn_b Kmm + Gyy  (regressors.py:162), factored as one batch
    assemble_rec_batched_kernel<<<dim3((m + 127) / 128, m, nlam), block, 0, stream>>>(m, dgn, jitter, Gyy, G->ld_gyy, Kzz, ld_kzz, Lb, ldm, sM);
    h->launches++;
    potrf_batched(h, nlam, m, Lb, ldm, sM, Ltb, ldm, sM, dinv, dinvT, sD, dinfo, stream);
    // (4) Ta_b = GYy rec_b^-1   (d rows)
    broadcast_rows_kernel<<<dim3((m + 127) / 128, d, nlam), block, 0, stream>>>(d, m, GYy, G->ld_gYy, Ta, ldm, (long long)d * ldm);
    h->launches++;
    trsm_fwd_t_rl(h, nlam, m, d, Lb, ldm, sM, dinv, sD, Ta, ldm, (long long)d * ldm, stream);
    trsm_bwd_t_rl(h, nlam, m, d, Ltb, ldm, sM, dinvT, sD, Ta, ldm, (long long)d * ldm, stream);
    // (5) Tb_b = Ta_b [Gyx | Gyu]   (d x (m+p))
    gemm_nt_batched(h, nlam, d, N1, m, 1.0, Ta, ldm, (long long)d * ldm, crossT, ldm, 0, 0.0, Tb, ld1, (long long)d * ld1, 0.0, 0, nullptr, 0, 0, stream);
    // (6) dynamics systems inner_b (regressors.py:148,151), one batch;  V_b = Tb_b inner_b^-1
    assemble_inner_batched_kernel<<<dim3((N1 + 127) / 128, N1, nlam), block, 0, stream>>>(m, p, dgn, jitter, Gxx, G->ld_gxx, Gxu, G->ld_gxu, Guu,
                                                                                          G->ld_guu, Kzz, ld_kzz, Lb, ld1, sM);
    h->launches++;
    potrf_batched(h, nlam, N1, Lb, ld1, sM, Ltb, ld1, sM, dinv, dinvT, sD, dinfo + nlam, stream);
    trsm_fwd_t_rl(h, nlam, N1, d, Lb, ld1, sM, dinv, sD, Tb, ld1, (long long)d * ld1, stream);
    trsm_bwd_t_rl(h, nlam, N1, d, Ltb, ld1, sM, dinvT, sD, Tb, ld1, (long long)d * ld1, stream);
    // (7) Tc_b = V_phi,b Kzz Kmm^-1 : all nlam*d rows against the one factor of Kmm
    gemm_nt_batched(h, nlam, d, m, m, 1.0, Tb, ld1, (long long)d * ld1, Kzz, ld_kzz, 0, 0.0, Tc, ldm, (long long)d * ldm, 0.0, 0, nullptr, 0, 0, stream);
    trsm_fwd_t_rl(h, 1, m, nlam * d, Lk, ldm, 0, dinvk, 0, Tc, ldm, 0, stream);
    trsm_bwd_t_rl(h, 1, m, nlam * d, Lkt, ldm, 0, dinvkT, 0, Tc, ldm, 0, stream);
    // (8) Wk_b = [Tc_b | V_u,b]
    NK_CUDA(h, cudaMemcpy2DAsync(Wk, (size_t)ld_wk * 8, Tc, (size_t)ldm * 8, (size_t)m * 8, (size_t)nlam * d, cudaMemcpyDeviceToDevice, stream));
    if (p) NK_CUDA(h, cudaMemcpy2DAsync(Wk + m, (size_t)ld_wk * 8, Tb + m, (size_t)ld1 * 8, (size_t)p * 8, (size_t)nlam * d, cudaMemcpyDeviceToDevice, stream));
    std::vector<int> hinfo(2 * nlam + 1, 0);
    NK_CUDA(h, cudaMemcpyAsync(hinfo.data(), dinfo, sizeof(int) * (2 * nlam + 1), cudaMemcpyDeviceToHost, stream));
    NK_CUDA(h, cudaStreamSynchronize(stream));
    NK_CUDA(h, cudaGetLastError());
    if ((rc = gram_watchdog_verdict(h)) != NK_OK) return rc;
    int bad = 0;
    for (int b = 0; b < nlam; b++) {
        int v = 0;
        if (hinfo[2 * nlam] != 0) v = 3;
        else if (hinfo[b] != 0) v = 2;
        else if (hinfo[nlam + b] != 0) v = 1;
        if (info) info[b] = v;
        if (v) bad++;
    }
    if (bad) return set_err(h, NK_E_NOT_SPD, "nk_cv_weights: " + std::to_string(bad) + " of " + std::to_string(nlam) + " regularised systems are not positive definite (see info[])");
    return NK_OK;
}

int nk_cv_score(nk_handle *h, const double *Z, long long ldz, int m, int d, int p, const double *inv_ls, int kind, const double *Wk,
                int R, const double *X_aug, long long ldx, const double *Y, long long ldy, long long N, double *sse, void *stream_) {
    if (!h) return NK_E_INVALID;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!Z || !inv_ls || !Wk || !X_aug || !Y || !sse || m < 1 || d < 1 || p < 0 || R < 1 || N < 0 || ldx < d + p || ldy < d)
        return set_err(h, NK_E_INVALID, "nk_cv_score: bad argument");
    if (kind != NK_KERNEL_RBF && kind != NK_KERNEL_MATERN52) return set_err(h, NK_E_INVALID, "nk_cv_score: unsupported kernel kind");
    if (N == 0) return NK_OK;
    NK_ON_DEVICE(h);
    // Predictions of all stacked weight rows at once on the packed persistent GEMM (nk_pgemm.cu): the held-out block's kernel
    // rows are written straight into the packed operand layout by the kernel-lift GEMM epilogue, the controls are dropped
    // into the p spare contraction columns, the stacked weights are packed once per call.
    int rc;
    const int N1 = m + p, KS = (N1 + kSlabK - 1) / kSlabK, Kpad = KS * kSlabK;
    const int RP = (int)pad_to(R, kTile);
    long long nb = (1LL << 28) / (Kpad + R);        // rows per block: <= 2 GiB of scratch for [K^T | U] and the predictions
    nb = (nb / kTile) * kTile;
    if (nb < kTile) nb = kTile;
    if (nb > pad_to(N, kTile)) nb = pad_to(N, kTile);
    double *Fp = dense_scratch(h, 2, (size_t)nb * Kpad, &rc); if (rc) return rc;
    double *Yh = dense_scratch(h, 3, (size_t)nb * R, &rc); if (rc) return rc;
    double *Wp = dense_scratch(h, 4, (size_t)RP * Kpad, &rc); if (rc) return rc;
    NK_CUDA(h, cudaMemsetAsync(Fp, 0, (size_t)nb * Kpad * 8, stream));       // contraction padding must be exact zeros
    NK_CUDA(h, cudaMemsetAsync(Wp, 0, (size_t)RP * Kpad * 8, stream));
    pack_rows(h, Wk, N1, R, N1, Wp, RP / kPanel, 0, 0, stream);
    const int rp_f = (int)(nb / kPanel);
    for (long long s = 0; s < N; s += nb) {
        const long long rows = (N - s < nb) ? N - s : nb;
        const double *Xb = X_aug + s * ldx;
        if ((rc = kernel_cross_t(h, Z, ldz, m, d, inv_ls, kind, Xb, ldx, rows, Fp, 0, stream, rp_f)) != NK_OK) return rc;   // k(x, Z), packed
        if (p) pack_rows(h, Xb + d, ldx, rows, p, Fp, rp_f, 0, m, stream);                                               // u
        PGemmParams P;
        P.M = (int)rows; P.N = R; P.KS = KS; P.Ap = Fp; P.a_rp = rp_f; P.Bp = Wp; P.b_rp = RP / kPanel; P.alpha = 1.0; P.beta = 0.0;
        P.C = Yh; P.ldc = R; P.c_col0 = 0; P.Cp = nullptr; P.c_rp = 0; P.cp_cols = 0;
        P.tiles_m = (int)(pad_to(rows, kTile) / kTile); P.tiles_n = RP / kTile;
        launch_pgemm(h, P, stream);
        sse_columns_kernel<<<(R + 31) / 32, dim3(32, 32), 0, stream>>>(rows, R, d, Yh, R, Y + s * ldy, ldy, sse);
        h->launches++;
    }
    NK_CUDA(h, cudaGetLastError());
    return NK_OK;
}

}  // extern "C"
