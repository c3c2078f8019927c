// nk_rollout.cu -- lift / predict (regressors.py:171-178, 48-55) and the batched open-loop rollout
// (benchmark_lqr_cloth.py:18-36; _classic.py:23-41; _hjb.py:23-44) on the DMMA NT-GEMM of nk_dense.cu.
//
// Points (samples or trajectories) are always ROWS, so every product is C = A B^T with contraction-contiguous
// operands: K^T = k(X, Z) (N,m) comes from the augmented-row GEMM with the kernel function as epilogue,
// Phi^T = K^T S^-1 (S^-1 symmetric), Yhat = [Phi^T | U] W^T, and a rollout step is Z_next = Z A^T (+ U_i B^T).
// The recurrence stays serial in time exactly like the reference loop; only trajectories are batched.
#include "nk_dense.cuh"
#include "nk_pgemm.cuh"

namespace nk {

static inline int even_i(int x) { return (x + 1) & ~1; }

// per-trajectory error sums for one time step: err[b] += sum_j (Yt[b,j]-Yh[b,j])^2, sim[b] += sum_j Yh[b,j]^2
__global__ void step_error_kernel(long long nb, int d, const double *Yh, const double *Yt, double *sq_err, double *sq_sim) {
    const long long b = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (b >= nb) return;
    double e = 0.0, s = 0.0;
    for (int j = lane; j < d; j += 32) {
        const double yh = Yh[b * d + j], df = Yt[b * d + j] - yh;
        e += df * df; s += yh * yh;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { e += __shfl_xor_sync(0xffffffffu, e, o); s += __shfl_xor_sync(0xffffffffu, s, o); }
    if (lane == 0) { sq_err[b] += e; sq_sim[b] += s; }
}

__global__ void copy_cols_kernel(long long rows, int cols, const double *src, long long lds, double *dst, long long ldd) {
    const long long total = rows * cols;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const long long r = idx / cols;
        const int c = (int)(idx % cols);
        dst[r * ldd + c] = src[r * lds + c];
    }
}
void copy_cols(nk_handle *h, long long rows, int cols, const double *src, long long lds, double *dst, long long ldd, cudaStream_t stream) {
    if (rows <= 0 || cols <= 0) return;
    const long long total = rows * cols;
    const int blocks = (int)((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
    copy_cols_kernel<<<blocks, 256, 0, stream>>>(rows, cols, src, lds, dst, ldd);
    h->launches++;
}

// out = a - b over a whole packed buffer (both operands share the layout, so the difference of two packed matrices is elementwise)
__global__ void sub_kernel(long long n, const double *a, const double *b, double *out) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) out[i] = a[i] - b[i];
}

// K^T = k(X, Z): (N, m), rows = points
// packed_rp > 0: Kt is a packed operand buffer with that many 8-row panels (rows = points, contraction index = landmark)
int kernel_cross_t(nk_handle *h, const double *Z, long long ldz, int m, int d, const double *inv_ls, int kind,
                   const double *X, long long ldx, long long N, double *Kt, long long ldkt, cudaStream_t stream, int packed_rp) {
    int rc;
    const int KA = even_i(d + 2);
    double *Za = dense_scratch(h, 0, (size_t)m * KA, &rc); if (rc) return rc;
    double *Xa = dense_scratch(h, 1, (size_t)N * KA, &rc); if (rc) return rc;
    double *ctr = dense_scratch(h, 7, (size_t)d, &rc); if (rc) return rc;
    landmark_center(Z, ldz, m, d, ctr, stream);
    augment_rows(h, Z, ldz, m, d, inv_ls, ctr, 1, Za, KA, stream);
    augment_rows(h, X, ldx, N, d, inv_ls, ctr, 0, Xa, KA, stream);
    if (packed_rp > 0) gemm_nt(h, (int)N, m, KA, 1.0, Xa, KA, Za, KA, 0.0, Kt, packed_rp, 0.0, kGemmPackedOut, nullptr, 0, stream, kind);
    else gemm_nt(h, (int)N, m, KA, 1.0, Xa, KA, Za, KA, 0.0, Kt, ldkt, 0.0, 0, nullptr, 0, stream, kind);
    return NK_OK;
}

}  // namespace nk

using namespace nk;

extern "C" {

int nk_lift(nk_handle *h, const double *Z, long long ldz, int m, int d, const double *inv_ls, int kind, const double *Sinv,
            long long ldsi, const double *X, long long ldx, long long N, double *Phi, long long ldphi, double *PhiT,
            long long ldphit, void *stream_) {
    if (!h) return NK_E_INVALID;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!Z || !inv_ls || !Sinv || !X || (!Phi && !PhiT) || m < 1 || d < 1 || N < 1 || N > 2000000000LL)
        return set_err(h, NK_E_INVALID, "nk_lift: bad argument");
    if (kind != NK_KERNEL_RBF && kind != NK_KERNEL_MATERN52) return set_err(h, NK_E_INVALID, "nk_lift: unsupported kernel kind");
    NK_ON_DEVICE(h);
    int rc;
    const int ldm = even_i(m);
    double *Kt = dense_scratch(h, 2, (size_t)N * ldm, &rc); if (rc) return rc;
    if ((rc = kernel_cross_t(h, Z, ldz, m, d, inv_ls, kind, X, ldx, N, Kt, ldm, stream, 0)) != NK_OK) return rc;
    // Phi^T (N,m) = K^T S^-1 ;  Phi (m,N) = S^-1 K  -- one product, stored both ways as requested
    const int flags = Phi ? kGemmStoreT : 0;
    gemm_nt(h, (int)N, m, m, 1.0, Kt, ldm, Sinv, ldsi, 0.0, PhiT, ldphit, 0.0, flags, Phi, ldphi, stream);
    NK_CUDA(h, cudaGetLastError());
    return NK_OK;
}

int nk_predict(nk_handle *h, const double *Z, long long ldz, int m, int d, int p, const double *inv_ls, int kind,
               const double *Sinv, long long ldsi, const double *W, long long ldw, const double *X_aug, long long ldx, long long N,
               double *Yhat, long long ldy, void *stream_) {
    if (!h) return NK_E_INVALID;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!Z || !inv_ls || !Sinv || !W || !X_aug || !Yhat || m < 1 || d < 1 || p < 0 || N < 1 || N > 2000000000LL)
        return set_err(h, NK_E_INVALID, "nk_predict: bad argument");
    if (kind != NK_KERNEL_RBF && kind != NK_KERNEL_MATERN52) return set_err(h, NK_E_INVALID, "nk_predict: unsupported kernel kind");
    NK_ON_DEVICE(h);
    int rc;
    const int ldm = even_i(m), ldf = even_i(m + p);
    double *Kt = dense_scratch(h, 2, (size_t)N * ldm, &rc); if (rc) return rc;
    double *F = dense_scratch(h, 3, (size_t)N * ldf, &rc); if (rc) return rc;    // [Phi^T | U] (N, m+p)
    if ((rc = kernel_cross_t(h, Z, ldz, m, d, inv_ls, kind, X_aug, ldx, N, Kt, ldm, stream, 0)) != NK_OK) return rc;
    gemm_nt(h, (int)N, m, m, 1.0, Kt, ldm, Sinv, ldsi, 0.0, F, ldf, 0.0, 0, nullptr, 0, stream);
    copy_cols(h, N, p, X_aug + d, ldx, F + m, ldf, stream);
    gemm_nt(h, (int)N, d, m + p, 1.0, F, ldf, W, ldw, 0.0, Yhat, ldy, 0.0, 0, nullptr, 0, stream);
    NK_CUDA(h, cudaGetLastError());
    return NK_OK;
}

int nk_rollout(nk_handle *h, int m, int p, int d, int T, long long nb, const double *A, const double *B, const double *C,
               const double *Z0, const double *U, double *Yhat, const double *Ytrue, double *sq_err, double *sq_sim,
               double *Zfinal, void *stream_) {
    if (!h) return NK_E_INVALID;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (m < 1 || p < 0 || d < 1 || T < 1 || nb < 1 || nb > 2000000000LL || !A || !C || !Z0 || (p && T > 1 && (!B || !U)))
        return set_err(h, NK_E_INVALID, "nk_rollout: bad argument");
    if (Ytrue && (!sq_err || !sq_sim)) return set_err(h, NK_E_INVALID, "nk_rollout: Ytrue needs sq_err and sq_sim");
    NK_ON_DEVICE(h);
    // One persistent packed-operand GEMM per time step (nk_pgemm.cu):
    //     [ Z_{i+1} | Yhat_i ] = [ Z_i | U_i ] * [ A  B ; C  0 ]^T
    // The lifted states stay in the packed operand layout from step to step (the epilogue writes Z_{i+1} with the result
    // column as contraction index), the stacked model [A B; C 0] is packed once, and U_i is dropped into the p spare
    // contraction columns of Z_i before each step.  The recurrence is serial in i exactly like the reference loop.
    int rc;
    const int K = m + p, KS = (K + kSlabK - 1) / kSlabK;
    const long long MPn = pad_to(nb, kTile);
    const int rp_z = (int)(MPn / kPanel);
    const int NWP = (int)pad_to(m + d, kTile), rp_w = NWP / kPanel;
    const int NCP = (int)pad_to(d, kTile), rp_c = NCP / kPanel;
    const size_t zdoubles = (size_t)KS * kSlabK * MPn;
    double *Zp[2];
    Zp[0] = dense_scratch(h, 0, zdoubles, &rc); if (rc) return rc;
    Zp[1] = dense_scratch(h, 1, zdoubles, &rc); if (rc) return rc;
    double *Ystep = nullptr;
    if (!Yhat) { Ystep = dense_scratch(h, 2, (size_t)nb * d, &rc); if (rc) return rc; }
    double *Wp = dense_scratch(h, 3, (size_t)KS * kSlabK * NWP, &rc); if (rc) return rc;
    double *Cpk = dense_scratch(h, 4, (size_t)KS * kSlabK * NCP, &rc); if (rc) return rc;
    NK_CUDA(h, cudaMemsetAsync(Zp[0], 0, zdoubles * 8, stream));
    NK_CUDA(h, cudaMemsetAsync(Zp[1], 0, zdoubles * 8, stream));
    NK_CUDA(h, cudaMemsetAsync(Wp, 0, (size_t)KS * kSlabK * NWP * 8, stream));
    NK_CUDA(h, cudaMemsetAsync(Cpk, 0, (size_t)KS * kSlabK * NCP * 8, stream));
    pack_rows(h, A, m, m, m, Wp, rp_w, 0, 0, stream);
    if (p) pack_rows(h, B, p, m, p, Wp, rp_w, 0, m, stream);
    pack_rows(h, C, m, d, m, Wp, rp_w, m, 0, stream);
    pack_rows(h, C, m, d, m, Cpk, rp_c, 0, 0, stream);
    pack_rows(h, Z0, m, nb, m, Zp[0], rp_z, 0, 0, stream);
    if (Ytrue) {
        NK_CUDA(h, cudaMemsetAsync(sq_err, 0, (size_t)nb * 8, stream));
        NK_CUDA(h, cudaMemsetAsync(sq_sim, 0, (size_t)nb * 8, stream));
    }
    const int warps = 8;
    int cur = 0;
    for (int i = 0; i < T; i++) {
        double *Yi = Yhat ? Yhat + (size_t)i * nb * d : Ystep;
        PGemmParams P;
        P.M = (int)nb; P.KS = KS; P.Ap = Zp[cur]; P.a_rp = rp_z; P.alpha = 1.0; P.beta = 0.0;
        P.C = Yi; P.ldc = d; P.tiles_m = (int)(MPn / kTile);
        if (i < T - 1) {
            if (p) pack_rows(h, U + (size_t)i * nb * p, p, nb, p, Zp[cur], rp_z, 0, m, stream);
            P.N = m + d; P.Bp = Wp; P.b_rp = rp_w; P.c_col0 = m;
            P.Cp = Zp[cur ^ 1]; P.c_rp = rp_z; P.cp_cols = m; P.tiles_n = NWP / kTile;
        } else {
            P.N = d; P.Bp = Cpk; P.b_rp = rp_c; P.c_col0 = 0;          // last step: only yhat_{T-1} = C z_{T-1}
            P.Cp = nullptr; P.c_rp = 0; P.cp_cols = 0; P.tiles_n = NCP / kTile;
        }
        launch_pgemm(h, P, stream);
        if (Ytrue) {
            step_error_kernel<<<(unsigned)((nb + warps - 1) / warps), warps * 32, 0, stream>>>(nb, d, Yi, Ytrue + (size_t)i * nb * d, sq_err, sq_sim);
            h->launches++;
        }
        if (i < T - 1) cur ^= 1;
    }
    if (Zfinal) unpack_rows(h, Zp[cur], rp_z, 0, 0, nb, m, Zfinal, m, stream);
    NK_CUDA(h, cudaGetLastError());
    return NK_OK;
}

int nk_closed_loop(nk_handle *h, int m, int p, int d, int steps, long long nb, const double *A, const double *B, const double *C,
                   const double *K, const double *Z0, const double *Zref, double *Xs, double *Us, double *Zfinal, void *stream_) {
    if (!h) return NK_E_INVALID;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (m < 1 || p < 1 || d < 1 || steps < 1 || nb < 1 || nb > 2000000000LL || !A || !B || !C || !K || !Z0 || !Zref || !Xs || !Us)
        return set_err(h, NK_E_INVALID, "nk_closed_loop: bad argument");
    NK_ON_DEVICE(h);
    // Per step, in the reference's order of operations (benchmark_lqr_cloth.py:80-84):
    //   u_i = K (phi_ref - z_i)          packed difference (elementwise), then a p-column packed GEMM
    //   [ z_{i+1} | x_i ] = [ z_i | u_i ] [ A B ; C 0 ]^T     the rollout step of nk_rollout
    int rc;
    const int Kc = m + p, KS = (Kc + kSlabK - 1) / kSlabK;
    const long long MPn = pad_to(nb, kTile);
    const int rp_z = (int)(MPn / kPanel);
    const int NWP = (int)pad_to(m + d, kTile), rp_w = NWP / kPanel;
    const int NKP = (int)pad_to(p, kTile), rp_k = NKP / kPanel;
    const size_t zdoubles = (size_t)KS * kSlabK * MPn;
    double *Zp[2];
    Zp[0] = dense_scratch(h, 0, zdoubles, &rc); if (rc) return rc;
    Zp[1] = dense_scratch(h, 1, zdoubles, &rc); if (rc) return rc;
    double *Rp = dense_scratch(h, 2, zdoubles, &rc); if (rc) return rc;      // packed phi_ref
    double *Wp = dense_scratch(h, 3, (size_t)KS * kSlabK * NWP, &rc); if (rc) return rc;
    double *Kp = dense_scratch(h, 4, (size_t)KS * kSlabK * NKP, &rc); if (rc) return rc;
    double *Dp = dense_scratch(h, 5, zdoubles, &rc); if (rc) return rc;      // packed phi_ref - z_i
    NK_CUDA(h, cudaMemsetAsync(Zp[0], 0, zdoubles * 8, stream));
    NK_CUDA(h, cudaMemsetAsync(Zp[1], 0, zdoubles * 8, stream));
    NK_CUDA(h, cudaMemsetAsync(Rp, 0, zdoubles * 8, stream));
    NK_CUDA(h, cudaMemsetAsync(Wp, 0, (size_t)KS * kSlabK * NWP * 8, stream));
    NK_CUDA(h, cudaMemsetAsync(Kp, 0, (size_t)KS * kSlabK * NKP * 8, stream));
    pack_rows(h, A, m, m, m, Wp, rp_w, 0, 0, stream);
    pack_rows(h, B, p, m, p, Wp, rp_w, 0, m, stream);
    pack_rows(h, C, m, d, m, Wp, rp_w, m, 0, stream);
    pack_rows(h, K, m, p, m, Kp, rp_k, 0, 0, stream);
    pack_rows(h, Z0, m, nb, m, Zp[0], rp_z, 0, 0, stream);
    pack_rows(h, Zref, m, nb, m, Rp, rp_z, 0, 0, stream);
    const long long sub_blocks = ((long long)zdoubles + 255) / 256;
    int cur = 0;
    for (int i = 0; i < steps; i++) {
        double *Ui = Us + (size_t)i * nb * p, *Xi = Xs + (size_t)i * nb * d;
        sub_kernel<<<(unsigned)(sub_blocks < 16384 ? sub_blocks : 16384), 256, 0, stream>>>((long long)zdoubles, Rp, Zp[cur], Dp);
        h->launches++;
        PGemmParams P;
        P.M = (int)nb; P.KS = KS; P.alpha = 1.0; P.beta = 0.0; P.tiles_m = (int)(MPn / kTile);
        P.Ap = Dp; P.a_rp = rp_z; P.Bp = Kp; P.b_rp = rp_k; P.N = p; P.C = Ui; P.ldc = p; P.c_col0 = 0;
        P.Cp = nullptr; P.c_rp = 0; P.cp_cols = 0; P.tiles_n = NKP / kTile;
        launch_pgemm(h, P, stream);                                            // u_i = (phi_ref - z_i) K^T
        pack_rows(h, Ui, p, nb, p, Zp[cur], rp_z, 0, m, stream);
        P.Ap = Zp[cur]; P.Bp = Wp; P.b_rp = rp_w; P.N = m + d; P.C = Xi; P.ldc = d; P.c_col0 = m;
        P.Cp = Zp[cur ^ 1]; P.c_rp = rp_z; P.cp_cols = m; P.tiles_n = NWP / kTile;
        launch_pgemm(h, P, stream);                                            // x_i = C z_i, z_{i+1} = A z_i + B u_i
        cur ^= 1;
    }
    if (Zfinal) unpack_rows(h, Zp[cur], rp_z, 0, 0, nb, m, Zfinal, m, stream);
    NK_CUDA(h, cudaGetLastError());
    return NK_OK;
}

}  // extern "C"
