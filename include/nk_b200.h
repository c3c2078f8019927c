/* nk_b200.h -- C ABI of the B200-native Nystrom-Koopman hot path (libnkb200.so).
 *
 * Drop-in boundary: these entry points are what a binding of the reference estimator
 * (LCSL/nys-koop-lqr, regressors.py::KoopmanNystromRegressor) calls instead of numpy/scipy/sklearn.
 * Every function cites the reference lines it replaces.  Conventions:
 *   - return value: 0 = ok, <0 = error (NK_E_*); nk_last_error_string() gives the text.  No exceptions.
 *   - all matrix pointers are CALLER-OWNED DEVICE pointers to IEEE float64, row-major with an explicit
 *     leading dimension (elements) unless a parameter is documented as "host".
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*); asynchronous failures
 *     surface at the next synchronising call.  A handle is bound to one device and is not thread-safe; every call runs
 *     on the handle's device and restores the caller's current CUDA device before it returns (streams and pointers passed
 *     in must belong to the handle's device).
 *   - there is NO CPU fallback: every function fails with NK_E_CUDA when no sm_100 device is usable.
 */
#ifndef NK_B200_H
#define NK_B200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct nk_handle nk_handle;

enum { NK_OK = 0, NK_E_INVALID = -1, NK_E_CUDA = -2, NK_E_NOT_SPD = -3, NK_E_NOMEM = -4, NK_E_STATE = -5 };
enum { NK_KERNEL_RBF = 0, NK_KERNEL_MATERN52 = 1 };   /* sklearn RBF / Matern(nu=2.5): regressors.py:15-26 */

int nk_version(void);
int nk_create(nk_handle **out, int device);
int nk_destroy(nk_handle *h);
const char *nk_last_error_string(nk_handle *h);   /* h may be NULL: last creation error */
int nk_device_sm_count(nk_handle *h);
/* Workspaces are grow-only per handle (a CV sweep at m=8192 leaves tens of GB behind): this frees all of them (synchronises
 * the device; the next call reallocates what it needs).  A Gram accumulation in progress is discarded (nk_gram_update /
 * nk_gram_finalize then fail with NK_E_STATE until the next nk_gram_begin). */
int nk_release_scratch(nk_handle *h);

/* ---- fused kernel lift + data-sample Grams  (regressors.py:141-142 lift; :147,151,153,162,164 Grams) ----
 * Streaming form: begin(landmarks, kernel) -> update(sample block)* -> finalize(outputs).
 *   Z (m,d) landmarks; inv_ls (d) = 1/length_scale per state dimension (device); kind = NK_KERNEL_*;
 *   chunk = samples per on-chip feature chunk (0 = default 512; rounded to a multiple of 128).
 * update: X (n,d+p) rows [x_t | u_t] (controls are the LAST p columns, regressors.py:123-126), Y (n,d) = x_{t+1}.
 * finalize writes (or adds to, if accumulate != 0):
 *   Gxx = Phi_x Phi_x^T (m,m)   Gyx = Phi_y Phi_x^T (m,m)   Gyy = Phi_y Phi_y^T (m,m)
 *   Gxu = Phi_x U (m,p)         Gyu = Phi_y U (m,p)         Guu = U^T U (p,p)      GYy = Y^T Phi_y^T (d,m)
 * with Phi_x = k(Z, X_state) and Phi_y = k(Z, Y), both (m,n) and never materialised.  Any output may be NULL. */
int nk_gram_begin(nk_handle *h, const double *Z, long long ldz, int m, int d, int p,
                  const double *inv_ls, int kind, int chunk, void *stream);
/* Same with DISTINCT input landmarks (regressors.py:133-134: a caller may inject nystrom_centers_input != _output; both (m,d)):
 * Phi_x = k(Z_in, X_state), Phi_y = k(Z_out, Y).  nk_gram_begin is this call with Z_in == Z_out. */
int nk_gram_begin_io(nk_handle *h, const double *Z_in, long long ldz_in, const double *Z_out, long long ldz_out, int m, int d, int p,
                     const double *inv_ls, int kind, int chunk, void *stream);
int nk_gram_update(nk_handle *h, const double *X, long long ldx, const double *Y, long long ldy,
                   long long n, void *stream);
int nk_gram_finalize(nk_handle *h, double *Gxx, long long ld_gxx, double *Gyx, long long ld_gyx,
                     double *Gyy, long long ld_gyy, double *Gxu, long long ld_gxu, double *Gyu, long long ld_gyu,
                     double *Guu, long long ld_guu, double *GYy, long long ld_gYy, int accumulate, void *stream);
/* The fused kernel has no grid-wide barrier: items wait on counters.  Every such wait carries a watchdog (20 s); if it fires the
 * kernel drains and the accumulation is abandoned.  nk_gram_finalize latches the flag (stream-ordered, no wait); it is reported as
 * NK_E_STATE by the next SYNCHRONISING call (nk_solve_abc / nk_solve_abc_part / nk_cv_weights) or by nk_gram_status, which
 * synchronises `stream` and returns NK_OK / NK_E_STATE. */
int nk_gram_status(nk_handle *h, void *stream);
/* ---- the one data-path collective of the sample-sharded fit: float64 sum of the packed Grams over the devices.
 * `packed` is this device's buffer of `count` doubles (any layout, typically [Gxx|Gyx|Gyy|Gxu|Gyu|Guu|GYy] as one allocation),
 * reduced IN PLACE on `stream` with ncclAllReduce(..., ncclDouble, ncclSum, comm, stream).  `nccl_comm` is the caller's
 * ncclComm_t for this device (the host owns communicator set-up: ncclCommInitAll / ncclCommInitRank, or torch.distributed's).
 * The library does not link NCCL: the symbol is taken from the NCCL already loaded in the process (dlsym), NK_E_STATE if none is.
 * In a single process driving several devices wrap the per-device calls in ncclGroupStart / ncclGroupEnd as usual.
 * With this entry point a host without Python shards a fit: per device nk_gram_begin / update / finalize on its sample block,
 * nk_allreduce_grams, then nk_solve_abc (or nk_solve_abc_part + an all-gather) -- tests/test_c_two_devices.py does exactly that. */
int nk_allreduce_grams(nk_handle *h, void *nccl_comm, double *packed, long long count, void *stream);

/* Introspection, HOST ONLY (no device needed): the work plan nk_gram_begin builds for these sizes on a GPU with sm_count SMs.
 * summary[12] = {chunk, MP, KLS, EP, psi_rows, nblk, ntiles, n_pack, n_lift, n_gram, period_len, nslots};  items (may be NULL)
 * receives up to items_cap entries of ONE period of the global work order, 4 ints each {type (0 pack, 1 lift, 2 Gram tile), a, b, c}
 * (pack: a = 128-sample strip; lift: a = side, b = landmark block, c = strip; Gram: a, b = row blocks of Psi, c = accumulator tile).
 * Period k holds the Gram items of chunk k and the pack / lift items of chunk k+1.  Returns period_len (<0: invalid argument).
 * tests/test_gram_plan.py checks on the CPU that every item depends only on items earlier in the claim order. */
int nk_gram_plan(int m, int d, int p, int chunk, int sm_count, int *summary, int *items, int items_cap);
/* executed FP64 flops of the last update (for roofline accounting), and launches issued so far */
double nk_gram_last_executed_flops(nk_handle *h);
long long nk_launch_count(nk_handle *h);

/* ---- measurement aid: register-only DMMA.8x8x4 issue-rate probe (the FP64 tensor roofline denominator, measured on
 * the device the handle is bound to).  Runs ~ms_target milliseconds, synchronises, writes TFLOP/s to *tflops (host). ---- */
int nk_probe_dmma_tflops(nk_handle *h, double ms_target, double *tflops);

/* ---- landmark kernel matrix K_zz = k(Z,Z)  (regressors.py:139,143,144,174); diagonal is exactly 1 ---- */
int nk_kzz(nk_handle *h, const double *Z, long long ldz, int m, int d, const double *inv_ls, int kind,
           double *Kzz, long long ldk, void *stream);

/* ---- the kernel function itself, elementwise: out[i] = k(exponent[i]) where exponent = -r^2/2 of the length-scaled distance r
 * (what the lift GEMMs accumulate): RBF exp(-r^2/2) (sklearn kernels.py RBF.__call__), Matern-5/2 (1+a+a^2/3)exp(-a), a=sqrt(5) r
 * (Matern.__call__, nu=2.5).  Exposed so that the device implementation can be checked against the library functions. ---- */
int nk_kernel_function(nk_handle *h, int kind, long long count, const double *exponent, double *out, void *stream);

/* ---- kernel cross matrix K = k(Z, X): (m, N) row-major from X (N, d) rows  (regressors.py:176) ---- */
int nk_kernel_cross(nk_handle *h, const double *Z, long long ldz, int m, int d, const double *inv_ls, int kind,
                    const double *X, long long ldx, long long N, double *K, long long ldk, void *stream);

/* ---- dense building blocks (all DMMA): C = alpha*op(A)*op(B) + beta*C, row-major ----
 * transa/transb: 0 = as stored, 1 = transposed.  A is (M,K) after op, B is (K,N) after op. */
int nk_gemm(nk_handle *h, int transa, int transb, int M, int N, int K, double alpha, const double *A, long long lda,
            const double *B, long long ldb, double beta, double *C, long long ldc, void *stream);
/* Cholesky A = L L^T in place (lower triangle; the strict upper triangle is zeroed). info (host int*) receives 0 or
 * the 1-based index of the first non-positive pivot; the call synchronises the stream to read it. */
int nk_potrf(nk_handle *h, int n, double *A, long long lda, int *info, void *stream);
/* B <- L^-1 B (trans=0) or L^-T B (trans=1), L lower (n,n), B (n,nrhs) */
int nk_trsm_lower(nk_handle *h, int trans, int n, int nrhs, const double *L, long long ldl, double *B, long long ldb, void *stream);
/* symmetric principal square root S = K^(1/2) and S^-1 of an SPD matrix (scipy.linalg.sqrtm at regressors.py:140,163,175
 * and the solves against it at :152,153,177).  lambda_min_bound > 0: a lower bound on the smallest eigenvalue (the
 * jitter 1e-6 for K_mm).  iters (host int*, may be NULL) receives the Newton-Schulz iteration count.
 * n <= 128 runs as one cooperative launch; from n = 1024 the iteration schedule starts from an inverse-iteration ESTIMATE of
 * lambda_min (never below lambda_min_bound) and a residual check adds steps if the estimate was too optimistic.  Synchronises `stream`. */
int nk_sym_sqrt(nk_handle *h, int n, const double *K, long long ldk, double lambda_min_bound,
                double *S, long long lds, double *Sinv, long long ldsi, int *iters, void *stream);

/* ---- Grams -> Koopman matrices  (regressors.py:147-169) ----
 * The seven data-sample Grams (outputs of nk_gram_finalize, possibly summed over devices) and the landmark-only matrices travel
 * as two plain structs of caller-owned device pointers, each row-major with its own leading dimension (elements). */
typedef struct nk_grams {
    const double *Gxx; long long ld_gxx;   /* (m,m) Phi_x Phi_x^T */
    const double *Gyx; long long ld_gyx;   /* (m,m) Phi_y Phi_x^T */
    const double *Gyy; long long ld_gyy;   /* (m,m) Phi_y Phi_y^T */
    const double *Gxu; long long ld_gxu;   /* (m,p) Phi_x U   (NULL when p == 0) */
    const double *Gyu; long long ld_gyu;   /* (m,p) Phi_y U */
    const double *Guu; long long ld_guu;   /* (p,p) U^T U */
    const double *GYy; long long ld_gYy;   /* (d,m) Y^T Phi_y^T */
} nk_grams;
typedef struct nk_landmarks {
    const double *Kzz; long long ld_kzz;       /* (m,m) k(Z_out, Z_out), no jitter (regressors.py:144) */
    const double *S; long long ld_s;           /* (m,m) (Kzz + jitter I)^(1/2), symmetric (regressors.py:140) */
    const double *Sinv; long long ld_sinv;     /* (m,m) its inverse */
    const double *Kzz_in; long long ld_kzz_in; /* (m,m) k(Z_in, Z_in) -- NULL when the input landmarks ARE the output landmarks */
    const double *Kio; long long ld_kio;       /* (m,m) k(Z_in, Z_out) (regressors.py:144) -- NULL together with Kzz_in */
} nk_landmarks;
/* inner = [[Gxx + gn*(Kzz_in + jitter I), Gxu],[Gxu^T, Guu + gn*I]],  G = S^-1 [Gyx|Gyu] inner^-1 blkdiag(Kio S^-1, I_p)
 * A = G[:, :m], B = G[:, m:];  C = GYy (gn*(Kzz + jitter I) + Gyy)^-1 S;  W = C G.   (Kzz_in = Kio = Kzz when NULL; then
 * Kzz S^-1 = S - jitter S^-1 is formed elementwise.)  Cholesky takes the place of scipy.linalg.lstsq (same solve while nothing
 * is truncated); info (host int*, may be NULL): 0, or 1 / 2 if inner_term / inner_term_rec is not positive definite.
 * Outputs: A (m,m), B (m,p), C (d,m), W (d,m+p), each with a leading dimension.  ONE stream synchronisation (the verdicts).
 *
 * nk_solve_abc_part computes a slice of the same solve, for sharding it over devices (each right-hand-side column is
 * independent): rows [g_row0, g_row0+g_rows) of G^T ((m+p) x m: row c = column c of [A|B]) and rows [c_row0, c_row0+c_rows) of
 * C^T (m x d).  Either range may be empty.  Every device factors both systems, solves only its columns; the slices are
 * exchanged by the caller (one all-gather each) and nk_solve_abc_finish turns the assembled G^T, C^T into A, B, C, W = C G
 * (any output may be NULL; W needs C).  nk_solve_abc is part(all rows) + finish on one device. */
int nk_solve_abc(nk_handle *h, int m, int p, int d, double gamma_n, double jitter, const nk_grams *G, const nk_landmarks *L,
                 double *A, long long lda, double *B, long long ldb, double *C, long long ldc, double *W, long long ldw, int *info,
                 void *stream);
int nk_solve_abc_part(nk_handle *h, int m, int p, int d, double gamma_n, double jitter, const nk_grams *G, const nk_landmarks *L,
                      int g_row0, int g_rows, double *GT, long long ld_gt, int c_row0, int c_rows, double *CT, long long ld_ct,
                      int *info, void *stream);
int nk_solve_abc_finish(nk_handle *h, int m, int p, int d, const double *GT, long long ld_gt, const double *CT, long long ld_ct,
                        double *A, long long lda, double *B, long long ldb, double *C, long long ldc, double *W, long long ldw, void *stream);

/* ---- lift  phi = S^-1 k(Z, X)  (regressors.py:171-178): X (N,d) rows -> Phi (m,N); PhiT (N,m) is the same data
 * transposed (either output may be NULL).  N is limited by scratch memory (N*(m+d+2) doubles); callers chunk. ---- */
int nk_lift(nk_handle *h, const double *Z, long long ldz, int m, int d, const double *inv_ls, int kind,
            const double *Sinv, long long ldsi, const double *X, long long ldx, long long N,
            double *Phi, long long ldphi, double *PhiT, long long ldphit, void *stream);

/* ---- predict (regressors.py:48-55): Yhat (N,d) = (W [phi(X_state); U])^T for X_aug (N, d+p) rows, W (d, m+p) ---- */
int nk_predict(nk_handle *h, const double *Z, long long ldz, int m, int d, int p, const double *inv_ls, int kind,
               const double *Sinv, long long ldsi, const double *W, long long ldw, const double *X_aug, long long ldx,
               long long N, double *Yhat, long long ldy, void *stream);

/* ---- batched open-loop rollout (benchmark_lqr_cloth.py:18-36 and its two copies in _classic.py:23-41, _hjb.py:23-44) ----
 * Trajectory-major storage (one row per trajectory):  for each of nb trajectories
 *     yhat_0 = C z_0;  z_{i+1} = A z_i + B u_i,  yhat_{i+1} = C z_{i+1},  i = 0..T-2   (serial in i, as the reference).
 *   Z0 (nb, m) lifted initial states; U (T-1, nb, p) controls, step-major; Yhat (T, nb, d) output, may be NULL;
 *   Ytrue (T, nb, d) optional: if given, sq_err (nb) = sum_{i,j} (Ytrue-Yhat)^2 and sq_sim (nb) = sum Yhat^2
 *   (the scripts' two RMSE definitions, _cloth.py:34 and _classic.py:39, are formed from these on the host).
 *   Zfinal (nb, m) optional: lifted state after the last step. */
int nk_rollout(nk_handle *h, int m, int p, int d, int T, long long nb, const double *A, const double *B, const double *C,
               const double *Z0, const double *U, double *Yhat, const double *Ytrue, double *sq_err, double *sq_sim,
               double *Zfinal, void *stream);

/* ---- batched closed loop in the lifted space (benchmark_lqr_cloth.py:69-104 `lqr_control`, lines 80-84 for one trajectory) ----
 *     u_i = K (phi_ref - z_i);   x_i = C z_i;   z_{i+1} = A z_i + B u_i,    i = 0..steps-1   (serial in i, same operation order)
 *   Z0 (nb, m) lifted initial states, Zref (nb, m) lifted references, K (p, m) gain (control.dlqr, computed on the host).
 *   Xs (steps, nb, d) visited states, Us (steps, nb, p) control increments, Zfinal (nb, m) optional. */
int nk_closed_loop(nk_handle *h, int m, int p, int d, int steps, long long nb, const double *A, const double *B, const double *C,
                   const double *K, const double *Z0, const double *Zref, double *Xs, double *Us, double *Zfinal, void *stream);

/* ---- cross-validation sweep over the (kernel, gamma) grid: learn_hyperparams in benchmark_lqr_hjb.py:47-71,
 * benchmark_lqr_classic.py:44-64, benchmark_lqr_cloth.py:39-66 (sklearn GridSearchCV cloning the estimator per candidate and
 * fold, `fit` on the training fold, score = RMSE of `predict` (regressors.py:48-55) on the held-out fold) ----
 * nk_cv_weights: for ONE kernel and ONE training fold, all nlam regularisation values at once.  The Grams are the fold's
 *   (nk_gram_* over the training samples); gamma_n (HOST array, nlam) = gamma * n_train (regressors.py:127).  The nlam
 *   pairs of regularised systems (regressors.py:151 inner_term, :162 inner_term_rec) are factored as one batch (blocked
 *   Cholesky batched over the regularisation grid) and solved for the d rows scoring needs:
 *     Wk[b] (d, m+p) = [ V_phi Kzz Kmm^-1 | V_u ],  V = GYy (gn_b Kmm + Gyy)^-1 [Gyx|Gyu] inner_b^-1,
 *   so that regressors.py:48-55 reads  Yhat = Wk[b] [k(Z,x); u]   (= weights [S^-1 k(Z,x); u]; the S factors cancel).
 *   Wk: device (nlam, d, ld_wk >= m+p).  info: HOST ints (nlam), 0 ok / 1 inner_term / 2 inner_term_rec / 3 K_mm not SPD.
 *   The call synchronises the stream.
 * nk_cv_score: sse[r] += sum_s (Yhat[s, r] - Y[s, r % d])^2 for the R = nlam*d stacked weight rows Wk (R, m+p) over
 *   the N held-out samples X_aug (N, d+p), Y (N, d).  sse (R) device, caller-zeroed (accumulates across calls).
 *   sklearn's 'neg_root_mean_squared_error' for value b is  -mean_j sqrt(sse[b*d + j] / N).
 * nk_axpy: y += alpha x (count doubles) -- combines per-fold Grams into training-fold Grams on the device. */
int nk_cv_weights(nk_handle *h, int m, int p, int d, int nlam, const double *gamma_n, double jitter, const nk_grams *G,
                  const double *Kzz, long long ld_kzz, double *Wk, long long ld_wk, int *info, void *stream);
int nk_cv_score(nk_handle *h, const double *Z, long long ldz, int m, int d, int p, const double *inv_ls, int kind,
                const double *Wk, int R, const double *X_aug, long long ldx, const double *Y, long long ldy, long long N,
                double *sse, void *stream);
int nk_axpy(nk_handle *h, long long count, double alpha, const double *x, double *y, void *stream);

#ifdef __cplusplus
}
#endif
#endif
