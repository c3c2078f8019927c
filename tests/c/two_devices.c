/* A host WITHOUT Python shards one Nystrom-Koopman fit over two devices through the C ABI alone (include/nk_b200.h):
 *   per device: nk_gram_begin / nk_gram_update (its half of the samples) / nk_gram_finalize into one packed buffer,
 *   nk_allreduce_grams (the only data-path collective; the host owns the NCCL communicators),
 *   nk_solve_abc_part (each device solves half of the right-hand-side columns), peer copies, nk_solve_abc_finish.
 * Checked against the same fit done by device 0 alone.  Exit code 0 = ok.  Built and run by tests/test_c_two_devices.py. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <cuda_runtime.h>
#include <nccl.h>
#include "nk_b200.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 2; } } while (0)
#define NK(h, x) do { int r_ = (x); if (r_ != NK_OK) { fprintf(stderr, "%s -> %d: %s\n", #x, r_, nk_last_error_string(h)); return 3; } } while (0)
#define NC(x) do { ncclResult_t r_ = (x); if (r_ != ncclSuccess) { fprintf(stderr, "%s: %s\n", #x, ncclGetErrorString(r_)); return 4; } } while (0)

static double urand(unsigned long long *s) { *s = *s * 6364136223846793005ULL + 1442695040888963407ULL; return (double)(*s >> 11) / 9007199254740992.0; }
static double nrand(unsigned long long *s) { double u = urand(s) + 1e-300, v = urand(s); return sqrt(-2.0 * log(u)) * cos(6.283185307179586 * v); }

static double relerr(const double *a, const double *b, long long n) {
    double num = 0, den = 0;
    for (long long i = 0; i < n; i++) { num += (a[i] - b[i]) * (a[i] - b[i]); den += b[i] * b[i]; }
    return sqrt(num / (den > 0 ? den : 1));
}

enum { N = 6000, D = 24, P = 3, M = 200, N1 = M + P };

typedef struct { nk_grams g; double *flat; long long count; } packed_t;

static int alloc_packed(packed_t *pk) {
    const long long sz[7] = {(long long)M * M, (long long)M * M, (long long)M * M, (long long)M * P, (long long)M * P, (long long)P * P, (long long)D * M};
    pk->count = 0;
    for (int i = 0; i < 7; i++) pk->count += sz[i];
    CK(cudaMalloc((void **)&pk->flat, pk->count * 8));
    double *q = pk->flat;
    pk->g.Gxx = q; pk->g.ld_gxx = M; q += sz[0];
    pk->g.Gyx = q; pk->g.ld_gyx = M; q += sz[1];
    pk->g.Gyy = q; pk->g.ld_gyy = M; q += sz[2];
    pk->g.Gxu = q; pk->g.ld_gxu = P; q += sz[3];
    pk->g.Gyu = q; pk->g.ld_gyu = P; q += sz[4];
    pk->g.Guu = q; pk->g.ld_guu = P; q += sz[5];
    pk->g.GYy = q; pk->g.ld_gYy = M;
    return 0;
}

static int grams(nk_handle *h, const double *Z, const double *il, const double *X, const double *Y, long long n, packed_t *pk, cudaStream_t st) {
    NK(h, nk_gram_begin(h, Z, D, M, D, P, il, NK_KERNEL_RBF, 0, st));
    NK(h, nk_gram_update(h, X, D + P, Y, D, n, st));
    NK(h, nk_gram_finalize(h, (double *)pk->g.Gxx, M, (double *)pk->g.Gyx, M, (double *)pk->g.Gyy, M, (double *)pk->g.Gxu, P, (double *)pk->g.Gyu, P,
                           (double *)pk->g.Guu, P, (double *)pk->g.GYy, M, 0, st));
    return 0;
}

int main(void) {
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    if (ndev < 2) { printf("SKIP: needs 2 GPUs, found %d\n", ndev); return 77; }
    /* ---- data on the host: x ~ N(0,1), y = tanh(0.3 x) + small control term; landmarks = the first M next-states ---- */
    unsigned long long seed = 12345;
    double *X = (double *)malloc(sizeof(double) * N * (D + P)), *Y = (double *)malloc(sizeof(double) * N * D);
    for (long long i = 0; i < (long long)N * (D + P); i++) X[i] = nrand(&seed);
    for (int s = 0; s < N; s++) for (int k = 0; k < D; k++) Y[s * D + k] = tanh(0.3 * X[s * (D + P) + k] + 0.2 * X[s * (D + P) + (k + 1) % D]) + 0.1 * X[s * (D + P) + D + k % P];
    double il_h[D];
    for (int k = 0; k < D; k++) il_h[k] = 1.0 / 4.0;
    const double gamma_n = 1e-2 * N, jitter = 1e-6;
    const long long half = N / 2 + 7;                         /* uneven shards on purpose */
    const long long off[2] = {0, half}, cnt[2] = {half, N - half};

    int devs[2] = {0, 1};
    ncclComm_t comms[2];
    NC(ncclCommInitAll(comms, 2, devs));
    nk_handle *h[2];
    cudaStream_t st[2];
    double *dX[2], *dY[2], *dZ[2], *dil[2], *Kzz[2], *Kmm[2], *S[2], *Si[2], *GT[2], *CT[2];
    packed_t pk[2], full;
    for (int r = 0; r < 2; r++) {
        CK(cudaSetDevice(r));
        NK(NULL, nk_create(&h[r], r));
        CK(cudaStreamCreate(&st[r]));
        CK(cudaMalloc((void **)&dX[r], sizeof(double) * N * (D + P))); CK(cudaMalloc((void **)&dY[r], sizeof(double) * N * D));
        CK(cudaMalloc((void **)&dZ[r], sizeof(double) * M * D)); CK(cudaMalloc((void **)&dil[r], sizeof(double) * D));
        CK(cudaMemcpy(dX[r], X, sizeof(double) * N * (D + P), cudaMemcpyHostToDevice)); CK(cudaMemcpy(dY[r], Y, sizeof(double) * N * D, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dZ[r], Y, sizeof(double) * M * D, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dil[r], il_h, sizeof(double) * D, cudaMemcpyHostToDevice));
        for (double ***q = (double **[]){&Kzz[r], &Kmm[r], &S[r], &Si[r], NULL}; *q; q++) CK(cudaMalloc((void **)*q, sizeof(double) * M * M));
        CK(cudaMalloc((void **)&GT[r], sizeof(double) * N1 * M)); CK(cudaMalloc((void **)&CT[r], sizeof(double) * M * D));
        if (alloc_packed(&pk[r])) return 2;
    }
    /* ---- sharded Gram pass + the one collective ---- */
    for (int r = 0; r < 2; r++) {
        CK(cudaSetDevice(r));
        if (grams(h[r], dZ[r], dil[r], dX[r] + off[r] * (D + P), dY[r] + off[r] * D, cnt[r], &pk[r], st[r])) return 3;
    }
    NC(ncclGroupStart());
    for (int r = 0; r < 2; r++) { CK(cudaSetDevice(r)); NK(h[r], nk_allreduce_grams(h[r], comms[r], pk[r].flat, pk[r].count, st[r])); }
    NC(ncclGroupEnd());
    /* ---- reference: device 0 alone over all samples ---- */
    CK(cudaSetDevice(0));
    if (alloc_packed(&full)) return 2;
    if (grams(h[0], dZ[0], dil[0], dX[0], dY[0], N, &full, st[0])) return 3;
    for (int r = 0; r < 2; r++) { CK(cudaSetDevice(r)); CK(cudaStreamSynchronize(st[r])); }
    double *a = (double *)malloc(8 * full.count), *b = (double *)malloc(8 * full.count), *c = (double *)malloc(8 * full.count);
    CK(cudaMemcpy(a, pk[0].flat, 8 * full.count, cudaMemcpyDeviceToHost));
    CK(cudaSetDevice(1)); CK(cudaMemcpy(b, pk[1].flat, 8 * full.count, cudaMemcpyDeviceToHost));
    CK(cudaSetDevice(0)); CK(cudaMemcpy(c, full.flat, 8 * full.count, cudaMemcpyDeviceToHost));
    const double e_gram = relerr(a, c, full.count);
    const int same = memcmp(a, b, 8 * full.count) == 0;
    printf("allreduced Grams vs one-device Grams: %.3e; both devices bit-identical: %d\n", e_gram, same);
    if (!(e_gram <= 1e-13) || !same) return 10;
    /* ---- landmark stage on every device (small here), column-sharded solve ---- */
    const int g_half = (N1 + 1) / 2, c_half = (M + 1) / 2;
    for (int r = 0; r < 2; r++) {
        CK(cudaSetDevice(r));
        NK(h[r], nk_kzz(h[r], dZ[r], D, M, D, dil[r], NK_KERNEL_RBF, Kzz[r], M, st[r]));
        CK(cudaMemcpyAsync(Kmm[r], Kzz[r], sizeof(double) * M * M, cudaMemcpyDeviceToDevice, st[r]));
        double *diag = (double *)malloc(8 * M * M);
        CK(cudaStreamSynchronize(st[r]));
        CK(cudaMemcpy(diag, Kmm[r], 8 * M * M, cudaMemcpyDeviceToHost));
        for (int i = 0; i < M; i++) diag[i * M + i] += jitter;
        CK(cudaMemcpy(Kmm[r], diag, 8 * M * M, cudaMemcpyHostToDevice));
        free(diag);
        NK(h[r], nk_sym_sqrt(h[r], M, Kmm[r], M, jitter, S[r], M, Si[r], M, NULL, st[r]));
        nk_landmarks lm = {Kzz[r], M, S[r], M, Si[r], M, NULL, 0, NULL, 0};
        const int g0 = r * g_half, gc = r ? N1 - g_half : g_half, c0 = r * c_half, cc = r ? M - c_half : c_half;
        int info = 0;
        NK(h[r], nk_solve_abc_part(h[r], M, P, D, gamma_n, jitter, &pk[r].g, &lm, g0, gc, GT[r] + (long long)g0 * M, M, c0, cc, CT[r] + (long long)c0 * D, D, &info, st[r]));
    }
    /* device 1's rows -> device 0 (what an all-gather does) */
    CK(cudaMemcpyPeer(GT[0] + (long long)g_half * M, 0, GT[1] + (long long)g_half * M, 1, sizeof(double) * (N1 - g_half) * M));
    CK(cudaMemcpyPeer(CT[0] + (long long)c_half * D, 0, CT[1] + (long long)c_half * D, 1, sizeof(double) * (M - c_half) * D));
    CK(cudaSetDevice(0));
    double *A, *B, *C, *W, *A1, *B1, *C1, *W1;
    for (double ***q = (double **[]){&A, &A1, NULL}; *q; q++) CK(cudaMalloc((void **)*q, 8 * M * M));
    for (double ***q = (double **[]){&B, &B1, NULL}; *q; q++) CK(cudaMalloc((void **)*q, 8 * M * P));
    for (double ***q = (double **[]){&C, &C1, NULL}; *q; q++) CK(cudaMalloc((void **)*q, 8 * D * M));
    for (double ***q = (double **[]){&W, &W1, NULL}; *q; q++) CK(cudaMalloc((void **)*q, 8 * D * N1));
    NK(h[0], nk_solve_abc_finish(h[0], M, P, D, GT[0], M, CT[0], D, A, M, B, P, C, M, W, N1, st[0]));
    nk_landmarks lm0 = {Kzz[0], M, S[0], M, Si[0], M, NULL, 0, NULL, 0};
    int info = 0;
    NK(h[0], nk_solve_abc(h[0], M, P, D, gamma_n, jitter, &full.g, &lm0, A1, M, B1, P, C1, M, W1, N1, &info, st[0]));
    CK(cudaStreamSynchronize(st[0]));
    double worst = 0;
    struct { double *x, *y; long long n; const char *nm; } cmp[4] = {{A, A1, (long long)M * M, "A"}, {B, B1, (long long)M * P, "B"}, {C, C1, (long long)D * M, "C"}, {W, W1, (long long)D * N1, "W"}};
    for (int i = 0; i < 4; i++) {
        double *hx = (double *)malloc(8 * cmp[i].n), *hy = (double *)malloc(8 * cmp[i].n);
        CK(cudaMemcpy(hx, cmp[i].x, 8 * cmp[i].n, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(hy, cmp[i].y, 8 * cmp[i].n, cudaMemcpyDeviceToHost));
        const double e = relerr(hx, hy, cmp[i].n);
        printf("two-device sharded %s vs one-device %s: %.3e\n", cmp[i].nm, cmp[i].nm, e);
        if (e > worst) worst = e;
        free(hx); free(hy);
    }
    for (int r = 0; r < 2; r++) { CK(cudaSetDevice(r)); nk_destroy(h[r]); ncclCommDestroy(comms[r]); }
    if (!(worst <= 1e-9)) return 11;
    printf("OK\n");
    return 0;
}
