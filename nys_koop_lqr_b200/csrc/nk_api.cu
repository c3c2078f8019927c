// nk_api.cu -- extern "C" entry points of libnkb200.so: handle life-cycle and the fused lift+Gram path.
// (dense stage: nk_dense.cu; lift/rollout: nk_rollout.cu)
#include <cstdio>
#include <cstring>
#include <algorithm>
#include <dlfcn.h>
#include "nk_handle.cuh"

namespace nk {

static std::string g_create_err;

int set_err(nk_handle *h, int code, const std::string &msg) {
    if (h) h->err = msg; else g_create_err = msg;
    return code;
}
int check_cuda(nk_handle *h, cudaError_t e, const char *what) {
    if (e == cudaSuccess) return NK_OK;
    return set_err(h, NK_E_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}
int gram_watchdog_verdict(nk_handle *h) {
    if (h->gram_err_host && *h->gram_err_host != 0) {
        *h->gram_err_host = 0;
        h->gram_open = false;
        return set_err(h, NK_E_STATE, "fused lift+Gram kernel: a dependence wait timed out (work-plan bug or a lost CTA); the accumulation was abandoned");
    }
    return NK_OK;
}
int ensure(nk_handle *h, nk_devbuf &b, size_t bytes) {
    if (bytes == 0) bytes = 8;
    if (b.bytes >= bytes) return NK_OK;
    if (b.ptr) { cudaFree(b.ptr); b.ptr = nullptr; b.bytes = 0; }
    cudaError_t e = cudaMalloc(&b.ptr, bytes);
    if (e != cudaSuccess) { cudaGetLastError(); return set_err(h, NK_E_NOMEM, std::string("cudaMalloc failed: ") + cudaGetErrorString(e)); }
    b.bytes = bytes;
    return NK_OK;
}

void launch_pack_landmarks(const double *Z, long long ldz, int m, int d, int MP, int KLS, const double *inv_ls,
                           const double *center, double *ZP, cudaStream_t stream);
void launch_unpack(const double *Gws, const int *tile_of, int nblk, int row0, int col0, int rows, int cols,
                   double *out, long long ld, int accumulate, cudaStream_t stream);

void landmark_center(const double *Z, long long ldz, int m, int d, double *center, cudaStream_t stream);   // nk_dense.cu

// register-only DMMA issue-rate probe: 32 independent accumulator tiles per warp, 8 warps per CTA, 2 CTAs per SM
__global__ void __launch_bounds__(256) dmma_probe_kernel(double *out, int iters, double a0, double b0) {
    double c[32][2];
#pragma unroll
    for (int i = 0; i < 32; i++) { c[i][0] = 0.0; c[i][1] = 0.0; }
    const double a = a0 + threadIdx.x * 1e-9, b = b0;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 32; i++) dmma884(c[i][0], c[i][1], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 32; i++) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}


// ---- the work plan of the fused engine (pure host code: no device needed; also exported as nk_gram_plan for the CPU tests) ----
struct GramPlan {
    int chunk, MP, KLS, EP, psi_rows, nblk, ntiles, n_sy, n_pk, n_lf, nslots;
    std::vector<GramItem> period;     // one period of the global work order
    std::vector<int> tile_of;         // (nblk x nblk) block pair -> accumulator tile, -1 where nothing is accumulated
};

void plan_gram_chunk(int m, int d, int p, int chunk, int sm_count, GramPlan &P);

// chunk <= 0: the library picks the samples per feature chunk.  512 to begin with; when the problem is large enough for two
// chunks in flight to fill the GPU (nslots == 2), the chunk is doubled (up to 2048) as long as ONE chunk of packed features
// (psi_rows x chunk doubles) stays within 72 MB, i.e. comfortably inside the 126 MB L2 next to the streaming accumulator
// traffic: every chunk costs one read-modify-write of the whole accumulator workspace in HBM, so the traffic per sample halves
// with each doubling (m = 4096: 1024 samples per chunk, measured +0.4% and 0.65 instead of 1.3 MB of DRAM traffic per sample;
// m = 8192 stays at 512).
void plan_gram(int m, int d, int p, int chunk, int sm_count, GramPlan &P) {
    if (chunk > 0) { plan_gram_chunk(m, d, p, chunk, sm_count, P); return; }
    plan_gram_chunk(m, d, p, 512, sm_count, P);
    while (P.nslots == 2 && P.chunk < 2048 && (size_t)P.psi_rows * (size_t)(2 * P.chunk) * 8 <= (size_t)72 << 20) {
        GramPlan Q;
        plan_gram_chunk(m, d, p, 2 * P.chunk, sm_count, Q);
        if (Q.nslots != 2) break;
        P = Q;
    }
}

void plan_gram_chunk(int m, int d, int p, int chunk, int sm_count, GramPlan &P) {
    chunk = ((chunk + kTile - 1) / kTile) * kTile;
    P.chunk = chunk;
    P.MP = ((m + kTile - 1) / kTile) * kTile;
    P.KLS = (d + 2 + kSlabK - 1) / kSlabK;
    P.EP = ((p + d + kTile - 1) / kTile) * kTile;
    P.psi_rows = 2 * P.MP + P.EP;
    const int MB = P.MP / kTile, EB = P.EP / kTile;
    P.nblk = 2 * MB + EB;

    // ---- accumulator tiles: lower triangle over the feature blocks, plus the [U;Y] products the fit needs ----
    std::vector<GramItem> sy;
    P.tile_of.assign((size_t)P.nblk * P.nblk, -1);
    auto add_tile = [&](int I, int J) {
        const int t = (int)sy.size();
        P.tile_of[(size_t)I * P.nblk + J] = t;
        sy.push_back(GramItem{kItemSyrk, I, J, t});
    };
    for (int I = 0; I < 2 * MB; I++) for (int J = 0; J <= I; J++) add_tile(I, J);
    const int u_blocks = (p + kTile - 1) / kTile;                 // E blocks holding control rows (0 or 1)
    for (int e = 0; e < EB; e++) {
        for (int J = 0; J < 2 * MB; J++) {
            const bool is_x = J < MB;
            if (is_x && e >= u_blocks) continue;                  // Y x Phi_x is not needed (regressors.py:164 pairs Y with Phi_y)
            add_tile(2 * MB + e, J);
        }
    }
    for (int e = 0; e < u_blocks; e++) for (int f = 0; f <= e; f++) add_tile(2 * MB + e, 2 * MB + f);   // U U^T
    P.ntiles = (int)sy.size();
    P.n_sy = P.ntiles;
    P.n_pk = chunk / kTile;
    P.n_lf = 2 * MB * (chunk / kTile);

    // ---- one period of the global work order: syrk(c) with pack(c+1) at 1/4 and lift(c+1) at 1/2 ----
    // Every item depends only on items EARLIER in this order (pack(c) <- syrk(c-S); lift(c) <- pack(c); syrk(c) <- lift(c) and
    // the same tile of syrk(c-1)): with all CTAs resident and claiming in order, nobody can wait for an unclaimed item.
    P.period.clear();
    const int q1 = P.n_sy / 4, q2 = P.n_sy / 2;
    for (int i = 0; i < q1; i++) P.period.push_back(sy[i]);
    for (int sb = 0; sb < P.n_pk; sb++) P.period.push_back(GramItem{kItemPack, sb, 0, 0});
    for (int i = q1; i < q2; i++) P.period.push_back(sy[i]);
    for (int sb = 0; sb < chunk / kTile; sb++)
        for (int side = 0; side < 2; side++)
            for (int lb = 0; lb < MB; lb++) P.period.push_back(GramItem{kItemLift, side, lb, sb});
    for (int i = q2; i < P.n_sy; i++) P.period.push_back(sy[i]);

    // Chunks in flight.  With many landmarks one chunk's items (thousands of Gram tiles) fill the GPU and two buffers suffice;
    // with few (the script configurations: m = 10 ... 400 gives 18 ... 150 items per chunk) the persistent CTAs would idle
    // behind the pack -> lift -> Gram chain of a single chunk, so several chunks are kept in flight (one buffer set each).
    const int period_len = (int)P.period.size();
    P.nslots = 2;
    if (period_len < 2 * sm_count) {
        P.nslots = (3 * sm_count + period_len - 1) / period_len;
        P.nslots = std::min(std::max(P.nslots, 2), kMaxSlots);
    }
}

}  // namespace nk

using namespace nk;

extern "C" {

int nk_version(void) { return 100; }

int nk_create(nk_handle **out, int device) {
    if (!out) return set_err(nullptr, NK_E_INVALID, "nk_create: out is NULL");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        return set_err(nullptr, NK_E_CUDA, "nk_create: no CUDA device (this library has no CPU fallback)");
    }
    if (device < 0 || device >= count) return set_err(nullptr, NK_E_INVALID, "nk_create: bad device index");
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) return set_err(nullptr, NK_E_CUDA, std::string("cudaGetDeviceProperties: ") + cudaGetErrorString(e));
    if (prop.major != 10) return set_err(nullptr, NK_E_CUDA, "nk_create: device is not sm_100 (kernels are built for sm_100a only)");
    {
        DeviceScope scope(device);      // makes sure the device can be selected; the caller's current device is left alone
        if (scope.err != cudaSuccess) return set_err(nullptr, NK_E_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(scope.err));
    }
    nk_handle *h = new nk_handle();
    h->device = device;
    h->sm_count = prop.multiProcessorCount;
    *out = h;
    return NK_OK;
}

int nk_destroy(nk_handle *h) {
    if (!h) return NK_OK;
    DeviceScope scope(h->device);
    nk_devbuf *bufs[] = {&h->zp, &h->zp_in, &h->gram_err, &h->inv_ls, &h->center, &h->gws, &h->items, &h->counters, &h->tile_of, &h->dinfo};
    for (nk_devbuf *b : bufs) if (b->ptr) cudaFree(b->ptr);
    for (int s = 0; s < kMaxSlots; s++) for (nk_devbuf *b : {&h->xp[s], &h->yp[s], &h->psi[s]}) if (b->ptr) cudaFree(b->ptr);
    for (nk_devbuf &b : h->dense) if (b.ptr) cudaFree(b.ptr);
    if (h->gram_err_host) cudaFreeHost(h->gram_err_host);
    if (h->side_stream) cudaStreamDestroy(h->side_stream);
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->ev_join) cudaEventDestroy(h->ev_join);
    delete h;
    return NK_OK;
}

int nk_release_scratch(nk_handle *h) {
    if (!h) return NK_E_INVALID;
    h->gram_open = false;      // an accumulation in progress is discarded: nk_gram_update / _finalize need a new nk_gram_begin
    NK_ON_DEVICE(h);
    NK_CUDA(h, cudaDeviceSynchronize());
    nk_devbuf *bufs[] = {&h->zp, &h->zp_in, &h->gram_err, &h->inv_ls, &h->center, &h->gws, &h->items, &h->counters, &h->tile_of, &h->dinfo};
    for (nk_devbuf *b : bufs) if (b->ptr) { cudaFree(b->ptr); b->ptr = nullptr; b->bytes = 0; }
    for (int s = 0; s < kMaxSlots; s++)
        for (nk_devbuf *b : {&h->xp[s], &h->yp[s], &h->psi[s]}) if (b->ptr) { cudaFree(b->ptr); b->ptr = nullptr; b->bytes = 0; }
    for (nk_devbuf &b : h->dense) if (b.ptr) { cudaFree(b.ptr); b.ptr = nullptr; b.bytes = 0; }
    return NK_OK;
}

const char *nk_last_error_string(nk_handle *h) { return h ? h->err.c_str() : g_create_err.c_str(); }
int nk_device_sm_count(nk_handle *h) { return h ? h->sm_count : 0; }
double nk_gram_last_executed_flops(nk_handle *h) { return h ? h->last_flops : 0.0; }
long long nk_launch_count(nk_handle *h) { return h ? h->launches : 0; }

int nk_probe_dmma_tflops(nk_handle *h, double ms_target, double *tflops) {
    if (!h || !tflops) return NK_E_INVALID;
    NK_ON_DEVICE(h);
    int rc;
    const int ctas = h->sm_count * 2;
    if ((rc = ensure(h, h->dense[11], (size_t)ctas * 256 * 8)) != NK_OK) return rc;
    cudaEvent_t e0, e1;
    NK_CUDA(h, cudaEventCreate(&e0));
    NK_CUDA(h, cudaEventCreate(&e1));
    int iters = 2000;
    float ms = 0.f;
    for (int pass = 0; pass < 2; pass++) {     // pass 0 calibrates the loop count, pass 1 is the measurement
        NK_CUDA(h, cudaEventRecord(e0, 0));
        dmma_probe_kernel<<<ctas, 256>>>((double *)h->dense[11].ptr, iters, 1.0, 1e-9);
        NK_CUDA(h, cudaEventRecord(e1, 0));
        NK_CUDA(h, cudaEventSynchronize(e1));
        NK_CUDA(h, cudaEventElapsedTime(&ms, e0, e1));
        if (pass == 0) { double scale = ms_target / (ms > 1e-3 ? ms : 1e-3); iters = (int)(iters * (scale < 1.0 ? 1.0 : scale)); }
    }
    h->launches += 2;
    *tflops = 2.0 * 256.0 * 32.0 * (double)iters * 8.0 * ctas / (ms * 1e-3) * 1e-12;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return NK_OK;
}

int nk_gram_plan(int m, int d, int p, int chunk, int sm_count, int *summary, int *items, int items_cap) {
    if (m < 1 || d < 1 || p < 0 || p > kTile || sm_count < 1 || !summary) return NK_E_INVALID;
    GramPlan P;
    plan_gram(m, d, p, chunk, sm_count, P);
    const int vals[12] = {P.chunk, P.MP, P.KLS, P.EP, P.psi_rows, P.nblk, P.ntiles, P.n_pk, P.n_lf, P.n_sy, (int)P.period.size(), P.nslots};
    for (int i = 0; i < 12; i++) summary[i] = vals[i];
    if (items) {
        const int n = std::min((int)P.period.size(), items_cap);
        for (int i = 0; i < n; i++) { items[4 * i] = P.period[i].type; items[4 * i + 1] = P.period[i].a; items[4 * i + 2] = P.period[i].b; items[4 * i + 3] = P.period[i].c; }
    }
    return (int)P.period.size();
}

int nk_gram_begin(nk_handle *h, const double *Z, long long ldz, int m, int d, int p, const double *inv_ls, int kind,
                  int chunk, void *stream_) {
    return nk_gram_begin_io(h, Z, ldz, Z, ldz, m, d, p, inv_ls, kind, chunk, stream_);
}

int nk_gram_begin_io(nk_handle *h, const double *Z_in, long long ldzi, const double *Z, long long ldz, int m, int d, int p,
                     const double *inv_ls, int kind, int chunk, void *stream_) {
    if (!h) return NK_E_INVALID;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!Z || !Z_in || !inv_ls || m < 1 || d < 1 || p < 0 || ldz < d || ldzi < d) return set_err(h, NK_E_INVALID, "nk_gram_begin: bad argument");
    if (kind != NK_KERNEL_RBF && kind != NK_KERNEL_MATERN52) return set_err(h, NK_E_INVALID, "nk_gram_begin: unsupported kernel kind");
    if (p > kTile) return set_err(h, NK_E_INVALID, "nk_gram_begin: more than 128 control inputs are not supported");
    NK_ON_DEVICE(h);
    GramPlan plan;
    plan_gram(m, d, p, chunk, h->sm_count, plan);
    chunk = plan.chunk;
    h->m = m; h->d = d; h->p = p; h->kind = kind; h->nk_chunk = chunk;
    h->MP = plan.MP; h->KLS = plan.KLS; h->EP = plan.EP; h->psi_rows = plan.psi_rows; h->nblk = plan.nblk;
    h->h_tile_of = plan.tile_of;
    h->ntiles = plan.ntiles; h->n_sy = plan.n_sy; h->n_pk = plan.n_pk; h->n_lf = plan.n_lf;
    h->period_len = (int)plan.period.size();
    h->nslots = plan.nslots;
    const std::vector<GramItem> &period = plan.period;

    int rc;
    if ((rc = ensure(h, h->zp, (size_t)h->MP * h->KLS * kSlabK * 8)) != NK_OK) return rc;
    h->distinct_in = (Z_in != Z);
    if (h->distinct_in && (rc = ensure(h, h->zp_in, (size_t)h->MP * h->KLS * kSlabK * 8)) != NK_OK) return rc;
    if ((rc = ensure(h, h->inv_ls, (size_t)d * 8)) != NK_OK) return rc;
    if ((rc = ensure(h, h->center, (size_t)d * 8)) != NK_OK) return rc;
    for (int s = 0; s < h->nslots; s++) {
        if ((rc = ensure(h, h->xp[s], (size_t)chunk * h->KLS * kSlabK * 8)) != NK_OK) return rc;
        if ((rc = ensure(h, h->yp[s], (size_t)chunk * h->KLS * kSlabK * 8)) != NK_OK) return rc;
        if ((rc = ensure(h, h->psi[s], (size_t)h->psi_rows * chunk * 8)) != NK_OK) return rc;
    }
    if ((rc = ensure(h, h->gws, (size_t)h->ntiles * kTile * kTile * 8)) != NK_OK) return rc;
    if ((rc = ensure(h, h->gram_err, 16)) != NK_OK) return rc;
    if (!h->gram_err_host) { NK_CUDA(h, cudaHostAlloc((void **)&h->gram_err_host, 16, cudaHostAllocDefault)); *h->gram_err_host = 0; }
    NK_CUDA(h, cudaMemsetAsync(h->gram_err.ptr, 0, 16, stream));
    if ((rc = ensure(h, h->items, period.size() * sizeof(GramItem))) != NK_OK) return rc;
    if ((rc = ensure(h, h->counters, (size_t)(kCounterTileVer + h->ntiles) * sizeof(int))) != NK_OK) return rc;
    if ((rc = ensure(h, h->tile_of, h->h_tile_of.size() * sizeof(int))) != NK_OK) return rc;

    NK_CUDA(h, cudaMemcpyAsync(h->items.ptr, period.data(), period.size() * sizeof(GramItem), cudaMemcpyHostToDevice, stream));
    NK_CUDA(h, cudaMemcpyAsync(h->tile_of.ptr, h->h_tile_of.data(), h->h_tile_of.size() * sizeof(int), cudaMemcpyHostToDevice, stream));
    NK_CUDA(h, cudaStreamSynchronize(stream));   // `period` and h_tile_of staging are host temporaries
    NK_CUDA(h, cudaMemcpyAsync(h->inv_ls.ptr, inv_ls, (size_t)d * 8, cudaMemcpyDeviceToDevice, stream));
    landmark_center(Z, ldz, m, d, (double *)h->center.ptr, stream);
    launch_pack_landmarks(Z, ldz, m, d, h->MP, h->KLS, (const double *)h->inv_ls.ptr, (const double *)h->center.ptr,
                          (double *)h->zp.ptr, stream);
    // distinct input landmarks share the shift of the output landmarks (distances are shift invariant; one shift per product)
    if (h->distinct_in) {
        launch_pack_landmarks(Z_in, ldzi, m, d, h->MP, h->KLS, (const double *)h->inv_ls.ptr, (const double *)h->center.ptr,
                              (double *)h->zp_in.ptr, stream);
        h->launches++;
    }
    NK_CUDA(h, cudaMemsetAsync(h->gws.ptr, 0, (size_t)h->ntiles * kTile * kTile * 8, stream));
    NK_CUDA(h, cudaGetLastError());
    h->launches += 2;
    h->gram_open = true;
    h->last_flops = 0.0;
    return NK_OK;
}

int nk_gram_update(nk_handle *h, const double *X, long long ldx, const double *Y, long long ldy, long long n, void *stream_) {
    if (!h) return NK_E_INVALID;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!h->gram_open) return set_err(h, NK_E_STATE, "nk_gram_update: call nk_gram_begin first");
    if (n == 0) return NK_OK;
    if (!X || !Y || n < 0 || ldx < h->d + h->p || ldy < h->d) return set_err(h, NK_E_INVALID, "nk_gram_update: bad argument");
    NK_ON_DEVICE(h);
    const long long n_chunks = (n + h->nk_chunk - 1) / h->nk_chunk;
    if ((n_chunks + 1) * (long long)h->period_len > 2000000000LL || (n_chunks / h->nslots + 1) * (long long)h->n_sy * kConsumerWarps > 2000000000LL
        || n_chunks * (long long)kConsumerWarps > 2000000000LL)
        return set_err(h, NK_E_INVALID, "nk_gram_update: too many chunks for one call; split the sample block");
    GramParams P;
    P.X = X; P.ldx = ldx; P.Y = Y; P.ldy = ldy; P.n = n;
    P.d = h->d; P.p = h->p; P.m = h->m; P.kind = h->kind;
    P.MP = h->MP; P.KLS = h->KLS; P.nk = h->nk_chunk; P.n_chunks = (int)n_chunks;
    P.psi_rp = h->psi_rows / kPanel; P.e_row0 = 2 * h->MP; P.EP = h->EP;
    P.ZP = (const double *)h->zp.ptr; P.ZPx = h->distinct_in ? (const double *)h->zp_in.ptr : P.ZP; P.inv_ls = (const double *)h->inv_ls.ptr; P.center = (const double *)h->center.ptr;
    P.nslots = h->nslots;
    for (int s = 0; s < kMaxSlots; s++) {
        const int u = s < h->nslots ? s : 0;
        P.XP[s] = (double *)h->xp[u].ptr; P.YP[s] = (double *)h->yp[u].ptr; P.PSI[s] = (double *)h->psi[u].ptr;
    }
    P.Gws = (double *)h->gws.ptr;
    P.items = (const GramItem *)h->items.ptr;
    P.period_len = h->period_len; P.n_pk = h->n_pk; P.n_lf = h->n_lf; P.n_sy = h->n_sy;
    P.counters = (int *)h->counters.ptr;
    P.err = (int *)h->gram_err.ptr;
#ifdef NK_GRAM_TIMING
    { int trc = ensure(h, h->dense[15], (size_t)h->sm_count * kConsumerWarps * 16 * sizeof(long long)); if (trc != NK_OK) return trc; }
    P.timing = (long long *)h->dense[15].ptr;
    NK_CUDA(h, cudaMemsetAsync(P.timing, 0, (size_t)h->sm_count * kConsumerWarps * 16 * sizeof(long long), stream));
#endif
    NK_CUDA(h, cudaMemsetAsync(h->counters.ptr, 0, (size_t)(kCounterTileVer + h->ntiles) * sizeof(int), stream));
    cudaError_t e = cudaSuccess;
    launch_gram(P, h->sm_count, stream, &e);
    NK_CUDA(h, e);
    h->launches += 1;
    const double per_tile = 2.0 * kTile * kTile;
    h->last_flops = (double)n_chunks * ((double)h->n_sy * per_tile * h->nk_chunk + (double)h->n_lf * per_tile * h->KLS * kSlabK);
    return NK_OK;
}

#ifdef NK_GRAM_TIMING
// development build only: copies the per-warp cycle counters of the last nk_gram_update to the host (synchronises)
int nk_debug_gram_timing(nk_handle *h, long long *out, int count) {
    if (!h || !out || !h->dense[15].ptr) return NK_E_INVALID;
    cudaDeviceSynchronize();
    const size_t have = (size_t)h->sm_count * kConsumerWarps * 16;
    cudaMemcpy(out, h->dense[15].ptr, sizeof(long long) * ((size_t)count < have ? (size_t)count : have), cudaMemcpyDeviceToHost);
    return NK_OK;
}
#endif

int nk_gram_finalize(nk_handle *h, double *Gxx, long long ld_gxx, double *Gyx, long long ld_gyx, double *Gyy, long long ld_gyy,
                     double *Gxu, long long ld_gxu, double *Gyu, long long ld_gyu, double *Guu, long long ld_guu,
                     double *GYy, long long ld_gYy, int accumulate, void *stream_) {
    if (!h) return NK_E_INVALID;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!h->gram_open) return set_err(h, NK_E_STATE, "nk_gram_finalize: call nk_gram_begin first");
    NK_ON_DEVICE(h);
    const double *G = (const double *)h->gws.ptr;
    const int *tof = (const int *)h->tile_of.ptr;
    const int m = h->m, p = h->p, d = h->d, MP = h->MP, E0 = 2 * h->MP, nb = h->nblk;
    if (Gxx) { launch_unpack(G, tof, nb, 0, 0, m, m, Gxx, ld_gxx, accumulate, stream); h->launches++; }
    if (Gyx) { launch_unpack(G, tof, nb, MP, 0, m, m, Gyx, ld_gyx, accumulate, stream); h->launches++; }
    if (Gyy) { launch_unpack(G, tof, nb, MP, MP, m, m, Gyy, ld_gyy, accumulate, stream); h->launches++; }
    if (Gxu && p) { launch_unpack(G, tof, nb, 0, E0, m, p, Gxu, ld_gxu, accumulate, stream); h->launches++; }
    if (Gyu && p) { launch_unpack(G, tof, nb, MP, E0, m, p, Gyu, ld_gyu, accumulate, stream); h->launches++; }
    if (Guu && p) { launch_unpack(G, tof, nb, E0, E0, p, p, Guu, ld_guu, accumulate, stream); h->launches++; }
    if (GYy) { launch_unpack(G, tof, nb, E0 + p, MP, d, m, GYy, ld_gYy, accumulate, stream); h->launches++; }
    // latch the watchdog flag for the next synchronising call (stream-ordered; nothing waits here)
    NK_CUDA(h, cudaMemcpyAsync(h->gram_err_host, h->gram_err.ptr, sizeof(int), cudaMemcpyDeviceToHost, stream));
    NK_CUDA(h, cudaGetLastError());
    return NK_OK;
}

int nk_allreduce_grams(nk_handle *h, void *nccl_comm, double *packed, long long count, void *stream_) {
    if (!h) return NK_E_INVALID;
    if (!nccl_comm || !packed || count < 0) return set_err(h, NK_E_INVALID, "nk_allreduce_grams: bad argument");
    if (count == 0) return NK_OK;
    // int ncclAllReduce(const void *send, void *recv, size_t count, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);  nccl.h:
    // ncclDouble = 8 (ncclFloat64), ncclSum = 0 -- stable across NCCL 2.x.  Resolved from the NCCL the HOST already loaded.
    typedef int (*AllReduceFn)(const void *, void *, size_t, int, int, void *, cudaStream_t);
    typedef const char *(*ErrStrFn)(int);
    static AllReduceFn fn = nullptr;
    static ErrStrFn errstr = nullptr;
    if (!fn) {
        fn = (AllReduceFn)dlsym(RTLD_DEFAULT, "ncclAllReduce");
        errstr = (ErrStrFn)dlsym(RTLD_DEFAULT, "ncclGetErrorString");
        if (!fn) {
            for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
                void *lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL | RTLD_NOLOAD);      // already mapped by the host (e.g. torch's bundled copy)?
                if (lib) { fn = (AllReduceFn)dlsym(lib, "ncclAllReduce"); errstr = (ErrStrFn)dlsym(lib, "ncclGetErrorString"); if (fn) break; }
            }
        }
    }
    if (!fn) return set_err(h, NK_E_STATE, "nk_allreduce_grams: no NCCL is loaded in this process (the host creates the communicator, so it must have loaded one)");
    NK_ON_DEVICE(h);
    const int rc = fn(packed, packed, (size_t)count, 8 /* ncclDouble */, 0 /* ncclSum */, nccl_comm, (cudaStream_t)stream_);
    if (rc != 0) return set_err(h, NK_E_CUDA, std::string("ncclAllReduce: ") + (errstr ? errstr(rc) : "error ") + " (" + std::to_string(rc) + ")");
    return NK_OK;
}

int nk_gram_status(nk_handle *h, void *stream_) {
    if (!h) return NK_E_INVALID;
    NK_ON_DEVICE(h);
    NK_CUDA(h, cudaStreamSynchronize((cudaStream_t)stream_));
    return gram_watchdog_verdict(h);
}

}  // extern "C"
