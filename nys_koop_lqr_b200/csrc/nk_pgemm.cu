// nk_pgemm.cu -- packed-operand persistent NT GEMM on FP64 tensor cores.
//
// The plain NT GEMM of nk_dense.cu gathers row-major operands with 16-byte cp.async and synchronises the whole CTA per
// 16-deep slab (82% of the DMMA issue rate).  Where the same operand is contracted many times -- the batched rollout
// z <- Az + Bu, y = Cz over 10^5 trajectories (benchmark_lqr_cloth.py:29-32 for one trajectory), CV scoring -- it pays to
// keep operands in the packed 8x8-block layout of the fused lift+Gram engine: a 128-row x 16-deep slab is then one
// contiguous 16 KB run that the TMA engine copies with a single cp.async.bulk (SASS UBLKCP), a dedicated producer lane
// feeds a 4-stage mbarrier ring, and the eight consumer warps run the same prefetched DMMA main loop as nk_gram.cu
// without any CTA-wide barrier.  One CTA per SM walks a static, band-swizzled tile order (8 row tiles share a B tile in
// L2).  The epilogue can write the result directly in packed form with the result column as contraction index, which
// makes the output of one rollout step the A operand of the next.
#include "nk_pgemm.cuh"
#include "nk_mainloop.cuh"

namespace nk {

constexpr int kPgStages = 4;
constexpr size_t kPgStageBytes = (size_t)kPgStages * 2 * kSlabTileDoubles * 8;   // 128 KB
constexpr size_t kPgSmemBytes = kPgStageBytes + 256;
constexpr int kBandH = 8;

struct PgCtl {
    uint64_t full[kPgStages];
    uint64_t empty[kPgStages];
};

__device__ __forceinline__ void tile_coords(int t, int tiles_m, int tiles_n, int &tm, int &tn) {
    const int per_band = kBandH * tiles_n;
    const int band = t / per_band, within = t - band * per_band;
    const int h = min(kBandH, tiles_m - band * kBandH);
    tn = within / h;
    tm = band * kBandH + (within - tn * h);
}

__global__ void __launch_bounds__(kThreads, 1) pgemm_kernel(const PGemmParams P) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    double *stage_base = reinterpret_cast<double *>(smem_raw);
    PgCtl *ctl = reinterpret_cast<PgCtl *>(smem_raw + kPgStageBytes);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < kPgStages; s++) { mbar_init(&ctl->full[s], 1); mbar_init(&ctl->empty[s], kConsumerWarps); }
        fence_mbar_init();
    }
    __syncthreads();
    const int n_tiles = P.tiles_m * P.tiles_n;

    if (warp >= kConsumerWarps) {
        setmaxnreg_dec<40>();
        if (warp == kConsumerWarps && lane == 0) {
            const uint64_t pol = policy_evict_last();
            uint32_t stage = 0, phase = 0;
            for (int tle = blockIdx.x; tle < n_tiles; tle += gridDim.x) {
                int tm, tn;
                tile_coords(tle, P.tiles_m, P.tiles_n, tm, tn);
                const double *Ab = P.Ap + (size_t)tm * 16 * 128, *Bb = P.Bp + (size_t)tn * 16 * 128;
                const size_t as = (size_t)P.a_rp * 128, bs = (size_t)P.b_rp * 128;
                for (int s = 0; s < P.KS; s++) {
                    mbar_wait(&ctl->empty[stage], phase ^ 1);
                    double *As = stage_base + (size_t)stage * 2 * kSlabTileDoubles;
                    mbar_arrive_expect_tx(&ctl->full[stage], 2 * kSlabTileDoubles * 8);
                    bulk_g2s(As, Ab + s * as, kSlabTileDoubles * 8, &ctl->full[stage], pol);
                    bulk_g2s(As + kSlabTileDoubles, Bb + s * bs, kSlabTileDoubles * 8, &ctl->full[stage], pol);
                    if (++stage == kPgStages) { stage = 0; phase ^= 1; }
                }
            }
        }
        return;
    }

    setmaxnreg_inc<232>();
    const int wr = warp >> 2, wc = warp & 3, g = lane >> 2, t = lane & 3;
    const uint32_t sm0 = opaque(smem_u32(smem_raw));
    const uint32_t a_off = opaque(sm0 + (uint32_t)(wr * 8 * 2) * 512u + lane * 16u);
    const uint32_t b_off = opaque(sm0 + 16384u + (uint32_t)(wc * 4 * 2) * 512u + lane * 16u);
    const uint32_t full0 = opaque(smem_u32(&ctl->full[0])), empty0 = opaque(smem_u32(&ctl->empty[0]));
    uint32_t stage = 0, sphase = 0;
    const bool c_vec = P.C != nullptr && (P.ldc % 2 == 0) && (((uintptr_t)P.C) % 16 == 0) && (P.c_col0 % 2 == 0);

    for (int tle = blockIdx.x; tle < n_tiles; tle += gridDim.x) {
        int tm, tn;
        tile_coords(tle, P.tiles_m, P.tiles_n, tm, tn);
        double acc[8][4][2];
#pragma unroll
        for (int i = 0; i < 8; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }

        mbar_wait_a(full0 + stage * 8, sphase);
        double2 b[4], nb[4], a0, na0;
        {
            const uint32_t so = stage * (uint32_t)(2 * kSlabTileDoubles * 8);
#pragma unroll
            for (int j = 0; j < 4; j++) b[j] = lds_v2(b_off + so + j * 1024);
            a0 = lds_v2(a_off + so);
        }
        for (int s = 0; s < P.KS; s++) {
            const uint32_t so = stage * (uint32_t)(2 * kSlabTileDoubles * 8);
            uint32_t nstage = stage + 1, nphase = sphase;
            if (nstage == kPgStages) { nstage = 0; nphase ^= 1; }
            const bool has_next = (s + 1 < P.KS);
            const uint32_t ready = has_next ? mbar_test_a(full0 + nstage * 8, nphase) : 1u;
            // q = 0 (prefetch q = 1 of this slab)
#pragma unroll
            for (int j = 0; j < 4; j++) nb[j] = lds_v2(b_off + so + 512 + j * 1024);
            na0 = lds_v2(a_off + so + 512);
            k8_step(acc, a_off + so, a0, b);
#pragma unroll
            for (int j = 0; j < 4; j++) b[j] = nb[j];
            a0 = na0;
            // q = 1 (prefetch q = 0 of the next slab)
            if (has_next) {
                if (!ready) mbar_wait_a(full0 + nstage * 8, nphase);
                const uint32_t no = nstage * (uint32_t)(2 * kSlabTileDoubles * 8);
#pragma unroll
                for (int j = 0; j < 4; j++) nb[j] = lds_v2(b_off + no + j * 1024);
                na0 = lds_v2(a_off + no);
            }
            k8_step(acc, a_off + so + 512, a0, b);
            __syncwarp();
            if (lane == 0) mbar_arrive_a(empty0 + stage * 8);
#pragma unroll
            for (int j = 0; j < 4; j++) b[j] = nb[j];
            a0 = na0;
            stage = nstage; sphase = nphase;
        }

        // ---- epilogue: packed destination (next product's operand) and / or row-major destination ----
        const int r_base = tm * kTile + wr * 64, c_base = tn * kTile + wc * 32;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int r = r_base + i * 8 + g;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int c0 = c_base + j * 8 + 2 * t;
                const double v0 = P.alpha * acc[i][j][0], v1 = P.alpha * acc[i][j][1];
                if (P.Cp != nullptr && c0 < P.cp_cols) {
                    double *o = P.Cp + packed_off(r, c0, P.c_rp);
                    if (c0 + 1 < P.cp_cols) { asm volatile("st.global.v2.f64 [%0], {%1,%2};" ::"l"(o), "d"(v0), "d"(v1) : "memory"); }
                    else *o = v0;
                }
                if (P.C != nullptr && r < P.M && c0 + 1 >= P.c_col0 && c0 < P.N) {
                    double *o = P.C + (long long)r * P.ldc + (c0 - P.c_col0);
                    const bool in0 = c0 >= P.c_col0, in1 = (c0 + 1) < P.N;
                    if (c_vec && in0 && in1) {
                        double w0 = v0, w1 = v1;
                        if (P.beta != 0.0) { w0 += P.beta * o[0]; w1 += P.beta * o[1]; }
                        asm volatile("st.global.v2.f64 [%0], {%1,%2};" ::"l"(o), "d"(w0), "d"(w1) : "memory");
                    } else {
                        if (in0) o[0] = v0 + (P.beta != 0.0 ? P.beta * o[0] : 0.0);
                        if (in1) o[1] = v1 + (P.beta != 0.0 ? P.beta * o[1] : 0.0);
                    }
                }
            }
        }
    }
}

void launch_pgemm(nk_handle *h, const PGemmParams &P, cudaStream_t stream) {
    static unsigned long long configured = 0;
    if (first_use_on_device(configured)) cudaFuncSetAttribute(pgemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPgSmemBytes);
    const int n_tiles = P.tiles_m * P.tiles_n;
    if (n_tiles <= 0 || P.KS <= 0) return;
    const int grid = n_tiles < h->sm_count ? n_tiles : h->sm_count;
    pgemm_kernel<<<grid, kThreads, kPgSmemBytes, stream>>>(P);
    h->launches++;
}

// ------------------------------------------------------------------------------------------------
// row-major <-> packed
// ------------------------------------------------------------------------------------------------
__global__ void pack_rows_kernel(const double *src, long long ld, long long rows, int cols, double *dst, int rp, long long row0, int k0) {
    const long long total = rows * cols;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const long long r = idx / cols;
        const int c = (int)(idx - r * cols);
        dst[packed_off((int)(row0 + r), k0 + c, rp)] = src[r * ld + c];
    }
}
__global__ void unpack_rows_kernel(const double *src, int rp, long long row0, int k0, long long rows, int cols, double *dst, long long ld) {
    const long long total = rows * cols;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const long long r = idx / cols;
        const int c = (int)(idx - r * cols);
        dst[r * ld + c] = src[packed_off((int)(row0 + r), k0 + c, rp)];
    }
}
static int blocks_for(long long total) { long long b = (total + 255) / 256; return (int)(b < 16384 ? (b > 0 ? b : 1) : 16384); }

void pack_rows(nk_handle *h, const double *src, long long ld, long long rows, int cols, double *dst, int rp, long long row0, int k0,
               cudaStream_t stream) {
    if (rows <= 0 || cols <= 0) return;
    pack_rows_kernel<<<blocks_for(rows * cols), 256, 0, stream>>>(src, ld, rows, cols, dst, rp, row0, k0);
    h->launches++;
}
void unpack_rows(nk_handle *h, const double *src, int rp, long long row0, int k0, long long rows, int cols, double *dst, long long ld,
                 cudaStream_t stream) {
    if (rows <= 0 || cols <= 0) return;
    unpack_rows_kernel<<<blocks_for(rows * cols), 256, 0, stream>>>(src, rp, row0, k0, rows, cols, dst, ld);
    h->launches++;
}

}  // namespace nk
