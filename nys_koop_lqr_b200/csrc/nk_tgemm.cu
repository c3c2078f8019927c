// nk_tgemm.cu -- the dense stage's NT GEMM on the TMA engine: C = alpha A B^T + beta C (+ diag I), row-major operands.
//
// The first dense GEMM (gemm_nt_kernel, nk_dense.cu) gathers operands with per-thread 16-byte cp.async (LDGSTS), synchronises
// the whole CTA once per 16-deep slab and runs one tile per CTA: 82% of the DMMA issue rate.  This kernel keeps the row-major
// operands where they are and lets the TMA engine do the gathering:
//
//   * one 3-D tensor map per operand (k, row, batch), box = 16 doubles x 128 rows = one 16 KB slab, SWIZZLE_128B: the engine
//     writes row r of the box at r*128 B with its eight 16-byte chunks XOR-permuted by (r & 7); out-of-range rows / depth are
//     zero-filled, so ragged M, N, K need no masks on the load side (SASS: UTMALDG);
//   * fragment loads stay one conflict-free LDS.128 per lane: lane (g, t) of k8-step q reads logical chunk 2t+q of its row,
//     i.e. physical chunk (2t+q) ^ g -- within a quarter warp (g in {0,1}, t in 0..3) that is all eight chunks, 32 banks once.
//     Which contraction indices a (step, lane) pair handles is the same for both operands, which is all the product needs;
//   * persistent CTAs (one per SM): a producer lane walks the CTA's tile list and keeps a 4-stage mbarrier ring full, eight
//     consumer warps run the prefetched 64x32 DMMA main loop of nk_mainloop.cuh with no CTA-wide barrier; the epilogue of a
//     tile overlaps with the loads of the next;
//   * the same epilogue options as gemm_nt (lower-triangle-only tile list with mirrored store, transposed second output,
//     diagonal shift, packed-operand output, kernel-function epilogue), batched over a third tensor-map dimension.
//
// gemm_nt_batched() routes a product here when the operands are 16-byte aligned with even leading dimensions and the product
// is large enough to fill the GPU; everything else (tiny products of the script-sized fits, odd strides) stays on gemm_nt_kernel.
#include <cuda.h>
#include <cstdlib>
#include "nk_dense.cuh"
#include "nk_mainloop.cuh"

namespace nk {

constexpr int kTgStages = 4;
constexpr size_t kTgStageBytes = (size_t)kTgStages * 2 * kSlabTileDoubles * 8;   // 128 KB
constexpr size_t kTgSmemBytes = kTgStageBytes + 256;
constexpr int kTgBandH = 8;

struct TgCtl {
    uint64_t full[kTgStages];
    uint64_t empty[kTgStages];
};

struct TGemmArgs {
    int M, N, K, KS;
    double alpha, beta, diag;
    double *C; long long ldc, sC;
    double *Ct; long long ldct, sCt;
    int flags, epi_kind;
    int tiles_m, tiles_n, tiles_per_mat, n_tiles;
    int a_batched, b_batched;
};

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *map, int c0, int c1, int c2, uint32_t bar, uint64_t policy) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3, %4}], [%5], %6;"
                 ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar), "l"(policy) : "memory");
}

__device__ __forceinline__ void tg_tile(const TGemmArgs &g, int t, int &bz, int &I, int &J) {
    bz = t / g.tiles_per_mat;
    const int w = t - bz * g.tiles_per_mat;
    if (g.flags & kGemmLowerOnly) {
        int r = (int)((sqrt(8.0 * w + 1.0) - 1.0) * 0.5);
        while ((r + 1) * (r + 2) / 2 <= w) r++;
        while (r * (r + 1) / 2 > w) r--;
        I = r; J = w - r * (r + 1) / 2;
    } else {
        const int per_band = kTgBandH * g.tiles_n;
        const int band = w / per_band, within = w - band * per_band;
        const int hgt = min(kTgBandH, g.tiles_m - band * kTgBandH);
        J = within / hgt;
        I = band * kTgBandH + (within - J * hgt);
    }
}

__global__ void __launch_bounds__(kThreads, 1) tgemm_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                                                            const TGemmArgs g) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    TgCtl *ctl = reinterpret_cast<TgCtl *>(smem_raw + kTgStageBytes);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < kTgStages; s++) { mbar_init(&ctl->full[s], 1); mbar_init(&ctl->empty[s], kConsumerWarps); }
        fence_mbar_init();
    }
    __syncthreads();
    const uint32_t sm0 = opaque(smem_u32(smem_raw));
    const uint32_t full0 = opaque(smem_u32(&ctl->full[0])), empty0 = opaque(smem_u32(&ctl->empty[0]));

    if (warp >= kConsumerWarps) {
        // =============================== producer warpgroup ===============================
        setmaxnreg_dec<40>();
        if (warp == kConsumerWarps && lane == 0) {
            const uint64_t pol = policy_evict_last();
            uint32_t stage = 0, phase = 0;
            for (int tle = blockIdx.x; tle < g.n_tiles; tle += gridDim.x) {
                int bz, I, J;
                tg_tile(g, tle, bz, I, J);
                const int za = g.a_batched ? bz : 0, zb = g.b_batched ? bz : 0;
                for (int s = 0; s < g.KS; s++) {
                    mbar_wait_a(empty0 + stage * 8, phase ^ 1);
                    const uint32_t dst = sm0 + stage * (uint32_t)(2 * kSlabTileDoubles * 8);
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full0 + stage * 8), "r"(2 * kSlabTileDoubles * 8) : "memory");
                    tma_load_3d(dst, &mapA, s * kSlabK, I * kTile, za, full0 + stage * 8, pol);
                    tma_load_3d(dst + kSlabTileDoubles * 8, &mapB, s * kSlabK, J * kTile, zb, full0 + stage * 8, pol);
                    if (++stage == kTgStages) { stage = 0; phase ^= 1; }
                }
            }
        }
        return;
    }

    // =============================== consumer warps ===============================
    setmaxnreg_inc<232>();
    const int wr = warp >> 2, wc = warp & 3, gq = lane >> 2, t = lane & 3;
    // swizzled fragment addresses of stage 0 for the two k8 steps of a slab (see the header comment)
    const uint32_t a_q0 = opaque(sm0 + (uint32_t)(wr * 64 + gq) * 128u + (uint32_t)(((2 * t) ^ gq) << 4));
    const uint32_t a_q1 = opaque(sm0 + (uint32_t)(wr * 64 + gq) * 128u + (uint32_t)(((2 * t + 1) ^ gq) << 4));
    const uint32_t b_q0 = opaque(sm0 + 16384u + (uint32_t)(wc * 32 + gq) * 128u + (uint32_t)(((2 * t) ^ gq) << 4));
    const uint32_t b_q1 = opaque(sm0 + 16384u + (uint32_t)(wc * 32 + gq) * 128u + (uint32_t)(((2 * t + 1) ^ gq) << 4));
    uint32_t stage = 0, sphase = 0;
    const bool fast_c = g.epi_kind < 0 && (g.flags & (kGemmMirror | kGemmStoreT | kGemmPackedOut)) == 0 && g.C != nullptr && (g.ldc % 2 == 0) &&
                        (g.sC % 2 == 0) && (((uintptr_t)g.C) % 16 == 0);
    const bool mirror = (g.flags & kGemmMirror) != 0;

    for (int tle = blockIdx.x; tle < g.n_tiles; tle += gridDim.x) {
        int bz, I, J;
        tg_tile(g, tle, bz, I, J);
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
            for (int j = 0; j < 4; j++) b[j] = lds_v2(b_q0 + so + j * 1024);
            a0 = lds_v2(a_q0 + so);
        }
        for (int s = 0; s < g.KS; s++) {
            const uint32_t so = stage * (uint32_t)(2 * kSlabTileDoubles * 8);
            uint32_t nstage = stage + 1, nphase = sphase;
            if (nstage == kTgStages) { nstage = 0; nphase ^= 1; }
            const bool has_next = (s + 1 < g.KS);
            const uint32_t ready = has_next ? mbar_test_a(full0 + nstage * 8, nphase) : 1u;
            // q = 0 (prefetch q = 1 of this slab)
#pragma unroll
            for (int j = 0; j < 4; j++) nb[j] = lds_v2(b_q1 + so + j * 1024);
            na0 = lds_v2(a_q1 + so);
            k8_step(acc, a_q0 + so, a0, b);
#pragma unroll
            for (int j = 0; j < 4; j++) b[j] = nb[j];
            a0 = na0;
            // q = 1 (prefetch q = 0 of the next slab)
            if (has_next) {
                if (!ready) mbar_wait_a(full0 + nstage * 8, nphase);
                const uint32_t no = nstage * (uint32_t)(2 * kSlabTileDoubles * 8);
#pragma unroll
                for (int j = 0; j < 4; j++) nb[j] = lds_v2(b_q0 + no + j * 1024);
                na0 = lds_v2(a_q0 + no);
            }
            k8_step(acc, a_q1 + so, a0, b);
            __syncwarp();
            if (lane == 0) mbar_arrive_a(empty0 + stage * 8);
#pragma unroll
            for (int j = 0; j < 4; j++) b[j] = nb[j];
            a0 = na0;
            stage = nstage; sphase = nphase;
        }

        // ---- epilogue (the producer is already loading the next tile) ----
        double *Cb = g.C ? g.C + (long long)bz * g.sC : nullptr;
        double *Ctb = g.Ct ? g.Ct + (long long)bz * g.sCt : nullptr;
        if (fast_c) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const int r = I * kTile + wr * 64 + i * 8 + gq;
                if (r >= g.M) continue;
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const int c0 = J * kTile + wc * 32 + j * 8 + 2 * t;
                    if (c0 >= g.N) continue;
                    double *o = Cb + (long long)r * g.ldc + c0;
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
            continue;
        }
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
                        if (g.beta != 0.0) v += g.beta * Cb[(long long)r * g.ldc + c];
                        if (r == c) v += g.diag;
                    }
                    if (g.flags & kGemmPackedOut) { Cb[packed_off(r, c, (int)g.ldc)] = v; continue; }
                    if (Cb) Cb[(long long)r * g.ldc + c] = v;
                    if (mirror && c != r) Cb[(long long)c * g.ldc + r] = v;
                    if (g.flags & kGemmStoreT) Ctb[(long long)c * g.ldct + r] = v;
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        const char *off = getenv("NK_TGEMM");
        if (off && off[0] == '0') return nullptr;      // development switch: A/B against gemm_nt_kernel
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
        cudaGetLastError();
    }
    return fn;
}

static bool make_map(EncodeTiledFn enc, CUtensorMap *map, const double *base, int rows, int K, long long ld, int batch, long long stride) {
    const bool batched = batch > 1 && stride != 0;
    cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)rows, (cuuint64_t)(batched ? batch : 1)};
    cuuint64_t strides[2] = {(cuuint64_t)ld * 8, (cuuint64_t)(batched ? stride : (long long)rows * ld) * 8};
    if (strides[1] == 0) strides[1] = 16;
    cuuint32_t box[3] = {(cuuint32_t)kSlabK, (cuuint32_t)kTile, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, (void *)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// returns true if the product was launched here
bool tgemm_try(nk_handle *h, int batch, int M, int N, int K, double alpha, const double *A, long long lda, long long sA, const double *B,
               long long ldb, long long sB, double beta, double *C, long long ldc, long long sC, double diag, int flags, double *Ct,
               long long ldct, long long sCt, cudaStream_t stream, int epi_kind) {
    EncodeTiledFn enc = encode_fn();
    if (!enc || K < 1) return false;
    if (((lda | ldb | sA | sB) & 1) || (((uintptr_t)A | (uintptr_t)B) & 15)) return false;      // TMA: 16-byte aligned base and strides
    if ((sA < 0) || (sB < 0)) return false;
    const int tm = (M + kTile - 1) / kTile, tn = (N + kTile - 1) / kTile;
    const long long per = (flags & kGemmLowerOnly) ? (long long)tm * (tm + 1) / 2 : (long long)tm * tn;
    const long long n_tiles = per * batch;
    const long long KS = (K + kSlabK - 1) / kSlabK;
    // small products (script-sized fits) are latency-bound: the plain kernel launches faster (no descriptors, no persistence)
    if (n_tiles < 16 || n_tiles * KS < 512 || n_tiles > 2000000000LL) return false;
    CUtensorMap mapA, mapB;
    if (!make_map(enc, &mapA, A, M, K, lda, batch, sA) || !make_map(enc, &mapB, B, N, K, ldb, batch, sB)) return false;
    static unsigned long long configured = 0;
    if (first_use_on_device(configured)) cudaFuncSetAttribute(tgemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTgSmemBytes);
    TGemmArgs g;
    g.M = M; g.N = N; g.K = K; g.KS = (int)KS; g.alpha = alpha; g.beta = beta; g.diag = diag;
    g.C = C; g.ldc = ldc; g.sC = sC; g.Ct = Ct; g.ldct = ldct; g.sCt = sCt; g.flags = flags; g.epi_kind = epi_kind;
    g.tiles_m = tm; g.tiles_n = tn; g.tiles_per_mat = (int)per; g.n_tiles = (int)n_tiles;
    g.a_batched = (batch > 1 && sA != 0) ? 1 : 0; g.b_batched = (batch > 1 && sB != 0) ? 1 : 0;
    const int grid = n_tiles < h->sm_count ? (int)n_tiles : h->sm_count;
    tgemm_kernel<<<grid, kThreads, kTgSmemBytes, stream>>>(mapA, mapB, g);
    h->launches++;
    return true;
}

}  // namespace nk
