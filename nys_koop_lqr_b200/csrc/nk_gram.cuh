// nk_gram.cuh -- parameter block of the fused lift+Gram engine (shared between nk_gram.cu and nk_api.cu)
#pragma once
#include "nk_common.cuh"

namespace nk {

enum ItemType : int { kItemPack = 0, kItemLift = 1, kItemSyrk = 2 };
constexpr int kMaxSlots = 16;

// One entry of the per-chunk work period.  pack: a = sample sub-block;  lift: a = side (0: x_t, 1: x_{t+1}),
// b = landmark block, c = sample sub-block;  syrk: a, b = 128-row blocks of the stacked feature matrix
// Psi = [Phi_x ; Phi_y ; U ; Y], c = accumulator tile.
struct GramItem { int type, a, b, c; };

struct GramParams {
    const double *X; long long ldx;   // (n, d+p) row-major: state then the p controls (regressors.py:123-126)
    const double *Y; long long ldy;   // (n, d) row-major next states
    long long n;
    int d, p, m, kind;
    int MP;        // landmarks padded to a multiple of 128
    int KLS;       // slabs of the lift contraction: ceil((d+2)/16)
    int nk;        // samples per chunk (multiple of 128)
    int n_chunks;
    int psi_rp;    // row panels of Psi (= (2*MP + EP)/8)
    int e_row0;    // first row of the [U;Y] block in Psi (= 2*MP)
    int EP;        // [U;Y] rows padded to a multiple of 128
    const double *ZP;       // packed, scaled, augmented OUTPUT landmarks (lift of x_{t+1}): MP/8 panels x KLS slabs
    const double *ZPx;      // the same for the INPUT landmarks (lift of x_t); == ZP unless the caller injected distinct input centres
    const double *inv_ls;   // (d) 1/length_scale
    const double *center;   // (d) shift applied to samples and landmarks before the norm expansion
    int nslots;             // chunk buffers in flight: chunk c lives in slot c % nslots (2 for large m; more when one chunk's items
                            // cannot fill the GPU, see nk_gram_begin)
    double *XP[kMaxSlots], *YP[kMaxSlots];  // packed scaled sample operands of the lift, one buffer per slot
    double *PSI[kMaxSlots];                 // packed feature chunk per slot: psi_rp panels x nk/16 slabs
    double *Gws;            // accumulator tiles in C-fragment order, 16384 doubles each
    const GramItem *items;  // one period: syrk(c) items with pack(c+1) and lift(c+1) items spliced in
    int period_len, n_pk, n_lf, n_sy;
#ifdef NK_GRAM_TIMING
    long long *timing;      // development build only: 16 cycle counters per consumer warp (tools/gram_timing.py)
#endif
    int *err;               // watchdog flag (handle-owned, zeroed by nk_gram_begin): set when a dependence wait timed out
    int *counters;          // [0] next item; per slot s: [kCtrPack+s] packs done, [kCtrLift+s] lift warps done, [kCtrSyrk+s] syrk warps done;
                            // [kCounterTileVer+t] tile versions
};

constexpr int kCtrPack = 16, kCtrLift = 32, kCtrSyrk = 48;
constexpr int kCounterTileVer = 64;
constexpr unsigned kSpinCap = 1u << 25;   // polls of a dependence wait before the watchdog fires (~1 us each)
// Operand ring depth: 5 stages of 2 x 16 KB slabs (160 KB) + 8 KB of accumulator / exponent staging per consumer warp (epilogues
// run in two passes of four fragment row blocks).  The first version had 3 stages + 16 KB staging (-DNK_GRAM_STAGES=3 still builds
// it); the deeper ring measured +0.4% at m = 4096.  Measured dead ends of round 2 (profiles/r02_gram_kernel_experiments.md):
// running the two warps of an SM sub-partition 1..4 slabs out of phase so that one's epilogue overlaps the other's main loop
// (-1.2 ... -2.2%: one warp alone cannot keep the FP64 tensor pipe fed, the lock-step pair can), and handing the reduce-add +
// completion signal of a finished Gram item to helper lanes of the producer warpgroup (-1.3%: the consumers' epilogue shrank
// by 1.6 k clk per item but their main loop grew by 2.7 k).
#ifndef NK_GRAM_STAGES
#define NK_GRAM_STAGES 5
#endif
constexpr int kGramStages = NK_GRAM_STAGES;
constexpr int kStagingPerWarp = kGramStages > 3 ? 8192 : 16384;
constexpr int kStagingHalves = 16384 / kStagingPerWarp;
constexpr int kItemQueue = 4;
constexpr size_t kGramStageBytes = (size_t)kGramStages * 2 * kSlabTileDoubles * 8;      // 96 KB operand ring
constexpr size_t kGramStagingBytes = (size_t)kConsumerWarps * kStagingPerWarp;           // accumulator / exponent staging
constexpr size_t kGramCtlBytes = 1024;
constexpr size_t kGramSmemBytes = kGramStageBytes + kGramStagingBytes + kGramCtlBytes;

void launch_gram(const GramParams &P, int sm_count, cudaStream_t stream, cudaError_t *err);

}  // namespace nk
