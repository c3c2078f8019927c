// nk_gram.cu -- the fused kernel-lift + Gram engine (reference regressors.py:141-142,147,151,153,162,164).
//
// What the reference does with three materialised m x n matrices (K_mn_out, K_mn_in_x, K_mn_in) and four
// dgemm calls is done here by ONE persistent kernel per fit.  Stacking Psi = [Phi_x ; Phi_y ; U ; Y]
// (rows: features of x_t, features of x_{t+1}, controls, next states) every Gram the fit needs is a block of
// the symmetric product Psi Psi^T, so the work is a rank-n update of the lower triangle of a (2m+p+d)^2
// matrix whose operand is produced on the fly:
//
//   pack(c)  : one 128-sample strip of chunk c -> scaled/centred/augmented sample operands for the lift GEMM
//              ([x', -|x'|^2/2, 1] so that the accumulated value IS the exponent -r^2/2) and the raw [U;Y] rows.
//   lift(c)  : 128 landmarks x 128 samples tile:  DMMA GEMM over d+2, kernel function applied on chip (exponents
//              staged through shared memory, rolled loop), tile stored in C-fragment order straight into the packed
//              feature chunk (L2-resident, evict_last).
//   syrk(c)  : 128 x 128 output tile: DMMA contraction over the chunk's samples, accumulators added into the
//              fragment-ordered accumulator workspace with cp.reduce.async.bulk ... add.f64 (UBLKRED).
//
// The n x m feature matrix never exists: only a double-buffered chunk of nk samples (2 x 34.6 MB at m=4096,
// nk=512) lives in the 126 MB L2.  (With few landmarks -- the script configurations, m = 10 ... 400 -- one chunk's
// items cannot fill 148 SMs, so up to 16 chunks are kept in flight, chunk c in buffer slot c % S, and a Gram item
// signals its completion at once instead of after its next main loop: a Duffing fit, n = 69 900, m = 20, went from
// 8.0 to 1.6 ms.)  Work items are claimed in a fixed global order from one atomic counter;
// items of chunk c+1's pack/lift are spliced into the middle of chunk c's syrk items, and every cross-CTA
// dependence (pack->lift->syrk->buffer reuse, and chunk order per accumulator tile) is a monotone counter in
// global memory, so there is no grid-wide barrier and the summation order is fixed (deterministic results).
//
// CTA = 8 consumer warps (2 x 4, warp tile 64 x 32, 64 FP64 accumulators per lane) + 1 producer warp that
// claims items one ahead, resolves their dependences and feeds a 5-stage ring of 2 x 16 KB operand slabs with
// cp.async.bulk (TMA engine, mbarrier transaction counts); 8 KB of shared memory per consumer warp stage the
// accumulator tiles for the asynchronous bulk reduce-adds of the Gram epilogue and the exponents of the kernel-function
// epilogue (two passes each).  Measured dead ends (profiles/r02_gram_kernel_experiments.md): two 4-warp
// CTAs per SM as a ping-pong, the two halves of one CTA out of phase, helper lanes for the completion signal.
#include "nk_gram.cuh"
#include "nk_mainloop.cuh"

namespace nk {

struct __align__(16) QueuedItem { int type, chunk, a, b, c, slot, pad1, pad2; };   // slot = chunk % nslots (set by the producer)

struct GramSmemCtl {
    uint64_t full[kGramStages];
    uint64_t empty[kGramStages];
    uint64_t iq_full[kItemQueue];
    uint64_t iq_empty[kItemQueue];
    QueuedItem iq[kItemQueue];
};
static_assert(sizeof(GramSmemCtl) <= kGramCtlBytes, "control block does not fit its reservation");

// Dependence wait with a watchdog.  A wait that outlasts kSpinCap polls (each poll is an L2 round trip plus a 64 ns sleep, i.e.
// roughly a microsecond: the cap is tens of seconds, a legitimate wait lasts microseconds), or that sees another CTA's timeout,
// raises *err and returns false: the caller abandons its item, the kernel drains instead of hanging the device, and the host
// turns the flag into NK_E_STATE at its next synchronising call (nk_gram_status, nk_solve_abc*, nk_cv_weights).
__device__ __forceinline__ bool spin_until_ge(const int *ctr, int target, int *err) {
    unsigned it = 0;
    while (ld_acquire(ctr) < target) {
        __nanosleep(64);
        if ((++it & 0xffffu) == 0 && (it >= kSpinCap || ld_acquire(err) != 0)) { atomicExch(err, 1); return false; }
    }
    return true;
}

// ------------------------------------------------------------------------------------------------
// pack item: 128 samples -> XP, YP (lift operands) and the [U;Y] rows of Psi
// ------------------------------------------------------------------------------------------------
__device__ void do_pack(const GramParams &P, int chunk, int slot, int sb, int tid) {
    const int warp = tid >> 5, lane = tid & 31;
    const long long s_base = (long long)chunk * P.nk + sb * kTile;
    const int KL = P.KLS * kSlabK;
    const int rp = P.nk / kPanel;  // row panels of XP / YP
    const uint64_t pol = policy_evict_last();
    // (1) scaled, centred, augmented operands.  one warp per sample row.
    for (int side = 0; side < 2; side++) {
        const double *src = side ? P.Y : P.X;
        const long long ld = side ? P.ldy : P.ldx;
        double *dst = side ? P.YP[slot] : P.XP[slot];
        for (int r = warp; r < kTile; r += kConsumerWarps) {
            const long long s = s_base + r;
            const int row = sb * kTile + r;
            double nrm = 0.0;
            for (int k = lane; k < KL; k += 32) {
                double v = 0.0;
                if (k < P.d && s < P.n) {
                    v = (src[s * ld + k] - P.center[k]) * P.inv_ls[k];
                    nrm += v * v;
                }
                if (k < P.d || k >= P.d + 2) dst[packed_off(row, k, rp)] = v;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) nrm += __shfl_xor_sync(0xffffffffu, nrm, o);
            if (lane == 0) {
                const bool live = s < P.n;
                dst[packed_off(row, P.d, rp)] = live ? -0.5 * nrm : 0.0;
                dst[packed_off(row, P.d + 1, rp)] = live ? 1.0 : 0.0;
            }
        }
    }
    // (2) raw [U ; Y] rows of Psi for these 128 samples (row e < p: control e; p <= e < p+d: next-state e-p)
    double *psi = P.PSI[slot];
    for (int e = tid; e < P.EP; e += kConsumerWarps * 32) {
        const double *src = nullptr; long long ld = 0; int col = 0;
        if (e < P.p) { src = P.X; ld = P.ldx; col = P.d + e; }
        else if (e < P.p + P.d) { src = P.Y; ld = P.ldy; col = e - P.p; }
        const int R = P.e_row0 + e;
        for (int s8 = 0; s8 < kTile / 8; s8++) {
            double v[8];
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const long long s = s_base + s8 * 8 + q;
                v[q] = (src != nullptr && s < P.n) ? src[s * ld + col] : 0.0;
            }
            double *o = psi + packed_off(R, sb * kTile + s8 * 8, P.psi_rp);
#pragma unroll
            for (int q = 0; q < 8; q += 2) st_v2_hint(o + q, v[q], v[q + 1], pol);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// kernel-function epilogue of a lift item: the warp's 64x32 tile of exponents (staged in shared memory, fragment order) ->
// kernel values -> packed feature chunk.  One fragment ROW block (8 landmarks x 32 samples: 4 column blocks, 8 values per lane)
// per iteration, evaluated as straight-line branch-free code so that the eight independent polynomial chains interleave: the
// first version went through the library exp one value at a time (a 16-deep dependent FP64 chain, 320 clk per value: the
// epilogue cost 24 k clk per item against 53 k clk for the item's main loop).  Masking (padded landmarks, samples past n) is a
// select after the evaluation.  Rolled over the 8 row blocks (fully unrolled it thrashed the instruction cache).
// ------------------------------------------------------------------------------------------------
template <int KIND>
__device__ __forceinline__ void lift_epilogue(const GramParams &P, uint32_t stg, int lane, int t, double *psi, long long s_chunk, int lm0,
                                              int sl0, int R0, uint64_t pol, int i_begin, int i_end) {
#pragma unroll 1
    for (int i = i_begin; i < i_end; i++) {
        double2 e[4];
#pragma unroll
        for (int j = 0; j < 4; j++) e[j] = lds_v2(stg + (uint32_t)((i - i_begin) * 4 + j) * 512u + lane * 16u);
        double v[4][2];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            v[j][0] = kernel_from_exponent_t<KIND>(e[j].x);
            v[j][1] = kernel_from_exponent_t<KIND>(e[j].y);
        }
        const bool lm_ok = (lm0 + i * 8) < P.m;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int s_local = sl0 + j * 8;
            const long long s0 = s_chunk + s_local + 2 * t;
            const double v0 = (lm_ok && s0 < P.n) ? v[j][0] : 0.0;
            const double v1 = (lm_ok && (s0 + 1) < P.n) ? v[j][1] : 0.0;
            double *o = psi + packed_off(R0 + i * 8, s_local, P.psi_rp) + lane * 2;
            st_v2_hint(o, v0, v1, pol);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// the persistent kernel
// ------------------------------------------------------------------------------------------------
// kMulti = false: two chunk buffers (slot = chunk parity), deferred completion signal -- the large-m configuration, kept as
// its own instantiation so that its instruction schedule does not depend on the small-problem generalisation (the merged
// runtime-S version measured 0.7% slower at m=4096 on the same box).  kMulti = true: S = P.nslots buffers, immediate signal.
template <bool kMulti>
__global__ void __launch_bounds__(kThreads, 1) gram_kernel(const GramParams P) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    double *stage_base = reinterpret_cast<double *>(smem_raw);
    GramSmemCtl *ctl = reinterpret_cast<GramSmemCtl *>(smem_raw + kGramStageBytes + kGramStagingBytes);

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        for (int s = 0; s < kGramStages; s++) { mbar_init(&ctl->full[s], 1); mbar_init(&ctl->empty[s], kConsumerWarps); }
        for (int s = 0; s < kItemQueue; s++) { mbar_init(&ctl->iq_full[s], 1); mbar_init(&ctl->iq_empty[s], kConsumerWarps); }
        fence_mbar_init();
    }
    __syncthreads();

    const int total_items = (P.n_chunks + 1) * P.period_len;
    const int slabs_syrk = P.nk / kSlabK;
    const int lift_warps_per_chunk = P.n_lf * kConsumerWarps;
    const int syrk_warps_per_chunk = P.n_sy * kConsumerWarps;

    if (warp >= kConsumerWarps) {
        // =============================== producer warpgroup ===============================
        setmaxnreg_dec<40>();
        if (warp == kConsumerWarps && lane == 0) {
            const uint64_t pol_keep = policy_evict_last();
            uint32_t stage = 0, sphase = 0;   // operand ring
            uint32_t qslot = 0, qphase = 0;   // item queue
            // claim(): next existing item in the global order (items of chunks that do not exist are skipped)
            auto claim = [&](QueuedItem &it) -> bool {
                for (;;) {
                    const int idx = atomicAdd(&P.counters[0], 1);
                    if (idx >= total_items) { it.type = -1; return false; }
                    const int period = idx / P.period_len - 1;
                    const GramItem g = P.items[idx % P.period_len];
                    const int chunk = (g.type == kItemSyrk) ? period : period + 1;
                    if (chunk < 0 || chunk >= P.n_chunks) continue;
                    it.type = g.type; it.chunk = chunk; it.a = g.a; it.b = g.b; it.c = g.c;
                    return true;
                }
            };
            QueuedItem it, nxt;
            it.slot = it.pad1 = it.pad2 = 0; it.chunk = it.a = it.b = it.c = 0;
            nxt = it;
            bool valid = claim(it);
            for (;;) {
                if (valid) {
                    // ---- dependences (all on items claimed earlier in the global order) ----
                    // Counters are split by buffer slot (chunk c lives in slot c % S, S = 2 for large problems):
                    // completions of chunk c+S can only start after everything of chunk c has finished (pack(c+S) waits
                    // for syrk(c)), so a per-slot count reaching its target means exactly "all items of chunks c, c-S, ...
                    // are done".  One running total would let early finishers of a later chunk stand in for a straggler
                    // of this one on small problems.
                    const int par = kMulti ? it.chunk % P.nslots : (it.chunk & 1);
                    const int gen = kMulti ? it.chunk / P.nslots : (it.chunk >> 1);
                    it.slot = par;
                    bool ok = true;
                    if (it.type == kItemPack) {
                        // buffers of this slot were last read by lift(chunk-S) / syrk(chunk-S)
                        if (gen >= 1) ok = spin_until_ge(&P.counters[kCtrSyrk + par], gen * syrk_warps_per_chunk, P.err);
                    } else if (it.type == kItemLift) {
                        ok = spin_until_ge(&P.counters[kCtrPack + par], (gen + 1) * P.n_pk, P.err);
                    } else {
                        ok = spin_until_ge(&P.counters[kCtrLift + par], (gen + 1) * lift_warps_per_chunk, P.err);
                    }
                    if (!ok) { valid = false; it.type = -1; }      // watchdog: stop this CTA (the flag stops the others)
                    fence_proxy_async();
                }
                // ---- publish to the consumer warps ----
                mbar_wait(&ctl->iq_empty[qslot], qphase ^ 1);
                ctl->iq[qslot] = it;
                mbar_arrive(&ctl->iq_full[qslot]);   // release: consumers acquire through the wait
                if (++qslot == kItemQueue) { qslot = 0; qphase ^= 1; }
                if (!valid) break;
                // claim the following item now, so that its round trip overlaps with feeding this one
                const bool nvalid = claim(nxt);
                // ---- feed operand slabs ----
                if (it.type != kItemPack) {
                    const double *Abase, *Bbase; size_t a_stride, b_stride; int nslabs;
                    const int slot = kMulti ? it.slot : (it.chunk & 1);
                    if (it.type == kItemLift) {
                        Abase = (it.a ? P.ZP : P.ZPx) + (size_t)it.b * 16 * 128; a_stride = (size_t)(P.MP / kPanel) * 128;
                        Bbase = (it.a ? P.YP[slot] : P.XP[slot]) + (size_t)it.c * 16 * 128; b_stride = (size_t)(P.nk / kPanel) * 128;
                        nslabs = P.KLS;
                    } else {
                        Abase = P.PSI[slot] + (size_t)it.a * 16 * 128; a_stride = (size_t)P.psi_rp * 128;
                        Bbase = P.PSI[slot] + (size_t)it.b * 16 * 128; b_stride = a_stride;
                        nslabs = slabs_syrk;
                    }
                    for (int s = 0; s < nslabs; s++) {
                        mbar_wait(&ctl->empty[stage], sphase ^ 1);
                        double *As = stage_base + (size_t)stage * 2 * kSlabTileDoubles;
                        mbar_arrive_expect_tx(&ctl->full[stage], 2 * kSlabTileDoubles * 8);
                        bulk_g2s(As, Abase + s * a_stride, kSlabTileDoubles * 8, &ctl->full[stage], pol_keep);
                        bulk_g2s(As + kSlabTileDoubles, Bbase + s * b_stride, kSlabTileDoubles * 8, &ctl->full[stage], pol_keep);
                        if (++stage == kGramStages) { stage = 0; sphase ^= 1; }
                    }
                }
                it = nxt; valid = nvalid;
            }
        }
    } else {
        // =============================== consumer warps ===============================
        setmaxnreg_inc<232>();
        const int wr = warp >> 2, wc = warp & 3;
        const int g = lane >> 2, t = lane & 3;
        uint32_t stage = 0, sphase = 0, qslot = 0, qphase = 0;
        const uint64_t pol_keep = policy_evict_last();
        // shared-window addresses, computed once
        const uint32_t sm0 = opaque(smem_u32(smem_raw));
        const uint32_t a_off = opaque(sm0 + (uint32_t)(wr * 8 * 2) * 512u + lane * 16u);            // A fragment (i=0,q=0) of stage 0
        const uint32_t b_off = opaque(sm0 + 16384u + (uint32_t)(wc * 4 * 2) * 512u + lane * 16u);  // B fragment (j=0,q=0) of stage 0
        const uint32_t full0 = opaque(smem_u32(&ctl->full[0])), empty0 = opaque(smem_u32(&ctl->empty[0]));
        const uint32_t stg = opaque(sm0 + (uint32_t)kGramStageBytes + (uint32_t)warp * (uint32_t)kStagingPerWarp);  // this warp's staging buffer
        // deferred completion signal of the previous syrk item of this warp (its bulk reduce is asynchronous)
        int *pend_ver = nullptr;
        int pend_par = 0;
        auto flush_pending = [&]() {
            if (pend_ver != nullptr) {
                if (lane == 0) {
                    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
                    __threadfence();
                    atomicAdd(pend_ver, 1);
                    atomicAdd(&P.counters[kCtrSyrk + pend_par], 1);
                }
                __syncwarp();
                pend_ver = nullptr;
            }
        };
#ifdef NK_GRAM_TIMING
        long long tm[16];
#pragma unroll
        for (int i = 0; i < 16; i++) tm[i] = 0;
        const long long tm_begin = clock64();
#define TM_START(v) long long v = clock64()
#define TM_ADD(slot, v) tm[slot] += clock64() - (v)
#define TM_COUNT(slot) tm[slot]++
#else
#define TM_START(v)
#define TM_ADD(slot, v)
#define TM_COUNT(slot)
#endif
        for (;;) {
            TM_START(t_iq);
            if (!mbar_try_wait(&ctl->iq_full[qslot], qphase)) {
                flush_pending();      // never sit on a completion signal while idle: later items may be waiting for it
                mbar_wait(&ctl->iq_full[qslot], qphase);
            }
            TM_ADD(1, t_iq);
            const QueuedItem it = ctl->iq[qslot];
            __syncwarp();
            if (lane == 0) mbar_arrive(&ctl->iq_empty[qslot]);
            if (++qslot == kItemQueue) { qslot = 0; qphase ^= 1; }
            if (it.type < 0) { flush_pending(); break; }

            if (it.type == kItemPack) {
                TM_START(t_pk);
                flush_pending();
                do_pack(P, it.chunk, kMulti ? it.slot : (it.chunk & 1), it.a, tid);
                fence_proxy_async();   // generic-proxy stores are read back through the async proxy (bulk copies)
                __threadfence();
                named_bar_sync(1, kConsumerWarps * 32);
                if (tid == 0) atomicAdd(&P.counters[kCtrPack + (kMulti ? it.slot : (it.chunk & 1))], 1);
                TM_ADD(2, t_pk);
                continue;
            }

            double acc[8][4][2];
#pragma unroll
            for (int i = 0; i < 8; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }

            // ---- main loop: k8 steps with the first fragments of the NEXT step (and next slab) prefetched ----
            const int nslabs = (it.type == kItemLift) ? P.KLS : slabs_syrk;
            TM_START(t_ff);
            mbar_wait_a(full0 + stage * 8, sphase);
            TM_ADD(3, t_ff);
            TM_START(t_ml);
            double2 b[4], nb[4], a0, na0;
            {
                const uint32_t so = stage * (uint32_t)(2 * kSlabTileDoubles * 8);
#pragma unroll
                for (int j = 0; j < 4; j++) b[j] = lds_v2(b_off + so + j * 1024);
                a0 = lds_v2(a_off + so);
            }
            for (int s = 0; s < nslabs; s++) {
                const uint32_t so = stage * (uint32_t)(2 * kSlabTileDoubles * 8);
                uint32_t nstage = stage + 1, nphase = sphase;
                if (nstage == kGramStages) { nstage = 0; nphase ^= 1; }
                const bool has_next = (s + 1 < nslabs);
                // early, non-blocking look at the next slab's barrier (its result is consumed after the q=0 DMMAs)
                uint32_t ready = has_next ? mbar_test_a(full0 + nstage * 8, nphase) : 1u;
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
                    if (!ready) {
                        TM_START(t_mw);
                        mbar_wait_a(full0 + nstage * 8, nphase);
                        TM_ADD(11, t_mw);
                        TM_COUNT(12);
                    }
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

            TM_ADD((it.type == kItemLift) ? 4 : 5, t_ml);
            TM_COUNT((it.type == kItemLift) ? 9 : 10);
            TM_START(t_fl);
            flush_pending();
            TM_ADD(6, t_fl);
            TM_START(t_ep);
            if (it.type == kItemLift) {
                // The raw exponents go through this warp's staging buffer so that the kernel function runs in a ROLLED
                // loop: applied straight to the 64 accumulator registers the code must be fully unrolled (static register
                // indices), which with the double-precision exp / Matern bodies is ~240 KB of SASS that thrashes the
                // instruction cache once per item (measured: the epilogue then costs as much as the item's main loop).
                // Each lane reads back exactly what it wrote (no cross-lane traffic).
                const int slot = kMulti ? it.slot : (it.chunk & 1);
                double *psi = P.PSI[slot];
                const long long s_chunk = (long long)it.chunk * P.nk;
                const int lm0 = it.b * kTile + wr * 64 + g;                       // landmark of fragment row block i = 0
                const int sl0 = it.c * kTile + wc * 32;                           // first sample (chunk-local) of this warp tile
                const int R0 = it.a * P.MP + it.b * kTile + wr * 64;              // first feature row of this warp tile
                constexpr int kRowsPerPass = 8 / kStagingHalves;                  // fragment row blocks that fit the staging buffer
#pragma unroll
                for (int hh = 0; hh < kStagingHalves; hh++) {
                    if (hh) __syncwarp();
#pragma unroll
                    for (int i = 0; i < kRowsPerPass; i++)
#pragma unroll
                        for (int j = 0; j < 4; j++)
                            sts_v2(stg + (uint32_t)(i * 4 + j) * 512u + lane * 16u, acc[hh * kRowsPerPass + i][j][0], acc[hh * kRowsPerPass + i][j][1]);
                    __syncwarp();
                    if (P.kind == kRBF) lift_epilogue<kRBF>(P, stg, lane, t, psi, s_chunk, lm0, sl0, R0, pol_keep, hh * kRowsPerPass, (hh + 1) * kRowsPerPass);
                    else lift_epilogue<kMatern52>(P, stg, lane, t, psi, s_chunk, lm0, sl0, R0, pol_keep, hh * kRowsPerPass, (hh + 1) * kRowsPerPass);
                }
                TM_ADD(13, t_ep);          // (development build) the kernel-function loop alone
                TM_START(t_fn);
                fence_proxy_async();
                __threadfence();
                __syncwarp();
                if (lane == 0) atomicAdd(&P.counters[kCtrLift + slot], 1);
                TM_ADD(14, t_fn);          // ... and the fences + signal that publish the tile
                TM_ADD(7, t_ep);
            } else {
                // Gram epilogue: accumulators -> this warp's staging buffer (8 KB: two passes of four fragment row blocks) ->
                // asynchronous bulk reduce-adds (TMA engine, SASS UBLKRED.ADD.F64) into the fragment-ordered accumulator tile.
                // The warp does not wait for the reduction: the completion signal is sent when the next item's main loop is
                // over (flush_pending).  Chunk order per tile is enforced by the tile's version counter, so the summation
                // order is fixed.
                int *ver = &P.counters[kCounterTileVer + it.c];
                constexpr int kRowsPerPass = 8 / kStagingHalves;
                bool tile_ok = true;
#pragma unroll
                for (int hh = 0; hh < kStagingHalves; hh++) {
                    if (hh) {      // the engine must have READ the previous pass out of the staging buffer before it is overwritten
                        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                        __syncwarp();
                    }
#pragma unroll
                    for (int i = 0; i < kRowsPerPass; i++)
#pragma unroll
                        for (int j = 0; j < 4; j++)
                            sts_v2(stg + (uint32_t)(i * 4 + j) * 512u + lane * 16u, acc[hh * kRowsPerPass + i][j][0], acc[hh * kRowsPerPass + i][j][1]);
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) {
                        if (hh == 0) tile_ok = spin_until_ge(ver, it.chunk * kConsumerWarps, P.err);
                        if (tile_ok) {
                            double *gt = P.Gws + (size_t)it.c * (kTile * kTile) + (size_t)warp * 32 * kBlk + (size_t)hh * (kStagingPerWarp / 8);
                            asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], %2;"
                                         ::"l"(gt), "r"(stg), "r"(kStagingPerWarp) : "memory");
                        }
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                }
                pend_ver = ver;
                pend_par = kMulti ? it.slot : (it.chunk & 1);
                if (kMulti) flush_pending();   // small problems: the same tile of the next chunk (another CTA) is waiting for this
                TM_ADD(8, t_ep);
            }
        }
#ifdef NK_GRAM_TIMING
        tm[0] = clock64() - tm_begin;
        if (lane == 0 && P.timing != nullptr) {
            long long *o = P.timing + ((size_t)blockIdx.x * kConsumerWarps + warp) * 16;
#pragma unroll
            for (int i = 0; i < 16; i++) o[i] = tm[i];
        }
#endif
    }
}

void launch_gram(const GramParams &P, int sm_count, cudaStream_t stream, cudaError_t *err) {
    static unsigned long long configured = 0;
    if (first_use_on_device(configured)) {
        *err = cudaFuncSetAttribute(gram_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGramSmemBytes);
        if (*err != cudaSuccess) return;
        *err = cudaFuncSetAttribute(gram_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGramSmemBytes);
        if (*err != cudaSuccess) return;
    }
    const void *kernel = P.nslots > 2 ? (const void *)gram_kernel<true> : (const void *)gram_kernel<false>;
    // every CTA must be resident (items wait on items claimed earlier): cooperative launch guarantees it
    int max_blocks = 0;
    *err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&max_blocks, gram_kernel<false>, kThreads, kGramSmemBytes);
    if (*err != cudaSuccess) return;
    if (max_blocks < 1) { *err = cudaErrorLaunchOutOfResources; return; }
    long long total_items = (long long)(P.n_chunks + 1) * P.period_len;
    int grid = sm_count;
    if (total_items < grid) grid = (int)total_items;
    if (grid < 1) grid = 1;
    void *args[] = {(void *)&P};
    *err = cudaLaunchCooperativeKernel(kernel, dim3(grid), dim3(kThreads), args, kGramSmemBytes, stream);
}

// ------------------------------------------------------------------------------------------------
// landmark packing (once per fit): ZP rows = [z', 1, -|z'|^2/2], z' = (z - center) * inv_ls
// ------------------------------------------------------------------------------------------------
__global__ void pack_landmarks_kernel(const double *Z, long long ldz, int m, int d, int MP, int KLS,
                                      const double *inv_ls, const double *center, double *ZP) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= MP) return;
    const int KL = KLS * kSlabK, rp = MP / kPanel;
    double nrm = 0.0;
    for (int k = lane; k < KL; k += 32) {
        double v = 0.0;
        if (k < d && row < m) { v = (Z[(long long)row * ldz + k] - center[k]) * inv_ls[k]; nrm += v * v; }
        if (k < d || k >= d + 2) ZP[packed_off(row, k, rp)] = v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nrm += __shfl_xor_sync(0xffffffffu, nrm, o);
    if (lane == 0) {
        ZP[packed_off(row, d, rp)] = (row < m) ? 1.0 : 0.0;
        ZP[packed_off(row, d + 1, rp)] = (row < m) ? -0.5 * nrm : 0.0;
    }
}

void launch_pack_landmarks(const double *Z, long long ldz, int m, int d, int MP, int KLS, const double *inv_ls,
                           const double *center, double *ZP, cudaStream_t stream) {
    const int warps = 8;
    pack_landmarks_kernel<<<(MP + warps - 1) / warps, warps * 32, 0, stream>>>(Z, ldz, m, d, MP, KLS, inv_ls, center, ZP);
}

// ------------------------------------------------------------------------------------------------
// accumulator workspace -> row-major Grams
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double gws_read(const double *Gws, const int *tile_of, int nblk, int R, int C) {
    // element (R, C) of Psi Psi^T, R >= C in block order guaranteed by the caller
    const int I = R / kTile, J = C / kTile;
    const int tile = tile_of[I * nblk + J];
    if (tile < 0) return 0.0;
    const int r = R % kTile, c = C % kTile;
    const int w = (r >> 6) * 4 + (c >> 5);
    const int i = (r & 63) >> 3, j = (c & 31) >> 3;
    const int lane = (r & 7) * 4 + ((c & 7) >> 1);
    return Gws[(size_t)tile * (kTile * kTile) + (size_t)(w * 32 + i * 4 + j) * 64 + lane * 2 + (c & 1)];
}

// out(r,c) = Psi Psi^T (row0 + r, col0 + c); if the block (I,J) lies above the computed lower triangle the
// mirrored element is read (symmetry of Psi Psi^T).
__global__ void unpack_kernel(const double *Gws, const int *tile_of, int nblk, int row0, int col0, int rows, int cols,
                              double *out, long long ld, int accumulate) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (c >= cols || r >= rows) return;
    int R = row0 + r, C = col0 + c;
    if (R / kTile < C / kTile || (R / kTile == C / kTile && R < C)) { int tmp = R; R = C; C = tmp; }
    double v = gws_read(Gws, tile_of, nblk, R, C);
    double *o = out + (long long)r * ld + c;
    *o = accumulate ? (*o + v) : v;
}

void launch_unpack(const double *Gws, const int *tile_of, int nblk, int row0, int col0, int rows, int cols,
                   double *out, long long ld, int accumulate, cudaStream_t stream) {
    if (rows <= 0 || cols <= 0) return;
    dim3 block(128), grid((cols + 127) / 128, rows);
    unpack_kernel<<<grid, block, 0, stream>>>(Gws, tile_of, nblk, row0, col0, rows, cols, out, ld, accumulate);
}

}  // namespace nk
