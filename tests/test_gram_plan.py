"""Work plan of the fused lift+Gram engine (nk_gram_plan: the same host function nk_gram_begin uses), checked on the CPU.

The persistent kernel has no grid-wide barrier: CTAs claim items in ONE global order and spin on counters for their dependences.
That cannot deadlock iff every item depends only on items EARLIER in the claim order (all CTAs are resident, a blocked CTA holds
exactly one item).  This test rebuilds the claim order exactly as the kernel's producer does and checks that invariant, the
per-chunk item counts, the accumulator-tile map and the chunks-in-flight rule -- for the script shapes, the headline shapes and
ragged ones."""
import ctypes as C

import pytest

PACK, LIFT, GRAM = 0, 1, 2
SHAPES = [  # m, d, p, chunk
    (4096, 192, 6, 0), (8192, 192, 6, 0), (20, 2, 1, 0), (100, 1, 1, 0), (100, 192, 6, 0), (398, 192, 6, 256), (1, 1, 0, 128),
    (129, 7, 0, 128), (1000, 4, 1, 128), (513, 33, 3, 1024), (260, 120, 128, 0),
]


def plan(m, d, p, chunk, sms=148):
    from nys_koop_lqr_b200 import _lib
    lib = _lib.load()
    summ = (C.c_int * 12)()
    n = lib.nk_gram_plan(m, d, p, chunk, sms, summ, None, 0)
    assert n > 0
    items = (C.c_int * (4 * n))()
    assert lib.nk_gram_plan(m, d, p, chunk, sms, summ, items, n) == n
    keys = ("chunk", "MP", "KLS", "EP", "psi_rows", "nblk", "ntiles", "n_pk", "n_lf", "n_sy", "period_len", "nslots")
    return dict(zip(keys, summ)), [tuple(items[4 * i:4 * i + 4]) for i in range(n)]


def claim_order(s, period, n_chunks):
    """The producer's claim(): idx -> period idx / period_len - 1; Gram items belong to that chunk, pack / lift items to the next."""
    out = []
    for idx in range((n_chunks + 1) * s["period_len"]):
        per = idx // s["period_len"] - 1
        t, a, b, c = period[idx % s["period_len"]]
        chunk = per if t == GRAM else per + 1
        if 0 <= chunk < n_chunks:
            out.append((t, chunk, a, b, c))
    return out


@pytest.mark.parametrize("m,d,p,chunk", SHAPES)
def test_every_item_depends_only_on_earlier_items(m, d, p, chunk):
    s, period = plan(m, d, p, chunk)
    assert len(period) == s["period_len"] == s["n_pk"] + s["n_lf"] + s["n_sy"]
    S = s["nslots"]
    n_chunks = 2 * S + 3
    order = claim_order(s, period, n_chunks)
    assert len(order) == n_chunks * s["period_len"] and len(set(order)) == len(order)
    first, last = {}, {}
    gram_pos = {}
    for pos, (t, ch, a, b, c) in enumerate(order):
        first.setdefault((t, ch), pos)
        last[(t, ch)] = pos
        if t == GRAM:
            gram_pos[(ch, c)] = pos
    for ch in range(n_chunks):
        cnt = {t: sum(1 for it in order if it[0] == t and it[1] == ch) for t in (PACK, LIFT, GRAM)}
        assert cnt == {PACK: s["n_pk"], LIFT: s["n_lf"], GRAM: s["n_sy"]}
        assert last[(PACK, ch)] < first[(LIFT, ch)]                      # lift(c) waits for all packs of chunk c
        assert last[(LIFT, ch)] < first[(GRAM, ch)]                      # Gram items of chunk c wait for all its lift items
        if ch >= S:
            assert last[(GRAM, ch - S)] < first[(PACK, ch)]              # pack(c) reuses the buffers of chunk c - S
        if ch >= 1:
            for tile in range(s["ntiles"]):
                assert gram_pos[(ch - 1, tile)] < gram_pos[(ch, tile)]   # per-tile chunk order (version counter, determinism)


@pytest.mark.parametrize("m,d,p,chunk", SHAPES)
def test_tiles_and_item_coverage(m, d, p, chunk):
    s, period = plan(m, d, p, chunk)
    MB, EB = s["MP"] // 128, s["EP"] // 128
    assert s["MP"] >= m and s["MP"] - m < 128 and s["EP"] >= p + d and s["EP"] - (p + d) < 128 and s["KLS"] * 16 >= d + 2
    assert s["nblk"] == 2 * MB + EB and s["psi_rows"] == 2 * s["MP"] + s["EP"] and s["chunk"] % 128 == 0
    grams = [(a, b, c) for t, a, b, c in period if t == GRAM]
    assert sorted(c for _, _, c in grams) == list(range(s["ntiles"]))
    have = {(a, b) for a, b, _ in grams}
    assert len(have) == len(grams)
    need = {(i, j) for i in range(2 * MB) for j in range(i + 1)}                         # Phi Phi^T blocks, lower triangle
    u_blocks = (p + 127) // 128
    for e in range(EB):
        need |= {(2 * MB + e, MB + j) for j in range(MB)}                                # [U;Y] x Phi_y  (G_yu, G_Yy)
        if e < u_blocks:
            need |= {(2 * MB + e, j) for j in range(MB)}                                 # U x Phi_x      (G_xu)
    need |= {(2 * MB + e, 2 * MB + f) for e in range(u_blocks) for f in range(e + 1)}    # U U^T
    assert have == need
    strips = s["chunk"] // 128
    assert sorted(a for t, a, _, _ in period if t == PACK) == list(range(strips))
    assert sorted((a, b, c) for t, a, b, c in period if t == LIFT) == sorted((side, lb, sb) for side in (0, 1) for lb in range(MB) for sb in range(strips))


def test_chunks_in_flight_rule():
    for m, d, p, chunk in SHAPES:
        for sms in (1, 16, 148, 160):
            s, _ = plan(m, d, p, chunk, sms)
            assert 2 <= s["nslots"] <= 16
            if s["period_len"] >= 2 * sms:
                assert s["nslots"] == 2
            else:
                assert s["nslots"] == min(16, max(2, -(-3 * sms // s["period_len"])))
    assert plan(4096, 192, 6, 0)[0]["nslots"] == 2 and plan(8192, 192, 6, 0)[0]["nslots"] == 2     # headline shapes: two buffers
    assert plan(20, 2, 1, 0)[0]["nslots"] == 16                                                    # Duffing script shape
