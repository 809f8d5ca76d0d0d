"""The paint kernel's algorithm (futspace_b200/csrc/fsb_paint.cu) restated in Python and checked on the CPU against the
plain statement of what it computes (fut/voxel_renderer.fut:244-251 -- scatter, fill_vline scan, sky,
transpose).  The model keeps the kernel's structure: one warp per group of 32 columns (or per segment of bands), lists
walked backward, a ring of 32 colours per lane, the rows at which they start as two 32-bit masks (band being painted,
next band), colour trips taken by every lane that has room and a candidate in reach, bands drained row by row.  It
asserts the invariants the kernel relies on: a lane that needs a trip always has room, the ring never holds more than 32
entries, every record is filtered exactly once per segment that paints it.  No GPU needed."""
import numpy as np
import pytest

RING, RUNAHEAD = 32, 32


def reference_columns(lists, h, sky, empty):
    """lists[j] = [(row, colour), ...] with rows strictly decreasing -> [h][len(lists)] pixels"""
    out = np.empty((h, len(lists)), np.uint32)
    for j, lst in enumerate(lists):
        at = {r: c for r, c in lst}
        cur = sky
        for r in range(h):
            if r in at and at[r] != empty:
                cur = at[r]
            out[r, j] = cur
    return out


def paint_group(lists, h, sky, empty, b_first, b_end, out, col0, stats):
    """one warp: 32 columns lists[0..31] (short groups padded with empty lists), bands [b_first, b_end)"""
    n = [len(l) for l in lists]
    p = [k - 1 for k in n]
    cur = [sky] * 32
    if b_first > 0:                                   # bisection + the running colour that enters the segment
        r_first = b_first * 32
        for l in range(32):
            lo = sum(1 for r, _ in lists[l] if r >= r_first)
            p[l] = lo - 1
            for q in range(lo, n[l]):
                stats["carry"] += 1
                if lists[l][q][1] != empty:
                    cur[l] = lists[l][q][1]
                    break
    ring = [[] for _ in range(32)]                    # FIFO of colours; capacity checked
    mask0, mask1 = [0] * 32, [0] * 32
    row_lim = b_end * 32
    for b in range(b_first, b_end):
        r_end = (b + 1) * 32
        r_ahead = min(r_end + RUNAHEAD, row_lim)
        while True:
            row1 = [lists[l][p[l]][0] if p[l] >= 0 else None for l in range(32)]
            need = [p[l] >= 0 and row1[l] < r_end for l in range(32)]
            if not any(need):
                break
            can = [p[l] >= 0 and len(ring[l]) < RING and row1[l] < r_ahead for l in range(32)]
            assert all(c for c, nd in zip(can, need) if nd), "a lane that needs the trip must have room"
            stats["trips"] += 1
            for l in range(32):
                if can[l]:
                    row, colour = lists[l][p[l]]
                    assert row >= b * 32, "records above the band were painted by earlier bands"
                    p[l] -= 1
                    ring[l].append(colour)
                    stats["filtered"] += 1
                    if row < r_end:
                        mask0[l] |= 1 << (row & 31)
                    else:
                        mask1[l] |= 1 << (row & 31)
            assert max(len(r) for r in ring) <= RING
        for r in range(min(32, h - b * 32)):
            for l in range(32):
                if mask0[l] >> r & 1:
                    e = ring[l].pop(0)
                    if e != empty:
                        cur[l] = e
                if col0 + l < out.shape[1]:
                    out[b * 32 + r, col0 + l] = cur[l]
        for l in range(32):
            assert bin(mask0[l]).count("1") <= 32
            mask0[l], mask1[l] = mask1[l], 0
    return out


def random_lists(rng, ncols, h, density, empty, sky):
    lists = []
    for _ in range(ncols):
        k = int(rng.binomial(h, min(1.0, density * rng.uniform(0.2, 1.8))))
        rows = np.sort(rng.choice(h, size=min(k, h), replace=False))[::-1]
        cols = rng.integers(1, 1 << 32, size=rows.size, dtype=np.uint64).astype(np.uint32)
        cols[rng.random(rows.size) < 0.1] = empty     # transparent records
        cols[rng.random(rows.size) < 0.05] = sky
        lists.append([(int(r), int(c)) for r, c in zip(rows, cols)])
    return lists


@pytest.mark.parametrize("h,ncols,density,seg_bands", [(1, 1, 1.0, 0), (31, 5, 0.5, 0), (32, 32, 1.0, 0), (33, 40, 0.9, 1),
                                                       (200, 70, 0.12, 0), (200, 70, 0.12, 2), (257, 33, 0.6, 3),
                                                       (1080, 64, 0.12, 0), (1080, 64, 0.12, 9), (700, 32, 1.0, 5)])
@pytest.mark.parametrize("sentinel", ["zero", "sky"])
def test_paint_model_matches_the_plain_statement(h, ncols, density, seg_bands, sentinel):
    rng = np.random.default_rng(h * 1000 + ncols + seg_bands)
    sky = 0xFF9090E0
    empty = 0 if sentinel == "zero" else sky
    lists = random_lists(rng, ncols, h, density, empty, sky)
    want = reference_columns(lists, h, sky, empty)
    n_bands = (h + 31) // 32
    out = np.full((h, ncols), 0xDEADBEEF, np.uint32)
    stats = {"trips": 0, "filtered": 0, "carry": 0}
    segs = [(0, n_bands)] if seg_bands == 0 else [(b, min(n_bands, b + seg_bands)) for b in range(0, n_bands, seg_bands)]
    for g in range((ncols + 31) // 32):
        group = lists[g * 32:(g + 1) * 32]
        group = group + [[] for _ in range(32 - len(group))]
        for b0, b1 in segs:
            paint_group(group, h, sky, empty, b0, b1, out, g * 32, stats)
    assert np.array_equal(out, want)
    records = sum(len(l) for l in lists)
    assert stats["filtered"] == records               # every record is filtered once by the segment that paints it
    if records:
        assert records <= 32 * stats["trips"]
