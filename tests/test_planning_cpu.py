"""CPU tests of the launch planning arithmetic. (1) The stream-K work decomposition (csrc/kernels_umma.cu: streamk_split / decode_work / sk_first_work / sk_next_work),
evaluated on the host through snnb_debug_streamk_schedule with the very functions the kernel's roles call: every K block of every
tile is computed exactly once, a cut tile's pieces are numbered 0 .. pieces-1 in K order with pieces <= 4 (the reducer's limit), every
CTA walks its pieces before its whole tiles, and the cut ranges are balanced to one K block."""
import ctypes as C

import numpy as np
import pytest

from shadernn_b200._lib import lib


def schedule(tiles, num_kb, sms):
    cap = tiles * 6 + 8 * sms + 16
    rows = np.zeros((cap, 6), np.int32)
    n = lib().snnb_debug_streamk_schedule(tiles, num_kb, sms, rows.ctypes.data_as(C.POINTER(C.c_int)), cap)
    assert n <= cap
    return None if n < 0 else rows[:n]


@pytest.mark.parametrize("tiles,num_kb,sms", [
    (224, 18, 148),   # ResNet-18 28x28x128: one whole wave + 76 tiles
    (128, 36, 148),   # 14x14x256: fewer tiles than SMs
    (52, 72, 148),    # 7x7x512
    (149, 9, 148), (295, 9, 148), (1, 72, 148), (3, 6, 148), (147, 7, 148), (1000, 11, 148), (17, 100, 16), (5, 6, 4), (443, 27, 132),
])
def test_streamk_schedule_covers_every_k_block_once(tiles, num_kb, sms):
    rows = schedule(tiles, num_kb, sms)
    assert rows is not None
    cover = np.zeros((tiles, num_kb), np.int32)
    per_tile = {}
    cut_units = {}
    seen_whole = set()
    for cta, tile, kb0, kb1, piece, pieces in rows.tolist():
        assert 0 <= cta < sms and 0 <= tile < tiles and 0 <= kb0 < kb1 <= num_kb, (cta, tile, kb0, kb1)
        cover[tile, kb0:kb1] += 1
        assert 1 <= pieces <= 4 and 0 <= piece < pieces
        per_tile.setdefault(tile, []).append((piece, pieces, kb0, kb1))
        if pieces == 1 and kb1 - kb0 == num_kb and tile < (tiles // sms) * sms:
            seen_whole.add(cta)
        else:
            assert cta not in seen_whole, "CTA %d took a piece after a whole tile" % cta  # pieces first
            cut_units[cta] = cut_units.get(cta, 0) + kb1 - kb0
    assert (cover == 1).all(), "K blocks computed %s times" % sorted(set(cover.ravel().tolist()))
    for tile, ps in per_tile.items():
        ps.sort()
        assert [q[0] for q in ps] == list(range(ps[0][1])) and all(q[1] == ps[0][1] for q in ps), (tile, ps)
        assert ps[0][2] == 0 and ps[-1][3] == num_kb and all(a[3] == b[2] for a, b in zip(ps, ps[1:])), (tile, ps)  # K order = piece order
    if cut_units:
        assert max(cut_units.values()) - min(cut_units.values()) <= 1, sorted(set(cut_units.values()))
        assert min(cut_units.values()) * 3 >= num_kb  # >= a third of a tile each: what bounds pieces at 4


def test_streamk_declines_when_there_is_nothing_to_cut():
    assert schedule(296, 18, 148) is None   # whole waves only
    assert lib().snnb_debug_streamk_schedule(0, 18, 148, None, 0) == -1


@pytest.mark.parametrize("k", [2, 3, 4, 5, 6, 7, 8, 9])
@pytest.mark.parametrize("pad_x", [0, 1, 2, 3, 4])
@pytest.mark.parametrize("ic", [1, 3, 4])
def test_feed_plan_invariants(k, pad_x, ic):
    # FeedPlan (csrc/snnb_internal.h): the window of output pixel m starts at pixel 2 m of the compact copy, whose image is shifted right by
    # px; tap t sits at window pixel d + t. What the kernel and the packer rely on:
    out = (C.c_int * 5)()
    ok = lib().snnb_debug_feed_plan(k, 2, pad_x, ic, out)
    px, d, nch, ksteps, rpp = list(out)
    if not ok:
        assert (k + (pad_x & 1) + 1) // 2 > 8  # only windows that need more than four K steps are refused for stride 2, ic <= 4
        return
    assert px % 2 == 0 and px >= pad_x and d == px - pad_x and d in (0, 1)   # 16-byte chunks start on even pixels
    assert nch % 2 == 0 and 2 * nch >= d + k and ksteps == nch // 2 and 1 <= ksteps <= 4   # every tap inside the window, whole K steps
    assert 2 * (nch - 2) < d + k + 2                                        # and no K step more than needed
    assert rpp * 16 * ksteps <= 64 and rpp in (1, 2, 4)                      # filter rows sharing a 128-byte weight row fit its 64 K columns
    # window pixel -> tap, as pack_feed_host maps it: every tap exactly once, the rest padding
    taps = [off - d for off in range(2 * nch) if 0 <= off - d < k]
    assert taps == list(range(k))


def test_feed_plan_refuses_other_layers():
    out = (C.c_int * 5)()
    assert lib().snnb_debug_feed_plan(7, 1, 3, 3, out) == 0    # stride 1: windows start on odd pixels
    assert lib().snnb_debug_feed_plan(7, 2, 3, 8, out) == 0    # more than 4 channels per pixel
    assert lib().snnb_debug_feed_plan(1, 2, 0, 3, out) == 0    # 1x1
