"""CPU: the work distribution of the tensor-core GEMM (gemm_tc.cu: TileSched, WorkIter, plan_pair), enumerated on the host
by the kernel's own scheduling code through gadcu_debug_gemm_schedule — no GPU involved. For every shape: each k-block of
each output tile is computed exactly once; a tile is either one whole unit, two k halves (tile-pushing mode) or one
head + one tail that meet at the same k-block; the hand-over between head and tail cannot deadlock (clusters that compute
tails never wait); the flags buffer is large enough; the split never makes the busiest cluster busier."""
import ctypes as C

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from gad_b200 import _lib


def schedule(M, N, K, sm=148, push=0):
    lib = _lib.load()
    info = np.zeros(8, np.uint32)
    n = lib.gadcu_debug_gemm_schedule(sm, M, N, K, push, None, 0, info.ctypes.data_as(C.c_void_p))
    rows = np.zeros((max(n, 1), 6), np.uint32)
    n2 = lib.gadcu_debug_gemm_schedule(sm, M, N, K, push, rows.ctypes.data_as(C.c_void_p), n, info.ctypes.data_as(C.c_void_p))
    assert n2 == n
    keys = ["clusters", "splits", "rem", "tail_kb", "num_kb", "band", "along_n", "tiles"]
    return rows[:n].astype(np.int64), dict(zip(keys, (int(v) for v in info)))


def check(M, N, K, sm=148, push=0):
    rows, info = schedule(M, N, K, sm, push)
    num_mt, num_nt, KB = (M + 255) // 256, (N + 255) // 256, info["num_kb"]
    assert info["tiles"] == num_mt * num_nt and KB == (K + 31) // 32
    assert 1 <= info["clusters"] <= sm // 2
    assert np.all(rows[:, 0] < info["clusters"]) and np.all(rows[:, 1] < num_mt) and np.all(rows[:, 2] < num_nt)
    assert np.all(rows[:, 3] < rows[:, 4]) and np.all(rows[:, 4] <= KB)
    # coverage: every k-block of every tile exactly once
    cover = np.zeros((num_mt, num_nt, KB), np.int32)
    by_tile = {}
    for c, mt, nt, k0, k1, part in rows:
        cover[mt, nt, k0:k1] += 1
        by_tile.setdefault((mt, nt), []).append((c, k0, k1, part))
    assert cover.min() == 1 and cover.max() == 1, (M, N, K, sm, push, info)
    split_tiles = 0
    tail_clusters, head_clusters = set(), set()
    for items in by_tile.values():
        parts = sorted(p for _, _, _, p in items)
        if parts == [0]:
            assert items[0][1:3] == (0, KB)
        elif parts == [0, 0]:  # two k halves (tile-pushing mode, or stream-K switched off)
            assert info["splits"] == 2 and sorted(i[1:3] for i in items) == [(0, KB // 2), (KB // 2, KB)]
        else:
            assert parts == [1, 2], parts
            tail = next(i for i in items if i[3] == 1)
            head = next(i for i in items if i[3] == 2)
            assert head[1] == 0 and head[2] == tail[1] and tail[2] == KB and tail[2] - tail[1] == info["tail_kb"]
            assert head[0] != tail[0]
            tail_clusters.add(tail[0])
            head_clusters.add(head[0])
            split_tiles += 1
    assert split_tiles == info["rem"] and split_tiles * 8 <= 1024  # (gadcu_ctx::gemm_flags holds 1024 words)
    # hand-over cannot deadlock: a cluster that computes tails never waits for anybody (it has no head) ...
    assert not (tail_clusters & head_clusters)
    # ... and per cluster the whole units come first, then the parts (every cluster reaches the split wave together)
    load = np.zeros(info["clusters"], np.int64)
    for c in range(info["clusters"]):
        mine = rows[rows[:, 0] == c]
        parts = list(mine[:, 5])
        assert parts == sorted(parts, key=lambda p: p != 0), parts
        load[c] = int((mine[:, 4] - mine[:, 3]).sum())
    assert load.sum() == info["tiles"] * KB
    if info["rem"]:
        whole_waves = -(-info["tiles"] // info["clusters"])  # what the busiest cluster would carry without the split
        assert load.max() < whole_waves * KB
        assert load.max() <= 0.97 * whole_waves * KB + 1  # (it is only used where it gains >= 3 %)
    if push:
        assert info["rem"] == 0 and info["splits"] == push
    return info


@pytest.mark.parametrize("M,N,K,expect_split", [
    (8192, 4096, 4096, False),   # 512 tiles = 6.92 waves: < 3 % to gain
    (4096, 4096, 8192, True),    # the dW GEMM of the 1-GPU step: 256 tiles, 3 waves + 34 split
    (4096, 4096, 4096, True),    # 2-GPU shards
    (2048, 4096, 4096, True),    # 4-GPU shards: 74 + 54
    (1024, 4096, 4096, False),   # 8-GPU shards: a lone wide wave (measured: no gain, hurts the overlapped exchange)
    (1024, 1024, 1024, True),    # 16 tiles on 32 clusters: every tile in two parts
    (256, 256, 64, False),       # two k-blocks: nothing to split
    (18944, 256, 4096, False),   # exactly one full wave
])
def test_named_shapes(M, N, K, expect_split):
    info = check(M, N, K)
    assert bool(info["rem"]) == expect_split, info


def test_tile_pushing_mode_keeps_whole_tiles_or_halves():
    for M, N, K, push in [(4096, 4096, 4096, 2), (4096, 4096, 1024, 1), (4096, 4096, 2048, 2)]:
        check(M, N, K, push=push)


@settings(max_examples=300, deadline=None)
@given(M=st.integers(64, 12000), N=st.integers(64, 12000), K=st.integers(32, 9000), sm=st.sampled_from([148, 132, 108, 16, 2]))
def test_random_shapes(M, N, K, sm):
    if ((M + 255) // 256) * ((N + 255) // 256) * ((K + 31) // 32) > 400_000:
        K = 1024
    check(M, N, K, sm)


def test_band_order_visits_every_tile_once_for_every_band_width():
    """TileSched::decode is a bijection whatever the band (including bands that do not divide the tile grid)."""
    import os
    for M, N in [(256 * 7, 256 * 5), (256 * 3, 256 * 11), (256, 256 * 40), (256 * 33, 256)]:
        rows, info = schedule(M, N, 64)
        tiles = {(mt, nt) for _, mt, nt, *_ in rows}
        assert len(tiles) == info["tiles"] == ((M + 255) // 256) * ((N + 255) // 256)
