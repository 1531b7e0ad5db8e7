"""The tile / z-chunk planner of the iteration kernels (bbpcg_plan_zchunks, csrc/bbpcg_solver.cu) as a pure host function:
whatever the block shape and the options, every plane 1..kn belongs to exactly one chunk, the chunk table is in claim order
with the block's top and bottom chunks first (the residual kernel pushes their planes to the z neighbours), and the item
count fits the reduction workspace.  A gap or an overlap here would silently skip or double-update planes on the GPU."""
import ctypes as C

import pytest
from hypothesis import given, settings, strategies as st

import bbpcg.lib as L

MAXZ, MAXBLOCKS, SLOTS = 1024, 65536, 296


def plan(in_, jn, kn, slots=SLOTS, ty=0, kc=0, guided=1, pct=0, cmin=0):
    lib = L.load_library()
    ztab = (C.c_int * (2 * MAXZ))()
    out = (C.c_int * 5)()
    rc = lib.bbpcg_plan_zchunks(in_, jn, kn, slots, ty, kc, guided, pct, cmin, ztab, 2 * MAXZ, out)
    if rc != 0:
        return rc, None, None
    ty_, nbx, nby, nbz, kc0 = list(out)
    return 0, (ty_, nbx, nby, nbz, kc0), [(ztab[2 * c], ztab[2 * c + 1]) for c in range(nbz)]


def check(in_, jn, kn, p, chunks):
    ty, nbx, nby, nbz, kc0 = p
    assert 1 <= ty <= 8 and nbx == -(-in_ // 128) and nby == -(-jn // ty) and nbz == len(chunks) >= 1
    seen = [0] * (kn + 2)
    for lo, hi in chunks:
        assert 1 <= lo <= hi <= kn
        for k in range(lo, hi + 1):
            seen[k] += 1
    assert seen[1:kn + 1] == [1] * kn                       # every plane exactly once
    assert chunks[0][1] - chunks[0][0] + 1 == kc0
    if nbz > 1:
        assert chunks[0][1] == kn and chunks[1][0] == 1     # boundary chunks are claimed first
        mids = chunks[1:]
        assert all(a[1] + 1 == b[0] for a, b in zip(mids, mids[1:])) and mids[-1][1] == chunks[0][0] - 1
    assert nbx * nby * nbz <= MAXBLOCKS


@settings(max_examples=300, deadline=None)
@given(st.integers(1, 700), st.integers(1, 700), st.integers(1, 1100), st.sampled_from([2, 64, 296, 2 * 132]),
       st.integers(0, 8), st.sampled_from([0, 0, 1, 7, 24, 1000]), st.integers(0, 1), st.sampled_from([0, 10, 60, 100, 400]),
       st.sampled_from([0, 2, 4, 8, 12]))
def test_every_plane_in_exactly_one_chunk(in_, jn, kn, slots, ty, kc, guided, pct, cmin):
    rc, p, chunks = plan(in_, jn, kn, slots, ty, kc, guided, pct, cmin)
    if rc != 0:                                             # refused (workspace limits) with a reason, never a broken table
        msg = L.load_library().bbpcg_last_error()
        assert b"z-chunks" in msg or b"workspace" in msg
        return
    check(in_, jn, kn, p, chunks)
    if ty:
        assert p[0] == ty


@pytest.mark.parametrize("cells,uniform", [((512, 512, 512), True), ((1024, 1024, 1024), True), ((512, 512, 256), True), ((256, 256, 256), False),
                                           ((512, 256, 128), False), ((512, 256, 256), False), ((256, 128, 128), False), ((48, 48, 48), False)])
def test_benchmarked_block_shapes(cells, uniform):
    """uniform ~24-plane chunks when that gives >= 7 items per resident CTA, guided (decreasing) chunks otherwise"""
    rc, p, chunks = plan(*cells)
    assert rc == 0
    check(*cells, p, chunks)
    ty, nbx, nby, nbz, kc0 = p
    sizes = [hi - lo + 1 for lo, hi in chunks]
    assert ty == 8
    if uniform:
        assert kc0 == 24 or nbx * nby * nbz == MAXBLOCKS or kc0 == -(-cells[2] // nbz)
        assert nbx * nby * nbz >= 7 * SLOTS and max(sizes) - min(sizes[:-1] or sizes) == 0
    else:
        assert all(a >= b for a, b in zip(sizes[:-1], sizes[1:-1])) and min(sizes) >= min(4, cells[2]) - 2
        assert sizes[0] <= max(4, -(-cells[2] * nbx * nby * 60 // (100 * SLOTS)))


def test_forced_uniform_chunks_and_tile_height():
    rc, p, chunks = plan(256, 256, 256, ty=7, kc=16)
    assert rc == 0 and p[0] == 7 and p[3] == 16 and p[4] == 16
    check(256, 256, 256, p, chunks)
    rc, p, chunks = plan(36, 28, 22, ty=5, kc=11)          # the decomposed ragged case of tests/test_gpu_parity.py
    assert rc == 0 and [hi - lo + 1 for lo, hi in chunks] == [11, 11] and chunks == [(12, 22), (1, 11)]


def test_bad_arguments_are_refused():
    lib = L.load_library()
    out = (C.c_int * 5)()
    ztab = (C.c_int * 8)()
    assert lib.bbpcg_plan_zchunks(0, 8, 8, SLOTS, 0, 0, 1, 0, 0, ztab, 8, out) != 0
    assert lib.bbpcg_plan_zchunks(128, 8, 4096, SLOTS, 0, 1, 1, 0, 0, ztab, 8, out) != 0        # table too small for 4096 one-plane chunks
    assert b"z-chunks" in lib.bbpcg_last_error() or b"workspace" in lib.bbpcg_last_error()
