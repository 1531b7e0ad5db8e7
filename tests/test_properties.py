"""Property-based CPU tests (hypothesis) of the host-side contracts that every multi-GPU run depends on: random block
decompositions, random boundary sets, ragged extents.  They hold the product's bb_domain_fill (csrc/bb_domain.c) to the
oracle's independent restatement and to structural invariants of src/domain.c:918-1486, and the oracle's four halo
exchanges to the copy semantics of src/mpi_comm.c:257-405 (a ghost face equals the neighbour's boundary plane)."""
import numpy as np
from hypothesis import given, settings, strategies as st

import bbpcg
from bbpcg.grid import GridInfo, NEUMANN, PERIODIC
from oracle import binding as ob

blocks_st = st.tuples(st.integers(1, 3), st.integers(1, 3), st.integers(1, 3))
per_block_st = st.tuples(st.integers(2, 6), st.integers(2, 5), st.integers(2, 5))
bc_axis = st.sampled_from([(PERIODIC, PERIODIC), (NEUMANN, NEUMANN)])
bc_st = st.tuples(bc_axis, bc_axis, bc_axis).map(lambda t: t[0] + t[1] + t[2])


@settings(max_examples=30, deadline=None, derandomize=True)
@given(blocks_st, per_block_st, bc_st)
def test_decomposition_invariants(blocks, per, bc):
    cells = tuple(b * p for b, p in zip(blocks, per))
    extent = (0., 1.5 * blocks[0], -1., -1. + blocks[1], 2., 2. + 0.5 * blocks[2])
    dec = bbpcg.Decomposition.uniform(extent, cells, blocks, bc)
    o = ob.Oracle(extent, cells, blocks, bc)
    n = dec.nranks
    covered = np.zeros(cells[::-1], dtype=np.int32)
    for r in range(n):
        d, od = dec.doms[r], o.dom(r)
        for gname in ("Gcc", "Gfx", "Gfy", "Gfz"):
            a, b = getattr(d, gname), getattr(od, gname)
            assert all(getattr(a, f) == getattr(b, f) for f, _ in GridInfo._fields_), (r, gname)
        assert d.rank == r == d.I + d.J * blocks[0] + d.K * blocks[0] * blocks[1]              # src/domain.c:141
        g = d.Gcc
        covered[g.get("ks") - 1:g.get("ke"), g.get("js") - 1:g.get("je"), g.get("is") - 1:g.get("ie")] += 1
        # a face grid has one more entry along its normal and shares the boundary face with its lower neighbour
        assert d.Gfx.get("in") == d.xn + 1 and d.Gfy.get("jn") == d.yn + 1 and d.Gfz.get("kn") == d.zn + 1
        if d.I > 0:
            assert d.Gfx.get("is") == dec.doms[r - 1].Gfx.get("ie")
        # neighbour links: reciprocal, and absent exactly on non-periodic outer faces
        for mine, theirs, outer, t in ((d.e, "w", d.I == blocks[0] - 1, bc[1]), (d.w, "e", d.I == 0, bc[0]),
                                       (d.n, "s", d.J == blocks[1] - 1, bc[3]), (d.s, "n", d.J == 0, bc[2]),
                                       (d.t, "b", d.K == blocks[2] - 1, bc[5]), (d.b, "t", d.K == 0, bc[4])):
            if outer and t != PERIODIC:
                assert mine < 0
            else:
                assert 0 <= mine < n and getattr(dec.doms[mine], theirs) == r
    assert (covered == 1).all()                                                                   # blocks tile the domain exactly


@settings(max_examples=20, deadline=None, derandomize=True)
@given(blocks_st, per_block_st, bc_st, st.integers(0, 2 ** 31 - 1))
def test_exchange_copies_the_neighbours_boundary_plane(blocks, per, bc, seed):
    """after mpi_cuda_exchange_G??: ghost face == the plane the neighbour sent (its _ie / _is; _ie-1 / _is+1 along a face
    grid's own normal), everything else unchanged -- for all four grids, any decomposition, any BC set"""
    cells = tuple(b * p for b, p in zip(blocks, per))
    o = ob.Oracle((0., 1., 0., 1., 0., 1.), cells, blocks, bc)
    rng = np.random.default_rng(seed)
    for aid, normal in ((ob.PB_Q, -1), (ob.U, 0), (ob.V, 1), (ob.W, 2)):
        before = []
        for r in range(o.nblocks):
            a = o.array(r, aid)
            a[...] = rng.standard_normal(a.shape)
            before.append(a.copy())
        o.exchange(aid)

        def ijk(a):                  # view indexed [i, j, k] whatever the grid's storage order (src/bluebottle.h:70-73)
            return {-1: a.transpose(2, 1, 0), 2: a.transpose(2, 1, 0), 0: a, 1: a.transpose(1, 0, 2)}[normal] if normal != 0 else a.transpose(0, 2, 1)
        for r in range(o.nblocks):
            d = o.dom(r)
            A, B = ijk(o.array(r, aid)), ijk(before[r])
            expect = B.copy()
            I = slice(1, -1)
            for axis, (lo_nb, hi_nb) in enumerate(((d.w, d.e), (d.s, d.n), (d.b, d.t))):
                shift = 1 if axis == normal else 0
                def plane(arr, ax, idx):
                    sl = [I, I, I]; sl[ax] = idx
                    return arr[tuple(sl)]
                if lo_nb >= 0:       # my low ghost <- the low neighbour's high plane (_ie, or _ie-1 for the shared face)
                    src = ijk(before[lo_nb])
                    sl = [I, I, I]; sl[axis] = 0
                    expect[tuple(sl)] = plane(src, axis, -2 - shift)
                if hi_nb >= 0:
                    src = ijk(before[hi_nb])
                    sl = [I, I, I]; sl[axis] = -1
                    expect[tuple(sl)] = plane(src, axis, 1 + shift)
            assert np.array_equal(A, expect), (aid, r)
