"""CPU tests (no GPU) of the host-side decomposition contract: bb_domain_* in libbbpcg.so
(csrc/bb_domain.c) against hand-computed index ranges (src/domain.c:918-1486), against the
oracle's independent restatement (oracle/pcg_ref.c bbo_domain_fill) and against the
decomp.config / flow.config grammar (src/domain.c:72-160, tools/src/decomp_reader.c:112-154)."""
import ctypes as C
import glob
import os

import numpy as np
import pytest

import bbpcg
from bbpcg import lib as L
from bbpcg import synth
from bbpcg.grid import BC_SETS, DomStruct, GridInfo, PressureBC, grid_shape, NEUMANN, PERIODIC, PROC_NULL
from oracle import binding as ob

REF = "/root/reference"


def _bytes(s):
    return bytes(C.string_at(C.byref(s), C.sizeof(s)))


def _grid_dict(g):
    return {n: getattr(g, n) for n, _ in GridInfo._fields_}


@pytest.mark.parametrize("blocks", [(1, 1, 1), (2, 1, 1), (1, 1, 2), (2, 2, 1), (2, 2, 2), (4, 1, 2), (3, 2, 1)])
@pytest.mark.parametrize("bc", ["cavity", "duct", "channel", "periodic", "box"])
def test_fill_matches_oracle_restatement(blocks, bc):
    cells = (24 * blocks[0] // blocks[0] * blocks[0], 12 * blocks[1], 6 * blocks[2])
    extent = (-6., 6., 0., 3., -1., 2.)
    dec = bbpcg.Decomposition.uniform(extent, cells, blocks, BC_SETS[bc])
    o = ob.Oracle(extent, cells, blocks, BC_SETS[bc])
    assert dec.nranks == o.nblocks
    for r in range(dec.nranks):
        a, b = dec.doms[r], o.dom(r)
        for gname in ("Gcc", "Gfx", "Gfy", "Gfz"):
            assert _grid_dict(getattr(a, gname)) == _grid_dict(getattr(b, gname)), (r, gname)
        for f, _ in DomStruct._fields_[4:]:
            assert getattr(a, f) == getattr(b, f), (r, f)


def test_fill_hand_computed_96_cube_2x2x1():
    """examples/channel: 96^3 on 2x2x1, block (1,1,0) = rank 3 (src/domain.c:141: rank = I + J*In + K*In*Jn)"""
    dec = bbpcg.Decomposition.uniform((-6, 6, -6, 6, -6, 6), (96, 96, 96), (2, 2, 1), BC_SETS["channel"])
    d = dec.doms[3]
    assert (d.I, d.J, d.K, d.rank) == (1, 1, 0, 3)
    assert (d.xn, d.yn, d.zn) == (48, 48, 96) and d.dx == 6.0 / 48 and d.xs == 0.0 and d.xe == 6.0
    g = d.Gcc
    # one ghost layer; local interior 1..n; global start chains off the west / south block
    assert (g.get("is"), g.get("ie"), g.get("in"), g.get("isb"), g.get("ieb"), g.get("inb")) == (49, 96, 48, 48, 97, 50)
    assert (g.get("js"), g.get("je")) == (49, 96) and (g.get("ks"), g.get("ke"), g.get("kn"), g.get("knb")) == (1, 96, 96, 98)
    assert (g.get("_is"), g.get("_ie"), g.get("_isb"), g.get("_ieb")) == (1, 48, 0, 49)
    assert (g.s1, g.s2, g.s3, g.s1b, g.s2b, g.s3b) == (48, 48 * 48, 48 * 48 * 96, 50, 50 * 50, 50 * 50 * 98)
    assert (g.s2_i, g.s2_j, g.s2_k) == (48 * 96, 48 * 96, 48 * 48)
    # face grids: one more face along their own axis, PERMUTED storage (src/bluebottle.h:70-73)
    fx, fy, fz = d.Gfx, d.Gfy, d.Gfz
    assert (fx.get("in"), fx.get("inb"), fx.get("is")) == (49, 51, 49)             # shares the block-boundary face
    assert (fx.s1b, fx.s2b, fx.s3b) == (50, 50 * 98, 50 * 98 * 51)                  # j fastest, then k, then i
    assert (fy.get("jn"), fy.get("jnb")) == (49, 51) and (fy.s1b, fy.s2b, fy.s3b) == (98, 98 * 50, 98 * 50 * 51)   # k, i, j
    assert (fz.get("kn"), fz.get("knb")) == (97, 99) and (fz.s1b, fz.s2b, fz.s3b) == (50, 50 * 50, 50 * 50 * 99)
    # neighbours from the PRESSURE BCs (src/domain.c:1147-1210): x periodic wraps, y is a wall, z periodic onto itself
    assert (d.w, d.e, d.s, d.n, d.b, d.t) == (2, 2, 1, PROC_NULL, 3, 3)
    d0 = dec.doms[0]
    assert (d0.w, d0.e, d0.s, d0.n, d0.b, d0.t) == (1, 1, PROC_NULL, 2, 0, 0)
    assert grid_shape(d, "Gfx") == (51, 98, 50) and grid_shape(d, "Gfy") == (51, 50, 98) and grid_shape(d, "Gcc") == (98, 50, 50)


@pytest.mark.parametrize("bc", sorted(BC_SETS))
def test_neighbour_links_are_reciprocal(bc):
    dec = bbpcg.Decomposition.uniform((0, 4, 0, 4, 0, 4), (16, 16, 16), (2, 2, 2), BC_SETS[bc])
    opp = {"e": "w", "w": "e", "n": "s", "s": "n", "t": "b", "b": "t"}
    for r in range(8):
        for side, back in opp.items():
            nb = getattr(dec.doms[r], side)
            if nb >= 0:
                assert getattr(dec.doms[nb], back) == r
    p = dec.bc
    assert (dec.doms[0].w >= 0) == (p.pW == PERIODIC) and (dec.doms[7].t >= 0) == (p.pT == PERIODIC)


def _write_flow(path, ext, cells, blocks, bc, extra=""):
    names = {PERIODIC: "PERIODIC", NEUMANN: "NEUMANN 0"}
    with open(path, "w") as f:
        f.write("GLOBAL DOMAIN\n(Xs, Xe, Xn) %g %g %d\n(Ys, Ye, Yn) %g %g %d\n(Zs, Ze, Zn) %g %g %d\n\n" %
                (ext[0], ext[1], cells[0], ext[2], ext[3], cells[1], ext[4], ext[5], cells[2]))
        f.write("MPI/GPU SUBDOMAIN DECOMPOSITION\n(In, Jn, Kn) %d %d %d\n\nPHYSICAL PARAMETERS\nrho_f 2.5\nnu 1.0\n\n" % blocks)
        f.write("SIMULATION PARAMETERS\nduration 1.0\nCFL 0.5\npp_max_iter 1234\npp_residual 1e-7\n" + extra)
        f.write("\nBOUNDARY CONDITIONS\nv_bc_tdelay 0\nPRESSURE\n")
        for k, v in zip(("pW", "pE", "pS", "pN", "pB", "pT"), bc):
            f.write("bc.%s %s\n" % (k, names[v]))
        f.write("X-VELOCITY\nbc.uW PERIODIC\n")


def test_decomp_config_round_trip(tmp_path):
    ext, cells, blocks = (-6., 6., -3., 3., 0., 12.), (96, 48, 96), (2, 2, 2)
    dec = bbpcg.Decomposition.uniform(ext, cells, blocks, BC_SETS["sedimentation"])
    dpath, fpath = str(tmp_path / "decomp.config"), str(tmp_path / "flow.config")
    dec.write_decomp(dpath)
    _write_flow(fpath, ext, cells, blocks, BC_SETS["sedimentation"])
    text = open(dpath).read()
    # record grammar of src/domain.c:138-159 as tools/src/decomp_reader.c:142-154 writes it
    assert text.startswith("(I, J, K) 0 0 0\n(Xs, Xe, Xn) -6.00 0.00 48\n(Ys, Ye, Yn) -3.00 0.00 24\n(Zs, Ze, Zn) 0.00 6.00 48\n\n(I, J, K) 1 0 0\n")
    assert text.count("(I, J, K)") == 8
    back = bbpcg.Decomposition.from_files(fpath, dpath)
    assert back.params == {"rho_f": 2.5, "pp_residual": 1e-7, "pp_max_iter": 1234}
    assert _bytes(back.DOM) == _bytes(dec.DOM) and _bytes(back.bc) == _bytes(dec.bc)
    for r in range(8):
        assert _bytes(back.doms[r]) == _bytes(dec.doms[r]), r
    # the oracle's reader agrees on the same file
    e, n, ijk = (C.c_double * 48)(), (C.c_int * 24)(), (C.c_int * 24)()
    assert ob.load().bbo_read_decomp(dpath.encode(), 8, e, n, ijk) == 8      # records read with the reference's own fscanf formats
    assert list(ijk[3:6]) == [1, 0, 0] and list(n[0:3]) == [48, 24, 48] and list(e[6:8]) == [0.0, 6.0]


def test_unequal_blocks_from_a_hand_written_decomp(tmp_path):
    """decomp.config may hold unequal blocks; dx is per block from the text file (src/domain.c:1219-1224)"""
    fpath, dpath = str(tmp_path / "flow.config"), str(tmp_path / "decomp.config")
    _write_flow(fpath, (0., 10., 0., 4., 0., 4.), (40, 16, 16), (2, 1, 1), BC_SETS["box"])
    open(dpath, "w").write("(I, J, K) 0 0 0\n(Xs, Xe, Xn) 0.0 2.5 10\n(Ys, Ye, Yn) 0.0 4.0 16\n(Zs, Ze, Zn) 0.0 4.0 16\n\n"
                           "(I, J, K) 1 0 0\n(Xs, Xe, Xn) 2.5 10.0 30\n(Ys, Ye, Yn) 0.0 4.0 16\n(Zs, Ze, Zn) 0.0 4.0 16\n\n")
    dec = bbpcg.Decomposition.from_files(fpath, dpath)
    a, b = dec.doms
    assert (a.xn, b.xn) == (10, 30) and a.dx == b.dx == 0.25
    assert (b.Gcc.get("is"), b.Gcc.get("ie")) == (11, 40) and b.Gfx.get("is") == a.Gfx.get("ie") == 11
    assert (a.e, b.w, a.w, b.e) == (1, 0, PROC_NULL, PROC_NULL)


def test_reader_errors(tmp_path):
    lib = bbpcg.load_library()
    DOM, pbc, fp, ptr = DomStruct(), PressureBC(), L.FlowParams(), C.POINTER(DomStruct)()
    rc = lib.bb_domain_read(b"/nonexistent/flow.config", b"/nonexistent/decomp.config", C.byref(DOM), C.byref(ptr), C.byref(pbc), C.byref(fp))
    assert rc == -5 and b"Could not open file" in lib.bbpcg_last_error()          # domain.c:83-86 message
    fpath, dpath = str(tmp_path / "flow.config"), str(tmp_path / "decomp.config")
    _write_flow(fpath, (0., 1., 0., 1., 0., 1.), (8, 8, 8), (2, 1, 1), BC_SETS["box"])
    open(dpath, "w").write("(I, J, K) 0 0 0\n(Xs, Xe, Xn) 0 0.5 4\n(Ys, Ye, Yn) 0 1 8\n(Zs, Ze, Zn) 0 1 8\n\n")   # one record short
    with pytest.raises(RuntimeError, match="record 1 of 2 unreadable"):
        bbpcg.Decomposition.from_files(fpath, dpath)
    open(dpath, "w").write("(I, J, K) 1 0 0\n(Xs, Xe, Xn) 0.5 1 4\n(Ys, Ye, Yn) 0 1 8\n(Zs, Ze, Zn) 0 1 8\n\n"
                           "(I, J, K) 0 0 0\n(Xs, Xe, Xn) 0 0.5 4\n(Ys, Ye, Yn) 0 1 8\n(Zs, Ze, Zn) 0 1 8\n\n")      # not rank order
    with pytest.raises(RuntimeError, match="I-fastest order"):
        bbpcg.Decomposition.from_files(fpath, dpath)
    # DIRICHLET is not a pressure BC (src/domain.c:216-287)
    txt = open(fpath).read().replace("bc.pW NEUMANN 0", "bc.pW DIRICHLET 0")
    open(fpath, "w").write(txt)
    with pytest.raises(RuntimeError, match="pressure boundary block"):
        bbpcg.Decomposition.from_files(fpath, dpath)


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree only exists in the authoring container")
@pytest.mark.parametrize("example", ["channel", "lid-driven-cavity", "pressure-driven-duct", "sedimentation", "shear",
                                     "cuboctohedron-collision"])
def test_reads_the_reference_examples(example):
    """every shipped examples/*/flow.config + decomp.config pair parses (by key: the shipped
    flow.config files are out of sync with the reference's positional parser)"""
    d = os.path.join(REF, "examples", example)
    if not (os.path.exists(os.path.join(d, "flow.config")) and os.path.exists(os.path.join(d, "decomp.config"))):
        pytest.skip("example has no config pair")
    dec = bbpcg.Decomposition.from_files(os.path.join(d, "flow.config"), os.path.join(d, "decomp.config"))
    D = dec.DOM
    assert dec.nranks == D.In * D.Jn * D.Kn >= 1
    assert sum(dec.doms[r].xn for r in range(D.In)) == D.xn
    assert sum(dec.doms[r * D.In].yn for r in range(D.Jn)) == D.yn
    assert dec.params["pp_max_iter"] == 2000 and dec.params["pp_residual"] == 1e-6
    expect = {"channel": "channel", "lid-driven-cavity": "cavity", "pressure-driven-duct": "duct", "sedimentation": "sedimentation"}
    if example in expect:
        assert tuple(_bytes(dec.bc)) == tuple(_bytes(PressureBC(*BC_SETS[expect[example]])))
    # same equal split as our writer produces
    same = bbpcg.Decomposition.uniform((D.xs, D.xe, D.ys, D.ye, D.zs, D.ze), (D.xn, D.yn, D.zn), (D.In, D.Jn, D.Kn),
                                       tuple(getattr(dec.bc, k) for k in ("pW", "pE", "pS", "pN", "pB", "pT")))
    for r in range(dec.nranks):
        assert _bytes(same.doms[r]) == _bytes(dec.doms[r])


# ---- synthetic inputs shard without communication ------------------------------------------------
@pytest.mark.parametrize("blocks", [(2, 1, 1), (1, 2, 2), (2, 2, 2)])
@pytest.mark.parametrize("bc", ["duct", "periodic", "box"])
def test_synthetic_velocity_is_decomposition_independent(blocks, bc):
    """each rank fills only its own block, yet the blocks tile the 1-block field bit for bit
    (duplicated block-boundary faces included)"""
    ext, cells = (0., 12., 0., 12., 0., 12.), (16, 12, 8)
    one = bbpcg.Decomposition.uniform(ext, cells, (1, 1, 1), BC_SETS[bc])
    many = bbpcg.Decomposition.uniform(ext, cells, blocks, BC_SETS[bc])
    U = synth.velocity_star(one.doms[0], one.DOM, one.bc)
    from bbpcg.grid import as_ijk
    for r in range(many.nranks):
        d = many.doms[r]
        u = synth.velocity_star(d, many.DOM, many.bc)
        for comp, grid in enumerate(("Gfx", "Gfy", "Gfz")):
            g = getattr(d, grid)
            loc = as_ijk(u[comp], grid)[1:g.get("in") + 1, 1:g.get("jn") + 1, 1:g.get("kn") + 1]
            i0, j0, k0 = g.get("is"), g.get("js"), g.get("ks")
            glob_ = as_ijk(U[comp], grid)[i0:i0 + g.get("in"), j0:j0 + g.get("jn"), k0:k0 + g.get("kn")]
            assert np.array_equal(loc, glob_), (r, grid)


def test_torch_and_numpy_synthetic_inputs_are_bit_identical():
    import torch
    dec = bbpcg.Decomposition.uniform((0., 12., 0., 12., 0., 12.), (20, 12, 16), (2, 1, 1), BC_SETS["duct"])
    for r in range(2):
        a = synth.velocity_star(dec.doms[r], dec.DOM, dec.bc)
        b = synth.velocity_star_torch(dec.doms[r], dec.DOM, dec.bc, torch.device("cpu"))
        for x, y in zip(a, b):
            assert np.array_equal(x, y.numpy())
    # flags: torch builder == the oracle's flag builder (cuda_particle.cu:1605-1639)
    for bc in ("duct", "box", "periodic", "cavity"):
        dec = bbpcg.Decomposition.uniform((0., 12., 0., 12., 0., 12.), (8, 6, 10), (2, 1, 2), BC_SETS[bc])
        o = ob.Oracle((0., 12., 0., 12., 0., 12.), (8, 6, 10), (2, 1, 2), BC_SETS[bc])
        o.build_flags_noparts()
        for r in range(4):
            fu, fv, fw = synth.flags_noparts_torch(dec.doms[r], dec.DOM, dec.bc, torch.device("cpu"))
            assert np.array_equal(fu.numpy(), o.array(r, ob.FLAG_U)) and np.array_equal(fv.numpy(), o.array(r, ob.FLAG_V))
            assert np.array_equal(fw.numpy(), o.array(r, ob.FLAG_W))


def test_solvability_of_the_synthetic_rhs():
    """wall-normal faces are zeroed so sum(rhs) = 0 to round-off (cuda_solvability, cuda_bluebottle.cu:2313-2492)"""
    from cases import Case
    for bc in ("box", "duct", "periodic"):
        case = Case((16, 12, 20), bc=bc)
        case.o.rhs(1.0, 1e-3)
        rhs = case.o.array(0, ob.RHS_P)
        assert abs(rhs.sum()) < 1e-9 * np.abs(rhs).sum()
