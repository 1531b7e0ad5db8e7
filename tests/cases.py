"""Shared case builders for the parity tests: the same seeded inputs go to the CPU oracle
(oracle/binding.py), to the reference's own kernels (oracle/_ref, when built) and to the CUDA
library through its C ABI."""
import ctypes as C
import os

import numpy as np

from bbpcg import synth
from bbpcg.grid import BC_SETS, DomStruct
from oracle import binding as ob

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_LIB = os.path.join(ROOT, "oracle", "_ref", "libbbref.so")


class Case:
    """One synthetic problem held in a multi-block oracle state."""

    def __init__(self, cells, blocks=(1, 1, 1), bc="cavity", extent=None, noise=1.0, nparts=0, radius=1.0, omp=False,
                 seed=7):
        self.cells, self.blocks, self.bcname = tuple(cells), tuple(blocks), bc
        self.extent = extent or (0., 12., 0., 12. * cells[1] / cells[0], 0., 12. * cells[2] / cells[0])
        self.o = ob.Oracle(self.extent, cells, blocks, BC_SETS[bc], omp=omp)
        self.nparts = nparts
        if nparts:
            px, py, pz, pr = synth.random_spheres(self.o.DOM, nparts, radius)
            self.parts = (px, py, pz, pr)
            self.o.build_cages(px, py, pz, pr)
        else:
            self.o.build_flags_noparts()
        for r in range(self.o.nblocks):
            u, v, w = synth.velocity_star(self.o.dom(r), self.o.DOM, self.o.bc, noise=noise, seed=seed)
            self.o.array(r, ob.U_STAR)[...] = u
            self.o.array(r, ob.V_STAR)[...] = v
            self.o.array(r, ob.W_STAR)[...] = w
        self.o.jacobi_init()

    def inputs(self, rank=0):
        a = self.o.array
        return dict(flag_u=a(rank, ob.FLAG_U), flag_v=a(rank, ob.FLAG_V), flag_w=a(rank, ob.FLAG_W),
                    phase=a(rank, ob.PHASE), phase_shell=a(rank, ob.PHASE_SHELL),
                    u_star=a(rank, ob.U_STAR), v_star=a(rank, ob.V_STAR), w_star=a(rank, ob.W_STAR))

    def solve_oracle(self, **kw):
        kw.setdefault("parts", self.nparts > 0)
        return self.o.solve(**kw)

    def seed_epilogue(self, seed=23, phi=True):
        """Seeded inputs of the solve epilogue in the oracle's arrays: p0 everywhere and -- unless the phi of a
        solve is to be kept -- phi INCLUDING its ghosts (so the exchange and the Neumann copy have work to undo)."""
        for r in range(self.o.nblocks):
            rng = np.random.default_rng(seed + 1000 * r)
            shape = self.o.array(r, ob.PHI).shape
            vals = rng.standard_normal(shape)
            if phi:
                self.o.array(r, ob.PHI)[...] = vals
            self.o.array(r, ob.P0)[...] = rng.standard_normal(shape)

    def epilogue_inputs(self, rank=0):
        return dict(phi=self.o.array(rank, ob.PHI).copy(), p0=self.o.array(rank, ob.P0).copy())


def rel_l2(a, b):
    return float(np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-300))


def load_ref():
    """The reference's own kernels (O1); None when oracle/_ref was not built."""
    if not os.path.exists(REF_LIB):
        return None
    lib = C.CDLL(REF_LIB)
    D = C.POINTER(DomStruct)
    vp = C.c_void_p
    lib.bbref_init.argtypes = [D, D]
    lib.bbref_set_inputs.argtypes = [vp] * 8 + [C.c_int]
    lib.bbref_set_inputs_dev.argtypes = [vp] * 8 + [C.c_int]
    lib.bbref_solve.argtypes = [C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, C.POINTER(C.c_int),
                                C.POINTER(C.c_double), C.POINTER(C.c_float)]
    lib.bbref_get.argtypes = [C.c_int, vp]
    lib.bbref_spmv.argtypes = [vp, C.c_int, vp]
    lib.bbref_exchange.argtypes = [vp]
    lib.bbref_epilogue.argtypes = [vp, vp, vp, C.c_double, C.c_double, C.c_double, vp, vp, vp, vp, vp, C.POINTER(C.c_float)]
    lib.bbref_exchange_face.argtypes = [vp, C.c_int]
    lib.bbref_solvability.argtypes = [vp, vp, vp, C.c_int, C.POINTER(C.c_double)]
    lib.bbref_dom_BC_star.argtypes = [vp, vp, vp, C.POINTER(C.c_int), C.POINTER(C.c_double)]
    lib.bbref_build_cages.argtypes = [C.c_int, C.c_int] + [vp] * 10
    lib.bbref_dev_ptr.argtypes = [C.c_int]
    lib.bbref_dev_ptr.restype = C.c_void_p
    return lib


def ref_epilogue(lib, case, phi, p0, rho_f=1.0, dt=1e-3):
    """The reference's own epilogue kernels (O1) on one block: returns u, v, w, p, ghost-filled phi and the ms."""
    from bbpcg.grid import BC_SETS, grid_shape
    d = case.o.dom(0)
    out = {k: np.zeros(grid_shape(d, g)) for k, g in (("u", "Gfx"), ("v", "Gfy"), ("w", "Gfz"), ("p", "Gcc"), ("phi", "Gcc"))}
    pbc = (C.c_int * 6)(*BC_SETS[case.bcname])
    P = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    ms = C.c_float()
    phi_c = None if phi is None else np.ascontiguousarray(phi)
    p0_c = np.ascontiguousarray(p0)
    rc = lib.bbref_epilogue(None if phi_c is None else P(phi_c), P(p0_c), C.cast(pbc, C.c_void_p), rho_f, dt, 0.01,
                            P(out["u"]), P(out["v"]), P(out["w"]), P(out["p"]), P(out["phi"]), C.byref(ms))
    assert rc == 0
    out["ms"] = ms.value
    return out


def face_interior(d, grid, a):
    """the entries cuda_project writes: Gf?._is.._ie x interior of the other two (all of a Gf? minus its ghosts)"""
    return a[1:-1, 1:-1, 1:-1]


def face_exchange_inputs(case, rank, seed):
    """seeded Gfx / Gfy / Gfz arrays (ghosts included) for the face-grid halo exchanges: {name: (array, BBPCG grid code)}"""
    from bbpcg.grid import grid_shape
    d = case.o.dom(rank)
    rng = np.random.default_rng(seed + 1000 * rank)
    return {k: (rng.standard_normal(grid_shape(d, g)), code) for k, g, code in (("u", "Gfx", 1), ("v", "Gfy", 2), ("w", "Gfz", 3))}


# ---- cuda_dom_BC_star: velocity BC tables, 18 (type, value) pairs, component-major (u on W,E,S,N,B,T, then v, then w) ----
_P, _D, _N, _PRE = 0, 1, 2, 3        # PERIODIC, DIRICHLET, NEUMANN, PRECURSOR (src/bluebottle.h:218-254)
BC_STAR_TABLES = {
    "dirichlet": ([_D] * 18, [0.1 * (e + 1) * (-1) ** e for e in range(18)]),
    "neumann": ([_N] * 18, [0.0] * 18),
    # no-slip walls with a lid: u = 1 on the NORTH wall, everything else 0 (the pattern of examples/lid-driven-cavity)
    "cavity_lid": ([_D] * 18, [0., 0., 0., 1., 0., 0.] + [0.] * 12),
    "mixed": ([_D, _N, _D, _D, _P, _PRE, _N, _D, _N, _D, _D, _N, _PRE, _P, _D, _N, _D, _D],
              [1.5, 9., -0.25, 0., 7., 7., 9., 0.75, 9., -1.25, 2., 9., 7., 7., 0.5, 9., -0.125, 3.]),
}


def velocity_bc(table):
    """lib.VelocityBC of a BC_STAR_TABLES entry"""
    from bbpcg.lib import VelocityBC
    types, vals = BC_STAR_TABLES[table] if isinstance(table, str) else table
    out = VelocityBC()
    for e in range(18):
        out.type[e // 6][e % 6] = types[e]
        out.val[e // 6][e % 6] = vals[e]
    return out


def ref_dom_BC_star(lib, case, arrs, table):
    """the reference's own BC_* kernels (O1) on host copies of one block's u*, v*, w* (in place)"""
    types, vals = BC_STAR_TABLES[table] if isinstance(table, str) else table
    P = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    assert lib.bbref_dom_BC_star(P(arrs["u"]), P(arrs["v"]), P(arrs["w"]), (C.c_int * 18)(*types), (C.c_double * 18)(*vals)) == 0
    return arrs


# ---- cuda_build_cages: particle lists (global coordinates) chosen to hit the corner cases of the cage logic ----
def cage_particles(case_name, extent, cells):
    """(x, y, z, r) arrays.  dx = extent / cells; radii are a few cells."""
    xs, xe, ys, ye, zs, ze = extent
    dx = (xe - xs) / cells[0]
    L = np.array([xe - xs, ye - ys, ze - zs])
    o = np.array([xs, ys, zs])
    frac = {
        # well inside; two spheres OVERLAPPING (the later index wins the shared cells); one tiny (r < dx)
        "inside": [(0.30, 0.35, 0.40, 3.3), (0.62, 0.55, 0.52, 2.6), (0.70, 0.60, 0.55, 2.2), (0.2, 0.75, 0.8, 0.7)],
        # cages cut by the domain faces: lower corner, upper face, one centred exactly on a face, one outside the block
        "faces": [(0.04, 0.05, 0.06, 2.9), (0.5, 0.97, 0.5, 2.4), (1.0, 0.5, 0.3, 2.0), (0.5, 0.5, -0.4, 1.5), (0.93, 0.08, 0.95, 3.1)],
        "single": [(0.5, 0.5, 0.5, 3.0)],
    }[case_name]
    a = np.array([[o[0] + f[0] * L[0], o[1] + f[1] * L[1], o[2] + f[2] * L[2], f[3] * dx] for f in frac])
    return a[:, 0].copy(), a[:, 1].copy(), a[:, 2].copy(), a[:, 3].copy()


def ref_build_cages(lib, case, parts, NPARTS=None):
    """the reference's own cage kernels (O1) on one block: returns flag_u, flag_v, flag_w, phase, phase_shell"""
    from bbpcg.grid import BC_SETS, grid_shape
    d = case.o.dom(0)
    px, py, pz, pr = [np.ascontiguousarray(v, dtype=np.float64) for v in parts]
    n = len(px)
    out = {k: np.full(grid_shape(d, g), 7, dtype=np.int32) for k, g in (("flag_u", "Gfx"), ("flag_v", "Gfy"), ("flag_w", "Gfz"),
                                                                          ("phase", "Gcc"), ("phase_shell", "Gcc"))}
    pbc = (C.c_int * 6)(*BC_SETS[case.bcname])
    P = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    assert lib.bbref_build_cages(n if NPARTS is None else NPARTS, n, P(px), P(py), P(pz), P(pr), C.cast(pbc, C.c_void_p),
                                 P(out["flag_u"]), P(out["flag_v"]), P(out["flag_w"]), P(out["phase"]), P(out["phase_shell"])) == 0
    return out
