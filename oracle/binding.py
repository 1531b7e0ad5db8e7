"""ctypes binding of the CPU oracle (oracle/pcg_ref.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(_HERE), "bluebottle-3.0_b200"))
from bbpcg.grid import DomStruct, PressureBC, grid_shape  # noqa: E402  (types only: the grid contract)

(FLAG_U, FLAG_V, FLAG_W, PHASE, PHASE_SHELL, U_STAR, V_STAR, W_STAR, RHS_P, PHI, PB_Q,
 INVM, R_Q, Z_Q, P_Q, APB_Q, U, V, W, P0, P) = range(21)

_GRID_OF = {FLAG_U: "Gfx", FLAG_V: "Gfy", FLAG_W: "Gfz", PHASE: "Gcc", PHASE_SHELL: "Gcc",
            U_STAR: "Gfx", V_STAR: "Gfy", W_STAR: "Gfz", RHS_P: "Gcc", PHI: "Gcc", PB_Q: "Gcc",
            U: "Gfx", V: "Gfy", W: "Gfz", P0: "Gcc", P: "Gcc"}
_INT_IDS = (FLAG_U, FLAG_V, FLAG_W, PHASE, PHASE_SHELL)


class Result(C.Structure):
    _fields_ = [("status", C.c_int), ("niter", C.c_int), ("resid", C.c_double),
                ("sp_rhs", C.c_double), ("sp_rq0", C.c_double)]


def build(omp=False):
    name = "liboracle_omp.so" if omp else "liboracle.so"
    path = os.path.join(_HERE, name)
    src = os.path.join(_HERE, "pcg_ref.c")
    if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, name], stdout=subprocess.DEVNULL)
    return path


_libs = {}


def load(omp=False):
    if omp in _libs:
        return _libs[omp]
    lib = C.CDLL(build(omp))
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
    lib.bbo_create.restype = C.c_void_p
    lib.bbo_create.argtypes = [dp, ip, ip, ip]
    lib.bbo_create_blocks.restype = C.c_void_p
    lib.bbo_create_blocks.argtypes = [dp, ip, ip, ip, dp, ip]
    lib.bbo_destroy.argtypes = [C.c_void_p]
    lib.bbo_nblocks.argtypes = [C.c_void_p]
    lib.bbo_dom.restype = C.POINTER(DomStruct)
    lib.bbo_dom.argtypes = [C.c_void_p, C.c_int]
    lib.bbo_DOM.restype = C.POINTER(DomStruct)
    lib.bbo_DOM.argtypes = [C.c_void_p]
    lib.bbo_array.restype = C.c_void_p
    lib.bbo_array.argtypes = [C.c_void_p, C.c_int, C.c_int]
    lib.bbo_write_decomp.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
    lib.bbo_read_decomp.argtypes = [C.c_char_p, C.c_int, dp, ip, ip]
    lib.bbo_build_flags_noparts.argtypes = [C.c_void_p]
    lib.bbo_build_cages.argtypes = [C.c_void_p, C.c_int, dp, dp, dp, dp]
    lib.bbo_jacobi_init.argtypes = [C.c_void_p]
    lib.bbo_rhs.argtypes = [C.c_void_p, C.c_double, C.c_double]
    lib.bbo_exchange_Gcc.argtypes = [C.c_void_p, C.c_int]
    lib.bbo_spmv_noparts.argtypes = [C.c_void_p, C.c_int]
    lib.bbo_spmv_parts.argtypes = [C.c_void_p, C.c_int]
    lib.bbo_solve.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int,
                              dp, C.c_int, C.POINTER(Result)]
    lib.bbo_iterate_fixed.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_int, C.c_int]
    lib.bbo_exchange.argtypes = [C.c_void_p, C.c_int]
    lib.bbo_solvability.argtypes = [C.c_void_p, C.c_int, dp]
    lib.bbo_dom_BC_p.argtypes = [C.c_void_p, C.c_int]
    lib.bbo_dom_BC_star.argtypes = [C.c_void_p, ip, dp]
    lib.bbo_project.argtypes = [C.c_void_p, C.c_double, C.c_double]
    lib.bbo_update_p.argtypes = [C.c_void_p]
    lib.bbo_update_p.restype = C.c_double
    lib.bbo_epilogue.argtypes = [C.c_void_p, C.c_double, C.c_double]
    lib.bbo_epilogue.restype = C.c_double
    lib.bbo_omp_threads.restype = C.c_int
    _libs[omp] = lib
    return lib


def _arr(vals, ctype):
    return (ctype * len(vals))(*vals)


class Oracle:
    """All blocks of one decomposition, held in one process (nblocks = In*Jn*Kn)."""

    def __init__(self, extent, cells, blocks=(1, 1, 1), bc=(0,) * 6, omp=False):
        self.lib = load(omp)
        self.h = self.lib.bbo_create(_arr([float(v) for v in extent], C.c_double), _arr(list(cells), C.c_int),
                                     _arr(list(blocks), C.c_int), _arr(list(bc), C.c_int))
        self.nblocks = self.lib.bbo_nblocks(self.h)
        self.bc = PressureBC(*bc)

    def close(self):
        if self.h:
            self.lib.bbo_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def DOM(self):
        return self.lib.bbo_DOM(self.h).contents

    def dom(self, rank=0):
        return self.lib.bbo_dom(self.h, rank).contents

    def array(self, rank, aid):
        """numpy view (no copy) of one of the oracle's arrays."""
        d = self.dom(rank)
        ptr = self.lib.bbo_array(self.h, rank, aid)
        if aid in _GRID_OF:
            shape = grid_shape(d, _GRID_OF[aid])
        else:
            shape = (d.Gcc.get("kn"), d.Gcc.get("jn"), d.Gcc.get("in"))
        ct = C.c_int if aid in _INT_IDS else C.c_double
        n = int(np.prod(shape))
        buf = (ct * n).from_address(ptr)
        return np.frombuffer(buf, dtype=np.int32 if aid in _INT_IDS else np.float64).reshape(shape)

    def build_flags_noparts(self):
        self.lib.bbo_build_flags_noparts(self.h)

    def build_cages(self, px, py, pz, pr):
        dp = C.POINTER(C.c_double)
        a = [np.ascontiguousarray(v, dtype=np.float64) for v in (px, py, pz, pr)]
        self.lib.bbo_build_cages(self.h, len(a[0]), *[v.ctypes.data_as(dp) for v in a])

    def jacobi_init(self):
        self.lib.bbo_jacobi_init(self.h)

    def rhs(self, rho_f, dt):
        self.lib.bbo_rhs(self.h, rho_f, dt)

    def exchange_Gcc(self, aid):
        self.lib.bbo_exchange_Gcc(self.h, aid)

    def exchange(self, aid):
        """mpi_cuda_exchange_G{cc,fx,fy,fz} on the array `aid` (the grid follows from the array)"""
        self.lib.bbo_exchange(self.h, aid)

    def spmv(self, aid, parts=False):
        (self.lib.bbo_spmv_parts if parts else self.lib.bbo_spmv_noparts)(self.h, aid)

    def solvability(self, out_plane=10):
        """cuda_solvability on the oracle's u_star / v_star / w_star (in place); returns eps[3]"""
        eps = (C.c_double * 3)()
        self.lib.bbo_solvability(self.h, out_plane, eps)
        return [eps[0], eps[1], eps[2]]

    def dom_BC_p(self, aid):
        self.lib.bbo_dom_BC_p(self.h, aid)

    def dom_BC_star(self, types, vals):
        """cuda_dom_BC_star on the oracle's u_star / v_star / w_star (in place); 18 types + 18 values, component-major
        (u on W,E,S,N,B,T, then v, then w)"""
        self.lib.bbo_dom_BC_star(self.h, _arr([int(t) for t in types], C.c_int), _arr([float(v) for v in vals], C.c_double))

    def project(self, rho_f=1.0, dt=1e-3):
        self.lib.bbo_project(self.h, rho_f, dt)

    def update_p(self):
        return self.lib.bbo_update_p(self.h)

    def epilogue(self, rho_f=1.0, dt=1e-3):
        """exchange_Gcc(phi); dom_BC_p(phi); project; update_p -- returns the subtracted mean"""
        return self.lib.bbo_epilogue(self.h, rho_f, dt)

    def solve(self, rho_f=1.0, dt=1e-3, pp_residual=1e-6, pp_max_iter=2000, parts=False):
        cap = pp_max_iter + 3
        hist = np.zeros(cap)
        res = Result()
        self.lib.bbo_solve(self.h, rho_f, dt, pp_residual, pp_max_iter, int(parts),
                           hist.ctypes.data_as(C.POINTER(C.c_double)), cap, C.byref(res))
        return res, hist[: res.niter + 1].copy()

    def iterate_fixed(self, niters, rho_f=1.0, dt=1e-3, parts=False):
        return self.lib.bbo_iterate_fixed(self.h, rho_f, dt, niters, int(parts))

    def gather_interior(self, aid):
        """Assemble the global interior field (Nz, Ny, Nx) from all blocks' ghosted Gcc arrays."""
        D = self.DOM
        out = np.zeros((D.zn, D.yn, D.xn))
        for r in range(self.nblocks):
            d = self.dom(r)
            a = self.array(r, aid)
            if aid in _GRID_OF:
                a = a[1:-1, 1:-1, 1:-1]
            i0, j0, k0 = d.Gcc.get("is") - 1, d.Gcc.get("js") - 1, d.Gcc.get("ks") - 1
            out[k0:k0 + d.zn, j0:j0 + d.yn, i0:i0 + d.xn] = a
        return out


def single_block_domain(extent, cells, bc, omp=False):
    """(DOM, dom) of a 1 x 1 x 1 decomposition filled by the oracle's OWN restatement of domain_fill (bbo_domain_fill,
    src/domain.c:918-1486) without allocating any field: what bench.py's reference arm hands to oracle/_ref so that the
    product library is not even mapped into that process."""
    lib = load(omp)
    D = C.POINTER(DomStruct)
    lib.bbo_domain_fill.argtypes = [D, D, C.POINTER(PressureBC)]
    lib.bbo_domain_fill.restype = None
    DOM, dom = DomStruct(), DomStruct()
    for d in (DOM, dom):
        d.xs, d.xe, d.ys, d.ye, d.zs, d.ze = [float(v) for v in extent]
        d.xn, d.yn, d.zn = [int(v) for v in cells]
    DOM.In = DOM.Jn = DOM.Kn = 1
    dom.I = dom.J = dom.K = 0
    pbc = PressureBC(*bc)
    lib.bbo_domain_fill(C.byref(DOM), C.byref(dom), C.byref(pbc))
    return DOM, dom, pbc
