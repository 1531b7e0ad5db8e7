"""Host-side mirror of the reference's operator interface for the pressure-Poisson path.

`Decomposition` = domain_read_input (decomp part) + domain_fill  (src/domain.c:91-160,918-1486)
`PoissonSolver.init_jacobi_preconditioner` = cuda_PP_init_jacobi_preconditioner (src/cuda_solver.cu:31-36)
`PoissonSolver.PP_cg / PP_cg_noparts`      = cuda_PP_cg / cuda_PP_cg_noparts (src/cuda_solver.cu:38-300,573-761)
`PoissonSolver.exchange_Gcc`               = mpi_cuda_exchange_Gcc (src/mpi_comm.c:257-315)
`PoissonSolver.exchange_Gfx/Gfy/Gfz`       = mpi_cuda_exchange_Gfx/_Gfy/_Gfz (src/mpi_comm.c:317-405)
`PoissonSolver.solvability`                = cuda_solvability (src/cuda_bluebottle.cu:2313-2492)
`PoissonSolver.dom_BC_p`                   = cuda_dom_BC_p (src/cuda_bluebottle.cu:2536-2589)
`PoissonSolver.project / update_p`         = cuda_project / cuda_update_p (src/cuda_bluebottle.cu:2495-2534)
`PoissonSolver.epilogue`                   = the sequence src/bluebottle.c:233-256 runs on phi, fused

Array arguments are torch CUDA tensors in the reference's ghosted layouts (grid.grid_shape).
"""
import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import lib as L
from .grid import DomStruct, PressureBC, grid_shape


@dataclass
class SolveResult:
    status: str
    niter: int
    resid: float
    sp_rhs: float
    sp_rq0: float
    ms_setup: float
    ms_iter: float
    ms_total: float
    launches: int


class Decomposition:
    """DOM + dom[] for an In x Jn x Kn block decomposition (one block per rank/GPU)."""

    def __init__(self, DOM, doms, bc, params=None):
        self.DOM, self.doms, self.bc, self.params = DOM, doms, bc, params

    @property
    def nranks(self):
        return self.DOM.In * self.DOM.Jn * self.DOM.Kn

    @classmethod
    def uniform(cls, extent, cells, blocks=(1, 1, 1), bc=(0,) * 6):
        """Equal splits, as tools/src/decomp_reader.c writes them."""
        lib = L.load_library()
        DOM = DomStruct()
        DOM.xs, DOM.xe, DOM.ys, DOM.ye, DOM.zs, DOM.ze = [float(v) for v in extent]
        DOM.xn, DOM.yn, DOM.zn = cells
        DOM.In, DOM.Jn, DOM.Kn = blocks
        n = blocks[0] * blocks[1] * blocks[2]
        doms = (DomStruct * n)()
        pbc = PressureBC(*bc)
        L.check(lib.bb_domain_split(C.byref(DOM), doms), "bb_domain_split")
        L.check(lib.bb_domain_fill(C.byref(DOM), doms, C.byref(pbc)), "bb_domain_fill")
        return cls(DOM, doms, pbc)

    @classmethod
    def from_files(cls, flow_config, decomp_config):
        """flow.config + decomp.config, src/domain.c:72-160."""
        lib = L.load_library()
        DOM, pbc, fp = DomStruct(), PressureBC(), L.FlowParams()
        ptr = C.POINTER(DomStruct)()
        L.check(lib.bb_domain_read(flow_config.encode(), decomp_config.encode(), C.byref(DOM), C.byref(ptr),
                                   C.byref(pbc), C.byref(fp)), "bb_domain_read")
        n = DOM.In * DOM.Jn * DOM.Kn
        doms = (DomStruct * n)()
        C.memmove(doms, ptr, C.sizeof(DomStruct) * n)
        lib.bb_domain_free(ptr)
        return cls(DOM, doms, pbc, {"rho_f": fp.rho_f, "pp_residual": fp.pp_residual, "pp_max_iter": fp.pp_max_iter})

    def write_decomp(self, path, prec=2):
        L.check(L.load_library().bb_domain_write_decomp(path.encode(), C.byref(self.DOM), self.doms, prec),
                "bb_domain_write_decomp")


RESTART_FIELDS = {"u": ("Gfx", np.float64), "v": ("Gfy", np.float64), "w": ("Gfz", np.float64),
                  "u_star": ("Gfx", np.float64), "v_star": ("Gfy", np.float64), "w_star": ("Gfz", np.float64),
                  "p": ("Gcc", np.float64), "phi": ("Gcc", np.float64), "p0": ("Gcc", np.float64),
                  "phase": ("Gcc", np.int32), "phase_shell": ("Gcc", np.int32),
                  "flag_u": ("Gfx", np.int32), "flag_v": ("Gfy", np.int32), "flag_w": ("Gfz", np.int32)}


def restart_path(directory, rank, nranks):
    """<dir>/restart.config-<rank>, zero-padded like out_restart (src/domain.c:3008-3017)"""
    buf = C.create_string_buffer(4096)
    L.check(L.load_library().bb_restart_path(buf, 4096, directory.encode(), rank, nranks), "bb_restart_path")
    return buf.value.decode()


def read_restart(path, dom):
    """One rank's Bluebottle restart file (out_restart, src/domain.c:3005-3092) -> dict of numpy arrays in the
    reference's ghosted layouts + the header scalars.  Host only."""
    lib = L.load_library()
    r = L.Restart()
    L.check(lib.bb_restart_read(path.encode(), C.byref(dom), C.byref(r)), "bb_restart_read")
    out = {k: getattr(r, k) for k in ("ttime", "dt0", "dt", "stepnum", "rec_vtk_stepnum_out", "rec_cgns_flow_ttime_out",
                                      "rec_cgns_part_ttime_out", "rec_vtk_ttime_out", "nparts_subdom")}
    for k, (grid, dt) in RESTART_FIELDS.items():
        shape = grid_shape(dom, grid)
        n = int(np.prod(shape))
        ct = C.c_double if dt == np.float64 else C.c_int
        out[k] = np.frombuffer((ct * n).from_address(getattr(r, k)), dtype=dt).reshape(shape).copy()
    lib.bb_restart_free(C.byref(r))
    return out


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


# head of one rank's attach record (struct ExportBlob in csrc/bbpcg_solver.cu): magic, rank,
# in, jn, kn, device, pid; the rest (arena pointer, size, CUDA IPC handle) is opaque here
RECORD_FMT = "<Iiiiiiq"
RECORD_MAGIC = 0xBB9C6001


def parse_record(blob):
    import struct
    magic, rank, in_, jn, kn, device, pid = struct.unpack_from(RECORD_FMT, blob)
    return {"magic": magic, "rank": rank, "in": in_, "jn": jn, "kn": kn, "device": device, "pid": pid}


def gather_records(mine, _shuffle_for_test=False):
    """All-gather the ranks' attach records through torch.distributed (what MPI_Allgather does
    in the Bluebottle host) and check they arrive in rank order.  Works on any
    backend: the records are host bytes."""
    import torch.distributed as dist
    n = dist.get_world_size()
    blobs = [None] * n
    dist.all_gather_object(blobs, bytes(mine))
    if _shuffle_for_test:
        blobs = blobs[1:] + blobs[:1]
    for r, b in enumerate(blobs):
        info = parse_record(b)
        if len(b) != L.BLOB_BYTES or info["magic"] != RECORD_MAGIC or info["rank"] != r:
            raise RuntimeError("attach record %d is not rank %d's export (got rank %d)" % (r, r, info["rank"]))
    return blobs


class PoissonSolver:
    """One rank's solver object (one per GPU)."""

    def __init__(self, decomp, rank=0, device=None):
        import torch
        self.torch = torch
        self.lib = L.load_library()
        if not torch.cuda.is_available():
            raise RuntimeError("bbpcg needs a CUDA device: there is no CPU path")
        self.decomp, self.rank = decomp, rank
        self.dom = decomp.doms[rank]
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        self.h = C.c_void_p()
        L.check(self.lib.bbpcg_create(C.byref(self.h), C.byref(self.dom), C.byref(decomp.DOM), C.byref(decomp.bc),
                                      self.device.index), "bbpcg_create")
        self._keep = []

    def close(self):
        if getattr(self, "h", None):
            self.lib.bbpcg_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- multi-GPU attach ------------------------------------------------------------------
    def comm_export(self):
        buf = C.create_string_buffer(L.BLOB_BYTES)
        L.check(self.lib.bbpcg_comm_export(self.h, buf), "bbpcg_comm_export")
        return buf.raw

    def comm_import(self, blobs):
        allb = b"".join(blobs)
        assert len(allb) == L.BLOB_BYTES * len(blobs)
        L.check(self.lib.bbpcg_comm_import(self.h, allb, len(blobs)), "bbpcg_comm_import")

    def comm_init_torch(self):
        """Exchange the attach records through torch.distributed (replaces the MPI window set-up)."""
        import torch.distributed as dist
        n = dist.get_world_size()
        if n == 1:
            return
        self.comm_import(gather_records(self.comm_export()))
        dist.barrier()

    # ---- helpers ---------------------------------------------------------------------------
    def _sync_caller_stream(self):
        """The library works on its own stream: the caller's torch work on the arrays must be
        complete first.  Only the caller's stream is waited for -- a device-wide synchronize would
        also wait for OTHER ranks' collective kernels when several ranks share one GPU."""
        self.torch.cuda.current_stream(self.device).synchronize()

    def empty(self, grid, dtype=None):
        t = self.torch
        return t.zeros(grid_shape(self.dom, grid), dtype=dtype or t.float64, device=self.device)

    def to_device(self, arr):
        return self.torch.from_numpy(np.ascontiguousarray(arr)).to(self.device)

    def set_option(self, key, value):
        L.check(self.lib.bbpcg_set_option(self.h, key.encode(), int(value)), "bbpcg_set_option")

    def info(self, key):
        return self.lib.bbpcg_get_info(self.h, key.encode())

    # ---- the reference's entry points ----------------------------------------------------------
    def init_jacobi_preconditioner(self, flag_u, flag_v, flag_w, phase=None):
        self._sync_caller_stream()
        L.check(self.lib.bbpcg_set_coefficients(self.h, _ptr(flag_u), _ptr(flag_v), _ptr(flag_w), _ptr(phase)),
                "bbpcg_set_coefficients")

    def build_cages(self, parts_xyzr, flag_u, flag_v, flag_w, phase=None, phase_shell=None, NPARTS=None):
        """cuda_build_cages() + the mask digestion of cuda_PP_init_jacobi_preconditioner(): parts_xyzr is a device tensor
        [nparts, 4] of (x, y, z, r) -- or None -- in global coordinates; fills the five int arrays and the solver's masks"""
        n = 0 if parts_xyzr is None else int(parts_xyzr.shape[0])
        NPARTS = n if NPARTS is None else NPARTS
        view = L.PartsView()
        if n:
            assert parts_xyzr.dtype == self.torch.float64 and parts_xyzr.is_contiguous() and parts_xyzr.shape[1] == 4
            view.base, view.stride = parts_xyzr.data_ptr(), 32
            view.off_x, view.off_y, view.off_z, view.off_r = 0, 8, 16, 24
        self._sync_caller_stream()
        L.check(self.lib.bbpcg_build_cages(self.h, NPARTS, n, C.byref(view), _ptr(flag_u), _ptr(flag_v), _ptr(flag_w),
                                           _ptr(phase) if phase is not None else None,
                                           _ptr(phase_shell) if phase_shell is not None else None), "bbpcg_build_cages")
        self._has_phase = NPARTS > 0

    def _solve(self, u_star, v_star, w_star, rhs_p, phi, rho_f, dt, pp_residual, pp_max_iter, use_phase,
               phase=None, phase_shell=None, fixed_iters=0):
        a = L.SolveArgs()
        a.u_star, a.v_star, a.w_star = u_star.data_ptr(), v_star.data_ptr(), w_star.data_ptr()
        a.rhs_p, a.phi = rhs_p.data_ptr(), phi.data_ptr()
        a.phase = phase.data_ptr() if phase is not None else None
        a.phase_shell = phase_shell.data_ptr() if phase_shell is not None else None
        a.rho_f, a.dt, a.pp_residual, a.pp_max_iter = rho_f, dt, pp_residual, pp_max_iter
        a.use_phase, a.fixed_iters = int(use_phase), int(fixed_iters)
        res = L.Result()
        self._sync_caller_stream()
        L.check(self.lib.bbpcg_solve(self.h, C.byref(a), C.byref(res)), "bbpcg_solve")
        return SolveResult(L.STATUS.get(res.status, str(res.status)), res.niter, res.resid, res.sp_rhs, res.sp_rq0,
                           res.ms_setup, res.ms_iter, res.ms_total, res.launches)

    def PP_cg_noparts(self, u_star, v_star, w_star, rhs_p, phi, rho_f=1.0, dt=1e-3, pp_residual=1e-6,
                      pp_max_iter=2000, fixed_iters=0):
        return self._solve(u_star, v_star, w_star, rhs_p, phi, rho_f, dt, pp_residual, pp_max_iter, False,
                           fixed_iters=fixed_iters)

    def PP_cg(self, u_star, v_star, w_star, rhs_p, phi, phase, phase_shell, rho_f=1.0, dt=1e-3, pp_residual=1e-6,
              pp_max_iter=2000, fixed_iters=0):
        return self._solve(u_star, v_star, w_star, rhs_p, phi, rho_f, dt, pp_residual, pp_max_iter, True,
                           phase=phase, phase_shell=phase_shell, fixed_iters=fixed_iters)

    def solve_host(self, u_h, v_h, w_h, phi_h, rho_f=1.0, dt=1e-3, pp_residual=1e-6, pp_max_iter=2000, fixed_iters=0):
        """Host (pinned) tensors in, phi host tensor out: the end-to-end form."""
        res = L.Result()
        L.check(self.lib.bbpcg_solve_host(self.h, _ptr(u_h), _ptr(v_h), _ptr(w_h), _ptr(phi_h), rho_f, dt, pp_residual,
                                          pp_max_iter, fixed_iters, C.byref(res)), "bbpcg_solve_host")
        return SolveResult(L.STATUS.get(res.status, str(res.status)), res.niter, res.resid, res.sp_rhs, res.sp_rq0,
                           res.ms_setup, res.ms_iter, res.ms_total, res.launches)

    def history(self, cap=70000):
        out = np.zeros(cap)
        n = self.lib.bbpcg_history(self.h, out.ctypes.data_as(C.POINTER(C.c_double)), cap)
        return out[:n].copy()

    def exchange_Gcc(self, array):
        self._sync_caller_stream()
        L.check(self.lib.bbpcg_exchange_Gcc(self.h, _ptr(array)), "bbpcg_exchange_Gcc")

    def exchange(self, array, grid):
        """mpi_cuda_exchange_G{cc,fx,fy,fz}(array): grid in {"Gcc", "Gfx", "Gfy", "Gfz"}"""
        self._sync_caller_stream()
        L.check(self.lib.bbpcg_exchange(self.h, _ptr(array), L.GRID_CODE[grid]), "bbpcg_exchange")

    def exchange_Gfx(self, array):
        self.exchange(array, "Gfx")

    def exchange_Gfy(self, array):
        self.exchange(array, "Gfy")

    def exchange_Gfz(self, array):
        self.exchange(array, "Gfz")

    def solvability(self, u_star, v_star, w_star, out_plane="HOMOGENEOUS"):
        """cuda_solvability(): remove the net boundary flux of u* from the outflow plane(s); returns eps[3]"""
        eps = (C.c_double * 3)()
        self._sync_caller_stream()
        L.check(self.lib.bbpcg_solvability(self.h, _ptr(u_star), _ptr(v_star), _ptr(w_star), L.OUT_PLANE[out_plane], eps),
                "bbpcg_solvability")
        return [eps[0], eps[1], eps[2]]

    def dom_BC_star(self, u_star, v_star, w_star, vbc):
        """cuda_dom_BC_star(): the velocity BC table on u*, v*, w* (in place); vbc: lib.VelocityBC or the dict its make() takes"""
        if not isinstance(vbc, L.VelocityBC):
            vbc = L.VelocityBC.make(vbc)
        self._sync_caller_stream()
        L.check(self.lib.bbpcg_dom_BC_star(self.h, _ptr(u_star), _ptr(v_star), _ptr(w_star), C.byref(vbc)), "bbpcg_dom_BC_star")

    def prologue(self, u_star, v_star, w_star, vbc, out_plane="HOMOGENEOUS"):
        """src/bluebottle.c:213-225 without particles: BC_star, exchanges, solvability, BC_star, exchanges; returns device ms"""
        if not isinstance(vbc, L.VelocityBC):
            vbc = L.VelocityBC.make(vbc)
        ms = C.c_double()
        self._sync_caller_stream()
        L.check(self.lib.bbpcg_prologue(self.h, _ptr(u_star), _ptr(v_star), _ptr(w_star), C.byref(vbc), L.OUT_PLANE[out_plane], C.byref(ms)),
                "bbpcg_prologue")
        return ms.value

    def dom_BC_p(self, array):
        self._sync_caller_stream()
        L.check(self.lib.bbpcg_dom_BC_p(self.h, _ptr(array)), "bbpcg_dom_BC_p")

    def epilogue(self, phi, u_star=None, v_star=None, w_star=None, flag_u=None, flag_v=None, flag_w=None,
                 u=None, v=None, w=None, p0=None, phase=None, p=None, rho_f=1.0, dt=1e-3, phi_ghosts_valid=False):
        """[exchange_Gcc(phi); dom_BC_p(phi);] cuda_project; cuda_update_p.  u=None skips the projection,
        p=None the pressure update.  Returns the device milliseconds of the call."""
        a = L.EpilogueArgs()
        for k, t in (("u_star", u_star), ("v_star", v_star), ("w_star", w_star), ("flag_u", flag_u), ("flag_v", flag_v),
                     ("flag_w", flag_w), ("phi", phi), ("u", u), ("v", v), ("w", w), ("p0", p0), ("phase", phase), ("p", p)):
            setattr(a, k, None if t is None else t.data_ptr())
        a.rho_f, a.dt, a.phi_ghosts_valid = rho_f, dt, int(phi_ghosts_valid)
        ms = C.c_double()
        self._sync_caller_stream()
        L.check(self.lib.bbpcg_epilogue(self.h, C.byref(a), C.byref(ms)), "bbpcg_epilogue")
        return ms.value

    def project(self, u_star, v_star, w_star, phi, flag_u, flag_v, flag_w, u, v, w, rho_f=1.0, dt=1e-3):
        """cuda_project(): phi's ghost faces must be current (exchange_Gcc + dom_BC_p), as in bluebottle.c:233-237."""
        return self.epilogue(phi, u_star, v_star, w_star, flag_u, flag_v, flag_w, u, v, w, rho_f=rho_f, dt=dt,
                             phi_ghosts_valid=True)

    def update_p(self, p0, phi, phase, p):
        """cuda_update_p(): p = (phase < 0)(p0 + phi) - global mean."""
        return self.epilogue(phi, p0=p0, phase=phase, p=p, phi_ghosts_valid=True)

    def rhs(self, u_star, v_star, w_star, rhs_p, rho_f=1.0, dt=1e-3):
        self._sync_caller_stream()
        L.check(self.lib.bbpcg_rhs(self.h, _ptr(u_star), _ptr(v_star), _ptr(w_star), rho_f, dt, _ptr(rhs_p)), "bbpcg_rhs")

    def spmv(self, src_s3b, use_phase=False):
        d = self.dom
        out = self.torch.zeros((d.Gcc.get("kn"), d.Gcc.get("jn"), d.Gcc.get("in")), dtype=self.torch.float64,
                               device=self.device)
        self._sync_caller_stream()
        L.check(self.lib.bbpcg_spmv(self.h, _ptr(src_s3b), _ptr(out), int(use_phase)), "bbpcg_spmv")
        return out
