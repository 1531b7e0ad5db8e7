"""ctypes declarations of include/bbpcg.h."""
import ctypes as C
import os

from .grid import DomStruct, PressureBC

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_DIR = os.path.join(os.path.dirname(_HERE), "lib")
LIB_PATH = os.environ.get("BBPCG_LIB_PATH") or os.path.join(LIB_DIR, "libbbpcg.so")   # BBPCG_LIB_PATH: a debug build of the SAME library (make trace)
DROPIN_PATH = os.path.join(LIB_DIR, "libbbpcg_dropin.so")

BLOB_BYTES = 256
OUT_PLANE = {"WEST": 0, "EAST": 1, "SOUTH": 2, "NORTH": 3, "BOTTOM": 4, "TOP": 5, "HOMOGENEOUS": 10}   # src/bluebottle.h:353-425
GRID_CODE = {"Gcc": 0, "Gfx": 1, "Gfy": 2, "Gfz": 3}      # BBPCG_GCC .. BBPCG_GFZ
OK = 0
STATUS = {0: "converged", 1: "tiny_rhs", 2: "max_iter", 3: "nan", 4: "comm_timeout"}


class LibraryMissing(RuntimeError):
    pass


class Result(C.Structure):
    _fields_ = [("status", C.c_int), ("niter", C.c_int), ("resid", C.c_double), ("sp_rhs", C.c_double),
                ("sp_rq0", C.c_double), ("ms_setup", C.c_double), ("ms_iter", C.c_double),
                ("ms_total", C.c_double), ("launches", C.c_longlong)]


class FlowParams(C.Structure):
    _fields_ = [("rho_f", C.c_double), ("pp_residual", C.c_double), ("pp_max_iter", C.c_int)]


PART_BC_FN = C.CFUNCTYPE(None)


class SolveArgs(C.Structure):
    _fields_ = [("u_star", C.c_void_p), ("v_star", C.c_void_p), ("w_star", C.c_void_p),
                ("rhs_p", C.c_void_p), ("phi", C.c_void_p), ("phase", C.c_void_p), ("phase_shell", C.c_void_p),
                ("rho_f", C.c_double), ("dt", C.c_double), ("pp_residual", C.c_double),
                ("pp_max_iter", C.c_int), ("use_phase", C.c_int), ("fixed_iters", C.c_int),
                ("part_bc", PART_BC_FN), ("no_refine", C.c_int)]


class EpilogueArgs(C.Structure):
    _fields_ = [("u_star", C.c_void_p), ("v_star", C.c_void_p), ("w_star", C.c_void_p),
                ("flag_u", C.c_void_p), ("flag_v", C.c_void_p), ("flag_w", C.c_void_p),
                ("phi", C.c_void_p), ("u", C.c_void_p), ("v", C.c_void_p), ("w", C.c_void_p),
                ("p0", C.c_void_p), ("phase", C.c_void_p), ("p", C.c_void_p),
                ("rho_f", C.c_double), ("dt", C.c_double), ("phi_ghosts_valid", C.c_int)]


class PartsView(C.Structure):
    """bbpcg_parts_view: a strided view of the rank's particle list (device memory)"""
    _fields_ = [("base", C.c_void_p), ("stride", C.c_size_t), ("off_x", C.c_size_t), ("off_y", C.c_size_t),
                ("off_z", C.c_size_t), ("off_r", C.c_size_t)]


class VelocityBC(C.Structure):
    """bb_velocity_bc: type[c][f], val[c][f]; component c = u, v, w; face f = W, E, S, N, B, T (the reference's order)"""
    _fields_ = [("type", (C.c_int * 6) * 3), ("val", (C.c_double * 6) * 3)]

    FACES = ("W", "E", "S", "N", "B", "T")

    @classmethod
    def make(cls, spec):
        """spec: {"uW": ("D", 1.5), "vN": "N", ...}; everything not named is PERIODIC (no action)"""
        out = cls()
        codes = {"P": 0, "D": 1, "N": 2, "PRECURSOR": 3}
        for key, v in (spec or {}).items():
            c, f = "uvw".index(key[0]), cls.FACES.index(key[1])
            kind, val = (v, 0.) if isinstance(v, str) else v
            out.type[c][f] = codes[kind]
            out.val[c][f] = float(val)
        return out

    def flat(self):
        return [self.type[c][f] for c in range(3) for f in range(6)], [self.val[c][f] for c in range(3) for f in range(6)]


class Restart(C.Structure):
    _fields_ = [("ttime", C.c_double), ("dt0", C.c_double), ("dt", C.c_double), ("stepnum", C.c_int),
                ("rec_vtk_stepnum_out", C.c_int), ("rec_cgns_flow_ttime_out", C.c_double),
                ("rec_cgns_part_ttime_out", C.c_double), ("rec_vtk_ttime_out", C.c_double),
                ("u", C.c_void_p), ("v", C.c_void_p), ("w", C.c_void_p),
                ("u_star", C.c_void_p), ("v_star", C.c_void_p), ("w_star", C.c_void_p),
                ("p", C.c_void_p), ("phi", C.c_void_p), ("p0", C.c_void_p),
                ("phase", C.c_void_p), ("phase_shell", C.c_void_p),
                ("flag_u", C.c_void_p), ("flag_v", C.c_void_p), ("flag_w", C.c_void_p), ("nparts_subdom", C.c_int)]


# every symbol include/bbpcg.h declares (checked by tests/test_abi.py)
SYMBOLS = [
    "bb_domain_read", "bb_domain_fill", "bb_domain_split", "bb_domain_write_decomp", "bb_domain_free",
    "bb_restart_path", "bb_restart_read", "bb_restart_free",
    "bb_recorder_PP_init", "bb_recorder_PP", "bb_recorder_PP_init_timed", "bb_recorder_PP_timed",
    "bbpcg_create", "bbpcg_destroy", "bbpcg_comm_export", "bbpcg_comm_import", "bbpcg_set_coefficients",
    "bbpcg_solve", "bbpcg_solve_host", "bbpcg_history", "bbpcg_exchange_Gcc", "bbpcg_rhs", "bbpcg_spmv",
    "bbpcg_set_option", "bbpcg_plan_zchunks", "bbpcg_get_info", "bbpcg_last_error", "bbpcg_version",
    "bbpcg_dom_BC_p", "bbpcg_epilogue", "bbpcg_exchange", "bbpcg_solvability", "bbpcg_dom_BC_star", "bbpcg_prologue", "bbpcg_build_cages",
]
DROPIN_SYMBOLS = ["cuda_PP_init_jacobi_preconditioner", "cuda_PP_cg", "cuda_PP_cg_noparts", "cuda_PP_cg_timed",
                  "mpi_cuda_exchange_Gcc", "mpi_cuda_exchange_Gfx", "mpi_cuda_exchange_Gfy", "mpi_cuda_exchange_Gfz", "cuda_solvability", "cuda_dom_BC_star", "cuda_build_cages", "cuda_dom_BC_p", "cuda_project", "cuda_update_p", "bbpcg_dropin_finalize"]

_lib = None


def load_library():
    """Load libbbpcg.so; raises LibraryMissing (never falls back to anything else)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LibraryMissing("%s not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                             "(there is no CPU fallback)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    vp, ip, dp = C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_double)
    D = C.POINTER(DomStruct)
    lib.bb_domain_read.argtypes = [C.c_char_p, C.c_char_p, D, C.POINTER(D), C.POINTER(PressureBC), C.POINTER(FlowParams)]
    lib.bb_domain_fill.argtypes = [D, D, C.POINTER(PressureBC)]
    lib.bb_domain_split.argtypes = [D, D]
    lib.bb_domain_write_decomp.argtypes = [C.c_char_p, D, D, C.c_int]
    lib.bb_domain_free.argtypes = [D]
    lib.bb_domain_free.restype = None
    lib.bb_restart_path.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_int, C.c_int]
    lib.bb_restart_read.argtypes = [C.c_char_p, D, C.POINTER(Restart)]
    lib.bb_restart_free.argtypes = [C.POINTER(Restart)]
    lib.bb_restart_free.restype = None
    lib.bb_recorder_PP_init.argtypes = [C.c_char_p, C.c_char_p]
    lib.bb_recorder_PP_init_timed.argtypes = [C.c_char_p, C.c_char_p]
    lib.bb_recorder_PP.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_double, C.c_double, C.c_int, C.c_double, C.c_double]
    lib.bb_recorder_PP_timed.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_double, C.c_double, C.c_int, C.c_double, C.c_double, dp]
    lib.bbpcg_plan_zchunks.argtypes = [C.c_int] * 9 + [ip, C.c_int, ip]
    lib.bbpcg_create.argtypes = [C.POINTER(vp), D, D, C.POINTER(PressureBC), C.c_int]
    lib.bbpcg_destroy.argtypes = [vp]
    lib.bbpcg_destroy.restype = None
    lib.bbpcg_comm_export.argtypes = [vp, vp]
    lib.bbpcg_comm_import.argtypes = [vp, vp, C.c_int]
    lib.bbpcg_set_coefficients.argtypes = [vp, vp, vp, vp, vp]
    lib.bbpcg_solve.argtypes = [vp, C.POINTER(SolveArgs), C.POINTER(Result)]
    lib.bbpcg_solve_host.argtypes = [vp, vp, vp, vp, vp, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, C.POINTER(Result)]
    lib.bbpcg_history.argtypes = [vp, dp, C.c_int]
    lib.bbpcg_exchange_Gcc.argtypes = [vp, vp]
    lib.bbpcg_rhs.argtypes = [vp, vp, vp, vp, C.c_double, C.c_double, vp]
    lib.bbpcg_spmv.argtypes = [vp, vp, vp, C.c_int]
    lib.bbpcg_exchange.argtypes = [vp, vp, C.c_int]
    lib.bbpcg_solvability.argtypes = [vp, vp, vp, vp, C.c_int, dp]
    lib.bbpcg_dom_BC_p.argtypes = [vp, vp]
    lib.bbpcg_build_cages.argtypes = [vp, C.c_int, C.c_int, C.POINTER(PartsView), vp, vp, vp, vp, vp]
    lib.bbpcg_dom_BC_star.argtypes = [vp, vp, vp, vp, C.POINTER(VelocityBC)]
    lib.bbpcg_prologue.argtypes = [vp, vp, vp, vp, C.POINTER(VelocityBC), C.c_int, dp]
    lib.bbpcg_epilogue.argtypes = [vp, C.POINTER(EpilogueArgs), dp]
    lib.bbpcg_set_option.argtypes = [vp, C.c_char_p, C.c_longlong]
    lib.bbpcg_get_info.argtypes = [vp, C.c_char_p]
    lib.bbpcg_get_info.restype = C.c_longlong
    lib.bbpcg_last_error.restype = C.c_char_p
    lib.bbpcg_version.restype = C.c_char_p
    _lib = lib
    return lib


def check(rc, what):
    if rc != OK:
        raise RuntimeError("%s failed (%d): %s" % (what, rc, load_library().bbpcg_last_error().decode()))
