"""CPU tests (no GPU): the C-ABI libraries load and export every symbol include/*.h declares;
the binary grid contract has the reference's size; the product fails loudly without a GPU."""
import ctypes as C
import os
import re
import subprocess

import pytest

import bbpcg
from bbpcg import lib as L
from bbpcg.grid import BC_SETS, DomStruct, GridInfo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INC = os.path.join(ROOT, "include")


def _declared(header):
    """names of the functions a header declares (C prototypes ending in ';')"""
    text = open(os.path.join(INC, header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"^\s*#.*$", "", text, flags=re.M)
    text = re.sub(r"typedef[^;{]*\{[^}]*\}[^;]*;", "", text, flags=re.S)      # struct typedefs
    text = re.sub(r"typedef[^;]*;", "", text)                                   # other typedefs (incl. fn pointers)
    return sorted(set(re.findall(r"\b([A-Za-z_]\w*)\s*\([^;{]*\)\s*;", text)) - {"static_assert", "_Static_assert"})


def _exported(path):
    out = subprocess.check_output(["nm", "-D", "--defined-only", path], text=True)
    return {ln.split()[-1] for ln in out.splitlines() if ln.strip()}


def _undefined(path):
    out = subprocess.check_output(["nm", "-D", "--undefined-only", path], text=True)
    return {ln.split()[-1] for ln in out.splitlines() if ln.strip()}


def test_headers_declare_what_the_binding_lists():
    assert _declared("bbpcg.h") == sorted(L.SYMBOLS)
    # bb_dropin_allgather is declared for the HOST to define (weak in the library)
    assert sorted(set(_declared("bb_dropin.h")) - {"bb_dropin_allgather"}) == sorted(L.DROPIN_SYMBOLS)


def test_libbbpcg_exports_every_declared_symbol():
    assert os.path.exists(L.LIB_PATH), "run __graft_entry__.build() first"
    exp = _exported(L.LIB_PATH)
    missing = [s for s in _declared("bbpcg.h") if s not in exp]
    assert not missing, missing
    lib = bbpcg.load_library()
    for s in L.SYMBOLS:
        assert getattr(lib, s) is not None
    assert lib.bbpcg_version().startswith(b"bbpcg")


def test_dropin_exports_the_reference_entry_points_and_imports_its_globals():
    assert os.path.exists(L.DROPIN_PATH)
    exp, und = _exported(L.DROPIN_PATH), _undefined(L.DROPIN_PATH)
    for s in L.DROPIN_SYMBOLS:
        assert s in exp, s
    # what the reference host program must provide (src/bluebottle.c:438-576, mpi_comm.c:26-27, particle.c:27-28)
    for g in ("dom", "DOM", "rank", "nprocs", "bc", "rho_f", "dt", "pp_residual", "pp_max_iter", "NPARTS", "nparts",
              "_u_star", "_v_star", "_w_star", "_flag_u", "_flag_v", "_flag_w", "_phase", "_phase_shell", "_rhs_p", "_phi",
              "cuda_part_BC_p", "recorder_PP",
              "_u", "_v", "_w", "_p", "_p0", "out_plane", "_parts"):  # epilogue / solvability / cage entry points
        assert g in und, g
    # private scratch of the reference solver is NOT touched
    for g in ("_invM", "_r_q", "_z_q", "_p_q", "_pb_q", "_Apb_q", "_dom"):
        assert g not in und, g
    # no MPI symbol anywhere
    assert not [s for s in und if s.startswith("MPI_") or s.startswith("ompi_")]


def test_product_does_not_link_or_import_the_oracle():
    und = _undefined(L.LIB_PATH) | _undefined(L.DROPIN_PATH)
    assert not [s for s in und if s.startswith("bbo_") or s.startswith("bbref_")]
    needed = subprocess.check_output(["readelf", "-d", L.LIB_PATH], text=True)
    assert "liboracle" not in needed and "libbbref" not in needed
    pkg = os.path.join(ROOT, "bluebottle-3.0_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".c", ".h")):
                txt = open(os.path.join(dp, f)).read()
                for needle in ("import oracle", "from oracle", "liboracle", "libbbref", "pcg_ref", "oracle/"):
                    assert needle not in txt, (os.path.join(dp, f), needle)


def test_part_struct_view_matches_the_reference_header(tmp_path):
    """the drop-in's cuda_build_cages reads x, y, z, r of the reference's part_struct through fixed offsets: re-derive them
    from the reference's own header whenever it is available (the authoring container)"""
    ref = "/root/reference/src"
    if not os.path.exists(os.path.join(ref, "particle.h")):
        pytest.skip("reference tree not present")
    src = tmp_path / "po.cu"
    src.write_text('#include <cstdio>\n#include <cstddef>\n#include "bluebottle.h"\n#include "particle.h"\n'
                   'int main(){ printf("%zu %zu %zu %zu %zu %zu\\n", sizeof(part_struct), offsetof(part_struct, r), offsetof(part_struct, x), '
                   'offsetof(part_struct, y), offsetof(part_struct, z), sizeof(BC)); }\n')
    exe = str(tmp_path / "po")
    subprocess.check_call(["nvcc", "-DDOUBLE", "-w", "-I", os.path.join(ROOT, "oracle", "stubs"), "-I", ref, "-o", exe, str(src)])
    size, off_r, off_x, off_y, off_z, size_bc = [int(v) for v in subprocess.check_output([exe], text=True).split()]
    txt = open(os.path.join(ROOT, "bluebottle-3.0_b200", "csrc", "bbpcg_dropin.cu")).read()
    got = {k: int(re.search(r"#define %s (\d+)" % k, txt).group(1)) for k in ("BB_PART_STRIDE", "BB_PART_OFF_R", "BB_PART_OFF_X", "BB_PART_OFF_Y", "BB_PART_OFF_Z")}
    assert got == {"BB_PART_STRIDE": size, "BB_PART_OFF_R": off_r, "BB_PART_OFF_X": off_x, "BB_PART_OFF_Y": off_y, "BB_PART_OFF_Z": off_z}
    assert size_bc == 648                       # include/bb_grid.h: bb_BC mirrors the reference's BC


def test_grid_contract_sizes():
    assert C.sizeof(GridInfo) == 42 * 4
    assert C.sizeof(DomStruct) == 880          # sizeof(dom_struct) in the reference build = 0x370
    assert DomStruct.xs.offset == 4 * 42 * 4   # 4 grid_info blocks first (src/domain.h:168-173)
    assert DomStruct.rank.offset == 4 * 42 * 4 + 3 * 40


def test_status_codes_match_header():
    text = open(os.path.join(INC, "bbpcg.h")).read()
    for name, val in (("CONVERGED", 0), ("TINY_RHS", 1), ("MAXITER", 2), ("NAN", 3)):
        assert re.search(r"#define\s+BBPCG_%s\s+%d\b" % (name, val), text)
    assert int(re.search(r"#define\s+BBPCG_BLOB_BYTES\s+(\d+)", text).group(1)) == L.BLOB_BYTES


def _no_gpu():
    import torch
    return not torch.cuda.is_available()


@pytest.mark.skipif(not _no_gpu(), reason="checks the behaviour WITHOUT a CUDA device")
def test_no_cpu_fallback_create_fails_loudly():
    """There is no CPU path: bbpcg_create reports BBPCG_ECUDA, the Python mirror raises."""
    dec = bbpcg.Decomposition.uniform((0, 1, 0, 1, 0, 1), (8, 8, 8), (1, 1, 1), BC_SETS["periodic"])
    lib = bbpcg.load_library()
    h = C.c_void_p()
    rc = lib.bbpcg_create(C.byref(h), C.byref(dec.doms[0]), C.byref(dec.DOM), C.byref(dec.bc), -1)
    assert rc == -2 and b"no CUDA device" in lib.bbpcg_last_error()
    with pytest.raises(RuntimeError, match="no CPU path"):
        bbpcg.PoissonSolver(dec, 0)


def test_missing_library_is_an_error_not_a_fallback(monkeypatch):
    monkeypatch.setattr(L, "_lib", None)
    monkeypatch.setattr(L, "LIB_PATH", "/nonexistent/libbbpcg.so")
    with pytest.raises(bbpcg.LibraryMissing):
        L.load_library()


def test_ctypes_mirrors_match_the_c_structs(tmp_path):
    """the Python mirror binds by struct layout: sizes and field offsets of every struct of bbpcg.h, from the C compiler"""
    src = tmp_path / "layout.c"
    fields = {
        "bbpcg_result": (L.Result, ["status", "niter", "resid", "sp_rhs", "sp_rq0", "ms_setup", "ms_iter", "ms_total", "launches"]),
        "bb_flow_params": (L.FlowParams, ["rho_f", "pp_residual", "pp_max_iter"]),
        "bbpcg_solve_args": (L.SolveArgs, ["u_star", "v_star", "w_star", "rhs_p", "phi", "phase", "phase_shell", "rho_f", "dt",
                                           "pp_residual", "pp_max_iter", "use_phase", "fixed_iters", "part_bc", "no_refine"]),
        "bbpcg_epilogue_args": (L.EpilogueArgs, ["u_star", "v_star", "w_star", "flag_u", "flag_v", "flag_w", "phi", "u", "v", "w",
                                                 "p0", "phase", "p", "rho_f", "dt", "phi_ghosts_valid"]),
        "bb_restart": (L.Restart, ["ttime", "dt0", "dt", "stepnum", "rec_vtk_stepnum_out", "rec_cgns_flow_ttime_out",
                                   "rec_cgns_part_ttime_out", "rec_vtk_ttime_out", "u", "v", "w", "u_star", "v_star", "w_star",
                                   "p", "phi", "p0", "phase", "phase_shell", "flag_u", "flag_v", "flag_w", "nparts_subdom"]),
        "bb_pressure_bc": (bbpcg.grid.PressureBC, ["pW", "pE", "pS", "pN", "pB", "pT"]),
    }
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "bbpcg.h"', "int main(void) {"]
    for name, (_, fl) in fields.items():
        lines.append('  printf("%s %%zu\\n", sizeof(%s));' % (name, name))
        for f in fl:
            lines.append('  printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (name, f, name, f))
    lines += ["  return 0;", "}"]
    src.write_text("\n".join(lines))
    exe = str(tmp_path / "layout")
    subprocess.check_call(["gcc", "-std=c99", "-I", INC, "-o", exe, str(src)])
    got = dict(ln.split() for ln in subprocess.check_output([exe], text=True).splitlines())
    for name, (ct, fl) in fields.items():
        assert C.sizeof(ct) == int(got[name]), name
        assert [f for f, _ in ct._fields_] == fl, name              # same fields, same order
        for f in fl:
            assert getattr(ct, f).offset == int(got["%s.%s" % (name, f)]), (name, f)
