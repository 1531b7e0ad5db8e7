"""Restart files as fixtures: bb_restart_read parses the per-rank binary file Bluebottle's
out_restart writes (src/domain.c:3005-3092).  The test writes files in that byte layout with numpy -- header scalars,
seven arrays per velocity component, p/phi/p0, phase/phase_shell, the three flag arrays, nparts_subdom, then trailing
bytes standing in for the particle structs -- and reads them back through the C ABI.  CPU only; the GPU replay test
(solve from a restart file == solve from the arrays) is in tests/test_gpu_epilogue.py."""
import os
import struct

import numpy as np
import pytest

import bbpcg
from bbpcg.grid import BC_SETS, grid_shape
from bbpcg.solver import RESTART_FIELDS


def write_restart(path, dom, fields, ttime=0.25, dt0=1e-3, dt=2e-3, stepnum=7, trailing=b"\x01" * 123, nparts=3):
    """out_restart, src/domain.c:3026-3078, field for field"""
    rng = np.random.default_rng(99)
    with open(path, "wb") as f:
        f.write(struct.pack("<dddiiddd", ttime, dt0, dt, stepnum, 11, 0.5, 0.6, 0.7))
        for vel, star, grid in (("u", "u_star", "Gfx"), ("v", "v_star", "Gfy"), ("w", "w_star", "Gfz")):
            shape = grid_shape(dom, grid)
            f.write(np.ascontiguousarray(fields[vel], dtype=np.float64).tobytes())
            for _ in range(5):                                    # X0, diff0, conv0, diff, conv: not needed by this path
                f.write(rng.standard_normal(shape).tobytes())
            f.write(np.ascontiguousarray(fields[star], dtype=np.float64).tobytes())
        for k in ("p", "phi", "p0"):
            f.write(np.ascontiguousarray(fields[k], dtype=np.float64).tobytes())
        for k in ("phase", "phase_shell", "flag_u", "flag_v", "flag_w"):
            f.write(np.ascontiguousarray(fields[k], dtype=np.int32).tobytes())
        f.write(struct.pack("<i", nparts))
        f.write(trailing)


def _fields(dom, seed):
    rng = np.random.default_rng(seed)
    out = {}
    for k, (grid, dt) in RESTART_FIELDS.items():
        shape = grid_shape(dom, grid)
        out[k] = rng.standard_normal(shape) if dt == np.float64 else rng.integers(-1, 3, size=shape, dtype=np.int32)
    return out


@pytest.mark.parametrize("cells,blocks,rank", [((6, 5, 4), (1, 1, 1), 0), ((8, 6, 6), (2, 1, 3), 4)])
def test_restart_round_trip(tmp_path, cells, blocks, rank):
    dec = bbpcg.Decomposition.uniform((0., 1., 0., 1., 0., 1.), cells, blocks, BC_SETS["duct"])
    dom = dec.doms[rank]
    fields = _fields(dom, 3 + rank)
    path = bbpcg.restart_path(str(tmp_path), rank, dec.nranks)
    write_restart(path, dom, fields)
    got = bbpcg.read_restart(path, dom)
    assert (got["ttime"], got["dt0"], got["dt"], got["stepnum"], got["rec_vtk_stepnum_out"]) == (0.25, 1e-3, 2e-3, 7, 11)
    assert (got["rec_cgns_flow_ttime_out"], got["rec_cgns_part_ttime_out"], got["rec_vtk_ttime_out"]) == (0.5, 0.6, 0.7)
    assert got["nparts_subdom"] == 3
    for k in RESTART_FIELDS:
        assert got[k].dtype == fields[k].dtype and np.array_equal(got[k], fields[k]), k


def test_restart_file_names_follow_out_restart(tmp_path):
    """restart.config-%0*d with floor(log10(S3 - 1)) + 1 digits (src/domain.c:3008-3017)"""
    d = str(tmp_path)
    assert os.path.basename(bbpcg.restart_path(d, 0, 1)) == "restart.config-0"
    assert os.path.basename(bbpcg.restart_path(d, 3, 8)) == "restart.config-3"
    assert os.path.basename(bbpcg.restart_path(d, 9, 10)) == "restart.config-9"
    assert os.path.basename(bbpcg.restart_path(d, 3, 11)) == "restart.config-03"
    assert os.path.basename(bbpcg.restart_path(d, 12, 16)) == "restart.config-12"
    assert os.path.basename(bbpcg.restart_path(d, 7, 101)) == "restart.config-007"
    with pytest.raises(RuntimeError):
        bbpcg.restart_path(d, 8, 8)


def test_restart_errors(tmp_path):
    dec = bbpcg.Decomposition.uniform((0., 1., 0., 1., 0., 1.), (6, 5, 4), (1, 1, 1), BC_SETS["duct"])
    dom = dec.doms[0]
    with pytest.raises(RuntimeError, match="could not be opened"):
        bbpcg.read_restart(str(tmp_path / "missing"), dom)
    path = str(tmp_path / "restart.config-0")
    write_restart(path, dom, _fields(dom, 1))
    size = os.path.getsize(path)
    with open(path, "rb") as f:
        blob = f.read()
    with open(path, "wb") as f:                                  # cut inside the flag arrays
        f.write(blob[: size - 123 - 4 - 200])
    with pytest.raises(RuntimeError, match="shorter than a restart file"):
        bbpcg.read_restart(path, dom)
    # a file written for another block size is refused or at least never over-read: reading with a LARGER block fails
    big = bbpcg.Decomposition.uniform((0., 1., 0., 1., 0., 1.), (12, 10, 8), (1, 1, 1), BC_SETS["duct"]).doms[0]
    with open(path, "wb") as f:
        f.write(blob)
    with pytest.raises(RuntimeError, match="shorter than a restart file"):
        bbpcg.read_restart(path, big)
