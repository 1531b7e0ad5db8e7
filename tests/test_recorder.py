"""The product's record-file writers (csrc/bb_recorder.c) against a restatement of the reference's fprintf formats:
recorder_PP_init / recorder_PP (src/recorder.c:157-221) and the timed pair (:223-336).  Host C, no GPU."""
import ctypes as C
import os

import bbpcg
from bbpcg import lib as L

SEG = ["spmv time (s)", "ip1 time (s)", "ar1 time (s)", "up1 time (s)", "ip2 time (s)", "ar2 time (s)", "up2 time (s)", "mpi time (s)"]


def _head(timed):
    s = "%-12s%-15s%-15s%-8s%-15s" % ("stepnum", "ttime", "dt", "niter", "resid")      # recorder.c:179-183
    if not timed:
        return s + "%-15s" % "time (s)"                                                 # :184
    return s + "%-16s" % "Total time (s)" + "".join("%-16s" % t for t in SEG)           # :249-257


def _line(stepnum, ttime, dt, niter, resid, etime, seg=None):
    s = "\n" + "%-12d%-15e%-15e%-8d%-15e" % (stepnum, ttime, dt, niter, resid)          # recorder.c:208-213
    if seg is None:
        return s + "%-15e" % etime                                                      # :215
    return s + "%-16e" % etime + "".join("%-16e" % v for v in seg)                      # :317-325


def test_recorder_PP_lines_are_byte_identical_to_the_reference_format(tmp_path):
    lib = bbpcg.load_library()
    root = str(tmp_path).encode()
    rows = [(1, 1e-3, 1e-3, 313, 9.87654321e-7, 0.5123), (2, 2e-3, 1e-3, 0, 0., 1.25e-4), (123456, 12.5, 2.5e-4, 2001, 3.3e-2, 77.)]
    for r in rows:                                        # no init call: the first line creates record/ and the header (:201-204)
        assert lib.bb_recorder_PP(root, b"solver_expd.rec", *r) == 0
    got = open(tmp_path / "record" / "solver_expd.rec").read()
    assert got == _head(False) + "".join(_line(*r) for r in rows)
    assert oct(os.stat(tmp_path / "record").st_mode & 0o777) == "0o700"                 # mkdir(buf, 0700), :166
    assert lib.bb_recorder_PP_init(root, b"solver_expd.rec") == 0                       # init truncates (fopen "w", :172)
    assert open(tmp_path / "record" / "solver_expd.rec").read() == _head(False)


def test_recorder_PP_timed_lines(tmp_path):
    lib = bbpcg.load_library()
    root = str(tmp_path).encode()
    seg = [0.11, 0., 0., 0.07, 0., 0., 0., 1e-9]
    arr = (C.c_double * 8)(*seg)
    assert lib.bb_recorder_PP_init_timed(root, b"solver_expd_timed.rec") == 0
    assert lib.bb_recorder_PP_timed(root, b"solver_expd_timed.rec", 7, 0.25, 1e-4, 150, 8.8e-7, 0.2, arr) == 0
    assert lib.bb_recorder_PP_timed(root, b"solver_expd_timed.rec", 8, 0.26, 1e-4, 151, 8.7e-7, 0.21, arr) == 0
    got = open(tmp_path / "record" / "solver_expd_timed.rec").read()
    assert got == _head(True) + _line(7, 0.25, 1e-4, 150, 8.8e-7, 0.2, seg) + _line(8, 0.26, 1e-4, 151, 8.7e-7, 0.21, seg)


def test_recorder_errors(tmp_path):
    lib = bbpcg.load_library()
    assert lib.bb_recorder_PP(None, b"x.rec", 1, 0., 0., 0, 0., 0.) == -1              # BBPCG_EINVAL
    assert lib.bb_recorder_PP_timed(str(tmp_path).encode(), b"x.rec", 1, 0., 0., 0, 0., 0., None) == -1
    missing = str(tmp_path / "no" / "such" / "dir").encode()
    assert lib.bb_recorder_PP(missing, b"x.rec", 1, 0., 0., 0, 0., 0.) == -5           # BBPCG_EIO
    assert b"record" in lib.bbpcg_last_error()
