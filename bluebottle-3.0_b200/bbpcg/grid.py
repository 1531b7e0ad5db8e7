"""ctypes mirrors of the grid contract in include/bb_grid.h.

Field order and types restate the reference's ``grid_info`` (src/domain.h:52-95),
``dom_struct`` (src/domain.h:168-210) and the six pressure entries that open ``BC``
(src/bluebottle.h:663-668).  ``sizeof(DomStruct)`` must be 880 bytes.
"""
import ctypes as C

import numpy as np

PERIODIC, DIRICHLET, NEUMANN = 0, 1, 2        # src/bluebottle.h:218,230,242
PROC_NULL = -2                                # MPI_PROC_NULL (OpenMPI); any negative = none
DOM_BUF = 1                                   # src/bluebottle.h:141

_GI_FIELDS = (
    "is ie in isb ieb inb js je jn jsb jeb jnb ks ke kn ksb keb knb "
    "_is _ie _isb _ieb _js _je _jsb _jeb _ks _ke _ksb _keb "
    "s1 s1b s2 s2b s3 s3b s2_i s2_j s2_k s2b_i s2b_j s2b_k"
).split()


class GridInfo(C.Structure):
    # python keywords (is, in) cannot be attribute names in source, use getattr(g, "in")
    _fields_ = [(n, C.c_int) for n in _GI_FIELDS]

    def get(self, name):
        return getattr(self, name)


class DomStruct(C.Structure):
    _fields_ = [
        ("Gcc", GridInfo), ("Gfx", GridInfo), ("Gfy", GridInfo), ("Gfz", GridInfo),
        ("xs", C.c_double), ("xe", C.c_double), ("xl", C.c_double), ("xn", C.c_int), ("dx", C.c_double),
        ("ys", C.c_double), ("ye", C.c_double), ("yl", C.c_double), ("yn", C.c_int), ("dy", C.c_double),
        ("zs", C.c_double), ("ze", C.c_double), ("zl", C.c_double), ("zn", C.c_int), ("dz", C.c_double),
        ("rank", C.c_int),
        ("e", C.c_int), ("w", C.c_int), ("n", C.c_int), ("s", C.c_int), ("t", C.c_int), ("b", C.c_int),
        ("I", C.c_int), ("Is", C.c_int), ("Ie", C.c_int), ("In", C.c_int),
        ("J", C.c_int), ("Js", C.c_int), ("Je", C.c_int), ("Jn", C.c_int),
        ("K", C.c_int), ("Ks", C.c_int), ("Ke", C.c_int), ("Kn", C.c_int),
        ("S1", C.c_int), ("S2", C.c_int), ("S3", C.c_int),
    ]


assert C.sizeof(GridInfo) == 42 * 4
assert C.sizeof(DomStruct) == 880, C.sizeof(DomStruct)


class PressureBC(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("pW", "pE", "pS", "pN", "pB", "pT")]


# the five pressure-BC sets of the shipped examples (examples/*/flow.config:26-31)
BC_SETS = {
    "cavity":        (NEUMANN, NEUMANN, NEUMANN, NEUMANN, PERIODIC, PERIODIC),   # lid-driven-cavity
    "duct":          (PERIODIC, PERIODIC, NEUMANN, NEUMANN, NEUMANN, NEUMANN),   # pressure-driven-duct
    "channel":       (PERIODIC, PERIODIC, NEUMANN, NEUMANN, PERIODIC, PERIODIC), # channel
    "sedimentation": (PERIODIC, PERIODIC, PERIODIC, PERIODIC, NEUMANN, NEUMANN), # sedimentation
    "periodic":      (PERIODIC,) * 6,
    "box":           (NEUMANN,) * 6,
}


def grid_shape(dom, grid):
    """numpy shape (slowest..fastest) of a ghosted (s3b) array on `grid` in {Gcc,Gfx,Gfy,Gfz}.

    Gcc/Gfz: i fastest -> a[k, j, i];  Gfx: j fastest -> a[i, k, j];  Gfy: k fastest -> a[j, i, k]
    (index macros, src/bluebottle.h:70-73)."""
    g = getattr(dom, grid)
    inb, jnb, knb = g.get("inb"), g.get("jnb"), g.get("knb")
    if grid in ("Gcc", "Gfz"):
        return (knb, jnb, inb)
    if grid == "Gfx":
        return (inb, knb, jnb)
    if grid == "Gfy":
        return (jnb, inb, knb)
    raise ValueError(grid)


def as_ijk(arr, grid):
    """View of a ghosted array indexed [i, j, k] regardless of the grid's storage order."""
    if grid in ("Gcc", "Gfz"):
        return arr.transpose(2, 1, 0)
    if grid == "Gfx":
        return arr.transpose(0, 2, 1)
    if grid == "Gfy":
        return arr.transpose(1, 0, 2)
    raise ValueError(grid)


def interior(arr_gcc):
    """Interior (ghost-free) view of a Gcc s3b array shaped (knb, jnb, inb)."""
    return arr_gcc[1:-1, 1:-1, 1:-1]


def copy_dom(d):
    out = DomStruct()
    C.memmove(C.byref(out), C.byref(d), C.sizeof(DomStruct))
    return out
