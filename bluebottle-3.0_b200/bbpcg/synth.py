"""Deterministic synthetic inputs for the pressure-Poisson path.

Everything is keyed on GLOBAL cell / face indices, so a field built block by block for any
decomposition is bit-identical to the same field built on one block: the projected
velocity u*, v*, w* = smooth sin/cos field + uniform noise in [-0.5, 0.5), with
wall-normal faces zeroed so that sum(rhs) = 0 (the solvability condition the reference
enforces in cuda_solvability, src/cuda_bluebottle.cu:2313-2492).

The noise is a counter-based hash (splitmix64 of the global face index) instead of a
sequential generator so that each rank can fill only its own block.
"""
import numpy as np

from .grid import PERIODIC, grid_shape

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(x):
    """Vectorised splitmix64 finaliser on uint64 arrays."""
    with np.errstate(over="ignore"):
        z = x + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return z


def hash_uniform(idx, stream):
    """uniform [-0.5, 0.5) from integer index array `idx` and an integer stream id."""
    z = splitmix64(idx.astype(np.uint64) + np.uint64(stream) * np.uint64(0x0123456789ABCDEF))
    return (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0) - 0.5


def _axis(dom, DOM, axis, staggered, periodic):
    """Per-axis helper: for local index l = 0..n+1(+1) returns (coordinate, global index, wall mask).

    Cell centres: x = (l - 0.5) dx + xs (particle_kernel.cu:1677-1679); faces: x = (l - 1) dx + xs.
    Global 0-based index = G.is - DOM_BUF + (l - 1); periodic face indices wrap modulo N so the
    two copies of the periodic face (and the duplicated block-boundary faces) agree."""
    n = getattr(dom, axis + "n")
    N = getattr(DOM, axis + "n")
    s = getattr(dom, axis + "s")
    d = getattr(dom, "d" + axis)
    gname = {"x": "Gfx", "y": "Gfy", "z": "Gfz"}[axis] if staggered else "Gcc"
    g0 = getattr(dom, gname).get({"x": "is", "y": "js", "z": "ks"}[axis]) - 1   # global index of local 1
    nb = n + 2 + (1 if staggered else 0)
    l = np.arange(nb)
    gi = g0 + (l - 1)
    if staggered:
        coord = (l - 1) * d + s
        wall = np.zeros(nb, dtype=bool) if periodic else ((gi == 0) | (gi == N))
        gi = np.where(gi == N, 0, gi) if periodic else gi
        valid = (l >= 1) & (l <= n + 1)
    else:
        coord = (l - 0.5) * d + s
        wall = np.zeros(nb, dtype=bool)
        valid = (l >= 1) & (l <= n)
    return coord, gi.astype(np.int64), wall, valid


def velocity_star(dom, DOM, bc, noise=1.0, seed=7):
    """u*, v*, w* for one block, in the reference's Gfx / Gfy / Gfz storage (ghosted, s3b).

    Returns three float64 arrays shaped grid_shape(dom, G).  Entries PP_rhs does not read
    (ghost rows, solver_kernel.cu:128-146) are left 0."""
    px = bc.pW == PERIODIC and bc.pE == PERIODIC
    py = bc.pS == PERIODIC and bc.pN == PERIODIC
    pz = bc.pB == PERIODIC and bc.pT == PERIODIC
    Lx, Ly, Lz = DOM.xe - DOM.xs, DOM.ye - DOM.ys, DOM.ze - DOM.zs
    Nx, Ny, Nz = DOM.xn, DOM.yn, DOM.zn
    out = []
    for comp, grid in enumerate(("Gfx", "Gfy", "Gfz")):
        sx, sy, sz = (comp == 0), (comp == 1), (comp == 2)
        X, GI, WX, VX = _axis(dom, DOM, "x", sx, px)
        Y, GJ, WY, VY = _axis(dom, DOM, "y", sy, py)
        Z, GK, WZ, VZ = _axis(dom, DOM, "z", sz, pz)
        ax = 2 * np.pi * (X - DOM.xs) / Lx
        ay = 2 * np.pi * (Y - DOM.ys) / Ly
        az = 2 * np.pi * (Z - DOM.zs) / Lz
        # component-specific smooth part
        if comp == 0:
            fx, fy, fz = np.sin(ax), np.cos(ay), np.cos(az)
        elif comp == 1:
            fx, fy, fz = np.cos(ax), np.sin(ay), np.cos(az)
        else:
            fx, fy, fz = np.cos(ax), np.cos(ay), np.sin(az)
        # broadcast in [i, j, k] then permute into the grid's storage order
        sI, sJ, sK = Nx + 1, Ny + 1, Nz + 1
        lin = (GI[:, None, None] + sI * (GJ[None, :, None] + sJ * GK[None, None, :]))
        f = fx[:, None, None] * fy[None, :, None] * fz[None, None, :]
        if noise != 0.0:
            f = f + noise * hash_uniform(lin, seed * 4 + comp)
        wall = (WX[:, None, None] if sx else False) | (WY[None, :, None] if sy else False) | \
               (WZ[None, None, :] if sz else False)
        valid = VX[:, None, None] & VY[None, :, None] & VZ[None, None, :]
        f = np.where(valid & ~wall, f, 0.0)
        if grid == "Gfx":
            a = f.transpose(0, 2, 1)       # [i, k, j]
        elif grid == "Gfy":
            a = f.transpose(1, 0, 2)       # [j, i, k]
        else:
            a = f.transpose(2, 1, 0)       # [k, j, i]
        a = np.ascontiguousarray(a, dtype=np.float64)
        assert a.shape == grid_shape(dom, grid), (a.shape, grid_shape(dom, grid))
        out.append(a)
    return out


def random_spheres(DOM, nparts, radius, seed=20240229, max_tries=200000):
    """Non-overlapping sphere centres by rejection sampling (config C4).
    Centres keep one radius + 2 cells clear of every domain face so cages never wrap."""
    rng = np.random.default_rng(seed)
    lo = np.array([DOM.xs, DOM.ys, DOM.zs]) + radius + 2 * max(DOM.dx, DOM.dy, DOM.dz)
    hi = np.array([DOM.xe, DOM.ye, DOM.ze]) - radius - 2 * max(DOM.dx, DOM.dy, DOM.dz)
    pts = np.empty((0, 3))
    tries = 0
    while len(pts) < nparts and tries < max_tries:
        cand = lo + (hi - lo) * rng.random(3)
        tries += 1
        if len(pts) == 0 or np.min(np.sum((pts - cand) ** 2, axis=1)) > (2.2 * radius) ** 2:
            pts = np.vstack([pts, cand])
    if len(pts) < nparts:
        raise RuntimeError("could not place %d spheres of radius %g" % (nparts, radius))
    return pts[:, 0].copy(), pts[:, 1].copy(), pts[:, 2].copy(), np.full(nparts, float(radius))


# ---- the same fields built with torch on any device (bench.py fills 512^3 blocks on the GPU) ----
def _i64(v):
    """two's-complement int64 value of an unsigned 64-bit constant"""
    v &= 0xFFFFFFFFFFFFFFFF
    return v - (1 << 64) if v >= (1 << 63) else v


def _lsr(t, n):
    """logical right shift of an int64 tensor holding uint64 bits"""
    return (t >> n) & ((1 << (64 - n)) - 1)


def hash_uniform_torch(idx, stream):
    """hash_uniform on an int64 torch tensor (wrap-around int64 arithmetic == uint64 arithmetic)."""
    import torch
    z = idx + _i64(stream * 0x0123456789ABCDEF) + _i64(0x9E3779B97F4A7C15)
    z = (z ^ _lsr(z, 30)) * _i64(0xBF58476D1CE4E5B9)
    z = (z ^ _lsr(z, 27)) * _i64(0x94D049BB133111EB)
    z = z ^ _lsr(z, 31)
    return _lsr(z, 11).to(torch.float64) * (1.0 / 9007199254740992.0) - 0.5


def velocity_star_torch(dom, DOM, bc, device, noise=1.0, seed=7):
    """velocity_star() evaluated with torch on `device`; bit-identical to the numpy version (the
    per-axis sin/cos factors are computed on the host in both, the rest is the same sequence of
    IEEE operations)."""
    import torch
    px = bc.pW == PERIODIC and bc.pE == PERIODIC
    py = bc.pS == PERIODIC and bc.pN == PERIODIC
    pz = bc.pB == PERIODIC and bc.pT == PERIODIC
    Lx, Ly, Lz = DOM.xe - DOM.xs, DOM.ye - DOM.ys, DOM.ze - DOM.zs
    Nx, Ny = DOM.xn, DOM.yn
    out = []
    # storage order (slowest..fastest) of each grid expressed as a permutation of (i, j, k)
    order = {"Gfx": (0, 2, 1), "Gfy": (1, 0, 2), "Gfz": (2, 1, 0)}
    for comp, grid in enumerate(("Gfx", "Gfy", "Gfz")):
        sx, sy, sz = (comp == 0), (comp == 1), (comp == 2)
        X, GI, WX, VX = _axis(dom, DOM, "x", sx, px)
        Y, GJ, WY, VY = _axis(dom, DOM, "y", sy, py)
        Z, GK, WZ, VZ = _axis(dom, DOM, "z", sz, pz)
        ax = 2 * np.pi * (X - DOM.xs) / Lx
        ay = 2 * np.pi * (Y - DOM.ys) / Ly
        az = 2 * np.pi * (Z - DOM.zs) / Lz
        if comp == 0:
            fx, fy, fz = np.sin(ax), np.cos(ay), np.cos(az)
        elif comp == 1:
            fx, fy, fz = np.cos(ax), np.sin(ay), np.cos(az)
        else:
            fx, fy, fz = np.cos(ax), np.cos(ay), np.sin(az)
        perm = order[grid]

        def shaped(v, axis, dtype=None):
            """1-D per-axis array -> broadcastable tensor in this grid's storage order"""
            shp = [1, 1, 1]
            shp[perm.index(axis)] = len(v)
            t = torch.from_numpy(np.ascontiguousarray(v))
            if dtype is not None:
                t = t.to(dtype)
            return t.to(device).reshape(shp)

        sI, sJ = Nx + 1, Ny + 1
        f = (shaped(fx, 0) * shaped(fy, 1)) * shaped(fz, 2)
        if noise != 0.0:
            lin = shaped(GI, 0) + sI * (shaped(GJ, 1) + sJ * shaped(GK, 2))
            f = f + noise * hash_uniform_torch(lin, seed * 4 + comp)
            del lin
        keep = (shaped(VX & ~(WX if sx else np.zeros_like(WX)), 0) & shaped(VY & ~(WY if sy else np.zeros_like(WY)), 1)
                & shaped(VZ & ~(WZ if sz else np.zeros_like(WZ)), 2))
        f = torch.where(keep, f, torch.zeros((), dtype=torch.float64, device=device)).contiguous()
        assert tuple(f.shape) == grid_shape(dom, grid), (tuple(f.shape), grid_shape(dom, grid))
        out.append(f)
    return out


def flags_noparts_torch(dom, DOM, bc, device):
    """flag_u, flag_v, flag_w (int32, Gfx/Gfy/Gfz s3b) for a particle-free block: 1 everywhere, 0 on
    external wall faces when BOTH sides of that direction are non-periodic and the block touches the
    wall (cuda_build_cages, src/cuda_particle.cu:1605-1639; kernels src/particle_kernel.cu:542-576)."""
    import torch
    fu = torch.ones(grid_shape(dom, "Gfx"), dtype=torch.int32, device=device)   # [i, k, j]
    fv = torch.ones(grid_shape(dom, "Gfy"), dtype=torch.int32, device=device)   # [j, i, k]
    fw = torch.ones(grid_shape(dom, "Gfz"), dtype=torch.int32, device=device)   # [k, j, i]
    if bc.pW != PERIODIC and bc.pE != PERIODIC:
        if dom.I == DOM.Is:
            fu[dom.Gfx.get("_is")] = 0
        if dom.I == DOM.Ie:
            fu[dom.Gfx.get("_ie")] = 0
    if bc.pS != PERIODIC and bc.pN != PERIODIC:
        if dom.J == DOM.Js:
            fv[dom.Gfy.get("_js")] = 0
        if dom.J == DOM.Je:
            fv[dom.Gfy.get("_je")] = 0
    if bc.pB != PERIODIC and bc.pT != PERIODIC:
        if dom.K == DOM.Ks:
            fw[dom.Gfz.get("_ks")] = 0
        if dom.K == DOM.Ke:
            fw[dom.Gfz.get("_ke")] = 0
    return fu, fv, fw


def cages_torch(dom, DOM, bc, parts, device):
    """phase, phase_shell (int32, Gcc s3b) and flag_u, flag_v, flag_w (int32, Gfx/Gfy/Gfz s3b) of one block for a list of
    spheres parts = (x, y, z, r) in global coordinates: cuda_build_cages (src/cuda_particle.cu:1516-1646) with cage_setup
    (src/particle_kernel.cu:135-146), build_phase (:148-253), build_phase_shell (:255-426), cage_flag_u/v/w (:482-540) and
    the external-wall flags (:542-576), evaluated with torch on `device` (bench.py builds the 1000-sphere case of
    BASELINE configs[3] on the GPU with it; tests/test_domain.py holds it to the CPU oracle bit for bit).
    The same IEEE operations in the same order as the reference's expressions, so floor(d / r) < 1 agrees on every cell."""
    import math
    import torch
    px, py, pz, pr = [np.asarray(v, dtype=np.float64) for v in parts]
    g = dom.Gcc
    knb, jnb, inb = g.get("knb"), g.get("jnb"), g.get("inb")
    phase = torch.full((knb, jnb, inb), -1, dtype=torch.int32, device=device)
    shell = torch.ones((knb, jnb, inb), dtype=torch.int32, device=device)
    S = [g.get("_is") if (dom.I == DOM.Is and bc.pW != PERIODIC) else g.get("_isb"),
         g.get("_js") if (dom.J == DOM.Js and bc.pS != PERIODIC) else g.get("_jsb"),
         g.get("_ks") if (dom.K == DOM.Ks and bc.pB != PERIODIC) else g.get("_ksb")]
    E = [g.get("_ie") if (dom.I == DOM.Ie and bc.pE != PERIODIC) else g.get("_ieb"),
         g.get("_je") if (dom.J == DOM.Je and bc.pN != PERIODIC) else g.get("_jeb"),
         g.get("_ke") if (dom.K == DOM.Ke and bc.pT != PERIODIC) else g.get("_keb")]
    dd = (dom.dx, dom.dy, dom.dz)
    ss = (dom.xs, dom.ys, dom.zs)
    nn = (dom.xn, dom.yn, dom.zn)

    def c_round(v):                                   # C round(): half away from zero
        return math.floor(v + 0.5) if v >= 0 else -math.floor(-v + 0.5)

    def cage(n):
        lo, hi = [0, 0, 0], [0, 0, 0]
        for a, pc in enumerate((px[n], py[n], pz[n])):
            cg = int(2. * math.ceil(pr[n] / dd[a])) + 2 - (nn[a] % 2)
            lo[a] = int(c_round((pc - ss[a]) * (1. / dd[a])) - 0.5 * cg + DOM_BUF_)
            hi[a] = lo[a] + cg
            lo[a] = min(max(lo[a], S[a]), E[a])
            hi[a] = min(max(hi[a], S[a]), E[a])
            if lo[a] == hi[a]:
                return None
        return lo, hi

    def axis(lo, hi, a, pc, off=0):
        l = torch.arange(lo, hi + 1, dtype=torch.float64, device=device)
        return (l + off - 0.5) * dd[a] - (pc - ss[a])

    cages = [cage(n) for n in range(len(pr))]
    for n, cg in enumerate(cages):                    # build_phase
        if cg is None:
            continue
        lo, hi = cg
        xx, yy, zz = axis(lo[0], hi[0], 0, px[n]), axis(lo[1], hi[1], 1, py[n]), axis(lo[2], hi[2], 2, pz[n])
        d2 = (xx * xx)[None, None, :] + (yy * yy)[None, :, None] + (zz * zz)[:, None, None]
        inside = torch.floor(torch.sqrt(d2) * (1. / pr[n])) < 1
        sub = phase[lo[2]:hi[2] + 1, lo[1]:hi[1] + 1, lo[0]:hi[0] + 1]
        sub[inside] = n
    for n, cg in enumerate(cages):                    # build_phase_shell
        if cg is None:
            continue
        lo, hi = cg
        ax = [[axis(lo[a], hi[a], a, pc, off) for off in (-1, 0, 1)] for a, pc in enumerate((px[n], py[n], pz[n]))]
        sq = [[v * v for v in row] for row in ax]

        def outside(ox, oy, oz):
            d2 = sq[0][ox + 1][None, None, :] + sq[1][oy + 1][None, :, None] + sq[2][oz + 1][:, None, None]
            return ~(torch.floor(torch.sqrt(d2) * (1. / pr[n])) < 1)
        any_out = outside(-1, 0, 0) | outside(1, 0, 0) | outside(0, -1, 0) | outside(0, 1, 0) | outside(0, 0, -1) | outside(0, 0, 1)
        sub_p = phase[lo[2]:hi[2] + 1, lo[1]:hi[1] + 1, lo[0]:hi[0] + 1]
        sub_s = shell[lo[2]:hi[2] + 1, lo[1]:hi[1] + 1, lo[0]:hi[0] + 1]
        sub_s[(sub_p == n) & any_out] = 0

    def face_flag(lo_p, hi_p, lo_s, hi_s):
        cut = ((lo_p < 0) & (hi_p > -1)) | ((lo_p > -1) & (hi_p < 0)) | ((hi_s < 1) & (lo_s < 1))
        return (1 - 2 * cut.to(torch.int32))
    fu = torch.ones(grid_shape(dom, "Gfx"), dtype=torch.int32, device=device)   # [i, k, j]
    fv = torch.ones(grid_shape(dom, "Gfy"), dtype=torch.int32, device=device)   # [j, i, k]
    fw = torch.ones(grid_shape(dom, "Gfz"), dtype=torch.int32, device=device)   # [k, j, i]
    # faces _is.._ie = 1..n+1 between cells i-1 and i; phase is [k, j, i]
    fu[1:-1] = face_flag(phase[:, :, :-1], phase[:, :, 1:], shell[:, :, :-1], shell[:, :, 1:]).permute(2, 0, 1)
    fv[1:-1] = face_flag(phase[:, :-1, :], phase[:, 1:, :], shell[:, :-1, :], shell[:, 1:, :]).permute(1, 2, 0)
    fw[1:-1] = face_flag(phase[:-1], phase[1:], shell[:-1], shell[1:])
    if bc.pW != PERIODIC and bc.pE != PERIODIC:
        if dom.I == DOM.Is:
            fu[dom.Gfx.get("_is")] = 0
        if dom.I == DOM.Ie:
            fu[dom.Gfx.get("_ie")] = 0
    if bc.pS != PERIODIC and bc.pN != PERIODIC:
        if dom.J == DOM.Js:
            fv[dom.Gfy.get("_js")] = 0
        if dom.J == DOM.Je:
            fv[dom.Gfy.get("_je")] = 0
    if bc.pB != PERIODIC and bc.pT != PERIODIC:
        if dom.K == DOM.Ks:
            fw[dom.Gfz.get("_ks")] = 0
        if dom.K == DOM.Ke:
            fw[dom.Gfz.get("_ke")] = 0
    return phase, shell, fu, fv, fw


DOM_BUF_ = 1                                          # src/bluebottle.h:141
