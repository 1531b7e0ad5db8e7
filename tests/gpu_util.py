"""Helpers that drive the CUDA library (through bbpcg's ctypes C-ABI binding) on a Case."""
import threading

import numpy as np

import bbpcg
from bbpcg.grid import BC_SETS


def decomposition_of(case):
    return bbpcg.Decomposition.uniform(case.extent, case.cells, case.blocks, BC_SETS[case.bcname])


class Product:
    """All ranks of a decomposition driven from ONE process on ONE GPU (each rank on its own
    stream and host thread); peers are plain device pointers.  The kernels, the in-kernel
    all-reduce and the peer halo stores are exactly those of a multi-GPU run."""

    def __init__(self, case, options=None):
        import torch
        self.torch = torch
        self.case = case
        self.dec = decomposition_of(case)
        self.n = self.dec.nranks
        self.solvers = [bbpcg.PoissonSolver(self.dec, r) for r in range(self.n)]
        if self.n > 1:
            blobs = [s.comm_export() for s in self.solvers]
            for s in self.solvers:
                s.comm_import(blobs)
        for s in self.solvers:
            s.set_option("comm_timeout_ms", 30000)      # a rank thread that page-faults on a fresh box can be seconds late (seen once at 4000)
            for k, v in (options or {}).items():
                s.set_option(k, v)
        self.dev = []
        for r, s in enumerate(self.solvers):
            inp = case.inputs(r)
            d = {k: s.to_device(v) for k, v in inp.items()}
            d["rhs_p"] = s.empty("Gcc")
            d["phi"] = s.empty("Gcc")
            self.dev.append(d)

    def each(self, fn):
        """Run fn(rank, solver, arrays) on every rank concurrently (collective calls)."""
        out, err = [None] * self.n, [None] * self.n

        def work(r):
            try:
                out[r] = fn(r, self.solvers[r], self.dev[r])
            except Exception as e:  # noqa: BLE001
                err[r] = e
        if self.n == 1:
            work(0)
        else:
            th = [threading.Thread(target=work, args=(r,)) for r in range(self.n)]
            [t.start() for t in th]
            [t.join() for t in th]
        for e in err:
            if e is not None:
                raise e
        return out

    def set_coefficients(self, parts=False):
        self.each(lambda r, s, d: s.init_jacobi_preconditioner(d["flag_u"], d["flag_v"], d["flag_w"],
                                                               d["phase"] if parts else None))

    def solve(self, parts=False, **kw):
        if parts:
            return self.each(lambda r, s, d: s.PP_cg(d["u_star"], d["v_star"], d["w_star"], d["rhs_p"], d["phi"],
                                                     d["phase"], d["phase_shell"], **kw))
        return self.each(lambda r, s, d: s.PP_cg_noparts(d["u_star"], d["v_star"], d["w_star"], d["rhs_p"], d["phi"], **kw))

    def gather(self, key):
        """global interior field (Nz, Ny, Nx) from the ranks' ghosted Gcc arrays"""
        D = self.dec.DOM
        out = np.zeros((D.zn, D.yn, D.xn))
        for r in range(self.n):
            d = self.dec.doms[r]
            a = self.dev[r][key].cpu().numpy()[1:-1, 1:-1, 1:-1]
            i0, j0, k0 = d.Gcc.get("is") - 1, d.Gcc.get("js") - 1, d.Gcc.get("ks") - 1
            out[k0:k0 + d.zn, j0:j0 + d.yn, i0:i0 + d.xn] = a
        return out

    def close(self):
        for s in self.solvers:
            s.close()


class TorchProduct:
    """Like Product, but every input is built on the GPU with bbpcg.synth's torch builders (bit-identical to the numpy
    ones), so grids of the BENCHMARKED sizes (256^3, 512x256x256, ...) need no host-side oracle state.  All ranks of the
    decomposition share one process and one GPU."""

    def __init__(self, extent, cells, blocks, bcname, nparts=0, radius=1.0, options=None):
        import torch
        from bbpcg import synth
        from bbpcg.grid import grid_shape
        self.torch = torch
        self.dec = bbpcg.Decomposition.uniform(extent, cells, blocks, BC_SETS[bcname])
        self.n, self.nparts = self.dec.nranks, nparts
        self.solvers = [bbpcg.PoissonSolver(self.dec, r) for r in range(self.n)]
        if self.n > 1:
            blobs = [s.comm_export() for s in self.solvers]
            for s in self.solvers:
                s.comm_import(blobs)
        self.dev = []
        parts = synth.random_spheres(self.dec.DOM, nparts, radius) if nparts else None
        for r, s in enumerate(self.solvers):
            s.set_option("comm_timeout_ms", 20000)
            for k, v in (options or {}).items():
                s.set_option(k, v)
            dom, dv = self.dec.doms[r], s.device
            d = {}
            if nparts:
                d["phase"], d["phase_shell"], d["flag_u"], d["flag_v"], d["flag_w"] = synth.cages_torch(dom, self.dec.DOM, self.dec.bc, parts, dv)
            else:
                d["flag_u"], d["flag_v"], d["flag_w"] = synth.flags_noparts_torch(dom, self.dec.DOM, self.dec.bc, dv)
                d["phase"] = torch.full(grid_shape(dom, "Gcc"), -1, dtype=torch.int32, device=dv)
                d["phase_shell"] = d["phase"]
            d["u_star"], d["v_star"], d["w_star"] = synth.velocity_star_torch(dom, self.dec.DOM, self.dec.bc, dv)
            d["rhs_p"], d["phi"] = s.empty("Gcc"), s.empty("Gcc")
            self.dev.append(d)

    each = Product.each

    def set_coefficients(self):
        self.each(lambda r, s, d: s.init_jacobi_preconditioner(d["flag_u"], d["flag_v"], d["flag_w"], d["phase"] if self.nparts else None))

    def solve(self, **kw):
        if self.nparts:
            return self.each(lambda r, s, d: s.PP_cg(d["u_star"], d["v_star"], d["w_star"], d["rhs_p"], d["phi"], d["phase"], d["phase_shell"], **kw))
        return self.each(lambda r, s, d: s.PP_cg_noparts(d["u_star"], d["v_star"], d["w_star"], d["rhs_p"], d["phi"], **kw))

    def gather(self, key):
        """global interior field (Nz, Ny, Nx) as a torch tensor on the GPU"""
        D = self.dec.DOM
        out = self.torch.zeros((D.zn, D.yn, D.xn), dtype=self.torch.float64, device=self.solvers[0].device)
        for r in range(self.n):
            d = self.dec.doms[r]
            i0, j0, k0 = d.Gcc.get("is") - 1, d.Gcc.get("js") - 1, d.Gcc.get("ks") - 1
            out[k0:k0 + d.zn, j0:j0 + d.yn, i0:i0 + d.xn] = self.dev[r][key][1:-1, 1:-1, 1:-1]
        return out

    def close(self):
        for s in self.solvers:
            s.close()
