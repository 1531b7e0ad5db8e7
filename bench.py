#!/usr/bin/env python
"""bench.py -- pressure-Poisson PCG throughput on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one complete pressure-Poisson solve through the reference-facing entry point
(cuda_PP_cg_noparts -> PoissonSolver.PP_cg_noparts -> bbpcg_solve; with --parts: cuda_PP_cg): PP_rhs, set-up, the PCG
iteration loop to pp_residual = 1e-6 and the phi write-back, on the synthetic 512^3 FP64 grid
(duct boundary set).  `value` = PCG iterations of all timed steps / their time,
i.e. whole-step iterations/s (set-up included).  N > 1: the same GLOBAL grid decomposed over N
ranks with Bluebottle's decomp.config block rule (strong scaling), one process per GPU, halos and
dot products through NVLink peer memory inside the kernels.

Rank 0 prints ONE JSON line.  Extra objects: `roofline` (dominant kernel k_search_tma, CUDA
events on the solver's stream, live in the timed region), `roofline_iteration` (72-B model over
the whole iteration loop), `cpu_baseline` (the OpenMP C port of the reference recurrence,
oracle/pcg_ref.c, on the host cores; bounded sample), `e2e` (host buffers -> bbpcg_solve_host ->
host buffer), `parity` (phi of this arm against the phi of the reference arm on the same inputs), `clocks`.

--impl reference: the reference's OWN unmodified CUDA kernels + host loop (oracle/_ref/libbbref.so,
compiled from /root/reference/src by oracle/Makefile) on one GPU, same workload, same metric.
The reference has no CPU implementation of this path (BASELINE.json north_star); if that library
cannot be loaded the arm falls back to the OpenMP C port on the host cores and says so.  That arm never maps the
product library: its dom_struct comes from the oracle's own domain_fill restatement.  It leaves its phi in
/tmp (REF_PHI_FMT) for the `parity` object of the bbpcg arm the driver runs right after it; when the file is absent
the bbpcg arm runs oracle/_ref once itself, AFTER its timed region, as the checker.

Other workloads (BASELINE.json configs): --grid 256 (configs[1]); --cells 512,256,256 --bc channel (configs[2]);
--parts 1000 --bc sedimentation --length 64 (configs[3]); --grid 1024 / --scaling weak (configs[4]).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "bluebottle-3.0_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

# algorithmic bytes per interior cell (docs/bytes_model.md).  The COMMITTED model of the iteration is 72 B
#  and the iteration roofline is always scored against it.  The library MOVES 64: the search kernel
# does not store q (40 B: r, p_prev, x read; p_new, x written) and the residual kernel re-applies the operator to p
# instead of reading q (24 B: p, r read; r written).
BYTES_SEARCH = 40
BYTES_RESID = 24
BYTES_ITER = 72
BYTES_MOVED = 64
# solve epilogue: project 8 (phi) + 24 (u*,v*,w*) + 12 (int flags) + 24 (u,v,w) = 68; update_p 8 (p0) + 4 (phase) + 8 (p) = 20;
# mean subtraction 16 (p read + write): 104 B per cell
BYTES_EPILOGUE = 104
BLOCKS_FOR = {1: (1, 1, 1), 2: (1, 1, 2), 4: (1, 2, 2), 8: (1, 2, 4)}   # x never split: 128-cell x-tiles stay full, halo pushes are contiguous rows (profiles/r02z_*: 4260 vs 4103 it/s for 2x2x2)
REF_PHI_FMT = "/tmp/bbpcg_ref_phi_%s.npy"
PP_RESIDUAL, PP_MAX_ITER, RHO_F, DT = 1e-6, 2000, 1.0, 1e-3
KT_SAMPLE = 64          # iterations of the first timed step whose kernels are bracketed by CUDA events (roofline.avg_launch_us)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="bbpcg", choices=["bbpcg", "reference"])
    ap.add_argument("--grid", type=int, default=512, help="global cells per side")
    ap.add_argument("--cells", default="", help="Nx,Ny,Nz override of --grid (BASELINE configs[2]: 512,256,256)")
    ap.add_argument("--length", type=float, default=12.0, help="domain length in x (dx = length / Nx in every direction)")
    ap.add_argument("--bc", default="duct")
    ap.add_argument("--parts", type=int, default=0, help="> 0: this many non-overlapping spheres of radius 1 through cuda_PP_cg "
                    "(BASELINE configs[3]: --parts 1000 --bc sedimentation --length 64)")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"])
    ap.add_argument("--blocks", default="", help="In,Jn,Kn override")
    ap.add_argument("--ty", type=int, default=-1, help="tile rows of the iteration kernels (default: planner)")
    ap.add_argument("--kc", type=int, default=-1, help="planes per z-chunk (default: planner)")
    ap.add_argument("--fixed-iters", type=int, default=0, help=">0: fixed iteration count per step, no stop test")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-epilogue", action="store_true", help="skip the solve-epilogue measurement (project + update_p)")
    ap.add_argument("--no-comm-split", action="store_true", help="N > 1: skip the exposed halo/all-reduce measurement")
    ap.add_argument("--no-parity", action="store_true", help="skip the phi comparison with the reference arm")
    ap.add_argument("--cpu-sample-grid", type=int, default=256)
    ap.add_argument("--cpu-sample-iters", type=int, default=400, help="fixed iterations of the CPU sample (~12 s of CPU work at 256^3)")
    ap.add_argument("--ref-kind", default="auto", choices=["auto", "cuda", "port"])
    ap.add_argument("--opt", action="append", default=[], help="solver tuning option key=value (bbpcg_set_option)")
    return ap.parse_args()


# ---- the workload, identical in both arms --------------------------------------------------------
def workload(args, n):
    blocks = tuple(int(v) for v in args.blocks.split(",")) if args.blocks else BLOCKS_FOR[n]
    base = tuple(int(v) for v in args.cells.split(",")) if args.cells else (args.grid,) * 3
    cells = tuple(b * k for b, k in zip(base, blocks)) if args.scaling == "weak" else base
    dx = args.length / base[0]
    extent = (0., dx * cells[0], 0., dx * cells[1], 0., dx * cells[2])
    return cells, blocks, extent


def config_of(args, cells):
    """the `config` object: the same keys and strings in the bbpcg arm and the reference arm"""
    entry = "cuda_PP_cg" if args.parts else "cuda_PP_cg_noparts"
    return {"workload": "synthetic FP64 pressure-Poisson, %dx%dx%d cells, %s boundary set%s, PP_rhs + Jacobi-PCG to pp_residual %g (%s)"
                        % (cells + (args.bc, ", %d spheres of radius 1" % args.parts if args.parts else "", PP_RESIDUAL, entry)),
            "cells": "%dx%dx%d" % cells, "bc": args.bc, "parts": args.parts, "length_x": args.length, "scaling": args.scaling,
            "pp_residual": PP_RESIDUAL, "pp_max_iter": PP_MAX_ITER, "rho_f": RHO_F, "dt": DT, "fixed_iters": args.fixed_iters,
            "rhs": "u* = sin cos cos + uniform noise (splitmix64 of the global face index), wall-normal faces zero",
            "l2": "inputs larger than L2 (4 FP64 solver vectors of %.2f GB vs 126 MB)" % (8 * cells[0] * cells[1] * cells[2] / 1e9)}


def metric_of(args, cells):
    return "Poisson PCG iterations/s (FP64, %s)" % ("%d^3" % cells[0] if cells[0] == cells[1] == cells[2] else "%dx%dx%d" % cells) + \
           (" per-GPU %d^3" % args.grid if args.scaling == "weak" else "")


def ref_phi_path(args, cells):
    return REF_PHI_FMT % ("%dx%dx%d_%s_p%d_L%g_f%d" % (cells + (args.bc, args.parts, args.length, args.fixed_iters)))


def build_inputs(args, dom, DOM, bc, dev):
    """flags, phase, phase_shell and u*, v*, w* of one block on `dev` (deterministic, keyed on global indices)"""
    import torch
    from bbpcg import synth
    from bbpcg.grid import grid_shape
    if args.parts:
        parts = synth.random_spheres(DOM, args.parts, 1.0)
        phase, shell, fu, fv, fw = synth.cages_torch(dom, DOM, bc, parts, dev)
    else:
        fu, fv, fw = synth.flags_noparts_torch(dom, DOM, bc, dev)
        phase = torch.full(grid_shape(dom, "Gcc"), -1, dtype=torch.int32, device=dev)
        shell = phase
    u, v, w = synth.velocity_star_torch(dom, DOM, bc, dev)
    return {"fu": fu, "fv": fv, "fw": fw, "phase": phase, "shell": shell, "u": u, "v": v, "w": w}


# ---- clocks -----------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi sampled every 200 ms DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index, period_ms=200):
        self.index, self.proc, self.lines, self.period_ms = index, None, [], period_ms

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", str(self.period_ms)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md, MEASURED_PEAKS.json absent)"


def traffic_from_profile():
    """per-launch DRAM bytes of the iteration kernels from the committed ncu --set full capture, if any"""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f)
    except Exception:
        return None


def median(v):
    s = sorted(v)
    return s[len(s) // 2] if len(s) % 2 else 0.5 * (s[len(s) // 2 - 1] + s[len(s) // 2])


# ---- CPU baseline (oracle port, OpenMP) --------------------------------------------------------
def cpu_port_sample(cells, sample_grid, iters, bc):
    """The OpenMP C port of the reference recurrence on a bounded sample: a sample_grid^3 block of
    the same discretisation and boundary set, `iters` fixed iterations; it/s scaled by the cell ratio to the bench grid."""
    from cases import Case
    from oracle import binding as ob
    case = Case((sample_grid,) * 3, bc=bc, omp=True)
    cores = ob.load(True).bbo_omp_threads()
    case.o.iterate_fixed(2)                       # touch pages
    t0 = time.perf_counter()
    case.o.iterate_fixed(iters)
    dt = time.perf_counter() - t0
    its_sample = iters / dt
    ncell = cells[0] * cells[1] * cells[2]
    scale = sample_grid ** 3 / float(ncell)
    return {"value": its_sample * scale, "unit": "PCG iterations/s", "cores": int(cores), "kind": "port",
            "sample": "%d^3 block (%.4g of the %dx%dx%d cells; no particles), %d fixed iterations in %.1f s incl. set-up; "
                      "%.2f it/s on the sample, scaled by the cell ratio (the recurrence is O(cells) per iteration)"
                      % ((sample_grid, scale) + cells + (iters, dt, its_sample)),
            "gbs_72B_model": BYTES_ITER * sample_grid ** 3 * its_sample / 1e9}


# ---- distributed plumbing ----------------------------------------------------------------------
def bind_near_gpu(local):
    """One process per GPU: run on the host cores NVML lists for this GPU, so that the pinned buffers of the end-to-end leg
    are allocated on the GPU's own NUMA node (eight ranks copying across the socket link measured 17 GB/s per GPU, r02q).
    Returns the number of cores bound to, or None when NVML cannot tell."""
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        pr = torch.cuda.get_device_properties(local)
        try:
            h = pynvml.nvmlDeviceGetHandleByPciBusId(("%08x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)).encode())
        except Exception:  # noqa: BLE001
            h = pynvml.nvmlDeviceGetHandleByIndex(local)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * i + b for i, w_ in enumerate(words) for b in range(64) if (int(w_) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:  # noqa: BLE001
        return None


class World:
    def __init__(self, want):
        self.rank = int(os.environ.get("RANK", "0"))
        self.size = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        self.dist = None
        self.bound_cpus = None
        if self.size > 1:
            import torch
            import torch.distributed as dist
            torch.cuda.set_device(self.local)
            self.bound_cpus = bind_near_gpu(self.local)
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
            self.dist = dist
        if want != self.size:
            if self.rank == 0:
                sys.stderr.write("bench.py: --gpus %d but WORLD_SIZE=%d (launch with torch.distributed.run)\n" % (want, self.size))
            sys.exit(2)

    def barrier(self):
        import torch
        if self.dist:
            self.dist.barrier()
        torch.cuda.synchronize()

    def _red(self, v, op):
        import torch
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        self.dist.all_reduce(t, op=op)
        return float(t.item())

    def max(self, v):
        return self._red(v, self.dist.ReduceOp.MAX) if self.dist else v

    def sum(self, v):
        return self._red(v, self.dist.ReduceOp.SUM) if self.dist else v

    def close(self):
        if self.dist:
            self.dist.barrier()
            self.dist.destroy_process_group()


# ---- the reference's own CUDA path (oracle/_ref), used by the reference arm and as the parity checker -----------
class RefSolver:
    """oracle/_ref/libbbref.so on ONE block: the reference's unmodified kernels + its host loop (cuda_PP_cg[_noparts])."""

    def __init__(self, args, cells, extent, dev):
        import torch
        from cases import load_ref
        from bbpcg.grid import BC_SETS
        from oracle import binding as ob
        self.torch = torch
        self.lib = load_ref()
        if self.lib is None:
            raise RuntimeError("oracle/_ref/libbbref.so is not built")
        self.DOM, self.dom, self.bc = ob.single_block_domain(extent, cells, BC_SETS[args.bc])
        self.parts = 1 if args.parts else 0
        assert self.lib.bbref_init(C.byref(self.dom), C.byref(self.DOM)) == 0
        self.inp = build_inputs(args, self.dom, self.DOM, self.bc, dev)
        torch.cuda.synchronize()
        P = self.P
        i = self.inp
        assert self.lib.bbref_set_inputs_dev(P(i["fu"]), P(i["fv"]), P(i["fw"]), P(i["phase"]), P(i["shell"]), P(i["u"]), P(i["v"]),
                                             P(i["w"]), args.parts) == 0
        self.lib.bbref_solve_host.argtypes = [C.c_void_p] * 4 + [C.c_double] * 3 + [C.c_int, C.c_int, C.POINTER(C.c_int),
                                                                                    C.POINTER(C.c_double), C.POINTER(C.c_float)]
        self.niter, self.resid, self.ms = C.c_int(), C.c_double(), C.c_float()

    @staticmethod
    def P(t):
        return C.c_void_p(t.data_ptr())

    def solve_dev(self):
        assert self.lib.bbref_solve(RHO_F, DT, PP_RESIDUAL, PP_MAX_ITER, self.parts, C.byref(self.niter), C.byref(self.resid),
                                    C.byref(self.ms)) == 0
        return self.niter.value, self.ms.value

    def solve_host(self, hu, hv, hw, hphi):
        P = self.P
        assert self.lib.bbref_solve_host(P(hu), P(hv), P(hw), P(hphi), RHO_F, DT, PP_RESIDUAL, PP_MAX_ITER, self.parts,
                                         C.byref(self.niter), C.byref(self.resid), C.byref(self.ms)) == 0
        return self.niter.value, self.ms.value

    def phi_interior(self):
        import numpy as np
        from bbpcg.grid import grid_shape
        a = np.zeros(grid_shape(self.dom, "Gcc"))
        assert self.lib.bbref_get(0, a.ctypes.data_as(C.c_void_p)) == 0
        return a[1:-1, 1:-1, 1:-1]

    def save_phi(self, path):
        import numpy as np
        tmp = path + ".tmp.npy"
        np.save(tmp, np.concatenate([np.array([float(self.niter.value), self.resid.value]), self.phi_interior().ravel()]))
        os.replace(tmp, path)


def parity_against_reference(args, w, cells, extent, dom, phi, niter, dev):
    """phi of this arm against the phi the reference's own kernels computed on the same inputs: relative L2 over the
    global interior (north_star: <= 1e-10) and the iteration counts (+-1).  The reference phi comes from the file the
    reference arm left (REF_PHI_FMT) or, when absent, from one run of oracle/_ref here -- after the timed region, as the
    checker only."""
    import numpy as np
    import torch
    path = ref_phi_path(args, cells)
    source = "file written by `bench.py --impl reference` (%s)" % path
    if args.fixed_iters:
        return {"skipped": "fixed-iteration mode: the reference has no such mode"}
    if not os.path.exists(path):
        source = "oracle/_ref run in this process after the timed region, as the checker (%s absent)" % path
        if w.rank == 0:
            try:
                ref = RefSolver(args, cells, extent, dev)
                ref.solve_dev()
                ref.save_phi(path)
                del ref
                torch.cuda.empty_cache()
            except Exception as e:  # noqa: BLE001
                sys.stderr.write("bench.py: parity checker unavailable: %s\n" % e)
        w.barrier()
    if not os.path.exists(path):
        return {"skipped": "no reference phi: oracle/_ref/libbbref.so not loadable"}
    raw = np.load(path, mmap_mode="r")
    niter_ref = int(raw[0])
    nz, ny, nx = cells[2], cells[1], cells[0]
    ref = raw[2:].reshape(nz, ny, nx)
    g = dom.Gcc
    i0, j0, k0 = g.get("is") - 1, g.get("js") - 1, g.get("ks") - 1
    blk = torch.from_numpy(np.ascontiguousarray(ref[k0:k0 + dom.zn, j0:j0 + dom.yn, i0:i0 + dom.xn])).to(dev)
    mine = phi[1:-1, 1:-1, 1:-1]
    num = w.sum(float(((mine - blk) ** 2).sum()))
    den = w.sum(float((blk ** 2).sum()))
    amax = w.max(float((mine - blk).abs().max()))
    rel = (num / den) ** 0.5 if den > 0 else None
    return {"rel_l2": rel, "max_abs_diff": amax, "niter": niter, "niter_ref": niter_ref,
            "tolerance": "rel_l2 <= 1e-10, |niter - niter_ref| <= 1 (BASELINE.json north_star)",
            "ok": bool(rel is not None and rel <= 1e-10 and abs(niter - niter_ref) <= 1), "reference": source}


# ---- our arm -----------------------------------------------------------------------------------
def run_bbpcg(args):
    import torch
    import bbpcg
    from bbpcg.grid import BC_SETS
    w = World(args.gpus)
    torch.cuda.set_device(w.local)
    dev = torch.device("cuda", w.local)
    cells, blocks, extent = workload(args, w.size)
    dec = bbpcg.Decomposition.uniform(extent, cells, blocks, BC_SETS[args.bc])
    assert dec.nranks == w.size, "decomposition %s needs %d ranks" % (blocks, dec.nranks)
    s = bbpcg.PoissonSolver(dec, w.rank, device=w.local)
    if w.size > 1:
        s.comm_init_torch()
    for key, val in (("ty", args.ty), ("kc", args.kc)):
        if val >= 0:
            s.set_option(key, val)
    for kv in args.opt:
        key, val = kv.split("=")
        s.set_option(key, int(val))
    dom = dec.doms[w.rank]
    inp = build_inputs(args, dom, dec.DOM, dec.bc, dev)
    u, v, wz = inp["u"], inp["v"], inp["w"]
    rhs, phi = s.empty("Gcc"), s.empty("Gcc")
    s.init_jacobi_preconditioner(inp["fu"], inp["fv"], inp["fw"], inp["phase"] if args.parts else None)
    ncell_rank = dom.xn * dom.yn * dom.zn
    ncell_glob = cells[0] * cells[1] * cells[2]
    kw = dict(rho_f=RHO_F, dt=DT, pp_residual=PP_RESIDUAL, pp_max_iter=PP_MAX_ITER, fixed_iters=args.fixed_iters)

    def solve(**over):
        k = dict(kw, **over)
        if args.parts:
            return s.PP_cg(u, v, wz, rhs, phi, inp["phase"], inp["shell"], **k)
        return s.PP_cg_noparts(u, v, wz, rhs, phi, **k)

    for _ in range(args.warmup):
        r = solve()
    clocks = ClockSampler(w.local) if w.rank == 0 else None
    w.barrier()
    if clocks:
        clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    iters = launches = 0
    ms_iter = ms_setup = 0.0
    step_ms = []
    kt = {k: 0 for k in ("kt_search_ns", "kt_resid_ns", "kt_refresh_ns", "kt_search_n", "kt_resid_n", "kt_refresh_n")}
    kt_ms_iter = 1e-9
    for step in range(args.steps):
        # per-kernel CUDA events (solver stream) around the launches of the first KT_SAMPLE iterations of the FIRST timed step
        # only: an event between two kernels forbids the programmatic dependent launch everything else runs with, and
        # instrumenting all 313 iterations made that step 8 % slower than its neighbours at N = 8 (r02aj: 72.1 mean / 71.0 median)
        timed_kernels = step == 0
        if timed_kernels:
            s.set_option("kernel_timing", KT_SAMPLE)
        r = solve()                                           # host-synchronous collective call
        assert r.status == "converged", r
        iters += r.niter; launches += r.launches; ms_iter += r.ms_iter; ms_setup += r.ms_total - r.ms_iter
        step_ms.append(r.ms_total)
        if timed_kernels:
            for k in kt:
                kt[k] += s.info(k)
            kt_ms_iter = r.ms_iter
            s.set_option("kernel_timing", 0)
    e1.record()
    w.barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    ms = max(e0.elapsed_time(e1), 0.0)
    clk = clocks.stop() if clocks else None
    ms = w.max(ms)
    ms_iter_max = w.max(ms_iter)
    launches_all = w.sum(launches)
    niter_last = r.niter
    value = iters / (ms * 1e-3)
    peak, peak_src = measured_peak()
    search_s = w.max(kt["kt_search_ns"] * 1e-9 / max(kt["kt_search_n"], 1))
    resid_s = w.max(kt["kt_resid_ns"] * 1e-9 / max(kt["kt_resid_n"], 1))
    ach = BYTES_SEARCH * ncell_rank / search_s / 1e9
    tr = traffic_from_profile() or {}
    tr_cells = tr.get("cells_per_launch", 512 ** 3)
    same_launch = (w.size == 1 and ncell_rank == tr_cells and not args.parts)

    def _traffic(key):
        val = tr.get(key)
        return None if val is None else val * (ncell_rank / float(tr_cells))
    roof = {"kernel": "k_search_tma<%s, 2> (q not stored: re-applied by k_resid_tma)" % ("true" if args.parts else "false"), "bound": "hbm",
            "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
            "traffic": _traffic("k_search_tma_bytes_per_launch"),
            "traffic_kind": "measured (ncu --set full of this launch shape)" if same_launch else
                            "extrapolated: the 512^3 single-GPU capture scaled by the cell count, not a measurement of this launch",
            "traffic_source": tr.get("source"), "peak_source": peak_src,
            "algorithmic_bytes_per_cell": BYTES_SEARCH, "cells_per_launch": ncell_rank,
            "avg_launch_us": search_s * 1e6, "launches_timed": kt["kt_search_n"],
            "timed": "CUDA events on the solver stream around every launch of the first %d iterations of the first timed step" % KT_SAMPLE,
            "share_of_iteration_loop": kt["kt_search_ns"] / max(kt["kt_search_ns"] + kt["kt_resid_ns"] + kt["kt_refresh_ns"], 1),
            "plan": {"ty": s.info("search_ty"), "planes_per_chunk": s.info("search_kc"), "ctas": s.info("search_grid"), "pdl": s.info("pdl")}}
    ach2 = BYTES_RESID * ncell_rank / resid_s / 1e9
    roof2 = {"kernel": "k_resid_tma (re-applies the operator to p)", "bound": "hbm", "achieved": ach2, "peak": peak, "unit": "GB/s",
             "frac": ach2 / peak, "traffic": _traffic("k_resid_tma_bytes_per_launch"), "algorithmic_bytes_per_cell": BYTES_RESID,
             "avg_launch_us": resid_s * 1e6, "launches_timed": kt["kt_resid_n"]}
    ach_it = BYTES_ITER * ncell_rank * iters / (ms_iter_max * 1e-3) / 1e9
    roof_it = {"bound": "hbm", "achieved": ach_it, "peak": peak, "unit": "GB/s", "frac": ach_it / peak,
               "model": "72 B/cell/iteration (committed model) over the iteration loop only, per GPU",
               "bytes_moved_per_cell": BYTES_MOVED,
               "achieved_moved": BYTES_MOVED * ncell_rank * iters / (ms_iter_max * 1e-3) / 1e9,
               "frac_moved": BYTES_MOVED * ncell_rank * iters / (ms_iter_max * 1e-3) / 1e9 / peak,
               "iter_loop_its": iters / (ms_iter_max * 1e-3), "us_per_iteration": ms_iter_max * 1e3 / max(iters, 1)}

    # ---- parity with the reference arm (same inputs), after the timed region --------------------------------------
    parity = None
    if not args.no_parity:
        parity = parity_against_reference(args, w, cells, extent, dom, phi, niter_last, dev)

    # ---- exposed communication (N > 1): the same per-rank block solved stand-alone (no peers: no halo push
    # over NVLink, no cross-rank all-reduce wait), same kernels, fixed iteration count ----------------------
    comm = None
    if w.size > 1 and not args.no_comm_split:
        ext1 = (dom.xs, dom.xe, dom.ys, dom.ye, dom.zs, dom.ze)
        dec1 = bbpcg.Decomposition.uniform(ext1, (dom.xn, dom.yn, dom.zn), (1, 1, 1), BC_SETS[args.bc])
        s1 = bbpcg.PoissonSolver(dec1, 0, device=w.local)
        for kv in args.opt:
            key, val = kv.split("=")
            s1.set_option(key, int(val))
        a1 = argparse.Namespace(**dict(vars(args), parts=0))
        i1 = build_inputs(a1, dec1.doms[0], dec1.DOM, dec1.bc, dev)
        s1.init_jacobi_preconditioner(i1["fu"], i1["fv"], i1["fw"])
        rhs1, phi1 = s1.empty("Gcc"), s1.empty("Gcc")
        s1.PP_cg_noparts(i1["u"], i1["v"], i1["w"], rhs1, phi1, fixed_iters=40)
        w.barrier()
        r1 = s1.PP_cg_noparts(i1["u"], i1["v"], i1["w"], rhs1, phi1, fixed_iters=200)
        local_us = w.max(r1.ms_iter * 1e3 / 200)
        w.barrier()
        rN = solve(fixed_iters=200)
        coll_us = w.max(rN.ms_iter * 1e3 / 200)
        comm = {"us_per_iteration": coll_us, "us_per_iteration_standalone_block": local_us,
                "exposed_halo_plus_allreduce_us": coll_us - local_us,
                "method": "200 fixed iterations of the decomposed solve vs the same per-rank block solved as a 1x1x1 domain "
                          "(no peer reads, no cross-rank wait; particle-free); max over ranks; the difference is the exposed halo + all-reduce time"}
        s1.close()
        del i1, rhs1, phi1

    # ---- end to end: pinned host buffers -> bbpcg_solve_host -> pinned host phi -------------------
    e2e = None
    if not args.no_e2e and not args.parts:
        hu, hv, hw = [t.cpu().pin_memory() for t in (u, v, wz)]
        hphi = torch.zeros(tuple(phi.shape), dtype=torch.float64).pin_memory()
        h2d = sum(t.numel() * 8 for t in (hu, hv, hw))
        d2h = hphi.numel() * 8
        s.solve_host(hu, hv, hw, hphi, **kw)                  # allocates the staging arrays
        k_e2e = max(1, min(args.steps, 3))
        w.barrier()
        it_e, t_steps = 0, []
        t0 = time.perf_counter()
        for _ in range(k_e2e):
            ts = time.perf_counter()
            re = s.solve_host(hu, hv, hw, hphi, **kw)
            t_steps.append(time.perf_counter() - ts)
            it_e += re.niter
        w.barrier()
        te = w.max(time.perf_counter() - t0)
        e2e = {"value": it_e / te, "unit": "PCG iterations/s", "h2d_bytes_per_step": int(w.sum(h2d)),
               "d2h_bytes_per_step": int(w.sum(d2h)), "steps": k_e2e, "ms_per_step": te * 1e3 / k_e2e,
               "ms_per_step_median": w.max(median(t_steps)) * 1e3,
               "api": "bbpcg_solve_host (C ABI): u*,v*,w* pinned host -> device, solve, phi -> pinned host",
               "host_cores_bound_per_rank": w.bound_cpus}
        del hu, hv, hw, hphi

    # ---- the solve epilogue: exchange(phi) + dom_BC_p + cuda_project + cuda_update_p, one fused
    # call per step; reported beside the headline, not part of `value` -------------------------------------
    epi = None
    if not args.no_epilogue:
        from bbpcg.grid import grid_shape
        un, vn, wn, pn = s.empty("Gfx"), s.empty("Gfy"), s.empty("Gfz"), s.empty("Gcc")
        p0 = torch.rand(grid_shape(dom, "Gcc"), dtype=torch.float64, device=dev)
        ea = (phi, u, v, wz, inp["fu"], inp["fv"], inp["fw"], un, vn, wn, p0, inp["phase"], pn)
        s.epilogue(*ea)
        w.barrier()
        n_epi = 5
        ms_epi = sum(s.epilogue(*ea) for _ in range(n_epi)) / n_epi
        ms_epi = w.max(ms_epi)
        gbs = BYTES_EPILOGUE * ncell_rank / (ms_epi * 1e-3) / 1e9
        epi = {"ms_per_call": ms_epi, "calls_timed": n_epi, "bound": "hbm",
               "algorithmic_bytes_per_cell": BYTES_EPILOGUE, "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak,
               "what": "bbpcg_epilogue (C ABI): mpi_cuda_exchange_Gcc(phi) + cuda_dom_BC_p(phi) + cuda_project + cuda_update_p "
                       "(src/bluebottle.c:233-250); CUDA events on the solver stream",
               "mean_p_after": w.sum(float(pn[1:-1, 1:-1, 1:-1].sum())) / ncell_glob}      # over ALL ranks' cells
        # solve prologue: cuda_solvability on u*, v*, w* (6 boundary planes, one 3-value all-reduce); wall clock around the
        # host-synchronous C-ABI call (2 launches); the correction it applies is undone by calling it on copies
        uc, vc, wc = u.clone(), v.clone(), wz.clone()
        s.solvability(uc, vc, wc, "HOMOGENEOUS")
        w.barrier()
        t0 = time.perf_counter()
        n_sol = 20
        for _ in range(n_sol):
            eps = s.solvability(uc, vc, wc, "HOMOGENEOUS")
        torch.cuda.synchronize()
        epi["solvability_us_per_call_wall"] = w.max((time.perf_counter() - t0) / n_sol * 1e6)
        epi["solvability_eps_after"] = eps
        del un, vn, wn, pn, p0, uc, vc, wc

    out = {"metric": metric_of(args, cells), "value": value, "unit": "PCG iterations/s", "n_gpus": w.size, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": ms / args.steps, "ms_per_step_median": w.max(median(step_ms)),
           "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": config_of(args, cells),
           "decomposition": {"blocks": "%dx%dx%d" % blocks, "cells_per_gpu": ncell_rank, "iterations_per_step": iters / args.steps},
           "impl": "bbpcg", "gpu_launches": int(launches_all), "e2e": e2e, "roofline": roof, "roofline_resid": roof2,
           "roofline_iteration": roof_it, "parity": parity, "comm": comm, "clocks": clk, "wall_ms_per_step": wall_ms / args.steps,
           "setup_ms_per_step": ms_setup / args.steps, "epilogue": epi,
           "hbm_gbs_72B_model_whole_step": BYTES_ITER * ncell_glob * value / w.size / 1e9}
    if w.rank == 0 and w.size == 1 and not args.no_cpu_baseline:
        try:
            out["cpu_baseline"] = cpu_port_sample(cells, args.cpu_sample_grid, args.cpu_sample_iters, args.bc)
        except Exception as e:  # noqa: BLE001
            out["cpu_baseline"] = {"value": None, "unit": "PCG iterations/s", "cores": 0, "kind": "port", "sample": "failed: %s" % e}
    s.close()
    if w.rank == 0:
        print(json.dumps(out))
    w.close()


# ---- reference arm ------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return                                    # the reference arm is single-rank (no MPI in the image)
    n_gpus = int(os.environ.get("WORLD_SIZE", "1"))
    cells, blocks, extent = workload(args, n_gpus if args.scaling == "weak" else 1)
    base = {"impl": "reference", "unit": "PCG iterations/s", "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "metric": metric_of(args, cells), "config": config_of(args, cells)}
    ref = None
    if args.ref_kind in ("auto", "cuda"):
        try:
            import torch
            assert torch.cuda.is_available()
            torch.cuda.set_device(0)
            ref = RefSolver(args, cells, extent, torch.device("cuda", 0))
        except Exception as e:  # noqa: BLE001
            sys.stderr.write("reference arm: oracle/_ref not usable (%s); falling back to the OpenMP C port\n" % e)
            ref = None
    if ref is None:
        # fallback: the OpenMP C port on the host cores, bounded sample per step
        vals = [cpu_port_sample(cells, args.cpu_sample_grid, args.cpu_sample_iters, args.bc) for _ in range(max(1, min(args.steps, 2)))]
        cb = vals[-1]
        out = dict(base, value=cb["value"], ms_per_step=None, cpu_baseline=cb,
                   reference_kind="OpenMP C port of the reference recurrence (oracle/pcg_ref.c); oracle/_ref/libbbref.so not loadable",
                   e2e={"value": cb["value"], "unit": "PCG iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0})
        print(json.dumps(out))
        return
    import torch
    from bbpcg.grid import grid_shape
    hu, hv, hw = [ref.inp[k].cpu().pin_memory() for k in ("u", "v", "w")]
    hphi = torch.zeros(grid_shape(ref.dom, "Gcc"), dtype=torch.float64).pin_memory()
    for _ in range(args.warmup):
        ref.solve_dev()
    # 1 s period: the reference calls cudaMalloc/cudaFree inside every thrust::inner_product (2 per iteration),
    # which serialise against NVML queries; a 200 ms sampler cost the reference arm ~35 % of its speed
    clocks = ClockSampler(0, period_ms=1000)
    torch.cuda.synchronize()
    clocks.start()
    iters, tot_ms, steps = 0, 0.0, []
    for _ in range(args.steps):
        n, m = ref.solve_dev()
        iters += n; tot_ms += m; steps.append((n, m))
        sys.stderr.write("reference arm: device-resident solve %d iterations %.1f ms\n" % (n, m))
    torch.cuda.synchronize()
    clk = clocks.stop()
    try:
        ref.save_phi(ref_phi_path(args, cells))                  # for the `parity` object of the bbpcg arm
    except Exception as e:  # noqa: BLE001
        sys.stderr.write("reference arm: could not save phi: %s\n" % e)
    k_e2e = max(1, min(args.steps, 3))
    ref.solve_host(hu, hv, hw, hphi)
    e_steps = []
    for _ in range(k_e2e):
        n, m = ref.solve_host(hu, hv, hw, hphi)
        e_steps.append((n, m))
        sys.stderr.write("reference arm: host-buffer solve %d iterations %.1f ms\n" % (n, m))
    # the MEDIAN step: single steps of this arm have been seen 2-5x slower than their neighbours on a fresh box
    n_med, ms_med = sorted(steps, key=lambda t: t[1])[len(steps) // 2]
    ne_med, mse_med = sorted(e_steps, key=lambda t: t[1])[len(e_steps) // 2]
    value = n_med / (ms_med * 1e-3)
    ncell = cells[0] * cells[1] * cells[2]
    peak, peak_src = measured_peak()
    epi = None
    if not args.no_epilogue and hasattr(ref.lib, "bbref_epilogue"):
        from bbpcg.grid import BC_SETS
        # the reference's own epilogue kernels on the phi its solve left on the device (src/bluebottle.c:233-250)
        hp0 = torch.rand(grid_shape(ref.dom, "Gcc"), dtype=torch.float64).pin_memory()
        pbc = (C.c_int * 6)(*BC_SETS[args.bc])
        ems = C.c_float()
        tot_e, n_epi = 0.0, 3
        for i in range(n_epi + 1):
            assert ref.lib.bbref_epilogue(None, ref.P(hp0), C.cast(pbc, C.c_void_p), RHO_F, DT, 0.01, None, None, None, None, None, C.byref(ems)) == 0
            if i:
                tot_e += ems.value
        epi = {"ms_per_call": tot_e / n_epi, "calls_timed": n_epi, "algorithmic_bytes_per_cell": BYTES_EPILOGUE,
               "achieved": BYTES_EPILOGUE * ncell / (tot_e / n_epi * 1e-3) / 1e9, "unit": "GB/s",
               "what": "the reference's pack/unpack + BC_p_*_N + project_u/v/w + update_p_laplacian + update_p + copy_p_p_noghost + "
                       "thrust::reduce + forcing_add_c_const, host sequence of cuda_bluebottle.cu:2495-2589 (CUDA events, default stream)"}
    out = dict(base, value=value, value_mean=iters / (tot_ms * 1e-3), ms_per_step=ms_med, ms_per_step_mean=tot_ms / args.steps,
               value_is="median step (iterations of that step / its CUDA-event time); value_mean is the mean over all timed steps",
               reference_kind="the reference's own cuda_PP_init_jacobi_preconditioner + %s, unmodified kernels recompiled for sm_100a "
                              "(oracle/_ref/libbbref.so), 1 rank on ONE GPU (no MPI in the image)" % ("cuda_PP_cg" if args.parts else "cuda_PP_cg_noparts"),
               iterations_per_step=iters / args.steps,
               cpu_baseline={"value": value, "unit": "PCG iterations/s", "cores": 0, "kind": "reference",
                             "sample": "full workload on ONE GPU: the reference has no CPU implementation of this path "
                                       "(BASELINE.json north_star); this is its CUDA path (oracle/_ref/libbbref.so)"},
               e2e={"value": ne_med / (mse_med * 1e-3), "unit": "PCG iterations/s",
                    "h2d_bytes_per_step": sum(t.numel() * 8 for t in (hu, hv, hw)), "d2h_bytes_per_step": hphi.numel() * 8,
                    "steps": k_e2e, "ms_per_step": mse_med, "value_is": "median step"},
               roofline_iteration={"bound": "hbm", "achieved": BYTES_ITER * ncell * value / 1e9, "peak": peak, "unit": "GB/s",
                                   "frac": BYTES_ITER * ncell * value / 1e9 / peak, "model": "72 B/cell/iteration, whole step"},
               clocks=clk, gpus_used=1, epilogue=epi)
    print(json.dumps(out))


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_bbpcg(args)


if __name__ == "__main__":
    main()
