#!/usr/bin/env python
"""bench.py -- pressure-Poisson PCG throughput on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one complete pressure-Poisson solve through the reference-facing entry point
(cuda_PP_cg_noparts -> PoissonSolver.PP_cg_noparts -> bbpcg_solve): PP_rhs, set-up, the PCG
iteration loop to pp_residual = 1e-6 and the phi write-back, on the synthetic 512^3 FP64 grid
(duct boundary set).  `value` = PCG iterations of all timed steps / their time,
i.e. whole-step iterations/s (set-up included).  N > 1: the same GLOBAL grid decomposed over N
ranks with Bluebottle's decomp.config block rule (strong scaling), one process per GPU, halos and
dot products through NVLink peer memory inside the kernels.

Rank 0 prints ONE JSON line.  Extra objects: `roofline` (dominant kernel k_search_spmv, CUDA
events on the solver's stream, live in the timed region), `roofline_iteration` (72-B model over
the whole iteration loop), `cpu_baseline` (the OpenMP C port of the reference recurrence,
oracle/pcg_ref.c, on the host cores; bounded sample), `e2e` (host buffers -> bbpcg_solve_host ->
host buffer), `clocks`.

--impl reference: the reference's OWN unmodified CUDA kernels + host loop (oracle/_ref/libbbref.so,
compiled from /root/reference/src by oracle/Makefile) on one GPU, same workload, same metric.
The reference has no CPU implementation of this path (BASELINE.json north_star); if that library
cannot be loaded the arm falls back to the OpenMP C port on the host cores and says so.
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

# algorithmic bytes per interior cell.  The COMMITTED model of the iteration is 72 B (BASELINE.md 3,
# SURVEY 8d) and the iteration roofline is always scored against it.  The library's default "recompute" variant
# moves 64: the search kernel does not store q (40 B: r, p_prev, x read; p_new, x written) and the residual
# kernel re-applies the operator to p instead of reading q (24 B: p, r read; r written).
BYTES_SEARCH = {0: 48, 1: 40}
BYTES_RESID = 24
BYTES_ITER = 72
# solve epilogue: project 8 (phi) + 24 (u*,v*,w*) + 12 (int flags) + 24 (u,v,w) = 68; update_p 8 (p0) + 4 (phase) + 8 (p) = 20;
# mean subtraction 16 (p read + write): 104 B per cell
BYTES_EPILOGUE = 104
BLOCKS_FOR = {1: (1, 1, 1), 2: (1, 1, 2), 4: (1, 2, 2), 8: (2, 2, 2)}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="bbpcg", choices=["bbpcg", "reference"])
    ap.add_argument("--grid", type=int, default=512, help="global cells per side")
    ap.add_argument("--bc", default="duct")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"])
    ap.add_argument("--blocks", default="", help="In,Jn,Kn override")
    ap.add_argument("--ty", type=int, default=-1, help="tile rows of the iteration kernels (default: planner)")
    ap.add_argument("--kc", type=int, default=-1, help="planes per z-chunk (default: planner)")
    ap.add_argument("--fixed-iters", type=int, default=0, help=">0: fixed iteration count per step, no stop test")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-epilogue", action="store_true", help="skip the solve-epilogue measurement (project + update_p)")
    ap.add_argument("--no-comm-split", action="store_true", help="N > 1: skip the exposed halo/all-reduce measurement")
    ap.add_argument("--cpu-sample-grid", type=int, default=256)
    ap.add_argument("--cpu-sample-iters", type=int, default=400, help="fixed iterations of the CPU sample (~12 s of CPU work at 256^3)")
    ap.add_argument("--ref-kind", default="auto", choices=["auto", "cuda", "port"])
    ap.add_argument("--opt", action="append", default=[], help="solver tuning option key=value (bbpcg_set_option)")
    return ap.parse_args()


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
    """per-launch DRAM bytes of k_search_spmv from the committed ncu --set full capture, if any"""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f)
    except Exception:
        return None


# ---- CPU baseline (oracle port, OpenMP) --------------------------------------------------------
def cpu_port_sample(grid, sample_grid, iters, bc):
    """The OpenMP C port of the reference recurrence on a bounded sample: a sample_grid^3 block of
    the same discretisation, `iters` fixed iterations; it/s scaled by the cell ratio to the bench grid."""
    from cases import Case
    from oracle import binding as ob
    case = Case((sample_grid,) * 3, bc=bc, omp=True)
    cores = ob.load(True).bbo_omp_threads()
    case.o.iterate_fixed(2)                       # touch pages
    t0 = time.perf_counter()
    case.o.iterate_fixed(iters)
    dt = time.perf_counter() - t0
    its_sample = iters / dt
    scale = (sample_grid / float(grid)) ** 3
    return {"value": its_sample * scale, "unit": "PCG iterations/s", "cores": int(cores), "kind": "port",
            "sample": "%d^3 block (%.4g of the %d^3 cells), %d fixed iterations in %.1f s incl. set-up; "
                      "%.2f it/s on the sample, scaled by the cell ratio" % (sample_grid, scale, grid, iters, dt, its_sample),
            "gbs_72B_model": BYTES_ITER * sample_grid ** 3 * its_sample / 1e9}


# ---- distributed plumbing ----------------------------------------------------------------------
class World:
    def __init__(self, want):
        self.rank = int(os.environ.get("RANK", "0"))
        self.size = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        self.dist = None
        if self.size > 1:
            import torch
            import torch.distributed as dist
            torch.cuda.set_device(self.local)
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

    def max(self, v):
        if not self.dist:
            return v
        import torch
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum(self, v):
        if not self.dist:
            return v
        import torch
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def close(self):
        if self.dist:
            self.dist.barrier()
            self.dist.destroy_process_group()


def workload(args, n):
    if args.blocks:
        blocks = tuple(int(v) for v in args.blocks.split(","))
    else:
        blocks = BLOCKS_FOR[n]
    g = args.grid
    if args.scaling == "weak":
        cells = (g * blocks[0], g * blocks[1], g * blocks[2])
    else:
        cells = (g, g, g)
    L = 12.0
    extent = (0., L * cells[0] / g, 0., L * cells[1] / g, 0., L * cells[2] / g)
    return cells, blocks, extent


# ---- our arm -----------------------------------------------------------------------------------
def run_bbpcg(args):
    import torch
    import bbpcg
    from bbpcg import synth
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
    fu, fv, fw = synth.flags_noparts_torch(dom, dec.DOM, dec.bc, dev)
    u, v, wz = synth.velocity_star_torch(dom, dec.DOM, dec.bc, dev)
    rhs, phi = s.empty("Gcc"), s.empty("Gcc")
    s.init_jacobi_preconditioner(fu, fv, fw)
    del fu, fv, fw
    ncell_rank = dom.xn * dom.yn * dom.zn
    ncell_glob = cells[0] * cells[1] * cells[2]
    kw = dict(rho_f=1.0, dt=1e-3, pp_residual=1e-6, pp_max_iter=2000, fixed_iters=args.fixed_iters)

    for _ in range(args.warmup):
        r = s.PP_cg_noparts(u, v, wz, rhs, phi, **kw)
    recompute = 1
    bytes_search = BYTES_SEARCH[recompute]
    clocks = ClockSampler(w.local) if w.rank == 0 else None
    w.barrier()
    if clocks:
        clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    iters = launches = 0
    ms_iter = ms_setup = 0.0
    kt = {k: 0 for k in ("kt_search_ns", "kt_resid_ns", "kt_refresh_ns", "kt_search_n", "kt_resid_n", "kt_refresh_n")}
    for step in range(args.steps):
        # per-kernel CUDA events (solver stream, around every launch) in the FIRST timed step only: an event between
        # two kernels forbids the programmatic dependent launch the other steps run with
        timed_kernels = step == 0
        if timed_kernels:
            s.set_option("kernel_timing", 1)
        r = s.PP_cg_noparts(u, v, wz, rhs, phi, **kw)      # host-synchronous collective call
        assert r.status == "converged", r
        iters += r.niter; launches += r.launches; ms_iter += r.ms_iter; ms_setup += r.ms_total - r.ms_iter
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
    value = iters / (ms * 1e-3)
    peak, peak_src = measured_peak()
    search_s = w.max(kt["kt_search_ns"] * 1e-9 / max(kt["kt_search_n"], 1))
    resid_s = w.max(kt["kt_resid_ns"] * 1e-9 / max(kt["kt_resid_n"], 1))
    ach = bytes_search * ncell_rank / search_s / 1e9
    tr = traffic_from_profile() or {}
    tr_cells = tr.get("cells_per_launch", 512 ** 3)
    tr_scale = ncell_rank / float(tr_cells)          # the capture is of the 512^3 1-GPU launch; bytes scale with the cells
    def _traffic(key):
        v = tr.get(key)
        return None if v is None else v * tr_scale
    roof = {"kernel": "k_search_tma (recompute variant: q not stored)" if recompute else "k_search_tma", "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
            "traffic": _traffic("k_search_spmv_bytes_per_launch"), "traffic_source": tr.get("source"), "peak_source": peak_src,
            "algorithmic_bytes_per_cell": bytes_search, "cells_per_launch": ncell_rank,
            "avg_launch_us": search_s * 1e6, "launches_timed": kt["kt_search_n"],
            "timed": "CUDA events on the solver stream around every launch of the first timed step",
            "share_of_iteration_loop": kt["kt_search_ns"] * 1e-6 / max(kt_ms_iter, 1e-9)}
    ach2 = BYTES_RESID * ncell_rank / resid_s / 1e9
    roof2 = {"kernel": "k_resid_tma (re-applies the operator to p)" if recompute else "k_resid", "bound": "hbm", "achieved": ach2, "peak": peak, "unit": "GB/s", "frac": ach2 / peak,
             "traffic": _traffic("k_resid_bytes_per_launch"), "algorithmic_bytes_per_cell": BYTES_RESID,
             "avg_launch_us": resid_s * 1e6, "launches_timed": kt["kt_resid_n"]}
    ach_it = BYTES_ITER * ncell_rank * iters / (ms_iter_max * 1e-3) / 1e9
    roof_it = {"bound": "hbm", "achieved": ach_it, "peak": peak, "unit": "GB/s", "frac": ach_it / peak,
               "model": "72 B/cell/iteration (committed model) over the iteration loop only, per GPU",
               "bytes_moved_per_cell": 64 if recompute else 72,
               "achieved_moved": (64 if recompute else 72) * ncell_rank * iters / (ms_iter_max * 1e-3) / 1e9,
               "iter_loop_its": iters / (ms_iter_max * 1e-3), "us_per_iteration": ms_iter_max * 1e3 / max(iters, 1)}

    # ---- exposed communication (N > 1): the same per-rank block solved stand-alone (no peers: no halo pull
    # over NVLink, no cross-rank all-reduce wait), same kernels, fixed iteration count ----------------------
    comm = None
    if w.size > 1 and not args.no_comm_split:
        ext1 = (dom.xs, dom.xe, dom.ys, dom.ye, dom.zs, dom.ze)
        dec1 = bbpcg.Decomposition.uniform(ext1, (dom.xn, dom.yn, dom.zn), (1, 1, 1), BC_SETS[args.bc])
        s1 = bbpcg.PoissonSolver(dec1, 0, device=w.local)
        f1 = synth.flags_noparts_torch(dec1.doms[0], dec1.DOM, dec1.bc, dev)
        s1.init_jacobi_preconditioner(*f1)
        u1, v1, w1 = synth.velocity_star_torch(dec1.doms[0], dec1.DOM, dec1.bc, dev)
        rhs1, phi1 = s1.empty("Gcc"), s1.empty("Gcc")
        s1.PP_cg_noparts(u1, v1, w1, rhs1, phi1, fixed_iters=40)
        w.barrier()
        r1 = s1.PP_cg_noparts(u1, v1, w1, rhs1, phi1, fixed_iters=200)
        local_us = w.max(r1.ms_iter * 1e3 / 200)
        rN = s.PP_cg_noparts(u, v, wz, rhs, phi, rho_f=1.0, dt=1e-3, fixed_iters=200)
        coll_us = w.max(rN.ms_iter * 1e3 / 200)
        comm = {"us_per_iteration": coll_us, "us_per_iteration_standalone_block": local_us,
                "exposed_halo_plus_allreduce_us": coll_us - local_us,
                "method": "200 fixed iterations of the decomposed solve vs the same per-rank block solved as a 1x1x1 domain "
                          "(no peer reads, no cross-rank wait); max over ranks; the difference is the exposed halo + all-reduce time"}
        s1.close()
        del u1, v1, w1, rhs1, phi1, f1

    # ---- end to end: pinned host buffers -> bbpcg_solve_host -> pinned host phi -------------------
    e2e = None
    if not args.no_e2e:
        hu, hv, hw = [t.cpu().pin_memory() for t in (u, v, wz)]
        hphi = torch.zeros(tuple(phi.shape), dtype=torch.float64).pin_memory()
        h2d = sum(t.numel() * 8 for t in (hu, hv, hw))
        d2h = hphi.numel() * 8
        s.solve_host(hu, hv, hw, hphi, **kw)                  # allocates the staging arrays
        k_e2e = max(1, min(args.steps, 3))
        w.barrier()
        t0 = time.perf_counter()
        it_e = 0
        for _ in range(k_e2e):
            re = s.solve_host(hu, hv, hw, hphi, **kw)
            it_e += re.niter
        w.barrier()
        te = w.max(time.perf_counter() - t0)
        e2e = {"value": it_e / te, "unit": "PCG iterations/s", "h2d_bytes_per_step": int(w.sum(h2d)),
               "d2h_bytes_per_step": int(w.sum(d2h)), "steps": k_e2e, "ms_per_step": te * 1e3 / k_e2e,
               "api": "bbpcg_solve_host (C ABI): u*,v*,w* pinned host -> device, solve, phi -> pinned host"}
        del hu, hv, hw, hphi

    # ---- the solve epilogue: exchange(phi) + dom_BC_p + cuda_project + cuda_update_p, one fused
    # call per step; reported beside the headline, not part of `value` -------------------------------------
    epi = None
    if not args.no_epilogue:
        from bbpcg.grid import grid_shape
        fu, fv, fw = synth.flags_noparts_torch(dom, dec.DOM, dec.bc, dev)
        un, vn, wn, pn = s.empty("Gfx"), s.empty("Gfy"), s.empty("Gfz"), s.empty("Gcc")
        p0 = torch.rand(grid_shape(dom, "Gcc"), dtype=torch.float64, device=dev)
        phase = torch.full(grid_shape(dom, "Gcc"), -1, dtype=torch.int32, device=dev)
        ea = (phi, u, v, wz, fu, fv, fw, un, vn, wn, p0, phase, pn)
        s.epilogue(*ea)
        w.barrier()
        n_epi = 5
        ms_epi = sum(s.epilogue(*ea) for _ in range(n_epi)) / n_epi
        ms_epi = w.max(ms_epi)
        gbs = BYTES_EPILOGUE * ncell_rank / (ms_epi * 1e-3) / 1e9
        epi = {"ms_per_call": ms_epi, "calls_timed": n_epi, "launches_per_call": 5, "bound": "hbm",
               "algorithmic_bytes_per_cell": BYTES_EPILOGUE, "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak,
               "what": "bbpcg_epilogue (C ABI): mpi_cuda_exchange_Gcc(phi) + cuda_dom_BC_p(phi) + cuda_project + cuda_update_p "
                       "(src/bluebottle.c:233-250) as k_xchg_send/recv + k_bc_p + k_epilogue + k_sub_mean; CUDA events on the solver stream",
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
        del fu, fv, fw, un, vn, wn, pn, p0, phase, uc, vc, wc

    out = {"metric": "Poisson PCG iterations/s (FP64, %d^3)" % args.grid if args.scaling == "strong" else
           "Poisson PCG iterations/s (FP64, %d^3 per GPU)" % args.grid,
           "value": value, "unit": "PCG iterations/s", "n_gpus": w.size, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
           "dtype": "f64", "data": "synthetic",
           "config": {"workload": "synthetic FP64 pressure-Poisson, %dx%dx%d cells, %s boundary set, PP_rhs + Jacobi-PCG to "
                                  "pp_residual 1e-6 (cuda_PP_cg_noparts)" % (cells + (args.bc,)),
                      "blocks": "%dx%dx%d" % blocks, "cells_per_gpu": ncell_rank, "iterations_per_step": iters / args.steps,
                      "l2": "inputs larger than L2 (%.1f GB of solver vectors per GPU vs 126 MB)" % (5 * 8 * ncell_rank / 1e9),
                      "fixed_iters": args.fixed_iters},
           "impl": "bbpcg", "gpu_launches": int(launches_all), "e2e": e2e, "roofline": roof, "roofline_resid": roof2,
           "roofline_iteration": roof_it, "comm": comm, "clocks": clk, "wall_ms_per_step": wall_ms / args.steps,
           "setup_ms_per_step": ms_setup / args.steps, "epilogue": epi,
           "hbm_gbs_72B_model_whole_step": BYTES_ITER * ncell_glob * value / w.size / 1e9}
    if w.rank == 0 and w.size == 1 and not args.no_cpu_baseline:
        try:
            out["cpu_baseline"] = cpu_port_sample(args.grid, args.cpu_sample_grid, args.cpu_sample_iters, args.bc)
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
    cells, blocks, extent = workload(args, 1)
    n_gpus = int(os.environ.get("WORLD_SIZE", "1"))
    lib = None
    if args.ref_kind in ("auto", "cuda"):
        try:
            import torch
            assert torch.cuda.is_available()
            from cases import load_ref
            lib = load_ref()
        except Exception:
            lib = None
    base = {"impl": "reference", "unit": "PCG iterations/s", "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "metric": "Poisson PCG iterations/s (FP64, %d^3)" % args.grid}
    if lib is None:
        # fallback: the OpenMP C port on the host cores, bounded sample per step
        vals = []
        for _ in range(max(1, min(args.steps, 2))):
            vals.append(cpu_port_sample(args.grid, args.cpu_sample_grid, args.cpu_sample_iters, args.bc))
        cb = vals[-1]
        out = dict(base, value=cb["value"], ms_per_step=None, cpu_baseline=cb,
                   config={"workload": "OpenMP C port of the reference recurrence (oracle/pcg_ref.c); oracle/_ref/libbbref.so not loadable"},
                   e2e={"value": cb["value"], "unit": "PCG iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0})
        print(json.dumps(out))
        return
    import torch
    import bbpcg
    from bbpcg import synth
    from bbpcg.grid import BC_SETS
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    dec = bbpcg.Decomposition.uniform(extent, cells, (1, 1, 1), BC_SETS[args.bc])
    dom = dec.doms[0]
    assert lib.bbref_init(C.byref(dom), C.byref(dec.DOM)) == 0
    fu, fv, fw = synth.flags_noparts_torch(dom, dec.DOM, dec.bc, dev)
    u, v, wz = synth.velocity_star_torch(dom, dec.DOM, dec.bc, dev)
    from bbpcg.grid import grid_shape
    phase = torch.full(grid_shape(dom, "Gcc"), -1, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    P = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    assert lib.bbref_set_inputs_dev(P(fu), P(fv), P(fw), P(phase), P(phase), P(u), P(v), P(wz), 0) == 0
    hu, hv, hw = [t.cpu().pin_memory() for t in (u, v, wz)]
    hphi = torch.zeros(grid_shape(dom, "Gcc"), dtype=torch.float64).pin_memory()
    del fu, fv, fw, u, v, wz, phase
    lib.bbref_solve_host.argtypes = [C.c_void_p] * 4 + [C.c_double] * 3 + [C.c_int, C.c_int, C.POINTER(C.c_int),
                                                                            C.POINTER(C.c_double), C.POINTER(C.c_float)]
    niter, resid, ms = C.c_int(), C.c_double(), C.c_float()

    def solve_dev():
        assert lib.bbref_solve(1.0, 1e-3, 1e-6, 2000, 0, C.byref(niter), C.byref(resid), C.byref(ms)) == 0
        return niter.value, ms.value

    def solve_host():
        assert lib.bbref_solve_host(P(hu), P(hv), P(hw), P(hphi), 1.0, 1e-3, 1e-6, 2000, 0, C.byref(niter),
                                    C.byref(resid), C.byref(ms)) == 0
        return niter.value, ms.value

    for _ in range(args.warmup):
        solve_dev()
    # 1 s period: the reference calls cudaMalloc/cudaFree inside every thrust::inner_product (2 per iteration),
    # which serialise against NVML queries; a 200 ms sampler cost the reference arm ~35 % of its speed
    clocks = ClockSampler(0, period_ms=1000)
    torch.cuda.synchronize()
    clocks.start()
    iters, tot_ms = 0, 0.0
    for _ in range(args.steps):
        n, m = solve_dev()
        iters += n; tot_ms += m
        sys.stderr.write("reference arm: device-resident solve %d iterations %.1f ms\n" % (n, m))
    torch.cuda.synchronize()
    clk = clocks.stop()
    k_e2e = max(1, min(args.steps, 3))
    solve_host()
    it_e, ms_e = 0, 0.0
    for _ in range(k_e2e):
        n, m = solve_host()
        it_e += n; ms_e += m
        sys.stderr.write("reference arm: host-buffer solve %d iterations %.1f ms\n" % (n, m))
    value = iters / (tot_ms * 1e-3)
    ncell = cells[0] * cells[1] * cells[2]
    peak, peak_src = measured_peak()
    epi = None
    if not args.no_epilogue and hasattr(lib, "bbref_epilogue"):
        # the reference's own epilogue kernels on the phi its solve left on the device (src/bluebottle.c:233-250)
        hp0 = torch.rand(grid_shape(dom, "Gcc"), dtype=torch.float64).pin_memory()
        pbc = (C.c_int * 6)(*BC_SETS[args.bc])
        ems = C.c_float()
        tot_e, n_epi = 0.0, 3
        for i in range(n_epi + 1):
            assert lib.bbref_epilogue(None, P(hp0), C.cast(pbc, C.c_void_p), 1.0, 1e-3, 0.01, None, None, None, None, None, C.byref(ems)) == 0
            if i:
                tot_e += ems.value
        epi = {"ms_per_call": tot_e / n_epi, "calls_timed": n_epi, "algorithmic_bytes_per_cell": BYTES_EPILOGUE,
               "achieved": BYTES_EPILOGUE * ncell / (tot_e / n_epi * 1e-3) / 1e9, "unit": "GB/s",
               "what": "the reference's pack/unpack + BC_p_*_N + project_u/v/w + update_p_laplacian + update_p + copy_p_p_noghost + "
                       "thrust::reduce + forcing_add_c_const, host sequence of cuda_bluebottle.cu:2495-2589 (CUDA events, default stream)"}
    out = dict(base, value=value, ms_per_step=tot_ms / args.steps,
               config={"workload": "synthetic FP64 pressure-Poisson, %dx%dx%d cells, %s boundary set, the reference's own "
                                   "cuda_PP_init_jacobi_preconditioner + cuda_PP_cg_noparts, unmodified kernels recompiled for "
                                   "sm_100a, 1 rank (no MPI in the image)" % (cells + (args.bc,)),
                       "iterations_per_step": iters / args.steps},
               cpu_baseline={"value": value, "unit": "PCG iterations/s", "cores": 0, "kind": "reference",
                             "sample": "full workload on ONE GPU: the reference has no CPU implementation of this path "
                                       "(BASELINE.json north_star); this is its CUDA path (oracle/_ref/libbbref.so)"},
               e2e={"value": it_e / (ms_e * 1e-3), "unit": "PCG iterations/s",
                    "h2d_bytes_per_step": sum(t.numel() * 8 for t in (hu, hv, hw)), "d2h_bytes_per_step": hphi.numel() * 8,
                    "steps": k_e2e, "ms_per_step": ms_e / k_e2e},
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
