#!/usr/bin/env python
"""One rank of a multi-process run (launched by torch.distributed.run from tests/test_multiprocess.py).

--mode host : gloo, no GPU.  Exercises the N>1 HOST path: decomposition per rank, inputs sharded
              without communication, the attach-record all-gather (same code path as
              PoissonSolver.comm_init_torch, with a CUDA-free record), and rank-ordered global
              reductions checked against the 1-block oracle.
--mode gpu  : nccl, one process per GPU.  The real thing: CUDA-IPC peer mapping, in-kernel halo
              pull over NVLink, in-kernel rank-ordered all-reduce; phi gathered on rank 0 and compared
              with the 1-block CPU oracle (the discrete solution is decomposition independent).
Prints "MP_WORKER_OK <json>" on rank 0 when every check passed; any failure raises (non-zero exit).
"""
import argparse
import json
import os
import struct
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "bluebottle-3.0_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def numpy_rhs(dom, u, v, w, rho_f, dt):
    """PP_rhs (src/solver_kernel.cu:152-157,173) on one block with numpy, interior only -> [k, j, i]"""
    from bbpcg.grid import as_ijk
    U, V, W = as_ijk(u, "Gfx"), as_ijk(v, "Gfy"), as_ijk(w, "Gfz")
    n = (dom.xn, dom.yn, dom.zn)
    sl = lambda a, ax, off: a[tuple(slice(1 + (off if d == ax else 0), 1 + n[d] + (off if d == ax else 0)) for d in range(3))]  # noqa: E731
    t = (sl(U, 0, 1) - sl(U, 0, 0)) * (1.0 / dom.dx)
    t = t + (sl(V, 1, 1) - sl(V, 1, 0)) * (1.0 / dom.dy)
    t = t + (sl(W, 2, 1) - sl(W, 2, 0)) * (1.0 / dom.dz)
    t = t * (rho_f / dt)
    return np.ascontiguousarray((-t).transpose(2, 1, 0))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", default="host", choices=["host", "gpu"])
    ap.add_argument("--cells", default="32,32,32")
    ap.add_argument("--blocks", default="")
    ap.add_argument("--bc", default="duct")
    ap.add_argument("--nparts", type=int, default=0)
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    import bbpcg
    from bbpcg import synth
    from bbpcg.grid import BC_SETS
    from bbpcg.solver import gather_records, parse_record, RECORD_FMT
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    cells = tuple(int(v) for v in a.cells.split(","))
    blocks = tuple(int(v) for v in a.blocks.split(",")) if a.blocks else {1: (1, 1, 1), 2: (1, 1, 2), 4: (1, 2, 2), 8: (2, 2, 2)}[world]
    assert blocks[0] * blocks[1] * blocks[2] == world
    extent = (0., 12., 0., 12. * cells[1] / cells[0], 0., 12. * cells[2] / cells[0])
    dec = bbpcg.Decomposition.uniform(extent, cells, blocks, BC_SETS[a.bc])
    dom = dec.doms[rank]
    out = {"mode": a.mode, "world": world, "blocks": blocks, "bc": a.bc}

    if a.mode == "host":
        dist.init_process_group("gloo")
        # (1) attach records travel in rank order through the same gather the GPU path uses
        rec = struct.pack(RECORD_FMT, 0xbb9c6001, rank, dom.xn, dom.yn, dom.zn, local, os.getpid())
        rec += b"\0" * (bbpcg.lib.BLOB_BYTES - len(rec))
        recs = gather_records(rec)
        assert len(recs) == world
        for r, b in enumerate(recs):
            info = parse_record(b)
            assert info["rank"] == r and (info["in"], info["jn"], info["kn"]) == (dec.doms[r].xn, dec.doms[r].yn, dec.doms[r].zn)
        # a record out of order must be refused
        try:
            gather_records(rec, _shuffle_for_test=True)
            raise AssertionError("shuffled records accepted")
        except RuntimeError:
            pass
        # (2) sharded inputs + global reductions: (b,b) summed over ranks in rank order == 1-block oracle
        u, v, w = synth.velocity_star(dom, dec.DOM, dec.bc)
        b = numpy_rhs(dom, u, v, w, 1.0, 1e-3)
        parts = [torch.zeros(2, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(parts, torch.tensor([float((b * b).sum()), float(b.sum())], dtype=torch.float64))
        bb = sum(float(p[0]) for p in parts)          # rank order, like the in-kernel mailbox sum
        sb = sum(float(p[1]) for p in parts)
        if rank == 0:
            from cases import Case
            from oracle import binding as ob
            one = Case(cells, bc=a.bc)
            one.o.rhs(1.0, 1e-3)
            ref = one.o.array(0, ob.RHS_P)[1:-1, 1:-1, 1:-1]
            assert abs(bb - float((ref * ref).sum())) <= 1e-12 * bb
            assert abs(sb) <= 1e-9 * float(np.abs(ref).sum())                   # solvability
            # my block is the matching slice of the 1-block field, bit for bit
            g = dom.Gcc
            i0, j0, k0 = g.get("is") - 1, g.get("js") - 1, g.get("ks") - 1
            assert np.array_equal(b, ref[k0:k0 + dom.zn, j0:j0 + dom.yn, i0:i0 + dom.xn])
            out["bb"] = bb
        # (3) neighbour tables agree across processes: my east's west is me
        nb = torch.tensor([dom.e, dom.w, dom.n, dom.s, dom.t, dom.b], dtype=torch.int64)
        allnb = [torch.zeros(6, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(allnb, nb)
        opp = [1, 0, 3, 2, 5, 4]
        for f in range(6):
            if int(nb[f]) >= 0:
                assert int(allnb[int(nb[f])][opp[f]]) == rank
        dist.barrier()
    else:
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        from cases import Case, rel_l2
        from oracle import binding as ob
        case = Case(cells, blocks=blocks, bc=a.bc, nparts=a.nparts, radius=2.5)     # inputs for every block (small grid)
        s = bbpcg.PoissonSolver(dec, rank, device=local)
        s.comm_init_torch()
        s.set_option("comm_timeout_ms", 8000)
        inp = {k: s.to_device(v) for k, v in case.inputs(rank).items()}
        rhs, phi = s.empty("Gcc"), s.empty("Gcc")
        parts = a.nparts > 0
        s.init_jacobi_preconditioner(inp["flag_u"], inp["flag_v"], inp["flag_w"], inp["phase"] if parts else None)
        # halo exchange on a caller array (mpi_cuda_exchange_Gcc)
        rng = np.random.default_rng(5)
        for r in range(world):
            arr = rng.standard_normal(case.o.array(r, ob.PHI).shape)
            case.o.array(r, ob.PHI)[...] = arr
            if r == rank:
                mine = s.to_device(arr)
        case.o.exchange_Gcc(ob.PHI)
        s.exchange_Gcc(mine)
        assert np.array_equal(mine.cpu().numpy(), case.o.array(rank, ob.PHI)), "halo exchange differs on rank %d" % rank
        # the solve
        if parts:
            res = s.PP_cg(inp["u_star"], inp["v_star"], inp["w_star"], rhs, phi, inp["phase"], inp["phase_shell"])
        else:
            res = s.PP_cg_noparts(inp["u_star"], inp["v_star"], inp["w_star"], rhs, phi)
        res2 = s.PP_cg(inp["u_star"], inp["v_star"], inp["w_star"], rhs, phi, inp["phase"], inp["phase_shell"]) if parts else \
            s.PP_cg_noparts(inp["u_star"], inp["v_star"], inp["w_star"], rhs, phi)
        hist = s.history()
        assert res.status == "converged" and res2.niter == res.niter
        single = Case(cells, bc=a.bc, nparts=a.nparts, radius=2.5)
        ores, ohist = single.solve_oracle()
        assert abs(res.niter - ores.niter) <= 1, (res.niter, ores.niter)
        n = min(len(hist), len(ohist))
        assert np.max(np.abs(hist[:n] - ohist[:n]) / ohist[:n]) < 1e-7
        # every rank took identical decisions
        hh = [None] * world
        dist.all_gather_object(hh, hist.tobytes())
        assert all(h == hh[0] for h in hh)
        # gather phi on rank 0
        blocks_phi = [None] * world
        dist.all_gather_object(blocks_phi, phi.cpu().numpy()[1:-1, 1:-1, 1:-1].copy())
        if rank == 0:
            D = dec.DOM
            full = np.zeros((D.zn, D.yn, D.xn))
            for r in range(world):
                d = dec.doms[r]
                i0, j0, k0 = d.Gcc.get("is") - 1, d.Gcc.get("js") - 1, d.Gcc.get("ks") - 1
                full[k0:k0 + d.zn, j0:j0 + d.yn, i0:i0 + d.xn] = blocks_phi[r]
            err = rel_l2(full, single.o.gather_interior(ob.PHI))
            if res.niter == ores.niter:
                assert err < 1e-10, err
            out.update(niter=res.niter, oracle_niter=ores.niter, rel_l2=err, ms_iter=res.ms_iter, launches=res.launches)
        # the solve epilogue across real GPUs (exchange of phi over NVLink, rank-ordered mean): seeded phi/p0 per block,
        # against the multi-block oracle
        case.seed_epilogue(77)
        ein = case.epilogue_inputs(rank)
        ephi, ep0 = s.to_device(ein["phi"]), s.to_device(ein["p0"])
        un, vn, wn, pn = s.empty("Gfx"), s.empty("Gfy"), s.empty("Gfz"), s.empty("Gcc")
        s.epilogue(ephi, inp["u_star"], inp["v_star"], inp["w_star"], inp["flag_u"], inp["flag_v"], inp["flag_w"], un, vn, wn,
                   ep0, inp["phase"], pn)
        case.o.epilogue(1.0, 1e-3)
        I = (slice(1, -1),) * 3
        assert np.array_equal(ephi.cpu().numpy(), case.o.array(rank, ob.PHI)), "epilogue: phi ghosts differ on rank %d" % rank
        for t, aid in ((un, ob.U), (vn, ob.V), (wn, ob.W)):
            ref = case.o.array(rank, aid)[I]
            assert np.abs(t.cpu().numpy()[I] - ref).max() <= 1e-14 * np.abs(ref).max(), "epilogue: velocity differs on rank %d" % rank
        ref = case.o.array(rank, ob.P)[I]
        assert np.abs(pn.cpu().numpy()[I] - ref).max() <= 1e-12 * np.abs(ref).max(), "epilogue: p differs on rank %d" % rank
        out["epilogue"] = "ok"
        s.close()
        dist.barrier()
    if rank == 0:
        print("MP_WORKER_OK " + json.dumps(out))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
