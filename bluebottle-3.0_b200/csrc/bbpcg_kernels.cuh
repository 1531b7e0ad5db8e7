/* bbpcg_kernels.cuh -- shared device code of the pressure-Poisson PCG path (hand-written sm_100a): the operator,
 * the deterministic grid reduction, the in-kernel rank all-reduce, and every kernel that is not one of the two
 * iteration kernels (those live in bbpcg_search_tma.cuh / bbpcg_resid_tma.cuh).
 *
 * One PCG iteration is TWO kernels (the reference uses 3 kernels + 2 Thrust reductions + 12 pack/unpack kernels +
 * 4 host syncs, src/cuda_solver.cu:196-263):
 *
 *   k_search_tma    p = z + beta p  (z = r*invM recomputed from the 1-byte mask)   [PP_update_search]
 *                   x += alpha_prev p_prev   (lazy phi update)                      [PP_update_soln_resid, phi part]
 *                   q = -A p  (7-point, flag^2 / phase coefficients), (p,q) partial [PP_spmv_shared_load(_noparts)]
 *                   last CTA: rank-ordered all-reduce, alpha                        [inner_product + MPI_Allreduce]
 *   k_resid_tma     q re-applied to the p just written; r -= alpha q; (r, r*invM) partial
 *                   [PP_update_soln_resid r/z part, inner_product, MPI_Allreduce]; last CTA: stop test, beta.
 *
 * Traffic actually moved: 40 B + 24 B = 64 B per cell per iteration (+2 mask bytes); the committed model every
 * roofline figure is scored against is 72 B (docs/bytes_model.md).  All scalars live in device memory
 * (struct Scal); a finished solve turns every later launch into a no-op through Scal::done.
 */
#ifndef BBPCG_KERNELS_CUH
#define BBPCG_KERNELS_CUH

#include "bbpcg_internal.h"

typedef unsigned char u8;

/* ------------------------------------------------------------------------------------ */
/* Jacobi diagonal from the 6 flag^2 bits: PP_jacobi_init, src/solver_kernel.cu:73-80.
 * invM = -1/M, M = -idx2(fE^2+fW^2) - idy2(fN^2+fS^2) - idz2(fT^2+fB^2).  128 entries:
 * masks without any flag bit (dead cells: ghosts behind walls) give 0 so that z = p = 0 there; bit 6 (FM_NEAR) is ignored. */
__global__ void k_build_tab(double *tab, double idx2, double idy2, double idz2)
{
  const int m = threadIdx.x;
  if (m >= 128) return;
  double v = 0.;
  if (m & 63) {
    int e = (m & FM_E) != 0, w = (m & FM_W) != 0, n = (m & FM_N) != 0, s = (m & FM_S) != 0,
        t = (m & FM_T) != 0, b = (m & FM_B) != 0;
    double M = -idx2 * (double)(e + w) - idy2 * (double)(n + s) - idz2 * (double)(t + b);
    v = -1. / M;
  }
  tab[m] = v;
}

/* every CTA copies the 1 KB table into shared memory (L2 hit after the first CTA) */
__device__ __forceinline__ void fill_invM_table(double *tab, const Dev &d)
{
  for (int m = threadIdx.x; m < 128; m += blockDim.x) tab[m] = __ldg(d.invM_tab + m);
}

/* -A p at one cell.  The three axis terms  X = fE^2 (pE - pC) - fW^2 (pC - pW), Y, Z  are formed per path and combined at
 * ONE code site with explicit roundings, so that the search kernel and the residual kernel (which re-applies the operator
 * instead of reading a stored q) produce the same bits whatever the compiler would have contracted:
 *     -A p = -idx2 X - idy2 Y - idz2 Z                                  src/solver_kernel.cu:824-829 */
__device__ __forceinline__ double stencil_combine(const Dev &d, double X, double Y, double Z)
{
  return __fma_rn(-d.idz2, Z, __fma_rn(-d.idy2, Y, __dmul_rn(-d.idx2, X)));
}

/* every flag of the cell is 1 (the common case away from walls and particles): multiplying by 1.0 is exact, so these
 * plain differences equal the flagged forms below bit for bit */
__device__ __forceinline__ double stencil_plain(const Dev &d, double pC, double pE, double pW, double pN, double pS, double pT, double pB)
{
  return stencil_combine(d, __dsub_rn(__dsub_rn(pE, pC), __dsub_rn(pC, pW)), __dsub_rn(__dsub_rn(pN, pC), __dsub_rn(pC, pS)),
                         __dsub_rn(__dsub_rn(pT, pC), __dsub_rn(pC, pB)));
}

/* A flag^2 in {0,1} times a difference: 1 * x is x and 0 * x is a zero, so a SELECT gives the product's value without an
 * FP64 multiply (the sign of a zero is the only thing that can differ, and it never reaches a stored value: a zero term
 * changes neither q = -A p, nor the dot products, nor r -= alpha q). */
__device__ __forceinline__ double flagged(unsigned bit, double x) { return bit ? x : 0.; }

/* noparts operator: src/solver_kernel.cu:824-829 (same association) */
__device__ __forceinline__ double stencil_noparts(const Dev &d, unsigned m, double pC, double pE, double pW,
                                                  double pN, double pS, double pT, double pB)
{
  return stencil_combine(d, __dsub_rn(flagged(m & FM_E, __dsub_rn(pE, pC)), flagged(m & FM_W, __dsub_rn(pC, pW))),
                         __dsub_rn(flagged(m & FM_N, __dsub_rn(pN, pC)), flagged(m & FM_S, __dsub_rn(pC, pS))),
                         __dsub_rn(flagged(m & FM_T, __dsub_rn(pT, pC)), flagged(m & FM_B, __dsub_rn(pC, pB))));
}

/* the particle factors of cell g (P-layout offset) from the solid bits of the cell and of its six neighbours */
__device__ __forceinline__ unsigned pm_of_neighbours(const u8 *__restrict__ fmask, long long g, const Layout &L)
{
  return ((fmask[g] & FM_SOLID) ? PM_C : 0u) | ((fmask[g + 1] & FM_SOLID) ? PM_E : 0u) | ((fmask[g - 1] & FM_SOLID) ? PM_W : 0u) |
         ((fmask[g + L.px] & FM_SOLID) ? PM_N : 0u) | ((fmask[g - L.px] & FM_SOLID) ? PM_S : 0u) |
         ((fmask[g + L.ps] & FM_SOLID) ? PM_T : 0u) | ((fmask[g - L.ps] & FM_SOLID) ? PM_B : 0u);
}

/* with particle masking: src/solver_kernel.cu:683-707.  The reference multiplies the centre value by pf{x,y,z} (1 in a
 * fluid cell, -d?^2/6 in a solid one) and every neighbour value by pf{e,w,n,s,t,b} in {0,1} (0 toward a solid neighbour and
 * on every coupling of a solid row): one rounded product for the centre of a solid cell, selects for everything else.
 * pm == 0 (fluid cell, no solid neighbour) reduces to the noparts form bit for bit. */
__device__ __forceinline__ double stencil_parts(const Dev &d, unsigned m, unsigned pm, double pC, double pE,
                                                double pW, double pN, double pS, double pT, double pB)
{
  const bool solid = (pm & PM_C) != 0;
  const double cx = solid ? __dmul_rn(-d.dx2_6, pC) : pC, cy = solid ? __dmul_rn(-d.dy2_6, pC) : pC,
               cz = solid ? __dmul_rn(-d.dz2_6, pC) : pC;
  const double zE = (pm & (PM_C | PM_E)) ? 0. : pE, zW = (pm & (PM_C | PM_W)) ? 0. : pW;
  const double zN = (pm & (PM_C | PM_N)) ? 0. : pN, zS = (pm & (PM_C | PM_S)) ? 0. : pS;
  const double zT = (pm & (PM_C | PM_T)) ? 0. : pT, zB = (pm & (PM_C | PM_B)) ? 0. : pB;
  return stencil_combine(d, __dsub_rn(flagged(m & FM_E, __dsub_rn(zE, cx)), flagged(m & FM_W, __dsub_rn(cx, zW))),
                         __dsub_rn(flagged(m & FM_N, __dsub_rn(zN, cy)), flagged(m & FM_S, __dsub_rn(cy, zS))),
                         __dsub_rn(flagged(m & FM_T, __dsub_rn(zT, cz)), flagged(m & FM_B, __dsub_rn(cz, zB))));
}

/* ------------------------------------------------------------------------------------ */
/* Deterministic grid reduction of NV values: warp shuffle tree -> per-warp slots -> CTA
 * partial in a fixed slot -> the LAST CTA to arrive sums the slots in a fixed order.
 * Returns true (all threads) only in that last CTA; tot[] valid in thread 0. */
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}

template <int NV>
__device__ __forceinline__ double block_sum(double v, double *sh /*[32]*/)
{
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  double s = 0.;
  if (threadIdx.x == 0) for (int w = 0; w < nw; w++) s += sh[w];
  return s;
}

/* Two levels so that kernels made of many tiny CTAs stay cheap: the last CTA of each group of
 * BB_GROUP CTAs sums that group's partials, the last group to finish sums the group sums.  Both
 * orders are fixed, so the result is bit-identical run to run. */
/* programmatic dependent launch (PDL): the NEXT kernel of the stream may be launched while this one drains;
 * it must not touch anything this kernel (or its predecessors) produce before pdl_wait(). */
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

/* This thread's share of a fixed-order sum over n partials written by OTHER CTAs (index i goes to thread i % blockDim.x, a
 * thread adds its values in index order: the result depends on n and the block size only).  The loads of a batch of 8 are
 * issued before the first add: written as `s += __ldcg(..)` in a plain loop the compiler kept ONE load in flight, i.e. one L2
 * round trip (~0.9 us) per 288 partials on the critical path of every iteration kernel -- 2.7 us of the 4.7 us tail at 256^3
 * (814 items, profiles/r02u_trace.json) and ~18 us per kernel at 512^3 (5632 items). */
__device__ __forceinline__ double strided_partial_sum(const double *__restrict__ part, int n)
{
  constexpr int U = 8;
  const int nt = blockDim.x;
  double s = 0.;
  for (int base = threadIdx.x; base < n; base += U * nt) {
    double v[U];
#pragma unroll
    for (int u = 0; u < U; u++) { const int i = base + u * nt; v[u] = i < n ? __ldcg(part + i) : 0.; }
#pragma unroll
    for (int u = 0; u < U; u++) s += v[u];
  }
  return s;
}

template <int NV>
__device__ bool grid_reduce(const Dev &d, double (&v)[NV], int bid, int nblocks, double (&tot)[NV], bool peer_stores)
{
  __shared__ double sh[32];
  __shared__ int s_last;
  double part[NV];
#pragma unroll
  for (int n = 0; n < NV; n++) part[n] = block_sum<NV>(v[n], sh);
  if (peer_stores) __threadfence_system();       /* halo stores visible before we count in */
  __syncthreads();
  /* small grids: ONE group = one level (fewer dependent atomics/fences in the tail of every kernel) */
  const int G = nblocks <= BB_FLAT_MAX ? nblocks : BB_GROUP;
  const int grp = bid / G, ngrp = (nblocks + G - 1) / G;
  const int gsize = min(G, nblocks - grp * G);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int n = 0; n < NV; n++) d.partials[n * BB_MAXBLOCKS + bid] = part[n];
    /* device scope is enough here even when peers will read what this CTA wrote: the last CTA
     * observes this release through the counter and issues ONE system-scope fence before it
     * signals the peers (rank_allreduce); causality order is transitive across scopes.  A
     * fence.sys in each of the ~10^4..10^5 CTAs of k_resid cost 200 us per iteration at N = 2. */
    __threadfence();
    unsigned t = atomicAdd(d.counter + 4 + grp, 1u);
    s_last = (t == (unsigned)(gsize - 1));
  }
  __syncthreads();
  if (!s_last) return false;
  __threadfence();
#pragma unroll
  for (int n = 0; n < NV; n++) {
    part[n] = block_sum<NV>(strided_partial_sum(d.partials + n * BB_MAXBLOCKS + grp * G, gsize), sh);
  }
  __syncthreads();
  if (ngrp == 1) {
    if (threadIdx.x == 0) {
      d.counter[4] = 0u;
#pragma unroll
      for (int n = 0; n < NV; n++) tot[n] = part[n];
    }
    return true;
  }
  if (threadIdx.x == 0) {
#pragma unroll
    for (int n = 0; n < NV; n++) d.gpartials[n * BB_MAXGROUPS + grp] = part[n];
    d.counter[4 + grp] = 0u;
    __threadfence();
    unsigned t = atomicAdd(d.counter, 1u);
    s_last = (t == (unsigned)(ngrp - 1));
  }
  __syncthreads();
  if (!s_last) return false;
  __threadfence();
#pragma unroll
  for (int n = 0; n < NV; n++) {
    tot[n] = block_sum<NV>(strided_partial_sum(d.gpartials + n * BB_MAXGROUPS, ngrp), sh);
  }
  if (threadIdx.x == 0) *d.counter = 0u;
  return true;
}

/* Rank-ordered all-reduce of up to 2 doubles through peer-mapped mailboxes (replaces
 * MPI_Allreduce, src/cuda_solver.cu:152,170,205,232).  Called by the last CTA only.  Every
 * rank writes its partial into every rank's mailbox over NVLink, then waits for the N entries of
 * its own mailbox and adds them in rank order: the sum is bit-identical on all ranks and run to
 * run.  nv == 0 makes it a barrier.  Threads 0..nranks-1 take part.
 *
 * Wire format ("LL": data and flag travel in the SAME 8-byte store, which is atomic, so no fence
 * is needed between payload and flag on either side): each double is sent as two u64 words
 * { 32 data bits | 32-bit sequence tag }.  A receiver spins until all four words of a sender
 * carry the current tag.  One NVLink one-way latency per all-reduce instead of
 * store / fence.sys / flag / spin / fence.sys.
 *
 * release = true: this kernel wrote data that PEERS will read after the barrier (r in the former pull
 * model, ghost pushes): thread 0 -- which has observed every CTA's device-scope release through
 * the counter -- issues the one system-scope fence before anything is signalled. */
__device__ __forceinline__ void fence_acq_rel_sys() { asm volatile("fence.acq_rel.sys;" ::: "memory"); }
__device__ __forceinline__ void st_relaxed_sys(unsigned long long *p, unsigned long long v)
{ asm volatile("st.relaxed.sys.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ unsigned long long ld_relaxed_sys(const unsigned long long *p)
{ unsigned long long v; asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v; }

__device__ __forceinline__ void rank_allreduce(const Dev &d, double *v /* thread 0's values, in/out */, int nv, bool release,
                                               const unsigned long long *seq_cur = nullptr /* thread 0: Scal::seq read earlier (off the critical path) */)
{
  const Comm &c = d.comm;
  if (c.nranks <= 1) return;
  __shared__ unsigned long long s_w[4];
  __shared__ double s_in[BB_MAXR][2];
  __shared__ unsigned long long s_seq;
  if (threadIdx.x == 0) {
    const unsigned long long seq = (seq_cur ? *seq_cur : d.sc->seq) + 1ull;
    d.sc->seq = seq;
    const unsigned long long tag = ((seq % 0xffffffffull) + 1ull) << 32;       /* never 0 */
    const unsigned long long b0 = (unsigned long long)__double_as_longlong(nv > 0 ? v[0] : 0.);
    const unsigned long long b1 = (unsigned long long)__double_as_longlong(nv > 1 ? v[1] : 0.);
    s_w[0] = (b0 & 0xffffffffull) | tag; s_w[1] = (b0 >> 32) | tag;
    s_w[2] = (b1 & 0xffffffffull) | tag; s_w[3] = (b1 >> 32) | tag;
    s_seq = seq;
    if (release) fence_acq_rel_sys();
  }
  __syncthreads();
  const unsigned long long seq = s_seq;
  const unsigned long long tag = s_w[0] >> 32;
  const int slot = (int)(seq & (BB_NSLOT - 1));
  const int t = threadIdx.x;
  if (t < c.nranks) {
    unsigned long long *dst = c.mbox[t] + (size_t)(slot * BB_MAXR + c.rank) * 4;
#pragma unroll
    for (int w = 0; w < 4; w++) st_relaxed_sys(dst + w, s_w[w]);
    /* wait for rank t's contribution in my own mailbox */
    const unsigned long long *src = c.mbox[c.rank] + (size_t)(slot * BB_MAXR + t) * 4;
    unsigned long long w0, w1, w2, w3;
    const long long t0 = clock64();
    bool ok = true;
    for (;;) {
      w0 = ld_relaxed_sys(src); w1 = ld_relaxed_sys(src + 1); w2 = ld_relaxed_sys(src + 2); w3 = ld_relaxed_sys(src + 3);
      if ((w0 >> 32) == tag && (w1 >> 32) == tag && (w2 >> 32) == tag && (w3 >> 32) == tag) break;
      if (c.timeout_cycles > 0 && clock64() - t0 > c.timeout_cycles) { ok = false; break; }      /* a peer is gone */
    }
    s_in[t][0] = __longlong_as_double((long long)((w0 & 0xffffffffull) | (w1 << 32)));
    s_in[t][1] = __longlong_as_double((long long)((w2 & 0xffffffffull) | (w3 << 32)));
    if (!ok) { d.sc->comm_timeout = 1; *c.host_flag = 1; }    /* sticky: the ranks now disagree, the solver object is dead */
    if (release) fence_acq_rel_sys();  /* acquire side: peers' released data (r, ghost pushes) is read by LATER kernels; a plain sum needs none */
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0., b = 0.;
    for (int p = 0; p < c.nranks; p++) { a += s_in[p][0]; b += s_in[p][1]; }
    if (nv > 0) v[0] = a;
    if (nv > 1) v[1] = b;
  }
  __syncthreads();
}

/* store a boundary value into the neighbours' ghost cells (what mpi_cuda_exchange_Gcc's
 * pack -> MPI_Put -> unpack achieves, src/mpi_comm.c:257-315, faces only).  `which` selects
 * the neighbour array: 0 r, 1 x.  Returns true if a (possibly remote) store was made. */
__device__ __forceinline__ bool push_halo(const Dev &d, int which, int i, int j, int k, double val)
{
  const Layout &L = d.L;
  bool any = false;
#define BB_PUSH(F, COND, II, JJ, KK)                                            \
  if (COND) {                                                                   \
    const NbrFace &nf = d.halo.f[F];                                            \
    double *base = which == 0 ? nf.r : nf.x;                                    \
    if (base) { base[pidx(nf.L, II, JJ, KK)] = val; any = true; }               \
  }
  BB_PUSH(0, i == L.in, 0, j, k)               /* my east face  -> east neighbour's west ghost  */
  BB_PUSH(1, i == 1, nf.L.in + 1, j, k)        /* my west face  -> west neighbour's east ghost  */
  BB_PUSH(2, j == L.jn, i, 0, k)
  BB_PUSH(3, j == 1, i, nf.L.jn + 1, k)
  BB_PUSH(4, k == L.kn, i, j, 0)
  BB_PUSH(5, k == 1, i, j, nf.L.kn + 1)
#undef BB_PUSH
  return any;
}

/* in-plane offset (i+XOFF) + j*px of this block -> correction for a neighbour whose pitch differs */
__device__ __forceinline__ long long j_px_fix(const Layout &L, const Layout &N, int goff)
{
  const int j = goff / L.px;
  return (long long)j * (L.px - N.px);
}

/* launch arguments of the two iteration kernels (bbpcg_search_tma.cuh, bbpcg_resid_tma.cuh) */
struct SearchArgs {
  int nbx, nby, nbz;   /* tiles in x, y; z-chunks: chunk c owns planes d.ztab[2c] .. d.ztab[2c+1] */
  int ty;              /* owned rows per tile (1..8): chosen by the host planner so that the CTA count fills the SM slots */
  const double *rhs;   /* refresh form of k_resid_tma only: the caller's right-hand side (Gcc s3b) and its strides */
  int s1b, s2b;
  int launch;          /* trace build: running launch number */
  int nitems;          /* nbx * nby * nbz work items, claimed dynamically by the resident CTAs */
  int producer;        /* the thread that issues the TMA loads: BB_PRODUCER (lane 0 of the 9th warp; default) or 0 (option tma_warp 0) */
};

/* ---- persistent, dynamically balanced work distribution of the two iteration kernels -------------------------------
 * A WORK ITEM is one (x-tile, y-tile, z-chunk): item = bx + nbx (by + nby cz).  The kernels run as ONE wave of resident
 * CTAs; a CTA's first item is its block index, every further one is claimed from an atomic counter in item order, and the
 * TMA producer keeps its D-plane lead ACROSS item boundaries, so a CTA streams without a pipeline refill.  Why: with a
 * static one-wave partition the 256^3 block of an 8-GPU run straggled -- identical CTAs took 84 .. 135 us depending on the
 * SM group they landed on (profiles/r02fg_trace_per_cta.jsonl: the two CTAs of an SM finish together, SM groups differ by
 * 40 %), the kernel ended with the slowest CTA and HBM idled behind it; hardware-scheduled small CTAs pay a launch + ramp
 * per CTA instead.  Dot products stay bit-reproducible under any assignment: every item stores ITS partial in its own
 * slot and the last CTA adds the slots in item order. */
#define BB_CLAIM_SEARCH 1          /* Dev::counter[1], [2]: claim counters of the two kernels; [3]: finished CTAs */
#define BB_CLAIM_RESID 2
#define BB_CLAIM_DONE 3
#define BB_QN 4                    /* item queue between the producer thread and the consumers (the producer is <= 2 items ahead) */

struct ItemGeom { int bx, by, k0, k1, nplanes; };

__device__ __forceinline__ ItemGeom decode_item(const Dev &d, const SearchArgs &a, int item)
{
  ItemGeom g;
  g.bx = item % a.nbx;
  const int t = item / a.nbx;
  g.by = t % a.nby;
  const int cz = t / a.nby;
  const int2 kk = __ldg(reinterpret_cast<const int2 *>(d.ztab) + cz);   /* host-written table: safe before pdl_wait() */
  g.k0 = kk.x;
  g.k1 = kk.y;
  g.nplanes = g.k1 - g.k0 + 3;                          /* planes k0-1 .. k1+1 */
  return g;
}

/* does the item own cells on a block face that has a neighbour (another rank, or this block itself for a periodic wrap)?
 * The residual kernel pushes the new r of those cells into the neighbour's ghost slots. */
__device__ __forceinline__ bool item_touches_nbr(const Dev &d, const SearchArgs &a, const ItemGeom &ig)
{
  const Layout &L = d.L;
  return (d.halo.f[1].r != nullptr && ig.bx == 0) || (d.halo.f[0].r != nullptr && ig.bx == a.nbx - 1) ||
         (d.halo.f[3].r != nullptr && ig.by == 0) || (d.halo.f[2].r != nullptr && ig.by == a.nby - 1) ||
         (d.halo.f[5].r != nullptr && ig.k0 == 1) || (d.halo.f[4].r != nullptr && ig.k1 == L.kn);
}

/* next item of this CTA (-1: none left) */
__device__ __forceinline__ int claim_item(const Dev &d, const SearchArgs &a, int which)
{
  const int item = (int)gridDim.x + (int)atomicAdd(d.counter + which, 1u);
  return item < a.nitems ? item : -1;
}

/* End of a persistent kernel: the LAST CTA to arrive adds the per-item partials in item order (bit-identical whatever CTA
 * computed which item) and re-arms the counters.  Returns true (all threads) in that CTA only; tot valid in thread 0. */
__device__ bool items_reduce(const Dev &d, int nitems, int which, double &tot, bool peer_stores)
{
  __shared__ double sh[32];
  __shared__ int s_last;
  __syncthreads();
  if (threadIdx.x == 0) {
    /* this CTA's item partials and its stores to r / p / x before the count; stores into PEER memory (the pushed ghost
     * values of r) are released at system scope by the CTA that made them -- once per CTA of a one-wave kernel */
    if (peer_stores) __threadfence_system(); else __threadfence();
    const unsigned t = atomicAdd(d.counter + BB_CLAIM_DONE, 1u);
    s_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last) return false;
  __threadfence();
  tot = block_sum<1>(strided_partial_sum(d.partials, nitems), sh);
  if (threadIdx.x == 0) { d.counter[BB_CLAIM_DONE] = 0u; d.counter[which] = 0u; }
  return true;
}

#ifdef BB_TRACE
__device__ __forceinline__ unsigned long long bb_gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ unsigned bb_smid() { unsigned v; asm volatile("mov.u32 %0, %%smid;" : "=r"(v)); return v; }
#define BB_TRACE_AT(d, a, ev, val) do { if (threadIdx.x == 0 && (d).trace) { const int bid_ = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z); \
  if (bid_ < BB_TRACE_CTAS) (d).trace[((size_t)((a).launch % BB_TRACE_LAUNCHES) * BB_TRACE_CTAS + bid_) * BB_TRACE_EV + (ev)] = (val); } } while (0)
#define BB_STAMP(d, a, ev) BB_TRACE_AT(d, a, ev, bb_gtimer())
#else
#define BB_TRACE_AT(d, a, ev, val) do { } while (0)
#define BB_STAMP(d, a, ev) do { } while (0)
#endif

/* ------------------------------------------------------------------------------------ */
/* end of an iteration, run by ONE thread once the global (r,z) is known: the host logic of src/cuda_solver.cu:231-267
 * (history, stop test, NaN test, iteration bound, beta) on the device-resident scalars */
/* The scalars the tail of an iteration kernel needs, none of which the running kernel changes: thread 0 of every CTA reads
 * them BEFORE it takes its ticket, so that in the last CTA no L2 round trip is left between the sum and the new alpha / beta
 * (each dependent load of Scal cost ~0.8 us on the critical path of every iteration). */
struct IterScal { double rz, bb, tol2, alpha; int q, fixed, max_q; unsigned long long seq; };
__device__ __forceinline__ IterScal load_iter_scal(const Dev &d)
{
  const Scal *sc = d.sc;
  IterScal p;                       /* L2 loads (ld.global.cg): never a line an earlier kernel's CTA left in this SM's L1 */
  p.rz = __ldcg(&sc->rz); p.bb = __ldcg(&sc->bb); p.tol2 = __ldcg(&sc->tol2); p.alpha = __ldcg(&sc->alpha);
  p.q = __ldcg(&sc->q); p.fixed = __ldcg(&sc->fixed); p.max_q = __ldcg(&sc->max_q); p.seq = __ldcg(&sc->seq);
  return p;
}

__device__ __forceinline__ void finish_iteration(const Dev &d, double rz_new, bool refreshed, const IterScal &p)
{
  Scal *sc = d.sc;
  const int qn = p.q + 1;
  sc->q = qn;
  if (qn < BB_HIST_CAP) d.history[qn] = rz_new;
  sc->alpha_x = refreshed ? 0. : p.alpha;
  /* the time-out flag is set by the all-reduce that just ran: the one load that cannot be made early (several ranks only) */
  if (d.comm.nranks > 1 && sc->comm_timeout) { sc->done = 1; sc->status = BBPCG_COMM_TIMEOUT; sc->resid = sqrt(rz_new) / sqrt(p.bb); return; }
  if (!p.fixed && rz_new <= p.tol2 * p.bb) {                     /* :235 */
    sc->done = 1; sc->status = BBPCG_CONVERGED; sc->resid = sqrt(rz_new) / sqrt(p.bb);
  } else if (!p.fixed && isnan(rz_new)) {                        /* :245 */
    sc->done = 1; sc->status = BBPCG_NAN; sc->resid = rz_new;
  } else if (!p.fixed && qn >= p.max_q) {                        /* :192,:271 */
    sc->done = 1; sc->status = BBPCG_MAXITER; sc->resid = sqrt(rz_new) / sqrt(p.bb);
  } else {
    sc->beta = rz_new / p.rz;                                    /* :256 */
    sc->rz = rz_new;                                             /* :267 */
  }
}

/* 256-bit global accesses (LDG.E.256 / STG.E.256 on sm_100a): one x-row chunk of 4 cells */
struct d4 { double a, b, c, d; };
__device__ __forceinline__ d4 ld256(const double *p)
{
  d4 v;
  asm("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.a), "=d"(v.b), "=d"(v.c), "=d"(v.d) : "l"(p));
  return v;
}
__device__ __forceinline__ d4 ld256_stream(const double *p)      /* read once, then dead: do not keep in L1 */
{
  d4 v;
  asm("ld.global.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.a), "=d"(v.b), "=d"(v.c), "=d"(v.d) : "l"(p));
  return v;
}
__device__ __forceinline__ void st256(double *p, const d4 &v)
{
  asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" :: "l"(p), "d"(v.a), "d"(v.b), "d"(v.c), "d"(v.d) : "memory");
}

struct ResidArgs { int cpr; int ncb; int npass; int ppc; };   /* k_refresh_x4: chunks per row, column blocks, passes, passes per CTA */

/* ------------------------------------------------------------------------------------ */
/* Every-50th-iteration true-residual refresh, src/cuda_solver.cu:209-223:
 * k_refresh_x4 : phi += alpha p (PP_update_solution, solver_kernel.cu:864-881) + halo of phi
 * then the REFRESH form of k_resid_tma: r = b - (-A)phi ; z ; (r,z)  (SpMV + PP_update_residual, :883-904) */
/* phi += alpha p as a streaming kernel: many tiny CTAs, 256-bit accesses, every load of a pass issued before its first
 * use; only boundary cells take the (possibly remote) halo-push path */
template <int XT, int UNR>
__global__ void __launch_bounds__(128, 4) k_refresh_x4(const __grid_constant__ Dev d, const ResidArgs a)
{
  constexpr int NT = 128, YT = NT / XT;
  const Layout L = d.L;
  Scal *sc = d.sc;
  if (sc->done) return;
  const double alpha = sc->alpha;
  const double *__restrict__ pcur = d.P[(sc->q + 1) & 1];     /* p of the iteration in flight */
  double *__restrict__ x = d.x;
  const unsigned nrows = (unsigned)L.jn * (unsigned)L.kn;
  const int tx = threadIdx.x % XT, ty = threadIdx.x / XT;
  bool pushed = false;
  const int pass0 = blockIdx.x * a.ppc, pass1 = min(pass0 + a.ppc, a.npass);
  for (int pass = pass0; pass < pass1; pass++) {
    const int cb = pass % a.ncb, rg = pass / a.ncb;
    const int c = cb * XT + tx;
    const unsigned row0 = (unsigned)rg * (YT * UNR) + ty;
    d4 xv[UNR], pv[UNR];
    long long g[UNR];
    int jj[UNR], kk[UNR];
#pragma unroll
    for (int u = 0; u < UNR; u++) {
      const unsigned row = row0 + u * YT;
      const unsigned k0 = row / (unsigned)L.jn;
      jj[u] = (int)(row - k0 * (unsigned)L.jn) + 1; kk[u] = (int)k0 + 1;
      g[u] = (long long)kk[u] * L.ps + (long long)jj[u] * L.px + (BB_XOFF + 1) + 4 * c;
      if (row < nrows && c < a.cpr) { xv[u] = ld256(x + g[u]); pv[u] = ld256(pcur + g[u]); }
      else kk[u] = -1;
    }
    const int nv = min(4, L.in - 4 * c);
#pragma unroll
    for (int u = 0; u < UNR; u++) {
      if (kk[u] < 0) continue;
      d4 v = xv[u];
      v.a += alpha * pv[u].a; v.b += alpha * pv[u].b; v.c += alpha * pv[u].c; v.d += alpha * pv[u].d;   /* solver_kernel.cu:876 */
      if (nv == 4) st256(x + g[u], v);
      else { double *xp = x + g[u]; xp[0] = v.a; if (nv > 1) xp[1] = v.b; if (nv > 2) xp[2] = v.c; }
      const int j = jj[u], k = kk[u], i0 = 4 * c + 1;
      if (d.any_nbr && (j == 1 || j == L.jn || k == 1 || k == L.kn || i0 == 1 || i0 + nv - 1 == L.in)) {
        const double e[4] = { v.a, v.b, v.c, v.d };
        for (int n = 0; n < nv; n++) {
          const int i = i0 + n;
          if (j == 1 || j == L.jn || k == 1 || k == L.kn || i == 1 || i == L.in) pushed |= push_halo(d, 1, i, j, k, e[n]);
        }
      }
    }
  }
  double v[1] = { 0. }, tot[1];
  if (grid_reduce<1>(d, v, blockIdx.x, gridDim.x, tot, pushed)) rank_allreduce(d, tot, 0, true);   /* barrier */
}

/* ------------------------------------------------------------------------------------ */
/* Set-up: r = b, x = 0, (b,b), (r,z) -- PP_cg_init (src/solver_kernel.cu:258-283) and the two
 * inner products of src/cuda_solver.cu:151,169.  p buffers are zeroed by the host (memset), so
 * the first k_search_tma computes p = z + 0*0. */
template <int NT>
__global__ void __launch_bounds__(NT) k_init(const Dev d, const double *__restrict__ rhs_s3b, int s1b, int s2b,
                                             double tol2, int max_q, int fixed)
{
  __shared__ double tab[128];
  fill_invM_table(tab, d);
  __syncthreads();
  const Layout L = d.L;
  const long long nrows = (long long)L.jn * L.kn;
  double bb = 0., rz = 0.;
  bool pushed = false;
  for (long long row = blockIdx.x; row < nrows; row += gridDim.x) {
    const int j = (int)(row % L.jn) + 1, k = (int)(row / L.jn) + 1;
    const long long base = pidx(L, 1, j, k);
    for (int i = threadIdx.x + 1; i <= L.in; i += NT) {
      const long long g = base + (i - 1);
      const double b = rhs_s3b[i + (long long)j * s1b + (long long)k * s2b];
      const double z = b * tab[d.fmask[g] & 127u];
      bb += b * b;  rz += b * z;
      d.r[g] = b;
      if (d.any_nbr && (i == 1 || i == L.in || j == 1 || j == L.jn || k == 1 || k == L.kn)) pushed |= push_halo(d, 0, i, j, k, b);   /* ghost r of the neighbours */
    }
  }
  double v[2] = { bb, rz }, tot[2];
  if (grid_reduce<2>(d, v, blockIdx.x, gridDim.x, tot, pushed)) {
    rank_allreduce(d, tot, 2, true);
    if (threadIdx.x == 0) {
      Scal *sc = d.sc;
      sc->bb = tot[0]; sc->rz = tot[1]; sc->rz0 = tot[1];
      sc->alpha = 0.; sc->alpha_x = 0.; sc->beta = 0.; sc->pAp = 0.; sc->resid = 0.;
      sc->tol2 = tol2; sc->max_q = max_q; sc->fixed = fixed; sc->q = 0;
      sc->status = BBPCG_CONVERGED;
      d.history[0] = tot[1];
      const double RHS_TOL = 1.e-8;                                          /* cuda_solver.cu:176 */
      if (!fixed && tot[0] < RHS_TOL * RHS_TOL) { sc->done = 1; sc->status = BBPCG_TINY_RHS; }
      else sc->done = sc->comm_timeout ? 1 : 0;
      if (sc->comm_timeout) sc->status = BBPCG_COMM_TIMEOUT;
    }
  }
}

/* phi (caller's Gcc s3b array) = x + alpha_x * p  for interior cells */
template <int NT>
__global__ void __launch_bounds__(NT) k_finish(const Dev d, double *__restrict__ phi_s3b, int s1b, int s2b)
{
  const Scal *sc = d.sc;
  const Layout L = d.L;
  const double ax = sc->alpha_x;
  const double *__restrict__ pcur = d.P[sc->q & 1];
  const long long nrows = (long long)L.jn * L.kn;
  for (long long row = blockIdx.x; row < nrows; row += gridDim.x) {
    const int j = (int)(row % L.jn) + 1, k = (int)(row / L.jn) + 1;
    const long long base = pidx(L, 1, j, k);
    for (int i = threadIdx.x + 1; i <= L.in; i += NT) {
      const long long g = base + (i - 1);
      phi_s3b[i + (long long)j * s1b + (long long)k * s2b] = d.x[g] + ax * pcur[g];
    }
  }
}

/* ------------------------------------------------------------------------------------ */
/* PP_rhs (src/solver_kernel.cu:89-176): rhs = -(((uE-uW) idx + (vN-vS) idy) + (wT-wB) idz) rho/dt
 * on interior cells, ghosts of rhs zero (the cudaMemset of cuda_solver.cu:122 is done by the
 * host before this kernel). u*: Gfx (j fastest), v*: Gfy (k fastest), w*: Gfz (i fastest). */
struct FaceStrides { int us1b, us2b, vs1b, vs2b, ws1b, ws2b, cs1b, cs2b; };

template <int NT>
__global__ void __launch_bounds__(NT) k_rhs(int in, int jn, int kn, FaceStrides st, const double *__restrict__ u,
                                            const double *__restrict__ v, const double *__restrict__ w,
                                            double *__restrict__ rhs, double idx, double idy, double idz, double rho_idt)
{
  const long long nrows = (long long)jn * kn;
  for (long long row = blockIdx.x; row < nrows; row += gridDim.x) {
    const int j = (int)(row % jn) + 1, k = (int)(row / jn) + 1;
    for (int i = threadIdx.x + 1; i <= in; i += NT) {
      const double uW = u[j + (long long)k * st.us1b + (long long)i * st.us2b];
      const double uE = u[j + (long long)k * st.us1b + (long long)(i + 1) * st.us2b];
      const double vS = v[k + (long long)i * st.vs1b + (long long)j * st.vs2b];
      const double vN = v[k + (long long)i * st.vs1b + (long long)(j + 1) * st.vs2b];
      const double wB = w[i + (long long)j * st.ws1b + (long long)k * st.ws2b];
      const double wT = w[i + (long long)j * st.ws1b + (long long)(k + 1) * st.ws2b];
      double t = (uE - uW) * idx;
      t += (vN - vS) * idy;
      t += (wT - wB) * idz;
      t *= rho_idt;
      rhs[i + (long long)j * st.cs1b + (long long)k * st.cs2b] = -t;
    }
  }
}

/* Tiled PP_rhs: u-star is stored j-fastest (Gfx) and v-star k-fastest (Gfy) while rhs is i-fastest, so a kernel that
 * walks rows in i reads u-star and v-star with a stride of a whole plane (6.1 ms at 512^3, DRAM sector efficiency 1/4).
 * Here a CTA owns a 32 x 8 x 8 (i,j,k) tile: u-star is fetched as runs of 8 consecutive j, v-star as runs of 8
 * consecutive k (64-byte pieces), transposed through shared memory, and rhs is written in 256-byte rows.
 * Same expression and association as k_rhs / src/solver_kernel.cu:152-157,173. */
#define RHS_TI 32
#define RHS_TJ 8
#define RHS_TK 8
__global__ void __launch_bounds__(256) k_rhs_tiled(int in, int jn, int kn, FaceStrides st, const double *__restrict__ u,
                                                   const double *__restrict__ v, const double *__restrict__ w,
                                                   double *__restrict__ rhs, double idx, double idy, double idz, double rho_idt)
{
  __shared__ double su[RHS_TI + 1][RHS_TK * RHS_TJ + 1];      /* [i'][k*8 + j], padded: lanes vary i' at compute time */
  __shared__ double sv[RHS_TJ + 1][RHS_TK][RHS_TI + 1];       /* [j'][k][i],    padded: lanes vary k at load, i at compute */
  const int nbi = (in + RHS_TI - 1) / RHS_TI, nbj = (jn + RHS_TJ - 1) / RHS_TJ;
  const int bi = blockIdx.x % nbi, bj = (blockIdx.x / nbi) % nbj, bk = blockIdx.x / (nbi * nbj);
  const int i0 = bi * RHS_TI + 1, j0 = bj * RHS_TJ + 1, k0 = bk * RHS_TK + 1;
  const int t = threadIdx.x;
  /* u*(i', j, k), i' = i0 .. i0+32: runs of 8 j for each (i', k) */
  for (int pr = t / 8; pr < (RHS_TI + 1) * RHS_TK; pr += 32) {
    const int ii = pr / RHS_TK, kk = pr % RHS_TK, jj = t % 8;
    const int i = i0 + ii, j = j0 + jj, k = k0 + kk;
    double val = 0.;
    if (i <= in + 1 && j <= jn && k <= kn) val = u[j + (long long)k * st.us1b + (long long)i * st.us2b];
    su[ii][kk * RHS_TJ + jj] = val;
  }
  /* v*(i, j', k), j' = j0 .. j0+8: runs of 8 k for each (i, j') */
  for (int pr = t / 8; pr < RHS_TI * (RHS_TJ + 1); pr += 32) {
    const int jj = pr / RHS_TI, ii = pr % RHS_TI, kk = t % 8;
    const int i = i0 + ii, j = j0 + jj, k = k0 + kk;
    double val = 0.;
    if (i <= in && j <= jn + 1 && k <= kn) val = v[k + (long long)i * st.vs1b + (long long)j * st.vs2b];
    sv[jj][kk][ii] = val;
  }
  __syncthreads();
  const int ii = t % RHS_TI, i = i0 + ii;
  for (int pr = t / RHS_TI; pr < RHS_TJ * RHS_TK; pr += 8) {
    const int jj = pr % RHS_TJ, kk = pr / RHS_TJ;
    const int j = j0 + jj, k = k0 + kk;
    if (i > in || j > jn || k > kn) continue;
    const double wB = w[i + (long long)j * st.ws1b + (long long)k * st.ws2b];
    const double wT = w[i + (long long)j * st.ws1b + (long long)(k + 1) * st.ws2b];
    double tt = (su[ii + 1][kk * RHS_TJ + jj] - su[ii][kk * RHS_TJ + jj]) * idx;
    tt += (sv[jj + 1][kk][ii] - sv[jj][kk][ii]) * idy;
    tt += (wT - wB) * idz;
    tt *= rho_idt;
    rhs[i + (long long)j * st.cs1b + (long long)k * st.cs2b] = -tt;
  }
}

/* flags (+ phase) -> 1-byte coefficient masks in the private layout, pushed into the
 * neighbours' ghost masks.  Replaces PP_jacobi_init (the diagonal is recomputed from the mask). */
template <int NT>
__global__ void __launch_bounds__(NT) k_masks(const Dev d, FaceStrides st, const int *__restrict__ fu,
                                              const int *__restrict__ fv, const int *__restrict__ fw,
                                              const int *__restrict__ phase)
{
  const Layout L = d.L;
  const long long nrows = (long long)L.jn * L.kn;
  bool pushed = false;
  for (long long row = blockIdx.x; row < nrows; row += gridDim.x) {
    const int j = (int)(row % L.jn) + 1, k = (int)(row / L.jn) + 1;
    for (int i = threadIdx.x + 1; i <= L.in; i += NT) {
      const int w_ = fu[j + (long long)k * st.us1b + (long long)i * st.us2b];
      const int e_ = fu[j + (long long)k * st.us1b + (long long)(i + 1) * st.us2b];
      const int s_ = fv[k + (long long)i * st.vs1b + (long long)j * st.vs2b];
      const int n_ = fv[k + (long long)i * st.vs1b + (long long)(j + 1) * st.vs2b];
      const int b_ = fw[i + (long long)j * st.ws1b + (long long)k * st.ws2b];
      const int t_ = fw[i + (long long)j * st.ws1b + (long long)(k + 1) * st.ws2b];
      unsigned m = (e_ * e_ ? FM_E : 0u) | (w_ * w_ ? FM_W : 0u) | (n_ * n_ ? FM_N : 0u) | (s_ * s_ ? FM_S : 0u) |
                   (t_ * t_ ? FM_T : 0u) | (b_ * b_ ? FM_B : 0u);
      const long long g = pidx(L, i, j, k);
      if (phase) {
        const long long C = i + (long long)j * st.cs1b + (long long)k * st.cs2b;
        const bool solid = phase[C] > -1;
        if (solid) m |= FM_SOLID;
        if (solid || phase[C + 1] > -1 || phase[C - 1] > -1 || phase[C + st.cs1b] > -1 || phase[C - st.cs1b] > -1 ||
            phase[C + st.cs2b] > -1 || phase[C - st.cs2b] > -1) m |= FM_NEAR;
      }
      d.fmask[g] = (u8)m;
      /* neighbours need my boundary masks to form z in their ghost copies of my cells */
#define BB_PUSHM(F, COND, II, JJ, KK) if (COND) { const NbrFace &nf = d.halo.f[F]; if (nf.fmask) { nf.fmask[pidx(nf.L, II, JJ, KK)] = (u8)m; pushed = true; } }
      BB_PUSHM(0, i == L.in, 0, j, k)  BB_PUSHM(1, i == 1, nf.L.in + 1, j, k)
      BB_PUSHM(2, j == L.jn, i, 0, k)  BB_PUSHM(3, j == 1, i, nf.L.jn + 1, k)
      BB_PUSHM(4, k == L.kn, i, j, 0)  BB_PUSHM(5, k == 1, i, j, nf.L.kn + 1)
#undef BB_PUSHM
    }
  }
  double v[1] = { 0. }, tot[1];
  if (grid_reduce<1>(d, v, blockIdx.x, gridDim.x, tot, pushed)) rank_allreduce(d, tot, 0, true);
}

/* ------------------------------------------------------------------------------------ */
/* particle right-hand-side patch, src/cuda_solver.cu:128-148 */
__global__ void k_part_rhs_net(double *rhs, const int *phase, const int *phase_shell, long long n)
{   /* net effect of part_BC_p, src/particle_kernel.cu:1753 */
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < n) rhs[i] = (double)(phase[i] < 0 && phase_shell[i]) * rhs[i];
}

template <int NT>
__global__ void __launch_bounds__(NT) k_coeffs_refine(int in, int jn, int kn, int s1b, int s2b, double *rhs,
                                                      const int *__restrict__ phase, double idx2, double idy2, double idz2)
{   /* coeffs_refine, src/solver_kernel.cu:212-256 (in place; only solid neighbours contribute
       and solid cells are never modified, so the in-place update is order independent) */
  const long long nrows = (long long)jn * kn;
  for (long long row = blockIdx.x; row < nrows; row += gridDim.x) {
    const int j = (int)(row % jn) + 1, k = (int)(row / jn) + 1;
    for (int i = threadIdx.x + 1; i <= in; i += NT) {
      const long long CC = i + (long long)j * s1b + (long long)k * s2b;
      const int is_fluid = (phase[CC] == -1);
      double v = rhs[CC];
      v += is_fluid * (phase[CC + 1] > -1) * idx2 * (-rhs[CC + 1]);
      v += is_fluid * (phase[CC - 1] > -1) * idx2 * (-rhs[CC - 1]);
      v += is_fluid * (phase[CC + s1b] > -1) * idy2 * (-rhs[CC + s1b]);
      v += is_fluid * (phase[CC - s1b] > -1) * idy2 * (-rhs[CC - s1b]);
      v += is_fluid * (phase[CC + s2b] > -1) * idz2 * (-rhs[CC + s2b]);
      v += is_fluid * (phase[CC - s2b] > -1) * idz2 * (-rhs[CC - s2b]);
      rhs[CC] = v;
    }
  }
}

/* zero_rhs_ghost_{i,j,k}, src/solver_kernel.cu:178-209 */
__global__ void k_zero_ghosts(double *a, int inb, int jnb, int knb)
{
  const long long n = (long long)inb * jnb * knb;
  for (long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x; c < n; c += (long long)gridDim.x * blockDim.x) {
    int i = (int)(c % inb), j = (int)((c / inb) % jnb), k = (int)(c / ((long long)inb * jnb));
    if (i == 0 || i == inb - 1 || j == 0 || j == jnb - 1 || k == 0 || k == knb - 1) a[c] = 0.;
  }
}

/* ------------------------------------------------------------------------------------ */
/* generic halo exchange on a caller-owned s3b array of ANY of the four grids: mpi_cuda_exchange_Gcc / _Gfx / _Gfy /
 * _Gfz, src/mpi_comm.c:257-405.  k_xchg_send reads my interior faces (what pack_planes_G??_* read,
 * src/bluebottle_kernel.cu:684-1081) and stores them directly into the neighbour's staging buffer for the
 * opposite face (that is the MPI_Put, mpi_comm.c:293-306), same buffer layouts E/W pp=(j-1)+jn(k-1),
 * N/S pp=(k-1)+kn(i-1), T/B pp=(i-1)+in(j-1) with the GRID's extents; the last CTA runs the rank barrier (the
 * MPI_Win_fence); k_xchg_recv is unpack_planes_G??_* (:1083-1466).
 * A face grid is one entry longer along its own normal and shares the block-boundary face with the neighbour
 * (src/domain.c:1292-1301), so along that axis the planes sent are _ie-1 / _is+1 instead of _ie / _is
 * (pack_planes_Gfx_east/west :782-815, Gfy_north/south :916-948, Gfz_top/bottom :1048-1081). */
struct XchgGrid {
  int n[3];            /* interior extents in, jn, kn of the grid                      */
  long long st[3];     /* array strides of i, j, k (elements): the grid's index macro  */
  int send_hi[3];      /* plane sent to the E / N / T neighbour (its ghost 0)          */
  int send_lo[3];      /* plane sent to the W / S / B neighbour (its ghost n+1)        */
};

__device__ __forceinline__ void xchg_decode(const XchgGrid &g, long long t, int &f, long long &pp, int &i, int &j, int &k)
{
  const long long fi = (long long)g.n[1] * g.n[2], fj = (long long)g.n[0] * g.n[2];
  long long u = t;
  if (u < 2 * fi) { f = (int)(u / fi); pp = u % fi; j = (int)(pp % g.n[1]) + 1; k = (int)(pp / g.n[1]) + 1; i = 0; }
  else if ((u -= 2 * fi) < 2 * fj) { f = 2 + (int)(u / fj); pp = u % fj; k = (int)(pp % g.n[2]) + 1; i = (int)(pp / g.n[2]) + 1; j = 0; }
  else { u -= 2 * fj; const long long fk = (long long)g.n[0] * g.n[1]; f = 4 + (int)(u / fk); pp = u % fk; i = (int)(pp % g.n[0]) + 1; j = (int)(pp / g.n[0]) + 1; k = 0; }
}

__global__ void k_xchg_send(const Dev d, const XchgGrid g, const double *__restrict__ a, int buf)
{
  const long long total = 2 * ((long long)g.n[1] * g.n[2] + (long long)g.n[0] * g.n[2] + (long long)g.n[0] * g.n[1]);
  bool pushed = false;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    int f, i, j, k; long long pp;
    xchg_decode(g, t, f, pp, i, j, k);
    const NbrFace &nf = d.halo.f[f];
    if (!nf.recv[buf]) continue;
    if (f < 2) i = (f == 0) ? g.send_hi[0] : g.send_lo[0];
    else if (f < 4) j = (f == 2) ? g.send_hi[1] : g.send_lo[1];
    else k = (f == 4) ? g.send_hi[2] : g.send_lo[2];
    nf.recv[buf][pp] = a[i * g.st[0] + j * g.st[1] + k * g.st[2]];
    pushed = true;
  }
  double v[1] = { 0. }, tot[1];
  if (grid_reduce<1>(d, v, blockIdx.x, gridDim.x, tot, pushed)) rank_allreduce(d, tot, 0, true);
}

__global__ void k_xchg_recv(const Dev d, const XchgGrid g, double *__restrict__ a, int buf)
{
  const long long total = 2 * ((long long)g.n[1] * g.n[2] + (long long)g.n[0] * g.n[2] + (long long)g.n[0] * g.n[1]);
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    int f, i, j, k; long long pp;
    xchg_decode(g, t, f, pp, i, j, k);
    if (!d.halo.f[f].recv[buf]) continue;           /* no neighbour on this side: ghost untouched */
    if (f < 2) i = (f == 0) ? g.n[0] + 1 : 0;       /* _ieb / _isb */
    else if (f < 4) j = (f == 2) ? g.n[1] + 1 : 0;
    else k = (f == 4) ? g.n[2] + 1 : 0;
    a[i * g.st[0] + j * g.st[1] + k * g.st[2]] = d.recv[buf][f][pp];
  }
}

/* ------------------------------------------------------------------------------------ */
/* unit entry: Ap (s3) = -A src (s3b, ghosts as given): same operator device functions */
template <int NT, bool PARTS>
__global__ void __launch_bounds__(NT) k_spmv_s3b(const Dev d, const double *__restrict__ src, int s1b, int s2b,
                                                 double *__restrict__ Ap, int s1, int s2)
{
  const Layout L = d.L;
  const long long nrows = (long long)L.jn * L.kn;
  for (long long row = blockIdx.x; row < nrows; row += gridDim.x) {
    const int j = (int)(row % L.jn) + 1, k = (int)(row / L.jn) + 1;
    for (int i = threadIdx.x + 1; i <= L.in; i += NT) {
      const long long C = i + (long long)j * s1b + (long long)k * s2b;
      const long long g = pidx(L, i, j, k);
      const unsigned m = d.fmask[g];
      double v;
      if (PARTS) v = stencil_parts(d, m, pm_of_neighbours(d.fmask, g, L), src[C], src[C + 1], src[C - 1], src[C + s1b], src[C - s1b], src[C + s2b], src[C - s2b]);
      else v = stencil_noparts(d, m, src[C], src[C + 1], src[C - 1], src[C + s1b], src[C - s1b], src[C + s2b], src[C - s2b]);
      Ap[(i - 1) + (long long)(j - 1) * s1 + (long long)(k - 1) * s2] = v;
    }
  }
}

#endif
