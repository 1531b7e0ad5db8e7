/* bbpcg_resid_tma.cuh -- k_resid_tma: the residual half of the iteration.
 *
 *   q = -A p          re-applied to the p just written by k_search_tma, NOT read back from HBM
 *                     (same device functions with explicit roundings, same operands => bit-identical to the q whose
 *                     dot product gave alpha)          PP_spmv_shared_load(_noparts)   src/solver_kernel.cu:528-836
 *   r -= alpha q      PP_update_soln_resid (r part)                                src/solver_kernel.cu:855
 *   (r, r*invM)       z recomputed from the mask; inner_product + MPI_Allreduce    src/cuda_solver.cu:231-232
 *   last CTA: rank-ordered all-reduce, stop test, beta                             src/cuda_solver.cu:235-267
 *
 * The iteration moves 40 + 24 = 64 B per cell instead of the 72 of a stored-q scheme (the q write and the q read are
 * gone; p is read here instead of q).
 *
 * CTA = 256 consumer threads + one producer warp (BB_NT_ITER), tile 128 x ty owned cells of one k-plane (same run-time ty and z-chunks as the search kernel).
 * The halo'd p tile and the mask tile arrive through TMA (cp.async.bulk.tensor.3d) into a shared-memory ring, D = 2
 * planes ahead; p ghost cells are already current in HBM (the search kernel keeps them so), so this kernel reads NO peer
 * memory.  Each thread owns the same (x,y) cells on every plane: p(k-1), p(k), p(k+1) of its cells stay in registers,
 * N/S/E/W, the mask and r come from the ring (r through TMA as well: a register prefetch ring stalls on the MOVs of
 * still-pending loads); r is written with 128-bit stores from registers.  As in the search kernel the XFULL form of the
 * plane loop carries no per-element predicates and warps of all-ones masks skip the flag decode and the table look-up.
 *
 * Algorithmic traffic: p, r read; r written = 24 B per cell (+1 B mask).
 */
#ifndef BBPCG_RESID_TMA_CUH
#define BBPCG_RESID_TMA_CUH

#include "bbpcg_search_tma.cuh"

template <bool PARTS, int DD = 2>
struct ResidGeom {
  static constexpr int TX = 128, NT = 256, HXP = TX + 4, HYMAX = BB_TYMAX + 2;
  static constexpr int MXP = 160, MX0 = 14;            /* same mask box as the search kernel (shares its tensor map) */
  static constexpr int a128(int v) { return (v + 127) / 128 * 128; }
  static constexpr int RT = a128(HXP * HYMAX * 8);     /* halo'd p tile */
  static constexpr int MT = a128(MXP * HYMAX);
  static constexpr int ROT = TX * BB_TYMAX * 8;        /* owned r tile */
  static constexpr int STAGE = MT + ROT;
  static constexpr int D = DD, NMS = D + 2, NPS = D + 2; /* a stage stays valid for the iteration after its arrival */
  static constexpr int NO = 2;
  static constexpr int OFF_STAGE = NPS * RT;
  static constexpr int OFF_TAB = OFF_STAGE + NMS * STAGE;
  static constexpr int OFF_BAR = OFF_TAB + 128 * 8;
  static constexpr int SMEM = OFF_BAR + 64;
  static_assert(D <= NMS, "the done-drain loop indexes barriers 0..D-1");
};

/* REFRESH = true is the every-50th-iteration true-residual form (src/cuda_solver.cu:209-223, PP_update_residual
 * src/solver_kernel.cu:883-904): the operator is applied to x (whose ghosts k_refresh_x4 just made current)
 * instead of p, and r = b - (-A x) with b read from the caller's right-hand side array. */
template <bool PARTS, int DD, bool REFRESH>
__device__ __forceinline__ void resid_issue(const Dev &d, const SearchMaps &tm, const SearchArgs &a, unsigned char *smem, const ProdCursor &c, int q)
{
  typedef ResidGeom<PARTS, DD> G;
  constexpr int TX = G::TX, HXP = G::HXP;
  const Layout &L = d.L;
  const int ty = a.ty, hy = ty + 2;
  const int bx = c.ig.bx, k0 = c.ig.k0, k1 = c.ig.k1;
  const int j0 = c.ig.by * ty + 1;
  const int x0 = BB_XOFF + 1 + bx * TX - 2;             /* array x index of tile column 0 */
  const int y0 = j0 - 1;
  const unsigned bar0 = tma::smem_u32(smem + G::OFF_BAR);
  const unsigned sP = tma::smem_u32(smem), sS = tma::smem_u32(smem + G::OFF_STAGE);
  const int pi = k0 - 1 + c.lp;
  const int ms = c.g % G::NMS, ps = c.g % G::NPS;
  const unsigned bar = bar0 + 8 * ms;
  const unsigned st = sS + ms * G::STAGE;
  unsigned bytes = HXP * hy * 8 + G::MXP * hy;
  const bool owned = !REFRESH && pi >= k0 && pi <= k1;
  if (owned) bytes += TX * ty * 8;
  tma::mbar_expect_tx(bar, bytes);
  if (owned) tma::load3d(st + G::MT, &tm.ro, BB_XOFF + 1 + bx * TX, j0, pi, bar);
  tma::load3d(sP + ps * G::RT, REFRESH ? &tm.xh : &tm.p[(q + 1) & 1], x0, y0, pi, bar);   /* p of the iteration in flight (ghosts current) / x */
  tma::load3d(st, &tm.fm, x0 - G::MX0, y0, pi, bar);
}

/* two neighbouring ghost values into a neighbour's array: one 128-bit store when both exist (the pair starts on an even
 * array index in every P-layout, whatever the neighbour's pitch) */
__device__ __forceinline__ void push_pair(double *dst, double a, double b, bool both)
{
  if (both) stg128(dst, a, b); else *dst = a;
}

/* the new r of cells on a y / z face of the block into the neighbour's ghost row / plane (push model, see resid_item);
 * faces: bit 0 S, 1 N, 2 B, 3 T -- the faces with a neighbour that THIS PLANE of the item touches (CTA-uniform).  Behind a
 * call on purpose: inlined, the address arithmetic of the four stores was speculated into the plane loop of every item. */
__device__ __noinline__ void push_yz_faces(const Dev &d, int faces, int iA, int j, int kc, double r0, double r1, bool both)
{
  const Layout &L = d.L;
  const NbrFace &nN = d.halo.f[2], &nS = d.halo.f[3], &nT = d.halo.f[4], &nB = d.halo.f[5];
  if ((faces & 1) && j == 1) push_pair(nS.r + (long long)kc * nS.L.ps + (unsigned)(iA + BB_XOFF) + (unsigned)(nS.L.jn + 1) * (unsigned)nS.L.px, r0, r1, both);
  if ((faces & 2) && j == L.jn) push_pair(nN.r + (long long)kc * nN.L.ps + (unsigned)(iA + BB_XOFF), r0, r1, both);
  if ((faces & 4) && kc == 1) push_pair(nB.r + (long long)(nB.L.kn + 1) * nB.L.ps + (unsigned)(iA + BB_XOFF) + (unsigned)j * (unsigned)nB.L.px, r0, r1, both);
  if ((faces & 8) && kc == L.kn) push_pair(nT.r + (unsigned)(iA + BB_XOFF) + (unsigned)j * (unsigned)nT.L.px, r0, r1, both);
}

/* the plane loop of ONE item; see search_item (bbpcg_search_tma.cuh) for the roles of g, pc and queue */
template <bool PARTS, int DD, bool REFRESH, bool XFULL, bool XPUSH>
__device__ __forceinline__ double resid_item(const Dev &d, const SearchMaps &tm, const SearchArgs &a, unsigned char *smem, const ItemGeom &ig,
                                             int &g, ProdCursor &pc, int *queue, int q, double alpha, double c63)
{
  typedef ResidGeom<PARTS, DD> G;
  constexpr int TX = G::TX, HXP = G::HXP, NO = G::NO;
  const double *tab = reinterpret_cast<const double *>(smem + G::OFF_TAB);
  unsigned long long *bars = reinterpret_cast<unsigned long long *>(smem + G::OFF_BAR);

  const Layout L = d.L;
  const int tid = threadIdx.x;
  const int bx = ig.bx, by = ig.by;
  const int ty = a.ty;
  const int i0 = bx * TX + 1, j0 = by * ty + 1;
  const int tyc = min(ty, L.jn - j0 + 1);
  const int k0 = ig.k0, k1 = ig.k1, nplanes = ig.nplanes;
  const int y0 = j0 - 1;
  const unsigned bar0 = tma::smem_u32(bars);

  /* per-thread geometry: two owned double2 items on tile rows rg+1, rg+5 */
  const int col2 = tid & 63, rg = tid >> 6;
  const int cA = 2 + 2 * col2;                          /* tile column of element 0 */
  const int iA = i0 + 2 * col2;                         /* its global i */
  const bool e0own = XFULL || iA <= L.in, e1own = XFULL || iA + 1 <= L.in;
  bool own[NO];
  int rowo[NO];
  unsigned goff[NO];
#pragma unroll
  for (int o = 0; o < NO; o++) {
    rowo[o] = rg + 1 + 4 * o;
    own[o] = rowo[o] <= tyc && e0own;
    goff[o] = (unsigned)(iA + BB_XOFF) + (unsigned)(y0 + rowo[o]) * (unsigned)L.px;
  }
  /* PUSH model of the halo: the new r of a block-boundary cell also goes into the neighbour's ghost slot (plain stores into
   * peer memory over NVLink, or into this block's own ghosts for a periodic self-wrap); the rank barrier at the end of
   * the kernel orders them before the neighbour's next search kernel reads its r tile, ghosts included. */
  const NbrFace &nE = d.halo.f[0], &nW = d.halo.f[1], &nN = d.halo.f[2], &nS = d.halo.f[3], &nT = d.halo.f[4], &nB = d.halo.f[5];
  /* XPUSH: the item owns a column on an x face that has a neighbour (CTA-uniform, chosen by the caller).  The other form
   * carries no x-push code at all: as predicated stores in the one plane loop their address arithmetic cost every warp of
   * every item ~30 instructions per plane; behind a call (divergent: one lane per row) they cost more still -- 520 -> 569 us
   * at 512^3 (profiles/r02ak_sweep_x_pushes_behind_call_rejected.jsonl). */
  const bool px_w = XPUSH && nW.r != nullptr && iA == 1;
  const bool px_e0 = XPUSH && nE.r != nullptr && iA == L.in, px_e1 = XPUSH && nE.r != nullptr && iA + 1 == L.in;
  /* element offsets of the pushed x-face values inside a plane of the NEIGHBOUR's array (its pitch may differ across an x
   * face), per item, so that a push costs one add per plane */
  unsigned oW[NO], oE[NO];
#pragma unroll
  for (int o = 0; o < NO; o++) {
    oW[o] = (unsigned)(nW.L.in + 1 + BB_XOFF) + (unsigned)(y0 + rowo[o]) * (unsigned)nW.L.px;
    oE[o] = (unsigned)BB_XOFF + (unsigned)(y0 + rowo[o]) * (unsigned)nE.L.px;
  }
  /* y / z faces: CTA-uniform per item, rare on large blocks */
  const bool ys_item = nS.r != nullptr && by == 0, yn_item = nN.r != nullptr && y0 + tyc == L.jn;
  const bool zb_item = nB.r != nullptr && k0 == 1, zt_item = nT.r != nullptr && k1 == L.kn;
  const bool yz_item = ys_item | yn_item | zb_item | zt_item;

  double *__restrict__ r = d.r;
  const bool consumer = tid < BB_PRODUCER;              /* warp-uniform; the producer warp only issues loads */

  double2 pB[NO], pC[NO], bn[NO];                       /* bn: refresh form, b of the NEXT plane to be computed */
  unsigned mBm[NO];                                     /* PARTS: the item's mask bytes one plane below the one being computed */
#pragma unroll
  for (int o = 0; o < NO; o++) { pB[o] = make_double2(0., 0.); pC[o] = make_double2(0., 0.); bn[o] = make_double2(0., 0.); mBm[o] = 0; }

  double dot = 0.;
  for (int lp = 0; lp < nplanes; lp++, g++) {
    const int pi = k0 - 1 + lp;
    const int ps = g % G::NPS;
    if (tid == a.producer) producer_step(d, a, pc, queue, BB_CLAIM_RESID, [&](const ProdCursor &c) { resid_issue<PARTS, DD, REFRESH>(d, tm, a, smem, c, q); });
    const bool plane_owned = pi >= k0 && pi <= k1;
    if (consumer) {
    double2 bc[NO];
    if (REFRESH) {                                      /* b(kc) was requested one iteration ago; request b(kc+1) = b(pi) now */
#pragma unroll
      for (int o = 0; o < NO; o++) {
        bc[o] = bn[o];
        if (plane_owned && own[o]) {
          const long long gb = (long long)iA + (long long)(y0 + rowo[o]) * a.s1b + (long long)pi * a.s2b;    /* caller's Gcc s3b index */
          bn[o].x = a.rhs[gb];
          if (e1own) bn[o].y = a.rhs[gb + 1];
        }
      }
    }
    tma::mbar_wait(bar0 + 8 * (g % G::NMS), (g / G::NMS) & 1);
    if (g == 0) BB_STAMP(d, a, 2);

    const double *Pt = reinterpret_cast<const double *>(smem + ps * G::RT);
    double2 pT[NO];
#pragma unroll
    for (int o = 0; o < NO; o++) pT[o] = own[o] ? *reinterpret_cast<const double2 *>(Pt + rowo[o] * HXP + cA) : make_double2(0., 0.);

    /* ---- plane kc = pi-1: q = -A p from registers + the previous ring slot; r -= alpha q; (r, z) ---- */
    const int kc = pi - 1;
    if (PARTS && lp == 1) {                             /* kc = k0 - 1, the halo plane below the first computed one: remember its solid bits */
      const unsigned char *Mb = smem + G::OFF_STAGE + ((g - 1) % G::NMS) * G::STAGE;
#pragma unroll
      for (int o = 0; o < NO; o++) if (own[o]) mBm[o] = *reinterpret_cast<const unsigned short *>(Mb + rowo[o] * G::MXP + G::MX0 + cA);
    }
    if (kc >= k0) {
      const double *Pc = reinterpret_cast<const double *>(smem + ((g - 1) % G::NPS) * G::RT);
      const unsigned char *Mc = smem + G::OFF_STAGE + ((g - 1) % G::NMS) * G::STAGE;        /* mask, pmask, r of plane kc */
      const unsigned char *Mtop = smem + G::OFF_STAGE + (g % G::NMS) * G::STAGE;           /* mask of plane kc + 1 (this step's arrival) */
      const double *Rc = reinterpret_cast<const double *>(Mc + G::MT);
      double *r_pl = r + (long long)kc * L.ps;
      double *nw_pl = nW.r + (long long)kc * nW.L.ps, *ne_pl = nE.r + (long long)kc * nE.L.ps;      /* uniform; only dereferenced when the neighbour exists */
      /* y / z faces that THIS PLANE of the item pushes (CTA-uniform): a y-boundary tile on every plane, a z-boundary chunk on
       * its one boundary plane only -- gated per item, the top and bottom chunks (half the volume of a 128-plane block under
       * the guided plan) paid the call on every plane */
      const int yz_faces = yz_item ? ((ys_item ? 1 : 0) | (yn_item ? 2 : 0) | ((zb_item && kc == 1) ? 4 : 0) | ((zt_item && kc == L.kn) ? 8 : 0)) : 0;
#pragma unroll
      for (int o = 0; o < NO; o++) {
        if (!own[o]) continue;                                           /* warp-uniform in the XFULL form */
        const int so = rowo[o] * HXP + cA;
        const double2 pN = *reinterpret_cast<const double2 *>(Pc + so + HXP);
        const double2 pS = *reinterpret_cast<const double2 *>(Pc + so - HXP);
        const double pW = Pc[so - 1], pE = Pc[so + 2];
        const unsigned m = *reinterpret_cast<const unsigned short *>(Mc + rowo[o] * G::MXP + G::MX0 + cA);
        unsigned pm = 0;
        if (PARTS) {
          if (m & ((FM_NEAR << 8) | FM_NEAR)) {                          /* rare: a particle in the 7-point neighbourhood of one of the two cells */
            const unsigned mt = *reinterpret_cast<const unsigned short *>(Mtop + rowo[o] * G::MXP + G::MX0 + cA);
            pm = gather_pm_inplane(Mc + rowo[o] * G::MXP + G::MX0 + cA, G::MXP, m) | pm_planes(mt, mBm[o]);
          }
          mBm[o] = m;                                                    /* plane kc is the next step's plane below */
        }
        double2 rc;
        if (REFRESH) rc = bc[o];
        else rc = *reinterpret_cast<const double2 *>(Rc + (rowo[o] - 1) * TX + 2 * col2);
        double q0, q1, c0 = c63, c1 = c63;
        const bool plain = XFULL && __all_sync(0xffffffffu, m == BB_FULLMASK2);
        if (plain) {
          q0 = stencil_plain(d, pC[o].x, pC[o].y, pW, pN.x, pS.x, pT[o].x, pB[o].x);
          q1 = stencil_plain(d, pC[o].y, pE, pC[o].x, pN.y, pS.y, pT[o].y, pB[o].y);
        } else {
          if (PARTS) {
            q0 = stencil_parts(d, m & 255u, pm & 255u, pC[o].x, pC[o].y, pW, pN.x, pS.x, pT[o].x, pB[o].x);
            q1 = stencil_parts(d, m >> 8, pm >> 8, pC[o].y, pE, pC[o].x, pN.y, pS.y, pT[o].y, pB[o].y);
          } else {
            q0 = stencil_noparts(d, m & 255u, pC[o].x, pC[o].y, pW, pN.x, pS.x, pT[o].x, pB[o].x);
            q1 = stencil_noparts(d, m >> 8, pC[o].y, pE, pC[o].x, pN.y, pS.y, pT[o].y, pB[o].y);
          }
          c0 = tab[m & 127u]; c1 = tab[(m >> 8) & 127u];
        }
        const double r0 = REFRESH ? __dsub_rn(rc.x, q0) : __fma_rn(-alpha, q0, rc.x);     /* solver_kernel.cu:897 / :855 */
        dot = __fma_rn(r0, __dmul_rn(r0, c0), dot);                                       /* z = r invM, :858 */
        const bool both = XFULL || e1own;
        double r1 = 0.;
        if (both) {
          r1 = REFRESH ? __dsub_rn(rc.y, q1) : __fma_rn(-alpha, q1, rc.y);
          dot = __fma_rn(r1, __dmul_rn(r1, c1), dot);
          stg128(r_pl + goff[o], r0, r1);
          if (XPUSH && px_e1) ne_pl[oE[o]] = r1;
        } else r_pl[goff[o]] = r0;                                                        /* odd row end: element 1 is the E ghost */
        if (XPUSH && px_w) nw_pl[oW[o]] = r0;
        if (XPUSH && px_e0) ne_pl[oE[o]] = r0;
        /* y / z faces: whole rows of the tile, rare on large blocks -- behind a CALL, so that none of the address arithmetic is
         * speculated into the plane loop (inlined, it cost every plane of every item ~80 instructions per warp: +30 % on the
         * kernel, 534 -> 572 us at 512^3 under the power cap; profiles/r02ae_512_ncu_full.md vs r02p_512_ncu_full.md) */
        if (yz_faces) push_yz_faces(d, yz_faces, iA, y0 + rowo[o], kc, r0, r1, both);
      }
    }
#pragma unroll
    for (int o = 0; o < NO; o++) { pB[o] = pC[o]; pC[o] = pT[o]; }
    }                                 /* consumer */
    __syncthreads();                  /* every thread is done with the slots the next issue overwrites */
  }
  return dot;
}

template <bool PARTS, int DD, bool REFRESH>
__global__ void __launch_bounds__(BB_NT_ITER, 2)
k_resid_tma(const __grid_constant__ Dev d, const __grid_constant__ SearchMaps tm, const SearchArgs a)
{
  typedef ResidGeom<PARTS, DD> G;
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ int queue[BB_QN];
  __shared__ double sh_sum[32];
  const int tid = threadIdx.x;
  if (tid == 0) {
    const unsigned bar0 = tma::smem_u32(smem + G::OFF_BAR);
    for (int s = 0; s < G::NMS; s++) tma::mbar_init(bar0 + 8 * s, 1);
    tma::fence_barrier_init();
  }
  if (tid < 128) reinterpret_cast<double *>(smem + G::OFF_TAB)[tid] = __ldg(d.invM_tab + tid);
  ProdCursor pc;
  pc.item = (int)blockIdx.x < a.nitems ? (int)blockIdx.x : -1; pc.lp = 0; pc.n = 0; pc.g = 0;
  if (pc.item >= 0) pc.ig = decode_item(d, a, pc.item);
  if (tid == a.producer) queue[0] = pc.item;
  __syncthreads();

  /* ---- from here on we read what the search kernel wrote ---- */
  BB_STAMP(d, a, 0);
  pdl_wait();
  BB_STAMP(d, a, 1);
  Scal *sc = d.sc;
  const int q = __ldcg(&sc->q);
  int issued = 0;
  if (tid == a.producer) {
#pragma unroll
    for (int l = 0; l < G::D; l++) if (pc.item >= 0) { producer_step(d, a, pc, queue, BB_CLAIM_RESID, [&](const ProdCursor &c) { resid_issue<PARTS, DD, REFRESH>(d, tm, a, smem, c, q); }); issued++; }
  }
  const double alpha = __ldcg(&sc->alpha);
  const int done = __ldcg(&sc->done);
  if (done) {                       /* a finished solve: drain the loads already issued, then leave */
    if (tid == a.producer) {
      const unsigned bar0 = tma::smem_u32(smem + G::OFF_BAR);
      for (int l = 0; l < issued; l++) tma::mbar_wait(bar0 + 8 * (l % G::NMS), (l / G::NMS) & 1);
    }
    return;
  }
  const double c63 = __ldg(d.invM_tab + 63);

  int g = 0, n = 0;
  int cur = queue[0];
  bool peer_push = false;           /* CTA-uniform: one of this CTA's items stored ghost values into a neighbour's array */
  while (cur >= 0) {
    const ItemGeom ig = decode_item(d, a, cur);
    const bool xfull = (ig.bx * G::TX + G::TX) <= d.L.in;
    peer_push |= item_touches_nbr(d, a, ig);
    const bool xpush = (d.halo.f[1].r != nullptr && ig.bx == 0) || (d.halo.f[0].r != nullptr && ig.bx == a.nbx - 1);
    const double dot = xfull ? (xpush ? resid_item<PARTS, DD, REFRESH, true, true>(d, tm, a, smem, ig, g, pc, queue, q, alpha, c63)
                                      : resid_item<PARTS, DD, REFRESH, true, false>(d, tm, a, smem, ig, g, pc, queue, q, alpha, c63))
                             : resid_item<PARTS, DD, REFRESH, false, true>(d, tm, a, smem, ig, g, pc, queue, q, alpha, c63);
    const double part = block_sum<1>(dot, sh_sum);      /* the item's (r,z) partial in the item's own slot */
    if (tid == 0) d.partials[cur] = part;
    n++;
    cur = queue[n % BB_QN];
  }
  BB_STAMP(d, a, 3);
  BB_TRACE_AT(d, a, 6, (unsigned long long)bb_smid() | ((unsigned long long)n << 32) | ((unsigned long long)g << 40));
  BB_TRACE_AT(d, a, 7, 2ull | ((unsigned long long)a.launch << 8));

  pdl_launch_dependents();
  /* ---- (r,z): item partials in item order, rank all-reduce, stop test, beta (cuda_solver.cu:231-267) ---- */
  double tot[1];
  IterScal isc;
  if (threadIdx.x == 0) isc = load_iter_scal(d);      /* in flight while this CTA waits for its ticket */
  /* system-scope release only in the CTAs that stored into PEER memory (boundary items of a decomposed run) */
  const bool last = items_reduce(d, a.nitems, BB_CLAIM_RESID, tot[0], d.comm.nranks > 1 && peer_push);
  BB_STAMP(d, a, 4);
  if (last) {
    rank_allreduce(d, tot, 1, true, &isc.seq);        /* the neighbours read the ghost values pushed here after this barrier */
    if (threadIdx.x == 0) finish_iteration(d, tot[0], REFRESH, isc);
    BB_STAMP(d, a, 5);
  }
}

#endif
