/* bbpcg_resid_tma.cuh -- k_resid_tma: the residual half of the iteration in the RECOMPUTE variant.
 *
 *   q = -A p          re-applied to the p just written by k_search_tma, NOT read back from HBM
 *                     (same device function, same operands => bit-identical to the q whose dot
 *                     product gave alpha)          PP_spmv_shared_load(_noparts)   src/solver_kernel.cu:528-836
 *   r -= alpha q      PP_update_soln_resid (r part)                                src/solver_kernel.cu:855
 *   (r, r*invM)       z recomputed from the mask; inner_product + MPI_Allreduce    src/cuda_solver.cu:231-232
 *   last CTA: rank-ordered all-reduce, stop test, beta                             src/cuda_solver.cu:235-267
 *
 * With it the search kernel no longer stores q: the iteration moves 40 + 24 = 64 B per cell
 * instead of 48 + 24 = 72 (the q write and the q read are gone; p is read here instead of q).
 *
 * CTA = 256 threads, tile 128 x TY owned cells of one k-plane, marching the same z-chunks as the search
 * kernel.  The halo'd p tile and the mask tile arrive through TMA (cp.async.bulk.tensor.3d) into a
 * shared-memory ring, D = 2 planes ahead; p ghost cells are already current in HBM (the search kernel keeps
 * them so), so this kernel reads NO peer memory.  Each thread owns the same (x,y) cells on every plane:
 * p(k-1), p(k), p(k+1) of its cells stay in registers, N/S/E/W, the mask and r come from the ring (r through
 * TMA as well: a register prefetch ring stalls on the MOVs of still-pending loads); r is written with
 * 128-bit stores from registers.
 *
 * Algorithmic traffic: p, r read; r written = 24 B per cell (+1 B mask).
 */
#ifndef BBPCG_RESID_TMA_CUH
#define BBPCG_RESID_TMA_CUH

#include "bbpcg_search_tma.cuh"

template <int TY, bool PARTS, int DD = 2>
struct ResidGeom {
  static constexpr int TX = 128, NT = 256, HXP = TX + 4, HY = TY + 2;
  static constexpr int MXP = 160, MX0 = 14;            /* same mask box as the search kernel (shares its tensor map) */
  static constexpr int a128(int v) { return (v + 127) / 128 * 128; }
  static constexpr int RT = a128(HXP * HY * 8);        /* halo'd p tile */
  static constexpr int MT = a128(MXP * HY);
  static constexpr int PMT = PARTS ? a128(TX * TY) : 0;
  static constexpr int ROT = TX * TY * 8;              /* owned r tile */
  static constexpr int STAGE = MT + PMT + ROT;
  static constexpr int D = DD, NMS = D + 2, NPS = D + 2; /* a stage stays valid for the iteration after its arrival */
  static constexpr int NO = TY / 4;
  static constexpr int OFF_STAGE = NPS * RT;
  static constexpr int OFF_TAB = OFF_STAGE + NMS * STAGE;
  static constexpr int OFF_BAR = OFF_TAB + 128 * 8;
  static constexpr int SMEM = OFF_BAR + 64;
};

/* REFRESH = true is the every-50th-iteration true-residual form (src/cuda_solver.cu:209-223, PP_update_residual
 * src/solver_kernel.cu:883-904): the operator is applied to x (whose ghosts k_refresh_x4 just made current)
 * instead of p, and r = b - (-A x) with b read from the caller's right-hand side array. */
template <int TY, bool PARTS, int MB, int DD, bool REFRESH>
__global__ void __launch_bounds__(256, MB)
k_resid_tma(const __grid_constant__ Dev d, const __grid_constant__ SearchMaps tm, const SearchArgs a)
{
  typedef ResidGeom<TY, PARTS, DD> G;
  constexpr int TX = G::TX, HXP = G::HXP, HY = G::HY, NO = G::NO;
  extern __shared__ __align__(128) unsigned char smem[];
  double *tab = reinterpret_cast<double *>(smem + G::OFF_TAB);
  unsigned long long *bars = reinterpret_cast<unsigned long long *>(smem + G::OFF_BAR);

  const Layout L = d.L;
  Scal *sc = d.sc;
  const int tid = threadIdx.x;
  const int bx = blockIdx.x, by = blockIdx.y;
  const int i0 = bx * TX + 1, j0 = by * TY + 1;
  const int k0 = __ldg(d.ztab + blockIdx.z) + 1;
  const int k1 = __ldg(d.ztab + blockIdx.z + 1);
  const int nplanes = k1 - k0 + 3;                      /* planes k0-1 .. k1+1 */
  const int x0 = BB_XOFF + 1 + bx * TX - 2;             /* array x index of tile column 0 */
  const int y0 = j0 - 1;

  const unsigned bar0 = tma::smem_u32(bars);
  const unsigned sP = tma::smem_u32(smem), sS = tma::smem_u32(smem + G::OFF_STAGE);
  if (tid == 0) {
    for (int s = 0; s < G::NMS; s++) tma::mbar_init(bar0 + 8 * s, 1);
    tma::fence_barrier_init();
  }
  if (tid < 128) tab[tid] = __ldg(d.invM_tab + tid);
  __syncthreads();

  /* per-thread geometry: NO owned double2 items on tile rows rg+1+4o */
  const int col2 = tid & 63, rg = tid >> 6;
  const int cA = 2 + 2 * col2;                          /* tile column of element 0 */
  const int iA = i0 + 2 * col2;                         /* its global i */
  bool e0[NO], e1[NO];
  int rowof[NO];
#pragma unroll
  for (int o = 0; o < NO; o++) {
    const int row = rg + 1 + 4 * o;
    const bool own = (y0 + row) <= L.jn;
    rowof[o] = row;
    e0[o] = own && iA <= L.in;
    e1[o] = own && iA + 1 <= L.in;
  }
  const long long gown0 = (long long)(iA + BB_XOFF);    /* + j*px + k*ps */

  /* ---- from here on we read what the search kernel wrote ---- */
  pdl_wait();
  const int q = sc->q;
  auto issue = [&](int lp) {
    const int pi = k0 - 1 + lp;
    const int ms = lp % G::NMS, ps = lp % G::NPS;
    const unsigned bar = bar0 + 8 * ms;
    const unsigned st = sS + ms * G::STAGE;
    const bool inner = pi >= 1 && pi <= L.kn;
    unsigned bytes = HXP * HY * 8 + G::MXP * HY;
    if (PARTS && inner) bytes += TX * TY;
    const bool owned = !REFRESH && pi >= k0 && pi <= k1;
    if (owned) bytes += G::ROT;
    tma::mbar_expect_tx(bar, bytes);
    if (owned) tma::load3d(st + G::MT + G::PMT, &tm.ro, BB_XOFF + 1 + bx * TX, j0, pi, bar);
    tma::load3d(sP + ps * G::RT, REFRESH ? &tm.xh : &tm.p[(q + 1) & 1], x0, y0, pi, bar);   /* p of the iteration in flight (ghosts current) / x */
    tma::load3d(st, &tm.fm, x0 - G::MX0, y0, pi, bar);
    if (PARTS && inner) tma::load3d(st + G::MT, &tm.pm, BB_XOFF + 1 + bx * TX, j0, pi, bar);
  };
  if (tid == 0) {
#pragma unroll
    for (int l = 0; l < G::D; l++) if (l < nplanes) issue(l);
  }
  const double alpha = sc->alpha;
  const int done = sc->done;
  double *__restrict__ r = d.r;
  if (done) {                       /* a finished solve: drain the loads already issued, then leave */
    if (tid == 0) {
#pragma unroll
      for (int l = 0; l < G::D; l++) if (l < nplanes) tma::mbar_wait(bar0 + 8 * l, 0);
    }
    return;
  }

  double2 pB[NO], pC[NO], bn[NO];                       /* bn: refresh form, b of the NEXT plane to be computed */
#pragma unroll
  for (int o = 0; o < NO; o++) { pB[o] = make_double2(0., 0.); pC[o] = make_double2(0., 0.); bn[o] = make_double2(0., 0.); }

  double dot = 0.;
  for (int lp = 0; lp < nplanes; lp++) {
    const int pi = k0 - 1 + lp;
    const int ms = lp % G::NMS, ps = lp % G::NPS;
    if (tid == 0 && lp + G::D < nplanes) issue(lp + G::D);
    const bool plane_owned = pi >= k0 && pi <= k1;
    double2 bc[NO];
    if (REFRESH) {                                      /* b(kc) was requested one iteration ago; request b(kc+1) = b(pi) now */
#pragma unroll
      for (int o = 0; o < NO; o++) {
        bc[o] = bn[o];
        if (plane_owned && e0[o]) {
          const long long gb = (long long)iA + (long long)(y0 + rowof[o]) * a.s1b + (long long)pi * a.s2b;    /* caller's Gcc s3b index */
          bn[o].x = a.rhs[gb];
          if (e1[o]) bn[o].y = a.rhs[gb + 1];
        }
      }
    }
    tma::mbar_wait(bar0 + 8 * ms, (lp / G::NMS) & 1);

    const double *Pt = reinterpret_cast<const double *>(smem + ps * G::RT);
    double2 pT[NO];
#pragma unroll
    for (int o = 0; o < NO; o++) pT[o] = *reinterpret_cast<const double2 *>(Pt + rowof[o] * HXP + cA);

    /* ---- plane kc = pi-1: q = -A p from registers + the previous ring slot; r -= alpha q; (r, z) ---- */
    const int kc = pi - 1;
    if (kc >= k0) {
      const double *Pc = reinterpret_cast<const double *>(smem + ((lp - 1) % G::NPS) * G::RT);
      const unsigned char *Mc = smem + G::OFF_STAGE + ((lp - 1) % G::NMS) * G::STAGE;       /* mask, pmask, r of plane kc */
      const unsigned char *PMc = Mc + G::MT;
      const double *Rc = reinterpret_cast<const double *>(Mc + G::MT + G::PMT);
      const long long gpc = (long long)kc * L.ps;
#pragma unroll
      for (int o = 0; o < NO; o++) {
        if (!e0[o]) continue;
        const int so = rowof[o] * HXP + cA;
        const double2 pN = *reinterpret_cast<const double2 *>(Pc + so + HXP);
        const double2 pS = *reinterpret_cast<const double2 *>(Pc + so - HXP);
        const double pW = Pc[so - 1], pE = Pc[so + 2];
        const unsigned m = *reinterpret_cast<const unsigned short *>(Mc + rowof[o] * G::MXP + G::MX0 + cA);
        double2 rc;
        if (REFRESH) rc = bc[o];
        else rc = *reinterpret_cast<const double2 *>(Rc + (rowof[o] - 1) * TX + 2 * col2);
        double q0, q1;
        if (PARTS) {
          const unsigned pm = *reinterpret_cast<const unsigned short *>(PMc + (rowof[o] - 1) * TX + 2 * col2);
          q0 = stencil_parts(d, m & 255u, pm & 255u, pC[o].x, pC[o].y, pW, pN.x, pS.x, pT[o].x, pB[o].x);
          q1 = stencil_parts(d, m >> 8, pm >> 8, pC[o].y, pE, pC[o].x, pN.y, pS.y, pT[o].y, pB[o].y);
        } else {
          q0 = stencil_noparts(d, m & 255u, pC[o].x, pC[o].y, pW, pN.x, pS.x, pT[o].x, pB[o].x);
          q1 = stencil_noparts(d, m >> 8, pC[o].y, pE, pC[o].x, pN.y, pS.y, pT[o].y, pB[o].y);
        }
        const long long g = gpc + gown0 + (long long)(y0 + rowof[o]) * L.px;
        double r0 = rc.x, r1 = rc.y;
        if (REFRESH) r0 -= q0; else r0 -= alpha * q0;                   /* solver_kernel.cu:897 / :855 */
        const double z0 = r0 * tab[m & 127u];                           /* :858 */
        if (e1[o]) {
          if (REFRESH) r1 -= q1; else r1 -= alpha * q1;
          const double z1 = r1 * tab[(m >> 8) & 127u];
          stg128(r + g, r0, r1);
          dot += r0 * z0; dot += r1 * z1;
          if (iA + 1 == L.in) store_xface(d, L.in, y0 + rowof[o], kc, r1);
        } else {                                                        /* odd row end: element 1 is the E ghost */
          r[g] = r0;
          dot += r0 * z0;
        }
        if (iA == 1 || iA == L.in) store_xface(d, iA, y0 + rowof[o], kc, r0);
      }
    }
#pragma unroll
    for (int o = 0; o < NO; o++) { pB[o] = pC[o]; pC[o] = pT[o]; }
    __syncthreads();                  /* every thread is done with the slots the next issue overwrites */
  }

  pdl_launch_dependents();
  /* ---- (r,z): grid reduction, rank all-reduce, stop test, beta (cuda_solver.cu:231-267) ---- */
  double v[1] = { dot }, tot[1];
  const int bid = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
  const int nblocks = gridDim.x * gridDim.y * gridDim.z;
  if (grid_reduce<1>(d, v, bid, nblocks, tot, false)) {
    rank_allreduce(d, tot, 1, true);          /* peers pull the r written here */
    if (threadIdx.x == 0) finish_iteration(d, tot[0], REFRESH);
  }
}

#endif
