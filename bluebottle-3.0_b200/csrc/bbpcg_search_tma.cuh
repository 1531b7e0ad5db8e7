/* bbpcg_search_tma.cuh -- k_search_tma: the dominant kernel of the PCG iteration, TMA-fed.
 *
 *   p = z + beta p   (z = r * invM[mask])            PP_update_search        src/solver_kernel.cu:906-927
 *   x += alpha_prev p_prev  (lazy phi update)        PP_update_soln_resid    src/solver_kernel.cu:852
 *   q = -A p  (7-point, flag^2 / phase coefficients) PP_spmv_shared_load(_noparts) src/solver_kernel.cu:528-836
 *   (p,q) partial -> last CTA: rank-ordered all-reduce, alpha     src/cuda_solver.cu:204-206
 *
 * CTA = 256 threads, tile TX = 128 x TY owned cells of one k-plane, marching KC planes in k.
 * All plane inputs arrive through TMA (cp.async.bulk.tensor.3d, SASS UTMALDG) into a
 * shared-memory ring, issued D (2 or 3) planes ahead by one thread and awaited on mbarriers, so DRAM
 * latency is covered without any register staging:
 *     P ring (D+2 slots): halo'd p_prev tile (TX+4) x (TY+2); converted IN PLACE to p_new
 *     per stage (D+1) : halo'd r tile, mask tile, [pmask tile], and -- PULL model -- the r values of
 *                       ghost cells straight from the NEIGHBOUR's r array (peer memory over NVLink or
 *                       this block itself for a periodic self-wrap): one row per y-ghost, one run of
 *                       HY values from the neighbour's compact x-face buffer per x-ghost, the whole
 *                       tile for a z-ghost plane.
 * Each thread owns the same (x,y) cells on every plane, so p(k-1), p(k), p(k+1) of its owned
 * cells stay in REGISTERS; only the N/S/E/W neighbours are read back from the P ring.  One
 * mbarrier wait + one __syncthreads per plane.  Stores (p_new, x, q) are 128-bit from registers.
 *
 * Algorithmic traffic: r, p_prev, x read; p_new, x, q written = 48 B per cell (+1 B mask);
 * 40 B in the recompute variant (SearchArgs::store_q = 0: q is only used for the dot product here and
 * re-applied by k_resid_tma, bbpcg_resid_tma.cuh).
 */
#ifndef BBPCG_SEARCH_TMA_CUH
#define BBPCG_SEARCH_TMA_CUH

#include <cuda.h>
#include "bbpcg_kernels.cuh"

struct SearchMaps {
  CUtensorMap r, p[2], fm, pm;       /* this block: halo'd f64 tiles, halo'd u8 mask tile, owned u8 pmask tile */
  CUtensorMap xo, ro;                /* owned (TX x TY) f64 tiles of x and r */
  CUtensorMap xh;                    /* halo'd tile of x (refresh form of k_resid_tma) */
  CUtensorMap nb[6];                 /* neighbours' r: E,W = HY-run box on the 2-D compact face buffer, N,S = row box, T,B = tile box */
};

namespace tma {
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count)
{ asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes)
{ asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity)
{
  unsigned ok;
  do {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void load3d(unsigned dst, const CUtensorMap *map, int c0, int c1, int c2, unsigned bar)
{
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
               :: "r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}
__device__ __forceinline__ void load2d(unsigned dst, const CUtensorMap *map, int c0, int c1, unsigned bar)
{
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               :: "r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ void prefetch_map(const CUtensorMap *map) { asm volatile("prefetch.tensormap [%0];" :: "l"(map) : "memory"); }
}

__device__ __forceinline__ void stg128(double *p, double a, double b)
{ asm volatile("st.global.v2.f64 [%0], {%1,%2};" :: "l"(p), "d"(a), "d"(b) : "memory"); }
__device__ __forceinline__ double2 ldg128(const double *p)
{ double2 v; asm volatile("ld.global.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p)); return v; }

/* geometry shared with the host (tensor-map boxes, dynamic shared memory size) */
template <int TY, bool PARTS, int DD = 2>
struct SearchGeom {
  static constexpr int TX = 128, NT = 256, HXP = TX + 4, HY = TY + 2;
  static constexpr int MXP = 160, MX0 = 14;            /* mask tile: row pitch; byte of tile column 0 (TMA box starts must be 16-B aligned) */
  static constexpr int a128(int v) { return (v + 127) / 128 * 128; }
  static constexpr int RT = a128(HXP * HY * 8);        /* halo'd f64 tile */
  static constexpr int MT = a128(MXP * HY);            /* halo'd mask tile */
  static constexpr int GYS = a128(HXP * 8), GY = 2 * GYS;      /* two y-ghost rows */
  static constexpr int GXS = a128(HY * 8), GX = 2 * GXS;       /* two x-ghost columns: HY contiguous values from the neighbour's face buffer */
  static constexpr int PMT = PARTS ? a128(TX * TY) : 0;
  static constexpr int XT = TX * TY * 8;               /* owned x tile */
  static constexpr int STAGE = RT + MT + GY + GX + PMT + XT;
  static constexpr int D = DD, NRS = DD + 1, NPS = DD + 2;   /* planes in flight, r/mask stages, p-ring slots */
  static constexpr int NO = TY / 4;                    /* owned double2 items per thread: rows rg+1+4n */
  static constexpr int NA = NO + 1;                    /* + one halo-row item for the threads rg = 0 (row 0), 1 (row HY-1) */
  static constexpr int OFF_STAGE = NPS * RT;
  static constexpr int OFF_TAB = OFF_STAGE + NRS * STAGE;
  static constexpr int OFF_BAR = OFF_TAB + 128 * 8;
  static constexpr int SMEM = OFF_BAR + 64;
};

/* item flags (per thread, fixed across planes) */
#define SF_ROWOK   0x001u   /* row exists in the block (ghost rows included)                 */
#define SF_OWNROW  0x002u   /* owned row                                                      */
#define SF_E0OWN   0x004u   /* element 0 / 1 is an owned cell (i <= in on an owned row)       */
#define SF_E1OWN   0x008u
#define SF_E0OK    0x010u   /* element exists in the block (i <= in+1)                        */
#define SF_E1OK    0x020u
#define SF_GY      0x040u   /* y-ghost row with a neighbour: r comes from the GY buffer       */
#define SF_GYSIDE  0x080u   /*   0: j = 0 (S neighbour)   1: j = jn+1 (N neighbour)           */
#define SF_MY      0x100u   /* y-ghost row whose p this CTA keeps current (store p_new)       */
#define SF_E0GX    0x200u   /* element is the x-ghost i = in+1 (inside a double2)             */
#define SF_E1GX    0x400u
#define SF_FAST    0x800u   /* owned row, both elements owned, nothing special                */

template <int TY, bool PARTS, int DD>
__global__ void __launch_bounds__(256, 2)
k_search_tma(const __grid_constant__ Dev d, const __grid_constant__ SearchMaps tm, const SearchArgs a)
{
  typedef SearchGeom<TY, PARTS, DD> G;
  constexpr int TX = G::TX, HXP = G::HXP, HY = G::HY, NA = G::NA, NO = G::NO;
  static_assert(TY % 4 == 0 && TY >= 4, "TY must be a multiple of 4");
  extern __shared__ __align__(128) unsigned char smem[];     /* plain pointer arithmetic only: keeps LDS/STS (no generic LD/ST) */
  double *tab = reinterpret_cast<double *>(smem + G::OFF_TAB);
  unsigned long long *bars = reinterpret_cast<unsigned long long *>(smem + G::OFF_BAR);

  const Layout L = d.L;
  Scal *sc = d.sc;
  const int tid = threadIdx.x;
  const int bx = blockIdx.x, by = blockIdx.y;
  const int i0 = bx * TX + 1, j0 = by * TY + 1;
  const int k0 = __ldg(d.ztab + blockIdx.z) + 1;        /* host-written table: safe before pdl_wait() */
  const int k1 = __ldg(d.ztab + blockIdx.z + 1);
  const int nplanes = k1 - k0 + 3;                      /* planes k0-1 .. k1+1 */
  const int x0 = BB_XOFF + 1 + bx * TX - 2;             /* array x index of tile column 0 */
  const int y0 = j0 - 1;

  /* which ghost buffers this tile needs (uniform per CTA) */
  const bool gy0 = (by == 0) && d.halo.f[3].r != nullptr;                    /* row 0 = j 0, from S  */
  const int rowN = L.jn + 1 - y0;                                            /* tile row of j = jn+1 */
  const bool gy1 = (rowN <= HY - 1) && d.halo.f[2].r != nullptr;             /* from N               */
  const bool gx0 = (bx == 0) && d.halo.f[1].r != nullptr;                    /* column i = 0, from W */
  const bool gx1 = (i0 + TX - 1 >= L.in) && d.halo.f[0].r != nullptr;        /* column i = in+1      */

  const unsigned bar0 = tma::smem_u32(bars);
  const unsigned sP = tma::smem_u32(smem), sS = tma::smem_u32(smem + G::OFF_STAGE);

  if (tid == 0) {
    for (int s = 0; s < G::NRS; s++) tma::mbar_init(bar0 + 8 * s, 1);
    tma::fence_barrier_init();
  }
  __syncthreads();

  /* ---- per-thread geometry: owned-row items n < NO on rows rg+1+4n, one halo-row item n = NO for
   * the thread rows rg = 0 (tile row 0) and rg = 1 (tile row HY-1), and at most one single ---- */
  const int col2 = tid & 63, rg = tid >> 6;
  const int cA = 2 + 2 * col2;                          /* tile column of element 0 */
  const int iA = i0 + 2 * col2;                         /* its global i */
  unsigned fl[NA];
  int rowof[NA];
#pragma unroll
  for (int n = 0; n < NA; n++) {
    const int row = n < NO ? rg + 1 + 4 * n : (rg == 0 ? 0 : rg == 1 ? HY - 1 : -1);
    const int j = y0 + row;
    unsigned f = 0;
    if (row >= 0 && j <= L.jn + 1) {
      f |= SF_ROWOK;
      const bool own = row >= 1 && row <= TY && j <= L.jn;
      if (own) f |= SF_OWNROW;
      if (iA <= L.in + 1) f |= SF_E0OK;
      if (iA + 1 <= L.in + 1) f |= SF_E1OK;
      if (own && iA <= L.in) f |= SF_E0OWN;
      if (own && iA + 1 <= L.in) f |= SF_E1OWN;
      if (j == 0 && gy0) f |= SF_GY;
      if (j == L.jn + 1 && gy1) f |= SF_GY | SF_GYSIDE;
      if ((j == 0 && by == 0) || j == L.jn + 1) f |= SF_MY;
      if (iA == L.in + 1) f |= SF_E0GX;
      if (iA + 1 == L.in + 1) f |= SF_E1GX;
      if ((f & (SF_E0OWN | SF_E1OWN)) == (SF_E0OWN | SF_E1OWN)) f |= SF_FAST;
    }
    fl[n] = f; rowof[n] = row;
  }
  /* singles: threads 0 .. 2*HY-1 handle the W (tile column 1) and E (column TX+2) halo columns */
  const bool has_single = tid < 2 * HY;
  const int s_side = tid / HY, s_row = tid % HY;
  const int s_col = s_side ? TX + 2 : 1;
  const int s_i = i0 + s_col - 2, s_j = y0 + s_row;
  const bool s_ok = has_single && s_j <= L.jn + 1 && s_i <= L.in + 1;
  const bool s_gx = s_ok && ((s_i == 0 && gx0) || (s_i == L.in + 1 && gx1));      /* r from the GX buffer */
  const bool s_store = s_ok && (s_i == 0 || s_i == L.in + 1) && s_j >= 1 && s_j <= L.jn;   /* x-ghost p kept current */

  /* ---- everything above is independent of the previous kernels; from here on we read what they wrote ---- */
  pdl_wait();
  const int q = sc->q;
  /* one thread issues every TMA load of plane lp (local index; global plane pi = k0-1+lp) */
  auto issue = [&](int lp) {
    const int pi = k0 - 1 + lp;
    const int rs = lp % G::NRS, ps = lp % G::NPS;
    const unsigned bar = bar0 + 8 * rs;
    const unsigned st = sS + rs * G::STAGE;
    const bool inner = pi >= 1 && pi <= L.kn;
    unsigned bytes = 2 * (HXP * HY * 8) + G::MXP * HY;
    if (inner) {
      if (gy0) bytes += HXP * 8;
      if (gy1) bytes += HXP * 8;
      if (gx0) bytes += HY * 8;
      if (gx1) bytes += HY * 8;
      if (PARTS) bytes += TX * TY;
    }
    const bool owned = pi >= k0 && pi <= k1;
    if (owned) bytes += G::XT;
    tma::mbar_expect_tx(bar, bytes);
    if (owned) tma::load3d(st + G::RT + G::MT + G::GY + G::GX + G::PMT, &tm.xo, BB_XOFF + 1 + bx * TX, j0, pi, bar);
    tma::load3d(sP + ps * G::RT, &tm.p[q & 1], x0, y0, pi, bar);
    tma::load3d(st + G::RT, &tm.fm, x0 - G::MX0, y0, pi, bar);
    if (pi == 0 && d.halo.f[5].r) tma::load3d(st, &tm.nb[5], x0, y0, d.halo.f[5].L.kn, bar);          /* B neighbour's top plane    */
    else if (pi == L.kn + 1 && d.halo.f[4].r) tma::load3d(st, &tm.nb[4], x0, y0, 1, bar);            /* T neighbour's bottom plane */
    else tma::load3d(st, &tm.r, x0, y0, pi, bar);
    if (inner) {
      if (gy0) tma::load3d(st + G::RT + G::MT, &tm.nb[3], x0, d.halo.f[3].L.jn, pi, bar);
      if (gy1) tma::load3d(st + G::RT + G::MT + G::GYS, &tm.nb[2], x0, 1, pi, bar);
      if (gx0) tma::load2d(st + G::RT + G::MT + G::GY, &tm.nb[1], y0, pi, bar);               /* W neighbour's E face, j = y0 .. */
      if (gx1) tma::load2d(st + G::RT + G::MT + G::GY + G::GXS, &tm.nb[0], y0, pi, bar);      /* E neighbour's W face          */
      if (PARTS) tma::load3d(st + G::RT + G::MT + G::GY + G::GX, &tm.pm, BB_XOFF + 1 + bx * TX, j0, pi, bar);
    }
  };
  if (tid == 0) {
#pragma unroll
    for (int l = 0; l < G::D; l++) if (l < nplanes) issue(l);
  }

  const int done = sc->done;
  const double beta = sc->beta, ax = sc->alpha_x;
  double *__restrict__ pnew = d.P[(q + 1) & 1];
  double *__restrict__ x = d.x;
  double *__restrict__ qv = d.q;
  if (tid < 128) tab[tid] = __ldg(d.invM_tab + tid);
  if (done) {                       /* a finished solve: drain the loads already issued, then leave */
    if (tid == 0) {
#pragma unroll
      for (int l = 0; l < G::D; l++) if (l < nplanes) tma::mbar_wait(bar0 + 8 * l, 0);
    }
    return;
  }


  /* register pipeline of the owned cells: p(kc-1), p(kc), masks of kc */
  double2 pB[NO], pC[NO];
  unsigned mC[NO], pmC[NO];
#pragma unroll
  for (int o = 0; o < NO; o++) { pB[o] = make_double2(0., 0.); pC[o] = make_double2(0., 0.); mC[o] = 0; pmC[o] = 0; }
  const long long gown0 = (long long)(iA + BB_XOFF);    /* + j*px + k*ps */

  double dot = 0.;
  __syncthreads();                                      /* table ready */

  for (int lp = 0; lp < nplanes; lp++) {
    const int pi = k0 - 1 + lp;
    const int rs = lp % G::NRS, ps = lp % G::NPS;
    if (tid == 0 && lp + G::D < nplanes) issue(lp + G::D);
    const bool plane_owned = pi >= k0 && pi <= k1;
    const bool plane_ghost = (pi == 0 || pi == L.kn + 1);
    tma::mbar_wait(bar0 + 8 * rs, (lp / G::NRS) & 1);

    double *Pt = reinterpret_cast<double *>(smem + ps * G::RT);
    const unsigned char *St = smem + G::OFF_STAGE + rs * G::STAGE;
    const double *Rt = reinterpret_cast<const double *>(St);
    const unsigned char *Mt = St + G::RT;
    const double *GYt = reinterpret_cast<const double *>(St + G::RT + G::MT);
    const double *GXt = reinterpret_cast<const double *>(St + G::RT + G::MT + G::GY);
    const unsigned char *PMt = St + G::RT + G::MT + G::GY + G::GX;
    const double *Xt = reinterpret_cast<const double *>(St + G::RT + G::MT + G::GY + G::GX + G::PMT);
    double2 xc[NO];                                     /* x of this plane: owned tile, row-major TX wide */
#pragma unroll
    for (int o = 0; o < NO; o++) xc[o] = plane_owned ? *reinterpret_cast<const double2 *>(Xt + (rowof[o] - 1) * TX + 2 * col2) : make_double2(0., 0.);
    const long long gplane = (long long)pi * L.ps;

    /* ---- phase A: p_new on the halo'd tile of plane pi ---- */
    double2 pT[NO];
    unsigned mT[NO], pmT[NO];
#pragma unroll
    for (int n = 0; n < NA; n++) {
      const unsigned f = fl[n];
      if (!(f & SF_ROWOK)) continue;
      const int row = rowof[n];
      const int oi = n < NO ? n : 0;                    /* owned slot (static for n < NO; unused for the halo item) */
      const int so = row * HXP + cA;
      double2 r2 = *reinterpret_cast<const double2 *>(Rt + so);
      const double2 p2 = *reinterpret_cast<const double2 *>(Pt + so);
      unsigned m2 = *reinterpret_cast<const unsigned short *>(Mt + row * G::MXP + G::MX0 + cA);
      if (!(f & SF_FAST)) {
        if (!plane_ghost) {
          if (f & SF_GY) r2 = *reinterpret_cast<const double2 *>(GYt + ((f & SF_GYSIDE) ? G::GYS / 8 : 0) + cA);
          if ((f & SF_E0GX) && gx1) r2.x = GXt[G::GXS / 8 + row];
          if ((f & SF_E1GX) && gx1) r2.y = GXt[G::GXS / 8 + row];
        }
        if (!(f & SF_E0OK)) m2 = (m2 & 0xff00u) | FM_DEAD;
        if (!(f & SF_E1OK)) m2 = (m2 & 0x00ffu) | (FM_DEAD << 8);
      }
      double2 pn;
      pn.x = r2.x * tab[m2 & 127u] + beta * p2.x;                       /* PP_update_search, solver_kernel.cu:921 */
      pn.y = r2.y * tab[(m2 >> 8) & 127u] + beta * p2.y;
      *reinterpret_cast<double2 *>(Pt + so) = pn;
      const long long g = gplane + gown0 + (long long)(y0 + row) * L.px;
      if (f & SF_OWNROW) {
        if (f & SF_FAST) {
          if (plane_owned) {
            stg128(pnew + g, pn.x, pn.y);
            stg128(x + g, xc[oi].x + ax * p2.x, xc[oi].y + ax * p2.y);    /* phi += alpha p, :852 */
          } else if (plane_ghost) stg128(pnew + g, pn.x, pn.y);           /* z-ghost copy of p kept current */
        } else {
          if (plane_owned || plane_ghost) {
            if (f & (SF_E0OWN | (plane_owned ? SF_E0GX : 0u))) pnew[g] = pn.x;
            if (f & (SF_E1OWN | (plane_owned ? SF_E1GX : 0u))) pnew[g + 1] = pn.y;
          }
          if (plane_owned) {
            if (f & SF_E0OWN) x[g] = xc[oi].x + ax * p2.x;
            if (f & SF_E1OWN) x[g + 1] = xc[oi].y + ax * p2.y;
          }
        }
        if (n < NO) {
          pT[oi] = pn; mT[oi] = m2;
          if (PARTS) pmT[oi] = plane_owned ? *reinterpret_cast<const unsigned short *>(PMt + (row - 1) * TX + 2 * col2) : 0u;
        }
      } else if ((f & SF_MY) && plane_owned) {                            /* y-ghost row: elements with 1 <= i <= in */
        if (iA <= L.in) pnew[g] = pn.x;
        if (iA + 1 <= L.in) pnew[g + 1] = pn.y;
      }
    }
    if (s_ok) {
      const int so = s_row * HXP + s_col;
      double rv = Rt[so];
      if (s_gx && !plane_ghost) rv = GXt[(s_i == 0 ? 0 : G::GXS / 8) + s_row];
      const double pn = rv * tab[Mt[s_row * G::MXP + G::MX0 + s_col] & 127u] + beta * Pt[so];
      Pt[so] = pn;
      if (s_store && plane_owned) pnew[gplane + (s_i + BB_XOFF) + (long long)s_j * L.px] = pn;
    }

    /* ---- phase B: q = -A p on plane kc = pi-1 (centre plane in the previous P slot) ---- */
    const int kc = pi - 1;
    if (kc >= k0) {
      const double *Pc = reinterpret_cast<const double *>(smem + ((lp - 1) % G::NPS) * G::RT);
      const long long gpc = (long long)kc * L.ps;
#pragma unroll
      for (int o = 0; o < NO; o++) {
        if (!(fl[o] & (SF_E0OWN | SF_E1OWN))) continue;
        const int so = rowof[o] * HXP + cA;
        const double2 pN = *reinterpret_cast<const double2 *>(Pc + so + HXP);
        const double2 pS = *reinterpret_cast<const double2 *>(Pc + so - HXP);
        const double pW = Pc[so - 1], pE = Pc[so + 2];
        const unsigned m = mC[o];
        double q0, q1;
        if (PARTS) {
          const unsigned pm = pmC[o];
          q0 = stencil_parts(d, m & 255u, pm & 255u, pC[o].x, pC[o].y, pW, pN.x, pS.x, pT[o].x, pB[o].x);
          q1 = stencil_parts(d, m >> 8, pm >> 8, pC[o].y, pE, pC[o].x, pN.y, pS.y, pT[o].y, pB[o].y);
        } else {
          q0 = stencil_noparts(d, m & 255u, pC[o].x, pC[o].y, pW, pN.x, pS.x, pT[o].x, pB[o].x);
          q1 = stencil_noparts(d, m >> 8, pC[o].y, pE, pC[o].x, pN.y, pS.y, pT[o].y, pB[o].y);
        }
        const long long g = gpc + gown0 + (long long)(y0 + rowof[o]) * L.px;
        if (iA + 1 <= L.in) {
          if (a.store_q) stg128(qv + g, q0, q1);
          dot += pC[o].x * q0; dot += pC[o].y * q1;
        } else {                                                          /* odd row end: element 1 is the E ghost */
          if (a.store_q) qv[g] = q0;
          dot += pC[o].x * q0;
        }
      }
    }
#pragma unroll
    for (int o = 0; o < NO; o++) { pB[o] = pC[o]; pC[o] = pT[o]; mC[o] = mT[o]; if (PARTS) pmC[o] = pmT[o]; }
    tma::fence_proxy_async();
    __syncthreads();
  }

  pdl_launch_dependents();          /* k_resid may be scheduled behind our tail; it blocks in pdl_wait() until alpha is final */
  /* ---- (p,q): grid reduction, rank all-reduce, alpha (cuda_solver.cu:204-206) ---- */
  double v[1] = { dot }, tot[1];
  const int bid = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
  const int nblocks = gridDim.x * gridDim.y * gridDim.z;
  if (grid_reduce<1>(d, v, bid, nblocks, tot, false)) {
    rank_allreduce(d, tot, 1, false);         /* this kernel writes nothing a peer reads */
    if (threadIdx.x == 0) {
      sc->pAp = tot[0];
      sc->alpha = sc->rz / tot[0];
    }
  }
}

#endif
