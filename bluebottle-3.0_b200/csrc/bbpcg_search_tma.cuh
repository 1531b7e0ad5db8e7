/* bbpcg_search_tma.cuh -- k_search_tma: the dominant kernel of the PCG iteration, TMA-fed.
 *
 *   p = z + beta p   (z = r * invM[mask])            PP_update_search        src/solver_kernel.cu:906-927
 *   x += alpha_prev p_prev  (lazy phi update)        PP_update_soln_resid    src/solver_kernel.cu:852
 *   q = -A p  (7-point, flag^2 / phase coefficients) PP_spmv_shared_load(_noparts) src/solver_kernel.cu:528-836
 *   (p,q) partial -> last CTA: rank-ordered all-reduce, alpha     src/cuda_solver.cu:204-206
 *
 * CTA = 256 consumer threads + one producer warp (BB_NT_ITER = 288), tile TX = 128 x ty owned cells of one k-plane (ty <= 8 is a RUN-TIME argument: the host planner
 * picks the tile height and the z-chunk count whose CTA count fills the 2 x 148 resident slots, e.g. 7 rows x 4 chunks
 * = 296 CTAs for a 256^3 block, 7 rows x 1 chunk = 296 CTAs at 512^3), marching its z-chunk in k.
 * All plane inputs arrive through TMA (cp.async.bulk.tensor.3d, SASS UTMALDG) into a shared-memory ring, issued D planes
 * ahead by one thread and awaited on mbarriers, so DRAM latency is covered without any register staging:
 *     P ring (D+2 slots): halo'd p_prev tile (TX+4) x (ty+2); converted IN PLACE to p_new
 *     per stage (D+1) : halo'd r tile, mask tile, owned x tile.  The r values of the block's GHOST cells are already in
 *                       this rank's r array: the neighbours' residual kernels (or this block itself for a periodic
 *                       self-wrap) PUSHED them there with plain stores over NVLink, so this kernel issues no remote load
 *                       (a pulled ghost value cost a 2-3 us NVLink round trip per plane on every boundary tile: 15 of the
 *                       147 us of a 256^3 block in a 2 x 2 x 2 run, profiles/r02q_bench_n8_strong512.json).
 * Each thread owns the same (x,y) cells on every plane, so p(k-1), p(k), p(k+1) of its owned
 * cells stay in REGISTERS; only the N/S/E/W neighbours are read back from the P ring.  One
 * mbarrier wait + one __syncthreads per plane.  Stores (p_new, x) are 128-bit from registers.
 *
 * Two instantiations of the plane loop per kernel: XFULL (every one of the tile's 128 columns is an interior column
 * of the block -- all tiles when `in` is a multiple of 128) runs without per-element predicates, and warps whose cells
 * all carry the all-ones flag mask (no wall, no particle nearby) skip the table look-ups and the flag decode; the
 * general form handles ragged row ends and an x-ghost inside the tile.
 *
 * Algorithmic traffic: r, p_prev, x read; p_new, x written = 40 B per cell (+1 B mask).  q is used for the dot
 * product only and re-applied by k_resid_tma (bbpcg_resid_tma.cuh), never stored.
 */
#ifndef BBPCG_SEARCH_TMA_CUH
#define BBPCG_SEARCH_TMA_CUH

#include <cuda.h>
#include "bbpcg_kernels.cuh"

struct SearchMaps {
  CUtensorMap r, p[2], fm;           /* this block: halo'd f64 tiles, halo'd u8 mask tile */
  CUtensorMap xo, ro;                /* owned (TX x ty) f64 tiles of x and r */
  CUtensorMap xh;                    /* halo'd tile of x (refresh form of k_resid_tma) */
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
}

__device__ __forceinline__ void stg128(double *p, double a, double b)
{ asm volatile("st.global.v2.f64 [%0], {%1,%2};" :: "l"(p), "d"(a), "d"(b) : "memory"); }

/* geometry shared with the host (tensor-map boxes, dynamic shared memory size): sized for the tallest tile, ty = 8 */
#define BB_TYMAX 8
/* CTA of the two iteration kernels: 8 consumer warps (threads 0..255 own the tile's cells) + 1 PRODUCER warp whose lane 0
 * (thread 256) issues every TMA load, D planes ahead.  With the issue on thread 0 the whole CTA waited at the per-plane
 * barrier for warp 0's serial expect_tx + 4..9 UTMALDG instructions; the producer warp issues plane lp + D while the
 * consumers compute plane lp and meets them at the same barrier. */
#define BB_NT_ITER 288
#define BB_PRODUCER 256
template <bool PARTS, int DD = 2>
struct SearchGeom {
  static constexpr int TX = 128, NT = 256, HXP = TX + 4, HYMAX = BB_TYMAX + 2;
  static constexpr int MXP = 160, MX0 = 14;            /* mask tile: row pitch; byte of tile column 0 (TMA box starts must be 16-B aligned) */
  static constexpr int a128(int v) { return (v + 127) / 128 * 128; }
  static constexpr int RT = a128(HXP * HYMAX * 8);     /* halo'd f64 tile */
  static constexpr int MT = a128(MXP * HYMAX);         /* halo'd mask tile */
  static constexpr int XT = TX * BB_TYMAX * 8;         /* owned x tile */
  static constexpr int STAGE = RT + MT + XT;
  static constexpr int D = DD, NRS = DD + 1, NPS = DD + 2;   /* planes in flight, r/mask stages, p-ring slots */
  static constexpr int NO = 2;                         /* owned double2 items per thread: tile rows rg+1, rg+5 */
  static constexpr int OFF_STAGE = NPS * RT;
  static constexpr int OFF_TAB = OFF_STAGE + NRS * STAGE;
  static constexpr int OFF_BAR = OFF_TAB + 128 * 8;
  static constexpr int SMEM = OFF_BAR + 64;
  static_assert(D <= NRS, "the done-drain loop indexes barriers 0..D-1");
};

#define BB_FULLMASK2 0x3f3fu        /* both cells of a double2 have all six flags set and no particle nearby (FM_NEAR, FM_SOLID clear) */

/* Particle factors (PM_* bits, low byte = element 0, high byte = element 1 of a double2 item) from the solid bits of the
 * halo'd mask tile.  mp -> the item's two mask bytes inside the tile, pitch = row pitch, m2 = those two bytes. */
__device__ __forceinline__ unsigned gather_pm_inplane(const unsigned char *mp, int pitch, unsigned m2)
{
  const unsigned wl = *reinterpret_cast<const unsigned short *>(mp - 2);        /* columns -2, -1 */
  const unsigned er = *reinterpret_cast<const unsigned short *>(mp + 2);        /* columns +2, +3 */
  const unsigned n2 = *reinterpret_cast<const unsigned short *>(mp + pitch), s2 = *reinterpret_cast<const unsigned short *>(mp - pitch);
  const unsigned c0 = (m2 >> 7) & 1u, c1 = (m2 >> 15) & 1u;
  const unsigned e0 = ((c1 ? PM_E : 0u) | ((wl >> 15) & 1u ? PM_W : 0u) | ((n2 >> 7) & 1u ? PM_N : 0u) | ((s2 >> 7) & 1u ? PM_S : 0u) | (c0 ? PM_C : 0u));
  const unsigned e1 = (((er >> 7) & 1u ? PM_E : 0u) | (c0 ? PM_W : 0u) | ((n2 >> 15) & 1u ? PM_N : 0u) | ((s2 >> 15) & 1u ? PM_S : 0u) | (c1 ? PM_C : 0u));
  return e0 | (e1 << 8);
}
/* ... and of the same two cells one plane up (mt) and one plane down (mb) */
__device__ __forceinline__ unsigned pm_planes(unsigned mt, unsigned mb)
{
  return ((mt >> 7) & 1u ? PM_T : 0u) | ((mb >> 7) & 1u ? PM_B : 0u) | ((mt >> 15) & 1u ? (PM_T << 8) : 0u) | ((mb >> 15) & 1u ? (PM_B << 8) : 0u);
}

/* ---- the TMA producer: lane 0 of the 9th warp ------------------------------------------------------------------
 * Its cursor (item, plane) runs D planes ahead of the consumers in the FLATTENED sequence of this CTA's items, so the
 * ring never drains at an item boundary.  A newly claimed item is published in the small queue the consumers read when
 * they get there (>= one CTA barrier later). */
struct ProdCursor {
  int item;            /* item being issued (-1: nothing left to issue) */
  int lp;              /* next plane of it to issue */
  int n;               /* sequence number of that item inside this CTA */
  int g;               /* flattened plane number of the next issue: ring slot g % N, mbarrier parity (g / N) & 1 */
  ItemGeom ig;
};

template <bool PARTS, int DD>
__device__ __forceinline__ void search_issue(const Dev &d, const SearchMaps &tm, const SearchArgs &a, unsigned char *smem, const ProdCursor &c, int q)
{
  typedef SearchGeom<PARTS, DD> G;
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
  const int rs = c.g % G::NRS, ps = c.g % G::NPS;
  const unsigned bar = bar0 + 8 * rs;
  const unsigned st = sS + rs * G::STAGE;
  unsigned bytes = 2 * (HXP * hy * 8) + G::MXP * hy;
  const bool owned = pi >= k0 && pi <= k1;
  if (owned) bytes += TX * ty * 8;
  tma::mbar_expect_tx(bar, bytes);
  if (owned) tma::load3d(st + G::RT + G::MT, &tm.xo, BB_XOFF + 1 + bx * TX, j0, pi, bar);
  tma::load3d(sP + ps * G::RT, &tm.p[q & 1], x0, y0, pi, bar);
  tma::load3d(st + G::RT, &tm.fm, x0 - G::MX0, y0, pi, bar);
  tma::load3d(st, &tm.r, x0, y0, pi, bar);              /* ghost rows / columns / planes included: the neighbours' residual kernels pushed them */
}

/* one step of the producer: issue the plane under the cursor, move on; at the end of an item claim the next one and
 * publish it.  ISSUE is the kernel's issue function (search_issue / resid_issue). */
template <typename ISSUE>
__device__ __forceinline__ void producer_step(const Dev &d, const SearchArgs &a, ProdCursor &c, int *queue, int which, ISSUE issue)
{
  if (c.item < 0) return;
  issue(c);
  c.g++;
  if (++c.lp == c.ig.nplanes) {
    c.item = claim_item(d, a, which);
    c.n++;
    queue[c.n % BB_QN] = c.item;
    c.lp = 0;
    if (c.item >= 0) c.ig = decode_item(d, a, c.item);
  }
}

/* The consumers' plane loop of ONE item (every thread of the CTA runs it: the producer warp follows the same control flow,
 * its lane 0 issues the loads).  g: flattened plane counter of the CTA.  XFULL: the tile's 128 columns are all interior
 * columns of the block.  Returns the thread's partial of (p, q) over the item. */
template <bool PARTS, int DD, bool XFULL>
__device__ __forceinline__ double search_item(const Dev &d, const SearchMaps &tm, const SearchArgs &a, unsigned char *smem, const ItemGeom &ig,
                                              int &g, ProdCursor &pc, int *queue, int q, double beta, double ax, double c63)
{
  typedef SearchGeom<PARTS, DD> G;
  constexpr int TX = G::TX, HXP = G::HXP, NO = G::NO;
  const double *tab = reinterpret_cast<const double *>(smem + G::OFF_TAB);
  unsigned long long *bars = reinterpret_cast<unsigned long long *>(smem + G::OFF_BAR);

  const Layout L = d.L;
  const int tid = threadIdx.x;
  const int bx = ig.bx, by = ig.by;
  const int ty = a.ty;
  const int i0 = bx * TX + 1, j0 = by * ty + 1;
  const int tyc = min(ty, L.jn - j0 + 1);               /* owned rows of THIS tile (the last tile of a column may be short) */
  const int k0 = ig.k0, k1 = ig.k1, nplanes = ig.nplanes;
  const int y0 = j0 - 1;

  const unsigned bar0 = tma::smem_u32(bars);

  /* ---- per-thread geometry: two owned-row items (tile rows rg+1, rg+5), one halo-row item for the row groups
   * rg = 0 (tile row 0) and rg = 1 (tile row tyc+1), and at most one single (W / E halo column) ---- */
  const int col2 = tid & 63, rg = tid >> 6;
  const int cA = 2 + 2 * col2;                          /* tile column of element 0 */
  const int iA = i0 + 2 * col2;                         /* its global i */
  /* column validity, general form only */
  const bool e0own = XFULL || iA <= L.in, e1own = XFULL || iA + 1 <= L.in;
  const bool e0ok = XFULL || iA <= L.in + 1, e1ok = XFULL || iA + 1 <= L.in + 1;
  const bool e0gx = !XFULL && iA == L.in + 1, e1gx = !XFULL && iA + 1 == L.in + 1;
  bool act[NO], own[NO];                                /* item exists in the block (E ghost included) / element 0 is an owned cell */
  int rowo[NO];
#pragma unroll
  for (int o = 0; o < NO; o++) { rowo[o] = rg + 1 + 4 * o; act[o] = rowo[o] <= tyc && e0ok; own[o] = rowo[o] <= tyc && e0own; }
  const int hrow = rg == 0 ? 0 : tyc + 1;
  const bool hvalid = rg < 2 && e0ok;
  const bool hmy = hvalid && (rg == 0 ? by == 0 : y0 + hrow == L.jn + 1);   /* a y-ghost row of the BLOCK: its p is kept current */
  /* singles: threads 0 .. 2*(tyc+2)-1 handle the W (tile column 1) and E (column TX+2) halo columns */
  const int nsr = tyc + 2;
  const bool has_single = tid < 2 * nsr;
  const int s_side = tid >= nsr ? 1 : 0, s_row = tid - s_side * nsr;
  const int s_col = s_side ? TX + 2 : 1;
  const int s_i = i0 + s_col - 2, s_j = y0 + s_row;
  const bool s_ok = has_single && s_i <= L.in + 1;
  const bool s_store = s_ok && (s_i == 0 || s_i == L.in + 1) && s_j >= 1 && s_j <= L.jn;   /* x-ghost p kept current */

  double *__restrict__ pnew = d.P[(q + 1) & 1];
  double *__restrict__ x = d.x;
  const bool consumer = tid < BB_PRODUCER;              /* warp-uniform */

  /* register pipeline of the owned cells: p(kc-1), p(kc), masks of kc */
  double2 pB[NO], pC[NO];
  unsigned mC[NO], pmC[NO], mBm[NO];                    /* masks of plane kc, its in-plane particle factors, masks of plane kc-1 */
#pragma unroll
  for (int o = 0; o < NO; o++) { pB[o] = make_double2(0., 0.); pC[o] = make_double2(0., 0.); mC[o] = 0; pmC[o] = 0; mBm[o] = 0; }
  /* element offsets of this thread's items inside a plane of the P-layout arrays */
  unsigned goff[NO];
#pragma unroll
  for (int o = 0; o < NO; o++) goff[o] = (unsigned)(iA + BB_XOFF) + (unsigned)(y0 + rowo[o]) * (unsigned)L.px;
  const unsigned hgoff = (unsigned)(iA + BB_XOFF) + (unsigned)(y0 + hrow) * (unsigned)L.px;

  double dot = 0.;

  for (int lp = 0; lp < nplanes; lp++, g++) {
    const int pi = k0 - 1 + lp;
    const int rs = g % G::NRS, ps = g % G::NPS;
    if (tid == a.producer) producer_step(d, a, pc, queue, BB_CLAIM_SEARCH, [&](const ProdCursor &c) { search_issue<PARTS, DD>(d, tm, a, smem, c, q); });
    const bool plane_owned = pi >= k0 && pi <= k1;
    const bool plane_ghost = (pi == 0 || pi == L.kn + 1);
    if (consumer) {
    tma::mbar_wait(bar0 + 8 * rs, (g / G::NRS) & 1);
    if (g == 0) BB_STAMP(d, a, 2);

    double *Pt = reinterpret_cast<double *>(smem + ps * G::RT);
    const unsigned char *St = smem + G::OFF_STAGE + rs * G::STAGE;
    const double *Rt = reinterpret_cast<const double *>(St);
    const unsigned char *Mt = St + G::RT;
    const double *Xt = reinterpret_cast<const double *>(St + G::RT + G::MT);
    double *pnew_pl = pnew + (long long)pi * L.ps;
    double *x_pl = x + (long long)pi * L.ps;

    /* ---- phase A: p_new on the halo'd tile of plane pi ---- */
    double2 pT[NO];
    unsigned mT[NO], pmT[NO];
#pragma unroll
    for (int o = 0; o < NO; o++) {
      pT[o] = make_double2(0., 0.); mT[o] = 0; pmT[o] = 0;
      if (!act[o]) continue;                                             /* warp-uniform in the XFULL form */
      const int row = rowo[o];
      const int so = row * HXP + cA;
      const double2 r2 = *reinterpret_cast<const double2 *>(Rt + so);
      const double2 p2 = *reinterpret_cast<const double2 *>(Pt + so);
      unsigned m2 = *reinterpret_cast<const unsigned short *>(Mt + row * G::MXP + G::MX0 + cA);
      if (!XFULL && !e1ok) m2 &= 0x00ffu;                                  /* ragged row end: no such cell -> dead */
      double c0 = c63, c1 = c63;
      if (!(XFULL && __all_sync(0xffffffffu, m2 == BB_FULLMASK2))) { c0 = tab[m2 & 127u]; c1 = tab[(m2 >> 8) & 127u]; }
      double2 pn;
      pn.x = __fma_rn(beta, p2.x, __dmul_rn(r2.x, c0));                    /* PP_update_search, solver_kernel.cu:921 */
      pn.y = __fma_rn(beta, p2.y, __dmul_rn(r2.y, c1));
      *reinterpret_cast<double2 *>(Pt + so) = pn;
      if (plane_owned) {
        const double2 xc = *reinterpret_cast<const double2 *>(Xt + (row - 1) * TX + 2 * col2);
        const double x0n = __fma_rn(ax, p2.x, xc.x), x1n = __fma_rn(ax, p2.y, xc.y);   /* phi += alpha p, :852 */
        if (XFULL || e1own) {
          stg128(pnew_pl + goff[o], pn.x, pn.y);
          stg128(x_pl + goff[o], x0n, x1n);
        } else if (e0own) {
          pnew_pl[goff[o]] = pn.x; x_pl[goff[o]] = x0n;
          if (e1gx) pnew_pl[goff[o] + 1] = pn.y;                           /* the E ghost's p is kept current */
        } else pnew_pl[goff[o]] = pn.x;                                    /* element 0 IS the E ghost */
        if (PARTS && (m2 & ((FM_NEAR << 8) | FM_NEAR))) pmT[o] = gather_pm_inplane(Mt + row * G::MXP + G::MX0 + cA, G::MXP, m2);   /* rare */
      } else if (plane_ghost) {                                            /* z-ghost copy of p kept current */
        if (XFULL || e1own) stg128(pnew_pl + goff[o], pn.x, pn.y);
        else if (e0own) pnew_pl[goff[o]] = pn.x;
      }
      pT[o] = pn; mT[o] = m2;
    }
    if (hvalid) {                                                          /* the two halo rows (row groups 0 and 1) */
      const int so = hrow * HXP + cA;
      const double2 r2 = *reinterpret_cast<const double2 *>(Rt + so);
      const double2 p2 = *reinterpret_cast<const double2 *>(Pt + so);
      unsigned m2 = *reinterpret_cast<const unsigned short *>(Mt + hrow * G::MXP + G::MX0 + cA);
      if (!XFULL && !e1ok) m2 &= 0x00ffu;
      double2 pn;
      pn.x = __fma_rn(beta, p2.x, __dmul_rn(r2.x, tab[m2 & 127u]));
      pn.y = __fma_rn(beta, p2.y, __dmul_rn(r2.y, tab[(m2 >> 8) & 127u]));
      *reinterpret_cast<double2 *>(Pt + so) = pn;
      if (hmy && plane_owned) {                                            /* block's y-ghost row: elements with 1 <= i <= in */
        if (XFULL || e1own) stg128(pnew_pl + hgoff, pn.x, pn.y);
        else if (e0own) pnew_pl[hgoff] = pn.x;
      }
    }
    if (s_ok) {
      const int so = s_row * HXP + s_col;
      const double rv = Rt[so];
      const double pn = __fma_rn(beta, Pt[so], __dmul_rn(rv, tab[Mt[s_row * G::MXP + G::MX0 + s_col] & 127u]));
      Pt[so] = pn;
      if (s_store && plane_owned) pnew_pl[(unsigned)(s_i + BB_XOFF) + (unsigned)s_j * (unsigned)L.px] = pn;
    }

    /* ---- phase B: q = -A p on plane kc = pi-1 (centre plane in the previous P slot) ---- */
    if (pi - 1 >= k0) {
      const double *Pc = reinterpret_cast<const double *>(smem + ((g - 1) % G::NPS) * G::RT);
#pragma unroll
      for (int o = 0; o < NO; o++) {
        if (!own[o]) continue;
        const int so = rowo[o] * HXP + cA;
        const double2 pN = *reinterpret_cast<const double2 *>(Pc + so + HXP);
        const double2 pS = *reinterpret_cast<const double2 *>(Pc + so - HXP);
        const double pW = Pc[so - 1], pE = Pc[so + 2];
        const unsigned m = mC[o];
        double q0, q1;
        bool plain = XFULL && __all_sync(0xffffffffu, m == BB_FULLMASK2);      /* all six flags set, no particle in the 7-point neighbourhood (FM_NEAR clear) */
        if (plain) {
          q0 = stencil_plain(d, pC[o].x, pC[o].y, pW, pN.x, pS.x, pT[o].x, pB[o].x);
          q1 = stencil_plain(d, pC[o].y, pE, pC[o].x, pN.y, pS.y, pT[o].y, pB[o].y);
        } else if (PARTS) {
          const unsigned pm = pmC[o] | pm_planes(mT[o], mBm[o]);
          q0 = stencil_parts(d, m & 255u, pm & 255u, pC[o].x, pC[o].y, pW, pN.x, pS.x, pT[o].x, pB[o].x);
          q1 = stencil_parts(d, m >> 8, pm >> 8, pC[o].y, pE, pC[o].x, pN.y, pS.y, pT[o].y, pB[o].y);
        } else {
          q0 = stencil_noparts(d, m & 255u, pC[o].x, pC[o].y, pW, pN.x, pS.x, pT[o].x, pB[o].x);
          q1 = stencil_noparts(d, m >> 8, pC[o].y, pE, pC[o].x, pN.y, pS.y, pT[o].y, pB[o].y);
        }
        dot = __fma_rn(pC[o].x, q0, dot);
        if (XFULL || e1own) dot = __fma_rn(pC[o].y, q1, dot);              /* odd row end: element 1 is the E ghost */
      }
    }
#pragma unroll
    for (int o = 0; o < NO; o++) { pB[o] = pC[o]; pC[o] = pT[o]; if (PARTS) { mBm[o] = mC[o]; pmC[o] = pmT[o]; } mC[o] = mT[o]; }
    tma::fence_proxy_async();
    }                                 /* consumer */
    __syncthreads();
  }
  return dot;
}

template <bool PARTS, int DD>
__global__ void __launch_bounds__(BB_NT_ITER, 2)
k_search_tma(const __grid_constant__ Dev d, const __grid_constant__ SearchMaps tm, const SearchArgs a)
{
  typedef SearchGeom<PARTS, DD> G;
  extern __shared__ __align__(128) unsigned char smem[];     /* plain pointer arithmetic only: keeps LDS/STS (no generic LD/ST) */
  __shared__ int queue[BB_QN];
  __shared__ double sh_sum[32];
  const int tid = threadIdx.x;
  if (tid == 0) {
    const unsigned bar0 = tma::smem_u32(smem + G::OFF_BAR);
    for (int s = 0; s < G::NRS; s++) tma::mbar_init(bar0 + 8 * s, 1);
    tma::fence_barrier_init();
  }
  if (tid < 128) reinterpret_cast<double *>(smem + G::OFF_TAB)[tid] = __ldg(d.invM_tab + tid);
  /* the first item is the block index (no atomic, the initial wave keeps its neighbourly layout) */
  ProdCursor pc;
  pc.item = (int)blockIdx.x < a.nitems ? (int)blockIdx.x : -1; pc.lp = 0; pc.n = 0; pc.g = 0;
  if (pc.item >= 0) pc.ig = decode_item(d, a, pc.item);
  if (tid == a.producer) queue[0] = pc.item;
  __syncthreads();

  /* ---- everything above is independent of the previous kernels; from here on we read what they wrote ---- */
  BB_STAMP(d, a, 0);
  pdl_wait();
  BB_STAMP(d, a, 1);
  Scal *sc = d.sc;
  const int q = __ldcg(&sc->q);     /* the scalars come from L2: with PDL this CTA may share an SM (and its L1) with the kernel that wrote them */
  int issued = 0;
  if (tid == a.producer) {
#pragma unroll
    for (int l = 0; l < G::D; l++) if (pc.item >= 0) { producer_step(d, a, pc, queue, BB_CLAIM_SEARCH, [&](const ProdCursor &c) { search_issue<PARTS, DD>(d, tm, a, smem, c, q); }); issued++; }
  }
  const int done = __ldcg(&sc->done);
  const double beta = __ldcg(&sc->beta), ax = __ldcg(&sc->alpha_x);
  if (done) {                       /* a finished solve: drain the loads already issued (no item was claimed yet: an item has > D planes), then leave */
    if (tid == a.producer) {
      const unsigned bar0 = tma::smem_u32(smem + G::OFF_BAR);
      for (int l = 0; l < issued; l++) tma::mbar_wait(bar0 + 8 * (l % G::NRS), (l / G::NRS) & 1);
    }
    return;
  }
  const double c63 = __ldg(d.invM_tab + 63);            /* Jacobi diagonal of a cell with all six flags set */

  int g = 0, n = 0;
  int cur = queue[0];
  while (cur >= 0) {                /* uniform across the CTA: every thread reads the same queue entry after a barrier */
    const ItemGeom ig = decode_item(d, a, cur);
    const bool xfull = (ig.bx * G::TX + G::TX) <= d.L.in;
    const double dot = xfull ? search_item<PARTS, DD, true>(d, tm, a, smem, ig, g, pc, queue, q, beta, ax, c63)
                             : search_item<PARTS, DD, false>(d, tm, a, smem, ig, g, pc, queue, q, beta, ax, c63);
    const double part = block_sum<1>(dot, sh_sum);      /* this ITEM's (p,q) partial in the item's own slot: the total does not depend on who computed what */
    if (tid == 0) d.partials[cur] = part;
    n++;
    cur = queue[n % BB_QN];         /* published by the producer at least one barrier ago */
  }
  BB_STAMP(d, a, 3);
  BB_TRACE_AT(d, a, 6, (unsigned long long)bb_smid() | ((unsigned long long)n << 32) | ((unsigned long long)g << 40));
  BB_TRACE_AT(d, a, 7, 1ull | ((unsigned long long)a.launch << 8));

  pdl_launch_dependents();          /* k_resid_tma may be scheduled behind our tail; it blocks in pdl_wait() until alpha is final */
  /* ---- (p,q): item partials in item order, rank all-reduce, alpha (cuda_solver.cu:204-206) ---- */
  double tot[1];
  IterScal isc;
  if (threadIdx.x == 0) isc = load_iter_scal(d);      /* rz, seq: in flight while this CTA waits for its ticket */
  const bool last = items_reduce(d, a.nitems, BB_CLAIM_SEARCH, tot[0], false);
  BB_STAMP(d, a, 4);
  if (last) {
    rank_allreduce(d, tot, 1, false, &isc.seq);       /* this kernel writes nothing a peer reads */
    if (threadIdx.x == 0) {
      d.sc->pAp = tot[0];
      d.sc->alpha = isc.rz / tot[0];
    }
    BB_STAMP(d, a, 5);
  }
}

#endif
