/* bbpcg_epilogue.cuh -- the solve EPILOGUE: what src/bluebottle.c:233-256 runs on phi right
 * after cuda_PP_cg, fused into one pass over the block.
 *
 *   cuda_dom_BC_p(phi)   src/cuda_bluebottle.cu:2536-2589, BC_p_{W,E,S,N,B,T}_N src/bluebottle_kernel.cu:26-102
 *                        -> k_bc_p: Neumann ghost copy on the faces that have no neighbour
 *   cuda_project()       src/cuda_bluebottle.cu:2495-2503, project_{u,v,w} src/bluebottle_kernel.cu:2303-2355
 *   cuda_update_p()      src/cuda_bluebottle.cu:2505-2534, update_p src/bluebottle_kernel.cu:2385-2402 (the Laplacian of
 *                        :2357-2383 is computed by the reference but NOT used, :2396 vs :2399), copy_p_p_noghost +
 *                        thrust::reduce + MPI_Allreduce (mean), forcing_add_c_const(-pmean) :1507-1518
 *                        -> k_epilogue<PROJECT, UPDATE_P> + k_sub_mean
 *
 * The reference walks each face grid with its SLOW index in the inner loop (project_u: threads over (j,k), loop over i),
 * so phi is read with a stride of a row per lane, three times, and update_p adds a cudaMalloc'ed Laplacian pass, a
 * ghost-free copy, a Thrust reduction and a host MPI_Allreduce.  Here a CTA owns a 16^3 tile of cells: phi is staged
 * once in shared memory (18^3 with the halo), and u*, v*, w*, the flags and the outputs are each touched in THEIR
 * fastest index (Gfx: j, Gfy: k, Gfz/Gcc: i) as 128-byte runs.  The mean of p is a deterministic grid + rank reduction
 * whose result stays on the device.
 *
 * Algorithmic traffic per cell: project 8 (phi) + 24 (u*,v*,w*) + 12 (int flags) + 24 (u,v,w) = 68 B;
 * update_p 8 (p0) + 4 (phase) + 8 (p) = 20 B fused (phi is already on chip) + 16 B for the mean subtraction.
 */
#ifndef BBPCG_EPILOGUE_CUH
#define BBPCG_EPILOGUE_CUH

#include "bbpcg_kernels.cuh"

#define EPI_T 16                       /* tile edge (cells) */
#define EPI_H (EPI_T + 2)              /* with the halo */
#define EPI_SJ 19                      /* smem row pitch: odd, so 16 lanes varying j hit 16 different bank pairs */
#define EPI_SK (EPI_H * EPI_SJ + 1)    /* smem plane pitch: odd as well (lanes varying k) */
#define EPI_SMEM (EPI_H * EPI_SK * 8)

struct EpiArgs {
  const double *u_star, *v_star, *w_star;     /* Gfx / Gfy / Gfz s3b */
  const int *flag_u, *flag_v, *flag_w;
  double *u, *v, *w;
  const double *phi;                          /* Gcc s3b, ghost FACES valid */
  const double *p0;  const int *phase;  double *p;
  double ddx, ddy, ddz;                       /* 1/dx ... (cuda_bluebottle.cu:2498-2502) */
  double dt_rho;                              /* dt / rho_f, evaluated first as in project_u (bluebottle_kernel.cu:2316) */
  int nti, ntj, ntk;                          /* tiles per direction */
};

/* cuda_dom_BC_p: ghost = adjacent interior cell on every face in `faces` (bit f: 0 E, 1 W, 2 N, 3 S, 4 T, 5 B), j/k/i
 * ranges 1..n as in BC_p_*_N (faces only, no edges) */
__global__ void k_bc_p(int in, int jn, int kn, int s1b, int s2b, double *__restrict__ a, unsigned faces)
{
  const long long fi = (long long)jn * kn, fj = (long long)in * kn, fk = (long long)in * jn;
  const long long total = 2 * (fi + fj + fk);
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    long long u = t;
    int f; long long pp;
    if (u < 2 * fi) { f = (int)(u / fi); pp = u % fi; }
    else if ((u -= 2 * fi) < 2 * fj) { f = 2 + (int)(u / fj); pp = u % fj; }
    else { u -= 2 * fj; f = 4 + (int)(u / fk); pp = u % fk; }
    if (!((faces >> f) & 1u)) continue;
    int i, j, k, gi, gj, gk;
    if (f < 2) { j = gj = (int)(pp % jn) + 1; k = gk = (int)(pp / jn) + 1; i = (f == 0) ? in : 1; gi = (f == 0) ? in + 1 : 0; }
    else if (f < 4) { i = gi = (int)(pp % in) + 1; k = gk = (int)(pp / in) + 1; j = (f == 2) ? jn : 1; gj = (f == 2) ? jn + 1 : 0; }
    else { i = gi = (int)(pp % in) + 1; j = gj = (int)(pp / in) + 1; k = (f == 4) ? kn : 1; gk = (f == 4) ? kn + 1 : 0; }
    a[gi + (long long)gj * s1b + (long long)gk * s2b] = a[i + (long long)j * s1b + (long long)k * s2b];
  }
}


/* (element indices are ints, as everywhere in the reference: every s3b of grid_info is an int)
 * one thread's line of faces: all loads of a batch are issued before the first store (the outputs never alias the
 * inputs: _u and _u_star are distinct arrays in the reference, src/bluebottle.h:1088,1137) */
#define EPI_NB 9
__device__ __forceinline__ void project_line(const double *__restrict__ star, const int *__restrict__ flag, double *__restrict__ out,
                                             int c0, int cstride, int nf, const double *ph, int pstride,
                                             double dd, double dt_rho)
{
  for (int f0 = 0; f0 < nf; f0 += EPI_NB) {
    asm volatile("" ::: "memory");            /* keep the batches apart: one batch of loads in flight per thread, not all phases' */
    double sv[EPI_NB]; int fv[EPI_NB];
#pragma unroll
    for (int b = 0; b < EPI_NB; b++) {
      if (f0 + b < nf) { const int c = c0 + (f0 + b) * cstride; sv[b] = __ldg(star + c); fv[b] = __ldg(flag + c); }
    }
#pragma unroll
    for (int b = 0; b < EPI_NB; b++) {
      if (f0 + b < nf) {
        const int f = f0 + b;
        const double gradPhi = abs(fv[b]) * dd * (ph[(f + 1) * pstride] - ph[f * pstride]);   /* bluebottle_kernel.cu:2315,2333,2351 */
        out[c0 + f * cstride] = (sv[b] - dt_rho * gradPhi);                          /* :2316,2334,2352 */
      }
    }
  }
}

/* stage the halo'd phi tile: element e of the (wi x wj x wk) box -> sphi[kk][jj][ii].
 * (An explicit 8-deep register batch of these loads, and a constant-divisor specialisation for full tiles, both pushed the
 * kernel over its 80-register budget: the spills showed up as +22 % DRAM writes and +0.5 ms in ncu,
 * profiles/r01f_epilogue_launches.md.  The plain loop it is.) */
__device__ __forceinline__ void load_phi_tile(double *sphi, const double *__restrict__ phi, int i0, int j0, int k0,
                                              int wi, int wj, int wk, int s1b, int s2b, int t)
{
  const int total = wi * wj * wk;
  const double *src = phi + (i0 - 1) + (long long)(j0 - 1) * s1b + (long long)(k0 - 1) * s2b;
  for (int e = t; e < total; e += 256) {
    const int ii = e % wi, r = e / wi, jj = r % wj, kk = r / wj;
    sphi[kk * EPI_SK + jj * EPI_SJ + ii] = src[ii + jj * s1b + (long long)kk * s2b];
  }
}

/* project_u / project_v / project_w / update_p on one 16^3 tile per loop trip.
 * Face ownership: a tile writes the W, S and B faces of its cells; the last tile of a direction also writes the closing
 * face (i = in+1 etc.), so every face of Gf?._is.._ie is written exactly once (project_u loops i = _is.._ie = 1..in+1). */
template <bool PROJECT, bool UPDATE_P>
__global__ void __launch_bounds__(256, 3) k_epilogue(const __grid_constant__ Dev d, const FaceStrides st, const EpiArgs a)
{
  extern __shared__ __align__(16) double sphi[];           /* [kk][jj][ii] at kk*EPI_SK + jj*EPI_SJ + ii: cell (i0-1+ii, ...) */
  const int in = d.L.in, jn = d.L.jn, kn = d.L.kn;
  const int t = threadIdx.x;
  const int lo = t & 15, hi = t >> 4;                      /* fast / slow lane coordinate of every phase */
  const long long ntiles = (long long)a.nti * a.ntj * a.ntk;
  double psum = 0.;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int bi = (int)(tile % a.nti), bj = (int)((tile / a.nti) % a.ntj), bk = (int)(tile / ((long long)a.nti * a.ntj));
    const int i0 = bi * EPI_T + 1, j0 = bj * EPI_T + 1, k0 = bk * EPI_T + 1;
    const int ni = min(EPI_T, in - i0 + 1), nj = min(EPI_T, jn - j0 + 1), nk = min(EPI_T, kn - k0 + 1);
    __syncthreads();                                       /* previous tile's readers are done */
    /* ---- phi tile with its halo: rows of ni+2 contiguous values ---- */
    load_phi_tile(sphi, a.phi, i0, j0, k0, ni + 2, nj + 2, nk + 2, st.cs1b, st.cs2b, t);
    __syncthreads();
    if (PROJECT) {
      /* u (Gfx, j fastest): lanes (j, k), walk the faces i';  v (Gfy, k fastest): lanes (k, i), walk j';
       * w (Gfz, i fastest): lanes (i, j), walk k'.  One copy of the line code (the phase loop is NOT unrolled: three
       * inlined copies made ptxas keep all three batches live, 158 registers). */
#pragma unroll 1
      for (int phase = 0; phase < 3; phase++) {
        const double *star; const int *flag; double *out; const double *ph;
        int c0, cstride, nf, pstride, nlo, nhi; double dd;
        if (phase == 0) {
          star = a.u_star; flag = a.flag_u; out = a.u; dd = a.ddx; nlo = nj; nhi = nk;
          c0 = (j0 + lo) + (k0 + hi) * st.us1b + i0 * st.us2b; cstride = st.us2b; nf = ni + (i0 + ni - 1 == in ? 1 : 0);
          ph = sphi + (hi + 1) * EPI_SK + (lo + 1) * EPI_SJ; pstride = 1;
        } else if (phase == 1) {
          star = a.v_star; flag = a.flag_v; out = a.v; dd = a.ddy; nlo = nk; nhi = ni;
          c0 = (k0 + lo) + (i0 + hi) * st.vs1b + j0 * st.vs2b; cstride = st.vs2b; nf = nj + (j0 + nj - 1 == jn ? 1 : 0);
          ph = sphi + (lo + 1) * EPI_SK + (hi + 1); pstride = EPI_SJ;
        } else {
          star = a.w_star; flag = a.flag_w; out = a.w; dd = a.ddz; nlo = ni; nhi = nj;
          c0 = (i0 + lo) + (j0 + hi) * st.ws1b + k0 * st.ws2b; cstride = st.ws2b; nf = nk + (k0 + nk - 1 == kn ? 1 : 0);
          ph = sphi + (hi + 1) * EPI_SJ + (lo + 1); pstride = EPI_SK;
        }
        if (lo < nlo && hi < nhi) project_line(star, flag, out, c0, cstride, nf, ph, pstride, dd, a.dt_rho);
      }
    }
    if (UPDATE_P) {
      /* ---- p = (phase < 0)(p0 + phi) on the cells of the tile (Gcc, i fastest); partial sum for the mean ---- */
      const int i = i0 + lo, j = j0 + hi;
      if (lo < ni && hi < nj) {
        const int base = i + j * st.cs1b + k0 * st.cs2b;
        const double *ph = sphi + EPI_SK + (hi + 1) * EPI_SJ + (lo + 1);
        const double *__restrict__ p0 = a.p0;
        const int *__restrict__ phase = a.phase;
        double *__restrict__ p = a.p;
        for (int f0 = 0; f0 < nk; f0 += 8) {
          double pv[8]; int fv[8];
#pragma unroll
          for (int b = 0; b < 8; b++) if (f0 + b < nk) { const int c = base + (f0 + b) * st.cs2b; pv[b] = __ldg(p0 + c); fv[b] = __ldg(phase + c); }
#pragma unroll
          for (int b = 0; b < 8; b++) {
            if (f0 + b < nk) {
              const double val = (fv[b] < 0) * (pv[b] + ph[(f0 + b) * EPI_SK]);        /* bluebottle_kernel.cu:2396 */
              p[base + (f0 + b) * st.cs2b] = val;
              psum += val;
            }
          }
        }
      }
    }
  }
  if (UPDATE_P) {
    /* mean pressure: thrust::reduce + MPI_Allreduce of cuda_bluebottle.cu:2526-2529, kept on the device */
    double v[1] = { psum }, tot[1];
    if (grid_reduce<1>(d, v, blockIdx.x, gridDim.x, tot, false)) {
      rank_allreduce(d, tot, 1, false);
      if (threadIdx.x == 0) d.sc->p_sum = tot[0];
    }
  }
}

/* ------------------------------------------------------------------------------------------------------------------
 * The same epilogue as TWO plane-marching streaming kernels (default; option epilogue_tiled 0 selects k_epilogue above).
 * k_epilogue keeps a 16^3 phi brick in shared memory and serves all three face grids from it: every access is a 128-byte
 * run on the reference's ghosted pitches (5 sectors fetched for 4 used), the brick is refilled 12-14 dependent DRAM round
 * trips per tile, and ncu showed 1.29 x the algorithmic bytes at 52 % of peak (profiles/r01h_epilogue_ncu_full.md).
 *
 *   k_epi_uwp  tile 32 (i) x 32 (j), marching k: phi row tile, w*, flag_w, p0, phase with lanes along i; u*, flag_u with lanes
 *              along j; the u lanes read phi(i) - phi(i-1) across the phi tile staged in shared memory; w uses phi(k-1) kept
 *              in registers; p and its partial sum from the same phi.  Every global access is a 256-byte run (128 for ints).
 *   k_epi_v    tile 32 (k) x 32 (i), marching j: phi rows along i for 32 planes, transposed through shared memory to the
 *              k-fastest Gfy layout; the previous j plane stays in shared memory (three phi buffers).
 * Both fetch every stream with plain coalesced loads into registers one plane ahead of its use (see the kernels).
 * phi is read twice (8 B/cell more than the fused brick) but nothing is fetched in 128-byte pieces any more.
 * Same expressions as project_line above, so u, v, w, p are bit-identical to k_epilogue's. */
#define EA_T 32
#define EA_PA 35                       /* pitch of the (j, i) phi tile of k_epi_uwp: odd multiple of the bank pair */
#define EA_PB 33                       /* pitch of the (k, i) phi tile of k_epi_v */

struct EpiPlan {
  int nti, ntj, ntk;                   /* 32-cell tiles per direction */
  int kc, nzc;                         /* k_epi_uwp: planes per z-chunk, chunks */
  int jc, njc;                         /* k_epi_v: faces per j-chunk, chunks */
};

/* Every stream is fetched by plain coalesced loads (ld.global.nc) into REGISTERS one plane ahead of its use:
 * w*, flag_w, p0, phase (lanes along i) and u*, flag_u (lanes along j) are private to the thread that loads them; only phi
 * crosses threads (u reads it with lanes along j) and goes registers -> shared memory, two buffers, ONE barrier per plane.
 * Order inside a plane: phi(k) to shared memory, w and p from the registers that have landed, the loads of plane k + 1 into
 * the same registers, barrier, u, the loads of u*(k + 1): every load has a barrier and half a plane of arithmetic between
 * issue and use without a second register set (109 registers, 2 CTAs per SM).
 * Measured at 512^3 (profiles/r02ab_*, r02ac_*): this form 1.75 ms at 5.35 TB/s of DRAM traffic; the earlier all-cp.async
 * form 1.87 ms (every stream a 4 / 8-byte LDGSTS: MIO-bound, two barriers per plane); a hybrid with u* by cp.async at 80
 * registers / 3 CTAs per SM 2.25 ms (a third CTA per SM widens the working set: L2 hit rate 19 -> 13 %, +1.5 GB of DRAM
 * reads); evict-first loads (ld.global.cs) 2.66 ms: they lose the L2 hits on the 32-byte sectors that neighbouring tiles
 * share on the reference's ghosted pitches (+2.3 GB). */
struct EpiSmemU { double phi[2][EA_T][EA_PA]; };                      /* phi[k & 1][jj][ii]: phi(i0 - 1 + ii, j0 + jj, k) */

template <bool PROJECT, bool UPDATE_P>
__global__ void __launch_bounds__(256, 2) k_epi_uwp(const __grid_constant__ Dev d, const FaceStrides st, const EpiArgs a, const EpiPlan pl)
{
  __shared__ EpiSmemU M;
  const int in = d.L.in, jn = d.L.jn, kn = d.L.kn;
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  const long long ntile = (long long)pl.nti * pl.ntj, nitems = ntile * pl.nzc;
  double psum = 0.;
  for (long long item = blockIdx.x; item < nitems; item += gridDim.x) {
    const int tile = (int)(item % ntile), zc = (int)(item / ntile);
    const int bi = tile % pl.nti, bj = tile / pl.nti;
    const int i0 = bi * EA_T + 1, j0 = bj * EA_T + 1;
    const int ni = min(EA_T, in - i0 + 1), nj = min(EA_T, jn - j0 + 1);
    const int k0 = zc * pl.kc + 1, k1 = min(kn, k0 + pl.kc - 1);
    const bool close_i = i0 + ni - 1 == in;                /* this tile also owns the closing face i = in + 1 (project_u loops _is.._ie) */
    const int nfi = ni + (close_i ? 1 : 0);
    const int klast = k1 + ((PROJECT && k1 == kn) ? 1 : 0);   /* ... and the last chunk the closing face k = kn + 1 of w */
    /* role A: cell (i0 + lane, j0 + wp + 8 r): phi, w*, flag_w, p0, phase -> w, p.  role B: face i0 + wp + 8 r, j0 + lane: u*, flag_u -> u */
    bool okA[4], okB[5];
    long long cA[4], fA[4], uB[5];
#pragma unroll
    for (int r = 0; r < 4; r++) {
      const int jj = wp + 8 * r;
      okA[r] = lane < ni && jj < nj;
      cA[r] = (i0 + lane) + (long long)(j0 + jj) * st.cs1b;             /* + k * cs2b */
      fA[r] = (i0 + lane) + (long long)(j0 + jj) * st.ws1b;             /* + k * ws2b */
    }
#pragma unroll
    for (int r = 0; r < 5; r++) {
      const int f = wp + 8 * r;
      okB[r] = PROJECT && f < nfi && lane < nj;
      uB[r] = (j0 + lane) + (long long)(i0 + f) * st.us2b;              /* + k * us1b */
    }
    /* the two halo columns of a row: lane 0 -> i0 - 1, lane 1 -> i0 + ni (needed only for the closing face); rows jj = wp + 8 r */
    const bool hal = PROJECT && (lane == 0 || (lane == 1 && close_i));
    const int hcol = lane == 0 ? 0 : ni + 1;
    const long long hoff = (lane == 0 ? i0 - 1 : i0 + ni);
    double pf[4], ws[4], p0[4], us[5], phh[4], pprev[4];
    int fw[4], ph[4], fu[5];
    auto loadA = [&](int k) {
      const bool cells = k <= k1;
#pragma unroll
      for (int r = 0; r < 4; r++) {
        pf[r] = okA[r] ? __ldg(a.phi + cA[r] + (long long)k * st.cs2b) : 0.;
        if (PROJECT) { ws[r] = okA[r] ? __ldg(a.w_star + fA[r] + (long long)k * st.ws2b) : 0.; fw[r] = okA[r] ? __ldg(a.flag_w + fA[r] + (long long)k * st.ws2b) : 0; }
        if (UPDATE_P) { p0[r] = (okA[r] && cells) ? __ldg(a.p0 + cA[r] + (long long)k * st.cs2b) : 0.; ph[r] = (okA[r] && cells) ? __ldg(a.phase + cA[r] + (long long)k * st.cs2b) : 0; }
        if (PROJECT) phh[r] = (hal && cells && wp + 8 * r < nj) ? __ldg(a.phi + hoff + (long long)(j0 + wp + 8 * r) * st.cs1b + (long long)k * st.cs2b) : 0.;
      }
    };
    auto loadB = [&](int k) {
#pragma unroll
      for (int r = 0; r < 5; r++) {
        us[r] = okB[r] ? __ldg(a.u_star + uB[r] + (long long)k * st.us1b) : 0.;
        fu[r] = okB[r] ? __ldg(a.flag_u + uB[r] + (long long)k * st.us1b) : 0;
      }
    };
    if (PROJECT) __syncthreads();                          /* the previous item's last u faces are done with the phi buffers */
#pragma unroll
    for (int r = 0; r < 4; r++) pprev[r] = (PROJECT && okA[r]) ? __ldg(a.phi + cA[r] + (long long)(k0 - 1) * st.cs2b) : 0.;
    loadA(k0);
    if (PROJECT) loadB(k0);
    for (int k = k0; k <= klast; k++) {
      const bool cells = k <= k1;                          /* k = kn + 1: only the closing w face */
      double (*P)[EA_PA] = M.phi[k & 1];
      /* ---- phi(k) to shared memory; w and p: lanes along i ---- */
#pragma unroll
      for (int r = 0; r < 4; r++) {
        const int jj = wp + 8 * r;
        if (PROJECT && cells) {
          if (okA[r]) P[jj][lane + 1] = pf[r];
          if (hal && jj < nj) P[jj][hcol] = phh[r];
        }
        if (okA[r]) {
          const double pc = pf[r];
          if (PROJECT) {
            const double gradPhi = abs(fw[r]) * a.ddz * (pc - pprev[r]);                             /* bluebottle_kernel.cu:2351 */
            a.w[fA[r] + (long long)k * st.ws2b] = (ws[r] - a.dt_rho * gradPhi);                      /* :2352 */
          }
          if (UPDATE_P && cells) {
            const double val = (ph[r] < 0) * (p0[r] + pc);                                           /* :2396 */
            a.p[cA[r] + (long long)k * st.cs2b] = val;
            psum += val;
          }
          pprev[r] = pc;
        }
      }
      if (k < klast) loadA(k + 1);                         /* in flight across the barrier and the u faces of plane k */
      if (PROJECT) {
        __syncthreads();                                   /* phi(k) visible; every thread is past the u faces of plane k - 1 (they read the other buffer) */
        if (cells) {
          /* ---- u: lanes along j, phi(i) - phi(i-1) read across the tile ---- */
#pragma unroll
          for (int r = 0; r < 5; r++) {
            if (okB[r]) {
              const int f = wp + 8 * r;
              const double gradPhi = abs(fu[r]) * a.ddx * (P[lane][f + 1] - P[lane][f]);             /* :2315 */
              a.u[uB[r] + (long long)k * st.us1b] = (us[r] - a.dt_rho * gradPhi);                    /* :2316 */
            }
          }
          if (k < k1) loadB(k + 1);
        }
      }
    }
  }
  if (UPDATE_P) {
    /* mean pressure: thrust::reduce + MPI_Allreduce of cuda_bluebottle.cu:2526-2529, kept on the device */
    double v[1] = { psum }, tot[1];
    if (grid_reduce<1>(d, v, blockIdx.x, gridDim.x, tot, false)) {
      rank_allreduce(d, tot, 1, false);
      if (threadIdx.x == 0) d.sc->p_sum = tot[0];
    }
  }
}

/* Every stream is fetched by plain coalesced loads into REGISTERS one face ahead of the computation (v*, flag_v and the v
 * written are private to a thread; only phi crosses threads -- it is loaded with lanes along i and read with lanes along k --
 * and goes registers -> shared memory, three buffers, ONE barrier per face).  0.67 ms at 512^3 = 6.03 TB/s of DRAM traffic
 * (profiles/r02ab_*); the earlier cp.async form issued 12 LDGSTS of 4 / 8 bytes per thread and face and stalled on the MIO
 * queue (mio_throttle 10.3 warps per issue, 1.02 ms, profiles/r02o_epilogue_ncu_full.md). */
struct EpiSmemV { double phi[3][EA_T][EA_PB]; };                      /* phi[j % 3][kk][ii]: phi(i0 + ii, j, k0 + kk) */
#define EPI_SMEM_V ((int)sizeof(EpiSmemV))

__global__ void __launch_bounds__(256, 3) k_epi_v(const __grid_constant__ Dev d, const FaceStrides st, const EpiArgs a, const EpiPlan pl)
{
  __shared__ EpiSmemV M;
  const int in = d.L.in, jn = d.L.jn, kn = d.L.kn;
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  const long long ntile = (long long)pl.ntk * pl.nti, nitems = ntile * pl.njc;
  for (long long item = blockIdx.x; item < nitems; item += gridDim.x) {
    const int tile = (int)(item % ntile), jcn = (int)(item / ntile);
    const int bi = tile % pl.nti, bk = tile / pl.nti;
    const int i0 = bi * EA_T + 1, k0 = bk * EA_T + 1;
    const int ni = min(EA_T, in - i0 + 1), nk = min(EA_T, kn - k0 + 1);
    const int f0 = jcn * pl.jc + 1, f1 = min(jn + 1, f0 + pl.jc - 1);      /* faces j = f0 .. f1 of Gfy._js.._je = 1 .. jn + 1 */
    /* this thread's four phi cells (k row wp + 8 r, i column lane) and four v faces (i row wp + 8 r, k column lane) */
    bool okp[4], okv[4];
    long long cp[4], cv[4];
#pragma unroll
    for (int r = 0; r < 4; r++) {
      const int rr = wp + 8 * r;
      okp[r] = lane < ni && rr < nk;
      okv[r] = rr < ni && lane < nk;
      cp[r] = (i0 + lane) + (long long)(k0 + rr) * st.cs2b;             /* + j * cs1b */
      cv[r] = (k0 + lane) + (long long)(i0 + rr) * st.vs1b;             /* + j * vs2b */
    }
    double pf[4], vs[4];
    int fv[4];
    __syncthreads();                                       /* the previous item's last face is computed: its phi buffers are free */
#pragma unroll
    for (int r = 0; r < 4; r++) if (okp[r]) M.phi[(f0 - 1) % 3][wp + 8 * r][lane] = __ldg(a.phi + cp[r] + (long long)(f0 - 1) * st.cs1b);
#pragma unroll
    for (int r = 0; r < 4; r++) {
      pf[r] = okp[r] ? __ldg(a.phi + cp[r] + (long long)f0 * st.cs1b) : 0.;
      vs[r] = okv[r] ? __ldg(a.v_star + cv[r] + (long long)f0 * st.vs2b) : 0.;
      fv[r] = okv[r] ? __ldg(a.flag_v + cv[r] + (long long)f0 * st.vs2b) : 0;
    }
    for (int j = f0; j <= f1; j++) {
      const int cur = j % 3, prev = (j + 2) % 3;
      double cvs[4];
      int cfv[4];
#pragma unroll
      for (int r = 0; r < 4; r++) {                         /* face j has landed: phi to its buffer, v* / flag_v to the working set */
        if (okp[r]) M.phi[cur][wp + 8 * r][lane] = pf[r];
        cvs[r] = vs[r]; cfv[r] = fv[r];
      }
      if (j < f1) {                                         /* face j + 1: in flight during the barrier and the computation of face j */
#pragma unroll
        for (int r = 0; r < 4; r++) {
          pf[r] = okp[r] ? __ldg(a.phi + cp[r] + (long long)(j + 1) * st.cs1b) : 0.;
          vs[r] = okv[r] ? __ldg(a.v_star + cv[r] + (long long)(j + 1) * st.vs2b) : 0.;
          fv[r] = okv[r] ? __ldg(a.flag_v + cv[r] + (long long)(j + 1) * st.vs2b) : 0;
        }
      }
      /* one barrier per face: phi(j) is visible, and every thread has finished face j - 1 (which read the buffer that
       * face j + 1 will overwrite after the NEXT barrier's predecessor, i.e. not before all threads passed this one) */
      __syncthreads();
#pragma unroll
      for (int r = 0; r < 4; r++) {
        if (okv[r]) {
          const int ii = wp + 8 * r;
          const double gradPhi = abs(cfv[r]) * a.ddy * (M.phi[cur][lane][ii] - M.phi[prev][lane][ii]);   /* bluebottle_kernel.cu:2333 */
          a.v[cv[r] + (long long)j * st.vs2b] = (cvs[r] - a.dt_rho * gradPhi);                              /* :2334 */
        }
      }
    }
  }
}

/* forcing_add_c_const(-pmean, p), src/bluebottle_kernel.cu:1507-1518: interior cells only.
 * pmean = sum / DOM.Gcc.s3 (cuda_bluebottle.cu:2528). */
__global__ void __launch_bounds__(256) k_sub_mean(const __grid_constant__ Dev d, double *__restrict__ p, int s1b, int s2b, double global_cells)
{
  constexpr int R = 4;                     /* rows in flight per thread: all loads of a batch before its stores */
  const int in = d.L.in, jn = d.L.jn;
  const long long nrows = (long long)jn * d.L.kn;
  const double val = -(d.sc->p_sum / global_cells);
  for (long long row0 = (long long)blockIdx.x * R; row0 < nrows; row0 += (long long)gridDim.x * R) {
    long long base[R];
#pragma unroll
    for (int r = 0; r < R; r++) {
      const long long row = row0 + r;
      base[r] = row < nrows ? (long long)((int)(row % jn) + 1) * s1b + (long long)((int)(row / jn) + 1) * s2b : -1;
    }
    for (int i = threadIdx.x + 1; i <= in; i += 256) {
      double v[R];
#pragma unroll
      for (int r = 0; r < R; r++) if (base[r] >= 0) v[r] = p[base[r] + i];
#pragma unroll
      for (int r = 0; r < R; r++) if (base[r] >= 0) p[base[r] + i] = v[r] + val;
    }
  }
}

/* ------------------------------------------------------------------------------------------------------------------
 * cuda_solvability (src/cuda_bluebottle.cu:2313-2492), the solve PROLOGUE's compatibility fix: the net flux of u* through
 * the six faces of the GLOBAL domain (surf_int_{x,y,z}{s,e}, src/bluebottle_kernel.cu:2135-2229, + thrust::reduce, scaled by
 * the face area), summed over ranks (MPI_Allreduce of 3 values), is removed from the outflow plane(s)
 * (plane_eps_*, :2231-2301).  The reference cudaMalloc's a plane, copies, reduces with Thrust six times and all-reduces on
 * the host; here one kernel sums the (contiguous: the normal index is the slowest of a face grid) boundary planes with the
 * deterministic grid + rank reduction and a second one applies the correction from device-resident scalars. */
struct SolvArgs {
  double *u, *v, *w;                   /* u*, v*, w*: Gfx / Gfy / Gfz s3b */
  unsigned planes;                     /* bit 0 xs, 1 xe, 2 ys, 3 ye, 4 zs, 5 ze: this block touches that global face */
  double ayz, azx, axy;                /* dy*dz, dz*dx, dx*dy of this block (cuda_bluebottle.cu:2345 ...) */
  double Ayz, Azx, Axy;                /* DOM.yl*DOM.zl, DOM.zl*DOM.xl, DOM.xl*DOM.yl (:2428 ...) */
  int out_plane;                       /* WEST 0 .. TOP 5, HOMOGENEOUS 10 (src/bluebottle.h:353-425) */
};

/* value (a, b) of a face plane: a = fast, b = middle index of the grid, both 1-based interior */
__device__ __forceinline__ long long plane_index(const FaceStrides &st, int axis, int n_normal, bool end, int a, int b)
{
  const int c = end ? n_normal : 1;    /* _ie = in (= xn+1 faces) / _is = 1 */
  if (axis == 0) return a + (long long)b * st.us1b + (long long)c * st.us2b;     /* Gfx: (j, k), i = c */
  if (axis == 1) return a + (long long)b * st.vs1b + (long long)c * st.vs2b;     /* Gfy: (k, i), j = c */
  return a + (long long)b * st.ws1b + (long long)c * st.ws2b;                    /* Gfz: (i, j), k = c */
}

__global__ void __launch_bounds__(256) k_solv_sum(const __grid_constant__ Dev d, const FaceStrides st, const SolvArgs a)
{
  const int in = d.L.in, jn = d.L.jn, kn = d.L.kn;
  /* plane p: axis p/2, end p&1; extents (fast, middle): x planes (jn, kn), y planes (kn, in), z planes (in, jn) */
  const int na[3] = { jn, kn, in }, nb[3] = { kn, in, jn }, nn[3] = { in + 1, jn + 1, kn + 1 };
  const double *arr[3] = { a.u, a.v, a.w };
  double acc[6] = { 0., 0., 0., 0., 0., 0. };
  const long long gtid = blockIdx.x * (long long)blockDim.x + threadIdx.x, gstride = (long long)gridDim.x * blockDim.x;
#pragma unroll
  for (int p = 0; p < 6; p++) {
    if (!((a.planes >> p) & 1u)) continue;
    const int ax = p >> 1;
    const long long total = (long long)na[ax] * nb[ax];
    for (long long t = gtid; t < total; t += gstride)
      acc[p] += arr[ax][plane_index(st, ax, nn[ax], p & 1, (int)(t % na[ax]) + 1, (int)(t / na[ax]) + 1)];
  }
  /* eps_?e - eps_?s, each scaled by the face area first (cuda_bluebottle.cu:2345,2358,...,2414-2416) */
  double v[3] = { acc[1] * a.ayz - acc[0] * a.ayz, acc[3] * a.azx - acc[2] * a.azx, acc[5] * a.axy - acc[4] * a.axy }, tot[3];
  if (grid_reduce<3>(d, v, blockIdx.x, gridDim.x, tot, false)) {
    rank_allreduce(d, tot, 2, false);                /* the MPI_Allreduce of 3 values (:2420) in two mailbox rounds */
    rank_allreduce(d, tot + 2, 1, false);
    if (threadIdx.x == 0) { d.sc->eps[0] = tot[0]; d.sc->eps[1] = tot[1]; d.sc->eps[2] = tot[2]; }
  }
}

__global__ void __launch_bounds__(256) k_solv_apply(const __grid_constant__ Dev d, const FaceStrides st, const SolvArgs a)
{
  const int in = d.L.in, jn = d.L.jn, kn = d.L.kn;
  const int na[3] = { jn, kn, in }, nb[3] = { kn, in, jn }, nn[3] = { in + 1, jn + 1, kn + 1 };
  double *arr[3] = { a.u, a.v, a.w };
  const double e0 = d.sc->eps[0], e1 = d.sc->eps[1], e2 = d.sc->eps[2];
  const double area[3] = { a.Ayz, a.Azx, a.Axy };
  const long long gtid = blockIdx.x * (long long)blockDim.x + threadIdx.x, gstride = (long long)gridDim.x * blockDim.x;
#pragma unroll
  for (int p = 0; p < 6; p++) {
    if (!((a.planes >> p) & 1u)) continue;
    const int ax = p >> 1;
    double val;
    if (a.out_plane == 10) val = 0.5 * (ax == 0 ? e0 : ax == 1 ? e1 : e2) / area[ax];      /* HOMOGENEOUS, :2470-2472 */
    else if (a.out_plane == p) val = (e0 + e1 + e2) / area[ax];                            /* :2428 ... */
    else continue;
    if (p & 1) val = -val;                                                                 /* start planes + eps, end planes - eps (:2239,:2251) */
    const long long total = (long long)na[ax] * nb[ax];
    for (long long t = gtid; t < total; t += gstride) {
      const long long c = plane_index(st, ax, nn[ax], p & 1, (int)(t % na[ax]) + 1, (int)(t / na[ax]) + 1);
      arr[ax][c] = arr[ax][c] + val;
    }
  }
}

/* ------------------------------------------------------------------------------------------------------------------
 * cuda_dom_BC_star (src/cuda_bluebottle.cu:2111-2311): the velocity boundary-condition table applied to u*, v*, w* on every
 * face of this block that has no neighbour -- 36 kernels BC_{u,v,w}_{W,E,S,N,B,T}_{D,N} (src/bluebottle_kernel.cu:104-598),
 * up to 18 launches per call.  Here ONE launch per axis (W/E, S/N, B/T): a thread owns one boundary line of one component
 * and applies the low face, then the high face -- the reference's order, which matters when a block is 1 cell thick.  The
 * three axis launches stay separate because a later face reads what an earlier one wrote (BC_u_S_D reads u on the plane
 * i = _is that BC_u_W_D just set to the wall value).
 *   normal component, DIRICHLET:   ghost = 2 bc - a(one face in);  wall face = bc                 (:104-132, 300-328, 492-520)
 *   tangential,      DIRICHLET:   ghost = 8/3 bc - 2 a(first) + 1/3 a(second)                    (:134-190, 270-298, ...)
 *   NEUMANN (both):               ghost = a(first)                                               (:192-268, 358-434, 522-598)
 * PERIODIC / PRECURSOR: no action (the switch has no such case). */
struct BcStarArgs {
  double *arr[3];          /* u*, v*, w* (Gfx / Gfy / Gfz s3b) */
  int n[3][3];             /* n[c][axis]: in, jn, kn of component c's grid */
  int st[3][3];            /* st[c][axis]: element stride of i, j, k in that grid (index macros, src/bluebottle.h:70-73) */
  int type[3][2];          /* [c][low / high face of this launch's axis]: 0 = leave alone, BB_DIRICHLET, BB_NEUMANN */
  double val[3][2];
  int axis;                /* 0: W/E   1: S/N   2: B/T */
  int a1[3], a2[3];        /* per component: the two tangential axes, a1 the one with the smaller stride */
};

__global__ void __launch_bounds__(256) k_bc_star(const BcStarArgs a)
{
  long long cnt[3];
#pragma unroll
  for (int c = 0; c < 3; c++) cnt[c] = (a.type[c][0] || a.type[c][1]) ? (long long)a.n[c][a.a1[c]] * a.n[c][a.a2[c]] : 0;
  const long long total = cnt[0] + cnt[1] + cnt[2];
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    long long u = t;
    int c = 0;
    if (u >= cnt[0]) { u -= cnt[0]; c = 1; if (u >= cnt[1]) { u -= cnt[1]; c = 2; } }
    const int n1 = a.n[c][a.a1[c]];
    const long long base = (long long)((int)(u % n1) + 1) * a.st[c][a.a1[c]] + (long long)((int)(u / n1) + 1) * a.st[c][a.a2[c]];
    double *__restrict__ f = a.arr[c] + base;
    const long long sN = a.st[c][a.axis];
    const int N = a.n[c][a.axis];             /* _is = 1, _ie = N, _isb = 0, _ieb = N + 1 along the normal */
    const bool normal = c == a.axis;
    if (a.type[c][0] == BB_DIRICHLET) {
      const double bc = a.val[c][0];
      if (normal) { f[0] = 2. * bc - f[2 * sN]; f[sN] = bc; }
      else f[0] = 8. / 3. * bc - 2. * f[sN] + 1. / 3. * f[2 * sN];
    } else if (a.type[c][0] == BB_NEUMANN) f[0] = f[sN];
    if (a.type[c][1] == BB_DIRICHLET) {
      const double bc = a.val[c][1];
      if (normal) { f[(N + 1) * sN] = 2. * bc - f[(N - 1) * sN]; f[N * sN] = bc; }
      else f[(N + 1) * sN] = 8. / 3. * bc - 2. * f[N * sN] + 1. / 3. * f[(N - 1) * sN];
    } else if (a.type[c][1] == BB_NEUMANN) f[(N + 1) * sN] = f[N * sN];
  }
}

#endif
