/* bbpcg_cages.cuh -- the COEFFICIENT PRODUCERS: cuda_build_cages (src/cuda_particle.cu:1516-1646)
 * and the flag -> mask digestion of cuda_PP_init_jacobi_preconditioner fused behind it.
 *
 * Reference: reset_flag_{u,v,w}, reset_phases (src/particle_kernel.cu:79-133); then PER PARTICLE cage_setup<<<1,1>>> + a
 * blocking 12-byte cudaMemcpy + build_phase over the particle's cage box (:135-253), the same again for build_phase_shell
 * (:255-426) -- 4 nparts launches and 2 nparts host round trips per time step (~4000 + ~2000 for the 1000-sphere case);
 * then cage_flag_{u,v,w} (:482-540), up to six flag_external_* launches (:542-576, cuda_particle.cu:1600-1639), and
 * PP_jacobi_init re-reading 12 B/cell of flags (src/solver_kernel.cu:26-87).
 *
 * Here: k_cage_reset, k_cage<false> (phase) and k_cage<true> (phase_shell) with ONE CTA PER PARTICLE -- the cage box is
 * computed on the device by every CTA from the same expressions as cage_setup / build_phase, no host round trip -- and
 * k_cage_flags, one pass over the block that writes the three flag arrays (what the reference's other kernels read) AND the
 * solver's 1-byte masks directly (fmask depends on the external walls only, because a cage flag is +-1 and enters the
 * operator squared; the particle factors come from one more bit, phase > -1), pushes the boundary masks into the
 * neighbours' ghosts and runs the rank barrier: bbpcg_set_coefficients is not needed afterwards.
 *
 * Sequential semantics kept: build_phase applies `phase += cutoff (n - phase)` for n = 0, 1, ... in order, so the LAST
 * particle covering a cell wins = the largest n = atomicMax; build_phase_shell multiplies by 0 or 1 depending only on the
 * FINAL phase and the particle's own geometry = an idempotent store of 0.
 */
#ifndef BBPCG_CAGES_CUH
#define BBPCG_CAGES_CUH

#include "bbpcg_kernels.cuh"

struct CageArgs {
  const char *parts;               /* device: the rank's particle list (the reference's `_parts`), addressed through a strided view */
  unsigned long long stride, ox, oy, oz, orad;
  int nparts;
  double xs, ys, zs, dx, dy, dz;   /* this block: _dom.xs, _dom.dx ... */
  int xn, yn, zn;
  int S[3], E[3];                  /* the range a cage is clipped to per axis: _is.._ie on a non-periodic global edge, _isb.._ieb elsewhere (:177-211) */
  int s1b, s2b;                    /* Gcc strides */
  int *phase, *phase_shell;
};

__device__ __forceinline__ double part_field(const CageArgs &a, int n, unsigned long long off)
{
  return *reinterpret_cast<const double *>(a.parts + (size_t)n * a.stride + off);
}

__global__ void __launch_bounds__(256) k_cage_reset(int *__restrict__ phase, int *__restrict__ phase_shell, long long n)
{   /* reset_phases, src/particle_kernel.cu:121-133 */
  for (long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x; c < n; c += (long long)gridDim.x * blockDim.x) {
    phase[c] = -1; phase_shell[c] = 1;
  }
}

/* one CTA per particle: cage_setup (:135-146) + the box / clipping logic and the cell test of build_phase (:148-253) or
 * build_phase_shell (:255-345) */
template <bool SHELL>
__global__ void __launch_bounds__(256) k_cage(const CageArgs a)
{
  const int n = blockIdx.x;
  if (n >= a.nparts) return;
  const double px = part_field(a, n, a.ox), py = part_field(a, n, a.oy), pz = part_field(a, n, a.oz), pr = part_field(a, n, a.orad);
  const double idx = 1. / a.dx, idy = 1. / a.dy, idz = 1. / a.dz, irad = 1. / pr;
  int cage[3], lo[3], hi[3];
  cage[0] = (int)(2. * ceil(pr / a.dx)) + 2 - (a.xn % 2);
  cage[1] = (int)(2. * ceil(pr / a.dy)) + 2 - (a.yn % 2);
  cage[2] = (int)(2. * ceil(pr / a.dz)) + 2 - (a.zn % 2);
  lo[0] = (int)(round((px - a.xs) * idx) - 0.5 * cage[0] + DOM_BUF);
  lo[1] = (int)(round((py - a.ys) * idy) - 0.5 * cage[1] + DOM_BUF);
  lo[2] = (int)(round((pz - a.zs) * idz) - 0.5 * cage[2] + DOM_BUF);
#pragma unroll
  for (int ax = 0; ax < 3; ax++) {
    hi[ax] = lo[ax] + cage[ax];
    lo[ax] += (a.S[ax] - lo[ax]) * (lo[ax] < a.S[ax]) + (a.E[ax] - lo[ax]) * (lo[ax] > a.E[ax]);
    hi[ax] += (a.S[ax] - hi[ax]) * (hi[ax] < a.S[ax]) + (a.E[ax] - hi[ax]) * (hi[ax] > a.E[ax]);
    if (lo[ax] == hi[ax]) return;                       /* `is != ie` guard (:230-232): a box clipped to one plane is skipped */
  }
  const int ni = hi[0] - lo[0] + 1, nj = hi[1] - lo[1] + 1, nk = hi[2] - lo[2] + 1;
  const int total = ni * nj * nk;
  for (int e = threadIdx.x; e < total; e += blockDim.x) {
    const int ti = lo[0] + e % ni, tj = lo[1] + (e / ni) % nj, tk = lo[2] + e / (ni * nj);
    const int C = ti + tj * a.s1b + tk * a.s2b;
    const double xx = (ti - 0.5) * a.dx - (px - a.xs);
    const double yy = (tj - 0.5) * a.dy - (py - a.ys);
    const double zz = (tk - 0.5) * a.dz - (pz - a.zs);
    if (!SHELL) {
      const double dd = sqrt(xx * xx + yy * yy + zz * zz);
      if (floor(dd * irad) < 1) atomicMax(a.phase + C, n);             /* phase[C] += cutoff * (n - phase[C]), n ascending (:249) */
    } else {
      if (a.phase[C] != n) continue;
      const double xx_w = (ti - 1 - 0.5) * a.dx - (px - a.xs), xx_e = (ti + 1 - 0.5) * a.dx - (px - a.xs);
      const double yy_s = (tj - 1 - 0.5) * a.dy - (py - a.ys), yy_n = (tj + 1 - 0.5) * a.dy - (py - a.ys);
      const double zz_b = (tk - 1 - 0.5) * a.dz - (pz - a.zs), zz_t = (tk + 1 - 0.5) * a.dz - (pz - a.zs);
      const double d_w = sqrt(xx_w * xx_w + yy * yy + zz * zz), d_e = sqrt(xx_e * xx_e + yy * yy + zz * zz);
      const double d_s = sqrt(xx * xx + yy_s * yy_s + zz * zz), d_n = sqrt(xx * xx + yy_n * yy_n + zz * zz);
      const double d_b = sqrt(xx * xx + yy * yy + zz_b * zz_b), d_t = sqrt(xx * xx + yy * yy + zz_t * zz_t);
      const int in_w = floor(d_w * irad) < 1, in_e = floor(d_e * irad) < 1, in_s = floor(d_s * irad) < 1,
                in_n = floor(d_n * irad) < 1, in_b = floor(d_b * irad) < 1, in_t = floor(d_t * irad) < 1;
      if (!(in_w && in_e && in_s && in_n && in_b && in_t)) a.phase_shell[C] = 0;       /* phase_shell[C] *= 1 - (...), :342-345 */
    }
  }
}

struct CageFlagArgs {
  int *flag_u, *flag_v, *flag_w;   /* OUT: Gfx / Gfy / Gfz s3b */
  const int *phase, *phase_shell;  /* Gcc s3b; NULL when NPARTS == 0 (the reference then leaves the flags at 1, :1524) */
  unsigned ext;                    /* bit 0: face plane Gfx._is is an external wall (flag_external_u), 1: Gfx._ie, 2: Gfy._js, 3: Gfy._je, 4: Gfz._ks, 5: Gfz._ke */
};

/* cage_flag_{u,v,w} value of the face between two cells (:495-497): 1, or -1 on a particle surface / inside the shell */
__device__ __forceinline__ int cage_flag(int p_lo, int p_hi, int s_lo, int s_hi)
{
  return 1 - 2 * ((p_lo < 0 && p_hi > -1) || (p_lo > -1 && p_hi < 0) || (s_hi < 1 && s_lo < 1));
}

/* one thread per cell of the GHOSTED Gcc grid: the cell's W, S, B faces (+ the reset value of the face-grid ghost planes),
 * and for interior cells the solver's masks */
template <int NT>
__global__ void __launch_bounds__(NT) k_cage_flags(const Dev d, const FaceStrides st, const CageFlagArgs a)
{
  const Layout L = d.L;
  const int inb = L.in + 2, jnb = L.jn + 2, knb = L.kn + 2;
  const long long nrows = (long long)jnb * knb;
  const bool parts = a.phase != nullptr;
  bool pushed = false;
  for (long long row = blockIdx.x; row < nrows; row += gridDim.x) {
    const int j = (int)(row % jnb), k = (int)(row / jnb);
    for (int i = threadIdx.x; i < inb; i += NT) {
      const long long C = i + (long long)j * st.cs1b + (long long)k * st.cs2b;
      int pC = -1, sC = 1, pW = -1, sW = 1, pS = -1, sS = 1, pB = -1, sB = 1;
      if (parts) {
        pC = a.phase[C]; sC = a.phase_shell[C];
        if (i >= 1) { pW = a.phase[C - 1]; sW = a.phase_shell[C - 1]; }
        if (j >= 1) { pS = a.phase[C - st.cs1b]; sS = a.phase_shell[C - st.cs1b]; }
        if (k >= 1) { pB = a.phase[C - st.cs2b]; sB = a.phase_shell[C - st.cs2b]; }
      }
      /* flag_u on Gfx (i = 0 .. in+2 along x): faces _is.._ie = 1 .. in+1 for EVERY (j,k) of the ghosted plane, ghost planes keep 1 */
      {
        const long long F = j + (long long)k * st.us1b + (long long)i * st.us2b;
        int f = (parts && i >= 1) ? cage_flag(pW, pC, sW, sC) : 1;
        if ((i == 1 && (a.ext & 1u)) || (i == L.in + 1 && (a.ext & 2u))) f = 0;          /* flag_external_u, :542-552 */
        a.flag_u[F] = f;
        if (i == L.in + 1) a.flag_u[F + st.us2b] = 1;                                    /* plane _ieb */
      }
      {
        const long long F = k + (long long)i * st.vs1b + (long long)j * st.vs2b;
        int f = (parts && j >= 1) ? cage_flag(pS, pC, sS, sC) : 1;
        if ((j == 1 && (a.ext & 4u)) || (j == L.jn + 1 && (a.ext & 8u))) f = 0;
        a.flag_v[F] = f;
        if (j == L.jn + 1) a.flag_v[F + st.vs2b] = 1;
      }
      {
        const long long F = i + (long long)j * st.ws1b + (long long)k * st.ws2b;
        int f = (parts && k >= 1) ? cage_flag(pB, pC, sB, sC) : 1;
        if ((k == 1 && (a.ext & 16u)) || (k == L.kn + 1 && (a.ext & 32u))) f = 0;
        a.flag_w[F] = f;
        if (k == L.kn + 1) a.flag_w[F + st.ws2b] = 1;
      }
      if (i < 1 || i > L.in || j < 1 || j > L.jn || k < 1 || k > L.kn) continue;
      /* ---- the solver's masks of an interior cell (what k_masks digests from the arrays above) ---- */
      const unsigned m = ((i == L.in && (a.ext & 2u)) ? 0u : FM_E) | ((i == 1 && (a.ext & 1u)) ? 0u : FM_W) |
                         ((j == L.jn && (a.ext & 8u)) ? 0u : FM_N) | ((j == 1 && (a.ext & 4u)) ? 0u : FM_S) |
                         ((k == L.kn && (a.ext & 32u)) ? 0u : FM_T) | ((k == 1 && (a.ext & 16u)) ? 0u : FM_B) |
                         ((parts && pC > -1) ? FM_SOLID : 0u) |           /* the particle factors are gathered from this bit of the cell and its neighbours */
                         ((parts && (pC > -1 || pW > -1 || pS > -1 || pB > -1 || a.phase[C + 1] > -1 || a.phase[C + st.cs1b] > -1 ||
                                     a.phase[C + st.cs2b] > -1)) ? FM_NEAR : 0u);
      const long long g = pidx(L, i, j, k);
      d.fmask[g] = (u8)m;
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

#endif
