/* bbpcg_internal.h -- device-side data layout shared by the kernels and the host driver.
 *
 * PRIVATE PADDED LAYOUT ("P-layout") of the solver's own vectors (r, p0, p1, x) and byte
 * mask.  Same logical extent as the reference's ghosted Gcc grid -- (in+2)(jn+2)(kn+2), one
 * ghost layer, interior 1..n (src/domain.c:1262-1289) -- but every x-row is padded so that
 * interior cell i = 1 starts on a 128-byte boundary:
 *
 *     offset(i,j,k) = (i + 15) + j*px + k*ps,   px = roundup(in + 17, 16),  ps = px*(jn+2)
 *
 * ghost i = 0 sits at 15, interior at 16..16+in-1, ghost i = in+1 at 16+in.  All four vectors
 * and the masks share one index, so one offset addresses a cell in every array.
 */
#ifndef BBPCG_INTERNAL_H
#define BBPCG_INTERNAL_H

#include <cuda_runtime.h>
#include "../../include/bbpcg.h"

#define BB_XOFF 15
#define BB_MAXR BBPCG_MAX_RANKS
#define BB_NSLOT 4                 /* mailbox slots (a peer is at most one stage ahead) */
#define BB_MAXBLOCKS 65536         /* max CTAs of a reducing kernel */
#define BB_GROUP 64                /* CTAs per first-level reduction group */
#define BB_MAXGROUPS (BB_MAXBLOCKS / BB_GROUP)
#define BB_FLAT_MAX 1024           /* grids up to this many CTAs reduce in ONE level (a single group) */
#define BB_MAXZ 1024               /* max z-chunks of the iteration kernels */

/* coefficient mask bits (fmask): squared face flags, src/solver_kernel.cu:824-829 */
#define FM_E 1u
#define FM_W 2u
#define FM_N 4u
#define FM_S 8u
#define FM_T 16u
#define FM_B 32u
/* A mask byte with NO flag bit set marks a DEAD cell -- a ghost behind an external wall (the arena is zero-filled and nobody
 * writes those): its Jacobi entry is 0, so z = p = 0 there.  (A real cell whose six faces are all walls, M = 0, would be
 * 1/0 in the reference; it gets 0 here.) */
#define FM_NEAR 64u                /* the cell or one of its six neighbours lies inside a particle: only then does a kernel
                                      gather the particle factors; the all-ones test of the fast path (0x3f) excludes it for free */
#define FM_SOLID 128u              /* phase > -1: the cell lies inside a particle (src/solver_kernel.cu:683-695).  The particle factors
                                      of a cell are gathered from this bit of the cell and of its six neighbours in the halo'd mask
                                      tile -- no second mask array, no second stream (a separate 1-byte pmask tile cost the search
                                      kernel 12 %: one more TMA request per plane) */
/* particle factors of one cell, as gathered by the kernels: solid bit of the cell and of its neighbours */
#define PM_C 1u                    /* phase[C] > -1 (solid) */
#define PM_E 2u
#define PM_W 4u
#define PM_N 8u
#define PM_S 16u
#define PM_T 32u
#define PM_B 64u

struct Layout {
  int in, jn, kn;
  int px;
  long long ps;
  long long n;                     /* elements per array = ps*(kn+2) */
};

static inline __host__ __device__ long long pidx(const Layout &L, int i, int j, int k)
{
  return (long long)(i + BB_XOFF) + (long long)j * L.px + (long long)k * L.ps;
}

static inline Layout make_layout(int in, int jn, int kn)
{
  Layout L;
  L.in = in; L.jn = jn; L.kn = kn;
  L.px = ((in + 17) + 15) / 16 * 16;
  L.ps = (long long)L.px * (jn + 2);
  L.n = L.ps * (kn + 2);
  return L;
}

/* byte offsets of everything inside a rank's single device allocation ("arena").  A pure
 * function of (in,jn,kn), so a peer can address a neighbour's arrays from its dimensions. */
struct ArenaMap {
  size_t r, p0, p1, x;             /* doubles[L.n] */
  size_t fmask;                    /* bytes[L.n]   */
  size_t recv[2][6];               /* doubles, generic exchange staging (double-buffered) */
  size_t partials;                 /* doubles[4*BB_MAXBLOCKS]: up to 4 values per CTA */
  size_t gpartials;                /* doubles[4*BB_MAXGROUPS]: group sums */
  size_t counter;                  /* unsigned[4 + BB_MAXGROUPS]: [0] groups done, [4+g] CTAs of group g done */
  size_t scal;                     /* Scal */
  size_t mbox;                     /* u64[BB_NSLOT][BB_MAXR][4]: {32 data bits | 32-bit tag} words */
  size_t history;                  /* doubles[hist_cap] */
  size_t invM_tab;                 /* doubles[128]: Jacobi diagonal per mask value */
  size_t ztab;                     /* ints[2 * BB_MAXZ]: first and last plane of every z-chunk of the iteration kernels, in claim order */
  size_t total;
};

#define BB_HIST_CAP 65536

static inline ArenaMap make_arena_map(const Layout &L)
{
  ArenaMap m;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) / 256 * 256; return o; };
  size_t vec = (size_t)L.n * sizeof(double);
  m.r = take(vec); m.p0 = take(vec); m.p1 = take(vec); m.x = take(vec);
  m.fmask = take((size_t)L.n);
  /* staging faces sized for the largest of the four grids (a face grid is one entry longer along its normal) */
  size_t fi = (size_t)(L.jn + 1) * (L.kn + 1), fj = (size_t)(L.in + 1) * (L.kn + 1), fk = (size_t)(L.in + 1) * (L.jn + 1);
  for (int b = 0; b < 2; b++) for (int f = 0; f < 6; f++)
    m.recv[b][f] = take(sizeof(double) * (f < 2 ? fi : f < 4 ? fj : fk));
  m.partials = take(sizeof(double) * 4 * BB_MAXBLOCKS);
  m.gpartials = take(sizeof(double) * 4 * BB_MAXGROUPS);
  m.counter = take(sizeof(unsigned) * (4 + BB_MAXGROUPS));
  m.scal = take(512);
  m.mbox = take(sizeof(unsigned long long) * BB_NSLOT * BB_MAXR * 4);
  m.history = take(sizeof(double) * BB_HIST_CAP);
  m.invM_tab = take(sizeof(double) * 128);
  m.ztab = take(sizeof(int) * (2 * BB_MAXZ));
  m.total = off;
  return m;
}

/* Device-resident solver scalars: no per-iteration host round trip (the reference syncs the
 * host four times per iteration, src/cuda_solver.cu:204-232). */
struct Scal {
  double bb;            /* (b,b)                       cuda_solver.cu:151 */
  double rz;            /* current (r,z) = sp_rq       :169 / :267        */
  double pAp;           /* DENOM                       :204               */
  double alpha;         /* step of the iteration in flight                */
  double alpha_x;       /* step not yet applied to x (lazy phi update)    */
  double beta;          /* for the next search update  :256               */
  double resid;         /* sqrt(rz)/sqrt(bb) at exit   :239               */
  double rz0;           /* initial (r,z)                                  */
  double tol2;          /* pp_residual^2                                  */
  int done;             /* 1: every later kernel of this solve is a no-op */
  int status;           /* BBPCG_CONVERGED ...                            */
  int q;                /* completed iterations                           */
  int max_q;            /* pp_max_iter + 1  (loop bound, :192)            */
  int fixed;            /* benchmark mode: no stop test                   */
  int comm_timeout;     /* set if a peer never showed up                  */
  int pad0, pad1;
  unsigned long long seq;   /* publish counter, never reset               */
  double p_sum;         /* epilogue: sum of p over all ranks' interior cells (cuda_bluebottle.cu:2526-2529) */
  double eps[3];        /* solvability: net outflow per axis over all ranks (cuda_bluebottle.cu:2414-2420) */
};

/* where this block's boundary values go: the neighbour's (or, for a periodic wrap onto the
 * same block, this rank's own) arrays.  r == NULL: external wall / no neighbour. */
struct NbrFace {
  double *r, *x;
  unsigned char *fmask;
  double *recv[2];      /* neighbour's staging buffer for the OPPOSITE face (generic exchange) */
  Layout L;
};
struct Halo { NbrFace f[6]; };     /* 0:E 1:W 2:N 3:S 4:T 5:B */

struct Comm {
  int rank, nranks;
  long long timeout_cycles;                   /* spin limit of the in-kernel all-reduce; <= 0: wait for ever (as MPI does) */
  int *host_flag;                             /* pinned host word, set when a wait timed out: every entry point reads it after its stream sync */
  unsigned long long *mbox[BB_MAXR];          /* rank p's mailbox (mapped) */
};

/* everything a kernel needs about this rank, passed by value */
struct Dev {
  Layout L;
  double *r, *P[2], *x;
  unsigned char *fmask;
  double *recv[2][6];
  double *partials, *gpartials;
  unsigned *counter;
  Scal *sc;
  double *history;
  const double *invM_tab;         /* [128], built once by k_build_tab */
  const int *ztab;                /* [2 nbz]: z-chunk c of the iteration kernels owns planes ztab[2c] .. ztab[2c+1] */
  double idx2, idy2, idz2;        /* 1/(dx*dx) ...  (per block, src/solver_kernel.cu:720-722) */
  double dx2_6, dy2_6, dz2_6;     /* dx*dx/6 ...    (src/solver_kernel.cu:683)                */
  Halo halo;
  Comm comm;
  int any_nbr;                    /* some face has a neighbour (another rank or a periodic self-wrap) */
#ifdef BB_TRACE
  unsigned long long *trace;      /* debug build only (make trace): per-CTA globaltimer stamps of the iteration kernels */
#endif
};

/* ---- optional per-CTA time line of the two iteration kernels (libbbpcg_trace.so, scripts/trace_timeline.py) ---- */
#ifdef BB_TRACE
#define BB_TRACE_EV 8
#define BB_TRACE_LAUNCHES 64
#define BB_TRACE_CTAS 4096
#endif

#endif
