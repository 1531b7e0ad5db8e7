/* bb_grid.h -- the grid / decomposition contract of Bluebottle's pressure-Poisson path.
 *
 * These types are the BINARY contract between the existing Bluebottle host code and the
 * bbpcg library.  They restate, field for field and in the same order, the reference's
 *   grid_info   (src/domain.h:52-95)
 *   dom_struct  (src/domain.h:168-210)
 *   BC          (src/bluebottle.h:662-747; the pressure types and, for cuda_dom_BC_star, the velocity types/values)
 * and the four index macros (src/bluebottle.h:70-73).  sizeof(dom_struct) must be 880
 * bytes (0x370) -- checked by a static assertion below -- so that a `dom_struct *dom`
 * owned by the reference host program can be handed to this library unchanged.
 *
 * Ghost-cell layout (src/domain.c:1262-1289): one ghost layer (DOM_BUF = 1); block-local
 * interior indices run 1..n, ghosts sit at 0 and n+1.
 */
#ifndef BB_GRID_H
#define BB_GRID_H

#ifdef __cplusplus
extern "C" {
#endif

typedef double real;            /* src/bluebottle.h:45-51 with -DDOUBLE (Makefile:41)  */

#define DOM_BUF 1               /* src/bluebottle.h:141 */

/* pressure boundary-condition codes, src/bluebottle.h:218,230,242 */
#define BB_PERIODIC  0
#define BB_DIRICHLET 1
#define BB_NEUMANN   2

/* "no neighbour" marker.  The reference stores MPI_PROC_NULL (OpenMPI: -2) in
 * dom[].e/w/n/s/t/b (src/domain.c:1147-1210).  The library treats every negative rank
 * as "no neighbour", so either convention works. */
#define BB_PROC_NULL (-2)

/* linear-index macros, src/bluebottle.h:70-73 (note the permuted face grids) */
#define GCC_LOC(II, JJ, KK, S1, S2) ((II) + (JJ)*(S1) + (KK)*(S2))
#define GFX_LOC(II, JJ, KK, S1, S2) ((JJ) + (KK)*(S1) + (II)*(S2))
#define GFY_LOC(II, JJ, KK, S1, S2) ((KK) + (II)*(S1) + (JJ)*(S2))
#define GFZ_LOC(II, JJ, KK, S1, S2) ((II) + (JJ)*(S1) + (KK)*(S2))

typedef struct grid_info {      /* src/domain.h:52-95 */
  int is, ie, in, isb, ieb, inb;
  int js, je, jn, jsb, jeb, jnb;
  int ks, ke, kn, ksb, keb, knb;
  int _is, _ie, _isb, _ieb;
  int _js, _je, _jsb, _jeb;
  int _ks, _ke, _ksb, _keb;
  int s1, s1b, s2, s2b, s3, s3b;
  int s2_i, s2_j, s2_k;
  int s2b_i, s2b_j, s2b_k;
} grid_info;

typedef struct dom_struct {     /* src/domain.h:168-210 */
  grid_info Gcc;
  grid_info Gfx;
  grid_info Gfy;
  grid_info Gfz;
  real xs, xe, xl; int xn; real dx;
  real ys, ye, yl; int yn; real dy;
  real zs, ze, zl; int zn; real dz;
  int rank;
  int e, w, n, s, t, b;
  int I, Is, Ie, In;
  int J, Js, Je, Jn;
  int K, Ks, Ke, Kn;
  int S1, S2, S3;
} dom_struct;

/* The six pressure boundary types, in the order they open the reference's BC struct
 * (src/bluebottle.h:663-668).  A pointer to the reference's `bc` global can be cast to
 * `const bb_pressure_bc *`. */
typedef struct bb_pressure_bc {
  int pW, pE, pS, pN, pB, pT;
} bb_pressure_bc;

/* One velocity entry of the reference's BC struct (src/bluebottle.h:669-740): the type, then the maximum, the current
 * and the acceleration of the DIRICHLET value.  cuda_dom_BC_star reads the type and the current value. */
#define BB_PRECURSOR 3          /* src/bluebottle.h:254: no action in cuda_dom_BC_star's switch */
typedef struct bb_bc_entry { int type; real Dm, D, Da; } bb_bc_entry;

/* The whole BC struct, field for field (src/bluebottle.h:662-747): six pressure types, then u, v, w on W, E, S, N, B, T,
 * then the six screen offsets.  The reference's global `bc` can be read through a `const bb_BC *`. */
typedef struct bb_BC {
  int pW, pE, pS, pN, pB, pT;
  bb_bc_entry u[6], v[6], w[6];                 /* [W, E, S, N, B, T] */
  real dsW, dsE, dsS, dsN, dsB, dsT;
} bb_BC;

/* Explicit-argument form of what cuda_dom_BC_star reads: type[c][f] / val[c][f] of component c (0 u, 1 v, 2 w) on face f
 * (0 W, 1 E, 2 S, 3 N, 4 B, 5 T -- the reference's order). */
typedef struct bb_velocity_bc {
  int  type[3][6];
  real val[3][6];
} bb_velocity_bc;

#if defined(__cplusplus)
static_assert(sizeof(bb_bc_entry) == 32, "BC entry layout (int + pad + 3 reals)");
static_assert(sizeof(bb_BC) == 24 + 18 * 32 + 48, "BC must match the reference (648 bytes)");
#else
_Static_assert(sizeof(bb_bc_entry) == 32, "BC entry layout (int + pad + 3 reals)");
_Static_assert(sizeof(bb_BC) == 24 + 18 * 32 + 48, "BC must match the reference (648 bytes)");
#endif

#if defined(__cplusplus)
static_assert(sizeof(grid_info) == 42 * sizeof(int), "grid_info layout");
static_assert(sizeof(dom_struct) == 880, "dom_struct must match the reference (0x370 bytes)");
#else
_Static_assert(sizeof(grid_info) == 42 * sizeof(int), "grid_info layout");
_Static_assert(sizeof(dom_struct) == 880, "dom_struct must match the reference (0x370 bytes)");
#endif

#ifdef __cplusplus
}
#endif
#endif /* BB_GRID_H */
