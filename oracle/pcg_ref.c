/* oracle/pcg_ref.c -- CPU restatement ("O2") of Bluebottle-3.0's pressure-Poisson PCG path.
 *
 * THIS FILE IS TEST INFRASTRUCTURE.  It is the checker for the CUDA product, never the
 * product: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
 * leg may build, load or call it.  Nothing under bluebottle-3.0_b200/ links or imports it.
 *
 * Parity status: the reference ships NO golden vector / known-answer test for this path
 * (SURVEY.md 8c).  The restatement is therefore pinned against the reference's own
 * kernels compiled unmodified from /root/reference/src into oracle/_ref/ (see
 * oracle/Makefile, oracle/ref_shim.cu) and run on a B200; the recorded outputs of that
 * run are committed under tests/golden/ (see tests/golden/README.md for what is pinned).
 *
 * Every function cites the reference file:line it follows (paths relative to
 * /root/reference).  The arithmetic keeps the reference's association order; it is
 * compiled with -ffp-contract=off so no FMA contraction sneaks in (the reference kernels
 * are nvcc-contracted, so agreement with O1 is to round-off, not bit-for-bit).
 *
 * Layout: all arrays are block-local and ghosted exactly like the reference
 * (include/bb_grid.h).  A bbo_state holds ALL blocks of a decomposition in one process, so
 * the multi-rank algorithm (halo exchange, rank-ordered allreduce) is executed faithfully
 * on the CPU; nblocks == 1 is the single-GPU case.
 *
 * OpenMP: every grid loop is `omp parallel for` over k-planes; dot products accumulate a
 * per-plane partial and the planes are then summed serially in k order, so results are
 * bit-identical for any thread count (and with OpenMP disabled).
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../include/bb_grid.h"

#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------------------ */
typedef struct bbo_block {
  int *flag_u, *flag_v, *flag_w;       /* Gfx/Gfy/Gfz s3b, int   (particle.h:256-331)      */
  int *phase, *phase_shell;            /* Gcc s3b, int           (particle.h:97,109)       */
  real *u_star, *v_star, *w_star;      /* Gfx/Gfy/Gfz s3b        (bluebottle.h:1088-1112)  */
  real *rhs_p, *phi, *pb_q;            /* Gcc s3b                (bluebottle.h:974-1013)   */
  real *invM, *r_q, *z_q, *p_q, *Apb_q;/* Gcc s3                                          */
  real *send[6], *recv[6];             /* e,w,n,s,t,b packing buffers (cuda_bluebottle.cu:255-319) */
  real *u, *v, *w;                     /* Gfx/Gfy/Gfz s3b: projected velocity (bluebottle.h:1137,1186,1235) */
  real *p0, *p;                        /* Gcc s3b: previous / updated pressure (bluebottle.h:961,1037)  */
} bbo_block;

typedef struct bbo_state {
  dom_struct DOM;
  dom_struct *dom;                     /* [S3] */
  bb_pressure_bc bc;
  int nblocks;
  bbo_block *blk;
  real *plane_partial;                 /* scratch for deterministic dots */
  int plane_cap;
} bbo_state;

enum { BBO_E = 0, BBO_W, BBO_N, BBO_S, BBO_T, BBO_B };

/* array ids for bbo_array() */
enum {
  BBO_FLAG_U = 0, BBO_FLAG_V, BBO_FLAG_W, BBO_PHASE, BBO_PHASE_SHELL,
  BBO_U_STAR, BBO_V_STAR, BBO_W_STAR, BBO_RHS_P, BBO_PHI, BBO_PB_Q,
  BBO_INVM, BBO_R_Q, BBO_Z_Q, BBO_P_Q, BBO_APB_Q,
  BBO_VEL_U, BBO_VEL_V, BBO_VEL_W, BBO_P0, BBO_P
};

/* ------------------------------------------------------------------------------------ */
/* domain_fill: src/domain.c:918-1486.  One helper per grid replaces the four unrolled
 * copies in the reference; the values produced are the same. */
static void fill_grid_local(grid_info *g, int in, int jn, int kn, int gis, int gjs, int gks,
                            int order /* 0: i,j,k  1: j,k,i (Gfx)  2: k,i,j (Gfy) */)
{
  g->is = gis; g->isb = g->is - DOM_BUF; g->in = in; g->inb = in + 2 * DOM_BUF;
  g->ie = g->isb + g->in; g->ieb = g->ie + DOM_BUF;
  g->js = gjs; g->jsb = g->js - DOM_BUF; g->jn = jn; g->jnb = jn + 2 * DOM_BUF;
  g->je = g->jsb + g->jn; g->jeb = g->je + DOM_BUF;
  g->ks = gks; g->ksb = g->ks - DOM_BUF; g->kn = kn; g->knb = kn + 2 * DOM_BUF;
  g->ke = g->ksb + g->kn; g->keb = g->ke + DOM_BUF;

  g->_is = DOM_BUF; g->_isb = g->_is - DOM_BUF; g->_ie = g->_isb + g->in; g->_ieb = g->_ie + DOM_BUF;
  g->_js = DOM_BUF; g->_jsb = g->_js - DOM_BUF; g->_je = g->_jsb + g->jn; g->_jeb = g->_je + DOM_BUF;
  g->_ks = DOM_BUF; g->_ksb = g->_ks - DOM_BUF; g->_ke = g->_ksb + g->kn; g->_keb = g->_ke + DOM_BUF;

  if (order == 0) {            /* Gcc domain.c:1277-1282, Gfz :1466-1471 */
    g->s1 = g->in;  g->s2 = g->s1 * g->jn;  g->s3 = g->s2 * g->kn;
    g->s1b = g->inb; g->s2b = g->s1b * g->jnb; g->s3b = g->s2b * g->knb;
  } else if (order == 1) {     /* Gfx domain.c:1340-1345 */
    g->s1 = g->jn;  g->s2 = g->s1 * g->kn;  g->s3 = g->s2 * g->in;
    g->s1b = g->jnb; g->s2b = g->s1b * g->knb; g->s3b = g->s2b * g->inb;
  } else {                     /* Gfy domain.c:1403-1408 */
    g->s1 = g->kn;  g->s2 = g->s1 * g->in;  g->s3 = g->s2 * g->jn;
    g->s1b = g->knb; g->s2b = g->s1b * g->inb; g->s3b = g->s2b * g->jnb;
  }
  g->s2_i = g->jn * g->kn;   g->s2_j = g->in * g->kn;   g->s2_k = g->in * g->jn;
  g->s2b_i = g->jnb * g->knb; g->s2b_j = g->inb * g->knb; g->s2b_k = g->inb * g->jnb;
}

/* Global (DOM) grids: domain.c:934-1134 -- same formulas with is=js=ks=DOM_BUF. */
static void fill_DOM(dom_struct *D)
{
  D->xl = D->xe - D->xs; D->yl = D->ye - D->ys; D->zl = D->ze - D->zs;
  D->dx = D->xl / D->xn; D->dy = D->yl / D->yn; D->dz = D->zl / D->zn;
  fill_grid_local(&D->Gcc, D->xn,     D->yn,     D->zn,     DOM_BUF, DOM_BUF, DOM_BUF, 0);
  fill_grid_local(&D->Gfx, D->xn + 1, D->yn,     D->zn,     DOM_BUF, DOM_BUF, DOM_BUF, 1);
  fill_grid_local(&D->Gfy, D->xn,     D->yn + 1, D->zn,     DOM_BUF, DOM_BUF, DOM_BUF, 2);
  fill_grid_local(&D->Gfz, D->xn,     D->yn,     D->zn + 1, DOM_BUF, DOM_BUF, DOM_BUF, 0);
  D->Is = 0; D->Ie = D->In - 1; D->Js = 0; D->Je = D->Jn - 1; D->Ks = 0; D->Ke = D->Kn - 1;
}

void bbo_domain_fill(dom_struct *DOM, dom_struct *dom, const bb_pressure_bc *bc)
{
  int i, j, k, c;
  DOM->S1 = DOM->In; DOM->S2 = DOM->S1 * DOM->Jn; DOM->S3 = DOM->S2 * DOM->Kn; /* domain.c:121-123 */
  fill_DOM(DOM);

  /* neighbours from the PRESSURE boundary types, domain.c:1147-1210 */
  for (c = 0; c < DOM->S3; c++) {
    dom_struct *d = &dom[c];
    d->rank = d->I + d->J * DOM->S1 + d->K * DOM->S2;               /* domain.c:141 */
    if (d->I == DOM->Is) d->w = (bc->pW == BB_PERIODIC) ? DOM->Ie + d->J * DOM->S1 + d->K * DOM->S2 : BB_PROC_NULL;
    else                 d->w = (d->I - 1) + d->J * DOM->S1 + d->K * DOM->S2;
    if (d->I == DOM->Ie) d->e = (bc->pE == BB_PERIODIC) ? DOM->Is + d->J * DOM->S1 + d->K * DOM->S2 : BB_PROC_NULL;
    else                 d->e = (d->I + 1) + d->J * DOM->S1 + d->K * DOM->S2;
    if (d->J == DOM->Js) d->s = (bc->pS == BB_PERIODIC) ? d->I + DOM->Je * DOM->S1 + d->K * DOM->S2 : BB_PROC_NULL;
    else                 d->s = d->I + (d->J - 1) * DOM->S1 + d->K * DOM->S2;
    if (d->J == DOM->Je) d->n = (bc->pN == BB_PERIODIC) ? d->I + DOM->Js * DOM->S1 + d->K * DOM->S2 : BB_PROC_NULL;
    else                 d->n = d->I + (d->J + 1) * DOM->S1 + d->K * DOM->S2;
    if (d->K == DOM->Ks) d->b = (bc->pB == BB_PERIODIC) ? d->I + d->J * DOM->S1 + DOM->Ke * DOM->S2 : BB_PROC_NULL;
    else                 d->b = d->I + d->J * DOM->S1 + (d->K - 1) * DOM->S2;
    if (d->K == DOM->Ke) d->t = (bc->pT == BB_PERIODIC) ? d->I + d->J * DOM->S1 + DOM->Ks * DOM->S2 : BB_PROC_NULL;
    else                 d->t = d->I + d->J * DOM->S1 + (d->K + 1) * DOM->S2;
    /* the block-count bookkeeping the reference leaves in each dom[] entry is unused on this path */
    d->Is = DOM->Is; d->Ie = DOM->Ie; d->In = DOM->In;
    d->Js = DOM->Js; d->Je = DOM->Je; d->Jn = DOM->Jn;
    d->Ks = DOM->Ks; d->Ke = DOM->Ke; d->Kn = DOM->Kn;
    d->S1 = DOM->S1; d->S2 = DOM->S2; d->S3 = DOM->S3;
  }

  /* per-block lengths and index ranges, domain.c:1213-1478.  Global start indices chain
   * off the west/south/bottom neighbour; face grids share the block-boundary face
   * (Gfx.is = west.Gfx.ie, domain.c:1292-1296).  (The reference's Gcc.is uses
   * dom[dom[i].w] at :1231 -- a latent typo for dom[dom[c].w]; the global `is` is not read
   * on this path, we use the evidently intended neighbour.) */
  for (k = 0; k < DOM->Kn; k++) for (j = 0; j < DOM->Jn; j++) for (i = 0; i < DOM->In; i++) {
    dom_struct *d;
    int gis, gjs, gks;
    c = GCC_LOC(i, j, k, DOM->S1, DOM->S2);
    d = &dom[c];
    d->xl = d->xe - d->xs; d->yl = d->ye - d->ys; d->zl = d->ze - d->zs;
    d->dx = d->xl / d->xn; d->dy = d->yl / d->yn; d->dz = d->zl / d->zn;
    {
      const dom_struct *W = (i == DOM->Is) ? NULL : &dom[GCC_LOC(i - 1, j, k, DOM->S1, DOM->S2)];
      const dom_struct *S = (j == DOM->Js) ? NULL : &dom[GCC_LOC(i, j - 1, k, DOM->S1, DOM->S2)];
      const dom_struct *B = (k == DOM->Ks) ? NULL : &dom[GCC_LOC(i, j, k - 1, DOM->S1, DOM->S2)];
      gis = W ? W->Gcc.ie + 1 : DOM_BUF; gjs = S ? S->Gcc.je + 1 : DOM_BUF; gks = B ? B->Gcc.ke + 1 : DOM_BUF;
      fill_grid_local(&d->Gcc, d->xn, d->yn, d->zn, gis, gjs, gks, 0);
      gis = W ? W->Gfx.ie : DOM_BUF; gjs = S ? S->Gfx.je + 1 : DOM_BUF; gks = B ? B->Gfx.ke + 1 : DOM_BUF;
      fill_grid_local(&d->Gfx, d->xn + 1, d->yn, d->zn, gis, gjs, gks, 1);
      gis = W ? W->Gfy.ie + 1 : DOM_BUF; gjs = S ? S->Gfy.je : DOM_BUF; gks = B ? B->Gfy.ke + 1 : DOM_BUF;
      fill_grid_local(&d->Gfy, d->xn, d->yn + 1, d->zn, gis, gjs, gks, 2);
      gis = W ? W->Gfz.ie + 1 : DOM_BUF; gjs = S ? S->Gfz.je + 1 : DOM_BUF; gks = B ? B->Gfz.ke : DOM_BUF;
      fill_grid_local(&d->Gfz, d->xn, d->yn, d->zn + 1, gis, gjs, gks, 0);
    }
  }
}

/* ------------------------------------------------------------------------------------ */
/* Equal-split decomposition exactly as tools/src/decomp_reader.c:112-129 computes it, i.e.
 * what a decomp.config written by that tool contains. */
static void split_axis(real Xs, real Xe, int Xn, int In, int i, real *xs, real *xe, int *xn)
{
  real Xl = Xe - Xs;
  real xl = Xl / In;
  *xn = Xn / In;
  *xs = Xs + i * xl;
  *xe = *xs + xl;
}

static void *xcalloc(size_t n, size_t sz)
{
  void *p = calloc(n ? n : 1, sz);
  if (!p) { fprintf(stderr, "bbo: out of memory (%zu x %zu)\n", n, sz); exit(EXIT_FAILURE); }
  return p;
}

static void alloc_block(bbo_block *b, const dom_struct *d)
{
  int f;
  b->flag_u = xcalloc(d->Gfx.s3b, sizeof(int)); b->flag_v = xcalloc(d->Gfy.s3b, sizeof(int));
  b->flag_w = xcalloc(d->Gfz.s3b, sizeof(int));
  b->phase = xcalloc(d->Gcc.s3b, sizeof(int)); b->phase_shell = xcalloc(d->Gcc.s3b, sizeof(int));
  b->u_star = xcalloc(d->Gfx.s3b, sizeof(real)); b->v_star = xcalloc(d->Gfy.s3b, sizeof(real));
  b->w_star = xcalloc(d->Gfz.s3b, sizeof(real));
  b->rhs_p = xcalloc(d->Gcc.s3b, sizeof(real)); b->phi = xcalloc(d->Gcc.s3b, sizeof(real));
  b->pb_q = xcalloc(d->Gcc.s3b, sizeof(real));   /* cudaMemset 0, cuda_bluebottle.cu:373 */
  b->u = xcalloc(d->Gfx.s3b, sizeof(real)); b->v = xcalloc(d->Gfy.s3b, sizeof(real)); b->w = xcalloc(d->Gfz.s3b, sizeof(real));
  b->p0 = xcalloc(d->Gcc.s3b, sizeof(real)); b->p = xcalloc(d->Gcc.s3b, sizeof(real));
  b->invM = xcalloc(d->Gcc.s3, sizeof(real)); b->r_q = xcalloc(d->Gcc.s3, sizeof(real));
  b->z_q = xcalloc(d->Gcc.s3, sizeof(real)); b->p_q = xcalloc(d->Gcc.s3, sizeof(real));
  b->Apb_q = xcalloc(d->Gcc.s3, sizeof(real));
  for (f = 0; f < 6; f++) {     /* sized for the largest of the four grids (cuda_bluebottle.cu:255-319 has one set per grid) */
    int n = (f < 2) ? (d->yn + 1) * (d->zn + 1) : (f < 4) ? (d->xn + 1) * (d->zn + 1) : (d->xn + 1) * (d->yn + 1);
    b->send[f] = xcalloc(n, sizeof(real)); b->recv[f] = xcalloc(n, sizeof(real));
  }
  /* host initialisation when NPARTS == 0: phase = phase_shell = -1 (particle.c:556-568) */
  for (f = 0; f < d->Gcc.s3b; f++) { b->phase[f] = -1; b->phase_shell[f] = -1; }
}

/* Build a state from explicit per-block extents (the content of a decomp.config:
 * S3 records in I-fastest order, domain.c:138-159).  ext[c] = {xs,xe,ys,ye,zs,ze},
 * nn[c] = {xn,yn,zn}. */
bbo_state *bbo_create_blocks(const real G[6], const int Gn[3], const int IJK[3], const int pbc[6],
                             const real *ext, const int *nn)
{
  bbo_state *s = xcalloc(1, sizeof(*s));
  int i, j, k, c, maxk = 0;
  s->DOM.xs = G[0]; s->DOM.xe = G[1]; s->DOM.xn = Gn[0];
  s->DOM.ys = G[2]; s->DOM.ye = G[3]; s->DOM.yn = Gn[1];
  s->DOM.zs = G[4]; s->DOM.ze = G[5]; s->DOM.zn = Gn[2];
  s->DOM.In = IJK[0]; s->DOM.Jn = IJK[1]; s->DOM.Kn = IJK[2];
  s->bc.pW = pbc[0]; s->bc.pE = pbc[1]; s->bc.pS = pbc[2]; s->bc.pN = pbc[3]; s->bc.pB = pbc[4]; s->bc.pT = pbc[5];
  s->nblocks = IJK[0] * IJK[1] * IJK[2];
  s->dom = xcalloc(s->nblocks, sizeof(dom_struct));
  for (k = 0; k < IJK[2]; k++) for (j = 0; j < IJK[1]; j++) for (i = 0; i < IJK[0]; i++) {
    c = i + j * IJK[0] + k * IJK[0] * IJK[1];
    s->dom[c].I = i; s->dom[c].J = j; s->dom[c].K = k;
    s->dom[c].xs = ext[6 * c + 0]; s->dom[c].xe = ext[6 * c + 1]; s->dom[c].xn = nn[3 * c + 0];
    s->dom[c].ys = ext[6 * c + 2]; s->dom[c].ye = ext[6 * c + 3]; s->dom[c].yn = nn[3 * c + 1];
    s->dom[c].zs = ext[6 * c + 4]; s->dom[c].ze = ext[6 * c + 5]; s->dom[c].zn = nn[3 * c + 2];
  }
  bbo_domain_fill(&s->DOM, s->dom, &s->bc);
  s->blk = xcalloc(s->nblocks, sizeof(bbo_block));
  for (c = 0; c < s->nblocks; c++) {
    alloc_block(&s->blk[c], &s->dom[c]);
    if (s->dom[c].Gcc.knb > maxk) maxk = s->dom[c].Gcc.knb;
  }
  s->plane_cap = maxk + 4;
  s->plane_partial = xcalloc(s->plane_cap, sizeof(real));
  return s;
}

/* Equal splits, as decomp_reader writes them. */
bbo_state *bbo_create(const real G[6], const int Gn[3], const int IJK[3], const int pbc[6])
{
  int S3 = IJK[0] * IJK[1] * IJK[2], i, j, k, c;
  real *ext = xcalloc(6 * S3, sizeof(real));
  int *nn = xcalloc(3 * S3, sizeof(int));
  bbo_state *s;
  for (k = 0; k < IJK[2]; k++) for (j = 0; j < IJK[1]; j++) for (i = 0; i < IJK[0]; i++) {
    c = i + j * IJK[0] + k * IJK[0] * IJK[1];
    split_axis(G[0], G[1], Gn[0], IJK[0], i, &ext[6 * c + 0], &ext[6 * c + 1], &nn[3 * c + 0]);
    split_axis(G[2], G[3], Gn[1], IJK[1], j, &ext[6 * c + 2], &ext[6 * c + 3], &nn[3 * c + 1]);
    split_axis(G[4], G[5], Gn[2], IJK[2], k, &ext[6 * c + 4], &ext[6 * c + 5], &nn[3 * c + 2]);
  }
  s = bbo_create_blocks(G, Gn, IJK, pbc, ext, nn);
  free(ext); free(nn);
  return s;
}

void bbo_destroy(bbo_state *s)
{
  int c, f;
  if (!s) return;
  for (c = 0; c < s->nblocks; c++) {
    bbo_block *b = &s->blk[c];
    free(b->flag_u); free(b->flag_v); free(b->flag_w); free(b->phase); free(b->phase_shell);
    free(b->u_star); free(b->v_star); free(b->w_star); free(b->rhs_p); free(b->phi); free(b->pb_q);
    free(b->invM); free(b->r_q); free(b->z_q); free(b->p_q); free(b->Apb_q);
    free(b->u); free(b->v); free(b->w); free(b->p0); free(b->p);
    for (f = 0; f < 6; f++) { free(b->send[f]); free(b->recv[f]); }
  }
  free(s->blk); free(s->dom); free(s->plane_partial); free(s);
}

int bbo_nblocks(const bbo_state *s) { return s->nblocks; }
const dom_struct *bbo_dom(const bbo_state *s, int rank) { return &s->dom[rank]; }
const dom_struct *bbo_DOM(const bbo_state *s) { return &s->DOM; }

void *bbo_array(bbo_state *s, int rank, int id)
{
  bbo_block *b = &s->blk[rank];
  switch (id) {
    case BBO_FLAG_U: return b->flag_u;  case BBO_FLAG_V: return b->flag_v;  case BBO_FLAG_W: return b->flag_w;
    case BBO_PHASE: return b->phase;    case BBO_PHASE_SHELL: return b->phase_shell;
    case BBO_U_STAR: return b->u_star;  case BBO_V_STAR: return b->v_star;  case BBO_W_STAR: return b->w_star;
    case BBO_RHS_P: return b->rhs_p;    case BBO_PHI: return b->phi;        case BBO_PB_Q: return b->pb_q;
    case BBO_INVM: return b->invM;      case BBO_R_Q: return b->r_q;        case BBO_Z_Q: return b->z_q;
    case BBO_P_Q: return b->p_q;        case BBO_APB_Q: return b->Apb_q;
    case BBO_VEL_U: return b->u;        case BBO_VEL_V: return b->v;        case BBO_VEL_W: return b->w;
    case BBO_P0: return b->p0;          case BBO_P: return b->p;
  }
  return NULL;
}

/* ------------------------------------------------------------------------------------ */
/* decomp.config writer/reader: record grammar of domain.c:138-159, written the way
 * tools/src/decomp_reader.c:142-154 writes it. */
int bbo_write_decomp(const bbo_state *s, const char *path, int prec)
{
  FILE *f = fopen(path, "w");
  int c;
  if (!f) return -1;
  for (c = 0; c < s->nblocks; c++) {
    const dom_struct *d = &s->dom[c];
    fprintf(f, "(I, J, K) %d %d %d\n", d->I, d->J, d->K);
    fprintf(f, "(Xs, Xe, Xn) %.*lf %.*lf %d\n", prec, d->xs, prec, d->xe, d->xn);
    fprintf(f, "(Ys, Ye, Yn) %.*lf %.*lf %d\n", prec, d->ys, prec, d->ye, d->yn);
    fprintf(f, "(Zs, Ze, Zn) %.*lf %.*lf %d\n\n", prec, d->zs, prec, d->ze, d->zn);
  }
  fclose(f);
  return 0;
}

/* Reads S3 records with the reference's own fscanf formats (domain.c:138-159).
 * ext/nn/ijk sized 6*S3 / 3*S3 / 3*S3.  Returns the number of records read. */
int bbo_read_decomp(const char *path, int S3, real *ext, int *nn, int *ijk)
{
  FILE *f = fopen(path, "r");
  int i, fret = 0, nread = 0;
  if (!f) return -1;
  for (i = 0; i < S3; i++) {
    fret = fscanf(f, "(I, J, K) %d %d %d\n", &ijk[3 * i], &ijk[3 * i + 1], &ijk[3 * i + 2]);
    if (fret != 3) break;
    fret = fscanf(f, "(Xs, Xe, Xn) %lf %lf %d\n", &ext[6 * i + 0], &ext[6 * i + 1], &nn[3 * i + 0]);
    fret += fscanf(f, "(Ys, Ye, Yn) %lf %lf %d\n", &ext[6 * i + 2], &ext[6 * i + 3], &nn[3 * i + 1]);
    fret += fscanf(f, "(Zs, Ze, Zn) %lf %lf %d\n", &ext[6 * i + 4], &ext[6 * i + 5], &nn[3 * i + 2]);
    if (fret != 9) break;
    fret = fscanf(f, "\n");
    nread++;
  }
  fclose(f);
  return nread;
}

/* ------------------------------------------------------------------------------------ */
/* Flags when NPARTS == 0: cuda_build_cages, src/cuda_particle.cu:1516-1521 (reset to 1:
 * particle_kernel.cu:79-119) then :1600-1639 (external walls -> 0 on the wall face plane,
 * all jnb x knb entries, particle_kernel.cu:542-576), applied only if BOTH sides of the
 * direction are non-periodic and only on blocks that touch the wall. */
void bbo_build_flags_noparts(bbo_state *s)
{
  int c, i, j, k;
  for (c = 0; c < s->nblocks; c++) {
    const dom_struct *d = &s->dom[c];
    bbo_block *b = &s->blk[c];
    for (i = 0; i < d->Gfx.s3b; i++) b->flag_u[i] = 1;
    for (i = 0; i < d->Gfy.s3b; i++) b->flag_v[i] = 1;
    for (i = 0; i < d->Gfz.s3b; i++) b->flag_w[i] = 1;
    if (s->bc.pW != BB_PERIODIC && s->bc.pE != BB_PERIODIC) {
      if (d->I == s->DOM.Is) for (k = 0; k < d->Gfx.knb; k++) for (j = 0; j < d->Gfx.jnb; j++)
        b->flag_u[GFX_LOC(d->Gfx._is, j, k, d->Gfx.s1b, d->Gfx.s2b)] = 0;
      if (d->I == s->DOM.Ie) for (k = 0; k < d->Gfx.knb; k++) for (j = 0; j < d->Gfx.jnb; j++)
        b->flag_u[GFX_LOC(d->Gfx._ie, j, k, d->Gfx.s1b, d->Gfx.s2b)] = 0;
    }
    if (s->bc.pS != BB_PERIODIC && s->bc.pN != BB_PERIODIC) {
      if (d->J == s->DOM.Js) for (k = 0; k < d->Gfy.knb; k++) for (i = 0; i < d->Gfy.inb; i++)
        b->flag_v[GFY_LOC(i, d->Gfy._js, k, d->Gfy.s1b, d->Gfy.s2b)] = 0;
      if (d->J == s->DOM.Je) for (k = 0; k < d->Gfy.knb; k++) for (i = 0; i < d->Gfy.inb; i++)
        b->flag_v[GFY_LOC(i, d->Gfy._je, k, d->Gfy.s1b, d->Gfy.s2b)] = 0;
    }
    if (s->bc.pB != BB_PERIODIC && s->bc.pT != BB_PERIODIC) {
      if (d->K == s->DOM.Ks) for (j = 0; j < d->Gfz.jnb; j++) for (i = 0; i < d->Gfz.inb; i++)
        b->flag_w[GFZ_LOC(i, j, d->Gfz._ks, d->Gfz.s1b, d->Gfz.s2b)] = 0;
      if (d->K == s->DOM.Ke) for (j = 0; j < d->Gfz.jnb; j++) for (i = 0; i < d->Gfz.inb; i++)
        b->flag_w[GFZ_LOC(i, j, d->Gfz._ke, d->Gfz.s1b, d->Gfz.s2b)] = 0;
    }
  }
}

/* Particle cages: cuda_build_cages, src/cuda_particle.cu:1523-1598, with cage_setup
 * (particle_kernel.cu:135-146), build_phase (:148-253), build_phase_shell (:255-426) and
 * cage_flag_{u,v,w} (:482-540).  Particle data are per-block arrays of centres/radii in
 * global coordinates (the reference keeps a per-rank `_parts` list; a particle is listed on
 * every block whose cage it can touch -- here simply every particle on every block, which
 * gives the same cells because the cage is clipped to the block).  Followed by the external
 * wall flags exactly as in the NPARTS == 0 case (:1600-1639). */
static void cage_range(const bbo_state *s, const dom_struct *d, real px, real py, real pz, real pr,
                       int lo[3], int hi[3], int *empty)
{
  int cage[3], a;
  int S[3], E[3];
  real idx = 1. / d->dx, idy = 1. / d->dy, idz = 1. / d->dz;
  cage[0] = (int)(2. * ceil(pr / d->dx)) + 2 - (d->xn % 2);
  cage[1] = (int)(2. * ceil(pr / d->dy)) + 2 - (d->yn % 2);
  cage[2] = (int)(2. * ceil(pr / d->dz)) + 2 - (d->zn % 2);
  lo[0] = (int)(round((px - d->xs) * idx) - 0.5 * cage[0] + DOM_BUF);
  lo[1] = (int)(round((py - d->ys) * idy) - 0.5 * cage[1] + DOM_BUF);
  lo[2] = (int)(round((pz - d->zs) * idz) - 0.5 * cage[2] + DOM_BUF);
  for (a = 0; a < 3; a++) hi[a] = lo[a] + cage[a];
  S[0] = (d->I == s->DOM.Is && s->bc.pW != BB_PERIODIC) ? d->Gcc._is : d->Gcc._isb;
  E[0] = (d->I == s->DOM.Ie && s->bc.pE != BB_PERIODIC) ? d->Gcc._ie : d->Gcc._ieb;
  S[1] = (d->J == s->DOM.Js && s->bc.pS != BB_PERIODIC) ? d->Gcc._js : d->Gcc._jsb;
  E[1] = (d->J == s->DOM.Je && s->bc.pN != BB_PERIODIC) ? d->Gcc._je : d->Gcc._jeb;
  S[2] = (d->K == s->DOM.Ks && s->bc.pB != BB_PERIODIC) ? d->Gcc._ks : d->Gcc._ksb;
  E[2] = (d->K == s->DOM.Ke && s->bc.pT != BB_PERIODIC) ? d->Gcc._ke : d->Gcc._keb;
  *empty = 0;
  for (a = 0; a < 3; a++) {
    if (lo[a] < S[a]) lo[a] = S[a]; else if (lo[a] > E[a]) lo[a] = E[a];
    if (hi[a] < S[a]) hi[a] = S[a]; else if (hi[a] > E[a]) hi[a] = E[a];
    if (lo[a] == hi[a]) *empty = 1;        /* `is != ie` guard, particle_kernel.cu:230-232 */
  }
}

void bbo_build_cages(bbo_state *s, int nparts, const real *px, const real *py, const real *pz, const real *pr)
{
  int c, n, i, j, k;
  for (c = 0; c < s->nblocks; c++) {
    const dom_struct *d = &s->dom[c];
    bbo_block *b = &s->blk[c];
    for (i = 0; i < d->Gfx.s3b; i++) b->flag_u[i] = 1;
    for (i = 0; i < d->Gfy.s3b; i++) b->flag_v[i] = 1;
    for (i = 0; i < d->Gfz.s3b; i++) b->flag_w[i] = 1;
    for (i = 0; i < d->Gcc.s3b; i++) { b->phase[i] = -1; b->phase_shell[i] = 1; }   /* :121-133 */
    for (n = 0; n < nparts; n++) {                                       /* build_phase */
      int lo[3], hi[3], empty;
      real irad = 1. / pr[n];
      cage_range(s, d, px[n], py[n], pz[n], pr[n], lo, hi, &empty);
      if (empty) continue;
      for (k = lo[2]; k <= hi[2]; k++) for (j = lo[1]; j <= hi[1]; j++) for (i = lo[0]; i <= hi[0]; i++) {
        real xx = (i - 0.5) * d->dx - (px[n] - d->xs);
        real yy = (j - 0.5) * d->dy - (py[n] - d->ys);
        real zz = (k - 0.5) * d->dz - (pz[n] - d->zs);
        real dd = sqrt(xx * xx + yy * yy + zz * zz);
        int C = GCC_LOC(i, j, k, d->Gcc.s1b, d->Gcc.s2b);
        int cutoff = floor(dd * irad) < 1;
        b->phase[C] += cutoff * (n - b->phase[C]);
      }
    }
    for (n = 0; n < nparts; n++) {                                       /* build_phase_shell */
      int lo[3], hi[3], empty;
      real irad = 1. / pr[n];
      cage_range(s, d, px[n], py[n], pz[n], pr[n], lo, hi, &empty);
      if (empty) continue;
      for (k = lo[2]; k <= hi[2]; k++) for (j = lo[1]; j <= hi[1]; j++) for (i = lo[0]; i <= hi[0]; i++) {
        int C = GCC_LOC(i, j, k, d->Gcc.s1b, d->Gcc.s2b);
        real xw = (i - 1 - 0.5) * d->dx - (px[n] - d->xs), xx = (i - 0.5) * d->dx - (px[n] - d->xs),
             xe = (i + 1 - 0.5) * d->dx - (px[n] - d->xs);
        real ys = (j - 1 - 0.5) * d->dy - (py[n] - d->ys), yy = (j - 0.5) * d->dy - (py[n] - d->ys),
             yn = (j + 1 - 0.5) * d->dy - (py[n] - d->ys);
        real zb = (k - 1 - 0.5) * d->dz - (pz[n] - d->zs), zz = (k - 0.5) * d->dz - (pz[n] - d->zs),
             zt = (k + 1 - 0.5) * d->dz - (pz[n] - d->zs);
        int pw = floor(sqrt(xw * xw + yy * yy + zz * zz) * irad) < 1;
        int pe = floor(sqrt(xe * xe + yy * yy + zz * zz) * irad) < 1;
        int ps = floor(sqrt(xx * xx + ys * ys + zz * zz) * irad) < 1;
        int pn = floor(sqrt(xx * xx + yn * yn + zz * zz) * irad) < 1;
        int pb = floor(sqrt(xx * xx + yy * yy + zb * zb) * irad) < 1;
        int pt = floor(sqrt(xx * xx + yy * yy + zt * zt) * irad) < 1;
        b->phase_shell[C] *= 1 - ((b->phase[C] == n) && (pw == 0 || pe == 0 || ps == 0 || pn == 0 || pb == 0 || pt == 0));
      }
    }
    /* cage_flag_u/v/w, particle_kernel.cu:482-540 (faces _is.._ie only) */
    for (k = 0; k < d->Gfx.knb; k++) for (j = 0; j < d->Gfx.jnb; j++) for (i = d->Gfx._is; i <= d->Gfx._ie; i++) {
      int CE = GCC_LOC(i, j, k, d->Gcc.s1b, d->Gcc.s2b), CW = GCC_LOC(i - 1, j, k, d->Gcc.s1b, d->Gcc.s2b);
      b->flag_u[GFX_LOC(i, j, k, d->Gfx.s1b, d->Gfx.s2b)] =
        1 - 2 * ((b->phase[CW] < 0 && b->phase[CE] > -1) || (b->phase[CW] > -1 && b->phase[CE] < 0) ||
                 (b->phase_shell[CE] < 1 && b->phase_shell[CW] < 1));
    }
    for (k = 0; k < d->Gfy.knb; k++) for (i = 0; i < d->Gfy.inb; i++) for (j = d->Gfy._js; j <= d->Gfy._je; j++) {
      int CN = GCC_LOC(i, j, k, d->Gcc.s1b, d->Gcc.s2b), CS = GCC_LOC(i, j - 1, k, d->Gcc.s1b, d->Gcc.s2b);
      b->flag_v[GFY_LOC(i, j, k, d->Gfy.s1b, d->Gfy.s2b)] =
        1 - 2 * ((b->phase[CS] < 0 && b->phase[CN] > -1) || (b->phase[CS] > -1 && b->phase[CN] < 0) ||
                 (b->phase_shell[CN] < 1 && b->phase_shell[CS] < 1));
    }
    for (j = 0; j < d->Gfz.jnb; j++) for (i = 0; i < d->Gfz.inb; i++) for (k = d->Gfz._ks; k <= d->Gfz._ke; k++) {
      int CT = GCC_LOC(i, j, k, d->Gcc.s1b, d->Gcc.s2b), CB = GCC_LOC(i, j, k - 1, d->Gcc.s1b, d->Gcc.s2b);
      b->flag_w[GFZ_LOC(i, j, k, d->Gfz.s1b, d->Gfz.s2b)] =
        1 - 2 * ((b->phase[CB] < 0 && b->phase[CT] > -1) || (b->phase[CB] > -1 && b->phase[CT] < 0) ||
                 (b->phase_shell[CT] < 1 && b->phase_shell[CB] < 1));
    }
  }
  /* external walls on top, cuda_particle.cu:1600-1639 */
  for (c = 0; c < s->nblocks; c++) {
    const dom_struct *d = &s->dom[c];
    bbo_block *b = &s->blk[c];
    if (s->bc.pW != BB_PERIODIC && s->bc.pE != BB_PERIODIC) {
      if (d->I == s->DOM.Is) for (k = 0; k < d->Gfx.knb; k++) for (j = 0; j < d->Gfx.jnb; j++)
        b->flag_u[GFX_LOC(d->Gfx._is, j, k, d->Gfx.s1b, d->Gfx.s2b)] = 0;
      if (d->I == s->DOM.Ie) for (k = 0; k < d->Gfx.knb; k++) for (j = 0; j < d->Gfx.jnb; j++)
        b->flag_u[GFX_LOC(d->Gfx._ie, j, k, d->Gfx.s1b, d->Gfx.s2b)] = 0;
    }
    if (s->bc.pS != BB_PERIODIC && s->bc.pN != BB_PERIODIC) {
      if (d->J == s->DOM.Js) for (k = 0; k < d->Gfy.knb; k++) for (i = 0; i < d->Gfy.inb; i++)
        b->flag_v[GFY_LOC(i, d->Gfy._js, k, d->Gfy.s1b, d->Gfy.s2b)] = 0;
      if (d->J == s->DOM.Je) for (k = 0; k < d->Gfy.knb; k++) for (i = 0; i < d->Gfy.inb; i++)
        b->flag_v[GFY_LOC(i, d->Gfy._je, k, d->Gfy.s1b, d->Gfy.s2b)] = 0;
    }
    if (s->bc.pB != BB_PERIODIC && s->bc.pT != BB_PERIODIC) {
      if (d->K == s->DOM.Ks) for (j = 0; j < d->Gfz.jnb; j++) for (i = 0; i < d->Gfz.inb; i++)
        b->flag_w[GFZ_LOC(i, j, d->Gfz._ks, d->Gfz.s1b, d->Gfz.s2b)] = 0;
      if (d->K == s->DOM.Ke) for (j = 0; j < d->Gfz.jnb; j++) for (i = 0; i < d->Gfz.inb; i++)
        b->flag_w[GFZ_LOC(i, j, d->Gfz._ke, d->Gfz.s1b, d->Gfz.s2b)] = 0;
    }
  }
}

/* ------------------------------------------------------------------------------------ */
/* PP_jacobi_init, src/solver_kernel.cu:26-87 (JACOBI branch): invM = -1/M,
 * M = -idx2(fE^2+fW^2) - idy2(fN^2+fS^2) - idz2(fT^2+fB^2). */
void bbo_jacobi_init(bbo_state *s)
{
  int c;
  for (c = 0; c < s->nblocks; c++) {
    const dom_struct *d = &s->dom[c];
    bbo_block *b = &s->blk[c];
    real idx2 = 1. / (d->dx * d->dx), idy2 = 1. / (d->dy * d->dy), idz2 = 1. / (d->dz * d->dz);
    int tk;
#pragma omp parallel for schedule(static)
    for (tk = 0; tk < d->Gcc.kn; tk++) {
      int ti, tj;
      for (tj = 0; tj < d->Gcc.jn; tj++) for (ti = 0; ti < d->Gcc.in; ti++) {
        int cc = GCC_LOC(ti, tj, tk, d->Gcc.s1, d->Gcc.s2);
        int TI = ti + DOM_BUF, TJ = tj + DOM_BUF, TK = tk + DOM_BUF;
        int Wfx = GFX_LOC(TI, TJ, TK, d->Gfx.s1b, d->Gfx.s2b), Efx = GFX_LOC(TI + 1, TJ, TK, d->Gfx.s1b, d->Gfx.s2b);
        int Sfy = GFY_LOC(TI, TJ, TK, d->Gfy.s1b, d->Gfy.s2b), Nfy = GFY_LOC(TI, TJ + 1, TK, d->Gfy.s1b, d->Gfy.s2b);
        int Bfz = GFZ_LOC(TI, TJ, TK, d->Gfz.s1b, d->Gfz.s2b), Tfz = GFZ_LOC(TI, TJ, TK + 1, d->Gfz.s1b, d->Gfz.s2b);
        real M = -idx2 * (b->flag_u[Efx] * b->flag_u[Efx] + b->flag_u[Wfx] * b->flag_u[Wfx])
                 - idy2 * (b->flag_v[Nfy] * b->flag_v[Nfy] + b->flag_v[Sfy] * b->flag_v[Sfy])
                 - idz2 * (b->flag_w[Tfz] * b->flag_w[Tfz] + b->flag_w[Bfz] * b->flag_w[Bfz]);
        b->invM[cc] = -1. / M;
      }
    }
  }
}

/* cudaMemset(_rhs_p,0) + PP_rhs, src/cuda_solver.cu:122-126 / solver_kernel.cu:89-176:
 * rhs = -( ((uE-uW)*idx + (vN-vS)*idy) + (wT-wB)*idz ) * (rho_f/dt), interior only. */
void bbo_rhs(bbo_state *s, real rho_f, real dt)
{
  int c;
  for (c = 0; c < s->nblocks; c++) {
    const dom_struct *d = &s->dom[c];
    bbo_block *b = &s->blk[c];
    real idx = 1. / d->dx, idy = 1. / d->dy, idz = 1. / d->dz, rho_idt = rho_f / dt;
    int TK;
    memset(b->rhs_p, 0, sizeof(real) * (size_t)d->Gcc.s3b);
#pragma omp parallel for schedule(static)
    for (TK = d->Gcc._ks; TK <= d->Gcc._ke; TK++) {
      int TI, TJ;
      for (TJ = d->Gcc._js; TJ <= d->Gcc._je; TJ++) for (TI = d->Gcc._is; TI <= d->Gcc._ie; TI++) {
        real uW = b->u_star[GFX_LOC(TI, TJ, TK, d->Gfx.s1b, d->Gfx.s2b)], uE = b->u_star[GFX_LOC(TI + 1, TJ, TK, d->Gfx.s1b, d->Gfx.s2b)];
        real vS = b->v_star[GFY_LOC(TI, TJ, TK, d->Gfy.s1b, d->Gfy.s2b)], vN = b->v_star[GFY_LOC(TI, TJ + 1, TK, d->Gfy.s1b, d->Gfy.s2b)];
        real wB = b->w_star[GFZ_LOC(TI, TJ, TK, d->Gfz.s1b, d->Gfz.s2b)], wT = b->w_star[GFZ_LOC(TI, TJ, TK + 1, d->Gfz.s1b, d->Gfz.s2b)];
        real t = (uE - uW) * idx;
        t += (vN - vS) * idy;
        t += (wT - wB) * idz;
        t *= rho_idt;
        b->rhs_p[GCC_LOC(TI, TJ, TK, d->Gcc.s1b, d->Gcc.s2b)] = -t;
      }
    }
  }
}

/* ------------------------------------------------------------------------------------ */
/* mpi_cuda_exchange_Gcc, src/mpi_comm.c:257-315: pack (bluebottle_kernel.cu:684-780) ->
 * MPI_Put into the neighbour's recv buffer (w->recv_e, e->recv_w, s->recv_n, n->recv_s,
 * b->recv_t, t->recv_b; PROC_NULL puts are no-ops) -> unpack into ghosts (:1083-1177).
 * Buffer layouts: E/W pp=(j-1)+jn(k-1); N/S pp=(k-1)+kn(i-1); T/B pp=(i-1)+in(j-1). */
static real *blk_gcc_array(bbo_block *b, int id)
{
  switch (id) { case BBO_RHS_P: return b->rhs_p; case BBO_PHI: return b->phi; case BBO_PB_Q: return b->pb_q;
                case BBO_P0: return b->p0; case BBO_P: return b->p; }
  return NULL;
}

void bbo_exchange_Gcc(bbo_state *s, int array_id)
{
  int c, i, j, k;
  for (c = 0; c < s->nblocks; c++) {                       /* pack */
    const dom_struct *d = &s->dom[c];
    const grid_info *g = &d->Gcc;
    bbo_block *b = &s->blk[c];
    real *a = blk_gcc_array(b, array_id);
    if (d->e >= 0) for (k = 1; k <= g->_ke; k++) for (j = 1; j <= g->_je; j++) b->send[BBO_E][(j - 1) + g->jn * (k - 1)] = a[GCC_LOC(g->_ie, j, k, g->s1b, g->s2b)];
    if (d->w >= 0) for (k = 1; k <= g->_ke; k++) for (j = 1; j <= g->_je; j++) b->send[BBO_W][(j - 1) + g->jn * (k - 1)] = a[GCC_LOC(g->_is, j, k, g->s1b, g->s2b)];
    if (d->n >= 0) for (i = 1; i <= g->_ie; i++) for (k = 1; k <= g->_ke; k++) b->send[BBO_N][(k - 1) + g->kn * (i - 1)] = a[GCC_LOC(i, g->_je, k, g->s1b, g->s2b)];
    if (d->s >= 0) for (i = 1; i <= g->_ie; i++) for (k = 1; k <= g->_ke; k++) b->send[BBO_S][(k - 1) + g->kn * (i - 1)] = a[GCC_LOC(i, g->_js, k, g->s1b, g->s2b)];
    if (d->t >= 0) for (j = 1; j <= g->_je; j++) for (i = 1; i <= g->_ie; i++) b->send[BBO_T][(i - 1) + g->in * (j - 1)] = a[GCC_LOC(i, j, g->_ke, g->s1b, g->s2b)];
    if (d->b >= 0) for (j = 1; j <= g->_je; j++) for (i = 1; i <= g->_ie; i++) b->send[BBO_B][(i - 1) + g->in * (j - 1)] = a[GCC_LOC(i, j, g->_ks, g->s1b, g->s2b)];
  }
  for (c = 0; c < s->nblocks; c++) {                       /* put: mpi_comm.c:293-306 */
    const dom_struct *d = &s->dom[c];
    bbo_block *b = &s->blk[c];
    if (d->w >= 0) memcpy(s->blk[d->w].recv[BBO_E], b->send[BBO_W], sizeof(real) * d->Gcc.s2_i);
    if (d->e >= 0) memcpy(s->blk[d->e].recv[BBO_W], b->send[BBO_E], sizeof(real) * d->Gcc.s2_i);
    if (d->s >= 0) memcpy(s->blk[d->s].recv[BBO_N], b->send[BBO_S], sizeof(real) * d->Gcc.s2_j);
    if (d->n >= 0) memcpy(s->blk[d->n].recv[BBO_S], b->send[BBO_N], sizeof(real) * d->Gcc.s2_j);
    if (d->b >= 0) memcpy(s->blk[d->b].recv[BBO_T], b->send[BBO_B], sizeof(real) * d->Gcc.s2_k);
    if (d->t >= 0) memcpy(s->blk[d->t].recv[BBO_B], b->send[BBO_T], sizeof(real) * d->Gcc.s2_k);
  }
  for (c = 0; c < s->nblocks; c++) {                       /* unpack, cuda_bluebottle.cu:1653-1675 */
    const dom_struct *d = &s->dom[c];
    const grid_info *g = &d->Gcc;
    bbo_block *b = &s->blk[c];
    real *a = blk_gcc_array(b, array_id);
    if (d->e >= 0) for (k = 1; k <= g->_ke; k++) for (j = 1; j <= g->_je; j++) a[GCC_LOC(g->_ieb, j, k, g->s1b, g->s2b)] = b->recv[BBO_E][(j - 1) + g->jn * (k - 1)];
    if (d->w >= 0) for (k = 1; k <= g->_ke; k++) for (j = 1; j <= g->_je; j++) a[GCC_LOC(g->_isb, j, k, g->s1b, g->s2b)] = b->recv[BBO_W][(j - 1) + g->jn * (k - 1)];
    if (d->n >= 0) for (i = 1; i <= g->_ie; i++) for (k = 1; k <= g->_ke; k++) a[GCC_LOC(i, g->_jeb, k, g->s1b, g->s2b)] = b->recv[BBO_N][(k - 1) + g->kn * (i - 1)];
    if (d->s >= 0) for (i = 1; i <= g->_ie; i++) for (k = 1; k <= g->_ke; k++) a[GCC_LOC(i, g->_jsb, k, g->s1b, g->s2b)] = b->recv[BBO_S][(k - 1) + g->kn * (i - 1)];
    if (d->t >= 0) for (j = 1; j <= g->_je; j++) for (i = 1; i <= g->_ie; i++) a[GCC_LOC(i, j, g->_keb, g->s1b, g->s2b)] = b->recv[BBO_T][(i - 1) + g->in * (j - 1)];
    if (d->b >= 0) for (j = 1; j <= g->_je; j++) for (i = 1; i <= g->_ie; i++) a[GCC_LOC(i, j, g->_ksb, g->s1b, g->s2b)] = b->recv[BBO_B][(i - 1) + g->in * (j - 1)];
  }
}

/* ------------------------------------------------------------------------------------ */
/* mpi_cuda_exchange_Gfx / _Gfy / _Gfz, src/mpi_comm.c:317-405, pack/unpack kernels src/bluebottle_kernel.cu:782-1081,
 * 1179-1466.  Same scheme as the Gcc exchange with the grid's own extents and index macro; along the grid's own normal
 * the block-boundary face is shared with the neighbour, so the planes _ie-1 / _is+1 are sent (:793,:811 for Gfx,
 * :927,:944 for Gfy, :1059,:1077 for Gfz) into the ghosts _isb / _ieb. */
static real *blk_face_array(bbo_block *b, int id, int *grid)
{
  switch (id) {
    case BBO_U_STAR: *grid = 1; return b->u_star;  case BBO_VEL_U: *grid = 1; return b->u;
    case BBO_V_STAR: *grid = 2; return b->v_star;  case BBO_VEL_V: *grid = 2; return b->v;
    case BBO_W_STAR: *grid = 3; return b->w_star;  case BBO_VEL_W: *grid = 3; return b->w;
  }
  *grid = 0;
  return blk_gcc_array(b, id);
}

static size_t grid_loc(int grid, const grid_info *g, int i, int j, int k)
{
  switch (grid) {
    case 1: return (size_t)GFX_LOC(i, j, k, g->s1b, g->s2b);
    case 2: return (size_t)GFY_LOC(i, j, k, g->s1b, g->s2b);
    case 3: return (size_t)GFZ_LOC(i, j, k, g->s1b, g->s2b);
  }
  return (size_t)GCC_LOC(i, j, k, g->s1b, g->s2b);
}

void bbo_exchange(bbo_state *s, int array_id)
{
  int c, i, j, k, grid = 0, pass;
  for (pass = 0; pass < 3; pass++) {                       /* 0 pack, 1 put, 2 unpack */
    for (c = 0; c < s->nblocks; c++) {
      const dom_struct *d = &s->dom[c];
      bbo_block *b = &s->blk[c];
      real *a = blk_face_array(b, array_id, &grid);
      const grid_info *g = grid == 1 ? &d->Gfx : grid == 2 ? &d->Gfy : grid == 3 ? &d->Gfz : &d->Gcc;
      /* planes sent east/west, north/south, top/bottom */
      const int ie = g->_ie - (grid == 1), is = g->_is + (grid == 1);
      const int je = g->_je - (grid == 2), js = g->_js + (grid == 2);
      const int ke = g->_ke - (grid == 3), ks = g->_ks + (grid == 3);
      if (pass == 0) {
        if (d->e >= 0) for (k = 1; k <= g->_ke; k++) for (j = 1; j <= g->_je; j++) b->send[BBO_E][(j - 1) + g->jn * (k - 1)] = a[grid_loc(grid, g, ie, j, k)];
        if (d->w >= 0) for (k = 1; k <= g->_ke; k++) for (j = 1; j <= g->_je; j++) b->send[BBO_W][(j - 1) + g->jn * (k - 1)] = a[grid_loc(grid, g, is, j, k)];
        if (d->n >= 0) for (i = 1; i <= g->_ie; i++) for (k = 1; k <= g->_ke; k++) b->send[BBO_N][(k - 1) + g->kn * (i - 1)] = a[grid_loc(grid, g, i, je, k)];
        if (d->s >= 0) for (i = 1; i <= g->_ie; i++) for (k = 1; k <= g->_ke; k++) b->send[BBO_S][(k - 1) + g->kn * (i - 1)] = a[grid_loc(grid, g, i, js, k)];
        if (d->t >= 0) for (j = 1; j <= g->_je; j++) for (i = 1; i <= g->_ie; i++) b->send[BBO_T][(i - 1) + g->in * (j - 1)] = a[grid_loc(grid, g, i, j, ke)];
        if (d->b >= 0) for (j = 1; j <= g->_je; j++) for (i = 1; i <= g->_ie; i++) b->send[BBO_B][(i - 1) + g->in * (j - 1)] = a[grid_loc(grid, g, i, j, ks)];
      } else if (pass == 1) {                              /* mpi_comm.c:326-343 */
        if (d->w >= 0) memcpy(s->blk[d->w].recv[BBO_E], b->send[BBO_W], sizeof(real) * g->s2_i);
        if (d->e >= 0) memcpy(s->blk[d->e].recv[BBO_W], b->send[BBO_E], sizeof(real) * g->s2_i);
        if (d->s >= 0) memcpy(s->blk[d->s].recv[BBO_N], b->send[BBO_S], sizeof(real) * g->s2_j);
        if (d->n >= 0) memcpy(s->blk[d->n].recv[BBO_S], b->send[BBO_N], sizeof(real) * g->s2_j);
        if (d->b >= 0) memcpy(s->blk[d->b].recv[BBO_T], b->send[BBO_B], sizeof(real) * g->s2_k);
        if (d->t >= 0) memcpy(s->blk[d->t].recv[BBO_B], b->send[BBO_T], sizeof(real) * g->s2_k);
      } else {
        if (d->e >= 0) for (k = 1; k <= g->_ke; k++) for (j = 1; j <= g->_je; j++) a[grid_loc(grid, g, g->_ieb, j, k)] = b->recv[BBO_E][(j - 1) + g->jn * (k - 1)];
        if (d->w >= 0) for (k = 1; k <= g->_ke; k++) for (j = 1; j <= g->_je; j++) a[grid_loc(grid, g, g->_isb, j, k)] = b->recv[BBO_W][(j - 1) + g->jn * (k - 1)];
        if (d->n >= 0) for (i = 1; i <= g->_ie; i++) for (k = 1; k <= g->_ke; k++) a[grid_loc(grid, g, i, g->_jeb, k)] = b->recv[BBO_N][(k - 1) + g->kn * (i - 1)];
        if (d->s >= 0) for (i = 1; i <= g->_ie; i++) for (k = 1; k <= g->_ke; k++) a[grid_loc(grid, g, i, g->_jsb, k)] = b->recv[BBO_S][(k - 1) + g->kn * (i - 1)];
        if (d->t >= 0) for (j = 1; j <= g->_je; j++) for (i = 1; i <= g->_ie; i++) a[grid_loc(grid, g, i, j, g->_keb)] = b->recv[BBO_T][(i - 1) + g->in * (j - 1)];
        if (d->b >= 0) for (j = 1; j <= g->_je; j++) for (i = 1; i <= g->_ie; i++) a[grid_loc(grid, g, i, j, g->_ksb)] = b->recv[BBO_B][(i - 1) + g->in * (j - 1)];
      }
    }
  }
}

/* ------------------------------------------------------------------------------------ */
/* cuda_solvability, src/cuda_bluebottle.cu:2313-2492: per-rank face integrals of u_star on the six faces of the GLOBAL
 * domain (surf_int_*, src/bluebottle_kernel.cu:2135-2229, thrust::reduce, times the face area), eps[axis] = end - start
 * summed over ranks (:2414-2420), then the correction of the outflow plane(s) (:2423-2491, plane_eps_* :2231-2301).
 * Summation order is ours (rows, then ranks in rank order). */
static real plane_sum(const real *a, int grid, const grid_info *g, int c)
{
  real tot = 0.;
  int p, q;
  if (grid == 1) { for (q = g->_ks; q <= g->_ke; q++) { real r = 0.; for (p = g->_js; p <= g->_je; p++) r += a[GFX_LOC(c, p, q, g->s1b, g->s2b)]; tot += r; } }
  else if (grid == 2) { for (q = g->_is; q <= g->_ie; q++) { real r = 0.; for (p = g->_ks; p <= g->_ke; p++) r += a[GFY_LOC(q, c, p, g->s1b, g->s2b)]; tot += r; } }
  else { for (q = g->_js; q <= g->_je; q++) { real r = 0.; for (p = g->_is; p <= g->_ie; p++) r += a[GFZ_LOC(p, q, c, g->s1b, g->s2b)]; tot += r; } }
  return tot;
}

static void plane_add(real *a, int grid, const grid_info *g, int c, real val)
{
  int p, q;
  if (grid == 1) { for (q = g->_ks; q <= g->_ke; q++) for (p = g->_js; p <= g->_je; p++) { size_t C = GFX_LOC(c, p, q, g->s1b, g->s2b); a[C] = a[C] + val; } }
  else if (grid == 2) { for (q = g->_is; q <= g->_ie; q++) for (p = g->_ks; p <= g->_ke; p++) { size_t C = GFY_LOC(q, c, p, g->s1b, g->s2b); a[C] = a[C] + val; } }
  else { for (q = g->_js; q <= g->_je; q++) for (p = g->_is; p <= g->_ie; p++) { size_t C = GFZ_LOC(p, q, c, g->s1b, g->s2b); a[C] = a[C] + val; } }
}

void bbo_solvability(bbo_state *s, int out_plane, real eps[3])
{
  const dom_struct *D = &s->DOM;
  int c;
  eps[0] = eps[1] = eps[2] = 0.;
  for (c = 0; c < s->nblocks; c++) {
    const dom_struct *d = &s->dom[c];
    bbo_block *b = &s->blk[c];
    real xs = 0., xe = 0., ys = 0., ye = 0., zs = 0., ze = 0.;
    if (d->I == 0)         { xs = plane_sum(b->u_star, 1, &d->Gfx, d->Gfx._is); xs *= d->dy * d->dz; }
    if (d->I == D->In - 1) { xe = plane_sum(b->u_star, 1, &d->Gfx, d->Gfx._ie); xe *= d->dy * d->dz; }
    if (d->J == 0)         { ys = plane_sum(b->v_star, 2, &d->Gfy, d->Gfy._js); ys *= d->dz * d->dx; }
    if (d->J == D->Jn - 1) { ye = plane_sum(b->v_star, 2, &d->Gfy, d->Gfy._je); ye *= d->dz * d->dx; }
    if (d->K == 0)         { zs = plane_sum(b->w_star, 3, &d->Gfz, d->Gfz._ks); zs *= d->dx * d->dy; }
    if (d->K == D->Kn - 1) { ze = plane_sum(b->w_star, 3, &d->Gfz, d->Gfz._ke); ze *= d->dx * d->dy; }
    eps[0] += xe - xs; eps[1] += ye - ys; eps[2] += ze - zs;
  }
  for (c = 0; c < s->nblocks; c++) {
    const dom_struct *d = &s->dom[c];
    bbo_block *b = &s->blk[c];
    const real all = eps[0] + eps[1] + eps[2];
    if (out_plane == 10) {                                   /* HOMOGENEOUS, :2469-2488 */
      const real sx = 0.5 * eps[0] / (D->yl * D->zl), sy = 0.5 * eps[1] / (D->zl * D->xl), sz = 0.5 * eps[2] / (D->xl * D->yl);
      if (d->I == 0) plane_add(b->u_star, 1, &d->Gfx, d->Gfx._is, sx);
      if (d->I == D->In - 1) plane_add(b->u_star, 1, &d->Gfx, d->Gfx._ie, -sx);
      if (d->J == 0) plane_add(b->v_star, 2, &d->Gfy, d->Gfy._js, sy);
      if (d->J == D->Jn - 1) plane_add(b->v_star, 2, &d->Gfy, d->Gfy._je, -sy);
      if (d->K == 0) plane_add(b->w_star, 3, &d->Gfz, d->Gfz._ks, sz);
      if (d->K == D->Kn - 1) plane_add(b->w_star, 3, &d->Gfz, d->Gfz._ke, -sz);
    } else if (out_plane == 0 && d->I == 0) plane_add(b->u_star, 1, &d->Gfx, d->Gfx._is, all / (D->yl * D->zl));
    else if (out_plane == 1 && d->I == D->In - 1) plane_add(b->u_star, 1, &d->Gfx, d->Gfx._ie, -(all / (D->yl * D->zl)));
    else if (out_plane == 2 && d->J == 0) plane_add(b->v_star, 2, &d->Gfy, d->Gfy._js, all / (D->zl * D->xl));
    else if (out_plane == 3 && d->J == D->Jn - 1) plane_add(b->v_star, 2, &d->Gfy, d->Gfy._je, -(all / (D->zl * D->xl)));
    else if (out_plane == 4 && d->K == 0) plane_add(b->w_star, 3, &d->Gfz, d->Gfz._ks, all / (D->xl * D->yl));
    else if (out_plane == 5 && d->K == D->Kn - 1) plane_add(b->w_star, 3, &d->Gfz, d->Gfz._ke, -(all / (D->xl * D->yl)));
  }
}

/* ------------------------------------------------------------------------------------ */
/* cuda_dom_BC_star, src/cuda_bluebottle.cu:2111-2311: on every side of a block that has no neighbour (MPI_PROC_NULL),
 * faces in the order W, E, S, N, B, T and components u, v, w inside a face, a switch on the velocity BC type launches
 * BC_{u,v,w}_{face}_D(array, value) or BC_{u,v,w}_{face}_N(array) (src/bluebottle_kernel.c:104-598); PERIODIC and
 * PRECURSOR fall through.  The later faces read what the earlier ones wrote, so the order is kept.
 *   DIRICHLET, wall-normal component (BC_u_W_D :104-117, BC_u_E_D :119-132, BC_v_S_D :300-313, BC_v_N_D :315-328,
 *     BC_w_B_D :492-505, BC_w_T_D :507-520):   ghost = 2.*bc - a[one face inside the wall face];  a[wall face] = bc
 *   DIRICHLET, tangential (e.g. BC_u_N_D :134-147, BC_v_W_D :270-283):
 *     ghost = 8./3.*bc - 2.*a[first interior] + 1./3.*a[second interior]
 *   NEUMANN (e.g. BC_u_W_N :192-203):          ghost = a[first: the wall face for the normal component]
 * Index ranges: the grid's own in / jn / kn of the two tangential directions (1..n), no edges.
 * type / val: 18 entries, component-major: u on W,E,S,N,B,T, then v, then w. */
static size_t comp_loc(int comp, const grid_info *g, int i, int j, int k)
{
  return comp == 0 ? (size_t)GFX_LOC(i, j, k, g->s1b, g->s2b) : comp == 1 ? (size_t)GFY_LOC(i, j, k, g->s1b, g->s2b)
                                                                            : (size_t)GFZ_LOC(i, j, k, g->s1b, g->s2b);
}

static void bc_star_face(real *a, int comp, const grid_info *g, int face, int type, real bc)
{
  const int axis = face / 2, high = face & 1;
  const int n[3] = { g->in, g->jn, g->kn };
  const int t1 = (axis + 1) % 3, t2 = (axis + 2) % 3;
  /* along the normal: _s = 1, _e = n, _sb = 0, _eb = n + 1 (src/domain.c:1262-1478) */
  const int gh = high ? n[axis] + 1 : 0, first = high ? n[axis] : 1, second = high ? n[axis] - 1 : 2;
  int p, q;
  if (type != BB_DIRICHLET && type != BB_NEUMANN) return;
  for (q = 1; q <= n[t2]; q++) for (p = 1; p <= n[t1]; p++) {
    int c[3], cg[3], c1[3], c2[3];
    c[t1] = p; c[t2] = q;
    memcpy(cg, c, sizeof(c)); memcpy(c1, c, sizeof(c)); memcpy(c2, c, sizeof(c));
    cg[axis] = gh; c1[axis] = first; c2[axis] = second;
    {
      const size_t G = comp_loc(comp, g, cg[0], cg[1], cg[2]), F = comp_loc(comp, g, c1[0], c1[1], c1[2]),
                   S = comp_loc(comp, g, c2[0], c2[1], c2[2]);
      if (type == BB_NEUMANN) a[G] = a[F];
      else if (comp == axis) { a[G] = 2. * bc - a[S]; a[F] = bc; }
      else a[G] = 8. / 3. * bc - 2. * a[F] + 1. / 3. * a[S];
    }
  }
}

void bbo_dom_BC_star(bbo_state *s, const int *type, const real *val)
{
  int c, face, comp;
  for (c = 0; c < s->nblocks; c++) {
    const dom_struct *d = &s->dom[c];
    bbo_block *b = &s->blk[c];
    const int nbr[6] = { d->w, d->e, d->s, d->n, d->b, d->t };
    real *arr[3] = { b->u_star, b->v_star, b->w_star };
    const grid_info *g[3] = { &d->Gfx, &d->Gfy, &d->Gfz };
    for (face = 0; face < 6; face++) {
      if (nbr[face] >= 0) continue;                        /* dom[rank].w == MPI_PROC_NULL, :2114 ... */
      for (comp = 0; comp < 3; comp++) bc_star_face(arr[comp], comp, g[comp], face, type[comp * 6 + face], val[comp * 6 + face]);
    }
  }
}

/* ------------------------------------------------------------------------------------ */
/* The solve epilogue, src/bluebottle.c:233-256.
 *
 * cuda_dom_BC_p, src/cuda_bluebottle.cu:2536-2589 with BC_p_{W,E,S,N,B,T}_N, src/bluebottle_kernel.cu:26-102:
 * on a side with no neighbour (MPI_PROC_NULL) whose pressure BC is NEUMANN, ghost = adjacent interior
 * cell, for the interior ranges of the other two indices only (faces, no edges). */
void bbo_dom_BC_p(bbo_state *s, int array_id)
{
  int c, i, j, k;
  for (c = 0; c < s->nblocks; c++) {
    const dom_struct *d = &s->dom[c];
    const grid_info *g = &d->Gcc;
    real *a = blk_gcc_array(&s->blk[c], array_id);
    if (d->w < 0 && s->bc.pW == BB_NEUMANN) for (k = 1; k <= g->kn; k++) for (j = 1; j <= g->jn; j++) a[GCC_LOC(g->_isb, j, k, g->s1b, g->s2b)] = a[GCC_LOC(g->_is, j, k, g->s1b, g->s2b)];
    if (d->e < 0 && s->bc.pE == BB_NEUMANN) for (k = 1; k <= g->kn; k++) for (j = 1; j <= g->jn; j++) a[GCC_LOC(g->_ieb, j, k, g->s1b, g->s2b)] = a[GCC_LOC(g->_ie, j, k, g->s1b, g->s2b)];
    if (d->s < 0 && s->bc.pS == BB_NEUMANN) for (k = 1; k <= g->kn; k++) for (i = 1; i <= g->in; i++) a[GCC_LOC(i, g->_jsb, k, g->s1b, g->s2b)] = a[GCC_LOC(i, g->_js, k, g->s1b, g->s2b)];
    if (d->n < 0 && s->bc.pN == BB_NEUMANN) for (k = 1; k <= g->kn; k++) for (i = 1; i <= g->in; i++) a[GCC_LOC(i, g->_jeb, k, g->s1b, g->s2b)] = a[GCC_LOC(i, g->_je, k, g->s1b, g->s2b)];
    if (d->b < 0 && s->bc.pB == BB_NEUMANN) for (j = 1; j <= g->jn; j++) for (i = 1; i <= g->in; i++) a[GCC_LOC(i, j, g->_ksb, g->s1b, g->s2b)] = a[GCC_LOC(i, j, g->_ks, g->s1b, g->s2b)];
    if (d->t < 0 && s->bc.pT == BB_NEUMANN) for (j = 1; j <= g->jn; j++) for (i = 1; i <= g->in; i++) a[GCC_LOC(i, j, g->_keb, g->s1b, g->s2b)] = a[GCC_LOC(i, j, g->_ke, g->s1b, g->s2b)];
  }
}

/* cuda_project, src/cuda_bluebottle.cu:2495-2503; project_u/v/w, src/bluebottle_kernel.cu:2303-2355:
 *   gradPhi = abs(flag) * ddx * (phi[C] - phi[W]);  u = u_star - dt / rho_f * gradPhi
 * over Gfx._is.._ie (in = xn+1 faces) x interior j,k; same for v, w.  phi ghosts as given. */
void bbo_project(bbo_state *s, real rho_f, real dt)
{
  int c;
  for (c = 0; c < s->nblocks; c++) {
    const dom_struct *d = &s->dom[c];
    bbo_block *b = &s->blk[c];
    const real ddx = 1. / d->dx, ddy = 1. / d->dy, ddz = 1. / d->dz;
    const int cs1 = d->Gcc.s1b, cs2 = d->Gcc.s2b;
    int k;
#pragma omp parallel for schedule(static)
    for (k = d->Gfx._ks; k <= d->Gfx._ke; k++) {
      int i, j;
      for (j = d->Gfx._js; j <= d->Gfx._je; j++) for (i = d->Gfx._is; i <= d->Gfx._ie; i++) {
        int cf = GFX_LOC(i, j, k, d->Gfx.s1b, d->Gfx.s2b);
        real gradPhi = abs(b->flag_u[cf]) * ddx * (b->phi[GCC_LOC(i, j, k, cs1, cs2)] - b->phi[GCC_LOC(i - 1, j, k, cs1, cs2)]);
        b->u[cf] = (b->u_star[cf] - dt / rho_f * gradPhi);
      }
    }
#pragma omp parallel for schedule(static)
    for (k = d->Gfy._ks; k <= d->Gfy._ke; k++) {
      int i, j;
      for (j = d->Gfy._js; j <= d->Gfy._je; j++) for (i = d->Gfy._is; i <= d->Gfy._ie; i++) {
        int cf = GFY_LOC(i, j, k, d->Gfy.s1b, d->Gfy.s2b);
        real gradPhi = abs(b->flag_v[cf]) * ddy * (b->phi[GCC_LOC(i, j, k, cs1, cs2)] - b->phi[GCC_LOC(i, j - 1, k, cs1, cs2)]);
        b->v[cf] = (b->v_star[cf] - dt / rho_f * gradPhi);
      }
    }
#pragma omp parallel for schedule(static)
    for (k = d->Gfz._ks; k <= d->Gfz._ke; k++) {
      int i, j;
      for (j = d->Gfz._js; j <= d->Gfz._je; j++) for (i = d->Gfz._is; i <= d->Gfz._ie; i++) {
        int cf = GFZ_LOC(i, j, k, d->Gfz.s1b, d->Gfz.s2b);
        real gradPhi = abs(b->flag_w[cf]) * ddz * (b->phi[GCC_LOC(i, j, k, cs1, cs2)] - b->phi[GCC_LOC(i, j, k - 1, cs1, cs2)]);
        b->w[cf] = (b->w_star[cf] - dt / rho_f * gradPhi);
      }
    }
  }
}

/* cuda_update_p, src/cuda_bluebottle.cu:2505-2534: update_p (src/bluebottle_kernel.cu:2385-2402; the Laplacian of
 * :2357-2383 is computed but unused, :2396 vs the commented :2399) then the mean over ALL ranks' interior cells
 * (copy_p_p_noghost + thrust::reduce + MPI_Allreduce, pmean /= DOM.Gcc.s3) is subtracted from the interior
 * (forcing_add_c_const(-pmean), :1507-1518).  Summation order is ours (rows -> planes -> ranks), see dot_s3. */
real bbo_update_p(bbo_state *s)
{
  int c, k;
  real total = 0., pmean;
  for (c = 0; c < s->nblocks; c++) {
    const grid_info *g = &s->dom[c].Gcc;
    bbo_block *b = &s->blk[c];
    real *pp = s->plane_partial, tot = 0.;
#pragma omp parallel for schedule(static)
    for (k = g->_ks; k <= g->_ke; k++) {
      real pk = 0.;
      int i, j;
      for (j = g->_js; j <= g->_je; j++) {
        real pj = 0.;
        for (i = g->_is; i <= g->_ie; i++) {
          int C = GCC_LOC(i, j, k, g->s1b, g->s2b);
          b->p[C] = (b->phase[C] < 0) * (b->p0[C] + b->phi[C]);
          pj += b->p[C];
        }
        pk += pj;
      }
      pp[k] = pk;
    }
    for (k = g->_ks; k <= g->_ke; k++) tot += pp[k];
    total += tot;
  }
  pmean = total / (real)s->DOM.Gcc.s3;
  for (c = 0; c < s->nblocks; c++) {
    const grid_info *g = &s->dom[c].Gcc;
    bbo_block *b = &s->blk[c];
#pragma omp parallel for schedule(static)
    for (k = g->_ks; k <= g->_ke; k++) {
      int i, j;
      for (j = g->_js; j <= g->_je; j++) for (i = g->_is; i <= g->_ie; i++) b->p[GCC_LOC(i, j, k, g->s1b, g->s2b)] += -pmean;
    }
  }
  return pmean;
}

/* bluebottle.c:233-256 restricted to what touches phi / p: exchange(phi), dom_BC_p(phi), project, update_p */
real bbo_epilogue(bbo_state *s, real rho_f, real dt)
{
  bbo_exchange_Gcc(s, BBO_PHI);
  bbo_dom_BC_p(s, BBO_PHI);
  bbo_project(s, rho_f, dt);
  return bbo_update_p(s);
}

/* ------------------------------------------------------------------------------------ */
/* Particle right-hand-side patch, cuda_solver.cu:128-148.
 * (1) cuda_part_BC_p -> part_BC_p (particle_kernel.cu:1655-1756): the last statement
 *     (:1753) multiplies rhs by (phase<0 && phase_shell), i.e. rhs = 0 in every solid cell,
 *     and leaves fluid cells (phase_shell == 1 there) unchanged; the Lamb-series value
 *     computed before it is discarded.  Restated as that net effect.
 * (2) exchange rhs; (3) coeffs_refine (solver_kernel.cu:212-256); (4) zero ghosts (:178-209). */
static void part_rhs_patch(bbo_state *s)
{
  int c, i, j, k;
  for (c = 0; c < s->nblocks; c++) {
    const grid_info *g = &s->dom[c].Gcc;
    bbo_block *b = &s->blk[c];
    for (k = g->_ks; k <= g->_ke; k++) for (j = g->_js; j <= g->_je; j++) for (i = g->_is; i <= g->_ie; i++) {
      int C = GCC_LOC(i, j, k, g->s1b, g->s2b);
      b->rhs_p[C] = (real)(b->phase[C] < 0 && b->phase_shell[C]) * b->rhs_p[C];
    }
  }
  bbo_exchange_Gcc(s, BBO_RHS_P);
  for (c = 0; c < s->nblocks; c++) {
    const dom_struct *d = &s->dom[c];
    const grid_info *g = &d->Gcc;
    bbo_block *b = &s->blk[c];
    real idx2 = 1. / (d->dx * d->dx), idy2 = 1. / (d->dy * d->dy), idz2 = 1. / (d->dz * d->dz);
    real *rhs = b->rhs_p; const int *ph = b->phase;
    for (k = g->_ks; k <= g->_ke; k++) for (j = g->_js; j <= g->_je; j++) for (i = g->_is; i <= g->_ie; i++) {
      int CC = GCC_LOC(i, j, k, g->s1b, g->s2b);
      int CE = CC + 1, CW = CC - 1, CN = CC + g->s1b, CS = CC - g->s1b, CT = CC + g->s2b, CB = CC - g->s2b;
      int is_fluid = (ph[CC] == -1);
      rhs[CC] += is_fluid * (ph[CE] > -1) * idx2 * (-rhs[CE]);
      rhs[CC] += is_fluid * (ph[CW] > -1) * idx2 * (-rhs[CW]);
      rhs[CC] += is_fluid * (ph[CN] > -1) * idy2 * (-rhs[CN]);
      rhs[CC] += is_fluid * (ph[CS] > -1) * idy2 * (-rhs[CS]);
      rhs[CC] += is_fluid * (ph[CT] > -1) * idz2 * (-rhs[CT]);
      rhs[CC] += is_fluid * (ph[CB] > -1) * idz2 * (-rhs[CB]);
    }
    for (k = 0; k < g->knb; k++) for (j = 0; j < g->jnb; j++) { rhs[GCC_LOC(g->_isb, j, k, g->s1b, g->s2b)] = 0.; rhs[GCC_LOC(g->_ieb, j, k, g->s1b, g->s2b)] = 0.; }
    for (k = 0; k < g->knb; k++) for (i = 0; i < g->inb; i++) { rhs[GCC_LOC(i, g->_jsb, k, g->s1b, g->s2b)] = 0.; rhs[GCC_LOC(i, g->_jeb, k, g->s1b, g->s2b)] = 0.; }
    for (j = 0; j < g->jnb; j++) for (i = 0; i < g->inb; i++) { rhs[GCC_LOC(i, j, g->_ksb, g->s1b, g->s2b)] = 0.; rhs[GCC_LOC(i, j, g->_keb, g->s1b, g->s2b)] = 0.; }
  }
}

/* ------------------------------------------------------------------------------------ */
/* Dot products.  Reference: thrust::inner_product per rank + MPI_Allreduce(SUM)
 * (cuda_solver.cu:151-152,169-170,204-205,231-232).  Thrust/CUB's summation tree is an
 * un-vendored implementation detail (CUDA toolkit), so the order here is ours: x-rows ->
 * j -> per-plane partial, planes summed in k order, ranks summed in rank order. */
static real dot_s3(bbo_state *s, int c, const real *a, const real *bvec)
{
  const grid_info *g = &s->dom[c].Gcc;
  real *pp = s->plane_partial, tot = 0.;
  int k;
#pragma omp parallel for schedule(static)
  for (k = 0; k < g->kn; k++) {
    real pk = 0.;
    int j, i;
    for (j = 0; j < g->jn; j++) {
      real pj = 0.;
      const real *ra = a + (size_t)k * g->s2 + (size_t)j * g->s1, *rb = bvec + (size_t)k * g->s2 + (size_t)j * g->s1;
      for (i = 0; i < g->in; i++) pj += ra[i] * rb[i];
      pk += pj;
    }
    pp[k] = pk;
  }
  for (k = 0; k < g->kn; k++) tot += pp[k];
  return tot;
}

static real dot_s3b(bbo_state *s, int c, const real *a, const real *bvec)
{
  const grid_info *g = &s->dom[c].Gcc;
  real *pp = s->plane_partial, tot = 0.;
  int k;
#pragma omp parallel for schedule(static)
  for (k = 0; k < g->knb; k++) {
    real pk = 0.;
    int j, i;
    for (j = 0; j < g->jnb; j++) {
      real pj = 0.;
      const real *ra = a + (size_t)k * g->s2b + (size_t)j * g->s1b, *rb = bvec + (size_t)k * g->s2b + (size_t)j * g->s1b;
      for (i = 0; i < g->inb; i++) pj += ra[i] * rb[i];
      pk += pj;
    }
    pp[k] = pk;
  }
  for (k = 0; k < g->knb; k++) tot += pp[k];
  return tot;
}

/* ------------------------------------------------------------------------------------ */
/* PP_cg_init, src/solver_kernel.cu:258-283 */
static void cg_init(bbo_state *s)
{
  int c;
  for (c = 0; c < s->nblocks; c++) {
    const grid_info *g = &s->dom[c].Gcc;
    bbo_block *b = &s->blk[c];
    int k;
#pragma omp parallel for schedule(static)
    for (k = 0; k < g->kn; k++) {
      int ti, tj;
      for (tj = 0; tj < g->jn; tj++) for (ti = 0; ti < g->in; ti++) {
        int cc = GCC_LOC(ti, tj, k, g->s1, g->s2);
        int C = GCC_LOC(ti + DOM_BUF, tj + DOM_BUF, k + DOM_BUF, g->s1b, g->s2b);
        real tmp = b->rhs_p[C], tmp2 = tmp * b->invM[cc];
        b->r_q[cc] = tmp; b->z_q[cc] = tmp2; b->p_q[cc] = tmp2; b->pb_q[C] = tmp2; b->phi[C] = 0.;
      }
    }
  }
}

/* PP_spmv_shared_load_noparts, src/solver_kernel.cu:715-836 (expression :824-829):
 * Ap = -idx2(fE^2(pE-pC) - fW^2(pC-pW)) - idy2(...) - idz2(...) on src (s3b) -> Apb_q (s3) */
void bbo_spmv_noparts(bbo_state *s, int src_id)
{
  int c;
  for (c = 0; c < s->nblocks; c++) {
    const dom_struct *d = &s->dom[c];
    const grid_info *g = &d->Gcc;
    bbo_block *b = &s->blk[c];
    const real *p = blk_gcc_array(b, src_id);
    real idx2 = 1. / (d->dx * d->dx), idy2 = 1. / (d->dy * d->dy), idz2 = 1. / (d->dz * d->dz);
    int TK;
#pragma omp parallel for schedule(static)
    for (TK = g->_ks; TK <= g->_ke; TK++) {
      int TI, TJ;
      for (TJ = g->_js; TJ <= g->_je; TJ++) for (TI = g->_is; TI <= g->_ie; TI++) {
        int C = GCC_LOC(TI, TJ, TK, g->s1b, g->s2b);
        int cc = GCC_LOC(TI - DOM_BUF, TJ - DOM_BUF, TK - DOM_BUF, g->s1, g->s2);
        int fw = b->flag_u[GFX_LOC(TI, TJ, TK, d->Gfx.s1b, d->Gfx.s2b)], fe = b->flag_u[GFX_LOC(TI + 1, TJ, TK, d->Gfx.s1b, d->Gfx.s2b)];
        int fs = b->flag_v[GFY_LOC(TI, TJ, TK, d->Gfy.s1b, d->Gfy.s2b)], fn = b->flag_v[GFY_LOC(TI, TJ + 1, TK, d->Gfy.s1b, d->Gfy.s2b)];
        int fb = b->flag_w[GFZ_LOC(TI, TJ, TK, d->Gfz.s1b, d->Gfz.s2b)], ft = b->flag_w[GFZ_LOC(TI, TJ, TK + 1, d->Gfz.s1b, d->Gfz.s2b)];
        real pc = p[C];
        b->Apb_q[cc] = -idx2 * (fe * fe * (p[C + 1] - pc) - fw * fw * (pc - p[C - 1]))
                       - idy2 * (fn * fn * (p[C + g->s1b] - pc) - fs * fs * (pc - p[C - g->s1b]))
                       - idz2 * (ft * ft * (p[C + g->s2b] - pc) - fb * fb * (pc - p[C - g->s2b]));
      }
    }
  }
}

/* PP_spmv_shared_load, src/solver_kernel.cu:528-713 (expression :683-707) */
void bbo_spmv_parts(bbo_state *s, int src_id)
{
  int c;
  for (c = 0; c < s->nblocks; c++) {
    const dom_struct *d = &s->dom[c];
    const grid_info *g = &d->Gcc;
    bbo_block *b = &s->blk[c];
    const real *p = blk_gcc_array(b, src_id);
    const int *ph = b->phase;
    real idx2 = 1. / (d->dx * d->dx), idy2 = 1. / (d->dy * d->dy), idz2 = 1. / (d->dz * d->dz);
    int TK;
#pragma omp parallel for schedule(static)
    for (TK = g->_ks; TK <= g->_ke; TK++) {
      int TI, TJ;
      for (TJ = g->_js; TJ <= g->_je; TJ++) for (TI = g->_is; TI <= g->_ie; TI++) {
        int C = GCC_LOC(TI, TJ, TK, g->s1b, g->s2b);
        int cc = GCC_LOC(TI - DOM_BUF, TJ - DOM_BUF, TK - DOM_BUF, g->s1, g->s2);
        int E = C + 1, W = C - 1, N = C + g->s1b, S = C - g->s1b, T = C + g->s2b, B = C - g->s2b;
        int fw = b->flag_u[GFX_LOC(TI, TJ, TK, d->Gfx.s1b, d->Gfx.s2b)], fe = b->flag_u[GFX_LOC(TI + 1, TJ, TK, d->Gfx.s1b, d->Gfx.s2b)];
        int fs = b->flag_v[GFY_LOC(TI, TJ, TK, d->Gfy.s1b, d->Gfy.s2b)], fn = b->flag_v[GFY_LOC(TI, TJ + 1, TK, d->Gfy.s1b, d->Gfy.s2b)];
        int fb = b->flag_w[GFZ_LOC(TI, TJ, TK, d->Gfz.s1b, d->Gfz.s2b)], ft = b->flag_w[GFZ_LOC(TI, TJ, TK + 1, d->Gfz.s1b, d->Gfz.s2b)];
        real pfx = -(d->dx * d->dx) / 6. * (ph[C] > -1) + (real)(ph[C] == -1);
        real pfy = -(d->dy * d->dy) / 6. * (ph[C] > -1) + (real)(ph[C] == -1);
        real pfz = -(d->dz * d->dz) / 6. * (ph[C] > -1) + (real)(ph[C] == -1);
        int pfe = (ph[C] == -1) * !(ph[C] == -1 && ph[E] > -1);
        int pfw = (ph[C] == -1) * !(ph[C] == -1 && ph[W] > -1);
        int pfn = (ph[C] == -1) * !(ph[C] == -1 && ph[N] > -1);
        int pfs = (ph[C] == -1) * !(ph[C] == -1 && ph[S] > -1);
        int pft = (ph[C] == -1) * !(ph[C] == -1 && ph[T] > -1);
        int pfb = (ph[C] == -1) * !(ph[C] == -1 && ph[B] > -1);
        real a;
        a = -idx2 * (fe * fe * (p[E] * pfe - pfx * p[C]) - fw * fw * (p[C] * pfx - pfw * p[W]));
        a += -idy2 * (fn * fn * (p[N] * pfn - pfy * p[C]) - fs * fs * (p[C] * pfy - pfs * p[S]));
        a += -idz2 * (ft * ft * (p[T] * pft - pfz * p[C]) - fb * fb * (p[C] * pfz - pfb * p[B]));
        b->Apb_q[cc] = a;
      }
    }
  }
}

static void spmv(bbo_state *s, int src_id, int parts) { if (parts) bbo_spmv_parts(s, src_id); else bbo_spmv_noparts(s, src_id); }

/* PP_update_soln_resid, src/solver_kernel.cu:838-862 */
static void update_soln_resid(bbo_state *s, real alpha)
{
  int c;
  for (c = 0; c < s->nblocks; c++) {
    const grid_info *g = &s->dom[c].Gcc;
    bbo_block *b = &s->blk[c];
    int k;
#pragma omp parallel for schedule(static)
    for (k = 0; k < g->kn; k++) {
      int ti, tj;
      for (tj = 0; tj < g->jn; tj++) for (ti = 0; ti < g->in; ti++) {
        int cc = GCC_LOC(ti, tj, k, g->s1, g->s2);
        int C = GCC_LOC(ti + DOM_BUF, tj + DOM_BUF, k + DOM_BUF, g->s1b, g->s2b);
        b->phi[C] += alpha * b->p_q[cc];
        b->r_q[cc] -= alpha * b->Apb_q[cc];
        b->z_q[cc] = b->r_q[cc] * b->invM[cc];
      }
    }
  }
}

/* PP_update_solution, src/solver_kernel.cu:864-881 */
static void update_solution(bbo_state *s, real alpha)
{
  int c;
  for (c = 0; c < s->nblocks; c++) {
    const grid_info *g = &s->dom[c].Gcc;
    bbo_block *b = &s->blk[c];
    int k;
#pragma omp parallel for schedule(static)
    for (k = 0; k < g->kn; k++) {
      int ti, tj;
      for (tj = 0; tj < g->jn; tj++) for (ti = 0; ti < g->in; ti++)
        b->phi[GCC_LOC(ti + DOM_BUF, tj + DOM_BUF, k + DOM_BUF, g->s1b, g->s2b)] += alpha * b->p_q[GCC_LOC(ti, tj, k, g->s1, g->s2)];
    }
  }
}

/* PP_update_residual, src/solver_kernel.cu:883-904 */
static void update_residual(bbo_state *s)
{
  int c;
  for (c = 0; c < s->nblocks; c++) {
    const grid_info *g = &s->dom[c].Gcc;
    bbo_block *b = &s->blk[c];
    int k;
#pragma omp parallel for schedule(static)
    for (k = 0; k < g->kn; k++) {
      int ti, tj;
      for (tj = 0; tj < g->jn; tj++) for (ti = 0; ti < g->in; ti++) {
        int cc = GCC_LOC(ti, tj, k, g->s1, g->s2);
        int C = GCC_LOC(ti + DOM_BUF, tj + DOM_BUF, k + DOM_BUF, g->s1b, g->s2b);
        b->r_q[cc] = b->rhs_p[C] - b->Apb_q[cc];
        b->z_q[cc] = b->r_q[cc] * b->invM[cc];
      }
    }
  }
}

/* PP_update_search, src/solver_kernel.cu:906-927 */
static void update_search(bbo_state *s, real beta)
{
  int c;
  for (c = 0; c < s->nblocks; c++) {
    const grid_info *g = &s->dom[c].Gcc;
    bbo_block *b = &s->blk[c];
    int k;
#pragma omp parallel for schedule(static)
    for (k = 0; k < g->kn; k++) {
      int ti, tj;
      for (tj = 0; tj < g->jn; tj++) for (ti = 0; ti < g->in; ti++) {
        int cc = GCC_LOC(ti, tj, k, g->s1, g->s2);
        int C = GCC_LOC(ti + DOM_BUF, tj + DOM_BUF, k + DOM_BUF, g->s1b, g->s2b);
        real np = b->z_q[cc] + beta * b->p_q[cc];
        b->p_q[cc] = np; b->pb_q[C] = np;
      }
    }
  }
}

/* ------------------------------------------------------------------------------------ */
/* Solve result */
typedef struct bbo_result {
  int status;        /* 0 converged, 1 tiny-rhs shortcut, 2 max-iter exceeded, 3 NaN */
  int niter;         /* q at exit (cuda_solver.cu:239) */
  real resid;        /* sqrt(sp_rq1)/sqrt(sp_rhs) as passed to recorder_PP */
  real sp_rhs;       /* (b,b) */
  real sp_rq0;       /* initial (r,z) */
} bbo_result;

/* cuda_PP_cg / cuda_PP_cg_noparts control flow, src/cuda_solver.cu:38-300 / :573-761.
 * parts != 0 selects the `if (NPARTS > 0)` block (:128-148) and the phase-aware SpMV.
 * history[q] (q = 1..niter, history[0] = initial (r,z)) receives sp_rq1 of every
 * iteration -- the reference never records it; it is what the +-1 iteration parity and the
 * residual-history comparison are made on.  Instead of exit(EXIT_FAILURE) (:245-251,
 * 271-279) a status is returned. */
int bbo_solve(bbo_state *s, real rho_f, real dt, real pp_residual, int pp_max_iter, int parts,
              real *history, int hist_cap, bbo_result *res)
{
  real sp_rhs = 0., sp_rq = 0., sp_rq1 = 0., alpha, beta, DENOM;
  int c, q = 0;
  const real RHS_TOL = 1.e-8;                                              /* :176 */

  bbo_rhs(s, rho_f, dt);                                                   /* :122-126 */
  if (parts) part_rhs_patch(s);                                            /* :128-148 */
  for (c = 0; c < s->nblocks; c++) sp_rhs += dot_s3b(s, c, s->blk[c].rhs_p, s->blk[c].rhs_p);  /* :151-152 */
  bbo_exchange_Gcc(s, BBO_RHS_P);                                          /* :155 */
  cg_init(s);                                                              /* :159 */
  bbo_exchange_Gcc(s, BBO_PB_Q);                                           /* :163 */
  for (c = 0; c < s->nblocks; c++) sp_rq += dot_s3(s, c, s->blk[c].r_q, s->blk[c].z_q);        /* :169-170 */
  res->sp_rhs = sp_rhs; res->sp_rq0 = sp_rq; res->resid = 0.; res->niter = 0;
  if (history && hist_cap > 0) history[0] = sp_rq;

  if (sp_rhs < RHS_TOL * RHS_TOL) { res->status = 1; return 1; }           /* :178-189 */

  while (q <= pp_max_iter) {                                               /* :192 */
    ++q;
    spmv(s, BBO_PB_Q, parts);                                              /* :201 */
    DENOM = 0.;
    for (c = 0; c < s->nblocks; c++) DENOM += dot_s3(s, c, s->blk[c].p_q, s->blk[c].Apb_q);    /* :204-205 */
    alpha = sp_rq / DENOM;                                                 /* :206 */
    if (q % 50 == 0) {                                                     /* :209-223 */
      update_solution(s, alpha);
      bbo_exchange_Gcc(s, BBO_PHI);
      spmv(s, BBO_PHI, parts);
      update_residual(s);
    } else {
      update_soln_resid(s, alpha);                                         /* :226 */
    }
    sp_rq1 = 0.;
    for (c = 0; c < s->nblocks; c++) sp_rq1 += dot_s3(s, c, s->blk[c].r_q, s->blk[c].z_q);     /* :231-232 */
    if (history && q < hist_cap) history[q] = sp_rq1;
    res->niter = q;
    if (sp_rq1 <= pp_residual * pp_residual * sp_rhs) {                    /* :235 */
      res->resid = sqrt(sp_rq1) / sqrt(sp_rhs);
      res->status = 0;
      return 0;
    } else if (isnan(sp_rq1)) {                                            /* :245 */
      res->resid = sp_rq1; res->status = 3; return 3;
    } else {
      beta = sp_rq1 / sp_rq;                                               /* :256 */
      update_search(s, beta);                                              /* :259 */
      bbo_exchange_Gcc(s, BBO_PB_Q);                                       /* :263 */
      sp_rq = sp_rq1;
    }
  }
  res->resid = sqrt(sp_rq1) / sqrt(sp_rhs);                                /* :271-279 */
  res->status = 2;
  return 2;
}

/* A fixed number of iterations of the hot loop with no stop test: used to time the CPU
 * baseline on a bounded sample (bench.py cpu_baseline / --impl reference fallback).
 * Same kernels, same order as bbo_solve's loop body. */
int bbo_iterate_fixed(bbo_state *s, real rho_f, real dt, int niters, int parts)
{
  real sp_rq = 0., sp_rq1, alpha, beta, DENOM;
  int c, q;
  bbo_rhs(s, rho_f, dt);
  if (parts) part_rhs_patch(s);
  bbo_exchange_Gcc(s, BBO_RHS_P);
  cg_init(s);
  bbo_exchange_Gcc(s, BBO_PB_Q);
  for (c = 0; c < s->nblocks; c++) sp_rq += dot_s3(s, c, s->blk[c].r_q, s->blk[c].z_q);
  for (q = 1; q <= niters; q++) {
    spmv(s, BBO_PB_Q, parts);
    DENOM = 0.;
    for (c = 0; c < s->nblocks; c++) DENOM += dot_s3(s, c, s->blk[c].p_q, s->blk[c].Apb_q);
    alpha = sp_rq / DENOM;
    update_soln_resid(s, alpha);
    sp_rq1 = 0.;
    for (c = 0; c < s->nblocks; c++) sp_rq1 += dot_s3(s, c, s->blk[c].r_q, s->blk[c].z_q);
    beta = sp_rq1 / sp_rq;
    update_search(s, beta);
    bbo_exchange_Gcc(s, BBO_PB_Q);
    sp_rq = sp_rq1;
  }
  return niters;
}

int bbo_omp_threads(void)
{
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
