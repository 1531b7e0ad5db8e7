/* bb_domain.c -- host side of the decomposition contract (no CUDA).
 *
 * Mirrors, for exactly what the pressure-Poisson path needs:
 *   domain_read_input   src/domain.c:72-160   (flow.config header + decomp.config records)
 *   pressure BC block   src/domain.c:216-287
 *   domain_fill         src/domain.c:918-1486 (index ranges, strides, neighbour ranks)
 *   decomp_reader       tools/src/decomp_reader.c:112-154 (equal splits + record writer)
 *
 * flow.config is read by KEY (a line scanner), not positionally: every flow.config shipped
 * under examples/ is out of sync with the reference's positional fscanf chain (SURVEY.md 5,
 * "Config / flags"), and a key scanner accepts both generations of the file.
 */
#include <ctype.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/bbpcg.h"

extern void bbpcg_set_error(const char *fmt, ...);

/* One axis of one grid: n points of which the first interior one has global index gs. */
typedef struct axis_rng { int s, e, n, sb, eb, nb, _s, _e, _sb, _eb; } axis_rng;

static axis_rng make_axis(int n, int gs)
{
  axis_rng a;
  a.s = gs;            a.sb = a.s - DOM_BUF;
  a.n = n;             a.nb = n + 2 * DOM_BUF;
  a.e = a.sb + n;      a.eb = a.e + DOM_BUF;
  a._s = DOM_BUF;      a._sb = 0;
  a._e = a._sb + n;    a._eb = a._e + DOM_BUF;
  return a;
}

/* storage order of each grid (src/bluebottle.h:70-73): which axis is fastest / middle / slowest */
enum { ORD_IJK = 0, ORD_JKI = 1, ORD_KIJ = 2 };

static void store_grid(grid_info *g, axis_rng ai, axis_rng aj, axis_rng ak, int ord)
{
  int f, m, l, fb, mb, lb;   /* fastest, middle, slowest extents (interior / with ghosts) */
  g->is = ai.s; g->ie = ai.e; g->in = ai.n; g->isb = ai.sb; g->ieb = ai.eb; g->inb = ai.nb;
  g->js = aj.s; g->je = aj.e; g->jn = aj.n; g->jsb = aj.sb; g->jeb = aj.eb; g->jnb = aj.nb;
  g->ks = ak.s; g->ke = ak.e; g->kn = ak.n; g->ksb = ak.sb; g->keb = ak.eb; g->knb = ak.nb;
  g->_is = ai._s; g->_ie = ai._e; g->_isb = ai._sb; g->_ieb = ai._eb;
  g->_js = aj._s; g->_je = aj._e; g->_jsb = aj._sb; g->_jeb = aj._eb;
  g->_ks = ak._s; g->_ke = ak._e; g->_ksb = ak._sb; g->_keb = ak._eb;
  switch (ord) {
    case ORD_JKI: f = aj.n; m = ak.n; l = ai.n; fb = aj.nb; mb = ak.nb; lb = ai.nb; break;
    case ORD_KIJ: f = ak.n; m = ai.n; l = aj.n; fb = ak.nb; mb = ai.nb; lb = aj.nb; break;
    default:      f = ai.n; m = aj.n; l = ak.n; fb = ai.nb; mb = aj.nb; lb = ak.nb; break;
  }
  g->s1 = f;   g->s2 = f * m;    g->s3 = f * m * l;
  g->s1b = fb; g->s2b = fb * mb; g->s3b = fb * mb * lb;
  g->s2_i = aj.n * ak.n;    g->s2_j = ai.n * ak.n;    g->s2_k = ai.n * aj.n;
  g->s2b_i = aj.nb * ak.nb; g->s2b_j = ai.nb * ak.nb; g->s2b_k = ai.nb * aj.nb;
}

/* cell-centred grid + the three face grids of a box with (xn,yn,zn) cells whose first cell
 * has global cell index (ci,cj,ck) and first face global face index (fi,fj,fk) */
static void store_four(dom_struct *d, int ci, int cj, int ck, int fi, int fj, int fk)
{
  store_grid(&d->Gcc, make_axis(d->xn, ci),     make_axis(d->yn, cj),     make_axis(d->zn, ck),     ORD_IJK);
  store_grid(&d->Gfx, make_axis(d->xn + 1, fi), make_axis(d->yn, cj),     make_axis(d->zn, ck),     ORD_JKI);
  store_grid(&d->Gfy, make_axis(d->xn, ci),     make_axis(d->yn + 1, fj), make_axis(d->zn, ck),     ORD_KIJ);
  store_grid(&d->Gfz, make_axis(d->xn, ci),     make_axis(d->yn, cj),     make_axis(d->zn + 1, fk), ORD_IJK);
}

static int block_rank(const dom_struct *DOM, int I, int J, int K) { return I + J * DOM->S1 + K * DOM->S2; }

/* neighbour across one side: the adjacent block, the wrap-around block if that side's
 * pressure BC is PERIODIC, else none (src/domain.c:1147-1210) */
static int side_nbr(const dom_struct *DOM, int I, int J, int K, int axis, int dir, int bctype)
{
  int c[3] = { I, J, K }, n[3] = { DOM->In, DOM->Jn, DOM->Kn };
  int v = c[axis] + dir;
  if (v < 0 || v >= n[axis]) {
    if (bctype != BB_PERIODIC) return BB_PROC_NULL;
    v = (v + n[axis]) % n[axis];
  }
  c[axis] = v;
  return block_rank(DOM, c[0], c[1], c[2]);
}

int bb_domain_fill(dom_struct *DOM, dom_struct *dom, const bb_pressure_bc *bc)
{
  int c, S3;
  if (!DOM || !dom || !bc || DOM->In < 1 || DOM->Jn < 1 || DOM->Kn < 1) { bbpcg_set_error("bb_domain_fill: bad arguments"); return BBPCG_EINVAL; }
  DOM->S1 = DOM->In; DOM->S2 = DOM->In * DOM->Jn; DOM->S3 = S3 = DOM->In * DOM->Jn * DOM->Kn;
  DOM->xl = DOM->xe - DOM->xs; DOM->yl = DOM->ye - DOM->ys; DOM->zl = DOM->ze - DOM->zs;
  DOM->dx = DOM->xl / DOM->xn; DOM->dy = DOM->yl / DOM->yn; DOM->dz = DOM->zl / DOM->zn;
  store_four(DOM, DOM_BUF, DOM_BUF, DOM_BUF, DOM_BUF, DOM_BUF, DOM_BUF);
  DOM->Is = DOM->Js = DOM->Ks = 0; DOM->Ie = DOM->In - 1; DOM->Je = DOM->Jn - 1; DOM->Ke = DOM->Kn - 1;
  DOM->I = DOM->J = DOM->K = 0; DOM->rank = 0;
  DOM->e = DOM->w = DOM->n = DOM->s = DOM->t = DOM->b = BB_PROC_NULL;

  /* records are in rank order (I fastest), so west/south/bottom neighbours precede a block */
  for (c = 0; c < S3; c++) {
    dom_struct *d = &dom[c];
    int ci = DOM_BUF, cj = DOM_BUF, ck = DOM_BUF, fi = DOM_BUF, fj = DOM_BUF, fk = DOM_BUF;
    if (d->I < 0 || d->I >= DOM->In || d->J < 0 || d->J >= DOM->Jn || d->K < 0 || d->K >= DOM->Kn ||
        block_rank(DOM, d->I, d->J, d->K) != c || d->xn < 1 || d->yn < 1 || d->zn < 1) {
      bbpcg_set_error("bb_domain_fill: record %d is not block (%d,%d,%d) in I-fastest order or has an empty extent", c, d->I, d->J, d->K);
      return BBPCG_EINVAL;
    }
    d->rank = c;
    d->w = side_nbr(DOM, d->I, d->J, d->K, 0, -1, bc->pW);  d->e = side_nbr(DOM, d->I, d->J, d->K, 0, +1, bc->pE);
    d->s = side_nbr(DOM, d->I, d->J, d->K, 1, -1, bc->pS);  d->n = side_nbr(DOM, d->I, d->J, d->K, 1, +1, bc->pN);
    d->b = side_nbr(DOM, d->I, d->J, d->K, 2, -1, bc->pB);  d->t = side_nbr(DOM, d->I, d->J, d->K, 2, +1, bc->pT);
    d->xl = d->xe - d->xs; d->yl = d->ye - d->ys; d->zl = d->ze - d->zs;       /* src/domain.c:1219-1224 */
    d->dx = d->xl / d->xn; d->dy = d->yl / d->yn; d->dz = d->zl / d->zn;
    /* global start indices chain off the lower neighbour; a face grid shares the block-boundary
     * face with it (src/domain.c:1229-1233,1292-1296) */
    if (d->I > 0) { const dom_struct *W = &dom[block_rank(DOM, d->I - 1, d->J, d->K)]; ci = W->Gcc.ie + 1; fi = W->Gfx.ie; }
    if (d->J > 0) { const dom_struct *S = &dom[block_rank(DOM, d->I, d->J - 1, d->K)]; cj = S->Gcc.je + 1; fj = S->Gfy.je; }
    if (d->K > 0) { const dom_struct *B = &dom[block_rank(DOM, d->I, d->J, d->K - 1)]; ck = B->Gcc.ke + 1; fk = B->Gfz.ke; }
    store_four(d, ci, cj, ck, fi, fj, fk);
    d->Is = DOM->Is; d->Ie = DOM->Ie; d->In = DOM->In; d->Js = DOM->Js; d->Je = DOM->Je; d->Jn = DOM->Jn;
    d->Ks = DOM->Ks; d->Ke = DOM->Ke; d->Kn = DOM->Kn; d->S1 = DOM->S1; d->S2 = DOM->S2; d->S3 = DOM->S3;
  }
  return BBPCG_OK;
}

int bb_domain_split(dom_struct *DOM, dom_struct *dom)
{
  int i, j, k;
  double xl, yl, zl;
  if (!DOM || !dom || DOM->In < 1 || DOM->Jn < 1 || DOM->Kn < 1) { bbpcg_set_error("bb_domain_split: bad arguments"); return BBPCG_EINVAL; }
  xl = (DOM->xe - DOM->xs) / DOM->In; yl = (DOM->ye - DOM->ys) / DOM->Jn; zl = (DOM->ze - DOM->zs) / DOM->Kn;
  for (k = 0; k < DOM->Kn; k++) for (j = 0; j < DOM->Jn; j++) for (i = 0; i < DOM->In; i++) {
    dom_struct *d = &dom[i + DOM->In * (j + DOM->Jn * k)];
    d->I = i; d->J = j; d->K = k;
    d->xn = DOM->xn / DOM->In; d->xs = DOM->xs + i * xl; d->xe = d->xs + xl;
    d->yn = DOM->yn / DOM->Jn; d->ys = DOM->ys + j * yl; d->ye = d->ys + yl;
    d->zn = DOM->zn / DOM->Kn; d->zs = DOM->zs + k * zl; d->ze = d->zs + zl;
  }
  return BBPCG_OK;
}

int bb_domain_write_decomp(const char *path, const dom_struct *DOM, const dom_struct *dom, int prec)
{
  FILE *f = fopen(path, "w");
  int c, S3 = DOM->In * DOM->Jn * DOM->Kn;
  if (!f) { bbpcg_set_error("cannot write %s", path); return BBPCG_EIO; }
  for (c = 0; c < S3; c++) {
    fprintf(f, "(I, J, K) %d %d %d\n", dom[c].I, dom[c].J, dom[c].K);
    fprintf(f, "(Xs, Xe, Xn) %.*f %.*f %d\n", prec, dom[c].xs, prec, dom[c].xe, dom[c].xn);
    fprintf(f, "(Ys, Ye, Yn) %.*f %.*f %d\n", prec, dom[c].ys, prec, dom[c].ye, dom[c].yn);
    fprintf(f, "(Zs, Ze, Zn) %.*f %.*f %d\n\n", prec, dom[c].zs, prec, dom[c].ze, dom[c].zn);
  }
  fclose(f);
  return BBPCG_OK;
}

/* ---- key-based scanners ---------------------------------------------------------------- */
static char *slurp(const char *path)
{
  FILE *f = fopen(path, "rb");
  long n;
  char *buf;
  if (!f) return NULL;
  fseek(f, 0, SEEK_END); n = ftell(f); fseek(f, 0, SEEK_SET);
  buf = (char *)malloc((size_t)n + 1);
  if (buf && fread(buf, 1, (size_t)n, f) != (size_t)n) { free(buf); buf = NULL; }
  if (buf) buf[n] = 0;
  fclose(f);
  return buf;
}

/* find `key` at the start of a line; returns pointer just past it, or NULL */
static const char *find_key(const char *text, const char *key)
{
  size_t kl = strlen(key);
  const char *p = text;
  while ((p = strstr(p, key)) != NULL) {
    if ((p == text || p[-1] == '\n') && (isspace((unsigned char)p[kl]) || p[kl] == 0)) return p + kl;
    p += kl;
  }
  return NULL;
}

static int read_triple(const char *text, const char *key, double *a, double *b, int *n)
{
  const char *p = find_key(text, key);
  return (p && sscanf(p, "%lf %lf %d", a, b, n) == 3) ? 0 : -1;
}

static int read_bc(const char *text, const char *key, int *out)
{
  char word[64];
  const char *p = find_key(text, key);
  if (!p || sscanf(p, "%63s", word) != 1) return -1;
  if (strcmp(word, "PERIODIC") == 0) *out = BB_PERIODIC;
  else if (strcmp(word, "NEUMANN") == 0) *out = BB_NEUMANN;
  else return -1;                          /* pressure is PERIODIC or NEUMANN only, domain.c:216-287 */
  return 0;
}

int bb_domain_read(const char *flow_config, const char *decomp_config, dom_struct *DOM,
                   dom_struct **dom_out, bb_pressure_bc *bc, bb_flow_params *params)
{
  char *flow = slurp(flow_config), *dec;
  const char *p;
  dom_struct *dom;
  int c, S3, rc;
  if (!flow) { bbpcg_set_error("Could not open file %s", flow_config); return BBPCG_EIO; }
  memset(DOM, 0, sizeof(*DOM));
  if (read_triple(flow, "(Xs, Xe, Xn)", &DOM->xs, &DOM->xe, &DOM->xn) ||
      read_triple(flow, "(Ys, Ye, Yn)", &DOM->ys, &DOM->ye, &DOM->yn) ||
      read_triple(flow, "(Zs, Ze, Zn)", &DOM->zs, &DOM->ze, &DOM->zn)) {
    bbpcg_set_error("%s: GLOBAL DOMAIN block not found", flow_config); free(flow); return BBPCG_EIO;
  }
  p = find_key(flow, "(In, Jn, Kn)");
  if (!p || sscanf(p, "%d %d %d", &DOM->In, &DOM->Jn, &DOM->Kn) != 3) {
    bbpcg_set_error("%s: (In, Jn, Kn) not found", flow_config); free(flow); return BBPCG_EIO;
  }
  if (read_bc(flow, "bc.pW", &bc->pW) || read_bc(flow, "bc.pE", &bc->pE) || read_bc(flow, "bc.pS", &bc->pS) ||
      read_bc(flow, "bc.pN", &bc->pN) || read_bc(flow, "bc.pB", &bc->pB) || read_bc(flow, "bc.pT", &bc->pT)) {
    bbpcg_set_error("%s: pressure boundary block (bc.pW..bc.pT PERIODIC|NEUMANN) not readable", flow_config);
    free(flow); return BBPCG_EIO;
  }
  if (params) {
    params->rho_f = 1.; params->pp_residual = 1e-6; params->pp_max_iter = 2000;
    if ((p = find_key(flow, "rho_f")) != NULL) sscanf(p, "%lf", &params->rho_f);
    if ((p = find_key(flow, "pp_residual")) != NULL) sscanf(p, "%lf", &params->pp_residual);
    if ((p = find_key(flow, "pp_max_iter")) != NULL) sscanf(p, "%d", &params->pp_max_iter);
  }
  free(flow);

  S3 = DOM->In * DOM->Jn * DOM->Kn;
  if (S3 < 1) { bbpcg_set_error("bad decomposition %d x %d x %d", DOM->In, DOM->Jn, DOM->Kn); return BBPCG_EINVAL; }
  dec = slurp(decomp_config);
  if (!dec) { bbpcg_set_error("Could not open file %s", decomp_config); return BBPCG_EIO; }
  dom = (dom_struct *)calloc((size_t)S3, sizeof(dom_struct));
  p = dec;
  for (c = 0; c < S3; c++) {              /* record grammar: src/domain.c:138-159 */
    const char *q = strstr(p, "(I, J, K)");
    dom_struct *d = &dom[c];
    if (!q || sscanf(q, "(I, J, K) %d %d %d (Xs, Xe, Xn) %lf %lf %d (Ys, Ye, Yn) %lf %lf %d (Zs, Ze, Zn) %lf %lf %d",
                     &d->I, &d->J, &d->K, &d->xs, &d->xe, &d->xn, &d->ys, &d->ye, &d->yn, &d->zs, &d->ze, &d->zn) != 12) {
      bbpcg_set_error("%s: record %d of %d unreadable", decomp_config, c, S3);
      free(dec); free(dom); return BBPCG_EIO;
    }
    p = q + 9;
  }
  free(dec);
  rc = bb_domain_fill(DOM, dom, bc);
  if (rc) { free(dom); return rc; }
  *dom_out = dom;
  return BBPCG_OK;
}

void bb_domain_free(dom_struct *dom) { free(dom); }

/* ---- restart files as fixtures: out_restart / in_restart, src/domain.c:3005-3260 ------------------------------- */
int bb_restart_path(char *out, size_t cap, const char *dir, int rank, int S3)
{
  int sigfigs = 1, v;
  if (!out || !dir || rank < 0 || S3 < 1 || rank >= S3) { bbpcg_set_error("bb_restart_path: bad arguments"); return BBPCG_EINVAL; }
  for (v = S3 - 1; v >= 10; v /= 10) sigfigs++;            /* floor(log10(S3 - 1)) + 1, 1 for S3 == 1 (domain.c:3009-3014) */
  if (snprintf(out, cap, "%s/restart.config-%0*d", dir, sigfigs, rank) >= (int)cap) { bbpcg_set_error("bb_restart_path: path too long"); return BBPCG_EINVAL; }
  return BBPCG_OK;
}

void bb_restart_free(bb_restart *r)
{
  if (!r) return;
  free(r->u); free(r->v); free(r->w); free(r->u_star); free(r->v_star); free(r->w_star);
  free(r->p); free(r->phi); free(r->p0); free(r->phase); free(r->phase_shell);
  free(r->flag_u); free(r->flag_v); free(r->flag_w);
  memset(r, 0, sizeof(*r));
}

static int rd(FILE *f, void *dst, size_t size, size_t n) { return fread(dst, size, n, f) == n ? 0 : -1; }
static int skip(FILE *f, size_t size, size_t n) { return fseek(f, (long)(size * n), SEEK_CUR); }
static void *grab(FILE *f, size_t size, size_t n, int *err)
{
  void *p = malloc(size * (n ? n : 1));
  if (!p || rd(f, p, size, n)) { free(p); *err = 1; return NULL; }
  return p;
}

int bb_restart_read(const char *path, const dom_struct *d, bb_restart *r)
{
  FILE *f;
  int err = 0, g;
  if (!path || !d || !r) { bbpcg_set_error("bb_restart_read: NULL argument"); return BBPCG_EINVAL; }
  memset(r, 0, sizeof(*r));
  f = fopen(path, "rb");
  if (!f) { bbpcg_set_error("File %s could not be opened.", path); return BBPCG_EIO; }          /* the reference's message, domain.c:3098 */
  /* header, domain.c:3026-3033 */
  err |= rd(f, &r->ttime, sizeof(real), 1) | rd(f, &r->dt0, sizeof(real), 1) | rd(f, &r->dt, sizeof(real), 1);
  err |= rd(f, &r->stepnum, sizeof(int), 1) | rd(f, &r->rec_vtk_stepnum_out, sizeof(int), 1);
  err |= rd(f, &r->rec_cgns_flow_ttime_out, sizeof(real), 1) | rd(f, &r->rec_cgns_part_ttime_out, sizeof(real), 1) |
         rd(f, &r->rec_vtk_ttime_out, sizeof(real), 1);
  /* seven arrays per velocity component, domain.c:3036-3058: X, X0, diff0, conv0, diff, conv, X_star */
  for (g = 0; g < 3 && !err; g++) {
    const size_t n = (size_t)(g == 0 ? d->Gfx.s3b : g == 1 ? d->Gfy.s3b : d->Gfz.s3b);
    real **vel = g == 0 ? &r->u : g == 1 ? &r->v : &r->w;
    real **star = g == 0 ? &r->u_star : g == 1 ? &r->v_star : &r->w_star;
    *vel = (real *)grab(f, sizeof(real), n, &err);
    if (!err && skip(f, sizeof(real), 5 * n)) err = 1;
    if (!err) *star = (real *)grab(f, sizeof(real), n, &err);
  }
  if (!err) {                                                                                   /* domain.c:3060-3068 */
    const size_t n = (size_t)d->Gcc.s3b;
    r->p = (real *)grab(f, sizeof(real), n, &err);
    if (!err) r->phi = (real *)grab(f, sizeof(real), n, &err);
    if (!err) r->p0 = (real *)grab(f, sizeof(real), n, &err);
    if (!err) r->phase = (int *)grab(f, sizeof(int), n, &err);
    if (!err) r->phase_shell = (int *)grab(f, sizeof(int), n, &err);
    if (!err) r->flag_u = (int *)grab(f, sizeof(int), (size_t)d->Gfx.s3b, &err);
    if (!err) r->flag_v = (int *)grab(f, sizeof(int), (size_t)d->Gfy.s3b, &err);
    if (!err) r->flag_w = (int *)grab(f, sizeof(int), (size_t)d->Gfz.s3b, &err);
    if (!err) err |= rd(f, &r->nparts_subdom, sizeof(int), 1);                                  /* :3070 */
  }
  fclose(f);
  if (err) {
    bb_restart_free(r);
    bbpcg_set_error("%s: shorter than a restart file of a %d x %d x %d block (or out of memory)", path, d->xn, d->yn, d->zn);
    return BBPCG_EIO;
  }
  if (r->nparts_subdom < 0) { bb_restart_free(r); bbpcg_set_error("%s: negative particle count -- not a restart file of this block", path); return BBPCG_EIO; }
  return BBPCG_OK;
}
