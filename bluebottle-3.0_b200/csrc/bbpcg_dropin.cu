/* bbpcg_dropin.cu -- the reference's entry points on top of libbbpcg (see include/bb_dropin.h).
 * Every function mirrors its reference namesake's observable behaviour: same globals read,
 * _phi written, recorder_PP called with (niter, resid, elapsed seconds), print + exit on
 * non-convergence / NaN (src/cuda_solver.cu:245-251, 271-279). */
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <sys/time.h>
#include <cuda_runtime.h>

#include "../../include/bbpcg.h"
#include "../../include/bb_dropin.h"

extern "C" {
/* globals owned by the Bluebottle host program */
extern dom_struct *dom;                       /* src/bluebottle.h:511 */
extern dom_struct DOM;                        /* src/bluebottle.h:487 */
extern int rank, nprocs;                      /* src/mpi_comm.h:218-230 */
extern bb_BC bc;                              /* the reference's BC struct, src/bluebottle.h:662-747 (mirrored field for field) */
extern real rho_f, dt, pp_residual, ttime;    /* src/bluebottle.h:524,560,572,2439 */
extern int pp_max_iter, stepnum;              /* src/bluebottle.h:2389,2487 */
extern int NPARTS, nparts;                    /* src/particle.h:319,331 */
extern real *_u_star, *_v_star, *_w_star, *_rhs_p, *_phi;
extern real *_u, *_v, *_w, *_p, *_p0;         /* src/bluebottle.h:961,1037,1137,1186,1235 (epilogue) */
extern int *_flag_u, *_flag_v, *_flag_w, *_phase, *_phase_shell;
extern int out_plane;                         /* src/bluebottle.h:650 (solvability) */
extern void *_parts;                          /* part_struct *_parts, src/particle.h:692: this rank's particles on the device */
void cuda_part_BC_p(void);                    /* src/cuda_particle.cu:1680 */
void recorder_PP(char *name, int niter, real resid, real etime);   /* src/recorder.c:190 */
/* cuda_PP_cg_timed's log (src/recorder.c:223-336); weak: a host linked without recorder.o's timed pair still loads */
void recorder_PP_init_timed(char *name) __attribute__((weak));
void recorder_PP_timed(char *name, int niter, real resid, real etime, real etime_spmv, real etime_ip1, real etime_AR1,
                       real etime_up1, real etime_ip2, real etime_AR2, real etime_up2, real etime_mpi) __attribute__((weak));
int bb_dropin_allgather(const void *send, void *recv, int bytes_per_rank) __attribute__((weak));
}

static bbpcg_solver *g_solver = NULL;

static void die(const char *what)
{
  fprintf(stderr, "N%d >> bbpcg: %s: %s\n", rank, what, bbpcg_last_error());
  exit(EXIT_FAILURE);
}

static bbpcg_solver *solver(void)
{
  if (g_solver) return g_solver;
  /* the reference selects the device before MPI_Init (src/mpi_comm.c:42-63); use it as is */
  if (bbpcg_create(&g_solver, &dom[rank], &DOM, (const bb_pressure_bc *)&bc, -1)) die("bbpcg_create");   /* BC opens with the six pressure types */
  if (nprocs > 1) {
    if (!bb_dropin_allgather) { fprintf(stderr, "N%d >> bbpcg: nprocs = %d but the host program does not define bb_dropin_allgather()\n", rank, nprocs); exit(EXIT_FAILURE); }
    char mine[BBPCG_BLOB_BYTES];
    char *all = (char *)malloc((size_t)nprocs * BBPCG_BLOB_BYTES);
    if (bbpcg_comm_export(g_solver, mine)) die("bbpcg_comm_export");
    if (bb_dropin_allgather(mine, all, BBPCG_BLOB_BYTES)) { fprintf(stderr, "N%d >> bbpcg: bb_dropin_allgather failed\n", rank); exit(EXIT_FAILURE); }
    if (bbpcg_comm_import(g_solver, all, nprocs)) die("bbpcg_comm_import");
    free(all);
  }
  return g_solver;
}

/* Layout of the reference's part_struct (src/particle.h:343-...) with -DDOUBLE: `int N; real r; real x; real y; real z; ...`,
 * sizeof 3128.  Only these four fields are read.  tests/test_abi.py re-derives the numbers from the reference header
 * whenever /root/reference is present; INTEGRATION.md shows the static_assert a maintainer adds on the Bluebottle side. */
#define BB_PART_STRIDE 3128
#define BB_PART_OFF_R 8
#define BB_PART_OFF_X 16
#define BB_PART_OFF_Y 24
#define BB_PART_OFF_Z 32

/* src/bluebottle.c:389 (+ :392-394: the preconditioner re-initialisation that follows it is folded in) */
extern "C" void cuda_build_cages(void)
{
  bbpcg_parts_view pv;
  pv.base = _parts; pv.stride = BB_PART_STRIDE;
  pv.off_x = BB_PART_OFF_X; pv.off_y = BB_PART_OFF_Y; pv.off_z = BB_PART_OFF_Z; pv.off_r = BB_PART_OFF_R;
  cudaDeviceSynchronize();
  if (bbpcg_build_cages(solver(), NPARTS, nparts, &pv, _flag_u, _flag_v, _flag_w, _phase, _phase_shell)) die("bbpcg_build_cages");
}

extern "C" void cuda_PP_init_jacobi_preconditioner(void)
{
  /* _phase is only meaningful when particles exist (src/cuda_particle.cu:66-69 vs
   * src/particle_kernel.cu:121-133), so it is digested only then */
  if (bbpcg_set_coefficients(solver(), _flag_u, _flag_v, _flag_w, NPARTS > 0 ? _phase : NULL)) die("bbpcg_set_coefficients");
}

/* The reference's eight wall-clock segments (src/cuda_solver.cu:372-390) do not exist as separate steps here: an iteration
 * is two fused kernels with the reductions, the all-reduces and the halo push inside them.  The columns are filled with
 * what can be measured (CUDA events around every launch, option kernel_timing):
 *   spmv  <- k_search_tma  (PP_update_search + SpMV + (p,Ap) + its all-reduce)
 *   up1   <- k_resid_tma   (PP_update_soln_resid + the halo push of r + (r,z) + its all-reduce), incl. the every-50th refresh pair
 *   ip1, ar1, ip2, ar2, up2, mpi <- 0 (fused into the two above) */
struct timed_segments { real spmv, up1; };

static void record(char *rname, int timed, int niter, real resid, real etime, const timed_segments &t)
{
  if (timed && recorder_PP_timed) recorder_PP_timed(rname, niter, resid, etime, t.spmv, 0., 0., t.up1, 0., 0., 0., 0.);
  else recorder_PP(rname, niter, resid, etime);
}

static void run(int use_phase, int timed = 0)
{
  struct timeval ts, te;
  gettimeofday(&ts, 0);                                     /* src/cuda_solver.cu:42-43 */
  if (timed) bbpcg_set_option(solver(), "kernel_timing", 1);
  bbpcg_solve_args a;
  bbpcg_result res;
  memset(&a, 0, sizeof(a));
  a.u_star = _u_star; a.v_star = _v_star; a.w_star = _w_star; a.rhs_p = _rhs_p; a.phi = _phi;
  a.phase = _phase; a.phase_shell = _phase_shell;
  a.rho_f = rho_f; a.dt = dt; a.pp_residual = pp_residual; a.pp_max_iter = pp_max_iter;
  a.use_phase = use_phase;
  a.part_bc = (use_phase && nparts > 0) ? cuda_part_BC_p : NULL;     /* :130-132 */
  if (use_phase && nparts <= 0) { a.phase_shell = NULL; a.no_refine = 1; }   /* no patch, no coeffs_refine on particle-free ranks (:130-142) */
  cudaDeviceSynchronize();          /* the caller's default-stream work on u*, flags is complete */
  if (bbpcg_solve(solver(), &a, &res)) die("bbpcg_solve");
  gettimeofday(&te, 0);
  real etime = (te.tv_sec - ts.tv_sec) + (te.tv_usec - ts.tv_usec) * 1.e-6;
  char rname_plain[] = "solver_expd.rec", rname_timed[] = "solver_expd_timed.rec";      /* :174, :367 */
  char *rname = (timed && recorder_PP_timed) ? rname_timed : rname_plain;
  timed_segments seg = { 0., 0. };
  if (timed) {
    seg.spmv = bbpcg_get_info(solver(), "kt_search_ns") * 1.e-9;
    seg.up1 = (bbpcg_get_info(solver(), "kt_resid_ns") + bbpcg_get_info(solver(), "kt_refresh_ns")) * 1.e-9;
    bbpcg_set_option(solver(), "kernel_timing", 0);
    if (stepnum == 1 && recorder_PP_init_timed) recorder_PP_init_timed(rname);          /* :392-394 */
  }
  switch (res.status) {
    case BBPCG_TINY_RHS:                                   /* :178-189 */
      record(rname, timed, 0, 0., etime, seg);
      if (rank == 0) printf("N%d >> Norm of the rhs is less than %.1e, exiting solver\n", rank, 1.e-8);
      break;
    case BBPCG_CONVERGED:                                  /* :235-241 */
      record(rname, timed, res.niter, res.resid, etime, seg);
      if (res.niter > pp_max_iter) {                       /* converged on iteration pp_max_iter + 1: the reference records the line,
                                                              breaks, and still fails its `q > pp_max_iter` test (:271-279) */
        printf("N%d >> The pressure-Poisson equation did not converge.\n", rank);
        printf("N%d >> (rhs, rhs) is %e\n", rank, res.sp_rhs);
        printf("N%d >> Residual at iteration %d is %lf\n", rank, res.niter, res.resid);
        exit(EXIT_FAILURE);
      }
      break;
    case BBPCG_NAN:                                        /* :245-251 */
      if (rank == 0) {
        printf("N%d >> The PP equation did not converge.\n", rank);
        printf("N%d >> The residual at iteration %d is nan (%lf).\n", rank, res.niter, res.resid);
      }
      exit(EXIT_FAILURE);
    default:                                               /* :271-279 */
      printf("N%d >> The pressure-Poisson equation did not converge.\n", rank);
      printf("N%d >> (rhs, rhs) is %e\n", rank, res.sp_rhs);
      printf("N%d >> Residual at iteration %d is %lf\n", rank, res.niter, res.resid);
      exit(EXIT_FAILURE);
  }
}

extern "C" void cuda_PP_cg(void) { run(NPARTS > 0 ? 1 : 0); }          /* the phase-aware operator equals the plain one when phase == -1 everywhere */
extern "C" void cuda_PP_cg_noparts(void) { run(0); }
extern "C" void cuda_PP_cg_timed(void) { run(0, 1); }                  /* reference: noparts variant + segment timers -> solver_expd_timed.rec (:302-571), no caller */

extern "C" void mpi_cuda_exchange_Gcc(real *array)
{
  cudaDeviceSynchronize();
  if (bbpcg_exchange_Gcc(solver(), array)) die("bbpcg_exchange_Gcc");
}

/* the face-grid exchanges on the same transport, src/mpi_comm.c:317-405 */
static void exchange_face(real *array, int grid)
{
  cudaDeviceSynchronize();
  if (bbpcg_exchange(solver(), array, grid)) die("bbpcg_exchange");
}
extern "C" void mpi_cuda_exchange_Gfx(real *array) { exchange_face(array, BBPCG_GFX); }
extern "C" void mpi_cuda_exchange_Gfy(real *array) { exchange_face(array, BBPCG_GFY); }
extern "C" void mpi_cuda_exchange_Gfz(real *array) { exchange_face(array, BBPCG_GFZ); }

/* ---- solve prologue: src/bluebottle.c:220, src/cuda_bluebottle.cu:2313-2492 ---- */
extern "C" void cuda_solvability(void)
{
  cudaDeviceSynchronize();
  if (bbpcg_solvability(solver(), _u_star, _v_star, _w_star, out_plane, NULL)) die("bbpcg_solvability");
}

/* src/bluebottle.c:214,222; src/cuda_bluebottle.cu:2111-2311: reads the types and the CURRENT Dirichlet values of bc */
extern "C" void cuda_dom_BC_star(void)
{
  bb_velocity_bc vbc;
  const bb_bc_entry *comp[3] = { bc.u, bc.v, bc.w };
  for (int c = 0; c < 3; c++) for (int f = 0; f < 6; f++) { vbc.type[c][f] = comp[c][f].type; vbc.val[c][f] = comp[c][f].D; }
  cudaDeviceSynchronize();
  if (bbpcg_dom_BC_star(solver(), _u_star, _v_star, _w_star, &vbc)) die("bbpcg_dom_BC_star");
}

/* ---- solve epilogue, src/bluebottle.c:233-256 ---- */
extern "C" void cuda_dom_BC_p(real *array)                  /* src/cuda_bluebottle.cu:2536-2589 */
{
  cudaDeviceSynchronize();
  if (bbpcg_dom_BC_p(solver(), array)) die("bbpcg_dom_BC_p");
}

static void epilogue(int project, int update)
{
  bbpcg_epilogue_args a;
  memset(&a, 0, sizeof(a));
  a.phi = _phi; a.rho_f = rho_f; a.dt = dt;
  a.phi_ghosts_valid = 1;                                   /* the caller ran mpi_cuda_exchange_Gcc(_phi); cuda_dom_BC_p(_phi) (bluebottle.c:233-234) */
  if (project) { a.u_star = _u_star; a.v_star = _v_star; a.w_star = _w_star; a.flag_u = _flag_u; a.flag_v = _flag_v; a.flag_w = _flag_w; a.u = _u; a.v = _v; a.w = _w; }
  if (update) { a.p0 = _p0; a.phase = _phase; a.p = _p; }
  cudaDeviceSynchronize();
  if (bbpcg_epilogue(solver(), &a, NULL)) die("bbpcg_epilogue");
}
extern "C" void cuda_project(void) { epilogue(1, 0); }      /* src/cuda_bluebottle.cu:2495-2503 */
extern "C" void cuda_update_p(void) { epilogue(0, 1); }     /* src/cuda_bluebottle.cu:2505-2534 */

extern "C" void bbpcg_dropin_finalize(void) { bbpcg_destroy(g_solver); g_solver = NULL; }
