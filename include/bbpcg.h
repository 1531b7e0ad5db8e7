/* bbpcg.h -- C ABI of the B200-native pressure-Poisson PCG library (libbbpcg.so).
 *
 * Explicit-argument form of the reference path (all paths relative to /root/reference):
 *   cuda_PP_init_jacobi_preconditioner   src/cuda_solver.cu:31-36      -> bbpcg_set_coefficients
 *   cuda_PP_cg / cuda_PP_cg_noparts      src/cuda_solver.cu:38-300,573-761 -> bbpcg_solve
 *   mpi_cuda_exchange_Gcc                src/mpi_comm.c:257-315         -> bbpcg_exchange_Gcc
 *   domain_read_input (decomp part) + domain_fill   src/domain.c:91-160,918-1486 -> bb_domain_*
 * The reference reads everything from globals; the symbols with the reference's own names and
 * `void f(void)` signatures are exported by the companion drop-in layer (include/bb_dropin.h),
 * which forwards to the functions below.
 *
 * Conventions: plain pointers and sizes, no C++ or torch types.  Every `const real *` /
 * `const int *` array argument of the solver calls is a DEVICE pointer in the reference's
 * ghosted layout (include/bb_grid.h); the caller owns it.  All functions return 0 on success,
 * a negative BBPCG_E* code on failure (bbpcg_last_error() has the text).  There is no CPU
 * path: without a CUDA device every compute entry point fails with BBPCG_ECUDA.
 *
 * Process model (reference: one MPI rank per GPU, src/mpi_comm.c:42-58): one solver object
 * per rank/GPU; bbpcg_solve, bbpcg_set_coefficients and bbpcg_exchange_Gcc are COLLECTIVE over
 * the ranks attached with bbpcg_comm_import().  Work is issued on a private non-blocking
 * stream and is complete (host-synchronised) on return.
 */
#ifndef BBPCG_H
#define BBPCG_H

#include <stddef.h>
#include "bb_grid.h"

#ifdef __cplusplus
extern "C" {
#endif

#define BBPCG_OK        0
#define BBPCG_EINVAL   (-1)
#define BBPCG_ECUDA    (-2)
#define BBPCG_ENOMEM   (-3)
#define BBPCG_ECOMM    (-4)
#define BBPCG_EIO      (-5)

/* solve status, mirrors the exits of cuda_PP_cg (src/cuda_solver.cu:178-189,235-279) */
#define BBPCG_CONVERGED   0   /* (r,z) <= pp_residual^2 (b,b)                               */
#define BBPCG_TINY_RHS    1   /* (b,b) < (1e-8)^2: phi = 0, 0 iterations                    */
#define BBPCG_MAXITER     2   /* pp_max_iter+1 iterations without convergence (reference: exit) */
#define BBPCG_NAN         3   /* (r,z) is NaN (reference: exit)                             */
#define BBPCG_COMM_TIMEOUT 4  /* a peer rank never arrived in an in-kernel collective: the call returns BBPCG_ECOMM and
                                 the solver objects of ALL ranks are dead (destroy + re-create); see option comm_timeout_ms */

#define BBPCG_MAX_RANKS 16
#define BBPCG_BLOB_BYTES 256  /* size of one rank's bbpcg_comm_export() record */

typedef struct bbpcg_solver bbpcg_solver;

typedef struct bbpcg_result {
  int    status;     /* BBPCG_CONVERGED ...                                                  */
  int    niter;      /* q at exit, as passed to recorder_PP (src/cuda_solver.cu:239)         */
  double resid;      /* sqrt((r,z)) / sqrt((b,b)) at exit                                    */
  double sp_rhs;     /* (b,b)                                                                */
  double sp_rq0;     /* initial (r,z)                                                        */
  double ms_setup;   /* device time: rhs + init (CUDA events)                                */
  double ms_iter;    /* device time: the iteration loop only                                 */
  double ms_total;   /* device time: whole call                                              */
  long long launches;/* kernels launched by this call                                        */
} bbpcg_result;

/* ---- host-side decomposition contract (no GPU needed) -------------------------------------
 * bb_domain_read: parse flow.config (the keys this path needs: GLOBAL DOMAIN, (In,Jn,Kn),
 * rho_f, pp_max_iter, pp_residual, the six bc.p* lines) and decomp.config (record grammar of
 * src/domain.c:138-159), then fill every index range like domain_fill.  *dom_out is malloc'ed
 * with DOM_out->S3 entries; free with bb_domain_free. */
typedef struct bb_flow_params {
  double rho_f;
  double pp_residual;
  int    pp_max_iter;
} bb_flow_params;

int  bb_domain_read(const char *flow_config, const char *decomp_config, dom_struct *DOM_out,
                    dom_struct **dom_out, bb_pressure_bc *bc_out, bb_flow_params *params_out);
/* dom[] entries must carry I,J,K and xs,xe,xn,ys,ye,yn,zs,ze,zn; DOM must carry the global
 * extents and In,Jn,Kn.  Fills everything else (neighbours from the pressure BCs). */
int  bb_domain_fill(dom_struct *DOM, dom_struct *dom, const bb_pressure_bc *bc);
/* equal splits as tools/src/decomp_reader.c:112-129 writes them */
int  bb_domain_split(dom_struct *DOM, dom_struct *dom);
int  bb_domain_write_decomp(const char *path, const dom_struct *DOM, const dom_struct *dom, int prec);
void bb_domain_free(dom_struct *dom);

/* ---- restart files as fixtures (no GPU needed) ---------------------------------------------
 * Reader for the per-rank binary `restart.config-<rank>` files Bluebottle writes (out_restart, src/domain.c:3005-3092;
 * read back by in_restart, :3094-3260), so that the state of a production run -- u*, v*, w*, flags, phase, phi, p, p0 --
 * can be replayed through this library without MPI.  Layout (sequential fwrite's, no padding): ttime, dt0, dt (real),
 * stepnum, rec_vtk_stepnum_out (int), three output times (real); then u, u0, diff0_u, conv0_u, diff_u, conv_u, u_star on
 * Gfx s3b, the same seven for v (Gfy) and w (Gfz); p, phi, p0 (real) and phase, phase_shell (int) on Gcc s3b; flag_u,
 * flag_v, flag_w (int); nparts_subdom (int); the particle structs and scalar-field arrays that follow are not read.
 * Every array is malloc'ed in the reference's ghosted layout (include/bb_grid.h); free with bb_restart_free. */
typedef struct bb_restart {
  real ttime, dt0, dt;
  int  stepnum, rec_vtk_stepnum_out;
  real rec_cgns_flow_ttime_out, rec_cgns_part_ttime_out, rec_vtk_ttime_out;
  real *u, *v, *w;                 /* Gfx / Gfy / Gfz s3b */
  real *u_star, *v_star, *w_star;
  real *p, *phi, *p0;              /* Gcc s3b */
  int  *phase, *phase_shell;       /* Gcc s3b */
  int  *flag_u, *flag_v, *flag_w;  /* Gfx / Gfy / Gfz s3b */
  int  nparts_subdom;
} bb_restart;

/* "<dir>/restart.config-<rank>" with the rank zero-padded to floor(log10(S3-1))+1 digits (src/domain.c:3008-3017) */
int  bb_restart_path(char *out, size_t cap, const char *dir, int rank, int S3);
int  bb_restart_read(const char *path, const dom_struct *dom_rank, bb_restart *out);
void bb_restart_free(bb_restart *r);

/* ---- record files (no GPU needed) ----------------------------------------------------------
 * Writers of the per-solve record lines, byte-compatible with recorder_PP_init / recorder_PP (src/recorder.c:157-221:
 * <root_dir>/record/solver_expd.rec) and recorder_PP_init_timed / recorder_PP_timed (:223-336: solver_expd_timed.rec with
 * the eight segment columns spmv, ip1, ar1, up1, ip2, ar2, up2, mpi).  The reference averages the times over the ranks
 * (MPI_Allreduce, :193-194) before rank 0 writes: pass averaged values and call on rank 0 only.  A missing file is
 * created with its header first (:201-204); a line is "\n" + fields, so the file never ends in a newline. */
int  bb_recorder_PP_init(const char *root_dir, const char *name);
int  bb_recorder_PP(const char *root_dir, const char *name, int stepnum, real ttime, real dt, int niter, real resid, real etime);
int  bb_recorder_PP_init_timed(const char *root_dir, const char *name);
int  bb_recorder_PP_timed(const char *root_dir, const char *name, int stepnum, real ttime, real dt, int niter, real resid,
                          real etime, const real seg[8]);

/* ---- solver object ------------------------------------------------------------------------ */
/* dom_rank: this rank's filled block; DOM: the global domain; bc: pressure BC types.
 * device: CUDA ordinal, or -1 for the current device.  Allocates the private workspace
 * (5 padded vectors + coefficient masks, ~41 B per cell) in ONE device allocation. */
int  bbpcg_create(bbpcg_solver **out, const dom_struct *dom_rank, const dom_struct *DOM,
                  const bb_pressure_bc *bc, int device);
void bbpcg_destroy(bbpcg_solver *s);

/* Multi-GPU attach (replaces mpi_dom_comm_init / create_windows, src/mpi_comm.c:81-251).
 * Each rank exports BBPCG_BLOB_BYTES describing its workspace (CUDA IPC handle + layout); the
 * host program all-gathers the records in rank order by whatever means it has (MPI_Allgather
 * in Bluebottle, torch.distributed in the harness) and hands the concatenation to
 * bbpcg_comm_import on every rank.  nranks == 1 needs neither call. */
int  bbpcg_comm_export(bbpcg_solver *s, void *blob /* BBPCG_BLOB_BYTES */);
int  bbpcg_comm_import(bbpcg_solver *s, const void *all_blobs, int nranks);

/* = cuda_PP_init_jacobi_preconditioner: digest the face flags (Gfx/Gfy/Gfz s3b ints) and,
 * if phase != NULL, the particle phase (Gcc s3b ints) into the solver's private coefficient
 * masks (the Jacobi diagonal is recomputed from them on the fly, no invM vector is stored).
 * Must be called again whenever flags/phase change (src/bluebottle.c:390-395). */
int  bbpcg_set_coefficients(bbpcg_solver *s, const int *flag_u, const int *flag_v,
                            const int *flag_w, const int *phase);

/* ---- coefficient producers ---------------------------------------------------------------------
 * = cuda_build_cages() (src/cuda_particle.cu:1516-1646) FUSED with the mask digestion of cuda_PP_init_jacobi_preconditioner:
 * from this rank's particle list (centres and radii in global coordinates; the reference's `_parts`, ghost particles
 * included) it fills phase / phase_shell (Gcc s3b ints: local particle index or -1; 0 on a particle's surface shell, else 1)
 * and flag_u / flag_v / flag_w (Gfx / Gfy / Gfz s3b ints: 1, -1 on cage faces, 0 on external walls) EXACTLY as the reference
 * does -- its other kernels read them -- and writes the solver's own 1-byte coefficient masks in the same pass, so no
 * bbpcg_set_coefficients call is needed afterwards.  NPARTS == 0 (no particle anywhere): phase arrays are not touched, flags
 * are 1 except on external walls.  4 launches, no host round trip (reference: 4 nparts + ~12 launches and 2 nparts blocking
 * copies).  COLLECTIVE.  All arrays are DEVICE pointers; the particle list is addressed through a strided view so that an
 * array of the reference's part_struct can be passed as is (stride sizeof(part_struct), offsets of x, y, z, r). */
typedef struct bbpcg_parts_view {
  const void *base;                 /* device */
  size_t stride;                    /* bytes between particles */
  size_t off_x, off_y, off_z, off_r;/* byte offsets of the `real` fields inside one particle */
} bbpcg_parts_view;

int  bbpcg_build_cages(bbpcg_solver *s, int NPARTS, int nparts, const bbpcg_parts_view *parts,
                       int *flag_u, int *flag_v, int *flag_w, int *phase, int *phase_shell);

/* Optional callback = the reference's cuda_part_BC_p() (src/cuda_particle.cu:1680), invoked
 * after PP_rhs when use_phase != 0 (src/cuda_solver.cu:128-132).  When NULL and phase_shell !=
 * NULL the library applies that kernel's net effect on rhs (rhs *= (phase<0 && phase_shell),
 * src/particle_kernel.cu:1753) itself. */
typedef void (*bbpcg_part_bc_fn)(void);

typedef struct bbpcg_solve_args {
  const real *u_star, *v_star, *w_star;   /* Gfx / Gfy / Gfz s3b, device                    */
  real       *rhs_p;                       /* Gcc s3b scratch, device (reference: _rhs_p)    */
  real       *phi;                         /* Gcc s3b, device: OUT, interior written        */
  const int  *phase, *phase_shell;         /* Gcc s3b, device; only read if use_phase        */
  real        rho_f, dt;
  real        pp_residual;
  int         pp_max_iter;
  int         use_phase;                   /* 0: cuda_PP_cg_noparts   1: cuda_PP_cg, NPARTS>0 */
  int         fixed_iters;                 /* >0: run exactly this many iterations, no stop
                                              test (benchmark mode; status CONVERGED)        */
  bbpcg_part_bc_fn part_bc;                /* may be NULL                                    */
  int         no_refine;                   /* 1: skip coeffs_refine, as the reference does on a rank
                                              that holds no particle (nparts == 0 while NPARTS > 0,
                                              src/cuda_solver.cu:139-142)                    */
} bbpcg_solve_args;

int  bbpcg_solve(bbpcg_solver *s, const bbpcg_solve_args *args, bbpcg_result *res);

/* Host-buffer form used by the end-to-end benchmark and by callers without device arrays:
 * copies u*,v*,w* host->device, solves, copies phi (s3b) device->host.  Same semantics. */
int  bbpcg_solve_host(bbpcg_solver *s, const real *u_star_h, const real *v_star_h,
                      const real *w_star_h, real *phi_h, real rho_f, real dt, real pp_residual,
                      int pp_max_iter, int fixed_iters, bbpcg_result *res);

/* (r,z) after every iteration of the last solve: out[0] = initial, out[q] = iteration q.
 * Returns the number of entries written (<= cap). */
int  bbpcg_history(bbpcg_solver *s, double *out, int cap);

/* = mpi_cuda_exchange_Gcc(array): fill the ghost faces of a caller-owned Gcc s3b device array
 * from the neighbouring blocks (periodic wrap included; faces only, no edges/corners). */
int  bbpcg_exchange_Gcc(bbpcg_solver *s, real *array);

/* The same transport for the three face grids: = mpi_cuda_exchange_Gfx / _Gfy / _Gfz (src/mpi_comm.c:317-405)
 * on a caller-owned Gfx / Gfy / Gfz s3b device array.  A face grid shares its block-boundary face with the
 * neighbour (src/domain.c:1292-1301), so along its own normal the planes _ie-1 / _is+1 travel
 * (src/bluebottle_kernel.cu:782-815,916-948,1048-1081); faces only, no edges/corners.  COLLECTIVE. */
#define BBPCG_GCC 0
#define BBPCG_GFX 1
#define BBPCG_GFY 2
#define BBPCG_GFZ 3
int  bbpcg_exchange(bbpcg_solver *s, real *array, int grid);

/* ---- solve epilogue (what src/bluebottle.c:233-256 runs on phi right after the solve) -------
 * = cuda_dom_BC_p(array) (src/cuda_bluebottle.cu:2536-2589): on every face of this block that has
 * no neighbour and whose pressure BC is NEUMANN, ghost = adjacent interior cell (faces only). */
int  bbpcg_dom_BC_p(bbpcg_solver *s, real *array);

typedef struct bbpcg_epilogue_args {
  const real *u_star, *v_star, *w_star;   /* Gfx / Gfy / Gfz s3b, device                           */
  const int  *flag_u, *flag_v, *flag_w;   /* Gfx / Gfy / Gfz s3b, device                           */
  real       *phi;                        /* Gcc s3b, device: interior from the solve               */
  real       *u, *v, *w;                  /* OUT (Gf?._is.._ie faces); NULL u: skip cuda_project    */
  const real *p0;                         /* Gcc s3b, device                                        */
  const int  *phase;                      /* Gcc s3b, device (all -1 without particles)             */
  real       *p;                          /* OUT interior; NULL: skip cuda_update_p                 */
  real        rho_f, dt;
  int         phi_ghosts_valid;           /* 0: run mpi_cuda_exchange_Gcc(phi) + cuda_dom_BC_p(phi)
                                             first (bluebottle.c:233-234); 1: the caller did        */
} bbpcg_epilogue_args;

/* = [mpi_cuda_exchange_Gcc(phi); cuda_dom_BC_p(phi);] cuda_project(); cuda_update_p()
 * (src/cuda_bluebottle.cu:2495-2534) in ONE pass over the block plus the mean subtraction:
 *   u = u* - dt/rho_f |flag_u| (phi_C - phi_W)/dx  (same for v, w; src/bluebottle_kernel.cu:2303-2355)
 *   p = (phase < 0)(p0 + phi) - mean over all ranks  (src/bluebottle_kernel.cu:2396, cuda_bluebottle.cu:2519-2533)
 * Either half may be skipped (u == NULL / p == NULL).  COLLECTIVE when p != NULL or the exchange runs.
 * The velocity BCs the reference applies between the two halves (bluebottle.c:243-248) touch neither
 * phi, p0, phase nor p, so running both halves together gives the same state.
 * ms_out (may be NULL): device time of the call (CUDA events). */
int  bbpcg_epilogue(bbpcg_solver *s, const bbpcg_epilogue_args *args, double *ms_out);

/* ---- solve prologue --------------------------------------------------------------------------
 * = cuda_solvability() (src/cuda_bluebottle.cu:2313-2492): the net flux of u* through the six faces of the GLOBAL domain
 * (face sums x face area, summed over all ranks) is removed from the outflow plane `out_plane` (WEST 0, EAST 1, SOUTH 2,
 * NORTH 3, BOTTOM 4, TOP 5; src/bluebottle.h:365-425) or, for HOMOGENEOUS (10, :353), half of each axis' imbalance from
 * both planes of that axis, so that the Poisson right-hand side sums to zero.  u*, v*, w* are modified in place on the
 * boundary planes only.  eps_out (may be NULL) receives the three per-axis imbalances (host).  COLLECTIVE. */
#define BBPCG_HOMOGENEOUS 10
int  bbpcg_solvability(bbpcg_solver *s, real *u_star, real *v_star, real *w_star, int out_plane, real *eps_out);

/* = cuda_dom_BC_star() (src/cuda_bluebottle.cu:2111-2311): the velocity boundary-condition table on u*, v*, w* -- on every
 * face of this block without a neighbour, per component: DIRICHLET (wall-normal component: ghost = 2 bc - inner face, wall
 * face = bc; tangential: ghost = 8/3 bc - 2 a1 + 1/3 a2) or NEUMANN (ghost = first value); PERIODIC / PRECURSOR entries are
 * left alone, faces only (BC_{u,v,w}_{W,E,S,N,B,T}_{D,N}, src/bluebottle_kernel.cu:104-598).  Three launches (one per
 * axis, W before E etc. inside a thread) instead of up to 18.  Not collective. */
int  bbpcg_dom_BC_star(bbpcg_solver *s, real *u_star, real *v_star, real *w_star, const bb_velocity_bc *vbc);

/* The particle-free prologue of the pressure solve, src/bluebottle.c:213-225, in ONE call with one host synchronisation:
 *   cuda_dom_BC_star; exchange Gfx/Gfy/Gfz;  cuda_solvability;  cuda_dom_BC_star; exchange Gfx/Gfy/Gfz
 * (with particles the caller interleaves cuda_part_BC_star and uses the separate entry points).  COLLECTIVE.
 * ms_out (may be NULL): device time of the call. */
int  bbpcg_prologue(bbpcg_solver *s, real *u_star, real *v_star, real *w_star, const bb_velocity_bc *vbc, int out_plane, double *ms_out);

/* Unit entry points used by the parity tests (same kernels the solve uses). */
int  bbpcg_rhs(bbpcg_solver *s, const real *u_star, const real *v_star, const real *w_star,
               real rho_f, real dt, real *rhs_p);
/* Ap (s3, ghost-free, device) = -A * src (Gcc s3b, device, ghosts as given) */
int  bbpcg_spmv(bbpcg_solver *s, const real *src_s3b, real *Ap_s3, int use_phase);

/* tuning / introspection.  Options (defaults are the measured best; none changes a result beyond the summation order of the
 * dot products, which stays deterministic for a given setting):
 *   comm_timeout_ms  spin limit of the in-kernel rank barriers (default ~70 000; <= 0: wait for ever, like MPI)
 *   ty, kc           tile height (1..8, 0 = automatic) and uniform z-chunk length (0 = automatic) of the iteration kernels
 *   guided, guided_pct, chunk_min   decreasing z-chunk lengths for small blocks (1, 60 % of the even share, >= 4 planes)
 *   pdl              programmatic dependent launch of the iteration kernels (on unless several ranks share one GPU)
 *   epilogue_tiled   1: the 16^3-brick epilogue kernel instead of the plane-marching pair;  epi_chunk: planes per chunk (32)
 *   rhs_tiled, tma_warp, stream_blocks, check_every   see csrc/bbpcg_solver.cu
 *   kernel_timing    1: CUDA events around every iteration kernel (bbpcg_get_info kt_*_ns / kt_*_n; switches PDL off)
 * Info keys: pitch, sm_count, nranks, tile_tx, search_ty, search_grid, search_items, search_kc, search_nbz, pdl, comm_timeout. */
int  bbpcg_set_option(bbpcg_solver *s, const char *key, long long value);
/* the tile / z-chunk planner of the iteration kernels as a pure host function (no GPU needed; CPU tests hold it to its
 * contract): ztab[2c], ztab[2c+1] = first / last plane of chunk c in claim order, plan[5] = { ty, nbx, nby, nbz, planes of
 * chunk 0 }; slots = resident CTAs (2 per SM); the other arguments are the options of the same names (0 = default) */
int  bbpcg_plan_zchunks(int in, int jn, int kn, int slots, int opt_ty, int opt_kc, int guided, int guided_pct, int chunk_min,
                        int *ztab, int ztab_cap, int *plan);
long long bbpcg_get_info(bbpcg_solver *s, const char *key);
const char *bbpcg_last_error(void);
const char *bbpcg_version(void);

#ifdef __cplusplus
}
#endif
#endif /* BBPCG_H */
