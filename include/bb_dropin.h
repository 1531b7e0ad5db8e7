/* bb_dropin.h -- the reference's own entry points, exported by libbbpcg_dropin.so.
 *
 * Link this library INSTEAD OF cuda_solver.o / solver_kernel.o (and the Gcc function of
 * mpi_comm.o) from the reference Makefile:81-91.  The functions keep the reference's names,
 * `void f(void)` signatures, globals-as-arguments convention and print+exit error behaviour
 * (all paths relative to /root/reference):
 *
 *   cuda_PP_init_jacobi_preconditioner   src/bluebottle.h:3073, called src/bluebottle.c:139,394
 *   cuda_PP_cg                           src/bluebottle.h:3085, called src/bluebottle.c:229
 *   cuda_PP_cg_noparts                   src/bluebottle.h:3112, called src/bluebottle.c:231
 *   cuda_PP_cg_timed                     src/bluebottle.h:3098, no caller in the reference
 *   mpi_cuda_exchange_Gcc(real *array)   src/mpi_comm.h:318, 19 call sites outside the solver
 *   mpi_cuda_exchange_Gfx/_Gfy/_Gfz(real *array)   src/mpi_comm.h:335,352,369; 24 call sites in src/bluebottle.c
 *   cuda_solvability                     src/cuda_bluebottle.cu:2313, called src/bluebottle.c:220 (reads the global out_plane)
 *   cuda_dom_BC_star                     src/cuda_bluebottle.cu:2111, called src/bluebottle.c:214,222 (reads the velocity entries of bc)
 *   cuda_build_cages                     src/cuda_particle.cu:1516, called src/bluebottle.c:389 (+ domain / restart set-up); reads _parts
 * and, for the solve epilogue (link instead of the same-named functions of cuda_bluebottle.o):
 *   cuda_dom_BC_p(real *array)           src/cuda_bluebottle.cu:2536, called src/bluebottle.c:234,255
 *   cuda_project                         src/cuda_bluebottle.cu:2495, called src/bluebottle.c:237
 *   cuda_update_p                        src/cuda_bluebottle.cu:2505, called src/bluebottle.c:250
 *
 * They read these globals of the host program (defined in src/bluebottle.c:438-576,
 * src/mpi_comm.c:26-27, src/particle.c:27-28):
 *   dom, DOM, rank, nprocs, bc, rho_f, dt, pp_residual, pp_max_iter, stepnum, ttime,
 *   NPARTS, nparts, _u_star, _v_star, _w_star, _flag_u, _flag_v, _flag_w, _phase,
 *   _phase_shell, _rhs_p, _phi, _parts, out_plane, and for the epilogue _u, _v, _w, _p, _p0
 * and call back into reference code at: cuda_part_BC_p() (src/cuda_particle.cu:1680) and
 * recorder_PP() (src/recorder.c:190).  `_invM,_r_q,_z_q,_p_q,_pb_q,_Apb_q` are NOT used: the
 * library keeps its own padded workspace.
 *
 * Multi-rank bootstrap: on first use the drop-in layer calls the weak hook
 *   int bb_dropin_allgather(const void *send, void *recv, int bytes_per_rank);
 * which the host program implements with MPI_Allgather (INTEGRATION.md shows the three-line
 * definition); a single-rank run needs no hook.
 */
#ifndef BB_DROPIN_H
#define BB_DROPIN_H
#include "bb_grid.h"
#ifdef __cplusplus
extern "C" {
#endif
void cuda_PP_init_jacobi_preconditioner(void);
void cuda_PP_cg(void);
void cuda_PP_cg_noparts(void);
void cuda_PP_cg_timed(void);
void mpi_cuda_exchange_Gcc(real *array);
void mpi_cuda_exchange_Gfx(real *array);
void mpi_cuda_exchange_Gfy(real *array);
void mpi_cuda_exchange_Gfz(real *array);
void cuda_solvability(void);
void cuda_dom_BC_star(void);
void cuda_build_cages(void);
void cuda_dom_BC_p(real *array);
void cuda_project(void);
void cuda_update_p(void);
/* extra: release the workspace before mpi_end(); optional */
void bbpcg_dropin_finalize(void);
/* host-provided (weak) all-gather used once at start-up when nprocs > 1 */
int bb_dropin_allgather(const void *send, void *recv, int bytes_per_rank);
#ifdef __cplusplus
}
#endif
#endif
