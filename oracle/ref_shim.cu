/* oracle/ref_shim.cu -- single-rank host shim around the reference's OWN hot-path objects
 * ("O1").  TEST INFRASTRUCTURE: the resulting oracle/_ref/libbbref.so is a
 * checker and the `--impl reference` arm of bench.py; it is never on the product path.
 *
 * What is linked: /root/reference/src/{solver_kernel,cuda_solver,bluebottle_kernel}.cu,
 * compiled UNMODIFIED where they lie (oracle/Makefile) with -DDOUBLE -DJACOBI for sm_100a.
 * This file supplies what the rest of the Bluebottle program would: the globals those
 * objects reference (definitions mirror bluebottle.c:438-576, mpi_comm.c:26-27,
 * particle.c:27-28, cuda_bluebottle.cu:34-35), a rank-0-of-1 stand-in for MPI
 * (allreduce = identity, self MPI_Put = device-to-device copy), the launch geometry of
 * cuda_blocks_init (cuda_bluebottle.cu:523-581, Gcc part) and the host sequence of
 * mpi_cuda_exchange_Gcc (mpi_comm.c:257-315) driving the reference's own pack/unpack
 * kernels.  No reference source text is copied into this repository.
 */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>
#include <thrust/device_ptr.h>
#include <thrust/reduce.h>

#include "cuda_solver.h"      /* reference header (from -I $(REF)/src): types + kernel prototypes */
#include "cuda_bluebottle.h"  /* pack/unpack kernel prototypes */
#include "cuda_particle.h"    /* cage-building kernel prototypes, part_struct */

/* ---- globals the reference objects link against --------------------------------------- */
__constant__ dom_struct _dom;            /* cuda_bluebottle.cu:34 */
__constant__ bin_struct _bins;           /* cuda_bluebottle.cu (referenced by bluebottle_kernel.o, unused here) */
cuda_blocks_struct blocks;               /* cuda_bluebottle.cu:35 */
dom_struct *dom;  dom_struct DOM;
int rank = 0, nprocs = 1;
int NPARTS = 0, nparts = 0;
real rho_f, dt, pp_residual, ttime;
int pp_max_iter, stepnum;
real *_phi, *_rhs_p, *_r_q, *_z_q, *_p_q, *_pb_q, *_Apb_q, *_invM;
real *_u_star, *_v_star, *_w_star;
int *_flag_u, *_flag_v, *_flag_w, *_phase, *_phase_shell;
static real *_send_Gcc[6], *_recv_Gcc[6];   /* e w n s t b */
/* epilogue (bluebottle.c:233-256): projected velocity, pressures, the BC table (only bc.p* is read) */
real *_u, *_v, *_w, *_p, *_p0;
real nu;
BC bc;

static int  g_niter = -1;
static real g_resid = -1., g_etime = 0.;

extern "C" {
/* MPI, rank 0 of 1 */
int MPI_Allreduce(const void *, void *, int, MPI_Datatype, MPI_Op, MPI_Comm) { return 0; }  /* IN_PLACE, 1 rank */
int MPI_Barrier(MPI_Comm) { return 0; }

/* recorder.c:190-221 -- the solver's only output besides _phi */
void recorder_PP(char *, int niter, real resid, real etime) { g_niter = niter; g_resid = resid; g_etime = etime; }
void recorder_PP_init_timed(char *) {}
void recorder_PP_timed(char *, int niter, real resid, real etime, real, real, real, real, real, real, real, real)
{ g_niter = niter; g_resid = resid; g_etime = etime; }   /* recorder.h:258-270: 8 segment timers */

/* Net effect of cuda_part_BC_p (cuda_particle.cu:1680 -> particle_kernel.cu:1655-1756): the
 * final statement (:1753) zeroes rhs in solid cells and keeps it in fluid cells. */
__global__ void shim_part_BC_p_net(real *rhs, const int *phase, const int *phase_shell, int n)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) rhs[i] = (real)(phase[i] < 0 && phase_shell[i]) * rhs[i];
}
void cuda_part_BC_p(void)
{
  int n = dom[rank].Gcc.s3b;
  shim_part_BC_p_net<<<(n + 255) / 256, 256>>>(_rhs_p, _phase, _phase_shell, n);
}

/* mpi_cuda_exchange_Gcc, mpi_comm.c:257-315, for one rank: a neighbour is either this rank
 * (periodic wrap = self put, domain.c:1151-1159) or MPI_PROC_NULL. */
void mpi_cuda_exchange_Gcc(real *array)
{
  const dom_struct *d = &dom[rank];
  if (d->e != MPI_PROC_NULL) pack_planes_Gcc_east  <<<blocks.Gcc.num_in, blocks.Gcc.dim_in>>>(array, _send_Gcc[0]);
  if (d->w != MPI_PROC_NULL) pack_planes_Gcc_west  <<<blocks.Gcc.num_in, blocks.Gcc.dim_in>>>(array, _send_Gcc[1]);
  if (d->n != MPI_PROC_NULL) pack_planes_Gcc_north <<<blocks.Gcc.num_jn, blocks.Gcc.dim_jn>>>(array, _send_Gcc[2]);
  if (d->s != MPI_PROC_NULL) pack_planes_Gcc_south <<<blocks.Gcc.num_jn, blocks.Gcc.dim_jn>>>(array, _send_Gcc[3]);
  if (d->t != MPI_PROC_NULL) pack_planes_Gcc_top   <<<blocks.Gcc.num_kn, blocks.Gcc.dim_kn>>>(array, _send_Gcc[4]);
  if (d->b != MPI_PROC_NULL) pack_planes_Gcc_bottom<<<blocks.Gcc.num_kn, blocks.Gcc.dim_kn>>>(array, _send_Gcc[5]);
  cudaDeviceSynchronize();
  /* w -> recv_e, e -> recv_w, s -> recv_n, n -> recv_s, b -> recv_t, t -> recv_b (mpi_comm.c:293-306) */
  if (d->w == rank) cudaMemcpy(_recv_Gcc[0], _send_Gcc[1], sizeof(real) * d->Gcc.s2_i, cudaMemcpyDeviceToDevice);
  if (d->e == rank) cudaMemcpy(_recv_Gcc[1], _send_Gcc[0], sizeof(real) * d->Gcc.s2_i, cudaMemcpyDeviceToDevice);
  if (d->s == rank) cudaMemcpy(_recv_Gcc[2], _send_Gcc[3], sizeof(real) * d->Gcc.s2_j, cudaMemcpyDeviceToDevice);
  if (d->n == rank) cudaMemcpy(_recv_Gcc[3], _send_Gcc[2], sizeof(real) * d->Gcc.s2_j, cudaMemcpyDeviceToDevice);
  if (d->b == rank) cudaMemcpy(_recv_Gcc[4], _send_Gcc[5], sizeof(real) * d->Gcc.s2_k, cudaMemcpyDeviceToDevice);
  if (d->t == rank) cudaMemcpy(_recv_Gcc[5], _send_Gcc[4], sizeof(real) * d->Gcc.s2_k, cudaMemcpyDeviceToDevice);
  cudaDeviceSynchronize();
  if (d->e != MPI_PROC_NULL) unpack_planes_Gcc_east  <<<blocks.Gcc.num_in, blocks.Gcc.dim_in>>>(array, _recv_Gcc[0]);
  if (d->w != MPI_PROC_NULL) unpack_planes_Gcc_west  <<<blocks.Gcc.num_in, blocks.Gcc.dim_in>>>(array, _recv_Gcc[1]);
  if (d->n != MPI_PROC_NULL) unpack_planes_Gcc_north <<<blocks.Gcc.num_jn, blocks.Gcc.dim_jn>>>(array, _recv_Gcc[2]);
  if (d->s != MPI_PROC_NULL) unpack_planes_Gcc_south <<<blocks.Gcc.num_jn, blocks.Gcc.dim_jn>>>(array, _recv_Gcc[3]);
  if (d->t != MPI_PROC_NULL) unpack_planes_Gcc_top   <<<blocks.Gcc.num_kn, blocks.Gcc.dim_kn>>>(array, _recv_Gcc[4]);
  if (d->b != MPI_PROC_NULL) unpack_planes_Gcc_bottom<<<blocks.Gcc.num_kn, blocks.Gcc.dim_kn>>>(array, _recv_Gcc[5]);
  cudaDeviceSynchronize();
}
} /* extern "C" */

/* launch geometry for the Gcc kernels: cuda_blocks_init, cuda_bluebottle.cu:523-581 */
static int thr(int n) { return n < MAX_THREADS_DIM ? n : MAX_THREADS_DIM; }
static int nblk(int n, int t) { return (n + t - 1) / t; }
static void shim_blocks_init(const dom_struct *d)
{
  int tx = thr(d->Gcc.in), ty = thr(d->Gcc.jn), tz = thr(d->Gcc.kn);
  int bx = nblk(d->Gcc.in, tx), by = nblk(d->Gcc.jn, ty), bz = nblk(d->Gcc.kn, tz);
  blocks.Gcc.dim_in = dim3(ty, tz); blocks.Gcc.dim_jn = dim3(tz, tx); blocks.Gcc.dim_kn = dim3(tx, ty);
  blocks.Gcc.num_in = dim3(by, bz); blocks.Gcc.num_jn = dim3(bz, bx); blocks.Gcc.num_kn = dim3(bx, by);
  tx = thr(d->Gcc.in + 2); ty = thr(d->Gcc.jn + 2); tz = thr(d->Gcc.kn + 2);
  bx = nblk(d->Gcc.in, tx - 2); by = nblk(d->Gcc.jn, ty - 2); bz = nblk(d->Gcc.kn, tz - 2);
  blocks.Gcc.dim_in_s = dim3(ty, tz); blocks.Gcc.dim_jn_s = dim3(tz, tx); blocks.Gcc.dim_kn_s = dim3(tx, ty);
  blocks.Gcc.num_in_s = dim3(by, bz); blocks.Gcc.num_jn_s = dim3(bz, bx); blocks.Gcc.num_kn_s = dim3(bx, by);
  /* face grids, cuda_bluebottle.cu:583-760: the same rule per grid (project_u/v/w and the Gf? pack/unpack kernels) */
#define SHIM_FACE_BLOCKS(G)                                                                              \
  tx = thr(d->G.in); ty = thr(d->G.jn); tz = thr(d->G.kn);                                               \
  bx = nblk(d->G.in, tx); by = nblk(d->G.jn, ty); bz = nblk(d->G.kn, tz);                                \
  blocks.G.dim_in = dim3(ty, tz); blocks.G.dim_jn = dim3(tz, tx); blocks.G.dim_kn = dim3(tx, ty);        \
  blocks.G.num_in = dim3(by, bz); blocks.G.num_jn = dim3(bz, bx); blocks.G.num_kn = dim3(bx, by);
  SHIM_FACE_BLOCKS(Gfx) SHIM_FACE_BLOCKS(Gfy) SHIM_FACE_BLOCKS(Gfz)
#undef SHIM_FACE_BLOCKS
  /* ghost-inclusive shapes (cuda_bluebottle.cu:762-790 Gcc, :822-850 Gfx, :882-910 Gfy, :942-970 Gfz): zero_rhs_ghost_{i,j,k},
   * reset_flag_*, reset_phases, cage_flag_*, flag_external_* */
#define SHIM_GHOST_BLOCKS(G)                                                                             \
  tx = thr(d->G.inb); ty = thr(d->G.jnb); tz = thr(d->G.knb);                                            \
  bx = nblk(d->G.inb, tx); by = nblk(d->G.jnb, ty); bz = nblk(d->G.knb, tz);                             \
  blocks.G.dim_inb = dim3(ty, tz); blocks.G.dim_jnb = dim3(tz, tx); blocks.G.dim_knb = dim3(tx, ty);     \
  blocks.G.num_inb = dim3(by, bz); blocks.G.num_jnb = dim3(bz, bx); blocks.G.num_knb = dim3(bx, by);
  SHIM_GHOST_BLOCKS(Gcc) SHIM_GHOST_BLOCKS(Gfx) SHIM_GHOST_BLOCKS(Gfy) SHIM_GHOST_BLOCKS(Gfz)
#undef SHIM_GHOST_BLOCKS
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "bbref: %s -> %s\n", #x, cudaGetErrorString(e_)); return -1; } } while (0)

extern "C" {

/* One block (rank 0 of 1); `d` must already be filled (domain_fill semantics). */
int bbref_init(const dom_struct *d, const dom_struct *D)
{
  dom = (dom_struct *)malloc(sizeof(dom_struct));
  memcpy(dom, d, sizeof(dom_struct)); memcpy(&DOM, D, sizeof(dom_struct));
  rank = 0; nprocs = 1;
  CK(cudaMemcpyToSymbol(_dom, dom, sizeof(dom_struct)));          /* cuda_bluebottle.cu:148 */
  shim_blocks_init(dom);
  size_t s3 = d->Gcc.s3, s3b = d->Gcc.s3b;
  /* cuda_dom_malloc_dev: cuda_bluebottle.cu:162-166,222-266,308-319,373 */
  CK(cudaMalloc(&_phi, s3b * sizeof(real)));    CK(cudaMemset(_phi, 0, s3b * sizeof(real)));
  CK(cudaMalloc(&_rhs_p, s3b * sizeof(real)));  CK(cudaMemset(_rhs_p, 0, s3b * sizeof(real)));
  CK(cudaMalloc(&_pb_q, s3b * sizeof(real)));   CK(cudaMemset(_pb_q, 0, s3b * sizeof(real)));
  CK(cudaMalloc(&_r_q, s3 * sizeof(real)));  CK(cudaMalloc(&_z_q, s3 * sizeof(real)));
  CK(cudaMalloc(&_p_q, s3 * sizeof(real)));  CK(cudaMalloc(&_Apb_q, s3 * sizeof(real)));
  CK(cudaMalloc(&_invM, s3 * sizeof(real)));
  CK(cudaMalloc(&_u_star, (size_t)d->Gfx.s3b * sizeof(real)));
  CK(cudaMalloc(&_v_star, (size_t)d->Gfy.s3b * sizeof(real)));
  CK(cudaMalloc(&_w_star, (size_t)d->Gfz.s3b * sizeof(real)));
  CK(cudaMalloc(&_flag_u, (size_t)d->Gfx.s3b * sizeof(int)));
  CK(cudaMalloc(&_flag_v, (size_t)d->Gfy.s3b * sizeof(int)));
  CK(cudaMalloc(&_flag_w, (size_t)d->Gfz.s3b * sizeof(int)));
  CK(cudaMalloc(&_phase, s3b * sizeof(int)));  CK(cudaMalloc(&_phase_shell, s3b * sizeof(int)));
  for (int f = 0; f < 6; f++) {
    size_t n = (f < 2) ? d->Gcc.s2_i : (f < 4) ? d->Gcc.s2_j : d->Gcc.s2_k;
    CK(cudaMalloc(&_send_Gcc[f], n * sizeof(real)));  CK(cudaMalloc(&_recv_Gcc[f], n * sizeof(real)));
  }
  CK(cudaMalloc(&_u, (size_t)d->Gfx.s3b * sizeof(real)));  CK(cudaMemset(_u, 0, (size_t)d->Gfx.s3b * sizeof(real)));
  CK(cudaMalloc(&_v, (size_t)d->Gfy.s3b * sizeof(real)));  CK(cudaMemset(_v, 0, (size_t)d->Gfy.s3b * sizeof(real)));
  CK(cudaMalloc(&_w, (size_t)d->Gfz.s3b * sizeof(real)));  CK(cudaMemset(_w, 0, (size_t)d->Gfz.s3b * sizeof(real)));
  CK(cudaMalloc(&_p, s3b * sizeof(real)));   CK(cudaMemset(_p, 0, s3b * sizeof(real)));
  CK(cudaMalloc(&_p0, s3b * sizeof(real)));  CK(cudaMemset(_p0, 0, s3b * sizeof(real)));
  return 0;
}

int bbref_set_inputs(const int *flag_u, const int *flag_v, const int *flag_w, const int *phase,
                     const int *phase_shell, const real *u, const real *v, const real *w, int n_parts)
{
  const dom_struct *d = &dom[rank];
  CK(cudaMemcpy(_flag_u, flag_u, (size_t)d->Gfx.s3b * sizeof(int), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(_flag_v, flag_v, (size_t)d->Gfy.s3b * sizeof(int), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(_flag_w, flag_w, (size_t)d->Gfz.s3b * sizeof(int), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(_phase, phase, (size_t)d->Gcc.s3b * sizeof(int), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(_phase_shell, phase_shell, (size_t)d->Gcc.s3b * sizeof(int), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(_u_star, u, (size_t)d->Gfx.s3b * sizeof(real), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(_v_star, v, (size_t)d->Gfy.s3b * sizeof(real), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(_w_star, w, (size_t)d->Gfz.s3b * sizeof(real), cudaMemcpyHostToDevice));
  NPARTS = nparts = n_parts;
  return 0;
}

/* device-resident inputs (bench: inputs generated on the GPU box by the harness) */
int bbref_set_inputs_dev(const int *flag_u, const int *flag_v, const int *flag_w, const int *phase,
                         const int *phase_shell, const real *u, const real *v, const real *w, int n_parts)
{
  const dom_struct *d = &dom[rank];
  CK(cudaMemcpy(_flag_u, flag_u, (size_t)d->Gfx.s3b * sizeof(int), cudaMemcpyDeviceToDevice));
  CK(cudaMemcpy(_flag_v, flag_v, (size_t)d->Gfy.s3b * sizeof(int), cudaMemcpyDeviceToDevice));
  CK(cudaMemcpy(_flag_w, flag_w, (size_t)d->Gfz.s3b * sizeof(int), cudaMemcpyDeviceToDevice));
  CK(cudaMemcpy(_phase, phase, (size_t)d->Gcc.s3b * sizeof(int), cudaMemcpyDeviceToDevice));
  CK(cudaMemcpy(_phase_shell, phase_shell, (size_t)d->Gcc.s3b * sizeof(int), cudaMemcpyDeviceToDevice));
  CK(cudaMemcpy(_u_star, u, (size_t)d->Gfx.s3b * sizeof(real), cudaMemcpyDeviceToDevice));
  CK(cudaMemcpy(_v_star, v, (size_t)d->Gfy.s3b * sizeof(real), cudaMemcpyDeviceToDevice));
  CK(cudaMemcpy(_w_star, w, (size_t)d->Gfz.s3b * sizeof(real), cudaMemcpyDeviceToDevice));
  NPARTS = nparts = n_parts;
  return 0;
}

/* Runs the reference entry points exactly as bluebottle.c:139 / :228-232 do.
 * Returns CUDA-event milliseconds of the solve call; niter/resid as given to recorder_PP.
 * NB: on non-convergence the reference calls exit(EXIT_FAILURE) (cuda_solver.cu:271-279). */
int bbref_solve(real rho_f_, real dt_, real pp_residual_, int pp_max_iter_, int use_parts,
                int *niter, real *resid, float *ms)
{
  rho_f = rho_f_; dt = dt_; pp_residual = pp_residual_; pp_max_iter = pp_max_iter_;
  stepnum = 0; ttime = 0.;
  g_niter = -1; g_resid = -1.;
  cuda_PP_init_jacobi_preconditioner();
  CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  CK(cudaEventRecord(e0, 0));
  if (use_parts) cuda_PP_cg(); else cuda_PP_cg_noparts();
  CK(cudaEventRecord(e1, 0));
  CK(cudaEventSynchronize(e1));
  CK(cudaGetLastError());
  if (ms) CK(cudaEventElapsedTime(ms, e0, e1));
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  if (niter) *niter = g_niter;
  if (resid) *resid = g_resid;
  return 0;
}

/* end-to-end form for bench.py's reference arm: u*, v*, w* from (pinned) host buffers, phi back
 * to the host; *ms covers copies + solve (CUDA events on the default stream the reference uses) */
int bbref_solve_host(const real *u_h, const real *v_h, const real *w_h, real *phi_h, real rho_f_, real dt_,
                     real pp_residual_, int pp_max_iter_, int use_parts, int *niter, real *resid, float *ms)
{
  const dom_struct *d = &dom[rank];
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  CK(cudaEventRecord(e0, 0));
  CK(cudaMemcpyAsync(_u_star, u_h, (size_t)d->Gfx.s3b * sizeof(real), cudaMemcpyHostToDevice, 0));
  CK(cudaMemcpyAsync(_v_star, v_h, (size_t)d->Gfy.s3b * sizeof(real), cudaMemcpyHostToDevice, 0));
  CK(cudaMemcpyAsync(_w_star, w_h, (size_t)d->Gfz.s3b * sizeof(real), cudaMemcpyHostToDevice, 0));
  float ms_solve = 0.f;
  if (bbref_solve(rho_f_, dt_, pp_residual_, pp_max_iter_, use_parts, niter, resid, &ms_solve)) return -1;
  CK(cudaMemcpyAsync(phi_h, _phi, (size_t)d->Gcc.s3b * sizeof(real), cudaMemcpyDeviceToHost, 0));
  CK(cudaEventRecord(e1, 0));
  CK(cudaEventSynchronize(e1));
  if (ms) CK(cudaEventElapsedTime(ms, e0, e1));
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  return 0;
}

int bbref_get(int which, real *host)   /* 0: phi (s3b)  1: rhs_p (s3b)  2: invM (s3)  3: Apb_q (s3)  4: pb_q (s3b) */
{
  const dom_struct *d = &dom[rank];
  const real *src = which == 0 ? _phi : which == 1 ? _rhs_p : which == 2 ? _invM : which == 3 ? _Apb_q : _pb_q;
  size_t n = (which == 2 || which == 3) ? d->Gcc.s3 : d->Gcc.s3b;
  CK(cudaMemcpy(host, src, n * sizeof(real), cudaMemcpyDeviceToHost));
  return 0;
}

/* one reference SpMV on a host-provided ghosted vector (unit parity of the operator) */
int bbref_spmv(const real *pb_host, int use_parts, real *Ap_host)
{
  const dom_struct *d = &dom[rank];
  CK(cudaMemcpy(_pb_q, pb_host, (size_t)d->Gcc.s3b * sizeof(real), cudaMemcpyHostToDevice));
  if (use_parts) PP_spmv_shared_load<<<blocks.Gcc.num_kn_s, blocks.Gcc.dim_kn_s>>>(_flag_u, _flag_v, _flag_w, _pb_q, _Apb_q, _phase);
  else PP_spmv_shared_load_noparts<<<blocks.Gcc.num_kn_s, blocks.Gcc.dim_kn_s>>>(_flag_u, _flag_v, _flag_w, _pb_q, _Apb_q);
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(Ap_host, _Apb_q, (size_t)d->Gcc.s3 * sizeof(real), cudaMemcpyDeviceToHost));
  return 0;
}

/* the reference halo exchange on a host-provided ghosted vector (cuda_BC_test_periodic analogue) */
int bbref_exchange(real *arr_host)
{
  const dom_struct *d = &dom[rank];
  CK(cudaMemcpy(_pb_q, arr_host, (size_t)d->Gcc.s3b * sizeof(real), cudaMemcpyHostToDevice));
  mpi_cuda_exchange_Gcc(_pb_q);
  CK(cudaMemcpy(arr_host, _pb_q, (size_t)d->Gcc.s3b * sizeof(real), cudaMemcpyDeviceToHost));
  return 0;
}

/* ---- the solve epilogue with the reference's own kernels -----------------------------------------
 * Host sequences restated from cuda_dom_BC_p (cuda_bluebottle.cu:2536-2589), cuda_project (:2495-2503) and
 * cuda_update_p (:2505-2534) -- cuda_bluebottle.cu itself is not linked (it drags in the whole program's
 * globals); every kernel launched here is the reference's, from bluebottle_kernel.cu. */
static void shim_dom_BC_p(real *array)
{
  const dom_struct *d = &dom[rank];
  if (d->w == MPI_PROC_NULL && bc.pW == NEUMANN) BC_p_W_N<<<blocks.Gcc.num_in, blocks.Gcc.dim_in>>>(array);
  if (d->e == MPI_PROC_NULL && bc.pE == NEUMANN) BC_p_E_N<<<blocks.Gcc.num_in, blocks.Gcc.dim_in>>>(array);
  if (d->s == MPI_PROC_NULL && bc.pS == NEUMANN) BC_p_S_N<<<blocks.Gcc.num_jn, blocks.Gcc.dim_jn>>>(array);
  if (d->n == MPI_PROC_NULL && bc.pN == NEUMANN) BC_p_N_N<<<blocks.Gcc.num_jn, blocks.Gcc.dim_jn>>>(array);
  if (d->b == MPI_PROC_NULL && bc.pB == NEUMANN) BC_p_B_N<<<blocks.Gcc.num_kn, blocks.Gcc.dim_kn>>>(array);
  if (d->t == MPI_PROC_NULL && bc.pT == NEUMANN) BC_p_T_N<<<blocks.Gcc.num_kn, blocks.Gcc.dim_kn>>>(array);
}

static void shim_project(void)
{
  project_u<<<blocks.Gfx.num_in, blocks.Gfx.dim_in>>>(_u_star, _phi, rho_f, dt, _u, 1. / dom[rank].dx, _flag_u);
  project_v<<<blocks.Gfy.num_jn, blocks.Gfy.dim_jn>>>(_v_star, _phi, rho_f, dt, _v, 1. / dom[rank].dy, _flag_v);
  project_w<<<blocks.Gfz.num_kn, blocks.Gfz.dim_kn>>>(_w_star, _phi, rho_f, dt, _w, 1. / dom[rank].dz, _flag_w);
}

static int shim_update_p(void)
{
  real *_Lp, *_p_mean;
  CK(cudaMalloc((void **)&_Lp, sizeof(real) * dom[rank].Gcc.s3b));
  update_p_laplacian<<<blocks.Gcc.num_kn, blocks.Gcc.dim_kn>>>(_Lp, _phi);
  update_p<<<blocks.Gcc.num_kn, blocks.Gcc.dim_kn>>>(_Lp, _p0, _p, _phi, nu, dt, _phase);
  CK(cudaFree(_Lp));
  CK(cudaMalloc((void **)&_p_mean, sizeof(real) * dom[rank].Gcc.s3));
  copy_p_p_noghost<<<blocks.Gcc.num_kn, blocks.Gcc.dim_kn>>>(_p_mean, _p);
  thrust::device_ptr<real> t_p_mean(_p_mean);
  real pmean = thrust::reduce(t_p_mean, t_p_mean + dom[rank].Gcc.s3, 0., thrust::plus<real>());
  pmean /= (real)DOM.Gcc.s3;                          /* 1 rank: the MPI_Allreduce is the identity */
  CK(cudaFree(_p_mean));
  forcing_add_c_const<<<blocks.Gcc.num_kn, blocks.Gcc.dim_kn>>>(-pmean, _p);
  return 0;
}

/* bluebottle.c:233-256 restricted to what touches phi / p.  phi_h != NULL replaces the solver's _phi interior+ghosts
 * first (a seeded test vector); p0_h is the previous pressure.  pbc[6] = bc.pW,pE,pS,pN,pB,pT.  Outputs to the host:
 * u, v, w (Gf? s3b), p and the ghost-filled phi (Gcc s3b).  *ms = CUDA-event time of the sequence. */
int bbref_epilogue(const real *phi_h, const real *p0_h, const int *pbc, real rho_f_, real dt_, real nu_,
                   real *u_h, real *v_h, real *w_h, real *p_h, real *phi_out_h, float *ms)
{
  const dom_struct *d = &dom[rank];
  const size_t s3b = d->Gcc.s3b;
  rho_f = rho_f_; dt = dt_; nu = nu_;
  memset(&bc, 0, sizeof(bc));
  bc.pW = pbc[0]; bc.pE = pbc[1]; bc.pS = pbc[2]; bc.pN = pbc[3]; bc.pB = pbc[4]; bc.pT = pbc[5];
  if (phi_h) CK(cudaMemcpy(_phi, phi_h, s3b * sizeof(real), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(_p0, p0_h, s3b * sizeof(real), cudaMemcpyHostToDevice));
  CK(cudaMemset(_p, 0, s3b * sizeof(real)));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  CK(cudaEventRecord(e0, 0));
  mpi_cuda_exchange_Gcc(_phi);                        /* bluebottle.c:233 */
  shim_dom_BC_p(_phi);                                /* :234 */
  shim_project();                                     /* :237 */
  if (shim_update_p()) return -1;                     /* :250 */
  CK(cudaEventRecord(e1, 0));
  CK(cudaEventSynchronize(e1));
  CK(cudaGetLastError());
  if (ms) CK(cudaEventElapsedTime(ms, e0, e1));
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  if (u_h) CK(cudaMemcpy(u_h, _u, (size_t)d->Gfx.s3b * sizeof(real), cudaMemcpyDeviceToHost));
  if (v_h) CK(cudaMemcpy(v_h, _v, (size_t)d->Gfy.s3b * sizeof(real), cudaMemcpyDeviceToHost));
  if (w_h) CK(cudaMemcpy(w_h, _w, (size_t)d->Gfz.s3b * sizeof(real), cudaMemcpyDeviceToHost));
  if (p_h) CK(cudaMemcpy(p_h, _p, s3b * sizeof(real), cudaMemcpyDeviceToHost));
  if (phi_out_h) CK(cudaMemcpy(phi_out_h, _phi, s3b * sizeof(real), cudaMemcpyDeviceToHost));
  return 0;
}

/* mpi_cuda_exchange_Gfx / _Gfy / _Gfz (mpi_comm.c:317-405) for one rank, with the reference's own pack / unpack kernels
 * (cuda_pack_planes_Gf?, cuda_bluebottle.cu:1576-1651; cuda_unpack_planes_Gf?, :1677-1752) and a device-to-device
 * copy for the self MPI_Put.  grid: 1 Gfx, 2 Gfy, 3 Gfz.  The array travels host -> device -> host. */
typedef void (*shim_plane_kernel)(real *, real *);
int bbref_exchange_face(real *arr_host, int grid)
{
  const dom_struct *d = &dom[rank];
  const grid_info *g = grid == 1 ? &d->Gfx : grid == 2 ? &d->Gfy : &d->Gfz;
  static const shim_plane_kernel pack[3][6] = {
    { pack_planes_Gfx_east, pack_planes_Gfx_west, pack_planes_Gfx_north, pack_planes_Gfx_south, pack_planes_Gfx_top, pack_planes_Gfx_bottom },
    { pack_planes_Gfy_east, pack_planes_Gfy_west, pack_planes_Gfy_north, pack_planes_Gfy_south, pack_planes_Gfy_top, pack_planes_Gfy_bottom },
    { pack_planes_Gfz_east, pack_planes_Gfz_west, pack_planes_Gfz_north, pack_planes_Gfz_south, pack_planes_Gfz_top, pack_planes_Gfz_bottom } };
  static const shim_plane_kernel unpack[3][6] = {
    { unpack_planes_Gfx_east, unpack_planes_Gfx_west, unpack_planes_Gfx_north, unpack_planes_Gfx_south, unpack_planes_Gfx_top, unpack_planes_Gfx_bottom },
    { unpack_planes_Gfy_east, unpack_planes_Gfy_west, unpack_planes_Gfy_north, unpack_planes_Gfy_south, unpack_planes_Gfy_top, unpack_planes_Gfy_bottom },
    { unpack_planes_Gfz_east, unpack_planes_Gfz_west, unpack_planes_Gfz_north, unpack_planes_Gfz_south, unpack_planes_Gfz_top, unpack_planes_Gfz_bottom } };
  const cuda_blocks_info *bi = grid == 1 ? &blocks.Gfx : grid == 2 ? &blocks.Gfy : &blocks.Gfz;
  const dim3 num[6] = { bi->num_in, bi->num_in, bi->num_jn, bi->num_jn, bi->num_kn, bi->num_kn };
  const dim3 dim[6] = { bi->dim_in, bi->dim_in, bi->dim_jn, bi->dim_jn, bi->dim_kn, bi->dim_kn };
  const int nbr[6] = { d->e, d->w, d->n, d->s, d->t, d->b };
  const int opp[6] = { 1, 0, 3, 2, 5, 4 };
  const size_t fsz[6] = { (size_t)g->s2_i, (size_t)g->s2_i, (size_t)g->s2_j, (size_t)g->s2_j, (size_t)g->s2_k, (size_t)g->s2_k };
  real *arr, *snd[6], *rcv[6];
  CK(cudaMalloc(&arr, (size_t)g->s3b * sizeof(real)));
  CK(cudaMemcpy(arr, arr_host, (size_t)g->s3b * sizeof(real), cudaMemcpyHostToDevice));
  for (int f = 0; f < 6; f++) { CK(cudaMalloc(&snd[f], fsz[f] * sizeof(real))); CK(cudaMalloc(&rcv[f], fsz[f] * sizeof(real))); }
  for (int f = 0; f < 6; f++) if (nbr[f] != MPI_PROC_NULL) pack[grid - 1][f]<<<num[f], dim[f]>>>(arr, snd[f]);
  CK(cudaDeviceSynchronize());
  /* w -> the west neighbour's recv_e, e -> recv_w, ... (mpi_comm.c:326-343); the only possible neighbour is this rank */
  for (int f = 0; f < 6; f++) if (nbr[f] == rank) CK(cudaMemcpy(rcv[opp[f]], snd[f], fsz[f] * sizeof(real), cudaMemcpyDeviceToDevice));
  for (int f = 0; f < 6; f++) if (nbr[f] != MPI_PROC_NULL) unpack[grid - 1][f]<<<num[f], dim[f]>>>(arr, rcv[f]);
  CK(cudaDeviceSynchronize());
  CK(cudaGetLastError());
  CK(cudaMemcpy(arr_host, arr, (size_t)g->s3b * sizeof(real), cudaMemcpyDeviceToHost));
  for (int f = 0; f < 6; f++) { cudaFree(snd[f]); cudaFree(rcv[f]); }
  cudaFree(arr);
  return 0;
}

/* cuda_solvability (cuda_bluebottle.cu:2313-2492) for one rank, with the reference's surf_int_* / plane_eps_* kernels and
 * thrust::reduce; the host sequence is restated (its TU is not linked).  One rank: it owns all six global faces and the
 * MPI_Allreduce is the identity.  u*, v*, w* travel host -> device -> host; eps_out[3] as after :2416. */
static int shim_face_sum(void (*kern)(real *, real *), dim3 num, dim3 dim, real *arr, int n, real *out)
{
  real *tmp;
  CK(cudaMalloc((void **)&tmp, (size_t)n * sizeof(real)));
  kern<<<num, dim>>>(arr, tmp);
  thrust::device_ptr<real> t(tmp);
  *out = thrust::reduce(t, t + n, 0., thrust::plus<real>());
  CK(cudaFree(tmp));
  return 0;
}

int bbref_solvability(real *u_h, real *v_h, real *w_h, int out_plane_, real *eps_out)
{
  const dom_struct *d = &dom[rank];
  CK(cudaMemcpy(_u_star, u_h, (size_t)d->Gfx.s3b * sizeof(real), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(_v_star, v_h, (size_t)d->Gfy.s3b * sizeof(real), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(_w_star, w_h, (size_t)d->Gfz.s3b * sizeof(real), cudaMemcpyHostToDevice));
  real exs, exe, eys, eye, ezs, eze, eps[3];
  if (shim_face_sum(surf_int_xs, blocks.Gfx.num_in, blocks.Gfx.dim_in, _u_star, d->Gfx.s2_i, &exs)) return -1;
  exs *= d->dy * d->dz;
  if (shim_face_sum(surf_int_xe, blocks.Gfx.num_in, blocks.Gfx.dim_in, _u_star, d->Gfx.s2_i, &exe)) return -1;
  exe *= d->dy * d->dz;
  if (shim_face_sum(surf_int_ys, blocks.Gfy.num_jn, blocks.Gfy.dim_jn, _v_star, d->Gfy.s2_j, &eys)) return -1;
  eys *= d->dz * d->dx;
  if (shim_face_sum(surf_int_ye, blocks.Gfy.num_jn, blocks.Gfy.dim_jn, _v_star, d->Gfy.s2_j, &eye)) return -1;
  eye *= d->dz * d->dx;
  if (shim_face_sum(surf_int_zs, blocks.Gfz.num_kn, blocks.Gfz.dim_kn, _w_star, d->Gfz.s2_k, &ezs)) return -1;
  ezs *= d->dx * d->dy;
  if (shim_face_sum(surf_int_ze, blocks.Gfz.num_kn, blocks.Gfz.dim_kn, _w_star, d->Gfz.s2_k, &eze)) return -1;
  eze *= d->dx * d->dy;
  eps[0] = exe - exs; eps[1] = eye - eys; eps[2] = eze - ezs;
  real sum;
  switch (out_plane_) {
    case WEST:   sum = (eps[0] + eps[1] + eps[2]) / (DOM.yl * DOM.zl); plane_eps_x_W<<<blocks.Gfx.num_in, blocks.Gfx.dim_in>>>(_u_star, sum); break;
    case EAST:   sum = (eps[0] + eps[1] + eps[2]) / (DOM.yl * DOM.zl); plane_eps_x_E<<<blocks.Gfx.num_in, blocks.Gfx.dim_in>>>(_u_star, sum); break;
    case SOUTH:  sum = (eps[0] + eps[1] + eps[2]) / (DOM.zl * DOM.xl); plane_eps_y_S<<<blocks.Gfy.num_jn, blocks.Gfy.dim_jn>>>(_v_star, sum); break;
    case NORTH:  sum = (eps[0] + eps[1] + eps[2]) / (DOM.zl * DOM.xl); plane_eps_y_N<<<blocks.Gfy.num_jn, blocks.Gfy.dim_jn>>>(_v_star, sum); break;
    case BOTTOM: sum = (eps[0] + eps[1] + eps[2]) / (DOM.xl * DOM.yl); plane_eps_z_B<<<blocks.Gfz.num_kn, blocks.Gfz.dim_kn>>>(_w_star, sum); break;
    case TOP:    sum = (eps[0] + eps[1] + eps[2]) / (DOM.xl * DOM.yl); plane_eps_z_T<<<blocks.Gfz.num_kn, blocks.Gfz.dim_kn>>>(_w_star, sum); break;
    case HOMOGENEOUS: {
      real sum_x = 0.5 * eps[0] / (DOM.yl * DOM.zl), sum_y = 0.5 * eps[1] / (DOM.zl * DOM.xl), sum_z = 0.5 * eps[2] / (DOM.xl * DOM.yl);
      plane_eps_x_W<<<blocks.Gfx.num_in, blocks.Gfx.dim_in>>>(_u_star, sum_x);
      plane_eps_x_E<<<blocks.Gfx.num_in, blocks.Gfx.dim_in>>>(_u_star, sum_x);
      plane_eps_y_S<<<blocks.Gfy.num_jn, blocks.Gfy.dim_jn>>>(_v_star, sum_y);
      plane_eps_y_N<<<blocks.Gfy.num_jn, blocks.Gfy.dim_jn>>>(_v_star, sum_y);
      plane_eps_z_B<<<blocks.Gfz.num_kn, blocks.Gfz.dim_kn>>>(_w_star, sum_z);
      plane_eps_z_T<<<blocks.Gfz.num_kn, blocks.Gfz.dim_kn>>>(_w_star, sum_z);
      break; }
    default: return -1;
  }
  CK(cudaDeviceSynchronize());
  CK(cudaGetLastError());
  CK(cudaMemcpy(u_h, _u_star, (size_t)d->Gfx.s3b * sizeof(real), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(v_h, _v_star, (size_t)d->Gfy.s3b * sizeof(real), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(w_h, _w_star, (size_t)d->Gfz.s3b * sizeof(real), cudaMemcpyDeviceToHost));
  if (eps_out) { eps_out[0] = eps[0]; eps_out[1] = eps[1]; eps_out[2] = eps[2]; }
  return 0;
}

/* cuda_dom_BC_star (cuda_bluebottle.cu:2111-2311; its translation unit is not linked -- it drags in the whole flow
 * solver): the host switch table restated, every KERNEL the reference's own BC_{u,v,w}_{W,E,S,N,B,T}_{D,N}
 * (bluebottle_kernel.cu:104-598) with the reference's launch shapes (blocks.Gf?.num_in / _jn / _kn). */
#define SHIM_BC_COMP(C, G, ARR, F, SHAPE)                                                                             \
  switch (bc.C##F) {                                                                                                  \
    case DIRICHLET: BC_##C##_##F##_D<<<blocks.G.num_##SHAPE, blocks.G.dim_##SHAPE>>>(ARR, bc.C##F##D); break;         \
    case NEUMANN:   BC_##C##_##F##_N<<<blocks.G.num_##SHAPE, blocks.G.dim_##SHAPE>>>(ARR); break;                     \
  }
#define SHIM_BC_FACE(NBR, F, SHAPE)                                                                                   \
  if (dom[rank].NBR == MPI_PROC_NULL) {                                                                               \
    SHIM_BC_COMP(u, Gfx, _u_star, F, SHAPE) SHIM_BC_COMP(v, Gfy, _v_star, F, SHAPE) SHIM_BC_COMP(w, Gfz, _w_star, F, SHAPE) \
  }
static void shim_dom_BC_star(void)
{
  SHIM_BC_FACE(w, W, in) SHIM_BC_FACE(e, E, in) SHIM_BC_FACE(s, S, jn) SHIM_BC_FACE(n, N, jn) SHIM_BC_FACE(b, B, kn) SHIM_BC_FACE(t, T, kn)
}

/* type / val: 18 entries, component-major (u on W,E,S,N,B,T, then v, then w); arrays are host, in place */
int bbref_dom_BC_star(real *u_h, real *v_h, real *w_h, const int *type, const real *val)
{
  const dom_struct *d = &dom[rank];
  int *ty[18] = { &bc.uW, &bc.uE, &bc.uS, &bc.uN, &bc.uB, &bc.uT, &bc.vW, &bc.vE, &bc.vS, &bc.vN, &bc.vB, &bc.vT,
                  &bc.wW, &bc.wE, &bc.wS, &bc.wN, &bc.wB, &bc.wT };
  real *vl[18] = { &bc.uWD, &bc.uED, &bc.uSD, &bc.uND, &bc.uBD, &bc.uTD, &bc.vWD, &bc.vED, &bc.vSD, &bc.vND, &bc.vBD, &bc.vTD,
                   &bc.wWD, &bc.wED, &bc.wSD, &bc.wND, &bc.wBD, &bc.wTD };
  for (int e = 0; e < 18; e++) { *ty[e] = type[e]; *vl[e] = val[e]; }
  CK(cudaMemcpy(_u_star, u_h, (size_t)d->Gfx.s3b * sizeof(real), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(_v_star, v_h, (size_t)d->Gfy.s3b * sizeof(real), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(_w_star, w_h, (size_t)d->Gfz.s3b * sizeof(real), cudaMemcpyHostToDevice));
  shim_dom_BC_star();
  CK(cudaDeviceSynchronize());
  CK(cudaGetLastError());
  CK(cudaMemcpy(u_h, _u_star, (size_t)d->Gfx.s3b * sizeof(real), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(v_h, _v_star, (size_t)d->Gfy.s3b * sizeof(real), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(w_h, _w_star, (size_t)d->Gfz.s3b * sizeof(real), cudaMemcpyDeviceToHost));
  return 0;
}

/* cuda_build_cages (cuda_particle.cu:1516-1646; its translation unit is not linked): the host sequence restated, every
 * KERNEL the reference's own (particle_kernel.cu:79-576): reset_flag_*, reset_phases, per particle cage_setup + build_phase,
 * then cage_setup + build_phase_shell, cage_flag_*, flag_external_*.  nparts_ = this rank's particle count (the reference's
 * `nparts`, ghost particles included); NPARTS > 0 selects the particle branch (:1524).  Outputs are host arrays (s3b). */
int bbref_build_cages(int NPARTS_, int nparts_, const real *px, const real *py, const real *pz, const real *pr, const int *pbc,
                      int *flag_u_h, int *flag_v_h, int *flag_w_h, int *phase_h, int *phase_shell_h)
{
  const dom_struct *d = &dom[rank];
  NPARTS = NPARTS_; nparts = nparts_;
  bc.pW = pbc[0]; bc.pE = pbc[1]; bc.pS = pbc[2]; bc.pN = pbc[3]; bc.pB = pbc[4]; bc.pT = pbc[5];
  part_struct *parts_h = (part_struct *)calloc(nparts_ > 0 ? nparts_ : 1, sizeof(part_struct));
  for (int n = 0; n < nparts_; n++) { parts_h[n].x = px[n]; parts_h[n].y = py[n]; parts_h[n].z = pz[n]; parts_h[n].r = pr[n]; }
  part_struct *_parts_d = NULL; dom_struct *_DOM_d = NULL; BC *_bc_d = NULL;
  CK(cudaMalloc(&_parts_d, (nparts_ > 0 ? nparts_ : 1) * sizeof(part_struct)));
  CK(cudaMemcpy(_parts_d, parts_h, (nparts_ > 0 ? nparts_ : 1) * sizeof(part_struct), cudaMemcpyHostToDevice));
  CK(cudaMalloc(&_DOM_d, sizeof(dom_struct))); CK(cudaMemcpy(_DOM_d, &DOM, sizeof(dom_struct), cudaMemcpyHostToDevice));
  CK(cudaMalloc(&_bc_d, sizeof(BC))); CK(cudaMemcpy(_bc_d, &bc, sizeof(BC), cudaMemcpyHostToDevice));

  reset_flag_u<<<blocks.Gfx.num_inb, blocks.Gfx.dim_inb>>>(_flag_u);                    /* :1520-1522 */
  reset_flag_v<<<blocks.Gfy.num_jnb, blocks.Gfy.dim_jnb>>>(_flag_v);
  reset_flag_w<<<blocks.Gfz.num_knb, blocks.Gfz.dim_knb>>>(_flag_w);
  if (NPARTS > 0) {
    reset_phases<<<blocks.Gcc.num_knb, blocks.Gcc.dim_knb>>>(_phase, _phase_shell);     /* :1526 */
    int tx = 0.5 * MAX_THREADS_DIM, ty = 0.5 * MAX_THREADS_DIM, tz = 0.5 * MAX_THREADS_DIM;
    real itx = 1. / tx, ity = 1. / ty, itz = 1. / tz;
    int cage_dim[3], *_cage_dim;
    CK(cudaMalloc(&_cage_dim, 3 * sizeof(int)));
    for (int pass = 0; pass < 2; pass++) {                                              /* :1541-1584: phase, then phase_shell */
      for (int n = 0; n < nparts; n++) {
        cage_setup<<<1, 1>>>(_parts_d, n, _cage_dim);
        CK(cudaMemcpy(cage_dim, _cage_dim, 3 * sizeof(int), cudaMemcpyDeviceToHost));
        int bx = (int)ceil((real)cage_dim[0] * itx), by = (int)ceil((real)cage_dim[1] * ity), bz = (int)ceil((real)cage_dim[2] * itz);
        dim3 dimb_3(tx, ty, tz), numb_3(bx, by, bz);
        if (bx > 0 && by > 0 && bz > 0) {
          if (pass == 0) build_phase<<<numb_3, dimb_3>>>(_parts_d, n, _cage_dim, _phase, _phase_shell, _DOM_d, _bc_d);
          else build_phase_shell<<<numb_3, dimb_3>>>(_parts_d, n, _cage_dim, _phase, _phase_shell, _DOM_d, _bc_d);
        }
      }
    }
    CK(cudaFree(_cage_dim));
    cage_flag_u<<<blocks.Gfx.num_inb, blocks.Gfx.dim_inb>>>(_flag_u, _phase, _phase_shell);   /* :1595-1597 */
    cage_flag_v<<<blocks.Gfy.num_jnb, blocks.Gfy.dim_jnb>>>(_flag_v, _phase, _phase_shell);
    cage_flag_w<<<blocks.Gfz.num_knb, blocks.Gfz.dim_knb>>>(_flag_w, _phase, _phase_shell);
  }
  if (bc.pW != PERIODIC && bc.pE != PERIODIC) {                                         /* :1605-1639 */
    if (d->I == DOM.Is) flag_external_u<<<blocks.Gfx.num_inb, blocks.Gfx.dim_inb>>>(_flag_u, d->Gfx._is);
    if (d->I == DOM.Ie) flag_external_u<<<blocks.Gfx.num_inb, blocks.Gfx.dim_inb>>>(_flag_u, d->Gfx._ie);
  }
  if (bc.pS != PERIODIC && bc.pN != PERIODIC) {
    if (d->J == DOM.Js) flag_external_v<<<blocks.Gfy.num_jnb, blocks.Gfy.dim_jnb>>>(_flag_v, d->Gfy._js);
    if (d->J == DOM.Je) flag_external_v<<<blocks.Gfy.num_jnb, blocks.Gfy.dim_jnb>>>(_flag_v, d->Gfy._je);
  }
  if (bc.pB != PERIODIC && bc.pT != PERIODIC) {
    if (d->K == DOM.Ks) flag_external_w<<<blocks.Gfz.num_knb, blocks.Gfz.dim_knb>>>(_flag_w, d->Gfz._ks);
    if (d->K == DOM.Ke) flag_external_w<<<blocks.Gfz.num_knb, blocks.Gfz.dim_knb>>>(_flag_w, d->Gfz._ke);
  }
  CK(cudaDeviceSynchronize());
  CK(cudaGetLastError());
  CK(cudaMemcpy(flag_u_h, _flag_u, (size_t)d->Gfx.s3b * sizeof(int), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(flag_v_h, _flag_v, (size_t)d->Gfy.s3b * sizeof(int), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(flag_w_h, _flag_w, (size_t)d->Gfz.s3b * sizeof(int), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(phase_h, _phase, (size_t)d->Gcc.s3b * sizeof(int), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(phase_shell_h, _phase_shell, (size_t)d->Gcc.s3b * sizeof(int), cudaMemcpyDeviceToHost));
  CK(cudaFree(_parts_d)); CK(cudaFree(_DOM_d)); CK(cudaFree(_bc_d));
  free(parts_h);
  return 0;
}

/* device pointers of the reference's arrays, for the benchmark's device-resident leg: 0 phi, 1 p0, 2 p */
void *bbref_dev_ptr(int which) { return which == 0 ? (void *)_phi : which == 1 ? (void *)_p0 : (void *)_p; }

} /* extern "C" */