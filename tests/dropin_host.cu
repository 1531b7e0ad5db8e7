/* dropin_host.cu -- a miniature stand-in for Bluebottle's host program (src/bluebottle.c) that exercises
 * libbbpcg_dropin.so exactly the way the reference does: it DEFINES the globals the reference defines
 * (src/bluebottle.c:438-576, src/mpi_comm.c:26-27, src/particle.c:27-28), allocates the device arrays in the
 * reference's layouts (src/cuda_bluebottle.cu:144-432), and calls the reference's entry-point names
 *
 *     cuda_PP_init_jacobi_preconditioner();   // src/bluebottle.c:139
 *     cuda_PP_cg_noparts();  or  cuda_PP_cg(); // src/bluebottle.c:228-232
 *     mpi_cuda_exchange_Gcc(_phi);             // src/bluebottle.c:233
 *     cuda_dom_BC_p(_phi); cuda_project(); cuda_update_p();   // src/bluebottle.c:234,237,250 (when an epilogue file is asked for)
 *
 * with `void f(void)` signatures.  recorder_PP() and cuda_part_BC_p() -- callees the library calls back -- are
 * this file's own small stand-ins (same signature; recorder_PP / recorder_PP_timed write the record lines through the
 * product's bb_recorder_PP* writers).  TEST CODE: built and run by tests/test_dropin.py.
 *
 * Multi-rank: one process per rank, as Bluebottle runs under mpirun.  The environment stands in for the launcher:
 * BB_RANK / BB_NPROCS (OMPI_COMM_WORLD_RANK / _SIZE) and BB_RDV, a directory through which bb_dropin_allgather() -- the
 * three-line MPI_Allgather hook of INTEGRATION.md -- exchanges the ranks' attach records as files.  Rank r uses GPU
 * r mod (device count), reads <inputs.bin>.<r> and writes <phi_out.bin>.<r>.
 *
 *   dropin_host <flow.config> <decomp.config> <inputs.bin> <phi_out.bin> <root>/record <noparts|parts|timed> [pp_max_iter [epilogue_out.bin]]
 *
 * epilogue_out.bin: u, v, w (float64, Gfx/Gfy/Gfz s3b) and p (Gcc s3b) after the epilogue with the previous pressure
 *             p0[C] = ((C * 2654435761 mod 2^32) >> 8) / 2^24 - 0.5 (exactly reproducible in numpy).
 *
 * inputs.bin: flag_u, flag_v, flag_w (int32, Gfx/Gfy/Gfz s3b), phase, phase_shell (int32, Gcc s3b),
 *             u_star, v_star, w_star (float64, Gfx/Gfy/Gfz s3b), in that order, raw.
 */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <unistd.h>
#include <cuda_runtime.h>

#include "../include/bbpcg.h"
#include "../include/bb_dropin.h"

extern "C" {
/* ---- the globals the drop-in layer imports ---- */
dom_struct *dom = NULL;
dom_struct DOM;
int rank = 0, nprocs = 1;
bb_BC bc;                               /* the reference's BC struct (include/bb_grid.h mirrors it field for field) */
real rho_f = 1., dt = 1e-3, pp_residual = 1e-6, ttime = 0.;
int pp_max_iter = 2000, stepnum = 0;
int NPARTS = 0, nparts = 0;
int out_plane = 10;                     /* HOMOGENEOUS, src/bluebottle.h:353 */
void *_parts = NULL;                    /* part_struct *_parts (src/particle.h:692); only cuda_build_cages reads it */
real *_u_star = NULL, *_v_star = NULL, *_w_star = NULL, *_rhs_p = NULL, *_phi = NULL;
real *_u = NULL, *_v = NULL, *_w = NULL, *_p = NULL, *_p0 = NULL;
int *_flag_u = NULL, *_flag_v = NULL, *_flag_w = NULL, *_phase = NULL, *_phase_shell = NULL;
static char g_record_dir[1024] = ".";      /* the reference's ROOT_DIR: records go to <ROOT_DIR>/record/ */

/* src/recorder.c:190-221 and :259-336 through the PRODUCT's record writers (bb_recorder_PP*, include/bbpcg.h): the header
 * is created on the first line, the columns are the reference's.  Multi-rank: the reference's rank 0 writes one line after
 * an MPI_Allreduce of the elapsed time; here every rank writes its own <name>.<rank> so that the test sees all of them. */
static const char *rec_name(char *buf, size_t cap, const char *name)
{
  if (nprocs > 1) { snprintf(buf, cap, "%s.%d", name, rank); return buf; }
  return name;
}
void recorder_PP(char *name, int niter, real resid, real etime)
{
  char nm[256];
  if (bb_recorder_PP(g_record_dir, rec_name(nm, sizeof(nm), name), stepnum, ttime, dt, niter, resid, etime)) { fprintf(stderr, "%s\n", bbpcg_last_error()); exit(EXIT_FAILURE); }
}
void recorder_PP_init_timed(char *name)
{
  char nm[256];
  if (bb_recorder_PP_init_timed(g_record_dir, rec_name(nm, sizeof(nm), name))) { fprintf(stderr, "%s\n", bbpcg_last_error()); exit(EXIT_FAILURE); }
}
void recorder_PP_timed(char *name, int niter, real resid, real etime, real etime_spmv, real etime_ip1, real etime_AR1,
                       real etime_up1, real etime_ip2, real etime_AR2, real etime_up2, real etime_mpi)
{
  char nm[256];
  const real seg[8] = { etime_spmv, etime_ip1, etime_AR1, etime_up1, etime_ip2, etime_AR2, etime_up2, etime_mpi };
  if (bb_recorder_PP_timed(g_record_dir, rec_name(nm, sizeof(nm), name), stepnum, ttime, dt, niter, resid, etime, seg)) { fprintf(stderr, "%s\n", bbpcg_last_error()); exit(EXIT_FAILURE); }
}
}

/* the multi-rank bootstrap hook (include/bb_dropin.h): MPI_Allgather in Bluebottle, files in a rendezvous directory here */
extern "C" int bb_dropin_allgather(const void *send, void *recv, int bytes_per_rank)
{
  static int seq = 0;
  const char *rdv = getenv("BB_RDV");
  if (!rdv) return -1;
  char path[2048], tmp[2048];
  snprintf(path, sizeof(path), "%s/gather.%d.%d", rdv, seq, rank);
  snprintf(tmp, sizeof(tmp), "%s.tmp", path);
  FILE *f = fopen(tmp, "wb");
  if (!f || fwrite(send, 1, bytes_per_rank, f) != (size_t)bytes_per_rank) return -1;
  fclose(f);
  if (rename(tmp, path)) return -1;
  for (int r = 0; r < nprocs; r++) {
    snprintf(path, sizeof(path), "%s/gather.%d.%d", rdv, seq, r);
    FILE *g = NULL;
    for (int tries = 0; tries < 60000 && !(g = fopen(path, "rb")); tries++) usleep(1000);      /* <= 60 s */
    if (!g || fread((char *)recv + (size_t)r * bytes_per_rank, 1, bytes_per_rank, g) != (size_t)bytes_per_rank) return -1;
    fclose(g);
  }
  seq++;
  return 0;
}

/* stand-in for cuda_part_BC_p (src/cuda_particle.cu:1680): the net effect of part_BC_p on rhs
 * (src/particle_kernel.cu:1753): zero in every cell that is solid or outside the particle shell mask */
__global__ void k_host_part_bc(real *rhs, const int *phase, const int *phase_shell, long long n)
{
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < n) rhs[i] = (real)(phase[i] < 0 && phase_shell[i]) * rhs[i];
}
extern "C" void cuda_part_BC_p(void)
{
  const long long n = dom[rank].Gcc.s3b;
  k_host_part_bc<<<(unsigned)((n + 255) / 256), 256>>>(_rhs_p, _phase, _phase_shell, n);   /* default stream, like the reference */
}

template <typename T> static T *upload(FILE *f, size_t n, bool have_gpu)
{
  std::vector<T> h(n);
  if (fread(h.data(), sizeof(T), n, f) != n) { fprintf(stderr, "inputs.bin too short\n"); exit(2); }
  if (!have_gpu) return NULL;
  T *d = NULL;
  if (cudaMalloc(&d, n * sizeof(T)) != cudaSuccess || cudaMemcpy(d, h.data(), n * sizeof(T), cudaMemcpyHostToDevice) != cudaSuccess) {
    fprintf(stderr, "device allocation failed\n"); exit(2);
  }
  return d;
}

int main(int argc, char **argv)
{
  if (argc < 7) { fprintf(stderr, "usage: %s flow.config decomp.config inputs.bin phi_out.bin record_dir noparts|parts [pp_max_iter]\n", argv[0]); return 2; }
  bb_flow_params fp;
  if (bb_domain_read(argv[1], argv[2], &DOM, &dom, (bb_pressure_bc *)&bc, &fp)) { fprintf(stderr, "%s\n", bbpcg_last_error()); return 2; }
  if (getenv("BB_RANK")) { rank = atoi(getenv("BB_RANK")); nprocs = atoi(getenv("BB_NPROCS")); }
  if (DOM.In * DOM.Jn * DOM.Kn != nprocs) { fprintf(stderr, "decomposition has %d blocks but BB_NPROCS = %d\n", DOM.In * DOM.Jn * DOM.Kn, nprocs); return 2; }
  rho_f = fp.rho_f; pp_residual = fp.pp_residual; pp_max_iter = argc > 7 ? atoi(argv[7]) : fp.pp_max_iter;
  snprintf(g_record_dir, sizeof(g_record_dir), "%s", argv[5]);
  {                                                     /* <root>/record is what the command line names; keep <root> */
    size_t n = strlen(g_record_dir);
    while (n > 1 && g_record_dir[n - 1] == '/') g_record_dir[--n] = 0;
    if (n >= 7 && !strcmp(g_record_dir + n - 7, "/record")) g_record_dir[n - 7] = 0;
  }
  const bool parts = strcmp(argv[6], "parts") == 0, timed = strcmp(argv[6], "timed") == 0;
  NPARTS = nparts = parts ? 1 : 0;
  int ndev = 0;
  const bool have_gpu = cudaGetDeviceCount(&ndev) == cudaSuccess && ndev > 0;
  if (have_gpu) cudaSetDevice(rank % ndev);             /* the reference selects the device before MPI_Init (src/mpi_comm.c:42-63) */
  const dom_struct *d = &dom[rank];
  const std::string sfx = nprocs > 1 ? "." + std::to_string(rank) : "";
  const std::string in_path = argv[3] + sfx, out_path = argv[4] + sfx;
  FILE *f = fopen(in_path.c_str(), "rb");
  if (!f) { fprintf(stderr, "cannot open %s\n", in_path.c_str()); return 2; }
  _flag_u = upload<int>(f, d->Gfx.s3b, have_gpu); _flag_v = upload<int>(f, d->Gfy.s3b, have_gpu); _flag_w = upload<int>(f, d->Gfz.s3b, have_gpu);
  _phase = upload<int>(f, d->Gcc.s3b, have_gpu); _phase_shell = upload<int>(f, d->Gcc.s3b, have_gpu);
  _u_star = upload<real>(f, d->Gfx.s3b, have_gpu); _v_star = upload<real>(f, d->Gfy.s3b, have_gpu); _w_star = upload<real>(f, d->Gfz.s3b, have_gpu);
  fclose(f);
  if (have_gpu) {
    cudaMalloc(&_rhs_p, sizeof(real) * d->Gcc.s3b); cudaMalloc(&_phi, sizeof(real) * d->Gcc.s3b);
    cudaMemset(_phi, 0, sizeof(real) * d->Gcc.s3b);
  }
  /* ---- what src/bluebottle.c does ---- */
  cuda_PP_init_jacobi_preconditioner();                 /* :139 -- without a GPU this prints and exits(EXIT_FAILURE) */
  stepnum = 1; ttime = dt;
  if (parts) cuda_PP_cg(); else if (timed) cuda_PP_cg_timed(); else cuda_PP_cg_noparts();   /* :228-232 */
  mpi_cuda_exchange_Gcc(_phi);                          /* :233 */
  if (argc > 8) {                                       /* the solve epilogue, src/bluebottle.c:234-250 */
    std::vector<real> p0(d->Gcc.s3b);
    for (size_t C = 0; C < p0.size(); C++) p0[C] = (real)(((unsigned)(C * 2654435761ull)) >> 8) / 16777216. - 0.5;
    cudaMalloc(&_u, sizeof(real) * d->Gfx.s3b); cudaMalloc(&_v, sizeof(real) * d->Gfy.s3b); cudaMalloc(&_w, sizeof(real) * d->Gfz.s3b);
    cudaMalloc(&_p, sizeof(real) * d->Gcc.s3b); cudaMalloc(&_p0, sizeof(real) * d->Gcc.s3b);
    cudaMemset(_u, 0, sizeof(real) * d->Gfx.s3b); cudaMemset(_v, 0, sizeof(real) * d->Gfy.s3b); cudaMemset(_w, 0, sizeof(real) * d->Gfz.s3b);
    cudaMemset(_p, 0, sizeof(real) * d->Gcc.s3b);
    cudaMemcpy(_p0, p0.data(), sizeof(real) * p0.size(), cudaMemcpyHostToDevice);
    cuda_dom_BC_p(_phi);                                /* :234 */
    cuda_project();                                     /* :237 */
    cuda_update_p();                                    /* :250 */
    FILE *e = fopen(argv[8], "wb");
    const size_t ns[4] = { (size_t)d->Gfx.s3b, (size_t)d->Gfy.s3b, (size_t)d->Gfz.s3b, (size_t)d->Gcc.s3b };
    real *src[4] = { _u, _v, _w, _p };
    for (int a = 0; a < 4; a++) {
      std::vector<real> h(ns[a]);
      cudaMemcpy(h.data(), src[a], sizeof(real) * ns[a], cudaMemcpyDeviceToHost);
      fwrite(h.data(), sizeof(real), ns[a], e);
    }
    fclose(e);
  }
  std::vector<real> phi(d->Gcc.s3b);
  cudaMemcpy(phi.data(), _phi, sizeof(real) * phi.size(), cudaMemcpyDeviceToHost);
  FILE *o = fopen(out_path.c_str(), "wb");
  fwrite(phi.data(), sizeof(real), phi.size(), o);
  fclose(o);
  bbpcg_dropin_finalize();
  bb_domain_free(dom);
  printf("dropin_host: done\n");
  return 0;
}
