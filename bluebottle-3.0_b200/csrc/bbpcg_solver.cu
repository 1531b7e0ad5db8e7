/* bbpcg_solver.cu -- host driver and C ABI of libbbpcg.so (include/bbpcg.h).
 *
 * Host control flow of cuda_PP_cg / cuda_PP_cg_noparts (src/cuda_solver.cu:38-300, 573-761)
 * re-expressed for a device-resident recurrence: the host only enqueues kernels in batches and
 * polls one `done` word; alpha, beta, the stop test and the iteration count live in device
 * memory (struct Scal).  There is NO CPU fallback: every entry point needs a CUDA device.
 */
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <unistd.h>

#include "bbpcg_kernels.cuh"
#include "bbpcg_search_tma.cuh"
#include "bbpcg_resid_tma.cuh"
#include "bbpcg_epilogue.cuh"
#include "bbpcg_cages.cuh"

/* ---- error plumbing ---------------------------------------------------------------------- */
static thread_local char g_err[512] = "";
extern "C" void bbpcg_set_error(const char *fmt, ...)
{
  va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap);
}
extern "C" const char *bbpcg_last_error(void) { return g_err; }
extern "C" const char *bbpcg_version(void) { return "bbpcg 0.1 (sm_100a)"; }

#define CU(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  bbpcg_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); return BBPCG_ECUDA; } } while (0)

/* ---- the solver object ------------------------------------------------------------------- */
struct ExportBlob {                 /* one rank's record for bbpcg_comm_import (<= BBPCG_BLOB_BYTES) */
  unsigned magic;
  int rank, in, jn, kn, device;
  long long pid;
  unsigned long long arena_ptr;     /* valid inside the exporting process */
  unsigned long long arena_bytes;
  cudaIpcMemHandle_t handle;
  unsigned char uuid[16];           /* of the exporting GPU: two PROCESSES on one GPU must be recognised as sharing it */
};
static_assert(sizeof(ExportBlob) <= BBPCG_BLOB_BYTES, "blob too large");
#define BB_MAGIC 0xbb9c6001u

struct bbpcg_solver {
  dom_struct dom, DOM;
  bb_pressure_bc bc;
  int device;
  cudaStream_t stream;
  cudaEvent_t ev[4];
  cudaEvent_t ev_poll[2];
  char *arena;
  ArenaMap amap;
  Dev dev;
  FaceStrides fst;
  int nranks;
  char *peer_arena[BB_MAXR];        /* mapped bases (own arena for self) */
  bool peer_opened[BB_MAXR];
  int has_phase;                    /* set_coefficients was given a phase array */
  int coeffs_set;
  /* launch configuration */
  int opt_ty, opt_kc;               /* user overrides of the plan (0 = automatic) */
  int stream_blocks;
  int check_every;                  /* iterations per polling batch */
  int sm_count;
  /* pinned poll words + host-mode buffers */
  int *h_poll;
  Scal *h_scal;
  double *hb_u, *hb_v, *hb_w, *hb_rhs, *hb_phi;
  long long launches;
  unsigned exchange_count;
  /* tile / z-chunk plan of the two iteration kernels (device table Dev::ztab) */
  int *h_ztab;                      /* pinned [2 * BB_MAXZ] */
  int plan_ok;                      /* the plan below, the uploaded table and the tensor maps are current */
  int plan_ty, plan_nbx, plan_nby, plan_nbz, plan_kc;
  int pdl;                          /* programmatic dependent launch of the two iteration kernels: 0 off, 1 on, 2 auto */
  int guided, guided_pct, chunk_min; /* z-chunk plan of small blocks: decreasing chunk lengths (make_plan) */
  int epilogue_tiled, epi_chunk;    /* 1: the 16^3-brick epilogue kernel instead of the streaming pair; planes per chunk of the pair */
  int tma_warp;                     /* 1 (default): a dedicated producer warp issues the iteration kernels' TMA loads; 0: thread 0 does */
  int rhs_tiled;                    /* PP_rhs through shared-memory transposes (default) or the row-walking kernel */
  int shared_device;                /* some peer rank lives on this same GPU (single-process harness, or two processes on one GPU) */
  unsigned char uuid[16];
  SearchMaps maps;                  /* tensor maps of the iteration kernels for the planned tile height */
  int kernel_timing;                /* 0 off, 1 every iteration, n > 1: the first n iterations of a solve (the others keep PDL) */
  int kt_now;                       /* the iteration being enqueued is instrumented: no PDL attribute on its launches */
  cudaEvent_t *kev;                 /* [2*BB_KT_CAP+1] */
  double kt_search_ms, kt_resid_ms, kt_refresh_ms;
  int kt_search_n, kt_resid_n, kt_refresh_n;
};
#define BB_KT_CAP 4096
#define BB_POLL_COMM 8                /* int index inside h_poll of the comm-timeout word the device sets */

static int preload_kernels();

static int nbr_rank(const dom_struct &d, int f)
{
  switch (f) { case 0: return d.e; case 1: return d.w; case 2: return d.n; case 3: return d.s; case 4: return d.t; default: return d.b; }
}

static void point_dev_at_arena(bbpcg_solver *s)
{
  Dev &d = s->dev;
  char *a = s->arena;
  const ArenaMap &m = s->amap;
  d.r = (double *)(a + m.r); d.P[0] = (double *)(a + m.p0); d.P[1] = (double *)(a + m.p1);
  d.x = (double *)(a + m.x);
  d.fmask = (u8 *)(a + m.fmask);
  for (int b = 0; b < 2; b++) for (int f = 0; f < 6; f++) d.recv[b][f] = (double *)(a + m.recv[b][f]);
  d.partials = (double *)(a + m.partials); d.gpartials = (double *)(a + m.gpartials); d.counter = (unsigned *)(a + m.counter);
  d.sc = (Scal *)(a + m.scal); d.history = (double *)(a + m.history);
  d.invM_tab = (const double *)(a + m.invM_tab);
  d.ztab = (const int *)(a + m.ztab);
}

/* ---- TMA tensor maps (cuTensorMapEncodeTiled through the runtime's driver entry point, so the
 * library needs no link-time dependency on libcuda) ---- */
typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled encode_fn()
{
  static PFN_encodeTiled fn = NULL;
  if (!fn) {
    void *p = NULL;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = (PFN_encodeTiled)p;
  }
  return fn;
}

/* 3-D map over one P-layout array (px x (jn+2) x (kn+2), x fastest) with box bx x by x 1 */
static int make_map(CUtensorMap *m, void *base, const Layout &L, bool u8, int bx, int by)
{
  PFN_encodeTiled fn = encode_fn();
  if (!fn) { bbpcg_set_error("cuTensorMapEncodeTiled is not available from this driver"); return BBPCG_ECUDA; }
  const size_t es = u8 ? 1 : 8;
  cuuint64_t dims[3] = { (cuuint64_t)L.px, (cuuint64_t)(L.jn + 2), (cuuint64_t)(L.kn + 2) };
  cuuint64_t strides[2] = { (cuuint64_t)L.px * es, (cuuint64_t)L.ps * es };
  cuuint32_t box[3] = { (cuuint32_t)bx, (cuuint32_t)by, 1u }, estr[3] = { 1u, 1u, 1u };
  CUresult r = fn(m, u8 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, base, dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { bbpcg_set_error("cuTensorMapEncodeTiled failed (%d) for box %d x %d", (int)r, bx, by); return BBPCG_ECUDA; }
  return BBPCG_OK;
}

/* boxes follow the planned tile height ty: halo'd tiles (TX+4) x (ty+2), owned tiles TX x ty */
static int build_search_maps(bbpcg_solver *s, int ty)
{
  typedef SearchGeom<true> G;
  const Dev &d = s->dev;
  SearchMaps *M = &s->maps;
  const int hy = ty + 2;
  int rc = 0;
  memset(M, 0, sizeof(*M));
  if (!rc) rc = make_map(&M->r, d.r, d.L, false, G::HXP, hy);
  if (!rc) rc = make_map(&M->p[0], d.P[0], d.L, false, G::HXP, hy);
  if (!rc) rc = make_map(&M->p[1], d.P[1], d.L, false, G::HXP, hy);
  if (!rc) rc = make_map(&M->fm, d.fmask, d.L, true, G::MXP, hy);
  if (!rc) rc = make_map(&M->xo, d.x, d.L, false, G::TX, ty);
  if (!rc) rc = make_map(&M->ro, d.r, d.L, false, G::TX, ty);
  if (!rc) rc = make_map(&M->xh, d.x, d.L, false, G::HXP, hy);
  return rc;
}

/* neighbour tables for a set of ranks whose arenas are addressable at peer_arena[] with
 * interior sizes dims[][3] */
static void build_halo(bbpcg_solver *s, const int (*dims)[3])
{
  static const int opposite[6] = { 1, 0, 3, 2, 5, 4 };
  Dev &d = s->dev;
  for (int f = 0; f < 6; f++) {
    NbrFace &nf = d.halo.f[f];
    memset(&nf, 0, sizeof(nf));
    int nb = nbr_rank(s->dom, f);
    if (nb < 0) continue;
    char *base = s->peer_arena[nb];
    if (!base) continue;
    Layout L = make_layout(dims[nb][0], dims[nb][1], dims[nb][2]);
    ArenaMap m = make_arena_map(L);
    nf.L = L;
    nf.r = (double *)(base + m.r); nf.x = (double *)(base + m.x); nf.fmask = (u8 *)(base + m.fmask);
    for (int b = 0; b < 2; b++) nf.recv[b] = (double *)(base + m.recv[b][opposite[f]]);
  }
  d.any_nbr = 0;
  for (int f = 0; f < 6; f++) if (d.halo.f[f].r) d.any_nbr = 1;
  d.comm.rank = s->dom.rank; d.comm.nranks = s->nranks;
  if (d.comm.timeout_cycles == 0) d.comm.timeout_cycles = 1ll << 37;       /* ~70 s default: far above any start-up or I/O skew; option comm_timeout_ms (0 = wait for ever) */
  for (int p = 0; p < BB_MAXR; p++) d.comm.mbox[p] = NULL;
  for (int p = 0; p < s->nranks; p++) {
    Layout L = make_layout(dims[p][0], dims[p][1], dims[p][2]);
    ArenaMap m = make_arena_map(L);
    d.comm.mbox[p] = (unsigned long long *)(s->peer_arena[p] + m.mbox);
  }
}

/* everything of bbpcg_create that can fail after the object exists; on failure the caller destroys the half-built
 * object (bbpcg_destroy tolerates missing pieces), so no error path leaks the arena, events or pinned buffers */
static int create_impl(bbpcg_solver *s, const dom_struct *dom_rank, const dom_struct *DOM)
{
  const grid_info &g = dom_rank->Gcc;
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, s->device));
  s->sm_count = prop.multiProcessorCount;
  memcpy(s->uuid, &prop.uuid, 16);
  Dev &d = s->dev;
  d.L = make_layout(g.in, g.jn, g.kn);
  s->amap = make_arena_map(d.L);
  if (cudaMalloc(&s->arena, s->amap.total) != cudaSuccess) {
    bbpcg_set_error("bbpcg_create: cudaMalloc of %zu bytes failed", s->amap.total); cudaGetLastError(); s->arena = NULL; return BBPCG_ENOMEM;
  }
  CU(cudaMemset(s->arena, 0, s->amap.total));
  point_dev_at_arena(s);
  /* the zero-filled mask array marks every cell dead; ghosts behind walls are never written and stay so */
  d.idx2 = 1. / (dom_rank->dx * dom_rank->dx); d.idy2 = 1. / (dom_rank->dy * dom_rank->dy); d.idz2 = 1. / (dom_rank->dz * dom_rank->dz);
  d.dx2_6 = (dom_rank->dx * dom_rank->dx) / 6.; d.dy2_6 = (dom_rank->dy * dom_rank->dy) / 6.; d.dz2_6 = (dom_rank->dz * dom_rank->dz) / 6.;
  s->fst.us1b = dom_rank->Gfx.s1b; s->fst.us2b = dom_rank->Gfx.s2b;
  s->fst.vs1b = dom_rank->Gfy.s1b; s->fst.vs2b = dom_rank->Gfy.s2b;
  s->fst.ws1b = dom_rank->Gfz.s1b; s->fst.ws2b = dom_rank->Gfz.s2b;
  s->fst.cs1b = g.s1b; s->fst.cs2b = g.s2b;
  { int rc = preload_kernels(); if (rc) return rc; }
  k_build_tab<<<1, 128>>>((double *)(s->arena + s->amap.invM_tab), d.idx2, d.idy2, d.idz2);
  CU(cudaDeviceSynchronize());
  CU(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
  for (int i = 0; i < 4; i++) CU(cudaEventCreate(&s->ev[i]));
  for (int i = 0; i < 2; i++) CU(cudaEventCreateWithFlags(&s->ev_poll[i], cudaEventDisableTiming));
  CU(cudaHostAlloc(&s->h_poll, 64, cudaHostAllocMapped | cudaHostAllocPortable));
  memset(s->h_poll, 0, 64);
  CU(cudaHostGetDevicePointer((void **)&d.comm.host_flag, (void *)&s->h_poll[BB_POLL_COMM], 0));
  CU(cudaHostAlloc(&s->h_scal, sizeof(Scal), cudaHostAllocDefault));
  CU(cudaHostAlloc(&s->h_ztab, sizeof(int) * (2 * BB_MAXZ), cudaHostAllocDefault));
  s->pdl = 2; s->rhs_tiled = 1; s->tma_warp = 1; s->guided = 1; s->guided_pct = 60; s->chunk_min = 4;
  /* single rank: neighbours are this block itself (periodic wrap) or nothing */
  s->nranks = 1;
  for (int p = 0; p < BB_MAXR; p++) { s->peer_arena[p] = NULL; s->peer_opened[p] = false; }
  if (DOM->In * DOM->Jn * DOM->Kn == 1) {
    s->peer_arena[0] = s->arena;
    int dims[1][3] = { { g.in, g.jn, g.kn } };
    build_halo(s, dims);
  } else {
    memset(&d.halo, 0, sizeof(d.halo));     /* until bbpcg_comm_import */
    d.comm.rank = dom_rank->rank; d.comm.nranks = 1;
  }
  s->stream_blocks = s->sm_count * 8;
  s->check_every = 10;
#ifdef BB_TRACE
  CU(cudaMalloc(&d.trace, (size_t)BB_TRACE_LAUNCHES * BB_TRACE_CTAS * BB_TRACE_EV * 8));
  CU(cudaMemset(d.trace, 0, (size_t)BB_TRACE_LAUNCHES * BB_TRACE_CTAS * BB_TRACE_EV * 8));
#endif
  return BBPCG_OK;
}

extern "C" int bbpcg_create(bbpcg_solver **out, const dom_struct *dom_rank, const dom_struct *DOM,
                            const bb_pressure_bc *bc, int device)
{
  if (!out || !dom_rank || !DOM || !bc) { bbpcg_set_error("bbpcg_create: NULL argument"); return BBPCG_EINVAL; }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    bbpcg_set_error("bbpcg_create: no CUDA device (this library has no CPU path)"); return BBPCG_ECUDA;
  }
  if (device < 0) CU(cudaGetDevice(&device));
  CU(cudaSetDevice(device));
  const grid_info &g = dom_rank->Gcc;
  if (g.in < 1 || g.jn < 1 || g.kn < 1 || g.s1b != g.in + 2 || g.s2b != g.s1b * (g.jn + 2)) {
    bbpcg_set_error("bbpcg_create: dom_struct is not filled (run bb_domain_fill)"); return BBPCG_EINVAL;
  }
  bbpcg_solver *s = new bbpcg_solver();
  memset(s, 0, sizeof(*s));
  s->dom = *dom_rank; s->DOM = *DOM; s->bc = *bc; s->device = device;
  const int rc = create_impl(s, dom_rank, DOM);
  if (rc) { bbpcg_destroy(s); return rc; }
  *out = s;
  return BBPCG_OK;
}

extern "C" void bbpcg_destroy(bbpcg_solver *s)
{
  if (!s) return;
  cudaSetDevice(s->device);
  if (s->stream) cudaStreamSynchronize(s->stream);
  for (int p = 0; p < BB_MAXR; p++) if (s->peer_opened[p]) cudaIpcCloseMemHandle(s->peer_arena[p]);
  if (s->arena) cudaFree(s->arena);
  cudaFree(s->hb_u); cudaFree(s->hb_v); cudaFree(s->hb_w); cudaFree(s->hb_rhs); cudaFree(s->hb_phi);
  if (s->h_poll) cudaFreeHost(s->h_poll);
  if (s->h_scal) cudaFreeHost(s->h_scal);
  if (s->h_ztab) cudaFreeHost(s->h_ztab);
  if (s->kev) { for (int i = 0; i <= 2 * BB_KT_CAP; i++) if (s->kev[i]) cudaEventDestroy(s->kev[i]); free(s->kev); }
  for (int i = 0; i < 4; i++) if (s->ev[i]) cudaEventDestroy(s->ev[i]);
  for (int i = 0; i < 2; i++) if (s->ev_poll[i]) cudaEventDestroy(s->ev_poll[i]);
  if (s->stream) cudaStreamDestroy(s->stream);
  cudaGetLastError();
  delete s;
}

/* ---- multi-GPU attach --------------------------------------------------------------------- */
extern "C" int bbpcg_comm_export(bbpcg_solver *s, void *blob)
{
  if (!s || !blob) { bbpcg_set_error("bbpcg_comm_export: NULL argument"); return BBPCG_EINVAL; }
  CU(cudaSetDevice(s->device));
  ExportBlob b;
  memset(&b, 0, sizeof(b));
  b.magic = BB_MAGIC; b.rank = s->dom.rank; b.in = s->dev.L.in; b.jn = s->dev.L.jn; b.kn = s->dev.L.kn;
  b.device = s->device; b.pid = (long long)getpid();
  b.arena_ptr = (unsigned long long)(uintptr_t)s->arena; b.arena_bytes = s->amap.total;
  CU(cudaIpcGetMemHandle(&b.handle, s->arena));
  memcpy(b.uuid, s->uuid, 16);
  memset(blob, 0, BBPCG_BLOB_BYTES);
  memcpy(blob, &b, sizeof(b));
  return BBPCG_OK;
}

extern "C" int bbpcg_comm_import(bbpcg_solver *s, const void *all_blobs, int nranks)
{
  if (!s || !all_blobs || nranks < 1 || nranks > BB_MAXR) { bbpcg_set_error("bbpcg_comm_import: bad arguments (nranks %d, max %d)", nranks, BB_MAXR); return BBPCG_EINVAL; }
  if (nranks != s->DOM.In * s->DOM.Jn * s->DOM.Kn) { bbpcg_set_error("bbpcg_comm_import: %d ranks but the decomposition has %d blocks", nranks, s->DOM.In * s->DOM.Jn * s->DOM.Kn); return BBPCG_EINVAL; }
  CU(cudaSetDevice(s->device));
  int dims[BB_MAXR][3];
  const long long mypid = (long long)getpid();
  for (int p = 0; p < nranks; p++) {
    ExportBlob b;
    memcpy(&b, (const char *)all_blobs + (size_t)p * BBPCG_BLOB_BYTES, sizeof(b));
    if (b.magic != BB_MAGIC || b.rank != p) { bbpcg_set_error("bbpcg_comm_import: record %d is not rank %d's export", p, p); return BBPCG_ECOMM; }
    dims[p][0] = b.in; dims[p][1] = b.jn; dims[p][2] = b.kn;
    if (p == s->dom.rank) { s->peer_arena[p] = s->arena; continue; }
    if (s->peer_opened[p]) continue;                       /* a repeated import: this peer's mapping is still open */
    if ((b.pid == mypid && b.device == s->device) || !memcmp(b.uuid, s->uuid, 16)) s->shared_device = 1;
    if (b.pid == mypid) {
      /* same process (several ranks driven from one process): the pointer is directly usable;
       * a different device needs peer access */
      if (b.device != s->device) {
        int can = 0;
        CU(cudaDeviceCanAccessPeer(&can, s->device, b.device));
        if (!can) { bbpcg_set_error("device %d cannot access device %d", s->device, b.device); return BBPCG_ECOMM; }
        cudaError_t e = cudaDeviceEnablePeerAccess(b.device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { bbpcg_set_error("cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(e)); return BBPCG_ECOMM; }
        cudaGetLastError();
      }
      s->peer_arena[p] = (char *)(uintptr_t)b.arena_ptr;
    } else {
      void *ptr = NULL;
      cudaError_t e = cudaIpcOpenMemHandle(&ptr, b.handle, cudaIpcMemLazyEnablePeerAccess);
      if (e != cudaSuccess) {
        bbpcg_set_error("cudaIpcOpenMemHandle(rank %d): %s", p, cudaGetErrorString(e)); cudaGetLastError();
        for (int o = 0; o < p; o++) if (s->peer_opened[o]) { cudaIpcCloseMemHandle(s->peer_arena[o]); s->peer_opened[o] = false; s->peer_arena[o] = NULL; }
        return BBPCG_ECOMM;
      }
      s->peer_arena[p] = (char *)ptr; s->peer_opened[p] = true;
    }
  }
  s->nranks = nranks;
  build_halo(s, dims);
  s->plan_ok = 0;                   /* the neighbour tensor maps are part of the plan */
  return BBPCG_OK;
}

/* ---- launch helpers ------------------------------------------------------------------------ */
/* Tile / z-chunk plan of the two iteration kernels.  A work item is (x-tile of 128, y-tile of ty rows, z-chunk); the
 * kernels run as ONE wave of resident CTAs (2 per SM) that claim items in index order -- x fastest, then y, then z-chunk,
 * so the y/x neighbours of one chunk are in flight together and their halo rows hit L2.  Measured (scripts/sweep.py,
 * scripts/trace_timeline.py; profiles/r02b_sweep*.jsonl, r02d_sweep*.jsonl, r02fg_trace_per_cta.jsonl):
 *   - many items (512^3: 256 columns): uniform ~24-plane chunks (longer ones lose L2 hits on the halo rows their y
 *     neighbours fetched, shorter ones re-read more halo planes);
 *   - few items (256^3 block, the 8-GPU share of 512^3): identical CTAs of a static one-wave split finished between 84 and
 *     135 us depending on the SM group they ran on, so the chunk lengths are GUIDED: every column is cut into chunks of
 *     decreasing length (a claim takes `guided_pct` = 60 % of cols/slots of what is left, never less than `chunk_min` = 4
 *     planes: 12 cost 35 % on a 256 x 128 x 128 block);
 *     fast SMs claim more of the short tail chunks and all CTAs finish together.
 * Tile height: 8 rows (fewest halo-row re-reads) unless option `ty` says otherwise; the kernels take any 1..8.
 * Option `kc` forces uniform chunks.  Uploads the chunk table (Dev::ztab) and rebuilds the tensor maps. */
/* The planner proper: pure host arithmetic, exported so that the CPU tests can hold it to its contract on any block shape
 * (every plane 1..kn in exactly one chunk, chunk lengths >= 1, boundary chunks first) without a GPU.
 * ztab[2c], ztab[2c+1] = first and last plane of chunk c in CLAIM order; plan[5] = { ty, nbx, nby, nbz, planes of chunk 0 }. */
extern "C" int bbpcg_plan_zchunks(int in, int jn, int kn, int slots, int opt_ty, int opt_kc, int guided, int guided_pct, int chunk_min,
                                  int *ztab, int ztab_cap, int *plan)
{
  if (in < 1 || jn < 1 || kn < 1 || slots < 1 || !ztab || !plan) { bbpcg_set_error("bbpcg_plan_zchunks: bad argument"); return BBPCG_EINVAL; }
  const int nbx = (in + 127) / 128;
  const int kc_uniform = opt_kc > 0 ? opt_kc : 24;
  const long long nz_uniform = (kn + kc_uniform - 1) / kc_uniform;
  const bool uniform = opt_kc > 0 || !guided || (long long)nbx * ((jn + BB_TYMAX - 1) / BB_TYMAX) * nz_uniform >= 7ll * slots;
  /* 8 rows: fewest halo-row re-reads.  Guided regime, four interleaved passes per setting on one box (profiles/r02y_sweep.jsonl,
   * medians): claims of 60 % of the even share down to 4 planes with 8 rows 220.3 us per iteration at 256^3 and 210.4 at
   * 512 x 256 x 128, against 225.0 / 217.8 for the earlier plan (100 %, 8 planes, 7 rows); boxes differ by more than that. */
  const int best_ty = opt_ty > 0 ? (opt_ty < BB_TYMAX ? opt_ty : BB_TYMAX) : BB_TYMAX;
  const int cols = nbx * ((jn + best_ty - 1) / best_ty);
  std::vector<int> sz;
  if (uniform) {
    int nz = (int)(nz_uniform > BB_MAXZ ? BB_MAXZ : nz_uniform);
    if ((long long)cols * nz > BB_MAXBLOCKS) nz = BB_MAXBLOCKS / cols;       /* the shortest chunks the reduction workspace allows */
    if (nz < 1) { bbpcg_set_error("grid too large for the reduction workspace"); return BBPCG_EINVAL; }
    const int kc = (kn + nz - 1) / nz;
    for (int r = kn; r > 0; r -= kc) sz.push_back(r < kc ? r : kc);
  } else {
    const int minc = chunk_min > 0 ? chunk_min : 4;
    const int pct = guided_pct > 0 ? guided_pct : 60;
    int r = kn;
    while (r > 0) {
      int c = (int)(((long long)r * cols * pct / 100 + slots - 1) / slots);
      if (c < minc) c = minc;
      if (c > r || r - c < (minc + 1) / 2) c = r;
      sz.push_back(c);
      r -= c;
    }
  }
  if (sz.size() > BB_MAXZ || 2 * sz.size() > (size_t)ztab_cap) { bbpcg_set_error("too many z-chunks"); return BBPCG_EINVAL; }
  if ((long long)cols * (long long)sz.size() > BB_MAXBLOCKS) { bbpcg_set_error("grid too large for the reduction workspace"); return BBPCG_EINVAL; }
  /* Placement of the chunks along z, in claim order: the first one ends at the top face k = kn, the second starts at the
   * bottom face k = 1, the others fill the middle upwards.  The residual kernel pushes the new r of the block's top / bottom
   * plane into the z neighbours' ghost planes (peer stores over NVLink when the block is split in z): claimed first, those
   * stores are under way a whole kernel before the rank barrier releases them, instead of in front of it. */
  const size_t n = sz.size();
  int lo = 1;
  for (size_t i = 0; i < n; i++) {
    if (i == 0 && n > 1) { ztab[0] = kn - sz[0] + 1; ztab[1] = kn; continue; }
    ztab[2 * i] = lo; ztab[2 * i + 1] = lo + sz[i] - 1;
    lo += sz[i];
  }
  plan[0] = best_ty; plan[1] = nbx; plan[2] = (jn + best_ty - 1) / best_ty; plan[3] = (int)n; plan[4] = sz[0];
  return BBPCG_OK;
}

static int make_plan(bbpcg_solver *s)
{
  if (s->plan_ok) return BBPCG_OK;
  const Layout &L = s->dev.L;
  /* the copy source must stay valid until the copy ran: it is only rewritten after a stream sync */
  CU(cudaStreamSynchronize(s->stream));
  int plan[5];
  int rc = bbpcg_plan_zchunks(L.in, L.jn, L.kn, s->sm_count * 2, s->opt_ty, s->opt_kc, s->guided, s->guided_pct, s->chunk_min, s->h_ztab, 2 * BB_MAXZ, plan);
  if (rc) return rc;
  CU(cudaMemcpyAsync((void *)s->dev.ztab, s->h_ztab, sizeof(int) * (2 * plan[3]), cudaMemcpyHostToDevice, s->stream));
  s->plan_ty = plan[0]; s->plan_nbx = plan[1]; s->plan_nby = plan[2]; s->plan_nbz = plan[3]; s->plan_kc = plan[4];
  rc = build_search_maps(s, s->plan_ty);
  if (rc) return rc;
  s->plan_ok = 1;
  return BBPCG_OK;
}

/* PDL policy: on (option pdl 0 switches it off), except when several ranks share one GPU.  With the round-1 kernels PDL cost
 * 13 us per iteration on decomposed 256^3 blocks (early-launched CTAs held SM slots while the last CTA waited for its
 * peers); with the persistent kernels a dependent CTA only becomes resident when a slot frees, and decomposed runs measure
 * the same with and without it (2 ranks x 256^3: 221.1 vs 221.4 us, profiles/r02l_comm_probe.jsonl), stand-alone 209 vs 213. */
static bool pdl_active(const bbpcg_solver *s)
{
  return !s->shared_device && s->pdl != 0;
}

/* launch with (optionally) the programmatic-stream-serialization attribute: the kernel may be scheduled
 * while its predecessor drains and blocks in pdl_wait() until that one is complete.  Never used when
 * several ranks share one GPU: their waiting CTAs could starve the peer whose arrival they wait for. */
template <typename... KArgs, typename... Args>
static cudaError_t launch_k(bbpcg_solver *s, void (*kernel)(KArgs...), dim3 grid, int nt, size_t smem, bool pdl, Args &&...args)
{
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = dim3(nt); cfg.dynamicSmemBytes = smem; cfg.stream = s->stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = (pdl && pdl_active(s) && !s->kt_now) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

static SearchArgs plan_args(const bbpcg_solver *s)
{
  SearchArgs a;
  memset(&a, 0, sizeof(a));
  a.nbx = s->plan_nbx; a.nby = s->plan_nby; a.nbz = s->plan_nbz; a.ty = s->plan_ty;
  a.producer = s->tma_warp ? BB_PRODUCER : 0;
  a.launch = (int)(s->launches & 0x7fffffff);
  a.nitems = a.nbx * a.nby * a.nbz;
  return a;
}

/* one wave of resident CTAs (2 per SM, bounded by launch_bounds and the shared-memory size), never more than items */
static unsigned iter_grid(const bbpcg_solver *s, const SearchArgs &a)
{
  const int slots = s->sm_count * 2;
  return (unsigned)(a.nitems < slots ? a.nitems : slots);
}

/* k_search_tma (bbpcg_search_tma.cuh) */
static int launch_search(bbpcg_solver *s, bool parts)
{
  static_assert(sizeof(Dev) + sizeof(SearchMaps) + sizeof(SearchArgs) + 192 <= 4096, "kernel parameters exceed 4 KB");
  int rc = make_plan(s);
  if (rc) return rc;
  const SearchArgs a = plan_args(s);
  const dim3 grid(iter_grid(s, a));
  if (parts) CU(launch_k(s, k_search_tma<true, 2>, grid, BB_NT_ITER, SearchGeom<true, 2>::SMEM, true, s->dev, s->maps, a));
  else CU(launch_k(s, k_search_tma<false, 2>, grid, BB_NT_ITER, SearchGeom<false, 2>::SMEM, true, s->dev, s->maps, a));
  s->launches++;
  return BBPCG_OK;
}

/* k_resid_tma (bbpcg_resid_tma.cuh): same tiles and z-chunks as the search kernel.  rhs != NULL: the true-residual
 * refresh form r = b - (-A x) */
static int launch_resid(bbpcg_solver *s, bool parts, const real *rhs)
{
  int rc = make_plan(s);
  if (rc) return rc;
  SearchArgs a = plan_args(s);
  a.rhs = rhs; a.s1b = s->fst.cs1b; a.s2b = s->fst.cs2b;
  const dim3 grid(iter_grid(s, a));
  if (rhs) {
    if (parts) CU(launch_k(s, k_resid_tma<true, 2, true>, grid, BB_NT_ITER, ResidGeom<true, 2>::SMEM, false, s->dev, s->maps, a));
    else CU(launch_k(s, k_resid_tma<false, 2, true>, grid, BB_NT_ITER, ResidGeom<false, 2>::SMEM, false, s->dev, s->maps, a));
  } else {
    if (parts) CU(launch_k(s, k_resid_tma<true, 2, false>, grid, BB_NT_ITER, ResidGeom<true, 2>::SMEM, true, s->dev, s->maps, a));
    else CU(launch_k(s, k_resid_tma<false, 2, false>, grid, BB_NT_ITER, ResidGeom<false, 2>::SMEM, true, s->dev, s->maps, a));
  }
  s->launches++;
  return BBPCG_OK;
}

template <int XT, int UNR>
static int launch_refresh_x4_t(bbpcg_solver *s)
{
  const Layout &L = s->dev.L;
  constexpr int YT = 128 / XT;
  ResidArgs a;
  a.cpr = (L.in + 3) / 4;
  a.ncb = (a.cpr + XT - 1) / XT;
  const long long nrows = (long long)L.jn * L.kn;
  const long long npass = (nrows + YT * UNR - 1) / (YT * UNR) * a.ncb;
  if (npass > 0x7fffffffll) { bbpcg_set_error("block too large"); return BBPCG_EINVAL; }
  a.npass = (int)npass;
  a.ppc = 1;
  if (a.npass > BB_MAXBLOCKS) a.ppc = (a.npass + BB_MAXBLOCKS - 1) / BB_MAXBLOCKS;
  k_refresh_x4<XT, UNR><<<(a.npass + a.ppc - 1) / a.ppc, 128, 0, s->stream>>>(s->dev, a);
  s->launches++;
  return BBPCG_OK;
}

static int launch_refresh_x4(bbpcg_solver *s)
{
  const int cpr = (s->dev.L.in + 3) / 4;
  if (cpr > 64) return launch_refresh_x4_t<128, 4>(s);
  if (cpr > 32) return launch_refresh_x4_t<64, 4>(s);
  return launch_refresh_x4_t<32, 4>(s);
}

/* Load every kernel of the library NOW.  CUDA loads device functions lazily at their first
 * launch, and that load takes a context-wide lock; when several ranks share one context (the
 * single-process test harness) a first launch on one rank's host thread would block the other
 * ranks' launches while a collective kernel already on the device waits for them. */
template <typename K> static int preload_one(K kernel, int dyn_smem = 0)
{
  cudaFuncAttributes at;
  CU(cudaFuncGetAttributes(&at, kernel));
  /* the > 48 KB opt-in is a per-device function attribute: set it here, once per solver (= per device), not behind a
   * process-wide flag at the first launch */
  if (dyn_smem > 48 * 1024) CU(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn_smem));
  return BBPCG_OK;
}
static int preload_kernels()
{
  int rc = 0;
#define PL(...) if (!rc) rc = preload_one(__VA_ARGS__)
  PL(k_search_tma<false, 2>, SearchGeom<false, 2>::SMEM); PL(k_search_tma<true, 2>, SearchGeom<true, 2>::SMEM);
  PL(k_resid_tma<false, 2, false>, ResidGeom<false, 2>::SMEM); PL(k_resid_tma<true, 2, false>, ResidGeom<true, 2>::SMEM);
  PL(k_resid_tma<false, 2, true>, ResidGeom<false, 2>::SMEM); PL(k_resid_tma<true, 2, true>, ResidGeom<true, 2>::SMEM);
  PL(k_refresh_x4<128, 4>); PL(k_refresh_x4<64, 4>); PL(k_refresh_x4<32, 4>);
  PL(k_build_tab); PL(k_init<256>); PL(k_finish<256>); PL(k_rhs<256>); PL(k_rhs_tiled); PL(k_masks<256>); PL(k_part_rhs_net);
  PL(k_coeffs_refine<256>); PL(k_zero_ghosts); PL(k_xchg_send); PL(k_xchg_recv);
  PL(k_spmv_s3b<256, false>); PL(k_spmv_s3b<256, true>);
  PL(k_solv_sum); PL(k_solv_apply); PL(k_bc_star);
  PL(k_cage_reset); PL(k_cage<false>); PL(k_cage<true>); PL(k_cage_flags<256>);
  PL(k_bc_p); PL(k_sub_mean);
  PL(k_epi_uwp<true, true>); PL(k_epi_uwp<true, false>); PL(k_epi_uwp<false, true>); PL(k_epi_v);
  PL(k_epilogue<true, true>, EPI_SMEM); PL(k_epilogue<true, false>, EPI_SMEM); PL(k_epilogue<false, true>, EPI_SMEM);
#undef PL
  return rc;
}

static int clampi(long long v, int lo, int hi) { return (int)(v < lo ? lo : v > hi ? hi : v); }

/* The in-kernel collectives give up after Comm::timeout_cycles instead of hanging (the reference's MPI calls block for
 * ever).  A time-out leaves the ranks with different sums and ghosts, so it is FATAL for the solver objects of all ranks:
 * the device sets a pinned host word, every collective entry point reads it after its stream sync (and on entry) and
 * returns BBPCG_ECOMM from then on; destroy and re-create the solvers to recover. */
static int comm_check(const bbpcg_solver *s, const char *who)
{
  if (((volatile int *)s->h_poll)[BB_POLL_COMM]) {
    bbpcg_set_error("%s: a peer rank did not arrive within the collective time-out (option comm_timeout_ms); ghosts and sums are "
                    "inconsistent across ranks -- destroy and re-create the solver on every rank", who);
    return BBPCG_ECOMM;
  }
  return BBPCG_OK;
}

/* ---- coefficients -------------------------------------------------------------------------- */
extern "C" int bbpcg_set_coefficients(bbpcg_solver *s, const int *flag_u, const int *flag_v, const int *flag_w, const int *phase)
{
  if (!s || !flag_u || !flag_v || !flag_w) { bbpcg_set_error("bbpcg_set_coefficients: NULL flag array"); return BBPCG_EINVAL; }
  CU(cudaSetDevice(s->device));
  if (s->nranks == 1 && s->DOM.In * s->DOM.Jn * s->DOM.Kn != 1) { bbpcg_set_error("decomposition has %d blocks: call bbpcg_comm_import first", s->DOM.In * s->DOM.Jn * s->DOM.Kn); return BBPCG_ECOMM; }
  const long long nrows = (long long)s->dev.L.jn * s->dev.L.kn;
  k_masks<256><<<clampi(nrows, 1, s->stream_blocks), 256, 0, s->stream>>>(s->dev, s->fst, flag_u, flag_v, flag_w, phase);
  s->launches++;
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(s->stream));
  if (comm_check(s, "bbpcg_set_coefficients")) return BBPCG_ECOMM;
  s->has_phase = phase != NULL;
  s->coeffs_set = 1;
  return BBPCG_OK;
}

/* ---- coefficient producers: cuda_build_cages (+ the mask digestion of cuda_PP_init_jacobi_preconditioner), bbpcg_cages.cuh ---- */
extern "C" int bbpcg_build_cages(bbpcg_solver *s, int NPARTS, int nparts, const bbpcg_parts_view *parts,
                                 int *flag_u, int *flag_v, int *flag_w, int *phase, int *phase_shell)
{
  if (!s || !flag_u || !flag_v || !flag_w) { bbpcg_set_error("bbpcg_build_cages: NULL flag array"); return BBPCG_EINVAL; }
  if (NPARTS > 0 && (!phase || !phase_shell)) { bbpcg_set_error("bbpcg_build_cages: NPARTS > 0 needs phase and phase_shell"); return BBPCG_EINVAL; }
  if (NPARTS > 0 && nparts > 0 && (!parts || !parts->base)) { bbpcg_set_error("bbpcg_build_cages: nparts > 0 needs the particle view"); return BBPCG_EINVAL; }
  if (nparts < 0) { bbpcg_set_error("bbpcg_build_cages: nparts < 0"); return BBPCG_EINVAL; }
  if (s->nranks == 1 && s->DOM.In * s->DOM.Jn * s->DOM.Kn != 1) { bbpcg_set_error("decomposition has %d blocks: call bbpcg_comm_import first", s->DOM.In * s->DOM.Jn * s->DOM.Kn); return BBPCG_ECOMM; }
  CU(cudaSetDevice(s->device));
  if (comm_check(s, "bbpcg_build_cages")) return BBPCG_ECOMM;
  const dom_struct &d = s->dom;
  const dom_struct &D = s->DOM;
  const grid_info &g = d.Gcc;
  if (NPARTS > 0) {                                                        /* cuda_particle.cu:1524-1598 */
    k_cage_reset<<<clampi(((long long)g.s3b + 255) / 256, 1, s->sm_count * 16), 256, 0, s->stream>>>(phase, phase_shell, (long long)g.s3b);
    s->launches++;
    if (nparts > 0) {
      CageArgs a;
      memset(&a, 0, sizeof(a));
      a.parts = (const char *)parts->base; a.stride = parts->stride; a.ox = parts->off_x; a.oy = parts->off_y; a.oz = parts->off_z; a.orad = parts->off_r;
      a.nparts = nparts;
      a.xs = d.xs; a.ys = d.ys; a.zs = d.zs; a.dx = d.dx; a.dy = d.dy; a.dz = d.dz;
      a.xn = d.xn; a.yn = d.yn; a.zn = d.zn;
      /* a cage is clipped to _is.._ie on a non-periodic edge of the GLOBAL domain, to _isb.._ieb elsewhere (particle_kernel.cu:177-211) */
      a.S[0] = (d.I == D.Is && s->bc.pW != BB_PERIODIC) ? g._is : g._isb;  a.E[0] = (d.I == D.Ie && s->bc.pE != BB_PERIODIC) ? g._ie : g._ieb;
      a.S[1] = (d.J == D.Js && s->bc.pS != BB_PERIODIC) ? g._js : g._jsb;  a.E[1] = (d.J == D.Je && s->bc.pN != BB_PERIODIC) ? g._je : g._jeb;
      a.S[2] = (d.K == D.Ks && s->bc.pB != BB_PERIODIC) ? g._ks : g._ksb;  a.E[2] = (d.K == D.Ke && s->bc.pT != BB_PERIODIC) ? g._ke : g._keb;
      a.s1b = g.s1b; a.s2b = g.s2b;
      a.phase = phase; a.phase_shell = phase_shell;
      k_cage<false><<<nparts, 256, 0, s->stream>>>(a);                    /* every build_phase before any build_phase_shell (:1541-1584) */
      k_cage<true><<<nparts, 256, 0, s->stream>>>(a);
      s->launches += 2;
    }
  }
  CageFlagArgs f;
  f.flag_u = flag_u; f.flag_v = flag_v; f.flag_w = flag_w;
  f.phase = NPARTS > 0 ? phase : NULL; f.phase_shell = NPARTS > 0 ? phase_shell : NULL;
  /* flag_external_*: only when neither side of the axis is periodic and the block touches the global face (cuda_particle.cu:1605-1639) */
  f.ext = 0;
  if (s->bc.pW != BB_PERIODIC && s->bc.pE != BB_PERIODIC) f.ext |= (d.I == D.Is ? 1u : 0u) | (d.I == D.Ie ? 2u : 0u);
  if (s->bc.pS != BB_PERIODIC && s->bc.pN != BB_PERIODIC) f.ext |= (d.J == D.Js ? 4u : 0u) | (d.J == D.Je ? 8u : 0u);
  if (s->bc.pB != BB_PERIODIC && s->bc.pT != BB_PERIODIC) f.ext |= (d.K == D.Ks ? 16u : 0u) | (d.K == D.Ke ? 32u : 0u);
  const long long nrows = (long long)(s->dev.L.jn + 2) * (s->dev.L.kn + 2);
  k_cage_flags<256><<<clampi(nrows, 1, s->stream_blocks), 256, 0, s->stream>>>(s->dev, s->fst, f);
  s->launches++;
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(s->stream));
  if (comm_check(s, "bbpcg_build_cages")) return BBPCG_ECOMM;
  s->has_phase = NPARTS > 0;
  s->coeffs_set = 1;
  return BBPCG_OK;
}

/* ---- right-hand side ----------------------------------------------------------------------- */
static int enqueue_rhs(bbpcg_solver *s, const real *u, const real *v, const real *w, real rho_f, real dt, real *rhs)
{
  const dom_struct &d = s->dom;
  CU(cudaMemsetAsync(rhs, 0, sizeof(real) * (size_t)d.Gcc.s3b, s->stream));            /* cuda_solver.cu:122 */
  const long long nrows = (long long)d.Gcc.jn * d.Gcc.kn;
  if (s->rhs_tiled) {
    const long long nb = (long long)((d.Gcc.in + RHS_TI - 1) / RHS_TI) * ((d.Gcc.jn + RHS_TJ - 1) / RHS_TJ) * ((d.Gcc.kn + RHS_TK - 1) / RHS_TK);
    if (nb > 0x7fffffffll) { bbpcg_set_error("block too large"); return BBPCG_EINVAL; }
    k_rhs_tiled<<<(unsigned)nb, 256, 0, s->stream>>>(d.Gcc.in, d.Gcc.jn, d.Gcc.kn, s->fst, u, v, w, rhs, 1. / d.dx, 1. / d.dy, 1. / d.dz, rho_f / dt);
  } else
    k_rhs<256><<<clampi(nrows, 1, s->stream_blocks), 256, 0, s->stream>>>(d.Gcc.in, d.Gcc.jn, d.Gcc.kn, s->fst, u, v, w, rhs,
                                                                          1. / d.dx, 1. / d.dy, 1. / d.dz, rho_f / dt);
  s->launches++;
  return BBPCG_OK;
}

extern "C" int bbpcg_rhs(bbpcg_solver *s, const real *u, const real *v, const real *w, real rho_f, real dt, real *rhs)
{
  if (!s || !u || !v || !w || !rhs) { bbpcg_set_error("bbpcg_rhs: NULL argument"); return BBPCG_EINVAL; }
  CU(cudaSetDevice(s->device));
  int rc = enqueue_rhs(s, u, v, w, rho_f, dt, rhs);
  if (rc) return rc;
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(s->stream));
  return BBPCG_OK;
}

/* ---- halo exchange on caller arrays -------------------------------------------------------- */
static XchgGrid xchg_grid(const bbpcg_solver *s, int grid)
{
  const dom_struct &d = s->dom;
  XchgGrid g;
  const grid_info &gi = grid == BBPCG_GFX ? d.Gfx : grid == BBPCG_GFY ? d.Gfy : grid == BBPCG_GFZ ? d.Gfz : d.Gcc;
  g.n[0] = gi.in; g.n[1] = gi.jn; g.n[2] = gi.kn;
  /* index macros, src/bluebottle.h:70-73 */
  if (grid == BBPCG_GFX) { g.st[0] = gi.s2b; g.st[1] = 1; g.st[2] = gi.s1b; }
  else if (grid == BBPCG_GFY) { g.st[0] = gi.s1b; g.st[1] = gi.s2b; g.st[2] = 1; }
  else { g.st[0] = 1; g.st[1] = gi.s1b; g.st[2] = gi.s2b; }
  for (int ax = 0; ax < 3; ax++) { g.send_hi[ax] = g.n[ax]; g.send_lo[ax] = 1; }         /* _ie -> nbr _isb, _is -> nbr _ieb */
  const int normal = grid == BBPCG_GFX ? 0 : grid == BBPCG_GFY ? 1 : grid == BBPCG_GFZ ? 2 : -1;
  if (normal >= 0) { g.send_hi[normal] = g.n[normal] - 1; g.send_lo[normal] = 2; }       /* _ie-1 / _is+1: the shared face is skipped */
  return g;
}

static int enqueue_exchange(bbpcg_solver *s, real *array, int grid = BBPCG_GCC)
{
  const XchgGrid g = xchg_grid(s, grid);
  const long long total = 2ll * ((long long)g.n[1] * g.n[2] + (long long)g.n[0] * g.n[2] + (long long)g.n[0] * g.n[1]);
  const int nb = clampi((total + 255) / 256, 1, s->sm_count * 4);
  const int buf = (int)(s->exchange_count++ & 1u);
  k_xchg_send<<<nb, 256, 0, s->stream>>>(s->dev, g, array, buf);
  k_xchg_recv<<<nb, 256, 0, s->stream>>>(s->dev, g, array, buf);
  s->launches += 2;
  return BBPCG_OK;
}

extern "C" int bbpcg_exchange(bbpcg_solver *s, real *array, int grid)
{
  if (!s || !array) { bbpcg_set_error("bbpcg_exchange: NULL argument"); return BBPCG_EINVAL; }
  if (grid < BBPCG_GCC || grid > BBPCG_GFZ) { bbpcg_set_error("bbpcg_exchange: grid must be BBPCG_GCC/GFX/GFY/GFZ"); return BBPCG_EINVAL; }
  if (s->nranks == 1 && s->DOM.In * s->DOM.Jn * s->DOM.Kn != 1) { bbpcg_set_error("decomposition has %d blocks: call bbpcg_comm_import first", s->DOM.In * s->DOM.Jn * s->DOM.Kn); return BBPCG_ECOMM; }
  CU(cudaSetDevice(s->device));
  if (comm_check(s, "bbpcg_exchange")) return BBPCG_ECOMM;
  int rc = enqueue_exchange(s, array, grid);
  if (rc) return rc;
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(s->stream));
  return comm_check(s, "bbpcg_exchange");
}

extern "C" int bbpcg_exchange_Gcc(bbpcg_solver *s, real *array) { return bbpcg_exchange(s, array, BBPCG_GCC); }

extern "C" int bbpcg_spmv(bbpcg_solver *s, const real *src_s3b, real *Ap_s3, int use_phase)
{
  if (!s || !src_s3b || !Ap_s3) { bbpcg_set_error("bbpcg_spmv: NULL argument"); return BBPCG_EINVAL; }
  if (!s->coeffs_set) { bbpcg_set_error("bbpcg_spmv: call bbpcg_set_coefficients first"); return BBPCG_EINVAL; }
  if (use_phase && !s->has_phase) { bbpcg_set_error("bbpcg_spmv: use_phase without a phase array in bbpcg_set_coefficients"); return BBPCG_EINVAL; }
  CU(cudaSetDevice(s->device));
  const grid_info &g = s->dom.Gcc;
  const long long nrows = (long long)g.jn * g.kn;
  const int nb = clampi(nrows, 1, s->stream_blocks);
  if (use_phase) k_spmv_s3b<256, true><<<nb, 256, 0, s->stream>>>(s->dev, src_s3b, g.s1b, g.s2b, Ap_s3, g.s1, g.s2);
  else k_spmv_s3b<256, false><<<nb, 256, 0, s->stream>>>(s->dev, src_s3b, g.s1b, g.s2b, Ap_s3, g.s1, g.s2);
  s->launches++;
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(s->stream));
  return BBPCG_OK;
}

/* ---- solve epilogue: cuda_dom_BC_p / cuda_project / cuda_update_p (bbpcg_epilogue.cuh) ------- */
static unsigned neumann_wall_faces(const bbpcg_solver *s)
{
  /* cuda_dom_BC_p: `dom[rank].w == MPI_PROC_NULL` and `bc.pW == NEUMANN` (cuda_bluebottle.cu:2541-2587) */
  const int bct[6] = { s->bc.pE, s->bc.pW, s->bc.pN, s->bc.pS, s->bc.pT, s->bc.pB };
  unsigned m = 0;
  for (int f = 0; f < 6; f++) if (nbr_rank(s->dom, f) < 0 && bct[f] == BB_NEUMANN) m |= 1u << f;
  return m;
}

static int enqueue_bc_p(bbpcg_solver *s, real *array)
{
  const unsigned faces = neumann_wall_faces(s);
  if (!faces) return BBPCG_OK;
  const Layout &L = s->dev.L;
  const long long total = 2ll * ((long long)L.jn * L.kn + (long long)L.in * L.kn + (long long)L.in * L.jn);
  k_bc_p<<<clampi((total + 255) / 256, 1, s->sm_count * 8), 256, 0, s->stream>>>(L.in, L.jn, L.kn, s->fst.cs1b, s->fst.cs2b, array, faces);
  s->launches++;
  return BBPCG_OK;
}

extern "C" int bbpcg_dom_BC_p(bbpcg_solver *s, real *array)
{
  if (!s || !array) { bbpcg_set_error("bbpcg_dom_BC_p: NULL argument"); return BBPCG_EINVAL; }
  CU(cudaSetDevice(s->device));
  int rc = enqueue_bc_p(s, array);
  if (rc) return rc;
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(s->stream));
  return BBPCG_OK;
}

extern "C" int bbpcg_epilogue(bbpcg_solver *s, const bbpcg_epilogue_args *a, double *ms_out)
{
  if (!s || !a || !a->phi) { bbpcg_set_error("bbpcg_epilogue: NULL argument"); return BBPCG_EINVAL; }
  const bool project = a->u != NULL, update = a->p != NULL;
  if (project && (!a->v || !a->w || !a->u_star || !a->v_star || !a->w_star || !a->flag_u || !a->flag_v || !a->flag_w)) {
    bbpcg_set_error("bbpcg_epilogue: cuda_project needs u,v,w, u*,v*,w* and the three flag arrays"); return BBPCG_EINVAL;
  }
  if (update && (!a->p0 || !a->phase)) { bbpcg_set_error("bbpcg_epilogue: cuda_update_p needs p0 and phase"); return BBPCG_EINVAL; }
  if (!project && !update) { bbpcg_set_error("bbpcg_epilogue: nothing to do (u == NULL and p == NULL)"); return BBPCG_EINVAL; }
  if (s->nranks == 1 && s->DOM.In * s->DOM.Jn * s->DOM.Kn != 1) { bbpcg_set_error("decomposition has %d blocks: call bbpcg_comm_import first", s->DOM.In * s->DOM.Jn * s->DOM.Kn); return BBPCG_ECOMM; }
  CU(cudaSetDevice(s->device));
  const dom_struct &d = s->dom;
  const Layout &L = s->dev.L;
  CU(cudaEventRecord(s->ev[0], s->stream));
  if (!a->phi_ghosts_valid) {                               /* bluebottle.c:233-234 */
    int rc = enqueue_exchange(s, a->phi);
    if (!rc) rc = enqueue_bc_p(s, a->phi);
    if (rc) return rc;
  }
  EpiArgs e;
  memset(&e, 0, sizeof(e));
  e.u_star = a->u_star; e.v_star = a->v_star; e.w_star = a->w_star;
  e.flag_u = a->flag_u; e.flag_v = a->flag_v; e.flag_w = a->flag_w;
  e.u = a->u; e.v = a->v; e.w = a->w; e.phi = a->phi; e.p0 = a->p0; e.phase = a->phase; e.p = a->p;
  e.ddx = 1. / d.dx; e.ddy = 1. / d.dy; e.ddz = 1. / d.dz;
  e.dt_rho = a->dt / a->rho_f;
  e.nti = (L.in + EPI_T - 1) / EPI_T; e.ntj = (L.jn + EPI_T - 1) / EPI_T; e.ntk = (L.kn + EPI_T - 1) / EPI_T;
  if (s->epilogue_tiled) {            /* the 16^3-brick kernel (kept selectable; the streaming pair below is the default) */
    const long long ntiles = (long long)e.nti * e.ntj * e.ntk;
    const int grid = clampi(ntiles, 1, BB_MAXBLOCKS);
    if (project && update) k_epilogue<true, true><<<grid, 256, EPI_SMEM, s->stream>>>(s->dev, s->fst, e);
    else if (project) k_epilogue<true, false><<<grid, 256, EPI_SMEM, s->stream>>>(s->dev, s->fst, e);
    else k_epilogue<false, true><<<grid, 256, EPI_SMEM, s->stream>>>(s->dev, s->fst, e);
    s->launches++;
  } else {
    EpiPlan pl;
    pl.nti = (L.in + EA_T - 1) / EA_T; pl.ntj = (L.jn + EA_T - 1) / EA_T; pl.ntk = (L.kn + EA_T - 1) / EA_T;
    /* chunks of ~32 planes / faces: enough items for every resident CTA, one extra phi plane per chunk (3 %) */
    pl.kc = s->epi_chunk > 0 ? s->epi_chunk : 32; pl.nzc = (L.kn + pl.kc - 1) / pl.kc;
    pl.jc = pl.kc; pl.njc = (L.jn + 1 + pl.jc - 1) / pl.jc;
    const long long itemsA = (long long)pl.nti * pl.ntj * pl.nzc, itemsB = (long long)pl.ntk * pl.nti * pl.njc;
    const int gridA = clampi(itemsA, 1, s->sm_count * 2), gridB = clampi(itemsB, 1, s->sm_count * 3);   /* 109 / 72 registers */
    if (project && update) k_epi_uwp<true, true><<<gridA, 256, 0, s->stream>>>(s->dev, s->fst, e, pl);
    else if (project) k_epi_uwp<true, false><<<gridA, 256, 0, s->stream>>>(s->dev, s->fst, e, pl);
    else k_epi_uwp<false, true><<<gridA, 256, 0, s->stream>>>(s->dev, s->fst, e, pl);
    s->launches++;
    if (project) { k_epi_v<<<gridB, 256, 0, s->stream>>>(s->dev, s->fst, e, pl); s->launches++; }
  }
  if (update) {
    const long long nrows = (long long)L.jn * L.kn;
    k_sub_mean<<<clampi(nrows, 1, s->sm_count * 16), 256, 0, s->stream>>>(s->dev, a->p, s->fst.cs1b, s->fst.cs2b, (double)s->DOM.xn * (double)s->DOM.yn * (double)s->DOM.zn);
    s->launches++;
  }
  CU(cudaEventRecord(s->ev[1], s->stream));
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(s->stream));
  if (ms_out) { float ms = 0.f; cudaEventElapsedTime(&ms, s->ev[0], s->ev[1]); *ms_out = ms; }
  return comm_check(s, "bbpcg_epilogue");
}

/* ---- solve prologue: cuda_solvability (bbpcg_epilogue.cuh) -------------------------------------------------- */
static int enqueue_solvability(bbpcg_solver *s, real *u_star, real *v_star, real *w_star, int out_plane)
{
  const dom_struct &d = s->dom;
  SolvArgs a;
  a.u = u_star; a.v = v_star; a.w = w_star; a.out_plane = out_plane;
  a.planes = (d.I == 0 ? 1u : 0u) | (d.I == s->DOM.In - 1 ? 2u : 0u) | (d.J == 0 ? 4u : 0u) | (d.J == s->DOM.Jn - 1 ? 8u : 0u) |
             (d.K == 0 ? 16u : 0u) | (d.K == s->DOM.Kn - 1 ? 32u : 0u);              /* dom[rank].I == DOM.Is ... (cuda_bluebottle.cu:2332-2411) */
  a.ayz = d.dy * d.dz; a.azx = d.dz * d.dx; a.axy = d.dx * d.dy;
  a.Ayz = s->DOM.yl * s->DOM.zl; a.Azx = s->DOM.zl * s->DOM.xl; a.Axy = s->DOM.xl * s->DOM.yl;
  const Layout &L = s->dev.L;
  long long big = (long long)L.jn * L.kn;
  if ((long long)L.in * L.kn > big) big = (long long)L.in * L.kn;
  if ((long long)L.in * L.jn > big) big = (long long)L.in * L.jn;
  const int nb = clampi((big + 255) / 256, 1, s->sm_count * 2);
  k_solv_sum<<<nb, 256, 0, s->stream>>>(s->dev, s->fst, a);
  k_solv_apply<<<nb, 256, 0, s->stream>>>(s->dev, s->fst, a);
  s->launches += 2;
  return BBPCG_OK;
}

static int check_out_plane(int out_plane, const char *who)
{
  if (!((out_plane >= 0 && out_plane <= 5) || out_plane == 10)) { bbpcg_set_error("%s: out_plane must be WEST 0 .. TOP 5 or HOMOGENEOUS 10", who); return BBPCG_EINVAL; }
  return BBPCG_OK;
}

extern "C" int bbpcg_solvability(bbpcg_solver *s, real *u_star, real *v_star, real *w_star, int out_plane, real *eps_out)
{
  if (!s || !u_star || !v_star || !w_star) { bbpcg_set_error("bbpcg_solvability: NULL argument"); return BBPCG_EINVAL; }
  if (check_out_plane(out_plane, "bbpcg_solvability")) return BBPCG_EINVAL;
  if (s->nranks == 1 && s->DOM.In * s->DOM.Jn * s->DOM.Kn != 1) { bbpcg_set_error("decomposition has %d blocks: call bbpcg_comm_import first", s->DOM.In * s->DOM.Jn * s->DOM.Kn); return BBPCG_ECOMM; }
  CU(cudaSetDevice(s->device));
  int rc = enqueue_solvability(s, u_star, v_star, w_star, out_plane);
  if (rc) return rc;
  if (eps_out) CU(cudaMemcpyAsync(s->h_scal, s->dev.sc, sizeof(Scal), cudaMemcpyDeviceToHost, s->stream));
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(s->stream));
  if (eps_out) { eps_out[0] = s->h_scal->eps[0]; eps_out[1] = s->h_scal->eps[1]; eps_out[2] = s->h_scal->eps[2]; }
  return comm_check(s, "bbpcg_solvability");
}

/* ---- solve prologue: cuda_dom_BC_star (bbpcg_epilogue.cuh) -------------------------------------------------- */
/* one launch per axis whose faces carry a DIRICHLET / NEUMANN entry on a side without a neighbour
 * (`dom[rank].w == MPI_PROC_NULL`, cuda_bluebottle.cu:2114 ...) */
static int enqueue_bc_star(bbpcg_solver *s, real *u_star, real *v_star, real *w_star, const bb_velocity_bc *vbc)
{
  const dom_struct &d = s->dom;
  const grid_info *g[3] = { &d.Gfx, &d.Gfy, &d.Gfz };
  const int nbr[6] = { d.w, d.e, d.s, d.n, d.b, d.t };           /* the reference's face order W, E, S, N, B, T */
  BcStarArgs a;
  memset(&a, 0, sizeof(a));
  a.arr[0] = u_star; a.arr[1] = v_star; a.arr[2] = w_star;
  for (int c = 0; c < 3; c++) { a.n[c][0] = g[c]->in; a.n[c][1] = g[c]->jn; a.n[c][2] = g[c]->kn; }
  /* index macros, src/bluebottle.h:70-73 */
  a.st[0][0] = d.Gfx.s2b; a.st[0][1] = 1;         a.st[0][2] = d.Gfx.s1b;
  a.st[1][0] = d.Gfy.s1b; a.st[1][1] = d.Gfy.s2b; a.st[1][2] = 1;
  a.st[2][0] = 1;         a.st[2][1] = d.Gfz.s1b; a.st[2][2] = d.Gfz.s2b;
  for (int axis = 0; axis < 3; axis++) {
    a.axis = axis;
    long long work = 0;
    for (int c = 0; c < 3; c++) {
      int t1 = (axis + 1) % 3, t2 = (axis + 2) % 3;
      if (a.st[c][t1] > a.st[c][t2]) { int t = t1; t1 = t2; t2 = t; }
      a.a1[c] = t1; a.a2[c] = t2;
      for (int side = 0; side < 2; side++) {
        const int f = 2 * axis + side;
        const int ty = vbc->type[c][f];
        a.type[c][side] = (nbr[f] < 0 && (ty == BB_DIRICHLET || ty == BB_NEUMANN)) ? ty : 0;
        a.val[c][side] = vbc->val[c][f];
      }
      if (a.type[c][0] || a.type[c][1]) work += (long long)a.n[c][t1] * a.n[c][t2];
    }
    if (!work) continue;
    k_bc_star<<<clampi((work + 255) / 256, 1, s->sm_count * 8), 256, 0, s->stream>>>(a);
    s->launches++;
  }
  return BBPCG_OK;
}

extern "C" int bbpcg_dom_BC_star(bbpcg_solver *s, real *u_star, real *v_star, real *w_star, const bb_velocity_bc *vbc)
{
  if (!s || !u_star || !v_star || !w_star || !vbc) { bbpcg_set_error("bbpcg_dom_BC_star: NULL argument"); return BBPCG_EINVAL; }
  CU(cudaSetDevice(s->device));
  int rc = enqueue_bc_star(s, u_star, v_star, w_star, vbc);
  if (rc) return rc;
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(s->stream));
  return BBPCG_OK;
}

/* the whole particle-free prologue of src/bluebottle.c:213-225 enqueued back to back, one host synchronisation at the end */
extern "C" int bbpcg_prologue(bbpcg_solver *s, real *u_star, real *v_star, real *w_star, const bb_velocity_bc *vbc, int out_plane, double *ms_out)
{
  if (!s || !u_star || !v_star || !w_star || !vbc) { bbpcg_set_error("bbpcg_prologue: NULL argument"); return BBPCG_EINVAL; }
  if (check_out_plane(out_plane, "bbpcg_prologue")) return BBPCG_EINVAL;
  if (s->nranks == 1 && s->DOM.In * s->DOM.Jn * s->DOM.Kn != 1) { bbpcg_set_error("decomposition has %d blocks: call bbpcg_comm_import first", s->DOM.In * s->DOM.Jn * s->DOM.Kn); return BBPCG_ECOMM; }
  CU(cudaSetDevice(s->device));
  if (comm_check(s, "bbpcg_prologue")) return BBPCG_ECOMM;
  CU(cudaEventRecord(s->ev[0], s->stream));
  int rc = 0;
  for (int pass = 0; pass < 2 && !rc; pass++) {
    if (pass == 1) rc = enqueue_solvability(s, u_star, v_star, w_star, out_plane);             /* bluebottle.c:220 */
    if (!rc) rc = enqueue_bc_star(s, u_star, v_star, w_star, vbc);                            /* :214, :222 */
    if (!rc) rc = enqueue_exchange(s, u_star, BBPCG_GFX);                                     /* :215-217, :223-225 */
    if (!rc) rc = enqueue_exchange(s, v_star, BBPCG_GFY);
    if (!rc) rc = enqueue_exchange(s, w_star, BBPCG_GFZ);
  }
  if (rc) { cudaStreamSynchronize(s->stream); return rc; }
  CU(cudaEventRecord(s->ev[1], s->stream));
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(s->stream));
  if (ms_out) { float ms = 0.f; cudaEventElapsedTime(&ms, s->ev[0], s->ev[1]); *ms_out = ms; }
  return comm_check(s, "bbpcg_prologue");
}

/* ---- the solve ----------------------------------------------------------------------------- */
static int kt_limit(const bbpcg_solver *s) { return s->kernel_timing > 1 && s->kernel_timing < BB_KT_CAP ? s->kernel_timing : BB_KT_CAP; }

static int enqueue_iteration(bbpcg_solver *s, int it, bool parts, const real *rhs)
{
  const bool kt = s->kernel_timing && it <= kt_limit(s);
  s->kt_now = kt;
  if (kt && it == 1) CU(cudaEventRecord(s->kev[0], s->stream));
  int rc = launch_search(s, parts);
  if (rc) return rc;
  if (kt) CU(cudaEventRecord(s->kev[2 * it - 1], s->stream));
  if (it % 50 == 0) {                                     /* cuda_solver.cu:209-223 */
    rc = launch_refresh_x4(s);
    if (!rc) rc = launch_resid(s, parts, rhs);
  } else rc = launch_resid(s, parts, NULL);
  if (rc) return rc;
  if (kt) CU(cudaEventRecord(s->kev[2 * it], s->stream));
  s->kt_now = 0;
  return BBPCG_OK;
}

extern "C" int bbpcg_solve(bbpcg_solver *s, const bbpcg_solve_args *a, bbpcg_result *res)
{
  if (!s || !a || !a->u_star || !a->v_star || !a->w_star || !a->rhs_p || !a->phi) { bbpcg_set_error("bbpcg_solve: NULL argument"); return BBPCG_EINVAL; }
  if (!s->coeffs_set) { bbpcg_set_error("bbpcg_solve: call bbpcg_set_coefficients first (cuda_PP_init_jacobi_preconditioner)"); return BBPCG_EINVAL; }
  if (a->use_phase && (!s->has_phase || !a->phase)) { bbpcg_set_error("bbpcg_solve: use_phase needs phase arrays"); return BBPCG_EINVAL; }
  CU(cudaSetDevice(s->device));
  const dom_struct &d = s->dom;
  const grid_info &g = d.Gcc;
  const Layout &L = s->dev.L;
  const bool parts = a->use_phase != 0;
  const long long launches0 = s->launches;
  const long long nrows = (long long)L.jn * L.kn;
  const int nbs = clampi(nrows, 1, s->stream_blocks);

  CU(cudaEventRecord(s->ev[0], s->stream));
  /* ---- set-up: cuda_solver.cu:122-170 ---- */
  int rc = enqueue_rhs(s, a->u_star, a->v_star, a->w_star, a->rho_f, a->dt, a->rhs_p);
  if (rc) return rc;
  if (parts) {                                             /* :128-148 */
    if (a->part_bc) {
      CU(cudaStreamSynchronize(s->stream));
      a->part_bc();                                        /* reference code, default stream */
      CU(cudaDeviceSynchronize());
    } else if (a->phase_shell) {
      k_part_rhs_net<<<(unsigned)((g.s3b + 255) / 256), 256, 0, s->stream>>>(a->rhs_p, a->phase, a->phase_shell, g.s3b);
      s->launches++;
    }
    rc = enqueue_exchange(s, a->rhs_p);
    if (rc) return rc;
    if (!a->no_refine) {                                   /* :139-142, guarded by the rank-local nparts > 0 in the reference */
      k_coeffs_refine<256><<<nbs, 256, 0, s->stream>>>(g.in, g.jn, g.kn, g.s1b, g.s2b, a->rhs_p, a->phase, s->dev.idx2, s->dev.idy2, s->dev.idz2);
      s->launches++;
    }
    k_zero_ghosts<<<s->sm_count * 4, 256, 0, s->stream>>>(a->rhs_p, g.inb, g.jnb, g.knb);
    s->launches++;
  }
  CU(cudaMemsetAsync(s->dev.x, 0, sizeof(double) * (size_t)L.n, s->stream));
  CU(cudaMemsetAsync(s->dev.P[0], 0, sizeof(double) * (size_t)L.n, s->stream));
  const int max_q = a->fixed_iters > 0 ? a->fixed_iters : a->pp_max_iter + 1;
  k_init<256><<<nbs, 256, 0, s->stream>>>(s->dev, a->rhs_p, g.s1b, g.s2b, a->pp_residual * a->pp_residual, max_q, a->fixed_iters > 0 ? 1 : 0);
  s->launches++;
  CU(cudaEventRecord(s->ev[1], s->stream));

  /* ---- iteration loop: kernels are enqueued in batches; each batch ends with an async read
   * of Scal::done.  Two batches stay in flight so the GPU never idles while the host polls;
   * once done is set the remaining launches are no-ops. ---- */
  int it = 0, nbatch = 0;
  bool finished = false;
  volatile int *poll = s->h_poll;
  poll[0] = poll[16 / 4] = 0;
  while (!finished && it < max_q) {
    const int upto = (it + s->check_every < max_q) ? it + s->check_every : max_q;
    while (it < upto) {
      ++it;
      rc = enqueue_iteration(s, it, parts, a->rhs_p);
      if (rc) { cudaStreamSynchronize(s->stream); return rc; }     /* nothing of this solve stays enqueued behind an error */
    }
    const int slot = nbatch & 1;
    CU(cudaMemcpyAsync((void *)&s->h_poll[slot * 4], &s->dev.sc->done, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
    CU(cudaEventRecord(s->ev_poll[slot], s->stream));
    if (nbatch >= 1) {
      const int prev = (nbatch - 1) & 1;
      CU(cudaEventSynchronize(s->ev_poll[prev]));
      if (poll[prev * 4]) finished = true;
    }
    nbatch++;
  }
  CU(cudaEventRecord(s->ev[2], s->stream));
  k_finish<256><<<nbs, 256, 0, s->stream>>>(s->dev, a->phi, g.s1b, g.s2b);
  s->launches++;
  CU(cudaMemcpyAsync(s->h_scal, s->dev.sc, sizeof(Scal), cudaMemcpyDeviceToHost, s->stream));
  CU(cudaEventRecord(s->ev[3], s->stream));
  CU(cudaStreamSynchronize(s->stream));
  CU(cudaGetLastError());

  const Scal &sc = *s->h_scal;
  if (s->kernel_timing) {
    s->kt_search_ms = s->kt_resid_ms = s->kt_refresh_ms = 0.; s->kt_search_n = s->kt_resid_n = s->kt_refresh_n = 0;
    for (int i = 1; i <= sc.q && i <= kt_limit(s) && i <= it; i++) {
      float a = 0.f, b = 0.f;
      cudaEventElapsedTime(&a, s->kev[2 * i - 2], s->kev[2 * i - 1]);
      cudaEventElapsedTime(&b, s->kev[2 * i - 1], s->kev[2 * i]);
      s->kt_search_ms += a; s->kt_search_n++;
      if (i % 50 == 0) { s->kt_refresh_ms += b; s->kt_refresh_n++; } else { s->kt_resid_ms += b; s->kt_resid_n++; }
    }
  }
  if (res) {
    float ms;
    res->status = sc.status; res->niter = sc.q; res->resid = sc.resid; res->sp_rhs = sc.bb; res->sp_rq0 = sc.rz0;
    if (a->fixed_iters <= 0 && !sc.done) res->status = BBPCG_MAXITER;
    cudaEventElapsedTime(&ms, s->ev[0], s->ev[1]); res->ms_setup = ms;
    cudaEventElapsedTime(&ms, s->ev[1], s->ev[2]); res->ms_iter = ms;
    cudaEventElapsedTime(&ms, s->ev[0], s->ev[3]); res->ms_total = ms;
    res->launches = s->launches - launches0;
  }
  if (sc.status == BBPCG_COMM_TIMEOUT) s->h_poll[BB_POLL_COMM] = 1;
  if (comm_check(s, "bbpcg_solve")) return BBPCG_ECOMM;
  return BBPCG_OK;
}

extern "C" int bbpcg_solve_host(bbpcg_solver *s, const real *u_h, const real *v_h, const real *w_h, real *phi_h,
                                real rho_f, real dt, real pp_residual, int pp_max_iter, int fixed_iters, bbpcg_result *res)
{
  if (!s || !u_h || !v_h || !w_h || !phi_h) { bbpcg_set_error("bbpcg_solve_host: NULL argument"); return BBPCG_EINVAL; }
  CU(cudaSetDevice(s->device));
  const dom_struct &d = s->dom;
  if (!s->hb_u) {
    CU(cudaMalloc(&s->hb_u, sizeof(real) * (size_t)d.Gfx.s3b)); CU(cudaMalloc(&s->hb_v, sizeof(real) * (size_t)d.Gfy.s3b));
    CU(cudaMalloc(&s->hb_w, sizeof(real) * (size_t)d.Gfz.s3b));
    CU(cudaMalloc(&s->hb_rhs, sizeof(real) * (size_t)d.Gcc.s3b)); CU(cudaMalloc(&s->hb_phi, sizeof(real) * (size_t)d.Gcc.s3b));
    CU(cudaMemset(s->hb_phi, 0, sizeof(real) * (size_t)d.Gcc.s3b));
  }
  CU(cudaMemcpyAsync(s->hb_u, u_h, sizeof(real) * (size_t)d.Gfx.s3b, cudaMemcpyHostToDevice, s->stream));
  CU(cudaMemcpyAsync(s->hb_v, v_h, sizeof(real) * (size_t)d.Gfy.s3b, cudaMemcpyHostToDevice, s->stream));
  CU(cudaMemcpyAsync(s->hb_w, w_h, sizeof(real) * (size_t)d.Gfz.s3b, cudaMemcpyHostToDevice, s->stream));
  bbpcg_solve_args a;
  memset(&a, 0, sizeof(a));
  a.u_star = s->hb_u; a.v_star = s->hb_v; a.w_star = s->hb_w; a.rhs_p = s->hb_rhs; a.phi = s->hb_phi;
  a.rho_f = rho_f; a.dt = dt; a.pp_residual = pp_residual; a.pp_max_iter = pp_max_iter; a.fixed_iters = fixed_iters;
  int rc = bbpcg_solve(s, &a, res);
  if (rc) return rc;
  CU(cudaMemcpyAsync(phi_h, s->hb_phi, sizeof(real) * (size_t)d.Gcc.s3b, cudaMemcpyDeviceToHost, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  return BBPCG_OK;
}

extern "C" int bbpcg_history(bbpcg_solver *s, double *out, int cap)
{
  if (!s || !out || cap < 1) return 0;
  if (cudaSetDevice(s->device) != cudaSuccess) return 0;
  int n = s->h_scal->q + 1;
  if (n > cap) n = cap;
  if (n > BB_HIST_CAP) n = BB_HIST_CAP;
  if (cudaMemcpy(out, s->dev.history, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost) != cudaSuccess) return 0;
  return n;
}

extern "C" int bbpcg_set_option(bbpcg_solver *s, const char *key, long long value)
{
  if (!s || !key) return BBPCG_EINVAL;
  if (!strcmp(key, "ty")) { if (value < 0 || value > BB_TYMAX) { bbpcg_set_error("ty must be 0 (automatic) .. %d", BB_TYMAX); return BBPCG_EINVAL; } s->opt_ty = (int)value; s->plan_ok = 0; }
  else if (!strcmp(key, "kc")) { s->opt_kc = value > 0 ? (int)value : 0; s->plan_ok = 0; }
  else if (!strcmp(key, "pdl")) s->pdl = clampi(value, 0, 2);
  else if (!strcmp(key, "rhs_tiled")) s->rhs_tiled = value != 0;
  else if (!strcmp(key, "tma_warp")) s->tma_warp = value != 0;
  else if (!strcmp(key, "epilogue_tiled")) s->epilogue_tiled = value != 0;
  else if (!strcmp(key, "epi_chunk")) s->epi_chunk = clampi(value, 0, 4096);
  else if (!strcmp(key, "guided")) { s->guided = value != 0; s->plan_ok = 0; }
  else if (!strcmp(key, "guided_pct")) { s->guided_pct = clampi(value, 10, 400); s->plan_ok = 0; }
  else if (!strcmp(key, "chunk_min")) { s->chunk_min = clampi(value, 2, 1024); s->plan_ok = 0; }
  else if (!strcmp(key, "stream_blocks")) s->stream_blocks = clampi(value, 1, BB_MAXBLOCKS);
  else if (!strcmp(key, "check_every")) s->check_every = clampi(value, 1, 1000);
  else if (!strcmp(key, "comm_timeout_ms")) s->dev.comm.timeout_cycles = value > 0 ? value * 2000000ll : -1;   /* ~2 GHz; <= 0: wait for ever, like MPI */
  else if (!strcmp(key, "comm_reset")) {                    /* tests only: clear a recorded time-out (all ranks must do it together) */
    CU(cudaSetDevice(s->device));
    CU(cudaStreamSynchronize(s->stream));
    CU(cudaMemset(&s->dev.sc->comm_timeout, 0, sizeof(int)));
    s->h_poll[BB_POLL_COMM] = 0;
  }
  else if (!strcmp(key, "kernel_timing")) {
    s->kernel_timing = (int)clampi(value, 0, BB_KT_CAP);
    if (s->kernel_timing && !s->kev) {
      CU(cudaSetDevice(s->device));
      s->kev = (cudaEvent_t *)calloc(2 * BB_KT_CAP + 1, sizeof(cudaEvent_t));
      for (int i = 0; i <= 2 * BB_KT_CAP; i++) CU(cudaEventCreate(&s->kev[i]));
    }
  }
  else { bbpcg_set_error("unknown option %s", key); return BBPCG_EINVAL; }
  return BBPCG_OK;
}

#ifdef BB_TRACE
/* debug build only: the stamps of the last BB_TRACE_LAUNCHES launches of the iteration kernels, raw u64 */
extern "C" int bbpcg_trace_dump(bbpcg_solver *s, const char *path)
{
  if (!s || !path || !s->dev.trace) return BBPCG_EINVAL;
  CU(cudaSetDevice(s->device));
  const size_t n = (size_t)BB_TRACE_LAUNCHES * BB_TRACE_CTAS * BB_TRACE_EV;
  std::vector<unsigned long long> h(n);
  CU(cudaMemcpy(h.data(), s->dev.trace, n * 8, cudaMemcpyDeviceToHost));
  FILE *f = fopen(path, "wb");
  if (!f) return BBPCG_EIO;
  const long long hdr[4] = { BB_TRACE_LAUNCHES, BB_TRACE_CTAS, BB_TRACE_EV, s->launches };
  fwrite(hdr, 8, 4, f); fwrite(h.data(), 8, n, f); fclose(f);
  return BBPCG_OK;
}
#endif

extern "C" long long bbpcg_get_info(bbpcg_solver *s, const char *key)
{
  if (!s || !key) return -1;
  if (!strcmp(key, "arena_bytes")) return (long long)s->amap.total;
  if (!strcmp(key, "launches")) return s->launches;
  if (!strcmp(key, "pitch")) return s->dev.L.px;
  if (!strcmp(key, "sm_count")) return s->sm_count;
  if (!strcmp(key, "nranks")) return s->nranks;
  if (!strcmp(key, "tile_tx")) return 128;
  if (!strcmp(key, "tile_ty") || !strcmp(key, "search_ty")) return make_plan(s) ? -1 : s->plan_ty;
  /* per-kernel device time of the last solve (kernel_timing = 1), nanoseconds / launch counts */
  if (!strcmp(key, "kt_search_ns")) return (long long)(s->kt_search_ms * 1e6);
  if (!strcmp(key, "kt_resid_ns")) return (long long)(s->kt_resid_ms * 1e6);
  if (!strcmp(key, "kt_refresh_ns")) return (long long)(s->kt_refresh_ms * 1e6);
  if (!strcmp(key, "kt_search_n")) return s->kt_search_n;
  if (!strcmp(key, "kt_resid_n")) return s->kt_resid_n;
  if (!strcmp(key, "kt_refresh_n")) return s->kt_refresh_n;
  if (!strcmp(key, "search_grid")) { if (make_plan(s)) return -1; const long long ni = (long long)s->plan_nbx * s->plan_nby * s->plan_nbz; return ni < 2 * s->sm_count ? ni : 2 * s->sm_count; }
  if (!strcmp(key, "search_items")) return make_plan(s) ? -1 : (long long)s->plan_nbx * s->plan_nby * s->plan_nbz;
  if (!strcmp(key, "search_kc")) return make_plan(s) ? -1 : s->plan_kc;
  if (!strcmp(key, "search_nbz")) return make_plan(s) ? -1 : s->plan_nbz;
  if (!strcmp(key, "pdl")) return pdl_active(s);
  if (!strcmp(key, "comm_timeout")) return s->h_poll ? s->h_poll[BB_POLL_COMM] : 0;
  return -1;
}
