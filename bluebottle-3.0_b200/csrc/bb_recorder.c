/* bb_recorder.c -- writer of the solver's record files (no CUDA), SURVEY.md 8f rank 4.
 *
 * Byte-compatible with what the reference appends per pressure solve:
 *   recorder_PP_init / recorder_PP              src/recorder.c:157-221   ->  record/solver_expd.rec
 *   recorder_PP_init_timed / recorder_PP_timed  src/recorder.c:223-336   ->  record/solver_expd_timed.rec
 * The reference averages the elapsed times over the ranks with MPI_Allreduce before rank 0 writes
 * (recorder.c:193-194, 278-296); here the caller passes the averaged values and only rank 0 calls.
 * A line is "\n" + the fields, i.e. the file never ends in a newline -- kept, since tools that tail
 * the file rely on it.
 */
#include <errno.h>
#include <stdio.h>
#include <string.h>
#include <sys/stat.h>
#include <sys/types.h>

#include "../../include/bbpcg.h"

extern void bbpcg_set_error(const char *fmt, ...);

#define REC_PATH 1024

static int rec_path(char *out, const char *root_dir, const char *name, int make_dir)
{
  if (!root_dir || !name) { bbpcg_set_error("bb_recorder: NULL root_dir / name"); return BBPCG_EINVAL; }
  char dir[REC_PATH];
  if (snprintf(dir, sizeof(dir), "%s/record", root_dir) >= (int)sizeof(dir) ||
      snprintf(out, REC_PATH, "%s/record/%s", root_dir, name) >= REC_PATH) {
    bbpcg_set_error("bb_recorder: path too long"); return BBPCG_EINVAL;
  }
  if (make_dir) {                                  /* recorder.c:161-168: create record/ with mode 0700 if missing */
    struct stat st;
    if (stat(dir, &st) == -1 && mkdir(dir, 0700) == -1 && errno != EEXIST) {
      bbpcg_set_error("bb_recorder: cannot create %s: %s", dir, strerror(errno)); return BBPCG_EIO;
    }
  }
  return BBPCG_OK;
}

static const char *const seg_titles[8] = {         /* recorder.c:250-257 */
  "spmv time (s)", "ip1 time (s)", "ar1 time (s)", "up1 time (s)", "ip2 time (s)", "ar2 time (s)", "up2 time (s)", "mpi time (s)"
};

static int rec_init(const char *root_dir, const char *name, int timed)
{
  char path[REC_PATH];
  int rc = rec_path(path, root_dir, name, 1);
  if (rc) return rc;
  FILE *rec = fopen(path, "w");
  if (!rec) { bbpcg_set_error("Could not open file %s", name); return BBPCG_EIO; }
  fprintf(rec, "%-12s", "stepnum");                /* recorder.c:179-184 / 245-249 */
  fprintf(rec, "%-15s", "ttime");
  fprintf(rec, "%-15s", "dt");
  fprintf(rec, "%-8s", "niter");
  fprintf(rec, "%-15s", "resid");
  if (!timed) fprintf(rec, "%-15s", "time (s)");
  else {
    fprintf(rec, "%-16s", "Total time (s)");
    for (int i = 0; i < 8; i++) fprintf(rec, "%-16s", seg_titles[i]);
  }
  fclose(rec);
  return BBPCG_OK;
}

static int rec_line(const char *root_dir, const char *name, int stepnum, real ttime, real dt, int niter, real resid,
                    real etime, const real *seg)
{
  char path[REC_PATH];
  int rc = rec_path(path, root_dir, name, 0);
  if (rc) return rc;
  FILE *rec = fopen(path, "r+");
  if (!rec) {                                      /* recorder.c:201-204: first line of a run creates the file */
    rc = rec_init(root_dir, name, seg != NULL);
    if (rc) return rc;
    rec = fopen(path, "r+");
    if (!rec) { bbpcg_set_error("Could not open file %s", name); return BBPCG_EIO; }
  }
  fseek(rec, 0, SEEK_END);
  fprintf(rec, "\n");
  fprintf(rec, "%-12d", stepnum);                  /* recorder.c:209-216 / 312-326 */
  fprintf(rec, "%-15e", ttime);
  fprintf(rec, "%-15e", dt);
  fprintf(rec, "%-8d", niter);
  fprintf(rec, "%-15e", resid);
  if (!seg) fprintf(rec, "%-15e", etime);
  else {
    fprintf(rec, "%-16e", etime);
    for (int i = 0; i < 8; i++) fprintf(rec, "%-16e", seg[i]);
  }
  fclose(rec);
  return BBPCG_OK;
}

int bb_recorder_PP_init(const char *root_dir, const char *name) { return rec_init(root_dir, name, 0); }
int bb_recorder_PP_init_timed(const char *root_dir, const char *name) { return rec_init(root_dir, name, 1); }

int bb_recorder_PP(const char *root_dir, const char *name, int stepnum, real ttime, real dt, int niter, real resid, real etime)
{
  return rec_line(root_dir, name, stepnum, ttime, dt, niter, resid, etime, NULL);
}

int bb_recorder_PP_timed(const char *root_dir, const char *name, int stepnum, real ttime, real dt, int niter, real resid,
                         real etime, const real seg[8])
{
  if (!seg) { bbpcg_set_error("bb_recorder_PP_timed: NULL segment array"); return BBPCG_EINVAL; }
  return rec_line(root_dir, name, stepnum, ttime, dt, niter, resid, etime, seg);
}
