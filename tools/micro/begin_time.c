// begin/end latency of the fused call from a plain C process (no Python, no torch): tools/micro, not product
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "texfusion.h"
static double now(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec * 1e6 + t.tv_nsec * 1e-3; }
static int cmp(const void* a, const void* b) { double x = *(const double*)a, y = *(const double*)b; return x < y ? -1 : x > y; }
int main(void) {
  tf_config cfg; memset(&cfg, 0, sizeof cfg);
  cfg.chunk_dim = 8; cfg.voxel_res = 0.005f; cfg.use_color = 1;
  cfg.trunc.quad = 0.0019f; cfg.trunc.lin = 0.00152f; cfg.trunc.cst = 0.001504f; cfg.trunc.scale = 6.0f; cfg.trunc.weight = 1.0f;
  cfg.n_ranks = 1; cfg.width = 640; cfg.height = 480; cfg.max_frames = 8;
  tf_map* m = NULL;
  if (tf_create(&m, &cfg) != 0) { printf("create failed: %s\n", tf_last_error(NULL)); return 1; }
  float* depth = (float*)tf_host_alloc(640 * 480 * 4);
  for (int i = 0; i < 640 * 480; i++) depth[i] = 1.5f + 0.3f * (float)((i % 640) + (i / 640)) / 1120.0f;
  tf_camera cam = {525.f, 525.f, 319.5f, 239.5f, 640, 480, 0.01f, 5.0f};
  tf_pose pose; memset(&pose, 0, sizeof pose); pose.m[0] = pose.m[5] = pose.m[10] = pose.m[15] = 1.0f;
  tf_frame_stats st;
  double tb[200], te[200]; int n = 0;
  for (int i = 0; i < 220; i++) {
    pose.m[12] = 0.001f * i;  // slide sideways
    tf_upload_frame(m, i, depth, NULL, NULL);
    tf_wait_upload(m, i);
    double a = now();
    int rc = tf_integrate_frame_begin(m, i, 0, &pose, &cam, NULL, NULL, NULL, NULL, 0);
    double b = now();
    rc |= tf_integrate_frame_end(m, &st);
    double c = now();
    if (rc) { printf("error: %s\n", tf_last_error(m)); return 1; }
    if (i >= 20) { tb[n] = b - a; te[n] = c - b; n++; }
  }
  qsort(tb, n, sizeof(double), cmp); qsort(te, n, sizeof(double), cmp);
  printf("plain C: begin median %.2f us, end median %.2f us, chunks %lld\n", tb[n / 2], te[n / 2], (long long)st.n_chunks);
  tf_destroy(m);
  return 0;
}
