/* Minimal C client of the drop-in boundary (include/pps_b200.h): what a cgo / JNI / ctypes binding would call.
 *
 *   gcc -O2 -Iinclude examples/evaluate_host.c -Lpps_b200/_C -lpps_b200 -Wl,-rpath,$PWD/pps_b200/_C -lm -o /tmp/evaluate_host
 *   /tmp/evaluate_host            (needs a B200: there is no CPU path)
 *
 * Builds a small synthetic re-ID set on the host (identity centres + noise, L2-normalised rows), then one call does
 * what reid_dataset_evaluator.py:104-122 does: distance -> junk mask -> ranking -> mAP / CMC, every host<->device
 * copy inside the call. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "pps_b200.h"

static float gauss(void) {            /* Box-Muller on rand(): good enough for a demo */
  const double u = (rand() + 1.0) / ((double)RAND_MAX + 2.0), v = (rand() + 1.0) / ((double)RAND_MAX + 2.0);
  return (float)(sqrt(-2.0 * log(u)) * cos(6.283185307179586 * v));
}

static void make_rows(float* x, const int64_t* ids, long long n, int dim, const float* centers) {
  for (long long i = 0; i < n; ++i) {
    double nrm = 0.0;
    for (int d = 0; d < dim; ++d) {
      const float c = ids[i] > 0 ? centers[(ids[i] - 1) * dim + d] : 0.f;
      x[i * dim + d] = c + 3.0f * gauss();
      nrm += (double)x[i * dim + d] * x[i * dim + d];
    }
    const float inv = (float)(1.0 / sqrt(nrm));
    for (int d = 0; d < dim; ++d) x[i * dim + d] *= inv;
  }
}

int main(void) {
  const long long nq = 200, ng = 3000;
  const int dim = 256, n_ids = 50, n_cams = 4, topk = 10;
  srand(7);
  float* centers = malloc(sizeof(float) * n_ids * dim);
  for (int i = 0; i < n_ids * dim; ++i) centers[i] = gauss();
  int64_t *qid = malloc(8 * nq), *qcam = malloc(8 * nq), *gid = malloc(8 * ng), *gcam = malloc(8 * ng);
  for (long long i = 0; i < nq; ++i) { qid[i] = 1 + rand() % n_ids; qcam[i] = rand() % n_cams; }
  for (long long i = 0; i < ng; ++i) { gid[i] = rand() % (n_ids + 1); gcam[i] = rand() % n_cams; }   /* id 0: distractors */
  float *q = malloc(sizeof(float) * nq * dim), *g = malloc(sizeof(float) * ng * dim);
  make_rows(q, qid, nq, dim, centers);
  make_rows(g, gid, ng, dim, centers);

  double map = 0.0, cmc[10];
  double* ap = malloc(8 * nq);
  uint8_t* valid = malloc(nq);
  int32_t* first = malloc(4 * nq);
  int32_t* top_idx = malloc(4 * nq * topk);
  float* top_d = malloc(4 * nq * topk);
  const int rc = pps_evaluate_host(q, nq, g, ng, dim, qid, qcam, gid, gcam, PPS_PREC_F16X3, /*cmc_topk=*/10, topk,
                                   /*device=*/0, &map, cmc, ap, valid, first, top_idx, top_d);
  if (rc != PPS_OK) {
    fprintf(stderr, "pps_evaluate_host: %s (%s)\n", pps_strerror(rc), pps_last_cuda_error());
    return 1;
  }
  printf("ABI %d  mAP %.4f  CMC-1 %.4f  CMC-5 %.4f  CMC-10 %.4f\n", pps_abi_version(), map, cmc[0], cmc[4], cmc[9]);
  printf("query 0: id %lld, nearest gallery rows:", (long long)qid[0]);
  for (int k = 0; k < 5; ++k) printf(" %d (id %lld, d %.3f)", top_idx[k], (long long)gid[top_idx[k]], top_d[k]);
  printf("\n%llu kernel launches\n", pps_kernel_launch_count());
  return 0;
}
