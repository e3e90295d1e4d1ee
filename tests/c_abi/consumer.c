/* A plain C99 consumer of include/i2c_b200.h: no Python, no torch, no C++ -- the drop-in boundary as a reference
 * maintainer binding it from C would use it.  Builds the batched pendulum swing-up problem of BASELINE configs[2] (small),
 * runs n EM iterations and prints checksums of the controllers and the alpha schedule; tests/test_abi.py compares them
 * with the numbers of the Python host path (same library, same inputs).
 *   gcc -std=c99 -Iinclude tests/c_abi/consumer.c -L<libdir> -li2c_b200 -Wl,-rpath,<libdir> -lm -o consumer
 *   ./consumer B T n_iter            exit code 0 = ran; 3 = no CUDA device (the library refuses: there is no CPU fallback) */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "i2c_b200.h"

static double lcg(unsigned long long* s) { /* deterministic inputs shared with the Python side of the test */
  *s = *s * 6364136223846793005ULL + 1442695040888963407ULL;
  return (double)((*s >> 11) & ((1ULL << 53) - 1)) / (double)(1ULL << 53) - 0.5;
}

int main(int argc, char** argv) {
  const int B = argc > 1 ? atoi(argv[1]) : 64, T = argc > 2 ? atoi(argv[2]) : 30, n_iter = argc > 3 ? atoi(argv[3]) : 3;
  int32_t dx, du, dz, dzt, np, dy;
  if (i2c_env_dims(I2C_ENV_PENDULUM, &dx, &du, &dz, &dzt, &np, &dy) != 0 || dx != 2 || du != 1 || dz != 4 || dzt != 3) {
    fprintf(stderr, "env dims: %s\n", i2c_last_error());
    return 2;
  }
  i2c_config_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.abi_version = I2C_ABI_VERSION;
  cfg.env = I2C_ENV_PENDULUM;
  cfg.inference = I2C_INF_CUBATURE;
  cfg.n_problems = B;
  cfg.horizon = T;
  cfg.max_iters = 16;
  cfg.quad_alpha = 1.0;
  size_t ws = 0;
  if (i2c_workspace_bytes(&cfg, &ws) != 0 || ws == 0) return 2;
  i2c_handle_t h = NULL;
  if (i2c_create(&cfg, NULL, 0, NULL, &h) != 0) {
    printf("no_device %s\n", i2c_last_error());
    return 3;
  }
  unsigned long long seed = 12345;
  double* x0 = malloc(sizeof(double) * B * 2);
  double* s0 = calloc((size_t)B * 4, sizeof(double));
  double* mu_u = malloc(sizeof(double) * B * T);
  double* alpha0 = malloc(sizeof(double) * B);
  double* z = calloc((size_t)T * 4, sizeof(double));
  for (int b = 0; b < B; ++b) {
    x0[2 * b] = 3.14159265358979323846 + 0.6 * lcg(&seed);
    x0[2 * b + 1] = 1.0 * lcg(&seed);
    s0[4 * b] = s0[4 * b + 3] = 1e-5;
    alpha0[b] = 100.0;
  }
  for (int i = 0; i < B * T; ++i) mu_u[i] = 0.02 * lcg(&seed);
  /* observe(): z = [sin th, cos th, th_dot, u]; goal = upright: [0, 1, 0, 0] (env_def.py:273-291) */
  for (int t = 0; t < T; ++t) z[4 * t + 1] = 1.0;
  const double z_graph[4] = {0.0, 1.0, 0.0, 0.0}, z_term[3] = {0.0, 1.0, 0.0};
  const double sig_eta[4] = {1e-5, 0.0, 0.0, 1e-5}, sig_u[1] = {2.0};
  const double QR[16] = {1, 0, 0, 0, 0, 100, 0, 0, 0, 0, 1, 0, 0, 0, 0, 2};
  const double Qf[9] = {1, 0, 0, 0, 100, 0, 0, 0, 1};
  if (i2c_set_problem(h, x0, s0, sig_eta, mu_u, sig_u, QR, Qf, z, z_graph, z_term, alpha0, 0.0, NULL, NULL, 1.0, NULL) != 0) {
    fprintf(stderr, "set_problem: %s\n", i2c_last_error());
    return 2;
  }
  if (i2c_run(h, n_iter, I2C_PH_FORWARD | I2C_PH_BACKWARD | I2C_PH_MSTEP | I2C_PH_UPDATE_PRIORS) != 0) {
    fprintf(stderr, "run: %s\n", i2c_last_error());
    return 2;
  }
  double* K = malloc(sizeof(double) * B * T * 2);
  double* k = malloc(sizeof(double) * B * T);
  double* sK = malloc(sizeof(double) * B * T);
  double* alpha = malloc(sizeof(double) * (size_t)n_iter * B);
  int32_t* st = malloc(sizeof(int32_t) * B);
  if (i2c_get_policy(h, K, k, sK) != 0 || i2c_get_metric(h, I2C_M_ALPHA, alpha, n_iter) != 0 || i2c_get_status(h, st, NULL) != 0) {
    fprintf(stderr, "getters: %s\n", i2c_last_error());
    return 2;
  }
  double cK = 0, ck = 0, cs = 0, ca = 0;
  int bad = 0;
  for (int i = 0; i < B * T * 2; ++i) cK += K[i] * (1.0 + (i % 7));
  for (int i = 0; i < B * T; ++i) ck += k[i] * (1.0 + (i % 5)), cs += sK[i];
  for (int i = 0; i < n_iter * B; ++i) ca += alpha[i];
  for (int b = 0; b < B; ++b) bad += st[b] != I2C_OK;
  printf("ok B=%d T=%d n_iter=%d failed=%d\nK %.17g\nk %.17g\nsigK %.17g\nalpha %.17g\n", B, T, n_iter, bad, cK, ck, cs, ca);
  i2c_destroy(h);
  free(x0), free(s0), free(mu_u), free(alpha0), free(z), free(K), free(k), free(sK), free(alpha), free(st);
  return 0;
}
