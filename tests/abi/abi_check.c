/* Plain-C client of include/piql_b200.h: what a non-Python host (the Julia ccall shim, or any C program) sees.
 * Without a CUDA device it checks the no-fallback contract (piql_create fails with PIQL_ERR_NO_DEVICE); with one it loads a
 * 6 x 7 gridworld, solves nu and one logz_updateV and checks the fixed-point algebra.  Exit code 0 = all checks passed. */
#include <math.h>
#include <stdio.h>
#include <string.h>

#include "piql_b200.h"

#define CHECK(cond, msg) do { if (!(cond)) { fprintf(stderr, "FAIL: %s\n", msg); return 1; } } while (0)

int main(void) {
  CHECK(piql_version() == PIQL_VERSION, "piql_version");
  int32_t ndev = -1;
  int32_t rc = piql_device_count(&ndev);
  piql_ctx* ctx = NULL;
  if (rc != PIQL_OK || ndev == 0) {
    CHECK(piql_create(&ctx, 0) == PIQL_ERR_NO_DEVICE && ctx == NULL, "no device: piql_create must fail, there is no CPU fallback");
    printf("no device: PIQL_ERR_NO_DEVICE as specified\n");
    return 0;
  }
  CHECK(piql_create(&ctx, 0) == PIQL_OK && ctx != NULL, "piql_create");
  enum { H = 6, W = 7 };
  uint8_t walls[H * W];
  for (int j = 0; j < W; j++)
    for (int i = 0; i < H; i++) walls[i + H * j] = (i == 0 || j == 0 || i == H - 1 || j == W - 1);
  walls[2 + H * 3] = 1; /* one inner barrier */
  const piql_grid_params p = {0.05, 0.02, 0.2, 0.95};
  CHECK(piql_grid_load(ctx, walls, H, W, 4, 5, &p) == PIQL_OK, piql_last_error(ctx));
  CHECK(piql_grid_load(ctx, walls, H, W, 0, 0, &p) == PIQL_ERR_ARG, "a goal inside a wall must be rejected");
  CHECK(piql_grid_load(ctx, walls, H, W, 4, 5, &p) == PIQL_OK, piql_last_error(ctx));
  int64_t nfree = 0;
  CHECK(piql_grid_free_count(ctx, &nfree) == PIQL_OK && nfree == 4 * 5 - 1, "free cells");
  const piql_solve_opts o = {1e-12, 100000, PIQL_F64, 0};
  double nu[H * W], v1[H * W], logz[H * W];
  piql_solve_stats st;
  CHECK(piql_solve_nu(ctx, 1.0, &o, nu, &st) == PIQL_OK && st.converged, piql_last_error(ctx));
  CHECK(nu[4 + H * 5] == 0.0 && isnan(nu[0]) && isnan(nu[2 + H * 3]), "nu(goal) = 0, NaN at walls");
  double in[H * W];
  for (int s = 0; s < H * W; s++) in[s] = isnan(nu[s]) ? 0.0 : nu[s];
  CHECK(piql_update_v(ctx, in, 0.5, &o, v1, logz, &st) == PIQL_OK, piql_last_error(ctx));
  double worst = 0.0;
  for (int s = 0; s < H * W; s++)
    if (!isnan(nu[s])) { worst = fmax(worst, fabs(logz[s])); worst = fmax(worst, fabs(v1[s] - nu[s])); }
  CHECK(worst < 1e-9, "logz_updateV leaves nu unchanged (log Z = 0 at V = nu)");
  CHECK(piql_solve_nu(ctx, 1.0, NULL, nu, &st) == PIQL_ERR_ARG && strlen(piql_last_error(ctx)) > 0, "null opts -> PIQL_ERR_ARG + message");
  CHECK(piql_destroy(ctx) == PIQL_OK, "piql_destroy");
  printf("device path: nu solved in %lld sweeps, fixed-point check %.1e\n", (long long)st.sweeps, worst);
  return 0;
}
