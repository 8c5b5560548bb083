/*
 * piql_oracle.c -- CPU restatement of the PIQL.jl hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the checker, never the product: only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.  Nothing under
 * piql.jl_b200/ links, imports or calls it.
 *
 * What it restates (reference = chelate/PIQL.jl, paths relative to /root/reference):
 *   KL-control algebra          src/PIQL.jl:55-130
 *   soft value iteration        src/value_iteration.jl:62-169 (+ Growable :3-18)
 *   gridworld problem           src/Problems/GridWorld.jl:20-121
 *   path sampling               src/path_functions.jl:12-50 (a broken sketch in the
 *                               reference; the Monte-Carlo estimator is specified from
 *                               z_recursion, src/PIQL.jl:109-115 -- see DESIGN.md)
 *
 * The reference is Julia and cannot run in this image (no julia binary), so there is no
 * oracle/_ref.  PARITY PIN: the value-iteration half is pinned against the Float64
 * results the reference itself stored in data/ (decoded into tests/golden by
 * tools/make_golden.py; tests/test_oracle_golden.py).  The sampling half follows
 * third-party code that is not under /root/reference -- StatsBase.jl `sample(a, Weights)`
 * (version unpinned: Project.toml has no [compat] entry, Manifest.toml is git-ignored) --
 * and the reference holds no trajectory fixture, so trajectory parity is "unpinned":
 * the injected-uniform contract below is defined by this build, and the estimator is
 * pinned only through its expectation (the exact generate_logz fixed point).
 *
 * Third-party arithmetic restated here: LogExpFunctions.jl logsumexp / logaddexp
 * (max-shifted log-sum-exp; unpinned version, results pinned to ~1e-12 by the goldens).
 *
 * Conventions: wall array is H x W *including* the wall border, column-major like
 * Julia (cell index s = i + H*j, 0-based); actions 0..3 with
 * step(0)=(+1,0) step(1)=(-1,0) step(2)=(0,+1) step(3)=(0,-1) on (i,j)
 * (src/Problems/GridWorld.jl:24-26).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define NA 4

typedef struct {
  int64_t H, W;
  const uint8_t *walls;
  int64_t goal;          /* linear index */
  double randomness, step_cost, reward, gamma;
  double reward_scale;   /* modify(ctrl, reward_function = scale * r)  test/runtests.jl:7 */
} orc_grid;

/* A "callable" value function, like Growable or a lambda in the reference. */
typedef double (*orc_fn)(void *env, int64_t s);

/* ------------------------------------------------------------------------------------
 * LogExpFunctions
 * ---------------------------------------------------------------------------------- */
static double lse4(const double *x) {
  double m = x[0];
  for (int a = 1; a < NA; a++) if (x[a] > m) m = x[a];
  if (isinf(m)) return m;
  double s = 0.0;
  for (int a = 0; a < NA; a++) s += exp(x[a] - m);
  return m + log(s);
}
static double logaddexp_(double x, double y) {
  if (isinf(x) && x < 0) return y;
  if (isinf(y) && y < 0) return x;
  double m = x > y ? x : y;
  return m + log1p(exp(-fabs(x - y)));
}

/* ------------------------------------------------------------------------------------
 * GridWorld.jl
 * ---------------------------------------------------------------------------------- */
static inline int64_t step_off(const orc_grid *g, int a) {
  /* step(a): (1-2*mod(a,2)) along dimension div(a,2)   GridWorld.jl:24-26 */
  int64_t sgn = 1 - 2 * (a & 1);
  return (a >> 1) == 0 ? sgn : sgn * g->H;
}
static inline int64_t take_step(const orc_grid *g, int64_t s, int a) {
  /* ifelse(walls[x1], x0, x1)   GridWorld.jl:29-32 */
  int64_t s1 = s + step_off(g, a);
  return g->walls[s1] ? s : s1;
}
static inline double slip_p(const orc_grid *g, int a, int aa) {
  /* (a == aa)*(1-randomness) + randomness/len   GridWorld.jl:101 */
  return (a == aa ? 1.0 : 0.0) * (1.0 - g->randomness) + g->randomness / NA;
}
static double pa(const orc_grid *g, int64_t s, int a, orc_fn f, void *env) {
  /* propagator_average   GridWorld.jl:98-104 */
  double out = 0.0;
  for (int aa = 0; aa < NA; aa++) out += slip_p(g, a, aa) * f(env, take_step(g, s, aa));
  return out;
}
static double log_pa(const orc_grid *g, int64_t s, int a, orc_fn f, void *env) {
  /* logpropagatorexp   GridWorld.jl:105-109 */
  double x[NA];
  for (int aa = 0; aa < NA; aa++) x[aa] = log(slip_p(g, a, aa)) + f(env, take_step(g, s, aa));
  return lse4(x);
}
static double reward_fn(const orc_grid *g, int64_t s1) {
  /* ifelse(s1 == goal, reward, -step_cost), times the modify() scale   GridWorld.jl:116 */
  return g->reward_scale * (s1 == g->goal ? g->reward : -g->step_cost);
}
static double reward_env(void *env, int64_t s1) { return reward_fn((const orc_grid *)env, s1); }

/* ------------------------------------------------------------------------------------
 * PIQL.jl algebra
 * ---------------------------------------------------------------------------------- */
static double average_reward(const orc_grid *g, int64_t s, int a) { /* PIQL.jl:55-57 */
  return pa(g, s, a, reward_env, (void *)g);
}
static double controlV(const orc_grid *g, orc_fn V, void *env, int64_t s, int a) {
  /* PIQL.jl:63-66 */
  return g->gamma * pa(g, s, a, V, env) - V(env, s) + average_reward(g, s, a);
}
static const double LOG_PRIOR = -1.3862943611198906; /* log(1.0/4) : action_prior  GridWorld.jl:112 */

/* Q is either V-derived ((s,a)->controlV(ctrl,V,s,a), value_iteration.jl:32,97,157) or a table */
typedef struct {
  const orc_grid *g;
  orc_fn V; void *Venv;
  const double *Qtab; /* NULL -> V-derived; else Q[s*4+a] */
} orc_q;
static double Qeval(const orc_q *q, int64_t s, int a) {
  return q->Qtab ? q->Qtab[s * NA + a] : controlV(q->g, q->V, q->Venv, s, a);
}
static double Qnormalization(const orc_q *q, int64_t s) { /* PIQL.jl:72-75 */
  double x[NA];
  for (int a = 0; a < NA; a++) x[a] = LOG_PRIOR + Qeval(q, s, a);
  return lse4(x);
}
static double controlQ(const orc_q *q, int64_t s, int a) { /* PIQL.jl:79-81 */
  return Qeval(q, s, a) - Qnormalization(q, s);
}
static double action_probability(const orc_q *q, int64_t s, int a) { /* PIQL.jl:83-85 */
  return 0.25 * exp(controlQ(q, s, a));
}
static double fitness(const orc_q *q, int64_t s, int a) { /* PIQL.jl:87-89 */
  return controlV(q->g, q->V, q->Venv, s, a) - controlQ(q, s, a);
}
static double truevalue_recursion(const orc_grid *g, int64_t s, orc_fn nu, void *env) {
  /* PIQL.jl:95-101 */
  double x[NA];
  for (int a = 0; a < NA; a++) x[a] = LOG_PRIOR + g->gamma * pa(g, s, a, nu, env) + average_reward(g, s, a);
  return lse4(x);
}
static double logz_recursion(const orc_q *q, int64_t s, orc_fn logz, void *zenv, double beta) {
  /* PIQL.jl:125-130 */
  const orc_grid *g = q->g;
  double x[NA];
  for (int a = 0; a < NA; a++)
    x[a] = LOG_PRIOR + beta * controlV(g, q->V, q->Venv, s, a) + (1 - beta) * controlQ(q, s, a) +
           logaddexp_(log(g->gamma) + log_pa(g, s, a, logz, zenv), log(1 - g->gamma));
  return lse4(x);
}
static double z_recursion(const orc_q *q, int64_t s, orc_fn Z, void *zenv, double beta) {
  /* PIQL.jl:109-115 */
  const orc_grid *g = q->g;
  double out = 0.0;
  for (int a = 0; a < NA; a++)
    out += 0.25 * exp(beta * controlV(g, q->V, q->Venv, s, a) + (1 - beta) * controlQ(q, s, a)) *
           (g->gamma * pa(g, s, a, Z, zenv) + 1 - g->gamma);
  return out;
}
static double optimality_gap_recursion(const orc_q *q, orc_fn nu, void *nuenv, int64_t s, orc_fn gap,
                                       void *gapenv) {
  /* value_iteration.jl:160-169 -- note: no gamma on the gap term */
  const orc_grid *g = q->g;
  double out = 0.0;
  for (int a = 0; a < NA; a++)
    out += action_probability(q, s, a) *
           (controlQ(q, s, a) - controlV(g, nu, nuenv, s, a) + pa(g, s, a, gap, gapenv));
  return out;
}

/* ------------------------------------------------------------------------------------
 * Value functions: dense array, and the reference's Growable (Dict + default)
 * ---------------------------------------------------------------------------------- */
static double dense_fn(void *env, int64_t s) { return ((const double *)env)[s]; }

typedef struct {        /* value_iteration.jl:3-18 */
  int64_t *key; double *val; int64_t cap, n;
  double dflt;
  int64_t *order;       /* insertion order, so keys() can be enumerated */
} growable;
static uint64_t hash64(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
  return x;
}
static void gr_init(growable *g, double dflt) {
  g->cap = 64; g->n = 0; g->dflt = dflt;
  g->key = (int64_t *)malloc(sizeof(int64_t) * g->cap);
  g->val = (double *)malloc(sizeof(double) * g->cap);
  g->order = (int64_t *)malloc(sizeof(int64_t) * g->cap);
  for (int64_t i = 0; i < g->cap; i++) g->key[i] = -1;
}
static void gr_free(growable *g) { free(g->key); free(g->val); free(g->order); }
static int64_t gr_slot(const growable *g, int64_t k) {
  int64_t i = (int64_t)(hash64((uint64_t)k) & (uint64_t)(g->cap - 1));
  while (g->key[i] != -1 && g->key[i] != k) i = (i + 1) & (g->cap - 1);
  return i;
}
static void gr_grow(growable *g) {
  growable o = *g;
  g->cap *= 2;
  g->key = (int64_t *)malloc(sizeof(int64_t) * g->cap);
  g->val = (double *)malloc(sizeof(double) * g->cap);
  g->order = (int64_t *)realloc(o.order, sizeof(int64_t) * g->cap);
  for (int64_t i = 0; i < g->cap; i++) g->key[i] = -1;
  for (int64_t i = 0; i < o.cap; i++)
    if (o.key[i] != -1) { int64_t j = gr_slot(g, o.key[i]); g->key[j] = o.key[i]; g->val[j] = o.val[i]; }
  free(o.key); free(o.val);
}
static void gr_set(growable *g, int64_t k, double v) {
  if (2 * (g->n + 1) > g->cap) gr_grow(g);
  int64_t i = gr_slot(g, k);
  if (g->key[i] == -1) { g->key[i] = k; g->order[g->n++] = k; }
  g->val[i] = v;
}
static double gr_call(void *env, int64_t k) {
  /* haskey, then push the default, then getindex: the two-probe pattern of :12-18 */
  growable *g = (growable *)env;
  int64_t i = gr_slot(g, k);
  if (g->key[i] == -1) { gr_set(g, k, g->dflt); i = gr_slot(g, k); }
  return g->val[i];
}

/* xoshiro256++ for shuffle(); not Julia's stream -- sweep order only changes the path to the
 * fixed point (SURVEY 8c iii) */
static uint64_t rng_s[4];
static inline uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
static uint64_t rng_next(void) {
  uint64_t r = rotl(rng_s[0] + rng_s[3], 23) + rng_s[0], t = rng_s[1] << 17;
  rng_s[2] ^= rng_s[0]; rng_s[3] ^= rng_s[1]; rng_s[1] ^= rng_s[2]; rng_s[0] ^= rng_s[3];
  rng_s[2] ^= t; rng_s[3] = rotl(rng_s[3], 45);
  return r;
}
static void rng_seed(uint64_t seed) {
  for (int i = 0; i < 4; i++) { seed += 0x9e3779b97f4a7c15ULL; rng_s[i] = hash64(seed); }
}
static void shuffle(int64_t *a, int64_t n) {
  for (int64_t i = n - 1; i > 0; i--) {
    int64_t j = (int64_t)(rng_next() % (uint64_t)(i + 1));
    int64_t t = a[i]; a[i] = a[j]; a[j] = t;
  }
}

/* ------------------------------------------------------------------------------------
 * Solvers.  kind: 0 = nu, 1 = logz, 2 = gap, 3 = z (linear)
 * mode: 0 = Jacobi over all free cells (dense), 1 = shuffled Gauss-Seidel over all free cells
 * (dense), 2 = faithful: Growable discovery from `start`, shuffled Gauss-Seidel, nested
 * recomputation exactly as the reference's closures do it.
 * Stop rule: sum |new - old| (+ newly discovered states) < tol   value_iteration.jl:74,82
 * ---------------------------------------------------------------------------------- */
typedef struct {
  int64_t sweeps;
  double diff;      /* last sweep's sum |delta| */
  double seconds;
  int64_t backups;  /* state backups performed */
  int64_t nstates;
} orc_stats;

static double now_s(void) {
  struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

static double backup_one(int kind, const orc_grid *g, const orc_q *q, orc_fn nu, void *nuenv, orc_fn X,
                         void *Xenv, int64_t s, double beta) {
  int terminal = (s == g->goal); /* terminal_condition   GridWorld.jl:118 */
  switch (kind) {
    case 0: return terminal ? 0.0 : truevalue_recursion(g, s, X, Xenv);            /* :66-70 */
    case 1: return terminal ? 0.0 - q->V(q->Venv, s) : logz_recursion(q, s, X, Xenv, beta); /* :93-98 */
    case 2: return terminal ? 0.0 : optimality_gap_recursion(q, nu, nuenv, s, X, Xenv);     /* :133-137 */
    default: return terminal ? exp(0.0 - q->V(q->Venv, s)) : z_recursion(q, s, X, Xenv, beta); /* :28-34 */
  }
}
static double damp(int kind, double new_, double old) {
  if (kind == 1) return logaddexp_(log(0.5) + new_, log(1 - 0.5) + old); /* :101-102 */
  if (kind == 3) return 0.5 * new_ + (1 - 0.5) * old;                    /* :37-38 */
  return new_;
}
static double absdiff(int kind, double new_, double old) {
  return kind == 3 ? fabs(log(new_) - log(old)) : fabs(new_ - old);      /* :35,:71,:100,:139 */
}

int orc_solve(int kind, int mode, int64_t H, int64_t W, const uint8_t *walls, int64_t goal,
              double randomness, double step_cost, double reward, double gamma, double reward_scale,
              const double *V, const double *Qtab, const double *nu, double beta, double tol,
              int64_t max_sweeps, int64_t start, uint64_t seed, int nthreads, double *out, orc_stats *st) {
  orc_grid g = {H, W, walls, goal, randomness, step_cost, reward, gamma, reward_scale};
  orc_q q = {&g, dense_fn, (void *)V, Qtab};
  int64_t n = H * W;
  double dflt = kind == 3 ? 1.0 : 0.0;
  double t0 = now_s();
  int64_t sweeps = 0, backups = 0;
  double diff = INFINITY;
  rng_seed(seed);
  (void)nthreads;
  if (mode == 2) {
    /* faithful: the iterated function is a Growable seeded at `start`; nu iterates over its
     * own keys (:65), logz/gap/z over keys(V.dict) (:27,:92,:132) -- V here being dense, all
     * free cells -- and every evaluation of a neighbour goes through the Dict. */
    growable X; gr_init(&X, dflt); gr_set(&X, start, dflt);
    growable Vg, Ng; int useV = (V != NULL), useN = (nu != NULL);
    if (useV) { gr_init(&Vg, 0.0); for (int64_t s = 0; s < n; s++) if (!walls[s]) gr_set(&Vg, s, V[s]); q.V = gr_call; q.Venv = &Vg; }
    if (useN) { gr_init(&Ng, 0.0); for (int64_t s = 0; s < n; s++) if (!walls[s]) gr_set(&Ng, s, nu[s]); }
    int64_t *keys = (int64_t *)malloc(sizeof(int64_t) * n);
    while (sweeps < max_sweeps) {
      int64_t old_len = X.n, nk;
      if (kind == 0) { nk = X.n; memcpy(keys, X.order, sizeof(int64_t) * nk); }
      else { nk = Vg.n; memcpy(keys, Vg.order, sizeof(int64_t) * nk); }
      shuffle(keys, nk);
      diff = 0.0;
      for (int64_t k = 0; k < nk; k++) {
        int64_t s = keys[k];
        double new_ = backup_one(kind, &g, &q, gr_call, useN ? (void *)&Ng : NULL, gr_call, &X, s, beta);
        double old = gr_call(&X, s);
        diff += absdiff(kind, new_, old);
        gr_set(&X, s, damp(kind, new_, gr_call(&X, s)));
      }
      backups += nk; sweeps++;
      diff += (double)(X.n - old_len);
      if (diff < tol) break;
    }
    for (int64_t s = 0; s < n; s++) out[s] = NAN;
    for (int64_t k = 0; k < X.n; k++) out[X.order[k]] = gr_call(&X, X.order[k]);
    st->nstates = X.n;
    free(keys); gr_free(&X); if (useV) gr_free(&Vg); if (useN) gr_free(&Ng);
  } else {
    int64_t nf = 0;
    int64_t *cells = (int64_t *)malloc(sizeof(int64_t) * n);
    for (int64_t s = 0; s < n; s++) if (!walls[s]) cells[nf++] = s;
    double *X = (double *)malloc(sizeof(double) * n), *Y = (double *)malloc(sizeof(double) * n);
    for (int64_t s = 0; s < n; s++) X[s] = Y[s] = dflt;
    while (sweeps < max_sweeps) {
      diff = 0.0;
      if (mode == 1) {
        shuffle(cells, nf);
        for (int64_t k = 0; k < nf; k++) {
          int64_t s = cells[k];
          double new_ = backup_one(kind, &g, &q, dense_fn, (void *)nu, dense_fn, X, s, beta);
          diff += absdiff(kind, new_, X[s]);
          X[s] = damp(kind, new_, X[s]);
        }
      } else {
#ifdef _OPENMP
#pragma omp parallel for reduction(+ : diff) schedule(static) num_threads(nthreads > 0 ? nthreads : 1)
#endif
        for (int64_t k = 0; k < nf; k++) {
          int64_t s = cells[k];
          double new_ = backup_one(kind, &g, &q, dense_fn, (void *)nu, dense_fn, X, s, beta);
          diff += absdiff(kind, new_, X[s]);
          Y[s] = damp(kind, new_, X[s]);
        }
        double *t = X; X = Y; Y = t;
      }
      backups += nf; sweeps++;
      if (diff < tol) break;
    }
    for (int64_t s = 0; s < n; s++) out[s] = walls[s] ? NAN : X[s];
    st->nstates = nf;
    free(cells); free(X); free(Y);
  }
  st->sweeps = sweeps; st->diff = diff; st->backups = backups; st->seconds = now_s() - t0;
  return 0;
}

/* One backup of every free cell from X into out (no damping unless kind 1/3): used to check a
 * single GPU sweep bit-for-structure. */
int orc_sweep(int kind, int64_t H, int64_t W, const uint8_t *walls, int64_t goal, double randomness,
              double step_cost, double reward, double gamma, double reward_scale, const double *V,
              const double *Qtab, const double *nu, double beta, const double *X, double *out,
              double *diff_out) {
  orc_grid g = {H, W, walls, goal, randomness, step_cost, reward, gamma, reward_scale};
  orc_q q = {&g, dense_fn, (void *)V, Qtab};
  double diff = 0.0;
  for (int64_t s = 0; s < H * W; s++) {
    if (walls[s]) { out[s] = NAN; continue; }
    double new_ = backup_one(kind, &g, &q, dense_fn, (void *)nu, dense_fn, (void *)X, s, beta);
    diff += absdiff(kind, new_, X[s]);
    out[s] = damp(kind, new_, X[s]);
  }
  if (diff_out) *diff_out = diff;
  return 0;
}

/* logz_updateV: V + logz/beta   value_iteration.jl:119-124 */
int orc_update_v(int64_t n, const uint8_t *walls, const double *V, const double *logz, double beta, double *out) {
  for (int64_t s = 0; s < n; s++) out[s] = walls[s] ? NAN : V[s] + logz[s] / beta;
  return 0;
}

/* kl_cost   path_functions.jl:17-20 ;  klterm(x,y) = xlogx(x) - xlogy(x,y) */
double orc_kl_cost(int64_t H, int64_t W, const uint8_t *walls, int64_t goal, double randomness,
                   double step_cost, double reward, double gamma, double reward_scale, const double *V,
                   const double *Qtab, int64_t s) {
  orc_grid g = {H, W, walls, goal, randomness, step_cost, reward, gamma, reward_scale};
  orc_q q = {&g, dense_fn, (void *)V, Qtab};
  double out = 0.0;
  for (int a = 0; a < NA; a++) {
    double x = action_probability(&q, s, a), y = 0.25;
    out += (x == 0 ? 0.0 : x * log(x)) - (x == 0 ? 0.0 : x * log(y));
  }
  return out;
}

/* ------------------------------------------------------------------------------------
 * Policy table: the per-state record the sampler reads.  8 x 4 bytes per cell:
 *   [0..2] float  c_k = (w_0+..+w_k)/sum(w): the running sums StatsBase's walk compares
 *   [3]    uint32 bit a set <=> take_step(s,a) is blocked by a wall
 *   [4..7] float  beta * fitness(s,a)
 * w_a = action_probability(ctrl,Q,s,a)  (choose_action, path_functions.jl:12-15)
 * ---------------------------------------------------------------------------------- */
static int orc_table_threads_ = 1; /* cells are independent: the full-size test grids build their table on all cores */
void orc_set_table_threads(int n) { orc_table_threads_ = n > 0 ? n : 1; }
int orc_policy_table(int64_t H, int64_t W, const uint8_t *walls, int64_t goal, double randomness,
                     double step_cost, double reward, double gamma, double reward_scale, const double *V,
                     const double *Qtab, double beta, float *table) {
  orc_grid g = {H, W, walls, goal, randomness, step_cost, reward, gamma, reward_scale};
  orc_q q = {&g, dense_fn, (void *)V, Qtab};
  memset(table, 0, sizeof(float) * 8 * H * W);
#ifdef _OPENMP
#pragma omp parallel for schedule(static, 4096) num_threads(orc_table_threads_)
#endif
  for (int64_t s = 0; s < H * W; s++) {
    if (walls[s]) continue;
    double w[NA], sum = 0.0, cw = 0.0;
    for (int a = 0; a < NA; a++) { w[a] = action_probability(&q, s, a); sum += w[a]; }
    uint32_t mask = 0;
    for (int a = 0; a < NA; a++) {
      cw += w[a];
      if (a < 3) table[8 * s + a] = (float)(cw / sum);
      table[8 * s + 4 + a] = (float)(beta * fitness(&q, s, a));
      if (walls[s + step_off(&g, a)]) mask |= 1u << a;
    }
    memcpy(&table[8 * s + 3], &mask, 4);
  }
  return 0;
}

/* ------------------------------------------------------------------------------------
 * Philox4x32-10 (Salmon et al., SC'11), counter-based.  counter = (step, 0, path_lo, path_hi),
 * key = (seed_lo, seed_hi).  Four 32-bit words per step = the four uniforms of the contract.
 * ---------------------------------------------------------------------------------- */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
  for (int r = 0; r < 10; r++) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
static inline float u01(uint32_t x) { return (float)(x >> 8) * 5.9604644775390625e-08f; } /* 2^-24 */

/* ------------------------------------------------------------------------------------
 * Rollout with the fp32 policy table (what the CUDA kernel must reproduce bit-for-bit).
 *
 * Injected-uniform contract, per step k of path p, four uniforms in this order:
 *   u0 action: StatsBase inverse-CDF walk  i=0; while (c_i < u0 && i < 3) i++   (t = u*sum(w))
 *   u1 stop  : estimator A only, after the weight update: u1 >= gamma stops
 *   u2 slip  : u2 < randomness            (rand() < randomness, GridWorld.jl:114)
 *   u3 slip action floor(4*u3), always consumed (eager ifelse, GridWorld.jl:114)
 * Estimator 0 (A): logw += beta*fit[a]; stop w.p. 1-gamma; at the goal logw -= V(goal).
 * Estimator 1 (B): no sampled stop; acc = logaddexp(acc, t*log(gamma) + log(1-gamma) + logW)
 *                  per step, at the goal acc (+)= t*log(gamma) + logW - V(goal); horizon cap.
 * Path p starts at starts[p % nstarts].  Record: logw[p], meta[p] = a0 | (len << 2), len =
 * number of sampled transitions (s,a,s').  Paths starting at the goal have a0 = 0, len = 0.
 * ---------------------------------------------------------------------------------- */
static inline float logaddexpf_(float x, float y) {
  if (x == -INFINITY) return y;
  if (y == -INFINITY) return x;
  float m = x > y ? x : y;
  return m + log1pf(expf(-fabsf(x - y)));
}
int orc_rollout(int64_t H, int64_t goal, const float *table, double gamma, double randomness,
                double Vgoal, int estimator, int64_t horizon, int64_t path0, int64_t npaths,
                const int32_t *starts, int64_t nstarts, uint64_t seed, const float *inj,
                int64_t inj_steps, float *logw_out, uint32_t *meta_out, int32_t *trace,
                int64_t trace_steps, int nthreads, int rec_bytes) {
  /* the stop and slip tests compare the (float) uniform with the Float64 gamma / randomness of the problem, as the
   * reference's rand() < randomness does (GridWorld.jl:114): for u = k 2^-24 that is k >= ceil(gamma 2^24) exactly */
  const double gam = gamma, rho = randomness;
  const float vgoal = (float)Vgoal;
  const float lgam = (float)log(gamma), l1mg = (float)log(1 - gamma);
  const uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  const int64_t off[4] = {1, -1, H, -H};
  (void)nthreads;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1024) num_threads(nthreads > 0 ? nthreads : 1)
#endif
  for (int64_t q = 0; q < npaths; q++) {
    int64_t p = path0 + q;
    int64_t s = starts[p % nstarts];
    float logw = 0.0f, acc = -INFINITY;
    uint32_t a0 = 0, len = 0;
    int64_t t = 0;
    for (;; t++) {
      if (s == goal) {
        if (estimator == 0) logw -= vgoal;
        else acc = logaddexpf_(acc, (float)t * lgam + (logw - vgoal));
        break;
      }
      if (t >= horizon) break;
      float u[4];
      if (inj) {
        if (t >= inj_steps) break;
        memcpy(u, inj + ((size_t)q * inj_steps + t) * 4, sizeof u);
      } else {
        uint32_t ctr[4] = {(uint32_t)t, 0u, (uint32_t)p, (uint32_t)((uint64_t)p >> 32)}, r[4];
        orc_philox4x32_10(ctr, key, r);
        for (int i = 0; i < 4; i++) u[i] = u01(r[i]);
      }
      float rec16[8];
      const float *rec = table + 8 * s;
      if (rec_bytes == 16) {
        /* the 16-byte record of large grids (V-determined Q: fitness does not depend on the action, PIQL.jl:87-89):
         * {c0 | (mask & 3) << 30, c1 | (mask >> 2) << 30, c2, beta*fitness}; probabilities are < 2, so bits 30-31 are free */
        uint32_t w[4]; memcpy(w, table + 4 * s, 16);
        uint32_t c0 = w[0] & 0x3fffffffu, c1 = w[1] & 0x3fffffffu, m = (w[0] >> 30) | ((w[1] >> 30) << 2);
        memcpy(&rec16[0], &c0, 4); memcpy(&rec16[1], &c1, 4); memcpy(&rec16[2], &w[2], 4); memcpy(&rec16[3], &m, 4);
        for (int a = 0; a < 4; a++) memcpy(&rec16[4 + a], &w[3], 4);
        rec = rec16;
      }
      uint32_t mask; memcpy(&mask, rec + 3, 4);
      int a = 0;
      while (rec[a] < u[0] && a < 3) a++;
      logw += rec[4 + a];
      if (t == 0) a0 = (uint32_t)a;
      if (estimator == 0) {
        if ((double)u[1] >= gam) { t++; break; }
      } else {
        acc = logaddexpf_(acc, ((float)t * lgam + l1mg) + logw);
      }
      int as = (int)(4.0f * u[3]); if (as > 3) as = 3;
      int aa = ((double)u[2] < rho) ? as : a;
      int64_t s1 = ((mask >> aa) & 1u) ? s : s + off[aa];
      if (trace && t < trace_steps) {
        int32_t *tr = trace + ((size_t)q * trace_steps + t) * 3;
        tr[0] = (int32_t)s; tr[1] = a; tr[2] = (int32_t)s1;
      }
      s = s1; len++;
    }
    logw_out[q] = estimator == 0 ? logw : acc;
    meta_out[q] = a0 | (len << 2);
  }
  return 0;
}

/* ------------------------------------------------------------------------------------
 * fitness_reward_estimate / direct_reward_estimate   path_functions.jl:34-50 over path_transitions :23-32
 * (as evidently intended: the file does not load as written -- `ctrl.reward` for ctrl.reward_function, `return out`
 * of an undefined name in the generator, no return of the running totals).  Exactly nsteps transitions per path, no
 * stop rule, the goal is not special.  After each transition:
 *   fit[t] = V(s0) + sum fitness(s,a)            dir[t] = V(s0) + sum [reward_function(s,a,s') - kl_cost(s)]
 * fp32 version (what the CUDA kernel must reproduce bit-for-bit): reads the fp32 policy records of beta = 1
 * (rec[4+a] = fitness) and an fp32 kl_cost per cell, sums in float in the order below.
 * ---------------------------------------------------------------------------------- */
int orc_path_estimates(int64_t H, int64_t goal, const float *table, const float *kl, const double *V,
                       double randomness, double step_cost, double reward, int64_t nsteps, int64_t path0,
                       int64_t npaths, const int32_t *starts, int64_t nstarts, uint64_t seed, const float *inj,
                       int64_t inj_steps, float *fit, float *dir, int32_t *trace) {
  const double rho = randomness;
  const float rg = (float)reward, rs = (float)(-step_cost);
  const uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  const int64_t off[4] = {1, -1, H, -H};
  for (int64_t q = 0; q < npaths; q++) {
    int64_t p = path0 + q;
    int64_t s = starts[p % nstarts];
    float tf = (float)V[s], td = tf;
    for (int64_t t = 0; t < nsteps; t++) {
      float u[4];
      if (inj) memcpy(u, inj + ((size_t)q * inj_steps + t) * 4, sizeof u);
      else {
        uint32_t ctr[4] = {(uint32_t)t, 0u, (uint32_t)p, (uint32_t)((uint64_t)p >> 32)}, r[4];
        orc_philox4x32_10(ctr, key, r);
        for (int i = 0; i < 4; i++) u[i] = u01(r[i]);
      }
      const float *rec = table + 8 * s;
      uint32_t mask; memcpy(&mask, rec + 3, 4);
      int a = 0;
      while (rec[a] < u[0] && a < 3) a++;                        /* choose_action :12-15 */
      int as = (int)(4.0f * u[3]); if (as > 3) as = 3;
      int aa = ((double)u[2] < rho) ? as : a;                    /* propagator  GridWorld.jl:113-115 */
      int64_t s1 = ((mask >> aa) & 1u) ? s : s + off[aa];
      tf = tf + rec[4 + a];                                      /* total += fitness(ctrl,V,Q,s,a)  :38 */
      float rw = (s1 == goal) ? rg : rs;                         /* reward_function  GridWorld.jl:116 */
      td = td + (rw - kl[s]);                                    /* total += reward - kl_cost  :47 */
      fit[q * nsteps + t] = tf;
      dir[q * nsteps + t] = td;
      if (trace) { int32_t *tr = trace + (q * nsteps + t) * 3; tr[0] = (int32_t)s; tr[1] = a; tr[2] = (int32_t)s1; }
      s = s1;
    }
  }
  return 0;
}

/* The same two estimates evaluated the reference's way -- Float64, every fitness / kl_cost / reward_function call
 * through the closures -- along a GIVEN sequence of (s, a, s') triples: pins the fp32 tables against the algebra. */
int orc_path_estimates_f64(int64_t H, int64_t W, const uint8_t *walls, int64_t goal, double randomness,
                           double step_cost, double reward, double gamma, double reward_scale, const double *V,
                           const double *Qtab, const int32_t *triples, int64_t nsteps, double *fit, double *dir) {
  orc_grid g = {H, W, walls, goal, randomness, step_cost, reward, gamma, reward_scale};
  orc_q q = {&g, dense_fn, (void *)V, Qtab};
  double tf = V[triples[0]], td = tf;
  for (int64_t t = 0; t < nsteps; t++) {
    int64_t s = triples[3 * t], s1 = triples[3 * t + 2];
    int a = triples[3 * t + 1];
    tf += fitness(&q, s, a);
    td += reward_fn(&g, s1) - orc_kl_cost(H, W, walls, goal, randomness, step_cost, reward, gamma, reward_scale, V, Qtab, s);
    fit[t] = tf; dir[t] = td;
  }
  return 0;
}

/* Faithful fp64 sampler: choose_action + propagator exactly as the reference evaluates them
 * (weights from action_probability in Float64, t = u*sum(w), walk cw < t), for one path with
 * injected uniforms.  Used to show the fp32-table sampler takes the same actions. */
int orc_rollout_f64(int64_t H, int64_t W, const uint8_t *walls, int64_t goal, double randomness,
                    double step_cost, double reward, double gamma, double reward_scale, const double *V,
                    const double *Qtab, double beta, int64_t start, const double *inj, int64_t nsteps,
                    int stop_rule, int32_t *trace, double *logw_out, int64_t *len_out) {
  orc_grid g = {H, W, walls, goal, randomness, step_cost, reward, gamma, reward_scale};
  orc_q q = {&g, dense_fn, (void *)V, Qtab};
  int64_t s = start, t = 0;
  double logw = 0.0;
  for (; t < nsteps; t++) {
    if (s == goal) { logw -= V[s]; break; }
    const double *u = inj + 4 * t;
    double w[NA], sum = 0.0;
    for (int a = 0; a < NA; a++) { w[a] = action_probability(&q, s, a); sum += w[a]; }
    double tt = u[0] * sum, cw = w[0];
    int a = 0;
    while (cw < tt && a < 3) { a++; cw += w[a]; }
    logw += beta * fitness(&q, s, a);
    if (stop_rule && u[1] >= gamma) { t++; break; }
    int as = (int)(4.0 * u[3]); if (as > 3) as = 3;
    int aa = (u[2] < randomness) ? as : a;
    int64_t s1 = take_step(&g, s, aa);
    if (trace) { trace[3 * t] = (int32_t)s; trace[3 * t + 1] = a; trace[3 * t + 2] = (int32_t)s1; }
    s = s1;
  }
  *logw_out = logw; *len_out = t;
  return 0;
}

/* ------------------------------------------------------------------------------------
 * Log-sum-exp reduction of path records into per-(s0,a0) statistics (fp64 accumulation):
 * lse_max, lse_sum (= sum exp(logw - max)), count; transitions = sum of len.
 * Path q (global id path0+q) belongs to start slot (path0+q) % nstarts.
 * ---------------------------------------------------------------------------------- */
int orc_reduce(int64_t path0, int64_t npaths, int64_t nstarts, const float *logw, const uint32_t *meta,
               double *mx /*nstarts*4*/, double *sm, uint64_t *cnt, uint64_t *transitions) {
  for (int64_t i = 0; i < nstarts * NA; i++) { mx[i] = -INFINITY; sm[i] = 0.0; cnt[i] = 0; }
  uint64_t tr = 0;
  for (int64_t q = 0; q < npaths; q++) {
    int64_t b = ((path0 + q) % nstarts) * NA + (meta[q] & 3u);
    if ((double)logw[q] > mx[b]) mx[b] = logw[q];
    tr += meta[q] >> 2;
  }
  for (int64_t q = 0; q < npaths; q++) {
    int64_t b = ((path0 + q) % nstarts) * NA + (meta[q] & 3u);
    sm[b] += exp((double)logw[q] - mx[b]);
    cnt[b]++;
  }
  *transitions = tr;
  return 0;
}

/* ------------------------------------------------------------------------------------
 * make_gridworld's flood fill: wall every cell not reachable from the goal
 * (get_reachable + the loop in make_gridworld, GridWorld.jl:63-88).  In place.
 * ---------------------------------------------------------------------------------- */
int64_t orc_wall_unreachable(int64_t H, int64_t W, uint8_t *walls, int64_t goal) {
  int64_t n = H * W, head = 0, tail = 0, nreach = 0;
  uint8_t *seen = (uint8_t *)calloc(n, 1);
  int64_t *queue = (int64_t *)malloc(sizeof(int64_t) * n);
  const int64_t off[4] = {1, -1, H, -H};
  queue[tail++] = goal; seen[goal] = 1;
  while (head < tail) {
    int64_t s = queue[head++]; nreach++;
    for (int a = 0; a < 4; a++) {
      int64_t s1 = s + off[a];
      if (s1 < 0 || s1 >= n || walls[s1] || seen[s1]) continue;
      seen[s1] = 1; queue[tail++] = s1;
    }
  }
  for (int64_t s = 0; s < n; s++) if (!seen[s]) walls[s] = 1;
  free(seen); free(queue);
  return nreach;
}

/* ------------------------------------------------------------------------------------
 * "Optimised CPU variant" of update_ν!: dense arrays, Jacobi, OpenMP over
 * cells, no closures or Dict -- the same arithmetic in the same order as truevalue_recursion/pa
 * above.  Runs exactly `sweeps` sweeps from x = 0 and returns the last sweep's sum |delta|.
 * ---------------------------------------------------------------------------------- */
double orc_nu_jacobi_dense(int64_t H, int64_t W, const uint8_t *walls, int64_t goal, double randomness,
                           double step_cost, double reward, double gamma, double reward_scale, int64_t sweeps,
                           int nthreads, double *out) {
  const int64_t n = H * W;
  const int64_t off[4] = {1, -1, H, -H};
  const double w_on = 1.0 * (1.0 - randomness) + randomness / NA, w_off = 0.0 * (1.0 - randomness) + randomness / NA;
  const double r_goal = reward_scale * reward, r_step = reward_scale * -step_cost;
  double *X = (double *)calloc(n, sizeof(double)), *Y = (double *)calloc(n, sizeof(double));
  double diff = 0.0;
  (void)nthreads;
  for (int64_t it = 0; it < sweeps; it++) {
    diff = 0.0;
#ifdef _OPENMP
#pragma omp parallel for reduction(+ : diff) schedule(static) num_threads(nthreads > 0 ? nthreads : 1)
#endif
    for (int64_t s = 0; s < n; s++) {
      if (walls[s]) continue;
      double nw = 0.0;
      if (s != goal) {
        double f[NA], r[NA], x[NA];
        for (int aa = 0; aa < NA; aa++) {
          const int64_t s1 = walls[s + off[aa]] ? s : s + off[aa];
          f[aa] = X[s1];
          r[aa] = s1 == goal ? r_goal : r_step;
        }
        for (int a = 0; a < NA; a++) {
          double pv = 0.0, pr = 0.0;
          for (int aa = 0; aa < NA; aa++) {
            const double w = a == aa ? w_on : w_off;
            pv += w * f[aa];
            pr += w * r[aa];
          }
          x[a] = LOG_PRIOR + gamma * pv + pr;
        }
        nw = lse4(x);
      }
      diff += fabs(nw - X[s]);
      Y[s] = nw;
    }
    double *t = X; X = Y; Y = t;
  }
  for (int64_t s = 0; s < n; s++) out[s] = walls[s] ? NAN : X[s];
  free(X); free(Y);
  return diff;
}

int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* ======================================================================================================
 * Tabulated ControlProblems (any finite problem behind the closure interface of src/PIQL.jl:12-23):
 * states 0..ns-1, actions 0..na-1, CSR over (state, action) of (successor, probability, reward).
 *   propagator_average(s,a,f) = sum_e p_e f(s'_e);  logpropagatorexp(s,a,f) = log sum_e p_e exp f(s'_e)
 * The same recursions as above (PIQL.jl:55-130, value_iteration.jl:62-169), V-determined Q, Jacobi sweeps.
 * Used to check the CUDA table path and, through a tabulated gridworld, against the grid restatement.
 * ====================================================================================================== */
typedef struct {
  int64_t ns; int na;
  const int64_t *ptr; const int32_t *succ; const double *prob, *rew, *lprior; const uint8_t *term;
  double gamma, scale;
} orc_table;

static double lse_n(const double *x, int n) {
  double m = -INFINITY;
  for (int i = 0; i < n; i++) if (x[i] > m) m = x[i];
  if (m == -INFINITY) return m;
  double s = 0.0;
  for (int i = 0; i < n; i++) s += exp(x[i] - m);
  return m + log(s);
}
static double t_pa(const orc_table *t, int64_t s, int a, const double *f) { /* propagator_average */
  double out = 0.0;
  for (int64_t e = t->ptr[s * t->na + a]; e < t->ptr[s * t->na + a + 1]; e++) out += t->prob[e] * f[t->succ[e]];
  return out;
}
static double t_rbar(const orc_table *t, int64_t s, int a) { /* average_reward  PIQL.jl:55-57 */
  double out = 0.0;
  for (int64_t e = t->ptr[s * t->na + a]; e < t->ptr[s * t->na + a + 1]; e++) out += t->prob[e] * (t->scale * t->rew[e]);
  return out;
}
static double t_cV(const orc_table *t, const double *V, int64_t s, int a) { /* controlV  PIQL.jl:63-66 */
  return t->gamma * t_pa(t, s, a, V) - V[s] + t_rbar(t, s, a);
}
static double t_log_pa(const orc_table *t, int64_t s, int a, const double *f) { /* logpropagatorexp */
  double m = -INFINITY;
  const int64_t e0 = t->ptr[s * t->na + a], e1 = t->ptr[s * t->na + a + 1];
  for (int64_t e = e0; e < e1; e++) { const double x = log(t->prob[e]) + f[t->succ[e]]; if (x > m) m = x; }
  if (m == -INFINITY) return m;
  double acc = 0.0;
  for (int64_t e = e0; e < e1; e++) acc += exp(log(t->prob[e]) + f[t->succ[e]] - m);
  return m + log(acc);
}
/* Q: optional table Q[s*na + a] (any Q callable of PIQL.jl:72-89, tabulated); NULL is the V-determined Q (s,a) -> controlV */
static double t_backup(int kind, const orc_table *t, const double *V, const double *Q, const double *nu, double beta, const double *X,
                       int64_t s) {
  double x[64] = {0}, cv[64] = {0}, q[64] = {0};
  const int na = t->na;
  if (t->term[s]) return kind == 1 ? 0.0 - V[s] : 0.0; /* value_iteration.jl:66-67,93-94,133-134 */
  if (kind == 0) { /* truevalue_recursion  PIQL.jl:95-101 */
    for (int a = 0; a < na; a++) x[a] = t->lprior[s * na + a] + t->gamma * t_pa(t, s, a, X) + t_rbar(t, s, a);
    return lse_n(x, na);
  }
  for (int a = 0; a < na; a++) {
    cv[a] = t_cV(t, V, s, a);
    q[a] = Q ? Q[s * na + a] : cv[a];
    x[a] = t->lprior[s * na + a] + q[a];
  }
  const double norm = lse_n(x, na); /* Qnormalization  PIQL.jl:72-75 */
  if (kind == 1) { /* logz_recursion  PIQL.jl:125-130 */
    for (int a = 0; a < na; a++)
      x[a] = t->lprior[s * na + a] + beta * cv[a] + (1.0 - beta) * (q[a] - norm) +
             logaddexp_(log(t->gamma) + t_log_pa(t, s, a, X), log(1.0 - t->gamma));
    return lse_n(x, na);
  }
  double out = 0.0; /* optimality_gap_recursion  value_iteration.jl:160-169 */
  for (int a = 0; a < na; a++) {
    const double cq = q[a] - norm;
    out += exp(t->lprior[s * na + a] + cq) * (cq - t_cV(t, nu, s, a) + t_pa(t, s, a, X));
  }
  return out;
}
int orc_table_solve(int kind, int64_t ns, int na, const int64_t *ptr, const int32_t *succ, const double *prob,
                    const double *rew, const double *lprior, const uint8_t *term, double gamma, double scale,
                    const double *V, const double *nu, double beta, const double *x0, double tol, int64_t max_sweeps,
                    double *out, int64_t *sweeps_out, double *residual_out, const double *Q) {
  if (na > 64) return 1;
  const orc_table t = {ns, na, ptr, succ, prob, rew, lprior, term, gamma, scale};
  double *X = (double *)malloc(sizeof(double) * (size_t)ns), *Y = (double *)malloc(sizeof(double) * (size_t)ns);
  for (int64_t s = 0; s < ns; s++) X[s] = x0 ? x0[s] : 0.0;
  int64_t sweeps = 0;
  double res = INFINITY;
  while (sweeps < max_sweeps) {
    double d = 0.0;
    for (int64_t s = 0; s < ns; s++) {
      const double nw = t_backup(kind, &t, V, Q, nu, beta, X, s);
      d += fabs(nw - X[s]);
      Y[s] = kind == 1 ? logaddexp_(log(0.5) + nw, log(0.5) + X[s]) : nw; /* eta = 0.5  value_iteration.jl:101-102 */
    }
    double *tmp = X; X = Y; Y = tmp;
    sweeps++;
    res = d;
    if (res < tol) break;
  }
  memcpy(out, X, sizeof(double) * (size_t)ns);
  free(X); free(Y);
  if (sweeps_out) *sweeps_out = sweeps;
  if (residual_out) *residual_out = res;
  return 0;
}

/* Path sampling on a tabulated problem with the sampler's float tables (running action probabilities cdf, beta*fitness bf,
 * running transition probabilities cum, terminal values vterm): the contract of piql_table_mc_logz, one Philox4x32-10 call
 * per step, words 0 / 1 / 2 -> action walk / stop test u >= gamma / successor walk.  Records: logw, meta = a0 | len << 4. */
int orc_table_rollout(int64_t ns, int na, const int64_t *ptr, const int32_t *succ, const uint8_t *term, const float *cdf,
                      const float *bf, const float *cum, const float *vterm, double gamma, const int32_t *starts,
                      int64_t nstarts, int64_t path0, int64_t npaths, uint64_t seed, int64_t horizon, float *logw_out,
                      uint32_t *meta_out) {
  (void)ns;
  const double gam = gamma;
  for (int64_t q = 0; q < npaths; q++) {
    const uint64_t gp = (uint64_t)(path0 + q);
    int64_t s = starts ? starts[gp % (uint64_t)nstarts] : (int64_t)(gp % (uint64_t)nstarts);
    uint32_t t = 0, a0 = 0;
    float logw = 0.f;
    for (;;) {
      if (term[s]) { logw = logw - vterm[s]; break; }
      if ((int64_t)t >= horizon) break;
      const uint32_t ctr[4] = {t, 0u, (uint32_t)gp, (uint32_t)(gp >> 32)}, key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
      uint32_t r[4];
      orc_philox4x32_10(ctr, key, r);
      const float u0 = u01(r[0]), u1 = u01(r[1]), u2 = u01(r[2]);
      int a = 0;
      while (a < na - 1 && cdf[s * na + a] < u0) a++;
      logw = logw + bf[s * na + a];
      if (t == 0) a0 = (uint32_t)a;
      if ((double)u1 >= gam) break;
      int64_t e = ptr[s * na + a];
      const int64_t last = ptr[s * na + a + 1] - 1;
      while (e < last && cum[e] < u2) e++;
      s = succ[e];
      t++;
    }
    logw_out[q] = logw;
    meta_out[q] = a0 | (t << 4);
  }
  return 0;
}
