/*
 * mppi_oracle.c -- CPU restatement of AutoRally's MPPI hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (autorally_b200/, include/) links,
 * loads or calls this file.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may use it, and there only as the checker / the CPU baseline.
 *
 * Parity status: PINNED.  The reference (rdesc/autorally) ships no tests, golden vectors or known-answer
 * fixtures for this path, so its own compiled code is the authority:
 *   - oracle/refbuild.py compiles the reference's MPPI sources from /root/reference for sm_100a
 *     (oracle/_ref/libautorally_ref.so); tests/golden/ref_gpu_golden.npz holds inputs + outputs of that
 *     library (made on a B200 by tests/golden/make_ref_gpu_golden.py) and pins costs, bookkeeping,
 *     weighting, smoothing and the nominal trajectory of this file in the CPU suite
 *     (tests/test_oracle_golden.py); tests/test_reference_gpu.py repeats the comparison live on the GPU box;
 *   - the dynamics (R7/R9: MLP forward + kinematics + Euler step) are also pinned against the reference's
 *     Python model (ml_pipeline/utils.py, imported unmodified by tests/golden/make_golden.py);
 *   - the Philox4x32-10 core against the published Random123 known-answer vectors (R1 is defined on
 *     injected identical noise: the generator changes from cuRAND XORWOW to Philox by specification).
 *
 * All citations are relative to /root/reference/autorally_control/include/autorally_control/
 * path_integral/ ("PI/").  Arithmetic is float32 with the reference's double-precision spots kept.
 * Where nvcc's default -fmad=true contracts a*b+c on the device, fmaf() is written explicitly and
 * the file is compiled with -ffp-contract=off, so the host compiler adds no contractions of its own.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

#define S_DIM 7
#define C_DIM 2
#define ORACLE_MAX_WIDTH 256 /* widest supported MLP layer */

/* Same field order as MPPICosts::CostParams (PI/costs.cuh:67-85) + the class member l1_cost_. */
typedef struct {
  float desired_speed, speed_coeff, track_coeff, max_slip_ang, slip_penalty, track_slop, crash_coeff;
  float steering_coeff, throttle_coeff, boundary_threshold, discount;
  int num_timesteps, grid_res;
  float r_c1[3], r_c2[3], trs[3];
  int l1_cost;
} oracle_cost_params;

typedef struct {
  int kind;                 /* 0 = NeuralNetModel, 1 = GeneralizedLinear<CarBasisFuncs,...> */
  const float *theta;       /* NN: [W1|b1|W2|b2|...] row-major (PI/neural_net_model.cu:120-141); BF: 4x25 */
  const int *net_structure; /* NN layer widths, e.g. 6,32,32,4 */
  int num_layers;           /* NN: number of entries of net_structure */
  int bdim_y;               /* BF: reference BLOCKSIZE_Y, fixes the partial-sum grouping (R8) */
  float dt;
  int negate_yaw_der;       /* NN only; BF always negates (PI/generalized_linear.cu:222) */
  float ctrl_lo[C_DIM], ctrl_hi[C_DIM];
} oracle_dynamics;

typedef struct {
  const float *channel0; /* row-major H x W, row = y, col = x (PI/costs.cu:218-223) */
  int width, height;
} oracle_costmap;

/* ------------------------------------------------------------------ dynamics ---- */

/* PI/neural_net_model.cu:311-323 (same if / else-if form for BF, PI/generalized_linear.cu:182-194) */
static void enforce_constraints(const oracle_dynamics *d, float *u) {
  for (int i = 0; i < C_DIM; i++) {
    if (u[i] < d->ctrl_lo[i]) u[i] = d->ctrl_lo[i];
    else if (u[i] > d->ctrl_hi[i]) u[i] = d->ctrl_hi[i];
  }
}

/* Device MLP, PI/neural_net_model.cu:357-410: k ascending FMA chain from 0, bias added after,
 * tanhf on hidden layers only.  fused != 0 selects the device (FMA-contracted) arithmetic,
 * fused == 0 the host twin used by computeNominalTraj (PI/neural_net_model.cu:205-230). */
static void nn_forward(const oracle_dynamics *d, const float *in, float *out, int fused) {
  float a[ORACLE_MAX_WIDTH], b[ORACLE_MAX_WIDTH];
  float *cur = a, *nxt = b;
  const int *ns = d->net_structure;
  for (int i = 0; i < ns[0]; i++) cur[i] = in[i];
  int off = 0;
  for (int l = 0; l < d->num_layers - 1; l++) {
    const float *W = d->theta + off;
    const float *bias = W + ns[l + 1] * ns[l];
    for (int j = 0; j < ns[l + 1]; j++) {
      float t = 0.0f;
      for (int k = 0; k < ns[l]; k++) {
        if (fused) t = fmaf(W[j * ns[l] + k], cur[k], t);
        else t = t + W[j * ns[l] + k] * cur[k];
      }
      t += bias[j];
      if (l < d->num_layers - 2) t = tanhf(t);
      nxt[j] = t;
    }
    off += ns[l + 1] * ns[l] + ns[l + 1];
    float *tmp = cur; cur = nxt; nxt = tmp;
  }
  for (int i = 0; i < ns[d->num_layers - 1]; i++) out[i] = cur[i];
}

/* The 25 hand-written basis functions of PI/car_bfs.cuh:44-120, with the reference's mixed
 * float/double/integer-literal arithmetic kept (C and C++ promote identically here; fabsf is used
 * where device C++ resolves fabs(float) to the float overload). */
static float car_basis(int idx, const float *s, const float *u) {
  const float vx = s[4], vy = s[5], wz = s[6], roll = s[3];
  const float steer = u[0], thr = u[1];
  const int moving = vx > .1;
  /* tan(alpha_f - steer): front slip term; atanf's argument is formed in double. */
  float tf = moving ? tanf(atanf(vy / vx + .45 * wz / vx) - steer) : tanf(-steer);
  float phi = 0;
  switch (idx) {
    case 0: phi = thr; break;
    case 1: phi = vx / 10.0; break;
    case 2: phi = sinf(steer) * tf / 1200.0; break;
    case 3: phi = sinf(steer) * tf * fabsf(tf) / 1440000.0; break;
    case 4: phi = sinf(steer) * powf(tf, 3) / 1728000000.0; break;
    case 5: phi = wz * vy / 25.0; break;
    case 6: phi = wz / 10.0; break;
    case 7: phi = vy / 10.0; break;
    case 8: phi = sinf(steer); break;
    case 9: phi = moving ? vy / vx / 40.0 : 0; break;
    case 10: phi = tf / 1400.0; break;
    case 11: phi = tf * fabsf(tf) / 1960000; break;          /* integer literal -> float divide */
    case 12: phi = powf(tf, 3) / 2744000000; break;          /* long literal -> float divide */
    case 13: phi = moving ? (vy / vx - .35 * wz / vx) / 40.0 : 0; break;
    case 14: {
      if (moving) { double r = vy / vx - .35 * wz / vx; phi = r * fabs(r) / 1600.0; }
      break;
    }
    case 15: phi = moving ? powf(vy / vx - .35 * wz / vx, 3) / 64000.0 : 0; break;
    case 16: phi = wz * vx / 50.0; break;
    case 17: phi = roll; break;
    case 18: phi = roll * wz; break;
    case 19: phi = roll * vx / 3.0; break;
    case 20: phi = roll * vx * wz / 5.0; break;
    case 21: phi = powf(vx, 2) / 100.0; break;
    case 22: phi = powf(vx, 3) / 1000.0; break;
    case 23: phi = powf(thr, 2); break;
    case 24: phi = powf(thr, 3); break;
  }
  return phi;
}

/* state derivative; sder[3..6] must be zero on entry for BF (they are accumulated into). */
static void state_deriv(const oracle_dynamics *d, const float *s, const float *u, float *sder, int fused) {
  /* kinematics: PI/neural_net_model.cu:346-355, PI/generalized_linear.cu:217-223 */
  float c = cosf(s[2]), sn = sinf(s[2]);
  if (fused) {
    sder[0] = fmaf(c, s[4], -(sn * s[5]));
    sder[1] = fmaf(sn, s[4], c * s[5]);
  } else {
    sder[0] = c * s[4] - sn * s[5];
    sder[1] = sn * s[4] + c * s[5];
  }
  sder[2] = s[6];
  if (d->kind == 1 || d->negate_yaw_der) sder[2] = -s[6];
  if (d->kind == 0) {
    float in[ORACLE_MAX_WIDTH], out[ORACLE_MAX_WIDTH];
    in[0] = s[3]; in[1] = s[4]; in[2] = s[5]; in[3] = s[6]; in[4] = u[0]; in[5] = u[1];
    nn_forward(d, in, out, fused);
    for (int i = 0; i < 4; i++) sder[3 + i] = out[i];
  } else if (fused) {
    /* PI/generalized_linear.cu:225-245: y-thread tdy sums basis functions tdy, tdy+BY, ... into a
     * private partial, then atomicAdd()s it into shared s_der (order not defined by the reference;
     * tdy ascending is used here). */
    int by = d->bdim_y > 0 ? d->bdim_y : 1;
    for (int tdy = 0; tdy < by; tdy++) {
      float part[4] = {0, 0, 0, 0};
      for (int i = tdy; i < 25; i += by) {
        float bf = car_basis(i, s, u);
        for (int j = 0; j < 4; j++) part[j] = fmaf(d->theta[j * 25 + i], bf, part[j]);
      }
      for (int j = 0; j < 4; j++) sder[3 + j] += part[j];
    }
  } else {
    /* host twin: theta_ * bf_vec_ (PI/generalized_linear.cu:159-166) */
    for (int j = 0; j < 4; j++) {
      float t = 0;
      for (int i = 0; i < 25; i++) t = t + d->theta[j * 25 + i] * car_basis(i, s, u);
      sder[3 + j] = t;
    }
  }
}

/* ------------------------------------------------------------------ costs ---- */

/* CUDA texture fetch, normalised coordinates, point filter, clamp (PI/costs.cu:143-149).  The texture
 * unit is third-party arithmetic; its texel selection was measured on B200
 * with tools/texprobe.cu (63 758 probes at and around every texel boundary of 5 widths, 0 mismatches,
 * profiles/texprobe_r01.txt): the normalised coordinate is TRUNCATED to 21 fractional bits, then
 * texel = clamp(floor(coord_q * dim), 0, dim-1).  A plain floorf(u * W) differs exactly at texel
 * boundaries, which grid-aligned start poses hit systematically. */
static int tex_index(float coord, int dim) {
  if (!(coord > 0.0f)) return 0; /* negative, zero and NaN clamp to texel 0 */
  double q = floor((double)coord * 2097152.0) / 2097152.0;
  double t = floor(q * (double)dim);
  return t >= (double)dim ? dim - 1 : (int)t;
}

static float tex_lookup(const oracle_costmap *m, float un, float vn) {
  return m->channel0[(size_t)tex_index(vn, m->height) * m->width + tex_index(un, m->width)];
}

static void coor_transform(const oracle_cost_params *p, float x, float y, float *u, float *v, float *w) {
  /* PI/costs.cu:351-357 */
  *u = fmaf(p->r_c1[0], x, p->r_c2[0] * y) + p->trs[0];
  *v = fmaf(p->r_c1[1], x, p->r_c2[1] * y) + p->trs[1];
  *w = fmaf(p->r_c1[2], x, p->r_c2[2] * y) + p->trs[2];
}

/* PI/costs.cu:396-409 and its parts :307-393.  ORDER MATTERS: the track cost sets the crash flag
 * before the crash cost is charged. */
static float compute_cost(const oracle_cost_params *p, const oracle_costmap *m, const float *s,
                          const float *u, const float *du, const float *vars, int *crash) {
  /* control cost :307-313 */
  float control = 0;
  control += p->steering_coeff * du[0] * (u[0] - du[0]) / (vars[0] * vars[0]);
  control += p->throttle_coeff * du[1] * (u[1] - du[1]) / (vars[1] * vars[1]);
  /* track cost :359-393 (device uses __cosf/__sinf; precise libm here) */
  float cy = cosf(s[2]), sy = sinf(s[2]);
  float xf = fmaf(0.5f, cy, s[0]), yf = fmaf(0.5f, sy, s[1]);
  float xb = fmaf(-0.5f, cy, s[0]), yb = fmaf(-0.5f, sy, s[1]);
  float uu, vv, ww;
  coor_transform(p, xf, yf, &uu, &vv, &ww);
  float front = tex_lookup(m, uu / ww, vv / ww);
  coor_transform(p, xb, yb, &uu, &vv, &ww);
  float back = tex_lookup(m, uu / ww, vv / ww);
  float track = (fabsf(front) + fabsf(back)) / 2.0;
  if (fabsf(track) < p->track_slop) track = 0;
  else track = p->track_coeff * track;
  if (front >= p->boundary_threshold || back >= p->boundary_threshold) crash[0] = 1;
  /* speed cost :315-326 */
  float err = s[4] - p->desired_speed;
  float sc = p->l1_cost ? fabsf(err) : err * err;
  float speed = p->speed_coeff * sc;
  /* crash cost :328-335, scaled in double with the host-side discount :402 */
  float crash_cost = (1.0 - p->discount) * (crash[0] > 0 ? p->crash_coeff : 0.0f);
  /* stabilizing cost :337-349 */
  float stab = 0;
  if (fabsf(s[4]) > 0.001) {
    float slip = -atanf(s[5] / fabsf(s[4]));
    stab = p->slip_penalty * (slip * slip);
    if (fabsf(slip) > p->max_slip_ang) stab += p->crash_coeff;
  }
  float cost = control + speed + crash_cost + track + stab;
  if (cost > 1e12 || isnan(cost)) cost = 1e12;
  return cost;
}

/* ------------------------------------------------------------------ rollouts ---- */

typedef struct {
  const oracle_dynamics *dyn; const oracle_cost_params *cp; const oracle_costmap *map;
  int n_global, r_begin, r_count, T, opt_delay;
  const float *state0, *U, *nu;
  float *du; float *costs; int *crash; float *final_state;
  int lo, hi;
} rollout_job;

/* rolloutKernel, PI/mppi_controller.cu:72-184.  du holds eps on entry ([r][t][j], index
 * C*T*r + t*C + j, :133) and the sampled un-clamped controls on exit (:153). */
static void *rollout_range(void *arg) {
  rollout_job *jb = (rollout_job *)arg;
  const int T = jb->T;
  for (int lr = jb->lo; lr < jb->hi; lr++) {
    const int r = jb->r_begin + lr; /* global rollout index drives the bookkeeping */
    float s[S_DIM], sder[S_DIM], u[C_DIM], du[C_DIM];
    int crash = 0;
    float running = 0;
    for (int i = 0; i < S_DIM; i++) { s[i] = jb->state0[i]; sder[i] = 0; }
    float *row = jb->du + (size_t)lr * T * C_DIM;
    for (int i = 0; i < T; i++) {
      for (int j = 0; j < C_DIM; j++) {
        if (r == 0 || i < jb->opt_delay) { du[j] = 0.0f; u[j] = jb->U[i * C_DIM + j]; }
        else if ((double)r >= .99 * (double)jb->n_global) { du[j] = row[i * C_DIM + j] * jb->nu[j]; u[j] = du[j]; }
        else { du[j] = row[i * C_DIM + j] * jb->nu[j]; u[j] = jb->U[i * C_DIM + j] + du[j]; }
        row[i * C_DIM + j] = u[j];
      }
      enforce_constraints(jb->dyn, u);
      if (i > 0) {
        float c = compute_cost(jb->cp, jb->map, s, u, du, jb->nu, &crash);
        running += (c - running) / (1.0 * i); /* float diff, double divide, double add, float store (:164) */
      }
      state_deriv(jb->dyn, s, u, sder, 1);
      for (int k = 0; k < S_DIM; k++) { s[k] = fmaf(sder[k], jb->dyn->dt, s[k]); sder[k] = 0; } /* :334-344 */
      if (fabsf(s[3]) > 1.57) crash = 1; /* getCrash, PI/costs.cu:301-305 */
    }
    jb->costs[lr] = running + 0.0f; /* terminalCost == 0, PI/costs.cu:411-414 */
    if (jb->crash) jb->crash[lr] = crash;
    if (jb->final_state) memcpy(jb->final_state + (size_t)lr * S_DIM, s, sizeof(s));
  }
  return NULL;
}

void oracle_rollouts(const oracle_dynamics *dyn, const oracle_cost_params *cp, const oracle_costmap *map,
                     int n_global, int r_begin, int r_count, int T, int opt_delay,
                     const float *state0, const float *U, const float *nu,
                     float *du, float *costs, int *crash, float *final_state, int num_threads) {
  if (num_threads < 1) num_threads = 1;
  if (num_threads > 256) num_threads = 256;
  if (num_threads > r_count) num_threads = r_count > 0 ? r_count : 1;
  pthread_t th[256];
  rollout_job jobs[256];
  for (int t = 0; t < num_threads; t++) {
    rollout_job jb = {dyn, cp, map, n_global, r_begin, r_count, T, opt_delay, state0, U, nu,
                      du, costs, crash, final_state,
                      (int)((long long)r_count * t / num_threads), (int)((long long)r_count * (t + 1) / num_threads)};
    jobs[t] = jb;
  }
  if (num_threads == 1) { rollout_range(&jobs[0]); return; }
  for (int t = 0; t < num_threads; t++) pthread_create(&th[t], NULL, rollout_range, &jobs[t]);
  for (int t = 0; t < num_threads; t++) pthread_join(th[t], NULL);
}

/* ---------------------------------------------------- importance weighting ---- */

/* PI/mppi_controller.cu:627-656 (host min / sums), :193-203 (normExpKernel), :219-267
 * (weightedReductionKernel: 64-rollout serial chunks, then the chunk partials serially). */
void oracle_weighting(const float *costs, const float *V, int N, int T, float gamma,
                      float *w_out, float *U_new, float *stats /* baseline, normalizer, trajectory_cost */) {
  float baseline = costs[0];
  for (int i = 0; i < N; i++) if (costs[i] < baseline) baseline = costs[i];
  float *w = w_out ? w_out : (float *)malloc(sizeof(float) * (size_t)N);
  for (int i = 0; i < N; i++) { float c2g = costs[i] - baseline; w[i] = expf(-gamma * c2g); }
  float Z = 0;
  for (int i = 0; i < N; i++) Z += w[i];
  float tc = 0;
  for (int i = 0; i < N; i++) tc += w[i] * w[i] / Z;
  const int stride = 64, nchunks = (N - 1) / stride + 1;
  for (int t = 0; t < T; t++) {
    float tot[C_DIM] = {0, 0};
    for (int c = 0; c < nchunks; c++) {
      float part[C_DIM] = {0, 0};
      for (int i = 0; i < stride; i++) {
        int r = stride * c + i;
        if (r < N) {
          float weight = w[r] / Z;
          for (int j = 0; j < C_DIM; j++) part[j] = fmaf(weight, V[(size_t)r * T * C_DIM + t * C_DIM + j], part[j]);
        }
      }
      for (int j = 0; j < C_DIM; j++) tot[j] += part[j];
    }
    for (int j = 0; j < C_DIM; j++) U_new[t * C_DIM + j] = tot[j];
  }
  stats[0] = baseline; stats[1] = Z; stats[2] = tc;
  if (!w_out) free(w);
}

/* Shard partials for the multi-GPU exchange (no reference counterpart):
 * b_g, Z_g, Q_g, W_g[T][2] relative to the shard's own minimum.  Straightforward float sums. */
void oracle_shard_partials(const float *costs, const float *V, int n, int T, float gamma, float *out /* 3 + 2T */) {
  float b = costs[0];
  for (int i = 0; i < n; i++) if (costs[i] < b) b = costs[i];
  double Z = 0, Q = 0;
  double *W = (double *)calloc((size_t)T * C_DIM, sizeof(double));
  for (int i = 0; i < n; i++) {
    float w = expf(-gamma * (costs[i] - b));
    Z += w; Q += (double)w * w;
    for (int k = 0; k < T * C_DIM; k++) W[k] += (double)w * V[(size_t)i * T * C_DIM + k];
  }
  out[0] = b; out[1] = (float)Z; out[2] = (float)Q;
  for (int k = 0; k < T * C_DIM; k++) out[3 + k] = (float)W[k];
  free(W);
}

/* ------------------------------------------------------ smoothing / nominal ---- */

/* PI/mppi_controller.cu:468-499: 5-tap filter over [hist0, hist1, U_0..U_{T-1}, U_{T-1}, U_{T-1}]. */
void oracle_savitsky_golay(float *U, const float *hist /* 2*C */, int T) {
  float filt[5] = {-3, 12, 17, 12, -3};
  for (int i = 0; i < 5; i++) filt[i] /= 35.0; /* Eigen: MatrixXf /= 35.0 -> float scalar */
  float *P = (float *)malloc(sizeof(float) * (size_t)(T + 4) * C_DIM);
  for (int i = 0; i < T + 4; i++)
    for (int j = 0; j < C_DIM; j++) {
      if (i < 2) P[i * C_DIM + j] = hist[C_DIM * i + j];
      else if (i < T + 2) P[i * C_DIM + j] = U[C_DIM * (i - 2) + j];
      else P[i * C_DIM + j] = U[C_DIM * (T - 1) + j];
    }
  for (int i = 0; i < T; i++)
    for (int j = 0; j < C_DIM; j++) {
      float acc = 0;
      for (int k = 0; k < 5; k++) acc = acc + filt[k] * P[(i + k) * C_DIM + j];
      U[C_DIM * i + j] = acc;
    }
  free(P);
}

/* host updateState, PI/neural_net_model.cu:280-288 / PI/generalized_linear.cu:140-147 */
void oracle_update_state(const oracle_dynamics *d, float *s, float *u) {
  float sder[S_DIM] = {0};
  enforce_constraints(d, u);
  state_deriv(d, s, u, sder, 0);
  for (int k = 0; k < S_DIM; k++) s[k] = s[k] + sder[k] * d->dt;
}

/* PI/mppi_controller.cu:501-519 */
void oracle_nominal_traj(const oracle_dynamics *d, const float *state, const float *U, int T,
                         float *state_solution, float *control_solution) {
  float s[S_DIM], u[C_DIM];
  memcpy(s, state, sizeof(s));
  for (int i = 0; i < T; i++) {
    for (int j = 0; j < S_DIM; j++) state_solution[i * S_DIM + j] = s[j];
    u[0] = U[2 * i]; u[1] = U[2 * i + 1];
    oracle_update_state(d, s, u);
    control_solution[2 * i] = u[0]; control_solution[2 * i + 1] = u[1];
  }
}

/* PI/mppi_controller.cu:527-554, including the stride != 1 branch that indexes the flat U_. */
void oracle_slide_control_seq(float *U, float *hist, const float *init_u, int T, int stride) {
  if (stride == 1) { hist[0] = hist[2]; hist[1] = hist[3]; hist[2] = U[0]; hist[3] = U[1]; }
  else { int t = stride - 2; for (int i = 0; i < 4; i++) hist[i] = U[t + i]; }
  for (int i = 0; i < T - stride; i++)
    for (int j = 0; j < C_DIM; j++) U[i * C_DIM + j] = U[(i + stride) * C_DIM + j];
  for (int j = 1; j <= stride; j++)
    for (int i = 0; i < C_DIM; i++) U[(T - j) * C_DIM + i] = init_u[i];
}

/* PI/mppi_controller.cu:560-568 */
void oracle_slide_state_seq(float *state_solution, int T, int stride) {
  for (int i = 0; i < T - stride; i++)
    for (int j = 0; j < S_DIM; j++) state_solution[i * S_DIM + j] = state_solution[(i + stride) * S_DIM + j];
}

/* computeControl(state), PI/mppi_controller.cu:600-675.  eps: [num_iters][N][T][2] injected noise.
 * Outputs: U (in/out, smoothed), V/costs/w of the last iteration (optional), stats[3],
 * state_solution[T*7], control_solution[T*2]. */
void oracle_compute_control(const oracle_dynamics *dyn, const oracle_cost_params *cp, const oracle_costmap *map,
                            int N, int T, int opt_delay, int num_iters, float gamma,
                            const float *state, float *U, const float *hist, const float *nu,
                            const float *eps, float *V_out, float *costs_out, int *crash_out, float *w_out,
                            float *stats, float *state_solution, float *control_solution, int num_threads) {
  float *V = V_out ? V_out : (float *)malloc(sizeof(float) * (size_t)N * T * C_DIM);
  float *costs = costs_out ? costs_out : (float *)malloc(sizeof(float) * (size_t)N);
  float *Unew = (float *)malloc(sizeof(float) * (size_t)T * C_DIM);
  for (int it = 0; it < num_iters; it++) {
    memcpy(V, eps + (size_t)it * N * T * C_DIM, sizeof(float) * (size_t)N * T * C_DIM);
    oracle_rollouts(dyn, cp, map, N, 0, N, T, opt_delay, state, U, nu, V, costs, crash_out, NULL, num_threads);
    oracle_weighting(costs, V, N, T, gamma, w_out, Unew, stats);
    memcpy(U, Unew, sizeof(float) * (size_t)T * C_DIM); /* U_ = du_ (replace), :663-667 */
  }
  oracle_savitsky_golay(U, hist, T);
  oracle_nominal_traj(dyn, state, U, T, state_solution, control_solution);
  free(Unew);
  if (!V_out) free(V);
  if (!costs_out) free(costs);
}

/* Single dynamics step with the device arithmetic (used to pin R7/R9 against the reference's
 * Python model): s <- s + f(s, clamp(u)) * dt; also returns the state derivative. */
void oracle_dynamics_step(const oracle_dynamics *d, float *s, float *u, float *sder_out) {
  float sder[S_DIM] = {0};
  enforce_constraints(d, u);
  state_deriv(d, s, u, sder, 1);
  if (sder_out) memcpy(sder_out, sder, sizeof(sder));
  for (int k = 0; k < S_DIM; k++) s[k] = fmaf(sder[k], d->dt, s[k]);
}

/* "Straight loop over the same weights" CPU baseline of BASELINE.md section 3 row B: dynamics only
 * (clamp -> MLP -> kinematics -> Euler), n rollouts x T steps, controls = U + nu * eps. */
void oracle_dynamics_rollouts(const oracle_dynamics *d, int n, int T, const float *state0, const float *U,
                              const float *nu, const float *eps, float *final_state) {
  for (int r = 0; r < n; r++) {
    float s[S_DIM];
    memcpy(s, state0, sizeof(s));
    for (int i = 0; i < T; i++) {
      float u[C_DIM];
      for (int j = 0; j < C_DIM; j++) u[j] = U[i * C_DIM + j] + eps[((size_t)r * T + i) * C_DIM + j] * nu[j];
      oracle_dynamics_step(d, s, u, NULL);
    }
    memcpy(final_state + (size_t)r * S_DIM, s, sizeof(s));
  }
}

/* ------------------------------------------------------------- noise sampler ---- */

/* Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11); constants and
 * round structure as published in Random123 (philox.h).  The reference uses cuRAND XORWOW
 * (PI/mppi_controller.cu:330-331,612); north_star replaces it, so this is the sampler's own oracle. */
void oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
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

/* The sampler's stream definition: for global rollout r and timestep pair q
 * (t = 2q, 2q+1) the counter is (q, r, call, controller) and the key (seed_lo, seed_hi); the four
 * outputs x0..x3 give eps[r][2q][0], eps[r][2q][1], eps[r][2q+1][0], eps[r][2q+1][1] by Box-Muller:
 *   radius = sqrt(-2 ln((x_a + 0.5) 2^-32)),  angle = 2 pi (x_b + 0.5) 2^-32 - pi,
 *   (x0,x1) -> radius*cos, radius*sin ; (x2,x3) likewise.  Evaluated in double here. */
void oracle_sample_noise(uint64_t seed, uint32_t call, uint32_t controller, int r_begin, int r_count, int T, float *eps) {
  uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  const double two_m32 = 1.0 / 4294967296.0, pi = 3.14159265358979323846;
  int Q = (T + 1) / 2;
  for (int lr = 0; lr < r_count; lr++)
    for (int q = 0; q < Q; q++) {
      uint32_t ctr[4] = {(uint32_t)q, (uint32_t)(r_begin + lr), call, controller}, x[4];
      oracle_philox4x32_10(ctr, key, x);
      double z[4];
      for (int h = 0; h < 2; h++) {
        double ua = ((double)x[2 * h] + 0.5) * two_m32;
        double ang = 2.0 * pi * ((double)x[2 * h + 1] + 0.5) * two_m32 - pi;
        double rad = sqrt(-2.0 * log(ua));
        z[2 * h] = rad * cos(ang); z[2 * h + 1] = rad * sin(ang);
      }
      for (int k = 0; k < 4; k++) {
        int t = 2 * q + k / 2;
        if (t < T) eps[((size_t)lr * T + t) * C_DIM + (k & 1)] = (float)z[k];
      }
    }
}
