/*
 * quadsim_oracle.c -- CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C, float64, one-env-at-a-time restatement of the reference's per-control-step hot path
 * (priban42/quad-swarm-rl-stable-baselines3, `gym_art/quadrotor_multi`).  It exists to CHECK the CUDA
 * product (tests/, __graft_entry__.smoke()) and to serve as the timed CPU baseline in bench.py; the
 * product never links, imports or calls it.
 *
 * Parity status: the reference's own tests hold no usable golden vectors for this path,
 * so this oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF: tests/golden/make_golden.py runs the
 * unmodified reference (numba JIT disabled so every random draw can be taped) and commits state/obs/reward/
 * done traces plus the tape of unit draws; tests/test_oracle_golden.py replays those tapes through this file.
 *
 * Each function cites the reference file:line it follows (paths relative to gym_art/quadrotor_multi/).
 * Iteration order deliberately mirrors the reference (per-drone loop, then env-level passes) so that in
 * "tape" mode the draws are consumed in the reference's order.  In "philox" mode the draws come from the
 * counter-based generator specified in DESIGN.md ("RNG contract"), identical to the CUDA kernels'.
 */
#include <math.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/quadsim.h"

#define QO_EPS_DYN 1e-6   /* quadrotor_dynamics.py:13 */
#define QO_EPS_COL 1e-5   /* quad_utils.py:11 */
#define QO_PI 3.14159265358979323846
#define QO_GRAV 9.81      /* quadrotor_dynamics.py:12 (floor code uses the module constant) */

/* RNG sites (DESIGN.md "RNG contract") */
enum { SITE_OU = 0, SITE_SENSOR = 1, SITE_SENSOR_IMPULSE = 2, SITE_SENSOR_RESET = 3, SITE_FLOOR_YAW = 4,
       SITE_PAIR = 5, SITE_OBST = 6, SITE_WALL = 7, SITE_CEILING = 8, SITE_DOWNWASH = 9,
       SITE_SPAWN = 10, SITE_SCENARIO = 11, SITE_CAMERA = 12 };

typedef struct {
    double pos[3], vel[3], rot[9], omega[3];
    double rot_damp[4], cmds_damp[4], ou[4];
    double acc[3], accm[3];
    double goal[3], spawn_point[3];
    int on_floor, crashed_floor, crashed_wall, crashed_ceiling;
    int prev_new_wall, prev_new_ceiling, prev_new_room, prev_obst_hit;
    uint32_t col_mask;              /* previous-step collision row */
    double dist_hist[5]; int dist_n;/* last 5 entries of distance_to_goal[i] (quadrotor_multi.py:651-655) */
    double sum1, sum3, sum5; int n1, n3, n5;
    int reached_goal, col_agent_ok, col_obst_ok;
    double ep_reward;
    /* fork mode: pre_controller state.  pid[2k] = last_error, pid[2k+1] = integral of PID k (cascade order) */
    double pid[24], angle, ang_vel;
} qo_drone;

/* formation scenarios (scenario_oracle.inc): the state of the reference's QuadrotorScenario object */
typedef struct {
    int formation, per_layer;           /* index into QUADS_FORMATION_LIST; num_agents_per_layer */
    double size, lowest, highest, layer_dist, center[3];
    int ctl_steps;                      /* control_step_for_sec */
    int increase; double speed;         /* dynamic_formations */
    double c1[3], c2[3];                /* swarm_vs_swarm goal centres */
    double bez[3][3];                   /* ep_rand_bezier control points */
    double goals[2 * QS_MAX_AGENTS + 6][3];
    int n_goals, n1_rows;
} qo_scen;

enum { QO_MG_PAIR = 0, QO_MG_FALLOFF, QO_MG_OBST, QO_MG_FLOOR, QO_MG_WALL, QO_MG_CEIL, QO_MG_YAW, QO_MG_RANK, QO_MG_REACH, QO_MG_LIFTOFF, QO_MG_COUNT };
/* fp32 unit in the last place at magnitude |x| (>= 2^-126) */
static double ulp32(double x) { x = fabs(x); if (x < 1.17549435e-38) x = 1.17549435e-38; int ex; frexp(x, &ex); return ldexp(1.0, ex - 24); }
struct qo_env;
static void mg_note(struct qo_env *e, int cls, double value, double threshold, double scale);

typedef struct qo_env {
    qs_config c;
    qo_scen sc;
    int K;
    uint32_t gid;
    qo_drone d[QS_MAX_AGENTS];
    int tick;
    int svd_ctr;
    uint32_t step_ctr;
    double snap_vel[QS_MAX_AGENTS][3];  /* self.vel: refreshed in step (quadrotor_multi.py:705-709), stale at reset (:477) */
    double obst_xy[QS_MAX_OBSTACLES][2];
    int n_obst;
    int scenario_now;                   /* QS_SCENARIO_O_RANDOM / O_STATIC_SAME_GOAL / STATIC_SAME_GOAL */
    double evader[2];                   /* fork mode: Scenario_dynamic_repulsive.pos */
    int episode_success;                /* fork mode: quadrotor_multi_rewards.py:757,625-627 */
    double snap_heading[QS_MAX_AGENTS]; /* fork mode: self.heading, refreshed in step only (stale in the reset-time neighbour obs) */
    double cam_n1[QS_MAX_AGENTS], cam_n2[QS_MAX_AGENTS];   /* camera pixel-noise normals of the current neighbour pass */
    int chasers_placed;                 /* fork mode: 0 until the first reset has given the dynamics a position; the evader's very
                                           first step ignores the chasers (`hasattr(env.dynamics, "pos")`, dynamic_repulsive.py:44) */
    double approach_metric;
    /* per-episode counters (quadrotor_multi.py:153-171) */
    int collisions_per_episode, collisions_after_settle, collisions_final_5s;
    int col_room, col_floor, col_wall, col_ceiling;
    int obst_col_per_episode, obst_col_after_settle;
    qs_stats stats;
    /* record of the last finished episode: infos[i]['episode_extra_stats'] (QS_ER_* layout of include/quadsim.h) */
    int32_t ep_rec[QS_ER_COUNT];
    double ep_agent[QS_MAX_AGENTS][4];
    /* last-step diagnostics for tests */
    uint32_t last_new_pairs[QS_MAX_AGENTS];
    int32_t last_neighbors[QS_MAX_AGENTS][QS_MAX_AGENTS];
    int last_impulse_flag;
    /* distance of every threshold decision of the last step from its threshold, in fp32 ulps of the operands' magnitude (min per class):
     * the parity tests accept a discrete disagreement with the fp32 kernel only where this proves a tie */
    double margin[QO_MG_COUNT];
    /* infos[i]["rewards"] raw terms of the last step (QS_RI_* order; fork mode: goal_dist in slot 0) */
    double rew_info[QS_MAX_AGENTS][QS_RI_COUNT];
    /* tape */
    const double *tn, *tu, *tc;
    int nn, nu, nc, in_, iu, ic;
    int use_tape;
} qo_env;

static void mg_note(struct qo_env *e, int cls, double value, double threshold, double scale)
{
    double m = fabs(value - threshold) / ulp32(fmax(fabs(scale), fabs(threshold)));
    if (m < e->margin[cls]) e->margin[cls] = m;
}
static void mg_clear(struct qo_env *e) { for (int k = 0; k < QO_MG_COUNT; ++k) e->margin[k] = 1e300; }

/* ------------------------------------------------------------------------------------------------ */
/* Philox4x32-10 + unit transforms (DESIGN.md "RNG contract")                                         */
/* ------------------------------------------------------------------------------------------------ */
static void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                          uint32_t out[4])
{
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static double u23(uint32_t x) { return ((double)(x >> 9) + 0.5) * (1.0 / 8388608.0); }

static void philox_block(const qo_env *e, int site, int drone, int aux, int block, uint32_t out[4])
{
    uint32_t c2 = (uint32_t)site | ((uint32_t)drone << 8) | ((uint32_t)aux << 16);
    philox4x32_10(e->gid, e->step_ctr, c2, (uint32_t)block, (uint32_t)e->c.seed, (uint32_t)(e->c.seed >> 32), out);
}

/* idx-th uniform in (0,1) of stream (site, drone, aux) */
static double rnd_u(qo_env *e, int site, int drone, int aux, int idx)
{
    if (e->use_tape) { if (e->iu >= e->nu) { e->iu++; return 0.5; } return e->tu[e->iu++]; }
    uint32_t r[4];
    philox_block(e, site, drone, aux, idx >> 2, r);
    return u23(r[idx & 3]);
}

/* idx-th standard normal of stream (site, drone, aux): Box-Muller on lanes (0,1) and (2,3) of a block */
static double rnd_n(qo_env *e, int site, int drone, int aux, int idx)
{
    if (e->use_tape) { if (e->in_ >= e->nn) { e->in_++; return 0.0; } return e->tn[e->in_++]; }
    uint32_t r[4];
    philox_block(e, site, drone, aux, idx >> 2, r);
    int p = (idx & 3) >> 1;
    double u1 = u23(r[2 * p]), u2 = u23(r[2 * p + 1]);
    double rad = sqrt(-2.0 * log(u1)), ang = 2.0 * QO_PI * u2;
    return (idx & 1) ? rad * sin(ang) : rad * cos(ang);
}

/* draws the reference makes but whose value never reaches an output: consumed only when replaying a tape */
static void burn_n(qo_env *e, int n) { if (e->use_tape) e->in_ += n; }
static void burn_u(qo_env *e, int n) { if (e->use_tape) e->iu += n; }

/* k distinct ids out of n.  tape: the reference's np.random.choice(..., replace=False) results;
 * philox: partial Fisher-Yates, draw t swaps slot t with slot t + floor(u * (n - t)). */
static void rnd_choice(qo_env *e, int site, int aux, int n, int k, int *out)
{
    if (e->use_tape) {
        for (int t = 0; t < k; ++t) out[t] = (e->ic < e->nc) ? (int)e->tc[e->ic] : t, e->ic++;
        return;
    }
    int a[256];
    for (int i = 0; i < n; ++i) a[i] = i;
    for (int t = 0; t < k; ++t) {
        double u = rnd_u(e, site, 0xFF, aux, t);
        int r = t + (int)floor(u * (double)(n - t));
        if (r > n - 1) r = n - 1;
        int tmp = a[t]; a[t] = a[r]; a[r] = tmp;
        out[t] = a[t];
    }
}

/* ------------------------------------------------------------------------------------------------ */
/* small math                                                                                         */
/* ------------------------------------------------------------------------------------------------ */
static double clampd(double x, double lo, double hi) { return fmin(fmax(x, lo), hi); }
static double norm3(const double *v) { return sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); }
static double dot3(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static void matvec3(const double *R, const double *v, double *o)
{
    for (int i = 0; i < 3; ++i) o[i] = R[3 * i] * v[0] + R[3 * i + 1] * v[1] + R[3 * i + 2] * v[2];
}
static void matmul3(const double *A, const double *B, double *C)
{
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}
static void yaw_rot(double theta, double *R)
{
    double c = cos(theta), s = sin(theta);
    R[0] = c; R[1] = -s; R[2] = 0; R[3] = s; R[4] = c; R[5] = 0; R[6] = 0; R[7] = 0; R[8] = 1;
}

/* U V^T of the SVD of R == orthogonal polar factor of R (quadrotor_dynamics.py:556-557).  Newton iteration
 * X <- (X + X^-T)/2 converges quadratically to it for det(R) > 0; 12 iterations are exact in float64 for the
 * near-orthonormal matrices this is called on. */
static void polar_orthonormalise(double *R)
{
    double X[9];
    memcpy(X, R, sizeof(X));
    for (int it = 0; it < 12; ++it) {
        double c00 = X[4] * X[8] - X[5] * X[7], c01 = X[5] * X[6] - X[3] * X[8], c02 = X[3] * X[7] - X[4] * X[6];
        double c10 = X[2] * X[7] - X[1] * X[8], c11 = X[0] * X[8] - X[2] * X[6], c12 = X[1] * X[6] - X[0] * X[7];
        double c20 = X[1] * X[5] - X[2] * X[4], c21 = X[2] * X[3] - X[0] * X[5], c22 = X[0] * X[4] - X[1] * X[3];
        double det = X[0] * c00 + X[1] * c01 + X[2] * c02;
        double id = 1.0 / det;
        /* X^-T = cofactor / det */
        double T[9] = { c00 * id, c01 * id, c02 * id, c10 * id, c11 * id, c12 * id, c20 * id, c21 * id, c22 * id };
        for (int i = 0; i < 9; ++i) X[i] = 0.5 * (X[i] + T[i]);
    }
    memcpy(R, X, sizeof(X));
}

/* ------------------------------------------------------------------------------------------------ */
/* a1-a4: QuadrotorDynamics.step                                                                      */
/* ------------------------------------------------------------------------------------------------ */

/* OUNoiseNumba.noise, numba_utils.py:101-105 (jitclass fields are float32: theta, sigma) */
static void ou_noise(qo_env *e, int i)
{
    qo_drone *q = &e->d[i];
    double theta = (double)(float)e->c.ou_theta, sigma = (double)(float)e->c.ou_sigma;
    for (int m = 0; m < 4; ++m) {
        double x = q->ou[m];
        double dx = theta * (0.0 - x) + sigma * rnd_n(e, SITE_OU, i, 0, m);
        q->ou[m] = x + dx;
    }
}

/* one physics sub-step: step1_numba, quadrotor_dynamics.py:355-390 */
static void dynamics_substep(qo_env *e, int i, const double *cmd_in, int substep, int do_svd)
{
    const qs_config *c = &e->c;
    qo_drone *q = &e->d[i];
    double dt = c->dt;

    /* --- calculate_torque_integrate_rotations_and_update_omega, quadrotor_dynamics.py:504-573 --- */
    double cmd[4], thrusts[4], torque[3] = { 0, 0, 0 }, thrust_sum = 0;
    for (int m = 0; m < 4; ++m) {
        cmd[m] = clampd(cmd_in[m], 0.0, 1.0);                                         /* :511 */
        double tau = (cmd[m] < q->cmds_damp[m]) ? c->motor_tau_down : c->motor_tau_up; /* :512-513 */
        if (tau > 1.0) tau = 1.0;                                                     /* :514 */
        double thrust_rot = sqrt(cmd[m]);                                             /* :517 */
        q->rot_damp[m] = tau * (thrust_rot - q->rot_damp[m]) + q->rot_damp[m];        /* :518 */
        q->cmds_damp[m] = q->rot_damp[m] * q->rot_damp[m];                            /* :519 */
        q->cmds_damp[m] = clampd(q->cmds_damp[m] + cmd[m] * q->ou[m], 0.0, 1.0);      /* :522-523 */
        double w = q->cmds_damp[m];
        thrusts[m] = c->thrust_max[m] * ((1.0 - c->motor_linearity) * w * w + c->motor_linearity * w); /* :524 */
        torque[0] += c->prop_cross[m][0] * thrusts[m];                                /* :527,533 */
        torque[1] += c->prop_cross[m][1] * thrusts[m];
        torque[2] += c->prop_cross[m][2] * thrusts[m] + c->torque_max[m] * c->prop_ccw[m] * w; /* :530 */
        thrust_sum += thrusts[m];                                                     /* :540 */
    }
    /* rotational dynamics, :544-551 */
    double wv[3];
    matvec3(q->rot, q->omega, wv);
    double wn = norm3(wv);
    if (wn != 0.0) {
        double kx = wv[0] / wn, ky = wv[1] / wn, kz = wv[2] / wn;
        double Kk[9] = { 0, -kz, ky, kz, 0, -kx, -ky, kx, 0 }, K2[9], dR[9], Rn[9];
        matmul3(Kk, Kk, K2);
        double ang = wn * dt, s = sin(ang), cc = 1.0 - cos(ang);
        for (int a = 0; a < 9; ++a) dR[a] = ((a % 4 == 0) ? 1.0 : 0.0) + s * Kk[a] + cc * K2[a];
        matmul3(dR, q->rot, Rn);
        memcpy(q->rot, Rn, sizeof(Rn));
    }
    /* :554-558 -- `since_last_svd += dt; if > 0.5: svd`.  All drones of an env step in lockstep from construction,
     * so the env keeps ONE integer sub-step counter; svd_period (host-computed by literally accumulating dt in
     * float64) is the number of sub-steps after which the float accumulator first exceeds the limit. */
    if (do_svd) polar_orthonormalise(q->rot);
    /* omega update, :562-567 */
    double Iw[3] = { c->inertia[0] * q->omega[0], c->inertia[1] * q->omega[1], c->inertia[2] * q->omega[2] };
    double nw[3] = { -q->omega[0], -q->omega[1], -q->omega[2] };
    double cr[3] = { nw[1] * Iw[2] - nw[2] * Iw[1], nw[2] * Iw[0] - nw[0] * Iw[2], nw[0] * Iw[1] - nw[1] * Iw[0] };
    for (int a = 0; a < 3; ++a) {
        double wdot = (1.0 / c->inertia[a]) * (cr[a] + torque[a]);
        double dq = clampd(c->damp_omega_quadratic * q->omega[a] * q->omega[a], 0.0, 1.0);
        q->omega[a] = clampd(q->omega[a] + (1.0 - dq) * dt * wdot, -c->omega_max, c->omega_max);
    }
    /* :570 */
    for (int a = 0; a < 3; ++a) q->pos[a] += dt * q->vel[a];

    /* --- room clip, quadrotor_dynamics.py:367-374 --- */
    double hx = c->room_dims[0] / 2, hy = c->room_dims[1] / 2, hz = c->room_dims[2];
    double bx = q->pos[0], by = q->pos[1], bz = q->pos[2];
    q->pos[0] = clampd(bx, -hx, hx); q->pos[1] = clampd(by, -hy, hy); q->pos[2] = clampd(bz, 0.0, hz);
    q->crashed_wall = !(bx == q->pos[0] && by == q->pos[1]);
    q->crashed_ceiling = bz > q->pos[2];
    mg_note(e, QO_MG_WALL, fabs(bx), hx, hx); mg_note(e, QO_MG_WALL, fabs(by), hy, hy); mg_note(e, QO_MG_CEIL, bz, hz, hz);
    mg_note(e, QO_MG_FLOOR, q->pos[2], c->arm, fmax(fabs(q->pos[2]), fabs(dt * q->vel[2])));

    /* --- floor_interaction_numba, quadrotor_dynamics.py:576-646 (floor_threshold = arm, :385) --- */
    double thr[3] = { 0, 0, thrust_sum }, force[3];
    q->crashed_floor = 0;
    if (q->pos[2] <= c->arm) {
        q->pos[2] = c->arm;
        matvec3(q->rot, thr, force);
        if (q->on_floor) {
            mg_note(e, QO_MG_LIFTOFF, force[2] / c->mass, QO_GRAV, QO_GRAV);    /* acc_z = max(0, .): whether a resting drone lifts off */
            double theta = atan2(q->rot[3], q->rot[0] + QO_EPS_DYN);
            yaw_rot(theta, q->rot);
            double fr = c->floor_mu * (c->mass * QO_GRAV - force[2]);
            if (norm3(q->vel) < QO_EPS_DYN) {
                double fm = sqrt(force[0] * force[0] + force[1] * force[1]);
                fm = fmax(fm - fr, 0.0);
                if (fm == 0.0) { force[0] = 0; force[1] = 0; }
                else { double fa = atan2(force[1], force[0]); force[0] = fm * cos(fa); force[1] = fm * sin(fa); }
            } else {
                double fa = atan2(q->vel[1], q->vel[0]);     /* :608 (numba path: friction opposes velocity) */
                force[0] -= cos(fa) * fr; force[1] -= sin(fa) * fr;
            }
        } else {
            q->on_floor = 1; q->crashed_floor = 1;
            for (int a = 0; a < 3; ++a) { q->vel[a] = 0; q->omega[a] = 0; }
            double theta = atan2(q->rot[3], q->rot[0] + QO_EPS_DYN);
            if (q->rot[8] < 0) theta = -QO_PI + 2.0 * QO_PI * rnd_u(e, SITE_FLOOR_YAW, i, substep, 0); /* :623-626 */
            yaw_rot(theta, q->rot);
            for (int m = 0; m < 4; ++m) { q->cmds_damp[m] = 0; q->rot_damp[m] = 0; }
        }
        q->acc[0] = force[0] / c->mass; q->acc[1] = force[1] / c->mass;
        q->acc[2] = fmax(0.0, -QO_GRAV + force[2] / c->mass);
    } else {
        q->on_floor = 0;
        matvec3(q->rot, thr, force);
        q->acc[0] = force[0] / c->mass; q->acc[1] = force[1] / c->mass; q->acc[2] = -QO_GRAV + force[2] / c->mass;
    }
    /* --- compute_velocity_and_acceleration, quadrotor_dynamics.py:649-656 --- */
    for (int a = 0; a < 3; ++a) q->vel[a] = (1.0 - c->vel_damp) * q->vel[a] + dt * q->acc[a];
    double ag[3] = { q->acc[0], q->acc[1], q->acc[2] + c->gravity };
    for (int a = 0; a < 3; ++a) q->accm[a] = q->rot[a] * ag[0] + q->rot[3 + a] * ag[1] + q->rot[6 + a] * ag[2];
}

/* QuadrotorDynamics.step (quadrotor_dynamics.py:215-221) behind RawControl.step (quadrotor_control.py:53-57) */
static void drone_control_step(qo_env *e, int i, const double *action, const int *svd_fire)
{
    double cmd[4];
    for (int m = 0; m < 4; ++m) cmd[m] = 0.5 * (clampd(action[m], -1.0, 1.0) + 1.0);
    ou_noise(e, i);
    for (int s = 0; s < e->c.sim_steps; ++s) dynamics_substep(e, i, cmd, s, svd_fire[s]);
}

/* compute_reward_weighted, quadrotor_single.py:34-92 (dt = physics dt, :362-364) */
static double base_reward(qo_env *e, int i, const double *action, double *rewraw_pos)
{
    const qs_config *c = &e->c;
    const qo_drone *q = &e->d[i];
    double dp[3] = { q->goal[0] - q->pos[0], q->goal[1] - q->pos[1], q->goal[2] - q->pos[2] };
    double dist = norm3(dp);
    double effort = sqrt(action[0] * action[0] + action[1] * action[1] + action[2] * action[2] + action[3] * action[3]);
    double orient = q->on_floor ? 1.0 : -q->rot[8];
    double spin = sqrt(q->omega[0] * q->omega[0] + q->omega[1] * q->omega[1] + q->omega[2] * q->omega[2]);
    double crash = q->on_floor ? 1.0 : 0.0;
    *rewraw_pos = c->dt * (-dist);
    double *ri = e->rew_info[i];                                  /* rew_info entries are dt * (-raw cost), quadrotor_single.py:69-84 */
    ri[QS_RI_RAW_POS] = c->dt * -dist; ri[QS_RI_RAW_ACTION] = c->dt * -effort; ri[QS_RI_RAW_CRASH] = c->dt * -crash;
    ri[QS_RI_RAW_ORIENT] = c->dt * -orient; ri[QS_RI_RAW_SPIN] = c->dt * -spin;
    return -c->dt * (c->rew_pos * dist + c->rew_effort * effort + c->rew_crash * crash + c->rew_orient * orient +
                     c->rew_spin * spin);
}

/* ------------------------------------------------------------------------------------------------ */
/* a9: self observation                                                                               */
/* ------------------------------------------------------------------------------------------------ */

/* rot2quat, sensor_noise.py:34-63 followed by quat2R, quad_utils.py:162-168 (theta noise is 0 so the
 * small-angle quaternion is identity; sensor_noise.py:205-210) */
static void rot_quat_roundtrip(const double *r, double *o)
{
    double tr = r[0] + r[4] + r[8], qw, qx, qy, qz, S;
    if (tr > 0) { S = sqrt(tr + 1.0) * 2; qw = 0.25 * S; qx = (r[7] - r[5]) / S; qy = (r[2] - r[6]) / S; qz = (r[3] - r[1]) / S; }
    else if (r[0] > r[4] && r[0] > r[8]) { S = sqrt(1.0 + r[0] - r[4] - r[8]) * 2; qw = (r[7] - r[5]) / S; qx = 0.25 * S; qy = (r[1] + r[3]) / S; qz = (r[2] + r[6]) / S; }
    else if (r[4] > r[8]) { S = sqrt(1.0 + r[4] - r[0] - r[8]) * 2; qw = (r[2] - r[6]) / S; qx = (r[1] + r[3]) / S; qy = 0.25 * S; qz = (r[5] + r[7]) / S; }
    else { S = sqrt(1.0 + r[8] - r[0] - r[4]) * 2; qw = (r[3] - r[1]) / S; qx = (r[2] + r[6]) / S; qy = (r[5] + r[7]) / S; qz = 0.25 * S; }
    o[0] = 1.0 - 2 * qy * qy - 2 * qz * qz; o[1] = 2 * qx * qy - 2 * qz * qw; o[2] = 2 * qx * qz + 2 * qy * qw;
    o[3] = 2 * qx * qy + 2 * qz * qw; o[4] = 1.0 - 2 * qx * qx - 2 * qz * qz; o[5] = 2 * qy * qz - 2 * qx * qw;
    o[6] = 2 * qx * qz - 2 * qy * qw; o[7] = 2 * qy * qz + 2 * qx * qw; o[8] = 1.0 - 2 * qx * qx - 2 * qy * qy;
}

static int self_obs_dim(const qs_config *c)
{
    return c->obs_repr == QS_OBS_XYZ_VXYZ_R_OMEGA_FLOOR ? 19 : (c->obs_repr == QS_OBS_XYZ_VXYZ_R_OMEGA_WALL ? 24 : 18);
}

/* state_xyz_vxyz_R_omega{,_floor,_wall} (get_state.py:226-292) with SensorNoise.add_noise_numba
 * (sensor_noise.py:172-218, 235-261).  Draw order of the reference: pos 3N 3U, vel 3N 3U, omega 3N,
 * theta 3N 3U, acc 3N 3N. */
static void self_obs(qo_env *e, int i, int site, double *o)
{
    const qs_config *c = &e->c;
    const qo_drone *q = &e->d[i];
    double p[3], v[3], w[3], R[9];
    if (c->sense_noise == QS_SENSE_NOISE_NONE) {
        memcpy(p, q->pos, sizeof(p)); memcpy(v, q->vel, sizeof(v)); memcpy(w, q->omega, sizeof(w));
        memcpy(R, q->rot, sizeof(R));
    } else {
        for (int a = 0; a < 3; ++a) p[a] = q->pos[a] + c->sense_pos_std * rnd_n(e, site, i, 0, a);
        burn_u(e, 3);
        for (int a = 0; a < 3; ++a) v[a] = q->vel[a] + c->sense_vel_std * rnd_n(e, site, i, 0, 3 + a);
        burn_u(e, 3);
        for (int a = 0; a < 3; ++a) w[a] = q->omega[a] + c->sense_gyro_std * rnd_n(e, site, i, 0, 6 + a);
        burn_n(e, 3); burn_u(e, 3); burn_n(e, 6);
        rot_quat_roundtrip(q->rot, R);
    }
    for (int a = 0; a < 3; ++a) { o[a] = p[a] - q->goal[a]; o[3 + a] = v[a]; o[15 + a] = w[a]; }
    for (int a = 0; a < 9; ++a) o[6 + a] = R[a];
    if (c->obs_repr == QS_OBS_XYZ_VXYZ_R_OMEGA_FLOOR) o[18] = p[2];
    if (c->obs_repr == QS_OBS_XYZ_VXYZ_R_OMEGA_WALL) {
        double lo[3] = { -c->room_dims[0] / 2, -c->room_dims[1] / 2, 0 }, hi[3] = { c->room_dims[0] / 2, c->room_dims[1] / 2, c->room_dims[2] };
        for (int a = 0; a < 3; ++a) { o[18 + a] = clampd(p[a] - lo[a], 0.0, 5.0); o[21 + a] = clampd(hi[a] - p[a], 0.0, 5.0); }
    }
}

/* ------------------------------------------------------------------------------------------------ */
/* a15: neighbour observations, quadrotor_multi.py:275-380                                            */
/* ------------------------------------------------------------------------------------------------ */
static void neighbor_obs(qo_env *e, int i, double *o)
{
    const qs_config *c = &e->c;
    int K = e->K, V = c->neighbor_visible_num;
    if (c->neighbor_obs_type != QS_NEIGHBOR_POS_VEL || V <= 0) return;
    int cand[QS_MAX_AGENTS], n = 0;
    double rel[QS_MAX_AGENTS][6], metric[QS_MAX_AGENTS];
    for (int j = 0; j < K; ++j) {
        if (j == i) continue;
        for (int a = 0; a < 3; ++a) { rel[n][a] = e->d[j].pos[a] - e->d[i].pos[a]; rel[n][3 + a] = e->snap_vel[j][a] - e->snap_vel[i][a]; }
        double s = 0; for (int a = 0; a < 6; ++a) s += rel[n][a] * rel[n][a];
        metric[n] = fmax(sqrt(s), 0.01);                         /* :357-359: norm of the 6-vector */
        cand[n++] = j;
    }
    int order[QS_MAX_AGENTS];
    for (int a = 0; a < n; ++a) order[a] = a;
    if (V < K - 1) {                                             /* :352-371 argsort (ties: lowest index first) */
        for (int a = 1; a < n; ++a) { int x = order[a], b = a - 1; while (b >= 0 && metric[order[b]] > metric[x]) { order[b + 1] = order[b]; --b; } order[b + 1] = x; }
    }
    if (V < K - 1) {                                             /* a tie = two candidates whose metrics are within fp32 round-off, one of them chosen */
        for (int a = 0; a < V && a + 1 < n; ++a) mg_note(e, QO_MG_RANK, metric[order[a]], metric[order[a + 1]], metric[order[a + 1]]);
    }
    double lim_p[3] = { c->room_dims[0], c->room_dims[1], c->room_dims[2] };   /* room_range, quadrotor_single.py:279,295 */
    double lim_v = 2.0 * 3.0;                                                  /* 2 * vxyz_max, quadrotor_single.py:296 */
    for (int s = 0; s < V && s < n; ++s) {
        int a = order[s];
        e->last_neighbors[i][s] = cand[a];
        for (int k = 0; k < 3; ++k) {
            /* clip in float32 space bounds like the Box (quadrotor_multi.py:121-125,337-339) */
            o[6 * s + k] = clampd(rel[a][k], -(double)(float)lim_p[k], (double)(float)lim_p[k]);
            o[6 * s + 3 + k] = clampd(rel[a][3 + k], -lim_v, lim_v);
        }
    }
}

/* a14: get_surround_sdfs, obstacles/utils.py:5-27 */
static void sdf_obs(const qo_env *e, int i, double *o)
{
    const qs_config *c = &e->c;
    double res = c->sdf_resolution, rad = c->obst_size / 2.0;
    double gx[3] = { e->d[i].pos[0] - res, e->d[i].pos[0], e->d[i].pos[0] + res };
    double gy[3] = { e->d[i].pos[1] - res, e->d[i].pos[1], e->d[i].pos[1] + res };
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) {
            double md = 100.0;
            for (int m = 0; m < e->n_obst; ++m) {
                double dx = gx[a] - e->obst_xy[m][0], dy = gy[b] - e->obst_xy[m][1];
                double dd = sqrt(dx * dx + dy * dy);
                if (dd < md) md = dd;
            }
            o[a * 3 + b] = md - rad;
        }
}

static int obs_dim(const qs_config *c)
{
    int D = self_obs_dim(c);
    if (c->neighbor_obs_type == QS_NEIGHBOR_POS_VEL) D += 6 * c->neighbor_visible_num;
    if (c->use_obstacles) D += 9;
    return D;
}

/* ------------------------------------------------------------------------------------------------ */
/* a12-a16: impulses                                                                                  */
/* ------------------------------------------------------------------------------------------------ */

/* compute_new_vel, collisions/utils.py:8-19 */
static void compute_new_vel(double max_mag, double *vel, const double *shift, double decay)
{
    double vn[3] = { vel[0] + shift[0], vel[1] + shift[1], vel[2] + shift[2] };
    double mag = norm3(vn), den = (mag == 0.0) ? mag + QO_EPS_COL : mag;
    double dir[3] = { vn[0] / den, vn[1] / den, vn[2] / den };
    mag = fmin(mag * decay, max_mag);
    for (int a = 0; a < 3; ++a) { double nv = dir[a] * mag; double sh = nv - vel[a]; vel[a] += sh; }
}

/* compute_new_omega, collisions/utils.py:22-33: 3 U(-1,1) then 1 U(max/2, max) */
static void compute_new_omega(const double *u4, double magn_scale, double *out)
{
    double omax = magn_scale * QO_PI;
    double w[3] = { -1.0 + 2.0 * u4[0], -1.0 + 2.0 * u4[1], -1.0 + 2.0 * u4[2] };
    double mag = norm3(w), den = (mag == 0.0) ? mag + QO_EPS_COL : mag;
    double m2 = omax / 2 + (omax - omax / 2) * u4[3];
    for (int a = 0; a < 3; ++a) out[a] = w[a] / den * m2;
}

/* perform_collision_between_drones, collisions/quadrotors.py:9-59 */
static void pair_impulse(qo_env *e, int i, int j)
{
    qo_drone *p = &e->d[i], *q = &e->d[j];
    double n[3] = { p->pos[0] - q->pos[0], p->pos[1] - q->pos[1], p->pos[2] - q->pos[2] };
    double nm = norm3(n), den = (nm == 0.0) ? nm + QO_EPS_COL : nm;
    for (int a = 0; a < 3; ++a) n[a] /= den;
    double v1n = dot3(p->vel, n), v2n = dot3(q->vel, n);
    double vc[3] = { (v2n - v1n) * n[0], (v2n - v1n) * n[1], (v2n - v1n) * n[2] };
    double s1[3] = { vc[0], vc[1], vc[2] }, s2[3] = { -vc[0], -vc[1], -vc[2] };
    for (int att = 0; att < 3; ++att) {
        double cons[3], n1[3], n2[3];
        for (int a = 0; a < 3; ++a) cons[a] = 0.8 * rnd_n(e, SITE_PAIR, i, j, att * 12 + a);
        for (int a = 0; a < 3; ++a) n1[a] = cons[a] + 0.15 * rnd_n(e, SITE_PAIR, i, j, att * 12 + 3 + a);
        for (int a = 0; a < 3; ++a) n2[a] = -cons[a] + 0.15 * rnd_n(e, SITE_PAIR, i, j, att * 12 + 6 + a);
        for (int a = 0; a < 3; ++a) { s1[a] = vc[a] + n1[a]; s2[a] = -vc[a] + n2[a]; }
        double t1[3] = { p->vel[0] + s1[0], p->vel[1] + s1[1], p->vel[2] + s1[2] };
        double t2[3] = { q->vel[0] + s2[0], q->vel[1] + s2[1], q->vel[2] + s2[2] };
        if (dot3(t1, n) > 0 && 0 > dot3(t2, n)) break;
    }
    double maxv = fmax(norm3(p->vel), norm3(q->vel));
    double d1 = 0.2 + 0.6 * rnd_u(e, SITE_PAIR, i, j, 36);
    compute_new_vel(maxv, p->vel, s1, d1);
    double d2 = 0.2 + 0.6 * rnd_u(e, SITE_PAIR, i, j, 37);
    compute_new_vel(maxv, q->vel, s2, d2);
    double u4[4], w[3];
    for (int a = 0; a < 4; ++a) u4[a] = rnd_u(e, SITE_PAIR, i, j, 38 + a);
    compute_new_omega(u4, 20.0, w);
    for (int a = 0; a < 3; ++a) { p->omega[a] += w[a]; q->omega[a] -= w[a]; }
}

/* perform_collision_with_obstacle, collisions/obstacles.py:9-50 */
/* compute_col_norm_and_new_vel_obst, collisions/obstacles.py:8-21: unit horizontal normal from the obstacle axis to the drone and
 * the velocity component along it (the reference's own unit test holds one known answer for it,
 * collisions/test/unit_test/obstacles.py:6-18) */
static double col_norm_and_new_vel_obst(const double *pos, const double *vel, const double *op, double *n)
{
    n[0] = pos[0] - op[0]; n[1] = pos[1] - op[1]; n[2] = 0.0;
    double nm = norm3(n), den = (nm == 0.0) ? nm + QO_EPS_COL : nm;
    for (int a = 0; a < 3; ++a) n[a] /= den;
    return dot3(vel, n);
}

static void obstacle_impulse(qo_env *e, int i, int m)
{
    const qs_config *c = &e->c;
    qo_drone *q = &e->d[i];
    double op[3] = { e->obst_xy[m][0], e->obst_xy[m][1], c->room_dims[2] / 2.0 };
    double n[3];
    col_norm_and_new_vel_obst(q->pos, q->vel, op, n);
    double vm = norm3(q->vel);
    double nv[3] = { vm * n[0], vm * n[1], vm * n[2] }, noise[3] = { 0, 0, 0 };
    for (int att = 0; att < 3; ++att) {
        double t[3], s[3];
        for (int a = 0; a < 3; ++a) t[a] = 0.1 * rnd_n(e, SITE_OBST, i, 0, att * 8 + a);
        for (int a = 0; a < 3; ++a) t[a] += 0.05 * rnd_n(e, SITE_OBST, i, 0, att * 8 + 3 + a);
        for (int a = 0; a < 3; ++a) s[a] = nv[a] + t[a];
        if (dot3(s, n) > 0) { memcpy(noise, t, sizeof(t)); break; }
    }
    double dp[3] = { q->pos[0] - op[0], q->pos[1] - op[1], q->pos[2] - op[2] };
    double shift[3] = { nv[0] - q->vel[0] + noise[0], nv[1] - q->vel[1] + noise[1], nv[2] - q->vel[2] + noise[2] };
    double u = rnd_u(e, SITE_OBST, i, 0, 24);
    double decay = (norm3(dp) < c->obst_size / 2) ? 1.0 : 0.2 + 0.6 * u;   /* :41-46 (3-D distance to the mid-height centre) */
    compute_new_vel(vm, q->vel, shift, decay);
    double u4[4], w[3];
    for (int a = 0; a < 4; ++a) u4[a] = rnd_u(e, SITE_OBST, i, 0, 25 + a);
    compute_new_omega(u4, 1.0, w);
    for (int a = 0; a < 3; ++a) q->omega[a] += w[a];
}

/* perform_collision_with_wall / _ceiling, collisions/room.py:6-45, 91-113 */
static void room_impulse(qo_env *e, int i, int is_wall)
{
    const qs_config *c = &e->c;
    qo_drone *q = &e->d[i];
    int site = is_wall ? SITE_WALL : SITE_CEILING, k = 0;
    double sp = norm3(q->vel);
    double real = 0.2 * sp + (0.8 * sp - 0.2 * sp) * rnd_u(e, site, i, 0, k++);
    real = clampd(real, 0.1, 6.0);
    double dir[3];
    for (int a = 0; a < 3; ++a) dir[a] = -1.0 + 2.0 * rnd_u(e, site, i, 0, k++);
    if (is_wall) {
        double hx = c->room_dims[0] / 2, hy = c->room_dims[1] / 2;
        /* philox streams use fixed slots 4 (x override) and 5 (y override); a tape just pops in order */
        if (q->pos[0] == -hx) dir[0] = 0.1 + 0.9 * rnd_u(e, site, i, 0, 4);
        else if (q->pos[0] == hx) dir[0] = -1.0 + 0.9 * rnd_u(e, site, i, 0, 4);
        if (q->pos[1] == -hy) dir[1] = 0.1 + 0.9 * rnd_u(e, site, i, 0, 5);
        else if (q->pos[1] == hy) dir[1] = -1.0 + 0.9 * rnd_u(e, site, i, 0, 5);
        k = 6;
    }
    dir[2] = -1.0 + 0.5 * rnd_u(e, site, i, 0, k++);
    double dm = norm3(dir);
    for (int a = 0; a < 3; ++a) q->vel[a] = real * (dir[a] / (dm + 1e-5));
    double w[3];
    for (int a = 0; a < 3; ++a) w[a] = -1.0 + 2.0 * rnd_u(e, site, i, 0, k++);
    double wm = norm3(w) + 1e-5;
    double omax = 20 * QO_PI, mag = omax / 2 + (omax - omax / 2) * rnd_u(e, site, i, 0, k++);
    for (int a = 0; a < 3; ++a) q->omega[a] += w[a] / wm * mag;
}

/* perform_downwash, aerodynamics/downwash.py:4-66 */
static int downwash(qo_env *e)
{
    int K = e->K, any = 0;
    double dt = e->c.dt * e->c.sim_steps;
    double P[QS_MAX_AGENTS][3], Z[QS_MAX_AGENTS][3];
    for (int i = 0; i < K; ++i) for (int a = 0; a < 3; ++a) { P[i][a] = e->d[i].pos[a]; Z[i][a] = e->d[i].rot[3 * a + 2]; }
    for (int i = 0; i < K; ++i) {
        double ua = rnd_u(e, SITE_DOWNWASH, i, 0xFF, 0), uw = rnd_u(e, SITE_DOWNWASH, i, 0xFF, 1);
        for (int j = 0; j < K; ++j) {
            if (j == i) continue;
            double r[3] = { P[j][0] - P[i][0], P[j][1] - P[i][1], P[j][2] - P[i][2] };
            double dist = norm3(r);
            double acc = fmax(1e-6, (6.0 / 17.0) * (-10.0 * dist + 7.0) + (-0.1 + 0.2 * ua));
            double ow = fmax(1e-6, 0.3 * (dist - 1.0) * (dist - 1.0) + (-0.01 + 0.02 * uw));
            double rz = dot3(r, Z[i]);
            double rxy = sqrt(dist * dist - rz * rz);
            if (-0.7 < rz && rz < 0 && rxy < 0.1) {
                double nz[3], dw[3];
                for (int a = 0; a < 3; ++a) nz[a] = Z[i][a] + (-0.1 + 0.2 * rnd_u(e, SITE_DOWNWASH, i, j, a));
                double nm = norm3(nz), den = (nm == 0.0) ? nm + 1e-6 : nm;
                for (int a = 0; a < 3; ++a) dw[a] = -1.0 + 2.0 * rnd_u(e, SITE_DOWNWASH, i, j, 3 + a);
                double wm = norm3(dw), wden = (wm == 0.0) ? wm + 1e-6 : wm;
                for (int a = 0; a < 3; ++a) {
                    e->d[j].vel[a] += acc * (-nz[a] / den) * dt;
                    e->d[j].omega[a] += ow * (dw[a] / wden) * dt;
                }
                any = 1;
            }
        }
    }
    return any;
}

/* ------------------------------------------------------------------------------------------------ */
/* a17: reset                                                                                         */
/* ------------------------------------------------------------------------------------------------ */

/* get_cell_centers, obstacles/utils.py:47-58 : index = i * W + jj, y descending in jj */
static void cell_center(const qs_config *c, int index, double *xy)
{
    int L = c->obst_area_len, W = c->obst_area_wid;
    int i = index / W, jj = index % W, j = W - 1 - jj;
    xy[0] = i + 0.5 - (double)(L / 2);
    xy[1] = j + 0.5 - (double)(W / 2);
}

/* obst_generation_given_density (quadrotor_multi.py:405-426) + Scenario_o_X.reset (scenarios/obstacles/o_X.py) */
static void obstacle_scenario_reset(qo_env *e)
{
    const qs_config *c = &e->c;
    int L = c->obst_area_len, W = c->obst_area_wid, M = c->num_obstacles, K = e->K;
    int ids[QS_MAX_OBSTACLES], map[64][64];
    memset(map, 0, sizeof(map));
    rnd_choice(e, SITE_SCENARIO, 0, L * W, M, ids);
    e->n_obst = M;
    for (int m = 0; m < M; ++m) {
        int rid = ids[m] / W, cid = ids[m] - rid * W;
        map[rid][cid] = 1;
        cell_center(c, rid + L * cid, e->obst_xy[m]);                 /* quadrotor_multi.py:420-424 */
    }
    /* mode_index = rng.integers(0, 100) (quadrotor_multi.py:452), Scenario_mix.reset picks list[idx % 2] (mix.py:79-86) */
    int scen = c->scenario;
    if (scen == QS_SCENARIO_O_MIX) {
        int mode_index = (int)floor(rnd_u(e, SITE_SCENARIO, 0xFF, 1, 0) * 100.0);
        /* a single drone draws from QUADS_MODE_LIST_OBSTACLES_SINGLE = ['o_random'] (mix.py:49-51, utils.py:23) */
        scen = (K == 1 || mode_index % 2 == 0) ? QS_SCENARIO_O_RANDOM : QS_SCENARIO_O_STATIC_SAME_GOAL;
    }
    e->scenario_now = scen;
    /* free cells in row-major order (np.where, o_static_same_goal.py:37-38) */
    int fr[4096], fc[4096], nf = 0;
    for (int r = 0; r < L; ++r) for (int cc = 0; cc < W; ++cc) if (!map[r][cc]) { fr[nf] = r; fc[nf] = cc; ++nf; }
    int pick[QS_MAX_AGENTS];
    if (scen == QS_SCENARIO_O_STATIC_SAME_GOAL) {
        burn_u(e, 1);                                                /* duration_time, o_static_same_goal.py:29 */
        rnd_choice(e, SITE_SCENARIO, 2, nf, K, pick);                /* generate_pos_obst_map_2, o_base.py:69-81 */
        for (int i = 0; i < K; ++i) {
            double xy[2];
            cell_center(c, fr[pick[i]] + L * fc[pick[i]], xy);
            e->d[i].spawn_point[0] = xy[0]; e->d[i].spawn_point[1] = xy[1];
            e->d[i].spawn_point[2] = 1.0 + 2.0 * rnd_u(e, SITE_SCENARIO, 0xFF, 3, i);
        }
        /* max_square_area_center, o_base.py:124-153 (dp row/col 0 copy the obstacle map itself) */
        int dp[64][64], max_size = 0, cx = 0, cy = 0;
        memset(dp, 0, sizeof(dp));
        for (int j = 0; j < W; ++j) dp[0][j] = map[0][j];
        for (int r = 0; r < L; ++r) dp[r][0] = map[r][0];
        for (int r = 1; r < L; ++r)
            for (int j = 1; j < W; ++j)
                if (map[r][j] == 0) {
                    int mn = dp[r - 1][j]; if (dp[r][j - 1] < mn) mn = dp[r][j - 1]; if (dp[r - 1][j - 1] < mn) mn = dp[r - 1][j - 1];
                    dp[r][j] = mn + 1;
                    if (dp[r][j] > max_size) { max_size = dp[r][j]; cx = r - (max_size - 1) / 2; cy = j - (max_size - 1) / 2; }
                }
        double gxy[2];
        cell_center(c, cx + W * cy, gxy);
        double gz = 1.5 + 1.5 * rnd_u(e, SITE_SCENARIO, 0xFF, 4, 0);
        for (int i = 0; i < K; ++i) { e->d[i].goal[0] = gxy[0]; e->d[i].goal[1] = gxy[1]; e->d[i].goal[2] = gz; }
        e->approach_metric = 1.0;
    } else {                                                         /* o_random.py:26-52 */
        if (e->use_tape) { e->ic += 2 * K; e->iu += 2 * K; }           /* 2K generate_pos_obst_map() calls, results discarded */
        rnd_choice(e, SITE_SCENARIO, 2, nf, K, pick);
        for (int i = 0; i < K; ++i) {
            double xy[2];
            cell_center(c, fr[pick[i]] + L * fc[pick[i]], xy);
            e->d[i].spawn_point[0] = xy[0]; e->d[i].spawn_point[1] = xy[1];
            e->d[i].spawn_point[2] = 1.0 + 2.0 * rnd_u(e, SITE_SCENARIO, 0xFF, 3, i);
        }
        rnd_choice(e, SITE_SCENARIO, 5, nf, K, pick);
        for (int i = 0; i < K; ++i) {
            double xy[2];
            cell_center(c, fr[pick[i]] + L * fc[pick[i]], xy);
            e->d[i].goal[0] = xy[0]; e->d[i].goal[1] = xy[1];
            e->d[i].goal[2] = 1.0 + 2.0 * rnd_u(e, SITE_SCENARIO, 0xFF, 6, i);
        }
        burn_u(e, 1);                                                /* duration_step, o_random.py:46 */
        e->approach_metric = 0.5;
    }
}

#include "scenario_oracle.inc"

/* QuadrotorSingle._reset, quadrotor_single.py:401-469 */
static void drone_reset(qo_env *e, int i)
{
    const qs_config *c = &e->c;
    qo_drone *q = &e->d[i];
    for (int a = 0; a < 3; ++a) q->pos[a] = (-c->spawn_box + 2.0 * c->spawn_box * rnd_u(e, SITE_SPAWN, i, 0, a)) + q->spawn_point[a];
    if (q->pos[2] < c->spawn_min_z) q->pos[2] = c->spawn_min_z;
    /* randyaw() until the body x axis points within 60 deg of the origin, :454-456 */
    double hx = -q->pos[0], hy = -q->pos[1], hn = sqrt(hx * hx + hy * hy);
    if (!(hn < 0.00001)) { hx /= hn; hy /= hn; }
    double theta = atan2(hy, hx);
    for (int att = 0; att < 64; ++att) {
        double t = -QO_PI + 2.0 * QO_PI * rnd_u(e, SITE_SPAWN, i, 1, att);
        mg_note(e, QO_MG_YAW, cos(t) * hx + sin(t) * hy, 0.5, 1.0);
        if (!(cos(t) * hx + sin(t) * hy < 0.5)) { theta = t; break; }
    }
    yaw_rot(theta, q->rot);
    for (int a = 0; a < 3; ++a) { q->vel[a] = 0; q->omega[a] = 0; q->acc[a] = 0; }
    q->accm[0] = 0; q->accm[1] = 0; q->accm[2] = QO_GRAV;
    for (int m = 0; m < 4; ++m) { q->rot_damp[m] = 0; q->cmds_damp[m] = 0; }
    q->on_floor = q->crashed_floor = q->crashed_wall = q->crashed_ceiling = 0;
}

/* QuadrotorEnvMulti.reset, quadrotor_multi.py:440-519 */
static void env_reset(qo_env *e, double *obs)
{
    const qs_config *c = &e->c;
    int K = e->K;
    if (c->use_obstacles) obstacle_scenario_reset(e);
    else {
        /* scenario.reset() then goal / spawn_point = scenario.goals[i] (quadrotor_multi.py:459,467-472).  static_same_goal draws
         * 3 + (K-1) unit values that cannot change its K identical goals */
        formation_scenario_reset(e);
        for (int i = 0; i < K; ++i) { memcpy(e->d[i].goal, e->sc.goals[i], sizeof(double) * 3); memcpy(e->d[i].spawn_point, e->d[i].goal, sizeof(double) * 3); }
    }
    int D = obs_dim(c), S = self_obs_dim(c), NB = (c->neighbor_obs_type == QS_NEIGHBOR_POS_VEL) ? 6 * c->neighbor_visible_num : 0;
    for (int i = 0; i < K; ++i) {
        drone_reset(e, i);
        self_obs(e, i, SITE_SENSOR_RESET, obs + (size_t)i * D);
    }
    e->tick = 0;
    for (int i = 0; i < K; ++i) neighbor_obs(e, i, obs + (size_t)i * D + S);    /* stale snap_vel, :477-481 */
    if (c->use_obstacles) for (int i = 0; i < K; ++i) sdf_obs(e, i, obs + (size_t)i * D + S + NB);
    e->collisions_per_episode = e->collisions_after_settle = e->collisions_final_5s = 0;
    e->col_room = e->col_floor = e->col_wall = e->col_ceiling = 0;
    e->obst_col_per_episode = e->obst_col_after_settle = 0;
    for (int i = 0; i < K; ++i) {
        qo_drone *q = &e->d[i];
        q->col_mask = 0; q->prev_new_wall = q->prev_new_ceiling = q->prev_new_room = q->prev_obst_hit = 0;
        q->dist_n = 0; q->sum1 = q->sum3 = q->sum5 = 0; q->n1 = q->n3 = q->n5 = 0;
        q->reached_goal = 0; q->col_agent_ok = 1; q->col_obst_ok = 1; q->ep_reward = 0;
    }
}

/* infos[i]['episode_extra_stats'] of the episode that just ended (quadrotor_multi.py:739-831), in the QS_ER_* layout */
static void write_episode_record(qo_env *e, int ep_len, int success, int nonfinite)
{
    int32_t *r = e->ep_rec;
    int K = e->K, n_succ = 0, n_dead = 0, n_col = 0, n_ncol = 0, n_ocol = 0;
    for (int i = 0; i < K; ++i) {
        const qo_drone *q = &e->d[i];
        int ok = q->col_agent_ok && q->col_obst_ok;
        n_succ += ok && q->reached_goal; n_dead += ok && !q->reached_goal; n_col += !ok;
        n_ncol += !q->col_agent_ok; n_ocol += !q->col_obst_ok;
    }
    r[QS_ER_SEQ] += 1; r[QS_ER_SCENARIO] = e->scenario_now;
    r[QS_ER_NUM_COLLISIONS] = e->collisions_per_episode; r[QS_ER_COLLISIONS_AFTER_SETTLE] = e->collisions_after_settle;
    r[QS_ER_COLLISIONS_FINAL_5S] = e->collisions_final_5s; r[QS_ER_COLLISIONS_ROOM] = e->col_room; r[QS_ER_COLLISIONS_FLOOR] = e->col_floor;
    r[QS_ER_COLLISIONS_WALL] = e->col_wall; r[QS_ER_COLLISIONS_CEILING] = e->col_ceiling; r[QS_ER_COLLISIONS_OBST] = e->obst_col_per_episode;
    r[QS_ER_COLLISIONS_OBST_AFTER_SETTLE] = e->obst_col_after_settle;
    r[QS_ER_AGENTS_SUCCESS] = n_succ; r[QS_ER_AGENTS_DEADLOCK] = n_dead; r[QS_ER_AGENTS_COLLIDED] = n_col;
    r[QS_ER_AGENTS_NEIGHBOR_COL] = n_ncol; r[QS_ER_AGENTS_OBST_COL] = n_ocol;
    r[QS_ER_EP_LEN] = ep_len; r[QS_ER_SUCCESS] = success; r[QS_ER_NONFINITE] = nonfinite;
}

/* ------------------------------------------------------------------------------------------------ */
/* QuadrotorEnvMulti.step, quadrotor_multi.py:521-842                                                 */
/* ------------------------------------------------------------------------------------------------ */
static void env_step(qo_env *e, const double *actions, double *obs, double *rew, uint8_t *done, double *terminal_obs)
{
    const qs_config *c = &e->c;
    int K = e->K;
    double rewraw_pos[QS_MAX_AGENTS];
    int time_remain = c->ep_len - e->tick;                      /* quadrotor_single.py:361 */
    int D = obs_dim(c), S = self_obs_dim(c), NB = (c->neighbor_obs_type == QS_NEIGHBOR_POS_VEL) ? 6 * c->neighbor_visible_num : 0;
    double control_freq = floor(1.0 / c->dt + 0.5) / c->sim_steps;  /* sim_freq / sim_steps, quadrotor_single.py:160 */
    int svd_fire[16];
    for (int s = 0; s < c->sim_steps; ++s) { e->svd_ctr += 1; svd_fire[s] = e->svd_ctr >= c->svd_period; if (svd_fire[s]) e->svd_ctr = 0; }
    /* per-drone QuadrotorSingle._step, quadrotor_single.py:355-371: control+dynamics, reward, self observation */
    for (int i = 0; i < K; ++i) {
        drone_control_step(e, i, actions + 4 * i, svd_fire);
        rew[i] = base_reward(e, i, actions + 4 * i, &rewraw_pos[i]);
        self_obs(e, i, SITE_SENSOR, obs + (size_t)i * D);
    }
    e->tick += 1;
    int all_done = e->tick > c->ep_len;                           /* :366-367 */
    int tick = e->tick;

    /* 1.1 drone-drone collisions, collisions/quadrotors.py:63-91 + quadrotor_multi.py:537-568 */
    double thr_col = c->collision_hitbox_radius * c->arm, thr_fall = c->collision_falloff_radius * c->arm;
    uint32_t row[QS_MAX_AGENTS];
    double prox[QS_MAX_AGENTS];
    memset(row, 0, sizeof(row)); memset(prox, 0, sizeof(prox));
    double pen_ratio = -c->rew_quadcol_bin_smooth_max / thr_fall;
    for (int i = 0; i < K; ++i)
        for (int j = i + 1; j < K; ++j) {
            double dx = e->d[i].pos[0] - e->d[j].pos[0], dy = e->d[i].pos[1] - e->d[j].pos[1], dz = e->d[i].pos[2] - e->d[j].pos[2];
            double dist = sqrt(dx * dx + dy * dy + dz * dz);
            {
                double sc = 0; for (int a = 0; a < 3; ++a) sc = fmax(sc, fmax(fabs(e->d[i].pos[a]), fabs(e->d[j].pos[a])));
                mg_note(e, QO_MG_PAIR, dist, thr_col, sc); mg_note(e, QO_MG_FALLOFF, dist, thr_fall, sc);
            }
            if (dist <= thr_col) { row[i] |= 1u << j; row[j] |= 1u << i; }
            if (dist <= thr_fall) { double pen = pen_ratio * dist + c->rew_quadcol_bin_smooth_max; prox[i] += pen; prox[j] += pen; }  /* quadrotors.py:95-103 */
        }
    /* setdiff1d(flatten(curr pairs), flatten(prev pairs)) over drone ids, :548 */
    int n_unique = 0, unique_any_nonzero = 0;
    int is_unique[QS_MAX_AGENTS];
    for (int i = 0; i < K; ++i) {
        is_unique[i] = (row[i] != 0) && (e->d[i].col_mask == 0);
        n_unique += is_unique[i];
        if (is_unique[i] && i != 0) unique_any_nonzero = 1;
        e->last_new_pairs[i] = row[i] & ~e->d[i].col_mask;        /* :545-546 */
    }
    int col_tick = n_unique / 2;                                  /* :557 */
    e->collisions_per_episode += col_tick;
    double grace = 1.5 * control_freq, final_grace = 5.0 * control_freq;   /* :156,160 */
    if (col_tick > 0 && (double)tick >= grace) {                   /* :560 */
        e->collisions_after_settle += col_tick;
        for (int i = 0; i < K; ++i) if (is_unique[i]) e->d[i].col_agent_ok = 0;
    }
    if (col_tick > 0 && (double)time_remain <= final_grace) e->collisions_final_5s += col_tick;   /* :564 */

    /* 1.2 obstacles, obstacles/utils.py:31-43 + quadrotor_multi.py:571-597 */
    int obst_hit[QS_MAX_AGENTS], obst_new[QS_MAX_AGENTS], n_obst_new = 0;
    memset(obst_new, 0, sizeof(obst_new));
    for (int i = 0; i < K; ++i) obst_hit[i] = -1;
    if (c->use_obstacles) {
        double thr_o = c->arm + c->obst_size / 2.0;
        for (int i = 0; i < K; ++i)
            for (int m = 0; m < e->n_obst; ++m) {
                double dx = e->d[i].pos[0] - e->obst_xy[m][0], dy = e->d[i].pos[1] - e->obst_xy[m][1];
                mg_note(e, QO_MG_OBST, sqrt(dx * dx + dy * dy), thr_o, fmax(fabs(e->d[i].pos[0]), fabs(e->d[i].pos[1])));
                if (sqrt(dx * dx + dy * dy) <= thr_o) { obst_hit[i] = m; break; }
            }
        for (int i = 0; i < K; ++i) {
            obst_new[i] = (obst_hit[i] >= 0) && !e->d[i].prev_obst_hit;
            n_obst_new += obst_new[i];
        }
        e->obst_col_per_episode += n_obst_new;
        if (n_obst_new > 0 && (double)tick >= grace) {
            e->obst_col_after_settle += n_obst_new;
            for (int i = 0; i < K; ++i) if (obst_new[i]) e->d[i].col_obst_ok = 0;
        }
        for (int i = 0; i < K; ++i) e->d[i].prev_obst_hit = obst_hit[i] >= 0;
    }

    /* 1.3 room, quadrotor_multi.py:390-403, 600-606 */
    int new_wall[QS_MAX_AGENTS], new_ceil[QS_MAX_AGENTS], n_floor = 0, n_wall = 0, n_ceil = 0, n_room = 0;
    for (int i = 0; i < K; ++i) {
        qo_drone *q = &e->d[i];
        new_wall[i] = q->crashed_wall && !q->prev_new_wall;
        new_ceil[i] = q->crashed_ceiling && !q->prev_new_ceiling;
        int in_room = q->crashed_floor || new_wall[i] || new_ceil[i];
        int new_room = in_room && !q->prev_new_room;
        n_floor += q->crashed_floor; n_wall += new_wall[i]; n_ceil += new_ceil[i]; n_room += new_room;
        q->prev_new_wall = new_wall[i]; q->prev_new_ceiling = new_ceil[i]; q->prev_new_room = new_room;
    }
    if ((double)tick >= grace) { e->col_room += n_room; e->col_floor += n_floor; e->col_wall += n_wall; e->col_ceiling += n_ceil; }  /* :631-635 */

    /* 2. rewards, :610-655 */
    for (int i = 0; i < K; ++i) {
        double rc = (unique_any_nonzero && is_unique[i]) ? -1.0 : 0.0;   /* `.any()` guard, :611 */
        rew[i] += c->rew_quadcol_bin * rc;
        rew[i] += -1.0 * ((c->dt * c->sim_steps) * prox[i]);
        if (c->use_obstacles) rew[i] += c->rew_quadcol_bin_obst * (obst_new[i] ? -1.0 : 0.0);   /* :594-597 */
        e->rew_info[i][QS_RI_RAW_QUADCOL] = rc; e->rew_info[i][QS_RI_PROXIMITY] = -1.0 * ((c->dt * c->sim_steps) * prox[i]);   /* :642-649 */
        e->rew_info[i][QS_RI_RAW_QUADCOL_OBST] = (c->use_obstacles && obst_new[i]) ? -1.0 : 0.0;
        qo_drone *q = &e->d[i];
        double dlog = -rewraw_pos[i];                             /* :651 */
        q->dist_hist[q->dist_n % 5] = dlog; q->dist_n++;
        if (q->dist_n >= 5 && !q->reached_goal) {
            double m5 = 0; for (int a = 0; a < 5; ++a) m5 += q->dist_hist[a];
            mg_note(e, QO_MG_REACH, (m5 / 5.0) / c->dt, e->approach_metric, (m5 / 5.0) / c->dt);
            if ((m5 / 5.0) / c->dt < e->approach_metric) q->reached_goal = 1;
        }
        /* running sums for distance_to_goal_{1,3,5}s (:762-767): windows are the last 100/300/500 control steps of a
         * fixed-length episode (ticks ep_len+1-n+1 .. ep_len+1) */
        int steps_left = c->ep_len + 1 - tick;
        if (steps_left < 100) { q->sum1 += dlog; q->n1++; }
        if (steps_left < 300) { q->sum3 += dlog; q->n3++; }
        if (steps_left < 500) { q->sum5 += dlog; q->n5++; }
    }

    /* 3. impulses, :659-698 */
    int flag = 0;
    if (c->use_downwash) flag |= downwash(e);
    if (c->apply_collision_force) {
        for (int i = 0; i < K; ++i)
            for (int j = i + 1; j < K; ++j)
                if (e->last_new_pairs[i] & (1u << j)) { pair_impulse(e, i, j); flag = 1; }
        if (c->use_obstacles) for (int i = 0; i < K; ++i) if (obst_new[i]) { obstacle_impulse(e, i, obst_hit[i]); flag = 1; }
        for (int i = 0; i < K; ++i) if (new_wall[i]) { room_impulse(e, i, 1); flag = 1; }
        for (int i = 0; i < K; ++i) if (new_ceil[i]) { room_impulse(e, i, 0); flag = 1; }
    }
    for (int i = 0; i < K; ++i) e->d[i].col_mask = row[i];       /* :568 */
    e->last_impulse_flag = flag;

    /* 4. scenario.step(), :701 -- goals may move; the self observations computed above keep the old goal unless an impulse
     * forces their recomputation below */
    if (!c->use_obstacles) formation_scenario_step(e);

    /* 5. observations, :703-720 */
    for (int i = 0; i < K; ++i) memcpy(e->snap_vel[i], e->d[i].vel, sizeof(double) * 3);
    if (flag) for (int i = 0; i < K; ++i) self_obs(e, i, SITE_SENSOR_IMPULSE, obs + (size_t)i * D);   /* :711-712, fresh noise */
    for (int i = 0; i < K; ++i) neighbor_obs(e, i, obs + (size_t)i * D + S);
    if (c->use_obstacles) for (int i = 0; i < K; ++i) sdf_obs(e, i, obs + (size_t)i * D + S + NB);
    for (int i = 0; i < K; ++i) { e->d[i].ep_reward += rew[i]; done[i] = (uint8_t)all_done; }

    /* 7. dones, :739-838 */
    if (all_done) {
        if (terminal_obs) memcpy(terminal_obs, obs, sizeof(double) * (size_t)K * D);
        qs_stats *s = &e->stats;
        s->episodes += 1;
        s->num_collisions += e->collisions_per_episode; s->num_collisions_after_settle += e->collisions_after_settle;
        s->num_collisions_final_5s += e->collisions_final_5s; s->num_collisions_with_room += e->col_room;
        s->num_collisions_with_floor += e->col_floor; s->num_collisions_with_wall += e->col_wall;
        s->num_collisions_with_ceiling += e->col_ceiling; s->num_collisions_obst_quad += e->obst_col_per_episode;
        s->num_collisions_obst_quad_after_settle += e->obst_col_after_settle;
        write_episode_record(e, tick, 0, 0);
        for (int i = 0; i < K; ++i) {
            qo_drone *q = &e->d[i];
            int ok = q->col_agent_ok && q->col_obst_ok;
            s->agents_success += ok && q->reached_goal; s->agents_deadlock += ok && !q->reached_goal; s->agents_collided += !ok;
            double d1 = q->n1 ? (1.0 / c->dt) * q->sum1 / q->n1 : 0.0, d3 = q->n3 ? (1.0 / c->dt) * q->sum3 / q->n3 : 0.0,
                   d5 = q->n5 ? (1.0 / c->dt) * q->sum5 / q->n5 : 0.0;
            s->distance_to_goal_1s += d1; s->distance_to_goal_3s += d3; s->distance_to_goal_5s += d5;
            e->ep_agent[i][0] = d1; e->ep_agent[i][1] = d3; e->ep_agent[i][2] = d5; e->ep_agent[i][3] = 0.0;
            s->reward_sum += q->ep_reward;
        }
        env_reset(e, obs);                                        /* :836 */
    }
    e->step_ctr += 1;
}

#include "fork_oracle.inc"

/* ------------------------------------------------------------------------------------------------ */
/* exported API (ctypes)                                                                              */
/* ------------------------------------------------------------------------------------------------ */
qo_env *qo_create(const qs_config *cfg, int env_index)
{
    if (!cfg || cfg->num_agents < 1 || cfg->num_agents > QS_MAX_AGENTS) return NULL;
    qo_env *e = (qo_env *)calloc(1, sizeof(qo_env));
    e->c = *cfg;
    e->K = cfg->num_agents;
    e->gid = (uint32_t)(cfg->env_id_offset + env_index);
    e->approach_metric = cfg->approach_goal_metric;
    e->scenario_now = cfg->scenario;
    for (int i = 0; i < e->K; ++i) { double I[9] = { 1, 0, 0, 0, 1, 0, 0, 0, 1 }; memcpy(e->d[i].rot, I, sizeof(I)); e->d[i].col_agent_ok = e->d[i].col_obst_ok = 1; }
    return e;
}
void qo_destroy(qo_env *e) { free(e); }
static int is_fork(const qo_env *e) { return e->c.env_mode == QS_MODE_FORK; }
static int any_obs_dim(const qs_config *c) { return c->env_mode == QS_MODE_FORK ? fork_obs_dim(c) : obs_dim(c); }
static int any_act_dim(const qs_config *c) { return c->env_mode == QS_MODE_FORK ? 2 : 4; }
int qo_obs_dim(const qo_env *e) { return any_obs_dim(&e->c); }
int qo_act_dim(const qo_env *e) { return any_act_dim(&e->c); }

void qo_set_tape(qo_env *e, const double *normals, int nn, const double *uniforms, int nu, const double *choices, int nc)
{
    e->tn = normals; e->nn = nn; e->tu = uniforms; e->nu = nu; e->tc = choices; e->nc = nc;
    e->in_ = e->iu = e->ic = 0;
    e->use_tape = (normals != NULL || uniforms != NULL || choices != NULL);
}
/* how far the tape was consumed (tests assert it was consumed exactly) */
void qo_tape_pos(const qo_env *e, int *in_, int *iu, int *ic) { *in_ = e->in_; *iu = e->iu; *ic = e->ic; }

void qo_reset(qo_env *e, double *obs) { mg_clear(e); if (is_fork(e)) fork_env_reset(e, obs); else env_reset(e, obs); e->step_ctr += 1; }
void qo_step(qo_env *e, const double *actions, double *obs, double *rew, uint8_t *done, double *terminal_obs, uint8_t *reset_success)
{
    mg_clear(e);
    if (is_fork(e)) fork_env_step(e, actions, obs, rew, done, terminal_obs, reset_success);
    else { env_step(e, actions, obs, rew, done, terminal_obs); if (reset_success && done[0]) *reset_success = 0; }
}
/* fork-mode state: pid [K,24], heading [K,3] = (angle, angular_velocity, self.heading snapshot), evader [2] */
void qo_get_fork_state(const qo_env *e, double *pid, double *heading, double *evader, int32_t *env_flags /* [2]: episode_success, chasers_placed */)
{
    for (int i = 0; i < e->K; ++i) {
        if (pid) memcpy(pid + 24 * i, e->d[i].pid, sizeof(double) * 24);
        if (heading) { heading[3 * i] = e->d[i].angle; heading[3 * i + 1] = e->d[i].ang_vel; heading[3 * i + 2] = e->snap_heading[i]; }
    }
    if (evader) { evader[0] = e->evader[0]; evader[1] = e->evader[1]; }
    if (env_flags) { env_flags[0] = e->episode_success; env_flags[1] = e->chasers_placed; }
}
void qo_set_fork_state(qo_env *e, const double *pid, const double *heading, const double *evader, const int32_t *env_flags)
{
    if (env_flags) { e->episode_success = env_flags[0]; e->chasers_placed = env_flags[1]; }
    for (int i = 0; i < e->K; ++i) {
        if (pid) memcpy(e->d[i].pid, pid + 24 * i, sizeof(double) * 24);
        if (heading) { e->d[i].angle = heading[3 * i]; e->d[i].ang_vel = heading[3 * i + 1]; e->snap_heading[i] = heading[3 * i + 2]; }
    }
    if (evader) { e->evader[0] = evader[0]; e->evader[1] = evader[1]; }
}

/* one free-flight control step of drone 0 only (golden vector A.1 of SURVEY.md) */
void qo_dynamics_only(qo_env *e, int drone, const double *thrust_cmd01)
{
    for (int s = 0; s < e->c.sim_steps; ++s) dynamics_substep(e, drone, thrust_cmd01, s, 0);
}

/* state access: arrays of K rows */
void qo_get_state(const qo_env *e, double *pos, double *vel, double *rot, double *omega, double *rot_damp, double *cmds_damp,
                  double *ou, double *goal, int32_t *flags, uint32_t *col_mask, int32_t *tick_svd_step)
{
    for (int i = 0; i < e->K; ++i) {
        const qo_drone *q = &e->d[i];
        if (pos) memcpy(pos + 3 * i, q->pos, 24);
        if (vel) memcpy(vel + 3 * i, q->vel, 24);
        if (rot) memcpy(rot + 9 * i, q->rot, 72);
        if (omega) memcpy(omega + 3 * i, q->omega, 24);
        if (rot_damp) memcpy(rot_damp + 4 * i, q->rot_damp, 32);
        if (cmds_damp) memcpy(cmds_damp + 4 * i, q->cmds_damp, 32);
        if (ou) memcpy(ou + 4 * i, q->ou, 32);
        if (goal) memcpy(goal + 3 * i, q->goal, 24);
        if (flags) flags[i] = q->on_floor | q->crashed_floor << 1 | q->crashed_wall << 2 | q->crashed_ceiling << 3 |
                              q->prev_new_wall << 4 | q->prev_new_ceiling << 5 | q->prev_new_room << 6 | q->prev_obst_hit << 7;
        if (col_mask) col_mask[i] = q->col_mask;
    }
    if (tick_svd_step) { tick_svd_step[0] = e->tick; tick_svd_step[1] = e->svd_ctr; tick_svd_step[2] = (int32_t)e->step_ctr; }
}

void qo_set_state(qo_env *e, const double *pos, const double *vel, const double *rot, const double *omega, const double *rot_damp,
                  const double *cmds_damp, const double *ou, const double *goal, const int32_t *flags, const uint32_t *col_mask,
                  const int32_t *tick_svd_step)
{
    for (int i = 0; i < e->K; ++i) {
        qo_drone *q = &e->d[i];
        if (pos) memcpy(q->pos, pos + 3 * i, 24);
        if (vel) { memcpy(q->vel, vel + 3 * i, 24); memcpy(e->snap_vel[i], vel + 3 * i, 24); }
        if (rot) memcpy(q->rot, rot + 9 * i, 72);
        if (omega) memcpy(q->omega, omega + 3 * i, 24);
        if (rot_damp) memcpy(q->rot_damp, rot_damp + 4 * i, 32);
        if (cmds_damp) memcpy(q->cmds_damp, cmds_damp + 4 * i, 32);
        if (ou) memcpy(q->ou, ou + 4 * i, 32);
        if (goal) { memcpy(q->goal, goal + 3 * i, 24); memcpy(e->sc.goals[i], goal + 3 * i, 24); }
        if (flags) {
            q->on_floor = flags[i] & 1; q->crashed_floor = (flags[i] >> 1) & 1; q->crashed_wall = (flags[i] >> 2) & 1;
            q->crashed_ceiling = (flags[i] >> 3) & 1; q->prev_new_wall = (flags[i] >> 4) & 1; q->prev_new_ceiling = (flags[i] >> 5) & 1;
            q->prev_new_room = (flags[i] >> 6) & 1; q->prev_obst_hit = (flags[i] >> 7) & 1;
        }
        if (col_mask) q->col_mask = col_mask[i];
    }
    if (tick_svd_step) { e->tick = tick_svd_step[0]; e->svd_ctr = tick_svd_step[1]; e->step_ctr = (uint32_t)tick_svd_step[2]; }
}

/* formation-scenario state in the flat layout of qs_state_view.scenario (include/quadsim.h, QS_SC_*) */
void qo_get_scenario(const qo_env *e, double *o)
{
    const qo_scen *s = &e->sc;
    memset(o, 0, sizeof(double) * QS_SC_COUNT);
    o[QS_SC_SCENARIO] = e->scenario_now; o[QS_SC_FORMATION] = s->formation; o[QS_SC_SIZE] = s->size; o[QS_SC_LAYER_DIST] = s->layer_dist;
    o[QS_SC_HIGHEST] = s->highest; o[QS_SC_LOWEST] = s->lowest;
    for (int a = 0; a < 3; ++a) o[QS_SC_CENTER + a] = s->center[a];
    o[QS_SC_CTL_STEPS] = s->ctl_steps; o[QS_SC_INCREASE] = s->increase; o[QS_SC_SPEED] = s->speed;
    if (e->scenario_now == QS_SCENARIO_EP_RAND_BEZIER) for (int k = 0; k < 9; ++k) o[QS_SC_AUX + k] = s->bez[k / 3][k % 3];
    else for (int a = 0; a < 3; ++a) { o[QS_SC_AUX + a] = s->c1[a]; o[QS_SC_AUX + 3 + a] = s->c2[a]; }
}
void qo_set_scenario(qo_env *e, const double *o)
{
    qo_scen *s = &e->sc;
    e->scenario_now = (int)o[QS_SC_SCENARIO]; s->formation = (int)o[QS_SC_FORMATION]; s->size = o[QS_SC_SIZE]; s->layer_dist = o[QS_SC_LAYER_DIST];
    s->highest = o[QS_SC_HIGHEST]; s->lowest = o[QS_SC_LOWEST];
    s->per_layer = (s->formation >= QF_GRID_H && s->formation <= QF_GRID_YZ) ? 50 : 8;
    for (int a = 0; a < 3; ++a) s->center[a] = o[QS_SC_CENTER + a];
    s->ctl_steps = (int)o[QS_SC_CTL_STEPS]; s->increase = (int)o[QS_SC_INCREASE]; s->speed = o[QS_SC_SPEED];
    s->n_goals = e->K;                                          /* swap_goals permutes the K rows the drones hold */
    if (e->scenario_now == QS_SCENARIO_EP_RAND_BEZIER) for (int k = 0; k < 9; ++k) s->bez[k / 3][k % 3] = o[QS_SC_AUX + k];
    else for (int a = 0; a < 3; ++a) { s->c1[a] = o[QS_SC_AUX + a]; s->c2[a] = o[QS_SC_AUX + 3 + a]; }
}
/* QuadrotorScenario.generate_goals on its own (formation fixture of tests/golden/formations.npz) */
int qo_generate_goals(int formation, double size, int n, const double *center, double layer_dist, int cube_dim, double *out)
{
    qo_scen s;
    memset(&s, 0, sizeof(s));
    s.formation = formation; s.size = size;
    s.per_layer = (formation >= QF_GRID_H && formation <= QF_GRID_YZ) ? 50 : 8;
    return scen_generate_goals(&s, n, center, layer_dist, (double (*)[3])out, cube_dim);
}

double qo_col_norm_and_new_vel_obst(const double *pos, const double *vel, const double *obst_pos, double *norm_out)
{
    return col_norm_and_new_vel_obst(pos, vel, obst_pos, norm_out);
}

void qo_set_obstacles(qo_env *e, const double *xy, int n) { e->n_obst = n; for (int m = 0; m < n; ++m) { e->obst_xy[m][0] = xy[2 * m]; e->obst_xy[m][1] = xy[2 * m + 1]; } }
void qo_get_obstacles(const qo_env *e, double *xy, int *n) { *n = e->n_obst; for (int m = 0; m < e->n_obst; ++m) { xy[2 * m] = e->obst_xy[m][0]; xy[2 * m + 1] = e->obst_xy[m][1]; } }
void qo_get_stats(const qo_env *e, qs_stats *out) { *out = e->stats; }
void qo_get_reward_info(const qo_env *e, double *out) { for (int i = 0; i < e->K; ++i) memcpy(out + (size_t)i * QS_RI_COUNT, e->rew_info[i], sizeof(double) * QS_RI_COUNT); }
/* min distance-to-threshold of the last qo_step / qo_reset per decision class, in fp32 ulps (QO_MG_* order; 1e300 = class not exercised) */
void qo_get_margins(const qo_env *e, double *out) { for (int k = 0; k < QO_MG_COUNT; ++k) out[k] = e->margin[k]; }
int qo_margin_count(void) { return QO_MG_COUNT; }
void qo_get_record(const qo_env *e, int32_t *env_rec, double *agent_rec)
{
    memcpy(env_rec, e->ep_rec, sizeof(e->ep_rec));
    for (int i = 0; i < e->K; ++i) memcpy(agent_rec + 4 * i, e->ep_agent[i], sizeof(double) * 4);
}
void qo_get_diag(const qo_env *e, uint32_t *new_pairs, int32_t *neighbors, int32_t *impulse_flag)
{
    for (int i = 0; i < e->K; ++i) {
        if (new_pairs) new_pairs[i] = e->last_new_pairs[i];
        if (neighbors) for (int s = 0; s < e->c.neighbor_visible_num; ++s) neighbors[i * e->c.neighbor_visible_num + s] = e->last_neighbors[i][s];
    }
    if (impulse_flag) *impulse_flag = e->last_impulse_flag;
}
void qo_set_param(qo_env *e, int key, double v)
{
    double *p[QS_PARAM_COUNT] = { &e->c.rew_pos, &e->c.rew_effort, &e->c.rew_crash, &e->c.rew_orient, &e->c.rew_spin,
                                  &e->c.rew_quadcol_bin, &e->c.rew_quadcol_bin_smooth_max, &e->c.rew_quadcol_bin_obst,
                                  &e->c.fork.capture_radius };
    if (key >= 0 && key < QS_PARAM_COUNT) *p[key] = v;
}

/* raw generator access so tests can pin the RNG contract bit-for-bit against the CUDA side */
void qo_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t *out) { philox4x32_10(c0, c1, c2, c3, k0, k1, out); }

/* batched stepping over a slice of independent envs: the CPU baseline bench.py times.  The Python wrapper calls this
 * from one thread per host core on disjoint slices (ctypes drops the GIL), mirroring the reference's
 * one-process-per-env data parallelism (swarm_rl/env_wrappers/subproc_vec_env_custom.py:112-153). */
void qo_batch_step(qo_env **envs, int n, const double *actions, double *obs, double *rew, uint8_t *done)
{
    if (n <= 0) return;
    int K = envs[0]->K, D = any_obs_dim(&envs[0]->c), A = any_act_dim(&envs[0]->c);
    for (int k = 0; k < n; ++k)
        qo_step(envs[k], actions + (size_t)k * K * A, obs + (size_t)k * K * D, rew + (size_t)k * K, done + (size_t)k * K, NULL, NULL);
}
void qo_batch_reset(qo_env **envs, int n, double *obs)
{
    if (n <= 0) return;
    int K = envs[0]->K, D = any_obs_dim(&envs[0]->c);
    for (int k = 0; k < n; ++k) qo_reset(envs[k], obs + (size_t)k * K * D);
}

/* `steps` lock-step batch steps on `threads` OpenMP threads inside ONE native call (no Python dispatch per step): envs are cut
 * into static slices, every thread steps its slice, one barrier per step -- the data parallelism of the reference's
 * SubprocVecEnvCustom (one worker per env group, step_wait = barrier) at its cheapest.  actions: pool of `pool` action sets
 * [pool][n*K*A], step s uses set s % pool.  Returns the number of env resets (finished episodes) seen. */
long long qo_batch_run(qo_env **envs, int n, const double *actions, int pool, int steps, double *obs, double *rew, uint8_t *done, int threads)
{
    if (n <= 0 || steps <= 0 || pool <= 0) return 0;
    const int K = envs[0]->K, D = any_obs_dim(&envs[0]->c), A = any_act_dim(&envs[0]->c);
    long long resets = 0;
#ifdef _OPENMP
    if (threads < 1) threads = omp_get_max_threads();
#pragma omp parallel num_threads(threads) reduction(+ : resets)
#endif
    {
        for (int s = 0; s < steps; ++s) {
            const double *a = actions + (size_t)(s % pool) * n * K * A;
#ifdef _OPENMP
#pragma omp for schedule(static)
#endif
            for (int k = 0; k < n; ++k) {
                qo_step(envs[k], a + (size_t)k * K * A, obs + (size_t)k * K * D, rew + (size_t)k * K, done + (size_t)k * K, NULL, NULL);
                resets += done[(size_t)k * K] ? 1 : 0;
            }
        }
    }
    return resets;
}
int qo_omp_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void qo_set_tick(qo_env *e, int tick) { e->tick = tick; }
