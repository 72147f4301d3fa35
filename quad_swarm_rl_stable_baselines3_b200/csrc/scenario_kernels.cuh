// scenario_kernels.cuh -- formation scenarios of the upstream env on the device, included by
// quadsim_kernels.cuh.  Reference: gym_art/quadrotor_multi/scenarios/{base,utils,static_diff_goal,dynamic_same_goal,
// dynamic_diff_goal,swap_goals,dynamic_formations,ep_lissajous3D,ep_rand_bezier,swarm_vs_swarm,mix}.py.
//
// The reference keeps one Python scenario object per env; here its state is a row of QS_SC_COUNT floats per env
// (include/quadsim.h QS_SC_*) and every lane of the env's lane group evaluates the scenario redundantly from the same
// counter-based draws (same keys -> same values), keeping only its own goal row.  Goal rows are produced by formula from
// (formation, size, centre, row index), so a shuffle is a permutation of row indices and needs no data exchange; only
// swap_goals permutes stored goals and goes through the warp's exchange buffer.  The draw addressing is the table in
// oracle/scenario_oracle.inc.
#pragma once

namespace qs {

enum { QF_CIRCLE_H = 0, QF_CIRCLE_XZ, QF_CIRCLE_YZ, QF_SPHERE, QF_GRID_H, QF_GRID_XZ, QF_GRID_YZ, QF_CUBE };   // scenarios/utils.py:25-26
#define QS_BEZIER_MAX_TRIES 1024

// QUADS_PARAMS_DICT, scenarios/utils.py:31-55
__device__ __forceinline__ void scen_params(int scen, int &nform, float &low, float &high)
{
    nform = 1; low = 0.f; high = 0.f;
    if (scen == QS_SCENARIO_STATIC_DIFF_GOAL || scen == QS_SCENARIO_DYNAMIC_DIFF_GOAL || scen == QS_SCENARIO_SWARM_VS_SWARM || scen == QS_SCENARIO_RUN_AWAY) { nform = 8; low = 0.25f; high = 0.5f; }
    else if (scen == QS_SCENARIO_SWAP_GOALS) { nform = 8; low = 0.4f; high = 0.8f; }
    else if (scen == QS_SCENARIO_DYNAMIC_FORMATIONS) { nform = 8; low = 0.f; high = 1.0f; }
}

// get_grid_dim_number, scenarios/utils.py:113-125 (integer square root: the approximate sqrt may round 9 to 2.9999)
__device__ __forceinline__ void grid_dims(int num, int &d1, int &d2)
{
    int a = 1;
#pragma unroll 1
    while ((a + 1) * (a + 1) <= num) ++a;
#pragma unroll 1
    while (a > 1 && (num % a) != 0) --a;
    d1 = a; d2 = num / a;
}

struct Formation { int formation; float size, layer, lowest, highest; };

// update_formation_and_relate_param, scenarios/base.py:119-131; u = draws (formation index, size, layer distance)
__device__ __forceinline__ void scen_update_formation(const DevConst &c, int scen, const float *u, Formation &f)
{
    int nform; float low, high;
    scen_params(scen, nform, low, high);
    f.formation = min((int)floorf(u[0] * (float)nform), nform - 1);
    if (f.formation <= QF_CIRCLE_YZ) {                                  // get_circle_radius(8, dist), utils.py:107-110
        f.lowest = (0.5f * low) / 0.3826834323650898f; f.highest = (0.5f * high) / 0.3826834323650898f;
    } else if (f.formation == QF_SPHERE) {                              // get_sphere_radius, utils.py:96-104
        const int n = (scen == QS_SCENARIO_SWARM_VS_SWARM) ? c.K / 2 : c.K;
        const float ratio = (1.75388487222762f - 0.0920858134405214f) / (1.0f + powf((float)n / 10.3632729642351f, 0.860487305801679f)) + 0.0920858134405214f;
        f.lowest = low / ratio; f.highest = high / ratio;
    } else { f.lowest = low; f.highest = high; }
    f.size = f.lowest + (f.highest - f.lowest) * u[1];
    f.layer = f.lowest + (f.highest - f.lowest) * u[2];
}

__device__ __forceinline__ void goal_by_formation(int orient, float p0, float p1, float layer, float *g)   // utils.py:154-165
{
    if (orient == 0) { g[0] = p0; g[1] = p1; g[2] = layer; }
    else if (orient == 1) { g[0] = p0; g[1] = layer; g[2] = p1; }
    else { g[0] = layer; g[1] = p0; g[2] = p1; }
}

// generate_points, scenarios/utils.py:77-93: point j of the m-point spiral on the unit sphere.  Out of line: the two general-range
// sincosf carry ~1400 instructions of slow-path code that would otherwise sit inside every caller (instruction-cache budget).
static __device__ __noinline__ float3 sphere_point(float m, int r)
{
    const float x = 0.1f + 1.2f * m, start = -1.0f + 1.0f / (m - 1.0f), inc = (2.0f - 2.0f / (m - 1.0f)) / (m - 1.0f);
    const float t = start + (float)r * inc, sg = (t > 0.f) ? 1.f : ((t < 0.f) ? -1.f : 0.f);
    const float b = 1.5707963267948966f * sg * (1.0f - sqrtf(fmaxf(1.0f - fabsf(t), 0.f)));
    float sa, ca, sb, cb;
    sincosf(t * x, &sa, &ca); sincosf(b, &sb, &cb);
    return make_float3(ca * cb, sa * cb, sb);
}

// Row r of QuadrotorScenario.generate_goals(n, centre, layer_dist), scenarios/base.py:42-116.  n <= 32 < 50, so a grid is
// always a single layer; circles stack layers of 8.
__device__ __forceinline__ void formation_goal(int f, float size, int n, const float *center, float layer_dist, int cube_fd, int r, float *out)
{
    if (f <= QF_CIRCLE_YZ) {
        const int per = 8, whole = n / per, rest = n % per;
        const int cur = (n <= per) ? n : ((r / per < whole) ? per : rest);
        float sn, cs;
        sincospif(2.0f * (float)(r % cur) / (float)cur, &sn, &cs);
        goal_by_formation(f, size * cs, size * sn, (float)(r / per) * layer_dist, out);
#pragma unroll
        for (int a = 0; a < 3; ++a) out[a] += center[a];
    } else if (f == QF_SPHERE) {
        const float3 pt = sphere_point((float)max(n, 3), r);
        out[0] = size * pt.x + center[0]; out[1] = size * pt.y + center[1]; out[2] = size * pt.z + center[2];
    } else if (f <= QF_GRID_YZ) {
        int d1, d2;
        grid_dims(n, d1, d2);
        // d1 * d2 == n (d1 divides n), so the column / row indices i % d2 and (i / d2) % d1 average to (d2 - 1) / 2 and (d1 - 1) / 2
        const float m0 = size * 0.5f * (float)(d2 - 1), m1 = size * 0.5f * (float)(d1 - 1);
        float g[3], mg[3];
        goal_by_formation(f - QF_GRID_H, size * (float)(r % d2), size * (float)((r / d2) % d1), 0.f, g);
        goal_by_formation(f - QF_GRID_H, m0, m1, 0.f, mg);
#pragma unroll
        for (int a = 0; a < 3; ++a) out[a] = g[a] - mg[a] + center[a];
    } else {                                                            // cube, base.py:100-112 (x is offset by formation_center[2])
        const int fd = max(cube_fd, 1);
        int s0 = 0, s1 = 0, s2 = 0;
#pragma unroll 1
        for (int i = 0; i < n; ++i) { s0 += i / (fd * fd); s1 += (i / fd) % fd; s2 += i % fd; }
        const float inv = 1.0f / (float)n;
        const float g0 = center[2] + size * (float)(r / (fd * fd)), m0 = center[2] + size * (float)s0 * inv;
        out[0] = g0 - m0 + center[0];
        out[1] = size * (float)((r / fd) % fd) - size * (float)s1 * inv + center[1];
        out[2] = size * (float)(r % fd) - size * (float)s2 * inv + center[2];
    }
}

__device__ __forceinline__ int formation_rows(int f, int n) { return (f == QF_SPHERE && n < 3) ? 3 : n; }

// rng.shuffle as a permutation of row indices: rows[k] afterwards = rows[perm[k]] before (Fisher-Yates from the top)
__device__ __forceinline__ void scen_shuffle_perm(const Rng &g, int aux, int n, unsigned char *perm)
{
    for (int i = 0; i < n; ++i) perm[i] = (unsigned char)i;
    float u[4];
    int t = 0;
    for (int i = n - 1; i > 0; --i, ++t) {
        if ((t & 3) == 0) rng_u4(g, SITE_SCENARIO, 0xFF, aux, t >> 2, u);
        const int j = min((int)floorf(u[t & 3] * (float)(i + 1)), i);
        const unsigned char tmp = perm[i]; perm[i] = perm[j]; perm[j] = tmp;
    }
}

// get_z_value, scenarios/utils.py:168-180
__device__ __forceinline__ float scen_z_value(const DevConst &c, int formation, float size, float u)
{
    const float box = c.spawn_box;
    const float z = (-0.5f * box + box * u) + 2.0f;
    float lb = 0.25f;
    if (formation == QF_SPHERE || formation == QF_CIRCLE_XZ || formation == QF_CIRCLE_YZ) lb = size + 0.25f;
    else if (formation == QF_GRID_XZ || formation == QF_GRID_YZ) { int d1, d2; grid_dims(c.K, d1, d2); lb = (float)d1 * size + 0.25f; }
    return fmaxf(lb, z);
}

// goal row d of the concatenated swarm_vs_swarm formations (create_formations, swarm_vs_swarm.py:60-65)
__device__ __forceinline__ void swarm_goal(const DevConst &c, const Formation &f, const float *c1, const float *c2, int d,
                                           const unsigned char *perm1, const unsigned char *perm2, float *goal)
{
    const int n1 = c.K / 2, n2 = c.K - n1, rows1 = formation_rows(f.formation, n1);
    if (d < rows1) formation_goal(f.formation, f.size, n1, c1, f.layer, c.cube_dim[1], perm1 ? perm1[d] : d, goal);
    else formation_goal(f.formation, f.size, n2, c2, f.layer, c.cube_dim[2], perm2 ? perm2[d - rows1] : d - rows1, goal);
}

__device__ __forceinline__ void scen_store_row(float *row, int scen, const Formation &f, const float *center, int ctl, int increase, float speed,
                                               const float *aux9)
{
    float4 *r4 = reinterpret_cast<float4 *>(row);
    r4[0] = make_float4((float)scen, (float)f.formation, f.size, f.layer);
    r4[1] = make_float4(f.highest, f.lowest, center[0], center[1]);
    r4[2] = make_float4(center[2], (float)ctl, (float)increase, speed);
    r4[3] = make_float4(aux9[0], aux9[1], aux9[2], aux9[3]);
    r4[4] = make_float4(aux9[4], aux9[5], aux9[6], aux9[7]);
    r4[5] = make_float4(aux9[8], 0.f, 0.f, 0.f);
}

// scenario.reset() of the non-obstacle env (quadrotor_multi.py:459) for lane d: this drone's goal (= its spawn point, :467-470)
static __device__ __noinline__ void formation_reset(const DevConst &c, const Rng g, int d, bool leader, float *row, float *goal)
{
    const int K = c.K;
    int scen = c.scenario;
    if (scen == QS_SCENARIO_MIX) {                                      // mix.py:79-99, mode lists utils.py:7-15
        const int n = (K == 1) ? 5 : 9;
        const int mi = min((int)floorf(rng_u(g, SITE_SCENARIO, 0xFF, 1, 0) * (float)n), n - 1);
        const int multi[9] = { QS_SCENARIO_STATIC_SAME_GOAL, QS_SCENARIO_STATIC_DIFF_GOAL, QS_SCENARIO_EP_LISSAJOUS3D, QS_SCENARIO_EP_RAND_BEZIER,
                               QS_SCENARIO_DYNAMIC_SAME_GOAL, QS_SCENARIO_DYNAMIC_DIFF_GOAL, QS_SCENARIO_DYNAMIC_FORMATIONS,
                               QS_SCENARIO_SWAP_GOALS, QS_SCENARIO_SWARM_VS_SWARM };
        scen = (K == 1 && mi == 4) ? QS_SCENARIO_DYNAMIC_SAME_GOAL : multi[mi];   // the single-drone list is the first four + dynamic_same_goal
    }
    float u0[4], u1[4], u2[4], aux9[9];
    rng_u4(g, SITE_SCENARIO, 0xFF, 7, 0, u0);                           // formation index, size, layer distance, duration
#pragma unroll
    for (int k = 0; k < 9; ++k) aux9[k] = 0.f;
    int ctl = 0, increase = 0;
    float speed = 0.f, center[3] = { 0.f, 0.f, 2.0f };
    bool shuffle = true;
    if (scen == QS_SCENARIO_DYNAMIC_SAME_GOAL || scen == QS_SCENARIO_DYNAMIC_DIFF_GOAL || scen == QS_SCENARIO_SWAP_GOALS ||
        scen == QS_SCENARIO_SWARM_VS_SWARM)                             // duration_time ~ U(4, 6) s, in double like the reference's int()
        ctl = (int)((4.0 + 2.0 * (double)u0[3]) * (double)c.control_freq);
    else if (scen == QS_SCENARIO_DYNAMIC_FORMATIONS) {                  // dynamic_formations.py:42-48
        rng_u4(g, SITE_SCENARIO, 0xFF, 7, 1, u1);
        increase = u1[0] < 0.5f; speed = 1.0f + 2.0f * u1[1];
    } else if (scen == QS_SCENARIO_EP_LISSAJOUS3D) { center[0] = -2.0f; shuffle = false; }   // ep_lissajous3D.py:30-38
    Formation f;
    scen_update_formation(c, scen, u0, f);
    if (scen == QS_SCENARIO_SWARM_VS_SWARM) {                           // formation_centers, swarm_vs_swarm.py:18-58
        rng_u4(g, SITE_SCENARIO, 0xFF, 7, 1, u1); rng_u4(g, SITE_SCENARIO, 0xFF, 7, 2, u2);
        const float box = c.spawn_box;
        float *c1 = aux9, *c2 = aux9 + 3;
        c1[0] = -box + 2.0f * box * u1[2]; c1[1] = -box + 2.0f * box * u1[3]; c1[2] = scen_z_value(c, f.formation, f.size, u2[0]);
        const float dist = box / 4 + (box - box / 4) * u2[1];
        float sp, cp, st, ct;
        sincospif(-1.0f + 2.0f * u2[2], &sp, &cp); sincospif(-0.5f + u2[3], &st, &ct);
        c2[0] = c1[0] + dist * (st * cp); c2[1] = c1[1] + dist * (st * sp); c2[2] = c1[2] + dist * ct;
        int axis = -1;
        if (f.formation == QF_CIRCLE_H || f.formation == QF_GRID_H) axis = 2;
        else if (f.formation == QF_CIRCLE_XZ || f.formation == QF_GRID_XZ) axis = 1;
        else if (f.formation == QF_CIRCLE_YZ || f.formation == QF_GRID_YZ) axis = 0;
        if (axis >= 0) {
            const float diff = c2[axis] - c1[axis];
            if (fabsf(diff) < f.lowest) c2[axis] = ((diff > 0.f) ? 1.f : ((diff < 0.f) ? -1.f : 0.f)) * f.lowest + c1[axis];
        }
        swarm_goal(c, f, c1, c2, d, nullptr, nullptr, goal);
#pragma unroll
        for (int a = 0; a < 3; ++a) center[a] = (c1[a] + c2[a]) / 2;
    } else {
        unsigned char perm[QS_MAX_AGENTS + 3];
        const int rows = formation_rows(f.formation, K);
        int r = d;
        if (shuffle) { scen_shuffle_perm(g, 8, rows, perm); r = perm[d]; }
        formation_goal(f.formation, f.size, K, center, (scen == QS_SCENARIO_EP_LISSAJOUS3D) ? 0.f : f.layer, c.cube_dim[0], r, goal);
    }
    if (leader) scen_store_row(row, scen, f, center, ctl, increase, speed, aux9);
}

// Goal changes that happen every 4-6 s (or every 5 s for the Bezier resampling): rare, out of line.
// Every lane of the group calls it; returns this lane's new goal.  `stage` is the warp's exchange buffer (swap_goals).
static __device__ __noinline__ float3 scenario_event(const DevConst &c, const Rng g, int scen, int d, bool leader, uint32_t gmask, int lane, int base,
                                                     float *row, float4 *stage, float3 goal_in)
{
    float goal[3] = { goal_in.x, goal_in.y, goal_in.z };                // by value: the caller's goal stays in registers
    const int K = c.K;
    const float box = c.spawn_box;
    float4 *r4 = reinterpret_cast<float4 *>(row);
    const float4 a0 = r4[0], a1 = r4[1], a2 = r4[2];
    Formation f; f.formation = (int)a0.y; f.size = a0.z; f.layer = a0.w; f.highest = a1.x; f.lowest = a1.y;
    float center[3] = { a1.z, a1.w, a2.x }, aux9[9];
    {
        const float4 b3 = r4[3], b4 = r4[4], b5 = r4[5];
        aux9[0] = b3.x; aux9[1] = b3.y; aux9[2] = b3.z; aux9[3] = b3.w; aux9[4] = b4.x; aux9[5] = b4.y; aux9[6] = b4.z; aux9[7] = b4.w; aux9[8] = b5.x;
    }
    const int ctl = (int)a2.y, increase = (int)a2.z;
    const float speed = a2.w;
    __syncwarp(gmask);                                                  // every lane has read the row before the leader rewrites it
    float u0[4], u1[4];
    unsigned char perm[QS_MAX_AGENTS + 3], perm2[QS_MAX_AGENTS + 3];
    if (scen == QS_SCENARIO_DYNAMIC_SAME_GOAL) {                        // dynamic_same_goal.py:18-31
        rng_u4(g, SITE_SCENARIO, 0xFF, 10, 0, u0); rng_u4(g, SITE_SCENARIO, 0xFF, 10, 1, u1);
        center[0] = -box + 2.0f * box * u0[3]; center[1] = -box + 2.0f * box * u1[0];
        center[2] = fmaxf(0.25f, (-0.5f * box + box * u1[1]) + 2.0f);
        formation_goal(f.formation, f.size, K, center, 0.f, c.cube_dim[0], d, goal);
    } else if (scen == QS_SCENARIO_DYNAMIC_DIFF_GOAL) {                 // dynamic_diff_goal.py:25-43
        rng_u4(g, SITE_SCENARIO, 0xFF, 10, 0, u0); rng_u4(g, SITE_SCENARIO, 0xFF, 10, 1, u1);
        center[0] = -box + 2.0f * box * u0[3]; center[1] = -box + 2.0f * box * u1[0];
        center[2] = scen_z_value(c, f.formation, f.size, u1[1]);        // with the OLD formation and size
        scen_update_formation(c, scen, u0, f);
        const int rows = formation_rows(f.formation, K);
        scen_shuffle_perm(g, 11, rows, perm);
        formation_goal(f.formation, f.size, K, center, f.layer, c.cube_dim[0], perm[d], goal);
    } else if (scen == QS_SCENARIO_SWAP_GOALS) {                        // swap_goals.py:14-27: permute the goals the drones hold
        scen_shuffle_perm(g, 11, K, perm);
        stage[2 * lane] = make_float4(goal[0], goal[1], goal[2], 0.f);
        __syncwarp(gmask);
        const float4 o = stage[2 * (base + perm[d < K ? d : 0])];
        goal[0] = o.x; goal[1] = o.y; goal[2] = o.z;
        __syncwarp(gmask);
    } else if (scen == QS_SCENARIO_RUN_AWAY) {                          // run_away.py:18-25: goals[0] <- goals[g0], goals[1] <- goals[g1], g in [1, K)
        rng_u4(g, SITE_SCENARIO, 0xFF, 10, 0, u0);
        const int g0 = min(1 + (int)floorf(u0[0] * (float)(K - 1)), K - 1), g1 = min(1 + (int)floorf(u0[1] * (float)(K - 1)), K - 1);
        stage[2 * lane] = make_float4(goal[0], goal[1], goal[2], 0.f);
        __syncwarp(gmask);
        if (d < 2) {                                                    // g1 >= 1: the second copy never reads the row the first one wrote
            const float4 o = stage[2 * (base + (d == 0 ? g0 : g1))];
            goal[0] = o.x; goal[1] = o.y; goal[2] = o.z;
        }
        __syncwarp(gmask);
    } else if (scen == QS_SCENARIO_SWARM_VS_SWARM) {                    // swarm_vs_swarm.py:67-90
#pragma unroll
        for (int a = 0; a < 3; ++a) { const float t = aux9[a]; aux9[a] = aux9[3 + a]; aux9[3 + a] = t; }
        rng_u4(g, SITE_SCENARIO, 0xFF, 10, 0, u0);
        scen_update_formation(c, scen, u0, f);
        const int n1 = K / 2, rows1 = formation_rows(f.formation, n1), rows2 = formation_rows(f.formation, K - n1);
        scen_shuffle_perm(g, 11, rows1, perm); scen_shuffle_perm(g, 12, rows2, perm2);
        swarm_goal(c, f, aux9, aux9 + 3, d, perm, perm2, goal);
    } else if (scen == QS_SCENARIO_EP_RAND_BEZIER) {                    // ep_rand_bezier.py:19-37: new control points inside the room
        const float rd[3] = { c.room_l - f.size, c.room_w - f.size, c.room_h - f.size };
        const float max_dist = fminf(30.0f, fmaxf(rd[0], fmaxf(rd[1], rd[2]))), min_dist = max_dist / 2;
        const float low[3] = { -rd[0] / 2, -rd[1] / 2, 0.f }, high[3] = { rd[0] / 2, rd[1] / 2, rd[2] };
        const int lo_i = (int)min_dist, hi_i = (int)(max_dist + 1.0f);
        bool found = false;
        for (int att = 0; att < QS_BEZIER_MAX_TRIES && !found; ++att) {
            rng_u4(g, SITE_SCENARIO, 0xFF, 13, 2 * att, u0); rng_u4(g, SITE_SCENARIO, 0xFF, 13, 2 * att + 1, u1);
            const float flat[6] = { -high[0] + 2.0f * high[0] * u0[0], -high[1] + 2.0f * high[1] * u0[1], -high[2] + 2.0f * high[2] * u0[2],
                                    -high[0] + 2.0f * high[0] * u0[3], -high[1] + 2.0f * high[1] * u1[0], -high[2] + 2.0f * high[2] * u1[1] };
            const float mag = (float)(lo_i + (int)floorf(u1[2] * (float)(hi_i - lo_i)));
            found = true;
            float np_[6];
#pragma unroll
            for (int col = 0; col < 2; ++col) {                         // (2,3).reshape(3,2): element [r][col] = flat[2r + col]
                const float nrm = sqrtf(flat[col] * flat[col] + flat[2 + col] * flat[2 + col] + flat[4 + col] * flat[4 + col]);
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    const float v = goal[r] + flat[2 * r + col] * mag / nrm;
                    np_[3 * col + r] = v;
                    if (!(v > low[r] + 0.5f) || !(v < high[r] - 0.5f)) found = false;
                }
            }
            if (found) {
#pragma unroll
                for (int k = 0; k < 6; ++k) aux9[3 + k] = np_[k];
            }
        }
#pragma unroll
        for (int r = 0; r < 3; ++r) { aux9[r] = goal[r]; if (!found) { aux9[3 + r] = goal[r]; aux9[6 + r] = goal[r]; } }
    }
    if (leader) scen_store_row(row, scen, f, center, ctl, increase, speed, aux9);
    return make_float3(goal[0], goal[1], goal[2]);
}

// scenario.step() (quadrotor_multi.py:701) with the already incremented tick.  Group-uniform control flow.
// `rs`: the env's scenario row as staged in shared memory at the top of the kernel (its global load travels with the state
// loads instead of being waited for here); `row`: the row in global memory, written by the leader lane when it changes.
template <int KG>
__device__ __forceinline__ void formation_scenario_step(const DevConst &c, const Rng &g, int d, bool leader, uint32_t gmask, int lane, int tick,
                                                        const float4 *rs, float *row, float4 *stage, float *goal)
{
    float4 *r4 = reinterpret_cast<float4 *>(row);
    const float4 a0 = rs[0];
    const int scen = (int)a0.x;
    if (scen == QS_SCENARIO_STATIC_SAME_GOAL || scen == QS_SCENARIO_STATIC_DIFF_GOAL) return;
    const int base = lane & ~(KG - 1);
    if (scen == QS_SCENARIO_DYNAMIC_FORMATIONS) {                       // dynamic_formations.py:22-40, every step
        const float4 a1 = rs[1], a2 = rs[2];
        float size = a0.z, speed = a2.w;
        int increase = (int)a2.z;
        if (size <= -a1.x) { increase = 1; speed = 1.0f + 2.0f * rng_u(g, SITE_SCENARIO, 0xFF, 10, 6); }
        else if (size >= a1.x) { increase = 0; speed = 1.0f + 2.0f * rng_u(g, SITE_SCENARIO, 0xFF, 10, 6); }
        size += increase ? 0.001f * speed : -0.001f * speed;
        const float center[3] = { a1.z, a1.w, a2.x };
        formation_goal((int)a0.y, size, c.K, center, a0.w, c.cube_dim[0], d, goal);
        if (leader) { r4[0] = make_float4(a0.x, a0.y, size, a0.w); r4[2] = make_float4(a2.x, a2.y, (float)increase, speed); }
    } else if (scen == QS_SCENARIO_EP_LISSAJOUS3D) {                    // ep_lissajous3D.py:9-25: a=0.03 b=c=0.01 n=m=2 phi=psi=90 rad
        const float t = (float)tick / c.control_freq;
        goal[0] += 0.03f * sinf(t); goal[1] += 0.01f * sinf(2.0f * t + 90.0f); goal[2] += 0.01f * cosf(2.0f * t + 90.0f);
    } else if (scen == QS_SCENARIO_EP_RAND_BEZIER) {                    // ep_rand_bezier.py:8-45
        const int control_steps = (int)(5.0f * c.control_freq), t = tick % control_steps;
        if (t == 0 || tick == 1) {
            const float3 ng = scenario_event(c, g, scen, d, leader, gmask, lane, base, row, stage, make_float3(goal[0], goal[1], goal[2]));
            goal[0] = ng.x; goal[1] = ng.y; goal[2] = ng.z;
        }
        if (t != 0 && tick > 1) {
            const float4 b3 = rs[3], b4 = rs[4], b5 = rs[5];
            const float u = (float)t / (float)(control_steps - 1), w0 = (1.0f - u) * (1.0f - u), w1 = 2.0f * (1.0f - u) * u, w2 = u * u;
            goal[0] = w0 * b3.x + w1 * b3.w + w2 * b4.z; goal[1] = w0 * b3.y + w1 * b4.x + w2 * b4.w; goal[2] = w0 * b3.z + w1 * b4.y + w2 * b5.x;
        }
    } else {                                                            // timer scenarios: goals change when tick % control_step_for_sec == 0
        const int ctl = (scen == QS_SCENARIO_RUN_AWAY) ? (int)c.control_freq : (int)rs[2].y;   // run_away.py:16: a local of step(), every second
        if (ctl > 0 && tick % ctl == 0 && tick > 0) {
            const float3 ng = scenario_event(c, g, scen, d, leader, gmask, lane, base, row, stage, make_float3(goal[0], goal[1], goal[2]));
            goal[0] = ng.x; goal[1] = ng.y; goal[2] = ng.z;
        }
    }
}

}  // namespace qs
