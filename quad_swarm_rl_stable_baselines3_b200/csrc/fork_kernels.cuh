// fork_kernels.cuh -- sm_100a device code of the fork env (QS_MODE_FORK): what swarm_rl/sb_train.py trains on.
//
// One launch = one VecEnv step = up to `substeps` (8) control steps of every env, each of them: the cascaded PID
// pre-controller + mixer (Controller/*.py), the same fused dynamics as the upstream path, collision / room
// bookkeeping, the capture reward and the evader's motion; then the 2-D observations, done, and the VecEnv worker's
// auto-reset.  The drone state, its 12 PIDs and the heading stay in registers across the 8 control steps, so HBM sees
// one read and one write of the state per 8 drone-steps (537 B per call = 67 B per drone-step).
// Same thread mapping as quadsim_kernels.cuh: one thread per drone, an env's K drones in KG adjacent lanes.
//
// Reference (paths relative to gym_art/quadrotor_multi/): quadrotor_multi_rewards.py:541-629,632-993;
// quadrotor_single_rewards.py:418-457,487-559; Controller/*.py; get_state.py:7-103; scenarios/dynamic_repulsive.py:41-82;
// swarm_rl/env_wrappers/subproc_vec_env_custom.py:35-52.
#pragma once
#include "quadsim_kernels.cuh"

namespace qs {

struct ForkConst {
    int substeps;
    float capture_radius, rew_existence, rew_captor, rew_helper, max_angular_rate, chaser_speed;
    float ev_vmax, ev_dt, ev_arena, spawn_ring, ev_rmin, ev_rspan;
    float pid[12][5];
    float rate_scale, mixer[4][4], mass, g, inv_kf4, min_rpm, inv_rpm_span, half_len;
    // camera model (simulate_camera_measurement_vect, quadrotor_multi_rewards.py:278-324)
    float cam_r, cam_noise_tan;      // target radius; pixel-noise std already scaled to tan units: noise_px * w / (resolution * f)
    float cam_seg, cam_inv_seg;      // 2 pi / n_cameras and its inverse
    int cam_num;
};

enum { FF_SUCCESS = 1, FF_PLACED = 2 };               // per-env fork flags (ForkPtrs::flags); planes: quadsim_kernels.cuh

// _pid_update_numba, Controller/Pid.py:7-26.  Branch-free: the host stores a disabled saturation as +inf and a disabled
// anti-windup limit as -1 (the open interval (-aw, aw) is then empty), so 12 PIDs x 8 sub-steps cost no branches.
__device__ __forceinline__ float pid_update(float &last, float &integ, const float *p, float err, float inv_dt, float dt)
{
    float diff = (err - last) * inv_dt;
    last = err;
    float out = p[0] * err + p[1] * diff + p[2] * integ;
    out = fminf(fmaxf(out, -p[3]), p[3]);
    integ = (-p[4] < out && out < p[4]) ? integ + err * dt : integ;
    return out;
}

__device__ __forceinline__ float wrap_pi(float a)
{
    const float two_pi = 6.283185307179586f;
    float r = a + QS_PI_F;
    r = r - two_pi * floorf(r / two_pi);          // python's float % for a positive divisor
    return r - QS_PI_F;
}

// |a + d| - |a| without cancellation: (2 a.d + d.d) / (|a + d| + |a|)
__device__ __forceinline__ float norm_increment(float ax, float ay, float dx, float dy, float na)
{
    float bx = ax + dx, by = ay + dy;
    float nb = sqrtf(bx * bx + by * by);
    float den = nb + na;
    return den > 0.f ? (2.0f * (ax * dx + ay * dy) + dx * dx + dy * dy) / den : 0.f;
}

// Controller.update_vel_height_dir, Controller/Controller.py:76-101 -> thrust commands in [0,1] after the fork's
// reorder / arctan squash (quadrotor_single_rewards.py:436-437) and CustomPidControl.step (quadrotor_control.py:90-94)
__device__ __forceinline__ void fork_controller(const DevConst &c, const ForkConst &f, Drone &q, float *pid, float &angle, float &ang_vel,
                                                float a0, float *cmd)
{
    const float dt = c.dt, inv_dt = 1.0f / c.dt;
    ang_vel = a0;
    angle = wrap_pi(angle + a0 * dt * f.max_angular_rate);
    float sa, ca;
    sincosf(angle, &sa, &ca);
    // position PIDs (x, y are overwritten below but their state still advances), PositionController.py:62-75
    float v0 = pid_update(pid[0], pid[1], f.pid[0], 0.f - q.p[0], inv_dt, dt);
    float v1 = pid_update(pid[2], pid[3], f.pid[1], 0.f - q.p[1], inv_dt, dt);
    float v2 = pid_update(pid[4], pid[5], f.pid[2], q.goal[2] - q.p[2], inv_dt, dt);
    v0 = ca * f.chaser_speed; v1 = sa * f.chaser_speed;
    // velocity PIDs, VelocityController.py:68-82
    float a_x = pid_update(pid[6], pid[7], f.pid[3], v0 - q.v[0], inv_dt, dt);
    float a_y = pid_update(pid[8], pid[9], f.pid[4], v1 - q.v[1], inv_dt, dt);
    float a_z = pid_update(pid[10], pid[11], f.pid[5], v2 - q.v[2], inv_dt, dt);
    // AccelerationController.get_control_signal (heading 0), AccelerationController.py:18-108
    float fx = a_x * f.mass, fy = a_y * f.mass, fz = (a_z + f.g) * f.mass;
    float fn = norm3f(fx, fy, fz), ifn = 1.0f / fn;
    float zx = fx * ifn, zy = fy * ifn, zz = fz * ifn;
    float A00 = 1.0f - zx * zx, A01 = -zx * zy, A10 = -zy * zx, A11 = 1.0f - zy * zy, A20 = -zz * zx, A21 = -zz * zy;
    float idet = 1.0f / (A00 * A11 - A01 * A10);
    float c0 = A11 * idet, c1 = -A10 * idet;                       // Bt_A2_inv @ [1, 0]
    float xx = A00 * c0 + A01 * c1, xy = A10 * c0 + A11 * c1, xz = A20 * c0 + A21 * c1;
    float ixn = 1.0f / norm3f(xx, xy, xz);
    xx *= ixn; xy *= ixn; xz *= ixn;
    float yx = zy * xz - zz * xy, yy = zz * xx - zx * xz, yz = zx * xy - zy * xx;
    float iyn = 1.0f / norm3f(yx, yy, yz);
    yx *= iyn; yy *= iyn; yz *= iyn;
    const float *R = q.R;
    float thrust_force = fmaxf(fx * R[2] + fy * R[5] + fz * R[8], 0.f);
    float throttle = clampf((sqrtf(thrust_force * f.inv_kf4) - f.min_rpm) * f.inv_rpm_span, 0.f, 1.f);
    // AttitudeController.get_control_signal, AttitudeController.py:60-82.  Rd columns = (x, y, z); M = Rd^T R
    float M01 = xx * R[1] + xy * R[4] + xz * R[7], M02 = xx * R[2] + xy * R[5] + xz * R[8];
    float M10 = yx * R[0] + yy * R[3] + yz * R[6], M12 = yx * R[2] + yy * R[5] + yz * R[8];
    float M20 = zx * R[0] + zy * R[3] + zz * R[6], M21 = zx * R[1] + zy * R[4] + zz * R[7];
    // E = (M - M^T)/2, vee / 2: ev0 = (E12 - E21)/2 = (M12 - M21)/2 ...
    float e0 = 0.5f * (M12 - M21), e1 = 0.5f * (M20 - M02), e2 = 0.5f * (M01 - M10);
    float r0 = pid_update(pid[12], pid[13], f.pid[6], e0, inv_dt, dt);
    float r1 = pid_update(pid[14], pid[15], f.pid[7], e1, inv_dt, dt);
    float r2 = pid_update(pid[16], pid[17], f.pid[8], e2, inv_dt, dt);
    // RateController.get_control_signal, RateController.py:71-89
    float g0 = pid_update(pid[18], pid[19], f.pid[9], r0 - q.w[0], inv_dt, dt) * f.rate_scale;
    float g1 = pid_update(pid[20], pid[21], f.pid[10], r1 - q.w[1], inv_dt, dt) * f.rate_scale;
    float g2 = pid_update(pid[22], pid[23], f.pid[11], r2 - q.w[2], inv_dt, dt) * f.rate_scale;
    // Mixer.get_control_signal, Mixer.py:69-111
    float m[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) m[k] = f.mixer[k][0] * g0 + f.mixer[k][1] * g1 + f.mixer[k][2] * g2 + f.mixer[k][3] * throttle;
    float mn = fminf(fminf(m[0], m[1]), fminf(m[2], m[3]));
    const float lift = fmaxf(-mn, 0.f);                                  // if mn < 0: motors += |mn|
#pragma unroll
    for (int k = 0; k < 4; ++k) m[k] += lift;
    float mx = fmaxf(fmaxf(m[0], m[1]), fmaxf(m[2], m[3]));
    if (mx > 1.0f) {
        if (throttle > 1e-2f) {
            float iscale = throttle / (0.25f * (m[0] + m[1] + m[2] + m[3]));
            g0 *= iscale; g1 *= iscale; g2 *= iscale;
#pragma unroll
            for (int k = 0; k < 4; ++k) m[k] = f.mixer[k][0] * g0 + f.mixer[k][1] * g1 + f.mixer[k][2] * g2 + f.mixer[k][3] * throttle;
        } else {
            float imx = 1.0f / mx;
#pragma unroll
            for (int k = 0; k < 4; ++k) m[k] *= imx;
        }
    }
    const float u[4] = { m[0] * 2.f - 1.f, m[3] * 2.f - 1.f, m[1] * 2.f - 1.f, m[2] * 2.f - 1.f };
#pragma unroll
    for (int k = 0; k < 4; ++k) cmd[k] = 0.5f * (clampf(atanf(u[k]), -1.0f, 1.0f) + 1.0f);
}

// state_{cdist_cdistdot,aw_awdot}_dist_distdot_{angle,sangle}_angledot, get_state.py:7-103
__device__ __forceinline__ void camera_measure(const ForkConst &f, float rx, float ry, float global_angle, float n1, float n2, float &dist, float &angle_rel);

__device__ __forceinline__ void fork_self_obs(const DevConst &c, const ForkConst &f, const Rng &g, int site, int drone, const Drone &q, float angle,
                                              float ang_vel, float *o)
{
    float p0 = q.p[0], p1 = q.p[1], v0 = q.v[0], v1 = q.v[1];
    if (c.sense_noise) {
        const float4 n = rng_n4v(g, site, drone, 0, 0), m = rng_n4v(g, site, drone, 0, 1);
        p0 += c.s_pos * n.x; p1 += c.s_pos * n.y;
        v0 += c.s_vel * n.w; v1 += c.s_vel * m.x;
    }
    const float dt = c.dt, inv_dt = 1.0f / c.dt;
    float rx = q.goal[0] - p0, ry = q.goal[1] - p1;
    float rel_dist = sqrtf(rx * rx + ry * ry);
    float dot_rel = norm_increment(rx, ry, v0 * dt, v1 * dt, rel_dist) * inv_dt;      // (|rel + v dt| - |rel|)/dt  [sic]
    float rel_angle = wrap_pi(atan2f(ry, rx) - angle);
    float cdist = sqrtf(p0 * p0 + p1 * p1);
    float cdistdot = norm_increment(p0, p1, v0 * dt, v1 * dt, cdist) * inv_dt;
    float prod = ang_vel * rel_angle;
    float adot = (prod > 0.f) ? -fabsf(ang_vel) : ((prod < 0.f) ? fabsf(ang_vel) : 0.f * fabsf(ang_vel));
    if (c.obs_repr == QS_OBS_AW_AWDOT_DIST_DISTDOT_ANGLE_ANGLEDOT) { o[0] = angle; o[1] = ang_vel; }
    else { o[0] = cdist; o[1] = cdistdot; }
    o[2] = rel_dist; o[3] = dot_rel;
    if (c.obs_repr == QS_OBS_CDIST_CDISTDOT_DIST_DISTDOT_SANGLE_ANGLEDOT) { float s, cc; sincosf(rel_angle, &s, &cc); o[4] = cc; o[5] = s; o[6] = adot; }
    else if (c.obs_repr == QS_OBS_CDIST_CDISTDOT_NDIST_DISTDOT_NSANGLE_ANGLEDOT) {
        // get_state.py:190-224: the goal as the camera model sees it from the (noisy) own position -- distance clipped to 0..10, cos / sin of
        // the measured bearing; the sign of angledot still comes from the exact relative angle.  Own noise stream: aux 0xF0 + site.
        const float4 n = rng_n4v(g, SITE_CAMERA, drone, 0xF0 + site, 0);
        float nd, na, s, cc;
        camera_measure(f, rx, ry, angle, n.x, n.y, nd, na);
        sincosf(na, &s, &cc);
        o[2] = clampf(nd, 0.f, 10.f); o[4] = cc; o[5] = s; o[6] = adot;
    } else { o[4] = rel_angle; o[5] = adot; }
}

// simulate_camera_measurement_vect (quadrotor_multi_rewards.py:278-324): the two tangent rays from the camera to a disc of
// radius r around the neighbour, seen by the best of n cameras, plus pixel noise.  The reference intersects the disc with the
// circle whose diameter is [camera, centre]; with c2 = c/2 and r2 = d = |c|/2 its `a` is exactly r^2 / (2 d), which is what is
// evaluated here (the literal r1^2 - r2^2 + d^2 cancels catastrophically in fp32).  NaN (target closer than its radius) -> 0.
__device__ __forceinline__ void camera_measure(const ForkConst &f, float rx, float ry, float global_angle, float n1, float n2,
                                               float &dist, float &angle_rel)
{
    float sg, cg;
    sincosf(global_angle, &sg, &cg);
    const float px = cg * rx + sg * ry, py = -sg * rx + cg * ry;         // R(-global_angle) rel_pos
    const float two_pi = 6.283185307179586f;
    float ao = atan2f(py, px);
    ao = ao - two_pi * floorf(ao / two_pi);                             // python %
    int cam_idx = (int)rintf(ao * f.cam_inv_seg) % f.cam_num;           // np.round: half to even
    const float cam = (float)cam_idx * f.cam_seg;
    float sc, cc;
    sincosf(cam, &sc, &cc);
    const float cx = cc * px + sc * py, cy = -sc * px + cc * py;
    const float r = f.cam_r;
    const float d = 0.5f * sqrtf(cx * cx + cy * cy);
    const float a = r * r / (2.0f * d);
    const float h = sqrtf(r * r - a * a);
    const float ux = -0.5f * cx / d, uy = -0.5f * cy / d;               // (c2 - c1) / d
    const float mx = cx + a * ux, my = cy + a * uy;
    const float x1x = mx - h * uy, x1y = my + h * ux, x2x = mx + h * uy, x2y = my - h * ux;
    const float t1 = x1y / x1x + f.cam_noise_tan * n1, t2 = x2y / x2x + f.cam_noise_tan * n2;   // u / f
    const float a1 = atanf(t1), a2 = atanf(t2);
    float l = r / sinf(0.5f * fabsf(a1 - a2));
    float ar = wrap_pi(0.5f * (a1 + a2) + cam);
    dist = isnan(l) ? 0.f : l;
    angle_rel = isnan(ar) ? 0.f : ar;
}

__device__ __forceinline__ int fork_nbr_width(int type)
{
    return (type == QS_NEIGHBOR_DIST_ANGLE) ? 2 : ((type == QS_NEIGHBOR_DIST_SANGLE || type == QS_NEIGHBOR_DIST_ANGLE_HEADING ||
            type == QS_NEIGHBOR_NDIST_NSANGLE) ? 3 : ((type == QS_NEIGHBOR_DIST_SANGLE_SHEADING) ? 5 : 0));
}

// one feature row of get_rel_pos_vel_item (quadrotor_multi_rewards.py:326-420) for neighbour j (position + heading snapshot `nb`).
// call 0: rows of the observation; call 1: the evaluation neighborhood_indices makes for the ranking (own camera noise).
__device__ __forceinline__ void fork_nbr_row(const DevConst &c, const ForkConst &f, const Rng &g, int d, int j, int call, const Drone &q,
                                             float angle, float hsnap, float4 nb, float *row)
{
    const float dx = nb.x - q.p[0], dy = nb.y - q.p[1], dz = nb.z - q.p[2];
    const float dist = norm3f(dx, dy, dz);
    row[0] = dist; row[1] = 0.f; row[2] = 0.f; row[3] = 0.f; row[4] = 0.f;
    if (c.nbr_type == QS_NEIGHBOR_NDIST_NSANGLE) {
        const float4 n = rng_n4v(g, SITE_CAMERA, d, j, call);
        float nd, na;
        camera_measure(f, dx, dy, angle, n.x, n.y, nd, na);
        row[0] = clampf(nd, 0.f, 10.f);
        sincosf(na, &row[2], &row[1]);
        return;
    }
    const float ang = wrap_pi(atan2f(dy, dx) - angle);
    const float hd = wrap_pi(nb.w - hsnap);
    if (c.nbr_type == QS_NEIGHBOR_DIST_ANGLE) row[1] = ang;
    else if (c.nbr_type == QS_NEIGHBOR_DIST_ANGLE_HEADING) { row[1] = ang; row[2] = hd; }
    else {
        sincosf(ang, &row[2], &row[1]);
        if (c.nbr_type == QS_NEIGHBOR_DIST_SANGLE_SHEADING) sincosf(hd, &row[4], &row[3]);
    }
}

// get_rel_pos_vel_item / neighborhood_indices / extend_obs_space, quadrotor_multi_rewards.py:326-476
template <int KG>
__device__ __forceinline__ void fork_neighbor_obs(const DevConst &c, const ForkConst &f, const Rng &g, int d, int lane, uint32_t gmask,
                                                  bool valid, const Drone &q, float angle, float hsnap, float *o, float4 *stage)
{
    const int W = fork_nbr_width(c.nbr_type);
    if (W == 0 || c.V <= 0 || KG == 1) return;
    const int base = lane & ~(KG - 1);
    __syncwarp(gmask);
    stage[2 * lane] = make_float4(q.p[0], q.p[1], q.p[2], hsnap);
    __syncwarp(gmask);
    if (!valid) return;
    const float INF = __int_as_float(0x7f800000);
    float met[KG];
    const bool ranked = c.V < c.K - 1;
    if (ranked) {
#pragma unroll 1
        for (int j = 0; j < KG; ++j) {
            float row[5], m = INF;
            if (j < c.K && j != d) {
                fork_nbr_row(c, f, g, d, j, 1, q, angle, hsnap, stage[2 * (base + j)], row);
                m = fmaxf(row[0] * row[0] + row[1] * row[1] + row[2] * row[2] + row[3] * row[3] + row[4] * row[4], 1.0e-4f);   // squared metric: same order
            }
#pragma unroll
            for (int k = 0; k < KG; ++k) if (k == j) met[k] = m;
        }
    }
#pragma unroll 1
    for (int j = 0; j < c.K; ++j) {
        if (j == d) continue;
        int slot;
        if (ranked) {
            float mj = INF;
#pragma unroll
            for (int k = 0; k < KG; ++k) if (k == j) mj = met[k];
            slot = 0;                                                  // stable rank among the candidates
#pragma unroll
            for (int k = 0; k < KG; ++k) slot += (met[k] < mj || (met[k] == mj && k < j)) ? 1 : 0;
        } else slot = j - (j > d ? 1 : 0);                             // all others in index order
        if (slot < c.V) {
            float row[5];
            fork_nbr_row(c, f, g, d, j, 0, q, angle, hsnap, stage[2 * (base + j)], row);
            float *r = o + W * slot;
            r[0] = clampf(row[0], -f.half_len, f.half_len);
            if (c.nbr_type == QS_NEIGHBOR_DIST_ANGLE || c.nbr_type == QS_NEIGHBOR_DIST_ANGLE_HEADING) {
                r[1] = clampf(row[1], -QS_PI_F, QS_PI_F);
                if (W == 3) r[2] = clampf(row[2], -QS_PI_F, QS_PI_F);
            } else {
                for (int k = 1; k < W; ++k) r[k] = clampf(row[k], -1.f, 1.f);
            }
        }
    }
}

template <int KG>
__device__ __forceinline__ float group_sum(float v, uint32_t gmask)
{
#pragma unroll
    for (int off = KG / 2; off > 0; off >>= 1) v += __shfl_xor_sync(gmask, v, off);
    return v;
}

// Scenario_dynamic_repulsive.step, scenarios/dynamic_repulsive.py:41-64: (fx, fy) = sum over the chasers of r / |r|^2
__device__ __forceinline__ void evader_advance(const ForkConst &f, float fx, float fy, float &ex, float &ey)
{
    float de = sqrtf(ex * ex + ey * ey);
    float iden = 1.0f / (de * fmaxf(f.ev_arena - de, 0.1f));
    float vx = fx - ex * iden, vy = fy - ey * iden;
    float vs = sqrtf(vx * vx + vy * vy), k = fminf(vs, f.ev_vmax) / vs * f.ev_dt;
    ex += vx * k; ey += vy * k;
}
// every lane of the group ends with the same evader; `mask` must name converged lanes (callers pass the full warp)
template <int KG>
__device__ __forceinline__ void evader_step(const ForkConst &f, const Drone &q, bool valid, bool placed, uint32_t mask, float &ex, float &ey)
{
    float fx = 0.f, fy = 0.f;
    if (valid && placed) {
        float rx = ex - q.p[0], ry = ey - q.p[1];
        float id2 = 1.0f / (rx * rx + ry * ry);
        fx = rx * id2; fy = ry * id2;
    }
    fx = group_sum<KG>(fx, mask); fy = group_sum<KG>(fy, mask);
    evader_advance(f, fx, fy, ex, ey);
}

// The fork step kernel.
template <int KG>
__global__ void __launch_bounds__(128, 2) fork_step_kernel(const __grid_constant__ DevConst c, const __grid_constant__ ForkConst f,
                                                           const __grid_constant__ DevPtrs P, const __grid_constant__ ForkPtrs F,
                                                           const float2 *__restrict__ actions, float *__restrict__ obs,
                                                           float *__restrict__ rew, uint8_t *__restrict__ done,
                                                           float *__restrict__ term_obs, uint8_t *__restrict__ reset_success)
{
    extern __shared__ __align__(16) float smem[];
    const int lane = threadIdx.x & 31, warp_in_block = threadIdx.x >> 5;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int env = tid / KG, d = tid % KG;
    const bool valid = env < c.N && d < c.K;
    const bool env_ok = env < c.N;
    const int gi = env * c.K + d;
    const uint32_t gmask = group_mask<KG>(lane);
    const int base = lane & ~(KG - 1);
    constexpr int GPW = 32 / KG;
    const int rows_per_warp = GPW * c.K;
    float4 *stage = reinterpret_cast<float4 *>(smem) + (size_t)warp_in_block * 64;
    float *tile = smem + (size_t)(blockDim.x >> 5) * 256 + (size_t)warp_in_block * rows_per_warp * c.D;
    const int row = (lane / KG) * c.K + d;
    float *orow = tile + (size_t)row * c.D;
    const int warp_env0 = (tid - lane) / KG;
    const int warp_rows = max(0, min(GPW, c.N - warp_env0)) * c.K;

    Drone q;
    float pid[24], angle = 0.f, ang_vel = 0.f, hsnap = 0.f, ex = 1.f, ey = 0.f;
    float2 act = make_float2(0.f, 0.f);
    int tick = 0, svd = 0, fflags = 0;
    Rng g; g.k0 = c.key0; g.k1 = c.key1; g.gid = 0; g.step = c.rng_step;      // sub-step s of this call draws from counter rng_step + s
    if (env_ok) {
        tick = P.tick[env]; svd = P.svd_ctr[env];
        g.gid = (uint32_t)(c.env_id_offset + env);
        float2 e2 = F.evader[env]; ex = e2.x; ey = e2.y;
        fflags = F.flags[env];
    }
    if (valid) {
        load_drone(P, gi, q);
        act = actions[gi];
#pragma unroll
        for (int k = 0; k < 6; ++k) { float4 v = F.plane[FP_PID0 + k][gi]; pid[4 * k] = v.x; pid[4 * k + 1] = v.y; pid[4 * k + 2] = v.z; pid[4 * k + 3] = v.w; }
        float4 h = F.plane[FP_HEADING][gi]; angle = h.x; ang_vel = h.y; hsnap = h.z;
    } else {
#pragma unroll
        for (int a = 0; a < 3; ++a) { q.p[a] = 1.0e6f * (float)(lane + 1); q.v[a] = 0.f; q.w[a] = 0.f; q.goal[a] = 0.f; }
#pragma unroll
        for (int a = 0; a < 9; ++a) q.R[a] = (a % 4 == 0) ? 1.f : 0.f;
#pragma unroll
        for (int a = 0; a < 4; ++a) { q.rd[a] = q.cd[a] = q.ou[a] = 0.f; }
#pragma unroll
        for (int a = 0; a < 24; ++a) pid[a] = 0.f;
        q.flags = 0; q.colmask = 0; q.goal[2] = 2.f;
    }

    int ec[EC_COUNT];
#pragma unroll
    for (int k = 0; k < EC_COUNT; ++k) ec[k] = 0;                       // this launch's increments (leader lane)
    float reward = 0.f;
    bool any_done = false, bad_any = false;
    // The sub-step loop is WARP-uniform: env groups that finish early (capture / timeout) idle under a predicate instead of
    // leaving the loop, so every collective below runs once per warp on the full mask.  (Group-masked collectives inside a
    // group-divergent loop are serialised per group by the compiler: 8 passes per warp for K = 4 -- that was 14 % of the
    // stall samples of the first version, profiles/r1_v4_fork_step_kernel_ncu_full.json.)
    bool grp_done = false;
    for (int sub = 0; sub < f.substeps; ++sub) {
        const bool gact = !grp_done, lact = valid && gact;
        const int time_remain = c.ep_len - tick;
        // ---- QuadrotorSingle._step, quadrotor_single_rewards.py:418-457
        if (lact) {
            float cmd[4];
            fork_controller(c, f, q, pid, angle, ang_vel, act.x, cmd);
            hsnap = angle;                                               // self.heading[i] = pre_controller.angle (:647)
            const float4 nv = rng_n4v(g, SITE_OU, d, 0, 0);
            const float n[4] = { nv.x, nv.y, nv.z, nv.w };
#pragma unroll
            for (int m = 0; m < 4; ++m) q.ou[m] = q.ou[m] + (c.ou_theta * (0.0f - q.ou[m]) + c.ou_sigma * n[m]);
            for (int s = 0; s < c.sim_steps; ++s) {
                svd += 1;
                bool fire = svd >= c.svd_period;
                if (fire) svd = 0;
                dynamics_substep(c, g, d, q, cmd, s, fire);
            }
        } else if (gact) {
            for (int s = 0; s < c.sim_steps; ++s) { svd += 1; if (svd >= c.svd_period) svd = 0; }
        }
        if (gact) tick += 1;
        const bool timeout = tick > c.ep_len;
        float chk = q.p[0] + q.p[1] + q.p[2] + q.v[0] + q.v[1] + q.v[2] + q.w[0] + q.w[1] + q.w[2] + q.R[0] + q.R[4] + q.R[8] + pid[0] + pid[12] + pid[18];
        const bool bad = lact && !isfinite(chk);
        const bool bad_now = (__ballot_sync(QS_FULL, bad) & gmask) != 0u;
        // ---- drone-drone collision bookkeeping (counters only: rewards and impulses are off in the fork, :775-786,818)
        uint32_t rowmask = 0u;
        if (KG > 1) {
            __syncwarp();
            stage[2 * lane] = make_float4(q.p[0], q.p[1], q.p[2], 0.f);
            __syncwarp();
            const float col2 = c.thr_col * c.thr_col * 1.0001f;
#pragma unroll
            for (int j = 0; j < KG; ++j) {
                float4 o4 = stage[2 * (base + j)];
                float dx = q.p[0] - o4.x, dy = q.p[1] - o4.y, dz = q.p[2] - o4.z, d2 = dx * dx + dy * dy + dz * dz;
                if ((j != d) && (j < c.K) && lact && d2 <= col2 && __fsqrt_rn(d2) <= c.thr_col) rowmask |= 1u << j;
            }
        }
        const bool is_unique = lact && (rowmask != 0u) && (q.colmask == 0u);
        const int col_tick = __popc(__ballot_sync(QS_FULL, is_unique) & gmask) >> 1;
        const bool settled = (float)tick >= c.grace_steps;
        // ---- room bookkeeping (:715-721, 760-764)
        const bool new_wall = (q.flags & F_CR_WALL) && !(q.flags & F_PREV_WALL);
        const bool new_ceil = (q.flags & F_CR_CEIL) && !(q.flags & F_PREV_CEIL);
        const bool cr_floor = (q.flags & F_CR_FLOOR) != 0;
        const bool new_room = (cr_floor || new_wall || new_ceil) && !(q.flags & F_PREV_ROOM);
        const uint32_t wall_b = __ballot_sync(QS_FULL, new_wall && lact) & gmask, ceil_b = __ballot_sync(QS_FULL, new_ceil && lact) & gmask;
        const uint32_t floor_b = __ballot_sync(QS_FULL, cr_floor && lact) & gmask, room_b = __ballot_sync(QS_FULL, new_room && lact) & gmask;
        // ---- capture reward (:733-758): distance to envs[0].goal = the evader before this sub-step's scenario.step
        const float rdx = ex - q.p[0], rdy = ey - q.p[1], rd = __fsqrt_rn(rdx * rdx + rdy * rdy);
        const bool cap = lact && (f.capture_radius > rd);
        const bool any_cap = (__ballot_sync(QS_FULL, cap) & gmask) != 0u;
        // the evader's push from the chasers of this group (scenario.step, :848) -- reduced on the full mask as well
        float efx = 0.f, efy = 0.f;
        if (lact && (fflags & FF_PLACED)) { const float id2 = 1.0f / (rdx * rdx + rdy * rdy); efx = rdx * id2; efy = rdy * id2; }
        efx = group_sum<KG>(efx, QS_FULL); efy = group_sum<KG>(efy, QS_FULL);
        if (gact) {
            if (col_tick > 0 && settled && is_unique) q.flags |= F_COL_AGENT;
            q.colmask = rowmask;
            q.flags = (q.flags & ~(F_PREV_WALL | F_PREV_CEIL | F_PREV_ROOM)) | (new_wall ? F_PREV_WALL : 0) | (new_ceil ? F_PREV_CEIL : 0) |
                      (new_room ? F_PREV_ROOM : 0);
            ec[EC_COL] += col_tick;
            if (settled) { ec[EC_COL_SETTLE] += col_tick; ec[EC_ROOM] += __popc(room_b); ec[EC_FLOOR] += __popc(floor_b); ec[EC_WALL] += __popc(wall_b); ec[EC_CEIL] += __popc(ceil_b); }
            if ((float)time_remain <= c.final_grace_steps) ec[EC_COL_FINAL] += col_tick;
            reward = f.rew_existence;                                    // the list is rebuilt every sub-step (:634)
            if (any_cap) { reward += cap ? f.rew_captor : 0.f; reward += (f.capture_radius < rd) ? f.rew_helper : 0.f; fflags |= FF_SUCCESS; }
            bad_any = bad_now;
            any_done = any_cap || timeout || bad_now;
            // ---- self observation of the last executed sub-step (goal = evader before scenario.step)
            if (valid && (any_done || sub == f.substeps - 1)) {
                fork_self_obs(c, f, g, SITE_SENSOR, d, q, angle, ang_vel, orow);
                // infos[i]['goal_dist'] of the last executed sub-step (quadrotor_single_rewards.py:457): 3-D distance to self.goal
                if (__builtin_expect(P.rew_info != nullptr, 0))
                    P.rew_info[2 * gi] = make_float4(norm3f(q.p[0] - q.goal[0], q.p[1] - q.goal[1], q.p[2] - q.goal[2]), 0.f, 0.f, 0.f);
            }
            // ---- scenario.step (:848)
            evader_advance(f, efx, efy, ex, ey);
            q.goal[0] = ex; q.goal[1] = ey; q.goal[2] = 2.0f;
            grp_done = any_done;
            if (!any_done && sub + 1 < f.substeps) g.step += 1u;         // next control step -> next RNG counter
        }
        if (__all_sync(QS_FULL, grp_done)) break;                        // warp-uniform exit
    }
    // ---- neighbour observations once, after the sub-steps (:990-991)
    fork_neighbor_obs<KG>(c, f, g, d, lane, QS_FULL, valid, q, angle, hsnap, orow + c.S, stage);
    if (valid) { rew[gi] = reward; done[gi] = any_done ? 1 : 0; }

    __syncwarp();
    const uint32_t done_ballot = __ballot_sync(QS_FULL, any_done && valid);
    if (done_ballot) {
        if (term_obs != nullptr) {
            for (int r = 0; r < warp_rows; ++r) {
                int src_lane = (r / c.K) * KG;
                if ((done_ballot >> src_lane) & 1u) {
                    const float *s = tile + (size_t)r * c.D;
                    float *t = term_obs + ((size_t)warp_env0 * c.K + r) * c.D;
                    for (int k = lane; k < c.D; k += 32) t[k] = s[k];
                }
            }
        }
        __syncwarp();
    }
    // ---- episode counters: accumulate, and on done flush to the aggregate and reset (the VecEnv worker's env.reset())
    uint32_t b_col = 0u;
    if (any_done) b_col = __ballot_sync(gmask, valid && (q.flags & F_COL_AGENT)) & gmask;
    if (env_ok && d == 0) {
        int *pe = P.ecnt + env * EC_COUNT;
        if (any_done) {
            qs_stats *st = P.stats;
            atomicAdd((unsigned long long *)&st->episodes, 1ull);
            atomicAdd((unsigned long long *)&st->episodes_success, (fflags & FF_SUCCESS) ? 1ull : 0ull);
            atomicAdd((unsigned long long *)&st->num_collisions, (unsigned long long)(pe[EC_COL] + ec[EC_COL]));
            atomicAdd((unsigned long long *)&st->num_collisions_after_settle, (unsigned long long)(pe[EC_COL_SETTLE] + ec[EC_COL_SETTLE]));
            atomicAdd((unsigned long long *)&st->num_collisions_final_5s, (unsigned long long)(pe[EC_COL_FINAL] + ec[EC_COL_FINAL]));
            atomicAdd((unsigned long long *)&st->num_collisions_with_room, (unsigned long long)(pe[EC_ROOM] + ec[EC_ROOM]));
            atomicAdd((unsigned long long *)&st->num_collisions_with_floor, (unsigned long long)(pe[EC_FLOOR] + ec[EC_FLOOR]));
            atomicAdd((unsigned long long *)&st->num_collisions_with_wall, (unsigned long long)(pe[EC_WALL] + ec[EC_WALL]));
            atomicAdd((unsigned long long *)&st->num_collisions_with_ceiling, (unsigned long long)(pe[EC_CEIL] + ec[EC_CEIL]));
            atomicAdd((unsigned long long *)&st->agents_collided, (unsigned long long)__popc(b_col));
            if (bad_any) atomicAdd((unsigned long long *)&st->nonfinite_resets, 1ull);
            // per-episode record (infos[i]['episode_extra_stats'], quadrotor_multi_rewards.py:886-978): the fork env never marks a
            // goal as reached (its distance log is commented out, :797), so every collision-free agent counts as deadlocked
            int *er = P.ep_rec + (size_t)env * QS_ER_COUNT;
            er[QS_ER_SEQ] += 1;
            er[QS_ER_SCENARIO] = QS_SCENARIO_DYNAMIC_REPULSIVE;
#pragma unroll
            for (int k = 0; k < 9; ++k) er[QS_ER_NUM_COLLISIONS + k] = pe[k] + ec[k];
            er[QS_ER_AGENTS_SUCCESS] = 0; er[QS_ER_AGENTS_DEADLOCK] = c.K - __popc(b_col); er[QS_ER_AGENTS_COLLIDED] = __popc(b_col);
            er[QS_ER_AGENTS_NEIGHBOR_COL] = __popc(b_col); er[QS_ER_AGENTS_OBST_COL] = 0;
            er[QS_ER_EP_LEN] = tick; er[QS_ER_SUCCESS] = (fflags & FF_SUCCESS) ? 1 : 0; er[QS_ER_NONFINITE] = bad_any ? 1 : 0;
#pragma unroll
            for (int k = 0; k < EC_COUNT; ++k) pe[k] = 0;
            if (reset_success != nullptr) reset_success[env] = (fflags & FF_SUCCESS) ? 1 : 0;
        } else {
#pragma unroll
            for (int k = 0; k < EC_COUNT; ++k) if (ec[k]) pe[k] += ec[k];
        }
    }
    if (any_done) {
        if (valid) P.ep_agent[gi] = make_float4(__int_as_float(0x7fc00000), __int_as_float(0x7fc00000), __int_as_float(0x7fc00000), 0.f);
        // ---- QuadrotorEnvMulti.reset (quadrotor_multi_rewards.py:541-629) + Scenario_dynamic_repulsive.reset (:66-82)
        if (bad_any) {
#pragma unroll
            for (int m = 0; m < 4; ++m) q.ou[m] = isfinite(q.ou[m]) ? q.ou[m] : 0.f;
#pragma unroll
            for (int a = 0; a < 24; ++a) pid[a] = isfinite(pid[a]) ? pid[a] : 0.f;
            if (!isfinite(q.p[0] + q.p[1])) { q.p[0] = 0.f; q.p[1] = 0.f; }
            if (!isfinite(ang_vel)) ang_vel = 0.f;
        }
        float u[4], t[4];
        rng_u4(g, SITE_SCENARIO, 0xFF, 7, d >> 1, u);                    // draws 2d, 2d+1 of the spawn-direction stream
        float sa_ = u[(2 * d) & 3] - 0.5f, sb_ = u[(2 * d + 1) & 3] - 0.5f, sn_ = sqrtf(sa_ * sa_ + sb_ * sb_);
        rng_u4(g, SITE_SCENARIO, 0xFF, 8, 0, t);
        const float ring = t[0] * f.spawn_ring;
        rng_u4(g, SITE_SCENARIO, 0xFF, 9, 0, t);
        float ea = t[0] - 0.5f, eb = t[1] - 0.5f, en = sqrtf(ea * ea + eb * eb), rad = t[2] * f.ev_rspan + f.ev_rmin;
        ex = ea / en * rad; ey = eb / en * rad;
        evader_step<KG>(f, q, valid, (fflags & FF_PLACED) != 0, gmask, ex, ey);   // one evader step against the OLD chaser positions (:82)
        q.goal[0] = ex; q.goal[1] = ey; q.goal[2] = 2.0f;
        rng_u4(g, SITE_SPAWN, d, 2, 0, u);
        angle = (u[0] - 0.5f) * 2.0f * QS_PI_F;                          // pre_controller.angle, :575
        q.p[0] = sa_ / sn_ * ring; q.p[1] = sb_ / sn_ * ring; q.p[2] = fmaxf(q.goal[2], c.spawn_min_z);
        rng_u4(g, SITE_SPAWN, d, 1, 0, u);
        float sy, cy;
        sincospif(-1.0f + 2.0f * u[0], &sy, &cy);                        // randyaw(), no rejection loop (quadrotor_single_rewards.py:540-543)
        set_yaw(q.R, cy, sy);
        q.v[0] = q.v[1] = q.v[2] = 0.f; q.w[0] = q.w[1] = q.w[2] = 0.f;
#pragma unroll
        for (int m = 0; m < 4; ++m) { q.rd[m] = 0.f; q.cd[m] = 0.f; }
        q.flags = 0; q.colmask = 0u;
        tick = 0;
        fflags = FF_PLACED;
        if (valid) fork_self_obs(c, f, g, SITE_SENSOR_RESET, d, q, angle, ang_vel, orow);
        fork_neighbor_obs<KG>(c, f, g, d, lane, gmask, valid, q, angle, hsnap, orow + c.S, stage);   // self.heading is NOT refreshed by reset()
    }
    // ---- write back
    if (valid) {
        store_drone(P, gi, q, true);
#pragma unroll
        for (int k = 0; k < 6; ++k) F.plane[FP_PID0 + k][gi] = make_float4(pid[4 * k], pid[4 * k + 1], pid[4 * k + 2], pid[4 * k + 3]);
        F.plane[FP_HEADING][gi] = make_float4(angle, ang_vel, hsnap, 0.f);
    }
    if (env_ok && d == 0) {
        P.tick[env] = tick; P.svd_ctr[env] = svd;
        F.evader[env] = make_float2(ex, ey); F.flags[env] = fflags;
    }
    __syncwarp();
    if (warp_rows > 0) warp_store_tile(tile, obs + (size_t)warp_env0 * c.K * c.D, warp_rows * c.D, lane);
}

// explicit reset of all / masked envs in fork mode
template <int KG>
__global__ void __launch_bounds__(128, 2) fork_reset_kernel(const __grid_constant__ DevConst c, const __grid_constant__ ForkConst f,
                                                            const __grid_constant__ DevPtrs P, const __grid_constant__ ForkPtrs F,
                                                            const uint8_t *__restrict__ env_mask, float *__restrict__ obs)
{
    extern __shared__ __align__(16) float smem[];
    const int lane = threadIdx.x & 31, warp_in_block = threadIdx.x >> 5;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int env = tid / KG, d = tid % KG;
    const bool env_ok = env < c.N;
    const bool sel = env_ok && (env_mask == nullptr || env_mask[env] != 0);
    const bool valid = sel && d < c.K;
    const int gi = env * c.K + d;
    const uint32_t gmask = group_mask<KG>(lane);
    constexpr int GPW = 32 / KG;
    const int rows_per_warp = GPW * c.K;
    float4 *stage = reinterpret_cast<float4 *>(smem) + (size_t)warp_in_block * 64;
    float *tile = smem + (size_t)(blockDim.x >> 5) * 256 + (size_t)warp_in_block * rows_per_warp * c.D;
    const int row = (lane / KG) * c.K + d;
    float *orow = tile + (size_t)row * c.D;

    Drone q;
    float angle = 0.f, ang_vel = 0.f, hsnap = 0.f, ex = 1.f, ey = 0.f;
    int fflags = 0;
    Rng g; g.k0 = c.key0; g.k1 = c.key1; g.gid = 0; g.step = c.rng_step;      // sub-step s of this call draws from counter rng_step + s
#pragma unroll
    for (int a = 0; a < 3; ++a) { q.p[a] = 1.0e6f * (float)(lane + 1); q.v[a] = 0.f; q.w[a] = 0.f; q.goal[a] = 0.f; }
#pragma unroll
    for (int a = 0; a < 9; ++a) q.R[a] = (a % 4 == 0) ? 1.f : 0.f;
#pragma unroll
    for (int a = 0; a < 4; ++a) { q.rd[a] = q.cd[a] = q.ou[a] = 0.f; }
    q.flags = 0; q.colmask = 0;
    if (env_ok) { g.gid = (uint32_t)(c.env_id_offset + env); fflags = F.flags[env]; }
    if (env_ok && d < c.K) { load_drone(P, gi, q); float4 h = F.plane[FP_HEADING][gi]; ang_vel = h.y; hsnap = h.z; }
    float u[4], t[4];
    rng_u4(g, SITE_SCENARIO, 0xFF, 7, d >> 1, u);
    float sa_ = u[(2 * d) & 3] - 0.5f, sb_ = u[(2 * d + 1) & 3] - 0.5f, sn_ = sqrtf(sa_ * sa_ + sb_ * sb_);
    rng_u4(g, SITE_SCENARIO, 0xFF, 8, 0, t);
    const float ring = t[0] * f.spawn_ring;
    rng_u4(g, SITE_SCENARIO, 0xFF, 9, 0, t);
    float ea = t[0] - 0.5f, eb = t[1] - 0.5f, en = sqrtf(ea * ea + eb * eb), rad = t[2] * f.ev_rspan + f.ev_rmin;
    ex = ea / en * rad; ey = eb / en * rad;
    evader_step<KG>(f, q, valid, (fflags & FF_PLACED) != 0, gmask, ex, ey);   // the very first reset ignores the chasers (dynamic_repulsive.py:44)
    q.goal[0] = ex; q.goal[1] = ey; q.goal[2] = 2.0f;
    rng_u4(g, SITE_SPAWN, d, 2, 0, u);
    angle = (u[0] - 0.5f) * 2.0f * QS_PI_F;
    q.p[0] = sa_ / sn_ * ring; q.p[1] = sb_ / sn_ * ring; q.p[2] = fmaxf(q.goal[2], c.spawn_min_z);
    rng_u4(g, SITE_SPAWN, d, 1, 0, u);
    float sy, cy;
    sincospif(-1.0f + 2.0f * u[0], &sy, &cy);
    set_yaw(q.R, cy, sy);
    q.v[0] = q.v[1] = q.v[2] = 0.f; q.w[0] = q.w[1] = q.w[2] = 0.f;
#pragma unroll
    for (int m = 0; m < 4; ++m) { q.rd[m] = 0.f; q.cd[m] = 0.f; }
    q.flags = 0; q.colmask = 0u;
    if (valid) {
        store_drone(P, gi, q, true);
        F.plane[FP_HEADING][gi] = make_float4(angle, ang_vel, hsnap, 0.f);
        fork_self_obs(c, f, g, SITE_SENSOR_RESET, d, q, angle, ang_vel, orow);
    }
    if (sel && d == 0) {
        int *pe = P.ecnt + env * EC_COUNT;
#pragma unroll
        for (int k = 0; k < EC_COUNT; ++k) pe[k] = 0;
        P.tick[env] = 0;
        F.evader[env] = make_float2(ex, ey); F.flags[env] = FF_PLACED;
    }
    fork_neighbor_obs<KG>(c, f, g, d, lane, gmask, valid, q, angle, hsnap, orow + c.S, stage);
    __syncwarp();
    const int warp_env0 = (tid - lane) / KG;
    const int warp_rows = max(0, min(GPW, c.N - warp_env0)) * c.K;
    for (int r = 0; r < warp_rows; ++r) {
        int e = warp_env0 + r / c.K;
        if (env_mask == nullptr || env_mask[e] != 0) {
            const float *s = tile + (size_t)r * c.D;
            float *tt = obs + ((size_t)warp_env0 * c.K + r) * c.D;
            for (int k = lane; k < c.D; k += 32) tt[k] = s[k];
        }
    }
}

}  // namespace qs
