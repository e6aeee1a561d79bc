// Ground-truth scoring rollout on the GPU: the hard-disc world that the reference builds with
// pymunk / Chipmunk2D in utils.py:1009-1125 (discs r=20, m=1, elasticity 1, friction 0; four
// static wall segments of radius 1 on the edges of the 200 x 200 box; gravity 0;
// space.step(1/60)), advanced with Chipmunk's fixed-step scheme:
//   1. integrate positions with (v + v_bias), clear v_bias
//   2. collide: disc-wall (closest point on the segment), disc-disc
//   3. per contact: nMass, penetration bias (slop 0.1, bias coefficient 1-0.9 per 1/60 s), bounce = e * v_rel.n
//   4. warm start: re-apply the accumulated normal impulse of contacts that persist
//   5. 10 sequential-impulse iterations (bias impulses into v_bias, normal impulses clamped >= 0)
// Contact persistence follows Chipmunk's arbiter cache (collisionPersistence = 3 steps).  Friction
// is zero and contact normals pass through the disc centres, so no angular state is carried.
// Chipmunk2D itself is not vendored in the reference and not installed here: this restatement is
// from its published algorithm (cpSpaceStep.c / cpArbiter.c / cpCollision.c, version 7.0.x) and
// parity with pymunk is UNPINNED.  One thread per design, all state thread-local,
// double precision, compiled with -fmad=false so it matches the C oracle bit for bit.
#include <cmath>

#include "engine.h"

namespace cindm {

namespace {

constexpr int kMaxBodies = 8;
constexpr int kMaxArb = kMaxBodies * (kMaxBodies - 1) / 2 + 4 * kMaxBodies;
constexpr double kRadius = 20.0, kWallRadius = 1.0, kBox = 200.0;
constexpr double kDt = 1.0 / 60.0;
constexpr double kSlop = 0.1;
constexpr int kIterations = 10;
constexpr int kPersistence = 3;

enum { ARB_NONE = 0, ARB_FIRST = 1, ARB_NORMAL = 2, ARB_CACHED = 3 };

struct World {
    int n;
    double px[kMaxBodies], py[kMaxBodies], vx[kMaxBodies], vy[kMaxBodies], bx[kMaxBodies], by[kMaxBodies];
    // persistent arbiter cache, slot = wall contacts (body*4 + wall) then disc pairs (lexicographic)
    unsigned char state[kMaxArb];
    int stamp[kMaxArb];
    double jn_acc[kMaxArb];
    // contacts of the current step, in solver order
    int n_active;
    unsigned char act_slot[kMaxArb];
    signed char act_a[kMaxArb], act_b[kMaxArb];       // b == -1: static wall
    double nx[kMaxArb], ny[kMaxArb], n_mass[kMaxArb], bias[kMaxArb], bounce[kMaxArb], j_bias[kMaxArb];
};

__device__ __forceinline__ void add_contact(World& w, int slot, int a, int b, double nx, double ny, double dist, int step,
                                            double bias_coef) {
    // cpArbiterUpdate: a persisting or cached arbiter hands its accumulated impulse to the new contact
    if (w.state[slot] == ARB_NONE) { w.jn_acc[slot] = 0.0; w.state[slot] = ARB_FIRST; }
    else if (w.state[slot] == ARB_CACHED) w.state[slot] = ARB_FIRST;
    w.stamp[slot] = step;
    const int k = w.n_active++;
    w.act_slot[k] = (unsigned char)slot; w.act_a[k] = (signed char)a; w.act_b[k] = (signed char)b;
    w.nx[k] = nx; w.ny[k] = ny;
    // cpArbiterPreStep
    const double inv_mass_sum = b >= 0 ? 2.0 : 1.0;       // m = 1 for discs, walls are static
    w.n_mass[k] = 1.0 / inv_mass_sum;
    const double pen = dist + kSlop;
    w.bias[k] = -bias_coef * (pen < 0.0 ? pen : 0.0) / kDt;
    w.j_bias[k] = 0.0;
    double rvx = -w.vx[a], rvy = -w.vy[a];
    if (b >= 0) { rvx = w.vx[b] - w.vx[a]; rvy = w.vy[b] - w.vy[a]; }
    w.bounce[k] = (rvx * nx + rvy * ny) * 1.0;              // e = 1.0 * 1.0
}

__device__ void world_step(World& w, int step, double bias_coef, double dt_coef) {
    const int n = w.n;
    // arbiters used last step become NORMAL (start of cpSpaceStep)
    for (int k = 0; k < w.n_active; ++k) w.state[w.act_slot[k]] = ARB_NORMAL;
    w.n_active = 0;
    // 1. positions
    for (int i = 0; i < n; ++i) {
        w.px[i] = w.px[i] + (w.vx[i] + w.bx[i]) * kDt;
        w.py[i] = w.py[i] + (w.vy[i] + w.by[i]) * kDt;
        w.bx[i] = 0.0; w.by[i] = 0.0;
    }
    // 2-3. collide + pre-step.  Walls first (static index), then disc pairs.
    const double wall_ax[4] = {0.0, 0.0, kBox, kBox}, wall_ay[4] = {0.0, kBox, kBox, 0.0};
    const double wall_bx[4] = {0.0, kBox, kBox, 0.0}, wall_by[4] = {kBox, kBox, 0.0, 0.0};
    const double min_wall = kRadius + kWallRadius;
    for (int i = 0; i < n; ++i) {
        for (int s = 0; s < 4; ++s) {
            const double dxs = wall_bx[s] - wall_ax[s], dys = wall_by[s] - wall_ay[s];
            double t = (dxs * (w.px[i] - wall_ax[s]) + dys * (w.py[i] - wall_ay[s])) / (dxs * dxs + dys * dys);
            t = t < 0.0 ? 0.0 : (t > 1.0 ? 1.0 : t);
            const double cx = wall_ax[s] + dxs * t, cy = wall_ay[s] + dys * t;
            const double ddx = cx - w.px[i], ddy = cy - w.py[i];
            const double d2 = ddx * ddx + ddy * ddy;
            if (d2 < min_wall * min_wall) {
                const double d = sqrt(d2);
                double nx, ny;
                if (d != 0.0) { nx = ddx * (1.0 / d); ny = ddy * (1.0 / d); }
                else { const double len = sqrt(dxs * dxs + dys * dys); nx = dys / len; ny = -dxs / len; }   // segment normal
                add_contact(w, i * 4 + s, i, -1, nx, ny, d - min_wall, step, bias_coef);
            }
        }
    }
    const int pair_base = 4 * kMaxBodies;
    int slot = pair_base;
    const double min_disc = 2.0 * kRadius;
    for (int i = 0; i < n; ++i) {
        for (int j = i + 1; j < n; ++j, ++slot) {
            const double ddx = w.px[j] - w.px[i], ddy = w.py[j] - w.py[i];
            const double d2 = ddx * ddx + ddy * ddy;
            if (d2 < min_disc * min_disc) {
                const double d = sqrt(d2);
                double nx = 1.0, ny = 0.0;
                if (d != 0.0) { nx = ddx * (1.0 / d); ny = ddy * (1.0 / d); }
                add_contact(w, slot, i, j, nx, ny, d - min_disc, step, bias_coef);
            }
        }
    }
    // cached-arbiter filter (cpSpaceArbiterSetFilter)
    const int n_slots = pair_base + n * (n - 1) / 2;
    for (int s = 0; s < n_slots; ++s) {
        if (w.state[s] == ARB_NONE) continue;
        const int ticks = step - w.stamp[s];
        if (ticks >= 1 && w.state[s] != ARB_CACHED) w.state[s] = ARB_CACHED;
        if (ticks >= kPersistence) w.state[s] = ARB_NONE;
    }
    // (velocity integration is the identity: no gravity, damping 1)
    // 4. warm start
    for (int k = 0; k < w.n_active; ++k) {
        const int s = w.act_slot[k];
        if (w.state[s] == ARB_FIRST) continue;
        const double jx = w.nx[k] * w.jn_acc[s] * dt_coef, jy = w.ny[k] * w.jn_acc[s] * dt_coef;
        const int a = w.act_a[k], b = w.act_b[k];
        w.vx[a] = w.vx[a] - jx; w.vy[a] = w.vy[a] - jy;
        if (b >= 0) { w.vx[b] = w.vx[b] + jx; w.vy[b] = w.vy[b] + jy; }
    }
    // 5. sequential impulses
    for (int it = 0; it < kIterations; ++it) {
        for (int k = 0; k < w.n_active; ++k) {
            const int s = w.act_slot[k], a = w.act_a[k], b = w.act_b[k];
            const double nx = w.nx[k], ny = w.ny[k];
            double vbx = -w.bx[a], vby = -w.by[a], vrx = -w.vx[a], vry = -w.vy[a];
            if (b >= 0) { vbx = w.bx[b] - w.bx[a]; vby = w.by[b] - w.by[a]; vrx = w.vx[b] - w.vx[a]; vry = w.vy[b] - w.vy[a]; }
            const double vbn = vbx * nx + vby * ny;
            const double vrn = vrx * nx + vry * ny;
            const double jbn = (w.bias[k] - vbn) * w.n_mass[k];
            const double jbn_old = w.j_bias[k];
            const double jb_new = jbn_old + jbn;
            w.j_bias[k] = jb_new > 0.0 ? jb_new : 0.0;
            const double jn = -(w.bounce[k] + vrn) * w.n_mass[k];
            const double jn_old = w.jn_acc[s];
            const double jn_new = jn_old + jn;
            w.jn_acc[s] = jn_new > 0.0 ? jn_new : 0.0;
            const double db = w.j_bias[k] - jbn_old, dj = w.jn_acc[s] - jn_old;
            w.bx[a] = w.bx[a] - nx * db; w.by[a] = w.by[a] - ny * db;
            w.vx[a] = w.vx[a] - nx * dj; w.vy[a] = w.vy[a] - ny * dj;
            if (b >= 0) {
                w.bx[b] = w.bx[b] + nx * db; w.by[b] = w.by[b] + ny * db;
                w.vx[b] = w.vx[b] + nx * dj; w.vy[b] = w.vy[b] + ny * dj;
            }
        }
    }
}

__device__ __forceinline__ void world_init(World& w, int n) {
    w.n = n;
    w.n_active = 0;
    for (int s = 0; s < kMaxArb; ++s) { w.state[s] = ARB_NONE; w.stamp[s] = 0; w.jn_acc[s] = 0.0; }
    for (int i = 0; i < kMaxBodies; ++i) { w.bx[i] = 0.0; w.by[i] = 0.0; }
}

// space->collisionBias = pow(1 - 0.1, 60);  biasCoef = 1 - pow(collisionBias, dt): evaluated on the host (libm)
// and passed to the kernels so that the CPU oracle and the GPU use the very same double.
static double bias_coefficient() { return 1.0 - pow(pow(1.0 - 0.1, 60.0), kDt); }

__global__ void __launch_bounds__(128) nbody_rollout_kernel(const double* __restrict__ state0, double* __restrict__ traj,
                                                            int B, int n, int n_steps, int stride, double bc) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    World w;
    world_init(w, n);
    for (int i = 0; i < n; ++i) {
        const double* s = state0 + ((long long)b * n + i) * 4;
        w.px[i] = s[0]; w.py[i] = s[1]; w.vx[i] = s[2]; w.vy[i] = s[3];
    }
    const int frames = n_steps / stride;
    for (int step = 0; step < n_steps; ++step) {
        // traj[k] is the state after k steps (utils.py:1052-1054 records before stepping); keep k = stride-1, 2*stride-1, ...
        if (step % stride == stride - 1) {
            double* o = traj + (((long long)b * frames + step / stride) * n) * 4;
            for (int i = 0; i < n; ++i) { o[4 * i] = w.px[i]; o[4 * i + 1] = w.py[i]; o[4 * i + 2] = w.vx[i]; o[4 * i + 3] = w.vy[i]; }
        }
        world_step(w, step, bc, step == 0 ? 0.0 : 1.0);
    }
}

// Fused scoring (inference/inverse_design_diffusion_1d.py:316-337): frame 0 of each design (x200) is rolled out
// for (T-1)*4 steps; the frames after 3, 7, 11, ... steps (/200) are compared with the design's own frames.
__global__ void __launch_bounds__(128) score_designs_kernel(const float* __restrict__ pred, double* __restrict__ mae,
                                                            double* __restrict__ objective, int B, int T, int n,
                                                            double tx, double ty, double bc) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    World w;
    world_init(w, n);
    const int F = 4 * n;
    const float* p0 = pred + (long long)b * T * F;
    for (int i = 0; i < n; ++i) {
        // cond_design[:, -1, :] * 200. is an fp32 product in the reference (utils.py:1139)
        w.px[i] = (double)(p0[4 * i] * 200.0f); w.py[i] = (double)(p0[4 * i + 1] * 200.0f);
        w.vx[i] = (double)(p0[4 * i + 2] * 200.0f); w.vy[i] = (double)(p0[4 * i + 3] * 200.0f);
    }
    const int n_steps = (T - 1) * 4;
    double abs_sum = 0.0, obj = 0.0;
    for (int step = 0; step < n_steps; ++step) {
        if (step % 4 == 3) {
            const int t = step / 4 + 1;
            const float* pt = p0 + (long long)t * F;
            for (int i = 0; i < n; ++i) {
                const double sx = w.px[i] / 200.0, sy = w.py[i] / 200.0, svx = w.vx[i] / 200.0, svy = w.vy[i] / 200.0;
                abs_sum += fabs(sx - (double)pt[4 * i]) + fabs(sy - (double)pt[4 * i + 1]) +
                           fabs(svx - (double)pt[4 * i + 2]) + fabs(svy - (double)pt[4 * i + 3]);
                if (t == T - 1) {
                    const double dx = sx - tx, dy = sy - ty;
                    obj += sqrt(dx * dx + dy * dy);
                }
            }
        }
        world_step(w, step, bc, step == 0 ? 0.0 : 1.0);
    }
    mae[b] = abs_sum / (double)(T * F);          // frame 0 contributes |pred - pred| = 0 to the mean over T frames
    objective[b] = obj / (double)n;
}

}  // namespace

int nbody_rollout(const double* state0, double* traj, int B, int n, int n_steps, int stride, cudaStream_t st) {
    if (n < 1 || n > kMaxBodies) return fail(-2, "nbody rollout supports 1..8 bodies");
    if (stride < 1 || n_steps < 0 || n_steps % stride) return fail(-2, "n_steps must be a multiple of stride");
    if (B == 0 || n_steps == 0) return 0;
    KernelTimer kt("nbody_rollout", st, (double)B * n * 32.0 * (1 + n_steps / stride));
    nbody_rollout_kernel<<<(B + 127) / 128, 128, 0, st>>>(state0, traj, B, n, n_steps, stride, bias_coefficient());
    CINDM_CHECK_LAUNCH();
    return 0;
}

int score_designs(const float* pred, double* mae, double* objective, int B, int T, int n, double tx, double ty,
                  cudaStream_t st) {
    if (n < 1 || n > kMaxBodies) return fail(-2, "scoring supports 1..8 bodies");
    if (T < 2) return fail(-2, "need at least two frames");
    if (B == 0) return 0;
    KernelTimer kt("score_designs", st, (double)B * (T * n * 16.0 + 16.0));
    score_designs_kernel<<<(B + 127) / 128, 128, 0, st>>>(pred, mae, objective, B, T, n, tx, ty, bias_coefficient());
    CINDM_CHECK_LAUNCH();
    return 0;
}

}  // namespace cindm
