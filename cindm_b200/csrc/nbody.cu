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
// parity with pymunk is UNPINNED.  One thread per design, body state in registers,
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

// Per-design state.  The body arrays are only ever indexed by the induction variables of fully unrolled loops, so they
// live in REGISTERS; the arbiter cache and the (rarely non-empty) active-contact list are dynamically indexed and live
// in local memory, touched only on steps that have, or recently had, a contact.  (Round 1 kept everything in one struct
// indexed by runtime loop bounds: the whole ~4 KB world sat in local memory and every step walked it - ncu: 5.5 G local
// load/store sectors for 1e5 designs, profiles/r2_nbody_score_full.md.)
struct Bodies {
    double px[kMaxBodies], py[kMaxBodies], vx[kMaxBodies], vy[kMaxBodies], bx[kMaxBodies], by[kMaxBodies];
};

struct Arbiters {
    // persistent arbiter cache, slot = wall contacts (body*4 + wall) then disc pairs (lexicographic)
    unsigned char state[kMaxArb];
    int stamp[kMaxArb];
    double jn_acc[kMaxArb];
    int n_live;                                        // slots whose state is not ARB_NONE
    // contacts of the current step, in solver order
    int n_active;
    unsigned char act_slot[kMaxArb];
    signed char act_a[kMaxArb], act_b[kMaxArb];       // b == -1: static wall
    double nx[kMaxArb], ny[kMaxArb], n_mass[kMaxArb], bias[kMaxArb], bounce[kMaxArb], j_bias[kMaxArb];
};

// The contact paths are rare (a quarter of the steps have any contact) and are kept OUT of line, with every body value passed
// by value: the per-step code stays small (instruction cache) and nothing takes the address of the register-resident bodies.
// (va, vb): velocities of the two bodies at detection time (vb unused for a wall)
__device__ __noinline__ void add_contact(Arbiters& w, int slot, int a, int b, double nx, double ny, double dist, int step,
                                         double bias_coef, double vax, double vay, double vbx, double vby) {
    // cpArbiterUpdate: a persisting or cached arbiter hands its accumulated impulse to the new contact
    if (w.state[slot] == ARB_NONE) { w.jn_acc[slot] = 0.0; w.state[slot] = ARB_FIRST; ++w.n_live; }
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
    double rvx = -vax, rvy = -vay;
    if (b >= 0) { rvx = vbx - vax; rvy = vby - vay; }
    w.bounce[k] = (rvx * nx + rvy * ny) * 1.0;              // e = 1.0 * 1.0
}

// One wall test, the arithmetic of round 1 (closest point on the segment a -> b, cpCollision.c CircleToSegment).
__device__ __noinline__ void wall_test(Arbiters& w, double px, double py, double vx, double vy, int i, int s, double ax, double ay,
                                       double bx, double by, int step, double bias_coef) {
    const double min_wall = kRadius + kWallRadius;
    const double dxs = bx - ax, dys = by - ay;
    double t = (dxs * (px - ax) + dys * (py - ay)) / (dxs * dxs + dys * dys);
    t = t < 0.0 ? 0.0 : (t > 1.0 ? 1.0 : t);
    const double cx = ax + dxs * t, cy = ay + dys * t;
    const double ddx = cx - px, ddy = cy - py;
    const double d2 = ddx * ddx + ddy * ddy;
    if (d2 < min_wall * min_wall) {
        const double d = sqrt(d2);
        double nx, ny;
        if (d != 0.0) { nx = ddx * (1.0 / d); ny = ddy * (1.0 / d); }
        else { const double len = sqrt(dxs * dxs + dys * dys); nx = dys / len; ny = -dxs / len; }   // segment normal
        add_contact(w, i * 4 + s, i, -1, nx, ny, d - min_wall, step, bias_coef, vx, vy, 0.0, 0.0);
    }
}

// Two discs closer than 2r (d2 = |p_j - p_i|^2 already found below the threshold).
__device__ __noinline__ void pair_contact(Arbiters& w, int i, int j, int n, double ddx, double ddy, double d2, double vxi, double vyi,
                                          double vxj, double vyj, int step, double bias_coef) {
    const double d = sqrt(d2);
    double nx = 1.0, ny = 0.0;
    if (d != 0.0) { nx = ddx * (1.0 / d); ny = ddy * (1.0 / d); }
    const int slot = 4 * kMaxBodies + i * (2 * n - i - 1) / 2 + (j - i - 1);      // lexicographic slot of (i, j) among n bodies
    add_contact(w, slot, i, j, nx, ny, d - 2.0 * kRadius, step, bias_coef, vxi, vyi, vxj, vyj);
}

// cached-arbiter filter (cpSpaceArbiterSetFilter)
__device__ __noinline__ void filter_arbiters(Arbiters& w, int n, int step) {
    const int n_slots = 4 * kMaxBodies + n * (n - 1) / 2;
    for (int s = 0; s < n_slots; ++s) {
        if (w.state[s] == ARB_NONE) continue;
        const int ticks = step - w.stamp[s];
        if (ticks >= 1 && w.state[s] != ARB_CACHED) w.state[s] = ARB_CACHED;
        if (ticks >= kPersistence) { w.state[s] = ARB_NONE; --w.n_live; }
    }
}

// Warm start + 10 sequential-impulse iterations over the active contacts, on dynamically indexed copies of the velocities
// (lv) and bias velocities (lb): [x | y][body].
__device__ __noinline__ void solve_contacts(Arbiters& w, double* lvx, double* lvy, double* lbx, double* lby, double dt_coef) {
    // 4. warm start
    for (int k = 0; k < w.n_active; ++k) {
        const int s = w.act_slot[k];
        if (w.state[s] == ARB_FIRST) continue;
        const double jx = w.nx[k] * w.jn_acc[s] * dt_coef, jy = w.ny[k] * w.jn_acc[s] * dt_coef;
        const int a = w.act_a[k], b = w.act_b[k];
        lvx[a] = lvx[a] - jx; lvy[a] = lvy[a] - jy;
        if (b >= 0) { lvx[b] = lvx[b] + jx; lvy[b] = lvy[b] + jy; }
    }
    // 5. sequential impulses
    for (int it = 0; it < kIterations; ++it) {
        for (int k = 0; k < w.n_active; ++k) {
            const int s = w.act_slot[k], a = w.act_a[k], b = w.act_b[k];
            const double nx = w.nx[k], ny = w.ny[k];
            double vbx = -lbx[a], vby = -lby[a], vrx = -lvx[a], vry = -lvy[a];
            if (b >= 0) { vbx = lbx[b] - lbx[a]; vby = lby[b] - lby[a]; vrx = lvx[b] - lvx[a]; vry = lvy[b] - lvy[a]; }
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
            lbx[a] = lbx[a] - nx * db; lby[a] = lby[a] - ny * db;
            lvx[a] = lvx[a] - nx * dj; lvy[a] = lvy[a] - ny * dj;
            if (b >= 0) {
                lbx[b] = lbx[b] + nx * db; lby[b] = lby[b] + ny * db;
                lvx[b] = lvx[b] + nx * dj; lvy[b] = lvy[b] + ny * dj;
            }
        }
    }
}

__device__ __forceinline__ void world_step(Bodies& q, Arbiters& w, int n, int step, double bias_coef, double dt_coef) {
    // arbiters used last step become NORMAL (start of cpSpaceStep)
    for (int k = 0; k < w.n_active; ++k) w.state[w.act_slot[k]] = ARB_NORMAL;
    w.n_active = 0;
    // 1. positions
#pragma unroll
    for (int i = 0; i < kMaxBodies; ++i) {
        if (i < n) {
            q.px[i] = q.px[i] + (q.vx[i] + q.bx[i]) * kDt;
            q.py[i] = q.py[i] + (q.vy[i] + q.by[i]) * kDt;
            q.bx[i] = 0.0; q.by[i] = 0.0;
        }
    }
    // 2-3. collide + pre-step.  Walls first (static index), then disc pairs.  A disc whose centre is at least 21 away from
    // a wall's LINE cannot be within 21 of the segment, so the (division-heavy) closest-point test is skipped exactly;
    // the negated comparisons send NaN coordinates to the full test.
    const double min_wall = kRadius + kWallRadius;
#pragma unroll
    for (int i = 0; i < kMaxBodies; ++i) {
        if (i < n) {
            if (!(q.px[i] >= min_wall)) wall_test(w, q.px[i], q.py[i], q.vx[i], q.vy[i], i, 0, 0.0, 0.0, 0.0, kBox, step, bias_coef);            // x = 0
            if (!(q.py[i] <= kBox - min_wall)) wall_test(w, q.px[i], q.py[i], q.vx[i], q.vy[i], i, 1, 0.0, kBox, kBox, kBox, step, bias_coef);   // y = 200
            if (!(q.px[i] <= kBox - min_wall)) wall_test(w, q.px[i], q.py[i], q.vx[i], q.vy[i], i, 2, kBox, kBox, kBox, 0.0, step, bias_coef);   // x = 200
            if (!(q.py[i] >= min_wall)) wall_test(w, q.px[i], q.py[i], q.vx[i], q.vy[i], i, 3, kBox, 0.0, 0.0, 0.0, step, bias_coef);            // y = 0
        }
    }
    const double min_disc = 2.0 * kRadius;
#pragma unroll
    for (int i = 0; i < kMaxBodies; ++i) {
#pragma unroll
        for (int j = i + 1; j < kMaxBodies; ++j) {
            if (j < n) {
                const double ddx = q.px[j] - q.px[i], ddy = q.py[j] - q.py[i];
                const double d2 = ddx * ddx + ddy * ddy;
                if (d2 < min_disc * min_disc)
                    pair_contact(w, i, j, n, ddx, ddy, d2, q.vx[i], q.vy[i], q.vx[j], q.vy[j], step, bias_coef);
            }
        }
    }
    if (w.n_live > 0) filter_arbiters(w, n, step);          // nothing to do while every slot is NONE
    if (w.n_active == 0) return;
    // (velocity integration is the identity: no gravity, damping 1)
    // The solver indexes bodies by contact: it works on a dynamically indexed copy of the velocities / bias velocities.
    double lvx[kMaxBodies], lvy[kMaxBodies], lbx[kMaxBodies], lby[kMaxBodies];
#pragma unroll
    for (int i = 0; i < kMaxBodies; ++i) { lvx[i] = q.vx[i]; lvy[i] = q.vy[i]; lbx[i] = q.bx[i]; lby[i] = q.by[i]; }
    solve_contacts(w, lvx, lvy, lbx, lby, dt_coef);
#pragma unroll
    for (int i = 0; i < kMaxBodies; ++i) { q.vx[i] = lvx[i]; q.vy[i] = lvy[i]; q.bx[i] = lbx[i]; q.by[i] = lby[i]; }
}

__device__ __forceinline__ void world_init(Bodies& q, Arbiters& w) {
    w.n_active = 0;
    w.n_live = 0;
    for (int s = 0; s < kMaxArb; ++s) { w.state[s] = ARB_NONE; w.stamp[s] = 0; w.jn_acc[s] = 0.0; }
#pragma unroll
    for (int i = 0; i < kMaxBodies; ++i) { q.px[i] = 0.0; q.py[i] = 0.0; q.vx[i] = 0.0; q.vy[i] = 0.0; q.bx[i] = 0.0; q.by[i] = 0.0; }
}

// space->collisionBias = pow(1 - 0.1, 60);  biasCoef = 1 - pow(collisionBias, dt): evaluated on the host (libm)
// and passed to the kernels so that the CPU oracle and the GPU use the very same double.
static double bias_coefficient() { return 1.0 - pow(pow(1.0 - 0.1, 60.0), kDt); }

__global__ void __launch_bounds__(128) nbody_rollout_kernel(const double* __restrict__ state0, double* __restrict__ traj,
                                                            int B, int n, int n_steps, int stride, double bc) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    Bodies q;
    Arbiters w;
    world_init(q, w);
#pragma unroll
    for (int i = 0; i < kMaxBodies; ++i) {
        if (i < n) {
            const double* s = state0 + ((long long)b * n + i) * 4;
            q.px[i] = s[0]; q.py[i] = s[1]; q.vx[i] = s[2]; q.vy[i] = s[3];
        }
    }
    const int frames = n_steps / stride;
    for (int step = 0; step < n_steps; ++step) {
        // traj[k] is the state after k steps (utils.py:1052-1054 records before stepping); keep k = stride-1, 2*stride-1, ...
        if (step % stride == stride - 1) {
            double* o = traj + (((long long)b * frames + step / stride) * n) * 4;
#pragma unroll
            for (int i = 0; i < kMaxBodies; ++i)
                if (i < n) { o[4 * i] = q.px[i]; o[4 * i + 1] = q.py[i]; o[4 * i + 2] = q.vx[i]; o[4 * i + 3] = q.vy[i]; }
        }
        world_step(q, w, n, step, bc, step == 0 ? 0.0 : 1.0);
    }
}

// Fused scoring (inference/inverse_design_diffusion_1d.py:316-337): frame 0 of each design (x200) is rolled out
// for (T-1)*4 steps; the frames after 3, 7, 11, ... steps (/200) are compared with the design's own frames.
__global__ void __launch_bounds__(128) score_designs_kernel(const float* __restrict__ pred, double* __restrict__ mae,
                                                            double* __restrict__ objective, int B, int T, int n,
                                                            double tx, double ty, double bc) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    Bodies q;
    Arbiters w;
    world_init(q, w);
    const int F = 4 * n;
    const float* p0 = pred + (long long)b * T * F;
#pragma unroll
    for (int i = 0; i < kMaxBodies; ++i) {
        if (i < n) {
            // cond_design[:, -1, :] * 200. is an fp32 product in the reference (utils.py:1139)
            q.px[i] = (double)(p0[4 * i] * 200.0f); q.py[i] = (double)(p0[4 * i + 1] * 200.0f);
            q.vx[i] = (double)(p0[4 * i + 2] * 200.0f); q.vy[i] = (double)(p0[4 * i + 3] * 200.0f);
        }
    }
    const int n_steps = (T - 1) * 4;
    double abs_sum = 0.0, obj = 0.0;
    for (int step = 0; step < n_steps; ++step) {
        if (step % 4 == 3) {
            const int t = step / 4 + 1;
            const float* pt = p0 + (long long)t * F;
#pragma unroll
            for (int i = 0; i < kMaxBodies; ++i) {
                if (i < n) {
                    const double sx = q.px[i] / 200.0, sy = q.py[i] / 200.0, svx = q.vx[i] / 200.0, svy = q.vy[i] / 200.0;
                    abs_sum += fabs(sx - (double)pt[4 * i]) + fabs(sy - (double)pt[4 * i + 1]) +
                               fabs(svx - (double)pt[4 * i + 2]) + fabs(svy - (double)pt[4 * i + 3]);
                    if (t == T - 1) {
                        const double dx = sx - tx, dy = sy - ty;
                        obj += sqrt(dx * dx + dy * dy);
                    }
                }
            }
        }
        world_step(q, w, n, step, bc, step == 0 ? 0.0 : 1.0);
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
