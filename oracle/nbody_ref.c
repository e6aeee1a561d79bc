/* CPU oracle for the scoring simulator (TEST INFRASTRUCTURE, not product).
 *
 * Plain-C restatement of what the reference asks pymunk / Chipmunk2D to do in
 * /root/reference/utils.py:1009-1125 (add_body :1009-1021, add_walls :1023-1033, run_simulation
 * :1041-1068, simulation :1071-1125) and data/nbody_simulation.py:54-116: discs of radius 20 and
 * mass 1 (elasticity 1, friction 0) in a 200 x 200 box bounded by four static segments of radius 1,
 * gravity 0, space.step(1/60), state recorded BEFORE each step.
 *
 * The arithmetic lives in a third-party dependency that is NOT in /root/reference and is not
 * installed in this image: pymunk (requirements.txt:5, unpinned; the notebook environment implies
 * pymunk 6.x wrapping Chipmunk2D 7.0.3).  This file restates Chipmunk2D 7.0.x's published stepping
 * algorithm (cpSpaceStep.c, cpArbiter.c, cpCollision.c, cpBody.c):
 *   cpBodyUpdatePosition, CircleToSegment / CircleToCircle, cpArbiterUpdate (persistent jnAcc,
 *   FIRST_COLLISION / NORMAL / CACHED states, collisionPersistence 3), cpArbiterPreStep (slop 0.1,
 *   biasCoef = 1 - collisionBias^dt with collisionBias = 0.9^60), cpArbiterApplyCachedImpulse,
 *   10 iterations of cpArbiterApplyImpulse.
 * Known deviations (documented in DESIGN.md): contacts are solved walls-first then disc pairs in
 * lexicographic order (Chipmunk's order comes from its spatial index), and the ~1e-15 angular
 * velocities Chipmunk picks up from rounding in r x j are not carried.
 *
 * A contact-order switch (nbody_ref_set_order) exists for the sensitivity study of DESIGN.md section 4;
 * the default order 0 is the one the CUDA kernel implements and every parity test uses.
 *
 * PARITY UNPINNED: the reference has no test or golden vector at this boundary and pymunk cannot be
 * run here; this oracle is pinned only by analytic known-answer tests (tests/test_nbody_oracle.py).
 *
 * Build: gcc -O2 -ffp-contract=off -shared -fPIC oracle/nbody_ref.c -o oracle/_build/libnbody_ref.so -lm
 */
#include <math.h>
#include <string.h>

#define MAXB 8
#define MAXARB (MAXB * (MAXB - 1) / 2 + 4 * MAXB)

enum { NONE = 0, FIRST = 1, NORMAL = 2, CACHED = 3 };

/* Contact solve order (sensitivity studies only; the CUDA kernel and every parity test use order 0).
 * Chipmunk2D iterates its arbiters in the order its bounding-box tree reported the pairs, which depends
 * on insertion history and on which leaves moved; that order cannot be reproduced without the library,
 * so profiles/contact_order_study.py measures how far the scores move under other orders:
 *   0 walls first (body-major), then disc pairs in lexicographic order          [default]
 *   1 disc pairs first, then walls
 *   2 the reverse of order 0
 *   3 a fresh random permutation every step (xorshift64*, seeded per design)
 *   4 body-major: for body i, its wall contacts, then its pairs (i, j > i)      [closest to a tree walk] */
static int g_order_mode = 0;
static unsigned long long g_order_seed = 0x9E3779B97F4A7C15ull;
void nbody_ref_set_order(int mode, unsigned long long seed) { g_order_mode = mode; g_order_seed = seed ? seed : 1ull; }
/* simultaneous-contact census of the last rollout call: steps with >= 2 active contacts sharing a body */
static long g_multi_steps = 0, g_contact_steps = 0;
void nbody_ref_census(long* multi_steps, long* contact_steps) { *multi_steps = g_multi_steps; *contact_steps = g_contact_steps; }

static unsigned long long rng_next(unsigned long long* s) {
    unsigned long long x = *s;
    x ^= x >> 12; x ^= x << 25; x ^= x >> 27;
    *s = x;
    return x * 0x2545F4914F6CDD1Dull;
}

typedef struct {
    int n;
    double p[MAXB][2], v[MAXB][2], vb[MAXB][2];
    int state[MAXARB], stamp[MAXARB];
    double jn_acc[MAXARB];
    int n_active, slot[MAXARB], a[MAXARB], b[MAXARB];
    double nrm[MAXARB][2], n_mass[MAXARB], bias[MAXARB], bounce[MAXARB], j_bias[MAXARB];
    int ord[MAXARB];               /* solve order: position -> index into the active list */
    unsigned long long rng;
} world_t;

/* fill w->ord for the active contacts of this step (order 0 is the identity) */
static void order_contacts(world_t* w) {
    int m = w->n_active;
    for (int k = 0; k < m; ++k) w->ord[k] = k;
    if (g_order_mode == 1) {                      /* pairs, then walls; each group keeps its lexicographic order */
        int q = 0;
        for (int k = 0; k < m; ++k) if (w->b[k] >= 0) w->ord[q++] = k;
        for (int k = 0; k < m; ++k) if (w->b[k] < 0) w->ord[q++] = k;
    } else if (g_order_mode == 2) {
        for (int k = 0; k < m; ++k) w->ord[k] = m - 1 - k;
    } else if (g_order_mode == 3) {               /* Fisher-Yates */
        for (int k = m - 1; k > 0; --k) {
            int r = (int)(rng_next(&w->rng) % (unsigned long long)(k + 1));
            int t = w->ord[k]; w->ord[k] = w->ord[r]; w->ord[r] = t;
        }
    } else if (g_order_mode == 4) {               /* stable sort by first body: walls of i, then pairs (i, j) */
        int q = 0;
        for (int i = 0; i < w->n; ++i)
            for (int k = 0; k < m; ++k) if (w->a[k] == i) w->ord[q++] = k;
    }
    /* census */
    if (m > 0) ++g_contact_steps;
    int touched[MAXB] = {0}, multi = 0;
    for (int k = 0; k < m; ++k) {
        if (++touched[w->a[k]] > 1) multi = 1;
        if (w->b[k] >= 0 && ++touched[w->b[k]] > 1) multi = 1;
    }
    if (multi) ++g_multi_steps;
}

static const double RADIUS = 20.0, WALL_RADIUS = 1.0, BOX = 200.0, DT = 1.0 / 60.0, SLOP = 0.1;

static void contact(world_t* w, int slot, int a, int b, double nx, double ny, double dist, int step, double bias_coef) {
    if (w->state[slot] == NONE) { w->jn_acc[slot] = 0.0; w->state[slot] = FIRST; }
    else if (w->state[slot] == CACHED) w->state[slot] = FIRST;
    w->stamp[slot] = step;
    int k = w->n_active++;
    w->slot[k] = slot; w->a[k] = a; w->b[k] = b;
    w->nrm[k][0] = nx; w->nrm[k][1] = ny;
    w->n_mass[k] = 1.0 / (b >= 0 ? 2.0 : 1.0);
    double pen = dist + SLOP;
    w->bias[k] = -bias_coef * (pen < 0.0 ? pen : 0.0) / DT;
    w->j_bias[k] = 0.0;
    double rvx = -w->v[a][0], rvy = -w->v[a][1];
    if (b >= 0) { rvx = w->v[b][0] - w->v[a][0]; rvy = w->v[b][1] - w->v[a][1]; }
    w->bounce[k] = (rvx * nx + rvy * ny) * 1.0;
}

static void step_world(world_t* w, int step, double bias_coef, double dt_coef) {
    static const double ax[4] = {0.0, 0.0, 200.0, 200.0}, ay[4] = {0.0, 200.0, 200.0, 0.0};
    static const double bx[4] = {0.0, 200.0, 200.0, 0.0}, by[4] = {200.0, 200.0, 0.0, 0.0};
    int n = w->n;
    for (int k = 0; k < w->n_active; ++k) w->state[w->slot[k]] = NORMAL;
    w->n_active = 0;
    for (int i = 0; i < n; ++i) {
        w->p[i][0] = w->p[i][0] + (w->v[i][0] + w->vb[i][0]) * DT;
        w->p[i][1] = w->p[i][1] + (w->v[i][1] + w->vb[i][1]) * DT;
        w->vb[i][0] = 0.0; w->vb[i][1] = 0.0;
    }
    const double min_wall = RADIUS + WALL_RADIUS;
    for (int i = 0; i < n; ++i)
        for (int s = 0; s < 4; ++s) {
            double dxs = bx[s] - ax[s], dys = by[s] - ay[s];
            double t = (dxs * (w->p[i][0] - ax[s]) + dys * (w->p[i][1] - ay[s])) / (dxs * dxs + dys * dys);
            t = t < 0.0 ? 0.0 : (t > 1.0 ? 1.0 : t);
            double cx = ax[s] + dxs * t, cy = ay[s] + dys * t;
            double ddx = cx - w->p[i][0], ddy = cy - w->p[i][1];
            double d2 = ddx * ddx + ddy * ddy;
            if (d2 < min_wall * min_wall) {
                double d = sqrt(d2), nx, ny;
                if (d != 0.0) { nx = ddx * (1.0 / d); ny = ddy * (1.0 / d); }
                else { double len = sqrt(dxs * dxs + dys * dys); nx = dys / len; ny = -dxs / len; }
                contact(w, i * 4 + s, i, -1, nx, ny, d - min_wall, step, bias_coef);
            }
        }
    int slot = 4 * MAXB;
    const double min_disc = 2.0 * RADIUS;
    for (int i = 0; i < n; ++i)
        for (int j = i + 1; j < n; ++j, ++slot) {
            double ddx = w->p[j][0] - w->p[i][0], ddy = w->p[j][1] - w->p[i][1];
            double d2 = ddx * ddx + ddy * ddy;
            if (d2 < min_disc * min_disc) {
                double d = sqrt(d2), nx = 1.0, ny = 0.0;
                if (d != 0.0) { nx = ddx * (1.0 / d); ny = ddy * (1.0 / d); }
                contact(w, slot, i, j, nx, ny, d - min_disc, step, bias_coef);
            }
        }
    int n_slots = 4 * MAXB + n * (n - 1) / 2;
    for (int s = 0; s < n_slots; ++s) {
        if (w->state[s] == NONE) continue;
        int ticks = step - w->stamp[s];
        if (ticks >= 1 && w->state[s] != CACHED) w->state[s] = CACHED;
        if (ticks >= 3) w->state[s] = NONE;
    }
    order_contacts(w);
    for (int kk = 0; kk < w->n_active; ++kk) {
        int k = w->ord[kk];
        int s = w->slot[k];
        if (w->state[s] == FIRST) continue;
        double jx = w->nrm[k][0] * w->jn_acc[s] * dt_coef, jy = w->nrm[k][1] * w->jn_acc[s] * dt_coef;
        int a = w->a[k], b = w->b[k];
        w->v[a][0] = w->v[a][0] - jx; w->v[a][1] = w->v[a][1] - jy;
        if (b >= 0) { w->v[b][0] = w->v[b][0] + jx; w->v[b][1] = w->v[b][1] + jy; }
    }
    for (int it = 0; it < 10; ++it)
        for (int kk = 0; kk < w->n_active; ++kk) {
            int k = w->ord[kk];
            int s = w->slot[k], a = w->a[k], b = w->b[k];
            double nx = w->nrm[k][0], ny = w->nrm[k][1];
            double vbx = -w->vb[a][0], vby = -w->vb[a][1], vrx = -w->v[a][0], vry = -w->v[a][1];
            if (b >= 0) {
                vbx = w->vb[b][0] - w->vb[a][0]; vby = w->vb[b][1] - w->vb[a][1];
                vrx = w->v[b][0] - w->v[a][0]; vry = w->v[b][1] - w->v[a][1];
            }
            double vbn = vbx * nx + vby * ny, vrn = vrx * nx + vry * ny;
            double jbn = (w->bias[k] - vbn) * w->n_mass[k];
            double jbn_old = w->j_bias[k];
            double jb_new = jbn_old + jbn;
            w->j_bias[k] = jb_new > 0.0 ? jb_new : 0.0;
            double jn = -(w->bounce[k] + vrn) * w->n_mass[k];
            double jn_old = w->jn_acc[s];
            double jn_new = jn_old + jn;
            w->jn_acc[s] = jn_new > 0.0 ? jn_new : 0.0;
            double db = w->j_bias[k] - jbn_old, dj = w->jn_acc[s] - jn_old;
            w->vb[a][0] = w->vb[a][0] - nx * db; w->vb[a][1] = w->vb[a][1] - ny * db;
            w->v[a][0] = w->v[a][0] - nx * dj; w->v[a][1] = w->v[a][1] - ny * dj;
            if (b >= 0) {
                w->vb[b][0] = w->vb[b][0] + nx * db; w->vb[b][1] = w->vb[b][1] + ny * db;
                w->v[b][0] = w->v[b][0] + nx * dj; w->v[b][1] = w->v[b][1] + ny * dj;
            }
        }
}

/* state0[B][n][4] (x, y, vx, vy in pixel units) -> traj[B][n_steps/stride][n][4]: the state after
 * stride-1, 2*stride-1, ... steps (utils.py:1144 keeps traj[:, time_interval-1::time_interval]). */
void nbody_ref_rollout(const double* state0, double* traj, int B, int n, int n_steps, int stride) {
    double bias_coef = 1.0 - pow(pow(1.0 - 0.1, 60.0), DT);
    int frames = n_steps / stride;
    g_multi_steps = 0; g_contact_steps = 0;
    for (int b = 0; b < B; ++b) {
        world_t w;
        memset(&w, 0, sizeof w);
        w.n = n;
        w.rng = g_order_seed ^ (0xD1B54A32D192ED03ull * (unsigned long long)(b + 1));
        if (w.rng == 0) w.rng = 1;
        for (int i = 0; i < n; ++i) {
            const double* s = state0 + ((long)b * n + i) * 4;
            w.p[i][0] = s[0]; w.p[i][1] = s[1]; w.v[i][0] = s[2]; w.v[i][1] = s[3];
        }
        for (int step = 0; step < n_steps; ++step) {
            if (step % stride == stride - 1) {
                double* o = traj + (((long)b * frames + step / stride) * n) * 4;
                for (int i = 0; i < n; ++i) { o[4*i] = w.p[i][0]; o[4*i+1] = w.p[i][1]; o[4*i+2] = w.v[i][0]; o[4*i+3] = w.v[i][1]; }
            }
            step_world(&w, step, bias_coef, step == 0 ? 0.0 : 1.0);
        }
    }
}
