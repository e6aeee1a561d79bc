// Guidance gradient, fused DDPM update and the reverse-diffusion loop.
//
// One DDPM step of p_sample_compose_inside (reference model/diffusion_1d.py:1189-1376, SURVEY
// Appendix B) is:   repeat R x { eps = compose(x);  pred = mu(x, eps) - g(x);  x = renoise(pred) }
// then out = pred + sigma_t * noise.   Here everything after `eps` is ONE kernel per evaluation
// (x0 + clamp + posterior mean + closed-form objective gradient + re-noise / final noise), the
// noise comes either from a caller-supplied tensor (parity runs) or from Philox4x32-10 keyed by
// (seed, global candidate id, t, draw) so results do not depend on how candidates are sharded,
// and the whole step is captured once in a CUDA graph and replayed with a device-resident t.
#include "engine.h"

namespace cindm {

// ---------------------------------------------------------------- Philox4x32-10 + Box-Muller
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
    uint32_t hi0 = __umulhi(M0, c[0]), lo0 = M0 * c[0];
    uint32_t hi1 = __umulhi(M1, c[2]), lo1 = M1 * c[2];
    uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}

__device__ __forceinline__ float4 philox_normal4(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) {
    uint32_t c[4] = {c0, c1, c2, c3};
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        philox_round(c, k0, k1);
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    const float two_pow_m32 = 2.3283064365386963e-10f;
    float u0 = fmaf((float)c[0], two_pow_m32, 0.5f * two_pow_m32);
    float u1 = fmaf((float)c[1], two_pow_m32, 0.5f * two_pow_m32);
    float u2 = fmaf((float)c[2], two_pow_m32, 0.5f * two_pow_m32);
    float u3 = fmaf((float)c[3], two_pow_m32, 0.5f * two_pow_m32);
    u0 = fminf(u0, 0.99999994f); u2 = fminf(u2, 0.99999994f);
    float r0 = sqrtf(-2.0f * __logf(u0)), r1 = sqrtf(-2.0f * __logf(u2));
    float s0, c0f, s1, c1f;
    __sincosf(6.283185307179586f * u1, &s0, &c0f);
    __sincosf(6.283185307179586f * u3, &s1, &c1f);
    return make_float4(r0 * c0f, r0 * s0, r1 * c1f, r1 * s1);
}

// counter layout: (element/4 within the candidate, global candidate id lo, candidate hi, (t << 16) | draw)
__device__ __forceinline__ float4 noise4(uint64_t seed, long long cand, int vec_in_cand, int t, int draw) {
    return philox_normal4(seed, (uint32_t)vec_in_cand, (uint32_t)cand, (uint32_t)((unsigned long long)cand >> 32),
                          ((uint32_t)t << 16) | (uint32_t)(draw & 0xFFFF));
}

__global__ void __launch_bounds__(256) fill_noise_kernel(float4* __restrict__ x, long long nvec, int vec_per_cand,
                                                         uint64_t seed, long long cand_off, int t, int draw) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
        long long b = i / vec_per_cand;
        int v = (int)(i - b * vec_per_cand);
        x[i] = noise4(seed, cand_off + b, v, t, draw);
    }
}

int launch_fill_noise(float* x, int B, int T, int n, uint64_t seed, int64_t cand_off, int t, int draw, cudaStream_t st) {
    long long nvec = (long long)B * T * n;
    if (nvec == 0) return 0;
    int blocks = (int)((nvec + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    fill_noise_kernel<<<blocks, 256, 0, st>>>((float4*)x, nvec, T * n, seed, cand_off, t, draw);
    CINDM_CHECK_LAUNCH();
    return 0;
}

__global__ void step_counter_kernel(int* t, int delta) { pdl_wait(); pdl_trigger(); *t += delta; }
__global__ void ddim_set_time_kernel(int* t, const int* times, const int* step) { pdl_wait(); pdl_trigger(); *t = times[*step]; }
int launch_step_counter(int* t_dev, int delta, cudaStream_t st) {
    CINDM_CHECK_CUDA(launch_chain(step_counter_kernel, dim3(1), dim3(1), 0, st, t_dev, delta));
    CINDM_CHECK_LAUNCH();
    return 0;
}

// ---------------------------------------------------------------- chained conditions (composing_time_sample :1827-1829)
__global__ void __launch_bounds__(256) chain_condition_kernel(float4* __restrict__ x, int rows_per_block, int blocks, int T, int n,
                                                              int cond_rows) {
    pdl_wait();
    pdl_trigger();
    const long long per = (long long)rows_per_block * cond_rows * n;          // float4 per block
    const long long total = per * (blocks - 1);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(i / per) + 1;                                     // destination block
        long long r = i - (long long)(k - 1) * per;
        const int j = (int)(r % n); r /= n;
        const int row = (int)(r % cond_rows);
        const long long b = r / cond_rows;
        const long long src = (((long long)(k - 1) * rows_per_block + b) * T + (T - cond_rows + row)) * n + j;
        const long long dst = (((long long)k * rows_per_block + b) * T + row) * n + j;
        x[dst] = x[src];
    }
}

int launch_chain_condition(float* x, int rows_per_block, int blocks, int T, int n, int cond_rows, cudaStream_t st) {
    if (blocks < 2 || cond_rows <= 0) return 0;
    const long long total = (long long)rows_per_block * cond_rows * n * (blocks - 1);
    int grid = (int)((total + 255) / 256);
    if (grid > 148 * 8) grid = 148 * 8;
    CINDM_CHECK_CUDA(launch_chain(chain_condition_kernel, dim3(grid), dim3(256), 0, st, (float4*)x, rows_per_block, blocks, T, n, cond_rows));
    CINDM_CHECK_LAUNCH();
    return 0;
}

// ---------------------------------------------------------------- objective gradient (closed form)
// J = coef * sum_b sum_j d(p[b,T-1,j], target) + cc * sum_b mean_t sum_{j,xy} (p[t+1]-p[t])^2
//   d = ||.||_2 (L2) or ||.||^2 (L2square); the distance term is evaluated in fp64 because the
//   driver's target tensor is fp64 (inference/inverse_design_diffusion_1d.py:281), the consistency
//   term in fp32; velocities get no gradient.  Returns (dJ/dx, dJ/dy) of body j at (b, t).
__device__ __forceinline__ float2 objective_grad(const float* __restrict__ x, long long bt_base, int t, int T, int F,
                                                 int j, float px, float py, const cindm_objective& o) {
    float gx = 0.f, gy = 0.f;
    if (o.consistency_coef > 0.f) {
        const float k = o.consistency_coef / (float)(T - 1);
        if (t >= 1) {
            const float* q = x + bt_base - F + 4 * j;
            gx = __fmul_rn(__fmul_rn(__fsub_rn(px, q[0]), 2.0f), k);
            gy = __fmul_rn(__fmul_rn(__fsub_rn(py, q[1]), 2.0f), k);
        }
        if (t <= T - 2) {
            const float* q = x + bt_base + F + 4 * j;
            gx = __fsub_rn(gx, __fmul_rn(__fmul_rn(__fsub_rn(q[0], px), 2.0f), k));
            gy = __fsub_rn(gy, __fmul_rn(__fmul_rn(__fsub_rn(q[1], py), 2.0f), k));
        }
    }
    if (t == T - 1) {
        double dx = (double)px - o.target_x, dy = (double)py - o.target_y;
        double lx, ly;
        if (o.mode == CINDM_OBJ_L2) {
            double nrm = sqrt(dx * dx + dy * dy);
            lx = (double)o.coef * dx / nrm;          // 0/0 -> NaN, as autograd produces
            ly = (double)o.coef * dy / nrm;
        } else {
            lx = (double)o.coef * 2.0 * dx;
            ly = (double)o.coef * 2.0 * dy;
        }
        gx = __fadd_rn(gx, (float)lx);
        gy = __fadd_rn(gy, (float)ly);
    }
    return make_float2(gx, gy);
}

__global__ void __launch_bounds__(256) design_grad_kernel(const float* __restrict__ x, float* __restrict__ g, int B, int T,
                                                          int n, cindm_objective o) {
    const long long total = (long long)B * T * n;
    const int F = 4 * n;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int j = (int)(i % n);
        long long bt = i / n;
        int t = (int)(bt % T);
        float4 v = reinterpret_cast<const float4*>(x)[i];
        float2 gr = objective_grad(x, bt * F, t, T, F, j, v.x, v.y, o);
        reinterpret_cast<float4*>(g)[i] = make_float4(gr.x, gr.y, 0.f, 0.f);
    }
}

int launch_design_grad(const float* x, float* g, int B, int T, int n, const cindm_objective& obj, cudaStream_t st) {
    long long total = (long long)B * T * n;
    if (total == 0) return 0;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    design_grad_kernel<<<blocks, 256, 0, st>>>(x, g, B, T, n, obj);
    CINDM_CHECK_LAUNCH();
    return 0;
}

// ---------------------------------------------------------------- fused DDPM update
struct UpdateParams {
    const float* x; const float* eps; const float* noise; float* x_out; float* pred_out; float* x0_out;
    const float* sched; const int* t_dev;
    long long cand_off; unsigned long long seed;
    int B, T, n, timesteps, t_host, renoise, t_start, draws_per_step, draw, use_philox;
    cindm_objective obj;
    int ddim; const float* ddim_coef; const int* step_dev;
    const float* mean_in; const float* x0_in;
    int cond_rows;
    const float* overwrite; int ow_rows;
};

__global__ void __launch_bounds__(256) ddpm_update_kernel(UpdateParams p) {
    pdl_wait();
    pdl_trigger();
    const int t = p.t_dev ? *p.t_dev : p.t_host;
    const int TS = p.timesteps;
    // coefficients, read from the fp32 buffers exactly as `extract` does (reference :454-462)
    const float A = p.sched[TAB_SQRT_RECIP_ACP * TS + t];
    const float Bc = p.sched[TAB_SQRT_RECIPM1_ACP * TS + t];
    const float c1 = p.sched[TAB_POST_C1 * TS + t];
    const float c2 = p.sched[TAB_POST_C2 * TS + t];
    float gscale = 0.f;
    if (p.obj.guidance == CINDM_GUIDE_STANDARD) gscale = 1.f;
    else if (p.obj.guidance == CINDM_GUIDE_STANDARD_ALPHA)
        gscale = __fdiv_rn(p.sched[TAB_BETAS * TS + t], __fsqrt_rn(p.sched[TAB_ACP_PREV * TS + t]));
    const int step = p.step_dev ? *p.step_dev : p.t_start - t;      // sampling step index (row of the explicit noise tensor)
    float na, nb;   // out = na * pred + nb * noise
    float d_san = 0.f, d_c = 0.f;
    bool d_last = false;
    if (p.ddim) {
        const float* cf = p.ddim_coef + 4 * step;
        d_san = cf[0]; d_c = cf[1];
        d_last = cf[3] != 0.f;
        na = 1.0f;
        nb = d_last ? 0.f : cf[2];                                      // sigma (0 when eta == 0; NaN on the last pair, unused there)
    } else if (p.renoise) {
        float ratio = __fdiv_rn(p.sched[TAB_ACP * TS + t], p.sched[TAB_ACP_PREV * TS + t]);   // fp32 ratio (:1366)
        na = __fsqrt_rn(ratio);
        nb = __fsqrt_rn(__fsub_rn(1.0f, ratio));
    } else {
        na = 1.0f;
        nb = (t > 0) ? expf(__fmul_rn(0.5f, p.sched[TAB_POST_LOGVAR * TS + t])) : 0.f;       // no noise at t == 0
    }
    const bool have_noise = nb != 0.f && (p.noise != nullptr || p.use_philox);
    const float4* noise = nullptr;
    const int Tn = p.T - p.cond_rows;                               // frames an explicit noise tensor covers
    if (p.noise)
        noise = reinterpret_cast<const float4*>(
            p.noise + ((long long)step * p.draws_per_step + p.draw) * ((long long)p.B * Tn * p.n * 4));

    const long long total = (long long)p.B * p.T * p.n;
    const int F = 4 * p.n;
    const int vec_per_cand = p.T * p.n;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int j = (int)(i % p.n);
        long long bt = i / p.n;
        int tt = (int)(bt % p.T);
        long long b = bt / p.T;
        float4 xv = reinterpret_cast<const float4*>(p.x)[i];
        if (tt < p.cond_rows) {                                   // condition frames: seen by the model, never updated
            reinterpret_cast<float4*>(p.x_out)[i] = xv;
            if (p.pred_out) reinterpret_cast<float4*>(p.pred_out)[i] = xv;
            if (p.x0_out) reinterpret_cast<float4*>(p.x0_out)[i] = xv;
            continue;
        }
        float4 ev = p.mean_in ? make_float4(0.f, 0.f, 0.f, 0.f) : reinterpret_cast<const float4*>(p.eps)[i];
        float xs[4] = {xv.x, xv.y, xv.z, xv.w}, es[4] = {ev.x, ev.y, ev.z, ev.w};
        float mi[4] = {0.f, 0.f, 0.f, 0.f}, x0i[4] = {0.f, 0.f, 0.f, 0.f};
        if (p.mean_in) {                                      // "mean" (outside) composition: posterior already composed
            const float4 m4 = reinterpret_cast<const float4*>(p.mean_in)[i], z4 = reinterpret_cast<const float4*>(p.x0_in)[i];
            mi[0] = m4.x; mi[1] = m4.y; mi[2] = m4.z; mi[3] = m4.w;
            x0i[0] = z4.x; x0i[1] = z4.y; x0i[2] = z4.z; x0i[3] = z4.w;
        }
        float g[4] = {0.f, 0.f, 0.f, 0.f};
        if (gscale != 0.f) {
            float2 gr = objective_grad(p.x, bt * F, tt, p.T, F, j, xv.x, xv.y, p.obj);
            g[0] = gr.x; g[1] = gr.y;
            if (p.obj.guidance == CINDM_GUIDE_STANDARD_ALPHA) { g[0] = __fmul_rn(gscale, g[0]); g[1] = __fmul_rn(gscale, g[1]); }
        }
        float4 nz = make_float4(0.f, 0.f, 0.f, 0.f);
        if (have_noise)
            nz = noise ? noise[(b * Tn + (tt - p.cond_rows)) * p.n + j] : noise4(p.seed, p.cand_off + b, (int)(i - b * vec_per_cand), t, p.draw);
        float ns[4] = {nz.x, nz.y, nz.z, nz.w};
        float x0[4], pr[4], out[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float v = __fsub_rn(__fmul_rn(A, xs[q]), __fmul_rn(Bc, es[q]));                 // (:914-918)
            v = (v != v) ? v : fminf(fmaxf(v, -1.0f), 1.0f);                                // clamp_(-1, 1) (:1039); torch.clamp keeps NaN
            float mu = __fadd_rn(__fmul_rn(c1, v), __fmul_rn(c2, xs[q]));                   // (:943-946)
            if (p.mean_in) { v = x0i[q]; mu = mi[q]; }
            x0[q] = v;
            pr[q] = __fsub_rn(mu, g[q]);                                                    // (:1349)
            if (p.overwrite && tt < p.ow_rows) pr[q] = p.overwrite[((b * p.ow_rows + tt) * p.n + j) * 4 + q];   // (:1361-1362)
            if (p.ddim) {
                // pred_noise + grad_design_final (:1375), then img = x_start * sqrt(alpha_next) + c * pred_noise + sigma * noise (:1786-1788)
                const float pn = __fadd_rn(es[q], g[q]);
                float o = __fadd_rn(__fmul_rn(v, d_san), __fmul_rn(d_c, pn));
                if (have_noise) o = __fadd_rn(o, __fmul_rn(nb, ns[q]));
                out[q] = d_last ? v : o;                                                    // time_next < 0: img = x_start (:1789-1793)
            } else {
                out[q] = have_noise ? __fadd_rn(__fmul_rn(na, pr[q]), __fmul_rn(nb, ns[q])) : __fmul_rn(na, pr[q]);
            }
        }
        reinterpret_cast<float4*>(p.x_out)[i] = make_float4(out[0], out[1], out[2], out[3]);
        if (p.pred_out) reinterpret_cast<float4*>(p.pred_out)[i] = make_float4(pr[0], pr[1], pr[2], pr[3]);
        if (p.x0_out) reinterpret_cast<float4*>(p.x0_out)[i] = make_float4(x0[0], x0[1], x0[2], x0[3]);
    }
}

int launch_update(const UpdateLaunch& u, cudaStream_t st) {
    if (!u.sched) return fail(-4, "schedule tables not set (cindm_set_schedule)");
    if (u.x == u.x_out) return fail(-2, "ddpm update cannot run in place (the consistency gradient is a stencil in t)");
    UpdateParams p;
    p.x = u.x; p.eps = u.eps; p.noise = u.noise; p.x_out = u.x_out; p.pred_out = u.pred_out; p.x0_out = u.x0_out;
    p.sched = u.sched; p.t_dev = u.t_dev; p.cand_off = u.cand_off; p.seed = u.seed;
    p.B = u.B; p.T = u.T; p.n = u.n; p.timesteps = u.timesteps; p.t_host = u.t_host; p.renoise = u.renoise;
    p.t_start = u.t_start; p.draws_per_step = u.draws_per_step; p.draw = u.draw; p.use_philox = u.use_philox;
    p.obj = u.obj;
    p.ddim = u.ddim; p.ddim_coef = u.ddim_coef; p.step_dev = u.step_dev;
    p.mean_in = u.mean_in; p.x0_in = u.x0_in;
    p.cond_rows = u.cond_rows;
    p.overwrite = u.overwrite; p.ow_rows = u.overwrite ? u.ow_rows : 0;
    if (u.overwrite && (u.ow_rows <= 0 || u.ow_rows > u.T || u.ddim)) return fail(-2, "initial_state_overwrite: bad frame count (or DDIM)");
    if (u.cond_rows < 0 || u.cond_rows >= u.T) return fail(-2, "cond_rows must be in [0, T)");
    if ((u.mean_in == nullptr) != (u.x0_in == nullptr)) return fail(-2, "composed posterior mean and x_start come together");
    if (u.mean_in && u.ddim) return fail(-5, "DDIM runs on the *-inside composition only (reference ddim_sample :1758-1771)");
    if (u.ddim && (!u.ddim_coef || !u.step_dev)) return fail(-2, "DDIM update needs the coefficient table and the step counter");
    long long total = (long long)u.B * u.T * u.n;
    if (total == 0) return 0;
    KernelTimer kt("ddpm_update", st, (double)total * 16.0 * (u.noise ? 4.0 : 3.0));
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    CINDM_CHECK_CUDA(launch_chain(ddpm_update_kernel, dim3(blocks), dim3(256), 0, st, p));
    CINDM_CHECK_LAUNCH();
    return 0;
}

// ---------------------------------------------------------------- unadjusted Langevin step (sample_step_ULA :2047-2073)
__global__ void __launch_bounds__(256) ula_step_kernel(const float4* __restrict__ x, const float4* __restrict__ eps,
                                                       const float4* __restrict__ noise, float4* __restrict__ out, long long nvec,
                                                       int vec_per_cand, float grad_scale, float ss, float std, unsigned long long seed,
                                                       long long cand_off, int t, int draw) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
        const long long b = i / vec_per_cand;
        const float4 xv = x[i], ev = eps[i];
        const float4 nz = noise ? noise[i] : noise4(seed, cand_off + b, (int)(i - b * vec_per_cand), t, draw);
        // grad = grad_scale * eps;  x + grad * ss + (randn * std)      (:2056-2061, three rounded operations per term)
        float4 o;
        o.x = __fadd_rn(__fadd_rn(xv.x, __fmul_rn(__fmul_rn(grad_scale, ev.x), ss)), __fmul_rn(nz.x, std));
        o.y = __fadd_rn(__fadd_rn(xv.y, __fmul_rn(__fmul_rn(grad_scale, ev.y), ss)), __fmul_rn(nz.y, std));
        o.z = __fadd_rn(__fadd_rn(xv.z, __fmul_rn(__fmul_rn(grad_scale, ev.z), ss)), __fmul_rn(nz.z, std));
        o.w = __fadd_rn(__fadd_rn(xv.w, __fmul_rn(__fmul_rn(grad_scale, ev.w), ss)), __fmul_rn(nz.w, std));
        out[i] = o;
    }
}

int launch_ula_step(const float* x, const float* eps, const float* noise, float* out, int B, int T, int n, float grad_scale,
                    float ss, uint64_t seed, int64_t cand_off, int t, int draw, cudaStream_t st) {
    const long long nvec = (long long)B * T * n;
    if (nvec == 0) return 0;
    int blocks = (int)((nvec + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    const float std = sqrtf(2.0f * ss);                              // (2 * ss) ** .5 on an fp32 tensor element (:2055)
    ula_step_kernel<<<blocks, 256, 0, st>>>((const float4*)x, (const float4*)eps, (const float4*)noise, (float4*)out, nvec, T * n,
                                            grad_scale, ss, std, seed, cand_off, t, draw);
    CINDM_CHECK_LAUNCH();
    return 0;
}

// ---------------------------------------------------------------- x_start from epsilon (predict_start_from_noise :914-918)
__global__ void __launch_bounds__(256) predict_start_kernel(const float4* __restrict__ x, const float4* __restrict__ eps,
                                                            float4* __restrict__ x0, long long nvec, const float* __restrict__ sched,
                                                            int timesteps, int t, int clip) {
    const float A = sched[TAB_SQRT_RECIP_ACP * timesteps + t], Bc = sched[TAB_SQRT_RECIPM1_ACP * timesteps + t];
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
        const float4 a = x[i], b = eps[i];
        float v[4] = {__fsub_rn(__fmul_rn(A, a.x), __fmul_rn(Bc, b.x)), __fsub_rn(__fmul_rn(A, a.y), __fmul_rn(Bc, b.y)),
                      __fsub_rn(__fmul_rn(A, a.z), __fmul_rn(Bc, b.z)), __fsub_rn(__fmul_rn(A, a.w), __fmul_rn(Bc, b.w))};
        if (clip) {
#pragma unroll
            for (int q = 0; q < 4; ++q) v[q] = (v[q] != v[q]) ? v[q] : fminf(fmaxf(v[q], -1.0f), 1.0f);   // torch.clamp keeps NaN
        }
        x0[i] = make_float4(v[0], v[1], v[2], v[3]);
    }
}

int launch_predict_start(const float* x, const float* eps, float* x0, long long elems, const float* sched, int timesteps, int t,
                         int clip, cudaStream_t st) {
    if (!sched) return fail(-4, "schedule tables not set (cindm_set_schedule)");
    if (elems % 4) return fail(-2, "element count must be a multiple of 4");
    const long long nvec = elems / 4;
    if (nvec == 0) return 0;
    int blocks = (int)((nvec + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    predict_start_kernel<<<blocks, 256, 0, st>>>((const float4*)x, (const float4*)eps, (float4*)x0, nvec, sched, timesteps, t, clip);
    CINDM_CHECK_LAUNCH();
    return 0;
}

// ---------------------------------------------------------------- composed epsilon + loop
int composed_eps(cindm_engine* e, const float* x, float* eps, int B, int n, int nc, int start, int mode, int t,
                 const int* t_dev, int prec, int conv_engine, cudaStream_t st, float* x0_composed, float ebm_coef) {
    const int H = e->cfg.horizon;
    if (e->cfg.transition_dim != 8) return fail(-2, "the composition operator runs on the body-pair model (transition_dim 8)");
    if (mode == CINDM_COMPOSE_EBM) {
        // gradient() (:1856-1982): pair terms summed per receiver (one window), minus coef x the unconditional single-body model
        cindm_engine* u = e->uncond;
        if (!u) return fail(-4, "compose mode EBM needs an unconditional single-body engine (cindm_attach_unconditioned)");
        if (nc != 0) return fail(-5, "the EBM body composition has one window (n_composed = 0)");
        if (n < 3) return fail(-2, "the EBM body composition needs at least 3 bodies");
        const int64_t S1 = (int64_t)n * B;
        if (S1 > u->ws.max_slices || u->ws.precision != prec)
            return fail(-6, "unconditional engine: workspace not reserved for this slice count / precision");
        CINDM_TRY(composed_eps(e, x, eps, B, n, 0, start, CINDM_COMPOSE_SUM_INSIDE, t, t_dev, prec, conv_engine, st));
        if (prec == PREC_F32) {
            CINDM_TRY(launch_body_gather(x, u->ws.slices, B, n, H, H, st));
            CINDM_TRY(unet_forward(u, u->ws.slices, S1, t, t_dev, u->ws.eps_pair, prec, conv_engine, st));
        } else {
            GatherSpec gs{x, B, n, 0, start};
            CINDM_TRY(unet_forward(u, nullptr, S1, t, t_dev, u->ws.eps_pair, prec, conv_engine, st, &gs));
        }
        return launch_ebm_subtract(eps, u->ws.eps_pair, B, n, H, ebm_coef, st);
    }
    if (mode == CINDM_COMPOSE_NOISE_SUM) mode = CINDM_COMPOSE_SUM_INSIDE;      // same operator (:1452-1457 vs :997-999)
    if (mode == CINDM_COMPOSE_MEAN_OUTSIDE && !x0_composed)
        return fail(-5, "compose_mode 'mean' composes the posterior, not epsilon: there is no composed epsilon to return");
    if (mode < 0 || mode > CINDM_COMPOSE_NOISE_SUM) return fail(-2, "bad compose_mode");
    if (n < 2) return fail(-2, "compose_n_bodies must be at least 2");
    if (nc < 0 || start <= 0) return fail(-2, "bad composition window parameters");
    const int64_t S = (int64_t)(nc + 1) * (n * (n - 1) / 2) * B;
    if (S > e->ws.max_slices || e->ws.precision != prec)
        return fail(-6, "workspace not reserved for this slice count / precision (call cindm_reserve)");
    if (prec == PREC_F32) {
        CINDM_TRY(launch_compose_gather(x, e->ws.slices, B, n, nc, start, H, st));
        CINDM_TRY(unet_forward(e, e->ws.slices, S, t, t_dev, e->ws.eps_pair, prec, conv_engine, st));
    } else {
        GatherSpec gs{x, B, n, nc, start};          // the stem kernel gathers straight from x
        CINDM_TRY(unet_forward(e, nullptr, S, t, t_dev, e->ws.eps_pair, prec, conv_engine, st, &gs));
    }
    if (mode == CINDM_COMPOSE_MEAN_OUTSIDE)      // `eps` receives the composed posterior mean, x0_composed the composed x_start
        return launch_compose_scatter_posterior(e->ws.eps_pair, x, eps, x0_composed, B, n, nc, start, H, e->sched_dev,
                                                e->cfg.timesteps, t, t_dev, st);
    return launch_compose_scatter(e->ws.eps_pair, eps, B, n, nc, start, H, mode, st);
}

void graph_cache_clear(cindm_engine* e) {
    if (e->sb.graph_exec) {
        cudaDeviceSynchronize();
        cudaGraphExecDestroy(e->sb.graph_exec);
        e->sb.graph_exec = nullptr;
    }
    e->sb.graph_key.clear();
    e->sb.graph_nodes = 0;
}

static int ensure_sample_buffers(cindm_engine* e, size_t elems) {
    SampleBuffers& sb = e->sb;
    if (!sb.t_dev) CINDM_CHECK_CUDA(cudaMalloc(&sb.t_dev, sizeof(int)));
    if (sb.elems >= elems) return 0;
    CINDM_CHECK_CUDA(cudaDeviceSynchronize());
    graph_cache_clear(e);
    if (sb.x_alt) cudaFree(sb.x_alt);
    if (sb.eps) cudaFree(sb.eps);
    if (sb.x0c) cudaFree(sb.x0c);
    sb.x_alt = sb.eps = sb.x0c = nullptr;
    CINDM_CHECK_CUDA(cudaMalloc(&sb.x_alt, elems * sizeof(float)));
    CINDM_CHECK_CUDA(cudaMalloc(&sb.eps, elems * sizeof(float)));
    CINDM_CHECK_CUDA(cudaMalloc(&sb.x0c, elems * sizeof(float)));
    sb.elems = elems;
    return 0;
}

// Issue the kernels of ONE sampling step.  x lives in bufs[cur]; returns (through cur) where the result is.
// ddim: the step is one (time, time_next) pair of ddim_sample (:1751-1797): the timestep comes from the device table
// indexed by the device step counter, the recurrence runs as in the DDPM step, and the last evaluation feeds the DDIM
// update.  Random draws per step, in the reference's order: R re-noise draws, the (unused) posterior noise, the DDIM noise.
static int issue_step(cindm_engine* e, const cindm_sample_config& c, float* bufs[2], int& cur, const float* noise,
                      float* x0_out, int t, const int* t_dev, cudaStream_t st, bool ddim = false) {
    const int T = e->cfg.horizon + c.n_composed * c.compose_start_step;
    const bool guided = c.objective.guidance != CINDM_GUIDE_NONE;
    const int iters = ddim ? (guided ? c.recurrence : 1) : (c.recurrence > 0 ? c.recurrence : 1);
    const int draws = ddim ? (guided ? c.recurrence + 2 : 1) : (c.recurrence > 0 ? c.recurrence + 1 : 1);
    if (ddim) {
        CINDM_CHECK_CUDA(launch_chain(ddim_set_time_kernel, dim3(1), dim3(1), 0, st, e->sb.t_dev, (const int*)e->sb.ddim_times,
                                      (const int*)e->sb.step_dev));
        CINDM_CHECK_LAUNCH();
    }
    if (c.chain_blocks > 1)
        CINDM_TRY(launch_chain_condition(bufs[cur], c.batch / c.chain_blocks, c.chain_blocks, T, c.n_bodies, c.cond_rows, st));
    for (int r = 0; r < iters; ++r) {
        const bool outside = c.compose_mode == CINDM_COMPOSE_MEAN_OUTSIDE;
        CINDM_TRY(composed_eps(e, bufs[cur], e->sb.eps, c.batch, c.n_bodies, c.n_composed, c.compose_start_step,
                               c.compose_mode, t, t_dev, c.precision, c.conv_engine, st, outside ? e->sb.x0c : nullptr,
                               c.ebm_uncond_coef));
        const bool last = r == iters - 1;
        UpdateLaunch u;
        u.x = bufs[cur]; u.eps = e->sb.eps; u.x_out = bufs[cur ^ 1];
        if (outside) { u.mean_in = e->sb.eps; u.x0_in = e->sb.x0c; }
        u.pred_out = nullptr; u.x0_out = last ? x0_out : nullptr;
        u.B = c.batch; u.T = T; u.n = c.n_bodies; u.sched = e->sched_dev; u.timesteps = e->cfg.timesteps;
        u.t_dev = t_dev; u.t_host = t;
        u.noise = noise; u.t_start = c.t_start; u.draws_per_step = draws;
        u.use_philox = noise == nullptr; u.seed = c.seed; u.cand_off = c.candidate_offset;
        u.obj = c.objective;
        u.cond_rows = c.cond_rows;
        u.overwrite = e->overwrite; u.ow_rows = e->overwrite_rows;
        // the reference re-noises after the last recurrence too and then discards it (:1365-1370):
        // the last evaluation goes straight to the final posterior noise (draw id R)
        u.renoise = (c.recurrence > 0 && !last) ? 1 : 0;
        u.draw = (c.recurrence > 0 && last) ? c.recurrence : r;
        if (ddim) {
            u.step_dev = e->sb.step_dev;
            if (last) { u.ddim = 1; u.ddim_coef = e->sb.ddim_coef; u.renoise = 0; u.draw = guided ? c.recurrence + 1 : 0; }
            else { u.renoise = 1; u.draw = r; }
        }
        CINDM_TRY(launch_update(u, st));
        cur ^= 1;
    }
    if (ddim) CINDM_TRY(launch_step_counter(e->sb.step_dev, 1, st));
    return 0;
}

static int check_condition_config(const cindm_sample_config& c, int T) {
    if (c.cond_rows < 0 || c.cond_rows >= T) return fail(-2, "cond_rows must be in [0, T)");
    if (c.chain_blocks > 1 && (c.cond_rows == 0 || c.batch % c.chain_blocks != 0))
        return fail(-2, "chain_blocks needs cond_rows > 0 and a batch that is a multiple of chain_blocks");
    if (c.cond_rows > 0 && c.objective.guidance != CINDM_GUIDE_NONE)
        return fail(-5, "design guidance on a conditioned model is not on the CUDA fast path");
    return 0;
}

static int sample_loop_on(cindm_engine* e, const cindm_sample_config& c, float* x, const float* noise, float* x0_out,
                          cudaStream_t st);

int sample_loop(cindm_engine* e, const cindm_sample_config& c, float* x, const float* noise, float* x0_out,
                cudaStream_t caller) {
    // the legacy / per-thread default streams cannot be captured: run on an engine-owned stream that is
    // ordered after the caller's stream, and make the caller's stream wait for it
    const bool special = caller == nullptr || caller == cudaStreamLegacy || caller == cudaStreamPerThread;
    if (!c.use_graph || !special) return sample_loop_on(e, c, x, noise, x0_out, caller);
    SampleBuffers& sb = e->sb;
    if (!sb.capture_stream) {
        CINDM_CHECK_CUDA(cudaStreamCreateWithFlags(&sb.capture_stream, cudaStreamNonBlocking));
        CINDM_CHECK_CUDA(cudaEventCreateWithFlags(&sb.ev_in, cudaEventDisableTiming));
        CINDM_CHECK_CUDA(cudaEventCreateWithFlags(&sb.ev_out, cudaEventDisableTiming));
    }
    CINDM_CHECK_CUDA(cudaEventRecord(sb.ev_in, caller));
    CINDM_CHECK_CUDA(cudaStreamWaitEvent(sb.capture_stream, sb.ev_in, 0));
    int rc = sample_loop_on(e, c, x, noise, x0_out, sb.capture_stream);
    CINDM_CHECK_CUDA(cudaEventRecord(sb.ev_out, sb.capture_stream));
    CINDM_CHECK_CUDA(cudaStreamWaitEvent(caller, sb.ev_out, 0));
    return rc;
}

static int sample_loop_on(cindm_engine* e, const cindm_sample_config& c, float* x, const float* noise, float* x0_out,
                          cudaStream_t st) {
    if (!e->finalized) return fail(-4, "weights not finalized");
    if (!e->sched_dev) return fail(-4, "schedule tables not set (cindm_set_schedule)");
    if (c.t_start >= e->cfg.timesteps || c.t_end < 0 || c.t_end > c.t_start) return fail(-2, "bad timestep range");
    if (c.compose_start_step >= e->cfg.horizon) return fail(-2, "compose_start_step must be < horizon");   // (:1679)
    if (c.batch <= 0) return fail(-2, "batch must be positive");
    if (e->cfg.transition_dim != 8) return fail(-2, "sampling runs on the body-pair engine (transition_dim 8)");
    if (c.compose_mode == CINDM_COMPOSE_EBM && c.objective.guidance != CINDM_GUIDE_NONE)
        return fail(-5, "design guidance with the EBM body composition is not on the CUDA fast path");
    const int T = e->cfg.horizon + c.n_composed * c.compose_start_step;
    CINDM_TRY(check_condition_config(c, T));
    const size_t elems = (size_t)c.batch * T * c.n_bodies * 4;
    const int64_t S = (int64_t)(c.n_composed + 1) * (c.n_bodies * (c.n_bodies - 1) / 2) * c.batch;
    CINDM_TRY(reserve_workspace(e, S > e->ws.max_slices ? S : e->ws.max_slices, c.precision));
    if (c.compose_mode == CINDM_COMPOSE_EBM) {
        if (!e->uncond) return fail(-4, "compose mode EBM needs an unconditional single-body engine (cindm_attach_unconditioned)");
        if (!e->uncond->finalized) return fail(-4, "unconditional engine: weights not finalized");
        const int64_t S1 = (int64_t)c.n_bodies * c.batch;
        CINDM_TRY(reserve_workspace(e->uncond, S1 > e->uncond->ws.max_slices ? S1 : e->uncond->ws.max_slices, c.precision));
    }
    CINDM_TRY(ensure_sample_buffers(e, elems));
    float* bufs[2] = {x, e->sb.x_alt};
    int cur = 0;
    const int n_steps = c.t_start - c.t_end + 1;

    if (!c.use_graph) {
        for (int t = c.t_start; t >= c.t_end; --t)
            CINDM_TRY(issue_step(e, c, bufs, cur, noise, x0_out, t, nullptr, st));
    } else {
        // One graph = two DDPM steps when a step flips the ping-pong parity, so that every replay
        // starts with x in bufs[0]; t is read from device memory and decremented inside the graph.
        const int iters = c.recurrence > 0 ? c.recurrence : 1;
        const int steps_per_graph = (iters % 2) ? 2 : 1;
        CINDM_CHECK_CUDA(cudaMemcpyAsync(e->sb.t_dev, &c.t_start, sizeof(int), cudaMemcpyHostToDevice, st));
        // the captured step depends on everything but the timestep range (t lives in device memory), unless an
        // explicit noise tensor is indexed relative to t_start
        cindm_sample_config kc = c;
        if (!noise) { kc.t_start = 0; kc.t_end = 0; }
        std::string key(reinterpret_cast<const char*>(&kc), sizeof(kc));
        const void* ptrs[6] = {x, noise, x0_out, (const void*)st, e->overwrite, (const void*)(intptr_t)e->overwrite_rows};
        key.append(reinterpret_cast<const char*>(ptrs), sizeof(ptrs));
        SampleBuffers& sb = e->sb;
        if (!sb.graph_exec || sb.graph_key != key) {
            graph_cache_clear(e);
            cudaGraph_t graph = nullptr;
            CINDM_CHECK_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
            const long long before = launch_count();
            int rc = 0, gcur = 0;
            for (int k = 0; k < steps_per_graph && rc == 0; ++k) {
                rc = issue_step(e, c, bufs, gcur, noise, x0_out, 0, e->sb.t_dev, st);
                if (rc == 0) rc = launch_step_counter(e->sb.t_dev, -1, st);
            }
            const long long nodes = launch_count() - before;
            add_launches(-nodes);                                  // captured, not executed
            cudaError_t ce = cudaStreamEndCapture(st, &graph);
            if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
            if (ce != cudaSuccess) return fail(-100, std::string("graph capture: ") + cudaGetErrorString(ce));
            ce = cudaGraphInstantiate(&sb.graph_exec, graph, 0);
            cudaGraphDestroy(graph);
            if (ce != cudaSuccess) { sb.graph_exec = nullptr; return fail(-100, std::string("graph instantiate: ") + cudaGetErrorString(ce)); }
            sb.graph_key = key;
            sb.graph_nodes = nodes;
        }
        int done = 0;
        for (; done + steps_per_graph <= n_steps; done += steps_per_graph) {
            CINDM_CHECK_CUDA(cudaGraphLaunch(sb.graph_exec, st));
            add_launches(sb.graph_nodes);
        }
        for (int t = c.t_start - done; t >= c.t_end; --t)       // odd remainder, issued directly
            CINDM_TRY(issue_step(e, c, bufs, cur, noise, x0_out, t, nullptr, st));
    }
    if (cur != 0) CINDM_CHECK_CUDA(cudaMemcpyAsync(x, e->sb.x_alt, elems * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return 0;
}

// ---------------------------------------------------------------- DDIM loop (sampling_timesteps < timesteps)
static int sample_ddim_on(cindm_engine* e, const cindm_sample_config& c, int n_pairs, const int32_t* times,
                          const int32_t* times_next, const float* coef3, float* x, const float* noise, float* x0_out,
                          cudaStream_t st) {
    if (!e->finalized) return fail(-4, "weights not finalized");
    if (!e->sched_dev) return fail(-4, "schedule tables not set (cindm_set_schedule)");
    if (n_pairs <= 0 || !times || !times_next || !coef3) return fail(-2, "DDIM needs the (time, time_next) pairs and their coefficients");
    if (c.compose_start_step >= e->cfg.horizon) return fail(-2, "compose_start_step must be < horizon");
    if (c.batch <= 0) return fail(-2, "batch must be positive");
    const bool guided = c.objective.guidance != CINDM_GUIDE_NONE;
    // only the recurrence branch of p_sample_compose_inside returns (pred_noise + grad, x_start) (:1372-1376); the
    // single-pass "standard" branch hands ddim_sample the posterior sample in place of epsilon (:1283)
    if (c.compose_mode != CINDM_COMPOSE_MEAN_INSIDE && c.compose_mode != CINDM_COMPOSE_SUM_INSIDE)
        return fail(-5, "DDIM runs on the *-inside composition only (reference ddim_sample :1758-1771)");
    if (e->cfg.transition_dim != 8) return fail(-2, "sampling runs on the body-pair engine (transition_dim 8)");
    if (guided && c.recurrence <= 0)
        return fail(-5, "DDIM sampling with guidance needs a '-recurrence-K' design_guidance (reference :1283 vs :1372-1376)");
    for (int i = 0; i < n_pairs; ++i)
        if (times[i] < 0 || times[i] >= e->cfg.timesteps) return fail(-2, "DDIM timestep out of range");
    const int T = e->cfg.horizon + c.n_composed * c.compose_start_step;
    CINDM_TRY(check_condition_config(c, T));
    const size_t elems = (size_t)c.batch * T * c.n_bodies * 4;
    const int64_t S = (int64_t)(c.n_composed + 1) * (c.n_bodies * (c.n_bodies - 1) / 2) * c.batch;
    CINDM_TRY(reserve_workspace(e, S > e->ws.max_slices ? S : e->ws.max_slices, c.precision));
    CINDM_TRY(ensure_sample_buffers(e, elems));
    SampleBuffers& sb = e->sb;
    if (!sb.step_dev) CINDM_CHECK_CUDA(cudaMalloc(&sb.step_dev, sizeof(int)));
    if (sb.ddim_capacity < n_pairs) {
        CINDM_CHECK_CUDA(cudaStreamSynchronize(st));
        graph_cache_clear(e);
        if (sb.ddim_times) cudaFree(sb.ddim_times);
        if (sb.ddim_coef) cudaFree(sb.ddim_coef);
        sb.ddim_times = nullptr; sb.ddim_coef = nullptr; sb.ddim_capacity = 0;
        CINDM_CHECK_CUDA(cudaMalloc(&sb.ddim_times, (size_t)n_pairs * sizeof(int)));
        CINDM_CHECK_CUDA(cudaMalloc(&sb.ddim_coef, (size_t)n_pairs * 4 * sizeof(float)));
        sb.ddim_capacity = n_pairs;
    }
    std::vector<float> coef4((size_t)n_pairs * 4);
    for (int i = 0; i < n_pairs; ++i) {
        coef4[4 * i] = coef3[3 * i]; coef4[4 * i + 1] = coef3[3 * i + 1]; coef4[4 * i + 2] = coef3[3 * i + 2];
        coef4[4 * i + 3] = times_next[i] < 0 ? 1.f : 0.f;
    }
    const int zero = 0;
    // pageable sources: these copies return once the host buffers have been staged
    CINDM_CHECK_CUDA(cudaMemcpyAsync(sb.ddim_times, times, (size_t)n_pairs * sizeof(int), cudaMemcpyHostToDevice, st));
    CINDM_CHECK_CUDA(cudaMemcpyAsync(sb.ddim_coef, coef4.data(), coef4.size() * sizeof(float), cudaMemcpyHostToDevice, st));
    CINDM_CHECK_CUDA(cudaMemcpyAsync(sb.step_dev, &zero, sizeof(int), cudaMemcpyHostToDevice, st));
    CINDM_CHECK_CUDA(cudaStreamSynchronize(st));

    float* bufs[2] = {x, sb.x_alt};
    int cur = 0;
    const int iters = guided ? c.recurrence : 1;
    if (!c.use_graph) {
        for (int i = 0; i < n_pairs; ++i) CINDM_TRY(issue_step(e, c, bufs, cur, noise, x0_out, 0, sb.t_dev, st, true));
    } else {
        const int steps_per_graph = (iters % 2) ? 2 : 1;
        cindm_sample_config kc = c;
        kc.t_start = -1; kc.t_end = -1;                              // marks a DDIM graph; the pairs live in device tables
        std::string key(reinterpret_cast<const char*>(&kc), sizeof(kc));
        const void* ptrs[4] = {x, noise, x0_out, (const void*)st};
        key.append(reinterpret_cast<const char*>(ptrs), sizeof(ptrs));
        if (!sb.graph_exec || sb.graph_key != key) {
            graph_cache_clear(e);
            cudaGraph_t graph = nullptr;
            CINDM_CHECK_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
            const long long before = launch_count();
            int rc = 0, gcur = 0;
            for (int k = 0; k < steps_per_graph && rc == 0; ++k) rc = issue_step(e, c, bufs, gcur, noise, x0_out, 0, sb.t_dev, st, true);
            const long long nodes = launch_count() - before;
            add_launches(-nodes);
            cudaError_t ce = cudaStreamEndCapture(st, &graph);
            if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
            if (ce != cudaSuccess) return fail(-100, std::string("graph capture: ") + cudaGetErrorString(ce));
            ce = cudaGraphInstantiate(&sb.graph_exec, graph, 0);
            cudaGraphDestroy(graph);
            if (ce != cudaSuccess) { sb.graph_exec = nullptr; return fail(-100, std::string("graph instantiate: ") + cudaGetErrorString(ce)); }
            sb.graph_key = key;
            sb.graph_nodes = nodes;
        }
        int done = 0;
        for (; done + steps_per_graph <= n_pairs; done += steps_per_graph) {
            CINDM_CHECK_CUDA(cudaGraphLaunch(sb.graph_exec, st));
            add_launches(sb.graph_nodes);
        }
        for (; done < n_pairs; ++done) CINDM_TRY(issue_step(e, c, bufs, cur, noise, x0_out, 0, sb.t_dev, st, true));
    }
    if (cur != 0) CINDM_CHECK_CUDA(cudaMemcpyAsync(x, sb.x_alt, elems * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return 0;
}

int sample_ddim(cindm_engine* e, const cindm_sample_config& c, int n_pairs, const int32_t* times, const int32_t* times_next,
                const float* coef3, float* x, const float* noise, float* x0_out, cudaStream_t caller) {
    const bool special = caller == nullptr || caller == cudaStreamLegacy || caller == cudaStreamPerThread;
    if (!c.use_graph || !special) return sample_ddim_on(e, c, n_pairs, times, times_next, coef3, x, noise, x0_out, caller);
    SampleBuffers& sb = e->sb;
    if (!sb.capture_stream) {
        CINDM_CHECK_CUDA(cudaStreamCreateWithFlags(&sb.capture_stream, cudaStreamNonBlocking));
        CINDM_CHECK_CUDA(cudaEventCreateWithFlags(&sb.ev_in, cudaEventDisableTiming));
        CINDM_CHECK_CUDA(cudaEventCreateWithFlags(&sb.ev_out, cudaEventDisableTiming));
    }
    CINDM_CHECK_CUDA(cudaEventRecord(sb.ev_in, caller));
    CINDM_CHECK_CUDA(cudaStreamWaitEvent(sb.capture_stream, sb.ev_in, 0));
    int rc = sample_ddim_on(e, c, n_pairs, times, times_next, coef3, x, noise, x0_out, sb.capture_stream);
    CINDM_CHECK_CUDA(cudaEventRecord(sb.ev_out, sb.capture_stream));
    CINDM_CHECK_CUDA(cudaStreamWaitEvent(caller, sb.ev_out, 0));
    return rc;
}

}  // namespace cindm
