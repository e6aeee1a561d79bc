// The composition operator of GaussianDiffusion1D.model_predictions (reference
// model/diffusion_1d.py:959-1001) as two bandwidth-bound kernels over precomputed index maps:
//   gather : x[B][T][4n]            -> slices[(kk*P+p)*B + b][24][8]
//   scatter: eps_pair[S][24][8]     -> eps[B][T][4n]   (mean over senders, mean over covering windows)
// Both move 16-byte (one body's x,y,vx,vy) vectors; no atomics: every output element gathers
// its own contributions in a fixed order, so the result is deterministic.
#include "engine.h"

namespace cindm {

// pair index of (ii < jj) in lexicographic order
__host__ __device__ __forceinline__ int pair_index(int ii, int jj, int n) {
    return ii * (2 * n - ii - 1) / 2 + (jj - ii - 1);
}

__global__ void __launch_bounds__(256) compose_gather_kernel(const float4* __restrict__ x, float4* __restrict__ slices,
                                                             int B, int n, int W, int start, int H, int T) {
    // one thread per (slice, row, body-half): 16 bytes
    const int P = n * (n - 1) / 2;
    const long long total = (long long)W * P * B * H * 2;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        int half = (int)(i & 1);
        long long r = i >> 1;
        int h = (int)(r % H);
        long long s = r / H;
        int b = (int)(s % B);
        int wp = (int)(s / B);
        int p = wp % P, kk = wp / P;
        // invert the lexicographic pair index
        int ii = 0, rem = p;
        while (rem >= n - 1 - ii) { rem -= n - 1 - ii; ++ii; }
        int jj = ii + 1 + rem;
        int body = half ? jj : ii;
        slices[i] = x[((long long)b * T + kk * start + h) * n + body];
    }
}

__global__ void __launch_bounds__(256) compose_scatter_kernel(const float4* __restrict__ eps_pair, float4* __restrict__ eps,
                                                              int B, int n, int W, int start, int H, int T, int mode) {
    pdl_wait();
    pdl_trigger();
    // one thread per (b, t, receiver): 16 bytes out, (n-1) * cover(t) 16-byte reads
    const int P = n * (n - 1) / 2;
    const long long total = (long long)B * T * n;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        int r = (int)(i % n);
        long long bt = i / n;
        int t = (int)(bt % T);
        int b = (int)(bt / T);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        int cover = 0;
        for (int kk = 0; kk < W; ++kk) {
            int h = t - kk * start;
            if (h < 0 || h >= H) continue;
            ++cover;
            float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int s = 0; s < n; ++s) {
                if (s == r) continue;
                int ii = r < s ? r : s, jj = r < s ? s : r;
                int p = pair_index(ii, jj, n);
                int half = (r == ii) ? 0 : 1;          // eps[..., :4] -> smaller index, [..., 4:] -> larger
                long long slice = ((long long)kk * P + p) * B + b;
                float4 v = eps_pair[(slice * H + h) * 2 + half];
                w.x += v.x; w.y += v.y; w.z += v.z; w.w += v.w;
            }
            if (mode == CINDM_COMPOSE_MEAN_INSIDE) {
                float d = (float)(n - 1);
                w.x /= d; w.y /= d; w.z /= d; w.w /= d;
            }
            acc.x += w.x; acc.y += w.y; acc.z += w.z; acc.w += w.w;
        }
        // mean-inside: / mask.sum(0) = cover;  sum-inside: / mask.mean(0) = cover / W   (:996, :999)
        float div = mode == CINDM_COMPOSE_MEAN_INSIDE ? (float)cover : (float)cover / (float)W;
        acc.x /= div; acc.y /= div; acc.z /= div; acc.w /= div;
        eps[i] = acc;
    }
}

// compose_mode "mean" of p_sample_compose_outside (reference :1414-1451): each (window, pair) slice has its own
// x_start = clamp(A_t x - B_t eps) and posterior mean c1 x_start + c2 x (p_mean_variance on the slice, :1427-1433);
// both are summed over senders, / (n-1), summed over covering windows, / cover.
__global__ void __launch_bounds__(256) compose_scatter_posterior_kernel(
    const float4* __restrict__ eps_pair, const float4* __restrict__ x, float4* __restrict__ mean_out,
    float4* __restrict__ x0_out, int B, int n, int W, int start, int H, int T, const float* __restrict__ sched, int TS,
    int t_host, const int* __restrict__ t_dev) {
    pdl_wait();
    pdl_trigger();
    const int tt = t_dev ? *t_dev : t_host;
    const float A = sched[TAB_SQRT_RECIP_ACP * TS + tt], Bc = sched[TAB_SQRT_RECIPM1_ACP * TS + tt];
    const float c1 = sched[TAB_POST_C1 * TS + tt], c2 = sched[TAB_POST_C2 * TS + tt];
    const int P = n * (n - 1) / 2;
    const long long total = (long long)B * T * n;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        int r = (int)(i % n);
        long long bt = i / n;
        int t = (int)(bt % T);
        int b = (int)(bt / T);
        const float4 xv4 = x[i];
        const float xv[4] = {xv4.x, xv4.y, xv4.z, xv4.w};
        float am[4] = {0.f, 0.f, 0.f, 0.f}, a0[4] = {0.f, 0.f, 0.f, 0.f};
        int cover = 0;
        for (int kk = 0; kk < W; ++kk) {
            int h = t - kk * start;
            if (h < 0 || h >= H) continue;
            ++cover;
            float wm[4] = {0.f, 0.f, 0.f, 0.f}, w0[4] = {0.f, 0.f, 0.f, 0.f};
            for (int s = 0; s < n; ++s) {
                if (s == r) continue;
                int ii = r < s ? r : s, jj = r < s ? s : r;
                int p = pair_index(ii, jj, n);
                int half = (r == ii) ? 0 : 1;
                long long slice = ((long long)kk * P + p) * B + b;
                const float4 e4 = eps_pair[(slice * H + h) * 2 + half];
                const float ev[4] = {e4.x, e4.y, e4.z, e4.w};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    float v = __fsub_rn(__fmul_rn(A, xv[q]), __fmul_rn(Bc, ev[q]));
                    v = (v != v) ? v : fminf(fmaxf(v, -1.0f), 1.0f);          // x_start.clamp_(-1, 1): torch.clamp keeps NaN
                    w0[q] += v;
                    wm[q] += __fadd_rn(__fmul_rn(c1, v), __fmul_rn(c2, xv[q]));
                }
            }
            const float d = (float)(n - 1);
#pragma unroll
            for (int q = 0; q < 4; ++q) { am[q] += wm[q] / d; a0[q] += w0[q] / d; }
        }
        const float div = (float)cover;
        mean_out[i] = make_float4(am[0] / div, am[1] / div, am[2] / div, am[3] / div);
        x0_out[i] = make_float4(a0[0] / div, a0[1] / div, a0[2] / div, a0[3] / div);
    }
}

int launch_compose_scatter_posterior(const float* eps_pair, const float* x, float* mean_out, float* x0_out, int B, int n,
                                     int nc, int start, int H, const float* sched, int timesteps, int t, const int* t_dev,
                                     cudaStream_t st) {
    if (!sched) return fail(-4, "schedule tables not set (cindm_set_schedule)");
    const int W = nc + 1, T = H + nc * start;
    long long total = (long long)B * T * n;
    if (total == 0) return 0;
    KernelTimer kt("compose_scatter", st, (double)total * 48.0 + (double)W * (n * (n - 1) / 2) * B * H * 32.0);
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    CINDM_CHECK_CUDA(launch_chain(compose_scatter_posterior_kernel, dim3(blocks), dim3(256), 0, st, (const float4*)eps_pair,
                                  (const float4*)x, (float4*)mean_out, (float4*)x0_out, B, n, W, start, H, T, sched, timesteps, t, t_dev));
    CINDM_CHECK_LAUNCH();
    return 0;
}

int launch_compose_gather(const float* x, float* slices, int B, int n, int nc, int start, int H, cudaStream_t st) {
    const int W = nc + 1, T = H + nc * start, P = n * (n - 1) / 2;
    long long total = (long long)W * P * B * H * 2;
    if (total == 0) return 0;
    KernelTimer kt("compose_gather", st, (double)total * 16.0 + (double)B * T * n * 16.0);
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    compose_gather_kernel<<<blocks, 256, 0, st>>>((const float4*)x, (float4*)slices, B, n, W, start, H, T);
    CINDM_CHECK_LAUNCH();
    return 0;
}

int launch_compose_scatter(const float* eps_pair, float* eps, int B, int n, int nc, int start, int H, int mode,
                           cudaStream_t st) {
    const int W = nc + 1, T = H + nc * start;
    long long total = (long long)B * T * n;
    if (total == 0) return 0;
    KernelTimer kt("compose_scatter", st, (double)total * 16.0 + (double)W * (n * (n - 1) / 2) * B * H * 32.0);
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    CINDM_CHECK_CUDA(launch_chain(compose_scatter_kernel, dim3(blocks), dim3(256), 0, st, (const float4*)eps_pair, (float4*)eps,
                                  B, n, W, start, H, T, mode));
    CINDM_CHECK_LAUNCH();
    return 0;
}

// ------------------------------------------------------------------------------------------
// EBM body composition with the unconditional single-body model (reference gradient() :1856-1982):
//   eps[b][t][4r + f] = sum_{s != r} eps_pair({r, s})[slot(r)][f]  -  coef * eps_single(body r)[f]
// The pair sum is the "sum-inside" operator with one window; the two kernels below feed and subtract the second term.
// Single-body slice order: slice = body * B + b (the reference calls the unconditional model once per body, :1895-1898).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) body_gather_kernel(const float4* __restrict__ x, float4* __restrict__ slices, int B, int n,
                                                          int H, int T) {
    const long long total = (long long)n * B * H;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int h = (int)(i % H);
        const long long s = i / H;
        const int b = (int)(s % B), body = (int)(s / B);
        slices[i] = x[((long long)b * T + h) * n + body];
    }
}

__global__ void __launch_bounds__(256) ebm_subtract_kernel(float4* __restrict__ eps, const float4* __restrict__ eps_single, int B, int n,
                                                           int T, float coef) {
    pdl_wait();
    pdl_trigger();
    const long long total = (long long)B * T * n;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(i % n);
        const long long bt = i / n;
        const int t = (int)(bt % T);
        const long long b = bt / T;
        const float4 u = eps_single[((long long)r * B + b) * T + t];
        float4 e = eps[i];
        // (sum of the pair terms) - (coef * unconditional): two rounded operations, as the reference's tensor expression
        e.x = __fsub_rn(e.x, __fmul_rn(coef, u.x)); e.y = __fsub_rn(e.y, __fmul_rn(coef, u.y));
        e.z = __fsub_rn(e.z, __fmul_rn(coef, u.z)); e.w = __fsub_rn(e.w, __fmul_rn(coef, u.w));
        eps[i] = e;
    }
}

int launch_body_gather(const float* x, float* slices, int B, int n, int H, int T, cudaStream_t st) {
    const long long total = (long long)n * B * H;
    if (total == 0) return 0;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    body_gather_kernel<<<blocks, 256, 0, st>>>((const float4*)x, (float4*)slices, B, n, H, T);
    CINDM_CHECK_LAUNCH();
    return 0;
}

int launch_ebm_subtract(float* eps, const float* eps_single, int B, int n, int T, float coef, cudaStream_t st) {
    const long long total = (long long)B * T * n;
    if (total == 0) return 0;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    CINDM_CHECK_CUDA(launch_chain(ebm_subtract_kernel, dim3(blocks), dim3(256), 0, st, (float4*)eps, (const float4*)eps_single, B, n, T, coef));
    CINDM_CHECK_LAUNCH();
    return 0;
}

}  // namespace cindm
