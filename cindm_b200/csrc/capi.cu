// extern "C" boundary of libcindm_b200.so (declared in include/cindm_b200.h).
#include <cstdlib>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <tuple>

#include "engine.h"

namespace cindm {

static thread_local std::string g_last_error;
void set_error(const std::string& msg) { g_last_error = msg; }
int fail(int code, const std::string& msg) {
    g_last_error = msg;
    return code;
}

// ---- tracing state
static long long g_launches = 0;
static bool g_profiling = false;
struct ProfEntry { std::string tag; cudaEvent_t e0, e1; double work; };
static std::vector<ProfEntry> g_prof;
static std::vector<size_t> g_prof_open;
void count_launch() { ++g_launches; }
void add_launches(long long n) { g_launches += n; }
long long launch_count() { return g_launches; }
bool profiling_enabled() { return g_profiling; }
bool pdl_enabled() {
    static int mode = -1;
    if (mode < 0) { const char* e = getenv("CINDM_PDL"); mode = (e && e[0] == '0') ? 0 : 1; }
    return mode == 1;
}
void profile_record(const char* tag, cudaStream_t st, bool begin, double work) {
    if (begin) {
        ProfEntry pe; pe.tag = tag; pe.work = work;
        cudaEventCreate(&pe.e0); cudaEventCreate(&pe.e1);
        cudaEventRecord(pe.e0, st);
        g_prof.push_back(pe);
        g_prof_open.push_back(g_prof.size() - 1);
    } else if (!g_prof_open.empty()) {
        cudaEventRecord(g_prof[g_prof_open.back()].e1, st);
        g_prof_open.pop_back();
    }
}

int nbody_rollout(const double* state0, double* traj, int B, int n, int n_steps, int stride, cudaStream_t st);
int score_designs(const float* pred, double* mae, double* objective, int B, int T, int n, double tx, double ty,
                  cudaStream_t st);

}  // namespace cindm

using namespace cindm;

namespace {
template <typename T>
__global__ void tap_to_f32_cf(const T* __restrict__ in, float* __restrict__ out, long long S, int C, int H) {
    long long total = S * C * H;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int h = (int)(i % H);
        long long r = i / H;
        int c = (int)(r % C);
        long long s = r / C;
        out[i] = to_f32<T>(in[(s * H + h) * C + c]);       // channels-last -> channels-first
    }
}
}  // namespace

#define API_BEGIN try {
#define API_END                                                             \
    }                                                                       \
    catch (const std::exception& ex) { return fail(-1, std::string("exception: ") + ex.what()); } \
    catch (...) { return fail(-1, "unknown exception"); }

extern "C" {

const char* cindm_last_error(void) { return g_last_error.c_str(); }
int cindm_version(void) { return 200; }

long long cindm_launch_count(void) { return launch_count(); }

int cindm_profile_enable(int enable) {
    API_BEGIN
    cudaDeviceSynchronize();
    for (auto& pe : g_prof) { cudaEventDestroy(pe.e0); cudaEventDestroy(pe.e1); }
    g_prof.clear(); g_prof_open.clear();
    g_profiling = enable != 0;
    return 0;
    API_END
}

// Writes "tag,launch_groups,total_ms,work" lines (one per kernel class) into buf; returns bytes needed.
int cindm_profile_report(char* buf, int capacity) {
    API_BEGIN
    cudaDeviceSynchronize();
    std::map<std::string, std::tuple<int, double, double>> acc;
    std::vector<std::string> order;
    for (auto& pe : g_prof) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, pe.e0, pe.e1) != cudaSuccess) continue;
        if (!acc.count(pe.tag)) order.push_back(pe.tag);
        auto& a = acc[pe.tag];
        std::get<0>(a) += 1; std::get<1>(a) += ms; std::get<2>(a) += pe.work;
    }
    std::string out;
    for (auto& tag : order) {
        auto& a = acc[tag];
        char line[256];
        snprintf(line, sizeof line, "%s,%d,%.6f,%.6e\n", tag.c_str(), std::get<0>(a), std::get<1>(a), std::get<2>(a));
        out += line;
    }
    if (buf && capacity > 0) {
        int n = (int)out.size() < capacity - 1 ? (int)out.size() : capacity - 1;
        memcpy(buf, out.data(), n);
        buf[n] = 0;
    }
    return (int)out.size() + 1;
    API_END
}

int cindm_create(const cindm_config* cfg, cindm_engine** out) {
    API_BEGIN
    if (!cfg || !out) return fail(-2, "null argument");
    // horizon 24 / dim 64 is the model every 16-bit tensor-core kernel is built for; the reference's other model names
    // (44-step rollout, Unet_dim 96: inference/inverse_design_diffusion_1d.py:150-154) run on the generic fp32 kernels
    if (cfg->dim < 16 || cfg->dim > 128 || cfg->dim % 16 != 0)
        return fail(-2, "Unet_dim must be a multiple of 16 in [16, 128] (dim_mults (1,2,4,8); 64 and 96 are the reference's models)");
    if (cfg->horizon < 8 || cfg->horizon > 48 || cfg->horizon % 2 != 0)
        return fail(-2, "horizon must be even and in [8, 48] (24: the 24-step models, 44: the 44-step models)");
    if (cfg->transition_dim != 8 && cfg->transition_dim != 4)
        return fail(-2, "transition_dim must be 8 (two bodies x 4 features) or 4 (the unconditional single-body model)");
    if (cfg->timesteps < 1 || cfg->timesteps > 65535) return fail(-2, "timesteps out of range");
    cindm_engine* e = new cindm_engine();
    e->cfg = *cfg;
    // A/B switches for measurements (both default on)
    if (const char* v = getenv("CINDM_TOEPLITZ")) e->use_toeplitz = v[0] != '0';
    if (const char* v = getenv("CINDM_FUSED_ATTN")) e->use_fused_attn = v[0] != '0';
    if (const char* v = getenv("CINDM_FORK_RES")) e->fork_residual = v[0] != '0';
    *out = e;
    return 0;
    API_END
}

int cindm_destroy(cindm_engine* e) {
    API_BEGIN
    if (!e) return 0;
    cudaDeviceSynchronize();
    for (void* p : e->allocations) cudaFree(p);
    for (void* p : e->tap_allocs) cudaFree(p);
    if (e->ws.base) cudaFree(e->ws.base);
    if (e->sched_dev) cudaFree(e->sched_dev);
    if (e->sb.x_alt) cudaFree(e->sb.x_alt);
    if (e->sb.eps) cudaFree(e->sb.eps);
    if (e->sb.x0c) cudaFree(e->sb.x0c);
    if (e->sb.t_dev) cudaFree(e->sb.t_dev);
    if (e->sb.step_dev) cudaFree(e->sb.step_dev);
    if (e->sb.ddim_times) cudaFree(e->sb.ddim_times);
    if (e->sb.ddim_coef) cudaFree(e->sb.ddim_coef);
    graph_cache_clear(e);
    if (e->side_stream) cudaStreamDestroy(e->side_stream);
    if (e->ev_fork) cudaEventDestroy(e->ev_fork);
    if (e->ev_join) cudaEventDestroy(e->ev_join);
    if (e->sb.capture_stream) cudaStreamDestroy(e->sb.capture_stream);
    if (e->sb.ev_in) cudaEventDestroy(e->sb.ev_in);
    if (e->sb.ev_out) cudaEventDestroy(e->sb.ev_out);
    delete e;
    return 0;
    API_END
}

int cindm_load_weight(cindm_engine* e, const char* name, const float* host, const int64_t* shape, int ndim) {
    API_BEGIN
    if (!e || !name || !host || !shape) return fail(-2, "null argument");
    if (e->finalized) return fail(-4, "weights already finalized");
    HostTensor t;
    int64_t n = 1;
    for (int i = 0; i < ndim; ++i) { t.shape.push_back(shape[i]); n *= shape[i]; }
    t.data.assign(host, host + n);
    e->host_weights[name] = std::move(t);
    return 0;
    API_END
}

int cindm_finalize_weights(cindm_engine* e, void* stream) {
    API_BEGIN
    if (!e) return fail(-2, "null engine");
    return finalize_weights(e, (cudaStream_t)stream);
    API_END
}

int cindm_set_schedule(cindm_engine* e, const float* tables13, int timesteps) {
    API_BEGIN
    if (!e || !tables13) return fail(-2, "null argument");
    if (timesteps != e->cfg.timesteps) return fail(-2, "schedule length does not match the engine's timesteps");
    size_t n = (size_t)TAB_COUNT * timesteps;
    e->sched_host.assign(tables13, tables13 + n);
    if (!e->sched_dev) CINDM_CHECK_CUDA(cudaMalloc(&e->sched_dev, n * sizeof(float)));
    CINDM_CHECK_CUDA(cudaMemcpy(e->sched_dev, tables13, n * sizeof(float), cudaMemcpyHostToDevice));
    return 0;
    API_END
}

int cindm_reserve(cindm_engine* e, int64_t max_slices, int precision) {
    API_BEGIN
    if (!e) return fail(-2, "null engine");
    return reserve_workspace(e, max_slices, precision);
    API_END
}

int64_t cindm_workspace_bytes(int64_t max_slices, int precision) { return workspace_bytes(max_slices, precision, 24, 64); }
int64_t cindm_model_workspace_bytes(int horizon, int dim, int64_t max_slices, int precision) {
    return workspace_bytes(max_slices, precision, horizon, dim);
}

int cindm_schedule_tables(int timesteps, float* out) {
    API_BEGIN
    if (!out || timesteps < 1) return fail(-2, "bad argument");
    const int T = timesteps;
    const double s = 0.008;
    std::vector<double> ac(T + 1), betas(T), acp(T), acp_prev(T);
    for (int i = 0; i <= T; ++i) {
        double x = (double)i;     // torch.linspace(0, T, T+1) is exact for integers
        double c = cos(((x / T) + s) / (1 + s) * M_PI * 0.5);
        ac[i] = c * c;
    }
    double ac0 = ac[0];
    for (int i = 0; i <= T; ++i) ac[i] /= ac0;
    double run = 1.0;
    for (int i = 0; i < T; ++i) {
        double b = 1.0 - ac[i + 1] / ac[i];
        b = b < 0.0 ? 0.0 : (b > 0.999 ? 0.999 : b);
        betas[i] = b;
        acp_prev[i] = run;
        run *= (1.0 - b);
        acp[i] = run;
    }
    auto put = [&](int tab, int i, double v) { out[(size_t)tab * T + i] = (float)v; };
    for (int i = 0; i < T; ++i) {
        double b = betas[i], a = acp[i], ap = acp_prev[i];
        double pv = b * (1.0 - ap) / (1.0 - a);
        put(TAB_BETAS, i, b);
        put(TAB_ACP, i, a);
        put(TAB_ACP_PREV, i, ap);
        put(TAB_SQRT_ACP, i, sqrt(a));
        put(TAB_SQRT_1M_ACP, i, sqrt(1.0 - a));
        put(TAB_LOG_1M_ACP, i, log(1.0 - a));
        put(TAB_SQRT_RECIP_ACP, i, sqrt(1.0 / a));
        put(TAB_SQRT_RECIPM1_ACP, i, sqrt(1.0 / a - 1.0));
        put(TAB_POST_VAR, i, pv);
        put(TAB_POST_LOGVAR, i, log(pv < 1e-20 ? 1e-20 : pv));
        put(TAB_POST_C1, i, b * sqrt(ap) / (1.0 - a));
        put(TAB_POST_C2, i, (1.0 - ap) * sqrt(1.0 - b) / (1.0 - a));
        put(TAB_LOSS_W, i, 1.0);
    }
    return 0;
    API_END
}

int cindm_build_index_maps(int n, int nc, int start, int H, int32_t* win_t0, int32_t* pair_i, int32_t* pair_j,
                           int32_t* cover) {
    API_BEGIN
    if (n < 2 || nc < 0 || start <= 0 || H <= 0) return fail(-2, "bad argument");
    const int T = H + nc * start;
    if (win_t0)
        for (int kk = 0; kk <= nc; ++kk) win_t0[kk] = kk * start;
    int p = 0;
    for (int ii = 0; ii < n; ++ii)
        for (int jj = ii + 1; jj < n; ++jj, ++p) {
            if (pair_i) pair_i[p] = ii;
            if (pair_j) pair_j[p] = jj;
        }
    if (cover)
        for (int t = 0; t < T; ++t) {
            int c = 0;
            for (int kk = 0; kk <= nc; ++kk) c += (t >= kk * start && t < kk * start + H) ? 1 : 0;
            cover[t] = c;
        }
    return 0;
    API_END
}

int cindm_compose_gather(const float* x, float* slices, int B, int n, int nc, int start, int H, void* stream) {
    API_BEGIN
    if (n < 2 || nc < 0 || start <= 0) return fail(-2, "bad composition parameters");
    return launch_compose_gather(x, slices, B, n, nc, start, H, (cudaStream_t)stream);
    API_END
}

int cindm_compose_scatter_mean(const float* eps_pair, float* eps, int B, int n, int nc, int start, int H, int mode,
                               void* stream) {
    API_BEGIN
    if (n < 2 || nc < 0 || start <= 0) return fail(-2, "bad composition parameters");
    if (mode != CINDM_COMPOSE_MEAN_INSIDE && mode != CINDM_COMPOSE_SUM_INSIDE) return fail(-2, "bad compose mode");
    return launch_compose_scatter(eps_pair, eps, B, n, nc, start, H, mode, (cudaStream_t)stream);
    API_END
}

int cindm_unet_forward(cindm_engine* e, const float* slices, int64_t S, int t, float* eps_pair, int precision,
                       int conv_engine, void* stream) {
    API_BEGIN
    if (!e) return fail(-2, "null engine");
    CINDM_TRY(reserve_workspace(e, S > e->ws.max_slices ? S : e->ws.max_slices, precision));
    return unet_forward(e, slices, S, t, nullptr, eps_pair, precision, conv_engine, (cudaStream_t)stream);
    API_END
}

int cindm_unet_enable_taps(cindm_engine* e, int enable) {
    if (!e) return fail(-2, "null engine");
    e->taps_enabled = enable != 0;
    return 0;
}


int cindm_unet_read_tap(cindm_engine* e, const char* name, float* host, int64_t capacity, int64_t* s, int64_t* c,
                        int64_t* h) {
    API_BEGIN
    if (!e || !name) return fail(-2, "null argument");
    auto it = e->taps.find(name);
    if (it == e->taps.end()) return fail(-7, std::string("no such tap: ") + name);
    const Tap& tp = it->second;
    if (s) *s = tp.s;
    if (c) *c = tp.c;
    if (h) *h = tp.h;
    int64_t n = tp.s * tp.c * tp.h;
    if (!host) return 0;
    if (capacity < n) return fail(-2, "tap buffer too small");
    float* tmp = nullptr;
    CINDM_CHECK_CUDA(cudaMalloc(&tmp, n * sizeof(float)));
    int blocks = (int)((n + 255) / 256);
    if (tp.precision == PREC_F32) tap_to_f32_cf<float><<<blocks, 256>>>((const float*)tp.ptr, tmp, tp.s, tp.c, tp.h);
    else if (tp.precision == PREC_F16) tap_to_f32_cf<__half><<<blocks, 256>>>((const __half*)tp.ptr, tmp, tp.s, tp.c, tp.h);
    else tap_to_f32_cf<__nv_bfloat16><<<blocks, 256>>>((const __nv_bfloat16*)tp.ptr, tmp, tp.s, tp.c, tp.h);
    cudaError_t ce = cudaMemcpy(host, tmp, n * sizeof(float), cudaMemcpyDeviceToHost);
    cudaFree(tmp);
    if (ce != cudaSuccess) return fail(-100, cudaGetErrorString(ce));
    return 0;
    API_END
}

int cindm_composed_eps(cindm_engine* e, const float* x, float* eps, int B, int n, int nc, int start, int mode, int t,
                       int precision, int conv_engine, void* stream) {
    API_BEGIN
    if (!e) return fail(-2, "null engine");
    if (n < 2 || nc < 0 || start <= 0) return fail(-2, "bad composition parameters");
    const int64_t S = (int64_t)(nc + 1) * (n * (n - 1) / 2) * B;
    CINDM_TRY(reserve_workspace(e, S > e->ws.max_slices ? S : e->ws.max_slices, precision));
    return composed_eps(e, x, eps, B, n, nc, start, mode, t, nullptr, precision, conv_engine, (cudaStream_t)stream);
    API_END
}

int cindm_design_grad(const float* x, float* g, int B, int T, int n, const cindm_objective* obj, void* stream) {
    API_BEGIN
    if (!obj) return fail(-2, "null objective");
    if (T < 2) return fail(-2, "need at least two time steps");
    return launch_design_grad(x, g, B, T, n, *obj, (cudaStream_t)stream);
    API_END
}

int cindm_composed_posterior(cindm_engine* e, const float* x, float* mean_out, float* x0_out, int B, int n, int nc,
                             int start, int t, int precision, int conv_engine, void* stream) {
    API_BEGIN
    if (!e || !x || !mean_out || !x0_out) return fail(-2, "null argument");
    if (!e->finalized) return fail(-4, "weights not finalized");
    if (t < 0 || t >= e->cfg.timesteps) return fail(-2, "timestep out of range");
    const int64_t S = (int64_t)(nc + 1) * (n * (n - 1) / 2) * B;
    CINDM_TRY(reserve_workspace(e, S > e->ws.max_slices ? S : e->ws.max_slices, precision));
    return composed_eps(e, x, mean_out, B, n, nc, start, CINDM_COMPOSE_MEAN_OUTSIDE, t, nullptr, precision, conv_engine,
                        (cudaStream_t)stream, x0_out);
    API_END
}

int cindm_posterior_update(cindm_engine* e, const float* x, const float* eps, const float* noise, float* x_out,
                           float* pred_out, float* x0_out, int B, int T, int n, int t, int renoise,
                           const cindm_objective* obj, void* stream) {
    API_BEGIN
    if (!e) return fail(-2, "null engine");
    if (t < 0 || t >= e->cfg.timesteps) return fail(-2, "timestep out of range");
    UpdateLaunch u;
    u.x = x; u.eps = eps; u.x_out = x_out; u.pred_out = pred_out; u.x0_out = x0_out;
    u.B = B; u.T = T; u.n = n; u.sched = e->sched_dev; u.timesteps = e->cfg.timesteps;
    u.t_host = t; u.renoise = renoise; u.noise = noise; u.t_start = t; u.draws_per_step = 1; u.draw = 0;
    if (obj) u.obj = *obj;
    else { memset(&u.obj, 0, sizeof(u.obj)); u.obj.guidance = CINDM_GUIDE_NONE; }
    return launch_update(u, (cudaStream_t)stream);
    API_END
}

int cindm_attach_unconditioned(cindm_engine* e, cindm_engine* single) {
    API_BEGIN
    if (!e) return fail(-2, "null engine");
    if (e->cfg.transition_dim != 8) return fail(-2, "the unconditional model attaches to a body-pair engine (transition_dim 8)");
    if (single && (single->cfg.transition_dim != 4 || single->cfg.horizon != e->cfg.horizon || single->cfg.timesteps != e->cfg.timesteps))
        return fail(-2, "the unconditional engine must have transition_dim 4 and the pair engine's horizon / timesteps");
    if (e->uncond != single) { cudaDeviceSynchronize(); graph_cache_clear(e); }
    e->uncond = single;
    return 0;
    API_END
}

int cindm_ebm_eps(cindm_engine* e, const float* x, float* eps, int B, int n, float uncond_coef, int t, int precision,
                  int conv_engine, void* stream) {
    API_BEGIN
    if (!e || !x || !eps) return fail(-2, "null argument");
    if (!e->uncond) return fail(-4, "no unconditional single-body engine attached (cindm_attach_unconditioned)");
    if (t < 0 || t >= e->cfg.timesteps) return fail(-2, "timestep out of range");
    if (n < 3) return fail(-2, "the EBM body composition needs at least 3 bodies");
    const int64_t S = (int64_t)(n * (n - 1) / 2) * B, S1 = (int64_t)n * B;
    CINDM_TRY(reserve_workspace(e, S > e->ws.max_slices ? S : e->ws.max_slices, precision));
    CINDM_TRY(reserve_workspace(e->uncond, S1 > e->uncond->ws.max_slices ? S1 : e->uncond->ws.max_slices, precision));
    return composed_eps(e, x, eps, B, n, 0, 1, CINDM_COMPOSE_EBM, t, nullptr, precision, conv_engine, (cudaStream_t)stream, nullptr,
                        uncond_coef);
    API_END
}

int cindm_ula_step(const float* x, const float* eps, const float* noise, float* x_out, int B, int T, int n, float grad_scale,
                   float step_size, uint64_t seed, int64_t cand_off, int t, int draw, void* stream) {
    API_BEGIN
    if (!x || !eps || !x_out) return fail(-2, "null argument");
    return launch_ula_step(x, eps, noise, x_out, B, T, n, grad_scale, step_size, seed, cand_off, t, draw, (cudaStream_t)stream);
    API_END
}

int cindm_predict_start(cindm_engine* e, const float* x, const float* eps, float* x0_out, int64_t elems, int t, int clip,
                        void* stream) {
    API_BEGIN
    if (!e || !x || !eps || !x0_out) return fail(-2, "null argument");
    if (t < 0 || t >= e->cfg.timesteps) return fail(-2, "timestep out of range");
    return launch_predict_start(x, eps, x0_out, elems, e->sched_dev, e->cfg.timesteps, t, clip, (cudaStream_t)stream);
    API_END
}

int cindm_sample(cindm_engine* e, const cindm_sample_config* cfg, float* x, const float* noise, float* x0_out,
                 void* stream) {
    API_BEGIN
    if (!e || !cfg || !x) return fail(-2, "null argument");
    return sample_loop(e, *cfg, x, noise, x0_out, (cudaStream_t)stream);
    API_END
}

int cindm_set_initial_state_overwrite(cindm_engine* e, const float* ow, int rows) {
    API_BEGIN
    if (!e) return fail(-2, "null engine");
    if ((ow == nullptr) != (rows <= 0)) return fail(-2, "initial_state_overwrite: pointer and frame count come together");
    e->overwrite = ow;
    e->overwrite_rows = ow ? rows : 0;
    return 0;
    API_END
}

int cindm_sample_ddim(cindm_engine* e, const cindm_sample_config* cfg, int n_pairs, const int32_t* times,
                      const int32_t* times_next, const float* coef, float* x, const float* noise, float* x0_out,
                      void* stream) {
    API_BEGIN
    if (!e || !cfg || !x) return fail(-2, "null argument");
    return sample_ddim(e, *cfg, n_pairs, times, times_next, coef, x, noise, x0_out, (cudaStream_t)stream);
    API_END
}

int cindm_fill_initial_noise(float* x, int B, int T, int n, uint64_t seed, int64_t cand_off, int timesteps,
                             void* stream) {
    API_BEGIN
    return launch_fill_noise(x, B, T, n, seed, cand_off, timesteps, 0xFFFF, (cudaStream_t)stream);
    API_END
}

int cindm_nbody_rollout(const double* state0, double* traj, int B, int n, int n_steps, int stride, void* stream) {
    API_BEGIN
    return nbody_rollout(state0, traj, B, n, n_steps, stride, (cudaStream_t)stream);
    API_END
}

int cindm_score_designs(const float* pred, double* mae, double* objective, int B, int T, int n, double tx, double ty,
                        void* stream) {
    API_BEGIN
    return score_designs(pred, mae, objective, B, T, n, tx, ty, (cudaStream_t)stream);
    API_END
}

}  // extern "C"
