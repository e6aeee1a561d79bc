// Engine state: weights (fp32 masters + repacked operands), time-bias tables, workspace.
#pragma once
#include <map>
#include <string>
#include <vector>

#include "common.cuh"
#include "../../include/cindm_b200.h"

namespace cindm {

struct HostTensor {
    std::vector<int64_t> shape;
    std::vector<float> data;
};

// One convolution's parameters.  `w` is the fp32 operand of the SIMT kernels, laid out
// [tap][cin][cout]; `w16[p]` (p = PREC_F16 / PREC_BF16) is the K-major tensor-core operand
// [tap][cout][cin] consumed through a TMA descriptor.
struct ConvW {
    int cin = 0, cout = 0, taps = 0;
    float* w = nullptr;
    float* bias = nullptr;        // [cout] or null
    void* w16[3] = {nullptr, nullptr, nullptr};
    void* w16t[3] = {nullptr, nullptr, nullptr};   // H=3 block-Toeplitz operand [(g,p,c)][(q,ci)], see conv_tc.h
};

struct NormW {
    float* gamma = nullptr;
    float* beta = nullptr;
};

struct ResBlockW {               // ResidualTemporalBlock, reference model/diffusion_1d.py:483-511
    ConvW conv0, conv1, res;
    NormW gn0, gn1;
    bool has_res = false;
    float* time_bias = nullptr;  // [timesteps][cout]: Linear(Mish(time_mlp(t))) precomputed per t
    std::string name;
};

struct AttnW {                   // Residual(PreNorm(LayerNorm, LinearAttentionTemporal)) :272-291
    float* g = nullptr;          // [C]
    ConvW qkv, out;
    // tcgen05 engine: to_qkv with the LayerNorm gain folded in, K-major [384][C] per 16-bit precision, and the row sums
    // of that (rounded) operand: qkv = rstd * (W' x - mean * wsum)   (attn_tc.cu)
    void* wln16[3] = {nullptr, nullptr, nullptr};
    float* wsum[3] = {nullptr, nullptr, nullptr};
    std::string name;
};

struct Workspace {
    int64_t max_slices = 0;
    int precision = -1;
    void* act[3] = {nullptr, nullptr, nullptr};
    void* skip[3] = {nullptr, nullptr, nullptr};
    void* res = nullptr;
    void* ln = nullptr;
    void* qkv = nullptr;
    void* att = nullptr;
    float* scratch = nullptr;    // fp32 pre-norm conv output (SIMT path)
    float* slices = nullptr;     // [S][24][8] gathered model inputs
    float* eps_pair = nullptr;   // [S][24][8] per-slice model outputs
    void* base = nullptr;        // single allocation everything above points into
    size_t bytes = 0;
};

struct SampleBuffers {           // sized by cindm_sample on first use
    cudaStream_t capture_stream = nullptr;   // used when the caller's stream is the (uncapturable) legacy stream
    cudaEvent_t ev_in = nullptr, ev_out = nullptr;
    size_t elems = 0;
    // cached CUDA graph of one (or two) DDPM steps
    std::string graph_key;
    cudaGraphExec_t graph_exec = nullptr;
    long long graph_nodes = 0;
    float* x_alt = nullptr;
    float* eps = nullptr;
    float* x0c = nullptr;          // composed x_start of the "mean" (outside) composition
    int* t_dev = nullptr;
    // DDIM: per-step table {time} / {sqrt(alpha_next), c, sigma, last-step flag} and the device-resident step index
    int* ddim_times = nullptr;
    float* ddim_coef = nullptr;
    int* step_dev = nullptr;
    int ddim_capacity = 0;
};

struct Tap {
    void* ptr;
    int c, h;
    int precision;               // element type of ptr
    int64_t s;
};

}  // namespace cindm

struct cindm_engine {
    cindm_config cfg;
    std::map<std::string, cindm::HostTensor> host_weights;
    bool finalized = false;

    // device weights
    std::vector<void*> allocations;
    cindm::ResBlockW downs_rb[4][2], ups_rb[3][2], mid_rb[2];
    cindm::AttnW downs_at[4], ups_at[3], mid_at;
    cindm::ConvW down_conv[3], up_conv[3];
    cindm::ConvW final_block;
    cindm::NormW final_gn;
    cindm::ConvW final_out;
    float* temb_table = nullptr;           // [timesteps][dim]

    // schedule (device, 13 x timesteps) and host copy
    float* sched_dev = nullptr;
    std::vector<float> sched_host;

    cindm::Workspace ws;
    cindm::SampleBuffers sb;

    // EBM body composition: the unconditional single-body engine (transition_dim 4) attached with
    // cindm_attach_unconditioned; not owned
    cindm_engine* uncond = nullptr;
    // initial_state_overwrite (cindm_set_initial_state_overwrite): [B][rows][4n] device tensor owned by the caller
    const float* overwrite = nullptr;
    int overwrite_rows = 0;

    // small batches: the 1x1 residual conv of a ResidualTemporalBlock runs on a side stream next to the block's first conv
    // (fork / join with events; the dependency is captured into the step's CUDA graph like any other edge)
    cudaStream_t side_stream = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    int fork_residual = -1;                // -1: by slice count; 0 / 1: CINDM_FORK_RES

    bool use_toeplitz = true;              // H=3 convs as one dense block-Toeplitz GEMM (tcgen05 engine)
    bool use_fused_attn = true;            // LayerNorm + to_qkv + attention core as one kernel (tcgen05 engine)
    bool taps_enabled = false;
    std::map<std::string, cindm::Tap> taps;
    std::vector<void*> tap_allocs;
};

namespace cindm {

// schedule table indices (order of cindm_schedule_tables)
enum {
    TAB_BETAS = 0, TAB_ACP, TAB_ACP_PREV, TAB_SQRT_ACP, TAB_SQRT_1M_ACP, TAB_LOG_1M_ACP, TAB_SQRT_RECIP_ACP,
    TAB_SQRT_RECIPM1_ACP, TAB_POST_VAR, TAB_POST_LOGVAR, TAB_POST_C1, TAB_POST_C2, TAB_LOSS_W, TAB_COUNT
};

// ---- launchers (each returns 0 or a negative error code) -------------------------------------

struct ConvLaunch {
    const void* in0 = nullptr; int c0 = 0;      // [S][Hin][c0]
    const void* in1 = nullptr; int c1 = 0;      // optional channel-concatenated second input
    const ConvW* w = nullptr;
    const void* res = nullptr;                  // optional residual [S][Hout][cout], same type as out
    void* out = nullptr;
    int64_t S = 0;
    int Hin = 0, Hout = 0, stride = 1, pad = 0, transposed = 0;
    int in_prec = PREC_F32, out_prec = PREC_F32;
};
int launch_conv_simt(const ConvLaunch& a, cudaStream_t st);

// GroupNorm(8) + Mish over fp32 pre-norm activations, then + per-channel vector or + residual tensor
// add_vec: per-channel vector; when t_dev != nullptr the row (*t_dev) of a [timesteps][C] table at add_vec
int launch_gn_mish(const float* in, const NormW& gn, const float* add_vec, const int* t_dev, const void* add_res,
                   void* out, int64_t S, int H, int C, int out_prec, cudaStream_t st);
int launch_layernorm(const void* in, const float* g, void* out, int64_t rows, int C, int prec, cudaStream_t st);
int launch_attn_core(const void* qkv, void* out, int64_t S, int n, int prec, cudaStream_t st);
// LayerNorm + to_qkv + attention core in one tcgen05 kernel: x [S][H][C] -> out [S][H][128] (attn_tc.cu)
int launch_qkv_attn_tc(const AttnW& a, const void* x, void* out, int64_t S, int H, int C, int prec, cudaStream_t st, int reverse = 0);

// fused first block (16-bit paths): see kernels_fused.cu
struct StemLaunch {
    const float* x = nullptr;                     // slices [S][24][8], or the design tensor when gather != 0
    int gather = 0, B = 0, n = 0, start = 0, T = 0;
    const ConvW* conv0 = nullptr; const NormW* gn = nullptr; const ConvW* res = nullptr;
    const float* tbias = nullptr; const int* t_dev = nullptr;
    void* out_b0 = nullptr; void* out_res = nullptr;
    int64_t S = 0; int prec = PREC_F16;
};
int launch_stem(const StemLaunch& a, cudaStream_t st);
int launch_head(const void* in, const ConvW& w, float* out, int64_t rows, int prec, cudaStream_t st);

// when non-null, the U-Net reads its slices straight out of the design tensor x[B][T][4n]
struct GatherSpec { const float* x; int B, n, nc, start; };   // (an engine with transition_dim 4 gathers single bodies: slice = body * B + b)
int launch_body_gather(const float* x, float* slices, int B, int n, int H, int T, cudaStream_t st);
int launch_ebm_subtract(float* eps, const float* eps_single, int B, int n, int T, float coef, cudaStream_t st);

// t_dev == nullptr: use the host value t; else the device integer *t_dev (CUDA-graph replay)
int unet_forward(cindm_engine* e, const float* slices, int64_t S, int t, const int* t_dev, float* eps_pair,
                 int precision, int conv_engine, cudaStream_t st, const GatherSpec* gather = nullptr);

int launch_compose_gather(const float* x, float* slices, int B, int n, int nc, int start, int H, cudaStream_t st);
// "mean" composition of p_sample_compose_outside: per-slice clamped x_start and posterior mean, averaged (compose.cu)
int launch_compose_scatter_posterior(const float* eps_pair, const float* x, float* mean_out, float* x0_out, int B, int n,
                                     int nc, int start, int H, const float* sched, int timesteps, int t, const int* t_dev,
                                     cudaStream_t st);
int launch_compose_scatter(const float* eps_pair, float* eps, int B, int n, int nc, int start, int H, int mode,
                           cudaStream_t st);
int launch_design_grad(const float* x, float* g, int B, int T, int n, const cindm_objective& obj, cudaStream_t st);

struct UpdateLaunch {
    const float* x = nullptr; const float* eps = nullptr;
    float* x_out = nullptr; float* pred_out = nullptr; float* x0_out = nullptr;
    int B = 0, T = 0, n = 0;
    const float* sched = nullptr; int timesteps = 0;   // device tables [13][timesteps]
    const int* t_dev = nullptr; int t_host = 0;        // *t_dev (if non-null) overrides t_host
    int renoise = 0;                                   // 1: recurrence re-noise, 0: final posterior noise
    // noise source: explicit tensor (noise + ((t_start - t)*draws_per_step + draw) * B*T*4n), or Philox
    const float* noise = nullptr; int t_start = 0, draws_per_step = 0, draw = 0;
    int use_philox = 0; uint64_t seed = 0; int64_t cand_off = 0;
    cindm_objective obj;
    // DDIM outer update (reference ddim_sample :1781-1797) instead of the posterior sample: out = x0 * coef[0] +
    // coef[1] * (eps + g) + coef[2] * noise, or x0 on the last step (coef[3] != 0); row = *step_dev of ddim_coef,
    // which also replaces (t_start - t) as the row of the explicit noise tensor
    int ddim = 0; const float* ddim_coef = nullptr; const int* step_dev = nullptr;
    // "mean" (outside) composition: the posterior mean and x_start are inputs (eps is not read)
    const float* mean_in = nullptr; const float* x0_in = nullptr;
    // conditioned models: frames [0, cond_rows) of every candidate are the condition: copied through unchanged, and an
    // explicit noise tensor only covers the T - cond_rows sampled frames
    int cond_rows = 0;
    // initial_state_overwrite (:1273-1276, :1355-1362): pred = overwrite[b][t] for t < ow_rows, before the noise is added
    const float* overwrite = nullptr; int ow_rows = 0;
};
int launch_update(const UpdateLaunch& u, cudaStream_t st);
int launch_fill_noise(float* x, int B, int T, int n, uint64_t seed, int64_t cand_off, int t, int draw,
                      cudaStream_t st);
int launch_step_counter(int* t_dev, int delta, cudaStream_t st);
int launch_ula_step(const float* x, const float* eps, const float* noise, float* out, int B, int T, int n, float grad_scale,
                    float ss, uint64_t seed, int64_t cand_off, int t, int draw, cudaStream_t st);
int launch_predict_start(const float* x, const float* eps, float* x0, long long elems, const float* sched, int timesteps, int t,
                         int clip, cudaStream_t st);
// composing_time_sample (:1827-1829): block k+1's first cond_rows frames <- block k's last cond_rows frames
int launch_chain_condition(float* x, int rows_per_block, int blocks, int T, int n, int cond_rows, cudaStream_t st);
int sample_ddim(cindm_engine* e, const cindm_sample_config& c, int n_pairs, const int32_t* times, const int32_t* times_next,
                const float* coef3, float* x, const float* noise, float* x0_out, cudaStream_t caller);

void graph_cache_clear(cindm_engine* e);
int finalize_weights(cindm_engine* e, cudaStream_t st);
int reserve_workspace(cindm_engine* e, int64_t S, int prec);
int64_t workspace_bytes(int64_t S, int prec, int horizon, int dim);
int down_samplings(int horizon);                 // Downsample1d stages of a model with this horizon (reference :549-554)
// x0_composed: required for (and only used by) CINDM_COMPOSE_MEAN_OUTSIDE, where `eps` receives the composed posterior mean
// mode CINDM_COMPOSE_EBM: sum over pairs minus ebm_coef * the attached unconditional single-body model (nc must be 0)
int composed_eps(cindm_engine* e, const float* x, float* eps, int B, int n, int nc, int start, int mode, int t,
                 const int* t_dev, int prec, int conv_engine, cudaStream_t st, float* x0_composed = nullptr, float ebm_coef = 0.f);
int sample_loop(cindm_engine* e, const cindm_sample_config& cfg, float* x, const float* noise, float* x0_out,
                cudaStream_t st);

size_t elem_size(int prec);

}  // namespace cindm
