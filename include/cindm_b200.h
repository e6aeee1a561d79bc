/*
 * cindm_b200 — C ABI of the B200-native (sm_100a) compositional-sampling hot path of CinDM.
 *
 * The reference (AI4Science-WestlakeU/cindm) has no FFI layer: its "operator API" for this
 * path is the Python class surface of model/diffusion_1d.py.  Each entry point below names
 * the reference code it replaces (file:line under /root/reference).  The Python mirror of
 * that class surface lives in cindm_b200/model/diffusion_1d.py and calls these functions
 * through ctypes (see INTEGRATION.md for the binding a reference maintainer would add).
 *
 * Conventions
 *   - every function returns 0 on success, a negative code on failure; the message is
 *     available from cindm_last_error() (thread-local).  Nothing throws across the boundary.
 *   - all tensors are row-major, contiguous, fp32 at the boundary.  Pointers named *_dev
 *     are device pointers owned by the caller; the library owns only what cindm_create /
 *     cindm_reserve allocated.  Step functions never allocate (CUDA-graph capturable).
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *   - layouts: design x[B][T][4n] (features per body: x, y, vx, vy); model slices
 *     [S][24][8] with slice id s = (kk*P + pair)*B + b  (kk window, pair lexicographic ii<jj).
 */
#ifndef CINDM_B200_H
#define CINDM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct cindm_engine cindm_engine;

/* precision of the U-Net activations / GEMM operands (accumulation is always fp32) */
enum { CINDM_PREC_F32 = 0, CINDM_PREC_F16 = 1, CINDM_PREC_BF16 = 2 };
/* conv engine: SIMT FMA kernels, or tcgen05 tensor-core kernels (16-bit precisions only) */
enum { CINDM_CONV_SIMT = 0, CINDM_CONV_TCGEN05 = 1 };
/* composition reduce: "mean-inside" / "sum-inside" (model/diffusion_1d.py:994-999) */
/* compose_mode of sample() / p_sample_loop (model/diffusion_1d.py:1683-1715):
 *   "mean-inside" / "sum-inside": composed epsilon inside model_predictions (:959-1001), then one posterior;
 *   "mean" (the API default): p_sample_compose_outside (:1379-1652) - every (window, pair) slice gets its own clamped
 *       x_start and posterior mean, and THOSE are averaged over senders and covering windows (:1446-1451);
 *   "noise_sum": the summed epsilon of :1452-1461, which is the sum-inside operator followed by the same posterior.
 *   CINDM_COMPOSE_EBM: gradient() (:1856-1982, reached through model_predictions when model_unconditioned is set, :1003):
 *       sum over the pairs containing a body of the pair model's epsilon for it, minus ebm_uncond_coef x the epsilon of the
 *       unconditional single-body model (an engine with transition_dim 4 attached by cindm_attach_unconditioned). */
enum { CINDM_COMPOSE_MEAN_INSIDE = 0, CINDM_COMPOSE_SUM_INSIDE = 1, CINDM_COMPOSE_MEAN_OUTSIDE = 2, CINDM_COMPOSE_NOISE_SUM = 3,
       CINDM_COMPOSE_EBM = 4 };
/* design objective: get_design_fn design_fn_mode (inference/inverse_design_diffusion_1d.py:215-222) */
enum { CINDM_OBJ_L2 = 0, CINDM_OBJ_L2SQUARE = 1 };
/* guidance scaling: "standard*" (g) or "standard-alpha*" (beta_t/sqrt(abar_prev_t) * g) (:1321-1324) */
enum { CINDM_GUIDE_NONE = 0, CINDM_GUIDE_STANDARD = 1, CINDM_GUIDE_STANDARD_ALPHA = 2 };

typedef struct {
    int horizon;        /* 24: TemporalUnet1D(horizon=...)            model/diffusion_1d.py:521.  Even, 8..48: 24 is the model the
                           16-bit tensor-core kernels are built for; the 44-step models (inverse_design_diffusion_1d.py:150-154)
                           and any other even horizon run with CINDM_PREC_F32 on CINDM_CONV_SIMT */
    int transition_dim; /* 8 : two bodies x (x,y,vx,vy); 4: the unconditional single-body model  :522 */
    int dim;            /* 64: Unet_dim (:524); a multiple of 16 up to 128 (96: the "_Unet_dim-96" model), fp32 / simt when != 64 */
    int timesteps;      /* 1000: GaussianDiffusion1D(timesteps=...)                         :810 */
} cindm_config;

typedef struct {
    double target_x, target_y; /* pos_target (fp64 in the driver, inverse_design_diffusion_1d.py:281) */
    float coef;                /* design_coef */
    float consistency_coef;    /* time_consistency_coef */
    int mode;                  /* CINDM_OBJ_* */
    int guidance;              /* CINDM_GUIDE_* */
} cindm_objective;

typedef struct {
    int batch;              /* B candidates on this device */
    int n_bodies;           /* compose_n_bodies */
    int n_composed;         /* extra windows; W = n_composed + 1 */
    int compose_start_step; /* window stride */
    int compose_mode;       /* CINDM_COMPOSE_* */
    int recurrence;         /* 0: "standard" single pass; K>0: "...-recurrence-K" */
    int precision;          /* CINDM_PREC_* */
    int conv_engine;        /* CINDM_CONV_* */
    int t_start, t_end;     /* inclusive start (e.g. 999) down to inclusive end (e.g. 0) */
    uint64_t seed;          /* Philox key when noise_dev == NULL */
    int64_t candidate_offset; /* global id of local candidate 0 (sharding-invariant noise) */
    int use_graph;          /* capture one DDPM step in a CUDA graph and replay it */
    cindm_objective objective;
    /* conditioned models (GaussianDiffusion1D(conditioned_steps=k), model/diffusion_1d.py:956-957, :1028-1030): the first
     * cond_rows frames of every x[b] hold the condition; the model sees them (x = cat(cond, x)), the update leaves them
     * untouched and explicit noise tensors do not cover them ([.., T - cond_rows, 4n]).  0 = unconditioned. */
    int cond_rows;
    /* composing_time_sample (:1806-1854): the batch is chain_blocks blocks of batch / chain_blocks trajectories; before
     * every step block k+1 takes the last cond_rows frames of block k as its condition (:1827-1829).  0 / 1 = off. */
    int chain_blocks;
    /* CINDM_COMPOSE_EBM only: coefficient of the unconditional single-body epsilon (1.4 for 4 bodies, :1904) */
    float ebm_uncond_coef;
} cindm_sample_config;

const char* cindm_last_error(void);
int cindm_version(void);

/* ---- tracing (the reference has none) -----------------------------------
 * cindm_launch_count: kernels this library has launched in this process (graph replays count
 * their kernel nodes).  cindm_profile_enable(1) makes every launcher bracket its kernels with CUDA
 * events on the launching stream (do not use while capturing a graph); cindm_profile_report
 * writes "class,launch_groups,total_ms,algorithmic_work" lines and returns the bytes needed. */
long long cindm_launch_count(void);
int cindm_profile_enable(int enable);
int cindm_profile_report(char* buf, int capacity);

/* ---- engine life cycle; replaces TemporalUnet1D.__init__ / GaussianDiffusion1D.__init__ /
 *      load_state_dict (model/diffusion_1d.py:519-608, :802-910; driver :163-180) ------------- */
int cindm_create(const cindm_config* cfg, cindm_engine** out);
int cindm_destroy(cindm_engine* e);
/* one call per U-Net state-dict entry, names WITHOUT the "model." prefix, host fp32 data */
int cindm_load_weight(cindm_engine* e, const char* name, const float* host, const int64_t* shape, int ndim);
/* repack weights for the kernels, build the per-timestep time-embedding bias tables
 * (time_mlp :537-542 and the 16 per-block Mish->Linear :493-497 are batch-invariant) */
int cindm_finalize_weights(cindm_engine* e, void* stream);
/* the 13 schedule buffers of GaussianDiffusion1D (:873-897), each [timesteps] fp32, host pointers
 * in the order of cindm_schedule_tables */
int cindm_set_schedule(cindm_engine* e, const float* tables13, int timesteps);
/* allocate the activation workspace for up to max_slices slices per forward */
int cindm_reserve(cindm_engine* e, int64_t max_slices, int precision);
int64_t cindm_workspace_bytes(int64_t max_slices, int precision);      /* the horizon-24, dim-64 model */
/* the same for any model cindm_create accepts (TemporalUnet1D(horizon, dim), model/diffusion_1d.py:519-608: the level
 * structure follows horizon % 8 / % 4 / % 2, :549-554) */
int64_t cindm_model_workspace_bytes(int horizon, int dim, int64_t max_slices, int precision);

/* ---- host-side helpers ------------------------------------------------------------------- */
/* cosine_beta_schedule + derived buffers (model/diffusion_1d.py:470-480, :853-897): writes
 * 13*timesteps floats: betas, alphas_cumprod, alphas_cumprod_prev, sqrt_alphas_cumprod,
 * sqrt_one_minus_alphas_cumprod, log_one_minus_alphas_cumprod, sqrt_recip_alphas_cumprod,
 * sqrt_recipm1_alphas_cumprod, posterior_variance, posterior_log_variance_clipped,
 * posterior_mean_coef1, posterior_mean_coef2, loss_weight */
int cindm_schedule_tables(int timesteps, float* out_host);
/* integer maps of the composition operator (model/diffusion_1d.py:977-990); arrays are host:
 * win_t0[W], pair_i[P], pair_j[P], cover[T_total] */
int cindm_build_index_maps(int n_bodies, int n_composed, int compose_start_step, int horizon,
                           int32_t* win_t0, int32_t* pair_i, int32_t* pair_j, int32_t* cover);

/* ---- composition operator (model/diffusion_1d.py:959-1001) ---------------------------------- */
/* x[B][T][4n] -> slices[W*P*B][24][8]   (gather at :985) */
int cindm_compose_gather(const float* x_dev, float* slices_dev, int batch, int n_bodies, int n_composed,
                         int compose_start_step, int horizon, void* stream);
/* eps_pair[W*P*B][24][8] -> eps[B][T][4n]   (scatter :989-990, reduce :994-999) */
int cindm_compose_scatter_mean(const float* eps_pair_dev, float* eps_dev, int batch, int n_bodies,
                               int n_composed, int compose_start_step, int horizon, int compose_mode,
                               void* stream);

/* ---- epsilon model: TemporalUnet1D.forward (model/diffusion_1d.py:610-646) ------------------ */
/* slices[S][24][8], one integer timestep for the whole batch -> eps_pair[S][24][8] */
int cindm_unet_forward(cindm_engine* e, const float* slices_dev, int64_t n_slices, int t,
                       float* eps_pair_dev, int precision, int conv_engine, void* stream);
/* debug / parity: copy a named intermediate activation of the LAST forward to the host as
 * channels-first fp32 [S][C][H]; names as in oracle/unet_ref.py taps */
int cindm_unet_read_tap(cindm_engine* e, const char* name, float* host, int64_t capacity_elems,
                        int64_t* s, int64_t* c, int64_t* h);
int cindm_unet_enable_taps(cindm_engine* e, int enable);

/* gather -> U-Net -> scatter-mean in one call: the compose branch of model_predictions (:959-1001) */
int cindm_composed_eps(cindm_engine* e, const float* x_dev, float* eps_dev, int batch, int n_bodies,
                       int n_composed, int compose_start_step, int compose_mode, int t, int precision,
                       int conv_engine, void* stream);
/* compose_mode "mean" of p_sample_compose_outside (:1414-1451): gather -> U-Net -> per-slice clamped x_start and
 * posterior mean (p_mean_variance on every (window, pair) slice, :1427-1433) -> mean over senders and covering
 * windows.  Outputs the composed posterior mean and x_start, both [B][T][4n]. */
int cindm_composed_posterior(cindm_engine* e, const float* x_dev, float* mean_dev, float* x0_dev, int batch, int n_bodies,
                             int n_composed, int compose_start_step, int t, int precision, int conv_engine, void* stream);

/* ---- EBM body composition (inference_1d_composing_multibodies.py:224-226 -> sample_compose_multibodies :1985-2042) ----
 * cindm_attach_unconditioned: `single` (created with transition_dim = 4, weights of the unconditional single-body model
 * loaded) becomes the model_unconditioned of `e` (:827); not owned, must outlive e's sampling calls; NULL detaches.
 * cindm_ebm_eps: gradient() for t <= 400 (:1856-1982): x[B][24][4n] -> eps[B][24][4n] =
 *   sum_{s != r} eps_pair({r,s})[slot(r)] - uncond_coef * eps_single(body r)      (uncond_coef 1.4 for 4 bodies :1904, 1 for 3 :1961)
 * cindm_ula_step: one unadjusted-Langevin update of sample_step_ULA (:2047-2073) on all frames:
 *   x_out = x + (grad_scale * eps) * step_size + sqrt(2 step_size) * noise      (grad_scale = -scalar_for_gradient[t] for t > 400)
 *   noise_dev == NULL draws Philox N(0,1) keyed by (seed, candidate, t, draw). */
int cindm_attach_unconditioned(cindm_engine* e, cindm_engine* single);
int cindm_ebm_eps(cindm_engine* e, const float* x_dev, float* eps_dev, int batch, int n_bodies, float uncond_coef, int t,
                  int precision, int conv_engine, void* stream);
int cindm_ula_step(const float* x_dev, const float* eps_dev, const float* noise_dev, float* x_out_dev, int batch, int t_total,
                   int n_bodies, float grad_scale, float step_size, uint64_t seed, int64_t candidate_offset, int t, int draw,
                   void* stream);

/* ---- guidance gradient: replaces torch.autograd.grad(design_fn(x), x) (:1316-1320) with the
 *      closed form of get_design_fn (inference/inverse_design_diffusion_1d.py:211-229) --------- */
int cindm_design_grad(const float* x_dev, float* grad_dev, int batch, int t_total, int n_bodies,
                      const cindm_objective* obj, void* stream);

/* ---- fused DDPM update: predict_start_from_noise + clamp + q_posterior + guidance + re-noise /
 *      final noise (:914-918, :938-949, :1039, :1349, :1365-1370) ------------------------------
 * pred = c1*clamp(A x - B eps) + c2 x - gscale*grad(x)
 * renoise != 0: x_out = sqrt(r_t) pred + sqrt(1-r_t) noise       (recurrence iteration)
 * renoise == 0: x_out = pred + exp(0.5 logvar_t) noise            (end of the step; noise may be NULL)
 * pred_out / x0_out may be NULL.  Needs cindm_set_schedule. */
int cindm_posterior_update(cindm_engine* e, const float* x_dev, const float* eps_dev, const float* noise_dev,
                           float* x_out_dev, float* pred_out_dev, float* x0_out_dev, int batch, int t_total,
                           int n_bodies, int t, int renoise, const cindm_objective* obj, void* stream);

/* x_start = sqrt(1/abar_t) x - sqrt(1/abar_t - 1) eps, optionally clamped to [-1, 1]: predict_start_from_noise (:914-918) with
 * the `maybe_clip` of model_predictions (:1009-1013).  elems = number of floats (a multiple of 4). */
int cindm_predict_start(cindm_engine* e, const float* x_dev, const float* eps_dev, float* x0_out_dev, int64_t elems, int t,
                        int clip, void* stream);

/* ---- the sampling loop: p_sample_loop / p_sample_compose_inside (:1655-1720, :1189-1376) ----
 * x_dev[B][T][4n] holds the initial noise on entry and the designs on exit.  noise_dev == NULL
 * draws N(0,1) from Philox4x32-10 keyed by (seed, global candidate, t, draw, element); otherwise
 * noise_dev[step][draw][B][T][4n] with draw = 0..R-1 (re-noise) then R (final), as the reference
 * consumes torch.randn_like.  x0_out_dev (optional) receives the last x_start. */
int cindm_sample(cindm_engine* e, const cindm_sample_config* cfg, float* x_dev, const float* noise_dev,
                 float* x0_out_dev, void* stream);
/* initial_state_overwrite of p_sample_compose_inside / p_sample_loop (:1273-1276, :1355-1362): ow_dev[B][rows][4n] (device, fp32,
 * owned by the caller, must stay valid while it is set) replaces the first `rows` frames of pred_img = mu - g in every
 * evaluation of the following cindm_sample calls, BEFORE the re-noise / final noise is added.  NULL / 0 clears it. */
int cindm_set_initial_state_overwrite(cindm_engine* e, const float* ow_dev, int rows);
/* DDIM sampling, `sampling_timesteps < timesteps`: replaces ddim_sample (model/diffusion_1d.py:1723-1804) with the
 * guidance / composition of p_sample_compose_inside in its epsilon-returning mode (:1372-1376).  The host passes the
 * reference's schedule for the run: n_pairs (time, time_next) pairs (:1741-1743; time_next = -1 on the last one) and,
 * per pair, coef[3] = { sqrt(alpha_next), c, sigma } evaluated in fp32 as :1778-1782 does (eta = ddim_sampling_eta).
 * Per pair: with guidance, cfg->recurrence evaluations {compose, posterior mean - g, re-noise} at `time`, then
 * x <- x_start * coef[0] + coef[1] * (eps + g) + coef[2] * noise (x <- x_start on the last pair); without guidance one
 * evaluation (model_predictions :1755).  cfg->t_start / t_end are ignored.  Guidance without recurrence is refused
 * (-5): that branch of the reference returns the posterior sample where ddim_sample expects epsilon (:1283).
 * noise_dev (optional) is [pair][draw][B][T][4n] with draws = R re-noise draws, the unused posterior draw, the DDIM draw
 * (R + 2 per pair; 1 per pair without guidance); NULL = Philox keyed by (seed, candidate, time, draw). */
int cindm_sample_ddim(cindm_engine* e, const cindm_sample_config* cfg, int n_pairs, const int32_t* times_host,
                      const int32_t* times_next_host, const float* coef_host, float* x_dev, const float* noise_dev,
                      float* x0_out_dev, void* stream);
/* fill x with the Philox N(0,1) stream used for the initial img (draw id 0xFFFF, t = timesteps) */
int cindm_fill_initial_noise(float* x_dev, int batch, int t_total, int n_bodies, uint64_t seed,
                             int64_t candidate_offset, int timesteps, void* stream);

/* ---- scoring rollout: eval_simu / simulation (utils.py:1071-1148) on the GPU ----------------
 * state0[B][n][4] in pixel units (positions 0..200, velocities px/s); runs n_steps of dt=1/60
 * hard-disc dynamics (r=20, m=1, e=1, walls of radius 1 on the 200x200 box) and writes the state
 * after steps stride-1, 2*stride-1, ... (traj[:, stride-1::stride]) into traj[B][n_steps/stride][n][4]. */
int cindm_nbody_rollout(const double* state0_dev, double* traj_dev, int batch, int n_bodies, int n_steps,
                        int stride, void* stream);
/* fused metrics of the driver (inverse_design_diffusion_1d.py:316-337): pred[B][T][4n] fp32 (normalised
 * units), runs the rollout from frame 0 and returns, per candidate, mean |cat(frame0, simulated) - pred| over
 * all T*4n entries and the mean over bodies of the last simulated frame's distance to the target (fp64) */
int cindm_score_designs(const float* pred_dev, double* mae_dev, double* objective_dev, int batch, int t_total,
                        int n_bodies, double target_x, double target_y, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CINDM_B200_H */
