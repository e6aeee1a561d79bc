// U-Net weight preparation and forward orchestration.
// Layer order and wiring follow TemporalUnet1D.forward (reference model/diffusion_1d.py:610-646);
// the schedule of kernels is ours: channels-last activations, per-t bias tables, one batched
// forward over all (window, pair, candidate) slices.
#include <cmath>
#include <cstring>

#include <cstdlib>
#include <unordered_map>
#include "engine.h"
#include "conv_tc.h"

namespace cindm {

int launch_temb_table(const float* w1, const float* b1, const float* w3, const float* b3, float* table, int dim,
                      int timesteps, cudaStream_t st);
int launch_block_time_bias(const float* temb, const float* w, const float* b, float* out, int dim, int cout,
                           int timesteps, cudaStream_t st);

// Number of Downsample1d / Upsample1d stages the reference builds for a horizon (model/diffusion_1d.py:549-554, :575-599):
// three when horizon % 8 == 0 (24 -> 12 -> 6 -> 3), two when only % 4 (the 44-step models: 44 -> 22 -> 11, and the 512-channel
// level stays at 11 positions), one when only % 2; the other levels keep nn.Identity in those slots.
int down_samplings(int horizon) {
    if (horizon <= 0) return 0;
    if (horizon % 8 == 0) return 3;
    if (horizon % 4 == 0) return 2;
    if (horizon % 2 == 0) return 1;
    return 0;
}

namespace {

struct Uploader {
    cindm_engine* e;
    cudaStream_t st;
    int rc = 0;

    const HostTensor* find(const std::string& name, bool required = true) {
        auto it = e->host_weights.find(name);
        if (it == e->host_weights.end()) {
            if (required && rc == 0) rc = fail(-3, "missing weight: " + name);
            return nullptr;
        }
        return &it->second;
    }

    void* upload(const void* host, size_t bytes) {
        void* d = nullptr;
        if (cudaMalloc(&d, bytes) != cudaSuccess) {
            if (rc == 0) rc = fail(-100, "cudaMalloc failed while uploading weights");
            return nullptr;
        }
        e->allocations.push_back(d);
        if (cudaMemcpyAsync(d, host, bytes, cudaMemcpyHostToDevice, st) != cudaSuccess && rc == 0)
            rc = fail(-100, "cudaMemcpy failed while uploading weights");
        cudaStreamSynchronize(st);   // host staging buffers are temporaries
        return d;
    }

    float* vec(const std::string& name, int64_t expect, bool required = true) {
        const HostTensor* t = find(name, required);
        if (!t) return nullptr;
        if ((int64_t)t->data.size() != expect) {
            if (rc == 0) rc = fail(-3, "bad size for " + name);
            return nullptr;
        }
        return (float*)upload(t->data.data(), t->data.size() * sizeof(float));
    }

    // torch Conv1d weight [cout][cin][k] (or ConvTranspose1d [cin][cout][k]) -> operands
    void conv(ConvW& c, const std::string& prefix, int cin, int cout, int taps, bool bias, bool transposed = false) {
        c.cin = cin; c.cout = cout; c.taps = taps;
        const HostTensor* t = find(prefix + ".weight");
        if (!t) return;
        if ((int64_t)t->data.size() != (int64_t)cin * cout * taps) {
            if (rc == 0) rc = fail(-3, "bad size for " + prefix + ".weight");
            return;
        }
        std::vector<float> w((size_t)taps * cin * cout);
        std::vector<__half> wh(w.size());
        std::vector<__nv_bfloat16> wb(w.size());
        for (int k = 0; k < taps; ++k)
            for (int ci = 0; ci < cin; ++ci)
                for (int co = 0; co < cout; ++co) {
                    float v = transposed ? t->data[((size_t)ci * cout + co) * taps + k]
                                         : t->data[((size_t)co * cin + ci) * taps + k];
                    w[((size_t)k * cin + ci) * cout + co] = v;
                    size_t km = ((size_t)k * cout + co) * cin + ci;       // K-major tensor-core operand
                    wh[km] = __float2half_rn(v);
                    wb[km] = __float2bfloat16_rn(v);
                }
        c.w = (float*)upload(w.data(), w.size() * sizeof(float));
        c.w16[PREC_F16] = upload(wh.data(), wh.size() * sizeof(__half));
        c.w16[PREC_BF16] = upload(wb.data(), wb.size() * sizeof(__nv_bfloat16));
        c.bias = bias ? vec(prefix + ".bias", cout) : nullptr;
    }

    // Block-Toeplitz operand of a k=5, pad=2 conv over 3 positions: rows n = (group, out position p, channel in
    // group), columns k = (in position q, ci); entry = W[co][ci][q - p + 2] (every tap index is in range at H=3).
    void toeplitz3(ConvW& c, const std::string& prefix) {
        const HostTensor* t = find(prefix + ".weight");
        if (!t || rc) return;
        const int cin = c.cin, cout = c.cout, cpg = cout / 8;
        std::vector<__half> wh((size_t)9 * cin * cout);
        std::vector<__nv_bfloat16> wb(wh.size());
        for (int g = 0; g < 8; ++g)
            for (int p = 0; p < 3; ++p)
                for (int cc = 0; cc < cpg; ++cc) {
                    const int co = g * cpg + cc;
                    const size_t n = ((size_t)g * 3 + p) * cpg + cc;
                    for (int q = 0; q < 3; ++q)
                        for (int ci = 0; ci < cin; ++ci) {
                            float v = t->data[((size_t)co * cin + ci) * 5 + (q - p + 2)];
                            size_t idx = n * (3 * (size_t)cin) + (size_t)q * cin + ci;
                            wh[idx] = __float2half_rn(v);
                            wb[idx] = __float2bfloat16_rn(v);
                        }
                }
        c.w16t[PREC_F16] = upload(wh.data(), wh.size() * sizeof(__half));
        c.w16t[PREC_BF16] = upload(wb.data(), wb.size() * sizeof(__nv_bfloat16));
    }

    void resblock(ResBlockW& r, const std::string& prefix, int cin, int cout, bool bottom = false) {
        r.name = prefix;
        conv(r.conv0, prefix + ".blocks.0.block.0", cin, cout, 5, true);
        if (bottom && e->cfg.horizon == 24) {
            toeplitz3(r.conv0, prefix + ".blocks.0.block.0");
        }
        r.gn0.gamma = vec(prefix + ".blocks.0.block.2.weight", cout);
        r.gn0.beta = vec(prefix + ".blocks.0.block.2.bias", cout);
        conv(r.conv1, prefix + ".blocks.1.block.0", cout, cout, 5, true);
        if (bottom && e->cfg.horizon == 24) toeplitz3(r.conv1, prefix + ".blocks.1.block.0");
        r.gn1.gamma = vec(prefix + ".blocks.1.block.2.weight", cout);
        r.gn1.beta = vec(prefix + ".blocks.1.block.2.bias", cout);
        r.has_res = cin != cout;
        if (r.has_res) conv(r.res, prefix + ".residual_conv", cin, cout, 1, true);
        // per-timestep bias table
        const int dim = e->cfg.dim, T = e->cfg.timesteps;
        float* w = vec(prefix + ".time_mlp.1.weight", (int64_t)cout * dim);
        float* b = vec(prefix + ".time_mlp.1.bias", cout);
        if (rc) return;
        float* table = nullptr;
        if (cudaMalloc(&table, (size_t)T * cout * sizeof(float)) != cudaSuccess) {
            rc = fail(-100, "cudaMalloc(time bias table)");
            return;
        }
        e->allocations.push_back(table);
        r.time_bias = table;
        int k = launch_block_time_bias(e->temb_table, w, b, table, dim, cout, T, st);
        if (k && rc == 0) rc = k;
    }

    void attn(AttnW& a, const std::string& prefix, int c) {
        a.name = prefix;
        a.g = vec(prefix + ".fn.norm.g", c);
        conv(a.qkv, prefix + ".fn.fn.to_qkv", c, 384, 1, false);
        conv(a.out, prefix + ".fn.fn.to_out", 128, c, 1, true);
        // folded operand of the fused tcgen05 attention kernel: W'[n][ci] = W[n][ci] * g[ci] (the q rows also carry the
        // 32^-0.5 query scale of :284), and its row sums
        const HostTensor* tg = find(prefix + ".fn.norm.g");
        const HostTensor* tw = find(prefix + ".fn.fn.to_qkv.weight");
        if (!tg || !tw || rc) return;
        std::vector<__half> wh((size_t)384 * c);
        std::vector<__nv_bfloat16> wb(wh.size());
        std::vector<float> sh(384), sb(384);
        for (int n = 0; n < 384; ++n) {
            float ah = 0.f, ab = 0.f;
            for (int ci = 0; ci < c; ++ci) {
                const float v = tw->data[(size_t)n * c + ci] * tg->data[ci] * (n < 128 ? 0.17677669529663687f : 1.0f);
                wh[(size_t)n * c + ci] = __float2half_rn(v);
                wb[(size_t)n * c + ci] = __float2bfloat16_rn(v);
                ah += __half2float(wh[(size_t)n * c + ci]);
                ab += __bfloat162float(wb[(size_t)n * c + ci]);
            }
            sh[n] = ah; sb[n] = ab;
        }
        a.wln16[PREC_F16] = upload(wh.data(), wh.size() * sizeof(__half));
        a.wln16[PREC_BF16] = upload(wb.data(), wb.size() * sizeof(__nv_bfloat16));
        a.wsum[PREC_F16] = (float*)upload(sh.data(), sh.size() * sizeof(float));
        a.wsum[PREC_BF16] = (float*)upload(sb.data(), sb.size() * sizeof(float));
    }
};

}  // namespace

int finalize_weights(cindm_engine* e, cudaStream_t st) {
    if (e->finalized) return fail(-4, "weights already finalized");
    const int dim = e->cfg.dim, T = e->cfg.timesteps, F = e->cfg.transition_dim;
    const int n_down = down_samplings(e->cfg.horizon);
    if (n_down == 0) return fail(-2, "horizon must be even (the reference defines no U-Net for odd horizons, :549-554)");
    Uploader up{e, st};
    // time_mlp -> temb table [T][dim]
    float* w1 = up.vec("time_mlp.1.weight", (int64_t)4 * dim * dim);
    float* b1 = up.vec("time_mlp.1.bias", 4 * dim);
    float* w3 = up.vec("time_mlp.3.weight", (int64_t)4 * dim * dim);
    float* b3 = up.vec("time_mlp.3.bias", dim);
    if (up.rc) return up.rc;
    CINDM_CHECK_CUDA(cudaMalloc(&e->temb_table, (size_t)T * dim * sizeof(float)));
    e->allocations.push_back(e->temb_table);
    CINDM_TRY(launch_temb_table(w1, b1, w3, b3, e->temb_table, dim, T, st));

    const int ch[5] = {F, dim, dim * 2, dim * 4, dim * 8};
    for (int i = 0; i < 4; ++i) {
        std::string p = "downs." + std::to_string(i);
        up.resblock(e->downs_rb[i][0], p + ".0", ch[i], ch[i + 1], i == 3);
        up.resblock(e->downs_rb[i][1], p + ".1", ch[i + 1], ch[i + 1], i == 3);
        up.attn(e->downs_at[i], p + ".2", ch[i + 1]);
        if (i < n_down) up.conv(e->down_conv[i], p + ".3.conv", ch[i + 1], ch[i + 1], 3, true);
    }
    up.resblock(e->mid_rb[0], "mid_block1", ch[4], ch[4], true);
    up.attn(e->mid_at, "mid_attn", ch[4]);
    up.resblock(e->mid_rb[1], "mid_block2", ch[4], ch[4], true);
    for (int i = 0; i < 3; ++i) {
        std::string p = "ups." + std::to_string(i);
        int co = ch[4 - i], ci = ch[3 - i];      // (dim_in, dim_out) of reversed(in_out[1:]): ci -> co going down
        up.resblock(e->ups_rb[i][0], p + ".0", co * 2, co, i == 0);
        up.resblock(e->ups_rb[i][1], p + ".1", co, ci, i == 0);
        up.attn(e->ups_at[i], p + ".2", ci);
        if (i >= 3 - n_down) up.conv(e->up_conv[i], p + ".3.conv", ci, ci, 4, true, /*transposed=*/true);
    }
    up.conv(e->final_block, "final_conv.0.block.0", dim, dim, 5, true);
    e->final_gn.gamma = up.vec("final_conv.0.block.2.weight", dim);
    e->final_gn.beta = up.vec("final_conv.0.block.2.bias", dim);
    up.conv(e->final_out, "final_conv.1", dim, F, 1, true);
    if (up.rc) return up.rc;
    CINDM_CHECK_CUDA(cudaStreamSynchronize(st));
    e->finalized = true;
    e->host_weights.clear();
    return 0;
}

// ---------------------------------------------------------------------------------------------

// Positions x channels of the largest activation tensor of one slice.  With three down-samplings (horizon % 8 == 0) every level
// holds horizon * dim values; a model that keeps its resolution on the lower levels (horizon % 4 / % 2, reference :549-554)
// holds more there: level l has dim * 2^l channels at horizon / 2^min(l, n_down) positions.
int64_t act_elems_per_slice(int horizon, int dim) {
    const int n_down = down_samplings(horizon);
    int64_t best = 0;
    for (int l = 0; l < 4; ++l) {
        const int64_t v = (int64_t)(horizon >> (l < n_down ? l : n_down)) * dim * (1 << l);
        if (v > best) best = v;
    }
    return best;
}

int64_t workspace_bytes(int64_t S, int prec, int horizon, int dim) {
    const size_t es = elem_size(prec);
    const int64_t per_act = S * act_elems_per_slice(horizon, dim);
    auto al = [](int64_t b) { return (b + 1023) / 1024 * 1024; };
    int64_t total = 0;
    total += 3 * al(per_act * es);                  // act[3]
    total += 3 * al(per_act * es);                  // skip[3]
    total += al(per_act * es) * 2;                  // res, ln
    total += al(S * (int64_t)horizon * 384 * es);   // qkv
    total += al(S * (int64_t)horizon * 128 * es);   // att
    total += al(per_act * 4);                       // scratch fp32
    total += 2 * al(S * (int64_t)horizon * 8 * 4);  // slices, eps_pair
    return total;
}

int reserve_workspace(cindm_engine* e, int64_t S, int prec) {
    if (prec < 0 || prec > 2) return fail(-2, "bad precision");
    Workspace& w = e->ws;
    if (w.base && w.max_slices >= S && w.precision == prec) return 0;
    if (w.base) {
        CINDM_CHECK_CUDA(cudaDeviceSynchronize());
        graph_cache_clear(e);
        CINDM_CHECK_CUDA(cudaFree(w.base));
        w = Workspace();
    }
    const int H = e->cfg.horizon;
    const size_t es = elem_size(prec);
    const int64_t per_act = S * act_elems_per_slice(H, e->cfg.dim);
    size_t bytes = (size_t)workspace_bytes(S, prec, H, e->cfg.dim);
    CINDM_CHECK_CUDA(cudaMalloc(&w.base, bytes));
    w.bytes = bytes;
    char* p = (char*)w.base;
    auto take = [&](int64_t b) { void* r = p; p += (b + 1023) / 1024 * 1024; return r; };
    for (int i = 0; i < 3; ++i) w.act[i] = take(per_act * es);
    for (int i = 0; i < 3; ++i) w.skip[i] = take(per_act * es);
    w.res = take(per_act * es);
    w.ln = take(per_act * es);
    w.qkv = take(S * (int64_t)H * 384 * es);
    w.att = take(S * (int64_t)H * 128 * es);
    w.scratch = (float*)take(per_act * 4);
    w.slices = (float*)take(S * (int64_t)H * 8 * 4);
    w.eps_pair = (float*)take(S * (int64_t)H * 8 * 4);
    w.max_slices = S;
    w.precision = prec;
    return 0;
}

// ---------------------------------------------------------------------------------------------

namespace {

struct Fwd {
    cindm_engine* e;
    int64_t S;
    int t;
    const int* t_dev;
    int prec, engine;
    cudaStream_t st;
    bool fork = false;                      // run residual 1x1 convs on e->side_stream (small batches)
    // Tile-walk direction of the tensor-core launches ("snake").  An activation tensor (132 MB at C4) does not fit the 126 MB L2
    // next to the traffic of the layer that wrote it, but its LAST-written part does: a consumer that walks its tiles in the
    // opposite direction to the producer of its main input starts on data that is still in L2 instead of re-reading HBM.
    // wdir[buffer] = direction the buffer was last written in (stem and SIMT kernels: front to back).
    // CINDM_SNAKE=0 walks every layer front to back (A/B runs).
    std::unordered_map<const void*, int> wdir;
    int dir_for(const void* main_input, const void* out) {
        static int snake = -1;
        if (snake < 0) { const char* v = getenv("CINDM_SNAKE"); snake = (v && v[0] == '0') ? 0 : 1; }
        if (!snake) return 0;
        auto it = wdir.find(main_input);
        const int d = (it == wdir.end() ? 0 : it->second) ^ 1;
        wdir[out] = d;
        return d;
    }

    int record_tap(const std::string& name, const void* ptr, int c, int h) {
        if (!e->taps_enabled) return 0;
        size_t bytes = (size_t)S * c * h * elem_size(prec);
        void* copy = nullptr;
        CINDM_CHECK_CUDA(cudaMalloc(&copy, bytes));
        e->tap_allocs.push_back(copy);
        CINDM_CHECK_CUDA(cudaMemcpyAsync(copy, ptr, bytes, cudaMemcpyDeviceToDevice, st));
        e->taps[name] = Tap{copy, c, h, prec, S};
        return 0;
    }

    // Conv1dBlock: conv(k5,pad2) -> GroupNorm(8) -> Mish, then + (time bias | residual)
    int conv_block(const ConvW& w, const NormW& gn, const void* in0, int c0, const void* in1, int c1, int in_prec,
                   int H, const float* add_vec, const void* add_res, void* out) {
        if (engine == CINDM_CONV_TCGEN05 && in_prec != PREC_F32) {
            ConvTcLaunch a;
            a.in0 = in0; a.c0 = c0; a.in1 = in1; a.c1 = c1; a.w = &w; a.gn = &gn;
            a.add_vec = add_vec; a.t_dev = add_vec ? t_dev : nullptr; a.add_res = add_res; a.out = out; a.S = S;
            a.H = H; a.prec = prec;
            a.epilogue = (H == 3 && w.w16t[prec] != nullptr && e->use_toeplitz) ? EPI_GN_MISH_T3 : EPI_GN_MISH;
            a.reverse = dir_for(in0, out);
            return launch_conv_tc(a, st);
        }
        ConvLaunch a;
        a.in0 = in0; a.c0 = c0; a.in1 = in1; a.c1 = c1; a.w = &w; a.out = e->ws.scratch; a.S = S;
        a.Hin = H; a.Hout = H; a.stride = 1; a.pad = w.taps / 2; a.in_prec = in_prec; a.out_prec = PREC_F32;
        CINDM_TRY(launch_conv_simt(a, st));
        return launch_gn_mish(e->ws.scratch, gn, add_vec, add_vec ? t_dev : nullptr, add_res, out, S, H, w.cout, prec, st);
    }

    // plain conv (1x1, strided, transposed) with bias and optional residual
    int conv_plain(const ConvW& w, const void* in0, int c0, const void* in1, int c1, int in_prec, int Hin, int Hout,
                   int stride, int pad, int transposed, const void* res, void* out, int out_prec, cudaStream_t on = nullptr) {
        cudaStream_t st = on ? on : this->st;
        if (engine == CINDM_CONV_TCGEN05 && in_prec != PREC_F32 && out_prec != PREC_F32 &&
            (w.cin % 64) == 0 && (w.cout % 64) == 0) {
            ConvTcLaunch a;
            a.in0 = in0; a.c0 = c0; a.in1 = in1; a.c1 = c1; a.w = &w; a.gn = nullptr;
            a.add_vec = nullptr; a.add_res = res; a.out = out; a.S = S; a.H = Hin; a.prec = prec;
            a.epilogue = EPI_BIAS;
            a.mode = transposed ? TC_UP : (stride == 2 ? TC_DOWN : TC_SAME);
            if (transposed && fork && !on) { a.side = e->side_stream; a.ev_fork = e->ev_fork; a.ev_join = e->ev_join; }
            a.reverse = dir_for(in0, out);
            return launch_conv_tc(a, st);
        }
        ConvLaunch a;
        a.in0 = in0; a.c0 = c0; a.in1 = in1; a.c1 = c1; a.w = &w; a.res = res; a.out = out; a.S = S;
        a.Hin = Hin; a.Hout = Hout; a.stride = stride; a.pad = pad; a.transposed = transposed;
        a.in_prec = in_prec; a.out_prec = out_prec;
        return launch_conv_simt(a, st);
    }

    // ResidualTemporalBlock; input may be a channel concat (in1 != null). `tmp` holds the first block's output.
    int resblock(const ResBlockW& r, const void* in0, int c0, const void* in1, int c1, int in_prec, int H, void* tmp,
                 void* out) {
        const int cout = r.conv0.cout;
        const float* tb = t_dev ? r.time_bias : r.time_bias + (size_t)t * cout;
        const void* resid = in0;
        if (r.has_res && fork) {
            // fork: the residual conv only needs the block input; it runs beside conv0 and is joined before conv1 reads it.
            // (The fork event also orders it after the previous block's conv1, the last reader of ws.res.)
            CINDM_CHECK_CUDA(cudaEventRecord(e->ev_fork, st));
            CINDM_CHECK_CUDA(cudaStreamWaitEvent(e->side_stream, e->ev_fork, 0));
            CINDM_TRY(conv_plain(r.res, in0, c0, in1, c1, in_prec, H, H, 1, 0, 0, nullptr, e->ws.res, prec, e->side_stream));
            CINDM_CHECK_CUDA(cudaEventRecord(e->ev_join, e->side_stream));
            CINDM_TRY(conv_block(r.conv0, r.gn0, in0, c0, in1, c1, in_prec, H, tb, nullptr, tmp));
            CINDM_CHECK_CUDA(cudaStreamWaitEvent(st, e->ev_join, 0));
            return conv_block(r.conv1, r.gn1, tmp, cout, nullptr, 0, prec, H, nullptr, e->ws.res, out);
        }
        CINDM_TRY(conv_block(r.conv0, r.gn0, in0, c0, in1, c1, in_prec, H, tb, nullptr, tmp));
        if (r.has_res) {
            CINDM_TRY(conv_plain(r.res, in0, c0, in1, c1, in_prec, H, H, 1, 0, 0, nullptr, e->ws.res, prec));
            resid = e->ws.res;
        } else if (in1 != nullptr || in_prec != prec) {
            return fail(-5, "identity residual needs a single input of the activation type");
        }
        return conv_block(r.conv1, r.gn1, tmp, cout, nullptr, 0, prec, H, nullptr, resid, out);
    }

    // x + to_out(linattn(LayerNorm(x)))
    int attention(const AttnW& a, const void* x, int C, int H, void* out) {
        if (engine == CINDM_CONV_TCGEN05 && prec != PREC_F32 && e->use_fused_attn) {
            CINDM_TRY(launch_qkv_attn_tc(a, x, e->ws.att, S, H, C, prec, st, dir_for(x, e->ws.att)));
            return conv_plain(a.out, e->ws.att, 128, nullptr, 0, prec, H, H, 1, 0, 0, x, out, prec);
        }
        CINDM_TRY(launch_layernorm(x, a.g, e->ws.ln, S * H, C, prec, st));
        CINDM_TRY(conv_plain(a.qkv, e->ws.ln, C, nullptr, 0, prec, H, H, 1, 0, 0, nullptr, e->ws.qkv, prec));
        CINDM_TRY(launch_attn_core(e->ws.qkv, e->ws.att, S, H, prec, st));
        return conv_plain(a.out, e->ws.att, 128, nullptr, 0, prec, H, H, 1, 0, 0, x, out, prec);
    }
};

}  // namespace

int unet_forward(cindm_engine* e, const float* slices, int64_t S, int t, const int* t_dev, float* eps_pair,
                 int precision, int conv_engine, cudaStream_t st, const GatherSpec* gather) {
    if (gather && precision == PREC_F32) return fail(-5, "fused gather is part of the 16-bit stem kernel");
    if (!e->finalized) return fail(-4, "weights not finalized");
    if (!t_dev && (t < 0 || t >= e->cfg.timesteps)) return fail(-2, "timestep out of range");
    if (S > e->ws.max_slices || e->ws.precision != precision)
        return fail(-6, "workspace not reserved for this slice count / precision (call cindm_reserve)");
    if (conv_engine == CINDM_CONV_TCGEN05 && precision == PREC_F32)
        return fail(-2, "the tcgen05 conv engine needs a 16-bit precision");
    if (precision != PREC_F32 && !(e->cfg.horizon == 24 && e->cfg.dim == 64))
        return fail(-2, "the 16-bit kernels (fused stem / head, tcgen05 convs) are built for the horizon-24, dim-64 model; "
                        "other models (44-step, Unet_dim 96) run with precision fp32 on the simt engine");
    const int n_down = down_samplings(e->cfg.horizon);
    if (e->taps_enabled) {
        for (void* p : e->tap_allocs) cudaFree(p);
        e->tap_allocs.clear();
        e->taps.clear();
    }
    Workspace& w = e->ws;
    Fwd f{e, S, t, t_dev, precision, conv_engine, st};
    // small batches leave most SMs idle inside every kernel: overlap what is independent (CINDM_FORK_RES=0 / 1 overrides)
    const bool want_fork = e->fork_residual >= 0 ? e->fork_residual == 1 : S <= 2048;
    if (want_fork && conv_engine == CINDM_CONV_TCGEN05 && precision != PREC_F32 && !profiling_enabled()) {
        if (!e->side_stream) {
            CINDM_CHECK_CUDA(cudaStreamCreateWithFlags(&e->side_stream, cudaStreamNonBlocking));
            CINDM_CHECK_CUDA(cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming));
            CINDM_CHECK_CUDA(cudaEventCreateWithFlags(&e->ev_join, cudaEventDisableTiming));
        }
        f.fork = true;
    }
    const int dim = e->cfg.dim, F = e->cfg.transition_dim;
    int H = e->cfg.horizon;
    const int ch[5] = {F, dim, dim * 2, dim * 4, dim * 8};

    // rotating activation buffers
    void* cur = nullptr;
    int cur_i = -1;
    auto next_buf = [&](int avoid_a, int avoid_b) {
        for (int i = 0; i < 3; ++i)
            if (i != avoid_a && i != avoid_b) return i;
        return 0;
    };

    const void* x_in = slices;
    int x_prec = PREC_F32;
    for (int i = 0; i < 4; ++i) {
        int tmp_i = next_buf(cur_i, -1), out_i = next_buf(cur_i, tmp_i);
        if (i == 0 && precision != PREC_F32) {
            // fused stem: [gather +] conv0 + GN + Mish + time bias, and the 1x1 residual conv, in one kernel
            const ResBlockW& r = e->downs_rb[0][0];
            StemLaunch sl;
            sl.x = gather ? gather->x : slices;
            if (gather) { sl.gather = 1; sl.B = gather->B; sl.n = gather->n; sl.start = gather->start;
                          sl.T = e->cfg.horizon + gather->nc * gather->start; }
            sl.conv0 = &r.conv0; sl.gn = &r.gn0; sl.res = &r.res;
            sl.tbias = t_dev ? r.time_bias : r.time_bias + (size_t)t * r.conv0.cout; sl.t_dev = t_dev;
            sl.out_b0 = w.act[tmp_i]; sl.out_res = w.res; sl.S = S; sl.prec = precision;
            CINDM_TRY(launch_stem(sl, st));
            CINDM_TRY(f.conv_block(r.conv1, r.gn1, w.act[tmp_i], r.conv0.cout, nullptr, 0, precision, H, nullptr, w.res,
                                   w.act[out_i]));
        } else {
            CINDM_TRY(f.resblock(e->downs_rb[i][0], x_in, ch[i], nullptr, 0, x_prec, H, w.act[tmp_i], w.act[out_i]));
        }
        cur = w.act[out_i]; cur_i = out_i; x_in = cur; x_prec = precision;
        CINDM_TRY(f.record_tap("downs." + std::to_string(i) + ".0", cur, ch[i + 1], H));
        tmp_i = next_buf(cur_i, -1); out_i = next_buf(cur_i, tmp_i);
        CINDM_TRY(f.resblock(e->downs_rb[i][1], cur, ch[i + 1], nullptr, 0, precision, H, w.act[tmp_i], w.act[out_i]));
        cur = w.act[out_i]; cur_i = out_i;
        CINDM_TRY(f.record_tap("downs." + std::to_string(i) + ".1", cur, ch[i + 1], H));
        // attention output goes to the skip buffer when this level feeds an up block, else rotates
        if (i >= 1) {
            CINDM_TRY(f.attention(e->downs_at[i], cur, ch[i + 1], H, w.skip[i - 1]));
            cur = w.skip[i - 1]; cur_i = -1;
        } else {
            out_i = next_buf(cur_i, -1);
            CINDM_TRY(f.attention(e->downs_at[i], cur, ch[i + 1], H, w.act[out_i]));
            cur = w.act[out_i]; cur_i = out_i;
        }
        CINDM_TRY(f.record_tap("downs." + std::to_string(i) + ".2", cur, ch[i + 1], H));
        if (i < n_down) {
            out_i = next_buf(cur_i, -1);
            CINDM_TRY(f.conv_plain(e->down_conv[i], cur, ch[i + 1], nullptr, 0, precision, H, H / 2, 2, 1, 0, nullptr,
                                   w.act[out_i], precision));
            H /= 2;
            cur = w.act[out_i]; cur_i = out_i;
            CINDM_TRY(f.record_tap("downs." + std::to_string(i) + ".3", cur, ch[i + 1], H));
        }
        x_in = cur;
    }
    // here cur == skip[2] (dim*8 channels at the lowest resolution); it is both the mid input and the first up-block's skip
    {
        int tmp_i = next_buf(cur_i, -1), out_i = next_buf(cur_i, tmp_i);
        CINDM_TRY(f.resblock(e->mid_rb[0], cur, ch[4], nullptr, 0, precision, H, w.act[tmp_i], w.act[out_i]));
        cur = w.act[out_i]; cur_i = out_i;
        CINDM_TRY(f.record_tap("mid_block1", cur, ch[4], H));
        out_i = next_buf(cur_i, -1);
        CINDM_TRY(f.attention(e->mid_at, cur, ch[4], H, w.act[out_i]));
        cur = w.act[out_i]; cur_i = out_i;
        CINDM_TRY(f.record_tap("mid_attn", cur, ch[4], H));
        tmp_i = next_buf(cur_i, -1); out_i = next_buf(cur_i, tmp_i);
        CINDM_TRY(f.resblock(e->mid_rb[1], cur, ch[4], nullptr, 0, precision, H, w.act[tmp_i], w.act[out_i]));
        cur = w.act[out_i]; cur_i = out_i;
        CINDM_TRY(f.record_tap("mid_block2", cur, ch[4], H));
    }
    for (int i = 0; i < 3; ++i) {
        const int co = ch[4 - i], ci = ch[3 - i];
        const void* skip = w.skip[2 - i];
        int tmp_i = next_buf(cur_i, -1), out_i = next_buf(cur_i, tmp_i);
        // torch.cat((x, h.pop()), dim=1): x first, then the skip (:637)
        CINDM_TRY(f.resblock(e->ups_rb[i][0], cur, co, skip, co, precision, H, w.act[tmp_i], w.act[out_i]));
        cur = w.act[out_i]; cur_i = out_i;
        CINDM_TRY(f.record_tap("ups." + std::to_string(i) + ".0", cur, co, H));
        tmp_i = next_buf(cur_i, -1); out_i = next_buf(cur_i, tmp_i);
        CINDM_TRY(f.resblock(e->ups_rb[i][1], cur, co, nullptr, 0, precision, H, w.act[tmp_i], w.act[out_i]));
        cur = w.act[out_i]; cur_i = out_i;
        CINDM_TRY(f.record_tap("ups." + std::to_string(i) + ".1", cur, ci, H));
        out_i = next_buf(cur_i, -1);
        CINDM_TRY(f.attention(e->ups_at[i], cur, ci, H, w.act[out_i]));
        cur = w.act[out_i]; cur_i = out_i;
        CINDM_TRY(f.record_tap("ups." + std::to_string(i) + ".2", cur, ci, H));
        if (i < 3 - n_down) continue;          // nn.Identity in the Upsample1d slot (:585, :594)
        out_i = next_buf(cur_i, -1);
        CINDM_TRY(f.conv_plain(e->up_conv[i], cur, ci, nullptr, 0, precision, H, H * 2, 2, 1, 1, nullptr, w.act[out_i],
                               precision));
        H *= 2;
        cur = w.act[out_i]; cur_i = out_i;
        CINDM_TRY(f.record_tap("ups." + std::to_string(i) + ".3", cur, ci, H));
    }
    {
        int out_i = next_buf(cur_i, -1);
        CINDM_TRY(f.conv_block(e->final_block, e->final_gn, cur, dim, nullptr, 0, precision, H, nullptr, nullptr,
                               w.act[out_i]));
        cur = w.act[out_i]; cur_i = out_i;
        CINDM_TRY(f.record_tap("final_conv.0", cur, dim, H));
        if (precision != PREC_F32) CINDM_TRY(launch_head(cur, e->final_out, eps_pair, S * H, precision, st));
        else CINDM_TRY(f.conv_plain(e->final_out, cur, dim, nullptr, 0, precision, H, H, 1, 0, 0, nullptr, eps_pair, PREC_F32));
    }
    return 0;
}

}  // namespace cindm
