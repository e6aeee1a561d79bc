// LayerNorm -> to_qkv -> linear-attention core of one attention block as ONE kernel (tcgen05 engine, 16-bit).
//
//   reference: Residual(PreNorm(dim, LinearAttentionTemporal(dim)))   model/diffusion_1d.py:272-291, LayerNorm :233-243
//
// * LayerNorm is folded into the GEMM:  qkv[r][n] = rstd_r * (sum_c (W[n][c] g[c]) x[r][c]  -  mean_r * sum_c W[n][c] g[c]),
//   so the tensor cores read the raw 16-bit activation rows (TMA, 128-byte swizzle) and the row statistics are a
//   per-row fix-up in the epilogue.  The operand W*g and its row sums are prepared once at weight load.
// * An M tile is 128/H whole slices.  The CTA that owns it runs the three 128-column passes (q | k | v) of the GEMM
//   through a double-buffered TMEM accumulator; the epilogue warps normalise and park the 16-bit q|k|v rows in a
//   shared-memory tile instead of HBM.
// * After the third pass the same sixteen warps run the attention core on that tile, one (slice, head) task per warp
//   with mma.sync (softmax over positions of k, ctx = k_s^T v, out = 32^-0.5 q ctx) while the MMA warp is already
//   working on the next tile's q and k passes.  Only the [S][H][128] attention output goes to HBM; to_out + residual
//   is the plain 1x1 tcgen05 conv that follows.
// Warp roles (576 threads): warp 0 TMA producer, warp 1 MMA issuer + TMEM allocator, warps 2-17 epilogue / attention
// (four warps per TMEM lane quarter, each owning 32 of a pass's 128 columns).
#include "tc_common.cuh"

namespace cindm {

namespace {

constexpr int kAttnStages = 3;
constexpr int kAttnStageBytes = kATileBytes + 128 * 128;      // A (128 rows) + B (128 weight rows), 64 channels each
constexpr int kTileRow = 392;                                 // halves per row of the q|k|v tile (784 B: ldmatrix conflict-free)
constexpr int kEpiWarps = 16;
constexpr int kAttnThreads = 64 + 32 * kEpiWarps;

struct AttnTcParams {
    const void* x;            // [S][H][C] 16-bit block input
    const float* wsum;        // [384] row sums of the folded weight operand
    void* out;                // [S][H][128] 16-bit attention output (before to_out)
    long long S;
    int H, C;
    int slices_per_tile, rows_used, m_tiles, k_chunks;
    int reverse;              // walk the M tiles from the last one down (see ConvTcLaunch::reverse)
};

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t addr, uint32_t (&r)[4]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
template <typename T>
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1);
template <>
__device__ __forceinline__ void mma16816<__half>(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <>
__device__ __forceinline__ void mma16816<__nv_bfloat16>(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <typename T16, int H>       // H = positions per slice at this level (3, 6, 12 or 24)
__global__ void __launch_bounds__(kAttnThreads, 1)
qkv_attn_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const AttnTcParams p) {
    constexpr int KS = H <= 16 ? 1 : 2;                           // 16-position steps of the attention products
    constexpr int n = H;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* tiles = smem;                                                             // kAttnStages x (A | B)
    T16* qkv = reinterpret_cast<T16*>(smem + kAttnStages * kAttnStageBytes);           // [128][kTileRow]
    float* wsum = reinterpret_cast<float*>(qkv + 128 * kTileRow);                      // [384]
    float2* part = reinterpret_cast<float2*>(wsum + 384);                              // [4 channel quarters][128 rows]
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(part + 512);
    uint64_t* empty_bar = full_bar + kAttnStages;
    uint64_t* tmem_full = empty_bar + kAttnStages;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < kAttnStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], kEpiWarps); }
        fence_barrier_init();
        tma_prefetch_desc(&map_a);
        tma_prefetch_desc(&map_b);
    }
    if (warp == 1) tmem_alloc(tmem_slot, 256);
    pdl_wait();                  // the previous kernel of the chain is complete: global memory may be touched from here on
    pdl_trigger();
    for (int c = threadIdx.x; c < 384; c += blockDim.x) wsum[c] = p.wsum[c];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // =============================== TMA producer ===============================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            const uint32_t tx_bytes = (uint32_t)(p.rows_used * 128 + 128 * 128);
            for (int it = blockIdx.x; it < p.m_tiles; it += gridDim.x) {
                const int m_tile = p.reverse ? p.m_tiles - 1 - it : it;
                const int s0 = m_tile * p.slices_per_tile;
                for (int pass = 0; pass < 3; ++pass) {
                    for (int kc = 0; kc < p.k_chunks; ++kc) {
                        mbar_wait_backoff(&empty_bar[stage], phase ^ 1);
                        uint8_t* a_dst = tiles + stage * kAttnStageBytes;
                        mbar_expect_tx(&full_bar[stage], tx_bytes);
                        tma_load_3d(&map_a, &full_bar[stage], a_dst, kc * kBlockK, 0, s0);
                        tma_load_2d(&map_b, &full_bar[stage], a_dst + kATileBytes, kc * kBlockK, pass * 128);
                        if (++stage == kAttnStages) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // =============================== MMA issuer ===============================
        if (lane == 0) {
            constexpr uint32_t idesc = (1u << 4) | (Fmt<T16>::kind << 7) | (Fmt<T16>::kind << 10) |
                                       ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            for (int m_tile = blockIdx.x; m_tile < p.m_tiles; m_tile += gridDim.x) {
                for (int pass = 0; pass < 3; ++pass) {
                    mbar_wait_backoff(&tmem_empty[acc], acc_phase ^ 1);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + (uint32_t)(acc * 128);
                    for (int kc = 0; kc < p.k_chunks; ++kc) {
                        mbar_wait_backoff(&full_bar[stage], phase);
                        tc_fence_after();
                        const uint32_t a_addr = smem_u32(tiles + stage * kAttnStageBytes);
                        const uint32_t b_addr = a_addr + kATileBytes;
#pragma unroll
                        for (int k = 0; k < kBlockK / 16; ++k)
                            tc_mma_f16(d_tmem, umma_smem_desc(a_addr + k * 32), umma_smem_desc(b_addr + k * 32), idesc,
                                       (kc > 0 || k > 0) ? 1u : 0u);
                        tc_commit(&empty_bar[stage]);
                        if (++stage == kAttnStages) { stage = 0; phase ^= 1; }
                    }
                    tc_commit(&tmem_full[acc]);
                    if (++acc == 2) { acc = 0; acc_phase ^= 1; }
                }
            }
        }
        __syncwarp();
    } else {
        // =============================== epilogue + attention core (16 warps) ===============================
        const int ew = warp - 2;                                  // 0..15
        const int q = warp & 3;                                   // TMEM lane quarter this warp may access
        const int cp = ew >> 2;                                   // which 32 of a pass's 128 columns (and channel quarter) this warp owns
        const int row = q * 32 + lane;                            // tile row == TMEM lane
        const int C = p.C;
        const T16* xin = reinterpret_cast<const T16*>(p.x);
        T16* out = reinterpret_cast<T16*>(p.out);
        const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(cp * 32);
        const int g = lane >> 2, t4 = lane & 3;
        // A 16-row fragment step covers more positions than a slice has: the surplus ldmatrix rows re-read the slice's own
        // last row (never another slice's rows, which another warp may be rewriting) and are cleared in the fragments
        // (k index 16*ks + 2*t4 (+1) in the low registers, + 8 in the high ones)
        uint32_t mlo[KS], mhi[KS];
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
            const int j = 16 * ks + 2 * t4;
            mlo[ks] = (j < H ? 0x0000FFFFu : 0u) | (j + 1 < H ? 0xFFFF0000u : 0u);
            mhi[ks] = (j + 8 < H ? 0x0000FFFFu : 0u) | (j + 9 < H ? 0xFFFF0000u : 0u);
        }
        const float inv_c = 1.0f / (float)C;
        int acc = 0; uint32_t acc_phase = 0;
        // The first 32 bytes of this thread's channel quarter of its row (all of it at C = 64) are fetched one tile ahead,
        // before the attention stage of the previous tile, so the statistics do not start on an exposed global load.
        uint4 pre[2] = {make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0)};
        T16 pre_x0 = (T16)0.f;
        auto prefetch_row = [&](int mt) {
            if (mt >= p.m_tiles) return;                    // (mt = position in the walk)
            const long long ps0 = (long long)(p.reverse ? p.m_tiles - 1 - mt : mt) * p.slices_per_tile;
            if (row < p.rows_used && (ps0 + row / n) < p.S) {
                const uint4* xp = reinterpret_cast<const uint4*>(xin + (ps0 * n + row) * (long long)C + cp * (C >> 2));
                pre[0] = xp[0]; pre[1] = xp[1];
                pre_x0 = xin[(ps0 * n + row) * (long long)C];
            }
        };
        prefetch_row(blockIdx.x);
        for (int it = blockIdx.x; it < p.m_tiles; it += gridDim.x) {
            const int m_tile = p.reverse ? p.m_tiles - 1 - it : it;
            const long long s0 = (long long)m_tile * p.slices_per_tile;
            const bool valid = row < p.rows_used && (s0 + row / n) < p.S;
            // ---- LayerNorm statistics of this thread's row (a quarter of the channels each; shifted sums, fixed order) ----
            float mean = 0.f, rstd = 0.f;
            {
                float s1 = 0.f, s2 = 0.f, x0 = 0.f;
                if (valid) {
                    const T16* xr = xin + (s0 * n + row) * (long long)C;
                    x0 = (float)pre_x0;
                    const uint4* xp = reinterpret_cast<const uint4*>(xr + cp * (C >> 2));
                    for (int i = 0; i < (C >> 5); i += 2) {
                        uint4 u[2];
                        if (i == 0) { u[0] = pre[0]; u[1] = pre[1]; }
                        else { u[0] = xp[i]; u[1] = xp[i + 1]; }
#pragma unroll
                        for (int j = 0; j < 2; ++j) {
                            const uint32_t w[4] = {u[j].x, u[j].y, u[j].z, u[j].w};
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const float2 f = unpack2<T16>(w[k]);
                                const float d0 = f.x - x0, d1 = f.y - x0;
                                s1 += d0 + d1;
                                s2 = fmaf(d0, d0, fmaf(d1, d1, s2));
                            }
                        }
                    }
                }
                part[cp * 128 + row] = make_float2(s1, s2);
                asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");
                const float2 p0 = part[row], p1 = part[128 + row], p2 = part[256 + row], p3 = part[384 + row];
                const float a = (p0.x + p1.x) + (p2.x + p3.x), b = (p0.y + p1.y) + (p2.y + p3.y);
                const float md = a * inv_c;
                mean = x0 + md;
                rstd = rsqrtf(fmaxf(b * inv_c - md * md, 0.f) + 1e-5f);
            }
            const float sc = rstd, nm = -mean * rstd;
            // ---- q | k | v passes: TMEM -> normalise -> 16-bit rows of the shared-memory tile.  Rows beyond the tile's slices
            //      carry stale operands; no task reads them (tasks exist for real slices only and stay inside their rows) ----
            for (int pass = 0; pass < 3; ++pass) {
                mbar_wait_backoff(&tmem_full[acc], acc_phase);
                tc_fence_after();
                {
                    float v[32];
                    tmem_ld32(lane_base + (uint32_t)(acc * 128), v);
                    const int col = pass * 128 + cp * 32;
                    const float4* ws4 = reinterpret_cast<const float4*>(wsum + col);
                    uint32_t packed[16];
                    // two columns per instruction (packed fp32: mul.f32x2 / fma.f32x2, same IEEE operations as the scalar form)
                    const unsigned long long sc_2 = f32x2_pack(sc, sc), nm_2 = f32x2_pack(nm, nm);
#pragma unroll
                    for (int i = 0; i < 32; i += 4) {
                        const float4 w4 = ws4[i >> 2];
                        float y0, y1, y2, y3;
                        f32x2_unpack(f32x2_fma(f32x2_pack(v[i], v[i + 1]), sc_2, f32x2_mul(nm_2, f32x2_pack(w4.x, w4.y))), y0, y1);
                        f32x2_unpack(f32x2_fma(f32x2_pack(v[i + 2], v[i + 3]), sc_2, f32x2_mul(nm_2, f32x2_pack(w4.z, w4.w))), y2, y3);
                        packed[i >> 1] = pack2<T16>(y0, y1);
                        packed[(i >> 1) + 1] = pack2<T16>(y2, y3);
                    }
                    uint4* dst = reinterpret_cast<uint4*>(qkv + row * kTileRow + col);
#pragma unroll
                    for (int j = 0; j < 4; ++j) dst[j] = make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tmem_empty[acc]);
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
            prefetch_row(it + gridDim.x);
            asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");             // the whole q|k|v tile is in shared memory
            // ---- attention core: one (slice, head) task per warp ----
            const long long left = p.S - s0;
            const int slices_here = left < p.slices_per_tile ? (int)left : p.slices_per_tile;
            for (int task = ew; task < slices_here * 4; task += kEpiWarps) {
                const int sl = task >> 2, h = task & 3, r0 = sl * n;
                const T16* tq = qkv + h * 32;
                // K: softmax over positions, one channel per lane, written back in place
                {
                    float kv[H];
                    float m = -INFINITY;
                    T16* tk = qkv + h * 32 + 128 + r0 * kTileRow + lane;
#pragma unroll
                    for (int j = 0; j < H; ++j) {
                        kv[j] = (float)tk[j * kTileRow];
                        m = fmaxf(m, kv[j]);
                    }
                    float sum = 0.f;
#pragma unroll
                    const float m2 = m * 1.4426950408889634f;
                    for (int j = 0; j < H; ++j) { kv[j] = ex2_approx(fmaf(kv[j], 1.4426950408889634f, -m2)); sum += kv[j]; }
                    const float inv = 1.0f / sum;
#pragma unroll
                    for (int j = 0; j < H; ++j) tk[j * kTileRow] = (T16)(kv[j] * inv);
                }
                __syncwarp();
                // ctx^T = V^T K_s : M = e (2 tiles of 16), N = d (4 tiles of 8), K = positions; positions >= n masked out of both operands
                float ct[2][4][4];
#pragma unroll
                for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
                        for (int c = 0; c < 4; ++c) ct[mt][nt][c] = 0.f;
#pragma unroll
                for (int ks = 0; ks < KS; ++ks) {
                    uint32_t bk[2][4];
#pragma unroll
                    for (int np = 0; np < 2; ++np) {
                        const int r = r0 + min(16 * ks + (lane & 7) + ((lane >> 3) & 1) * 8, n - 1), col = 16 * np + (lane >> 4) * 8;
                        ldsm_x4_trans(smem_u32(tq + 128 + r * kTileRow + col), bk[np]);
                        bk[np][0] &= mlo[ks]; bk[np][1] &= mhi[ks]; bk[np][2] &= mlo[ks]; bk[np][3] &= mhi[ks];
                    }
#pragma unroll
                    for (int mt = 0; mt < 2; ++mt) {
                        uint32_t av[4];
                        const int r = r0 + min(16 * ks + (lane & 7) + (lane >> 4) * 8, n - 1), col = 16 * mt + ((lane >> 3) & 1) * 8;
                        ldsm_x4_trans(smem_u32(tq + 256 + r * kTileRow + col), av);
                        av[0] &= mlo[ks]; av[1] &= mlo[ks]; av[2] &= mhi[ks]; av[3] &= mhi[ks];
#pragma unroll
                        for (int nt = 0; nt < 4; ++nt) mma16816<T16>(ct[mt][nt], av, bk[nt >> 1][(nt & 1) * 2], bk[nt >> 1][(nt & 1) * 2 + 1]);
                    }
                }
                // out = Q ctx : M = positions, N = e (4 tiles of 8), K = d (2 steps of 16); ctx^T accumulators are the B fragments
                T16* dst = out + (s0 + sl) * (long long)n * 128 + h * 32;
#pragma unroll
                for (int mt = 0; mt < KS; ++mt) {
                    float oc[4][4];
#pragma unroll
                    for (int ne = 0; ne < 4; ++ne)
#pragma unroll
                        for (int c = 0; c < 4; ++c) oc[ne][c] = 0.f;
#pragma unroll
                    for (int kd = 0; kd < 2; ++kd) {
                        uint32_t aq[4];
                        const int r = r0 + min(16 * mt + (lane & 7) + ((lane >> 3) & 1) * 8, n - 1), col = 16 * kd + (lane >> 4) * 8;
                        ldsm_x4(smem_u32(tq + r * kTileRow + col), aq);
#pragma unroll
                        for (int ne = 0; ne < 4; ++ne) {
                            const int m2 = ne >> 1, hi = (ne & 1) * 2;
                            const uint32_t b0 = pack2<T16>(ct[m2][2 * kd][hi], ct[m2][2 * kd][hi + 1]);
                            const uint32_t b1 = pack2<T16>(ct[m2][2 * kd + 1][hi], ct[m2][2 * kd + 1][hi + 1]);
                            mma16816<T16>(oc[ne], aq, b0, b1);
                        }
                    }
                    const int j0 = 16 * mt + g, j1 = j0 + 8;                   // (the 32^-0.5 query scale is folded into the q weights)
#pragma unroll
                    for (int ne = 0; ne < 4; ++ne) {
                        const int e = 8 * ne + 2 * t4;
                        if (j0 < n) *reinterpret_cast<uint32_t*>(dst + j0 * 128 + e) = pack2<T16>(oc[ne][0], oc[ne][1]);
                        if (j1 < n) *reinterpret_cast<uint32_t*>(dst + j1 * 128 + e) = pack2<T16>(oc[ne][2], oc[ne][3]);
                    }
                }
            }
            asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");             // the tile is rewritten by the next M tile's q pass
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 256);
    }
}

constexpr size_t attn_smem_bytes() {
    return (size_t)kAttnStages * kAttnStageBytes + 128 * kTileRow * 2 + 384 * 4 + 512 * 8 + (2 * kAttnStages + 4) * 8 + 16;
}

template <typename T16, int H>
int launch_instance(const CUtensorMap& ma, const CUtensorMap& mb, const AttnTcParams& p, cudaStream_t st) {
    auto kern = qkv_attn_kernel<T16, H>;
    constexpr size_t smem = attn_smem_bytes();
    static DeviceOnce once;
    if (once.first_time()) {
        CINDM_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    const int grid = p.m_tiles < num_sms() ? p.m_tiles : num_sms();
    CINDM_CHECK_CUDA(launch_chain(kern, dim3(grid), dim3(kAttnThreads), smem, st, ma, mb, p));
    CINDM_CHECK_LAUNCH();
    return 0;
}

}  // namespace

int launch_qkv_attn_tc(const AttnW& a, const void* x, void* out, int64_t S, int H, int C, int prec, cudaStream_t st, int reverse) {
    if (prec != PREC_F16 && prec != PREC_BF16) return fail(-2, "attn_tc: 16-bit precisions only");
    if (C % 64 || C > 512 || (H != 3 && H != 6 && H != 12 && H != 24)) return fail(-2, "attn_tc: unsupported block shape");
    if (!a.wln16[prec] || !a.wsum[prec]) return fail(-4, "attn_tc: folded LayerNorm operand missing");
    if (S == 0) return 0;
    AttnTcParams p;
    p.x = x; p.wsum = a.wsum[prec]; p.out = out; p.S = S; p.H = H; p.C = C; p.reverse = reverse;
    // whole slices per 128-row tile; at H <= 6 a multiple of 4, so that the (slice, head) tasks split evenly over the
    // 16 warps (at H = 12 / 24 the extra tiles that would take cost more than the last, partly filled round of tasks)
    p.slices_per_tile = H <= 6 ? (128 / H) / 4 * 4 : 128 / H;
    p.rows_used = p.slices_per_tile * H;
    p.m_tiles = (int)((S + p.slices_per_tile - 1) / p.slices_per_tile);
    p.k_chunks = C / kBlockK;
    char tag[64];
    snprintf(tag, sizeof tag, "attn_tc ln+qkv+core H%d C%d", H, C);
    KernelTimer kt(tag, st, 2.0 * (double)S * H * (384.0 * C + 2.0 * 128 * 32));
    CUtensorMap ma, mb;
    CINDM_TRY(encode_act_map(&ma, x, prec, S, H, C, p.slices_per_tile, H, 1));
    CINDM_TRY(encode_weight_map(&mb, a.wln16[prec], prec, 384, C, 128));
    switch (H * 2 + (prec == PREC_F16 ? 0 : 1)) {
        case 6: return launch_instance<__half, 3>(ma, mb, p, st);
        case 7: return launch_instance<__nv_bfloat16, 3>(ma, mb, p, st);
        case 12: return launch_instance<__half, 6>(ma, mb, p, st);
        case 13: return launch_instance<__nv_bfloat16, 6>(ma, mb, p, st);
        case 24: return launch_instance<__half, 12>(ma, mb, p, st);
        case 25: return launch_instance<__nv_bfloat16, 12>(ma, mb, p, st);
        case 48: return launch_instance<__half, 24>(ma, mb, p, st);
        default: return launch_instance<__nv_bfloat16, 24>(ma, mb, p, st);
    }
}

}  // namespace cindm
