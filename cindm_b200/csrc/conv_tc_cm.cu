// GroupNorm + Mish temporal convolution with a CHANNEL-MAJOR accumulator (sm_100a, tcgen05 / TMEM / TMA).
//
// The same implicit GEMM as conv_tc.cu with the operand roles swapped:
//
//   D[co][row] = sum_{tap, ci} W[tap][co][ci] * X_tap[row][ci]        M = output channels, N = 240 rows = whole slices
//
// so that a TMEM lane is an output CHANNEL and the columns of one accumulator are (slice, position).  For the epilogue
// that turns the row-major kernel's costs inside out:
//   * bias / gamma / beta / time-embedding value of a thread's channel are four registers (the row-major epilogue reads
//     them from shared memory once per element: ~1.1 LDS.128 per output element-warp);
//   * a slice's H positions of one channel sit in one thread: per-thread sums, then a shuffle tree over the CPG adjacent
//     lanes of the channel group give the GroupNorm statistics - no shared-memory partials, no CTA barriers, ONE pass
//     over the accumulator, every epilogue warp streams through its slices independently;
//   * normalisation and affine collapse into one FFMA per element (per-(slice, channel) scale / shift).
// The price: activations are channels-last, so a thread's outputs are 2-byte values C apart; a warp instruction stores 32
// consecutive channels of one row (64 contiguous bytes, two full sectors).
//
// cout = 64 (DUAL): M = 64 MMAs occupy 16 lanes of every 32-lane TMEM sub-partition; two row tiles are accumulated side
// by side (lane offset 16), so all 128 lanes carry work.
// Warp roles: warp 0 TMA producer, warp 1 MMA issuer + TMEM allocator, warps 2.. epilogue (EW / 4 per lane quarter, the
// slices of a tile dealt round-robin among the warps of a quarter).
#include "conv_tc.h"
#include "tc_common.cuh"

namespace cindm {

namespace {

constexpr int kRows = 240;                    // rows (slice positions) per tile: 10 x 24 = 20 x 12 = 40 x 6
constexpr int kXTileBytes = kRows * 128;      // one K block (64 channels) of a row tile
constexpr int kAccStride = 256;               // TMEM columns between the two accumulators

struct CmParams {
    const float* bias;        // [cout] or null
    const float* gamma;       // [cout]
    const float* beta;        // [cout]
    const float* add_vec;     // [cout] or [timesteps][cout] when t_dev != null
    const int* t_dev;
    const void* add_res;      // [S][H][cout] 16-bit or null
    void* out;                // [S][H][cout] 16-bit
    long long S;
    int cout, c0, c1, taps, k_chunks_per_tap;
    int m_tiles;              // cout / 128 (1 when DUAL)
    int row_tiles;            // ceil(S / slices per tile)  (DUAL: pairs of row tiles)
    int base_off_mode;        // halo kernel: fill the descriptors' matrix-base-offset field (A/B probe)
    int reverse;              // walk the tiles from the last one down (see ConvTcLaunch::reverse)
};

constexpr int cm_stages(bool dual) { return dual ? 3 : 4; }
constexpr int cm_stage_bytes(bool dual) { return dual ? 64 * 128 + 2 * kXTileBytes : 128 * 128 + kXTileBytes; }
constexpr size_t cm_smem_bytes(bool dual) { return 1024 + (size_t)cm_stages(dual) * cm_stage_bytes(dual) + (2 * cm_stages(dual) + 4) * 8 + 16; }

// ---- TMEM loads of N consecutive columns of this thread's lane (no wait inside)
__device__ __forceinline__ void tmem_ld_x2(uint32_t a, float* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=f"(v[0]), "=f"(v[1]) : "r"(a));
}
__device__ __forceinline__ void tmem_ld_x4(uint32_t a, float* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "r"(a));
}
__device__ __forceinline__ void tmem_ld_x8(uint32_t a, float* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "r"(a));
}
__device__ __forceinline__ void tmem_ld_x16(uint32_t a, float* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]),
          "=f"(v[8]), "=f"(v[9]), "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15]) : "r"(a));
}
template <int H> __device__ __forceinline__ void tmem_ld_slice(uint32_t a, float (&v)[H]);
template <> __device__ __forceinline__ void tmem_ld_slice<24>(uint32_t a, float (&v)[24]) { tmem_ld_x16(a, v); tmem_ld_x8(a + 16, v + 16); }
template <> __device__ __forceinline__ void tmem_ld_slice<12>(uint32_t a, float (&v)[12]) { tmem_ld_x8(a, v); tmem_ld_x4(a + 8, v + 8); }
template <> __device__ __forceinline__ void tmem_ld_slice<6>(uint32_t a, float (&v)[6]) { tmem_ld_x4(a, v); tmem_ld_x2(a + 4, v + 4); }

// tcgen05.wait::ld that CARRIES the loaded registers: their consumers cannot be scheduled above the wait
template <int H> __device__ __forceinline__ void tmem_wait_slice(float (&v)[H]);
template <> __device__ __forceinline__ void tmem_wait_slice<6>(float (&v)[6]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;" : "+f"(v[0]), "+f"(v[1]), "+f"(v[2]), "+f"(v[3]), "+f"(v[4]), "+f"(v[5])::"memory");
}
template <> __device__ __forceinline__ void tmem_wait_slice<12>(float (&v)[12]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+f"(v[0]), "+f"(v[1]), "+f"(v[2]), "+f"(v[3]), "+f"(v[4]), "+f"(v[5]), "+f"(v[6]), "+f"(v[7]), "+f"(v[8]), "+f"(v[9]),
                   "+f"(v[10]), "+f"(v[11])::"memory");
}
template <> __device__ __forceinline__ void tmem_wait_slice<24>(float (&v)[24]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+f"(v[0]), "+f"(v[1]), "+f"(v[2]), "+f"(v[3]), "+f"(v[4]), "+f"(v[5]), "+f"(v[6]), "+f"(v[7]), "+f"(v[8]), "+f"(v[9]),
                   "+f"(v[10]), "+f"(v[11]), "+f"(v[12]), "+f"(v[13]), "+f"(v[14]), "+f"(v[15]), "+f"(v[16]), "+f"(v[17]), "+f"(v[18]),
                   "+f"(v[19]), "+f"(v[20]), "+f"(v[21]), "+f"(v[22]), "+f"(v[23])::"memory");
}

__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// ------------------------------------------------------------------ the epilogue (shared by both kernels)
// One accumulator holds, per TMEM lane = output channel, SPT slices of H positions each, slice j starting at column j * HP.
// EW epilogue warps (8, 12 or 16): EW / 4 per lane quarter, the slices dealt round-robin among them.  RES: a residual
// tensor is added last.  COUT == 64 (DUAL): lanes 0-15 / 16-31 of a sub-partition hold the same 16 channels of two row tiles.
// The channel count is a template parameter so that a thread's H outputs (COUT elements apart) get immediate offsets.
template <typename T16, int H, int COUT, int EW, bool RES, int HP, int SPT>
__device__ __forceinline__ void cm_epilogue(const CmParams& p, uint32_t tmem_base, uint64_t* tmem_full, uint64_t* tmem_empty,
                                            int warp, int lane, int tile_first, int tile_step, int num_tiles) {
    constexpr bool DUAL = COUT == 64;
    constexpr int CPG = COUT / 8;                  // channels per GroupNorm group (8 groups)
    constexpr int PARTS = EW / 4;                  // epilogue warps per TMEM lane quarter
    constexpr int MC = DUAL ? 64 : 128;            // output channels per MMA
    constexpr int XT = DUAL ? 2 : 1;               // row tiles per accumulator
    {
        const int q = warp & 3;                                   // TMEM lane quarter this warp may access
        const int part = (warp - 2) >> 2;                         // which share of the tile's slices
        T16* out = reinterpret_cast<T16*>(p.out);
        const T16* res = reinterpret_cast<const T16*>(p.add_res);
        const float* addv = p.add_vec;
        if (addv && p.t_dev) addv += (long long)(*p.t_dev) * COUT;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
        const int cl = DUAL ? q * 16 + (lane & 15) : q * 32 + lane;       // channel inside the M tile
        const int xt = DUAL ? (lane >> 4) : 0;                            // which row tile of the accumulator this lane holds
        constexpr float inv_cnt = 1.0f / (float)(H * CPG);
        constexpr float kLog2e = 1.4426950408889634f;
        int acc = 0; uint32_t acc_phase = 0;
        for (int tile = tile_first; tile < num_tiles; tile += tile_step) {
            const int vt = p.reverse ? num_tiles - 1 - tile : tile;
            const int rt = vt / p.m_tiles, mt = vt - rt * p.m_tiles;
            const int c = mt * MC + cl;
            const float bi = p.bias ? p.bias[c] : 0.f, ga = p.gamma[c], be = p.beta[c], ad = addv ? addv[c] : 0.f;
            const long long s_first = (long long)rt * (SPT * XT) + xt * SPT;      // first slice of this lane's row tile
            // raw residual values of slice j of this lane's row tile (slices past the end read slice 0: never stored)
            constexpr int RH = RES ? H : 1;
            auto res_load = [&](T16 (&r)[RH], int j) {
                const long long s = s_first + j;
                const T16* rp = res + ((s < p.S ? s : 0) * H) * COUT + c;
#pragma unroll
                for (int h = 0; h < RH; ++h) r[h] = rp[h * COUT];
            };
            T16 ra[RH], rb[RH];
            if (RES) {
                // the first slice's residual does not depend on the accumulator: fetch it before the wait, and pull the
                // NEXT tile's residual block (one contiguous range of rows) into L2 a whole tile ahead of its use
                res_load(ra, part);
                const int nt = tile + tile_step;
                const int nvt = p.reverse ? num_tiles - 1 - nt : nt;
                if (nt < num_tiles && (nvt % p.m_tiles) == 0) {
                    const long long row0 = (long long)(nvt / p.m_tiles) * (SPT * H * XT);
                    const char* base = reinterpret_cast<const char*>(res + row0 * COUT);
                    const long long left = (p.S * H - row0) * COUT * 2;           // bytes up to the end of the tensor
                    constexpr int kBlock = SPT * H * XT * COUT * 2;
                    for (int b = ((warp - 2) * 32 + lane) * 128; b < kBlock && b < left; b += EW * 32 * 128)
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(base + b));
                }
            }
            mbar_wait_backoff(&tmem_full[acc], acc_phase);
            tc_fence_after();
            const uint32_t taddr = lane_addr + (uint32_t)(acc * kAccStride);

            auto compute = [&](float (&v)[H], T16 (&r16)[RH], int j) {
                const long long s = s_first + j;
                const bool valid = s < p.S;
                T16* op = out + (s * H) * COUT + c;
                // Packed fp32 (add.f32x2 / fma.f32x2 = two IEEE operations per instruction at half the issue rate, the same
                // arithmetic throughput): the statistics, x and the ex2 argument of two adjacent positions take one instruction
                // each (their inputs sit in adjacent registers straight out of tcgen05.ld); everything downstream of the scalar
                // MUFU results stays scalar (re-packing them costs the moves it would save).  Same operations in the same order
                // as the scalar form, which H = 24 with a residual keeps (two more live register pairs would spill there).
                constexpr bool PACKED = !(H == 24 && RES);
                float a0, a1, b0, b1;
                unsigned long long v2[PACKED ? H / 2 : 1];
                if (PACKED) {
                    unsigned long long a2 = 0ull, b2 = 0ull;
#pragma unroll
                    for (int h = 0; h < H; h += 2) {
                        v2[PACKED ? h >> 1 : 0] = f32x2_pack(v[h], v[h + 1]);
                        a2 = f32x2_add(a2, v2[PACKED ? h >> 1 : 0]);
                        b2 = f32x2_fma(v2[PACKED ? h >> 1 : 0], v2[PACKED ? h >> 1 : 0], b2);
                    }
                    f32x2_unpack(a2, a0, a1);
                    f32x2_unpack(b2, b0, b1);
                } else {
                    a0 = a1 = b0 = b1 = 0.f;
#pragma unroll
                    for (int h = 0; h < H; h += 2) {
                        a0 += v[h]; b0 = fmaf(v[h], v[h], b0);
                        a1 += v[h + 1]; b1 = fmaf(v[h + 1], v[h + 1], b1);
                    }
                }
                // per-thread sums (even / odd positions, combined in a fixed order), conv bias folded in
                const float t1 = a0 + a1, t2 = b0 + b1;
                float s1 = fmaf((float)H, bi, t1);
                float s2 = fmaf(bi, fmaf((float)H, bi, 2.0f * t1), t2);   // sum (v + b)^2 = sum v^2 + b (2 sum v + H b)
                // ... over the CPG adjacent lanes of the channel group: fixed shuffle tree, independent of the slice's place
#pragma unroll
                for (int o = 1; o < CPG; o <<= 1) {
                    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
                    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
                }
                const float mean = s1 * inv_cnt;
                const float rstd = rsqrtf(fmaxf(fmaf(-mean, mean, s2 * inv_cnt), 0.f) + 1e-5f);
                const float sc = rstd * ga;                               // x = (v + b - mean) rstd gamma + beta = v sc + sh
                const float sh = fmaf(bi - mean, sc, be);
                const float scl = sc * kLog2e, shl = sh * kLog2e;
                const unsigned long long sc_2 = f32x2_pack(sc, sc), sh_2 = f32x2_pack(sh, sh);
                const unsigned long long scl_2 = f32x2_pack(scl, scl), shl_2 = f32x2_pack(shl, shl);
#pragma unroll
                for (int h = 0; h < H; h += 2) {
                    // mish(x) = x (1 - 2 / (w (w + 2) + 2)),  w = e^x
                    float xx[2], tt[2];
                    if (PACKED) {
                        f32x2_unpack(f32x2_fma(v2[PACKED ? h >> 1 : 0], sc_2, sh_2), xx[0], xx[1]);
                        f32x2_unpack(f32x2_fma(v2[PACKED ? h >> 1 : 0], scl_2, shl_2), tt[0], tt[1]);
                    } else {
                        xx[0] = fmaf(v[h], sc, sh); xx[1] = fmaf(v[h + 1], sc, sh);
                        tt[0] = fmaf(v[h], scl, shl); tt[1] = fmaf(v[h + 1], scl, shl);
                    }
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        const float w = ex2_approx(tt[u]);
                        const float r = rcp_approx(fmaf(w, w + 2.0f, 2.0f));
                        // (a layer adds either the time-embedding row or a residual, so with RES the residual rides in the FMA)
                        const float y = fmaf(xx[u], fmaf(r, -2.0f, 1.0f), RES ? to_f32<T16>(r16[RES ? h + u : 0]) : ad);
                        if (valid) op[(h + u) * COUT] = from_f32<T16>(y);
                    }
                }
            };

            // slices part, part + PARTS, ...: the TMEM load (and the residual values) of the next one are in flight while
            // this one is processed
            float va[H], vb[H];
            tmem_ld_slice<H>(taddr + part * HP, va);
            for (int j = part; j < SPT; j += 2 * PARTS) {
                const int j1 = j + PARTS, j2 = j + 2 * PARTS;
                tmem_wait_slice<H>(va);
                if (j1 < SPT) {
                    tmem_ld_slice<H>(taddr + j1 * HP, vb);
                    if (RES) res_load(rb, j1);
                }
                compute(va, ra, j);
                if (j1 < SPT) {
                    tmem_wait_slice<H>(vb);
                    if (j2 < SPT) {
                        tmem_ld_slice<H>(taddr + j2 * HP, va);
                        if (RES) res_load(ra, j2);
                    }
                    compute(vb, rb, j1);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }
}

// ------------------------------------------------------------------ the kernel
// Per-tap operand loads: every (tap, 64-channel block) is one pipeline stage holding the weight tile and the row tile
// shifted by the tap (zero-filled outside the slice by TMA).
template <typename T16, int H, int COUT, int EW, bool RES>
__global__ void __launch_bounds__(64 + 32 * EW, 1)
conv_tc_cm_kernel(const __grid_constant__ CUtensorMap map_x0, const __grid_constant__ CUtensorMap map_x1,
                  const __grid_constant__ CUtensorMap map_w, const CmParams p) {
    constexpr bool DUAL = COUT == 64;
    constexpr int CPG = COUT / 8;                  // channels per GroupNorm group (8 groups)
    constexpr int kStages = cm_stages(DUAL);
    constexpr int kStageBytes = cm_stage_bytes(DUAL);
    constexpr int kWTileBytes = DUAL ? 64 * 128 : 128 * 128;
    constexpr int SPT = kRows / H;                 // slices per row tile
    constexpr int PARTS = EW / 4;                  // epilogue warps per TMEM lane quarter
    constexpr int MC = DUAL ? 64 : 128;            // output channels per MMA
    constexpr int XT = DUAL ? 2 : 1;               // row tiles per accumulator
    static_assert(kRows % H == 0, "a tile holds whole slices");

    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* tiles = smem;                                             // kStages x (W | X [| X])
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
    uint64_t* empty_bar = full_bar + kStages;
    uint64_t* tmem_full = empty_bar + kStages;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_tiles = p.row_tiles * p.m_tiles;
    const int tile_first = blockIdx.x, tile_step = gridDim.x;
    const int k_chunks = p.taps * p.k_chunks_per_tap;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], EW); }
        fence_barrier_init();
        tma_prefetch_desc(&map_x0);
        tma_prefetch_desc(&map_w);
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    // everything above depends on nothing; the previous kernel of the chain must be complete before any global access
    pdl_wait();
    pdl_trigger();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // =============================== TMA producer ===============================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int tile = tile_first; tile < num_tiles; tile += tile_step) {
                const int vt = p.reverse ? num_tiles - 1 - tile : tile;
                const int rt = vt / p.m_tiles, mt = vt - rt * p.m_tiles;
                const int s0 = rt * (SPT * XT);
                for (int tap = 0; tap < p.taps; ++tap) {
                    for (int kc = 0; kc < p.k_chunks_per_tap; ++kc) {
                        mbar_wait_backoff(&empty_bar[stage], phase ^ 1);
                        uint8_t* w_dst = tiles + stage * kStageBytes;
                        uint8_t* x_dst = w_dst + kWTileBytes;
                        mbar_expect_tx(&full_bar[stage], (uint32_t)(kWTileBytes + XT * kXTileBytes));
                        const int ci = kc * kBlockK;
                        const bool second = ci >= p.c0;
                        const CUtensorMap* mx = second ? &map_x1 : &map_x0;
                        const int xc = second ? ci - p.c0 : ci;
                        tma_load_2d(&map_w, &full_bar[stage], w_dst, ci, tap * COUT + mt * MC);
                        tma_load_3d(mx, &full_bar[stage], x_dst, xc, tap - p.taps / 2, s0);
                        if (DUAL) tma_load_3d(mx, &full_bar[stage], x_dst + kXTileBytes, xc, tap - p.taps / 2, s0 + SPT);
                        if (++stage == kStages) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // =============================== MMA issuer ===============================
        if (lane == 0) {
            constexpr uint32_t idesc = (1u << 4) | (Fmt<T16>::kind << 7) | (Fmt<T16>::kind << 10) |
                                       ((uint32_t)(kRows >> 3) << 17) | ((uint32_t)(MC >> 4) << 24);
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            for (int tile = tile_first; tile < num_tiles; tile += tile_step) {
                mbar_wait_backoff(&tmem_empty[acc], acc_phase ^ 1);       // the epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * kAccStride);
                for (int kc = 0; kc < k_chunks; ++kc) {
                    mbar_wait_backoff(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t w_addr = smem_u32(tiles + stage * kStageBytes);
                    const uint32_t x_addr = w_addr + kWTileBytes;
#pragma unroll
                    for (int k = 0; k < kBlockK / 16; ++k) {
                        const uint32_t accumulate = (kc | k) ? 1u : 0u;
                        tc_mma_f16(d_tmem, umma_smem_desc(w_addr + k * 32), umma_smem_desc(x_addr + k * 32), idesc, accumulate);
                        if (DUAL)
                            tc_mma_f16(d_tmem + (16u << 16), umma_smem_desc(w_addr + k * 32),
                                       umma_smem_desc(x_addr + kXTileBytes + k * 32), idesc, accumulate);
                    }
                    tc_commit(&empty_bar[stage]);                         // frees the smem stage
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
                tc_commit(&tmem_full[acc]);                               // accumulator ready for the epilogue
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
        __syncwarp();
    } else {
        cm_epilogue<T16, H, COUT, EW, RES, H, SPT>(p, tmem_base, tmem_full, tmem_empty, warp, lane, tile_first, tile_step, num_tiles);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// ------------------------------------------------------------------ the halo kernel (H = 24 / 12)
// The per-tap kernel above streams every row tile through the L2 -> SM fabric five times (once per tap) and is bound by
// exactly that (ncu: 1.15 GB of fabric reads per 128->128 launch at the fabric's ~11.5 TB/s = the kernel's 100 us).  Here a
// 64-channel block of the row tile is loaded ONCE, with a zero halo of two positions on both sides of every slice (TMA box
// of H + 4 positions starting at position -2: the out-of-range rows are zero-filled), and the five taps are five MMAs whose
// B descriptors start 0..4 rows (128 B each) into that tile.  Slice j then occupies accumulator columns j (H + 4) .. + H - 1;
// the 4 columns between slices are junk and never read.  Two rings: row-tile blocks (X) and weight tiles (W, one per tap).
template <int H> struct HaloGeom;
template <> struct HaloGeom<24> { static constexpr int SPT = 9, N = 248; };    // 9 x 28 = 252 rows in shared memory
template <> struct HaloGeom<12> { static constexpr int SPT = 16, N = 256; };   // 16 x 16 = 256 rows (taps 1-4 read up to 4 rows
                                                                               // of the NEXT buffer into junk columns only)
constexpr int kXHaloBytes = 32 * 1024;
constexpr int halo_nx(bool dual) { return dual ? 2 : 3; }       // X ring depth (a dual slot holds two row tiles)
constexpr int halo_nw(bool dual) { return dual ? 8 : 6; }       // W ring depth
constexpr size_t halo_smem_bytes(bool dual) {
    return 1024 + (size_t)halo_nx(dual) * (dual ? 2 : 1) * kXHaloBytes + (size_t)halo_nw(dual) * (dual ? 64 : 128) * 128 + 1024 +
           (2 * halo_nx(dual) + 2 * halo_nw(dual) + 4) * 8 + 16;
}

// SW128 K-major descriptor whose start is `row_off` rows into a 1024-byte-aligned tile
__device__ __forceinline__ uint64_t umma_smem_desc_rows(uint32_t tile_addr, int row_off, int k_off_bytes, uint32_t base_off_mode) {
    uint64_t d = umma_smem_desc(tile_addr + row_off * 128 + k_off_bytes);
    if (base_off_mode) d |= (uint64_t)(row_off & 7) << 49;         // matrix base offset = (start address >> 7) & 7
    return d;
}

template <typename T16, int H, int COUT, int EW, bool RES>
__global__ void __launch_bounds__(64 + 32 * EW, 1)
conv_tc_cm_halo_kernel(const __grid_constant__ CUtensorMap map_x0, const __grid_constant__ CUtensorMap map_x1,
                       const __grid_constant__ CUtensorMap map_w, const CmParams p) {
    constexpr bool DUAL = COUT == 64;
    constexpr int SPT = HaloGeom<H>::SPT, N = HaloGeom<H>::N, HP = H + 4;
    constexpr int NX = halo_nx(DUAL), NW = halo_nw(DUAL);
    constexpr int MC = DUAL ? 64 : 128, XT = DUAL ? 2 : 1;
    constexpr int kWTileBytes = MC * 128;
    constexpr int kXSlotBytes = XT * kXHaloBytes;
    constexpr uint32_t kXBytes = (uint32_t)(SPT * HP * 128);       // bytes one TMA box delivers

    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* x_ring = smem;
    uint8_t* w_ring = smem + NX * kXSlotBytes;
    uint64_t* x_full = reinterpret_cast<uint64_t*>(w_ring + NW * kWTileBytes + 1024);   // (1 KB of slack: the 4-row overrun)
    uint64_t* x_empty = x_full + NX;
    uint64_t* w_full = x_empty + NX;
    uint64_t* w_empty = w_full + NW;
    uint64_t* tmem_full = w_empty + NW;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_tiles = p.row_tiles * p.m_tiles;
    const int tile_first = blockIdx.x, tile_step = gridDim.x;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NX; ++s) { mbar_init(&x_full[s], 1); mbar_init(&x_empty[s], 1); }
        for (int s = 0; s < NW; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], EW); }
        fence_barrier_init();
        tma_prefetch_desc(&map_x0);
        tma_prefetch_desc(&map_w);
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    pdl_wait();
    pdl_trigger();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // =============================== TMA producer ===============================
        if (lane == 0) {
            int xs = 0, ws = 0; uint32_t xph = 0, wph = 0;
            for (int tile = tile_first; tile < num_tiles; tile += tile_step) {
                const int vt = p.reverse ? num_tiles - 1 - tile : tile;
                const int rt = vt / p.m_tiles, mt = vt - rt * p.m_tiles;
                const int s0 = rt * (SPT * XT);
                for (int kc = 0; kc < p.k_chunks_per_tap; ++kc) {
                    const int ci = kc * kBlockK;
                    const bool second = ci >= p.c0;
                    const CUtensorMap* mx = second ? &map_x1 : &map_x0;
                    const int xc = second ? ci - p.c0 : ci;
                    mbar_wait_backoff(&x_empty[xs], xph ^ 1);
                    uint8_t* x_dst = x_ring + xs * kXSlotBytes;
                    mbar_expect_tx(&x_full[xs], kXBytes * XT);
                    tma_load_3d(mx, &x_full[xs], x_dst, xc, -2, s0);
                    if (DUAL) tma_load_3d(mx, &x_full[xs], x_dst + kXHaloBytes, xc, -2, s0 + SPT);
                    if (++xs == NX) { xs = 0; xph ^= 1; }
                    for (int tap = 0; tap < 5; ++tap) {
                        mbar_wait_backoff(&w_empty[ws], wph ^ 1);
                        mbar_expect_tx(&w_full[ws], (uint32_t)kWTileBytes);
                        tma_load_2d(&map_w, &w_full[ws], w_ring + ws * kWTileBytes, ci, tap * COUT + mt * MC);
                        if (++ws == NW) { ws = 0; wph ^= 1; }
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // =============================== MMA issuer ===============================
        if (lane == 0) {
            constexpr uint32_t idesc = (1u << 4) | (Fmt<T16>::kind << 7) | (Fmt<T16>::kind << 10) |
                                       ((uint32_t)(N >> 3) << 17) | ((uint32_t)(MC >> 4) << 24);
            const uint32_t bo = (uint32_t)p.base_off_mode;
            int xs = 0, ws = 0; uint32_t xph = 0, wph = 0;
            int acc = 0; uint32_t acc_phase = 0;
            for (int tile = tile_first; tile < num_tiles; tile += tile_step) {
                mbar_wait_backoff(&tmem_empty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * kAccStride);
                for (int kc = 0; kc < p.k_chunks_per_tap; ++kc) {
                    mbar_wait_backoff(&x_full[xs], xph);
                    const uint32_t x_addr = smem_u32(x_ring + xs * kXSlotBytes);
                    for (int tap = 0; tap < 5; ++tap) {
                        mbar_wait_backoff(&w_full[ws], wph);
                        tc_fence_after();
                        const uint32_t w_addr = smem_u32(w_ring + ws * kWTileBytes);
#pragma unroll
                        for (int k = 0; k < kBlockK / 16; ++k) {
                            const uint32_t accumulate = (kc | tap | k) ? 1u : 0u;
                            tc_mma_f16(d_tmem, umma_smem_desc(w_addr + k * 32), umma_smem_desc_rows(x_addr, tap, k * 32, bo), idesc, accumulate);
                            if (DUAL)
                                tc_mma_f16(d_tmem + (16u << 16), umma_smem_desc(w_addr + k * 32),
                                           umma_smem_desc_rows(x_addr + kXHaloBytes, tap, k * 32, bo), idesc, accumulate);
                        }
                        tc_commit(&w_empty[ws]);
                        if (++ws == NW) { ws = 0; wph ^= 1; }
                    }
                    tc_commit(&x_empty[xs]);
                    if (++xs == NX) { xs = 0; xph ^= 1; }
                }
                tc_commit(&tmem_full[acc]);
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
        __syncwarp();
    } else {
        cm_epilogue<T16, H, COUT, EW, RES, HP, SPT>(p, tmem_base, tmem_full, tmem_empty, warp, lane, tile_first, tile_step, num_tiles);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

template <typename T16, int H, int COUT, int EW, bool RES>
int launch_cm_halo(const CUtensorMap& x0, const CUtensorMap& x1, const CUtensorMap& w, const CmParams& p, cudaStream_t st) {
    auto kern = conv_tc_cm_halo_kernel<T16, H, COUT, EW, RES>;
    constexpr size_t smem = halo_smem_bytes(COUT == 64);
    static_assert(smem <= 227 * 1024, "shared memory budget");
    static DeviceOnce once;
    if (once.first_time()) CINDM_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int tiles = p.row_tiles * p.m_tiles;
    const int grid = tiles < num_sms() ? tiles : num_sms();
    CINDM_CHECK_CUDA(launch_chain(kern, dim3(grid), dim3(64 + 32 * EW), smem, st, x0, x1, w, p));
    CINDM_CHECK_LAUNCH();
    return 0;
}

template <typename T16, int H, int COUT, int EW, bool RES>
int launch_cm(const CUtensorMap& x0, const CUtensorMap& x1, const CUtensorMap& w, const CmParams& p, cudaStream_t st) {
    auto kern = conv_tc_cm_kernel<T16, H, COUT, EW, RES>;
    constexpr size_t smem = cm_smem_bytes(COUT == 64);
    static DeviceOnce once;
    if (once.first_time()) CINDM_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int tiles = p.row_tiles * p.m_tiles;
    const int grid = tiles < num_sms() ? tiles : num_sms();
    CINDM_CHECK_CUDA(launch_chain(kern, dim3(grid), dim3(64 + 32 * EW), smem, st, x0, x1, w, p));
    CINDM_CHECK_LAUNCH();
    return 0;
}

// which channel counts take this kernel: bit 0 -> 128, bit 1 -> 64, bit 2 -> 256 / 512.  CINDM_CONV_CM=<mask> (A/B runs)
int cm_mask() {
    static int mask = -1;
    if (mask < 0) { const char* e = getenv("CINDM_CONV_CM"); mask = e ? atoi(e) : 7; }
    return mask;
}
int cm_epilogue_warps() {
    static int ew = -1;
    if (ew < 0) { const char* e = getenv("CINDM_CONV_CM_EW"); ew = (e && atoi(e) == 12) ? 12 : 16; }
    return ew;
}
// single row-tile load per K block with per-slice zero halos (H = 24 / 12).  CINDM_CONV_CM_HALO: 0 = never, 1 = wherever it is
// built, 3 (default) = where it measured faster (the 64-channel layers that also read a residual), 2 = like 1 but with the
// descriptors' base-offset field filled in (a probe of the swizzle convention: WRONG results, kept for the record)
int cm_halo_mode() {
    static int mode = -1;
    if (mode < 0) { const char* e = getenv("CINDM_CONV_CM_HALO"); mode = e ? atoi(e) : 3; }
    return mode;
}

template <typename T16, int H, int COUT>
int dispatch_cm2(bool res, bool halo, const CUtensorMap& x0, const CUtensorMap& x1, const CUtensorMap& w, const CmParams& p, cudaStream_t st) {
    // 16 epilogue warps (4 per SM sub-partition) measured best or equal on every shape except H = 24 with a residual, whose two
    // raw-residual buffers do not fit the 96 registers of an 18-warp CTA (12 warps there).  CINDM_CONV_CM_EW=12 overrides.
    const int ew = (H == 24 && res) ? 12 : cm_epilogue_warps();
    if constexpr (H == 24 || H == 12) {
        if (halo) {
            if (res) return ew == 12 ? launch_cm_halo<T16, H, COUT, 12, true>(x0, x1, w, p, st) : launch_cm_halo<T16, H, COUT, 16, true>(x0, x1, w, p, st);
            return ew == 12 ? launch_cm_halo<T16, H, COUT, 12, false>(x0, x1, w, p, st) : launch_cm_halo<T16, H, COUT, 16, false>(x0, x1, w, p, st);
        }
    }
    if (res) return ew == 12 ? launch_cm<T16, H, COUT, 12, true>(x0, x1, w, p, st) : launch_cm<T16, H, COUT, 16, true>(x0, x1, w, p, st);
    return ew == 12 ? launch_cm<T16, H, COUT, 12, false>(x0, x1, w, p, st) : launch_cm<T16, H, COUT, 16, false>(x0, x1, w, p, st);
}

template <typename T16>
int dispatch_cm(int H, int cout, bool res, bool halo, const CUtensorMap& x0, const CUtensorMap& x1, const CUtensorMap& w, const CmParams& p,
                cudaStream_t st) {
    if (cout == 64 && H == 24) return dispatch_cm2<T16, 24, 64>(res, halo, x0, x1, w, p, st);
    if (cout == 64 && H == 12) return dispatch_cm2<T16, 12, 64>(res, halo, x0, x1, w, p, st);
    if (cout == 128 && H == 12) return dispatch_cm2<T16, 12, 128>(res, halo, x0, x1, w, p, st);
    if (cout == 128 && H == 6) return dispatch_cm2<T16, 6, 128>(res, halo, x0, x1, w, p, st);
    if (cout == 256 && H == 6) return dispatch_cm2<T16, 6, 256>(res, halo, x0, x1, w, p, st);
    return fail(-2, "conv_tc_cm: no instance for this (channels, H)");
}

}  // namespace

bool conv_tc_cm_eligible(const ConvTcLaunch& a) {
    if (a.epilogue != EPI_GN_MISH || a.mode != TC_SAME || !a.w || a.w->taps != 5) return false;
    if (a.add_res && a.add_vec) return false;      // (no layer of the U-Net adds both; the row-major kernel handles it)
    const int cout = a.w->cout, H = a.H, m = cm_mask();
    if (cout == 128 && (H == 12 || H == 6)) return m & 1;
    if (cout == 64 && (H == 24 || H == 12)) return m & 2;
    // (256 channels = two M tiles that both stream the row tile: with a long K loop, cin = 512, the row-major CTA-pair kernel wins)
    if (cout == 256 && H == 6 && a.w->cin <= 256) return m & 4;
    return false;
}

int launch_conv_tc_cm(const ConvTcLaunch& a, cudaStream_t st) {
    const ConvW& w = *a.w;
    const bool dual = w.cout == 64;
    CmParams p;
    p.bias = w.bias; p.gamma = a.gn->gamma; p.beta = a.gn->beta;
    p.add_vec = a.add_vec; p.t_dev = a.t_dev; p.add_res = a.add_res; p.out = a.out;
    p.S = a.S; p.cout = w.cout; p.c0 = a.c0; p.c1 = a.in1 ? a.c1 : 0; p.taps = w.taps;
    p.k_chunks_per_tap = w.cin / kBlockK;
    const int hm = cm_halo_mode();
    const bool halo = (a.H == 24 || a.H == 12) && hm != 0 && (hm != 3 || (w.cout == 64 && a.add_res));
    const int spt = halo ? (a.H == 24 ? HaloGeom<24>::SPT : HaloGeom<12>::SPT) : kRows / a.H;
    p.base_off_mode = hm == 2 ? 1 : 0;
    p.reverse = a.reverse;
    p.m_tiles = dual ? 1 : w.cout / 128;
    p.row_tiles = (int)((a.S + spt * (dual ? 2 : 1) - 1) / (spt * (dual ? 2 : 1)));
    char tag[96];
    snprintf(tag, sizeof tag, "conv_tc same H%d %d->%d k%d gn%s", a.H, w.cin, w.cout, w.taps, a.add_res ? "+res" : "");
    KernelTimer kt(tag, st, 2.0 * (double)a.S * (double)(5 * a.H - 6) * w.cin * w.cout);
    CUtensorMap m0, m1, mw;
    const int box_rows = halo ? a.H + 4 : a.H;
    CINDM_TRY(encode_act_map(&m0, a.in0, a.prec, a.S, a.H, a.c0, spt, box_rows, 1));
    if (a.in1) CINDM_TRY(encode_act_map(&m1, a.in1, a.prec, a.S, a.H, a.c1, spt, box_rows, 1));
    else m1 = m0;
    CINDM_TRY(encode_weight_map(&mw, w.w16[a.prec], a.prec, w.taps * w.cout, w.cin, dual ? 64 : 128));
    if (a.prec == PREC_F16) return dispatch_cm<__half>(a.H, w.cout, a.add_res != nullptr, halo, m0, m1, mw, p, st);
    return dispatch_cm<__nv_bfloat16>(a.H, w.cout, a.add_res != nullptr, halo, m0, m1, mw, p, st);
}

}  // namespace cindm
