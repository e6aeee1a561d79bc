// Temporal convolution as an implicit GEMM on the 5th-generation tensor cores (sm_100a).
//
//   D[row][co] = sum_{tap, ci} A_tap[row][ci] * W[tap][co][ci]        row = (slice, position), fp32 accumulate
//
// * A operand: the channels-last activation tensor [S][H][C] is read through a 3-D TMA descriptor
//   with box (64 channels, H positions, 128/H slices) whose H coordinate starts at tap - pad: the
//   hardware zero-fills the out-of-range rows, which is exactly the conv's zero padding at slice
//   boundaries, so a tile's M rows are whole slices and each tap is one more K block of the GEMM.
// * B operand: repacked weights [tap][cout][cin] (K-major) through a 2-D descriptor.
// * Both land in 128-byte-swizzled shared memory; one elected thread issues tcgen05.mma
//   (cta_group::1, kind::f16, M=128, N=N_TILE, K=16) into a double-buffered TMEM accumulator.
// * Epilogue warps read TMEM with tcgen05.ld and apply, fused: conv bias, GroupNorm(8) over
//   (C/8 channels x H positions) of each slice (a tile owns whole slices, so the statistics are
//   CTA-local), Mish, then the time-embedding bias or the residual, and store 16-bit activations.
// Warp roles (320 threads): warp 0 TMA producer, warp 1 MMA issuer + TMEM allocator, warps 2-9 epilogue
// (two warps per TMEM lane quarter, each owning half of the tile's columns).
#include "conv_tc.h"
#include "tc_common.cuh"

namespace cindm {

namespace {

// pipeline depth per N tile: as deep as the 227 KB of shared memory allow (A 16 KB + B N*128 B per stage,
// plus the output staging slabs of the N <= 128 kernels and ~20 KB of vectors / partials / barriers)
// OCC == 2 (two CTAs per SM, N <= 128): each CTA gets half of the shared memory and of the tensor memory, so that while
// one CTA's epilogue warps sit at a barrier or wait for their accumulator the other CTA's can issue.
constexpr int stages_for(int n_tile, int cg, int occ = 1) {
    return occ == 2 ? (n_tile == 128 ? 2 : 3)
                    : (cg == 2 ? (n_tile == 256 ? 6 : 7) : (n_tile == 256 ? 4 : (n_tile == 192 ? 5 : (n_tile == 128 ? 5 : 6))));
}
// floats per channel vector in shared memory, and how many vectors (bias | gamma | beta | add): the GroupNorm kernels with
// N <= 128 always cover all output channels with one N tile; the plain kernels only need the bias (up to 512 channels)
constexpr int vec_cap_for(int n_tile, int epi, int occ) { return occ == 2 ? (epi == EPI_BIAS ? 512 : n_tile) : 512; }
constexpr int vec_count_for(int epi, int occ) { return (occ == 2 && epi == EPI_BIAS) ? 1 : 4; }
constexpr int part_floats_for(int epi, int occ) { return (occ == 2 && epi == EPI_BIAS) ? 0 : 2048 + 704; }
// warp 0 TMA, warp 1 MMA, then the epilogue warps: 4*PARTS of them, PARTS per TMEM lane quarter, each owning 1/PARTS of
// the tile's columns.  PARTS = 2 everywhere: 4 (16 epilogue warps) was measured slower (96-register cap, spills).
constexpr int epi_warps_for(int /*n_tile*/, int /*epi*/) { return 8; }
constexpr int threads_for(int n_tile, int epi) { return 64 + 32 * epi_warps_for(n_tile, epi); }

struct TcParams {
    const float* bias;        // [cout] or null
    const float* gamma;       // [cout]  (GN)
    const float* beta;        // [cout]  (GN)
    const float* add_vec;     // [cout] or [timesteps][cout] when t_dev != null
    const int* t_dev;
    const void* add_res;      // [S][H][cout] 16-bit or null
    void* out;                // [S][H][cout] 16-bit
    long long S;
    int H;                    // rows per slice inside an M tile (output positions of this GEMM)
    int cout, c0, c1, cin, taps;
    int kpp;                  // Toeplitz: 64-channel K chunks per input position (= cin / 64)
    int tap_hoff[5];          // H coordinate where the A box of each tap starts (out-of-range rows read as zero)
    int tap_wrow[5];          // first row of each tap's block in the [taps*cout][cin] weight matrix
    int out_mul, out_add;     // output row = tile row * out_mul + out_add (2, parity for the transposed conv)
    int tma_out;              // stage the output tile in shared memory and write it with TMA (N_TILE <= 128 kernels)
    int slices_per_tile, rows_used, m_tiles, n_tiles, k_chunks_per_tap;
    int reverse;              // walk the (pair-)tiles from the last one down (see ConvTcLaunch::reverse)
};

// ------------------------------------------------------------------ the kernel
template <typename T16, int N_TILE, int CPG, int EPI, int CG, int OCC>
__global__ void __launch_bounds__(threads_for(N_TILE, EPI), OCC)
conv_tc_kernel(const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_a1,
               const __grid_constant__ CUtensorMap map_b, const __grid_constant__ CUtensorMap map_out, const TcParams p) {
    // CG == 2: two CTAs of a cluster form one 256 x N_TILE MMA; each stages its own 128 A rows and N_TILE/2 rows of B
    static_assert(OCC == 1 || (CG == 1 && N_TILE <= 128 && EPI != EPI_GN_MISH_T3), "two CTAs per SM: single-CTA N <= 128 kernels only");
    constexpr int kStages = stages_for(N_TILE, CG, OCC);
    constexpr int kBTileBytes = N_TILE * 128 / CG;
    constexpr int kStageBytes = kATileBytes + kBTileBytes;
    constexpr int ACC_STRIDE = (N_TILE == 192) ? 256 : N_TILE;        // TMEM columns between the two accumulators
    constexpr bool HAS_GN = (EPI != EPI_BIAS);

    // dynamic shared memory starts 1024-byte aligned (required by the 128-byte swizzle); using the array directly
    // (no integer round-trip) keeps every access in the shared address space (LDS/STS instead of generic LD/ST)
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* tiles = smem;                                             // kStages x (A | B)
    constexpr int kStagingBytes = (N_TILE <= 128) ? (N_TILE / 64) * kATileBytes : 0;   // [N_TILE/64 slabs][128 rows][128 B]
    uint8_t* staging = smem + kStages * kStageBytes;                   // 1024-byte aligned (stage sizes are multiples of 1 KB)
    constexpr int VC = vec_cap_for(N_TILE, EPI, OCC), VN = vec_count_for(EPI, OCC);
    float* vec_bias = reinterpret_cast<float*>(staging + kStagingBytes);
    float* vec_gamma = vec_bias + (VN == 4 ? VC : 0);                  // (the plain kernels at OCC == 2 keep the bias only)
    float* vec_beta = vec_gamma + (VN == 4 ? VC : 0);
    float* vec_add = vec_beta + (VN == 4 ? VC : 0);
    float* part = vec_bias + VN * VC;                                  // GroupNorm partial sums [128][8] float2 + totals [43][8] float2
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(part + part_floats_for(EPI, OCC));
    uint64_t* empty_bar = full_bar + kStages;
    uint64_t* tmem_full = empty_bar + kStages;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t cta_rank = (CG == 2) ? cluster_ctarank() : 0u;
    const bool leader = cta_rank == 0;
    // tiles are handed out per cluster: pair-tile t = (m_pair, n_tile); this CTA works on m_tile = m_pair * CG + rank
    // the generic GroupNorm kernels with N <= 128 always cover all output channels with ONE N tile: knowing it at compile time
    // removes four integer divisions per tile from every role's loop
    constexpr bool ONE_N = (EPI == EPI_GN_MISH) && N_TILE <= 128;
    const int n_tiles = ONE_N ? 1 : p.n_tiles;
    const int num_tiles = ((p.m_tiles + CG - 1) / CG) * n_tiles;
    const int tile_first = blockIdx.x / CG, tile_step = gridDim.x / CG;
    const int k_chunks = p.taps * p.k_chunks_per_tap;
    auto vtile = [&](int t) { return p.reverse ? num_tiles - 1 - t : t; };      // position in the walk -> tile

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], epi_warps_for(N_TILE, EPI) * CG); }
        fence_barrier_init();
        tma_prefetch_desc(&map_a0);
        tma_prefetch_desc(&map_b);
    }
    constexpr uint32_t kTmemCols = OCC == 2 ? 2 * ACC_STRIDE : 512;    // (a power of two >= 32: 128 or 256 at OCC == 2)
    if (warp == 1) { if (CG == 2) tmem_alloc_pair(tmem_slot, kTmemCols); else tmem_alloc(tmem_slot, kTmemCols); }
    // everything above depends on nothing; the previous kernel of the chain must be complete before any global access
    pdl_wait();
    pdl_trigger();
    // per-layer channel vectors -> smem (epilogue broadcast reads)
    {
        const float* addv = p.add_vec;
        if (addv && p.t_dev) addv += (long long)(*p.t_dev) * p.cout;
        for (int c = threadIdx.x; c < p.cout; c += blockDim.x) {
            vec_bias[c] = p.bias ? p.bias[c] : 0.f;
            if (HAS_GN) { vec_gamma[c] = p.gamma[c]; vec_beta[c] = p.beta[c]; }
            if (VN == 4) vec_add[c] = addv ? addv[c] : 0.f;
        }
    }
    tc_fence_before();
    if (CG == 2) cluster_sync_all(); else __syncthreads();      // barrier inits / TMEM allocation visible to the whole pair
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // =============================== TMA producer ===============================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int tile = tile_first; tile < num_tiles; tile += tile_step) {
                const int vt = vtile(tile);
                const int m_pair = vt / n_tiles, n_tile = vt - m_pair * n_tiles;
                const int m_tile = m_pair * CG + (int)cta_rank;
                const int s0 = m_tile * p.slices_per_tile;
                const uint32_t tx_bytes = (uint32_t)(p.rows_used * 128 + kBTileBytes) * CG;   // leader counts both CTAs' bytes
                const int b_row0 = n_tile * N_TILE + (int)cta_rank * (N_TILE / CG);
                for (int tap = 0; tap < p.taps; ++tap) {
                    for (int kc = 0; kc < p.k_chunks_per_tap; ++kc) {
                        mbar_wait_backoff(&empty_bar[stage], phase ^ 1);
                        uint8_t* a_dst = tiles + stage * kStageBytes;
                        uint8_t* b_dst = a_dst + kATileBytes;
                        if (leader) mbar_expect_tx(&full_bar[stage], tx_bytes);
                        // A coordinates (channel, H start, first slice) and B coordinates (k, weight row) of this K block
                        int a_c, a_h, b_k, b_r;
                        bool second;
                        if (EPI == EPI_GN_MISH_T3) {
                            // K index = (input position q, channel ci); the A matrix is the [S][3*C] view of the tensor
                            const int qpos = kc / p.kpp, ci = (kc - qpos * p.kpp) * kBlockK;
                            second = ci >= p.c0;
                            a_c = second ? qpos * p.c1 + ci - p.c0 : qpos * p.c0 + ci;
                            a_h = 0; b_k = kc * kBlockK; b_r = b_row0;
                        } else {
                            const int ci = kc * kBlockK;
                            second = ci >= p.c0;
                            a_c = second ? ci - p.c0 : ci;
                            a_h = p.tap_hoff[tap]; b_k = ci; b_r = p.tap_wrow[tap] + b_row0;
                        }
                        if (CG == 2) {
                            tma_load_3d_pair(second ? &map_a1 : &map_a0, &full_bar[stage], a_dst, a_c, a_h, s0);
                            tma_load_2d_pair(&map_b, &full_bar[stage], b_dst, b_k, b_r);
                        } else {
                            tma_load_3d(second ? &map_a1 : &map_a0, &full_bar[stage], a_dst, a_c, a_h, s0);
                            tma_load_2d(&map_b, &full_bar[stage], b_dst, b_k, b_r);
                        }
                        if (++stage == kStages) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // =============================== MMA issuer ===============================
        if (lane == 0 && leader) {
            constexpr uint32_t idesc = (1u << 4) | (Fmt<T16>::kind << 7) | (Fmt<T16>::kind << 10) |
                                       ((uint32_t)(N_TILE >> 3) << 17) | ((uint32_t)((128 * CG) >> 4) << 24);
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            for (int tile = tile_first; tile < num_tiles; tile += tile_step) {
                mbar_wait_backoff(&tmem_empty[acc], acc_phase);   // the epilogue has pre-loaded this accumulator with the bias
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * ACC_STRIDE);
                for (int kc = 0; kc < k_chunks; ++kc) {
                    mbar_wait_backoff(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(tiles + stage * kStageBytes);
                    const uint32_t b_addr = a_addr + kATileBytes;
#pragma unroll
                    for (int k = 0; k < kBlockK / 16; ++k) {
                        if (CG == 2) tc_mma_f16_pair(d_tmem, umma_smem_desc(a_addr + k * 32), umma_smem_desc(b_addr + k * 32), idesc, 1u);
                        else tc_mma_f16(d_tmem, umma_smem_desc(a_addr + k * 32), umma_smem_desc(b_addr + k * 32), idesc, 1u);
                    }
                    if (CG == 2) tc_commit_pair(&empty_bar[stage]); else tc_commit(&empty_bar[stage]);   // frees the smem stage (both CTAs)
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
                if (CG == 2) tc_commit_pair(&tmem_full[acc]); else tc_commit(&tmem_full[acc]);           // accumulator ready for the epilogue(s)
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
        __syncwarp();
    } else {
        // =============================== epilogue (8 warps) ===============================
        // Two warps share each TMEM lane quarter (hardware: warp w may read lanes 32*(w%4)..+31); each takes
        // half of the tile's columns.  The conv bias is not added here: the epilogue writes it into the
        // accumulator (tcgen05.st) before the MMA warp starts the tile, so TMEM already holds conv + bias.
        const int q = warp & 3;                                   // TMEM lane quarter this warp may access
        constexpr int EW = epi_warps_for(N_TILE, EPI);            // epilogue warps
        constexpr int PARTS = EW / 4;                             // column parts per lane quarter (2 or 4)
        const int half = (warp - 2) >> 2;                         // which column part of the N tile this warp owns
        const int row = q * 32 + lane;                            // tile row == TMEM lane
        const int et = (warp - 2) * 32 + lane;                    // 0..255 index among epilogue threads
        constexpr bool T3 = (EPI == EPI_GN_MISH_T3);
        constexpr int HALF_N = T3 ? 96 : N_TILE / PARTS;          // columns per thread
        constexpr int NCHUNK = HALF_N / 32;
        constexpr int HG = (EPI == EPI_GN_MISH) ? HALF_N / CPG : 1;   // groups owned by one thread (generic GN path)
        constexpr int GCOLS = 3 * CPG;                            // T3: accumulator columns of one GroupNorm group
        constexpr int PART_BUF = 1024;                            // T3: floats per parity buffer of the half-row exchange
        T16* out = reinterpret_cast<T16*>(p.out);
        const T16* res = reinterpret_cast<const T16*>(p.add_res);
        const float4* bias4 = reinterpret_cast<const float4*>(vec_bias);
        const float4* gamma4 = reinterpret_cast<const float4*>(vec_gamma);
        const float4* beta4 = reinterpret_cast<const float4*>(vec_beta);
        const float4* add4 = reinterpret_cast<const float4*>(vec_add);
        const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * HALF_N);
        const bool stage_out = (N_TILE <= 128) && p.tma_out;

        // first channel of 32-column chunk `cc` of this thread for N tile `nt`
        auto chunk_channel = [&](int nt, int cc) -> int {
            if (T3) {
                const int j0 = half * 96 + cc * 32, gl = j0 / GCOLS, rem = j0 - gl * GCOLS, pp = rem / CPG;
                return (nt * (192 / GCOLS) + gl) * CPG + (rem - pp * CPG);
            }
            return nt * N_TILE + half * HALF_N + cc * 32;
        };
        // write the conv bias of N tile `nt` into accumulator buffer `buf` (own lane, own column half)
        auto init_accumulator = [&](int buf, int nt) {
#pragma unroll
            for (int cc = 0; cc < NCHUNK; ++cc) {
                const int chb = chunk_channel(nt, cc);
                float bv[32];
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    const float4 b4 = bias4[(chb + i) >> 2];
                    bv[i] = b4.x; bv[i + 1] = b4.y; bv[i + 2] = b4.z; bv[i + 3] = b4.w;
                }
                tmem_st32(lane_base + (uint32_t)(buf * ACC_STRIDE + cc * 32), bv);
            }
            tmem_st_wait();
        };
        // loop-invariant GroupNorm geometry of this thread's row (generic path: rows of a slice are H consecutive lanes)
        const int sl = min(row / p.H, p.slices_per_tile - 1);     // slice within the tile
        const float inv_cnt = 1.0f / (float)(p.H * CPG);

        {   // both accumulators start out holding the bias of the first two tiles of this CTA
            const int t0 = tile_first, t1 = tile_first + tile_step;
            if (t0 < num_tiles) init_accumulator(0, vtile(t0) % n_tiles);
            if (t1 < num_tiles) init_accumulator(1, vtile(t1) % n_tiles);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (CG == 2) { mbar_arrive_on_cta(&tmem_empty[0], 0); mbar_arrive_on_cta(&tmem_empty[1], 0); }
                else { mbar_arrive(&tmem_empty[0]); mbar_arrive(&tmem_empty[1]); }
            }
        }
        int acc = 0; uint32_t acc_phase = 0;
        for (int tile = tile_first; tile < num_tiles; tile += tile_step) {
            const int vt = vtile(tile);
            const int m_pair = vt / n_tiles, n_tile = vt - m_pair * n_tiles;
            const int m_tile = m_pair * CG + (int)cta_rank;
            const long long s0 = (long long)m_tile * p.slices_per_tile;
            const int n0 = n_tile * N_TILE;
            const bool valid = T3 ? (s0 + row) < p.S : (row < p.rows_used && (s0 + sl) < p.S);
            const long long grow = s0 * p.H + row;                // global row (slices are contiguous rows); T3: slice index
            // element offset of chunk cc of this thread's row in the [S][H][cout] output / residual tensors
            auto chunk_offset = [&](int cc) -> long long {
                if (T3) {
                    const int j0 = half * 96 + cc * 32, gl = j0 / GCOLS, rem = j0 - gl * GCOLS, pp = rem / CPG;
                    return (grow * 3 + pp) * p.cout + chunk_channel(n_tile, cc);
                }
                return grow * p.cout + chunk_channel(n_tile, cc);
            };
            // the residual rows of this tile do not depend on the accumulator: start fetching them before the wait
            uint4 rv[4];
            if (res != nullptr && valid) {
                const uint4* rp = reinterpret_cast<const uint4*>(res + chunk_offset(0));
#pragma unroll
                for (int j = 0; j < 4; ++j) rv[j] = rp[j];
            }
            // ... and pull the NEXT tile's residual rows into L2 now, a whole tile ahead of their use
            if (res != nullptr) {
                const int nt_tile = tile + tile_step;
                if (nt_tile < num_tiles) {
                    const int nvt = vtile(nt_tile);
                    const int nm_pair = nvt / n_tiles, nn_tile = nvt - nm_pair * n_tiles;
                    const long long ns0 = (long long)(nm_pair * CG + (int)cta_rank) * p.slices_per_tile;
                    const long long ngrow = ns0 * p.H + row;
                    const bool nvalid = T3 ? (ns0 + row) < p.S : (row < p.rows_used && (ns0 + sl) < p.S);
                    if (nvalid) {
#pragma unroll
                        for (int cc = 0; cc < NCHUNK; cc += 2) {       // 64 channels = one 128-byte line
                            long long off;
                            if (T3) {
                                const int j0 = half * 96 + cc * 32, gl = j0 / GCOLS, rem = j0 - gl * GCOLS, pp = rem / CPG;
                                off = (ngrow * 3 + pp) * p.cout + chunk_channel(nn_tile, cc);
                            } else {
                                off = ngrow * p.cout + chunk_channel(nn_tile, cc);
                            }
                            asm volatile("prefetch.global.L2 [%0];" ::"l"(res + off));
                        }
                    }
                }
            }
            mbar_wait_backoff(&tmem_full[acc], acc_phase);
            tc_fence_after();
            const uint32_t taddr = lane_base + (uint32_t)(acc * ACC_STRIDE);
            // accumulator values are read from TMEM once when they fit in registers (<= 64 columns per thread)
            constexpr bool KEEP = HAS_GN && (NCHUNK == 1 || (NCHUNK == 2 && EW == 8 && OCC == 1));   // (96-register budget at OCC == 2)
            float vk[KEEP ? NCHUNK : 1][32];
            float v[32];
            float g_sc[HG], g_sh[HG];                             // per group: rstd and -mean * rstd
            float t3_rstd = 0.f, t3_nm = 0.f;
            if (EPI == EPI_GN_MISH) {
                // ---- pass 1: per-row sums of every group in this thread's column half ----
                float s1[HG], s2[HG];
#pragma unroll
                for (int g = 0; g < HG; ++g) { s1[g] = 0.f; s2[g] = 0.f; }
#pragma unroll
                for (int cc = 0; cc < NCHUNK; ++cc) {
                    float (&vv)[32] = KEEP ? vk[KEEP ? cc : 0] : v;
                    tmem_ld32(taddr + cc * 32, vv);
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const int g = (cc * 32 + i) / CPG;
                        s1[g] += vv[i];
                        s2[g] = fmaf(vv[i], vv[i], s2[g]);
                    }
                }
                // ---- reduce over the H rows of each slice in a FIXED order (h = 0..H-1), so that a slice's statistics do
                //      not depend on where it sits inside the tile: results are invariant to batch size / sharding ----
                float2* pr = reinterpret_cast<float2*>(part);                       // [128 rows][8 groups]
                float2* tot = reinterpret_cast<float2*>(part + 2048);               // [43 slices][8 groups]
#pragma unroll
                for (int g = 0; g < HG; ++g) pr[row * 8 + half * HG + g] = make_float2(s1[g], s2[g]);
                asm volatile("bar.sync 1, %0;" ::"n"(EW * 32) : "memory");
                {
                    // (slice, group) totals, each the sum of H row partials.  A team of TS adjacent lanes owns one total:
                    // every member adds H / TS consecutive rows in order and the members are combined by a fixed
                    // shuffle tree, so the association order depends on h only.
                    constexpr int G = N_TILE / CPG;                                  // groups per N tile
                    const int n_tot = p.slices_per_tile * G;
                    const int ts = (p.H % 4 == 0 && n_tot * 4 <= EW * 32) ? 4 : ((p.H % 2 == 0 && n_tot * 2 <= EW * 32) ? 2 : 1);
                    const int per = p.H / ts;
                    for (int base = 0; base < n_tot * ts; base += EW * 32) {
                        const int item = base + et;
                        const bool live = item < n_tot * ts;
                        const int tid = live ? item / ts : 0, part = live ? item - tid * ts : 0;
                        const int tsl = tid / G, tg = tid - tsl * G;
                        float a = 0.f, b2 = 0.f;
                        const float2* src = pr + (tsl * p.H + part * per) * 8 + tg;
                        if (live)
                            for (int h = 0; h < per; ++h) { const float2 t = src[h * 8]; a += t.x; b2 += t.y; }
                        if (ts >= 2) { a += __shfl_xor_sync(0xffffffffu, a, 1); b2 += __shfl_xor_sync(0xffffffffu, b2, 1); }
                        if (ts == 4) { a += __shfl_xor_sync(0xffffffffu, a, 2); b2 += __shfl_xor_sync(0xffffffffu, b2, 2); }
                        if (live && part == 0) tot[tsl * 8 + tg] = make_float2(a, b2);
                    }
                }
                if (stage_out && et == 0) tma_store_wait_read();  // staging is free again after this barrier
                asm volatile("bar.sync 1, %0;" ::"n"(EW * 32) : "memory");
#pragma unroll
                for (int g = 0; g < HG; ++g) {
                    const float2 t = tot[sl * 8 + half * HG + g];
                    const float mean = t.x * inv_cnt;
                    const float rstd = rsqrtf(fmaxf(t.y * inv_cnt - mean * mean, 0.f) + 1e-5f);
                    g_sc[g] = rstd; g_sh[g] = -mean * rstd;
                }
            } else if (T3) {
                // ---- block-Toeplitz tile: row = slice; the group statistics live in this row's own columns ----
                float s1 = 0.f, s2 = 0.f;
#pragma unroll
                for (int cc = 0; cc < NCHUNK; ++cc) {
                    tmem_ld32(taddr + cc * 32, v);
#pragma unroll
                    for (int i = 0; i < 32; ++i) { s1 += v[i]; s2 = fmaf(v[i], v[i], s2); }
                }
                if (CPG == 64) {
                    // the group spans both column halves: exchange partial sums with the partner warp
                    float* ex = part + acc * PART_BUF;
                    *reinterpret_cast<float2*>(ex + (half * 128 + row) * 2) = make_float2(s1, s2);
                    asm volatile("bar.sync 1, %0;" ::"n"(EW * 32) : "memory");
                    const float2 o = *reinterpret_cast<const float2*>(ex + ((half ^ 1) * 128 + row) * 2);
                    s1 += o.x; s2 += o.y;
                }
                const float mean = s1 * (1.0f / (float)GCOLS);
                t3_rstd = rsqrtf(fmaxf(s2 * (1.0f / (float)GCOLS) - mean * mean, 0.f) + 1e-5f);
                t3_nm = -mean * t3_rstd;
            } else if (stage_out) {
                if (et == 0) tma_store_wait_read();               // the previous tile's store has finished reading staging
                asm volatile("bar.sync 1, %0;" ::"n"(EW * 32) : "memory");
            }
            // ---- pass 2 (or the only pass): normalise / activate / add / store ----
#pragma unroll
            for (int cc = 0; cc < NCHUNK; ++cc) {
                if (KEEP && EPI == EPI_GN_MISH) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = vk[KEEP ? cc : 0][i];
                } else {
                    tmem_ld32(taddr + cc * 32, v);
                }
                const int chb = chunk_channel(n_tile, cc);
                uint32_t packed[16];
                uint4 rn[4];
                if (res != nullptr && valid && cc + 1 < NCHUNK) {  // next chunk's residual, in flight during this chunk's math
                    const uint4* rp = reinterpret_cast<const uint4*>(res + chunk_offset(cc + 1));
#pragma unroll
                    for (int j = 0; j < 4; ++j) rn[j] = rp[j];
                }
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    float y[4];
                    if (HAS_GN) {
                        const int ch4 = (chb + i) >> 2;
                        const float4 ga4 = gamma4[ch4], be4 = beta4[ch4], ad4 = add4[ch4];
                        const float ga[4] = {ga4.x, ga4.y, ga4.z, ga4.w}, be[4] = {be4.x, be4.y, be4.z, be4.w};
                        const float ad[4] = {ad4.x, ad4.y, ad4.z, ad4.w};
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int g = T3 ? 0 : (cc * 32 + i + u) / CPG;
                            const float rs = T3 ? t3_rstd : g_sc[g], nm = T3 ? t3_nm : g_sh[g];
                            const float x = fmaf(fmaf(v[i + u], rs, nm), ga[u], be[u]);      // ((v - mean) * rstd) * gamma + beta
                            y[u] = mish_fast(x) + ad[u];
                        }
                    } else {
#pragma unroll
                        for (int u = 0; u < 4; ++u) y[u] = v[i + u];
                    }
                    if (res != nullptr) {
                        const uint32_t* rw = reinterpret_cast<const uint32_t*>(rv);
                        const float2 r0 = unpack2<T16>(rw[i >> 1]), r1 = unpack2<T16>(rw[(i >> 1) + 1]);
                        y[0] += r0.x; y[1] += r0.y; y[2] += r1.x; y[3] += r1.y;
                    }
                    packed[i >> 1] = pack2<T16>(y[0], y[1]);
                    packed[(i >> 1) + 1] = pack2<T16>(y[2], y[3]);
                }
                if (stage_out) {
                    // 128-byte-swizzled staging slab of 64 channels: 16-byte chunk index XOR (row & 7)
                    const int col = half * HALF_N + cc * 32;       // column inside the N tile
                    uint8_t* slab = staging + (col >> 6) * kATileBytes + row * 128;
                    const int ck = (col & 63) >> 3;
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        *reinterpret_cast<uint4*>(slab + (((ck + j) ^ (row & 7)) << 4)) =
                            make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
                } else if (valid) {
                    const long long off = T3 ? chunk_offset(cc) : (grow * p.out_mul + p.out_add) * p.cout + chb;
                    uint4* op = reinterpret_cast<uint4*>(out + off);
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        op[j] = make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
                }
                if (res != nullptr && cc + 1 < NCHUNK) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) rv[j] = rn[j];
                }
            }
            // ---- hand the accumulator back, pre-loaded with the bias of the tile that will use it next ----
            {
                const int nxt = tile + 2 * tile_step;
                if (nxt < num_tiles) init_accumulator(acc, vtile(nxt) % n_tiles);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { if (CG == 2) mbar_arrive_on_cta(&tmem_empty[acc], 0); else mbar_arrive(&tmem_empty[acc]); }
            if (stage_out) {
                fence_proxy_async();                               // generic-proxy smem writes -> visible to the TMA engine
                asm volatile("bar.sync 1, %0;" ::"n"(EW * 32) : "memory");
                if (et == 0) {
#pragma unroll
                    for (int sl64 = 0; sl64 < N_TILE / 64; ++sl64)
                        tma_store_3d(&map_out, staging + sl64 * kATileBytes, n0 + sl64 * 64, 0, (int)s0);
                    tma_store_commit();
                }
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        if (stage_out && et == 0) tma_store_wait_all();           // global writes complete before the CTA exits
    }
    tc_fence_before();
    if (CG == 2) cluster_sync_all(); else __syncthreads();      // nobody leaves while its partner may still touch its smem / TMEM
    if (warp == 1) {
        tc_fence_after();
        if (CG == 2) tmem_dealloc_pair(tmem_base, kTmemCols); else tmem_dealloc(tmem_base, kTmemCols);
    }
}

// ------------------------------------------------------------------ host side
template <int N_TILE, int CG, int OCC, int EPI>
constexpr size_t smem_bytes_for() {
    return 1024 + (size_t)stages_for(N_TILE, CG, OCC) * (kATileBytes + N_TILE * 128 / CG) + (N_TILE <= 128 ? (N_TILE / 64) * kATileBytes : 0) +
           (size_t)vec_count_for(EPI, OCC) * vec_cap_for(N_TILE, EPI, OCC) * 4 + (size_t)part_floats_for(EPI, OCC) * 4 +
           (2 * stages_for(N_TILE, CG, OCC) + 4) * 8 + 16;
}

template <typename T16, int N_TILE, int CPG, int EPI, int CG, int OCC = 1>
int launch_instance(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b, const CUtensorMap& mo, const TcParams& p,
                    cudaStream_t st) {
    auto kern = conv_tc_kernel<T16, N_TILE, CPG, EPI, CG, OCC>;
    constexpr size_t smem = smem_bytes_for<N_TILE, CG, OCC, EPI>();
    static_assert(OCC == 1 || smem <= 113 * 1024, "two CTAs per SM need <= 113 KB of shared memory each");
    static DeviceOnce once;
    if (once.first_time()) {
        CINDM_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    const int tiles = ((p.m_tiles + CG - 1) / CG) * p.n_tiles;          // (pair-)tiles
    const int slots = num_sms() * OCC / CG;                             // CTAs (clusters) resident on the machine at once
    const int grid = (tiles < slots ? tiles : slots) * CG;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(threads_for(N_TILE, EPI)); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CG; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 2 : 1;
    CINDM_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, a0, a1, b, mo, p));
    CINDM_CHECK_LAUNCH();
    return 0;
}

// two CTAs per SM for the N = 64 kernels (measured, same box: 64->64@H24 GroupNorm conv 0.145 -> 0.125 ms, the N = 64 plain
// convs -5..-10 %; the N = 128 kernels got 10-15 % SLOWER with the two 32 KB stages that fit, so they stay at one CTA
// per SM); CINDM_CONV_OCC2=0 selects the one-CTA-per-SM variant (A/B runs)
bool occ2_mode() {
    static int mode = -1;
    if (mode < 0) { const char* e = getenv("CINDM_CONV_OCC2"); mode = (e && e[0] == '0') ? 0 : 1; }
    return mode == 1;
}

// N tile of the plain (1x1 / resampling) convs.  Measured per layer (same box, profiles/r2_plain_conv_tiles.md): with a short
// K loop (1x1 convs from 64 / 128 channels: the attention output projections, two residual convs) the kernel is bound by its
// epilogue and the 64-column tiles of the two-CTAs-per-SM kernel win (-8..-18 %); with a long K loop (cin >= 256, the 3- / 4-tap
// resampling convs) re-reading the A tile once per N tile costs more and 128-column CTA pairs win (by 20-35 %).
// CINDM_CONV_BIAS64=1 / 0 forces 64 / 128 everywhere (A/B runs).
int plain_n_tile(int cout, int taps, int cin) {
    static int mode = -2;
    if (mode == -2) { const char* e = getenv("CINDM_CONV_BIAS64"); mode = !e ? -1 : (e[0] == '1' ? 1 : 0); }
    if (cout % 128 != 0) return 64;
    if (mode >= 0) return mode == 1 ? 64 : 128;
    return taps * cin > 128 ? 128 : 64;
}

// CTA pairs for the N = 128 kernels too (each CTA then stages only half of the B tile: -25 % operand traffic per tile on the
// L2-fabric-bound 128-channel layers); CINDM_CONV_PAIR128=0 keeps them single-CTA (A/B runs)
bool pair128_mode() {
    static int mode = -1;
    if (mode < 0) { const char* e = getenv("CINDM_CONV_PAIR128"); mode = (e && e[0] == '0') ? 0 : 1; }
    return mode == 1;
}

// CTA-pair (cta_group::2) MMA for the N >= 192 kernels; CINDM_CONV_PAIR=0 selects the single-CTA variant (A/B runs)
bool pair_mode() {
    static int mode = -1;
    if (mode < 0) { const char* e = getenv("CINDM_CONV_PAIR"); mode = (e && e[0] == '0') ? 0 : 1; }
    return mode == 1;
}

template <typename T16>
int dispatch(const ConvTcLaunch& a, const CUtensorMap& m0, const CUtensorMap& m1, const CUtensorMap& mb, const CUtensorMap& mo,
             const TcParams& p, int n_tile, cudaStream_t st) {
    if (a.epilogue == EPI_GN_MISH_T3) {
        switch (p.cout) {
            case 256: return pair_mode() ? launch_instance<T16, 192, 32, EPI_GN_MISH_T3, 2>(m0, m1, mb, mo, p, st)
                                         : launch_instance<T16, 192, 32, EPI_GN_MISH_T3, 1>(m0, m1, mb, mo, p, st);
            case 512: return pair_mode() ? launch_instance<T16, 192, 64, EPI_GN_MISH_T3, 2>(m0, m1, mb, mo, p, st)
                                         : launch_instance<T16, 192, 64, EPI_GN_MISH_T3, 1>(m0, m1, mb, mo, p, st);
        }
        return fail(-2, "conv_tc: the block-Toeplitz path is built for 256 / 512 output channels");
    }
    if (a.epilogue == EPI_GN_MISH) {
        switch (p.cout) {
            case 64: return occ2_mode() ? launch_instance<T16, 64, 8, EPI_GN_MISH, 1, 2>(m0, m1, mb, mo, p, st)
                                        : launch_instance<T16, 64, 8, EPI_GN_MISH, 1>(m0, m1, mb, mo, p, st);
            case 128: return pair128_mode() ? launch_instance<T16, 128, 16, EPI_GN_MISH, 2>(m0, m1, mb, mo, p, st)
                                            : launch_instance<T16, 128, 16, EPI_GN_MISH, 1>(m0, m1, mb, mo, p, st);
            case 256: return pair_mode() ? launch_instance<T16, 256, 32, EPI_GN_MISH, 2>(m0, m1, mb, mo, p, st)
                                         : launch_instance<T16, 256, 32, EPI_GN_MISH, 1>(m0, m1, mb, mo, p, st);
            case 512: return pair_mode() ? launch_instance<T16, 256, 64, EPI_GN_MISH, 2>(m0, m1, mb, mo, p, st)
                                         : launch_instance<T16, 256, 64, EPI_GN_MISH, 1>(m0, m1, mb, mo, p, st);
        }
        return fail(-2, "conv_tc: unsupported channel count for the GroupNorm epilogue");
    }
    switch (n_tile) {
        case 64: return occ2_mode() ? launch_instance<T16, 64, 8, EPI_BIAS, 1, 2>(m0, m1, mb, mo, p, st)
                                    : launch_instance<T16, 64, 8, EPI_BIAS, 1>(m0, m1, mb, mo, p, st);
        case 128: return pair128_mode() ? launch_instance<T16, 128, 8, EPI_BIAS, 2>(m0, m1, mb, mo, p, st)
                                        : launch_instance<T16, 128, 8, EPI_BIAS, 1>(m0, m1, mb, mo, p, st);
        case 256: return launch_instance<T16, 256, 8, EPI_BIAS, 1>(m0, m1, mb, mo, p, st);
    }
    return fail(-2, "conv_tc: unsupported N tile");
}

}  // namespace

static int launch_conv_tc_one(const ConvTcLaunch& a, int parity, cudaStream_t st) {
    const ConvW& w = *a.w;
    TcParams p;
    p.bias = w.bias;
    p.gamma = a.gn ? a.gn->gamma : nullptr;
    p.beta = a.gn ? a.gn->beta : nullptr;
    p.add_vec = a.add_vec; p.t_dev = a.t_dev; p.add_res = a.add_res; p.out = a.out;
    p.S = a.S; p.cout = w.cout; p.c0 = a.c0; p.c1 = a.in1 ? a.c1 : 0; p.cin = w.cin; p.kpp = w.cin / kBlockK;
    p.out_mul = 1; p.out_add = 0; p.reverse = a.reverse;
    if (a.epilogue == EPI_GN_MISH_T3) {
        // one dense GEMM: rows = slices, K = 3*cin, N = 3*cout in (group, position, channel) order
        p.H = 1; p.taps = 1; p.tap_hoff[0] = 0; p.tap_wrow[0] = 0;
        p.slices_per_tile = 128; p.rows_used = 128;
        p.m_tiles = (int)((a.S + 127) / 128);
        p.n_tiles = 3 * w.cout / 192;
        p.k_chunks_per_tap = 3 * w.cin / kBlockK;
        char tag[96];
        snprintf(tag, sizeof tag, "conv_tc toep H3 %d->%d k5 gn", w.cin, w.cout);
        KernelTimer kt(tag, st, 2.0 * (double)a.S * 9.0 * w.cin * w.cout);
        CUtensorMap m0, m1, mb;
        CINDM_TRY(encode_act_map(&m0, a.in0, a.prec, a.S, 1, 3 * a.c0, 128, 1, 1));
        if (a.in1) CINDM_TRY(encode_act_map(&m1, a.in1, a.prec, a.S, 1, 3 * a.c1, 128, 1, 1));
        else m1 = m0;
        CINDM_TRY(encode_weight_map(&mb, w.w16t[a.prec], a.prec, 3 * w.cout, 3 * w.cin, pair_mode() ? 96 : 192));   // CTA pair: half of B each
        p.tma_out = 0;
        if (a.prec == PREC_F16) return dispatch<__half>(a, m0, m1, mb, m0, p, 192, st);
        return dispatch<__nv_bfloat16>(a, m0, m1, mb, m0, p, 192, st);
    }
    int h_stride = 1;
    double nz_taps;                         // in-range taps summed over one slice's output positions
    if (a.mode == TC_SAME) {
        p.H = a.H; p.taps = w.taps;
        for (int k = 0; k < w.taps; ++k) { p.tap_hoff[k] = k - w.taps / 2; p.tap_wrow[k] = k * w.cout; }
        nz_taps = w.taps == 1 ? (double)a.H : (double)(5 * a.H - 6);
    } else if (a.mode == TC_DOWN) {         // out[p] = sum_k in[2p + k - 1] W[k]
        p.H = a.H / 2; p.taps = 3; h_stride = 2;
        for (int k = 0; k < 3; ++k) { p.tap_hoff[k] = k - 1; p.tap_wrow[k] = k * w.cout; }
        nz_taps = 3.0 * p.H - 1.0;
    } else {                                // out[2m] = in[m] W[1] + in[m-1] W[3];  out[2m+1] = in[m] W[2] + in[m+1] W[0]
        p.H = a.H; p.taps = 2; p.out_mul = 2; p.out_add = parity;
        p.tap_hoff[0] = 0; p.tap_wrow[0] = (parity ? 2 : 1) * w.cout;
        p.tap_hoff[1] = parity ? 1 : -1; p.tap_wrow[1] = (parity ? 0 : 3) * w.cout;
        nz_taps = 2.0 * a.H - 1.0;
    }
    p.slices_per_tile = 128 / p.H;
    p.rows_used = p.slices_per_tile * p.H;
    p.m_tiles = (int)((a.S + p.slices_per_tile - 1) / p.slices_per_tile);
    int n_tile;
    if (a.epilogue == EPI_GN_MISH) n_tile = w.cout < 256 ? w.cout : 256;
    else n_tile = plain_n_tile(w.cout, p.taps, w.cin);   // (N <= 128 kernels write their output tile with TMA)
    p.n_tiles = w.cout / n_tile;
    p.k_chunks_per_tap = w.cin / kBlockK;

    // algorithmic work: 2 x nonzero-tap MACs
    char tag[96];
    snprintf(tag, sizeof tag, "conv_tc %s H%d %d->%d k%d %s", a.mode == TC_SAME ? "same" : (a.mode == TC_DOWN ? "down" : "up"),
             a.H, w.cin, w.cout, w.taps, a.epilogue == EPI_GN_MISH ? "gn" : "bias");
    KernelTimer kt(tag, st, 2.0 * (double)a.S * nz_taps * w.cin * w.cout);
    CUtensorMap m0, m1, mb;
    CINDM_TRY(encode_act_map(&m0, a.in0, a.prec, a.S, a.H, a.c0, p.slices_per_tile, p.H, h_stride));
    if (a.in1) CINDM_TRY(encode_act_map(&m1, a.in1, a.prec, a.S, a.H, a.c1, p.slices_per_tile, p.H, h_stride));
    else m1 = m0;
    // the N = 256 GroupNorm kernels and (all) the N = 128 kernels run as CTA pairs
    const int cg = ((a.epilogue == EPI_GN_MISH && n_tile == 256 && pair_mode()) || (n_tile == 128 && pair128_mode())) ? 2 : 1;
    CINDM_TRY(encode_weight_map(&mb, w.w16[a.prec], a.prec, w.taps * w.cout, w.cin, n_tile / cg));
    // output tile through shared memory + TMA store: whole-row 128-byte bursts instead of 32 scattered 16-byte
    // stores per warp instruction (the transposed conv's interleaved rows keep the direct stores)
    CUtensorMap mo = m0;
    p.tma_out = (n_tile <= 128 && a.mode != TC_UP) ? 1 : 0;
    if (p.tma_out) CINDM_TRY(encode_act_map(&mo, a.out, a.prec, a.S, p.H, w.cout, p.slices_per_tile, p.H, 1));
    if (a.prec == PREC_F16) return dispatch<__half>(a, m0, m1, mb, mo, p, n_tile, st);
    return dispatch<__nv_bfloat16>(a, m0, m1, mb, mo, p, n_tile, st);
}

int launch_conv_tc(const ConvTcLaunch& a, cudaStream_t st) {
    const ConvW& w = *a.w;
    if (a.prec != PREC_F16 && a.prec != PREC_BF16) return fail(-2, "conv_tc: 16-bit precisions only");
    const int want_taps = a.mode == TC_SAME ? w.taps : (a.mode == TC_DOWN ? 3 : 4);
    if (w.taps != want_taps || (a.mode == TC_SAME && w.taps != 1 && w.taps != 5))
        return fail(-2, "conv_tc: tap count does not match the conv mode");
    if (w.cin % 64 || w.cout % 64) return fail(-2, "conv_tc: channel counts must be multiples of 64");
    if (a.c0 + (a.in1 ? a.c1 : 0) != w.cin) return fail(-2, "conv_tc: input channels do not match the weight");
    if (a.in1 && (a.c0 % 64)) return fail(-2, "conv_tc: concat split must be a multiple of 64 channels");
    if (a.H < 1 || a.H > 128 || (a.mode == TC_DOWN && (a.H % 2))) return fail(-2, "conv_tc: bad H");
    if (a.mode != TC_SAME && (a.epilogue != EPI_BIAS || a.add_res)) return fail(-2, "conv_tc: resampling convs take the plain epilogue");
    if (a.S == 0) return 0;
    if (a.epilogue != EPI_BIAS && !a.gn) return fail(-2, "conv_tc: GroupNorm parameters missing");
    if (a.epilogue == EPI_GN_MISH_T3 && (a.H != 3 || w.taps != 5 || a.mode != TC_SAME || !w.w16t[a.prec]))
        return fail(-2, "conv_tc: the block-Toeplitz path needs H == 3, k == 5 and the repacked operand");
    if (a.mode == TC_UP) {
        if (a.side) {
            CINDM_CHECK_CUDA(cudaEventRecord(a.ev_fork, st));
            CINDM_CHECK_CUDA(cudaStreamWaitEvent(a.side, a.ev_fork, 0));
            CINDM_TRY(launch_conv_tc_one(a, 1, a.side));
            CINDM_CHECK_CUDA(cudaEventRecord(a.ev_join, a.side));
            CINDM_TRY(launch_conv_tc_one(a, 0, st));
            CINDM_CHECK_CUDA(cudaStreamWaitEvent(st, a.ev_join, 0));
            return 0;
        }
        CINDM_TRY(launch_conv_tc_one(a, 0, st));
        return launch_conv_tc_one(a, 1, st);
    }
    if (conv_tc_cm_eligible(a)) return launch_conv_tc_cm(a, st);
    return launch_conv_tc_one(a, 0, st);
}

}  // namespace cindm
