// Per-slice fused kernels of the throughput path (16-bit activations, fp32 math):
//   stem  : [composition gather] + conv(8->64,k5) + GroupNorm + Mish + time bias, and the 1x1 residual conv
//   head  : final 1x1 conv 64 -> 8 to fp32 eps_pair
//   attn  : linear-attention core (softmax_n(k), k v^T, ctx^T q) with register-tiled 4x4 outer products
// They replace generic SIMT GEMM launches whose K or N extent (8 channels) is too small to tile well.
#include <cstdlib>

#include "engine.h"

namespace cindm {

// ------------------------------------------------------------------------------------------
// Linear attention core, one CTA (256 threads = 4 heads x 64) per slice.
//   qkv: [S][n][384] (q | k | v, each 4 heads x 32), out: [S][n][128]
//   reference LinearAttentionTemporal.forward, model/diffusion_1d.py:281-291
// Shared memory rows are padded to 132 floats so that both the float4 channel reads (stage 3)
// and the per-position scalar reads (stage 4) are bank-conflict free.
// ------------------------------------------------------------------------------------------
constexpr int kAttnRow = 132;
constexpr int kCtxRow = 36;

template <typename T>
__device__ __forceinline__ void load8(const T* p, float (&v)[8]);
template <>
__device__ __forceinline__ void load8<__half>(const __half* p, float (&v)[8]) {
    uint4 u = *reinterpret_cast<const uint4*>(p);
    const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) { float2 f = __half22float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}
template <>
__device__ __forceinline__ void load8<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[8]) {
    uint4 u = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) { float2 f = __bfloat1622float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}
template <>
__device__ __forceinline__ void load8<float>(const float* p, float (&v)[8]) {
    float4 a = reinterpret_cast<const float4*>(p)[0], b = reinterpret_cast<const float4*>(p)[1];
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

template <typename T>
__device__ __forceinline__ void unpack8(const uint4& raw, const T* p, float (&v)[8]);
template <>
__device__ __forceinline__ void unpack8<__half>(const uint4& raw, const __half*, float (&v)[8]) {
    const __half2* h = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
    for (int i = 0; i < 4; ++i) { float2 f = __half22float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}
template <>
__device__ __forceinline__ void unpack8<__nv_bfloat16>(const uint4& raw, const __nv_bfloat16*, float (&v)[8]) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
    for (int i = 0; i < 4; ++i) { float2 f = __bfloat1622float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}
template <>
__device__ __forceinline__ void unpack8<float>(const uint4& raw, const float* p, float (&v)[8]) {
    const float4 b = reinterpret_cast<const float4*>(p)[1];
    v[0] = __uint_as_float(raw.x); v[1] = __uint_as_float(raw.y); v[2] = __uint_as_float(raw.z); v[3] = __uint_as_float(raw.w);
    v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

template <typename T>
__device__ __forceinline__ void store4(T* p, float a, float b, float c, float d);
template <>
__device__ __forceinline__ void store4<__half>(__half* p, float a, float b, float c, float d) {
    __half2 lo = __floats2half2_rn(a, b), hi = __floats2half2_rn(c, d);
    uint2 u = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
    *reinterpret_cast<uint2*>(p) = u;
}
template <>
__device__ __forceinline__ void store4<__nv_bfloat16>(__nv_bfloat16* p, float a, float b, float c, float d) {
    __nv_bfloat162 lo = __floats2bfloat162_rn(a, b), hi = __floats2bfloat162_rn(c, d);
    uint2 u = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
    *reinterpret_cast<uint2*>(p) = u;
}
template <>
__device__ __forceinline__ void store4<float>(float* p, float a, float b, float c, float d) {
    *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d);
}

// SLOTS = position slots of 8 a thread covers in stage 4: 3 for n <= 24 (every level of the 24-step model), 6 for n <= 48
// (the 44-step models); every accumulator sums in the same order either way.
template <typename T, int SLOTS = 3>
__global__ void __launch_bounds__(256) attn_core_tiled_kernel(const T* __restrict__ qkv, T* __restrict__ out, int n,
                                                              long long S) {
    extern __shared__ __align__(16) float sm[];
    float* sq = sm;                          // [n][132]  q * 32^-0.5
    float* sk = sq + n * kAttnRow;           // [n][132]
    float* sv = sk + n * kAttnRow;           // [n][132]
    float* ctx = sk;                         // [4][32][36], written over k/v once they are consumed
    const int tid = threadIdx.x;
    const int h = tid >> 6, l = tid & 63;
    constexpr int kMaxVec = (SLOTS * 8 * 48 + 255) / 256;    // n*48 <= 1152 (2304) 16-byte vectors per slice, <= 5 (9) per thread
    constexpr int kVecStride = (int)(sizeof(T) * 8 / 16);
    const int nvec = n * 48;
    uint4 raw[kMaxVec];
    // persistent CTA: slices blockIdx.x, +gridDim.x, ...; the next slice's global reads are in flight while
    // the current one is computed
    long long s = blockIdx.x;
    if (s < S) {
        const T* src = qkv + s * (long long)n * 384;
#pragma unroll
        for (int r = 0; r < kMaxVec; ++r) {
            const int i = tid + r * 256;
            if (i < nvec) raw[r] = reinterpret_cast<const uint4*>(src)[i * kVecStride];
        }
    }
    for (; s < S; s += gridDim.x) {
        const T* src = qkv + s * (long long)n * 384;
        // ---- 1. registers -> shared (fp32)
#pragma unroll
        for (int r = 0; r < kMaxVec; ++r) {
            const int i = tid + r * 256;
            if (i >= nvec) continue;
            const int pos = i / 48, c8 = (i - pos * 48) * 8;
            float v[8];
            unpack8<T>(raw[r], src + (long long)i * 8, v);
            float* dst;
            float scale = 1.0f;
            if (c8 < 128) { dst = sq + pos * kAttnRow + c8; scale = 0.17677669529663687f; }
            else if (c8 < 256) dst = sk + pos * kAttnRow + (c8 - 128);
            else dst = sv + pos * kAttnRow + (c8 - 256);
            *reinterpret_cast<float4*>(dst) = make_float4(v[0] * scale, v[1] * scale, v[2] * scale, v[3] * scale);
            *reinterpret_cast<float4*>(dst + 4) = make_float4(v[4] * scale, v[5] * scale, v[6] * scale, v[7] * scale);
        }
        {
            const long long sn = s + gridDim.x;
            if (sn < S) {
                const T* nsrc = qkv + sn * (long long)n * 384;
#pragma unroll
                for (int r = 0; r < kMaxVec; ++r) {
                    const int i = tid + r * 256;
                    if (i < nvec) raw[r] = reinterpret_cast<const uint4*>(nsrc)[i * kVecStride];
                }
            }
        }
        __syncthreads();
        // ---- 2. softmax over positions for every k channel
        if (tid < 128) {
            float m = -INFINITY;
            for (int j = 0; j < n; ++j) m = fmaxf(m, sk[j * kAttnRow + tid]);
            float sum = 0.f;
            for (int j = 0; j < n; ++j) {
                float e = __expf(sk[j * kAttnRow + tid] - m);
                sk[j * kAttnRow + tid] = e;
                sum += e;
            }
            const float inv = 1.0f / sum;
            for (int j = 0; j < n; ++j) sk[j * kAttnRow + tid] *= inv;
        }
        __syncthreads();
        // ---- 3. ctx[d][e] = sum_j k[d][j] v[e][j], 4x4 register tile per thread
        {
            const int bi = l >> 3, bj = l & 7;
            float acc[4][4];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
            for (int j = 0; j < n; ++j) {
                const float4 kd = *reinterpret_cast<const float4*>(sk + j * kAttnRow + h * 32 + 4 * bi);
                const float4 ve = *reinterpret_cast<const float4*>(sv + j * kAttnRow + h * 32 + 4 * bj);
                const float ka[4] = {kd.x, kd.y, kd.z, kd.w}, va[4] = {ve.x, ve.y, ve.z, ve.w};
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(ka[a], va[b], acc[a][b]);
            }
            __syncthreads();                 // everyone is done reading k / v: reuse their storage for ctx
#pragma unroll
            for (int a = 0; a < 4; ++a)
                *reinterpret_cast<float4*>(ctx + (h * 32 + 4 * bi + a) * kCtxRow + 4 * bj) =
                    make_float4(acc[a][0], acc[a][1], acc[a][2], acc[a][3]);
        }
        __syncthreads();
        // ---- 4. out[e][j] = sum_d ctx[d][e] q[d][j]; thread = 4 channels x positions {ng, ng+8, ng+16, ...}
        {
            const int be = l & 7, ng = l >> 3;
            float acc[SLOTS][4];
            bool on[SLOTS];
#pragma unroll
            for (int k = 0; k < SLOTS; ++k) {
                on[k] = ng + 8 * k < n;
#pragma unroll
                for (int b = 0; b < 4; ++b) acc[k][b] = 0.f;
            }
#pragma unroll 4
            for (int d = 0; d < 32; ++d) {
                const float4 c4 = *reinterpret_cast<const float4*>(ctx + (h * 32 + d) * kCtxRow + 4 * be);
#pragma unroll
                for (int k = 0; k < SLOTS; ++k) {
                    if (!on[k]) continue;
                    const float qv = sq[(ng + 8 * k) * kAttnRow + h * 32 + d];
                    acc[k][0] = fmaf(c4.x, qv, acc[k][0]);
                    acc[k][1] = fmaf(c4.y, qv, acc[k][1]);
                    acc[k][2] = fmaf(c4.z, qv, acc[k][2]);
                    acc[k][3] = fmaf(c4.w, qv, acc[k][3]);
                }
            }
            T* dst = out + s * (long long)n * 128;
#pragma unroll
            for (int k = 0; k < SLOTS; ++k)
                if (on[k]) store4<T>(dst + (ng + 8 * k) * 128 + h * 32 + 4 * be, acc[k][0], acc[k][1], acc[k][2], acc[k][3]);
        }
        __syncthreads();                     // shared memory is rewritten by the next slice
    }
}

static size_t attn_smem_bytes(int n) {
    size_t kv = (size_t)2 * n * kAttnRow, cx = (size_t)4 * 32 * kCtxRow;
    return ((size_t)n * kAttnRow + (kv > cx ? kv : cx)) * sizeof(float);
}

// ------------------------------------------------------------------------------------------
// Tensor-core version for the 16-bit precisions: one warp per (slice, head), no block-level barriers.
//   ctx^T[e][d] = sum_j V[j][e] * softmax_j(K)[j][d]       mma.m16n8k16, A = V^T (ldmatrix.trans), B = K_s (ldmatrix.trans)
//   out[j][e]   = 32^-0.5 * sum_d Q[j][d] * ctx[d][e]       A = Q (ldmatrix), B = ctx^T accumulators re-used in registers:
// the C-fragment of ctx^T (row e = lane/4, cols d = 2*(lane%4)..+1) is exactly the B-fragment (k = d, n = e) that
// the second product needs, so ctx never leaves the register file.  Positions are padded to 16 or 32 with zeros.
// ------------------------------------------------------------------------------------------
constexpr int kMmaRow = 40;                  // halves per shared-memory row (80 B: ldmatrix conflict-free)

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
template <typename T> __device__ __forceinline__ uint32_t pack_pair(float a, float b);
template <> __device__ __forceinline__ uint32_t pack_pair<__half>(float a, float b) {
    __half2 h = __floats2half2_rn(a, b); return *reinterpret_cast<uint32_t*>(&h);
}
template <> __device__ __forceinline__ uint32_t pack_pair<__nv_bfloat16>(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b); return *reinterpret_cast<uint32_t*>(&h);
}

template <typename T, int KS>      // KS = number of 16-position steps (1: n <= 16, 2: n <= 32)
__global__ void __launch_bounds__(256) attn_core_mma_kernel(const T* __restrict__ qkv, T* __restrict__ out, int n,
                                                            long long tasks) {
    extern __shared__ __align__(16) unsigned char attn_mma_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    T* sK = reinterpret_cast<T*>(attn_mma_smem) + warp * (3 * 32 * kMmaRow);
    T* sV = sK + 32 * kMmaRow;
    T* sQ = sV + 32 * kMmaRow;
    const int g = lane >> 2, t = lane & 3;
    constexpr int ROWS = 16 * KS;
    const T zero = from_f32<T>(0.f);
    // Software pipeline (KS == 1, i.e. n <= 16: seven of the eight attention blocks): the NEXT task's k / v / q rows are
    // fetched into registers while the current task runs its tensor-core phase, so a warp never sits on DRAM latency.
    constexpr bool PREFETCH = (KS == 1);
    T kraw[PREFETCH ? ROWS : 1];
    uint4 vraw[PREFETCH ? ROWS / 8 : 1], qraw[PREFETCH ? ROWS / 8 : 1];
    auto fetch = [&](long long tk) {
        const long long fs = tk >> 2;
        const T* fb = qkv + fs * (long long)n * 384 + (int)(tk & 3) * 32;
#pragma unroll
        for (int j = 0; j < ROWS; ++j) kraw[PREFETCH ? j : 0] = j < n ? fb[j * 384 + 128 + lane] : zero;
#pragma unroll
        for (int r0 = 0; r0 < ROWS; r0 += 8) {
            const int j = r0 + (lane >> 2), part = lane & 3;
            uint4 vv = make_uint4(0, 0, 0, 0), qq = make_uint4(0, 0, 0, 0);
            if (j < n) {
                vv = *reinterpret_cast<const uint4*>(fb + j * 384 + 256 + part * 8);
                qq = *reinterpret_cast<const uint4*>(fb + j * 384 + part * 8);
            }
            vraw[PREFETCH ? r0 / 8 : 0] = vv; qraw[PREFETCH ? r0 / 8 : 0] = qq;
        }
    };
    const long long task0 = (long long)blockIdx.x * 8 + warp, task_step = (long long)gridDim.x * 8;
    if (PREFETCH && task0 < tasks) fetch(task0);
    for (long long task = task0; task < tasks; task += task_step) {
        const long long s = task >> 2;
        const int h = (int)(task & 3);
        const T* base = qkv + s * (long long)n * 384 + h * 32;
        // ---- K: softmax over positions, one channel per lane
        {
            float kv[ROWS];
            float m = -INFINITY;
#pragma unroll
            for (int j = 0; j < ROWS; ++j) {
                if (PREFETCH) kv[j] = j < n ? to_f32<T>(kraw[PREFETCH ? j : 0]) : -INFINITY;
                else kv[j] = j < n ? to_f32<T>(base[j * 384 + 128 + lane]) : -INFINITY;
                m = fmaxf(m, kv[j]);
            }
            float sum = 0.f;
#pragma unroll
            for (int j = 0; j < ROWS; ++j) { kv[j] = __expf(kv[j] - m); sum += kv[j]; }      // exp(-inf) = 0 pads the rows >= n
            const float inv = 1.0f / sum;
#pragma unroll
            for (int j = 0; j < ROWS; ++j) sK[j * kMmaRow + lane] = from_f32<T>(kv[j] * inv);
        }
        // ---- V and Q rows: 64 bytes each, 4 lanes per row
#pragma unroll
        for (int r0 = 0; r0 < ROWS; r0 += 8) {
            const int j = r0 + (lane >> 2), part = lane & 3;
            uint4 vv = make_uint4(0, 0, 0, 0), qq = make_uint4(0, 0, 0, 0);
            if (PREFETCH) { vv = vraw[PREFETCH ? r0 / 8 : 0]; qq = qraw[PREFETCH ? r0 / 8 : 0]; }
            else if (j < n) {
                vv = *reinterpret_cast<const uint4*>(base + j * 384 + 256 + part * 8);
                qq = *reinterpret_cast<const uint4*>(base + j * 384 + part * 8);
            }
            *reinterpret_cast<uint4*>(sV + j * kMmaRow + part * 8) = vv;
            *reinterpret_cast<uint4*>(sQ + j * kMmaRow + part * 8) = qq;
        }
        if (PREFETCH && task + task_step < tasks) fetch(task + task_step);      // in flight during the MMA phase below
        __syncwarp();
        // ---- ctx^T = V^T K_s : M = e (2 tiles of 16), N = d (4 tiles of 8), K = positions
        float ct[2][4][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < 4; ++nt)
#pragma unroll
                for (int c = 0; c < 4; ++c) ct[mt][nt][c] = 0.f;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
            // B fragments of K_s for all four d tiles: stored [j][d]; matrices (j0, d0) (j0+8, d0) (j0, d0+8) (j0+8, d0+8)
            uint32_t bk[2][4];
#pragma unroll
            for (int np = 0; np < 2; ++np) {
                const int row = 16 * ks + (lane & 7) + ((lane >> 3) & 1) * 8, col = 16 * np + (lane >> 4) * 8;
                ldsm_x4_trans((uint32_t)__cvta_generic_to_shared(sK + row * kMmaRow + col), bk[np]);
            }
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                // A fragments of V^T: stored [j][e]; matrices (j0, e0) (j0, e0+8) (j0+8, e0) (j0+8, e0+8)
                uint32_t av[4];
                const int row = 16 * ks + (lane & 7) + (lane >> 4) * 8, col = 16 * mt + ((lane >> 3) & 1) * 8;
                ldsm_x4_trans((uint32_t)__cvta_generic_to_shared(sV + row * kMmaRow + col), av);
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) mma16816<T>(ct[mt][nt], av, bk[nt >> 1][(nt & 1) * 2], bk[nt >> 1][(nt & 1) * 2 + 1]);
            }
        }
        // ---- out = Q ctx : M = positions (KS tiles of 16), N = e (4 tiles of 8), K = d (2 steps of 16)
        T* dst = out + s * (long long)n * 128 + h * 32;
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
                const int row = 16 * mt + (lane & 7) + ((lane >> 3) & 1) * 8, col = 16 * kd + (lane >> 4) * 8;
                ldsm_x4((uint32_t)__cvta_generic_to_shared(sQ + row * kMmaRow + col), aq);
#pragma unroll
                for (int ne = 0; ne < 4; ++ne) {
                    // B[k = d][n = e] = ctx^T[e][d]: rows e of tile ne live in ct[ne/2] (upper half: c0,c1; lower: c2,c3)
                    const int m2 = ne >> 1, hi = (ne & 1) * 2;
                    const uint32_t b0 = pack_pair<T>(ct[m2][2 * kd][hi], ct[m2][2 * kd][hi + 1]);
                    const uint32_t b1 = pack_pair<T>(ct[m2][2 * kd + 1][hi], ct[m2][2 * kd + 1][hi + 1]);
                    mma16816<T>(oc[ne], aq, b0, b1);
                }
            }
            const float scale = 0.17677669529663687f;                 // 32^-0.5 (q * scale in the reference, :284)
            const int j0 = 16 * mt + g, j1 = j0 + 8;
#pragma unroll
            for (int ne = 0; ne < 4; ++ne) {
                const int e = 8 * ne + 2 * t;
                if (j0 < n) *reinterpret_cast<uint32_t*>(dst + j0 * 128 + e) = pack_pair<T>(oc[ne][0] * scale, oc[ne][1] * scale);
                if (j1 < n) *reinterpret_cast<uint32_t*>(dst + j1 * 128 + e) = pack_pair<T>(oc[ne][2] * scale, oc[ne][3] * scale);
            }
        }
        __syncwarp();            // this warp's shared memory is rewritten by its next task
    }
}

template <typename T>
static int launch_attn_mma(const T* qkv, T* out, int64_t S, int n, cudaStream_t st) {
    const size_t smem = (size_t)8 * 3 * 32 * kMmaRow * sizeof(T);
    static DeviceOnce once;
    if (once.first_time()) {
        CINDM_CHECK_CUDA(cudaFuncSetAttribute(attn_core_mma_kernel<T, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CINDM_CHECK_CUDA(cudaFuncSetAttribute(attn_core_mma_kernel<T, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long tasks = S * 4;
    long long want = (long long)sms * 3;
    const long long need = (tasks + 7) / 8;
    const unsigned grid = (unsigned)(need < want ? need : want);
    if (n <= 16) attn_core_mma_kernel<T, 1><<<grid, 256, smem, st>>>(qkv, out, n, tasks);
    else attn_core_mma_kernel<T, 2><<<grid, 256, smem, st>>>(qkv, out, n, tasks);
    CINDM_CHECK_LAUNCH();
    return 0;
}

int launch_attn_core(const void* qkv, void* out, int64_t S, int n, int prec, cudaStream_t st) {
    if (S == 0) return 0;
    if (n > 48) return fail(-2, "attention core supports at most 48 positions");
    if (n > 24 && prec != PREC_F32) return fail(-2, "the 16-bit attention core supports at most 24 positions");
    KernelTimer kt("attn_core", st, (double)S * n * 512.0 * elem_size(prec));
    if (prec == PREC_F16) return launch_attn_mma<__half>((const __half*)qkv, (__half*)out, S, n, st);
    if (prec == PREC_BF16) return launch_attn_mma<__nv_bfloat16>((const __nv_bfloat16*)qkv, (__nv_bfloat16*)out, S, n, st);
    const size_t smem = attn_smem_bytes(n);
    static DeviceOnce once;
    if (once.first_time()) {
        const int mx = (int)attn_smem_bytes(24);
        CINDM_CHECK_CUDA(cudaFuncSetAttribute(attn_core_tiled_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
        CINDM_CHECK_CUDA(cudaFuncSetAttribute(attn_core_tiled_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
        CINDM_CHECK_CUDA(cudaFuncSetAttribute(attn_core_tiled_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
        CINDM_CHECK_CUDA(cudaFuncSetAttribute(attn_core_tiled_kernel<float, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              (int)attn_smem_bytes(48)));
    }
    // persistent grid: as many CTAs as fit on the machine at this shared-memory footprint
    int per_sm = (int)((size_t)(227 * 1024) / (smem + 1024));
    if (per_sm > 8) per_sm = 8;
    if (per_sm < 1) per_sm = 1;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    long long want = (long long)sms * per_sm;
    const unsigned grid = (unsigned)(S < want ? S : want);
    switch (prec) {
        case PREC_F32:
            if (n > 24) attn_core_tiled_kernel<float, 6><<<grid, 256, smem, st>>>((const float*)qkv, (float*)out, n, S);
            else attn_core_tiled_kernel<float><<<grid, 256, smem, st>>>((const float*)qkv, (float*)out, n, S);
            break;
        case PREC_F16: attn_core_tiled_kernel<__half><<<grid, 256, smem, st>>>((const __half*)qkv, (__half*)out, n, S); break;
        case PREC_BF16:
            attn_core_tiled_kernel<__nv_bfloat16><<<grid, 256, smem, st>>>((const __nv_bfloat16*)qkv, (__nv_bfloat16*)out, n, S);
            break;
        default: return fail(-2, "attn: bad precision");
    }
    CINDM_CHECK_LAUNCH();
    return 0;
}

// ------------------------------------------------------------------------------------------
// Stem: the first ResidualTemporalBlock's first Conv1dBlock (8 -> 64, k=5) + GroupNorm(8) + Mish + time
// bias, and its 1x1 residual conv (8 -> 64), straight from the fp32 design tensor.  With
// gather != 0 the composition gather (reference :985) is fused into the loader: slice s =
// (kk*P + pair)*B + b reads x[b][kk*start + h][4*body + f].
// One warp-pair (64 threads = 64 output channels) per slice; weights live in registers.
// ------------------------------------------------------------------------------------------
struct StemParams {
    const float* x;            // gather: [B][T][4n];  else slices [S][24][F]   (F = 8: body pair, F = 4: single body)
    const float* w0;           // [5][F][64]
    const float* b0;           // [64]
    const float* gamma; const float* beta;
    const float* tbias;        // [64] or [timesteps][64] with t_dev
    const int* t_dev;
    const float* wr;           // [1][F][64]
    const float* br;           // [64]
    void* out_b0;              // [S][24][64]
    void* out_res;             // [S][24][64]
    long long S;
    int gather, B, n, P, start, T;
};

template <typename T, int F>
__global__ void __launch_bounds__(256, 2) stem_kernel(StemParams p) {
    __shared__ float xs[4][24 + 4][F];                      // per-slice input with 2 zero rows of padding each side
    pdl_wait();
    pdl_trigger();
    const int sub = threadIdx.x >> 6, co = threadIdx.x & 63;
    const long long s = (long long)blockIdx.x * 4 + sub;
    const bool active = s < p.S;
    // ---- load the slice (24 x F floats = 24 * F/4 float4; 64 threads)
    constexpr int BODIES = F / 4;                           // bodies per slice: 2 (pair model) or 1 (unconditional single-body model)
    if (co < 4) {
#pragma unroll
        for (int c = 0; c < F; ++c) { xs[sub][co < 2 ? co : 24 + co][c] = 0.f; }
    }
    if (active && co < 24 * BODIES) {
        const int h = co / BODIES, half = co - h * BODIES;
        float4 v;
        if (p.gather && BODIES == 2) {
            const int b = (int)(s % p.B);
            const int wp = (int)(s / p.B);
            const int pr = wp % p.P, kk = wp / p.P;
            int ii = 0, rem = pr;
            while (rem >= p.n - 1 - ii) { rem -= p.n - 1 - ii; ++ii; }
            const int jj = ii + 1 + rem;
            const int body = half ? jj : ii;
            v = reinterpret_cast<const float4*>(p.x)[((long long)b * p.T + kk * p.start + h) * p.n + body];
        } else if (p.gather) {
            // single-body slices of the EBM composition (reference gradient() :1866-1870): slice = body * B + b
            const int b = (int)(s % p.B), body = (int)(s / p.B);
            v = reinterpret_cast<const float4*>(p.x)[((long long)b * p.T + h) * p.n + body];
        } else {
            v = reinterpret_cast<const float4*>(p.x)[(s * 24 + h) * BODIES + half];
        }
        *reinterpret_cast<float4*>(&xs[sub][h + 2][half * 4]) = v;
    }
    // ---- weights of this output channel
    float w[5][F], wr[F];
#pragma unroll
    for (int k = 0; k < 5; ++k)
#pragma unroll
        for (int c = 0; c < F; ++c) w[k][c] = p.w0[(k * F + c) * 64 + co];
#pragma unroll
    for (int c = 0; c < F; ++c) wr[c] = p.wr[c * 64 + co];
    const float bias0 = p.b0[co], biasr = p.br[co];
    const float ga = p.gamma[co], be = p.beta[co];
    const float tb = p.t_dev ? p.tbias[(long long)(*p.t_dev) * 64 + co] : p.tbias[co];
    __syncthreads();
    if (!active) return;
    float y[24];
    float sum = 0.f;
#pragma unroll
    for (int h = 0; h < 24; ++h) {
        float a = bias0;
#pragma unroll
        for (int k = 0; k < 5; ++k)
#pragma unroll
            for (int c = 0; c < F; ++c) a = fmaf(w[k][c], xs[sub][h + k][c], a);
        y[h] = a;
        sum += a;
    }
    // GroupNorm group = 8 consecutive channels (lanes) x 24 positions
    sum += __shfl_xor_sync(0xffffffffu, sum, 1);
    sum += __shfl_xor_sync(0xffffffffu, sum, 2);
    sum += __shfl_xor_sync(0xffffffffu, sum, 4);
    const float mean = sum * (1.0f / 192.0f);
    float sq = 0.f;
#pragma unroll
    for (int h = 0; h < 24; ++h) { float d = y[h] - mean; sq = fmaf(d, d, sq); }
    sq += __shfl_xor_sync(0xffffffffu, sq, 1);
    sq += __shfl_xor_sync(0xffffffffu, sq, 2);
    sq += __shfl_xor_sync(0xffffffffu, sq, 4);
    const float rstd = rsqrtf(sq * (1.0f / 192.0f) + 1e-5f);
    T* ob = reinterpret_cast<T*>(p.out_b0) + s * 24 * 64 + co;
    T* orr = reinterpret_cast<T*>(p.out_res) + s * 24 * 64 + co;
#pragma unroll
    for (int h = 0; h < 24; ++h) {
        float v = (y[h] - mean) * rstd * ga + be;
        v = mish_fast(v) + tb;
        ob[h * 64] = from_f32<T>(v);
        float r = biasr;
#pragma unroll
        for (int c = 0; c < F; ++c) r = fmaf(wr[c], xs[sub][h + 2][c], r);
        orr[h * 64] = from_f32<T>(r);
    }
}

// ------------------------------------------------------------------------------------------
// Stem on the tensor cores (body-pair model, F = 8): the same block as stem_kernel, one WARP per slice.
//   conv0 as a GEMM  [24 (32) rows] x [K = 5 taps x 8 channels = 40 (48)] x [64]  with mma.sync.m16n8k16, the A fragments
//   read straight out of a 16-bit copy of the (zero-padded) slice: A[h][8 tap + c] = x[h + tap - 2][c];
//   the 1x1 residual conv as a second GEMM with K = 8 (16) on the centre tap.
// A GroupNorm group (8 channels x 24 positions) is exactly one n8 accumulator tile, so its statistics are a warp reduction.
// Outputs are staged in shared memory and written with 16-byte row-major stores.  The SIMT stem_kernel spent 1 300 FFMA /
// LDS instructions per (slice, channel) thread (ncu: issue-active 68 %, 162 M warp instructions per launch at 43 008 slices);
// here the contraction is 64 MMAs per slice.  Operands are rounded to the activation type (fp16 / bf16) like every other
// layer's; accumulation, GroupNorm and Mish stay fp32.
// ------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ void mma_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1);
template <>
__device__ __forceinline__ void mma_16816<__half>(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <>
__device__ __forceinline__ void mma_16816<__nv_bfloat16>(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <typename T>
__device__ __forceinline__ uint32_t pack_pair(float a, float b) {
    T lo = from_f32<T>(a), hi = from_f32<T>(b);
    return (uint32_t)(*reinterpret_cast<unsigned short*>(&lo)) | ((uint32_t)(*reinterpret_cast<unsigned short*>(&hi)) << 16);
}

constexpr int kStemWarps = 8;
constexpr int kStemXRows = 40;              // 2 zero rows + 24 positions + zero rows up to the last row a fragment can touch
constexpr int kStemStageRow = 72;           // halves per staged output row (144 B: 16-byte aligned, conflict-free)

template <typename T>
__global__ void __launch_bounds__(kStemWarps * 32, 2) stem_mma_kernel(StemParams p) {
    __shared__ __align__(16) T xs[kStemWarps][kStemXRows][8];
    __shared__ __align__(16) T stage[kStemWarps][24 * kStemStageRow];
    __shared__ uint2 bw[3][8][32];          // conv0 B fragments (b0, b1) of k-step ks, n tile nt, per lane
    __shared__ uint32_t br[8][32];          // residual conv B fragment b0 (its K rows 8..15 are zero)
    __shared__ float vbias[64], vgamma[64], vbeta[64], vtb[64], vbr[64];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t4 = lane & 3;
    pdl_wait();
    pdl_trigger();
    // ---- per-block set-up: fragments of the (rounded) weights, channel vectors, zero padding rows
    for (int i = threadIdx.x; i < 3 * 8 * 32; i += blockDim.x) {
        const int l = i & 31, nt = (i >> 5) & 7, ks = i >> 8;
        const int gg = l >> 2, tt = l & 3, co = nt * 8 + gg;
        auto wv = [&](int kk) { return kk < 40 ? p.w0[kk * 64 + co] : 0.f; };
        const int k0 = 16 * ks + 2 * tt;
        bw[ks][nt][l] = make_uint2(pack_pair<T>(wv(k0), wv(k0 + 1)), pack_pair<T>(wv(k0 + 8), wv(k0 + 9)));
    }
    for (int i = threadIdx.x; i < 8 * 32; i += blockDim.x) {
        const int l = i & 31, nt = i >> 5;
        const int gg = l >> 2, tt = l & 3, co = nt * 8 + gg;
        br[nt][l] = pack_pair<T>(p.wr[(2 * tt) * 64 + co], p.wr[(2 * tt + 1) * 64 + co]);
    }
    if (threadIdx.x < 64) {
        const int c = threadIdx.x;
        vbias[c] = p.b0[c]; vgamma[c] = p.gamma[c]; vbeta[c] = p.beta[c]; vbr[c] = p.br[c];
        vtb[c] = p.t_dev ? p.tbias[(long long)(*p.t_dev) * 64 + c] : p.tbias[c];
    }
    for (int i = lane; i < kStemXRows * 8; i += 32) (&xs[warp][0][0])[i] = from_f32<T>(0.f);
    __syncthreads();
    T* xw = &xs[warp][0][0];
    T* sw = &stage[warp][0];
    for (long long s = (long long)blockIdx.x * kStemWarps + warp; s < p.S; s += (long long)gridDim.x * kStemWarps) {
        // ---- the slice: 24 rows x 8 features, fp32 -> 16-bit, into rows 2..25 of the padded copy
        if (lane < 24) {
            const int h = lane;
            float4 v0, v1;
            if (p.gather) {
                const int b = (int)(s % p.B);
                const int wp = (int)(s / p.B);
                const int pr = wp % p.P, kk = wp / p.P;
                int ii = 0, rem = pr;
                while (rem >= p.n - 1 - ii) { rem -= p.n - 1 - ii; ++ii; }
                const int jj = ii + 1 + rem;
                const float4* row = reinterpret_cast<const float4*>(p.x) + ((long long)b * p.T + kk * p.start + h) * p.n;
                v0 = row[ii]; v1 = row[jj];
            } else {
                const float4* row = reinterpret_cast<const float4*>(p.x) + (s * 24 + h) * 2;
                v0 = row[0]; v1 = row[1];
            }
            *reinterpret_cast<uint4*>(xw + (h + 2) * 8) =
                make_uint4(pack_pair<T>(v0.x, v0.y), pack_pair<T>(v0.z, v0.w), pack_pair<T>(v1.x, v1.y), pack_pair<T>(v1.z, v1.w));
        }
        __syncwarp();
        // ---- A fragments: A[row][16 ks + k] = x[row + 2 ks (+1 for k >= 8)][k & 7]; rows = padded rows m + tap
        uint32_t af[2][3][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int ks = 0; ks < 3; ++ks) {
                const T* base = xw + (mt * 16 + g + 2 * ks) * 8 + 2 * t4;
                af[mt][ks][0] = *reinterpret_cast<const uint32_t*>(base);
                af[mt][ks][1] = *reinterpret_cast<const uint32_t*>(base + 8 * 8);
                af[mt][ks][2] = *reinterpret_cast<const uint32_t*>(base + 8);
                af[mt][ks][3] = *reinterpret_cast<const uint32_t*>(base + 9 * 8);
            }
        // ---- conv0 + GroupNorm + Mish + time bias, four n8 tiles (= four GroupNorm groups) at a time
#pragma unroll
        for (int nh = 0; nh < 2; ++nh) {
            float acc[2][4][4];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int q = 0; q < 4; ++q)
#pragma unroll
                    for (int c = 0; c < 4; ++c) acc[mt][q][c] = 0.f;
#pragma unroll
            for (int ks = 0; ks < 3; ++ks)
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const uint2 b = bw[ks][nh * 4 + q][lane];
#pragma unroll
                    for (int mt = 0; mt < 2; ++mt) mma_16816<T>(acc[mt][q], af[mt][ks], b.x, b.y);
                }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int nt = nh * 4 + q, c0 = nt * 8 + 2 * t4;
                const float2 bi = *reinterpret_cast<const float2*>(&vbias[c0]);
                // this thread's six valid outputs of the group: rows g, g + 8, 16 + g (row 24 + g is padding), two channels
                float y[6] = {acc[0][q][0] + bi.x, acc[0][q][1] + bi.y, acc[0][q][2] + bi.x, acc[0][q][3] + bi.y,
                              acc[1][q][0] + bi.x, acc[1][q][1] + bi.y};
                float sum = ((y[0] + y[1]) + (y[2] + y[3])) + (y[4] + y[5]);
                sum = warp_sum(sum);
                const float mean = sum * (1.0f / 192.0f);
                float sq = 0.f;
#pragma unroll
                for (int i = 0; i < 6; ++i) { const float d = y[i] - mean; sq = fmaf(d, d, sq); }
                sq = warp_sum(sq);
                const float rstd = rsqrtf(sq * (1.0f / 192.0f) + 1e-5f);
                const float2 ga = *reinterpret_cast<const float2*>(&vgamma[c0]), be = *reinterpret_cast<const float2*>(&vbeta[c0]);
                const float2 tb = *reinterpret_cast<const float2*>(&vtb[c0]);
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    const int row = i == 0 ? g : (i == 1 ? g + 8 : 16 + g);
                    const float v0 = mish_fast((y[2 * i] - mean) * rstd * ga.x + be.x) + tb.x;
                    const float v1 = mish_fast((y[2 * i + 1] - mean) * rstd * ga.y + be.y) + tb.y;
                    *reinterpret_cast<uint32_t*>(sw + row * kStemStageRow + c0) = pack_pair<T>(v0, v1);
                }
            }
        }
        __syncwarp();
        {
            T* dst = reinterpret_cast<T*>(p.out_b0) + s * 24 * 64;
#pragma unroll
            for (int i = 0; i < 6; ++i) {
                const int idx = i * 32 + lane, row = idx >> 3, c8 = idx & 7;
                *reinterpret_cast<uint4*>(dst + row * 64 + c8 * 8) = *reinterpret_cast<const uint4*>(sw + row * kStemStageRow + c8 * 8);
            }
        }
        __syncwarp();
        // ---- residual 1x1 conv on the centre tap (K = 8 channels, padded to 16 with zeros)
#pragma unroll
        for (int nh = 0; nh < 2; ++nh) {
            float acc[2][4][4];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int q = 0; q < 4; ++q)
#pragma unroll
                    for (int c = 0; c < 4; ++c) acc[mt][q][c] = 0.f;
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                // centre tap = k-step 1, low half (tap 2): the fragment registers already loaded for conv0
                const uint32_t ar[4] = {af[mt][1][0], af[mt][1][1], 0u, 0u};
#pragma unroll
                for (int q = 0; q < 4; ++q) mma_16816<T>(acc[mt][q], ar, br[nh * 4 + q][lane], 0u);
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int c0 = (nh * 4 + q) * 8 + 2 * t4;
                const float2 bi = *reinterpret_cast<const float2*>(&vbr[c0]);
                *reinterpret_cast<uint32_t*>(sw + g * kStemStageRow + c0) = pack_pair<T>(acc[0][q][0] + bi.x, acc[0][q][1] + bi.y);
                *reinterpret_cast<uint32_t*>(sw + (g + 8) * kStemStageRow + c0) = pack_pair<T>(acc[0][q][2] + bi.x, acc[0][q][3] + bi.y);
                *reinterpret_cast<uint32_t*>(sw + (16 + g) * kStemStageRow + c0) = pack_pair<T>(acc[1][q][0] + bi.x, acc[1][q][1] + bi.y);
            }
        }
        __syncwarp();
        {
            T* dst = reinterpret_cast<T*>(p.out_res) + s * 24 * 64;
#pragma unroll
            for (int i = 0; i < 6; ++i) {
                const int idx = i * 32 + lane, row = idx >> 3, c8 = idx & 7;
                *reinterpret_cast<uint4*>(dst + row * 64 + c8 * 8) = *reinterpret_cast<const uint4*>(sw + row * kStemStageRow + c8 * 8);
            }
        }
        __syncwarp();
    }
}

int launch_stem(const StemLaunch& a, cudaStream_t st) {
    if (a.S == 0) return 0;
    const int F = a.conv0->cin;
    if (F != 8 && F != 4) return fail(-2, "stem kernel is built for 8 (body pair) or 4 (single body) input features");
    KernelTimer kt("stem", st, (double)a.S * 24 * (4.0 * F + 2.0 * 64 * elem_size(a.prec)));
    StemParams p;
    p.x = a.x; p.w0 = a.conv0->w; p.b0 = a.conv0->bias; p.gamma = a.gn->gamma; p.beta = a.gn->beta;
    p.tbias = a.tbias; p.t_dev = a.t_dev; p.wr = a.res->w; p.br = a.res->bias;
    p.out_b0 = a.out_b0; p.out_res = a.out_res; p.S = a.S;
    p.gather = a.gather; p.B = a.B; p.n = a.n; p.P = a.n * (a.n - 1) / 2; p.start = a.start; p.T = a.T;
    static int stem_simt = -1;
    if (stem_simt < 0) { const char* e = getenv("CINDM_STEM_SIMT"); stem_simt = (e && e[0] == '1') ? 1 : 0; }
    if (F == 8 && !stem_simt) {
        if (a.prec != PREC_F16 && a.prec != PREC_BF16) return fail(-2, "stem kernel is built for the 16-bit precisions");
        long long need = (a.S + kStemWarps - 1) / kStemWarps;
        const unsigned grid = (unsigned)(need < 148LL * 8 ? need : 148LL * 8);
        if (a.prec == PREC_F16) CINDM_CHECK_CUDA(launch_chain(stem_mma_kernel<__half>, dim3(grid), dim3(kStemWarps * 32), 0, st, p));
        else CINDM_CHECK_CUDA(launch_chain(stem_mma_kernel<__nv_bfloat16>, dim3(grid), dim3(kStemWarps * 32), 0, st, p));
        CINDM_CHECK_LAUNCH();
        return 0;
    }
    const unsigned blocks = (unsigned)((a.S + 3) / 4);
    if (a.prec == PREC_F16 && F == 8) CINDM_CHECK_CUDA(launch_chain(stem_kernel<__half, 8>, dim3(blocks), dim3(256), 0, st, p));
    else if (a.prec == PREC_F16) CINDM_CHECK_CUDA(launch_chain(stem_kernel<__half, 4>, dim3(blocks), dim3(256), 0, st, p));
    else if (a.prec == PREC_BF16 && F == 8) CINDM_CHECK_CUDA(launch_chain(stem_kernel<__nv_bfloat16, 8>, dim3(blocks), dim3(256), 0, st, p));
    else if (a.prec == PREC_BF16) CINDM_CHECK_CUDA(launch_chain(stem_kernel<__nv_bfloat16, 4>, dim3(blocks), dim3(256), 0, st, p));
    else return fail(-2, "stem kernel is built for the 16-bit precisions");
    CINDM_CHECK_LAUNCH();
    return 0;
}

// ------------------------------------------------------------------------------------------
// Head: final 1x1 conv 64 -> 8 (+bias) from 16-bit activations to fp32 eps_pair [S][24][8]
// (reference final_conv[1], model/diffusion_1d.py:607).  One thread per (slice, position) row.
// ------------------------------------------------------------------------------------------
constexpr int kHeadRows = 512;                 // rows per block: two per thread, so every broadcast weight load feeds two rows
constexpr int kHeadRowHalves = 64 + 8;         // row stride 144 B: the per-thread 16-byte reads of the second phase are conflict-free

template <typename T, int F>
__global__ void __launch_bounds__(256) head_kernel(const T* __restrict__ in, const float* __restrict__ w,
                                                   const float* __restrict__ bias, float* __restrict__ out,
                                                   long long rows) {
    // The block's 512 input rows (128 B each) are staged in shared memory with fully coalesced 16-byte loads (round 1 let
    // every thread walk its own row: 32 different 128-byte lines per load instruction, 1.9 TB/s).  The FMA phase is bound by
    // the shared-memory loads of the weights (two 16-byte broadcast loads per input channel), so a thread takes TWO rows per
    // weight load, on packed fp32 FMAs (fma.f32x2: the same IEEE operations, half the issue slots).
    extern __shared__ __align__(16) unsigned char head_smem[];
    T* tile = reinterpret_cast<T*>(head_smem);
    __shared__ float sw[64][F];
    __shared__ float sb[F];
    pdl_wait();
    pdl_trigger();
    for (int i = threadIdx.x; i < 64 * F; i += 256) sw[i / F][i % F] = w[i];      // w: [1][64][F]
    if (threadIdx.x < F) sb[threadIdx.x] = bias[threadIdx.x];
    const long long row0 = (long long)blockIdx.x * kHeadRows;
    const long long n_here = rows - row0 < kHeadRows ? rows - row0 : kHeadRows;
    const uint4* src = reinterpret_cast<const uint4*>(in + row0 * 64);
#pragma unroll
    for (int it = 0; it < kHeadRows * 8 / 256; ++it) {
        const int v = it * 256 + threadIdx.x;                     // 16-byte vector index inside the block's rows
        const int r = v >> 3, c = v & 7;
        if (r < n_here) *reinterpret_cast<uint4*>(tile + r * kHeadRowHalves + c * 8) = src[v];
    }
    __syncthreads();
    if (threadIdx.x >= n_here) return;
    const bool two = threadIdx.x + 256 < n_here;                  // (the second row of the last block may not exist)
    unsigned long long acc_a[F / 2], acc_b[F / 2];
#pragma unroll
    for (int o = 0; o < F; o += 2) acc_a[o >> 1] = acc_b[o >> 1] = f32x2_pack(sb[o], sb[o + 1]);
    const T* row_a = tile + threadIdx.x * kHeadRowHalves;
    const T* row_b = tile + (two ? threadIdx.x + 256 : threadIdx.x) * kHeadRowHalves;
#pragma unroll
    for (int c8 = 0; c8 < 64; c8 += 8) {
        float va[8], vb[8];
        load8<T>(row_a + c8, va);
        load8<T>(row_b + c8, vb);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const unsigned long long a2 = f32x2_pack(va[k], va[k]), b2 = f32x2_pack(vb[k], vb[k]);
#pragma unroll
            for (int o = 0; o < F; o += 2) {
                const unsigned long long w2 = f32x2_pack(sw[c8 + k][o], sw[c8 + k][o + 1]);
                acc_a[o >> 1] = f32x2_fma(a2, w2, acc_a[o >> 1]);
                acc_b[o >> 1] = f32x2_fma(b2, w2, acc_b[o >> 1]);
            }
        }
    }
    float acc[F];
#pragma unroll
    for (int o = 0; o < F; o += 2) f32x2_unpack(acc_a[o >> 1], acc[o], acc[o + 1]);
    float4* dst = reinterpret_cast<float4*>(out + (row0 + threadIdx.x) * F);
#pragma unroll
    for (int o = 0; o < F; o += 4) dst[o >> 2] = make_float4(acc[o], acc[o + 1], acc[o + 2], acc[o + 3]);
    if (two) {
#pragma unroll
        for (int o = 0; o < F; o += 2) f32x2_unpack(acc_b[o >> 1], acc[o], acc[o + 1]);
        dst = reinterpret_cast<float4*>(out + (row0 + threadIdx.x + 256) * F);
#pragma unroll
        for (int o = 0; o < F; o += 4) dst[o >> 2] = make_float4(acc[o], acc[o + 1], acc[o + 2], acc[o + 3]);
    }
}

int launch_head(const void* in, const ConvW& w, float* out, int64_t rows, int prec, cudaStream_t st) {
    if (rows == 0) return 0;
    const int F = w.cout;
    if (F != 8 && F != 4) return fail(-2, "head kernel is built for 8 or 4 output features");
    KernelTimer kt("head", st, (double)rows * (64.0 * elem_size(prec) + 4.0 * F));
    const unsigned blocks = (unsigned)((rows + kHeadRows - 1) / kHeadRows);
    const size_t smem = (size_t)kHeadRows * kHeadRowHalves * 2;
    if (prec != PREC_F16 && prec != PREC_BF16) return fail(-2, "head kernel is built for the 16-bit precisions");
#define CINDM_HEAD(T, FF)                                                                                                     \
    do {                                                                                                                  \
        static DeviceOnce once;                                                                                           \
        if (once.first_time())                                                                                            \
            CINDM_CHECK_CUDA(cudaFuncSetAttribute(head_kernel<T, FF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        CINDM_CHECK_CUDA(launch_chain(head_kernel<T, FF>, dim3(blocks), dim3(256), smem, st, (const T*)in, (const float*)w.w,   \
                                      (const float*)w.bias, out, (long long)rows));                                        \
    } while (0)
    if (prec == PREC_F16 && F == 8) CINDM_HEAD(__half, 8);
    else if (prec == PREC_F16) CINDM_HEAD(__half, 4);
    else if (F == 8) CINDM_HEAD(__nv_bfloat16, 8);
    else CINDM_HEAD(__nv_bfloat16, 4);
#undef CINDM_HEAD
    CINDM_CHECK_LAUNCH();
    return 0;
}

}  // namespace cindm
