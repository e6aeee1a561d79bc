// Shared declarations for the cindm_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <atomic>
#include <string>
#include <utility>

namespace cindm {

// Storage / operand precision of the U-Net activations.  Accumulation is always fp32.
enum Precision { PREC_F32 = 0, PREC_F16 = 1, PREC_BF16 = 2 };

void set_error(const std::string& msg);          // capi.cu; message returned by cindm_last_error()
int fail(int code, const std::string& msg);      // records msg, returns code

#define CINDM_CHECK_CUDA(expr)                                                              \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess)                                                              \
            return ::cindm::fail(-100, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
    } while (0)

#define CINDM_CHECK_LAUNCH()                                                                      \
    do {                                                                                          \
        ::cindm::count_launch();                                                                  \
        cudaError_t _e = cudaGetLastError();                                                      \
        if (_e != cudaSuccess)                                                                    \
            return ::cindm::fail(-101, std::string("kernel launch (") + __FILE__ + ":" +          \
                                           std::to_string(__LINE__) + "): " + cudaGetErrorString(_e)); \
    } while (0)

// cudaFuncSetAttribute is per device.  A launcher keeps `static DeviceOnce once;` and configures its kernel the first
// time it runs on each device (one process may drive several GPUs through several engine handles).
struct DeviceOnce {
    std::atomic<unsigned long long> done{0};
    bool first_time() {
        int d = 0;
        cudaGetDevice(&d);
        if (d < 0 || d >= 64) return true;
        const unsigned long long bit = 1ull << d;
        return (done.fetch_or(bit) & bit) == 0;
    }
};

// ---- programmatic dependent launch (PDL) ------------------------------------------------------------------
// The kernels of one evaluation form a chain.  Launched with programmaticStreamSerializationAllowed, kernel k+1 may
// become resident while kernel k is still running (on SMs k does not use - small batches - or has already left -
// the tail of a persistent kernel) and do the work that depends on nothing: barrier init, TMEM allocation, tensor-map
// prefetch.  It then blocks in pdl_wait() until kernel k has COMPLETED and its memory is visible; every global
// access of every such kernel comes after that call.  pdl_trigger() lets the kernel after it be scheduled.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
bool pdl_enabled();        // CINDM_PDL=0 turns the launch attribute off (the device calls are then no-ops)

template <typename... KArgs, typename... Args>
cudaError_t launch_chain(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

// ---- built-in tracing: launches counted always; per-kernel-class CUDA-event timing when enabled
void count_launch();
void add_launches(long long n);
long long launch_count();
bool profiling_enabled();
void profile_record(const char* tag, cudaStream_t st, bool begin, double work);

// RAII: brackets the kernel launches of one launcher with two events on the launching stream.
struct KernelTimer {
    cudaStream_t st; bool on;
    KernelTimer(const char* t, cudaStream_t s, double work = 0.0) : st(s), on(profiling_enabled()) {
        if (on) profile_record(t, st, true, work);
    }
    ~KernelTimer() { if (on) profile_record("", st, false, 0.0); }
};

#define CINDM_TRY(expr)            \
    do {                           \
        int _rc = (expr);          \
        if (_rc != 0) return _rc;  \
    } while (0)

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __half from_f32<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// x * tanh(softplus(x)) exactly as torch evaluates nn.Mish in fp32 (reference model/diffusion_1d.py:210).
__device__ __forceinline__ float mish_exact(float x) { return x * tanhf(log1pf(expf(x))); }

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// Same function with one ex2 and one rcp: with w = e^x, tanh(log(1+w)) = (w^2+2w)/(w^2+2w+2) = 1 - 2/(w^2+2w+2), so
// mish(x) = x - 2x / (w(w+2) + 2).  No clamp is needed: for large x the denominator overflows to +inf, its
// reciprocal is 0 and the result is x; for very negative x, w = 0 and the result is x - x = 0.
__device__ __forceinline__ float mish_fast(float x) {
    float w, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(w) : "f"(x * 1.4426950408889634f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(fmaf(w, w + 2.0f, 2.0f)));
    return fmaf(x * r, -2.0f, x);
}

// Packed fp32 FMA (sm_100: fma.rn.f32x2 -> FFMA2): two IEEE fp32 FMAs on a 64-bit register pair per instruction.  Measured on B200
// (profiles/micro/ffma2_throughput.cu): the same 126 lane-FMA/clk/SM as scalar FFMA at HALF the instruction rate, i.e. it frees
// every second issue slot of an FMA-bound loop for the shared-memory loads.
__device__ __forceinline__ unsigned long long f32x2_pack(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void f32x2_unpack(unsigned long long v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long f32x2_fma(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ unsigned long long f32x2_mul(unsigned long long a, unsigned long long b) {
    unsigned long long d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ unsigned long long f32x2_add(unsigned long long a, unsigned long long b) {
    unsigned long long d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

}  // namespace cindm
