// tcgen05 / TMEM / TMA implicit-GEMM convolution for 16-bit activations (see conv_tc.cu).
#pragma once
#include "engine.h"

namespace cindm {

// EPI_GN_MISH_T3: k=5 conv at H=3 as ONE dense GEMM [S, 3*Cin] x [3*Cin, 3*Cout] (block-Toeplitz weights, no
// zero taps: 9 instead of 15 tap-blocks), rows = slices, columns ordered (group, position, channel) so that a
// GroupNorm group is 192 / 96 consecutive accumulator columns of one row.
enum { EPI_BIAS = 0, EPI_GN_MISH = 1, EPI_GN_MISH_T3 = 2 };
// TC_SAME: k in {1,5}, stride 1, pad k/2 (H -> H).  TC_DOWN: Downsample1d k=3 s=2 p=1 (H -> H/2).
// TC_UP: Upsample1d ConvTranspose k=4 s=2 p=1 (H -> 2H), run as two 2-tap GEMMs (even / odd outputs).
enum { TC_SAME = 0, TC_DOWN = 1, TC_UP = 2 };

struct ConvTcLaunch {
    const void* in0 = nullptr; int c0 = 0;      // [S][H][c0] 16-bit
    const void* in1 = nullptr; int c1 = 0;      // optional channel-concatenated second input
    const ConvW* w = nullptr;                   // taps in {1, 5}, stride 1, pad taps/2
    const NormW* gn = nullptr;                  // EPI_GN_MISH
    const float* add_vec = nullptr;             // per-channel vector added after the activation
    const int* t_dev = nullptr;                 // if set, add_vec is a [timesteps][cout] table indexed by *t_dev
    const void* add_res = nullptr;              // residual tensor [S][H][cout] added last
    void* out = nullptr;                        // [S][H][cout] 16-bit
    int64_t S = 0;
    int H = 0;                                  // input positions per slice
    int mode = TC_SAME;
    int prec = PREC_F16;
    int epilogue = EPI_BIAS;
    // walk the tiles from the last one down: consecutive layers alternate, so that a layer starts on the part of its input the
    // previous layer wrote LAST (still in L2) instead of the part written first (long evicted: a tensor is 132 MB, L2 126 MB)
    int reverse = 0;
    // small batches: the two parity GEMMs of a transposed conv are independent; when `side` is set the odd outputs are
    // computed on it (event fork / join around it, graph-capturable)
    cudaStream_t side = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
};

int launch_conv_tc(const ConvTcLaunch& a, cudaStream_t st);

// conv_tc_cm.cu: the k = 5 GroupNorm + Mish convs at 64 / 128 (/ 256) channels with a channel-major accumulator
// (output channels on the TMEM lanes).  launch_conv_tc routes the eligible layers there.
bool conv_tc_cm_eligible(const ConvTcLaunch& a);
int launch_conv_tc_cm(const ConvTcLaunch& a, cudaStream_t st);

}  // namespace cindm
