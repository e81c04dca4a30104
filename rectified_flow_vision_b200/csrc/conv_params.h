// Parameters shared by the two implicit-GEMM convolution kernels (tcgen05 and mma.sync).
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

namespace rfv {

// out[n,ho,wo,co] = sum_{tap,c} A0[n, ho*stride+dy, wo*stride+dx, c] * W[co, tap*C0 + c]          (segment 0: ks x ks)
//                 + sum_c  S1[n,ho,wo,c] * W[co, K0 + c]                                          (segment 1: 1x1 shortcut)
//                 + (temb ? temb[n*temb_stride + co] : bias[co]) + resid[n,ho,wo,co]
//                   (when a time projection is present the conv bias is already folded into it)
// All activations NHWC bf16.  Segment 1 may read a virtual channel-concat of two tensors (s1a | s1b).
// `ups`: segment 0 input is nearest-upsampled x2 on the fly (logical input = 2*H0 x 2*W0).
// GroupNorm partial statistics of the fp32 result are accumulated per (n, 8*k-channel slab):
//   stats[(n * (Cout >> slab_shift) + (co >> slab_shift)) * 2 + {0: sum, 1: sum of squares}].
struct ConvParams {
    __nv_bfloat16* out;
    const __nv_bfloat16* a0;
    const __nv_bfloat16* s1a;
    const __nv_bfloat16* s1b;
    const __nv_bfloat16* w;      // [Cout][Ktot], K-major
    const float* bias;           // [Cout]
    const float* temb;           // or nullptr
    const __nv_bfloat16* resid;  // or nullptr
    float* stats;                // or nullptr
    int B, Ho, Wo, Cout;
    int H0, W0, C0;              // physical dims of a0
    int ks, stride, ups;
    int C1a, C1b;
    int K0, Ktot;
    int temb_stride;
    int slab_shift;
    // fused GroupNorm(+SiLU) on the segment-0 operand (conv_wa.cuh, FUSE): per-(image, channel) scale / shift pairs
    const float* gn_coef;        // [B][gn_C][2] or nullptr
    int gn_C;
    int gn_silu;
};

}  // namespace rfv
