// Fused multi-head self-attention core, softmax(q^T k / sqrt(d)) v, head dim 64 (models/unet.py:84-97).
// Flash-style: the N x N score matrix never leaves the SM.  Tensor-core math via mma.sync m16n8k16 (bf16 in,
// fp32 accumulate); online softmax in fp32 with exp2.  qkv is the NHWC output of the 1x1 qkv conv:
// [B, N, 3C] with q = channels [0,C), k = [C,2C), v = [2C,3C); head h owns the contiguous 64-channel block h*64
// inside each third (the reference's chunk(3, dim=1) + view(B, heads, C/heads, HW)).
// grid (N/64, heads, B), 128 threads: 4 warps x 16 queries; keys/values streamed in blocks of 64.
#pragma once
#include "common.cuh"

namespace rfv {

template <int D>  // head dim: 32, 64 or 128
__global__ void __launch_bounds__(128) attn_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ out, int N, int C,
                                                   float scale_log2) {
    constexpr int ATT_LD = D + 8;  // padded row length (bf16) -> conflict-free ldmatrix
    constexpr int KS = D / 16;     // k16 steps over the head dim
    __shared__ __align__(16) bf16 Qs[64][ATT_LD];
    __shared__ __align__(16) bf16 Ks[64][ATT_LD];
    __shared__ __align__(16) bf16 Vs[64][ATT_LD];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int qb = blockIdx.x, head = blockIdx.y, b = blockIdx.z;
    const size_t row_stride = (size_t)3 * C;
    const bf16* base = qkv + (size_t)b * N * row_stride + head * D;

    auto load_tile = [&](bf16(*dst)[ATT_LD], const bf16* src, int row0) {
#pragma unroll
        for (int i = 0; i < D / 16; ++i) {
            const int id = tid + i * 128, r = id / (D / 8), c = id % (D / 8);
            cp_async16(smem_u32(&dst[r][c * 8]), src + (size_t)(row0 + r) * row_stride + c * 8, true);
        }
    };
    load_tile(Qs, base, qb * 64);
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    uint32_t qf[KS][4];
#pragma unroll
    for (int kk = 0; kk < KS; ++kk)
        ldmatrix_x4(smem_u32(&Qs[warp * 16 + (lane & 15)][kk * 16 + (lane >> 4) * 8]), qf[kk][0], qf[kk][1], qf[kk][2], qf[kk][3]);

    float o[D / 8][4];
#pragma unroll
    for (int i = 0; i < D / 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) o[i][j] = 0.f;
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;

    for (int kb = 0; kb < N / 64; ++kb) {
        __syncthreads();  // previous K/V tile fully consumed
        load_tile(Ks, base + C, kb * 64);
        load_tile(Vs, base + 2 * C, kb * 64);
        cp_async_commit();
        cp_async_wait<0>();
        __syncthreads();

        float s[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
#pragma unroll
        for (int kk = 0; kk < KS; ++kk)
#pragma unroll
            for (int np = 0; np < 4; ++np) {
                uint32_t b0, b1, b2, b3;
                ldmatrix_x4(smem_u32(&Ks[np * 16 + (lane & 7) + ((lane >> 4) << 3)][kk * 16 + ((lane >> 3) & 1) * 8]), b0, b1, b2, b3);
                mma_bf16_16816(s[2 * np], qf[kk], b0, b1);
                mma_bf16_16816(s[2 * np + 1], qf[kk], b2, b3);
            }
        // online softmax (rows g and g+8 of this warp's 16 queries)
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            mx0 = fmaxf(mx0, fmaxf(s[i][0], s[i][1]));
            mx1 = fmaxf(mx1, fmaxf(s[i][2], s[i][3]));
        }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
        const float mn0 = fmaxf(m0, mx0 * scale_log2), mn1 = fmaxf(m1, mx1 * scale_log2);
        const float a0 = exp2f(m0 - mn0), a1 = exp2f(m1 - mn1);
        m0 = mn0;
        m1 = mn1;
        l0 *= a0;
        l1 *= a1;
#pragma unroll
        for (int i = 0; i < D / 8; ++i) {
            o[i][0] *= a0; o[i][1] *= a0; o[i][2] *= a1; o[i][3] *= a1;
        }
        uint32_t pf[4][4];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float p0 = exp2f(s[i][0] * scale_log2 - mn0), p1 = exp2f(s[i][1] * scale_log2 - mn0);
            const float p2 = exp2f(s[i][2] * scale_log2 - mn1), p3 = exp2f(s[i][3] * scale_log2 - mn1);
            l0 += p0 + p1;
            l1 += p2 + p3;
            pf[i >> 1][(i & 1) * 2] = pack_bf16x2(p0, p1);
            pf[i >> 1][(i & 1) * 2 + 1] = pack_bf16x2(p2, p3);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)      // key k16-steps
#pragma unroll
            for (int np = 0; np < D / 16; ++np) {  // pairs of d n-tiles
                uint32_t b0, b1, b2, b3;
                ldmatrix_x4_trans(smem_u32(&Vs[j * 16 + (lane & 7) + ((lane >> 3) & 1) * 8][np * 16 + (lane >> 4) * 8]), b0, b1, b2, b3);
                mma_bf16_16816(o[2 * np], pf[j], b0, b1);
                mma_bf16_16816(o[2 * np + 1], pf[j], b2, b3);
            }
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float i0 = 1.0f / l0, i1 = 1.0f / l1;
    const int g = lane >> 2, tq = lane & 3;
    const size_t r0 = (size_t)b * N + qb * 64 + warp * 16 + g;
#pragma unroll
    for (int i = 0; i < D / 8; ++i) {
        const int c = head * D + i * 8 + tq * 2;
        *reinterpret_cast<uint32_t*>(out + r0 * C + c) = pack_bf16x2(o[i][0] * i0, o[i][1] * i0);
        *reinterpret_cast<uint32_t*>(out + (r0 + 8) * C + c) = pack_bf16x2(o[i][2] * i1, o[i][3] * i1);
    }
}

}  // namespace rfv
