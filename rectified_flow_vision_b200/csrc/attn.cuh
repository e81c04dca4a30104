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
                                                   float scale_log2, float* __restrict__ lse) {
    pdl_wait();   // programmatic dependent launch (common.cuh): no-ops unless launched with the attribute
    pdl_launch();
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
    if (lse && tq == 0) {  // log2-domain log-sum-exp per query row: the backward pass rebuilds P = exp2(s*scale_log2 - lse)
        float* l = lse + ((size_t)b * gridDim.y + head) * N + qb * 64 + warp * 16 + g;
        l[0] = m0 + log2f(l0);
        l[8] = m1 + log2f(l1);
    }
#pragma unroll
    for (int i = 0; i < D / 8; ++i) {
        const int c = head * D + i * 8 + tq * 2;
        *reinterpret_cast<uint32_t*>(out + r0 * C + c) = pack_bf16x2(o[i][0] * i0, o[i][1] * i0);
        *reinterpret_cast<uint32_t*>(out + (r0 + 8) * C + c) = pack_bf16x2(o[i][2] * i1, o[i][3] * i1);
    }
}

// ---------------------------------------------------------------------------------------------------------
// Backward of the attention core.  With S = scale * Q K^T, P = softmax(S), O = P V and delta_i = dO_i . O_i:
//   dV = P^T dO      dP = dO V^T      dS = P o (dP - delta)      dQ = scale * dS K      dK = scale * dS^T Q
// Two kernels in the forward kernel's fragment conventions, P recomputed from the saved log-sum-exp:
//   attn_bwd_dq_kernel   grid (N/64 query blocks, heads, B): streams key/value blocks, also writes delta
//   attn_bwd_dkv_kernel  grid (N/64 key blocks,   heads, B): streams query blocks (everything transposed)
// dqkv has the layout of qkv ([B, N, 3C]); o / dout are [B, N, C].
// ---------------------------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(128) attn_bwd_dq_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ o,
                                                          const bf16* __restrict__ dout, const float* __restrict__ lse,
                                                          float* __restrict__ delta, bf16* __restrict__ dqkv, int N, int C,
                                                          float scale, float scale_log2) {
    constexpr int ATT_LD = D + 8;
    constexpr int KS = D / 16;
    __shared__ __align__(16) bf16 Qs[64][ATT_LD];
    __shared__ __align__(16) bf16 Gs[64][ATT_LD];   // dO rows of this query block
    __shared__ __align__(16) bf16 Ks[64][ATT_LD];
    __shared__ __align__(16) bf16 Vs[64][ATT_LD];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int qb = blockIdx.x, head = blockIdx.y, b = blockIdx.z;
    const size_t rs3 = (size_t)3 * C;
    const bf16* base = qkv + (size_t)b * N * rs3 + head * D;
    auto load_tile = [&](bf16(*dst)[ATT_LD], const bf16* src, size_t row_stride, int row0) {
#pragma unroll
        for (int i = 0; i < D / 16; ++i) {
            const int id = tid + i * 128, r = id / (D / 8), c = id % (D / 8);
            cp_async16(smem_u32(&dst[r][c * 8]), src + (size_t)(row0 + r) * row_stride + c * 8, true);
        }
    };
    load_tile(Qs, base, rs3, qb * 64);
    load_tile(Gs, dout + (size_t)b * N * C + head * D, (size_t)C, qb * 64);
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    uint32_t qf[KS][4], gf[KS][4];
#pragma unroll
    for (int kk = 0; kk < KS; ++kk) {
        ldmatrix_x4(smem_u32(&Qs[warp * 16 + (lane & 15)][kk * 16 + (lane >> 4) * 8]), qf[kk][0], qf[kk][1], qf[kk][2], qf[kk][3]);
        ldmatrix_x4(smem_u32(&Gs[warp * 16 + (lane & 15)][kk * 16 + (lane >> 4) * 8]), gf[kk][0], gf[kk][1], gf[kk][2], gf[kk][3]);
    }
    const int g = lane >> 2, tq = lane & 3;
    const size_t r0 = (size_t)b * N + qb * 64 + warp * 16 + g;
    float d0 = 0.f, d1 = 0.f;
#pragma unroll
    for (int i = 0; i < D / 8; ++i) {
        const int c = head * D + i * 8 + tq * 2;
        const float2 a0 = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(o + r0 * C + c));
        const float2 a1 = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(o + (r0 + 8) * C + c));
        const float2 b0 = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(dout + r0 * C + c));
        const float2 b1 = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(dout + (r0 + 8) * C + c));
        d0 += a0.x * b0.x + a0.y * b0.y;
        d1 += a1.x * b1.x + a1.y * b1.y;
    }
    d0 += __shfl_xor_sync(0xffffffffu, d0, 1); d0 += __shfl_xor_sync(0xffffffffu, d0, 2);
    d1 += __shfl_xor_sync(0xffffffffu, d1, 1); d1 += __shfl_xor_sync(0xffffffffu, d1, 2);
    const size_t lrow = ((size_t)b * gridDim.y + head) * N + qb * 64 + warp * 16 + g;
    if (tq == 0) { delta[lrow] = d0; delta[lrow + 8] = d1; }
    const float l0 = lse[lrow], l1 = lse[lrow + 8];

    float dq[D / 8][4];
#pragma unroll
    for (int i = 0; i < D / 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dq[i][j] = 0.f;
    for (int kb = 0; kb < N / 64; ++kb) {
        __syncthreads();
        load_tile(Ks, base + C, rs3, kb * 64);
        load_tile(Vs, base + 2 * C, rs3, kb * 64);
        cp_async_commit();
        cp_async_wait<0>();
        __syncthreads();
        float s[8][4], dp[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) { s[i][j] = 0.f; dp[i][j] = 0.f; }
#pragma unroll
        for (int kk = 0; kk < KS; ++kk)
#pragma unroll
            for (int np = 0; np < 4; ++np) {
                uint32_t b0, b1, b2, b3;
                ldmatrix_x4(smem_u32(&Ks[np * 16 + (lane & 7) + ((lane >> 4) << 3)][kk * 16 + ((lane >> 3) & 1) * 8]), b0, b1, b2, b3);
                mma_bf16_16816(s[2 * np], qf[kk], b0, b1);
                mma_bf16_16816(s[2 * np + 1], qf[kk], b2, b3);
                ldmatrix_x4(smem_u32(&Vs[np * 16 + (lane & 7) + ((lane >> 4) << 3)][kk * 16 + ((lane >> 3) & 1) * 8]), b0, b1, b2, b3);
                mma_bf16_16816(dp[2 * np], gf[kk], b0, b1);
                mma_bf16_16816(dp[2 * np + 1], gf[kk], b2, b3);
            }
        uint32_t dsf[4][4];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float p0 = exp2f(s[i][0] * scale_log2 - l0), p1 = exp2f(s[i][1] * scale_log2 - l0);
            const float p2 = exp2f(s[i][2] * scale_log2 - l1), p3 = exp2f(s[i][3] * scale_log2 - l1);
            dsf[i >> 1][(i & 1) * 2] = pack_bf16x2(p0 * (dp[i][0] - d0), p1 * (dp[i][1] - d0));
            dsf[i >> 1][(i & 1) * 2 + 1] = pack_bf16x2(p2 * (dp[i][2] - d1), p3 * (dp[i][3] - d1));
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int np = 0; np < D / 16; ++np) {
                uint32_t b0, b1, b2, b3;
                ldmatrix_x4_trans(smem_u32(&Ks[j * 16 + (lane & 7) + ((lane >> 3) & 1) * 8][np * 16 + (lane >> 4) * 8]), b0, b1, b2, b3);
                mma_bf16_16816(dq[2 * np], dsf[j], b0, b1);
                mma_bf16_16816(dq[2 * np + 1], dsf[j], b2, b3);
            }
    }
#pragma unroll
    for (int i = 0; i < D / 8; ++i) {
        const int c = head * D + i * 8 + tq * 2;
        *reinterpret_cast<uint32_t*>(dqkv + r0 * rs3 + c) = pack_bf16x2(dq[i][0] * scale, dq[i][1] * scale);
        *reinterpret_cast<uint32_t*>(dqkv + (r0 + 8) * rs3 + c) = pack_bf16x2(dq[i][2] * scale, dq[i][3] * scale);
    }
}

template <int D>
__global__ void __launch_bounds__(128) attn_bwd_dkv_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ dout,
                                                           const float* __restrict__ lse, const float* __restrict__ delta,
                                                           bf16* __restrict__ dqkv, int N, int C, float scale, float scale_log2) {
    constexpr int ATT_LD = D + 8;
    constexpr int KS = D / 16;
    __shared__ __align__(16) bf16 Ks[64][ATT_LD];
    __shared__ __align__(16) bf16 Vs[64][ATT_LD];
    __shared__ __align__(16) bf16 Qs[64][ATT_LD];
    __shared__ __align__(16) bf16 Gs[64][ATT_LD];
    __shared__ float lse_s[64], del_s[64];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int kb = blockIdx.x, head = blockIdx.y, b = blockIdx.z;
    const size_t rs3 = (size_t)3 * C;
    const bf16* base = qkv + (size_t)b * N * rs3 + head * D;
    const bf16* gbase = dout + (size_t)b * N * C + head * D;
    auto load_tile = [&](bf16(*dst)[ATT_LD], const bf16* src, size_t row_stride, int row0) {
#pragma unroll
        for (int i = 0; i < D / 16; ++i) {
            const int id = tid + i * 128, r = id / (D / 8), c = id % (D / 8);
            cp_async16(smem_u32(&dst[r][c * 8]), src + (size_t)(row0 + r) * row_stride + c * 8, true);
        }
    };
    load_tile(Ks, base + C, rs3, kb * 64);
    load_tile(Vs, base + 2 * C, rs3, kb * 64);
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    uint32_t kf[KS][4], vf[KS][4];
#pragma unroll
    for (int kk = 0; kk < KS; ++kk) {
        ldmatrix_x4(smem_u32(&Ks[warp * 16 + (lane & 15)][kk * 16 + (lane >> 4) * 8]), kf[kk][0], kf[kk][1], kf[kk][2], kf[kk][3]);
        ldmatrix_x4(smem_u32(&Vs[warp * 16 + (lane & 15)][kk * 16 + (lane >> 4) * 8]), vf[kk][0], vf[kk][1], vf[kk][2], vf[kk][3]);
    }
    float dk[D / 8][4], dv[D / 8][4];
#pragma unroll
    for (int i = 0; i < D / 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { dk[i][j] = 0.f; dv[i][j] = 0.f; }
    const int g = lane >> 2, tq = lane & 3;
    const size_t lbase = ((size_t)b * gridDim.y + head) * N;
    for (int qb = 0; qb < N / 64; ++qb) {
        __syncthreads();
        load_tile(Qs, base, rs3, qb * 64);
        load_tile(Gs, gbase, (size_t)C, qb * 64);
        if (tid < 64) { lse_s[tid] = lse[lbase + qb * 64 + tid]; del_s[tid] = delta[lbase + qb * 64 + tid]; }
        cp_async_commit();
        cp_async_wait<0>();
        __syncthreads();
        float s[8][4], dp[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) { s[i][j] = 0.f; dp[i][j] = 0.f; }
#pragma unroll
        for (int kk = 0; kk < KS; ++kk)
#pragma unroll
            for (int np = 0; np < 4; ++np) {
                uint32_t b0, b1, b2, b3;
                ldmatrix_x4(smem_u32(&Qs[np * 16 + (lane & 7) + ((lane >> 4) << 3)][kk * 16 + ((lane >> 3) & 1) * 8]), b0, b1, b2, b3);
                mma_bf16_16816(s[2 * np], kf[kk], b0, b1);       // S^T[key][query]
                mma_bf16_16816(s[2 * np + 1], kf[kk], b2, b3);
                ldmatrix_x4(smem_u32(&Gs[np * 16 + (lane & 7) + ((lane >> 4) << 3)][kk * 16 + ((lane >> 3) & 1) * 8]), b0, b1, b2, b3);
                mma_bf16_16816(dp[2 * np], vf[kk], b0, b1);      // dP^T[key][query]
                mma_bf16_16816(dp[2 * np + 1], vf[kk], b2, b3);
            }
        uint32_t pf[4][4], dsf[4][4];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int c0 = i * 8 + tq * 2;
            const float la = lse_s[c0], lb = lse_s[c0 + 1], da = del_s[c0], db = del_s[c0 + 1];
            const float p0 = exp2f(s[i][0] * scale_log2 - la), p1 = exp2f(s[i][1] * scale_log2 - lb);
            const float p2 = exp2f(s[i][2] * scale_log2 - la), p3 = exp2f(s[i][3] * scale_log2 - lb);
            pf[i >> 1][(i & 1) * 2] = pack_bf16x2(p0, p1);
            pf[i >> 1][(i & 1) * 2 + 1] = pack_bf16x2(p2, p3);
            dsf[i >> 1][(i & 1) * 2] = pack_bf16x2(p0 * (dp[i][0] - da), p1 * (dp[i][1] - db));
            dsf[i >> 1][(i & 1) * 2 + 1] = pack_bf16x2(p2 * (dp[i][2] - da), p3 * (dp[i][3] - db));
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int np = 0; np < D / 16; ++np) {
                uint32_t b0, b1, b2, b3;
                ldmatrix_x4_trans(smem_u32(&Gs[j * 16 + (lane & 7) + ((lane >> 3) & 1) * 8][np * 16 + (lane >> 4) * 8]), b0, b1, b2, b3);
                mma_bf16_16816(dv[2 * np], pf[j], b0, b1);
                mma_bf16_16816(dv[2 * np + 1], pf[j], b2, b3);
                ldmatrix_x4_trans(smem_u32(&Qs[j * 16 + (lane & 7) + ((lane >> 3) & 1) * 8][np * 16 + (lane >> 4) * 8]), b0, b1, b2, b3);
                mma_bf16_16816(dk[2 * np], dsf[j], b0, b1);
                mma_bf16_16816(dk[2 * np + 1], dsf[j], b2, b3);
            }
    }
    const size_t r0 = (size_t)b * N + kb * 64 + warp * 16 + g;
#pragma unroll
    for (int i = 0; i < D / 8; ++i) {
        const int c = head * D + i * 8 + tq * 2;
        *reinterpret_cast<uint32_t*>(dqkv + r0 * rs3 + C + c) = pack_bf16x2(dk[i][0] * scale, dk[i][1] * scale);
        *reinterpret_cast<uint32_t*>(dqkv + (r0 + 8) * rs3 + C + c) = pack_bf16x2(dk[i][2] * scale, dk[i][3] * scale);
        *reinterpret_cast<uint32_t*>(dqkv + r0 * rs3 + 2 * C + c) = pack_bf16x2(dv[i][0], dv[i][1]);
        *reinterpret_cast<uint32_t*>(dqkv + (r0 + 8) * rs3 + 2 * C + c) = pack_bf16x2(dv[i][2], dv[i][3]);
    }
}

}  // namespace rfv
