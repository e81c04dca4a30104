// Shared epilogue of the tcgen05 convolution kernels: one 128-row x BN-column fp32 accumulator tile in TMEM ->
// + bias | (bias + time projection)  -> + identity residual -> GroupNorm partial sums -> bf16 NHWC store.
// Executed by ONE warp-group (4 warps = the 4 TMEM lane quadrants; thread = one output pixel).  The kernels run two
// such groups, each bound to one of the two accumulator stages, so the epilogues of consecutive tiles overlap; the
// residual (which does not depend on the accumulator) is prefetched before the accumulator-full wait.
#pragma once
#include "common.cuh"
#include "conv_params.h"

namespace rfv {

template <int BN>
__device__ __forceinline__ void conv_epilogue_tile(const ConvParams& p, uint32_t taddr, int n, bool n_ok, bool valid,
                                                   size_t pix, int nt, int lane, uint64_t* tfull, uint32_t parity,
                                                   uint64_t* tempty) {
    const bf16* rbase = p.resid ? p.resid + pix * p.Cout + (size_t)nt * BN : nullptr;
    const bool has_res = rbase != nullptr && valid;
    uint4 rcur[4];
    if (has_res) {
#pragma unroll
        for (int i = 0; i < 4; ++i) rcur[i] = reinterpret_cast<const uint4*>(rbase)[i];
    }
    const float* addbase = p.temb ? p.temb + (size_t)(n_ok ? n : 0) * p.temb_stride + nt * BN : p.bias + nt * BN;
    mbar_wait(tfull, parity);
    tc_fence_after();
#pragma unroll 1
    for (int ch = 0; ch < BN / 32; ++ch) {
        uint32_t acc[32];
        tmem_ld32(taddr + ch * 32, acc);
        uint4 rnext[4];
        if (has_res && ch + 1 < BN / 32) {
#pragma unroll
            for (int i = 0; i < 4; ++i) rnext[i] = reinterpret_cast<const uint4*>(rbase + (ch + 1) * 32)[i];
        }
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
            const float4 b4 = *reinterpret_cast<const float4*>(addbase + ch * 32 + i);
            v[i] = b4.x; v[i + 1] = b4.y; v[i + 2] = b4.z; v[i + 3] = b4.w;
        }
        tmem_ld_wait();
        if (ch == BN / 32 - 1) {  // accumulator fully in registers: hand the TMEM stage back to the MMA issuer
            tc_fence_before();
            mbar_arrive(tempty);
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] += __uint_as_float(acc[i]);
        if (has_res) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float f[8];
                unpack8(rcur[i], f);
#pragma unroll
                for (int j = 0; j < 8; ++j) v[i * 8 + j] += f[j];
            }
        }
        const int c0 = nt * BN + ch * 32;
        if (valid) {
            uint4* op = reinterpret_cast<uint4*>(p.out + pix * p.Cout + c0);
#pragma unroll
            for (int i = 0; i < 4; ++i) op[i] = pack8(v + i * 8);
        }
        if (p.stats) {
            // 8 partial sums per lane (4 slabs x {sum, sum of squares}); after the transposing butterfly lane L
            // (L % 4 == 0) holds the warp total of value L >> 2.  A warp's 32 rows lie inside one image.
            float t8[8];
#pragma unroll
            for (int sl = 0; sl < 4; ++sl) {
                float s = 0.f, ss = 0.f;
#pragma unroll
                for (int j = 0; j < 8; ++j) { const float x = valid ? v[sl * 8 + j] : 0.f; s += x; ss += x * x; }
                t8[sl * 2] = s;
                t8[sl * 2 + 1] = ss;
            }
            warp_reduce8(t8, lane);
            const int n_w = __shfl_sync(0xffffffffu, n, 0);
            const bool ok_w = __shfl_sync(0xffffffffu, (int)n_ok, 0) != 0;
            if ((lane & 3) == 0 && ok_w) {
                const int idx = lane >> 2;
                float* dst = p.stats + ((size_t)n_w * (p.Cout >> p.slab_shift) + ((c0 + (idx >> 1) * 8) >> p.slab_shift)) * 2;
                atomicAdd(dst + (idx & 1), t8[0]);
            }
        }
        if (has_res && ch + 1 < BN / 32) {
#pragma unroll
            for (int i = 0; i < 4; ++i) rcur[i] = rnext[i];
        }
    }
}

}  // namespace rfv
