// Shared epilogue of the tcgen05 convolution kernels: one 128-row x BN-column fp32 accumulator tile in TMEM ->
// + bias | (bias + time projection)  -> + identity residual -> GroupNorm partial sums -> bf16 NHWC store.
// Executed by ONE warp-group (4 warps = the 4 TMEM lane quadrants; thread = one output pixel).  The kernels run two
// such groups, each bound to one of the two accumulator stages, so the epilogues of consecutive tiles overlap.
// Nothing the accumulator does not depend on may sit between the accumulator-full wait and the stores (ncu, round 2: with
// two warps per scheduler the exposed L2 latency of the bias row and of the residual made the epilogue of a 64-channel layer
// with a residual take 4 us per tile against 1.7 us of MMAs): the residual of the first TWO 32-channel chunks is in
// registers and the bias / time-projection row of the tile is in L1 before the wait; later residual chunks are fetched two
// chunks ahead.
// `release` = false: the stage holds a second accumulator that is still to be drained.
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "conv_params.h"

namespace rfv {

__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

template <int BN>
__device__ __forceinline__ void conv_epilogue_tile(const ConvParams& p, uint32_t taddr, int n, bool n_ok, bool valid,
                                                   size_t pix, int nt, int lane, uint64_t* tfull, uint32_t parity,
                                                   uint64_t* tempty, bool release = true, bool pair = false) {
    constexpr int NCH = BN / 32;
    const bf16* rbase = p.resid ? p.resid + pix * p.Cout + (size_t)nt * BN : nullptr;
    const bool has_res = rbase != nullptr && valid;
    uint4 r0[4], r1[4];   // residual of chunk c (even c: r0, odd c: r1)
    if (has_res) {
#pragma unroll
        for (int i = 0; i < 4; ++i) r0[i] = reinterpret_cast<const uint4*>(rbase)[i];
        if (NCH > 1) {
#pragma unroll
            for (int i = 0; i < 4; ++i) r1[i] = reinterpret_cast<const uint4*>(rbase + 32)[i];
        }
    }
    const float* addbase = p.temb ? p.temb + (size_t)(n_ok ? n : 0) * p.temb_stride + nt * BN : p.bias + nt * BN;
    if (lane < NCH) prefetch_l1(addbase + lane * 32);   // one 128-byte line per 32-channel chunk
    mbar_wait(tfull, parity);
    tc_fence_after();
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
        uint32_t acc[32];
        tmem_ld32(taddr + ch * 32, acc);
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
            const float4 b4 = *reinterpret_cast<const float4*>(addbase + ch * 32 + i);
            v[i] = b4.x; v[i + 1] = b4.y; v[i + 2] = b4.z; v[i + 3] = b4.w;
        }
        tmem_ld_wait();
        if (ch == NCH - 1 && release) {  // accumulator fully in registers: hand the TMEM stage back to the MMA issuer
            tc_fence_before();
            if (pair) mbar_arrive_remote(tempty, 0);   // CTA pair: the leader's MMA issuer waits for both CTAs' epilogues
            else mbar_arrive(tempty);
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] += __uint_as_float(acc[i]);
        if (has_res) {
            uint4* rc = (ch & 1) ? r1 : r0;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float f[8];
                unpack8(rc[i], f);
#pragma unroll
                for (int j = 0; j < 8; ++j) v[i * 8 + j] += f[j];
            }
            if (ch + 2 < NCH) {
#pragma unroll
                for (int i = 0; i < 4; ++i) rc[i] = reinterpret_cast<const uint4*>(rbase + (ch + 2) * 32)[i];
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
    }
}

}  // namespace rfv
