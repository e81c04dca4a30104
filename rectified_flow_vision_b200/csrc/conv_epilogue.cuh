// Shared epilogue of the tcgen05 convolution kernels: one 128-row x BN-column fp32 accumulator tile in TMEM ->
// + bias | (bias + time projection)  -> + identity residual -> GroupNorm partial sums -> bf16 NHWC store.
// Executed by ONE warp-group (4 warps = the 4 TMEM lane quadrants; thread = one output pixel).  The kernels run two
// such groups, each bound to one of the two accumulator stages, so the epilogues of consecutive tiles overlap.
// Nothing the accumulator does not depend on may sit between the accumulator-full wait and the stores (ncu, round 2: with
// two warps per scheduler the exposed L2 latency of the bias row and of the residual made the epilogue of a 64-channel layer
// with a residual take 4 us per tile against 1.7 us of MMAs): the residual of the first TWO 32-channel chunks is in
// registers and the bias / time-projection row of the tile is in L1 before the wait; later residual chunks are fetched two
// chunks ahead.
// `release` = false: the stage holds a second accumulator that is still to be drained (double tiles, conv_halo.cuh).
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "conv_params.h"

namespace rfv {

__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

template <int BN>
__device__ __forceinline__ void conv_epilogue_tile(const ConvParams& p, uint32_t taddr, int n, bool n_ok, bool valid,
                                                   size_t pix, int nt, int lane, uint64_t* tfull, uint32_t parity,
                                                   uint64_t* tempty, bool release = true) {
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
            mbar_arrive(tempty);
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

// Same epilogue for the halo-reuse kernels, with the output leaving through the TMA unit.  Why (ncu, round 2,
// conv_halo_kernel<64> at 512 images): with thread = pixel every 16-byte STG / LDG of a warp touches 32 different 128-byte
// lines, L1TEX sat at 58-65 % and a layer with a residual was L1-bound (274 us against 204 us without).  Here each warp
// packs its 32 positions x 64 channels into a 4 KB shared-memory box (128B-swizzled, conflict-free 16-byte stores) and one
// lane issues ONE tiled TMA store per box.  A run of 32 flat padded positions (pitch W+1 > 32) holds at most one position
// on the shared zero column; that row is dropped while packing, so the box is 32 or 31 CONSECUTIVE pixels of the per-image
// flat pixel dimension {C, H*W, N} (two tensor maps, box heights 32 and 31); rows behind the image are clipped by the
// tensor's upper bound.  (A clipped per-row store would be simpler but needs negative start coordinates, which TMA stores
// reject -- tools/micro/tma_store_test.cu.)  No global store instruction is left in the kernel.
//   stage: this warp's 4 KB box (1024-byte aligned); p0: flat padded position of the warp's first row.
constexpr int HALO_STAGE_BYTES = 4096;

template <int BN>
__device__ __forceinline__ void conv_epilogue_halo(const ConvParams& p, const CUtensorMap* map32, const CUtensorMap* map31, uint8_t* stage,
                                                   uint32_t taddr, int n, bool valid, size_t pix, int nt, int lane, int p0, int pitch,
                                                   uint64_t* tfull, uint32_t parity, uint64_t* tempty, bool release) {
    constexpr int NP = BN / 64;   // passes of 64 channels = one TMA-store box each
    const bf16* rbase = p.resid ? p.resid + pix * p.Cout + (size_t)nt * BN : nullptr;
    const bool has_res = rbase != nullptr && valid;
    uint4 r0[4], r1[4];   // residual of the pass's first / second 32-channel chunk
    if (has_res) {
#pragma unroll
        for (int i = 0; i < 4; ++i) r0[i] = reinterpret_cast<const uint4*>(rbase)[i];
#pragma unroll
        for (int i = 0; i < 4; ++i) r1[i] = reinterpret_cast<const uint4*>(rbase + 32)[i];
    }
    const float* addbase = p.temb ? p.temb + (size_t)n * p.temb_stride + nt * BN : p.bias + nt * BN;
    if (lane < BN / 32) prefetch_l1(addbase + lane * 32);   // one 128-byte line per 32-channel chunk
    const int rr0 = p0 / pitch, cc0 = p0 - rr0 * pitch;
    const int kz = cc0 == 0 ? 0 : pitch - cc0;                    // index of the zero-column position inside the run (>= 32: none)
    const int pix0 = rr0 * (pitch - 1) + (cc0 == 0 ? 0 : cc0 - 1);   // first pixel of the run inside the image
    const int row = lane - (lane > kz ? 1 : 0);                   // box row of this thread's pixel
    const uint32_t srow = smem_u32(stage) + (uint32_t)row * 128;
    mbar_wait(tfull, parity);
    tc_fence_after();
#pragma unroll 1
    for (int h = 0; h < NP; ++h) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const int ch = 2 * h + c;
            uint32_t acc[32];
            tmem_ld32(taddr + ch * 32, acc);
            float v[32];
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
                const float4 b4 = *reinterpret_cast<const float4*>(addbase + ch * 32 + i);
                v[i] = b4.x; v[i + 1] = b4.y; v[i + 2] = b4.z; v[i + 3] = b4.w;
            }
            if (c == 0) {   // the previous box must have been read by the TMA unit before it is overwritten
                if (lane == 0) bulk_wait_read0();
                __syncwarp();
            }
            tmem_ld_wait();
            if (ch == BN / 32 - 1 && release) {  // accumulator fully in registers: hand the TMEM stage back to the MMA issuer
                tc_fence_before();
                mbar_arrive(tempty);
            }
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] += __uint_as_float(acc[i]);
            if (has_res) {
                uint4* rc = c ? r1 : r0;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float f[8];
                    unpack8(rc[i], f);
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[i * 8 + j] += f[j];
                }
                if (h + 1 < NP) {   // the same chunk of the next pass
#pragma unroll
                    for (int i = 0; i < 4; ++i) rc[i] = reinterpret_cast<const uint4*>(rbase + (ch + 2) * 32)[i];
                }
            }
            // logical 16-byte vector j of box row r sits at (j ^ (r & 7)) * 16 (the box is 1024-byte aligned)
            if (lane != kz) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const uint4 q = pack8(v + i * 8);
                    const uint32_t a = srow + (uint32_t)(((c * 4 + i) ^ (row & 7)) << 4);
                    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(q.x), "r"(q.y), "r"(q.z), "r"(q.w) : "memory");
                }
            }
            if (c == 1) {
                fence_async_smem();
                __syncwarp();
                if (lane == 0) {
                    tma_store_3d(kz < 32 ? map31 : map32, stage, nt * BN + h * 64, pix0, n);
                    bulk_commit();
                }
            }
            if (p.stats) {
                const int c0 = nt * BN + ch * 32;
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
                if ((lane & 3) == 0) {   // tiles never span images: n is warp-uniform
                    const int idx = lane >> 2;
                    float* dst = p.stats + ((size_t)n * (p.Cout >> p.slab_shift) + ((c0 + (idx >> 1) * 8) >> p.slab_shift)) * 2;
                    atomicAdd(dst + (idx & 1), t8[0]);
                }
            }
        }
    }
}

}  // namespace rfv
