// 3x3 stride-1 convolution with 64 output channels on tcgen05: halo reuse (conv_halo.cuh) + TWO TAPS PER MMA.
//
// Why: in SS mode every tcgen05.mma with M = 128 re-reads its 4 KB A operand from shared memory, which takes ~81 cycles
// whatever N is (tools/micro/umma_rate.cu: N=64 81.8, N=128 81.2, N=256 128 clk per instruction).  A 64-output-channel
// layer issued as N = 64 MMAs therefore cannot exceed 39 % of the tensor pipe.  Here the B operand of one MMA is the
// weight block of tap (dy,-1) stacked on the block of tap (dy,0): N = 128 for the price of N = 64.  Both halves see the
// SAME A rows (the shift of tap (dy,-1)), so the second half computes the (dy,0) contribution one position early:
//      D1[q] += A[q + s(dy,-1)] . W(dy,-1)          (exact contribution to out[q])
//      D2[q] += A[q + s(dy,-1)] . W(dy, 0)  = A[(q-1) + s(dy,0)] . W(dy,0)      (the contribution to out[q-1])
// and the epilogue forms out[p] = D1[p] + D2[p+1] (one shuffle per value; the lane-31 neighbour comes through shared
// memory).  Tap (dy,+1) stays an N = 64 MMA into D1.  Nine taps = 3 x (N=128) + 3 x (N=64) instructions per 64-channel
// chunk instead of 9.  Tiles advance by 127 positions (position 127 of a tile has no right neighbour in its own tile).
//
//   warp 0  A producer    warp 1  MMA issuer    warp 2  TMEM allocator    warp 3  B producer    warps 4-11  epilogue
// Weights live in shared memory as [chunk][tap] 8 KB blocks so that a (dy,-1),(dy,0) pair is one contiguous 128-row
// K-major operand; resident for the whole kernel when they fit (64->64: 72 KB), else streamed per (chunk, dy) triple.
// Measured on B200 (micro-batch 256): 192->64 @64x64 0.403 -> 0.283 ms.  For single-chunk 64->64 layers with resident
// weights the MMA stream drops from 0.107 to 0.091 ms but the heavier epilogue (two accumulator halves, neighbour exchange)
// shares the LSU / shared-memory path with the tensor core's operand fetch and the layer ends up at 0.124 ms, so the engine
// uses this kernel only where the K loop is long enough to hide the epilogue (C_in >= 128 or several N tiles).
// Replaces nn.Conv2d call sites models/unet.py:38,41(+51) with 64 output channels (and their data-gradient twins).
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "conv_halo.cuh"
#include "conv_params.h"

namespace rfv {

constexpr int HP_BLK = 64 * 128;   // one (tap, chunk) weight block: 64 rows x 128 B
constexpr int HP_TRI = 3 * HP_BLK;
constexpr int HP_XCH_BYTES = 4 * 2 * 4 * 16 * 4;   // [epilogue warp-group][16-column sub-chunk][warp][16 floats]
constexpr int HP_THREADS = 128 + 16 * 32;           // 4 control warps + 16 epilogue warps

__global__ void __launch_bounds__(HP_THREADS, 1)
conv_halo_pair_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
                      const __grid_constant__ CUtensorMap mapA2, const __grid_constant__ CUtensorMap mapW, const ConvParams p,
                      const HaloGeom g) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* smem_a = smem + 1024;   // [1 KB guard][A ring][B region][barriers][exchange]
    uint8_t* smem_b = smem_a + g.a_stages * g.a_stage_bytes;
    const int nch = g.cch0 + g.cch1a + g.cch1b;
    const int b_region = g.resident_b ? (9 * g.cch0 + g.cch1a + g.cch1b) * HP_BLK : g.b_stages * HP_TRI;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_b + b_region);
    uint64_t* afull = bars;
    uint64_t* aempty = afull + g.a_stages;
    uint64_t* bfull = aempty + g.a_stages;
    uint64_t* bempty = bfull + g.b_stages;
    uint64_t* tfull = bempty + g.b_stages;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
    float* xch = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&mapA0);
        tma_prefetch_desc(&mapW);
        for (int s = 0; s < g.a_stages; ++s) { mbar_init(&afull[s], 1); mbar_init(&aempty[s], 1); }
        for (int s = 0; s < g.b_stages; ++s) { mbar_init(&bfull[s], 1); mbar_init(&bempty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&tfull[s], 1); mbar_init(&tempty[s], 256); }
        mbar_fence_init();
    }
    if (threadIdx.x < 32)   // the row after each box must read as zero
        for (int s = 0; s < g.a_stages; ++s)
            reinterpret_cast<uint32_t*>(smem_a + (size_t)s * g.a_stage_bytes + g.a_box_bytes)[threadIdx.x] = 0u;
    if (warp == 2) tmem_alloc(tmem_slot, 256);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int total_tiles = g.m_tiles * g.n_tiles;

    if (warp == 0) {
        // ===================== A producer: one halo box per (tile, 64-channel chunk) =====================
        uint32_t st = 0, ph = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const int mt = tile / g.n_tiles;
            const int n = mt / g.tiles_per_img, ti = mt - n * g.tiles_per_img;
            const int rbox = (ti * 127) / g.pitch - 1;
            for (int ch = 0; ch < nch; ++ch) {
                mbar_wait(&aempty[st], ph ^ 1);
                if (elect_one()) {
                    mbar_arrive_expect_tx(&afull[st], g.a_box_bytes);
                    uint8_t* dst = smem_a + (size_t)st * g.a_stage_bytes;
                    if (ch < g.cch0) tma_load_4d(dst, &mapA0, &afull[st], ch * 64, -1, rbox, n);
                    else if (ch - g.cch0 < g.cch1a) tma_load_4d(dst, &mapA1, &afull[st], (ch - g.cch0) * 64, -1, rbox, n);
                    else tma_load_4d(dst, &mapA2, &afull[st], (ch - g.cch0 - g.cch1a) * 64, -1, rbox, n);
                }
                __syncwarp();
                if (++st == (uint32_t)g.a_stages) { st = 0; ph ^= 1; }
            }
        }
    } else if (warp == 3) {
        // ===================== B producer =====================
        const int nkb0 = 9 * g.cch0;
        if (g.resident_b) {
            if ((int)blockIdx.x < total_tiles && elect_one()) {
                mbar_arrive_expect_tx(&bfull[0], (nkb0 + g.cch1a + g.cch1b) * HP_BLK);
                for (int ch = 0; ch < g.cch0; ++ch)
                    for (int tap = 0; tap < 9; ++tap)
                        tma_load_2d(smem_b + (size_t)(ch * 9 + tap) * HP_BLK, &mapW, &bfull[0], (tap * g.cch0 + ch) * 64, 0);
                for (int k = 0; k < g.cch1a + g.cch1b; ++k)
                    tma_load_2d(smem_b + (size_t)(nkb0 + k) * HP_BLK, &mapW, &bfull[0], (nkb0 + k) * 64, 0);
            }
        } else {
            uint32_t st = 0, ph = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int nt = tile % g.n_tiles;
                for (int ch = 0; ch < nch; ++ch) {
                    const int nr = ch < g.cch0 ? 3 : 1;
                    for (int r = 0; r < nr; ++r) {
                        mbar_wait(&bempty[st], ph ^ 1);
                        if (elect_one()) {
                            uint8_t* dst = smem_b + (size_t)st * HP_TRI;
                            if (ch < g.cch0) {
                                mbar_arrive_expect_tx(&bfull[st], HP_TRI);
                                for (int j = 0; j < 3; ++j)
                                    tma_load_2d(dst + j * HP_BLK, &mapW, &bfull[st], ((3 * r + j) * g.cch0 + ch) * 64, nt * 64);
                            } else {
                                mbar_arrive_expect_tx(&bfull[st], HP_BLK);
                                tma_load_2d(dst, &mapW, &bfull[st], (nkb0 + ch - g.cch0) * 64, nt * 64);
                            }
                        }
                        __syncwarp();
                        if (++st == (uint32_t)g.b_stages) { st = 0; ph ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        constexpr uint32_t idesc128 = umma_idesc_bf16(UMMA_BM, 128), idesc64 = umma_idesc_bf16(UMMA_BM, 64);
        const int nkb0 = 9 * g.cch0;
        uint32_t ast = 0, aph = 0, bst = 0, bph = 0, it = 0;
        if (g.resident_b && (int)blockIdx.x < total_tiles) mbar_wait(&bfull[0], 0);
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
            const int mt = tile / g.n_tiles;
            const int ti = mt % g.tiles_per_img;
            const int q0 = ti * 127;
            const int idx0 = q0 - (q0 / g.pitch - 1) * g.pitch;
            const uint32_t as = it & 1, aphase = (it >> 1) & 1;
            mbar_wait(&tempty[as], aphase ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + as * 128;
            for (int ch = 0; ch < nch; ++ch) {
                mbar_wait(&afull[ast], aph);
                tc_fence_after();
                const uint32_t abase = smem_u32(smem_a + (size_t)ast * g.a_stage_bytes) + (uint32_t)(idx0 * 128);
                const bool seg0 = ch < g.cch0;
                const int nr = seg0 ? 3 : 1;
                if (g.resident_b && seg0) {
                    // all weights are in shared memory: issue the three N=128 groups back to back, then the three N=64
                    // groups (switching the instruction shape between consecutive MMAs drains the tensor pipe)
                    if (elect_one()) {
                        const uint32_t bb = smem_u32(smem_b + (size_t)(ch * 9) * HP_BLK);
#pragma unroll
                        for (int r = 0; r < 3; ++r) {
                            const uint64_t a1 = umma_desc_sw128(abase + (uint32_t)(((r - 1) * g.pitch - 1) * 128));
                            const uint64_t b1 = umma_desc_sw128(bb + 3 * r * HP_BLK);
#pragma unroll
                            for (int j = 0; j < 4; ++j) umma_bf16(d_tmem, a1 + 2 * j, b1 + 2 * j, idesc128, (ch | r | j) != 0);
                        }
#pragma unroll
                        for (int r = 0; r < 3; ++r) {
                            const uint64_t a2 = umma_desc_sw128(abase + (uint32_t)(((r - 1) * g.pitch + 1) * 128));
                            const uint64_t b2 = umma_desc_sw128(bb + (3 * r + 2) * HP_BLK);
#pragma unroll
                            for (int j = 0; j < 4; ++j) umma_bf16(d_tmem, a2 + 2 * j, b2 + 2 * j, idesc64, 1u);
                        }
                        umma_commit(&aempty[ast]);
                        if (ch == nch - 1) umma_commit(&tfull[as]);
                    }
                    __syncwarp();
                } else
                for (int r = 0; r < nr; ++r) {
                    uint32_t bbase;
                    if (g.resident_b) bbase = smem_u32(smem_b + (size_t)(nkb0 + ch - g.cch0) * HP_BLK);
                    else {
                        mbar_wait(&bfull[bst], bph);
                        tc_fence_after();
                        bbase = smem_u32(smem_b + (size_t)bst * HP_TRI);
                    }
                    if (elect_one()) {
                        if (seg0) {
                            const int s0 = (r - 1) * g.pitch;
                            const uint64_t a1 = umma_desc_sw128(abase + (uint32_t)((s0 - 1) * 128)), b1 = umma_desc_sw128(bbase);
#pragma unroll
                            for (int j = 0; j < 4; ++j) umma_bf16(d_tmem, a1 + 2 * j, b1 + 2 * j, idesc128, (ch | r | j) != 0);
                            const uint64_t a2 = umma_desc_sw128(abase + (uint32_t)((s0 + 1) * 128)), b2 = umma_desc_sw128(bbase + 2 * HP_BLK);
#pragma unroll
                            for (int j = 0; j < 4; ++j) umma_bf16(d_tmem, a2 + 2 * j, b2 + 2 * j, idesc64, 1u);
                        } else {
                            const uint64_t a1 = umma_desc_sw128(abase), b1 = umma_desc_sw128(bbase);
#pragma unroll
                            for (int j = 0; j < 4; ++j) umma_bf16(d_tmem, a1 + 2 * j, b1 + 2 * j, idesc64, 1u);
                        }
                        if (!g.resident_b) umma_commit(&bempty[bst]);
                        if (r == nr - 1) {
                            umma_commit(&aempty[ast]);
                            if (ch == nch - 1) umma_commit(&tfull[as]);
                        }
                    }
                    __syncwarp();
                    if (!g.resident_b && ++bst == (uint32_t)g.b_stages) { bst = 0; bph ^= 1; }
                }
                if (++ast == (uint32_t)g.a_stages) { ast = 0; aph ^= 1; }
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue: out[p] = D1[p] + D2[p+1] (+ bias/time, residual, GroupNorm sums) =====================
        // 16 warps = {accumulator stage (tile parity)} x {32-channel half} x {TMEM lane quadrant}; a thread owns one output
        // position and 32 channels, processed as two 16-column sub-chunks (keeps the register arrays small at 640 threads)
        const int e = warp - 4, q = e & 3, grp = (e >> 2) & 1, half = e >> 3;
        const int r = q * 32 + lane;
        const int wg = grp * 2 + half;                       // warp-group id: named barrier 1 + wg, 128 threads
        float* xg = xch + wg * (2 * 4 * 16);
        uint32_t it = grp;
        for (int tile = blockIdx.x + grp * gridDim.x; tile < total_tiles; tile += 2 * gridDim.x, it += 2) {
            const int mt = tile / g.n_tiles, nt = tile - mt * g.n_tiles;
            const int n = mt / g.tiles_per_img, ti = mt - n * g.tiles_per_img;
            const int pos = ti * 127 + r;
            const int rr = pos / g.pitch, cc = pos - rr * g.pitch;
            const bool n_ok = n < p.B;
            const bool valid = n_ok && r < 127 && cc >= 1 && rr < g.H;
            const size_t pix = ((size_t)n * g.H + rr) * g.W + (cc - 1);
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + grp * 128 + half * 32;
            const int cbase = nt * 64 + half * 32;           // first output channel of this thread
            const bf16* rbase = p.resid ? p.resid + pix * p.Cout + cbase : nullptr;
            const bool has_res = rbase != nullptr && valid;
            uint4 rv[4];
            if (has_res) {
#pragma unroll
                for (int i = 0; i < 4; ++i) rv[i] = reinterpret_cast<const uint4*>(rbase)[i];
            }
            const float* addbase = (p.temb ? p.temb + (size_t)(n_ok ? n : 0) * p.temb_stride : p.bias) + cbase;
            mbar_wait(&tfull[grp], (it >> 1) & 1);
            tc_fence_after();
#pragma unroll
            for (int sub = 0; sub < 2; ++sub) {
                uint32_t a1[16], a2[16];
                tmem_ld16(taddr + sub * 16, a1);
                tmem_ld16(taddr + 64 + sub * 16, a2);
                float v[16];
#pragma unroll
                for (int i = 0; i < 16; i += 4) {
                    const float4 b4 = *reinterpret_cast<const float4*>(addbase + sub * 16 + i);
                    v[i] = b4.x; v[i + 1] = b4.y; v[i + 2] = b4.z; v[i + 3] = b4.w;
                }
                tmem_ld_wait();
                if (sub == 1) {   // accumulator fully in registers: hand the TMEM stage back to the MMA issuer
                    tc_fence_before();
                    mbar_arrive(&tempty[grp]);
                }
                float* xs = xg + sub * (4 * 16);
                if (lane == 0) {
#pragma unroll
                    for (int i = 0; i < 16; i += 4)
                        *reinterpret_cast<float4*>(xs + q * 16 + i) =
                            make_float4(__uint_as_float(a2[i]), __uint_as_float(a2[i + 1]), __uint_as_float(a2[i + 2]), __uint_as_float(a2[i + 3]));
                }
                asm volatile("bar.sync %0, 128;" ::"r"(1 + wg) : "memory");
                float nb31[16];
                if (lane == 31) {
#pragma unroll
                    for (int i = 0; i < 16; i += 4) {
                        const float4 t = q < 3 ? *reinterpret_cast<const float4*>(xs + (q + 1) * 16 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
                        nb31[i] = t.x; nb31[i + 1] = t.y; nb31[i + 2] = t.z; nb31[i + 3] = t.w;
                    }
                }
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    float nb = __shfl_down_sync(0xffffffffu, __uint_as_float(a2[i]), 1);
                    if (lane == 31) nb = nb31[i];
                    v[i] += __uint_as_float(a1[i]) + nb;
                }
                if (has_res) {
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        float f[8];
                        unpack8(rv[sub * 2 + i], f);
#pragma unroll
                        for (int j = 0; j < 8; ++j) v[i * 8 + j] += f[j];
                    }
                }
                const int c0 = cbase + sub * 16;
                if (valid) {
                    uint4* op = reinterpret_cast<uint4*>(p.out + pix * p.Cout + c0);
                    op[0] = pack8(v);
                    op[1] = pack8(v + 8);
                }
                if (p.stats) {
                    // 4 partial sums per lane (2 slabs x {sum, sum of squares}); after the transposing butterfly every lane
                    // holds the warp total of value ((lane >> 4) & 1) * 2 + ((lane >> 3) & 1)
                    float t4[4];
#pragma unroll
                    for (int sl = 0; sl < 2; ++sl) {
                        float s_ = 0.f, ss = 0.f;
#pragma unroll
                        for (int j = 0; j < 8; ++j) { const float x = valid ? v[sl * 8 + j] : 0.f; s_ += x; ss += x * x; }
                        t4[sl * 2] = s_;
                        t4[sl * 2 + 1] = ss;
                    }
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const float send = (lane & 16) ? t4[i] : t4[i + 2], keep = (lane & 16) ? t4[i + 2] : t4[i];
                        t4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
                    }
                    {
                        const float send = (lane & 8) ? t4[0] : t4[1], keep = (lane & 8) ? t4[1] : t4[0];
                        t4[0] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
                    }
                    t4[0] += __shfl_xor_sync(0xffffffffu, t4[0], 4);
                    t4[0] += __shfl_xor_sync(0xffffffffu, t4[0], 2);
                    t4[0] += __shfl_xor_sync(0xffffffffu, t4[0], 1);
                    if ((lane & 7) == 0 && n_ok) {   // a warp's 32 positions lie inside one image (tiles never span images)
                        const int idx = ((lane >> 4) & 1) * 2 + ((lane >> 3) & 1);
                        float* dst = p.stats + ((size_t)n * (p.Cout >> p.slab_shift) + ((c0 + (idx >> 1) * 8) >> p.slab_shift)) * 2;
                        atomicAdd(dst + (idx & 1), t4[0]);
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 256);
}

}  // namespace rfv
