// conv_halo_kernel with GroupNorm(+SiLU) applied to the segment-0 operand INSIDE the kernel (sampling path only).
//
// Why: gn_apply_kernel is 22 % of a velocity evaluation and half of its activation traffic (read 2 B + write 2 B per element
// just to hand a normalised copy to the next conv).  With halo reuse a 64-channel chunk of the input box lands in shared
// memory ONCE per tile, so the normalisation can be applied there, in place, by four extra warps between the TMA write and
// the MMAs: y = silu(x * sc[n,c] + sh[n,c]) with per-(image, channel) coefficients precomputed from the producer's
// GroupNorm statistics (gn_coef_kernel).  Positions outside the image stay exactly zero (the conv pads AFTER the
// normalisation in the reference: models/unet.py:56 -> nn.Conv2d(padding=1)), so the transform is masked by position.
// The decoder's virtual concat [h | skip] is read from its two source tensors directly (segment-0 chunks from two maps).
// The 128B swizzle is an XOR on absolute shared-memory address bits, so logical 16-byte vector j of the row at address a
// sits at a + ((j ^ ((a >> 7) & 7)) << 4).
//
// MEASURED (B200, micro-batch 256): gn_apply drops from 1.22 to 0.34 ms but the fused convs go from 2.15 to 4.63 ms: the box a
// tile needs (5 image rows for 2 rows of outputs) is 2.5x larger than the tile, SiLU costs two MUFU operations per element
// (16 per clock per SM), and four transform warps -- one per scheduler -- cannot hide their own latencies.  The kernel is
// kept behind RFV_FLAG_FUSE_GN as a correct (parity-tested) starting point; making it pay needs >= 8 transform warps, a
// one-MUFU sigmoid and, ideally, row-aligned tiles without halo redundancy.
//
//   warp 0  A producer   warp 1  MMA issuer   warp 2  TMEM alloc   warp 3  B producer   warps 4-11 epilogue   warps 12-15 transform
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "conv_halo.cuh"
#include "conv_params.h"

namespace rfv {

constexpr int HF_THREADS = 768;      // 4 control warps, 8 epilogue warps, 12 transform warps (register budgets re-split below)
constexpr int HF_TWARPS = 12;

// scale / shift per (image, channel) of a GroupNorm(8) over a virtual concat of up to two tensors:
// coef[(n*C + c)*2] = rstd*gamma, coef[..+1] = beta - mean*rstd*gamma
__global__ void __launch_bounds__(256) gn_coef_kernel(const float* __restrict__ stats_a, const float* __restrict__ stats_b,
                                                      const float* __restrict__ gamma, const float* __restrict__ beta,
                                                      float* __restrict__ coef, int Ca, int Cb, int HW, int slab_shift, float eps) {
    __shared__ float gmean[8], grstd[8];
    const int C = Ca + Cb, n = blockIdx.x, cpg = C / 8;
    if (threadIdx.x < 8) {
        const int g = threadIdx.x, slab = 1 << slab_shift;
        float s = 0.f, ss = 0.f;
        for (int c = g * cpg; c < (g + 1) * cpg; c += slab) {
            const float* src = (c < Ca) ? stats_a + ((size_t)n * (Ca >> slab_shift) + (c >> slab_shift)) * 2
                                        : stats_b + ((size_t)n * (Cb >> slab_shift) + ((c - Ca) >> slab_shift)) * 2;
            s += src[0];
            ss += src[1];
        }
        const float cnt = (float)cpg * (float)HW;
        const float mean = s / cnt;
        gmean[g] = mean;
        grstd[g] = rsqrtf(fmaxf(ss / cnt - mean * mean, 0.f) + eps);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const int g = c / cpg;
        const float sc = grstd[g] * gamma[c];
        coef[((size_t)n * C + c) * 2] = sc;
        coef[((size_t)n * C + c) * 2 + 1] = beta[c] - gmean[g] * sc;
    }
}

template <int BN>
__global__ void __launch_bounds__(HF_THREADS, 1)
conv_halo_fused_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA0b,
                       const __grid_constant__ CUtensorMap mapA1, const __grid_constant__ CUtensorMap mapA2,
                       const __grid_constant__ CUtensorMap mapW, const __grid_constant__ CUtensorMap mapO32, const __grid_constant__ CUtensorMap mapO31, const ConvParams p,
                       const HaloGeom g) {
    constexpr int B_BYTES = BN * 128;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* smem_a = smem + 1024;
    uint8_t* smem_b = smem_a + g.a_stages * g.a_stage_bytes;
    const int nkb0 = 9 * g.cch0, nkb = nkb0 + g.cch1a + g.cch1b;
    const int nch = g.cch0 + g.cch1a + g.cch1b;
    const int b_slots = g.resident_b ? nkb : g.b_stages;
    uint8_t* smem_o = smem_b + (size_t)b_slots * B_BYTES;   // epilogue boxes: 8 warps x 4 KB (conv_epilogue_halo)
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_o + 8 * HALO_STAGE_BYTES);
    uint64_t* afull = bars;                      // TMA -> transform
    uint64_t* aready = afull + g.a_stages;       // transform -> MMA
    uint64_t* aempty = aready + g.a_stages;      // MMA -> TMA
    uint64_t* bfull = aempty + g.a_stages;
    uint64_t* bempty = bfull + g.b_stages;
    uint64_t* tfull = bempty + g.b_stages;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&mapA0);
        tma_prefetch_desc(&mapW);
        tma_prefetch_desc(&mapO32);
        tma_prefetch_desc(&mapO31);
        for (int s = 0; s < g.a_stages; ++s) { mbar_init(&afull[s], 1); mbar_init(&aready[s], HF_TWARPS); mbar_init(&aempty[s], 1); }
        for (int s = 0; s < g.b_stages; ++s) { mbar_init(&bfull[s], 1); mbar_init(&bempty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&tfull[s], 1); mbar_init(&tempty[s], 128); }
        mbar_fence_init();
    }
    if (threadIdx.x < 32)
        for (int s = 0; s < g.a_stages; ++s)
            reinterpret_cast<uint32_t*>(smem_a + (size_t)s * g.a_stage_bytes + g.a_box_bytes)[threadIdx.x] = 0u;
    const int tile_pos = 128 * g.sub;   // double tiles: see conv_halo.cuh
    if (warp == 2) tmem_alloc(tmem_slot, 2 * g.sub * BN);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int total_tiles = g.m_tiles * g.n_tiles;
    // 768 threads start with 80 registers each; the epilogue needs ~126.  Warp-group-wide re-split (setmaxnreg):
    // control 32, transform 64, epilogue 128  ->  128*32 + 384*64 + 256*128 = 61,440 = 768*80: the pool a CTA can re-split is
    // what it was launched with, not the SM's register file (an `inc` beyond it waits forever).
    // (each setmaxnreg sits at the top of the branch it governs so that ptxas allocates that branch with its own budget)
    if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
    if (warp == 0) {
        uint32_t st = 0, ph = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const int mt = tile / g.n_tiles;
            const int n = mt / g.tiles_per_img, ti = mt - n * g.tiles_per_img;
            const int rbox = (ti * tile_pos) / g.pitch - 1;
            for (int ch = 0; ch < nch; ++ch) {
                mbar_wait(&aempty[st], ph ^ 1);
                if (elect_one()) {
                    mbar_arrive_expect_tx(&afull[st], g.a_box_bytes);
                    uint8_t* dst = smem_a + (size_t)st * g.a_stage_bytes;
                    if (ch < g.cch0a) tma_load_4d(dst, &mapA0, &afull[st], ch * 64, -1, rbox, n);
                    else if (ch < g.cch0) tma_load_4d(dst, &mapA0b, &afull[st], (ch - g.cch0a) * 64, -1, rbox, n);
                    else if (ch - g.cch0 < g.cch1a) tma_load_4d(dst, &mapA1, &afull[st], (ch - g.cch0) * 64, -1, rbox, n);
                    else tma_load_4d(dst, &mapA2, &afull[st], (ch - g.cch0 - g.cch1a) * 64, -1, rbox, n);
                }
                __syncwarp();
                if (++st == (uint32_t)g.a_stages) { st = 0; ph ^= 1; }
            }
        }
    } else if (warp == 3) {
        if (g.resident_b) {
            if ((int)blockIdx.x < total_tiles && elect_one()) {
                const int nt = blockIdx.x % g.n_tiles;
                mbar_arrive_expect_tx(&bfull[0], nkb * B_BYTES);
                for (int kb = 0; kb < nkb; ++kb) tma_load_2d(smem_b + (size_t)kb * B_BYTES, &mapW, &bfull[0], kb * 64, nt * BN);
            }
        } else {
            uint32_t st = 0, ph = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int nt = tile % g.n_tiles;
                for (int ch = 0; ch < nch; ++ch) {
                    const int ntaps = ch < g.cch0 ? 9 : 1;
                    for (int tap = 0; tap < ntaps; ++tap) {
                        const int kb = ch < g.cch0 ? tap * g.cch0 + ch : nkb0 + (ch - g.cch0);
                        mbar_wait(&bempty[st], ph ^ 1);
                        if (elect_one()) {
                            mbar_arrive_expect_tx(&bfull[st], B_BYTES);
                            tma_load_2d(smem_b + (size_t)st * B_BYTES, &mapW, &bfull[st], kb * 64, nt * BN);
                        }
                        __syncwarp();
                        if (++st == (uint32_t)g.b_stages) { st = 0; ph ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc = umma_idesc_bf16(UMMA_BM, BN);
        uint32_t ast = 0, aph = 0, bst = 0, bph = 0, it = 0;
        if (g.resident_b && (int)blockIdx.x < total_tiles) mbar_wait(&bfull[0], 0);
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
            const int mt = tile / g.n_tiles;
            const int ti = mt % g.tiles_per_img;
            const int q0 = ti * tile_pos;
            const int idx0 = q0 - (q0 / g.pitch - 1) * g.pitch;
            const int nsub = (g.sub == 2 && q0 + 128 < g.H * g.pitch) ? 2 : 1;
            const uint32_t as = it & 1, aphase = (it >> 1) & 1;
            mbar_wait(&tempty[as], aphase ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + as * (g.sub * BN);
            for (int ch = 0; ch < nch; ++ch) {
                mbar_wait(&aready[ast], aph);       // the chunk has been normalised in place
                tc_fence_after();
                const uint32_t abase = smem_u32(smem_a + (size_t)ast * g.a_stage_bytes) + (uint32_t)(idx0 * 128);
                const bool seg0 = ch < g.cch0;
                if (g.resident_b) {
                    if (elect_one()) {
                        if (seg0) {
#pragma unroll
                            for (int tap = 0; tap < 9; ++tap) {
                                const int shift = (tap / 3 - 1) * g.pitch + (tap % 3 - 1);
                                const uint64_t adesc = umma_desc_sw128(abase + (uint32_t)(shift * 128));
                                const uint64_t bdesc = umma_desc_sw128(smem_u32(smem_b + (size_t)(tap * g.cch0 + ch) * B_BYTES));
                                for (int sb = 0; sb < nsub; ++sb) {
#pragma unroll
                                    for (int j = 0; j < 4; ++j)
                                        umma_bf16(d_tmem + sb * BN, adesc + sb * 1024 + 2 * j, bdesc + 2 * j, idesc, (ch | tap | j) != 0);
                                }
                            }
                        } else {
                            const uint64_t adesc = umma_desc_sw128(abase);
                            const uint64_t bdesc = umma_desc_sw128(smem_u32(smem_b + (size_t)(nkb0 + ch - g.cch0) * B_BYTES));
                            for (int sb = 0; sb < nsub; ++sb) {
#pragma unroll
                                for (int j = 0; j < 4; ++j) umma_bf16(d_tmem + sb * BN, adesc + sb * 1024 + 2 * j, bdesc + 2 * j, idesc, 1u);
                            }
                        }
                        umma_commit(&aempty[ast]);
                        if (ch == nch - 1) umma_commit(&tfull[as]);
                    }
                    __syncwarp();
                } else {
                    const int ntaps = seg0 ? 9 : 1;
                    for (int tap = 0; tap < ntaps; ++tap) {
                        const int shift = seg0 ? (tap / 3 - 1) * g.pitch + (tap % 3 - 1) : 0;
                        mbar_wait(&bfull[bst], bph);
                        tc_fence_after();
                        if (elect_one()) {
                            const uint64_t adesc = umma_desc_sw128(abase + (uint32_t)(shift * 128));
                            const uint64_t bdesc = umma_desc_sw128(smem_u32(smem_b + (size_t)bst * B_BYTES));
                            for (int sb = 0; sb < nsub; ++sb) {
#pragma unroll
                                for (int j = 0; j < 4; ++j)
                                    umma_bf16(d_tmem + sb * BN, adesc + sb * 1024 + 2 * j, bdesc + 2 * j, idesc, (ch | tap | j) != 0);
                            }
                            umma_commit(&bempty[bst]);
                            if (tap == ntaps - 1) {
                                umma_commit(&aempty[ast]);
                                if (ch == nch - 1) umma_commit(&tfull[as]);
                            }
                        }
                        __syncwarp();
                        if (++bst == (uint32_t)g.b_stages) { bst = 0; bph ^= 1; }
                    }
                }
                if (++ast == (uint32_t)g.a_stages) { ast = 0; aph ^= 1; }
            }
        }
    }
    } else if (warp >= 12) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
        // ===================== transform: GroupNorm(+SiLU) in place on every segment-0 chunk =====================
        const int tt = threadIdx.x - 12 * 32;          // 0 .. 32*HF_TWARPS-1
        constexpr int PL = HF_TWARPS * 4;              // position lanes
        const int j = tt & 7, pl = tt >> 3;            // logical 16-byte vector (8 channels) / position lane
        const int npos = g.rows * g.pitch;
        uint32_t st = 0, ph = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const int mt = tile / g.n_tiles;
            const int n = mt / g.tiles_per_img, ti = mt - n * g.tiles_per_img;
            const int rbox = (ti * tile_pos) / g.pitch - 1;
            for (int ch = 0; ch < nch; ++ch) {
                float sc[8], sh[8];
                const bool seg0 = ch < g.cch0;
                if (seg0) {   // coefficients of this (image, chunk): issued before the wait so the latency overlaps it
                    const float4* cp = reinterpret_cast<const float4*>(p.gn_coef + ((size_t)n * p.gn_C + ch * 64 + j * 8) * 2);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float4 v = cp[i];
                        sc[2 * i] = v.x; sh[2 * i] = v.y; sc[2 * i + 1] = v.z; sh[2 * i + 1] = v.w;
                    }
                    if (p.gn_silu) {   // silu(y) = h (1 + tanh h), h = y / 2: one MUFU op per element (see gn_apply_kernel)
#pragma unroll
                        for (int i = 0; i < 8; ++i) { sc[i] *= 0.5f; sh[i] *= 0.5f; }
                    }
                }
                mbar_wait(&afull[st], ph);
                if (seg0) {
                    // Stage buffers are 1024-byte aligned and the position stride PL is a multiple of 8, so the 128B-swizzle term
                    // of this thread's 16-byte vector, j ^ (position & 7), is the same for every position it visits: the address
                    // just advances by PL rows.  Two positions per iteration give the scheduler two independent dependency
                    // chains (load -> unpack -> FMA -> MUFU -> FMA -> pack -> store); the loop is issue-bound.
                    const uint32_t base = smem_u32(smem_a + (size_t)st * g.a_stage_bytes);
                    const int row_lo = max(0, -rbox), row_hi = min(g.rows, g.H - rbox);   // box rows inside the image
                    uint32_t addr = base + (uint32_t)pl * 128 + (uint32_t)((j ^ (pl & 7)) << 4);
                    int rb = pl / g.pitch, cb = pl - rb * g.pitch;
                    auto advance = [&](int& r_, int& c_) {
                        c_ += PL;
                        while (c_ >= g.pitch) { c_ -= g.pitch; ++r_; }
                    };
                    auto xform = [&](uint4& q) {
                        float f[8];
                        unpack8(q, f);
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const float y = fmaf(f[i], sc[i], sh[i]);
                            if (p.gn_silu) {
                                float t;
                                asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(y));
                                f[i] = fmaf(y, t, y);
                            } else {
                                f[i] = y;
                            }
                        }
                        q = pack8(f);
                    };
                    for (int pos = pl; pos < npos; pos += 2 * PL, addr += 2 * PL * 128) {
                        int rb1 = rb, cb1 = cb;
                        advance(rb1, cb1);
                        // padding positions (column 0, rows outside the image) stay zero
                        const bool v0 = cb >= 1 && rb >= row_lo && rb < row_hi;
                        const bool v1 = pos + PL < npos && cb1 >= 1 && rb1 >= row_lo && rb1 < row_hi;
                        uint4 q0, q1;
                        if (v0) asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(q0.x), "=r"(q0.y), "=r"(q0.z), "=r"(q0.w) : "r"(addr));
                        if (v1) asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(q1.x), "=r"(q1.y), "=r"(q1.z), "=r"(q1.w) : "r"(addr + PL * 128));
                        if (v0) xform(q0);
                        if (v1) xform(q1);
                        if (v0) asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(q0.x), "r"(q0.y), "r"(q0.z), "r"(q0.w) : "memory");
                        if (v1) asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr + PL * 128), "r"(q1.x), "r"(q1.y), "r"(q1.z), "r"(q1.w) : "memory");
                        rb = rb1; cb = cb1;
                        advance(rb, cb);
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes -> visible to the UMMA reads
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&aready[st]);
                if (++st == (uint32_t)g.a_stages) { st = 0; ph ^= 1; }
            }
        }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 128;");
        const int q = warp & 3, grp = (warp - 4) >> 2;
        const int r = q * 32 + lane;
        uint32_t it = grp;
        for (int tile = blockIdx.x + grp * gridDim.x; tile < total_tiles; tile += 2 * gridDim.x, it += 2) {
            const int mt = tile / g.n_tiles, nt = tile - mt * g.n_tiles;
            const int n = mt / g.tiles_per_img, ti = mt - n * g.tiles_per_img;
            const int nsub = (g.sub == 2 && ti * tile_pos + 128 < g.H * g.pitch) ? 2 : 1;
            for (int sb = 0; sb < nsub; ++sb) {
                const int pos = ti * tile_pos + sb * 128 + r;
                const int rr = pos / g.pitch, cc = pos - rr * g.pitch;
                const bool valid = cc >= 1 && rr < g.H;     // (m_tiles = B * tiles_per_img: every tile lies inside the batch)
                const size_t pix = ((size_t)n * g.H + rr) * g.W + (cc - 1);
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + grp * (g.sub * BN) + sb * BN;
                conv_epilogue_halo<BN>(p, &mapO32, &mapO31, smem_o + (warp - 4) * HALO_STAGE_BYTES, taddr, n, valid, pix, nt, lane, pos - lane,
                                       g.pitch, &tfull[grp], (it >> 1) & 1, &tempty[grp], sb == nsub - 1);
            }
        }
        if (lane == 0) bulk_wait_all0();   // this lane's TMA stores must have left shared memory (and landed) before the CTA exits
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 2 * g.sub * BN);
}

}  // namespace rfv
