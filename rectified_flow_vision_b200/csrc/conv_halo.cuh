// 3x3 stride-1 convolution on tcgen05 with HALO REUSE: the input box of a tile is fetched from L2 once per
// 64-channel chunk and all nine taps are issued as shifted views of it.
//
// Why: the plain implicit-GEMM kernel (conv_umma.cuh) re-fetches a 16 KB A box per tap and the weight slice per
// tile; ncu shows it pinned at 8-11 TB/s of L2->SM TMA traffic with the tensor pipe 16-53 % busy
// (profiles/r1_ncu_conv_umma_full.md).  Here the GEMM rows do not enumerate output pixels but 128 consecutive
// positions q of a FLAT PADDED image whose rows have pitch W+1 (one shared zero column between rows):
//       q = r * (W+1) + c + 1      (position c = -1 of every row is the zero column)
// In that space tap (dy,dx) is the constant shift dy*(W+1)+dx, so the A operand of tap t is simply the same shared
// memory buffer addressed `shift(t)` rows (of 128 B) further -- one UMMA descriptor per tap, no data movement.
// The padded layout exists only in shared memory: one 4-D TMA box {64 ch, W+1, ROWS, 1} at (c=-1, r=r_first-1) is
// zero-filled by the TMA unit wherever it leaves the image.  Shifted start addresses are 128-byte but not 1024-byte
// aligned.  Measured on B200: the UMMA unit applies the 128B swizzle XOR on ABSOLUTE shared-memory address bits
// (like the TMA unit that wrote the box), so such descriptors work with base offset 0; setting the base-offset
// field to (addr >> 7) & 7 double-counts the phase and gives wrong results (RFV_FLAG_BASEOFF reproduces that).
// Positions on the zero column (and past the last image row) are computed and discarded (3 % at 64x64, 11 % at 32x32).
// Weights: streamed per (chunk, tap) through their own ring by a second producer thread, or -- when the whole
// [C_out][K] matrix fits (64->64 layers: 72 KB) -- loaded once per CTA and kept resident.
//
//   warp 0  A producer        warp 1  MMA issuer        warp 2  TMEM allocator        warp 3  B producer
//   warps 4-11  epilogue (same as conv_umma.cuh: bias/time projection, residual, GroupNorm partial sums, bf16 store)
// A second K segment (1x1 shortcut over up to two raw tensors) uses the same boxes with the centre shift.
// DOUBLE TILES (g.sub == 2, streamed weights): with 128-position tiles a 128-output-channel layer re-streams its whole
// weight matrix (288 KB for 128->128) per tile and is L2->SM-bound (ncu round 1: 7.6 TB/s of TMA traffic, tensor pipe 43 %).
// A double tile covers 256 consecutive positions: ONE halo box per chunk, every streamed weight block feeds two MMA groups
// (A descriptors 128 rows apart) into two accumulators of the same TMEM stage -- weight traffic per output halves and the
// box redundancy drops from 2.5x / 1.5x to 1.8x / 1.3x (64- / 32-pixel rows).  A trailing sub-tile that lies entirely
// behind the image is skipped.
// Replaces nn.Conv2d call sites models/unet.py:38,41(+51) at the 64x64 / 32x32 (/128x128) levels.
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "conv_params.h"
#include "conv_umma.cuh"

namespace rfv {

struct HaloGeom {
    int W, H, pitch, rows;     // pitch = W + 1; rows = box height (tile rows + 2 halo rows)
    int tiles_per_img, m_tiles, n_tiles;
    int cch0, cch1a, cch1b;    // 64-channel chunks of segment 0 / of the two shortcut sources
    int a_stage_bytes;         // 1024-aligned, >= rows*pitch*128 + 128 (trailing zero row)
    int a_box_bytes;           // rows*pitch*128
    int a_stages, b_stages;
    int resident_b;            // whole weight matrix lives in shared memory
    int base_offset_mode;      // 0 (default, correct on B200): base offset 0; 1: (addr >> 7) & 7 (experiment, wrong)
    int cch0a;                 // fused-GroupNorm kernel: segment-0 chunks [0, cch0a) come from map A0, the rest from A0b
    int sub;                   // M sub-tiles (128 positions each) per tile: 1, or 2 = DOUBLE TILE (see below)
};

__device__ __forceinline__ uint64_t umma_desc_sw128_bo(uint32_t smem_addr, int mode) {
    uint64_t d = umma_desc_sw128(smem_addr);
    if (mode) d |= (uint64_t)((smem_addr >> 7) & 7) << 49;
    return d;
}

template <int BN>
__global__ void __launch_bounds__(UMMA_THREADS, 1)
conv_halo_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
                 const __grid_constant__ CUtensorMap mapA2, const __grid_constant__ CUtensorMap mapW,
                 const __grid_constant__ CUtensorMap mapO32, const __grid_constant__ CUtensorMap mapO31, const ConvParams p, const HaloGeom g) {
    constexpr int B_BYTES = BN * 128;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    // [1 KB guard][A ring][B ring or resident B][epilogue boxes: 8 warps x 4 KB][barriers]
    uint8_t* smem_a = smem + 1024;
    uint8_t* smem_b = smem_a + g.a_stages * g.a_stage_bytes;
    const int nkb0 = 9 * g.cch0, nkb = nkb0 + g.cch1a + g.cch1b;
    const int b_slots = g.resident_b ? nkb : g.b_stages;
    uint8_t* smem_o = smem_b + (size_t)b_slots * B_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_o + 8 * HALO_STAGE_BYTES);
    uint64_t* afull = bars;                      // [a_stages]
    uint64_t* aempty = afull + g.a_stages;       // [a_stages]
    uint64_t* bfull = aempty + g.a_stages;       // [b_stages] (slot 0 doubles as the "resident weights landed" barrier)
    uint64_t* bempty = bfull + g.b_stages;       // [b_stages]
    uint64_t* tfull = bempty + g.b_stages;       // [2]
    uint64_t* tempty = tfull + 2;                // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&mapA0);
        tma_prefetch_desc(&mapW);
        tma_prefetch_desc(&mapO32);
        tma_prefetch_desc(&mapO31);
        for (int s = 0; s < g.a_stages; ++s) { mbar_init(&afull[s], 1); mbar_init(&aempty[s], 1); }
        for (int s = 0; s < g.b_stages; ++s) { mbar_init(&bfull[s], 1); mbar_init(&bempty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&tfull[s], 1); mbar_init(&tempty[s], 128); }
        mbar_fence_init();
    }
    // the row after each box must read as zero (tap (+1,+1) of the last pixel of the last box row lands there)
    if (threadIdx.x < 32)
        for (int s = 0; s < g.a_stages; ++s)
            reinterpret_cast<uint32_t*>(smem_a + (size_t)s * g.a_stage_bytes + g.a_box_bytes)[threadIdx.x] = 0u;
    const int tile_pos = 128 * g.sub;
    if (warp == 2) tmem_alloc(tmem_slot, 2 * g.sub * BN);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy zero fill -> visible to UMMA reads
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int total_tiles = g.m_tiles * g.n_tiles;

    if (warp == 0) {
        {
            // ===================== A producer: one halo box per (tile, 64-channel chunk) =====================
            uint32_t st = 0, ph = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int mt = tile / g.n_tiles;
                const int n = mt / g.tiles_per_img, ti = mt - n * g.tiles_per_img;
                const int rbox = (ti * tile_pos) / g.pitch - 1;
                for (int ch = 0; ch < g.cch0 + g.cch1a + g.cch1b; ++ch) {
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
        }
        __syncwarp();
    } else if (warp == 3) {
        {
            // ===================== B producer =====================
            if (g.resident_b) {
                if ((int)blockIdx.x < total_tiles && elect_one()) {
                    const int nt = blockIdx.x % g.n_tiles;  // resident mode is only used with n_tiles == 1
                    mbar_arrive_expect_tx(&bfull[0], nkb * B_BYTES);
                    for (int kb = 0; kb < nkb; ++kb) tma_load_2d(smem_b + (size_t)kb * B_BYTES, &mapW, &bfull[0], kb * 64, nt * BN);
                }
            } else {
                uint32_t st = 0, ph = 0;
                for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                    const int nt = tile % g.n_tiles;
                    for (int ch = 0; ch < g.cch0 + g.cch1a + g.cch1b; ++ch) {
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
        }
        __syncwarp();
    } else if (warp == 1) {
        {
            // ===================== MMA issuer (whole warp walks the loop, one elected lane issues) =====================
            constexpr uint32_t idesc = umma_idesc_bf16(UMMA_BM, BN);
            uint32_t ast = 0, aph = 0, bst = 0, bph = 0, it = 0;
            if (g.resident_b && (int)blockIdx.x < total_tiles) mbar_wait(&bfull[0], 0);
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
                const int mt = tile / g.n_tiles;
                const int ti = mt % g.tiles_per_img;
                const int q0 = ti * tile_pos;
                const int idx0 = q0 - (q0 / g.pitch - 1) * g.pitch;  // row (128 B) of position q0 inside the box buffer
                const int nsub = (g.sub == 2 && q0 + 128 < g.H * g.pitch) ? 2 : 1;
                const uint32_t as = it & 1, aphase = (it >> 1) & 1;
                mbar_wait(&tempty[as], aphase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * (g.sub * BN);
                const int nch = g.cch0 + g.cch1a + g.cch1b;
                for (int ch = 0; ch < nch; ++ch) {
                    mbar_wait(&afull[ast], aph);
                    tc_fence_after();
                    const uint32_t abase = smem_u32(smem_a + (size_t)ast * g.a_stage_bytes) + (uint32_t)(idx0 * 128);
                    const bool seg0 = ch < g.cch0;
                    if (g.resident_b) {
                        // weights already in shared memory: the whole chunk (9 taps x 4 K-slices) is issued back to back
                        if (elect_one()) {
                            if (seg0) {
#pragma unroll
                                for (int tap = 0; tap < 9; ++tap) {
                                    const int shift = (tap / 3 - 1) * g.pitch + (tap % 3 - 1);
                                    const uint64_t adesc = umma_desc_sw128_bo(abase + (uint32_t)(shift * 128), g.base_offset_mode);
                                    const uint64_t bdesc = umma_desc_sw128(smem_u32(smem_b + (size_t)(tap * g.cch0 + ch) * B_BYTES));
                                    for (int sb = 0; sb < nsub; ++sb) {
#pragma unroll
                                        for (int j = 0; j < 4; ++j)
                                            umma_bf16(d_tmem + sb * BN, adesc + sb * 1024 + 2 * j, bdesc + 2 * j, idesc, (ch | tap | j) != 0);
                                    }
                                }
                            } else {
                                const uint64_t adesc = umma_desc_sw128_bo(abase, g.base_offset_mode);
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
                                const uint64_t adesc = umma_desc_sw128_bo(abase + (uint32_t)(shift * 128), g.base_offset_mode);
                                const uint64_t bdesc = umma_desc_sw128(smem_u32(smem_b + (size_t)bst * B_BYTES));
                                for (int sb = 0; sb < nsub; ++sb) {   // both sub-tiles consume the weight block while it is here
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
        __syncwarp();
    } else if (warp >= 4) {
        // ===================== epilogue: warp-group g drains accumulator stage g (tiles it = g, g+2, ...) =====================
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
