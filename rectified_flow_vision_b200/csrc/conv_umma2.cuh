// conv_umma_kernel<256 / 128> on a CTA PAIR (tcgen05 cta_group::2): the 256-output-channel convolutions of the 16x16 level
// and the 128-channel sub-pixel upsample conv
// (3x3 stride 1, also as the four sub-pixel phases of an upsample conv; models/unet.py:38,41,51,217 at the lowest resolution).
//
// Why: with one CTA per tile every (tap, 64-channel chunk) stage moves a 16 KB pixel box AND a 32 KB weight slice into
// shared memory for four M128 x N256 x K16 MMAs (512 cycles) that read 4 KB of A and 8 KB of B each: 94 B/clk of TMA writes
// plus 96 B/clk of operand reads against the 128 B/clk a shared memory delivers -- those launches sat at 71 % of the MMA
// rate (and 12.4 TB/s of L2->SM traffic chip-wide).  In pair mode the two CTAs of a cluster (the two SMs of a TPC) execute
// ONE M = 256 instruction: each CTA holds its own 128-pixel A box and HALF of the weight slice (128 of the 256 output
// channels, 16 KB), the tensor core reads the other half from the peer, and each CTA keeps its own 128 accumulator rows x 256
// columns in its TMEM.  Per SM: 64 B/clk of TMA writes + 64 B/clk of operand reads, and the weight slice crosses L2->SM once
// per pair instead of once per CTA.
//
// Protocol (everything else -- tile enumeration, im2col by TMA, epilogue -- is conv_umma_kernel's):
//   * both CTAs run a producer warp; their TMA loads (cta_group::2) complete on the LEADER's full barrier, which the leader
//     arms with the bytes of both CTAs; stage-free and accumulator-full signals are multicast commits to both CTAs;
//   * only the leader's MMA warp issues; it waits on its accumulator-empty barrier for the epilogue threads of BOTH CTAs
//     (the peer's arrive remotely);
//   * TMEM is allocated / freed with the cta_group::2 forms by the same warp of both CTAs.
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "conv_epilogue.cuh"
#include "conv_params.h"
#include "conv_umma.cuh"

namespace rfv {

template <int BN>
struct PairCfg {
    static constexpr int B_BYTES = (BN / 2) * 128;                  // this CTA's half of the weight slice
    static constexpr int STAGE_BYTES = UMMA_A_BYTES + B_BYTES;      // 32 KB (BN = 256) / 24 KB (BN = 128)
    static constexpr int STAGES = BN == 256 ? 6 : 8;
    static constexpr int TMEM_COLS = 2 * BN;                        // two accumulator stages
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
};

template <int U2_BN>
__global__ void __launch_bounds__(UMMA_THREADS, 1)
conv_umma2_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
                  const __grid_constant__ CUtensorMap mapA2, const __grid_constant__ CUtensorMap mapW, const ConvParams p,
                  const UmmaGeom g) {
    constexpr int U2_B_BYTES = PairCfg<U2_BN>::B_BYTES, U2_STAGE_BYTES = PairCfg<U2_BN>::STAGE_BYTES, U2_STAGES = PairCfg<U2_BN>::STAGES;
    constexpr int U2_TMEM = PairCfg<U2_BN>::TMEM_COLS;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + U2_STAGES * UMMA_A_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + U2_STAGES * U2_STAGE_BYTES);
    uint64_t* full_bar = bars;                      // [STAGES]  both CTAs' TMA -> the leader's MMA warp (leader's copy is used)
    uint64_t* empty_bar = bars + U2_STAGES;         // [STAGES]  MMA -> TMA, multicast to both CTAs
    uint64_t* tfull_bar = bars + 2 * U2_STAGES;     // [2]       MMA -> epilogue, multicast to both CTAs
    uint64_t* tempty_bar = tfull_bar + 2;           // [2]       both CTAs' epilogues -> the leader's MMA warp
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int crank = (int)cluster_ctarank();
    const bool leader = crank == 0;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&mapA0);
        tma_prefetch_desc(&mapW);
        for (int s = 0; s < U2_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&tfull_bar[s], 1); mbar_init(&tempty_bar[s], 256); }
        mbar_fence_init();
    }
    if (warp == 2) tmem_alloc_pair(tmem_slot, U2_TMEM);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();   // the peer's barriers exist before anything is signalled into them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (warp != 1) pdl_wait();   // programmatic dependent launch (common.cuh): everything above overlapped the predecessor's tail
    if (threadIdx.x == 0) pdl_launch();

    const int nkb0 = g.taps * g.cch0;
    const int nkb = nkb0 + g.cch1a + g.cch1b;
    // super-tile = (pair of consecutive m-tiles, n-tile); this CTA takes m-tile 2*group + crank.  An m-tile past the end is
    // computed on zero-filled boxes and never stored.
    const int m_groups = (g.m_tiles + 1) / 2;
    const int phases = g.ups ? 4 : 1;   // sub-pixel phases of a nearest-x2-upsample + 3x3 conv (see conv_umma.cuh)
    const int total_tiles = m_groups * g.n_tiles * phases;
    const int tile0 = blockIdx.x >> 1, tile_step = gridDim.x >> 1;
    const int box_shift = g.bw_shift + g.bh_shift;
    const int tiles_per_img = g.tiles_w * g.tiles_h;

    if (warp == 0) {
        // ===================== TMA producer (both CTAs): own pixel box + own half of the weight slice =====================
        uint32_t stage = 0, phase = 0;
        for (int tile = tile0; tile < total_tiles; tile += tile_step) {
            const int rest = tile / g.n_tiles, nt = tile - rest * g.n_tiles;
            const int mt = (rest / phases) * 2 + crank, sp = rest % phases;
            const int py = sp >> 1, px = sp & 1;
            int n0, h0, w0;
            if (box_shift >= 7) {
                n0 = mt / tiles_per_img;
                const int r = mt - n0 * tiles_per_img;
                h0 = (r / g.tiles_w) << g.bh_shift;
                w0 = (r - (r / g.tiles_w) * g.tiles_w) << g.bw_shift;
            } else {
                n0 = mt << (7 - box_shift);
                h0 = 0;
                w0 = 0;
            }
            for (int kb = 0; kb < nkb; ++kb) {
                mbar_wait(&empty_bar[stage], phase ^ 1);
                if (elect_one()) {
                    if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * U2_STAGE_BYTES);   // the bytes of BOTH CTAs land here
                    uint8_t* sa = smem_a + stage * UMMA_A_BYTES;
                    uint8_t* sb = smem_b + stage * U2_B_BYTES;
                    if (kb < nkb0) {
                        const int tap = kb / g.cch0, cc = kb - tap * g.cch0;
                        int dy = 0, dx = 0;
                        if (g.taps == 9) { dy = tap / 3; dx = tap - dy * 3; dy -= 1; dx -= 1; }
                        else if (g.taps == 4) { dy = (tap >> 1) + py - 1; dx = (tap & 1) + px - 1; }
                        tma_load_4d_pair(sa, &mapA0, &full_bar[stage], cc * 64, w0 + dx, h0 + dy, n0);
                    } else {
                        const int k1 = kb - nkb0;
                        if (k1 < g.cch1a) tma_load_4d_pair(sa, &mapA1, &full_bar[stage], k1 * 64, w0, h0, n0);
                        else tma_load_4d_pair(sa, &mapA2, &full_bar[stage], (k1 - g.cch1a) * 64, w0, h0, n0);
                    }
                    tma_load_2d_pair(sb, &mapW, &full_bar[stage], kb * 64, sp * p.Cout + nt * U2_BN + crank * (U2_BN / 2));
                }
                __syncwarp();
                if (++stage == U2_STAGES) { stage = 0; phase ^= 1; }
            }
        }
        __syncwarp();
    } else if (warp == 1 && leader) {
        // ===================== MMA issuer (leader only): one M = 256 instruction per K step for the pair =====================
        constexpr uint32_t idesc = umma_idesc_bf16(256, U2_BN);
        uint32_t stage = 0, phase = 0, it = 0;
        for (int tile = tile0; tile < total_tiles; tile += tile_step, ++it) {
            const uint32_t as = it & 1, aphase = (it >> 1) & 1;
            mbar_wait(&tempty_bar[as], aphase ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + as * U2_BN;
            for (int kb = 0; kb < nkb; ++kb) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                if (elect_one()) {
                    const uint64_t adesc = umma_desc_sw128(smem_u32(smem_a + stage * UMMA_A_BYTES));
                    const uint64_t bdesc = umma_desc_sw128(smem_u32(smem_b + stage * U2_B_BYTES));
#pragma unroll
                    for (int j = 0; j < 4; ++j) umma_bf16_pair(d_tmem, adesc + 2 * j, bdesc + 2 * j, idesc, (kb | j) != 0);
                    umma_commit_pair(&empty_bar[stage]);
                    if (kb == nkb - 1) umma_commit_pair(&tfull_bar[as]);
                }
                __syncwarp();
                if (++stage == U2_STAGES) { stage = 0; phase ^= 1; }
            }
        }
        __syncwarp();
    } else if (warp >= 4) {
        // ===================== epilogue (both CTAs): own 128 pixels x 256 channels =====================
        const int q = warp & 3;
        const int eg = (warp - 4) >> 2;
        const int r = q * 32 + lane;
        uint32_t it = eg;
        for (int tile = tile0 + eg * tile_step; tile < total_tiles; tile += 2 * tile_step, it += 2) {
            const int rest = tile / g.n_tiles, nt = tile - rest * g.n_tiles;
            const int mt = (rest / phases) * 2 + crank, sp = rest % phases;
            int n, h, w;
            if (box_shift >= 7) {
                n = mt / tiles_per_img;
                const int rr = mt - n * tiles_per_img;
                h = ((rr / g.tiles_w) << g.bh_shift) + (r >> g.bw_shift);
                w = ((rr - (rr / g.tiles_w) * g.tiles_w) << g.bw_shift) + (r & ((1 << g.bw_shift) - 1));
            } else {
                n = (mt << (7 - box_shift)) + (r >> box_shift);
                h = (r >> g.bw_shift) & ((1 << g.bh_shift) - 1);
                w = r & ((1 << g.bw_shift) - 1);
            }
            if (g.ups) { h = 2 * h + (sp >> 1); w = 2 * w + (sp & 1); }
            const bool valid = n < p.B && mt < g.m_tiles;
            const size_t pix = ((size_t)n * p.Ho + h) * p.Wo + w;
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + eg * U2_BN;
            conv_epilogue_tile<U2_BN>(p, taddr, n, valid, valid, pix, nt, lane, &tfull_bar[eg], (it >> 1) & 1, &tempty_bar[eg], true, true);
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();   // nobody exits while the peer may still signal into, or read operands from, its shared memory
    if (warp == 2) tmem_dealloc_pair(tmem_base, U2_TMEM);
}

}  // namespace rfv
