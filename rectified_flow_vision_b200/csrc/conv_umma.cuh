// Implicit-GEMM convolution on the 5th-gen tensor cores: TMA (tiled, zero-filled halo) -> 128B-swizzled shared
// memory -> tcgen05.mma with fp32 accumulators in TMEM -> tcgen05.ld epilogue.  Persistent, warp-specialised:
//   warp 0  TMA producer (one lane)      warp 1  MMA issuer (one lane)      warp 2  TMEM allocator
//   warps 4-11 epilogue (one TMEM lane = one output pixel per thread; two warps per lane quadrant split the columns)
// GEMM view: M = 128 output pixels (a bn x bh x bw box of the NHWC tensor), N = BN output channels,
// K = taps x C_in walked in chunks of 64 channels (one 128-byte swizzle row per pixel).  The im2col gather is done
// by the TMA unit itself: for tap (dy,dx) the same 4-D box is fetched at (w0+dx-1, h0+dy-1); out-of-bounds
// coordinates are zero-filled, which is exactly the conv's zero padding.  Stride-2 convs fetch from four
// parity views of the input (one tensor map each).  A second K segment (1x1 shortcut over a virtual concat of up
// to two tensors) accumulates into the same TMEM tile, so a ResidualBlock's conv2 + shortcut is one kernel.
// Nearest-x2 upsample + 3x3 conv (models/unet.py:215-218) is never materialised: it is evaluated as four sub-pixel
// phases, out[2i+py, 2j+px] = sum_{a,b in {0,1}} W'[py,px][a,b] . in[i+a+py-1, j+b+px-1], with the 3x3 taps that
// alias onto the same input pixel pre-summed at weight-pack time (2.25x fewer MACs than the literal formulation).
// Replaces nn.Conv2d call sites models/unet.py:38,41,51,76,77,185 (see SURVEY §2.4 for the shapes).
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "conv_epilogue.cuh"
#include "conv_params.h"

namespace rfv {

struct UmmaGeom {
    int bw_shift, bh_shift;   // log2 box width / height (pixels); images per box = 128 >> (bw_shift + bh_shift)
    int tiles_w, tiles_h;     // boxes per image along W and H (1,1 when a box covers whole images)
    int m_tiles, n_tiles;
    int cch0, cch1a, cch1b;   // 64-channel chunks per tap of segment 0 / per source of segment 1
    int taps;                 // 9, 1, or 4 (sub-pixel phases of a nearest-x2-upsample + 3x3 conv)
    int stride2;              // 1: segment 0 reads the four parity maps
    int ups;                  // 1: output is 2x the input; tiles enumerate (pixel box, phase py/px, channel tile)
    int cluster;              // CTAs per cluster (1, 2 or 4): they run m-tiles cl*g+rank of the same (phase, n-tile) in
                              // lock-step and each multicasts 1/cluster of the weight slice to all of them
};

constexpr int UMMA_BM = 128;
constexpr int UMMA_THREADS = 384;  // 4 control warps + 8 epilogue warps
constexpr int UMMA_A_BYTES = UMMA_BM * 128;  // 128 pixels x 64 bf16

template <int BN>
struct UmmaCfg {
    static constexpr int B_BYTES = BN * 128;
    static constexpr int STAGE_BYTES = UMMA_A_BYTES + B_BYTES;
    static constexpr int STAGES = (BN == 256) ? 4 : (BN == 128 ? 6 : 8);
    static constexpr int TMEM_COLS = 2 * BN;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

template <int BN>
__global__ void __launch_bounds__(UMMA_THREADS, 1)
conv_umma_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
                 const __grid_constant__ CUtensorMap mapA2, const __grid_constant__ CUtensorMap mapA3,
                 const __grid_constant__ CUtensorMap mapW, const ConvParams p, const UmmaGeom g) {
    using Cfg = UmmaCfg<BN>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + Cfg::STAGES * UMMA_A_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
    uint64_t* full_bar = bars;                       // [STAGES]  TMA -> MMA
    uint64_t* empty_bar = bars + Cfg::STAGES;        // [STAGES]  MMA -> TMA
    uint64_t* tfull_bar = bars + 2 * Cfg::STAGES;    // [2]       MMA -> epilogue
    uint64_t* tempty_bar = tfull_bar + 2;            // [2]       epilogue -> MMA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int CL = g.cluster;
    const int crank = CL > 1 ? (int)cluster_ctarank() : 0;
    const uint16_t cmask = (uint16_t)((1u << CL) - 1);

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&mapA0);
        tma_prefetch_desc(&mapW);
        for (int s = 0; s < Cfg::STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], CL);  // one arrival per CTA of the cluster (multicast commit)
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&tfull_bar[s], 1);
            mbar_init(&tempty_bar[s], 128);
        }
        mbar_fence_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync_all();  // peers' barriers must be initialised before anyone multicasts into them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (warp != 1) pdl_wait();   // programmatic dependent launch (common.cuh): everything above overlapped the predecessor's tail
    if (threadIdx.x == 0) pdl_launch();

    const int nkb0 = g.taps * g.cch0;
    const int nkb = nkb0 + g.cch1a + g.cch1b;
    const int phases = g.ups ? 4 : 1;
    // Tiles are enumerated per cluster: "super-tile" = (group of CL consecutive m-tiles, phase, n-tile); CTA `crank`
    // takes m-tile group*CL + crank.  m-tiles past the end are computed on zero-filled boxes and never stored.
    const int m_groups = (g.m_tiles + CL - 1) / CL;
    const int total_tiles = m_groups * g.n_tiles * phases;
    const int tile0 = blockIdx.x / CL, tile_step = gridDim.x / CL;
    const int box_shift = g.bw_shift + g.bh_shift;        // log2 pixels per image inside a box
    const int tiles_per_img = g.tiles_w * g.tiles_h;

    if (warp == 0) {
        {
            // ===================== TMA producer (whole warp walks the loop, one elected lane issues) =====================
            uint32_t stage = 0, phase = 0;
            for (int tile = tile0; tile < total_tiles; tile += tile_step) {
                const int rest = tile / g.n_tiles, nt = tile - rest * g.n_tiles;
                const int mt = (rest / phases) * CL + crank, sp = rest % phases;  // sp: sub-pixel phase of an upsample conv
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
                    mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
                    uint8_t* sa = smem_a + stage * UMMA_A_BYTES;
                    uint8_t* sb = smem_b + stage * Cfg::B_BYTES;
                    if (kb < nkb0) {
                        const int tap = kb / g.cch0, cc = kb - tap * g.cch0;
                        int dy = 0, dx = 0;
                        if (g.taps == 9) { dy = tap / 3; dx = tap - dy * 3; dy -= 1; dx -= 1; }
                        else if (g.taps == 4) { dy = (tap >> 1) + py - 1; dx = (tap & 1) + px - 1; }
                        if (!g.stride2) {
                            tma_load_4d(sa, &mapA0, &full_bar[stage], cc * 64, w0 + dx, h0 + dy, n0);
                        } else if (g.taps == 16) {
                            // data gradient of a nearest-x2-upsample + 3x3 conv: tap = (phase, a, b) reads the parity view
                            // `phase` of the high-resolution gradient at (i + 1 - a - py, j + 1 - b - px)
                            const int ph = tap >> 2, a = (tap >> 1) & 1, b = tap & 1;
                            const CUtensorMap* mp = ph == 0 ? &mapA0 : (ph == 1 ? &mapA1 : (ph == 2 ? &mapA2 : &mapA3));
                            tma_load_4d(sa, mp, &full_bar[stage], cc * 64, w0 + 1 - b - (ph & 1), h0 + 1 - a - (ph >> 1), n0);
                        } else {
                            // input row 2*ho + dy: dy=-1 -> odd rows at ho-1; dy=0 -> even rows at ho; dy=+1 -> odd rows at ho
                            const int ph = (dy == 0) ? 0 : 1, pw = (dx == 0) ? 0 : 1;
                            const int hh = h0 + (dy < 0 ? -1 : 0), ww = w0 + (dx < 0 ? -1 : 0);
                            const CUtensorMap* mp = ph ? (pw ? &mapA3 : &mapA2) : (pw ? &mapA1 : &mapA0);
                            tma_load_4d(sa, mp, &full_bar[stage], cc * 64, ww, hh, n0);
                        }
                    } else {
                        const int k1 = kb - nkb0;
                        if (k1 < g.cch1a) tma_load_4d(sa, &mapA1, &full_bar[stage], k1 * 64, w0, h0, n0);
                        else tma_load_4d(sa, &mapA2, &full_bar[stage], (k1 - g.cch1a) * 64, w0, h0, n0);
                    }
                    if (CL == 1) tma_load_2d(sb, &mapW, &full_bar[stage], kb * 64, sp * p.Cout + nt * BN);
                    else  // this CTA's 1/CL of the weight rows, delivered to every CTA of the cluster
                        tma_load_2d_mc(sb + crank * (Cfg::B_BYTES / CL), &mapW, &full_bar[stage], kb * 64,
                                       sp * p.Cout + nt * BN + crank * (BN / CL), cmask);
                    }
                    __syncwarp();
                    if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        {
            // ===================== MMA issuer (whole warp walks the loop, one elected lane issues) =====================
            constexpr uint32_t idesc = umma_idesc_bf16(UMMA_BM, BN);
            uint32_t stage = 0, phase = 0, it = 0;
            for (int tile = tile0; tile < total_tiles; tile += tile_step, ++it) {
                const uint32_t as = it & 1, aphase = (it >> 1) & 1;
                mbar_wait(&tempty_bar[as], aphase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * BN;
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint64_t adesc = umma_desc_sw128(smem_u32(smem_a + stage * UMMA_A_BYTES));
                        const uint64_t bdesc = umma_desc_sw128(smem_u32(smem_b + stage * Cfg::B_BYTES));
#pragma unroll
                        for (int j = 0; j < 4; ++j)  // 4 x (K = 16) per 64-channel chunk: +32 bytes = +2 in descriptor units
                            umma_bf16(d_tmem, adesc + 2 * j, bdesc + 2 * j, idesc, (kb | j) != 0);
                        if (CL == 1) umma_commit(&empty_bar[stage]);
                        else umma_commit_mc(&empty_bar[stage], cmask);
                        if (kb == nkb - 1) umma_commit(&tfull_bar[as]);
                    }
                    __syncwarp();
                    if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
        __syncwarp();
    } else if (warp >= 4) {
        // ===================== epilogue: warp-group g drains accumulator stage g (tiles it = g, g+2, ...) =====================
        const int q = warp & 3;                 // TMEM lane quadrant this warp may access (warp id mod 4)
        const int grp = (warp - 4) >> 2;
        const int r = q * 32 + lane;            // row of the tile = output pixel
        uint32_t it = grp;
        for (int tile = tile0 + grp * tile_step; tile < total_tiles; tile += 2 * tile_step, it += 2) {
            const int rest = tile / g.n_tiles, nt = tile - rest * g.n_tiles;
            const int mt = (rest / phases) * CL + crank, sp = rest % phases;
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
            const bool valid = n < p.B;
            const size_t pix = ((size_t)n * p.Ho + h) * p.Wo + w;
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + grp * BN;
            conv_epilogue_tile<BN>(p, taddr, n, valid, valid, pix, nt, lane, &tfull_bar[grp], (it >> 1) & 1, &tempty_bar[grp]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync_all();  // nobody exits while a peer may still multicast into / arrive on its shared memory
    if (warp == 2) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

}  // namespace rfv
