// 3x3 stride-1 convolution on tcgen05 with the WEIGHTS as the A operand and up to 256 PIXELS as the B operand.
//
// Why: in SS mode an M=128 tcgen05.mma re-reads its 128-row A operand from shared memory in ~82 cycles whatever N is
// (tools/micro/umma_rate2.cu: N=64 82.8, N=128 81.8, N=192 96, N=256 128 cycles per instruction), so the pixel-major
// formulation (A = 128 positions, B = C_out weights: round 1's conv_halo kernels) cannot exceed 39 % of the tensor pipe on 64-channel
// layers and 78 % on 128-channel ones.  Here the roles are swapped: A = one 128-row weight block (16 KB, K-major), B = N
// consecutive positions of the flat padded image (row pitch W+1: the column shared between rows is zero-filled by the TMA unit, so tap (dy,dx) is the constant row shift
// dy*(W+1)+dx of ONE shared-memory box per 64-channel chunk -- nine UMMA descriptors over the same bytes; the 128B swizzle is
// applied on absolute shared-memory address bits, so 128-byte-aligned shifted starts work with base offset 0),
// N = 192..256, which is the shape the pipe runs at full rate.  The accumulator is D[channel (TMEM lane)][position (column)].
//
//   * C_out tile = 128: block rows = output channels, nine taps per 64-channel chunk.
//   * C_out tile = 64 (PAIR): the 128 rows hold TWO taps of the same 64 channels -- within every 16 rows, rows 0-7 are tap
//     (dy,-1) and rows 8-15 tap (dy,0) of the same 8 channels.  Both halves see the same B operand (start = shift of the first
//     tap), so the second tap's contribution to position p lands one column to the right: out[p] = top[p] + bot[p+1]; tiles
//     advance by N-1 positions.  Three pair blocks + three single blocks ((dy,+1), second half zero) per chunk: 6 MMA groups
//     for 9 taps = 75 % of the pipe (a common column offset admits at most three disjoint pairs in a 3x3 stencil).
//   * The 1x1 shortcut of a ResidualBlock is extra K chunks at the centre shift ; the IDENTITY residual
//     is one more: a chunk of the residual tensor multiplied by an identity block (exact in fp32 accumulation), so the
//     epilogue never touches it.  (Tried instead: adding the residual in the epilogue from global memory, 2-byte loads in the
//     fragment layout -- 0.55 ms on a 64->64 layer issued per half-unit, 0.31 ms with L2 prefetch and loads one half-unit
//     ahead, against 0.24-0.26 ms for the identity chunk: every half-unit exposes a memory latency the MMA path hides.)
//   * Epilogue in the m16n8 fragment layout: tcgen05.ld.16x256b gives thread t rows t/4 and t/4+8 of a 16-lane group and
//     columns 2(t%4), 2(t%4)+1 of every 8-column group.  In PAIR mode both taps of a channel sit in the SAME thread, so the
//     one-column shift is one in-thread add plus one shuffle per two outputs.  GroupNorm partial sums are per-thread running sums
//     (channels are lanes) reduced once per tile; the bf16 result is transposed to [pixel][channel] rows with stmatrix.trans
//     into a per-warp staging buffer and leaves with 16-byte stores (register mapping verified by tools/micro/frag_test.cu).
//   * FUSE: GroupNorm(+SiLU) is applied to the segment-0 boxes in shared memory by 8 transform warps between the TMA write
//     and the MMAs (coefficients per (image, channel) from gn_coef_kernel).
//
//   warp 0  box producer   warp 1  MMA issuer   warp 2  TMEM allocator   warp 3  weight producer   warps 4-19  epilogue
//   (FUSE: warps 4-15 epilogue, warps 16-23 transform)
// Replaces nn.Conv2d call sites models/unet.py:38,41(+51) at the 64x64 / 32x32 (/128x128) levels, and their data gradients.
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "conv_params.h"

namespace rfv {

struct WaGeom {
    int W, H, pitch, rows;        // pitch = W + 1; rows = box height
    int N, adv;                   // MMA N (positions per tile, multiple of 32) / positions a tile advances by (N or N-1)
    int tiles_per_img, m_tiles, n_tiles;
    int cch0, cch0a;              // 64-channel chunks of segment 0 (the first cch0a from map A0, the rest from A0b)
    int cch1a, cch1b, cchr;       // chunks of the two shortcut sources / of the identity residual (cchr: 0 = none)
    int slots0;                   // weight blocks per segment-0 chunk: 9, or 6 (PAIR)
    int nblk;                     // weight blocks per channel tile in the packed buffer (incl. the identity blocks)
    int box_bytes, stage_bytes;   // rows*pitch*128 / 1024-aligned stage (box + one trailing zero row)
    int a_stages, w_stages, resident;
    int ctile;                    // output channels per tile: 128, or 64 (PAIR)
    uint32_t inv_pitch;           // ceil(2^32 / pitch)
    uint32_t inv_tpi;             // ceil(2^32 / tiles_per_img)
    int tstages, tstride;         // accumulator stages in TMEM (2, or 3 when N <= 160) and their column stride
    int pf;                       // L2 prefetch distance in tiles (0 = off)
    int rs;                       // box rows per TMA request (a box is fetched as ceil(rows / rs) requests: one request streams at
                                  // only ~12 B/clk, several in flight overlap)
    int tblk;                     // resident weight blocks held in TENSOR MEMORY (the first tblk blocks of a tile's sequence, all in chunk 0;
                                  // 32 columns each at column tcol0 + 32 b) and multiplied in TS mode (A operand from TMEM); the rest
                                  // of the resident blocks stay in shared memory.  0 = off.
    int tcol0;
    const void* wa;               // the packed weight blocks in global memory (the TMEM preload reads them directly)
    int dbg;                      // experiments (RFV_WA_DBG): 1 = epilogue only waits / releases, 2 = no MMAs issued, 4 = no global stores,
                                  // 8 = no TMA loads (barriers only)
};

// Epilogue warps: four per TMEM lane quarter (16 half-units of an N = 256 tile split evenly, four each), three in the FUSE
// kernels whose register file also has to hold the transform warps.  ncu (r2f, 64->64 PAIR layer, 12 warps): 780 instructions
// per warp and tile at an IPC of 0.18 per warp -- the warps are bound by their own dependent-issue latency, not by the
// schedulers (issue slots 56 % busy), so more warps with less work each is what shortens the per-tile epilogue.
__host__ __device__ constexpr int wa_ewarps(bool fuse) { return fuse ? 12 : 16; }
constexpr int WA_TWARPS = 8;                               // transform warps (FUSE)
__host__ __device__ constexpr int wa_threads(bool fuse) { return (4 + wa_ewarps(fuse) + (fuse ? WA_TWARPS : 0)) * 32; }
constexpr int WA_BLK = 128 * 128;       // one weight block: 128 rows x 64 bf16

__host__ __device__ constexpr int wa_stage_pitch(bool pair) { return pair ? 48 : 80; }   // bytes per staged pixel row (+16 pad)
__host__ __device__ constexpr int wa_staging_bytes(bool pair, bool fuse) { return wa_ewarps(fuse) * 16 * wa_stage_pitch(pair); }   // 16 pixel rows per warp

// scale / shift per (image, channel) of a GroupNorm(8) over a virtual concat of up to two tensors:
// coef[(n*C + c)*2] = rstd*gamma, coef[..+1] = beta - mean*rstd*gamma
__global__ void __launch_bounds__(256) gn_coef_kernel(const float* __restrict__ stats_a, const float* __restrict__ stats_b,
                                                      const float* __restrict__ gamma, const float* __restrict__ beta,
                                                      float* __restrict__ coef, int Ca, int Cb, int HW, int slab_shift, float eps) {
    pdl_wait();   // programmatic dependent launch (common.cuh): no-ops unless launched with the attribute
    pdl_launch();
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

// [Cout][Ktot] K-major bf16 (K = tap*C0 + c | K0 + shortcut c)  ->  [n_tiles][nblk][128 rows][64] blocks in consumption order
__global__ void pack_wa_kernel(const bf16* __restrict__ src, bf16* __restrict__ dst, int Cout, int C0, int K0, int Ktot, int cch0,
                               int cch1, int cchr, int pair, int n_tiles) {
    const int slots0 = pair ? 6 : 9;
    const int nblk = cch0 * slots0 + cch1 + cchr;
    const size_t total = (size_t)n_tiles * nblk * (WA_BLK / 2);   // elements
    const bf16 zero = __float2bfloat16_rn(0.f), one = __float2bfloat16_rn(1.f);
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const int k = (int)(e & 63), r = (int)((e >> 6) & 127);
        const int blk = (int)((e >> 13) % nblk), nt = (int)((e >> 13) / nblk);
        int co, half = 0;
        if (pair) { co = nt * 64 + (r >> 4) * 8 + (r & 7); half = (r >> 3) & 1; }
        else co = nt * 128 + r;
        bf16 v = zero;
        if (blk < cch0 * slots0) {
            const int ch = blk / slots0, slot = blk - ch * slots0;
            int tap = slot;
            if (pair) tap = slot < 3 ? slot * 3 + half : (half ? -1 : (slot - 3) * 3 + 2);
            if (tap >= 0) v = src[(size_t)co * Ktot + (size_t)tap * C0 + ch * 64 + k];
        } else if (blk < cch0 * slots0 + cch1) {
            const int j = blk - cch0 * slots0;
            if (!half) v = src[(size_t)co * Ktot + K0 + j * 64 + k];
        } else {
            const int j = blk - cch0 * slots0 - cch1;
            const int local = pair ? (r >> 4) * 8 + (r & 7) : r;
            if (!half && local == j * 64 + k) v = one;
        }
        dst[e] = v;
    }
}

__device__ __forceinline__ void tmem_ld_16x256_x4(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.16x256b.x4.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_16x256_x2(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_ld_16x256_x1(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(taddr)
                 : "memory");
}
// one 32-lane x 32-column slab of tensor memory from registers: thread = lane (row), register j = 32-bit column j
__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%32], "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31};"
        ::"r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
          "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
          "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]),
          "r"(r[31]), "r"(taddr)
        : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc]: the A operand (128 rows = lanes, K = 16 bf16 = 8 columns per instruction) comes from
// tensor memory (tools/micro/ts_mma_test.cu: layout and results verified; same rate as the shared-memory form at N >= 164)
__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void stmatrix_x4_trans(uint32_t addr, uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3) {
    asm volatile("stmatrix.sync.aligned.m8n8.x4.trans.shared.b16 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(r0), "r"(r1), "r"(r2), "r"(r3)
                 : "memory");
}

// (not inlined: its 32 staging registers must not count against the epilogue's register budget)
__device__ __noinline__ void wa_preload_tmem(const void* wa, int tblk, uint32_t taddr, int row) {
    for (int b = 0; b < tblk; ++b) {
        const uint4* src = reinterpret_cast<const uint4*>(static_cast<const bf16*>(wa) + ((size_t)b * 128 + row) * 64);
        uint32_t v[32];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const uint4 t = __ldg(src + i);
            v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
        }
        tmem_st_32x32(taddr + (uint32_t)(32 * b), v);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

template <bool PAIR, bool FUSE>
__global__ void __launch_bounds__(wa_threads(FUSE), 1)
conv_wa_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA0b,
               const __grid_constant__ CUtensorMap mapA1, const __grid_constant__ CUtensorMap mapA2,
               const __grid_constant__ CUtensorMap mapR, const __grid_constant__ CUtensorMap mapW, const ConvParams p, const WaGeom g) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    // [1 KB guard][box ring][weight blocks: ring or resident][epilogue staging: 8 warps][barriers]
    uint8_t* smem_x = smem + 1024;
    uint8_t* smem_w = smem_x + (size_t)g.a_stages * g.stage_bytes;
    const int nchunks = g.cch0 + g.cch1a + g.cch1b + g.cchr;
    const int nblk_used = g.cch0 * g.slots0 + g.cch1a + g.cch1b + g.cchr;
    const int w_slots = g.resident ? nblk_used - g.tblk : g.w_stages;
    uint8_t* smem_o = smem_w + (size_t)w_slots * WA_BLK;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_o + wa_staging_bytes(PAIR, FUSE));
    uint64_t* xfull = bars;                       // TMA -> (transform | MMA)
    uint64_t* xready = xfull + g.a_stages;        // transform -> MMA (FUSE only)
    uint64_t* xempty = xready + g.a_stages;       // MMA -> TMA
    uint64_t* wfull = xempty + g.a_stages;        // [w_stages] (slot 0 doubles as "resident weights landed")
    uint64_t* wempty = wfull + g.w_stages;
    uint64_t* tfull = wempty + g.w_stages;        // [<= 4] MMA -> epilogue
    uint64_t* tempty = tfull + 4;                 // [<= 4] epilogue -> MMA (one arrival per epilogue warp)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&mapA0);
        tma_prefetch_desc(&mapW);
        for (int s = 0; s < g.a_stages; ++s) { mbar_init(&xfull[s], 1); mbar_init(&xready[s], WA_TWARPS); mbar_init(&xempty[s], 1); }
        for (int s = 0; s < g.w_stages; ++s) { mbar_init(&wfull[s], 1); mbar_init(&wempty[s], 1); }
        for (int s = 0; s < 4; ++s) { mbar_init(&tfull[s], 1); mbar_init(&tempty[s], wa_ewarps(FUSE)); }
        mbar_fence_init();
    }
    // the row after each box must read as zero (tap (+1,+1) of the last position of the last box row lands there)
    if (threadIdx.x < 32)
        for (int s = 0; s < g.a_stages; ++s)
            reinterpret_cast<uint32_t*>(smem_x + (size_t)s * g.stage_bytes + g.box_bytes)[threadIdx.x] = 0u;
    if (warp == 2) tmem_alloc(tmem_slot, 512);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int total_tiles = g.m_tiles * g.n_tiles;
    if (g.tblk > 0) {
        // weight blocks that live in tensor memory for the whole launch: warp 4 + q writes rows 32 q .. 32 q + 31 (its lane
        // quarter), thread = row, 32 registers = the row's 64 bf16 (column j = K elements 2j, 2j+1).  Weights are constant
        // while a sampling chain runs, so this precedes pdl_wait like the resident shared-memory blocks.
        if (warp >= 4 && warp < 8) wa_preload_tmem(g.wa, g.tblk, tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)g.tcol0, (warp & 3) * 32 + lane);
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
    }
    // programmatic dependent launch: the box producer, the transform and the epilogue warps touch activations / statistics /
    // coefficients of neighbouring kernels and wait for the predecessor here; the weight producer (weights are constant
    // while a sampling chain runs) and the MMA warp (no global accesses) go ahead, so the resident weight blocks land under
    // the predecessor's tail
    if (warp != 1 && warp != 3) pdl_wait();
    if (threadIdx.x == 0) pdl_launch();

    if (warp < 4) {
        if (FUSE) asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
        if (warp == 0) {
            // ===================== box producer: one halo box per (tile, 64-channel chunk) =====================
            uint32_t st = 0, ph = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                int mt = tile, nt = 0;
                if (g.n_tiles > 1) { mt = tile / g.n_tiles; nt = tile - mt * g.n_tiles; }
                const int n = (int)__umulhi((uint32_t)mt, g.inv_tpi), ti = mt - n * g.tiles_per_img;
                const int rbox = (int)__umulhi((uint32_t)(ti * g.adv), g.inv_pitch) - 1;
                if (g.pf > 0) {
                    // L2 prefetch of the boxes of the tile this CTA runs g.pf iterations from now: the ring only keeps one or two
                    // boxes in flight, and a box that misses L2 streams in at ~20 B/clk (measured: loads alone take 1.3 us per 50 KB)
                    const int tp = tile + g.pf * (int)gridDim.x;
                    if (tp < total_tiles && elect_one()) {
                        int mtp = tp, ntp = 0;
                        if (g.n_tiles > 1) { mtp = tp / g.n_tiles; ntp = tp - mtp * g.n_tiles; }
                        const int np = (int)__umulhi((uint32_t)mtp, g.inv_tpi), tip = mtp - np * g.tiles_per_img;
                        const int rbp = (int)__umulhi((uint32_t)(tip * g.adv), g.inv_pitch) - 1;
                        for (int ch = 0; ch < nchunks; ++ch) {
                            int c = ch;
                            const CUtensorMap* mp;
                            if (c < g.cch0a) mp = &mapA0;
                            else if (c < g.cch0) { mp = &mapA0b; c -= g.cch0a; }
                            else if ((c -= g.cch0) < g.cch1a) mp = &mapA1;
                            else if ((c -= g.cch1a) < g.cch1b) mp = &mapA2;
                            else { mp = &mapR; c = ntp * (g.ctile / 64) + (c - g.cch1b); }
                            const int trim = (ch >= g.cch0 && g.rs == 1) ? 1 : 0;   // centre-only chunks: see the loads below
                            for (int r0 = trim; r0 < g.rows - trim; r0 += g.rs) tma_prefetch_4d(mp, c * 64, -1, rbp + r0, np);
                        }
                    }
                    __syncwarp();
                }
                for (int ch = 0; ch < nchunks; ++ch) {
                    mbar_wait(&xempty[st], ph ^ 1);
                    if (g.dbg & 8) { if (elect_one()) mbar_arrive(&xfull[st]); }
                    else if (elect_one()) {
                        // shortcut / residual chunks are multiplied at the centre shift only: their MMAs read box rows
                        // 1 .. rows-2 (idx0 lies in [pitch, 2 pitch), idx0 + N - 1 < (rows - 1) pitch), so the two halo rows are
                        // not fetched (these layers are bound by the box loads: two to four boxes per tile at ~24 B/clk per SM)
                        const int trim = (ch >= g.cch0 && g.rs == 1) ? 1 : 0;
                        mbar_arrive_expect_tx(&xfull[st], g.box_bytes - 2 * trim * g.pitch * 128);
                        uint8_t* dst = smem_x + (size_t)st * g.stage_bytes;
                        int c = ch;
                        const CUtensorMap* mp;
                        if (c < g.cch0a) mp = &mapA0;
                        else if (c < g.cch0) { mp = &mapA0b; c -= g.cch0a; }
                        else if ((c -= g.cch0) < g.cch1a) mp = &mapA1;
                        else if ((c -= g.cch1a) < g.cch1b) mp = &mapA2;
                        else { mp = &mapR; c = nt * (g.ctile / 64) + (c - g.cch1b); }
                        for (int r0 = trim; r0 < g.rows - trim; r0 += g.rs)
                            tma_load_4d(dst + (size_t)r0 * g.pitch * 128, mp, &xfull[st], c * 64, -1, rbox + r0, n);
                    }
                    __syncwarp();
                    if (++st == (uint32_t)g.a_stages) { st = 0; ph ^= 1; }
                }
            }
        } else if (warp == 3) {
            // ===================== weight producer: 16 KB blocks in consumption order =====================
            if (g.resident) {
                if ((int)blockIdx.x < total_tiles && nblk_used > g.tblk && elect_one()) {
                    mbar_arrive_expect_tx(&wfull[0], (nblk_used - g.tblk) * WA_BLK);
                    for (int b = g.tblk; b < nblk_used; ++b) tma_load_2d(smem_w + (size_t)(b - g.tblk) * WA_BLK, &mapW, &wfull[0], 0, b * 128);
                }
            } else {
                uint32_t st = 0, ph = 0;
                for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                    const int nt = tile % g.n_tiles;
                    for (int b = 0; b < nblk_used; ++b) {
                        mbar_wait(&wempty[st], ph ^ 1);
                        if (g.dbg & 8) { if (elect_one()) mbar_arrive(&wfull[st]); }
                        else if (elect_one()) {
                            mbar_arrive_expect_tx(&wfull[st], WA_BLK);
                            tma_load_2d(smem_w + (size_t)st * WA_BLK, &mapW, &wfull[st], 0, (nt * g.nblk + b) * 128);
                        }
                        __syncwarp();
                        if (++st == (uint32_t)g.w_stages) { st = 0; ph ^= 1; }
                    }
                }
            }
        } else if (warp == 1) {
            // ===================== MMA issuer (whole warp walks the loop, one elected lane issues) =====================
            const int positions = g.H * g.pitch;
            uint32_t xst = 0, xph = 0, wst = 0, wph = 0, as = 0, aph = 0;
            if (g.resident && nblk_used > g.tblk && (int)blockIdx.x < total_tiles) mbar_wait(&wfull[0], 0);
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int mt = g.n_tiles > 1 ? tile / g.n_tiles : tile;
                const int ti = mt - (int)__umulhi((uint32_t)mt, g.inv_tpi) * g.tiles_per_img;
                const int q0 = ti * g.adv;
                const int idx0 = q0 - ((int)__umulhi((uint32_t)q0, g.inv_pitch) - 1) * g.pitch;   // row (128 B) of position q0 inside the box buffer
                // the last tile of an image covers what is left of it: a narrower instruction (N in steps of 16; PAIR needs
                // one more column for the second tap of its last position)
                const uint32_t idesc = umma_idesc_bf16(128, min(g.N, (min(g.adv, positions - q0) + (PAIR ? 1 : 0) + 15) & ~15));
                mbar_wait(&tempty[as], aph ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * g.tstride;
                int blk = 0;
                for (int ch = 0; ch < nchunks; ++ch) {
                    const bool seg0 = ch < g.cch0;
                    mbar_wait(FUSE ? &xready[xst] : &xfull[xst], xph);   // FUSE: the transform warps pass every chunk on
                    tc_fence_after();
                    const uint32_t xbase = smem_u32(smem_x + (size_t)xst * g.stage_bytes) + (uint32_t)(idx0 * 128);
                    const int nslots = seg0 ? g.slots0 : 1;
                    auto shift_of = [&](int sl) {
                        if (!seg0) return 0;
                        if (PAIR) return sl < 3 ? (sl - 1) * g.pitch - 1 : (sl - 4) * g.pitch + 1;
                        return (sl / 3 - 1) * g.pitch + (sl % 3 - 1);
                    };
                    int sl = 0;
                    // blocks held in tensor memory (the first g.tblk of chunk 0): TS-mode MMAs.  A loop of its own -- a predicated-off
                    // tcgen05.mma inside a shared loop body costs ~40 cycles per instruction (tools/micro/ts_mma_test.cu)
                    const int nts = ch == 0 ? min(g.tblk, nslots) : 0;
                    for (; sl < nts; ++sl, ++blk) {
                        if (elect_one()) {
                            const uint32_t a_tm = tmem_base + (uint32_t)(g.tcol0 + 32 * blk);
                            const uint64_t bdesc = umma_desc_sw128(xbase + (uint32_t)(shift_of(sl) * 128));
#pragma unroll
                            for (int j = 0; j < 4; ++j)
                                if (!(g.dbg & 2)) umma_bf16_ts(d_tmem, a_tm + 8 * j, bdesc + 2 * j, idesc, (sl | j) != 0);
                            if (sl == nslots - 1) {
                                umma_commit(&xempty[xst]);
                                if (ch == nchunks - 1) umma_commit(&tfull[as]);
                            }
                        }
                        __syncwarp();
                    }
                    for (; sl < nslots; ++sl, ++blk) {
                        const int shift = shift_of(sl);
                        uint32_t wslot = (uint32_t)(blk - g.tblk);
                        if (!g.resident) {
                            mbar_wait(&wfull[wst], wph);
                            tc_fence_after();
                            wslot = wst;
                        }
                        if (elect_one()) {
                            const uint64_t adesc = umma_desc_sw128(smem_u32(smem_w + (size_t)wslot * WA_BLK));
                            const uint64_t bdesc = umma_desc_sw128(xbase + (uint32_t)(shift * 128));
#pragma unroll
                            for (int j = 0; j < 4; ++j)
                                if (!(g.dbg & 2)) umma_bf16(d_tmem, adesc + 2 * j, bdesc + 2 * j, idesc, (ch | sl | j) != 0);
                            if (!g.resident) umma_commit(&wempty[wst]);
                            if (sl == nslots - 1) {
                                umma_commit(&xempty[xst]);
                                if (ch == nchunks - 1) umma_commit(&tfull[as]);
                            }
                        }
                        __syncwarp();
                        if (!g.resident && ++wst == (uint32_t)g.w_stages) { wst = 0; wph ^= 1; }
                    }
                    if (++xst == (uint32_t)g.a_stages) { xst = 0; xph ^= 1; }
                }
                if (++as == (uint32_t)g.tstages) { as = 0; aph ^= 1; }
            }
        }
    } else if (FUSE && warp >= 4 + wa_ewarps(FUSE)) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
        // ===================== transform: GroupNorm(+SiLU) in place on every segment-0 box =====================
        const int tt = threadIdx.x - (4 + wa_ewarps(FUSE)) * 32;
        constexpr int PL = WA_TWARPS * 4;              // position lanes
        const int j = tt & 7, pl = tt >> 3;            // logical 16-byte vector (8 channels) / position lane
        const int npos = g.rows * g.pitch;
        uint32_t st = 0, ph = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const int mt = g.n_tiles > 1 ? tile / g.n_tiles : tile;
            const int n = (int)__umulhi((uint32_t)mt, g.inv_tpi), ti = mt - n * g.tiles_per_img;
            const int rbox = (int)__umulhi((uint32_t)(ti * g.adv), g.inv_pitch) - 1;
            for (int ch = 0; ch < nchunks; ++ch) {
                if (ch < g.cch0) {
                    float sc[8], sh[8];
                    const float4* cp = reinterpret_cast<const float4*>(p.gn_coef + ((size_t)n * p.gn_C + ch * 64 + j * 8) * 2);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float4 v = cp[i];
                        sc[2 * i] = v.x; sh[2 * i] = v.y; sc[2 * i + 1] = v.z; sh[2 * i + 1] = v.w;
                    }
                    if (p.gn_silu) {   // silu(y) = h (1 + tanh h), h = y / 2: one MUFU op per element
#pragma unroll
                        for (int i = 0; i < 8; ++i) { sc[i] *= 0.5f; sh[i] *= 0.5f; }
                    }
                    mbar_wait(&xfull[st], ph);
                    // stage buffers are 1024-byte aligned and PL is a multiple of 8: the swizzle term j ^ (position & 7) of this
                    // thread's vector is the same for every position it visits
                    const uint32_t base = smem_u32(smem_x + (size_t)st * g.stage_bytes);
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
                    fence_async_smem();   // generic writes -> visible to the UMMA reads
                } else {
                    mbar_wait(&xfull[st], ph);   // shortcut / residual chunks pass through untouched
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&xready[st]);
                if (++st == (uint32_t)g.a_stages) { st = 0; ph ^= 1; }
            }
        }
    } else {
        // 768 threads x 80 registers = 61,440: control 4 x 32 x 32, transform 8 x 32 x 64, epilogue 12 x 32 x 104 = 60,416
        if (FUSE) asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
        // ===================== epilogue: all twelve warps drain one accumulator stage =====================
        // warp quarter q = TMEM lanes 32q..32q+31; the three warps of a quarter take 16-column half-units round-robin.
        // ncu on the first version (8 warps, 32-column units, per-column index arithmetic): 8.7 k warp-instructions per
        // 192-column PAIR tile against 2.3 k cycles of MMAs, issue slots 52 % busy -> per-column bookkeeping is done once per
        // half-unit with lane = column (one ballot gives the validity mask, one shuffle per stored row its pixel offset), the
        // statistics are reduced with the transposing butterfly, and three warps per scheduler hide each other's
        // TMEM-load -> shuffle -> stmatrix -> copy-out latency chain.
        const int q = warp & 3, slot = (warp - 4) >> 2;
        const int t4 = lane & 3, t8 = lane >> 2;
        constexpr int SP = wa_stage_pitch(PAIR);
        constexpr int CV = PAIR ? 2 : 4;                     // channel sub-blocks of 8 per thread
        constexpr int EQ = wa_ewarps(FUSE) / 4;              // warps per quarter
        const uint32_t stage = smem_u32(smem_o + (warp - 4) * 16 * SP);
        const int HW = g.H * g.W;
        uint32_t as = 0, aph = 0;
        // per-thread channels: cbase + 8*v + t8, v = 0..CV-1 (fragment rows t8 / t8+8 of the two 16-lane halves).  The per-channel
        // addend (bias, or bias + time projection of the tile's image) of the NEXT tile is fetched while the current one is
        // drained: loaded at the top of its own tile, the global-load latency sat on the epilogue's critical path (ncu: 7 % of
        // the kernel's stall samples on that one line)
        float addn[CV];
        auto load_addend = [&](int tile_, float* dst) {
            int mt_ = tile_, nt_ = 0;
            if (g.n_tiles > 1) { mt_ = tile_ / g.n_tiles; nt_ = tile_ - mt_ * g.n_tiles; }
            const int n_ = (int)__umulhi((uint32_t)mt_, g.inv_tpi);
            const float* ab = p.temb ? p.temb + (size_t)n_ * p.temb_stride : p.bias;
            const int cb_ = nt_ * g.ctile + q * (PAIR ? 16 : 32);
#pragma unroll
            for (int v = 0; v < CV; ++v) dst[v] = ab[cb_ + 8 * v + t8];
        };
        if ((int)blockIdx.x < total_tiles) load_addend(blockIdx.x, addn);
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            int mt = tile, nt = 0;
            if (g.n_tiles > 1) { mt = tile / g.n_tiles; nt = tile - mt * g.n_tiles; }
            const int n = (int)__umulhi((uint32_t)mt, g.inv_tpi), ti = mt - n * g.tiles_per_img;
            const int q0 = ti * g.adv;
            const int nhu = (min(g.adv, HW + g.H - q0) + 15) >> 4;   // half-units with positions of the image (fewer in its last tile)
            const int cbase = nt * g.ctile + q * (PAIR ? 16 : 32);   // first channel of this warp
            bf16* const obase = p.out + (size_t)n * HW * p.Cout + cbase;
            float addv[CV];
            float2 s1[CV], s2[CV];   // statistics as (even column, odd column) pairs: packed fp32x2 adds / FMAs
#pragma unroll
            for (int v = 0; v < CV; ++v) { addv[v] = addn[v]; s1[v] = make_float2(0.f, 0.f); s2[v] = make_float2(0.f, 0.f); }
            if (tile + (int)gridDim.x < total_tiles) load_addend(tile + gridDim.x, addn);
            mbar_wait(&tfull[as], aph);
            tc_fence_after();
            const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16) + as * g.tstride;
            if (slot >= nhu || (g.dbg & 1)) {   // (never with N >= 48; keeps the barrier protocol total)
                tc_fence_before();
                if (lane == 0) mbar_arrive(&tempty[as]);
            }
            for (int u = (g.dbg & 1) ? nhu : slot; u < nhu; u += EQ) {
                const int col0 = u * 16;
                uint32_t r[2][8];
                uint32_t nx[2][4];
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) tmem_ld_16x256_x2(tbase + ((uint32_t)(hh * 16) << 16) + col0, r[hh]);
                if (PAIR) {
                    const int cn = min(col0 + 16, g.N - 8);   // first group of the next half-unit (the tile's last column pairs with nothing)
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) tmem_ld_16x256_x1(tbase + ((uint32_t)(hh * 16) << 16) + cn, nx[hh]);
                }
                // lane & 15 = column of the half-unit: validity and pixel offset inside the image (overlaps the TMEM load latency)
                const int pos_l = q0 + col0 + (lane & 15);
                const int rr_l = (int)__umulhi((uint32_t)pos_l, g.inv_pitch);
                const bool ok_l = pos_l - rr_l * g.pitch >= 1 && rr_l < g.H && col0 + (lane & 15) < g.adv;
                const uint32_t vmask = __ballot_sync(0xffffffffu, ok_l);
                const int pixoff_l = pos_l - rr_l - 1;
                const uint32_t m2 = vmask >> (2 * t4);
                tmem_ld_wait();
                if (u + EQ >= nhu) {   // this warp's last half-unit: its share of the stage is in registers
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tempty[as]);
                }
                uint32_t mprev[2] = {0u, 0u};
#pragma unroll
                for (int gp = 0; gp < 2; ++gp) {
                    // columns outside the image are SELECTED away, never multiplied by zero: their accumulators come from shared-memory
                    // rows behind the box (whatever bytes lie there -- NaN / inf bit patterns included)
                    const bool vA = (m2 & (1u << (8 * gp))) != 0, vB = (m2 & (2u << (8 * gp))) != 0;
                    uint32_t m[4];
                    if (PAIR) {
#pragma unroll
                        for (int hh = 0; hh < 2; ++hh) {
                            const float a0 = __uint_as_float(r[hh][4 * gp]), a1 = __uint_as_float(r[hh][4 * gp + 1]);
                            const float b0 = __uint_as_float(r[hh][4 * gp + 2]), b1 = __uint_as_float(r[hh][4 * gp + 3]);
                            const float b0n = __uint_as_float(gp < 1 ? r[hh][4 * gp + 6] : nx[hh][2]);   // b0 of the next column group
                            const float send = t4 == 0 ? b0n : b0;
                            const float got = __shfl_sync(0xffffffffu, send, (lane & ~3) | ((lane + 1) & 3));
                            const float2 o = fadd2(fadd2(make_float2(a0, a1), make_float2(b1, got)), make_float2(addv[hh], addv[hh]));
                            const float2 x = make_float2(vA ? o.x : 0.f, vB ? o.y : 0.f);
                            s1[hh] = fadd2(s1[hh], x);
                            s2[hh] = ffma2(x, x, s2[hh]);
                            m[hh] = pack_bf16x2(o.x, o.y);
                        }
                        // both column groups in one stmatrix: matrices {gp 0: hh0, hh1, gp 1: hh0, hh1}
                        if (gp & 1) {
                            const uint32_t a = stage + (uint32_t)((8 * (lane >> 4) + (lane & 7)) * SP + ((lane >> 3) & 1) * 16);
                            stmatrix_x4_trans(a, mprev[0], mprev[1], m[0], m[1]);
                        } else {
                            mprev[0] = m[0]; mprev[1] = m[1];
                        }
                    } else {
#pragma unroll
                        for (int v = 0; v < 4; ++v) {
                            const int hh = v >> 1, k = v & 1;
                            const float2 o = fadd2(make_float2(__uint_as_float(r[hh][4 * gp + 2 * k]), __uint_as_float(r[hh][4 * gp + 2 * k + 1])),
                                                   make_float2(addv[v], addv[v]));
                            const float2 x = make_float2(vA ? o.x : 0.f, vB ? o.y : 0.f);
                            s1[v] = fadd2(s1[v], x);
                            s2[v] = ffma2(x, x, s2[v]);
                            m[v] = pack_bf16x2(o.x, o.y);
                        }
                        const uint32_t a = stage + (uint32_t)((8 * gp + (lane & 7)) * SP + (lane >> 3) * 16);
                        stmatrix_x4_trans(a, m[0], m[1], m[2], m[3]);
                    }
                }
                __syncwarp();
                // copy-out: 16 pixel rows x (CV 16-byte parts); CV consecutive lanes write one pixel's contiguous channels
#pragma unroll
                for (int k2 = 0; k2 < CV / 2; ++k2) {
                    const int px = k2 * (32 / CV) + lane / CV, part = lane % CV;
                    const int pixoff = __shfl_sync(0xffffffffu, pixoff_l, px);
                    if (((vmask >> px) & 1u) && !(g.dbg & 4)) {
                        uint4 val;
                        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(val.x), "=r"(val.y), "=r"(val.z), "=r"(val.w) : "r"(stage + (uint32_t)(px * SP + part * 16)));
                        *reinterpret_cast<uint4*>(obase + (uint32_t)(pixoff * p.Cout + part * 8)) = val;
                    }
                }
                __syncwarp();
            }
            if (p.stats) {
                // transposing butterfly: every value is summed over all 32 lanes (4 column lanes x 8 channels of an 8-block)
                float t[8];
#pragma unroll
                for (int v = 0; v < CV; ++v) { t[2 * v] = s1[v].x + s1[v].y; t[2 * v + 1] = s2[v].x + s2[v].y; }
                int idx;
                if (PAIR) {
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const float send = (lane & 16) ? t[i] : t[i + 2], keep = (lane & 16) ? t[i + 2] : t[i];
                        t[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
                    }
                    {
                        const float send = (lane & 8) ? t[0] : t[1], keep = (lane & 8) ? t[1] : t[0];
                        t[0] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
                    }
                    t[0] += __shfl_xor_sync(0xffffffffu, t[0], 4);
                    t[0] += __shfl_xor_sync(0xffffffffu, t[0], 2);
                    t[0] += __shfl_xor_sync(0xffffffffu, t[0], 1);
                    idx = ((lane >> 4) & 1) * 2 + ((lane >> 3) & 1);
                } else {
                    warp_reduce8(t, lane);
                    idx = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
                }
                if ((lane & (PAIR ? 7 : 3)) == 0) {
                    float* dst = p.stats + ((size_t)n * (p.Cout >> p.slab_shift) + ((cbase + 8 * (idx >> 1)) >> p.slab_shift)) * 2;
                    atomicAdd(dst + (idx & 1), t[0]);
                }
            }
            if (++as == (uint32_t)g.tstages) { as = 0; aph ^= 1; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 512);
}

}  // namespace rfv
