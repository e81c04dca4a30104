// Convolution weight gradient on tcgen05:  dW[co][tap][ci] = sum over (n, pixel) of dY[n,pix,co] * A[n,pix+tap,ci]
// (the wgrad of nn.Conv2d at models/unet.py:38,41,51,76,77,185,217 -- what loss.backward() at
//  models/rectified_flow.py:235 computes for every conv weight).
//
// GEMM view: the reduction (K) dimension is the PIXEL index, so both operands are read "MN-major": a TMA box of an
// NHWC tensor {64 channels, pitch, rows} lands in shared memory as 128-byte rows (one pixel's 64 channels) in 8-row
// 128B-swizzle atoms -- exactly the canonical MN-major SWIZZLE_128B layout ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte
// units: SBO = 1024 B between 8-pixel groups, LBO = distance between 64-channel blocks of the M (or N) dimension.
//   * B operand (N = 64 output channels): the dY tile, R whole image rows in the flat padded space of conv_wa.cuh
//     (row pitch W+1, position 0 of each row is a zero column supplied by the TMA unit's out-of-bounds fill).
//   * A operand (M = 128): TWO taps of the input halo box {64 ch, W+1, R+2 rows}.  In the padded space tap (dy,dx) is
//     the constant row shift (1+dy)*(W+1)+dx of the box, so "the same 64 channels one tap further" is simply another
//     64-row block LBO = (shift1-shift0)*128 bytes away: one MMA accumulates two taps, nine taps take five MMAs per
//     16-pixel K step (the odd tap is paired with its neighbour again and that half is discarded).
// Accumulators (5 units x 64 fp32 columns of TMEM) persist across all pixel tiles a CTA walks for one
// (input-chunk, output-chunk) pair; they are flushed with coalesced fp32 reductions (red.global.add) into the packed
// [C_out][K] gradient, so the split over pixels (split-K) costs one flush per CTA per pair.
// Stride-2 convs read four parity views of the input (four tensor maps) whose taps are again constant shifts;
// 1x1 convs are the single-atom case.
//
//   warp 0  TMA producer     warp 1  MMA issuer     warp 2  TMEM allocator     warps 4-7  flush
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace rfv {

constexpr int WG_MAX_UNITS = 5;
constexpr int WG_THREADS = 256;

struct WgradGeom {
    int pitch, R, tiles_per_img, ksteps;
    int cchA, cchB, nvar;          // 64-channel chunks of A; N-chunks (64*ncob channels) of dY; variants (tap groups / parity views)
    int ncob;                      // 64-channel dY boxes per N-chunk: N = 64 (1) or 128 (2) per MMA
    int b_sub_bytes;               // bytes from one dY box to the next inside a stage
    int a_stage_bytes, b_stage_bytes, stages;
    int a_box_bytes, b_box_bytes;
    int ldw;                       // floats per gradient row
    int num_tiles;                 // images * tiles_per_img (launch time)
    int nunits[4];
    int u_shift[4][WG_MAX_UNITS];  // row (128 B) of the unit's first atom inside the A box, relative to dY position 0
    int u_lbo[4][WG_MAX_UNITS];    // bytes from the first to the second atom
    int u_koff0[4][WG_MAX_UNITS];  // gradient column offset of the first / second atom (before + chunk*64 + ci); -1: discard
    int u_koff1[4][WG_MAX_UNITS];
    int y_per_var;                 // 1: variant v reads its own dY map (sub-pixel phases of an upsample conv)
    int u_kx0[4][WG_MAX_UNITS][3]; // further gradient columns the first / second atom is ALSO added to (-1: none): a
    int u_kx1[4][WG_MAX_UNITS][3]; // pre-summed sub-pixel tap is the sum of up to four 3x3 taps, so its gradient goes to each
};

// Matrix descriptor, MN-major operand, 128-byte swizzle (see header comment).
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_umma_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
                  const __grid_constant__ CUtensorMap mapA2, const __grid_constant__ CUtensorMap mapA3,
                  const __grid_constant__ CUtensorMap mapY, const __grid_constant__ CUtensorMap mapY1,
                  const __grid_constant__ CUtensorMap mapY2, const __grid_constant__ CUtensorMap mapY3, float* __restrict__ dW,
                  const WgradGeom g) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int stage_bytes = g.a_stage_bytes + g.b_stage_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)g.stages * stage_bytes);
    uint64_t* full = bars;                 // [stages]
    uint64_t* empty = full + g.stages;     // [stages]
    uint64_t* acc_full = empty + g.stages;
    uint64_t* acc_empty = acc_full + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // guards / tails of every stage must read as zero (dY) or at least finite (A): clear everything once
    for (int i = threadIdx.x; i < g.stages * stage_bytes / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&mapA0);
        tma_prefetch_desc(&mapY);
        for (int s = 0; s < g.stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(acc_full, 1);
        mbar_init(acc_empty, 128);
        mbar_fence_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, 512);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // this CTA's contiguous slice of the (pair-major) work list: w = pair * num_tiles + tile
    const long long total = (long long)g.nvar * g.cchA * g.cchB * g.num_tiles;
    const long long w0 = total * blockIdx.x / gridDim.x, w1 = total * (blockIdx.x + 1) / gridDim.x;

    if (warp == 0) {
        uint32_t st = 0, ph = 0;
        for (long long w = w0; w < w1; ++w) {
            const int pair = (int)(w / g.num_tiles), tile = (int)(w - (long long)pair * g.num_tiles);
            const int b = pair % g.cchB, va = pair / g.cchB, a = va % g.cchA, var = va / g.cchA;
            const int n = tile / g.tiles_per_img, r0 = (tile - n * g.tiles_per_img) * g.R;
            mbar_wait(&empty[st], ph ^ 1);
            if (elect_one()) {
                uint8_t* sa = smem + (size_t)st * stage_bytes;
                uint8_t* sb = sa + g.a_stage_bytes;
                mbar_arrive_expect_tx(&full[st], g.a_box_bytes + g.ncob * g.b_box_bytes);
                const CUtensorMap* ma = var == 0 ? &mapA0 : (var == 1 ? &mapA1 : (var == 2 ? &mapA2 : &mapA3));
                tma_load_4d(sa + 1024, ma, &full[st], a * 64, -1, r0 - 1, n);
                const CUtensorMap* my = (!g.y_per_var || var == 0) ? &mapY : (var == 1 ? &mapY1 : (var == 2 ? &mapY2 : &mapY3));
                for (int j = 0; j < g.ncob; ++j) tma_load_4d(sb + j * g.b_sub_bytes, my, &full[st], (b * g.ncob + j) * 64, -1, r0, n);
            }
            __syncwarp();
            if (++st == (uint32_t)g.stages) { st = 0; ph ^= 1; }
        }
    } else if (warp == 1) {
        const int N = 64 * g.ncob;
        const uint32_t idesc = umma_idesc_bf16(128, N) | (1u << 15) | (1u << 16);  // both operands MN-major
        uint32_t st = 0, ph = 0, accph = 0;
        int cur_pair = -1;
        bool first = true;
        for (long long w = w0; w < w1; ++w) {
            const int pair = (int)(w / g.num_tiles);
            const int var = pair / (g.cchB * g.cchA);
            if (pair != cur_pair) {
                if (cur_pair >= 0) {  // the flush warps must have drained the previous pair's accumulators
                    mbar_wait(acc_empty, accph);
                    accph ^= 1;
                    tc_fence_after();
                }
                cur_pair = pair;
                first = true;
            }
            mbar_wait(&full[st], ph);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t sa = smem_u32(smem + (size_t)st * stage_bytes) + 1024;
                const uint32_t sb = sa - 1024 + g.a_stage_bytes;
                const int nu = g.nunits[var];
                for (int ks = 0; ks < g.ksteps; ++ks) {
                    const uint64_t bdesc = umma_desc_mn_sw128(sb + ks * 2048, (uint32_t)g.b_sub_bytes);  // LBO: next 64 output channels
                    for (int u = 0; u < nu; ++u) {
                        const uint64_t adesc = umma_desc_mn_sw128(sa + (uint32_t)((g.u_shift[var][u] + ks * 16) * 128), (uint32_t)g.u_lbo[var][u]);
                        umma_bf16(tmem_base + u * N, adesc, bdesc, idesc, (first && ks == 0) ? 0u : 1u);
                    }
                }
                umma_commit(&empty[st]);
                const bool last_of_pair = (w + 1 == w1) || ((int)((w + 1) / g.num_tiles) != pair);
                if (last_of_pair) umma_commit(acc_full);
            }
            __syncwarp();
            first = false;
            if (++st == (uint32_t)g.stages) { st = 0; ph ^= 1; }
        }
    } else if (warp >= 4) {
        const int q = warp & 3;
        const int row = q * 32 + lane;
        uint32_t fph = 0;
        long long w = w0;
        while (w < w1) {
            const int pair = (int)(w / g.num_tiles);
            const long long pair_end = (long long)(pair + 1) * g.num_tiles;
            w = pair_end < w1 ? pair_end : w1;
            const int b = pair % g.cchB, va = pair / g.cchB, a = va % g.cchA, var = va / g.cchA;
            mbar_wait(acc_full, fph);
            fph ^= 1;
            tc_fence_after();
            const int nu = g.nunits[var], N = 64 * g.ncob;
            for (int u = 0; u < nu; ++u) {
                const int koff = row < 64 ? g.u_koff0[var][u] : g.u_koff1[var][u];  // warp-uniform (q < 2 / q >= 2)
#pragma unroll 1
                for (int ch = 0; ch < N / 32; ++ch) {
                    uint32_t acc[32];
                    tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + u * N + ch * 32, acc);
                    tmem_ld_wait();
                    if (koff >= 0) {
                        float* dst = dW + (size_t)(b * N + ch * 32) * g.ldw + koff + a * 64 + (row & 63);
#pragma unroll
                        for (int j = 0; j < 32; ++j) atomicAdd(dst + (size_t)j * g.ldw, __uint_as_float(acc[j]));
                        for (int x = 0; x < 3; ++x) {   // sub-pixel atoms: the same values belong to further 3x3 taps
                            const int kx = row < 64 ? g.u_kx0[var][u][x] : g.u_kx1[var][u][x];
                            if (kx < 0) break;
                            float* d2 = dW + (size_t)(b * N + ch * 32) * g.ldw + kx + a * 64 + (row & 63);
#pragma unroll
                            for (int j = 0; j < 32; ++j) atomicAdd(d2 + (size_t)j * g.ldw, __uint_as_float(acc[j]));
                        }
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(acc_empty);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// ---- host-side geometry ---------------------------------------------------------------------------------
// W, H: dY (conv output) grid.  kind 0: 3x3 stride 1 (taps = 9 shifts of one box); kind 1: 1x1; kind 2: 3x3 stride 2
// (A variants = the four parity views (ph,pw) of the input, dims W x H each).  Cin, Cout multiples of 64.
// koff_base: column of this conv's first weight inside the gradient row (shortcut segments live after the 3x3 part).
inline bool make_wgrad_geom(WgradGeom* g, int W, int H, int Cin, int Cout, int kind, int ldw, int koff_base, int allow_n128 = 1) {
    *g = WgradGeom{};
    g->pitch = W + 1;
    // N = 128 per MMA runs the tensor pipe at 79 % instead of 39 % (measured 81-cycle floor per instruction); it needs the
    // nine taps of a 3x3 conv split into two variants (5 + 4 taps) because 9 x 64 x 128 fp32 accumulators exceed TMEM
    g->ncob = (allow_n128 && Cout % 128 == 0) ? 2 : 1;
    const int avail = 227 * 1024 - 1024 - 256;
    int R = 1;
    while (R * 2 <= H && R * 2 * g->pitch <= 272) R *= 2;
    for (;; R /= 2) {
        g->R = R;
        g->ksteps = (R * g->pitch + 15) / 16;
        g->a_box_bytes = (R + 2) * g->pitch * 128;
        g->b_box_bytes = R * g->pitch * 128;
        g->a_stage_bytes = 1024 + ((g->a_box_bytes + 17 * 128 + 1023) & ~1023);
        g->b_sub_bytes = (g->ksteps * 2048 + 1023) & ~1023;
        g->b_stage_bytes = g->ncob * g->b_sub_bytes;
        g->stages = avail / (g->a_stage_bytes + g->b_stage_bytes);
        if (g->stages >= 2 || R == 1) break;
    }
    if (g->stages > 4) g->stages = 4;
    if (g->stages < 2) return false;
    g->tiles_per_img = (H + g->R - 1) / g->R;
    g->cchA = Cin / 64; g->cchB = Cout / (64 * g->ncob);
    const int max_units = 512 / (64 * g->ncob) > WG_MAX_UNITS ? WG_MAX_UNITS : 512 / (64 * g->ncob);
    g->ldw = ldw;
    // atoms (tap shifts) per variant
    int vshift[4][9], vkoff[4][9], vn[4] = {0, 0, 0, 0};
    int vkx[4][9][3];
    for (int v = 0; v < 4; ++v)
        for (int i = 0; i < 9; ++i) vkx[v][i][0] = vkx[v][i][1] = vkx[v][i][2] = -1;
    if (kind == 0) {
        const int split = g->ncob == 2 ? 5 : 9;   // variant 0: taps [0, split), variant 1: the rest
        g->nvar = g->ncob == 2 ? 2 : 1;
        for (int tap = 0; tap < 9; ++tap) {
            const int v = tap < split ? 0 : 1;
            vshift[v][vn[v]] = (tap / 3) * g->pitch + (tap % 3 - 1);
            vkoff[v][vn[v]] = koff_base + tap * Cin;
            ++vn[v];
        }
    } else if (kind == 1) {
        g->nvar = 1;
        vshift[0][0] = g->pitch; vkoff[0][0] = koff_base; vn[0] = 1;
    } else if (kind == 3) {
        // nearest-x2 upsample + 3x3 conv as four sub-pixel phases (W, H: the LOW resolution = grid of each dY parity view):
        // out[2i+py, 2j+px] = sum_{a,b} W'[py,px][a,b] . in[i+a+py-1, j+b+px-1];  W'[..][a,b] = sum of the 3x3 taps (ky,kx) with
        // ky in KY(py,a), kx in KX(px,b):  (0,0)->{0} (0,1)->{1,2} (1,0)->{0,1} (1,1)->{2}.  dW[ky,kx] += dW'[phase][a,b].
        g->nvar = 4;
        g->y_per_var = 1;
        for (int var = 0; var < 4; ++var) {
            const int py = var >> 1, px = var & 1;
            for (int a = 0; a < 2; ++a)
                for (int b = 0; b < 2; ++b) {
                    const int i = vn[var]++;
                    vshift[var][i] = (a + py) * g->pitch + (b + px - 1);
                    const int ky0 = py == 0 ? (a == 0 ? 0 : 1) : (a == 0 ? 0 : 2), ky1 = py == 0 ? (a == 0 ? 0 : 2) : (a == 0 ? 1 : 2);
                    const int kx0 = px == 0 ? (b == 0 ? 0 : 1) : (b == 0 ? 0 : 2), kx1 = px == 0 ? (b == 0 ? 0 : 2) : (b == 0 ? 1 : 2);
                    int cnt = 0;
                    for (int ky = ky0; ky <= ky1; ++ky)
                        for (int kx = kx0; kx <= kx1; ++kx) {
                            const int ko = koff_base + (ky * 3 + kx) * Cin;
                            if (cnt == 0) vkoff[var][i] = ko; else vkx[var][i][cnt - 1] = ko;
                            ++cnt;
                        }
                }
        }
    } else {
        // parity view (ph,pw): in[2i+ph, 2j+pw]; tap ky reads row 2i+ky-1: ky=1 -> ph 0 shift 0; ky=0 -> ph 1 shift -1; ky=2 -> ph 1 shift 0
        g->nvar = 4;
        for (int var = 0; var < 4; ++var) {
            const int ph = var >> 1, pw = var & 1;
            for (int ky = 0; ky < 3; ++ky)
                for (int kx = 0; kx < 3; ++kx) {
                    if (((ky + 1) & 1) != ph || ((kx + 1) & 1) != pw) continue;
                    const int dy = ky == 0 ? -1 : 0, dx = kx == 0 ? -1 : 0;
                    vshift[var][vn[var]] = (1 + dy) * g->pitch + dx;
                    vkoff[var][vn[var]] = koff_base + (ky * 3 + kx) * Cin;
                    ++vn[var];
                }
        }
    }
    for (int var = 0; var < g->nvar; ++var) {
        const int na = vn[var];
        const int* shift = vshift[var];
        const int* koff = vkoff[var];
        int nu = 0;
        for (int i = 0; i + 1 < na; i += 2) {
            g->u_shift[var][nu] = shift[i]; g->u_lbo[var][nu] = (shift[i + 1] - shift[i]) * 128;
            g->u_koff0[var][nu] = koff[i]; g->u_koff1[var][nu] = koff[i + 1];
            for (int x = 0; x < 3; ++x) { g->u_kx0[var][nu][x] = vkx[var][i][x]; g->u_kx1[var][nu][x] = vkx[var][i + 1][x]; }
            ++nu;
        }
        for (int x = 0; x < 3 && (na & 1); ++x) g->u_kx0[var][nu][x] = g->u_kx1[var][nu][x] = -1;   // odd tails: 3x3 / 1x1 only
        if (na & 1) {
            if (na == 1) { g->u_shift[var][nu] = shift[0]; g->u_lbo[var][nu] = 1024; g->u_koff0[var][nu] = koff[0]; g->u_koff1[var][nu] = -1; }
            else { g->u_shift[var][nu] = shift[na - 2]; g->u_lbo[var][nu] = (shift[na - 1] - shift[na - 2]) * 128; g->u_koff0[var][nu] = -1; g->u_koff1[var][nu] = koff[na - 1]; }
            ++nu;
        }
        if (nu > max_units) return false;
        g->nunits[var] = nu;
    }
    return true;
}
inline size_t wgrad_smem_bytes(const WgradGeom& g) { return 1024 + (size_t)g.stages * (g.a_stage_bytes + g.b_stage_bytes) + 256; }

}  // namespace rfv
