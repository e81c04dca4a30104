// HBM-bound / small kernels around the convolutions: time embedding, input conv (+NCHW fp32 -> NHWC bf16),
// GroupNorm-apply(+SiLU)(+virtual concat), output conv fused with the Euler update (+NHWC -> NCHW fp32),
// weight packing, flow-matching interpolation.
#pragma once
#include "common.cuh"
#include "train_kernels.cuh"

namespace rfv {

// ---------------------------------------------------------------------------------------------------------
// Time embedding: temb_act[b, :] = SiLU(W2 . SiLU(W1 . sincos(t_b) + b1) + b2)        (models/unet.py:20-27,157-162)
// The trailing SiLU is the first op of every ResidualBlock.time_mlp (models/unet.py:43-46), shared by all blocks.
// grid (rows, TEMB_SLICES): one batch row per blockIdx.x (a single row when t is uniform over the batch, which is the
// sampling case); the second layer's outputs are split over blockIdx.y, every slice recomputing the small first layer, so
// that the uniform-t case is not one lonely block on a 148-SM GPU.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) temb_kernel(const float* __restrict__ t, int t_stride_is_zero, float t_scalar,
                                                   const float* __restrict__ w1, const float* __restrict__ b1,
                                                   const float* __restrict__ w2, const float* __restrict__ b2,
                                                   float* __restrict__ out, int mc, int td, float* __restrict__ save_emb,
                                                   float* __restrict__ save_z1, float* __restrict__ save_h1,
                                                   float* __restrict__ save_z2) {
    pdl_wait();   // programmatic dependent launch (common.cuh): no-ops unless launched with the attribute
    pdl_launch();
    // save_*: training only -- the sinusoidal embedding and the two pre-activations, read by the time-MLP backward
    extern __shared__ float sm[];
    float* emb = sm;        // [mc]
    float* h1 = sm + mc;    // [td]
    const int b = blockIdx.x;
    const float tv = t ? t[b] : t_scalar;
    const int half = mc / 2;
    const float step = logf(10000.0f) / (float)(half - 1);
    for (int j = threadIdx.x; j < half; j += blockDim.x) {
        const float f = expf(-step * (float)j);
        const float a = tv * f;
        emb[j] = sinf(a);
        emb[j + half] = cosf(a);
    }
    __syncthreads();
    if (save_emb && blockIdx.y == 0)
        for (int j = threadIdx.x; j < mc; j += blockDim.x) save_emb[(size_t)b * mc + j] = emb[j];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int o = warp; o < td; o += nw) {
        float s = 0.f;
        for (int j = lane; j < mc; j += 32) s += w1[(size_t)o * mc + j] * emb[j];
        s = warp_sum(s);
        if (lane == 0) {
            h1[o] = silu_f(s + b1[o]);
            if (save_z1 && blockIdx.y == 0) { save_z1[(size_t)b * td + o] = s + b1[o]; save_h1[(size_t)b * td + o] = h1[o]; }
        }
    }
    __syncthreads();
    const int per = (td + gridDim.y - 1) / gridDim.y, o_lo = blockIdx.y * per, o_hi = min(td, o_lo + per);
    for (int o = o_lo + warp; o < o_hi; o += nw) {
        float s = 0.f;
        for (int j = lane; j < td; j += 32) s += w2[(size_t)o * td + j] * h1[j];
        s = warp_sum(s);
        if (lane == 0) {
            out[(size_t)b * td + o] = silu_f(s + b2[o]);
            if (save_z2) save_z2[(size_t)b * td + o] = s + b2[o];
        }
    }
}

// All ResidualBlock time projections at once: proj[b, c] = Wcat[c, :] . temb_act[b, :] + bcat[c]   (models/unet.py:59-60)
// grid (ceil(sumC/64), ceil(rows/8)): each block keeps 8 batch rows in smem so a weight row is read once per 8 rows.
__global__ void __launch_bounds__(256) temb_proj_kernel(const float* __restrict__ act, const float* __restrict__ wcat,
                                                        const float* __restrict__ bcat, float* __restrict__ proj,
                                                        int rows, int td, int sumC) {
    pdl_wait();   // programmatic dependent launch (common.cuh): no-ops unless launched with the attribute
    pdl_launch();
    extern __shared__ float sm[];  // [8][td]
    const int r0 = blockIdx.y * 8;
    const int nr = min(8, rows - r0);
    for (int i = threadIdx.x; i < nr * td; i += blockDim.x) sm[i] = act[(size_t)r0 * td + i];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int oc = warp; oc < 64; oc += 8) {
        const int c = blockIdx.x * 64 + oc;
        if (c >= sumC) break;
        float s[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) s[i] = 0.f;
        for (int j = lane; j < td; j += 32) {
            const float wv = wcat[(size_t)c * td + j];
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if (i < nr) s[i] += wv * sm[i * td + j];
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float v = warp_sum(s[i]);
            if (lane == 0 && i < nr) proj[(size_t)(r0 + i) * sumC + c] = v + bcat[c];
        }
    }
}

// t_i = i * dt in double, rounded once to fp32: what `torch.full((B,), i * dt)` holds in models/base_flow.py:163-166
__global__ void fill_step_times_kernel(float* __restrict__ t, int n, double dt) {
    pdl_wait();   // programmatic dependent launch (common.cuh): no-ops unless launched with the attribute
    pdl_launch();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) t[i] = (float)((double)i * dt);
}

// ---------------------------------------------------------------------------------------------------------
// Input conv 3x3, C_in small (<= 4): fp32 NCHW in -> bf16 NHWC out (+ GroupNorm slab statistics).  models/unet.py:165,234
// Optionally computes the flow-matching interpolation on the fly: x = (1-t) x0 + t x1   (models/base_flow.py:84).
// One thread = one output pixel x all C_out channels: its 9*C_in inputs live in registers, weights are broadcast from
// shared memory ([k][co] fp32, pre-transposed at upload time), output rows are written as contiguous 16-byte vectors.
// ---------------------------------------------------------------------------------------------------------
template <int CIN>
__global__ void __launch_bounds__(256) input_conv_kernel(const float* __restrict__ x, const float* __restrict__ x1,
                                                         const float* __restrict__ tvec, const float* __restrict__ wt,
                                                         const float* __restrict__ bias, bf16* __restrict__ out,
                                                         float* __restrict__ stats, int H, int W, int Cout, int slab_shift) {
    pdl_wait();   // programmatic dependent launch (common.cuh): no-ops unless launched with the attribute
    pdl_launch();
    // One thread = TWO horizontally adjacent output pixels x all C_out channels (16 at a time): the 3x4 input patch per
    // channel lives in registers and every broadcast weight vector read from shared memory feeds 8 FMAs instead of 4 (the
    // one-pixel version was bound by the shared-memory pipe: ncu l1tex 89 %, short-scoreboard stalls).
    constexpr int K = CIN * 9;
    extern __shared__ float sm[];
    float* wsm = sm;               // [K][Cout]
    float* bsm = sm + K * Cout;    // [Cout]
    float* st = bsm + Cout;        // [Cout/8][2]
    const int HW = H * W;
    const int n = blockIdx.y, pix = (blockIdx.x * 256 + threadIdx.x) * 2;   // W is even: both pixels are in the same row
    for (int i = threadIdx.x; i < K * Cout; i += 256) wsm[i] = wt[i];
    for (int i = threadIdx.x; i < Cout; i += 256) bsm[i] = bias[i];
    for (int i = threadIdx.x; i < (Cout / 8) * 2; i += 256) st[i] = 0.f;
    const bool active = pix < HW;   // the last block of a small image may be partly empty
    const int h = pix / W, w = pix - h * W;
    const float tb = (x1 != nullptr) ? tvec[n] : 0.f;
    float in[CIN][3][4];
#pragma unroll
    for (int ci = 0; ci < CIN; ++ci)
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 4; ++kx) {
                const int hh = h + ky - 1, ww = w + kx - 1;
                float v = 0.f;
                if (active && hh >= 0 && hh < H && ww >= 0 && ww < W) {
                    const size_t o = ((size_t)n * CIN + ci) * HW + (size_t)hh * W + ww;
                    v = x[o];
                    if (x1 != nullptr) v = (1.0f - tb) * v + tb * x1[o];
                }
                in[ci][ky][kx] = v;
            }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    bf16* orow = out + ((size_t)n * HW + pix) * Cout;
    for (int c0 = 0; c0 < Cout; c0 += 16) {
        float acc[2][16];
#pragma unroll
        for (int j = 0; j < 16; ++j) { acc[0][j] = bsm[c0 + j]; acc[1][j] = acc[0][j]; }
#pragma unroll
        for (int ci = 0; ci < CIN; ++ci)
#pragma unroll
            for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const float a0 = in[ci][ky][kx], a1 = in[ci][ky][kx + 1];
                    const float4* wr = reinterpret_cast<const float4*>(wsm + (ci * 9 + ky * 3 + kx) * Cout + c0);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float4 w4 = wr[j];
                        acc[0][4 * j] = fmaf(a0, w4.x, acc[0][4 * j]);         acc[1][4 * j] = fmaf(a1, w4.x, acc[1][4 * j]);
                        acc[0][4 * j + 1] = fmaf(a0, w4.y, acc[0][4 * j + 1]); acc[1][4 * j + 1] = fmaf(a1, w4.y, acc[1][4 * j + 1]);
                        acc[0][4 * j + 2] = fmaf(a0, w4.z, acc[0][4 * j + 2]); acc[1][4 * j + 2] = fmaf(a1, w4.z, acc[1][4 * j + 2]);
                        acc[0][4 * j + 3] = fmaf(a0, w4.w, acc[0][4 * j + 3]); acc[1][4 * j + 3] = fmaf(a1, w4.w, acc[1][4 * j + 3]);
                    }
                }
#pragma unroll
        for (int px = 0; px < 2 && active; ++px) {
            *reinterpret_cast<uint4*>(orow + (size_t)px * Cout + c0) = pack8(acc[px]);
            *reinterpret_cast<uint4*>(orow + (size_t)px * Cout + c0 + 8) = pack8(acc[px] + 8);
        }
        if (stats) {
            // 4 partial sums per lane (2 slabs x {sum, sum of squares}) over both pixels; every lane ends up with the warp
            // total of value ((lane >> 4) & 1) * 2 + ((lane >> 3) & 1)
            float t4[4];
#pragma unroll
            for (int sl = 0; sl < 2; ++sl) {
                float s = 0.f, ss = 0.f;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float v0 = active ? acc[0][sl * 8 + j] : 0.f, v1 = active ? acc[1][sl * 8 + j] : 0.f;
                    s += v0 + v1;
                    ss += v0 * v0 + v1 * v1;
                }
                t4[sl * 2] = s;
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
            if ((lane & 7) == 0) {
                const int idx = ((lane >> 4) & 1) * 2 + ((lane >> 3) & 1);
                atomicAdd(&st[((c0 >> 3) + (idx >> 1)) * 2 + (idx & 1)], t4[0]);
            }
        }
    }
    if (stats) {
        __syncthreads();
        for (int i = threadIdx.x; i < Cout / 8; i += 256) {
            float* dst = stats + ((size_t)n * (Cout >> slab_shift) + ((i * 8) >> slab_shift)) * 2;
            atomicAdd(dst, st[i * 2]);
            atomicAdd(dst + 1, st[i * 2 + 1]);
        }
    }
}
// ---------------------------------------------------------------------------------------------------------
// Input conv on the tensor cores (mma.sync m16n8k16, bf16 operands, fp32 accumulation): the FMA version above tops out at
// ~45 % of the fp32 FMA peak, 5x above the kernel's HBM time.  GEMM per 8x32-pixel tile: M = pixels, K = 9*C_in padded to a
// multiple of 16, N = C_out in chunks of 8*NTC channels.  The fp32 halo of the tile is staged in shared memory; im2col is a
// per-thread table of K offsets into it (each A-fragment register is two shared-memory reads and one cvt.bf16x2); the weight
// fragments of the current channel chunk live in registers.  Same outputs as input_conv_kernel: NHWC bf16 + GroupNorm slab
// statistics of the fp32 result, optional x_t interpolation.  grid (tile groups, B), IM_TPB tiles per block, warp = tile row.
// ---------------------------------------------------------------------------------------------------------
constexpr int IM_TH = 8, IM_TW = 32, IM_XP = 36, IM_TPB = 4;
template <int CIN, int NTC>
__global__ void __launch_bounds__(256, 2) input_conv_mma_kernel(const float* __restrict__ x, const float* __restrict__ x1,
                                                                const float* __restrict__ tvec, const float* __restrict__ wt,
                                                                const float* __restrict__ bias, bf16* __restrict__ out,
                                                                float* __restrict__ stats, int H, int W, int Cout, int slab_shift) {
    pdl_wait();   // programmatic dependent launch (common.cuh): no-ops unless launched with the attribute
    pdl_launch();
    constexpr int K = CIN * 9, KS = (K + 15) / 16;
    extern __shared__ float ism[];
    float* xs = ism;                                     // [CIN][IM_TH + 2][IM_XP]
    float* st = ism + CIN * (IM_TH + 2) * IM_XP;         // [Cout >> slab_shift][2]
    // per-warp output staging [32 pixels][NTC*8 channels] bf16, row pitch +16 B: fragment-order writes and the 16-byte row
    // reads that follow are both bank-conflict free; rows then leave as full 16-byte vectors (4-byte fragment stores straight
    // to global memory halve the store efficiency and cost as much as the FMA version's arithmetic)
    constexpr int OP = NTC * 16 + 16;
    uint8_t* ostage = reinterpret_cast<uint8_t*>(st + (((Cout >> slab_shift) * 2 + 3) & ~3)) + (size_t)(threadIdx.x >> 5) * 32 * OP;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, tq = lane & 3;
    const int n = blockIdx.y;
    const int tw = (W + IM_TW - 1) / IM_TW, th = (H + IM_TH - 1) / IM_TH, tiles_img = tw * th;
    const int t_begin = blockIdx.x * IM_TPB, t_end = min(tiles_img, t_begin + IM_TPB);
    const int nslab = Cout >> slab_shift;
    for (int i = tid; i < nslab * 2; i += 256) st[i] = 0.f;
    // im2col table: k = s*16 + 2*tq + (j & 1) + 8*(j >> 1)  ->  offset of (ci, ky, kx) in the staged halo; -1: zero padding of K
    int aoff[KS][4];
#pragma unroll
    for (int s = 0; s < KS; ++s)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int k = s * 16 + 2 * tq + (j & 1) + 8 * (j >> 1);
            const int ci = k / 9, r9 = k - ci * 9, ky = r9 / 3, kx = r9 - ky * 3;
            aoff[s][j] = k < K ? (ci * (IM_TH + 2) + ky) * IM_XP + kx : -1;
        }
    const float tb = (x1 != nullptr) ? tvec[n] : 0.f;
    const size_t HW = (size_t)H * W;
    for (int c0 = 0; c0 < Cout; c0 += NTC * 8) {
        // B fragments of this channel chunk: b0 = (W[k0+2tq][n], W[k0+2tq+1][n]), b1 = the same 8 rows further down; n = nt*8 + g
        uint32_t bfr[KS][NTC][2];
#pragma unroll
        for (int s = 0; s < KS; ++s)
#pragma unroll
            for (int nt = 0; nt < NTC; ++nt)
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    const int k = s * 16 + 2 * tq + 8 * hh, col = c0 + nt * 8 + g;
                    const float w0 = k < K ? wt[(size_t)k * Cout + col] : 0.f;
                    const float w1 = k + 1 < K ? wt[(size_t)(k + 1) * Cout + col] : 0.f;
                    bfr[s][nt][hh] = pack_bf16x2(w0, w1);
                }
        float sacc[NTC][2];
#pragma unroll
        for (int nt = 0; nt < NTC; ++nt) { sacc[nt][0] = 0.f; sacc[nt][1] = 0.f; }
        // the halo of tile t+1 is fetched into registers while tile t is multiplied (a block has no other way to hide the
        // global-memory latency: two blocks of eight warps per SM)
        constexpr int NPF = (CIN * (IM_TH + 2) * (IM_TW + 2) + 255) / 256;
        float pv[NPF];
        auto fetch = [&](int t) {
            const int h0 = (t / tw) * IM_TH, w0 = (t - (t / tw) * tw) * IM_TW;
#pragma unroll
            for (int u = 0; u < NPF; ++u) {
                const int i = tid + u * 256;
                const int ci = i / ((IM_TH + 2) * (IM_TW + 2)), r = i - ci * ((IM_TH + 2) * (IM_TW + 2));
                const int yy = r / (IM_TW + 2), xx = r - yy * (IM_TW + 2);
                const int hh = h0 + yy - 1, ww = w0 + xx - 1;
                float v = 0.f;
                if (ci < CIN && hh >= 0 && hh < H && ww >= 0 && ww < W) {
                    const size_t o = ((size_t)n * CIN + ci) * HW + (size_t)hh * W + ww;
                    v = x[o];
                    if (x1 != nullptr) v = (1.0f - tb) * v + tb * x1[o];
                }
                pv[u] = v;
            }
        };
        if (t_begin < t_end) fetch(t_begin);
        for (int t = t_begin; t < t_end; ++t) {
            const int h0 = (t / tw) * IM_TH, w0 = (t - (t / tw) * tw) * IM_TW;
            __syncthreads();   // previous tile's fragments are built; st[] zeroed
#pragma unroll
            for (int u = 0; u < NPF; ++u) {
                const int i = tid + u * 256;
                const int ci = i / ((IM_TH + 2) * (IM_TW + 2)), r = i - ci * ((IM_TH + 2) * (IM_TW + 2));
                const int yy = r / (IM_TW + 2), xx = r - yy * (IM_TW + 2);
                if (ci < CIN) xs[(ci * (IM_TH + 2) + yy) * IM_XP + xx] = pv[u];
            }
            __syncthreads();
            if (t + 1 < t_end) fetch(t + 1);
            const int h = h0 + warp;
#pragma unroll
            for (int m = 0; m < 2; ++m) {
                const float* xr = xs + warp * IM_XP + m * 16 + g;
                float acc[NTC][4];
#pragma unroll
                for (int nt = 0; nt < NTC; ++nt)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[nt][j] = 0.f;
#pragma unroll
                for (int s = 0; s < KS; ++s) {
                    uint32_t af[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {   // q = (k half) * 2 + (row half): a0 (g, k lo), a1 (g+8, k lo), a2 (g, k hi), a3 (g+8, k hi)
                        const int j0 = (q >> 1) * 2, ro = (q & 1) * 8;
                        const float v0 = aoff[s][j0] >= 0 ? xr[aoff[s][j0] + ro] : 0.f;
                        const float v1 = aoff[s][j0 + 1] >= 0 ? xr[aoff[s][j0 + 1] + ro] : 0.f;
                        af[q] = pack_bf16x2(v0, v1);
                    }
#pragma unroll
                    for (int nt = 0; nt < NTC; ++nt) mma_bf16_16816(acc[nt], af, bfr[s][nt][0], bfr[s][nt][1]);
                }
                const int wa = w0 + m * 16 + g, wb = wa + 8;
                const bool va = h < H && wa < W, vb = h < H && wb < W;
                uint8_t* sa = ostage + (m * 16 + g) * OP + tq * 4;
                uint8_t* sb = sa + 8 * OP;
#pragma unroll
                for (int nt = 0; nt < NTC; ++nt) {
                    const float b0 = bias[c0 + nt * 8 + 2 * tq], b1 = bias[c0 + nt * 8 + 2 * tq + 1];
                    const float v0 = acc[nt][0] + b0, v1 = acc[nt][1] + b1, v2 = acc[nt][2] + b0, v3 = acc[nt][3] + b1;
                    *reinterpret_cast<uint32_t*>(sa + nt * 16) = pack_bf16x2(v0, v1);
                    *reinterpret_cast<uint32_t*>(sb + nt * 16) = pack_bf16x2(v2, v3);
                    if (va) {
                        sacc[nt][0] += v0 + v1;
                        sacc[nt][1] += v0 * v0 + v1 * v1;
                    }
                    if (vb) {
                        sacc[nt][0] += v2 + v3;
                        sacc[nt][1] += v2 * v2 + v3 * v3;
                    }
                }
            }
            __syncwarp();
            if (h < H) {   // this warp's 32-pixel row: NTC 16-byte vectors per pixel
#pragma unroll
                for (int i = 0; i < NTC; ++i) {
                    const int v = i * 32 + lane, px = v / NTC, ch = v - px * NTC;
                    if (w0 + px < W)
                        *reinterpret_cast<uint4*>(out + (((size_t)n * H + h) * W + w0 + px) * Cout + c0 + ch * 8) =
                            *reinterpret_cast<const uint4*>(ostage + px * OP + ch * 16);
                }
            }
            __syncwarp();
        }
        if (stats) {
#pragma unroll
            for (int nt = 0; nt < NTC; ++nt)
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    float v = sacc[nt][q];
                    v += __shfl_xor_sync(0xffffffffu, v, 4);
                    v += __shfl_xor_sync(0xffffffffu, v, 8);
                    v += __shfl_xor_sync(0xffffffffu, v, 16);
                    if (g == 0) atomicAdd(&st[((c0 + nt * 8 + 2 * tq) >> slab_shift) * 2 + q], v);   // channel pair -> its slab
                }
        }
    }
    if (stats) {
        __syncthreads();
        for (int i = tid; i < nslab * 2; i += 256) atomicAdd(stats + (size_t)n * nslab * 2 + i, st[i]);
    }
}
// OIHW fp32 -> [ci*9+tap][co] fp32 (the layout input_conv_kernel stages into shared memory)
__global__ void transpose_input_weight_kernel(const float* __restrict__ src, float* __restrict__ dst, int O, int K) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < O * K) { const int o = i / K, k = i - o * K; dst[k * O + o] = src[i]; }
}

// ---------------------------------------------------------------------------------------------------------
// GroupNorm(8) apply (+SiLU) over a virtual channel-concat of up to two NHWC bf16 tensors.  models/unet.py:56,62,82,224
// Statistics come from the producers' epilogues as per-(n, slab) sums; groups are unions of whole slabs.
// grid (pixel chunks, B); each block first derives per-channel scale/shift for its image, then streams pixels
// with 16-byte loads/stores.  Pure HBM traffic: reads 2 B/elem, writes 2 B/elem.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gn_apply_kernel(const bf16* __restrict__ xa, const bf16* __restrict__ xb,
                                                       const float* __restrict__ stats_a, const float* __restrict__ stats_b,
                                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                                       bf16* __restrict__ out, int Ca, int Cb, int HW, int slab_shift,
                                                       int apply_silu, int pix_per_block, float eps, uint32_t drop_thresh,
                                                       uint32_t drop_seed, float drop_scale) {
    pdl_wait();   // programmatic dependent launch (common.cuh): no-ops unless launched with the attribute
    pdl_launch();
    // drop_thresh != 0 (training only): nn.Dropout after the SiLU (models/unet.py:62), mask from train_kernels.cuh
    // blockDim.x is a multiple of C/8, so every thread owns ONE 8-channel vector position for the whole block and
    // keeps its scale/shift in registers; the streaming loop is then load -> 8 FMAs (+SiLU) -> store.
    extern __shared__ float sm[];
    const int C = Ca + Cb;
    float* scale = sm;       // [C]
    float* shift = sm + C;   // [C]
    __shared__ float gmean[8], grstd[8];
    const int n = blockIdx.y;
    const int cpg = C / 8;
    if (threadIdx.x < 8) {
        const int g = threadIdx.x;
        const int slab = 1 << slab_shift;
        float s = 0.f, ss = 0.f;
        for (int c = g * cpg; c < (g + 1) * cpg; c += slab) {
            const float* src = (c < Ca) ? stats_a + ((size_t)n * (Ca >> slab_shift) + (c >> slab_shift)) * 2
                                        : stats_b + ((size_t)n * (Cb >> slab_shift) + ((c - Ca) >> slab_shift)) * 2;
            s += src[0];
            ss += src[1];
        }
        const float cnt = (float)cpg * (float)HW;
        const float mean = s / cnt;
        const float var = fmaxf(ss / cnt - mean * mean, 0.f);
        gmean[g] = mean;
        grstd[g] = rsqrtf(var + eps);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const int g = c / cpg;
        const float sc = grstd[g] * gamma[c];
        scale[c] = sc;
        shift[c] = beta[c] - gmean[g] * sc;
    }
    __syncthreads();
    const int vpp = C >> 3;                       // 16-byte vectors per pixel
    const int cv = threadIdx.x % vpp;             // this thread's vector position (fixed)
    const int pl = threadIdx.x / vpp;             // pixel lane
    const int pstride = blockDim.x / vpp;
    float sc[8], sh[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { sc[j] = scale[cv * 8 + j]; sh[j] = shift[cv * 8 + j]; }
    // apply_silu == 2: silu(y) = h * (1 + tanh(h)), h = y / 2 -- with the 1/2 folded into scale / shift that is FMA, MUFU.TANH,
    // FMA per element instead of FMA, FMUL, MUFU.EX2, FADD, MUFU.RCP, FMUL (the kernel is issue- and MUFU-bound, not only
    // HBM-bound).  tanh.approx is good to ~2.5e-4 absolute, i.e. |h| * 2.5e-4 on the result: below bf16 resolution of O(1) values.
    if (apply_silu == 2) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { sc[j] *= 0.5f; sh[j] *= 0.5f; }
    }
    const bool from_a = cv * 8 < Ca;
    const bf16* src = from_a ? xa + cv * 8 : xb + (cv * 8 - Ca);
    const int cs = from_a ? Ca : Cb;
    const int p0 = blockIdx.x * pix_per_block;
    const int np = min(pix_per_block, HW - p0);
    const size_t base = (size_t)n * HW + p0;
    constexpr int U = 4;
    for (int pp = pl; pp < np; pp += U * pstride) {
        uint4 q[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int px = pp + u * pstride;
            if (px < np) q[u] = *reinterpret_cast<const uint4*>(src + (base + px) * cs);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int px = pp + u * pstride;
            if (px < np) {
                float f[8];
                unpack8(q[u], f);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float y = fmaf(f[j], sc[j], sh[j]);
                    if (apply_silu == 2) {
                        float t;
                        asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(y));
                        f[j] = fmaf(y, t, y);
                    } else {
                        f[j] = apply_silu ? silu_f(y) : y;
                    }
                }
                if (drop_thresh) {
                    float mk[8];
                    dropout_mask8(drop_seed, (uint32_t)((base + px) * C + cv * 8), drop_thresh, drop_scale, mk);
#pragma unroll
                    for (int j = 0; j < 8; ++j) f[j] *= mk[j];
                }
                *reinterpret_cast<uint4*>(out + (base + px) * C + cv * 8) = pack8(f);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// Output conv 3x3, C -> C_out (<= 4), fused with what follows it at the call site:
//   mode 0  v = conv(a)                                      (UNet.forward result, models/unet.py:275)
//   mode 1  x += dt * conv(a)   [+ snapshot into traj]       (Euler update, models/base_flow.py:170)
//   mode 2  nothing is written (loss only)
//   mode 3  xv = (conv(a) - (x1 - x0)) * dt : the loss gradient w.r.t. the prediction (dt carries 2 / numel)
// and optionally accumulates sum((conv(a) - target)^2), target = x1 - x0, for the loss / straightness metrics
// (models/rectified_flow.py:118,231).  a: NHWC bf16 (already GroupNorm+SiLU'd); x, v, x0, x1: NCHW fp32.
// Tensor-core formulation: per 8x32-pixel tile the halo of `a` is staged in shared memory (cp.async, double
// buffered across the tiles a persistent block walks); im2col is free -- each ldmatrix row address simply points at
// the tap-shifted pixel of the halo tile.  GEMM per warp: M = 32 pixels (one tile row), N = 8 (C_out padded),
// K = 9*C, mma.sync m16n8k16 with bf16 weights [tap][n][C].
// ---------------------------------------------------------------------------------------------------------
constexpr int OC_TH = 8, OC_TW = 32;
__global__ void __launch_bounds__(256) output_conv_kernel(const bf16* __restrict__ a, const bf16* __restrict__ wpk,
                                                          const float* __restrict__ bias, float* __restrict__ xv,
                                                          float* __restrict__ traj, const float* __restrict__ x0,
                                                          const float* __restrict__ x1, float* __restrict__ mse_acc,
                                                          int C, int H, int W, int Cout, int B, int mode, float dt) {
    pdl_wait();   // programmatic dependent launch (common.cuh): no-ops unless launched with the attribute
    pdl_launch();
    extern __shared__ __align__(16) uint8_t smraw[];
    const int pitch = C * 2 + 16;                                  // bytes per staged pixel / weight row
    const int tile_bytes = (OC_TH + 2) * (OC_TW + 2) * pitch;
    uint8_t* tiles = smraw;                                        // 2 x [(TH+2)*(TW+2)][pitch]
    uint8_t* wsm = smraw + 2 * tile_bytes;                         // [9][8][pitch]
    __shared__ float red[8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tw = (W + OC_TW - 1) / OC_TW, th = (H + OC_TH - 1) / OC_TH;
    const int ntiles = tw * th * B;
    const int vec = C >> 3;
    for (int i = tid; i < 9 * 8 * vec; i += 256) {  // packed weights [tap][8][C] bf16 -> padded rows
        const int row = i / vec, cv = i - row * vec;
        *reinterpret_cast<uint4*>(wsm + row * pitch + cv * 16) = *reinterpret_cast<const uint4*>(wpk + (size_t)row * C + cv * 8);
    }
    auto stage_tile = [&](int tile, int buf) {
        const int n = tile / (tw * th), r = tile - n * (tw * th);
        const int h0 = (r / tw) * OC_TH, w0 = (r - (r / tw) * tw) * OC_TW;
        uint8_t* dst = tiles + buf * tile_bytes;
        for (int i = tid; i < (OC_TH + 2) * (OC_TW + 2) * vec; i += 256) {
            const int pp = i / vec, cv = i - pp * vec;
            const int hh = h0 + pp / (OC_TW + 2) - 1, ww = w0 + pp % (OC_TW + 2) - 1;
            const bool ok = hh >= 0 && hh < H && ww >= 0 && ww < W;
            const bf16* src = ok ? a + (((size_t)n * H + hh) * W + ww) * C + cv * 8 : a;
            cp_async16(smem_u32(dst + pp * pitch + cv * 16), src, ok);
        }
    };
    float sq = 0.f;
    int buf = 0;
    if ((int)blockIdx.x < ntiles) stage_tile(blockIdx.x, 0);
    cp_async_commit();
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, buf ^= 1) {
        const int nxt = tile + gridDim.x;
        if (nxt < ntiles) stage_tile(nxt, buf ^ 1);
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();
        const uint8_t* tb = tiles + buf * tile_bytes;
        float acc[2][4];
#pragma unroll
        for (int m = 0; m < 2; ++m)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[m][j] = 0.f;
        for (int tap = 0; tap < 9; ++tap) {
            const int ky = tap / 3, kx = tap - ky * 3;
            // A rows: pixel (warp + ky, m*16 + lane%16 + kx) of the halo tile; k half selected by lane/16
            const uint8_t* arow = tb + ((warp + ky) * (OC_TW + 2) + (lane & 15) + kx) * pitch + (lane >> 4) * 16;
            // B rows: weight row n = lane%8 of this tap; lanes 8-15 address the k+8 half (ldmatrix.x2 uses lanes 0-15)
            const uint8_t* brow = wsm + (tap * 8 + (lane & 7)) * pitch + ((lane >> 3) & 1) * 16;
            for (int kc = 0; kc < C / 16; ++kc) {
                uint32_t b0, b1;
                asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(b0), "=r"(b1) : "r"(smem_u32(brow + kc * 32)));
#pragma unroll
                for (int m = 0; m < 2; ++m) {
                    uint32_t af[4];
                    ldmatrix_x4(smem_u32(arow + m * 16 * pitch + kc * 32), af[0], af[1], af[2], af[3]);
                    mma_bf16_16816(acc[m], af, b0, b1);
                }
            }
        }
        // epilogue: thread holds (pixel g / g+8 of each m-tile) x (channels 2t, 2t+1)
        const int n = tile / (tw * th), r = tile - n * (tw * th);
        const int h = (r / tw) * OC_TH + warp, w0 = (r - (r / tw) * tw) * OC_TW;
        const int g = lane >> 2, tq = lane & 3;
#pragma unroll
        for (int m = 0; m < 2; ++m)
#pragma unroll
            for (int hh = 0; hh < 2; ++hh)
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int co = tq * 2 + j, w = w0 + m * 16 + g + hh * 8;
                    if (co < Cout && h < H && w < W) {
                        const float val = acc[m][hh * 2 + j] + bias[co];
                        const size_t o = (((size_t)n * Cout + co) * H + h) * W + w;
                        if (mse_acc) { const float d = val - (x1[o] - x0[o]); sq += d * d; }
                        if (mode == 0) xv[o] = val;
                        else if (mode == 3) xv[o] = (val - (x1[o] - x0[o])) * dt;
                        else if (mode == 1) {
                            const float nx = xv[o] + val * dt;
                            xv[o] = nx;
                            if (traj) traj[o] = nx;
                        }
                    }
                }
        __syncthreads();  // all warps done with this buffer before it is refilled two iterations later
    }
    cp_async_wait<0>();
    if (mse_acc) {
        sq = warp_sum(sq);
        if (lane == 0) red[warp] = sq;
        __syncthreads();
        if (tid == 0) {
            float s = 0.f;
            for (int i = 0; i < 8; ++i) s += red[i];
            atomicAdd(mse_acc, s);
        }
    }
}
// ---------------------------------------------------------------------------------------------------------
// Output conv, second formulation (C_out <= 3, C = 16*KS): instead of shifting the A operand once per tap (nine ldmatrix
// passes over the staged tile), every staged pixel is multiplied ONCE by all 9*C_out (<= 27, padded to 32) weight columns:
//   Z[q][tap*C_out + co] = sum_c a[q][c] * W[co][c][tap]          (mma.sync m16n8k16, B fragments live in registers)
//   out[p][co]           = bias[co] + sum_tap Z[p + shift(tap)][tap*C_out + co]      (27 shared-memory reads per pixel)
// Shared-memory traffic per 8x32 tile drops from ~370 KB to ~120 KB (the first formulation is bound by exactly that), and
// the MMA count from 576 to 352.  Same modes / epilogue as output_conv_kernel.  One tile buffer per block, two blocks per SM.
// ---------------------------------------------------------------------------------------------------------
constexpr int OZ_PIX = (OC_TH + 2) * (OC_TW + 2);        // 340 staged pixels
constexpr int OZ_ROWS = ((OZ_PIX + 15) / 16) * 16;       // 352: whole m16 tiles
constexpr int OZ_ZP = 29;                                // Z row pitch in floats (odd: conflict-free gathers)
__host__ __device__ inline size_t output_conv_z_smem(int C) { return (size_t)OZ_ROWS * (C * 2 + 16) + (size_t)OZ_ROWS * OZ_ZP * 4; }
template <int KS>
__global__ void __launch_bounds__(256, 2) output_conv_z_kernel(const bf16* __restrict__ a, const bf16* __restrict__ wpk,
                                                               const float* __restrict__ bias, float* __restrict__ xv,
                                                               float* __restrict__ traj, const float* __restrict__ x0,
                                                               const float* __restrict__ x1, float* __restrict__ mse_acc,
                                                               int H, int W, int Cout, int B, int mode, float dt) {
    pdl_wait();   // programmatic dependent launch (common.cuh): no-ops unless launched with the attribute
    pdl_launch();
    constexpr int C = KS * 16;
    constexpr int pitch = C * 2 + 16;
    extern __shared__ __align__(16) uint8_t smraw[];
    uint8_t* tile = smraw;                                               // [OZ_ROWS][pitch] bf16 (rows >= 340 never staged / used)
    float* Zs = reinterpret_cast<float*>(smraw + (size_t)OZ_ROWS * pitch);   // [OZ_ROWS][OZ_ZP]
    __shared__ float red[8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, tq = lane & 3;
    const int tw = (W + OC_TW - 1) / OC_TW, th = (H + OC_TH - 1) / OC_TH;
    const int ntiles = tw * th * B;
    const int NC = 9 * Cout;                                             // used Z columns
    // B fragments: column n = tap*Cout + co of the [C x 32] weight matrix, rows k = channels.  wpk is [tap][8][C].
    uint32_t bf[KS][4][2];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
        const int n = nt * 8 + g;
        const bool ok = n < NC;
        const int tap = ok ? n / Cout : 0, co = ok ? n - tap * Cout : 0;
        const bf16* wr = wpk + (size_t)(tap * 8 + co) * C;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
            bf[ks][nt][0] = ok ? *reinterpret_cast<const uint32_t*>(wr + ks * 16 + 2 * tq) : 0u;
            bf[ks][nt][1] = ok ? *reinterpret_cast<const uint32_t*>(wr + ks * 16 + 2 * tq + 8) : 0u;
        }
    }
    float bco[3];
#pragma unroll
    for (int co = 0; co < 3; ++co) bco[co] = co < Cout ? bias[co] : 0.f;
    constexpr int vec = C >> 3;
    float sq = 0.f;
    auto stage_tile = [&](int t) {
        const int n = t / (tw * th), r = t - n * (tw * th);
        const int h0 = (r / tw) * OC_TH, w0 = (r - (r / tw) * tw) * OC_TW;
        for (int i = tid; i < OZ_PIX * vec; i += 256) {
            const int pp = i / vec, cv = i - pp * vec;
            const int hh = h0 + pp / (OC_TW + 2) - 1, ww = w0 + pp % (OC_TW + 2) - 1;
            const bool ok = hh >= 0 && hh < H && ww >= 0 && ww < W;
            const bf16* src = ok ? a + (((size_t)n * H + hh) * W + ww) * C + cv * 8 : a;
            cp_async16(smem_u32(tile + pp * pitch + cv * 16), src, ok);
        }
        cp_async_commit();
    };
    if ((int)blockIdx.x < ntiles) stage_tile(blockIdx.x);
    for (int tile_i = blockIdx.x; tile_i < ntiles; tile_i += gridDim.x) {
        const int n = tile_i / (tw * th), r = tile_i - n * (tw * th);
        const int h0 = (r / tw) * OC_TH, w0 = (r - (r / tw) * tw) * OC_TW;
        cp_async_wait<0>();
        __syncthreads();   // tile landed; every thread is past the previous gather (Z may be overwritten)
        // Z = tile x W for every staged pixel: m16 tiles mt = warp, warp + 8, ...
        for (int mt = warp; mt < OZ_ROWS / 16; mt += 8) {
            float acc[4][4];
#pragma unroll
            for (int nt = 0; nt < 4; ++nt)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[nt][j] = 0.f;
            const uint8_t* arow = tile + (mt * 16 + (lane & 15)) * pitch + (lane >> 4) * 16;
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
                uint32_t af[4];
                ldmatrix_x4(smem_u32(arow + ks * 32), af[0], af[1], af[2], af[3]);
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) mma_bf16_16816(acc[nt], af, bf[ks][nt][0], bf[ks][nt][1]);
            }
            float* z0 = Zs + (mt * 16 + g) * OZ_ZP, * z1 = z0 + 8 * OZ_ZP;
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                const int c = nt * 8 + 2 * tq;
                if (c < NC) { z0[c] = acc[nt][0]; z1[c] = acc[nt][2]; }
                if (c + 1 < NC) { z0[c + 1] = acc[nt][1]; z1[c + 1] = acc[nt][3]; }
            }
        }
        __syncthreads();
        // the tile buffer is free: fetch the next tile while this one's outputs are gathered from Z
        if (tile_i + (int)gridDim.x < ntiles) stage_tile(tile_i + gridDim.x);
        // gather: thread = output pixel (row tid / 32, column tid % 32) of the 8x32 tile
        {
            const int pr = tid >> 5, pc = tid & 31;
            const int h = h0 + pr, w = w0 + pc;
            if (h < H && w < W) {
                float o[3] = {bco[0], bco[1], bco[2]};
#pragma unroll
                for (int tap = 0; tap < 9; ++tap) {
                    const float* z = Zs + ((pr + tap / 3) * (OC_TW + 2) + pc + tap % 3) * OZ_ZP + tap * Cout;
#pragma unroll
                    for (int co = 0; co < 3; ++co)
                        if (co < Cout) o[co] += z[co];
                }
#pragma unroll
                for (int co = 0; co < 3; ++co) {
                    if (co < Cout) {
                        const float val = o[co];
                        const size_t oi = (((size_t)n * Cout + co) * H + h) * W + w;
                        if (mse_acc) { const float d = val - (x1[oi] - x0[oi]); sq += d * d; }
                        if (mode == 0) xv[oi] = val;
                        else if (mode == 3) xv[oi] = (val - (x1[oi] - x0[oi])) * dt;
                        else if (mode == 1) {
                            const float nx = xv[oi] + val * dt;
                            xv[oi] = nx;
                            if (traj) traj[oi] = nx;
                        }
                    }
                }
            }
        }
    }
    if (mse_acc) {
        sq = warp_sum(sq);
        if (lane == 0) red[warp] = sq;
        __syncthreads();
        if (tid == 0) {
            float t = 0.f;
            for (int i = 0; i < 8; ++i) t += red[i];
            atomicAdd(mse_acc, t);
        }
    }
}

// OIHW fp32 [co][c][3][3] -> bf16 [tap][8][C] (rows co >= Cout are zero): the B operand of output_conv_kernel
__global__ void pack_output_weight_kernel(const float* __restrict__ src, bf16* __restrict__ dst, int Cout, int C) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 9 * 8 * C) {
        const int c = i % C, nrow = (i / C) % 8, tap = i / (8 * C);
        dst[i] = __float2bfloat16_rn(nrow < Cout ? src[((size_t)nrow * C + c) * 9 + tap] : 0.f);
    }
}

// ---------------------------------------------------------------------------------------------------------
// small utilities
// ---------------------------------------------------------------------------------------------------------
// fp32 OIHW conv weight -> bf16 [O][k_off + (kh*KW+kw)*I + i] inside a row of length Ktot.
__global__ void pack_conv_weight_kernel(const float* __restrict__ src, bf16* __restrict__ dst, int O, int I, int KK,
                                        int Ktot, int k_off) {
    const size_t total = (size_t)O * I * KK;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int o = (int)(idx / ((size_t)I * KK));
        const int r = (int)(idx - (size_t)o * I * KK);
        const int i = r / KK, tap = r - i * KK;
        dst[(size_t)o * Ktot + k_off + tap * I + i] = __float2bfloat16_rn(src[idx]);
    }
}
__global__ void unpack_conv_weight_kernel(const bf16* __restrict__ src, float* __restrict__ dst, int O, int I, int KK,
                                          int Ktot, int k_off) {
    const size_t total = (size_t)O * I * KK;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int o = (int)(idx / ((size_t)I * KK));
        const int r = (int)(idx - (size_t)o * I * KK);
        const int i = r / KK, tap = r - i * KK;
        dst[idx] = __bfloat162float(src[(size_t)o * Ktot + k_off + tap * I + i]);
    }
}
// Nearest-x2 upsample folded into a 3x3 conv: per output phase (py,px) a 2x2 kernel whose taps are sums of the 3x3
// taps that land on the same input pixel.  dst[((py*2+px)*O + o)][(a*2+b)*I + i], bf16, sums formed in fp32.
__global__ void pack_upsample_weight_kernel(const float* __restrict__ src, bf16* __restrict__ dst, int O, int I) {
    const size_t total = (size_t)16 * O * I;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(idx % I);
        const int tap = (int)((idx / I) % 4);
        const int o = (int)((idx / ((size_t)4 * I)) % O);
        const int phase = (int)(idx / ((size_t)4 * I * O));
        const int py = phase >> 1, px = phase & 1, a = tap >> 1, b = tap & 1;
        // rows of the 3x3 kernel that alias onto input row i+a+py-1:  (py,a): (0,0)->{0} (0,1)->{1,2} (1,0)->{0,1} (1,1)->{2}
        const int ky0 = (py == 0) ? (a == 0 ? 0 : 1) : (a == 0 ? 0 : 2), ky1 = (py == 0) ? (a == 0 ? 0 : 2) : (a == 0 ? 1 : 2);
        const int kx0 = (px == 0) ? (b == 0 ? 0 : 1) : (b == 0 ? 0 : 2), kx1 = (px == 0) ? (b == 0 ? 0 : 2) : (b == 0 ? 1 : 2);
        const float* w = src + ((size_t)o * I + i) * 9;
        float acc = 0.f;
        for (int ky = ky0; ky <= ky1; ++ky)
            for (int kx = kx0; kx <= kx1; ++kx) acc += w[ky * 3 + kx];
        dst[idx] = __float2bfloat16_rn(acc);
    }
}
__global__ void add_vec_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = a[i] + (b ? b[i] : 0.f);
}
// NHWC bf16 -> NCHW fp32 (debug hook only).
__global__ void nhwc_to_nchw_kernel(const bf16* __restrict__ src, float* __restrict__ dst, int B, int C, int HW) {
    const size_t total = (size_t)B * C * HW;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int p = (int)(idx % HW);
        const int c = (int)((idx / HW) % C);
        const int n = (int)(idx / ((size_t)HW * C));
        dst[idx] = __bfloat162float(src[((size_t)n * HW + p) * C + c]);
    }
}
__global__ void scale_kernel(float* __restrict__ v, int n, float s) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] *= s;
}

}  // namespace rfv
