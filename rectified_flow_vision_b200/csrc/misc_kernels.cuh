// HBM-bound / small kernels around the convolutions: time embedding, input conv (+NCHW fp32 -> NHWC bf16),
// GroupNorm-apply(+SiLU)(+virtual concat), output conv fused with the Euler update (+NHWC -> NCHW fp32),
// weight packing, flow-matching interpolation.
#pragma once
#include "common.cuh"

namespace rfv {

// ---------------------------------------------------------------------------------------------------------
// Time embedding: temb_act[b, :] = SiLU(W2 . SiLU(W1 . sincos(t_b) + b1) + b2)        (models/unet.py:20-27,157-162)
// The trailing SiLU is the first op of every ResidualBlock.time_mlp (models/unet.py:43-46), shared by all blocks.
// One block per batch row (a single row when t is uniform over the batch, which is the sampling case).
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) temb_kernel(const float* __restrict__ t, int t_stride_is_zero, float t_scalar,
                                                   const float* __restrict__ w1, const float* __restrict__ b1,
                                                   const float* __restrict__ w2, const float* __restrict__ b2,
                                                   float* __restrict__ out, int mc, int td) {
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
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int o = warp; o < td; o += nw) {
        float s = 0.f;
        for (int j = lane; j < mc; j += 32) s += w1[(size_t)o * mc + j] * emb[j];
        s = warp_sum(s);
        if (lane == 0) h1[o] = silu_f(s + b1[o]);
    }
    __syncthreads();
    for (int o = warp; o < td; o += nw) {
        float s = 0.f;
        for (int j = lane; j < td; j += 32) s += w2[(size_t)o * td + j] * h1[j];
        s = warp_sum(s);
        if (lane == 0) out[(size_t)b * td + o] = silu_f(s + b2[o]);
    }
}

// All ResidualBlock time projections at once: proj[b, c] = Wcat[c, :] . temb_act[b, :] + bcat[c]   (models/unet.py:59-60)
// grid (ceil(sumC/64), ceil(rows/8)): each block keeps 8 batch rows in smem so a weight row is read once per 8 rows.
__global__ void __launch_bounds__(256) temb_proj_kernel(const float* __restrict__ act, const float* __restrict__ wcat,
                                                        const float* __restrict__ bcat, float* __restrict__ proj,
                                                        int rows, int td, int sumC) {
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

// ---------------------------------------------------------------------------------------------------------
// Input conv 3x3, C_in small (3): fp32 NCHW in -> bf16 NHWC out (+ GroupNorm slab statistics).  models/unet.py:165,234
// Optionally computes the flow-matching interpolation on the fly: x = (1-t) x0 + t x1   (models/base_flow.py:84).
// Block = 64 consecutive pixels of one image x all C_out channels; 256 threads = 64 pixels x 4 channel phases.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) input_conv_kernel(const float* __restrict__ x, const float* __restrict__ x1,
                                                         const float* __restrict__ tvec, const float* __restrict__ w,
                                                         const float* __restrict__ bias, bf16* __restrict__ out,
                                                         float* __restrict__ stats, int Cin, int H, int W, int Cout,
                                                         int slab_shift) {
    extern __shared__ float sm[];
    const int K = Cin * 9;
    float* wsm = sm;                   // [K][Cout]
    float* patch = sm + K * Cout;      // [64][K+1]
    float* st = patch + 64 * (K + 1);  // [Cout/8][2]
    const int HW = H * W;
    const int n = blockIdx.y, p0 = blockIdx.x * 64;
    for (int i = threadIdx.x; i < K * Cout; i += 256) {  // w is OIHW: [co][ci][kh][kw] -> wsm[(ci*9+kh*3+kw)][co]
        const int co = i / K, k = i - co * K;
        wsm[k * Cout + co] = w[i];
    }
    for (int i = threadIdx.x; i < (Cout / 8) * 2; i += 256) st[i] = 0.f;
    const float tb = (x1 != nullptr) ? tvec[n] : 0.f;
    for (int i = threadIdx.x; i < 64 * K; i += 256) {
        const int k = i / 64, pp = i - k * 64;  // consecutive threads -> consecutive pixels (coalesced)
        const int ci = k / 9, tap = k - ci * 9;
        const int pix = p0 + pp;
        const int h = pix / W + tap / 3 - 1, ww = pix % W + tap % 3 - 1;
        float v = 0.f;
        if (h >= 0 && h < H && ww >= 0 && ww < W) {
            const size_t o = ((size_t)n * Cin + ci) * HW + (size_t)h * W + ww;
            v = x[o];
            if (x1 != nullptr) v = (1.0f - tb) * v + tb * x1[o];
        }
        patch[pp * (K + 1) + k] = v;
    }
    __syncthreads();
    const int pp = threadIdx.x >> 2, q = threadIdx.x & 3;
    const float* pr = patch + pp * (K + 1);
    for (int co = q * 8; co < Cout; co += 32) {
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = bias[co + j];
        for (int k = 0; k < K; ++k) {
            const float a = pr[k];
            const float4 w0 = *reinterpret_cast<const float4*>(wsm + k * Cout + co);
            const float4 w1 = *reinterpret_cast<const float4*>(wsm + k * Cout + co + 4);
            acc[0] += a * w0.x; acc[1] += a * w0.y; acc[2] += a * w0.z; acc[3] += a * w0.w;
            acc[4] += a * w1.x; acc[5] += a * w1.y; acc[6] += a * w1.z; acc[7] += a * w1.w;
        }
        *reinterpret_cast<uint4*>(out + ((size_t)n * HW + p0 + pp) * Cout + co) = pack8(acc);
        if (stats) {
            float s = 0.f, ss = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) { s += acc[j]; ss += acc[j] * acc[j]; }
            // reduce over the 8 pixels of this warp that share q (lane bits 2..4)
#pragma unroll
            for (int o = 4; o < 32; o <<= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); ss += __shfl_xor_sync(0xffffffffu, ss, o); }
            if ((threadIdx.x & 31) < 4) { atomicAdd(&st[(co >> 3) * 2], s); atomicAdd(&st[(co >> 3) * 2 + 1], ss); }
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
// GroupNorm(8) apply (+SiLU) over a virtual channel-concat of up to two NHWC bf16 tensors.  models/unet.py:56,62,82,224
// Statistics come from the producers' epilogues as per-(n, slab) sums; groups are unions of whole slabs.
// grid (pixel chunks, B); each block first derives per-channel scale/shift for its image, then streams pixels
// with 16-byte loads/stores.  Pure HBM traffic: reads 2 B/elem, writes 2 B/elem.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gn_apply_kernel(const bf16* __restrict__ xa, const bf16* __restrict__ xb,
                                                       const float* __restrict__ stats_a, const float* __restrict__ stats_b,
                                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                                       bf16* __restrict__ out, int Ca, int Cb, int HW, int slab_shift,
                                                       int apply_silu, int pix_per_block, float eps) {
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
                    f[j] = apply_silu ? silu_f(y) : y;
                }
                *reinterpret_cast<uint4*>(out + (base + px) * C + cv * 8) = pack8(f);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// Output conv 3x3, C -> C_out (3), fused with what follows it at the call site:
//   mode 0  v = conv(a)                                      (UNet.forward result, models/unet.py:275)
//   mode 1  x += dt * conv(a)   [+ snapshot into traj]       (Euler update, models/base_flow.py:170)
// and optionally accumulates sum((conv(a) - target)^2), target = x1 - x0, for the loss / straightness metrics
// (models/rectified_flow.py:118,231).  a: NHWC bf16 (already GroupNorm+SiLU'd); x, v, x0, x1: NCHW fp32.
// Block = 8 x 32 output pixels, halo tile of `a` staged in shared memory (pixel pitch padded by 16 B).
// ---------------------------------------------------------------------------------------------------------
constexpr int OC_TH = 8, OC_TW = 32;
__global__ void __launch_bounds__(256) output_conv_kernel(const bf16* __restrict__ a, const float* __restrict__ w,
                                                          const float* __restrict__ bias, float* __restrict__ xv,
                                                          float* __restrict__ traj, const float* __restrict__ x0,
                                                          const float* __restrict__ x1, float* __restrict__ mse_acc,
                                                          int C, int H, int W, int Cout, int mode, float dt) {
    extern __shared__ __align__(16) uint8_t smraw[];
    const int pitch = C * 2 + 16;                                  // bytes per staged pixel
    uint8_t* tile = smraw;                                         // [(TH+2)*(TW+2)][pitch]
    float* wsm = reinterpret_cast<float*>(smraw + (OC_TH + 2) * (OC_TW + 2) * pitch);  // [Cout][9][C]
    __shared__ float red[8];
    const int n = blockIdx.z, h0 = blockIdx.y * OC_TH, w0 = blockIdx.x * OC_TW;
    for (int i = threadIdx.x; i < Cout * 9 * C; i += 256) {  // OIHW [co][c][tap] -> wsm[co][tap][c]
        const int co = i / (9 * C), r = i - co * 9 * C, c = r / 9, tap = r - c * 9;
        wsm[(co * 9 + tap) * C + c] = w[i];
    }
    const int vec = C >> 3;
    for (int i = threadIdx.x; i < (OC_TH + 2) * (OC_TW + 2) * vec; i += 256) {
        const int pp = i / vec, cv = i - pp * vec;
        const int hh = h0 + pp / (OC_TW + 2) - 1, ww = w0 + pp % (OC_TW + 2) - 1;
        uint4 q = make_uint4(0, 0, 0, 0);
        if (hh >= 0 && hh < H && ww >= 0 && ww < W)
            q = *reinterpret_cast<const uint4*>(a + (((size_t)n * H + hh) * W + ww) * C + cv * 8);
        *reinterpret_cast<uint4*>(tile + pp * pitch + cv * 16) = q;
    }
    __syncthreads();
    const int ty = threadIdx.x >> 5, tx = threadIdx.x & 31;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};  // Cout <= 4
    for (int tap = 0; tap < 9; ++tap) {
        const uint8_t* src = tile + ((ty + tap / 3) * (OC_TW + 2) + tx + tap % 3) * pitch;
        for (int cv = 0; cv < vec; ++cv) {
            float f[8];
            unpack8(*reinterpret_cast<const uint4*>(src + cv * 16), f);
            for (int co = 0; co < Cout; ++co) {
                const float* wr = wsm + (co * 9 + tap) * C + cv * 8;
                const float4 wa = *reinterpret_cast<const float4*>(wr), wb = *reinterpret_cast<const float4*>(wr + 4);
                acc[co] += f[0] * wa.x + f[1] * wa.y + f[2] * wa.z + f[3] * wa.w + f[4] * wb.x + f[5] * wb.y + f[6] * wb.z + f[7] * wb.w;
            }
        }
    }
    const int h = h0 + ty, ww = w0 + tx;
    float sq = 0.f;
    if (h < H && ww < W) {
        for (int co = 0; co < Cout; ++co) {
            const float val = acc[co] + bias[co];
            const size_t o = (((size_t)n * Cout + co) * H + h) * W + ww;
            if (mse_acc) { const float d = val - (x1[o] - x0[o]); sq += d * d; }
            if (mode == 0) xv[o] = val;
            else {
                const float nx = xv[o] + val * dt;
                xv[o] = nx;
                if (traj) traj[o] = nx;
            }
        }
    }
    if (mse_acc) {
        sq = warp_sum(sq);
        if (tx == 0) red[ty] = sq;
        __syncthreads();
        if (threadIdx.x == 0) {
            float s = 0.f;
            for (int i = 0; i < 8; ++i) s += red[i];
            atomicAdd(mse_acc, s);
        }
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
