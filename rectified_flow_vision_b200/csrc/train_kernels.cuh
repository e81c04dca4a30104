// Backward / optimizer kernels of the reflow training step (models/rectified_flow.py:217-238: loss.backward(),
// clip_grad_norm_, AdamW.step) other than the tensor-core GEMMs (conv dgrad = the forward conv kernels on transposed
// weights, conv wgrad = wgrad.cuh, attention = attn.cuh).  Everything here is HBM- or latency-bound.
#pragma once
#include "common.cuh"

namespace rfv {

// Counter-based dropout mask: one 32-bit hash covers two consecutive elements (16 bits each); an element is KEPT when
// its 16 bits are >= thresh (thresh = round(p * 65536)).  The backward pass regenerates the mask from the same
// (seed, element index) instead of storing it.
__device__ __forceinline__ uint32_t mix32(uint32_t h) {
    h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
    return h;
}
__device__ __forceinline__ void dropout_mask8(uint32_t seed, uint32_t elem0, uint32_t thresh, float scale, float* m) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        uint32_t h = (seed ^ ((elem0 >> 1) + k)) * 0x9E3779B1u;   // two multiply / xor-shift rounds: enough to decorrelate
        h ^= h >> 15; h *= 0x85ebca6bu; h ^= h >> 13;                 // neighbouring counters for a Bernoulli mask
        m[2 * k] = (h & 0xFFFFu) >= thresh ? scale : 0.f;
        m[2 * k + 1] = (h >> 16) >= thresh ? scale : 0.f;
    }
}

// Per-channel block reductions.  Threads are laid out as (pixel lane, channel vector cv = tid % vpp); shared-memory fp32
// atomicAdd is a compare-and-swap loop (SASS ATOMS.CAST.SPIN), so letting all blockDim / vpp owners of a channel hit the
// same word costs a 32-way serialised spin for 64 channels.  When vpp divides 32 the lanes of a warp that own the same
// vector (lane % vpp) are first summed with shuffles; only the first vpp lanes of each warp then touch shared memory.
// Returns whether this thread still has to publish its values.
template <int K>
__device__ __forceinline__ bool warp_sum_same_cv(float* v, int vpp) {
    if (vpp >= 32 || (32 % vpp) != 0) return true;
    for (int off = vpp; off < 32; off <<= 1) {
#pragma unroll
        for (int k = 0; k < K; ++k) v[k] += __shfl_xor_sync(0xffffffffu, v[k], off);
    }
    return (int)(threadIdx.x & 31) < vpp;
}

// ---------------------------------------------------------------------------------------------------------
// GroupNorm(8)(+SiLU)(+dropout) backward over a virtual concat of up to two NHWC bf16 tensors (models/unet.py:56,62,82,224).
//   y = drop(silu(z)),  z = gamma * xhat + beta,  xhat = (x - mean) * rstd
//   dz = dy * mask * silu'(z);  dbeta = sum dz;  dgamma = sum dz * xhat
//   dx = rstd * (gamma * dz - S1/m - xhat * S2/m),  S1 = sum_group gamma*dz,  S2 = sum_group gamma*dz*xhat,  m = group size
// Pass 1 (reduce): per-(image, channel) sums A = sum_pix dz, B = sum_pix dz*xhat into cs[n][c][2].  S1, S2, dgamma,
// dbeta all derive from them.  Pass 2 (apply): dx, plus up to two addends (the block's identity / shortcut path),
// written (or accumulated) into the gradient tensors of the two sources.
// Thread mapping as gn_apply_kernel: a thread owns one 8-channel vector position for the whole block.
// ---------------------------------------------------------------------------------------------------------
struct GnBwdArgs {
    const bf16* dy;            // [B][HW][C] gradient w.r.t. the GroupNorm(+SiLU) output (C = Ca + Cb)
    const bf16* xa; const bf16* xb;
    const float* stats_a; const float* stats_b;
    const float* gamma; const float* beta;
    float* cs;                 // [B][C][2]
    uint32_t* arrive; uint32_t* ticket;   // single-pass kernel: per-image arrival counters [B], work-item ticket
    const bf16* add_cat;       // [B][HW][C]  optional addend in concat layout
    const bf16* add_a;         // [B][HW][Ca] optional addend for source a
    bf16* out_a; bf16* out_b;  // gradient tensors of the sources
    float* dgamma; float* dbeta;
    float* out_colsum;         // optional [B][ld_colsum]: per-(image, channel) sums of the produced gradient (source a only)
    int ld_colsum;
    int acc_a, acc_b;          // accumulate into out_a / out_b instead of overwriting
    int Ca, Cb, HW, slab_shift, silu, pix_per_block;
    float eps;
    uint32_t drop_thresh, seed; float drop_scale;
};

__device__ __forceinline__ void gn_group_stats(const GnBwdArgs& a, int n, int g, int cpg, float* mean, float* rstd) {
    const int slab = 1 << a.slab_shift;
    float s = 0.f, ss = 0.f;
    for (int c = g * cpg; c < (g + 1) * cpg; c += slab) {
        const float* src = (c < a.Ca) ? a.stats_a + ((size_t)n * (a.Ca >> a.slab_shift) + (c >> a.slab_shift)) * 2
                                      : a.stats_b + ((size_t)n * (a.Cb >> a.slab_shift) + ((c - a.Ca) >> a.slab_shift)) * 2;
        s += src[0];
        ss += src[1];
    }
    const float cnt = (float)cpg * (float)a.HW;
    const float m = s / cnt;
    *mean = m;
    *rstd = rsqrtf(fmaxf(ss / cnt - m * m, 0.f) + a.eps);
}

// sigmoid through the hardware tanh (one MUFU op instead of ex2 + rcp): |error| ~ 2.5e-4, far below bf16 resolution
__device__ __forceinline__ float sigmoid_fast(float z) {
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * z));
    return fmaf(0.5f, t, 0.5f);
}

// The element-wise cores below work on (even, odd) channel PAIRS with the packed fp32x2 instructions of sm_100 (FADD2 / FMUL2 /
// FFMA2: one issue slot for two elements) -- these kernels are issue-bound, and the FP32 arithmetic was half of their
// instruction stream.  dz[p] *= silu'(z), z = x*sc + sh, silu'(z) = sg * (1 + z * (1 - sg)), sg = sigmoid(z) via one tanh each.
__device__ __forceinline__ void unpack8p(const uint4& q, float2* f) {
    f[0] = unpack_bf16x2(q.x); f[1] = unpack_bf16x2(q.y); f[2] = unpack_bf16x2(q.z); f[3] = unpack_bf16x2(q.w);
}
__device__ __forceinline__ uint4 pack8p(const float2* f) {
    uint4 q;
    q.x = pack_bf16x2(f[0].x, f[0].y); q.y = pack_bf16x2(f[1].x, f[1].y);
    q.z = pack_bf16x2(f[2].x, f[2].y); q.w = pack_bf16x2(f[3].x, f[3].y);
    return q;
}
__device__ __forceinline__ void silu_grad4(float2* dz, const float2* x, const float2* sc, const float2* sh) {
    const float2 half = make_float2(0.5f, 0.5f), one = make_float2(1.f, 1.f), neg = make_float2(-1.f, -1.f);
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const float2 z = ffma2(x[p], sc[p], sh[p]);
        const float2 zh = fmul2(z, half);
        float2 t;
        asm("tanh.approx.f32 %0, %1;" : "=f"(t.x) : "f"(zh.x));
        asm("tanh.approx.f32 %0, %1;" : "=f"(t.y) : "f"(zh.y));
        const float2 sg = ffma2(half, t, half);
        const float2 q = ffma2(z, ffma2(sg, neg, one), one);
        dz[p] = fmul2(dz[p], fmul2(sg, q));
    }
}

// Arithmetic is arranged so that the per-element work is a handful of FMAs (these kernels are issue-bound, not HBM-bound,
// when written naively):  z = x*sc + sh;  silu'(z) = sg + z*(sg - sg^2);  pass 1 accumulates sum(dz) and sum(dz*x) and
// converts to sum(dz*xhat) at the end;  pass 2 is dx = dz*k1[c] + x*k2 + k3 with per-group constants k2, k3.
template <bool APPLY>
__global__ void __launch_bounds__(256, 2) gn_bwd_kernel(const GnBwdArgs a) {
    extern __shared__ float sm[];   // reduce: [2*C] partial sums
    __shared__ float gmean[8], grstd[8], gS1[8], gS2[8];
    const int C = a.Ca + a.Cb, n = blockIdx.y, cpg = C / 8;
    if (threadIdx.x < 8) {
        const int g = threadIdx.x;
        gn_group_stats(a, n, g, cpg, &gmean[g], &grstd[g]);
        if (APPLY) {
            float s1 = 0.f, s2 = 0.f;
            for (int c = g * cpg; c < (g + 1) * cpg; ++c) {
                const float gm = a.gamma[c];
                s1 += gm * a.cs[((size_t)n * C + c) * 2];
                s2 += gm * a.cs[((size_t)n * C + c) * 2 + 1];
            }
            const float inv = 1.0f / ((float)cpg * (float)a.HW);
            gS1[g] = s1 * inv;
            gS2[g] = s2 * inv;
        }
    }
    if (!APPLY)
        for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sm[i] = 0.f;
    if (APPLY && blockIdx.x == 0)
        for (int c = threadIdx.x; c < C; c += blockDim.x) {
            atomicAdd(a.dbeta + c, a.cs[((size_t)n * C + c) * 2]);
            atomicAdd(a.dgamma + c, a.cs[((size_t)n * C + c) * 2 + 1]);
        }
    __syncthreads();
    const int vpp = C >> 3, cv = threadIdx.x % vpp, pl = threadIdx.x / vpp, pstride = blockDim.x / vpp;
    const int grp = (cv * 8) / cpg;
    const float mean = gmean[grp], rstd = grstd[grp];
    float2 sc[4], sh[4];   // z = x*sc + sh with sc = rstd*gamma (also the dz coefficient of dx); (even, odd) channel pairs
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        sc[j] = make_float2(rstd * a.gamma[cv * 8 + 2 * j], rstd * a.gamma[cv * 8 + 2 * j + 1]);
        sh[j] = make_float2(a.beta[cv * 8 + 2 * j] - mean * sc[j].x, a.beta[cv * 8 + 2 * j + 1] - mean * sc[j].y);
    }
    const bool from_a = cv * 8 < a.Ca;
    const bf16* src = from_a ? a.xa + cv * 8 : a.xb + (cv * 8 - a.Ca);
    const int cs_ = from_a ? a.Ca : a.Cb;
    bf16* dst = from_a ? a.out_a + cv * 8 : a.out_b + (cv * 8 - a.Ca);
    const bool acc = from_a ? a.acc_a != 0 : a.acc_b != 0;
    const int p0 = blockIdx.x * a.pix_per_block;
    const int np = min(a.pix_per_block, a.HW - p0);
    const size_t base = (size_t)n * a.HW + p0;
    float2 accA[4], accB[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) { accA[j] = make_float2(0.f, 0.f); accB[j] = make_float2(0.f, 0.f); }
    const float s1 = APPLY ? gS1[grp] : 0.f, s2 = APPLY ? gS2[grp] : 0.f;
    const float k2s = -rstd * rstd * s2, k3s = -rstd * s1 - mean * k2s;   // dx = dz*sc + x*k2 + k3
    const float2 k2 = make_float2(k2s, k2s), k3 = make_float2(k3s, k3s);
    const bool has_cat = APPLY && a.add_cat != nullptr, has_a = APPLY && a.add_a != nullptr && from_a;
    constexpr int U = 2;
    for (int pp = pl; pp < np; pp += U * pstride) {
        uint4 qx[U], qd[U], qc[U], qa[U], qo[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {   // all loads of the batch first (memory-level parallelism)
            const int px = pp + u * pstride;
            if (px < np) {
                qx[u] = *reinterpret_cast<const uint4*>(src + (base + px) * cs_);
                qd[u] = *reinterpret_cast<const uint4*>(a.dy + (base + px) * C + cv * 8);
                if (has_cat) qc[u] = *reinterpret_cast<const uint4*>(a.add_cat + (base + px) * C + cv * 8);
                if (has_a) qa[u] = *reinterpret_cast<const uint4*>(a.add_a + (base + px) * a.Ca + cv * 8);
                if (APPLY && acc) qo[u] = *reinterpret_cast<const uint4*>(dst + (base + px) * cs_);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int px = pp + u * pstride;
            if (px >= np) continue;
            float2 x[4], dz[4];
            unpack8p(qx[u], x);
            unpack8p(qd[u], dz);
            if (a.drop_thresh) {
                float mk[8];
                dropout_mask8(a.seed, (uint32_t)((base + px) * C + cv * 8), a.drop_thresh, a.drop_scale, mk);
#pragma unroll
                for (int j = 0; j < 4; ++j) dz[j] = fmul2(dz[j], make_float2(mk[2 * j], mk[2 * j + 1]));
            }
            if (a.silu) silu_grad4(dz, x, sc, sh);
            if (!APPLY) {
#pragma unroll
                for (int j = 0; j < 4; ++j) { accA[j] = fadd2(accA[j], dz[j]); accB[j] = ffma2(dz[j], x[j], accB[j]); }
            } else {
                float2 r[4], f[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) r[j] = ffma2(dz[j], sc[j], ffma2(x[j], k2, k3));
                if (has_cat) {
                    unpack8p(qc[u], f);
#pragma unroll
                    for (int j = 0; j < 4; ++j) r[j] = fadd2(r[j], f[j]);
                }
                if (has_a) {
                    unpack8p(qa[u], f);
#pragma unroll
                    for (int j = 0; j < 4; ++j) r[j] = fadd2(r[j], f[j]);
                }
                if (acc) {
                    unpack8p(qo[u], f);
#pragma unroll
                    for (int j = 0; j < 4; ++j) r[j] = fadd2(r[j], f[j]);
                }
                if (a.out_colsum) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) accA[j] = fadd2(accA[j], r[j]);
                }
                *reinterpret_cast<uint4*>(dst + (base + px) * cs_) = pack8p(r);
            }
        }
    }
    float fA[8], fB[8];   // the pair accumulators as per-channel scalars for the reductions
#pragma unroll
    for (int j = 0; j < 4; ++j) { fA[2 * j] = accA[j].x; fA[2 * j + 1] = accA[j].y; fB[2 * j] = accB[j].x; fB[2 * j + 1] = accB[j].y; }
    if (!APPLY) {
        const bool pub_a = warp_sum_same_cv<8>(fA, vpp), pub_b = warp_sum_same_cv<8>(fB, vpp);
        if (pub_a && pub_b) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                atomicAdd(&sm[(cv * 8 + j) * 2], fA[j]);
                atomicAdd(&sm[(cv * 8 + j) * 2 + 1], rstd * (fB[j] - mean * fA[j]));   // sum dz*xhat
            }
        }
        __syncthreads();
        for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) atomicAdd(a.cs + (size_t)n * C * 2 + i, sm[i]);
    } else if (a.out_colsum) {   // dynamic shared memory [C] (the launcher sizes it)
        for (int i = threadIdx.x; i < C; i += blockDim.x) sm[i] = 0.f;
        __syncthreads();
        if (warp_sum_same_cv<8>(fA, vpp)) {
#pragma unroll
            for (int j = 0; j < 8; ++j) atomicAdd(&sm[cv * 8 + j], fA[j]);
        }
        __syncthreads();
        for (int i = threadIdx.x; i < a.Ca; i += blockDim.x) atomicAdd(a.out_colsum + (size_t)n * a.ld_colsum + i, sm[i]);
    }
}

// ---------------------------------------------------------------------------------------------------------
// Single-pass GroupNorm backward.  An image is cut into SL pixel slices, one CTA each.  A CTA bulk-copies (TMA engine, one
// mbarrier per stage) its slice of x and dy into shared memory ONCE, reduces its part of the per-channel sums while the later
// stages are still landing, adds them to the image's totals in cs[n][c][2] (global atomics), and waits on the image's arrival
// counter until all SL slices have contributed; the second pass (dx) then runs out of shared memory.  x and dy cross HBM once
// instead of twice (5 tensor passes -> 3) and arrive as asynchronous bulk copies rather than register-held vectors, which is
// what the two-pass kernels above are latency-bound on.  Slices are ~64 KB so that three CTAs share an SM and the load /
// reduce / wait / apply phases of different CTAs overlap.
// Forward progress of the wait: work items are handed out by an atomic ticket, so the SL slices of an image are taken by
// CTAs that are already running, in start order -- a waiting CTA only ever depends on CTAs that started before it or that
// start as soon as ANY earlier image (which depends on nothing later) retires.
// The kernel is instruction-bound (ncu: ~50 % issue-slot use, DRAM < 20 %), so pass 1 writes dz = dy * mask * silu'(z) back
// over its dy vector in shared memory (bf16, the precision dy itself has) and pass 2 is just dx = dz*k1 + x*k2 + k3; the
// per-channel sums go through a per-warp table with plain stores instead of shared-memory atomics (CAS loops).
// Thread mapping and argument block as gn_bwd_kernel (a.pix_per_block = HW / SL; a.arrive / a.ticket are zeroed with cs
// once per backward pass).  Shared memory: dy slice | xa slice | xb slice | red[rows][2C] | tot[2C] | mbarriers.
// ---------------------------------------------------------------------------------------------------------
constexpr int GNF_STAGES = 4;
// rows of the partial-sum table: one per warp (plain stores after warp_sum_same_cv) for up to 128 channels; wider tensors
// keep one row and add into it with shared-memory atomics (at most blockDim / (C/8) <= 8 owners per word there)
__host__ __device__ inline int gn_bwd_fused_rows(int C, int threads) {
    const int vpp = C / 8;
    return (vpp <= 16 && 32 % vpp == 0) ? threads / 32 : 1;
}
__host__ __device__ inline size_t gn_bwd_fused_smem(int C, int np, int threads) {
    return (size_t)np * C * 4 + (size_t)(gn_bwd_fused_rows(C, threads) + 1) * C * 8 + GNF_STAGES * 8 + 16;
}
__global__ void __launch_bounds__(256, 3) gn_bwd_fused_kernel(const GnBwdArgs a, int SL) {
    extern __shared__ __align__(16) uint8_t gsm[];
    __shared__ float gmean[8], grstd[8], gS1[8], gS2[8];
    __shared__ uint32_t s_ticket, s_last;
    const int C = a.Ca + a.Cb, cpg = C / 8;
    const int np = a.pix_per_block, ps = np / GNF_STAGES;          // pixels of this CTA / per stage
    bf16* dys = reinterpret_cast<bf16*>(gsm);
    bf16* xsa = dys + (size_t)np * C;
    bf16* xsb = xsa + (size_t)np * a.Ca;
    const int rows = gn_bwd_fused_rows(C, blockDim.x);
    float* red = reinterpret_cast<float*>(xsb + (size_t)np * a.Cb);    // [rows][2C] partial sums of this CTA
    float* tot = red + (size_t)rows * 2 * C;                            // [2C] whole-image sums
    uint64_t* bars = reinterpret_cast<uint64_t*>(tot + 2 * C);
    if (threadIdx.x == 0) {
        s_ticket = atomicAdd(a.ticket, 1u);
        for (int s = 0; s < GNF_STAGES; ++s) mbar_init(&bars[s], 1);
        mbar_fence_init();
    }
    if (rows == 1)
        for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) red[i] = 0.f;
    __syncthreads();
    const int n = s_ticket / SL, p0 = (s_ticket % SL) * np;
    const size_t base = (size_t)n * a.HW + p0;
    if (threadIdx.x == 0) {
        for (int s = 0; s < GNF_STAGES; ++s) {
            const size_t r0 = base + (size_t)s * ps;
            mbar_arrive_expect_tx(&bars[s], (uint32_t)(ps * C * 4));
            bulk_load(dys + (size_t)s * ps * C, a.dy + r0 * C, (uint32_t)(ps * C * 2), &bars[s]);
            bulk_load(xsa + (size_t)s * ps * a.Ca, a.xa + r0 * a.Ca, (uint32_t)(ps * a.Ca * 2), &bars[s]);
            if (a.Cb) bulk_load(xsb + (size_t)s * ps * a.Cb, a.xb + r0 * a.Cb, (uint32_t)(ps * a.Cb * 2), &bars[s]);
        }
    }
    if (threadIdx.x >= 32 && threadIdx.x < 40) gn_group_stats(a, n, threadIdx.x - 32, cpg, &gmean[threadIdx.x - 32], &grstd[threadIdx.x - 32]);
    __syncthreads();
    const int vpp = C >> 3, cv = threadIdx.x % vpp, pl = threadIdx.x / vpp, pstride = blockDim.x / vpp;
    const int grp = (cv * 8) / cpg;
    const float mean = gmean[grp], rstd = grstd[grp];
    float2 sc[4], sh[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        sc[j] = make_float2(rstd * a.gamma[cv * 8 + 2 * j], rstd * a.gamma[cv * 8 + 2 * j + 1]);
        sh[j] = make_float2(a.beta[cv * 8 + 2 * j] - mean * sc[j].x, a.beta[cv * 8 + 2 * j + 1] - mean * sc[j].y);
    }
    const bool from_a = cv * 8 < a.Ca;
    const bf16* xs = from_a ? xsa + cv * 8 : xsb + (cv * 8 - a.Ca);
    const int cs_ = from_a ? a.Ca : a.Cb;
    float2 accA[4], accB[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) { accA[j] = make_float2(0.f, 0.f); accB[j] = make_float2(0.f, 0.f); }
    for (int s = 0; s < GNF_STAGES; ++s) {
        mbar_wait(&bars[s], 0);
        for (int pp = s * ps + pl; pp < (s + 1) * ps; pp += pstride) {
            float2 x[4], dz[4];
            uint4* dzp = reinterpret_cast<uint4*>(dys + (size_t)pp * C + cv * 8);
            unpack8p(*reinterpret_cast<const uint4*>(xs + (size_t)pp * cs_), x);
            unpack8p(*dzp, dz);
            if (a.drop_thresh) {
                float mk[8];
                dropout_mask8(a.seed, (uint32_t)((base + pp) * C + cv * 8), a.drop_thresh, a.drop_scale, mk);
#pragma unroll
                for (int j = 0; j < 4; ++j) dz[j] = fmul2(dz[j], make_float2(mk[2 * j], mk[2 * j + 1]));
            }
            if (a.silu) silu_grad4(dz, x, sc, sh);
#pragma unroll
            for (int j = 0; j < 4; ++j) { accA[j] = fadd2(accA[j], dz[j]); accB[j] = ffma2(dz[j], x[j], accB[j]); }
            if (a.drop_thresh || a.silu) *dzp = pack8p(dz);   // pass 2 reads dz, not dy
        }
    }
    {
        float fA[8], fB[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) { fA[2 * j] = accA[j].x; fA[2 * j + 1] = accA[j].y; fB[2 * j] = accB[j].x; fB[2 * j + 1] = accB[j].y; }
        const bool pub_a = warp_sum_same_cv<8>(fA, vpp), pub_b = warp_sum_same_cv<8>(fB, vpp);
        if (pub_a && pub_b) {
            float* dstp = red + (size_t)(rows > 1 ? threadIdx.x >> 5 : 0) * 2 * C + cv * 16;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float va = fA[j], vb = rstd * (fB[j] - mean * fA[j]);   // sum dz, sum dz*xhat
                if (rows > 1) { dstp[j * 2] = va; dstp[j * 2 + 1] = vb; }
                else { atomicAdd(dstp + j * 2, va); atomicAdd(dstp + j * 2 + 1, vb); }
            }
        }
    }
    __syncthreads();
    float* csn = a.cs + (size_t)n * C * 2;
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
        float t = 0.f;
        for (int r = 0; r < rows; ++r) t += red[(size_t)r * 2 * C + i];
        atomicAdd(csn + i, t);
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t prev = atomicAdd(a.arrive + n, 1u);
        s_last = prev == (uint32_t)SL - 1 ? 1u : 0u;
        if (!s_last) {
            const volatile uint32_t* flag = a.arrive + n;
            while (*flag < (uint32_t)SL) __nanosleep(100);
        }
        __threadfence();
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) tot[i] = __ldcg(csn + i);
    __syncthreads();
    if (threadIdx.x < 8) {
        const int g = threadIdx.x;
        float s1 = 0.f, s2 = 0.f;
        for (int c = g * cpg; c < (g + 1) * cpg; ++c) {
            const float gm = a.gamma[c];
            s1 += gm * tot[c * 2];
            s2 += gm * tot[c * 2 + 1];
        }
        const float inv = 1.0f / ((float)cpg * (float)a.HW);
        gS1[g] = s1 * inv;
        gS2[g] = s2 * inv;
    }
    if (s_last)   // the CTA that completed the image folds its totals into the parameter gradients
        for (int c = threadIdx.x; c < C; c += blockDim.x) {
            atomicAdd(a.dbeta + c, tot[c * 2]);
            atomicAdd(a.dgamma + c, tot[c * 2 + 1]);
        }
    __syncthreads();
    const float k2s = -rstd * rstd * gS2[grp], k3s = -rstd * gS1[grp] - mean * k2s;   // dx = dz*sc + x*k2 + k3
    const float2 k2 = make_float2(k2s, k2s), k3 = make_float2(k3s, k3s);
    bf16* dst = from_a ? a.out_a + cv * 8 : a.out_b + (cv * 8 - a.Ca);
    const bool acc = from_a ? a.acc_a != 0 : a.acc_b != 0;
    const bool has_cat = a.add_cat != nullptr, has_a = a.add_a != nullptr && from_a;
#pragma unroll
    for (int j = 0; j < 4; ++j) accA[j] = make_float2(0.f, 0.f);
    for (int pp = pl; pp < np; pp += pstride) {
        uint4 qc, qa, qo;   // optional addends straight from global memory, issued before the arithmetic
        if (has_cat) qc = *reinterpret_cast<const uint4*>(a.add_cat + (base + pp) * C + cv * 8);
        if (has_a) qa = *reinterpret_cast<const uint4*>(a.add_a + (base + pp) * a.Ca + cv * 8);
        if (acc) qo = *reinterpret_cast<const uint4*>(dst + (base + pp) * cs_);
        float2 x[4], dz[4], r[4], f[4];
        unpack8p(*reinterpret_cast<const uint4*>(xs + (size_t)pp * cs_), x);
        unpack8p(*reinterpret_cast<const uint4*>(dys + (size_t)pp * C + cv * 8), dz);
#pragma unroll
        for (int j = 0; j < 4; ++j) r[j] = ffma2(dz[j], sc[j], ffma2(x[j], k2, k3));
        if (has_cat) {
            unpack8p(qc, f);
#pragma unroll
            for (int j = 0; j < 4; ++j) r[j] = fadd2(r[j], f[j]);
        }
        if (has_a) {
            unpack8p(qa, f);
#pragma unroll
            for (int j = 0; j < 4; ++j) r[j] = fadd2(r[j], f[j]);
        }
        if (acc) {
            unpack8p(qo, f);
#pragma unroll
            for (int j = 0; j < 4; ++j) r[j] = fadd2(r[j], f[j]);
        }
        if (a.out_colsum) {
#pragma unroll
            for (int j = 0; j < 4; ++j) accA[j] = fadd2(accA[j], r[j]);
        }
        *reinterpret_cast<uint4*>(dst + (base + pp) * cs_) = pack8p(r);
    }
    if (a.out_colsum) {
        __syncthreads();
        for (int i = threadIdx.x; i < C; i += blockDim.x) tot[i] = 0.f;
        __syncthreads();
        float fA[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) { fA[2 * j] = accA[j].x; fA[2 * j + 1] = accA[j].y; }
        if (warp_sum_same_cv<8>(fA, vpp)) {
#pragma unroll
            for (int j = 0; j < 8; ++j) atomicAdd(&tot[cv * 8 + j], fA[j]);
        }
        __syncthreads();
        for (int i = threadIdx.x; i < a.Ca; i += blockDim.x) atomicAdd(a.out_colsum + (size_t)n * a.ld_colsum + i, tot[i]);
    }
}

// Per-(image, channel) and per-channel sums of an NHWC bf16 tensor: conv-bias gradients and the gradient of the
// per-block time projection (models/unet.py:59-60 adds it to every pixel).  grid (pixel chunks, B).
__global__ void __launch_bounds__(256) colsum_kernel(const bf16* __restrict__ dy, float* __restrict__ out_nc, int ld_nc,
                                                     float* __restrict__ out_c1, float* __restrict__ out_c2, int C, int HW,
                                                     int pix_per_block) {
    extern __shared__ float sm[];  // [C]
    const int n = blockIdx.y;
    for (int i = threadIdx.x; i < C; i += blockDim.x) sm[i] = 0.f;
    __syncthreads();
    const int vpp = C >> 3, cv = threadIdx.x % vpp, pl = threadIdx.x / vpp, pstride = blockDim.x / vpp;
    const int p0 = blockIdx.x * pix_per_block, np = min(pix_per_block, HW - p0);
    const size_t base = (size_t)n * HW + p0;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    constexpr int U = 8;
    for (int pp = pl; pp < np; pp += U * pstride) {
        uint4 q[U];
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (pp + u * pstride < np) q[u] = *reinterpret_cast<const uint4*>(dy + (base + pp + u * pstride) * C + cv * 8);
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (pp + u * pstride < np) {
                float f[8];
                unpack8(q[u], f);
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[j] += f[j];
            }
    }
    if (warp_sum_same_cv<8>(acc, vpp) && threadIdx.x < pstride * vpp) {
#pragma unroll
        for (int j = 0; j < 8; ++j) atomicAdd(&sm[cv * 8 + j], acc[j]);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const float v = sm[c];
        if (out_nc) atomicAdd(out_nc + (size_t)n * ld_nc + c, v);
        if (out_c1) atomicAdd(out_c1 + c, v);
        if (out_c2) atomicAdd(out_c2 + c, v);
    }
}

// Nearest x2 upsample of an NHWC bf16 tensor (materialised only for the upsample conv's weight gradient).
__global__ void upsample2x_kernel(const bf16* __restrict__ in, bf16* __restrict__ out, int B, int H, int W, int C) {
    const int vpp = C >> 3;
    const size_t total = (size_t)B * 4 * H * W * vpp;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int cv = (int)(idx % vpp);
        size_t r = idx / vpp;
        const int ow = (int)(r % (2 * W)); r /= (2 * W);
        const int oh = (int)(r % (2 * H));
        const int n = (int)(r / (2 * H));
        reinterpret_cast<uint4*>(out)[idx] = *reinterpret_cast<const uint4*>(in + (((size_t)n * H + (oh >> 1)) * W + (ow >> 1)) * C + cv * 8);
    }
}
// 2x2 sum pool (the adjoint of the nearest upsample): out[n,i,j,:] (+)= sum of the four children.
__global__ void sumpool2x2_kernel(const bf16* __restrict__ in, bf16* __restrict__ out, int B, int H, int W, int C, int accumulate) {
    const int vpp = C >> 3;
    const size_t total = (size_t)B * H * W * vpp;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int cv = (int)(idx % vpp);
        size_t r = idx / vpp;
        const int w = (int)(r % W); r /= W;
        const int h = (int)(r % H);
        const int n = (int)(r / H);
        float s[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) s[j] = 0.f;
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int b = 0; b < 2; ++b) {
                float f[8];
                unpack8(*reinterpret_cast<const uint4*>(in + (((size_t)n * 2 * H + 2 * h + a) * 2 * W + 2 * w + b) * C + cv * 8), f);
#pragma unroll
                for (int j = 0; j < 8; ++j) s[j] += f[j];
            }
        uint4* o = reinterpret_cast<uint4*>(out) + idx;
        if (accumulate) {
            float f[8];
            unpack8(*o, f);
#pragma unroll
            for (int j = 0; j < 8; ++j) s[j] += f[j];
        }
        *o = pack8(s);
    }
}

// ---------------------------------------------------------------------------------------------------------
// The two thin convolutions (input conv 3->mc, output conv mc->3; models/unet.py:165,226) get their weight gradient from
// the tcgen05 wgrad kernel too: the 3-channel NCHW fp32 tensor (x_t, or the loss gradient dv) is widened to a 64-channel
// NHWC bf16 tensor (zero padding), the kernel writes a padded [O][9][Ipad] fp32 gradient into a scratch buffer and the real
// rows / columns are added into the parameter's slot in reference layout.
// ---------------------------------------------------------------------------------------------------------
// out[n, pix, 0:C] = (1-t) x0 + t x1 (or x0), out[n, pix, C:64] = 0
__global__ void __launch_bounds__(256) pad_to_nhwc64_kernel(const float* __restrict__ x0, const float* __restrict__ x1,
                                                            const float* __restrict__ tvec, bf16* __restrict__ out, int B, int C, int HW) {
    const size_t total = (size_t)B * HW;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const size_t n = idx / HW, p = idx - n * HW;
        const float tb = x1 ? tvec[n] : 0.f;
        float f[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float v = 0.f;
            if (j < C) {
                const size_t o = (n * C + j) * HW + p;
                v = x0[o];
                if (x1) v = (1.0f - tb) * v + tb * x1[o];
            }
            f[j] = v;
        }
        uint4* o = reinterpret_cast<uint4*>(out + idx * 64);
        o[0] = pack8(f);
#pragma unroll
        for (int i = 1; i < 8; ++i) o[i] = make_uint4(0, 0, 0, 0);
    }
}
// g[o][i][tap] += scratch[o*(9*Ipad) + tap*Ipad + i]   for o < O, i < I
__global__ void extract_wgrad_kernel(const float* __restrict__ scratch, float* __restrict__ g, int O, int I, int Ipad) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= O * I * 9) return;
    const int tap = idx % 9, i = (idx / 9) % I, o = idx / (9 * I);
    g[idx] += scratch[(size_t)o * 9 * Ipad + tap * Ipad + i];
}

// Per-channel sums of an NCHW fp32 tensor with few channels (output-conv bias gradient).
__global__ void nchw_channel_sum_kernel(const float* __restrict__ x, float* __restrict__ out, int B, int C, int HW) {
    __shared__ float red[8];
    const int c = blockIdx.y;
    float s = 0.f;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (size_t)B * HW; i += (size_t)gridDim.x * blockDim.x) {
        const size_t n = i / HW, p = i - n * HW;
        s += x[(n * C + c) * HW + p];
    }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
        atomicAdd(out + c, t);
    }
}

// ---------------------------------------------------------------------------------------------------------
// Small dense layers of the time MLP (models/unet.py:157-162 and ResidualBlock.time_mlp :43-46), fp32.
// ---------------------------------------------------------------------------------------------------------
// dW[o][i] += sum_b dy[b*ldy + o] * x[b*ldx + i];  db[o] += sum_b dy[b*ldy + o].  The output rows may be scattered over
// several parameter tensors (the per-block time projections): `segs` maps row ranges to destinations.
struct LinSeg { int row0, rows; float* dW; float* db; float* db2; };
constexpr int LIN_MAX_SEGS = 48;
struct LinSegs { int n; LinSeg s[LIN_MAX_SEGS]; };
__global__ void __launch_bounds__(256) lin_wgrad_kernel(const float* __restrict__ dy, int ldy, const float* __restrict__ x, int ldx,
                                                        const __grid_constant__ LinSegs segs, int rows, int O, int I) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= O * I) return;
    const int o = idx / I, i = idx - o * I;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f, sb = 0.f;
    int b = 0;
    for (; b + 4 <= rows; b += 4) {
        const float d0 = dy[(size_t)b * ldy + o], d1 = dy[(size_t)(b + 1) * ldy + o], d2 = dy[(size_t)(b + 2) * ldy + o],
                    d3 = dy[(size_t)(b + 3) * ldy + o];
        const float x0 = x[(size_t)b * ldx + i], x1 = x[(size_t)(b + 1) * ldx + i], x2 = x[(size_t)(b + 2) * ldx + i],
                    x3 = x[(size_t)(b + 3) * ldx + i];
        s0 = fmaf(d0, x0, s0); s1 = fmaf(d1, x1, s1); s2 = fmaf(d2, x2, s2); s3 = fmaf(d3, x3, s3);
        sb += (d0 + d1) + (d2 + d3);
    }
    for (; b < rows; ++b) {
        const float d = dy[(size_t)b * ldy + o];
        s0 = fmaf(d, x[(size_t)b * ldx + i], s0);
        sb += d;
    }
    const float s = (s0 + s1) + (s2 + s3);
    int k = 0;
    while (k + 1 < segs.n && o >= segs.s[k].row0 + segs.s[k].rows) ++k;
    const LinSeg& sg = segs.s[k];
    const int ro = o - sg.row0;
    sg.dW[(size_t)ro * I + i] += s;
    if (i == 0) {
        if (sg.db) sg.db[ro] += sb;
        if (sg.db2) sg.db2[ro] += sb;
    }
}
// dx[b][i] = sum_o dy[b*ldy + o] * W[o][i]   (optionally * silu'(z[b][i]))
__global__ void __launch_bounds__(256) lin_dgrad_kernel(const float* __restrict__ dy, int ldy, const float* __restrict__ Wt,
                                                        float* __restrict__ dx, const float* __restrict__ z, int rows, int O, int I) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows * I) return;
    const int b = idx / I, i = idx - b * I;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    const float* dr = dy + (size_t)b * ldy;
    int o = 0;
    for (; o + 4 <= O; o += 4) {
        const float d0 = dr[o], d1 = dr[o + 1], d2 = dr[o + 2], d3 = dr[o + 3];
        const float w0 = Wt[(size_t)o * I + i], w1 = Wt[(size_t)(o + 1) * I + i], w2 = Wt[(size_t)(o + 2) * I + i],
                    w3 = Wt[(size_t)(o + 3) * I + i];
        s0 = fmaf(d0, w0, s0); s1 = fmaf(d1, w1, s1); s2 = fmaf(d2, w2, s2); s3 = fmaf(d3, w3, s3);
    }
    for (; o < O; ++o) s0 = fmaf(dr[o], Wt[(size_t)o * I + i], s0);
    float s = (s0 + s1) + (s2 + s3);
    if (z) {
        const float zz = z[idx];
        const float sg = 1.0f / (1.0f + __expf(-zz));
        s *= sg * (1.0f + zz * (1.0f - sg));
    }
    dx[idx] = s;
}

// ---------------------------------------------------------------------------------------------------------
// Weight layouts of the data-gradient convolutions (the forward kernels run on these)
// ---------------------------------------------------------------------------------------------------------
// OIHW fp32 [O][I][KK] -> bf16 [I][KK*O]:  dst[i][t*O + o] = src[o][i][KK-1-t]  (transposed, spatially flipped)
__global__ void pack_conv_weight_T_kernel(const float* __restrict__ src, bf16* __restrict__ dst, int O, int I, int KK, int i_total, int i_off) {
    const size_t total = (size_t)O * I * KK;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int o = (int)(idx / ((size_t)I * KK));
        const int r = (int)(idx - (size_t)o * I * KK);
        const int i = r / KK, tap = r - i * KK;
        (void)i_total;
        dst[((size_t)(i_off + i) * KK + (KK - 1 - tap)) * O + o] = __float2bfloat16_rn(src[idx]);
    }
}
// Data gradient of a 3x3 stride-2 pad-1 conv as four sub-pixel phases of 2x2 taps over dY (the layout the
// sub-pixel upsample kernel consumes): dst[(phase*I + i)][(a*2+b)*O + o], rows (py,a) -> ky: (0,1)->1 (1,0)->2 (1,1)->0.
__global__ void pack_down_dgrad_weight_kernel(const float* __restrict__ src, bf16* __restrict__ dst, int O, int I) {
    const size_t total = (size_t)16 * O * I;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int o = (int)(idx % O);
        const int tap = (int)((idx / O) % 4);
        const int i = (int)((idx / ((size_t)4 * O)) % I);
        const int phase = (int)(idx / ((size_t)4 * O * I));
        const int py = phase >> 1, px = phase & 1, a = tap >> 1, b = tap & 1;
        const int ky = py == 0 ? (a == 1 ? 1 : -1) : (a == 0 ? 2 : 0);
        const int kx = px == 0 ? (b == 1 ? 1 : -1) : (b == 0 ? 2 : 0);
        const float v = (ky >= 0 && kx >= 0) ? src[((size_t)o * I + i) * 9 + ky * 3 + kx] : 0.f;
        dst[idx] = __float2bfloat16_rn(v);
    }
}
// Data gradient of nearest-x2 upsample + 3x3 conv as a 16-tap stride-2 conv over the four parity views of dY:
// dst[i][((py*2+px)*4 + a*2+b)*O + o] = sum of src[o][i][ky][kx] over ky in KY(py,a), kx in KX(px,b)  (fp32 sum, then bf16)
__global__ void pack_up_dgrad_weight_kernel(const float* __restrict__ src, bf16* __restrict__ dst, int O, int I) {
    const size_t total = (size_t)16 * O * I;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int o = (int)(idx % O);
        const int tap = (int)((idx / O) % 16);
        const int i = (int)(idx / ((size_t)16 * O));
        const int py = tap >> 3, px = (tap >> 2) & 1, a = (tap >> 1) & 1, b = tap & 1;
        const int ky0 = py == 0 ? (a == 0 ? 0 : 1) : (a == 0 ? 0 : 2), ky1 = py == 0 ? (a == 0 ? 0 : 2) : (a == 0 ? 1 : 2);
        const int kx0 = px == 0 ? (b == 0 ? 0 : 1) : (b == 0 ? 0 : 2), kx1 = px == 0 ? (b == 0 ? 0 : 2) : (b == 0 ? 1 : 2);
        const float* w = src + ((size_t)o * I + i) * 9;
        float acc = 0.f;
        for (int ky = ky0; ky <= ky1; ++ky)
            for (int kx = kx0; kx <= kx1; ++kx) acc += w[ky * 3 + kx];
        dst[idx] = __float2bfloat16_rn(acc);
    }
}
// Output conv data gradient in the shape input_conv_kernel consumes: wt[(s*9 + t)][c] = W[s][c][8-t], fp32.
__global__ void pack_output_dgrad_weight_kernel(const float* __restrict__ src, float* __restrict__ dst, int Co, int C) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < Co * 9 * C) {
        const int c = i % C, t = (i / C) % 9, s = i / (9 * C);
        dst[i] = src[((size_t)s * C + c) * 9 + (8 - t)];
    }
}

// ---------------------------------------------------------------------------------------------------------
// Global-norm clipping + AdamW (torch.nn.utils.clip_grad_norm_(params, 1.0); torch.optim.AdamW defaults)
// ---------------------------------------------------------------------------------------------------------
// Deterministic two-stage sum of squares (fixed grid, fixed summation order): data-parallel replicas must compute the SAME
// clip coefficient from the same all-reduced gradient or they drift apart bit by bit.
constexpr int SUMSQ_BLOCKS = 1024;
__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ g, size_t n, float scale, float* __restrict__ partial) {
    __shared__ float red[256];
    float s = 0.f;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float v = g[i] * scale;
        s = fmaf(v, v, s);
    }
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = red[0];
}
__global__ void __launch_bounds__(256) sumsq_final_kernel(const float* __restrict__ partial, int n, float* __restrict__ out) {
    __shared__ float red[256];
    float s = 0.f;
    for (int i = threadIdx.x; i < n; i += 256) s += partial[i];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = red[0];
}

struct AdamSeg {      // one parameter tensor
    long long goff;   // offset of its slot in the flat gradient / moment buffers
    long long numel;
    float* master;    // engine fp32 copy (reference layout)
    float* bound;     // caller's parameter storage (torch), or null
    int O, I, KK;     // conv weight: slot layout [O][KK][I] vs reference [O][I][KK]; KK == 0: identical layouts
    int ld;           // slot row length (floats) -- rows of a packed conv gradient may be longer than KK*I
    int koff;         // column of this tensor inside the slot row
    int no_decay;     // unused (torch AdamW decays every parameter)
};
struct AdamHyper { float lr, beta1, beta2, eps, wd, max_norm, grad_scale, bc1, bc2; };

constexpr int ADAM_CHUNK = 4096;
__global__ void __launch_bounds__(256) adamw_kernel(const AdamSeg* __restrict__ segs, const int2* __restrict__ blocks,
                                                    const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                                    const float* __restrict__ norm2, AdamHyper h) {
    const int2 bk = blocks[blockIdx.x];
    const AdamSeg s = segs[bk.x];
    float clip = 1.0f;
    if (h.max_norm > 0.f) {
        const float tn = sqrtf(*norm2);
        clip = fminf(1.0f, h.max_norm / (tn + 1e-6f));
    }
    const float gs = h.grad_scale * clip;
    const long long j0 = (long long)bk.y * ADAM_CHUNK;
    for (long long j = j0 + threadIdx.x; j < j0 + ADAM_CHUNK && j < s.numel; j += blockDim.x) {
        long long gi, ri;   // gradient-slot index, reference-layout index
        if (s.KK > 0) {     // j enumerates the slot order [o][tap][i]
            const int o = (int)(j / ((long long)s.KK * s.I));
            const int r = (int)(j - (long long)o * s.KK * s.I);
            const int tap = r / s.I, i = r - tap * s.I;
            gi = s.goff + (long long)o * s.ld + s.koff + r;
            ri = ((long long)o * s.I + i) * s.KK + tap;
        } else { gi = s.goff + j; ri = j; }
        const float gr = g[gi] * gs;
        const float mm = h.beta1 * m[gi] + (1.0f - h.beta1) * gr;
        const float vv = h.beta2 * v[gi] + (1.0f - h.beta2) * gr * gr;
        m[gi] = mm;
        v[gi] = vv;
        float p = s.master[ri];
        p *= (1.0f - h.lr * h.wd);
        p -= h.lr * (mm / h.bc1) / (sqrtf(vv / h.bc2) + h.eps);
        s.master[ri] = p;
        if (s.bound) s.bound[ri] = p;
    }
}
__global__ void sqrt_kernel(float* v) { if (threadIdx.x == 0) *v = sqrtf(*v); }
// Gradient slot -> reference layout (tests / rfv_get_grad).
__global__ void unpack_grad_kernel(const float* __restrict__ g, float* __restrict__ dst, AdamSeg s, float scale) {
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < s.numel; j += (long long)gridDim.x * blockDim.x) {
        long long gi, ri;
        if (s.KK > 0) {
            const int o = (int)(j / ((long long)s.KK * s.I));
            const int r = (int)(j - (long long)o * s.KK * s.I);
            const int tap = r / s.I, i = r - tap * s.I;
            gi = s.goff + (long long)o * s.ld + s.koff + r;
            ri = ((long long)o * s.I + i) * s.KK + tap;
        } else { gi = s.goff + j; ri = j; }
        dst[ri] = g[gi] * scale;
    }
}

}  // namespace rfv
