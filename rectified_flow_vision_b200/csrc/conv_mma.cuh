// Generic implicit-GEMM convolution on the legacy tensor-core path (mma.sync m16n8k16, cp.async pipeline).
// Handles everything the tcgen05 kernel does not take: odd spatial sizes, folded nearest-upsampling, and serves as
// the A/B reference for the tcgen05 kernel on the device.  Replaces nn.Conv2d call sites models/unet.py:38,41,51,185,217.
#pragma once
#include "common.cuh"
#include "conv_params.h"

namespace rfv {

constexpr int MMA_BM = 128, MMA_BN = 64, MMA_BK = 32, MMA_STAGES = 3, MMA_LD = 40;

__global__ void __launch_bounds__(256) conv_mma_kernel(const ConvParams p) {
    pdl_wait();   // programmatic dependent launch (common.cuh): no-ops unless launched with the attribute
    pdl_launch();
    __shared__ __align__(16) bf16 As[MMA_STAGES][MMA_BM][MMA_LD];
    __shared__ __align__(16) bf16 Bs[MMA_STAGES][MMA_BN][MMA_LD];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int warp_m = warp >> 1, warp_n = warp & 1;
    const long m0 = (long)blockIdx.x * MMA_BM;
    const int n0 = blockIdx.y * MMA_BN;
    const int HoWo = p.Ho * p.Wo;
    const long M = (long)p.B * HoWo;
    const int pad = p.ks >> 1;
    const int Hl = p.ups ? 2 * p.H0 : p.H0, Wl = p.ups ? 2 * p.W0 : p.W0;

    // fixed per-thread gather coordinates: two A rows and one B row, one 16-byte chunk (8 k-values) each
    const int jchunk = tid & 3;
    int an[2], aho[2], awo[2];
    bool aok[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        long m = m0 + (tid >> 2) + 64 * i;
        aok[i] = m < M;
        long mm = aok[i] ? m : 0;
        an[i] = (int)(mm / HoWo);
        int r = (int)(mm - (long)an[i] * HoWo);
        aho[i] = r / p.Wo;
        awo[i] = r - aho[i] * p.Wo;
    }
    const bf16* wrow = p.w + (size_t)(n0 + (tid >> 2)) * p.Ktot;

    auto load_stage = [&](int stage, int kb) {
        const int k = kb * MMA_BK + jchunk * 8;
        cp_async16(smem_u32(&Bs[stage][tid >> 2][jchunk * 8]), k < p.Ktot ? (const void*)(wrow + k) : (const void*)p.w, k < p.Ktot);
        if (k < p.K0) {
            const int tap = k / p.C0;
            const int c = k - tap * p.C0;
            const int dy = tap / p.ks - pad, dx = tap - (tap / p.ks) * p.ks - pad;
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                int hi = aho[i] * p.stride + dy, wi = awo[i] * p.stride + dx;
                bool v = aok[i] && hi >= 0 && hi < Hl && wi >= 0 && wi < Wl;
                if (p.ups) { hi >>= 1; wi >>= 1; }
                const bf16* src = v ? p.a0 + (((size_t)an[i] * p.H0 + hi) * p.W0 + wi) * p.C0 + c : p.a0;
                cp_async16(smem_u32(&As[stage][(tid >> 2) + 64 * i][jchunk * 8]), src, v);
            }
        } else {
            int k1 = k - p.K0;
            const bf16* base = p.s1a;
            int C = p.C1a;
            if (k1 >= p.C1a) { base = p.s1b; C = p.C1b; k1 -= p.C1a; }
            const bool kin = k < p.Ktot;
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                bool v = kin && aok[i];
                const bf16* src = v ? base + (((size_t)an[i] * p.Ho + aho[i]) * p.Wo + awo[i]) * C + k1 : p.a0;
                cp_async16(smem_u32(&As[stage][(tid >> 2) + 64 * i][jchunk * 8]), src, v);
            }
        }
    };

    float acc[2][4][4];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[a][b][c] = 0.f;

    const int nk = (p.Ktot + MMA_BK - 1) / MMA_BK;
#pragma unroll
    for (int s = 0; s < MMA_STAGES - 1; ++s) {
        if (s < nk) load_stage(s, s);
        cp_async_commit();
    }
    for (int kb = 0; kb < nk; ++kb) {
        cp_async_wait<MMA_STAGES - 2>();
        __syncthreads();
        {
            const int nx = kb + MMA_STAGES - 1;
            if (nx < nk) load_stage(nx % MMA_STAGES, nx);
            cp_async_commit();
        }
        const int st = kb % MMA_STAGES;
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
            uint32_t af[2][4], bfr[2][4];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
                ldmatrix_x4(smem_u32(&As[st][warp_m * 32 + mt * 16 + (lane & 15)][kk * 16 + (lane >> 4) * 8]), af[mt][0],
                            af[mt][1], af[mt][2], af[mt][3]);
#pragma unroll
            for (int np = 0; np < 2; ++np)
                ldmatrix_x4(smem_u32(&Bs[st][warp_n * 32 + np * 16 + (lane & 7) + ((lane >> 4) << 3)][kk * 16 + ((lane >> 3) & 1) * 8]),
                            bfr[np][0], bfr[np][1], bfr[np][2], bfr[np][3]);
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int nt = 0; nt < 4; ++nt)
                    mma_bf16_16816(acc[mt][nt], af[mt], bfr[nt >> 1][(nt & 1) * 2], bfr[nt >> 1][(nt & 1) * 2 + 1]);
        }
    }
    cp_async_wait<0>();

    // ---- epilogue: bias + time-embedding + residual, GroupNorm partial stats, bf16 store --------------------------
    const long wm0 = m0 + warp_m * 32;
    const bool warp_ok = wm0 < M;  // M % 32 == 0 is a launch precondition
    const int n_img = warp_ok ? (int)(wm0 / HoWo) : 0;
    const int g = lane >> 2, tq = lane & 3;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
        const int col = n0 + warp_n * 32 + nt * 8 + tq * 2;
        // per-channel addend: the conv bias, or (bias + time projection) pre-summed by the temb kernel
        const float* addp = p.temb ? p.temb + (size_t)n_img * p.temb_stride : p.bias;
        const float add0 = addp[col], add1 = addp[col + 1];
        float s = 0.f, ss = 0.f;
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                const long m = wm0 + mt * 16 + g + hh * 8;
                float v0 = acc[mt][nt][hh * 2] + add0, v1 = acc[mt][nt][hh * 2 + 1] + add1;
                if (warp_ok) {
                    const size_t o = (size_t)m * p.Cout + col;
                    if (p.resid) {
                        float2 r = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(p.resid + o));
                        v0 += r.x;
                        v1 += r.y;
                    }
                    *reinterpret_cast<uint32_t*>(p.out + o) = pack_bf16x2(v0, v1);
                    s += v0 + v1;
                    ss += v0 * v0 + v1 * v1;
                }
            }
        if (p.stats) {
            s = warp_sum(s);
            ss = warp_sum(ss);
            if (lane == 0 && warp_ok) {
                float* dst = p.stats + ((size_t)n_img * (p.Cout >> p.slab_shift) + ((n0 + warp_n * 32 + nt * 8) >> p.slab_shift)) * 2;
                atomicAdd(dst, s);
                atomicAdd(dst + 1, ss);
            }
        }
    }
}

}  // namespace rfv
