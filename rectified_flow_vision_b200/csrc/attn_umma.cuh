// Self-attention core on tcgen05 for the default network's shape (N = 256 tokens, head dim 64; models/unet.py:84-97):
//   S = Q K^T  : one UMMA group  M = 128 queries, N = 256 keys, K = 64   (Q, K tiles K-major straight from TMA)
//   P = softmax(S / 8) in registers from TMEM, written as bf16 into 128B-swizzled shared memory (the A operand of the next GEMM)
//   O = P V    : M = 128, N = 64, K = 256 keys; V is read as an MN-major B operand (rows = keys, as TMA delivers it)
// One CTA per (128-query block, head, image), two CTAs per SM: P (64 KB) is staged over the Q / K tiles, which are dead once
// S has been computed, and O reuses the first 64 of S's 256 TMEM columns (every softmax thread has read its S row before the
// second GEMM is issued), so a CTA needs 96 KB of shared memory and 256 TMEM columns.  warp 0: TMA + MMA issue (one lane);  warps 1-4: softmax / epilogue (thread = query row = TMEM lane).
// Measured at 256 images (B200): 0.060 ms against 0.086 ms for the mma.sync kernel (attn.cuh), which stays for other shapes.
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "wgrad.cuh"   // umma_desc_mn_sw128

namespace rfv {

constexpr int AU_THREADS = 160;
constexpr int AU_SMEM = 1024 + 65536 + 32768 + 256;

__global__ void __launch_bounds__(AU_THREADS, 2)
attn_umma_kernel(const __grid_constant__ CUtensorMap mapQKV, bf16* __restrict__ out, int N, int C, float scale_log2,
                 float* __restrict__ lse) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sQ = smem;                 // [128 queries][64] bf16, 128-byte rows, swizzled
    uint8_t* sK = sQ + 16384;           // [256 keys][64]
    uint8_t* sP = smem;                 // 4 x [128 queries][64 keys]: overlays Q and K (dead after the first GEMM)
    uint8_t* sV = smem + 65536;         // [256 keys][64]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sV + 32768);
    uint64_t* ld_bar = bars;            // TMA -> MMA
    uint64_t* s_bar = bars + 1;         // S ready
    uint64_t* p_bar = bars + 2;         // P staged (128 arrivals)
    uint64_t* o_bar = bars + 3;         // O ready
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qb = blockIdx.x, head = blockIdx.y, b = blockIdx.z;
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&mapQKV);
        mbar_init(ld_bar, 1); mbar_init(s_bar, 1); mbar_init(p_bar, 128); mbar_init(o_bar, 1);
        mbar_fence_init();
    }
    if (warp == 0) tmem_alloc(tmem_slot, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    pdl_wait();   // programmatic dependent launch (common.cuh)
    if (threadIdx.x == 0) pdl_launch();

    if (warp == 0) {
        if (elect_one()) {
            const int row0 = b * N;
            mbar_arrive_expect_tx(ld_bar, 16384 + 32768 + 32768);
            tma_load_2d(sQ, &mapQKV, ld_bar, head * 64, row0 + qb * 128);
            tma_load_2d(sK, &mapQKV, ld_bar, C + head * 64, row0);
            tma_load_2d(sK + 16384, &mapQKV, ld_bar, C + head * 64, row0 + 128);
            tma_load_2d(sV, &mapQKV, ld_bar, 2 * C + head * 64, row0);
            tma_load_2d(sV + 16384, &mapQKV, ld_bar, 2 * C + head * 64, row0 + 128);
        }
        __syncwarp();
        mbar_wait(ld_bar, 0);
        tc_fence_after();
        if (elect_one()) {
            constexpr uint32_t idS = umma_idesc_bf16(128, 256);
            const uint64_t qd = umma_desc_sw128(smem_u32(sQ)), kd = umma_desc_sw128(smem_u32(sK));
#pragma unroll
            for (int j = 0; j < 4; ++j) umma_bf16(tmem, qd + 2 * j, kd + 2 * j, idS, j != 0);
            umma_commit(s_bar);
        }
        __syncwarp();
        mbar_wait(p_bar, 0);
        tc_fence_after();
        if (elect_one()) {
            constexpr uint32_t idO = umma_idesc_bf16(128, 64) | (1u << 16);   // B (= V) is MN-major
#pragma unroll
            for (int ks = 0; ks < 16; ++ks) {
                const uint64_t pd = umma_desc_sw128(smem_u32(sP + (ks >> 2) * 16384)) + 2 * (ks & 3);
                const uint64_t vd = umma_desc_mn_sw128(smem_u32(sV + ks * 2048), 1024);
                umma_bf16(tmem, pd, vd, idO, ks != 0);
            }
            umma_commit(o_bar);
        }
        __syncwarp();
    } else {
        const int q = warp & 3;                       // TMEM lane quadrant this warp may access
        const int r = q * 32 + lane;                  // query row inside the block
        const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16);
        mbar_wait(s_bar, 0);
        tc_fence_after();
        float mx = -INFINITY;
#pragma unroll 1
        for (int c = 0; c < 8; ++c) {
            uint32_t v[32];
            tmem_ld32(trow + c * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(v[i]));
        }
        const float mb = mx * scale_log2;
        float sum = 0.f;
#pragma unroll 1
        for (int c = 0; c < 8; ++c) {
            uint32_t v[32];
            tmem_ld32(trow + c * 32, v);
            tmem_ld_wait();
            float p[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) { p[i] = exp2f(fmaf(__uint_as_float(v[i]), scale_log2, -mb)); sum += p[i]; }
            // keys [c*32, c*32+32) of this row -> sub-tile c/2, 16-byte chunks (c%2)*4 .. +3
            const uint32_t base = smem_u32(sP + (c >> 1) * 16384) + (uint32_t)r * 128;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint4 pk = pack8(p + j * 8);
                const uint32_t addr = base + (uint32_t)((((c & 1) * 4 + j) ^ (r & 7)) << 4);
                asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(pk.x), "r"(pk.y), "r"(pk.z), "r"(pk.w) : "memory");
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        tc_fence_before();
        mbar_arrive(p_bar);
        mbar_wait(o_bar, 0);
        tc_fence_after();
        const float inv = 1.0f / sum;
        // log2-domain log-sum-exp per query row, same definition as attn.cuh (the backward kernels rebuild P from it)
        if (lse) lse[((size_t)b * gridDim.y + head) * N + qb * 128 + r] = mb + log2f(sum);
        bf16* orow = out + ((size_t)b * N + qb * 128 + r) * C + head * 64;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            uint32_t v[32];
            tmem_ld32(trow + c * 32, v);
            tmem_ld_wait();
            float o[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __uint_as_float(v[i]) * inv;
#pragma unroll
            for (int j = 0; j < 4; ++j) *reinterpret_cast<uint4*>(orow + c * 32 + j * 8) = pack8(o + j * 8);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 256);
}

// ---------------------------------------------------------------------------------------------------------
// The same attention core for longer sequences (N a multiple of 256, e.g. 1024 tokens at 128x128 images: BASELINE configs[4]):
// a loop over 256-key blocks with the online softmax.  Per block j:  S = Q K_j^T (one N = 256 UMMA group)  ->  softmax threads
// (thread = query row) take the block maximum, rescale their running sum by alpha = 2^((m_old - m_new) * scale), write
// P_j = 2^(S * scale - m_new * scale) as the bf16 A operand  ->  O_j = P_j V_j into 64 further TMEM columns (fresh
// accumulator)  ->  every thread folds it into its 64 fp32 output registers: o = o * alpha + O_j.  K / V blocks are double
// buffered (the TMA loads of block j+1 fly during block j); Q stays resident.  One CTA per SM (Q 16 + K 2x32 + V 2x32 + P 64 KB).
// ---------------------------------------------------------------------------------------------------------
constexpr int AK_THREADS = 160;
constexpr int AK_SMEM = 1024 + 16384 + 2 * 32768 + 2 * 32768 + 65536 + 256;

__global__ void __launch_bounds__(AK_THREADS, 1)
attn_umma_kv_kernel(const __grid_constant__ CUtensorMap mapQKV, bf16* __restrict__ out, int N, int C, float scale_log2,
                    float* __restrict__ lse) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sQ = smem;                         // [128 queries][64]
    uint8_t* sK = sQ + 16384;                   // 2 x [256 keys][64]
    uint8_t* sV = sK + 2 * 32768;               // 2 x [256 keys][64]
    uint8_t* sP = sV + 2 * 32768;               // 4 x [128 queries][64 keys]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 65536);
    uint64_t* q_bar = bars;                     // Q landed
    uint64_t* kv_bar = bars + 1;                // [2] K / V block landed
    uint64_t* s_bar = bars + 3;                 // S_j ready
    uint64_t* p_bar = bars + 4;                 // P_j staged (128 arrivals)
    uint64_t* o_bar = bars + 5;                 // O_j ready
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qb = blockIdx.x, head = blockIdx.y, b = blockIdx.z;
    const int nkb = N >> 8;
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&mapQKV);
        mbar_init(q_bar, 1); mbar_init(&kv_bar[0], 1); mbar_init(&kv_bar[1], 1);
        mbar_init(s_bar, 1); mbar_init(p_bar, 128); mbar_init(o_bar, 1);
        mbar_fence_init();
    }
    if (warp == 0) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    pdl_wait();   // programmatic dependent launch (common.cuh)
    if (threadIdx.x == 0) pdl_launch();
    const uint32_t tmem_o = tmem + 256;

    if (warp == 0) {
        const int row0 = b * N;
        auto load_kv = [&](int j) {
            const int buf = j & 1;
            mbar_arrive_expect_tx(&kv_bar[buf], 2 * 32768);
            tma_load_2d(sK + buf * 32768, &mapQKV, &kv_bar[buf], C + head * 64, row0 + j * 256);
            tma_load_2d(sK + buf * 32768 + 16384, &mapQKV, &kv_bar[buf], C + head * 64, row0 + j * 256 + 128);
            tma_load_2d(sV + buf * 32768, &mapQKV, &kv_bar[buf], 2 * C + head * 64, row0 + j * 256);
            tma_load_2d(sV + buf * 32768 + 16384, &mapQKV, &kv_bar[buf], 2 * C + head * 64, row0 + j * 256 + 128);
        };
        if (elect_one()) {
            mbar_arrive_expect_tx(q_bar, 16384);
            tma_load_2d(sQ, &mapQKV, q_bar, head * 64, row0 + qb * 128);
            load_kv(0);
        }
        __syncwarp();
        mbar_wait(q_bar, 0);
        for (int j = 0; j < nkb; ++j) {
            const int buf = j & 1;
            // buffer buf^1 held block j-1: K_{j-1} is dead since S_{j-1} completed, V_{j-1} once O_{j-1} has
            if (j >= 1) mbar_wait(o_bar, (j - 1) & 1);
            if (j + 1 < nkb && elect_one()) load_kv(j + 1);
            __syncwarp();
            mbar_wait(&kv_bar[buf], (j >> 1) & 1);
            tc_fence_after();
            if (elect_one()) {
                constexpr uint32_t idS = umma_idesc_bf16(128, 256);
                const uint64_t qd = umma_desc_sw128(smem_u32(sQ)), kd = umma_desc_sw128(smem_u32(sK + buf * 32768));
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16(tmem, qd + 2 * k, kd + 2 * k, idS, k != 0);
                umma_commit(s_bar);
            }
            __syncwarp();
            mbar_wait(p_bar, j & 1);
            tc_fence_after();
            if (elect_one()) {
                constexpr uint32_t idO = umma_idesc_bf16(128, 64) | (1u << 16);   // B (= V) is MN-major
#pragma unroll
                for (int ks = 0; ks < 16; ++ks) {
                    const uint64_t pd = umma_desc_sw128(smem_u32(sP + (ks >> 2) * 16384)) + 2 * (ks & 3);
                    const uint64_t vd = umma_desc_mn_sw128(smem_u32(sV + buf * 32768 + ks * 2048), 1024);
                    umma_bf16(tmem_o, pd, vd, idO, ks != 0);
                }
                umma_commit(o_bar);
            }
            __syncwarp();
        }
    } else {
        const int q = warp & 3;                       // TMEM lane quadrant this warp may access
        const int r = q * 32 + lane;                  // query row inside the block
        const uint32_t lane_off = (uint32_t)(q * 32) << 16;
        float m_run = -INFINITY, l_run = 0.f;
        float o[64];
#pragma unroll
        for (int i = 0; i < 64; ++i) o[i] = 0.f;
        for (int j = 0; j < nkb; ++j) {
            mbar_wait(s_bar, j & 1);
            tc_fence_after();
            float mx = m_run;
#pragma unroll 1
            for (int c = 0; c < 8; ++c) {
                uint32_t v[32];
                tmem_ld32(tmem + lane_off + c * 32, v);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(v[i]));
            }
            const float mb = mx * scale_log2;
            const float alpha = exp2f(m_run * scale_log2 - mb);   // 0 for the first block (m_run = -inf)
            float sum = 0.f;
#pragma unroll 1
            for (int c = 0; c < 8; ++c) {
                uint32_t v[32];
                tmem_ld32(tmem + lane_off + c * 32, v);
                tmem_ld_wait();
                float pv[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) { pv[i] = exp2f(fmaf(__uint_as_float(v[i]), scale_log2, -mb)); sum += pv[i]; }
                const uint32_t base = smem_u32(sP + (c >> 1) * 16384) + (uint32_t)r * 128;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint4 pk = pack8(pv + k * 8);
                    const uint32_t addr = base + (uint32_t)((((c & 1) * 4 + k) ^ (r & 7)) << 4);
                    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(pk.x), "r"(pk.y), "r"(pk.z), "r"(pk.w) : "memory");
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            tc_fence_before();
            mbar_arrive(p_bar);
            m_run = mx;
            l_run = fmaf(l_run, alpha, sum);
            mbar_wait(o_bar, j & 1);
            tc_fence_after();
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                uint32_t v[32];
                tmem_ld32(tmem_o + lane_off + c * 32, v);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) o[c * 32 + i] = fmaf(o[c * 32 + i], alpha, __uint_as_float(v[i]));
            }
            tc_fence_before();   // the next P_j+1 arrival orders these reads before the next O accumulation
        }
        const float inv = 1.0f / l_run;
        if (lse) lse[((size_t)b * gridDim.y + head) * N + qb * 128 + r] = m_run * scale_log2 + log2f(l_run);
        bf16* orow = out + ((size_t)b * N + qb * 128 + r) * C + head * 64;
#pragma unroll
        for (int i = 0; i < 64; ++i) o[i] *= inv;
#pragma unroll
        for (int k = 0; k < 8; ++k) *reinterpret_cast<uint4*>(orow + k * 8) = pack8(o + k * 8);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

}  // namespace rfv
