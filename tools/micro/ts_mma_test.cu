// Stand-alone check and timing of tcgen05.mma with the A operand in TENSOR MEMORY (TS mode), against the SS mode every kernel
// here uses.  Question: can a resident 128-row weight block live in TMEM (written once with tcgen05.st.32x32b: thread = row =
// TMEM lane, 32-bit column j = K elements 2j, 2j+1), so that an M128 x N x K16 MMA no longer re-reads 4 KB of A from shared
// memory -- the ~82-cycle floor that makes N < 256 tiles slow and costs shared-memory bandwidth at N = 256?
//   D[128][N] = A[128][64] . B[N][64]^T, bf16 in, fp32 out; (1) SS reference, (2) TS; both checked against the CPU; then the
//   instruction rate of both forms at N = 160 / 192 / 256.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I rectified_flow_vision_b200/csrc -o tools/micro/ts_mma_test tools/micro/ts_mma_test.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "common.cuh"
using namespace rfv;

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* r) {
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
__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}

constexpr int NMAX = 256;
// smem: A tile [128][64] bf16 (16 KB, 128B-swizzled), B tile [NMAX][64] bf16 (32 KB)
__global__ void __launch_bounds__(128, 1) k(const __nv_bfloat16* A, const __nv_bfloat16* B, float* Dss, float* Dts, int N, int reps,
                                            long long* cyc) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;
    uint8_t* sB = smem + 16384;
    __shared__ uint32_t tmem_slot;
    __shared__ uint64_t bar;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // fill the swizzled operand tiles: element (r, k) of a K-major tile sits at r*128 + ((k/8) ^ (r%8))*16 + (k%8)*2
    for (int i = tid; i < 128 * 8; i += 128) {
        const int r = i >> 3, c = i & 7;
        *reinterpret_cast<uint4*>(sA + r * 128 + ((c ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(A + r * 64 + c * 8);
    }
    for (int i = tid; i < N * 8; i += 128) {
        const int r = i >> 3, c = i & 7;
        *reinterpret_cast<uint4*>(sB + r * 128 + ((c ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(B + r * 64 + c * 8);
    }
    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc(&tmem_slot, 512);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = tmem_slot;
    const uint32_t d_ss = tb, d_ts = tb + 256 - 0, a_tm = tb + 256;   // D(ss) cols 0.., A cols 256..287, D(ts) reuses cols 0.. after readback
    (void)d_ts;
    // A into TMEM: thread = row r = lane (warp*32 + lane), 32 columns = the row's 64 bf16
    {
        uint32_t v[32];
        const uint32_t* row = reinterpret_cast<const uint32_t*>(A + (size_t)tid * 64);
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = row[j];
        tmem_st32(a_tm + ((uint32_t)(warp * 32) << 16), v);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t idesc = umma_idesc_bf16(128, N);
    const uint64_t adesc = umma_desc_sw128(smem_u32(sA)), bdesc = umma_desc_sw128(smem_u32(sB));
    uint32_t parity = 0;
    for (int mode = 0; mode < 2; ++mode) {
        if (tid == 0) {
            for (int j = 0; j < 4; ++j) {
                if (mode == 0) umma_bf16(d_ss, adesc + 2 * j, bdesc + 2 * j, idesc, j != 0);
                else umma_bf16_ts(d_ss, a_tm + 8 * j, bdesc + 2 * j, idesc, j != 0);
            }
            umma_commit(&bar);
        }
        mbar_wait(&bar, parity);
        parity ^= 1;
        tc_fence_after();
        float* out = mode == 0 ? Dss : Dts;
        for (int c = 0; c < N; c += 32) {
            uint32_t v[32];
            tmem_ld32(tb + ((uint32_t)(warp * 32) << 16) + c, v);
            tmem_ld_wait();
            for (int j = 0; j < 32 && c + j < N; ++j) out[(size_t)tid * N + c + j] = __uint_as_float(v[j]);
        }
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
    }
    // instruction rate: `reps` groups of 4 MMAs back to back
    for (int mode = 0; mode < 2; ++mode) {
        long long t0 = 0;
        if (tid == 0) {
            t0 = clock64();
            for (int i = 0; i < reps; ++i)
                for (int j = 0; j < 4; ++j) {
                    if (mode == 0) umma_bf16(d_ss, adesc + 2 * j, bdesc + 2 * j, idesc, 1);
                    else umma_bf16_ts(d_ss, a_tm + 8 * j, bdesc + 2 * j, idesc, 1);
                }
            umma_commit(&bar);
        }
        mbar_wait(&bar, parity);
        parity ^= 1;
        if (tid == 0) cyc[mode] = clock64() - t0;
        __syncthreads();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tb, 512);
}

int main() {
    const int Ns[3] = {160, 192, 256};
    std::vector<__nv_bfloat16> hA(128 * 64), hB(NMAX * 64);
    srand(1);
    for (auto& v : hA) v = __float2bfloat16((rand() % 17 - 8) / 8.0f);
    for (auto& v : hB) v = __float2bfloat16((rand() % 13 - 6) / 4.0f);
    __nv_bfloat16 *dA, *dB;
    float *dS, *dT;
    long long* dc;
    cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2);
    cudaMalloc(&dS, 128 * NMAX * 4); cudaMalloc(&dT, 128 * NMAX * 4); cudaMalloc(&dc, 16);
    cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    for (int N : Ns) {
        const int reps = 2000;
        k<<<1, 128, 50 * 1024>>>(dA, dB, dS, dT, N, reps, dc);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("N=%d: CUDA error %s\n", N, cudaGetErrorString(e)); return 1; }
        std::vector<float> S(128 * N), T(128 * N);
        long long c[2];
        cudaMemcpy(S.data(), dS, S.size() * 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(T.data(), dT, T.size() * 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(c, dc, 16, cudaMemcpyDeviceToHost);
        double es = 0, et = 0;
        for (int r = 0; r < 128; ++r)
            for (int n = 0; n < N; ++n) {
                double ref = 0;
                for (int kk = 0; kk < 64; ++kk) ref += (double)__bfloat162float(hA[r * 64 + kk]) * (double)__bfloat162float(hB[n * 64 + kk]);
                es = fmax(es, fabs(S[r * N + n] - ref));
                et = fmax(et, fabs(T[r * N + n] - ref));
            }
        printf("N=%d: max|err| SS %.3g  TS %.3g   cycles per MMA: SS %.1f  TS %.1f\n", N, es, et, c[0] / (4.0 * reps), c[1] / (4.0 * reps));
    }
    return 0;
}
