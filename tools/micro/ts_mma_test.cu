// Stand-alone check and timing of tcgen05.mma with the A operand in TENSOR MEMORY (TS mode) against the SS mode every kernel
// here uses.  Question: would a resident 128-row weight block in TMEM (written once with tcgen05.st.32x32b: thread = row = TMEM
// lane, 32-bit column j = K elements 2j, 2j+1 -- that layout is confirmed: results are exact) make the M128 x N x K16 MMAs of
// the weights-as-A kernel faster?
//   D[128][N] = A[128][64] . B[N][64]^T, bf16 in, fp32 out; SS and TS both checked against the CPU, then timed.
// Measured on B200 (cycles per MMA, elected-lane issue): N = 160: SS 83.8 / TS 80.0;  N = 192: 96 / 96;  N = 256: 128 / 128.
// So TS mode only removes the ~82-cycle A-fetch floor below N = 164; at the N = 256 shape the kernels use it buys nothing, and
// the TMEM it would cost (32 columns per 64-channel weight block) is the accumulator double buffer.  Two further findings:
//   * putting BOTH forms into one loop body behind a run-time switch (the inactive one predicated off) made every MMA ~40 cycles
//     slower (168.5 instead of 128 at N = 256) -- a predicated-off UTCHMMA is not free; the conv kernels have none;
//   * a tcgen05.commit per MMA group, taken or predicated off, costs nothing (the stage-free / accumulator-full signals of the
//     conv kernels are free); alternating between two accumulators changes nothing either (no accumulate-dependency bubble).
//   usage: ts_mma_test [const] [148 CTAs] [384 threads] [225 KB smem] [switches: 1 no TMEM store, 2 no check, 4 / 8 commits]
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
__global__ void __launch_bounds__(384, 1) k(const __nv_bfloat16* A, const __nv_bfloat16* B, float* Dss, float* Dts, int N, int reps,
                                            long long* cyc, int quiet, int skip) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;
    uint8_t* sB = smem + 16384;
    __shared__ uint32_t tmem_slot;
    __shared__ uint64_t bar, bar2;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // fill the swizzled operand tiles: element (r, k) of a K-major tile sits at r*128 + ((k/8) ^ (r%8))*16 + (k%8)*2
    for (int i = tid; i < 128 * 8; i += blockDim.x) {
        const int r = i >> 3, c = i & 7;
        *reinterpret_cast<uint4*>(sA + r * 128 + ((c ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(A + r * 64 + c * 8);
    }
    for (int i = tid; i < N * 8; i += blockDim.x) {
        const int r = i >> 3, c = i & 7;
        *reinterpret_cast<uint4*>(sB + r * 128 + ((c ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(B + r * 64 + c * 8);
    }
    if (tid == 0) { mbar_init(&bar, 1); mbar_init(&bar2, 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc(&tmem_slot, 512);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = tmem_slot;
    const uint32_t d_ss = tb, d_ts = tb + 256 - 0, a_tm = tb + 256;   // D(ss) cols 0.., A cols 256..287, D(ts) reuses cols 0.. after readback
    (void)d_ts;
    // A into TMEM: thread = row r = lane (warp*32 + lane), 32 columns = the row's 64 bf16
    if (tid < 128 && !(skip & 1)) {
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
    for (int mode = 0; mode < ((skip & 2) ? 0 : 2); ++mode) {
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
        if (tid < 128) for (int c = 0; c < N; c += 32) {
            uint32_t v[32];
            tmem_ld32(tb + ((uint32_t)(warp * 32) << 16) + c, v);
            tmem_ld_wait();
            if (blockIdx.x == 0) for (int j = 0; j < 32 && c + j < N; ++j) out[(size_t)tid * N + c + j] = __uint_as_float(v[j]);
        }
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
    }
    // instruction rate: `reps` groups of 4 MMAs back to back.  modes 0 / 1: SS / TS accumulating into ONE tile (what a conv
    // tile does: every MMA depends on the previous one's accumulator); modes 2 / 3: the same alternating between TWO tiles
    // (second accumulator at column 288; only when N <= 192 so that it fits behind the A block)
    for (int mode = 0; mode < 4; ++mode) {
        long long t0 = 0;
        const bool two = mode >= 2;
        if (two && N > 192) { if (tid == 0 && blockIdx.x == 0) cyc[mode] = 0; continue; }
        // issued the way the kernels do: the whole (converged) warp walks the loop, one ELECTED lane issues.  (A first version
        // issued from `if (tid == 0)` inside a diverged warp: every MMA then cost ~40 cycles more in SS mode, 168.5 instead of
        // 128 at N = 256 -- the elected-lane form is what reaches the documented rate.)
        if (warp == 0) {
            t0 = clock64();
            // (separate loops per form: with both forms predicated inside one loop body the SASS issues the inactive form
            // predicated-off, and every MMA measured ~40 cycles slower -- 168.5 instead of 128 at N = 256)
            if ((mode & 1) == 0) {
                for (int i = 0; i < reps; ++i) {
                    if (elect_one()) {
                        const uint32_t d = (two && (i & 1)) ? tb + 288 : d_ss;
#pragma unroll
                        for (int j = 0; j < 4; ++j) umma_bf16(d, adesc + 2 * j, bdesc + 2 * j, idesc, 1);
                        // skip & 4: a commit that is never taken (what `if (last k block) commit(accumulator full)` is on all
                        // but one iteration of a conv kernel's loop) -- does a predicated-off UTCBAR cost tensor-queue time?
                        if ((skip & 4) && i == reps + 5) umma_commit(&bar);
                        if ((skip & 8)) umma_commit(&bar2);   // skip & 8: a commit that IS taken every group (stage-free signal)
                    }
                    __syncwarp();
                }
            } else {
                for (int i = 0; i < reps; ++i) {
                    if (elect_one()) {
                        const uint32_t d = (two && (i & 1)) ? tb + 288 : d_ss;
#pragma unroll
                        for (int j = 0; j < 4; ++j) umma_bf16_ts(d, a_tm + 8 * j, bdesc + 2 * j, idesc, 1);
                    }
                    __syncwarp();
                }
            }
            if (elect_one()) umma_commit(&bar);
            __syncwarp();
        }
        // quiet = 1: only the issuing thread polls the mbarrier, everyone else parks on the hardware barrier (does polling
        // by other warps slow the tensor core's shared-memory operand fetch?)
        if (!quiet || tid == 0) mbar_wait(&bar, parity);
        parity ^= 1;
        if (tid == 0 && blockIdx.x == 0) cyc[mode] = clock64() - t0;
        __syncthreads();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tb, 512);
}

int main(int argc, char** argv) {
    const int Ns[3] = {160, 192, 256};
    std::vector<__nv_bfloat16> hA(128 * 64), hB(NMAX * 64);
    srand(1);
    const bool constant = argc > 1;   // any argument: constant operands (what tools/micro/umma_rate2.cu measured with)
    for (auto& v : hA) v = __float2bfloat16(constant ? 1.0f : (rand() % 17 - 8) / 8.0f);
    for (auto& v : hB) v = __float2bfloat16(constant ? 1.0f : (rand() % 13 - 6) / 4.0f);
    printf("%s operands\n", constant ? "constant" : "random");
    __nv_bfloat16 *dA, *dB;
    float *dS, *dT;
    long long* dc;
    cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2);
    cudaMalloc(&dS, 128 * NMAX * 4); cudaMalloc(&dT, 128 * NMAX * 4); cudaMalloc(&dc, 32);
    cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024);
    for (int N : Ns) {
        const int reps = 2000;
      for (int quiet = 0; quiet < 1; ++quiet) {
        k<<<(argc > 2 ? 148 : 1), (argc > 3 ? 384 : 128), (argc > 4 ? 225 : 50) * 1024>>>(dA, dB, dS, dT, N, reps, dc, quiet, argc > 5 ? atoi(argv[5]) : 0);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("N=%d: CUDA error %s\n", N, cudaGetErrorString(e)); return 1; }
        std::vector<float> S(128 * N), T(128 * N);
        long long c[4];
        cudaMemcpy(S.data(), dS, S.size() * 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(T.data(), dT, T.size() * 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(c, dc, 32, cudaMemcpyDeviceToHost);
        double es = 0, et = 0;
        for (int r = 0; r < 128; ++r)
            for (int n = 0; n < N; ++n) {
                double ref = 0;
                for (int kk = 0; kk < 64; ++kk) ref += (double)__bfloat162float(hA[r * 64 + kk]) * (double)__bfloat162float(hB[n * 64 + kk]);
                es = fmax(es, fabs(S[r * N + n] - ref));
                et = fmax(et, fabs(T[r * N + n] - ref));
            }
        printf("N=%d: max|err| SS %.3g  TS %.3g   cycles per MMA, one accumulator: SS %.1f  TS %.1f;  two accumulators alternating per group: SS %.1f  TS %.1f\n",
               N, es, et, c[0] / (4.0 * reps), c[1] / (4.0 * reps), c[2] / (4.0 * reps), c[3] / (4.0 * reps));
        printf("      (other warps %s)\n", quiet ? "parked on bar.sync" : "polling the mbarrier");
      }
    }
    return 0;
}
