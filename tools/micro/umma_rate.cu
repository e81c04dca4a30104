// Microbenchmark: sustained tcgen05.mma (kind::f16, bf16, M=128, cta_group::1, SS mode) rate on one SM per N,
// with operands already resident in 128B-swizzled shared memory (no TMA, no epilogue).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I rectified_flow_vision_b200/csrc -o umma_rate tools/micro/umma_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "common.cuh"
using namespace rfv;

template <int N>
__global__ void __launch_bounds__(128, 1) rate_kernel(long long* out, int iters, int nstages, int a_shift_rows) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < (nstages * (16384 + N * 128) + 4096) / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc(&slot, 512);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = slot;
    if (warp == 1) {
        constexpr uint32_t idesc = umma_idesc_bf16(128, N);
        long long t0 = 0, t1 = 0;
        uint32_t ph = 0;
        for (int rep = 0; rep < 2; ++rep) {   // rep 0 = warm-up
            t0 = clock64();
            for (int it = 0; it < iters; ++it) {
                const int st = it % nstages;
                if (elect_one()) {
                    const uint32_t a = smem_u32(smem + st * 16384) + a_shift_rows * 128;
                    const uint32_t b = smem_u32(smem + nstages * 16384 + st * N * 128);
                    const uint64_t ad = umma_desc_sw128(a), bd = umma_desc_sw128(b);
#pragma unroll
                    for (int j = 0; j < 4; ++j) umma_bf16(tmem + (it & 1) * N * 0, ad + 2 * j, bd + 2 * j, idesc, 1u);
                }
                __syncwarp();
            }
            if (elect_one()) umma_commit(&bar);
            __syncwarp();
            mbar_wait(&bar, ph);
            ph ^= 1;
            t1 = clock64();
        }
        if (threadIdx.x == 32) out[blockIdx.x] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

template <int N>
void run(int blocks, int iters, int nstages, int shift) {
    long long* d;
    cudaMalloc(&d, blocks * sizeof(long long));
    size_t smem = nstages * (16384 + N * 128) + 4096 + 1024;
    cudaFuncSetAttribute(rate_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    rate_kernel<N><<<blocks, 128, smem>>>(d, iters, nstages, shift);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[256];
    cudaMemcpy(h, d, blocks * sizeof(long long), cudaMemcpyDeviceToHost);
    double clk = (double)h[0] / (iters * 4.0);
    printf("N=%3d blocks=%3d stages=%d shift=%d: %.1f clk per MMA(128xNx16) -> %.1f%% of 8192 flop/clk/SM  [%s]\n", N, blocks,
           nstages, shift, clk, 100.0 * (2.0 * 128 * N * 16 / clk) / 8192.0, cudaGetErrorString(e));
    cudaFree(d);
}

int main() {
    for (int blocks : {1, 148}) {
        run<64>(blocks, 2000, 4, 0);
        run<128>(blocks, 2000, 4, 0);
        run<256>(blocks, 2000, 4, 0);
    }
    run<64>(148, 2000, 1, 0);
    run<256>(148, 2000, 1, 0);
    run<64>(148, 2000, 4, 3);   // A start address shifted by 3 rows (128-byte, not 1024-byte aligned)
    run<256>(148, 2000, 4, 3);
    return 0;
}
