// Stand-alone check of the TMA store the halo-conv epilogue relies on.  Finding (B200, driver 580): a tiled TMA STORE whose
// start coordinate is negative raises "illegal instruction" (loads accept it); a box that extends past the upper bound is
// clipped.  So a warp's run of 32 flat padded positions (pitch W+1) is COMPACTED in shared memory -- the one position that
// falls on the shared zero column is dropped -- and leaves as ONE store of 32 or 31 consecutive pixels of the flat per-image
// pixel dimension {C, H*W, N}; positions behind the image are clipped by the upper bound.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I rectified_flow_vision_b200/csrc -o tools/micro/tma_store_test tools/micro/tma_store_test.cu -lcuda
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>
#include "common.cuh"
using namespace rfv;

__device__ __forceinline__ void tma_store_3d(const void* map, const void* smem, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"(map), "r"(smem_u32(smem)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

__global__ void k(const __grid_constant__ CUtensorMap map32, const __grid_constant__ CUtensorMap map31, int W, int H) {
    extern __shared__ uint8_t raw[];
    uint8_t* stage = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    const int lane = threadIdx.x, pitch = W + 1;
    for (int p0 = blockIdx.x * 32; p0 < ((H * pitch + 127) / 128) * 128; p0 += gridDim.x * 32) {
        if (lane == 0) bulk_wait_read0();
        __syncwarp();
        const int rr0 = p0 / pitch, cc0 = p0 - rr0 * pitch;
        const int kz = cc0 == 0 ? 0 : pitch - cc0;          // index of the zero-column position inside the run (>= 32: none)
        const int P = rr0 * W + (cc0 == 0 ? 0 : cc0 - 1);   // first pixel of the run
        const int pos = p0 + lane;
        const int row = lane - (lane > kz ? 1 : 0);
        if (lane != kz)
            for (int j = 0; j < 8; ++j) {
                uint4 q;
                q.x = q.y = q.z = q.w = (uint32_t)pos;
                const uint32_t a = smem_u32(stage) + row * 128 + ((j ^ (row & 7)) << 4);
                asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(q.x), "r"(q.y), "r"(q.z), "r"(q.w) : "memory");
            }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {
            if (kz < 32) tma_store_3d(&map31, stage, 0, P, 0);
            else tma_store_3d(&map32, stage, 0, P, 0);
            bulk_commit();
        }
    }
    if (lane == 0) bulk_wait_all0();
}

int main(int argc, char** argv) {
    const int W = argc > 1 ? atoi(argv[1]) : 32, H = W, C = 64;
    uint32_t* d;
    cudaMalloc(&d, (size_t)(H * W + 64) * C * 2);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    auto encode = (PFN_cuTensorMapEncodeTiled_v12000)fn;
    CUtensorMap m32, m31;
    cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)H * W, 1};
    cuuint64_t strides[2] = {(cuuint64_t)C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t box[3] = {64, 32, 1}, es[3] = {1, 1, 1};
    CUresult r = encode(&m32, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    box[1] = 31;
    CUresult r2 = encode(&m31, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode: %d %d\n", (int)r, (int)r2);
    cudaMemset(d, 0xff, (size_t)(H * W + 64) * C * 2);
    k<<<4, 32, 8192>>>(m32, m31, W, H);
    cudaError_t e = cudaDeviceSynchronize();
    printf("run: %s\n", cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    std::vector<uint32_t> h((size_t)(H * W + 64) * C / 2);
    cudaMemcpy(h.data(), d, h.size() * 4, cudaMemcpyDeviceToHost);
    long bad = 0, spill = 0;
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x)
            for (int c = 0; c < C / 2; ++c)
                if (h[((size_t)y * W + x) * (C / 2) + c] != (uint32_t)(y * (W + 1) + x + 1)) ++bad;
    for (size_t i = (size_t)H * W * C / 2; i < h.size(); ++i)
        if (h[i] != 0xffffffffu) ++spill;
    printf("W=%d: mismatches %ld of %d, words written behind the image %ld\n", W, bad, H * W * C / 2, spill);
    return bad || spill;
}
