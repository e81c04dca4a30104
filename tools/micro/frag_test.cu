// Stand-alone check of the register <-> (TMEM lane, column) mapping of tcgen05.ld.16x256b and of the stmatrix.trans
// staging the weights-as-A convolution epilogue (conv_wa.cuh) relies on.
//   TMEM is filled with tcgen05.st.32x32b (thread = lane, register = column: the mapping every kernel here already uses),
//   value(lane, col) = (col + 1) * 2^(lane - 64) (exact in bf16), then read back with 16x256b.x4 and the fragment is written through
//   stmatrix.x4.trans; the host checks (a) each register against the documented m16n8 accumulator layout and
//   (b) that the staged bytes are [pixel][channel] rows.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I rectified_flow_vision_b200/csrc -o tools/micro/frag_test tools/micro/frag_test.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
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
__device__ __forceinline__ void tmem_ld_16x256_x4(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.16x256b.x4.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void stmatrix_x4_trans(uint32_t addr, uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3) {
    asm volatile("stmatrix.sync.aligned.m8n8.x4.trans.shared.b16 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(r0), "r"(r1), "r"(r2), "r"(r3) : "memory");
}

constexpr int PITCH = 80;   // staging row pitch in bytes (32 channels = 64 B + 16 B pad)

__global__ void k(uint32_t* frag_out, uint16_t* stage_out, int col0) {
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(16) uint8_t stage[4][32 * PITCH];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) tmem_alloc(&tmem_slot, 64);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = tmem_slot;
    uint32_t v[32];
    for (int c = 0; c < 32; ++c) v[c] = __float_as_uint(ldexpf((float)(c + 1), warp * 32 + lane - 64));
    tmem_st32(tb + ((uint32_t)(warp * 32) << 16), v);
    for (int c = 0; c < 32; ++c) v[c] = __float_as_uint(ldexpf((float)(32 + c + 1), warp * 32 + lane - 64));
    tmem_st32(tb + ((uint32_t)(warp * 32) << 16) + 32, v);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t r[2][16];
    for (int hh = 0; hh < 2; ++hh) tmem_ld_16x256_x4(tb + ((uint32_t)(warp * 32 + hh * 16) << 16) + col0, r[hh]);
    tmem_ld_wait();
    for (int hh = 0; hh < 2; ++hh)
        for (int i = 0; i < 16; ++i) frag_out[((warp * 32 + lane) * 2 + hh) * 16 + i] = r[hh][i];
    // stage: per 8-column group g one stmatrix.x4.trans: matrices (hh=0: rows +0, +8; hh=1: rows +0, +8) -> [8 px][32 ch]
    for (int g = 0; g < 4; ++g) {
        uint32_t m[4];
        for (int hh = 0; hh < 2; ++hh) {
            m[hh * 2] = pack_bf16x2(__uint_as_float(r[hh][4 * g]), __uint_as_float(r[hh][4 * g + 1]));
            m[hh * 2 + 1] = pack_bf16x2(__uint_as_float(r[hh][4 * g + 2]), __uint_as_float(r[hh][4 * g + 3]));
        }
        const uint32_t a = smem_u32(stage[warp]) + (uint32_t)((8 * g + (lane & 7)) * PITCH + (lane >> 3) * 16);
        stmatrix_x4_trans(a, m[0], m[1], m[2], m[3]);
    }
    __syncwarp();
    for (int i = lane; i < 32 * 32; i += 32)   // [px][32 ch] of this warp
        stage_out[warp * 1024 + i] = *reinterpret_cast<uint16_t*>(stage[warp] + (i / 32) * PITCH + (i % 32) * 2);
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tb, 64);
}

static float bf16_to_f(uint16_t h) { uint32_t u = (uint32_t)h << 16; float f; memcpy(&f, &u, 4); return f; }
#include <cstring>

int main(int argc, char** argv) {
    const int col0 = argc > 1 ? atoi(argv[1]) : 0;
    uint32_t* d_frag; uint16_t* d_stage;
    cudaMalloc(&d_frag, 128 * 2 * 16 * 4);
    cudaMalloc(&d_stage, 4 * 1024 * 2);
    k<<<1, 128>>>(d_frag, d_stage, col0);
    cudaError_t e = cudaDeviceSynchronize();
    printf("run: %s\n", cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    std::vector<uint32_t> f(128 * 2 * 16);
    std::vector<uint16_t> s(4 * 1024);
    cudaMemcpy(f.data(), d_frag, f.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(s.data(), d_stage, s.size() * 2, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int t = 0; t < 128; ++t)
        for (int hh = 0; hh < 2; ++hh)
            for (int i = 0; i < 16; ++i) {
                const int warp = t / 32, lane = t % 32, g = i / 4, j = i % 4;
                const int tl = warp * 32 + hh * 16 + lane / 4 + (j >= 2 ? 8 : 0);
                const int col = col0 + 8 * g + 2 * (lane % 4) + (j & 1);
                float got; memcpy(&got, &f[(t * 2 + hh) * 16 + i], 4);
                const float want = ldexpf((float)(col + 1), tl - 64);
                if (got != want) { if (bad < 10) printf("frag mismatch t=%d hh=%d i=%d got %g want lane %d col %d (%g)\n", t, hh, i, got, tl, col, want); ++bad; }
            }
    printf("16x256b.x4 mapping mismatches: %d\n", bad);
    int bad2 = 0;
    for (int w = 0; w < 4; ++w)
        for (int px = 0; px < 32; ++px)
            for (int ch = 0; ch < 32; ++ch) {
                const float got = bf16_to_f(s[w * 1024 + px * 32 + ch]);
                const float ex = ldexpf((float)(col0 + px + 1), w * 32 + ch - 64);
                if (got != ex) { if (bad2 < 10) printf("stage mismatch w=%d px=%d ch=%d got %g want %g\n", w, px, ch, got, ex); ++bad2; }
            }
    printf("stmatrix.trans staging mismatches: %d\n", bad2);
    return bad || bad2;
}
