// Stand-alone check of wgrad_umma_kernel (MN-major tcgen05 operands) against a CPU double-precision reference.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I rectified_flow_vision_b200/csrc -o tools/micro/wgrad_test tools/micro/wgrad_test.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>
#include "wgrad.cuh"
using namespace rfv;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode;

static void make_map4(CUtensorMap* m, const bf16* base, int C, int Wd, int Hd, int Nd, size_t sW, size_t sH, size_t sN, int bw, int bh) {
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)Wd, (cuuint64_t)Hd, (cuuint64_t)Nd};
    cuuint64_t strides[3] = {sW * 2, sH * 2, sN * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)bw, (cuuint32_t)bh, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
}

static float bf(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

static int run_up(int Wl, int Hl, int Cin, int Cout, int B, int n128) {
    const int Wh = 2 * Wl, Hh = 2 * Hl, K = 9 * Cin;
    std::vector<float> x((size_t)B * Hl * Wl * Cin), dy((size_t)B * Hh * Wh * Cout);
    srand(4321 + Wl + Cin);
    for (auto& v : x) v = bf((rand() % 2001 - 1000) / 1000.0f);
    for (auto& v : dy) v = bf((rand() % 2001 - 1000) / 1000.0f);
    std::vector<bf16> xb(x.size()), dyb(dy.size());
    for (size_t i = 0; i < x.size(); ++i) xb[i] = __float2bfloat16_rn(x[i]);
    for (size_t i = 0; i < dy.size(); ++i) dyb[i] = __float2bfloat16_rn(dy[i]);
    bf16 *dx, *ddy; float* dW;
    cudaMalloc(&dx, xb.size() * 2); cudaMalloc(&ddy, dyb.size() * 2); cudaMalloc(&dW, (size_t)Cout * K * 4);
    cudaMemcpy(dx, xb.data(), xb.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(ddy, dyb.data(), dyb.size() * 2, cudaMemcpyHostToDevice);
    cudaMemset(dW, 0, (size_t)Cout * K * 4);
    WgradGeom g;
    if (!make_wgrad_geom(&g, Wl, Hl, Cin, Cout, 3, K, 0, n128)) { printf("geom failed\n"); return 1; }
    g.num_tiles = B * g.tiles_per_img;
    CUtensorMap ma, my[4];
    make_map4(&ma, dx, Cin, Wl, Hl, B, Cin, (size_t)Wl * Cin, (size_t)Hl * Wl * Cin, g.pitch, g.R + 2);
    for (int py = 0; py < 2; ++py)
        for (int px = 0; px < 2; ++px)
            make_map4(&my[py * 2 + px], ddy + ((size_t)py * Wh + px) * Cout, Cout, Wl, Hl, B, (size_t)2 * Cout, (size_t)2 * Wh * Cout,
                      (size_t)Hh * Wh * Cout, g.pitch, g.R);
    const size_t smem = wgrad_smem_bytes(g);
    cudaFuncSetAttribute(wgrad_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const long long total = (long long)g.nvar * g.cchA * g.cchB * g.num_tiles;
    const int grid = (int)std::min<long long>(total, 148);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    wgrad_umma_kernel<<<grid, WG_THREADS, smem>>>(ma, ma, ma, ma, my[0], my[1], my[2], my[3], dW, g);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); return 1; }
    std::vector<float> out((size_t)Cout * K);
    cudaMemcpy(out.data(), dW, out.size() * 4, cudaMemcpyDeviceToHost);
    std::vector<double> ref((size_t)Cout * K, 0.0);
    for (int n = 0; n < B; ++n)
        for (int oh = 0; oh < Hh; ++oh)
            for (int ow = 0; ow < Wh; ++ow) {
                const float* dyp = &dy[(((size_t)n * Hh + oh) * Wh + ow) * Cout];
                for (int tap = 0; tap < 9; ++tap) {
                    const int uh = oh + tap / 3 - 1, uw = ow + tap % 3 - 1;   // position in the upsampled image
                    if (uh < 0 || uh >= Hh || uw < 0 || uw >= Wh) continue;
                    const float* xp = &x[(((size_t)n * Hl + uh / 2) * Wl + uw / 2) * Cin];
                    for (int co = 0; co < Cout; ++co) {
                        const double d = dyp[co];
                        double* r = &ref[(size_t)co * K + tap * Cin];
                        for (int ci = 0; ci < Cin; ++ci) r[ci] += d * xp[ci];
                    }
                }
            }
    double num = 0, den = 0;
    for (size_t i = 0; i < ref.size(); ++i) { const double d = out[i] - ref[i]; num += d * d; den += ref[i] * ref[i]; }
    const double rel = std::sqrt(num / std::max(den, 1e-30));
    printf("kind=3 N=%d up %dx%d->%dx%d Cin=%d Cout=%d B=%d R=%d: rel-L2 %.3e  %.3f ms  %s\n", 64 * g.ncob, Wl, Hl, Wh, Hh, Cin, Cout, B, g.R, rel, ms,
           rel < 1e-3 ? "OK" : "MISMATCH");
    if (rel >= 1e-3)
        for (int tap = 0; tap < 9; ++tap) {
            double n2 = 0, d2 = 0;
            for (int co = 0; co < Cout; ++co)
                for (int ci = 0; ci < Cin; ++ci) { const size_t i = (size_t)co * K + tap * Cin + ci; n2 += (out[i] - ref[i]) * (out[i] - ref[i]); d2 += ref[i] * ref[i]; }
            printf("   tap %d rel %.3e   out[0]=%.4f ref[0]=%.4f\n", tap, std::sqrt(n2 / d2), out[tap * Cin], ref[tap * Cin]);
        }
    cudaFree(dx); cudaFree(ddy); cudaFree(dW);
    return rel < 1e-3 ? 0 : 1;
}

// kind 0: 3x3 s1, 1: 1x1, 2: 3x3 s2, 3: nearest-x2 upsample + 3x3 (Wo,Ho = LOW resolution).  Wo,Ho = output grid otherwise.
static int run(int kind, int Wo, int Ho, int Cin, int Cout, int B, int n128 = 1) {
    if (kind == 3) return run_up(Wo, Ho, Cin, Cout, B, n128);
    const int Wi = kind == 2 ? 2 * Wo : Wo, Hi = kind == 2 ? 2 * Ho : Ho;
    const int taps = kind == 1 ? 1 : 9;
    const int K = taps * Cin;
    std::vector<float> x((size_t)B * Hi * Wi * Cin), dy((size_t)B * Ho * Wo * Cout);
    srand(1234 + kind + Wo + Cin);
    for (auto& v : x) v = bf((rand() % 2001 - 1000) / 1000.0f);
    for (auto& v : dy) v = bf((rand() % 2001 - 1000) / 1000.0f);
    std::vector<bf16> xb(x.size()), dyb(dy.size());
    for (size_t i = 0; i < x.size(); ++i) xb[i] = __float2bfloat16_rn(x[i]);
    for (size_t i = 0; i < dy.size(); ++i) dyb[i] = __float2bfloat16_rn(dy[i]);
    bf16 *dx, *ddy; float* dW;
    cudaMalloc(&dx, xb.size() * 2); cudaMalloc(&ddy, dyb.size() * 2); cudaMalloc(&dW, (size_t)Cout * K * 4);
    cudaMemcpy(dx, xb.data(), xb.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(ddy, dyb.data(), dyb.size() * 2, cudaMemcpyHostToDevice);
    cudaMemset(dW, 0, (size_t)Cout * K * 4);
    WgradGeom g;
    if (!make_wgrad_geom(&g, Wo, Ho, Cin, Cout, kind, K, 0, n128)) { printf("geom failed\n"); return 1; }
    g.num_tiles = B * g.tiles_per_img;
    CUtensorMap ma[4], my;
    if (kind != 2) {
        make_map4(&ma[0], dx, Cin, Wi, Hi, B, Cin, (size_t)Wi * Cin, (size_t)Hi * Wi * Cin, g.pitch, g.R + 2);
        ma[1] = ma[2] = ma[3] = ma[0];
    } else {
        for (int ph = 0; ph < 2; ++ph)
            for (int pw = 0; pw < 2; ++pw)
                make_map4(&ma[ph * 2 + pw], dx + ((size_t)ph * Wi + pw) * Cin, Cin, Wo, Ho, B, (size_t)2 * Cin, (size_t)2 * Wi * Cin,
                          (size_t)Hi * Wi * Cin, g.pitch, g.R + 2);
    }
    make_map4(&my, ddy, Cout, Wo, Ho, B, Cout, (size_t)Wo * Cout, (size_t)Ho * Wo * Cout, g.pitch, g.R);
    const size_t smem = wgrad_smem_bytes(g);
    cudaFuncSetAttribute(wgrad_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const long long total = (long long)g.nvar * g.cchA * g.cchB * g.num_tiles;
    const int grid = (int)std::min<long long>(total, 148);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    wgrad_umma_kernel<<<grid, WG_THREADS, smem>>>(ma[0], ma[1], ma[2], ma[3], my, my, my, my, dW, g);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); return 1; }
    std::vector<float> out((size_t)Cout * K);
    cudaMemcpy(out.data(), dW, out.size() * 4, cudaMemcpyDeviceToHost);
    // reference
    std::vector<double> ref((size_t)Cout * K, 0.0);
    const int st = kind == 2 ? 2 : 1;
    for (int n = 0; n < B; ++n)
        for (int oh = 0; oh < Ho; ++oh)
            for (int ow = 0; ow < Wo; ++ow) {
                const float* dyp = &dy[(((size_t)n * Ho + oh) * Wo + ow) * Cout];
                for (int tap = 0; tap < taps; ++tap) {
                    const int ky = kind == 1 ? 1 : tap / 3, kx = kind == 1 ? 1 : tap % 3;
                    const int ih = oh * st + ky - 1, iw = ow * st + kx - 1;
                    if (ih < 0 || ih >= Hi || iw < 0 || iw >= Wi) continue;
                    const float* xp = &x[(((size_t)n * Hi + ih) * Wi + iw) * Cin];
                    for (int co = 0; co < Cout; ++co) {
                        const double d = dyp[co];
                        double* r = &ref[(size_t)co * K + tap * Cin];
                        for (int ci = 0; ci < Cin; ++ci) r[ci] += d * xp[ci];
                    }
                }
            }
    double num = 0, den = 0, maxd = 0;
    for (size_t i = 0; i < ref.size(); ++i) { const double d = out[i] - ref[i]; num += d * d; den += ref[i] * ref[i]; maxd = std::max(maxd, std::fabs(d)); }
    const double rel = std::sqrt(num / std::max(den, 1e-30));
    const double fl = 2.0 * B * Ho * Wo * (double)Cout * K;
    printf("kind=%d N=%d %dx%d Cin=%d Cout=%d B=%d R=%d ksteps=%d stages=%d grid=%d: rel-L2 %.3e max|d| %.3e  %.3f ms %.1f TFLOP/s  %s\n", kind, 64 * g.ncob, Wo, Ho,
           Cin, Cout, B, g.R, g.ksteps, g.stages, grid, rel, maxd, ms, fl / ms / 1e9, rel < 1e-3 ? "OK" : "MISMATCH");
    if (rel >= 1e-3) {
        // diagnose: per-tap relative error for co=0..1
        for (int tap = 0; tap < taps; ++tap) {
            double n2 = 0, d2 = 0;
            for (int co = 0; co < Cout; ++co)
                for (int ci = 0; ci < Cin; ++ci) { const size_t i = (size_t)co * K + tap * Cin + ci; n2 += (out[i] - ref[i]) * (out[i] - ref[i]); d2 += ref[i] * ref[i]; }
            printf("   tap %d rel %.3e   out[0]=%.4f ref[0]=%.4f\n", tap, std::sqrt(n2 / d2), out[tap * Cin], ref[tap * Cin]);
        }
    }
    cudaFree(dx); cudaFree(ddy); cudaFree(dW);
    return rel < 1e-3 ? 0 : 1;
}

int main() {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaFree(0);
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    g_encode = (EncodeTiledFn)fn;
    int bad = 0;
    bad += run(1, 16, 16, 64, 64, 2);
    bad += run(0, 16, 16, 64, 64, 2);
    bad += run(0, 32, 32, 128, 64, 3);
    bad += run(0, 64, 64, 64, 128, 2);
    bad += run(2, 16, 16, 64, 64, 2);
    bad += run(1, 16, 16, 256, 768, 4);
    bad += run(0, 8, 8, 64, 64, 5);
    bad += run(0, 64, 64, 64, 64, 64);    // timing-sized
    bad += run(0, 16, 16, 256, 256, 64, 0);
    bad += run(0, 16, 16, 256, 256, 64, 1);
    bad += run(0, 32, 32, 128, 128, 64, 0);
    bad += run(0, 32, 32, 128, 128, 64, 1);
    bad += run(0, 64, 64, 128, 128, 32, 0);
    bad += run(0, 64, 64, 128, 128, 32, 1);
    bad += run(2, 16, 16, 128, 128, 8, 1);
    bad += run(3, 16, 16, 64, 64, 3, 0);
    bad += run(3, 16, 16, 256, 256, 8, 1);
    bad += run(3, 32, 32, 128, 128, 16, 1);
    bad += run(1, 16, 16, 256, 768, 64, 0);
    bad += run(1, 16, 16, 256, 768, 64, 1);
    printf(bad ? "FAILED %d\n" : "ALL OK\n", bad);
    return bad;
}
