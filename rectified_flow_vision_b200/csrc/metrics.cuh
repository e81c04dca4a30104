// GPU-side quality metrics on the sampler's output tensors: the pixel-space Frechet statistics of
// utils/metrics.py:73-116 (MetricsCalculator.compute_fid_statistics / compute_fid) and the structural similarity of
// utils/metrics.py:39-53 (skimage.metrics.structural_similarity with its defaults: 7x7 uniform window, K1 = 0.01, K2 = 0.03,
// sample covariance, mean over the image cropped by 3 pixels, channels averaged).  Inputs are the fp32 NCHW tensors the
// sampler produces, resident in HBM -- no round trip through host numpy arrays.
//
//   * mean:        mu[j] = mean_i x[i][j]                                  one pass over x, HBM-bound (4 B per element)
//   * covariance:  sigma = (x - mu)^T (x - mu) / (n - 1)   [d x d] fp64     what np.cov(rowvar=False) returns
//   * cross Gram:  m = (x1 - mu1)(x2 - mu2)^T / sqrt((n1-1)(n2-1))  [n1 x n2]:  the non-zero eigenvalues of sigma1 sigma2
//     are the squared singular values of m (sigma_i = a_i^T a_i with a_i the centred, scaled sample matrices, so
//     sigma1 sigma2 = a1^T (a1 a2^T) a2 has the spectrum of (a1 a2^T)(a1 a2^T)^T), hence tr sqrtm(sigma1 sigma2) = sum of the
//     singular values of m -- an n1 x n2 problem instead of a d x d matrix square root (d = 12,288 at 64x64).
//   Both products run as one tiled fp32-FMA kernel (64 x 64 tile per CTA, 4 x 4 per thread) whose partial sums move to fp64
//   every 256 k-steps: the reference computes them in float64, and tensor-core input rounding (bf16 / tf32) would not do.
//   * SSIM: one CTA per (row strip, channel, image); the two planes' strip (+3 halo rows) is staged in shared memory and each
//     interior pixel accumulates its five 7x7 moments in fp64.
#pragma once
#include "common.cuh"

namespace rfv {

// column sums of x [n][d] (fp32) into acc [d] (fp64, zeroed by the caller): grid (ceil(d / 256), row slices)
__global__ void __launch_bounds__(256) metrics_colsum_kernel(const float* __restrict__ x, int n, int d, double* __restrict__ acc) {
    const int j = blockIdx.x * 256 + threadIdx.x;
    if (j >= d) return;
    const int rows = (n + gridDim.y - 1) / gridDim.y, i0 = blockIdx.y * rows, i1 = min(n, i0 + rows);
    double s = 0.0;
    for (int i = i0; i < i1; ++i) s += (double)x[(size_t)i * d + j];
    if (i1 > i0) atomicAdd(acc + j, s);
}
__global__ void metrics_scale_kernel(double* __restrict__ v, int n, double s) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] *= s;
}
// acc[0] += scale * sum_{i,j} (x[i][j] - mu[j])^2   (trace of the covariance when scale = 1 / (n - 1))
__global__ void __launch_bounds__(256) metrics_sqdev_kernel(const float* __restrict__ x, const double* __restrict__ mu, int n, int d,
                                                            double scale, double* __restrict__ acc) {
    __shared__ double red[8];
    const size_t total = (size_t)n * d;
    double s = 0.0;
    for (size_t e = (size_t)blockIdx.x * 256 + threadIdx.x; e < total; e += (size_t)gridDim.x * 256) {
        const double v = (double)x[e] - mu[e % d];
        s += v * v;
    }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += red[w];
        atomicAdd(acc, t * scale);
    }
}
__global__ void metrics_sqdiff_kernel(const double* __restrict__ a, const double* __restrict__ b, int n, double* __restrict__ acc) {
    double s = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) s += (a[i] - b[i]) * (a[i] - b[i]);
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(acc, s);
}

// C[i][j] = scale * sum_k (A(i,k) - ma) (B(j,k) - mb), fp64 out.
//   FEATURES = true  (covariance):  i, j = feature, k = sample:  A(i,k) = xa[k][i] - ma[i]   (M = Nn = d, K = n, ld = d)
//   FEATURES = false (cross Gram):  i, j = sample,  k = feature: A(i,k) = xa[i][k] - ma[k]   (M = n1, Nn = n2, K = d, ld = d)
constexpr int MG_T = 64, MG_K = 16;
template <bool FEATURES>
__global__ void __launch_bounds__(256) metrics_gram_kernel(const float* __restrict__ xa, const float* __restrict__ xb,
                                                           const double* __restrict__ ma, const double* __restrict__ mb, int M, int Nn,
                                                           int K, int ld, double scale, double* __restrict__ C) {
    __shared__ float As[MG_K][MG_T + 4], Bs[MG_K][MG_T + 4];
    const int i0 = blockIdx.y * MG_T, j0 = blockIdx.x * MG_T;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    float acc[4][4];
    double tot[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) { acc[a][b] = 0.f; tot[a][b] = 0.0; }
    int since = 0;
    for (int k0 = 0; k0 < K; k0 += MG_K) {
        // stage the centred operands: 16 x 64 each, four elements per thread
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int idx = threadIdx.x + e * 256;
            int kk, mm;
            if (FEATURES) { kk = idx >> 6; mm = idx & 63; }     // consecutive threads = consecutive features (contiguous)
            else { mm = idx >> 4; kk = idx & 15; }              // consecutive threads = consecutive features of one sample
            const int k = k0 + kk;
            float va = 0.f, vb = 0.f;
            if (k < K) {
                const int ia = i0 + mm, jb = j0 + mm;
                if (ia < M) va = FEATURES ? (float)((double)xa[(size_t)k * ld + ia] - ma[ia]) : (float)((double)xa[(size_t)ia * ld + k] - ma[k]);
                if (jb < Nn) vb = FEATURES ? (float)((double)xb[(size_t)k * ld + jb] - mb[jb]) : (float)((double)xb[(size_t)jb * ld + k] - mb[k]);
            }
            As[kk][mm] = va;
            Bs[kk][mm] = vb;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < MG_K; ++kk) {
            float a[4], b[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) { a[q] = As[kk][ty * 4 + q]; b[q] = Bs[kk][tx * 4 + q]; }
#pragma unroll
            for (int p = 0; p < 4; ++p)
#pragma unroll
                for (int q = 0; q < 4; ++q) acc[p][q] = fmaf(a[p], b[q], acc[p][q]);
        }
        __syncthreads();
        if (++since == 16) {   // 256 k-steps of fp32 partial sums, then into the fp64 totals
            since = 0;
#pragma unroll
            for (int p = 0; p < 4; ++p)
#pragma unroll
                for (int q = 0; q < 4; ++q) { tot[p][q] += (double)acc[p][q]; acc[p][q] = 0.f; }
        }
    }
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int i = i0 + ty * 4 + p, j = j0 + tx * 4 + q;
            if (i < M && j < Nn) C[(size_t)i * Nn + j] = (tot[p][q] + (double)acc[p][q]) * scale;
        }
}

// mean SSIM of image pairs: a, b [B][C][H][W] fp32, out [B] fp64 (zeroed by the caller).  grid (strips, C, B), strip = 8 rows.
constexpr int SS_ROWS = 8, SS_WIN = 7, SS_PAD = 3;
__global__ void __launch_bounds__(256) metrics_ssim_kernel(const float* __restrict__ a, const float* __restrict__ b, int C, int H, int W,
                                                           double c1, double c2, double inv_count, double* __restrict__ out) {
    extern __shared__ float ssm[];   // [2][SS_ROWS + 6][W]
    __shared__ double red[8];
    const int strip = blockIdx.x, ch = blockIdx.y, n = blockIdx.z;
    const int y0 = SS_PAD + strip * SS_ROWS;                  // first output row of this strip
    const int rows_out = min(SS_ROWS, H - SS_PAD - y0);
    const int rows_in = rows_out + 2 * SS_PAD;
    float* pa = ssm;
    float* pb = ssm + (SS_ROWS + 2 * SS_PAD) * W;
    const size_t plane = ((size_t)n * C + ch) * H * W + (size_t)(y0 - SS_PAD) * W;
    for (int e = threadIdx.x; e < rows_in * W; e += 256) { pa[e] = a[plane + e]; pb[e] = b[plane + e]; }
    __syncthreads();
    const double cov_norm = (double)(SS_WIN * SS_WIN) / (double)(SS_WIN * SS_WIN - 1);   // sample covariance (skimage default)
    const double inv_np = 1.0 / (double)(SS_WIN * SS_WIN);
    const int wout = W - 2 * SS_PAD;
    double s = 0.0;
    for (int e = threadIdx.x; e < rows_out * wout; e += 256) {
        const int r = e / wout, x = e - r * wout;             // window rows r .. r+6 of the strip, columns x .. x+6
        double sa = 0, sb = 0, saa = 0, sbb = 0, sab = 0;
        for (int dy = 0; dy < SS_WIN; ++dy) {
            const float* ra = pa + (r + dy) * W + x;
            const float* rb = pb + (r + dy) * W + x;
#pragma unroll
            for (int dx = 0; dx < SS_WIN; ++dx) {
                const double va = ra[dx], vb = rb[dx];
                sa += va; sb += vb; saa += va * va; sbb += vb * vb; sab += va * vb;
            }
        }
        const double ux = sa * inv_np, uy = sb * inv_np;
        const double vx = cov_norm * (saa * inv_np - ux * ux), vy = cov_norm * (sbb * inv_np - uy * uy);
        const double vxy = cov_norm * (sab * inv_np - ux * uy);
        s += ((2.0 * ux * uy + c1) * (2.0 * vxy + c2)) / ((ux * ux + uy * uy + c1) * (vx + vy + c2));
    }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += red[w];
        atomicAdd(out + n, t * inv_count);
    }
}

}  // namespace rfv
