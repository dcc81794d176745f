// Hand-crafted orientation and affine shape: the estimators the reference uses when no OriNet /
// AffNet is given.  Replaces OrientationDetector.forward (HandCraftedModules.py:168-192) and AffineShapeEstimator.forward
// (HandCraftedModules.py:94-132).  One warp per patch; the patch (PS x PS, PS <= 41) is staged in shared memory.
#include <math.h>

#include "common.cuh"

namespace ag {

constexpr int HC_MAXPS = 41, HC_WARPS = 2;

__device__ __forceinline__ float warp_sum_f(float v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// gradient with replicate padding; wx = (w_minus, w_plus) of the (1x3) cross-correlation
__device__ __forceinline__ void grad_at(const float* p, int PS, int i, int j, float wm, float wp, float& gx, float& gy) {
    const int jm = max(j - 1, 0), jp = min(j + 1, PS - 1), im = max(i - 1, 0), ip = min(i + 1, PS - 1);
    gx = __fadd_rn(__fmul_rn(wm, p[i * PS + jm]), __fmul_rn(wp, p[i * PS + jp]));
    gy = __fadd_rn(__fmul_rn(wm, p[im * PS + j]), __fmul_rn(wp, p[ip * PS + j]));
}

__global__ void __launch_bounds__(HC_WARPS * 32) orientation_hist_kernel(const float* __restrict__ patches, int n, int PS, const float* __restrict__ gk,
                                                                        float* __restrict__ angle) {
    __shared__ float s_p[HC_WARPS][HC_MAXPS * HC_MAXPS];
    __shared__ float s_w[HC_WARPS][HC_MAXPS * HC_MAXPS];
    __shared__ unsigned char s_b[HC_WARPS][HC_MAXPS * HC_MAXPS];
    __shared__ float s_h[HC_WARPS][40];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int pi = blockIdx.x * HC_WARPS + warp;
    if (pi >= n) return;
    const int NP = PS * PS;
    float* p = s_p[warp];
    for (int i = lane; i < NP; i += 32) p[i] = patches[(size_t)pi * NP + i];
    __syncwarp();
    const float PI_F = 3.14159265358979323846f;
    for (int i = lane; i < NP; i += 32) {
        float gx, gy;
        grad_at(p, PS, i / PS, i % PS, 0.5f, -0.5f, gx, gy);   // weights (0.5, 0, -0.5): 0.5*x[j-1] - 0.5*x[j+1]
        const float mag = __fmul_rn(__fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(gx, gx), __fmul_rn(gy, gy)), 1e-10f)), gk[i]);
        const float ori = atan2f(gy, gx);
        const float o_big = __fdiv_rn(__fmul_rn(36.0f, __fadd_rn(ori, PI_F)), 2.0f * PI_F);
        float bo0 = floorf(o_big);
        const float wo1 = __fsub_rn(o_big, bo0);
        bo0 = fmodf(bo0, 36.0f);
        s_b[warp][i] = (unsigned char)(int)bo0;
        s_w[warp][i] = __fmul_rn(__fsub_rn(1.0f, wo1), mag);   // only the lower-bin weight is accumulated (as the reference does)
    }
    __syncwarp();
    // deterministic histogram: lane b sums bin b (and b+32) over all pixels in raster order
    for (int b = lane; b < 36; b += 32) {
        float acc = 0.f;
        for (int i = 0; i < NP; i++)
            if (s_b[warp][i] == b) acc += s_w[warp][i];
        s_h[warp][b + 1] = acc / (float)NP;   // adaptive_avg_pool2d -> mean
    }
    if (lane == 0) { s_h[warp][0] = 0.f; s_h[warp][37] = 0.f; }   // conv1d zero padding
    __syncwarp();
    float best = -INFINITY;
    int bidx = 0;
    for (int b = lane; b < 36; b += 32) {
        const float v = fmaf(0.33f, s_h[warp][b + 2], fmaf(0.34f, s_h[warp][b + 1], 0.33f * s_h[warp][b]));
        if (v > best) { best = v; bidx = b; }
    }
    // first maximum wins (torch.max on CPU)
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
        if (ov > best || (ov == best && oi < bidx)) { best = ov; bidx = oi; }
    }
    if (lane == 0) angle[pi] = -__fsub_rn(__fdiv_rn(__fmul_rn(2.0f * PI_F, (float)bidx), 36.0f), PI_F);
}

__global__ void __launch_bounds__(HC_WARPS * 32) baumberg_kernel(const float* __restrict__ patches, int n, int PS, const float* __restrict__ gk,
                                                                float* __restrict__ A) {
    __shared__ float s_p[HC_WARPS][HC_MAXPS * HC_MAXPS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int pi = blockIdx.x * HC_WARPS + warp;
    if (pi >= n) return;
    const int NP = PS * PS;
    float* p = s_p[warp];
    for (int i = lane; i < NP; i += 32) p[i] = patches[(size_t)pi * NP + i];
    __syncwarp();
    float a = 0.f, b = 0.f, c = 0.f;
    for (int i = lane; i < NP; i += 32) {
        float gx, gy;
        grad_at(p, PS, i / PS, i % PS, -1.0f, 1.0f, gx, gy);   // weights (-1, 0, 1)
        const float g = gk[i];
        a += gx * gx * g; b += gx * gy * g; c += gy * gy * g;
    }
    a = warp_sum_f(a) / (float)NP; b = warp_sum_f(b) / (float)NP; c = warp_sum_f(c) / (float)NP;
    if (lane != 0) return;
    // invSqrt (HandCraftedModules.py:94-117)
    const float eps = 1e-12f;
    const float mask = (b != 0.f) ? 1.f : 0.f;
    const float r1 = mask * (c - a) / (2.f * b + eps);
    const float sgn = (r1 > 0.f) ? 1.f : ((r1 < 0.f) ? -1.f : 0.f);
    const float t1 = sgn / (fabsf(r1) + sqrtf(1.f + r1 * r1));
    float r = 1.0f / sqrtf(1.f + t1 * t1);
    float t = t1 * r;
    r = r * mask + 1.0f * (1.0f - mask);
    t = t * mask;
    float x = 1.f / sqrtf(r * r * a - 2.0f * r * t * b + t * t * c);
    float z = 1.f / sqrtf(t * t * a + 2.0f * r * t * b + r * r * c);
    const float d = sqrtf(x * z);
    x = x / d; z = z / d;
    const float na = r * r * x + t * t * z, nb = -r * t * x + t * r * z, nc = t * t * x + r * r * z;
    // abc2A + rectifyAffineTransformationUpIsUp (LAF.py:285-291)
    const float a00 = na, a01 = nb, a10 = nb, a11 = nc;
    const float det = sqrtf(fabsf(a00 * a11 - a10 * a01 + 1e-10f));
    const float b2a2 = sqrtf(a01 * a01 + a00 * a00);
    float* o = A + (size_t)pi * 4;
    o[0] = b2a2 / det; o[1] = 0.f;
    o[2] = (a11 * a01 + a10 * a00) / (b2a2 * det); o[3] = det / b2a2;
}

}  // namespace ag

using namespace ag;

extern "C" {

int ag_circular_gauss_kernel(int kernlen, double sigma, float* h_out) {
    AG_REQUIRE(h_out && kernlen >= 1, "bad arguments");
    // CircularGaussKernel(kernlen, sigma, circ_zeros=False, norm=True) under python3 (Utils.py:92-114)
    const double half = kernlen / 2.0, r2 = half * half;
    const double sigma2 = (sigma > 0.0) ? 2.0 * sigma * sigma : 0.9 * r2;
    const double step = (kernlen > 1) ? (2.0 * half) / (kernlen - 1) : 0.0;
    double sum = 0.0;
    for (int i = 0; i < kernlen; i++)
        for (int j = 0; j < kernlen; j++) {
            const double y = (i == kernlen - 1) ? half : -half + i * step, x = (j == kernlen - 1) ? half : -half + j * step;
            sum += exp(-(x * x + y * y) / sigma2);
        }
    for (int i = 0; i < kernlen; i++)
        for (int j = 0; j < kernlen; j++) {
            const double y = (i == kernlen - 1) ? half : -half + i * step, x = (j == kernlen - 1) ? half : -half + j * step;
            h_out[i * kernlen + j] = (float)(exp(-(x * x + y * y) / sigma2) / sum);
        }
    return AG_OK;
}

int ag_orientation_hist(const float* d_patches, int n, int PS, const float* d_gk, float* d_angle, void* stream) {
    AG_REQUIRE(d_patches && d_gk && d_angle, "NULL argument");
    AG_REQUIRE(PS >= 3 && PS <= HC_MAXPS, "patch size out of range (3..41)");
    if (n <= 0) return AG_OK;
    orientation_hist_kernel<<<cdiv(n, HC_WARPS), HC_WARPS * 32, 0, (cudaStream_t)stream>>>(d_patches, n, PS, d_gk, d_angle);
    AG_CHECK_LAUNCH("orientation_hist_kernel");
    return AG_OK;
}

int ag_baumberg_shape(const float* d_patches, int n, int PS, const float* d_gk, float* d_A, void* stream) {
    AG_REQUIRE(d_patches && d_gk && d_A, "NULL argument");
    AG_REQUIRE(PS >= 3 && PS <= HC_MAXPS, "patch size out of range (3..41)");
    if (n <= 0) return AG_OK;
    baumberg_kernel<<<cdiv(n, HC_WARPS), HC_WARPS * 32, 0, (cudaStream_t)stream>>>(d_patches, n, PS, d_gk, d_A);
    AG_CHECK_LAUNCH("baumberg_kernel");
    return AG_OK;
}

}  // extern "C"
