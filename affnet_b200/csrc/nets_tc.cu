// Tensor-core trunks of AffNet / OriNet / HardNet: first layer (K = 9, fp32 SIMT, fused with the per-patch input
// normalisation) writing the fp16 canonical layout, then five tcgen05 shifted-window convolutions (tc_conv.cuh).
// Replaces the conv stacks of architectures.py:207-226 / 36-55 and HardNet.py:67-85 (BatchNorm folded, ReLU fused).
// Numerics: fp16 operands, fp32 accumulation in TMEM, first layer and heads in fp32 (SURVEY.md §7 hard part 1).
#include <vector>

#include "net_impl.cuh"
#include "tc_conv.cuh"
#include "tc_head.cuh"

namespace ag {
namespace tc {

// ---- layer 1: input_norm + conv3x3(1 -> C) + ReLU -> fp16 PLAIN(32) ------------------------------------------------
template <int C, int OSA>
__global__ void __launch_bounds__(256) first_layer_kernel(const float* __restrict__ patches, __half* __restrict__ out,
                                                          const float* __restrict__ wpk /*[9][C]*/, const float* __restrict__ bias,
                                                          int group, const int* __restrict__ count) {
    using Lay = InLay<32, 1>;
    __shared__ float s_in[34][35];
    __shared__ float s_w[9][C];
    __shared__ float s_b[C];
    __shared__ float s_red[8][2];
    const int pi = blockIdx.x;
    if (count != nullptr && (pi % group) >= count[pi / group]) return;
    const float* src = patches + (size_t)pi * 1024;
    float v4[4];
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 4; k++) { v4[k] = src[threadIdx.x + k * 256]; s += v4[k]; }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5][0] = s;
    for (int i = threadIdx.x; i < 9 * C; i += 256) s_w[i / C][i % C] = wpk[i];
    if (threadIdx.x < C) s_b[threadIdx.x] = bias[threadIdx.x];
    for (int i = threadIdx.x; i < 34 * 35; i += 256) (&s_in[0][0])[i] = 0.f;
    __syncthreads();
    s = 0.f;
    for (int i = 0; i < 8; i++) s += s_red[i][0];
    const float mean = s / 1024.f;
    float q = 0.f;
#pragma unroll
    for (int k = 0; k < 4; k++) { const float d = v4[k] - mean; q = fmaf(d, d, q); }
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5][1] = q;
    __syncthreads();
    q = 0.f;
    for (int i = 0; i < 8; i++) q += s_red[i][1];
    const float inv = 1.f / (sqrtf(q / 1023.f) + 1e-7f);  // unbiased std + 1e-7 (architectures.py:234)
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int p = threadIdx.x + k * 256;
        s_in[(p >> 5) + 1][(p & 31) + 1] = (v4[k] - mean) * inv;
    }
    __syncthreads();
    unsigned char* outp = reinterpret_cast<unsigned char*>(out) + (size_t)pi * (C / 8) * (1 + OSA) * Lay::NPIX * 16;
    // zero border of the padded plane
    for (int i = threadIdx.x; i < 4 * 33; i += 256) {
        const int side = i / 33, k = i - side * 33;
        int Y, X;
        if (side == 0) { Y = 0; X = k; } else if (side == 1) { Y = 33; X = k + 1; } else if (side == 2) { Y = k + 1; X = 0; } else { Y = k; X = 33; }
#pragma unroll
        for (int g = 0; g < (C / 8) * (1 + OSA); g++) *reinterpret_cast<uint4*>(outp + ((size_t)g * Lay::NPIX + Lay::slot(Y, X)) * 16) = make_uint4(0, 0, 0, 0);
    }
#pragma unroll 1
    for (int k = 0; k < 4; k++) {
        const int p = threadIdx.x + k * 256, y = p >> 5, x = p & 31;
        float acc[C];
#pragma unroll
        for (int c = 0; c < C; c++) acc[c] = s_b[c];
#pragma unroll
        for (int tap = 0; tap < 9; tap++) {
            const float a = s_in[y + tap / 3][x + tap % 3];
#pragma unroll
            for (int c = 0; c < C; c++) acc[c] = fmaf(a, s_w[tap][c], acc[c]);
        }
        const int slot = Lay::slot(y + 1, x + 1);
#pragma unroll
        for (int g = 0; g < C / 8; g++) {
            uint4 pk;
            pk.x = pack_h2(fmaxf(acc[g * 8 + 0], 0.f), fmaxf(acc[g * 8 + 1], 0.f));
            pk.y = pack_h2(fmaxf(acc[g * 8 + 2], 0.f), fmaxf(acc[g * 8 + 3], 0.f));
            pk.z = pack_h2(fmaxf(acc[g * 8 + 4], 0.f), fmaxf(acc[g * 8 + 5], 0.f));
            pk.w = pack_h2(fmaxf(acc[g * 8 + 6], 0.f), fmaxf(acc[g * 8 + 7], 0.f));
            *reinterpret_cast<uint4*>(outp + ((size_t)g * Lay::NPIX + slot) * 16) = pk;
            if (OSA) {
                float l[8];
#pragma unroll
                for (int e = 0; e < 8; e++) { const float v = fmaxf(acc[g * 8 + e], 0.f); l[e] = v - __half2float(__float2half_rn(v)); }
                pk.x = pack_h2(l[0], l[1]); pk.y = pack_h2(l[2], l[3]); pk.z = pack_h2(l[4], l[5]); pk.w = pack_h2(l[6], l[7]);
                *reinterpret_cast<uint4*>(outp + ((size_t)(C / 8 + g) * Lay::NPIX + slot) * 16) = pk;
            }
        }
    }
}

static int g_num_sms = 0;
static int num_sms() {
    if (g_num_sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
        if (g_num_sms <= 0) g_num_sms = 148;
    }
    return g_num_sms;
}

template <int CIN, int COUT, int H, int STRIDE, int NSPLIT, int STAGES, int OUT, int SA = 0, int SW = 0, int OSA = 0>
static int launch_tc(const __half* in, void* out, const __half* w, const float* b, int n, int group, const int* count, cudaStream_t st) {
    using Cfg = ConvCfg<CIN, COUT, H, STRIDE, NSPLIT, STAGES, OUT, SA, SW, OSA>;
    auto kern = tc_conv_kernel<CIN, COUT, H, STRIDE, NSPLIT, STAGES, OUT, SA, SW, OSA>;
    static bool configured = false;
    if (!configured) {
        int rc = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM), "tc_conv smem attr");
        if (rc != AG_OK) return rc;
        configured = true;
    }
    ConvArgs a;
    a.in = in; a.out = out; a.wpk = w; a.bias = b; a.n = n; a.group = group; a.count = count;
    int gx = num_sms() / NSPLIT;
    if (gx > n) gx = n;
    if (gx < 1) gx = 1;
    kern<<<dim3(gx, NSPLIT), 192, Cfg::SMEM, st>>>(a);
    AG_CHECK_LAUNCH("tc_conv_kernel");
    return AG_OK;
}

}  // namespace tc

// bytes per patch of each of the two ping-pong fp16 activation buffers (largest layer output)
size_t tc_act_bytes(int kind) {
    using namespace tc;
    if (kind == AG_NET_HARDNET) return ConvCfg<32, 32, 32, 1, 1, 2, PHASE>::OUT_BYTES;        // 82,944 B (L2 out)
    if (kind == AG_NET_ORINET) return ConvCfg<16, 16, 32, 1, 1, 2, PHASE, 1, 1, 1>::OUT_BYTES;  // hi+lo planes
    return ConvCfg<16, 16, 32, 1, 1, 2, PHASE, 0, 1, 0>::OUT_BYTES;
}

// HardNet: trunk (fp16 operands) + tensor-core head -> L2-normalised descriptors [n,128] in `out`.
// `headbuf` holds the last layer's output in the HEADL layout for ceil(n/128)*128 patches (16 KiB each).
int tc_hardnet_forward(const ag_net* net, const float* patches, int n, int group, const int* count, void* bufA, void* bufB, void* headbuf,
                       float* out, cudaStream_t st) {
    using namespace tc;
    __half* A = (__half*)bufA;
    __half* B = (__half*)bufB;
    first_layer_kernel<32, 0><<<n, 256, 0, st>>>(patches, A, net->d_w1, net->d_b[0], group, count);
    AG_CHECK_LAUNCH("first_layer_kernel");
    int rc;
    if ((rc = launch_tc<32, 32, 32, 1, 1, 2, PHASE>(A, B, net->d_wh[1], net->d_b[1], n, group, count, st))) return rc;
    if ((rc = launch_tc<32, 64, 32, 2, 1, 2, PLAIN>(B, A, net->d_wh[2], net->d_b[2], n, group, count, st))) return rc;
    if ((rc = launch_tc<64, 64, 16, 1, 1, 2, PHASE>(A, B, net->d_wh[3], net->d_b[3], n, group, count, st))) return rc;
    if ((rc = launch_tc<64, 128, 16, 2, 2, 2, PLAIN>(B, A, net->d_wh[4], net->d_b[4], n, group, count, st))) return rc;
    if ((rc = launch_tc<128, 128, 8, 1, 2, 2, HEADL>(A, headbuf, net->d_wh[5], net->d_b[5], n, group, count, st))) return rc;
    static bool configured = false;
    if (!configured) {
        rc = check_cuda(cudaFuncSetAttribute(tc_head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HEAD_SMEM), "tc_head smem attr");
        if (rc != AG_OK) return rc;
        configured = true;
    }
    tc_head_kernel<<<(n + 127) / 128, 192, HEAD_SMEM, st>>>((const __half*)headbuf, net->d_headh, net->d_head_b, out, n, group, count);
    AG_CHECK_LAUNCH("tc_head_kernel");
    return AG_OK;
}

// AffNet trunk -> fp32 features [n,64,8,8].  Weights split hi/lo (A error 1.8e-4; plain fp16 weights give 1.8e-3).
int tc_trunk_affnet(const ag_net* net, const float* patches, int n, int group, const int* count, void* bufA, void* bufB, float* feat,
                    cudaStream_t st) {
    using namespace tc;
    __half* A = (__half*)bufA;
    __half* B = (__half*)bufB;
    first_layer_kernel<16, 0><<<n, 256, 0, st>>>(patches, A, net->d_w1, net->d_b[0], group, count);
    AG_CHECK_LAUNCH("first_layer_kernel");
    int rc;
    if ((rc = launch_tc<16, 16, 32, 1, 1, 2, PHASE, 0, 1, 0>(A, B, net->d_wh[1], net->d_b[1], n, group, count, st))) return rc;
    if ((rc = launch_tc<16, 32, 32, 2, 1, 2, PLAIN, 0, 1, 0>(B, A, net->d_wh[2], net->d_b[2], n, group, count, st))) return rc;
    if ((rc = launch_tc<32, 32, 16, 1, 1, 2, PHASE, 0, 1, 0>(A, B, net->d_wh[3], net->d_b[3], n, group, count, st))) return rc;
    if ((rc = launch_tc<32, 64, 16, 2, 1, 2, PLAIN, 0, 1, 0>(B, A, net->d_wh[4], net->d_b[4], n, group, count, st))) return rc;
    if ((rc = launch_tc<64, 64, 8, 1, 1, 2, FINAL, 0, 1, 0>(A, feat, net->d_wh[5], net->d_b[5], n, group, count, st))) return rc;
    return AG_OK;
}

// OriNet trunk -> fp32 features [n,64,8,8].  The angle is ill-conditioned in the features (fp16 activations give 8e-3 rad),
// so both operands are split: three MMAs per K step, fp32-grade result (1.6e-5 rad in emulation).
int tc_trunk_orinet(const ag_net* net, const float* patches, int n, int group, const int* count, void* bufA, void* bufB, float* feat,
                    cudaStream_t st) {
    using namespace tc;
    __half* A = (__half*)bufA;
    __half* B = (__half*)bufB;
    first_layer_kernel<16, 1><<<n, 256, 0, st>>>(patches, A, net->d_w1, net->d_b[0], group, count);
    AG_CHECK_LAUNCH("first_layer_kernel");
    int rc;
    if ((rc = launch_tc<16, 16, 32, 1, 1, 2, PHASE, 1, 1, 1>(A, B, net->d_wh[1], net->d_b[1], n, group, count, st))) return rc;
    if ((rc = launch_tc<16, 32, 32, 2, 1, 2, PLAIN, 1, 1, 1>(B, A, net->d_wh[2], net->d_b[2], n, group, count, st))) return rc;
    if ((rc = launch_tc<32, 32, 16, 1, 1, 2, PHASE, 1, 1, 1>(A, B, net->d_wh[3], net->d_b[3], n, group, count, st))) return rc;
    if ((rc = launch_tc<32, 64, 16, 2, 1, 2, PLAIN, 1, 1, 1>(B, A, net->d_wh[4], net->d_b[4], n, group, count, st))) return rc;
    if ((rc = launch_tc<64, 64, 8, 1, 1, 2, FINAL, 1, 1, 0>(A, feat, net->d_wh[5], net->d_b[5], n, group, count, st))) return rc;
    return AG_OK;
}

int tc_nsplit(int kind, int layer) { return (kind == AG_NET_HARDNET && layer >= 4) ? 2 : 1; }
int tc_split_w(int kind) { return kind == AG_NET_HARDNET ? 0 : 1; }

}  // namespace ag
