// Tensor-core engine of AffNet / OriNet / HardNet (replaces the conv stacks and heads of architectures.py:207-235 / 36-82 and
// HardNet.py:67-101; BatchNorm folded, ReLU fused).  Per net:
//   tc_first2_kernel    sampler + input_norm + conv1 + conv2 (tc_first.cuh), patches and layer-1 activations stay on the SM
//   tc_conv_kernel      conv3 .. conv6 as shifted-window implicit GEMMs (tc_conv.cuh)
//   tc_conv_pair_kernel HardNet conv5 / conv6 on CTA pairs, cta_group::2 (tc_pair.cuh)
//   tc_head_kernel / tc_headx_kernel   the 8x8 heads as GEMMs over 128-patch tiles (tc_head.cuh)
// Numerics: fp16 operands with fp16 residual planes where a net needs them, fp32 accumulation in TMEM.
#include <vector>

#include "net_impl.cuh"
#include "tc_conv.cuh"
#include "tc_first.cuh"
#include "tc_head.cuh"
#include "tc_pair.cuh"
#include <stdlib.h>
#include <string.h>

namespace ag {
namespace tc {

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

template <int CIN, int COUT, int H, int STRIDE, int NSPLIT, int STAGES, int OUT, int SA = 0, int SW = 0, int OSA = 0, int FIRST = 0>
static int launch_tc(const __half* in, void* out, const __half* w, const float* b, float inv_scale, int n, int group, const int* count, cudaStream_t st,
                     const FirstSrc* src = nullptr) {
    using Cfg = ConvCfg<CIN, COUT, H, STRIDE, NSPLIT, STAGES, OUT, SA, SW, OSA, FIRST>;
    auto kern = tc_conv_kernel<CIN, COUT, H, STRIDE, NSPLIT, STAGES, OUT, SA, SW, OSA, FIRST>;
    static SmemAttrOnce attr_once;
    {
        int rc = attr_once.ensure(kern, Cfg::SMEM, "tc_conv smem attr");
        if (rc != AG_OK) return rc;
    }
    ConvArgs a;
    a.in = in; a.out = out; a.wpk = w; a.bias = b; a.inv_scale = inv_scale; a.prof_id = 0; a.n = n; a.group = group; a.count = count;
    a.prof_id = (H == 32) ? (FIRST ? 0 : 2) : (H == 16 ? (STRIDE == 1 ? 3 : 4) : 5);
    int gx = num_sms() / NSPLIT;
    if (gx > n) gx = n;
    if (gx < 1) gx = 1;
    FirstSrc fs;
    if (src) fs = *src; else memset(&fs, 0, sizeof(fs));
    kern<<<dim3(gx, NSPLIT), Cfg::THREADS, Cfg::SMEM, st>>>(a, fs);
    AG_CHECK_LAUNCH(FIRST ? "tc_conv_kernel<first>" : "tc_conv_kernel");
    return AG_OK;
}

template <int C1, int COUT, int SA, int SW, int OSA>
static int launch_first2(void* out, const __half* w, const float* b, float inv_scale, int n, int group, const int* count, cudaStream_t st, const FirstSrc& src) {
    using Cfg = FirstCfg<C1, COUT, SA, SW, OSA>;
    auto kern = tc_first2_kernel<C1, COUT, SA, SW, OSA>;
    static SmemAttrOnce attr_once;
    {
        int rc = attr_once.ensure(kern, Cfg::SMEM, "tc_first2 smem attr");
        if (rc != AG_OK) return rc;
    }
    ConvArgs a;
    a.in = nullptr; a.out = out; a.wpk = w; a.bias = b; a.inv_scale = inv_scale; a.prof_id = 0; a.n = n; a.group = group; a.count = count;
    int gx = num_sms();
    if (gx > n) gx = n;
    if (gx < 1) gx = 1;
    kern<<<gx, 448, Cfg::SMEM, st>>>(a, src);
    AG_CHECK_LAUNCH("tc_first2_kernel");
    return AG_OK;
}

// cta_group::2 launch: clusters of two CTAs, one patch per CTA (tc_pair.cuh)
template <int CIN, int COUT, int H, int STRIDE, int STAGES, int OUT>
static int launch_pair(const __half* in, void* out, const __half* w, const float* b, float inv_scale, int n, int group, const int* count, cudaStream_t st) {
    using Cfg = PairCfg<CIN, COUT, H, STRIDE, STAGES, OUT>;
    auto kern = tc_conv_pair_kernel<CIN, COUT, H, STRIDE, STAGES, OUT>;
    static SmemAttrOnce attr_once;
    {
        int rc = attr_once.ensure(kern, Cfg::SMEM, "tc_conv_pair smem attr");
        if (rc != AG_OK) return rc;
    }
    ConvArgs a;
    a.in = in; a.out = out; a.wpk = w; a.bias = b; a.inv_scale = inv_scale; a.prof_id = 0; a.n = n; a.group = group; a.count = count;
    int pairs = num_sms() / 2;
    if (pairs > (n + 1) / 2) pairs = (n + 1) / 2;
    if (pairs < 1) pairs = 1;
    kern<<<2 * pairs, 224, Cfg::SMEM, st>>>(a);
    AG_CHECK_LAUNCH("tc_conv_pair_kernel");
    return AG_OK;
}

static bool no_pair() {
    static const bool v = getenv("AG_NO_PAIR") != nullptr;   // A/B switch: wide HardNet layers as two independent COUT halves
    return v;
}

static bool first_simt() {
    static const bool v = getenv("AG_FIRST_SIMT") != nullptr;   // A/B switch: layer 1 on CUDA cores inside the layer-2 kernel
    return v;
}

}  // namespace tc

tc::FirstSrc tc_src_patches(const float* patches) {
    tc::FirstSrc s;
    memset(&s, 0, sizeof(s));
    s.patches = patches;
    s.cap = 1;
    return s;
}

tc::FirstSrc tc_src_pyramid(const ag_pyramid_plan_t* p, const float* pyr, const float* lafs, const int* oct, const int* lvl, int cap) {
    tc::FirstSrc s;
    memset(&s, 0, sizeof(s));
    s.pyr = pyr; s.lafs = lafs; s.oct = oct; s.lvl = lvl; s.cap = cap;
    s.geom.n_octaves = p->n_octaves; s.geom.n_levels = p->n_levels;
    for (int o = 0; o < AG_MAX_OCTAVES; o++) {
        s.geom.h[o] = p->h[o]; s.geom.w[o] = p->w[o];
        for (int l = 0; l < AG_MAX_LEVELS; l++) s.geom.off[o][l] = p->level_offset[o][l];
    }
    return s;
}

// bytes per patch of each of the two ping-pong fp16 activation buffers (largest layer output)
size_t tc_act_bytes(int kind) {
    using namespace tc;
    if (kind == AG_NET_HARDNET) return ConvCfg<32, 32, 32, 1, 1, 2, PHASE>::OUT_BYTES;        // 82,944 B (L2 out)
    if (kind == AG_NET_ORINET) return ConvCfg<16, 16, 32, 1, 1, 2, PHASE, 1, 1, 1>::OUT_BYTES;  // hi+lo planes
    return ConvCfg<16, 16, 32, 1, 1, 2, PHASE, 0, 1, 0>::OUT_BYTES;
}

// HardNet: trunk (fp16 operands) + tensor-core head -> L2-normalised descriptors [n,128] in `out`.
// `headbuf` holds the last layer's output in the HEADL layout for ceil(n/128)*128 patches (16 KiB each).
int tc_hardnet_forward(const ag_net* net, const tc::FirstSrc& src0, int n, int group, const int* count, void* bufA, void* bufB, void* headbuf,
                       float* out, cudaStream_t st) {
    using namespace tc;
    __half* A = (__half*)bufA;
    __half* B = (__half*)bufB;
    FirstSrc src = src0;
    src.w1 = net->d_w1; src.b1 = net->d_b[0]; src.w1_inv = net->w_inv_scale[0]; src.w1_scale = 1.0f / net->w_inv_scale[0];
    int rc;
    if (first_simt()) rc = launch_tc<32, 32, 32, 1, 1, 2, PHASE, 0, 0, 0, 1>(nullptr, B, net->d_wh[1], net->d_b[1], net->w_inv_scale[1], n, group, count, st, &src);
    else rc = launch_first2<32, 32, 0, 0, 0>(B, net->d_wh[1], net->d_b[1], net->w_inv_scale[1], n, group, count, st, src);
    if (rc) return rc;
    if ((rc = launch_tc<32, 64, 32, 2, 1, 2, PLAIN>(B, A, net->d_wh[2], net->d_b[2], net->w_inv_scale[2], n, group, count, st))) return rc;
    if ((rc = launch_tc<64, 64, 16, 1, 1, 2, PHASE>(A, B, net->d_wh[3], net->d_b[3], net->w_inv_scale[3], n, group, count, st))) return rc;
    if (no_pair()) {
        if ((rc = launch_tc<64, 128, 16, 2, 2, 2, PLAIN>(B, A, net->d_wh[4], net->d_b[4], net->w_inv_scale[4], n, group, count, st))) return rc;
        if ((rc = launch_tc<128, 128, 8, 1, 2, 2, HEADL>(A, headbuf, net->d_wh[5], net->d_b[5], net->w_inv_scale[5], n, group, count, st))) return rc;
    } else {   // two SMs per MMA: every patch is read and multiplied once, each SM holds half of the weights
        if ((rc = launch_pair<64, 128, 16, 2, 3, PLAIN>(B, A, net->d_wh[4], net->d_b[4], net->w_inv_scale[4], n, group, count, st))) return rc;
        if ((rc = launch_pair<128, 128, 8, 1, 2, HEADL>(A, headbuf, net->d_wh[5], net->d_b[5], net->w_inv_scale[5], n, group, count, st))) return rc;
    }
    return tc_hardnet_head(net, headbuf, n, group, count, out, st);
}

// HardNet 8x8 head GEMM + BatchNorm + L2 norm over the head operand a trunk left in `headbuf`
int tc_hardnet_head(const ag_net* net, const void* headbuf, int n, int group, const int* count, float* out, cudaStream_t st, int bf16) {
    using namespace tc;
    static SmemAttrOnce once0, once1;
    {
        int rc = once0.ensure(tc_head_kernel<0>, HEAD_SMEM, "tc_head smem attr");
        if (rc == AG_OK) rc = once1.ensure(tc_head_kernel<1>, HEAD_SMEM, "tc_head smem attr");
        if (rc != AG_OK) return rc;
    }
    if (bf16) tc_head_kernel<1><<<(n + 127) / 128, 192, HEAD_SMEM, st>>>((const __half*)headbuf, net->d_headh_bf, net->d_head_b, out, n, group, count);
    else tc_head_kernel<0><<<(n + 127) / 128, 192, HEAD_SMEM, st>>>((const __half*)headbuf, net->d_headh, net->d_head_b, out, n, group, count);
    AG_CHECK_LAUNCH("tc_head_kernel");
    return AG_OK;
}

// AffNet trunk -> features as fp16 hi + lo planes in the head-GEMM layout (tc_head.cuh).  Weights split hi/lo (A error 1.8e-4; plain fp16 weights give 1.8e-3).
int tc_trunk_affnet(const ag_net* net, const tc::FirstSrc& src0, int n, int group, const int* count, void* bufA, void* bufB, void* feat,
                    cudaStream_t st) {
    using namespace tc;
    __half* A = (__half*)bufA;
    __half* B = (__half*)bufB;
    FirstSrc src = src0;
    src.w1 = net->d_w1; src.b1 = net->d_b[0]; src.w1_inv = net->w_inv_scale[0]; src.w1_scale = 1.0f / net->w_inv_scale[0];
    int rc;
    if (first_simt()) rc = launch_tc<16, 16, 32, 1, 1, 2, PHASE, 0, 1, 0, 1>(nullptr, B, net->d_wh[1], net->d_b[1], net->w_inv_scale[1], n, group, count, st, &src);
    else rc = launch_first2<16, 16, 0, 1, 0>(B, net->d_wh[1], net->d_b[1], net->w_inv_scale[1], n, group, count, st, src);
    if (rc) return rc;
    if ((rc = launch_tc<16, 32, 32, 2, 1, 4, PLAIN, 0, 1, 0>(B, A, net->d_wh[2], net->d_b[2], net->w_inv_scale[2], n, group, count, st))) return rc;
    if ((rc = launch_tc<32, 32, 16, 1, 1, 6, PHASE, 0, 1, 0>(A, B, net->d_wh[3], net->d_b[3], net->w_inv_scale[3], n, group, count, st))) return rc;
    if ((rc = launch_tc<32, 64, 16, 2, 1, 5, PLAIN, 0, 1, 0>(B, A, net->d_wh[4], net->d_b[4], net->w_inv_scale[4], n, group, count, st))) return rc;
    if ((rc = launch_tc<64, 64, 8, 1, 1, 4, HEADL, 0, 1, 1>(A, feat, net->d_wh[5], net->d_b[5], net->w_inv_scale[5], n, group, count, st))) return rc;
    return AG_OK;
}

// OriNet trunk (and AffNet under the exact engine) -> features as fp16 hi + lo planes in the head-GEMM layout, or fp32 [n,64,8,8].  The angle is ill-conditioned in the features (fp16 activations give 8e-3 rad),
// so both operands are split: three MMAs per K step, fp32-grade result (1.6e-5 rad in emulation).
int tc_trunk_orinet(const ag_net* net, const tc::FirstSrc& src0, int n, int group, const int* count, void* bufA, void* bufB, void* feat,
                    cudaStream_t st) {
    using namespace tc;
    __half* A = (__half*)bufA;
    __half* B = (__half*)bufB;
    FirstSrc src = src0;
    src.w1 = net->d_w1; src.b1 = net->d_b[0]; src.w1_inv = net->w_inv_scale[0]; src.w1_scale = 1.0f / net->w_inv_scale[0];
    int rc;
    if (first_simt()) rc = launch_tc<16, 16, 32, 1, 1, 2, PHASE, 1, 1, 1, 1>(nullptr, B, net->d_wh[1], net->d_b[1], net->w_inv_scale[1], n, group, count, st, &src);
    else rc = launch_first2<16, 16, 1, 1, 1>(B, net->d_wh[1], net->d_b[1], net->w_inv_scale[1], n, group, count, st, src);
    if (rc) return rc;
    if ((rc = launch_tc<16, 32, 32, 2, 1, 2, PLAIN, 1, 1, 1>(B, A, net->d_wh[2], net->d_b[2], net->w_inv_scale[2], n, group, count, st))) return rc;
    if ((rc = launch_tc<32, 32, 16, 1, 1, 3, PHASE, 1, 1, 1>(A, B, net->d_wh[3], net->d_b[3], net->w_inv_scale[3], n, group, count, st))) return rc;
    if ((rc = launch_tc<32, 64, 16, 2, 1, 2, PLAIN, 1, 1, 1>(B, A, net->d_wh[4], net->d_b[4], net->w_inv_scale[4], n, group, count, st))) return rc;
    // default engine: hi/lo head-GEMM operand; exact engine: fp32 NCHW features for the fp32 FMA-chain heads of nets_simt.cu
    if (net->engine == AG_ENGINE_TC_EXACT) return launch_tc<64, 64, 8, 1, 1, 2, FINAL, 1, 1, 0>(A, feat, net->d_wh[5], net->d_b[5], net->w_inv_scale[5], n, group, count, st);
    return launch_tc<64, 64, 8, 1, 1, 2, HEADL, 1, 1, 1>(A, feat, net->d_wh[5], net->d_b[5], net->w_inv_scale[5], n, group, count, st);
}

// AffNet / OriNet head on tensor cores over the hi/lo feature planes the trunks above leave in `feat`
int tc_headx_forward(const ag_net* net, const void* feat, int n, int group, const int* count, float* out, float* angle, cudaStream_t st, float* raw) {
    using namespace tc;
    static SmemAttrOnce once0, once1;
    {
        int rc = once0.ensure(tc_headx_kernel<0>, HX_SMEM, "tc_headx smem attr");
        if (rc == AG_OK) rc = once1.ensure(tc_headx_kernel<1>, HX_SMEM, "tc_headx smem attr");
        if (rc != AG_OK) return rc;
    }
    const int tiles = (n + 127) / 128;
    if (net->kind == AG_NET_AFFNET) tc_headx_kernel<0><<<tiles, 192, HX_SMEM, st>>>((const __half*)feat, net->d_headh, net->d_head_b, net->head_inv_scale, out, nullptr, raw, n, group, count);
    else tc_headx_kernel<1><<<tiles, 192, HX_SMEM, st>>>((const __half*)feat, net->d_headh, net->d_head_b, net->head_inv_scale, out, angle, raw, n, group, count);
    AG_CHECK_LAUNCH("tc_headx_kernel");
    return AG_OK;
}

// bytes of the head-GEMM operand of n patches (hi + lo planes, padded to whole 128-patch tiles)
size_t tc_headx_bytes(int n) { return (size_t)((n + 127) / 128) * 2 * tc::HX_PLANE_TILE; }

int tc_nsplit(int kind, int layer) { return (kind == AG_NET_HARDNET && layer >= 4) ? 2 : 1; }
int tc_split_w(int kind) { return kind == AG_NET_HARDNET ? 0 : 1; }

}  // namespace ag

#ifdef AG_ROLE_PROF
// developer-only: per-CTA role cycle counters of the last tc_first2_kernel launch (tc_first.cuh)
extern "C" int ag_debug_role_prof(unsigned long long* out) {
    if (cudaMemcpyFromSymbol(out, ag::tc::g_role_prof, sizeof(unsigned long long) * 8 * 160 * 20) != cudaSuccess) return 1;
    void* p = nullptr;
    return (cudaGetSymbolAddress(&p, ag::tc::g_role_prof) == cudaSuccess && cudaMemset(p, 0, sizeof(unsigned long long) * 8 * 160 * 20) == cudaSuccess) ? 0 : 1;
}
#endif
