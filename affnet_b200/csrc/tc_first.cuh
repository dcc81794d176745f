// First two conv layers of AffNet / OriNet / HardNet in ONE kernel, both on tensor cores (sm_100a):
//
//   sampler (LAF.py:313-372) -> input_norm (architectures.py:231-235) -> conv3x3(1 -> C1)+BN+ReLU -> conv3x3(C1 -> COUT)+BN+ReLU
//
// 32x32 patches and the layer-1 activations never exist in HBM.  Layer 1 (K = 9) is made tensor-core shaped with a sliding
// window plane: P[slot] = 4 consecutive normalised pixels of the slot's row and 4 of the row below (fp16), so one 16-byte
// core-matrix row holds two tap rows (elements 0..2 and 4..6; the weights of elements 3 and 7 are zero) and the third tap row
// is the same plane two pitches further, reached through the descriptor's leading-byte offset: ONE M=128, K=16 MMA per tile
// covers all nine taps (every MMA costs a 4 KB shared-memory read of its A operand, ~64 clk, whatever its N).  For the
// input and weights carry fp16 residual planes, so layer 1 is fp32-grade, as the accuracy
// budget requires (SURVEY.md §7 hard part 1).  (HardNet with the single hi*hi product is 6 % faster in this kernel but doubles
// the descriptor error to 1.1e-3; it keeps the residual terms as well.)
//
// Warp roles (448 threads):  0 weight loader | 1 MMA issuer | 2-5 layer-2 epilogue (TMEM -> global, next layout)
//                            6-9 layer-1 epilogue (TMEM -> bias/ReLU -> fp16 stage in shared memory) | 10-13 sampler + norm + P planes
#pragma once
#include "tc_conv.cuh"

namespace ag {
namespace tc {

template <int C1, int COUT, int SA, int SW, int OSA>
struct FirstCfg {
    using In = InLay<32, 1>;     // layout of the stage (input of layer 2)
    using OutS = InLay<32, 2>;   // layout of the output (input of the stride-2 layer 3)
    static constexpr int KC = C1 / 8, NT = COUT, TILES = In::TILES;
    static constexpr int NPIXP = 1280;                         // slots of a P plane: 9*128 rows + 3 pitches of look-ahead, zero tail
    static constexpr int X1 = 1;                              // layer 1 with residual planes (x = hi + lo, w = hi + lo) for every net
    // stacked-N operands (see ConvCfg::ACCW): layer 2 when SW; layer 1 when its nine accumulators still fit TMEM (C1 = 16)
    static constexpr int S1 = (X1 && TILES * 2 * C1 + 2 * NT * (1 + SW) <= 512) ? 1 : 0;
    static constexpr int ACC1 = C1 * (1 + S1);                 // layer-1 accumulator width
    static constexpr int ACCW = NT * (1 + SW);                 // layer-2 accumulator width
    static constexpr int C1COLS = TILES * ACC1;                // TMEM columns of the layer-1 accumulators (one buffer per tile)
    static constexpr int NACC = ((512 - C1COLS) / ACCW) < 8 ? ((512 - C1COLS) / ACCW) : 8;
    static constexpr uint32_t IN_BYTES = (uint32_t)KC * (1 + SA) * In::NPIX * 16;
    static constexpr uint32_t W_HALF = 9u * KC * NT * 16, W_BYTES = W_HALF * (1 + SW);
    static constexpr uint32_t W1_BYTES = 2u * 2u * C1 * 16;   // [K chunk 0|1][hi rows | lo rows][8]
    static constexpr uint32_t P_BYTES = 2u * NPIXP * 16;
    static constexpr int SX = 1320;                            // floats of one padded fp32 patch buffer (34*34 + zero tail)
    static constexpr size_t SMEM = 1024 + (size_t)W_BYTES + 2 * (size_t)IN_BYTES + P_BYTES + W1_BYTES + 2 * SX * 4 + 256;
    static constexpr int OUT_NPIX = OutS::NPIX;
    static constexpr size_t OUT_BYTES = (size_t)(COUT / 8) * (1 + OSA) * OUT_NPIX * 16;
    static_assert(C1 % 16 == 0 && NT % 16 == 0 && NT <= 128 && NACC >= 2 && ACCW <= 256, "shape");
    static_assert(SMEM <= 232448, "shared memory budget");
    static_assert(TILES * 128 + 2 * In::PITCH <= NPIXP && NPIXP + In::PITCH + 4 <= SX, "P plane look-ahead");
};

template <int C1, int COUT, int SA, int SW, int OSA>
__global__ void __launch_bounds__(448, 1) tc_first2_kernel(const ConvArgs a, const FirstSrc src) {
    using Cfg = FirstCfg<C1, COUT, SA, SW, OSA>;
    using In = typename Cfg::In;
    constexpr int KC = Cfg::KC, NT = Cfg::NT, NACC = Cfg::NACC, TILES = Cfg::TILES, NPIXP = Cfg::NPIXP, SX = Cfg::SX;
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);   // [2]   layer-2 stage filled (128 layer-1 epilogue threads)
    uint64_t* empty = full + 2;                             // [2]   layer-2 MMAs done with the stage
    uint64_t* tfull = empty + 2;                            // [8]
    uint64_t* tempty = tfull + 8;                           // [8]
    uint64_t* wbar = tempty + 8;
    uint64_t* p_full = wbar + 1;                            // P planes written (128 producer threads)
    uint64_t* p_empty = p_full + 1;                         // layer-1 MMAs done with the P planes
    uint64_t* c1_full = p_empty + 1;                        // [9]
    uint64_t* c1_empty = c1_full + 9;                       // [9]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(c1_empty + 9);
    float* s_bias1 = reinterpret_cast<float*>(smem + 384);  // [C1]
    float* s_bias = reinterpret_cast<float*>(smem + 512);   // [NT]
    unsigned char* sW = smem + 1024;
    unsigned char* sIn = sW + Cfg::W_BYTES;
    unsigned char* sP = sIn + 2 * (size_t)Cfg::IN_BYTES;    // [hi|lo][NPIXP][8] fp16
    unsigned char* sW1 = sP + Cfg::P_BYTES;                 // [chunk][hi|lo][C1][8] fp16
    float* s_x = reinterpret_cast<float*>(sW1 + Cfg::W1_BYTES);   // [2][SX]
    float* s_red = s_x + 2 * SX;                            // [2][4][2]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    auto valid = [&](int pi) -> bool { return a.count == nullptr || (pi % a.group) < a.count[pi / a.group]; };
    auto next_valid = [&](int pi) -> int {
        while (pi < a.n && !valid(pi)) pi += gridDim.x;
        return pi;
    };

    // ---- one-time setup by all threads: barriers, biases, layer-1 weight operand, zeroed planes / stages ----
    if (threadIdx.x < NT) s_bias[threadIdx.x] = a.bias[threadIdx.x];
    if (threadIdx.x < C1) s_bias1[threadIdx.x] = src.b1[threadIdx.x];
    if (threadIdx.x == 0) {
        for (int s = 0; s < 2; s++) { mbar_init(&full[s], 128); mbar_init(&empty[s], 1); }
        for (int i = 0; i < 8; i++) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 4); }
        mbar_init(wbar, 1); mbar_init(p_full, 128); mbar_init(p_empty, 1);
        for (int i = 0; i < 9; i++) { mbar_init(&c1_full[i], 1); mbar_init(&c1_empty[i], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < 2 * 2 * C1 * 8; i += blockDim.x) {   // W1[chunk][hi rows | lo rows][e]: chunk 0 = tap rows 0 (e 0..2), 1 (e 4..6); chunk 1 = tap row 2
        const int e = i & 7, co = (i >> 3) % C1, part = (i / (8 * C1)) & 1, ch = i / (8 * C1 * 2);
        const int dy = ch == 0 ? (e >> 2) : 2, dx = e & 3;
        float v = 0.f;
        if (dx < 3 && (ch == 0 || e < 4)) {
            const float wv = src.w1[(dy * 3 + dx) * C1 + co] * src.w1_scale;   // power-of-two scale, undone in the epilogue
            const __half hi = __float2half_rn(wv);
            v = part == 0 ? __half2float(hi) : wv - __half2float(hi);
        }
        reinterpret_cast<__half*>(sW1)[i] = __float2half_rn(v);
    }
    for (int i = threadIdx.x; i < (int)((2 * Cfg::IN_BYTES + Cfg::P_BYTES) / 16); i += blockDim.x) reinterpret_cast<uint4*>(sIn)[i] = make_uint4(0, 0, 0, 0);
    for (int i = threadIdx.x; i < 2 * SX; i += blockDim.x) s_x[i] = 0.f;
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t tmem_l2 = tmem + (uint32_t)Cfg::C1COLS;

    if (warp == 0) {
        if (lane == 0) {
            mbar_expect_tx(wbar, Cfg::W_BYTES);
            bulk_g2s(sW, a.wpk, Cfg::W_BYTES, wbar);
        }
    } else if (warp == 1) {
        // ===== MMA issuer: layer 1 runs one patch ahead of layer 2 =====
        constexpr uint32_t idesc2 = (1u << 4) | ((uint32_t)(Cfg::ACCW >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        constexpr uint32_t idesc2_hi = (1u << 4) | ((uint32_t)(NT >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        constexpr uint32_t idesc1 = (1u << 4) | ((uint32_t)(C1 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        constexpr uint32_t idesc1_st = (1u << 4) | ((uint32_t)((2 * C1) >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t leader = elect_one();
        mbar_wait(wbar, 0);
        tc_fence_after();
        RP_DECL;
        const uint32_t w_lo = desc_lo(smem_u32(sW), Cfg::ACCW * 16u);
        const uint32_t w1_lo = desc_lo(smem_u32(sW1), 2 * C1 * 16u);   // K chunks are 2*C1 rows apart (hi rows, then lo rows)
        const uint32_t p_lo = desc_lo(smem_u32(sP), 2 * In::PITCH * 16u);   // leading-byte offset = two tap rows
        auto issue_conv1 = [&](int n1) {
            RP_WAIT(0, mbar_wait(p_full, n1 & 1));
            tc_fence_after();
#pragma unroll 1
            for (int t = 0; t < TILES; t++) {
                RP_WAIT(1, mbar_wait(&c1_empty[t], (n1 & 1) ^ 1));
                tc_fence_after();
                if (leader) {
                    const uint32_t d = tmem + (uint32_t)(t * Cfg::ACC1);
                    const uint32_t alo = p_lo + (uint32_t)(t * 128);
                    const uint32_t mfix = (In::TAIL64 && t == TILES - 1) ? M64_FIX : 0u;
                    if (Cfg::S1) {   // x_hi * [w_hi ; w_lo] in one MMA, then x_lo * w_hi
                        umma_f16_lo<0>(d, alo, w1_lo, idesc1_st - mfix);
                        umma_f16_lo<1>(d, alo + (uint32_t)NPIXP, w1_lo, idesc1 - mfix);
                    } else {
                        umma_f16_lo<0>(d, alo, w1_lo, idesc1 - mfix);
                        if (Cfg::X1) {
                            umma_f16_lo<1>(d, alo + (uint32_t)NPIXP, w1_lo, idesc1 - mfix);             // x_lo * w_hi
                            umma_f16_lo<1>(d, alo, w1_lo + (uint32_t)C1, idesc1 - mfix);                 // x_hi * w_lo (lo rows follow the hi rows)
                        }
                    }
                    umma_commit(&c1_full[t]);
                }
                __syncwarp();
            }
            if (leader) umma_commit(p_empty);
            __syncwarp();
        };
        int tcnt = 0;
        auto issue_l2 = [&](int it) {
            const int s = it & 1;
            RP_WAIT(2, mbar_wait(&full[s], (it >> 1) & 1));
            tc_fence_after();
            const uint32_t in_lo = desc_lo(smem_u32(sIn + (size_t)s * Cfg::IN_BYTES), In::NPIX * 16u);
#pragma unroll 1
            for (int t = 0; t < TILES; t++, tcnt++) {
                const int ab = tcnt % NACC;
                RP_WAIT(3, mbar_wait(&tempty[ab], ((tcnt / NACC) & 1) ^ 1));
                tc_fence_after();
                if (leader) {
                    const uint32_t d = tmem_l2 + (uint32_t)(ab * Cfg::ACCW);
                    const uint32_t a_t = in_lo + (uint32_t)(t * 128);
                    const uint32_t mfix = (In::TAIL64 && t == TILES - 1) ? M64_FIX : 0u;
#pragma unroll
                    for (int tap = 0; tap < 9; tap++) {
#pragma unroll
                        for (int j = 0; j < KC / 2; j++) {
                            const uint32_t alo = a_t + (uint32_t)(In::tap_off(tap / 3, tap % 3) + 2 * j * In::NPIX);
                            const uint32_t blo = w_lo + (uint32_t)((tap * KC + 2 * j) * Cfg::ACCW);
                            if (tap == 0 && j == 0) umma_f16_lo<0>(d, alo, blo, idesc2 - mfix); else umma_f16_lo<1>(d, alo, blo, idesc2 - mfix);
                            if (SA) umma_f16_lo<1>(d, alo + (uint32_t)(KC * In::NPIX), blo, idesc2_hi - mfix);
                        }
                    }
                    umma_commit(&tfull[ab]);
                }
                __syncwarp();
            }
            if (leader) umma_commit(&empty[s]);
            __syncwarp();
        };
        int pi = next_valid(blockIdx.x);
        if (pi < a.n) issue_conv1(0);
        int it = 0;
        while (pi < a.n) {
            const int pn = next_valid(pi + gridDim.x);
            if (pn < a.n) issue_conv1(it + 1);
            issue_l2(it);
            it++;
            pi = pn;
        }
        RP_STORE(0, 0);
    } else if (warp < 6) {
        // ===== layer-2 epilogue: TMEM -> bias + ReLU -> fp16 -> global (layout of the stride-2 consumer) =====
        const int q = warp & 3, et = (warp - 2) * 32 + lane;
        int tcnt = 0;
        RP_DECL;
        for (int pi = next_valid(blockIdx.x); pi < a.n; pi = next_valid(pi + gridDim.x)) {
            unsigned char* outp = reinterpret_cast<unsigned char*>(a.out) + (size_t)pi * Cfg::OUT_BYTES;
            constexpr int HB = 33;
            for (int i = et; i < 4 * HB; i += 128) {   // zero border of the consumer's padded plane
                const int side = i / HB, k = i - side * HB;
                int Y, X;
                if (side == 0) { Y = 0; X = k; } else if (side == 1) { Y = 33; X = k + 1; } else if (side == 2) { Y = k + 1; X = 0; } else { Y = k; X = 33; }
                const int slot = Cfg::OutS::slot(Y, X);
#pragma unroll
                for (int g = 0; g < (NT / 8) * (1 + OSA); g++) *reinterpret_cast<uint4*>(outp + ((size_t)g * Cfg::OUT_NPIX + slot) * 16) = make_uint4(0, 0, 0, 0);
            }
#pragma unroll 1
            for (int t = 0; t < TILES; t++, tcnt++) {
                const int ab = tcnt % NACC;
                RP_WAIT(0, mbar_wait(&tfull[ab], (tcnt / NACC) & 1));
                tc_fence_after();
                const bool t64 = In::TAIL64 && t == TILES - 1;
                const int m = t * 128 + (t64 ? q * 16 : q * 32) + lane;
                const int y = m / In::PITCH, x = m - y * In::PITCH;
                const bool ok = (y < 32) && (x < 32) && !(t64 && lane >= 16);
                const uint32_t taddr = tmem_l2 + ((uint32_t)(q * 32) << 16) + (uint32_t)(ab * Cfg::ACCW);
                uint32_t r[32];
                if (NT >= 32) {
                    tmem_ld32(taddr, r);
                    if (SW) {
                        uint32_t r2[32];
                        tmem_ld32(taddr + NT, r2);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 32; i++) r[i] = __float_as_uint(__uint_as_float(r[i]) + __uint_as_float(r2[i]));
                    }
                } else {
                    uint32_t r16[16];
                    tmem_ld16(taddr, r16);
#pragma unroll
                    for (int i = 0; i < 16; i++) { r[i] = r16[i]; r[16 + i] = 0; }
                    if (SW) {
                        tmem_ld16(taddr + NT, r16);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 16; i++) r[i] = __float_as_uint(__uint_as_float(r[i]) + __uint_as_float(r16[i]));
                    }
                }
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty[ab]);
                if (ok) {
                    const int slot = Cfg::OutS::slot(y + 1, x + 1);
#pragma unroll
                    for (int g = 0; g < NT / 8; g++) {
                        float v[8];
                        const float4 b0 = *reinterpret_cast<const float4*>(s_bias + g * 8), b1 = *reinterpret_cast<const float4*>(s_bias + g * 8 + 4);
                        v[0] = fmaxf(fmaf(__uint_as_float(r[g * 8 + 0]), a.inv_scale, b0.x), 0.f); v[1] = fmaxf(fmaf(__uint_as_float(r[g * 8 + 1]), a.inv_scale, b0.y), 0.f);
                        v[2] = fmaxf(fmaf(__uint_as_float(r[g * 8 + 2]), a.inv_scale, b0.z), 0.f); v[3] = fmaxf(fmaf(__uint_as_float(r[g * 8 + 3]), a.inv_scale, b0.w), 0.f);
                        v[4] = fmaxf(fmaf(__uint_as_float(r[g * 8 + 4]), a.inv_scale, b1.x), 0.f); v[5] = fmaxf(fmaf(__uint_as_float(r[g * 8 + 5]), a.inv_scale, b1.y), 0.f);
                        v[6] = fmaxf(fmaf(__uint_as_float(r[g * 8 + 6]), a.inv_scale, b1.z), 0.f); v[7] = fmaxf(fmaf(__uint_as_float(r[g * 8 + 7]), a.inv_scale, b1.w), 0.f);
                        uint4 pk;
                        pk.x = pack_h2(v[0], v[1]); pk.y = pack_h2(v[2], v[3]); pk.z = pack_h2(v[4], v[5]); pk.w = pack_h2(v[6], v[7]);
                        *reinterpret_cast<uint4*>(outp + ((size_t)g * Cfg::OUT_NPIX + slot) * 16) = pk;
                        if (OSA) {
                            float l[8];
#pragma unroll
                            for (int e = 0; e < 8; e++) l[e] = v[e] - __half2float(__float2half_rn(v[e]));
                            pk.x = pack_h2(l[0], l[1]); pk.y = pack_h2(l[2], l[3]); pk.z = pack_h2(l[4], l[5]); pk.w = pack_h2(l[6], l[7]);
                            *reinterpret_cast<uint4*>(outp + ((size_t)(NT / 8 + g) * Cfg::OUT_NPIX + slot) * 16) = pk;
                        }
                    }
                }
            }
        }
        RP_STORE(0, 1);
    } else if (warp < 10) {
        // ===== layer-1 epilogue: TMEM -> bias + ReLU -> fp16 (hi [+lo]) -> shared-memory stage of layer 2 =====
        const int q = warp & 3;
        int it = 0;
        RP_DECL;
        for (int pi = next_valid(blockIdx.x); pi < a.n; pi = next_valid(pi + gridDim.x), it++) {
            const int s = it & 1;
            RP_WAIT(0, mbar_wait(&empty[s], ((it >> 1) & 1) ^ 1));
            unsigned char* st = sIn + (size_t)s * Cfg::IN_BYTES;
#pragma unroll 1
            for (int t = 0; t < TILES; t++) {
                RP_WAIT(1, mbar_wait(&c1_full[t], it & 1));
                tc_fence_after();
                const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(t * Cfg::ACC1);
                uint32_t r[32];
                if (C1 >= 32) {
                    tmem_ld32(taddr, r);
                } else if (Cfg::S1) {   // [x*w_hi | x_hi*w_lo] side by side: one 32-column load, add the halves
                    tmem_ld32(taddr, r);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 16; i++) r[i] = __float_as_uint(__uint_as_float(r[i]) + __uint_as_float(r[16 + i]));
                } else {
                    uint32_t r16[16];
                    tmem_ld16(taddr, r16);
#pragma unroll
                    for (int i = 0; i < 16; i++) { r[i] = r16[i]; r[16 + i] = 0; }
                }
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&c1_empty[t]);
                const bool t64 = In::TAIL64 && t == TILES - 1;
                const int m = t * 128 + (t64 ? q * 16 : q * 32) + lane;
                const int y = m / In::PITCH, x = m - y * In::PITCH;
                if (y < 32 && x < 32 && !(t64 && lane >= 16)) {
                    const int slot = In::slot(y + 1, x + 1);
#pragma unroll
                    for (int g = 0; g < C1 / 8; g++) {
                        float v[8];
                        const float4 b0 = *reinterpret_cast<const float4*>(s_bias1 + g * 8), b1 = *reinterpret_cast<const float4*>(s_bias1 + g * 8 + 4);
                        v[0] = fmaxf(fmaf(__uint_as_float(r[g * 8 + 0]), src.w1_inv, b0.x), 0.f); v[1] = fmaxf(fmaf(__uint_as_float(r[g * 8 + 1]), src.w1_inv, b0.y), 0.f);
                        v[2] = fmaxf(fmaf(__uint_as_float(r[g * 8 + 2]), src.w1_inv, b0.z), 0.f); v[3] = fmaxf(fmaf(__uint_as_float(r[g * 8 + 3]), src.w1_inv, b0.w), 0.f);
                        v[4] = fmaxf(fmaf(__uint_as_float(r[g * 8 + 4]), src.w1_inv, b1.x), 0.f); v[5] = fmaxf(fmaf(__uint_as_float(r[g * 8 + 5]), src.w1_inv, b1.y), 0.f);
                        v[6] = fmaxf(fmaf(__uint_as_float(r[g * 8 + 6]), src.w1_inv, b1.z), 0.f); v[7] = fmaxf(fmaf(__uint_as_float(r[g * 8 + 7]), src.w1_inv, b1.w), 0.f);
                        uint4 pk;
                        pk.x = pack_h2(v[0], v[1]); pk.y = pack_h2(v[2], v[3]); pk.z = pack_h2(v[4], v[5]); pk.w = pack_h2(v[6], v[7]);
                        *reinterpret_cast<uint4*>(st + ((size_t)g * In::NPIX + slot) * 16) = pk;
                        if (SA) {
                            float l[8];
#pragma unroll
                            for (int e = 0; e < 8; e++) l[e] = v[e] - __half2float(__float2half_rn(v[e]));
                            pk.x = pack_h2(l[0], l[1]); pk.y = pack_h2(l[2], l[3]); pk.z = pack_h2(l[4], l[5]); pk.w = pack_h2(l[6], l[7]);
                            *reinterpret_cast<uint4*>(st + ((size_t)(KC + g) * In::NPIX + slot) * 16) = pk;
                        }
                    }
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive(&full[s]);
        }
        RP_STORE(0, 2);
    } else {
        // ===== producers: sampler (or patch load) -> input_norm -> sliding-window planes P_hi / P_lo =====
        const int pt = threadIdx.x - 320;   // 0..127
        const int pw = warp - 10;
        float tp[8][4], fx[8], fy[8];
        // pixel k of this thread: warp pw owns patch block-rows {2pw, 2pw+1} (4 rows each); k = block (4 rows x 8 cols), lane = cell
        auto pix_of = [&](int k) -> int { return ((pw * 2 + (k >> 2)) * 4 + (lane >> 3)) * 32 + (k & 3) * 8 + (lane & 7); };
        auto issue_fetch = [&](int pi) {
            if (src.patches != nullptr) {
                const float* pp = src.patches + (size_t)pi * 1024;
#pragma unroll
                for (int k = 0; k < 8; k++) { tp[k][0] = pp[pix_of(k)]; tp[k][1] = tp[k][2] = tp[k][3] = 0.f; fx[k] = 0.f; fy[k] = 0.f; }
            } else {
                const int b = pi / src.cap;
                const int o = min(max(src.oct[pi], 0), src.geom.n_octaves - 1), l = min(max(src.lvl[pi], 0), src.geom.n_levels - 1);
                const int h = src.geom.h[o], w = src.geom.w[o];
                const float* img = src.pyr + src.geom.off[o][l] + (size_t)b * h * w;
                const float* Lf = src.lafs + (size_t)pi * 6;
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    const int p = pix_of(k);
                    float px, py;
                    laf_sample_xy(Lf, h, w, p >> 5, p & 31, 1.0f / 32.0f, px, py);
                    bilinear_taps(img, h, w, px, py, tp[k], fx[k], fy[k]);
                }
            }
        };
        int pi = next_valid(blockIdx.x);
        if (pi < a.n) issue_fetch(pi);
        int it = 0;
        RP_DECL;
        while (pi < a.n) {
            float* sx = s_x + (it & 1) * SX;
            float* red = s_red + (it & 1) * 8;
            float v[8];
#ifdef AG_ROLE_PROF
            const unsigned long long rp_tc = clock64();   // stalls in the combine = gather latency the prefetch did not hide
#endif
#pragma unroll
            for (int k = 0; k < 8; k++) v[k] = bilinear_combine(tp[k], fx[k], fy[k]);
#ifdef AG_ROLE_PROF
            rp_w[1] += clock64() - rp_tc + (unsigned long long)(__float_as_uint(v[0]) & 0u);
#endif
            const int pn = next_valid(pi + gridDim.x);
            if (pn < a.n) issue_fetch(pn);
            // input_norm: mean, unbiased std + 1e-7 (two passes, as the reference)
            float sm = ((v[0] + v[1]) + (v[2] + v[3])) + ((v[4] + v[5]) + (v[6] + v[7]));
            for (int o = 16; o > 0; o >>= 1) sm += __shfl_xor_sync(0xffffffffu, sm, o);
            if (lane == 0) red[pw * 2] = sm;
            asm volatile("bar.sync 1, 128;" ::: "memory");
            const float mean = ((red[0] + red[2]) + (red[4] + red[6])) / 1024.f;
            float qs = 0.f;
#pragma unroll
            for (int k = 0; k < 8; k++) { const float d = v[k] - mean; qs = fmaf(d, d, qs); }
            for (int o = 16; o > 0; o >>= 1) qs += __shfl_xor_sync(0xffffffffu, qs, o);
            if (lane == 0) red[pw * 2 + 1] = qs;
            asm volatile("bar.sync 1, 128;" ::: "memory");
            const float inv = 1.f / (sqrtf(((red[1] + red[3]) + (red[5] + red[7])) / 1023.f) + 1e-7f);
#pragma unroll
            for (int k = 0; k < 8; k++) { const int p = pix_of(k); sx[((p >> 5) + 1) * 34 + (p & 31) + 1] = (v[k] - mean) * inv; }
            asm volatile("bar.sync 1, 128;" ::: "memory");
            RP_WAIT(0, mbar_wait(p_empty, (it & 1) ^ 1));   // layer-1 MMAs of the previous patch have consumed the planes
#pragma unroll 1
            for (int k = 0; k < NPIXP / 128; k++) {
                const int s0 = pt + k * 128;
                float xv[8];
#pragma unroll
                for (int e = 0; e < 4; e++) { xv[e] = sx[s0 + e]; xv[4 + e] = sx[s0 + In::PITCH + e]; }
                uint4 hi;
                hi.x = pack_h2(xv[0], xv[1]); hi.y = pack_h2(xv[2], xv[3]); hi.z = pack_h2(xv[4], xv[5]); hi.w = pack_h2(xv[6], xv[7]);
                *reinterpret_cast<uint4*>(sP + (size_t)s0 * 16) = hi;
                if (Cfg::X1) {
                    uint4 lo;
                    float r8[8];
#pragma unroll
                    for (int e = 0; e < 8; e++) r8[e] = xv[e] - __half2float(__float2half_rn(xv[e]));
                    lo.x = pack_h2(r8[0], r8[1]); lo.y = pack_h2(r8[2], r8[3]); lo.z = pack_h2(r8[4], r8[5]); lo.w = pack_h2(r8[6], r8[7]);
                    *reinterpret_cast<uint4*>(sP + (size_t)(NPIXP + s0) * 16) = lo;
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive(p_full);
            it++;
            pi = pn;
        }
        RP_STORE(0, 3);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
    }
}

}  // namespace tc
}  // namespace ag
