// HardNet 8x8 head on tensor cores: conv8x8(128->128, no bias) == GEMM [n, 8192] x [8192, 128], then BatchNorm and
// L2 normalisation (HardNet.py:86-101, 12-19).  A = trunk features in the HEADL layout written by the last conv layer
// ([patch/128][k/8][patch%128][8] fp16: 128 patches are the M rows of one tile), B = head weights [k/8][cout][8] fp16.
// One CTA per 128-patch tile streams K in 64-wide stages (A 16 KiB + B 16 KiB per stage, bulk copies, 6-stage ring),
// accumulates 128x128 fp32 in TMEM, and the epilogue thread of each row does BN + sum of squares + scale in registers.
#pragma once
#include "tc_conv.cuh"

namespace ag {
namespace tc {

constexpr int HEAD_K = 8192, HEAD_N = 128, HEAD_KS = 64, HEAD_STAGES = 6;
constexpr uint32_t HEAD_STAGE_A = (HEAD_KS / 8) * 128 * 16, HEAD_STAGE_B = (HEAD_KS / 8) * HEAD_N * 16;
constexpr size_t HEAD_SMEM = 1024 + (size_t)HEAD_STAGES * (HEAD_STAGE_A + HEAD_STAGE_B);

__global__ void __launch_bounds__(192, 1) tc_head_kernel(const __half* __restrict__ feat, const __half* __restrict__ wh,
                                                          const float* __restrict__ bn /*scale[128], shift[128]*/, float* __restrict__ out,
                                                          int n, int group, const int* __restrict__ count) {
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    uint64_t* empty = full + HEAD_STAGES;
    uint64_t* done = empty + HEAD_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
    unsigned char* sA = smem + 1024;
    unsigned char* sB = sA + (size_t)HEAD_STAGES * HEAD_STAGE_A;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile = blockIdx.x;
    constexpr int NK = HEAD_K / HEAD_KS;  // 128 stages

    if (threadIdx.x == 0) {
        for (int s = 0; s < HEAD_STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(128));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            const unsigned char* ga = reinterpret_cast<const unsigned char*>(feat) + (size_t)tile * (HEAD_K / 8) * 128 * 16;
            const unsigned char* gb = reinterpret_cast<const unsigned char*>(wh);
            for (int k = 0; k < NK; k++) {
                const int s = k % HEAD_STAGES;
                mbar_wait(&empty[s], ((k / HEAD_STAGES) & 1) ^ 1);
                mbar_expect_tx(&full[s], HEAD_STAGE_A + HEAD_STAGE_B);
                bulk_g2s(sA + (size_t)s * HEAD_STAGE_A, ga + (size_t)k * HEAD_STAGE_A, HEAD_STAGE_A, &full[s]);
                bulk_g2s(sB + (size_t)s * HEAD_STAGE_B, gb + (size_t)k * HEAD_STAGE_B, HEAD_STAGE_B, &full[s]);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(HEAD_N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            for (int k = 0; k < NK; k++) {
                const int s = k % HEAD_STAGES;
                mbar_wait(&full[s], (k / HEAD_STAGES) & 1);
                tc_fence_after();
                const uint32_t a0 = smem_u32(sA + (size_t)s * HEAD_STAGE_A), b0 = smem_u32(sB + (size_t)s * HEAD_STAGE_B);
#pragma unroll
                for (int j = 0; j < HEAD_KS / 16; j++) {
                    const uint64_t da = make_desc(a0 + (uint32_t)(2 * j) * 128u * 16u, 128u * 16u, 128u);
                    const uint64_t db = make_desc(b0 + (uint32_t)(2 * j) * HEAD_N * 16u, HEAD_N * 16u, 128u);
                    umma_f16(tmem, da, db, idesc, (k | j) != 0);
                }
                umma_commit(&empty[s]);
            }
            umma_commit(done);
        }
    } else {
        const int q = warp & 3;
        mbar_wait(done, 0);
        tc_fence_after();
        const int pi = tile * 128 + q * 32 + lane;
        const bool ok = pi < n && (count == nullptr || (pi % group) < count[pi / group]);
        const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16);
        float ss = 0.f;
        // pass 1: sum of squares of the BatchNorm-ed row; pass 2: reload, scale, store (keeps registers low)
#pragma unroll 1
        for (int c0 = 0; c0 < HEAD_N; c0 += 32) {
            uint32_t r[32];
            tmem_ld32(taddr + c0, r);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 32; e++) {
                const float v = fmaf(__uint_as_float(r[e]), __ldg(bn + c0 + e), __ldg(bn + 128 + c0 + e));
                ss = fmaf(v, v, ss);
            }
        }
        const float inv = 1.0f / sqrtf(ss + 1e-8f);
#pragma unroll 1
        for (int c0 = 0; c0 < HEAD_N; c0 += 32) {
            uint32_t r[32];
            tmem_ld32(taddr + c0, r);
            tmem_ld_wait();
            if (ok) {
                float4* o = reinterpret_cast<float4*>(out + (size_t)pi * 128 + c0);
#pragma unroll
                for (int e = 0; e < 32; e += 4) {
                    float4 v;
                    v.x = fmaf(__uint_as_float(r[e + 0]), __ldg(bn + c0 + e + 0), __ldg(bn + 128 + c0 + e + 0)) * inv;
                    v.y = fmaf(__uint_as_float(r[e + 1]), __ldg(bn + c0 + e + 1), __ldg(bn + 128 + c0 + e + 1)) * inv;
                    v.z = fmaf(__uint_as_float(r[e + 2]), __ldg(bn + c0 + e + 2), __ldg(bn + 128 + c0 + e + 2)) * inv;
                    v.w = fmaf(__uint_as_float(r[e + 3]), __ldg(bn + c0 + e + 3), __ldg(bn + 128 + c0 + e + 3)) * inv;
                    o[e / 4] = v;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128));
    }
}

}  // namespace tc
}  // namespace ag
