// 2-CTA (cta_group::2) variant of the tensor-core conv layer for the wide HardNet layers whose weights do not fit one SM.
//
// A cluster of two CTAs on one TPC works on TWO patches at once: every CTA stages its own patch (A operand, 128 rows) and HALF of
// the layer's weights (B operand rows [64 r, 64 r + 64) of each K chunk); the leader CTA issues `tcgen05.mma.cta_group::2` with
// M = 256, N = COUT, which makes each SM multiply its own 128 rows by ALL COUT columns, fetching the other half of B from the peer
// SM.  Compared with splitting COUT over two independent CTAs (NSPLIT = 2) every patch is read and multiplied once instead of
// twice, at the same shared-memory footprint.
//
// Synchronisation (per CTA unless noted):
//   full[s]   : own bulk copy landed (tx)            -> relay thread -> pair_full[s] ON THE LEADER (count 2, remote arrive)
//   empty[s]  : MMAs reading stage s done            <- tcgen05.commit.cta_group::2 multicast to both CTAs
//   tfull[a]  : accumulator a complete               <- multicast commit
//   tempty[a] : ON THE LEADER, count 8               <- the 4 epilogue warps of both CTAs (remote arrive from the peer)
// Only layers with ONE M tile per patch are instantiated (HardNet layers 5 and 6).
#pragma once
#include "tc_conv.cuh"

namespace ag {
namespace tc {

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {   // arrive on the barrier at the same offset in CTA 0 of the cluster
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(remote) : "r"(smem_u32(bar)));
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
// Accumulator hand-over (tempty): carries no data, only "these TMEM columns were read", which tcgen05.fence::before_thread_sync
// orders.  The release form above makes the epilogue warp wait for its global stores first (20-30 % membar stall samples in ncu);
// the relaxed form (default since r02: tc_conv_pair 1.36 -> 1.22 ms per step, results bit-identical; -DAG_PAIR_RELEASE_ARRIVE restores the release form).
__device__ __forceinline__ void mbar_arrive_leader_nodata(uint64_t* bar) {
#ifndef AG_PAIR_RELEASE_ARRIVE
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(remote) : "r"(smem_u32(bar)));
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
#else
    mbar_arrive_leader(bar);
#endif
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n .reg .pred P;\n WAIT_%=:\n mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P, [%0], %1;\n @P bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}\n" ::"r"(
            smem_u32(bar)),
        "r"(parity)
        : "memory");
}
template <int ACC>
__device__ __forceinline__ void umma2_f16_lo(uint32_t tmem_d, uint32_t alo, uint32_t blo, uint32_t idesc) {
    asm volatile(
        "{\n .reg .pred p;\n .reg .b64 da, db;\n setp.ne.u32 p, %4, 0;\n mov.b64 da, {%1, %5};\n mov.b64 db, {%2, %5};\n"
        " tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %3, p;\n}\n" ::"r"(tmem_d),
        "r"(alo), "r"(blo), "r"(idesc), "n"(ACC), "r"(DESC_HI)
        : "memory");
}
__device__ __forceinline__ void umma2_commit_mc(uint64_t* bar) {   // arrives on `bar` (same offset) in both CTAs of the pair
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
                 "h"((uint16_t)3)
                 : "memory");
}

template <int CIN, int COUT, int H, int STRIDE, int STAGES, int OUT>
struct PairCfg {
    using In = InLay<H, STRIDE>;
    static constexpr int HOUT = H / STRIDE, KC = CIN / 8, NH = COUT / 2;   // NH: B rows held by each CTA
    static constexpr int NACC = (512 / COUT) < 4 ? (512 / COUT) : 4;
    static constexpr uint32_t IN_BYTES = (uint32_t)KC * In::NPIX * 16;
    static constexpr uint32_t W_BYTES = 9u * KC * NH * 16;                   // this CTA's half
    static constexpr size_t SMEM = 1024 + (size_t)W_BYTES + (size_t)STAGES * IN_BYTES;
    using OutP = InLay<HOUT, 1>;
    static constexpr int OUT_NPIX = (OUT == PLAIN) ? OutP::NPIX : 0;
    static constexpr size_t OUT_BYTES = (OUT == FINAL) ? (size_t)COUT * HOUT * HOUT * 4 : (OUT == HEADL) ? (size_t)COUT * HOUT * HOUT * 2 : (size_t)(COUT / 8) * OUT_NPIX * 16;
    static_assert(In::TILES == 1, "pair kernel: one M tile per patch");
    static_assert(COUT == 128 && CIN % 16 == 0, "pair kernel shapes");
    static_assert(OUT == PLAIN || OUT == FINAL || OUT == HEADL, "output layouts of the wide layers");
    static_assert(SMEM <= 232448, "shared memory budget");
};

template <int CIN, int COUT, int H, int STRIDE, int STAGES, int OUT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(224, 1) tc_conv_pair_kernel(const ConvArgs a) {
    using Cfg = PairCfg<CIN, COUT, H, STRIDE, STAGES, OUT>;
    using In = typename Cfg::In;
    constexpr int KC = Cfg::KC, NH = Cfg::NH, NACC = Cfg::NACC, HOUT = Cfg::HOUT;
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);   // [STAGES]
    uint64_t* empty = full + STAGES;                        // [STAGES]
    uint64_t* pair_full = empty + STAGES;                   // [STAGES]  (used on the leader)
    uint64_t* tfull = pair_full + STAGES;                   // [NACC]
    uint64_t* tempty = tfull + NACC;                        // [NACC]    (used on the leader)
    uint64_t* wbar = tempty + NACC;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wbar + 1);
    float* s_bias = reinterpret_cast<float*>(smem + 512);   // [COUT]
    unsigned char* sW = smem + 1024;
    unsigned char* sIn = sW + Cfg::W_BYTES;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int pair = blockIdx.x >> 1, npairs_grid = gridDim.x >> 1;
    const int n_pairs = (a.n + 1) >> 1;

    if (threadIdx.x < COUT) s_bias[threadIdx.x] = a.bias[threadIdx.x];
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); mbar_init(&pair_full[s], 2); }
        for (int i = 0; i < NACC; i++) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 8); }
        mbar_init(wbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    cluster_sync_all();   // barriers of both CTAs exist before any remote arrive / multicast commit
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    // patch of this CTA in pair-iteration pp: 2*pp + rank; a CTA whose patch is out of range / not counted still takes part in every barrier
    auto patch_of = [&](int pp) -> int { return 2 * pp + (int)rank; };
    auto valid = [&](int pi) -> bool { return pi < a.n && (a.count == nullptr || (pi % a.group) < a.count[pi / a.group]); };
    auto pair_valid = [&](int pp) -> bool { return valid(2 * pp) || valid(2 * pp + 1); };   // same verdict in both CTAs and all roles

    if (warp == 0) {
        // ===== producer: own half of the weights once, then own patch per pair iteration =====
        if (lane == 0) {
            mbar_expect_tx(wbar, Cfg::W_BYTES);
            bulk_g2s(sW, reinterpret_cast<const unsigned char*>(a.wpk) + (size_t)rank * Cfg::W_BYTES, Cfg::W_BYTES, wbar);
            int itc = 0;
            for (int pp = pair; pp < n_pairs; pp += npairs_grid) {
                if (!pair_valid(pp)) continue;
                const int it = itc++;
                const int s = it % STAGES, pi = patch_of(pp);
                mbar_wait(&empty[s], ((it / STAGES) & 1) ^ 1);
                if (valid(pi)) {
                    mbar_expect_tx(&full[s], (uint32_t)KC * In::USED * 16u);
                    const unsigned char* gsrc = reinterpret_cast<const unsigned char*>(a.in) + (size_t)pi * Cfg::IN_BYTES;
#pragma unroll
                    for (int g = 0; g < KC; g++)
                        bulk_g2s(sIn + (size_t)s * Cfg::IN_BYTES + (size_t)g * In::NPIX * 16, gsrc + (size_t)g * In::NPIX * 16, In::USED * 16u, &full[s]);
                } else {
                    mbar_arrive(&full[s]);   // nothing to load: the stage keeps stale (finite) data, its rows are never stored
                }
            }
        }
    } else if (warp == 6) {
        // ===== relay: "my operands are in place" -> pair_full on the leader =====
        if (lane == 0) {
            mbar_wait(wbar, 0);
            int itc = 0;
            for (int pp = pair; pp < n_pairs; pp += npairs_grid) {
                if (!pair_valid(pp)) continue;
                const int it = itc++;
                const int s = it % STAGES;
                mbar_wait(&full[s], (it / STAGES) & 1);
                mbar_arrive_leader(&pair_full[s]);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (leader CTA only): one cta_group::2 MMA feeds both SMs =====
        if (rank == 0) {
            constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(COUT >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);   // M = 256 over the pair
            const uint32_t leader = elect_one();
            const uint32_t w_lo = desc_lo(smem_u32(sW), NH * 16u);
            int itc = 0;
            for (int pp = pair; pp < n_pairs; pp += npairs_grid) {
                if (!pair_valid(pp)) continue;
                const int it = itc++;
                const int s = it % STAGES, ab = it % NACC;
                mbar_wait_cluster(&pair_full[s], (it / STAGES) & 1);
                mbar_wait_cluster(&tempty[ab], ((it / NACC) & 1) ^ 1);
                tc_fence_after();
                if (leader) {
                    const uint32_t d_tmem = tmem + (uint32_t)(ab * COUT);
                    const uint32_t in_lo = desc_lo(smem_u32(sIn + (size_t)s * Cfg::IN_BYTES), In::NPIX * 16u);
#pragma unroll
                    for (int tap = 0; tap < 9; tap++) {
#pragma unroll
                        for (int j = 0; j < KC / 2; j++) {
                            const uint32_t alo = in_lo + (uint32_t)(In::tap_off(tap / 3, tap % 3) + 2 * j * In::NPIX);
                            const uint32_t blo = w_lo + (uint32_t)((tap * KC + 2 * j) * NH);
                            if (tap == 0 && j == 0) umma2_f16_lo<0>(d_tmem, alo, blo, idesc);
                            else umma2_f16_lo<1>(d_tmem, alo, blo, idesc);
                        }
                    }
                    umma2_commit_mc(&tfull[ab]);
                    umma2_commit_mc(&empty[s]);
                }
                __syncwarp();
            }
        }
    } else if (warp < 6) {
        // ===== epilogue (both CTAs): own 128 rows x COUT columns =====
        const int q = warp & 3, et = (warp - 2) * 32 + lane;
        int itc = 0;
        for (int pp = pair; pp < n_pairs; pp += npairs_grid) {
                if (!pair_valid(pp)) continue;
                const int it = itc++;
            const int ab = it % NACC, pi = patch_of(pp);
            const bool pv = valid(pi);
            unsigned char* outp = reinterpret_cast<unsigned char*>(a.out) + (size_t)pi * Cfg::OUT_BYTES;
            if (OUT == PLAIN && pv) {
                constexpr int HB = HOUT + 1;
                for (int i = et; i < 4 * HB; i += 128) {
                    const int side = i / HB, k = i - side * HB;
                    int Y, X;
                    if (side == 0) { Y = 0; X = k; } else if (side == 1) { Y = HOUT + 1; X = k + 1; } else if (side == 2) { Y = k + 1; X = 0; } else { Y = k; X = HOUT + 1; }
                    const int slot = Cfg::OutP::slot(Y, X);
#pragma unroll
                    for (int g = 0; g < COUT / 8; g++) *reinterpret_cast<uint4*>(outp + ((size_t)g * Cfg::OUT_NPIX + slot) * 16) = make_uint4(0, 0, 0, 0);
                }
            }
            mbar_wait(&tfull[ab], (it / NACC) & 1);
            tc_fence_after();
            const int m = q * 32 + lane;
            const int y = m / In::PITCH, x = m - y * In::PITCH;
            const bool ok = pv && (y < HOUT) && (x < HOUT);
            const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(ab * COUT);
#pragma unroll
            for (int c0 = 0; c0 < COUT; c0 += 32) {
                uint32_t r[32];
                tmem_ld32(taddr + c0, r);
                tmem_ld_wait();
                if (c0 + 32 >= COUT) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_leader_nodata(&tempty[ab]);
                }
                if (ok) {
#pragma unroll
                    for (int g = 0; g < 4; g++) {
                        const int ch = c0 + g * 8;
                        float v[8];
                        const float4 b0 = *reinterpret_cast<const float4*>(s_bias + ch), b1 = *reinterpret_cast<const float4*>(s_bias + ch + 4);
                        v[0] = fmaxf(fmaf(__uint_as_float(r[g * 8 + 0]), a.inv_scale, b0.x), 0.f); v[1] = fmaxf(fmaf(__uint_as_float(r[g * 8 + 1]), a.inv_scale, b0.y), 0.f);
                        v[2] = fmaxf(fmaf(__uint_as_float(r[g * 8 + 2]), a.inv_scale, b0.z), 0.f); v[3] = fmaxf(fmaf(__uint_as_float(r[g * 8 + 3]), a.inv_scale, b0.w), 0.f);
                        v[4] = fmaxf(fmaf(__uint_as_float(r[g * 8 + 4]), a.inv_scale, b1.x), 0.f); v[5] = fmaxf(fmaf(__uint_as_float(r[g * 8 + 5]), a.inv_scale, b1.y), 0.f);
                        v[6] = fmaxf(fmaf(__uint_as_float(r[g * 8 + 6]), a.inv_scale, b1.z), 0.f); v[7] = fmaxf(fmaf(__uint_as_float(r[g * 8 + 7]), a.inv_scale, b1.w), 0.f);
                        if (OUT == FINAL) {
                            float* o = reinterpret_cast<float*>(outp);
#pragma unroll
                            for (int e = 0; e < 8; e++) o[(size_t)(ch + e) * HOUT * HOUT + y * HOUT + x] = v[e];
                        } else {
                            uint4 pk;
                            pk.x = pack_h2(v[0], v[1]); pk.y = pack_h2(v[2], v[3]); pk.z = pack_h2(v[4], v[5]); pk.w = pack_h2(v[6], v[7]);
                            if (OUT == HEADL) {
                                const size_t kch = (size_t)(y * HOUT + x) * (COUT / 8) + ch / 8;
                                unsigned char* hb = reinterpret_cast<unsigned char*>(a.out);
                                *reinterpret_cast<uint4*>(hb + (((size_t)(pi >> 7) * (HOUT * HOUT * COUT / 8) + kch) * 128 + (pi & 127)) * 16) = pk;
                            } else {
                                const int slot = Cfg::OutP::slot(y + 1, x + 1);
                                *reinterpret_cast<uint4*>(outp + ((size_t)(ch / 8) * Cfg::OUT_NPIX + slot) * 16) = pk;
                            }
                        }
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();   // no CTA frees TMEM / exits while its peer may still signal or multiply
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
    }
}

}  // namespace tc
}  // namespace ag
