// Microbenchmark of the tcgen05 rates the conv engine's design rests on (sm_100a, one CTA on one SM):
//   1. tcgen05.mma kind::f16, K = 16, A and B from shared memory (no-swizzle K-major), M = 128 / 64, N = 16 .. 256: clk per MMA
//   2. the same with the A operand in TMEM
//   3. tcgen05.ld 32x32b (x32 / x64 / x128) with 1, 2, 4 and 8 reading warps: bytes per clk per SM
//   4. warp shuffles: clk per SHFL with 4 warps
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tc_rates tc_rates.cu ; run on a B200.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(2); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n .reg .pred P;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 P, [%0], %1;\n @P bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}\n"
        :: "r"(bar), "r"(parity) : "memory");
}
constexpr uint32_t DESC_HI = 0x4008u;   // SBO = 128 B, version 1
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr, uint32_t lbo_bytes) { return ((saddr >> 4) & 0x3FFFu) | ((lbo_bytes >> 4) << 16); }
__device__ __forceinline__ void umma_ss(uint32_t d, uint32_t alo, uint32_t blo, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n .reg .pred p;\n .reg .b64 da, db;\n setp.ne.u32 p, %4, 0;\n mov.b64 da, {%1, %5};\n mov.b64 db, {%2, %5};\n"
        " tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n}\n" ::"r"(d), "r"(alo), "r"(blo), "r"(idesc), "r"(acc), "r"(DESC_HI) : "memory");
}
__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a_tmem, uint32_t blo, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n .reg .pred p;\n .reg .b64 db;\n setp.ne.u32 p, %4, 0;\n mov.b64 db, {%2, %5};\n"
        " tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %3, p;\n}\n" ::"r"(d), "r"(a_tmem), "r"(blo), "r"(idesc), "r"(acc), "r"(DESC_HI) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

constexpr int NPIX = 1408;   // slots of the A buffer (11 tiles of 128 rows)

// MODE 0: SS, 1: TS (A in TMEM columns 256..).  The issue loop is fully unrolled with compile-time operand offsets (as the
// production issuers are): one thread, 9 taps x TILES tiles, accumulators rotating over NACC buffers per tile.
template <int MODE, int M, int N>
__global__ void __launch_bounds__(192) mma_rate(long long* out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base;
    unsigned char* sA = smem;                         // [2][NPIX][16 B]
    unsigned char* sB = smem + 2 * NPIX * 16;         // [9 taps][2][256][16 B]
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < (2 * NPIX * 16 + 9 * 2 * 256 * 16) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;  // fp16 1.0
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_base;
    if (threadIdx.x == 32) {
        constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        const uint32_t a_lo = desc_lo(smem_u32(sA), NPIX * 16u), b_lo = desc_lo(smem_u32(sB), 256 * 16u);
        constexpr int TILES = 8;
        for (int rep = 0; rep < 3; rep++) {   // rep 0 warms up
            const long long t0 = clock64();
#pragma unroll 1
            for (int round = 0; round < 4; round++) {
#pragma unroll
                for (int tile = 0; tile < TILES; tile++) {
                    const uint32_t d = tmem + (uint32_t)((tile & 1) * (N <= 128 ? 128 : 0));
#pragma unroll
                    for (int tap = 0; tap < 9; tap++) {
                        const uint32_t alo = a_lo + (uint32_t)(tile * 128 + (tap / 3) * 34 + tap % 3);
                        const uint32_t blo = b_lo + (uint32_t)(tap * 2 * 256);
                        if (MODE == 0) umma_ss(d, alo, blo, idesc, tap > 0);
                        else umma_ts(d, tmem + 256 + (uint32_t)((tap % 8) * 8), blo, idesc, tap > 0);
                    }
                }
            }
            umma_commit(smem_u32(&bar));
            mbar_wait(smem_u32(&bar), rep & 1);
            const long long t1 = clock64();
            if (rep == 2) out[0] = t1 - t0;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
    }
}

template <int MODE, int M, int N>
static double run_mma(long long* d_out, int smem) {
    CK(cudaFuncSetAttribute(mma_rate<MODE, M, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    mma_rate<MODE, M, N><<<1, 192, smem>>>(d_out);
    CK(cudaDeviceSynchronize());
    long long c;
    CK(cudaMemcpy(&c, d_out, 8, cudaMemcpyDeviceToHost));
    return (double)c / (4 * 8 * 9);
}
template <int MODE, int M>
static void run_mma_row(long long* d_out, int smem) {
    printf("%4s M=%3d  N: 16 %.1f | 32 %.1f | 48 %.1f | 64 %.1f | 96 %.1f | 128 %.1f | 160 %.1f | 192 %.1f | 256 %.1f   (clk per MMA; floor N/2)\n", MODE ? "TS" : "SS", M,
           run_mma<MODE, M, 16>(d_out, smem), run_mma<MODE, M, 32>(d_out, smem), run_mma<MODE, M, 48>(d_out, smem), run_mma<MODE, M, 64>(d_out, smem),
           run_mma<MODE, M, 96>(d_out, smem), run_mma<MODE, M, 128>(d_out, smem), run_mma<MODE, M, 160>(d_out, smem), run_mma<MODE, M, 192>(d_out, smem),
           run_mma<MODE, M, 256>(d_out, smem));
}


// The same issue loop from NI issuer threads (one per warp, warps 1 .. NI) running concurrently, each into its own accumulator columns
// and with its own completion barrier: does the small-N floor belong to the issuing thread or to the tensor pipe?
// out[0] = cycles from the earliest start to the latest completion; every issuer issues the full 288 MMAs.
template <int NI, int N>
__global__ void __launch_bounds__(192) mma_rate_multi(long long* out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar[4];
    __shared__ uint32_t tmem_base;
    __shared__ long long t_begin[4], t_end[4];
    unsigned char* sA = smem;
    unsigned char* sB = smem + 2 * NPIX * 16;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < (2 * NPIX * 16 + 9 * 2 * 256 * 16) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 4; i++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[i])));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_base;
    if ((threadIdx.x & 31) == 0 && warp >= 1 && warp <= NI) {
        const int me = warp - 1;
        constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t a_lo = desc_lo(smem_u32(sA), NPIX * 16u), b_lo = desc_lo(smem_u32(sB), 256 * 16u);
        const uint32_t dbase = tmem + (uint32_t)(me * 128);          // N <= 64: two buffers of 64 columns per issuer
        for (int rep = 0; rep < 3; rep++) {
            const long long t0 = clock64();
#pragma unroll 1
            for (int round = 0; round < 4; round++) {
#pragma unroll
                for (int tile = 0; tile < 8; tile++) {
                    const uint32_t d = dbase + (uint32_t)((tile & 1) * 64);
#pragma unroll
                    for (int tap = 0; tap < 9; tap++) {
                        const uint32_t alo = a_lo + (uint32_t)(tile * 128 + (tap / 3) * 34 + tap % 3);
                        const uint32_t blo = b_lo + (uint32_t)(tap * 2 * 256);
                        umma_ss(d, alo, blo, idesc, tap > 0);
                    }
                }
            }
            umma_commit(smem_u32(&bar[me]));
            mbar_wait(smem_u32(&bar[me]), rep & 1);
            const long long t1 = clock64();
            if (rep == 2) { t_begin[me] = t0; t_end[me] = t1; }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        long long b = t_begin[0], e = t_end[0];
        for (int w = 1; w < NI; w++) { b = min(b, t_begin[w]); e = max(e, t_end[w]); }
        out[0] = e - b;
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
    }
}
template <int NI, int N>
static double run_multi(long long* d_out, int smem) {
    CK(cudaFuncSetAttribute(mma_rate_multi<NI, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    mma_rate_multi<NI, N><<<1, 192, smem>>>(d_out);
    CK(cudaDeviceSynchronize());
    long long c;
    CK(cudaMemcpy(&c, d_out, 8, cudaMemcpyDeviceToHost));
    return (double)c / (4 * 8 * 9 * NI);     // clk per MMA of the SM (all issuers together)
}

template <int X>
__device__ __forceinline__ uint32_t tmem_ld_sum(uint32_t taddr);
template <>
__device__ __forceinline__ uint32_t tmem_ld_sum<32>(uint32_t taddr) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
          "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
          "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 32; i++) s ^= r[i];
    return s;
}
template <>
__device__ __forceinline__ uint32_t tmem_ld_sum<64>(uint32_t taddr) {
    uint32_t r[64];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, "
        "%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, "
        "%48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
          "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
          "=r"(r[31]), "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]), "=r"(r[40]),
          "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]), "=r"(r[48]), "=r"(r[49]), "=r"(r[50]),
          "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]), "=r"(r[56]), "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]),
          "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 64; i++) s ^= r[i];
    return s;
}


__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
          "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
          "=r"(r[31])
        : "r"(taddr));
}
template <>
__device__ __forceinline__ uint32_t tmem_ld_sum<33>(uint32_t taddr) {   // 33 = two x32 loads in flight, one wait
    uint32_t r[32], q[32];
    tmem_ld32_nowait(taddr, r);
    tmem_ld32_nowait(taddr ^ 32u, q);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 32; i++) s ^= r[i] ^ q[i];
    return s;
}

// nw reading warps (warps 0..nw-1; quadrant = warp % 4), each doing `iters` loads of X columns; pipelined = two loads in flight
template <int X>
__global__ void __launch_bounds__(512) ldtm_rate(int nw, int iters, long long* out, uint32_t* sink) {
    __shared__ uint32_t tmem_base;
    __shared__ long long t_begin[16], t_end[16];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_base;
    uint32_t s = 0;
    if (warp < nw) {
        const uint32_t base = tmem + ((uint32_t)((warp & 3) * 32) << 16);
        for (int rep = 0; rep < 2; rep++) {
            const long long t0 = clock64();
            for (int i = 0; i < iters; i++) s ^= tmem_ld_sum<X>(base + (uint32_t)((i * 64) & 255) + (uint32_t)((warp >> 2) & 1) * 256u);
            const long long t1 = clock64();
            if (lane == 0) { t_begin[warp] = t0; t_end[warp] = t1; }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        long long b = t_begin[0], e = t_end[0];
        for (int w = 1; w < nw; w++) { b = min(b, t_begin[w]); e = max(e, t_end[w]); }
        out[0] = e - b;
    }
    if (s == 0x12345u) sink[threadIdx.x] = s;
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
    }
}

__global__ void __launch_bounds__(256) shfl_rate(int nw, int iters, long long* out, float* sink) {
    __shared__ long long t_begin[8], t_end[8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float v[8];
    for (int k = 0; k < 8; k++) v[k] = (float)(threadIdx.x * 8 + k);
    if (warp < nw) {
        const long long t0 = clock64();
        for (int i = 0; i < iters; i++) {
#pragma unroll
            for (int k = 0; k < 8; k++) v[k] += __shfl_up_sync(0xffffffffu, v[(k + 1) & 7], 1);
        }
        const long long t1 = clock64();
        if (lane == 0) { t_begin[warp] = t0; t_end[warp] = t1; }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        long long b = t_begin[0], e = t_end[0];
        for (int w = 1; w < nw; w++) { b = min(b, t_begin[w]); e = max(e, t_end[w]); }
        out[0] = e - b;
    }
    float s = 0;
    for (int k = 0; k < 8; k++) s += v[k];
    if (s == 12345.f) sink[threadIdx.x] = s;
}

int main() {
    long long* d_out; uint32_t* d_sink;
    CK(cudaMalloc(&d_out, 64)); CK(cudaMalloc(&d_sink, 4096));
    const int smem = 2 * NPIX * 16 + 9 * 2 * 256 * 16 + 1024;
    printf("== tcgen05.mma kind::f16 K=16: clk per MMA, 288 MMAs issued back to back by one thread (unrolled)\n");
    run_mma_row<0, 128>(d_out, smem);
    run_mma_row<0, 64>(d_out, smem);
    run_mma_row<1, 128>(d_out, smem);
    run_mma_row<1, 64>(d_out, smem);
    printf("== the same from 1 / 2 / 4 issuer warps at once (M=128, SS): clk per MMA of the SM\n");
    printf("  N=16: 1 issuer %.1f | 2 issuers %.1f | 4 issuers %.1f\n", run_multi<1, 16>(d_out, smem), run_multi<2, 16>(d_out, smem), run_multi<4, 16>(d_out, smem));
    printf("  N=48: 1 issuer %.1f | 2 issuers %.1f | 4 issuers %.1f\n", run_multi<1, 48>(d_out, smem), run_multi<2, 48>(d_out, smem), run_multi<4, 48>(d_out, smem));
    printf("  N=64: 1 issuer %.1f | 2 issuers %.1f | 4 issuers %.1f\n", run_multi<1, 64>(d_out, smem), run_multi<2, 64>(d_out, smem), run_multi<4, 64>(d_out, smem));
    printf("== tcgen05.ld 32x32b: bytes per clk per SM\n");
    for (int nw : {1, 2, 4, 8, 16}) {
        long long c32, c64;
        const int iters = 512;
        ldtm_rate<32><<<1, 512>>>(nw, iters, d_out, d_sink); CK(cudaDeviceSynchronize()); CK(cudaMemcpy(&c32, d_out, 8, cudaMemcpyDeviceToHost));
        ldtm_rate<64><<<1, 512>>>(nw, iters, d_out, d_sink); CK(cudaDeviceSynchronize()); CK(cudaMemcpy(&c64, d_out, 8, cudaMemcpyDeviceToHost));
        long long c33;
        ldtm_rate<33><<<1, 512>>>(nw, iters, d_out, d_sink); CK(cudaDeviceSynchronize()); CK(cudaMemcpy(&c33, d_out, 8, cudaMemcpyDeviceToHost));
        printf("warps %d: x32 %.1f B/clk (%.1f clk per load) | x64 %.1f B/clk (%.1f clk per load) | 2 x x32 in flight %.1f B/clk (%.1f clk per pair)\n", nw, (double)nw * iters * 32 * 32 * 4 / c32, (double)c32 / iters,
               (double)nw * iters * 32 * 64 * 4 / c64, (double)c64 / iters, (double)nw * iters * 32 * 64 * 4 / c33, (double)c33 / iters);
    }
    printf("== SHFL: clk per warp-shuffle (per warp), 8 independent chains\n");
    for (int nw : {1, 4, 8}) {
        long long c;
        shfl_rate<<<1, 256>>>(nw, 1024, d_out, (float*)d_sink); CK(cudaDeviceSynchronize()); CK(cudaMemcpy(&c, d_out, 8, cudaMemcpyDeviceToHost));
        printf("warps %d: %.2f clk per SHFL+FADD per warp\n", nw, (double)c / (1024 * 8));
    }
    return 0;
}
