// Standalone probe of the tcgen05 conventions the tensor-core conv engine relies on (sm_100a):
// no-swizzle K-major shared-memory descriptors (LBO/SBO, 16-byte start-address shifts), instruction
// descriptor for kind::f16 (fp16 x fp16 -> fp32), TMEM alloc / tcgen05.ld 32x32b, tcgen05.commit -> mbarrier,
// cp.async.bulk global -> shared.   Build: nvcc -gencode arch=compute_100a,code=sm_100a -o tc_probe tc_probe.cu
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(2); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;  // version = 1 (Blackwell)
    return d;                // base_offset 0, lbo_mode 0, layout_type 0 (SWIZZLE_NONE / interleave)
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n .reg .pred P;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 P, [%0], %1;\n @P bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}\n"
        :: "r"(bar), "r"(parity) : "memory");
}

constexpr int M = 128, N = 32, K = 32, NPIX = 200, SHIFT = 3;

__global__ void __launch_bounds__(192) probe(const __half* __restrict__ gA, const __half* __restrict__ gB, float* __restrict__ gD) {
    extern __shared__ __align__(128) unsigned char smem[];
    __half* sA = reinterpret_cast<__half*>(smem);                       // [K/8][NPIX][8]
    __half* sB = reinterpret_cast<__half*>(smem + (K / 8) * NPIX * 16);  // [K/8][N][8]
    __shared__ __align__(8) uint64_t bar_b, bar_mma;
    __shared__ uint32_t tmem_base;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar_b)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar_mma)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base)), "r"(32));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    // A via ordinary stores (generic proxy) -> needs fence.proxy.async before the MMA reads it
    for (int i = threadIdx.x; i < (K / 8) * NPIX * 8; i += blockDim.x) sA[i] = gA[i];
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_base;

    if (warp == 0 && lane == 0) {  // B via bulk copy (async proxy)
        const uint32_t bytes = (K / 8) * N * 16;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(&bar_b)), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     :: "r"(smem_u32(sB)), "l"(gB), "r"(bytes), "r"(smem_u32(&bar_b)) : "memory");
    }
    if (warp == 1 && lane == 0) {
        mbar_wait(smem_u32(&bar_b), 0);
        asm volatile("tcgen05.fence::after_thread_sync;");
        // instruction descriptor: D=F32 (bit 4), A=B=F16 (0), K-major both, N>>3 at bit 17, M>>4 at bit 24
        const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        for (int k = 0; k < K / 16; k++) {
            const uint64_t da = make_desc(smem_u32(sA) + (2 * k) * NPIX * 16 + SHIFT * 16, NPIX * 16, 128);
            const uint64_t db = make_desc(smem_u32(sB) + (2 * k) * N * 16, N * 16, 128);
            const uint32_t acc = k > 0;
            asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                         :: "r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bar_mma)) : "memory");
    }
    if (warp >= 2) {  // warps 2..5 -> TMEM lane quadrants (warp % 4)
        mbar_wait(smem_u32(&bar_mma), 0);
        asm volatile("tcgen05.fence::after_thread_sync;");
        const int q = warp & 3;
        uint32_t r[32];
        const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16);
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                     "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                     "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                       "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                       "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                       "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                     : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        const int m = q * 32 + lane;
        for (int n = 0; n < 32; n++) gD[m * N + n] = __uint_as_float(r[n]);
        asm volatile("tcgen05.fence::before_thread_sync;");
    }
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(32));
}

int main() {
    // logical A[m][k] lives at pixel slot (m + SHIFT); other slots hold poison to catch addressing errors
    std::vector<__half> hA((K / 8) * NPIX * 8), hB((K / 8) * N * 8);
    std::vector<float> A(M * K), B(N * K);
    for (auto& v : hA) v = __float2half(777.f);
    srand(1);
    for (int m = 0; m < M; m++)
        for (int k = 0; k < K; k++) {
            float v = (float)((rand() % 17) - 8) / 8.f;
            A[m * K + k] = v;
            hA[((k / 8) * NPIX + (m + SHIFT)) * 8 + (k % 8)] = __float2half(v);
        }
    for (int n = 0; n < N; n++)
        for (int k = 0; k < K; k++) {
            float v = (float)((rand() % 13) - 6) / 4.f;
            B[n * K + k] = v;
            hB[((k / 8) * N + n) * 8 + (k % 8)] = __float2half(v);
        }
    __half *dA, *dB; float* dD;
    CK(cudaMalloc(&dA, hA.size() * 2)); CK(cudaMalloc(&dB, hB.size() * 2)); CK(cudaMalloc(&dD, M * N * 4));
    CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemset(dD, 0, M * N * 4));
    const int smem = (K / 8) * NPIX * 16 + (K / 8) * N * 16;
    probe<<<1, 192, smem>>>(dA, dB, dD);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    std::vector<float> D(M * N);
    CK(cudaMemcpy(D.data(), dD, M * N * 4, cudaMemcpyDeviceToHost));
    double maxerr = 0; int bad = 0;
    for (int m = 0; m < M; m++)
        for (int n = 0; n < N; n++) {
            double ref = 0;
            for (int k = 0; k < K; k++) ref += (double)A[m * K + k] * B[n * K + k];
            double e = fabs(ref - D[m * N + n]);
            if (e > maxerr) maxerr = e;
            if (e > 1e-3 && bad++ < 8) printf("mismatch m=%d n=%d got %f want %f\n", m, n, D[m * N + n], ref);
        }
    printf("tc_probe: max abs err %.3e, %d mismatches -> %s\n", maxerr, bad, bad ? "FAIL" : "PASS");
    return bad ? 1 : 0;
}
