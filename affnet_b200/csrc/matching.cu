// Descriptor matching right after the hot path: distance_matrix_vector (Losses.py:5-13) and the
// second-nearest-neighbour ratio test of train_AffNet_test_on_graffity.py:292-298, including its quirk: before the second
// minimum is taken, EVERY column that is the nearest neighbour of ANY row is set to 100000 (`dist_matrix[:, idxs_in_2] = 100000`).
// fp32 throughout (the ratio test is a decision; a 2000 x 2000 x 128 product is 1 GFLOP, far below anything worth tensor cores).
#include "common.cuh"

namespace ag {

constexpr int MT = 64, MK = 16;

// dist[i][j] = sqrt((|a_i|^2 + |b_j|^2 - 2 a_i.b_j) + 1e-6)
__global__ void __launch_bounds__(256) dist_matrix_kernel(const float* __restrict__ a, int n1, const float* __restrict__ b, int n2, int D,
                                                           float* __restrict__ out) {
    __shared__ float sa[MK][MT + 1], sb[MK][MT + 1];
    const int i0 = blockIdx.y * MT, j0 = blockIdx.x * MT;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;   // 16 x 16 threads, 4 x 4 outputs each
    float acc[4][4] = {}, na[4] = {}, nb[4] = {};
    for (int k0 = 0; k0 < D; k0 += MK) {
        for (int t = threadIdx.x; t < MT * MK; t += 256) {
            const int r = t / MK, k = t - r * MK;
            sa[k][r] = (i0 + r < n1 && k0 + k < D) ? a[(size_t)(i0 + r) * D + k0 + k] : 0.f;
            sb[k][r] = (j0 + r < n2 && k0 + k < D) ? b[(size_t)(j0 + r) * D + k0 + k] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < MK; k++) {
            float va[4], vb[4];
#pragma unroll
            for (int q = 0; q < 4; q++) { va[q] = sa[k][ty * 4 + q]; vb[q] = sb[k][tx * 4 + q]; }
#pragma unroll
            for (int q = 0; q < 4; q++) {
                na[q] = fmaf(va[q], va[q], na[q]); nb[q] = fmaf(vb[q], vb[q], nb[q]);
#pragma unroll
                for (int p = 0; p < 4; p++) acc[q][p] = fmaf(va[q], vb[p], acc[q][p]);
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int q = 0; q < 4; q++)
#pragma unroll
        for (int p = 0; p < 4; p++) {
            const int i = i0 + ty * 4 + q, j = j0 + tx * 4 + p;
            if (i < n1 && j < n2) out[(size_t)i * n2 + j] = sqrtf(__fadd_rn(__fsub_rn(__fadd_rn(na[q], nb[p]), __fmul_rn(2.0f, acc[q][p])), 1e-6f));
        }
}

// per row: minimum and first arg-min; marks the arg-min column as taken
__global__ void row_min_kernel(const float* __restrict__ dist, int n1, int n2, float* __restrict__ mn, int* __restrict__ arg,
                               unsigned char* __restrict__ colmask) {
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (i >= n1) return;
    float best = INFINITY;
    int bj = 0x7fffffff;   // stays so only when the lane saw nothing smaller than +inf (e.g. an all-NaN row): resolved to column 0 below
    for (int j = lane; j < n2; j += 32) {
        const float v = dist[(size_t)i * n2 + j];
        if (v < best) { best = v; bj = j; }
    }
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oj = __shfl_xor_sync(0xffffffffu, bj, o);
        if (ov < best || (ov == best && oj < bj)) { best = ov; bj = oj; }
    }
    if (lane == 0) {
        if (bj < 0 || bj >= n2) bj = 0;       // a row of NaNs has no minimum: never index colmask out of bounds
        mn[i] = best; arg[i] = bj; colmask[bj] = 1;
    }
}

// per row: minimum after `dist[:, idxs] = 100000`, then the ratio test
__global__ void row_second_kernel(const float* __restrict__ dist, int n1, int n2, const unsigned char* __restrict__ colmask,
                                  const float* __restrict__ mn, float ratio, float* __restrict__ second, unsigned char* __restrict__ keep) {
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (i >= n1) return;
    float best = INFINITY;
    for (int j = lane; j < n2; j += 32) best = fminf(best, colmask[j] ? 100000.0f : dist[(size_t)i * n2 + j]);
    for (int o = 16; o > 0; o >>= 1) best = fminf(best, __shfl_xor_sync(0xffffffffu, best, o));
    if (lane == 0) {
        second[i] = best;
        keep[i] = (__fdiv_rn(mn[i], __fadd_rn(best, 1e-8f)) <= ratio) ? 1 : 0;
    }
}

}  // namespace ag

using namespace ag;

extern "C" {

int ag_distance_matrix(const float* d_a, int n1, const float* d_b, int n2, int dim, float* d_out, void* stream) {
    AG_REQUIRE(d_a && d_b && d_out && n1 >= 1 && n2 >= 1 && dim >= 1, "bad arguments");
    dist_matrix_kernel<<<dim3(cdiv(n2, MT), cdiv(n1, MT)), 256, 0, (cudaStream_t)stream>>>(d_a, n1, d_b, n2, dim, d_out);
    AG_CHECK_LAUNCH("dist_matrix_kernel");
    return AG_OK;
}

size_t ag_match_snn_workspace_bytes(int n1, int n2) { return align_up((size_t)n1 * n2 * sizeof(float), 256) + align_up((size_t)n2, 256); }

int ag_match_snn(const float* d_desc1, int n1, const float* d_desc2, int n2, int dim, float ratio, void* d_ws, size_t ws_bytes, int* d_idx2,
                 float* d_min, float* d_second, unsigned char* d_keep, void* stream) {
    AG_REQUIRE(d_desc1 && d_desc2 && d_ws && d_idx2 && d_min && d_second && d_keep, "NULL argument");
    AG_REQUIRE(n1 >= 1 && n2 >= 1 && dim >= 1, "bad sizes");
    if (ws_bytes < ag_match_snn_workspace_bytes(n1, n2)) { set_error("ag_match_snn: workspace too small"); return AG_ERR_CAPACITY; }
    cudaStream_t st = (cudaStream_t)stream;
    float* dist = (float*)d_ws;
    unsigned char* colmask = (unsigned char*)d_ws + align_up((size_t)n1 * n2 * sizeof(float), 256);
    int rc = check_cuda(cudaMemsetAsync(colmask, 0, n2, st), "memset column mask");
    if (rc) return rc;
    if ((rc = ag_distance_matrix(d_desc1, n1, d_desc2, n2, dim, dist, stream))) return rc;
    row_min_kernel<<<cdiv(n1, 8), 256, 0, st>>>(dist, n1, n2, d_min, d_idx2, colmask);
    AG_CHECK_LAUNCH("row_min_kernel");
    row_second_kernel<<<cdiv(n1, 8), 256, 0, st>>>(dist, n1, n2, colmask, d_min, ratio, d_second, d_keep);
    AG_CHECK_LAUNCH("row_second_kernel");
    return AG_OK;
}

}  // extern "C"
