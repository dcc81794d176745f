// Internal definition of the opaque ag_net_t handle shared by the fp32 SIMT engine (nets_simt.cu) and the
// tensor-core engine (nets_tc.cu).
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"

#define AG_ENGINE_SIMT 0  /* exact fp32 direct convolution */
#define AG_ENGINE_TC 1    /* tcgen05: fp16 operands, fp32 accumulate; layer 1 and heads in fp32 */
#define AG_ENGINE_TC_EXACT 2 /* tcgen05 with fp16 residual planes of weights AND activations (what OriNet always uses) */

struct ag_net {
    int kind;
    int engine;
    float* d_w[6];     // fp32 [9][cin][cout], BN folded   (d_w[0] doubles as the [9][C] first-layer weights)
    float* d_b[6];     // fp32 [cout]  (BN shift)
    float* d_w1;       // == d_w[0]
    __half* d_wh[6];   // fp16 [nsplit][hi|lo][9][cin/8][cout/nsplit][8] for layers 1..5 (index 0 unused; lo only for AffNet/OriNet)
    __half* d_headh;   // HardNet head for the tensor-core GEMM: fp16 [8192/8][128][8], k = (pixel*16 + c/8)*8 + c%8
    float* d_head_w;   // AffNet [3][4096], OriNet w_eff[4096][18] (per-position shifted copies), HardNet [8192][128]
    float* d_head_b;   // AffNet bias[3], OriNet bias[2], HardNet {scale[128], shift[128]}
    float* d_all;      // fp32 allocation
    __half* d_all_h;   // fp16 allocation
};

#include "tc_conv.cuh"

namespace ag {
size_t tc_act_bytes(int kind);
int tc_nsplit(int kind, int layer);
int tc_split_w(int kind);
tc::FirstSrc tc_src_patches(const float* patches);
tc::FirstSrc tc_src_pyramid(const ag_pyramid_plan_t* p, const float* pyr, const float* lafs, const int* oct, const int* lvl, int cap);
int tc_hardnet_forward(const ag_net* net, const tc::FirstSrc& src, int n, int group, const int* count, void* bufA, void* bufB, void* headbuf,
                       float* out, cudaStream_t st);
int tc_trunk_orinet(const ag_net* net, const tc::FirstSrc& src, int n, int group, const int* count, void* bufA, void* bufB, float* feat,
                    cudaStream_t st);
int tc_trunk_affnet(const ag_net* net, const tc::FirstSrc& src, int n, int group, const int* count, void* bufA, void* bufB, float* feat,
                    cudaStream_t st);
}  // namespace ag
