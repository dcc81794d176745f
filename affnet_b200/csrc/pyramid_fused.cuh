// One launch per octave of the Gaussian scale-space pyramid (replaces five / four chained blur launches of ScalePyramid.forward,
// HandCraftedModules.py:23-56): all levels of the octave are produced from ONE read of the octave's input, rows streamed through
// shared memory, the stride-2 decimated seed of the next octave written from the same registers.
//
// A CTA owns a column strip x [x0, x0+SW) and a row band [y0, y1) of one image and runs the chain of NB separable blurs as a software
// pipeline over rows ("steps").  Stage l (64 threads = 64 column quads of a 256-float row buffer) does per step
//   H phase: the horizontal pass of its newest input row (image row from a cp.async.bulk ring for stage 0, the row the previous stage
//            produced in the previous step otherwise) -> ring_l[row & 15]
//   V phase: the vertical pass over ring_l rows v-R .. v+R (replicate-clamped at the image border) -> level l row v: global memory
//            (128-bit stores, only inside the CTA's own strip and band) + the shared-memory row that feeds stage l+1 + the decimated seed.
// Halos: stage l computes columns / rows [x0 - A_l, x1 + A_l) with A_l = sum of the radii of the later stages, clipped to the image; the
// replicate border (Utils.py:160-166: F.pad(mode='replicate')) is materialised in the row buffers by the thread that owns the edge pixel.
// Arithmetic is exactly blur_kernel's (pyramid.cu): acc = 0; acc = fmaf(w[k], x[k], acc) for k = 0 .. 2R, horizontal then vertical, so
// the levels are bit-identical to the per-level launches (tests/test_gpu_parity.py compares them).
#pragma once
#include "common.cuh"

namespace ag {
namespace pf {

constexpr int BW = 256;          // floats of a row buffer: 64 column quads
constexpr int HALO = 32;         // buffer column 0 is image column x0 - HALO
constexpr int SWMAX = BW - 2 * HALO;   // 192 output columns per strip at most
constexpr int RING = 16;         // rows of a stage's horizontal-pass ring (>= 2 * 7 + 1, power of two)
constexpr int NIMG = 8, PD = 5;  // image-row ring and prefetch distance
constexpr int MAXB = 5;

struct OctArgs {
    const float* src;            // [B][h][w] input of stage 0
    float* out[MAXB];            // [B][h][w] output of each stage
    float* dec;                  // [B][h2][w2] decimated copy of stage `seed`'s output (or NULL)
    int seed;
    int B, h, w;
    int sw, nstrips, band, nbands;   // strip width (multiple of 4, <= SWMAX), rows per band
    float taps[MAXB][16];        // taps[l][0 .. 2 R_l]
};

// chain 0: octave 0 (image -> level 0 -> .. -> level 4), chain 1: octaves >= 1 (level 0 -> level 1 .. level 4); nlevels = 3, init_sigma = 1.6
template <int CH> struct Chain;
template <> struct Chain<0> { static constexpr int NB = 5; __host__ __device__ static constexpr int R(int l) { return l == 0 ? 5 : l == 1 ? 4 : l == 2 ? 5 : l == 3 ? 6 : 7; } };
template <> struct Chain<1> { static constexpr int NB = 4; __host__ __device__ static constexpr int R(int l) { return l == 0 ? 4 : l == 1 ? 5 : l == 2 ? 6 : 7; } };

template <int CH> __host__ __device__ constexpr int halo_after(int l) { int a = 0; for (int m = l + 1; m < Chain<CH>::NB; m++) a += Chain<CH>::R(m); return a; }
template <int CH> __host__ __device__ constexpr int delay_of(int l) { int d = 0; for (int m = 0; m < l; m++) d += Chain<CH>::R(m) + 1; return d; }

template <int CH>
struct Smem {
    float ring[Chain<CH>::NB][RING][BW];      // horizontal-pass rows
    float orow[Chain<CH>::NB][2][BW];         // vertical-pass output row of each stage (input of the next), double buffered by step parity
    float img[NIMG][BW];                      // image rows (cp.async.bulk)
    unsigned long long bar[NIMG];
};

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// stage L of chain CH: one thread = one column quad
template <int CH, int L>
__device__ __forceinline__ void stage_step(Smem<CH>& S, const OctArgs& a, int s, int q, int b, int x0, int x1, int y0, int y1, int i0, bool vphase) {
    constexpr int R = Chain<CH>::R(L), A = halo_after<CH>(L), D = delay_of<CH>(L), NB = Chain<CH>::NB;
    const int xb = x0 - HALO;
    const int c = xb + 4 * q;                                   // first image column of this quad
    const int clo = max(0, x0 - A), chi = min(a.w, x1 + A);     // columns this stage produces
    const bool col_on = (c + 4 > clo) && (c < chi);
    const float* tw = a.taps[L];
    if (!vphase) {
        // ---- horizontal pass of input row u ----
        const int u = i0 + s - D;
        const int ulo = max(0, y0 - A - R), uhi = min(a.h, y1 + A + R);
        if (u < ulo || u >= uhi) return;
        const float* in = (L == 0) ? S.img[(u - i0) & (NIMG - 1)] : S.orow[L > 0 ? L - 1 : 0][(s + 1) & 1];   // stage L-1 wrote it in step s-1
        if (L == 0) {   // the row was requested PD steps ago
            const int k = u - i0;
            const uint32_t parity = (uint32_t)((k / NIMG) & 1);
            asm volatile("{\n .reg .pred P;\n W_%=:\n mbarrier.try_wait.parity.shared::cta.b64 P, [%0], %1;\n @P bra D_%=;\n bra W_%=;\n D_%=:\n}\n" ::"r"(s32(&S.bar[k & (NIMG - 1)])), "r"(parity) : "memory");
        }
        if (!col_on) return;
        constexpr int R4 = (R + 3) / 4 * 4, OFF = R4 - R, NQ = (4 + 2 * R + OFF + 3) / 4;
        float v[NQ * 4];
        const bool edge = (c - R4 < 0) || (c + 4 + R4 > a.w);     // part of the window lies outside the image (replicate) - or, for stage 0, outside the loaded row
        if (!edge) {
            const float4* row = reinterpret_cast<const float4*>(in + 4 * q - R4);
#pragma unroll
            for (int j = 0; j < NQ; j++) { const float4 t = row[j]; v[4 * j] = t.x; v[4 * j + 1] = t.y; v[4 * j + 2] = t.z; v[4 * j + 3] = t.w; }
        } else {
#pragma unroll
            for (int j = 0; j < NQ * 4; j++) v[j] = in[clampi(c - R4 + j, 0, a.w - 1) - xb];
        }
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int k = 0; k <= 2 * R; k++) {
            const float wk = tw[k];
            o.x = fmaf(wk, v[OFF + k], o.x); o.y = fmaf(wk, v[OFF + k + 1], o.y);
            o.z = fmaf(wk, v[OFF + k + 2], o.z); o.w = fmaf(wk, v[OFF + k + 3], o.w);
        }
        *reinterpret_cast<float4*>(&S.ring[L][u & (RING - 1)][4 * q]) = o;
    } else {
        // ---- vertical pass: level row v ----
        const int v = i0 + s - D - R;
        const int vlo = max(0, y0 - A), vhi = min(a.h, y1 + A);
        if (v < vlo || v >= vhi || !col_on) return;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int k = 0; k <= 2 * R; k++) {
            const int rr = clampi(v - R + k, 0, a.h - 1);
            const float4 t = *reinterpret_cast<const float4*>(&S.ring[L][rr & (RING - 1)][4 * q]);
            const float wk = tw[k];
            acc.x = fmaf(wk, t.x, acc.x); acc.y = fmaf(wk, t.y, acc.y); acc.z = fmaf(wk, t.z, acc.z); acc.w = fmaf(wk, t.w, acc.w);
        }
        if (L + 1 < NB) *reinterpret_cast<float4*>(&S.orow[L][s & 1][4 * q]) = acc;
        if (v >= y0 && v < y1 && c >= x0 && c < x1) {
            *reinterpret_cast<float4*>(a.out[L] + ((size_t)b * a.h + v) * a.w + c) = acc;
            if (a.dec != nullptr && L == a.seed && (v & 1) == 0) {
                const int h2 = (a.h + 1) >> 1, w2 = (a.w + 1) >> 1;
                *reinterpret_cast<float2*>(a.dec + ((size_t)b * h2 + (v >> 1)) * w2 + (c >> 1)) = make_float2(acc.x, acc.z);
            }
        }
    }
}

template <int CH>
__global__ void __launch_bounds__(64 * Chain<CH>::NB) octave_kernel(const OctArgs a) {
    constexpr int NB = Chain<CH>::NB;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Smem<CH>& S = *reinterpret_cast<Smem<CH>*>(smem_raw);
    const int strip = blockIdx.x % a.nstrips, band = blockIdx.x / a.nstrips, b = blockIdx.y;
    const int x0 = strip * a.sw, x1 = min(a.w, x0 + a.sw);
    const int y0 = band * a.band, y1 = min(a.h, y0 + a.band);
    constexpr int ATOT = halo_after<CH>(-1);                       // sum of all radii
    const int i0 = max(0, y0 - ATOT), i1 = min(a.h, y1 + ATOT);    // input rows of the chain
    const int stage = threadIdx.x >> 6, q = threadIdx.x & 63;
    const int xb = x0 - HALO;
    const int lx0 = max(0, xb), lx1 = min(a.w, xb + BW);           // image columns that are loaded into a row buffer
    if (threadIdx.x == 0) {
        for (int i = 0; i < NIMG; i++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&S.bar[i])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto request = [&](int k) {   // input row i0 + k -> img[k % NIMG]
        const int u = i0 + k;
        if (u >= i1) return;
        const uint32_t bytes = (uint32_t)(lx1 - lx0) * 4u;
        const uint32_t bar = s32(&S.bar[k & (NIMG - 1)]);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(&S.img[k & (NIMG - 1)][lx0 - xb])),
                     "l"(a.src + ((size_t)b * a.h + u) * a.w + lx0), "r"(bytes), "r"(bar)
                     : "memory");
    };
    if (threadIdx.x == 0)
        for (int k = 0; k < PD; k++) request(k);
    constexpr int DLAST = delay_of<CH>(NB - 1) + Chain<CH>::R(NB - 1);
    const int steps = (y1 - i0) + DLAST;          // the last stage's vertical pass reaches row y1 - 1 in step y1 - 1 - i0 + DLAST
#pragma unroll 1
    for (int s = 0; s < steps; s++) {
        if (threadIdx.x == 0) request(s + PD);     // its slot held row s + PD - NIMG, consumed before the barrier that ended step s + PD - NIMG
        switch (stage) {
            case 0: stage_step<CH, 0>(S, a, s, q, b, x0, x1, y0, y1, i0, false); break;
            case 1: stage_step<CH, 1>(S, a, s, q, b, x0, x1, y0, y1, i0, false); break;
            case 2: stage_step<CH, 2>(S, a, s, q, b, x0, x1, y0, y1, i0, false); break;
            case 3: stage_step<CH, 3>(S, a, s, q, b, x0, x1, y0, y1, i0, false); break;
            default: if (NB > 4) stage_step<CH, (NB > 4 ? 4 : 0)>(S, a, s, q, b, x0, x1, y0, y1, i0, false); break;
        }
        __syncthreads();
        switch (stage) {
            case 0: stage_step<CH, 0>(S, a, s, q, b, x0, x1, y0, y1, i0, true); break;
            case 1: stage_step<CH, 1>(S, a, s, q, b, x0, x1, y0, y1, i0, true); break;
            case 2: stage_step<CH, 2>(S, a, s, q, b, x0, x1, y0, y1, i0, true); break;
            case 3: stage_step<CH, 3>(S, a, s, q, b, x0, x1, y0, y1, i0, true); break;
            default: if (NB > 4) stage_step<CH, (NB > 4 ? 4 : 0)>(S, a, s, q, b, x0, x1, y0, y1, i0, true); break;
        }
        __syncthreads();
    }
}

}  // namespace pf
}  // namespace ag
