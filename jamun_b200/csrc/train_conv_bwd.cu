// Backward of the equivariant convolution (Conv.forward, /root/reference/src/jamun/e3tools/nn/_conv.py:93-119) in the
// aggregate-then-transform formulation of the forward kernels:
//
//   forward   A_i = sum_{e->i} h'_e (x) f_e,  out_i = G-scale_i * A_i . M        (+ path 0e(x)1e->1e via Y = x_s . M2)
//   backward  G_i    = alpha / deg_i * dOut_i                                      jamun_conv_bwd_scale
//             dM     = sum_i A_i^T (x) G_i          (A recomputed by the builder)  jamun_stage_atb  ("A^T . B" over the nodes)
//             dA_i   = M . G_i                                                     jamun_gemm_tf32x3 (column blocks, tcgen05)
//             dh'_e  = <dA_i, f_e>,  df_e = h'_e . dA_i   (per edge, receiver CTA) jamun_conv_bwd_edge
//             path 2: dT_e = rhat_e . G1_i,  dY_j = sum_{e from j} h'_e (x) dT_e,  dh'_e += Y_j . dT_e     jamun_conv_bwd_p2
//                     dM2 = x_s^T . dY (jamun_stage_atb),  dx_s += dY . M2^T (jamun_gemm_tf32x3)
//             dx_j   = sum_{e from j} J_e^T df_e  (+ dx_s)  -- source-major, out-edge lists sorted: deterministic
//                                                                                  jamun_conv_bwd_gather
// Every reduction has a fixed order; no atomics.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "umma.cuh"

namespace {
using namespace jb;

constexpr float kInvSqrt3 = 0.57735026918962576451f;
constexpr float kInvSqrt2 = 0.70710678118654752440f;

// ---- G = alpha * dOut / deg ---------------------------------------------------------------------------------------------------
__global__ void conv_bwd_scale_kernel(const float* __restrict__ dout, const float* __restrict__ inv_deg, float a0, float a1, int N,
                                      float* __restrict__ g) {
    const size_t total = (size_t)N * JAMUN_GATE_IN;
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(t / JAMUN_GATE_IN), c = (int)(t % JAMUN_GATE_IN);
        g[t] = dout[t] * inv_deg[i] * (c < JAMUN_S + JAMUN_V ? a0 : a1);
    }
}

// ---- stage-major operand, transposed, times a row-major matrix ----------------------------------------------------------------
// acc[u, w] = sum_c sum_{r < rows} A_c[stage][r][u] * B[r][b_col0 + c*b_comp_stride + w]     (u < 32, w < W <= 160)
// A: jamun_gemm_tf32x3's A layout ([stage][rows_pad][32], 16-byte chunks XOR-swizzled with row & 7), component c at
// a + c*a_comp_stride.  Output element (stage = k*nslots + slot, u, w):
//   mode 0:  row = slot_row0[slot] + u (skipped if >= slot_row0[slot] + slot_rows[slot]);  out[(k*out_rows + row)*W + w]
//   mode 1:  transposed (the operand's 32 columns are the output's minor index):           out[(k*out_rows + w)*32 + u]
struct AtbParams {
    const float* a;
    long long a_comp_stride;
    int ncomp, rows, rows_pad, nslots;
    const float* b;
    int ldb, b_col0, b_comp_stride, W;
    float* out;
    int mode, out_rows;
    int slot_row0[8], slot_rows[8];
};

__global__ void __launch_bounds__(256) stage_atb_kernel(const AtbParams P) {
    __shared__ __align__(16) float As[32][32];
    __shared__ float Bs[32][161];
    const int stage = blockIdx.x;
    const int u4 = threadIdx.x & 7, wg = threadIdx.x >> 3;  // 4 operand columns x (up to 5) matrix columns wg + 32 q
    float acc[4][5];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int q = 0; q < 5; ++q) acc[i][q] = 0.f;
    for (int c = 0; c < P.ncomp; ++c) {
        const float* a = P.a + (size_t)c * P.a_comp_stride + (size_t)stage * P.rows_pad * 32;
        const int bc = P.b_col0 + c * P.b_comp_stride;
        for (int r0 = 0; r0 < P.rows; r0 += 32) {
            // operand tile: 32 rows x 128 B, contiguous; rows beyond `rows` hold stale data -> masked
            {
                const int rr = threadIdx.x >> 3, ch = threadIdx.x & 7;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (r0 + rr < P.rows) v = *reinterpret_cast<const float4*>(a + (size_t)(r0 + rr) * 32 + 4 * ch);
                *reinterpret_cast<float4*>(&As[rr][4 * ch]) = v;  // kept swizzled: un-swizzled on read
            }
            for (int t = threadIdx.x; t < 32 * P.W; t += 256) {
                const int rr = t / P.W, w = t - rr * P.W;
                Bs[rr][w] = (r0 + rr < P.rows) ? P.b[(size_t)(r0 + rr) * P.ldb + bc + w] : 0.f;
            }
            __syncthreads();
#pragma unroll 8
            for (int rr = 0; rr < 32; ++rr) {
                const float4 av = *reinterpret_cast<const float4*>(&As[rr][4 * (u4 ^ ((r0 + rr) & 7))]);
#pragma unroll
                for (int q = 0; q < 5; ++q) {
                    const float bv = Bs[rr][wg + 32 * q];
                    acc[0][q] = fmaf(av.x, bv, acc[0][q]);
                    acc[1][q] = fmaf(av.y, bv, acc[1][q]);
                    acc[2][q] = fmaf(av.z, bv, acc[2][q]);
                    acc[3][q] = fmaf(av.w, bv, acc[3][q]);
                }
            }
            __syncthreads();
        }
    }
    const int k = stage / P.nslots, slot = stage - k * P.nslots;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int u = 4 * u4 + i;
#pragma unroll
        for (int q = 0; q < 5; ++q) {
            const int w = wg + 32 * q;
            if (w >= P.W) continue;
            if (P.mode == 0) {
                if (u < P.slot_rows[slot]) P.out[((size_t)k * P.out_rows + P.slot_row0[slot] + u) * P.W + w] = acc[i][q];
            } else {
                P.out[((size_t)k * P.out_rows + w) * 32 + u] = acc[i][q];
            }
        }
    }
}

// ---- per-edge backward of the aggregated paths ---------------------------------------------------------------------------------
// One CTA per receiver i.  dA0_i [65][NSL0*32] and dA1_i[c] [65][64] (rows of the column-block GEMM outputs) are staged in shared
// memory once and reused by all in-edges of i, which are processed four at a time (features / gradients of the batch live in
// shared memory as float4 = one value per edge, so every dA element read from shared memory feeds four FMAs):
//   df0[u'] = sum_k' h'[k'] dA0[k'][u']      df1[c][u'] = sum_k' h'[k'] dA1[c][k'][u']        (threads over u')
//   dh[k']  = sum_u' f0[u'] dA0[k'][u'] + sum_c sum_u' f1[c][u'] dA1[c][k'][u']   (k' < 64)   (warps over k', lanes over u')
//   dxe[e]  = J_e^T df  (row of the per-edge input gradient, SoA layout; summed per source by jamun_conv_bwd_gather)
// The raw gathers (x row of the source, rhat, h) of batch b+1 are issued into registers before the arithmetic of batch b.
template <int S_IN, int V_IN>
__global__ void __launch_bounds__(384)
conv_bwd_edge_kernel(const float* __restrict__ x, const int* __restrict__ rowptr, const int* __restrict__ col,
                     const float* __restrict__ h, const float* __restrict__ rhat, const float* __restrict__ dA0, int ld0,
                     const float* __restrict__ dA1, int ld1, long long dA1_comp_stride, int N, float* __restrict__ dh,
                     float* __restrict__ dxe) {
    constexpr int D_IN = S_IN + 3 * V_IN;
    constexpr int NS = (S_IN + 31) / 32;
    constexpr int NS32 = NS * 32;
    constexpr int W0 = (NS + (V_IN > 0 ? 1 : 0)) * 32;  // padded 0e feature columns
    constexpr int W1 = V_IN > 0 ? 64 : 0;
    constexpr int NF = W0 + 3 * W1;
    constexpr int EB = 4;                                // edges per batch
    constexpr int XS_PER = (EB * NS32 + 383) / 384;      // scalar gathers per thread and batch
    extern __shared__ __align__(16) float sm[];
    float* sA0 = sm;                                         // [65][W0]
    float* sA1 = sA0 + 65 * W0;                              // [3][65][W1]
    float4* sf = reinterpret_cast<float4*>(sA1 + 3 * 65 * W1);  // [NF] features, one lane of the float4 per edge of the batch
    float4* sdf = sf + NF;                                   // [NF] their gradients
    float4* sh = sdf + NF;                                   // [65] h'
    float4* srh = sh + 65;                                   // [EB] rhat of the batch's edges
    const int i = blockIdx.x;
    if (i >= N) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int e0 = rowptr[i], e1 = rowptr[i + 1];
    if (e0 == e1) return;
    for (int t = tid; t < 65 * W0; t += 384) sA0[t] = dA0[(size_t)i * ld0 + t];
    if (V_IN > 0)
        for (int c = 0; c < 3; ++c)
            for (int t = tid; t < 65 * W1; t += 384) sA1[c * 65 * W1 + t] = dA1[(size_t)c * dA1_comp_stride + (size_t)i * ld1 + t];

    // raw gathers of one batch, held in registers: scalars (XS_PER per thread), vectors + rhat (threads < EB*32), h (threads >= 128)
    float r_xs[XS_PER], r_v[3] = {0.f, 0.f, 0.f}, r_h = 0.f;
    float4 r_rh = make_float4(0.f, 0.f, 0.f, 0.f);
    auto gather = [&](int eb) {
#pragma unroll
        for (int m = 0; m < XS_PER; ++m) {
            const int t = tid + 384 * m, q = t / NS32, c = t - q * NS32;
            r_xs[m] = (t < EB * NS32 && eb + q < e1 && c < S_IN) ? x[(size_t)col[eb + q] * D_IN + c] : 0.f;
        }
        if (V_IN > 0 && tid < EB * 32) {
            const int q = tid >> 5, w = tid & 31;
            if (eb + q < e1) {
                const float* xr = x + (size_t)col[eb + q] * D_IN + S_IN + w;
                r_v[0] = xr[0], r_v[1] = xr[V_IN], r_v[2] = xr[2 * V_IN];
                r_rh = *reinterpret_cast<const float4*>(rhat + 4 * (size_t)(eb + q));
            } else {
                r_v[0] = r_v[1] = r_v[2] = 0.f;
                r_rh = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        if (tid >= 128) {
            const int t2 = tid - 128, q = t2 >> 6, k = t2 & 63;
            r_h = eb + q < e1 ? h[(size_t)(eb + q) * JAMUN_EDGE_HID + k] : 0.f;
        }
    };
    gather(e0);
    for (int eb = e0; eb < e1; eb += EB) {
        const int nb = e1 - eb < EB ? e1 - eb : EB;
        __syncthreads();  // previous batch done with sf / sdf / sh (and the dA tiles are loaded)
        float* sff = reinterpret_cast<float*>(sf);
#pragma unroll
        for (int m = 0; m < XS_PER; ++m) {
            const int t = tid + 384 * m, q = t / NS32, c = t - q * NS32;
            if (t < EB * NS32) sff[4 * c + q] = r_xs[m];
        }
        if (V_IN > 0 && tid < EB * 32) {
            const int q = tid >> 5, w = tid & 31;
            const float vx = r_v[0], vy = r_v[1], vz = r_v[2];
            const float4 rh = r_rh;
            sff[4 * (NS32 + w) + q] = vx * rh.x + vy * rh.y + vz * rh.z;
            sff[4 * (W0 + 0 * 64 + w) + q] = vx * kInvSqrt3;
            sff[4 * (W0 + 1 * 64 + w) + q] = vy * kInvSqrt3;
            sff[4 * (W0 + 2 * 64 + w) + q] = vz * kInvSqrt3;
            sff[4 * (W0 + 0 * 64 + 32 + w) + q] = (vy * rh.z - vz * rh.y) * kInvSqrt2;
            sff[4 * (W0 + 1 * 64 + 32 + w) + q] = (vz * rh.x - vx * rh.z) * kInvSqrt2;
            sff[4 * (W0 + 2 * 64 + 32 + w) + q] = (vx * rh.y - vy * rh.x) * kInvSqrt2;
            if (w == 0) srh[q] = rh;
        }
        if (tid >= 128) {
            const int t2 = tid - 128, q = t2 >> 6, k = t2 & 63;
            reinterpret_cast<float*>(sh)[4 * k + q] = r_h;
        }
        if (tid < EB) reinterpret_cast<float*>(sh)[4 * 64 + tid] = tid < nb ? 1.f : 0.f;  // bias channel h' = 1
        __syncthreads();
        if (eb + EB < e1) gather(eb + EB);  // next batch's loads fly under this batch's arithmetic
        // df: one feature column per thread, four edges at once
        if (tid < NF) {
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            const float* A = tid < W0 ? sA0 + tid : sA1 + ((tid - W0) / 64) * 65 * W1 + (tid - W0) % 64;
            const int lda = tid < W0 ? W0 : W1;
#pragma unroll 5
            for (int k = 0; k < 65; ++k) {
                const float a = A[k * lda];
                const float4 hk = sh[k];
                jb::ffma2(acc.x, acc.y, a, hk.x, hk.y);  // packed FP32 FMA: two edges per issue slot
                jb::ffma2(acc.z, acc.w, a, hk.z, hk.w);
            }
            sdf[tid] = acc;
        }
        // dh: a warp takes eight channels at a time, four edges at once -- the edge features (float4 per lane: four shared-memory
        // wavefronts per load) are read once per eight channels instead of once per channel
        if (warp < 8) {
            const int k0 = warp * 8;
            float4 acc[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int u = lane; u < W0; u += 32) {
                const float4 f = sf[u];
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const float a = sA0[(k0 + c) * W0 + u];
                    jb::ffma2(acc[c].x, acc[c].y, a, f.x, f.y);
                    jb::ffma2(acc[c].z, acc[c].w, a, f.z, f.w);
                }
            }
            if (V_IN > 0) {
#pragma unroll
                for (int c3 = 0; c3 < 3; ++c3)
#pragma unroll
                    for (int hv = 0; hv < 2; ++hv) {
                        const float4 f = sf[W0 + c3 * 64 + 32 * hv + lane];
#pragma unroll
                        for (int c = 0; c < 8; ++c) {
                            const float a = sA1[(c3 * 65 + k0 + c) * W1 + 32 * hv + lane];
                            jb::ffma2(acc[c].x, acc[c].y, a, f.x, f.y);
                            jb::ffma2(acc[c].z, acc[c].w, a, f.z, f.w);
                        }
                    }
            }
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const float sx = warp_sum(acc[c].x), sy = warp_sum(acc[c].y), sz = warp_sum(acc[c].z), sw = warp_sum(acc[c].w);
                if (lane < nb) dh[(size_t)(eb + lane) * JAMUN_EDGE_HID + k0 + c] = lane == 0 ? sx : lane == 1 ? sy : lane == 2 ? sz : sw;
            }
        }
        __syncthreads();
        // dxe = J^T df
        const float* sdff = reinterpret_cast<const float*>(sdf);
        for (int t = tid; t < EB * S_IN; t += 384) {
            const int q = t / S_IN, c = t - q * S_IN;
            if (q < nb) dxe[(size_t)(eb + q) * D_IN + c] = sdff[4 * c + q];
        }
        if (V_IN > 0 && tid < EB * 32) {
            const int q = tid >> 5, w = tid & 31;
            if (q < nb) {
                const float4 rh = srh[q];
                float* o = dxe + (size_t)(eb + q) * D_IN;
                const float dd = sdff[4 * (NS32 + w) + q];
                const float cx = sdff[4 * (W0 + 0 * 64 + 32 + w) + q], cy = sdff[4 * (W0 + 1 * 64 + 32 + w) + q],
                            cz = sdff[4 * (W0 + 2 * 64 + 32 + w) + q];
                // cross = x_v x rhat  =>  d x_v = rhat x d cross
                o[S_IN + w] = dd * rh.x + sdff[4 * (W0 + 0 * 64 + w) + q] * kInvSqrt3 + (rh.y * cz - rh.z * cy) * kInvSqrt2;
                o[S_IN + V_IN + w] = dd * rh.y + sdff[4 * (W0 + 1 * 64 + w) + q] * kInvSqrt3 + (rh.z * cx - rh.x * cz) * kInvSqrt2;
                o[S_IN + 2 * V_IN + w] = dd * rh.z + sdff[4 * (W0 + 2 * 64 + w) + q] * kInvSqrt3 + (rh.x * cy - rh.y * cx) * kInvSqrt2;
            }
        }
    }
}

// ---- the same per-edge backward on the warp-level tensor cores (default) ---------------------------------------------------------
// Per receiver both products are small dense GEMMs over the staged dA_i [65][NFC] (NFC = compact feature columns: scalars, v.rhat,
// then per component v/sqrt3 and the cross product):
//   product 1   dF^T [NFC x edges]  = dA_i^T [NFC x 64] . H^T [64 x edges]   (+ the bias row dA_i[64][:])
//   product 2   dH^T [64 x edges]   = dA_i   [64 x NFC] . F^T [NFC x edges]
// with the edges as the N dimension of mma.sync.m16n8k8 (tf32 operands, fp32 accumulate), so in-degrees are padded to a multiple
// of 8 only.  fp32 accuracy comes from the same three-product split as the forward kernels (hi = rna_tf32(v), lo = rna_tf32(v-hi);
// lo.hi + hi.lo + hi.hi), done in registers as the fragments are loaded.  tcgen05 is the wrong tool here: its M is 64/128 rows, a
// receiver has <= 35 edges, and 2 x 91 KB of pre-split dA_i would not fit next to the operands -- measured on the aggregate
// builder, small SS-mode tcgen05 products retire at ~300 clk each, more than a warp-level m16n8k8 triple per tile costs here.
// One persistent CTA per SM (16 warps); per batch of <= 32 in-edges the tile tasks are dealt one per warp
// (product 2 split in four K slices whose partial tiles are summed in slice order: deterministic).  Shared-memory strides
// are chosen so that every fragment load but the product-2 A load (2-way) is conflict-free: LDA = NFC (24 mod 32), LDF = NFC + 4
// (28 mod 32), LDH = 68.
// hi = the tf32 the tensor core reads anyway (it ignores the low 13 mantissa bits), lo = the exact remainder (its own low bits are
// ignored in turn: 2^-20 relative).  cvt.rna.tf32 would cost four instructions per value on this architecture; this is two.
__device__ __forceinline__ void split_tf32(float v, uint32_t& hi, uint32_t& lo) {
    hi = __float_as_uint(v) & 0xFFFFE000u;
    lo = __float_as_uint(v - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int S_IN, int V_IN>
struct EdgeMmaCfg {
    static constexpr int D_IN = S_IN + 3 * V_IN;
    static constexpr int NS32 = ((S_IN + 31) / 32) * 32;
    static constexpr int W0 = NS32 + (V_IN > 0 ? 32 : 0);   // row length of dA0 in global memory (padded scalars + v.rhat)
    static constexpr int NFC = S_IN + V_IN + 6 * V_IN;      // compact feature columns
    static constexpr int LDA = NFC, LDF = NFC + 4, LDH = 68;
    static constexpr int NB = 32;                           // in-edges per batch (four n-tiles)
    static constexpr int LDP = NB + 1;                      // partial-tile rows (odd: the channel-major read-back is conflict-free)
    static constexpr int MT1 = (NFC + 15) / 16;             // product 1: m-tiles over the features
    static constexpr int KS2 = NFC / 8;                     // product 2: k-steps over the features
    static constexpr int NSL = 4;                           // ... in four slices
    static constexpr int kThreads = 512;
    static constexpr size_t kSmemFloats = (size_t)65 * LDA + 8 + 2 * NB * LDF + NB * LDH + NSL * 64 * LDP + 4 * NB + 4;  // + mbarrier
    static_assert(S_IN % 4 == 0 && V_IN % 4 == 0 && NFC % 8 == 0, "16-byte staging chunks / whole k-steps");
    static_assert(LDA % 32 == 24 || LDA % 32 == 8, "conflict-free product-1 A fragments");
    static_assert(LDF % 32 == 28 || LDF % 32 == 4, "conflict-free B fragments");
    static_assert((65 * LDA + 8 + 2 * NB * LDF + NB * LDH + NSL * 64 * LDP) % 4 == 0, "float4 alignment of the rhat slots");
};

// The tile tasks of one batch with NTILES n-tiles of 8 edges, one per warp: warps 0..7 = product 2 (a pair of m-tiles = 32
// channels, one of four K slices), warps 8..15 = product 1 (ceil(MT1/8) m-tiles -- three = 48 feature columns for the hidden layers --, all eight k-steps).  A task splits
// its B fragments once per k-step and reuses them for all of its m-tiles; 21.5 / 24 (m-tile, k-step) units per warp.
template <int S_IN, int V_IN, int NTILES>
__device__ __forceinline__ void edge_mma_tasks(const float* __restrict__ sA, const float* __restrict__ sF, const float* __restrict__ sH,
                                               float* __restrict__ sDF, float* __restrict__ sP, int warp, int g, int tig) {
    using C = EdgeMmaCfg<S_IN, V_IN>;
    constexpr int NFC = C::NFC, LDA = C::LDA, LDF = C::LDF, LDH = C::LDH, LDP = C::LDP;
    static_assert(C::kThreads == 512 && C::NSL == 4, "one task per warp");
    if (warp < 8) {
        const int pair = warp & 1, sl = warp >> 1;
        const int ks0 = sl * C::KS2 / C::NSL, ks1 = (sl + 1) * C::KS2 / C::NSL;
        float acc[2][NTILES][4];
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int nt = 0; nt < NTILES; ++nt) acc[i][nt][0] = acc[i][nt][1] = acc[i][nt][2] = acc[i][nt][3] = 0.f;
        const float* Ar = sA + (32 * pair + g) * LDA + tig;
        const float* Br = sF + g * LDF + tig;
        for (int ks = ks0; ks < ks1; ++ks) {
            uint32_t bh[NTILES][2], bl[NTILES][2];
#pragma unroll
            for (int nt = 0; nt < NTILES; ++nt) {
                split_tf32(Br[8 * nt * LDF + 8 * ks], bh[nt][0], bl[nt][0]);
                split_tf32(Br[8 * nt * LDF + 8 * ks + 4], bh[nt][1], bl[nt][1]);
            }
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                uint32_t ahi[4], alo[4];
                split_tf32(Ar[16 * i * LDA + 8 * ks], ahi[0], alo[0]);
                split_tf32(Ar[(16 * i + 8) * LDA + 8 * ks], ahi[1], alo[1]);
                split_tf32(Ar[16 * i * LDA + 8 * ks + 4], ahi[2], alo[2]);
                split_tf32(Ar[(16 * i + 8) * LDA + 8 * ks + 4], ahi[3], alo[3]);
                // term by term across the n-tiles: consecutive MMAs never wait on each other's accumulator
#pragma unroll
                for (int nt = 0; nt < NTILES; ++nt) mma_tf32(acc[i][nt], alo, bh[nt][0], bh[nt][1]);
#pragma unroll
                for (int nt = 0; nt < NTILES; ++nt) mma_tf32(acc[i][nt], ahi, bl[nt][0], bl[nt][1]);
#pragma unroll
                for (int nt = 0; nt < NTILES; ++nt) mma_tf32(acc[i][nt], ahi, bh[nt][0], bh[nt][1]);
            }
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            float* P = sP + (sl * 64 + 32 * pair + 16 * i + g) * LDP + 2 * tig;
#pragma unroll
            for (int nt = 0; nt < NTILES; ++nt) {
                P[8 * nt] = acc[i][nt][0], P[8 * nt + 1] = acc[i][nt][1];
                P[8 * LDP + 8 * nt] = acc[i][nt][2], P[8 * LDP + 8 * nt + 1] = acc[i][nt][3];
            }
        }
    } else {
        constexpr int MG = (C::MT1 + 7) / 8;  // m-tiles per product-1 task
        const int mt0 = MG * (warp - 8);
        const int nmt = C::MT1 - mt0 < MG ? C::MT1 - mt0 : MG;
        if (nmt <= 0) return;
        float acc[MG][NTILES][4];
#pragma unroll
        for (int i = 0; i < MG; ++i)
#pragma unroll
            for (int nt = 0; nt < NTILES; ++nt) acc[i][nt][0] = acc[i][nt][1] = acc[i][nt][2] = acc[i][nt][3] = 0.f;
        const float* Ar = sA + tig * LDA + 16 * mt0 + g;  // A[row = feature][col = channel] = dA[channel][feature]
        const float* Br = sH + g * LDH + tig;
#pragma unroll 2
        for (int ks = 0; ks < 8; ++ks) {
            uint32_t bh[NTILES][2], bl[NTILES][2];
#pragma unroll
            for (int nt = 0; nt < NTILES; ++nt) {
                split_tf32(Br[8 * nt * LDH + 8 * ks], bh[nt][0], bl[nt][0]);
                split_tf32(Br[8 * nt * LDH + 8 * ks + 4], bh[nt][1], bl[nt][1]);
            }
#pragma unroll
            for (int i = 0; i < MG; ++i)
                if (i < nmt) {
                    uint32_t ahi[4], alo[4];
                    split_tf32(Ar[8 * ks * LDA + 16 * i], ahi[0], alo[0]);
                    split_tf32(Ar[8 * ks * LDA + 16 * i + 8], ahi[1], alo[1]);
                    split_tf32(Ar[(8 * ks + 4) * LDA + 16 * i], ahi[2], alo[2]);
                    split_tf32(Ar[(8 * ks + 4) * LDA + 16 * i + 8], ahi[3], alo[3]);
#pragma unroll
                    for (int nt = 0; nt < NTILES; ++nt) mma_tf32(acc[i][nt], alo, bh[nt][0], bh[nt][1]);
#pragma unroll
                    for (int nt = 0; nt < NTILES; ++nt) mma_tf32(acc[i][nt], ahi, bl[nt][0], bl[nt][1]);
#pragma unroll
                    for (int nt = 0; nt < NTILES; ++nt) mma_tf32(acc[i][nt], ahi, bh[nt][0], bh[nt][1]);
                }
        }
        // + bias channel (h' = 1), transposed into sDF[edge][feature]
#pragma unroll
        for (int i = 0; i < MG; ++i)
            if (i < nmt) {
                const int fa = 16 * (mt0 + i) + g, fb = fa + 8;
                const float ba = fa < NFC ? sA[64 * LDA + fa] : 0.f, bb = fb < NFC ? sA[64 * LDA + fb] : 0.f;
#pragma unroll
                for (int nt = 0; nt < NTILES; ++nt) {
                    float* D = sDF + (8 * nt + 2 * tig) * LDF;
                    if (fa < NFC) D[fa] = acc[i][nt][0] + ba, D[LDF + fa] = acc[i][nt][1] + ba;
                    if (fb < NFC) D[fb] = acc[i][nt][2] + bb, D[LDF + fb] = acc[i][nt][3] + bb;
                }
            }
    }
}

template <int S_IN, int V_IN>
__global__ void __launch_bounds__(512, 1)
conv_bwd_edge_mma_kernel(const float* __restrict__ x, const int* __restrict__ rowptr, const int* __restrict__ col,
                         const float* __restrict__ h, const float* __restrict__ rhat, const float* __restrict__ dA0, int ld0,
                         const float* __restrict__ dA1, int ld1, long long dA1_comp_stride, int N, float* __restrict__ dh,
                         float* __restrict__ dxe) {
    using C = EdgeMmaCfg<S_IN, V_IN>;
    constexpr int D_IN = C::D_IN, NFC = C::NFC, LDA = C::LDA, LDF = C::LDF, LDH = C::LDH, LDP = C::LDP, NB = C::NB, NW = C::kThreads / 32;
    constexpr int B1 = S_IN + V_IN;  // first per-component column
    extern __shared__ __align__(16) float sm[];
    float* sA = sm;                       // [65][LDA] (+8 floats: the last partial m-tile reads past row 64's end)
    float* sF = sA + 65 * LDA + 8;        // [NB][LDF] edge features
    float* sDF = sF + NB * LDF;           // [NB][LDF] their gradients
    float* sH = sDF + NB * LDF;           // [NB][LDH] radial hidden channels
    float* sP = sH + NB * LDH;            // [NSL][64][LDP] product-2 partial tiles
    float4* srh = reinterpret_cast<float4*>(sP + C::NSL * 64 * LDP);  // [NB]
    uint64_t* bar = reinterpret_cast<uint64_t*>(srh + NB);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, tig = lane & 3;
    static_assert(NB == 2 * NW, "the gather takes two edges per warp");
    if (tid == 0) {
        umma::mbar_init(bar, 1);
        umma::fence_barrier_init();
    }
    __syncthreads();
    uint32_t phase = 0;

    // raw gathers of one batch (two edges per warp), held in registers: the first batch of the NEXT receiver is loaded before the
    // tile tasks of the current receiver's last batch, so its latency is covered by the tensor-core phase
    constexpr int NSC = (S_IN + 31) / 32;
    float xs[2][NSC], xv[2][3], hh[2][2];
    float4 rh[2];
    auto gather = [&](int eb, int e1) {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int q = warp + NW * u;
            const bool on = eb + q < e1;
            const float* xr = x + (size_t)(on ? col[eb + q] : 0) * D_IN;
#pragma unroll
            for (int m = 0; m < NSC; ++m) xs[u][m] = (on && 32 * m + lane < S_IN) ? xr[32 * m + lane] : 0.f;
            if (V_IN > 0) {
#pragma unroll
                for (int c = 0; c < 3; ++c) xv[u][c] = on ? xr[S_IN + c * V_IN + lane] : 0.f;
                rh[u] = on ? *reinterpret_cast<const float4*>(rhat + 4 * (size_t)(eb + q)) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            hh[u][0] = on ? h[(size_t)(eb + q) * JAMUN_EDGE_HID + lane] : 0.f;
            hh[u][1] = on ? h[(size_t)(eb + q) * JAMUN_EDGE_HID + 32 + lane] : 0.f;
        }
    };
    int e0 = 0, e1 = 0;
    if ((int)blockIdx.x < N) {
        e0 = rowptr[blockIdx.x], e1 = rowptr[blockIdx.x + 1];
        gather(e0, e1);
    }
    for (int i = blockIdx.x; i < N; i += gridDim.x) {
        // the next receiver's row extent (its gather is issued below); `continue` for an isolated node still advances to it
        const int i_next = i + gridDim.x;
        int e0n = 0, e1n = 0;
        if (i_next < N) e0n = rowptr[i_next], e1n = rowptr[i_next + 1];
        if (e0 == e1) {
            e0 = e0n, e1 = e1n;
            gather(e0, e1);
            continue;
        }
        // stage dA_i (compacted) with 1-D bulk async copies, one contiguous piece per thread: per channel row the scalars, the
        // v.rhat columns and the three components (the previous receiver's last barrier ordered all reads of sA before this)
        {
            constexpr int PIECES = V_IN > 0 ? 5 : 1;
            if (tid == 0) umma::mbar_arrive_expect_tx(bar, 65 * NFC * (uint32_t)sizeof(float));
            if (tid < 65 * PIECES) {
                const int k = tid / PIECES, pc = tid - k * PIECES;
                const float* a0 = dA0 + (size_t)i * ld0 + k * C::W0;
                float* dst = sA + k * LDA;
                if (pc == 0) umma::bulk_g2s(dst, a0, S_IN * sizeof(float), bar);
                else if (pc == 1) umma::bulk_g2s(dst + S_IN, a0 + C::NS32, V_IN * sizeof(float), bar);
                else
                    umma::bulk_g2s(dst + B1 + (pc - 2) * 2 * V_IN, dA1 + (size_t)(pc - 2) * dA1_comp_stride + (size_t)i * ld1 + k * 2 * V_IN,
                                   2 * V_IN * sizeof(float), bar);
            }
        }
        for (int eb = e0; eb < e1; eb += NB) {
            const int nb = e1 - eb < NB ? e1 - eb : NB;
            const int ntiles = (nb + 7) >> 3, npad = ntiles * 8;
            // ---- features, hidden channels, rhat of the batch from the gathered registers (rows nb..npad-1 zero)
            if (eb != e0) gather(eb, e1);  // later batches of a high-degree receiver: loaded here
            {
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int q = warp + NW * u;
                    if (q >= npad) break;
                    float* f = sF + q * LDF;
#pragma unroll
                    for (int m = 0; m < NSC; ++m)
                        if (32 * m + lane < S_IN) f[32 * m + lane] = xs[u][m];
                    sH[q * LDH + lane] = hh[u][0], sH[q * LDH + 32 + lane] = hh[u][1];
                    if (V_IN > 0) {
                        const int w = lane;
                        const float vx = xv[u][0], vy = xv[u][1], vz = xv[u][2];
                        const float4 r = rh[u];
                        f[S_IN + w] = vx * r.x + vy * r.y + vz * r.z;
                        f[B1 + 0 * 2 * V_IN + w] = vx * kInvSqrt3;
                        f[B1 + 1 * 2 * V_IN + w] = vy * kInvSqrt3;
                        f[B1 + 2 * 2 * V_IN + w] = vz * kInvSqrt3;
                        f[B1 + 0 * 2 * V_IN + V_IN + w] = (vy * r.z - vz * r.y) * kInvSqrt2;
                        f[B1 + 1 * 2 * V_IN + V_IN + w] = (vz * r.x - vx * r.z) * kInvSqrt2;
                        f[B1 + 2 * 2 * V_IN + V_IN + w] = (vx * r.y - vy * r.x) * kInvSqrt2;
                        if (w == 0) srh[q] = r;
                    }
                }
            }
            if (eb == e0) umma::mbar_wait(bar, phase);
            __syncthreads();
            if (eb + NB >= e1) gather(e0n, e1n);  // registers are free again: the next receiver's first batch flies under the MMAs
            switch (ntiles) {
                case 1: edge_mma_tasks<S_IN, V_IN, 1>(sA, sF, sH, sDF, sP, warp, g, tig); break;
                case 2: edge_mma_tasks<S_IN, V_IN, 2>(sA, sF, sH, sDF, sP, warp, g, tig); break;
                case 3: edge_mma_tasks<S_IN, V_IN, 3>(sA, sF, sH, sDF, sP, warp, g, tig); break;
                default: edge_mma_tasks<S_IN, V_IN, 4>(sA, sF, sH, sDF, sP, warp, g, tig); break;
            }
            __syncthreads();
            // ---- dh = sum of the four K slices (ascending), dxe = J^T df: a warp per edge
            for (int q = warp; q < nb; q += NW) {
                const float* d = sDF + q * LDF;
                float* o = dxe + (size_t)(eb + q) * D_IN;
#pragma unroll
                for (int hv = 0; hv < 2; ++hv) {
                    const float* P = sP + (32 * hv + lane) * LDP + q;
                    dh[(size_t)(eb + q) * JAMUN_EDGE_HID + 32 * hv + lane] = ((P[0] + P[64 * LDP]) + P[2 * 64 * LDP]) + P[3 * 64 * LDP];
                }
                for (int c = lane; c < S_IN; c += 32) o[c] = d[c];
                if (V_IN > 0) {
                    const int w = lane;
                    const float4 rh = srh[q];
                    const float dd = d[S_IN + w];
                    const float cx = d[B1 + 0 * 2 * V_IN + V_IN + w], cy = d[B1 + 1 * 2 * V_IN + V_IN + w],
                                cz = d[B1 + 2 * 2 * V_IN + V_IN + w];
                    // cross = x_v x rhat  =>  d x_v = rhat x d cross
                    o[S_IN + w] = dd * rh.x + d[B1 + 0 * 2 * V_IN + w] * kInvSqrt3 + (rh.y * cz - rh.z * cy) * kInvSqrt2;
                    o[S_IN + V_IN + w] = dd * rh.y + d[B1 + 1 * 2 * V_IN + w] * kInvSqrt3 + (rh.z * cx - rh.x * cz) * kInvSqrt2;
                    o[S_IN + 2 * V_IN + w] = dd * rh.z + d[B1 + 2 * 2 * V_IN + w] * kInvSqrt3 + (rh.x * cy - rh.y * cx) * kInvSqrt2;
                }
            }
            __syncthreads();  // sF / sH / sDF / sP (and, after the last batch, sA) are free again
        }
        phase ^= 1;
        e0 = e0n, e1 = e1n;
    }
}

// ---- path 0e(x)1e->1e, source-major ---------------------------------------------------------------------------------------------
// One warp per source j (lane = output channel w): Y_j [65][32] in registers; for every out-edge e (receiver i):
//   dT[w] = sum_c rhat_e[c] G1_i[c][w];  dY[k'][w] += h'_e[k'] dT[w];  dh_e[k'] += sum_w Y[k'][w] dT[w]   (k' < 64)
// dY_j is written as a stage-major, chunk-swizzled GEMM operand ([65][rows_pad][32]).
constexpr int kP2Warps = 4;
__global__ void __launch_bounds__(32 * kP2Warps)
conv_bwd_p2_kernel(const int* __restrict__ src_rowptr, const int* __restrict__ src_eid, const int* __restrict__ edst,
                   const float* __restrict__ h, const float* __restrict__ rhat, const float* __restrict__ y, int y_ld,
                   const float* __restrict__ g, int N, int rows_pad, float* __restrict__ dh, float* __restrict__ dy_op) {
    __shared__ float prod[kP2Warps][64][33];
    __shared__ float hs[kP2Warps][64];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int j = blockIdx.x * kP2Warps + wib;
    if (j >= N) return;
    const int s0 = src_rowptr[j], s1 = src_rowptr[j + 1];
    float yr[65], dy[65];
#pragma unroll
    for (int k = 0; k < 65; ++k) {
        yr[k] = y[(size_t)j * y_ld + k * JAMUN_V + lane];
        dy[k] = 0.f;
    }
    for (int q = s0; q < s1; ++q) {
        const int e = src_eid[q];
        const int i = edst[e];
        const float4 rh = *reinterpret_cast<const float4*>(rhat + 4 * (size_t)e);
        const float* gi = g + (size_t)i * JAMUN_GATE_IN + JAMUN_S + JAMUN_V;
        const float dT = rh.x * gi[lane] + rh.y * gi[JAMUN_V + lane] + rh.z * gi[2 * JAMUN_V + lane];
        hs[wib][lane] = h[(size_t)e * JAMUN_EDGE_HID + lane];
        hs[wib][32 + lane] = h[(size_t)e * JAMUN_EDGE_HID + 32 + lane];
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 64; ++k) {
            dy[k] = fmaf(hs[wib][k], dT, dy[k]);
            prod[wib][k][lane] = yr[k] * dT;
        }
        dy[64] += dT;
        __syncwarp();
        // lane sums rows lane and lane + 32 of the product tile (ascending w: fixed order)
        float a0 = 0.f, a1 = 0.f;
#pragma unroll 8
        for (int w = 0; w < 32; ++w) {
            a0 += prod[wib][lane][w];
            a1 += prod[wib][32 + lane][w];
        }
        float* dhe = dh + (size_t)e * JAMUN_EDGE_HID;
        dhe[lane] += a0;
        dhe[32 + lane] += a1;
        __syncwarp();
    }
    const int pos = (((lane >> 2) ^ (j & 7)) << 2) | (lane & 3);
#pragma unroll
    for (int k = 0; k < 65; ++k) dy_op[((size_t)k * rows_pad + j) * 32 + pos] = dy[k];
}

// ---- the same on the warp-level tensor cores (default) ---------------------------------------------------------------------------
// Per source j with out-edges q = 0..n-1 (receiver i_q) both sums are small dense products over the source's out-edges:
//   DH [n x 64]  = DT [n x 32] . Y_j^T [32 x 64]        (dh'_e += ...)         M = edges, N = channels, K = 32
//   DY [64 x 32] = H^T [64 x n] . DT [n x 32]           (+ bias row: column sums of DT)   M = channels, N = 32, K = edges
// as mma.sync.m16n8k8 (tf32, three-product split).  One warp per source; DT (<= 32 edges per chunk) and Y_j live in per-warp
// shared memory, the H^T fragments are read straight from h (each element is used once).  Same fixed summation order for every
// run: deterministic.
constexpr int kP2mWarps = 4, kP2mLdY = 36, kP2mLdT = 40;
__global__ void __launch_bounds__(32 * kP2mWarps, 3)
conv_bwd_p2_mma_kernel(const int* __restrict__ src_rowptr, const int* __restrict__ src_eid, const int* __restrict__ edst,
                       const float* __restrict__ h, const float* __restrict__ rhat, const float* __restrict__ y, int y_ld,
                       const float* __restrict__ gm, int N, int rows_pad, float* __restrict__ dh, float* __restrict__ dy_op) {
    extern __shared__ __align__(16) float sm[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, g = lane >> 2, tig = lane & 3;
    float* sY = sm + wib * (64 * kP2mLdY + 32 * kP2mLdT + 32);  // [64][36]
    float* sT = sY + 64 * kP2mLdY;                              // [32][40]  DT, rows >= n zero
    int* sE = reinterpret_cast<int*>(sT + 32 * kP2mLdT);        // [32]      edge ids of the chunk (-1 beyond n)
    const int j = blockIdx.x * kP2mWarps + wib;
    if (j >= N) return;
    const int s0 = src_rowptr[j], s1 = src_rowptr[j + 1];
    // Y_j (rows 0..63) -> shared memory
#pragma unroll 8
    for (int k = 0; k < 64; ++k) sY[k * kP2mLdY + lane] = y[(size_t)j * y_ld + k * JAMUN_V + lane];
    float accB[4][4][4];
#pragma unroll
    for (int mt = 0; mt < 4; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int r = 0; r < 4; ++r) accB[mt][nt][r] = 0.f;
    float bsum = 0.f;
    for (int c0 = s0; c0 < s1; c0 += 32) {
        const int n = s1 - c0 < 32 ? s1 - c0 : 32;
        __syncwarp();
        const int e_mine = lane < n ? src_eid[c0 + lane] : -1;
        sE[lane] = e_mine;
        // DT rows: lane = w (rows n .. 8 ceil(n/8) - 1 zero; rows beyond hold stale finite values that only reach discarded outputs)
        const int qpad = ((n + 7) >> 3) << 3;
#pragma unroll 4
        for (int q = 0; q < qpad; ++q) {
            const int e = __shfl_sync(0xffffffffu, e_mine, q);
            float dT = 0.f;
            if (e >= 0) {
                const float4 rh = *reinterpret_cast<const float4*>(rhat + 4 * (size_t)e);
                const float* gi = gm + (size_t)edst[e] * JAMUN_GATE_IN + JAMUN_S + JAMUN_V;
                dT = rh.x * gi[lane] + rh.y * gi[JAMUN_V + lane] + rh.z * gi[2 * JAMUN_V + lane];
            }
            sT[q * kP2mLdT + lane] = dT;
            bsum += dT;
        }
        __syncwarp();
        const int mtiles = (n + 15) >> 4, ksteps = (n + 7) >> 3;
        // ---- DH = DT . Y^T, added to dh
        for (int mt = 0; mt < mtiles; ++mt) {
            float acc[8][4];
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
            const float* Ar = sT + (16 * mt + g) * kP2mLdT + tig;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                uint32_t ahi[4], alo[4];
                split_tf32(Ar[8 * ks], ahi[0], alo[0]);
                split_tf32(Ar[8 * ks + 8 * kP2mLdT], ahi[1], alo[1]);
                split_tf32(Ar[8 * ks + 4], ahi[2], alo[2]);
                split_tf32(Ar[8 * ks + 8 * kP2mLdT + 4], ahi[3], alo[3]);
#pragma unroll
                for (int nt = 0; nt < 8; ++nt) {
                    const float* Br = sY + (8 * nt + g) * kP2mLdY + 8 * ks + tig;
                    uint32_t bh0, bl0, bh1, bl1;
                    split_tf32(Br[0], bh0, bl0);
                    split_tf32(Br[4], bh1, bl1);
                    mma_tf32(acc[nt], alo, bh0, bh1);
                    mma_tf32(acc[nt], ahi, bl0, bl1);
                    mma_tf32(acc[nt], ahi, bh0, bh1);
                }
            }
            const int ea = sE[16 * mt + g], eb = sE[16 * mt + g + 8];
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                if (ea >= 0) {
                    float2* o = reinterpret_cast<float2*>(dh + (size_t)ea * JAMUN_EDGE_HID + 8 * nt + 2 * tig);
                    float2 v = *o;
                    v.x += acc[nt][0], v.y += acc[nt][1];
                    *o = v;
                }
                if (eb >= 0) {
                    float2* o = reinterpret_cast<float2*>(dh + (size_t)eb * JAMUN_EDGE_HID + 8 * nt + 2 * tig);
                    float2 v = *o;
                    v.x += acc[nt][2], v.y += acc[nt][3];
                    *o = v;
                }
            }
        }
        // ---- DY += H^T . DT
        for (int ks = 0; ks < ksteps; ++ks) {
            const int qa = sE[8 * ks + tig], qb = sE[8 * ks + tig + 4];
            uint32_t bh[4][2], bl[4][2];
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                split_tf32(sT[(8 * ks + tig) * kP2mLdT + 8 * nt + g], bh[nt][0], bl[nt][0]);
                split_tf32(sT[(8 * ks + tig + 4) * kP2mLdT + 8 * nt + g], bh[nt][1], bl[nt][1]);
            }
            const float* ha = h + (size_t)(qa >= 0 ? qa : 0) * JAMUN_EDGE_HID + g;
            const float* hb = h + (size_t)(qb >= 0 ? qb : 0) * JAMUN_EDGE_HID + g;
#pragma unroll
            for (int mt = 0; mt < 4; ++mt) {
                // A[row = channel][col = edge] = h[edge][channel]
                uint32_t ahi[4], alo[4];
                split_tf32(qa >= 0 ? ha[16 * mt] : 0.f, ahi[0], alo[0]);
                split_tf32(qa >= 0 ? ha[16 * mt + 8] : 0.f, ahi[1], alo[1]);
                split_tf32(qb >= 0 ? hb[16 * mt] : 0.f, ahi[2], alo[2]);
                split_tf32(qb >= 0 ? hb[16 * mt + 8] : 0.f, ahi[3], alo[3]);
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) {
                    mma_tf32(accB[mt][nt], alo, bh[nt][0], bh[nt][1]);
                    mma_tf32(accB[mt][nt], ahi, bl[nt][0], bl[nt][1]);
                    mma_tf32(accB[mt][nt], ahi, bh[nt][0], bh[nt][1]);
                }
            }
        }
    }
    // dY_j as a stage-major, chunk-swizzled GEMM operand ([65][rows_pad][32])
#pragma unroll
    for (int mt = 0; mt < 4; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            const int w = 8 * nt + 2 * tig;
            const int pos = (((w >> 2) ^ (j & 7)) << 2) | (w & 3);
            const int ka = 16 * mt + g, kb = ka + 8;
            *reinterpret_cast<float2*>(dy_op + ((size_t)ka * rows_pad + j) * 32 + pos) = make_float2(accB[mt][nt][0], accB[mt][nt][1]);
            *reinterpret_cast<float2*>(dy_op + ((size_t)kb * rows_pad + j) * 32 + pos) = make_float2(accB[mt][nt][2], accB[mt][nt][3]);
        }
    dy_op[((size_t)64 * rows_pad + j) * 32 + ((((lane >> 2) ^ (j & 7)) << 2) | (lane & 3))] = bsum;
}

// ---- dx_j = sum over the out-edges of j (ascending edge id) of dxe[e]  (+ extra[j, :n_extra]) ----------------------------------
__global__ void __launch_bounds__(256)
conv_bwd_gather_kernel(const int* __restrict__ src_rowptr, const int* __restrict__ src_eid, const float* __restrict__ dxe, int D,
                       const float* __restrict__ extra, int extra_ld, int n_extra, int N, float* __restrict__ dx) {
    const int lane = threadIdx.x & 31;
    const int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (j >= N) return;
    const int s0 = src_rowptr[j], s1 = src_rowptr[j + 1];
    for (int c0 = 0; c0 < D; c0 += 32) {
        const int c = c0 + lane;
        if (c >= D) break;
        float acc = (extra && c < n_extra) ? extra[(size_t)j * extra_ld + c] : 0.f;
        for (int q = s0; q < s1; ++q) acc += dxe[(size_t)src_eid[q] * D + c];
        dx[(size_t)j * D + c] = acc;
    }
}

template <int S_IN, int V_IN>
int launch_edge(const float* x, const int* rowptr, const int* col, const float* h, const float* rhat, const float* dA0, int ld0,
                const float* dA1, int ld1, long long comp, int N, float* dh, float* dxe, cudaStream_t s) {
    constexpr int NS = (S_IN + 31) / 32;
    constexpr int W0 = (NS + (V_IN > 0 ? 1 : 0)) * 32, W1 = V_IN > 0 ? 64 : 0, NF = W0 + 3 * W1;
    constexpr size_t smem = (size_t)(65 * W0 + 3 * 65 * W1 + 4 * (2 * NF + 65 + 4)) * sizeof(float);
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(conv_bwd_edge_kernel<S_IN, V_IN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) {
            jb::set_error("jamun_conv_bwd_edge: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
            return JAMUN_ECUDA;
        }
        attr_set = true;
    }
    conv_bwd_edge_kernel<S_IN, V_IN><<<N, 384, smem, s>>>(x, rowptr, col, h, rhat, dA0, ld0, dA1, ld1, comp, N, dh, dxe);
    return JAMUN_OK;
}

template <int S_IN, int V_IN>
int launch_edge_mma(const float* x, const int* rowptr, const int* col, const float* h, const float* rhat, const float* dA0, int ld0,
                    const float* dA1, int ld1, long long comp, int N, float* dh, float* dxe, cudaStream_t s) {
    using C = EdgeMmaCfg<S_IN, V_IN>;
    constexpr size_t smem = C::kSmemFloats * sizeof(float);
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(conv_bwd_edge_mma_kernel<S_IN, V_IN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) {
            jb::set_error("jamun_conv_bwd_edge: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
            return JAMUN_ECUDA;
        }
        attr_set = true;
    }
    const int grid = N < jb::kNumSMs ? N : jb::kNumSMs;
    conv_bwd_edge_mma_kernel<S_IN, V_IN><<<grid, C::kThreads, smem, s>>>(x, rowptr, col, h, rhat, dA0, ld0, dA1, ld1, comp, N, dh, dxe);
    return JAMUN_OK;
}

// JAMUN_B200_BWD_EDGE=simt selects the CUDA-core kernel (A/B reference); default: warp-level tensor cores
bool edge_use_mma() {
    static const bool v = [] {
        const char* e = getenv("JAMUN_B200_BWD_EDGE");
        return !(e && strcmp(e, "simt") == 0);
    }();
    return v;
}

}  // namespace

extern "C" int jamun_conv_bwd_scale(const float* dout, const float* inv_deg, float alpha0, float alpha1, int N, float* g,
                                    jamun_stream_t stream) {
    JB_CHECK_ARG(dout && inv_deg && g, "null argument");
    if (N == 0) return JAMUN_OK;
    size_t total = (size_t)N * JAMUN_GATE_IN;
    int blocks = (int)((total + 255) / 256);
    if (blocks > jb::kNumSMs * 16) blocks = jb::kNumSMs * 16;
    conv_bwd_scale_kernel<<<blocks, 256, 0, jb::as_stream(stream)>>>(dout, inv_deg, alpha0, alpha1, N, g);
    JB_CHECK_LAUNCH();
    return JAMUN_OK;
}

extern "C" int jamun_stage_atb(const float* a, long long a_comp_stride, int ncomp, int n_stages, int nslots, int rows, int rows_pad,
                               const float* b, int ldb, int b_col0, int b_comp_stride, int W, float* out, int mode, int out_rows,
                               const int* slot_row0, const int* slot_rows, jamun_stream_t stream) {
    JB_CHECK_ARG(a && b && out, "null argument");
    JB_CHECK_ARG(W >= 1 && W <= 160 && nslots >= 1 && nslots <= 8 && n_stages % nslots == 0 && ncomp >= 1, "bad shape");
    JB_CHECK_ARG(mode == 1 || (slot_row0 && slot_rows), "mode 0 needs the slot tables (host pointers)");
    if (n_stages == 0) return JAMUN_OK;
    AtbParams P{};
    P.a = a, P.a_comp_stride = a_comp_stride, P.ncomp = ncomp, P.rows = rows, P.rows_pad = rows_pad, P.nslots = nslots;
    P.b = b, P.ldb = ldb, P.b_col0 = b_col0, P.b_comp_stride = b_comp_stride, P.W = W;
    P.out = out, P.mode = mode, P.out_rows = out_rows;
    for (int s = 0; s < nslots && mode == 0; ++s) P.slot_row0[s] = slot_row0[s], P.slot_rows[s] = slot_rows[s];
    stage_atb_kernel<<<n_stages, 256, 0, jb::as_stream(stream)>>>(P);
    JB_CHECK_LAUNCH();
    return JAMUN_OK;
}

extern "C" int jamun_conv_bwd_edge(const float* x, int s_in, int v_in, const int* rowptr, const int* col, const float* h,
                                   const float* rhat, const float* dA0, int ld0, const float* dA1, int ld1,
                                   long long dA1_comp_stride, int N, float* dh, float* dxe, jamun_stream_t stream) {
    JB_CHECK_ARG(x && rowptr && col && h && rhat && dA0 && dh && dxe, "null argument");
    if (N == 0) return JAMUN_OK;
    cudaStream_t s = jb::as_stream(stream);
    int rc;
    if (s_in == JAMUN_S && v_in == JAMUN_V) {
        JB_CHECK_ARG(dA1, "dA1 required for vector inputs");
        JB_CHECK_ARG(ld0 % 4 == 0 && ld1 % 4 == 0 && dA1_comp_stride % 4 == 0, "dA rows must be 16-byte aligned");
        rc = edge_use_mma() ? launch_edge_mma<JAMUN_S, JAMUN_V>(x, rowptr, col, h, rhat, dA0, ld0, dA1, ld1, dA1_comp_stride, N, dh, dxe, s)
                            : launch_edge<JAMUN_S, JAMUN_V>(x, rowptr, col, h, rhat, dA0, ld0, dA1, ld1, dA1_comp_stride, N, dh, dxe, s);
    } else if (s_in == JAMUN_S0 && v_in == 0) {
        JB_CHECK_ARG(ld0 % 4 == 0, "dA rows must be 16-byte aligned");
        rc = edge_use_mma() ? launch_edge_mma<JAMUN_S0, 0>(x, rowptr, col, h, rhat, dA0, ld0, nullptr, 0, 0, N, dh, dxe, s)
                            : launch_edge<JAMUN_S0, 0>(x, rowptr, col, h, rhat, dA0, ld0, nullptr, 0, 0, N, dh, dxe, s);
    } else {
        jb::set_error("jamun_conv_bwd_edge: unsupported input irreps %dx0e+%dx1e", s_in, v_in);
        return JAMUN_EINVAL;
    }
    if (rc != JAMUN_OK) return rc;
    JB_CHECK_LAUNCH();
    return JAMUN_OK;
}

extern "C" int jamun_conv_bwd_p2(const int* src_rowptr, const int* src_eid, const int* edst, const float* h, const float* rhat,
                                 const float* y, int y_ld, const float* g, int N, int rows_pad, float* dh, float* dy_op,
                                 jamun_stream_t stream) {
    JB_CHECK_ARG(src_rowptr && src_eid && edst && h && rhat && y && g && dh && dy_op, "null argument");
    JB_CHECK_ARG(N <= rows_pad, "N exceeds rows_pad");
    if (N == 0) return JAMUN_OK;
    static const bool use_mma = [] {
        const char* e = getenv("JAMUN_B200_BWD_P2");  // "simt" selects the CUDA-core kernel (A/B reference)
        return !(e && strcmp(e, "simt") == 0);
    }();
    if (use_mma) {
        constexpr size_t smem = (size_t)kP2mWarps * (64 * kP2mLdY + 32 * kP2mLdT + 32) * sizeof(float);
        static bool attr_set = false;
        if (!attr_set) {
            cudaError_t e = cudaFuncSetAttribute(conv_bwd_p2_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) {
                jb::set_error("jamun_conv_bwd_p2: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
                return JAMUN_ECUDA;
            }
            attr_set = true;
        }
        conv_bwd_p2_mma_kernel<<<(N + kP2mWarps - 1) / kP2mWarps, 32 * kP2mWarps, smem, jb::as_stream(stream)>>>(
            src_rowptr, src_eid, edst, h, rhat, y, y_ld, g, N, rows_pad, dh, dy_op);
    } else {
        conv_bwd_p2_kernel<<<(N + kP2Warps - 1) / kP2Warps, 32 * kP2Warps, 0, jb::as_stream(stream)>>>(src_rowptr, src_eid, edst, h, rhat,
                                                                                                       y, y_ld, g, N, rows_pad, dh, dy_op);
    }
    JB_CHECK_LAUNCH();
    return JAMUN_OK;
}

extern "C" int jamun_conv_bwd_gather(const int* src_rowptr, const int* src_eid, const float* dxe, int D, const float* extra,
                                     int extra_ld, int n_extra, int N, float* dx, jamun_stream_t stream) {
    JB_CHECK_ARG(src_rowptr && src_eid && dxe && dx, "null argument");
    if (N == 0) return JAMUN_OK;
    const int blocks = (int)(((size_t)N * 32 + 255) / 256);
    conv_bwd_gather_kernel<<<blocks, 256, 0, jb::as_stream(stream)>>>(src_rowptr, src_eid, dxe, D, extra, extra_ld, n_extra, N, dx);
    JB_CHECK_LAUNCH();
    return JAMUN_OK;
}
