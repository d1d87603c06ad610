// Aggregate step of the equivariant convolution for the tensor-core path.  Per receiver node i
//   A_i[k', u'] = sum_{e -> i} h'_e[k'] * f_e[u']          (h' = [radial hidden (64), 1];  f = edge features)
// is written as the fp32 A operand of jamun_gemm_tf32x3 in its stage-major layout, for every tensor-product path whose
// edge feature is cheap to aggregate: 0e(x)0e->0e, 1e(x)1e->0e, 1e(x)0e->1e, 1e(x)1e->1e.
//
// The remaining path 0e(x)1e->1e (x_s * rhat_c, 120 scalars x 3 components: half of all aggregate FLOPs and operand
// bytes if treated the same way) is evaluated transform-then-aggregate instead: a per-node GEMM first forms
// Y_j[k', w] = sum_u M1[k', u, w] x_s,j[u] (jamun_gemm_tf32x3 with 13 column blocks), and this kernel gathers
//   p2_i[c, w] = sum_{e -> i} rhat_e[c] * sum_k' h'_e[k'] * Y_j(e)[k', w]
// with lanes over w -- 2 kFMA per edge instead of 23 k.  The GEMM epilogue adds p2 to the 1e accumulators.
//
// One warp per node.  Scalars: lane L owns x_s[4L..4L+3] (one 16-byte gather per edge, one 16-byte operand store per
// channel); vectors: lane owns multiplicity `lane`; RK = 4 radial channels are accumulated in registers per pass over the
// in-edges with packed fp32 FMAs (FFMA2).  Per-edge invariants (row offsets, rhat) are cached in shared memory per warp.
//
// Operand layout, NS = ceil(S_IN/32) scalar slots:
//   segment 0 (0e, nslots0 = NS + [V>0]):   slot s<NS : x_s[32s .. 32s+31]   slot NS : x_v . rhat
//   segment 1+c (1e, nslots1 = 2[V>0]):     slot 0 : x_v[c] / sqrt3          slot 1  : (x_v x rhat)[c] / sqrt2
//   stage index = k' * nslots + slot;  element (stage, row, p) at ((stage * rows_pad) + row) * 32 + swz(p, row),
//   swz = (((p / 4) ^ (row % 8)) * 4) + p % 4  (so the GEMM's per-row shared-memory reads are conflict-free).
#include <stdlib.h>
#include "common.cuh"

namespace {
using namespace jb;

constexpr float kInvSqrt3 = 0.57735026918962576451f;
constexpr float kInvSqrt2 = 0.70710678118654752440f;
constexpr int YLD = 17 * 128;  // row stride of Y: (64+1)*32 = 2080 columns padded to 17 column blocks of 128
constexpr int MAXD = 64;  // in-edges per node whose per-edge invariants are cached in shared memory

template <int S_IN, int V_IN, int RK, int MINB, bool CACHED>
__global__ void __launch_bounds__(256, MINB)
conv_build_kernel(const float* __restrict__ x, const int* __restrict__ rowptr, const int* __restrict__ col,
                  const float* __restrict__ h, const float* __restrict__ rhat, const float* __restrict__ y, int row0,
                  int nrows, int rows_pad, float* __restrict__ a0, float* __restrict__ a1, size_t a1_comp_stride,
                  float* __restrict__ p2, int p2_ld, float p2_scale, float* __restrict__ inv_deg) {
    static_assert(RK == 4 || RK == 8, "RK must be 4 or 8");
    constexpr int NKB = JAMUN_EDGE_HID / RK;  // full blocks; block NKB is the bias channel (h' = 1)
    constexpr int D_IN = S_IN + 3 * V_IN;
    constexpr int NS = (S_IN + 31) / 32;
    constexpr int NSL0 = NS + (V_IN > 0 ? 1 : 0), NSL1 = V_IN > 0 ? 2 : 0;
    // per warp, per in-edge: {x row offset, y row offset, rx, ry}, {rz, -, -, -} (CACHED: every node has <= MAXD in-edges)
    __shared__ float4 meta_s[CACHED ? 8 : 1][CACHED ? MAXD : 1][2];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;  // row within the chunk
    if (r >= nrows) return;
    const int i = row0 + r;
    const int e0 = rowptr[i], e1 = rowptr[i + 1];
    const int deg = e1 - e0;
    const float invd = 1.0f / (float)(deg > 0 ? deg : 1);
    constexpr bool DO_S = true, DO_V = V_IN > 0;
    if (lane == 0 && DO_S) inv_deg[i] = invd;
    if (CACHED) {
        for (int t = lane; t < deg; t += 32) {
            const int j = col[e0 + t];
            const float4 rh = *reinterpret_cast<const float4*>(rhat + 4 * (size_t)(e0 + t));
            meta_s[wib][t][0] = make_float4(__int_as_float(j * D_IN), __int_as_float(j * YLD), rh.x, rh.y);
            meta_s[wib][t][1] = make_float4(rh.z, 0.f, 0.f, 0.f);
        }
        __syncwarp();
    }
    const bool s_live = 4 * lane < NS * 32;
    const bool s_load = 4 * lane < S_IN;
    const int swz = (((lane >> 2) ^ (r & 7)) << 2) | (lane & 3);  // position of a 4-byte element owned by `lane`
    const int swz4 = ((lane & 7) ^ (r & 7)) << 2;                  // position of the 16-byte chunk of the scalar store
    const float* xl = x + (s_load ? 4 * lane : 0);                 // lane-adjusted bases: per edge only `+ offset` remains
    const float* xv = x + S_IN + lane;
    const float* yl = y + lane;
    const float* hrow = h + (size_t)e0 * JAMUN_EDGE_HID;
    float pacc[3] = {0.f, 0.f, 0.f};

    for (int kb = 0; kb <= NKB; ++kb) {
        const bool bias = kb == NKB;
        float s0[RK][4];
        float aq[RK], av[RK][3], ax[RK][3];
#pragma unroll
        for (int k = 0; k < RK; ++k) {
#pragma unroll
            for (int t = 0; t < 4; ++t) s0[k][t] = 0.f;
            aq[k] = 0.f;
#pragma unroll
            for (int c = 0; c < 3; ++c) av[k][c] = ax[k][c] = 0.f;
        }
        const float* hk = hrow + kb * RK;
        const float* yk = yl + kb * RK * JAMUN_V;
#pragma unroll 2
        for (int t = 0; t < deg; ++t, hk += JAMUN_EDGE_HID) {
            int xoff, yoff;
            float rx, ry, rz;
            if (CACHED) {
                const float4 m0 = meta_s[wib][t][0], m1 = meta_s[wib][t][1];
                xoff = __float_as_int(m0.x);
                yoff = __float_as_int(m0.y);
                rx = m0.z;
                ry = m0.w;
                rz = m1.x;
            } else {
                const int j = col[e0 + t];
                const float4 rh = *reinterpret_cast<const float4*>(rhat + 4 * (size_t)(e0 + t));
                xoff = j * D_IN;
                yoff = j * YLD;
                rx = rh.x;
                ry = rh.y;
                rz = rh.z;
            }
            float hq[RK];
            if (!bias) {
#pragma unroll
                for (int q = 0; q < RK / 4; ++q) {
                    const float4 w = *reinterpret_cast<const float4*>(hk + 4 * q);
                    hq[4 * q] = w.x;
                    hq[4 * q + 1] = w.y;
                    hq[4 * q + 2] = w.z;
                    hq[4 * q + 3] = w.w;
                }
            } else {
                hq[0] = 1.f;
#pragma unroll
                for (int k = 1; k < RK; ++k) hq[k] = 0.f;
            }
            float4 xs = make_float4(0.f, 0.f, 0.f, 0.f);
            if (DO_S) {
                if (s_load) xs = *reinterpret_cast<const float4*>(xl + xoff);
                // path 0e(x)1e->1e through the pre-transformed source rows
                const float* yj = yk + yoff;
                float tsum;
                if (!bias) {
                    float ta = hq[0] * yj[0], tb = 0.f;
#pragma unroll
                    for (int k = 1; k + 1 < RK; k += 2) ffma2v(ta, tb, hq[k], hq[k + 1], yj[k * JAMUN_V], yj[(k + 1) * JAMUN_V]);
                    ta = fmaf(hq[RK - 1], yj[(RK - 1) * JAMUN_V], ta);
                    tsum = ta + tb;
                } else {
                    tsum = yj[0];
                }
                ffma2(pacc[0], pacc[1], tsum, rx, ry);
                pacc[2] = fmaf(rz, tsum, pacc[2]);
            }
            float vx = 0.f, vy = 0.f, vz = 0.f, q = 0.f, cx = 0.f, cy = 0.f, cz = 0.f;
            if (DO_V) {
                const float* vj = xv + xoff;
                vx = vj[0];
                vy = vj[V_IN];
                vz = vj[2 * V_IN];
                q = vx * rx + vy * ry + vz * rz;
                cx = vy * rz - vz * ry;
                cy = vz * rx - vx * rz;
                cz = vx * ry - vy * rx;
            }
#pragma unroll
            for (int k = 0; k < RK; ++k) {
                if (DO_S) {
                    ffma2(s0[k][0], s0[k][1], hq[k], xs.x, xs.y);
                    ffma2(s0[k][2], s0[k][3], hq[k], xs.z, xs.w);
                }
                if (DO_V) {
                    ffma2(aq[k], av[k][0], hq[k], q, vx);
                    ffma2(av[k][1], av[k][2], hq[k], vy, vz);
                    ffma2(ax[k][0], ax[k][1], hq[k], cx, cy);
                    ax[k][2] = fmaf(hq[k], cz, ax[k][2]);
                }
            }
        }
        const int nk = bias ? 1 : RK;
#pragma unroll
        for (int k = 0; k < RK; ++k) {
            if (k >= nk) break;
            const int kp = kb * RK + k;
            if (s_live && DO_S) {
                float* ps = a0 + ((size_t)(kp * NSL0 + (lane >> 3)) * rows_pad + r) * 32 + swz4;
                __stcs(reinterpret_cast<float4*>(ps), make_float4(s0[k][0], s0[k][1], s0[k][2], s0[k][3]));
            }
            if (DO_V) {
                __stcs(a0 + ((size_t)(kp * NSL0 + NS) * rows_pad + r) * 32 + swz, aq[k]);
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    float* p1 = a1 + c * a1_comp_stride + ((size_t)(kp * NSL1) * rows_pad + r) * 32 + swz;
                    __stcs(p1, av[k][c] * kInvSqrt3);
                    __stcs(p1 + (size_t)rows_pad * 32, ax[k][c] * kInvSqrt2);
                }
            }
        }
    }
    // path-2 sums: raw (p2_scale == 0; the GEMM epilogue adds and scales them) or final (initial block: no 1e GEMM)
    if (DO_S) {
        const float sc = p2_scale != 0.f ? p2_scale * invd : 1.0f;
#pragma unroll
        for (int c = 0; c < 3; ++c) p2[(size_t)i * p2_ld + c * JAMUN_V + lane] = pacc[c] * sc;
    }
}

// x[rows, ld] columns [col0, col0+ncols) -> stage-major chunk-swizzled A operand, zero padded to 32-column stages
__global__ void pack_rows_kernel(const float* __restrict__ x, int ld, int col0, int ncols, int rows, int rows_pad,
                                 float* __restrict__ a) {
    const int nst = (ncols + 31) / 32;
    const size_t total = (size_t)nst * rows_pad * 32;
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const int lane = (int)(t & 31);
        const size_t sr = t >> 5;
        const int row = (int)(sr % rows_pad), st = (int)(sr / rows_pad);
        const int c = st * 32 + lane;
        const float v = (row < rows && c < ncols) ? x[(size_t)row * ld + col0 + c] : 0.f;
        const int swz = (((lane >> 2) ^ (row & 7)) << 2) | (lane & 3);
        a[sr * 32 + swz] = v;
    }
}

}  // namespace

// a0: [65*nslots0][rows_pad][32]; a1: 3 x [65*2][rows_pad][32] (component stride a1_comp_stride floats; unused when v_in == 0);
// y: [N, 65*32] pre-transformed source rows; p2: [N, p2_ld] path-2 sums (see the kernel for p2_scale).
extern "C" int jamun_conv_build_a(const float* x, int s_in, int v_in, const int* rowptr, const int* col, const float* h,
                                  const float* rhat, const float* y, int max_degree, int row0, int nrows, int rows_pad, float* a0, float* a1,
                                  long long a1_comp_stride, float* p2, int p2_ld, float p2_scale, float* inv_deg,
                                  jamun_stream_t stream) {
    JB_CHECK_ARG(x && rowptr && col && h && rhat && y && a0 && p2 && inv_deg, "null argument");
    JB_CHECK_ARG(nrows <= rows_pad, "nrows exceeds rows_pad");
    JB_CHECK_ARG(((size_t)x & 15) == 0 && ((size_t)a0 & 15) == 0, "x and a0 must be 16-byte aligned");
    if (nrows == 0) return JAMUN_OK;
    const int blocks = (nrows * 32 + 255) / 256;
    cudaStream_t s = jb::as_stream(stream);
#define JB_LAUNCH_BUILD(S_, V_, RK_, MB_, C_)                                                                                 \
    conv_build_kernel<S_, V_, RK_, MB_, C_><<<blocks, 256, 0, s>>>(x, rowptr, col, h, rhat, y, row0, nrows, rows_pad, a0, a1,    \
                                                                   (size_t)a1_comp_stride, p2, p2_ld, p2_scale, inv_deg)
    const bool cached = max_degree <= MAXD;  // per-edge invariants fit the per-warp shared-memory cache
    if (s_in == JAMUN_S && v_in == JAMUN_V) {
        JB_CHECK_ARG(a1, "a1 required for vector inputs");
        if (cached) JB_LAUNCH_BUILD(JAMUN_S, JAMUN_V, 4, 2, true);
        else JB_LAUNCH_BUILD(JAMUN_S, JAMUN_V, 4, 2, false);
    } else if (s_in == JAMUN_S0 && v_in == 0) {
        if (cached) JB_LAUNCH_BUILD(JAMUN_S0, 0, 8, 3, true);
        else JB_LAUNCH_BUILD(JAMUN_S0, 0, 4, 2, false);
    } else {
        jb::set_error("jamun_conv_build_a: unsupported input irreps %dx0e+%dx1e", s_in, v_in);
        return JAMUN_EINVAL;
    }
#undef JB_LAUNCH_BUILD
    JB_CHECK_LAUNCH();
    return JAMUN_OK;
}

extern "C" int jamun_pack_rows(const float* x, int ld, int col0, int ncols, int rows, int rows_pad, float* a,
                               jamun_stream_t stream) {
    JB_CHECK_ARG(x && a && ncols > 0 && rows <= rows_pad && rows_pad % 128 == 0, "bad argument");
    const size_t total = (size_t)((ncols + 31) / 32) * rows_pad * 32;
    int blocks = (int)((total + 255) / 256);
    if (blocks > jb::kNumSMs * 16) blocks = jb::kNumSMs * 16;
    pack_rows_kernel<<<blocks, 256, 0, jb::as_stream(stream)>>>(x, ld, col0, ncols, rows, rows_pad, a);
    JB_CHECK_LAUNCH();
    return JAMUN_OK;
}
