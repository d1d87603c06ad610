// Aggregate step of the equivariant convolution for the tensor-core path: per receiver node i
//   A_i[k', u'] = sum_{e -> i} h'_e[k'] * f_e[u']          (h' = [radial hidden (64), 1];  f = edge features)
// written as the fp32 A operand of jamun_gemm_tf32x3 in its stage-major layout.
//
// One warp per node; lanes run over a 32-wide feature slot (coalesced 128-byte gathers of the source row and
// coalesced 128-byte stores of the operand row); a block of 4 radial channels is accumulated in registers per pass
// over the node's in-edges, so each gathered source row feeds 4 x 23 FMAs.
//
// Operand layout (DESIGN.md "conv operand layout"), NS = ceil(S_IN/32) scalar slots:
//   segment 0 (0e, nslots0 = NS + [V>0]):  slot s<NS : x_s[32s+lane]          slot NS   : x_v . rhat
//   segment 1+c (1e, nslots1 = NS + 2[V>0]): slot s<NS : x_s[32s+lane] rhat_c   slot NS   : x_v[c] / sqrt3
//                                                                               slot NS+1 : (x_v x rhat)[c] / sqrt2
//   stage index = k' * nslots + slot;  element (stage, row, lane) at ((stage * rows_pad) + row) * 32 + swz(lane, row),
//   swz = (((lane / 4) ^ (row % 8)) * 4) + lane % 4  (so the GEMM's per-row shared-memory reads are conflict-free).
#include "common.cuh"

namespace {
using namespace jb;

constexpr int RK = 4;
constexpr float kInvSqrt3 = 0.57735026918962576451f;
constexpr float kInvSqrt2 = 0.70710678118654752440f;

template <int S_IN, int V_IN>
__global__ void __launch_bounds__(256)
conv_build_kernel(const float* __restrict__ x, const int* __restrict__ rowptr, const int* __restrict__ col,
                  const float* __restrict__ h, const float* __restrict__ rhat, int row0, int nrows, int rows_pad,
                  float* __restrict__ a0, float* __restrict__ a1, size_t a1_comp_stride, float* __restrict__ inv_deg) {
    constexpr int D_IN = S_IN + 3 * V_IN;
    constexpr int NS = (S_IN + 31) / 32;
    constexpr int NV0 = V_IN > 0 ? 1 : 0, NV1 = V_IN > 0 ? 2 : 0;
    constexpr int NSL0 = NS + NV0, NSL1 = NS + NV1;
    const int lane = threadIdx.x & 31;
    const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;  // row within the chunk
    if (r >= nrows) return;
    const int i = row0 + r;
    const int swz = (((lane >> 2) ^ (r & 7)) << 2) | (lane & 3);  // 16-byte chunks XOR-swizzled with (row & 7)
    const int e0 = rowptr[i], e1 = rowptr[i + 1];
    if (lane == 0) inv_deg[i] = 1.0f / (float)(e1 > e0 ? e1 - e0 : 1);

    for (int kb = 0; kb <= JAMUN_EDGE_HID / RK; ++kb) {
        const bool bias = kb == JAMUN_EDGE_HID / RK;
        float s0[RK][NS], s1[RK][3][NS];
        float aq[RK], av[RK][3], ax[RK][3];
#pragma unroll
        for (int k = 0; k < RK; ++k) {
#pragma unroll
            for (int s = 0; s < NS; ++s) s0[k][s] = s1[k][0][s] = s1[k][1][s] = s1[k][2][s] = 0.f;
            aq[k] = 0.f;
#pragma unroll
            for (int c = 0; c < 3; ++c) av[k][c] = ax[k][c] = 0.f;
        }
        for (int e = e0; e < e1; ++e) {
            const int j = col[e];
            const float4 rh = *reinterpret_cast<const float4*>(rhat + 4 * (size_t)e);
            float4 hq4 = make_float4(1.f, 0.f, 0.f, 0.f);
            if (!bias) hq4 = *reinterpret_cast<const float4*>(h + (size_t)e * JAMUN_EDGE_HID + kb * RK);
            const float hq[RK] = {hq4.x, hq4.y, hq4.z, hq4.w};
            const float* xj = x + (size_t)j * D_IN;
            float xs[NS];
#pragma unroll
            for (int s = 0; s < NS; ++s) xs[s] = (lane + 32 * s < S_IN) ? xj[lane + 32 * s] : 0.f;
            float vx = 0.f, vy = 0.f, vz = 0.f, q = 0.f, cx = 0.f, cy = 0.f, cz = 0.f;
            if (V_IN > 0) {
                vx = xj[S_IN + lane];
                vy = xj[S_IN + V_IN + lane];
                vz = xj[S_IN + 2 * V_IN + lane];
                q = vx * rh.x + vy * rh.y + vz * rh.z;
                cx = vy * rh.z - vz * rh.y;
                cy = vz * rh.x - vx * rh.z;
                cz = vx * rh.y - vy * rh.x;
            }
#pragma unroll
            for (int k = 0; k < RK; ++k) {
#pragma unroll
                for (int s = 0; s < NS; ++s) {
                    const float t = hq[k] * xs[s];
                    s0[k][s] += t;
                    s1[k][0][s] = fmaf(t, rh.x, s1[k][0][s]);
                    s1[k][1][s] = fmaf(t, rh.y, s1[k][1][s]);
                    s1[k][2][s] = fmaf(t, rh.z, s1[k][2][s]);
                }
                if (V_IN > 0) {
                    aq[k] = fmaf(hq[k], q, aq[k]);
                    av[k][0] = fmaf(hq[k], vx, av[k][0]);
                    av[k][1] = fmaf(hq[k], vy, av[k][1]);
                    av[k][2] = fmaf(hq[k], vz, av[k][2]);
                    ax[k][0] = fmaf(hq[k], cx, ax[k][0]);
                    ax[k][1] = fmaf(hq[k], cy, ax[k][1]);
                    ax[k][2] = fmaf(hq[k], cz, ax[k][2]);
                }
            }
        }
        const int nk = bias ? 1 : RK;
#pragma unroll
        for (int k = 0; k < RK; ++k) {
            if (k >= nk) break;
            const int kp = kb * RK + k;
            float* p0 = a0 + ((size_t)(kp * NSL0) * rows_pad + r) * 32 + swz;
#pragma unroll
            for (int s = 0; s < NS; ++s) __stcs(p0 + (size_t)s * rows_pad * 32, s0[k][s]);
            if (V_IN > 0) __stcs(p0 + (size_t)NS * rows_pad * 32, aq[k]);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                float* p1 = a1 + c * a1_comp_stride + ((size_t)(kp * NSL1) * rows_pad + r) * 32 + swz;
#pragma unroll
                for (int s = 0; s < NS; ++s) __stcs(p1 + (size_t)s * rows_pad * 32, s1[k][c][s]);
                if (V_IN > 0) {
                    __stcs(p1 + (size_t)NS * rows_pad * 32, av[k][c] * kInvSqrt3);
                    __stcs(p1 + (size_t)(NS + 1) * rows_pad * 32, ax[k][c] * kInvSqrt2);
                }
            }
        }
    }
}

}  // namespace

// a0: [65*nslots0][rows_pad][32], a1: 3 x [65*nslots1][rows_pad][32] (component stride a1_comp_stride floats).
extern "C" int jamun_conv_build_a(const float* x, int s_in, int v_in, const int* rowptr, const int* col, const float* h,
                                  const float* rhat, int row0, int nrows, int rows_pad, float* a0, float* a1,
                                  long long a1_comp_stride, float* inv_deg, jamun_stream_t stream) {
    JB_CHECK_ARG(x && rowptr && col && h && rhat && a0 && a1 && inv_deg, "null argument");
    JB_CHECK_ARG(nrows <= rows_pad, "nrows exceeds rows_pad");
    if (nrows == 0) return JAMUN_OK;
    const int blocks = (nrows * 32 + 255) / 256;
    cudaStream_t s = jb::as_stream(stream);
    if (s_in == JAMUN_S && v_in == JAMUN_V) {
        conv_build_kernel<JAMUN_S, JAMUN_V><<<blocks, 256, 0, s>>>(x, rowptr, col, h, rhat, row0, nrows, rows_pad, a0, a1,
                                                                  (size_t)a1_comp_stride, inv_deg);
    } else if (s_in == JAMUN_S0 && v_in == 0) {
        conv_build_kernel<JAMUN_S0, 0><<<blocks, 256, 0, s>>>(x, rowptr, col, h, rhat, row0, nrows, rows_pad, a0, a1,
                                                             (size_t)a1_comp_stride, inv_deg);
    } else {
        jb::set_error("jamun_conv_build_a: unsupported input irreps %dx0e+%dx1e", s_in, v_in);
        return JAMUN_EINVAL;
    }
    JB_CHECK_LAUNCH();
    return JAMUN_OK;
}
