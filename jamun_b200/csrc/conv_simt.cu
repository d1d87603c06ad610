// Node-side kernels, fp32 CUDA-core implementation (first correct path; the tcgen05 version replaces
// the inner contraction of conv_fwd): fused equivariant convolution in aggregate-then-transform form,
// gate + self-interaction + skip tail, output head.
//
// Work decomposition shared by all three kernels: a warp owns 8 consecutive nodes; lanes run over
// feature columns (coalesced gathers / weight reads); the per-node operand rows live in that warp's
// private slice of shared memory, so the main loops need only __syncwarp().
#include "common.cuh"

namespace {
using namespace jb;

constexpr int SO = JAMUN_S + JAMUN_V;  // 152 scalar outputs of a conv (120 scalars + 32 gates)
constexpr int VO = JAMUN_V;            // 32 vector outputs
constexpr int NPW = 8;                 // nodes per warp
constexpr int WARPS = 8;               // warps per CTA
constexpr int TM = NPW * WARPS;        // 64 nodes per CTA
constexpr float kInvSqrt3 = 0.57735026918962576451f;
constexpr float kInvSqrt2 = 0.70710678118654752440f;

template <int S_IN, int V_IN>
struct ConvCfg {
    static constexpr int D_IN = S_IN + 3 * V_IN;
    static constexpr int U0 = S_IN + V_IN;      // K-width of the 0e operand per radial channel
    static constexpr int U1 = S_IN + 2 * V_IN;  // K-width of the 1e operand per radial channel
    static constexpr int NS = (S_IN + 31) / 32; // lane slots covering the scalar inputs
    static constexpr int A0_FLOATS = NPW * U0;
    static constexpr int A1_FLOATS = 3 * NPW * U1;
    static constexpr int WARP_FLOATS = A0_FLOATS + A1_FLOATS;
    static constexpr size_t SMEM = (size_t)WARPS * WARP_FLOATS * sizeof(float);
    static_assert(U0 % 4 == 0 && U1 % 4 == 0, "operand rows must be float4-aligned");
    static_assert(V_IN == 0 || V_IN == 32, "vector multiplicity must be 0 or 32");
};

// out[i] = alpha/deg_i * sum_k' sum_u' A_i[k',u'] M[k',u',:],  A_i[k',u'] = sum_{e->i} h'_e[k'] f_e[u']
// f0 = [x_s, x_v.rhat], f1[c] = [x_s rhat_c, x_v[c]/sqrt3, (x_v x rhat)[c]/sqrt2]   (conv math)
template <int S_IN, int V_IN>
__global__ void __launch_bounds__(WARPS * 32, 1)
conv_simt_kernel(const float* __restrict__ x, const int* __restrict__ rowptr, const int* __restrict__ col,
                 const float* __restrict__ h, const float* __restrict__ rhat, const float* __restrict__ m0,
                 const float* __restrict__ m1, float alpha0, float alpha1, int N, float* __restrict__ out) {
    using C = ConvCfg<S_IN, V_IN>;
    extern __shared__ __align__(16) float smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* A0 = smem + warp * C::WARP_FLOATS;  // [NPW][U0]
    float* A1 = A0 + C::A0_FLOATS;             // [3][NPW][U1]
    const int node0 = blockIdx.x * TM + warp * NPW;
    if (node0 >= N) return;

    float acc0[NPW][5];
    float acc1[NPW][3];
#pragma unroll
    for (int r = 0; r < NPW; ++r) {
#pragma unroll
        for (int j = 0; j < 5; ++j) acc0[r][j] = 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) acc1[r][c] = 0.f;
    }

    for (int kp = 0; kp <= JAMUN_EDGE_HID; ++kp) {
        // ---- build this warp's operand rows for radial channel kp (kp == 64 is the bias row, h' = 1)
        for (int r = 0; r < NPW; ++r) {
            const int i = node0 + r;
            float s0[C::NS], s1[3][C::NS];
            float adot = 0.f, av[3] = {0.f, 0.f, 0.f}, ax[3] = {0.f, 0.f, 0.f};
#pragma unroll
            for (int s = 0; s < C::NS; ++s) s0[s] = s1[0][s] = s1[1][s] = s1[2][s] = 0.f;
            if (i < N) {
                const int e0 = rowptr[i], e1 = rowptr[i + 1];
                for (int e = e0; e < e1; ++e) {
                    const int j = col[e];
                    const float hk = kp < JAMUN_EDGE_HID ? h[(size_t)e * JAMUN_EDGE_HID + kp] : 1.0f;
                    const float4 rh = *reinterpret_cast<const float4*>(rhat + 4 * (size_t)e);
                    const float* xj = x + (size_t)j * C::D_IN;
#pragma unroll
                    for (int s = 0; s < C::NS; ++s) {
                        const int u = lane + 32 * s;
                        const float xs = u < S_IN ? xj[u] : 0.f;
                        const float t = hk * xs;
                        s0[s] += t;
                        s1[0][s] = fmaf(t, rh.x, s1[0][s]);
                        s1[1][s] = fmaf(t, rh.y, s1[1][s]);
                        s1[2][s] = fmaf(t, rh.z, s1[2][s]);
                    }
                    if (V_IN > 0) {
                        const float vx = xj[S_IN + lane], vy = xj[S_IN + V_IN + lane], vz = xj[S_IN + 2 * V_IN + lane];
                        adot = fmaf(hk, vx * rh.x + vy * rh.y + vz * rh.z, adot);
                        av[0] = fmaf(hk, vx, av[0]);
                        av[1] = fmaf(hk, vy, av[1]);
                        av[2] = fmaf(hk, vz, av[2]);
                        ax[0] = fmaf(hk, vy * rh.z - vz * rh.y, ax[0]);
                        ax[1] = fmaf(hk, vz * rh.x - vx * rh.z, ax[1]);
                        ax[2] = fmaf(hk, vx * rh.y - vy * rh.x, ax[2]);
                    }
                }
            }
#pragma unroll
            for (int s = 0; s < C::NS; ++s) {
                const int u = lane + 32 * s;
                if (u < S_IN) {
                    A0[r * C::U0 + u] = s0[s];
#pragma unroll
                    for (int c = 0; c < 3; ++c) A1[(c * NPW + r) * C::U1 + u] = s1[c][s];
                }
            }
            if (V_IN > 0) {
                A0[r * C::U0 + S_IN + lane] = adot;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    A1[(c * NPW + r) * C::U1 + S_IN + lane] = av[c] * kInvSqrt3;
                    A1[(c * NPW + r) * C::U1 + S_IN + V_IN + lane] = ax[c] * kInvSqrt2;
                }
            }
        }
        __syncwarp();

        // ---- contract with the packed radial-layer weights of channel kp
        const float* b0 = m0 + (size_t)kp * C::U0 * SO;
        for (int u = 0; u < C::U0; u += 4) {
            float4 a[NPW];
#pragma unroll
            for (int r = 0; r < NPW; ++r) a[r] = *reinterpret_cast<const float4*>(A0 + r * C::U0 + u);
#pragma unroll
            for (int uu = 0; uu < 4; ++uu) {
                const float* brow = b0 + (size_t)(u + uu) * SO;
                float b[5];
#pragma unroll
                for (int j = 0; j < 5; ++j) b[j] = (lane + 32 * j < SO) ? brow[lane + 32 * j] : 0.f;
#pragma unroll
                for (int r = 0; r < NPW; ++r) {
                    const float ar = uu == 0 ? a[r].x : uu == 1 ? a[r].y : uu == 2 ? a[r].z : a[r].w;
#pragma unroll
                    for (int j = 0; j < 5; ++j) acc0[r][j] = fmaf(ar, b[j], acc0[r][j]);
                }
            }
        }
        const float* b1 = m1 + (size_t)kp * C::U1 * VO;
        for (int u = 0; u < C::U1; u += 4) {
            float bq[4];
#pragma unroll
            for (int uu = 0; uu < 4; ++uu) bq[uu] = b1[(size_t)(u + uu) * VO + lane];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
#pragma unroll
                for (int r = 0; r < NPW; ++r) {
                    const float4 a = *reinterpret_cast<const float4*>(A1 + (c * NPW + r) * C::U1 + u);
                    acc1[r][c] = fmaf(a.x, bq[0], acc1[r][c]);
                    acc1[r][c] = fmaf(a.y, bq[1], acc1[r][c]);
                    acc1[r][c] = fmaf(a.z, bq[2], acc1[r][c]);
                    acc1[r][c] = fmaf(a.w, bq[3], acc1[r][c]);
                }
            }
        }
        __syncwarp();
    }

    // ---- epilogue: mean over in-edges and path normalisation
#pragma unroll
    for (int r = 0; r < NPW; ++r) {
        const int i = node0 + r;
        if (i >= N) break;
        const int deg = rowptr[i + 1] - rowptr[i];
        const float inv = 1.0f / (float)(deg > 0 ? deg : 1);
        float* o = out + (size_t)i * JAMUN_GATE_IN;
#pragma unroll
        for (int j = 0; j < 5; ++j)
            if (lane + 32 * j < SO) o[lane + 32 * j] = acc0[r][j] * alpha0 * inv;
#pragma unroll
        for (int c = 0; c < 3; ++c) o[SO + c * VO + lane] = acc1[r][c] * alpha1 * inv;
    }
}

// ---- tail: Gate -> self-interaction Linear, + skip Linear(x_in), noise-conditional skip / next-layer scaling
template <int S_IN, int V_IN>
__global__ void __launch_bounds__(WARPS * 32, 1)
block_tail_kernel(const float* __restrict__ conv, const float* __restrict__ vadd, const float* __restrict__ x_in,
                  const float* __restrict__ x_res,
                  const float* __restrict__ wself_s, const float* __restrict__ wself_v,
                  const float* __restrict__ wskip_s, const float* __restrict__ wskip_v,
                  const float* __restrict__ skip_w, const float* __restrict__ s_next, float c_act, float c_gate,
                  int N, float* __restrict__ x_new, float* __restrict__ x_scaled) {
    constexpr int D_IN = S_IN + 3 * V_IN;
    constexpr int S = JAMUN_S, V = JAMUN_V, HID = JAMUN_HID;
    constexpr int WARP_FLOATS = NPW * (HID + D_IN);
    extern __shared__ __align__(16) float smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* Gs = smem + warp * WARP_FLOATS;  // [NPW][216] gate output (SoA)
    float* Xs = Gs + NPW * HID;             // [NPW][D_IN] block input (scaled)
    const int node0 = blockIdx.x * TM + warp * NPW;
    if (node0 >= N) return;

    for (int r = 0; r < NPW; ++r) {
        const int i = node0 + r;
        if (i < N) {
            const float* o = conv + (size_t)i * JAMUN_GATE_IN;
            for (int t = lane; t < S; t += 32) {
                float v = o[t];
                Gs[r * HID + t] = c_act * (v > 0.f ? v : 0.01f * v);
            }
            const float gate = c_gate * sigmoidf_acc(o[S + lane]);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                float v = o[SO + c * V + lane];
                if (vadd) v += vadd[(size_t)i * (3 * V) + c * V + lane];  // 0e(x)1e->1e part gathered separately (jamun_conv_p2)
                Gs[r * HID + S + c * V + lane] = v * gate;
            }
            for (int t = lane; t < D_IN; t += 32) Xs[r * D_IN + t] = x_in[(size_t)i * D_IN + t];
        } else {
            for (int t = lane; t < HID; t += 32) Gs[r * HID + t] = 0.f;
            for (int t = lane; t < D_IN; t += 32) Xs[r * D_IN + t] = 0.f;
        }
    }
    __syncwarp();

    float ys[NPW][4], yv[NPW][3];
#pragma unroll
    for (int r = 0; r < NPW; ++r) {
#pragma unroll
        for (int j = 0; j < 4; ++j) ys[r][j] = 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) yv[r][c] = 0.f;
    }
    // scalars: self-interaction over gate scalars, skip over input scalars
    for (int u = 0; u < S; u += 4) {
        float4 a[NPW];
#pragma unroll
        for (int r = 0; r < NPW; ++r) a[r] = *reinterpret_cast<const float4*>(Gs + r * HID + u);
#pragma unroll
        for (int uu = 0; uu < 4; ++uu) {
            float b[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = (lane + 32 * j < S) ? wself_s[(size_t)(u + uu) * S + lane + 32 * j] : 0.f;
#pragma unroll
            for (int r = 0; r < NPW; ++r) {
                const float ar = uu == 0 ? a[r].x : uu == 1 ? a[r].y : uu == 2 ? a[r].z : a[r].w;
#pragma unroll
                for (int j = 0; j < 4; ++j) ys[r][j] = fmaf(ar, b[j], ys[r][j]);
            }
        }
    }
    for (int u = 0; u < S_IN; u += 4) {
        float4 a[NPW];
#pragma unroll
        for (int r = 0; r < NPW; ++r) a[r] = *reinterpret_cast<const float4*>(Xs + r * D_IN + u);
#pragma unroll
        for (int uu = 0; uu < 4; ++uu) {
            float b[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = (lane + 32 * j < S) ? wskip_s[(size_t)(u + uu) * S + lane + 32 * j] : 0.f;
#pragma unroll
            for (int r = 0; r < NPW; ++r) {
                const float ar = uu == 0 ? a[r].x : uu == 1 ? a[r].y : uu == 2 ? a[r].z : a[r].w;
#pragma unroll
                for (int j = 0; j < 4; ++j) ys[r][j] = fmaf(ar, b[j], ys[r][j]);
            }
        }
    }
    // vectors
    for (int u = 0; u < V; ++u) {
        const float b = wself_v[u * V + lane];
#pragma unroll
        for (int r = 0; r < NPW; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) yv[r][c] = fmaf(Gs[r * HID + S + c * V + u], b, yv[r][c]);
    }
    if (V_IN > 0) {
        for (int u = 0; u < V_IN; ++u) {
            const float b = wskip_v[u * V + lane];
#pragma unroll
            for (int r = 0; r < NPW; ++r)
#pragma unroll
                for (int c = 0; c < 3; ++c) yv[r][c] = fmaf(Xs[r * D_IN + S_IN + c * V_IN + u], b, yv[r][c]);
        }
    }

#pragma unroll
    for (int r = 0; r < NPW; ++r) {
        const int i = node0 + r;
        if (i >= N) break;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int w = lane + 32 * j;
            if (w < S) {
                float y = ys[r][j];
                if (skip_w) {
                    const float sw = skip_w[w];
                    y = x_res[(size_t)i * HID + w] * sw + y * (1.0f - sw);
                }
                x_new[(size_t)i * HID + w] = y;
                if (x_scaled) x_scaled[(size_t)i * HID + w] = s_next ? y * s_next[w] : y;
            }
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const int w = S + c * V + lane;
            float y = yv[r][c];
            if (skip_w) {
                const float sw = skip_w[S + lane];
                y = x_res[(size_t)i * HID + w] * sw + y * (1.0f - sw);
            }
            x_new[(size_t)i * HID + w] = y;
            if (x_scaled) x_scaled[(size_t)i * HID + w] = s_next ? y * s_next[S + lane] : y;
        }
    }
}

// ---- block tail on the tensor cores: Gate -> operand packer, [gate | x_in] . [W_self ; W_skip] by jamun_gemm_tf32x3, mix ----
// tail_pack: per node, the activated scalars (120 -> 4 stages of 32, zero padded), the block input scalars (NS stages), and per
// component the gated vectors (1 stage) and the block input vectors (1 stage) in the GEMM's stage-major chunk-swizzled layout.
template <int S_IN, int V_IN>
__global__ void __launch_bounds__(256)
tail_pack_kernel(const float* __restrict__ conv, const float* __restrict__ vadd, const float* __restrict__ x_in, float c_act,
                 float c_gate, int N, int rows_pad, float* __restrict__ a_s, float* __restrict__ a_v, size_t comp_stride,
                 const int* __restrict__ rowptr, const float* __restrict__ rhat, const float* __restrict__ t_edge, float p2_scale,
                 int conv_has_v) {
    constexpr int D_IN = S_IN + 3 * V_IN, S = JAMUN_S, V = JAMUN_V, NS = (S_IN + 31) / 32;
    const int lane = threadIdx.x & 31;
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= N) return;
    const int swz = (((lane >> 2) ^ (i & 7)) << 2) | (lane & 3);
    const float* o = conv + (size_t)i * JAMUN_GATE_IN;
    const float* xi = x_in + (size_t)i * D_IN;
#pragma unroll
    for (int st = 0; st < 4; ++st) {
        const int t = 32 * st + lane;
        float v = 0.f;
        if (t < S) {
            v = o[t];
            v = c_act * (v > 0.f ? v : 0.01f * v);
        }
        a_s[((size_t)st * rows_pad + i) * 32 + swz] = v;
    }
#pragma unroll
    for (int st = 0; st < NS; ++st) {
        const int t = 32 * st + lane;
        a_s[((size_t)(4 + st) * rows_pad + i) * 32 + swz] = t < S_IN ? xi[t] : 0.f;
    }
    const float gate = c_gate * sigmoidf_acc(o[S + lane]);
    // receiver-side sum of the 0e(x)1e->1e path over this node's (contiguous) in-edges: sum_e rhat_e[c] * T_e[w]
    float p2v[3] = {0.f, 0.f, 0.f};
    if (t_edge) {
        const int e0 = rowptr[i], e1 = rowptr[i + 1];
#pragma unroll 4
        for (int e = e0; e < e1; ++e) {
            const float4 rh = *reinterpret_cast<const float4*>(rhat + 4 * (size_t)e);
            const float t = t_edge[(size_t)e * V + lane];
            p2v[0] = fmaf(rh.x, t, p2v[0]);
            p2v[1] = fmaf(rh.y, t, p2v[1]);
            p2v[2] = fmaf(rh.z, t, p2v[2]);
        }
        const float sc = p2_scale / (float)(e1 > e0 ? e1 - e0 : 1);
#pragma unroll
        for (int c = 0; c < 3; ++c) p2v[c] *= sc;
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float v = (conv_has_v ? o[SO + c * V + lane] : 0.f) + p2v[c];
        if (vadd) v += vadd[(size_t)i * (3 * V) + c * V + lane];
        float* av = a_v + c * comp_stride;
        av[(size_t)i * 32 + swz] = v * gate;
        if (V_IN > 0) av[((size_t)rows_pad + i) * 32 + swz] = xi[S_IN + c * V_IN + lane];
    }
}

// tail_mix: noise-conditional skip (x_new = x_res*w + y*(1-w)) and the next block's input scaling
__global__ void __launch_bounds__(256)
tail_mix_kernel(const float* __restrict__ y, const float* __restrict__ x_res, const float* __restrict__ skip_w,
                const float* __restrict__ s_next, int N, float* __restrict__ x_new, float* __restrict__ x_scaled,
                float* __restrict__ xs_op, int rows_pad) {
    constexpr int S = JAMUN_S, V = JAMUN_V, HID = JAMUN_HID;
    const size_t total = (size_t)N * HID;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int w = (int)(idx % HID);
        const int cw = w < S ? w : S + (w - S) % V;  // per-irrep weight index
        float v = y[idx];
        if (skip_w) {
            const float sw = skip_w[cw];
            v = x_res[idx] * sw + v * (1.0f - sw);
        }
        x_new[idx] = v;
        if (x_scaled) {
            const float vs = s_next ? v * s_next[cw] : v;
            x_scaled[idx] = vs;
            if (xs_op && w < S) {  // scalars of the next block's input, already in the GEMM's stage-major operand layout
                const int i = (int)(idx / HID), p = w & 31;
                xs_op[((size_t)(w >> 5) * rows_pad + i) * 32 + ((((p >> 2) ^ (i & 7)) << 2) | (p & 3))] = vs;
            }
        }
    }
}

// ---- head: Linear(hidden -> 152x0e+32x1e) -> Gate -> Linear(-> 1x1e) * gain.  Only the 32 gate scalars and the
// gated vectors reach the output, so the 120 activated scalars are never formed.
__global__ void __launch_bounds__(WARPS * 32)
head_kernel(const float* __restrict__ x, const float* __restrict__ w1_s, const float* __restrict__ w1_v,
            const float* __restrict__ w2, float c_gate, int N, float* __restrict__ g) {
    constexpr int S = JAMUN_S, V = JAMUN_V, HID = JAMUN_HID;
    extern __shared__ __align__(16) float smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* Xs = smem + warp * NPW * HID;
    const int node0 = blockIdx.x * TM + warp * NPW;
    if (node0 >= N) return;
    for (int r = 0; r < NPW; ++r) {
        const int i = node0 + r;
        for (int t = lane; t < HID; t += 32) Xs[r * HID + t] = i < N ? x[(size_t)i * HID + t] : 0.f;
    }
    __syncwarp();
    float gs[NPW], hv[NPW][3];
#pragma unroll
    for (int r = 0; r < NPW; ++r) gs[r] = hv[r][0] = hv[r][1] = hv[r][2] = 0.f;
    for (int u = 0; u < S; ++u) {
        const float b = w1_s[(size_t)u * SO + S + lane];
#pragma unroll
        for (int r = 0; r < NPW; ++r) gs[r] = fmaf(Xs[r * HID + u], b, gs[r]);
    }
    for (int u = 0; u < V; ++u) {
        const float b = w1_v[u * V + lane];
#pragma unroll
        for (int r = 0; r < NPW; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) hv[r][c] = fmaf(Xs[r * HID + S + c * V + u], b, hv[r][c]);
    }
    const float w2l = w2[lane];
#pragma unroll
    for (int r = 0; r < NPW; ++r) {
        const float gate = c_gate * sigmoidf_acc(gs[r]);
        float o[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) o[c] = warp_sum(hv[r][c] * gate * w2l);
        const int i = node0 + r;
        if (i < N && lane < 3) g[3 * (size_t)i + lane] = lane == 0 ? o[0] : lane == 1 ? o[1] : o[2];
    }
}

template <typename K>
int set_smem(K kernel, size_t bytes) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) {
        jb::set_error("cudaFuncSetAttribute(%zu B): %s", bytes, cudaGetErrorString(e));
        return JAMUN_ECUDA;
    }
    return JAMUN_OK;
}

}  // namespace

extern "C" int jamun_conv_fwd(const float* x, int s_in, int v_in, const int* rowptr, const int* col, const float* h,
                              const float* rhat, const float* m0, const float* m1, float alpha0, float alpha1, int N,
                              float* out, jamun_stream_t stream) {
    JB_CHECK_ARG(x && rowptr && col && h && rhat && m0 && m1 && out, "null argument");
    if (N == 0) return JAMUN_OK;
    const int blocks = (N + TM - 1) / TM;
    cudaStream_t s = jb::as_stream(stream);
    if (s_in == JAMUN_S && v_in == JAMUN_V) {
        using C = ConvCfg<JAMUN_S, JAMUN_V>;
        if (int rc = set_smem(conv_simt_kernel<JAMUN_S, JAMUN_V>, C::SMEM)) return rc;
        conv_simt_kernel<JAMUN_S, JAMUN_V><<<blocks, WARPS * 32, C::SMEM, s>>>(x, rowptr, col, h, rhat, m0, m1, alpha0,
                                                                              alpha1, N, out);
    } else if (s_in == JAMUN_S0 && v_in == 0) {
        using C = ConvCfg<JAMUN_S0, 0>;
        if (int rc = set_smem(conv_simt_kernel<JAMUN_S0, 0>, C::SMEM)) return rc;
        conv_simt_kernel<JAMUN_S0, 0><<<blocks, WARPS * 32, C::SMEM, s>>>(x, rowptr, col, h, rhat, m0, m1, alpha0,
                                                                         alpha1, N, out);
    } else {
        jb::set_error("jamun_conv_fwd: unsupported input irreps %dx0e+%dx1e (built for 120x0e+32x1e and 56x0e)", s_in, v_in);
        return JAMUN_EINVAL;
    }
    JB_CHECK_LAUNCH();
    return JAMUN_OK;
}

extern "C" int jamun_block_tail(const float* conv, const float* vadd, const float* x_in, int s_in, int v_in, const float* x_res,
                                const float* wself_s, const float* wself_v, const float* wskip_s, const float* wskip_v,
                                const float* skip_w, const float* s_next, float c_act, float c_gate, int N,
                                float* x_new, float* x_scaled, jamun_stream_t stream) {
    JB_CHECK_ARG(conv && x_in && wself_s && wself_v && wskip_s && x_new, "null argument");
    JB_CHECK_ARG(!skip_w || x_res, "skip_w needs x_res");
    if (N == 0) return JAMUN_OK;
    const int blocks = (N + TM - 1) / TM;
    cudaStream_t s = jb::as_stream(stream);
    if (s_in == JAMUN_S && v_in == JAMUN_V) {
        JB_CHECK_ARG(wskip_v, "wskip_v required for vector inputs");
        constexpr size_t smem = (size_t)WARPS * NPW * (JAMUN_HID + JAMUN_HID) * sizeof(float);
        if (int rc = set_smem(block_tail_kernel<JAMUN_S, JAMUN_V>, smem)) return rc;
        block_tail_kernel<JAMUN_S, JAMUN_V><<<blocks, WARPS * 32, smem, s>>>(conv, vadd, x_in, x_res, wself_s, wself_v, wskip_s,
                                                                            wskip_v, skip_w, s_next, c_act, c_gate, N,
                                                                            x_new, x_scaled);
    } else if (s_in == JAMUN_S0 && v_in == 0) {
        constexpr size_t smem = (size_t)WARPS * NPW * (JAMUN_HID + JAMUN_S0) * sizeof(float);
        if (int rc = set_smem(block_tail_kernel<JAMUN_S0, 0>, smem)) return rc;
        block_tail_kernel<JAMUN_S0, 0><<<blocks, WARPS * 32, smem, s>>>(conv, vadd, x_in, x_res, wself_s, wself_v, wskip_s,
                                                                       wskip_v, skip_w, s_next, c_act, c_gate, N, x_new,
                                                                       x_scaled);
    } else {
        jb::set_error("jamun_block_tail: unsupported input irreps %dx0e+%dx1e", s_in, v_in);
        return JAMUN_EINVAL;
    }
    JB_CHECK_LAUNCH();
    return JAMUN_OK;
}

extern "C" int jamun_tail_pack(const float* conv, const float* vadd, const float* x_in, int s_in, int v_in, float c_act,
                               float c_gate, int N, int rows_pad, float* a_s, float* a_v, long long a_v_comp_stride,
                               const int* rowptr, const float* rhat, const float* t_edge, float p2_scale, int conv_has_v,
                               jamun_stream_t stream) {
    JB_CHECK_ARG(conv && x_in && a_s && a_v && N <= rows_pad, "bad argument");
    JB_CHECK_ARG(!t_edge || (rowptr && rhat), "t_edge needs rowptr and rhat");
    if (N == 0) return JAMUN_OK;
    const int blocks = (int)(((size_t)N * 32 + 255) / 256);
    cudaStream_t s = jb::as_stream(stream);
    if (s_in == JAMUN_S && v_in == JAMUN_V) {
        tail_pack_kernel<JAMUN_S, JAMUN_V><<<blocks, 256, 0, s>>>(conv, vadd, x_in, c_act, c_gate, N, rows_pad, a_s, a_v,
                                                                  (size_t)a_v_comp_stride, rowptr, rhat, t_edge, p2_scale,
                                                                  conv_has_v);
    } else if (s_in == JAMUN_S0 && v_in == 0) {
        tail_pack_kernel<JAMUN_S0, 0><<<blocks, 256, 0, s>>>(conv, vadd, x_in, c_act, c_gate, N, rows_pad, a_s, a_v,
                                                             (size_t)a_v_comp_stride, rowptr, rhat, t_edge, p2_scale, conv_has_v);
    } else {
        jb::set_error("jamun_tail_pack: unsupported input irreps %dx0e+%dx1e", s_in, v_in);
        return JAMUN_EINVAL;
    }
    JB_CHECK_LAUNCH();
    return JAMUN_OK;
}

extern "C" int jamun_tail_mix(const float* y, const float* x_res, const float* skip_w, const float* s_next, int N, float* x_new,
                              float* x_scaled, float* xs_op, int rows_pad, jamun_stream_t stream) {
    JB_CHECK_ARG(y && x_new, "null argument");
    JB_CHECK_ARG(!skip_w || x_res, "skip_w needs x_res");
    if (N == 0) return JAMUN_OK;
    size_t total = (size_t)N * JAMUN_HID;
    int blocks = (int)((total + 255) / 256);
    if (blocks > jb::kNumSMs * 16) blocks = jb::kNumSMs * 16;
    tail_mix_kernel<<<blocks, 256, 0, jb::as_stream(stream)>>>(y, x_res, skip_w, s_next, N, x_new, x_scaled, xs_op, rows_pad);
    JB_CHECK_LAUNCH();
    return JAMUN_OK;
}

extern "C" int jamun_head(const float* x, const float* w1_s, const float* w1_v, const float* w2, float c_gate, int N,
                          float* g, jamun_stream_t stream) {
    JB_CHECK_ARG(x && w1_s && w1_v && w2 && g, "null argument");
    if (N == 0) return JAMUN_OK;
    const int blocks = (N + TM - 1) / TM;
    constexpr size_t smem = (size_t)WARPS * NPW * JAMUN_HID * sizeof(float);
    if (int rc = set_smem(head_kernel, smem)) return rc;
    head_kernel<<<blocks, WARPS * 32, smem, jb::as_stream(stream)>>>(x, w1_s, w1_v, w2, c_gate, N, g);
    JB_CHECK_LAUNCH();
    return JAMUN_OK;
}
