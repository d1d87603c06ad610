// Training-side kernels around the convolution: gate / skip-mix / head elementwise backward, radial-MLP and embedding
// backward, noise-conditioning MLP backward, the denoising loss (forward + backward) and the batched Kabsch alignment.
// Reference: /root/reference/src/jamun/model/denoiser.py:87-109,219-319 (noise, align, loss, training_step),
// utils/align.py:9-56 (Kabsch), e3tools/nn/_gate.py:63, _interaction.py:26-30, model/noise_conditioning.py:27-73.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace {
using namespace jb;

constexpr int S = JAMUN_S, V = JAMUN_V, HID = JAMUN_HID, GIN = JAMUN_GATE_IN, SO = JAMUN_S + JAMUN_V;

// ---- Gate (e3nn.nn.Gate with normalize2mom'd LeakyReLU / sigmoid), SoA layout ---------------------------------------------------
// gated[i] = [c_act * lrelu(conv_s) | conv_v[c] * c_gate * sigmoid(conv_gate)]
__global__ void gate_fwd_kernel(const float* __restrict__ conv, float c_act, float c_gate, int N, float* __restrict__ gated) {
    const size_t total = (size_t)N * HID;
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(t / HID), c = (int)(t % HID);
        const float* o = conv + (size_t)i * GIN;
        float r;
        if (c < S) {
            const float v = o[c];
            r = c_act * (v > 0.f ? v : 0.01f * v);
        } else {
            const int w = (c - S) % V;
            r = o[SO + (c - S)] * c_gate * sigmoidf_acc(o[S + w]);
        }
        gated[t] = r;
    }
}

// dconv from dgated: one thread per (node, w < 152)
__global__ void gate_bwd_kernel(const float* __restrict__ conv, const float* __restrict__ dgated, float c_act, float c_gate, int N,
                                float* __restrict__ dconv) {
    const size_t total = (size_t)N * SO;
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(t / SO), c = (int)(t % SO);
        const float* o = conv + (size_t)i * GIN;
        const float* dg = dgated + (size_t)i * HID;
        float* d = dconv + (size_t)i * GIN;
        if (c < S) {
            d[c] = dg[c] * c_act * (o[c] > 0.f ? 1.f : 0.01f);
        } else {
            const int w = c - S;
            const float sg = sigmoidf_acc(o[S + w]);
            float dgate = 0.f;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const float dv = dg[S + k * V + w];
                dgate = fmaf(dv, o[SO + k * V + w], dgate);
                d[SO + k * V + w] = dv * c_gate * sg;
            }
            d[S + w] = dgate * c_gate * sg * (1.f - sg);
        }
    }
}

// ---- noise-conditional skip / scale (arch/e3conv.py:132-133), backward ----------------------------------------------------------
// forward:  x_new = skip_w ? x_res*w + y*(1-w) : y;   x_scaled = x_new * s   (w, s per irrep, broadcast over vector components)
// given dx_new (may be null) and dx_scaled (may be null):
//   dxn = dx_new + dx_scaled*s;  dy = dxn*(1-w);  dx_res = dxn*w;  prod_s = dx_scaled*x_new;  prod_w = dxn*(x_res - y)
// (prod_* are reduced over nodes and folded per irrep by jamun_colsum)
__global__ void mix_bwd_kernel(const float* __restrict__ dx_new, const float* __restrict__ dx_scaled, const float* __restrict__ y,
                               const float* __restrict__ x_res, const float* __restrict__ skip_w, const float* __restrict__ s_next,
                               int N, float* __restrict__ dy, float* __restrict__ dx_res, float* __restrict__ prod_s,
                               float* __restrict__ prod_w) {
    const size_t total = (size_t)N * HID;
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(t % HID);
        const int ir = c < S ? c : S + (c - S) % V;
        const float yv = y[t];
        const float w = skip_w ? skip_w[ir] : 0.f;
        const float xr = skip_w ? x_res[t] : 0.f;
        const float xn = skip_w ? xr * w + yv * (1.f - w) : yv;
        float dxn = dx_new ? dx_new[t] : 0.f;
        if (dx_scaled && s_next) {
            const float ds = dx_scaled[t];
            dxn = fmaf(ds, s_next[ir], dxn);
            prod_s[t] = ds * xn;
        }
        dy[t] = dxn * (1.f - w);
        if (skip_w) {
            dx_res[t] = dxn * w;
            prod_w[t] = dxn * (xr - yv);
        }
    }
}

// ---- output head (e3tools/nn/_mlp.py:37-114): g[c] = sum_w w2[w] * gate[w] * hv[c][w], gate = c_gate*sigmoid(pre) -------------
// given dg [N,3]: dhv[c][w] = dg[c] w2[w] gate[w];  dpre[w] = (sum_c dg[c] w2[w] hv[c][w]) c_gate s(1-s);
//                 prod_w2[i][w] = sum_c dg[c] gate[w] hv[c][w]   (reduced over nodes -> dw2)
__global__ void head_bwd_kernel(const float* __restrict__ pre, const float* __restrict__ hv, const float* __restrict__ w2,
                                const float* __restrict__ dg, float c_gate, int N, float* __restrict__ dpre,
                                float* __restrict__ dhv, float* __restrict__ prod_w2) {
    const size_t total = (size_t)N * V;
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(t / V), w = (int)(t % V);
        const float sg = sigmoidf_acc(pre[t]);
        const float gate = c_gate * sg;
        float dgate = 0.f, pw = 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float d = dg[(size_t)i * 3 + c], h = hv[((size_t)i * 3 + c) * V + w];
            dhv[((size_t)i * 3 + c) * V + w] = d * w2[w] * gate;
            dgate = fmaf(d * w2[w], h, dgate);
            pw = fmaf(d * gate, h, pw);
        }
        dpre[t] = dgate * c_gate * sg * (1.f - sg);
        prod_w2[t] = pw;
    }
}

// ---- radial MLP hidden layer backward: dz = dh * silu'(z), z = rb . w0r + b0eff[bond] ------------------------------------------
__global__ void radial_bwd_kernel(const float* __restrict__ rb, const unsigned char* __restrict__ ebond, const int* __restrict__ rowptr,
                                  int N, const float* __restrict__ w0r, const float* __restrict__ b0eff, const float* __restrict__ dh,
                                  float* __restrict__ dz) {
    __shared__ float ws[JAMUN_NBASIS][JAMUN_EDGE_HID + 1];
    __shared__ float bs[2][JAMUN_EDGE_HID];
    for (int t = threadIdx.x; t < JAMUN_NBASIS * JAMUN_EDGE_HID; t += blockDim.x) ws[t / JAMUN_EDGE_HID][t % JAMUN_EDGE_HID] = w0r[t];
    for (int t = threadIdx.x; t < 2 * JAMUN_EDGE_HID; t += blockDim.x) bs[t / JAMUN_EDGE_HID][t % JAMUN_EDGE_HID] = b0eff[t];
    __syncthreads();
    const int E = rowptr[N];
    const size_t total = (size_t)E * JAMUN_EDGE_HID;
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const size_t e = t / JAMUN_EDGE_HID;
        const int o = (int)(t % JAMUN_EDGE_HID);
        float z = bs[ebond[e] ? 1 : 0][o];
        const float* r = rb + e * JAMUN_NBASIS;
#pragma unroll 8
        for (int k = 0; k < JAMUN_NBASIS; ++k) z = fmaf(r[k], ws[k][o], z);
        const float sg = sigmoidf_acc(z);
        dz[t] = dh[t] * sg * (1.f + z * (1.f - sg));
    }
}

// The same with z on the warp-level tensor cores: mma.sync.m16n8k8 (tf32 operands, fp32 accumulate, three-product split
// hi = v & 0xFFFFE000, lo = v - hi for fp32 accuracy).  A warp owns 16 edges per pass (A fragments straight from rb), the weight
// is staged once per CTA (row stride 72: conflict-free B fragments); the epilogue reads dh and writes dz as float2 pairs.
__device__ __forceinline__ void rbw_split(float v, uint32_t& hi, uint32_t& lo) {
    hi = __float_as_uint(v) & 0xFFFFE000u;
    lo = __float_as_uint(v - __uint_as_float(hi));
}
__device__ __forceinline__ void rbw_mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__global__ void __launch_bounds__(256)
radial_bwd_mma_kernel(const float* __restrict__ rb, const unsigned char* __restrict__ ebond, const int* __restrict__ rowptr, int N,
                      const float* __restrict__ w0r, const float* __restrict__ b0eff, const float* __restrict__ dh,
                      float* __restrict__ dz) {
    constexpr int LDW = JAMUN_EDGE_HID + 8;
    __shared__ float ws[JAMUN_NBASIS * LDW];
    __shared__ float bs[2 * JAMUN_EDGE_HID];
    for (int t = threadIdx.x; t < JAMUN_NBASIS * JAMUN_EDGE_HID; t += 256) ws[(t / JAMUN_EDGE_HID) * LDW + t % JAMUN_EDGE_HID] = w0r[t];
    for (int t = threadIdx.x; t < 2 * JAMUN_EDGE_HID; t += 256) bs[t] = b0eff[t];
    __syncthreads();
    const int E = rowptr[N];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, tig = lane & 3;
    for (int e0 = 16 * (blockIdx.x * 8 + warp); e0 < E; e0 += 16 * 8 * gridDim.x) {
        const int ra = e0 + g, rbi = ra + 8;
        const bool ona = ra < E, onb = rbi < E;
        const float* pa = rb + (size_t)ra * JAMUN_NBASIS + tig;
        const float* pb = rb + (size_t)rbi * JAMUN_NBASIS + tig;
        float av[4][4];
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            av[ks][0] = ona ? pa[8 * ks] : 0.f;
            av[ks][1] = onb ? pb[8 * ks] : 0.f;
            av[ks][2] = ona ? pa[8 * ks + 4] : 0.f;
            av[ks][3] = onb ? pb[8 * ks + 4] : 0.f;
        }
        float acc[8][4];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            uint32_t ahi[4], alo[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) rbw_split(av[ks][q], ahi[q], alo[q]);
            const float* Br = ws + (8 * ks + tig) * LDW + g;
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                uint32_t bh0, bl0, bh1, bl1;
                rbw_split(Br[8 * nt], bh0, bl0);
                rbw_split(Br[4 * LDW + 8 * nt], bh1, bl1);
                rbw_mma(acc[nt], alo, bh0, bh1);
                rbw_mma(acc[nt], ahi, bl0, bl1);
                rbw_mma(acc[nt], ahi, bh0, bh1);
            }
        }
        const float* ba = bs + ((ona && ebond[ra]) ? JAMUN_EDGE_HID : 0) + 2 * tig;
        const float* bb = bs + ((onb && ebond[rbi]) ? JAMUN_EDGE_HID : 0) + 2 * tig;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            if (ona) {
                const size_t o = (size_t)ra * JAMUN_EDGE_HID + 8 * nt + 2 * tig;
                const float2 d = *reinterpret_cast<const float2*>(dh + o);
                const float z0 = acc[nt][0] + ba[8 * nt], z1 = acc[nt][1] + ba[8 * nt + 1];
                const float s0 = sigmoidf_acc(z0), s1 = sigmoidf_acc(z1);
                *reinterpret_cast<float2*>(dz + o) = make_float2(d.x * s0 * (1.f + z0 * (1.f - s0)), d.y * s1 * (1.f + z1 * (1.f - s1)));
            }
            if (onb) {
                const size_t o = (size_t)rbi * JAMUN_EDGE_HID + 8 * nt + 2 * tig;
                const float2 d = *reinterpret_cast<const float2*>(dh + o);
                const float z0 = acc[nt][2] + bb[8 * nt], z1 = acc[nt][3] + bb[8 * nt + 1];
                const float s0 = sigmoidf_acc(z0), s1 = sigmoidf_acc(z1);
                *reinterpret_cast<float2*>(dz + o) = make_float2(d.x * s0 * (1.f + z0 * (1.f - s0)), d.y * s1 * (1.f + z1 * (1.f - s1)));
            }
        }
    }
}

// ---- atom embedding backward (model/atom_embedding.py:58-76 fused with the initial noise scaling) ------------------------------
// x0[i, col_k + c] = tab_k[idx_k[i]][c] * scale[col_k + c].  One CTA per (table, row r): its 8 warps take contiguous eighths of
// the atoms, read 32 indices per step (coalesced), and for every atom of the step whose index is r (ballot, ascending order)
// add that atom's gradient row (lanes = the table's columns, dim <= 32); the eight partial rows are summed in warp order.
// Fixed order everywhere: bit-reproducible.
__global__ void __launch_bounds__(256)
embed_bwd_kernel(const int* __restrict__ idx, const float* __restrict__ scale, const float* __restrict__ dx0, int ld, int col0, int dim,
                 int n_rows, int N, int per_split, float* __restrict__ partial, float* __restrict__ dtab) {
    __shared__ float part[8][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int r = blockIdx.x;
    if (r >= n_rows) return;
    // grid.y splits the atoms (per_split a multiple of 256); within a split the 8 warps take contiguous eighths
    const int s0 = blockIdx.y * per_split, s1 = min(N, s0 + per_split);
    const int per = per_split / 8;
    const int i0 = s0 + warp * per, i1 = min(s1, i0 + per);
    float acc = 0.f;
    for (int base = i0; base < i1; base += 32) {
        const int i = base + lane;
        const int id = i < i1 ? (idx ? idx[i] : 0) : -1;
        unsigned m = __ballot_sync(0xffffffffu, id == r);
        while (m) {
            const int j = __ffs(m) - 1;
            m &= m - 1;
            if (lane < dim) acc += dx0[(size_t)(base + j) * ld + col0 + lane];
        }
    }
    part[warp][lane] = acc;
    __syncthreads();
    if (warp == 0 && lane < dim) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += part[w][lane];
        if (partial) partial[((size_t)blockIdx.y * n_rows + r) * 32 + lane] = t;
        else dtab[(size_t)r * dim + lane] = t * (scale ? scale[col0 + lane] : 1.f);
    }
}

// dtab[r] = scale * sum over the splits (ascending) of partial[split][r]
__global__ void embed_bwd_reduce_kernel(const float* __restrict__ partial, int splits, int n_rows, const float* __restrict__ scale,
                                        int col0, int dim, float* __restrict__ dtab) {
    const int r = blockIdx.x, lane = threadIdx.x;
    if (lane >= dim) return;
    float t = 0.f;
    for (int z = 0; z < splits; ++z) t += partial[((size_t)z * n_rows + r) * 32 + lane];
    dtab[(size_t)r * dim + lane] = t * (scale ? scale[col0 + lane] : 1.f);
}

// prod[i, c] = dx0[i, c] * emb[i, c] (unscaled embedding) -> colsum gives dscale
__global__ void embed_prod_kernel(const int* __restrict__ i0, const int* __restrict__ i1, const int* __restrict__ i2,
                                  const int* __restrict__ i3, const float* __restrict__ t0, const float* __restrict__ t1,
                                  const float* __restrict__ t2, const float* __restrict__ t3, int d0, int d1, int d2, int d3,
                                  const float* __restrict__ dx0, int N, float* __restrict__ prod) {
    const int D = d0 + d1 + d2 + d3;
    const size_t total = (size_t)N * D;
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(t / D);
        int c = (int)(t % D);
        float e;
        if (c < d0) e = t0[(size_t)i0[i] * d0 + c];
        else if ((c -= d0) < d1) e = t1[(size_t)i1[i] * d1 + c];
        else if ((c -= d1) < d2) e = t2[(size_t)i2[i] * d2 + c];
        else e = t3[(size_t)(i3 ? i3[i] : 0) * d3 + (c - d2)];
        prod[t] = dx0[t] * e;
    }
}

// ---- NoiseConditionalScaling MLP backward (noise_conditioning.py:33-38): out = W2 . selu(w1 c + b1) + b2 [, sigmoid] ----------
__global__ void __launch_bounds__(256)
noise_mlp_bwd_kernel(const float* __restrict__ w1, const float* __restrict__ b1, const float* __restrict__ w2, const float* __restrict__ b2,
                     float c_noise, int n, int apply_sigmoid, const float* __restrict__ dout, float* __restrict__ dw1,
                     float* __restrict__ db1, float* __restrict__ dw2, float* __restrict__ db2) {
    __shared__ float a[256], z1[256], dpre[256];
    const float kAlpha = 1.6732632423543772848170429916717f, kScale = 1.0507009873554804934193349852946f;
    const int t = threadIdx.x;
    if (t < n) {
        const float z = fmaf(w1[t], c_noise, b1[t]);
        z1[t] = z;
        a[t] = kScale * (z > 0.f ? z : kAlpha * expm1f(z));
    }
    __syncthreads();
    if (t < n) {
        float d = dout[t];
        if (apply_sigmoid) {
            float o = b2[t];
            for (int k = 0; k < n; ++k) o = fmaf(w2[(size_t)t * n + k], a[k], o);
            const float sg = sigmoidf_acc(o);
            d *= sg * (1.f - sg);
        }
        dpre[t] = d;
        db2[t] = d;
        for (int k = 0; k < n; ++k) dw2[(size_t)t * n + k] = d * a[k];
    }
    __syncthreads();
    if (t < n) {
        float da = 0.f;
        for (int o = 0; o < n; ++o) da = fmaf(w2[(size_t)o * n + t], dpre[o], da);
        const float z = z1[t];
        const float dz = da * kScale * (z > 0.f ? 1.f : kAlpha * expf(z));
        dw1[t] = dz * c_noise;
        db1[t] = dz;
    }
}

// ---- xhat = center(c_skip*ybar + c_out*g) (model/denoiser.py:200,213-215) and its backward; one warp per chain ----------------
__global__ void __launch_bounds__(256)
combine_xhat_kernel(const float* __restrict__ g, const float* __restrict__ ybar, const int* __restrict__ chain_ptr, int G, float c_skip,
                    float c_out, int center, float* __restrict__ xhat) {
    const int lane = threadIdx.x & 31;
    const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (c >= G) return;
    const int a0 = chain_ptr[c], a1 = chain_ptr[c + 1];
    float m[3] = {0.f, 0.f, 0.f};
    if (center) {
        for (int i = a0 + lane; i < a1; i += 32)
            for (int k = 0; k < 3; ++k) m[k] += (ybar ? c_skip * ybar[3 * i + k] : 0.f) + c_out * g[3 * i + k];
        const float inv = 1.f / (float)max(a1 - a0, 1);
        for (int k = 0; k < 3; ++k) m[k] = warp_sum(m[k]) * inv;
    }
    for (int i = a0 + lane; i < a1; i += 32)
        for (int k = 0; k < 3; ++k) xhat[3 * i + k] = (ybar ? c_skip * ybar[3 * i + k] : 0.f) + c_out * g[3 * i + k] - m[k];
}

// ---- coordinate loss (model/denoiser.py:251-287): raw_g = mean_i |xhat_i - x_i|^2; loss_g = raw_g * lw_g * scale;
// rmsd_g = mean_i |xhat_i - x_i| / (sigma sqrt3).  One warp per chain (fixed summation order).
__global__ void __launch_bounds__(256)
loss_fwd_kernel(const float* __restrict__ xhat, const float* __restrict__ x, const int* __restrict__ chain_ptr, int G, float scale,
                float sigma, const float* __restrict__ lw, float* __restrict__ loss, float* __restrict__ raw, float* __restrict__ rmsd) {
    const int lane = threadIdx.x & 31;
    const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (c >= G) return;
    const int a0 = chain_ptr[c], a1 = chain_ptr[c + 1];
    float sq = 0.f, rm = 0.f;
    for (int i = a0 + lane; i < a1; i += 32) {
        const float dx = xhat[3 * i] - x[3 * i], dy = xhat[3 * i + 1] - x[3 * i + 1], dz = xhat[3 * i + 2] - x[3 * i + 2];
        const float d2 = dx * dx + dy * dy + dz * dz;
        sq += d2;
        rm += sqrtf(d2);
    }
    sq = warp_sum(sq), rm = warp_sum(rm);
    if (lane == 0) {
        const float inv = 1.f / (float)max(a1 - a0, 1);
        raw[c] = sq * inv;
        loss[c] = sq * inv * (lw ? lw[c] : 1.f) * scale;
        rmsd[c] = rm * inv / (sigma * 1.7320508075688772f);
    }
}

// dxhat_i = dloss_g * 2 (xhat_i - x_i) lw_g scale / n_g
__global__ void loss_bwd_kernel(const float* __restrict__ xhat, const float* __restrict__ x, const int* __restrict__ chain_of,
                                const int* __restrict__ chain_ptr, int N, float scale, const float* __restrict__ lw,
                                const float* __restrict__ dloss, float* __restrict__ dxhat) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const int c = chain_of[i];
    const int n = chain_ptr[c + 1] - chain_ptr[c];
    const float coef = dloss[c] * 2.f * (lw ? lw[c] : 1.f) * scale / (float)max(n, 1);
    for (int k = 0; k < 3; ++k) dxhat[3 * i + k] = coef * (xhat[3 * i + k] - x[3 * i + k]);
}

// ---- batched Kabsch alignment (utils/align.py:9-56): y -> R y + t minimising |R y + t - x| per chain -----------------------------
// One warp per chain: centroids and covariance H = sum y_c x_c^T by ordered warp reductions; lane 0 runs a one-sided Jacobi SVD
// of the 3x3 H in fp64 (H V = U S), orders the singular values descending as torch.linalg.svd does, and forms
// R = V diag(1, 1, det(V U^T)) U^T.
__device__ void svd3(const double H[3][3], double U[3][3], double Vm[3][3], double sv[3]) {
    double A[3][3], Vv[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) A[i][j] = H[i][j];
    for (int sweep = 0; sweep < 30; ++sweep) {
        double off = 0.0;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                double alpha = 0, beta = 0, gamma = 0;
                for (int i = 0; i < 3; ++i) alpha += A[i][p] * A[i][p], beta += A[i][q] * A[i][q], gamma += A[i][p] * A[i][q];
                off = fmax(off, fabs(gamma) / (sqrt(alpha * beta) + 1e-300));
                if (fabs(gamma) < 1e-300) continue;
                const double zeta = (beta - alpha) / (2.0 * gamma);
                const double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
                for (int i = 0; i < 3; ++i) {
                    const double ap = A[i][p], aq = A[i][q];
                    A[i][p] = c * ap - s * aq, A[i][q] = s * ap + c * aq;
                    const double vp = Vv[i][p], vq = Vv[i][q];
                    Vv[i][p] = c * vp - s * vq, Vv[i][q] = s * vp + c * vq;
                }
            }
        if (off < 1e-15) break;
    }
    int ord[3] = {0, 1, 2};
    double nrm[3];
    for (int j = 0; j < 3; ++j) nrm[j] = sqrt(A[0][j] * A[0][j] + A[1][j] * A[1][j] + A[2][j] * A[2][j]);
    for (int a = 0; a < 2; ++a)
        for (int b = a + 1; b < 3; ++b)
            if (nrm[ord[b]] > nrm[ord[a]]) {
                const int tmp = ord[a];
                ord[a] = ord[b], ord[b] = tmp;
            }
    for (int j = 0; j < 3; ++j) {
        const int o = ord[j];
        sv[j] = nrm[o];
        for (int i = 0; i < 3; ++i) Vm[i][j] = Vv[i][o], U[i][j] = nrm[o] > 1e-300 ? A[i][o] / nrm[o] : 0.0;
    }
    // rank-deficient H (chains of one or two atoms, collinear chains): complete U to an orthonormal basis; the choice inside
    // the null space does not change R y for the chain's own points
    if (sv[0] <= 1e-300) {
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) U[i][j] = i == j ? 1.0 : 0.0;
    } else if (sv[1] <= 1e-14 * sv[0]) {
        int ax = 0;  // coordinate axis least aligned with U0
        for (int i = 1; i < 3; ++i)
            if (fabs(U[i][0]) < fabs(U[ax][0])) ax = i;
        double v[3] = {0, 0, 0};
        v[ax] = 1.0;
        const double d = U[ax][0];
        double nn = 0;
        for (int i = 0; i < 3; ++i) v[i] -= d * U[i][0], nn += v[i] * v[i];
        nn = sqrt(nn);
        for (int i = 0; i < 3; ++i) U[i][1] = v[i] / nn;
    }
    if (sv[2] <= 1e-14 * sv[0] || sv[2] <= 1e-300) {
        U[0][2] = U[1][0] * U[2][1] - U[2][0] * U[1][1];
        U[1][2] = U[2][0] * U[0][1] - U[0][0] * U[2][1];
        U[2][2] = U[0][0] * U[1][1] - U[1][0] * U[0][1];
    }
}

__global__ void __launch_bounds__(256)
kabsch_kernel(const float* __restrict__ y, const float* __restrict__ x, const int* __restrict__ chain_ptr, int G,
              float* __restrict__ out, float* __restrict__ rot) {
    const int lane = threadIdx.x & 31;
    const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (c >= G) return;
    const int a0 = chain_ptr[c], a1 = chain_ptr[c + 1];
    const int n = a1 - a0;
    float sy[3] = {0, 0, 0}, sx[3] = {0, 0, 0};
    for (int i = a0 + lane; i < a1; i += 32)
        for (int k = 0; k < 3; ++k) sy[k] += y[3 * i + k], sx[k] += x[3 * i + k];
    const float inv = 1.f / (float)max(n, 1);
    float ym[3], xm[3];
    for (int k = 0; k < 3; ++k) ym[k] = warp_sum(sy[k]) * inv, xm[k] = warp_sum(sx[k]) * inv;
    float h[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = a0 + lane; i < a1; i += 32)
        for (int p = 0; p < 3; ++p)
            for (int q = 0; q < 3; ++q) h[3 * p + q] = fmaf(y[3 * i + p] - ym[p], x[3 * i + q] - xm[q], h[3 * p + q]);
    for (int k = 0; k < 9; ++k) h[k] = warp_sum(h[k]);
    float R[9];
    if (lane == 0) {
        double H[3][3], U[3][3], Vm[3][3], sv[3];
        for (int p = 0; p < 3; ++p)
            for (int q = 0; q < 3; ++q) H[p][q] = (double)h[3 * p + q];
        svd3(H, U, Vm, sv);
        // H = U S V^T (y-side U, x-side V);  R = V diag(1,1,d) U^T,  d = det(V U^T)
        double M[3][3];
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) M[i][j] = Vm[i][0] * U[j][0] + Vm[i][1] * U[j][1] + Vm[i][2] * U[j][2];
        const double det = M[0][0] * (M[1][1] * M[2][2] - M[1][2] * M[2][1]) - M[0][1] * (M[1][0] * M[2][2] - M[1][2] * M[2][0]) +
                           M[0][2] * (M[1][0] * M[2][1] - M[1][1] * M[2][0]);
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) R[3 * i + j] = (float)(Vm[i][0] * U[j][0] + Vm[i][1] * U[j][1] + det * Vm[i][2] * U[j][2]);
    }
    for (int k = 0; k < 9; ++k) R[k] = __shfl_sync(0xffffffffu, R[k], 0);
    float t[3];
    for (int p = 0; p < 3; ++p) t[p] = xm[p] - (R[3 * p] * ym[0] + R[3 * p + 1] * ym[1] + R[3 * p + 2] * ym[2]);
    for (int i = a0 + lane; i < a1; i += 32) {
        const float y0 = y[3 * i], y1 = y[3 * i + 1], y2 = y[3 * i + 2];
        for (int p = 0; p < 3; ++p) out[3 * i + p] = R[3 * p] * y0 + R[3 * p + 1] * y1 + R[3 * p + 2] * y2 + t[p];
    }
    if (rot && lane < 9) rot[9 * c + lane] = R[lane];
}

// ---- average squared pair distance below a cutoff, per chain (utils/average_squared_distance.py:154-177) -------------------------
// sums[c] = (sum_{i>j, |x_i-x_j| < cutoff} |x_i-x_j|^2, number of such pairs); one CTA per chain, fixed-order reduction.
__global__ void __launch_bounds__(256)
avg_sq_dist_kernel(const float* __restrict__ pos, const int* __restrict__ chain_ptr, float cutoff, double* __restrict__ sums) {
    __shared__ double s_sum[8], s_cnt[8];
    const int c = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int a0 = chain_ptr[c], a1 = chain_ptr[c + 1];
    double sum = 0.0, cnt = 0.0;
    for (int i = a0 + threadIdx.x; i < a1; i += 256) {
        const float xi = pos[3 * (size_t)i], yi = pos[3 * (size_t)i + 1], zi = pos[3 * (size_t)i + 2];
        for (int j = a0; j < i; ++j) {
            const float dx = xi - pos[3 * (size_t)j], dy = yi - pos[3 * (size_t)j + 1], dz = zi - pos[3 * (size_t)j + 2];
            const float d2 = dx * dx + dy * dy + dz * dz;
            if (cutoff <= 0.f || sqrtf(d2) < cutoff) {
                sum += (double)d2;
                cnt += 1.0;
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    }
    if (lane == 0) s_sum[warp] = sum, s_cnt[warp] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, b = 0.0;
        for (int w = 0; w < 8; ++w) a += s_sum[w], b += s_cnt[w];
        sums[2 * c] = a;
        sums[2 * c + 1] = b;
    }
}

// ema = decay * ema + (1 - decay) * p over a flat parameter buffer (callbacks/_ema.py: EMAOptimizer.update)
__global__ void ema_update_kernel(float* __restrict__ ema, const float* __restrict__ p, float decay, size_t n) {
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < n; t += (size_t)gridDim.x * blockDim.x)
        ema[t] = fmaf(decay, ema[t] - p[t], p[t]);
}

// out[:, col0 : col0 + n] += add[:, :n]
__global__ void add_cols_kernel(float* __restrict__ out, int ld, int col0, const float* __restrict__ add, int add_ld, int n, int N) {
    const size_t total = (size_t)N * n;
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(t / n), c = (int)(t % n);
        out[(size_t)i * ld + col0 + c] += add[(size_t)i * add_ld + c];
    }
}

inline int ew_blocks(size_t total) {
    size_t b = (total + 255) / 256;
    if (b > (size_t)jb::kNumSMs * 16) b = (size_t)jb::kNumSMs * 16;
    return b < 1 ? 1 : (int)b;
}

}  // namespace

extern "C" int jamun_gate_fwd(const float* conv, float c_act, float c_gate, int N, float* gated, jamun_stream_t stream) {
    JB_CHECK_ARG(conv && gated, "null argument");
    if (N == 0) return JAMUN_OK;
    gate_fwd_kernel<<<ew_blocks((size_t)N * HID), 256, 0, jb::as_stream(stream)>>>(conv, c_act, c_gate, N, gated);
    JB_CHECK_LAUNCH();
    return JAMUN_OK;
}

extern "C" int jamun_gate_bwd(const float* conv, const float* dgated, float c_act, float c_gate, int N, float* dconv,
                              jamun_stream_t stream) {
    JB_CHECK_ARG(conv && dgated && dconv, "null argument");
    if (N == 0) return JAMUN_OK;
    gate_bwd_kernel<<<ew_blocks((size_t)N * SO), 256, 0, jb::as_stream(stream)>>>(conv, dgated, c_act, c_gate, N, dconv);
    JB_CHECK_LAUNCH();
    return JAMUN_OK;
}

extern "C" int jamun_mix_bwd(const float* dx_new, const float* dx_scaled, const float* y, const float* x_res, const float* skip_w,
                             const float* s_next, int N, float* dy, float* dx_res, float* prod_s, float* prod_w,
                             jamun_stream_t stream) {
    JB_CHECK_ARG(y && dy && (dx_new || dx_scaled), "null argument");
    JB_CHECK_ARG(!skip_w || (x_res && dx_res && prod_w), "skip_w needs x_res, dx_res, prod_w");
    JB_CHECK_ARG(!(dx_scaled && s_next) || prod_s, "dx_scaled needs prod_s");
    if (N == 0) return JAMUN_OK;
    mix_bwd_kernel<<<ew_blocks((size_t)N * HID), 256, 0, jb::as_stream(stream)>>>(dx_new, dx_scaled, y, x_res, skip_w, s_next, N, dy,
                                                                                  dx_res, prod_s, prod_w);
    JB_CHECK_LAUNCH();
    return JAMUN_OK;
}

extern "C" int jamun_head_bwd(const float* pre, const float* hv, const float* w2, const float* dg, float c_gate, int N, float* dpre,
                              float* dhv, float* prod_w2, jamun_stream_t stream) {
    JB_CHECK_ARG(pre && hv && w2 && dg && dpre && dhv && prod_w2, "null argument");
    if (N == 0) return JAMUN_OK;
    head_bwd_kernel<<<ew_blocks((size_t)N * V), 256, 0, jb::as_stream(stream)>>>(pre, hv, w2, dg, c_gate, N, dpre, dhv, prod_w2);
    JB_CHECK_LAUNCH();
    return JAMUN_OK;
}

extern "C" int jamun_radial_bwd(const float* rb, const unsigned char* ebond, const int* rowptr, int N, int cap, const float* w0r,
                                const float* b0eff, const float* dh, float* dz, jamun_stream_t stream) {
    JB_CHECK_ARG(rb && ebond && rowptr && w0r && b0eff && dh && dz, "null argument");
    if (N == 0 || cap == 0) return JAMUN_OK;
    static const bool use_mma = [] {
        const char* e = getenv("JAMUN_B200_RADIAL_BWD");  // "simt" selects the CUDA-core kernel (A/B reference)
        return !(e && strcmp(e, "simt") == 0);
    }();
    if (use_mma) {
        int blocks = (cap + 127) / 128;
        if (blocks > jb::kNumSMs * 8) blocks = jb::kNumSMs * 8;
        radial_bwd_mma_kernel<<<blocks, 256, 0, jb::as_stream(stream)>>>(rb, ebond, rowptr, N, w0r, b0eff, dh, dz);
    } else {
        radial_bwd_kernel<<<ew_blocks((size_t)cap * JAMUN_EDGE_HID), 256, 0, jb::as_stream(stream)>>>(rb, ebond, rowptr, N, w0r, b0eff, dh,
                                                                                                      dz);
    }
    JB_CHECK_LAUNCH();
    return JAMUN_OK;
}

extern "C" int jamun_embed_bwd(const int* idx0, const int* idx1, const int* idx2, const int* idx3, const float* tab0,
                               const float* tab1, const float* tab2, const float* tab3, int dim0, int dim1, int dim2, int dim3,
                               int rows0, int rows1, int rows2, int rows3, const float* scale, const float* dx0, int N,
                               float* dtab0, float* dtab1, float* dtab2, float* dtab3, float* prod, jamun_stream_t stream) {
    JB_CHECK_ARG(idx0 && idx1 && idx2 && tab0 && tab1 && tab2 && tab3 && dx0 && dtab0 && dtab1 && dtab2 && dtab3 && prod,
                 "null argument");
    JB_CHECK_ARG(dim0 <= 32 && dim1 <= 32 && dim2 <= 32 && dim3 <= 32, "embedding tables wider than 32 columns");
    cudaStream_t s = jb::as_stream(stream);
    const int D = dim0 + dim1 + dim2 + dim3;
    const int* idx[4] = {idx0, idx1, idx2, idx3};
    float* dt[4] = {dtab0, dtab1, dtab2, dtab3};
    const int dims[4] = {dim0, dim1, dim2, dim3}, rows[4] = {rows0, rows1, rows2, rows3};
    int col0 = 0;
    // the atoms are split over grid.y when there are enough of them; the per-split partial rows live in `prod` (free until
    // embed_prod_kernel below overwrites it; the stream orders the two)
    int splits = (N + 1023) / 1024;
    if (splits > 64) splits = 64;
    int max_rows = 1;
    for (int k = 0; k < 4; ++k) max_rows = rows[k] > max_rows ? rows[k] : max_rows;
    if (splits > 1 && (size_t)splits * max_rows * 32 > (size_t)N * D) splits = 1;
    const int per_split = splits > 1 ? ((N + splits - 1) / splits + 255) / 256 * 256 : ((N + 255) / 256 * 256 > 0 ? (N + 255) / 256 * 256 : 256);
    for (int k = 0; k < 4; ++k) {
        if (rows[k] > 0) {
            embed_bwd_kernel<<<::dim3(rows[k], splits), 256, 0, s>>>(idx[k], scale, dx0, D, col0, dims[k], rows[k], N, per_split,
                                                                  splits > 1 ? prod : nullptr, dt[k]);
            if (splits > 1) embed_bwd_reduce_kernel<<<rows[k], 32, 0, s>>>(prod, splits, rows[k], scale, col0, dims[k], dt[k]);
        }
        col0 += dims[k];
    }
    if (N > 0)
        embed_prod_kernel<<<ew_blocks((size_t)N * D), 256, 0, s>>>(idx0, idx1, idx2, idx3, tab0, tab1, tab2, tab3, dim0, dim1, dim2,
                                                                    dim3, dx0, N, prod);
    JB_CHECK_LAUNCH();
    return JAMUN_OK;
}

extern "C" int jamun_noise_mlp_bwd(const float* w1, const float* b1, const float* w2, const float* b2, float c_noise, int n,
                                   int apply_sigmoid, const float* dout, float* dw1, float* db1, float* dw2, float* db2,
                                   jamun_stream_t stream) {
    JB_CHECK_ARG(w1 && b1 && w2 && b2 && dout && dw1 && db1 && dw2 && db2, "null argument");
    JB_CHECK_ARG(n >= 1 && n <= 256, "n must be in [1, 256]");
    noise_mlp_bwd_kernel<<<1, 256, 0, jb::as_stream(stream)>>>(w1, b1, w2, b2, c_noise, n, apply_sigmoid, dout, dw1, db1, dw2, db2);
    JB_CHECK_LAUNCH();
    return JAMUN_OK;
}

// out = center(c_skip*ybar + c_mix*g) per chain (ybar may be NULL).  With ybar = NULL, c_mix = c_out it is also the backward
// of itself with respect to g: dg = c_out * (dxhat - mean_chain dxhat).
extern "C" int jamun_combine_xhat(const float* g, const float* ybar, const int* chain_ptr, int G, float c_skip, float c_mix,
                                  int center, float* xhat, jamun_stream_t stream) {
    JB_CHECK_ARG(g && chain_ptr && xhat, "null argument");
    if (G == 0) return JAMUN_OK;
    combine_xhat_kernel<<<(G * 32 + 255) / 256, 256, 0, jb::as_stream(stream)>>>(g, ybar, chain_ptr, G, c_skip, c_mix, center, xhat);
    JB_CHECK_LAUNCH();
    return JAMUN_OK;
}

extern "C" int jamun_loss_fwd(const float* xhat, const float* x, const int* chain_ptr, int G, float scale, float sigma,
                              const float* loss_weight, float* loss, float* raw, float* rmsd, jamun_stream_t stream) {
    JB_CHECK_ARG(xhat && x && chain_ptr && loss && raw && rmsd, "null argument");
    if (G == 0) return JAMUN_OK;
    loss_fwd_kernel<<<(G * 32 + 255) / 256, 256, 0, jb::as_stream(stream)>>>(xhat, x, chain_ptr, G, scale, sigma, loss_weight, loss, raw,
                                                                             rmsd);
    JB_CHECK_LAUNCH();
    return JAMUN_OK;
}

extern "C" int jamun_loss_bwd(const float* xhat, const float* x, const int* chain_of, const int* chain_ptr, int N, float scale,
                              const float* loss_weight, const float* dloss, float* dxhat, jamun_stream_t stream) {
    JB_CHECK_ARG(xhat && x && chain_of && chain_ptr && dloss && dxhat, "null argument");
    if (N == 0) return JAMUN_OK;
    loss_bwd_kernel<<<(N + 255) / 256, 256, 0, jb::as_stream(stream)>>>(xhat, x, chain_of, chain_ptr, N, scale, loss_weight, dloss, dxhat);
    JB_CHECK_LAUNCH();
    return JAMUN_OK;
}

extern "C" int jamun_kabsch_align(const float* y, const float* x, const int* chain_ptr, int G, float* out, float* rot,
                                  jamun_stream_t stream) {
    JB_CHECK_ARG(y && x && chain_ptr && out, "null argument");
    if (G == 0) return JAMUN_OK;
    kabsch_kernel<<<(G * 32 + 255) / 256, 256, 0, jb::as_stream(stream)>>>(y, x, chain_ptr, G, out, rot);
    JB_CHECK_LAUNCH();
    return JAMUN_OK;
}

extern "C" int jamun_avg_sq_dist(const float* pos, const int* chain_ptr, int G, float cutoff, double* sums, jamun_stream_t stream) {
    JB_CHECK_ARG(pos && chain_ptr && sums, "null argument");
    if (G == 0) return JAMUN_OK;
    avg_sq_dist_kernel<<<G, 256, 0, jb::as_stream(stream)>>>(pos, chain_ptr, cutoff, sums);
    JB_CHECK_LAUNCH();
    return JAMUN_OK;
}

extern "C" int jamun_ema_update(float* ema, const float* p, float decay, long long n, jamun_stream_t stream) {
    JB_CHECK_ARG(ema && p && n >= 0, "bad argument");
    if (n == 0) return JAMUN_OK;
    ema_update_kernel<<<ew_blocks((size_t)n), 256, 0, jb::as_stream(stream)>>>(ema, p, decay, (size_t)n);
    JB_CHECK_LAUNCH();
    return JAMUN_OK;
}

extern "C" int jamun_add_cols(float* out, int ld, int col0, const float* add, int add_ld, int n, int N, jamun_stream_t stream) {
    JB_CHECK_ARG(out && add, "null argument");
    if (N == 0 || n == 0) return JAMUN_OK;
    add_cols_kernel<<<ew_blocks((size_t)N * n), 256, 0, jb::as_stream(stream)>>>(out, ld, col0, add, add_ld, n, N);
    JB_CHECK_LAUNCH();
    return JAMUN_OK;
}
