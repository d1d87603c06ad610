// Walk-jump integrator kernels: fused (xhat, score, clip, BAOAB update, Philox draw, next-eval prologue),
// Gaussian axpy, and the small per-(model, sigma) constant kernels (noise MLP, atom embedding).
#include "common.cuh"

namespace {
using namespace jb;

// One warp per chain.  See include/jamun_b200.h::jamun_walk_step for the update it performs.
__global__ void walk_step_kernel(float* __restrict__ y, float* __restrict__ v, float* __restrict__ ybar,
                                 float* __restrict__ p, const float* __restrict__ g,
                                 const float* __restrict__ score_in, const int* __restrict__ chain_ptr,
                                 int G, jamun_walk_params prm, float sigma2, float half_delta, float u_half_delta,
                                 const float* __restrict__ noise, float* __restrict__ xhat, float* __restrict__ score,
                                 float* __restrict__ traj_y, float* __restrict__ traj_xhat,
                                 float* __restrict__ traj_score, const unsigned long long* __restrict__ dev_state) {
    const int chain = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (chain >= G) return;
    if (dev_state) {  // graph replay: step counter and trajectory frame live in device memory
        prm.step = dev_state[0];
        const size_t off = (size_t)dev_state[1] * 3 * (size_t)chain_ptr[G];
        if (traj_y) traj_y += off;
        if (traj_xhat) traj_xhat += off;
        if (traj_score) traj_score += off;
    }
    const int lo = chain_ptr[chain], hi = chain_ptr[chain + 1];
    const float inv_n = 1.0f / fmaxf(1.0f, (float)(hi - lo));

    // centroid of t = c_skip*ybar + c_out*g
    float sx = 0.f, sy = 0.f, sz = 0.f;
    if (!score_in) {
        for (int i = lo + lane; i < hi; i += 32) {
            sx += prm.c_skip * ybar[3 * i] + prm.c_out * g[3 * i];
            sy += prm.c_skip * ybar[3 * i + 1] + prm.c_out * g[3 * i + 1];
            sz += prm.c_skip * ybar[3 * i + 2] + prm.c_out * g[3 * i + 2];
        }
    }
    const float cen = prm.center ? inv_n : 0.f;
    const float mx = warp_sum(sx) * cen, my = warp_sum(sy) * cen, mz = warp_sum(sz) * cen;

    float nx = 0.f, ny = 0.f, nz = 0.f;  // centroid accumulators of the advanced y
    for (int i = lo + lane; i < hi; i += 32) {
        float yy[3] = {y[3 * i], y[3 * i + 1], y[3 * i + 2]};
        float xh[3] = {0.f, 0.f, 0.f};
        float sc[3], psi[3];
        if (score_in) {
#pragma unroll
            for (int c = 0; c < 3; ++c) sc[c] = score_in[3 * i + c];
        } else {
            xh[0] = prm.c_skip * ybar[3 * i] + prm.c_out * g[3 * i] - mx;
            xh[1] = prm.c_skip * ybar[3 * i + 1] + prm.c_out * g[3 * i + 1] - my;
            xh[2] = prm.c_skip * ybar[3 * i + 2] + prm.c_out * g[3 * i + 2] - mz;
#pragma unroll
            for (int c = 0; c < 3; ++c) sc[c] = (xh[c] - yy[c]) / sigma2;
        }
        if (prm.clip > 0.f) {
            const float nrm = sqrtf(sc[0] * sc[0] + sc[1] * sc[1] + sc[2] * sc[2]);
            const float cl = fminf(nrm, prm.clip);
#pragma unroll
            for (int c = 0; c < 3; ++c) psi[c] = nrm > 0.f ? (sc[c] / nrm) * cl : 0.f;
        } else {
#pragma unroll
            for (int c = 0; c < 3; ++c) psi[c] = sc[c];
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) psi[c] *= prm.beta;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            if (xhat && !score_in) xhat[3 * i + c] = xh[c];
            if (score) score[3 * i + c] = sc[c];
            if (traj_y) traj_y[3 * i + c] = yy[c];
            if (traj_xhat && !score_in) traj_xhat[3 * i + c] = xh[c];
            if (traj_score) traj_score[3 * i + c] = sc[c];
        }
        float vv[3] = {v[3 * i], v[3 * i + 1], v[3 * i + 2]};
        if (!prm.first) {
#pragma unroll
            for (int c = 0; c < 3; ++c) vv[c] = vv[c] + half_delta * psi[c];  // closing B (no u: reference quirk)
        }
        if (!prm.last) {
            float3 R;
            if (noise) R = make_float3(noise[3 * i], noise[3 * i + 1], noise[3 * i + 2]);
            else R = Philox::normal3(prm.seed, prm.step, (uint64_t)i);
            const float Rr[3] = {R.x, R.y, R.z};
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                vv[c] = vv[c] + u_half_delta * psi[c];        // B
                yy[c] = yy[c] + half_delta * vv[c];           // A
                vv[c] = prm.a * vv[c] + prm.z_sqrt_u * Rr[c]; // O
                yy[c] = yy[c] + half_delta * vv[c];           // A
                y[3 * i + c] = yy[c];
            }
            nx += yy[0];
            ny += yy[1];
            nz += yy[2];
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) v[3 * i + c] = vv[c];
    }
    if (prm.last) return;
    const float cx = warp_sum(nx) * cen, cy = warp_sum(ny) * cen, cz = warp_sum(nz) * cen;
    __syncwarp();
    for (int i = lo + lane; i < hi; i += 32) {
        const float bx = y[3 * i] - cx, by = y[3 * i + 1] - cy, bz = y[3 * i + 2] - cz;
        ybar[3 * i] = bx;
        ybar[3 * i + 1] = by;
        ybar[3 * i + 2] = bz;
        p[3 * i] = bx * prm.c_in;
        p[3 * i + 1] = by * prm.c_in;
        p[3 * i + 2] = bz * prm.c_in;
    }
}

__device__ __forceinline__ void clip_scale(const float* sc, float clip, float beta, float* psi) {
    if (clip > 0.f) {
        const float nrm = sqrtf(sc[0] * sc[0] + sc[1] * sc[1] + sc[2] * sc[2]);
        const float cl = fminf(nrm, clip);
#pragma unroll
        for (int c = 0; c < 3; ++c) psi[c] = nrm > 0.f ? (sc[c] / nrm) * cl : 0.f;
    } else {
#pragma unroll
        for (int c = 0; c < 3; ++c) psi[c] = sc[c];
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) psi[c] *= beta;
}

__global__ void aboba_drift_kernel(float* __restrict__ y, const float* __restrict__ v, float half_delta, int n) {
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < 3 * n; t += gridDim.x * blockDim.x) y[t] = y[t] + half_delta * v[t];
}

__global__ void aboba_kick_kernel(float* __restrict__ y, float* __restrict__ v, const float* __restrict__ score,
                                  jamun_walk_params prm, float half_delta, float u_half_delta,
                                  const float* __restrict__ noise, int n) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float sc[3] = {score[3 * i], score[3 * i + 1], score[3 * i + 2]}, psi[3];
        clip_scale(sc, prm.clip, prm.beta, psi);
        float3 R;
        if (noise) R = make_float3(noise[3 * i], noise[3 * i + 1], noise[3 * i + 2]);
        else R = Philox::normal3(prm.seed, prm.step, (uint64_t)i);
        const float Rr[3] = {R.x, R.y, R.z};
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float vv = v[3 * i + c] + u_half_delta * psi[c];
            vv = prm.a * vv + prm.z_sqrt_u * Rr[c];
            vv = vv + half_delta * psi[c];
            v[3 * i + c] = vv;
            y[3 * i + c] = y[3 * i + c] + half_delta * vv;
        }
    }
}

__global__ void gaussian_axpy_kernel(const float* __restrict__ x, float a, float b, const float* __restrict__ noise,
                                     uint64_t seed, uint64_t step, int n_atoms, float* __restrict__ out) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_atoms; i += gridDim.x * blockDim.x) {
        float3 R;
        if (noise) R = make_float3(noise[3 * i], noise[3 * i + 1], noise[3 * i + 2]);
        else R = Philox::normal3(seed, step, (uint64_t)i);
        const float x0 = x ? x[3 * i] : 0.f, x1 = x ? x[3 * i + 1] : 0.f, x2 = x ? x[3 * i + 2] : 0.f;
        out[3 * i] = a * x0 + b * R.x;
        out[3 * i + 1] = a * x1 + b * R.y;
        out[3 * i + 2] = a * x2 + b * R.z;
    }
}

__global__ void noise_mlp_kernel(const float* __restrict__ w1, const float* __restrict__ b1, const float* __restrict__ w2,
                                 const float* __restrict__ b2, float c, int n, int apply_sigmoid, float* __restrict__ out) {
    extern __shared__ float hid[];
    constexpr float kAlpha = 1.6732632423543772848170429916717f, kScale = 1.0507009873554804934193349852946f;
    for (int o = threadIdx.x; o < n; o += blockDim.x) {
        const float t = w1[o] * c + b1[o];
        hid[o] = kScale * (fmaxf(t, 0.f) + fminf(0.f, kAlpha * (expf(t) - 1.0f)));
    }
    __syncthreads();
    for (int o = threadIdx.x; o < n; o += blockDim.x) {
        float acc = 0.f;
        for (int k = 0; k < n; ++k) acc = fmaf(w2[(size_t)o * n + k], hid[k], acc);
        acc += b2[o];
        out[o] = apply_sigmoid ? sigmoidf_acc(acc) : acc;
    }
}

__global__ void atom_embed_kernel(const int* __restrict__ i0, const int* __restrict__ i1, const int* __restrict__ i2,
                                  const int* __restrict__ i3, const float* __restrict__ t0, const float* __restrict__ t1,
                                  const float* __restrict__ t2, const float* __restrict__ t3, int d0, int d1, int d2,
                                  int d3, const float* __restrict__ scale, int N, float* __restrict__ out) {
    const int D = d0 + d1 + d2 + d3;
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < (size_t)N * D; t += (size_t)gridDim.x * blockDim.x) {
        const int n = (int)(t / D);
        int k = (int)(t % D);
        float val;
        if (k < d0) val = t0[(size_t)i0[n] * d0 + k];
        else if ((k -= d0) < d1) val = t1[(size_t)i1[n] * d1 + k];
        else if ((k -= d1) < d2) val = t2[(size_t)i2[n] * d2 + k];
        else { k -= d2; val = t3[(size_t)(i3 ? i3[n] : 0) * d3 + k]; }
        out[t] = scale ? val * scale[t % D] : val;
    }
}

}  // namespace

extern "C" int jamun_walk_step(float* y, float* v, float* ybar, float* p, const float* g, const float* score_in,
                               const int* chain_ptr, int G, const jamun_walk_params* prm, const float* noise,
                               float* xhat, float* score, float* traj_y, float* traj_xhat, float* traj_score,
                               const unsigned long long* dev_state, jamun_stream_t stream) {
    JB_CHECK_ARG(y && v && ybar && p && (g || score_in) && chain_ptr && prm, "null argument");
    if (G == 0) return JAMUN_OK;
    const float sigma2 = prm->sigma2;
    const float half_delta = prm->delta * 0.5f;
    const float u_half_delta = prm->u * half_delta;
    int blocks = (G * 32 + 255) / 256;
    walk_step_kernel<<<blocks, 256, 0, jb::as_stream(stream)>>>(y, v, ybar, p, g, score_in, chain_ptr, G, *prm, sigma2, half_delta,
                                                                u_half_delta, noise, xhat, score, traj_y, traj_xhat,
                                                                traj_score, dev_state);
    JB_CHECK_LAUNCH();
    return JAMUN_OK;
}

__global__ void walk_advance_kernel(unsigned long long* st, int slot_inc) {
    st[0] += 1;
    st[1] += (unsigned long long)slot_inc;
}

extern "C" int jamun_walk_advance(unsigned long long* dev_state, int slot_inc, jamun_stream_t stream) {
    JB_CHECK_ARG(dev_state, "null argument");
    walk_advance_kernel<<<1, 1, 0, jb::as_stream(stream)>>>(dev_state, slot_inc);
    JB_CHECK_LAUNCH();
    return JAMUN_OK;
}

extern "C" int jamun_aboba_drift(float* y, const float* v, float half_delta, int n_atoms, jamun_stream_t stream) {
    JB_CHECK_ARG(y && v, "null argument");
    if (n_atoms == 0) return JAMUN_OK;
    int blocks = (3 * n_atoms + 255) / 256;
    if (blocks > jb::kNumSMs * 8) blocks = jb::kNumSMs * 8;
    aboba_drift_kernel<<<blocks, 256, 0, jb::as_stream(stream)>>>(y, v, half_delta, n_atoms);
    JB_CHECK_LAUNCH();
    return JAMUN_OK;
}

extern "C" int jamun_aboba_kick(float* y, float* v, const float* score, const jamun_walk_params* prm, const float* noise,
                                int n_atoms, jamun_stream_t stream) {
    JB_CHECK_ARG(y && v && score && prm, "null argument");
    if (n_atoms == 0) return JAMUN_OK;
    int blocks = (n_atoms + 255) / 256;
    if (blocks > jb::kNumSMs * 8) blocks = jb::kNumSMs * 8;
    const float hd = prm->delta * 0.5f;
    aboba_kick_kernel<<<blocks, 256, 0, jb::as_stream(stream)>>>(y, v, score, *prm, hd, prm->u * hd, noise, n_atoms);
    JB_CHECK_LAUNCH();
    return JAMUN_OK;
}

extern "C" int jamun_gaussian_axpy(const float* x, float a, float b, const float* noise, unsigned long long seed,
                                   unsigned long long step, int n_atoms, float* out, jamun_stream_t stream) {
    JB_CHECK_ARG(out, "null argument");
    if (n_atoms == 0) return JAMUN_OK;
    int blocks = (n_atoms + 255) / 256;
    if (blocks > jb::kNumSMs * 8) blocks = jb::kNumSMs * 8;
    gaussian_axpy_kernel<<<blocks, 256, 0, jb::as_stream(stream)>>>(x, a, b, noise, seed, step, n_atoms, out);
    JB_CHECK_LAUNCH();
    return JAMUN_OK;
}

extern "C" int jamun_noise_mlp(const float* w1, const float* b1, const float* w2, const float* b2, float c_noise, int n,
                               int apply_sigmoid, float* out, jamun_stream_t stream) {
    JB_CHECK_ARG(w1 && b1 && w2 && b2 && out, "null argument");
    JB_CHECK_ARG(n > 0 && n <= 4096, "n out of range");
    noise_mlp_kernel<<<1, 256, n * sizeof(float), jb::as_stream(stream)>>>(w1, b1, w2, b2, c_noise, n, apply_sigmoid, out);
    JB_CHECK_LAUNCH();
    return JAMUN_OK;
}

extern "C" int jamun_atom_embed(const int* idx0, const int* idx1, const int* idx2, const int* idx3, const float* tab0,
                                const float* tab1, const float* tab2, const float* tab3, int dim0, int dim1, int dim2,
                                int dim3, const float* scale, int N, float* out, jamun_stream_t stream) {
    JB_CHECK_ARG(idx0 && idx1 && idx2 && tab0 && tab1 && tab2 && tab3 && out, "null argument");
    if (N == 0) return JAMUN_OK;
    int blocks = jb::kNumSMs * 4;
    atom_embed_kernel<<<blocks, 256, 0, jb::as_stream(stream)>>>(idx0, idx1, idx2, idx3, tab0, tab1, tab2, tab3, dim0, dim1,
                                                                 dim2, dim3, scale, N, out);
    JB_CHECK_LAUNCH();
    return JAMUN_OK;
}

// ---- generic dense layer for module-level (compatibility) forwards: out[e][o] = act(sum_k w[o][k] in[e][k] + b[o]) ---------
namespace {
__global__ void linear_act_kernel(const float* __restrict__ in, const float* __restrict__ w, const float* __restrict__ b,
                                  int rows, int K, int O, int act, float* __restrict__ out) {
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < (size_t)rows * O; t += (size_t)gridDim.x * blockDim.x) {
        const int e = (int)(t / O), o = (int)(t % O);
        float acc = b ? b[o] : 0.f;
        for (int k = 0; k < K; ++k) acc = fmaf(w[(size_t)o * K + k], in[(size_t)e * K + k], acc);
        out[t] = act == 1 ? jb::siluf_acc(acc) : acc;
    }
}
}  // namespace

extern "C" int jamun_linear_act(const float* in, const float* w, const float* b, int rows, int K, int O, int act, float* out,
                                jamun_stream_t stream) {
    JB_CHECK_ARG(in && w && out && K > 0 && O > 0, "bad argument");
    if (rows == 0) return JAMUN_OK;
    linear_act_kernel<<<jb::kNumSMs * 8, 256, 0, jb::as_stream(stream)>>>(in, w, b, rows, K, O, act, out);
    JB_CHECK_LAUNCH();
    return JAMUN_OK;
}
