// FullyConnectedTensorProduct with per-row external weights, l <= 1 (e3nn.o3.FullyConnectedTensorProduct(in1, in2, out,
// shared_weights=False, internal_weights=False); call site /root/reference/src/jamun/e3tools/nn/_conv.py:94 `tp(x_src, sh, w)`).
// This is the module-level seam of the reference's plug-in API; the sampling path never materialises per-edge
// weights and does not use this kernel.
//
// One CTA per row z.  For every instruction (i1, i2, io) in e3nn's order, with the path normalisation
// sqrt(dim(ir_out) / sum_{paths -> io} mul1*mul2) and Wigner-3j tensors of unit Frobenius norm:
//   (0,0,0)  t = a b            (0,1,1)  t_k = a b_k / sqrt3      (1,0,1)  t_k = a_k b / sqrt3
//   (1,1,0)  t = a.b / sqrt3    (1,1,1)  t_k = (a x b)_k / sqrt6
//   out[io][w, k] += coeff * sum_{u,v} W[z][off + (u*mul2 + v)*mul_out + w] * t_k(x1[i1][u], x2[i2][v])
#include "common.cuh"

namespace {

struct TpInstr {
    int off1, mul1, l1, off2, mul2, l2, offo, mulo, lo, woff;
    float coeff;
    int pad;
};

constexpr int kMaxDim = 1024;

__global__ void __launch_bounds__(128)
tensor_product_kernel(const float* __restrict__ x1, int d1, const float* __restrict__ x2, int d2, const float* __restrict__ w,
                      long long w_ld, const TpInstr* __restrict__ instr, int n_instr, int d_out, float* __restrict__ out) {
    __shared__ float s1[kMaxDim], s2[64], so[kMaxDim];
    const int z = blockIdx.x;
    for (int t = threadIdx.x; t < d1; t += blockDim.x) s1[t] = x1[(size_t)z * d1 + t];
    for (int t = threadIdx.x; t < d2; t += blockDim.x) s2[t] = x2[(size_t)z * d2 + t];
    for (int t = threadIdx.x; t < d_out; t += blockDim.x) so[t] = 0.f;
    __syncthreads();
    const float* wz = w + (size_t)z * w_ld;
    const float kI3 = 0.57735026918962576451f, kI6 = 0.40824829046386301637f;
    for (int q = 0; q < n_instr; ++q) {
        const TpInstr I = instr[q];
        const int da = 2 * I.l1 + 1, db = 2 * I.l2 + 1, dc = 2 * I.lo + 1;
        for (int wo = threadIdx.x; wo < I.mulo; wo += blockDim.x) {  // the same thread owns (io, wo) in every instruction
            float acc[3] = {0.f, 0.f, 0.f};
            for (int u = 0; u < I.mul1; ++u) {
                const float* a = s1 + I.off1 + u * da;
                for (int v = 0; v < I.mul2; ++v) {
                    const float* b = s2 + I.off2 + v * db;
                    const float wt = wz[I.woff + (u * I.mul2 + v) * I.mulo + wo];
                    if (I.l1 == 0 && I.l2 == 0) {
                        acc[0] = fmaf(wt, a[0] * b[0], acc[0]);
                    } else if (I.l1 == 0 && I.l2 == 1) {
                        for (int k = 0; k < 3; ++k) acc[k] = fmaf(wt, a[0] * b[k] * kI3, acc[k]);
                    } else if (I.l1 == 1 && I.l2 == 0) {
                        for (int k = 0; k < 3; ++k) acc[k] = fmaf(wt, a[k] * b[0] * kI3, acc[k]);
                    } else if (I.lo == 0) {
                        acc[0] = fmaf(wt, (a[0] * b[0] + a[1] * b[1] + a[2] * b[2]) * kI3, acc[0]);
                    } else {
                        acc[0] = fmaf(wt, (a[1] * b[2] - a[2] * b[1]) * kI6, acc[0]);
                        acc[1] = fmaf(wt, (a[2] * b[0] - a[0] * b[2]) * kI6, acc[1]);
                        acc[2] = fmaf(wt, (a[0] * b[1] - a[1] * b[0]) * kI6, acc[2]);
                    }
                }
            }
            for (int k = 0; k < dc; ++k) so[I.offo + wo * dc + k] += I.coeff * acc[k];
        }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < d_out; t += blockDim.x) out[(size_t)z * d_out + t] = so[t];
}

}  // namespace

// instr: [n_instr][12] int32 device table (off1, mul1, l1, off2, mul2, l2, offo, mulo, lo, woff, coeff as float bits, 0)
extern "C" int jamun_tensor_product(const float* x1, int d1, const float* x2, int d2, const float* w, long long w_ld,
                                    const int* instr, int n_instr, int d_out, int Z, float* out, jamun_stream_t stream) {
    JB_CHECK_ARG(x1 && x2 && w && instr && out, "null argument");
    JB_CHECK_ARG(d1 >= 1 && d1 <= kMaxDim && d2 >= 1 && d2 <= 64 && d_out >= 1 && d_out <= kMaxDim, "irreps dimension out of range");
    static_assert(sizeof(TpInstr) == 12 * sizeof(int), "instruction record layout");
    if (Z == 0) return JAMUN_OK;
    tensor_product_kernel<<<Z, 128, 0, jb::as_stream(stream)>>>(x1, d1, x2, d2, w, w_ld, reinterpret_cast<const TpInstr*>(instr),
                                                                n_instr, d_out, out);
    JB_CHECK_LAUNCH();
    return JAMUN_OK;
}
