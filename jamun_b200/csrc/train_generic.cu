// Building blocks of the training (backward) path that are not conv-shaped: small dense products over the node / edge
// dimension and deterministic column reductions.  All FP32 on the CUDA cores: together they are < 2 % of a training step's
// FLOPs (35 kFLOP per node for the self-interaction / skip Linears against 3.8 MFLOP for one conv contraction).
//
//   jamun_rowmat_mul  Y[n, :b] (+)= X[n, :a] . W          (forward / recompute of the o3.Linear stand-ins; dX = dY . W^T)
//   jamun_rowmat_dw   dW (+)= X^T . dY                    (weight gradients: a reduction over nodes or edges)
//   jamun_colsum      out[c] (+)= sum_n M[n, c]           (bias-like gradients, per-irrep scale gradients)
// Reductions over rows run in a fixed order (row chunks -> partial sums -> ascending final sum): results are bit-reproducible.
// `rows_dev` (optional) is a device-side row count (e.g. rowptr + N for the live edge count) clamped to `rows`.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace {

constexpr int kMaxB = 160;

// ---- Y = X . W (transW: W stored [b, a]) ------------------------------------------------------------------------------------
// One CTA = 32 rows.  thread t: row t >> 3, columns (t & 7) + 8 q.  K is staged through shared memory in slabs of 32.
__global__ void __launch_bounds__(256)
rowmat_mul_kernel(const float* __restrict__ X, int ldx, const float* __restrict__ W, int ldw, int transW, float* __restrict__ Y,
                  int ldy, int rows, const int* __restrict__ rows_dev, int a, int b, int accumulate) {
    __shared__ float Xs[32][33];
    __shared__ float Ws[32][kMaxB + 1];
    if (rows_dev) rows = min(rows, *rows_dev);
    const int row0 = blockIdx.x * 32;
    if (row0 >= rows) return;
    const int r = threadIdx.x >> 3, cg = threadIdx.x & 7;
    float acc[kMaxB / 8];
#pragma unroll
    for (int q = 0; q < kMaxB / 8; ++q) acc[q] = 0.f;
    for (int k0 = 0; k0 < a; k0 += 32) {
        for (int t = threadIdx.x; t < 32 * 32; t += 256) {
            const int rr = t >> 5, kk = t & 31;
            Xs[rr][kk] = (row0 + rr < rows && k0 + kk < a) ? X[(size_t)(row0 + rr) * ldx + k0 + kk] : 0.f;
        }
        if (!transW) {
            for (int t = threadIdx.x; t < 32 * b; t += 256) {
                const int kk = t / b, j = t - kk * b;
                Ws[kk][j] = (k0 + kk < a) ? W[(size_t)(k0 + kk) * ldw + j] : 0.f;
            }
        } else {
            for (int t = threadIdx.x; t < 32 * b; t += 256) {
                const int j = t >> 5, kk = t & 31;
                Ws[kk][j] = (k0 + kk < a) ? W[(size_t)j * ldw + k0 + kk] : 0.f;
            }
        }
        __syncthreads();
#pragma unroll 4
        for (int kk = 0; kk < 32; ++kk) {
            const float x = Xs[r][kk];
#pragma unroll
            for (int q = 0; q < kMaxB / 8; ++q)
                if (cg + 8 * q < b) acc[q] = fmaf(x, Ws[kk][cg + 8 * q], acc[q]);
        }
        __syncthreads();
    }
    if (row0 + r < rows) {
        float* y = Y + (size_t)(row0 + r) * ldy;
#pragma unroll
        for (int q = 0; q < kMaxB / 8; ++q)
            if (cg + 8 * q < b) y[cg + 8 * q] = accumulate ? y[cg + 8 * q] + acc[q] : acc[q];
    }
}

// ---- Y = X . W on the warp-level tensor cores ------------------------------------------------------------------------------------
// mma.sync.m16n8k8 (tf32 operands, fp32 accumulate) with the three-product split (hi = v & 0xFFFFE000, lo = v - hi: fp32
// accuracy).  One CTA = 128 rows (a warp per 16 rows, N tiles of 8 columns, NT = ceil(b/8) accumulators per warp).  W is staged
// once per CTA into shared memory, K-major with a row stride of 8 mod 32 floats (conflict-free B fragments); the A fragments come
// straight from X (a load instruction covers eight rows x 16 bytes, both halves of every 32-byte sector are used by the
// instruction pair).  Used when there are enough rows to pay for staging W (rows >= 2048); a <= 160, b <= 160.
__device__ __forceinline__ void mm_split(float v, uint32_t& hi, uint32_t& lo) {
    hi = __float_as_uint(v) & 0xFFFFE000u;
    lo = __float_as_uint(v - __uint_as_float(hi));
}
__device__ __forceinline__ void mm_mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <int NT>
__global__ void __launch_bounds__(256)
rowmat_mul_mma_kernel(const float* __restrict__ X, int ldx, const float* __restrict__ W, int ldw, int transW, float* __restrict__ Y,
                      int ldy, int rows, const int* __restrict__ rows_dev, int a, int b, int accumulate) {
    constexpr int LDW = 8 * NT + (8 * NT % 32 == 8 ? 0 : (40 - 8 * NT % 32) % 32);  // >= 8 NT, == 8 (mod 32)
    extern __shared__ __align__(16) float mm_sm[];  // [a_pad][LDW], rows a..a_pad-1 and columns b.. zero
    if (rows_dev) rows = min(rows, *rows_dev);
    const int row0 = blockIdx.x * 128;
    if (row0 >= rows) return;
    const int a_pad = (a + 7) & ~7;
    for (int t = threadIdx.x; t < a_pad * LDW; t += 256) {
        const int k = t / LDW, j = t - k * LDW;
        float v = 0.f;
        if (k < a && j < b) v = transW ? W[(size_t)j * ldw + k] : W[(size_t)k * ldw + j];
        mm_sm[t] = v;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, tig = lane & 3;
    const int ra = row0 + 16 * warp + g, rb = ra + 8;
    if (row0 + 16 * warp >= rows) return;
    const bool ona = ra < rows, onb = rb < rows;
    const float* xa = X + (size_t)ra * ldx;
    const float* xb = X + (size_t)rb * ldx;
    float acc[NT][4];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
    for (int k0 = 0; k0 < a_pad; k0 += 8) {
        const int ka = k0 + tig, kb = ka + 4;
        uint32_t ahi[4], alo[4];
        mm_split((ona && ka < a) ? xa[ka] : 0.f, ahi[0], alo[0]);
        mm_split((onb && ka < a) ? xb[ka] : 0.f, ahi[1], alo[1]);
        mm_split((ona && kb < a) ? xa[kb] : 0.f, ahi[2], alo[2]);
        mm_split((onb && kb < a) ? xb[kb] : 0.f, ahi[3], alo[3]);
        const float* Br = mm_sm + ka * LDW + g;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            uint32_t bh0, bl0, bh1, bl1;
            mm_split(Br[8 * nt], bh0, bl0);
            mm_split(Br[4 * LDW + 8 * nt], bh1, bl1);
            mm_mma(acc[nt], alo, bh0, bh1);
            mm_mma(acc[nt], ahi, bl0, bl1);
            mm_mma(acc[nt], ahi, bh0, bh1);
        }
    }
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
        const int c = 8 * nt + 2 * tig;
        if (ona) {
            float* y = Y + (size_t)ra * ldy + c;
            if (c < b) y[0] = accumulate ? y[0] + acc[nt][0] : acc[nt][0];
            if (c + 1 < b) y[1] = accumulate ? y[1] + acc[nt][1] : acc[nt][1];
        }
        if (onb) {
            float* y = Y + (size_t)rb * ldy + c;
            if (c < b) y[0] = accumulate ? y[0] + acc[nt][2] : acc[nt][2];
            if (c + 1 < b) y[1] = accumulate ? y[1] + acc[nt][3] : acc[nt][3];
        }
    }
}

template <int NT>
int launch_rowmat_mul_mma(const float* X, int ldx, const float* W, int ldw, int transW, float* Y, int ldy, int rows, const int* rows_dev,
                          int a, int b, int accumulate, cudaStream_t s) {
    constexpr int LDW = 8 * NT + (8 * NT % 32 == 8 ? 0 : (40 - 8 * NT % 32) % 32);
    static_assert(LDW % 32 == 8 && LDW >= 8 * NT, "row stride of the staged W");
    constexpr size_t smem = (size_t)160 * LDW * sizeof(float);
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(rowmat_mul_mma_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) {
            jb::set_error("jamun_rowmat_mul: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
            return JAMUN_ECUDA;
        }
        attr_set = true;
    }
    const size_t need = (size_t)((a + 7) & ~7) * LDW * sizeof(float);
    rowmat_mul_mma_kernel<NT><<<(rows + 127) / 128, 256, need, s>>>(X, ldx, W, ldw, transW, Y, ldy, rows, rows_dev, a, b, accumulate);
    return JAMUN_OK;
}

// ---- dW = X^T . dY over a chunk of rows -------------------------------------------------------------------------------------
// grid (splits): a CTA reduces its chunk of rows into the whole a x b product (a, b <= 128), 32 rows at a time through shared
// memory; thread (ty, tx) of the 16 x 16 block owns the 8 x 8 register tile dW[ty + 16 i][tx + 16 j], so X and dY are read from
// global memory exactly once and every pair of shared-memory loads feeds eight FMAs.  partial[z][a][b], summed in split order by
// rowmat_dw_reduce_kernel (deterministic).  RI / RJ = register-tile extents actually needed (ceil(a/16), ceil(b/16)).
template <int RI, int RJ>
__global__ void __launch_bounds__(256)
rowmat_dw_kernel(const float* __restrict__ X, int ldx, const float* __restrict__ dY, int ldy, float* __restrict__ partial, int rows,
                 const int* __restrict__ rows_dev, int a, int b, int rows_per_split) {
    constexpr int RB = 32;
    __shared__ float Xs[RB][16 * RI + 1], Ys[RB][16 * RJ + 1];
    if (rows_dev) rows = min(rows, *rows_dev);
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
    const int n_begin = blockIdx.x * rows_per_split, n_end = min(rows, n_begin + rows_per_split);
    float acc[RI][RJ];
#pragma unroll
    for (int i = 0; i < RI; ++i)
#pragma unroll
        for (int j = 0; j < RJ; ++j) acc[i][j] = 0.f;
    for (int n0 = n_begin; n0 < n_end; n0 += RB) {
        for (int t = threadIdx.x; t < RB * 16 * RI; t += 256) {
            const int rr = t / (16 * RI), cc = t - rr * (16 * RI);
            Xs[rr][cc] = (n0 + rr < n_end && cc < a) ? X[(size_t)(n0 + rr) * ldx + cc] : 0.f;
        }
        for (int t = threadIdx.x; t < RB * 16 * RJ; t += 256) {
            const int rr = t / (16 * RJ), cc = t - rr * (16 * RJ);
            Ys[rr][cc] = (n0 + rr < n_end && cc < b) ? dY[(size_t)(n0 + rr) * ldy + cc] : 0.f;
        }
        __syncthreads();
#pragma unroll 4
        for (int rr = 0; rr < RB; ++rr) {
            float xv[RI], yv[RJ];
#pragma unroll
            for (int i = 0; i < RI; ++i) xv[i] = Xs[rr][ty + 16 * i];
#pragma unroll
            for (int j = 0; j < RJ; ++j) yv[j] = Ys[rr][tx + 16 * j];
#pragma unroll
            for (int i = 0; i < RI; ++i)
#pragma unroll
                for (int j = 0; j < RJ; ++j) acc[i][j] = fmaf(xv[i], yv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < RI; ++i)
#pragma unroll
        for (int j = 0; j < RJ; ++j) {
            const int k = ty + 16 * i, c = tx + 16 * j;
            if (k < a && c < b) partial[((size_t)blockIdx.x * a + k) * b + c] = acc[i][j];
        }
}

// The narrow shapes (a <= 32, b <= 64: the radial MLP's first layer over the ~20 edges per atom) on the warp-level tensor cores:
// dW [32 x 64] = X^T [32 x rows] . dY [rows x 64] as mma.sync.m16n8k8 (tf32 operands, fp32 accumulate, three-product split
// hi = v & 0xFFFFE000, lo = v - hi for fp32 accuracy) with the rows as the K dimension.  The fragments are read straight from
// global memory (every element is used once; a load instruction covers four rows x 32 bytes).  A warp accumulates the whole
// product over its k-steps (ascending), the eight warps of a CTA are summed in warp order: deterministic.
__device__ __forceinline__ void dw_split(float v, uint32_t& hi, uint32_t& lo) {
    hi = __float_as_uint(v) & 0xFFFFE000u;
    lo = __float_as_uint(v - __uint_as_float(hi));
}
__device__ __forceinline__ void dw_mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__global__ void __launch_bounds__(256)
rowmat_dw_mma_kernel(const float* __restrict__ X, int ldx, const float* __restrict__ dY, int ldy, float* __restrict__ partial, int rows,
                     const int* __restrict__ rows_dev, int a, int b, int rows_per_split) {
    __shared__ float red[4][32 * 64];  // warps 4..7 deposit first, warps 0..3 add theirs on top: fixed order
    if (rows_dev) rows = min(rows, *rows_dev);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, tig = lane & 3;
    const int n_begin = blockIdx.x * rows_per_split, n_end = min(rows, n_begin + rows_per_split);
    float acc[2][8][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = acc[mt][nt][2] = acc[mt][nt][3] = 0.f;
    for (int r0 = n_begin + 8 * warp; r0 < n_end; r0 += 64) {
        const int ra = r0 + tig, rb = r0 + tig + 4;
        const bool ona = ra < n_end, onb = rb < n_end;
        const float* xa = X + (size_t)ra * ldx;
        const float* xb = X + (size_t)rb * ldx;
        const float* ya = dY + (size_t)ra * ldy;
        const float* yb = dY + (size_t)rb * ldy;
        // A[m = column of X][k = row]: a0 = (g, tig), a1 = (g + 8, tig), a2 = (g, tig + 4), a3 = (g + 8, tig + 4)
        float av[2][4], bv[8][2];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
            const int c0 = 16 * mt + g, c1 = c0 + 8;
            av[mt][0] = (ona && c0 < a) ? xa[c0] : 0.f;
            av[mt][1] = (ona && c1 < a) ? xa[c1] : 0.f;
            av[mt][2] = (onb && c0 < a) ? xb[c0] : 0.f;
            av[mt][3] = (onb && c1 < a) ? xb[c1] : 0.f;
        }
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            const int c = 8 * nt + g;
            bv[nt][0] = (ona && c < b) ? ya[c] : 0.f;
            bv[nt][1] = (onb && c < b) ? yb[c] : 0.f;
        }
        uint32_t ahi[2][4], alo[2][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int q = 0; q < 4; ++q) dw_split(av[mt][q], ahi[mt][q], alo[mt][q]);
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            uint32_t bh0, bl0, bh1, bl1;
            dw_split(bv[nt][0], bh0, bl0);
            dw_split(bv[nt][1], bh1, bl1);
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                dw_mma(acc[mt][nt], alo[mt], bh0, bh1);
                dw_mma(acc[mt][nt], ahi[mt], bl0, bl1);
                dw_mma(acc[mt][nt], ahi[mt], bh0, bh1);
            }
        }
    }
    // D[m][n]: c0 = (g, 2 tig), c1 = (g, 2 tig + 1), c2 = (g + 8, 2 tig), c3 = (g + 8, 2 tig + 1)
#pragma unroll
    for (int round = 0; round < 2; ++round) {
        if ((warp >> 2) == 1 - round) {
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int nt = 0; nt < 8; ++nt) {
                    float* r = red[warp & 3] + (16 * mt + g) * 64 + 8 * nt + 2 * tig;
                    if (round == 0) {
                        r[0] = acc[mt][nt][0], r[1] = acc[mt][nt][1];
                        r[8 * 64] = acc[mt][nt][2], r[8 * 64 + 1] = acc[mt][nt][3];
                    } else {
                        r[0] += acc[mt][nt][0], r[1] += acc[mt][nt][1];
                        r[8 * 64] += acc[mt][nt][2], r[8 * 64 + 1] += acc[mt][nt][3];
                    }
                }
        }
        __syncthreads();
    }
    for (int t = threadIdx.x; t < a * b; t += 256) {
        const int k = t / b, c = t - k * b;
        float sum = 0.f;
#pragma unroll
        for (int w = 0; w < 4; ++w) sum += red[w][k * 64 + c];
        partial[(size_t)blockIdx.x * a * b + t] = sum;
    }
}

__global__ void rowmat_dw_reduce_kernel(const float* __restrict__ partial, int splits, int a, int b, float* __restrict__ dW, int lddw,
                                        int transW, int accumulate) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a * b) return;
    const int k = t / b, j = t - k * b;
    float acc = 0.f;
    for (int z = 0; z < splits; ++z) acc += partial[(size_t)z * a * b + t];
    float* o = dW + (transW ? (size_t)j * lddw + k : (size_t)k * lddw + j);
    *o = accumulate ? *o + acc : acc;
}

// ---- column sums ---------------------------------------------------------------------------------------------------------------
// grid (ceil(cols/32), splits); block (32 columns x 8 row lanes).  flag != null: only rows with flag[n] == flag_value.
__global__ void __launch_bounds__(256)
colsum_kernel(const float* __restrict__ M, int ld, int rows, const int* __restrict__ rows_dev, int cols,
              const unsigned char* __restrict__ flag, int flag_value, float* __restrict__ partial, int rows_per_split) {
    __shared__ float red[8][33];
    if (rows_dev) rows = min(rows, *rows_dev);
    const int c = blockIdx.x * 32 + (threadIdx.x & 31), ry = threadIdx.x >> 5;
    const int n_begin = blockIdx.y * rows_per_split, n_end = min(rows, n_begin + rows_per_split);
    float acc = 0.f;
    if (c < cols) {
        // four rows in flight per thread (independent partial sums, combined in a fixed order)
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        int n = n_begin + ry;
        for (; n + 24 < n_end; n += 32) {
            const float v0 = (!flag || flag[n] == flag_value) ? M[(size_t)n * ld + c] : 0.f;
            const float v1 = (!flag || flag[n + 8] == flag_value) ? M[(size_t)(n + 8) * ld + c] : 0.f;
            const float v2 = (!flag || flag[n + 16] == flag_value) ? M[(size_t)(n + 16) * ld + c] : 0.f;
            const float v3 = (!flag || flag[n + 24] == flag_value) ? M[(size_t)(n + 24) * ld + c] : 0.f;
            a0 += v0, a1 += v1, a2 += v2, a3 += v3;
        }
        for (; n < n_end; n += 8)
            if (!flag || flag[n] == flag_value) a0 += M[(size_t)n * ld + c];
        acc = (a0 + a1) + (a2 + a3);
    }
    red[ry][threadIdx.x & 31] = acc;
    __syncthreads();
    if (ry == 0 && c < cols) {
        float s = 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q) s += red[q][threadIdx.x];
        partial[(size_t)blockIdx.y * cols + c] = s;
    }
}

// out[fold(c)] (+)= sum_z partial[z][c]; fold: SoA rows [S scalars | V x | V y | V z] -> per-irrep [S + V] (fold_v > 0)
__global__ void colsum_reduce_kernel(const float* __restrict__ partial, int splits, int cols, int fold_s, int fold_v,
                                     float* __restrict__ out, int accumulate) {
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    const int n_out = fold_v > 0 ? fold_s + fold_v : cols;
    if (o >= n_out) return;
    float acc = 0.f;
    const int reps = (fold_v > 0 && o >= fold_s) ? (cols - fold_s) / fold_v : 1;
    for (int q = 0; q < reps; ++q) {
        const int c = o + q * fold_v;
        for (int z = 0; z < splits; ++z) acc += partial[(size_t)z * cols + c];
    }
    out[o] = accumulate ? out[o] + acc : acc;
}

inline int pick_splits(int rows, int per_min, int max_splits) {
    int s = (rows + per_min - 1) / per_min;
    if (s < 1) s = 1;
    if (s > max_splits) s = max_splits;
    return s;
}

}  // namespace

extern "C" int jamun_rowmat_mul(const float* X, int ldx, const float* W, int ldw, int transW, float* Y, int ldy, int rows,
                                const int* rows_dev, int a, int b, int accumulate, jamun_stream_t stream) {
    JB_CHECK_ARG(X && W && Y, "null argument");
    JB_CHECK_ARG(a >= 1 && b >= 1 && b <= kMaxB, "b must be in [1, 160]");
    if (rows == 0) return JAMUN_OK;
    cudaStream_t s = jb::as_stream(stream);
    static const bool use_mma = [] {
        const char* e = getenv("JAMUN_B200_ROWMAT");  // "simt" keeps the CUDA-core kernel for every size (A/B reference)
        return !(e && strcmp(e, "simt") == 0);
    }();
    if (use_mma && rows >= 2048 && a <= 160) {
        const int nt = (b + 7) / 8;
        int rc;
        if (nt <= 4) rc = launch_rowmat_mul_mma<4>(X, ldx, W, ldw, transW, Y, ldy, rows, rows_dev, a, b, accumulate, s);
        else if (nt <= 8) rc = launch_rowmat_mul_mma<8>(X, ldx, W, ldw, transW, Y, ldy, rows, rows_dev, a, b, accumulate, s);
        else if (nt <= 15) rc = launch_rowmat_mul_mma<15>(X, ldx, W, ldw, transW, Y, ldy, rows, rows_dev, a, b, accumulate, s);
        else rc = launch_rowmat_mul_mma<20>(X, ldx, W, ldw, transW, Y, ldy, rows, rows_dev, a, b, accumulate, s);
        if (rc != JAMUN_OK) return rc;
    } else {
        rowmat_mul_kernel<<<(rows + 31) / 32, 256, 0, s>>>(X, ldx, W, ldw, transW, Y, ldy, rows, rows_dev, a, b, accumulate);
    }
    JB_CHECK_LAUNCH();
    return JAMUN_OK;
}

// scratch: >= jamun_rowmat_dw_scratch(rows, a, b) floats
inline int dw_splits(int rows) { return pick_splits(rows, 64, 2 * jb::kNumSMs); }  // >= 64 rows per CTA, at most two CTAs per SM

extern "C" long long jamun_rowmat_dw_scratch(int rows, int a, int b) { return (long long)dw_splits(rows) * a * b; }

extern "C" int jamun_rowmat_dw(const float* X, int ldx, const float* dY, int ldy, float* dW, int lddw, int transW, int rows,
                               const int* rows_dev, int a, int b, int accumulate, float* scratch, jamun_stream_t stream) {
    JB_CHECK_ARG(X && dY && dW && scratch, "null argument");
    JB_CHECK_ARG(a >= 1 && b >= 1 && a <= 128 && b <= 128, "shape must be within 128 x 128");
    cudaStream_t s = jb::as_stream(stream);
    const int splits = dw_splits(rows);
    const int per = rows > 0 ? ((rows + splits - 1) / splits + 31) / 32 * 32 : 32;
    if (rows > 0) {
        const int ri = (a + 15) / 16, rj = (b + 15) / 16;
#define JB_DW(RI, RJ) rowmat_dw_kernel<RI, RJ><<<splits, 256, 0, s>>>(X, ldx, dY, ldy, scratch, rows, rows_dev, a, b, per)
        if (ri <= 2 && rj <= 2) JB_DW(2, 2);
        else if (ri <= 2 && rj <= 4 && rows >= 65536)  // many rows (edges): tensor-core kernel
            rowmat_dw_mma_kernel<<<splits, 256, 0, s>>>(X, ldx, dY, ldy, scratch, rows, rows_dev, a, b, per);
        else if (ri <= 2 && rj <= 4) JB_DW(2, 4);
        else if (ri <= 4 && rj <= 4) JB_DW(4, 4);
        else if (ri <= 2) JB_DW(2, 8);
        else if (rj <= 2) JB_DW(8, 2);
        else JB_DW(8, 8);
#undef JB_DW
    }
    rowmat_dw_reduce_kernel<<<(a * b + 255) / 256, 256, 0, s>>>(scratch, rows > 0 ? splits : 0, a, b, dW, lddw, transW, accumulate);
    JB_CHECK_LAUNCH();
    return JAMUN_OK;
}

extern "C" long long jamun_colsum_scratch(int rows, int cols) { return (long long)pick_splits(rows, 1024, 256) * cols; }

extern "C" int jamun_colsum(const float* M, int ld, int rows, const int* rows_dev, int cols, const unsigned char* flag,
                            int flag_value, int fold_s, int fold_v, float* out, int accumulate, float* scratch,
                            jamun_stream_t stream) {
    JB_CHECK_ARG(M && out && scratch && cols >= 1, "bad argument");
    JB_CHECK_ARG(fold_v == 0 || (cols > fold_s && (cols - fold_s) % fold_v == 0), "fold does not divide the columns");
    cudaStream_t s = jb::as_stream(stream);
    const int splits = pick_splits(rows, 1024, 256);
    const int per = rows > 0 ? (rows + splits - 1) / splits : 1;
    if (rows > 0)
        colsum_kernel<<<dim3((cols + 31) / 32, splits), 256, 0, s>>>(M, ld, rows, rows_dev, cols, flag, flag_value, scratch, per);
    const int n_out = fold_v > 0 ? fold_s + fold_v : cols;
    colsum_reduce_kernel<<<(n_out + 127) / 128, 128, 0, s>>>(scratch, rows > 0 ? splits : 0, cols, fold_s, fold_v, out, accumulate);
    JB_CHECK_LAUNCH();
    return JAMUN_OK;
}
