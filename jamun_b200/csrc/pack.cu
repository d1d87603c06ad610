// Weight re-layout on the device: row-major fp32 matrices -> the (hi | lo) stage images jamun_gemm_tf32x3 streams as its
// B operand (UMMA K-major SWIZZLE_128B; format described in jamun_b200/packing.py, which is the host-side statement of the
// same layout used by the tests).  One launch per operand: plan building costs a few dozen launches of this kernel instead
// of hundreds of indexing launches, and the training step can re-pack every weight after each optimiser update.
#include <cuda_fp16.h>
#include "common.cuh"

namespace {

// out[cb][stage][2][n_pad*32]; element (stage, n, kk) of column block cb is W[k = stage*32 + kk][col = cb*n_pad + n] with
//   source row  = row_map ? row_map[k] : k      (negative or >= K_src: zero)
//   source elem = src[(row + (col / n_inner) * outer_rows) * ld + (col % n_inner)]      (col >= N_valid: zero)
// transpose != 0: element (k, col) = src[row_map ? row_map[col] : col][k] for k < N_valid, col < K_src (K_src then counts the
// entries of row_map, or the source rows), i.e. the image of the transposed (and row-gathered) matrix.
// n_inner >= N_valid gives a plain [K, N] matrix; n_inner < N_valid addresses a [outer][K][n_inner] tensor whose leading index
// is spread along the columns (the per-node transform W_y[u, k'*32 + w] = m1[k', u, w]).
// F16: the fp16-split image of jamun_gemm_f16x3 instead -- per stage n_pad rows of 128 bytes [hi k0..31 | lo k0..31] (halves,
// hi = rn16(scale * w), lo = rn16(scale * w - hi)), 16-byte chunks of a row XOR-ed with (n & 7); out then holds
// [col_blocks][n_stages][n_pad * 32] 4-byte words.
template <bool F16>
__global__ void __launch_bounds__(256)
pack_b_kernel(const float* __restrict__ src, int ld, const int* __restrict__ row_map, int K_src, int n_stages, int N_valid,
              int n_inner, int outer_rows, int n_pad, int col_blocks, int transpose, float scale, float* __restrict__ out) {
    const long long per_stage = (long long)n_pad * 32;
    const long long total = (long long)col_blocks * n_stages * per_stage;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        // consecutive threads walk kk fastest inside a source row group: t -> (cb, stage, n, kk)
        const int kk = (int)(t & 31);
        long long q = t >> 5;
        const int n = (int)(q % n_pad);
        q /= n_pad;
        const int stage = (int)(q % n_stages);
        const int cb = (int)(q / n_stages);
        const int k = stage * 32 + kk;
        const int col = cb * n_pad + n;
        float v = 0.f;
        if (transpose) {
            // the image of the transposed matrix: element (k, col) = src[row(col)][k]   (backward GEMMs: dA = G . M^T)
            if (k < N_valid) {
                const int row = row_map ? (col < K_src ? row_map[col] : -1) : col;
                if (row >= 0 && (row_map || row < K_src)) v = src[(size_t)row * ld + k];
            }
        } else {
            const int row = row_map ? row_map[k] : k;
            if (row >= 0 && row < K_src && col < N_valid) {
                const int o = col / n_inner, ci = col - o * n_inner;
                v = src[((size_t)row + (size_t)o * outer_rows) * ld + ci];
            }
        }
        if constexpr (F16) {
            v *= scale;
            const __half hi = __float2half_rn(v), lo = __float2half_rn(v - __half2float(hi));
            __half* img = reinterpret_cast<__half*>(out + ((size_t)cb * n_stages + stage) * per_stage);
            const int base = (n >> 3) * 512 + (n & 7) * 64 + (kk & 7);  // half index of the row + position inside a 16-byte chunk
            img[base + ((((kk >> 3)) ^ (n & 7)) << 3)] = hi;
            img[base + (((4 + (kk >> 3)) ^ (n & 7)) << 3)] = lo;
            continue;
        }
        const float hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
        const int fidx = (n >> 3) * 256 + (n & 7) * 32 + (((kk >> 2) ^ (n & 7)) << 2) + (kk & 3);  // float index in the image
        float* img = out + ((size_t)cb * n_stages + stage) * 2 * per_stage;
        img[fidx] = hi;
        img[per_stage + fidx] = v - hi;
    }
}

}  // namespace

extern "C" int jamun_pack_b(const float* src, int ld, const int* row_map, int K_src, int n_stages, int N_valid, int n_inner,
                            int outer_rows, int n_pad, int col_blocks, int transpose, float* out, jamun_stream_t stream) {
    JB_CHECK_ARG(src && out, "null argument");
    JB_CHECK_ARG(n_stages >= 1 && n_pad >= 16 && n_pad % 8 == 0 && col_blocks >= 1 && n_inner >= 1 && ld >= 1, "bad shape");
    JB_CHECK_ARG(transpose || row_map || K_src <= n_stages * 32, "K_src exceeds the padded K extent");
    const long long total = (long long)col_blocks * n_stages * n_pad * 32;
    long long blocks = (total + 255) / 256;
    if (blocks > jb::kNumSMs * 16) blocks = jb::kNumSMs * 16;
    pack_b_kernel<false><<<(int)blocks, 256, 0, jb::as_stream(stream)>>>(src, ld, row_map, K_src, n_stages, N_valid, n_inner,
                                                                          outer_rows, n_pad, col_blocks, transpose, 1.0f, out);
    JB_CHECK_LAUNCH();
    return JAMUN_OK;
}

// The fp16-split images of jamun_gemm_f16x3: same arguments as jamun_pack_b plus the power-of-two pre-scale of the weights (the
// caller divides the GEMM's alpha by it).  out: col_blocks * n_stages * n_pad * 32 floats (half of jamun_pack_b's).
extern "C" int jamun_pack_b_f16(const float* src, int ld, const int* row_map, int K_src, int n_stages, int N_valid, int n_inner,
                                int outer_rows, int n_pad, int col_blocks, int transpose, float scale, float* out,
                                jamun_stream_t stream) {
    JB_CHECK_ARG(src && out, "null argument");
    JB_CHECK_ARG(n_stages >= 1 && n_pad >= 16 && n_pad % 8 == 0 && col_blocks >= 1 && n_inner >= 1 && ld >= 1, "bad shape");
    JB_CHECK_ARG(transpose || row_map || K_src <= n_stages * 32, "K_src exceeds the padded K extent");
    const long long total = (long long)col_blocks * n_stages * n_pad * 32;
    long long blocks = (total + 255) / 256;
    if (blocks > jb::kNumSMs * 16) blocks = jb::kNumSMs * 16;
    pack_b_kernel<true><<<(int)blocks, 256, 0, jb::as_stream(stream)>>>(src, ld, row_map, K_src, n_stages, N_valid, n_inner,
                                                                         outer_rows, n_pad, col_blocks, transpose, scale, out);
    JB_CHECK_LAUNCH();
    return JAMUN_OK;
}
