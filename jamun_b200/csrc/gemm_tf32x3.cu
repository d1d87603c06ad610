// Node-tile GEMM on the 5th-gen tensor cores: out[rows, N] = scale_row * alpha * A[rows, K] . B[K, N] with fp32 accuracy
// from three TF32 products (a_hi*b_hi + a_hi*b_lo + a_lo*b_hi, fp32 accumulation in TMEM).
//
// One CTA owns 128 rows (nodes).  Roles (192 threads):
//   warps 0-3  A producers: thread r streams row r of the fp32 A operand from HBM (stage-major layout, 128 B per stage),
//              splits each value into tf32 hi / exact remainder lo in registers and writes both straight into TMEM with
//              tcgen05.st -- the A operand never touches shared memory (TS-mode MMA).  Afterwards they run the epilogue.
//   warp 4     B producer: one thread issues a 1-D bulk async copy per stage of the pre-swizzled (hi|lo) weight image
//              (packed once on the host in the UMMA K-major SWIZZLE_128B layout) into a 4-deep shared-memory ring.
//   warp 5     MMA issuer: one thread issues 4 k-steps x 3 tcgen05.mma (kind::tf32, M=128, N=n_pad) per stage and
//              commits to the stage's "empty" mbarrier.
// Up to four K-segments, each with its own A, B, N and TMEM accumulator columns, run back to back in one launch
// (conv: one 0e segment with N=160 and three 1e segments with N=32).
#include "common.cuh"
#include "umma.cuh"

namespace {
using namespace jb;

constexpr int kSlots = 4;                  // pipeline depth (A slots in TMEM, B slots in shared memory)
constexpr int kBK = 32;                    // K per stage (one 128-byte swizzle row of tf32)
constexpr int kMaxN = 160;
constexpr int kBSlotBytes = 2 * kMaxN * 128;  // hi + lo images
constexpr int kTmemCols = 512;
constexpr int kACol0 = 256;                // A slots live in TMEM columns [256, 512): 64 columns (hi 32 | lo 32) each
constexpr int kThreads = 192;

struct Seg {
    const float* a;       // [n_stages][rows_pad][32]
    const float* b;       // [n_stages][2][n_pad*32] swizzled images
    float* out;           // output base (row-major, ld = out_ld)
    int n_stages, n_pad, n_valid, d_col, out_col;
    float alpha;
};
struct Params {
    Seg seg[4];
    int nseg, rows, rows_pad, out_ld;
    const float* row_scale;  // [rows] or null
};

struct __align__(1024) Smem {
    uint8_t b[kSlots][kBSlotBytes];
    uint64_t full_a[kSlots], full_b[kSlots], empty[kSlots], d_full;
    uint32_t tmem_base;
};

__device__ __forceinline__ void load_stage(const float* __restrict__ p, float4 (&v)[8]) {
#pragma unroll
    for (int q = 0; q < 8; ++q) v[q] = __ldg(reinterpret_cast<const float4*>(p) + q);
}

__global__ void __launch_bounds__(kThreads, 1) gemm_tf32x3_kernel(const Params P) {
    extern __shared__ uint8_t smem_raw[];
    Smem& S = *reinterpret_cast<Smem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile_row0 = blockIdx.x * 128;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kSlots; ++s) {
            umma::mbar_init(&S.full_a[s], 128);
            umma::mbar_init(&S.full_b[s], 1);
            umma::mbar_init(&S.empty[s], 1);
        }
        umma::mbar_init(&S.d_full, 1);
        umma::fence_barrier_init();
    }
    if (warp == 0) umma::tmem_alloc<kTmemCols>(&S.tmem_base);
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = S.tmem_base;

    int total_stages = 0;
    for (int s = 0; s < P.nseg; ++s) total_stages += P.seg[s].n_stages;

    if (warp < 4) {
        // ------------------------------------------------ A producers
        const int r = threadIdx.x;  // row within the tile == TMEM lane
        const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
        int g = 0;
        float4 cur[8], nxt[8];
        // flattened (segment, stage) iteration with a one-stage register prefetch
        int si = 0, st = 0;
        auto stage_ptr = [&](int seg_i, int stage) {
            return P.seg[seg_i].a + ((size_t)stage * P.rows_pad + tile_row0 + r) * kBK;
        };
        if (total_stages > 0) load_stage(stage_ptr(0, 0), cur);
        while (g < total_stages) {
            int nsi = si, nst = st + 1;
            if (nst == P.seg[si].n_stages) { nsi = si + 1; nst = 0; }
            if (g + 1 < total_stages) load_stage(stage_ptr(nsi, nst), nxt);
            const int slot = g % kSlots;
            const uint32_t par = (g / kSlots) & 1;
            umma::mbar_wait(&S.empty[slot], par ^ 1);
            umma::fence_after_sync();
            uint32_t hi[32], lo[32];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const float v[4] = {cur[q].x, cur[q].y, cur[q].z, cur[q].w};
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const uint32_t h = __float_as_uint(v[t]) & 0xFFFFE000u;
                    hi[4 * q + t] = h;
                    lo[4 * q + t] = __float_as_uint(v[t] - __uint_as_float(h));
                }
            }
            const uint32_t a_addr = tmem + lane_base + (uint32_t)(kACol0 + slot * 64);
            umma::tmem_st32(a_addr, hi);
            umma::tmem_st32(a_addr + 32, lo);
            umma::wait_st();
            umma::fence_before_sync();
            umma::mbar_arrive(&S.full_a[slot]);
#pragma unroll
            for (int q = 0; q < 8; ++q) cur[q] = nxt[q];
            si = nsi;
            st = nst;
            ++g;
        }
        // ------------------------------------------------ epilogue
        umma::mbar_wait(&S.d_full, 0);
        umma::fence_after_sync();
        const int row = tile_row0 + r;
        const float rs = (P.row_scale && row < P.rows) ? P.row_scale[row] : 1.0f;
        for (int s = 0; s < P.nseg; ++s) {
            const Seg& sg = P.seg[s];
            for (int c0 = 0; c0 < sg.n_pad; c0 += 32) {
                uint32_t v[32];
                umma::tmem_ld32(tmem + lane_base + (uint32_t)(sg.d_col + c0), v);
                umma::wait_ld();
                if (row < P.rows) {
                    float* o = sg.out + (size_t)row * P.out_ld + sg.out_col + c0;
#pragma unroll
                    for (int c = 0; c < 32; ++c)
                        if (c0 + c < sg.n_valid) o[c] = __uint_as_float(v[c]) * sg.alpha * rs;
                }
            }
        }
        umma::fence_before_sync();
    } else if (warp == 4) {
        // ------------------------------------------------ B producer
        if (lane == 0) {
            int g = 0;
            for (int s = 0; s < P.nseg; ++s) {
                const Seg& sg = P.seg[s];
                const uint32_t bytes = 2u * sg.n_pad * 128u;
                for (int st = 0; st < sg.n_stages; ++st, ++g) {
                    const int slot = g % kSlots;
                    const uint32_t par = (g / kSlots) & 1;
                    umma::mbar_wait(&S.empty[slot], par ^ 1);
                    umma::mbar_arrive_expect_tx(&S.full_b[slot], bytes);
                    umma::bulk_g2s(S.b[slot], sg.b + (size_t)st * (bytes / 4), bytes, &S.full_b[slot]);
                }
            }
        }
    } else {
        // ------------------------------------------------ MMA issuer
        if (lane == 0) {
            int g = 0;
            for (int s = 0; s < P.nseg; ++s) {
                const Seg& sg = P.seg[s];
                const uint32_t idesc = umma::make_idesc_tf32(128, sg.n_pad);
                const uint32_t d_addr = tmem + (uint32_t)sg.d_col;
                for (int st = 0; st < sg.n_stages; ++st, ++g) {
                    const int slot = g % kSlots;
                    const uint32_t par = (g / kSlots) & 1;
                    umma::mbar_wait(&S.full_a[slot], par);
                    umma::mbar_wait(&S.full_b[slot], par);
                    umma::fence_after_sync();
                    const uint32_t a_hi = tmem + (uint32_t)(kACol0 + slot * 64), a_lo = a_hi + 32;
                    const uint32_t b_hi = umma::smem_u32(S.b[slot]), b_lo = b_hi + sg.n_pad * 128;
#pragma unroll
                    for (int k = 0; k < kBK / 8; ++k) {
                        const uint64_t dbh = umma::make_desc_kmajor_sw128(b_hi + k * 32);
                        const uint64_t dbl = umma::make_desc_kmajor_sw128(b_lo + k * 32);
                        umma::mma_tf32_ts(d_addr, a_lo + k * 8, dbh, idesc, (st | k) != 0);
                        umma::mma_tf32_ts(d_addr, a_hi + k * 8, dbl, idesc, 1);
                        umma::mma_tf32_ts(d_addr, a_hi + k * 8, dbh, idesc, 1);
                    }
                    umma::commit(&S.empty[slot]);
                }
            }
            umma::commit(&S.d_full);
        }
    }
    __syncthreads();
    if (warp == 0) {
        umma::fence_after_sync();
        umma::tmem_dealloc<kTmemCols>(tmem);
    }
}

}  // namespace

// Generic entry: out[rows, n_valid] (+ column offsets) from up to 4 segments.  Exposed for tests and for the conv path.
extern "C" int jamun_gemm_tf32x3(int nseg, const float* const* a, const float* const* b, const int* n_stages, const int* n_pad,
                                 const int* n_valid, const int* out_col, const float* alpha, int rows, int rows_pad,
                                 const float* row_scale, float* out, int out_ld, jamun_stream_t stream) {
    JB_CHECK_ARG(nseg >= 1 && nseg <= 4 && a && b && n_stages && n_pad && n_valid && out_col && alpha && out, "bad argument");
    JB_CHECK_ARG(rows_pad % 128 == 0 && rows <= rows_pad, "rows_pad must be a multiple of 128");
    if (rows == 0) return JAMUN_OK;
    Params P{};
    P.nseg = nseg;
    P.rows = rows;
    P.rows_pad = rows_pad;
    P.out_ld = out_ld;
    P.row_scale = row_scale;
    int dcol = 0;
    for (int s = 0; s < nseg; ++s) {
        JB_CHECK_ARG(n_pad[s] % 16 == 0 && n_pad[s] >= 16 && n_pad[s] <= kMaxN && n_valid[s] <= n_pad[s], "n_pad out of range");
        P.seg[s] = Seg{a[s], b[s], out, n_stages[s], n_pad[s], n_valid[s], dcol, out_col[s], alpha[s]};
        dcol += n_pad[s];
    }
    JB_CHECK_ARG(dcol <= kACol0, "accumulators exceed 256 TMEM columns");
    const size_t smem = sizeof(Smem) + 1024;
    cudaError_t e = cudaFuncSetAttribute(gemm_tf32x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
        jb::set_error("jamun_gemm_tf32x3: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        return JAMUN_ECUDA;
    }
    gemm_tf32x3_kernel<<<rows_pad / 128, kThreads, smem, jb::as_stream(stream)>>>(P);
    JB_CHECK_LAUNCH();
    return JAMUN_OK;
}
