// Weight-gradient GEMM of the training path on the 5th-gen tensor cores:  dW[m, w] = sum_c sum_r A_c[m, r] * B_c[r, w]
//   m = (stage, u)  rows of a stage-major operand A ([stage][rows_pad][32], the layout jamun_gemm_tf32x3 consumes as its A),
//   r = node (the reduction index),  w < W <= 160 columns of a row-major matrix (scaled output gradients G, or the block input).
// This is the "A^T . B" product of SURVEY Appendix D (dM = A^T dO): the *reduction* runs over the rows of both operands, so
// both are MN-major for the tensor core.  A goes through registers (column gather -> TMEM), so only B needs a tensor-core
// shared-memory layout: jamun_pack_rows_split writes it in the UMMA MN-major SWIZZLE_128B_BASE32B canonical form (the one
// layout tcgen05 takes for MN-major tf32: a row is 128 bytes = 32 consecutive w, four rows form a 512-byte swizzle atom whose
// 32-byte pieces are XOR-ed with (r & 3)), so tiles arrive by plain 1-D bulk copies:
//   A loader   4 x 4 KB per K-stage (32 nodes): the four stage arrays of this CTA's 128-row M tile
//   converters thread (stage w, u) gathers its column of the landed tile (32 nodes, conflict-free), splits every value into
//              tf32 hi + exact remainder lo and writes both to TMEM with tcgen05.st -> TS-mode MMA, A^T is never materialised
//   B loader   the pre-split (hi | lo) images of B (jamun_pack_rows_split): n_slots x 4 KB each per K-stage
//   MMA        per 8-node k-step: D += A_lo.B_hi + A_hi.B_lo + A_hi.B_hi, B described MN-major (LBO = 4 KB between 32-column
//              blocks, SBO = 1 KB between 8-node groups); fp32 accumulation in TMEM
// The node range is split over gridDim.y CTAs (few M tiles, long K; it also bounds the number of MMAs accumulated into one
// TMEM tile, whose truncating adds cost ~S^1.5 2^-24 of accuracy); partial tiles are summed in ascending order by a second
// kernel, which also scatters rows to the gradient's layout.  No atomics: bit-reproducible.
#include "common.cuh"
#include "umma.cuh"

namespace {
using namespace jb;

constexpr int kTSlots = 4;
constexpr int kASlots = 6;
constexpr int kBSlots = 3;
constexpr int kKRows = 32;                       // nodes per K-stage
constexpr int kATileBytes = 4 * kKRows * 128;    // 16 KB: four stage arrays x 32 rows x 128 B
constexpr int kMaxSlotsB = 5;                    // W <= 160
constexpr int kBSlotBytes = 2 * kMaxSlotsB * kKRows * 128;  // hi + lo, 40 KB
constexpr int kACol0 = 256;
constexpr int kConvWarps = 8;
constexpr int kThreads = (kConvWarps + 3) * 32;

struct AtbTcParams {
    const float* a;           // stage-major operand, component c at a + c * a_comp_stride
    long long a_comp_stride;
    const float* bsplit;      // [ncomp][2 (hi|lo)][nslots_b][rows_pad][32]
    int ncomp, n_stages, rows, rows_pad, nslots_b, n_pad;
    int k_splits;
    float* partial;           // [k_splits][m_tiles*128][n_pad]
};

struct __align__(1024) Smem {
    uint8_t b[kBSlots][kBSlotBytes];
    uint8_t a[kASlots][kATileBytes];
    uint64_t a_full[kASlots], a_empty[kASlots], b_full[kBSlots], b_empty[kBSlots], t_full[kTSlots], t_empty[kTSlots], d_full;
    uint32_t tmem_base;
};

__host__ __device__ constexpr uint32_t make_idesc_tf32_bmn(int M, int N) {  // kind::tf32, A K-major (TMEM), B MN-major
    return (1u << 4) | (2u << 7) | (2u << 10) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__global__ void __launch_bounds__(kThreads, 1) gemm_atb_kernel(const AtbTcParams P) {
    extern __shared__ uint8_t smem_raw[];
    Smem& S = *reinterpret_cast<Smem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int warp = threadIdx.x >> 5;
    const int mt = blockIdx.x, ky = blockIdx.y;
    if (threadIdx.x == 0) {
        for (int s = 0; s < kASlots; ++s) {
            umma::mbar_init(&S.a_full[s], 1);
            umma::mbar_init(&S.a_empty[s], 128);
        }
        for (int s = 0; s < kBSlots; ++s) {
            umma::mbar_init(&S.b_full[s], 1);
            umma::mbar_init(&S.b_empty[s], 1);
        }
        for (int s = 0; s < kTSlots; ++s) {
            umma::mbar_init(&S.t_full[s], 128);
            umma::mbar_init(&S.t_empty[s], 1);
        }
        umma::mbar_init(&S.d_full, 1);
        umma::fence_barrier_init();
    }
    if (warp == 0) umma::tmem_alloc<512>(&S.tmem_base);
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = S.tmem_base;

    // K-stages of this split: q in [q_lo, q_hi) over (component, 32-node tile), component-major
    const int tiles_per_comp = (P.rows + kKRows - 1) / kKRows;
    const int total_q = P.ncomp * tiles_per_comp;
    const int q_lo = (int)((long long)total_q * ky / P.k_splits), q_hi = (int)((long long)total_q * (ky + 1) / P.k_splits);
    const int nq = q_hi - q_lo;

    if (warp < kConvWarps) {
        // ------------------------------------------------ converters: column gather of the landed tile -> (hi, lo) in TMEM
        const int grp = warp >> 2, wq = warp & 3, lane = threadIdx.x & 31;
        const uint32_t lane_base = (uint32_t)(wq * 32) << 16;
        for (int g = grp; g < nq; g += kConvWarps / 4) {
            const int sa = g % kASlots, st = g % kTSlots;
            const int r0 = ((q_lo + g) % tiles_per_comp) * kKRows;
            umma::mbar_wait(&S.a_full[sa], (g / kASlots) & 1);
            const uint8_t* tile = S.a[sa] + wq * (kKRows * 128);
            uint32_t hi[32], lo[32];
#pragma unroll
            for (int k = 0; k < 32; ++k) {
                float v = *reinterpret_cast<const float*>(tile + k * 128 + ((((lane >> 2) ^ (k & 7)) << 4) | ((lane & 3) << 2)));
                if (r0 + k >= P.rows) v = 0.f;  // rows of the padded operand that were never written
                const uint32_t h = __float_as_uint(v) & 0xFFFFE000u;
                hi[k] = h;
                lo[k] = __float_as_uint(v - __uint_as_float(h));
            }
            umma::mbar_arrive(&S.a_empty[sa]);
            umma::mbar_wait(&S.t_empty[st], ((g / kTSlots) & 1) ^ 1);
            umma::fence_after_sync();
            const uint32_t a_addr = tmem + lane_base + (uint32_t)(kACol0 + st * 64);
            umma::tmem_st32(a_addr, hi);
            umma::tmem_st32(a_addr + 32, lo);
            umma::wait_st();
            umma::fence_before_sync();
            umma::mbar_arrive(&S.t_full[st]);
        }
        // ------------------------------------------------ epilogue: TMEM lane = row m of the tile -> partial[ky][m][:]
        umma::mbar_wait(&S.d_full, 0);
        umma::fence_after_sync();
        const int m = mt * 128 + wq * 32 + lane;
        float* o = P.partial + ((size_t)ky * gridDim.x * 128 + m) * P.n_pad;
        int chunk = 0;
        for (int c0 = 0; c0 < P.n_pad; c0 += 32, ++chunk) {
            if ((chunk & 1) != grp) continue;
            uint32_t v[32];
            umma::tmem_ld32(tmem + lane_base + (uint32_t)c0, v);
            umma::wait_ld();
            if (nq == 0) {
#pragma unroll
                for (int q = 0; q < 32; ++q) v[q] = 0u;
            }
#pragma unroll
            for (int q = 0; q < 8; ++q)
                if (c0 + 4 * q < P.n_pad)
                    *reinterpret_cast<float4*>(o + c0 + 4 * q) = make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]),
                                                                            __uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3]));
        }
        umma::fence_before_sync();
    } else if (warp == kConvWarps) {
        // ------------------------------------------------ A loader: four 4 KB bulk copies per K-stage
        for (int g = 0; g < nq; ++g) {
            const int q = q_lo + g, c = q / tiles_per_comp, r0 = (q % tiles_per_comp) * kKRows;
            const int sa = g % kASlots;
            umma::mbar_wait(&S.a_empty[sa], ((g / kASlots) & 1) ^ 1);
            if (umma::elect_one()) {
                umma::mbar_arrive_expect_tx(&S.a_full[sa], kATileBytes);
                const float* base = P.a + (size_t)c * P.a_comp_stride;
#pragma unroll
                for (int w = 0; w < 4; ++w) {
                    const int stage = min(4 * mt + w, P.n_stages - 1);  // a short last tile re-reads its last stage (masked at output)
                    umma::bulk_g2s(S.a[sa] + w * (kKRows * 128), base + ((size_t)stage * P.rows_pad + r0) * 32, kKRows * 128, &S.a_full[sa]);
                }
            }
            __syncwarp();
        }
    } else if (warp == kConvWarps + 1) {
        // ------------------------------------------------ B loader: (hi | lo) x nslots_b x 4 KB per K-stage
        const uint32_t bytes = 2u * P.nslots_b * (kKRows * 128);
        for (int g = 0; g < nq; ++g) {
            const int q = q_lo + g, c = q / tiles_per_comp, r0 = (q % tiles_per_comp) * kKRows;
            const int sb = g % kBSlots;
            umma::mbar_wait(&S.b_empty[sb], ((g / kBSlots) & 1) ^ 1);
            if (umma::elect_one()) {
                umma::mbar_arrive_expect_tx(&S.b_full[sb], bytes);
                for (int h = 0; h < 2; ++h)
                    for (int j = 0; j < P.nslots_b; ++j) {
                        const float* src = P.bsplit + ((((size_t)c * 2 + h) * P.nslots_b + j) * P.rows_pad + r0) * 32;
                        umma::bulk_g2s(S.b[sb] + (h * kMaxSlotsB + j) * (kKRows * 128), src, kKRows * 128, &S.b_full[sb]);
                    }
            }
            __syncwarp();
        }
    } else {
        // ------------------------------------------------ MMA issuer
        const uint32_t idesc = make_idesc_tf32_bmn(128, P.n_pad);
        // MN-major SWIZZLE_128B_BASE32B descriptor (layout type 1): a swizzle atom is 4 nodes x 128 B; LBO = 4096 B between
        // 32-column blocks, SBO = 512 B between 4-node groups; one K = 8 instruction spans two atoms (1 KB)
        constexpr uint32_t kLbo = (uint32_t)(kKRows * 128) >> 4;
        constexpr uint32_t kDescHiMnSw128B32 = (512u >> 4) | (1u << 14) | (1u << 29);
        const uint32_t lo_off = (uint32_t)(kMaxSlotsB * kKRows * 128) >> 4;
        for (int g = 0; g < nq; ++g) {
            const int ts = g % kTSlots, sb = g % kBSlots;
            umma::mbar_wait(&S.t_full[ts], (g / kTSlots) & 1);
            umma::mbar_wait(&S.b_full[sb], (g / kBSlots) & 1);
            umma::fence_after_sync();
            if (umma::elect_one()) {
                const uint32_t a_hi = tmem + (uint32_t)(kACol0 + ts * 64), a_lo = a_hi + 32;
                const uint32_t bh = ((umma::smem_u32(S.b[sb]) >> 4) & 0x3FFF) | (kLbo << 16), bl = bh + lo_off;
#pragma unroll
                for (int k = 0; k < kKRows / 8; ++k) {
                    const uint64_t dbh = umma::make_desc(bh + 64 * k, kDescHiMnSw128B32);  // + k * 1024 B
                    const uint64_t dbl = umma::make_desc(bl + 64 * k, kDescHiMnSw128B32);
                    umma::mma_tf32_ts(tmem, a_lo + k * 8, dbh, idesc, (g != 0) || k != 0);
                    umma::mma_tf32_ts(tmem, a_hi + k * 8, dbl, idesc, 1);
                    umma::mma_tf32_ts(tmem, a_hi + k * 8, dbh, idesc, 1);
                }
                umma::commit(&S.t_empty[ts]);
                umma::commit(&S.b_empty[sb]);
            }
            __syncwarp();
        }
        if (umma::elect_one()) umma::commit(&S.d_full);
        __syncwarp();
    }
    __syncthreads();
    if (warp == 0) {
        umma::fence_after_sync();
        umma::tmem_dealloc<512>(tmem);
    }
}

// out <- sum over the K splits (ascending) of partial rows, scattered to the gradient layout (same modes as jamun_stage_atb)
struct AtbOut {
    float* out;
    int mode, out_rows, W, nslots, n_stages, n_pad, k_splits, m_rows;
    int slot_row0[8], slot_rows[8];
};
__global__ void gemm_atb_reduce_kernel(const float* __restrict__ partial, const AtbOut O) {
    const long long total = (long long)O.n_stages * 32 * O.W;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int w = (int)(t % O.W);
        const int m = (int)(t / O.W);  // = stage*32 + u
        const int stage = m >> 5, u = m & 31;
        const int k = stage / O.nslots, slot = stage - k * O.nslots;
        float acc = 0.f;
        for (int y = 0; y < O.k_splits; ++y) acc += partial[((size_t)y * O.m_rows + m) * O.n_pad + w];
        if (O.mode == 0) {
            if (u < O.slot_rows[slot]) O.out[((size_t)k * O.out_rows + O.slot_row0[slot] + u) * O.W + w] = acc;
        } else {
            O.out[((size_t)k * O.out_rows + w) * 32 + u] = acc;
        }
    }
}

// x[:, col0 : col0+ncols] -> (hi | lo) stage-major, chunk-swizzled images [2][nslots][rows_pad][32]; rows >= rows and columns
// >= ncols are zero
__global__ void pack_rows_split_kernel(const float* __restrict__ x, int ld, int col0, int ncols, int rows, int rows_pad, int nslots,
                                       float* __restrict__ out) {
    const size_t per = (size_t)nslots * rows_pad * 32;
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < per; t += (size_t)gridDim.x * blockDim.x) {
        const int lane = (int)(t & 31);
        const size_t sr = t >> 5;
        const int row = (int)(sr % rows_pad), slot = (int)(sr / rows_pad);
        const int c = slot * 32 + lane;
        const float v = (row < rows && c < ncols) ? x[(size_t)row * ld + col0 + c] : 0.f;
        const float hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
        // SWIZZLE_128B_BASE32B (the only shared-memory layout tcgen05 accepts for MN-major tf32): 32-byte pieces XOR (row & 3)
        const int swz = (((lane >> 3) ^ (row & 3)) << 3) | (lane & 7);
        out[sr * 32 + swz] = hi;
        out[per + sr * 32 + swz] = v - hi;
    }
}

}  // namespace

extern "C" int jamun_pack_rows_split(const float* x, int ld, int col0, int ncols, int rows, int rows_pad, int nslots, float* out,
                                     jamun_stream_t stream) {
    JB_CHECK_ARG(x && out && rows <= rows_pad && rows_pad % 32 == 0 && nslots * 32 >= ncols, "bad argument");
    const size_t per = (size_t)nslots * rows_pad * 32;
    if (per == 0) return JAMUN_OK;
    size_t blocks = (per + 255) / 256;
    if (blocks > (size_t)jb::kNumSMs * 16) blocks = (size_t)jb::kNumSMs * 16;
    pack_rows_split_kernel<<<(int)blocks, 256, 0, jb::as_stream(stream)>>>(x, ld, col0, ncols, rows, rows_pad, nslots, out);
    JB_CHECK_LAUNCH();
    return JAMUN_OK;
}

extern "C" long long jamun_stage_atb_tc_scratch(int n_stages, int W, int k_splits) {
    const int m_tiles = (n_stages + 3) / 4, n_pad = (W + 15) / 16 * 16;
    return (long long)k_splits * m_tiles * 128 * n_pad;
}

// Tensor-core form of jamun_stage_atb.  bsplit: ncomp images from jamun_pack_rows_split ([ncomp][2][nslots_b][rows_pad][32],
// nslots_b = ceil(W/32)); partial: jamun_stage_atb_tc_scratch(n_stages, W, k_splits) floats.
extern "C" int jamun_stage_atb_tc(const float* a, long long a_comp_stride, int ncomp, int n_stages, int nslots, int rows, int rows_pad,
                                  const float* bsplit, int W, float* out, int mode, int out_rows, const int* slot_row0,
                                  const int* slot_rows, int k_splits, float* partial, jamun_stream_t stream) {
    JB_CHECK_ARG(a && bsplit && out && partial, "null argument");
    JB_CHECK_ARG(W >= 1 && W <= 160 && nslots >= 1 && nslots <= 8 && n_stages % nslots == 0 && ncomp >= 1, "bad shape");
    JB_CHECK_ARG(rows_pad % 32 == 0 && rows <= rows_pad && k_splits >= 1, "rows_pad must be a multiple of 32");
    JB_CHECK_ARG(mode == 1 || (slot_row0 && slot_rows), "mode 0 needs the slot tables (host pointers)");
    if (n_stages == 0) return JAMUN_OK;
    cudaStream_t s = jb::as_stream(stream);
    AtbTcParams P{};
    P.a = a, P.a_comp_stride = a_comp_stride, P.bsplit = bsplit, P.ncomp = ncomp, P.n_stages = n_stages, P.rows = rows;
    P.rows_pad = rows_pad, P.nslots_b = (W + 31) / 32, P.n_pad = (W + 15) / 16 * 16, P.k_splits = k_splits, P.partial = partial;
    const int m_tiles = (n_stages + 3) / 4;
    const size_t smem = sizeof(Smem) + 1024;
    static_assert(sizeof(Smem) + 1024 <= 227 * 1024, "shared memory budget");
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(gemm_atb_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) {
            jb::set_error("jamun_stage_atb_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
            return JAMUN_ECUDA;
        }
        attr_set = true;
    }
    gemm_atb_kernel<<<dim3(m_tiles, k_splits), kThreads, smem, s>>>(P);
    AtbOut O{};
    O.out = out, O.mode = mode, O.out_rows = out_rows, O.W = W, O.nslots = nslots, O.n_stages = n_stages, O.n_pad = P.n_pad;
    O.k_splits = k_splits, O.m_rows = m_tiles * 128;
    for (int q = 0; q < nslots && mode == 0; ++q) O.slot_row0[q] = slot_row0[q], O.slot_rows[q] = slot_rows[q];
    const long long total = (long long)n_stages * 32 * W;
    int blocks = (int)((total + 255) / 256);
    if (blocks > jb::kNumSMs * 16) blocks = jb::kNumSMs * 16;
    gemm_atb_reduce_kernel<<<blocks, 256, 0, s>>>(partial, O);
    JB_CHECK_LAUNCH();
    return JAMUN_OK;
}
