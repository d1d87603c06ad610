// Node-tile GEMM on the 5th-gen tensor cores: out[rows, N] = scale_row * alpha * A[rows, K] . B[K, N] with fp32 accuracy
// from three TF32 products (a_hi*b_hi + a_hi*b_lo + a_lo*b_hi, fp32 accumulation in TMEM).
//
// One CTA owns 128 rows (nodes).  Roles (352 threads; warp numbers below are for one converter group):
//   warp 4     A loader: one thread issues a 16 KB 1-D bulk async copy per stage (the tile's 128 rows x 128 B are contiguous
//              in the stage-major fp32 A layout) into a 6-deep shared-memory ring -- HBM latency is hidden by the copy engine.
//   warps 0-3  converters: thread r reads row r of the landed tile (bank-conflict-free thanks to the XOR chunk swizzle the
//              builder applied), splits each value into tf32 hi / exact remainder lo and writes both straight into TMEM
//              with tcgen05.st (TS-mode MMA: the A operand is never re-read from shared memory).  Then the epilogue.
//   warp 5     B loader: one bulk async copy per stage of the pre-swizzled (hi|lo) weight image (packed once on the host
//              in the UMMA K-major SWIZZLE_128B layout) into a 3-deep shared-memory ring.
//   warp 6     MMA issuer: one thread issues 4 k-steps x 3 tcgen05.mma (kind::tf32, M=128, N=n_pad) per stage and commits
//              to the stage's "empty" mbarriers.
// Up to four K-segments, each with its own A, B, N and TMEM accumulator columns, run back to back in one launch
// (conv: one 0e segment with N=160 and three 1e segments with N=32).
//
// F16 variant (jamun_gemm_f16x3): the same pipeline with the split taken in fp16 instead of tf32 -- both formats carry an 11-bit
// significand, so a_hi*b_hi + a_hi*b_lo + a_lo*b_hi keeps the same ~2^-21 per-product accuracy, but kind::f16 contracts K = 16
// per instruction at twice the tf32 rate: half the MMA instructions per stage.  fp16's range is the price: weights are
// pre-scaled by a power of two when their images are packed (jamun_pack_b_f16; undone through `alpha`), and an operand value
// beyond +-65504 raises bit 0 of the caller's status word (the caller then has to use the tf32 kernel).  A converter thread
// packs its row as 16 (hi,hi) + 16 (lo,lo) half2 words -> one 32-column TMEM slot; a weight stage image is n_pad rows of
// 128 bytes [hi k0..31 | lo k0..31] (K-major SWIZZLE_128B), half the bytes of the tf32 image.
#include <cuda_fp16.h>
#include <stdlib.h>
#include "common.cuh"
#include "umma.cuh"

namespace {
using namespace jb;

constexpr int kTSlots = 4;                 // A slots in TMEM
#ifndef JAMUN_GEMM_ASLOTS
#define JAMUN_GEMM_ASLOTS 6
#define JAMUN_GEMM_BSLOTS 3
#endif
constexpr int kBK = 32;                    // K per stage (one 128-byte row of the fp32 A tile)
constexpr int kMaxN = 160;
constexpr int kATileBytes = 128 * kBK * 4;    // 16 KB
template <bool F16>
struct Cfg {
    static constexpr int kASlots = F16 ? 8 : JAMUN_GEMM_ASLOTS;  // raw fp32 A tiles in shared memory (bulk-copied from HBM)
    static constexpr int kBSlots = F16 ? 4 : JAMUN_GEMM_BSLOTS;  // weight images in shared memory
    static constexpr int kBRowBytes = F16 ? 128 : 256;           // image bytes per output column and stage (hi + lo)
    static constexpr int kBSlotBytes = kMaxN * kBRowBytes;       // 20 KB (fp16) / 40 KB (tf32)
    static constexpr int kASlotCols = F16 ? 32 : 64;             // TMEM columns of one converted A stage (hi | lo)
};
constexpr int kTmemCols = 512;
constexpr int kACol0 = 256;                // A slots live in TMEM columns [256, 512): 64 columns (hi 32 | lo 32) each
constexpr int kConvWarps = 8;              // two converter groups of 4 warps (TMEM lane quarters), alternating stages
constexpr int kThreads = (kConvWarps + 3) * 32;  // + A loader, B loader, MMA issuer

struct Seg {
    const float* a;       // [n_stages][rows_pad][32], 16-byte chunks of a row XOR-swizzled with (row & 7)
    const float* b;       // [n_stages][2][n_pad*32] swizzled images
    float* out;           // output base (row-major, ld = out_ld)
    const float* addend;  // optional [rows, addend_ld]: out = (acc + addend[row, n]) * alpha * row_scale
    int n_stages, n_pad, n_valid, d_col, out_col, addend_ld;
    float alpha;
    float addend_scale;   // out = (acc + addend_scale * addend) * alpha * row_scale (fp16 form: the weights' pre-scale)
};
struct Params {
    Seg seg[4];
    int nseg, rows, rows_pad, out_ld;
    // column blocks (same A, passes over B): pass y uses b += y * b_block_floats, out_col += y * n_valid (all segments)
    long long b_block_floats;
    int col_blocks;
    const float* row_scale;  // [rows] or null
    int coalesce;            // stationary mode: transpose output blocks through shared memory (full-line stores)
    // split-K (small row counts: few tiles): CTA (tile, y) accumulates stages [n*y/k_splits, n*(y+1)/k_splits) of every segment
    // and writes its scaled partial result to partial + y * partial_stride; gemm_splitk_reduce_kernel sums them in order
    int cb_groups;  // column-block passes spread over gridDim.y CTAs per tile (few tiles: the passes of one tile run in parallel)
    int k_splits;
    float* partial;
    long long partial_stride;
    int* status;  // F16 only: bit 0 is set when an A value does not fit fp16 (null: not reported)
    int a_tile_major;  // A is [row / 128][stage][128][32] (a tile's stages contiguous) instead of [stage][rows_pad][32]
    jamun_gemm_epilogue epi;  // fused ConvBlock epilogues (mode 0: plain store)
};

template <bool F16>
struct __align__(1024) SmemT {
    uint8_t b[Cfg<F16>::kBSlots][Cfg<F16>::kBSlotBytes];
    uint8_t a[Cfg<F16>::kASlots][kATileBytes];
    uint64_t a_full[Cfg<F16>::kASlots], a_empty[Cfg<F16>::kASlots], b_full[Cfg<F16>::kBSlots], b_empty[Cfg<F16>::kBSlots],
        t_full[kTSlots], t_empty[kTSlots], d_full[2], d_empty[2];
    uint32_t tmem_base;
};

__device__ __forceinline__ uint32_t h2_bits(__half2 h) { return *reinterpret_cast<uint32_t*>(&h); }

// one 128-byte row of a stage-major operand ([stage][rows_pad][32], 16-byte chunks XOR-ed with row & 7) from 32 registers
__device__ __forceinline__ void store_op_row(float* row_base, int sw, const float (&f)[32]) {
#pragma unroll
    for (int q = 0; q < 8; ++q)
        *reinterpret_cast<float4*>(row_base + ((q ^ sw) << 2)) = make_float4(f[4 * q], f[4 * q + 1], f[4 * q + 2], f[4 * q + 3]);
}

// Fused epilogue of the ConvBlock contraction (mode 1): Gate (e3tools/nn/_gate.py:63: LeakyReLU on the 120 scalars, sigmoid
// gates on the vectors) applied to the accumulators and written straight into the block-tail GEMM's operands -- the conv
// output never reaches memory.  Segment 0 = [120 scalars | 32 gate scalars | pad], segments 1-3 = the 1e components (their
// addend is the receiver-side sum of the 0e(x)1e->1e path).  `row` is the operand row (== output row of this launch).
__device__ __forceinline__ void epi_gate(const Params& P, int s, int c0, const uint32_t (&v)[32], int row, bool live, float rs_own,
                                         uint32_t tmem_lane) {
    const jamun_gemm_epilogue& E = P.epi;
    const Seg& sg = P.seg[s];
    const float sc = sg.alpha * rs_own;
    const int sw = row & 7;
    float f[32];
    if (s == 0) {
        if (c0 >= 128 || !live) return;  // the gate scalars are consumed by the vector chunks below
#pragma unroll
        for (int q = 0; q < 32; ++q) {
            const float x = __uint_as_float(v[q]) * sc;
            f[q] = (c0 + q < JAMUN_S) ? E.c_act * (x > 0.f ? x : 0.01f * x) : 0.f;
        }
        store_op_row(E.op_s + ((size_t)(c0 >> 5) * E.op_rows_pad + row) * 32, sw, f);
    } else {
        uint32_t g[32];
        umma::tmem_ld32(tmem_lane + (uint32_t)(P.seg[0].d_col + JAMUN_S), g);  // this row's 32 gate pre-activations
        umma::wait_ld();
        if (!live) return;
        const float sc0 = P.seg[0].alpha * rs_own;
        const float* ad = sg.addend ? sg.addend + (size_t)row * sg.addend_ld : nullptr;
#pragma unroll
        for (int q = 0; q < 32; ++q) {
            float x = __uint_as_float(v[q]);
            if (ad) x += ad[q] * sg.addend_scale;
            f[q] = x * sc * (E.c_gate * sigmoidf_acc(__uint_as_float(g[q]) * sc0));
        }
        store_op_row(E.op_v + (size_t)(s - 1) * E.op_v_comp_stride + (size_t)row * 32, sw, f);
    }
}

// Fused epilogue of the block-tail GEMM (mode 2): noise-conditional skip x_new = x_res*w + y*(1-w) and the next block's
// input scaling x_scaled = x_new * s_next (model/noise_conditioning.py:50-73, arch/e3conv.py:131-133), written row-major for
// the gathers of the next aggregate and, packed, as the x_in halves of the next block-tail operands (scalars: stages 4-7 of
// op_s, which are also the per-node transform's operand; vectors: stage 1 of op_v).  Weights are per irrep: [120 | 32].
// The warp's 32 x 32 accumulator block (thread = row) is transposed through shared memory first, so that every global access
// below is a float4 with eight lanes covering one 128-byte piece of a row: 4 rows per instruction instead of 32 scattered
// 16-byte pieces.  Called by all 32 lanes (rows beyond the batch only skip the memory accesses).
__device__ __forceinline__ void epi_mix(const Params& P, int s, int c0, const uint32_t (&v)[32], int warp_row0, int lane, float* stg) {
    const jamun_gemm_epilogue& E = P.epi;
    const Seg& sg = P.seg[s];
    const int nval = sg.n_valid - c0 < 32 ? sg.n_valid - c0 : 32;  // 24 in the last scalar chunk
    if (nval <= 0) return;
#pragma unroll
    for (int q = 0; q < 8; ++q)
        *reinterpret_cast<float4*>(stg + lane * 32 + ((q ^ (lane & 7)) << 2)) =
            make_float4(__uint_as_float(v[4 * q]) * sg.alpha, __uint_as_float(v[4 * q + 1]) * sg.alpha,
                        __uint_as_float(v[4 * q + 2]) * sg.alpha, __uint_as_float(v[4 * q + 3]) * sg.alpha);
    __syncwarp();
    const int ch = lane & 7;                               // 16-byte piece of the 32-column chunk
    const int j0 = sg.out_col + c0 + 4 * ch;               // column of the [216]-wide node row
    const int w0 = (s == 0 ? c0 : JAMUN_S) + 4 * ch;       // per-irrep weight index
    const bool cols_live = 4 * ch < nval;
    float4 w4 = make_float4(0.f, 0.f, 0.f, 0.f), s4 = make_float4(1.f, 1.f, 1.f, 1.f);
    if (cols_live) {
        if (E.skip_w) w4 = *reinterpret_cast<const float4*>(E.skip_w + w0);
        if (E.s_next) s4 = *reinterpret_cast<const float4*>(E.s_next + w0);
    }
    const bool pack = E.x_scaled && E.op_s;
#pragma unroll
    for (int it = 0; it < 8; ++it) {
        const int rl = it * 4 + (lane >> 3);
        const int row = warp_row0 + rl;
        if (row >= P.rows) continue;
        float4 y = *reinterpret_cast<const float4*>(stg + rl * 32 + ((ch ^ (rl & 7)) << 2));
        float4 xs = make_float4(0.f, 0.f, 0.f, 0.f);
        if (cols_live) {
            if (E.skip_w) {
                const float4 r4 = *reinterpret_cast<const float4*>(E.x_res + (size_t)row * JAMUN_HID + j0);
                y.x = r4.x * w4.x + y.x * (1.0f - w4.x);
                y.y = r4.y * w4.y + y.y * (1.0f - w4.y);
                y.z = r4.z * w4.z + y.z * (1.0f - w4.z);
                y.w = r4.w * w4.w + y.w * (1.0f - w4.w);
            }
            *reinterpret_cast<float4*>(E.x_new + (size_t)row * JAMUN_HID + j0) = y;
            xs = make_float4(y.x * s4.x, y.y * s4.y, y.z * s4.z, y.w * s4.w);
            if (E.x_scaled) *reinterpret_cast<float4*>(E.x_scaled + (size_t)row * JAMUN_HID + j0) = xs;
        }
        if (pack) {
            float* dst = s == 0 ? E.op_s + ((size_t)(4 + (c0 >> 5)) * E.op_rows_pad + row) * 32
                                : E.op_v + (size_t)(s - 1) * E.op_v_comp_stride + ((size_t)E.op_rows_pad + row) * 32;
            *reinterpret_cast<float4*>(dst + ((ch ^ (row & 7)) << 2)) = xs;  // columns beyond the segment: zeros
        }
    }
    __syncwarp();
}

// COALESCE: epilogue variant for wide outputs in stationary mode (kept out of the contraction instantiation: +40 registers)
template <bool COALESCE, bool F16>
__global__ void __launch_bounds__(kThreads, 1) gemm_tf32x3_kernel(const Params P) {
    using Smem = SmemT<F16>;
    constexpr int kASlots = Cfg<F16>::kASlots, kBSlots = Cfg<F16>::kBSlots, kASlotCols = Cfg<F16>::kASlotCols;
    extern __shared__ uint8_t smem_raw[];
    Smem& S = *reinterpret_cast<Smem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int warp = threadIdx.x >> 5;
    const int tile_row0 = blockIdx.x * 128;

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
        for (int s = 0; s < 2; ++s) {
            umma::mbar_init(&S.d_full[s], 1);
            umma::mbar_init(&S.d_empty[s], kConvWarps * 32);
        }
        umma::fence_barrier_init();
    }
    if (warp == 0) umma::tmem_alloc<kTmemCols>(&S.tmem_base);
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = S.tmem_base;

    const int ksp = P.k_splits, ky = P.k_splits > 1 ? blockIdx.y : 0;
    // this CTA's column blocks: [cb_base, cb_base + ncb) -- `cb` below is relative to cb_base
    const int cb_base = P.cb_groups > 1 ? (int)((long long)P.col_blocks * blockIdx.y / P.cb_groups) : 0;
    const int ncb = P.cb_groups > 1 ? (int)((long long)P.col_blocks * (blockIdx.y + 1) / P.cb_groups) - cb_base : P.col_blocks;
    auto st_lo = [&](const Seg& sg) { return (int)((long long)sg.n_stages * ky / ksp); };
    auto st_hi = [&](const Seg& sg) { return (int)((long long)sg.n_stages * (ky + 1) / ksp); };
    int total_stages = 0;
    for (int s = 0; s < P.nseg; ++s) total_stages += st_hi(P.seg[s]) - st_lo(P.seg[s]);
    // A-stationary mode (wide per-node transforms: few K stages, many column-block passes): the converted operand stays in its
    // TMEM slots after pass 0; later passes only stream weight images and issue MMAs.
    const bool stationary = P.col_blocks > 1 && total_stages <= kTSlots;
    // ... and with one segment of N <= 128 the accumulator is double-buffered (TMEM columns [0,128) / [128,256)) so the MMAs of
    // pass p+1 overlap the epilogue of pass p.
    const bool dbuf = stationary && P.nseg == 1 && P.seg[0].n_pad <= 128;

    if (warp < kConvWarps) {
        // ------------------------------------------------ converters: smem fp32 tile -> (hi, lo) in TMEM
        const int grp = warp >> 2;              // group g handles stages == g (mod 2)
        const int r = threadIdx.x & 127;        // row within the tile == TMEM lane
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        const int sw = r & 7;
        const float rs_own = (P.row_scale && tile_row0 + r < P.rows) ? P.row_scale[tile_row0 + r] : 1.0f;
        for (int cb = 0; cb < ncb; ++cb) {
        for (int g = cb * total_stages + grp; g < (cb + 1) * total_stages && !(stationary && cb > 0); g += kConvWarps / 4) {
            const int sa = g % kASlots, st = g % kTSlots;
            umma::mbar_wait(&S.a_full[sa], (g / kASlots) & 1);
            const float4* row = reinterpret_cast<const float4*>(S.a[sa] + r * 128);
            const uint32_t a_addr = tmem + lane_base + (uint32_t)(kACol0 + st * kASlotCols);
            if constexpr (F16) {
                // columns 0-15: (hi[2j], hi[2j+1]) packed halves, columns 16-31: the remainders; one 32-column store
                uint32_t pk[32];
                uint32_t ovf = 0;
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const float4 c = row[q ^ sw];
                    const __half2 h01 = __floats2half2_rn(c.x, c.y), h23 = __floats2half2_rn(c.z, c.w);
                    const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
                    pk[2 * q] = h2_bits(h01);
                    pk[2 * q + 1] = h2_bits(h23);
                    pk[16 + 2 * q] = h2_bits(__floats2half2_rn(c.x - f01.x, c.y - f01.y));
                    pk[16 + 2 * q + 1] = h2_bits(__floats2half2_rn(c.z - f23.x, c.w - f23.y));
                    // exponent field all ones (inf / nan) in either half sets bit 15 / 31
                    ovf |= ((pk[2 * q] & 0x7C007C00u) + 0x04000400u) | ((pk[2 * q + 1] & 0x7C007C00u) + 0x04000400u);
                }
                if ((ovf & 0x80008000u) && P.status && tile_row0 + r < P.rows) atomicOr(P.status, 1);
                umma::mbar_arrive(&S.a_empty[sa]);
                umma::mbar_wait(&S.t_empty[st], ((g / kTSlots) & 1) ^ 1);
                umma::fence_after_sync();
                umma::tmem_st32(a_addr, pk);
            } else {
            uint32_t hi[32], lo[32];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const float4 c = row[q ^ sw];
                const float v[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const uint32_t h = __float_as_uint(v[t]) & 0xFFFFE000u;
                    hi[4 * q + t] = h;
                    lo[4 * q + t] = __float_as_uint(v[t] - __uint_as_float(h));
                }
            }
            umma::mbar_arrive(&S.a_empty[sa]);
            umma::mbar_wait(&S.t_empty[st], ((g / kTSlots) & 1) ^ 1);
            umma::fence_after_sync();
            umma::tmem_st32(a_addr, hi);
            umma::tmem_st32(a_addr + 32, lo);
            }
            umma::wait_st();
            umma::fence_before_sync();
            umma::mbar_arrive(&S.t_full[st]);
        }
        // ------------------------------------------------ epilogue: TMEM -> registers (thread = row) -> scaled 16-byte stores
        const int db = dbuf ? (cb & 1) : 0;
        umma::mbar_wait(&S.d_full[db], dbuf ? (cb >> 1) & 1 : cb & 1);
        umma::fence_after_sync();
        int chunk = 0;
        for (int s = 0; s < P.nseg; ++s) {
            const Seg& sg = P.seg[s];
            for (int c0 = 0; c0 < sg.n_pad; c0 += 32, ++chunk) {
                if ((chunk & 1) != grp) continue;  // the two converter groups alternate 32-column chunks
                uint32_t v[32];
                umma::tmem_ld32(tmem + lane_base + (uint32_t)(sg.d_col + db * 128 + c0), v);
                umma::wait_ld();
                if (COALESCE && stationary && !sg.addend && sg.n_valid == sg.n_pad && (P.out_ld & 3) == 0) {
                    // wide outputs (per-node transform): transpose the warp's 32 x 32 block through shared memory (the A ring's
                    // slots 4-5 are never used in stationary mode) so that every store instruction writes four full 128-byte
                    // lines instead of 32 scattered 16-byte pieces
                    const int lane = threadIdx.x & 31;
                    float* stg = reinterpret_cast<float*>(S.a[4]) + warp * 1024;
                    const float sc = sg.alpha * rs_own;
#pragma unroll
                    for (int q = 0; q < 8; ++q)
                        *reinterpret_cast<float4*>(stg + lane * 32 + ((q ^ (lane & 7)) << 2)) =
                            make_float4(__uint_as_float(v[4 * q]) * sc, __uint_as_float(v[4 * q + 1]) * sc,
                                        __uint_as_float(v[4 * q + 2]) * sc, __uint_as_float(v[4 * q + 3]) * sc);
                    __syncwarp();
                    const int ch = lane & 7;
                    float* obase = sg.out + sg.out_col + (cb_base + cb) * sg.n_valid + c0 + 4 * ch;
#pragma unroll
                    for (int it = 0; it < 8; ++it) {
                        const int rl = it * 4 + (lane >> 3);
                        const int row = tile_row0 + (warp & 3) * 32 + rl;
                        const float4 val = *reinterpret_cast<const float4*>(stg + rl * 32 + ((ch ^ (rl & 7)) << 2));
                        if (row < P.rows) *reinterpret_cast<float4*>(obase + (size_t)row * P.out_ld) = val;
                    }
                    __syncwarp();
                    continue;
                }
                // thread = row: 8 x 16-byte stores per 32-column chunk (few instructions; the 8 pieces of a 128-byte line
                // merge in L2)
                const int row = tile_row0 + r;
                if (P.epi.mode != 0) {
                    // (epi_gate reads TMEM with a warp-collective tcgen05.ld: every thread calls it, rows beyond the batch only
                    // skip the memory accesses)
                    if (P.epi.mode == 1) epi_gate(P, s, c0, v, row, row < P.rows, rs_own, tmem + lane_base);
                    else epi_mix(P, s, c0, v, tile_row0 + (warp & 3) * 32, threadIdx.x & 31, reinterpret_cast<float*>(S.a[4]) + warp * 1024);
                    continue;
                }
                if (row < P.rows) {
                    const float sc = sg.alpha * rs_own;
                    float* o = (ksp > 1 ? P.partial + (size_t)ky * P.partial_stride : sg.out) + (size_t)row * P.out_ld + sg.out_col +
                               (cb_base + cb) * sg.n_valid + c0;
                    if (st_hi(sg) == st_lo(sg)) {  // this split holds no stage of the segment: its accumulator was never written
#pragma unroll
                        for (int q = 0; q < 32; ++q) v[q] = 0u;
                    }
                    const float* ad = sg.addend ? sg.addend + (size_t)row * sg.addend_ld + c0 : nullptr;
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        float f[4] = {__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]), __uint_as_float(v[4 * q + 2]),
                                      __uint_as_float(v[4 * q + 3])};
                        if (c0 + 4 * q + 3 < sg.n_valid) {
                            if (ad) {
                                const float4 a4 = *reinterpret_cast<const float4*>(ad + 4 * q);
                                f[0] += a4.x * sg.addend_scale;
                                f[1] += a4.y * sg.addend_scale;
                                f[2] += a4.z * sg.addend_scale;
                                f[3] += a4.w * sg.addend_scale;
                            }
                            *reinterpret_cast<float4*>(o + 4 * q) = make_float4(f[0] * sc, f[1] * sc, f[2] * sc, f[3] * sc);
                        } else {
#pragma unroll
                            for (int t = 0; t < 4; ++t)
                                if (c0 + 4 * q + t < sg.n_valid) o[4 * q + t] = (f[t] + (ad ? ad[4 * q + t] * sg.addend_scale : 0.f)) * sc;
                        }
                    }
                }
            }
        }
        umma::fence_before_sync();
        umma::mbar_arrive(&S.d_empty[db]);  // accumulators drained: a later pass may overwrite them
        }  // column-block pass
    } else if (warp == kConvWarps) {
        // ------------------------------------------------ A loader: one 16 KB bulk copy per stage
        int g = 0;
        for (int cb = 0; cb < (stationary ? 1 : ncb); ++cb) {
            for (int s = 0; s < P.nseg; ++s) {
                const Seg& sg = P.seg[s];
                const int st_first = st_lo(sg);
                for (int st = st_first, st_end = st_hi(sg); st < st_end; ++st, ++g) {
                    const int sa = g % kASlots;
                    umma::mbar_wait(&S.a_empty[sa], ((g / kASlots) & 1) ^ 1);
                    if (umma::elect_one()) {
                        umma::mbar_arrive_expect_tx(&S.a_full[sa], kATileBytes);
                        const float* src = P.a_tile_major ? sg.a + ((size_t)blockIdx.x * sg.n_stages + st) * (128 * kBK)
                                                          : sg.a + ((size_t)st * P.rows_pad + tile_row0) * kBK;
                        umma::bulk_g2s(S.a[sa], src, kATileBytes, &S.a_full[sa]);
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp == kConvWarps + 1) {
        // ------------------------------------------------ B loader
        int g = 0;
        for (int cb = 0; cb < ncb; ++cb) {
            for (int s = 0; s < P.nseg; ++s) {
                const Seg& sg = P.seg[s];
                const uint32_t bytes = (uint32_t)sg.n_pad * Cfg<F16>::kBRowBytes;
                const int st_first = st_lo(sg);
                for (int st = st_first, st_end = st_hi(sg); st < st_end; ++st, ++g) {
                    const int sb = g % kBSlots;
                    umma::mbar_wait(&S.b_empty[sb], ((g / kBSlots) & 1) ^ 1);
                    if (umma::elect_one()) {
                        umma::mbar_arrive_expect_tx(&S.b_full[sb], bytes);
                        umma::bulk_g2s(S.b[sb], sg.b + (size_t)(cb_base + cb) * P.b_block_floats + (size_t)st * (bytes / 4), bytes, &S.b_full[sb]);
                    }
                    __syncwarp();
                }
            }
        }
    } else {
        // ------------------------------------------------ MMA issuer (whole warp converged; one elected lane issues)
        int g = 0;
        for (int cb = 0; cb < ncb; ++cb) {
            const int db = dbuf ? (cb & 1) : 0;
            if (dbuf ? cb >= 2 : cb > 0) {  // the epilogue that last used this accumulator buffer must have drained it
                umma::mbar_wait(&S.d_empty[db], dbuf ? ((cb - 2) >> 1) & 1 : (cb - 1) & 1);
                umma::fence_after_sync();
            }
            for (int s = 0; s < P.nseg; ++s) {
                const Seg& sg = P.seg[s];
                const uint32_t idesc = F16 ? umma::make_idesc_f16(128, sg.n_pad) : umma::make_idesc_tf32(128, sg.n_pad);
                const uint32_t d_addr = tmem + (uint32_t)(sg.d_col + db * 128);
                const uint32_t lo_off = (uint32_t)(sg.n_pad * 128) >> 4;
                const int st_first = st_lo(sg);
                for (int st = st_first, st_end = st_hi(sg); st < st_end; ++st, ++g) {
                    const int gl = g - cb * total_stages;  // stage index within the pass
                    const int ts = stationary ? gl : g % kTSlots, sb = g % kBSlots;
                    if (!stationary || cb == 0) umma::mbar_wait(&S.t_full[ts], stationary ? 0 : (g / kTSlots) & 1);
                    umma::mbar_wait(&S.b_full[sb], (g / kBSlots) & 1);
                    umma::fence_after_sync();
                    if (umma::elect_one()) {
                        const uint32_t a_hi = tmem + (uint32_t)(kACol0 + ts * kASlotCols), a_lo = a_hi + kASlotCols / 2;
                        const uint32_t bh = umma::desc_lo_kmajor_sw128(umma::smem_u32(S.b[sb]));
                        if constexpr (F16) {
                            // a row of the image: [hi k0..31 | lo k0..31] halves; a k-step is 16 halves = 32 bytes = 8 TMEM columns
#pragma unroll
                            for (int k = 0; k < kBK / 16; ++k) {
                                const uint64_t dbh = umma::make_desc(bh + 2 * k, umma::kDescHiKmajorSw128);
                                const uint64_t dbl = umma::make_desc(bh + 4 + 2 * k, umma::kDescHiKmajorSw128);
                                umma::mma_f16_ts(d_addr, a_lo + k * 8, dbh, idesc, (st != st_first) || k != 0);
                                umma::mma_f16_ts(d_addr, a_hi + k * 8, dbl, idesc, 1);
                                umma::mma_f16_ts(d_addr, a_hi + k * 8, dbh, idesc, 1);
                            }
                        } else {
                        const uint32_t bl = bh + lo_off;
#pragma unroll
                        for (int k = 0; k < kBK / 8; ++k) {
                            const uint64_t dbh = umma::make_desc(bh + 2 * k, umma::kDescHiKmajorSw128);
                            const uint64_t dbl = umma::make_desc(bl + 2 * k, umma::kDescHiKmajorSw128);
                            umma::mma_tf32_ts(d_addr, a_lo + k * 8, dbh, idesc, (st != st_first) || k != 0);
                            umma::mma_tf32_ts(d_addr, a_hi + k * 8, dbl, idesc, 1);
                            umma::mma_tf32_ts(d_addr, a_hi + k * 8, dbh, idesc, 1);
                        }
                        }
                        if (!stationary) umma::commit(&S.t_empty[ts]);
                        umma::commit(&S.b_empty[sb]);
                    }
                    __syncwarp();
                }
            }
            if (umma::elect_one()) umma::commit(&S.d_full[db]);
            __syncwarp();
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
namespace {
// out[row, c] = sum_y partial[y][row, c] over the segments' column ranges, y ascending (deterministic)
__global__ void gemm_splitk_reduce_kernel(const float* __restrict__ partial, long long stride, int k_splits, int rows, int out_ld,
                                          int4 c0, int4 nc, int nseg, float* __restrict__ out) {
    const int col0[4] = {c0.x, c0.y, c0.z, c0.w}, ncol[4] = {nc.x, nc.y, nc.z, nc.w};
    int width = 0;
    for (int s = 0; s < nseg; ++s) width += ncol[s];
    const long long total = (long long)rows * width;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int row = (int)(t / width);
        int c = (int)(t % width), s = 0;
        while (c >= ncol[s]) c -= ncol[s++];
        const size_t off = (size_t)row * out_ld + col0[s] + c;
        float acc = 0.f;
        for (int y = 0; y < k_splits; ++y) acc += partial[(size_t)y * stride + off];
        out[off] = acc;
    }
}
}  // namespace

static int gemm_launch(bool f16, int nseg, const float* const* a, const float* const* b, const int* n_stages, const int* n_pad,
                       const int* n_valid, const int* out_col, const float* alpha, const float* const* addend,
                       const int* addend_ld, int col_blocks, long long b_block_floats, int rows, int rows_pad,
                       const float* row_scale, float* out, int out_ld, int k_splits, float* partial, int* status,
                       const float* addend_scale, int a_tile_major, const jamun_gemm_epilogue* epi, jamun_stream_t stream);

extern "C" int jamun_gemm_tf32x3(int nseg, const float* const* a, const float* const* b, const int* n_stages, const int* n_pad,
                                 const int* n_valid, const int* out_col, const float* alpha, const float* const* addend,
                                 const int* addend_ld, int col_blocks, long long b_block_floats, int rows, int rows_pad,
                                 const float* row_scale, float* out, int out_ld, jamun_stream_t stream) {
    return gemm_launch(false, nseg, a, b, n_stages, n_pad, n_valid, out_col, alpha, addend, addend_ld, col_blocks, b_block_floats,
                       rows, rows_pad, row_scale, out, out_ld, 1, nullptr, nullptr, nullptr, 0, nullptr, stream);
}

// Split-K form for small row counts (few 128-row tiles): k_splits CTAs per tile, partial: [k_splits, rows, out_ld] scratch.
extern "C" int jamun_gemm_tf32x3_splitk(int nseg, const float* const* a, const float* const* b, const int* n_stages,
                                        const int* n_pad, const int* n_valid, const int* out_col, const float* alpha, int rows,
                                        int rows_pad, const float* row_scale, float* out, int out_ld, int k_splits, float* partial,
                                        jamun_stream_t stream) {
    JB_CHECK_ARG(k_splits >= 1 && k_splits <= 64 && (k_splits == 1 || partial), "bad k_splits / partial");
    return gemm_launch(false, nseg, a, b, n_stages, n_pad, n_valid, out_col, alpha, nullptr, nullptr, 1, 0, rows, rows_pad, row_scale,
                       out, out_ld, k_splits, partial, nullptr, nullptr, 0, nullptr, stream);
}

// fp16-split form (same A operand; B images from jamun_pack_b_f16; k_splits == 1: plain launch, partial unused).
extern "C" int jamun_gemm_f16x3(int nseg, const float* const* a, const float* const* b, const int* n_stages, const int* n_pad,
                                const int* n_valid, const int* out_col, const float* alpha, const float* const* addend,
                                const int* addend_ld, const float* addend_scale, int col_blocks, long long b_block_floats, int rows,
                                int rows_pad, const float* row_scale, float* out, int out_ld, int k_splits, float* partial,
                                int* status, int a_tile_major, jamun_stream_t stream) {
    JB_CHECK_ARG(k_splits >= 1 && k_splits <= 64 && (k_splits == 1 || (partial && !addend && col_blocks == 1)), "bad k_splits / partial");
    return gemm_launch(true, nseg, a, b, n_stages, n_pad, n_valid, out_col, alpha, addend, addend_ld, col_blocks, b_block_floats, rows,
                       rows_pad, row_scale, out, out_ld, k_splits, partial, status, addend_scale, a_tile_major, nullptr, stream);
}

// jamun_gemm_f16x3 with one of the fused ConvBlock epilogues (include/jamun_b200.h: jamun_gemm_epilogue); `out` is not
// written (may be NULL).  mode 1 expects the contraction's segments [160 | 32 | 32 | 32], mode 2 the block tail's
// [128 | 32 | 32 | 32] with out_col [0, 120, 152, 184].  Single pass only: no split-K, no column blocks.
extern "C" int jamun_gemm_f16x3_fused(int nseg, const float* const* a, const float* const* b, const int* n_stages, const int* n_pad,
                                      const int* n_valid, const int* out_col, const float* alpha, const float* const* addend,
                                      const int* addend_ld, const float* addend_scale, int rows, int rows_pad,
                                      const float* row_scale, int* status, int a_tile_major, const jamun_gemm_epilogue* epi,
                                      jamun_stream_t stream) {
    JB_CHECK_ARG(epi && (epi->mode == 1 || epi->mode == 2) && nseg == 4 && n_pad && n_valid, "bad epilogue / segments");
    JB_CHECK_ARG(n_pad[1] == 32 && n_pad[2] == 32 && n_pad[3] == 32 && n_valid[1] == 32, "segments 1-3 must be 32 columns wide");
    if (epi->mode == 1) {
        JB_CHECK_ARG(n_pad[0] == 160 && n_valid[0] == JAMUN_S + JAMUN_V && epi->op_s && epi->op_v, "mode 1: contraction segments");
    } else {
        JB_CHECK_ARG(n_pad[0] == 128 && n_valid[0] == JAMUN_S && out_col && out_col[0] == 0 && out_col[1] == JAMUN_S && epi->x_new &&
                         (!epi->skip_w || epi->x_res) && (!epi->op_s || epi->op_v),
                     "mode 2: block-tail segments");
    }
    JB_CHECK_ARG(epi->op_rows_pad % 8 == 0 && (((size_t)epi->op_s | (size_t)epi->op_v) & 127) == 0, "operand buffers must be 128-byte aligned");
    return gemm_launch(true, nseg, a, b, n_stages, n_pad, n_valid, out_col, alpha, addend, addend_ld, 1, 0, rows, rows_pad, row_scale,
                       nullptr, 0, 1, nullptr, status, addend_scale, a_tile_major, epi, stream);
}

template <bool F16>
static int gemm_launch_t(Params& P, int nseg, const int* n_valid, const int* out_col, int col_blocks, int rows, int rows_pad, float* out,
                         int out_ld, int k_splits, float* partial, jamun_stream_t stream) {
    using Smem = SmemT<F16>;
    const size_t smem = sizeof(Smem) + 1024;
    static_assert(sizeof(Smem) + 1024 <= 227 * 1024, "shared memory budget");
    const bool wide = P.coalesce && col_blocks > 1;
    cudaError_t e = wide ? cudaFuncSetAttribute(gemm_tf32x3_kernel<true, F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                         : cudaFuncSetAttribute(gemm_tf32x3_kernel<false, F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
        jb::set_error("jamun_gemm: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        return JAMUN_ECUDA;
    }
    if (wide) {
        const int tiles = rows_pad / 128;
        int groups = jb::kNumSMs / tiles;  // few tiles: run the column-block passes of a tile on several CTAs
        if (groups > col_blocks) groups = col_blocks;
        if (groups < 1) groups = 1;
        P.cb_groups = groups;
        gemm_tf32x3_kernel<true, F16><<<dim3(tiles, groups), kThreads, smem, jb::as_stream(stream)>>>(P);
    }
    else gemm_tf32x3_kernel<false, F16><<<dim3(rows_pad / 128, k_splits), kThreads, smem, jb::as_stream(stream)>>>(P);
    if (k_splits > 1) {
        int4 c0 = {0, 0, 0, 0}, nc = {0, 0, 0, 0};
        int* pc0 = &c0.x;
        int* pnc = &nc.x;
        long long width = 0;
        for (int s = 0; s < nseg; ++s) pc0[s] = out_col[s], pnc[s] = n_valid[s], width += n_valid[s];
        long long total = (long long)rows * width;
        int blocks = (int)((total + 255) / 256);
        if (blocks > jb::kNumSMs * 8) blocks = jb::kNumSMs * 8;
        gemm_splitk_reduce_kernel<<<blocks, 256, 0, jb::as_stream(stream)>>>(partial, P.partial_stride, k_splits, rows, out_ld, c0, nc,
                                                                            nseg, out);
    }
    JB_CHECK_LAUNCH();
    return JAMUN_OK;
}

static int gemm_launch(bool f16, int nseg, const float* const* a, const float* const* b, const int* n_stages, const int* n_pad,
                       const int* n_valid, const int* out_col, const float* alpha, const float* const* addend,
                       const int* addend_ld, int col_blocks, long long b_block_floats, int rows, int rows_pad,
                       const float* row_scale, float* out, int out_ld, int k_splits, float* partial, int* status,
                       const float* addend_scale, int a_tile_major, const jamun_gemm_epilogue* epi, jamun_stream_t stream) {
    JB_CHECK_ARG(nseg >= 1 && nseg <= 4 && a && b && n_stages && n_pad && n_valid && out_col && alpha && (out || (epi && epi->mode != 0)),
                 "bad argument");
    JB_CHECK_ARG(rows_pad % 128 == 0 && rows <= rows_pad, "rows_pad must be a multiple of 128");
    if (rows == 0) return JAMUN_OK;
    Params P{};
    P.nseg = nseg;
    P.k_splits = k_splits;
    P.cb_groups = 1;
    P.partial = partial;
    P.partial_stride = (long long)rows * out_ld;
    P.status = status;
    P.a_tile_major = a_tile_major;
    if (epi) P.epi = *epi;
    {
        const char* e = getenv("JAMUN_GEMM_COALESCE");
        P.coalesce = e ? atoi(e) : 1;
    }
    P.rows = rows;
    P.rows_pad = rows_pad;
    P.out_ld = out_ld;
    P.row_scale = row_scale;
    P.b_block_floats = b_block_floats;
    P.col_blocks = col_blocks;
    JB_CHECK_ARG(col_blocks >= 1, "col_blocks must be >= 1");
    int dcol = 0;
    for (int s = 0; s < nseg; ++s) {
        JB_CHECK_ARG(n_pad[s] % 16 == 0 && n_pad[s] >= 16 && n_pad[s] <= kMaxN && n_valid[s] <= n_pad[s], "n_pad out of range");
        P.seg[s] = Seg{a[s], b[s], out, addend ? addend[s] : nullptr, n_stages[s], n_pad[s], n_valid[s], dcol, out_col[s],
                       addend_ld ? addend_ld[s] : 0, alpha[s], addend_scale ? addend_scale[s] : 1.0f};
        dcol += n_pad[s];
    }
    JB_CHECK_ARG(dcol <= kACol0, "accumulators exceed 256 TMEM columns");
    return f16 ? gemm_launch_t<true>(P, nseg, n_valid, out_col, col_blocks, rows, rows_pad, out, out_ld, k_splits, partial, stream)
               : gemm_launch_t<false>(P, nseg, n_valid, out_col, col_blocks, rows, rows_pad, out, out_ld, k_splits, partial, stream);
}
