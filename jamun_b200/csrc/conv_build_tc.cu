// Aggregate step of the equivariant convolution on the tensor cores (tcgen05, kind::tf32, 3xTF32 split).
//
// Per receiver node i the aggregate   A_i[k', u'] = sum_{e -> i} h'_e[k'] * f_e[u']      (conv_build.cu, same output)
// is a small GEMM  F_i^T . H_i  with K = in-degree (<= 33 for the capped radius graph), M = 344 edge-feature columns and
// N = 65 radial channels.  conv_build_kernel evaluates it with FFMA2 at ~25 % of the FP32 pipe because every pass over
// the in-edges can keep only 4 channels x 11 columns in registers; here the FP32 pipe only *stages* the operands and the
// product runs as tcgen05.mma with the accumulator in tensor memory:
//
//   producers (8 warps)   one item = (node, chunk of <= 32 in-edges).  Warp g gathers in-edges 4g..4g+3 (lanes over the
//                         feature columns), forms the 1e features (x_v . rhat, x_v/sqrt3, x_v x rhat/sqrt2), splits every
//                         value into tf32 hi + lo and writes 16-byte K-quads into K-major SWIZZLE_128B tiles
//                         (rows = feature columns for F, channels for H; one 128-byte row = 32 in-edges): conflict-free.
//   MMA warp (1 thread)   per item and column group t (scalars | x_v.rhat, x_v | x_v x rhat: M tiles of 128, 128, 96 rows):
//                         D_t[u', k'] (+)= Fhi.Hhi + Fhi.Hlo + Flo.Hhi  for each 8-edge K step (N = 80 >= 65);
//                         commit -> accumulator full, commit -> operand slot free.  Row 64 of H is the constant 1
//                         (the bias channel h' = 1).
//   epilogue (2 x 4 warps) TMEM lane = feature column, so for a fixed channel the 32 lanes of a warp hold the 32
//                         consecutive elements of one 128-byte operand row of jamun_gemm_tf32x3's A layout: every store
//                         instruction writes one full line.
//
// Two operand slots and two sets of three 80-column accumulators keep the roles overlapped; the kernel is persistent
// with one CTA per SM (222 KB of shared memory) and is bound by the HBM write of A (91.5 KB per node).
// The path 0e(x)1e->1e gather (p2) is conv_p2_kernel below.
#include <cuda_fp16.h>
#include <stdlib.h>
#include "common.cuh"
#include "umma.cuh"

namespace {
using namespace jb;

constexpr float kInvSqrt3 = 0.57735026918962576451f;
constexpr float kInvSqrt2 = 0.70710678118654752440f;
constexpr int YLD = 17 * 128;
constexpr int kAccCols = 80;     // accumulator width: 65 channels padded to a legal N
#ifndef JAMUN_BUILD_PRODUCERS
#define JAMUN_BUILD_PRODUCERS 8
#endif
constexpr int kProducerWarps = JAMUN_BUILD_PRODUCERS;  // 8; 11 (= 5 warps per scheduler at 96 registers) measured: 2AA +9 %, 4AA -1 %, protein1000 -5 %
__host__ __device__ constexpr int pmod(int x) { return ((x % kProducerWarps) + kProducerWarps) % kProducerWarps; }
constexpr int kEpiWarps = 8;     // two sets of four (TMEM lane quarter = warp % 4)
constexpr int kMmaWarp = kEpiWarps;
constexpr int kThreads = 32 * (kEpiWarps + 1 + kProducerWarps);
constexpr int kHBytes = kAccCols * 128;  // H tile: channel rows 0..64 (rows 65..79 are never written nor used)

// F16: operands split in fp16 and interleaved along K -- a 16-byte chunk of a tile row = [hi(e0..e3) | lo(e0..e3)] halves, so
// one kind::f16 instruction (K = 16) covers 8 edges and  F.H1 + F.H2  with H1 = [hh | hh], H2 = [hl | 0] gives
// fh.hh + fl.hh + fh.hl: 2 MMAs per 8-edge step instead of 3, one F tile instead of two (half the shared-memory stores), and
// three operand slots instead of two in the same shared memory.  (tf32: F16 = false, the round-1 form.)
template <int S_IN, int V_IN, bool F16 = false>
struct Shape {
    static constexpr int D_IN = S_IN + 3 * V_IN;
    static constexpr int NS = (S_IN + 31) / 32;
    static constexpr int NSL0 = NS + (V_IN > 0 ? 1 : 0), NSL1 = V_IN > 0 ? 2 : 0;
    static constexpr int NCOL = NS + (V_IN > 0 ? 7 : 0);  // 32-column groups of F
    static constexpr int F_BYTES = NCOL * 32 * 128;
    static constexpr int NT = V_IN > 0 ? 3 : 1;           // M tiles (accumulators) per node
    static constexpr int kSlots = F16 ? 3 : 2;  // operand slots
    static constexpr int SLOT_BYTES = 2 * kHBytes + (F16 ? 1 : 2) * F_BYTES;
    // barriers | slots | tail the last M tile's 128-row read may run into
    static constexpr int SMEM_BYTES = 1024 /*alignment*/ + 1024 /*barriers*/ + kSlots * SLOT_BYTES + 8192;
    __host__ __device__ static constexpr int tgroups(int t) { return t == 0 ? NS : t == 1 ? 4 : 3; }  // live 32-column groups
    __host__ __device__ static constexpr int trow0(int t) { return t == 0 ? 0 : t == 1 ? NS * 32 : NS * 32 + 128; }
};

__device__ unsigned long long g_tc_trace[64 * 16];
#define TC_TRACE(slot_, k_)                                                                      \
    if (TRACE && blockIdx.x == 0 && lane == 0 && (slot_) < 64) g_tc_trace[(slot_) * 16 + (k_)] = clock64()

__device__ __forceinline__ float tf32_hi(float v) { return __uint_as_float(__float_as_uint(v) & 0xFFFFE000u); }

__device__ __forceinline__ void store_split(unsigned char* hi_tile, unsigned char* lo_tile, uint32_t off, float a, float b,
                                            float c, float d) {
    const float ah = tf32_hi(a), bh = tf32_hi(b), ch = tf32_hi(c), dh = tf32_hi(d);
    *reinterpret_cast<float4*>(hi_tile + off) = make_float4(ah, bh, ch, dh);
    *reinterpret_cast<float4*>(lo_tile + off) = make_float4(a - ah, b - bh, c - ch, d - dh);
}

__device__ __forceinline__ uint32_t h2u(__half2 h) { return *reinterpret_cast<uint32_t*>(&h); }
// fp16 split of a K-quad: hi = rn16(v), lo = rn16(v - hi) as packed half2 words (hi01, hi23, lo01, lo23)
__device__ __forceinline__ uint4 split_f16(float a, float b, float c, float d) {
    const __half2 h01 = __floats2half2_rn(a, b), h23 = __floats2half2_rn(c, d);
    const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
    return make_uint4(h2u(h01), h2u(h23), h2u(__floats2half2_rn(a - f01.x, b - f01.y)), h2u(__floats2half2_rn(c - f23.x, d - f23.y)));
}
// F operand chunk: [hi | lo];  H operand chunks: H1 = [hi | hi], H2 = [lo | 0]
__device__ __forceinline__ void store_f_f16(unsigned char* tile, uint32_t off, float a, float b, float c, float d) {
    *reinterpret_cast<uint4*>(tile + off) = split_f16(a, b, c, d);
}
__device__ __forceinline__ void store_h_f16(unsigned char* h1, unsigned char* h2, uint32_t off, float a, float b, float c, float d) {
    const uint4 s = split_f16(a, b, c, d);
    *reinterpret_cast<uint4*>(h1 + off) = make_uint4(s.x, s.y, s.x, s.y);
    *reinterpret_cast<uint4*>(h2 + off) = make_uint4(s.z, s.w, 0u, 0u);
}

template <int S_IN, int V_IN, bool TRACE = false, bool F16 = false>
__global__ void __launch_bounds__(kThreads, 1)
conv_build_tc_kernel(const float* __restrict__ x, const int* __restrict__ rowptr, const int* __restrict__ col,
                     const float* __restrict__ h, const float* __restrict__ rhat, int row0, int nrows, int rows_pad,
                     float* __restrict__ a0, float* __restrict__ a1, size_t a1_comp_stride, float* __restrict__ inv_deg,
                     int tile_major, int pf_dist) {
    using SH = Shape<S_IN, V_IN, F16>;
    constexpr int NS = SH::NS, NCOL = SH::NCOL, NT = SH::NT, kSlots = SH::kSlots;
    constexpr int kBufs = 2 * NT;
    extern __shared__ unsigned char dsm_raw[];
    unsigned char* dsm = reinterpret_cast<unsigned char*>(((uintptr_t)dsm_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(dsm);
    uint64_t* full = bars;                  // [kSlots]  producers -> MMA
    uint64_t* empty = bars + kSlots;        // [kSlots]  MMA -> producers
    uint64_t* tfull = bars + 2 * kSlots;    // [kBufs]   MMA -> epilogue
    uint64_t* tempty = tfull + kBufs;       // [kBufs]   epilogue -> MMA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + kBufs);
    unsigned char* slots = dsm + 1024;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < kSlots; ++s) {
            umma::mbar_init(&full[s], kProducerWarps);
            umma::mbar_init(&empty[s], 1);
        }
        for (int b = 0; b < kBufs; ++b) {
            umma::mbar_init(&tfull[b], 1);
            umma::mbar_init(&tempty[b], 4);
        }
        umma::fence_barrier_init();
    }
    if (warp == kMmaWarp) umma::tmem_alloc<512>(tmem_slot);
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    // rows are dealt round-robin: at any moment the CTAs work on ~gridDim.x consecutive rows, so the 128-byte lines they
    // write into each operand stage are neighbours (DRAM page locality of the 1.6 GB output stream)
    const int r_begin = blockIdx.x, r_step = gridDim.x, r_end = nrows;

    if (warp > kMmaWarp) {
        // ------------------------------------------------------------------ producers
        // The (item, K-quad) tasks of consecutive items are dealt round-robin to the 8 warps: an item with 16 in-edges
        // occupies 4 warps while the other 4 already gather the next item.  tb = index (mod 8) of the item's first task;
        // this warp owns K-quad kq = (w - tb) mod 8 of the item if kq < ntasks.
        const int w = warp - kMmaWarp - 1;
        const uint32_t off_lane = (uint32_t)(lane >> 3) * 1024u + (uint32_t)(lane & 7) * 128u;
        int it = 0, tb = 0;
        auto tasks_of = [](int deg) {  // K-quads (incl. zero padding) over all chunks of a node
            int t = 0;
            for (int c = 0; c == 0 || 32 * c < deg; ++c) {
                const int n = min(32, deg - 32 * c);
                t += 2 * (n > 8 ? (n + 7) >> 3 : 1);
            }
            return t;
        };
        // index prefetch: row extents two rows ahead, this warp's four source indices one row ahead -- the gathers of a
        // row then start without the rowptr -> col -> x chain of dependent L2 round trips
        int e0_n = 0, deg_n = 0, e0_nn = 0, deg_nn = 0;
        int jn[4] = {0, 0, 0, 0};
        if (r_begin < r_end) {
            e0_n = rowptr[row0 + r_begin];
            deg_n = rowptr[row0 + r_begin + 1] - e0_n;
#pragma unroll
            for (int q = 0; q < 4; ++q) jn[q] = 4 * w + q < deg_n ? col[e0_n + 4 * w + q] : 0;
        }
        if (r_begin + r_step < r_end) {
            e0_nn = rowptr[row0 + r_begin + r_step];
            deg_nn = rowptr[row0 + r_begin + r_step + 1] - e0_nn;
        }
        // L2 prefetch of the read-once streams: the radial channels h and unit vectors rhat of a node's in-edges are contiguous
        // (receiver-sorted CSR) and are touched exactly once, so every gather of them is a compulsory L2 miss served by an HBM
        // that is saturated with this kernel's own 1.6 GB of writes (measured: 31 % of the read sectors miss, the gather of an
        // item takes ~5.5 k cycles and the epilogue / MMA roles starve behind it).  One lane asks the L2 for the rows of the
        // item `pf_dist` nodes ahead (cp.async.bulk.prefetch.L2: no registers, no shared memory); its extents are loaded one
        // iteration earlier so the request never waits on its own index loads.
        int pf_a = 0, pf_b = 0;
        if (pf_dist > 0 && w == 0 && lane == 0 && r_begin + pf_dist * r_step < r_end) {
            pf_a = rowptr[row0 + r_begin + pf_dist * r_step];
            pf_b = rowptr[row0 + r_begin + pf_dist * r_step + 1];
        }
        for (int r = r_begin; r < r_end; r += r_step) {
            if (pf_dist > 0 && w == 0 && lane == 0) {
                if (pf_b > pf_a) {
                    const uint32_t ne = (uint32_t)(pf_b - pf_a);
                    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(h + (size_t)pf_a * JAMUN_EDGE_HID), "r"(ne * 256u)
                                 : "memory");
                    if constexpr (V_IN > 0)
                        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(rhat + 4 * (size_t)pf_a), "r"(ne * 16u) : "memory");
                }
                const int rn = r + (pf_dist + 1) * r_step;
                pf_a = pf_b = 0;
                if (rn < r_end) {
                    pf_a = rowptr[row0 + rn];
                    pf_b = rowptr[row0 + rn + 1];
                }
            }
            const int e0 = e0_n, deg = deg_n;
            int jc[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) jc[q] = jn[q];
            e0_n = e0_nn, deg_n = deg_nn;
            if (r + r_step < r_end) {
                const int kn = pmod(w - tb - tasks_of(deg));  // this warp's K-quad in the next node's first chunk
#pragma unroll
                for (int q = 0; q < 4; ++q) jn[q] = 4 * kn + q < deg_n ? col[e0_n + 4 * kn + q] : 0;
            }
            if (r + 2 * r_step < r_end) {
                e0_nn = rowptr[row0 + r + 2 * r_step];
                deg_nn = rowptr[row0 + r + 2 * r_step + 1] - e0_nn;
            }
            const int nchunks = deg > 32 ? (deg + 31) >> 5 : 1;
            if (w == 0 && lane == 0 && inv_deg) inv_deg[row0 + r] = 1.0f / (float)(deg > 0 ? deg : 1);
            for (int c = 0; c < nchunks; ++c, ++it) {
                const int n = min(32, deg - 32 * c);
                const int ksteps = n > 8 ? (n + 7) >> 3 : 1;
                const int ntasks = 2 * ksteps;
                const int g = pmod(w - tb);  // K-quad owned by this warp (none if g >= ntasks)
                tb = pmod(tb + ntasks);
                const uint32_t off = off_lane + (uint32_t)((g ^ (lane & 7)) << 4);
                if (g == 0) TC_TRACE(it, 0);
                const int s = it % kSlots;
                const uint32_t empty_parity = (((uint32_t)it / kSlots) & 1u) ^ 1u;  // waited on just before the stores
                unsigned char* slot = slots + s * SH::SLOT_BYTES;
                unsigned char* Hhi = slot;
                unsigned char* Hlo = slot + kHBytes;
                unsigned char* Fhi = slot + 2 * kHBytes;
                unsigned char* Flo = Fhi + SH::F_BYTES;
                if (4 * g < n) {
                    const int nv = min(4, n - 4 * g);
                    const int eb = e0 + 32 * c + 4 * g;
                    // phase 1: every gather of the four in-edges is issued before any use (missing edges re-read edge 0
                    // and are masked below), so the warp pays one L2 round trip per item, not one per edge
                    float xs[NS][4], xv[V_IN > 0 ? 3 : 1][4], hh[2][4];
                    float4 rh[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const bool ok = q < nv;
                        const int e = ok ? eb + q : eb;
                        const int j = (c == 0 && ok) ? jc[q] : col[e];
                        const float* xr = x + (size_t)j * SH::D_IN;
#pragma unroll
                        for (int sl = 0; sl < NS; ++sl) xs[sl][q] = xr[(lane + 32 * sl < S_IN) ? lane + 32 * sl : lane];
                        if constexpr (V_IN > 0) {
                            rh[q] = *reinterpret_cast<const float4*>(rhat + 4 * (size_t)e);
                            xv[0][q] = xr[S_IN + lane];
                            xv[1][q] = xr[S_IN + V_IN + lane];
                            xv[2][q] = xr[S_IN + 2 * V_IN + lane];
                        }
                        const float* he = h + (size_t)e * JAMUN_EDGE_HID;
                        hh[0][q] = he[lane];
                        hh[1][q] = he[32 + lane];
                    }
                    // phase 2: features
                    float f[NCOL][4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float m = q < nv ? 1.f : 0.f;
#pragma unroll
                        for (int sl = 0; sl < NS; ++sl) f[sl][q] = (lane + 32 * sl < S_IN) ? m * xs[sl][q] : 0.f;
                        if constexpr (V_IN > 0) {
                            const float vx = m * xv[0][q], vy = m * xv[1][q], vz = m * xv[2][q];
                            f[NS][q] = vx * rh[q].x + vy * rh[q].y + vz * rh[q].z;
                            f[NS + 1][q] = vx * kInvSqrt3;
                            f[NS + 2][q] = vy * kInvSqrt3;
                            f[NS + 3][q] = vz * kInvSqrt3;
                            f[NS + 4][q] = (vy * rh[q].z - vz * rh[q].y) * kInvSqrt2;
                            f[NS + 5][q] = (vz * rh[q].x - vx * rh[q].z) * kInvSqrt2;
                            f[NS + 6][q] = (vx * rh[q].y - vy * rh[q].x) * kInvSqrt2;
                        }
                        hh[0][q] *= m;
                        hh[1][q] *= m;
                    }
                    if (g == 0) TC_TRACE(it, 1);
                    umma::mbar_wait(&empty[s], empty_parity);  // gathers above overlap the MMAs still reading this slot
                    if (g == 0) TC_TRACE(it, 2);
                    if constexpr (F16) {  // Hhi / Hlo hold H1 = [hi | hi] / H2 = [lo | 0]; Fhi is the one F tile
#pragma unroll
                        for (int sl = 0; sl < NCOL; ++sl) store_f_f16(Fhi, off + sl * 4096, f[sl][0], f[sl][1], f[sl][2], f[sl][3]);
                        store_h_f16(Hhi, Hlo, off, hh[0][0], hh[0][1], hh[0][2], hh[0][3]);
                        store_h_f16(Hhi, Hlo, off + 4096, hh[1][0], hh[1][1], hh[1][2], hh[1][3]);
                        if (lane == 0) {  // row 64: the bias channel h' = 1 (row % 8 == 0: chunk g is not permuted)
                            const uint32_t o64 = 8 * 1024 + (uint32_t)(g << 4);
                            const uint32_t w01 = 0x3C00u | (nv > 1 ? 0x3C000000u : 0u), w23 = (nv > 2 ? 0x3C00u : 0u) | (nv > 3 ? 0x3C000000u : 0u);
                            *reinterpret_cast<uint4*>(Hhi + o64) = make_uint4(w01, w23, w01, w23);
                            *reinterpret_cast<uint4*>(Hlo + o64) = make_uint4(0u, 0u, 0u, 0u);
                        }
                    } else {
#pragma unroll
                    for (int sl = 0; sl < NCOL; ++sl) store_split(Fhi, Flo, off + sl * 4096, f[sl][0], f[sl][1], f[sl][2], f[sl][3]);
                    store_split(Hhi, Hlo, off, hh[0][0], hh[0][1], hh[0][2], hh[0][3]);
                    store_split(Hhi, Hlo, off + 4096, hh[1][0], hh[1][1], hh[1][2], hh[1][3]);
                    if (lane == 0) {  // row 64: the bias channel h' = 1 (row % 8 == 0: chunk g is not permuted)
                        const uint32_t o64 = 8 * 1024 + (uint32_t)(g << 4);
                        *reinterpret_cast<float4*>(Hhi + o64) =
                            make_float4(1.f, nv > 1 ? 1.f : 0.f, nv > 2 ? 1.f : 0.f, nv > 3 ? 1.f : 0.f);
                        *reinterpret_cast<float4*>(Hlo + o64) = make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                    }
                } else if (g < 2 * ksteps) {  // zero K-quad completing the last 8-edge step (or an isolated node)
                    umma::mbar_wait(&empty[s], empty_parity);
                    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int sl = 0; sl < NCOL; ++sl) {
                        *reinterpret_cast<float4*>(Fhi + off + sl * 4096) = z;
                        if constexpr (!F16) *reinterpret_cast<float4*>(Flo + off + sl * 4096) = z;
                    }
                    *reinterpret_cast<float4*>(Hhi + off) = z;
                    *reinterpret_cast<float4*>(Hlo + off) = z;
                    *reinterpret_cast<float4*>(Hhi + off + 4096) = z;
                    *reinterpret_cast<float4*>(Hlo + off + 4096) = z;
                    if (lane == 0) {
                        const uint32_t o64 = 8 * 1024 + (uint32_t)(g << 4);
                        *reinterpret_cast<float4*>(Hhi + o64) = z;
                        *reinterpret_cast<float4*>(Hlo + o64) = z;
                    }
                } else {
                    // no task in this item: still observe the slot's phase and arrive -- every warp takes part in every phase,
                    // which keeps all parity waits within one phase of their barrier
                    umma::mbar_wait(&empty[s], empty_parity);
                }
                umma::fence_proxy_async();  // generic-proxy writes -> visible to the tensor core's async-proxy reads
                __syncwarp();
                if (lane == 0) umma::mbar_arrive(&full[s]);
                if (g == 0) TC_TRACE(it, 3);
            }
        }
    } else if (warp == kMmaWarp) {
        // ------------------------------------------------------------------ MMA issuer
        constexpr uint32_t idesc = F16 ? umma::make_idesc_f16(128, kAccCols) : umma::make_idesc_tf32(128, kAccCols);
        int it = 0;
        uint32_t na = 0;  // node counter of this CTA: accumulator set = na & 1
        for (int r = r_begin; r < r_end; r += r_step, ++na) {
            const int i = row0 + r;
            const int deg = rowptr[i + 1] - rowptr[i];
            const int nchunks = deg > 32 ? (deg + 31) >> 5 : 1;
            const uint32_t set = na & 1u;
            for (int c = 0; c < nchunks; ++c, ++it) {
                const int n = min(32, deg - 32 * c);
                const int ksteps = n > 8 ? (n + 7) >> 3 : 1;
                const int s = it % kSlots;
                TC_TRACE(it, 4);
                umma::mbar_wait(&full[s], ((uint32_t)it / kSlots) & 1u);
                TC_TRACE(it, 5);
                if (c == 0) {
#pragma unroll
                    for (int t = 0; t < NT; ++t) umma::mbar_wait(&tempty[set * NT + t], ((na >> 1) & 1u) ^ 1u);
                }
                TC_TRACE(it, 6);
                umma::fence_after_sync();
                if (umma::elect_one()) {
                    const uint32_t slot = umma::smem_u32(slots + s * SH::SLOT_BYTES);
                    const uint32_t hhi = umma::desc_lo_kmajor_sw128(slot), hlo = umma::desc_lo_kmajor_sw128(slot + kHBytes);
#pragma unroll
                    for (int t = 0; t < NT; ++t) {
                        const uint32_t d = tmem_base + (set * NT + t) * kAccCols;
                        const uint32_t fhi = umma::desc_lo_kmajor_sw128(slot + 2 * kHBytes + SH::trow0(t) * 128);
                        const uint32_t flo = umma::desc_lo_kmajor_sw128(slot + 2 * kHBytes + SH::F_BYTES + SH::trow0(t) * 128);
                        for (int ks = 0; ks < ksteps; ++ks) {
                            const uint32_t k2 = 2u * ks;  // 32 bytes per K step, in 16-byte descriptor units
                            if constexpr (F16) {  // F.[hh | hh] + F.[hl | 0]  (hhi / hlo address the H1 / H2 tiles)
                                umma::mma_f16_ss(d, umma::make_desc(fhi + k2, umma::kDescHiKmajorSw128),
                                                 umma::make_desc(hhi + k2, umma::kDescHiKmajorSw128), idesc, (c > 0 || ks > 0) ? 1u : 0u);
                                umma::mma_f16_ss(d, umma::make_desc(fhi + k2, umma::kDescHiKmajorSw128),
                                                 umma::make_desc(hlo + k2, umma::kDescHiKmajorSw128), idesc, 1u);
                                continue;
                            }
                            umma::mma_tf32_ss(d, umma::make_desc(fhi + k2, umma::kDescHiKmajorSw128),
                                              umma::make_desc(hhi + k2, umma::kDescHiKmajorSw128), idesc, (c > 0 || ks > 0) ? 1u : 0u);
                            umma::mma_tf32_ss(d, umma::make_desc(fhi + k2, umma::kDescHiKmajorSw128),
                                              umma::make_desc(hlo + k2, umma::kDescHiKmajorSw128), idesc, 1u);
                            umma::mma_tf32_ss(d, umma::make_desc(flo + k2, umma::kDescHiKmajorSw128),
                                              umma::make_desc(hhi + k2, umma::kDescHiKmajorSw128), idesc, 1u);
                        }
                        if (c == nchunks - 1) umma::commit(&tfull[set * NT + t]);
                    }
                    umma::commit(&empty[s]);
                }
                __syncwarp();
                TC_TRACE(it, 7);
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue (set 0: warps 0-3, set 1: warps 4-7)
        const uint32_t es = warp >> 2, wq = warp & 3;
        uint32_t na = 0;
        for (int r = r_begin; r < r_end; r += r_step, ++na) {
            if ((na & 1u) != es) continue;
            const int swz = (((lane >> 2) ^ (r & 7)) << 2) | (lane & 3);
#pragma unroll
            for (int t = 0; t < NT; ++t) {
                const uint32_t b = es * NT + t;
                const bool live = (int)wq < SH::tgroups(t);  // this warp's 32 feature columns exist
                // operand row of channel k' for this warp's column group: base + k' * kstride
                // stage-major: [stage][rows_pad][32];  tile-major: [row / 128][stage][128][32] (all stages of a 128-row tile
                // contiguous: the 715 lines a node writes fall into ~12 MB instead of being 2.3 MB apart across 1.6 GB)
                float* dst;
                size_t kstride;
                const bool zero = t == 0 || (t == 1 && wq == 0);          // 0e operand (a0) or a 1e component (a1)
                const int nsl = zero ? SH::NSL0 : SH::NSL1;
                const int slot = t == 0 ? (int)wq : zero ? NS : (t == 1 ? 0 : 1);
                float* base = zero ? a0 : a1 + (size_t)(t == 1 ? (int)wq - 1 : (int)wq) * a1_comp_stride;
                if (tile_major) {
                    dst = base + ((size_t)(r >> 7) * (65 * nsl) + slot) * 4096 + (size_t)(r & 127) * 32;
                    kstride = (size_t)nsl * 4096;
                } else {
                    dst = base + ((size_t)slot * rows_pad + r) * 32;
                    kstride = (size_t)nsl * rows_pad * 32;
                }
                dst += swz;
                if (wq == 0 && t == 0) TC_TRACE(na, 8);
                umma::mbar_wait(&tfull[b], (na >> 1) & 1u);
                if (wq == 0) TC_TRACE(na, 9 + t);
                umma::fence_after_sync();
                const uint32_t taddr = tmem_base + ((32u * wq) << 16) + b * kAccCols;
                if (live) {
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        uint32_t v[32];
                        umma::tmem_ld32(taddr + 32u * half, v);
                        umma::wait_ld();
#pragma unroll
                        for (int k = 0; k < 32; ++k) __stcs(dst + (size_t)(32 * half + k) * kstride, __uint_as_float(v[k]));
                    }
                    const uint32_t vb = umma::tmem_ld1(taddr + 64u);
                    umma::wait_ld();
                    __stcs(dst + (size_t)64 * kstride, __uint_as_float(vb));
                }
                umma::fence_before_sync();
                __syncwarp();
                if (lane == 0) umma::mbar_arrive(&tempty[b]);
                if (wq == 0) TC_TRACE(na, 12 + t);
            }
        }
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == kMmaWarp) umma::tmem_dealloc<512>(tmem_base);
}

// ---- path 0e(x)1e->1e ------------------------------------------------------------------------------------------------------
//   p2_i[c, w] = sum_{e -> i} rhat_e[c] * T_e[w],     T_e[w] = sum_k' h'_e[k'] * Y_j(e)[k', w]
// T is evaluated source-major: one warp per source node j keeps its transformed row Y_j (65 x 32) in registers and walks
// j's out-edges (CSR by source), so Y is read once (158 MB) instead of once per edge (2.4 GB of L2 gathers); the per-edge
// channels h_e are warp-uniform 16-byte loads.  The receiver-side sum is a contiguous pass over T (edges are receiver-major).
constexpr int kP2Chunk = 16;  // out-edges whose radial channels are staged per pass
constexpr int kP2Warps = 4;
__global__ void __launch_bounds__(32 * kP2Warps, 5)
conv_p2_edge_kernel(const int* __restrict__ src_rowptr, const int* __restrict__ src_eid, const float* __restrict__ h,
                    const float* __restrict__ y, int N, float* __restrict__ t_edge) {
    __shared__ __align__(16) float hs[kP2Warps][kP2Chunk][JAMUN_EDGE_HID];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (j >= N) return;
    const int s0 = src_rowptr[j], s1 = src_rowptr[j + 1];
    if (s0 == s1) return;
    int eid = (lane < kP2Chunk && s0 + lane < s1) ? src_eid[s0 + lane] : 0;
    float yr[JAMUN_EDGE_HID + 1];
    const float* yj = y + (size_t)j * YLD + lane;
#pragma unroll
    for (int k = 0; k <= JAMUN_EDGE_HID; ++k) yr[k] = yj[k * JAMUN_V];
    const int half = lane >> 4, l16 = lane & 15;
    for (int c0 = s0; c0 < s1; c0 += kP2Chunk) {
        const int cnt = min(kP2Chunk, s1 - c0);
        // stage the chunk's channel rows with 16-byte async copies (two 256-byte rows per instruction, no registers held):
        // the whole chunk is in flight together, one L2 round trip per chunk
#pragma unroll
        for (int q2 = 0; q2 < kP2Chunk / 2; ++q2) {
            const int q = 2 * q2 + half;
            const int e = __shfl_sync(0xffffffffu, eid, q);
            if (q < cnt) {
                const uint32_t dst = umma::smem_u32(&hs[wib][q][4 * l16]);
                const float* src = h + (size_t)e * JAMUN_EDGE_HID + 4 * l16;
                asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        const int eid_cur = eid;
        if (c0 + kP2Chunk < s1) eid = (lane < kP2Chunk && c0 + kP2Chunk + lane < s1) ? src_eid[c0 + kP2Chunk + lane] : 0;
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
        for (int q = 0; q < cnt; q += 2) {  // two edges per pass: 8 independent FMA chains (row q+1 of a ragged tail is stale, unused)
            const float4* ha = reinterpret_cast<const float4*>(hs[wib][q]);
            const float4* hb = reinterpret_cast<const float4*>(hs[wib][q + 1 < kP2Chunk ? q + 1 : q]);
            float a[4] = {yr[JAMUN_EDGE_HID], 0.f, 0.f, 0.f};  // bias channel h' = 1
            float b[4] = {yr[JAMUN_EDGE_HID], 0.f, 0.f, 0.f};
#pragma unroll
            for (int k4 = 0; k4 < JAMUN_EDGE_HID / 4; ++k4) {
                const float4 wa = ha[k4], wb = hb[k4];
                a[0] = fmaf(wa.x, yr[4 * k4], a[0]);
                b[0] = fmaf(wb.x, yr[4 * k4], b[0]);
                a[1] = fmaf(wa.y, yr[4 * k4 + 1], a[1]);
                b[1] = fmaf(wb.y, yr[4 * k4 + 1], b[1]);
                a[2] = fmaf(wa.z, yr[4 * k4 + 2], a[2]);
                b[2] = fmaf(wb.z, yr[4 * k4 + 2], b[2]);
                a[3] = fmaf(wa.w, yr[4 * k4 + 3], a[3]);
                b[3] = fmaf(wb.w, yr[4 * k4 + 3], b[3]);
            }
            const int ea = __shfl_sync(0xffffffffu, eid_cur, q);
            const int eb = __shfl_sync(0xffffffffu, eid_cur, (q + 1) & (kP2Chunk - 1));
            t_edge[(size_t)ea * JAMUN_V + lane] = (a[0] + a[1]) + (a[2] + a[3]);
            if (q + 1 < cnt) t_edge[(size_t)eb * JAMUN_V + lane] = (b[0] + b[1]) + (b[2] + b[3]);
        }
        __syncwarp();
    }
}

__global__ void __launch_bounds__(256)
conv_p2_reduce_kernel(const int* __restrict__ rowptr, const float* __restrict__ rhat, const float* __restrict__ t_edge, int N,
                      float* __restrict__ p2, int p2_ld, float p2_scale, float* __restrict__ inv_deg) {
    const int lane = threadIdx.x & 31;
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= N) return;
    const int e0 = rowptr[i], e1 = rowptr[i + 1];
    const float invd = 1.0f / (float)(e1 > e0 ? e1 - e0 : 1);
    if (lane == 0 && inv_deg) inv_deg[i] = invd;
    float px = 0.f, py = 0.f, pz = 0.f;
#pragma unroll 4
    for (int e = e0; e < e1; ++e) {
        const float4 rh = *reinterpret_cast<const float4*>(rhat + 4 * (size_t)e);
        const float t = t_edge[(size_t)e * JAMUN_V + lane];
        px = fmaf(rh.x, t, px);
        py = fmaf(rh.y, t, py);
        pz = fmaf(rh.z, t, pz);
    }
    const float sc = p2_scale != 0.f ? p2_scale * invd : 1.0f;
    float* out = p2 + (size_t)i * p2_ld + lane;
    out[0] = px * sc;
    out[JAMUN_V] = py * sc;
    out[2 * JAMUN_V] = pz * sc;
}

template <int S_IN, int V_IN, bool TRACE, bool F16>
int launch_tc_t(const float* x, const int* rowptr, const int* col, const float* h, const float* rhat, int row0, int nrows,
                int rows_pad, float* a0, float* a1, size_t comp, float* inv_deg, int tile_major, cudaStream_t s) {
    using SH = Shape<S_IN, V_IN, F16>;
    static_assert(SH::SMEM_BYTES <= 227 * 1024, "shared memory budget");
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(conv_build_tc_kernel<S_IN, V_IN, TRACE, F16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             SH::SMEM_BYTES);
        if (e != cudaSuccess) {
            jb::set_error("jamun_conv_build_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
            return JAMUN_ECUDA;
        }
        attr_set = true;
    }
    const int blocks = nrows < jb::kNumSMs ? nrows : jb::kNumSMs;
    static int pf_dist = -1;  // nodes of look-ahead of the L2 prefetch (JAMUN_BUILD_PF, 0 = off)
    if (pf_dist < 0) {
        const char* e = getenv("JAMUN_BUILD_PF");
        pf_dist = e ? atoi(e) : 4;
    }
    conv_build_tc_kernel<S_IN, V_IN, TRACE, F16><<<blocks, kThreads, SH::SMEM_BYTES, s>>>(x, rowptr, col, h, rhat, row0, nrows,
                                                                                          rows_pad, a0, a1, comp, inv_deg, tile_major,
                                                                                          pf_dist);
    return JAMUN_OK;
}

template <int S_IN, int V_IN, bool TRACE>
int launch_tc(const float* x, const int* rowptr, const int* col, const float* h, const float* rhat, int row0, int nrows,
              int rows_pad, float* a0, float* a1, size_t comp, float* inv_deg, int tile_major, cudaStream_t s) {
    // operand split of the per-node products: tf32 (default) or fp16-interleaved (JAMUN_B200_BUILD_SPLIT=f16).  The fp16 form issues
    // 2/3 of the MMAs and half the shared-memory stores, but measured no faster (2AA 1.89 vs 1.73 ms per six launches, 4AA 3.72 vs
    // 3.65, protein1000 8.0 vs 8.2): the per-item time is not set by the MMA count
    static int f16 = -1;
    if (f16 < 0) {
        const char* e = getenv("JAMUN_B200_BUILD_SPLIT");
        f16 = (e && e[0] == 'f') ? 1 : 0;
    }
    return f16 ? launch_tc_t<S_IN, V_IN, TRACE, true>(x, rowptr, col, h, rhat, row0, nrows, rows_pad, a0, a1, comp, inv_deg, tile_major, s)
               : launch_tc_t<S_IN, V_IN, TRACE, false>(x, rowptr, col, h, rhat, row0, nrows, rows_pad, a0, a1, comp, inv_deg, tile_major, s);
}

}  // namespace

// a0: [65*nslots0][rows_pad][32]; a1: 3 x [65*2][rows_pad][32] (component stride a1_comp_stride floats; unused when v_in == 0)
static int build_tc(const float* x, int s_in, int v_in, const int* rowptr, const int* col, const float* h, const float* rhat,
                    int row0, int nrows, int rows_pad, float* a0, float* a1, long long a1_comp_stride, float* inv_deg,
                    int tile_major, jamun_stream_t stream) {
    JB_CHECK_ARG(x && rowptr && col && h && rhat && a0, "null argument");
    JB_CHECK_ARG(nrows <= rows_pad && (!tile_major || rows_pad % 128 == 0), "nrows exceeds rows_pad / rows_pad not a multiple of 128");
    JB_CHECK_ARG(((size_t)a0 & 127) == 0 && ((size_t)rhat & 15) == 0, "a0 must be 128-byte, rhat 16-byte aligned");
    if (nrows == 0) return JAMUN_OK;
    cudaStream_t s = jb::as_stream(stream);
    int rc;
    if (s_in == JAMUN_S && v_in == JAMUN_V) {
        JB_CHECK_ARG(a1 && ((size_t)a1 & 127) == 0, "a1 (128-byte aligned) required for vector inputs");
        const char* t = getenv("JAMUN_TC_TRACE");  // debug: per-item role timestamps of CTA 0 (jamun_debug_tc_trace)
        rc = (t && atoi(t)) ? launch_tc<JAMUN_S, JAMUN_V, true>(x, rowptr, col, h, rhat, row0, nrows, rows_pad, a0, a1,
                                                                (size_t)a1_comp_stride, inv_deg, tile_major, s)
                            : launch_tc<JAMUN_S, JAMUN_V, false>(x, rowptr, col, h, rhat, row0, nrows, rows_pad, a0, a1,
                                                                 (size_t)a1_comp_stride, inv_deg, tile_major, s);
    } else if (s_in == JAMUN_S0 && v_in == 0) {
        rc = launch_tc<JAMUN_S0, 0, false>(x, rowptr, col, h, rhat, row0, nrows, rows_pad, a0, a1, 0, inv_deg, tile_major, s);
    } else {
        jb::set_error("jamun_conv_build_tc: unsupported input irreps %dx0e+%dx1e", s_in, v_in);
        return JAMUN_EINVAL;
    }
    if (rc != JAMUN_OK) return rc;
    JB_CHECK_LAUNCH();
    return JAMUN_OK;
}

extern "C" int jamun_conv_build_tc(const float* x, int s_in, int v_in, const int* rowptr, const int* col, const float* h,
                                   const float* rhat, int row0, int nrows, int rows_pad, float* a0, float* a1,
                                   long long a1_comp_stride, float* inv_deg, jamun_stream_t stream) {
    return build_tc(x, s_in, v_in, rowptr, col, h, rhat, row0, nrows, rows_pad, a0, a1, a1_comp_stride, inv_deg, 0, stream);
}

// The same aggregate in the tile-major operand layout [row / 128][stage][128][32] (jamun_gemm_f16x3 with a_tile_major = 1).
extern "C" int jamun_conv_build_tc_tiled(const float* x, int s_in, int v_in, const int* rowptr, const int* col, const float* h,
                                         const float* rhat, int row0, int nrows, int rows_pad, float* a0, float* a1,
                                         long long a1_comp_stride, float* inv_deg, jamun_stream_t stream) {
    return build_tc(x, s_in, v_in, rowptr, col, h, rhat, row0, nrows, rows_pad, a0, a1, a1_comp_stride, inv_deg, 1, stream);
}

extern "C" int jamun_conv_p2(const int* rowptr, const int* src_rowptr, const int* src_eid, const float* h, const float* rhat,
                             const float* y, int N, float* t_edge, float* p2, int p2_ld, float p2_scale, float* inv_deg,
                             jamun_stream_t stream) {
    JB_CHECK_ARG(rowptr && src_rowptr && src_eid && h && rhat && y && t_edge, "null argument");
    if (N == 0) return JAMUN_OK;
    cudaStream_t s = jb::as_stream(stream);
    const int blocks = (int)(((size_t)N * 32 + 255) / 256);
    conv_p2_edge_kernel<<<(N + kP2Warps - 1) / kP2Warps, 32 * kP2Warps, 0, s>>>(src_rowptr, src_eid, h, y, N, t_edge);
    if (p2) conv_p2_reduce_kernel<<<blocks, 256, 0, s>>>(rowptr, rhat, t_edge, N, p2, p2_ld, p2_scale, inv_deg);
    JB_CHECK_LAUNCH();
    return JAMUN_OK;
}

// debug: per-item role timestamps of CTA 0 (JAMUN_TC_TRACE=1)
extern "C" int jamun_debug_tc_trace(unsigned long long* out) {
    return cudaMemcpyFromSymbol(out, g_tc_trace, sizeof(g_tc_trace)) == cudaSuccess ? 0 : 1;
}
