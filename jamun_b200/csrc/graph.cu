// Graph-side kernels: per-chain mean-centre, capped radius graph -> receiver-sorted CSR, edge geometry,
// radial-MLP hidden layer.  All HBM/latency-bound integer/elementwise work (kernels K0-K2).
#include <stdarg.h>
#include <string.h>
#include "common.cuh"
#include "umma.cuh"

namespace jb {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace jb

extern "C" int jamun_abi_version(void) { return 1; }
extern "C" const char* jamun_last_error(void) { return jb::g_err; }

namespace {
using namespace jb;

// ---- K0: mean-centre + scale, one warp per chain -------------------------------------------------
__global__ void center_scale_kernel(const float* __restrict__ y, const int* __restrict__ chain_ptr, int G,
                                    int center, float c_in, float* __restrict__ ybar, float* __restrict__ p) {
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= G) return;
    int lo = chain_ptr[warp], hi = chain_ptr[warp + 1];
    float sx = 0.f, sy = 0.f, sz = 0.f;
    for (int i = lo + lane; i < hi; i += 32) {
        sx += y[3 * i];
        sy += y[3 * i + 1];
        sz += y[3 * i + 2];
    }
    sx = warp_sum(sx);
    sy = warp_sum(sy);
    sz = warp_sum(sz);
    float inv = 1.0f / fmaxf(1.0f, (float)(hi - lo));
    float mx = center ? sx * inv : 0.f, my = center ? sy * inv : 0.f, mz = center ? sz * inv : 0.f;
    for (int i = lo + lane; i < hi; i += 32) {
        float bx = y[3 * i] - mx, by = y[3 * i + 1] - my, bz = y[3 * i + 2] - mz;
        if (ybar) {
            ybar[3 * i] = bx;
            ybar[3 * i + 1] = by;
            ybar[3 * i + 2] = bz;
        }
        if (p) {
            p[3 * i] = bx * c_in;
            p[3 * i + 1] = by * c_in;
            p[3 * i + 2] = bz * c_in;
        }
    }
}

// ---- K1: radius graph.  One thread per receiver scans its chain in ascending index. ------------------
// Distance in unfused fp32 (x,y,z order) so that edge sets are bit-reproducible against the oracle.
__device__ __forceinline__ bool in_range(const float* __restrict__ pos, float xi, float yi, float zi, int j, float r2) {
    float dx = __fsub_rn(xi, pos[3 * j]), dy = __fsub_rn(yi, pos[3 * j + 1]), dz = __fsub_rn(zi, pos[3 * j + 2]);
    float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
    return d2 < r2;
}

template <bool FILL>
__global__ void radius_kernel(const float* __restrict__ pos, const int* __restrict__ chain_of,
                              const int* __restrict__ chain_ptr, int N, float r2, int max_hits,
                              const int* __restrict__ bond_rowptr, const int* __restrict__ bond_src,
                              int* __restrict__ count, const int* __restrict__ rowptr, int* __restrict__ col,
                              int* __restrict__ edst, unsigned char* __restrict__ ebond) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    int c = chain_of[i];
    int lo = chain_ptr[c], hi = chain_ptr[c + 1];
    float xi = pos[3 * i], yi = pos[3 * i + 1], zi = pos[3 * i + 2];
    int hits = 0, n = 0;
    int w = FILL ? rowptr[i] : 0;
    for (int j = lo; j < hi; ++j) {
        if (in_range(pos, xi, yi, zi, j, r2)) {
            if (j != i) {
                if (FILL) {
                    col[w + n] = j;
                    edst[w + n] = i;
                    ebond[w + n] = 0;
                }
                ++n;
            }
            ++hits;
            if (max_hits > 0 && hits >= max_hits) break;
        }
    }
    int b0 = bond_rowptr ? bond_rowptr[i] : 0, b1 = bond_rowptr ? bond_rowptr[i + 1] : 0;
    if (FILL) {
        for (int b = b0; b < b1; ++b) {
            col[w + n] = bond_src[b];
            edst[w + n] = i;
            ebond[w + n] = 1;
            ++n;
        }
    } else {
        count[i] = n + (b1 - b0);
    }
}

// ---- K1 (long chains): cell-list neighbour search -------------------------------------------------------------------------------
// One CTA per chain.  The chain's atoms are binned into cubic cells of edge >= r_cut held in shared memory (counting sort by
// cell: count -> exclusive scan -> fill); a receiver then tests only the atoms of the 27 cells around its own with the same
// unfused-fp32 predicate as the brute-force kernel, so the hit set is identical.  The reference's cap rule ("the first
// max_hits hits scanning the chain in ascending index, self included, then self removed" -- torch_cluster's radius kernel,
// SURVEY A.3) is restated as "the max_hits smallest hit indices": every thread keeps its hits in a sorted insertion buffer.
// Neighbour lists go to a fixed-stride scratch (nbr [N][max_hits]) and are compacted into the CSR after the row scan, so the
// search runs once.  Atomics are used on integer counters only and the per-receiver lists are sorted: the CSR is deterministic.
constexpr int kCellMaxHits = 64;   // max_num_neighbors + 1 <= 64
constexpr int kCellMaxCells = 4096;
constexpr int kCellMaxAtoms = 8192;
constexpr int kHitBuf = 160;       // per-warp hit buffer (>= max_hits + 32, multiple of 32)

__global__ void __launch_bounds__(256)
radius_cell_kernel(const float* __restrict__ pos, const int* __restrict__ chain_ptr, float r2, float r_cut, int max_hits,
                   const int* __restrict__ bond_rowptr, int* __restrict__ count, int* __restrict__ nbr) {
    extern __shared__ __align__(16) unsigned char cell_smem[];
    const int c = blockIdx.x;
    const int lo = chain_ptr[c], n = chain_ptr[c + 1] - lo;
    if (n <= 0) return;
    float* px = reinterpret_cast<float*>(cell_smem);  // [3][n] coordinates (SoA)
    float* py = px + n;
    float* pz = py + n;
    int* cell_of = reinterpret_cast<int*>(pz + n);    // [n]
    int* sorted = cell_of + n;                        // [n] atom indices (chain-local) grouped by cell
    int* cstart = sorted + n;                         // [ncell + 1]
    __shared__ float red[6][8];
    __shared__ float bb[6];
    __shared__ int dims[3];
    __shared__ float cell_edge;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float mn[3] = {3.4e38f, 3.4e38f, 3.4e38f}, mx[3] = {-3.4e38f, -3.4e38f, -3.4e38f};
    for (int a = tid; a < n; a += 256) {
        const float x = pos[3 * (size_t)(lo + a)], y = pos[3 * (size_t)(lo + a) + 1], z = pos[3 * (size_t)(lo + a) + 2];
        px[a] = x, py[a] = y, pz[a] = z;
        mn[0] = fminf(mn[0], x), mn[1] = fminf(mn[1], y), mn[2] = fminf(mn[2], z);
        mx[0] = fmaxf(mx[0], x), mx[1] = fmaxf(mx[1], y), mx[2] = fmaxf(mx[2], z);
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn[k] = fminf(mn[k], __shfl_xor_sync(0xffffffffu, mn[k], o));
            mx[k] = fmaxf(mx[k], __shfl_xor_sync(0xffffffffu, mx[k], o));
        }
        if (lane == 0) red[k][warp] = mn[k], red[3 + k][warp] = mx[k];
    }
    __syncthreads();
    if (tid == 0) {
        float ext = 0.f;
        for (int k = 0; k < 3; ++k) {
            float a = red[k][0], b = red[3 + k][0];
            for (int w = 1; w < 8; ++w) a = fminf(a, red[k][w]), b = fmaxf(b, red[3 + k][w]);
            bb[k] = a, bb[3 + k] = b;
            ext = fmaxf(ext, b - a);
        }
        // cell edge: r_cut (slightly enlarged so that rounding of the cell index can never hide an in-range pair), grown
        // until the grid fits the shared-memory budget
        float edge = r_cut * 1.0001f + 1e-6f;
        for (;;) {
            int tot = 1;
            for (int k = 0; k < 3; ++k) {
                dims[k] = (int)floorf((bb[3 + k] - bb[k]) / edge) + 1;
                tot *= dims[k];
            }
            if (tot <= kCellMaxCells) break;
            edge *= 1.26f;
        }
        cell_edge = edge;
    }
    __syncthreads();
    const int nx = dims[0], ny = dims[1], nz = dims[2], ncell = nx * ny * nz;
    const float inv_edge = 1.0f / cell_edge;
    for (int q = tid; q <= ncell; q += 256) cstart[q] = 0;
    __syncthreads();
    for (int a = tid; a < n; a += 256) {
        const int cx = min(nx - 1, (int)((px[a] - bb[0]) * inv_edge)), cy = min(ny - 1, (int)((py[a] - bb[1]) * inv_edge)),
                  cz = min(nz - 1, (int)((pz[a] - bb[2]) * inv_edge));
        const int cid = (cz * ny + cy) * nx + cx;
        cell_of[a] = cid;
        atomicAdd(&cstart[cid + 1], 1);  // integer histogram: order-independent
    }
    __syncthreads();
    if (warp == 0) {  // inclusive scan of the histogram by one warp (ncell <= 4096)
        int carry = 0;
        for (int base = 1; base <= ncell; base += 32) {
            const int idx = base + lane;
            int v = idx <= ncell ? cstart[idx] : 0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, v, o);
                if (lane >= o) v += t;
            }
            if (idx <= ncell) cstart[idx] = v + carry;
            carry += __shfl_sync(0xffffffffu, v, 31);
        }
    }
    __syncthreads();
    // fill: cursor = cstart copy held in `count` scratch?  no global traffic: reuse cell_of as (cell id) and a second histogram
    // pass with atomic cursors kept in the upper half of cstart's allocation
    int* cursor = cstart + ncell + 1;
    for (int q = tid; q < ncell; q += 256) cursor[q] = cstart[q];
    __syncthreads();
    for (int a = tid; a < n; a += 256) sorted[atomicAdd(&cursor[cell_of[a]], 1)] = a;
    __syncthreads();
    // search: one warp per receiver.  Lanes test 32 candidates at a time and append the hits to a per-warp buffer; the hits are
    // then ranked by counting (rank = number of smaller hit indices), which orders them without a sort and makes "keep the
    // max_hits smallest" a comparison of the rank.  A nearly full buffer is compacted the same way and the scan continues.
    int* hitbuf = cursor + ncell + (size_t)warp * kHitBuf;  // [8 warps][kHitBuf]
    for (int a = warp; a < n; a += 8) {
        const float xi = px[a], yi = py[a], zi = pz[a];
        const int cid = cell_of[a];
        const int cx = cid % nx, cy = (cid / nx) % ny, cz = cid / (nx * ny);
        int cnt = 0;
        auto compact = [&](bool final_pass) {
            // rank every buffered hit; keep those of rank < max_hits at position rank (ascending)
            __syncwarp();
            int keep_val[kHitBuf / 32], keep_rank[kHitBuf / 32];
#pragma unroll
            for (int t = 0; t < kHitBuf / 32; ++t) {
                const int idx = lane + 32 * t;
                keep_val[t] = idx < cnt ? hitbuf[idx] : 0x7fffffff;
                int rk = 0;
                if (idx < cnt)
                    for (int q = 0; q < cnt; ++q) rk += hitbuf[q] < keep_val[t];
                keep_rank[t] = idx < cnt ? rk : 0x7fffffff;
            }
            __syncwarp();
#pragma unroll
            for (int t = 0; t < kHitBuf / 32; ++t)
                if (keep_rank[t] < max_hits) hitbuf[keep_rank[t]] = keep_val[t];
            cnt = min(cnt, max_hits);
            __syncwarp();
            (void)final_pass;
        };
        for (int dz = -1; dz <= 1; ++dz) {
            const int z = cz + dz;
            if (z < 0 || z >= nz) continue;
            for (int dy = -1; dy <= 1; ++dy) {
                const int y = cy + dy;
                if (y < 0 || y >= ny) continue;
                const int x0 = max(cx - 1, 0), x1 = min(cx + 1, nx - 1);
                const int row = (z * ny + y) * nx;
                const int q1 = cstart[row + x1 + 1];
                for (int q0 = cstart[row + x0]; q0 < q1; q0 += 32) {  // the x-neighbours are contiguous in `sorted`
                    const int q = q0 + lane;
                    bool hit = false;
                    int j = 0;
                    if (q < q1) {
                        j = sorted[q];
                        const float ddx = __fsub_rn(xi, px[j]), ddy = __fsub_rn(yi, py[j]), ddz = __fsub_rn(zi, pz[j]);
                        const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(ddx, ddx), __fmul_rn(ddy, ddy)), __fmul_rn(ddz, ddz));
                        hit = d2 < r2;
                    }
                    const unsigned m = __ballot_sync(0xffffffffu, hit);
                    if (cnt + 32 > kHitBuf) compact(false);
                    if (hit) hitbuf[cnt + __popc(m & ((1u << lane) - 1))] = j;
                    cnt += __popc(m);
                }
            }
        }
        compact(true);
        // hitbuf[0 .. cnt) ascending; drop self, write the row
        const int i = lo + a;
        int self_pos = cnt;
        for (int t = lane; t < cnt; t += 32)
            if (hitbuf[t] == a) self_pos = t;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) self_pos = min(self_pos, __shfl_xor_sync(0xffffffffu, self_pos, o));
        for (int t = lane; t < cnt; t += 32)
            if (t != self_pos) nbr[(size_t)i * max_hits + (t < self_pos ? t : t - 1)] = lo + hitbuf[t];
        if (lane == 0) count[i] = cnt - (self_pos < cnt ? 1 : 0) + (bond_rowptr ? bond_rowptr[i + 1] - bond_rowptr[i] : 0);
        __syncwarp();
    }
}

// rows of the fixed-stride neighbour scratch -> CSR (radial sources ascending, then the bonded in-edges in their original order)
__global__ void radius_compact_kernel(const int* __restrict__ nbr, int max_hits, const int* __restrict__ bond_rowptr,
                                      const int* __restrict__ bond_src, const int* __restrict__ rowptr, int N, int* __restrict__ col,
                                      int* __restrict__ edst, unsigned char* __restrict__ ebond) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const int w = rowptr[i];
    const int b0 = bond_rowptr ? bond_rowptr[i] : 0, b1 = bond_rowptr ? bond_rowptr[i + 1] : 0;
    const int m = rowptr[i + 1] - w - (b1 - b0);
    for (int k = 0; k < m; ++k) {
        col[w + k] = nbr[(size_t)i * max_hits + k];
        edst[w + k] = i;
        ebond[w + k] = 0;
    }
    for (int b = b0; b < b1; ++b) {
        col[w + m + b - b0] = bond_src[b];
        edst[w + m + b - b0] = i;
        ebond[w + m + b - b0] = 1;
    }
}

// single-CTA exclusive scan; out[n] = total.  n <= a few 1e5, off the critical path.
__global__ void exclusive_scan_kernel(const int* __restrict__ in, int* __restrict__ out, int n) {
    __shared__ int warp_tot[32];
    __shared__ int warp_excl[32];
    __shared__ int block_total;
    __shared__ int carry_s;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int nwarps = blockDim.x >> 5;
    if (tid == 0) carry_s = 0;
    __syncthreads();
    constexpr int PER = 4;
    for (int base = 0; base < n; base += blockDim.x * PER) {
        int idx = base + tid * PER;
        int v[PER], s = 0;
#pragma unroll
        for (int k = 0; k < PER; ++k) {
            v[k] = (idx + k < n) ? in[idx + k] : 0;
            s += v[k];
        }
        int incl = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) warp_tot[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            int t = (lane < nwarps) ? warp_tot[lane] : 0;
            int ti = t;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int q = __shfl_up_sync(0xffffffffu, ti, o);
                if (lane >= o) ti += q;
            }
            warp_excl[lane] = ti - t;
            if (lane == 31) block_total = ti;
        }
        __syncthreads();
        int excl = carry_s + warp_excl[wid] + incl - s;
#pragma unroll
        for (int k = 0; k < PER; ++k) {
            if (idx + k < n) out[idx + k] = excl;
            excl += v[k];
        }
        __syncthreads();
        if (tid == 0) carry_s += block_total;
        __syncthreads();
    }
    if (tid == 0) out[n] = carry_s;
}

// ---- K2a: edge geometry.  A warp takes 32 consecutive edges: lane = edge for the geometry (all gathers of the 32 edges in
// flight together, one 16-byte rhat store per lane), then lane = radial basis index for the 32 x 32 Gaussians (one 128-byte
// row per edge). ----------------
__global__ void edge_geom_kernel(const float* __restrict__ p, const int* __restrict__ rowptr,
                                 const int* __restrict__ col, const int* __restrict__ edst, int N,
                                 const float* __restrict__ mu, float step, float* __restrict__ rhat,
                                 float* __restrict__ rb) {
    const int E = rowptr[N];
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    const float mu_k = mu[lane];
    const float inv_step = 1.0f / step;
    for (int e0 = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 32; e0 < E; e0 += warps * 32) {
        const int e = e0 + lane;
        float d = 0.f;
        if (e < E) {
            const int j = col[e], i = edst[e];
            const float dx = p[3 * j] - p[3 * i], dy = p[3 * j + 1] - p[3 * i + 1], dz = p[3 * j + 2] - p[3 * i + 2];
            d = sqrtf(dx * dx + dy * dy + dz * dz);
            const float inv = 1.0f / fmaxf(d, 1e-12f);
            *reinterpret_cast<float4*>(rhat + 4 * (size_t)e) = make_float4(dx * inv, dy * inv, dz * inv, d);
        }
        const int cnt = min(32, E - e0);
        for (int q = 0; q < cnt; ++q) {
            const float t = (__shfl_sync(0xffffffffu, d, q) - mu_k) * inv_step;
            rb[(size_t)(e0 + q) * JAMUN_NBASIS + lane] = __expf(-(t * t)) * (1.0f / 1.12f);
        }
    }
}

// ---- CSR by source: for every node j the ids (positions in the receiver-major edge list) of its out-edges -----------
__global__ void count_sources_kernel(const int* __restrict__ rowptr, const int* __restrict__ col, int N, int* __restrict__ counts) {
    const int E = rowptr[N];
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < E; e += gridDim.x * blockDim.x) atomicAdd(&counts[col[e]], 1);
}
__global__ void fill_sources_kernel(const int* __restrict__ rowptr, const int* __restrict__ col, int N,
                                    const int* __restrict__ src_rowptr, int* __restrict__ cursor, int* __restrict__ src_eid) {
    const int E = rowptr[N];
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < E; e += gridDim.x * blockDim.x) {
        const int j = col[e];
        src_eid[src_rowptr[j] + atomicAdd(&cursor[j], 1)] = e;
    }
}

// The atomic cursor above leaves each source's out-edge list in arrival order; an insertion sort of every (short) list makes
// src_eid ascending per source, i.e. bit-reproducible, so source-major reductions (backward dx) are deterministic.
__global__ void sort_sources_kernel(const int* __restrict__ src_rowptr, int N, int* __restrict__ src_eid) {
    // one warp per source: out-degree <= 64 is ranked by counting in registers (two entries per lane); longer lists fall
    // back to an insertion sort by lane 0
    const int lane = threadIdx.x & 31;
    const int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (j >= N) return;
    const int s0 = src_rowptr[j], s1 = src_rowptr[j + 1];
    const int n = s1 - s0;
    if (n <= 1) return;
    if (n <= 64) {
        const int v0 = lane < n ? src_eid[s0 + lane] : 0x7fffffff, v1 = lane + 32 < n ? src_eid[s0 + 32 + lane] : 0x7fffffff;
        int r0 = 0, r1 = 0;
        for (int q = 0; q < 32; ++q) {
            const int a = __shfl_sync(0xffffffffu, v0, q), b = __shfl_sync(0xffffffffu, v1, q);
            r0 += (a < v0) + (b < v0);
            r1 += (a < v1) + (b < v1);
        }
        __syncwarp();
        if (lane < n) src_eid[s0 + r0] = v0;
        if (lane + 32 < n) src_eid[s0 + r1] = v1;
        return;
    }
    if (lane == 0)
        for (int a = s0 + 1; a < s1; ++a) {
            const int v = src_eid[a];
            int b = a - 1;
            while (b >= s0 && src_eid[b] > v) {
                src_eid[b + 1] = src_eid[b];
                --b;
            }
            src_eid[b + 1] = v;
        }
}

// ---- K2b: radial MLP hidden layer: h[e][o] = SiLU(sum_k w0rt[k][o] rb[e][k] + b0eff[flag][o]) --------
// A [E x 32] . [32 x 64] product on the FP32 pipe.  One CTA of 128 threads owns a tile of 128 edges: the tile's radial
// basis (16 KB, contiguous) and the transposed weight (8 KB) arrive by two bulk async copies; each thread keeps an
// 8-edge x 8-channel register tile (packed FFMA2), so every operand fetched from shared memory feeds 8 FMAs.  Lane layout
// (lane = channel group + 8 * edge group) makes each quarter-warp read one 128-byte row (weights) or one address
// (basis, broadcast), and lets 8 lanes store one contiguous 128-byte half row of h.
constexpr int kRhTile = 128;
__global__ void __launch_bounds__(128, 4) edge_radial_hidden_kernel(const float* __restrict__ rb,
                                                                    const unsigned char* __restrict__ ebond,
                                                                    const int* __restrict__ rowptr, int N,
                                                                    const float* __restrict__ w0rt,
                                                                    const float* __restrict__ b0eff,
                                                                    float* __restrict__ h, int L, size_t h_stride) {
    // L layers share the edge tile: layer l uses w0rt + l*2048, b0eff + l*128 and writes h + l*h_stride; the weight images
    // are double-buffered (layer l+2 streams in while layer l+1 computes)
    constexpr uint32_t kWBytes = JAMUN_NBASIS * JAMUN_EDGE_HID * 4u;
    __shared__ __align__(128) float s_rb[kRhTile * JAMUN_NBASIS];
    __shared__ __align__(128) float s_w2[2][JAMUN_NBASIS * JAMUN_EDGE_HID];
    __shared__ float s_b2[2][2 * JAMUN_EDGE_HID];
    __shared__ __align__(8) uint64_t bar[2];
    const int E = rowptr[N];
    const int e0 = blockIdx.x * kRhTile;
    if (e0 >= E) return;
    const int nvalid = min(kRhTile, E - e0);
    const int tid = threadIdx.x;
    if (tid == 0) {
        umma::mbar_init(&bar[0], 1);
        umma::mbar_init(&bar[1], 1);
        umma::fence_barrier_init();
    }
    __syncthreads();
    if (tid == 0) {
        const uint32_t rb_bytes = (uint32_t)nvalid * JAMUN_NBASIS * 4u;
        umma::mbar_arrive_expect_tx(&bar[0], rb_bytes + kWBytes);
        umma::bulk_g2s(s_rb, rb + (size_t)e0 * JAMUN_NBASIS, rb_bytes, &bar[0]);
        umma::bulk_g2s(s_w2[0], w0rt, kWBytes, &bar[0]);
        if (L > 1) {
            umma::mbar_arrive_expect_tx(&bar[1], kWBytes);
            umma::bulk_g2s(s_w2[1], w0rt + JAMUN_NBASIS * JAMUN_EDGE_HID, kWBytes, &bar[1]);
        }
    }
    const int og = tid & 7, eg = tid >> 3;  // channels {4 og .. +3} and {32 + 4 og .. +3};  edges 8 eg .. 8 eg + 7
    for (int l = 0; l < L; ++l) {
    const float* s_w = s_w2[l & 1];
    float* s_b = s_b2[l & 1];
    s_b[tid] = b0eff[(size_t)l * 2 * JAMUN_EDGE_HID + tid];
    umma::mbar_wait(&bar[l & 1], (uint32_t)(l >> 1) & 1u);
    __syncthreads();
    float* hl = h + (size_t)l * h_stride;
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    const float4* rb4 = reinterpret_cast<const float4*>(s_rb) + eg * 8 * (JAMUN_NBASIS / 4);
    const float4* w4 = reinterpret_cast<const float4*>(s_w) + og;
#pragma unroll 2
    for (int k4 = 0; k4 < JAMUN_NBASIS / 4; ++k4) {
        float4 r[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) r[i] = rb4[i * (JAMUN_NBASIS / 4) + k4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const float4 wa = w4[(k4 * 4 + c) * (JAMUN_EDGE_HID / 4)];
            const float4 wb = w4[(k4 * 4 + c) * (JAMUN_EDGE_HID / 4) + 8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float rv = c == 0 ? r[i].x : c == 1 ? r[i].y : c == 2 ? r[i].z : r[i].w;
                jb::ffma2(acc[i][0], acc[i][1], rv, wa.x, wa.y);
                jb::ffma2(acc[i][2], acc[i][3], rv, wa.z, wa.w);
                jb::ffma2(acc[i][4], acc[i][5], rv, wb.x, wb.y);
                jb::ffma2(acc[i][6], acc[i][7], rv, wb.z, wb.w);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int el = eg * 8 + i;
        if (el < nvalid) {
            const int e = e0 + el;
            const float* bias = s_b + (ebond[e] ? JAMUN_EDGE_HID : 0) + 4 * og;
            float4 lo, hi;
            lo.x = jb::siluf_fast(acc[i][0] + bias[0]);
            lo.y = jb::siluf_fast(acc[i][1] + bias[1]);
            lo.z = jb::siluf_fast(acc[i][2] + bias[2]);
            lo.w = jb::siluf_fast(acc[i][3] + bias[3]);
            hi.x = jb::siluf_fast(acc[i][4] + bias[32]);
            hi.y = jb::siluf_fast(acc[i][5] + bias[33]);
            hi.z = jb::siluf_fast(acc[i][6] + bias[34]);
            hi.w = jb::siluf_fast(acc[i][7] + bias[35]);
            float4* dst = reinterpret_cast<float4*>(hl + (size_t)e * JAMUN_EDGE_HID) + og;
            dst[0] = lo;
            dst[8] = hi;
        }
    }
    __syncthreads();  // this layer's weight image and bias rows are free
    if (tid == 0 && l + 2 < L) {
        umma::mbar_arrive_expect_tx(&bar[l & 1], kWBytes);
        umma::bulk_g2s(s_w2[l & 1], w0rt + (size_t)(l + 2) * JAMUN_NBASIS * JAMUN_EDGE_HID, kWBytes, &bar[l & 1]);
    }
    }  // layers
}

// ---- layout conversion ---------------------------------------------------------------------------------
template <bool TO_SOA>
__global__ void layout_kernel(const float* __restrict__ in, int s, int v, int N, float* __restrict__ out) {
    const int D = s + 3 * v;
    size_t total = (size_t)N * D;
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        int n = (int)(t / D), k = (int)(t % D);
        int k2 = k;
        if (k >= s) {
            int q = k - s;
            if (TO_SOA) {  // t indexes the SoA output: q = c*v + u  <- e3nn index u*3 + c
                int c = q / v, u = q % v;
                k2 = s + u * 3 + c;
            } else {  // t indexes the e3nn output: q = u*3 + c <- SoA c*v + u
                int u = q / 3, c = q % 3;
                k2 = s + c * v + u;
            }
        }
        out[t] = in[(size_t)n * D + k2];
    }
}

}  // namespace

extern "C" int jamun_center_scale(const float* y, const int* chain_ptr, int G, int center, float c_in, float* ybar, float* p,
                                  jamun_stream_t stream) {
    JB_CHECK_ARG(y && chain_ptr && G >= 0, "null argument");
    if (G == 0) return JAMUN_OK;
    int blocks = (G * 32 + 255) / 256;
    center_scale_kernel<<<blocks, 256, 0, jb::as_stream(stream)>>>(y, chain_ptr, G, center, c_in, ybar, p);
    JB_CHECK_LAUNCH();
    return JAMUN_OK;
}

extern "C" int jamun_radius_csr(const float* pos, const int* chain_of, const int* chain_ptr, int N, float r2,
                                int max_num_neighbors, const int* bond_rowptr, const int* bond_src, int* scratch,
                                int* rowptr, int* col, int* edst, unsigned char* ebond, jamun_stream_t stream) {
    JB_CHECK_ARG(pos && chain_of && chain_ptr && scratch && rowptr && col && edst && ebond, "null argument");
    JB_CHECK_ARG(N >= 0, "negative N");
    cudaStream_t s = jb::as_stream(stream);
    int max_hits = max_num_neighbors < 0 ? 0 : max_num_neighbors + 1;
    int blocks = (N + 127) / 128;
    if (N > 0) {
        radius_kernel<false><<<blocks, 128, 0, s>>>(pos, chain_of, chain_ptr, N, r2, max_hits, bond_rowptr, bond_src,
                                                    scratch, nullptr, nullptr, nullptr, nullptr);
        JB_CHECK_LAUNCH();
    }
    exclusive_scan_kernel<<<1, 1024, 0, s>>>(scratch, rowptr, N);
    JB_CHECK_LAUNCH();
    if (N > 0) {
        radius_kernel<true><<<blocks, 128, 0, s>>>(pos, chain_of, chain_ptr, N, r2, max_hits, bond_rowptr, bond_src,
                                                   nullptr, rowptr, col, edst, ebond);
        JB_CHECK_LAUNCH();
    }
    return JAMUN_OK;
}

// Cell-list form of jamun_radius_csr for long chains (same CSR, bit for bit).  nbr: [N, max_num_neighbors + 1] int scratch;
// max_chain: longest chain of the batch (host-side constant of the topology).
extern "C" int jamun_radius_csr_cells(const float* pos, const int* chain_ptr, int G, int N, int max_chain, float r2, float r_cut,
                                      int max_num_neighbors, const int* bond_rowptr, const int* bond_src, int* scratch, int* nbr,
                                      int* rowptr, int* col, int* edst, unsigned char* ebond, jamun_stream_t stream) {
    JB_CHECK_ARG(pos && chain_ptr && scratch && nbr && rowptr && col && edst && ebond, "null argument");
    JB_CHECK_ARG(max_num_neighbors >= 0 && max_num_neighbors + 1 <= kCellMaxHits, "the cell-list search needs 0 <= max_num_neighbors < 64");
    JB_CHECK_ARG(max_chain <= kCellMaxAtoms, "chains longer than 8192 atoms: use jamun_radius_csr");
    cudaStream_t s = jb::as_stream(stream);
    const int max_hits = max_num_neighbors + 1;
    if (N > 0 && G > 0) {
        const size_t smem = (size_t)max_chain * 20 + (size_t)(2 * kCellMaxCells + 2 + 8 * kHitBuf) * sizeof(int);
        static size_t smem_set = 0;
        if (smem > smem_set) {
            cudaError_t e = cudaFuncSetAttribute(radius_cell_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) {
                jb::set_error("jamun_radius_csr_cells: cudaFuncSetAttribute(%zu B): %s", smem, cudaGetErrorString(e));
                return JAMUN_ECUDA;
            }
            smem_set = smem;
        }
        radius_cell_kernel<<<G, 256, smem, s>>>(pos, chain_ptr, r2, r_cut, max_hits, bond_rowptr, scratch, nbr);
        JB_CHECK_LAUNCH();
    }
    exclusive_scan_kernel<<<1, 1024, 0, s>>>(scratch, rowptr, N);
    JB_CHECK_LAUNCH();
    if (N > 0) {
        radius_compact_kernel<<<(N + 127) / 128, 128, 0, s>>>(nbr, max_hits, bond_rowptr, bond_src, rowptr, N, col, edst, ebond);
        JB_CHECK_LAUNCH();
    }
    return JAMUN_OK;
}

extern "C" int jamun_csr_by_source(const int* rowptr, const int* col, int N, int cap, int* scratch, int* src_rowptr, int* src_eid,
                                   jamun_stream_t stream) {
    JB_CHECK_ARG(rowptr && col && scratch && src_rowptr && src_eid, "null argument");
    cudaStream_t s = jb::as_stream(stream);
    if (N == 0 || cap == 0) return JAMUN_OK;
    int blocks = (cap + 255) / 256;
    if (blocks > jb::kNumSMs * 8) blocks = jb::kNumSMs * 8;
    cudaMemsetAsync(scratch, 0, (size_t)(N + 1) * sizeof(int), s);
    count_sources_kernel<<<blocks, 256, 0, s>>>(rowptr, col, N, scratch);
    exclusive_scan_kernel<<<1, 1024, 0, s>>>(scratch, src_rowptr, N);
    cudaMemsetAsync(scratch, 0, (size_t)(N + 1) * sizeof(int), s);
    fill_sources_kernel<<<blocks, 256, 0, s>>>(rowptr, col, N, src_rowptr, scratch, src_eid);
    sort_sources_kernel<<<(int)(((size_t)N * 32 + 255) / 256), 256, 0, s>>>(src_rowptr, N, src_eid);
    JB_CHECK_LAUNCH();
    return JAMUN_OK;
}

extern "C" int jamun_edge_geom(const float* p, const int* rowptr, const int* col, const int* edst, int N, int cap,
                               const float* mu, float step, float* rhat, float* rb, jamun_stream_t stream) {
    JB_CHECK_ARG(p && rowptr && col && edst && mu && rhat && rb, "null argument");
    if (N == 0 || cap == 0) return JAMUN_OK;
    long long warps = (cap + 31) / 32;
    int blocks = (int)((warps * 32 + 255) / 256);
    int max_blocks = jb::kNumSMs * 16;
    if (blocks > max_blocks) blocks = max_blocks;
    edge_geom_kernel<<<blocks, 256, 0, jb::as_stream(stream)>>>(p, rowptr, col, edst, N, mu, step, rhat, rb);
    JB_CHECK_LAUNCH();
    return JAMUN_OK;
}

extern "C" int jamun_edge_radial_hidden(const float* rb, const unsigned char* ebond, const int* rowptr, int N, int cap,
                                        const float* w0r, const float* b0eff, float* h, jamun_stream_t stream) {
    JB_CHECK_ARG(rb && ebond && rowptr && w0r && b0eff && h, "null argument");
    if (N == 0 || cap == 0) return JAMUN_OK;
    // one CTA per 128-edge tile of the capacity; the live edge count is device-side (rowptr[N]), surplus CTAs exit
    int blocks = (cap + kRhTile - 1) / kRhTile;
    edge_radial_hidden_kernel<<<blocks, 128, 0, jb::as_stream(stream)>>>(rb, ebond, rowptr, N, w0r, b0eff, h, 1, 0);
    JB_CHECK_LAUNCH();
    return JAMUN_OK;
}

extern "C" int jamun_edge_radial_hidden_all(const float* rb, const unsigned char* ebond, const int* rowptr, int N, int cap,
                                            const float* w0r_all, const float* b0eff_all, int layers, float* h_all,
                                            jamun_stream_t stream) {
    JB_CHECK_ARG(rb && ebond && rowptr && w0r_all && b0eff_all && h_all && layers >= 1, "bad argument");
    if (N == 0 || cap == 0) return JAMUN_OK;
    int blocks = (cap + kRhTile - 1) / kRhTile;
    edge_radial_hidden_kernel<<<blocks, 128, 0, jb::as_stream(stream)>>>(rb, ebond, rowptr, N, w0r_all, b0eff_all, h_all, layers,
                                                                         (size_t)cap * JAMUN_EDGE_HID);
    JB_CHECK_LAUNCH();
    return JAMUN_OK;
}

// ---- K2b on the warp-level tensor cores ---------------------------------------------------------------------------------------
// The same [E x 32] . [32 x 64] product per layer as mma.sync.m16n8k8 (tf32 operands, fp32 accumulate) with the three-product
// split for fp32 accuracy (hi = v & 0xFFFFE000, lo = v - hi; lo.hi + hi.lo + hi.hi).  A warp owns 16 edges: their radial basis is
// loaded and split once (32 registers) and reused by every layer; the weights arrive as pre-split, fragment-ordered images
// ([layer][k-step][n-tile][lane] x {b0.hi, b1.hi, b0.lo, b1.lo}, built once per plan by radial_pack_frag_kernel; 96 KB for six
// layers, L1-resident) so that a B fragment pair is one coalesced 16-byte load per lane.  Bias + SiLU in the epilogue; every store
// instruction writes eight full 32-byte sectors.  Per 16 edges and layer: 96 HMMA + 32 loads, against 2048 FFMA2 before.
__global__ void radial_pack_frag_kernel(const float* __restrict__ w0r_all, int layers, uint4* __restrict__ img) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= layers * 4 * 8 * 32) return;
    const int lane = t & 31, nt = (t >> 5) & 7, ks = (t >> 8) & 3, l = t >> 10;
    const int g = lane >> 2, tig = lane & 3;
    const float* w = w0r_all + (size_t)l * JAMUN_NBASIS * JAMUN_EDGE_HID;
    const float b0 = w[(8 * ks + tig) * JAMUN_EDGE_HID + 8 * nt + g], b1 = w[(8 * ks + tig + 4) * JAMUN_EDGE_HID + 8 * nt + g];
    uint4 o;
    o.x = __float_as_uint(b0) & 0xFFFFE000u, o.y = __float_as_uint(b1) & 0xFFFFE000u;
    o.z = __float_as_uint(b0 - __uint_as_float(o.x)), o.w = __float_as_uint(b1 - __uint_as_float(o.y));
    img[t] = o;
}

__device__ __forceinline__ void mma_tf32_16x8x8(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(128) edge_radial_hidden_mma_kernel(const float* __restrict__ rb, const unsigned char* __restrict__ ebond,
                                                                     const int* __restrict__ rowptr, int N,
                                                                     const uint4* __restrict__ img, const float* __restrict__ b0eff,
                                                                     float* __restrict__ h, int L, size_t h_stride) {
    const int E = rowptr[N];
    const int lane = threadIdx.x & 31, g = lane >> 2, tig = lane & 3;
    const int e_base = 16 * (blockIdx.x * 4 + (threadIdx.x >> 5));
    if (e_base >= E) return;
    const int r0 = e_base + g, r1 = r0 + 8;
    const bool on0 = r0 < E, on1 = r1 < E;
    uint32_t ahi[4][4], alo[4][4];
    {
        float a[4][4];
        const float* p0 = rb + (size_t)r0 * JAMUN_NBASIS + tig;
        const float* p1 = rb + (size_t)r1 * JAMUN_NBASIS + tig;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            a[ks][0] = on0 ? p0[8 * ks] : 0.f;
            a[ks][1] = on1 ? p1[8 * ks] : 0.f;
            a[ks][2] = on0 ? p0[8 * ks + 4] : 0.f;
            a[ks][3] = on1 ? p1[8 * ks + 4] : 0.f;
        }
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                ahi[ks][q] = __float_as_uint(a[ks][q]) & 0xFFFFE000u;
                alo[ks][q] = __float_as_uint(a[ks][q] - __uint_as_float(ahi[ks][q]));
            }
    }
    const int f0 = on0 && ebond[r0] ? JAMUN_EDGE_HID : 0, f1 = on1 && ebond[r1] ? JAMUN_EDGE_HID : 0;
    for (int l = 0; l < L; ++l) {
        float acc[8][4];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
        const uint4* im = img + (size_t)l * 4 * 8 * 32 + lane;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            uint4 b[8];
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) b[nt] = im[(ks * 8 + nt) * 32];
            // term by term across the n-tiles: consecutive MMAs never wait on each other's accumulator
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) mma_tf32_16x8x8(acc[nt], alo[ks], b[nt].x, b[nt].y);
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) mma_tf32_16x8x8(acc[nt], ahi[ks], b[nt].z, b[nt].w);
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) mma_tf32_16x8x8(acc[nt], ahi[ks], b[nt].x, b[nt].y);
        }
        const float* bias = b0eff + (size_t)l * 2 * JAMUN_EDGE_HID + 2 * tig;
        float* hl = h + (size_t)l * h_stride + 2 * tig;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            if (on0) {
                const float2 bb = *reinterpret_cast<const float2*>(bias + f0 + 8 * nt);
                *reinterpret_cast<float2*>(hl + (size_t)r0 * JAMUN_EDGE_HID + 8 * nt) =
                    make_float2(jb::siluf_fast(acc[nt][0] + bb.x), jb::siluf_fast(acc[nt][1] + bb.y));
            }
            if (on1) {
                const float2 bb = *reinterpret_cast<const float2*>(bias + f1 + 8 * nt);
                *reinterpret_cast<float2*>(hl + (size_t)r1 * JAMUN_EDGE_HID + 8 * nt) =
                    make_float2(jb::siluf_fast(acc[nt][2] + bb.x), jb::siluf_fast(acc[nt][3] + bb.y));
            }
        }
    }
}

extern "C" int jamun_radial_pack_frag(const float* w0r_all, int layers, float* img, jamun_stream_t stream) {
    JB_CHECK_ARG(w0r_all && img && layers >= 1, "bad argument");
    const int total = layers * 4 * 8 * 32;
    radial_pack_frag_kernel<<<(total + 255) / 256, 256, 0, jb::as_stream(stream)>>>(w0r_all, layers, reinterpret_cast<uint4*>(img));
    JB_CHECK_LAUNCH();
    return JAMUN_OK;
}

extern "C" int jamun_edge_radial_hidden_mma(const float* rb, const unsigned char* ebond, const int* rowptr, int N, int cap,
                                            const float* img, const float* b0eff_all, int layers, float* h_all,
                                            jamun_stream_t stream) {
    JB_CHECK_ARG(rb && ebond && rowptr && img && b0eff_all && h_all && layers >= 1, "bad argument");
    if (N == 0 || cap == 0) return JAMUN_OK;
    // one warp per 16-edge tile of the capacity; the live edge count is device-side (rowptr[N]), surplus warps exit
    const int blocks = (cap + 63) / 64;
    edge_radial_hidden_mma_kernel<<<blocks, 128, 0, jb::as_stream(stream)>>>(rb, ebond, rowptr, N, reinterpret_cast<const uint4*>(img),
                                                                             b0eff_all, h_all, layers, (size_t)cap * JAMUN_EDGE_HID);
    JB_CHECK_LAUNCH();
    return JAMUN_OK;
}

extern "C" int jamun_layout_to_soa(const float* in, int s, int v, int N, float* out, jamun_stream_t stream) {
    JB_CHECK_ARG(in && out && s >= 0 && v >= 0, "bad argument");
    if (N == 0) return JAMUN_OK;
    layout_kernel<true><<<jb::kNumSMs * 4, 256, 0, jb::as_stream(stream)>>>(in, s, v, N, out);
    JB_CHECK_LAUNCH();
    return JAMUN_OK;
}

extern "C" int jamun_layout_from_soa(const float* in, int s, int v, int N, float* out, jamun_stream_t stream) {
    JB_CHECK_ARG(in && out && s >= 0 && v >= 0, "bad argument");
    if (N == 0) return JAMUN_OK;
    layout_kernel<false><<<jb::kNumSMs * 4, 256, 0, jb::as_stream(stream)>>>(in, s, v, N, out);
    JB_CHECK_LAUNCH();
    return JAMUN_OK;
}
