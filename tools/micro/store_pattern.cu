// Micro-benchmark: HBM write rate of the aggregate builder's store pattern (a node's 715 operand lines, each a full 128-byte
// line, 16 KB x nslots apart, streaming stores) against a sequential stream of the same size, 148 persistent CTAs x 8 warps.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/store_pattern tools/micro/store_pattern.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int kRows = 18304, kTiles = kRows / 128;
constexpr int kSt0 = 65 * 5, kSt1 = 65 * 2;  // stages of the 0e operand and of each 1e component

// mode 0: builder pattern, tile-major; mode 1: same lines, but the 65 channel lines of a (node, slot) adjacent (8.3 KB runs);
// mode 2: sequential stream
__global__ void __launch_bounds__(256, 1) store_kernel(float* a0, float* a1, size_t comp, int mode, int rows) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wq = warp & 3, es = warp >> 2;
    unsigned na = 0;
    for (int r = blockIdx.x; r < rows; r += gridDim.x, ++na) {
        if ((na & 1u) != (unsigned)es) continue;
        for (int t = 0; t < 3; ++t) {
            const bool zero = t == 0 || (t == 1 && wq == 0);
            const int nsl = zero ? 5 : 2;
            const int slot = t == 0 ? wq : zero ? 4 : (t == 1 ? 0 : 1);
            if (t == 2 && wq == 3) continue;
            float* base = zero ? a0 : a1 + (size_t)(t == 1 ? wq - 1 : wq) * comp;
            float* dst;
            size_t kstride;
            if (mode == 0) {
                dst = base + ((size_t)(r >> 7) * (65 * nsl) + slot) * 4096 + (size_t)(r & 127) * 32;
                kstride = (size_t)nsl * 4096;
            } else {
                dst = base + (((size_t)(r >> 7) * nsl + slot) * 128 + (r & 127)) * (65 * 32);
                kstride = 32;
            }
#pragma unroll 13
            for (int k = 0; k < 65; ++k) __stcs(dst + k * kstride + lane, (float)k);
        }
    }
}
__global__ void seq_kernel(float* a, size_t n) {
    const size_t per = n / gridDim.x;
    float* p = a + (size_t)blockIdx.x * per;
    for (size_t i = threadIdx.x; i < per; i += blockDim.x) __stcs(p + i, 1.0f);
}
__global__ void seq4_kernel(float4* a, size_t n4) {
    const size_t per = n4 / gridDim.x;
    float4* p = a + (size_t)blockIdx.x * per;
    for (size_t i = threadIdx.x; i < per; i += blockDim.x) __stcs(p + i, make_float4(1.f, 2.f, 3.f, 4.f));
}
// builder pattern with `nw` warps per CTA: warp w takes lane quarter w & 3 and the channels k == (w >> 2) mod (nw / 4)
__global__ void store_kernel_w(float* a0, float* a1, size_t comp, int rows, int v4) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wq = warp & 3, part = warp >> 2, nparts = blockDim.x >> 7;
    for (int r = blockIdx.x; r < rows; r += gridDim.x) {
        for (int t = 0; t < 3; ++t) {
            const bool zero = t == 0 || (t == 1 && wq == 0);
            const int nsl = zero ? 5 : 2;
            const int slot = t == 0 ? wq : zero ? 4 : (t == 1 ? 0 : 1);
            if (t == 2 && wq == 3) continue;
            float* base = zero ? a0 : a1 + (size_t)(t == 1 ? wq - 1 : wq) * comp;
            float* dst = base + ((size_t)(r >> 7) * (65 * nsl) + slot) * 4096 + (size_t)(r & 127) * 32;
            const size_t kstride = (size_t)nsl * 4096;
            if (v4) {  // 8 lanes x 16 B per line, 4 lines (channels) per instruction
                for (int k = part * 4 + (lane >> 3); k < 65; k += nparts * 4)
                    __stcs(reinterpret_cast<float4*>(dst + k * kstride) + (lane & 7), make_float4(1.f, 2.f, 3.f, 4.f));
            } else {
                for (int k = part; k < 65; k += nparts) __stcs(dst + k * kstride + lane, (float)k);
            }
        }
    }
}

int main() {
    const size_t n0 = (size_t)kSt0 * kRows * 32, n1 = (size_t)kSt1 * kRows * 32, n = n0 + 3 * n1;
    float* a;
    cudaMalloc(&a, n * sizeof(float));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const double bytes_b = 18222.0 * 65 * (160 + 3 * 64) * 4;
    auto run = [&](const char* name, double bytes, auto launch) {
        float best = 1e9f;
        for (int rep = 0; rep < 6; ++rep) {
            cudaEventRecord(e0);
            launch();
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            if (rep > 0 && ms < best) best = ms;
        }
        printf("%-64s %.3f ms, %5.0f GB/s  [%s]\n", name, best, bytes / best / 1e6, cudaGetErrorString(cudaGetLastError()));
    };
    run("builder pattern, 8 warps, two alternating sets (as the kernel)", bytes_b, [&] { store_kernel<<<148, 256>>>(a, a + n0, n1, 0, 18222); });
    run("channel lines adjacent, 8 warps", bytes_b, [&] { store_kernel<<<148, 256>>>(a, a + n0, n1, 1, 18222); });
    for (int nw : {4, 8, 16, 32}) {
        char nm[96];
        snprintf(nm, sizeof nm, "builder pattern, %d warps on one node, 4 B per lane", nw);
        run(nm, bytes_b, [&] { store_kernel_w<<<148, 32 * nw>>>(a, a + n0, n1, 18222, 0); });
        snprintf(nm, sizeof nm, "builder pattern, %d warps on one node, 16 B per lane", nw);
        run(nm, bytes_b, [&] { store_kernel_w<<<148, 32 * nw>>>(a, a + n0, n1, 18222, 1); });
    }
    for (int thr : {256, 512, 1024}) {
        char nm[96];
        snprintf(nm, sizeof nm, "sequential, %d threads, 4 B per lane", thr);
        run(nm, (double)n * 4, [&] { seq_kernel<<<148, thr>>>(a, n); });
        snprintf(nm, sizeof nm, "sequential, %d threads, 16 B per lane", thr);
        run(nm, (double)n * 4, [&] { seq4_kernel<<<148, thr>>>(reinterpret_cast<float4*>(a), n / 4); });
    }
    run("sequential, 592 CTAs x 256 threads, 16 B per lane", (double)n * 4, [&] { seq4_kernel<<<592, 256>>>(reinterpret_cast<float4*>(a), n / 4); });
    return 0;
}
