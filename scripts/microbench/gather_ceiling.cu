// gather_ceiling.cu -- what random 512-byte row gathers can reach on this GPU (the real ceiling of
// K1's dominant access pattern), as a function of loads in flight per warp and warps per SM.
// Each warp reads `iters` batches of U random rows (one 16-byte load per lane per row) and folds
// them into a checksum. Independent batches (no dependent chain): this is the bandwidth ceiling,
// not a latency test. Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gather_ceiling gather_ceiling.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

template <int U>
__global__ void gather_kernel(const float4 *__restrict__ arena, uint32_t n_rows, uint32_t iters, float *out) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    uint64_t s = 0x9E3779B97F4A7C15ull * (warp + 1);
    float acc = 0.f;
    for (uint32_t it = 0; it < iters; ++it) {
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            s = s * 6364136223846793005ull + 1442695040888963407ull;
            const uint32_t row = static_cast<uint32_t>((s >> 33) % n_rows);
            v[u] = __ldg(arena + static_cast<size_t>(row) * 32 + lane);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) acc += v[u].x + v[u].y + v[u].z + v[u].w;
    }
    if (acc == 123.456f) out[0] = acc;
}

template <int U>
void run(const float4 *arena, uint32_t n_rows, int warps_per_sm, int sms, float *out) {
    const int block = 32;
    const uint32_t iters = 4096 / U;
    // occupancy is forced with dynamic shared memory: 227 KB / warps_per_sm per block
    int smem = (227 * 1024) / warps_per_sm - 1024;
    if (smem < 0) smem = 0;
    cudaFuncSetAttribute(gather_kernel<U>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int grid = sms * warps_per_sm;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    gather_kernel<U><<<grid, block, smem>>>(arena, n_rows, iters, out);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    gather_kernel<U><<<grid, block, smem>>>(arena, n_rows, iters, out);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double bytes = static_cast<double>(grid) * iters * U * 512.0;
    printf("{\"rows\": %u, \"U\": %d, \"warps_per_sm\": %d, \"ms\": %.3f, \"GBps\": %.1f}\n", n_rows, U, warps_per_sm, ms, bytes / ms / 1e6);
}

int main(int argc, char **argv) {
    const uint32_t n_rows = argc > 1 ? atoi(argv[1]) : 1000000;
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    float4 *arena; float *out;
    cudaMalloc(&arena, static_cast<size_t>(n_rows) * 512);
    cudaMemset(arena, 0, static_cast<size_t>(n_rows) * 512);
    cudaMalloc(&out, 4);
    for (int w : {8, 16, 24, 32}) {
        run<2>(arena, n_rows, w, p.multiProcessorCount, out);
        run<4>(arena, n_rows, w, p.multiProcessorCount, out);
        run<8>(arena, n_rows, w, p.multiProcessorCount, out);
        run<16>(arena, n_rows, w, p.multiProcessorCount, out);
    }
    return 0;
}
