// Developer microbenchmark (not part of the product): cost of one mma_unit of tsp_fast.cuh in isolation.
#include <cstdio>
#include <cuda_runtime.h>
#include "../tree_search_planning_b200/csrc/tsp_fast.cuh"
using namespace tsp;

template <int KC, int VAR>
__global__ void bench(long long* cyc, int reps, int wlo, int active_warps) {
    extern __shared__ __align__(16) float smem[];
    const int tid = threadIdx.x, lane = tid & 31, wg = tid >> 5;
    for (int i = tid; i < 40000; i += blockDim.x) smem[i] = 1.0f + (i % 97) * 1e-3f;
    __shared__ int row_action[16];
    if (tid < 16) row_action[tid] = tid % 5;
    __syncthreads();
    const int rs = 624;  // 2 * 312
    float* W = smem;            // 16k floats of "weights"
    float* act = smem + 20000;  // strip
    const int g = lane >> 2, t = lane & 3;
    const float* actL = act + g * rs + 4 * t;
    const float* wL = W + lane * 4;
    const float* wT = W + 2 * t;
    int4 ua = make_int4(0, wg * KC * 128, 2 * (64 + 8 * wg), KC | 0x100 | 0x200 | (VAR == 1 ? 0x400 : 0));
    int4 ub = make_int4(8000 + 8 * wg, 9000, 32, 0);
    long long t0 = 0, t1 = 0;
    if (wg < active_warps) {
        for (int r = 0; r < 3; ++r) mma_unit<KC>(ua, ub, actL, wL, wT, 8 * rs, wlo ? 10000 : 0, row_action, g);
        t0 = clock64();
        for (int r = 0; r < reps; ++r) {
            mma_unit<KC>(ua, ub, actL, wL, wT, 8 * rs, wlo ? 10000 : 0, row_action, g);
            if (VAR == 2) asm volatile("bar.sync 1, 256;" ::: "memory");
        }
        t1 = clock64();
    } else if (VAR == 2) {
        for (int r = 0; r < reps; ++r) asm volatile("bar.sync 1, 256;" ::: "memory");
    }
    if (lane == 0) cyc[blockIdx.x * 32 + wg] = t1 - t0;
}

template <int KC, int VAR>
void run(const char* name, int threads, int active, int wlo, int ctas) {
    long long* cyc;
    cudaMalloc(&cyc, 8 * 32 * 256);
    cudaFuncSetAttribute(bench<KC, VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200000);
    const int reps = 2000;
    bench<KC, VAR><<<ctas, threads, 200000>>>(cyc, reps, wlo, active);
    cudaDeviceSynchronize();
    long long h[32];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    printf("%-34s KC=%d threads=%4d active_warps=%2d wlo=%d ctas=%3d: warp0 %7.1f cyc/unit  last %7.1f  (%s)\n", name, KC, threads,
           active, wlo, ctas, (double)h[0] / reps, (double)h[active - 1] / reps, cudaGetErrorString(cudaGetLastError()));
    cudaFree(cyc);
}

int main() {
    for (int wlo = 0; wlo < 2; ++wlo) {
        run<4, 0>("unit", 32, 1, wlo, 1);
        run<4, 0>("unit", 128, 4, wlo, 1);
        run<4, 0>("unit", 256, 8, wlo, 1);
        run<4, 0>("unit", 512, 16, wlo, 1);
        run<4, 1>("unit + one-hot", 256, 8, wlo, 1);
        run<4, 2>("unit + group barrier", 256, 8, wlo, 1);
        run<4, 2>("unit + group barrier, 4 of 8 work", 256, 4, wlo, 1);
        run<2, 0>("unit", 256, 8, wlo, 1);
        run<2, 2>("unit + group barrier", 256, 8, wlo, 1);
        run<4, 0>("unit, all SMs", 256, 8, wlo, 148);
    }
    return 0;
}
