// Developer microbenchmark (not part of the product): single-warp latencies / issue rates on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
#define N 512
__device__ __forceinline__ unsigned tf32_hi(float x) { unsigned r; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x)); return r; }
__device__ __forceinline__ void mma(float (&d)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__global__ void bench(float* out, long long* cyc, const int* chase_g, int chase_n, int mode_warps) {
    __shared__ int chase_s[1024];
    __shared__ float fs[2048];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) chase_s[i] = (i + 33) & 1023;
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) fs[i] = 1.0f + i * 1e-6f;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    long long t0, t1;
    float acc = lane * 1e-3f, x = 1.0001f;
    // 1. dependent FFMA
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) acc = fmaf(acc, x, 0.5f);
    t1 = clock64(); if (threadIdx.x == 0) cyc[0] = t1 - t0;
    // 2. 8 independent FFMA chains
    float a8[8]; for (int j = 0; j < 8; ++j) a8[j] = acc + j;
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) a8[j] = fmaf(a8[j], x, 0.5f);
    }
    t1 = clock64(); if (threadIdx.x == 0) cyc[1] = t1 - t0;
    for (int j = 0; j < 8; ++j) acc += a8[j];
    // 3. dependent LDS.32 chain
    int p = lane;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) p = chase_s[p];
    t1 = clock64(); if (threadIdx.x == 0) cyc[2] = t1 - t0;
    acc += p;
    // 4. dependent mma chain, 5. four independent mma chains
    unsigned fa[4] = {tf32_hi(acc), tf32_hi(x), tf32_hi(1.5f), tf32_hi(0.25f)}, fb[2] = {tf32_hi(0.5f), tf32_hi(0.125f)};
    float d[4] = {0, 0, 0, 0};
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) mma(d, fa, fb);
    t1 = clock64(); if (threadIdx.x == 0) cyc[3] = t1 - t0;
    float d4[4][4] = {};
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N; ++i) {
#pragma unroll
        for (int j = 0; j < 4; ++j) mma(d4[j], fa, fb);
    }
    t1 = clock64(); if (threadIdx.x == 0) cyc[4] = t1 - t0;
    for (int j = 0; j < 4; ++j) acc += d4[j][0] + d4[j][3];
    acc += d[0] + d[1] + d[2] + d[3];
    // 6. dependent DFMA, 7. dependent DDIV
    double da = acc, dx = 1.0000001;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) da = fma(da, dx, 0.5);
    t1 = clock64(); if (threadIdx.x == 0) cyc[5] = t1 - t0;
    t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < N; ++i) da = __ddiv_rn(da + 3.0, dx);
    t1 = clock64(); if (threadIdx.x == 0) cyc[6] = t1 - t0;
    // 8. dependent shfl
    int sv = lane;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) sv = __shfl_xor_sync(0xffffffffu, sv, 1) + 1;
    t1 = clock64(); if (threadIdx.x == 0) cyc[7] = t1 - t0;
    // 9. global pointer chase: first pass (L2 / DRAM), second pass over a small ring (L1)
    int q = lane * 32 % chase_n;
    t0 = clock64();
    for (int i = 0; i < N; ++i) q = chase_g[q];
    t1 = clock64(); if (threadIdx.x == 0) cyc[8] = t1 - t0;
    int q2 = 0;
    for (int i = 0; i < 64; ++i) q2 = chase_g[chase_n + q2];   // warm a 64-entry ring
    t0 = clock64();
    for (int i = 0; i < N; ++i) q2 = chase_g[chase_n + q2];
    t1 = clock64(); if (threadIdx.x == 0) cyc[9] = t1 - t0;
    // 10. __syncthreads
    t0 = clock64();
    for (int i = 0; i < N; ++i) __syncthreads();
    t1 = clock64(); if (threadIdx.x == 0) cyc[10] = t1 - t0;
    // 11. LDS.32 x6 + 8 dependent cvt/sub + 3 mma (one k-step of the MLP tile), loop-carried through d only
    float dd[4] = {0, 0, 0, 0};
    t0 = clock64();
#pragma unroll 2
    for (int i = 0; i < N; ++i) {
        const float* s = fs + ((i * 8) & 1023) + (lane & 3);
        float a0 = s[0], a1 = s[40], a2 = s[4], a3 = s[44], b0 = s[80], b1 = s[120];
        unsigned ah[4] = {tf32_hi(a0), tf32_hi(a1), tf32_hi(a2), tf32_hi(a3)}, bh[2] = {tf32_hi(b0), tf32_hi(b1)};
        unsigned al[4] = {__float_as_uint(a0 - __uint_as_float(ah[0])), __float_as_uint(a1 - __uint_as_float(ah[1])),
                          __float_as_uint(a2 - __uint_as_float(ah[2])), __float_as_uint(a3 - __uint_as_float(ah[3]))};
        unsigned bl[2] = {__float_as_uint(b0 - __uint_as_float(bh[0])), __float_as_uint(b1 - __uint_as_float(bh[1]))};
        mma(dd, al, bh); mma(dd, ah, bl); mma(dd, ah, bh);
    }
    t1 = clock64(); if (threadIdx.x == 0) cyc[11] = t1 - t0;
    out[threadIdx.x] = acc + (float)da + sv + q + q2 + dd[0] + dd[3];
}
int main() {
    float* out; long long* cyc; int* chase;
    const int n = 16 << 20;  // 64 MB ring -> mostly L2 / DRAM misses
    cudaMalloc(&out, 4096); cudaMalloc(&cyc, 16 * 8); cudaMalloc(&chase, (n + 64) * 4);
    int* h = (int*)malloc((n + 64) * 4);
    for (int i = 0; i < n; ++i) h[i] = (int)(((long long)i * 1000003 + 12345) % n);
    for (int i = 0; i < 64; ++i) h[n + i] = (i + 1) & 63;
    cudaMemcpy(chase, h, (n + 64) * 4, cudaMemcpyHostToDevice);
    const char* names[] = {"dependent FFMA", "8 independent FFMA chains (per FFMA)", "dependent LDS.32", "dependent mma.m16n8k8.tf32",
        "4 independent mma chains (per mma)", "dependent DFMA", "dependent DDIV(+DADD)", "dependent SHFL(+IADD)", "global chase 64MB (L2/DRAM)",
        "global chase 64-entry ring (L1)", "__syncthreads", "MLP k-step (6 LDS + split + 3 mma)"};
    const double div[] = {N, N * 8.0, N, N, N * 4.0, N, N, N, N, N, N, N};
    for (int warps : {1, 8}) {
        bench<<<1, 32 * warps>>>(out, cyc, chase, n, warps);
        bench<<<1, 32 * warps>>>(out, cyc, chase, n, warps);
        cudaDeviceSynchronize();
        long long c[16]; cudaMemcpy(c, cyc, 16 * 8, cudaMemcpyDeviceToHost);
        printf("---- %d warp(s) in the CTA, warp 0 timed; err=%s\n", warps, cudaGetErrorString(cudaGetLastError()));
        for (int i = 0; i < 12; ++i) printf("%-42s %8.1f cycles\n", names[i], c[i] / div[i]);
    }
    return 0;
}
