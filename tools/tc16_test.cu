// Developer bring-up test (not part of the product): the scaled fp16 split with the A operand (weights) in TENSOR MEMORY.
//   D[128 features x 16 trees] = W[128 x K] * X[16 x K]^T,  v = v1 + v2 / 2048 (v1 = fp16(v), v2 = fp16((v - v1) * 2048))
//   acc0 | acc1 (32 columns) += W1 [X1 | X2]   (one MMA, N = 32),   acc1 += W2 X1  (N = 16):  D = acc0 + acc1 / 2048
// A in TMEM for kind::f16: lane = row, one 32-bit column holds TWO consecutive inputs (k = 2c in the low half, 2c + 1 in the
// high half), so K = 16 per MMA is 8 columns -- this is the layout assumption under test. B in shared memory, MN-major.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/tc16_test tools/tc16_test.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <vector>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

constexpr int M = 128, NT = 16, K = 64;
constexpr float SPLIT = 2048.0f;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mma_f16_ts(uint32_t d, uint32_t a_tmem, uint32_t b_lo, uint32_t b_hi, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 bd;\n\tsetp.ne.b32 p, %5, 0;\n\tmov.b64 bd, {%2, %3};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], bd, %4, p;\n\t}\n" ::"r"(d), "r"(a_tmem), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(acc)
        : "memory");
}

// b_mn: 1 = MN-major (element (n, k) at (n/8) SBO + (k/8) 128 + (k%8) 16 + (n%8) 2), 0 = K-major ((n/8) SBO + (k/8) 128 + (n%8) 16 + (k%8) 2)
__global__ void __launch_bounds__(128) test(const float* __restrict__ W, const float* __restrict__ X, float* __restrict__ D, int b_mn,
                                            int iters, long long* cycles) {
    __shared__ __align__(128) unsigned char xb[4 * (K / 8) * 128];  // X1: 2 groups of 8 trees, X2: 2 more
    __shared__ __align__(8) unsigned long long mbar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int sbo = (K / 8) * 128;
    for (int i = tid; i < NT * K; i += 128) {
        const int n = i / K, k = i % K;
        const float x = X[i];
        const __half h1 = __float2half_rn(x), h2 = __float2half_rn((x - __half2float(h1)) * SPLIT);
        const int off = b_mn ? (n >> 3) * sbo + (k >> 3) * 128 + (k & 7) * 16 + (n & 7) * 2
                             : (n >> 3) * sbo + (k >> 3) * 128 + (n & 7) * 16 + (k & 7) * 2;
        *reinterpret_cast<__half*>(xb + off) = h1;
        *reinterpret_cast<__half*>(xb + 2 * sbo + off) = h2;
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "n"(128));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    // weights -> TMEM: columns [0, K/2) hi pairs, [K/2, K) lo pairs; accumulators at column K (acc0: 16, acc1: 16)
    {
        const int row = 32 * warp + lane;
        for (int c0 = 0; c0 < K; c0 += 8) {
            uint32_t r[8];
            for (int i = 0; i < 8; ++i) {
                const int c = c0 + i, lo = c >= K / 2, cc = lo ? c - K / 2 : c;
                __half h[2];
                for (int u = 0; u < 2; ++u) {
                    const float w = W[row * K + 2 * cc + u];
                    const __half h1 = __float2half_rn(w);
                    h[u] = lo ? __float2half_rn((w - __half2float(h1)) * SPLIT) : h1;
                }
                r[i] = (uint32_t)__half_as_ushort(h[0]) | ((uint32_t)__half_as_ushort(h[1]) << 16);
            }
            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(tmem + ((32u * warp) << 16) + c0),
                         "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (tid == 0) {
        const uint32_t idesc32 = (1u << 4) | ((uint32_t)b_mn << 16) | ((uint32_t)(32 >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        const uint32_t idesc16 = (1u << 4) | ((uint32_t)b_mn << 16) | ((uint32_t)(16 >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        const uint32_t b_lo0 = ((smem_u32(xb) & 0x3FFFF) >> 4) | ((128u >> 4) << 16), b_hi = ((uint32_t)sbo >> 4) | (1u << 14);
        const uint32_t d = tmem + K;
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int q = 0; q < K / 16; ++q) {
                mma_f16_ts(d, tmem + 8 * q, b_lo0 + 16 * q, b_hi, idesc32, q > 0 ? 1u : 0u);           // W1 [X1 | X2] -> acc0 | acc1
                mma_f16_ts(d + 16, tmem + K / 2 + 8 * q, b_lo0 + 16 * q, b_hi, idesc16, 1u);            // W2 X1 -> acc1
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
        uint32_t done = 0;
        while (!done)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                         : "=r"(done) : "r"(smem_u32(&mbar)), "r"(0) : "memory");
        if (cycles) *cycles = clock64() - t0;
    }
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t v[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
          "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
          "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
          "=r"(v[31])
        : "r"(tmem + ((32u * warp) << 16) + K));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int n = 0; n < NT; ++n) D[(32 * warp + lane) * NT + n] = fmaf(__uint_as_float(v[16 + n]), 1.0f / SPLIT, __uint_as_float(v[n]));
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(128));
}

int main() {
    std::vector<float> W((size_t)M * K), X((size_t)NT * K), D((size_t)M * NT), ref((size_t)M * NT);
    srand(5);
    for (auto& x : W) x = (rand() % 20001 - 10000) / 10000.0f * 0.4f;
    for (auto& x : X) x = (rand() % 20001) / 20000.0f;
    for (int i = 0; i < M; ++i)
        for (int n = 0; n < NT; ++n) {
            double s = 0;
            for (int k = 0; k < K; ++k) s += (double)W[(size_t)i * K + k] * X[(size_t)n * K + k];
            ref[(size_t)i * NT + n] = (float)s;
        }
    float *dW, *dX, *dD; long long* dC;
    cudaMalloc(&dW, W.size() * 4); cudaMalloc(&dX, X.size() * 4); cudaMalloc(&dD, D.size() * 4); cudaMalloc(&dC, 8);
    cudaMemcpy(dW, W.data(), W.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dX, X.data(), X.size() * 4, cudaMemcpyHostToDevice);
    for (int b_mn = 0; b_mn < 2; ++b_mn) {
        cudaMemset(dD, 0, D.size() * 4);
        test<<<1, 128>>>(dW, dX, dD, b_mn, 1, dC);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
        double err = 0, mag = 0;
        for (size_t i = 0; i < D.size(); ++i) { err = fmax(err, fabs(D[i] - ref[i])); mag = fmax(mag, fabs(ref[i])); }
        printf("A in TMEM (f16 pairs), B %s-major: cuda=%s max|D-ref|=%.3g (max|ref|=%.3g) D[0]=%g ref=%g D[last]=%g ref=%g\n",
               b_mn ? "MN" : "K", cudaGetErrorString(e), err, mag, D[0], ref[0], D.back(), ref.back());
        if (e != cudaSuccess) return 2;
    }
    test<<<1, 128>>>(dW, dX, dD, 1, 1000, dC);
    cudaDeviceSynchronize();
    long long c = 0;
    cudaMemcpy(&c, dC, 8, cudaMemcpyDeviceToHost);
    printf("K = %d layer = %d MMAs (N = 32 + N = 16 per 16 inputs): %.1f cycles per layer, %.1f per MMA\n", K, 2 * K / 16, c / 1000.0,
           c / 1000.0 / (2 * K / 16));
    return 0;
}
