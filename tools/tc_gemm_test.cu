// Developer bring-up test (not part of the product): D[128 x 64] = A[128 x K] * B[64 x K]^T with tcgen05.mma
// kind::tf32, operands in shared memory in the K-major no-swizzle core-matrix layout, accumulator in TMEM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/tc_gemm_test tools/tc_gemm_test.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

constexpr int BM = 128, BN = 64;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, SWIZZLE_NONE: element (r, k) of an operand tile with K columns (fp32) lives at
//   (r / 8) * SBO + (k / 4) * 128 + (r % 8) * 16 + (k % 4) * 4 bytes,   SBO = (K / 4) * 128
__host__ __device__ inline int canon_off_floats(int r, int k, int K) {
    return (r >> 3) * (K >> 2) * 32 + (k >> 2) * 32 + (r & 7) * 4 + (k & 3);
}

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)(lbo_bytes >> 4) << 16;
    d |= (uint64_t)(sbo_bytes >> 4) << 32;
    d |= 1ull << 46;  // descriptor version for sm_100
    return d;         // layout_type 0 = no swizzle
}

__global__ void __launch_bounds__(128) gemm_test(const float* __restrict__ A, const float* __restrict__ B,
                                                 float* __restrict__ D, int K, int* status) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    float* sA = reinterpret_cast<float*>(smem_raw);
    float* sB = sA + BM * K;
    __shared__ __align__(8) unsigned long long mbar;
    __shared__ uint32_t tmem_base_sh;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < BM * K; i += 128) {
        const int r = i / K, k = i % K;
        sA[canon_off_floats(r, k, K)] = A[i];
    }
    for (int i = tid; i < BN * K; i += 128) {
        const int r = i / K, k = i % K;
        sB[canon_off_floats(r, k, K)] = B[i];
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_sh)), "n"(64));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy smem writes -> visible to the async proxy
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_sh;
    if (tid == 0) {
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
        const uint32_t sbo = (uint32_t)(K >> 2) * 128u;
        const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
        for (int ks = 0; ks < K / 8; ++ks) {
            const uint64_t da = make_desc(a0 + ks * 256, 128, sbo);
            const uint64_t db = make_desc(b0 + ks * 256, 128, sbo);
            const uint32_t acc = ks > 0 ? 1u : 0u;
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_base),
                "l"(da), "l"(db), "r"(idesc), "r"(acc)
                : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
    }
    // wait for the MMAs (bounded spin: a bring-up bug must not hang the GPU)
    {
        uint32_t done = 0;
        for (int spin = 0; spin < (1 << 22) && !done; ++spin) {
            asm volatile(
                "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                : "=r"(done)
                : "r"(smem_u32(&mbar)), "r"(0u)
                : "memory");
        }
        if (!done) {
            if (tid == 0) *status = 1;
            __syncthreads();
            if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(64));
            return;
        }
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // epilogue: warp w owns TMEM lanes 32 w .. 32 w + 31 (== rows), 32 columns per load
    for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t v[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(32 * warp) << 16) + (uint32_t)c0;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
              "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
              "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
              "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        const int row = 32 * warp + lane;
        for (int c = 0; c < 32; ++c) D[row * BN + c0 + c] = __uint_as_float(v[c]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(64));
}

static float tf32_trunc(float x) {
    uint32_t u;
    memcpy(&u, &x, 4);
    u &= 0xFFFFE000u;
    memcpy(&x, &u, 4);
    return x;
}

int main() {
    for (int K : {32, 64, 480}) {
        std::vector<float> A(BM * K), B(BN * K), D(BM * BN), ref(BM * BN), ref_t(BM * BN);
        srand(1);
        for (auto& x : A) x = (rand() % 2001 - 1000) / 1000.0f;
        for (auto& x : B) x = (rand() % 2001 - 1000) / 1000.0f;
        for (int i = 0; i < BM; ++i)
            for (int j = 0; j < BN; ++j) {
                double s = 0, st = 0;
                for (int k = 0; k < K; ++k) {
                    s += (double)A[i * K + k] * B[j * K + k];
                    st += (double)tf32_trunc(A[i * K + k]) * tf32_trunc(B[j * K + k]);
                }
                ref[i * BN + j] = (float)s;
                ref_t[i * BN + j] = (float)st;
            }
        float *dA, *dB, *dD;
        int* dS;
        cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4); cudaMalloc(&dS, 4);
        cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
        cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
        cudaMemset(dD, 0, D.size() * 4); cudaMemset(dS, 0, 4);
        const size_t smem = (size_t)(BM + BN) * K * 4 + 1024;
        cudaFuncSetAttribute(gemm_test, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        gemm_test<<<1, 128, smem>>>(dA, dB, dD, K, dS);
        cudaError_t e = cudaDeviceSynchronize();
        int st = 0;
        cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost);
        double err = 0, err_t = 0;
        for (size_t i = 0; i < D.size(); ++i) {
            err = fmax(err, fabs(D[i] - ref[i]));
            err_t = fmax(err_t, fabs(D[i] - ref_t[i]));
        }
        printf("K=%d cuda=%s status=%d max|D-ref|=%g max|D-ref(tf32-truncated operands)|=%g  D[0]=%g ref=%g D[last]=%g ref=%g\n", K,
               cudaGetErrorString(e), st, err, err_t, D[0], ref[0], D.back(), ref.back());
        cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dS);
    }
    return 0;
}
