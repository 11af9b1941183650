// Developer bring-up test (not part of the product): the core of the weight-STREAMING tcgen05 MLP (tsp_stream.cuh).
//   D[128 features x 64 trees] = W[128 x K] * X[64 x K]^T at float32-level accuracy with kind::f16 and the SCALED fp16 split
//       v = v1 + v2 / 2048,  v1 = fp16(v), v2 = fp16((v - v1) * 2048)
//       acc0 += W1 X1,  acc1 += W1 X2 + W2 X1,  D = acc0 + acc1 / 2048         (3 MMAs per 16 inputs)
//   * A = weight tiles (128 rows x 64 inputs, fp16, 16 KB) streamed from global memory by TMA (cp.async.bulk.tensor.2d,
//     a ring of mbarrier-guarded slots), host-arranged in the K-major no-swizzle core-matrix order so that the bytes land
//     where the MMA descriptor expects them;
//   * B = the activations in shared memory, either K-major (what the gather of a hidden state writes: 16-byte chunks of 8
//     inputs, LBO = 144) or MN-major (what an epilogue thread = feature writes: 16-byte chunks of 8 trees);
//   * warp 4 (one elected lane) is the TMA producer, warp 5 (one elected lane) the MMA issuer, warps 0-3 run the epilogue.
// Also measures cycles per pass over the weights with 1 .. 148 CTAs streaming the same image (L2 -> SM bandwidth), the MMA
// rate alone (no TMA) and the TMA rate alone (no MMA).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/tcs_gemm_test tools/tcs_gemm_test.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <cmath>
#include <vector>
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

constexpr int BM = 128, BN = 64, KC = 64, K = 256;
constexpr int T = 2 * (K / KC);          // tiles per pass: W1[kc], W2[kc] alternating
constexpr int TILE_BYTES = BM * KC * 2;  // 16 KB
constexpr float SPLIT = 2048.0f, INV_SPLIT = 1.0f / 2048.0f;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr, uint32_t lbo) { return ((saddr & 0x3FFFF) >> 4) | ((lbo >> 4) << 16); }
__device__ __forceinline__ uint32_t desc_hi(uint32_t sbo) { return (sbo >> 4) | (1u << 14); }
__device__ __forceinline__ bool mbar_wait(uint32_t mbar, uint32_t parity) {
    uint32_t done = 0;
    for (int spin = 0; spin < (1 << 22) && !done; ++spin)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(done) : "r"(mbar), "r"(parity) : "memory");
    return done != 0;
}
// D[tmem] (+)= A[smem desc] * B[smem desc], kind::f16, descriptors as two 32-bit words each
__device__ __forceinline__ void mma_f16(uint32_t d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 ad, bd;\n\tsetp.ne.b32 p, %6, 0;\n\tmov.b64 ad, {%1, %2};\n\tmov.b64 bd, {%3, %4};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], ad, bd, %5, p;\n\t}\n" ::"r"(d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void commit(uint32_t mbar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar) : "memory");
}

// b_mn: B operand layout. 0: K-major, element (n, k) at (n/8) SBO + (k/8) 144 + (n%8) 16 + (k%8) 2, SBO = (K/8) 144;
//                         1: MN-major, element (n, k) at (n/8) SBO + (k/8) 128 + (k%8) 16 + (n%8) 2, SBO = (K/8) 128.
// mode: 0 full pipeline; 1 MMAs only (no TMA after the first NS tiles, no waits); 2 TMA only (the consumer frees the slot at once)
template <int NS, int NMMA>  // NMMA: columns of the MMA (trees)
__global__ void __launch_bounds__(192, 1) stream_test(const __grid_constant__ CUtensorMap tmap, const float* __restrict__ X,
                                                      float* __restrict__ D, int b_mn, int iters, int sync_epi, int mode,
                                                      int* status, long long* cycles) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);  // TMA destinations: 128-byte aligned
    unsigned char* ring = smem;                          // NS slots of 16 KB
    unsigned char* xb = smem + NS * TILE_BYTES;          // X1 then X2
    const int lbo_b = b_mn ? 128 : 144;
    const int sbo_b = (K / 8) * lbo_b;
    const int xbytes = (BN / 8) * sbo_b;
    __shared__ __align__(8) unsigned long long bars[2 * NS + 2];
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < BN * K; i += blockDim.x) {  // split the activations into the B operand (both images)
        const int n = i / K, k = i % K;
        const float x = X[i];
        const __half h1 = __float2half_rn(x);
        const __half h2 = __float2half_rn((x - __half2float(h1)) * SPLIT);
        const int off = b_mn ? (n >> 3) * sbo_b + (k >> 3) * 128 + (k & 7) * 16 + (n & 7) * 2
                             : (n >> 3) * sbo_b + (k >> 3) * 144 + (n & 7) * 16 + (k & 7) * 2;
        *reinterpret_cast<__half*>(xb + off) = h1;
        *reinterpret_cast<__half*>(xb + xbytes + off) = h2;
    }
    if (tid == 0) {
        for (int i = 0; i < 2 * NS + 2; ++i)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bars[i])), "r"(i == 2 * NS + 1 ? 4 : 1));
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
    const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[NS]), acc_ready = smem_u32(&bars[2 * NS]),
                   acc_free = smem_u32(&bars[2 * NS + 1]);
    const int total = iters * T;
    bool ok = true;
    const int warp_u = __shfl_sync(0xffffffffu, tid >> 5, 0);
    if (warp_u == 4) {
        // ================= TMA producer =================
        if (elect_one() && mode != 1) {
            int t = 0;
            for (int g = 0; g < total; ++g) {
                const int s = g & (NS - 1);
                if (g >= NS) ok = mbar_wait(empty0 + 8 * s, (uint32_t)(((g / NS) - 1) & 1)) && ok;
                const uint32_t dst = smem_u32(ring) + s * TILE_BYTES, mb = full0 + 8 * s;
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(TILE_BYTES) : "memory");
                asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                             ::"r"(dst), "l"(&tmap), "r"(0), "r"(t * 128), "r"(mb) : "memory");
                t = t + 1 == T ? 0 : t + 1;
            }
        }
        __syncwarp();
    } else if (warp_u == 5) {
        // ================= MMA issuer =================
        if (elect_one()) {
            const uint32_t idesc = (1u << 4) | ((uint32_t)b_mn << 16) | ((uint32_t)(NMMA >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
            const uint32_t a_hi = desc_hi((KC / 8) * 128), b_hi = desc_hi(sbo_b);
            const uint32_t bstep = (uint32_t)(2 * lbo_b) >> 4;
            const long long t0 = clock64();
            int g = 0;
            for (int it = 0; it < iters; ++it) {
                if (it > 0 && sync_epi) {  // the epilogue of the previous pass has read the accumulators
                    ok = mbar_wait(acc_free, (uint32_t)((it - 1) & 1)) && ok;
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                }
#pragma unroll
                for (int t = 0; t < T; ++t, ++g) {
                    const int s = g & (NS - 1), kc = t >> 1;
                    if (mode != 1) ok = mbar_wait(full0 + 8 * s, (uint32_t)((g / NS) & 1)) && ok;
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    if (mode == 2) {
                        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(empty0 + 8 * s) : "memory");
                        continue;
                    }
                    const uint32_t a_lo = desc_lo(smem_u32(ring) + s * TILE_BYTES, 128);
                    const uint32_t b1 = desc_lo(smem_u32(xb) + kc * (KC / 8) * lbo_b, lbo_b);
                    const uint32_t b2 = desc_lo(smem_u32(xb) + xbytes + kc * (KC / 8) * lbo_b, lbo_b);
#pragma unroll
                    for (int j = 0; j < KC / 16; ++j) {
                        if ((t & 1) == 0) {
                            mma_f16(tmem, a_lo + 16 * j, a_hi, b1 + bstep * j, b_hi, idesc, (t == 0 && j == 0) ? 0u : 1u);       // acc0 += W1 X1
                            mma_f16(tmem + 64, a_lo + 16 * j, a_hi, b2 + bstep * j, b_hi, idesc, (t == 0 && j == 0) ? 0u : 1u);  // acc1 += W1 X2
                        } else {
                            mma_f16(tmem + 64, a_lo + 16 * j, a_hi, b1 + bstep * j, b_hi, idesc, 1u);                            // acc1 += W2 X1
                        }
                    }
                    commit(empty0 + 8 * s);
                }
                if (sync_epi || it == iters - 1) commit(acc_ready);
            }
            ok = mbar_wait(empty0 + 8 * ((total - 1) & (NS - 1)), (uint32_t)(((total - 1) / NS) & 1)) && ok;
            if (cycles) cycles[blockIdx.x] = clock64() - t0;
        }
        __syncwarp();
    } else if (mode == 0) {
        for (int it = sync_epi ? 0 : iters - 1; it < iters; ++it) {
            ok = mbar_wait(acc_ready, sync_epi ? (uint32_t)(it & 1) : 0u) && ok;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            for (int c0 = 0; c0 < NMMA; c0 += 16) {
                uint32_t v0[16], v1[16];
                const uint32_t ta = tmem + ((uint32_t)(32 * warp) << 16) + (uint32_t)c0;
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                             : "=r"(v0[0]), "=r"(v0[1]), "=r"(v0[2]), "=r"(v0[3]), "=r"(v0[4]), "=r"(v0[5]), "=r"(v0[6]), "=r"(v0[7]),
                               "=r"(v0[8]), "=r"(v0[9]), "=r"(v0[10]), "=r"(v0[11]), "=r"(v0[12]), "=r"(v0[13]), "=r"(v0[14]), "=r"(v0[15])
                             : "r"(ta));
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                             : "=r"(v1[0]), "=r"(v1[1]), "=r"(v1[2]), "=r"(v1[3]), "=r"(v1[4]), "=r"(v1[5]), "=r"(v1[6]), "=r"(v1[7]),
                               "=r"(v1[8]), "=r"(v1[9]), "=r"(v1[10]), "=r"(v1[11]), "=r"(v1[12]), "=r"(v1[13]), "=r"(v1[14]), "=r"(v1[15])
                             : "r"(ta + 64));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (it == iters - 1 && blockIdx.x == 0)
                    for (int c = 0; c < 16; ++c)
                        D[(32 * warp + lane) * BN + c0 + c] = fmaf(__uint_as_float(v1[c]), INV_SPLIT, __uint_as_float(v0[c]));
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(acc_free) : "memory");
        }
    }
    if (!ok) atomicExch(status, 1);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(128));
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    EncodeFn encode = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qres) != cudaSuccess || !encode) {
        printf("cuTensorMapEncodeTiled not available\n");
        return 1;
    }
    std::vector<float> W((size_t)BM * K), X((size_t)BN * K), D((size_t)BM * BN), ref((size_t)BM * BN);
    srand(3);
    for (auto& x : W) x = (rand() % 20001 - 10000) / 10000.0f * 0.3f;
    for (auto& x : X) x = (rand() % 20001 - 10000) / 10000.0f * 2.0f;
    X[5] = 3.1e-6f; X[77] = -901.25f; W[9] = 1.7e-7f;
    for (int i = 0; i < BM; ++i)
        for (int j = 0; j < BN; ++j) {
            double s = 0;
            for (int k = 0; k < K; ++k) s += (double)W[(size_t)i * K + k] * X[(size_t)j * K + k];
            ref[(size_t)i * BN + j] = (float)s;
        }
    // weight image: tile t = 2 kc + lo, 16 KB each: [(m/8)][(k/8)][m%8][k%8] fp16 (K-major core matrices, LBO 128, SBO 1024)
    std::vector<__half> img((size_t)T * BM * KC);
    for (int kc = 0; kc < K / KC; ++kc)
        for (int m = 0; m < BM; ++m)
            for (int kk = 0; kk < KC; ++kk) {
                const float w = W[(size_t)m * K + kc * KC + kk];
                const __half h1 = __float2half_rn(w);
                const __half h2 = __float2half_rn((w - __half2float(h1)) * SPLIT);
                const size_t off = ((size_t)(m >> 3) * (KC / 8) + (kk >> 3)) * 64 + (m & 7) * 8 + (kk & 7);
                img[(size_t)(2 * kc) * BM * KC + off] = h1;
                img[(size_t)(2 * kc + 1) * BM * KC + off] = h2;
            }
    __half* dImg; float *dX, *dD; int* dS; long long* dC;
    cudaMalloc(&dImg, img.size() * 2); cudaMalloc(&dX, X.size() * 4); cudaMalloc(&dD, D.size() * 4); cudaMalloc(&dS, 4);
    cudaMalloc(&dC, 160 * 8);
    cudaMemcpy(dImg, img.data(), img.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dX, X.data(), X.size() * 4, cudaMemcpyHostToDevice);
    CUtensorMap tmap;
    const cuuint64_t gdim[2] = {64, (cuuint64_t)(img.size() / 64)};
    const cuuint64_t gstr[1] = {128};
    const cuuint32_t box[2] = {64, 128}, estr[2] = {1, 1};  // one box = one 16 KB tile
    CUresult cr = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, dImg, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)cr); return 1; }
    const size_t smem = (size_t)8 * TILE_BYTES + 2 * (BN / 8) * (K / 8) * 144 + 1024 + 128;
    cudaFuncSetAttribute(stream_test<4, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(stream_test<8, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(stream_test<8, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(stream_test<8, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int b_mn = 0; b_mn < 2; ++b_mn) {
        cudaMemset(dD, 0, D.size() * 4); cudaMemset(dS, 0, 4);
        stream_test<4, 64><<<1, 192, smem>>>(tmap, dX, dD, b_mn, 3, 1, 0, dS, dC);
        cudaError_t e = cudaDeviceSynchronize();
        int st = 0;
        cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost);
        double err = 0, mag = 0;
        for (size_t i = 0; i < D.size(); ++i) { err = fmax(err, fabs(D[i] - ref[i])); mag = fmax(mag, fabs(ref[i])); }
        printf("K=%d B %s-major: cuda=%s status=%d max|D-ref|=%.3g (max|ref|=%.3g) D[0]=%g ref=%g D[last]=%g ref=%g\n", K,
               b_mn ? "MN" : "K", cudaGetErrorString(e), st, err, mag, D[0], ref[0], D.back(), ref.back());
        if (e != cudaSuccess) return 2;
    }
    const char* names[] = {"ring NS=4", "ring NS=8", "ring NS=8 + epilogue per pass", "MMAs only N=64", "MMAs only N=32", "MMAs only N=16",
                           "TMA only NS=4", "TMA only NS=8", "ring NS=8 B K-major"};
    for (int cfg = 0; cfg < 9; ++cfg)
        for (int grid : {1, 148}) {
            const int iters = 200;
            const int sync_epi = cfg == 2 ? 1 : 0;
            const int mode = (cfg >= 3 && cfg <= 5) ? 1 : (cfg >= 6 && cfg <= 7 ? 2 : 0);
            const int b_mn = cfg == 8 ? 0 : 1;
            if (cfg == 0 || cfg == 6) stream_test<4, 64><<<grid, 192, smem>>>(tmap, dX, dD, b_mn, iters, sync_epi, mode, dS, dC);
            else if (cfg == 4) stream_test<8, 32><<<grid, 192, smem>>>(tmap, dX, dD, b_mn, iters, sync_epi, mode, dS, dC);
            else if (cfg == 5) stream_test<8, 16><<<grid, 192, smem>>>(tmap, dX, dD, b_mn, iters, sync_epi, mode, dS, dC);
            else stream_test<8, 64><<<grid, 192, smem>>>(tmap, dX, dD, b_mn, iters, sync_epi, mode, dS, dC);
            cudaError_t e = cudaDeviceSynchronize();
            std::vector<long long> c(grid);
            cudaMemcpy(c.data(), dC, grid * 8, cudaMemcpyDeviceToHost);
            long long mx = 0; double mean = 0;
            for (auto v : c) { mx = v > mx ? v : mx; mean += (double)v / grid; }
            const double bytes = (double)iters * T * TILE_BYTES;
            printf("%-30s grid %3d: %s  %.0f cycles per pass (max %.0f) = %d tiles of 16 KB, %d MMAs; %.1f B/cycle/SM, %.1f cycles per MMA\n",
                   names[cfg], grid, cudaGetErrorString(e), mean / iters, (double)mx / iters, T, T / 2 * 12, bytes / mean,
                   mean / iters / (T / 2 * 12));
        }
    cudaFree(dImg); cudaFree(dX); cudaFree(dD); cudaFree(dS); cudaFree(dC);
    return 0;
}
