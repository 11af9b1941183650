// Developer bring-up test (not part of the product): the MLP stage of a 16-tree group on tcgen05 in the
// TRANSPOSED formulation D^T[feature][tree] = W[feature][k] * X^T:
//   A = weights, resident in TMEM (lane = output feature, column = k), written once with tcgen05.st;
//   B = activations of the 16 trees in shared memory, K-major no-swizzle core matrices with a padded K pitch
//       (LBO = 144 bytes: conflict-free epilogue stores);
//   3xTF32 (hi*hi + lo*hi + hi*lo) accumulated in one TMEM accumulator of 16 columns;
//   epilogue: tcgen05.ld 32x32b (thread = feature, registers = trees), bias + ELU, hi / lo restaged as the next B.
// Checks the numerics of a 5-stage chain against float64 and times a stage.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/tc_stage_test tools/tc_stage_test.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

constexpr int NT = 16;        // trees per group == MMA N
constexpr int KMAX = 64;
constexpr int LBO = 144;      // bytes between core matrices along K (128 data + 16 pad)
constexpr int NSTAGE = 5;
constexpr int WCOLS = 64;     // K columns of one weight set (hi), followed by 64 lo columns
constexpr int NSETS = 3;
constexpr int D_COL = NSETS * 2 * WCOLS;  // accumulators behind the weight sets

__host__ __device__ inline int b_off(int n, int k, int K) {  // byte offset of X[n][k] inside a B buffer
    return (n >> 3) * ((K >> 2) * LBO) + (k >> 2) * LBO + (n & 7) * 16 + (k & 3) * 4;
}
constexpr int BUF_BYTES = 2 * (KMAX / 4) * LBO;  // one buffer (hi or lo) of K = 64

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)(lbo_bytes >> 4) << 16;
    d |= (uint64_t)(sbo_bytes >> 4) << 32;
    d |= 1ull << 46;
    return d;
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d), "r"(a), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ float tf32_lo(float x) { return x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }
__device__ __forceinline__ float elu_fast(float v) {
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fminf(v, 0.0f) * 1.4426950408889634f));
    return v > 0.0f ? v : e - 1.0f;
}

// REP = 4: the 32 output features are replicated in the four lane quarters, warp q handles trees 4q..4q+3
// REP = 1: warp 0 handles all 16 trees
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
    return pred != 0;
}

template <int REP, int K, int SPLIT>
__global__ void __launch_bounds__(128) stage_chain(const float* __restrict__ timg, const float* __restrict__ X0,
                                                   const float* __restrict__ bias, float* __restrict__ out,
                                                   int iters, long long* cycles, int* status, int n_mma = NT) {
    __shared__ __align__(128) unsigned char bufs[2][2][BUF_BYTES];  // [buffer][hi / lo]
    __shared__ __align__(8) unsigned long long mbar;
    __shared__ uint32_t tmem_base_sh;
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);  // provably warp-uniform: the MMA issue stays on the uniform datapath
    long long ph[4] = {0, 0, 0, 0};
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&mbar)), "n"(SPLIT));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_sh)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    for (int i = tid; i < 2 * 2 * BUF_BYTES / 4; i += 128) reinterpret_cast<float*>(bufs)[i] = 0.0f;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_sh;
    const uint32_t lane_base = (uint32_t)(32 * warp) << 16;
    // weights -> TMEM: thread = lane (output feature), 8 columns per store
    for (int c0 = 0; c0 < NSETS * 2 * WCOLS; c0 += 8) {
        uint32_t r[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) r[i] = __float_as_uint(timg[(size_t)(c0 + i) * 128 + tid]);
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(tmem_base + lane_base + c0),
                     "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                     : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    for (int i = tid; i < NT * K; i += 128) {
        const int n = i / K, k = i % K;
        const float x = X0[i];
        *reinterpret_cast<float*>(&bufs[0][0][b_off(n, k, K)]) = x;
        *reinterpret_cast<float*>(&bufs[0][1][b_off(n, k, K)]) = tf32_lo(x);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n_mma >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);  // n_mma > 16: timing only
    const uint32_t sbo = (uint32_t)(K >> 2) * LBO;
    uint32_t parity = 0;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        for (int s = 0; s < NSTAGE; ++s) {
            const int cur = s & 1, nxt = cur ^ 1;
            const uint32_t d_tmem = tmem_base + D_COL + 16 * (s & 1);  // split accumulators at + 32 * warp
            const long long q0 = clock64();
            if (warp < SPLIT) {  // SPLIT issuing warps: each takes a share of the K steps into its own accumulator
                const uint32_t a_hi = tmem_base + (s % NSETS) * 2 * WCOLS, a_lo = a_hi + WCOLS;
                const uint32_t bh = smem_u32(&bufs[cur][0][0]), bl = smem_u32(&bufs[cur][1][0]);
                const uint32_t dd = d_tmem + 32 * warp;
                if (elect_one()) {
#pragma unroll
                    for (int q = 0; q < K / 8 / SPLIT; ++q) {
                        const int ks = warp * (K / 8 / SPLIT) + q;
                        const uint64_t dh = make_desc(bh + ks * 2 * LBO, LBO, sbo), dl = make_desc(bl + ks * 2 * LBO, LBO, sbo);
                        mma_ts(dd, a_hi + 8 * ks, dh, idesc, q > 0 ? 1u : 0u);
                        mma_ts(dd, a_lo + 8 * ks, dh, idesc, 1u);
                        mma_ts(dd, a_hi + 8 * ks, dl, idesc, 1u);
                    }
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
                }
                __syncwarp();
            }
            const long long q1 = clock64();
            {
                uint32_t done = 0;
                for (int spin = 0; spin < (1 << 22) && !done; ++spin) {
                    asm volatile(
                        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                        : "=r"(done) : "r"(smem_u32(&mbar)), "r"(parity) : "memory");
                }
                if (!done) {
                    if (tid == 0) *status = 1 + s;
                    break;
                }
                parity ^= 1u;
            }
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const long long q2 = clock64();
            const float bv = bias[s * 128 + tid];
            if (REP == 4) {
                uint32_t v[4];
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                             : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(d_tmem + lane_base + 4 * warp));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                for (int sp = 1; sp < SPLIT; ++sp) {
                    uint32_t w[4];
                    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                                 : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "r"(d_tmem + 32 * sp + lane_base + 4 * warp));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    for (int i = 0; i < 4; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) + __uint_as_float(w[i]));
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int n = 4 * warp + i;
                    const float y = elu_fast(__uint_as_float(v[i]) + bv);
                    const float yl = tf32_lo(y);
                    if (s == NSTAGE - 1 && it == 0) out[n * 32 + lane] = y;
                    *reinterpret_cast<float*>(&bufs[nxt][0][b_off(n, lane, K)]) = y;
                    *reinterpret_cast<float*>(&bufs[nxt][1][b_off(n, lane, K)]) = yl;
                    if (K == 64) {
                        *reinterpret_cast<float*>(&bufs[nxt][0][b_off(n, lane + 32, K)]) = y;
                        *reinterpret_cast<float*>(&bufs[nxt][1][b_off(n, lane + 32, K)]) = yl;
                    }
                }
            } else if (warp == 0) {
                uint32_t v[16];
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                      "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                    : "r"(d_tmem + lane_base));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int n = 0; n < 16; ++n) {
                    const float y = elu_fast(__uint_as_float(v[n]) + bv);
                    const float yl = tf32_lo(y);
                    if (s == NSTAGE - 1 && it == 0) out[n * 32 + lane] = y;
                    *reinterpret_cast<float*>(&bufs[nxt][0][b_off(n, lane, K)]) = y;
                    *reinterpret_cast<float*>(&bufs[nxt][1][b_off(n, lane, K)]) = yl;
                    if (K == 64) {
                        *reinterpret_cast<float*>(&bufs[nxt][0][b_off(n, lane + 32, K)]) = y;
                        *reinterpret_cast<float*>(&bufs[nxt][1][b_off(n, lane + 32, K)]) = yl;
                    }
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            const long long q3 = clock64();
            __syncthreads();
            const long long q4 = clock64();
            ph[0] += q1 - q0; ph[1] += q2 - q1; ph[2] += q3 - q2; ph[3] += q4 - q3;
        }
    }
    long long t1 = clock64();
    if (tid == 0) {
        cycles[0] = t1 - t0;
        for (int i = 0; i < 4; ++i) cycles[1 + i] = ph[i];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
}

static float tf32_lo_h(float x) {
    uint32_t u;
    memcpy(&u, &x, 4);
    u &= 0xFFFFE000u;
    float t;
    memcpy(&t, &u, 4);
    return x - t;
}

int main(int argc, char** argv) {
    const int n_mma = argc > 1 ? atoi(argv[1]) : NT;
    printf("MMA N = %d%s\n", n_mma, n_mma != NT ? " (timing only: results are not meaningful)" : "");
    for (int K : {64, 32}) {
        for (int rep : {4, 2, 3, 1}) {  // 2 / 3: rep 4 with the issue split over 2 / 4 warps
            // weights W[s][32][K]; stage s uses set s % NSETS, so stages 3, 4 reuse the weights of 0, 1
            std::vector<float> W(NSETS * 32 * K), X0(NT * K), bias(NSTAGE * 128), out(NT * 32);
            srand(7);
            for (auto& x : W) x = (rand() % 2001 - 1000) / 1000.0f * (1.5f / sqrtf((float)K));
            for (auto& x : X0) x = (rand() % 2001 - 1000) / 1000.0f;
            for (int s = 0; s < NSTAGE; ++s)
                for (int m = 0; m < 128; ++m) bias[s * 128 + m] = ((s * 131 + (m % 32) * 17) % 200 - 100) / 500.0f;
            // TMEM image: [column][lane]; set q: 64 hi columns then 64 lo columns; feature = lane % 32 (replicated)
            std::vector<float> timg((size_t)NSETS * 2 * WCOLS * 128, 0.0f);
            for (int q = 0; q < NSETS; ++q)
                for (int k = 0; k < K; ++k)
                    for (int l = 0; l < 128; ++l) {
                        const float w = W[(q * 32 + (l % 32)) * K + k];
                        timg[(size_t)(q * 2 * WCOLS + k) * 128 + l] = w;
                        timg[(size_t)(q * 2 * WCOLS + WCOLS + k) * 128 + l] = tf32_lo_h(w);
                    }
            // float64 reference of the chain
            std::vector<double> x(NT * K), y(NT * K);
            for (int i = 0; i < NT * K; ++i) x[i] = X0[i];
            std::vector<double> ref(NT * 32);
            for (int s = 0; s < NSTAGE; ++s) {
                for (int n = 0; n < NT; ++n)
                    for (int m = 0; m < 32; ++m) {
                        double a = bias[s * 128 + m];
                        for (int k = 0; k < K; ++k) a += (double)W[((s % NSETS) * 32 + m) * K + k] * x[n * K + k];
                        a = a > 0 ? a : expm1(a);
                        a = (double)(float)a;  // the device stores float32 activations
                        y[n * K + m] = a;
                        if (K == 64) y[n * K + m + 32] = a;
                        if (s == NSTAGE - 1) ref[n * 32 + m] = a;
                    }
                x = y;
            }
            float *dT, *dX, *dB, *dO;
            long long* dC;
            int* dS;
            cudaMalloc(&dT, timg.size() * 4); cudaMalloc(&dX, X0.size() * 4); cudaMalloc(&dB, bias.size() * 4);
            cudaMalloc(&dO, out.size() * 4); cudaMalloc(&dC, 40); cudaMalloc(&dS, 4);
            cudaMemcpy(dT, timg.data(), timg.size() * 4, cudaMemcpyHostToDevice);
            cudaMemcpy(dX, X0.data(), X0.size() * 4, cudaMemcpyHostToDevice);
            cudaMemcpy(dB, bias.data(), bias.size() * 4, cudaMemcpyHostToDevice);
            cudaMemset(dO, 0, out.size() * 4); cudaMemset(dS, 0, 4);
            for (int iters : {1, 200}) {
                if (rep == 4 && K == 64) stage_chain<4, 64, 1><<<1, 128>>>(dT, dX, dB, dO, iters, dC, dS, n_mma);
                else if (rep == 4) stage_chain<4, 32, 1><<<1, 128>>>(dT, dX, dB, dO, iters, dC, dS, n_mma);
                else if (rep == 2 && K == 64) stage_chain<4, 64, 2><<<1, 128>>>(dT, dX, dB, dO, iters, dC, dS, n_mma);
                else if (rep == 2) stage_chain<4, 32, 2><<<1, 128>>>(dT, dX, dB, dO, iters, dC, dS, n_mma);
                else if (rep == 3 && K == 64) stage_chain<4, 64, 4><<<1, 128>>>(dT, dX, dB, dO, iters, dC, dS, n_mma);
                else if (rep == 3) stage_chain<4, 32, 4><<<1, 128>>>(dT, dX, dB, dO, iters, dC, dS, n_mma);
                else if (K == 64) stage_chain<1, 64, 1><<<1, 128>>>(dT, dX, dB, dO, iters, dC, dS, n_mma);
                else stage_chain<1, 32, 1><<<1, 128>>>(dT, dX, dB, dO, iters, dC, dS, n_mma);
                cudaError_t e1 = cudaGetLastError();
                cudaError_t e = cudaDeviceSynchronize();
                long long cy[5] = {0, 0, 0, 0, 0};
                int st = 0;
                cudaMemcpy(cy, dC, 40, cudaMemcpyDeviceToHost);
                const long long cyc = cy[0];
                cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost);
                cudaMemcpy(out.data(), dO, out.size() * 4, cudaMemcpyDeviceToHost);
                double err = 0, mag = 0;
                for (size_t i = 0; i < out.size(); ++i) {
                    err = fmax(err, fabs(out[i] - ref[i]));
                    mag = fmax(mag, fabs(ref[i]));
                }
                printf("K=%d rep=%d iters=%d launch=%s sync=%s status=%d max|out-ref|=%.3g (max|ref|=%.3g) out[0]=%g ref[0]=%g cycles/stage=%.1f (issue %.0f, mma wait %.0f, epilogue %.0f, barrier %.0f)\n",
                       K, rep, iters, cudaGetErrorString(e1), cudaGetErrorString(e), st, err, mag, out[0], ref[0],
                       (double)cyc / (iters * NSTAGE), (double)cy[1] / (iters * NSTAGE), (double)cy[2] / (iters * NSTAGE),
                       (double)cy[3] / (iters * NSTAGE), (double)cy[4] / (iters * NSTAGE));
            }
            cudaFree(dT); cudaFree(dX); cudaFree(dB); cudaFree(dO); cudaFree(dC); cudaFree(dS);
        }
    }
    return 0;
}
