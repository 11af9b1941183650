// tsp_tc.cuh -- the search MLP of a 16-tree group on the 5th-generation tensor cores (tcgen05 + TMEM), used by
// search_fast_kernel<..., TC = true> for the fully-connected MuZero networks whose search weights fit TMEM
// (highway / roundabout: models.py:136-288 with one hidden layer per MLP).
//
// TRANSPOSED formulation: a Linear layer is D^T[feature][tree] = W[feature][k] * X^T, so that the 16 trees of a
// group are the MMA's N dimension (N = 16 is legal; M = 128 trees is not available at 28-32 trees per SM):
//   * A = the weights, resident in TENSOR MEMORY for the whole search (lane = output feature, column = k; written
//     once per CTA with tcgen05.st from a host-built image). Layers that share K sit side by side in the lanes;
//     small layers are REPLICATED across the four lane quarters so that the four epilogue warps of a group split
//     the trees instead of one warp owning all 16.
//   * B = the activations of the 16 trees in shared memory, K-major / no swizzle: element (n, k) at
//     (n / 8) * SBO + (k / 4) * LBO + (n % 8) * 16 + (k % 4) * 4 bytes with LBO = 144 (128 + 16 of padding: the
//     epilogue's stores of 32 consecutive features hit 32 different banks).
//   * fp32-level accuracy by the 3xTF32 split (hi*hi + lo*hi + hi*lo into one fp32 TMEM accumulator): the tensor
//     core ignores the 13 low mantissa bits of a kind::tf32 operand, so the raw float32 word is the hi operand;
//     lo = x - trunc(x) is stored next to it (weights: second half of the TMEM image; activations: twin buffer).
//   * One elected thread of a warp that has no epilogue share in the stage issues the stage's MMAs and commits them
//     to the group's mbarrier; all warps of the group run the epilogue (warp w reads TMEM lane quarter w % 4; with 8 warps per group
//     two warps split the trees of a quarter): tcgen05.ld 32x32b
//     (thread = feature, registers = trees), bias (+ the gathered one-hot action row of models.py:236-243), ELU,
//     and the restaging of hi / lo as the next layer's B operand.
// Measured on B200 (tools/tc_stage_test.cu): a K = 64 stage (24 MMAs) 800 cycles, a K = 32 stage 590, of which the
// single-thread issue is 12 cycles per MMA -- against ~2,400 cycles per stage of the mma.sync unit lists.
#pragma once
#include <cstdint>

namespace tsp {

constexpr int TC_LBO = 144;        // bytes between core matrices along K
constexpr int TC_MAX_MMA = 8;
constexpr int TC_MAX_STAGES = 6;
constexpr int TC_MAX_PASS = 6;
constexpr int TC_MAX_CTAB = 1024;  // floats: biases and one-hot rows
constexpr int TC_TMEM_COLS = 512;

struct TcMma {
    int a_col;    // TMEM column of the hi weights of this layer group (lo at + a_lo)
    int b_off;    // byte offset of the B operand (hi) inside the group's arena (lo at + b_lo)
    int ksteps;   // K / 8
    int sbo;      // bytes between 8-tree groups of the B operand
    int d_col;    // accumulator column, relative to the group's accumulator base
    int n;        // MMA columns: 16 trees (0 reads as 16), or 32 = both futures of a transition expansion side by side
    int pad[2];
};
struct TcEpi {        // one warp (TMEM lane quarter) of one epilogue pass
    short d_col;      // accumulator column (group-relative) of tree 0
    short n0, ncnt;   // columns handled by this warp: n0 .. n0 + ncnt - 1, ncnt in {0, 2, 4, 8, 16, 32}; column n = tree n % 16
                      // (columns 16..31: the second future of a merged transition expansion)
    short rows;       // valid lanes: lane < rows holds feature feat0 + lane
    int dst;          // B destination: byte offset of the hi buffer in the arena; RAW: column of lane 0 in the out strip
    int dst_sbo;      // (RAW: out-strip column of lane 0 for the columns 16..31)
    int feat0;        // B destination: k of lane 0
    int cidx;         // ctab offset of lane 0's bias
    int eidx;         // ctab offset of the one-hot rows (row a at eidx + a * e_ld), -1: none
    int e_ld;
    int zidx;         // ctab offset of the one-hot FUTURE rows of MuZeroStochasticConcat (models.py:568-580; row z at zidx + z * e_ld), -1: none
    int flags;        // 1: ELU, 2: RAW (float32 logits to the out strip), 4: BOTH futures from one accumulator (rows zidx and
                      // zidx + e_ld; the second future's operand 16 columns further), 8: no lo twin (no MMA reads this buffer)
};
static_assert(sizeof(TcEpi) == 40, "TcEpi: four shorts and eight ints");
struct TcStage {
    int mma0, n_mma;  // MMAs issued at the start of the stage
    int commit;       // commit them (and everything issued before) to the group's mbarrier
    int row_op;       // 1: min-max normalise xs -> xn (models.py:248-255) while the MMAs run
    int epi0, n_epi;  // epilogue passes after the mbarrier wait (a lane quarter with two accumulators to read: two passes)
    int issuer;       // warp of the group that issues the MMAs: one WITHOUT a share of this stage's epilogue, so that the
                      // epilogue warps fetch their operands while the MMAs are issued instead of after
    int pad;
};
struct TcSched {
    int n_stages, n_cols, a_lo, b_lo;
    int d_col0, d_cols_group, arena_bytes, out_pitch;
    int off_vl, off_rl, off_pl, ctab_floats;
    int xh_off, xh_sbo, xs_off, xs_sbo;
    int xn_off, xn_sbo, H, off_tp;
    int merged, off_rl2;  // chance-node searches: both futures in one pass (columns 16..31 = future 1); its reward logits
    TcStage stage[TC_MAX_STAGES];
    TcMma mma[TC_MAX_MMA];
    TcEpi epi[TC_MAX_PASS][8];  // [pass][warp of the group]: warps w and w + 4 share TMEM lane quarter w % 4
};
static_assert(sizeof(TcSched) <= 3200, "TcSched travels as a kernel parameter (constant bank: uniform reads)");

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
    return pred != 0;
}
// K-major, no swizzle, descriptor version 1 (sm_100): low word = start address >> 4 | LBO >> 4 << 16, high word =
// SBO >> 4 | 1 << 14
__device__ __forceinline__ uint32_t tc_desc_lo(uint32_t saddr) { return ((saddr & 0x3FFFF) >> 4) | ((uint32_t)(TC_LBO >> 4) << 16); }
__device__ __forceinline__ uint32_t tc_desc_hi(uint32_t sbo_bytes) { return (sbo_bytes >> 4) | (1u << 14); }
// D[tmem] (+)= A[tmem] * B[smem descriptor], kind::tf32, M = 128, N = 16, K = 8. The descriptor travels as two 32-bit
// words: a K step only moves the start-address field of the low word (no 64-bit carry chain on the uniform datapath).
__device__ __forceinline__ void tc_mma(uint32_t d, uint32_t a, uint32_t bdesc_lo, uint32_t bdesc_hi, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 bd;\n\tsetp.ne.b32 p, %5, 0;\n\tmov.b64 bd, {%2, %3};\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], bd, %4, p;\n\t}\n" ::"r"(d), "r"(a), "r"(bdesc_lo), "r"(bdesc_hi),
        "r"(idesc), "r"(acc)
        : "memory");
}
constexpr uint32_t TC_IDESC0 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(128 >> 4) << 24);
__device__ __forceinline__ uint32_t tc_idesc(int n) { return TC_IDESC0 | ((uint32_t)(n >> 3) << 17); }

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Bounded wait (a lost commit must not hang the GPU): returns false on time-out.
__device__ __forceinline__ bool mbar_wait(uint32_t mbar, uint32_t parity) {
    uint32_t done = 0;
    for (int spin = 0; spin < (1 << 24) && !done; ++spin)
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
            : "=r"(done) : "r"(mbar), "r"(parity) : "memory");
    return done != 0;
}

template <int N>
__device__ __forceinline__ void tmem_ld(uint32_t taddr, uint32_t (&v)[N]);
template <>
__device__ __forceinline__ void tmem_ld<2>(uint32_t taddr, uint32_t (&v)[2]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];" : "=r"(v[0]), "=r"(v[1]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
template <>
__device__ __forceinline__ void tmem_ld<4>(uint32_t taddr, uint32_t (&v)[4]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
template <>
__device__ __forceinline__ void tmem_ld<8>(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
template <>
__device__ __forceinline__ void tmem_ld<16>(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// One warp's share of an epilogue pass: NCNT columns n0 .. n0 + NCNT - 1 (n0 a multiple of NCNT) of the lanes' features.
// Everything that does not depend on the accumulator (bias, the trees' one-hot rows, addresses) is fetched BEFORE the
// wait on the MMAs' mbarrier; all per-tree addresses are the pass's base plus compile-time offsets.
// MODE 0: MuZero; 1: chance-node searches, one future per pass (the future's one-hot row joins the bias); 2: chance-node searches,
// both futures side by side (columns 16..31, two passes per stage, operands without a lo twin). What a mode does not need is compiled out.
template <int NCNT, int MODE>
__device__ __forceinline__ bool tc_epi_warp(const TcEpi& ep, int n0, uint32_t d_tmem, unsigned char* __restrict__ arena, int b_lo,
                                            float* __restrict__ outs, int out_pitch, const float* __restrict__ ctab,
                                            const int* __restrict__ row_action, int lane, uint32_t mbar, uint32_t parity, int z,
                                            bool wait) {
    const bool on = lane < ep.rows;
    const int flags = ep.flags;
    const int t0 = n0 & 15;  // first tree
    float add[NCNT];
    float z0 = 0.0f, z1 = 0.0f;
    {
        float bias = on ? ctab[ep.cidx + lane] : 0.0f;
        if (MODE >= 1 && ep.zidx >= 0 && on) {
            if (MODE == 2 && (flags & 4)) {
                z0 = ctab[ep.zidx + lane];
                z1 = ctab[ep.zidx + ep.e_ld + lane];
            } else {
                bias += ctab[ep.zidx + z * ep.e_ld + lane];  // the same future for every tree of the pass
            }
        }
#pragma unroll
        for (int i = 0; i < NCNT; ++i) add[i] = bias;
        if (ep.eidx >= 0 && on) {  // one-hot action input == one gathered weight row per tree (models.py:236-243)
            const float* erow = ctab + ep.eidx + lane;
            const int e_ld = ep.e_ld;
            int act[NCNT];
            if constexpr (NCNT >= 4) {
#pragma unroll
                for (int i = 0; i < NCNT; i += 4) {
                    const int4 a4 = *reinterpret_cast<const int4*>(row_action + t0 + i);
                    act[i] = a4.x; act[i + 1] = a4.y; act[i + 2] = a4.z; act[i + 3] = a4.w;
                }
            } else {
                const int2 a2 = *reinterpret_cast<const int2*>(row_action + t0);
                act[0] = a2.x; act[1] = a2.y;
            }
#pragma unroll
            for (int i = 0; i < NCNT; ++i) add[i] += erow[act[i] * e_ld];
        }
    }
    const int k = ep.feat0 + lane;
    unsigned char* hi = (flags & 2) ? reinterpret_cast<unsigned char*>(outs + t0 * out_pitch + ((MODE == 2 && n0 >= 16) ? ep.dst_sbo : ep.dst) + lane)
                                    : arena + ep.dst + (k >> 2) * TC_LBO + (k & 3) * 4 + (n0 >> 3) * ep.dst_sbo + (n0 & 7) * 16;
    const int sbo = ep.dst_sbo;
    const uint32_t taddr = d_tmem + (uint32_t)(ep.d_col + n0);
    // ---- the accumulator
    bool ok = true;
    if (MODE != 2 || wait) {
        ok = mbar_wait(mbar, parity);
        tc_fence_after();
    }
    uint32_t v[NCNT];
    tmem_ld<NCNT>(taddr, v);
    if (!on) return ok;
    float y[NCNT];
#pragma unroll
    for (int i = 0; i < NCNT; ++i) y[i] = __uint_as_float(v[i]) + add[i];
    if (flags & 2) {  // raw logits for decode_row: out strip [tree][column]
        float* o = reinterpret_cast<float*>(hi);
#pragma unroll
        for (int i = 0; i < NCNT; ++i) {
            *o = y[i];
            o += out_pitch;
        }
        return ok;
    }
    if (MODE == 2 && (flags & 4)) {  // the hidden layer of the dynamics for both futures: they differ by one gathered weight row
        unsigned char* lo = hi + b_lo;
#pragma unroll
        for (int f = 0; f < 2; ++f) {
            const float zr = f ? z1 : z0;
#pragma unroll
            for (int i = 0; i < NCNT; ++i) {
                const float e = elu_fast(y[i] + zr);
                const int off = (NCNT == 16 && i >= 8) ? sbo + 16 * (i - 8) : 16 * i;
                *reinterpret_cast<float*>(hi + off) = e;
                *reinterpret_cast<float*>(lo + off) = tf32_lo(e);
            }
            hi += 2 * sbo;
            lo += 2 * sbo;
        }
        return ok;
    }
    if (flags & 1) {
#pragma unroll
        for (int i = 0; i < NCNT; ++i) y[i] = elu_fast(y[i]);
    }
    if (MODE == 2 && (flags & 8)) {
#pragma unroll
        for (int i = 0; i < NCNT; ++i) *reinterpret_cast<float*>(hi + ((NCNT == 16 && i >= 8) ? sbo + 16 * (i - 8) : 16 * i)) = y[i];
        return ok;
    }
    unsigned char* lo = hi + b_lo;
#pragma unroll
    for (int i = 0; i < (NCNT < 8 ? NCNT : 8); ++i) {
        *reinterpret_cast<float*>(hi + 16 * i) = y[i];
        *reinterpret_cast<float*>(lo + 16 * i) = tf32_lo(y[i]);
    }
    if constexpr (NCNT == 16) {
        hi += sbo;
        lo += sbo;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            *reinterpret_cast<float*>(hi + 16 * i) = y[8 + i];
            *reinterpret_cast<float*>(lo + 16 * i) = tf32_lo(y[8 + i]);
        }
    }
    return ok;
}

// The 3 * KS MMAs of one layer (3xTF32), fully unrolled: one thread issues them.
template <int KS>
__device__ __forceinline__ void tc_issue_layer(uint32_t d, uint32_t a_hi, uint32_t a_lo, uint32_t bh_lo, uint32_t bl_lo, uint32_t b_hi,
                                               int ksteps, uint32_t idesc) {
#pragma unroll
    for (int ks = 0; ks < (KS ? KS : 1); ++ks) {
        if (KS == 0 && ks == 0) {
            for (int q = 0; q < ksteps; ++q) {
                tc_mma(d, a_hi + 8 * q, bh_lo + q * ((2 * TC_LBO) >> 4), b_hi, idesc, q > 0 ? 1u : 0u);
                tc_mma(d, a_lo + 8 * q, bh_lo + q * ((2 * TC_LBO) >> 4), b_hi, idesc, 1u);
                tc_mma(d, a_hi + 8 * q, bl_lo + q * ((2 * TC_LBO) >> 4), b_hi, idesc, 1u);
            }
        } else if (KS) {
            tc_mma(d, a_hi + 8 * ks, bh_lo + ks * ((2 * TC_LBO) >> 4), b_hi, idesc, ks > 0 ? 1u : 0u);
            tc_mma(d, a_lo + 8 * ks, bh_lo + ks * ((2 * TC_LBO) >> 4), b_hi, idesc, 1u);
            tc_mma(d, a_hi + 8 * ks, bl_lo + ks * ((2 * TC_LBO) >> 4), b_hi, idesc, 1u);
        }
    }
}

// All stages of the recurrent inference for the 16 trees of one group (models.py:233-257, 283-287); called by every
// thread of the group. gw: warp of the group (provably warp-uniform); hq: the normalised hidden state quads of this
// lane's tree (quad i = j + q * LPT), returned for the expansion's hidden-state store.
template <int LPT, int MODE = 0>
__device__ __forceinline__ bool tc_mlp(const TcSched& ts, uint32_t tmem_base, uint32_t d_group, unsigned char* __restrict__ arena,
                                       float* __restrict__ outs, const float* __restrict__ ctab,
                                       const int* __restrict__ row_action, uint32_t mbar, uint32_t& parity, int gw, int lane,
                                       int row, int j, int barid, float4 (&hq)[2], unsigned long long* dbg = nullptr,
                                       int s_begin = 0, int s_end = -1, int z = 0, float* __restrict__ hst = nullptr) {
    constexpr int GTHREADS = GT * LPT;
    const uint32_t arena_s = smem_u32(arena);
    const uint32_t lane_q = (uint32_t)(32 * (gw & 3)) << 16;
    bool ok = true;
    const int n_stages = s_end < 0 ? ts.n_stages : s_end;
    for (int s = s_begin; s < n_stages; ++s) {
        const TcStage& st = ts.stage[s];
        const bool tm = dbg && threadIdx.x == 0;
        long long q0 = 0, q1 = 0, q2 = 0, q3 = 0, q4 = 0;
        if (tm) q0 = clock64();
        if (gw == st.issuer) {
            tc_fence_after();
            if (elect_one()) {
                for (int m = st.mma0; m < st.mma0 + st.n_mma; ++m) {
                    const TcMma& mm = ts.mma[m];
                    const uint32_t a_hi = tmem_base + mm.a_col, a_lo = a_hi + ts.a_lo;
                    const uint32_t bh_lo = tc_desc_lo(arena_s + mm.b_off), bl_lo = tc_desc_lo(arena_s + mm.b_off + ts.b_lo);
                    const uint32_t b_hi = tc_desc_hi(mm.sbo);
                    const uint32_t d = d_group + mm.d_col;
                    const uint32_t idesc = tc_idesc((MODE == 2 && mm.n) ? mm.n : 16);
                    if (mm.ksteps == 8) tc_issue_layer<8>(d, a_hi, a_lo, bh_lo, bl_lo, b_hi, 8, idesc);
                    else if (mm.ksteps == 4) tc_issue_layer<4>(d, a_hi, a_lo, bh_lo, bl_lo, b_hi, 4, idesc);
                    else tc_issue_layer<0>(d, a_hi, a_lo, bh_lo, bl_lo, b_hi, mm.ksteps, idesc);
                }
                if (st.commit)
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar) : "memory");
            }
            __syncwarp();
        }
        if (tm) q1 = clock64();
        if (st.row_op) {
            // models.py:248-255: (x - min) / scale, scale += 1e-5 only where scale < 1e-5; the tree's lanes each own
            // quads of 4 consecutive features == one 16-byte row of a core matrix
            const int nq = ts.H >> 2;
            const int nz = MODE == 2 ? 2 : 1;  // merged transition expansion: future 1 of tree `row` is column row + 16
            for (int zz = 0; zz < nz; ++zz) {
                const int r2 = row + 16 * zz;
                const unsigned char* src = arena + ts.xs_off + (r2 >> 3) * ts.xs_sbo + (r2 & 7) * 16;
                float mn = CUDART_INF_F, mx = -CUDART_INF_F;
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int i = j + q * LPT;
                    if (i < nq) {
                        hq[q] = *reinterpret_cast<const float4*>(src + i * TC_LBO);
                        mn = fminf(mn, fminf(fminf(hq[q].x, hq[q].y), fminf(hq[q].z, hq[q].w)));
                        mx = fmaxf(mx, fmaxf(fmaxf(hq[q].x, hq[q].y), fmaxf(hq[q].z, hq[q].w)));
                    }
                }
#pragma unroll
                for (int d = 1; d < LPT; d <<= 1) {
                    mn = fminf(mn, __shfl_xor_sync(FULL, mn, d));
                    mx = fmaxf(mx, __shfl_xor_sync(FULL, mx, d));
                }
                float scale = __fsub_rn(mx, mn);
                if (scale < 1e-5f) scale = __fadd_rn(scale, 1e-5f);
                const float inv = __frcp_rn(scale);
                unsigned char* dst = arena + ts.xn_off + (row >> 3) * ts.xn_sbo + (row & 7) * 16;
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int i = j + q * LPT;
                    if (i < nq) {
                        float4 v = hq[q];
                        v.x = __fmul_rn(__fsub_rn(v.x, mn), inv); v.y = __fmul_rn(__fsub_rn(v.y, mn), inv);
                        v.z = __fmul_rn(__fsub_rn(v.z, mn), inv); v.w = __fmul_rn(__fsub_rn(v.w, mn), inv);
                        hq[q] = v;
                        if (MODE == 2) {  // straight into the hidden-state slots of the new futures
                            if (hst) reinterpret_cast<float4*>(hst + (size_t)zz * ts.H)[i] = v;
                        } else if (ts.xn_off >= 0) {  // (no layer of a transition expansion reads the normalised state)
                            *reinterpret_cast<float4*>(dst + i * TC_LBO) = v;
                            *reinterpret_cast<float4*>(dst + i * TC_LBO + ts.b_lo) = make_float4(tf32_lo(v.x), tf32_lo(v.y), tf32_lo(v.z), tf32_lo(v.w));
                        }
                    }
                }
            }
            fence_async_smem();
        }
        if (st.n_epi > 0) {  // every warp of the group waits on the stage's commit, once
            const uint32_t d_tmem = d_group + lane_q;
            const int n_epi = MODE == 2 ? st.n_epi : 1;
            for (int e = 0; e < n_epi; ++e) {
                const TcEpi& ep = ts.epi[st.epi0 + e][gw];
                const bool wt = e == 0;
                bool w = true;
                if (MODE != 2 && ep.ncnt == 2) w = tc_epi_warp<2, MODE>(ep, ep.n0, d_tmem, arena, ts.b_lo, outs, ts.out_pitch, ctab, row_action, lane, mbar, parity, z, wt);
                else if (MODE != 2 && ep.ncnt == 4) w = tc_epi_warp<4, MODE>(ep, ep.n0, d_tmem, arena, ts.b_lo, outs, ts.out_pitch, ctab, row_action, lane, mbar, parity, z, wt);
                else if (ep.ncnt == 8) w = tc_epi_warp<8, MODE>(ep, ep.n0, d_tmem, arena, ts.b_lo, outs, ts.out_pitch, ctab, row_action, lane, mbar, parity, z, wt);
                else if (ep.ncnt >= 16) {
                    const int reps = (MODE == 2 && ep.ncnt == 32) ? 2 : 1;  // 32 columns: two rounds of 16 through one copy of the code
#pragma unroll 1
                    for (int h = 0; h < reps; ++h)
                        w = tc_epi_warp<16, MODE>(ep, ep.n0 + 16 * h, d_tmem, arena, ts.b_lo, outs, ts.out_pitch, ctab, row_action, lane, mbar,
                                                  parity, z, wt && h == 0) && w;
                } else if (wt) {
                    w = mbar_wait(mbar, parity);
                    if (MODE == 2) tc_fence_after();
                }
                ok = ok && w;
            }
            parity ^= 1u;
            if (tm) q2 = clock64();
            fence_async_smem();
            tc_fence_before();
        }
        if (tm) q3 = clock64();
        group_barrier<GTHREADS>(barid);
        if (tm) {  // stage s: {issue, row op / mbarrier wait, epilogue, barrier}
            q4 = clock64();
            if (q2 == 0) q2 = q3;
            atomicAdd(dbg + 8 + 2 * s, (unsigned long long)(q3 - q0));
            atomicAdd(dbg + 9 + 2 * s, (unsigned long long)(q4 - q3));
            atomicAdd(dbg + 44 + 4 * s, (unsigned long long)(q1 - q0));
            atomicAdd(dbg + 45 + 4 * s, (unsigned long long)(q2 - q1));
            atomicAdd(dbg + 46 + 4 * s, (unsigned long long)(q3 - q2));
        }
    }
    return ok;
}

}  // namespace tsp
