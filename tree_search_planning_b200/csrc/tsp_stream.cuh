// tsp_stream.cuh -- the search MLP of 64 lock-stepped trees on tcgen05 with the weights STREAMED by TMA, used by
// search_tcs_kernel for fully-connected MuZero networks whose search weights fit neither tensor memory nor shared
// memory (cross-merge: games/cross_merge_env.py:112-118, H = 256, dynamics 261 -> 128 -> 256, heads 256 -> 32 -> n;
// models.py:209-287).
//
// TRANSPOSED formulation like tsp_tc.cuh -- D^T[feature][tree] = W[feature][k] X^T[k][tree] -- with
//   * N = 64: the 64 trees of the CTA are the MMA's N dimension and run the network evaluation in lock step;
//   * A = weight tiles of 16 KB (fp16: 128 rows x 64 inputs, 64 x 128 or 32 x 256 -- narrow layers get long tiles so that
//     the bytes in flight do not shrink with the layer), host-arranged in the K-major no-swizzle core-matrix order and
//     streamed from global memory / L2 through a ring of TS_NS shared-memory slots by TMA (cp.async.bulk.tensor.2d +
//     mbarrier complete_tx); one elected lane of a PRODUCER warp runs ahead of the MMAs by the depth of the ring, across
//     layer and simulation boundaries;
//   * B = the activations in shared memory: the gathered hidden state K-major (written by the trees' lanes, 16-byte
//     chunks of 8 inputs), every later operand MN-major (written by the epilogue, thread = feature, 16-byte chunks of 8
//     trees);
//   * float32-level accuracy by the SCALED fp16 split on kind::f16 (K = 16 per MMA, half the MMAs of 3xTF32):
//         v = v1 + v2 / 2048,  v1 = fp16(v),  v2 = fp16((v - v1) * 2048)          (22 significant bits)
//         acc0 += W1 X1,  acc1 += W1 X2 + W2 X1,  D = acc0 + acc1 / 2048          (fp32 accumulators in TMEM)
//     The X2 image of an operand follows its X1 image, so [X1 | X2] is ONE B operand of N = 128 columns and acc0 | acc1
//     one accumulator of 128 columns: a W1 tile issues one N = 128 MMA per 16 inputs, a W2 tile one N = 64 MMA --
//     two MMAs per 16 inputs instead of three, and the W1 tile is read from shared memory once;
//     the dropped term W2 X2 / 2048^2 is 2^-22 relative (tools/tcs_gemm_test.cu: 5.8e-7 of max |D| at K = 256);
//     activations beyond fp16's range (|x| > 6e4) raise the status flag instead of overflowing silently;
//   * one elected lane of an ISSUER warp issues the MMAs of a tile when its slot is full (and, at a layer boundary, when
//     all 16 tree warps have published the layer's B operand on an mbarrier), commits the slot back to the producer and
//     the finished accumulator tile to the epilogue;
//   * the 16 tree warps run the epilogues: warp w owns TMEM lane quarter w % 4 (32 features) and the trees
//     16 (w / 4) .. + 15: tcgen05.ld, bias (+ the gathered one-hot action row of models.py:236-243), ELU, fp16 split,
//     STS.128 of 8 trees. The next state is min-max normalised (models.py:248-255) in the accumulator's own layout:
//     per-tree minimum / maximum by REDUX + shared-memory atomics on order-preserving integer keys, then a second pass
//     over the accumulator writes the normalised state as the B operand of the prediction heads and straight into the
//     hidden-state pool (thread = feature: a warp stores 128 contiguous bytes per tree).
// Stages (cross-merge): dynamics 1 (8 tiles) -> dynamics 2 (+ the first reward layer folded in like tsp_tc.cuh: 10
// tiles) -> policy 1 | value 1 (4 tiles) -> the three output layers as one block-diagonal K = 96 layer (2 tiles):
// 24 tiles = 384 KB and 124 MMAs per simulation of 64 trees.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cstdint>

namespace tsp {

constexpr int TS_N = 64;          // trees per CTA == N of every MMA of the search kernel (the root kernel runs 32 rows per CTA)
constexpr int TS_NS = 4;          // ring slots (8 were measured: no gain -- the stream is bound by shared-memory bandwidth, not latency)
constexpr int TS_X0_LBO = 144;    // gathered state, K-major: bytes between core matrices along K (128 + 16 of padding: the
                                  // 16-byte stores of a tree's 8 lanes hit different banks; dense 128 cost 1k cycles per simulation)
constexpr int TS_SLOT = 16384;    // bytes per slot: 128 rows x 64 inputs fp16
constexpr int TS_MAX_TILES = 72;
constexpr int TS_MAX_ACC = 8;
constexpr int TS_STAGES = 4;      // stages of the recurrent inference (search kernel)
constexpr int TS_MAX_STAGES = 7;  // the root network: representation layers + 2
constexpr int TS_TMEM_COLS = 512;
constexpr int TS_MAX_CTAB = 2048;  // floats: biases and one-hot rows
constexpr float TS_SPLIT = 2048.0f, TS_INV_SPLIT = 1.0f / 2048.0f;
constexpr float TS_FP16_MAX = 6.0e4f;

struct TsTile {       // one weight tile: one ring slot fill and the MMAs that consume it
    int img_row;      // first 128-byte row of the tile in the weight image (TMA coordinate); a tile is 128 image rows
    short a_sbo;      // bytes between 8-row groups of the tile: 128 * (inputs of the tile / 8) = 1024 / 2048 / 4096
    short ksteps;     // valid inputs of the tile / 16, <= 16
    int b_off;        // B operand (X1 image) at the tile's first input: byte offset in the arena
    int b_lo;         // distance X1 -> X2 image of that operand: always 8 * b_sbo ([X1 | X2] is one operand of 128 columns)
    int b_sbo;        // bytes between 8-tree groups
    short b_lbo;      // bytes between core matrices along K: TS_X0_LBO (K-major) or 128 (MN-major)
    short b_step;     // descriptor units (16 bytes) per k-step
    short d_col;      // accumulator: acc0 at d_col, acc1 at d_col + nt
    char kind;        // 0: W1 tile (acc0 += W1 X1, acc1 += W1 X2), 1: W2 tile (acc1 += W2 X1)
    char first;       // first tile of its accumulator: k-step 0 overwrites
    char b_mn;        // B operand major: 0 K, 1 MN
    char acc_done;    // 1 + index of the accumulator tile this tile completes, 0: none
    short pad;
};
static_assert(sizeof(TsTile) == 32, "TsTile");
struct TsEpi {        // one accumulator tile (128 TMEM lanes x {acc0, acc1} x 64 trees) and what its rows become
    short d_col;
    short rows;       // valid rows: lane quarter q holds rows 32 q .. 32 q + 31
    short kind;       // 0: bias (+ one-hot action row) + ELU -> fp16 split -> B operand (MN-major)
                      // 1: tile of the next state: bias, min-max normalised -> B operand + hidden-state pool
                      // 2: raw float32 logits -> out strip [tree][column]
    short feat0;      // input index k (kind 0, 1: of the destination operand) or out-strip column (kind 2) of row 0
    int dst;          // destination operand (X1 image): byte offset in the arena; kind 2: unused
    int dst_sbo;
    int dst_lo;
    int cidx;         // ctab offset of row 0's bias
    int eidx;         // ctab offset of the one-hot action rows (row a at eidx + a * e_ld), -1: none
    int e_ld;
};
static_assert(sizeof(TsEpi) == 32, "TsEpi");
struct TsSched {
    int n_tiles, n_acc, arena_bytes, out_pitch;
    int off_vl, off_rl, off_pl, ctab_floats;
    int x0_off, x0_sbo, x0_lo, H;       // gathered hidden state: K-major, LBO TS_X0_LBO
    int stage_tile[TS_MAX_STAGES + 1];  // tiles [stage_tile[s], stage_tile[s + 1]) read the B operand published before stage s
    int stage_acc[TS_MAX_STAGES + 1];
    int img_rows, n_stages;
    int nt, in_dim;                     // columns of the MMAs (trees / rows per CTA: 64 or 32); root network: observation width
    TsTile tile[TS_MAX_TILES];
    TsEpi acc[TS_MAX_ACC];
};
static_assert(sizeof(TsSched) <= 3072, "TsSched travels as a kernel parameter (constant bank: uniform reads)");

// shared-memory control block of the streamed path
struct TsCtl {
    unsigned long long full[TS_NS], empty[TS_NS], acc[TS_MAX_ACC], bready;
    uint32_t tmem_slot;
    int abort;
    int mmk[2][TS_N][2];                 // [simulation parity][tree]{min key, max key} of the next state
    unsigned long long hdst[TS_N];       // where the hidden state of the tree's new node goes (0: nowhere)
    float2 mnv[16][16];                  // per tree warp: {minimum, 1 / scale} of its 16 trees' next states
};

__device__ __forceinline__ uint32_t ts_desc_lo(uint32_t saddr, uint32_t lbo) { return ((saddr & 0x3FFFF) >> 4) | ((lbo >> 4) << 16); }
__device__ __forceinline__ uint32_t ts_desc_hi(uint32_t sbo) { return (sbo >> 4) | (1u << 14); }
// D[tmem] (+)= A[smem descriptor] * B[smem descriptor], kind::f16 (fp16 operands, fp32 accumulate), M = 128, N = 64, K = 16
__device__ __forceinline__ void ts_mma(uint32_t d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 ad, bd;\n\tsetp.ne.b32 p, %6, 0;\n\tmov.b64 ad, {%1, %2};\n\tmov.b64 bd, {%3, %4};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], ad, bd, %5, p;\n\t}\n" ::"r"(d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(acc)
        : "memory");
}
// instruction descriptor: D f32, A / B f16, A K-major, M = 128, N columns
__device__ __forceinline__ uint32_t ts_idesc(int n) { return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24); }

// Bounded wait that gives up at once when another thread of the CTA has already timed out (a lost TMA / commit must
// neither hang the GPU nor stretch into minutes of serial time-outs).
__device__ __forceinline__ bool ts_wait(uint32_t mbar, uint32_t parity, volatile int* abort) {
    uint32_t done = 0;
    for (int spin = 0; spin < (1 << 22); ++spin) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(done) : "r"(mbar), "r"(parity) : "memory");
        if (done) return true;
        if ((spin & 255) == 255 && *abort) return false;
    }
    *abort = 1;
    return false;
}
__device__ __forceinline__ void ts_arrive(uint32_t mbar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(mbar) : "memory"); }
__device__ __forceinline__ void ts_commit(uint32_t mbar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar) : "memory");
}

// TMA producer: one elected lane; the whole search (S simulations x n_tiles tiles) in ring order.
__device__ __forceinline__ bool ts_producer(const TsSched& ts, const CUtensorMap* tmap, uint32_t ring, uint32_t full0, uint32_t empty0, int S, volatile int* abort,
                                            unsigned long long* dbg) {
    bool ok = true;
    int g = 0;
    long long w_empty = 0;
    const long long t_begin = dbg ? clock64() : 0;
    for (int it = 0; it < S && ok; ++it)
        for (int t = 0; t < ts.n_tiles; ++t, ++g) {
            const int s = g & (TS_NS - 1);
            const long long c0 = dbg ? clock64() : 0;
            if (g >= TS_NS) ok = ts_wait(empty0 + 8 * s, (uint32_t)(((g / TS_NS) - 1) & 1), abort) && ok;
            if (dbg) w_empty += clock64() - c0;
            const uint32_t dst = ring + s * TS_SLOT, mb = full0 + 8 * s;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "n"(TS_SLOT) : "memory");
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                         ::"r"(dst), "l"(tmap), "r"(0), "r"(ts.tile[t].img_row), "r"(mb) : "memory");
        }
    if (dbg) {  // [24] producer: waiting for a free slot, [25] producer: total
        atomicAdd(dbg + 24, (unsigned long long)w_empty);
        atomicAdd(dbg + 25, (unsigned long long)(clock64() - t_begin));
    }
    return ok;
}

// MMA issuer: one elected lane.
__device__ __forceinline__ bool ts_issuer(const TsSched& ts, uint32_t tmem, uint32_t arena_s, uint32_t ring, uint32_t full0,
                                          uint32_t empty0, uint32_t acc0, uint32_t bready, int S, volatile int* abort,
                                          unsigned long long* dbg) {
    bool ok = true;
    int g = 0;
    uint32_t br = 0;
    long long w_b[TS_MAX_STAGES] = {0, 0, 0, 0, 0, 0, 0}, w_full = 0, c0 = 0, c1 = 0;
    const int n_stages = ts.n_stages, nt = ts.nt;
    const uint32_t idesc_w1 = ts_idesc(2 * nt), idesc_w2 = ts_idesc(nt);  // W1 [X1 | X2]: 2 nt columns; W2 X1: nt
    const long long t_begin = dbg ? clock64() : 0;
    for (int it = 0; it < S && ok; ++it)
        for (int st = 0; st < n_stages; ++st) {
            if (dbg) c0 = clock64();
            ok = ts_wait(bready, br, abort) && ok;  // all tree warps have published this stage's B operand
            if (dbg) w_b[st] += clock64() - c0;
            br ^= 1u;
            tc_fence_after();
            for (int t = ts.stage_tile[st]; t < ts.stage_tile[st + 1]; ++t, ++g) {
                const TsTile& tl = ts.tile[t];
                const int s = g & (TS_NS - 1);
                if (dbg) c1 = clock64();
                ok = ts_wait(full0 + 8 * s, (uint32_t)((g / TS_NS) & 1), abort) && ok;
                if (dbg) w_full += clock64() - c1;
                tc_fence_after();
                // A tile: core matrices of 8 rows x 8 inputs (128 bytes), a_sbo bytes per 8-row group
                uint32_t a_lo = ts_desc_lo(ring + s * TS_SLOT, 128);
                const uint32_t a_hi = ts_desc_hi((uint32_t)tl.a_sbo);
                uint32_t b1 = ts_desc_lo(arena_s + tl.b_off, (uint32_t)tl.b_lbo);
                const uint32_t b_hi = ts_desc_hi((uint32_t)tl.b_sbo);
                const uint32_t bstep = (uint32_t)tl.b_step;
                const int ks = tl.ksteps;
                if (tl.kind == 0) {  // W1 [X1 | X2] -> acc0 | acc1: one MMA of N = 128
                    const uint32_t idesc = idesc_w1 | ((uint32_t)tl.b_mn << 16);
                    const uint32_t d = tmem + (uint32_t)tl.d_col;
                    uint32_t acc = tl.first ? 0u : 1u;
#pragma unroll 4
                    for (int j = 0; j < ks; ++j) {
                        ts_mma(d, a_lo, a_hi, b1, b_hi, idesc, acc);
                        a_lo += 16; b1 += bstep; acc = 1u;
                    }
                } else {             // W2 X1 -> acc1
                    const uint32_t idesc = idesc_w2 | ((uint32_t)tl.b_mn << 16);
                    const uint32_t d = tmem + (uint32_t)(tl.d_col + nt);
#pragma unroll 4
                    for (int j = 0; j < ks; ++j) {
                        ts_mma(d, a_lo, a_hi, b1, b_hi, idesc, 1u);
                        a_lo += 16; b1 += bstep;
                    }
                }
                ts_commit(empty0 + 8 * s);
                if (tl.acc_done) ts_commit(acc0 + 8 * (tl.acc_done - 1));
            }
        }
    if (dbg) {  // issuer: [16 + s] waiting for the B operand of stage s, [20] waiting for full slots, [21] total
        for (int st = 0; st < TS_STAGES; ++st) atomicAdd(dbg + 16 + st, (unsigned long long)w_b[st]);  // (search kernel: 4 stages)
        atomicAdd(dbg + 20, (unsigned long long)w_full);
        atomicAdd(dbg + 21, (unsigned long long)(clock64() - t_begin));
    }
    return ok;
}

__device__ __forceinline__ void ts_ld8(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr));
}
__device__ __forceinline__ void ts_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// v = v1 + v2 / 2048 for 8 trees of one feature -> two 16-byte chunks (MN-major core-matrix rows)
__device__ __forceinline__ void ts_split8(const float (&y)[8], uint4& hi, uint4& lo) {
    __half2 h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {  // (packed conversions: one cvt.rn.f16x2.f32 per pair)
        h[i] = __floats2half2_rn(y[2 * i], y[2 * i + 1]);
        const float2 f = __half22float2(h[i]);
        l[i] = __floats2half2_rn((y[2 * i] - f.x) * TS_SPLIT, (y[2 * i + 1] - f.y) * TS_SPLIT);
    }
    hi = make_uint4(*reinterpret_cast<uint32_t*>(&h[0]), *reinterpret_cast<uint32_t*>(&h[1]), *reinterpret_cast<uint32_t*>(&h[2]),
                    *reinterpret_cast<uint32_t*>(&h[3]));
    lo = make_uint4(*reinterpret_cast<uint32_t*>(&l[0]), *reinterpret_cast<uint32_t*>(&l[1]), *reinterpret_cast<uint32_t*>(&l[2]),
                    *reinterpret_cast<uint32_t*>(&l[3]));
}

// Epilogue of one accumulator tile of kind 0 (hidden layer -> B operand) or 2 (logits -> out strip) by one tree warp:
// lane quarter q, trees c0 .. c0 + 15, eight trees per round.
__device__ __forceinline__ bool ts_epi_warp(const TsEpi& ep, uint32_t tmem_q, int q, int c0, int lane, unsigned char* __restrict__ arena,
                                            float* __restrict__ outs, int out_pitch, const float* __restrict__ ctab,
                                            const int* __restrict__ row_action, uint32_t accbar, uint32_t parity, volatile int* abort,
                                            float& amax, int nt) {
    // A warp whose lane quarter holds no row of the tile still waits for the tile: a warp must not publish the NEXT stage's
    // operand (mbarrier arrival) before the CURRENT stage has started, or its arrival would be counted in the phase that
    // is still collecting the other warps' arrivals and release the MMAs early. The accumulator barrier of this stage
    // cannot complete before that phase has.
    if (32 * q >= ep.rows) return ts_wait(accbar, parity, abort);  // (warp-uniform)
    const int r = 32 * q + lane;
    const bool on = r < ep.rows;
    const float bias = on ? ctab[ep.cidx + r] : 0.0f;
    const float* erow = ctab + ep.eidx + r;
    const int e_ld = ep.e_ld;
    const bool onehot = ep.eidx >= 0 && on;
    const int k = ep.feat0 + r;
    const bool ok = ts_wait(accbar, parity, abort);
    tc_fence_after();
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int n0 = c0 + 8 * h;
        uint32_t v0[8], v1[8];
        ts_ld8(tmem_q + (uint32_t)(ep.d_col + n0), v0);
        ts_ld8(tmem_q + (uint32_t)(ep.d_col + nt + n0), v1);
        float add[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) add[i] = bias;
        if (onehot) {  // one-hot action input == one gathered weight row per tree (models.py:236-243)
            const int4 a0 = *reinterpret_cast<const int4*>(row_action + n0), a1 = *reinterpret_cast<const int4*>(row_action + n0 + 4);
            add[0] += erow[a0.x * e_ld]; add[1] += erow[a0.y * e_ld]; add[2] += erow[a0.z * e_ld]; add[3] += erow[a0.w * e_ld];
            add[4] += erow[a1.x * e_ld]; add[5] += erow[a1.y * e_ld]; add[6] += erow[a1.z * e_ld]; add[7] += erow[a1.w * e_ld];
        }
        ts_ld_wait();
        float y[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) y[i] = fmaf(__uint_as_float(v1[i]), TS_INV_SPLIT, __uint_as_float(v0[i])) + add[i];
        if (!on) continue;
        if (ep.kind == 2) {
            float* o = outs + n0 * out_pitch + k;
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i * out_pitch] = y[i];
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                y[i] = elu_fast(y[i]);
                amax = fmaxf(amax, fabsf(y[i]));  // (NaN-free inputs; an overflow shows as inf)
            }
            uint4 hi, lo;
            ts_split8(y, hi, lo);
            unsigned char* d = arena + ep.dst + (n0 >> 3) * ep.dst_sbo + k * 16;  // (k / 8) * 128 + (k % 8) * 16
            *reinterpret_cast<uint4*>(d) = hi;
            *reinterpret_cast<uint4*>(d + ep.dst_lo) = lo;
        }
    }
    return ok;
}

// float <-> int with the same ordering (an involution)
__device__ __forceinline__ int ts_key(float f) {
    const int i = __float_as_int(f);
    return i ^ ((i >> 31) & 0x7fffffff);
}
__device__ __forceinline__ float ts_unkey(int k) { return __int_as_float(k ^ ((k >> 31) & 0x7fffffff)); }

// The network evaluation of the CTA's 64 leaves (models.py:233-257, 283-287); called by all 512 tree threads after
// the hidden states have been gathered and ctl.hdst / row_action written. w: warp (provably uniform), it: simulation.
template <int TREE_THREADS>
__device__ __forceinline__ bool ts_mlp(const TsSched& ts, TsCtl* __restrict__ ctl, uint32_t tmem_base, unsigned char* __restrict__ arena,
                                       float* __restrict__ outs, const float* __restrict__ ctab, const int* __restrict__ row_action, int w,
                                       int lane, int it, float& amax, unsigned long long* dbg) {
    const int q = w & 3, c0 = 16 * (w >> 2);
    const uint32_t tmem_q = tmem_base + ((uint32_t)(32 * q) << 16);
    const uint32_t acc0 = smem_u32(&ctl->acc[0]), bready = smem_u32(&ctl->bready);
    const uint32_t parity = (uint32_t)(it & 1);
    volatile int* abort = &ctl->abort;
    bool ok = true;
    const bool tm = dbg && threadIdx.x == 0;
    long long t0 = 0;
    // the gathered hidden state is in place: publish it
    fence_async_smem();
    __syncwarp();
    if (lane == 0) ts_arrive(bready);
    const int n_stages = ts.n_stages;
    for (int st = 0; st < n_stages; ++st) {
        if (tm) t0 = clock64();
        const int a_begin = ts.stage_acc[st], a_end = ts.stage_acc[st + 1];
        bool has_xs = false;
        for (int a = a_begin; a < a_end; ++a) has_xs = has_xs || ts.acc[a].kind == 1;
        if (has_xs) {  // (the next-state tiles are issued first: they are on the critical path)
            // ---- pass 1: minimum / maximum of every tree's next state (models.py:248-250) ----
            // (the second pass reads the accumulators again: keeping the 32 values per thread in registers instead was
            // measured -- at the 96 registers this CTA shape allows it spills, and the spills slow the selection down more
            // than the 2k cycles of tensor-memory reads it saves)
            int kmn[16], kmx[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) { kmn[i] = INT_MAX; kmx[i] = INT_MIN; }
            for (int a = a_begin; a < a_end; ++a) {
                const TsEpi& ep = ts.acc[a];
                if (ep.kind != 1) continue;
                ok = ts_wait(acc0 + 8 * a, parity, abort) && ok;
                tc_fence_after();
                if (32 * q >= ep.rows) continue;
                const int r = 32 * q + lane;
                const bool on = r < ep.rows;
                const float bias = on ? ctab[ep.cidx + r] : 0.0f;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    uint32_t v0[8], v1[8];
                    ts_ld8(tmem_q + (uint32_t)(ep.d_col + c0 + 8 * h), v0);
                    ts_ld8(tmem_q + (uint32_t)(ep.d_col + ts.nt + c0 + 8 * h), v1);
                    ts_ld_wait();
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float y = fmaf(__uint_as_float(v1[i]), TS_INV_SPLIT, __uint_as_float(v0[i])) + bias;
                        const int ky = ts_key(y);
                        if (on) {
                            kmn[8 * h + i] = min(kmn[8 * h + i], ky);
                            kmx[8 * h + i] = max(kmx[8 * h + i], ky);
                        }
                    }
                }
            }
            long long u0 = 0, u1 = 0, u2 = 0;
            if (tm) u0 = clock64();
            int mine = 0;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int rmn = __reduce_min_sync(FULL, kmn[i]), rmx = __reduce_max_sync(FULL, kmx[i]);
                if (lane == i) mine = rmn;
                if (lane == 16 + i) mine = rmx;
            }
            int (*mmk)[2] = ctl->mmk[it & 1];
            if (lane < 16) atomicMin(&mmk[c0 + lane][0], mine);
            else atomicMax(&mmk[c0 + lane - 16][1], mine);
            if (tm) u1 = clock64();
            asm volatile("bar.sync 1, %0;" ::"n"(TREE_THREADS) : "memory");
            if (tm) u2 = clock64();
            {   // the other parity's keys are free again: reset them for the next simulation
                const int t = threadIdx.x;
                if (t < TS_N) {
                    ctl->mmk[(it + 1) & 1][t][0] = INT_MAX;
                    ctl->mmk[(it + 1) & 1][t][1] = INT_MIN;
                }
            }
            // ---- pass 2: normalise -> B operand of the prediction heads + hidden-state pool ----
            float2* mnv = ctl->mnv[w];  // {minimum, 1 / scale} of this warp's 16 trees: computed once, by lanes 0..15
            if (lane < 16) {
                const int2 kk = *reinterpret_cast<const int2*>(&mmk[c0 + lane][0]);
                const float mn = ts_unkey(kk.x);
                float sc = __fsub_rn(ts_unkey(kk.y), mn);
                if (sc < 1e-5f) sc = __fadd_rn(sc, 1e-5f);  // scale += 1e-5 only where scale < 1e-5
                mnv[lane] = make_float2(mn, __frcp_rn(sc));
            }
            __syncwarp();
            for (int a = a_begin; a < a_end; ++a) {
                const TsEpi& ep = ts.acc[a];
                if (ep.kind != 1 || 32 * q >= ep.rows) continue;
                const int r = 32 * q + lane;
                const bool on = r < ep.rows;
                const float bias = on ? ctab[ep.cidx + r] : 0.0f;
                const int k = ep.feat0 + r;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int n0 = c0 + 8 * h;
                    uint32_t v0[8], v1[8];
                    ts_ld8(tmem_q + (uint32_t)(ep.d_col + n0), v0);
                    ts_ld8(tmem_q + (uint32_t)(ep.d_col + ts.nt + n0), v1);
                    float mn[8], inv[8];
#pragma unroll
                    for (int i = 0; i < 8; i += 2) {
                        const float4 m2 = *reinterpret_cast<const float4*>(&mnv[8 * h + i]);
                        mn[i] = m2.x; inv[i] = m2.y; mn[i + 1] = m2.z; inv[i + 1] = m2.w;
                    }
                    ts_ld_wait();
                    if (!on) continue;
                    float y[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float x = fmaf(__uint_as_float(v1[i]), TS_INV_SPLIT, __uint_as_float(v0[i])) + bias;
                        y[i] = __fmul_rn(__fsub_rn(x, mn[i]), inv[i]);
                    }
                    uint4 hi, lo;
                    ts_split8(y, hi, lo);
                    unsigned char* d = arena + ep.dst + (n0 >> 3) * ep.dst_sbo + k * 16;
                    *reinterpret_cast<uint4*>(d) = hi;
                    *reinterpret_cast<uint4*>(d + ep.dst_lo) = lo;
#pragma unroll
                    for (int i = 0; i < 8; i += 2) {
                        const ulonglong2 hp = *reinterpret_cast<const ulonglong2*>(&ctl->hdst[n0 + i]);
                        // (streaming stores: see the gather in search_fast_body)
                        if (hp.x) asm volatile("st.global.cs.f32 [%0], %1;" ::"l"(reinterpret_cast<float*>(hp.x) + k), "f"(y[i]) : "memory");
                        if (hp.y) asm volatile("st.global.cs.f32 [%0], %1;" ::"l"(reinterpret_cast<float*>(hp.y) + k), "f"(y[i + 1]) : "memory");
                    }
                }
            }
            if (tm) {  // [28] accumulators + minimum / maximum, [29] REDUX + atomics, [30] barrier, [31] second pass
                atomicAdd(dbg + 28, (unsigned long long)(u0 - t0));
                atomicAdd(dbg + 29, (unsigned long long)(u1 - u0));
                atomicAdd(dbg + 30, (unsigned long long)(u2 - u1));
                atomicAdd(dbg + 31, (unsigned long long)(clock64() - u2));
            }
        }
        for (int a = a_begin; a < a_end; ++a)
            if (ts.acc[a].kind != 1)
                ok = ts_epi_warp(ts.acc[a], tmem_q, q, c0, lane, arena, outs, ts.out_pitch, ctab, row_action, acc0 + 8 * a, parity, abort,
                                 amax, ts.nt) && ok;
        if (st + 1 < n_stages) {  // this warp's share of the next stage's B operand is in place
            fence_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) ts_arrive(bready);
        }
        if (tm) atomicAdd(dbg + 8 + 2 * st, (unsigned long long)(clock64() - t0));
    }
    tc_fence_before();
    asm volatile("bar.sync 1, %0;" ::"n"(TREE_THREADS) : "memory");  // the logits of every tree are in the out strip
    return ok;
}

// ------------------------------------------------------------------------------------------------ root network ----
// initial_inference (models.py:259-281: representation -> min-max scale -> policy / value heads) + support_to_scalar +
// the root of the search (self_play.py:335-376: softmax over the ordered legal list, exploration noise, root block and
// tree header) for TR_N observation rows per CTA on the same streamed tcgen05 machinery: the rows are the MMAs' columns,
// every Linear layer streams its fp16-split weight tiles by TMA, the min-max scaling runs in the accumulator layout and
// writes the hidden state straight to where it is kept (the tree's hidden-state slot 0, or the caller's array).
constexpr int TR_N = 32;           // rows per CTA: the K-major image of 32 observation rows (up to ~1,000 inputs) fits beside the ring
constexpr int TR_THREADS = TR_N * GW + 64;

struct RootTcsParams {
    int B, b0;                 // rows [b0, B) of this launch
    int A, H, support, in_dim;
    const float* in;           // observations [B][in_dim]
    const float* ctab;
    int* status;
    // outputs of tsp_initial_inference (device, may be null)
    float* value; float* policy_logits; float* hidden_out; float* value_logits;
    // the root of the search (tree_init != 0)
    int tree_init, add_noise, uniform_policy, blocks_per_tree, hidden_per_tree;
    double one_minus_frac, frac;
    void* blocks; float* hidden_pool; TreeHeader* hdr; double* rootprior; int* rootact;
    const int* legal; const int* num_legal; const double* noise;
    float* trace_root; float* root_pred; float* root_hidden;
};

__global__ void __launch_bounds__(TR_THREADS, 1)
root_tcs_kernel(const __grid_constant__ RootTcsParams p, const __grid_constant__ TsSched ts, const __grid_constant__ CUtensorMap tmap) {
    extern __shared__ __align__(16) float smem[];
    constexpr int ROW_THREADS = TR_N * GW;
    const int tid = threadIdx.x;
    float* ctab = smem;
    const int ctab_n = (ts.ctab_floats + 3) & ~3;
    TsCtl* ctl = reinterpret_cast<TsCtl*>(smem + ctab_n);
    const uint32_t base_s = smem_u32(smem);
    const int goff = (int)(((base_s + (uint32_t)(reinterpret_cast<unsigned char*>(ctl + 1) - reinterpret_cast<unsigned char*>(smem)) + 127u) & ~127u) - base_s);
    unsigned char* gbase = reinterpret_cast<unsigned char*>(smem) + goff;  // TMA destinations: 128-byte aligned
    const uint32_t ring = base_s + (uint32_t)goff;
    unsigned char* arena = gbase + TS_NS * TS_SLOT;
    float* outs = reinterpret_cast<float*>(arena + ts.arena_bytes);
    __builtin_assume(__isShared(arena));
    __builtin_assume(__isShared(outs));
    __builtin_assume(__isShared(ctab));
    __builtin_assume(__isShared(ctl));
    const int warp_u = __shfl_sync(FULL, tid >> 5, 0);
    for (int i = tid; i < ts.ctab_floats; i += TR_THREADS) ctab[i] = p.ctab[i];
    for (int i = tid; i < (ts.arena_bytes >> 4); i += TR_THREADS) reinterpret_cast<uint4*>(arena)[i] = make_uint4(0u, 0u, 0u, 0u);
    if (tid == 0) {
        for (int i = 0; i < TS_NS; ++i) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&ctl->full[i])));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&ctl->empty[i])));
        }
        for (int i = 0; i < TS_MAX_ACC; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&ctl->acc[i])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&ctl->bready)), "n"(ROW_THREADS / 32));
        ctl->abort = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < TS_N) {
        ctl->mmk[0][tid][0] = INT_MAX; ctl->mmk[0][tid][1] = INT_MIN;
        ctl->mmk[1][tid][0] = INT_MAX; ctl->mmk[1][tid][1] = INT_MIN;
        ctl->hdst[tid] = 0ull;
    }
    if (warp_u == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&ctl->tmem_slot)), "n"(TS_TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = ctl->tmem_slot;
    bool ok = true;
    float amax = 0.0f;
    const int n = tid >> 3, j = tid & 7;          // row of the CTA, lane of the row
    const int b = p.b0 + blockIdx.x * TR_N + n;
    const bool valid = tid < ROW_THREADS && b < p.B;
    float* hid = nullptr;                          // where this row's hidden state is kept
    if (valid) hid = p.tree_init ? p.hidden_pool + (size_t)b * p.hidden_per_tree * p.H : p.hidden_out + (size_t)b * p.H;
    if (warp_u >= ROW_THREADS / 32) {  // service warps: TMA producer, MMA issuer
        if (elect_one()) {
            const uint32_t full0 = smem_u32(&ctl->full[0]), empty0 = smem_u32(&ctl->empty[0]);
            if (warp_u == ROW_THREADS / 32)
                ok = ts_producer(ts, &tmap, ring, full0, empty0, 1, &ctl->abort, nullptr);
            else
                ok = ts_issuer(ts, tmem_base, smem_u32(arena), ring, full0, empty0, smem_u32(&ctl->acc[0]), smem_u32(&ctl->bready), 1,
                               &ctl->abort, nullptr);
        }
        __syncwarp();
    } else {
        // ---- the observation row as the K-major B operand of the first layer: chunks of 8 inputs, zero padded ----
        if (j == 0) ctl->hdst[n] = valid ? (unsigned long long)hid : 0ull;
        {
            const int nchunks = ts.x0_sbo / TS_X0_LBO;  // padded input width / 8
            unsigned char* d = arena + ts.x0_off + (n >> 3) * ts.x0_sbo + (n & 7) * 16;
            const float* src = p.in + (size_t)(valid ? b : 0) * p.in_dim;
            // (four chunks = 32 independent loads in flight per lane: the row comes from HBM, and one chunk at a time made
            // this loop a chain of eight memory latencies -- a quarter of the kernel)
            for (int c0 = j; c0 < nchunks; c0 += 4 * GW) {
                float y[4][8];
#pragma unroll
                for (int u = 0; u < 4; ++u)
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int k = 8 * (c0 + u * GW) + i;
                        y[u][i] = (valid && k < p.in_dim) ? __ldg(src + k) : 0.0f;
                    }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int c = c0 + u * GW;
                    if (c < nchunks) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) amax = fmaxf(amax, fabsf(y[u][i]));
                        uint4 hi, lo;
                        ts_split8(y[u], hi, lo);
                        *reinterpret_cast<uint4*>(d + c * TS_X0_LBO) = hi;
                        *reinterpret_cast<uint4*>(d + c * TS_X0_LBO + ts.x0_lo) = lo;
                    }
                }
            }
        }
        ok = ts_mlp<ROW_THREADS>(ts, ctl, tmem_base, arena, outs, ctab, nullptr, warp_u, tid & 31, 0, amax, nullptr);
        // ---- the row's outputs: 8 lanes per row ----
        const unsigned gmask = 0xFFu << ((tid & 31) & 24);
        if (valid) {
            const float* o = outs + n * ts.out_pitch;
            const int V = 2 * p.support + 1;
            const float value = support_to_scalar_group(o + ts.off_vl, p.support, j, gmask);
            if (p.value && j == 0) p.value[b] = value;
            if (p.policy_logits && j < p.A) p.policy_logits[(size_t)b * p.A + j] = o[ts.off_pl + j];
            if (p.value_logits)
                for (int i = j; i < V; i += GW) p.value_logits[(size_t)b * V + i] = o[ts.off_vl + i];
            if (p.tree_init) {
                // ---- root of the search: self_play.py:335-376 ----
                const int nl = p.num_legal ? p.num_legal[b] : p.A;
                const int a_j = (j < nl) ? (p.legal ? p.legal[(size_t)b * p.A + j] : j) : 0;
                const float lg = (j < nl && !p.uniform_policy) ? o[ts.off_pl + a_j] : 0.0f;
                const float prior = softmax_group(lg, j < nl, gmask);  // Node.expand, self_play.py:532-538
                if (p.trace_root) {
                    float* tr = p.trace_root + (size_t)b * REC;
                    for (int i = j; i < REC; i += GW) tr[i] = 0.0f;
                    __syncwarp(gmask);
                    if (j == 0) { tr[0] = 1.0f; tr[1] = value; }
                    if (j < nl) tr[4 + j] = prior;
                }
                double pr64 = (double)prior;
                if (p.add_noise && j < nl)  // add_exploration_noise, self_play.py:540-549
                    pr64 = __dadd_rn(__dmul_rn(pr64, p.one_minus_frac), __dmul_rn(p.noise[(size_t)b * p.A + j], p.frac));
                p.rootprior[(size_t)b * GW + j] = (j < nl) ? pr64 : 0.0;
                p.rootact[(size_t)b * GW + j] = a_j;
                NodeBlock<5>* blk = reinterpret_cast<NodeBlock<5>*>(p.blocks) + (size_t)b * p.blocks_per_tree;
                if (j < 5) {
                    blk->visit[j] = 0;
                    blk->child[j] = -1;
                    blk->prior[j] = prior;
                    blk->reward[j] = 0.0f;
                    blk->vsum[j] = 0.0;
                }
                if (j == 0) {
                    blk->parent = -1;
                    blk->aux = 0;
                    TreeHeader h;
                    h.mm_min = CUDART_INF;
                    h.mm_max = -CUDART_INF;
                    h.root_vsum = 0.0;
                    h.root_visit = 0;
                    h.n_blocks = 1;
                    h.n_tgroups = 0;
                    h.max_depth = 0;
                    h.n_sims = 0;
                    h.k_tie = 0;
                    h.k_chance = 0;
                    h.n_rec = 0;
                    h.root_pred = value;
                    h.n_root = nl;
                    p.hdr[b] = h;
                    if (p.root_pred) p.root_pred[b] = value;
                }
                // (the epilogue of the last representation layer has written the hidden state into slot 0 of the tree)
                if (p.root_hidden)
                    for (int i = j; i < p.H; i += GW) p.root_hidden[(size_t)b * p.H + i] = hid[i];
            }
        }
    }
    if (!ok && p.status) atomicExch(p.status, 1);
    if (!(amax <= TS_FP16_MAX) && p.status) atomicCAS(p.status, 0, 2);
    tc_fence_before();
    __syncthreads();
    if (warp_u == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TS_TMEM_COLS));
}

}  // namespace tsp
