// tsp_fast.cuh -- latency-engineered search kernel for the deterministic MuZero search
// (self_play.py:378-427) with the network weights resident in shared memory.
//
// Organisation: a CTA holds G independent GROUPS of 16 trees; a group is 8 warps (16 lanes per tree, two
// trees per warp), owns one 16-row activation strip and synchronises only with itself (named barrier,
// 256 threads), so the groups of an SM drift out of phase and one group's tensor-core MLP overlaps another
// group's selection / backup. All groups of a CTA share one shared-memory copy of the weights. The search
// is bound by the serial instruction stream of a warp (ncu: ~8 cycles per issued instruction, issue slots
// 25 % busy), so the design spreads a group's work over many warps: two trees per warp keep the number of
// tree levels a warp walks close to one tree's depth, and the same 8 warps split the 16 x 8 MLP tiles.
//
// MLP: every Linear layer is cut into 16 x 8 output tiles (mma.sync.m16n8k8, 3xTF32 split, fp32
// accumulate -- float32-level accuracy, see tsp_device.cuh). Both operands live in shared memory in
// FRAGMENT ORDER so that one LDS.128 yields the registers of an mma operand with no shuffling:
//   * activations: rows g and g + 8 of the 16-row tile are interleaved, element (r, k) of a group's strip at
//     (r & 7) * rs + (k >> 1) * 4 + (k & 1) * 2 + (r >> 3); lane (g, t) reads {X[g][k], X[g+8][k], X[g][k+1],
//     X[g+8][k+1]}, k = 8 s + 2 t, == {a0, a1, a2, a3} of k-step s (k-slots t, t + 4 <-> inputs k, k + 1). The
//     same float4 is what the accumulator fragment {c0, c2, c1, c3} of a lane is stored as. The producer also
//     stores lo = x - trunc_tf32(x) in a twin strip, so consumers never split activations (the tensor core
//     ignores the 13 low mantissa bits of a TF32 operand: the raw float32 IS the hi operand).
//   * weights: per 8-column tile, chunk of 16 inputs and lane: {W[n][16c+2t], W[n][16c+2t+1], W[n][16c+8+2t],
//     W[n][16c+9+2t]}, n = 8 j + g: the B operands of two k-steps in one LDS.128 (split in registers).
// A host-built tile list per stage and warp replaces all index arithmetic; tiles that share their input are
// batched so the A fragments are loaded once.
//
// Selection runs a float32 FILTER with a rigorous error bound in front of the float64 arithmetic of the
// reference: the exact code path only decides near-ties and exact ties (about 0.6 % of the levels), so
// the result is bit-identical to the float64 arg-max of self_play.py:436-477 at a fraction of its cost.
#pragma once
#include "tsp_device.cuh"

namespace tsp {

constexpr int GT = 16;               // trees per group (one mma row tile)
// Lanes per tree (template parameter LPT of the kernel): lanes 0..7 score the children, all LPT lanes move data.
//   LPT = 16: 8 warps per group, two trees per warp  -- shortest warp programs, for batches that leave the GPU
//             latency-bound (<= ~2 groups per SM, e.g. 4,096 roots);
//   LPT =  8: 4 warps per group, four trees per warp -- fewer issued instructions per simulation and 64 trees per SM
//             (four groups) at the same 16 warps, for large batches (throughput-bound).
constexpr int MAX_GWARPS = 8;
constexpr int FAST_MAX_TILES = 200;
constexpr int FAST_MAX_STAGES = 12;
constexpr int PATH_LEVELS = 32;      // levels recorded per descent for the lane-parallel backup
constexpr unsigned FULL = 0xffffffffu;  // the two trees of a warp run in lock step: every shuffle is full-mask

// One 16 x 8 output tile of a Linear layer, all offsets in floats and final (no arithmetic left for the
// device but adding the lane / group bases).
struct alignas(16) FUnit {
    int x_off;    // A operand: 2 * (column offset of the input buffer) inside the interleaved strip
    int w_off;    // fragment block [kc][32 lanes][4] of this tile in the weight arena
    int y_off;    // 2 * (column offset of the tile's 8 output columns)
    int flags;    // bits 0..7: kc (chunks of 16 inputs), bit 8: ELU, bit 9: store lo parts, bit 10: one-hot action
                  // rows, bit 11: one-hot future rows (MuZeroStochasticConcat, models.py:568-580)
    int b_off;    // 8 bias values of this tile
    int e_off;    // one-hot rows: action a at e_off + a * e_ld, future z at e_off + (e_fut + z) * e_ld
    int e_ld;
    int e_fut;    // number of action rows in front of the future rows
};
static_assert(sizeof(FUnit) == 32, "FUnit is two 16-byte words");

struct FastSched {
    int n_stages, row_stride, n_units, w_floats;
    short off_hidden, off_pl, off_vl, off_rl, off_tl, off_tp, pad[2];
    int4 stage[FAST_MAX_STAGES];          // {row_op, 2 * r_in, 2 * r_out, r_n}; row_op 1: min-max normalise
    int ulist[FAST_MAX_STAGES][MAX_GWARPS];   // units of warp w in stage s: first | count << 16
    FUnit units[FAST_MAX_TILES];
};
static_assert(sizeof(FastSched) % 16 == 0, "FastSched is copied in 16-byte words");

struct FastParams {
    int B, S, A, H, support, mask_absorbing, uniform_policy;
    int blocks_per_tree, hidden_per_tree, max_rec, T, pb_stride, w_lo, two_player;
    int S_tab;                // simulations the trees hold after this run (size of the visit-count tables)
    int grp_base, grp_extra;  // groups of 16 trees per CTA: grp_base, one more in the first grp_extra CTAs (multi-wave launches of
                              // the one-CTA-per-SM kernels spread the groups evenly over full waves instead of leaving the last
                              // wave short: 7 groups per SM as 4 + 3 rather than 4 + 4 on three quarters of the SMs)
    int resume, add_noise;    // resume: continue the trees of the previous run (override_root_with, self_play.py:331-333)
    double one_minus_frac, frac;
    const double* noise;      // [B][A], mixed into the root priors again on resume
    float discount_f;  // (float)discount for the selection filter, negated for two players (-child.value())
    double discount;
    const double* pb_table;
    void* blocks; float* hidden_pool; TreeHeader* hdr; double* rootprior; const int* rootact;
    const unsigned char* tie;
    const float* fed; float* trace;
    const FastSched* sched;  // device copy (deterministic: recurrent inference; chance-node searches: one future of the
                             // transition expansion)
    const FastSched* sched_b;  // chance-node searches: prediction
    int nf, K, max_iters;      // chance-node searches: futures, chance stream length, loop bound
    const unsigned char* chance;
    const float* wfrag;      // fragment-ordered weight arena of the search MLPs
    const float* tc_img;     // tcgen05 path: TMEM image of the weights [column][128 lanes], hi columns then lo columns
    const float* tc_ctab;    // tcgen05 path: biases and one-hot rows
    int* tc_status;          // tcgen05 path: set to 1 if an mbarrier wait timed out
    int* visit_counts; double* root_value; double* child_value; double* child_prior; double* child_reward;
    int* max_depth; int* num_records; int* total_depth;
    unsigned long long* phase_cycles;
};

__device__ __forceinline__ float rcp_approx(float x) {  // MUFU.RCP, <= 1 ulp
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
#ifndef TSP_REDUX_LPT8
#define TSP_REDUX_LPT8 0  // four trees per warp: four REDUX are no faster than the three shuffle rounds (measured: 4.34 vs 4.31 ms at 65,536 x 50)
#endif
// float -> int with the same ordering (finite values and infinities)
__device__ __forceinline__ int ordered_key(float f) {
    const int i = __float_as_int(f);
    return i ^ ((i >> 31) & 0x7fffffff);
}
template <int GTHREADS>
__device__ __forceinline__ void group_barrier(int id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(GTHREADS) : "memory"); }

// column k of a row inside the interleaved strip (see the header comment)
__device__ __forceinline__ int kofs(int k) { return ((k >> 1) << 2) + ((k & 1) << 1); }
__device__ __forceinline__ float tf32_lo(float x) { return x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }
__device__ __forceinline__ void mma_tf32v(float (&d)[4], const uint4& a, unsigned b0, unsigned b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b0), "r"(b1));
}

// ELU(alpha = 1) on the fast path: ex2.approx (2^-22 relative) keeps |error| < 5e-7 on (-inf, 0].
__device__ __forceinline__ float elu_fast(float v) {
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fminf(v, 0.0f) * 1.4426950408889634f));
    return v > 0.0f ? v : e - 1.0f;
}

// One unit: a 16 x 8 output tile over all K, by one warp. KCT > 0: compile-time chunk count, fully unrolled
// (the compiler then batches every fragment load in front of the mma chain). wlo_off: float offset from a
// weight fragment to its host-computed lo part (0: split the weights in registers).
template <int KCT>
__device__ __forceinline__ void mma_unit(const int4 ua, const int4 ub, const float* __restrict__ actL,
                                         const float* __restrict__ wL, const float* __restrict__ wT, int lo_off,
                                         int wlo_off, const int* __restrict__ row_action, int g, int z = 0) {
    const int kc = KCT ? KCT : (ua.w & 0xFF);
    const float* x = actL + ua.x;   // hi operand == the raw float32 activations; lo parts at + lo_off
    const float* w = wL + ua.y;
    float dm[4], dc[4] = {0.f, 0.f, 0.f, 0.f}, de[4] = {0.f, 0.f, 0.f, 0.f};  // hi*hi, lo*hi, hi*lo chains
    {
        const float2 b = *reinterpret_cast<const float2*>(wT + ub.x);
        dm[0] = b.x; dm[1] = b.y; dm[2] = b.x; dm[3] = b.y;
        if (ua.w & 0x400) {  // one-hot action input == one gathered weight row (models.py:236-243)
            const float2 e0 = *reinterpret_cast<const float2*>(wT + ub.y + row_action[g] * ub.z);
            const float2 e1 = *reinterpret_cast<const float2*>(wT + ub.y + row_action[g + 8] * ub.z);
            dm[0] += e0.x; dm[1] += e0.y; dm[2] += e1.x; dm[3] += e1.y;
        }
        if (ua.w & 0x800) {  // one-hot future z: the same row for every tree of the pass
            const float2 f = *reinterpret_cast<const float2*>(wT + ub.y + (ub.w + z) * ub.z);
            dm[0] += f.x; dm[1] += f.y; dm[2] += f.x; dm[3] += f.y;
        }
    }
#pragma unroll(KCT ? KCT : 2)
    for (int c = 0; c < kc; ++c) {
        const uint4 ah0 = *reinterpret_cast<const uint4*>(x + 32 * c), ah1 = *reinterpret_cast<const uint4*>(x + 32 * c + 16);
        const uint4 al0 = *reinterpret_cast<const uint4*>(x + lo_off + 32 * c);
        const uint4 al1 = *reinterpret_cast<const uint4*>(x + lo_off + 32 * c + 16);
        const uint4 bh = *reinterpret_cast<const uint4*>(w + 128 * c);
        uint4 bl;
        if (wlo_off) {
            bl = *reinterpret_cast<const uint4*>(w + wlo_off + 128 * c);
        } else {
            bl.x = __float_as_uint(tf32_lo(__uint_as_float(bh.x))); bl.y = __float_as_uint(tf32_lo(__uint_as_float(bh.y)));
            bl.z = __float_as_uint(tf32_lo(__uint_as_float(bh.z))); bl.w = __float_as_uint(tf32_lo(__uint_as_float(bh.w)));
        }
        mma_tf32v(dm, ah0, bh.x, bh.y);
        mma_tf32v(dc, al0, bh.x, bh.y);
        mma_tf32v(de, ah0, bl.x, bl.y);
        mma_tf32v(dm, ah1, bh.z, bh.w);
        mma_tf32v(dc, al1, bh.z, bh.w);
        mma_tf32v(de, ah1, bl.z, bl.w);
    }
    float d0 = dm[0] + (dc[0] + de[0]), d1 = dm[1] + (dc[1] + de[1]);
    float d2 = dm[2] + (dc[2] + de[2]), d3 = dm[3] + (dc[3] + de[3]);
    if (ua.w & 0x100) {
        d0 = elu_fast(d0); d1 = elu_fast(d1); d2 = elu_fast(d2); d3 = elu_fast(d3);
    }
    // {c0, c2, c1, c3} = {X[g][n], X[g+8][n], X[g][n+1], X[g+8][n+1]}, n = out column + 2 t
    float* y = const_cast<float*>(actL) + ua.z;
    *reinterpret_cast<float4*>(y) = make_float4(d0, d2, d1, d3);
    if (ua.w & 0x200)  // a later layer reads this buffer as an mma operand
        *reinterpret_cast<float4*>(y + lo_off) = make_float4(tf32_lo(d0), tf32_lo(d2), tf32_lo(d1), tf32_lo(d3));
}


// The same unit with the weights streamed from global memory / L2 (networks whose search MLPs do not fit shared
// memory, e.g. cross-merge: 365 KB): fragment-ordered weights make every warp load a contiguous 512-byte line;
// a 4-deep register ring keeps four of them in flight ahead of the mma chain (L2 latency ~300 cycles).
__device__ __forceinline__ void mma_unit_g(const int4 ua, const int4 ub, const float* __restrict__ actL,
                                           const float* __restrict__ wG, const int* __restrict__ row_action,
                                           int g, int lane, int z = 0) {
    const int kc = ua.w & 0xFF;
    const int t = lane & 3;
    const float* x = actL + ua.x;
    const float4* w = reinterpret_cast<const float4*>(wG + ua.y) + lane;  // chunk c at w + 32 * c
    float4 ring[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) ring[q] = __ldg(w + 32 * (q < kc ? q : 0));
    float dm[4], dc[4] = {0.f, 0.f, 0.f, 0.f}, de[4] = {0.f, 0.f, 0.f, 0.f};
    {
        const float2 b = __ldg(reinterpret_cast<const float2*>(wG + ub.x + 2 * t));
        dm[0] = b.x; dm[1] = b.y; dm[2] = b.x; dm[3] = b.y;
        if (ua.w & 0x400) {
            const float2 e0 = __ldg(reinterpret_cast<const float2*>(wG + ub.y + row_action[g] * ub.z + 2 * t));
            const float2 e1 = __ldg(reinterpret_cast<const float2*>(wG + ub.y + row_action[g + 8] * ub.z + 2 * t));
            dm[0] += e0.x; dm[1] += e0.y; dm[2] += e1.x; dm[3] += e1.y;
        }
        if (ua.w & 0x800) {
            const float2 f = __ldg(reinterpret_cast<const float2*>(wG + ub.y + (ub.w + z) * ub.z + 2 * t));
            dm[0] += f.x; dm[1] += f.y; dm[2] += f.x; dm[3] += f.y;
        }
    }
    for (int c0 = 0; c0 < kc; c0 += 4) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int c = c0 + q;
            if (c < kc) {
                const float4 Bv = ring[q];
                if (c + 4 < kc) ring[q] = __ldg(w + 32 * (c + 4));
                // (this path is tensor-pipe bound: the split arithmetic hides behind the mma of the other warps, and
                // without the twin strip of lo parts twice as many groups fit beside each other)
                const uint4 ah0 = *reinterpret_cast<const uint4*>(x + 32 * c), ah1 = *reinterpret_cast<const uint4*>(x + 32 * c + 16);
                uint4 al0, al1;
                al0.x = __float_as_uint(tf32_lo(__uint_as_float(ah0.x))); al0.y = __float_as_uint(tf32_lo(__uint_as_float(ah0.y)));
                al0.z = __float_as_uint(tf32_lo(__uint_as_float(ah0.z))); al0.w = __float_as_uint(tf32_lo(__uint_as_float(ah0.w)));
                al1.x = __float_as_uint(tf32_lo(__uint_as_float(ah1.x))); al1.y = __float_as_uint(tf32_lo(__uint_as_float(ah1.y)));
                al1.z = __float_as_uint(tf32_lo(__uint_as_float(ah1.z))); al1.w = __float_as_uint(tf32_lo(__uint_as_float(ah1.w)));
                const unsigned bl0 = __float_as_uint(tf32_lo(Bv.x)), bl1 = __float_as_uint(tf32_lo(Bv.y));
                const unsigned bl2 = __float_as_uint(tf32_lo(Bv.z)), bl3 = __float_as_uint(tf32_lo(Bv.w));
                mma_tf32v(dm, ah0, __float_as_uint(Bv.x), __float_as_uint(Bv.y));
                mma_tf32v(dc, al0, __float_as_uint(Bv.x), __float_as_uint(Bv.y));
                mma_tf32v(de, ah0, bl0, bl1);
                mma_tf32v(dm, ah1, __float_as_uint(Bv.z), __float_as_uint(Bv.w));
                mma_tf32v(dc, al1, __float_as_uint(Bv.z), __float_as_uint(Bv.w));
                mma_tf32v(de, ah1, bl2, bl3);
            }
        }
    }
    float d0 = dm[0] + (dc[0] + de[0]), d1 = dm[1] + (dc[1] + de[1]);
    float d2 = dm[2] + (dc[2] + de[2]), d3 = dm[3] + (dc[3] + de[3]);
    if (ua.w & 0x100) {
        d0 = elu_fast(d0); d1 = elu_fast(d1); d2 = elu_fast(d2); d3 = elu_fast(d3);
    }
    float* y = const_cast<float*>(actL) + ua.z;
    *reinterpret_cast<float4*>(y) = make_float4(d0, d2, d1, d3);
}

// All stages of the schedule for the 16 rows of one group; called by the warps of the group.
// WG: the weights stay in global memory (W is then the device arena).
template <int LPT, bool WG = false>
__device__ __forceinline__ void fast_mlp(const FastSched& fs, const float* __restrict__ W, float* __restrict__ act,
                                         const int* __restrict__ row_action, const unsigned char* __restrict__ row_active,
                                         int gt, int barid, int wlo_off, unsigned long long* dbg, int z = 0) {
    const int rs = fs.row_stride;
    const int wg = gt >> 5, lane = gt & 31;
    const int g = lane >> 2, t = lane & 3;
    const float* actL = act + g * rs + 4 * t;
    const float* wL = W + lane * 4;
    const float* wT = W + 2 * t;
    const int lo_off = 8 * rs;
    const int n_stages = fs.n_stages;
    // the descriptors of a stage are fetched before the barrier that ends the previous stage
    int ul = fs.ulist[0][wg];
    int4 ua = *reinterpret_cast<const int4*>(&fs.units[ul & 0xFFFF]);
    int4 ub = *reinterpret_cast<const int4*>(&fs.units[ul & 0xFFFF].b_off);
    for (int s = 0; s < n_stages; ++s) {
        long long d0 = 0, d1 = 0;
        if (dbg && threadIdx.x == 0) d0 = clock64();
        int u = ul & 0xFFFF;
        const int uend = u + (ul >> 16);
        long long q0 = 0, q1 = 0;
        while (u < uend) {
            const int kc = ua.w & 0xFF;
            if (dbg && threadIdx.x == 0) q0 = clock64();
            if (WG) mma_unit_g(ua, ub, actL, W, row_action, g, lane, z);
            else if (kc == 4) mma_unit<4>(ua, ub, actL, wL, wT, lo_off, wlo_off, row_action, g, z);
            else if (kc == 2) mma_unit<2>(ua, ub, actL, wL, wT, lo_off, wlo_off, row_action, g, z);
            else mma_unit<0>(ua, ub, actL, wL, wT, lo_off, wlo_off, row_action, g, z);
            if (dbg && threadIdx.x == 0) {
                q1 = clock64();
                atomicAdd(dbg + 40, (unsigned long long)(q0 - d0));  // stage start -> unit start (dispatch)
                atomicAdd(dbg + 41, (unsigned long long)(q1 - q0));  // the unit itself
                atomicAdd(dbg + 43, 1ull);
            }
            if (++u < uend) {
                ua = *reinterpret_cast<const int4*>(&fs.units[u]);
                ub = *reinterpret_cast<const int4*>(&fs.units[u].b_off);
            }
        }
        if (s + 1 < n_stages) {
            ul = fs.ulist[s + 1][wg];
            ua = *reinterpret_cast<const int4*>(&fs.units[ul & 0xFFFF]);
            ub = *reinterpret_cast<const int4*>(&fs.units[ul & 0xFFFF].b_off);
        }
        const int4 sg = fs.stage[s];
        if (sg.x == 1) {
            // models.py:223-231 / 248-255: (x - min) / scale, scale += 1e-5 only where scale < 1e-5
            const int row = gt / LPT, j = gt % LPT;
            {   // both trees of a warp run this in lock step (full-mask shuffles); idle rows compute on stale data
                const float* x = act + (row & 7) * rs + (row >> 3) + sg.y;
                float mn = CUDART_INF_F, mx = -CUDART_INF_F;
                for (int q = j; q < sg.w; q += LPT) {
                    const float v = x[kofs(q)];
                    mn = fminf(mn, v);
                    mx = fmaxf(mx, v);
                }
#pragma unroll
                for (int d = 1; d < LPT; d <<= 1) {
                    mn = fminf(mn, __shfl_xor_sync(FULL, mn, d));
                    mx = fmaxf(mx, __shfl_xor_sync(FULL, mx, d));
                }
                float scale = __fsub_rn(mx, mn);
                if (scale < 1e-5f) scale = __fadd_rn(scale, 1e-5f);
                const float inv = __frcp_rn(scale);
                float* y = act + (row & 7) * rs + (row >> 3) + sg.z;
                for (int q = j; q < sg.w; q += LPT) {
                    const float v = __fmul_rn(__fsub_rn(x[kofs(q)], mn), inv);
                    y[kofs(q)] = v;
                    if (!WG) y[kofs(q) + lo_off] = tf32_lo(v);
                }
            }
        }
        if (dbg && threadIdx.x == 0) d1 = clock64();
        group_barrier<GT * LPT>(barid);
        if (dbg && threadIdx.x == 0) {
            const long long d2 = clock64();
            atomicAdd(dbg + 8 + 2 * s, (unsigned long long)(d1 - d0));
            atomicAdd(dbg + 9 + 2 * s, (unsigned long long)(d2 - d1));
        }
    }
}

// value and reward decode (models.py:1180-1201) and the policy softmax (self_play.py:532-538) of one row by the LPT
// lanes of its tree. The lanes split by DISTRIBUTION: the first half of the tree's lanes decodes the value support, the
// second half the reward support -- one softmax-expectation and ONE inverse value transform (a float32 square root and two
// divisions in IEEE rounding) per lane instead of two, reductions over LPT / 2 lanes; the policy softmax (at most 8 logits,
// one per lane) rides in the same shuffle rounds. vl / rl / pl point at column 0 of the row inside the interleaved strip
// (column k at kofs(k)). LIN: the logits are plain arrays (tcgen05 path) instead of columns of the interleaved strip.
template <int LPT, bool LIN = false>
__device__ __forceinline__ void decode_row(const float* __restrict__ vl, const float* __restrict__ rl,
                                           const float* __restrict__ pl, int support, int A, bool uniform, int j,
                                           float& value, float& reward, float& prior) {
    auto kofs = [](int k) { return LIN ? k : ((k >> 1) << 2) + ((k & 1) << 1); };
    constexpr int HL = LPT / 2;
    const int n = 2 * support + 1;
    const float* sl = (j >= HL) ? rl : vl;
    const int jj = j & (HL - 1);
    float m = -CUDART_INF_F;
    for (int i = jj; i < n; i += HL) m = fmaxf(m, sl[kofs(i)]);
    const bool pon = j < A;
    const float lg = (pon && !uniform) ? pl[kofs(j)] : 0.0f;
    float mp = pon ? lg : -CUDART_INF_F;
#pragma unroll
    for (int d = 1; d < HL; d <<= 1) {
        m = fmaxf(m, __shfl_xor_sync(FULL, m, d));
        mp = fmaxf(mp, __shfl_xor_sync(FULL, mp, d));
    }
    mp = fmaxf(mp, __shfl_xor_sync(FULL, mp, HL));
    float sm = 0.f, xm = 0.f;
    for (int i = jj; i < n; i += HL) {
        const float e = __expf(__fsub_rn(sl[kofs(i)], m));
        sm = __fadd_rn(sm, e);
        xm = fmaf((float)(i - support), e, xm);
    }
    const float ep = pon ? __expf(__fsub_rn(lg, mp)) : 0.0f;
    float sp = ep;
#pragma unroll
    for (int d = 1; d < HL; d <<= 1) {
        sm = __fadd_rn(sm, __shfl_xor_sync(FULL, sm, d));
        xm = __fadd_rn(xm, __shfl_xor_sync(FULL, xm, d));
        sp = __fadd_rn(sp, __shfl_xor_sync(FULL, sp, d));
    }
    sp = __fadd_rn(sp, __shfl_xor_sync(FULL, sp, HL));
    const float sc = inverse_value_transform(__fdiv_rn(xm, sm));  // the value in the first half of the lanes, the reward in the second
    value = __shfl_sync(FULL, sc, 0, LPT);
    reward = __shfl_sync(FULL, sc, HL, LPT);
    prior = __fdiv_rn(ep, sp);
}

// ------------------------------------------------------------------ selection ----
// Exact pUCT arg-max of one tree level in float64 (self_play.py:436-477), the reference's operation
// order; out of line so that the float32 filter stays small. Lanes 0..7 of the tree hold the children.
__device__ __noinline__ int select_exact(double pbN, double sqN, int vis, double vs, double prior, float rw,
                                         double discount, double mn, double mx, bool on, int gl0,
                                         const unsigned char* tie_row, int T, int& k_tie, bool two_player) {
    MinMax mm;
    mm.mn = mn;
    mm.mx = mx;
    double score = -CUDART_INF;
    if (on) score = ucb_score<0>(pbN, sqN, vis, vs, prior, rw, discount, mm, two_player);
    double best = score;
#pragma unroll
    for (int d = 1; d < GW; d <<= 1) {
        const double o = shfl_xor_d(FULL, best, d);
        best = o > best ? o : best;
    }
    best = shfl_d(FULL, best, gl0);  // lanes 8..15 hold no child: take the maximum of lanes 0..7
    const unsigned tied = (__ballot_sync(FULL, on && score == best) >> gl0) & 0xFFu;
    const int nt = __popc(tied);
    if (nt <= 1) return __ffs(tied) - 1;
    int v = 0;  // self_play.py:444-450 with the external tie stream
    if (tie_row) v = tie_row[k_tie % T];
    k_tie++;
    return __fns(tied, 0, (v % nt) + 1);
}

// Sequential backup for paths deeper than the lane-parallel version records (self_play.py:485-500).
__device__ __noinline__ void backup_walk(NodeBlock<5>* blocks, int cur, int cs, double v, double discount, MinMax& mm,
                                         double& root_vsum, int& root_visit, bool two, bool masked) {
    int k = 0;  // plies above the leaf
    while (true) {
        NodeBlock<5>* bk = blocks + cur;
        const int ps = bk->aux, pc = bk->parent;
        const bool same = two && ((k & 1) == 0) && !(masked && k == 0);
        const double vs = __dadd_rn(bk->vsum[cs], (two && !same) ? -v : v);
        const int vis = bk->visit[cs] + 1;
        const double rw = (double)bk->reward[cs];
        bk->vsum[cs] = vs;
        bk->visit[cs] = vis;
        const double q = __ddiv_rn(vs, (double)vis);
        mm.update(__dadd_rn(rw, __dmul_rn(discount, two ? -q : q)));
        v = __dadd_rn(same ? -rw : rw, __dmul_rn(discount, v));
        ++k;
        if (cur == 0) break;
        cs = ps & 0xFF;
        cur = pc;
    }
    const bool same = two && ((k & 1) == 0);
    root_vsum = __dadd_rn(root_vsum, (two && !same) ? -v : v);
    root_visit += 1;
    const double q = __ddiv_rn(root_vsum, (double)root_visit);
    mm.update(__dadd_rn(0.0, __dmul_rn(discount, two ? -q : q)));
}

}  // namespace tsp
#include "tsp_tc.cuh"
#include "tsp_stream.cuh"
namespace tsp {

// what the streamed tcgen05 path (TC == 2, tsp_stream.cuh) needs besides FastParams: all kernel parameters (constant bank)
struct TsArgs {
    const TsSched* ts;
    const CUtensorMap* tmap;  // the weight image as boxes of 128 rows of 128 bytes: one 16 KB tile
};
constexpr int TS_SERVICE_THREADS = 64;    // TC == 2: one TMA producer warp and one MMA issuer warp behind the tree threads

// TC == 1: the MLP runs on tcgen05 with the weights resident in tensor memory (tsp_tc.cuh); ts then points at the
// kernel's TcSched parameter (constant bank). TC == 2: tcgen05 with the weights streamed by TMA (tsp_stream.cuh): the
// CTA's G * 16 = 64 trees evaluate the network in lock step, threads [THREADS, THREADS + 64) are the service warps.
template <int G, int LPT, bool WG, int TC>
__device__ __forceinline__ void search_fast_body(const FastParams& p, const TcSched* __restrict__ tsp_, const TsArgs* __restrict__ sx = nullptr) {
    extern __shared__ __align__(16) float smem[];
    constexpr int GTHREADS = GT * LPT;
    constexpr int THREADS = G * GTHREADS;
    static_assert(LPT == 8 || LPT == 16, "8 or 16 lanes per tree");
    static_assert(THREADS <= 512, "128 registers per thread");
    static_assert(TC != 2 || (G * GT == TS_N && LPT == 8), "streamed path: 64 trees, 8 lanes per tree");
    constexpr int NC = 5;
    using Block = NodeBlock<NC>;
    // shared-memory carve-out: schedule (compile-time offset 0; not on the streamed path), tables, weights, per-group strips
    constexpr int SCHED_FLOATS = TC == 2 ? 0 : (int)(sizeof(FastSched) / 4);
    FastSched* fsp = reinterpret_cast<FastSched*>(smem);
    const int pb_n = (p.S_tab + 2 + 1) & ~1;  // S_tab: largest parent visit count (sums over continued searches)
    const int pbf_n = (pb_n + 3) & ~3;
    double* pb = reinterpret_cast<double*>(smem + SCHED_FLOATS);  // pb[N] = log((N + base + 1) / base) + init
    double* sq = pb + pb_n;                                                // sq[N] = sqrt(N), host libm
    float* pbf = reinterpret_cast<float*>(sq + pb_n);                      // float32 pb[N] * sq[N] for the filter
    float* wsm = pbf + pbf_n;
    const int tid = threadIdx.x;
    if (!TC) {
        const int4* src = reinterpret_cast<const int4*>(p.sched);
        int4* dst = reinterpret_cast<int4*>(fsp);
        for (int i = tid; i < (int)(sizeof(FastSched) / 16); i += THREADS) dst[i] = src[i];
    }
    for (int i = tid; i < p.S_tab + 2; i += THREADS) {
        const double a = p.pb_table[i], c = p.pb_table[p.pb_stride + i];
        pb[i] = a;
        sq[i] = c;
        pbf[i] = (float)(a * c);
    }
    const bool fed = !TC && p.fed != nullptr;
    __syncthreads();
    const FastSched& fs = *fsp;
    const int rs = fs.row_stride;
    // p.w_lo: the arena is followed by the lo parts of the weights (same fragment layout) -- shared memory permitting
    const int wlo_off = (p.w_lo && !WG && !TC) ? fs.w_floats : 0;
    const int w_total = (WG || TC) ? 0 : fs.w_floats + wlo_off;  // WG: the weights are streamed from global memory
    if (!fed && !WG && !TC) {
        const float4* src = reinterpret_cast<const float4*>(p.wfrag);
        float4* dst = reinterpret_cast<float4*>(wsm);
        for (int i = tid; i < (w_total >> 2); i += THREADS) dst[i] = __ldg(src + i);
    }
    const int grp = tid / GTHREADS, gt = tid % GTHREADS;
    const int row = gt / LPT;                                // tree of the group == row of the strip
    const int barid = 1 + grp;
    float* act = nullptr;
    float* rowp = nullptr;
    int* row_action;
    unsigned char* row_active;
    int* path;
    // ---- TC: constant table, mbarriers and per group {B-operand arena, out strip, row tables}; weights -> TMEM ----
    const TcSched& ts = *tsp_;
    [[maybe_unused]] unsigned char* tc_arena = nullptr;
    [[maybe_unused]] float* tc_outs = nullptr;
    [[maybe_unused]] const float* ctab = wsm;
    [[maybe_unused]] uint32_t tmem_base = 0, d_group = 0, tc_mbar = 0, tc_parity = 0;
    [[maybe_unused]] int gw_u = 0;
    [[maybe_unused]] float4 hq[2];
    [[maybe_unused]] bool tc_ok = true;
    [[maybe_unused]] TsCtl* ts_ctl = nullptr;
    [[maybe_unused]] float ts_amax = 0.0f;  // largest |activation| that went through the fp16 split
    [[maybe_unused]] uint32_t ts_ring = 0;
    if constexpr (TC == 2) {
        // ---- streamed tcgen05 path: constant table, control block, weight ring, B-operand arena, out strip, row tables ----
        const TsSched& tx = *sx->ts;
        constexpr int ALL = THREADS + TS_SERVICE_THREADS;
        const int ctab_n = (tx.ctab_floats + 3) & ~3;
        ts_ctl = reinterpret_cast<TsCtl*>(wsm + ctab_n);  // 16-byte aligned: every table in front of it is a multiple of 16 bytes
        const uint32_t base_s = smem_u32(smem);
        const int goff = (int)(((base_s + (uint32_t)(reinterpret_cast<unsigned char*>(ts_ctl + 1) - reinterpret_cast<unsigned char*>(smem)) + 127u) & ~127u) - base_s);
        unsigned char* gbase = reinterpret_cast<unsigned char*>(smem) + goff;  // TMA destinations: 128-byte aligned
        ts_ring = base_s + (uint32_t)goff;
        tc_arena = gbase + TS_NS * TS_SLOT;
        tc_outs = reinterpret_cast<float*>(tc_arena + tx.arena_bytes);
        row_action = reinterpret_cast<int*>(tc_outs + TS_N * tx.out_pitch);
        row_active = reinterpret_cast<unsigned char*>(row_action + TS_N);
        path = row_action + TS_N + TS_N / 4 + ((tid / LPT) & (TS_N - 1)) * PATH_LEVELS;
        __builtin_assume(__isShared(tc_arena));
        __builtin_assume(__isShared(tc_outs));
        __builtin_assume(__isShared(row_action));
        __builtin_assume(__isShared(path));
        __builtin_assume(__isShared(ctab));
        __builtin_assume(__isShared(ts_ctl));
        gw_u = __shfl_sync(FULL, tid >> 5, 0);  // warp of the CTA, provably warp-uniform
        for (int i = tid; i < tx.ctab_floats; i += ALL) wsm[i] = p.tc_ctab[i];
        // zero arena: the padded inputs of an operand must be finite (their weights are zero)
        for (int i = tid; i < (tx.arena_bytes >> 4); i += ALL) reinterpret_cast<uint4*>(tc_arena)[i] = make_uint4(0u, 0u, 0u, 0u);
        if (tid == 0) {
            for (int i = 0; i < TS_NS; ++i) {
                asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&ts_ctl->full[i])));
                asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&ts_ctl->empty[i])));
            }
            for (int i = 0; i < TS_MAX_ACC; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&ts_ctl->acc[i])));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&ts_ctl->bready)), "n"(THREADS / 32));
            ts_ctl->abort = 0;
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        if (tid < TS_N) {
            ts_ctl->mmk[0][tid][0] = INT_MAX; ts_ctl->mmk[0][tid][1] = INT_MIN;
            ts_ctl->mmk[1][tid][0] = INT_MAX; ts_ctl->mmk[1][tid][1] = INT_MIN;
            ts_ctl->hdst[tid] = 0ull;
        }
        if (gw_u == 0) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&ts_ctl->tmem_slot)), "n"(TS_TMEM_COLS));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
        tmem_base = ts_ctl->tmem_slot;
    } else if constexpr (TC == 1) {
        const int ctab_n = (ts.ctab_floats + 3) & ~3;
        unsigned long long* mbars = reinterpret_cast<unsigned long long*>(wsm + ctab_n);  // 8-byte aligned: ctab_n % 4 == 0
        uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbars + 4);
        // (offsets from the shared-memory base, not integer-cast pointers: the compiler keeps the address space and emits
        // STS / LDS with 32-bit addresses instead of generic ST.E / LD.E with 64-bit address arithmetic)
        const int goff = ((int)(reinterpret_cast<unsigned char*>(tmem_slot + 4) - reinterpret_cast<unsigned char*>(smem)) + 127) & ~127;
        unsigned char* gbase = reinterpret_cast<unsigned char*>(smem) + goff;
        const int gstride = ts.arena_bytes + (GT * ts.out_pitch + GT + GT / 4 + GT * PATH_LEVELS) * 4;
        tc_arena = gbase + grp * gstride;
        tc_outs = reinterpret_cast<float*>(tc_arena + ts.arena_bytes);
        row_action = reinterpret_cast<int*>(tc_outs + GT * ts.out_pitch);
        row_active = reinterpret_cast<unsigned char*>(row_action + GT);
        path = row_action + GT + GT / 4 + row * PATH_LEVELS;
        __builtin_assume(__isShared(tc_arena));
        __builtin_assume(__isShared(tc_outs));
        __builtin_assume(__isShared(row_action));
        __builtin_assume(__isShared(path));
        __builtin_assume(__isShared(ctab));
        const int warp_u = __shfl_sync(FULL, tid >> 5, 0);  // provably warp-uniform: keeps the MMA issue on the uniform datapath
        const int grp_u = warp_u / (GTHREADS / 32);
        gw_u = warp_u % (GTHREADS / 32);
        for (int i = tid; i < ts.ctab_floats; i += THREADS) wsm[i] = p.tc_ctab[i];
        if (tid < G) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(mbars + tid)));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        if (warp_u == 0) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TC_TMEM_COLS));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
        }
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
        tmem_base = *tmem_slot;
        tc_mbar = smem_u32(mbars + grp_u);
        d_group = tmem_base + ts.d_col0 + grp_u * ts.d_cols_group;
        {   // weight image [column][128 lanes] -> TMEM: warp w writes lane quarter w % 4, 8 columns per store
            const uint32_t q = (uint32_t)(warp_u & 3);
            const float* img = p.tc_img + 32 * q + (tid & 31);
            for (int c0 = 8 * (warp_u >> 2); c0 < ts.n_cols; c0 += 8 * (THREADS / 128)) {
                uint32_t r[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) r[i] = __float_as_uint(__ldg(img + (size_t)(c0 + i) * 128));
                asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(tmem_base + ((32u * q) << 16) + c0),
                             "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                             : "memory");
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
        tc_fence_before();
    } else {
        // per group: strip of 8 interleaved row pairs (rs floats each) + its twin of lo parts (not with streamed
        // weights), then the row tables
        constexpr int STRIP_LINES = WG ? GT / 2 : GT;
        act = wsm + w_total + grp * (STRIP_LINES * rs + GT + GT + GT * PATH_LEVELS);
        row_action = reinterpret_cast<int*>(act + STRIP_LINES * rs);
        row_active = reinterpret_cast<unsigned char*>(row_action + GT);
        path = row_action + 2 * GT + row * PATH_LEVELS;     // (block << 8 | slot) per level of this tree's descent
        rowp = act + (row & 7) * rs + (row >> 3);           // column 0 of this tree's row in the interleaved strip
    }

    const int lane = tid & 31, j = tid % LPT;
    const int gl0 = lane & (32 - LPT);  // first lane of this tree inside the warp
    const int cta = blockIdx.x;
    const bool grp_on = grp < p.grp_base + (cta < p.grp_extra ? 1 : 0);
    const int b = (cta * p.grp_base + min(cta, p.grp_extra) + grp) * GT + row;
    const bool valid = grp_on && b < p.B && (TC != 2 || tid < THREADS);  // (the service warps of the streamed path own no tree)
    const int A = p.A, H = p.H;
    const double discount = p.discount;

    Block* blocks = reinterpret_cast<Block*>(p.blocks) + (size_t)(valid ? b : 0) * p.blocks_per_tree;
    float* hpool = p.hidden_pool + (size_t)(valid ? b : 0) * p.hidden_per_tree * H;
    // Opaque to the compiler from here on: at the 128-register cap it would otherwise re-derive these bases
    // (64-bit multiplies) inside every level of the selection loop instead of keeping them.
    asm volatile("" : "+l"(blocks), "+l"(hpool));
    __builtin_assume(__isGlobal(blocks));
    __builtin_assume(__isGlobal(hpool));
    const float* fed_rec = fed ? p.fed + (size_t)(valid ? b : 0) * p.max_rec * REC : nullptr;
    float* trace_rec = p.trace ? p.trace + (size_t)(valid ? b : 0) * p.max_rec * REC : nullptr;

    MinMax mm;
    mm.mn = CUDART_INF;
    mm.mx = -CUDART_INF;
    double root_vsum = 0.0, root_prior = 0.0;
    int root_visit = 0, n_blocks = 1, max_depth = 0, n_sims = 0, k_tie = 0, n_rec = 0;
    int n_root = 0, root_act = 0, depth_total = 0, n_exact = 0;
    if (valid) {
        const TreeHeader h = p.hdr[b];
        n_root = h.n_root;
        if (j < GW) {
            root_prior = p.rootprior[(size_t)b * GW + j];
            root_act = p.rootact[(size_t)b * GW + j];
        }
        if (p.resume) {  // MCTS.run(..., override_root_with=root): keep the tree, fresh MinMaxStats / depth / streams
            root_vsum = h.root_vsum;
            root_visit = h.root_visit;
            n_blocks = h.n_blocks;
            if (p.add_noise && j < n_root) {  // add_exploration_noise once more, self_play.py:372-376, 540-549
                root_prior = __dadd_rn(__dmul_rn(root_prior, p.one_minus_frac), __dmul_rn(p.noise[(size_t)b * A + j], p.frac));
                p.rootprior[(size_t)b * GW + j] = root_prior;
            }
        }
    }
    const float root_prior_f = (float)root_prior;
    __syncthreads();
    if constexpr (TC) tc_fence_after();

    int n_it = grp_on ? p.S : 0;  // (a group slot this CTA does not use sits the search out)
    if constexpr (TC == 2) {
        if (gw_u >= THREADS / 32) {  // service warps: the TMA producer and the MMA issuer of the whole search
            n_it = 0;
            bool ok = true;
            if (elect_one()) {
                const uint32_t full0 = smem_u32(&ts_ctl->full[0]), empty0 = smem_u32(&ts_ctl->empty[0]);
                if (gw_u == THREADS / 32)
                    ok = ts_producer(*sx->ts, sx->tmap, ts_ring, full0, empty0, p.S, &ts_ctl->abort, p.phase_cycles);
                else
                    ok = ts_issuer(*sx->ts, tmem_base, smem_u32(tc_arena), ts_ring, full0, empty0, smem_u32(&ts_ctl->acc[0]),
                                   smem_u32(&ts_ctl->bready), p.S, &ts_ctl->abort, p.phase_cycles);
            }
            __syncwarp();
            tc_ok = ok;
        }
    }
    for (int it = 0; it < n_it; ++it) {
        if (fed && n_rec >= p.max_rec) n_sims = p.S;  // fed records exhausted: stop this tree
        const bool active = valid && n_sims < p.S;
        long long tc0 = 0, tc1 = 0, tc2 = 0, tc3 = 0;
        if (p.phase_cycles && tid == 0) tc0 = clock64();
        // ================= selection, self_play.py:388-397 =================
        // The two trees of a warp descend in lock step (a finished tree idles), so every shuffle is full-mask.
        int node = 0, slot = 0, depth = 0, Np = root_visit;
        {
            // float32 image of MinMaxStats.normalize for the filter: (x - mnf) * idf, identity while max <= min
            float mnf = 0.0f, idf = 1.0f;
            if (mm.mx > mm.mn) {
                mnf = (float)mm.mn;
                idf = rcp_approx((float)__dsub_rn(mm.mx, mm.mn));
            }
            const float amnf = fabsf(mnf);
            bool walking = active;
            while (__any_sync(FULL, walking)) {
                // straight-line level: every lane loads (lanes past the children re-read child 0), masks decide
                const Block* blk = blocks + node;
                const int nch = (node == 0) ? n_root : A;
                const bool on = walking && j < nch;
                const int jc = j < NC ? j : 0;
                const int vis = blk->visit[jc];
                const int ch = on ? blk->child[jc] : -1;
                const double vs = blk->vsum[jc];
                const float rw = blk->reward[jc];
                const float prf = (node == 0) ? root_prior_f : blk->prior[jc];
                // (prefetching the children's blocks / the node's hidden state towards L1 here was measured: no effect while
                // the pools are L2-resident; the streamed path's large batches of 201-node trees are not: the five candidate
                // blocks of the next level are requested while this level is scored)
                if constexpr (TC == 2) {
                    if (ch >= 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(blocks + ch));
                }
                // ---- float32 filter: score_j +- E_j encloses the reference's float64 ucb_score; when one
                // interval lies strictly above all others, the exact arg-max is that child and it is not tied.
                const float ps = pbf[Np] * rcp_approx((float)(vis + 1)) * prf;
                const float gq = p.discount_f * ((float)vs * rcp_approx((float)vis));  // vis == 0: discarded below
                float V = ((gq + rw) - mnf) * idf;
                float aux = idf * (fabsf(gq) + fabsf(rw) + amnf);
                if (vis <= 0) {
                    V = 0.0f;
                    aux = 0.0f;
                }
                const float sc = ps + V;
                const float E = fmaf(1.9073486e-6f, fabsf(ps) + fabsf(V) + aux, 1e-30f);  // 16 * 2^-23 * magnitudes
                const float lo = on ? sc - E : -CUDART_INF_F;
                const float hi = on ? sc + E : -CUDART_INF_F;
                unsigned cand;
                if constexpr (LPT == 16 || TSP_REDUX_LPT8) {
                    // the maximum of the lower bounds by one REDUX.MAX per tree of the warp on order-preserving integer
                    // keys instead of three shuffle rounds and a broadcast (-0.0 sorts below +0.0: only ever more conservative);
                    // a non-finite bound sends the level to the float64 path like the float comparison did
                    const int klo = ordered_key(lo), khi = ordered_key(hi);
                    const int tw = lane / LPT;
                    int m1k = INT_MIN;
#pragma unroll
                    for (int t = 0; t < 32 / LPT; ++t) {
                        const int r = __reduce_max_sync(FULL, tw == t ? klo : INT_MIN);
                        if (tw == t) m1k = r;
                    }
                    const bool finite = fabsf(sc) <= 3.0e38f;
                    const unsigned bal = __ballot_sync(FULL, on && khi >= m1k);
                    const unsigned bad = __ballot_sync(FULL, on && !finite);
                    cand = (bal >> gl0) & 0xFFu;
                    if ((bad >> gl0) & 0xFFu) cand = 0;
                } else {
                    float m1 = lo;
#pragma unroll
                    for (int d = 1; d < GW; d <<= 1) m1 = fmaxf(m1, __shfl_xor_sync(FULL, m1, d));
                    m1 = __shfl_sync(FULL, m1, gl0);
                    cand = (__ballot_sync(FULL, on && hi >= m1) >> gl0) & 0xFFu;
                    // a non-finite score (NaN bounds drop out of fmaxf and of the comparison above) must not leave a
                    // lone candidate behind: the level goes to the float64 path
                    const unsigned bad = __ballot_sync(FULL, on && !(fabsf(sc) <= 3.0e38f));
                    if ((bad >> gl0) & 0xFFu) cand = 0;
                }
                int sl = __ffs(cand) - 1;
                const bool need = walking && __popc(cand) != 1;
                if (__any_sync(FULL, need)) {  // near-tie, exact tie or non-finite filter value: decide in float64
                    int k2 = k_tie;
                    const int se = select_exact(pb[Np], sq[Np], vis, vs, (node == 0) ? root_prior : (double)prf, rw, discount,
                                                mm.mn, mm.mx, on && need, gl0, p.tie ? p.tie + (size_t)b * p.T : nullptr, p.T, k2,
                                                p.two_player != 0);
                    if (need) {
                        sl = se;
                        k_tie = k2;
                        n_exact++;
                    }
                }
                const int vis_sel = __shfl_sync(FULL, vis, (gl0 + sl) & 31);
                const int child_sel = __shfl_sync(FULL, ch, (gl0 + sl) & 31);
                if (walking) {
                    slot = sl;
                    if (j == 0 && depth < PATH_LEVELS) path[depth] = (node << 8) | slot;
                    depth++;
                    if (child_sel < 0) {
                        walking = false;
                    } else {
                        node = child_sel;
                        Np = vis_sel;
                    }
                }
            }
        }
        const int ract = __shfl_sync(FULL, root_act, (gl0 + slot) & 31);
        int action = 0;
        if (active) {
            action = (node == 0) ? ract : slot;
            if (depth > max_depth) max_depth = depth;
            depth_total += depth;
        }

        // ================= network evaluation of the group's 16 leaves =================
        float value = 0.0f, reward = 0.0f, terminal = -1.0f, prior = 0.0f;
        if constexpr (TC == 2) {
            const TsSched& tx = *sx->ts;
            const int n = tid / LPT;  // tree of the CTA == column of the MMAs
            if (j == 0) {
                row_action[n] = action;
                row_active[n] = active ? 1 : 0;
                // the new node's hidden state is written by the epilogue of the next-state tiles (no terminal head on this
                // path: an active tree always expands, self_play.py:408-417)
                ts_ctl->hdst[n] = active ? (unsigned long long)(hpool + (size_t)n_blocks * H) : 0ull;
            }
            if (active) {  // gather the parent's hidden state as the K-major B operand of the first dynamics layer
                const float4* src = reinterpret_cast<const float4*>(hpool + (size_t)node * H);
                unsigned char* d = tc_arena + tx.x0_off + (n >> 3) * tx.x0_sbo + (n & 7) * 16;
                for (int c = j; c < (H >> 3); c += LPT) {
                    // (hidden states are written once and read once or twice: streaming accesses keep them from pushing the
                    // node blocks, which every simulation walks again, out of L2 -- 1 % at the HBM-resident sizes of this kernel)
                    const float4 u = __ldcs(src + 2 * c), v = __ldcs(src + 2 * c + 1);
                    const float y[8] = {u.x, u.y, u.z, u.w, v.x, v.y, v.z, v.w};
                    uint4 hi, lo;
                    ts_split8(y, hi, lo);
                    *reinterpret_cast<uint4*>(d + c * TS_X0_LBO) = hi;
                    *reinterpret_cast<uint4*>(d + c * TS_X0_LBO + tx.x0_lo) = lo;
                }
            }
            if (p.phase_cycles && tid == 0) tc1 = tc2 = clock64();
            tc_ok = ts_mlp<THREADS>(tx, ts_ctl, tmem_base, tc_arena, tc_outs, ctab, row_action, gw_u, lane, it, ts_amax, p.phase_cycles) && tc_ok;
            if (p.phase_cycles && tid == 0) tc3 = clock64();
            const float* o = tc_outs + n * tx.out_pitch;
            decode_row<LPT, true>(o + tx.off_vl, o + tx.off_rl, o + tx.off_pl, p.support, A, p.uniform_policy != 0, j, value, reward,
                                  prior);
        } else if constexpr (TC == 1) {
            if (j == 0) {
                row_action[row] = action;
                row_active[row] = active ? 1 : 0;
            }
            if (active) {  // gather the parent's hidden state as the B operand of the first dynamics layer
                const float4* src = reinterpret_cast<const float4*>(hpool + (size_t)node * H);
                unsigned char* d = tc_arena + ts.xh_off + (row >> 3) * ts.xh_sbo + (row & 7) * 16;
                for (int i = j; i < (H >> 2); i += LPT) {
                    const float4 h4 = src[i];  // (streaming loads / stores were measured here: 2-3 % slower while the pools fit L2)
                    *reinterpret_cast<float4*>(d + i * TC_LBO) = h4;
                    *reinterpret_cast<float4*>(d + i * TC_LBO + ts.b_lo) =
                        make_float4(tf32_lo(h4.x), tf32_lo(h4.y), tf32_lo(h4.z), tf32_lo(h4.w));
                }
            }
            fence_async_smem();
            if (p.phase_cycles && tid == 0) tc1 = clock64();
            group_barrier<GTHREADS>(barid);
            if (p.phase_cycles && tid == 0) tc2 = clock64();
            tc_ok = tc_mlp<LPT>(ts, tmem_base, d_group, tc_arena, tc_outs, ctab, row_action, tc_mbar, tc_parity, gw_u, lane, row, j,
                                barid, hq, p.phase_cycles) && tc_ok;
            if (p.phase_cycles && tid == 0) tc3 = clock64();
            const float* o = tc_outs + row * ts.out_pitch;
            decode_row<LPT, true>(o + ts.off_vl, o + ts.off_rl, o + ts.off_pl, p.support, A, p.uniform_policy != 0, j, value, reward,
                                  prior);
        } else if (!fed) {
            if (j == 0) {
                row_action[row] = action;
                row_active[row] = active ? 1 : 0;
            }
            if (active) {  // gather the parent's hidden state (block id == hidden slot) into columns [0, H)
                const float4* src = reinterpret_cast<const float4*>(hpool + (size_t)node * H);
                for (int i = j; i < (H >> 2); i += LPT) {
                    const float4 h4 = src[i];
                    float* d = rowp + 8 * i;  // kofs(4 i)
                    d[0] = h4.x; d[2] = h4.y; d[4] = h4.z; d[6] = h4.w;
                    if (!WG) {
                        d += 8 * rs;
                        d[0] = tf32_lo(h4.x); d[2] = tf32_lo(h4.y); d[4] = tf32_lo(h4.z); d[6] = tf32_lo(h4.w);
                    }
                }
            }
            if (p.phase_cycles && tid == 0) tc1 = clock64();
            group_barrier<GTHREADS>(barid);
            if (p.phase_cycles && tid == 0) tc2 = clock64();
            fast_mlp<LPT, WG>(fs, WG ? p.wfrag : wsm, act, row_action, row_active, gt, barid, wlo_off, p.phase_cycles);
            if (p.phase_cycles && tid == 0) tc3 = clock64();
            decode_row<LPT>(rowp + 2 * fs.off_vl, rowp + 2 * fs.off_rl, rowp + 2 * fs.off_pl, p.support, A, p.uniform_policy != 0,
                       j, value, reward, prior);
            if (p.mask_absorbing) terminal = rowp[2 * fs.off_tl];
        } else if (active) {
            const float* fr = fed_rec + (size_t)n_rec * REC;
            value = fr[1];
            reward = fr[2];
            terminal = fr[3];
            prior = (j < A) ? fr[4 + j] : 0.0f;
        }

        // ================= expansion + backup =================
        {
            // record: [0]=1 [1]=value [2]=reward [3]=terminal [4..4+A) priors; the LPT lanes write REC / LPT entries each
#pragma unroll
            for (int q = 0; q < REC / LPT; ++q) {
                const int k = j + q * LPT;
                const float pk = __shfl_sync(FULL, prior, (gl0 + ((k - 4) & (LPT - 1))) & 31);
                if (active && trace_rec && n_rec < p.max_rec) {
                    float tv = 0.0f;
                    if (k == 0) tv = 1.0f;
                    if (k == 1) tv = value;
                    if (k == 2) tv = reward;
                    if (k == 3) tv = terminal;
                    if (k >= 4 && k < 4 + A) tv = pk;
                    trace_rec[(size_t)n_rec * REC + k] = tv;
                }
            }
            if (active) n_rec++;
            Block* pblk = blocks + node;
            const bool masked = p.mask_absorbing && (terminal >= 0.0f);  // self_play.py:408-417
            double v = masked ? 0.0 : (double)value;
            if (active && !masked) {
                // ---- Node.expand, self_play.py:524-538 ----
                const int nb = n_blocks++;
                Block* nblk = blocks + nb;
                if (j < NC) {
                    nblk->visit[j] = 0;
                    nblk->child[j] = -1;
                    nblk->prior[j] = prior;
                    nblk->reward[j] = 0.0f;
                    nblk->vsum[j] = 0.0;
                }
                if (j == 0) {
                    nblk->parent = node;
                    nblk->aux = slot | (nb << 8);
                    pblk->child[slot] = nb;
                    pblk->reward[slot] = reward;
                }
                if constexpr (TC == 2) {
                    // (the epilogue of the next-state tiles has written the hidden state of the new node, tsp_stream.cuh)
                } else if constexpr (TC == 1) {  // hidden state of the new node: the normalised quads this lane produced
                    float4* dst = reinterpret_cast<float4*>(hpool + (size_t)nb * H);
#pragma unroll
                    for (int q = 0; q < 2; ++q)
                        if (j + q * LPT < (H >> 2)) dst[j + q * LPT] = hq[q];
                } else if (!fed) {  // hidden state of the new node
                    float4* dst = reinterpret_cast<float4*>(hpool + (size_t)nb * H);
                    const float* src = rowp + 2 * fs.off_hidden;
                    for (int i = j; i < (H >> 2); i += LPT) {
                        const float* q = src + 8 * i;
                        dst[i] = make_float4(q[0], q[2], q[4], q[6]);
                    }
                }
            }
            __syncwarp();
            // ---- backup, self_play.py:485-490, lane-parallel: entry e = 0 is the root, entry e >= 1 the child chosen
            // at level e - 1 of the descent. The value chain v <- r + discount * v is replayed by every lane (one
            // shuffle + 2 flops per level); the expensive per-node part (value_sum / visit_count for MinMaxStats)
            // runs once, one entry per lane. Both trees of the warp run it together (warp-uniform trip counts).
            // two players (self_play.py:492-500): entry e is the leaf player's own node iff depth - e is even, except
            // the unexpanded (masked terminal) leaf whose to_play is still -1
            const bool two = p.two_player != 0;
            const int dact = active ? depth : 0;
            int dmax = dact;  // the deepest descent among the trees of this warp
#pragma unroll
            for (int d = LPT; d < 32; d <<= 1) dmax = max(dmax, __shfl_xor_sync(FULL, dmax, d));
            if (dmax <= PATH_LEVELS) {
                double pmn = CUDART_INF, pmx = -CUDART_INF;
                for (int r = dmax / LPT; r >= 0; --r) {
                    const int e = LPT * r + j;
                    const bool mine = active && e <= depth;
                    Block* bk = blocks;
                    int cs = 0, vis0 = root_visit;
                    double vs0 = root_vsum;
                    float rwf = 0.0f;
                    if (mine && e >= 1) {
                        const int ent = path[e - 1];
                        bk = blocks + (ent >> 8);
                        cs = ent & 0xFF;
                        vs0 = bk->vsum[cs];
                        vis0 = bk->visit[cs];
                        rwf = bk->reward[cs];
                    }
                    double myv = 0.0;
                    const int e_hi = min(dmax, LPT * r + LPT - 1), e_lo = max(LPT * r, 1);
                    for (int ee = e_hi; ee >= e_lo; --ee) {
                        const double rwe = (double)__shfl_sync(FULL, rwf, gl0 + (ee & (LPT - 1)));
                        if (ee <= depth) {
                            if (ee == e) myv = v;
                            const bool same_ee = two && (((depth - ee) & 1) == 0) && !(masked && ee == depth);
                            v = __dadd_rn(same_ee ? -rwe : rwe, __dmul_rn(discount, v));
                        }
                    }
                    if (e == 0) myv = v;
                    if (mine) {
                        const bool same_e = two && (((depth - e) & 1) == 0) && !(masked && e == depth);
                        const double vs = __dadd_rn(vs0, (two && !same_e) ? -myv : myv);
                        const int vis = vis0 + 1;
                        const double q = __ddiv_rn(vs, (double)vis);
                        const double x = __dadd_rn((double)rwf, __dmul_rn(discount, two ? -q : q));
                        if (e >= 1) {
                            bk->vsum[cs] = vs;
                            bk->visit[cs] = vis;
                        }
                        pmn = pmn < x ? pmn : x;
                        pmx = pmx > x ? pmx : x;
                    }
                }
#pragma unroll
                for (int d = 1; d < LPT; d <<= 1) {
                    const double on_ = shfl_xor_d(FULL, pmn, d), ox_ = shfl_xor_d(FULL, pmx, d);
                    pmn = on_ < pmn ? on_ : pmn;
                    pmx = ox_ > pmx ? ox_ : pmx;
                }
                if (active) {
                    root_vsum = __dadd_rn(root_vsum, (two && (depth & 1)) ? -v : v);
                    root_visit += 1;
                    mm.mn = mm.mn < pmn ? mm.mn : pmn;
                    mm.mx = mm.mx > pmx ? mm.mx : pmx;
                }
            } else if (active) {
                backup_walk(blocks, node, slot, v, discount, mm, root_vsum, root_visit, two, masked);
            }
            if (active) n_sims++;
        }
        if (p.phase_cycles && tid == 0 && !fed) {
            const long long tc4 = clock64();
            atomicAdd(p.phase_cycles + 0, (unsigned long long)(tc1 - tc0));
            atomicAdd(p.phase_cycles + 1, (unsigned long long)(tc2 - tc1));
            atomicAdd(p.phase_cycles + 2, (unsigned long long)(tc3 - tc2));
            atomicAdd(p.phase_cycles + 3, (unsigned long long)(tc4 - tc3));
            atomicAdd(p.phase_cycles + 4, 1ull);
        }
        // a row's strip is re-filled by the lanes that read it; the last stage barrier ordered all tiles
        __syncwarp();
    }

    // ================= outputs: what consumers read from `root` =================
    if (valid) {
        if (j == 0) {
            if (p.root_value) p.root_value[b] = root_visit == 0 ? 0.0 : __ddiv_rn(root_vsum, (double)root_visit);
            if (p.max_depth) p.max_depth[b] = max_depth;
            if (p.num_records) p.num_records[b] = n_rec;
            if (p.total_depth) p.total_depth[b] = depth_total;
            p.hdr[b].root_vsum = root_vsum;  // what a continued search (resume) needs besides the pools
            p.hdr[b].root_visit = root_visit;
            p.hdr[b].n_blocks = n_blocks;
            if (p.phase_cycles) {  // how often the float32 filter had to defer to float64
                atomicAdd(p.phase_cycles + 5, (unsigned long long)n_exact);
                atomicAdd(p.phase_cycles + 6, (unsigned long long)depth_total);
            }
        }
        if (j < A) {
            if (p.visit_counts) p.visit_counts[(size_t)b * A + j] = 0;
            if (p.child_value) p.child_value[(size_t)b * A + j] = 0.0;
            if (p.child_prior) p.child_prior[(size_t)b * A + j] = 0.0;
            if (p.child_reward) p.child_reward[(size_t)b * A + j] = 0.0;
        }
    }
    // (outside the branch: with an odd number of roots a warp holds one valid and one invalid tree, and the lanes of
    // the invalid one do not exit here when the kernel still has to release its tensor memory)
    __syncwarp();
    if (valid) {
        const Block* rb = blocks;
        if (j < n_root) {
            const int a = root_act;
            const int vis = rb->visit[j];
            const double vs = rb->vsum[j];
            if (p.visit_counts) p.visit_counts[(size_t)b * A + a] = vis;
            if (p.child_value) p.child_value[(size_t)b * A + a] = vis == 0 ? 0.0 : __ddiv_rn(vs, (double)vis);
            if (p.child_prior) p.child_prior[(size_t)b * A + a] = root_prior;
            if (p.child_reward) p.child_reward[(size_t)b * A + a] = (double)rb->reward[j];
        }
    }
    if constexpr (TC) {
        if (!tc_ok && p.tc_status) atomicExch(p.tc_status, 1);  // an MMA commit never arrived (bounded mbarrier wait)
        if (!(ts_amax <= TS_FP16_MAX) && p.tc_status) atomicCAS(p.tc_status, 0, 2);  // an activation left the fp16 range of the streamed path
        tc_fence_before();
        __syncthreads();
        if (__shfl_sync(FULL, tid >> 5, 0) == 0)
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TC_TMEM_COLS));
    }
}

template <int G, int LPT, bool WG = false>
__global__ void __launch_bounds__(G * GT * LPT, 512 / (G * GT * LPT)) search_fast_kernel(const __grid_constant__ FastParams p) {
    search_fast_body<G, LPT, WG, 0>(p, nullptr);
}

// The same search with the MLP on tcgen05 / TMEM (tsp_tc.cuh). One CTA per SM: it owns all 512 TMEM columns.
template <int G, int LPT>
__global__ void __launch_bounds__(G * GT * LPT, 1) search_tc_kernel(const __grid_constant__ FastParams p,
                                                                    const __grid_constant__ TcSched ts) {
    search_fast_body<G, LPT, false, 1>(p, &ts);
}

// The same search with the MLP on tcgen05 and the weights streamed by TMA (tsp_stream.cuh): 64 trees per CTA (8 lanes
// per tree, 16 tree warps) + the TMA producer warp + the MMA issuer warp. One CTA per SM.
__global__ void __launch_bounds__(TS_N * 8 + TS_SERVICE_THREADS, 1)
search_tcs_kernel(const __grid_constant__ FastParams p, const __grid_constant__ TsSched ts, const __grid_constant__ CUtensorMap tmap) {
    const TsArgs sx{&ts, &tmap};
    search_fast_body<TS_N / GT, 8, false, 2>(p, nullptr, &sx);
}

// Per-thread view of the tcgen05 path's state (tsp_tc.cuh) and its set-up: constant table, mbarriers, the group's
// {B-operand arena, out strip, row tables} behind `base`, tensor-memory allocation and the weight image -> TMEM.
// Must be followed by __syncthreads() and tc_fence_after().
struct TcCtx {
    unsigned char* arena;
    float* outs;
    const float* ctab;
    int* row_action;
    unsigned char* row_active;
    int* path;
    uint32_t tmem_base, d_group, mbar, parity;
    int gw;
};
template <int G, int GTHREADS, int PATH>
__device__ __forceinline__ void tc_setup(TcCtx& c, const FastParams& p, const TcSched& ts, float* base, int tid, int grp, int row) {
    constexpr int THREADS = G * GTHREADS;
    const int ctab_n = (ts.ctab_floats + 3) & ~3;
    unsigned long long* mbars = reinterpret_cast<unsigned long long*>(base + ctab_n);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbars + 4);
    extern __shared__ __align__(16) float smem_base_[];
    const int goff = ((int)(reinterpret_cast<unsigned char*>(tmem_slot + 4) - reinterpret_cast<unsigned char*>(smem_base_)) + 127) & ~127;
    unsigned char* gbase = reinterpret_cast<unsigned char*>(smem_base_) + goff;
    const int gstride = ts.arena_bytes + (GT * ts.out_pitch + GT + GT / 4 + GT * PATH) * 4;
    c.ctab = base;
    c.arena = gbase + grp * gstride;
    c.outs = reinterpret_cast<float*>(c.arena + ts.arena_bytes);
    c.row_action = reinterpret_cast<int*>(c.outs + GT * ts.out_pitch);
    c.row_active = reinterpret_cast<unsigned char*>(c.row_action + GT);
    c.path = c.row_action + GT + GT / 4 + row * PATH;
    __builtin_assume(__isShared(c.arena));
    __builtin_assume(__isShared(c.outs));
    __builtin_assume(__isShared(c.row_action));
    __builtin_assume(__isShared(c.ctab));
    const int warp_u = __shfl_sync(FULL, tid >> 5, 0);  // provably warp-uniform
    const int grp_u = warp_u / (GTHREADS / 32);
    c.gw = warp_u % (GTHREADS / 32);
    c.parity = 0;
    for (int i = tid; i < ts.ctab_floats; i += THREADS) base[i] = p.tc_ctab[i];
    if (tid < G) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(mbars + tid)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp_u == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TC_TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    c.tmem_base = *tmem_slot;
    c.mbar = smem_u32(mbars + grp_u);
    c.d_group = c.tmem_base + ts.d_col0 + grp_u * ts.d_cols_group;
    const uint32_t q = (uint32_t)(warp_u & 3);
    const float* img = p.tc_img + 32 * q + (tid & 31);
    for (int c0 = 8 * (warp_u >> 2); c0 < ts.n_cols; c0 += 8 * (THREADS / 128)) {
        uint32_t r[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) r[i] = __float_as_uint(__ldg(img + (size_t)(c0 + i) * 128));
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(c.tmem_base + ((32u * q) << 16) + c0),
                     "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                     : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    tc_fence_before();
}
template <int LPT>
__device__ __forceinline__ void tc_gather_hidden(const TcSched& ts, unsigned char* arena, const float* __restrict__ hsrc, int H, int row, int j) {
    const float4* src = reinterpret_cast<const float4*>(hsrc);
    unsigned char* d = arena + ts.xh_off + (row >> 3) * ts.xh_sbo + (row & 7) * 16;
    for (int i = j; i < (H >> 2); i += LPT) {
        const float4 h4 = src[i];
        *reinterpret_cast<float4*>(d + i * TC_LBO) = h4;
        *reinterpret_cast<float4*>(d + i * TC_LBO + ts.b_lo) = make_float4(tf32_lo(h4.x), tf32_lo(h4.y), tf32_lo(h4.z), tf32_lo(h4.w));
    }
}

// ------------------------------------------------------- chance-node searches ----
// OR of a predicate over the threads of a group (named barrier reduction)
template <int GTHREADS>
__device__ __forceinline__ bool group_any(int id, bool pred) {
    int out;
    asm volatile(
        "{ .reg .pred q, r; setp.ne.s32 q, %2, 0; barrier.red.or.pred r, %1, %3, q; selp.s32 %0, 1, 0, r; }"
        : "=r"(out) : "r"(id), "r"((int)pred), "n"(GTHREADS) : "memory");
    return out != 0;
}

// Exact pUCT arg-max of a StateNode level of the chance-node searches (self_play_stochastic.py:471-508,
// self_play_risk_sensitive.py:479-520) in float64; same contract as select_exact.
template <int VARIANT>
__device__ __noinline__ int select_exact_v(double pbN, double sqN, int vis, double vs, double prior, float rw,
                                           double discount, double mn, double mx, bool on, int gl0,
                                           const unsigned char* tie_row, int T, int& k_tie) {
    MinMax mm;
    mm.mn = mn;
    mm.mx = mx;
    double score = -CUDART_INF;
    if (on) score = ucb_score<VARIANT>(pbN, sqN, vis, vs, prior, rw, discount, mm);
    double best = score;
#pragma unroll
    for (int d = 1; d < GW; d <<= 1) {
        const double o = shfl_xor_d(FULL, best, d);
        best = o > best ? o : best;
    }
    best = shfl_d(FULL, best, gl0);
    const unsigned tied = (__ballot_sync(FULL, on && score == best) >> gl0) & 0xFFu;
    const int nt = __popc(tied);
    if (nt <= 1) return __ffs(tied) - 1;
    int v = 0;
    if (tie_row) v = tie_row[k_tie % T];
    k_tie++;
    return __fns(tied, 0, (v % nt) + 1);
}

// Stochastic (VARIANT 1, self_play_stochastic.py:329-418) and risk-sensitive (VARIANT 2,
// self_play_risk_sensitive.py:329-537) searches on the group / unit structure of search_fast_kernel:
// the same float32-filtered selection at StateNode levels, chance outcomes from the external stream at
// TransitionNode levels, two unit schedules (one future of a transition expansion, evaluated once per future;
// the prediction of a state expansion), the reference's backups walked by one lane per tree.
// TC: both networks on tcgen05 with the weights resident in tensor memory (tsp_tc.cuh): stages [0, 4) of the TcSched are
// one future of a transition expansion, stages [4, 6) the prediction of a state expansion.
template <int G, int LPT, int VARIANT, bool TC, bool MERGED = false>
__device__ __forceinline__ void search_fast_variant_body(const FastParams& p, const TcSched* __restrict__ tsp_) {
    extern __shared__ __align__(16) float smem[];
    constexpr int GTHREADS = GT * LPT;
    constexpr int THREADS = G * GTHREADS;
    constexpr int NC = 5;
    using Block = NodeBlock<NC>;
    static_assert(VARIANT == 1 || VARIANT == 2, "chance-node searches");
    FastSched* fsa = reinterpret_cast<FastSched*>(smem);
    FastSched* fsb = fsa + 1;
    const int pb_n = (p.S_tab + 2 + 1) & ~1;
    const int pbf_n = (pb_n + 3) & ~3;
    double* pb = reinterpret_cast<double*>(smem + (TC ? 0 : 2 * sizeof(FastSched) / 4));  // TC: no unit schedules in shared memory
    double* sq = pb + pb_n;
    float* pbf = reinterpret_cast<float*>(sq + pb_n);
    float* wsm = pbf + pbf_n;
    const int tid = threadIdx.x;
    if (!TC) {
        const int4* sa = reinterpret_cast<const int4*>(p.sched);
        const int4* sb = reinterpret_cast<const int4*>(p.sched_b);
        int4* da = reinterpret_cast<int4*>(fsa);
        int4* db = reinterpret_cast<int4*>(fsb);
        for (int i = tid; i < (int)(sizeof(FastSched) / 16); i += THREADS) {
            da[i] = sa[i];
            db[i] = sb[i];
        }
    }
    for (int i = tid; i < p.S_tab + 2; i += THREADS) {
        const double a = p.pb_table[i], c = p.pb_table[p.pb_stride + i];
        pb[i] = a;
        sq[i] = c;
        pbf[i] = (float)(a * c);
    }
    const bool fed = !TC && p.fed != nullptr;
    __syncthreads();
    const FastSched& fA = *fsa;
    const FastSched& fB = *fsb;
    const int rs = TC ? 0 : fA.row_stride;  // the host gives both schedules the same strip geometry
    const int wlo_off = (p.w_lo && !TC) ? fA.w_floats : 0;
    const int w_total = TC ? 0 : fA.w_floats + wlo_off;
    if (!fed && !TC) {
        const float4* src = reinterpret_cast<const float4*>(p.wfrag);
        float4* dst = reinterpret_cast<float4*>(wsm);
        for (int i = tid; i < (w_total >> 2); i += THREADS) dst[i] = __ldg(src + i);
    }
    const int grp = tid / GTHREADS, gt = tid % GTHREADS;
    const int row = gt / LPT;
    const int barid = 1 + grp;
    float* act = nullptr;
    float* rowp = nullptr;
    int* row_action;
    unsigned char* row_active;
    const TcSched& ts = *tsp_;
    [[maybe_unused]] TcCtx tc;
    [[maybe_unused]] float4 hq[2];
    [[maybe_unused]] bool tc_ok = true;
    if constexpr (TC) {
        tc_setup<G, GTHREADS, 0>(tc, p, ts, wsm, tid, grp, row);  // (these searches record no path)
        row_action = tc.row_action;
        row_active = tc.row_active;
    } else {
        act = wsm + w_total + grp * (GT * rs + GT + GT + GT * PATH_LEVELS);
        row_action = reinterpret_cast<int*>(act + GT * rs);
        row_active = reinterpret_cast<unsigned char*>(row_action + GT);
        rowp = act + (row & 7) * rs + (row >> 3);
    }

    const int lane = tid & 31, j = tid % LPT;
    const int gl0 = lane & (32 - LPT);
    const int cta = blockIdx.x;
    const bool grp_on = grp < p.grp_base + (cta < p.grp_extra ? 1 : 0);
    const int b = (cta * p.grp_base + min(cta, p.grp_extra) + grp) * GT + row;
    const bool valid = grp_on && b < p.B;  // (an unused group slot has no active tree: it leaves the loop at once)
    const int A = p.A, H = p.H, nf = p.nf;
    const double discount = p.discount;

    Block* blocks = reinterpret_cast<Block*>(p.blocks) + (size_t)(valid ? b : 0) * p.blocks_per_tree;
    float* hpool = p.hidden_pool + (size_t)(valid ? b : 0) * p.hidden_per_tree * H;
    asm volatile("" : "+l"(blocks), "+l"(hpool));
    __builtin_assume(__isGlobal(blocks));
    __builtin_assume(__isGlobal(hpool));
    const float* fed_rec = fed ? p.fed + (size_t)(valid ? b : 0) * p.max_rec * REC : nullptr;
    float* trace_rec = p.trace ? p.trace + (size_t)(valid ? b : 0) * p.max_rec * REC : nullptr;
    const unsigned char* chance_row = p.chance ? p.chance + (size_t)(valid ? b : 0) * p.K : nullptr;

    MinMax mm;
    mm.mn = CUDART_INF;
    mm.mx = -CUDART_INF;
    double root_vsum = 0.0, root_prior = 0.0;
    int root_visit = 0, n_blocks = 1, n_tgroups = 0, max_depth = 0, n_sims = 0, k_tie = 0, k_chance = 0, n_rec = 0;
    int n_root = 0, root_act = 0, depth_total = 0, n_exact = 0;
    if (valid) {
        const TreeHeader h = p.hdr[b];
        n_root = h.n_root;
        if (j < GW) {
            root_prior = p.rootprior[(size_t)b * GW + j];
            root_act = p.rootact[(size_t)b * GW + j];
        }
        if (p.resume) {
            root_vsum = h.root_vsum;
            root_visit = h.root_visit;
            n_blocks = h.n_blocks;
            n_tgroups = h.n_tgroups;
            if (p.add_noise && j < n_root) {
                root_prior = __dadd_rn(__dmul_rn(root_prior, p.one_minus_frac), __dmul_rn(p.noise[(size_t)b * A + j], p.frac));
                p.rootprior[(size_t)b * GW + j] = root_prior;
            }
        }
    }
    const float root_prior_f = (float)root_prior;
    __syncthreads();
    if constexpr (TC) tc_fence_after();

    for (int it = 0; it < p.max_iters; ++it) {
        if (fed && n_rec >= p.max_rec) n_sims = p.S;
        const bool active = valid && n_sims < p.S;
        if (!group_any<GTHREADS>(barid, active)) break;
        [[maybe_unused]] long long tc0 = 0, tc1 = 0, tc2 = 0, tc3 = 0;
        if (TC && p.phase_cycles && tid == 0) tc0 = clock64();

        // ================= selection: StateNode levels pUCT, TransitionNode levels chance =================
        int node = 0, slot = 0, depth = 0, Np = root_visit;
        bool parent_is_state = true;  // kind of the node that owns block `node`
        {
            float mnf = 0.0f, idf = 1.0f;
            if (mm.mx > mm.mn) {
                mnf = (float)mm.mn;
                idf = rcp_approx((float)__dsub_rn(mm.mx, mm.mn));
            }
            const float amnf = fabsf(mnf);
            bool walking = active;
            while (__any_sync(FULL, walking)) {
                const Block* blk = blocks + node;
                const int nch = (node == 0) ? n_root : A;
                const bool state_lvl = walking && parent_is_state;
                const bool on = state_lvl && j < nch;
                const int jc = j < NC ? j : 0;
                const int vis = blk->visit[jc];
                const int ch = blk->child[jc];
                int sl = 0;
                // (the trees of a warp descend in step -- state levels on even rounds, transition levels on odd ones -- so a
                // transition round skips the whole pUCT evaluation)
                if (__any_sync(FULL, state_lvl)) {
                const double vs = blk->vsum[jc];
                const float rw = blk->reward[jc];
                const float prf = (node == 0) ? root_prior_f : blk->prior[jc];
                const float ps = pbf[Np] * rcp_approx((float)(vis + 1)) * prf;
                float V, aux;
                if (VARIANT == 1) {  // normalize(child.value()), self_play_stochastic.py:500-503
                    const float q = (float)vs * rcp_approx((float)vis);
                    V = (q - mnf) * idf;
                    aux = idf * (fabsf(q) + amnf);
                } else {  // normalize(child.reward + discount * child.value), self_play_risk_sensitive.py:511-516
                    const float gq = p.discount_f * (float)vs;
                    V = ((gq + rw) - mnf) * idf;
                    aux = idf * (fabsf(gq) + fabsf(rw) + amnf);
                }
                if (vis <= 0) {
                    V = 0.0f;
                    aux = 0.0f;
                }
                const float sc = ps + V;
                const float E = fmaf(1.9073486e-6f, fabsf(ps) + fabsf(V) + aux, 1e-30f);
                const float lo = on ? sc - E : -CUDART_INF_F;
                const float hi = on ? sc + E : -CUDART_INF_F;
                float m1 = lo;
#pragma unroll
                for (int d = 1; d < GW; d <<= 1) m1 = fmaxf(m1, __shfl_xor_sync(FULL, m1, d));
                m1 = __shfl_sync(FULL, m1, gl0);
                unsigned cand = (__ballot_sync(FULL, on && hi >= m1) >> gl0) & 0xFFu;
                {   // a non-finite score must not leave a lone candidate behind (NaN drops out of fmaxf / >=): float64 decides
                    const unsigned bad = __ballot_sync(FULL, on && !(fabsf(sc) <= 3.0e38f));
                    if ((bad >> gl0) & 0xFFu) cand = 0;
                }
                sl = __ffs(cand) - 1;
                const bool need = state_lvl && __popc(cand) != 1;
                if (__any_sync(FULL, need)) {
                    int k2 = k_tie;
                    const int se = select_exact_v<VARIANT>(pb[Np], sq[Np], vis, vs, (node == 0) ? root_prior : (double)prf, rw,
                                                           discount, mm.mn, mm.mx, on && need, gl0,
                                                           p.tie ? p.tie + (size_t)b * p.T : nullptr, p.T, k2);
                    if (need) {
                        sl = se;
                        k_tie = k2;
                        n_exact++;
                    }
                }
                }
                if (walking && !parent_is_state) {  // TransitionNode.select_child, self_play_stochastic.py:517-533
                    sl = 0;
                    if (nf > 1) {
                        const int v = chance_row ? chance_row[k_chance % p.K] : 0;
                        k_chance++;
                        sl = v % nf;
                    }
                }
                const int vis_sel = __shfl_sync(FULL, vis, (gl0 + sl) & 31);
                const int child_sel = __shfl_sync(FULL, ch, (gl0 + sl) & 31);
                if (walking) {
                    slot = sl;
                    depth++;
                    if (child_sel < 0) {
                        walking = false;
                    } else {
                        node = child_sel;
                        Np = vis_sel;
                        parent_is_state = !parent_is_state;
                    }
                }
            }
        }
        // leaf = child `slot` of block `node`
        const bool leaf_is_state = !parent_is_state;
        const int kind = !active ? 0 : (leaf_is_state ? 1 : 2);
        const int ract = __shfl_sync(FULL, root_act, (gl0 + slot) & 31);
        int action = 0, hslot = 0;
        if (active) {
            const int aux_node = blocks[node].aux >> 8;
            if (!leaf_is_state) {
                action = (node == 0) ? ract : slot;
                hslot = (node == 0) ? 0 : aux_node;
            } else {
                hslot = 1 + aux_node * nf + slot;
            }
            if (depth > max_depth) max_depth = depth;
            depth_total += depth;
        }

        // ================= network evaluation =================
        float value = 0.0f, prior = 0.0f;    // state-leaf results
        float tprior = 0.0f, treward = 0.0f;  // transition-leaf results (lane z)
        if constexpr (TC) {
            if (j == 0) {
                row_action[row] = action;
                row_active[row] = active ? 1 : 0;
            }
            if (p.phase_cycles && tid == 0) tc1 = clock64();
            if (active) tc_gather_hidden<LPT>(ts, tc.arena, hpool + (size_t)hslot * H, H, row, j);
            fence_async_smem();
            const bool any_t = group_any<GTHREADS>(barid, kind == 2);  // (also orders the gather before the MMAs)
            const bool any_s = MERGED ? true : group_any<GTHREADS>(barid, kind == 1);
            if (p.phase_cycles && tid == 0) tc2 = clock64();
            const int lane_ = tid & 31;
            const float* o = tc.outs + row * ts.out_pitch;
            if constexpr (MERGED) {
                // both futures of a transition expansion AND the prediction of a state expansion in one pass of three stages
                // (the trees of a group are rarely all of one kind; what a tree does not need is computed and dropped)
                float* hst = (kind == 2) ? hpool + (size_t)(1 + n_tgroups * nf) * H : nullptr;
                tc_ok = tc_mlp<LPT, 2>(ts, tc.tmem_base, tc.d_group, tc.arena, tc.outs, tc.ctab, row_action, tc.mbar, tc.parity, tc.gw,
                                    lane_, row, j, barid, hq, p.phase_cycles, 0, -1, 0, hst) && tc_ok;
                float v1, rw0, p1, dv, rw1, tp;
                decode_row<LPT, true>(o + ts.off_vl, o + ts.off_rl, o + ts.off_pl, p.support, A, false, j, v1, rw0, p1);
                decode_row<LPT, true>(o + ts.off_rl2, o + ts.off_rl2, o + ts.off_tp, p.support, nf, false, j, dv, rw1, tp);
                if (kind == 2) {
                    treward = (j == 0) ? rw0 : (j == 1 ? rw1 : 0.0f);
                    tprior = tp;  // softmax of the prior logits, self_play_stochastic.py:553-554
                } else if (kind == 1) {
                    value = v1;
                    prior = p1;
                }
            } else {
            if (any_t) {
                for (int z = 0; z < nf; ++z) {
                    tc_ok = tc_mlp<LPT, 1>(ts, tc.tmem_base, tc.d_group, tc.arena, tc.outs, tc.ctab, row_action, tc.mbar, tc.parity, tc.gw,
                                        lane_, row, j, barid, hq, nullptr, 0, 4, z) && tc_ok;
                    float dv, rwz, tp;
                    decode_row<LPT, true>(o + ts.off_rl, o + ts.off_rl, o + ts.off_tp, p.support, nf, false, j, dv, rwz, tp);
                    if (kind == 2) {
                        if (j == z) treward = rwz;
                        if (z == 0) tprior = tp;  // softmax of the prior logits, self_play_stochastic.py:553-554
                        float4* dst = reinterpret_cast<float4*>(hpool + (size_t)(1 + n_tgroups * nf + z) * H);
#pragma unroll
                        for (int q = 0; q < 2; ++q)
                            if (j + q * LPT < (H >> 2)) dst[j + q * LPT] = hq[q];
                    }
                }
            }
            if (any_s) {
                tc_ok = tc_mlp<LPT, 1>(ts, tc.tmem_base, tc.d_group, tc.arena, tc.outs, tc.ctab, row_action, tc.mbar, tc.parity, tc.gw,
                                    lane_, row, j, barid, hq, nullptr, 4, 6, 0) && tc_ok;
                float dr, v1, p1;
                decode_row<LPT, true>(o + ts.off_vl, o + ts.off_vl, o + ts.off_pl, p.support, A, false, j, v1, dr, p1);
                if (kind == 1) {
                    value = v1;
                    prior = p1;
                }
            }
            }
        } else if (!fed) {
            if (j == 0) {
                row_action[row] = action;
                row_active[row] = active ? 1 : 0;
            }
            if (active) {
                const float4* src = reinterpret_cast<const float4*>(hpool + (size_t)hslot * H);
                for (int i = j; i < (H >> 2); i += LPT) {
                    const float4 h4 = src[i];
                    float* d = rowp + 8 * i;
                    d[0] = h4.x; d[2] = h4.y; d[4] = h4.z; d[6] = h4.w;
                    d += 8 * rs;
                    d[0] = tf32_lo(h4.x); d[2] = tf32_lo(h4.y); d[4] = tf32_lo(h4.z); d[6] = tf32_lo(h4.w);
                }
            }
            const bool any_t = group_any<GTHREADS>(barid, kind == 2);  // (also orders the gather before the tiles)
            const bool any_s = group_any<GTHREADS>(barid, kind == 1);
            if (any_t) {
                for (int z = 0; z < nf; ++z) {
                    fast_mlp<LPT>(fA, wsm, act, row_action, row_active, gt, barid, wlo_off, nullptr, z);
                    float dv, rwz, tp;
                    decode_row<LPT>(rowp + 2 * fA.off_rl, rowp + 2 * fA.off_rl, rowp + 2 * fA.off_tp, p.support, nf, false, j,
                                    dv, rwz, tp);
                    if (kind == 2) {
                        if (j == z) treward = rwz;
                        if (z == 0) tprior = tp;  // softmax of the prior logits, self_play_stochastic.py:553-554
                        float4* dst = reinterpret_cast<float4*>(hpool + (size_t)(1 + n_tgroups * nf + z) * H);
                        const float* src = rowp + 2 * fA.off_hidden;
                        for (int i = j; i < (H >> 2); i += LPT) {
                            const float* q = src + 8 * i;
                            dst[i] = make_float4(q[0], q[2], q[4], q[6]);
                        }
                    }
                    group_barrier<GTHREADS>(barid);  // the strip is read above and rewritten by the next pass
                }
            }
            if (any_s) {
                fast_mlp<LPT>(fB, wsm, act, row_action, row_active, gt, barid, wlo_off, nullptr, 0);
                float dr, v1, p1;
                decode_row<LPT>(rowp + 2 * fB.off_vl, rowp + 2 * fB.off_vl, rowp + 2 * fB.off_pl, p.support, A, false, j, v1, dr, p1);
                if (kind == 1) {
                    value = v1;
                    prior = p1;
                }
            }
        } else if (active) {
            const float* fr = fed_rec + (size_t)n_rec * REC;
            if (kind == 1) {
                value = fr[1];
                prior = (j < A) ? fr[4 + j] : 0.0f;
            } else {
                tprior = (j < nf) ? fr[1 + j] : 0.0f;
                treward = (j < nf) ? fr[1 + nf + j] : 0.0f;
            }
        }

        if (TC && p.phase_cycles && tid == 0) tc3 = clock64();
        // ================= expansion + backup =================
        {
#pragma unroll
            for (int q = 0; q < REC / LPT; ++q) {  // evaluation record (tsp.h)
                const int k = j + q * LPT;
                const float pk = __shfl_sync(FULL, prior, (gl0 + ((k - 4) & (LPT - 1))) & 31);
                const float tpk = __shfl_sync(FULL, tprior, (gl0 + ((k - 1) & (LPT - 1))) & 31);
                const float trk = __shfl_sync(FULL, treward, (gl0 + ((k - 1 - nf) & (LPT - 1))) & 31);
                if (active && trace_rec && n_rec < p.max_rec) {
                    float tv = 0.0f;
                    if (kind == 1) {
                        if (k == 0) tv = 1.0f;
                        if (k == 1) tv = value;
                        if (k >= 4 && k < 4 + A) tv = pk;
                    } else {
                        if (k == 0) tv = 2.0f;
                        if (k >= 1 && k < 1 + nf) tv = tpk;
                        if (k >= 1 + nf && k < 1 + 2 * nf) tv = trk;
                    }
                    trace_rec[(size_t)n_rec * REC + k] = tv;
                }
            }
            if (active) n_rec++;
            Block* pblk = blocks + node;
            if (kind == 2) {
                // ---- TransitionNode.expand, self_play_stochastic.py:535-557 (no backup, no sim count) ----
                const int nb = n_blocks++;
                Block* nblk = blocks + nb;
                if (j < NC) {
                    nblk->visit[j] = 0;
                    nblk->child[j] = -1;
                    nblk->prior[j] = tprior;
                    nblk->reward[j] = (j < nf) ? treward : 0.0f;
                    nblk->vsum[j] = 0.0;
                }
                if (j == 0) {
                    nblk->parent = node;
                    nblk->aux = slot | (n_tgroups << 8);
                    pblk->child[slot] = nb;
                }
                n_tgroups++;
            } else if (kind == 1) {
                // ---- StateNode.expand (self_play_stochastic.py:455-469) ----
                const int nb = n_blocks++;
                Block* nblk = blocks + nb;
                if (j < NC) {
                    nblk->visit[j] = 0;
                    nblk->child[j] = -1;
                    nblk->prior[j] = prior;
                    nblk->reward[j] = 0.0f;
                    nblk->vsum[j] = 0.0;
                }
                if (j == 0) {
                    nblk->parent = node;
                    nblk->aux = slot | (hslot << 8);
                    pblk->child[slot] = nb;
                }
            }
            __syncwarp();
            // ---- backup by one lane per tree (the walk is a dependent chain; no lane needs another's view) ----
            if (kind == 1 && j == 0) {
                double v = (double)value;
                int cur = node, cs = slot;
                if (VARIANT == 1) {
                    // self_play_stochastic.py:408-418; entries alternate State (leaf), Transition, State, ...
                    bool is_state = true;
                    while (true) {
                        Block* bk = blocks + cur;
                        const double vs = __dadd_rn(bk->vsum[cs], v);
                        const int vis = bk->visit[cs] + 1;
                        const double rw = (double)bk->reward[cs];
                        bk->vsum[cs] = vs;
                        bk->visit[cs] = vis;
                        if (is_state) {
                            v = __dadd_rn(rw, __dmul_rn(discount, v));
                            mm.update(__dadd_rn(rw, __dmul_rn(discount, __ddiv_rn(vs, (double)vis))));
                        }
                        is_state = !is_state;
                        if (cur == 0) break;
                        cs = bk->aux & 0xFF;
                        cur = bk->parent;
                    }
                    root_vsum = __dadd_rn(root_vsum, v);
                    root_visit += 1;
                    mm.update(__dadd_rn(0.0, __dmul_rn(discount, __ddiv_rn(root_vsum, (double)root_visit))));
                } else {
                    // self_play_risk_sensitive.py:408-420, 452-461, 529-537
                    pblk->vsum[slot] = v;  // search_path[-1].value = value
                    pblk->visit[slot] = pblk->visit[slot] + 1;
                    bool is_transition = true;  // kind of the node that OWNS block `cur`
                    while (true) {
                        Block* bk = blocks + cur;
                        double nv;
                        if (is_transition) {
                            double best = 0.0;
                            bool have = false;
                            for (int k = 0; k < nf; ++k) {
                                if (bk->visit[k] > 0) {
                                    const double t = __dadd_rn((double)bk->reward[k], __dmul_rn(discount, bk->vsum[k]));
                                    if (!have || t < best) best = t;
                                    have = true;
                                }
                            }
                            nv = best;
                        } else {
                            const int nch = (cur == 0) ? n_root : A;
                            double sum = 0.0;
                            int counts = 0;
                            for (int k = 0; k < nch; ++k) {
                                const int c = bk->visit[k];
                                sum = __dadd_rn(sum, __dmul_rn((double)c, bk->vsum[k]));
                                counts += c;
                            }
                            nv = __ddiv_rn(sum, (double)counts);
                        }
                        if (cur == 0) {
                            root_vsum = nv;
                            root_visit += 1;
                            mm.update(__dadd_rn(0.0, __dmul_rn(discount, nv)));
                            break;
                        }
                        const int ps = bk->aux & 0xFF, pc = bk->parent;
                        Block* pb2 = blocks + pc;
                        const int vis = pb2->visit[ps] + 1;
                        const double rw = (double)pb2->reward[ps];
                        pb2->vsum[ps] = nv;
                        pb2->visit[ps] = vis;
                        if (!is_transition) mm.update(__dadd_rn(rw, __dmul_rn(discount, nv)));
                        is_transition = !is_transition;
                        cur = pc;
                    }
                }
            }
            __syncwarp();
            // lane 0 of the tree carries the updated statistics: hand them to the other lanes
            mm.mn = shfl_d(FULL, mm.mn, gl0);
            mm.mx = shfl_d(FULL, mm.mx, gl0);
            root_vsum = shfl_d(FULL, root_vsum, gl0);
            root_visit = __shfl_sync(FULL, root_visit, gl0);
            if (kind == 1) n_sims++;
        }
        __syncwarp();
        if (TC && p.phase_cycles && tid == 0) {
            const long long tc4 = clock64();
            atomicAdd(p.phase_cycles + 0, (unsigned long long)(tc1 - tc0));
            atomicAdd(p.phase_cycles + 1, (unsigned long long)(tc2 - tc1));
            atomicAdd(p.phase_cycles + 2, (unsigned long long)(tc3 - tc2));
            atomicAdd(p.phase_cycles + 3, (unsigned long long)(tc4 - tc3));
            atomicAdd(p.phase_cycles + 4, 1ull);
        }
    }

    if (valid) {
        if (j == 0) {
            if (p.root_value)
                p.root_value[b] = (VARIANT == 2) ? root_vsum : (root_visit == 0 ? 0.0 : __ddiv_rn(root_vsum, (double)root_visit));
            if (p.max_depth) p.max_depth[b] = max_depth;
            if (p.num_records) p.num_records[b] = n_rec;
            if (p.total_depth) p.total_depth[b] = depth_total;
            p.hdr[b].root_vsum = root_vsum;
            p.hdr[b].root_visit = root_visit;
            p.hdr[b].n_blocks = n_blocks;
            p.hdr[b].n_tgroups = n_tgroups;
        }
        if (j < A) {
            if (p.visit_counts) p.visit_counts[(size_t)b * A + j] = 0;
            if (p.child_value) p.child_value[(size_t)b * A + j] = 0.0;
            if (p.child_prior) p.child_prior[(size_t)b * A + j] = 0.0;
            if (p.child_reward) p.child_reward[(size_t)b * A + j] = 0.0;
        }
    }
    // (outside the branch: with an odd number of roots a warp holds one valid and one invalid tree, and the lanes of
    // the invalid one do not exit here when the kernel still has to release its tensor memory)
    __syncwarp();
    if (valid) {
        const Block* rb = blocks;
        if (j < n_root) {
            const int a = root_act;
            const int vis = rb->visit[j];
            const double vs = rb->vsum[j];
            if (p.visit_counts) p.visit_counts[(size_t)b * A + a] = vis;
            if (p.child_value)
                p.child_value[(size_t)b * A + a] = (VARIANT == 2) ? vs : (vis == 0 ? 0.0 : __ddiv_rn(vs, (double)vis));
            if (p.child_prior) p.child_prior[(size_t)b * A + a] = root_prior;
            if (p.child_reward) p.child_reward[(size_t)b * A + a] = (double)rb->reward[j];
        }
    }
    if constexpr (TC) {
        if (!tc_ok && p.tc_status) atomicExch(p.tc_status, 1);
        tc_fence_before();
        __syncthreads();
        if (__shfl_sync(FULL, tid >> 5, 0) == 0)
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tc.tmem_base), "n"(TC_TMEM_COLS));
    }
}

template <int G, int LPT, int VARIANT>
__global__ void __launch_bounds__(G * GT * LPT, 512 / (G * GT * LPT)) search_fast_variant_kernel(const __grid_constant__ FastParams p) {
    search_fast_variant_body<G, LPT, VARIANT, false>(p, nullptr);
}

template <int G, int LPT, int VARIANT, bool MERGED>
__global__ void __launch_bounds__(G * GT * LPT, 1) search_tc_variant_kernel(const __grid_constant__ FastParams p,
                                                                            const __grid_constant__ TcSched ts) {
    search_fast_variant_body<G, LPT, VARIANT, true, MERGED>(p, &ts);
}

}  // namespace tsp
