// tsp_device.cuh -- device side of the B200 batched MuZero tree-search engine (sm_100a).
//
// Data layout in HBM (per engine, sized at tsp_create):
//   node pool      [max_roots][blocks_per_tree] NodeBlock<NC>   one 128-byte (NC=5) or 256-byte (NC=8)
//                  block per EXPANDED node holding the statistics of all its children
//                  (visit, child index, prior, reward, value sum) + parent link: one cache line
//                  per tree level during selection.
//   hidden pool    [max_roots][hidden_per_tree][H] float32
//   tree headers   [max_roots] TreeHeader (min-max stats, root statistics, counters)
//   root priors    [max_roots][8] float64 (noise-mixed priors stay double, self_play.py:540-549)
//
// Execution model: ONE persistent kernel runs the whole search of TM trees per CTA; a group of
// 8 lanes owns one tree (lane j <-> child j, shuffle arg-max), all threads of the CTA evaluate
// the leaf batch through the fused MLP schedule, then each group expands and backs up its tree.
// Trees are independent, so there is no grid-wide synchronisation and no launch per simulation.
//
// Tree arithmetic is IEEE double in exactly the reference's operation order (explicit _rn
// intrinsics so that nvcc never contracts a*b+c into an FMA); network arithmetic is float32.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

namespace tsp {

constexpr int GW = 8;    // lanes per tree
constexpr int REC = 16;  // floats per evaluation record (TSP_REC_WIDTH)
constexpr int MAX_STAGES = 14;
constexpr int MAX_OPS = 4;

template <int NC>
struct alignas(NC <= 5 ? 128 : 256) NodeBlock {
    int visit[NC];     // child.visit_count
    int child[NC];     // block id of the child once expanded, -1 before
    float prior[NC];   // child.prior (float32 softmax output; the root's double priors live apart)
    float reward[NC];  // child.reward (set when the child is expanded)
    int parent;        // block id of the parent (-1 for the root)
    int aux;           // bits 0..7: slot of this node in its parent; bits 8..31: hidden slot
                       // (state node) or transition ordinal (transition node)
    double vsum[NC];   // child.value_sum  (risk-sensitive: child.value)
};
static_assert(sizeof(NodeBlock<5>) == 128, "NodeBlock<5> must be one 128-byte line");
static_assert(sizeof(NodeBlock<8>) == 256, "NodeBlock<8> must be two 128-byte lines");

struct alignas(16) TreeHeader {
    double mm_min, mm_max;  // MinMaxStats, self_play.py:661-678
    double root_vsum;       // root.value_sum (risk-sensitive: root.value)
    int root_visit;         // root.visit_count
    int n_blocks;           // expanded nodes so far
    int n_tgroups;          // expanded transition nodes so far (stochastic variants)
    int max_depth;
    int n_sims;
    int k_tie, k_chance;    // stream cursors
    int n_rec;              // network evaluations so far
    float root_pred;
    int n_root;             // number of root children (legal actions)
};
static_assert(sizeof(TreeHeader) == 64, "TreeHeader must be 64 bytes");

// ------------------------------------------------------------------ MLP schedule ----
struct LayerOp {
    int w_off;  // float offset of Wt [K_total][Npad] (transposed, zero padded) in the weight arena
    int b_off;  // float offset of the bias [Npad]
    short K;    // dense input width (one-hot rows follow at K, K+1, ...)
    short N, Npad;
    short in_off, out_off;  // float offsets inside a row's activation strip
    unsigned char act;      // 0 identity, 1 ELU
    unsigned char extra;    // 0 none, 1 + Wt[K + action], 2 + Wt[K + action] + Wt[K + A + z]
    unsigned char ct;       // SIMT path: output columns per thread (4, 8 or 16; divides Npad), set per launch
    unsigned char ng;       // tensor-core path: adjacent 8-column tiles per warp unit (1..4), set per launch
    short wld;              // leading dimension of Wt in floats (== 8 or 24 mod 32)
    short units;            // tensor-core path: warp units of this op = (TM / 16) * groups, set per launch
    short groups;           // ceil((Npad / 8) / ng)
    short pad2;
};
struct Stage {
    unsigned char n_ops;
    unsigned char row_op;  // 0 none, 1 min-max normalise r_in -> r_out over r_n elements
    short r_in, r_out, r_n;
    LayerOp ops[MAX_OPS];
};
struct Schedule {
    int n_stages;
    int row_stride;  // floats per row strip, == 4 (mod 32): conflict-free LDS.128 across rows
    int in_dim;
    short off_in, off_hidden, off_pl, off_vl, off_rl, off_tl, off_tp, pad0;
    int f_stride;    // MuZeroStochastic (one dynamics MLP per future, models.py:334-345): the weights of the transition
                     // expansion of future z start z * f_stride floats further on (same layout per future); else 0
    int pad1[3];
    Stage st[MAX_STAGES];
};

// ELU(alpha = 1): exp(v) - 1 for v <= 0. ex2.approx keeps the absolute error below 5e-7 on (-inf, 0],
// two orders of magnitude inside the 1e-5 network tolerance, at 3 instructions instead of ~25 (expm1f).
__device__ __forceinline__ float elu(float v) {
    const float e = __fsub_rn(__expf(fminf(v, 0.0f)), 1.0f);
    return v > 0.0f ? v : e;
}

// Weights come either from the shared-memory copy (plain loads: the pointer is derived from the
// extern __shared__ array by offsets only, so the compiler keeps the shared address space and emits
// LDS.128 that it is free to batch ahead of the FMAs) or from global memory (read-only path).
template <bool WSMEM>
__device__ __forceinline__ float4 ldw4(const float4* p) {
    if (WSMEM) return *p;
    return __ldg(p);
}

// One work unit of a Linear layer: ONE row (tree) x CT output columns per thread, lanes of a warp
// holding different rows. The weight reads are then warp-uniform (one shared-memory broadcast per
// LDS.128 instead of four 128-byte phases), the activation reads are one conflict-free LDS.128 per lane
// (row stride == 4 mod 32 words), every lane does useful FMAs, and a thread issues CT*4 FMAs per
// (1 + CT) loads. Measured alternatives (1x4 tiles: shared-memory pipe bound; 4x4 tiles: the work of a
// stage collapses onto one warp) are recorded in DESIGN.md.
template <bool WSMEM, int CT>
__device__ __forceinline__ void dense_unit(const LayerOp& op, const float* __restrict__ W, float* __restrict__ act,
                                           int rs, int row, int c0, int action, bool active, int A, int z) {
    const int npad = op.wld;  // row stride of Wt
    const int K = op.K;
    const float* x = act + row * rs + op.in_off;
    const float* wp = W + op.w_off + c0;  // row k of Wt starts at wp + k * wld
    float acc[CT];
#pragma unroll
    for (int c = 0; c < CT; c += 4) {
        const float4 b4 = ldw4<WSMEM>(reinterpret_cast<const float4*>(W + op.b_off + c0 + c));
        acc[c] = b4.x; acc[c + 1] = b4.y; acc[c + 2] = b4.z; acc[c + 3] = b4.w;
    }
    if (op.extra) {
        const float* we = wp + (K + action) * npad;
#pragma unroll
        for (int c = 0; c < CT; c += 4) {
            const float4 e = ldw4<WSMEM>(reinterpret_cast<const float4*>(we + c));
            acc[c] += e.x; acc[c + 1] += e.y; acc[c + 2] += e.z; acc[c + 3] += e.w;
        }
        if (op.extra == 2) {
            const float* wf = wp + (K + A + z) * npad;
#pragma unroll
            for (int c = 0; c < CT; c += 4) {
                const float4 f = ldw4<WSMEM>(reinterpret_cast<const float4*>(wf + c));
                acc[c] += f.x; acc[c + 1] += f.y; acc[c + 2] += f.z; acc[c + 3] += f.w;
            }
        }
    }
    int k = 0;
    for (; k + 4 <= K; k += 4) {
        const float4 xv = *reinterpret_cast<const float4*>(x + k);
        const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
            for (int c = 0; c < CT; c += 4) {
                const float4 w4 = ldw4<WSMEM>(reinterpret_cast<const float4*>(wp + c));
                acc[c] = fmaf(xs[kk], w4.x, acc[c]);
                acc[c + 1] = fmaf(xs[kk], w4.y, acc[c + 1]);
                acc[c + 2] = fmaf(xs[kk], w4.z, acc[c + 2]);
                acc[c + 3] = fmaf(xs[kk], w4.w, acc[c + 3]);
            }
            wp += npad;
        }
    }
    for (; k < K; ++k) {
        const float xk = x[k];
#pragma unroll
        for (int c = 0; c < CT; c += 4) {
            const float4 w4 = ldw4<WSMEM>(reinterpret_cast<const float4*>(wp + c));
            acc[c] = fmaf(xk, w4.x, acc[c]);
            acc[c + 1] = fmaf(xk, w4.y, acc[c + 1]);
            acc[c + 2] = fmaf(xk, w4.z, acc[c + 2]);
            acc[c + 3] = fmaf(xk, w4.w, acc[c + 3]);
        }
        wp += npad;
    }
    if (!active) return;
    float* y = act + row * rs + op.out_off + c0;
#pragma unroll
    for (int c = 0; c < CT; c += 4) {
        float4 v = make_float4(acc[c], acc[c + 1], acc[c + 2], acc[c + 3]);
        if (op.act) {
            v.x = elu(v.x); v.y = elu(v.y); v.z = elu(v.z); v.w = elu(v.w);
        }
        *reinterpret_cast<float4*>(y + c) = v;
    }
}

// Register tile for weights that stay in global memory / L2 (networks too large for the shared-memory
// copy): RT rows x 4 columns per thread, so that one 16-byte weight load feeds RT * 4 FMAs.
template <int RT>
__device__ __forceinline__ void dense_tile_g(const LayerOp& op, const float* __restrict__ W, float* __restrict__ act,
                                             int rs, int tt, const int* __restrict__ row_action,
                                             const unsigned char* __restrict__ row_active, int A, int z) {
    const int ncgs = op.Npad >> 2;   // column groups of the tile grid
    const int ncg = op.wld >> 2;     // row stride of Wt in float4
    const int rg = tt / ncgs, cg = tt - rg * ncgs;
    const int row0 = rg * RT;
    bool any = false;
#pragma unroll
    for (int r = 0; r < RT; ++r) any |= (row_active[row0 + r] != 0);
    if (!any) return;
    const int K = op.K;
    const float* x = act + row0 * rs + op.in_off;
    const float4* w = reinterpret_cast<const float4*>(W + op.w_off) + cg;
    float4 acc[RT];
    {
        const float4 bias = __ldg(reinterpret_cast<const float4*>(W + op.b_off) + cg);
#pragma unroll
        for (int r = 0; r < RT; ++r) acc[r] = bias;
    }
    if (op.extra) {
#pragma unroll
        for (int r = 0; r < RT; ++r) {
            const float4 e = __ldg(w + (K + row_action[row0 + r]) * ncg);
            acc[r].x += e.x; acc[r].y += e.y; acc[r].z += e.z; acc[r].w += e.w;
        }
        if (op.extra == 2) {
            const float4 f = __ldg(w + (K + A + z) * ncg);
#pragma unroll
            for (int r = 0; r < RT; ++r) {
                acc[r].x += f.x; acc[r].y += f.y; acc[r].z += f.z; acc[r].w += f.w;
            }
        }
    }
    int k = 0;
    for (; k + 4 <= K; k += 4) {
        const float4 w0 = __ldg(w);
        const float4 w1 = __ldg(w + ncg);
        const float4 w2 = __ldg(w + 2 * ncg);
        const float4 w3 = __ldg(w + 3 * ncg);
        w += 4 * ncg;
#pragma unroll
        for (int r = 0; r < RT; ++r) {
            const float4 xv = *reinterpret_cast<const float4*>(x + r * rs + k);
            acc[r].x = fmaf(xv.x, w0.x, acc[r].x); acc[r].y = fmaf(xv.x, w0.y, acc[r].y);
            acc[r].z = fmaf(xv.x, w0.z, acc[r].z); acc[r].w = fmaf(xv.x, w0.w, acc[r].w);
            acc[r].x = fmaf(xv.y, w1.x, acc[r].x); acc[r].y = fmaf(xv.y, w1.y, acc[r].y);
            acc[r].z = fmaf(xv.y, w1.z, acc[r].z); acc[r].w = fmaf(xv.y, w1.w, acc[r].w);
            acc[r].x = fmaf(xv.z, w2.x, acc[r].x); acc[r].y = fmaf(xv.z, w2.y, acc[r].y);
            acc[r].z = fmaf(xv.z, w2.z, acc[r].z); acc[r].w = fmaf(xv.z, w2.w, acc[r].w);
            acc[r].x = fmaf(xv.w, w3.x, acc[r].x); acc[r].y = fmaf(xv.w, w3.y, acc[r].y);
            acc[r].z = fmaf(xv.w, w3.z, acc[r].z); acc[r].w = fmaf(xv.w, w3.w, acc[r].w);
        }
    }
    for (; k < K; ++k) {
        const float4 w0 = __ldg(w);
        w += ncg;
#pragma unroll
        for (int r = 0; r < RT; ++r) {
            const float xv = x[r * rs + k];
            acc[r].x = fmaf(xv, w0.x, acc[r].x); acc[r].y = fmaf(xv, w0.y, acc[r].y);
            acc[r].z = fmaf(xv, w0.z, acc[r].z); acc[r].w = fmaf(xv, w0.w, acc[r].w);
        }
    }
#pragma unroll
    for (int r = 0; r < RT; ++r) {
        if (!row_active[row0 + r]) continue;
        float4 v = acc[r];
        if (op.act) {
            v.x = elu(v.x); v.y = elu(v.y); v.z = elu(v.z); v.w = elu(v.w);
        }
        *reinterpret_cast<float4*>(act + (row0 + r) * rs + op.out_off + cg * 4) = v;
    }
}

// ------------------------------------------------------------- tensor-core MLP tiles ----
// One warp computes a 16 (trees) x 8 (outputs) tile of a Linear layer with mma.sync.m16n8k8 on TF32
// operands, using the 3xTF32 split (x = hi + lo, hi = rna-tf32(x), lo = x - hi;
// x*w ~= hi_x*hi_w + hi_x*lo_w + lo_x*hi_w) accumulated in fp32: the dropped lo*lo term is ~2^-22 relative,
// i.e. float32-level accuracy, which keeps the 1e-5 network parity of the CUDA-core path.
// Why not tcgen05 here: a CTA holds 32 trees, a tcgen05 tile needs M >= 64 rows and a TMEM -> register ->
// swizzled-bf16 shared-memory epilogue after EVERY layer (~500 instructions per thread per stage), more than
// the whole warp-level stage costs (~170). Activations stay float32 in shared memory between layers.
// hi = x with the 13 low mantissa bits cleared (exactly representable in TF32), lo = x - hi (exact in
// fp32, |lo| < 2^-10 |x|). cvt.rna.tf32 would halve |lo| but is emulated with ~6 integer instructions
// on sm_100a (measured: 1.4k of 13.5k instructions per step); the dropped lo*lo term stays < 2^-20.
__device__ __forceinline__ unsigned tf32_hi(float x) { return __float_as_uint(x) & 0xFFFFE000u; }
__device__ __forceinline__ void mma_tf32(float (&d)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <bool WSMEM>
__device__ __forceinline__ float ldw1(const float* p) {
    if (WSMEM) return *p;
    return __ldg(p);
}

// Unit (mt, nt0 .. nt0 + NT - 1): rows [16 mt, 16 mt + 16) x NT adjacent 8-column tiles, one warp.
// The A fragments (activations) are loaded and split once per k-step and reused by the NT tiles; every
// tile has two independent accumulators (hi*hi, and the two correction products), so that consecutive
// mma.sync instructions never wait on each other's ~35-cycle latency.
template <bool WSMEM, int NT>
__device__ __forceinline__ void mma_tiles(const LayerOp& op, const float* __restrict__ W, float* __restrict__ act,
                                          int rs, int mt, int nt0, const int* __restrict__ row_action,
                                          const unsigned char* __restrict__ row_active, int A, int z,
                                          unsigned long long* dbg = nullptr) {
    long long q0 = 0, q1 = 0, q2 = 0;
    if (dbg && threadIdx.x == 0) q0 = clock64();
    const int lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const int r0 = mt * 16 + g, r1 = r0 + 8;
    const int K = op.K, wld = op.wld;
    const bool on0 = row_active[r0] != 0, on1 = row_active[r1] != 0;
    const int ra0 = row_action[r0], ra1 = row_action[r1];
    const float* x0 = act + r0 * rs + op.in_off;
    const float* x1 = act + r1 * rs + op.in_off;
    const float* wb = W + op.w_off + nt0 * 8;  // Wt[k][8 nt0 + n] at wb[k * wld + n]
    float dm[NT][4], dc[NT][4], de[NT][4];  // hi*hi, lo*hi, hi*lo: three independent mma chains per tile
#pragma unroll
    for (int i = 0; i < NT; ++i) {
        dc[i][0] = dc[i][1] = dc[i][2] = dc[i][3] = 0.f;
        de[i][0] = de[i][1] = de[i][2] = de[i][3] = 0.f;
        dm[i][0] = dm[i][1] = dm[i][2] = dm[i][3] = 0.f;
        {  // accumulator init: bias (+ the one-hot rows of the action / future)
            const float* bp = W + op.b_off + (nt0 + i) * 8 + 2 * t;
            const float b0 = ldw1<WSMEM>(bp), b1 = ldw1<WSMEM>(bp + 1);
            dm[i][0] = b0; dm[i][1] = b1; dm[i][2] = b0; dm[i][3] = b1;
            if (op.extra) {
                const float* e0 = wb + (K + ra0) * wld + i * 8 + 2 * t;
                const float* e1 = wb + (K + ra1) * wld + i * 8 + 2 * t;
                dm[i][0] += ldw1<WSMEM>(e0); dm[i][1] += ldw1<WSMEM>(e0 + 1);
                dm[i][2] += ldw1<WSMEM>(e1); dm[i][3] += ldw1<WSMEM>(e1 + 1);
                if (op.extra == 2) {
                    const float* f = wb + (K + A + z) * wld + i * 8 + 2 * t;
                    const float f0 = ldw1<WSMEM>(f), f1 = ldw1<WSMEM>(f + 1);
                    dm[i][0] += f0; dm[i][1] += f1; dm[i][2] += f0; dm[i][3] += f1;
                }
            }
        }
    }
    if (dbg && threadIdx.x == 0) q1 = clock64();
    const float* wk = wb + t * wld + g;  // B fragment: b0 = Wt[k0 + t][n + g], b1 = Wt[k0 + t + 4][n + g]
    const int kfull = K & ~7;
#pragma unroll 2
    for (int k0 = 0; k0 < kfull; k0 += 8) {
        const float a0f = x0[k0 + t], a1f = x1[k0 + t], a2f = x0[k0 + t + 4], a3f = x1[k0 + t + 4];
        float bf[NT][2];
#pragma unroll
        for (int i = 0; i < NT; ++i) {
            bf[i][0] = ldw1<WSMEM>(wk + i * 8);
            bf[i][1] = ldw1<WSMEM>(wk + i * 8 + 4 * wld);
        }
        wk += 8 * wld;
        unsigned ah[4], al[4];
        ah[0] = tf32_hi(a0f); ah[1] = tf32_hi(a1f); ah[2] = tf32_hi(a2f); ah[3] = tf32_hi(a3f);
        al[0] = __float_as_uint(a0f - __uint_as_float(ah[0]));
        al[1] = __float_as_uint(a1f - __uint_as_float(ah[1]));
        al[2] = __float_as_uint(a2f - __uint_as_float(ah[2]));
        al[3] = __float_as_uint(a3f - __uint_as_float(ah[3]));
#pragma unroll
        for (int i = 0; i < NT; ++i) {
            {
                unsigned bh[2], bl[2];
                bh[0] = tf32_hi(bf[i][0]); bh[1] = tf32_hi(bf[i][1]);
                bl[0] = __float_as_uint(bf[i][0] - __uint_as_float(bh[0]));
                bl[1] = __float_as_uint(bf[i][1] - __uint_as_float(bh[1]));
                mma_tf32(dm[i], ah, bh);
                mma_tf32(dc[i], al, bh);
                mma_tf32(de[i], ah, bl);
            }
        }
    }
    if (kfull < K) {  // K not a multiple of 8 (e.g. the 465-wide observation): zero the out-of-range operands
        const int k0 = kfull;
        const bool lo_ok = (k0 + t) < K, hi_ok = (k0 + t + 4) < K;
        const float a0f = lo_ok ? x0[k0 + t] : 0.f, a1f = lo_ok ? x1[k0 + t] : 0.f;
        const float a2f = hi_ok ? x0[k0 + t + 4] : 0.f, a3f = hi_ok ? x1[k0 + t + 4] : 0.f;
        unsigned ah[4], al[4];
        ah[0] = tf32_hi(a0f); ah[1] = tf32_hi(a1f); ah[2] = tf32_hi(a2f); ah[3] = tf32_hi(a3f);
        al[0] = __float_as_uint(a0f - __uint_as_float(ah[0]));
        al[1] = __float_as_uint(a1f - __uint_as_float(ah[1]));
        al[2] = __float_as_uint(a2f - __uint_as_float(ah[2]));
        al[3] = __float_as_uint(a3f - __uint_as_float(ah[3]));
#pragma unroll
        for (int i = 0; i < NT; ++i) {
            {
                const float b0f = lo_ok ? ldw1<WSMEM>(wk + i * 8) : 0.f;
                const float b1f = hi_ok ? ldw1<WSMEM>(wk + i * 8 + 4 * wld) : 0.f;
                unsigned bh[2], bl[2];
                bh[0] = tf32_hi(b0f); bh[1] = tf32_hi(b1f);
                bl[0] = __float_as_uint(b0f - __uint_as_float(bh[0]));
                bl[1] = __float_as_uint(b1f - __uint_as_float(bh[1]));
                mma_tf32(dm[i], ah, bh);
                mma_tf32(dc[i], al, bh);
                mma_tf32(de[i], ah, bl);
            }
        }
    }
    if (dbg && threadIdx.x == 0) q2 = clock64();
    // C fragment: d0,d1 = C[g][2t, 2t+1], d2,d3 = C[g+8][2t, 2t+1]
#pragma unroll
    for (int i = 0; i < NT; ++i) {
        {
            float d0 = dm[i][0] + (dc[i][0] + de[i][0]), d1 = dm[i][1] + (dc[i][1] + de[i][1]);
            float d2 = dm[i][2] + (dc[i][2] + de[i][2]), d3 = dm[i][3] + (dc[i][3] + de[i][3]);
            if (op.act) {
                d0 = elu(d0); d1 = elu(d1); d2 = elu(d2); d3 = elu(d3);
            }
            float* y = act + op.out_off + (nt0 + i) * 8 + 2 * t;
            if (on0) *reinterpret_cast<float2*>(y + r0 * rs) = make_float2(d0, d1);
            if (on1) *reinterpret_cast<float2*>(y + r1 * rs) = make_float2(d2, d3);
        }
    }
    if (dbg && threadIdx.x == 0) {
        const long long q3 = clock64();
        atomicAdd(dbg + 40, (unsigned long long)(q1 - q0));  // accumulator init (bias, one-hot rows)
        atomicAdd(dbg + 41, (unsigned long long)(q2 - q1));  // k loop
        atomicAdd(dbg + 42, (unsigned long long)(q3 - q2));  // ELU + store
        atomicAdd(dbg + 43, 1ull);
    }
}

template <bool WSMEM>
__device__ __forceinline__ void mma_tiles_n(int n, const LayerOp& op, const float* __restrict__ W, float* __restrict__ act,
                                            int rs, int mt, int nt0, const int* __restrict__ row_action,
                                            const unsigned char* __restrict__ row_active, int A, int z,
                                            unsigned long long* dbg = nullptr) {
    switch (n) {
        case 1: mma_tiles<WSMEM, 1>(op, W, act, rs, mt, nt0, row_action, row_active, A, z, dbg); break;
        case 2: mma_tiles<WSMEM, 2>(op, W, act, rs, mt, nt0, row_action, row_active, A, z, dbg); break;
        case 3: mma_tiles<WSMEM, 3>(op, W, act, rs, mt, nt0, row_action, row_active, A, z, dbg); break;
        default: mma_tiles<WSMEM, 4>(op, W, act, rs, mt, nt0, row_action, row_active, A, z, dbg); break;
    }
}

// Runtime-indexed reads of kernel parameters compile to indexed constant loads (LDC) with a latency of
// hundreds of cycles; the schedules are therefore copied to shared memory once per CTA.
__device__ __forceinline__ void copy_schedule(Schedule* dst, const Schedule& src, int tid, int nthreads) {
    const int* s = reinterpret_cast<const int*>(&src);
    int* d = reinterpret_cast<int*>(dst);
    for (int i = tid; i < (int)(sizeof(Schedule) / sizeof(int)); i += nthreads) d[i] = s[i];
}

// Evaluate all stages for the TM rows of this CTA. Every thread of the CTA must call it.
template <int TM, int THREADS, bool WSMEM, bool MMA = true>
__device__ void run_schedule(const Schedule& sc, const float* __restrict__ W, float* __restrict__ act,
                             const int* __restrict__ row_action, const unsigned char* __restrict__ row_active,
                             int A, int z, unsigned long long* dbg = nullptr) {
    const int tid = threadIdx.x;
    const int rs = sc.row_stride;
    for (int s = 0; s < sc.n_stages; ++s) {
        const Stage& st = sc.st[s];
        long long d0 = 0, d1 = 0;
        if (dbg && tid == 0) d0 = clock64();
        if (MMA) {
            // tensor-core path: one warp per unit = 16 rows x ng adjacent 8-column tiles
            // (unit counts are precomputed on the host: no integer divisions by runtime values here)
            const int nops = st.n_ops;
            int total = 0;
            for (int o = 0; o < nops; ++o) total += st.ops[o].units;
            for (int u = tid >> 5; u < total; u += THREADS / 32) {
                int o = 0, uu = u;
                while (uu >= st.ops[o].units) {
                    uu -= st.ops[o].units;
                    ++o;
                }
                const LayerOp& op = st.ops[o];
                const int groups = op.groups;
                const int mt = (TM == 16 || uu < groups) ? 0 : 1;  // TM / 16 <= 2 row tiles
                const int nt0 = (uu - mt * groups) * op.ng;
                const int rem = (op.Npad >> 3) - nt0;
                mma_tiles_n<WSMEM>(rem < op.ng ? rem : op.ng, op, W, act, rs, mt, nt0, row_action, row_active, A, z, dbg);
            }
        } else if (!WSMEM) {
            // weights in global memory: 4 x 4 register tiles
            int total = 0;
            for (int o = 0; o < st.n_ops; ++o) total += (TM / 4) * (st.ops[o].Npad >> 2);
            for (int t = tid; t < total; t += THREADS) {
                int o = 0, tt = t;
                while (tt >= (TM / 4) * (st.ops[o].Npad >> 2)) {
                    tt -= (TM / 4) * (st.ops[o].Npad >> 2);
                    ++o;
                }
                dense_tile_g<4>(st.ops[o], W, act, rs, tt, row_action, row_active, A, z);
            }
        } else {
            // units of this stage: (op, chunk of ct columns); thread -> (unit = tid / TM, row = tid % TM)
            int total = 0;
            for (int o = 0; o < st.n_ops; ++o) total += st.ops[o].Npad / st.ops[o].ct;
            const int urow = tid % TM;
            const bool uactive = row_active[urow] != 0;
            const int uaction = row_action[urow];
            for (int u = tid / TM; u < total; u += THREADS / TM) {
                int o = 0, uu = u;
                while (uu >= st.ops[o].Npad / st.ops[o].ct) {
                    uu -= st.ops[o].Npad / st.ops[o].ct;
                    ++o;
                }
                const LayerOp& op = st.ops[o];
                switch (op.ct) {
                    case 4: dense_unit<WSMEM, 4>(op, W, act, rs, urow, uu * 4, uaction, uactive, A, z); break;
                    case 8: dense_unit<WSMEM, 8>(op, W, act, rs, urow, uu * 8, uaction, uactive, A, z); break;
                    default: dense_unit<WSMEM, 16>(op, W, act, rs, urow, uu * 16, uaction, uactive, A, z); break;
                }
            }
        }
        if (st.row_op == 1) {
            // models.py:223-231 / 248-255: (x - min) / scale, scale += 1e-5 only where scale < 1e-5
            const int row = tid >> 3, j = tid & 7;
            const unsigned gmask = 0xFFu << ((tid & 31) & 24);
            if (row < TM && row_active[row]) {
                const float* x = act + row * rs + st.r_in;
                float mn = CUDART_INF_F, mx = -CUDART_INF_F;
                for (int i = j; i < st.r_n; i += GW) {
                    const float v = x[i];
                    mn = fminf(mn, v);
                    mx = fmaxf(mx, v);
                }
#pragma unroll
                for (int d = 1; d < GW; d <<= 1) {
                    mn = fminf(mn, __shfl_xor_sync(gmask, mn, d));
                    mx = fmaxf(mx, __shfl_xor_sync(gmask, mx, d));
                }
                float scale = __fsub_rn(mx, mn);
                if (scale < 1e-5f) scale = __fadd_rn(scale, 1e-5f);
                // one IEEE reciprocal per row instead of one division per element (<= 1 ulp from torch's
                // (x - min) / scale, far inside the 1e-5 network tolerance)
                const float inv = __frcp_rn(scale);
                float* y = act + row * rs + st.r_out;
                for (int i = j; i < st.r_n; i += GW) y[i] = __fmul_rn(__fsub_rn(x[i], mn), inv);
            }
        }
        if (dbg && tid == 0) d1 = clock64();
        __syncthreads();
        if (dbg && tid == 0) {
            const long long d2 = clock64();
            atomicAdd(dbg + 8 + 2 * s, (unsigned long long)(d1 - d0));      // thread 0's own work in stage s
            atomicAdd(dbg + 9 + 2 * s, (unsigned long long)(d2 - d1));      // its wait at the stage barrier
        }
    }
}

// Warp-resident MLP: ONE warp evaluates every stage of the schedule for the 16 rows (trees) of row tile
// `mt`. The A fragments of a layer are that same warp's C fragments of the previous layer (round-tripped
// through its rows of the activation strip), so stages are separated by __syncwarp() only -- no CTA
// barrier, no tile decode across warps -- and with 4 column tiles in flight (12 independent mma chains)
// the warp issues HMMAs back to back. Measured against the CTA-wide stage-by-stage version in DESIGN.md.
template <bool WSMEM>
__device__ void run_schedule_warp(const Schedule& sc, const float* __restrict__ W, float* __restrict__ act,
                                  const int* __restrict__ row_action, const unsigned char* __restrict__ row_active,
                                  int A, int z, int mt, unsigned long long* dbg = nullptr) {
    const int lane = threadIdx.x & 31;
    const int rs = sc.row_stride;
    for (int s = 0; s < sc.n_stages; ++s) {
        const Stage& st = sc.st[s];
        long long d0 = 0;
        if (dbg && threadIdx.x == 0) d0 = clock64();
        const int nops = st.n_ops;
        for (int o = 0; o < nops; ++o) {
            const LayerOp& op = st.ops[o];
            const int ntiles = op.Npad >> 3;
            for (int nt0 = 0; nt0 < ntiles; nt0 += 4)
                mma_tiles_n<WSMEM>(ntiles - nt0, op, W, act, rs, mt, nt0, row_action, row_active, A, z, nullptr);
        }
        if (st.row_op == 1) {  // min-max scaling of the 16 rows of this tile, 8 lanes per row, 4 rows per pass
            const int j = lane & 7;
            const unsigned gmask = 0xFFu << (lane & 24);
#pragma unroll 1
            for (int r = 0; r < 16; r += 4) {
                const int row = mt * 16 + r + (lane >> 3);
                if (row_active[row]) {
                    const float* x = act + row * rs + st.r_in;
                    float mn = CUDART_INF_F, mx = -CUDART_INF_F;
                    for (int i = j; i < st.r_n; i += GW) {
                        const float v = x[i];
                        mn = fminf(mn, v);
                        mx = fmaxf(mx, v);
                    }
#pragma unroll
                    for (int d = 1; d < GW; d <<= 1) {
                        mn = fminf(mn, __shfl_xor_sync(gmask, mn, d));
                        mx = fmaxf(mx, __shfl_xor_sync(gmask, mx, d));
                    }
                    float scale = __fsub_rn(mx, mn);
                    if (scale < 1e-5f) scale = __fadd_rn(scale, 1e-5f);
                    const float inv = __frcp_rn(scale);
                    float* y = act + row * rs + st.r_out;
                    for (int i = j; i < st.r_n; i += GW) y[i] = __fmul_rn(__fsub_rn(x[i], mn), inv);
                }
            }
        }
        __syncwarp();
        if (dbg && threadIdx.x == 0) atomicAdd(dbg + 8 + 2 * s, (unsigned long long)(clock64() - d0));
    }
}

// ------------------------------------------------------------ row-group helpers ----
__device__ __forceinline__ float group_max(float v, unsigned gmask) {
#pragma unroll
    for (int d = 1; d < GW; d <<= 1) v = fmaxf(v, __shfl_xor_sync(gmask, v, d));
    return v;
}
__device__ __forceinline__ float group_sum(float v, unsigned gmask) {
#pragma unroll
    for (int d = 1; d < GW; d <<= 1) v = __fadd_rn(v, __shfl_xor_sync(gmask, v, d));
    return v;
}
__device__ __forceinline__ double shfl_d(unsigned gmask, double v, int src_lane_in_warp) {
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_sync(gmask, lo, src_lane_in_warp);
    hi = __shfl_sync(gmask, hi, src_lane_in_warp);
    return __hiloint2double(hi, lo);
}
__device__ __forceinline__ double shfl_xor_d(unsigned gmask, double v, int d) {
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_xor_sync(gmask, lo, d);
    hi = __shfl_xor_sync(gmask, hi, d);
    return __hiloint2double(hi, lo);
}

// models.py:1192-1200 in torch's float32 operation order (no FMA contraction)
__device__ __forceinline__ float inverse_value_transform(float x) {
    const float c4e = (float)(4 * 0.001);
    const float c2e = (float)(2 * 0.001);
    float t = __fadd_rn(fabsf(x), 1.0f);
    t = __fadd_rn(t, 0.001f);
    t = __fmul_rn(c4e, t);
    t = __fadd_rn(1.0f, t);
    t = __fsqrt_rn(t);
    t = __fsub_rn(t, 1.0f);
    t = __fdiv_rn(t, c2e);
    t = __fmul_rn(t, t);
    t = __fsub_rn(t, 1.0f);
    const float sgn = (x > 0.0f) ? 1.0f : ((x < 0.0f) ? -1.0f : 0.0f);
    return __fmul_rn(sgn, t);
}

// models.py:1180-1201 by one 8-lane group; every lane returns the scalar
__device__ __forceinline__ float support_to_scalar_group(const float* logits, int support, int j, unsigned gmask) {
    const int n = 2 * support + 1;
    float m = -CUDART_INF_F;
    for (int i = j; i < n; i += GW) m = fmaxf(m, logits[i]);
    m = group_max(m, gmask);
    float s = 0.0f, xs = 0.0f;
    for (int i = j; i < n; i += GW) {
        const float e = expf(__fsub_rn(logits[i], m));
        s = __fadd_rn(s, e);
        xs = fmaf((float)(i - support), e, xs);
    }
    s = group_sum(s, gmask);
    xs = group_sum(xs, gmask);
    // sum_i support_i * (e_i / s): one division instead of n (differs from torch by rounding only)
    return inverse_value_transform(__fdiv_rn(xs, s));
}

// softmax over up to 8 logits held one per lane (lanes >= n pass -inf); returns this lane's probability
__device__ __forceinline__ float softmax_group(float logit, bool active, unsigned gmask) {
    const float m = group_max(active ? logit : -CUDART_INF_F, gmask);
    const float e = active ? expf(__fsub_rn(logit, m)) : 0.0f;
    const float s = group_sum(e, gmask);
    return __fdiv_rn(e, s);
}

// ------------------------------------------------------------------- tree math ----
struct MinMax {
    double mn, mx;
    __device__ __forceinline__ void update(double v) {  // self_play.py:669-671
        mx = mx > v ? mx : v;
        mn = mn < v ? mn : v;
    }
    __device__ __forceinline__ double normalize(double v) const {  // self_play.py:673-677
        if (mx > mn) return __ddiv_rn(__dsub_rn(v, mn), __dsub_rn(mx, mn));
        return v;
    }
};

// self_play.py:453-477 (VARIANT 0), self_play_stochastic.py:488-508 (1), self_play_risk_sensitive.py:496-520 (2)
template <int VARIANT>
__device__ __forceinline__ double ucb_score(double pb_log, double sqrt_parent_visits, int visits, double vsum,
                                            double prior, float reward, double discount, const MinMax& mm,
                                            bool two_player = false) {
    double pb_c = __dmul_rn(pb_log, __ddiv_rn(sqrt_parent_visits, (double)(visits + 1)));
    const double prior_score = __dmul_rn(pb_c, prior);
    double value_score = 0.0;
    if (visits > 0) {
        if (VARIANT == 0) {
            const double q = __ddiv_rn(vsum, (double)visits);
            // two players: -child.value() (self_play.py:469-472)
            value_score = mm.normalize(__dadd_rn((double)reward, __dmul_rn(discount, two_player ? -q : q)));
        } else if (VARIANT == 1) {
            value_score = mm.normalize(__ddiv_rn(vsum, (double)visits));
        } else {
            value_score = mm.normalize(__dadd_rn((double)reward, __dmul_rn(discount, vsum)));
        }
    }
    return __dadd_rn(prior_score, value_score);
}

// --------------------------------------------------------------------- params ----
struct NetParams {
    int B, A, H, nf, support;  // B: one past the last row of this launch
    int b0;                    // first row of this launch (tsp_run overlaps the root network with the upload of later rows)
    int mode;  // 0 initial_inference (+ optional tree init), 1 recurrent, 2 stochastic dynamics, 3 prediction
    const float* W;
    const float* in;      // observations [B][obs_dim] or hidden [B][H]
    const int* action;    // [B] (modes 1, 2)
    // outputs (device, may be null)
    float* value; float* reward; float* terminal; float* policy_logits; float* hidden_out;
    float* value_logits; float* reward_logits; float* transition_logits;
    // tree init (mode 0 when tree_init != 0)
    int tree_init, add_noise, uniform_policy, blocks_per_tree, hidden_per_tree;
    double one_minus_frac, frac;
    void* blocks; float* hidden_pool; TreeHeader* hdr; double* rootprior; int* rootact;
    const int* legal; const int* num_legal; const double* noise;
    const float* fed_root; float* trace_root; float* root_pred; float* root_hidden;
    Schedule sched;
};

struct SearchParams {
    int B, S, A, H, nf, support, mask_absorbing, uniform_policy;
    int blocks_per_tree, hidden_per_tree, max_rec, T, K, max_iters, w_smem_floats, pb_stride, warp_mlp, two_player;
    int S_tab;                 // simulations the trees hold after this run (size of the visit-count tables)
    int resume, add_noise;     // resume: continue the trees of the previous run (override_root_with, self_play.py:331-333)
    double one_minus_frac, frac;
    const double* noise;      // [B][A], mixed into the root priors again on resume (self_play.py:372-376)
    double discount;
    const double* pb_table;  // [2][S + 2]: log((N + base + 1) / base) + init, then sqrt(N); host libm (bit-equal to math.log / math.sqrt)
    void* blocks; float* hidden_pool; TreeHeader* hdr; double* rootprior; const int* rootact;
    const unsigned char* tie; const unsigned char* chance;
    const float* fed; float* trace;
    const float* W;
    int* visit_counts; double* root_value; double* child_value; double* child_prior; double* child_reward;
    int* max_depth; int* num_records; int* total_depth;
    unsigned long long* phase_cycles;  // optional [8]: summed clock64 deltas of thread 0 of every CTA
    Schedule sched_a;  // VARIANT 0: recurrent inference; else: transition expansion (dynamics, per future z)
    Schedule sched_b;  // VARIANT != 0: prediction
};

// ------------------------------------------------------------- network kernel ----
template <int NC, int TM>
__global__ void __launch_bounds__(TM * GW) net_kernel(const __grid_constant__ NetParams p) {
    extern __shared__ __align__(16) float smem[];
    constexpr int THREADS = TM * GW;
    __shared__ Schedule s_sched;
    copy_schedule(&s_sched, p.sched, threadIdx.x, THREADS);
    __syncthreads();
    const Schedule& sc = s_sched;
    const int rs = sc.row_stride;
    float* act = smem;
    int* row_action = reinterpret_cast<int*>(act + TM * rs);
    unsigned char* row_active = reinterpret_cast<unsigned char*>(row_action + TM);
    const int tid = threadIdx.x, j = tid & 7, row = tid >> 3;
    const unsigned gmask = 0xFFu << ((tid & 31) & 24);
    const int b = p.b0 + blockIdx.x * TM + row;
    const bool valid = b < p.B;
    const bool fed = p.fed_root != nullptr;

    if (j == 0) {
        row_active[row] = valid ? 1 : 0;
        row_action[row] = (valid && p.action) ? p.action[b] : 0;
    }
    if (valid && !fed) {
        const float* src = p.in + (size_t)b * sc.in_dim;
        float* dst = act + row * rs + sc.off_in;
        for (int i = j; i < sc.in_dim; i += GW) dst[i] = src[i];
    }
    __syncthreads();
    const int nz = (p.mode == 2) ? p.nf : 1;
    for (int z = 0; z < nz; ++z) {
        if (!fed) run_schedule<TM, THREADS, false>(sc, p.W + (size_t)z * sc.f_stride, act, row_action, row_active, p.A, z);
        if (valid && !fed) {
            const float* r = act + row * rs;
            const int V = 2 * p.support + 1;
            if (p.mode == 2) {
                if (p.hidden_out)
                    for (int i = j; i < p.H; i += GW) p.hidden_out[((size_t)b * p.nf + z) * p.H + i] = r[sc.off_hidden + i];
                const float rw = support_to_scalar_group(r + sc.off_rl, p.support, j, gmask);
                if (p.reward && j == 0) p.reward[(size_t)b * p.nf + z] = rw;
                if (z == 0 && p.transition_logits && j < p.nf) p.transition_logits[(size_t)b * p.nf + j] = r[sc.off_tp + j];
            } else {
                if (p.hidden_out && sc.off_hidden >= 0)
                    for (int i = j; i < p.H; i += GW) p.hidden_out[(size_t)b * p.H + i] = r[sc.off_hidden + i];
                if (p.policy_logits && j < p.A) p.policy_logits[(size_t)b * p.A + j] = r[sc.off_pl + j];
                if (p.value_logits)
                    for (int i = j; i < V; i += GW) p.value_logits[(size_t)b * V + i] = r[sc.off_vl + i];
                if (p.value) {
                    const float v = support_to_scalar_group(r + sc.off_vl, p.support, j, gmask);
                    if (j == 0) p.value[b] = v;
                }
                if (p.mode == 1) {
                    if (p.reward_logits)
                        for (int i = j; i < V; i += GW) p.reward_logits[(size_t)b * V + i] = r[sc.off_rl + i];
                    if (p.reward) {
                        const float v = support_to_scalar_group(r + sc.off_rl, p.support, j, gmask);
                        if (j == 0) p.reward[b] = v;
                    }
                    if (p.terminal && j == 0) p.terminal[b] = sc.off_tl >= 0 ? r[sc.off_tl] : -1.0f;
                }
            }
        }
        if (nz > 1) __syncthreads();
    }

    if (p.tree_init && valid) {
        // ---- root of the search: self_play.py:335-376 ----
        const float* r = act + row * rs;
        const int nl = p.num_legal ? p.num_legal[b] : p.A;
        const int a_j = (j < nl) ? (p.legal ? p.legal[(size_t)b * p.A + j] : j) : 0;
        float prior, value;
        if (fed) {
            value = p.fed_root[(size_t)b * REC + 1];
            prior = (j < nl) ? p.fed_root[(size_t)b * REC + 4 + j] : 0.0f;
        } else {
            value = support_to_scalar_group(r + sc.off_vl, p.support, j, gmask);
            const float lg = (j < nl && !p.uniform_policy) ? r[sc.off_pl + a_j] : 0.0f;
            prior = softmax_group(lg, j < nl, gmask);  // Node.expand, self_play.py:532-538
        }
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
        NodeBlock<NC>* blk = reinterpret_cast<NodeBlock<NC>*>(p.blocks) + (size_t)b * p.blocks_per_tree;
        if (j < NC) {
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
        if (!fed) {
            float* hp = p.hidden_pool + (size_t)b * p.hidden_per_tree * p.H;
            for (int i = j; i < p.H; i += GW) {
                const float v = r[sc.off_hidden + i];
                hp[i] = v;
                if (p.root_hidden) p.root_hidden[(size_t)b * p.H + i] = v;
            }
        }
    }
}

// -------------------------------------------------------------- search kernel ----
// VARIANT 0: MCTS.run of self_play.py:378-427; 1: self_play_stochastic.py:329-393; 2: risk-sensitive.
// XW: thread multiplier. The first TM * 8 threads own the trees (8 lanes each); the other
// (XW - 1) * TM * 8 threads only help with the MLP tiles of phase 2.
template <int VARIANT, int NC, int TM, bool WSMEM, int XW>
__global__ void __launch_bounds__(TM * GW * XW) search_kernel(const __grid_constant__ SearchParams p) {
    extern __shared__ __align__(16) float smem[];
    constexpr int THREADS = TM * GW * XW;
    using Block = NodeBlock<NC>;
    __shared__ Schedule s_sched[VARIANT == 0 ? 1 : 2];
    copy_schedule(&s_sched[0], p.sched_a, threadIdx.x, THREADS);
    if (VARIANT != 0) copy_schedule(&s_sched[VARIANT == 0 ? 0 : 1], p.sched_b, threadIdx.x, THREADS);
    const Schedule& sched_a = s_sched[0];
    const Schedule& sched_b = s_sched[VARIANT == 0 ? 0 : 1];
    const int rs = p.sched_a.row_stride;  // the host gives sched_a and sched_b the same row stride
    // shared-memory carve-out (offsets only -- no integer round trips, so every pointer stays "shared")
    double* pb = reinterpret_cast<double*>(smem);  // pb[N] = log((N + base + 1) / base) + init
    const int pb_n = (p.S_tab + 2 + 1) & ~1;  // S_tab: largest parent visit count (sums over continued searches)
    double* sq = pb + pb_n;                        // sq[N] = sqrt(N), correctly rounded (host libm)
    float* wsm = smem + 4 * pb_n;                  // optional copy of the search MLP weights
    float* act = wsm + (WSMEM ? p.w_smem_floats : 0);
    int* row_action = reinterpret_cast<int*>(act + TM * rs);
    unsigned char* row_active = reinterpret_cast<unsigned char*>(row_action + TM);
    unsigned char* row_kind = row_active + TM;
    const float* Wp;
    if (WSMEM) Wp = wsm; else Wp = p.W;

    const int tid = threadIdx.x, lane = tid & 31, j = tid & 7, row = tid >> 3;
    const unsigned gmask = 0xFFu << (lane & 24);
    const int gl0 = lane & 24;  // first lane of this group inside the warp
    const int b = blockIdx.x * TM + row;
    const bool valid = (row < TM) && (b < p.B);  // helper threads (row >= TM) own no tree
    const bool fed = p.fed != nullptr;
    const int A = p.A, H = p.H, nf = p.nf;
    const double discount = p.discount;

    for (int i = tid; i < p.S_tab + 2; i += THREADS) {
        pb[i] = p.pb_table[i];
        sq[i] = p.pb_table[p.pb_stride + i];
    }
    if (WSMEM) {
        const float4* src = reinterpret_cast<const float4*>(p.W);
        float4* dst = reinterpret_cast<float4*>(wsm);
        for (int i = tid; i < (p.w_smem_floats >> 2); i += THREADS) dst[i] = __ldg(src + i);
    }

    Block* blocks = reinterpret_cast<Block*>(p.blocks) + (size_t)(valid ? b : 0) * p.blocks_per_tree;
    float* hpool = p.hidden_pool + (size_t)(valid ? b : 0) * p.hidden_per_tree * H;
    const float* fed_rec = fed ? p.fed + (size_t)(valid ? b : 0) * p.max_rec * REC : nullptr;
    float* trace_rec = p.trace ? p.trace + (size_t)(valid ? b : 0) * p.max_rec * REC : nullptr;

    // per-tree state, replicated in the 8 lanes of the group
    MinMax mm;
    mm.mn = CUDART_INF;
    mm.mx = -CUDART_INF;
    double root_vsum = 0.0, root_prior = 0.0;
    int root_visit = 0, n_blocks = 1, n_tgroups = 0, max_depth = 0, n_sims = 0, k_tie = 0, k_chance = 0, n_rec = 0;
    int n_root = 0, root_act = 0, depth_total = 0;
    if (valid) {
        const TreeHeader h = p.hdr[b];
        n_root = h.n_root;
        root_prior = p.rootprior[(size_t)b * GW + j];
        root_act = p.rootact[(size_t)b * GW + j];
        if (p.resume) {  // MCTS.run(..., override_root_with=root): keep the tree, fresh MinMaxStats / depth / streams
            root_vsum = h.root_vsum;
            root_visit = h.root_visit;
            n_blocks = h.n_blocks;
            n_tgroups = h.n_tgroups;
            if (p.add_noise && j < n_root) {  // add_exploration_noise once more, self_play.py:372-376, 540-549
                root_prior = __dadd_rn(__dmul_rn(root_prior, p.one_minus_frac), __dmul_rn(p.noise[(size_t)b * A + j], p.frac));
                p.rootprior[(size_t)b * GW + j] = root_prior;
            }
        }
    }
    __syncthreads();

    for (int it = 0; it < p.max_iters; ++it) {
        if (fed && n_rec >= p.max_rec) n_sims = p.S;  // fed records exhausted: stop this tree (results will not match)
        const bool active = valid && n_sims < p.S;
        if (VARIANT != 0) {
            if (!__syncthreads_or(active ? 1 : 0)) break;
        } else {
            if (it >= p.S) break;
        }

        long long tc0 = 0, tc1 = 0, tc2 = 0, tc3 = 0;
        if (p.phase_cycles && tid == 0) tc0 = clock64();
        // ================= phase 1: selection (one 8-lane group per tree) =================
        int node = 0, slot = 0, depth = 0, Np = root_visit;
        bool parent_is_state = true;  // kind of block `node`
        if (active) {
            while (true) {
                const Block* blk = blocks + node;
                int vis_sel, child_sel;
                if (VARIANT == 0 || parent_is_state) {
                    const int nch = (node == 0) ? n_root : A;
                    const bool on = j < nch;
                    int vis = 0, ch = -1;
                    double vs = 0.0, prior = 0.0;
                    float rw = 0.0f;
                    if (on) {
                        vis = blk->visit[j];
                        ch = blk->child[j];
                        vs = blk->vsum[j];
                        rw = blk->reward[j];
                        prior = (node == 0) ? root_prior : (double)blk->prior[j];
                        // pull every expanded child's block towards L1 while this level's float64 scores
                        // are computed: the next level's five loads then hit instead of paying L2 latency
                        if (ch >= 0) asm volatile("prefetch.global.L1 [%0];" ::"l"(blocks + ch));
                    }
                    double score = -CUDART_INF;
                    if (on) score = ucb_score<VARIANT>(pb[Np], sq[Np], vis, vs, prior, rw, discount, mm, VARIANT == 0 && p.two_player);
                    double best = score;
#pragma unroll
                    for (int d = 1; d < GW; d <<= 1) {
                        const double o = shfl_xor_d(gmask, best, d);
                        best = o > best ? o : best;
                    }
                    const unsigned tied = (__ballot_sync(gmask, on && score == best) >> gl0) & 0xFFu;
                    const int nt = __popc(tied);
                    if (nt == 1) {
                        slot = __ffs(tied) - 1;
                    } else {  // self_play.py:444-450 with the external tie stream
                        int v = 0;
                        if (p.tie) v = p.tie[(size_t)b * p.T + (k_tie % p.T)];
                        k_tie++;
                        slot = __fns(tied, 0, (v % nt) + 1);
                    }
                    vis_sel = __shfl_sync(gmask, vis, gl0 + slot);
                    child_sel = __shfl_sync(gmask, ch, gl0 + slot);
                } else {  // TransitionNode.select_child, self_play_stochastic.py:517-533
                    slot = 0;
                    if (nf > 1) {
                        const int v = p.chance ? p.chance[(size_t)b * p.K + (k_chance % p.K)] : 0;
                        k_chance++;
                        slot = v % nf;
                    }
                    vis_sel = blk->visit[slot];
                    child_sel = blk->child[slot];
                }
                depth++;
                if (child_sel < 0) break;
                node = child_sel;
                Np = vis_sel;
                if (VARIANT != 0) parent_is_state = !parent_is_state;
            }
        }
        // leaf = child `slot` of block `node`
        const bool leaf_is_state = (VARIANT == 0) ? true : !parent_is_state;
        int action = 0, hslot = 0;
        if (active) {
            if (VARIANT == 0) {
                action = (node == 0) ? __shfl_sync(gmask, root_act, gl0 + slot) : slot;
                hslot = node;
            } else if (!leaf_is_state) {
                action = (node == 0) ? __shfl_sync(gmask, root_act, gl0 + slot) : slot;
                hslot = (node == 0) ? 0 : (blocks[node].aux >> 8);
            } else {
                hslot = 1 + (blocks[node].aux >> 8) * nf + slot;
            }
            if (depth > max_depth) max_depth = depth;
            depth_total += depth;
        }
        const int kind = !active ? 0 : (leaf_is_state ? 1 : 2);

        // ================= phase 2: network evaluation of the leaf batch =================
        float value = 0.0f, reward = 0.0f, terminal = -1.0f, prior = 0.0f;  // state-leaf results
        float tprior = 0.0f, treward = 0.0f;                               // transition-leaf results (lane z)
        if (!fed) {
            if (j == 0 && row < TM) {
                row_action[row] = action;
                row_kind[row] = (unsigned char)kind;
                if (VARIANT == 0) row_active[row] = active ? 1 : 0;
            }
            if (active) {  // gather the input hidden state of this row
                const float4* src = reinterpret_cast<const float4*>(hpool + (size_t)hslot * H);
                float4* dst = reinterpret_cast<float4*>(act + row * rs);  // off_in == 0 in every search schedule
                for (int i = j; i < (H >> 2); i += GW) dst[i] = src[i];
            }
            if (p.phase_cycles && tid == 0) tc1 = clock64();
            __syncthreads();
            if (p.phase_cycles && tid == 0) tc2 = clock64();
            if (VARIANT == 0) {
                if (p.warp_mlp) {  // developer alternative: one warp per 16-row tile walks all stages
                    if ((tid >> 5) < TM / 16)
                        run_schedule_warp<WSMEM>(sched_a, Wp, act, row_action, row_active, A, 0, tid >> 5, p.phase_cycles);
                    __syncthreads();
                } else {
                    run_schedule<TM, THREADS, WSMEM>(sched_a, Wp, act, row_action, row_active, A, 0, p.phase_cycles);
                }
                if (p.phase_cycles && tid == 0) tc3 = clock64();
                if (active) {
                    const float* r = act + row * rs;
                    value = support_to_scalar_group(r + sched_a.off_vl, p.support, j, gmask);
                    reward = support_to_scalar_group(r + sched_a.off_rl, p.support, j, gmask);
                    if (p.mask_absorbing) terminal = r[sched_a.off_tl];
                    const float lg = (j < A && !p.uniform_policy) ? r[sched_a.off_pl + j] : 0.0f;
                    prior = softmax_group(lg, j < A, gmask);
                }
            } else {
                const int any_t = __syncthreads_or(kind == 2);
                const int any_s = __syncthreads_or(kind == 1);
                if (any_t) {
                    if (j == 0 && row < TM) row_active[row] = (kind == 2) ? 1 : 0;
                    __syncthreads();
                    for (int z = 0; z < nf; ++z) {
                        run_schedule<TM, THREADS, WSMEM>(sched_a, Wp + z * sched_a.f_stride, act, row_action, row_active, A, z);
                        if (kind == 2) {
                            const float* r = act + row * rs;
                            const float rw = support_to_scalar_group(r + sched_a.off_rl, p.support, j, gmask);
                            if (j == z) treward = rw;
                            if (z == 0) {
                                const float lg = (j < nf) ? r[sched_a.off_tp + j] : 0.0f;
                                tprior = softmax_group(lg, j < nf, gmask);  // self_play_stochastic.py:553-554
                            }
                            // children's hidden states: slot 1 + t_ord * nf + z  (t_ord == n_tgroups)
                            float4* dst = reinterpret_cast<float4*>(hpool + (size_t)(1 + n_tgroups * nf + z) * H);
                            const float4* src = reinterpret_cast<const float4*>(r + sched_a.off_hidden);
                            for (int i = j; i < (H >> 2); i += GW) dst[i] = src[i];
                        }
                        __syncthreads();
                    }
                }
                if (any_s) {
                    if (j == 0 && row < TM) row_active[row] = (kind == 1) ? 1 : 0;
                    __syncthreads();
                    run_schedule<TM, THREADS, WSMEM>(sched_b, Wp, act, row_action, row_active, A, 0);
                    if (kind == 1) {
                        const float* r = act + row * rs;
                        value = support_to_scalar_group(r + sched_b.off_vl, p.support, j, gmask);
                        const float lg = (j < A) ? r[sched_b.off_pl + j] : 0.0f;
                        prior = softmax_group(lg, j < A, gmask);
                    }
                }
            }
        } else if (active) {
            const float* fr = fed_rec + (size_t)n_rec * REC;
            if (kind == 1) {
                value = fr[1];
                reward = fr[2];
                terminal = fr[3];
                prior = (j < A) ? fr[4 + j] : 0.0f;
            } else {
                tprior = (j < nf) ? fr[1 + j] : 0.0f;
                treward = (j < nf) ? fr[1 + nf + j] : 0.0f;
            }
        }

        // ================= phase 3: expansion + backup (one group per tree) =================
        if (active) {
            if (trace_rec && n_rec < p.max_rec) {
                float* tr = trace_rec + (size_t)n_rec * REC;
                for (int i = j; i < REC; i += GW) tr[i] = 0.0f;
                __syncwarp(gmask);
                if (kind == 1) {
                    if (j == 0) {
                        tr[0] = 1.0f;
                        tr[1] = value;
                        if (VARIANT == 0) {
                            tr[2] = reward;
                            tr[3] = terminal;
                        }
                    }
                    if (j < A) tr[4 + j] = prior;
                } else {
                    if (j == 0) tr[0] = 2.0f;
                    if (j < nf) {
                        tr[1 + j] = tprior;
                        tr[1 + nf + j] = treward;
                    }
                }
            }
            n_rec++;
            Block* pblk = blocks + node;
            if (kind == 2) {
                // ---- TransitionNode.expand, self_play_stochastic.py:535-557 (no backup, no sim count) ----
                const int nb = n_blocks++;
                Block* nblk = blocks + nb;
                if (j < NC) {
                    nblk->visit[j] = 0;
                    nblk->child[j] = -1;
                    nblk->prior[j] = tprior;
                    nblk->reward[j] = (j < nf) ? treward : 0.0f;  // decoded child_rewards[z]; read only once visited
                    nblk->vsum[j] = 0.0;
                }
                if (j == 0) {
                    nblk->parent = node;
                    nblk->aux = slot | (n_tgroups << 8);
                    pblk->child[slot] = nb;
                }
                n_tgroups++;
            } else {
                const bool masked = (VARIANT == 0) && p.mask_absorbing && (terminal >= 0.0f);  // self_play.py:408-417
                double v = masked ? 0.0 : (double)value;
                if (!masked) {
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
                        nblk->aux = slot | ((VARIANT == 0 ? nb : hslot) << 8);
                        pblk->child[slot] = nb;
                        if (VARIANT == 0) pblk->reward[slot] = reward;
                    }
                    if (VARIANT == 0 && !fed) {  // hidden state of the new node
                        float4* dst = reinterpret_cast<float4*>(hpool + (size_t)nb * H);
                        const float4* src = reinterpret_cast<const float4*>(act + row * rs + sched_a.off_hidden);
                        for (int i = j; i < (H >> 2); i += GW) dst[i] = src[i];
                    }
                }
                __syncwarp(gmask);
                // ---- backup: every lane of the group walks the path redundantly, lane 0 stores ----
                int cur = node, cs = slot;
                if (VARIANT == 0) {
                    // self_play.py:485-490
                    // The 8 lanes of the group execute this loop in lock step (identical control flow) and ALL
                    // of them store the identical updated statistics: no lane ever needs another lane's view
                    // of memory inside the loop, so no warp-level sync per level is required.
                    // two players (self_play.py:492-500): a node k plies above the leaf belongs to the leaf's player
                    // iff k is even -- except the unexpanded (masked terminal) leaf, whose to_play is still -1
                    const bool two = p.two_player != 0;
                    int k = 0;
                    while (true) {
                        Block* bk = blocks + cur;
                        const int ps = bk->aux, pc = bk->parent;  // issued with the statistics loads
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
                    {
                        const bool same = two && ((k & 1) == 0);
                        root_vsum = __dadd_rn(root_vsum, (two && !same) ? -v : v);
                        root_visit += 1;
                        const double q = __ddiv_rn(root_vsum, (double)root_visit);
                        mm.update(__dadd_rn(0.0, __dmul_rn(discount, two ? -q : q)));
                    }
                } else if (VARIANT == 1) {
                    // self_play_stochastic.py:408-418; entries alternate State (leaf), Transition, State, ...
                    bool is_state = true;
                    while (true) {
                        Block* bk = blocks + cur;
                        const double vs = __dadd_rn(bk->vsum[cs], v);
                        const int vis = bk->visit[cs] + 1;
                        const double rw = (double)bk->reward[cs];
                        __syncwarp(gmask);
                        if (j == 0) {
                            bk->vsum[cs] = vs;
                            bk->visit[cs] = vis;
                        }
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
                    v = __dadd_rn(0.0, __dmul_rn(discount, v));
                    mm.update(__dadd_rn(0.0, __dmul_rn(discount, __ddiv_rn(root_vsum, (double)root_visit))));
                } else {
                    // self_play_risk_sensitive.py:408-420, 452-461, 529-537
                    if (j == 0) {
                        pblk->vsum[slot] = v;  // search_path[-1].value = value
                        pblk->visit[slot] = pblk->visit[slot] + 1;
                    }
                    __syncwarp(gmask);
                    bool is_transition = true;  // kind of the node that OWNS block `cur`
                    while (true) {
                        Block* bk = blocks + cur;  // children statistics of the node being updated
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
                        if (cur == 0) {  // the root (a StateNode, reward 0)
                            root_vsum = nv;
                            root_visit += 1;
                            mm.update(__dadd_rn(0.0, __dmul_rn(discount, nv)));
                            break;
                        }
                        const int ps = bk->aux & 0xFF, pc = bk->parent;
                        Block* pb2 = blocks + pc;
                        const int vis = pb2->visit[ps] + 1;
                        const double rw = (double)pb2->reward[ps];
                        __syncwarp(gmask);
                        if (j == 0) {
                            pb2->vsum[ps] = nv;
                            pb2->visit[ps] = vis;
                        }
                        __syncwarp(gmask);
                        if (!is_transition) mm.update(__dadd_rn(rw, __dmul_rn(discount, nv)));
                        is_transition = !is_transition;
                        cur = pc;
                    }
                }
                n_sims++;
            }
        }
        if (p.phase_cycles && tid == 0 && VARIANT == 0 && !fed) {
            const long long tc4 = clock64();
            atomicAdd(p.phase_cycles + 0, (unsigned long long)(tc1 - tc0));  // select + gather (tree 0 of the CTA)
            atomicAdd(p.phase_cycles + 1, (unsigned long long)(tc2 - tc1));  // wait for the slowest tree of the CTA
            atomicAdd(p.phase_cycles + 2, (unsigned long long)(tc3 - tc2));  // MLP schedule
            atomicAdd(p.phase_cycles + 3, (unsigned long long)(tc4 - tc3));  // decode + expand + backup
            atomicAdd(p.phase_cycles + 4, 1ull);
        }
        // No CTA barrier here: a row's activation strip is re-filled by the very group that read it in
        // phase 3, and every tile of this iteration finished before run_schedule's last barrier.
        __syncwarp(gmask);
    }

    // ================= outputs: what consumers read from `root` =================
    if (valid) {
        const Block* rb = blocks;
        if (j == 0) {
            if (p.root_value)
                p.root_value[b] = (VARIANT == 2) ? root_vsum
                                                 : (root_visit == 0 ? 0.0 : __ddiv_rn(root_vsum, (double)root_visit));
            if (p.max_depth) p.max_depth[b] = max_depth;
            if (p.num_records) p.num_records[b] = n_rec;
            if (p.total_depth) p.total_depth[b] = depth_total;
            // what a continued search (resume) needs besides the pools
            p.hdr[b].root_vsum = root_vsum;
            p.hdr[b].root_visit = root_visit;
            p.hdr[b].n_blocks = n_blocks;
            p.hdr[b].n_tgroups = n_tgroups;
        }
        if (j < A) {  // zero-fill, then scatter the legal children to their action index
            if (p.visit_counts) p.visit_counts[(size_t)b * A + j] = 0;
            if (p.child_value) p.child_value[(size_t)b * A + j] = 0.0;
            if (p.child_prior) p.child_prior[(size_t)b * A + j] = 0.0;
            if (p.child_reward) p.child_reward[(size_t)b * A + j] = 0.0;
        }
        __syncwarp(gmask);
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
}

}  // namespace tsp
