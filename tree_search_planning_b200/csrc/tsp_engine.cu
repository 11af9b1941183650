// tsp_engine.cu -- host side of libtsp_b200.so: engine object, weight packing, MLP schedule
// construction, pool management, kernel launches and the extern "C" ABI of include/tsp.h.
// There is no CPU fallback in this file: every entry point either runs CUDA kernels or fails.
#include "tsp_device.cuh"
#include "tsp_fast.cuh"
#include "tsp_rootio.cuh"
#include "../../include/tsp.h"

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

using namespace tsp;

namespace {

thread_local std::string g_create_error;

struct HostLayer {
    int out_dim = 0, in_dim = 0;
    std::vector<float> W, b;
    int w_off = -1, b_off = -1, npad = 0, wld = 0;  // arena placement: Wt[in_dim][wld], bias[npad]
    // fragment-ordered arena of the fast search kernel (tsp_fast.cuh): tiles, bias[npad], one-hot rows
    int f_off = -1, fb_off = -1, fe_off = -1, f_dense = 0;
};

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
};

// ---- host-side schedule description, lowered to tsp::Schedule ----
struct HOp {
    int mlp, layer, in_buf, out_buf, act, extra;
};
struct HStage {
    std::vector<HOp> ops;
    int rop = 0, r_in = -1, r_out = -1, r_n = 0;
};
struct HSched {
    std::vector<int> width;  // per buffer
    std::vector<int> ready;  // first stage at which the buffer can be read
    std::vector<HStage> stages;
    int in_buf = -1, hidden = -1, pl = -1, vl = -1, rl = -1, tl = -1, tp = -1;
    int in_dim = 0;
    int new_buf(int w, int ready_stage) {
        width.push_back((w + 7) & ~7);  // layers are written in 8-column mma tiles
        ready.push_back(ready_stage);
        return (int)width.size() - 1;
    }
    HStage& stage(int s) {
        while ((int)stages.size() <= s) stages.emplace_back();
        return stages[s];
    }
};

}  // namespace

struct tsp_engine {
    tsp_config cfg{};
    std::string err;
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    // HOST-memory searches: the observations are uploaded in row chunks on a second stream and the root network of a
    // chunk starts as soon as its rows have arrived (copy / compute overlap inside tsp_run)
    cudaStream_t copy_stream = nullptr;
    static constexpr int OBS_CHUNKS = 4;
    cudaEvent_t ev_chunk[OBS_CHUNKS] = {};
    cudaEvent_t ev_root_out = nullptr;  // the root kernel's HOST outputs have left on the copy stream
    // Two sets of input staging buffers, used alternately: the inputs of run i + 1 cross PCIe on the copy stream while
    // the search of run i (which still reads its own set: tie / chance streams, noise) is running.
    struct InSet {
        DevBuf obs, legal, nlegal, noise, tie, chance;
        cudaEvent_t done = nullptr;  // recorded behind the search that read this set
    };
    InSet in_set[2];
    int in_flip = 0;
    bool no_overlap = false;  // TSP_NO_OVERLAP
    bool no_balance = false;  // TSP_NO_BALANCE: G groups per CTA in every wave
    bool no_merge = false;    // TSP_TC_NOMERGE: chance-node searches evaluate the futures of a transition expansion one by one
    bool no_zero_copy = false;  // TSP_NO_ZERO_COPY: stage + copy HOST outputs even when they are pinned
    // ring of event triples (before root kernel, between, after search kernel), drained lazily
    static constexpr int EV_RING = 1024;
    std::vector<cudaEvent_t> ev;
    int ev_head = 0, ev_pending = 0;
    std::vector<HostLayer> mlp[TSP_NUM_MLP];
    bool finalized = false;
    int nc = 5;  // node-block child capacity (5 or 8)
    int tm_override = 0;
    bool no_wsmem = false;
    int xw_override = 0;
    int ct_override = 0;
    int ng_override = 0;
    bool warp_mlp = false;
    // trees currently held by the pools (for tsp_run_args.resume)
    int tree_roots = 0, tree_sims = 0;
    bool tree_fed = false;
    bool no_fast = false;
    bool no_wlo = false;
    bool no_wglobal = false, force_wglobal = false;
    int fast_groups = 0;  // TSP_FAST_GROUPS override (1, 2 or 4 groups of 16 trees per CTA)

    // device memory
    DevBuf arena;  // packed weights
    int64_t arena_floats = 0;
    int64_t search_floats = 0;  // leading part of the arena used by the search kernel (everything but the representation net)
    // fast search kernel: fragment-ordered weights; one FastSched per group shape ([0]: 8 warps / 16 lanes per tree,
    // [1]: 4 warps / 8 lanes per tree)
    // [shape][0]: recurrent inference (deterministic) or one future of a transition expansion; [shape][1]: prediction
    DevBuf farena, fsched_dev[2][2];
    int64_t farena_floats = 0;
    FastSched fsched[2][2]{};
    bool has_fast = false;
    int fast_lpt = 0;     // TSP_FAST_LPT override (8 or 16 lanes per tree)
    // tcgen05 search MLP (tsp_tc.cuh): [0] latency shape (2 groups x 8 warps, replicated weights), [1] throughput
    // shape (4 groups x 4 warps); TMEM image of the weights and the bias / one-hot table per shape
    TcSched tcs[2]{};
    DevBuf tc_img[2], tc_ctab[2], tc_status;
    bool has_tc = false;
    bool no_tc = false;   // TSP_NO_TC: keep the mma.sync kernels
    bool tc_nofold = false;  // TSP_TC_NOFOLD: evaluate the first reward layer as its own MMA on the un-normalised state
    // streamed tcgen05 search MLP (tsp_stream.cuh): networks too large for tensor memory (cross-merge). fp16-split weight
    // image in global memory (16 KB tiles in TMA order), bias / one-hot table, tile / epilogue tables, the image's tensor map
    TsSched tss{};
    DevBuf ts_img, ts_ctab;
    CUtensorMap ts_map{};
    bool has_tcs = false;
    // streamed tcgen05 root network (root_tcs_kernel): initial_inference for every fully-connected network of the reference
    TsSched rts{};
    DevBuf rts_img, rts_ctab, rts_hidden;
    CUtensorMap rts_map{};
    bool has_rtcs = false;
    bool no_rtcs = false;  // TSP_NO_RTCS: keep net_kernel (mma.sync) for the root
    bool separate_heads = false;  // MuZeroStochastic: one dynamics MLP per future (TSP_MLP_DYNAMICS_F1..)
    int f_stride = 0;             // floats between the per-future weight groups {dynamics z, reward, prior} of the arena
    bool no_tcs = false;  // TSP_NO_TCS: keep the mma.sync kernels for these networks
    DevBuf blocks, hidden, hdr, rootprior, rootact, pbtable, phase;
    bool phase_timing = false;
    int blocks_per_tree = 0, hidden_per_tree = 0, max_iters = 0;
    // staging for HOST-memory calls
    DevBuf s_obs, s_legal, s_nlegal, s_noise, s_tie, s_chance, s_fedroot, s_fed;
    DevBuf s_visits, s_rootv, s_cval, s_cprior, s_crew, s_depth, s_rpred, s_rhid, s_nrec, s_troot, s_trace, s_tdepth;
    DevBuf n_in, n_act, n_o[8];

    Schedule sch_root{}, sch_recur{}, sch_trans{}, sch_pred{};
    bool has_root = false, has_recur = false, has_trans = false, has_pred = false;

    tsp_stats stats{};
    int64_t pool_bytes = 0;
};

namespace {

int fail(tsp_engine* e, int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (e)
        e->err = buf;
    else
        g_create_error = buf;
    return code;
}

#define CK(e, call)                                                                                   \
    do {                                                                                              \
        cudaError_t _st = (call);                                                                     \
        if (_st != cudaSuccess)                                                                       \
            return fail(e, TSP_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_st), __FILE__, __LINE__); \
    } while (0)

int ensure(tsp_engine* e, DevBuf& b, size_t bytes) {
    if (bytes <= b.cap) return TSP_OK;
    if (b.p) {
        CK(e, cudaStreamSynchronize(e->stream));
        if (e->copy_stream) CK(e, cudaStreamSynchronize(e->copy_stream));
        CK(e, cudaFree(b.p));
        e->pool_bytes -= (int64_t)b.cap;
        b.p = nullptr;
        b.cap = 0;
    }
    cudaError_t st = cudaMalloc(&b.p, bytes);
    if (st != cudaSuccess) {
        b.p = nullptr;
        return fail(e, TSP_ENOMEM, "cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(st));
    }
    b.cap = bytes;
    e->pool_bytes += (int64_t)bytes;
    return TSP_OK;
}

void release(DevBuf& b) {
    if (b.p) cudaFree(b.p);
    b.p = nullptr;
    b.cap = 0;
}

// ---- schedule construction ----
// Appends the Linear layers of `mlp_id` reading `in_buf`; returns the output buffer.
int add_chain(tsp_engine* e, HSched& hs, int mlp_id, int in_buf, int extra_first) {
    const auto& layers = e->mlp[mlp_id];
    int cur = in_buf;
    for (size_t l = 0; l < layers.size(); ++l) {
        int s = hs.ready[cur];
        while ((int)hs.stage(s).ops.size() >= MAX_OPS) ++s;
        const int out = hs.new_buf(layers[l].out_dim, s + 1);
        HOp op{mlp_id, (int)l, cur, out, l + 1 < layers.size() ? 1 : 0, l == 0 ? extra_first : 0};
        hs.stage(s).ops.push_back(op);
        cur = out;
    }
    return cur;
}

int add_normalise(HSched& hs, int in_buf, int n) {
    int s = hs.ready[in_buf];
    while (hs.stage(s).rop != 0) ++s;
    const int out = hs.new_buf(n, s + 1);
    HStage& st = hs.stage(s);
    st.rop = 1;
    st.r_in = in_buf;
    st.r_out = out;
    st.r_n = n;
    return out;
}

// Liveness-based first-fit placement of the activation buffers inside a row strip: two buffers may share
// space iff their [first write, last read] stage ranges are disjoint. Returns the strip width in floats.
int strip_offsets(const HSched& hs, const std::vector<int>& width, bool keep_input, std::vector<int>& off) {
    const int nb = (int)hs.width.size();
    const int ns = (int)hs.stages.size();
    const int INF = 1 << 20;
    std::vector<int> def(nb, INF), last(nb, -1);
    off.assign(nb, -1);
    def[hs.in_buf] = -1;
    for (int s = 0; s < ns; ++s) {
        for (const HOp& op : hs.stages[s].ops) {
            def[op.out_buf] = std::min(def[op.out_buf], s);
            last[op.in_buf] = std::max(last[op.in_buf], s);
        }
        if (hs.stages[s].rop) {
            def[hs.stages[s].r_out] = std::min(def[hs.stages[s].r_out], s);
            last[hs.stages[s].r_in] = std::max(last[hs.stages[s].r_in], s);
        }
    }
    for (int o : {hs.hidden, hs.pl, hs.vl, hs.rl, hs.tl, hs.tp})
        if (o >= 0) last[o] = INF;
    last[hs.in_buf] = std::max(last[hs.in_buf], 0);
    if (keep_input) last[hs.in_buf] = INF;  // the schedule is evaluated once per future z on the same input
    std::vector<int> order(nb);
    for (int i = 0; i < nb; ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return def[a] < def[b]; });
    int total = 0;
    for (int idx = 0; idx < nb; ++idx) {
        const int b = order[idx];
        int pos = 0;
        bool moved = true;
        while (moved) {
            moved = false;
            for (int j = 0; j < idx; ++j) {
                const int o = order[j];
                const bool live_overlap = !(last[o] < def[b] || last[b] < def[o]);
                if (live_overlap && pos < off[o] + width[o] && off[o] < pos + width[b]) {
                    pos = off[o] + width[o];
                    moved = true;
                }
            }
        }
        off[b] = pos;
        total = std::max(total, pos + width[b]);
    }
    return total;
}

int lower(tsp_engine* e, const HSched& hs, Schedule& sc, bool keep_input = false) {
    const int ns = (int)hs.stages.size();
    if (ns > MAX_STAGES) return fail(e, TSP_EINVAL, "network too deep: %d stages > %d", ns, MAX_STAGES);
    std::vector<int> off;
    int total = strip_offsets(hs, hs.width, keep_input, off);
    total += ((4 - total % 32) + 32) % 32;  // row stride == 4 (mod 32)
    if (total > 32000) return fail(e, TSP_EINVAL, "activation strip too wide (%d floats)", total);
    memset(&sc, 0, sizeof(sc));
    sc.n_stages = ns;
    sc.row_stride = total;
    sc.in_dim = hs.in_dim;
    auto o16 = [&](int b) { return (short)(b >= 0 ? off[b] : -1); };
    sc.off_in = o16(hs.in_buf);
    sc.off_hidden = o16(hs.hidden);
    sc.off_pl = o16(hs.pl);
    sc.off_vl = o16(hs.vl);
    sc.off_rl = o16(hs.rl);
    sc.off_tl = o16(hs.tl);
    sc.off_tp = o16(hs.tp);
    for (int s = 0; s < ns; ++s) {
        Stage& st = sc.st[s];
        const HStage& h = hs.stages[s];
        st.n_ops = (unsigned char)h.ops.size();
        st.row_op = (unsigned char)h.rop;
        if (h.rop) {
            st.r_in = o16(h.r_in);
            st.r_out = o16(h.r_out);
            st.r_n = (short)h.r_n;
        }
        for (size_t i = 0; i < h.ops.size(); ++i) {
            const HOp& op = h.ops[i];
            const HostLayer& L = e->mlp[op.mlp][op.layer];
            LayerOp& d = st.ops[i];
            d.w_off = L.w_off;
            d.b_off = L.b_off;
            d.N = (short)L.out_dim;
            d.Npad = (short)L.npad;
            d.wld = (short)L.wld;
            d.in_off = o16(op.in_buf);
            d.out_off = o16(op.out_buf);
            d.act = (unsigned char)op.act;
            d.extra = (unsigned char)op.extra;
            int k = L.in_dim;
            if (op.extra == 1) k -= e->cfg.num_actions;
            if (op.extra == 2) k -= e->cfg.num_actions + e->cfg.n_futures;
            d.K = (short)k;
        }
    }
    return TSP_OK;
}

// Fast search kernel (tsp_fast.cuh): lower a schedule to per-stage, per-warp lists of 16 x 8 output tiles.
// Returns TSP_OK with e->has_fast == false when the network does not fit the fast path's constraints
// (the generic search kernel then runs it).
int lower_fast(tsp_engine* e, const HSched& hs, int which, int GWARPS, int slot = 0, bool keep_input = false) {
    e->has_fast = false;
    const int ns = (int)hs.stages.size();
    if (ns > FAST_MAX_STAGES) return TSP_OK;
    std::vector<int> width(hs.width.size());
    for (size_t i = 0; i < width.size(); ++i) width[i] = (hs.width[i] + 15) & ~15;  // A operands are read in chunks of 16
    std::vector<int> off;
    int total = strip_offsets(hs, width, keep_input, off);
    // two rows are interleaved per strip line: line stride 2 * total == 16 (mod 32) words keeps the LDS.128 /
    // STS.128 of lanes g, g + 1 conflict-free
    total += ((8 - total % 16) + 16) % 16;
    if (off[hs.in_buf] != 0 || total > 4000) return TSP_OK;
    FastSched& fs = e->fsched[which][slot];
    memset(&fs, 0, sizeof(fs));
    fs.n_stages = ns;
    fs.row_stride = 2 * total;
    auto o16 = [&](int b) { return (short)(b >= 0 ? off[b] : -1); };
    // buffers that a later layer reads as an mma operand get their lo parts stored by the producer
    std::vector<char> is_operand(hs.width.size(), 0);
    for (const HStage& h : hs.stages)
        for (const HOp& op : h.ops) is_operand[op.in_buf] = 1;
    fs.off_hidden = o16(hs.hidden);
    fs.off_pl = o16(hs.pl);
    fs.off_vl = o16(hs.vl);
    fs.off_rl = o16(hs.rl);
    fs.off_tl = o16(hs.tl);
    fs.off_tp = o16(hs.tp);
    int nu = 0;
    for (int s = 0; s < ns; ++s) {
        const HStage& h = hs.stages[s];
        fs.stage[s] = make_int4(h.rop, h.rop ? 2 * off[h.r_in] : 0, h.rop ? 2 * off[h.r_out] : 0, h.r_n);
        std::vector<FUnit> units;
        for (const HOp& op : h.ops) {
            const HostLayer& L = e->mlp[op.mlp][op.layer];
            if (L.f_off < 0 || L.f_dense % 16) return TSP_OK;
            const int kc = L.f_dense / 16;
            if (kc > 255) return TSP_OK;
            for (int jt = 0; jt < L.npad / 8; ++jt) {
                FUnit u;
                memset(&u, 0, sizeof(u));
                u.x_off = 2 * off[op.in_buf];
                u.w_off = L.f_off + jt * kc * 128;
                u.y_off = 2 * (off[op.out_buf] + 8 * jt);
                u.flags = kc | (op.act ? 0x100 : 0) | (is_operand[op.out_buf] ? 0x200 : 0) | (op.extra ? 0x400 : 0) |
                          (op.extra == 2 ? 0x800 : 0);
                u.b_off = L.fb_off + 8 * jt;
                u.e_off = op.extra ? L.fe_off + 8 * jt : 0;
                u.e_ld = L.npad;
                u.e_fut = e->cfg.num_actions;
                units.push_back(u);
            }
        }
        if (nu + (int)units.size() > FAST_MAX_TILES) return TSP_OK;
        // longest-processing-time assignment to the warps of a group: heaviest (deepest K) units first
        std::vector<int> order(units.size());
        for (size_t i = 0; i < order.size(); ++i) order[i] = (int)i;
        std::stable_sort(order.begin(), order.end(),
                         [&](int a, int b) { return (units[a].flags & 0xFF) > (units[b].flags & 0xFF); });
        std::vector<std::vector<int>> mine(GWARPS);
        std::vector<int> load(GWARPS, 0);
        for (int idx : order) {
            int w = 0;
            for (int k = 1; k < GWARPS; ++k)
                if (load[k] < load[w]) w = k;
            mine[w].push_back(idx);
            load[w] += units[idx].flags & 0xFF;
        }
        for (int w = 0; w < GWARPS; ++w) {
            fs.ulist[s][w] = nu | ((int)mine[w].size() << 16);
            for (int idx : mine[w]) fs.units[nu++] = units[idx];
        }
    }
    fs.n_units = nu;
    fs.w_floats = (int)e->farena_floats;
    int rc = ensure(e, e->fsched_dev[which][slot], sizeof(FastSched));
    if (rc) return rc;
    CK(e, cudaMemcpyAsync(e->fsched_dev[which][slot].p, &fs, sizeof(FastSched), cudaMemcpyHostToDevice, e->stream));
    CK(e, cudaStreamSynchronize(e->stream));
    e->has_fast = true;
    return TSP_OK;
}

// Pack the search MLPs (everything but the representation network) in fragment order for tsp_fast.cuh.
int pack_fast_arena(tsp_engine* e) {
    const tsp_config& c = e->cfg;
    e->farena_floats = 0;
    const bool chance = c.variant != TSP_VARIANT_MUZERO;
    auto skipped = [&](int m) {
        if (m == TSP_MLP_REPRESENTATION || m >= TSP_MLP_DYNAMICS_F1) return true;
        if (m == TSP_MLP_PRIOR) return !chance;
        if (m == TSP_MLP_TERMINAL) return chance || !c.mask_absorbing_states;
        return false;
    };
    int64_t total = 0;
    for (int m = 0; m < TSP_NUM_MLP; ++m) {
        if (skipped(m)) continue;
        auto& v = e->mlp[m];
        for (size_t l = 0; l < v.size(); ++l) {
            HostLayer& L = v[l];
            const int n_extra = (m == TSP_MLP_DYNAMICS && l == 0) ? c.num_actions + (chance ? c.n_futures : 0) : 0;
            L.f_dense = L.in_dim - n_extra;
            L.f_off = -1;
            if (L.f_dense <= 0 || L.f_dense % 16) continue;
            L.f_off = (int)total;
            total += (int64_t)(L.npad / 8) * (L.f_dense / 16) * 128;
            L.fb_off = (int)total;
            total += L.npad;
            L.fe_off = (int)total;
            total += (int64_t)n_extra * L.npad;
        }
    }
    total = (total + 3) & ~(int64_t)3;
    if (total == 0) return TSP_OK;
    std::vector<float> host((size_t)total, 0.0f);
    for (int m = 0; m < TSP_NUM_MLP; ++m)
        for (HostLayer& L : e->mlp[m]) {
            if (skipped(m) || L.f_off < 0) continue;
            const int kc = L.f_dense / 16;
            // tile jt, chunk ch, lane (g, t): W[8 jt + g][16 ch + {2 t, 2 t + 1, 8 + 2 t, 9 + 2 t}]
            for (int jt = 0; jt < L.npad / 8; ++jt)
                for (int ch = 0; ch < kc; ++ch)
                    for (int lane = 0; lane < 32; ++lane) {
                        const int g = lane >> 2, t = lane & 3, n = 8 * jt + g;
                        float* dst = &host[(size_t)L.f_off + ((size_t)jt * kc + ch) * 128 + lane * 4];
                        const int kk[4] = {2 * t, 2 * t + 1, 8 + 2 * t, 9 + 2 * t};
                        for (int q = 0; q < 4; ++q)
                            dst[q] = n < L.out_dim ? L.W[(size_t)n * L.in_dim + 16 * ch + kk[q]] : 0.0f;
                    }
            for (int o = 0; o < L.out_dim; ++o) host[(size_t)L.fb_off + o] = L.b[o];
            for (int a = 0; a < L.in_dim - L.f_dense; ++a)
                for (int o = 0; o < L.out_dim; ++o)
                    host[(size_t)L.fe_off + (size_t)a * L.npad + o] = L.W[(size_t)o * L.in_dim + L.f_dense + a];
        }
    // second half: lo parts (x - trunc_tf32(x), the same float32 expression the kernel evaluates) of every value,
    // same layout; only the weight fragments of it are read, and only when it fits in shared memory
    host.resize((size_t)2 * total);
    for (int64_t i = 0; i < total; ++i) {
        uint32_t bits;
        memcpy(&bits, &host[(size_t)i], 4);
        bits &= 0xFFFFE000u;
        float hi;
        memcpy(&hi, &bits, 4);
        host[(size_t)(total + i)] = host[(size_t)i] - hi;
    }
    int rc = ensure(e, e->farena, host.size() * sizeof(float));
    if (rc) return rc;
    CK(e, cudaMemcpyAsync(e->farena.p, host.data(), host.size() * sizeof(float), cudaMemcpyHostToDevice, e->stream));
    CK(e, cudaStreamSynchronize(e->stream));
    e->farena_floats = total;
    return TSP_OK;
}

// tcgen05 search MLP (tsp_tc.cuh): place the weights of the recurrent inference in tensor memory and lower the
// network to MMA / epilogue tables. Supported: the fully-connected MuZero network with one hidden layer in the
// dynamics and in every head (all game files of the reference: fc_*_layers = [n]), widths that are multiples of 8,
// H <= 64, head widths <= 32, no terminal head. Anything else keeps the mma.sync kernels (e->has_tc == false).
int build_tc(tsp_engine* e) {
    e->has_tc = false;
    const tsp_config& c = e->cfg;
    if (c.variant != TSP_VARIANT_MUZERO || e->nc != 5 || c.mask_absorbing_states) return TSP_OK;
    for (int m : {TSP_MLP_DYNAMICS, TSP_MLP_REWARD, TSP_MLP_POLICY, TSP_MLP_VALUE})
        if (e->mlp[m].size() != 2) return TSP_OK;
    const HostLayer &D1 = e->mlp[TSP_MLP_DYNAMICS][0], &D2 = e->mlp[TSP_MLP_DYNAMICS][1];
    const HostLayer &R1 = e->mlp[TSP_MLP_REWARD][0], &R2 = e->mlp[TSP_MLP_REWARD][1];
    const HostLayer &P1 = e->mlp[TSP_MLP_POLICY][0], &P2 = e->mlp[TSP_MLP_POLICY][1];
    const HostLayer &V1 = e->mlp[TSP_MLP_VALUE][0], &V2 = e->mlp[TSP_MLP_VALUE][1];
    const int H = c.encoding_size, A = c.num_actions;
    const int DH = D1.out_dim, RH = R1.out_dim, PH = P1.out_dim, VH = V1.out_dim;
    const int NR = R2.out_dim, NV = V2.out_dim;
    if (H % 8 || H > 64 || H < 8 || DH % 8 || DH > 32 || RH % 8 || PH % 8 || VH % 8 || RH > 32 || PH > 32 || VH > 32) return TSP_OK;
    if (NR > 32 || NV > 32 || A > 32 || D1.in_dim != H + A || D2.out_dim != H) return TSP_OK;
    const int HK = std::max(RH, std::max(PH, VH));
    auto bufsz = [](int K) { return 2 * (K / 4) * TC_LBO; };
    auto sbo = [](int K) { return (K / 4) * TC_LBO; };
    for (int shape = 0; shape < 2; ++shape) {
        TcSched& ts = e->tcs[shape];
        memset(&ts, 0, sizeof(ts));
        const bool rep = shape == 0;  // replicate small layers over the lane quarters (latency shape)
        const int G = shape == 0 ? 2 : 4;
        // The second dynamics layer has no activation (models.py:1150-1162: Identity at the output), so the first reward
        // layer on the un-normalised state is a Linear layer of the dynamics hidden layer: Wr1 (W2 d + b2) + br1 =
        // (Wr1 W2) d + (Wr1 b2 + br1). The product (float64 on the host, rounded once to float32) rides in the MMA of
        // dynamics layer 2 as RH more rows -- one MMA sequence and one stage less per simulation.
        const bool fold = !e->tc_nofold && H + RH <= 128;
        std::vector<float> FW((size_t)RH * DH), Fb((size_t)RH);
        for (int r = 0; r < RH; ++r) {
            for (int k = 0; k < DH; ++k) {
                double acc = 0.0;
                for (int h = 0; h < H; ++h) acc += (double)R1.W[(size_t)r * H + h] * (double)D2.W[(size_t)h * DH + k];
                FW[(size_t)r * DH + k] = (float)acc;
            }
            double acc = R1.b[(size_t)r];
            for (int h = 0; h < H; ++h) acc += (double)R1.W[(size_t)r * H + h] * (double)D2.b[(size_t)h];
            Fb[(size_t)r] = (float)acc;
        }
        // ---- TMEM columns ----
        const int c1 = 0, c2 = c1 + H, c3 = rep ? c2 + DH : c1, c4 = (rep ? c3 + H : c2 + DH);
        const int hi_cols = c4 + HK;
        ts.a_lo = hi_cols;
        ts.n_cols = 2 * hi_cols;
        ts.d_col0 = ts.n_cols;
        ts.d_cols_group = 48;
        if (ts.n_cols % 8 || ts.d_col0 + G * ts.d_cols_group > TC_TMEM_COLS) return TSP_OK;
        // ---- B arena of a group ----
        int o = 0;
        const int xh = o; o += bufsz(H);
        const int xd = o; o += bufsz(DH);
        const int xs = o; o += bufsz(H);
        const int xn = o; o += bufsz(H);
        const int xr = o; o += bufsz(RH);
        const int xp = o; o += bufsz(PH);
        const int xv = o; o += bufsz(VH);
        ts.b_lo = o;
        ts.arena_bytes = (2 * o + 127) & ~127;
        ts.xh_off = xh; ts.xh_sbo = sbo(H);
        ts.xs_off = xs; ts.xs_sbo = sbo(H);
        ts.xn_off = xn; ts.xn_sbo = sbo(H);
        ts.H = H;
        // ---- out strip ----
        ts.off_vl = 0;
        ts.off_rl = NV;
        ts.off_pl = NV + NR;
        int pitch = NV + NR + A;
        if (shape == 0) while (pitch % 32 != 16) ++pitch;
        else while (pitch % 32 != 8 && pitch % 32 != 24) ++pitch;
        ts.out_pitch = pitch;
        // ---- constant table ----
        std::vector<float> ctab;
        auto put = [&](const std::vector<float>& v, int n) {
            const int at = (int)ctab.size();
            for (int i = 0; i < n; ++i) ctab.push_back(v[(size_t)i]);
            while (ctab.size() % 32) ctab.push_back(0.0f);  // lanes past `rows` still form an address
            return at;
        };
        const int t_b1 = put(D1.b, DH);
        std::vector<float> e1((size_t)A * 32, 0.0f);
        for (int a = 0; a < A; ++a)
            for (int r = 0; r < DH; ++r) e1[(size_t)a * 32 + r] = D1.W[(size_t)r * D1.in_dim + H + a];
        const int t_e1 = put(e1, A * 32);
        const int t_b2 = put(D2.b, H);
        const int t_b3r = put(fold ? Fb : R1.b, RH), t_b3p = put(P1.b, PH), t_b3v = put(V1.b, VH);
        const int t_b4r = put(R2.b, NR), t_b4v = put(V2.b, NV), t_b4p = put(P2.b, A);
        for (int i = 0; i < 64; ++i) ctab.push_back(0.0f);
        if ((int)ctab.size() > TC_MAX_CTAB) return TSP_OK;
        ts.ctab_floats = (int)ctab.size();
        // ---- TMEM image [column][lane] ----
        std::vector<float> img((size_t)ts.n_cols * 128, 0.0f);
        auto place = [&](const HostLayer& L, int col0, int lane0, int K) {  // rows of L at lanes lane0.., k < K at columns col0..
            for (int r = 0; r < L.out_dim; ++r)
                for (int k = 0; k < K; ++k) img[(size_t)(col0 + k) * 128 + lane0 + r] = L.W[(size_t)r * L.in_dim + k];
        };
        const int h_q = (H + 31) / 32;  // lane quarters of the next state; the folded reward layer sits behind them
        auto place_fold = [&](int col0, int lane0) {
            for (int r = 0; r < RH; ++r)
                for (int k = 0; k < DH; ++k) img[(size_t)(col0 + k) * 128 + lane0 + r] = FW[(size_t)r * DH + k];
        };
        if (rep) {
            for (int q = 0; q < 4; ++q) place(D1, c1, 32 * q, H);
            if (fold) {
                place(D2, c2, 0, DH);
                place_fold(c2, 32 * h_q);
                place(P1, c3, 0, H);
                place(V1, c3, 32, H);
            } else {
                const int r2 = H <= 32 ? 4 : 2;
                for (int q = 0; q < r2; ++q) place(D2, c2, (128 / r2) * q, DH);
                place(R1, c3, 0, H); place(P1, c3, 32, H); place(V1, c3, 64, H);
            }
        } else if (fold) {  // lane quarter 3 stays free in every stage: its warp issues the MMAs
            place(D1, c1, 0, H); place(P1, c1, 32, H); place(V1, c1, 64, H);
            place(D2, c2, 0, DH);
            place_fold(c2, 32 * h_q);
        } else {
            place(D1, c1, 0, H); place(P1, c1, 64, H); place(V1, c1, 96, H);
            place(D2, c2, 0, DH);
            place(R1, c1, 32, H);
        }
        place(R2, c4, 0, RH); place(V2, c4, 32, VH); place(P2, c4, 64, PH);
        for (int col = 0; col < hi_cols; ++col)
            for (int l = 0; l < 128; ++l) {
                const float w = img[(size_t)col * 128 + l];
                uint32_t bits;
                memcpy(&bits, &w, 4);
                bits &= 0xFFFFE000u;
                float hi;
                memcpy(&hi, &bits, 4);
                img[(size_t)(hi_cols + col) * 128 + l] = w - hi;
            }
        // ---- MMAs ----
        auto mma = [&](int i, int a_col, int b_off, int K, int d_col) {
            ts.mma[i].a_col = a_col; ts.mma[i].b_off = b_off; ts.mma[i].ksteps = K / 8; ts.mma[i].sbo = sbo(K); ts.mma[i].d_col = d_col;
        };
        const int DA = 0, DB = 16, DC = 32;
        int n_m = 0;
        const int m_d1 = n_m++; mma(m_d1, c1, xh, H, DA);
        const int m_d2 = n_m++; mma(m_d2, c2, xd, DH, DB);
        int m_r1 = -1;
        if (!fold) { m_r1 = n_m++; mma(m_r1, c3, xs, H, DA); }
        const int m_pv = n_m++; mma(m_pv, c3, xn, H, DC);
        const int m_out = n_m;
        mma(n_m++, c4, xr, RH, DA);
        mma(n_m++, c4, xv, VH, DB);
        mma(n_m++, c4, xp, PH, DC);
        // ---- epilogue passes ----
        // a pass is described per TMEM lane quarter q; with 8 warps per group (latency shape) warps q and q + 4 split
        // the quarter's trees
        auto epi = [&](int pass, int q, int d_col, int n0, int ncnt, int rows, int dst, int dst_sbo, int feat0, int cidx, int eidx,
                       int flags) {
            const int parts = rep ? 2 : 1;
            for (int h = 0; h < parts; ++h) {
                TcEpi& x = ts.epi[pass][q + 4 * h];
                x.d_col = (short)d_col; x.n0 = (short)(n0 + h * (ncnt / parts)); x.ncnt = (short)(ncnt / parts); x.rows = (short)rows;
                x.dst = dst; x.dst_sbo = dst_sbo; x.feat0 = feat0; x.cidx = cidx; x.eidx = eidx; x.e_ld = 32; x.zidx = -1; x.flags = flags;
            }
        };
        if (rep) {
            if (fold) {
                // quarters 0 and 1 (two replicas of dynamics layer 1, 8 trees each, split over warps q and q + 4)
                for (int q = 0; q < 2; ++q) epi(0, q, DA, 8 * q, 8, DH, xd, sbo(DH), 0, t_b1, t_e1, 1);
                for (int q = 0; q < h_q; ++q) epi(1, q, DB, 0, 16, std::min(32, H - 32 * q), xs, sbo(H), 32 * q, t_b2 + 32 * q, -1, 0);
                epi(1, h_q, DB, 0, 16, RH, xr, sbo(RH), 0, t_b3r, -1, 1);
                epi(2, 0, DC, 0, 16, PH, xp, sbo(PH), 0, t_b3p, -1, 1);
                epi(2, 1, DC, 0, 16, VH, xv, sbo(VH), 0, t_b3v, -1, 1);
            } else {
                for (int q = 0; q < 4; ++q) epi(0, q, DA, 4 * q, 4, DH, xd, sbo(DH), 0, t_b1, t_e1, 1);
                if (H <= 32) {
                    for (int q = 0; q < 4; ++q) epi(1, q, DB, 4 * q, 4, H, xs, sbo(H), 0, t_b2, -1, 0);
                } else {
                    for (int q = 0; q < 4; ++q)
                        epi(1, q, DB, 8 * (q >> 1), 8, std::min(32, H - 32 * (q & 1)), xs, sbo(H), 32 * (q & 1), t_b2 + 32 * (q & 1), -1, 0);
                }
                epi(2, 0, DA, 0, 16, RH, xr, sbo(RH), 0, t_b3r, -1, 1);
                epi(2, 1, DC, 0, 16, PH, xp, sbo(PH), 0, t_b3p, -1, 1);
                epi(2, 2, DC, 0, 16, VH, xv, sbo(VH), 0, t_b3v, -1, 1);
            }
        } else {
            epi(0, 0, DA, 0, 16, DH, xd, sbo(DH), 0, t_b1, t_e1, 1);
            for (int q = 0; q < h_q; ++q) epi(1, q, DB, 0, 16, std::min(32, H - 32 * q), xs, sbo(H), 32 * q, t_b2 + 32 * q, -1, 0);
            if (fold) {
                epi(1, h_q, DB, 0, 16, RH, xr, sbo(RH), 0, t_b3r, -1, 1);
                epi(2, 1, DC, 0, 16, PH, xp, sbo(PH), 0, t_b3p, -1, 1);
                epi(2, 2, DC, 0, 16, VH, xv, sbo(VH), 0, t_b3v, -1, 1);
            } else {
                epi(2, 1, DA, 0, 16, RH, xr, sbo(RH), 0, t_b3r, -1, 1);
                epi(2, 2, DC, 0, 16, PH, xp, sbo(PH), 0, t_b3p, -1, 1);
                epi(2, 3, DC, 0, 16, VH, xv, sbo(VH), 0, t_b3v, -1, 1);
            }
        }
        epi(3, 0, DA, 0, 16, NR, ts.off_rl, 0, 0, t_b4r, -1, 2);
        epi(3, 1, DB, 0, 16, NV, ts.off_vl, 0, 0, t_b4v, -1, 2);
        epi(3, 2, DC, 0, 16, A, ts.off_pl, 0, 0, t_b4p, -1, 2);
        // ---- stages ----
        auto stage = [&](int s, int mma0, int n_mma, int commit, int row_op, int epi0, int n_epi) {
            ts.stage[s].mma0 = mma0; ts.stage[s].n_mma = n_mma; ts.stage[s].commit = commit; ts.stage[s].row_op = row_op;
            ts.stage[s].epi0 = epi0; ts.stage[s].n_epi = n_epi;
        };
        ts.n_stages = 5;
        // with the fold, lane quarter 3 has no epilogue share in any stage: the last warp of the group (quarter 3) issues
        const int issuer = fold ? (rep ? 7 : 3) : 0;
        for (int s2 = 0; s2 < TC_MAX_STAGES; ++s2) ts.stage[s2].issuer = issuer;
        stage(0, m_d1, 1, 1, 0, 0, 1);   // dynamics layer 1
        stage(1, m_d2, 1, 1, 0, 1, 1);   // dynamics layer 2 -> un-normalised next state (+ folded reward layer 1)
        if (fold) stage(2, 0, 0, 0, 1, 0, 0);     // the state is normalised
        else stage(2, m_r1, 1, 0, 1, 0, 0);       // reward layer 1 in flight while the state is normalised
        stage(3, m_pv, 1, 1, 0, 2, 1);   // policy / value layer 1; epilogue of the hidden layers
        stage(4, m_out, 3, 1, 0, 3, 1);  // output layers -> logits
        int rc;
        if ((rc = ensure(e, e->tc_img[shape], img.size() * sizeof(float)))) return rc;
        if ((rc = ensure(e, e->tc_ctab[shape], ctab.size() * sizeof(float)))) return rc;
        CK(e, cudaMemcpyAsync(e->tc_img[shape].p, img.data(), img.size() * sizeof(float), cudaMemcpyHostToDevice, e->stream));
        CK(e, cudaMemcpyAsync(e->tc_ctab[shape].p, ctab.data(), ctab.size() * sizeof(float), cudaMemcpyHostToDevice, e->stream));
        CK(e, cudaStreamSynchronize(e->stream));
    }
    int rc;
    if ((rc = ensure(e, e->tc_status, sizeof(int)))) return rc;
    CK(e, cudaMemset(e->tc_status.p, 0, sizeof(int)));
    e->has_tc = true;
    return TSP_OK;
}

// Streamed tcgen05 search MLP (tsp_stream.cuh): the weights of the recurrent inference as fp16-split tiles in TMA order
// plus the tile / epilogue tables. Supported: the fully-connected MuZero network with one hidden layer in the dynamics
// and in every head, 64 < H <= 256 (multiples of 64), dynamics hidden width <= 128 (multiples of 64), head widths
// <= 32 (multiples of 16), no terminal head: the cross-merge network (games/cross_merge_env.py:112-118). Smaller
// networks live in tensor memory (build_tc); anything else keeps the mma.sync kernels (e->has_tcs == false).
typedef CUresult (*TspEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// v = v1 + v2 / 2048 (the same split the kernel applies to activations, tsp_stream.cuh)
void ts_split_host(float v, __half& h1, __half& h2) {
    h1 = __float2half_rn(v);
    h2 = __float2half_rn((v - __half2float(h1)) * TS_SPLIT);
}

size_t root_tcs_smem(const TsSched& ts) {
    return (size_t)((ts.ctab_floats + 3) & ~3) * 4 + sizeof(TsCtl) + 256 + (size_t)TS_NS * TS_SLOT + (size_t)ts.arena_bytes +
           (size_t)TR_N * ts.out_pitch * 4;
}

// Tables and weight image of a streamed tcgen05 MLP (tsp_stream.cuh): tiles of 16 KB = `rows` x KT inputs (128 x 64, 64 x 128
// or 32 x 256), element (m, k) of a tile at ((m / 8) * (KT / 8) + k / 8) * 64 + (m % 8) * 8 + k % 8 (K-major core matrices:
// LBO 128, SBO 128 * KT / 8); a tile is 128 image rows of 128 bytes.
struct TsBuilder {
    TsSched& ts;
    int nt;  // columns of the MMAs: trees / rows per CTA
    std::vector<__half> img;
    int n_tiles = 0, n_acc = 0;
    bool overflow = false;
    static int box(int rows) { return rows > 64 ? 128 : (rows > 32 ? 64 : 32); }
    static int mn_sbo(int K) { return (K / 8) * 128; }
    // One accumulator tile `acc` of `rows_valid` rows over K inputs: K chunks of KT inputs, per chunk the W1 tile then the
    // W2 tile. W(m, k): weight of row m, input k. The B operand starts at b_off (X1 image; X2 follows (nt / 8) * b_sbo on).
    template <typename WF>
    void add_layer(int acc, int rows_valid, int rows_box, int K, WF W, int b_off, int b_sbo, bool b_mn, int d_col) {
        const int KT = 8192 / rows_box;
        for (int k0 = 0; k0 < K; k0 += KT) {
            const int kw = std::min(KT, K - k0);
            for (int part = 0; part < 2; ++part) {
                if (n_tiles >= TS_MAX_TILES) { overflow = true; return; }
                TsTile& tl = ts.tile[n_tiles++];
                tl.img_row = (int)(img.size() / 64);
                tl.a_sbo = (short)(128 * (KT / 8));
                tl.ksteps = (short)((kw + 15) / 16);
                const int lbo = b_mn ? 128 : TS_X0_LBO;
                tl.b_off = b_off + (k0 / 8) * lbo;
                tl.b_lo = (nt / 8) * b_sbo;
                tl.b_sbo = b_sbo;
                tl.b_lbo = (short)lbo;
                tl.b_step = (short)((2 * lbo) >> 4);
                tl.d_col = (short)d_col;
                tl.kind = (char)part;
                tl.first = (char)(k0 == 0 && part == 0);
                tl.b_mn = (char)(b_mn ? 1 : 0);
                tl.acc_done = (char)((k0 + KT >= K && part == 1) ? acc + 1 : 0);
                const size_t base = img.size();
                img.resize(base + 8192, __float2half_rn(0.0f));
                for (int m = 0; m < rows_valid; ++m)
                    for (int kk = 0; kk < kw; ++kk) {
                        __half h1, h2;
                        ts_split_host(W(m, k0 + kk), h1, h2);
                        img[base + ((size_t)(m >> 3) * (KT / 8) + (kk >> 3)) * 64 + (m & 7) * 8 + (kk & 7)] = part == 0 ? h1 : h2;
                    }
            }
        }
    }
    void add_acc(int idx, int d_col, int rows, int kind, int feat0, int dst, int dst_sbo, int cidx, int eidx, int e_ld) {
        n_acc = std::max(n_acc, idx + 1);
        TsEpi& a = ts.acc[idx];
        a.d_col = (short)d_col; a.rows = (short)rows; a.kind = (short)kind; a.feat0 = (short)feat0;
        a.dst = dst; a.dst_sbo = dst_sbo; a.dst_lo = (nt / 8) * dst_sbo; a.cidx = cidx; a.eidx = eidx; a.e_ld = e_ld;
    }
};

// upload + tensor map (cuTensorMapEncodeTiled through the runtime: no link against libcuda)
int ts_upload(tsp_engine* e, const std::vector<__half>& img, const std::vector<float>& ctab, DevBuf& d_img, DevBuf& d_ctab,
              CUtensorMap& map, int img_rows) {
    int rc;
    if ((rc = ensure(e, d_img, img.size() * sizeof(__half)))) return rc;
    if ((rc = ensure(e, d_ctab, ctab.size() * sizeof(float)))) return rc;
    CK(e, cudaMemcpyAsync(d_img.p, img.data(), img.size() * sizeof(__half), cudaMemcpyHostToDevice, e->stream));
    CK(e, cudaMemcpyAsync(d_ctab.p, ctab.data(), ctab.size() * sizeof(float), cudaMemcpyHostToDevice, e->stream));
    CK(e, cudaStreamSynchronize(e->stream));
    TspEncodeTiled encode = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qres) != cudaSuccess || !encode)
        return fail(e, TSP_ECUDA, "cuTensorMapEncodeTiled is not available from this driver (TMA weight streaming needs it)");
    const cuuint64_t gdim[2] = {64, (cuuint64_t)img_rows};
    const cuuint64_t gstr[1] = {128};
    const cuuint32_t estr[2] = {1, 1};
    const cuuint32_t bx[2] = {64, 128};  // one box == one 16 KB tile
    const CUresult cr = encode(&map, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, d_img.p, gdim, gstr, bx, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) return fail(e, TSP_ECUDA, "cuTensorMapEncodeTiled failed (%d)", (int)cr);
    if ((rc = ensure(e, e->tc_status, sizeof(int)))) return rc;
    CK(e, cudaMemset(e->tc_status.p, 0, sizeof(int)));
    return TSP_OK;
}

int build_tcs(tsp_engine* e) {
    e->has_tcs = false;
    const tsp_config& c = e->cfg;
    if (c.variant != TSP_VARIANT_MUZERO || e->nc != 5 || c.mask_absorbing_states) return TSP_OK;
    for (int m : {TSP_MLP_DYNAMICS, TSP_MLP_REWARD, TSP_MLP_POLICY, TSP_MLP_VALUE})
        if (e->mlp[m].size() != 2) return TSP_OK;
    const HostLayer &D1 = e->mlp[TSP_MLP_DYNAMICS][0], &D2 = e->mlp[TSP_MLP_DYNAMICS][1];
    const HostLayer &R1 = e->mlp[TSP_MLP_REWARD][0], &R2 = e->mlp[TSP_MLP_REWARD][1];
    const HostLayer &P1 = e->mlp[TSP_MLP_POLICY][0], &P2 = e->mlp[TSP_MLP_POLICY][1];
    const HostLayer &V1 = e->mlp[TSP_MLP_VALUE][0], &V2 = e->mlp[TSP_MLP_VALUE][1];
    const int H = c.encoding_size, A = c.num_actions;
    const int DH = D1.out_dim, RH = R1.out_dim, PH = P1.out_dim, VH = V1.out_dim;
    const int NR = R2.out_dim, NV = V2.out_dim;
    if (H % 64 || H <= 64 || H > 256 || DH % 64 || DH > 128 || DH < 64) return TSP_OK;
    if (RH % 16 || PH % 16 || VH % 16 || RH > 32 || PH > 32 || VH > 32 || RH < 16 || PH < 16 || VH < 16) return TSP_OK;
    if (NR + NV + A > 64 || D1.in_dim != H + A || D2.out_dim != H || D2.in_dim != DH) return TSP_OK;
    {   // fp16-split range, decided at weight load: every input of the search MLP is a min-max normalised state in [0, 1]
        // plus a one-hot action, so worst-case magnitudes of all activations follow from the weights alone. If none can reach
        // fp16's range the kernel's overflow flag can never fire for this network; otherwise the mma.sync kernels search it.
        auto bound = [](const HostLayer& L, double in_max, int dense, bool elu) {
            double worst = 0.0;
            for (int r = 0; r < L.out_dim; ++r) {
                double acc = std::fabs((double)L.b[(size_t)r]), onehot = 0.0;
                for (int k = 0; k < L.in_dim; ++k) {
                    const double w = std::fabs((double)L.W[(size_t)r * L.in_dim + k]);
                    if (k < dense) acc += w * in_max;
                    else onehot = std::max(onehot, w);  // one of the one-hot inputs is 1
                }
                worst = std::max(worst, acc + onehot);
            }
            return elu ? std::max(worst, 1.0) : worst;  // ELU: (-1, x]
        };
        const double d1 = bound(D1, 1.0, H, true), xs = bound(D2, d1, DH, false);
        const double r1 = bound(R1, xs, H, true), p1 = bound(P1, 1.0, H, true), v1 = bound(V1, 1.0, H, true);
        const double worst = std::max(std::max(d1, xs), std::max(r1, std::max(p1, v1)));
        if (!(worst < 0.5 * TS_FP16_MAX)) return TSP_OK;
    }
    TsSched& ts = e->tss;
    memset(&ts, 0, sizeof(ts));
    // ---- B-operand arena: the gathered state X0 (K-major, LBO 144) shares its space with the normalised state XN (MN-major;
    // X0 is dead once dynamics layer 1 has completed), then the dynamics hidden layer XD and the heads' hidden layers XH
    // (policy at inputs [0, 32), value at [32, 64), reward at [64, 96)), all MN-major. (Measured: letting XD and XH take
    // turns in the same region as well -- folded reward tile issued first, its epilogue deferred -- frees 56 KB for an
    // 8-slot ring, but the ring is not what limits the stream (shared-memory bandwidth is) and the longer critical path
    // cost 8 %.) ----
    auto mn_sbo = [](int K) { return (K / 8) * 128; };
    const int x0_sbo = (H / 8) * TS_X0_LBO, xn_sbo = mn_sbo(H);
    const int HK = 96;
    int o = 0;
    const int x0 = o; o += 16 * std::max(x0_sbo, xn_sbo);  // X1 image (8 groups of 8 trees) + X2 image
    const int xd = o; o += 16 * mn_sbo(DH);
    const int xh = o; o += 16 * mn_sbo(HK);
    ts.arena_bytes = (o + 127) & ~127;
    ts.x0_off = x0; ts.x0_sbo = x0_sbo; ts.x0_lo = 8 * x0_sbo; ts.H = H;
    // ---- out strip ----
    ts.off_vl = 0; ts.off_rl = NV; ts.off_pl = NV + NR;
    int pitch = NV + NR + A;
    while (pitch % 32 != 8 && pitch % 32 != 24) ++pitch;
    ts.out_pitch = pitch;
    // ---- folded reward layer 1 (see build_tc): (Wr1 W2) d + (Wr1 b2 + br1), product in float64, rounded once ----
    std::vector<float> FW((size_t)RH * DH), Fb((size_t)RH);
    for (int r = 0; r < RH; ++r) {
        for (int k = 0; k < DH; ++k) {
            double acc = 0.0;
            for (int h = 0; h < H; ++h) acc += (double)R1.W[(size_t)r * H + h] * (double)D2.W[(size_t)h * DH + k];
            FW[(size_t)r * DH + k] = (float)acc;
        }
        double acc = R1.b[(size_t)r];
        for (int h = 0; h < H; ++h) acc += (double)R1.W[(size_t)r * H + h] * (double)D2.b[(size_t)h];
        Fb[(size_t)r] = (float)acc;
    }
    // ---- constant table ----
    std::vector<float> ctab;
    auto put = [&](const std::vector<float>& v, int n) {
        const int at = (int)ctab.size();
        for (int i = 0; i < n; ++i) ctab.push_back(v[(size_t)i]);
        while (ctab.size() % 32) ctab.push_back(0.0f);
        return at;
    };
    const int t_b1 = put(D1.b, DH);
    std::vector<float> e1((size_t)A * 128, 0.0f);
    for (int a = 0; a < A; ++a)
        for (int r = 0; r < DH; ++r) e1[(size_t)a * 128 + r] = D1.W[(size_t)r * D1.in_dim + H + a];
    const int t_e1 = put(e1, A * 128);
    const int t_b2 = put(D2.b, H);
    const int t_bf = put(Fb, RH);
    std::vector<float> bpv(64, 0.0f);
    for (int r = 0; r < PH; ++r) bpv[(size_t)r] = P1.b[(size_t)r];
    for (int r = 0; r < VH; ++r) bpv[(size_t)(32 + r)] = V1.b[(size_t)r];
    const int t_bpv = put(bpv, 64);
    std::vector<float> bout(64, 0.0f);  // rows of the output tile: value | reward | policy == columns of the out strip
    for (int r = 0; r < NV; ++r) bout[(size_t)(ts.off_vl + r)] = V2.b[(size_t)r];
    for (int r = 0; r < NR; ++r) bout[(size_t)(ts.off_rl + r)] = R2.b[(size_t)r];
    for (int r = 0; r < A; ++r) bout[(size_t)(ts.off_pl + r)] = P2.b[(size_t)r];
    const int t_bo = put(bout, 64);
    for (int i = 0; i < 128; ++i) ctab.push_back(0.0f);  // lanes past `rows` still form an address
    if ((int)ctab.size() > TS_MAX_CTAB) return TSP_OK;
    ts.ctab_floats = (int)ctab.size();
    TsBuilder tb{ts, TS_N};
    ts.nt = TS_N; ts.n_stages = TS_STAGES;
    auto box = [](int rows) { return TsBuilder::box(rows); };
    const int n_xs = (H + 127) / 128;
    const int A_D1 = 0, A_XS = 1, A_FOLD = 1 + n_xs, A_PV = 2 + n_xs, A_OUT = 3 + n_xs;
    if (A_OUT >= TS_MAX_ACC) return TSP_OK;
    // stage 0: dynamics layer 1 (the one-hot action input is a gathered row in the epilogue, models.py:236-243)
    ts.stage_tile[0] = 0; ts.stage_acc[0] = A_D1;
    tb.add_layer(A_D1, DH, box(DH), H, [&](int m, int k) { return D1.W[(size_t)m * D1.in_dim + k]; }, x0, x0_sbo, false, 0);
    tb.add_acc(A_D1, 0, DH, 0, 0, xd, mn_sbo(DH), t_b1, t_e1, 128);
    // stage 1: dynamics layer 2 -> the un-normalised next state in tiles of 128 features (first: they are on the critical
    // path), then the folded reward layer 1
    ts.stage_tile[1] = tb.n_tiles; ts.stage_acc[1] = A_XS;
    for (int t = 0; t < n_xs; ++t) {
        const int rows = std::min(128, H - 128 * t);
        tb.add_layer(A_XS + t, rows, box(rows), DH, [&](int m, int k) { return D2.W[(size_t)(128 * t + m) * DH + k]; }, xd, mn_sbo(DH),
                     true, 128 * t);
        tb.add_acc(A_XS + t, 128 * t, rows, 1, 128 * t, x0, xn_sbo, t_b2 + 128 * t, -1, 0);
    }
    tb.add_layer(A_FOLD, RH, 32, DH, [&](int m, int k) { return FW[(size_t)m * DH + k]; }, xd, mn_sbo(DH), true, 256);
    tb.add_acc(A_FOLD, 256, RH, 0, 64, xh, mn_sbo(HK), t_bf, -1, 0);
    // stage 2: policy layer 1 (rows 0..31) | value layer 1 (rows 32..63) on the normalised state
    ts.stage_tile[2] = tb.n_tiles; ts.stage_acc[2] = A_PV;
    tb.add_layer(A_PV, 64, 64, H, [&](int m, int k) {
        if (m < 32) return m < PH ? P1.W[(size_t)m * H + k] : 0.0f;
        return m - 32 < VH ? V1.W[(size_t)(m - 32) * H + k] : 0.0f;
    }, x0, xn_sbo, true, 384);
    tb.add_acc(A_PV, 384, 64, 0, 0, xh, mn_sbo(HK), t_bpv, -1, 0);
    // stage 3: the three output layers as one block-diagonal layer over the 96 hidden inputs; rows == out-strip columns
    ts.stage_tile[3] = tb.n_tiles; ts.stage_acc[3] = A_OUT;
    tb.add_layer(A_OUT, NV + NR + A, box(NV + NR + A), HK, [&](int m, int k) {
        if (m >= ts.off_pl) return (k < PH) ? P2.W[(size_t)(m - ts.off_pl) * PH + k] : 0.0f;
        if (m >= ts.off_rl) return (k >= 64 && k - 64 < RH) ? R2.W[(size_t)(m - ts.off_rl) * RH + (k - 64)] : 0.0f;
        return (k >= 32 && k - 32 < VH) ? V2.W[(size_t)m * VH + (k - 32)] : 0.0f;
    }, xh, mn_sbo(HK), true, 0);
    tb.add_acc(A_OUT, 0, NV + NR + A, 2, 0, 0, 0, t_bo, -1, 0);
    ts.stage_tile[4] = tb.n_tiles; ts.stage_acc[4] = tb.n_acc;
    if (tb.overflow || tb.n_acc > TS_MAX_ACC) return TSP_OK;
    ts.n_tiles = tb.n_tiles; ts.n_acc = tb.n_acc;
    ts.img_rows = (int)(tb.img.size() / 64);
    int rc;
    if ((rc = ts_upload(e, tb.img, ctab, e->ts_img, e->ts_ctab, e->ts_map, ts.img_rows))) return rc;
    e->has_tcs = true;
    return TSP_OK;
}

// Streamed tcgen05 root network (root_tcs_kernel, tsp_stream.cuh): initial_inference (models.py:259-281) = the representation
// MLP (any depth; widths <= 256, multiples of 16) -> min-max scale -> policy | value heads (one hidden layer <= 32 each).
// TR_N = 32 observation rows per CTA are the MMAs' columns. Stages: one per representation layer (the last one has no
// activation: its tiles are the hidden state, normalised in the epilogue), policy 1 | value 1, the two output layers
// as one block-diagonal layer. Anything else keeps net_kernel (e->has_rtcs == false).
int build_root_tcs(tsp_engine* e) {
    e->has_rtcs = false;
    const tsp_config& c = e->cfg;
    if (e->nc != 5) return TSP_OK;
    const auto& R = e->mlp[TSP_MLP_REPRESENTATION];
    if (R.empty() || e->mlp[TSP_MLP_POLICY].size() != 2 || e->mlp[TSP_MLP_VALUE].size() != 2) return TSP_OK;
    const HostLayer &P1 = e->mlp[TSP_MLP_POLICY][0], &P2 = e->mlp[TSP_MLP_POLICY][1];
    const HostLayer &V1 = e->mlp[TSP_MLP_VALUE][0], &V2 = e->mlp[TSP_MLP_VALUE][1];
    const int H = c.encoding_size, A = c.num_actions, L = (int)R.size();
    const int PH = P1.out_dim, VH = V1.out_dim, NV = V2.out_dim;
    if (L + 2 > TS_MAX_STAGES || H % 16 || H > 256 || R.back().out_dim != H || R.front().in_dim != c.obs_dim) return TSP_OK;
    if (PH % 16 || VH % 16 || PH > 32 || VH > 32 || PH < 16 || VH < 16 || NV + A > 64) return TSP_OK;
    int wmax = 64, n_acc_need = 2;
    for (int l = 0; l < L; ++l) {
        if (R[l].out_dim % 16 || R[l].out_dim > 256) return TSP_OK;
        wmax = std::max(wmax, R[l].out_dim);
        n_acc_need += (R[l].out_dim + 127) / 128;
    }
    if (n_acc_need > TS_MAX_ACC) return TSP_OK;
    TsSched& ts = e->rts;
    memset(&ts, 0, sizeof(ts));
    TsBuilder tb{ts, TR_N};
    ts.nt = TR_N; ts.n_stages = L + 2; ts.in_dim = c.obs_dim; ts.H = H;
    auto mn_sbo = [](int K) { return TsBuilder::mn_sbo(K); };
    // ---- B operands: X0 = the observation rows (K-major, inputs padded to 16), then two MN-major buffers used alternately by
    // the layers; the second one takes X0's place (X0 is dead once the first layer has completed) ----
    const int K0 = (c.obs_dim + 15) & ~15;
    const int x0_sbo = (K0 / 8) * TS_X0_LBO;
    const int buf_bytes = 2 * (TR_N / 8) * mn_sbo(wmax);
    const int x0_bytes = 2 * (TR_N / 8) * x0_sbo;
    const int bufA = (std::max(x0_bytes, buf_bytes) + 127) & ~127, bufB = 0;
    ts.arena_bytes = (bufA + buf_bytes + 127) & ~127;
    ts.x0_off = 0; ts.x0_sbo = x0_sbo; ts.x0_lo = (TR_N / 8) * x0_sbo;
    // ---- out strip: value logits, then policy logits ----
    ts.off_vl = 0; ts.off_rl = 0; ts.off_pl = NV;
    int pitch = NV + A;
    while (pitch % 32 != 8 && pitch % 32 != 24) ++pitch;
    ts.out_pitch = pitch;
    // ---- constant table ----
    std::vector<float> ctab;
    auto put = [&](const std::vector<float>& v, int n) {
        const int at = (int)ctab.size();
        for (int i = 0; i < n; ++i) ctab.push_back(v[(size_t)i]);
        while (ctab.size() % 32) ctab.push_back(0.0f);
        return at;
    };
    std::vector<int> t_b((size_t)L);
    for (int l = 0; l < L; ++l) t_b[(size_t)l] = put(R[l].b, R[l].out_dim);
    std::vector<float> bpv(64, 0.0f);
    for (int r = 0; r < PH; ++r) bpv[(size_t)r] = P1.b[(size_t)r];
    for (int r = 0; r < VH; ++r) bpv[(size_t)(32 + r)] = V1.b[(size_t)r];
    const int t_bpv = put(bpv, 64);
    std::vector<float> bout(64, 0.0f);
    for (int r = 0; r < NV; ++r) bout[(size_t)(ts.off_vl + r)] = V2.b[(size_t)r];
    for (int r = 0; r < A; ++r) bout[(size_t)(ts.off_pl + r)] = P2.b[(size_t)r];
    const int t_bo = put(bout, 64);
    for (int i = 0; i < 128; ++i) ctab.push_back(0.0f);
    if ((int)ctab.size() > TS_MAX_CTAB) return TSP_OK;
    ts.ctab_floats = (int)ctab.size();
    // ---- stages ----
    int acc = 0;
    int in_off = 0, in_sbo = x0_sbo;
    bool in_mn = false;
    for (int l = 0; l < L; ++l) {
        const HostLayer& Ly = R[l];
        const int K = l == 0 ? K0 : Ly.in_dim;  // (inputs past obs_dim: zero weights against the zero padding of X0)
        const int out_off = (l % 2 == 0) ? bufA : bufB, out_sbo = mn_sbo(Ly.out_dim);
        ts.stage_tile[l] = tb.n_tiles; ts.stage_acc[l] = acc;
        for (int t = 0; t < (Ly.out_dim + 127) / 128; ++t) {
            const int rows = std::min(128, Ly.out_dim - 128 * t);
            tb.add_layer(acc, rows, TsBuilder::box(rows), K,
                         [&](int m, int k) { return k < Ly.in_dim ? Ly.W[(size_t)(128 * t + m) * Ly.in_dim + k] : 0.0f; }, in_off, in_sbo,
                         in_mn, 2 * TR_N * acc);
            // hidden layers: ELU -> operand of the next layer; the last layer (Identity, models.py:1150-1162): min-max scaled
            tb.add_acc(acc, 2 * TR_N * acc, rows, l + 1 < L ? 0 : 1, 128 * t, out_off, out_sbo, t_b[(size_t)l] + 128 * t, -1, 0);
            ++acc;
        }
        in_off = out_off; in_sbo = out_sbo; in_mn = true;
    }
    // policy layer 1 (rows 0..31) | value layer 1 (rows 32..63) on the normalised state
    const int hh_off = (L % 2 == 0) ? bufA : bufB, hh_sbo = mn_sbo(64);
    ts.stage_tile[L] = tb.n_tiles; ts.stage_acc[L] = acc;
    tb.add_layer(acc, 64, 64, H, [&](int m, int k) {
        if (m < 32) return m < PH ? P1.W[(size_t)m * H + k] : 0.0f;
        return m - 32 < VH ? V1.W[(size_t)(m - 32) * H + k] : 0.0f;
    }, in_off, in_sbo, true, 2 * TR_N * acc);
    tb.add_acc(acc, 2 * TR_N * acc, 64, 0, 0, hh_off, hh_sbo, t_bpv, -1, 0);
    ++acc;
    // the two output layers as one block-diagonal layer over the 64 hidden inputs; rows == out-strip columns
    ts.stage_tile[L + 1] = tb.n_tiles; ts.stage_acc[L + 1] = acc;
    tb.add_layer(acc, NV + A, TsBuilder::box(NV + A), 64, [&](int m, int k) {
        if (m >= ts.off_pl) return (k < PH) ? P2.W[(size_t)(m - ts.off_pl) * PH + k] : 0.0f;
        return (k >= 32 && k - 32 < VH) ? V2.W[(size_t)m * VH + (k - 32)] : 0.0f;
    }, hh_off, hh_sbo, true, 2 * TR_N * acc);
    tb.add_acc(acc, 2 * TR_N * acc, NV + A, 2, 0, 0, 0, t_bo, -1, 0);
    ++acc;
    ts.stage_tile[L + 2] = tb.n_tiles; ts.stage_acc[L + 2] = acc;
    if (tb.overflow || acc > TS_MAX_ACC || 2 * TR_N * acc > TS_TMEM_COLS) return TSP_OK;
    ts.n_tiles = tb.n_tiles; ts.n_acc = acc;
    ts.img_rows = (int)(tb.img.size() / 64);
    if (root_tcs_smem(ts) > 227 * 1024) return TSP_OK;
    int rc;
    if ((rc = ts_upload(e, tb.img, ctab, e->rts_img, e->rts_ctab, e->rts_map, ts.img_rows))) return rc;
    e->has_rtcs = true;
    return TSP_OK;
}

// tcgen05 tables of the chance-node searches (MuZeroStochasticConcat, models.py:465-595): stages [0, 4) evaluate ONE future z
// of a transition expansion (dynamics on [h | onehot(a) | onehot(z)] -> reward head on the un-normalised state (first
// layer folded into the MMA of dynamics layer 2, see build_tc) -> min-max normalise; plus the transition prior head on h),
// stages [4, 6) the prediction of a state expansion. One K = H weight set serves both (lanes: dynamics 1 | prior 1 |
// policy 1 | value 1), one K = 32 set holds the four output layers (lanes: reward | prior | policy | value).
int build_tc_variant(tsp_engine* e) {
    e->has_tc = false;
    const tsp_config& c = e->cfg;
    if (c.variant == TSP_VARIANT_MUZERO || e->nc != 5) return TSP_OK;
    for (int m : {TSP_MLP_DYNAMICS, TSP_MLP_REWARD, TSP_MLP_POLICY, TSP_MLP_VALUE, TSP_MLP_PRIOR})
        if (e->mlp[m].size() != 2) return TSP_OK;
    const HostLayer &D1 = e->mlp[TSP_MLP_DYNAMICS][0], &D2 = e->mlp[TSP_MLP_DYNAMICS][1];
    const HostLayer &R1 = e->mlp[TSP_MLP_REWARD][0], &R2 = e->mlp[TSP_MLP_REWARD][1];
    const HostLayer &P1 = e->mlp[TSP_MLP_POLICY][0], &P2 = e->mlp[TSP_MLP_POLICY][1];
    const HostLayer &V1 = e->mlp[TSP_MLP_VALUE][0], &V2 = e->mlp[TSP_MLP_VALUE][1];
    const HostLayer &Q1 = e->mlp[TSP_MLP_PRIOR][0], &Q2 = e->mlp[TSP_MLP_PRIOR][1];
    const int H = c.encoding_size, A = c.num_actions, nf = c.n_futures;
    const int DH = D1.out_dim, RH = R1.out_dim, PH = P1.out_dim, VH = V1.out_dim, QH = Q1.out_dim;
    const int NR = R2.out_dim, NV = V2.out_dim;
    for (int w : {H, DH, RH, PH, VH, QH})
        if (w % 8 || w < 8) return TSP_OK;
    if (H > 64 || DH > 32 || RH > 32 || PH > 32 || VH > 32 || QH > 32 || NR > 32 || NV > 32 || A > 32 || nf > 32) return TSP_OK;
    if (D1.in_dim != H + A + nf || D2.out_dim != H || H + RH > 128) return TSP_OK;
    const int HK = std::max(std::max(RH, QH), std::max(PH, VH));
    auto bufsz = [](int K) { return 2 * (K / 4) * TC_LBO; };
    auto sbo = [](int K) { return (K / 4) * TC_LBO; };
    std::vector<float> FW((size_t)RH * DH), Fb((size_t)RH);
    for (int r = 0; r < RH; ++r) {
        for (int k = 0; k < DH; ++k) {
            double acc = 0.0;
            for (int h = 0; h < H; ++h) acc += (double)R1.W[(size_t)r * H + h] * (double)D2.W[(size_t)h * DH + k];
            FW[(size_t)r * DH + k] = (float)acc;
        }
        double acc = R1.b[(size_t)r];
        for (int h = 0; h < H; ++h) acc += (double)R1.W[(size_t)r * H + h] * (double)D2.b[(size_t)h];
        Fb[(size_t)r] = (float)acc;
    }
    for (int shape = 0; shape < 2; ++shape) {
        TcSched& ts = e->tcs[shape];
        memset(&ts, 0, sizeof(ts));
        const int G = shape == 0 ? 2 : 4;
        const int parts = shape == 0 ? 2 : 1;  // 8 warps per group: warps q and q + 4 split the trees of lane quarter q
        const int h_q = (H + 31) / 32;         // lane quarters of the next state; the folded reward layer sits behind them
        const int c1 = 0, c2 = c1 + H, c4 = c2 + DH, hi_cols = c4 + HK;
        ts.a_lo = hi_cols;
        ts.n_cols = 2 * hi_cols;
        ts.d_col0 = ts.n_cols;
        // Two futures: BOTH in one pass -- the second future's trees are the MMA columns 16..31 (N = 32) from the second dynamics
        // layer on (the first one differs between futures by one gathered weight row only), the prediction's layers ride in
        // the same three stages: 3 commit round trips per iteration instead of 10.
        const bool merged = nf == 2 && !e->no_merge && ts.n_cols + G * 64 <= TC_TMEM_COLS;
        ts.merged = merged ? 1 : 0;
        ts.d_cols_group = merged ? 64 : 48;
        if (ts.n_cols % 8 || ts.d_col0 + G * ts.d_cols_group > TC_TMEM_COLS) return TSP_OK;
        int o = 0;
        const int xh = o; o += bufsz(H);
        const int xd = o; o += bufsz(DH) * (merged ? 2 : 1);
        const int xq = o; o += bufsz(QH);
        // merged: the un-normalised next states (32 columns, no lo twin) reuse the gathered state and the dynamics hidden
        // layer -- both are dead once the second stage's MMAs have completed
        const int xs = (merged && 2 * bufsz(H) <= bufsz(H) + 2 * bufsz(DH)) ? xh : o;
        if (xs == o) o += bufsz(H) * (merged ? 2 : 1);
        const int xr = o; o += bufsz(RH) * (merged ? 2 : 1);
        const int xp = o; o += bufsz(PH);
        const int xv = o; o += bufsz(VH);
        ts.b_lo = o;
        ts.arena_bytes = (2 * o + 127) & ~127;
        ts.xh_off = xh; ts.xh_sbo = sbo(H);
        ts.xs_off = xs; ts.xs_sbo = sbo(H);
        ts.xn_off = -1; ts.xn_sbo = 0;  // no layer reads the normalised state of a future: it goes to the hidden-state pool only
        ts.H = H;
        ts.off_vl = 0; ts.off_rl = NV; ts.off_pl = NV + NR; ts.off_tp = NV + NR + A;
        ts.off_rl2 = NV + NR + A + nf;  // merged: reward logits of future 1
        int pitch = NV + NR + A + nf + (merged ? NR : 0);
        if (shape == 0) while (pitch % 32 != 16) ++pitch;
        else while (pitch % 32 != 8 && pitch % 32 != 24) ++pitch;
        ts.out_pitch = pitch;
        std::vector<float> ctab;
        auto put = [&](const std::vector<float>& v, int n) {
            const int at = (int)ctab.size();
            for (int i = 0; i < n; ++i) ctab.push_back(v[(size_t)i]);
            while (ctab.size() % 32) ctab.push_back(0.0f);
            return at;
        };
        const int t_b1 = put(D1.b, DH);
        std::vector<float> e1((size_t)(A + nf) * 32, 0.0f);  // one-hot action rows, then one-hot future rows
        for (int a = 0; a < A + nf; ++a)
            for (int r = 0; r < DH; ++r) e1[(size_t)a * 32 + r] = D1.W[(size_t)r * D1.in_dim + H + a];
        const int t_e1 = put(e1, (A + nf) * 32);
        const int t_bq1 = put(Q1.b, QH), t_b2 = put(D2.b, H), t_bf = put(Fb, RH);
        const int t_bp1 = put(P1.b, PH), t_bv1 = put(V1.b, VH);
        const int t_br2 = put(R2.b, NR), t_bq2 = put(Q2.b, nf), t_bp2 = put(P2.b, A), t_bv2 = put(V2.b, NV);
        for (int i = 0; i < 64; ++i) ctab.push_back(0.0f);
        if ((int)ctab.size() > TC_MAX_CTAB) return TSP_OK;
        ts.ctab_floats = (int)ctab.size();
        std::vector<float> img((size_t)ts.n_cols * 128, 0.0f);
        auto place = [&](const HostLayer& L, int col0, int lane0, int K) {
            for (int r = 0; r < L.out_dim; ++r)
                for (int k = 0; k < K; ++k) img[(size_t)(col0 + k) * 128 + lane0 + r] = L.W[(size_t)r * L.in_dim + k];
        };
        place(D1, c1, 0, H); place(Q1, c1, 32, H); place(P1, c1, 64, H); place(V1, c1, 96, H);
        place(D2, c2, 0, DH);
        for (int r = 0; r < RH; ++r)
            for (int k = 0; k < DH; ++k) img[(size_t)(c2 + k) * 128 + 32 * h_q + r] = FW[(size_t)r * DH + k];
        place(R2, c4, 0, RH); place(Q2, c4, 32, QH); place(P2, c4, 64, PH); place(V2, c4, 96, VH);
        for (int col = 0; col < hi_cols; ++col)
            for (int l = 0; l < 128; ++l) {
                const float w = img[(size_t)col * 128 + l];
                uint32_t bits;
                memcpy(&bits, &w, 4);
                bits &= 0xFFFFE000u;
                float hi;
                memcpy(&hi, &bits, 4);
                img[(size_t)(hi_cols + col) * 128 + l] = w - hi;
            }
        auto mma = [&](int i, int a_col, int b_off, int K, int d_col, int n = 16) {
            ts.mma[i].a_col = a_col; ts.mma[i].b_off = b_off; ts.mma[i].ksteps = K / 8; ts.mma[i].sbo = sbo(K); ts.mma[i].d_col = d_col;
            ts.mma[i].n = n;
        };
        auto stage = [&](int s, int mma0, int n_mma, int commit, int row_op, int epi0, int n_epi) {
            ts.stage[s].mma0 = mma0; ts.stage[s].n_mma = n_mma; ts.stage[s].commit = commit; ts.stage[s].row_op = row_op;
            ts.stage[s].epi0 = epi0; ts.stage[s].n_epi = n_epi;
        };
        if (merged) {
            const int DA = 0, DB = 32;
            mma(0, c1, xh, H, DA);           // dynamics 1 | prior 1 | policy 1 | value 1 of the gathered state
            mma(1, c2, xd, DH, DB, 32);      // dynamics 2 | folded reward 1, both futures
            mma(2, c4, xp, PH, DA);          // policy 2 (lanes 64..95)
            mma(3, c4, xv, VH, DA + 16);     // value 2 (lanes 96..127)
            mma(4, c4, xr, RH, DB, 32);      // reward 2 (lanes 0..31), both futures
            mma(5, c4, xq, QH, DA);          // prior 2 (lanes 32..63)
            auto epi = [&](int pass, int q, int d_col, int ncols, int rows, int dst, int dst_sbo, int feat0, int cidx, int eidx, int zidx,
                           int flags) {
                for (int h = 0; h < parts; ++h) {
                    TcEpi& x = ts.epi[pass][q + 4 * h];
                    x.d_col = (short)d_col; x.n0 = (short)(h * (ncols / parts)); x.ncnt = (short)(ncols / parts); x.rows = (short)rows;
                    x.dst = dst; x.dst_sbo = dst_sbo; x.feat0 = feat0; x.cidx = cidx; x.eidx = eidx; x.e_ld = 32; x.zidx = zidx; x.flags = flags;
                }
            };
            epi(0, 0, DA, 16, DH, xd, sbo(DH), 0, t_b1, t_e1, t_e1 + A * 32, 1 | 4);
            epi(0, 1, DA, 16, QH, xq, sbo(QH), 0, t_bq1, -1, -1, 1);
            epi(0, 2, DA, 16, PH, xp, sbo(PH), 0, t_bp1, -1, -1, 1);
            epi(0, 3, DA, 16, VH, xv, sbo(VH), 0, t_bv1, -1, -1, 1);
            for (int q = 0; q < h_q; ++q) epi(1, q, DB, 32, std::min(32, H - 32 * q), xs, sbo(H), 32 * q, t_b2 + 32 * q, -1, -1, 8);
            epi(1, h_q, DB, 32, RH, xr, sbo(RH), 0, t_bf, -1, -1, 1);
            epi(1, 3, DA + 16, 16, NV, ts.off_vl, ts.off_vl, 0, t_bv2, -1, -1, 2);
            epi(2, 2, DA, 16, A, ts.off_pl, ts.off_pl, 0, t_bp2, -1, -1, 2);
            epi(3, 0, DB, 32, NR, ts.off_rl, ts.off_rl2, 0, t_br2, -1, -1, 2);
            epi(3, 1, DA, 16, nf, ts.off_tp, ts.off_tp, 0, t_bq2, -1, -1, 2);
            ts.n_stages = 3;
            stage(0, 0, 1, 1, 0, 0, 1);
            stage(1, 1, 3, 1, 0, 1, 2);
            stage(2, 4, 2, 1, 1, 3, 1);
            for (int s2 = 0; s2 < 3; ++s2) ts.stage[s2].issuer = parts == 2 ? 7 : 3;
        } else {
        const int DA = 0, DB = 16;
        mma(0, c1, xh, H, DA);    // dynamics 1 | prior 1
        mma(1, c2, xd, DH, DB);   // dynamics 2 | folded reward 1
        mma(2, c4, xr, RH, DA);   // reward 2 (lanes 0..31)
        mma(3, c4, xq, QH, DB);   // prior 2 (lanes 32..63)
        mma(4, c1, xh, H, DA);    // policy 1 | value 1 (lanes 64..127)
        mma(5, c4, xp, PH, DA);   // policy 2 (lanes 64..95)
        mma(6, c4, xv, VH, DB);   // value 2 (lanes 96..127)
        auto epi = [&](int pass, int q, int d_col, int rows, int dst, int dst_sbo, int feat0, int cidx, int eidx, int zidx, int flags) {
            for (int h = 0; h < parts; ++h) {
                TcEpi& x = ts.epi[pass][q + 4 * h];
                x.d_col = (short)d_col; x.n0 = (short)(h * (16 / parts)); x.ncnt = (short)(16 / parts); x.rows = (short)rows;
                x.dst = dst; x.dst_sbo = dst_sbo; x.feat0 = feat0; x.cidx = cidx; x.eidx = eidx; x.e_ld = 32; x.zidx = zidx; x.flags = flags;
            }
        };
        epi(0, 0, DA, DH, xd, sbo(DH), 0, t_b1, t_e1, t_e1 + A * 32, 1);
        epi(0, 1, DA, QH, xq, sbo(QH), 0, t_bq1, -1, -1, 1);
        for (int q = 0; q < h_q; ++q) epi(1, q, DB, std::min(32, H - 32 * q), xs, sbo(H), 32 * q, t_b2 + 32 * q, -1, -1, 0);
        epi(1, h_q, DB, RH, xr, sbo(RH), 0, t_bf, -1, -1, 1);
        epi(2, 0, DA, NR, ts.off_rl, 0, 0, t_br2, -1, -1, 2);
        epi(2, 1, DB, nf, ts.off_tp, 0, 0, t_bq2, -1, -1, 2);
        epi(3, 2, DA, PH, xp, sbo(PH), 0, t_bp1, -1, -1, 1);
        epi(3, 3, DA, VH, xv, sbo(VH), 0, t_bv1, -1, -1, 1);
        epi(4, 2, DA, A, ts.off_pl, 0, 0, t_bp2, -1, -1, 2);
        epi(4, 3, DB, NV, ts.off_vl, 0, 0, t_bv2, -1, -1, 2);
        ts.n_stages = 6;
        stage(0, 0, 1, 1, 0, 0, 1);
        stage(1, 1, 1, 1, 0, 1, 1);
        stage(2, 0, 0, 0, 1, 0, 0);
        stage(3, 2, 2, 1, 0, 2, 1);
        stage(4, 4, 1, 1, 0, 3, 1);
        stage(5, 5, 2, 1, 0, 4, 1);
        // the MMAs of a stage are issued by a warp without a share of its epilogue: quarter 3 is idle in the stages of a
        // future, quarter 0 in those of the prediction
        for (int s2 = 0; s2 < 4; ++s2) ts.stage[s2].issuer = parts == 2 ? 7 : 3;
        for (int s2 = 4; s2 < 6; ++s2) ts.stage[s2].issuer = parts == 2 ? 4 : 0;
        }
        int rc;
        if ((rc = ensure(e, e->tc_img[shape], img.size() * sizeof(float)))) return rc;
        if ((rc = ensure(e, e->tc_ctab[shape], ctab.size() * sizeof(float)))) return rc;
        CK(e, cudaMemcpyAsync(e->tc_img[shape].p, img.data(), img.size() * sizeof(float), cudaMemcpyHostToDevice, e->stream));
        CK(e, cudaMemcpyAsync(e->tc_ctab[shape].p, ctab.data(), ctab.size() * sizeof(float), cudaMemcpyHostToDevice, e->stream));
        CK(e, cudaStreamSynchronize(e->stream));
    }
    int rc;
    if ((rc = ensure(e, e->tc_status, sizeof(int)))) return rc;
    CK(e, cudaMemset(e->tc_status.p, 0, sizeof(int)));
    e->has_tc = true;
    return TSP_OK;
}

int build_schedules(tsp_engine* e) {
    const tsp_config& c = e->cfg;
    const int H = c.encoding_size, A = c.num_actions;
    const bool stochastic = c.variant != TSP_VARIANT_MUZERO;
    auto have = [&](int m) { return !e->mlp[m].empty(); };
    auto check_io = [&](int m, int in, int out, const char* name) -> int {
        if (!have(m)) return fail(e, TSP_ESTATE, "weights of the %s network were not loaded", name);
        if (e->mlp[m].front().in_dim != in || e->mlp[m].back().out_dim != out)
            return fail(e, TSP_EINVAL, "%s network has shape %d -> %d, expected %d -> %d", name,
                        e->mlp[m].front().in_dim, e->mlp[m].back().out_dim, in, out);
        return TSP_OK;
    };
    const int V = 2 * c.support_size + 1;
    int rc;
    if ((rc = check_io(TSP_MLP_REPRESENTATION, c.obs_dim, H, "representation"))) return rc;
    if ((rc = check_io(TSP_MLP_DYNAMICS, H + A + (stochastic && !e->separate_heads ? c.n_futures : 0), H, "dynamics"))) return rc;
    if ((rc = check_io(TSP_MLP_REWARD, H, V, "reward"))) return rc;
    if ((rc = check_io(TSP_MLP_POLICY, H, A, "policy"))) return rc;
    if ((rc = check_io(TSP_MLP_VALUE, H, V, "value"))) return rc;
    if (stochastic && (rc = check_io(TSP_MLP_PRIOR, H, c.n_futures, "transition prior"))) return rc;
    if (c.mask_absorbing_states && (rc = check_io(TSP_MLP_TERMINAL, H, 1, "terminal"))) return rc;

    {  // root: models.py:259-281 (reconstruction head skipped: MCTS discards it)
        HSched hs;
        hs.in_dim = c.obs_dim;
        hs.in_buf = hs.new_buf(c.obs_dim, 0);
        const int raw = add_chain(e, hs, TSP_MLP_REPRESENTATION, hs.in_buf, 0);
        hs.hidden = add_normalise(hs, raw, H);
        hs.pl = add_chain(e, hs, TSP_MLP_POLICY, hs.hidden, 0);
        hs.vl = add_chain(e, hs, TSP_MLP_VALUE, hs.hidden, 0);
        if ((rc = lower(e, hs, e->sch_root))) return rc;
        e->has_root = true;
        if ((rc = build_root_tcs(e))) return rc;
    }
    {  // prediction: models.py:209-212
        HSched hs;
        hs.in_dim = H;
        hs.in_buf = hs.new_buf(H, 0);
        hs.pl = add_chain(e, hs, TSP_MLP_POLICY, hs.in_buf, 0);
        hs.vl = add_chain(e, hs, TSP_MLP_VALUE, hs.in_buf, 0);
        if ((rc = lower(e, hs, e->sch_pred))) return rc;
        e->has_pred = true;
    }
    if (!stochastic) {  // recurrent inference: models.py:233-257, 283-287
        const bool have_terminal = have(TSP_MLP_TERMINAL) && e->mlp[TSP_MLP_TERMINAL].back().out_dim == 1 &&
                                   e->mlp[TSP_MLP_TERMINAL].front().in_dim == H;
        auto make = [&](bool with_terminal) {
            HSched hs;
            hs.in_dim = H;
            hs.in_buf = hs.new_buf(H, 0);
            const int raw = add_chain(e, hs, TSP_MLP_DYNAMICS, hs.in_buf, 1);
            hs.rl = add_chain(e, hs, TSP_MLP_REWARD, raw, 0);  // reward / terminal read the UN-normalised state
            if (with_terminal) hs.tl = add_chain(e, hs, TSP_MLP_TERMINAL, raw, 0);
            hs.hidden = add_normalise(hs, raw, H);
            hs.pl = add_chain(e, hs, TSP_MLP_POLICY, hs.hidden, 0);
            hs.vl = add_chain(e, hs, TSP_MLP_VALUE, hs.hidden, 0);
            return hs;
        };
        if ((rc = lower(e, make(have_terminal), e->sch_recur))) return rc;
        e->has_recur = true;
        // the search evaluates the terminal head only under mask_absorbing_states (self_play.py:415)
        if (e->nc == 5 && e->farena_floats > 0 && H % 16 == 0 && (!c.mask_absorbing_states || have_terminal)) {
            const HSched fh = make(c.mask_absorbing_states != 0);
            if ((rc = lower_fast(e, fh, 0, 8))) return rc;
            if (e->has_fast && (rc = lower_fast(e, fh, 1, 4))) return rc;
        }
        if ((rc = build_tc(e))) return rc;
        e->has_tcs = false;
        if (!e->has_tc && (rc = build_tcs(e))) return rc;
    } else {  // one future of MuZeroStochasticConcat.dynamics: models.py:520-595
        HSched hs;
        hs.in_dim = H;
        hs.in_buf = hs.new_buf(H, 0);
        // (separate heads: [h | onehot(a)] into the dynamics MLP of future z, whose weights sit z * f_stride further on)
        const int raw = add_chain(e, hs, TSP_MLP_DYNAMICS, hs.in_buf, e->separate_heads ? 1 : 2);
        hs.rl = add_chain(e, hs, TSP_MLP_REWARD, raw, 0);
        hs.hidden = add_normalise(hs, raw, H);
        hs.tp = add_chain(e, hs, TSP_MLP_PRIOR, hs.in_buf, 0);
        if ((rc = lower(e, hs, e->sch_trans, true))) return rc;
        e->sch_trans.f_stride = e->separate_heads ? e->f_stride : 0;
        e->has_trans = true;
        // the search kernel addresses both schedules with one row stride
        const int rs = std::max(e->sch_trans.row_stride, e->sch_pred.row_stride);
        e->sch_trans.row_stride = rs;
        e->sch_pred.row_stride = rs;
        e->has_fast = false;
        // (the separate-heads network runs on the generic kernel: the fast / tcgen05 tables hold one dynamics MLP)
        if (e->nc == 5 && e->farena_floats > 0 && H % 16 == 0 && !e->separate_heads) {  // the chance-node searches on the fast kernel's structure
            HSched hp;
            hp.in_dim = H;
            hp.in_buf = hp.new_buf(H, 0);
            hp.pl = add_chain(e, hp, TSP_MLP_POLICY, hp.in_buf, 0);
            hp.vl = add_chain(e, hp, TSP_MLP_VALUE, hp.in_buf, 0);
            bool ok = true;
            for (int shape = 0; shape < 2 && ok; ++shape) {
                const int gw = shape == 0 ? 8 : 4;
                if ((rc = lower_fast(e, hs, shape, gw, 0, true))) return rc;
                ok = e->has_fast;
                if (ok && (rc = lower_fast(e, hp, shape, gw, 1, true))) return rc;
                ok = ok && e->has_fast;
                if (ok) {  // one strip geometry for both schedules
                    const int frs = std::max(e->fsched[shape][0].row_stride, e->fsched[shape][1].row_stride);
                    for (int k = 0; k < 2; ++k) {
                        e->fsched[shape][k].row_stride = frs;
                        CK(e, cudaMemcpyAsync(e->fsched_dev[shape][k].p, &e->fsched[shape][k], sizeof(FastSched),
                                              cudaMemcpyHostToDevice, e->stream));
                    }
                    CK(e, cudaStreamSynchronize(e->stream));
                }
            }
            e->has_fast = ok;
        }
        e->has_tc = false;
        if (!e->separate_heads && (rc = build_tc_variant(e))) return rc;
    }
    for (Schedule* s : {&e->sch_root, &e->sch_pred, &e->sch_recur, &e->sch_trans})
        if (s->n_stages && s->off_in != 0) return fail(e, TSP_EINVAL, "internal: schedule input not at offset 0");
    return TSP_OK;
}

// Columns per thread for every op of a schedule evaluated by a CTA with `slots` = threads / TM
// unit slots: the widest chunk (fewest loads per FMA) that still gives every slot a unit.
void set_chunks(Schedule& sc, int slots, int forced) {
    for (int s = 0; s < sc.n_stages; ++s) {
        Stage& st = sc.st[s];
        int best = 4;
        for (int ct : {16, 8, 4}) {
            int units = 0;
            for (int o = 0; o < st.n_ops; ++o) {
                int c = ct;
                while (st.ops[o].Npad % c) c >>= 1;
                units += st.ops[o].Npad / c;
            }
            if (units >= slots || ct == 4) {
                best = ct;
                break;
            }
        }
        if (forced == 4 || forced == 8 || forced == 16) best = forced;
        for (int o = 0; o < st.n_ops; ++o) {
            int c = best;
            while (st.ops[o].Npad % c) c >>= 1;
            st.ops[o].ct = (unsigned char)c;
        }
    }
}

// Tensor-core path: 8-column tiles per warp unit. The smallest group size (most warps busy, shortest
// serial stream per warp) for which all units of a stage fit in one pass over the CTA's warps.
void set_tile_groups(Schedule& sc, int TM, int warps, int forced) {
    for (int s = 0; s < sc.n_stages; ++s) {
        Stage& st = sc.st[s];
        int best = 4;
        for (int ng = 1; ng <= 4; ++ng) {
            int units = 0;
            for (int o = 0; o < st.n_ops; ++o) units += (TM / 16) * (((st.ops[o].Npad >> 3) + ng - 1) / ng);
            if (units <= warps) {
                best = ng;
                break;
            }
        }
        if (forced >= 1 && forced <= 4) best = forced;
        for (int o = 0; o < st.n_ops; ++o) {
            LayerOp& op = st.ops[o];
            op.ng = (unsigned char)best;
            op.groups = (short)(((op.Npad >> 3) + best - 1) / best);
            op.units = (short)((TM / 16) * op.groups);
        }
    }
}

template <typename K>
int set_smem(tsp_engine* e, K kernel, size_t bytes) {
    if (bytes > 227 * 1024) return fail(e, TSP_EINVAL, "kernel needs %zu bytes of shared memory (> 227 KB)", bytes);
    CK(e, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return TSP_OK;
}

size_t net_smem(const Schedule& sc, int TM) { return (size_t)TM * sc.row_stride * 4 + TM * 4 + TM * 2 + 16; }
size_t search_smem(const Schedule& sc, int TM, int S, size_t w_floats = 0) {
    return (size_t)((S + 3) & ~1) * 16 + (size_t)TM * sc.row_stride * 4 + TM * 4 + TM * 2 + 32 + w_floats * 4;
}

template <int NC, int TM>
int launch_net(tsp_engine* e, const NetParams& p_in) {
    NetParams p = p_in;
    set_chunks(p.sched, GW, e->ct_override);
    set_tile_groups(p.sched, TM, TM * GW / 32, e->ng_override);
    const size_t smem = net_smem(p.sched, TM);
    int rc = set_smem(e, net_kernel<NC, TM>, smem);
    if (rc) return rc;
    const int grid = (p.B - p.b0 + TM - 1) / TM;
    net_kernel<NC, TM><<<grid, TM * GW, smem, e->stream>>>(p);
    CK(e, cudaGetLastError());
    e->stats.kernel_launches++;
    return TSP_OK;
}

// initial_inference (+ the root of the search) on tcgen05 with the weights streamed by TMA (root_tcs_kernel)
int launch_root_tcs(tsp_engine* e, const NetParams& np) {
    RootTcsParams p;
    memset(&p, 0, sizeof(p));
    const tsp_config& c = e->cfg;
    p.B = np.B; p.b0 = np.b0; p.A = np.A; p.H = np.H; p.support = np.support; p.in_dim = c.obs_dim;
    p.in = np.in;
    p.ctab = reinterpret_cast<const float*>(e->rts_ctab.p);
    p.status = reinterpret_cast<int*>(e->tc_status.p);
    p.value = np.value; p.policy_logits = np.policy_logits; p.hidden_out = np.hidden_out; p.value_logits = np.value_logits;
    p.tree_init = np.tree_init; p.add_noise = np.add_noise; p.uniform_policy = np.uniform_policy;
    p.blocks_per_tree = np.blocks_per_tree; p.hidden_per_tree = np.hidden_per_tree;
    p.one_minus_frac = np.one_minus_frac; p.frac = np.frac;
    p.blocks = np.blocks; p.hidden_pool = np.hidden_pool; p.hdr = np.hdr; p.rootprior = np.rootprior; p.rootact = np.rootact;
    p.legal = np.legal; p.num_legal = np.num_legal; p.noise = np.noise;
    p.trace_root = np.trace_root; p.root_pred = np.root_pred; p.root_hidden = np.root_hidden;
    if (!p.tree_init && !p.hidden_out) {  // the epilogue writes the hidden state somewhere: a scratch array
        int rc = ensure(e, e->rts_hidden, (size_t)np.B * np.H * sizeof(float));
        if (rc) return rc;
        p.hidden_out = reinterpret_cast<float*>(e->rts_hidden.p);
    }
    const size_t smem = std::max(root_tcs_smem(e->rts), (size_t)120 * 1024);  // one CTA per SM (it owns the SM's tensor memory)
    int rc = set_smem(e, root_tcs_kernel, smem);
    if (rc) return rc;
    const int grid = (np.B - np.b0 + TR_N - 1) / TR_N;
    root_tcs_kernel<<<grid, TR_THREADS, smem, e->stream>>>(p, e->rts, e->rts_map);
    CK(e, cudaGetLastError());
    e->stats.kernel_launches++;
    e->stats.root_kernel = 1;
    return TSP_OK;
}

int launch_net_any(tsp_engine* e, const NetParams& p) {
    if (p.mode == 0 && !p.fed_root && e->has_rtcs && !e->no_rtcs && e->nc == 5) return launch_root_tcs(e, p);
    if (p.mode == 0) e->stats.root_kernel = 0;
    // wide observation strips: fewer rows per CTA keep the strip under the shared-memory limit
    const int TM = (net_smem(p.sched, 32) <= 100 * 1024) ? 32 : 16;
    if (e->nc == 5) return TM == 32 ? launch_net<5, 32>(e, p) : launch_net<5, 16>(e, p);
    return TM == 32 ? launch_net<8, 32>(e, p) : launch_net<8, 16>(e, p);
}

template <int VARIANT, int NC, int TM, bool WSMEM, int XW>
int launch_search_x(tsp_engine* e, const SearchParams& p_in) {
    SearchParams p = p_in;
    set_chunks(p.sched_a, GW * XW, e->ct_override);
    set_chunks(p.sched_b, GW * XW, e->ct_override);
    set_tile_groups(p.sched_a, TM, TM * GW * XW / 32, e->ng_override);
    set_tile_groups(p.sched_b, TM, TM * GW * XW / 32, e->ng_override);
    const size_t smem = search_smem(p.sched_a, TM, p.S_tab, WSMEM ? (size_t)p.w_smem_floats : 0);
    int rc = set_smem(e, search_kernel<VARIANT, NC, TM, WSMEM, XW>, smem);
    if (rc) return rc;
    const int grid = (p.B + TM - 1) / TM;
    search_kernel<VARIANT, NC, TM, WSMEM, XW><<<grid, TM * GW * XW, smem, e->stream>>>(p);
    CK(e, cudaGetLastError());
    e->stats.kernel_launches++;
    e->stats.search_ctas = grid;
    e->stats.search_smem_bytes = (int)smem;
    e->stats.trees_per_cta = TM;
    e->stats.threads_per_cta = TM * GW * XW;
    e->stats.search_kernel = 0;
    return TSP_OK;
}

template <int VARIANT, int NC, int TM, bool WSMEM>
int launch_search_w(tsp_engine* e, const SearchParams& p, int xw) {
    if (NC == 5 && VARIANT == 0) {
        if (xw == 2) return launch_search_x<VARIANT, NC, TM, WSMEM, (NC == 5 && VARIANT == 0 ? 2 : 1)>(e, p);
    }
    return launch_search_x<VARIANT, NC, TM, WSMEM, 1>(e, p);
}

template <int VARIANT, int NC, int TM>
int launch_search(tsp_engine* e, const SearchParams& p, int xw) {
    return p.w_smem_floats > 0 ? launch_search_w<VARIANT, NC, TM, true>(e, p, xw)
                               : launch_search_w<VARIANT, NC, TM, false>(e, p, xw);
}

template <int VARIANT, int NC>
int launch_search_tm(tsp_engine* e, const SearchParams& p, int TM, int xw) {
    if (NC == 5 && TM == 16) return launch_search<VARIANT, NC, (NC == 5 ? 16 : 32)>(e, p, xw);
    return launch_search<VARIANT, NC, 32>(e, p, xw);
}

size_t fast_smem(const FastSched& fs, int G, int S, bool w_lo = false, int n_sched = 1, bool w_global = false) {
    const size_t pb_n = (size_t)((S + 3) & ~1);
    return pb_n * 16 + ((pb_n + 3) & ~(size_t)3) * 4 + n_sched * sizeof(FastSched) +
           (w_global ? 0 : (size_t)fs.w_floats * (w_lo ? 8 : 4)) +
           (size_t)G * ((w_global ? GT / 2 : GT) * fs.row_stride + 2 * GT + GT * PATH_LEVELS) * 4;
}

// Groups of 16 trees per CTA. One wave (or kernels that may share an SM): G per CTA. Several waves of a one-CTA-per-SM
// kernel (`balance`): the groups are spread evenly over full waves -- sm_count x rounds CTAs with floor / ceil of the
// average -- so that no SM idles through a short last wave while others run G groups.
int groups_per_cta(tsp_engine* e, FastParams& p, int G, bool balance) {
    const int NG = (p.B + GT - 1) / GT;
    int grid = (NG + G - 1) / G;
    p.grp_base = G;
    p.grp_extra = 0;
    if (balance && grid > e->sm_count && !e->no_balance) {
        const int rounds = (grid + e->sm_count - 1) / e->sm_count;
        grid = e->sm_count * rounds;
        p.grp_base = NG / grid;
        p.grp_extra = NG % grid;
    }
    return grid;
}

template <int G, int LPT, bool WG = false>
int launch_fast(tsp_engine* e, const FastParams& p_in) {
    FastParams p = p_in;
    const int grid = groups_per_cta(e, p, G, false);
    const size_t smem = fast_smem(e->fsched[LPT == 16 ? 0 : 1][0], G, p.S_tab, p.w_lo != 0, 1, WG);
    int rc = set_smem(e, search_fast_kernel<G, LPT, WG>, smem);
    if (rc) return rc;
    constexpr int GTHREADS = GT * LPT;
    search_fast_kernel<G, LPT, WG><<<grid, G * GTHREADS, smem, e->stream>>>(p);
    CK(e, cudaGetLastError());
    e->stats.kernel_launches++;
    e->stats.search_ctas = grid;
    e->stats.search_smem_bytes = (int)smem;
    e->stats.trees_per_cta = G * GT;
    e->stats.threads_per_cta = G * GTHREADS;
    e->stats.search_kernel = 1;
    return TSP_OK;
}

size_t tc_smem(const TcSched& ts, int G, int S) {
    const size_t pb_n = (size_t)((S + 3) & ~1);
    return pb_n * 16 + ((pb_n + 3) & ~(size_t)3) * 4 + sizeof(FastSched) + (size_t)((ts.ctab_floats + 3) & ~3) * 4 + 64 + 256 +
           (size_t)G * (ts.arena_bytes + (GT * ts.out_pitch + GT + GT / 4 + GT * PATH_LEVELS) * 4);
}

template <int G, int LPT>
int launch_tc(tsp_engine* e, const FastParams& p_in, const TcSched& ts) {
    FastParams p = p_in;
    const int grid = groups_per_cta(e, p, G, true);
    // at least half of an SM's shared memory: one CTA per SM, which owns all of the SM's tensor memory
    const size_t smem = std::max(tc_smem(ts, G, p.S_tab), (size_t)120 * 1024);
    int rc = set_smem(e, search_tc_kernel<G, LPT>, smem);
    if (rc) return rc;
    constexpr int GTHREADS = GT * LPT;
    search_tc_kernel<G, LPT><<<grid, G * GTHREADS, smem, e->stream>>>(p, ts);
    CK(e, cudaGetLastError());
    e->stats.kernel_launches++;
    e->stats.search_ctas = grid;
    e->stats.search_smem_bytes = (int)smem;
    e->stats.trees_per_cta = G * GT;
    e->stats.threads_per_cta = G * GTHREADS;
    e->stats.search_kernel = 2;
    return TSP_OK;
}

size_t tcs_smem(const TsSched& ts, int S) {
    const size_t pb_n = (size_t)((S + 3) & ~1);
    return pb_n * 16 + ((pb_n + 3) & ~(size_t)3) * 4 + (size_t)((ts.ctab_floats + 3) & ~3) * 4 + sizeof(TsCtl) + 256 +
           (size_t)TS_NS * TS_SLOT + (size_t)ts.arena_bytes + (size_t)(TS_N * ts.out_pitch + TS_N + TS_N / 4 + TS_N * PATH_LEVELS) * 4;
}

// The search with the MLP on tcgen05 and the weights streamed by TMA (tsp_stream.cuh): 64 trees per CTA, one CTA per SM.
int launch_tcs(tsp_engine* e, const FastParams& p_in) {
    FastParams p = p_in;
    p.grp_base = TS_N / GT;  // the CTA's 64 trees are one lock-stepped group of four 16-tree slots
    p.grp_extra = 0;
    const TsSched& ts = e->tss;
    const size_t smem = std::max(tcs_smem(ts, p.S_tab), (size_t)120 * 1024);
    int rc = set_smem(e, search_tcs_kernel, smem);
    if (rc) return rc;
    const int grid = (p.B + TS_N - 1) / TS_N;
    constexpr int THREADS = TS_N * 8 + TS_SERVICE_THREADS;
    search_tcs_kernel<<<grid, THREADS, smem, e->stream>>>(p, ts, e->ts_map);
    CK(e, cudaGetLastError());
    e->stats.kernel_launches++;
    e->stats.search_ctas = grid;
    e->stats.search_smem_bytes = (int)smem;
    e->stats.trees_per_cta = TS_N;
    e->stats.threads_per_cta = THREADS;
    e->stats.search_kernel = 3;
    return TSP_OK;
}

size_t tc_variant_smem(const TcSched& ts, int G, int S) {
    const size_t pb_n = (size_t)((S + 3) & ~1);
    return pb_n * 16 + ((pb_n + 3) & ~(size_t)3) * 4 + (size_t)((ts.ctab_floats + 3) & ~3) * 4 + 64 + 256 +
           (size_t)G * (ts.arena_bytes + (GT * ts.out_pitch + GT + GT / 4) * 4);
}

template <int G, int LPT, int VARIANT, bool MERGED>
int launch_tc_variant(tsp_engine* e, const FastParams& p_in, const TcSched& ts) {
    FastParams p = p_in;
    const int grid = groups_per_cta(e, p, G, true);
    const size_t smem = std::max(tc_variant_smem(ts, G, p.S_tab), (size_t)120 * 1024);  // one CTA per SM (it owns the SM's tensor memory)
    int rc = set_smem(e, search_tc_variant_kernel<G, LPT, VARIANT, MERGED>, smem);
    if (rc) return rc;
    constexpr int GTHREADS = GT * LPT;
    search_tc_variant_kernel<G, LPT, VARIANT, MERGED><<<grid, G * GTHREADS, smem, e->stream>>>(p, ts);
    CK(e, cudaGetLastError());
    e->stats.kernel_launches++;
    e->stats.search_ctas = grid;
    e->stats.search_smem_bytes = (int)smem;
    e->stats.trees_per_cta = G * GT;
    e->stats.threads_per_cta = G * GTHREADS;
    e->stats.search_kernel = 2;
    return TSP_OK;
}

template <int G, int LPT, int VARIANT>
int launch_fast_variant(tsp_engine* e, const FastParams& p_in) {
    FastParams p = p_in;
    const int grid = groups_per_cta(e, p, G, false);
    const size_t smem = fast_smem(e->fsched[LPT == 16 ? 0 : 1][0], G, p.S_tab, p.w_lo != 0, 2);
    int rc = set_smem(e, search_fast_variant_kernel<G, LPT, VARIANT>, smem);
    if (rc) return rc;
    constexpr int GTHREADS = GT * LPT;
    search_fast_variant_kernel<G, LPT, VARIANT><<<grid, G * GTHREADS, smem, e->stream>>>(p);
    CK(e, cudaGetLastError());
    e->stats.kernel_launches++;
    e->stats.search_ctas = grid;
    e->stats.search_smem_bytes = (int)smem;
    e->stats.trees_per_cta = G * GT;
    e->stats.threads_per_cta = G * GTHREADS;
    e->stats.search_kernel = 1;
    return TSP_OK;
}

// Search on the fast kernels: groups of 16 trees, G groups per CTA sharing one weight copy.
int launch_fast_any(tsp_engine* e, const SearchParams& sp) {
    FastParams p;
    memset(&p, 0, sizeof(p));
    p.B = sp.B; p.S = sp.S; p.A = sp.A; p.H = sp.H; p.support = sp.support;
    p.mask_absorbing = sp.mask_absorbing; p.uniform_policy = sp.uniform_policy;
    p.blocks_per_tree = sp.blocks_per_tree; p.hidden_per_tree = sp.hidden_per_tree;
    p.max_rec = sp.max_rec; p.T = sp.T; p.pb_stride = sp.pb_stride;
    p.discount = sp.discount; p.pb_table = sp.pb_table;
    p.two_player = sp.two_player;
    p.grp_base = 0; p.grp_extra = 0;  // set per launch (groups_per_cta)
    p.S_tab = sp.S_tab; p.resume = sp.resume; p.add_noise = sp.add_noise; p.one_minus_frac = sp.one_minus_frac; p.frac = sp.frac; p.noise = sp.noise;
    p.discount_f = sp.two_player ? -(float)sp.discount : (float)sp.discount;
    p.blocks = sp.blocks; p.hidden_pool = sp.hidden_pool; p.hdr = sp.hdr; p.rootprior = sp.rootprior; p.rootact = sp.rootact;
    p.tie = sp.tie; p.fed = sp.fed; p.trace = sp.trace;
    p.wfrag = reinterpret_cast<const float*>(e->farena.p);
    p.visit_counts = sp.visit_counts; p.root_value = sp.root_value; p.child_value = sp.child_value;
    p.child_prior = sp.child_prior; p.child_reward = sp.child_reward;
    p.max_depth = sp.max_depth; p.num_records = sp.num_records; p.total_depth = sp.total_depth;
    p.phase_cycles = sp.phase_cycles;
    // Group shape. Few roots (<= 2 groups of 16 trees per SM): the search is latency-bound -- 8 warps per group,
    // two groups per CTA sharing one weight copy. Many roots: 4 warps per group (four trees per warp: fewer issued
    // instructions per simulation) and four groups per CTA = 64 trees per SM at the same 16 warps.
    const int groups = (p.B + GT - 1) / GT;
    int lpt = groups <= 2 * e->sm_count ? 16 : 8;
    if (e->fast_lpt == 8 || e->fast_lpt == 16) lpt = e->fast_lpt;
    const int which = lpt == 16 ? 0 : 1;
    p.sched = reinterpret_cast<const FastSched*>(e->fsched_dev[which][0].p);
    p.sched_b = reinterpret_cast<const FastSched*>(e->fsched_dev[which][1].p);
    p.nf = sp.nf; p.K = sp.K; p.max_iters = sp.max_iters; p.chance = sp.chance;
    const int variant = e->cfg.variant, n_sched = variant == 0 ? 1 : 2;
    if (variant == 0 && e->has_tcs && !e->no_tcs && !p.fed && !p.mask_absorbing && tcs_smem(e->tss, p.S_tab) <= 227 * 1024) {
        // the MLP on tcgen05 with the weights streamed by TMA (tsp_stream.cuh)
        p.tc_ctab = reinterpret_cast<const float*>(e->ts_ctab.p);
        p.tc_status = reinterpret_cast<int*>(e->tc_status.p);
        return launch_tcs(e, p);
    }
    if (variant == 0 && e->has_tc && !e->no_tc && !p.fed && !p.mask_absorbing) {
        // the MLP on tcgen05 with the weights resident in tensor memory (tsp_tc.cuh)
        const TcSched& ts = e->tcs[which];
        p.tc_img = reinterpret_cast<const float*>(e->tc_img[which].p);
        p.tc_ctab = reinterpret_cast<const float*>(e->tc_ctab[which].p);
        p.tc_status = reinterpret_cast<int*>(e->tc_status.p);
        if (lpt == 16 && e->fast_groups == 1) return launch_tc<1, 16>(e, p, ts);  // (developer: one group alone on an SM)
        if (lpt == 16 && tc_smem(ts, 2, p.S_tab) <= 227 * 1024) return launch_tc<2, 16>(e, p, ts);
        if (lpt == 8 && tc_smem(ts, 4, p.S_tab) <= 227 * 1024) return launch_tc<4, 8>(e, p, ts);
        if (lpt == 8 && tc_smem(ts, 2, p.S_tab) <= 227 * 1024) return launch_tc<2, 8>(e, p, ts);
        if (!e->has_fast) return fail(e, TSP_EINVAL, "internal: tcgen05 search kernel does not fit shared memory");
    }
    if (variant != 0 && e->has_tc && !e->no_tc && !p.fed) {
        const TcSched& ts = e->tcs[which];
        p.tc_img = reinterpret_cast<const float*>(e->tc_img[which].p);
        p.tc_ctab = reinterpret_cast<const float*>(e->tc_ctab[which].p);
        p.tc_status = reinterpret_cast<int*>(e->tc_status.p);
#define TSP_LAUNCH_TV(GG, LL)                                                                                                   \
    (ts.merged ? (variant == 1 ? launch_tc_variant<GG, LL, 1, true>(e, p, ts) : launch_tc_variant<GG, LL, 2, true>(e, p, ts))  \
               : (variant == 1 ? launch_tc_variant<GG, LL, 1, false>(e, p, ts) : launch_tc_variant<GG, LL, 2, false>(e, p, ts)))
        if (lpt == 16 && tc_variant_smem(ts, 2, p.S_tab) <= 227 * 1024) return TSP_LAUNCH_TV(2, 16);
        if (lpt == 8 && tc_variant_smem(ts, 4, p.S_tab) <= 227 * 1024) return TSP_LAUNCH_TV(4, 8);
        if (lpt == 8 && tc_variant_smem(ts, 2, p.S_tab) <= 227 * 1024) return TSP_LAUNCH_TV(2, 8);
#undef TSP_LAUNCH_TV
        if (!e->has_fast) return fail(e, TSP_EINVAL, "internal: tcgen05 search kernel does not fit shared memory");
    }
    if (variant == 0 && fast_smem(e->fsched[which][0], 1, p.S_tab, false, 1) > 227 * 1024) {
        // the weights do not fit shared memory (cross-merge): stream them from L2, strips only in shared memory
        int G = lpt == 16 ? 2 : 4;
        if (e->fast_groups == 1 || e->fast_groups == 2 || (e->fast_groups == 4 && lpt == 8)) G = e->fast_groups;
        while (G > 1 && fast_smem(e->fsched[which][0], G, p.S_tab, false, 1, true) > 227 * 1024) G >>= 1;
        p.w_lo = 0;
        if (lpt == 8) {
            if (G == 4) return launch_fast<4, 8, true>(e, p);
            if (G == 2) return launch_fast<2, 8, true>(e, p);
            return launch_fast<1, 8, true>(e, p);
        }
        if (G == 2) return launch_fast<2, 16, true>(e, p);
        return launch_fast<1, 16, true>(e, p);
    }
    int G = lpt == 16 ? 2 : 4;
    if (e->fast_groups == 1 || e->fast_groups == 2 || (e->fast_groups == 4 && lpt == 8)) G = e->fast_groups;
    while (G > 1 && fast_smem(e->fsched[which][0], G, p.S_tab, false, n_sched) > 227 * 1024) G >>= 1;
    // host-split weights (no split arithmetic in the mma loop) when the doubled arena still fits
    p.w_lo = (!e->no_wlo && fast_smem(e->fsched[which][0], G, p.S_tab, true, n_sched) <= 227 * 1024) ? 1 : 0;
    if (variant != 0) {
#define TSP_LAUNCH_V(GG, LL) (variant == 1 ? launch_fast_variant<GG, LL, 1>(e, p) : launch_fast_variant<GG, LL, 2>(e, p))
        if (lpt == 8) return G == 4 ? TSP_LAUNCH_V(4, 8) : (G == 2 ? TSP_LAUNCH_V(2, 8) : TSP_LAUNCH_V(1, 8));
        return G == 2 ? TSP_LAUNCH_V(2, 16) : TSP_LAUNCH_V(1, 16);
#undef TSP_LAUNCH_V
    }
    if (lpt == 8) {
        if (G == 4) return launch_fast<4, 8>(e, p);
        if (G == 2) return launch_fast<2, 8>(e, p);
        return launch_fast<1, 8>(e, p);
    }
    if (G == 2) return launch_fast<2, 16>(e, p);
    return launch_fast<1, 16>(e, p);
}


int launch_search_any(tsp_engine* e, const SearchParams& p_in) {
    SearchParams p = p_in;
    // Networks whose weights do not fit shared memory (cross-merge) run on the fast kernel with the weights streamed
    // from L2 only while the batch is latency-bound (<= 2 groups per SM: 1.94 vs 2.19 ms at 4,096 x 50); for large
    // batches the generic kernel's 32-row tiles reuse every streamed weight twice as often (12.2 vs 13.5 ms at 8,192 x 200).
    const bool small_batch = (p.B + GT - 1) / GT <= 2 * e->sm_count;
    const bool stream_ok = e->cfg.variant == 0 && !e->no_wglobal && (small_batch || e->force_wglobal);
    if (e->nc == 5 && e->has_tc && !e->no_tc && !e->no_fast && e->finalized && e->cfg.variant == 0 && !p.fed && !p.mask_absorbing &&
        tc_smem(e->tcs[0], 2, p.S_tab) <= 227 * 1024 && tc_smem(e->tcs[1], 2, p.S_tab) <= 227 * 1024)
        return launch_fast_any(e, p);  // tcgen05 kernel (weights in tensor memory)
    if (e->nc == 5 && e->has_tcs && !e->no_tcs && !e->no_fast && e->finalized && e->cfg.variant == 0 && !p.fed && !p.mask_absorbing &&
        tcs_smem(e->tss, p.S_tab) <= 227 * 1024)
        return launch_fast_any(e, p);  // tcgen05 kernel, weights streamed by TMA
    if (e->nc == 5 && e->has_tc && !e->no_tc && !e->no_fast && e->finalized && e->cfg.variant != 0 && !p.fed &&
        tc_variant_smem(e->tcs[0], 2, p.S_tab) <= 227 * 1024 && tc_variant_smem(e->tcs[1], 2, p.S_tab) <= 227 * 1024)
        return launch_fast_any(e, p);
    if (e->nc == 5 && e->has_fast && !e->no_fast && e->finalized &&
        fast_smem(e->fsched[0][0], 1, p.S_tab, false, e->cfg.variant == 0 ? 1 : 2, stream_ok) <= 227 * 1024)
        return launch_fast_any(e, p);
    // 32 trees per CTA measured best at every batch size (4,096 .. 65,536 roots, both network sizes):
    // lane == row in the MLP phase wants a full warp of rows; 16 is the fallback when the activation
    // strips of 32 rows do not fit next to the weights.
    int TM = 32;
    if (search_smem(p.sched_a, 32, p.S_tab) > 110 * 1024) TM = 16;
    if (e->tm_override == 16 || e->tm_override == 32) TM = e->tm_override;
    if (e->nc != 5) TM = 32;
    // keep the search MLP weights in shared memory when two CTAs per SM still fit
    p.w_smem_floats = 0;
    if (p.fed == nullptr && !e->no_wsmem &&
        search_smem(p.sched_a, TM, p.S_tab, (size_t)e->search_floats) <= 112 * 1024)
        p.w_smem_floats = (int)e->search_floats;
    const int v = e->cfg.variant;
    // helper warps for the MLP units pay off while there is at most one CTA per SM (they double the
    // registers of a CTA); measured: 4,096 roots 1.13 -> 0.99 ms, 65,536 roots slower
    int xw = ((p.B + TM - 1) / TM <= e->sm_count && TM == 32) ? 2 : 1;
    p.warp_mlp = e->warp_mlp ? 1 : 0;
    if (e->xw_override == 1 || e->xw_override == 2) xw = e->xw_override;
    if (e->nc == 5) {
        if (v == 0) return launch_search_tm<0, 5>(e, p, TM, xw);
        if (v == 1) return launch_search_tm<1, 5>(e, p, TM, xw);
        return launch_search_tm<2, 5>(e, p, TM, xw);
    }
    if (v == 0) return launch_search_tm<0, 8>(e, p, 32, 1);
    if (v == 1) return launch_search_tm<1, 8>(e, p, 32, 1);
    return launch_search_tm<2, 8>(e, p, 32, 1);
}

// copy-in helper: returns the device pointer to use for an input array
template <typename T>
int stage_in(tsp_engine* e, int memory, const T* src, size_t count, DevBuf& buf, const T** out) {
    if (!src) {
        *out = nullptr;
        return TSP_OK;
    }
    if (memory == TSP_MEM_DEVICE) {
        *out = src;
        return TSP_OK;
    }
    int rc = ensure(e, buf, count * sizeof(T));
    if (rc) return rc;
    CK(e, cudaMemcpyAsync(buf.p, src, count * sizeof(T), cudaMemcpyHostToDevice, e->stream));
    e->stats.bytes_h2d += (int64_t)(count * sizeof(T));
    *out = reinterpret_cast<const T*>(buf.p);
    return TSP_OK;
}

// A HOST output array in pinned (page-locked) memory is written by the kernels directly: under unified addressing such
// memory is mapped into the device's address space, the few hundred KB of root statistics leave as posted PCIe writes while
// the kernel's tail still runs, and no device->host copy is queued behind the search (each one costs ~8 us of stream
// time). Pageable host arrays get a staging buffer and a copy. Returns the device alias or nullptr.
// (Only arrays up to 4 MB -- the root statistics; a pinned trace buffer of many MB written word by word from inside the
// simulation loop is better off staged.)
template <typename T>
T* pinned_alias(tsp_engine* e, T* host, size_t count) {
    if (e->no_zero_copy || count * sizeof(T) > ((size_t)4 << 20)) return nullptr;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, host) != cudaSuccess) {
        cudaGetLastError();  // (pageable memory on older drivers: an error, not a type)
        return nullptr;
    }
    if (at.type != cudaMemoryTypeHost || !at.devicePointer) return nullptr;
    return reinterpret_cast<T*>(at.devicePointer);
}

// `in_place`: false for the outputs of the ROOT kernel -- its few scattered words per row would leave as thousands of
// small PCIe writes the kernel has to drain before the search can start (measured: +24 us); they are staged and copied on
// the copy stream while the search runs instead.
template <typename T>
int stage_out(tsp_engine* e, int memory, T* dst, size_t count, DevBuf& buf, T** out, bool in_place = true) {
    if (!dst) {
        *out = nullptr;
        return TSP_OK;
    }
    if (memory == TSP_MEM_DEVICE) {
        *out = dst;
        return TSP_OK;
    }
    if (T* alias = in_place ? pinned_alias(e, dst, count) : nullptr) {
        *out = alias;
        e->stats.bytes_d2h += (int64_t)(count * sizeof(T));
        return TSP_OK;
    }
    int rc = ensure(e, buf, count * sizeof(T));
    if (rc) return rc;
    *out = reinterpret_cast<T*>(buf.p);
    return TSP_OK;
}

template <typename T>
int copy_out(tsp_engine* e, int memory, T* dst, const T* dev, size_t count, cudaStream_t stream = nullptr) {
    if (!dst || memory == TSP_MEM_DEVICE) return TSP_OK;
    {   // written in place: `dev` is the device alias of a pinned HOST array (stage_out)
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, dev) == cudaSuccess) {
            if (at.type == cudaMemoryTypeHost) return TSP_OK;
        } else {
            cudaGetLastError();
        }
    }
    CK(e, cudaMemcpyAsync(dst, dev, count * sizeof(T), cudaMemcpyDeviceToHost, stream ? stream : e->stream));
    e->stats.bytes_d2h += (int64_t)(count * sizeof(T));
    return TSP_OK;
}

// accumulate the device times of every completed run (waits for the pending ones)
int drain_events(tsp_engine* e) {
    while (e->ev_pending > 0) {
        cudaEvent_t* ev = &e->ev[3 * e->ev_head];
        CK(e, cudaEventSynchronize(ev[2]));
        float r = 0.f, s = 0.f;
        CK(e, cudaEventElapsedTime(&r, ev[0], ev[1]));
        CK(e, cudaEventElapsedTime(&s, ev[1], ev[2]));
        e->stats.last_root_ms = r;
        e->stats.last_search_ms = s;
        e->stats.sum_root_ms += r;
        e->stats.sum_search_ms += s;
        e->stats.timed_runs++;
        e->ev_head = (e->ev_head + 1) % tsp_engine::EV_RING;
        e->ev_pending--;
    }
    return TSP_OK;
}

}  // namespace

// ================================================================== C ABI ====
extern "C" {

int tsp_abi_version(void) { return TSP_ABI_VERSION; }

const char* tsp_last_error(tsp_handle h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int tsp_create(const tsp_config* cfg, tsp_handle* out) {
    if (!cfg || !out) return fail(nullptr, TSP_EINVAL, "tsp_create: null argument");
    *out = nullptr;
    if (cfg->num_players < 1 || cfg->num_players > 2)  // self_play.py:502-503 raises NotImplementedError
        return fail(nullptr, TSP_ENOTIMPL, "more than two player mode not implemented (len(players) == %d)", cfg->num_players);
    if (cfg->num_players == 2 && cfg->variant != TSP_VARIANT_MUZERO)  // self_play_stochastic.py:407 asserts one player
        return fail(nullptr, TSP_EINVAL, "the chance-node searches require len(players) == 1");
    if (cfg->variant < 0 || cfg->variant > 2) return fail(nullptr, TSP_EINVAL, "unknown variant %d", cfg->variant);
    if (cfg->num_actions < 1 || cfg->num_actions > TSP_MAX_ACTIONS)
        return fail(nullptr, TSP_EINVAL, "num_actions must be in 1..%d", TSP_MAX_ACTIONS);
    if (cfg->encoding_size < 4 || cfg->encoding_size % 4)
        return fail(nullptr, TSP_EINVAL, "encoding_size must be a positive multiple of 4");
    if (cfg->support_size < 0 || cfg->support_size > 300) return fail(nullptr, TSP_EINVAL, "bad support_size");
    if (cfg->max_roots < 1 || cfg->max_simulations < 0) return fail(nullptr, TSP_EINVAL, "bad capacity");
    if (cfg->variant != 0 && (cfg->n_futures < 1 || cfg->n_futures > TSP_MAX_ACTIONS))
        return fail(nullptr, TSP_EINVAL, "n_futures must be in 1..%d", TSP_MAX_ACTIONS);
    int ndev = 0;
    cudaError_t st = cudaGetDeviceCount(&ndev);
    if (st != cudaSuccess || ndev == 0)
        return fail(nullptr, TSP_ENODEVICE, "no CUDA device (%s); this engine has no CPU fallback",
                    st == cudaSuccess ? "device count is 0" : cudaGetErrorString(st));
    if (cfg->device < 0 || cfg->device >= ndev) return fail(nullptr, TSP_EINVAL, "device %d out of range", cfg->device);
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, cfg->device) != cudaSuccess)
        return fail(nullptr, TSP_ECUDA, "cudaGetDeviceProperties failed");
    if (prop.major != 10)
        return fail(nullptr, TSP_ENODEVICE, "device %d is sm_%d%d; this library is built for sm_100a (B200) only",
                    cfg->device, prop.major, prop.minor);
    tsp_engine* e = new tsp_engine();
    e->cfg = *cfg;
    if (e->cfg.variant == 0) e->cfg.n_futures = 1;
    e->device = cfg->device;
    e->sm_count = prop.multiProcessorCount;
    e->nc = (std::max(cfg->num_actions, e->cfg.n_futures) <= 5) ? 5 : 8;
    if (const char* t = getenv("TSP_TREES_PER_CTA")) e->tm_override = atoi(t);
    if (const char* t = getenv("TSP_NO_WSMEM")) e->no_wsmem = atoi(t) != 0;
    if (const char* t = getenv("TSP_THREAD_MULT")) e->xw_override = atoi(t);
    if (const char* t = getenv("TSP_PHASE_TIMING")) e->phase_timing = atoi(t) != 0;
    if (const char* t = getenv("TSP_COLS_PER_THREAD")) e->ct_override = atoi(t);
    if (const char* t = getenv("TSP_TILES_PER_WARP")) e->ng_override = atoi(t);
    if (const char* t = getenv("TSP_WARP_MLP")) e->warp_mlp = atoi(t) != 0;
    if (const char* t = getenv("TSP_NO_FAST")) e->no_fast = atoi(t) != 0;
    if (const char* t = getenv("TSP_NO_WLO")) e->no_wlo = atoi(t) != 0;
    if (const char* t = getenv("TSP_NO_TC")) e->no_tc = atoi(t) != 0;
    if (const char* t = getenv("TSP_NO_TCS")) e->no_tcs = atoi(t) != 0;
    if (const char* t = getenv("TSP_NO_RTCS")) e->no_rtcs = atoi(t) != 0;
    if (const char* t = getenv("TSP_NO_OVERLAP")) e->no_overlap = atoi(t) != 0;
    if (const char* t = getenv("TSP_NO_ZERO_COPY")) e->no_zero_copy = atoi(t) != 0;
    if (const char* t = getenv("TSP_NO_BALANCE")) e->no_balance = atoi(t) != 0;
    if (const char* t = getenv("TSP_TC_NOMERGE")) e->no_merge = atoi(t) != 0;
    if (const char* t = getenv("TSP_TC_NOFOLD")) e->tc_nofold = atoi(t) != 0;
    if (const char* t = getenv("TSP_NO_WGLOBAL")) e->no_wglobal = atoi(t) != 0;
    if (const char* t = getenv("TSP_FORCE_WGLOBAL")) e->force_wglobal = atoi(t) != 0;
    if (const char* t = getenv("TSP_FAST_GROUPS")) e->fast_groups = atoi(t);
    if (const char* t = getenv("TSP_FAST_LPT")) e->fast_lpt = atoi(t);
    auto bail = [&](int rc) {
        g_create_error = e->err;
        tsp_destroy(e);
        return rc;
    };
    if (cudaSetDevice(e->device) != cudaSuccess) return bail(fail(e, TSP_ECUDA, "cudaSetDevice failed"));
    if (cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking) != cudaSuccess)
        return bail(fail(e, TSP_ECUDA, "cudaStreamCreate failed"));
    if (cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking) != cudaSuccess)
        return bail(fail(e, TSP_ECUDA, "cudaStreamCreate failed"));
    for (auto& ce : e->ev_chunk)
        if (cudaEventCreateWithFlags(&ce, cudaEventDisableTiming) != cudaSuccess) return bail(fail(e, TSP_ECUDA, "cudaEventCreate failed"));
    if (cudaEventCreateWithFlags(&e->ev_root_out, cudaEventDisableTiming) != cudaSuccess) return bail(fail(e, TSP_ECUDA, "cudaEventCreate failed"));
    for (auto& is : e->in_set)
        if (cudaEventCreateWithFlags(&is.done, cudaEventDisableTiming) != cudaSuccess) return bail(fail(e, TSP_ECUDA, "cudaEventCreate failed"));
    e->ev.resize(3 * tsp_engine::EV_RING, nullptr);
    for (auto& ev : e->ev)
        if (cudaEventCreate(&ev) != cudaSuccess) return bail(fail(e, TSP_ECUDA, "cudaEventCreate failed"));

    const int S = cfg->max_simulations, A = cfg->num_actions, nf = e->cfg.n_futures, H = cfg->encoding_size;
    const size_t B = (size_t)cfg->max_roots;
    if (cfg->variant == 0) {
        e->blocks_per_tree = S + 1;
        e->hidden_per_tree = S + 1;
        e->max_iters = S;
    } else {
        // state nodes <= S + 1; every expanded state node has A transition children
        e->blocks_per_tree = (S + 1) * (A + 1);
        e->hidden_per_tree = 1 + (S + 1) * A * nf;
        e->max_iters = S + (S + 1) * A + 1;
    }
    const size_t block_bytes = e->nc == 5 ? sizeof(NodeBlock<5>) : sizeof(NodeBlock<8>);
    int rc;
    if ((rc = ensure(e, e->blocks, B * e->blocks_per_tree * block_bytes))) return bail(rc);
    if ((rc = ensure(e, e->hidden, B * e->hidden_per_tree * H * sizeof(float)))) return bail(rc);
    if ((rc = ensure(e, e->hdr, B * sizeof(TreeHeader)))) return bail(rc);
    if ((rc = ensure(e, e->rootprior, B * GW * sizeof(double)))) return bail(rc);
    if ((rc = ensure(e, e->rootact, B * GW * sizeof(int)))) return bail(rc);
    // pb_c table: log((N + pb_c_base + 1) / pb_c_base) + pb_c_init with the host libm, the same
    // correctly-rounded log CPython's math.log calls (self_play.py:457-462)
    // ... followed by sqrt(N) (IEEE-exact, same as math.sqrt) so that the kernel never runs a double sqrt.
    // The tables are laid out for max_simulations; tsp_run passes S == max_simulations-independent offsets.
    std::vector<double> pb(2 * (S + 2));
    for (int n = 0; n < S + 2; ++n) {
        pb[n] = std::log(((double)n + cfg->pb_c_base + 1.0) / cfg->pb_c_base) + cfg->pb_c_init;
        pb[S + 2 + n] = std::sqrt((double)n);
    }
    if ((rc = ensure(e, e->pbtable, pb.size() * sizeof(double)))) return bail(rc);
    if (cudaMemcpy(e->pbtable.p, pb.data(), pb.size() * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess)
        return bail(fail(e, TSP_ECUDA, "pb table upload failed"));
    if (e->phase_timing) {
        if ((rc = ensure(e, e->phase, 64 * sizeof(unsigned long long)))) return bail(rc);
        cudaMemset(e->phase.p, 0, 64 * sizeof(unsigned long long));
    }
    e->stats.sm_count = e->sm_count;
    *out = e;
    return TSP_OK;
}

int tsp_destroy(tsp_handle e) {
    if (!e) return TSP_OK;
    cudaSetDevice(e->device);
    if (e->stream) cudaStreamSynchronize(e->stream);
    for (DevBuf* b : {&e->arena, &e->farena, &e->fsched_dev[0][0], &e->fsched_dev[0][1], &e->fsched_dev[1][0],
                      &e->fsched_dev[1][1], &e->blocks, &e->hidden, &e->hdr, &e->rootprior, &e->rootact, &e->pbtable, &e->phase, &e->s_obs,
                      &e->s_legal, &e->s_nlegal, &e->s_noise, &e->s_tie, &e->s_chance, &e->s_fedroot, &e->s_fed,
                      &e->s_visits, &e->s_rootv, &e->s_cval, &e->s_cprior, &e->s_crew, &e->s_depth, &e->s_rpred,
                      &e->s_rhid, &e->s_nrec, &e->s_troot, &e->s_trace, &e->s_tdepth, &e->n_in, &e->n_act, &e->tc_img[0],
                      &e->tc_img[1], &e->tc_ctab[0], &e->tc_ctab[1], &e->tc_status, &e->ts_img, &e->ts_ctab, &e->rts_img, &e->rts_ctab, &e->rts_hidden})
        release(*b);
    for (auto& b : e->n_o) release(b);
    for (auto& ev : e->ev)
        if (ev) cudaEventDestroy(ev);
    if (e->ev_root_out) cudaEventDestroy(e->ev_root_out);
    for (auto& ce : e->ev_chunk)
        if (ce) cudaEventDestroy(ce);
    for (auto& is : e->in_set) {
        if (is.done) cudaEventDestroy(is.done);
        for (DevBuf* b : {&is.obs, &is.legal, &is.nlegal, &is.noise, &is.tie, &is.chance}) release(*b);
    }
    if (e->copy_stream) cudaStreamDestroy(e->copy_stream);
    if (e->stream) cudaStreamDestroy(e->stream);
    delete e;
    return TSP_OK;
}

int tsp_set_layer(tsp_handle e, int32_t mlp, int32_t layer, const float* W, const float* b, int32_t out_dim,
                  int32_t in_dim) {
    if (!e) return TSP_EINVAL;
    if (mlp < 0 || mlp >= TSP_NUM_MLP || layer < 0 || layer >= TSP_MAX_LAYERS || !W || !b || out_dim < 1 || in_dim < 1)
        return fail(e, TSP_EINVAL, "tsp_set_layer: bad argument (mlp %d layer %d %dx%d)", mlp, layer, out_dim, in_dim);
    auto& v = e->mlp[mlp];
    if ((int)v.size() <= layer) v.resize(layer + 1);
    HostLayer& L = v[layer];
    L.out_dim = out_dim;
    L.in_dim = in_dim;
    L.W.assign(W, W + (size_t)out_dim * in_dim);
    L.b.assign(b, b + out_dim);
    e->finalized = false;
    return TSP_OK;
}

int tsp_finalize_weights(tsp_handle e) {
    if (!e) return TSP_EINVAL;
    CK(e, cudaSetDevice(e->device));
    // pack: per layer Wt[in_dim][wld] (transposed: row k holds the weights of input k for all outputs),
    // zero padded to npad = a multiple of 8 columns (one mma n-tile), followed by the bias [npad].
    // The leading dimension wld is == 8 (mod 32) words so that the B fragments of mma.m16n8k8
    // (rows k0 + lane % 4, columns n0 + lane / 4) hit 32 distinct shared-memory banks.
    int64_t total = 0;
    // MuZeroStochastic (models.py:298-462): one dynamics MLP per future. The arena then starts with n_futures weight
    // groups of identical layout, {dynamics of future z, reward head, prior head} (the shared heads are repeated), so
    // that the transition schedule of future z is the schedule of future 0 on a base pointer z * f_stride further on.
    const int nf = e->cfg.n_futures;
    e->separate_heads = !e->mlp[TSP_MLP_DYNAMICS_F1].empty();
    e->f_stride = 0;
    if (e->separate_heads) {
        if (e->cfg.variant == TSP_VARIANT_MUZERO || nf < 2 || nf > 4)
            return fail(e, TSP_EINVAL, "per-future dynamics heads need a chance-node variant with 2..4 futures");
        const auto& d0 = e->mlp[TSP_MLP_DYNAMICS];
        for (int z = 1; z < 4; ++z) {
            const auto& dz = e->mlp[TSP_MLP_DYNAMICS_F1 + z - 1];
            if (z >= nf) {
                if (!dz.empty()) return fail(e, TSP_EINVAL, "dynamics head of future %d set, but n_futures is %d", z, nf);
                continue;
            }
            if (dz.size() != d0.size()) return fail(e, TSP_EINVAL, "dynamics head of future %d: %zu layers, future 0 has %zu", z, dz.size(), d0.size());
            for (size_t l = 0; l < dz.size(); ++l)
                if (dz[l].out_dim != d0[l].out_dim || dz[l].in_dim != d0[l].in_dim)
                    return fail(e, TSP_EINVAL, "dynamics head of future %d: layer %zu shape differs from future 0", z, l);
        }
    }
    const int order[TSP_NUM_MLP] = {TSP_MLP_DYNAMICS, TSP_MLP_REWARD, TSP_MLP_PRIOR, TSP_MLP_DYNAMICS_F1, TSP_MLP_DYNAMICS_F2,
                                    TSP_MLP_DYNAMICS_F3, TSP_MLP_TERMINAL, TSP_MLP_POLICY, TSP_MLP_VALUE, TSP_MLP_REPRESENTATION};
    for (int oi = 0; oi < TSP_NUM_MLP; ++oi) {
        auto& v = e->mlp[order[oi]];
        if (order[oi] == TSP_MLP_REPRESENTATION) e->search_floats = total;
        if (order[oi] == TSP_MLP_DYNAMICS_F1 && e->separate_heads) e->f_stride = (int)total;  // == size of group 0
        for (size_t l = 0; l < v.size(); ++l) {
            HostLayer& L = v[l];
            if (L.out_dim == 0) return fail(e, TSP_ESTATE, "a Linear layer is missing (set every layer index)");
            if (l > 0 && v[l - 1].out_dim != L.in_dim) return fail(e, TSP_EINVAL, "layer shapes do not chain");
            L.npad = (L.out_dim + 7) & ~7;
            L.wld = L.npad;
            while (L.wld % 32 != 8 && L.wld % 32 != 24) L.wld += 8;
            L.w_off = (int)total;
            total += (int64_t)L.in_dim * L.wld;
            L.b_off = (int)total;
            total += L.npad;
        }
        // group z >= 1: the dynamics MLP of future z is followed by room for the repeated reward / prior heads
        if (e->separate_heads && order[oi] >= TSP_MLP_DYNAMICS_F1 && order[oi] <= TSP_MLP_DYNAMICS_F3 && !v.empty()) {
            const int z = order[oi] - TSP_MLP_DYNAMICS_F1 + 1;
            total = (int64_t)(z + 1) * e->f_stride;
        }
    }
    if (total == 0) return fail(e, TSP_ESTATE, "no weights loaded");
    std::vector<float> host((size_t)total, 0.0f);
    for (auto& v : e->mlp)
        for (HostLayer& L : v) {
            for (int o = 0; o < L.out_dim; ++o)
                for (int i = 0; i < L.in_dim; ++i) host[(size_t)L.w_off + (size_t)i * L.wld + o] = L.W[(size_t)o * L.in_dim + i];
            for (int o = 0; o < L.out_dim; ++o) host[(size_t)L.b_off + o] = L.b[o];
        }
    if (e->separate_heads) {  // repeat the shared heads of group 0 behind every other future's dynamics MLP
        const int64_t heads0 = e->mlp[TSP_MLP_REWARD].front().w_off;
        for (int z = 1; z < nf; ++z)
            std::copy(host.begin() + heads0, host.begin() + e->f_stride, host.begin() + (int64_t)z * e->f_stride + heads0);
    }
    int rc = ensure(e, e->arena, (size_t)total * sizeof(float));
    if (rc) return rc;
    CK(e, cudaMemcpyAsync(e->arena.p, host.data(), (size_t)total * sizeof(float), cudaMemcpyHostToDevice, e->stream));
    CK(e, cudaStreamSynchronize(e->stream));
    e->arena_floats = total;
    if ((rc = pack_fast_arena(e))) return rc;
    if ((rc = build_schedules(e))) return rc;
    e->finalized = true;
    return TSP_OK;
}

int tsp_weight_arena(tsp_handle e, void** device_ptr, int64_t* num_bytes) {
    if (!e || !e->finalized) return e ? fail(e, TSP_ESTATE, "weights not finalized") : TSP_EINVAL;
    if (device_ptr) *device_ptr = e->arena.p;
    if (num_bytes) *num_bytes = e->arena_floats * (int64_t)sizeof(float);
    return TSP_OK;
}

int tsp_run(tsp_handle e, const tsp_run_args* a) {
    if (!e || !a) return TSP_EINVAL;
    const tsp_config& c = e->cfg;
    const int B = a->num_roots, S = a->num_simulations, A = c.num_actions, H = c.encoding_size;
    const bool fed = a->fed_root != nullptr || a->fed_records != nullptr;
    if (!fed && !e->finalized) return fail(e, TSP_ESTATE, "tsp_run before tsp_finalize_weights");
    if (B < 1 || B > c.max_roots) return fail(e, TSP_EINVAL, "num_roots %d outside 1..%d", B, c.max_roots);
    if (S < 0 || S > c.max_simulations) return fail(e, TSP_EINVAL, "num_simulations %d outside 0..%d", S, c.max_simulations);
    if (fed && ((!a->fed_root && !a->resume) || (S > 0 && !a->fed_records)))
        return fail(e, TSP_EINVAL, "fed mode needs fed_root and fed_records");
    const bool resume = a->resume != 0;
    if (resume) {
        if (c.variant == TSP_VARIANT_STOCHASTIC)  // its first assert fails on a visited root, self_play_stochastic.py:333
            return fail(e, TSP_EINVAL, "resume: the stochastic search cannot be continued (the reference asserts "
                                       "num_sims == sum of the root's child visits, self_play_stochastic.py:333)");
        if (e->tree_roots != B) return fail(e, TSP_ESTATE, "resume: the previous tsp_run held %d roots, not %d", e->tree_roots, B);
        if (e->tree_sims + S > c.max_simulations)
            return fail(e, TSP_EINVAL, "resume: %d + %d simulations exceed max_simulations %d", e->tree_sims, S, c.max_simulations);
        if (fed != e->tree_fed) return fail(e, TSP_EINVAL, "resume: fed / network mode must match the search being continued");
    }
    if (!fed && !resume && !a->observations) return fail(e, TSP_EINVAL, "observations missing");
    if (a->add_exploration_noise && !a->noise) return fail(e, TSP_EINVAL, "add_exploration_noise without a noise array");
    if ((a->legal_actions == nullptr) != (a->num_legal == nullptr))
        return fail(e, TSP_EINVAL, "legal_actions and num_legal must be given together");
    if ((fed || a->trace_records) && a->max_records < 1 && S > 0) return fail(e, TSP_EINVAL, "max_records must be positive");
    if (a->tie_stream && a->tie_len < 1) return fail(e, TSP_EINVAL, "tie_stream with tie_len < 1");
    if (a->chance_stream && a->chance_len < 1) return fail(e, TSP_EINVAL, "chance_stream with chance_len < 1");
    if (c.variant != 0 && c.n_futures > 1 && !a->chance_stream && S > 0)
        return fail(e, TSP_EINVAL, "the stochastic variants need a chance_stream");
    if (a->memory == TSP_MEM_HOST && a->legal_actions) {
        // mirrors the asserts of self_play.py:354-359
        for (int b = 0; b < B; ++b) {
            const int nl = a->num_legal[b];
            if (nl < 1 || nl > A) return fail(e, TSP_EINVAL, "root %d: legal actions should not be an empty array (got %d)", b, nl);
            for (int i = 0; i < nl; ++i) {
                const int x = a->legal_actions[(size_t)b * A + i];
                if (x < 0 || x >= A) return fail(e, TSP_EINVAL, "root %d: legal actions should be a subset of the action space", b);
            }
        }
    }
    CK(e, cudaSetDevice(e->device));
    const int mem = a->memory;
    const int omem = (mem == TSP_MEM_HOST && a->outputs_on_device) ? TSP_MEM_DEVICE : mem;  // where the OUTPUT arrays live
    const size_t nb = (size_t)B;
    int rc;

    NetParams np;
    memset(&np, 0, sizeof(np));
    np.B = B; np.A = A; np.H = H; np.nf = c.n_futures; np.support = c.support_size; np.mode = 0;
    np.W = reinterpret_cast<const float*>(e->arena.p);
    np.tree_init = 1;
    np.add_noise = a->add_exploration_noise ? 1 : 0;
    np.uniform_policy = a->uniform_policy ? 1 : 0;
    np.blocks_per_tree = e->blocks_per_tree;
    np.hidden_per_tree = e->hidden_per_tree;
    np.frac = c.root_exploration_fraction;
    np.one_minus_frac = 1 - c.root_exploration_fraction;
    np.blocks = e->blocks.p;
    np.hidden_pool = reinterpret_cast<float*>(e->hidden.p);
    np.hdr = reinterpret_cast<TreeHeader*>(e->hdr.p);
    np.rootprior = reinterpret_cast<double*>(e->rootprior.p);
    np.rootact = reinterpret_cast<int*>(e->rootact.p);
    np.sched = e->sch_root;
    if (fed) {  // the network is bypassed: any well-formed schedule shape will do
        memset(&np.sched, 0, sizeof(np.sched));
        np.sched.row_stride = 36;
    }
    // Large HOST batches: every input crosses PCIe on the copy stream into one of two staging sets (used alternately), so
    // that the upload of this run overlaps the search of the previous one; the launches below wait for the copies.
    const bool overlap = mem == TSP_MEM_HOST && !resume && !fed && a->observations && B >= 1024 && !e->no_overlap && e->copy_stream;
    tsp_engine::InSet* is = nullptr;
    // on the copy stream: host array -> staging buffer of the set
    auto upload = [&](const void* src, size_t bytes, DevBuf& buf, const void** out) -> int {
        *out = nullptr;
        if (!src || bytes == 0) return TSP_OK;
        int r = ensure(e, buf, bytes);
        if (r) return r;
        CK(e, cudaMemcpyAsync(buf.p, src, bytes, cudaMemcpyHostToDevice, e->copy_stream));
        e->stats.bytes_h2d += (int64_t)bytes;
        *out = buf.p;
        return TSP_OK;
    };
    if (overlap) {
        is = &e->in_set[e->in_flip];
        e->in_flip ^= 1;
        if ((rc = ensure(e, is->obs, nb * c.obs_dim * sizeof(float)))) return rc;
        np.in = reinterpret_cast<const float*>(is->obs.p);
        // the set is free once the search that read it (two runs ago) is done
        CK(e, cudaStreamWaitEvent(e->copy_stream, is->done, 0));
        const void* d = nullptr;
        if ((rc = upload(a->legal_actions, nb * A * sizeof(int32_t), is->legal, &d))) return rc;
        np.legal = reinterpret_cast<const int32_t*>(d);
        if ((rc = upload(a->num_legal, nb * sizeof(int32_t), is->nlegal, &d))) return rc;
        np.num_legal = reinterpret_cast<const int32_t*>(d);
        if ((rc = upload(a->noise, nb * A * sizeof(double), is->noise, &d))) return rc;
        np.noise = reinterpret_cast<const double*>(d);
    } else {
        if ((rc = stage_in(e, mem, a->observations, nb * c.obs_dim, e->s_obs, &np.in))) return rc;
        if ((rc = stage_in(e, mem, a->legal_actions, nb * A, e->s_legal, &np.legal))) return rc;
        if ((rc = stage_in(e, mem, a->num_legal, nb, e->s_nlegal, &np.num_legal))) return rc;
        if ((rc = stage_in(e, mem, a->noise, nb * A, e->s_noise, &np.noise))) return rc;
    }
    if ((rc = stage_in(e, mem, a->fed_root, nb * REC, e->s_fedroot, &np.fed_root))) return rc;
    if ((rc = stage_out(e, omem, a->trace_root, nb * REC, e->s_troot, &np.trace_root, false))) return rc;
    if ((rc = stage_out(e, omem, a->root_predicted_value, nb, e->s_rpred, &np.root_pred, false))) return rc;
    if ((rc = stage_out(e, omem, a->root_hidden, nb * H, e->s_rhid, &np.root_hidden, false))) return rc;

    SearchParams sp;
    memset(&sp, 0, sizeof(sp));
    sp.B = B; sp.S = S; sp.A = A; sp.H = H; sp.nf = c.n_futures; sp.support = c.support_size;
    sp.mask_absorbing = c.mask_absorbing_states ? 1 : 0;
    sp.uniform_policy = a->uniform_policy ? 1 : 0;
    sp.blocks_per_tree = e->blocks_per_tree;
    sp.hidden_per_tree = e->hidden_per_tree;
    sp.max_rec = a->max_records;
    sp.T = a->tie_stream ? a->tie_len : 0;
    sp.K = a->chance_stream ? a->chance_len : 0;
    // chance-node variants: every iteration either counts a simulation or expands one of the <= (states) * A
    // transition nodes; a continued search may still find unexpanded transitions of the earlier searches
    const int S_tree = (resume ? e->tree_sims : 0) + S;
    sp.S_tab = S_tree;
    sp.max_iters = (c.variant == 0) ? S : S + (S_tree + 1) * A + 1;
    sp.discount = c.discount;
    sp.two_player = c.num_players == 2 ? 1 : 0;
    sp.resume = resume ? 1 : 0;
    sp.add_noise = a->add_exploration_noise ? 1 : 0;
    sp.frac = c.root_exploration_fraction;
    sp.one_minus_frac = 1 - c.root_exploration_fraction;
    sp.noise = np.noise;
    sp.pb_table = reinterpret_cast<const double*>(e->pbtable.p);
    sp.pb_stride = c.max_simulations + 2;
    sp.blocks = e->blocks.p;
    sp.hidden_pool = reinterpret_cast<float*>(e->hidden.p);
    sp.hdr = reinterpret_cast<TreeHeader*>(e->hdr.p);
    sp.rootprior = reinterpret_cast<double*>(e->rootprior.p);
    sp.rootact = reinterpret_cast<const int*>(e->rootact.p);
    sp.W = reinterpret_cast<const float*>(e->arena.p);
    sp.phase_cycles = reinterpret_cast<unsigned long long*>(e->phase.p);
    if (c.variant == 0) {
        sp.sched_a = e->sch_recur;
    } else {
        sp.sched_a = e->sch_trans;
        sp.sched_b = e->sch_pred;
    }
    if (fed) {
        memset(&sp.sched_a, 0, sizeof(sp.sched_a));
        memset(&sp.sched_b, 0, sizeof(sp.sched_b));
        sp.sched_a.row_stride = 4;
        sp.sched_b.row_stride = 4;
    }
    const size_t nrec = nb * (size_t)std::max(a->max_records, 0);
    if (overlap) {
        const void* d = nullptr;
        if ((rc = upload(a->tie_stream, nb * sp.T, is->tie, &d))) return rc;
        sp.tie = reinterpret_cast<const unsigned char*>(d);
        if ((rc = upload(a->chance_stream, nb * sp.K, is->chance, &d))) return rc;
        sp.chance = reinterpret_cast<const unsigned char*>(d);
    } else {
        if ((rc = stage_in(e, mem, a->tie_stream, nb * sp.T, e->s_tie, &sp.tie))) return rc;
        if ((rc = stage_in(e, mem, a->chance_stream, nb * sp.K, e->s_chance, &sp.chance))) return rc;
    }
    if ((rc = stage_in(e, mem, a->fed_records, nrec * REC, e->s_fed, &sp.fed))) return rc;
    if ((rc = stage_out(e, omem, a->trace_records, nrec * REC, e->s_trace, &sp.trace))) return rc;
    if ((rc = stage_out(e, omem, a->visit_counts, nb * A, e->s_visits, &sp.visit_counts))) return rc;
    if ((rc = stage_out(e, omem, a->root_value, nb, e->s_rootv, &sp.root_value))) return rc;
    if ((rc = stage_out(e, omem, a->child_value, nb * A, e->s_cval, &sp.child_value))) return rc;
    if ((rc = stage_out(e, omem, a->child_prior, nb * A, e->s_cprior, &sp.child_prior))) return rc;
    if ((rc = stage_out(e, omem, a->child_reward, nb * A, e->s_crew, &sp.child_reward))) return rc;
    if ((rc = stage_out(e, omem, a->max_tree_depth, nb, e->s_depth, &sp.max_depth))) return rc;
    if ((rc = stage_out(e, omem, a->num_records, nb, e->s_nrec, &sp.num_records))) return rc;
    if ((rc = stage_out(e, omem, a->total_select_depth, nb, e->s_tdepth, &sp.total_depth))) return rc;
    if (fed && a->fed_records == nullptr) sp.fed = reinterpret_cast<const float*>(e->hdr.p);  // S == 0: never read

    if (e->ev_pending == tsp_engine::EV_RING && (rc = drain_events(e))) return rc;
    cudaEvent_t* ev = &e->ev[3 * ((e->ev_head + e->ev_pending) % tsp_engine::EV_RING)];
    CK(e, cudaEventRecord(ev[0], e->stream));
    if (overlap) {
        // A search still in flight hides the whole upload: one copy, ONE root launch behind it. On an idle engine the
        // observations go up in OBS_CHUNKS row chunks and the root network of a chunk starts as soon as its rows have
        // arrived, so that only the last chunk's root kernel is exposed (end to end = upload + search).
        const bool busy = cudaStreamQuery(e->stream) == cudaErrorNotReady;
        const int want = busy ? 1 : tsp_engine::OBS_CHUNKS;
        const int per = ((B + want - 1) / want + 31) & ~31;
        int n_chunks = 0;
        for (int r0 = 0; r0 < B; r0 += per, ++n_chunks) {
            const size_t rows = (size_t)std::min(per, B - r0);
            CK(e, cudaMemcpyAsync(reinterpret_cast<float*>(is->obs.p) + (size_t)r0 * c.obs_dim, a->observations + (size_t)r0 * c.obs_dim,
                                  rows * c.obs_dim * sizeof(float), cudaMemcpyHostToDevice, e->copy_stream));
            CK(e, cudaEventRecord(e->ev_chunk[n_chunks], e->copy_stream));
        }
        e->stats.bytes_h2d += (int64_t)(nb * c.obs_dim * sizeof(float));
        n_chunks = 0;
        for (int r0 = 0; r0 < B; r0 += per, ++n_chunks) {
            CK(e, cudaStreamWaitEvent(e->stream, e->ev_chunk[n_chunks], 0));
            NetParams part = np;
            part.b0 = r0;
            part.B = std::min(B, r0 + per);
            if ((rc = launch_net_any(e, part))) return rc;
        }
    } else if (!resume && (rc = launch_net_any(e, np))) {
        return rc;
    }
    CK(e, cudaEventRecord(ev[1], e->stream));
    // the root kernel's HOST outputs leave on the copy stream while the search runs (HOST batches with the overlapped upload)
    const bool root_out_early = overlap && !resume && omem == TSP_MEM_HOST && (a->trace_root || a->root_predicted_value || a->root_hidden);
    if (root_out_early) {
        CK(e, cudaStreamWaitEvent(e->copy_stream, ev[1], 0));
        if ((rc = copy_out(e, omem, a->trace_root, np.trace_root, nb * REC, e->copy_stream))) return rc;
        if ((rc = copy_out(e, omem, a->root_predicted_value, np.root_pred, nb, e->copy_stream))) return rc;
        if ((rc = copy_out(e, omem, a->root_hidden, np.root_hidden, nb * H, e->copy_stream))) return rc;
        CK(e, cudaEventRecord(e->ev_root_out, e->copy_stream));
    }
    if ((rc = launch_search_any(e, sp))) return rc;
    CK(e, cudaEventRecord(ev[2], e->stream));
    if (is) CK(e, cudaEventRecord(is->done, e->stream));
    e->ev_pending++;

    if (root_out_early) {
        CK(e, cudaStreamWaitEvent(e->stream, e->ev_root_out, 0));  // (long done: keeps "the engine's stream orders everything")
    } else if (!resume) {  // the root is not evaluated again on resume (root_predicted_value is None in the reference)
        if ((rc = copy_out(e, omem, a->trace_root, np.trace_root, nb * REC))) return rc;
        if ((rc = copy_out(e, omem, a->root_predicted_value, np.root_pred, nb))) return rc;
        if ((rc = copy_out(e, omem, a->root_hidden, np.root_hidden, nb * H))) return rc;
    }
    if ((rc = copy_out(e, omem, a->trace_records, sp.trace, nrec * REC))) return rc;
    if ((rc = copy_out(e, omem, a->visit_counts, sp.visit_counts, nb * A))) return rc;
    if ((rc = copy_out(e, omem, a->root_value, sp.root_value, nb))) return rc;
    if ((rc = copy_out(e, omem, a->child_value, sp.child_value, nb * A))) return rc;
    if ((rc = copy_out(e, omem, a->child_prior, sp.child_prior, nb * A))) return rc;
    if ((rc = copy_out(e, omem, a->child_reward, sp.child_reward, nb * A))) return rc;
    if ((rc = copy_out(e, omem, a->max_tree_depth, sp.max_depth, nb))) return rc;
    if ((rc = copy_out(e, omem, a->num_records, sp.num_records, nb))) return rc;
    if ((rc = copy_out(e, omem, a->total_select_depth, sp.total_depth, nb))) return rc;
    e->stats.searches++;
    e->tree_roots = B;
    e->tree_sims = (resume ? e->tree_sims : 0) + S;
    e->tree_fed = fed;
    return TSP_OK;
}

static int net_call(tsp_handle e, int mode, const Schedule& sc, int memory, int B, const float* in, int in_dim,
                    const int32_t* action, float* outs[8], const size_t out_counts[8]) {
    if (!e) return TSP_EINVAL;
    if (!e->finalized) return fail(e, TSP_ESTATE, "network call before tsp_finalize_weights");
    if (B < 1 || !in) return fail(e, TSP_EINVAL, "bad batch");
    CK(e, cudaSetDevice(e->device));
    const tsp_config& c = e->cfg;
    NetParams p;
    memset(&p, 0, sizeof(p));
    p.B = B; p.A = c.num_actions; p.H = c.encoding_size; p.nf = c.n_futures; p.support = c.support_size; p.mode = mode;
    p.W = reinterpret_cast<const float*>(e->arena.p);
    p.sched = sc;
    int rc;
    if ((rc = stage_in(e, memory, in, (size_t)B * in_dim, e->n_in, &p.in))) return rc;
    if ((rc = stage_in(e, memory, action, (size_t)B, e->n_act, &p.action))) return rc;
    float* dev[8];
    for (int i = 0; i < 8; ++i)
        if ((rc = stage_out(e, memory, outs[i], out_counts[i], e->n_o[i], &dev[i]))) return rc;
    p.value = dev[0]; p.reward = dev[1]; p.terminal = dev[2]; p.policy_logits = dev[3]; p.hidden_out = dev[4];
    p.value_logits = dev[5]; p.reward_logits = dev[6]; p.transition_logits = dev[7];
    if ((rc = launch_net_any(e, p))) return rc;
    for (int i = 0; i < 8; ++i)
        if ((rc = copy_out(e, memory, outs[i], dev[i], out_counts[i]))) return rc;
    return TSP_OK;
}

int tsp_initial_inference(tsp_handle e, int32_t memory, int32_t B, const float* observations, float* value,
                          float* policy_logits, float* hidden, float* value_logits) {
    if (!e) return TSP_EINVAL;
    const tsp_config& c = e->cfg;
    const size_t n = (size_t)std::max(B, 0), V = 2 * c.support_size + 1;
    float* outs[8] = {value, nullptr, nullptr, policy_logits, hidden, value_logits, nullptr, nullptr};
    const size_t cnt[8] = {n, 0, 0, n * c.num_actions, n * c.encoding_size, n * V, 0, 0};
    return net_call(e, 0, e->sch_root, memory, B, observations, c.obs_dim, nullptr, outs, cnt);
}

int tsp_recurrent_inference(tsp_handle e, int32_t memory, int32_t B, const float* hidden, const int32_t* action,
                            float* value, float* reward, float* terminal, float* policy_logits, float* next_hidden,
                            float* value_logits, float* reward_logits) {
    if (!e) return TSP_EINVAL;
    if (!e->has_recur) return fail(e, TSP_ESTATE, "recurrent_inference is only defined for the deterministic network");
    if (!action) return fail(e, TSP_EINVAL, "action missing");
    const tsp_config& c = e->cfg;
    const size_t n = (size_t)std::max(B, 0), V = 2 * c.support_size + 1;
    float* outs[8] = {value, reward, terminal, policy_logits, next_hidden, value_logits, reward_logits, nullptr};
    const size_t cnt[8] = {n, n, n, n * c.num_actions, n * c.encoding_size, n * V, n * V, 0};
    return net_call(e, 1, e->sch_recur, memory, B, hidden, c.encoding_size, action, outs, cnt);
}

int tsp_stochastic_dynamics(tsp_handle e, int32_t memory, int32_t B, const float* hidden, const int32_t* action,
                            float* transition_logits, float* next_hidden, float* reward) {
    if (!e) return TSP_EINVAL;
    if (!e->has_trans) return fail(e, TSP_ESTATE, "dynamics with futures is only defined for the stochastic network");
    if (!action) return fail(e, TSP_EINVAL, "action missing");
    const tsp_config& c = e->cfg;
    const size_t n = (size_t)std::max(B, 0);
    float* outs[8] = {nullptr, reward, nullptr, nullptr, next_hidden, nullptr, nullptr, transition_logits};
    const size_t cnt[8] = {0, n * c.n_futures, 0, 0, n * c.n_futures * c.encoding_size, 0, 0, n * c.n_futures};
    return net_call(e, 2, e->sch_trans, memory, B, hidden, c.encoding_size, action, outs, cnt);
}

int tsp_prediction(tsp_handle e, int32_t memory, int32_t B, const float* hidden, float* value, float* policy_logits) {
    if (!e) return TSP_EINVAL;
    const tsp_config& c = e->cfg;
    const size_t n = (size_t)std::max(B, 0);
    float* outs[8] = {value, nullptr, nullptr, policy_logits, nullptr, nullptr, nullptr, nullptr};
    const size_t cnt[8] = {n, 0, 0, n * c.num_actions, 0, 0, 0, 0};
    return net_call(e, 3, e->sch_pred, memory, B, hidden, c.encoding_size, nullptr, outs, cnt);
}

int tsp_stack_observations(tsp_handle e, int32_t memory, int32_t B, int32_t so, int32_t F, int32_t HW, const float* frame_pool,
                           const int32_t* frame_index, const float* action_value, float* out) {
    if (!e) return TSP_EINVAL;
    if (B < 1 || so < 0 || F < 1 || HW < 0 || !frame_pool || !frame_index || !out || (so > 0 && !action_value))
        return fail(e, TSP_EINVAL, "tsp_stack_observations: bad argument");
    if (F * (so + 1) + so * HW != e->cfg.obs_dim)
        return fail(e, TSP_EINVAL, "tsp_stack_observations: frame_size * (so + 1) + so * plane_size = %d, obs_dim is %d",
                    F * (so + 1) + so * HW, e->cfg.obs_dim);
    if (memory != TSP_MEM_DEVICE)
        return fail(e, TSP_EINVAL, "tsp_stack_observations takes DEVICE buffers (the frame pool stays on the device)");
    CK(e, cudaSetDevice(e->device));
    stack_observations_kernel<<<B, 128, 0, e->stream>>>(B, so, F, HW, frame_pool, frame_index, action_value, out);
    CK(e, cudaGetLastError());
    e->stats.kernel_launches++;
    return TSP_OK;
}

int tsp_select_action(tsp_handle e, int32_t memory, int32_t B, const int32_t* visit_counts, const int32_t* legal_actions,
                      const int32_t* num_legal, double temperature, const double* uniform, int32_t* action) {
    if (!e) return TSP_EINVAL;
    const int A = e->cfg.num_actions;
    if (B < 1 || !visit_counts || !action || (legal_actions == nullptr) != (num_legal == nullptr) || temperature < 0)
        return fail(e, TSP_EINVAL, "tsp_select_action: bad argument");
    if (temperature != 0.0 && !uniform) return fail(e, TSP_EINVAL, "tsp_select_action: uniform samples required for temperature > 0");
    CK(e, cudaSetDevice(e->device));
    const size_t nb = (size_t)B;
    const int32_t *d_v, *d_l, *d_n;
    const double* d_u;
    int32_t* d_a;
    int rc;
    if ((rc = stage_in(e, memory, visit_counts, nb * A, e->s_visits, &d_v))) return rc;
    if ((rc = stage_in(e, memory, legal_actions, nb * A, e->s_legal, &d_l))) return rc;
    if ((rc = stage_in(e, memory, num_legal, nb, e->s_nlegal, &d_n))) return rc;
    if ((rc = stage_in(e, memory, uniform, nb, e->s_noise, &d_u))) return rc;
    if ((rc = stage_out(e, memory, action, nb, e->s_depth, &d_a))) return rc;
    select_action_kernel<<<(B + 127) / 128, 128, 0, e->stream>>>(B, A, d_v, d_l, d_n, temperature, d_u, d_a);
    CK(e, cudaGetLastError());
    e->stats.kernel_launches++;
    return copy_out(e, memory, action, d_a, nb);
}

int tsp_search_statistics(tsp_handle e, int32_t memory, int32_t B, const int32_t* visit_counts, double* child_visits) {
    if (!e) return TSP_EINVAL;
    const int A = e->cfg.num_actions;
    if (B < 1 || !visit_counts || !child_visits) return fail(e, TSP_EINVAL, "tsp_search_statistics: bad argument");
    CK(e, cudaSetDevice(e->device));
    const size_t nb = (size_t)B;
    const int32_t* d_v;
    double* d_c;
    int rc;
    if ((rc = stage_in(e, memory, visit_counts, nb * A, e->s_visits, &d_v))) return rc;
    if ((rc = stage_out(e, memory, child_visits, nb * A, e->s_cval, &d_c))) return rc;
    search_statistics_kernel<<<(B + 127) / 128, 128, 0, e->stream>>>(B, A, d_v, d_c);
    CK(e, cudaGetLastError());
    e->stats.kernel_launches++;
    return copy_out(e, memory, child_visits, d_c, nb * A);
}

int tsp_peer_put(tsp_handle e, const void* src, int64_t num_bytes, void* const* peers, int32_t n_peers) {
    if (!e) return TSP_EINVAL;
    if (!src || !peers || n_peers < 1 || n_peers > 16 || num_bytes < 16 || num_bytes % 16 || ((uintptr_t)src & 15))
        return fail(e, TSP_EINVAL, "tsp_peer_put: 1..16 peers, a 16-byte aligned source, a multiple of 16 bytes");
    CK(e, cudaSetDevice(e->device));
    PeerPtrs pp;
    memset(&pp, 0, sizeof(pp));
    for (int r = 0; r < n_peers; ++r) {
        if (!peers[r] || ((uintptr_t)peers[r] & 15)) return fail(e, TSP_EINVAL, "tsp_peer_put: peer %d: null or misaligned", r);
        pp.p[r] = reinterpret_cast<unsigned char*>(peers[r]);
    }
    const long long n16 = num_bytes / 16;
    peer_put_kernel<<<(unsigned)((n16 + 255) / 256), 256, 0, e->stream>>>(reinterpret_cast<const uint4*>(src), n16, pp, n_peers);
    CK(e, cudaGetLastError());
    e->stats.kernel_launches++;
    return TSP_OK;
}

// tcgen05 path: a bounded mbarrier wait timed out (an MMA commit never arrived) -- the results are invalid
int tsp_export_tree(tsp_handle e, int32_t root_index, tsp_tree_info* info, void* blocks, int32_t max_blocks, float* hidden,
                    int32_t max_hidden) {
    if (!e || !info) return TSP_EINVAL;
    if (root_index < 0 || root_index >= e->tree_roots)
        return fail(e, TSP_EINVAL, "tsp_export_tree: root %d is not one of the %d trees of the latest search", root_index, e->tree_roots);
    CK(e, cudaSetDevice(e->device));
    CK(e, cudaStreamSynchronize(e->stream));
    const size_t bb = e->nc == 5 ? sizeof(NodeBlock<5>) : sizeof(NodeBlock<8>);
    TreeHeader h;
    CK(e, cudaMemcpy(&h, reinterpret_cast<const TreeHeader*>(e->hdr.p) + root_index, sizeof(h), cudaMemcpyDeviceToHost));
    memset(info, 0, sizeof(*info));
    info->n_blocks = h.n_blocks;
    info->block_bytes = (int32_t)bb;
    info->num_children = e->nc;
    info->n_root = h.n_root;
    info->root_visit = h.root_visit;
    info->root_value_sum = h.root_vsum;
    // hidden-state slots: block id == slot (deterministic); one slot per future of every expanded transition + the root
    // (a fed-network search keeps none: 0)
    info->n_hidden = e->tree_fed ? 0 : (e->cfg.variant == 0 ? h.n_blocks : 1 + h.n_tgroups * e->cfg.n_futures);
    CK(e, cudaMemcpy(info->root_prior, reinterpret_cast<const double*>(e->rootprior.p) + (size_t)root_index * GW, GW * sizeof(double),
                     cudaMemcpyDeviceToHost));
    CK(e, cudaMemcpy(info->root_action, reinterpret_cast<const int*>(e->rootact.p) + (size_t)root_index * GW, GW * sizeof(int),
                     cudaMemcpyDeviceToHost));
    if (blocks) {
        if (max_blocks < h.n_blocks) return fail(e, TSP_EINVAL, "tsp_export_tree: %d blocks, room for %d", h.n_blocks, max_blocks);
        CK(e, cudaMemcpy(blocks, reinterpret_cast<const unsigned char*>(e->blocks.p) + (size_t)root_index * e->blocks_per_tree * bb,
                         (size_t)h.n_blocks * bb, cudaMemcpyDeviceToHost));
    }
    if (hidden && info->n_hidden > 0) {
        if (max_hidden < info->n_hidden) return fail(e, TSP_EINVAL, "tsp_export_tree: %d hidden states, room for %d", info->n_hidden, max_hidden);
        const size_t H = (size_t)e->cfg.encoding_size;
        CK(e, cudaMemcpy(hidden, reinterpret_cast<const float*>(e->hidden.p) + (size_t)root_index * e->hidden_per_tree * H,
                         (size_t)info->n_hidden * H * sizeof(float), cudaMemcpyDeviceToHost));
    }
    return TSP_OK;
}

static int check_tc_status(tsp_engine* e) {
    if ((!e->has_tc && !e->has_tcs && !e->has_rtcs) || !e->tc_status.p) return TSP_OK;
    int st = 0;
    CK(e, cudaMemcpy(&st, e->tc_status.p, sizeof(int), cudaMemcpyDeviceToHost));
    if (st) {  // report once: the flag is cleared so that later searches are judged on their own
        cudaMemset(e->tc_status.p, 0, sizeof(int));
        if (st == 2)
            return fail(e, TSP_ECUDA, "tcgen05 kernels with streamed weights: an observation or activation left the fp16-split range (|x| > 6e4); "
                                      "the results of that call are invalid -- set TSP_NO_TCS=1 / TSP_NO_RTCS=1 to run this network on the mma.sync kernels");
        return fail(e, TSP_ECUDA, "tcgen05 search kernel: mbarrier wait timed out (the results of that search are invalid)");
    }
    return TSP_OK;
}

int tsp_synchronize(tsp_handle e) {
    if (!e) return TSP_EINVAL;
    CK(e, cudaSetDevice(e->device));
    CK(e, cudaStreamSynchronize(e->stream));
    return check_tc_status(e);
}

int tsp_get_stats(tsp_handle e, tsp_stats* out) {
    if (!e || !out) return TSP_EINVAL;
    CK(e, cudaSetDevice(e->device));
    int rc = drain_events(e);
    if (rc) return rc;
    if ((rc = check_tc_status(e))) return rc;
    e->stats.pool_bytes = e->pool_bytes;
    *out = e->stats;
    return TSP_OK;
}

int tsp_reset_timing(tsp_handle e) {
    if (!e) return TSP_EINVAL;
    CK(e, cudaSetDevice(e->device));
    int rc = drain_events(e);
    if (rc) return rc;
    e->stats.sum_search_ms = 0.0;
    e->stats.sum_root_ms = 0.0;
    e->stats.timed_runs = 0;
    return TSP_OK;
}

int tsp_debug_phase_cycles(tsp_handle e, uint64_t out[64], int32_t reset) {
    if (!e || !out) return TSP_EINVAL;
    for (int i = 0; i < 64; ++i) out[i] = 0;
    if (!e->phase.p) return TSP_OK;
    CK(e, cudaSetDevice(e->device));
    CK(e, cudaStreamSynchronize(e->stream));
    CK(e, cudaMemcpy(out, e->phase.p, 64 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    if (reset) CK(e, cudaMemset(e->phase.p, 0, 64 * sizeof(uint64_t)));
    return TSP_OK;
}

int tsp_stream(tsp_handle e, void** stream) {
    if (!e || !stream) return TSP_EINVAL;
    *stream = reinterpret_cast<void*>(e->stream);
    return TSP_OK;
}

}  // extern "C"
