/*
 * tsp.h -- C ABI of the B200 batched MuZero tree-search engine (libtsp_b200.so).
 *
 * The reference (Disiok/tree-search-planning) is pure Python and has no FFI of its own; this
 * header is the boundary a maintainer binds (ctypes stub in INTEGRATION.md) to replace the
 * planning hot path. Every entry point cites the reference interface it replaces
 * (paths under /root/reference/muzero-general/).
 *
 * Conventions
 *   - plain C: pointers + sizes, no C++/torch types; every function returns 0 on success or a
 *     negative TSP_E* code, with a message retrievable through tsp_last_error().
 *   - ownership: the caller owns every buffer it passes; the engine owns its node pools, hidden
 *     state pools, weight arena and staging buffers, sized at tsp_create (max_roots, max_simulations).
 *   - memory spaces: each call states whether its I/O pointers are HOST or DEVICE memory. HOST
 *     buffers are staged through engine-owned device buffers with cudaMemcpyAsync on the engine's
 *     stream (pinned host memory makes these truly asynchronous). All calls are asynchronous on
 *     that stream; tsp_synchronize() waits. One host thread per handle.
 *   - there is NO CPU fallback: without a CUDA device tsp_create fails with TSP_ENODEVICE.
 */
#ifndef TSP_H
#define TSP_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TSP_ABI_VERSION 1
#define TSP_REC_WIDTH 16     /* floats per network-evaluation record (see below) */
#define TSP_MAX_ACTIONS 8    /* |action_space| supported by the node-block layouts */
#define TSP_MAX_LAYERS 8     /* Linear layers per MLP */

typedef struct tsp_engine *tsp_handle;

/* error codes */
enum {
    TSP_OK = 0,
    TSP_EINVAL = -1,      /* bad argument (mirrors the reference's asserts, self_play.py:354-359) */
    TSP_ENODEVICE = -2,   /* no CUDA device / wrong architecture */
    TSP_ECUDA = -3,       /* CUDA runtime error, text in tsp_last_error */
    TSP_ENOMEM = -4,
    TSP_ENOTIMPL = -5,    /* > 2 players (self_play.py:503 raises NotImplementedError) */
    TSP_ESTATE = -6       /* weights not loaded / not finalized */
};

/* which MCTS the engine runs; same dispatch as muzero.py:232-250 */
enum {
    TSP_VARIANT_MUZERO = 0,         /* self_play.py:303-503 */
    TSP_VARIANT_STOCHASTIC = 1,     /* self_play_stochastic.py:259-557 (chance nodes) */
    TSP_VARIANT_RISK_SENSITIVE = 2  /* self_play_risk_sensitive.py:259-580 */
};

/* sub-MLPs of models.MuZeroFullyConnectedNetwork / MuZeroStochasticConcat (models.py:156-195, 502-513) */
enum {
    TSP_MLP_REPRESENTATION = 0, /* representation_network */
    TSP_MLP_DYNAMICS = 1,       /* dynamics_encoded_state_network, input [h | onehot(a) (| onehot(z))] */
    TSP_MLP_REWARD = 2,         /* dynamics_reward_network */
    TSP_MLP_TERMINAL = 3,       /* dynamics_terminal_network (only evaluated with mask_absorbing_states) */
    TSP_MLP_POLICY = 4,         /* prediction_policy_network */
    TSP_MLP_VALUE = 5,          /* prediction_value_network */
    TSP_MLP_PRIOR = 6,          /* transition_prior_network (stochastic variants) */
    /* models.MuZeroStochastic (models.py:298-462): dynamics_encoded_state_network is a ModuleList with ONE MLP PER FUTURE,
     * each on [h | onehot(a)]. Future 0 is TSP_MLP_DYNAMICS, futures 1..3 are the ids below; setting any of them selects
     * the separate-heads network (n_futures of them must then be set, with equal shapes). */
    TSP_MLP_DYNAMICS_F1 = 7,
    TSP_MLP_DYNAMICS_F2 = 8,
    TSP_MLP_DYNAMICS_F3 = 9,
    TSP_NUM_MLP = 10
};

enum { TSP_MEM_HOST = 0, TSP_MEM_DEVICE = 1 };

/* The attributes of the reference's MuZeroConfig that the hot path reads ("Config"). */
typedef struct {
    int32_t variant;               /* TSP_VARIANT_* */
    int32_t num_actions;           /* len(config.action_space), actions are 0..A-1 */
    int32_t obs_dim;               /* flattened stacked observation, models.py:157-163 */
    int32_t encoding_size;         /* config.encoding_size */
    int32_t support_size;          /* config.support_size */
    int32_t n_futures;             /* config.n_futures (stochastic variants), else 1 */
    int32_t mask_absorbing_states; /* config.mask_absorbing_states, self_play.py:415 */
    int32_t num_players;           /* len(config.players): 1, or 2 for the deterministic search (self_play.py:469-472,
                                      492-500); > 2 -> TSP_ENOTIMPL like the reference's NotImplementedError */
    double discount;               /* config.discount */
    double pb_c_base;              /* config.pb_c_base */
    double pb_c_init;              /* config.pb_c_init */
    double root_exploration_fraction; /* config.root_exploration_fraction */
    int32_t max_roots;             /* capacity: concurrent roots per tsp_run call */
    int32_t max_simulations;       /* capacity: config.num_simulations upper bound */
    int32_t device;                /* CUDA device ordinal */
    int32_t reserved;
} tsp_config;

/* Network-evaluation record, float32[TSP_REC_WIDTH], one per network call in call order per root.
 *   state record      (recurrent_inference / prediction):  [0]=1 [1]=value [2]=reward [3]=terminal logit
 *                                                           [4..4+A) priors = softmax(policy logits)
 *   transition record (MuZeroStochasticConcat.dynamics):   [0]=2 [1..1+nf) softmax(prior logits)
 *                                                           [1+nf..1+2nf) decoded rewards
 * `trace_*` outputs receive the engine's own network outputs; `fed_*` inputs replace the network
 * (tree arithmetic with network outputs fed identically -- the parity mode of BASELINE.json). */

typedef struct {
    int32_t num_roots;             /* B <= max_roots */
    int32_t num_simulations;       /* S <= max_simulations; 0 == policy_only (self_play.py:381) */
    int32_t memory;                /* TSP_MEM_HOST / TSP_MEM_DEVICE for every pointer below */
    int32_t add_exploration_noise; /* MCTS.run argument, self_play.py:372 */
    int32_t uniform_policy;        /* MCTS.run argument, self_play.py:361,411 */
    int32_t max_records;           /* rows per root in fed_records / trace_records */
    int32_t tie_len;               /* T */
    int32_t chance_len;            /* K */
    int32_t resume;                /* != 0: MCTS.run(..., override_root_with=root) (self_play.py:331-333, 372-378) for every
                                      root of the previous tsp_run on this handle: the trees are kept, no root inference,
                                      a fresh MinMaxStats / max_tree_depth, the exploration noise mixed into the root
                                      priors AGAIN, then num_simulations more simulations (the total must stay within
                                      max_simulations). observations are ignored, root_predicted_value is not written
                                      (the reference returns None). Deterministic and risk-sensitive search only: the
                                      stochastic reference fails its own assert (self_play_stochastic.py:333). */
    int32_t outputs_on_device;     /* != 0 with memory == TSP_MEM_HOST: the INPUT arrays are host memory (uploaded by the
                                      engine, overlapped with the previous search) while every OUTPUT array is device
                                      memory -- the multi-GPU path gathers the root statistics over NCCL before they
                                      leave the device (sharding.ShardedEngine). Ignored with TSP_MEM_DEVICE. */
    /* inputs */
    const float *observations;     /* [B][obs_dim]; torch.tensor(observation).float() of self_play.py:336-341 */
    const int32_t *legal_actions;  /* [B][A] ordered legal-action lists (dict insertion order matters,
                                      self_play.py:364-370), or NULL = config.action_space */
    const int32_t *num_legal;      /* [B], or NULL with legal_actions == NULL */
    const double *noise;           /* [B][A] Dirichlet sample aligned with the legal list (self_play.py:546);
                                      required when add_exploration_noise */
    const uint8_t *tie_stream;     /* [B][T] or NULL (= lowest index); replaces numpy.random.choice of
                                      self_play.py:444-450: tied[stream[k % T] % n_tied], k++ per real tie */
    const uint8_t *chance_stream;  /* [B][K]; replaces numpy.random.choice of self_play_stochastic.py:530-532:
                                      z = stream[k % K] % nf, k++ per chance-node traversal (nf > 1) */
    const float *fed_root;         /* [B][REC] or NULL */
    const float *fed_records;      /* [B][max_records][REC] or NULL */
    /* outputs (any may be NULL) -- what consumers read from `root` / `extra_info` */
    int32_t *visit_counts;         /* [B][A]  root.children[a].visit_count (0 for illegal a) */
    double *root_value;            /* [B]     root.value() (attribute .value in the risk-sensitive variant) */
    double *child_value;           /* [B][A]  root.children[a].value() */
    double *child_prior;           /* [B][A]  root.children[a].prior (after noise) */
    double *child_reward;          /* [B][A]  root.children[a].reward */
    int32_t *max_tree_depth;       /* [B]     extra_info["max_tree_depth"] */
    float *root_predicted_value;   /* [B]     extra_info["root_predicted_value"] */
    float *root_hidden;            /* [B][H]  root.hidden_state */
    int32_t *num_records;          /* [B]     network evaluations made per root (excluding the root one) */
    int32_t *total_select_depth;   /* [B]     sum over all descents of the number of levels walked (for the
                                              algorithmic-traffic figure of the roofline report) */
    float *trace_root;             /* [B][REC] */
    float *trace_records;          /* [B][max_records][REC] */
} tsp_run_args;

typedef struct {
    int64_t kernel_launches;       /* kernels launched by this handle so far */
    int64_t searches;              /* tsp_run calls */
    int64_t bytes_h2d, bytes_d2h;  /* staged through the engine so far */
    int64_t pool_bytes;            /* device memory owned by the engine */
    float last_search_ms;          /* device time of the search kernel of the last tsp_run (CUDA events) */
    float last_root_ms;            /* device time of the root kernel of the last tsp_run */
    double sum_search_ms;          /* accumulated over every tsp_run since tsp_reset_timing (CUDA events on the */
    double sum_root_ms;            /*   engine's stream, one event triple per run)                              */
    int64_t timed_runs;
    int32_t sm_count;
    int32_t search_ctas;           /* grid of the last search kernel */
    int32_t search_smem_bytes;
    int32_t trees_per_cta;
    int32_t threads_per_cta;
    int32_t search_kernel;         /* last search: 3 = tsp::search_tcs_kernel (tcgen05, weights streamed by TMA), 2 = tsp::search_tc_kernel (tcgen05, weights in tensor memory), 1 = tsp::search_fast_kernel (weights in shared memory), 0 = tsp::search_kernel */
    int32_t root_kernel;           /* last initial inference: 1 = tsp::root_tcs_kernel (tcgen05, weights streamed by TMA), 0 = tsp::net_kernel */
} tsp_stats;

int tsp_abi_version(void);

/* replaces MCTS(config) construction (self_play.py:311-312) plus the model's placement on a device */
int tsp_create(const tsp_config *cfg, tsp_handle *out);
int tsp_destroy(tsp_handle h);
/* message of the last error on this handle (h == NULL: last tsp_create failure) */
const char *tsp_last_error(tsp_handle h);

/* replaces model.set_weights(state_dict) (models.py:128-129): one call per torch.nn.Linear,
 * W is [out_dim][in_dim] row-major float32 HOST memory exactly as in the state dict
 * ("<net>.module.<2*layer>.weight"), b is [out_dim]. Then tsp_finalize_weights() packs and uploads. */
int tsp_set_layer(tsp_handle h, int32_t mlp, int32_t layer, const float *W, const float *b,
                  int32_t out_dim, int32_t in_dim);
int tsp_finalize_weights(tsp_handle h);
/* packed device-resident weight arena (for an NCCL broadcast / hot refresh between ranks) */
int tsp_weight_arena(tsp_handle h, void **device_ptr, int64_t *num_bytes);

/* replaces MCTS.run (self_play.py:314-434; self_play_stochastic.py:270-400;
 * self_play_risk_sensitive.py:270-399) for B independent roots at once */
int tsp_run(tsp_handle h, const tsp_run_args *args);

/* replaces model.initial_inference + support_to_scalar (models.py:259-281, 1180-1201) for B rows.
 * outputs (any may be NULL): value[B], policy_logits[B][A], hidden[B][H], value_logits[B][2s+1] */
int tsp_initial_inference(tsp_handle h, int32_t memory, int32_t B, const float *observations,
                          float *value, float *policy_logits, float *hidden, float *value_logits);
/* replaces model.recurrent_inference + support_to_scalar (models.py:283-287) for B rows
 * (deterministic network). outputs (any may be NULL): value[B], reward[B], terminal[B],
 * policy_logits[B][A], next_hidden[B][H], value_logits[B][2s+1], reward_logits[B][2s+1] */
int tsp_recurrent_inference(tsp_handle h, int32_t memory, int32_t B, const float *hidden, const int32_t *action,
                            float *value, float *reward, float *terminal, float *policy_logits,
                            float *next_hidden, float *value_logits, float *reward_logits);
/* replaces MuZeroStochasticConcat.dynamics (models.py:520-550) for B rows:
 * transition_logits[B][nf], next_hidden[B][nf][H], reward[B][nf] (decoded) */
int tsp_stochastic_dynamics(tsp_handle h, int32_t memory, int32_t B, const float *hidden, const int32_t *action,
                            float *transition_logits, float *next_hidden, float *reward);
/* replaces model.prediction + support_to_scalar (models.py:209-212): value[B], policy_logits[B][A] */
int tsp_prediction(tsp_handle h, int32_t memory, int32_t B, const float *hidden, float *value,
                   float *policy_logits);

/* ---- the data formats on either side of the search, on the device (batched self-play) ---- */

/* replaces GameHistory.get_stacked_observations (self_play.py:587-624) for B rows at once.
 *   frame_pool   [N][frame_size] float32 observation frames (frame_size = C*H*W of config.observation_shape)
 *   frame_index  [B][so + 1] rows of frame_pool: the frame at `index`, then index-1, index-2, ...; -1 = before the
 *                start of the game (the reference pads with zeros)
 *   action_value [B][so] value of the action plane that follows past frame i (action_history[past + 1]; 0 when padded)
 *   out          [B][obs_dim], obs_dim = frame_size*(so+1) + so*plane_size (plane_size = H*W) -- the layout
 *                tsp_run / tsp_initial_inference take as `observations` */
int tsp_stack_observations(tsp_handle h, int32_t memory, int32_t B, int32_t stacked_observations, int32_t frame_size,
                           int32_t plane_size, const float *frame_pool, const int32_t *frame_index,
                           const float *action_value, float *out);
/* replaces SelfPlay.select_action (self_play.py:277-299) for B roots: temperature 0 = first arg-max of the visit
 * counts over the ordered legal list, +inf = uniform, else numpy.random.choice(actions, p = visits^(1/T) / sum)
 * with the caller's uniform samples u[B] in [0, 1) (inverse CDF, exactly numpy's algorithm).
 * legal_actions / num_legal as in tsp_run (NULL = all actions). */
int tsp_select_action(tsp_handle h, int32_t memory, int32_t B, const int32_t *visit_counts, const int32_t *legal_actions,
                      const int32_t *num_legal, double temperature, const double *uniform, int32_t *action);
/* replaces the policy target of GameHistory.store_search_statistics (self_play.py:570-585):
 * child_visits[B][A] = visit_count / max(sum of the root's visit counts, 1) in float64 */
int tsp_search_statistics(tsp_handle h, int32_t memory, int32_t B, const int32_t *visit_counts, double *child_visits);

/* The device-resident tree of one root of the latest tsp_run, for consumers that walk below the root's children
 * (diagnose_model.py:156-185 recurses through node.children; SelfPlay itself only reads depth 1, which tsp_run returns).
 * One fixed-size record per EXPANDED node ("block", id 0 = the root, ids in expansion order) holding the statistics of
 * that node's children, i.e. exactly the fields of the reference's Node (self_play.py:506-522):
 *   int32 visit[NC] | int32 child[NC] (block id of the expanded child, -1: not expanded) | float32 prior[NC] |
 *   float32 reward[NC] | int32 parent | int32 aux (bits 0..7: slot in the parent; bits 8..: hidden-state slot of a state
 *   node) | float64 value_sum[NC] (risk-sensitive: value),  NC = 5 (block_bytes 128) or 8 (256).
 * Chance-node variants: node kinds alternate by depth (root = StateNode, its children = TransitionNodes whose slots
 * are the futures z, ...). The root's own float64 priors and the action of each root slot come back in `info`.
 * blocks: HOST [max_blocks][block_bytes] or NULL (query sizes only); hidden: HOST [max_hidden][H] hidden states by
 * slot, or NULL (a search on fed network outputs keeps none: n_hidden == 0). Synchronises the engine's stream. */
typedef struct tsp_tree_info {
    int32_t n_blocks;        /* expanded nodes of this tree */
    int32_t block_bytes;
    int32_t num_children;    /* NC */
    int32_t n_root;          /* children of the root (legal actions) */
    int32_t root_visit;
    int32_t n_hidden;        /* hidden-state slots in use */
    double root_value_sum;   /* risk-sensitive: root.value */
    double root_prior[8];
    int32_t root_action[8];
} tsp_tree_info;
int tsp_export_tree(tsp_handle h, int32_t root_index, tsp_tree_info *info, void *blocks, int32_t max_blocks, float *hidden,
                    int32_t max_hidden);

/* Multi-GPU: the all-gather of a rank's packed root statistics ("NCCL ... to gather root visit distributions",
 * BASELINE.json north_star) as ONE kernel of peer stores over NVLink / NVSwitch: `src` (DEVICE, num_bytes, 16-byte
 * granular) is written to peers[0..n_peers) -- device pointers into the other ranks' (and this rank's own) symmetric
 * buffers at this rank's offset, obtained by the caller (torch symmetric memory / cudaIpc). Asynchronous on the
 * engine's stream; the cross-GPU barrier that publishes the writes is the caller's. */
int tsp_peer_put(tsp_handle h, const void *src, int64_t num_bytes, void *const *peers, int32_t n_peers);

int tsp_synchronize(tsp_handle h);
int tsp_get_stats(tsp_handle h, tsp_stats *out);
/* zero sum_search_ms / sum_root_ms / timed_runs (synchronizes the stream) */
int tsp_reset_timing(tsp_handle h);
/* developer instrumentation (env TSP_PHASE_TIMING=1 at tsp_create): summed clock64() deltas of thread 0 of
 * every CTA of the deterministic search kernel: [0] select+gather [1] barrier wait [2] MLP [3] expand+backup
 * [4] number of samples; [8 + 2s] own work in MLP stage s, [9 + 2s] wait at the barrier of stage s */
int tsp_debug_phase_cycles(tsp_handle h, uint64_t out[64], int32_t reset);
/* the CUDA stream the engine launches on (cudaStream_t as void*), for event timing by the caller */
int tsp_stream(tsp_handle h, void **stream);

#ifdef __cplusplus
}
#endif
#endif /* TSP_H */
