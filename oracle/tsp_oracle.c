/*
 * TEST INFRASTRUCTURE ONLY -- CPU restatement ("oracle") of the reference planning hot path.
 *
 * Nothing under tree_search_planning_b200/ links, loads or calls this file. Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do, and only
 * as the checker or the reported CPU baseline. The product path has no CPU fallback.
 *
 * What is restated (reference file:line, all under muzero-general/):
 *   MCTS.run / select_child / ucb_score / backpropagate      self_play.py:314-503
 *   Node.expand / add_exploration_noise / value              self_play.py:506-549
 *   MinMaxStats                                              self_play.py:661-678
 *   stochastic MCTS (StateNode / TransitionNode)             self_play_stochastic.py:270-557
 *   risk-sensitive backup / update_value / ucb               self_play_risk_sensitive.py:401-537
 *   MuZeroFullyConnectedNetwork representation / dynamics /
 *     prediction / initial_inference / recurrent_inference   models.py:209-287
 *   MuZeroStochasticConcat dynamics / _dynamics              models.py:520-595
 *   MuZeroStochastic (separate heads) dynamics / _dynamics    models.py:365-428
 *   mlp (Linear + ELU, identity output), support_to_scalar   models.py:1150-1162, 1180-1201
 *
 * Parity pinning: the reference has no tests or golden vectors for this path, so
 * this restatement is pinned against the LIVE reference imported in the build container
 * (oracle/reference_live.py) and against the fixtures that script wrote to tests/golden/
 * (tests/test_oracle_golden.py): tree arithmetic bit-exact when network outputs are fed
 * identically, network arithmetic within 1e-5 of torch fp32.
 *
 * Arithmetic: tree statistics in IEEE double exactly in the reference's operation order
 * (compile with -ffp-contract=off); network in float32.
 *
 * Randomness is external (same semantics as reference_live.ExternalStreams):
 *   noise[B][A]      Dirichlet sample per root, aligned with the legal-action list
 *   tie[B][T]        consumed only when >1 children share the maximum score:
 *                    pick tied[ tie[b][k % T] % n_tied ], k++   (NULL: lowest index)
 *   chance[B][K]     consumed once per TransitionNode traversal with nf > 1:
 *                    z = chance[b][k % K] % nf, k++             (NULL: 0)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

#define ORC_MAX_LAYERS 8
#define ORC_MAX_ACTIONS 32
#define ORC_REC_WIDTH 16

/* ORC_DYN_F1.. : dynamics_encoded_state_network[z], z >= 1, of MuZeroStochastic (separate heads per future,
 * models.py:334-345); z == 0 is ORC_DYN */
enum { ORC_REP = 0, ORC_DYN, ORC_REWARD, ORC_TERMINAL, ORC_POLICY, ORC_VALUE, ORC_PRIOR, ORC_DYN_F1, ORC_DYN_F2, ORC_DYN_F3,
       ORC_NUM_MLP };

typedef struct {
    int n_layers;
    int in_dim[ORC_MAX_LAYERS];
    int out_dim[ORC_MAX_LAYERS];
    float *W[ORC_MAX_LAYERS]; /* [out][in], torch Linear layout */
    float *b[ORC_MAX_LAYERS];
} orc_mlp;

typedef struct {
    int variant;       /* 0 deterministic, 1 stochastic, 2 risk-sensitive */
    int num_actions;   /* A */
    int obs_dim;
    int encoding_size; /* H */
    int support_size;  /* s; bins = 2s+1 */
    int n_futures;     /* nf (variants 1,2) */
    int mask_absorbing_states;
    int num_players;   /* 1 or 2 (deterministic MCTS only; the chance-node variants assert 1) */
    double discount;
    double pb_c_base;
    double pb_c_init;
    double root_exploration_fraction;
} orc_config;

typedef struct {
    orc_config cfg;
    orc_mlp mlp[ORC_NUM_MLP];
    int max_width;
} orc_net;

/* ------------------------------------------------------------------ network ---- */

orc_net *orc_net_create(const orc_config *cfg) {
    orc_net *n = (orc_net *)calloc(1, sizeof(orc_net));
    n->cfg = *cfg;
    n->max_width = cfg->obs_dim > cfg->encoding_size ? cfg->obs_dim : cfg->encoding_size;
    n->max_width += ORC_MAX_ACTIONS + 8;
    return n;
}

void orc_net_destroy(orc_net *n) {
    if (!n) return;
    for (int m = 0; m < ORC_NUM_MLP; ++m)
        for (int l = 0; l < n->mlp[m].n_layers; ++l) {
            free(n->mlp[m].W[l]);
            free(n->mlp[m].b[l]);
        }
    free(n);
}

int orc_net_set_layer(orc_net *n, int mlp_id, int layer, const float *W, const float *b, int out_dim, int in_dim) {
    if (mlp_id < 0 || mlp_id >= ORC_NUM_MLP || layer < 0 || layer >= ORC_MAX_LAYERS) return -1;
    orc_mlp *m = &n->mlp[mlp_id];
    if (layer >= m->n_layers) m->n_layers = layer + 1;
    free(m->W[layer]);
    free(m->b[layer]);
    m->W[layer] = (float *)malloc(sizeof(float) * (size_t)out_dim * in_dim);
    m->b[layer] = (float *)malloc(sizeof(float) * (size_t)out_dim);
    memcpy(m->W[layer], W, sizeof(float) * (size_t)out_dim * in_dim);
    memcpy(m->b[layer], b, sizeof(float) * (size_t)out_dim);
    m->in_dim[layer] = in_dim;
    m->out_dim[layer] = out_dim;
    if (out_dim > n->max_width) n->max_width = out_dim;
    if (in_dim > n->max_width) n->max_width = in_dim;
    return 0;
}

/* models.py:1150-1162: Linear, ELU between layers, Identity after the last one. */
static void mlp_forward(const orc_mlp *m, const float *x, float *y, float *tmp0, float *tmp1) {
    const float *in = x;
    float *bufs[2] = {tmp0, tmp1};
    for (int l = 0; l < m->n_layers; ++l) {
        const int I = m->in_dim[l], O = m->out_dim[l];
        const int last = (l == m->n_layers - 1);
        float *out = last ? y : bufs[l & 1];
        const float *W = m->W[l];
        for (int o = 0; o < O; ++o) {
            float acc = m->b[l][o];
            const float *w = W + (size_t)o * I;
            for (int i = 0; i < I; ++i) acc += in[i] * w[i];
            if (!last) acc = acc > 0.0f ? acc : expm1f(acc); /* ELU(alpha=1) */
            out[o] = acc;
        }
        in = out;
    }
}

/* models.py:223-231 / 248-255: per-row min-max scaling, scale += 1e-5 only where scale < 1e-5 */
static void minmax_scale(const float *x, float *y, int n) {
    float mn = x[0], mx = x[0];
    for (int i = 1; i < n; ++i) {
        if (x[i] < mn) mn = x[i];
        if (x[i] > mx) mx = x[i];
    }
    float scale = mx - mn;
    if (scale < 1e-5f) scale += 1e-5f;
    for (int i = 0; i < n; ++i) y[i] = (x[i] - mn) / scale;
}

void orc_softmax(const float *x, float *p, int n) {
    float m = x[0];
    for (int i = 1; i < n; ++i)
        if (x[i] > m) m = x[i];
    float s = 0.0f;
    for (int i = 0; i < n; ++i) {
        p[i] = expf(x[i] - m);
        s += p[i];
    }
    for (int i = 0; i < n; ++i) p[i] = p[i] / s;
}

/* models.py:1192-1200: x -> sign(x) * (((sqrt(1 + 4*eps*(|x| + 1 + eps)) - 1) / (2*eps))^2 - 1),
 * evaluated in float32 in torch's operation order (python scalars 4*0.001 and 2*0.001 are
 * folded in double first, then rounded to float32 when they meet the tensor). */
float orc_inverse_value_transform(float x) {
    const float c4e = (float)(4 * 0.001);
    const float c2e = (float)(2 * 0.001);
    float t = fabsf(x) + 1.0f;
    t = t + 0.001f;
    t = c4e * t;
    t = 1.0f + t;
    t = sqrtf(t);
    t = t - 1.0f;
    t = t / c2e;
    t = t * t;
    t = t - 1.0f;
    float sgn = (x > 0.0f) ? 1.0f : ((x < 0.0f) ? -1.0f : 0.0f);
    return sgn * t;
}

/* models.py:1180-1201 */
float orc_support_to_scalar(const float *logits, int support_size) {
    float p[2 * 64 + 1];
    const int n = 2 * support_size + 1;
    orc_softmax(logits, p, n);
    float x = 0.0f;
    for (int i = 0; i < n; ++i) x += (float)(i - support_size) * p[i];
    return orc_inverse_value_transform(x);
}

typedef struct {
    float *a, *b, *c; /* scratch, each max_width */
} orc_scratch;

static orc_scratch scratch_new(const orc_net *n) {
    orc_scratch s;
    s.a = (float *)malloc(sizeof(float) * n->max_width);
    s.b = (float *)malloc(sizeof(float) * n->max_width);
    s.c = (float *)malloc(sizeof(float) * n->max_width);
    return s;
}
static void scratch_free(orc_scratch *s) {
    free(s->a);
    free(s->b);
    free(s->c);
}

/* models.py:209-212 + support_to_scalar; value_logits may be NULL */
static void net_prediction(const orc_net *n, const float *h, float *policy_logits, float *value_scalar,
                           float *value_logits, orc_scratch *s) {
    float vl[2 * 64 + 1];
    mlp_forward(&n->mlp[ORC_POLICY], h, policy_logits, s->a, s->b);
    mlp_forward(&n->mlp[ORC_VALUE], h, vl, s->a, s->b);
    *value_scalar = orc_support_to_scalar(vl, n->cfg.support_size);
    if (value_logits) memcpy(value_logits, vl, sizeof(float) * (2 * n->cfg.support_size + 1));
}

/* models.py:219-231, 259-281 (reconstruction head is discarded by MCTS and not computed) */
static void net_initial(const orc_net *n, const float *obs, float *hidden, float *policy_logits,
                        float *value_scalar, float *value_logits, orc_scratch *s) {
    mlp_forward(&n->mlp[ORC_REP], obs, s->c, s->a, s->b);
    minmax_scale(s->c, hidden, n->cfg.encoding_size);
    net_prediction(n, hidden, policy_logits, value_scalar, value_logits, s);
}

/* models.py:233-257 (deterministic) and models.py:568-595 (stochastic concat, future z >= 0):
 * x = [h | onehot(a) | onehot(z)] -> dynamics MLP -> next (un-normalised);
 * reward / terminal heads read the UN-normalised next; then min-max scale. */
static void net_dynamics(const orc_net *n, const float *h, int action, int future, float *next_hidden,
                         float *reward_scalar, float *terminal, float *reward_logits, orc_scratch *s) {
    const int H = n->cfg.encoding_size, A = n->cfg.num_actions;
    /* MuZeroStochastic._dynamics (models.py:401-428): one dynamics MLP per future on [h | onehot(a)], no one-hot z */
    const int separate = future >= 0 && n->mlp[ORC_DYN_F1].n_layers > 0;
    const orc_mlp *dyn = (separate && future >= 1) ? &n->mlp[ORC_DYN_F1 + future - 1] : &n->mlp[ORC_DYN];
    const int nf = (future >= 0 && !separate) ? n->cfg.n_futures : 0;
    float *x = (float *)malloc(sizeof(float) * (H + A + nf));
    float rl[2 * 64 + 1];
    memcpy(x, h, sizeof(float) * H);
    for (int i = 0; i < A + nf; ++i) x[H + i] = 0.0f;
    x[H + action] = 1.0f;
    if (nf > 0) x[H + A + future] = 1.0f;
    mlp_forward(dyn, x, s->c, s->a, s->b);
    free(x);
    mlp_forward(&n->mlp[ORC_REWARD], s->c, rl, s->a, s->b);
    *reward_scalar = orc_support_to_scalar(rl, n->cfg.support_size);
    if (reward_logits) memcpy(reward_logits, rl, sizeof(float) * (2 * n->cfg.support_size + 1));
    if (terminal) {
        if (n->mlp[ORC_TERMINAL].n_layers > 0) {
            float t;
            mlp_forward(&n->mlp[ORC_TERMINAL], s->c, &t, s->a, s->b);
            *terminal = t;
        } else {
            *terminal = -1.0f;
        }
    }
    minmax_scale(s->c, next_hidden, H);
}

/* batched network entry points (for the network-parity tests) */
void orc_initial_inference(const orc_net *n, const float *obs, int B, float *value_scalar, float *policy_logits,
                           float *hidden, float *value_logits) {
    const int H = n->cfg.encoding_size, A = n->cfg.num_actions, V = 2 * n->cfg.support_size + 1;
    orc_scratch s = scratch_new(n);
    for (int b = 0; b < B; ++b)
        net_initial(n, obs + (size_t)b * n->cfg.obs_dim, hidden + (size_t)b * H, policy_logits + (size_t)b * A,
                    value_scalar + b, value_logits ? value_logits + (size_t)b * V : NULL, &s);
    scratch_free(&s);
}

void orc_recurrent_inference(const orc_net *n, const float *hidden, const int *action, int B, float *value_scalar,
                             float *reward_scalar, float *terminal, float *policy_logits, float *next_hidden,
                             float *value_logits, float *reward_logits) {
    const int H = n->cfg.encoding_size, A = n->cfg.num_actions, V = 2 * n->cfg.support_size + 1;
    orc_scratch s = scratch_new(n);
    for (int b = 0; b < B; ++b) {
        net_dynamics(n, hidden + (size_t)b * H, action[b], -1, next_hidden + (size_t)b * H, reward_scalar + b,
                     terminal + b, reward_logits ? reward_logits + (size_t)b * V : NULL, &s);
        net_prediction(n, next_hidden + (size_t)b * H, policy_logits + (size_t)b * A, value_scalar + b,
                       value_logits ? value_logits + (size_t)b * V : NULL, &s);
    }
    scratch_free(&s);
}

/* models.py:520-550: all futures of (h, a) + prior logits P(z | h). */
void orc_stochastic_dynamics(const orc_net *n, const float *hidden, const int *action, int B,
                             float *transition_logits /*[B][nf]*/, float *next_hidden /*[B][nf][H]*/,
                             float *reward_scalar /*[B][nf]*/) {
    const int H = n->cfg.encoding_size, nf = n->cfg.n_futures;
    orc_scratch s = scratch_new(n);
    for (int b = 0; b < B; ++b) {
        for (int z = 0; z < nf; ++z)
            net_dynamics(n, hidden + (size_t)b * H, action[b], z, next_hidden + ((size_t)b * nf + z) * H,
                         reward_scalar + (size_t)b * nf + z, NULL, NULL, &s);
        mlp_forward(&n->mlp[ORC_PRIOR], hidden + (size_t)b * H, transition_logits + (size_t)b * nf, s.a, s.b);
    }
    scratch_free(&s);
}

void orc_prediction(const orc_net *n, const float *hidden, int B, float *value_scalar, float *policy_logits) {
    const int H = n->cfg.encoding_size, A = n->cfg.num_actions;
    orc_scratch s = scratch_new(n);
    for (int b = 0; b < B; ++b)
        net_prediction(n, hidden + (size_t)b * H, policy_logits + (size_t)b * A, value_scalar + b, NULL, &s);
    scratch_free(&s);
}

/* --------------------------------------------------------------------- tree ---- */

enum { KIND_STATE = 0, KIND_TRANSITION = 1 };

typedef struct Node {
    int visit_count;
    int kind;
    int n_children; /* 0 = not expanded */
    int to_play;    /* Node.to_play: -1 until expanded (self_play.py:509, 529) */
    double prior;
    double value_sum; /* variants 0,1: sum; variant 2: the `.value` attribute */
    double reward;
    struct Node **children;   /* by slot */
    int *child_action;        /* slot -> action (root: legal list order; others identity) */
    float *hidden;            /* state node: its hidden state (may alias parent's child_hidden) */
    float *child_hidden;      /* transition node: [nf][H] */
    float *child_reward;      /* transition node: decoded rewards [nf] */
} Node;

typedef struct {
    char *base;
    size_t cap, used;
} Arena;

static void *arena_alloc(Arena *a, size_t n) {
    n = (n + 15) & ~(size_t)15;
    if (a->used + n > a->cap) {
        fprintf(stderr, "tsp_oracle: arena overflow\n");
        abort();
    }
    void *p = a->base + a->used;
    a->used += n;
    memset(p, 0, n);
    return p;
}

static Node *node_new(Arena *a, int kind, double prior) {
    Node *n = (Node *)arena_alloc(a, sizeof(Node));
    n->kind = kind;
    n->prior = prior;
    n->to_play = -1;
    return n;
}

typedef struct {
    double maximum, minimum;
} MinMax;

/* self_play.py:669-678 */
static void minmax_update(MinMax *m, double v) {
    m->maximum = m->maximum > v ? m->maximum : v; /* max(self.maximum, value) */
    m->minimum = m->minimum < v ? m->minimum : v;
}
static double minmax_normalize(const MinMax *m, double v) {
    if (m->maximum > m->minimum) return (v - m->minimum) / (m->maximum - m->minimum);
    return v;
}

static double node_value(const Node *n, int variant) {
    if (variant == 2) return n->value_sum; /* attribute, self_play_risk_sensitive.py:426 */
    if (n->visit_count == 0) return 0.0;
    return n->value_sum / (double)n->visit_count;
}

/* self_play.py:453-477, self_play_stochastic.py:488-508, self_play_risk_sensitive.py:496-520 */
static double ucb_score(const orc_config *c, const Node *parent, const Node *child, const MinMax *mm) {
    double pb_c = log(((double)parent->visit_count + c->pb_c_base + 1.0) / c->pb_c_base) + c->pb_c_init;
    pb_c *= sqrt((double)parent->visit_count) / (double)(child->visit_count + 1);
    double prior_score = pb_c * child->prior;
    double value_score;
    if (child->visit_count > 0) {
        if (c->variant == 1)
            value_score = minmax_normalize(mm, node_value(child, 1));
        else if (c->variant == 0 && c->num_players != 1) /* self_play.py:469-472: -child.value() for two players */
            value_score = minmax_normalize(mm, child->reward + c->discount * -node_value(child, 0));
        else
            value_score = minmax_normalize(mm, child->reward + c->discount * node_value(child, c->variant));
    } else {
        value_score = 0.0;
    }
    return prior_score + value_score;
}

typedef struct {
    const uint8_t *tie;
    int T;
    int k_tie;
    const uint8_t *chance;
    int K;
    int k_chance;
} Streams;

/* self_play.py:436-451 */
static int select_child(const orc_config *c, const Node *node, const MinMax *mm, Streams *st) {
    double score[ORC_MAX_ACTIONS];
    double best = 0.0;
    for (int i = 0; i < node->n_children; ++i) {
        score[i] = ucb_score(c, node, node->children[i], mm);
        if (i == 0 || score[i] > best) best = score[i];
    }
    int tied[ORC_MAX_ACTIONS], nt = 0;
    for (int i = 0; i < node->n_children; ++i)
        if (score[i] == best) tied[nt++] = i;
    if (nt == 1) return tied[0];
    int v = 0;
    if (st->tie && st->T > 0) v = st->tie[st->k_tie % st->T];
    st->k_tie++;
    return tied[v % nt];
}

/* self_play_stochastic.py:517-533 */
static int select_outcome(const Node *node, Streams *st) {
    if (node->n_children == 1) return 0;
    int v = 0;
    if (st->chance && st->K > 0) v = st->chance[st->k_chance % st->K];
    st->k_chance++;
    return v % node->n_children;
}

/* self_play.py:524-538: children get softmax priors (float32 values widened to double). */
static void expand_state(Arena *ar, Node *node, int child_kind, const int *actions, int n_actions, double reward,
                         const float *priors, float *hidden) {
    node->reward = reward;
    node->hidden = hidden;
    node->n_children = n_actions;
    node->children = (Node **)arena_alloc(ar, sizeof(Node *) * n_actions);
    node->child_action = (int *)arena_alloc(ar, sizeof(int) * n_actions);
    for (int i = 0; i < n_actions; ++i) {
        node->children[i] = node_new(ar, child_kind, (double)priors[i]);
        node->child_action[i] = actions ? actions[i] : i;
    }
}

typedef struct {
    /* inputs */
    const orc_net *net;
    int S;
    int add_noise, uniform_policy;
    /* per-call outputs for one tree */
    int32_t *visit_counts; /* [A] */
    double *root_value;
    double *child_value, *child_prior, *child_reward; /* [A] */
    int32_t *max_depth;
    float *root_pred_value;
    float *root_hidden;   /* [H] or NULL */
    int32_t *n_records;
    int32_t *n_nodes;
    float *trace_root;    /* [REC] or NULL */
    float *trace_records; /* [max_records][REC] or NULL */
    const float *fed_root;
    const float *fed_records;
    int max_records;
    int to_play;          /* MCTS.run argument (root player) */
    /* n_phases > 1: the search is continued n_phases - 1 times with override_root_with=root
     * (self_play.py:331-333): the tree is kept, MinMaxStats / max_tree_depth / the random streams start
     * afresh and the exploration noise is mixed into the root priors again. Per-phase inputs and
     * records are `stride` elements apart. */
    int n_phases;
    size_t noise_stride, tie_stride, chance_stride, rec_stride;
} TreeJob;

static void rec_clear(float *r) {
    for (int i = 0; i < ORC_REC_WIDTH; ++i) r[i] = 0.0f;
}

static int run_tree(const TreeJob *job, const float *obs, const int *legal, int n_legal, const double *noise,
                    Streams *st, Arena *ar, orc_scratch *sc) {
    const orc_net *net = job->net;
    const orc_config *c = &net->cfg;
    const int A = c->num_actions, H = c->encoding_size, nf = c->n_futures, variant = c->variant;
    const int fed = job->fed_records != NULL || job->fed_root != NULL;
    int all_actions[ORC_MAX_ACTIONS];
    for (int i = 0; i < A; ++i) all_actions[i] = i;
    if (!legal) {
        legal = all_actions;
        n_legal = A;
    }
    ar->used = 0;
    int n_rec = 0, n_nodes = 1;
    float logits[ORC_MAX_ACTIONS], sel[ORC_MAX_ACTIONS], pri[ORC_MAX_ACTIONS];
    float rec[ORC_REC_WIDTH];

    /* ---- root: self_play.py:335-376 ---- */
    Node *root = node_new(ar, KIND_STATE, 0.0);
    float root_pred = 0.0f;
    float *root_h = NULL;
    if (fed) {
        root_pred = job->fed_root[1];
        for (int i = 0; i < n_legal; ++i) pri[i] = job->fed_root[4 + i];
    } else {
        root_h = (float *)arena_alloc(ar, sizeof(float) * H);
        net_initial(net, obs, root_h, logits, &root_pred, NULL, sc);
        for (int i = 0; i < n_legal; ++i) sel[i] = job->uniform_policy ? 0.0f : logits[legal[i]];
        orc_softmax(sel, pri, n_legal);
    }
    if (job->trace_root) {
        rec_clear(job->trace_root);
        job->trace_root[0] = 1.0f;
        job->trace_root[1] = root_pred;
        for (int i = 0; i < n_legal; ++i) job->trace_root[4 + i] = pri[i];
    }
    if (job->root_hidden && root_h) memcpy(job->root_hidden, root_h, sizeof(float) * H);
    expand_state(ar, root, variant == 0 ? KIND_STATE : KIND_TRANSITION, legal, n_legal, 0.0, pri, root_h);
    root->to_play = job->to_play;
    if (job->add_noise) { /* self_play.py:540-549 */
        const double frac = c->root_exploration_fraction;
        for (int i = 0; i < n_legal; ++i)
            root->children[i]->prior = root->children[i]->prior * (1 - frac) + noise[i] * frac;
    }

    const int n_phases = job->n_phases > 1 ? job->n_phases : 1;
    Node **path = (Node **)arena_alloc(ar, sizeof(Node *) * (size_t)(2 * job->S * n_phases + 4) * 2);
    int path_cap = (2 * job->S * n_phases + 4) * 2;
    int max_tree_depth = 0;
    const uint8_t *tie0 = st->tie, *chance0 = st->chance;
    for (int phase = 0; phase < n_phases; ++phase) {
    const float *fed_records = job->fed_records ? job->fed_records + phase * job->rec_stride : NULL;
    float *trace_records = job->trace_records ? job->trace_records + phase * job->rec_stride : NULL;
    if (phase > 0) {
        /* MCTS.run(..., override_root_with=root): self_play.py:331-333, 372-378 */
        if (job->add_noise) {
            const double frac = c->root_exploration_fraction;
            const double *nz = noise + phase * job->noise_stride;
            for (int i = 0; i < n_legal; ++i)
                root->children[i]->prior = root->children[i]->prior * (1 - frac) + nz[i] * frac;
        }
        st->tie = tie0 ? tie0 + phase * job->tie_stride : NULL;
        st->chance = chance0 ? chance0 + phase * job->chance_stride : NULL;
        st->k_tie = 0;
        st->k_chance = 0;
        n_rec = 0;
    }
    MinMax mm;
    mm.maximum = -INFINITY;
    mm.minimum = INFINITY;
    max_tree_depth = 0;

    int num_sims = 0;
    while (num_sims < job->S) {
        Node *node = root;
        int depth = 0, slot = 0, plen = 0;
        int virtual_to_play = job->to_play;
        path[plen++] = node;
        while (node->n_children > 0) {
            depth++;
            if (node->kind == KIND_STATE)
                slot = select_child(c, node, &mm, st);
            else
                slot = select_outcome(node, st);
            node = node->children[slot];
            if (plen >= path_cap) return -2;
            path[plen++] = node;
            /* players play turn by turn, self_play.py:393-397 (config.players == [0, .., n-1]) */
            if (virtual_to_play + 1 < c->num_players)
                virtual_to_play = virtual_to_play + 1;
            else
                virtual_to_play = 0;
        }
        Node *parent = path[plen - 2];
        const int action = parent->child_action[slot];
        double value = 0.0;
        int do_backprop = 1;

        if (variant == 0) {
            /* self_play.py:401-427 */
            float v, r, term;
            float *h = NULL;
            if (fed) {
                if (n_rec >= job->max_records) return -3;
                const float *fr = fed_records + (size_t)n_rec * ORC_REC_WIDTH;
                v = fr[1];
                r = fr[2];
                term = fr[3];
                for (int i = 0; i < A; ++i) pri[i] = fr[4 + i];
            } else {
                h = (float *)arena_alloc(ar, sizeof(float) * H);
                net_dynamics(net, parent->hidden, action, -1, h, &r, &term, NULL, sc);
                net_prediction(net, h, logits, &v, NULL, sc);
                if (job->uniform_policy)
                    for (int i = 0; i < A; ++i) logits[i] = 0.0f;
                orc_softmax(logits, pri, A);
            }
            if (trace_records && n_rec < job->max_records) {
                float *tr = trace_records + (size_t)n_rec * ORC_REC_WIDTH;
                rec_clear(tr);
                tr[0] = 1.0f;
                tr[1] = v;
                tr[2] = r;
                tr[3] = term;
                for (int i = 0; i < A; ++i) tr[4 + i] = pri[i];
            }
            n_rec++;
            const int is_terminal = term >= 0.0f;
            if (is_terminal && c->mask_absorbing_states) {
                value = 0.0;
            } else {
                expand_state(ar, node, KIND_STATE, NULL, A, (double)r, pri, h);
                node->to_play = virtual_to_play; /* Node.expand, self_play.py:529 */
                n_nodes += A;
                value = (double)v;
            }
            if (c->num_players == 1) {
                /* self_play.py:485-490 */
                for (int i = plen - 1; i >= 0; --i) {
                    Node *x = path[i];
                    x->value_sum += value;
                    x->visit_count += 1;
                    minmax_update(&mm, x->reward + c->discount * node_value(x, 0));
                    value = x->reward + c->discount * value;
                }
            } else {
                /* self_play.py:492-500 */
                for (int i = plen - 1; i >= 0; --i) {
                    Node *x = path[i];
                    x->value_sum += (x->to_play == virtual_to_play) ? value : -value;
                    x->visit_count += 1;
                    minmax_update(&mm, x->reward + c->discount * -node_value(x, 0));
                    value = ((x->to_play == virtual_to_play) ? -x->reward : x->reward) + c->discount * value;
                }
            }
            num_sims++;
            (void)do_backprop;
        } else if (node->kind == KIND_STATE) {
            /* self_play_stochastic.py:351-372: parent is a TransitionNode */
            float v;
            float *h = NULL;
            const double r = (double)parent->child_reward[slot];
            if (fed) {
                if (n_rec >= job->max_records) return -3;
                const float *fr = fed_records + (size_t)n_rec * ORC_REC_WIDTH;
                if (fr[0] != 1.0f) return -4;
                v = fr[1];
                for (int i = 0; i < A; ++i) pri[i] = fr[4 + i];
            } else {
                h = parent->child_hidden + (size_t)slot * H;
                net_prediction(net, h, logits, &v, NULL, sc);
                orc_softmax(logits, pri, A);
            }
            if (trace_records && n_rec < job->max_records) {
                float *tr = trace_records + (size_t)n_rec * ORC_REC_WIDTH;
                rec_clear(tr);
                tr[0] = 1.0f;
                tr[1] = v;
                for (int i = 0; i < A; ++i) tr[4 + i] = pri[i];
            }
            n_rec++;
            expand_state(ar, node, KIND_TRANSITION, NULL, A, r, pri, h);
            n_nodes += A;
            value = (double)v;
            if (variant == 1) {
                /* self_play_stochastic.py:408-418 */
                for (int i = plen - 1; i >= 0; --i) {
                    Node *x = path[i];
                    x->value_sum += value;
                    x->visit_count += 1;
                    if (x->kind == KIND_STATE) {
                        value = x->reward + c->discount * value;
                        minmax_update(&mm, x->reward + c->discount * node_value(x, 1));
                    }
                }
            } else {
                /* self_play_risk_sensitive.py:408-420, 452-461, 529-537 */
                path[plen - 1]->value_sum = value;
                path[plen - 1]->visit_count += 1;
                for (int i = plen - 2; i >= 0; --i) {
                    Node *x = path[i];
                    x->visit_count += 1;
                    if (x->kind == KIND_TRANSITION) {
                        double best = 0.0;
                        int have = 0;
                        for (int k = 0; k < x->n_children; ++k) {
                            Node *ch = x->children[k];
                            if (ch->visit_count > 0) {
                                double t = ch->reward + c->discount * ch->value_sum;
                                if (!have || t < best) best = t;
                                have = 1;
                            }
                        }
                        x->value_sum = best;
                    } else {
                        double sum = 0.0;
                        int counts = 0;
                        for (int k = 0; k < x->n_children; ++k) {
                            Node *ch = x->children[k];
                            sum += (double)ch->visit_count * ch->value_sum;
                            counts += ch->visit_count;
                        }
                        x->value_sum = sum / (double)counts;
                        minmax_update(&mm, x->reward + c->discount * x->value_sum);
                    }
                }
            }
            num_sims++;
        } else {
            /* self_play_stochastic.py:374-386: TransitionNode leaf, parent is a StateNode */
            float tp[ORC_MAX_ACTIONS], tl[ORC_MAX_ACTIONS];
            node->child_reward = (float *)arena_alloc(ar, sizeof(float) * nf);
            if (fed) {
                if (n_rec >= job->max_records) return -3;
                const float *fr = fed_records + (size_t)n_rec * ORC_REC_WIDTH;
                if (fr[0] != 2.0f) return -4;
                for (int z = 0; z < nf; ++z) {
                    tp[z] = fr[1 + z];
                    node->child_reward[z] = fr[1 + nf + z];
                }
            } else {
                node->child_hidden = (float *)arena_alloc(ar, sizeof(float) * (size_t)nf * H);
                for (int z = 0; z < nf; ++z)
                    net_dynamics(net, parent->hidden, action, z, node->child_hidden + (size_t)z * H,
                                 node->child_reward + z, NULL, NULL, sc);
                mlp_forward(&net->mlp[ORC_PRIOR], parent->hidden, tl, sc->a, sc->b);
                orc_softmax(tl, tp, nf);
            }
            if (trace_records && n_rec < job->max_records) {
                float *tr = trace_records + (size_t)n_rec * ORC_REC_WIDTH;
                rec_clear(tr);
                tr[0] = 2.0f;
                for (int z = 0; z < nf; ++z) {
                    tr[1 + z] = tp[z];
                    tr[1 + nf + z] = node->child_reward[z];
                }
            }
            n_rec++;
            /* self_play_stochastic.py:535-557: phantom node, reward 0, children StateNodes */
            node->reward = 0.0;
            node->n_children = nf;
            node->children = (Node **)arena_alloc(ar, sizeof(Node *) * nf);
            node->child_action = (int *)arena_alloc(ar, sizeof(int) * nf);
            for (int z = 0; z < nf; ++z) {
                node->children[z] = node_new(ar, KIND_STATE, (double)tp[z]);
                node->child_action[z] = z;
            }
            n_nodes += nf;
        }
        if (depth > max_tree_depth) max_tree_depth = depth;
    }
    } /* phases */

    /* ---- what consumers read from root ---- */
    for (int a = 0; a < A; ++a) {
        job->visit_counts[a] = 0;
        job->child_value[a] = 0.0;
        job->child_prior[a] = 0.0;
        job->child_reward[a] = 0.0;
    }
    for (int i = 0; i < root->n_children; ++i) {
        const int a = root->child_action[i];
        const Node *ch = root->children[i];
        job->visit_counts[a] = ch->visit_count;
        job->child_value[a] = node_value(ch, variant);
        job->child_prior[a] = ch->prior;
        job->child_reward[a] = ch->reward;
    }
    *job->root_value = node_value(root, variant);
    *job->max_depth = max_tree_depth;
    *job->root_pred_value = root_pred;
    *job->n_records = n_rec;
    *job->n_nodes = n_nodes + (int)root->n_children;
    return 0;
}

typedef struct {
    int B, S;
    int add_noise, uniform_policy;
    const float *obs;        /* [B][obs_dim] (ignored when fed) */
    const int32_t *legal;    /* [B][A] ordered legal-action lists, or NULL = all */
    const int32_t *n_legal;  /* [B] or NULL */
    const double *noise;     /* [B][A] aligned with the legal list, or NULL */
    const uint8_t *tie;      /* [B][T] or NULL */
    int T;
    const uint8_t *chance;   /* [B][K] or NULL */
    int K;
    const float *fed_root;    /* [B][REC] or NULL */
    const float *fed_records; /* [B][max_records][REC] or NULL */
    int max_records;
    /* outputs */
    int32_t *visit_counts;   /* [B][A] */
    double *root_value;      /* [B] */
    double *child_value;     /* [B][A] */
    double *child_prior;     /* [B][A] */
    double *child_reward;    /* [B][A] */
    int32_t *max_depth;      /* [B] */
    float *root_pred_value;  /* [B] */
    float *root_hidden;      /* [B][H] or NULL */
    int32_t *n_records;      /* [B] */
    int32_t *n_nodes;        /* [B] */
    float *trace_root;       /* [B][REC] or NULL */
    float *trace_records;    /* [B][max_records][REC] or NULL */
    int n_threads;           /* 0 = OpenMP default */
    const int32_t *to_play;  /* [B] MCTS.run's to_play per root, or NULL = 0 */
    int n_phases;            /* > 1: continue the search n_phases - 1 times (override_root_with); noise / tie /
                                chance / fed_records / trace_records then carry a leading [n_phases] dimension */
} orc_run_args;

#include <pthread.h>
#include <stdatomic.h>
#include <unistd.h>

int orc_max_threads(void) {
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

typedef struct {
    const orc_net *net;
    const orc_run_args *a;
    atomic_int *next;
    atomic_int *status;
    size_t arena_cap;
} Worker;

#define ORC_CHUNK 16

static void *worker_main(void *p) {
    Worker *w = (Worker *)p;
    const orc_net *net = w->net;
    const orc_run_args *a = w->a;
    const orc_config *c = &net->cfg;
    const int A = c->num_actions, H = c->encoding_size;
    Arena ar;
    ar.base = (char *)malloc(w->arena_cap);
    ar.cap = w->arena_cap;
    ar.used = 0;
    orc_scratch sc = scratch_new(net);
    for (;;) {
        const int b0 = atomic_fetch_add(w->next, ORC_CHUNK);
        if (b0 >= a->B) break;
        const int b1 = b0 + ORC_CHUNK < a->B ? b0 + ORC_CHUNK : a->B;
        for (int b = b0; b < b1; ++b) {
            TreeJob job;
            memset(&job, 0, sizeof(job));
            job.net = net;
            job.S = a->S;
            job.to_play = a->to_play ? a->to_play[b] : 0;
            job.add_noise = a->add_noise;
            job.uniform_policy = a->uniform_policy;
            job.visit_counts = a->visit_counts + (size_t)b * A;
            job.root_value = a->root_value + b;
            job.child_value = a->child_value + (size_t)b * A;
            job.child_prior = a->child_prior + (size_t)b * A;
            job.child_reward = a->child_reward + (size_t)b * A;
            job.max_depth = a->max_depth + b;
            job.root_pred_value = a->root_pred_value + b;
            job.root_hidden = a->root_hidden ? a->root_hidden + (size_t)b * H : NULL;
            job.n_records = a->n_records + b;
            job.n_nodes = a->n_nodes + b;
            job.trace_root = a->trace_root ? a->trace_root + (size_t)b * ORC_REC_WIDTH : NULL;
            job.trace_records = a->trace_records ? a->trace_records + (size_t)b * a->max_records * ORC_REC_WIDTH : NULL;
            job.fed_root = a->fed_root ? a->fed_root + (size_t)b * ORC_REC_WIDTH : NULL;
            job.fed_records = a->fed_records ? a->fed_records + (size_t)b * a->max_records * ORC_REC_WIDTH : NULL;
            job.max_records = a->max_records;
            job.n_phases = a->n_phases;
            job.noise_stride = (size_t)a->B * A;
            job.tie_stride = (size_t)a->B * a->T;
            job.chance_stride = (size_t)a->B * a->K;
            job.rec_stride = (size_t)a->B * a->max_records * ORC_REC_WIDTH;
            Streams st;
            st.tie = a->tie ? a->tie + (size_t)b * a->T : NULL;
            st.T = a->T;
            st.k_tie = 0;
            st.chance = a->chance ? a->chance + (size_t)b * a->K : NULL;
            st.K = a->K;
            st.k_chance = 0;
            const int32_t *lg = a->legal ? a->legal + (size_t)b * A : NULL;
            int nl = (a->legal && a->n_legal) ? a->n_legal[b] : A;
            int lgi[ORC_MAX_ACTIONS];
            if (lg)
                for (int i = 0; i < nl; ++i) lgi[i] = lg[i];
            int rc = run_tree(&job, a->obs ? a->obs + (size_t)b * c->obs_dim : NULL, lg ? lgi : NULL, nl,
                              a->noise ? a->noise + (size_t)b * A : NULL, &st, &ar, &sc);
            if (rc != 0) atomic_store(w->status, rc);
        }
    }
    scratch_free(&sc);
    free(ar.base);
    return NULL;
}

int orc_run(const orc_net *net, const orc_run_args *a) {
    const orc_config *c = &net->cfg;
    const int A = c->num_actions, H = c->encoding_size;
    if (c->num_players != 1 && !(c->num_players == 2 && c->variant == 0)) return -10;
    if (A > ORC_MAX_ACTIONS) return -11;
    const int S_all = a->S * (a->n_phases > 1 ? a->n_phases : 1);
    const int iters_cap = (c->variant == 0) ? S_all : S_all * (A + 2) + A + 2;
    const size_t node_bytes = sizeof(Node) + 64;
    const size_t per_iter = (size_t)(A + 2) * (node_bytes + 16) + sizeof(float) * (size_t)(c->n_futures + 1) * (H + 8) + 256;
    const size_t arena_cap = (size_t)(iters_cap + 4) * per_iter + sizeof(Node *) * (size_t)(4 * S_all + 16) * 2 + (1 << 16);
    int nthreads = a->n_threads > 0 ? a->n_threads : orc_max_threads();
    if (nthreads > (a->B + ORC_CHUNK - 1) / ORC_CHUNK) nthreads = (a->B + ORC_CHUNK - 1) / ORC_CHUNK;
    if (nthreads < 1) nthreads = 1;
    atomic_int next = 0, status = 0;
    Worker w = {net, a, &next, &status, arena_cap};
    if (nthreads == 1) {
        worker_main(&w);
    } else {
        pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * nthreads);
        for (int i = 0; i < nthreads; ++i) pthread_create(&th[i], NULL, worker_main, &w);
        for (int i = 0; i < nthreads; ++i) pthread_join(th[i], NULL);
        free(th);
    }
    return atomic_load(&status);
}
