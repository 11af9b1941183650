// tsp_rootio.cuh -- the data formats on either side of the search, on the device:
//   stack_observations_kernel   GameHistory.get_stacked_observations        self_play.py:587-624
//   select_action_kernel        SelfPlay.select_action                      self_play.py:277-299
//   search_statistics_kernel    GameHistory.store_search_statistics         self_play.py:570-585
// so that a batch of games never round-trips per root: the host only uploads the newest frame of every
// game and downloads one action per game. All three are HBM-bound gathers / element-wise passes.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

namespace tsp {

// out[b] = [ frame(idx[b][0]) | for i = 1..so: frame(idx[b][i]), plane filled with action_value[b][i-1] ]
// (index -1: the reference pads missing history with zeros, self_play.py:611-617). One CTA per row,
// float4 when the frame size allows it.
__global__ void stack_observations_kernel(int B, int so, int F, int HW, const float* __restrict__ pool,
                                          const int* __restrict__ frame_index, const float* __restrict__ action_value,
                                          float* __restrict__ out) {
    const int b = blockIdx.x;
    if (b >= B) return;
    const int row = F * (so + 1) + so * HW;
    float* o = out + (size_t)b * row;
    const int* idx = frame_index + (size_t)b * (so + 1);
    for (int i = 0; i <= so; ++i) {
        const int fi = idx[i];
        const float* src = fi >= 0 ? pool + (size_t)fi * F : nullptr;
        float* dst = o + (i == 0 ? 0 : F + (i - 1) * (F + HW));
        for (int k = threadIdx.x; k < F; k += blockDim.x) dst[k] = src ? src[k] : 0.0f;
        if (i > 0) {
            const float a = action_value[(size_t)b * so + (i - 1)];
            for (int k = threadIdx.x; k < HW; k += blockDim.x) dst[F + k] = a;
        }
    }
}

// One thread per root. temperature == 0: first arg-max over the ordered legal list (numpy.argmax);
// +inf: actions[floor(u * n)]; else numpy.random.choice(actions, p = v^(1/T) / sum) with the caller's
// uniform sample: cdf = cumsum(p); cdf /= cdf[-1]; searchsorted(cdf, u, side='right').
__global__ void select_action_kernel(int B, int A, const int* __restrict__ visits, const int* __restrict__ legal,
                                     const int* __restrict__ num_legal, double temperature,
                                     const double* __restrict__ uniform, int* __restrict__ action) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const int n = num_legal ? num_legal[b] : A;
    const int* v = visits + (size_t)b * A;
    const int* la = legal ? legal + (size_t)b * A : nullptr;
    int pick = 0;
    if (temperature == 0.0) {
        int best = -1;
        for (int i = 0; i < n; ++i) {
            const int c = v[la ? la[i] : i];
            if (c > best) {
                best = c;
                pick = i;
            }
        }
    } else if (isinf(temperature)) {
        pick = (int)floor(uniform[b] * (double)n);
        if (pick >= n) pick = n - 1;
    } else {
        double p[8], sum = 0.0;
        const double inv_t = 1.0 / temperature;
        for (int i = 0; i < n; ++i) {
            p[i] = pow((double)v[la ? la[i] : i], inv_t);  // visit_counts ** (1 / temperature)
            sum += p[i];                                   // python sum(): left to right
        }
        double cdf[8], run = 0.0;
        for (int i = 0; i < n; ++i) {
            run += p[i] / sum;  // numpy cumsum of p / sum
            cdf[i] = run;
        }
        const double last = cdf[n - 1], u = uniform[b];
        pick = n - 1;
        for (int i = 0; i < n; ++i)
            if (u < cdf[i] / last) {  // searchsorted(..., side='right'): first i with u < cdf[i]
                pick = i;
                break;
            }
    }
    action[b] = la ? la[pick] : pick;
}

// child_visits[b][a] = visit_count / max(sum_visits, 1) in float64 (Python int / int), 0 for illegal actions
__global__ void search_statistics_kernel(int B, int A, const int* __restrict__ visits, double* __restrict__ child_visits) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    int sum = 0;
    for (int a = 0; a < A; ++a) sum += visits[(size_t)b * A + a];
    const double den = (double)(sum > 1 ? sum : 1);
    for (int a = 0; a < A; ++a) child_visits[(size_t)b * A + a] = __ddiv_rn((double)visits[(size_t)b * A + a], den);
}

// The all-gather of the packed root statistics as peer stores (NVLink / NVSwitch): every rank writes its slice into the
// same offset of every rank's symmetric buffer -- one kernel, no ring, no protocol; the cross-GPU barrier behind it is the
// caller's (torch symmetric memory). Independent roots shard across GPUs, so this is the only exchange of a search.
struct PeerPtrs {
    unsigned char* p[16];
};
__global__ void peer_put_kernel(const uint4* __restrict__ src, long long n16, PeerPtrs peers, int n_peers) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n16) return;
    const uint4 v = src[i];
    for (int r = 0; r < n_peers; ++r) reinterpret_cast<uint4*>(peers.p[r])[i] = v;
}

}  // namespace tsp
