// sacb_act.cuh -- low-latency acting path: Actor.forward (reference networks/linear.py:32-53) for a handful of rows
// (the per-environment-step call `action, _ = actor(state)`, sac/algorithm.py:233, and run_agent.py:37).
//
// ONE launch of ONE thread-block cluster of 8 CTAs.  Every CTA owns 1/8 of the output neurons of each layer, so the
// 294 KB of actor weights are pulled from L2 by 8 SMs in parallel; the activations never touch global memory: each
// CTA publishes its slice in its own shared memory, a cluster barrier (barrier.cluster, ~0.2 us) replaces the grid
// barrier, and the next layer first pulls the full activation vector out of the 8 CTAs' shared memory (DSMEM).
// CTA 0 gathers the [mu | log_std] heads, applies the tanh-Gaussian (same head_eval as the training kernel) and writes
// the action and log-prob straight into the caller's arrays.
#pragma once
#include <cooperative_groups.h>
#include <curand_kernel.h>

namespace cg = cooperative_groups;

constexpr int kActCluster = 8;
constexpr int kActMaxRows = 4;        // measured: 1 row 20.5 us host-bound (phase path 28.7); beyond ~4 rows the batched phase path wins

struct ActArgs {
    const float* x;            // [n, O] caller's states
    const float* w0; const float* b0; const float* w1; const float* b1; const float* wh; const float* bh;
    const float* eps;          // [n, A] injected noise, or nullptr
    float* pi; float* logp;    // outputs [n, A], [n] (logp may be nullptr)
    int n, O, A, H1, H2;
    int deterministic;
    unsigned long long seed, counter;
    float log_min_std, log_max_std, act_limit;
    unsigned int flags;
};

__host__ __device__ inline size_t act_smem_floats(int n, int O, int A, int H1, int H2) {
    const int c1 = (H1 + kActCluster - 1) / kActCluster, c2 = (H2 + kActCluster - 1) / kActCluster, ch = (2 * A + kActCluster - 1) / kActCluster;
    return (size_t)n * (O + c1 + c2 + ch + H1 + H2 + 2 * A);
}

// y[row][j] for this CTA's neuron slice: warp per neuron, lanes stride the reduction (coalesced weight reads), rows in
// chunks of 8 accumulators; the input [n][K] is in local shared memory.
template <typename OutFn>
__device__ __forceinline__ void act_layer(const float* __restrict__ W, const float* __restrict__ bias, const float* in, int K,
                                          int j_begin, int j_end, int n, OutFn out) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
    for (int j = j_begin + warp; j < j_end; j += nwarp) {
        const float* __restrict__ wr = W + (size_t)j * K;
        const float bj = __ldg(bias + j);
        for (int r0 = 0; r0 < n; r0 += 8) {
            float acc[8];
#pragma unroll
            for (int r = 0; r < 8; ++r) acc[r] = 0.f;
            for (int k = lane; k < K; k += 32) {
                const float w = __ldg(wr + k);
#pragma unroll
                for (int r = 0; r < 8; ++r)
                    if (r0 + r < n) acc[r] = fmaf(w, in[(r0 + r) * K + k], acc[r]);
            }
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const float s = warp_sum(acc[r]);
                if (lane == 0 && r0 + r < n) out(r0 + r, j, s + bj);
            }
        }
    }
}

// pull the full [n][H] activation out of the cluster: neuron k lives in CTA k / c, slot k % c
__device__ __forceinline__ void act_pull(cg::cluster_group& cluster, float* slice, float* full, int n, int H, int c) {
    for (int i = threadIdx.x; i < n * H; i += blockDim.x) {
        const int r = i / H, k = i - r * H, q = k / c;
        const float* src = cluster.map_shared_rank(slice, q);
        full[i] = src[r * c + (k - q * c)];
    }
    __syncthreads();
}

__global__ void __cluster_dims__(kActCluster, 1, 1) __launch_bounds__(256, 1) actor_act_kernel(const __grid_constant__ ActArgs a) {
    extern __shared__ __align__(16) float sm[];
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int n = a.n, O = a.O, A = a.A, H1 = a.H1, H2 = a.H2;
    const int c1 = (H1 + kActCluster - 1) / kActCluster, c2 = (H2 + kActCluster - 1) / kActCluster;
    const int ch = (2 * A + kActCluster - 1) / kActCluster;
    // shared layout (same in every CTA): x [n][O] | h1 slice [n][c1] | h2 slice [n][c2] | heads slice [n][ch] |
    //                                    h1 full [n][H1] | h2 full [n][H2] | heads full [n][2A]
    float* sx = sm;
    float* s1 = sx + n * O;
    float* s2 = s1 + n * c1;
    float* sh = s2 + n * c2;
    float* f1 = sh + n * ch;
    float* f2 = f1 + n * H1;
    float* fh = f2 + n * H2;
    for (int i = threadIdx.x; i < n * O; i += blockDim.x) sx[i] = __ldg(a.x + i);
    __syncthreads();

    act_layer(a.w0, a.b0, sx, O, min(rank * c1, H1), min((rank + 1) * c1, H1), n,
              [&](int r, int j, float v) { s1[r * c1 + (j - rank * c1)] = fmaxf(v, 0.f); });
    cluster.sync();
    act_pull(cluster, s1, f1, n, H1, c1);
    act_layer(a.w1, a.b1, f1, H1, min(rank * c2, H2), min((rank + 1) * c2, H2), n,
              [&](int r, int j, float v) { s2[r * c2 + (j - rank * c2)] = fmaxf(v, 0.f); });
    cluster.sync();
    act_pull(cluster, s2, f2, n, H2, c2);
    // heads: [mu ; log_std] = one [2A, H2] matrix
    act_layer(a.wh, a.bh, f2, H2, min(rank * ch, 2 * A), min((rank + 1) * ch, 2 * A), n,
              [&](int r, int j, float v) { sh[r * ch + (j - rank * ch)] = v; });
    cluster.sync();
    if (rank == 0) {
        act_pull(cluster, sh, fh, n, 2 * A, ch);
        // tanh-Gaussian rows (networks/linear.py:37-51): 16 lanes per row, as in the training kernel's ROW_HEADS job
        RunArgs hp;                                   // only the scalars head_eval reads
        hp.log_min_std = a.log_min_std; hp.log_max_std = a.log_max_std; hp.act_limit = a.act_limit; hp.flags = a.flags;
        const int g = threadIdx.x >> 4, l = threadIdx.x & 15;
        for (int row = g; row < ((n + 15) & ~15); row += blockDim.x >> 4) {
            float sl = 0.f, sc = 0.f;
            if (row < n) {
                for (int act = l; act < A; act += 16) {
                    const float mu = fh[row * 2 * A + act], lsr = fh[row * 2 * A + A + act];
                    float e = 0.f;
                    if (!a.deterministic) {
                        if (a.eps) e = __ldg(a.eps + (size_t)row * A + act);
                        else {
                            curandStatePhilox4_32_10_t st;
                            curand_init(a.seed, a.counter + (unsigned long long)(row * A + act), 0ULL, &st);
                            e = curand_normal(&st);
                        }
                    }
                    const HeadVals h = head_eval(mu, lsr, e, a.deterministic != 0, hp);
                    a.pi[(size_t)row * A + act] = __fmul_rn(h.th, a.act_limit);
                    sl += h.logp_term; sc += h.corr_term;
                }
            }
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) { sl += __shfl_xor_sync(0xffffffffu, sl, o); sc += __shfl_xor_sync(0xffffffffu, sc, o); }
            if (row < n && l == 0 && a.logp) a.logp[row] = __fsub_rn(sl, sc);
        }
    }
    cluster.sync();            // peers' shared memory must stay alive until CTA 0 has read the heads
}
