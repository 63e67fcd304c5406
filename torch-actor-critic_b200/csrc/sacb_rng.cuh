// sacb_rng.cuh -- on-device draws shared by the MLP engine (sacb_engine.cu) and the visual engine (sacv_engine.cu):
// duplicate-free uniform batch indices (ReplayBuffer.sample's random.sample) and the standard normal noise of Normal.rsample.
#pragma once
#include <cuda_runtime.h>
#include <curand_kernel.h>

namespace sacb_rng {

// Duplicate-free uniform indices (random.sample(range(size), B), buffer/replay_buffer.py:46).  One CTA per update.
// Sparse regime (size >= 4B): every position draws from its own Philox subsequence; a position that collides with an
// EARLIER position redraws (next counter of its subsequence) until the batch is duplicate free (each round succeeds
// with probability >= 3/4 per position).  The procedure only compares values for equality, so the result is uniform
// over ordered duplicate-free B-tuples.
// Dense regime (size < 4B, e.g. the first bursts after update_after): exact partial Fisher-Yates over an explicit
// permutation of range(size) in global scratch (parallel init, B sequential swaps by one thread).
static __global__ void __launch_bounds__(256) draw_indices_kernel(long long* __restrict__ out, int B, long long size,
                                                           unsigned long long seed, unsigned long long counter0,
                                                           int* __restrict__ perm_scratch) {
    extern __shared__ long long cand[];
    __shared__ int any_dup;
    const unsigned long long u = counter0 + blockIdx.x;
    auto draw_below = [&](int i, int round, unsigned long long bound) -> long long {
        curandStatePhilox4_32_10_t st;
        // Philox offsets count single 32-bit outputs and curand4 consumes four: round r starts at output 4r
        curand_init(seed, (u << 20) + (unsigned long long)i, 4ULL * (unsigned long long)round, &st);
        uint4 r = curand4(&st);
        unsigned long long x = ((unsigned long long)r.x << 32) | r.y;
        return (long long)__umul64hi(x, bound);
    };
    if (size < 4LL * B) {
        int* perm = perm_scratch + (size_t)blockIdx.x * 4 * B;
        const int n = (int)size;
        for (int i = threadIdx.x; i < n; i += blockDim.x) perm[i] = i;
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int i = 0; i < B; ++i) {
                int j = i + (int)draw_below(i, 0, (unsigned long long)(n - i));
                int a = perm[i], b = perm[j];
                perm[i] = b; perm[j] = a;
                out[(size_t)blockIdx.x * B + i] = b;
            }
        }
        return;
    }
    for (int i = threadIdx.x; i < B; i += blockDim.x) cand[i] = draw_below(i, 0, (unsigned long long)size);
    __syncthreads();
    for (int round = 1; round < 256; ++round) {
        if (threadIdx.x == 0) any_dup = 0;
        __syncthreads();
        long long redraw_mask = 0;       // up to 64 positions per thread (B <= 16384)
        int q = 0;
        for (int i = threadIdx.x; i < B; i += blockDim.x, ++q) {
            const long long v = cand[i];
            bool dup = false;
            for (int j = 0; j < i; ++j) dup |= (cand[j] == v);
            if (dup) { redraw_mask |= (1LL << q); any_dup = 1; }
        }
        __syncthreads();
        if (!any_dup) break;
        q = 0;
        for (int i = threadIdx.x; i < B; i += blockDim.x, ++q)
            if (redraw_mask & (1LL << q)) cand[i] = draw_below(i, round, (unsigned long long)size);
        __syncthreads();
    }
    for (int i = threadIdx.x; i < B; i += blockDim.x) out[(size_t)blockIdx.x * B + i] = cand[i];
}

// Standard normal noise for Normal.rsample (networks/linear.py:45): 4 values per thread, Philox + Box-Muller.
static __global__ void draw_normal_kernel(float* __restrict__ out, long long n, unsigned long long seed, unsigned long long counter0) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long base = t * 4;
    if (base >= n) return;
    curandStatePhilox4_32_10_t st;
    curand_init(seed ^ 0x9E3779B97F4A7C15ULL, counter0 + (unsigned long long)t, 0ULL, &st);
    float4 v = curand_normal4(&st);
    float vals[4] = {v.x, v.y, v.z, v.w};
    for (int i = 0; i < 4 && base + i < n; ++i) out[base + i] = vals[i];
}

}  // namespace sacb_rng
