// sacv_kernels.cuh -- device kernels of the VISUAL SAC engine (reference networks/convolutional.py, buffer/visual_replay_buffer.py).
//
// The visual update (BASELINE configs[4]: 3x84x84 frames, batch 256) is 61 GFLOP of convolutions per update, i.e. GEMM
// shaped and throughput bound -- unlike the 0.6 GFLOP latency-bound MLP update that the persistent kernel of
// sacb_device.cuh serves.  It therefore runs as a stream of ordinary launches: uint8 frames are gathered from the HBM ring straight into im2col form, every convolution / dense layer is one
// tensor-core GEMM with a fused bias / ReLU epilogue (mma.sync m16n8k8 TF32, optionally 3xTF32 = fp32-level accuracy like the
// MLP engine), weight gradients are split-K GEMMs with a fixed-order reduction (deterministic), col2im is a gather.
// Activations live in [batch*pixels, channels] layout (what the GEMMs produce); the NCHW order of the reference only
// matters at Flatten (convolutional.py:45) and is restored by a small permutation kernel.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sacv {

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp16(void* s, const void* g, int bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(smem_u32(s)), "l"(g), "r"(bytes));
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }
__device__ __forceinline__ unsigned f2tf32(float x) { unsigned r; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x)); return r; }
__device__ __forceinline__ void mma_tf32(float (&d)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
    asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// ------------------------------------------------------------------------------------------------------------
// GEMM  C[M,N] = epi( sum_k A(m,k) * B(n,k) ),  row-major C with pitch ldc.
//   AK: A(m,k) = A[m*lda + k]  (k contiguous)   else A[k*lda + m]  (m contiguous: the transposed use of a stored matrix)
//   BK: B(n,k) = B[n*ldb + k]                    else B[k*ldb + n]
// CTA tile 128 x BN x 32, 8 warps (4 along M x 2 along N), 3-stage cp.async pipeline, fragments split into hi/lo TF32 at
// load for the 3xTF32 path.  gridDim.y = split-K parts (each writes its own partial C at + y*split_stride, epilogue
// applied by the reduction kernel), gridDim.z = batch (twin critics).
// ------------------------------------------------------------------------------------------------------------
enum Epi : int { EPI_NONE = 0, EPI_BIAS = 1, EPI_BIAS_RELU = 2 };
struct GemmArgs {
    const float* A; const float* B; float* C; const float* bias;
    int M, N, K, lda, ldb, ldc, epi, x3;             // x3: 1 = 3xTF32 error compensation, 0 = plain TF32
    long long sA, sB, sC, sBias;                     // batch strides (floats)
    int k_per_split; long long split_stride;         // split-K: this CTA reduces k in [y*k_per_split, ...)
};
constexpr int kGM = 128, kGK = 32, kGStages = 3, kGThreads = 256;
constexpr int kPadK = kGK + 4;                       // k-contiguous operand rows: [rows][36]
template <int BN> constexpr int gemm_smem_floats() {
    // worst case of the two layouts per operand
    constexpr int a = (kGM * kPadK > kGK * (kGM + 8)) ? kGM * kPadK : kGK * (kGM + 8);
    constexpr int b = (BN * kPadK > kGK * (BN + 8)) ? BN * kPadK : kGK * (BN + 8);
    return kGStages * (a + b);
}

// one operand tile [ROWS x 32(k)] into shared memory; 16-byte cp.async when the source allows, guarded scalar loads otherwise
template <int ROWS, bool KMAJOR>
__device__ __forceinline__ void load_tile(float* s, const float* __restrict__ g, int ld, int row0, int rows_valid, int k0, int k_valid, bool vec) {
    const int t = threadIdx.x;
    if (KMAJOR) {                                    // smem [ROWS][kPadK]
        for (int v = t; v < ROWS * (kGK / 4); v += kGThreads) {
            const int r = v >> 3, c = (v & 7) << 2;
            float* dst = s + r * kPadK + c;
            const int nv = (r < rows_valid) ? min(4, k_valid - c) : 0;
            const float* src = g + (size_t)(row0 + r) * ld + k0 + c;
            if (vec) cp16(dst, nv > 0 ? src : g, nv > 0 ? nv * 4 : 0);
            else {
#pragma unroll
                for (int q = 0; q < 4; ++q) dst[q] = q < nv ? __ldg(src + q) : 0.f;
            }
        }
    } else {                                         // smem [32][ROWS + 8]
        constexpr int P = ROWS + 8;
        for (int v = t; v < kGK * (ROWS / 4); v += kGThreads) {
            const int k = v / (ROWS / 4), c = (v % (ROWS / 4)) << 2;
            float* dst = s + k * P + c;
            const int nv = (k < k_valid) ? min(4, rows_valid - c) : 0;
            const float* src = g + (size_t)(k0 + k) * ld + row0 + c;
            if (vec) cp16(dst, nv > 0 ? src : g, nv > 0 ? nv * 4 : 0);
            else {
#pragma unroll
                for (int q = 0; q < 4; ++q) dst[q] = q < nv ? __ldg(src + q) : 0.f;
            }
        }
    }
}

template <int BN, bool AK, bool BK>
__global__ void __launch_bounds__(kGThreads) gemm_kernel(const GemmArgs g) {
    extern __shared__ __align__(16) float sm[];
    constexpr int kA = (kGM * kPadK > kGK * (kGM + 8)) ? kGM * kPadK : kGK * (kGM + 8);
    constexpr int kB = (BN * kPadK > kGK * (BN + 8)) ? BN * kPadK : kGK * (BN + 8);
    constexpr int NI = BN / 16;                      // n8 fragments per warp (warp tile 32 x BN/2)
    const int tn = (g.N + BN - 1) / BN;
    const int tile_m = blockIdx.x / tn, tile_n = blockIdx.x - tile_m * tn;
    const int m0 = tile_m * kGM, n0 = tile_n * BN;
    const int z = blockIdx.z;
    const float* A = g.A + (size_t)z * g.sA;
    const float* B = g.B + (size_t)z * g.sB;
    const int kb = blockIdx.y * g.k_per_split, ke = min(g.K, kb + g.k_per_split);
    const int m_valid = min(kGM, g.M - m0), n_valid = min(BN, g.N - n0);
    const bool vecA = ((reinterpret_cast<uintptr_t>(A) & 15) == 0) && (g.lda % 4 == 0) && (AK ? (kb % 4 == 0) : (m0 % 4 == 0));
    const bool vecB = ((reinterpret_cast<uintptr_t>(B) & 15) == 0) && (g.ldb % 4 == 0) && (BK ? (kb % 4 == 0) : (n0 % 4 == 0));
    const int nk = (ke - kb + kGK - 1) / kGK;
    auto issue = [&](int it) {
        float* As = sm + (it % kGStages) * (kA + kB);
        float* Bs = As + kA;
        const int k0 = kb + it * kGK, kv = min(kGK, ke - k0);
        load_tile<kGM, AK>(As, A, g.lda, m0, m_valid, k0, kv, vecA);
        load_tile<BN, BK>(Bs, B, g.ldb, n0, n_valid, k0, kv, vecB);
    };
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, gq = lane >> 2, t4 = lane & 3;
    const int wm = (warp & 3) * 32, wn = (warp >> 2) * (BN / 2);
    float acc[2][NI][4];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < NI; ++j)
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[i][j][q] = 0.f;
#pragma unroll
    for (int s = 0; s < kGStages - 1; ++s) { if (s < nk) issue(s); cp_commit(); }
    for (int it = 0; it < nk; ++it) {
        if (it + kGStages - 1 < nk) issue(it + kGStages - 1);
        cp_commit();
        cp_wait<kGStages - 1>();
        __syncthreads();
        const float* As = sm + (it % kGStages) * (kA + kB);
        const float* Bs = As + kA;
#pragma unroll
        for (int ks = 0; ks < kGK; ks += 8) {
            unsigned ah[2][4], al[2][4], bh[NI][2], bl[NI][2];
#pragma unroll
            for (int mi = 0; mi < 2; ++mi) {
                const int r0 = wm + mi * 16 + gq, r1 = r0 + 8, k0 = ks + t4, k1 = k0 + 4;
                float x[4];
                if (AK) { x[0] = As[r0 * kPadK + k0]; x[1] = As[r1 * kPadK + k0]; x[2] = As[r0 * kPadK + k1]; x[3] = As[r1 * kPadK + k1]; }
                else { constexpr int P = kGM + 8; x[0] = As[k0 * P + r0]; x[1] = As[k0 * P + r1]; x[2] = As[k1 * P + r0]; x[3] = As[k1 * P + r1]; }
#pragma unroll
                for (int e = 0; e < 4; ++e) { ah[mi][e] = f2tf32(x[e]); al[mi][e] = f2tf32(x[e] - __uint_as_float(ah[mi][e])); }
            }
#pragma unroll
            for (int ni = 0; ni < NI; ++ni) {
                const int n = wn + ni * 8 + gq, k0 = ks + t4, k1 = k0 + 4;
                float x[2];
                if (BK) { x[0] = Bs[n * kPadK + k0]; x[1] = Bs[n * kPadK + k1]; }
                else { constexpr int P = BN + 8; x[0] = Bs[k0 * P + n]; x[1] = Bs[k1 * P + n]; }
#pragma unroll
                for (int e = 0; e < 2; ++e) { bh[ni][e] = f2tf32(x[e]); bl[ni][e] = f2tf32(x[e] - __uint_as_float(bh[ni][e])); }
            }
            if (g.x3) {
#pragma unroll
                for (int mi = 0; mi < 2; ++mi)
#pragma unroll
                    for (int ni = 0; ni < NI; ++ni) { mma_tf32(acc[mi][ni], al[mi], bh[ni]); mma_tf32(acc[mi][ni], ah[mi], bl[ni]); }
            }
#pragma unroll
            for (int mi = 0; mi < 2; ++mi)
#pragma unroll
                for (int ni = 0; ni < NI; ++ni) mma_tf32(acc[mi][ni], ah[mi], bh[ni]);
        }
        __syncthreads();
    }
    cp_wait<0>();
    float* C = g.C + (size_t)z * g.sC + (size_t)blockIdx.y * g.split_stride;
    const float* bias = g.bias ? g.bias + (size_t)z * g.sBias : nullptr;
    const int epi = gridDim.y > 1 ? EPI_NONE : g.epi;
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int ni = 0; ni < NI; ++ni)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int m = m0 + wm + mi * 16 + gq + h * 8;
                const int n = n0 + wn + ni * 8 + 2 * t4;
                if (m < g.M) {
#pragma unroll
                    for (int q = 0; q < 2; ++q)
                        if (n + q < g.N) {
                            float v = acc[mi][ni][h * 2 + q];
                            if (epi != EPI_NONE && bias) v += __ldg(bias + n + q);
                            if (epi == EPI_BIAS_RELU) v = fmaxf(v, 0.f);
                            C[(size_t)m * g.ldc + n + q] = v;
                        }
                }
            }
}

// split-K reduction in fixed order + epilogue: C[m,n] = epi( sum_s P[s][m,n] )
__global__ void splitk_reduce_kernel(const float* __restrict__ P, float* __restrict__ C, const float* __restrict__ bias, int M, int N, int ldc,
                                     int splits, long long split_stride, int epi, long long sP, long long sC, long long sBias) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int z = blockIdx.z;
    if (i >= (long long)M * N) return;
    const int m = (int)(i / N), n = (int)(i - (long long)m * N);
    const float* p = P + (size_t)z * sP + (size_t)m * ldc + n;
    float s = 0.f;
    for (int k = 0; k < splits; ++k) s += p[(size_t)k * split_stride];
    if (epi != EPI_NONE && bias) s += bias[(size_t)z * sBias + n];
    if (epi == EPI_BIAS_RELU) s = fmaxf(s, 0.f);
    C[(size_t)z * sC + (size_t)m * ldc + n] = s;
}

// ------------------------------------------------------------------------------------------------------------
// replay ring -> batch: features / actions / rewards / done rows, and uint8 frames straight into im2col form for conv 0
//   col[(b*OH + oy)*OW + ox][(c*KH + ky)*KW + kx] = float(frame[idx[b]][c][oy*S + ky][ox*S + kx])     (frames raw 0..255)
// ------------------------------------------------------------------------------------------------------------
__global__ void gather_rows_kernel(const long long* __restrict__ idx, int B, int O, int A, const float* __restrict__ feat,
                                   const float* __restrict__ nfeat, const float* __restrict__ act, const float* __restrict__ rew,
                                   const unsigned char* __restrict__ done, float* xf, float* xnf, float* xa, float* xr, float* xd) {
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (b >= B) return;
    const long long row = idx[b];
    for (int c = lane; c < O; c += 32) { xf[(size_t)b * O + c] = feat[(size_t)row * O + c]; xnf[(size_t)b * O + c] = nfeat[(size_t)row * O + c]; }
    for (int c = lane; c < A; c += 32) xa[(size_t)b * A + c] = act[(size_t)row * A + c];
    if (lane == 0) { xr[b] = rew[row]; xd[b] = done[row] ? 1.f : 0.f; }
}

template <typename T>
__global__ void im2col_frames_kernel(const T* __restrict__ frames, const long long* __restrict__ idx, float* __restrict__ col,
                                     int B, int C, int H, int W, int KH, int S, int OH, int OW) {
    // one thread per (output pixel, c, ky): writes KW consecutive column entries from KW consecutive source pixels
    const int K = C * KH * KH;
    const long long total = (long long)B * OH * OW * C * KH;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int ky = (int)(i % KH); long long r = i / KH;
        const int c = (int)(r % C); r /= C;
        const int ox = (int)(r % OW); r /= OW;
        const int oy = (int)(r % OH); const int b = (int)(r / OH);
        const long long frow = idx ? idx[b] : b;
        const T* src = frames + ((size_t)frow * C + c) * H * W + (size_t)(oy * S + ky) * W + ox * S;
        float* dst = col + ((size_t)(b * OH + oy) * OW + ox) * K + (c * KH + ky) * KH;
        for (int kx = 0; kx < KH; ++kx) dst[kx] = (float)src[kx];
    }
}

// im2col of an activation in [B*IH*IW, C] layout.  Column order (ky, kx, c): for a fixed output pixel and ky the KW*C source
// floats are ONE contiguous run of the activation, so the whole transform is coalesced 16-byte copies.  (The weights of
// conv layers 1 and 2 are stored in the matching [F, KH, KW, C] order in the arena; the Python modules expose them as
// permuted views with the reference's [F, C, KH, KW] shape.)
__global__ void im2col_act_kernel(const float* __restrict__ act, float* __restrict__ col, int B, int C, int IH, int IW, int KH, int S,
                                  int OH, int OW) {
    const int seg = KH * C, K = KH * seg;
    const int V = (C & 3) ? 1 : 4, segv = seg / V;
    const long long total = (long long)B * OH * OW * KH * segv;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int q = (int)(i % segv); long long r = i / segv;
        const int ky = (int)(r % KH); const long long pix = r / KH;
        const int ox = (int)(pix % OW); const long long r2 = pix / OW;
        const int oy = (int)(r2 % OH); const int b = (int)(r2 / OH);
        const float* src = act + ((size_t)(b * IH + oy * S + ky) * IW + ox * S) * C + (size_t)q * V;
        float* dst = col + (size_t)pix * K + (size_t)ky * seg + (size_t)q * V;
        if (V == 4) *reinterpret_cast<float4*>(dst) = *reinterpret_cast<const float4*>(src);
        else *dst = *src;
    }
}

// col2im as a gather + ReLU mask of the layer input: dact[b,iy,ix,c] = [act > 0] * sum over the patches that cover the pixel
// (same (ky, kx, c) column order: consecutive channels are consecutive in dcol)
__global__ void col2im_mask_kernel(const float* __restrict__ dcol, const float* __restrict__ act, float* __restrict__ dact, int B, int C,
                                   int IH, int IW, int KH, int S, int OH, int OW) {
    const int K = C * KH * KH;
    const int V = (C & 3) ? 1 : 4, Cv = C / V;
    const long long total = (long long)B * IH * IW * Cv;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int cq = (int)(i % Cv); long long r = i / Cv;
        const int ix = (int)(r % IW); r /= IW;
        const int iy = (int)(r % IH); const int b = (int)(r / IH);
        const size_t e = ((size_t)(b * IH + iy) * IW + ix) * C + (size_t)cq * V;
        float s[4] = {0.f, 0.f, 0.f, 0.f};
        for (int ky = iy % S; ky < KH; ky += S) {
            const int oy = (iy - ky) / S;
            if (iy - ky < 0 || oy >= OH) continue;
            for (int kx = ix % S; kx < KH; kx += S) {
                const int ox = (ix - kx) / S;
                if (ix - kx < 0 || ox >= OW) continue;
                const float* p = dcol + ((size_t)(b * OH + oy) * OW + ox) * K + (size_t)(ky * KH + kx) * C + (size_t)cq * V;
                if (V == 4) { const float4 v = *reinterpret_cast<const float4*>(p); s[0] += v.x; s[1] += v.y; s[2] += v.z; s[3] += v.w; }
                else s[0] += *p;
            }
        }
        if (V == 4) {
            const float4 a = *reinterpret_cast<const float4*>(act + e);
            *reinterpret_cast<float4*>(dact + e) = make_float4(a.x > 0.f ? s[0] : 0.f, a.y > 0.f ? s[1] : 0.f, a.z > 0.f ? s[2] : 0.f, a.w > 0.f ? s[3] : 0.f);
        } else {
            dact[e] = act[e] > 0.f ? s[0] : 0.f;
        }
    }
}

// Flatten order: [B, P, F] (what the conv GEMM writes) <-> [B, F*P] (NCHW flatten, convolutional.py:45).  back = 1 also applies
// the ReLU mask of the forward activation: dact[b,p,f] = [act > 0] * dflat[b, f*P + p]
__global__ void permute_flat_kernel(const float* __restrict__ src, const float* __restrict__ act, float* __restrict__ dst, int B, int P, int F,
                                    int back) {
    const long long total = (long long)B * P * F;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int f = (int)(i % F); const long long r = i / F;
        const int p = (int)(r % P); const int b = (int)(r / P);
        const size_t j = ((size_t)b * F + f) * P + p;
        if (!back) dst[j] = src[i];
        else dst[i] = act[i] > 0.f ? src[j] : 0.f;
    }
}

// column sums in fixed order: out[n] = sum_m X[m*ld + n]  (bias gradients); one warp-wide strip of columns per CTA row group.
// gridDim.y > 1: row block y of `rows_per_block` rows writes its partial sums to out + y*part_stride (a second launch of the
// same kernel over the [gridDim.y, N] partials finishes the sum) -- a conv bias gradient sums 100k rows.
__global__ void colsum_kernel(const float* __restrict__ X, float* __restrict__ out, int M, int N, int ld, long long sX, long long sOut,
                              int rows_per_block = 0, long long part_stride = 0) {
    __shared__ float red[8][33];
    const int z = blockIdx.z;
    const int n = blockIdx.x * 32 + (threadIdx.x & 31), rg = threadIdx.x >> 5;      // 8 row groups x 32 columns
    const float* x = X + (size_t)z * sX;
    const int mb = rows_per_block ? blockIdx.y * rows_per_block : 0, me = rows_per_block ? min(M, mb + rows_per_block) : M;
    out += (size_t)blockIdx.y * part_stride;
    float s = 0.f;
    if (n < N) {
        int m = mb + rg;
        for (; m + 24 < me; m += 32) {                  // 4 independent loads in flight
            const float a0 = x[(size_t)m * ld + n], a1 = x[(size_t)(m + 8) * ld + n], a2 = x[(size_t)(m + 16) * ld + n], a3 = x[(size_t)(m + 24) * ld + n];
            s += (a0 + a1) + (a2 + a3);
        }
        for (; m < me; m += 8) s += x[(size_t)m * ld + n];
    }
    red[rg][threadIdx.x & 31] = s;
    __syncthreads();
    if (rg == 0 && n < N) {
        float t = 0.f;
#pragma unroll
        for (int r = 0; r < 8; ++r) t += red[r][threadIdx.x & 31];
        out[(size_t)z * sOut + n] = t;
    }
}

// ------------------------------------------------------------------------------------------------------------
// heads of simple_cnn: c[b] = d[b,:] . wf + bf (forward, convolutional.py:49);  backward of the same dot
// ------------------------------------------------------------------------------------------------------------
__global__ void rowdot_kernel(const float* __restrict__ X, int ld, const float* __restrict__ w, const float* __restrict__ b, float* __restrict__ out,
                              int out_stride, int M, int K, long long sX, long long sW, long long sB, long long sOut) {
    const int z = blockIdx.z;
    const int m = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (m >= M) return;
    const float* x = X + (size_t)z * sX + (size_t)m * ld;
    const float* ww = w + (size_t)z * sW;
    float s = 0.f;
    for (int k = lane; k < K; k += 32) s = fmaf(x[k], ww[k], s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) out[(size_t)z * sOut + (size_t)m * out_stride] = s + b[(size_t)z * sB];
}
// dd[b,k] = dc[b] * wf[k]
__global__ void outer_kernel(const float* __restrict__ dc, int dc_stride, const float* __restrict__ w, float* __restrict__ out, int M, int K,
                             long long sDc, long long sW, long long sOut) {
    const int z = blockIdx.z;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)M * K) return;
    const int m = (int)(i / K), k = (int)(i - (long long)m * K);
    out[(size_t)z * sOut + i] = dc[(size_t)z * sDc + (size_t)m * dc_stride] * w[(size_t)z * sW + k];
}
// dw[k] = sum_b dc[b] * d[b,k]; db = sum_b dc[b]        (fixed order)
__global__ void wdot_grad_kernel(const float* __restrict__ dc, int dc_stride, const float* __restrict__ D, int ld, float* __restrict__ dw,
                                 float* __restrict__ db, int M, int K, long long sDc, long long sD, long long sDw, long long sDb) {
    const int z = blockIdx.z;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const float* dcz = dc + (size_t)z * sDc;
    if (k < K) {
        float s = 0.f;
        for (int m = 0; m < M; ++m) s = fmaf(dcz[(size_t)m * dc_stride], D[(size_t)z * sD + (size_t)m * ld + k], s);
        dw[(size_t)z * sDw + k] = s;
    }
    if (k == 0) {
        float s = 0.f;
        for (int m = 0; m < M; ++m) s += dcz[(size_t)m * dc_stride];
        db[(size_t)z * sDb] = s;
    }
}
// elementwise ReLU mask: dst = [act > 0] * src
__global__ void relu_mask_kernel(const float* __restrict__ src, int lds, const float* __restrict__ act, int lda, float* __restrict__ dst, int ldd,
                                 int M, int N, long long sSrc, long long sAct, long long sDst) {
    const int z = blockIdx.z;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)M * N) return;
    const int m = (int)(i / N), n = (int)(i - (long long)m * N);
    dst[(size_t)z * sDst + (size_t)m * ldd + n] = act[(size_t)z * sAct + (size_t)m * lda + n] > 0.f ? src[(size_t)z * sSrc + (size_t)m * lds + n] : 0.f;
}
__global__ void copy_cols_kernel(const float* __restrict__ src, int lds, float* __restrict__ dst, int ldd, int M, int N) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)M * N) return;
    const int m = (int)(i / N), n = (int)(i - (long long)m * N);
    dst[(size_t)m * ldd + n] = src[(size_t)m * lds + n];
}

// ------------------------------------------------------------------------------------------------------------
// row math of the visual nets (one warp per batch row)
// ------------------------------------------------------------------------------------------------------------
struct Scal { float alpha, gamma, reward_scale, log_min_std, log_max_std, act_limit; unsigned flags; int B, A; };
__device__ __forceinline__ float softplusf(float x) { return x > 20.f ? x : log1pf(expf(x)); }

// tanh-Gaussian heads (convolutional.py:103-121): hm = [mu | log_std_raw] rows (pitch ldh), eps [B,A] or null (deterministic)
//   act_out[b*ldo + a] = tanh(u) * act_limit ; logp[b]
__global__ void heads_rows_kernel(const float* __restrict__ hm, int ldh, const float* __restrict__ eps, float* __restrict__ act_out, int ldo,
                                  float* __restrict__ logp, Scal sc, int rows) {
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (b >= rows) return;
    float sl = 0.f, scr = 0.f;
    for (int a = lane; a < sc.A; a += 32) {
        const float mu = hm[(size_t)b * ldh + a], lsr = hm[(size_t)b * ldh + sc.A + a];
        const float ls = fminf(fmaxf(lsr, sc.log_min_std), sc.log_max_std);
        const float std = expf(ls);
        const float u = eps ? __fadd_rn(mu, __fmul_rn(eps[(size_t)b * sc.A + a], std)) : mu;
        const float th = tanhf(u);
        act_out[(size_t)b * ldo + a] = __fmul_rn(th, sc.act_limit);
        const float zc = __fsub_rn(u, mu), var = __fmul_rn(std, std);
        sl += __fsub_rn(__fsub_rn(-__fdiv_rn(__fmul_rn(zc, zc), __fmul_rn(2.f, var)), logf(std)), 0.918938533204672742f);
        const float sp = softplusf(-2.f * u);
        scr += (sc.flags & 1u) ? 2.f * (0.693147180559945309f - u - sp) : __fsub_rn(__fsub_rn(1.386294361119890618f, u), sp);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { sl += __shfl_xor_sync(0xffffffffu, sl, o); scr += __shfl_xor_sync(0xffffffffu, scr, o); }
    if (lane == 0 && logp) logp[b] = __fsub_rn(sl, scr);
}

// VisualCritic tail (convolutional.py:156-162): q = wf[0]*relu(m) + wf[1]*c + bf   for both critics (z = blockIdx.y)
__global__ void critic_tail_kernel(const float* __restrict__ m, const float* __restrict__ c, const float* __restrict__ wf, const float* __restrict__ bf,
                                   float* __restrict__ q, int B, long long sM, long long sC, long long sW, long long sQ) {
    const int z = blockIdx.y, b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const float mv = fmaxf(m[(size_t)z * sM + b], 0.f);
    q[(size_t)z * sQ + b] = fmaf(wf[(size_t)z * sW], mv, fmaf(wf[(size_t)z * sW + 1], c[(size_t)z * sC + b], bf[(size_t)z * sW]));
}

// critic loss rows (sac/algorithm.py:63-72): backup, dq_z = 2 (q_z - backup) / B ; dm_z = dq_z * wf_z[0] * [m_z > 0] ; dc_z = dq_z * wf_z[1]
// per-row loss terms -> lrow[b][4] = (e1^2, e2^2, dq1, dq2)
__global__ void critic_loss_rows_kernel(const float* __restrict__ q, const float* __restrict__ qt, const float* __restrict__ logp,
                                        const float* __restrict__ rew, const float* __restrict__ done, const float* __restrict__ m,
                                        const float* __restrict__ wf, float* __restrict__ backup_out, float* __restrict__ dm, float* __restrict__ dc,
                                        float* __restrict__ dq_out, float* __restrict__ lrow, Scal sc, long long sW) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x, B = sc.B;
    if (b >= B) return;
    const float qtm = fminf(qt[b], qt[B + b]);
    const float backup = __fadd_rn(__fmul_rn(sc.reward_scale, rew[b]),
                                   __fmul_rn(__fmul_rn(sc.gamma, __fsub_rn(1.f, done[b])), __fsub_rn(qtm, __fmul_rn(sc.alpha, logp[b]))));
    backup_out[b] = backup;
    float e[2];
#pragma unroll
    for (int z = 0; z < 2; ++z) {
        e[z] = __fsub_rn(q[(size_t)z * B + b], backup);
        const float dq = 2.f * e[z] / (float)B;
        dq_out[(size_t)z * B + b] = dq;
        dm[(size_t)z * B + b] = m[(size_t)z * B + b] > 0.f ? dq * wf[(size_t)z * sW] : 0.f;
        dc[(size_t)z * B + b] = dq * wf[(size_t)z * sW + 1];
        lrow[(size_t)b * 4 + 2 + z] = dq;
    }
    lrow[(size_t)b * 4 + 0] = e[0] * e[0]; lrow[(size_t)b * 4 + 1] = e[1] * e[1];
}
// d loss / d final.{weight,bias} of both critics + the loss itself: one CTA of 256 threads, 8 quantities, fixed-order tree
__global__ void critic_final_grads_kernel(const float* __restrict__ lrow, const float* __restrict__ m, const float* __restrict__ c, float* __restrict__ gw,
                                          float* __restrict__ gb, float* __restrict__ loss_out, int B, long long sW) {
    __shared__ float red[8][256];
    float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};       // z0: g0 g1 gb | z1: g0 g1 gb | l0 l1
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
#pragma unroll
        for (int z = 0; z < 2; ++z) {
            const float dq = lrow[(size_t)b * 4 + 2 + z];
            v[3 * z + 0] = fmaf(dq, fmaxf(m[(size_t)z * B + b], 0.f), v[3 * z + 0]);
            v[3 * z + 1] = fmaf(dq, c[(size_t)z * B + b], v[3 * z + 1]);
            v[3 * z + 2] += dq;
        }
        v[6] += lrow[(size_t)b * 4]; v[7] += lrow[(size_t)b * 4 + 1];
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) red[q][threadIdx.x] = v[q];
    __syncthreads();
    for (int o = blockDim.x >> 1; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o)
#pragma unroll
            for (int q = 0; q < 8; ++q) red[q][threadIdx.x] += red[q][threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        for (int z = 0; z < 2; ++z) { gw[(size_t)z * sW] = red[3 * z][0]; gw[(size_t)z * sW + 1] = red[3 * z + 1][0]; gb[(size_t)z * sW] = red[3 * z + 2][0]; }
        loss_out[0] = __fadd_rn(red[6][0] / (float)B, red[7][0] / (float)B);
    }
}

// actor loss rows (sac/algorithm.py:40-41): loss terms, routing of d min(q1,q2) into dm_z = dq_z * wf_z[0] * [m_z > 0]
__global__ void actor_loss_rows_kernel(const float* __restrict__ q, const float* __restrict__ logp, const float* __restrict__ m,
                                       const float* __restrict__ wf, float* __restrict__ dm, float* __restrict__ lrow, Scal sc, long long sW) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x, B = sc.B;
    if (b >= B) return;
    const float q1 = q[b], q2 = q[B + b], gq = -1.f / (float)B;
    const float dq1 = q1 < q2 ? gq : (q1 == q2 ? 0.5f * gq : 0.f), dq2 = q2 < q1 ? gq : (q1 == q2 ? 0.5f * gq : 0.f);
    dm[b] = m[b] > 0.f ? dq1 * wf[0] : 0.f;
    dm[(size_t)B + b] = m[(size_t)B + b] > 0.f ? dq2 * wf[sW] : 0.f;
    lrow[b] = __fsub_rn(__fmul_rn(sc.alpha, logp[b]), fminf(q1, q2));
}
__global__ void mean_kernel(const float* __restrict__ x, float* __restrict__ out, int B) {      // one CTA of 256 threads, fixed-order tree
    __shared__ float red[256];
    float s = 0.f;
    for (int b = threadIdx.x; b < B; b += blockDim.x) s += x[b];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = blockDim.x >> 1; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = red[0] / (float)B;
}
// gradient w.r.t. [mu | log_std_raw] (the MLP engine's row_ghead_tile, same formulas): dpi [B, ldp] (first A columns)
__global__ void ghead_rows_kernel(const float* __restrict__ hm, int ldh, const float* __restrict__ eps, const float* __restrict__ dpi, int ldp,
                                  float* __restrict__ g, Scal sc) {
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (b >= sc.B) return;
    const float glogp = sc.alpha / (float)sc.B;
    for (int a = lane; a < sc.A; a += 32) {
        const float mu = hm[(size_t)b * ldh + a], lsr = hm[(size_t)b * ldh + sc.A + a];
        const float ls = fminf(fmaxf(lsr, sc.log_min_std), sc.log_max_std), std = expf(ls);
        const float e = eps ? eps[(size_t)b * sc.A + a] : 0.f;
        const float u = eps ? __fadd_rn(mu, __fmul_rn(e, std)) : mu;
        const float th = tanhf(u);
        const float dcorr = (sc.flags & 1u) ? 2.f * th : th;
        const float gu = dpi[(size_t)b * ldp + a] * sc.act_limit * (1.f - th * th) + glogp * dcorr;
        const bool inside = (lsr >= sc.log_min_std) && (lsr <= sc.log_max_std);
        g[(size_t)b * ldh + a] = gu;
        g[(size_t)b * ldh + sc.A + a] = inside ? (gu * e * std - glogp) : 0.f;
    }
}

// ------------------------------------------------------------------------------------------------------------
// Adam (+ optional Polyak of the target) over a flat arena -- the arithmetic of sacb::adam_elem (torch _single_tensor_adam)
// ------------------------------------------------------------------------------------------------------------
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, float* __restrict__ tgt,
                            long long n, float w1, float b2, float w2, float bc2s, float eps, float nstep, float pk, float pc) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float gi = g[i];
        float mi = m[i], vi = v[i], pi = p[i];
        mi = __fadd_rn(mi, __fmul_rn(w1, __fsub_rn(gi, mi)));
        vi = __fadd_rn(__fmul_rn(vi, b2), __fmul_rn(__fmul_rn(w2, gi), gi));
        const float denom = __fadd_rn(__fdiv_rn(sqrtf(vi), bc2s), eps);
        pi = __fadd_rn(pi, __fdiv_rn(__fmul_rn(nstep, mi), denom));
        p[i] = pi; m[i] = mi; v[i] = vi;
        if (tgt) tgt[i] = __fadd_rn(__fmul_rn(tgt[i], pk), __fmul_rn(pc, pi));
    }
}
__global__ void polyak_kernel(float* __restrict__ tgt, const float* __restrict__ src, long long n, float pk, float pc) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        tgt[i] = __fadd_rn(__fmul_rn(tgt[i], pk), __fmul_rn(pc, src[i]));
}

}  // namespace sacv
