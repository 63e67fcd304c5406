// sacv_engine.cu -- host side of the VISUAL SAC engine: C ABI `sacv_*` of include/sacb200.h.
//
// Replaces, for BASELINE configs[4], the reference's visual path: buffer/visual_replay_buffer.py (VisualReplayBuffer.store /
// sample, VisualBatch), networks/convolutional.py (simple_cnn, VisualActor, VisualCritic, VisualDoubleCritic) and the update
// sac/algorithm.py:30-81,115-162 driven on VisualBatches.  One update = sample -> update_critic -> update_policy ->
// update_targets, a stream of tensor-core GEMM / elementwise launches (kernels in sacv_kernels.cuh).
#include "sacv_kernels.cuh"
#include "sacb_rng.cuh"

#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/sacb200.h"

using namespace sacv;

static thread_local std::string gv_err;
static int vfail(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    gv_err = buf;
    return code;
}
#define CUV(call)                                                                                         \
    do {                                                                                                  \
        cudaError_t e_ = (call);                                                                          \
        if (e_ != cudaSuccess) return vfail(SACB_ERR_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

static inline long long r4(long long x) { return (x + 3) & ~3LL; }

// buffers of one CNN evaluation (nz nets side by side; every array is [nz][...] with the listed per-net size)
struct CnnBufs {
    int nz = 1;
    float *a1 = nullptr, *col2 = nullptr, *a2 = nullptr, *col3 = nullptr, *a3 = nullptr, *a3n = nullptr, *d = nullptr, *c = nullptr;
};

struct sacv_engine {
    sacv_config cfg{};
    sacv_layout_info lay{};
    int P[3], K[3], Cin[3], IH[3], IW[3];       // per conv layer: output pixels, reduction length, input channels / size
    // bound arenas
    float *pa = nullptr, *pc = nullptr, *pt = nullptr, *ma = nullptr, *va = nullptr, *mc = nullptr, *vc = nullptr;
    long long step_pi = 0, step_q = 0;
    // replay ring
    float *rf = nullptr, *rnf = nullptr, *ra_ = nullptr, *rr = nullptr; uint8_t *rfr = nullptr, *rnfr = nullptr, *rd = nullptr;
    long long cap = 0, ptr = 0, size = 0;
    // workspace
    std::vector<void*> allocs;
    float *xf, *xnf, *xa, *xr, *xd, *xs, *xp, *x2, *col1_s, *col1_n;
    int ldx = 0, ldxh = 0, ldh = 0;
    CnnBufs cq, cs, ca;                          // critics on s (kept for backward), scratch twin, actor on s'
    float *ah1, *xh, *hm, *logp, *ch1, *ch2, *cm, *qv, *qt, *qpi, *backup, *dq, *dm, *dcq, *lrow, *lrow_pi, *loss;
    float *dh2, *dh1, *dxs, *dpi, *g, *dxh, *adh2, *adh1, *dd, *da3n, *da3, *dcol3, *da2, *dcol2, *da1, *partial;
    size_t partial_floats = 0;
    float *ga = nullptr, *gc = nullptr;
    long long* idx_buf = nullptr; int* perm_buf = nullptr; float* eps_buf = nullptr;
    unsigned long long rng_counter = 0;
    uint8_t* stage = nullptr; size_t stage_bytes = 0; cudaEvent_t stage_ev = nullptr;
    long long launches = 0;
    bool batch_loaded = false;
};

// ------------------------------------------------------------------------------------------------------------
// layout
// ------------------------------------------------------------------------------------------------------------
static int check_vcfg(const sacv_config* c) {
    if (!c) return vfail(SACB_ERR_INVALID, "null config");
    if (c->struct_size != sizeof(sacv_config)) return vfail(SACB_ERR_INVALID, "sacv_config size mismatch (ABI): got %u want %zu", c->struct_size, sizeof(sacv_config));
    if (c->obs_dim < 1 || c->act_dim < 1 || c->hidden1 < 1 || c->hidden2 < 1 || c->dense < 1) return vfail(SACB_ERR_INVALID, "dims must be positive");
    if (c->batch_size < 1 || c->batch_size > 16384) return vfail(SACB_ERR_INVALID, "batch_size out of range (1..16384)");
    int h = c->vis_h, w = c->vis_w;
    if (c->vis_c < 1 || h < 1 || w < 1) return vfail(SACB_ERR_INVALID, "bad frame shape");
    for (int i = 0; i < 3; ++i) {
        if (c->filters[i] < 1 || c->kernels[i] < 1 || c->strides[i] < 1) return vfail(SACB_ERR_INVALID, "bad conv layer %d", i);
        h = (h - c->kernels[i]) / c->strides[i] + 1; w = (w - c->kernels[i]) / c->strides[i] + 1;
        if (h < 1 || w < 1 || (c->kernels[i] > (i ? 1 << 30 : c->vis_h))) return vfail(SACB_ERR_INVALID, "frame too small for conv layer %d", i);
    }
    return 0;
}

static void compute_vlayout(const sacv_config& c, sacv_layout_info& L) {
    memset(&L, 0, sizeof L);
    int h = c.vis_h, w = c.vis_w, cin = c.vis_c;
    long long o = 0;
    for (int i = 0; i < 3; ++i) {
        L.cnn_w[i] = o; o = r4(o + (long long)c.filters[i] * cin * c.kernels[i] * c.kernels[i]);
        L.cnn_b[i] = o; o = r4(o + c.filters[i]);
        h = (h - c.kernels[i]) / c.strides[i] + 1; w = (w - c.kernels[i]) / c.strides[i] + 1;
        L.conv_h[i] = h; L.conv_w[i] = w; cin = c.filters[i];
    }
    L.flat = c.filters[2] * h * w;
    L.cnn_lw = o; o = r4(o + (long long)c.dense * L.flat);
    L.cnn_lb = o; o = r4(o + c.dense);
    L.cnn_fw = o; o = r4(o + c.dense);
    L.cnn_fb = o; o = r4(o + 1);
    L.cnn_floats = o;
    const long long O = c.obs_dim, A = c.act_dim, H1 = c.hidden1, H2 = c.hidden2;
    o = 0;
    L.a_l0w = o; o = r4(o + H1 * O); L.a_l0b = o; o = r4(o + H1);
    L.a_l1w = o; o = r4(o + H2 * H1); L.a_l1b = o; o = r4(o + H2);
    L.a_cnn = o; o += L.cnn_floats;
    L.a_muw = o; L.a_lsw = o + A * (H2 + 1); o = r4(o + 2 * A * (H2 + 1));      // [mu ; log_std] = one [2A, H2+1] matrix
    L.a_mub = o; L.a_lsb = o + A; o = r4(o + 2 * A);
    L.actor_floats = o;
    o = 0;
    L.c_l0w = o; o = r4(o + H1 * (O + A)); L.c_l0b = o; o = r4(o + H1);
    L.c_l1w = o; o = r4(o + H2 * H1); L.c_l1b = o; o = r4(o + H2);
    L.c_l2w = o; o = r4(o + H2); L.c_l2b = o; o = r4(o + 1);
    L.c_fw = o; o = r4(o + 2); L.c_fb = o; o = r4(o + 1);
    L.c_cnn = o; o += L.cnn_floats;
    L.critic_net_floats = o;
    L.critic_floats = 2 * o;
}

// ------------------------------------------------------------------------------------------------------------
// launch helpers
// ------------------------------------------------------------------------------------------------------------
template <int BN, bool AK, bool BK>
static void launch_gemm(const GemmArgs& g, int tiles, int splits, int batch, cudaStream_t st) {
    static bool attr[64] = {false};                       // the attribute is per device
    constexpr int smem = gemm_smem_floats<BN>() * 4;
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attr[dev & 63]) { cudaFuncSetAttribute(gemm_kernel<BN, AK, BK>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); attr[dev & 63] = true; }
    gemm_kernel<BN, AK, BK><<<dim3(tiles, splits, batch), kGThreads, smem, st>>>(g);
}

// C[M,N] = epi(A(m,k) B(n,k)); ak / bk: the operand is stored k-contiguous.  batch = 1 or 2 (strides sA.. in floats)
static int gemm(sacv_engine* e, cudaStream_t st, bool ak, bool bk, const float* A, int lda, long long sA, const float* B, int ldb, long long sB,
                float* C, int ldc, long long sC, const float* bias, long long sBias, int M, int N, int K, int epi, int batch = 1) {
    GemmArgs g;
    memset(&g, 0, sizeof g);
    g.A = A; g.B = B; g.C = C; g.bias = bias; g.M = M; g.N = N; g.K = K; g.lda = lda; g.ldb = ldb; g.ldc = ldc; g.epi = epi;
    g.x3 = e->cfg.precision != 0; g.sA = sA; g.sB = sB; g.sC = sC; g.sBias = sBias;
    const int bn = N <= 32 ? 32 : 64;
    const int tiles = ((M + kGM - 1) / kGM) * ((N + bn - 1) / bn);
    // split-K for the skinny weight-gradient products (tiny output, K = batch*pixels): fill ~2 waves of SMs, fixed-order reduce
    int splits = 1;
    if (tiles * batch < 96 && K >= 1024) {
        splits = (296 + tiles * batch - 1) / (tiles * batch);
        const int max_s = K / 256;
        if (splits > max_s) splits = max_s;
        if (splits < 1) splits = 1;
        while (splits > 1 && (size_t)splits * batch * M * ldc > e->partial_floats) --splits;
    }
    int kps = (K + splits - 1) / splits;
    kps = (kps + kGK - 1) / kGK * kGK;
    splits = (K + kps - 1) / kps;
    g.k_per_split = kps;
    float* Cfinal = C;
    if (splits > 1) { g.C = e->partial; g.sC = (long long)splits * M * ldc; g.split_stride = (long long)M * ldc; }
#define GL(BN_, AK_, BK_) launch_gemm<BN_, AK_, BK_>(g, tiles, splits, batch, st)
    if (bn == 32) { if (ak) { if (bk) GL(32, true, true); else GL(32, true, false); } else { if (bk) GL(32, false, true); else GL(32, false, false); } }
    else { if (ak) { if (bk) GL(64, true, true); else GL(64, true, false); } else { if (bk) GL(64, false, true); else GL(64, false, false); } }
#undef GL
    CUV(cudaGetLastError());
    e->launches++;
    if (splits > 1) {
        const long long n = (long long)M * N;
        splitk_reduce_kernel<<<dim3((unsigned)((n + 255) / 256), 1, batch), 256, 0, st>>>(e->partial, Cfinal, bias, M, N, ldc, splits, g.split_stride, epi,
                                                                                          g.sC, sC, sBias);
        CUV(cudaGetLastError());
        e->launches++;
    }
    return 0;
}

// column sums of X [M, N] (pitch ld) for nz batched matrices; tall matrices go through per-row-block partials (fixed order)
static int colsum(sacv_engine* e, cudaStream_t st, const float* X, float* out, int M, int N, int ld, long long sX, long long sOut, int nz);

static inline unsigned nblk(long long n, int per = 256) { long long b = (n + per - 1) / per; return (unsigned)(b > 65535LL * 16 ? 65535LL * 16 : b); }
#define KCHK() do { CUV(cudaGetLastError()); e->launches++; } while (0)

// simple_cnn forward for nz nets whose parameter blocks start at P (+ z*sP): col1 -> a1 .. c   (convolutional.py:32-51)
static int cnn_forward(sacv_engine* e, cudaStream_t st, const float* P, long long sP, const float* col1, CnnBufs& b, float* c_out, int c_stride,
                       long long c_zstride) {
    const sacv_config& c = e->cfg; const sacv_layout_info& L = e->lay;
    const int B = c.batch_size, nz = b.nz;
    const int F0 = c.filters[0], F1 = c.filters[1], F2 = c.filters[2], D = c.dense;
    const long long M1 = (long long)B * e->P[0], M2 = (long long)B * e->P[1], M3 = (long long)B * e->P[2];
    if (int r = gemm(e, st, true, true, col1, e->K[0], 0, P + L.cnn_w[0], e->K[0], sP, b.a1, F0, M1 * F0, P + L.cnn_b[0], sP, (int)M1, F0, e->K[0], EPI_BIAS_RELU, nz)) return r;
    for (int z = 0; z < nz; ++z) {
        im2col_act_kernel<<<nblk(M2 * e->K[1]), 256, 0, st>>>(b.a1 + z * M1 * F0, b.col2 + z * M2 * e->K[1], B, F0, L.conv_h[0], L.conv_w[0], c.kernels[1], c.strides[1], L.conv_h[1], L.conv_w[1]);
        KCHK();
    }
    if (int r = gemm(e, st, true, true, b.col2, e->K[1], M2 * e->K[1], P + L.cnn_w[1], e->K[1], sP, b.a2, F1, M2 * F1, P + L.cnn_b[1], sP, (int)M2, F1, e->K[1], EPI_BIAS_RELU, nz)) return r;
    for (int z = 0; z < nz; ++z) {
        im2col_act_kernel<<<nblk(M3 * e->K[2]), 256, 0, st>>>(b.a2 + z * M2 * F1, b.col3 + z * M3 * e->K[2], B, F1, L.conv_h[1], L.conv_w[1], c.kernels[2], c.strides[2], L.conv_h[2], L.conv_w[2]);
        KCHK();
    }
    if (int r = gemm(e, st, true, true, b.col3, e->K[2], M3 * e->K[2], P + L.cnn_w[2], e->K[2], sP, b.a3, F2, M3 * F2, P + L.cnn_b[2], sP, (int)M3, F2, e->K[2], EPI_BIAS_RELU, nz)) return r;
    for (int z = 0; z < nz; ++z) {
        permute_flat_kernel<<<nblk(M3 * F2), 256, 0, st>>>(b.a3 + z * M3 * F2, nullptr, b.a3n + (long long)z * B * L.flat, B, e->P[2], F2, 0);
        KCHK();
    }
    if (int r = gemm(e, st, true, true, b.a3n, L.flat, (long long)B * L.flat, P + L.cnn_lw, L.flat, sP, b.d, D, (long long)B * D, P + L.cnn_lb, sP, B, D, L.flat, EPI_BIAS, nz)) return r;
    rowdot_kernel<<<dim3((B + 7) / 8, 1, nz), 256, 0, st>>>(b.d, D, P + L.cnn_fw, P + L.cnn_fb, c_out, c_stride, B, D, (long long)B * D, sP, sP, c_zstride);
    KCHK();
    return 0;
}

// simple_cnn backward: dc[b] (stride dc_stride) -> gradients into G (+ z*sG), same layout as the parameters
static int cnn_backward(sacv_engine* e, cudaStream_t st, const float* P, float* G, long long sP, const float* col1, CnnBufs& b, const float* dc,
                        int dc_stride, long long dc_zstride) {
    const sacv_config& c = e->cfg; const sacv_layout_info& L = e->lay;
    const int B = c.batch_size, nz = b.nz;
    const int F0 = c.filters[0], F1 = c.filters[1], F2 = c.filters[2], D = c.dense;
    const long long M1 = (long long)B * e->P[0], M2 = (long long)B * e->P[1], M3 = (long long)B * e->P[2];
    // final: c = d . wf + bf
    wdot_grad_kernel<<<dim3((D + 127) / 128, 1, nz), 128, 0, st>>>(dc, dc_stride, b.d, D, G + L.cnn_fw, G + L.cnn_fb, B, D, dc_zstride, (long long)B * D, sP, sP);
    KCHK();
    outer_kernel<<<dim3(nblk((long long)B * D), 1, nz), 256, 0, st>>>(dc, dc_stride, P + L.cnn_fw, e->dd, B, D, dc_zstride, sP, (long long)B * D);
    KCHK();
    // linear: d = a3n . Wl^T + bl
    if (int r = gemm(e, st, false, false, e->dd, D, (long long)B * D, b.a3n, L.flat, (long long)B * L.flat, G + L.cnn_lw, L.flat, sP, nullptr, 0, D, L.flat, B, EPI_NONE, nz)) return r;
    if (int r = colsum(e, st, e->dd, G + L.cnn_lb, B, D, D, (long long)B * D, sP, nz)) return r;
    if (int r = gemm(e, st, true, false, e->dd, D, (long long)B * D, P + L.cnn_lw, L.flat, sP, e->da3n, L.flat, (long long)B * L.flat, nullptr, 0, B, L.flat, D, EPI_NONE, nz)) return r;
    for (int z = 0; z < nz; ++z) {
        permute_flat_kernel<<<nblk(M3 * F2), 256, 0, st>>>(e->da3n + (long long)z * B * L.flat, b.a3 + z * M3 * F2, e->da3 + z * M3 * F2, B, e->P[2], F2, 1);
        KCHK();
    }
    // conv 2
    if (int r = gemm(e, st, false, false, e->da3, F2, M3 * F2, b.col3, e->K[2], M3 * e->K[2], G + L.cnn_w[2], e->K[2], sP, nullptr, 0, F2, e->K[2], (int)M3, EPI_NONE, nz)) return r;
    if (int r = colsum(e, st, e->da3, G + L.cnn_b[2], (int)M3, F2, F2, M3 * F2, sP, nz)) return r;
    if (int r = gemm(e, st, true, false, e->da3, F2, M3 * F2, P + L.cnn_w[2], e->K[2], sP, e->dcol3, e->K[2], M3 * e->K[2], nullptr, 0, (int)M3, e->K[2], F2, EPI_NONE, nz)) return r;
    for (int z = 0; z < nz; ++z) {
        col2im_mask_kernel<<<nblk(M2 * F1), 256, 0, st>>>(e->dcol3 + z * M3 * e->K[2], b.a2 + z * M2 * F1, e->da2 + z * M2 * F1, B, F1, L.conv_h[1], L.conv_w[1], c.kernels[2], c.strides[2], L.conv_h[2], L.conv_w[2]);
        KCHK();
    }
    // conv 1
    if (int r = gemm(e, st, false, false, e->da2, F1, M2 * F1, b.col2, e->K[1], M2 * e->K[1], G + L.cnn_w[1], e->K[1], sP, nullptr, 0, F1, e->K[1], (int)M2, EPI_NONE, nz)) return r;
    if (int r = colsum(e, st, e->da2, G + L.cnn_b[1], (int)M2, F1, F1, M2 * F1, sP, nz)) return r;
    if (int r = gemm(e, st, true, false, e->da2, F1, M2 * F1, P + L.cnn_w[1], e->K[1], sP, e->dcol2, e->K[1], M2 * e->K[1], nullptr, 0, (int)M2, e->K[1], F1, EPI_NONE, nz)) return r;
    for (int z = 0; z < nz; ++z) {
        col2im_mask_kernel<<<nblk(M1 * F0), 256, 0, st>>>(e->dcol2 + z * M2 * e->K[1], b.a1 + z * M1 * F0, e->da1 + z * M1 * F0, B, F0, L.conv_h[0], L.conv_w[0], c.kernels[1], c.strides[1], L.conv_h[1], L.conv_w[1]);
        KCHK();
    }
    // conv 0 (no input gradient: the frames are data)
    if (int r = gemm(e, st, false, false, e->da1, F0, M1 * F0, col1, e->K[0], 0, G + L.cnn_w[0], e->K[0], sP, nullptr, 0, F0, e->K[0], (int)M1, EPI_NONE, nz)) return r;
    if (int r = colsum(e, st, e->da1, G + L.cnn_b[0], (int)M1, F0, F0, M1 * F0, sP, nz)) return r;
    return 0;
}

static int colsum(sacv_engine* e, cudaStream_t st, const float* X, float* out, int M, int N, int ld, long long sX, long long sOut, int nz) {
    const int cb = (N + 31) / 32;
    if (M <= 4096) {
        colsum_kernel<<<dim3(cb, 1, nz), 256, 0, st>>>(X, out, M, N, ld, sX, sOut);
        KCHK();
        return 0;
    }
    int S = (M + 1023) / 1024;
    if (S > 128) S = 128;
    const int rpb = (M + S - 1) / S;
    // partials [nz][S][N] live in the split-K scratch (a GEMM never runs concurrently on this stream)
    colsum_kernel<<<dim3(cb, S, nz), 256, 0, st>>>(X, e->partial, M, N, ld, sX, (long long)S * N, rpb, N);
    KCHK();
    colsum_kernel<<<dim3(cb, 1, nz), 256, 0, st>>>(e->partial, out, S, N, N, (long long)S * N, sOut);
    KCHK();
    return 0;
}

static Scal scal_of(const sacv_engine* e) {
    const sacv_config& c = e->cfg;
    Scal s;
    s.alpha = (float)c.alpha; s.gamma = (float)c.gamma; s.reward_scale = (float)c.reward_scale; s.log_min_std = c.log_min_std;
    s.log_max_std = c.log_max_std; s.act_limit = c.act_limit; s.flags = c.flags; s.B = c.batch_size; s.A = c.act_dim;
    return s;
}

// VisualDoubleCritic MLP forward on X = [features | action] (pitch ldx) with CNN scalars cz [2][B]; activations into h1 / h2 / m
static int critic_mlp_forward(sacv_engine* e, cudaStream_t st, const float* P, const float* X, const float* cz, float* h1, float* h2, float* m, float* q) {
    const sacv_config& c = e->cfg; const sacv_layout_info& L = e->lay;
    const int B = c.batch_size, O = c.obs_dim, A = c.act_dim, H1 = c.hidden1, H2 = c.hidden2;
    const long long sP = L.critic_net_floats;
    if (int r = gemm(e, st, true, true, X, e->ldx, 0, P + L.c_l0w, O + A, sP, h1, H1, (long long)B * H1, P + L.c_l0b, sP, B, H1, O + A, EPI_BIAS_RELU, 2)) return r;
    if (int r = gemm(e, st, true, true, h1, H1, (long long)B * H1, P + L.c_l1w, H1, sP, h2, H2, (long long)B * H2, P + L.c_l1b, sP, B, H2, H1, EPI_BIAS_RELU, 2)) return r;
    rowdot_kernel<<<dim3((B + 7) / 8, 1, 2), 256, 0, st>>>(h2, H2, P + L.c_l2w, P + L.c_l2b, m, 1, B, H2, (long long)B * H2, sP, sP, B);
    KCHK();
    critic_tail_kernel<<<dim3((B + 127) / 128, 2), 128, 0, st>>>(m, cz, P + L.c_fw, P + L.c_fb, q, B, B, B, sP, B);
    KCHK();
    return 0;
}

// dh2 = [h2 > 0] * dm (x) w2
__global__ void outer_mask_kernel(const float* __restrict__ dm, const float* __restrict__ w, const float* __restrict__ h, float* __restrict__ out, int M, int K,
                                  long long sDm, long long sW, long long sH) {
    const int z = blockIdx.z;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)M * K) return;
    const int m = (int)(i / K), k = (int)(i - (long long)m * K);
    out[(size_t)z * sH + i] = h[(size_t)z * sH + i] > 0.f ? dm[(size_t)z * sDm + m] * w[(size_t)z * sW + k] : 0.f;
}
__global__ void add_dpi_kernel(const float* __restrict__ dx, int ldx, int O, int A, float* __restrict__ dpi, int ldp, int B, long long sDx) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * A) return;
    const int b = i / A, a = i - b * A;
    dpi[(size_t)b * ldp + a] = dx[(size_t)b * ldx + O + a] + dx[(size_t)sDx + (size_t)b * ldx + O + a];
}
__global__ void build_x_kernel(const float* __restrict__ f, int O, const float* __restrict__ a, int lda_, int A, float* __restrict__ x, int ldx, int B) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * (O + A)) return;
    const int b = i / (O + A), c = i - b * (O + A);
    x[(size_t)b * ldx + c] = c < O ? f[(size_t)b * O + c] : a[(size_t)b * lda_ + (c - O)];
}

static int adam_arena(sacv_engine* e, cudaStream_t st, float* p, const float* g, float* m, float* v, float* tgt, long long n, double lr, long long t) {
    const sacv_config& c = e->cfg;
    const double bc1 = 1.0 - std::pow(c.beta1, (double)t), bc2 = 1.0 - std::pow(c.beta2, (double)t);
    adam_kernel<<<nblk(n), 256, 0, st>>>(p, g, m, v, tgt, n, (float)(1.0 - c.beta1), (float)c.beta2, (float)(1.0 - c.beta2), (float)std::sqrt(bc2), (float)c.adam_eps,
                                         (float)(-(lr / bc1)), (float)c.polyak, (float)(1.0 - c.polyak));
    KCHK();
    return 0;
}

static int vneed(bool cond, const char* what) { return cond ? 0 : vfail(SACB_ERR_STATE, "%s not bound", what); }

// ---- the three pieces of one update on the batch that sits in the staging tiles -------------------------------------
// actor trunk + CNN on next_states and the two tanh-Gaussian draws (the actor does not change between them)
static int actor_eval(sacv_engine* e, cudaStream_t st, const float* eps1, const float* eps2, int passes) {
    const sacv_config& c = e->cfg; const sacv_layout_info& L = e->lay;
    const int B = c.batch_size, O = c.obs_dim, A = c.act_dim, H1 = c.hidden1, H2 = c.hidden2;
    if (int r = gemm(e, st, true, true, e->xnf, O, 0, e->pa + L.a_l0w, O, 0, e->ah1, H1, 0, e->pa + L.a_l0b, 0, B, H1, O, EPI_BIAS_RELU)) return r;
    if (int r = gemm(e, st, true, true, e->ah1, H1, 0, e->pa + L.a_l1w, H1, 0, e->xh, e->ldxh, 0, e->pa + L.a_l1b, 0, B, H2, H1, EPI_BIAS_RELU)) return r;
    if (int r = cnn_forward(e, st, e->pa + L.a_cnn, 0, e->col1_n, e->ca, e->xh + H2, e->ldxh, 0)) return r;
    if (int r = gemm(e, st, true, true, e->xh, e->ldxh, 0, e->pa + L.a_muw, H2 + 1, 0, e->hm, e->ldh, 0, e->pa + L.a_mub, 0, B, 2 * A, H2 + 1, EPI_BIAS)) return r;
    const Scal sc = scal_of(e);
    if (passes & 1) { heads_rows_kernel<<<(B + 7) / 8, 256, 0, st>>>(e->hm, e->ldh, eps1, e->x2 + O, e->ldx, e->logp, sc, B); KCHK(); }
    if (passes & 2) { heads_rows_kernel<<<(B + 7) / 8, 256, 0, st>>>(e->hm, e->ldh, eps2, e->xp + O, e->ldx, e->logp + B, sc, B); KCHK(); }
    return 0;
}

static int critic_step(sacv_engine* e, cudaStream_t st, float* loss_out, bool polyak_fused) {
    const sacv_config& c = e->cfg; const sacv_layout_info& L = e->lay;
    const int B = c.batch_size, O = c.obs_dim, A = c.act_dim, H1 = c.hidden1, H2 = c.hidden2;
    const long long sP = L.critic_net_floats;
    const Scal sc = scal_of(e);
    // target critics on (s', a2): CNNs on next frames (scratch buffers), MLP on x2 = [f' | a2]
    if (int r = cnn_forward(e, st, e->pt + L.c_cnn, sP, e->col1_n, e->cs, e->dcq, 1, B)) return r;      // dcq reused as scratch for c_target
    if (int r = critic_mlp_forward(e, st, e->pt, e->x2, e->dcq, e->ch1, e->ch2, e->cm, e->qt)) return r;
    // online critics on (s, a)
    if (int r = cnn_forward(e, st, e->pc + L.c_cnn, sP, e->col1_s, e->cq, e->cq.c, 1, B)) return r;
    if (int r = critic_mlp_forward(e, st, e->pc, e->xs, e->cq.c, e->ch1, e->ch2, e->cm, e->qv)) return r;
    critic_loss_rows_kernel<<<(B + 127) / 128, 128, 0, st>>>(e->qv, e->qt, e->logp, e->xr, e->xd, e->cm, e->pc + L.c_fw, e->backup, e->dm, e->dcq, e->dq, e->lrow, sc, sP);
    KCHK();
    critic_final_grads_kernel<<<1, 256, 0, st>>>(e->lrow, e->cm, e->cq.c, e->gc + L.c_fw, e->gc + L.c_fb, e->loss, B, sP);
    KCHK();
    // MLP backward (both critics, z = 2)
    wdot_grad_kernel<<<dim3((H2 + 127) / 128, 1, 2), 128, 0, st>>>(e->dm, 1, e->ch2, H2, e->gc + L.c_l2w, e->gc + L.c_l2b, B, H2, B, (long long)B * H2, sP, sP);
    KCHK();
    outer_mask_kernel<<<dim3(nblk((long long)B * H2), 1, 2), 256, 0, st>>>(e->dm, e->pc + L.c_l2w, e->ch2, e->dh2, B, H2, B, sP, (long long)B * H2);
    KCHK();
    if (int r = gemm(e, st, false, false, e->dh2, H2, (long long)B * H2, e->ch1, H1, (long long)B * H1, e->gc + L.c_l1w, H1, sP, nullptr, 0, H2, H1, B, EPI_NONE, 2)) return r;
    colsum_kernel<<<dim3((H2 + 31) / 32, 1, 2), 256, 0, st>>>(e->dh2, e->gc + L.c_l1b, B, H2, H2, (long long)B * H2, sP);
    KCHK();
    if (int r = gemm(e, st, true, false, e->dh2, H2, (long long)B * H2, e->pc + L.c_l1w, H1, sP, e->dh1, H1, (long long)B * H1, nullptr, 0, B, H1, H2, EPI_NONE, 2)) return r;
    relu_mask_kernel<<<dim3(nblk((long long)B * H1), 1, 2), 256, 0, st>>>(e->dh1, H1, e->ch1, H1, e->dh1, H1, B, H1, (long long)B * H1, (long long)B * H1, (long long)B * H1);
    KCHK();
    if (int r = gemm(e, st, false, false, e->dh1, H1, (long long)B * H1, e->xs, e->ldx, 0, e->gc + L.c_l0w, O + A, sP, nullptr, 0, H1, O + A, B, EPI_NONE, 2)) return r;
    colsum_kernel<<<dim3((H1 + 31) / 32, 1, 2), 256, 0, st>>>(e->dh1, e->gc + L.c_l0b, B, H1, H1, (long long)B * H1, sP);
    KCHK();
    if (int r = cnn_backward(e, st, e->pc + L.c_cnn, e->gc + L.c_cnn, sP, e->col1_s, e->cq, e->dcq, 1, B)) return r;
    if (int r = adam_arena(e, st, e->pc, e->gc, e->mc, e->vc, polyak_fused ? e->pt : nullptr, L.critic_floats, c.lr_q, e->step_q + 1)) return r;
    if (loss_out) CUV(cudaMemcpyAsync(loss_out, e->loss, sizeof(float), cudaMemcpyDeviceToDevice, st));
    return 0;
}

static int policy_step(sacv_engine* e, cudaStream_t st, const float* eps2, float* loss_out) {
    const sacv_config& c = e->cfg; const sacv_layout_info& L = e->lay;
    const int B = c.batch_size, O = c.obs_dim, A = c.act_dim, H1 = c.hidden1, H2 = c.hidden2;
    const long long sP = L.critic_net_floats;
    const Scal sc = scal_of(e);
    // critics (just stepped) on (s, pi): the CNN scalars do not depend on pi but on the new weights -> forward only
    if (int r = cnn_forward(e, st, e->pc + L.c_cnn, sP, e->col1_s, e->cs, e->cs.c, 1, B)) return r;
    if (int r = critic_mlp_forward(e, st, e->pc, e->xp, e->cs.c, e->ch1, e->ch2, e->cm, e->qpi)) return r;
    actor_loss_rows_kernel<<<(B + 127) / 128, 128, 0, st>>>(e->qpi, e->logp + B, e->cm, e->pc + L.c_fw, e->dm, e->lrow_pi, sc, sP);
    KCHK();
    mean_kernel<<<1, 256, 0, st>>>(e->lrow_pi, e->loss + 1, B);
    KCHK();
    // back through the critics' MLPs to the action columns of their input
    outer_mask_kernel<<<dim3(nblk((long long)B * H2), 1, 2), 256, 0, st>>>(e->dm, e->pc + L.c_l2w, e->ch2, e->dh2, B, H2, B, sP, (long long)B * H2);
    KCHK();
    if (int r = gemm(e, st, true, false, e->dh2, H2, (long long)B * H2, e->pc + L.c_l1w, H1, sP, e->dh1, H1, (long long)B * H1, nullptr, 0, B, H1, H2, EPI_NONE, 2)) return r;
    relu_mask_kernel<<<dim3(nblk((long long)B * H1), 1, 2), 256, 0, st>>>(e->dh1, H1, e->ch1, H1, e->dh1, H1, B, H1, (long long)B * H1, (long long)B * H1, (long long)B * H1);
    KCHK();
    if (int r = gemm(e, st, true, false, e->dh1, H1, (long long)B * H1, e->pc + L.c_l0w, O + A, sP, e->dxs, e->ldx, (long long)B * e->ldx, nullptr, 0, B, O + A, H1, EPI_NONE, 2)) return r;
    add_dpi_kernel<<<(B * A + 255) / 256, 256, 0, st>>>(e->dxs, e->ldx, O, A, e->dpi, e->ldh, B, (long long)B * e->ldx);
    KCHK();
    ghead_rows_kernel<<<(B + 7) / 8, 256, 0, st>>>(e->hm, e->ldh, eps2, e->dpi, e->ldh, e->g, sc);
    KCHK();
    // heads: [mu | ls] = xh . Wh^T + bh with xh = [h2 | c]
    if (int r = gemm(e, st, false, false, e->g, e->ldh, 0, e->xh, e->ldxh, 0, e->ga + L.a_muw, H2 + 1, 0, nullptr, 0, 2 * A, H2 + 1, B, EPI_NONE)) return r;
    colsum_kernel<<<dim3((2 * A + 31) / 32, 1, 1), 256, 0, st>>>(e->g, e->ga + L.a_mub, B, 2 * A, e->ldh, 0, 0);
    KCHK();
    if (int r = gemm(e, st, true, false, e->g, e->ldh, 0, e->pa + L.a_muw, H2 + 1, 0, e->dxh, e->ldxh, 0, nullptr, 0, B, H2 + 1, 2 * A, EPI_NONE)) return r;
    relu_mask_kernel<<<dim3(nblk((long long)B * H2), 1, 1), 256, 0, st>>>(e->dxh, e->ldxh, e->xh, e->ldxh, e->adh2, H2, B, H2, 0, 0, 0);
    KCHK();
    if (int r = gemm(e, st, false, false, e->adh2, H2, 0, e->ah1, H1, 0, e->ga + L.a_l1w, H1, 0, nullptr, 0, H2, H1, B, EPI_NONE)) return r;
    colsum_kernel<<<dim3((H2 + 31) / 32, 1, 1), 256, 0, st>>>(e->adh2, e->ga + L.a_l1b, B, H2, H2, 0, 0);
    KCHK();
    if (int r = gemm(e, st, true, false, e->adh2, H2, 0, e->pa + L.a_l1w, H1, 0, e->adh1, H1, 0, nullptr, 0, B, H1, H2, EPI_NONE)) return r;
    relu_mask_kernel<<<dim3(nblk((long long)B * H1), 1, 1), 256, 0, st>>>(e->adh1, H1, e->ah1, H1, e->adh1, H1, B, H1, 0, 0, 0);
    KCHK();
    if (int r = gemm(e, st, false, false, e->adh1, H1, 0, e->xnf, O, 0, e->ga + L.a_l0w, O, 0, nullptr, 0, H1, O, B, EPI_NONE)) return r;
    colsum_kernel<<<dim3((H1 + 31) / 32, 1, 1), 256, 0, st>>>(e->adh1, e->ga + L.a_l0b, B, H1, H1, 0, 0);
    KCHK();
    if (int r = cnn_backward(e, st, e->pa + L.a_cnn, e->ga + L.a_cnn, 0, e->col1_n, e->ca, e->dxh + H2, e->ldxh, 0)) return r;
    if (int r = adam_arena(e, st, e->pa, e->ga, e->ma, e->va, nullptr, L.actor_floats, c.lr_pi, e->step_pi + 1)) return r;
    if (loss_out) CUV(cudaMemcpyAsync(loss_out, e->loss + 1, sizeof(float), cudaMemcpyDeviceToDevice, st));
    return 0;
}

// batch -> staging tiles: rows from explicit arrays (idx == null, frames float or u8) or from the ring at idx
template <typename T>
static int stage_frames(sacv_engine* e, cudaStream_t st, const T* frames, const T* nframes, const long long* idx) {
    const sacv_config& c = e->cfg; const sacv_layout_info& L = e->lay;
    const long long n = (long long)c.batch_size * L.conv_h[0] * L.conv_w[0] * c.vis_c * c.kernels[0];
    im2col_frames_kernel<T><<<nblk(n), 256, 0, st>>>(frames, idx, e->col1_s, c.batch_size, c.vis_c, c.vis_h, c.vis_w, c.kernels[0], c.strides[0], L.conv_h[0], L.conv_w[0]);
    KCHK();
    im2col_frames_kernel<T><<<nblk(n), 256, 0, st>>>(nframes, idx, e->col1_n, c.batch_size, c.vis_c, c.vis_h, c.vis_w, c.kernels[0], c.strides[0], L.conv_h[0], L.conv_w[0]);
    KCHK();
    return 0;
}
static int finish_staging(sacv_engine* e, cudaStream_t st) {
    const sacv_config& c = e->cfg;
    const int B = c.batch_size, O = c.obs_dim, A = c.act_dim;
    build_x_kernel<<<(B * (O + A) + 255) / 256, 256, 0, st>>>(e->xf, O, e->xa, A, A, e->xs, e->ldx, B);      KCHK();
    copy_cols_kernel<<<(B * O + 255) / 256, 256, 0, st>>>(e->xf, O, e->xp, e->ldx, B, O);                    KCHK();
    copy_cols_kernel<<<(B * O + 255) / 256, 256, 0, st>>>(e->xnf, O, e->x2, e->ldx, B, O);                   KCHK();
    e->batch_loaded = true;
    return 0;
}

// ------------------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------------------
extern "C" {

const char* sacv_last_error(void) { return gv_err.c_str(); }

int sacv_layout(const sacv_config* cfg, sacv_layout_info* out) {
    if (int r = check_vcfg(cfg)) return r;
    if (!out) return vfail(SACB_ERR_INVALID, "null out");
    compute_vlayout(*cfg, *out);
    return 0;
}

int sacv_create(const sacv_config* cfg, sacv_handle* out) {
    if (int r = check_vcfg(cfg)) return r;
    if (!out) return vfail(SACB_ERR_INVALID, "null out");
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0) return vfail(SACB_ERR_CUDA, "no CUDA device available (%s); this engine has no CPU fallback", cudaGetErrorString(ce));
    if (cfg->device < 0 || cfg->device >= ndev) return vfail(SACB_ERR_INVALID, "device %d out of range", cfg->device);
    CUV(cudaSetDevice(cfg->device));
    sacv_engine* e = new sacv_engine();
    e->cfg = *cfg;
    compute_vlayout(*cfg, e->lay);
    const sacv_config& c = e->cfg; const sacv_layout_info& L = e->lay;
    int cin = c.vis_c, ih = c.vis_h, iw = c.vis_w;
    for (int i = 0; i < 3; ++i) {
        e->Cin[i] = cin; e->IH[i] = ih; e->IW[i] = iw;
        e->K[i] = cin * c.kernels[i] * c.kernels[i];
        e->P[i] = L.conv_h[i] * L.conv_w[i];
        cin = c.filters[i]; ih = L.conv_h[i]; iw = L.conv_w[i];
    }
    const long long B = c.batch_size, O = c.obs_dim, A = c.act_dim, H1 = c.hidden1, H2 = c.hidden2, D = c.dense;
    e->ldx = (int)r4(O + A); e->ldxh = (int)r4(H2 + 1); e->ldh = (int)r4(2 * A);
    auto alloc = [&](float*& p, long long n) -> int {
        void* q = nullptr;
        cudaError_t er = cudaMalloc(&q, (size_t)(n > 0 ? n : 1) * sizeof(float));
        if (er != cudaSuccess) return vfail(SACB_ERR_CUDA, "cudaMalloc(%lld floats): %s", n, cudaGetErrorString(er));
        cudaMemset(q, 0, (size_t)(n > 0 ? n : 1) * sizeof(float));
        e->allocs.push_back(q); p = (float*)q; return 0; };
#define AL(p, n) do { if (int r_ = alloc(p, (n))) { return r_; } } while (0)
    const long long M1 = B * e->P[0], M2 = B * e->P[1], M3 = B * e->P[2];
    AL(e->xf, B * O); AL(e->xnf, B * O); AL(e->xa, B * A); AL(e->xr, B); AL(e->xd, B);
    AL(e->xs, B * e->ldx); AL(e->xp, B * e->ldx); AL(e->x2, B * e->ldx);
    AL(e->col1_s, M1 * e->K[0]); AL(e->col1_n, M1 * e->K[0]);
    auto alloc_cnn = [&](CnnBufs& b, int nz) -> int {
        b.nz = nz;
        AL(b.a1, nz * M1 * c.filters[0]); AL(b.col2, nz * M2 * e->K[1]); AL(b.a2, nz * M2 * c.filters[1]); AL(b.col3, nz * M3 * e->K[2]);
        AL(b.a3, nz * M3 * c.filters[2]); AL(b.a3n, nz * B * L.flat); AL(b.d, nz * B * D); AL(b.c, nz * B);
        return 0; };
    if (int r = alloc_cnn(e->cq, 2)) return r;
    if (int r = alloc_cnn(e->cs, 2)) return r;
    if (int r = alloc_cnn(e->ca, 1)) return r;
    AL(e->ah1, B * H1); AL(e->xh, B * e->ldxh); AL(e->hm, B * e->ldh); AL(e->logp, 2 * B);
    AL(e->ch1, 2 * B * H1); AL(e->ch2, 2 * B * H2); AL(e->cm, 2 * B); AL(e->qv, 2 * B); AL(e->qt, 2 * B); AL(e->qpi, 2 * B);
    AL(e->backup, B); AL(e->dq, 2 * B); AL(e->dm, 2 * B); AL(e->dcq, 2 * B); AL(e->lrow, 4 * B); AL(e->lrow_pi, B); AL(e->loss, 4);
    AL(e->dh2, 2 * B * H2); AL(e->dh1, 2 * B * H1); AL(e->dxs, 2 * B * e->ldx); AL(e->dpi, B * e->ldh); AL(e->g, B * e->ldh);
    AL(e->dxh, B * e->ldxh); AL(e->adh2, B * H2); AL(e->adh1, B * H1);
    AL(e->dd, 2 * B * D); AL(e->da3n, 2 * B * L.flat); AL(e->da3, 2 * M3 * c.filters[2]); AL(e->dcol3, 2 * M3 * e->K[2]);
    AL(e->da2, 2 * M2 * c.filters[1]); AL(e->dcol2, 2 * M2 * e->K[1]); AL(e->da1, 2 * M1 * c.filters[0]);
    e->partial_floats = (size_t)16 << 20;                  // 64 MB of split-K partials
    AL(e->partial, (long long)e->partial_floats);
    AL(e->ga, L.actor_floats); AL(e->gc, L.critic_floats);
    AL(e->eps_buf, 2 * B * A);
#undef AL
    void* q = nullptr;
    CUV(cudaMalloc(&q, (size_t)B * sizeof(long long))); e->allocs.push_back(q); e->idx_buf = (long long*)q;
    CUV(cudaMalloc(&q, (size_t)4 * B * sizeof(int))); e->allocs.push_back(q); e->perm_buf = (int*)q;
    CUV(cudaFuncSetAttribute(sacb_rng::draw_indices_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * (int)sizeof(long long)));
    CUV(cudaEventCreateWithFlags(&e->stage_ev, cudaEventDisableTiming));
    *out = e;
    return 0;
}

int sacv_destroy(sacv_handle e) {
    if (!e) return 0;
    cudaSetDevice(e->cfg.device);
    cudaDeviceSynchronize();
    for (void* p : e->allocs) cudaFree(p);
    if (e->stage) cudaFreeHost(e->stage);
    if (e->stage_ev) cudaEventDestroy(e->stage_ev);
    delete e;
    return 0;
}

int sacv_bind_params(sacv_handle e, float* actor, float* critic, float* target) {
    if (!e) return vfail(SACB_ERR_INVALID, "null handle");
    e->pa = actor; e->pc = critic; e->pt = target;
    return 0;
}
int sacv_bind_adam(sacv_handle e, float* ma, float* va, int64_t step_a, float* mc, float* vc, int64_t step_c) {
    if (!e) return vfail(SACB_ERR_INVALID, "null handle");
    e->ma = ma; e->va = va; e->mc = mc; e->vc = vc; e->step_pi = step_a; e->step_q = step_c;
    return 0;
}
int sacv_get_adam_steps(sacv_handle e, int64_t* sa, int64_t* sc) {
    if (!e) return vfail(SACB_ERR_INVALID, "null handle");
    if (sa) *sa = e->step_pi;
    if (sc) *sc = e->step_q;
    return 0;
}
int sacv_set_scalar(sacv_handle e, const char* name, double v) {
    if (!e || !name) return vfail(SACB_ERR_INVALID, "null argument");
    sacv_config& c = e->cfg;
    std::string n(name);
    if (n == "alpha") c.alpha = v; else if (n == "gamma") c.gamma = v; else if (n == "polyak") c.polyak = v;
    else if (n == "reward_scale") c.reward_scale = v; else if (n == "lr_pi") c.lr_pi = v; else if (n == "lr_q") c.lr_q = v;
    else if (n == "beta1") c.beta1 = v; else if (n == "beta2") c.beta2 = v; else if (n == "adam_eps") c.adam_eps = v;
    else if (n == "precision") c.precision = (int)v;
    else return vfail(SACB_ERR_INVALID, "unknown scalar '%s'", name);
    return 0;
}

int sacv_bind_buffer(sacv_handle e, float* feat, uint8_t* frames, float* act, float* rew, float* nfeat, uint8_t* nframes, uint8_t* done, int64_t cap) {
    if (!e) return vfail(SACB_ERR_INVALID, "null handle");
    if (!feat || !frames || !act || !rew || !nfeat || !nframes || !done || cap < 1) return vfail(SACB_ERR_INVALID, "bad buffer arrays");
    e->rf = feat; e->rfr = frames; e->ra_ = act; e->rr = rew; e->rnf = nfeat; e->rnfr = nframes; e->rd = done; e->cap = cap;
    if (e->ptr >= cap) e->ptr = 0;
    if (e->size > cap) e->size = cap;
    return 0;
}
int sacv_buffer_state(sacv_handle e, int64_t* ptr, int64_t* size, int64_t* cap) {
    if (!e) return vfail(SACB_ERR_INVALID, "null handle");
    if (ptr) *ptr = e->ptr;
    if (size) *size = e->size;
    if (cap) *cap = e->cap;
    return 0;
}
int sacv_set_buffer_state(sacv_handle e, int64_t ptr, int64_t size) {
    if (!e) return vfail(SACB_ERR_INVALID, "null handle");
    if (ptr < 0 || size < 0 || size > e->cap || (e->cap && ptr >= e->cap)) return vfail(SACB_ERR_INVALID, "ptr/size out of range");
    e->ptr = ptr; e->size = size;
    return 0;
}

// VisualReplayBuffer.store (visual_replay_buffer.py:33-47) for `count` transitions: host rows -> pinned staging -> HBM ring
int sacv_store(sacv_handle e, const float* feat, const uint8_t* frames, const float* act, const float* rew, const float* nfeat,
               const uint8_t* nframes, const uint8_t* done, int64_t count, uintptr_t stream) {
    if (!e) return vfail(SACB_ERR_INVALID, "null handle");
    if (int r = vneed(e->rf != nullptr, "replay buffer")) return r;
    if (count < 0) return vfail(SACB_ERR_INVALID, "negative count");
    cudaStream_t st = (cudaStream_t)stream;
    CUV(cudaSetDevice(e->cfg.device));
    const size_t O = e->cfg.obs_dim, A = e->cfg.act_dim, F = (size_t)e->cfg.vis_c * e->cfg.vis_h * e->cfg.vis_w;
    const size_t row_bytes = (2 * O + A + 1) * sizeof(float) + 2 * F + 1;
    int64_t chunk_rows = (int64_t)((32u << 20) / row_bytes);
    if (chunk_rows < 1) chunk_rows = 1;
    if (!e->stage) { e->stage_bytes = (size_t)chunk_rows * row_bytes + 256; CUV(cudaMallocHost(&e->stage, e->stage_bytes)); }
    int64_t done_rows = 0;
    while (done_rows < count) {
        int64_t n = count - done_rows;
        if (n > chunk_rows) n = chunk_rows;
        if (n > e->cap - e->ptr) n = e->cap - e->ptr;
        CUV(cudaEventSynchronize(e->stage_ev));
        float* so = (float*)e->stage; float* sn = so + n * O; float* sa = sn + n * O; float* sr = sa + n * A;
        uint8_t* sf = (uint8_t*)(sr + n); uint8_t* snf = sf + n * F; uint8_t* sd = snf + n * F;
        memcpy(so, feat + done_rows * O, n * O * sizeof(float)); memcpy(sn, nfeat + done_rows * O, n * O * sizeof(float));
        memcpy(sa, act + done_rows * A, n * A * sizeof(float)); memcpy(sr, rew + done_rows, n * sizeof(float));
        memcpy(sf, frames + done_rows * F, n * F); memcpy(snf, nframes + done_rows * F, n * F);
        for (int64_t i = 0; i < n; ++i) sd[i] = done[done_rows + i] ? 1 : 0;
        CUV(cudaMemcpyAsync(e->rf + e->ptr * O, so, n * O * sizeof(float), cudaMemcpyHostToDevice, st));
        CUV(cudaMemcpyAsync(e->rnf + e->ptr * O, sn, n * O * sizeof(float), cudaMemcpyHostToDevice, st));
        CUV(cudaMemcpyAsync(e->ra_ + e->ptr * A, sa, n * A * sizeof(float), cudaMemcpyHostToDevice, st));
        CUV(cudaMemcpyAsync(e->rr + e->ptr, sr, n * sizeof(float), cudaMemcpyHostToDevice, st));
        CUV(cudaMemcpyAsync(e->rfr + e->ptr * F, sf, n * F, cudaMemcpyHostToDevice, st));
        CUV(cudaMemcpyAsync(e->rnfr + e->ptr * F, snf, n * F, cudaMemcpyHostToDevice, st));
        CUV(cudaMemcpyAsync(e->rd + e->ptr, sd, n, cudaMemcpyHostToDevice, st));
        CUV(cudaEventRecord(e->stage_ev, st));
        e->ptr = (e->ptr + n) % e->cap;                                   // visual_replay_buffer.py:46
        e->size = e->size + n > e->cap ? e->cap : e->size + n;            // :47
        done_rows += n;
    }
    return 0;
}

static int vdraw_indices(sacv_engine* e, long long* out, int B, cudaStream_t st) {
    if (e->size < B) return vfail(SACB_ERR_SAMPLE, "Sample larger than population or is negative (size=%lld, batch=%d)", e->size, B);
    const unsigned long long seed = e->cfg.seed + 0x51ED270B9F5A3C1DULL;
    sacb_rng::draw_indices_kernel<<<1, 256, (size_t)B * sizeof(long long), st>>>(out, B, e->size, seed, e->rng_counter, e->perm_buf);
    CUV(cudaGetLastError());
    e->launches++; e->rng_counter += 1;
    return 0;
}
static int vdraw_normals(sacv_engine* e, float* out, long long n, cudaStream_t st) {
    const unsigned long long seed = e->cfg.seed + 0xA24BAED4963EE407ULL;
    const long long threads = (n + 3) / 4;
    sacb_rng::draw_normal_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(out, n, seed, e->rng_counter << 24);
    CUV(cudaGetLastError());
    e->launches++; e->rng_counter += 1;
    return 0;
}

// gather the batch at idx from the ring into the staging tiles (VisualReplayBuffer.sample, visual_replay_buffer.py:49-66)
static int stage_from_ring(sacv_engine* e, const long long* idx, cudaStream_t st) {
    const sacv_config& c = e->cfg;
    gather_rows_kernel<<<(c.batch_size + 7) / 8, 256, 0, st>>>(idx, c.batch_size, c.obs_dim, c.act_dim, e->rf, e->rnf, e->ra_, e->rr, e->rd, e->xf, e->xnf, e->xa, e->xr, e->xd);
    KCHK();
    if (int r = stage_frames<unsigned char>(e, st, e->rfr, e->rnfr, idx)) return r;
    return finish_staging(e, st);
}

// VisualReplayBuffer.sample into dense user arrays (frames as float32 0..255, like the reference's stacked FloatTensors)
__global__ void sample_visual_kernel(const long long* __restrict__ idx, int B, int O, int A, long long F, const float* __restrict__ rf,
                                     const unsigned char* __restrict__ rfr, const float* __restrict__ ra, const float* __restrict__ rr,
                                     const float* __restrict__ rnf, const unsigned char* __restrict__ rnfr, const unsigned char* __restrict__ rd,
                                     float* f, float* fr, float* a, float* r, float* nf, float* nfr, float* d) {
    const int b = blockIdx.x;
    const long long row = idx[b];
    for (int c = threadIdx.x; c < O; c += blockDim.x) { f[(size_t)b * O + c] = rf[(size_t)row * O + c]; nf[(size_t)b * O + c] = rnf[(size_t)row * O + c]; }
    for (int c = threadIdx.x; c < A; c += blockDim.x) a[(size_t)b * A + c] = ra[(size_t)row * A + c];
    for (long long c = threadIdx.x; c < F; c += blockDim.x) { fr[(size_t)b * F + c] = (float)rfr[(size_t)row * F + c]; nfr[(size_t)b * F + c] = (float)rnfr[(size_t)row * F + c]; }
    if (threadIdx.x == 0) { r[b] = rr[row]; d[b] = rd[row] ? 1.f : 0.f; }
}

int sacv_sample(sacv_handle e, int32_t batch, const int64_t* idx_in, float* feat, float* frames, float* act, float* rew, float* nfeat, float* nframes,
                float* done, int64_t* idx_out, uintptr_t stream) {
    if (!e) return vfail(SACB_ERR_INVALID, "null handle");
    if (int r = vneed(e->rf != nullptr, "replay buffer")) return r;
    if (batch < 1 || e->size < batch) return vfail(SACB_ERR_SAMPLE, "Sample larger than population or is negative (size=%lld, batch=%d)", e->size, batch);
    if (batch > 16384) return vfail(SACB_ERR_INVALID, "batch too large for the on-device sampler");
    cudaStream_t st = (cudaStream_t)stream;
    CUV(cudaSetDevice(e->cfg.device));
    const long long* idx = (const long long*)idx_in;
    long long* tmp = nullptr;
    if (!idx) {
        CUV(cudaMallocAsync(&tmp, (size_t)batch * sizeof(long long), st));
        int* perm = nullptr;
        CUV(cudaMallocAsync(&perm, (size_t)4 * batch * sizeof(int), st));
        const unsigned long long seed = e->cfg.seed + 0x51ED270B9F5A3C1DULL;
        sacb_rng::draw_indices_kernel<<<1, 256, (size_t)batch * sizeof(long long), st>>>(tmp, batch, e->size, seed, e->rng_counter++, perm);
        CUV(cudaGetLastError());
        CUV(cudaFreeAsync(perm, st));
        idx = tmp;
    }
    const long long F = (long long)e->cfg.vis_c * e->cfg.vis_h * e->cfg.vis_w;
    sample_visual_kernel<<<batch, 256, 0, st>>>(idx, batch, e->cfg.obs_dim, e->cfg.act_dim, F, e->rf, e->rfr, e->ra_, e->rr, e->rnf, e->rnfr, e->rd, feat, frames,
                                                act, rew, nfeat, nframes, done);
    CUV(cudaGetLastError());
    if (idx_out) CUV(cudaMemcpyAsync(idx_out, idx, (size_t)batch * sizeof(long long), cudaMemcpyDeviceToDevice, st));
    if (tmp) CUV(cudaFreeAsync(tmp, st));
    e->launches++;
    return 0;
}

// an explicit VisualBatch (device arrays, frames float32) -> staging tiles
int sacv_load_batch(sacv_handle e, const float* feat, const float* frames, const float* act, const float* rew, const float* nfeat, const float* nframes,
                    const float* done, uintptr_t stream) {
    if (!e) return vfail(SACB_ERR_INVALID, "null handle");
    if (!feat || !frames || !act || !rew || !nfeat || !nframes || !done) return vfail(SACB_ERR_INVALID, "null batch array");
    cudaStream_t st = (cudaStream_t)stream;
    CUV(cudaSetDevice(e->cfg.device));
    const sacv_config& c = e->cfg;
    const size_t B = c.batch_size, O = c.obs_dim, A = c.act_dim;
    CUV(cudaMemcpyAsync(e->xf, feat, B * O * 4, cudaMemcpyDeviceToDevice, st)); CUV(cudaMemcpyAsync(e->xnf, nfeat, B * O * 4, cudaMemcpyDeviceToDevice, st));
    CUV(cudaMemcpyAsync(e->xa, act, B * A * 4, cudaMemcpyDeviceToDevice, st)); CUV(cudaMemcpyAsync(e->xr, rew, B * 4, cudaMemcpyDeviceToDevice, st));
    CUV(cudaMemcpyAsync(e->xd, done, B * 4, cudaMemcpyDeviceToDevice, st));
    if (int r = stage_frames<float>(e, st, frames, nframes, nullptr)) return r;
    return finish_staging(e, st);
}

static int vcheck_train(sacv_engine* e, int which) {
    if (int r = vneed(e->pa && e->pc, "actor/critic parameters")) return r;
    if (which & 1) { if (int r = vneed(e->pt != nullptr, "target parameters")) return r; if (int r = vneed(e->mc && e->vc, "critic Adam state")) return r; }
    if (which & 2) if (int r = vneed(e->ma && e->va, "actor Adam state")) return r;
    return 0;
}

// n_updates x (sample -> update_critic -> update_policy -> update_targets), sac/algorithm.py:274-278 on VisualBatches
int sacv_update(sacv_handle e, int32_t n_updates, const int64_t* idx_in, const float* eps_in, float* loss_q, float* loss_pi, uintptr_t stream) {
    if (!e) return vfail(SACB_ERR_INVALID, "null handle");
    if (int r = vcheck_train(e, 3)) return r;
    if (int r = vneed(e->rf != nullptr, "replay buffer")) return r;
    if (n_updates < 0) return vfail(SACB_ERR_INVALID, "negative n_updates");
    cudaStream_t st = (cudaStream_t)stream;
    CUV(cudaSetDevice(e->cfg.device));
    const int B = e->cfg.batch_size, A = e->cfg.act_dim;
    if (e->size < B) return vfail(SACB_ERR_SAMPLE, "Sample larger than population or is negative (size=%lld, batch=%d)", e->size, B);
    for (int k = 0; k < n_updates; ++k) {
        const long long* idx = idx_in ? (const long long*)idx_in + (size_t)k * B : nullptr;
        if (!idx) { if (int r = vdraw_indices(e, e->idx_buf, B, st)) return r; idx = e->idx_buf; }
        const float* eps = eps_in ? eps_in + (size_t)k * 2 * B * A : nullptr;
        if (!eps) { if (int r = vdraw_normals(e, e->eps_buf, 2LL * B * A, st)) return r; eps = e->eps_buf; }
        if (int r = stage_from_ring(e, idx, st)) return r;
        if (int r = actor_eval(e, st, eps, eps + (size_t)B * A, 3)) return r;
        if (int r = critic_step(e, st, loss_q ? loss_q + k : nullptr, true)) return r;
        e->step_q += 1;
        if (int r = policy_step(e, st, eps + (size_t)B * A, loss_pi ? loss_pi + k : nullptr)) return r;
        e->step_pi += 1;
    }
    return 0;
}

int sacv_update_critic(sacv_handle e, const float* eps_in, float* loss_q, uintptr_t stream) {
    if (!e) return vfail(SACB_ERR_INVALID, "null handle");
    if (int r = vcheck_train(e, 1)) return r;
    if (!e->batch_loaded) return vfail(SACB_ERR_STATE, "no batch loaded (sacv_load_batch)");
    cudaStream_t st = (cudaStream_t)stream;
    CUV(cudaSetDevice(e->cfg.device));
    const float* eps = eps_in;
    if (!eps) { if (int r = vdraw_normals(e, e->eps_buf, (long long)e->cfg.batch_size * e->cfg.act_dim, st)) return r; eps = e->eps_buf; }
    if (int r = actor_eval(e, st, eps, nullptr, 1)) return r;
    if (int r = critic_step(e, st, loss_q, false)) return r;
    e->step_q += 1;
    return 0;
}
int sacv_update_policy(sacv_handle e, const float* eps_in, float* loss_pi, uintptr_t stream) {
    if (!e) return vfail(SACB_ERR_INVALID, "null handle");
    if (int r = vcheck_train(e, 2)) return r;
    if (!e->batch_loaded) return vfail(SACB_ERR_STATE, "no batch loaded (sacv_load_batch)");
    cudaStream_t st = (cudaStream_t)stream;
    CUV(cudaSetDevice(e->cfg.device));
    const float* eps = eps_in;
    if (!eps) { if (int r = vdraw_normals(e, e->eps_buf, (long long)e->cfg.batch_size * e->cfg.act_dim, st)) return r; eps = e->eps_buf; }
    if (int r = actor_eval(e, st, nullptr, eps, 2)) return r;
    if (int r = policy_step(e, st, eps, loss_pi)) return r;
    e->step_pi += 1;
    return 0;
}
int sacv_update_targets(sacv_handle e, uintptr_t stream) {
    if (!e) return vfail(SACB_ERR_INVALID, "null handle");
    if (int r = vneed(e->pc && e->pt, "critic/target parameters")) return r;
    cudaStream_t st = (cudaStream_t)stream;
    CUV(cudaSetDevice(e->cfg.device));
    polyak_kernel<<<nblk(e->lay.critic_floats), 256, 0, st>>>(e->pt, e->pc, e->lay.critic_floats, (float)e->cfg.polyak, (float)(1.0 - e->cfg.polyak));
    KCHK();
    return 0;
}

// VisualActor.forward (convolutional.py:85-121) on n <= batch_size rows (device arrays, frames float32)
int sacv_actor_forward(sacv_handle e, const float* feat, const float* frames, int32_t n, const float* eps_in, int32_t deterministic, float* pi, float* logp,
                       uintptr_t stream) {
    if (!e) return vfail(SACB_ERR_INVALID, "null handle");
    if (int r = vneed(e->pa != nullptr, "actor parameters")) return r;
    const sacv_config& c = e->cfg; const sacv_layout_info& L = e->lay;
    const int B = c.batch_size, O = c.obs_dim, A = c.act_dim;
    if (n < 1 || n > B || !feat || !frames || !pi) return vfail(SACB_ERR_INVALID, "bad arguments (1 <= n <= batch_size)");
    cudaStream_t st = (cudaStream_t)stream;
    CUV(cudaSetDevice(c.device));
    const size_t F = (size_t)c.vis_c * c.vis_h * c.vis_w;
    // rows beyond n: whatever the staging tile holds (finite); they are computed and discarded
    CUV(cudaMemcpyAsync(e->xnf, feat, (size_t)n * O * 4, cudaMemcpyDeviceToDevice, st));
    float* tmp = nullptr;
    const float* fr = frames;
    if (n < B) {
        CUV(cudaMallocAsync(&tmp, (size_t)B * F * sizeof(float), st));
        CUV(cudaMemsetAsync(tmp, 0, (size_t)B * F * sizeof(float), st));
        CUV(cudaMemcpyAsync(tmp, frames, (size_t)n * F * 4, cudaMemcpyDeviceToDevice, st));
        fr = tmp;
    }
    const long long nn = (long long)B * L.conv_h[0] * L.conv_w[0] * c.vis_c * c.kernels[0];
    im2col_frames_kernel<float><<<nblk(nn), 256, 0, st>>>(fr, nullptr, e->col1_n, B, c.vis_c, c.vis_h, c.vis_w, c.kernels[0], c.strides[0], L.conv_h[0], L.conv_w[0]);
    KCHK();
    const float* eps = nullptr;
    if (!deterministic) {
        if (eps_in) { CUV(cudaMemcpyAsync(e->eps_buf, eps_in, (size_t)n * A * 4, cudaMemcpyDeviceToDevice, st)); eps = e->eps_buf; }
        else { if (int r = vdraw_normals(e, e->eps_buf, (long long)B * A, st)) return r; eps = e->eps_buf; }
    }
    if (int r = actor_eval(e, st, eps, nullptr, 1)) return r;
    CUV(cudaMemcpy2DAsync(pi, A * sizeof(float), e->x2 + O, e->ldx * sizeof(float), A * sizeof(float), n, cudaMemcpyDeviceToDevice, st));
    if (logp) CUV(cudaMemcpyAsync(logp, e->logp, (size_t)n * 4, cudaMemcpyDeviceToDevice, st));
    if (tmp) CUV(cudaFreeAsync(tmp, st));
    e->batch_loaded = false;
    return 0;
}

// VisualDoubleCritic.forward (convolutional.py:144-183) on n <= batch_size rows; which: 0 critic, 1 target
int sacv_critic_forward(sacv_handle e, int32_t which, const float* feat, const float* frames, const float* act, int32_t n, float* q1, float* q2,
                        uintptr_t stream) {
    if (!e) return vfail(SACB_ERR_INVALID, "null handle");
    if (which != 0 && which != 1) return vfail(SACB_ERR_INVALID, "which must be 0 (critic) or 1 (target)");
    const float* P = which ? e->pt : e->pc;
    if (int r = vneed(P != nullptr, which ? "target parameters" : "critic parameters")) return r;
    const sacv_config& c = e->cfg; const sacv_layout_info& L = e->lay;
    const int B = c.batch_size, O = c.obs_dim, A = c.act_dim;
    if (n < 1 || n > B || !feat || !frames || !act || !q1 || !q2) return vfail(SACB_ERR_INVALID, "bad arguments (1 <= n <= batch_size)");
    cudaStream_t st = (cudaStream_t)stream;
    CUV(cudaSetDevice(c.device));
    const size_t F = (size_t)c.vis_c * c.vis_h * c.vis_w;
    CUV(cudaMemcpyAsync(e->xf, feat, (size_t)n * O * 4, cudaMemcpyDeviceToDevice, st));
    CUV(cudaMemcpyAsync(e->xa, act, (size_t)n * A * 4, cudaMemcpyDeviceToDevice, st));
    build_x_kernel<<<(B * (O + A) + 255) / 256, 256, 0, st>>>(e->xf, O, e->xa, A, A, e->xs, e->ldx, B);
    KCHK();
    float* tmp = nullptr;
    const float* fr = frames;
    if (n < B) {
        CUV(cudaMallocAsync(&tmp, (size_t)B * F * sizeof(float), st));
        CUV(cudaMemsetAsync(tmp, 0, (size_t)B * F * sizeof(float), st));
        CUV(cudaMemcpyAsync(tmp, frames, (size_t)n * F * 4, cudaMemcpyDeviceToDevice, st));
        fr = tmp;
    }
    const long long nn = (long long)B * L.conv_h[0] * L.conv_w[0] * c.vis_c * c.kernels[0];
    im2col_frames_kernel<float><<<nblk(nn), 256, 0, st>>>(fr, nullptr, e->col1_s, B, c.vis_c, c.vis_h, c.vis_w, c.kernels[0], c.strides[0], L.conv_h[0], L.conv_w[0]);
    KCHK();
    if (int r = cnn_forward(e, st, P + L.c_cnn, L.critic_net_floats, e->col1_s, e->cs, e->cs.c, 1, B)) return r;
    if (int r = critic_mlp_forward(e, st, P, e->xs, e->cs.c, e->ch1, e->ch2, e->cm, e->qpi)) return r;
    CUV(cudaMemcpyAsync(q1, e->qpi, (size_t)n * 4, cudaMemcpyDeviceToDevice, st));
    CUV(cudaMemcpyAsync(q2, e->qpi + B, (size_t)n * 4, cudaMemcpyDeviceToDevice, st));
    if (tmp) CUV(cudaFreeAsync(tmp, st));
    e->batch_loaded = false;
    return 0;
}

int sacv_debug_read(sacv_handle e, const char* name, float* dst, int64_t max_floats, int64_t* n_out) {
    if (!e || !name || !dst) return vfail(SACB_ERR_INVALID, "null argument");
    const long long B = e->cfg.batch_size, A = e->cfg.act_dim, O = e->cfg.obs_dim;
    std::string n(name);
    const float* src = nullptr;
    long long cnt = 0, rows = 0, cols = 0, ld = 0;
    if (n == "q1") { src = e->qv; cnt = B; } else if (n == "q2") { src = e->qv + B; cnt = B; }
    else if (n == "q1_target") { src = e->qt; cnt = B; } else if (n == "q2_target") { src = e->qt + B; cnt = B; }
    else if (n == "backup") { src = e->backup; cnt = B; }
    else if (n == "q1_pi") { src = e->qpi; cnt = B; } else if (n == "q2_pi") { src = e->qpi + B; cnt = B; }
    else if (n == "logp_next") { src = e->logp; cnt = B; } else if (n == "logp_pi") { src = e->logp + B; cnt = B; }
    else if (n == "a2") { src = e->x2 + O; rows = B; cols = A; ld = e->ldx; } else if (n == "pi") { src = e->xp + O; rows = B; cols = A; ld = e->ldx; }
    else if (n == "grad_actor") { src = e->ga; cnt = e->lay.actor_floats; } else if (n == "grad_critic") { src = e->gc; cnt = e->lay.critic_floats; }
    else if (n == "loss") { src = e->loss; cnt = 2; } else if (n == "rewards") { src = e->xr; cnt = B; } else if (n == "done") { src = e->xd; cnt = B; }
    else if (n == "features") { src = e->xf; cnt = B * O; }
    else return vfail(SACB_ERR_INVALID, "unknown debug array '%s'", name);
    CUV(cudaSetDevice(e->cfg.device));
    CUV(cudaDeviceSynchronize());
    if (rows) {
        if (rows * cols > max_floats) return vfail(SACB_ERR_INVALID, "destination too small");
        CUV(cudaMemcpy2D(dst, cols * sizeof(float), src, ld * sizeof(float), cols * sizeof(float), rows, cudaMemcpyDeviceToHost));
        if (n_out) *n_out = rows * cols;
    } else {
        if (cnt > max_floats) return vfail(SACB_ERR_INVALID, "destination too small");
        CUV(cudaMemcpy(dst, src, cnt * sizeof(float), cudaMemcpyDeviceToHost));
        if (n_out) *n_out = cnt;
    }
    return 0;
}

int64_t sacv_launch_count(sacv_handle e) { return e ? e->launches : 0; }

}  // extern "C"
