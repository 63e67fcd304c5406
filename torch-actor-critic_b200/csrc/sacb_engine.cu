// sacb_engine.cu -- host side of libsacb200.so: engine state, phase-program builder, launches and the C ABI
// declared in include/sacb200.h.  Compile: nvcc -gencode arch=compute_100a,code=sm_100a (see __graft_entry__.build).
#include "sacb_device.cuh"

#include <cuda.h>                  // CUtensorMap types only: cuTensorMapEncodeTiled is fetched through cudaGetDriverEntryPoint
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/sacb200.h"

using namespace sacb;
#include "sacb_act.cuh"
#include "sacb_rng.cuh"
using sacb_rng::draw_indices_kernel;
using sacb_rng::draw_normal_kernel;

// ------------------------------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static int fail(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}
#define CU(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess) return fail(SACB_ERR_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

// ------------------------------------------------------------------------------------------------------------
// minimal NCCL binding (dlopen; declarations restated from the public nccl.h API)
// ------------------------------------------------------------------------------------------------------------
typedef struct { char internal[128]; } nccl_uid_t;
typedef void* nccl_comm_t;
struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(nccl_uid_t*) = nullptr;
    int (*CommInitRank)(nccl_comm_t*, int, nccl_uid_t, int) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*Broadcast)(const void*, void*, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*CommDestroy)(nccl_comm_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;
static int load_nccl(const char* path) {
    if (g_nccl.lib) return 0;
    const char* cands[] = {path, "libnccl.so.2", "libnccl.so"};
    void* lib = nullptr;
    for (const char* c : cands) {
        if (!c || !*c) continue;
        lib = dlopen(c, RTLD_NOW | RTLD_GLOBAL);
        if (lib) break;
    }
    if (!lib) return fail(SACB_ERR_NCCL, "dlopen libnccl failed: %s", dlerror());
    g_nccl.lib = lib;
#define SYM(field, name)                                                               \
    *(void**)(&g_nccl.field) = dlsym(lib, name);                                       \
    if (!g_nccl.field) return fail(SACB_ERR_NCCL, "libnccl misses symbol %s", name);
    SYM(GetUniqueId, "ncclGetUniqueId");
    SYM(CommInitRank, "ncclCommInitRank");
    SYM(AllReduce, "ncclAllReduce");
    SYM(Broadcast, "ncclBroadcast");
    SYM(CommDestroy, "ncclCommDestroy");
    SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
    return 0;
}
#define NC(call)                                                                                          \
    do {                                                                                                  \
        int r_ = (call);                                                                                  \
        if (r_ != 0) return fail(SACB_ERR_NCCL, "%s: %s", #call, g_nccl.GetErrorString ? g_nccl.GetErrorString(r_) : "?"); \
    } while (0)
constexpr int kNcclFloat = 7, kNcclSum = 0;

// ------------------------------------------------------------------------------------------------------------
// engine
// ------------------------------------------------------------------------------------------------------------
static inline long long r4(long long x) { return (x + 3) & ~3LL; }

struct ProgRange { int begin = 0, end = 0; };

struct sacb_engine {
    sacb_config cfg{};
    sacb_layout_info lay{};
    int num_sms = 0, grid_coop = 0, grid_std = 0, grid_t5 = 0;
    int tile_m = kBM, tile_n = kBN;       // GEMM tile of the active arithmetic path (64x16 for tcgen05)
    int mode = 0;
    int tf32x3 = 1;                 // GEMM arithmetic: 0 fp32 FFMA, 1 3xTF32 mma.sync (default), 2 3xTF32 tcgen05, 3 bf16 tcgen05
    long long launches = 0;
    // bound arenas
    float *pa = nullptr, *pc = nullptr, *pt = nullptr;
    float *ma = nullptr, *va = nullptr, *mc = nullptr, *vc = nullptr;
    long long step_pi = 0, step_q = 0;
    long long exch_pi = 0, exch_q = 0;          // gradient exchanges done so far (peer-flag epochs): never reset by a re-bind
    float *bs = nullptr, *ba = nullptr, *br = nullptr, *bs2 = nullptr;
    uint8_t* bd = nullptr;
    float* brec = nullptr; int rec_pitch = 0;   // packed ring (sacb_bind_buffer_packed): [s | a | r | s' | d | pad] per transition
    long long cap = 0, ptr = 0, size = 0;
    // workspace
    float* ws = nullptr;
    size_t ws_floats = 0;
    int ldx = 0, lds = 0, ldh = 0, lddpi = 0, tnH2 = 0, row_tiles = 0;
    float *xs, *xp, *x2, *s2buf, *rd, *ah1, *ah2, *heads, *logp, *ch1, *ch2, *th1, *th2, *qpart, *qtpart, *qbuf, *qfw,
        *dh2, *dh1, *dpi, *ghead, *hpart, *dpart, *adh2, *adh1, *lparts_q, *lparts_pi, *loss_scalar, *ga, *gc;
    // folded program only: critic activations of the actor-loss pass, its Q-dot partials, per-row-block column-sum partials
    float *cp1, *cp2, *qppart, *colp_b0, *colp_b1, *colp_w2, *colp_ab0, *colp_ab1, *colp_bh;
    int row_blocks = 0;
    bool fold_ok = false;           // the folded program exists for this configuration / arithmetic path
    int last_parity = 0;            // staging parity of the last update executed (debug reads of the folded program)
    long long* idx_buf = nullptr;
    int* perm_buf = nullptr;
    size_t perm_cap = 0;
    float* eps_buf = nullptr;
    unsigned long long* barrier = nullptr;
    unsigned long long rng_counter = 0;
    // programs
    bool dirty = true;
    std::vector<Job> jobs;
    std::vector<Phase> phases;
    Job* d_jobs = nullptr;
    Phase* d_phases = nullptr;
    size_t d_jobs_cap = 0, d_phases_cap = 0;
    ProgRange p_update, p_policy, p_targets, p_actor_fwd, p_critic_fwd[2], p_fold;
    // bf16 shadows of the GEMM operands + their tensor maps (precision 3, csrc/sacb_tma.cuh)
    struct ShReg { const float* base; long long rows; int ld, cols; long long off16; int ld16; };
    std::vector<ShReg> shreg;                   // fp32 region -> shadow
    std::vector<CUtensorMap> tmaps;
    unsigned short* sh16 = nullptr; size_t sh16_elems = 0;
    void* d_tmaps = nullptr; size_t d_tmaps_cap = 0;
    Sh16 sh{};
    bool bf_tma = false;                        // precision 3 and every GEMM operand of every program is shadowed
    int n_refresh = 0;                          // regions the pre-launch refresh kernel converts (weights + staging tiles)
    void* d_refresh = nullptr;
    int ph_crit_begin = 0, ph_crit_adam = 0, ph_pi_begin = 0, ph_pi_adam = 0;   // split points inside p_update
    // store staging
    uint8_t* stage = nullptr;
    size_t stage_bytes = 0;
    cudaEvent_t stage_ev = nullptr;
    // comm
    nccl_comm_t comm = nullptr;
    int rank = 0, world = 1;
    // peer mode: in-kernel NVLink gradient exchange (workspaces mapped across ranks with CUDA IPC)
    int peer_mode = 0;
    float* peer_ws[kMaxPeers] = {nullptr};
    float* peer_flags = nullptr;
    unsigned int* abort_host = nullptr;        // host-mapped abort word of the peer exchange (see RunArgs.abort_word)
    unsigned int* abort_dev = nullptr;
    unsigned long long peer_timeout_ns = 20000000000ULL;
};

static void compute_layout(const sacb_config& c, sacb_layout_info& L) {
    const long long O = c.obs_dim, A = c.act_dim, H1 = c.hidden1, H2 = c.hidden2;
    long long o = 0;
    L.actor_w0 = o; o = r4(o + H1 * O);
    L.actor_b0 = o; o = r4(o + H1);
    L.actor_w1 = o; o = r4(o + H2 * H1);
    L.actor_b1 = o; o = r4(o + H2);
    L.actor_w_mu = o; L.actor_w_ls = o + A * H2; o = r4(o + 2 * A * H2);     // [mu ; log_std] form one [2A,H2] matrix
    L.actor_b_mu = o; L.actor_b_ls = o + A; o = r4(o + 2 * A);
    L.actor_floats = o;
    o = 0;
    const long long W0 = H1 * (O + A);
    L.critic_w0[0] = o; L.critic_w0[1] = o + W0; o = r4(o + 2 * W0);         // twin layer-0 weights form one [2H1,O+A] matrix
    L.critic_b0[0] = o; L.critic_b0[1] = o + H1; o = r4(o + 2 * H1);
    L.critic_w1[0] = o; L.critic_w1[1] = o + H2 * H1; o = r4(o + 2 * H2 * H1);
    L.critic_b1[0] = o; L.critic_b1[1] = o + H2; o = r4(o + 2 * H2);
    L.critic_w2[0] = o; L.critic_w2[1] = o + H2; o = r4(o + 2 * H2);
    L.critic_b2[0] = o; L.critic_b2[1] = o + 1; o = r4(o + 2);
    L.critic_floats = o;
}

static int check_cfg(const sacb_config* c) {
    if (!c) return fail(SACB_ERR_INVALID, "null config");
    if (c->struct_size != sizeof(sacb_config)) return fail(SACB_ERR_INVALID, "sacb_config size mismatch (ABI): got %u want %zu", c->struct_size, sizeof(sacb_config));
    if (c->obs_dim < 1 || c->act_dim < 1 || c->hidden1 < 1 || c->hidden2 < 1) return fail(SACB_ERR_INVALID, "dims must be positive");
    if (c->batch_size < 1 || c->batch_size > 16384) return fail(SACB_ERR_INVALID, "batch_size out of range (1..16384, the on-device sampler's limit)");
    if (c->max_burst < 1 || c->max_burst > kMaxBurst) return fail(SACB_ERR_INVALID, "max_burst must be in [1, %d]", kMaxBurst);
    return 0;
}

// ------------------------------------------------------------------------------------------------------------
// program builder
// ------------------------------------------------------------------------------------------------------------
struct Builder {
    sacb_engine* e;
    std::vector<Job>& jobs;
    std::vector<Phase>& phases;
    int cur = -1;
    explicit Builder(sacb_engine* e_) : e(e_), jobs(e_->jobs), phases(e_->phases) {}
    int phase() {
        Phase p{(int)jobs.size(), 0, 0, 0};
        phases.push_back(p);
        cur = (int)phases.size() - 1;
        return cur;
    }
    void add(Job j) {
        Phase& p = phases[cur];
        j.tile_begin = p.ntiles;
        p.ntiles += j.ntiles;
        p.job_count += 1;
        jobs.push_back(j);
    }
    static bool al16(const void* p) { return ((uintptr_t)p & 15) == 0; }
    static Job zero() { Job j; memset(&j, 0, sizeof j); j.nz = 1; return j; }
    // C[M,N] = epi( sum_r A(i,r) B(j,r) )
    Job gemm(int mode, int epi, const float* A, int lda, long long sA, const float* B, int ldb, long long sB, float* C,
             int ldc, long long sC, int M, int N, int K, int nz, int dyn = 0) {
        Job j = zero();
        j.type = J_GEMM; j.mode = mode; j.epi = epi;
        j.A = A; j.lda = lda; j.sA = sA; j.B = B; j.ldb = ldb; j.sB = sB; j.C = C; j.ldc = ldc; j.sC = sC;
        j.M = M; j.N = N; j.K = K; j.nz = nz; j.rows_dynamic = dyn;
        j.a_vec = al16(A) && (lda % 4 == 0) && (sA % 4 == 0);
        j.b_vec = al16(B) && (ldb % 4 == 0) && (sB % 4 == 0);
        j.tn = (N + e->tile_n - 1) / e->tile_n;
        j.tpz = ((M + e->tile_m - 1) / e->tile_m) * j.tn;
        j.ntiles = j.tpz * nz;
        if (nz > 2) { fprintf(stderr, "sacb: nz > 2 unsupported\n"); abort(); }
        return j;
    }
    Job colsum(const float* X, int ldx, long long sX, const float* w, long long sW, float* out, long long sOut, int M, int N, int nz,
               int width = 32) {
        Job j = zero();
        j.type = J_COLSUM; j.A = X; j.lda = ldx; j.sA = sX; j.aux0 = w; j.sAux = sW; j.C = out; j.sC = sOut;
        j.M = M; j.N = N; j.nz = nz; j.i0 = width; j.ntiles = ((N + width - 1) / width) * nz;
        return j;
    }
};

#include "sacb_program.inl"
#include "sacb_tma_host.inl"

static int upload_programs(sacb_engine* e, cudaStream_t st) {
    if (!e->dirty) return 0;
    build_programs(e);
    if (int r = build_tma(e, st)) return r;
    if (e->jobs.size() > e->d_jobs_cap) {
        if (e->d_jobs) CU(cudaFree(e->d_jobs));
        e->d_jobs_cap = e->jobs.size() + 16;
        CU(cudaMalloc(&e->d_jobs, e->d_jobs_cap * sizeof(Job)));
    }
    if (e->phases.size() > e->d_phases_cap) {
        if (e->d_phases) CU(cudaFree(e->d_phases));
        e->d_phases_cap = e->phases.size() + 16;
        CU(cudaMalloc(&e->d_phases, e->d_phases_cap * sizeof(Phase)));
    }
    // synchronous copies: the tables are rebuilt only on (re)bind
    CU(cudaStreamSynchronize(st));
    CU(cudaMemcpy(e->d_jobs, e->jobs.data(), e->jobs.size() * sizeof(Job), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(e->d_phases, e->phases.data(), e->phases.size() * sizeof(Phase), cudaMemcpyHostToDevice));
    e->dirty = false;
    return 0;
}

// ------------------------------------------------------------------------------------------------------------
// small standalone kernels: RNG draws, dense batch load, sample gather
// ------------------------------------------------------------------------------------------------------------
// Dense batch (the reference's Batch dataclass, buffer/replay_buffer.py:8-14) -> staging tiles.
__global__ void load_batch_kernel(const float* __restrict__ s, const float* __restrict__ a, const float* __restrict__ r,
                                  const float* __restrict__ s2, const float* __restrict__ d, int B, int O, int A, float* xs,
                                  float* xp, float* x2, float* s2buf, float* rd, int ldx, int lds) {
    const int m = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (m >= B) return;
    for (int c = lane; c < O; c += 32) {
        float v = s[(size_t)m * O + c], w = s2[(size_t)m * O + c];
        xs[(size_t)m * ldx + c] = v; xp[(size_t)m * ldx + c] = v; x2[(size_t)m * ldx + c] = w; s2buf[(size_t)m * lds + c] = w;
    }
    for (int c = lane; c < A; c += 32) xs[(size_t)m * ldx + O + c] = a[(size_t)m * A + c];
    if (lane == 0) { rd[m] = r[m]; rd[B + m] = d[m]; }
}

// ReplayBuffer.sample gather into dense user arrays (buffer/replay_buffer.py:48-54).  Warp per row.
__global__ void sample_gather_kernel(const long long* __restrict__ idx, int B, int O, int A, const float* __restrict__ bs,
                                     const float* __restrict__ ba, const float* __restrict__ br, const float* __restrict__ bs2,
                                     const unsigned char* __restrict__ bd, float* s, float* a, float* r, float* s2, float* d,
                                     long long* idx_out) {
    const int m = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (m >= B) return;
    const long long row = idx[m];
    for (int c = lane; c < O; c += 32) { s[(size_t)m * O + c] = bs[(size_t)row * O + c]; s2[(size_t)m * O + c] = bs2[(size_t)row * O + c]; }
    for (int c = lane; c < A; c += 32) a[(size_t)m * A + c] = ba[(size_t)row * A + c];
    if (lane == 0) { r[m] = br[row]; d[m] = bd[row] ? 1.f : 0.f; if (idx_out) idx_out[m] = row; }
}

// same from the packed ring: one record per row, 128-bit loads
__global__ void sample_gather_packed_kernel(const long long* __restrict__ idx, int B, int O, int A, const float* __restrict__ rec, int pitch,
                                            float* s, float* a, float* r, float* s2, float* d, long long* idx_out) {
    const int m = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (m >= B) return;
    const long long row = idx[m];
    const float* src = rec + (size_t)row * pitch;
    const int W = 2 * O + A + 2;
    for (int c = lane; c < W; c += 32) {
        const float v = src[c];
        if (c < O) s[(size_t)m * O + c] = v;
        else if (c < O + A) a[(size_t)m * A + c - O] = v;
        else if (c == O + A) r[m] = v;
        else if (c < 2 * O + A + 1) s2[(size_t)m * O + c - (O + A + 1)] = v;
        else d[m] = v != 0.f ? 1.f : 0.f;
    }
    if (lane == 0 && idx_out) idx_out[m] = row;
}

// ------------------------------------------------------------------------------------------------------------
// launch helpers
// ------------------------------------------------------------------------------------------------------------
static void fill_args(sacb_engine* e, RunArgs& ra) {
    memset(&ra, 0, sizeof ra);
    const sacb_config& c = e->cfg;
    ra.jobs = e->d_jobs; ra.phases = e->d_phases;
    ra.barrier = e->barrier;
    ra.buf_state = e->bs; ra.buf_actions = e->ba; ra.buf_rewards = e->br; ra.buf_next_state = e->bs2; ra.buf_done = e->bd;
    ra.buf_rec = e->brec; ra.rec_pitch = e->rec_pitch;
    ra.alpha = (float)c.alpha; ra.gamma = (float)c.gamma; ra.polyak = (float)c.polyak; ra.reward_scale = (float)c.reward_scale;
    ra.polyak_c = (float)(1.0 - c.polyak);          // `(1 - polyak) * src` with a python-float polyak (sac/algorithm.py:81)
    ra.log_min_std = c.log_min_std; ra.log_max_std = c.log_max_std; ra.act_limit = c.act_limit;
    ra.adam_w1 = (float)(1.0 - c.beta1); ra.adam_b2 = (float)c.beta2; ra.adam_w2 = (float)(1.0 - c.beta2);
    ra.adam_eps = (float)c.adam_eps;
    // torch._single_tensor_adam forms these in python doubles per step: bias_correction{1,2} = 1 - beta^t,
    // step_size = lr / bc1, bc2_sqrt = bc2 ** 0.5; the kernels consume them as float32 scalars like torch does.
    for (int k = 0; k < kMaxBurst; ++k) {
        const double ta = (double)(e->step_pi + k + 1), tq = (double)(e->step_q + k + 1);
        ra.adam_coef[k][0] = (float)(-(c.lr_pi / (1.0 - std::pow(c.beta1, ta))));
        ra.adam_coef[k][1] = (float)std::sqrt(1.0 - std::pow(c.beta2, ta));
        ra.adam_coef[k][2] = (float)(-(c.lr_q / (1.0 - std::pow(c.beta1, tq))));
        ra.adam_coef[k][3] = (float)std::sqrt(1.0 - std::pow(c.beta2, tq));
    }
    ra.tf32x3 = e->tf32x3;
    ra.peer_world = e->peer_mode ? e->world : 1;
    ra.peer_rank = e->rank;
    for (int r = 0; r < kMaxPeers; ++r) ra.peer_ws[r] = e->peer_ws[r];
    ra.peer_off_flags = e->peer_flags - e->ws;
    ra.epoch_q = e->exch_q; ra.epoch_pi = e->exch_pi;
    ra.abort_word = e->peer_mode ? e->abort_dev : nullptr;
    ra.peer_timeout_ns = e->peer_timeout_ns;
    ra.grad_div = (float)e->world;
    ra.flags = c.flags;
    ra.O = c.obs_dim; ra.A = c.act_dim; ra.H1 = c.hidden1; ra.H2 = c.hidden2; ra.B = c.batch_size;
    ra.s2_stride = (long long)c.batch_size * e->lds;
    ra.pass_mask = 3;
    if (e->bf_tma && e->tf32x3 == 3) ra.sh = e->sh;
    {   // the folded program's generators (sacb_fold.cuh)
        const sacb_layout_info& L = e->lay;
        FoldPtrs& f = ra.f;
        const long long B = c.batch_size;
        if (e->pa) { f.a_w0 = e->pa + L.actor_w0; f.a_b0 = e->pa + L.actor_b0; f.a_wh = e->pa + L.actor_w_mu; f.a_bh = e->pa + L.actor_b_mu; }
        if (e->pc) { f.c_w0 = e->pc + L.critic_w0[0]; f.c_b0 = e->pc + L.critic_b0[0]; f.c_w2 = e->pc + L.critic_w2[0]; f.c_b2 = e->pc + L.critic_b2[0]; }
        if (e->pt) { f.t_w0 = e->pt + L.critic_w0[0]; f.t_b0 = e->pt + L.critic_b0[0]; f.t_b2 = e->pt + L.critic_b2[0]; }
        f.xs = e->xs; f.s2 = e->s2buf; f.rd = e->rd;
        f.xs_stride = B * e->ldx; f.s2_stride = B * e->lds; f.rd_stride = 2 * B;
        f.ldx = e->ldx; f.lds = e->lds; f.ldh = e->ldh; f.tnH2 = e->tnH2; f.tnH1x2 = 2 * ((c.hidden1 + e->tile_n - 1) / e->tile_n);
        f.ah1 = e->ah1; f.ch1 = e->ch1; f.cp1 = e->cp1;
        f.hpart = e->hpart; f.heads = e->heads; f.logp = e->logp;
        f.dbg_a2 = e->x2 + c.obs_dim; f.dbg_pi = e->xp + c.obs_dim;
        f.qpart = e->qpart; f.qtpart = e->qtpart; f.qppart = e->qppart;
        f.qbuf = e->qbuf; f.lparts_q = e->lparts_q; f.lparts_pi = e->lparts_pi;
        f.dh2 = e->dh2; f.colp_b1 = e->colp_b1; f.colp_w2 = e->colp_w2;
        f.dpart = e->dpart; f.ghead = e->ghead; f.adh2 = e->adh2; f.colp_ab1 = e->colp_ab1; f.colp_bh = e->colp_bh;
    }
}

// Run phases [pb, pe) for updates [kb, ke).  Cooperative persistent launch unless mode==1 or a single phase.
static int run_phases(sacb_engine* e, RunArgs ra, int pb, int pe, int kb, int ke, cudaStream_t st) {
    if (pe <= pb || ke <= kb) return 0;
    {
        const int nj = e->phases[pe - 1].job_begin + e->phases[pe - 1].job_count - e->phases[pb].job_begin;
        if (nj > kMaxJobs || pe - pb > kMaxPhases) return fail(SACB_ERR_INVALID, "program slice too large (%d jobs, %d phases)", nj, pe - pb);
    }
    if (ra.sh.base) refresh_shadows(e, st);          // weights / staging tiles may have been written by the host side since
    const bool single = (pe - pb == 1) && (ke - kb == 1);
    if (e->mode == 1 || single) {
        for (int k = kb; k < ke; ++k)
            for (int p = pb; p < pe; ++p) {
                ra.phase_begin = p; ra.phase_end = p + 1; ra.k_begin = k; ra.k_end = k + 1; ra.use_barrier = 0;
                int nt = e->phases[p].ntiles;
                int grid = nt < 1 ? 1 : nt;
                if (e->tf32x3 >= 2) sac_phase_kernel<true><<<grid, kThreads, kSmemBytesT5, st>>>(ra);
                else sac_phase_kernel<false><<<grid, kThreads, kSmemBytes, st>>>(ra);
                CU(cudaGetLastError());
                e->launches++;
            }
        return 0;
    }
    ra.phase_begin = pb; ra.phase_end = pe; ra.k_begin = kb; ra.k_end = ke; ra.use_barrier = 1;
    CU(cudaMemsetAsync(e->barrier, 0, sizeof(unsigned long long), st));
    void* args[] = {&ra};
    if (e->tf32x3 >= 2)
        CU(cudaLaunchCooperativeKernel((void*)sac_phase_kernel<true>, dim3(e->grid_coop), dim3(kThreads), args, kSmemBytesT5, st));
    else
        CU(cudaLaunchCooperativeKernel((void*)sac_phase_kernel<false>, dim3(e->grid_coop), dim3(kThreads), args, kSmemBytes, st));
    e->launches++;
    return 0;
}

static int need(bool cond, const char* what) { return cond ? 0 : fail(SACB_ERR_STATE, "%s not bound", what); }

// ------------------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------------------
extern "C" {

int sacb_abi_version(void) { return SACB_ABI_VERSION; }
const char* sacb_last_error(void) { return g_err.c_str(); }

int sacb_layout(const sacb_config* cfg, sacb_layout_info* out) {
    if (int r = check_cfg(cfg)) return r;
    if (!out) return fail(SACB_ERR_INVALID, "null out");
    compute_layout(*cfg, *out);
    return 0;
}

int sacb_create(const sacb_config* cfg, sacb_handle* out) {
    if (int r = check_cfg(cfg)) return r;
    if (!out) return fail(SACB_ERR_INVALID, "null out");
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0)
        return fail(SACB_ERR_CUDA, "no CUDA device available (%s); this engine has no CPU fallback", cudaGetErrorString(ce));
    if (cfg->device < 0 || cfg->device >= ndev) return fail(SACB_ERR_INVALID, "device %d out of range", cfg->device);
    CU(cudaSetDevice(cfg->device));
    sacb_engine* e = new sacb_engine();
    e->cfg = *cfg;
    e->rank = cfg->rank; e->world = 1;
    compute_layout(*cfg, e->lay);
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, cfg->device));
    e->num_sms = prop.multiProcessorCount;
    CU(cudaFuncSetAttribute(sac_phase_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    CU(cudaFuncSetAttribute(sac_phase_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytesT5));
    CU(cudaFuncSetAttribute(draw_indices_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * (int)sizeof(long long)));
    CU(cudaFuncSetAttribute(actor_act_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));   // per device: set at every create
    int occ = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, sac_phase_kernel<false>, kThreads, kSmemBytes));
    if (occ < 1) { delete e; return fail(SACB_ERR_CUDA, "sac_phase_kernel does not fit on an SM"); }
    if (occ > 2) occ = 2;
    occ = 1;    // one CTA per SM measured faster than two (8742 vs 8267 updates/s at C1): half the barrier participants
    if (const char* o = getenv("SACB_OCC")) { if (atoi(o) == 2) occ = 2; }      // experiment knob
    e->grid_std = occ * e->num_sms;
    if (const char* gsz = getenv("SACB_GRID")) { const int v = atoi(gsz); if (v >= 1 && v <= e->grid_std) e->grid_std = v; }   // experiment knob
    int occ5 = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ5, sac_phase_kernel<true>, kThreads, kSmemBytesT5));
    e->grid_t5 = (occ5 < 1 ? 0 : 1) * e->num_sms;
    e->grid_coop = e->grid_std;

    const long long O = cfg->obs_dim, A = cfg->act_dim, H1 = cfg->hidden1, H2 = cfg->hidden2, B = cfg->batch_size;
    e->ldx = (int)r4(O + A); e->lds = (int)r4(O); e->ldh = (int)r4(2 * A); e->lddpi = (int)r4(A);
    e->tnH2 = (int)((H2 + e->tile_n - 1) / e->tile_n);
    const long long tn16_H2 = (H2 + 15) / 16, tn16_H1 = (H1 + 15) / 16;     // partial-tile counts of the finest tiling (tcgen05: N=16)
    e->row_tiles = (int)((B + kRowsWarp - 1) / kRowsWarp);
    // carve the workspace
    size_t off = 0;
    auto take = [&](long long n) { size_t o = off; off += (size_t)r4(n); return o; };
    struct Slot { float** p; size_t o; };
    std::vector<Slot> slots;
    auto want = [&](float*& p, long long n) { slots.push_back({&p, take(n)}); };
    // staging tiles: two parity buffers each (the folded program gathers the next batch while this one is still in use)
    want(e->xs, 2 * B * e->ldx); want(e->xp, B * e->ldx); want(e->x2, B * e->ldx); want(e->s2buf, 2 * B * e->lds); want(e->rd, 4 * B);
    e->row_blocks = (int)((B + kBM - 1) / kBM);
    const long long RB = e->row_blocks;
    want(e->cp1, B * 2 * H1); want(e->cp2, B * 2 * H2); want(e->qppart, 2LL * tn16_H2 * B);
    want(e->colp_b0, RB * 2 * H1); want(e->colp_b1, RB * 2 * H2); want(e->colp_w2, RB * 2 * H2);
    want(e->colp_ab0, RB * H1); want(e->colp_ab1, RB * H2); want(e->colp_bh, RB * 16);
    want(e->ah1, B * H1); want(e->ah2, B * H2); want(e->heads, B * e->ldh); want(e->logp, 2 * B);
    want(e->ch1, B * 2 * H1); want(e->ch2, B * 2 * H2); want(e->th1, B * 2 * H1); want(e->th2, B * 2 * H2);
    want(e->qpart, 2LL * tn16_H2 * B); want(e->qtpart, 2LL * tn16_H2 * B); want(e->qbuf, 9 * B); want(e->qfw, 2 * B);
    want(e->dh2, B * 2 * H2); want(e->dh1, B * 2 * H1); want(e->dpi, 2 * B * e->lddpi); want(e->ghead, B * e->ldh);
    want(e->adh2, B * H2); want(e->adh1, B * H1);
    want(e->hpart, tn16_H2 * 2 * A * B); want(e->dpart, 2LL * tn16_H1 * A * B);
    want(e->lparts_q, 4LL * e->row_tiles); want(e->lparts_pi, 4LL * e->row_tiles); want(e->loss_scalar, 4);
    want(e->ga, e->lay.actor_floats); want(e->gc, e->lay.critic_floats);
    want(e->peer_flags, 2 * kMaxPeers * 2);          // [2][kMaxPeers] u64 exchange epochs, written by the peers
    e->ws_floats = off;
    CU(cudaMalloc(&e->ws, off * sizeof(float)));
    CU(cudaMemset(e->ws, 0, off * sizeof(float)));
    for (auto& s : slots) *s.p = e->ws + s.o;
    CU(cudaMalloc(&e->idx_buf, (size_t)cfg->max_burst * B * sizeof(long long)));
    e->perm_cap = (size_t)cfg->max_burst * 4 * B;
    CU(cudaMalloc(&e->perm_buf, e->perm_cap * sizeof(int)));
    CU(cudaMalloc(&e->eps_buf, (size_t)cfg->max_burst * 2 * B * A * sizeof(float)));
    CU(cudaMalloc(&e->barrier, sizeof(unsigned long long)));
    CU(cudaMemset(e->barrier, 0, sizeof(unsigned long long)));
    CU(cudaEventCreateWithFlags(&e->stage_ev, cudaEventDisableTiming));
    *out = e;
    return 0;
}

int sacb_destroy(sacb_handle e) {
    if (!e) return 0;
    cudaSetDevice(e->cfg.device);
    cudaDeviceSynchronize();
    if (e->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(e->comm);
    for (int r = 0; r < kMaxPeers; ++r)
        if (e->peer_mode && e->peer_ws[r] && r != e->rank) cudaIpcCloseMemHandle(e->peer_ws[r]);
    cudaFree(e->ws); cudaFree(e->idx_buf); cudaFree(e->perm_buf); cudaFree(e->eps_buf); cudaFree(e->barrier);
    cudaFree(e->d_jobs); cudaFree(e->d_phases); cudaFree(e->sh16); cudaFree(e->d_tmaps); cudaFree(e->d_refresh);
    if (e->stage) cudaFreeHost(e->stage);
    if (e->abort_host) cudaFreeHost(e->abort_host);
    if (e->stage_ev) cudaEventDestroy(e->stage_ev);
    delete e;
    return 0;
}

int sacb_bind_params(sacb_handle e, float* actor, float* critic, float* target) {
    if (!e) return fail(SACB_ERR_INVALID, "null handle");
    for (float* p : {actor, critic, target})
        if (p && ((uintptr_t)p & 15)) return fail(SACB_ERR_INVALID, "parameter arenas must be 16-byte aligned");
    e->pa = actor; e->pc = critic; e->pt = target;
    e->dirty = true;
    return 0;
}

int sacb_bind_adam(sacb_handle e, float* ma, float* va, int64_t step_a, float* mc, float* vc, int64_t step_c) {
    if (!e) return fail(SACB_ERR_INVALID, "null handle");
    for (float* p : {ma, va, mc, vc})
        if (p && ((uintptr_t)p & 15)) return fail(SACB_ERR_INVALID, "Adam arenas must be 16-byte aligned");
    e->ma = ma; e->va = va; e->mc = mc; e->vc = vc;
    e->step_pi = step_a; e->step_q = step_c;
    e->dirty = true;
    return 0;
}

int sacb_get_adam_steps(sacb_handle e, int64_t* sa, int64_t* sc) {
    if (!e) return fail(SACB_ERR_INVALID, "null handle");
    if (sa) *sa = e->step_pi;
    if (sc) *sc = e->step_q;
    return 0;
}

int sacb_set_scalar(sacb_handle e, const char* name, double v) {
    if (!e || !name) return fail(SACB_ERR_INVALID, "null argument");
    sacb_config& c = e->cfg;
    std::string n(name);
    if (n == "alpha") c.alpha = v;
    else if (n == "gamma") c.gamma = v;
    else if (n == "polyak") c.polyak = v;
    else if (n == "reward_scale") c.reward_scale = v;
    else if (n == "lr_pi") c.lr_pi = v;
    else if (n == "lr_q") c.lr_q = v;
    else if (n == "beta1") c.beta1 = v;
    else if (n == "beta2") c.beta2 = v;
    else if (n == "adam_eps") c.adam_eps = v;
    else return fail(SACB_ERR_INVALID, "unknown scalar '%s'", name);
    return 0;
}

int sacb_bind_buffer(sacb_handle e, float* s, float* a, float* r, float* s2, uint8_t* d, int64_t cap) {
    if (!e) return fail(SACB_ERR_INVALID, "null handle");
    if (!s || !a || !r || !s2 || !d || cap < 1) return fail(SACB_ERR_INVALID, "bad buffer arrays");
    e->bs = s; e->ba = a; e->br = r; e->bs2 = s2; e->bd = d; e->cap = cap;
    e->brec = nullptr; e->rec_pitch = 0;
    if (e->ptr >= cap) e->ptr = 0;
    if (e->size > cap) e->size = cap;
    return 0;
}

int sacb_bind_buffer_packed(sacb_handle e, float* rec, int32_t pitch_floats, int64_t cap) {
    if (!e) return fail(SACB_ERR_INVALID, "null handle");
    const int W = 2 * e->cfg.obs_dim + e->cfg.act_dim + 2;
    if (!rec || cap < 1 || pitch_floats < W || (pitch_floats & 3) || ((uintptr_t)rec & 15))
        return fail(SACB_ERR_INVALID, "packed ring: need a 16-byte-aligned array with a pitch that is a multiple of 4 floats and >= %d", W);
    e->brec = rec; e->rec_pitch = pitch_floats; e->cap = cap;
    e->bs = rec; e->ba = rec; e->br = rec; e->bs2 = rec; e->bd = (uint8_t*)rec;      // "bound" markers for the SoA checks
    if (e->ptr >= cap) e->ptr = 0;
    if (e->size > cap) e->size = cap;
    return 0;
}

int sacb_buffer_state(sacb_handle e, int64_t* ptr, int64_t* size, int64_t* cap) {
    if (!e) return fail(SACB_ERR_INVALID, "null handle");
    if (ptr) *ptr = e->ptr;
    if (size) *size = e->size;
    if (cap) *cap = e->cap;
    return 0;
}

int sacb_set_buffer_state(sacb_handle e, int64_t ptr, int64_t size) {
    if (!e) return fail(SACB_ERR_INVALID, "null handle");
    if (ptr < 0 || size < 0 || size > e->cap || (e->cap && ptr >= e->cap)) return fail(SACB_ERR_INVALID, "ptr/size out of range");
    e->ptr = ptr; e->size = size;
    return 0;
}

int sacb_store(sacb_handle e, const float* obs, const float* act, const float* rew, const float* nobs, const uint8_t* done,
               int64_t count, uintptr_t stream) {
    if (!e) return fail(SACB_ERR_INVALID, "null handle");
    if (int r = need(e->bs != nullptr, "replay buffer")) return r;
    if (count < 0) return fail(SACB_ERR_INVALID, "negative count");
    cudaStream_t st = (cudaStream_t)stream;
    CU(cudaSetDevice(e->cfg.device));
    const size_t O = e->cfg.obs_dim, A = e->cfg.act_dim;
    if (e->brec) {
        // packed ring: rows are assembled as records in pinned staging and go to HBM with ONE copy per contiguous ring segment
        const size_t P = (size_t)e->rec_pitch;
        const int64_t chunk = 4096;
        if (!e->stage || e->stage_bytes < (size_t)chunk * P * sizeof(float)) {
            if (e->stage) { CU(cudaEventSynchronize(e->stage_ev)); CU(cudaFreeHost(e->stage)); e->stage = nullptr; }
            e->stage_bytes = (size_t)chunk * P * sizeof(float);
            CU(cudaMallocHost(&e->stage, e->stage_bytes));
        }
        int64_t done_rows = 0;
        while (done_rows < count) {
            int64_t n = count - done_rows;
            if (n > chunk) n = chunk;
            if (n > e->cap - e->ptr) n = e->cap - e->ptr;
            CU(cudaEventSynchronize(e->stage_ev));
            float* rec = (float*)e->stage;
            for (int64_t i = 0; i < n; ++i) {
                float* r = rec + i * P;
                const int64_t g = done_rows + i;
                memcpy(r, obs + g * O, O * sizeof(float));
                memcpy(r + O, act + g * A, A * sizeof(float));
                r[O + A] = rew[g];
                memcpy(r + O + A + 1, nobs + g * O, O * sizeof(float));
                r[2 * O + A + 1] = done[g] ? 1.f : 0.f;
                for (size_t c = 2 * O + A + 2; c < P; ++c) r[c] = 0.f;
            }
            CU(cudaMemcpyAsync(e->brec + e->ptr * P, rec, (size_t)n * P * sizeof(float), cudaMemcpyHostToDevice, st));
            CU(cudaEventRecord(e->stage_ev, st));
            e->ptr = (e->ptr + n) % e->cap;                                   // buffer/replay_buffer.py:42
            e->size = e->size + n > e->cap ? e->cap : e->size + n;            // :43
            done_rows += n;
        }
        return 0;
    }
    const size_t row_bytes = (2 * O + A + 1) * sizeof(float) + 1;
    const int64_t chunk_rows = 4096;
    if (!e->stage) {
        e->stage_bytes = (size_t)chunk_rows * row_bytes + 64;
        CU(cudaMallocHost(&e->stage, e->stage_bytes));
    }
    int64_t done_rows = 0;
    while (done_rows < count) {
        // contiguous segment of the ring: [ptr, min(ptr+n, cap))
        int64_t n = count - done_rows;
        if (n > chunk_rows) n = chunk_rows;
        if (n > e->cap - e->ptr) n = e->cap - e->ptr;
        CU(cudaEventSynchronize(e->stage_ev));            // previous chunk's copies have drained the staging area
        float* so = (float*)e->stage;
        float* sa = so + n * O;
        float* sr = sa + n * A;
        float* sn = sr + n;
        uint8_t* sd = (uint8_t*)(sn + n * O);
        memcpy(so, obs + done_rows * O, n * O * sizeof(float));
        memcpy(sa, act + done_rows * A, n * A * sizeof(float));
        memcpy(sr, rew + done_rows, n * sizeof(float));
        memcpy(sn, nobs + done_rows * O, n * O * sizeof(float));
        for (int64_t i = 0; i < n; ++i) sd[i] = done[done_rows + i] ? 1 : 0;
        CU(cudaMemcpyAsync(e->bs + e->ptr * O, so, n * O * sizeof(float), cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(e->ba + e->ptr * A, sa, n * A * sizeof(float), cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(e->br + e->ptr, sr, n * sizeof(float), cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(e->bs2 + e->ptr * O, sn, n * O * sizeof(float), cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(e->bd + e->ptr, sd, n, cudaMemcpyHostToDevice, st));
        CU(cudaEventRecord(e->stage_ev, st));
        e->ptr = (e->ptr + n) % e->cap;                                   // buffer/replay_buffer.py:42
        e->size = e->size + n > e->cap ? e->cap : e->size + n;            // :43
        done_rows += n;
    }
    return 0;
}

static int draw_indices(sacb_engine* e, long long* out, int n_sets, int B, cudaStream_t st) {
    if (e->size < B) return fail(SACB_ERR_SAMPLE, "Sample larger than population or is negative (size=%lld, batch=%d)", e->size, B);
    if (B > 16384) return fail(SACB_ERR_INVALID, "batch too large for the on-device sampler");
    const unsigned long long seed = e->cfg.seed + 0x51ED270B9F5A3C1DULL * (unsigned long long)(e->rank + 1);
    int* perm = e->perm_buf;
    int* tmp = nullptr;
    if (e->size < 4LL * B && (size_t)n_sets * 4 * B > e->perm_cap) {       // dense regime with an oversize batch
        CU(cudaMallocAsync(&tmp, (size_t)n_sets * 4 * B * sizeof(int), st));
        perm = tmp;
    }
    draw_indices_kernel<<<n_sets, 256, (size_t)B * sizeof(long long), st>>>(out, B, e->size, seed, e->rng_counter, perm);
    CU(cudaGetLastError());
    if (tmp) CU(cudaFreeAsync(tmp, st));
    e->launches++;
    e->rng_counter += n_sets;
    return 0;
}

static int draw_normals(sacb_engine* e, float* out, long long n, cudaStream_t st) {
    const unsigned long long seed = e->cfg.seed + 0xA24BAED4963EE407ULL * (unsigned long long)(e->rank + 1);
    const long long threads = (n + 3) / 4;
    const int grid = (int)((threads + 255) / 256);
    draw_normal_kernel<<<grid, 256, 0, st>>>(out, n, seed, e->rng_counter << 24);
    CU(cudaGetLastError());
    e->launches++;
    e->rng_counter += 1;
    return 0;
}

int sacb_sample(sacb_handle e, int32_t batch, const int64_t* idx_in, float* s, float* a, float* r, float* s2, float* d,
                int64_t* idx_out, uintptr_t stream) {
    if (!e) return fail(SACB_ERR_INVALID, "null handle");
    if (int rr = need(e->bs != nullptr, "replay buffer")) return rr;
    if (batch < 1) return fail(SACB_ERR_SAMPLE, "Sample larger than population or is negative");
    if (e->size < batch) return fail(SACB_ERR_SAMPLE, "Sample larger than population or is negative (size=%lld, batch=%d)", e->size, batch);
    cudaStream_t st = (cudaStream_t)stream;
    CU(cudaSetDevice(e->cfg.device));
    const long long* idx = (const long long*)idx_in;
    long long* tmp = nullptr;
    if (!idx) {
        if ((long long)batch > (long long)e->cfg.max_burst * e->cfg.batch_size) {
            CU(cudaMallocAsync(&tmp, (size_t)batch * sizeof(long long), st));
            if (int rr = draw_indices(e, tmp, 1, batch, st)) return rr;
            idx = tmp;
        } else {
            if (int rr = draw_indices(e, e->idx_buf, 1, batch, st)) return rr;
            idx = e->idx_buf;
        }
    }
    if (e->brec)
        sample_gather_packed_kernel<<<(batch + 7) / 8, 256, 0, st>>>(idx, batch, e->cfg.obs_dim, e->cfg.act_dim, e->brec, e->rec_pitch, s, a, r, s2, d,
                                                                     (long long*)idx_out);
    else
        sample_gather_kernel<<<(batch + 7) / 8, 256, 0, st>>>(idx, batch, e->cfg.obs_dim, e->cfg.act_dim, e->bs, e->ba, e->br, e->bs2,
                                                              e->bd, s, a, r, s2, d, (long long*)idx_out);
    CU(cudaGetLastError());
    e->launches++;
    if (tmp) CU(cudaFreeAsync(tmp, st));
    return 0;
}

// which: 1 = critic step (needs target + critic Adam), 2 = policy step (needs actor Adam), 3 = both
static int check_train_bound(sacb_engine* e, int which) {
    if (int r = need(e->pa && e->pc, "actor/critic parameters")) return r;
    if (which & 1) {
        if (int r = need(e->pt != nullptr, "target parameters")) return r;
        if (int r = need(e->mc && e->vc, "critic Adam state")) return r;
    }
    if (which & 2)
        if (int r = need(e->ma && e->va, "actor Adam state")) return r;
    return 0;
}

static int allreduce(sacb_engine* e, float* buf, long long n, cudaStream_t st) {
    NC(g_nccl.AllReduce(buf, buf, (size_t)n, kNcclFloat, kNcclSum, e->comm, st));
    return 0;
}

int sacb_update(sacb_handle e, int32_t n_updates, const int64_t* idx_in, const float* eps_in, float* loss_q, float* loss_pi,
                uintptr_t stream) {
    if (!e) return fail(SACB_ERR_INVALID, "null handle");
    if (int r = check_train_bound(e, 3)) return r;
    if (int r = need(e->bs != nullptr, "replay buffer")) return r;
    if (n_updates < 0 || n_updates > e->cfg.max_burst) return fail(SACB_ERR_INVALID, "n_updates %d exceeds max_burst %d", n_updates, e->cfg.max_burst);
    if (n_updates == 0) return 0;
    if (e->abort_host && *(volatile unsigned int*)e->abort_host)
        return fail(SACB_ERR_PEER, "a previous data-parallel burst timed out waiting for rank %u", *(volatile unsigned int*)e->abort_host - 1);
    cudaStream_t st = (cudaStream_t)stream;
    CU(cudaSetDevice(e->cfg.device));
    if (int r = upload_programs(e, st)) return r;
    const int B = e->cfg.batch_size, A = e->cfg.act_dim;
    if (e->size < B) return fail(SACB_ERR_SAMPLE, "Sample larger than population or is negative (size=%lld, batch=%d)", e->size, B);
    const long long* idx = (const long long*)idx_in;
    if (!idx) {
        if (int r = draw_indices(e, e->idx_buf, n_updates, B, st)) return r;
        idx = e->idx_buf;
    }
    const float* eps = eps_in;
    if (!eps) {
        if (int r = draw_normals(e, e->eps_buf, (long long)n_updates * 2 * B * A, st)) return r;
        eps = e->eps_buf;
    }
    RunArgs ra;
    fill_args(e, ra);
    ra.idx = idx; ra.idx_stride = B;
    ra.eps1 = eps; ra.eps2 = eps + (size_t)B * A; ra.eps_stride = 2LL * B * A;
    ra.loss_q_out = loss_q; ra.loss_pi_out = loss_pi;
    ra.fuse_polyak = 1; ra.pass_mask = 3; ra.k_total = n_updates;
    if (e->world <= 1 || e->peer_mode) {
        // single GPU, or peer mode: the gradient exchange is fused into the Adam phases (NVLink loads from the peers'
        // arenas), so a data-parallel burst is still ONE persistent launch -- of the folded 8-phase program when the
        // configuration allows it (build_fold_program), else of the first 16-phase program
        const ProgRange& pr = e->fold_ok ? e->p_fold : e->p_update;
        if (int r = run_phases(e, ra, pr.begin, pr.end, 0, n_updates, st)) return r;
        e->last_parity = e->fold_ok ? ((n_updates - 1) & 1) : 0;
    } else {
        const bool sync_actor = !(e->cfg.flags & SACB_FLAG_NO_ACTOR_GRAD_SYNC);
        for (int k = 0; k < n_updates; ++k) {
            if (int r = run_phases(e, ra, e->p_update.begin, e->ph_crit_adam, k, k + 1, st)) return r;
            if (int r = allreduce(e, e->gc, e->lay.critic_floats, st)) return r;         // mpi_avg_grads(critic), algorithm.py:138
            if (sync_actor) {
                if (int r = run_phases(e, ra, e->ph_crit_adam, e->ph_pi_adam, k, k + 1, st)) return r;
                if (int r = allreduce(e, e->ga, e->lay.actor_floats, st)) return r;      // intent of algorithm.py:155
                if (int r = run_phases(e, ra, e->ph_pi_adam, e->p_update.end, k, k + 1, st)) return r;
            } else {
                RunArgs rb = ra;
                if (int r = run_phases(e, rb, e->ph_crit_adam, e->ph_pi_adam, k, k + 1, st)) return r;
                rb.grad_div = 1.f;
                if (int r = run_phases(e, rb, e->ph_pi_adam, e->p_update.end, k, k + 1, st)) return r;
            }
        }
    }
    e->step_pi += n_updates; e->step_q += n_updates;
    e->exch_pi += n_updates; e->exch_q += n_updates;
    return 0;
}

int sacb_load_batch(sacb_handle e, const float* s, const float* a, const float* r, const float* s2, const float* d, uintptr_t stream) {
    if (!e) return fail(SACB_ERR_INVALID, "null handle");
    if (!s || !a || !r || !s2 || !d) return fail(SACB_ERR_INVALID, "null batch array");
    cudaStream_t st = (cudaStream_t)stream;
    CU(cudaSetDevice(e->cfg.device));
    const int B = e->cfg.batch_size;
    load_batch_kernel<<<(B + 7) / 8, 256, 0, st>>>(s, a, r, s2, d, B, e->cfg.obs_dim, e->cfg.act_dim, e->xs, e->xp, e->x2, e->s2buf,
                                                   e->rd, e->ldx, e->lds);
    CU(cudaGetLastError());
    e->launches++;
    e->last_parity = 0;
    return 0;
}

static int eps_or_draw(sacb_engine* e, const float* eps_in, const float** eps_out, cudaStream_t st) {
    if (eps_in) { *eps_out = eps_in; return 0; }
    const long long n = (long long)e->cfg.batch_size * e->cfg.act_dim;
    if (int r = draw_normals(e, e->eps_buf, n, st)) return r;
    *eps_out = e->eps_buf;
    return 0;
}

int sacb_update_critic(sacb_handle e, const float* eps_in, float* loss_q, uintptr_t stream) {
    if (!e) return fail(SACB_ERR_INVALID, "null handle");
    if (int r = check_train_bound(e, 1)) return r;
    cudaStream_t st = (cudaStream_t)stream;
    CU(cudaSetDevice(e->cfg.device));
    if (int r = upload_programs(e, st)) return r;
    const float* eps = nullptr;
    if (int r = eps_or_draw(e, eps_in, &eps, st)) return r;
    RunArgs ra;
    fill_args(e, ra);
    ra.eps1 = eps; ra.eps2 = nullptr; ra.eps_stride = 0; ra.pass_mask = 1; ra.fuse_polyak = 0;
    ra.loss_q_out = loss_q;
    if (e->world <= 1 || e->peer_mode) {
        if (int r = run_phases(e, ra, e->ph_crit_begin, e->ph_crit_adam + 1, 0, 1, st)) return r;
    } else {
        if (int r = run_phases(e, ra, e->ph_crit_begin, e->ph_crit_adam, 0, 1, st)) return r;
        if (int r = allreduce(e, e->gc, e->lay.critic_floats, st)) return r;
        if (int r = run_phases(e, ra, e->ph_crit_adam, e->ph_crit_adam + 1, 0, 1, st)) return r;
    }
    e->step_q += 1; e->exch_q += 1;
    return 0;
}

int sacb_update_policy(sacb_handle e, const float* eps_in, float* loss_pi, uintptr_t stream) {
    if (!e) return fail(SACB_ERR_INVALID, "null handle");
    if (int r = check_train_bound(e, 2)) return r;
    cudaStream_t st = (cudaStream_t)stream;
    CU(cudaSetDevice(e->cfg.device));
    if (int r = upload_programs(e, st)) return r;
    const float* eps = nullptr;
    if (int r = eps_or_draw(e, eps_in, &eps, st)) return r;
    RunArgs ra;
    fill_args(e, ra);
    ra.eps1 = nullptr; ra.eps2 = eps; ra.eps_stride = 0; ra.pass_mask = 2; ra.fuse_polyak = 0;
    ra.loss_pi_out = loss_pi;
    const int adam = e->p_policy.end - 1;
    const bool sync_actor = e->world > 1 && !e->peer_mode && !(e->cfg.flags & SACB_FLAG_NO_ACTOR_GRAD_SYNC);
    if (e->peer_mode) {
        if (int r = run_phases(e, ra, e->p_policy.begin, e->p_policy.end, 0, 1, st)) return r;
    } else if (!sync_actor) {
        ra.grad_div = 1.f;
        if (int r = run_phases(e, ra, e->p_policy.begin, e->p_policy.end, 0, 1, st)) return r;
    } else {
        if (int r = run_phases(e, ra, e->p_policy.begin, adam, 0, 1, st)) return r;
        if (int r = allreduce(e, e->ga, e->lay.actor_floats, st)) return r;
        if (int r = run_phases(e, ra, adam, e->p_policy.end, 0, 1, st)) return r;
    }
    e->step_pi += 1; e->exch_pi += 1;
    return 0;
}

int sacb_update_targets(sacb_handle e, uintptr_t stream) {
    if (!e) return fail(SACB_ERR_INVALID, "null handle");
    if (int r = need(e->pc && e->pt, "critic/target parameters")) return r;
    cudaStream_t st = (cudaStream_t)stream;
    CU(cudaSetDevice(e->cfg.device));
    if (int r = upload_programs(e, st)) return r;
    RunArgs ra;
    fill_args(e, ra);
    return run_phases(e, ra, e->p_targets.begin, e->p_targets.end, 0, 1, st);
}

int sacb_actor_forward(sacb_handle e, const float* x, int32_t n, const float* eps_in, int32_t deterministic, float* pi, float* logp,
                       uintptr_t stream) {
    if (!e) return fail(SACB_ERR_INVALID, "null handle");
    if (int r = need(e->pa != nullptr, "actor parameters")) return r;
    if (n < 0 || !x || !pi) return fail(SACB_ERR_INVALID, "bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    CU(cudaSetDevice(e->cfg.device));
    if (int r = upload_programs(e, st)) return r;
    const int B = e->cfg.batch_size, O = e->cfg.obs_dim, A = e->cfg.act_dim;
    {   // acting path (a handful of rows, e.g. one env step): one launch of an 8-CTA cluster, activations in DSMEM (sacb_act.cuh)
        const size_t smem = act_smem_floats(n, O, A, e->cfg.hidden1, e->cfg.hidden2) * sizeof(float);
        if (n >= 1 && n <= kActMaxRows && e->mode == 0 && smem <= 200 * 1024) {
            const sacb_layout_info& L = e->lay;
            ActArgs a;
            memset(&a, 0, sizeof a);
            a.x = x; a.w0 = e->pa + L.actor_w0; a.b0 = e->pa + L.actor_b0; a.w1 = e->pa + L.actor_w1; a.b1 = e->pa + L.actor_b1;
            a.wh = e->pa + L.actor_w_mu; a.bh = e->pa + L.actor_b_mu;
            a.eps = deterministic ? nullptr : eps_in; a.pi = pi; a.logp = logp;
            a.n = n; a.O = O; a.A = A; a.H1 = e->cfg.hidden1; a.H2 = e->cfg.hidden2; a.deterministic = deterministic ? 1 : 0;
            a.seed = e->cfg.seed + 0xA24BAED4963EE407ULL * (unsigned long long)(e->rank + 1);
            a.counter = (e->rng_counter++) << 24;
            a.log_min_std = e->cfg.log_min_std; a.log_max_std = e->cfg.log_max_std; a.act_limit = e->cfg.act_limit; a.flags = e->cfg.flags;
            actor_act_kernel<<<kActCluster, 256, smem, st>>>(a);
            CU(cudaGetLastError());
            e->launches++;
            return 0;
        }
    }
    for (int r0 = 0; r0 < n; r0 += B) {
        const int m = (n - r0) < B ? (n - r0) : B;
        CU(cudaMemcpy2DAsync(e->s2buf, e->lds * sizeof(float), x + (size_t)r0 * O, O * sizeof(float), O * sizeof(float), m,
                             cudaMemcpyDeviceToDevice, st));
        const float* eps = nullptr;
        if (!deterministic) {
            if (eps_in) eps = eps_in + (size_t)r0 * A;
            else {
                if (int r = draw_normals(e, e->eps_buf, (long long)m * A, st)) return r;
                eps = e->eps_buf;
            }
        }
        RunArgs ra;
        fill_args(e, ra);
        ra.m_rows = m; ra.pass_mask = 1; ra.eps1 = eps; ra.eps2 = nullptr; ra.eps_stride = 0;
        if (int r = run_phases(e, ra, e->p_actor_fwd.begin, e->p_actor_fwd.end, 0, 1, st)) return r;
        CU(cudaMemcpy2DAsync(pi + (size_t)r0 * A, A * sizeof(float), e->x2 + O, e->ldx * sizeof(float), A * sizeof(float), m,
                             cudaMemcpyDeviceToDevice, st));
        if (logp) CU(cudaMemcpyAsync(logp + r0, e->logp, m * sizeof(float), cudaMemcpyDeviceToDevice, st));
    }
    return 0;
}

int sacb_critic_forward(sacb_handle e, int32_t which, const float* s, const float* a, int32_t n, float* q1, float* q2, uintptr_t stream) {
    if (!e) return fail(SACB_ERR_INVALID, "null handle");
    if (which != 0 && which != 1) return fail(SACB_ERR_INVALID, "which must be 0 (critic) or 1 (target)");
    if (int r = need((which ? e->pt : e->pc) != nullptr, which ? "target parameters" : "critic parameters")) return r;
    if (n < 0 || !s || !a || !q1 || !q2) return fail(SACB_ERR_INVALID, "bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    CU(cudaSetDevice(e->cfg.device));
    if (int r = upload_programs(e, st)) return r;
    const int B = e->cfg.batch_size, O = e->cfg.obs_dim, A = e->cfg.act_dim;
    for (int r0 = 0; r0 < n; r0 += B) {
        const int m = (n - r0) < B ? (n - r0) : B;
        CU(cudaMemcpy2DAsync(e->xs, e->ldx * sizeof(float), s + (size_t)r0 * O, O * sizeof(float), O * sizeof(float), m,
                             cudaMemcpyDeviceToDevice, st));
        CU(cudaMemcpy2DAsync(e->xs + O, e->ldx * sizeof(float), a + (size_t)r0 * A, A * sizeof(float), A * sizeof(float), m,
                             cudaMemcpyDeviceToDevice, st));
        RunArgs ra;
        fill_args(e, ra);
        ra.m_rows = m;
        if (int r = run_phases(e, ra, e->p_critic_fwd[which].begin, e->p_critic_fwd[which].end, 0, 1, st)) return r;
        CU(cudaMemcpyAsync(q1 + r0, e->qfw, m * sizeof(float), cudaMemcpyDeviceToDevice, st));
        CU(cudaMemcpyAsync(q2 + r0, e->qfw + B, m * sizeof(float), cudaMemcpyDeviceToDevice, st));
    }
    return 0;
}

int sacb_comm_unique_id(const char* lib, uint8_t id_out[128]) {
    if (int r = load_nccl(lib)) return r;
    nccl_uid_t id;
    NC(g_nccl.GetUniqueId(&id));
    memcpy(id_out, id.internal, 128);
    return 0;
}

int sacb_comm_init(sacb_handle e, const char* lib, const uint8_t id_in[128], int32_t rank, int32_t world) {
    if (!e) return fail(SACB_ERR_INVALID, "null handle");
    if (world < 1 || rank < 0 || rank >= world) return fail(SACB_ERR_INVALID, "bad rank/world");
    if (int r = load_nccl(lib)) return r;
    CU(cudaSetDevice(e->cfg.device));
    nccl_uid_t id;
    memcpy(id.internal, id_in, 128);
    NC(g_nccl.CommInitRank(&e->comm, world, id, rank));
    e->rank = rank; e->world = world;
    e->dirty = true;                                   // the phase programs depend on the world size (fused vs separate Adam)
    return 0;
}

int sacb_peer_export(sacb_handle e, uint8_t handle_out[64], int64_t* ws_floats) {
    if (!e || !handle_out) return fail(SACB_ERR_INVALID, "null argument");
    CU(cudaSetDevice(e->cfg.device));
    cudaIpcMemHandle_t hnd;
    static_assert(sizeof(hnd) == 64, "cudaIpcMemHandle_t is 64 bytes");
    CU(cudaIpcGetMemHandle(&hnd, e->ws));
    memcpy(handle_out, &hnd, 64);
    if (ws_floats) *ws_floats = (int64_t)e->ws_floats;
    return 0;
}

int sacb_peer_bind(sacb_handle e, const uint8_t* handles, int32_t rank, int32_t world) {
    if (!e || !handles) return fail(SACB_ERR_INVALID, "null argument");
    if (world < 2 || world > kMaxPeers || rank < 0 || rank >= world) return fail(SACB_ERR_INVALID, "peer mode needs 2..%d ranks", kMaxPeers);
    if (e->cfg.flags & SACB_FLAG_NO_ACTOR_GRAD_SYNC)
        return fail(SACB_ERR_INVALID, "peer mode relies on both exchanges per update to interlock the ranks; use the NCCL communicator with SACB_FLAG_NO_ACTOR_GRAD_SYNC");
    CU(cudaSetDevice(e->cfg.device));
    for (int r = 0; r < world; ++r) {
        if (r == rank) { e->peer_ws[r] = e->ws; continue; }
        cudaIpcMemHandle_t hnd;
        memcpy(&hnd, handles + 64 * r, 64);
        void* p = nullptr;
        CU(cudaIpcOpenMemHandle(&p, hnd, cudaIpcMemLazyEnablePeerAccess));
        e->peer_ws[r] = (float*)p;
    }
    if (!e->abort_host) {
        CU(cudaHostAlloc((void**)&e->abort_host, sizeof(unsigned int), cudaHostAllocMapped));
        *e->abort_host = 0;
        CU(cudaHostGetDevicePointer((void**)&e->abort_dev, e->abort_host, 0));
    }
    if (const char* t = getenv("SACB_PEER_TIMEOUT_MS")) {
        const long long ms = atoll(t);
        if (ms > 0) e->peer_timeout_ns = (unsigned long long)ms * 1000000ULL;
    }
    e->rank = rank; e->world = world; e->peer_mode = 1;
    e->dirty = true;
    return 0;
}

// 0 = healthy.  Synchronises the stream first, so that a timeout raised by the launch in flight is seen.
int sacb_peer_status(sacb_handle e, uintptr_t stream) {
    if (!e) return fail(SACB_ERR_INVALID, "null handle");
    if (!e->peer_mode || !e->abort_host) return 0;
    CU(cudaSetDevice(e->cfg.device));
    CU(cudaStreamSynchronize((cudaStream_t)stream));
    const unsigned ab = *(volatile unsigned int*)e->abort_host;
    if (ab) return fail(SACB_ERR_PEER, "in-kernel gradient exchange timed out after %.1f s waiting for rank %u (a peer died or stalled); "
                                       "parameters of this replica are no longer valid", e->peer_timeout_ns * 1e-9, ab - 1);
    return 0;
}

int sacb_sync_params(sacb_handle e, uintptr_t stream) {
    if (!e) return fail(SACB_ERR_INVALID, "null handle");
    if (e->world <= 1) return 0;                      // sync_params early return, sac/mpi.py:94-95
    if (!e->comm) return fail(SACB_ERR_STATE, "sacb_sync_params needs the NCCL communicator (sacb_comm_init)");
    cudaStream_t st = (cudaStream_t)stream;
    CU(cudaSetDevice(e->cfg.device));
    if (e->pa) NC(g_nccl.Broadcast(e->pa, e->pa, (size_t)e->lay.actor_floats, kNcclFloat, 0, e->comm, st));
    if (e->pc) NC(g_nccl.Broadcast(e->pc, e->pc, (size_t)e->lay.critic_floats, kNcclFloat, 0, e->comm, st));
    if (e->pt) NC(g_nccl.Broadcast(e->pt, e->pt, (size_t)e->lay.critic_floats, kNcclFloat, 0, e->comm, st));
    return 0;
}

int sacb_debug_read(sacb_handle e, const char* name, float* dst, int64_t max_floats, int64_t* n_out) {
    if (!e || !name || !dst) return fail(SACB_ERR_INVALID, "null argument");
    const long long B = e->cfg.batch_size, A = e->cfg.act_dim, O = e->cfg.obs_dim;
    std::string n(name);
    const float* src = nullptr;
    long long cnt = 0, rows = 0, cols = 0, ld = 0;
    if (n == "q1") { src = e->qbuf; cnt = B; }
    else if (n == "q2") { src = e->qbuf + B; cnt = B; }
    else if (n == "q1_target") { src = e->qbuf + 2 * B; cnt = B; }
    else if (n == "q2_target") { src = e->qbuf + 3 * B; cnt = B; }
    else if (n == "backup") { src = e->qbuf + 4 * B; cnt = B; }
    else if (n == "dq1") { src = e->qbuf + 5 * B; cnt = B; }
    else if (n == "dq2") { src = e->qbuf + 6 * B; cnt = B; }
    else if (n == "q1_pi") { src = e->qbuf + 7 * B; cnt = B; }
    else if (n == "q2_pi") { src = e->qbuf + 8 * B; cnt = B; }
    else if (n == "logp_next") { src = e->logp; cnt = B; }
    else if (n == "logp_pi") { src = e->logp + B; cnt = B; }
    else if (n == "a2") { src = e->x2 + O; rows = B; cols = A; ld = e->ldx; }
    else if (n == "pi") { src = e->xp + O; rows = B; cols = A; ld = e->ldx; }
    else if (n == "grad_actor") { src = e->ga; cnt = e->lay.actor_floats; }
    else if (n == "grad_critic") { src = e->gc; cnt = e->lay.critic_floats; }
    else if (n == "loss") { src = e->loss_scalar; cnt = 2; }
    else if (n == "operand_path") {     // [0] = 1 when precision 3 runs on TMA-fed bf16 shadow operands (csrc/sacb_tma.cuh), [1] = tensor maps
        if (max_floats < 2) return fail(SACB_ERR_INVALID, "destination too small");
        CU(cudaSetDevice(e->cfg.device));
        if (int r = upload_programs(e, nullptr)) return r;
        dst[0] = (e->bf_tma && e->tf32x3 == 3) ? 1.f : 0.f; dst[1] = (float)e->tmaps.size();
        if (n_out) *n_out = 2;
        return 0;
    }
    else if (n == "rewards") { src = e->rd + (size_t)e->last_parity * 2 * B; cnt = B; }
    else if (n == "done") { src = e->rd + (size_t)e->last_parity * 2 * B + B; cnt = B; }
    else return fail(SACB_ERR_INVALID, "unknown debug array '%s'", name);
    CU(cudaSetDevice(e->cfg.device));
    CU(cudaDeviceSynchronize());
    if (rows) {
        if (rows * cols > max_floats) return fail(SACB_ERR_INVALID, "destination too small");
        CU(cudaMemcpy2D(dst, cols * sizeof(float), src, ld * sizeof(float), cols * sizeof(float), rows, cudaMemcpyDeviceToHost));
        if (n_out) *n_out = rows * cols;
    } else {
        if (cnt > max_floats) return fail(SACB_ERR_INVALID, "destination too small");
        CU(cudaMemcpy(dst, src, cnt * sizeof(float), cudaMemcpyDeviceToHost));
        if (n_out) *n_out = cnt;
    }
    return 0;
}

int sacb_debug_profile(sacb_handle e, int32_t n_updates, uint64_t* out, int64_t max_words, int32_t* n_phases, int32_t* n_ctas,
                       uintptr_t stream) {
    if (!e || !out) return fail(SACB_ERR_INVALID, "null argument");
    if (int r = check_train_bound(e, 3)) return r;
    if (int r = need(e->bs != nullptr, "replay buffer")) return r;
    if (n_updates < 1 || n_updates > e->cfg.max_burst) return fail(SACB_ERR_INVALID, "bad n_updates");
    if (e->world > 1 && !e->peer_mode) return fail(SACB_ERR_INVALID, "profile needs a single launch per burst: one GPU or peer mode");
    cudaStream_t st = (cudaStream_t)stream;
    CU(cudaSetDevice(e->cfg.device));
    if (int r = upload_programs(e, st)) return r;
    const ProgRange& pr = e->fold_ok ? e->p_fold : e->p_update;
    const int np = pr.end - pr.begin;
    const long long words = (long long)n_updates * np * e->grid_coop * 2;
    if (words > max_words) return fail(SACB_ERR_INVALID, "stamp buffer too small: need %lld words", words);
    unsigned long long* d = nullptr;
    CU(cudaMalloc(&d, words * sizeof(unsigned long long)));
    CU(cudaMemset(d, 0, words * sizeof(unsigned long long)));
    const int B = e->cfg.batch_size, A = e->cfg.act_dim;
    if (int r = draw_indices(e, e->idx_buf, n_updates, B, st)) return r;
    if (int r = draw_normals(e, e->eps_buf, (long long)n_updates * 2 * B * A, st)) return r;
#ifdef SACB_TRACE
    { int z = 0; cudaMemcpyToSymbol(g_trace_n, &z, sizeof(int));
      const char* tc = getenv("SACB_TRACE_CTA"); int cta = tc ? atoi(tc) : 0; cudaMemcpyToSymbol(g_trace_cta, &cta, sizeof(int)); }
#endif
    RunArgs ra;
    fill_args(e, ra);
    ra.idx = e->idx_buf; ra.idx_stride = B;
    ra.eps1 = e->eps_buf; ra.eps2 = e->eps_buf + (size_t)B * A; ra.eps_stride = 2LL * B * A;
    ra.fuse_polyak = 1; ra.pass_mask = 3; ra.prof = d; ra.k_total = n_updates;
    const int saved_mode = e->mode;
    e->mode = 0;
    int rc = run_phases(e, ra, pr.begin, pr.end, 0, n_updates, st);
    e->mode = saved_mode;
    e->last_parity = e->fold_ok ? ((n_updates - 1) & 1) : 0;
    if (rc) { cudaFree(d); return rc; }
    CU(cudaStreamSynchronize(st));
#ifdef SACB_TRACE
    {
        static unsigned long long h[1 << 16];
        int n = 0;
        cudaMemcpyFromSymbol(&n, g_trace_n, sizeof(int));
        cudaMemcpyFromSymbol(h, g_trace, sizeof(h));
        fprintf(stderr, "TRACE %d events (id, dcycles)\n", n);
        for (int i = 1; i < n; ++i)
            fprintf(stderr, "%llu %lld\n", h[2 * i], (long long)(h[2 * i + 1] - h[2 * i - 1]));
    }
#endif
    CU(cudaMemcpy(out, d, words * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    CU(cudaFree(d));
    e->step_pi += n_updates; e->step_q += n_updates;
    e->exch_pi += n_updates; e->exch_q += n_updates;
    if (n_phases) *n_phases = np;
    if (n_ctas) *n_ctas = e->grid_coop;
    return 0;
}

int64_t sacb_launch_count(sacb_handle e) { return e ? e->launches : 0; }

int sacb_set_precision(sacb_handle e, int32_t tf32x3) {
    if (!e) return fail(SACB_ERR_INVALID, "null handle");
    if (tf32x3 < 0 || tf32x3 > 3)
        return fail(SACB_ERR_INVALID, "precision must be 0 (fp32 FFMA), 1 (3xTF32 mma.sync), 2 (3xTF32 tcgen05) or 3 (bf16 tcgen05)");
    if (tf32x3 >= 2 && e->grid_t5 < 1) return fail(SACB_ERR_CUDA, "tcgen05 path does not fit on this device");
    e->tf32x3 = tf32x3;
    e->tile_m = tf32x3 == 3 ? kBfM : tf32x3 == 2 ? kT5M : kBM;
    e->tile_n = tf32x3 == 3 ? kBfN : tf32x3 == 2 ? kT5N : kBN;
    e->tnH2 = (e->cfg.hidden2 + e->tile_n - 1) / e->tile_n;
    e->grid_coop = tf32x3 >= 2 ? e->grid_t5 : e->grid_std;
    e->dirty = true;                                   // tile shapes are baked into the phase programs
    return 0;
}

int sacb_set_mode(sacb_handle e, int32_t mode) {
    if (!e) return fail(SACB_ERR_INVALID, "null handle");
    if (mode != 0 && mode != 1) return fail(SACB_ERR_INVALID, "mode must be 0 or 1");
    e->mode = mode;
    return 0;
}

}  // extern "C"
