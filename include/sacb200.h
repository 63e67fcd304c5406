/*
 * sacb200.h -- C ABI of the B200-native SAC update engine (libsacb200.so).
 *
 * The reference (dogeplusplus/torch-actor-critic) has no FFI: its boundary for the hot path is the Python
 * call surface  ReplayBuffer.store/sample  (buffer/replay_buffer.py:29-54),  Actor / DoubleCritic .forward
 * (networks/linear.py:32-53,61-79),  SAC.update_critic / update_policy (sac/algorithm.py:115-162),
 * update_targets (sac/algorithm.py:77-81), the burst loop in SAC.train (sac/algorithm.py:273-281) and the
 * gradient/parameter sync of sac/mpi.py:77-98.  Each entry point below names the reference code it replaces.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes only.  Every function returns 0 on success or a negative
 *     sacb_status; the message is available from sacb_last_error() (thread-local).  No exception crosses the ABI.
 *   - All `dev` pointers are CUDA device pointers on the handle's device that the CALLER owns (in the Python
 *     host layer: `tensor.data_ptr()` of contiguous torch CUDA tensors).  The engine borrows them until
 *     sacb_destroy / re-bind.  `host` pointers are ordinary host memory.
 *   - Every call that enqueues work takes the CUDA stream as a uintptr_t (`cudaStream_t`, 0 = legacy default)
 *     and is asynchronous with respect to the host unless stated otherwise.
 *   - A handle is not thread-safe; one handle per device.
 *   - There is NO CPU fallback: with no usable CUDA device sacb_create fails with SACB_ERR_CUDA.
 */
#ifndef SACB200_H
#define SACB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SACB_ABI_VERSION 1

typedef enum {
    SACB_OK = 0,
    SACB_ERR_INVALID = -1,   /* bad argument / unsupported shape            */
    SACB_ERR_CUDA = -2,      /* CUDA runtime error (message has the detail) */
    SACB_ERR_STATE = -3,     /* call out of order (e.g. update before bind)  */
    SACB_ERR_SAMPLE = -4,    /* sample larger than population (reference: ValueError, replay_buffer.py:46) */
    SACB_ERR_NCCL = -5,
    SACB_ERR_PEER = -6       /* in-kernel gradient exchange timed out on a dead / stalled rank (reference: a blocking
                                comm.recv that never returns, sac/algorithm.py:264-265) */
} sacb_status;

/* quirk switches.  Default 0 = reproduce the reference on one rank exactly.  One INTENTIONAL
 * deviation with world > 1: the default averages the actor gradient across ranks (the reference's evident intent; it
 * calls mpi_avg_grads(actor) BEFORE backward(), sac/algorithm.py:155-156, so it effectively never does).
 * SACB_FLAG_NO_ACTOR_GRAD_SYNC reproduces the reference's behaviour; it needs the NCCL transport (SACB_COMM=nccl),
 * because the in-kernel peer exchange relies on both exchanges per update to interlock the ranks. */
#define SACB_FLAG_CANONICAL_LOGPROB 1u   /* 2*(ln2-u-softplus(-2u)) instead of networks/linear.py:51 as written   */
#define SACB_FLAG_PI_FROM_STATES    2u   /* policy at `states` instead of `next_states` (sac/algorithm.py:37,148-153) */
#define SACB_FLAG_NO_ACTOR_GRAD_SYNC 4u  /* multi-rank: skip the actor allreduce, as the reference effectively does
                                            (mpi_avg_grads before backward, sac/algorithm.py:155-156)              */

typedef struct sacb_config {
    uint32_t struct_size;      /* = sizeof(sacb_config), ABI check */
    int32_t device;            /* CUDA device ordinal */
    int32_t obs_dim;           /* O */
    int32_t act_dim;           /* A */
    int32_t hidden1, hidden2;  /* MLP widths; the reference uses [256,256] (main.py:61) */
    int32_t batch_size;        /* B, rows per gradient update (main.py:153; BASELINE uses 256) */
    int32_t max_burst;         /* max updates per sacb_update call (reference update_every = 50, main.py:157) */
    float log_min_std, log_max_std, act_limit;       /* networks/linear.py:19-21 */
    int32_t reserved0;
    double alpha, gamma, polyak, reward_scale;       /* main.py:148-151; python floats in the reference, so doubles here:
                                                        (1 - polyak) is formed in double before it is rounded to f32 */
    double lr_pi, lr_q, beta1, beta2, adam_eps;      /* torch.optim.Adam, main.py:93-95 */
    uint32_t flags;            /* SACB_FLAG_* */
    uint64_t seed;             /* Philox seed for on-device index / noise draws */
    int32_t rank, world;       /* data-parallel rank / size (sac/mpi.py:37-43); Philox subsequence space is per rank */
} sacb_config;

/* Element offsets (in floats) of each parameter tensor inside the flat arenas the caller allocates.
 * The same offsets apply to the Adam exp_avg / exp_avg_sq arenas and to the engine's gradient arenas. */
typedef struct sacb_layout_info {
    int64_t actor_floats;      /* arena length (padded) */
    int64_t actor_w0, actor_b0, actor_w1, actor_b1;      /* layers.{0,1}.{weight,bias}  (networks/linear.py:23) */
    int64_t actor_w_mu, actor_w_ls, actor_b_mu, actor_b_ls; /* mu_layer / log_std_layer    (networks/linear.py:25-27) */
    int64_t critic_floats;
    int64_t critic_w0[2], critic_b0[2], critic_w1[2], critic_b1[2], critic_w2[2], critic_b2[2]; /* q1,q2 .layers.{0,1,2} */
} sacb_layout_info;

typedef struct sacb_engine* sacb_handle;

int sacb_abi_version(void);
const char* sacb_last_error(void);

/* Arena layout for a configuration (pure host computation). */
int sacb_layout(const sacb_config* cfg, sacb_layout_info* out);

/* Create / destroy.  Allocates the engine's private workspace (activations, gradient arenas, index / noise
 * staging, grid barrier) on cfg->device.  Replaces the object construction in main.py:54-97,140-166. */
int sacb_create(const sacb_config* cfg, sacb_handle* out);
int sacb_destroy(sacb_handle h);

/* Bind caller-owned flat arenas (layout: sacb_layout).  Any of the three may be NULL to leave it unbound
 * (a forward-only handle for one module binds only that module).
 * Replaces the torch module parameters of Actor / DoubleCritic / deepcopy(critic) (sac/algorithm.py:194). */
int sacb_bind_params(sacb_handle h, float* actor_dev, float* critic_dev, float* target_dev);
/* Adam moments (exp_avg, exp_avg_sq) for actor and critic and the step counts taken so far
 * (torch.optim.Adam state, main.py:93-95; checkpointed at sac/algorithm.py:176-178). */
int sacb_bind_adam(sacb_handle h, float* m_actor_dev, float* v_actor_dev, int64_t step_actor,
                   float* m_critic_dev, float* v_critic_dev, int64_t step_critic);
int sacb_get_adam_steps(sacb_handle h, int64_t* step_actor, int64_t* step_critic);
/* Change a scalar hyper-parameter after creation (name = "alpha","gamma","polyak","reward_scale","lr_pi","lr_q",
 * "beta1","beta2","adam_eps").  Takes effect at the next enqueue. */
int sacb_set_scalar(sacb_handle h, const char* name, double value);

/* Replay ring in HBM as a struct of arrays: state[N,O] f32, actions[N,A] f32, rewards[N] f32,
 * next_state[N,O] f32, done[N] u8 (0/1).  Replaces ReplayBuffer.__init__ (buffer/replay_buffer.py:18-27). */
int sacb_bind_buffer(sacb_handle h, float* state_dev, float* actions_dev, float* rewards_dev,
                     float* next_state_dev, uint8_t* done_dev, int64_t capacity);
/* Append `count` transitions given as HOST rows (obs[count,O], act[count,A], rew[count], next_obs[count,O],
 * done[count] u8) at the ring pointer; ptr/size advance exactly as ReplayBuffer.store does
 * (buffer/replay_buffer.py:29-43).  Rows are staged through pinned memory and copied on `stream`. */
int sacb_store(sacb_handle h, const float* obs_host, const float* act_host, const float* rew_host,
               const float* next_obs_host, const uint8_t* done_host, int64_t count, uintptr_t stream);
/* Ring state (ReplayBuffer.ptr / .size / .max_size). */
/* The same ring as ONE packed array [capacity][pitch] of float32 records  [ state(O) | action(A) | reward | next_state(O) |
 * done (0/1) | pad ]  with a 16-byte-aligned pitch (the Python ReplayBuffer uses a multiple of 64 bytes).  A sampled row is
 * then one aligned, contiguous read -- fetched inside the update kernel by one TMA bulk copy (cp.async.bulk) per row into
 * shared memory -- instead of five unaligned ones, and sacb_store ships rows with one H2D copy per ring segment. */
int sacb_bind_buffer_packed(sacb_handle h, float* records_dev, int32_t pitch_floats, int64_t capacity);
int sacb_buffer_state(sacb_handle h, int64_t* ptr, int64_t* size, int64_t* capacity);
int sacb_set_buffer_state(sacb_handle h, int64_t ptr, int64_t size);

/* Uniform batch WITHOUT replacement + gather (ReplayBuffer.sample, buffer/replay_buffer.py:45-54).
 * idx_in_dev == NULL: indices are drawn on the device (cuRAND Philox, duplicate-free); otherwise the given
 * int64 indices [batch] are used.  Outputs: states[B,O], actions[B,A], rewards[B], next_states[B,O], done[B]
 * (all f32, done as 0/1), idx_out_dev (int64 [B], may be NULL).  SACB_ERR_SAMPLE if size < batch. */
int sacb_sample(sacb_handle h, int32_t batch, const int64_t* idx_in_dev, float* states_dev, float* actions_dev,
                float* rewards_dev, float* next_states_dev, float* done_dev, int64_t* idx_out_dev, uintptr_t stream);

/* n_updates full gradient updates = the burst sac/algorithm.py:274-281:
 *   sample -> update_critic -> update_policy -> update_targets, n_updates times, no host round trip.
 * idx_in_dev  : NULL (device draws) or int64 [n_updates, B]
 * eps_in_dev  : NULL (device draws N(0,1)) or f32 [n_updates, 2, B, A]  ([.,0] = critic-target pass
 *               sac/algorithm.py:61, [.,1] = actor-loss pass :37)
 * loss_q_dev / loss_pi_dev : f32 [n_updates] outputs (the values the reference reads back at :280-281). */
int sacb_update(sacb_handle h, int32_t n_updates, const int64_t* idx_in_dev, const float* eps_in_dev,
                float* loss_q_dev, float* loss_pi_dev, uintptr_t stream);

/* Piecewise API mirroring the reference methods on an explicit batch (all f32 device arrays of B rows). */
int sacb_load_batch(sacb_handle h, const float* states_dev, const float* actions_dev, const float* rewards_dev,
                    const float* next_states_dev, const float* done_dev, uintptr_t stream);
/* SAC.update_critic (sac/algorithm.py:115-141) on the loaded batch; eps_dev f32 [B,A] or NULL. */
int sacb_update_critic(sacb_handle h, const float* eps_dev, float* loss_q_dev, uintptr_t stream);
/* SAC.update_policy (sac/algorithm.py:143-162) on the loaded batch. */
int sacb_update_policy(sacb_handle h, const float* eps_dev, float* loss_pi_dev, uintptr_t stream);
/* update_targets (sac/algorithm.py:77-81) critic -> target with the configured polyak. */
int sacb_update_targets(sacb_handle h, uintptr_t stream);

/* Actor.forward (networks/linear.py:32-53) on n rows.  eps_dev NULL + deterministic=0 draws noise on device.
 * logp_dev may be NULL (with_logprob=False). */
int sacb_actor_forward(sacb_handle h, const float* x_dev, int32_t n, const float* eps_dev, int32_t deterministic,
                       float* pi_dev, float* logp_dev, uintptr_t stream);
/* DoubleCritic.forward (networks/linear.py:78-79); which = 0 online critic, 1 target critic. */
int sacb_critic_forward(sacb_handle h, int32_t which, const float* state_dev, const float* action_dev, int32_t n,
                        float* q1_dev, float* q2_dev, uintptr_t stream);

/* Data-parallel communicator (replaces sac/mpi.py).  nccl_lib_path = the libnccl.so.2 to dlopen (NULL: default
 * search).  sacb_comm_unique_id fills 128 bytes on rank 0; the caller broadcasts them out of band.
 * After sacb_comm_init, sacb_update all-reduce-averages the critic gradient (sac/algorithm.py:138) and the actor
 * gradient (intent of :155) across ranks; sacb_sync_params broadcasts rank 0's arenas (sac/algorithm.py:198-200). */
int sacb_comm_unique_id(const char* nccl_lib_path, uint8_t id_out[128]);
int sacb_comm_init(sacb_handle h, const char* nccl_lib_path, const uint8_t id[128], int32_t rank, int32_t world);
int sacb_sync_params(sacb_handle h, uintptr_t stream);

/* Peer mode: the gradient averaging of sac/mpi.py:77-85 fused INTO the update kernel.  Every rank maps every other rank's
 * engine workspace (CUDA IPC, same node, NVLink peers); in the Adam phases one thread per peer publishes "my gradient
 * of update t is complete" (system-scope release store into the peer's flag word), every CTA waits for all ranks' flags
 * (system-scope acquire) and then sums the ranks' gradient arenas over NVLink in fixed rank order while it applies
 * Adam.  No NCCL call and no extra kernel launch per update: a data-parallel burst stays one persistent launch and the
 * replicas stay bit-identical.  sacb_peer_export: 64-byte cudaIpcMemHandle of this rank's workspace (+ its length, which
 * must agree on all ranks).  sacb_peer_bind: the handles of all ranks, [world][64] bytes in rank order (world <= 8).
 * sacb_comm_init may still be used next to it for sacb_sync_params. */
int sacb_peer_export(sacb_handle h, uint8_t handle_out[64], int64_t* ws_floats);
int sacb_peer_bind(sacb_handle h, const uint8_t* handles, int32_t rank, int32_t world);
/* Health of the peer exchange: a waiter that sees no progress for SACB_PEER_TIMEOUT_MS (environment, default 20000)
 * raises an abort word and every waiter of the launch stops waiting, so a dead rank cannot hang its peers forever
 * (the reference's blocking comm.recv, sac/algorithm.py:264-265, would).  Synchronises `stream`, then returns 0 or
 * SACB_ERR_PEER (the message names the missing rank); after a timeout sacb_update refuses further bursts. */
int sacb_peer_status(sacb_handle h, uintptr_t stream);

/* Test / debug: copy an internal f32 array to HOST memory (synchronous).  Names: "backup","q1","q2","q1_target",
 * "q2_target","a2","logp_next","pi","logp_pi","q1_pi","q2_pi","grad_actor","grad_critic","loss_q_parts", ...;
 * "operand_path" -> [1 if precision 3 feeds its GEMM tiles from bf16 shadow operands through the TMA, number of tensor maps].
 * Returns the number of floats written via *n_out. */
int sacb_debug_read(sacb_handle h, const char* name, float* dst_host, int64_t max_floats, int64_t* n_out);
/* Test / debug: run n_updates like sacb_update with device-drawn inputs while every CTA stamps %globaltimer after the
 * work and after the grid barrier of every phase.  stamps_host receives [n_updates][n_phases][n_ctas][2] uint64 ns;
 * n_phases / n_ctas are returned.  Synchronous. */
int sacb_debug_profile(sacb_handle h, int32_t n_updates, uint64_t* stamps_host, int64_t max_words, int32_t* n_phases,
                       int32_t* n_ctas, uintptr_t stream);
/* Number of kernel launches the engine has enqueued so far (bench `gpu_launches`). */
int64_t sacb_launch_count(sacb_handle h);
/* GEMM arithmetic.  1 (default) = mma.sync tensor cores with 3xTF32 error compensation (fp32-level accuracy: every
 * product is a_lo*b_hi + a_hi*b_lo + a_hi*b_hi with fp32 accumulation); 0 = plain fp32 FFMA; 2 = the same 3xTF32
 * arithmetic on tcgen05 with TMEM accumulators; 3 = tcgen05 with bf16 operands (rounded to nearest-even from the fp32
 * tensors on load), fp32 accumulation in TMEM, fp32 master weights / activations / Adam -- the "bf16 tensor-core MLP
 * path" of the Humanoid-shaped configuration.  Nothing the reference exposes corresponds to this call: its arithmetic is
 * whatever torch CPU does (networks/linear.py:33-38). */
int sacb_set_precision(sacb_handle h, int32_t tf32x3);
/* Execution mode: 0 = persistent cooperative kernel (default), 1 = one launch per phase (debug / profiling). */
int sacb_set_mode(sacb_handle h, int32_t mode);

/* =====================================================================================================================
 * Visual SAC (BASELINE configs[4]): the reference's networks/convolutional.py + buffer/visual_replay_buffer.py path.
 *
 * Same conventions as above (status codes, caller-owned device pointers, stream as uintptr_t, no CPU fallback); messages
 * of failing sacv_* calls are read with sacv_last_error().  A handle owns one (VisualActor, VisualDoubleCritic, target)
 * triple's workspace for a fixed batch size.  Frames live in the replay ring as uint8 [N, C, H, W] (the reference keeps
 * float tensors of raw 0..255 pixels, environments/wall_runner.py:54-58; integral pixel values survive the uint8 ring
 * unchanged) and are gathered straight into im2col form; explicit batches (sacv_load_batch, forward calls) pass float32
 * frames like the reference's VisualBatch / MultiObservation.
 * ===================================================================================================================== */
typedef struct sacv_config {
    uint32_t struct_size;          /* = sizeof(sacv_config) */
    int32_t device;
    int32_t obs_dim, act_dim;      /* feature width O (environments/wall_runner.py:21: 168), action width A (56) */
    int32_t hidden1, hidden2;      /* MLP widths (main.py:61) */
    int32_t batch_size;
    int32_t vis_c, vis_h, vis_w;   /* frame shape (main.py:68 uses 3x64x64; BASELINE configs[4] 3x84x84) */
    int32_t filters[3], kernels[3], strides[3];   /* simple_cnn, networks/convolutional.py:32-44 (main.py:65-67: 32/64/64, 8/4/3, 4/2/1) */
    int32_t dense;                 /* simple_cnn dense_size (512) */
    float log_min_std, log_max_std, act_limit;
    int32_t precision;             /* GEMM arithmetic: 1 = 3xTF32 (fp32-level accuracy, default), 0 = plain TF32 */
    double alpha, gamma, polyak, reward_scale, lr_pi, lr_q, beta1, beta2, adam_eps;
    uint32_t flags;                /* SACB_FLAG_CANONICAL_LOGPROB */
    uint32_t reserved0;
    uint64_t seed;
} sacv_config;

/* Offsets (floats) inside the flat arenas.  Actor arena: trunk, CNN block, heads ([mu ; log_std] as one [2A, H2+1] matrix,
 * convolutional.py:78-79).  Critic arena: q1 net then q2 net, `critic_net_floats` apart, each: MLP [O+A,H1,H2,1], final
 * Linear(2,1), CNN block (convolutional.py:140-143).  cnn_* are relative to the start of a CNN block. */
typedef struct sacv_layout_info {
    int64_t actor_floats, a_l0w, a_l0b, a_l1w, a_l1b, a_cnn, a_muw, a_lsw, a_mub, a_lsb;
    int64_t critic_floats, critic_net_floats, c_l0w, c_l0b, c_l1w, c_l1b, c_l2w, c_l2b, c_fw, c_fb, c_cnn;
    int64_t cnn_floats, cnn_w[3], cnn_b[3], cnn_lw, cnn_lb, cnn_fw, cnn_fb;
    int32_t conv_h[3], conv_w[3], flat, reserved0;      /* calculate_size, convolutional.py:14-29 */
} sacv_layout_info;

typedef struct sacv_engine* sacv_handle;

const char* sacv_last_error(void);
int sacv_layout(const sacv_config* cfg, sacv_layout_info* out);
int sacv_create(const sacv_config* cfg, sacv_handle* out);
int sacv_destroy(sacv_handle h);
int sacv_bind_params(sacv_handle h, float* actor_dev, float* critic_dev, float* target_dev);
int sacv_bind_adam(sacv_handle h, float* m_actor, float* v_actor, int64_t step_actor, float* m_critic, float* v_critic, int64_t step_critic);
int sacv_get_adam_steps(sacv_handle h, int64_t* step_actor, int64_t* step_critic);
int sacv_set_scalar(sacv_handle h, const char* name, double value);      /* alpha gamma polyak reward_scale lr_pi lr_q beta1 beta2 adam_eps precision */

/* VisualReplayBuffer.__init__ / store / sample (buffer/visual_replay_buffer.py:21-66).  Ring arrays (device): features
 * [N,O] f32, frames [N,C,H,W] u8, actions [N,A] f32, rewards [N] f32, next_features, next_frames, done [N] u8. */
int sacv_bind_buffer(sacv_handle h, float* features, uint8_t* frames, float* actions, float* rewards, float* next_features,
                     uint8_t* next_frames, uint8_t* done, int64_t capacity);
int sacv_buffer_state(sacv_handle h, int64_t* ptr, int64_t* size, int64_t* capacity);
int sacv_set_buffer_state(sacv_handle h, int64_t ptr, int64_t size);
int sacv_store(sacv_handle h, const float* features_host, const uint8_t* frames_host, const float* actions_host, const float* rewards_host,
               const float* next_features_host, const uint8_t* next_frames_host, const uint8_t* done_host, int64_t count, uintptr_t stream);
int sacv_sample(sacv_handle h, int32_t batch, const int64_t* idx_dev_or_null, float* features, float* frames_f32, float* actions, float* rewards,
                float* next_features, float* next_frames_f32, float* done, int64_t* idx_out_or_null, uintptr_t stream);

/* n_updates x (VisualReplayBuffer.sample -> SAC.update_critic -> SAC.update_policy -> update_targets), sac/algorithm.py:274-278;
 * idx [n,B] / eps [n,2,B,A] may be injected (device) or null (drawn on the device). */
int sacv_update(sacv_handle h, int32_t n_updates, const int64_t* idx_dev_or_null, const float* eps_dev_or_null, float* loss_q_dev,
                float* loss_pi_dev, uintptr_t stream);
/* piecewise API on an explicit VisualBatch (sac/algorithm.py:115-162,77-81) */
int sacv_load_batch(sacv_handle h, const float* features, const float* frames_f32, const float* actions, const float* rewards,
                    const float* next_features, const float* next_frames_f32, const float* done, uintptr_t stream);
int sacv_update_critic(sacv_handle h, const float* eps_dev_or_null, float* loss_q_dev, uintptr_t stream);
int sacv_update_policy(sacv_handle h, const float* eps_dev_or_null, float* loss_pi_dev, uintptr_t stream);
int sacv_update_targets(sacv_handle h, uintptr_t stream);
/* VisualActor.forward / VisualDoubleCritic.forward (networks/convolutional.py:85-121,144-183) on n <= batch_size rows */
int sacv_actor_forward(sacv_handle h, const float* features, const float* frames_f32, int32_t n, const float* eps_dev_or_null,
                       int32_t deterministic, float* pi, float* logp_or_null, uintptr_t stream);
int sacv_critic_forward(sacv_handle h, int32_t which, const float* features, const float* frames_f32, const float* actions, int32_t n,
                        float* q1, float* q2, uintptr_t stream);
int sacv_debug_read(sacv_handle h, const char* name, float* dst_host, int64_t max_floats, int64_t* n_out);
int64_t sacv_launch_count(sacv_handle h);

#ifdef __cplusplus
}
#endif
#endif /* SACB200_H */
