/* marl_b200.h — C-ABI of libmarl_b200.so
 *
 * Drop-in boundary of the B200-native DQN-MARL hot path.  The reference
 * (LX-530/DQN-MARL) is pure Python and has no FFI of its own: the boundary its
 * runners use is the Python class surface
 *     EvacuationEnv.reset()/step()           Louvre_Evacuation/envs/evacuation_env.py:61,122
 *     EvacuationEnvMulti.reset()/step()      Louvre_Evacuation/envs/evacuation_env_multi.py:31,55
 *     DQNAgent.act()/remember()/learn()      Louvre_Evacuation/agents/dqn_agent.py:97,101,126
 *     DQNAgent.update_target_network()       Louvre_Evacuation/agents/dqn_agent.py:170
 * The Python mirror of those classes lives in dqn_marl_b200/{envs,agents}; every
 * entry point below is what that mirror binds (ctypes) and cites the reference
 * code it replaces.  INTEGRATION.md shows the binding stub.
 *
 * Conventions
 *   - plain C types only; every pointer marked `dev` is a CUDA device pointer
 *     owned by the CALLER (PyTorch tensors), `host` pointers are host memory;
 *   - all GPU work is enqueued on the caller's stream (`stream` is a
 *     cudaStream_t passed as void*), the library never synchronises on its own
 *     unless the function is documented as host-returning;
 *   - return 0 on success, negative mq_status on failure; mq_last_error() gives
 *     a thread-local message;
 *   - no CPU fallback: every compute entry point fails with MQ_ERR_CUDA when no
 *     device is usable.
 */
#ifndef MARL_B200_H
#define MARL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MQ_ABI_VERSION 3      /* 2: mq_env_create_layouts, mq_layout_tables_device, mq_replay_restore; 3: mq_qnet_explore_draw */
#define MQ_MAX_ROBOTS 4
#define MQ_OBS_WIN 11                          /* evacuation_env.py:56  state_size = (11, 11, 6) */
#define MQ_OBS_CH 6
#define MQ_OBS_SIZE (MQ_OBS_WIN * MQ_OBS_WIN * MQ_OBS_CH)   /* 726 */
#define MQ_N_ACTIONS 5                         /* evacuation_env.py:57 */
#define MQ_OBS_WIRE_WORDS 136                  /* compact wire form of one observation window: 544 B (mq_env_set_obs_wire) */

typedef enum mq_status {
    MQ_OK = 0,
    MQ_ERR_ARG = -1,
    MQ_ERR_CUDA = -2,
    MQ_ERR_ALLOC = -3,
    MQ_ERR_UNSUPPORTED = -4
} mq_status;

const char* mq_last_error(void);
int mq_abi_version(void);

/* ------------------------------------------------------------------------
 * Static floor field (host, init time).  Replaces Map.Init_Potential
 * (map.py:127-148): 8-connected Dijkstra from the exits, step cost 1.0 / 1.4,
 * start value 1, then += add_term (200*danger^2 at fire step 0) on reached cells.
 *   wall      host u8  [(L+2)*(W+2)]  1 = inf before the search (map.py:44-57,66-71)
 *   exits     host i32 [n_exits][2]
 *   add_term  host f64 [(L+2)*(W+2)]
 *   space_out host f64 [(L+2)*(W+2)]  inf where unreached
 * ---------------------------------------------------------------------- */
int mq_floor_field(int32_t L, int32_t W, const uint8_t* wall, const int32_t* exits, int32_t n_exits,
                   const double* add_term, double* space_out);

/* The same field for a batch of layouts, on the device (device pointers; init-time: synchronises `stream` while it
 * waits for the relaxation sweeps to reach their fixed point).  Bit-identical to mq_floor_field per layout.
 *   wall      dev u8  [n_layouts][(L+2)*(W+2)]
 *   exits     dev i32 [n_layouts][max_exits][2],  n_exits dev i32 [n_layouts]
 *   add_term  dev f64 [n_layouts][(L+2)*(W+2)] or NULL
 *   space_out dev f64 [n_layouts][(L+2)*(W+2)];  sweeps_out (host, optional) = relaxation sweeps that were run */
int mq_floor_field_device(int32_t L, int32_t W, int32_t n_layouts, const uint8_t* wall, const int32_t* exits, int32_t max_exits,
                          const int32_t* n_exits, const double* add_term, double* space_out, int32_t* sweeps_out, void* stream);

/* dp5 / cellinfo (see mq_layout below) of a batch of layouts from their floor fields, on the device: what lets the output of
 * mq_floor_field_device feed mq_env_create_layouts without a round trip through the host.
 *   space     dev f64 [n_layouts][(L+2)*(W+2)]   (mq_floor_field_device)
 *   barrier   dev u8  [n_layouts][(L+2)*(W+2)]   Map.barrier_list membership (map.py:43-57,72-73): observation channel 3
 *   exits / n_exits as above;  obs_exit dev i32 [n_layouts][2]  (evacuation_env.py:113 exit_location)
 *   dp5_out   dev f64 [n_layouts][(L+2)*(W+2)][8];  cellinfo_out dev u8 [n_layouts][(L+2)*(W+2)] */
int mq_layout_tables_device(int32_t L, int32_t W, int32_t n_layouts, const double* space, const uint8_t* barrier, const int32_t* exits,
                            int32_t max_exits, const int32_t* n_exits, const int32_t* obs_exit, double* dp5_out, uint8_t* cellinfo_out,
                            void* stream);

/* ------------------------------------------------------------------------
 * Layout tables (host pointers; copied to the device by mq_env_create).
 * Cell index = x*(W+2)+y, the reference indexes space[x][y] (map.py:44).
 * ---------------------------------------------------------------------- */
typedef struct mq_layout {
    int32_t L, W;                       /* Map.Length / Map.Width (map.py:39-40) */
    int32_t n_fire_steps;               /* FireSpreadModel max_steps + 1 = 181 (fire_model.py:213) */
    int32_t ctr_box[4];                 /* x0, y0, w, h of danger_ctr */
    int32_t int_box[4];                 /* x0, y0, w, h of danger_int (x0/y0 may be negative) */
    int32_t robot_range[2];             /* map.py:75 */
    int32_t robot_start[MQ_MAX_ROBOTS][2];   /* map.py:76 / evacuation_env_multi.py:27 */
    int32_t reset_obs_center[2];        /* evacuation_env.py:64 */
    int32_t obs_exit[2];                /* evacuation_env.py:194 exit_location used by the reward */
    const double*  dp5;                 /* [G][8]  (space[c]-space[n])*5.0, -inf if n invalid (people.py:270,288) */
    const uint8_t* cellinfo;            /* [G] bit0 Check_Valid(map.py:85) bit1 obs ch3 bit2 obs ch4 bit3 checkSavefy(map.py:93) */
    const double*  danger_ctr;          /* [n_fire_steps][w][h]  get_max_danger at cell centres (people.py:205) */
    const double*  danger_int;          /* [n_fire_steps][w][h]  get_max_danger at integer coords (evacuation_env.py:106) */
} mq_layout;

/* ------------------------------------------------------------------------
 * Batched environment.  One handle = n_envs independent instances of
 * EvacuationEnv (n_robots = 1) / EvacuationEnvMulti (n_robots = 2) on one GPU.
 * ---------------------------------------------------------------------- */
typedef struct mq_env_cfg {
    int32_t  n_envs;
    int32_t  n_people;                  /* evacuation_env.py:24 */
    int32_t  n_robots;                  /* 1..MQ_MAX_ROBOTS */
    int32_t  device;                    /* CUDA ordinal */
    uint64_t seed;                      /* Philox key */
    int32_t  env_id_base;               /* global id of env 0 (rank offset) -> Philox counter word 0 */
    int32_t  max_steps;                 /* evacuation_env.py:27-30: 600 s / 0.5 s = 1200 */
    int32_t  reset_robots;              /* 0: keep robot cells across reset, obs from reset_obs_center (quirk Q7,
                                              EvacuationEnv); 1: robots return to robot_start (EvacuationEnvMulti) */
    int32_t  reset_fire;                /* 0: fire step survives reset (quirk Q6, reference); 1: restart at 0 */
    int32_t  auto_reset;                /* 1: an env that reports done is re-spawned inside the same step call and
                                              obs_out holds the first observation of the new episode */
    int32_t  reserved;
    double   evac_reward;               /* evacuation_env.py:16-19 class attributes */
    double   death_penalty;
    double   death_acc_penalty;
    double   alive_bonus;
} mq_env_cfg;

/* Device-resident state, allocated by the caller, SoA across envs.
 * n_pad = n_people rounded up to 16; rmap_words = (L+2) * ceil((W+2)/32). */
typedef struct mq_env_state {
    uint32_t* pos;        /* dev [n_envs][n_pad]  cell x | y<<16  (Person.pos is always a cell centre, people.py:303) */
    double*   health;     /* dev [n_envs][n_pad]  Person.health (people.py:19) */
    double*   acc;        /* dev [n_envs][n_pad]  Person.move_accumulator (people.py:22) */
    uint8_t*  flags;      /* dev [n_envs][n_pad]  bit0 savety, bit1 dead (people.py:18,20) */
    uint32_t* rmap;       /* dev [n_envs][rmap_words]  People.rmap as 1 bit/cell, row x, bit y (people.py:162) */
    int32_t*  robots;     /* dev [n_envs][MQ_MAX_ROBOTS][2]  Map.robot_positions (map.py:78) */
    int32_t*  scalars;    /* dev [n_envs][MQ_ENV_SCALARS]  see enum below */
} mq_env_state;

enum {
    MQ_S_FIRE_STEP = 0,   /* progressive_model.current_step (fire_model.py:63-67), never reset in the reference */
    MQ_S_CUR_STEP = 1,    /* EvacuationEnv.current_step (evacuation_env.py:71,148) */
    MQ_S_PREV_EVAC = 2,   /* evacuation_env.py:72,285 */
    MQ_S_PREV_DEAD = 3,   /* evacuation_env.py:73,286 */
    MQ_S_EPISODE = 4,     /* number of resets so far -> spawn draw key */
    MQ_S_TICK = 5,        /* number of steps since creation -> step draw key */
    MQ_S_EVAC = 6,        /* current evacuated count (info / get_performance_metrics) */
    MQ_S_DEAD = 7,        /* current dead count */
    MQ_S_ROBOT_POS_X = 8, /* Map.robot_position (map.py:76,200-201): aliases robots[0] except between a reset() */
    MQ_S_ROBOT_POS_Y = 9, /*   and the next valid move_robot call (evacuation_env.py:64, quirk Q7) */
    MQ_ENV_SCALARS = 16   /* 64 B per env; 10..15 reserved */
};

typedef struct mq_env mq_env;

int mq_env_state_sizes(const mq_env_cfg* cfg, const mq_layout* layout, int64_t* n_pad, int64_t* rmap_words);
int mq_env_create(mq_env** out, const mq_env_cfg* cfg, const mq_layout* layout, const mq_env_state* state);
/* Several layouts in one batch (per-env layouts; the reference builds one Map per env instance,
 * evacuation_env.py:42-43).  All layouts share L x W and n_fire_steps; walls, exits, floor fields, fire sources, robot starts
 * may differ.  env_layout host i32 [n_envs]: the layout index of every env (NULL when n_layouts == 1).
 * tables_on_device != 0: dp5 / cellinfo / danger_ctr / danger_int of every mq_layout are DEVICE pointers (e.g. written by
 * mq_layout_tables_device); they are copied, the caller keeps ownership.  Danger tables given by the same pointer are stored once. */
int mq_env_create_layouts(mq_env** out, const mq_env_cfg* cfg, const mq_layout* layouts, int32_t n_layouts,
                          const int32_t* env_layout, int32_t tables_on_device, const mq_env_state* state);
int mq_env_destroy(mq_env* env);
/* EvacuationEnv.EVAC_REWARD etc. are mutated at runtime by overnight_experiments.py:69-70 */
int mq_env_set_reward_coefs(mq_env* env, double evac_reward, double death_penalty, double death_acc_penalty,
                            double alive_bonus);

/* EvacuationEnv.reset (evacuation_env.py:61-82) + People.__init__ spawn (people.py:185-194).
 *   env_mask      dev u8 [n_envs] or NULL (= all): which envs to reset
 *   inject_spawn  dev i16 [n_envs][n_people][2] or NULL: cells to use instead of keyed draws (test mode)
 *   obs_out       dev f32 [n_envs][n_robots][11][11][6] or NULL
 *   obs64_out     dev f64 same shape or NULL (single-env facade returns float64 like the reference) */
int mq_env_reset(mq_env* env, const uint8_t* env_mask, const int16_t* inject_spawn, float* obs_out,
                 double* obs64_out, void* stream);

/* EvacuationEnv.step (evacuation_env.py:122-172): move_robot (map.py:160-202), People.run
 * (people.py:196-253), fire update, _calculate_reward (:174-288), done (:155-157), _get_state (:84-120).
 *   actions     dev i32 [n_envs][n_robots]; values outside 0..4 are ignored (map.py:180-181)
 *   reward_out  dev f64 [n_envs]
 *   done_out    dev u8  [n_envs] */
int mq_env_step(mq_env* env, const int32_t* actions, float* obs_out, double* obs64_out, double* reward_out,
                uint8_t* done_out, void* stream);

/* Compact wire form of the observations for callers that take them to the HOST every step (the reference's agent lives on the
 * host: dqn_agent.py:101-110).  Of the 6 channels of _get_state (evacuation_env.py:84-120) channel 0 is identically 0,
 * channels 1 / 3 / 4 hold 0 or 1 and channel 5 only marks the centre cell, so a window travels as MQ_OBS_WIRE_WORDS u32:
 *   words 0..120    channel 2 (danger) of cell c = i*11 + j as f32 bits
 *   words 121..124  channel 1 (People.rmap) bit plane, bit c        125..128  channel 3      129..132  channel 4      133..135  0
 * 544 B instead of 2904 B per window over PCIe.  wire_out dev u32 [n_envs][n_robots][MQ_OBS_WIRE_WORDS], written by every
 * following mq_env_step / mq_env_reset (in addition to obs_out / obs64_out, which may then be NULL); NULL switches it off. */
int mq_env_set_obs_wire(mq_env* env, uint32_t* wire_out);
/* Host routine (no GPU work): expand n_windows wire records (host memory) into dense f32 windows [n_windows][11][11][6],
 * bit-identical to what obs_out would have held.  n_threads host threads (<= 0: one per 2048 windows, at most 16). */
int mq_obs_wire_expand(const uint32_t* wire, int64_t n_windows, float* obs_out, int32_t n_threads);

/* People.rmap as bytes: dev u8 [n_envs][(L+2)*(W+2)] */
int mq_env_unpack_rmap(mq_env* env, uint8_t* rmap_out, void* stream);
/* number of kernels this handle has launched so far (bench.py gpu_launches) */
int64_t mq_env_launch_count(const mq_env* env);

/* ------------------------------------------------------------------------
 * Replay ring (device resident).  Replaces DQNAgent.memory = deque(maxlen)
 * (dqn_agent.py:88-89), remember() (:97-99) and random.sample + stacking in
 * learn() (:132-140).  Storage is caller-owned fp32 SoA.
 * ---------------------------------------------------------------------- */
typedef struct mq_replay_store {
    float*   state;       /* dev [capacity][726] */
    float*   next_state;  /* dev [capacity][726] */
    int32_t* action;      /* dev [capacity] */
    float*   reward;      /* dev [capacity] */
    uint8_t* done;        /* dev [capacity] */
} mq_replay_store;

typedef struct mq_replay mq_replay;

int mq_replay_create(mq_replay** out, int64_t capacity, int32_t device, const mq_replay_store* store);
int mq_replay_destroy(mq_replay* rb);
int64_t mq_replay_size(const mq_replay* rb);      /* len(agent.memory) */
int64_t mq_replay_cursor(const mq_replay* rb);
int64_t mq_replay_launch_count(const mq_replay* rb);
/* Checkpoint resume (the reference saves only the networks, dqn_agent.py:174-182; a deque has no other state than its
 * contents): after the caller has copied the saved transitions back into the storage tensors, restore len(memory) and the
 * write position.  0 <= size <= capacity, 0 <= cursor < capacity, cursor == size % capacity unless the ring is full. */
int mq_replay_restore(mq_replay* rb, int64_t size, int64_t cursor);
/* n transitions appended FIFO (oldest overwritten once full).  reward is the env's f64 reward,
 * cast to f32 exactly as torch.FloatTensor(rewards) does at dqn_agent.py:138. */
int mq_replay_push(mq_replay* rb, const float* state, const int32_t* action, const double* reward,
                   const float* next_state, const uint8_t* done, int64_t n, void* stream);
/* B transitions sampled uniformly WITHOUT replacement from the current contents.
 *   inject_idx  dev i64 [B] logical indices (0 = oldest) or NULL -> keyed permutation (seed, draw_id) */
int mq_replay_sample(mq_replay* rb, int64_t B, uint64_t seed, uint64_t draw_id, const int64_t* inject_idx,
                     float* state, int64_t* action, float* reward, float* next_state, uint8_t* done,
                     int64_t* idx_out, void* stream);

/* ------------------------------------------------------------------------
 * Q-network (DQNNetwork, dqn_agent.py:15-61) and learner (DQNAgent.act/learn, :101-172).
 * Parameters live in 12 caller-owned fp32 tensors in state_dict order (conv1.weight conv1.bias conv2.weight
 * conv2.bias conv3.weight conv3.bias fc1.weight fc1.bias fc2.weight fc2.bias fc3.weight fc3.bias) in the library's
 * kernel layouts: convK.weight as [(kh*3+kw)*Cin + c][Cout] (PyTorch [Cout][Cin][kh][kw]); fc1.weight as
 * [512][p*128 + c] with p = i*11+j (PyTorch [512][c*121 + p]); everything else as in PyTorch.  The Python mirror
 * permutes at the state_dict / checkpoint boundary (dqn_marl_b200/agents/qnet_params.py).
 * ---------------------------------------------------------------------- */
#define MQ_QNET_TENSORS 12
#define MQ_QNET_PARAMS 8157093

typedef struct mq_qnet_bind {
    float* online[MQ_QNET_TENSORS];   /* dev */
    float* target[MQ_QNET_TENSORS];   /* dev */
    float* grad[MQ_QNET_TENSORS];     /* dev */
    float* adam_m[MQ_QNET_TENSORS];   /* dev */
    float* adam_v[MQ_QNET_TENSORS];   /* dev */
} mq_qnet_bind;

typedef struct mq_hparams {
    float gamma;          /* dqn_agent.py:73 */
    float lr;             /* :77 */
    float beta1, beta2;   /* torch.optim.Adam defaults (:85) */
    float adam_eps;
    float clip_norm;      /* :158 */
    int32_t huber;        /* 0 = MSE (:151, reference), 1 = Huber(delta=1) option of north_star */
    int32_t adam_step;    /* t (1-based) of this update */
} mq_hparams;

typedef struct mq_qnet mq_qnet;

int mq_qnet_create(mq_qnet** out, int32_t device, int64_t max_batch, const mq_qnet_bind* bind);
int mq_qnet_destroy(mq_qnet* net);
/* which: 0 online, 1 target.  obs dev f32 [B][11][11][6] (NHWC as the env writes it; the permute of
 * dqn_agent.py:37-45 is folded into the conv1 loader).  drop_mask dev u8 [B][512] or NULL (= eval mode). */
int mq_qnet_forward(mq_qnet* net, int32_t which, const float* obs, int64_t B, const uint8_t* drop_mask,
                    float* q_out, void* stream);
/* epsilon-greedy (dqn_agent.py:101-124) fused on top of the online forward: action = keyed random action if
 * u <= eps else first argmax.  Sample b belongs to env env_id_base + b / n_robots, robot b % n_robots.
 * eps = 0 is `training=False`.  q_out may be NULL. */
int mq_qnet_act(mq_qnet* net, const float* obs, int64_t B, float eps, uint64_t seed, uint32_t env_id_base,
                uint32_t tick, int32_t n_robots, const uint8_t* drop_mask, int32_t* action_out, float* q_out,
                void* stream);
/* HOST function, no device work: the keyed exploration draw mq_qnet_act makes for sample (env, robot) at `tick` —
 * returns 1 and sets *action_out to the random action when u <= eps (`np.random.random() <= self.epsilon` ->
 * `random.randrange(action_size)`, dqn_agent.py:103-104), else 0.  The single-env drop-in agent asks this first and,
 * like the reference, skips the network forward when it explores; the draw is the one the kernel would make. */
int mq_qnet_explore_draw(float eps, uint64_t seed, uint32_t env, uint32_t tick, uint32_t robot, int32_t* action_out);
/* The differentiable half of DQNAgent.learn() (dqn_agent.py:143-155) on a sampled batch: target forward, online
 * forward, TD target, loss, full backward into the bound gradient tensors.  drop_online / drop_target: u8 [B][512]
 * keep-masks of the two forward passes or NULL (= .eval()).  loss_out dev f32 [1].  The caller may all-reduce the
 * gradients between this call and mq_qnet_clip_adam. */
int mq_qnet_td_backward(mq_qnet* net, const float* state, const int64_t* action, const float* reward,
                        const float* next_state, const uint8_t* done, int64_t B, const mq_hparams* hp,
                        const uint8_t* drop_online, const uint8_t* drop_target, float* loss_out, void* stream);
/* mq_qnet_td_backward in two calls (same arguments) so that a data-parallel caller overlaps the gradient exchange with the
 * backward: part 1 = both forwards, loss, backward of fc3 / fc2 / fc1 — gradient tensors 6..11 (99 % of the 8,157,093 floats)
 * are final when its work completes; part 2 = backward of the three convolutions (tensors 0..5). */
int mq_qnet_td_backward_part(mq_qnet* net, const float* state, const int64_t* action, const float* reward,
                             const float* next_state, const uint8_t* done, int64_t B, const mq_hparams* hp,
                             const uint8_t* drop_online, const uint8_t* drop_target, float* loss_out, int32_t part, void* stream);
/* Backward of the online network from an external dL/dQ (dev f32 [B][5]) — the autograd path a runner takes when it builds
 * its own loss on agent.q_network(states) (train_qmix.py:92-110: two agents' Q-values go through a mixing network before
 * the loss).  Recomputes the online forward on `state` with the same drop mask, then fills the bound gradient tensors. */
int mq_qnet_backward(mq_qnet* net, const float* state, const float* dq, int64_t B, const uint8_t* drop_online, void* stream);
/* clip_grad_norm_(params, clip_norm) + optimizer.step() (dqn_agent.py:158-160).  Gradients are multiplied by
 * grad_scale first (1/world_size after a summing all-reduce).  gnorm_out dev f32 [1] or NULL. */
int mq_qnet_clip_adam(mq_qnet* net, const mq_hparams* hp, float grad_scale, float* gnorm_out, void* stream);
/* update_target_network (dqn_agent.py:170-172): tau = 1 hard copy (reference); tau < 1 Polyak option */
int mq_qnet_sync_target(mq_qnet* net, float tau, void* stream);
/* nn.Dropout(0.2) keep-mask from keyed draws: mask dev u8 [n], keep probability 1-p (dqn_agent.py:33,57) */
int mq_qnet_dropout_mask(uint8_t* mask, int64_t n, float p, uint64_t seed, uint64_t counter, void* stream);
/* precision 0 = fp32 FFMA parity path (default, Q/loss within 1e-5 of the reference's fp32 agent);
 * precision 1 = bf16 tcgen05/TMEM path for conv2, conv3 and fc1 forward + backward (fp32 accumulation, fp32 master
 * weights and optimizer; reported separately).  Allocates the bf16 workspaces on first use. */
int mq_qnet_set_precision(mq_qnet* net, int32_t precision);
/* tell the library that the caller overwrote the bound parameter tensors (load_state_dict / checkpoint load) */
int mq_qnet_params_changed(mq_qnet* net);
int64_t mq_qnet_launch_count(const mq_qnet* net);

/* ------------------------------------------------------------------------
 * Stand-alone bf16 tensor-core GEMM (tcgen05 + TMEM + TMA), the building block of the Q-network's throughput
 * path: C[M][N] = A[M][K] bf16 * B[N][K]^T bf16 (both K-contiguous), written as f32 (C) and / or bf16 (C_bf16).
 * bn = tile selector: 32 / 64 / 128 / 256 = 128 x bn outputs per CTA; 384 = 256 x 128, 512 = 256 x 256 (two row tiles per CTA).
 * ---------------------------------------------------------------------- */
int mq_gemm_bf16(const void* A, const void* B, float* C, void* C_bf16, int32_t M, int32_t N, int32_t K, int32_t bn,
                 int32_t splits, float* workspace, void* stream);

/* C[M][N] f32 = At[K][M]^T * Bt[K][N]: both bf16 operands stored with K as the row index (MN-major UMMA descriptors).
 * The weight-gradient shapes of DQNAgent.learn (dqn_agent.py:153, loss.backward) reduce over the batch rows; this form
 * needs no transposed copy.  M and N multiples of 8. */
int mq_gemm_bf16_tn(const void* At, const void* Bt, float* C, int32_t M, int32_t N, int32_t K, int32_t splits, float* workspace,
                    void* stream);

/* 3x3 / padding 1 convolution over the 11x11 observation window as an implicit GEMM (nn.Conv2d(.., 3, padding=1) of
 * DQNNetwork, dqn_agent.py:22-24,48-50): Y[batch*121][Cout] f32 = im2col(X) * Wk^T with X [batch][11][11][Cin] bf16 (NHWC)
 * and Wk [Cout][9*Cin] bf16 (taps in (kh, kw, c) order).  The im2col matrix is never written: every tap is one shifted,
 * zero-filled 4-D TMA box.  flip = 1 mirrors the taps (the data-gradient convolution of loss.backward()).  Output in fp32
 * (Y) and / or bf16 (Y_bf16).  bn = tile width 128 / 64 / 32, or 0 = the persistent kernel (weights resident in shared memory,
 * double-buffered TMEM accumulators, TMA-store epilogue for a bf16-only output). */
int mq_conv3x3_bf16(const void* X, const void* Wk, float* Y, void* Y_bf16, int64_t batch, int32_t Cin, int32_t Cout, int32_t flip,
                    int32_t bn, void* stream);

/* Weight gradient of that convolution: dW[9*Cin][Cout] f32 = im2col(X)^T * dY, dY [batch*121][Cout] bf16.  splits > 1
 * partitions the samples (workspace >= splits * 9*Cin*Cout floats).  dBias (optional, [Cout]) = column sums of dY, computed
 * by the same kernel through a spare operand row of ones (needs 9*Cin % 128 != 0 and splits*Cout more workspace floats). */
int mq_conv3x3_wgrad_bf16(const void* X, const void* dY, float* dW, float* dBias, int64_t batch, int32_t Cin, int32_t Cout,
                          int32_t splits, float* workspace, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MARL_B200_H */
