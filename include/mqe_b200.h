/*
 * mqe_b200.h -- C ABI of libmqe_b200.so: the B200-native replacement for the Isaac Gym tensor-API
 * subset that mqe.envs.go1.Go1.step() drives (reference: mqe/envs/go1/go1.py:35-62,
 * mqe/envs/base/legged_robot.py:117-157, 394-470, 549-595).
 *
 * The reference has no FFI of its own: its "lower" interface is the closed isaacgym Python module.
 * Each entry point below names the gym call(s) it replaces.  All functions return 0 on success or a
 * negative MqeStatus; mqe_last_error() returns a thread-local message.  No function throws, none
 * synchronises the device unless it says so.  Pointers named d_* are DEVICE pointers, h_* HOST pointers.
 * Every kernel is enqueued on the stream given at creation (or set with mqe_sim_set_stream).
 *
 * Layout conventions: env-major, fp32, quaternions xyzw.
 *   root_states  [N][A+P][13]  pos3 quat4 linvel3(world) angvel3(world)   (gym.acquire_actor_root_state_tensor)
 *   dof_states   [N][12A+D][2] (pos, vel)                                  (gym.acquire_dof_state_tensor)
 *   contact_force[N][17A+P][3] net world-frame contact force per rigid body (gym.acquire_net_contact_force_tensor)
 */
#ifndef MQE_B200_H
#define MQE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MQE_ABI_VERSION 7
#define MQE_MAX_PROBES 32
#define MQE_MAX_CAPS 20
#define MQE_NUM_DOF 12
#define MQE_NUM_BODIES 17
#define MQE_LOC_OBS 70           /* walk-these-ways observation frame (go1.py:64-108)          */
#define MQE_HIST_FRAMES 30       /* history depth, 30 x 70 = 2100 (go1.py:395)                 */
#define MQE_OBS_FLOATS 71        /* per-agent observation struct (go1.py:153-196)              */

typedef enum {
    MQE_OK = 0,
    MQE_ERR_INVALID = -1,        /* bad argument / descriptor                                  */
    MQE_ERR_CUDA = -2,           /* a CUDA runtime call failed (message has the cudaError)     */
    MQE_ERR_NO_DEVICE = -3,      /* no sm_100 device visible: there is NO CPU fallback         */
    MQE_ERR_UNSUPPORTED = -4
} MqeStatus;

typedef enum { MQE_NPC_NONE = 0, MQE_NPC_RIGID = 1, MQE_NPC_SEESAW = 2, MQE_NPC_BOX = 3, MQE_NPC_PLATFORM = 4 } MqeNpcKind;   /* RIGID: capsule / sphere; BOX: free box; PLATFORM: fixed raised boxes */
typedef enum { MQE_NPC_PASSIVE = 0, MQE_NPC_SHEEP = 1 } MqeNpcCtrl;
typedef enum { MQE_POLICY_FP32 = 0, MQE_POLICY_BF16X3 = 1, MQE_POLICY_BF16 = 2 } MqePolicyMode;

/* Compiled robot tables (model.py; go1.urdf after Isaac Gym's fixed-joint collapse). */
typedef struct {
    float base_inertial[10];          /* m, com xyz, Ixx Ixy Ixz Iyy Iyz Izz about the COM, base frame     */
    float leg_offsets[4][4][3];       /* per leg FL,FR,RL,RR: hip joint pos, thigh off, calf off, foot off */
    float leg_inertial[4][3][10];     /* hip, thigh, calf(+foot)                                           */
    float q_lower[12], q_upper[12], qd_limit[12], tau_limit[12], q_default[12];
    int32_t n_probes, n_caps;
    float probes[MQE_MAX_PROBES][6];  /* link id, rigid-body id, centre xyz (link frame), radius           */
    float caps[MQE_MAX_CAPS][9];      /* link id, rigid-body id, p0 xyz, p1 xyz, radius                    */
} MqeRobotModel;

/* Frozen network weights, row-major [out][in] exactly as the TorchScript state_dicts hold them
 * (go1.py:367,397-398).  HOST pointers, copied at creation. */
typedef struct {
    const float *adapt_w0, *adapt_b0;   /* [256][2100]  */
    const float *adapt_w1, *adapt_b1;   /* [128][256]   */
    const float *adapt_w2, *adapt_b2;   /* [2][128]     */
    const float *body_w0, *body_b0;     /* [512][2102]  */
    const float *body_w1, *body_b1;     /* [256][512]   */
    const float *body_w2, *body_b2;     /* [128][256]   */
    const float *body_w3, *body_b3;     /* [12][128]    */
    const float *act_w0, *act_b0;       /* [32][6]      */
    const float *act_w1, *act_b1;       /* [32][32]     */
    const float *act_w2, *act_b2;       /* [1][32]      */
} MqeWeights;

/* Simulation descriptor: what gymapi.SimParams + the env cfg carry for this path. */
typedef struct {
    int32_t abi_version;
    int32_t num_envs, num_agents, num_npcs;     /* LOCAL shard sizes                                       */
    int32_t env_id_offset;                      /* global id of local env 0 (keys the counter-based RNG)   */
    int32_t npc_kind, npc_ctrl, npc_dofs;       /* MqeNpcKind, MqeNpcCtrl, NPC dofs per env                */
    int32_t decimation;                         /* go1_config.py:119                                       */
    int32_t solver_iters;                       /* contact impulse sweeps per substep                      */
    int32_t max_episode_length;                 /* ceil(episode_length_s / dt), legged_robot.py:1021       */
    int32_t term_mask;                          /* bit0 roll bit1 pitch bit2 z_low bit3 z_high bit4 base contact */
    int32_t quat_alias;                         /* 1 iff P==0: obs base_quat/base_rpy alias the sim tensor  */
    int32_t policy_mode;                        /* MqePolicyMode                                           */
    int32_t defender;                           /* 1: last agent is the scripted defender (go1_football_defender.py) */
    int32_t command_vel;                        /* cfg.command.cfg.vel: actions carry (vx, vy, wz) (go1.py:66-68)     */
    float sim_dt;                               /* legged_robot_config.py:212                              */
    float gravity_z;
    float friction, contact_offset, max_depen_vel, erp, cfm;
    float floor_z, wall_top_z;                  /* barrier_track.py:628-632 ground slab / wall_height       */
    float limit_margin;                         /* joint-limit rows become active inside this margin [rad]  */
    float term_roll, term_pitch, term_zlow, term_zhigh;
    float act_scale[3];                         /* wrapper action scale [2,.5,.5] applied before clip(+-1)  */
    float cmd_scale[3];                         /* obs_scales lin_vel, lin_vel, ang_vel (go1.py:66-68)      */
    float action_scale, hip_scale, clip_actions;/* go1_config.py:113,120,108                               */
    float loc_obs_default[MQE_LOC_OBS];         /* go1.py:411-479                                          */
    float dof_ratio_lo, dof_ratio_hi;           /* init_dof_pos_ratio_range                                */
    float base_vel_lo, base_vel_hi;
    int32_t has_base_pos_range, has_npc_pos_range, has_npc_rpy_range;
    int32_t max_pair_contacts;                  /* dynamic-vs-dynamic contacts kept per env and substep (1..16; 0 = 16)   */
    float base_pos_x[2], base_pos_y[2];
    float npc_pos_x[2], npc_pos_y[2];
    float npc_rpy_r[2], npc_rpy_p[2], npc_rpy_y[2];
    float npc_mass, npc_inertia, npc_radius, npc_halflen;
    float sheep_scale, sheep_randomness;        /* go1_sheep_config.py asset.sheep_movement_*              */
    float gate_x;                               /* defender: init+plane block length                       */
    float max_push_vel_xy;                      /* domain_rand.max_push_vel_xy (legged_robot.py:472-477)   */
    /* MQE_NPC_SEESAW geometry (resources/objects/seesaw.urdf): [0..2] revolute-y joint origin rel. the fixed base,
     * [3] plank box / COM x offset in the plank frame, [4..6] plank half extents, [7..9] platform (base box) half extents,
     * [10] column radius, [11] column length (0: none), [12] joint velocity limit [rad/s], [13] hinge axis (0: y seesaw,
     * 1: z revolving door, rotation_door.urdf), [14], [15] box centre y, z in the hinged frame.
     * MQE_NPC_BOX (resources/objects/box.urdf): [4..6] box half extents.
     * MQE_NPC_PLATFORM (fix_npc_base_link assets made of axis-aligned boxes: wrestling_field/urdf/wrestling.urdf,
     * bridge/urdf/bridge.urdf): [0] number of boxes (<= 3), then per box centre x, y, half extents x, y and top z, all
     * relative to the NPC root position.  Their tops are ground for the robots' probes; their sides are not modelled. */
    float npc_geom[16];
    uint64_t seed;
    /* static world: 2-D signed distance to the wall footprint on the BarrierTrack pixel grid */
    int32_t sdf_nx, sdf_ny;
    float sdf_cell;
    int32_t push_interval;                      /* domain_rand.push_robots: every push_interval policy steps (ceil(push_interval_s / dt),
                                                   legged_robot.py:1024) every robot's base velocity x, y is redrawn in +-max_push_vel_xy; 0 = off */
    int32_t control_type;                       /* cfg.control.control_type: 0 'C' actuator network behind the walk policy (go1.py:335-352, every
                                                   shipped task; mqe_sim_step takes 3-D commands), 1 'P' PD position targets, 2 'T' scaled torques,
                                                   3 'V' PD velocity targets (legged_robot.py:384-392; no hip scale there).  1..3 bypass the walk
                                                   policy as Go1.step does (go1.py:43-45): the caller hands [N][12A] joint actions to mqe_sim_step_joint */
    float stiffness, damping;                   /* cfg.control.stiffness / damping ['joint'] for control_type 'P'                  */
    int32_t lag_enabled;                        /* domain_rand.randomize_lag_timesteps (go1.py:337-339, 363): the actuator network's position
                                                   target is the scaled action of `lag_timesteps` _compute_torques calls (substeps) ago */
    int32_t lag_timesteps;
    float soft_dof_pos_limit;                   /* cfg.rewards.soft_dof_pos_limit (legged_robot.py:318-321): the limits substep_exceed_dof_pos_limits tests */
    const float *h_sdf;                         /* [nx][ny] host, copied                                   */
    /* per-env constants, host, copied */
    const float *h_env_origins;                 /* [N][3]                                                  */
    const float *h_agent_origins;               /* [N][A][3]                                               */
    const float *h_base_init_state;             /* [N*A][13]                                               */
    const float *h_npc_init_state;              /* [N*P][13] or NULL                                       */
    const float *h_npc_dof_default;             /* [D] or NULL                                             */
    const float *h_base_added_mass;             /* [N*A] mass added to each robot's base link (domain_rand.randomize_base_mass,
                                                   legged_robot.py:332-335: props[0].mass += U(added_mass_range)) or NULL */
    const float *h_env_friction;                /* [N] contact friction per env (domain_rand.randomize_friction, legged_robot.py:283-294,
                                                   already combined with the terrain's) or NULL: `friction` everywhere */
    const float *h_base_com_shift;              /* [N*A][3] shift of each robot's base-link centre of mass, base frame (domain_rand.randomize_com,
                                                   legged_robot_field.py:321-332: props[0].com += U(com_range)) or NULL */
    const float *h_motor_strength;              /* [N][12A] factor on the joint actions of control types P / V / T (domain_rand.randomize_motor,
                                                   legged_robot_field.py:283-291, 180-183; Go1's 'C' path never applies it, go1.py:315-354) or NULL */
    MqeRobotModel model;
    MqeWeights weights;
} MqeSimDesc;

/* Task-wrapper gather fused into the step: observation vector, per-agent reward and the running sums behind the
 * wrappers' `reward_buffer`, for the wrappers of the BASELINE tasks.  Scales are `cfg.rewards.scales.*` in the order given. */
typedef enum {
    MQE_WRAP_NONE = 0,
    MQE_WRAP_SHEEP = 1,              /* go1_sheep_wrapper.py:54-118   scale: success, contact_punishment, sheep_movement, mixed_sheep,
                                        sheep_pos_var_lin_punishment, sheep_pos_var_exp_punishment; D = 14 + 2 P + A */
    MQE_WRAP_SEESAW = 2,             /* go1_seesaw_wrapper.py:48-120  scale: x_movement, height, y_punishment, contact_punishment,
                                        agent_distance_punishment, success, fall_punishment; D = 12 + A */
    MQE_WRAP_FOOTBALL_DEFENDER = 3,  /* go1_football_wrapper.py:57-91 scale: goal, ball_gate_distance; D = 20, two reported agents */
    MQE_WRAP_PUSHBOX = 4,            /* go1_pushbox_wrapper.py        scale: box_x_movement; D = 20 + A (ids, pos/rpy self and other, gate xy,
                                        box xy, box quaternion); h_gate [N][2] */
    /* two-agent duel wrappers: D = 12 = (pos, rpy) self | other, agent 1 sees a mirrored world, the reward goes to agent 0 only */
    MQE_WRAP_WRESTLING = 5,          /* go1_wrestling_wrapper.py:9-89  scale: success, punishment */
    MQE_WRAP_BRIDGE = 6,             /* go1_bridge_wrapper.py:8-80     scale: success, punishment, target */
    MQE_WRAP_ROTATION = 7            /* go1_rotation_wrapper.py:8-103  scale: success, punishment, distance, then scale[3] = target x */
} MqeWrapperKind;
typedef struct {
    int32_t kind;
    float scale[8];
    const float *h_gate;             /* HOST: sheep / pushbox [N][2] gate position (env-relative), football defender [N][3] gate position (world) */
} MqeWrapperDesc;

typedef struct MqeSim MqeSim;

/* Buffers owned by the engine; mqe_sim_get_buffer returns the device pointer and shape so a host
 * framework can wrap them zero-copy (replaces gym.acquire_*_tensor + gymtorch.wrap_tensor,
 * legged_robot.py:554-595). */
typedef enum {
    MQE_BUF_ROOT_STATES = 0,    /* f32 [N][A+P][13]                         */
    MQE_BUF_DOF_STATES,         /* f32 [N][12A+D][2]                        */
    MQE_BUF_CONTACT_FORCES,     /* f32 [N][17A+P][3]                        */
    MQE_BUF_TORQUES,            /* f32 [N][12A]                             */
    MQE_BUF_ACTIONS,            /* f32 [N][12A]   clipped policy output     */
    MQE_BUF_LAST_ACTIONS,       /* f32 [N][12A]                             */
    MQE_BUF_OBS,                /* f32 [N*A][71]  see MQE_OBS_* offsets     */
    MQE_BUF_BASE_LIN_VEL,       /* f32 [N*A][3]                             */
    MQE_BUF_BASE_ANG_VEL,       /* f32 [N*A][3]                             */
    MQE_BUF_PROJ_GRAVITY,       /* f32 [N*A][3]                             */
    MQE_BUF_RESET,              /* u8  [N]                                  */
    MQE_BUF_TIMEOUT,            /* u8  [N]                                  */
    MQE_BUF_COLLIDE,            /* u8  [N]                                  */
    MQE_BUF_ROLL_TERM,          /* u8  [N]                                  */
    MQE_BUF_PITCH_TERM,         /* u8  [N]                                  */
    MQE_BUF_ZLOW_TERM,          /* u8  [N]                                  */
    MQE_BUF_ZHIGH_TERM,         /* u8  [N]                                  */
    MQE_BUF_EPISODE_LENGTH,     /* i64 [N]                                  */
    MQE_BUF_COMMANDS,           /* f32 [N*A][3]   clipped command given to the policy */
    MQE_BUF_LOC_OBS,            /* f32 [N*A][70]                            */
    MQE_BUF_LOC_ACTION,         /* f32 [N*A][12]  raw policy output (last_locomotion_action) */
    MQE_BUF_GAIT,               /* f32 [N*A]                                */
    MQE_BUF_HISTORY,            /* f32 [N*A][30][80] frame ring, see mqe_sim_history_head */
    MQE_BUF_SHEEP_STATS,        /* f32 [N][3]     sheep_pos_avg xy, sheep_pos_var */
    MQE_BUF_STATS,              /* i32 [8]        contact / row statistics of the last step */
    MQE_BUF_CLOCK,              /* f32 [N*A][4]   gait clock inputs (go1.py:240-279)        */
    MQE_BUF_WRAP_SUMS,          /* f64 [16]        running sums of the reward terms (scale order), [8] = steps      */
    MQE_BUF_WARP_TRACE,         /* i64 [warps][20] substep kernel trace per warp of envs: start ns, end ns, pair contacts, widest row count, then (MQE_TRACE=1) cycles per phase */
    /* post_decimation_step logs (legged_robot.py:112-115).  Nothing on the Go1 path reads them, so they are produced LAZILY: the first
     * mqe_sim_get_buffer on any of the three allocates them and switches the per-substep stores on for every following step. */
    MQE_BUF_SUBSTEP_TORQUES,    /* f32 [N][decimation][12A]                 */
    MQE_BUF_SUBSTEP_DOF_VEL,    /* f32 [N][decimation][12A]                 */
    MQE_BUF_SUBSTEP_EXCEED,     /* u8  [N][decimation][12A]  dof_pos outside the soft limits */
    MQE_BUF_HISTORY_HI,         /* u16 [ceil(N*A/128)][30][10][128][8] bf16 high plane of the frame ring (tensor-core policy modes) */
    MQE_BUF_HISTORY_LO,         /* u16 same layout, bf16 residual plane: frame = float(hi) + float(lo) to 2^-16 relative */
    MQE_BUF_STEP_RESULT,        /* u8  [2][total_bytes] what the learner reads after a step, packed for ONE copy / ONE exchange: wrapper obs
                                   f32 [N][Aw][D], reward f32 [N][Aw], done u8 [N]; see mqe_sim_step_result_layout.  Double buffered: step t
                                   writes half (t + 1) & 1 (mqe_sim_result_parity), so the previous step's result stays readable for one more
                                   step -- the reference returns fresh tensors every step, a learner may hold obs_t across env.step() */
    MQE_BUF_COUNT
} MqeBuffer;

/* offsets into one agent's MQE_BUF_OBS row (go1.py:153-196) */
#define MQE_OBS_BASE_POS 0
#define MQE_OBS_BASE_QUAT 3
#define MQE_OBS_DOF_POS 7
#define MQE_OBS_DOF_VEL 19
#define MQE_OBS_LIN_VEL 31
#define MQE_OBS_ANG_VEL 34
#define MQE_OBS_LAST_ACTION 37
#define MQE_OBS_LAST_LAST_ACTION 49
#define MQE_OBS_PROJ_GRAVITY 61
#define MQE_OBS_CLOCK 64
#define MQE_OBS_BASE_RPY 68

const char *mqe_last_error(void);
int mqe_abi_version(void);
int mqe_device_count(void);

/* gymapi.acquire_gym + create_sim + create_env/create_actor loop + prepare_sim
 * (base_task.py:40-95, legged_robot.py:255-261,754-923).  `stream` is a cudaStream_t (may be NULL). */
int mqe_sim_create(const MqeSimDesc *desc, int device, void *stream, MqeSim **out);
int mqe_sim_destroy(MqeSim *sim);
int mqe_sim_set_stream(MqeSim *sim, void *stream);
/* Scale applied to incoming actions after clip(+-1) and before Go1.step's own clip(+-1): [2,.5,.5] when the caller
 * is a task wrapper handing over raw policy actions (wrappers/go1_*_wrapper.py step()), [1,1,1] when the caller is
 * Go1.step() itself and the wrapper has already scaled (go1.py:38). */
int mqe_sim_set_action_scale(MqeSim *sim, const float scale[3]);

/* Enable the fused task-wrapper gather: every following mqe_sim_step also fills the obs / reward fields of MQE_BUF_STEP_RESULT and MQE_BUF_WRAP_SUMS, and
 * mqe_sim_reset fills the observation (the wrappers' reset()).  kind = MQE_WRAP_NONE switches it off again. */
int mqe_sim_set_wrapper(MqeSim *sim, const MqeWrapperDesc *desc);
/* the wrappers' reset() without an env reset: observation only, per-episode wrapper state cleared (used once, right after
 * mqe_sim_set_wrapper, when the env has already been reset) */
int mqe_sim_wrapper_reset(MqeSim *sim);

/* gym.acquire_*_tensor (legged_robot.py:554-557).  shape[4] is zero padded; elem_size in bytes. */
int mqe_sim_get_buffer(MqeSim *sim, int which, void **d_ptr, int64_t shape[4], int32_t *elem_size);

/* Go1.reset(): reset_idx(all) + compute_observations (go1.py:147-151). */
int mqe_sim_reset(MqeSim *sim);

/* One policy step, Go1.step() (go1.py:35-62): command -> obs frame + history -> walk policy ->
 * decimation x (actuator net -> simulate -> refresh) -> post_physics_step (terminate, NPC step,
 * indexed reset, observations).  d_actions: f32 [N][A_ctrl][3] wrapper-level actions in [-1,1]
 * (A_ctrl = A, or A-1 when desc.defender).  Asynchronous on the stream. */
int mqe_sim_step(MqeSim *sim, const float *d_actions);

/* Go1.step() for control_type 'P' / 'V' / 'T' (go1.py:43-45 -> legged_robot.py:108-110 pre_physics_step): d_joint_actions f32 [N][12A]
 * are clipped to +-clip_actions and become `actions`; no walk policy, no gait clock.  MQE_ERR_UNSUPPORTED for control_type 'C'
 * (and mqe_sim_step returns MQE_ERR_UNSUPPORTED for the other control types). */
int mqe_sim_step_joint(MqeSim *sim, const float *d_joint_actions);

/* Same step through HOST buffers: H2D of actions, step, D2H of obs rows / reset flags; blocks until
 * the results are in host memory.  h_obs: [N*A][71] (may be NULL), h_reset: [N] (may be NULL). */
int mqe_sim_step_host(MqeSim *sim, const float *h_actions, float *h_obs, uint8_t *h_reset);

/* What the learner reads after a step, packed so that ONE device->host copy (mqe_sim_step_host_result) or ONE peer exchange
 * (mqe_sim_gather_*) moves it: the fused task-wrapper observation and reward (empty when no wrapper is set, e.g. go1gate whose shipped
 * wrapper returns 0, go1_gate_wrapper.py:155) and the done flags.  Offsets in bytes into MQE_BUF_STEP_RESULT. */
typedef struct {
    int64_t obs_off, obs_bytes;          /* f32 [N][Aw][D]  */
    int64_t reward_off, reward_bytes;    /* f32 [N][Aw]     */
    int64_t done_off, done_bytes;        /* u8  [N]         */
    int64_t total_bytes;                 /* multiple of 16  */
    int32_t num_envs, Aw, D, reserved;
} MqeStepResultLayout;
int mqe_sim_step_result_layout(MqeSim *sim, MqeStepResultLayout *out);
/* half of MQE_BUF_STEP_RESULT that holds the latest result (policy steps done so far & 1; reset rewrites the current half) */
int mqe_sim_result_parity(MqeSim *sim);
/* mqe_sim_step through HOST buffers with the packed result: H2D of actions, step, ONE D2H of MQE_BUF_STEP_RESULT into h_result
 * (total_bytes); blocks until it has landed.  This is the call behind openrl_ws/utils.py:53-67 (`mqe_openrl_wrapper.step`). */
int mqe_sim_step_host_result(MqeSim *sim, const float *h_actions, void *h_result);

/* Per-step exchange between the ranks of one node over NVLink peer memory (replaces an NCCL all-gather after the step).
 * Every rank owns a receive buffer [2 parities][obs | reward | done regions, each world x local bytes]; the last kernel of the step
 * (inside the step graph) stores this rank's MQE_BUF_STEP_RESULT fields into EVERY peer's buffer, publishes a per-rank step flag
 * with system-scope release, and waits until all peers' flags for this step have arrived locally.  Equal shard sizes only.
 *   1. every rank: mqe_sim_gather_init (after mqe_sim_set_wrapper) -> 64-byte cudaIpcMemHandle;
 *   2. hosts exchange the handles (torch.distributed all_gather_object in mqe_b200/dist.py);
 *   3. every rank: mqe_sim_gather_connect(handles of all ranks, rank order).
 * After step t (t = 0, 1, ...) the global result is in the parity (t + 1) & 1 half; mqe_sim_gather_view returns its device pointer and
 * the GLOBAL layout (num_envs = world x N).  A peer that never arrives makes the wait give up after ~2 s and raises MQE_STAT_GATHER_TIMEOUT
 * in MQE_BUF_STATS[5] instead of hanging the device. */
int mqe_sim_gather_init(MqeSim *sim, int rank, int world, void *ipc_handle_out /* 64 bytes */);
int mqe_sim_gather_connect(MqeSim *sim, const void *ipc_handles /* [world][64] */);
int mqe_sim_gather_view(MqeSim *sim, int parity, void **d_ptr, MqeStepResultLayout *global_layout);
/* half of the receive buffer that holds the latest exchange (every mqe_sim_reset and every step performs one) */
int mqe_sim_gather_parity(MqeSim *sim);

/* Optional: register a caller-owned host range (cudaHostRegister) so that mqe_sim_step_host copies straight from / into it
 * instead of through the handle's staging buffers.  The range must stay mapped until mqe_sim_unpin_host / mqe_sim_destroy.
 * (The reference has no host path of its own here: openrl_ws/utils.py:55-60 does .cpu().numpy() on pageable memory.) */
int mqe_sim_pin_host(MqeSim *sim, void *h_ptr, size_t bytes);
int mqe_sim_unpin_host(MqeSim *sim, void *h_ptr);

/* Finer-grained entry points mirroring the individual gym calls (used by tests and by a host that
 * wants to keep the reference's loop structure). */
int mqe_sim_policy(MqeSim *sim, const float *d_actions);     /* preprocess_action, go1.py:64-108              */
int mqe_sim_substeps(MqeSim *sim, int count);                /* count x {_compute_torques; gym.simulate; refresh_dof_state} go1.py:48-58 */
int mqe_sim_post_physics(MqeSim *sim);                       /* post_physics_step, legged_robot.py:117-157    */

/* gym.set_actor_root_state_tensor_indexed / set_dof_state_tensor_indexed (legged_robot.py:419-421,
 * 468-470): the state buffers ARE the engine state, so writes through the wrapped views take effect
 * directly; these calls exist for hosts that stage state elsewhere.  d_actor_ids: i32 [n] actor
 * indices (DOMAIN_SIM: env*(A+P)+k). */
int mqe_sim_set_root_indexed(MqeSim *sim, const float *d_root_states, const int32_t *d_actor_ids, int n);
int mqe_sim_set_dof_indexed(MqeSim *sim, const float *d_dof_states, const int32_t *d_actor_ids, int n);

/* stand-alone operator entry points (parity tests against the TorchScript goldens) */
int mqe_policy_forward(MqeSim *sim, const float *d_history /* [rows][2100] */, int rows,
                       float *d_latent /* [rows][2] */, float *d_action /* [rows][12] */);
int mqe_actuator_forward(MqeSim *sim, const float *d_x /* [rows][6] */, int rows, float *d_torque);
/* slot of the 30-frame history ring that holds the newest frame (MQE_BUF_HISTORY is [N*A][30][80], slot-major) */
int mqe_sim_history_head(MqeSim *sim);

int mqe_sim_synchronize(MqeSim *sim);
/* number of kernels launched by this handle since creation (bench.py's gpu_launches) */
int64_t mqe_sim_launch_count(MqeSim *sim);
/* diagnostics: device time of the four stages of mqe_sim_step, from event marks that become nodes of the step graph (so they cost the
 * programmatic launch overlap between the stages: a few us; off by default).  ms4 = policy (frame, new-frame layer 0, fused tail) |
 * physics (k_substeps incl. the fused bookkeeping) | separate bookkeeping launches, task gather, peer exchange | wait for the
 * background layer-0 pass of the NEXT step.  Values are those of the last step on this handle. */
int mqe_sim_stage_timing(MqeSim *sim, int enable);
int mqe_sim_stage_ms(MqeSim *sim, float *ms4);

#ifdef __cplusplus
}
#endif
#endif /* MQE_B200_H */
