/*
 * mgb200.h -- C ABI of libmgb200.so, the B200 (sm_100a) batched environment engine for the two MetaGym dynamics
 * hot paths (quadrotor 6-DoF integrator + task logic; MetaMaze grid step + raycast render).
 *
 * The reference (PaddlePaddle/MetaGym) is pure Python and has no FFI of its own; the boundary it exposes is the
 * gym.Env protocol.  Each entry point below therefore names the reference METHOD it replaces (file:line relative to
 * the reference tree), for a batch of n independent env instances.  The Python classes in metagym_b200/ bind these
 * symbols with ctypes and re-expose the reference's method names (INTEGRATION.md shows the stub).
 *
 * Conventions
 *   - every function returns 0 on success, <0 on error; mgb_last_error() returns a thread-local message.
 *   - *_dev pointers are device pointers owned by the caller (e.g. torch tensor.data_ptr()); *_host are host pointers.
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream).  Nothing synchronises the host except
 *     the *_host entry points, mgb_*_create/destroy and the functions documented as synchronous.
 *   - a handle is bound to one device and is not thread-safe; distinct handles are independent.
 *   - there is NO CPU fallback: without a usable CUDA device every create call fails with MGB_ERR_CUDA.
 */
#ifndef MGB200_H
#define MGB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MGB_OK 0
#define MGB_ERR_ARG (-1)
#define MGB_ERR_CUDA (-2)
#define MGB_ERR_STATE (-3)

/* ------------------------------------------------------------------------------------------------------------ */
/* Quadrotor                                                                                                      */
/* ------------------------------------------------------------------------------------------------------------ */

#define MGB_TASK_NO_COLLISION 0     /* env.py:216-218 */
#define MGB_TASK_HOVERING_CONTROL 1 /* env.py:222-243 */
#define MGB_TASK_VELOCITY_CONTROL 2 /* env.py:219-221 */

#define MGB_INTEGRATOR_REFERENCE 0
#define MGB_INTEGRATOR_RK4 1

#define MGB_FAIL_NONE 0
#define MGB_FAIL_RANGE 1    /* quadrotorsim.py:213-214 */
#define MGB_FAIL_VELOCITY 2 /* quadrotorsim.py:216-217 */
#define MGB_FAIL_ANGULAR 3  /* quadrotorsim.py:219-221 */

/* Numbers of QuadrotorSim._parse_cfg (quadrotorsim.py:50-109) + the Quadrotor ctor kwargs (env.py:46-69).
 * Doubles carry the python floats of the config exactly; the library rounds derived constants to float32 the way
 * numpy >= 2 does ("weak" python scalars take the float32 type of the array they meet). */
typedef struct mgb_quad_cfg {
    double precision;          /* config.json "precision": substep h                         */
    double quality;            /* config.json "quality": mass                                */
    float inv_inertia[9];      /* np.linalg.inv(float32 inertia), row-major (quadrotorsim.py:64) */
    float drag_m[3];           /* diag of _drag_coeff_momentum                               */
    float drag_f[3];           /* diag of _drag_coeff_force                                  */
    float gravity_center[3];
    double ct[3];              /* thrust CT0..2                                              */
    double mm, jm, phi, ra;    /* thrust Mm, Jm, phi, RA                                     */
    double fail_velocity, fail_range, fail_w;
    float propeller[12];       /* 4 x (x,y,z)                                                */
    float propeller_norm[4];   /* np.linalg.norm(float32 coord) (quadrotorsim.py:146)        */
    double min_voltage, max_voltage;
    float init_velocity[3];    /* config.json "init_velocity" x,y,z                          */
    double init_velocity_noise;
    float init_angular_velocity[3];
    double init_angular_velocity_noise;
    /* Quadrotor(...) kwargs, env.py:46-53 */
    double dt;
    int32_t nt;
    int32_t task;              /* MGB_TASK_*                                                 */
    double healthy_reward;
    double z_offset;           /* env.py:112 (5.0 for the flat map; 0 for velocity_control)  */
    /* Integrator.  MGB_INTEGRATOR_REFERENCE: int(dt/precision) semi-implicit Euler substeps, the reference's only
     * integrator (quadrotorsim.py:122-208) and the parity-checked default.  MGB_INTEGRATOR_RK4: classical RK4 on the
     * same continuous-time model, rk4_steps steps of dt/rk4_steps per env step -- BASELINE.json's "RK4 dt=0.005";
     * the reference has no counterpart, so it is validated by convergence only. */
    int32_t integrator;
    int32_t rk4_steps;
} mgb_quad_cfg;

typedef struct mgb_quad mgb_quad;

/* Quadrotor.__init__ (env.py:46-114) for n_envs instances on `device`.  State starts at _zero_state
 * (quadrotorsim.py:20-28), ct = 0.  `env_index_base` is the global index of local env 0: per-env random streams are
 * keyed by the global index so that results do not depend on how envs are sharded over GPUs. */
int mgb_quad_create(mgb_quad **out, int64_t n_envs, const mgb_quad_cfg *cfg, int device, int64_t env_index_base);
void mgb_quad_destroy(mgb_quad *h);

/* Observation width: 16, or 19 for velocity_control (env.py:86-92). */
int mgb_quad_obs_dim(const mgb_quad *h);
int64_t mgb_quad_num_envs(const mgb_quad *h);

/* auto_reset != 0: an env whose step returns done is re-initialised in the same launch with counter-based noise
 * (Philox keyed by seed, global env index, episode count); obs then holds the first observation of the new episode
 * and final_obs (if given to mgb_quad_step) the terminal one.  auto_reset == 0: the reference behaviour -- state is
 * left as is and the caller resets (env.py:116).  Synchronous w.r.t. nothing; takes effect at the next launch. */
int mgb_quad_set_options(mgb_quad *h, int auto_reset, uint64_t seed);

/* Quadrotor.load_map + the map part of __init__ (env.py:97-114, 293-305): map_host [rows][cols] int32 with exactly one
 * -1 (the start cell, which becomes 0); x_offset / y_offset are its column / row.  Only the TRUTHINESS of the cells in
 * the window swept by a step matters to the reference's _check_collision (env.py:248-260: `z < np.any(taken_pos)`),
 * including python's negative-index slice semantics; both are reproduced.  NULL map = the flat default (env.py:295-298).
 * no_collision / hovering_control only.  Synchronous. */
int mgb_quad_set_map(mgb_quad *h, const int32_t *map_host, int32_t rows, int32_t cols);

/* Velocity targets of define_velocity_control_task (quadrotorsim.py:306-319): tbl [n_tasks][nt][3] float32 and the
 * task row of every local env, env2task [n_envs] int32.  Both are COPIED into the handle (synchronous). */
int mgb_quad_set_targets(mgb_quad *h, const float *tbl_dev, int32_t n_tasks, const int32_t *env2task_dev);

/* Runs define_velocity_control_task (quadrotorsim.py:306-319) on the device for n_tasks seeds: act_host [n_tasks][nt][4] float32 are the
 * np.random.uniform draws (host-replayed for RNG parity); tbl_dev [n_tasks][nt][3] receives global_velocity after
 * every step from the zero state.  Independent of the handle's env state. */
int mgb_quad_make_targets(mgb_quad *h, const float *act_dev, int32_t n_tasks, float *tbl_dev, void *stream);

/* Quadrotor.reset (env.py:116-125 -> quadrotorsim.py:239-258).  mask_dev [n] uint8 (NULL = all envs).
 * noise_dev [n][12] float64 = the twelve np.random.random() draws of one reset in reference order (sign_v[3],
 * mag_v[3], sign_w[3], mag_w[3]); NULL = counter-based draws.  ct is NOT touched (reference quirk, env.py:65,150).
 * obs_dev [n][obs_dim] (NULL = skip) receives the observation of every env (masked-out envs: current state). */
int mgb_quad_reset(mgb_quad *h, const uint8_t *mask_dev, const double *noise_dev, float *obs_dev, void *stream);

/* Quadrotor.step (env.py:127-165): int(dt/precision) substeps of QuadrotorSim._run_internal (quadrotorsim.py:122-221),
 * get_sensor/get_state (:260-293), reward / collision / done (env.py:211-260).
 *   act_dev  [n][4] float32           obs_dev [n][obs_dim] float32      rew_dev [n] float32     done_dev [n] uint8
 *   fail_dev [n] int32 or NULL  (MGB_FAIL_*: the reference raises, the batch reports done + code)
 *   final_obs_dev [n][obs_dim] or NULL (terminal observation of envs that finished, when auto_reset is on) */
int mgb_quad_step(mgb_quad *h, const float *act_dev, float *obs_dev, float *rew_dev, uint8_t *done_dev,
                  int32_t *fail_dev, float *final_obs_dev, void *stream);

/* T consecutive Quadrotor.step calls (env.py:127-165; the rollout loop of quadrotor/tests/test_env.py:22-28) in ONE launch
 * with the state held in registers (auto-reset semantics as configured).
 *   act_dev [T][n][4] or NULL: NULL draws U(min_voltage, max_voltage) actions from the counter-based generator
 *   (stream id `act_seed`), written to act_out_dev [T][n][4] if not NULL.
 *   obs_dev [T][n][obs_dim], rew_dev [T][n], done_dev [T][n]; any of them may be NULL to skip that output. */
int mgb_quad_rollout(mgb_quad *h, int32_t T, const float *act_dev, uint64_t act_seed, float *act_out_dev,
                     float *obs_dev, float *rew_dev, uint8_t *done_dev, void *stream);

/* Quadrotor.step as the reference's numpy users call it (env.py:127-165: ndarray in, ndarray out).
 * Same as mgb_quad_step with HOST buffers: stages through pinned memory (or, for pinned caller buffers, lets the kernel
 * read/write host memory directly), copies inside the call, returns when the outputs are on the host (synchronous).
 * All work is enqueued on `stream` (the caller's current stream), so the step is ordered after a preceding
 * mgb_quad_reset / mgb_quad_rollout / mgb_quad_state on that stream.  fail_host [n] int32 and final_obs_host
 * [n][obs_dim] may be NULL; they report what mgb_quad_step's fail_dev / final_obs_dev report (quadrotorsim.py:212-221).
 * This is the call a numpy user of the reference API makes. */
int mgb_quad_step_host(mgb_quad *h, const float *act_host, float *obs_host, float *rew_host, uint8_t *done_host,
                       int32_t *fail_host, float *final_obs_host, void *stream);

/* Checkpoint / inspection (quadrotorsim.py:30-48 _save_state/_restore_state): state_dev [n][22] float32 row-major
 * = p3 v3 w3 prop4 R9, ct_dev [n] int32.  load = 0 copies handle -> buffers, 1 buffers -> handle. */
int mgb_quad_state(mgb_quad *h, float *state_dev, int32_t *ct_dev, int load, void *stream);

/* Number of kernel launches issued through this handle so far (bench.py reports it as gpu_launches). */
int64_t mgb_quad_launch_count(const mgb_quad *h);

/* Name of the kernel an mgb_quad_step launch of this handle takes at its batch size ("quad_step_wide_kernel<..>": one
 * CTA per SM for single-wave batches; "quad_stream_kernel<..>": persistent TMA-pipelined variant for multi-wave batches;
 * "quad_step_kernel<..>": 64-env CTAs otherwise; "quad_step2_kernel<..>": the packed two-envs-per-thread variant, MGB_PACKED=1).
 * Reporting only (bench.py's roofline.kernel); no reference counterpart. */
const char *mgb_quad_step_kernel(const mgb_quad *h);

/* ------------------------------------------------------------------------------------------------------------ */
/* MetaMaze (2D grid + discrete-3D raycast)                                                                       */
/* ------------------------------------------------------------------------------------------------------------ */

#define MGB_MAZE_2D 0          /* MazeCore2D, maze_2d.py:13          */
#define MGB_MAZE_DISCRETE_3D 1 /* MazeCoreDiscrete3D, maze_discrete_3d.py:17 */
#define MGB_MAZE_CONTINUOUS_3D 2 /* MazeCoreContinuous3D, maze_continuous_3d.py:16 (+ dynamics.py) */

#define MGB_MAZE_SURVIVAL 0 /* maze_base.py:52-57,72-88 */
#define MGB_MAZE_ESCAPE 1   /* maze_base.py:58-60,90-93 */

#define MGB_OBS_U8 0  /* min(value, 255) as uint8 (3-D only; reference values can reach ~350 on near-floor pixels) */
#define MGB_OBS_I32 1 /* exact reference values as int32 (ray_caster_utils.py:79)                                 */
#define MGB_OBS_F32 2 /* the same exact values as float32: the dtype observation_space declares (maze_env.py:37-39)  */

/* Scalars of one TaskConfig (maze_task.py:15-17,176-190). */
typedef struct mgb_maze_task_scalars {
    int32_t start[2];
    int32_t goal[2];
    double cell_size, wall_height, agent_height;
    double initial_life, max_life, step_reward, goal_reward;
} mgb_maze_task_scalars;

typedef struct mgb_maze_cfg {
    int32_t kind;        /* MGB_MAZE_*                                                      */
    int32_t task_type;   /* MGB_MAZE_SURVIVAL / ESCAPE                                      */
    int32_t n_cells;     /* maze side n (odd, maze_task.py:56-57); <= 31                    */
    int32_t max_steps;   /* maze_base.py:191-192                                            */
    int32_t view_grid;   /* 2-D: half window g, obs (2g+1)^2 (maze_2d.py:89-121)            */
    int32_t res_h;       /* 3-D: resolution_horizon                                         */
    int32_t res_v;       /* 3-D: resolution_vertical                                        */
    int32_t obs_dtype;   /* 3-D: MGB_OBS_*                                                  */
    double max_vision;   /* 12.0, maze_discrete_3d.py:22                                    */
    double fov;          /* 0.6 * 3.1415926, maze_discrete_3d.py:23                         */
    double l_focal;      /* 0.20, maze_discrete_3d.py:116                                   */
    double text_size;    /* 1.0, maze_discrete_3d.py:116                                    */
} mgb_maze_cfg;

typedef struct mgb_maze mgb_maze;

/* MetaMaze2D.__init__ / MetaMazeDiscrete3D.__init__ (maze_env.py:156-172, 17-42) for n_envs instances. */
int mgb_maze_create(mgb_maze **out, int64_t n_envs, const mgb_maze_cfg *cfg, int device, int64_t env_index_base);
void mgb_maze_destroy(mgb_maze *h);
int64_t mgb_maze_obs_bytes_per_env(const mgb_maze *h);

/* MazeTaskManager textures (maze_task.py:19-35): grounds [n_tex][ts][ts][3] (x-major like pygame.surfarray) and
 * ceil [ts][ts][3], HOST pointers, uint8 (the reference stores the same integers as float32).  ts must be 64. */
int mgb_maze_set_textures(mgb_maze *h, const uint8_t *grounds_host, int32_t n_tex, const uint8_t *ceil_host,
                          int32_t tex_size);

/* MazeBase.set_task (maze_base.py:19-38) for a table of n_tasks TaskConfigs and the task of every local env.
 * HOST pointers, copied (synchronous): walls/texts int8 [n_tasks][n][n], food_rewards float64 [n_tasks][n][n],
 * food_interval int32 [n_tasks][n][n], scalars [n_tasks], env2task int32 [n_envs]. */
int mgb_maze_set_task(mgb_maze *h, int32_t n_tasks, const int8_t *walls_host, const int8_t *texts_host,
                      const double *food_rewards_host, const int32_t *food_interval_host,
                      const mgb_maze_task_scalars *scalars_host, const int32_t *env2task_host);

/* MazeTaskSampler keyword arguments (maze_task.py:41-54; defaults in metagym_b200/metamaze.py). */
typedef struct mgb_maze_sampler_cfg {
    int32_t allow_loops, n_texts, food_interval, pad;
    double cell_size, wall_height, agent_height, step_reward;
    double goal_reward;         /* <= 0: the reference's default -sqrt(n) * n * step_reward (maze_task.py:163-166) */
    double food_reward, initial_life, max_life, food_density, crowd_ratio;
} mgb_maze_sampler_cfg;

/* Per-episode task resampling ON THE DEVICE (maze_task.py:41-190 at the scale of SURVEY.md 8f row 3): every env e with
 * mask_dev[e] != 0 (NULL: all envs) gets a freshly drawn maze written into its task-table slot and starts an episode on
 * it; one kernel, stream-ordered, no host involvement (typical use: mask = the `done` array of the previous step).  Draws
 * come from a counter-based generator keyed by (seed, global env index, how often the env has been resampled), so results
 * do not depend on sharding.  The distribution family is MazeTaskSampler's (spanning tree of the room lattice, loops down
 * to crowd_ratio, textures, start/goal, thinned food); it is NOT sample-identical to the reference, which draws from
 * Python's and numpy's global MT19937 streams.  Needs one table slot per env (mgb_maze_set_task with an injective
 * env2task) and the direct renderer; food cells per task are capped at the table's largest task. */
int mgb_maze_resample_tasks(mgb_maze *h, const uint8_t *mask_dev, const mgb_maze_sampler_cfg *cfg, uint64_t seed,
                            void *stream);

/* Read tasks back from the table (synchronous; inspection / tests): arrays as for mgb_maze_set_task, [count] long. */
int mgb_maze_get_tasks(mgb_maze *h, int32_t count, const int32_t *task_slots_host, int8_t *walls_host, int8_t *texts_host,
                       double *food_rewards_host, int32_t *food_interval_host, mgb_maze_task_scalars *scalars_host);

/* MetaMazeDiscrete3D renderer choice.  enabled = 1 (default): static layers of every (task, cell, heading) are rendered once
 * and memoised (pose cache, within MGB_MAZE_CACHE_GB), a step composes / copies; 0: every frame is ray-cast directly
 * (ray_caster_utils.py:66-209 per frame, like the reference) -- the mode for task tables that change every episode. */
int mgb_maze_set_cache(mgb_maze *h, int enabled);

/* Pose-cache statistics after the first reset/step (reporting only): out[0] cached poses, out[1] extra variant frames,
 * out[2] variant bits in use (poses whose image depends on k <= bits foods have all 2^k finished frames), out[3] bytes,
 * out[4..12] poses by k (0..7, and 8 = eight or more), out[13] 1 if the cache is in use. */
int mgb_maze_cache_info(const mgb_maze *h, int64_t out[16]);

/* Per-episode task resampling (MazeBase.set_task on a fresh TaskConfig every episode, maze_base.py:19-38, at the scale of
 * SURVEY.md 8f row 3): replace `count` entries of the table mgb_maze_set_task built -- task_slots_host [count] indices into
 * it, the other arrays as for mgb_maze_set_task but [count] long -- STREAM-ORDERED and without any device synchronisation
 * (one pinned-staged copy + three small kernels on `stream`).  Every env whose env2task entry is one of the replaced slots
 * starts a new episode on its new task (agent at start, life = initial_life, food restored), like set_task + reset of that
 * env; other envs are untouched.  The table's shape is fixed by mgb_maze_set_task: a replacement may not have more food
 * cells than the table's largest task nor smaller cells than its smallest.  Needs the direct renderer (the pose cache
 * memoises whole task tables): create the env with the cache off or with more tasks than the cache budget holds. */
int mgb_maze_update_tasks(mgb_maze *h, int32_t count, const int32_t *task_slots_host, const int8_t *walls_host,
                          const int8_t *texts_host, const double *food_rewards_host, const int32_t *food_interval_host,
                          const mgb_maze_task_scalars *scalars_host, void *stream);

/* MazeBase.reset (maze_base.py:40-63, maze_discrete_3d.py:39-49).  mask_dev NULL = all.  obs_dev NULL = skip. */
int mgb_maze_reset(mgb_maze *h, const uint8_t *mask_dev, void *obs_dev, void *stream);

/* MetaMaze*.step (maze_env.py:59-75,189-206): DISCRETE_ACTIONS[a] (maze_env.py:14) -> do_action -> evaluation_rule
 * (maze_base.py:65-95) -> update_observation (maze_2d.py:89-121 | maze_discrete_3d.py:113-127 + ray_caster_utils.py).
 *   act_dev [n] int32 in 0..3; obs_dev: 2-D float32 [n][2g+1][2g+1]; 3-D uint8|int32 [n][res_h][res_v][3];
 *   rew_dev [n] float64 (the reference returns python/np float64); done_dev [n] uint8.
 * With auto_reset on (mgb_maze_set_options) a finished env is reset in the same launch and obs holds the first
 * observation of the next episode. */
int mgb_maze_step(mgb_maze *h, const int32_t *act_dev, void *obs_dev, double *rew_dev, uint8_t *done_dev,
                  void *stream);
int mgb_maze_set_options(mgb_maze *h, int auto_reset);

/* MetaMaze2D and MetaMazeDiscrete3D: T consecutive step() calls (maze_env.py:59-75,189-206; the random-action loops of
 * metamaze/test.py:9-47) in ONE launch (agent state in registers; auto-reset semantics
 * as configured; the 3-D form runs on the pose cache: one CTA per env, step logic by one thread, frame by the CTA).
 *   act_dev [T][n] int32 or NULL: NULL draws uniform {0..3} actions from the counter-based generator (stream id
 *   act_seed, keyed by the global env index), written to act_out_dev [T][n] if not NULL.
 *   obs_dev [T][n][obs of one env] (2-D: float32 [2g+1][2g+1]; 3-D: uint8 or int32 [res_h][res_v][3]),
 *   rew_dev [T][n] float64, done_dev [T][n] uint8 (any may be NULL). */
int mgb_maze_rollout(mgb_maze *h, int32_t T, const int32_t *act_dev, uint64_t act_seed, int32_t *act_out_dev,
                     void *obs_dev, double *rew_dev, uint8_t *done_dev, void *stream);

/* MetaMazeContinuous3D.step (maze_env.py:129-146 -> maze_continuous_3d.py:47-56, dynamics.py:58-92): act_dev [n][2]
 * float32 = (turn_rate, walk_speed), clipped to [-1, 1] like the reference; ten 10 ms sub-steps of turn/walk with the
 * soft wall-repulsion collision model, then evaluation_rule and the ray-cast observation (same renderer).  Typing follows
 * what the reference computes for float32 actions (its action_space.sample()): float32 position, float64 heading. */
int mgb_maze_step_continuous(mgb_maze *h, const float *act_dev, void *obs_dev, double *rew_dev, uint8_t *done_dev,
                             void *stream);
/* Continuous pose (maze_continuous_3d.py:47-56, dynamics.py:71-92): pos_dev [n][2] float32 (_agent_loc), ori_dev [n]
 * float64 (_agent_ori). */
int mgb_maze_pose(mgb_maze *h, float *pos_dev, double *ori_dev, void *stream);

/* Inspection (info["steps"] maze_env.py:73; _agent_grid, _agent_ori_index, _life of maze_base.py:40-63): agent [n][4]
 * int32 = grid_x, grid_y, ori_index, steps; life [n] float64. */
int mgb_maze_state(mgb_maze *h, int32_t *agent_dev, double *life_dev, void *stream);
int64_t mgb_maze_launch_count(const mgb_maze *h);

/* ------------------------------------------------------------------------------------------------------------ */
const char *mgb_last_error(void);
const char *mgb_version(void);
/* Number of CUDA devices visible, or <0 when the runtime cannot initialise (no fallback exists). */
int mgb_device_count(void);

/* ---- peer memory: the rollout kernels as their own all-gather (NVLink stores, no NCCL on the data) -----------------
 * Replaces the `comm.allgather(trajectories)` a distributed learner would do around N copies of the reference's
 * env.step loop.  One receive arena per rank, allocated with mgb_peer_alloc, exported as a 64-byte
 * cudaIpcMemHandle, opened by every other rank of the node; mgb_*_set_mirrors then makes the fused rollout kernels
 * store each output at `ptr` and at `ptr + byte_delta[i]`, i < count <= MGB_MAX_MIRRORS. */
#define MGB_MAX_MIRRORS 7
#define MGB_PEER_HANDLE_BYTES 64
int mgb_peer_alloc(int device, uint64_t bytes, void **ptr_out);
int mgb_peer_free(int device, void *ptr);
int mgb_peer_export(int device, void *ptr, uint8_t handle_out[MGB_PEER_HANDLE_BYTES]);
int mgb_peer_open(int device, const uint8_t handle[MGB_PEER_HANDLE_BYTES], void **ptr_out);
int mgb_peer_close(int device, void *ptr);
int mgb_quad_set_mirrors(mgb_quad *h, int count, const int64_t *byte_delta);
int mgb_maze_set_mirrors(mgb_maze *h, int count, const int64_t *byte_delta);
/* Guard: while mirrors (or multicast) are on, a rollout whose obs/rew/done/act_out pointers are not all inside
 * [base, base + bytes) -- this rank's slot of the arena the deltas were computed for -- is refused with MGB_ERR_ARG
 * instead of storing to `pointer + delta` somewhere else.  bytes = 0 removes the guard. */
int mgb_quad_set_mirror_window(mgb_quad *h, const void *base, uint64_t bytes);
int mgb_maze_set_mirror_window(mgb_maze *h, const void *base, uint64_t bytes);
/* NVSwitch multicast variant: the rollout outputs are stored ONLY at `ptr + byte_delta` with multimem.st, where
 * byte_delta = (multicast mapping base - local arena base) of a multicast object every rank has bound its arena to
 * (cuMulticast*; torch.distributed._symmetric_memory does that plumbing).  The switch replicates each store into every
 * rank's arena, this rank's included.  0 switches it off.  Needs num_envs % 4 == 0. */
int mgb_quad_set_multicast(mgb_quad *h, int64_t byte_delta);
int mgb_maze_set_multicast(mgb_maze *h, int64_t byte_delta);

#ifdef __cplusplus
}
#endif
#endif /* MGB200_H */
