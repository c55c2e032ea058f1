/* C ABI of libso100_b200.so -- the B200-native batched replacement for the hot path of
 * gym_so100's bin-a-cube env:  reset()/step() of SO100Env (task "so100_cube_to_bin") and
 * SO100GoalEnv, i.e. everything the reference executes between receiving an action and
 * returning (obs, reward, terminated, truncated, info), for N independent envs at once.
 *
 * Plain pointers and sizes only; no torch / CUDA types in the signatures.  Unless a parameter
 * says "host", pointers are DEVICE pointers on the device given to so100_create and calls are
 * enqueued on `stream` (a cudaStream_t passed as void*, NULL = legacy default stream) without
 * any host synchronisation.  Every function returns 0 on success or a negative so100_status;
 * so100_last_error() describes the last failure on the calling thread.  A handle is not
 * re-entrant; use one handle (and one process) per GPU.
 *
 * Batched layouts (row-major, one row per env):
 *   action  float32 [N,6]   normalised joint targets in [-1,1]          (env.py:75-77)
 *   obs     float32 [N,15]  box(3) bin(3) ee(3) qpos(6), "so100_state"  (env.py:137-145)
 *   achieved/desired float32 [N,3]  GoalEnv goals                        (env.py:360-370)
 *   reward  float32 [N];  terminated/truncated/success uint8 [N]
 *   qpos float32 [N,13], qvel [N,12], ctrl [N,6], warm [N,12]  (MuJoCo qpos/qvel/ctrl/qacc_warmstart)
 */
#ifndef SO100_B200_H_
#define SO100_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct so100_ctx* so100_handle;

enum so100_status {
  SO100_OK = 0,
  SO100_ERR_ARG = -1,      /* bad argument / model blob */
  SO100_ERR_CUDA = -2,     /* CUDA runtime error */
  SO100_ERR_MODEL = -3,    /* model does not fit the compiled kernel specialisation */
  SO100_ERR_NOMEM = -4
};

enum so100_task {
  SO100_TASK_CUBE_TO_BIN = 0,  /* SO100Env(task="so100_cube_to_bin"): staged reward, TimeLimit 700 */
  SO100_TASK_GOAL = 1,         /* SO100GoalEnv: sparse HER reward, truncation at 300 steps       */
  SO100_TASK_TOUCH_CUBE = 2,   /* SO100Env(task="so100_touch_cube"): shaped reward, TimeLimit 300 (single_arm.py:149-215) */
  SO100_TASK_TOUCH_CUBE_SPARSE = 3 /* SO100Env(task="so100_touch_cube_sparse"): -0.2 / 4, TimeLimit 300 (single_arm.py:246-285) */
};

#define SO100_MAX_CONTACTS 24   /* per-env contact capacity; overflow is counted in diagnostics */
#define SO100_NDIAG 8

/* Replaces: SO100Env.__init__/_make_env_task -> mujoco.Physics.from_xml_path + control.Environment
 * (gym_so100/env.py:29-77, 92-128).  `model_blob` (host) is the packed so100_model
 * (include/so100_model.h) produced by gym_so100_c_b200.model.pack(); `env_offset` is the global
 * index of this handle's env 0 (multi-GPU sharding: RNG streams depend on the global index only).  All live handles of a
 * process on one device must use the same model (its constants sit in that device's __constant__ memory); a different one
 * fails with SO100_ERR_MODEL.  Every entry point switches to the handle's device and restores the caller's current device;
 * on any failure nothing stays allocated. */
int so100_create(const void* model_blob, size_t nbytes, int num_envs, int device, int task,
                 uint64_t seed, int64_t env_offset, so100_handle* out);
int so100_destroy(so100_handle h);
int so100_num_envs(so100_handle h);
/* Kernel launches (CUDA-graph kernel nodes) one so100_step enqueues: 64 per env group (so100_b200.cu: EnvGroup). */
int so100_launches_per_step(so100_handle h);

/* Replaces: SO100Env.reset / SO100GoalEnv.reset (env.py:148-170, 302-320) incl.
 * sample_so100_box_pose (utils.py:18-29) and initialize_episode (single_arm.py:299-309).
 * mask: uint8 [N] or NULL (= all envs).  box_pose: float32 [N,7] (xyz + wxyz) or NULL; NULL draws
 * the pose on the device (Philox keyed by seed / global env index / episode).  Outputs may be NULL. */
int so100_reset(so100_handle h, const uint8_t* mask, const float* box_pose, float* obs,
                float* achieved, float* desired, void* stream);

/* Replaces: SO100Env.step / SO100GoalEnv.step (env.py:172-182, 372-406): before_step
 * (single_arm.py:33-38), physics.step(10) + trailing mj_step1, get_reward (single_arm.py:322-380),
 * get_observation / _format_raw_obs.  autoreset != 0 gives SB3-VecEnv semantics (the reference's
 * consumers, scripts/train_sac.py:294-301): envs that finish are reset in the same call, `obs`
 * holds the first observation of the new episode and `final_obs` (nullable) the terminal one.
 * All output pointers except `obs` may be NULL. */
int so100_step(so100_handle h, const float* action, int autoreset, float* obs, float* achieved,
               float* desired, float* reward, uint8_t* terminated, uint8_t* truncated,
               uint8_t* success, float* final_obs, void* stream);

/* Same as so100_step with HOST buffers (pageable or pinned; with page-locked result buffers every env group copies its slice
 * out inside the step's own pipeline, pageable ones are copied after the step): copies the actions in, steps, copies
 * the results out and synchronises `stream` before returning -- what a CPU-side caller of the
 * reference's env.step sees. */
int so100_step_host(so100_handle h, const float* action, int autoreset, float* obs, float* achieved,
                    float* desired, float* reward, uint8_t* terminated, uint8_t* truncated,
                    uint8_t* success, float* final_obs, void* stream);

/* Replaces: SO100GoalEnv.compute_reward on batches (env.py:341-353), used by HER relabelling. */
int so100_compute_reward(const float* achieved, const float* desired, int64_t n, float threshold,
                         float* reward, void* stream);

/* MuJoCo data.qpos/qvel/ctrl/qacc_warmstart access (physics.data.*, single_arm.py:46,54,304-307);
 * also the injection point of the parity tests.  Any pointer may be NULL. */
int so100_get_state(so100_handle h, float* qpos, float* qvel, float* ctrl, float* warm, void* stream);
int so100_set_state(so100_handle h, const float* qpos, const float* qvel, const float* ctrl,
                    const float* warm, void* stream);
/* goal float32 [N,3]; step_count int32 [N]; total_steps int32 [N]; episode uint32 [N] */
int so100_get_aux(so100_handle h, float* goal, int32_t* step_count, int32_t* total_steps,
                  uint32_t* episode, void* stream);
int so100_set_aux(so100_handle h, const float* goal, const int32_t* step_count,
                  const int32_t* total_steps, const uint32_t* episode, void* stream);

/* Parity / debug: advance the physics only (`nsub` x mj_step with the stored ctrl). */
int so100_substeps(so100_handle h, int nsub, void* stream);
/* Parity / debug: mj_forward on the stored state.  qacc float32 [N,12]; ncon int32 [N];
 * con_geom int32 [N,SO100_MAX_CONTACTS,2] (MuJoCo geom ids, geom1 first);
 * con_data float32 [N,SO100_MAX_CONTACTS,11] = dist, pos[3], normal[3], force[4];
 * sites float32 [N,3,3] = cube_site, bin_center, ee_site.  Any pointer may be NULL. */
int so100_forward(so100_handle h, float* qacc, int32_t* ncon, int32_t* con_geom, float* con_data,
                  float* sites, void* stream);

/* Host counters accumulated on the device since create (synchronises the stream):
 * [0] contact-capacity overflows, [1] solver runs that hit the iteration cap, [2] non-finite
 * states forced to reset, [3] episodes finished, [4] successes, [5] total Newton iterations,
 * [6] total solver runs, [7] total contacts seen. */
int so100_diagnostics(so100_handle h, int64_t* out8, void* stream);

/* Replaces: gymnasium's RecordEpisodeStatistics around the env (scripts/train_sac.py:290, scripts/train_sac_her.py:226).
 * The task layer keeps every env's running episode return and length on the device.  so100_set_episode_outputs registers
 * optional per-env outputs (device; float32 [N] / int32 [N]; either may be NULL): at the step an env's episode ends they
 * receive its return and length (info["episode"]["r"], ["l"]); other entries are left untouched.  Call it before the first
 * so100_step, or expect one more graph capture.  so100_episode_stats sums over all envs since create (host double[4]:
 * episodes finished, successes, sum of episode returns, sum of episode lengths; synchronises the stream): the vector that
 * gym_so100_c_b200.parallel.all_reduce_stats adds up over the GPUs. */
int so100_set_episode_outputs(so100_handle h, float* ep_return, int32_t* ep_length);
int so100_episode_stats(so100_handle h, double* out4, void* stream);

/* Replaces: stable_baselines3 HerReplayBuffer(n_sampled_goal, goal_selection_strategy="future") as the reference uses it
 * around SO100GoalEnv (scripts/train_sac_her.py:231-246), kept on the device.  The ring holds `capacity` steps of `num_envs`
 * envs in CALLER-owned device arrays (row-major [capacity, num_envs, width]); ep_start / ep_length / cur_* are its episode
 * bookkeeping (int32, zero-initialised by the caller).  capacity must exceed the longest episode (GoalEnv: 300 steps).
 * Per env step: so100_her_begin(ring, pos, obs, achieved, desired, action) with the observation the action was chosen on,
 * then so100_step(..., autoreset = 1, final_obs != NULL), then so100_her_commit(ring, pos, new obs, new achieved, final_obs,
 * reward, terminated, truncated); pos advances modulo capacity.  so100_her_sample draws `batch` transitions of finished
 * episodes (Philox keyed by seed / call / sample index): the first batch / (n_sampled_goal + 1) keep their goal and reward,
 * the others are relabelled with the achieved goal of a later step of the same episode and rewarded by the arithmetic of
 * so100_compute_reward (env.py:341-353).  index int32 [batch,3] = (ring position, env, relabelling position or -1);
 * (-1,-1,-1) marks a sample for which no finished episode was found. */
typedef struct so100_her_ring {
  int32_t capacity, num_envs;
  float *obs, *next_obs;                        /* [capacity, num_envs, 15] */
  float *achieved, *next_achieved, *desired;    /* [capacity, num_envs, 3]  */
  float *action;                                /* [capacity, num_envs, 6]  */
  float *reward;                                /* [capacity, num_envs]     */
  uint8_t *done;                                /* terminated and not truncated (SB3: dones * (1 - timeouts)) */
  int32_t *ep_start, *ep_length;                /* [capacity, num_envs]: start slot / length of the transition's episode (0: running) */
  int32_t *cur_start, *cur_length;              /* [num_envs]: the running episode */
} so100_her_ring;
int so100_her_begin(const so100_her_ring* ring, int32_t pos, const float* obs, const float* achieved, const float* desired,
                    const float* action, void* stream);
int so100_her_commit(const so100_her_ring* ring, int32_t pos, const float* obs, const float* achieved, const float* final_obs,
                     const float* reward, const uint8_t* terminated, const uint8_t* truncated, void* stream);
int so100_her_sample(const so100_her_ring* ring, int64_t batch, int32_t n_sampled_goal, float threshold, uint64_t seed, uint32_t call,
                     float* obs, float* action, float* next_obs, float* achieved, float* next_achieved, float* desired, float* reward,
                     uint8_t* done, int32_t* index, void* stream);

/* Replaces (approximately -- see DESIGN.md section 11): physics.render(height, width, camera_id="top") behind obs_type
 * "so100_pixels_agent_pos" (gym_so100/env.py:50-66, 79-90, 130-136; gym_so100/tasks/single_arm.py:87-91).  A ray-caster over the
 * scene's collision geometry, not MuJoCo's OpenGL pipeline.  so100_render_config (host arrays, copied): `planes` float32
 * [nplanes,4] = body-frame facets n.x <= w of the convex hulls; per collidable geom (library order, 25 entries) the first plane and
 * the plane count (> 0 hull, 0 = box geom drawn from its size, < 0 = not drawn) and an rgb colour; `camera13` = position, x (right),
 * y (up), z (backward) axes in the world and fovy in degrees; `lights` float32 [nlights,4] = direction of travel + diffuse
 * intensity (nlights <= 4); headlight ambient / diffuse.  so100_render writes uint8 [N, height, width, 3] (device) for the current
 * state of every env. */
int so100_render_config(so100_handle h, const float* planes, int32_t nplanes, const int32_t* geom_plane_adr,
                        const int32_t* geom_plane_num, const float* geom_rgb, const float* camera13, const float* lights,
                        int32_t nlights, float ambient, float head_diffuse, int32_t width, int32_t height);
int so100_render(so100_handle h, uint8_t* pixels, void* stream);

/* Step-graph cache of this handle (host ints, any may be NULL): graphs captured since create, graphs cached now, and whether
 * the handle has switched to staging outputs because the caller keeps rotating its output pointers (so100_b200.cu). */
int so100_graph_stats(so100_handle h, int32_t* captures, int32_t* cached, int32_t* staged);

/* Measurement aid (bench.py roofline): when enabled, every kernel launched by so100_step is bracketed by a CUDA
 * event pair on `stream`.  A call first reports (if ms6 / launches6 are non-NULL; synchronises the stream) the
 * accumulated device time and launch count per kernel class since the previous call -- [0] kinematics+dynamics,
 * [1] collision, box stage, [2] constraint solve + integration (<= 8 contacts), [3] task layer, [4] collision, GJK/EPA
 * queue, [5] constraint solve, heavy queue -- then clears the record and sets the mode. */
int so100_phase_timing(so100_handle h, int enable, float* ms6, int32_t* launches6, void* stream);

/* Measurement aid (SURVEY.md 8d "the builder must measure an FFMA-loop peak"): runs a register-resident FFMA loop (8
 * independent chains per thread, 8 blocks of 256 threads per SM) on `device` and returns the achieved FP32 rate in TFLOP/s
 * (FMA = 2 flops), timed with CUDA events.  bench.py quotes the step's FP32 fraction against this number. */
int so100_measure_fp32_peak(int device, float* tflops);

/* Development aid (SO100_GROUP_TIMES=1 in the environment at create): device time from the start of the last so100_step
 * to the completion of each env group's pipeline.  ms: host float[32]; *ngroups receives the count (0 when disabled). */
int so100_group_times(so100_handle h, float* ms, int32_t* ngroups, void* stream);

/* Development aid: raw copy of the per-env records.  what = 0: state record, 1: phase workspace (link frames, mass
 * matrix, contact list, collision statistics).  `words_per_env` (host, nullable) receives the record length in float
 * words; `out` (device, nullable) receives [N, words_per_env] floats. */
int so100_debug_read(so100_handle h, int what, float* out, int64_t* words_per_env, void* stream);

const char* so100_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* SO100_B200_H_ */
