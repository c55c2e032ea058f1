// Device-resident hindsight-experience-replay ring for the GoalEnv rollout (the consumer contract of
// scripts/train_sac_her.py:231-246: SB3 HerReplayBuffer(n_sampled_goal = 4, goal_selection_strategy = "future") around
// SO100GoalEnv.compute_reward, env.py:341-353).  The ring is [capacity, num_envs] transitions in caller-owned device arrays
// (so100_her_ring, include/so100_b200.h); three entry points keep it, each a per-env / per-sample bookkeeping kernel plus an
// element-wise (coalesced) copy kernel:
//   her_begin_kernel   before env.step: observation / goals / action of the step into slot `pos`; a finished episode that
//                      the slot still belongs to is invalidated as a whole (SB3: "the episode is overwritten")
//   her_commit_kernel  after env.step: reward, done, next observation (the terminal one for envs that were auto-reset) and
//                      the episode bookkeeping: when an env's episode ends, every transition of it learns its length
//   her_sample_kernel  a batch of transitions of FINISHED episodes; the first batch / (n_sampled_goal + 1) keep their goal,
//                      the rest take the achieved goal of a uniformly drawn later step of the same episode ("future") and get
//                      their reward from the same float32 arithmetic as so100_compute_reward
#pragma once
#include "so100_dev.cuh"

namespace so100 {

struct HerRing {            // mirror of so100_her_ring (plain pointers, same order)
  int capacity, num_envs;
  float *obs, *next_obs, *achieved, *next_achieved, *desired, *action, *reward;
  uint8_t* done;
  int *ep_start, *ep_length, *cur_start, *cur_length;
};

// ---- per-env bookkeeping kernels (one thread per env) and element-wise copy kernels (one thread per float, coalesced)
__global__ void her_begin_kernel(HerRing R, int pos) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= R.num_envs) return;
  const size_t s = (size_t)pos * R.num_envs + e;
  // the slot still holds a transition of an older, finished episode: that whole episode leaves the buffer
  const int old_len = R.ep_length[s];
  if (old_len > 0) {
    const int st = R.ep_start[s];
    for (int k = 0; k < old_len; k++) R.ep_length[(size_t)((st + k) % R.capacity) * R.num_envs + e] = 0;
  }
  // a running episode longer than the ring loses its oldest step (the stored trajectory then starts one step later)
  if (R.cur_length[e] >= R.capacity) { R.cur_start[e] = (R.cur_start[e] + 1) % R.capacity; R.cur_length[e] = R.capacity - 1; }
  R.ep_start[s] = R.cur_start[e];
  R.ep_length[s] = 0;
}
// obs[15] achieved[3] desired[3] action[6] of every env into slot `pos`: 27 floats per env
__global__ void her_begin_copy_kernel(HerRing R, int pos, const float* obs, const float* achieved, const float* desired, const float* action) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long e = i / 27;
  const int k = (int)(i - e * 27);
  if (e >= R.num_envs) return;
  const size_t s = (size_t)pos * R.num_envs + e;
  if (k < 15) R.obs[s * 15 + k] = obs[e * 15 + k];
  else if (k < 18) R.achieved[s * 3 + (k - 15)] = achieved[e * 3 + (k - 15)];
  else if (k < 21) R.desired[s * 3 + (k - 18)] = desired[e * 3 + (k - 18)];
  else R.action[s * 6 + (k - 21)] = action[e * 6 + (k - 21)];
}

__global__ void her_commit_kernel(HerRing R, int pos, const float* reward, const uint8_t* terminated, const uint8_t* truncated) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= R.num_envs) return;
  const size_t s = (size_t)pos * R.num_envs + e;
  const bool term = terminated[e] != 0, over = term || truncated[e] != 0;
  R.reward[s] = reward[e];
  R.done[s] = term ? 1 : 0;            // SB3 stores dones * (1 - timeouts): a truncation is not a terminal state for the critic
  const int len = R.cur_length[e] + 1;
  if (over) {
    const int st = R.cur_start[e];
    for (int k = 0; k < len; k++) R.ep_length[(size_t)((st + k) % R.capacity) * R.num_envs + e] = len;
    R.cur_start[e] = (pos + 1) % R.capacity;
    R.cur_length[e] = 0;
  } else {
    R.cur_length[e] = len;
  }
}
// next observation: the terminal one for an env the step has already reset (SB3's infos["terminal_observation"]); the achieved
// goal is the cube site = the first three observation entries (env.py:336-339, 137-145).  18 floats per env.
__global__ void her_commit_copy_kernel(HerRing R, int pos, const float* obs, const float* achieved, const float* final_obs,
                                       const uint8_t* terminated, const uint8_t* truncated) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long e = i / 18;
  const int k = (int)(i - e * 18);
  if (e >= R.num_envs) return;
  const size_t s = (size_t)pos * R.num_envs + e;
  const bool over = terminated[e] != 0 || truncated[e] != 0;
  if (k < 15) R.next_obs[s * 15 + k] = over ? final_obs[e * 15 + k] : obs[e * 15 + k];
  else R.next_achieved[s * 3 + (k - 15)] = over ? final_obs[e * 15 + (k - 15)] : achieved[e * 3 + (k - 15)];
}

// one thread per sample: which transition, which relabelling step, reward, done.  index[b] = (ring position, env, ring position of
// the relabelling step or -1)
__global__ void her_pick_kernel(HerRing R, long long batch, int n_sampled_goal, float thr, uint32_t seed_lo, uint32_t seed_hi,
                                uint32_t call, float* reward, uint8_t* done, int* index) {
  const long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (b >= batch) return;
  const long long n_real = batch / (n_sampled_goal + 1);      // her_ratio = 1 - 1 / (n_sampled_goal + 1)
  const unsigned long long cells = (unsigned long long)R.capacity * R.num_envs;
  size_t s = 0;
  bool found = false;
  uint32_t r[4] = {0, 0, 0, 0};
  // uniform over the transitions of finished episodes: rejection sampling over the ring cells
  for (uint32_t attempt = 0; attempt < 64 && !found; attempt++) {
    philox4x32((uint32_t)b, (uint32_t)((unsigned long long)b >> 32), call, attempt, seed_lo, seed_hi, r);
    s = (size_t)((((unsigned long long)r[0] << 32) | r[1]) % cells);
    found = R.ep_length[s] > 0;
  }
  int* ix = index + b * 3;
  if (!found) {                                               // (nearly) empty buffer: flagged, never silently filled
    ix[0] = -1; ix[1] = -1; ix[2] = -1;
    reward[b] = 0.0f; done[b] = 0;
    return;
  }
  const int pos = (int)(s / R.num_envs), e = (int)(s % R.num_envs);
  float rew = R.reward[s];
  int fut = -1;
  if (b >= n_real) {
    // "future": a step of the same episode at or after this one, uniformly (SB3: randint(current, ep_length))
    const int st = R.ep_start[s], len = R.ep_length[s];
    const int k = ((pos - st) % R.capacity + R.capacity) % R.capacity;
    const int f = k + (int)(r[2] % (uint32_t)(len - k));
    fut = (st + f) % R.capacity;
    const size_t sf = (size_t)fut * R.num_envs + e;
    // env.py:346-349 in float32, ((dx^2 + dy^2) + dz^2), as compute_reward_kernel
    const float dx = __fsub_rn(R.next_achieved[s * 3], R.next_achieved[sf * 3]), dy = __fsub_rn(R.next_achieved[s * 3 + 1], R.next_achieved[sf * 3 + 1]),
                dz = __fsub_rn(R.next_achieved[s * 3 + 2], R.next_achieved[sf * 3 + 2]);
    const float d = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
    rew = d < thr ? 0.0f : -1.0f;
  }
  reward[b] = rew;
  done[b] = R.done[s];
  ix[0] = pos; ix[1] = e; ix[2] = fut;
}
// one thread per float of the batch: obs[15] next_obs[15] action[6] achieved[3] next_achieved[3] desired[3] = 45 per sample
__global__ void her_gather_kernel(HerRing R, long long batch, const int* index, float* obs, float* action, float* next_obs, float* achieved,
                                  float* next_achieved, float* desired) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long b = i / 45;
  const int k = (int)(i - b * 45);
  if (b >= batch) return;
  const int pos = index[b * 3], e = index[b * 3 + 1], fut = index[b * 3 + 2];
  if (pos < 0) return;
  const size_t s = (size_t)pos * R.num_envs + e;
  if (k < 15) obs[b * 15 + k] = R.obs[s * 15 + k];
  else if (k < 30) next_obs[b * 15 + (k - 15)] = R.next_obs[s * 15 + (k - 15)];
  else if (k < 36) action[b * 6 + (k - 30)] = R.action[s * 6 + (k - 30)];
  else if (k < 39) achieved[b * 3 + (k - 36)] = R.achieved[s * 3 + (k - 36)];
  else if (k < 42) next_achieved[b * 3 + (k - 39)] = R.next_achieved[s * 3 + (k - 39)];
  else desired[b * 3 + (k - 42)] = fut < 0 ? R.desired[s * 3 + (k - 42)] : R.next_achieved[((size_t)fut * R.num_envs + e) * 3 + (k - 42)];
}

}  // namespace so100
