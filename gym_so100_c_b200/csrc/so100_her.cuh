// Device-resident hindsight-experience-replay ring for the GoalEnv rollout (the consumer contract of
// scripts/train_sac_her.py:231-246: SB3 HerReplayBuffer(n_sampled_goal = 4, goal_selection_strategy = "future") around
// SO100GoalEnv.compute_reward, env.py:341-353).  The ring is [capacity, num_envs] transitions in caller-owned device arrays
// (so100_her_ring, include/so100_b200.h); three kernels keep it:
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

__global__ void her_begin_kernel(HerRing R, int pos, const float* obs, const float* achieved, const float* desired, const float* action) {
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
  for (int k = 0; k < 15; k++) R.obs[s * 15 + k] = obs[(size_t)e * 15 + k];
  for (int k = 0; k < 3; k++) { R.achieved[s * 3 + k] = achieved[(size_t)e * 3 + k]; R.desired[s * 3 + k] = desired[(size_t)e * 3 + k]; }
  for (int k = 0; k < 6; k++) R.action[s * 6 + k] = action[(size_t)e * 6 + k];
  R.ep_start[s] = R.cur_start[e];
  R.ep_length[s] = 0;
}

__global__ void her_commit_kernel(HerRing R, int pos, const float* obs, const float* achieved, const float* final_obs, const float* reward,
                                  const uint8_t* terminated, const uint8_t* truncated) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= R.num_envs) return;
  const size_t s = (size_t)pos * R.num_envs + e;
  const bool term = terminated[e] != 0, over = term || truncated[e] != 0;
  // next observation: the terminal one for an env the step has already reset (SB3's infos["terminal_observation"]);
  // the achieved goal is the cube site = the first three observation entries (env.py:336-339, 137-145)
  const float* nx = over ? final_obs + (size_t)e * 15 : obs + (size_t)e * 15;
  for (int k = 0; k < 15; k++) R.next_obs[s * 15 + k] = nx[k];
  for (int k = 0; k < 3; k++) R.next_achieved[s * 3 + k] = over ? nx[k] : achieved[(size_t)e * 3 + k];
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

// one thread per sample.  index[b] = (ring position, env, ring position of the relabelling step or -1)
__global__ void her_sample_kernel(HerRing R, long long batch, int n_sampled_goal, float thr, uint32_t seed_lo, uint32_t seed_hi,
                                  uint32_t call, float* obs, float* action, float* next_obs, float* achieved, float* next_achieved,
                                  float* desired, float* reward, uint8_t* done, int* index) {
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
  float goal[3] = {R.desired[s * 3], R.desired[s * 3 + 1], R.desired[s * 3 + 2]};
  float rew = R.reward[s];
  int fut = -1;
  if (b >= n_real) {
    // "future": a step of the same episode at or after this one, uniformly (SB3: randint(current, ep_length))
    const int st = R.ep_start[s], len = R.ep_length[s];
    const int k = ((pos - st) % R.capacity + R.capacity) % R.capacity;
    const int f = k + (int)(r[2] % (uint32_t)(len - k));
    fut = (st + f) % R.capacity;
    const size_t sf = (size_t)fut * R.num_envs + e;
    for (int c = 0; c < 3; c++) goal[c] = R.next_achieved[sf * 3 + c];
    // env.py:346-349 in float32, ((dx^2 + dy^2) + dz^2), as compute_reward_kernel
    const float dx = __fsub_rn(R.next_achieved[s * 3], goal[0]), dy = __fsub_rn(R.next_achieved[s * 3 + 1], goal[1]),
                dz = __fsub_rn(R.next_achieved[s * 3 + 2], goal[2]);
    const float d = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
    rew = d < thr ? 0.0f : -1.0f;
  }
  for (int k = 0; k < 15; k++) { obs[b * 15 + k] = R.obs[s * 15 + k]; next_obs[b * 15 + k] = R.next_obs[s * 15 + k]; }
  for (int k = 0; k < 6; k++) action[b * 6 + k] = R.action[s * 6 + k];
  for (int k = 0; k < 3; k++) {
    achieved[b * 3 + k] = R.achieved[s * 3 + k]; next_achieved[b * 3 + k] = R.next_achieved[s * 3 + k]; desired[b * 3 + k] = goal[k];
  }
  reward[b] = rew;
  done[b] = R.done[s];
  ix[0] = pos; ix[1] = e; ix[2] = fut;
}

}  // namespace so100
