import sys, torch
sys.path.insert(0, '.')
from gym_so100_c_b200.her import HerRollout
from gym_so100_c_b200.vec_env import SO100GoalVecEnv
n = 65536
env = SO100GoalVecEnv(n, device="cuda:0", seed=0x50100)
roll = HerRollout(env, horizon=320, n_sampled_goal=4)
roll.reset()
env.sim.set_aux(step_count=torch.randint(0, 300, (n,), dtype=torch.int32))
g = torch.Generator(device="cuda").manual_seed(1)
for s in range(310):
    roll.step(torch.rand((n, 6), device="cuda", generator=g) * 2 - 1)
acts = torch.rand((10, n, 6), device="cuda", generator=g) * 2 - 1
def timeit(f, k=10):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(k): f(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k
print("roll.step        %.3f ms" % timeit(lambda i: roll.step(acts[i])))

print("roll.sample(5n)  %.3f ms" % timeit(lambda i: roll.sample(5 * n)))
print("roll.sample(256) %.3f ms" % timeit(lambda i: roll.sample(256)))
for rep in range(3):
    print("roll.sample(5n)  %.3f ms" % timeit(lambda i: roll.sample(5 * n)))
