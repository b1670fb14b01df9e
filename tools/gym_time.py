"""Where a GeneralsVecEnv.step() goes at 65,536 envs: CUDA-event time of the whole step vs wall time."""
import sys, time; sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import torch
from generalsreinforcementlearning_b200.gym_env import GeneralsVecEnv
W = int(sys.argv[1]) if len(sys.argv) > 1 else 15
env = GeneralsVecEnv(65536, W, W, max_turns=500, seed=3)
obs, info = env.reset()
acts = [env.sample_actions() for _ in range(8)]
for i in range(5): env.step(acts[i % 8])
torch.cuda.synchronize()
# (a) back-to-back steps, one sync at the end (what the device needs)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); e0.record()
for i in range(40): env.step(acts[i % 8])
e1.record(); torch.cuda.synchronize(); t1 = time.perf_counter()
print("40 steps: device %.3f ms/step, wall %.3f ms/step" % (e0.elapsed_time(e1) / 40, (t1 - t0) * 1e3 / 40))
# (b) the library call alone
t0 = time.perf_counter()
for i in range(40):
    o = env._out[0]
    env.engine.gym_step(env.max_turns, 5 + i, action=acts[i % 8], opponent_action=None, obs=env._obs, mask=env._mask, stats=env._stats,
                        actions=env._actions, prev_stats=env._prev_stats, turns=env._turns, calls=env._calls, reward=o["reward"],
                        terminated=o["terminated"], truncated=o["truncated"], valid=o["valid"], done=env._done, winner=o["winner"],
                        step_error=o["step_error"], n_finished=env._nfin)
torch.cuda.synchronize(); t1 = time.perf_counter()
print("40 grl_gym_step calls: wall %.3f ms/step" % ((t1 - t0) * 1e3 / 40))
t0 = time.perf_counter()
for i in range(40):
    torch.cuda.synchronize()
t1 = time.perf_counter()
print("empty sync %.4f ms" % ((t1 - t0) * 1e3 / 40))
