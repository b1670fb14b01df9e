import sys, time; sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import torch
from generalsreinforcementlearning_b200.gym_env import GeneralsVecEnv
env = GeneralsVecEnv(65536, 15, 15, max_turns=500, seed=3)
obs, info = env.reset()
acts = [env.sample_actions() for _ in range(8)]
for i in range(5): env.step(acts[i % 8])
torch.cuda.synchronize()
nf = []
for i in range(10):
    t0 = time.perf_counter()
    out = env.step(acts[i % 8]); torch.cuda.synchronize()
    t1 = time.perf_counter()
    nf.append((int(env._nfin.item()), round((t1 - t0) * 1e3, 3), int(out[4]["invalid_action"].sum())))
print(nf)
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for i in range(20): env.step(acts[i % 8])
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
