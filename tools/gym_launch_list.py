import sys, os
sys.path.insert(0, os.getcwd())
import torch
from generalsreinforcementlearning_b200.gym_env import GeneralsVecEnv
B, W = 65536, 15
env = GeneralsVecEnv(B, W, W, max_turns=500, seed=3, auto_reset="device")
env.reset()
env._calls.copy_(torch.randint(0, env.max_turns, (B,), device=env._calls.device, dtype=torch.int32))
for _ in range(12):
    env.step(env.sample_actions())
torch.cuda.synchronize()
