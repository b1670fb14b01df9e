"""A few fused gym steps (grl_gym_step) at 65,536 envs for an ncu capture: python tools/gym_ncu_target.py <board> [envs]
GRL_GYM_AGENT=in_step: the random agent is drawn inside the step (GeneralsVecEnv.step(None), the GYM = 2 instantiation)."""
import sys; sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import torch
from generalsreinforcementlearning_b200.gym_env import GeneralsVecEnv
W = int(sys.argv[1]) if len(sys.argv) > 1 else 20
B = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
env = GeneralsVecEnv(B, W, W, max_turns=500, seed=3)
obs, info = env.reset()
for _ in range(40):
    obs, r, te, tr, info = env.step(None if __import__("os").environ.get("GRL_GYM_AGENT") == "in_step" else env.sample_actions())
torch.cuda.synchronize()
print("ok", float(r.sum()))
