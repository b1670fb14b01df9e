import sys; sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from generalsreinforcementlearning_b200.gym_env import GeneralsVecEnv
W = int(sys.argv[1]) if len(sys.argv) > 1 else 15
env = GeneralsVecEnv(65536, W, W, max_turns=500, seed=3, auto_reset="device")
env.reset()
env._calls.copy_(torch.randint(0, 500, (65536,), device=env._calls.device, dtype=torch.int32))
for _ in range(5): env.step(env.sample_actions())
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(10):
        env.step(env.sample_actions())
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=16, max_name_column_width=50))
