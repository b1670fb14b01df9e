import sys; sys.path.insert(0,'/root/repo')
import torch
from torch.profiler import profile, ProfilerActivity
from generalsreinforcementlearning_b200.gym_env import GeneralsVecEnv
env=GeneralsVecEnv(65536,int(sys.argv[1]) if len(sys.argv)>1 else 15,int(sys.argv[1]) if len(sys.argv)>1 else 15,max_turns=500,seed=3)
obs,info=env.reset()
for _ in range(5): obs,r,te,tr,info=env.step(env.sample_actions())
acts=[env.sample_actions() for _ in range(1)]
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(10):
        obs,r,te,tr,info=env.step(acts[0])
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=60))
