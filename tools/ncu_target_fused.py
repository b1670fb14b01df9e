"""A few fused steps of one shape for an ncu capture: python tools/ncu_target_fused.py W H P B"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from generalsreinforcementlearning_b200 import load_library, _abi
from generalsreinforcementlearning_b200.engine import BatchedEngine, make_config
W, H, P, B = (int(v) for v in sys.argv[1:5])
lib = load_library(); dev = torch.device("cuda:0")
e = BatchedEngine(lib, make_config(lib, num_envs=B, width=W, height=H, num_players=P, host_threads=0))
e.use_torch_stream()
e.reset_seeded(np.arange(B, dtype=np.int64) + 12345)
obs = torch.empty((B, P, 9, H, W), dtype=torch.float32, device=dev)
mask = torch.empty((B, P, e.mask_words), dtype=torch.int32, device=dev)
reward = torch.empty((B, P), dtype=torch.float32, device=dev); done = torch.empty(B, dtype=torch.uint8, device=dev)
for t in range(40):
    e.step_fused(None, e.outputs(obs=obs, mask_bits=mask, reward=reward, done=done), _abi.STEP_FLAG_RANDOM_POLICY, 7)
torch.cuda.synchronize(); print("ok")
