import time, sys, numpy as np
sys.path.insert(0,'/root/repo')
import torch
from generalsreinforcementlearning_b200 import load_library
from generalsreinforcementlearning_b200.engine import BatchedEngine, make_config
lib=load_library()
for (W,H,B) in [(20,20,65536),(10,10,65536)]:
    e=BatchedEngine(lib, make_config(lib,num_envs=B,width=W,height=H,num_players=2,host_threads=0))
    seeds=np.arange(B,dtype=np.int64)+1
    e.reset_seeded(seeds)
    t0=time.perf_counter(); e.reset_seeded(seeds+B); dt=time.perf_counter()-t0
    ids=np.arange(0,B,64,dtype=np.int32)
    t0=time.perf_counter(); e.reset_seeded(seeds[:len(ids)]+3*B, ids); dt2=time.perf_counter()-t0
    import os
    print(W,H,B,'full reset s',round(dt,3),'per map us (wall)',round(dt/B*1e6,2),'partial',len(ids),'maps s',round(dt2,4), 'cores', os.cpu_count())
    e.close()
