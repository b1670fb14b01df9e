import sys, os, json
sys.path.insert(0, os.getcwd())
import torch
B, N = 65536, 225
dev = torch.device("cuda")
obs2 = torch.rand((B, 2, 9, 15, 15), device=dev)
obs = obs2[:, 0]
final = torch.rand((B, 9, 15, 15), device=dev)
ring = torch.zeros((8 * B, 9, 15, 15), device=dev)
done = torch.rand(B, device=dev) < 0.002
def timed(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
out = {}
out["copy_strided_to_ring"] = timed(lambda: ring[B:2*B].copy_(obs))
out["copy_contig_to_ring"] = timed(lambda: ring[B:2*B].copy_(final))
out["where_out"] = timed(lambda: torch.where(done.view(-1,1,1,1), final, obs, out=ring[2*B:3*B]))
out["where_alloc"] = timed(lambda: torch.where(done.view(-1,1,1,1), final, obs))
flat_ring = ring.view(8*B, -1)
def fix():
    ring[3*B:4*B].copy_(obs)
    idx = torch.nonzero_static(done, size=1024, fill_value=0).squeeze(1)
    flat_ring[3*B:4*B].index_copy_(0, idx, final.view(B,-1)[idx])
out["copy_plus_fix1024"] = timed(fix)
print(json.dumps(out))
