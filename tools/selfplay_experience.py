#!/usr/bin/env python3
"""BASELINE config 5: 2-player 20x20 self-play on every GPU of the box, experience gathered to the
learner rank over NCCL and served from there through the gRPC ExperienceService in batches of 32.

  torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/selfplay_experience.py \
      [--games-per-gpu 4096] [--turns 20] [--sample 64]

Every rank steps its shard with the in-kernel random policy (device-resident observations).  Each
turn a fixed-size sample of that turn's transitions (``--sample`` per rank) is packed on the device
(state, action, reward, next state, done, serializer mask), gathered to rank 0 with NCCL
(``sharding.gather_experience``), and fed into the service's store; a gRPC client on rank 0 drains
``StreamExperienceBatches(batch_size=32)``.  Rank 0 prints one JSON line."""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run(rank, world, games_per_gpu=4096, turns=20, sample=64, out_path=None, port_base=None):
    import grpc
    import numpy as np
    import torch
    import torch.distributed as dist

    from generalsreinforcementlearning_b200 import _abi, load_library, sharding
    from generalsreinforcementlearning_b200.grpc_schema import experience
    from generalsreinforcementlearning_b200.grpc_service import GameServer, Stub

    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    lib = load_library()
    W = H = 20
    P = 2
    sh = sharding.shard_for(games_per_gpu * world, world, rank)
    e = sharding.create_sharded_engine(lib, sh, device=local, width=W, height=H, num_players=P, host_threads=0)
    e.use_torch_stream()
    e.reset_seeded(sh.seeds(12345))
    B = sh.count
    mk = lambda *shape, dt=torch.float32: torch.zeros(shape, dtype=dt, device=dev)  # noqa: E731
    obs, prev = mk(B, P, 9, H, W), mk(B, P, 9, H, W)
    smask, pmask = mk(B, P, 4 * W * H, dt=torch.uint8), mk(B, P, 4 * W * H, dt=torch.uint8)
    reward, done, aidx = mk(B, P), mk(B, dt=torch.uint8), mk(B, P, dt=torch.int32)
    e.observe(e.outputs(obs=prev))
    e.mask(_abi.MASK_SERIALIZER_UDLR, pmask)

    server = gs = client = None
    streamed = []
    if rank == 0:
        from concurrent import futures

        gs = GameServer(lib=lib)
        server = grpc.server(futures.ThreadPoolExecutor(max_workers=4))
        gs.add_to_server(server)
        port = server.add_insecure_port("127.0.0.1:0")
        server.start()
        stub = Stub(grpc.insecure_channel(f"127.0.0.1:{port}"), "generals.experience.v1.ExperienceService")
        stop = threading.Event()

        def drain():
            call = stub.StreamExperienceBatches(experience.StreamExperiencesRequest(batch_size=32, follow=True))
            try:
                for b in call:
                    streamed.append(len(b.experiences))
                    if stop.is_set() and sum(streamed) >= drain.expect:
                        call.cancel()
                        return
            except grpc.RpcError:
                pass

        drain.expect = 1 << 60
        client = threading.Thread(target=drain, daemon=True)
        client.start()

    gathered = 0
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for t in range(turns):
        e.step_fused(None, e.outputs(obs=obs, reward=reward, done=done, action_index=aidx), _abi.STEP_FLAG_RANDOM_POLICY, 7)
        e.mask(_abi.MASK_SERIALIZER_UDLR, smask)
        rec = sharding.pack_experience(prev, obs, pmask, aidx, reward, done, t + 1, env_id_base=sh.first)
        rec = {k: v[:sample] for k, v in rec.items()}  # a fixed-size sample of this turn's transitions
        g = sharding.gather_experience(rec, capacity=sample, dst=0) if world > 1 else rec
        if rank == 0:
            gathered += gs.ingest_records(g, W, H)
        prev, obs = obs, prev
        pmask, smask = smask, pmask
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    stats = sharding.all_reduce_stats(e.stats(), device=dev)
    result = None
    if rank == 0:
        drain.expect = gathered
        stop.set()
        client.join(timeout=20)
        server.stop(0)
        result = dict(n_gpus=world, games_per_gpu=games_per_gpu, turns=turns, env_steps=int(stats[0]),
                      env_steps_per_s=float(stats[0]) / dt, gathered=gathered, streamed=int(sum(streamed)),
                      batches=len(streamed), full_batches=int(sum(1 for n in streamed if n == 32)),
                      note="step + observation + NCCL gather of a per-turn sample + proto conversion on the learner rank")
        if out_path:
            json.dump(result, open(out_path, "w"))
        gs.close()
    e.close()
    if world > 1:
        dist.barrier()
    return result


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--games-per-gpu", type=int, default=4096)
    ap.add_argument("--turns", type=int, default=20)
    ap.add_argument("--sample", type=int, default=64)
    a = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    r = run(rank, world, a.games_per_gpu, a.turns, a.sample)
    if rank == 0:
        print(json.dumps(r))
    if world > 1:
        import torch.distributed as dist

        dist.destroy_process_group()


if __name__ == "__main__":
    main()
