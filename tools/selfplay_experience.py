#!/usr/bin/env python3
"""BASELINE config 5: 2-player 20x20 self-play on every GPU of the box, experience gathered to the
learner rank over NCCL and served from there through the gRPC ExperienceService in batches of 32.

  torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/selfplay_experience.py \
      [--games-per-gpu 65536] [--turns 40] [--sample 256] [--out profiles/r2_selfplay.json]

Every rank steps its shard with the in-kernel random policy.  What leaves the step kernel for the hand-off is the
PACKED observation record of each game (grl_step_outputs.obs_packed, 1,120 B at 20x20) before and after the turn, so an
experience record is 2,264 B instead of the 29 KB of two float32 tensors; the learner rebuilds state, next_state and the
serializer action mask on its side (sharding.expand_experience).  Records travel as one byte buffer per rank with its
actual row count (sharding.gather_experience: one count all-gather + one variable-size transfer per rank, no padding).

Three measured phases, each `--turns` turns from the same mid-episode state (turn 100), timed on every rank between a
barrier + cuda synchronize on both sides, max over ranks:
  step_only     the turn kernel alone (the step path without any hand-off)
  gather_all    + EVERY transition of every turn packed and gathered into the learner GPU's HBM
  stream        + a per-turn sample of `--sample` records per rank gathered, expanded on the learner's host, framed as
                Experience messages and drained by a gRPC client through StreamExperienceBatches(batch_size=32)
Rank 0 prints one JSON line (and writes --out)."""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run(rank, world, games_per_gpu=4096, turns=20, sample=64, out_path=None, start_turn=100):
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
    B = sh.count
    RW = e.packed_words
    mk = lambda *shape, dt=torch.float32: torch.zeros(shape, dtype=dt, device=dev)  # noqa: E731
    pk = [mk(B, RW, dt=torch.int32), mk(B, RW, dt=torch.int32)]          # packed records before / after the turn
    reward, done, aidx = mk(B, P), mk(B, dt=torch.uint8), mk(B, P, dt=torch.int32)
    record_bytes = 2 * RW * 4 + 4 + 4 + 1 + 4 + 4 + 4

    def rewind():
        e.reset_seeded(sh.seeds(12345))
        for _ in range(start_turn):
            e.step(None, _abi.STEP_FLAG_RANDOM_POLICY, 7)
        e.observe(e.outputs(obs_packed=pk[0]))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(v):
        if world > 1:
            t = torch.tensor([v], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return v

    def step(cur):
        e.step_fused(None, e.outputs(obs_packed=pk[cur ^ 1], reward=reward, done=done, action_index=aidx),
                     _abi.STEP_FLAG_RANDOM_POLICY, 7)

    def phase(body):
        rewind()
        body(0, 0, warm=True)
        barrier()
        t0 = time.perf_counter()
        cur = 1
        for t in range(1, turns + 1):
            body(t, cur, warm=False)
            cur ^= 1
        barrier()
        return reduce_max(time.perf_counter() - t0)

    # ---- phase 1: the step path alone ---------------------------------------------------------------------------------
    dt_step = phase(lambda t, cur, warm: step(cur))

    # ---- phase 2: every transition of every turn into the learner GPU's HBM ---------------------------------------------
    got = {"records": 0, "nvlink_bytes": 0, "dropped": 0}

    def gather_all(t, cur, warm):
        step(cur)
        rec = sharding.pack_experience_packed(pk[cur], pk[cur ^ 1], aidx, reward, done, start_turn + t + 1, env_id_base=sh.first)
        g = sharding.gather_experience(rec, dst=0)
        if rank == 0 and not warm:
            counts = g["counts"].tolist()
            got["records"] += sum(counts)
            got["nvlink_bytes"] += (sum(counts) - counts[0]) * record_bytes
            got["dropped"] += int(g["dropped"].sum())

    dt_all = phase(gather_all)

    # ---- phase 3: a per-turn sample through the learner's host into the gRPC stream ---------------------------------------
    server = gs = client = None
    streamed = []
    if rank == 0:
        from concurrent import futures

        gs = GameServer(lib=lib)
        server = grpc.server(futures.ThreadPoolExecutor(max_workers=4), options=[("grpc.max_send_message_length", 64 << 20)])
        gs.add_to_server(server)
        port = server.add_insecure_port("127.0.0.1:0")
        server.start()
        channel = grpc.insecure_channel(f"127.0.0.1:{port}", options=[("grpc.max_receive_message_length", 64 << 20)])
        stub = Stub(channel, "generals.experience.v1.ExperienceService")
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
    ingested = {"n": 0}

    def stream(t, cur, warm):
        step(cur)
        rec = sharding.pack_experience_packed(pk[cur], pk[cur ^ 1], aidx, reward, done, start_turn + t + 1, env_id_base=sh.first,
                                              limit=sample)
        g = sharding.gather_experience(rec, dst=0)
        if rank == 0 and not warm:
            ingested["n"] += gs.ingest_records(g, W, H, players=P)

    dt_stream = phase(stream)
    stats = sharding.all_reduce_stats(e.stats(), device=dev)
    result = None
    if rank == 0:
        drain.expect = ingested["n"]
        stop.set()
        t_wait = time.perf_counter()
        client.join(timeout=60)
        drain_tail = time.perf_counter() - t_wait
        server.stop(0)
        env_steps = B * world * turns
        result = dict(
            config="BASELINE configs[4]: 2-player 20x20 self-play -> NCCL gather to the learner rank -> gRPC batch 32",
            n_gpus=world, games_per_gpu=games_per_gpu, turns_per_phase=turns, start_turn=start_turn,
            record_bytes=record_bytes, fp32_record_bytes=2 * 9 * W * H * 4 + 4 * W * H // 8 + 21,
            step_only=dict(ms_per_turn=1e3 * dt_step / turns, env_steps_per_s=env_steps / dt_step),
            gather_all=dict(ms_per_turn=1e3 * dt_all / turns, env_steps_per_s=env_steps / dt_all,
                            experiences_per_s_into_learner=got["records"] / dt_all, experiences_per_turn=got["records"] / turns,
                            nvlink_bytes_per_turn=got["nvlink_bytes"] / turns, nvlink_GBps_into_learner=got["nvlink_bytes"] / dt_all / 1e9,
                            dropped=got["dropped"], step_path_slowdown=dt_all / dt_step),
            stream=dict(sample_per_rank_per_turn=sample, ms_per_turn=1e3 * dt_stream / turns, env_steps_per_s=env_steps / dt_stream,
                        gathered=ingested["n"], streamed=int(sum(streamed)), batches=len(streamed),
                        full_batches=int(sum(1 for n in streamed if n == 32)),
                        experiences_per_s=ingested["n"] / dt_stream, grpc_batches_per_s=len(streamed) / (dt_stream + drain_tail),
                        nvlink_bytes_per_turn=(world - 1) * sample * record_bytes, step_path_slowdown=dt_stream / dt_step),
            env_steps_total=int(stats[0]),
            note="wall clock between barrier + cuda synchronize on both sides, max over ranks; the stream phase includes the "
                 "learner's host work (expansion to float32 tensors, protobuf framing) on rank 0's Python thread")
        if out_path:
            json.dump(result, open(out_path, "w"), indent=1)
        gs.close()
    e.close()
    if world > 1:
        dist.barrier()
    return result


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--games-per-gpu", type=int, default=65536)
    ap.add_argument("--turns", type=int, default=40)
    ap.add_argument("--sample", type=int, default=256)
    ap.add_argument("--start-turn", type=int, default=100)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    r = run(rank, world, a.games_per_gpu, a.turns, a.sample, a.out, a.start_turn)
    if rank == 0:
        print(json.dumps(r))
    if world > 1:
        import torch.distributed as dist

        dist.destroy_process_group()


if __name__ == "__main__":
    main()
