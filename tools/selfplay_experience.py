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

Four measured phases, each `--turns` turns from the same mid-episode state (turn 100), timed on every rank between a
barrier + cuda synchronize on both sides, max over ranks:
  step_only     the turn kernel alone (the step path without any hand-off)
  p2p_all       EVERY transition of every turn lands in the learner GPU's HBM with NO extra kernel and NO collective: the
                turn kernel's own output pointers (obs_packed, reward, done, action_index) are this rank's rows of planes
                that live on the learner GPU, mapped into every rank through NVLink peer memory
                (torch.distributed._symmetric_memory), so the kernel's stores ARE the transfer; one device-side barrier
                per turn tells the learner the turn has landed.  (+ the learner compacting the turn into records)
  gather_all    the same hand-off through NCCL: records packed on each rank and gathered with send/recv
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

    # ---- phase 2: the turn kernel writes straight into the learner's HBM over NVLink peer memory --------------------------
    p2p = None
    if world > 1:
        try:
            import torch.distributed._symmetric_memory as symm_mem

            total = sh.total_envs
            n_pk, n_small = 2 * total * RW, total * P
            words = n_pk + 2 * n_small + (total + 3) // 4
            buf = symm_mem.empty(words, dtype=torch.int32, device=dev)
            hdl = symm_mem.rendezvous(buf, dist.group.WORLD)
            learner = hdl.get_buffer(0, (words,), torch.int32, 0)            # rank 0's planes, as seen from this rank
            L_pk = learner[:n_pk].view(2, total, RW)
            L_reward = learner[n_pk:n_pk + n_small].view(torch.float32).view(total, P)
            L_aidx = learner[n_pk + n_small:n_pk + 2 * n_small].view(total, P)
            L_done = learner[n_pk + 2 * n_small:].view(torch.uint8)[:total]
            rows = slice(sh.first, sh.first + B)
            landed = {"records": 0}

            def p2p_body(compact):
                def body(t, cur, warm):
                    if warm:
                        e.observe(e.outputs(obs_packed=L_pk[0, rows]))
                    e.step_fused(None, e.outputs(obs_packed=L_pk[cur ^ 1, rows], reward=L_reward[rows], done=L_done[rows],
                                                 action_index=L_aidx[rows]), _abi.STEP_FLAG_RANDOM_POLICY, 7)
                    hdl.barrier()                                           # device-side: every rank's turn has landed
                    if rank == 0 and compact:                               # the learner forms this turn's records
                        rec = sharding.pack_experience_packed(L_pk[cur], L_pk[cur ^ 1], L_aidx, L_reward, L_done, start_turn + t + 1)
                        if not warm:
                            landed["records"] += int(rec["action"].shape[0])
                    if compact:
                        hdl.barrier()                                       # the planes may be overwritten again
                return body

            dt_p2p = phase(p2p_body(False))
            # what landed in the learner's planes is what this rank's kernel would have written at home
            final = (turns + 1) & 1                                          # parity of the last turn's "after" plane
            e.observe(e.outputs(obs_packed=pk[0], done=done))
            torch.cuda.synchronize()
            same = torch.equal(pk[0], L_pk[final, rows]) and torch.equal(done, L_done[rows])
            ok = torch.tensor([1 if same else 0], dtype=torch.int32, device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            dt_p2p_c = phase(p2p_body(True))
            per_turn = (world - 1) * B * (RW * 4 + P * 8 + 1)
            p2p = dict(ms_per_turn=1e3 * dt_p2p / turns, env_steps_per_s=B * world * turns / dt_p2p,
                       experiences_per_s_into_learner=2 * B * world * turns / dt_p2p, nvlink_bytes_per_turn=per_turn,
                       nvlink_GBps_into_learner=per_turn * turns / dt_p2p / 1e9, step_path_slowdown=dt_p2p / dt_step,
                       verified=bool(ok.item()),
                       with_learner_compaction=dict(ms_per_turn=1e3 * dt_p2p_c / turns, experiences_per_s=landed["records"] / dt_p2p_c if rank == 0 else None,
                                                    step_path_slowdown=dt_p2p_c / dt_step),
                       how="turn kernel stores into learner-resident planes through NVLink peer memory "
                           "(torch.distributed._symmetric_memory); one device-side barrier per turn; no NCCL call, no copy kernel")
            del learner, L_pk, L_reward, L_aidx, L_done, hdl, buf
        except Exception as exc:  # noqa: BLE001 - reported, the NCCL path below still runs
            p2p = {"unavailable": repr(exc)[:300]}

    # ---- phase 3: the same hand-off through NCCL: records packed per rank, gathered with send/recv --------------------------
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

    # ---- phase 4: a per-turn sample through the learner's host into the gRPC stream ---------------------------------------
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
            p2p_all=p2p,
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
