"""N>1 on real GPUs (skipped below 2 devices): game shards on separate B200s reproduce the
single-device run, and the learner hand-off (NCCL gather of experience records + all-reduce
of statistics) delivers exactly the records of the unsharded run.  The CPU twin of this test
is tests/test_sharding_gloo.py."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _worker(rank, world, port, total, T, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    from generalsreinforcementlearning_b200 import _abi, load_library, sharding

    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    lib = load_library()
    sh = sharding.shard_from_env(total)
    e = sharding.create_sharded_engine(lib, sh, device=rank, width=20, height=20, num_players=2, host_threads=0)
    e.use_torch_stream()
    e.reset_seeded(sh.seeds(12345))
    B, P, H, W = sh.count, 2, 20, 20
    mk = lambda *shape, dt=torch.float32: torch.zeros(shape, dtype=dt, device=dev)
    obs, prev = mk(B, P, 9, H, W), mk(B, P, 9, H, W)
    mask, pmask = mk(B, P, e.mask_words, dt=torch.int32), mk(B, P, e.mask_words, dt=torch.int32)
    reward, done, aidx = mk(B, P), mk(B, dt=torch.uint8), mk(B, P, dt=torch.int32)
    e.observe(e.outputs(obs=prev, mask_bits=pmask))
    hashes, counts = [], []
    for t in range(T):
        e.step_fused(None, e.outputs(obs=obs, mask_bits=mask, reward=reward, done=done, action_index=aidx),
                     _abi.STEP_FLAG_RANDOM_POLICY, 99)
        rec = sharding.pack_experience(prev, obs, pmask, aidx, reward, done, t + 1, env_id_base=sh.first)
        g = sharding.gather_experience(rec, capacity=2 * B, dst=0)
        if rank == 0:
            counts.append(int(g["action"].shape[0]))
            if t == T - 1:
                np.savez(os.path.join(tmp, "last.npz"), **{k: v.cpu().numpy() for k, v in g.items() if k not in ("state", "next_state")},
                         next_state_sum=g["next_state"].sum(dim=(1, 2, 3)).cpu().numpy())
        hashes.append(e.state_hash().copy())
        prev, obs = obs, prev
        pmask, mask = mask, pmask
    stats = sharding.all_reduce_stats(e.stats(), device=dev)
    np.savez(os.path.join(tmp, f"rank{rank}.npz"), hashes=np.stack(hashes), stats=stats, counts=np.array(counts))
    dist.barrier()
    dist.destroy_process_group()


def test_two_gpu_shards_and_nccl_gather(cuda_lib, tmp_path):
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from generalsreinforcementlearning_b200 import _abi
    from helpers import new_engine

    total, T, world = 4096, 20, 2
    port = 29600 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, total, T, str(tmp_path)), nprocs=world, join=True)
    e = new_engine(cuda_lib, 20, 20, 2, total, host_threads=0)
    e.reset_seeded(np.arange(total, dtype=np.int64) + 12345)
    out = e.alloc_outputs_host()
    r0, r1 = (np.load(os.path.join(tmp_path, f"rank{r}.npz")) for r in range(2))
    for t in range(T):
        e.step_fused(None, e.outputs(**out), _abi.STEP_FLAG_RANDOM_POLICY, 99)
        assert np.array_equal(np.concatenate([r0["hashes"][t], r1["hashes"][t]]), e.state_hash()), f"turn {t}"
        assert r0["counts"][t] == int((out["action_index"] >= 0).sum())
    last = np.load(os.path.join(tmp_path, "last.npz"))
    env, ply = np.nonzero(out["action_index"] >= 0)
    assert np.array_equal(last["env_id"], env.astype(np.int32)) and np.array_equal(last["player"], ply.astype(np.int32))
    assert np.array_equal(last["action"], out["action_index"][env, ply])
    assert np.array_equal(last["reward"].view(np.uint32), out["reward"][env, ply].view(np.uint32))
    assert np.allclose(last["next_state_sum"], out["obs"][env, ply].sum(axis=(1, 2, 3)))
    assert np.array_equal(r0["stats"], r1["stats"]) and np.array_equal(r0["stats"], e.stats().astype(np.int64))


def _selfplay_worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import selfplay_experience

    selfplay_experience.run(rank, world, games_per_gpu=1024, turns=6, sample=40,
                            out_path=os.path.join(tmp, "selfplay.json") if rank == 0 else None, start_turn=5)
    import torch.distributed as dist

    if dist.is_initialized():
        dist.destroy_process_group()


def test_selfplay_experience_to_grpc_learner(cuda_lib, tmp_path):
    """BASELINE config 5 in small: self-play shards -> NCCL gather -> gRPC batches of 32 on the learner rank."""
    import json

    import torch
    import torch.multiprocessing as mp

    world = min(torch.cuda.device_count(), 2)
    port = 29700 + (os.getpid() % 2000)
    mp.spawn(_selfplay_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r = json.load(open(os.path.join(tmp_path, "selfplay.json")))
    phases = 5 if world > 1 and "unavailable" not in r["p2p_all"] else 3
    assert r["env_steps_total"] == world * 1024 * phases * (5 + 1 + 6)     # every phase rewinds to turn 5, warms one turn, times six
    ga, st = r["gather_all"], r["stream"]
    # every transition of every turn reached the learner's HBM, nothing dropped, nothing padded
    assert ga["dropped"] == 0 and 1.9 * 1024 * world <= ga["experiences_per_turn"] <= 2 * 1024 * world
    assert ga["nvlink_bytes_per_turn"] == (ga["experiences_per_turn"] * (world - 1) / world) * r["record_bytes"] or world == 1 or \
        abs(ga["nvlink_bytes_per_turn"] - ga["experiences_per_turn"] * (world - 1) / world * r["record_bytes"]) < 0.1 * ga["nvlink_bytes_per_turn"]
    if world > 1:   # the turn kernel's stores into the learner's planes over NVLink peer memory delivered the very records
        assert r["p2p_all"].get("verified") is True, r["p2p_all"]
    assert st["gathered"] == world * 40 * 6          # the per-turn sample of every rank arrived
    assert st["streamed"] == st["gathered"]          # and all of it went out through the ExperienceService
    assert st["full_batches"] >= st["gathered"] // 32 - 2
