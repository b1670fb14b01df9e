"""World-size-2 gloo test of the multi-rank host logic (no GPU): the env->rank partition,
seed/policy keying through env_id_base, the statistics all-reduce and the experience gather.
Each rank steps its shard with the CPU oracle (tests may use it as the checker); the
concatenation must equal one unsharded run game for game."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

from generalsreinforcementlearning_b200 import sharding  # noqa: E402


def test_partition_covers_every_env_once():
    for total, world in [(65536, 8), (10, 3), (7, 7), (262144, 4), (1001, 8)]:
        seen = []
        for r in range(world):
            sh = sharding.shard_for(total, world, r)
            seen.extend(range(sh.first, sh.first + sh.count))
            for e in (sh.first, sh.first + sh.count - 1):
                assert sharding.owner_rank(total, world, e) == r
                assert sh.local(e) == e - sh.first
        assert seen == list(range(total))
    sh = sharding.shard_for(65536, 8, 3)
    assert (sh.first, sh.count) == (3 * 8192, 8192)  # gpu = env / (B_total / G), SURVEY 8e
    assert sh.seeds(12345)[0] == 12345 + 3 * 8192 and sh.seeds(12345, episode=1)[0] == 12345 + 3 * 8192 + 65536
    with pytest.raises(IndexError):
        sh.local(0)
    with pytest.raises(ValueError):
        sharding.shard_for(4, 8, 0)


def _worker(rank, world, port, total, T, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist

    from generalsreinforcementlearning_b200 import _abi
    from generalsreinforcementlearning_b200._abi import BoundLibrary

    dist.init_process_group("gloo", rank=rank, world_size=world)
    lib = BoundLibrary(os.path.join(ROOT, "oracle", "libgrloracle.so"), "grlo_")
    sh = sharding.shard_from_env(total)
    e = sharding.create_sharded_engine(lib, sh, width=10, height=10, num_players=2, host_threads=1)
    e.reset_seeded(sh.seeds(12345))
    out = e.alloc_outputs_host()
    prev = {k: v.copy() for k, v in out.items()}
    packed, prev_packed = (np.zeros((sh.count, e.packed_words), np.uint32) for _ in range(2))
    e.observe(e.outputs(obs=prev["obs"], mask_bits=prev["mask_bits"], obs_packed=prev_packed))
    hashes, gathered = [], []
    for t in range(T):
        ser_mask = e.mask(_abi.MASK_SERIALIZER_UDLR)          # of the state BEFORE the step
        e.step_fused(None, e.outputs(obs_packed=packed, **out), _abi.STEP_FLAG_RANDOM_POLICY, 99)
        hashes.append(e.state_hash().copy())
        # the compact records: packed observation records before/after, expanded again on the learner
        recp = sharding.pack_experience_packed(
            torch.from_numpy(prev_packed.view(np.int32)), torch.from_numpy(packed.view(np.int32)),
            torch.from_numpy(out["action_index"]), torch.from_numpy(out["reward"]), torch.from_numpy(out["done"]),
            t + 1, env_id_base=sh.first)
        gp = sharding.gather_experience(recp, dst=0)
        sm = sharding.gather_experience({"m": torch.from_numpy(ser_mask[out["action_index"] >= 0]),
                                         "k": torch.zeros(int((out["action_index"] >= 0).sum()), dtype=torch.int32)}, dst=0)
        prev_packed[:] = packed
        rec = sharding.pack_experience(
            torch.from_numpy(prev["obs"]), torch.from_numpy(out["obs"]), torch.from_numpy(prev["mask_bits"].view(np.int32)),
            torch.from_numpy(out["action_index"]), torch.from_numpy(out["reward"]), torch.from_numpy(out["done"]),
            t + 1, env_id_base=sh.first)
        g = sharding.gather_experience(rec, capacity=2 * sh.count, dst=0)
        # a capacity below what a rank holds is reported, never silent
        lossy = sharding.gather_experience(rec, capacity=5, dst=0)
        if rank == 0:
            assert int(g["dropped"].sum()) == 0 and g["counts"].tolist() == [int(c) for c in g["counts"]]
            assert lossy["action"].shape[0] == int(lossy["counts"].sum()) <= 5 * world
            assert (lossy["dropped"] + lossy["counts"]).tolist() == g["counts"].tolist()
            ex = sharding.expand_experience(gp, lib, 10, 10, 2, threads=1)
            for k in ("action", "reward", "done", "player", "turn", "env_id"):
                assert np.array_equal(ex[k], g[k].numpy()), k
            assert np.array_equal(ex["state"].view(np.uint32), g["state"].numpy().view(np.uint32))
            assert np.array_equal(ex["next_state"].view(np.uint32), g["next_state"].numpy().view(np.uint32))
            assert np.array_equal(ex["mask_bits"], sm["m"].numpy().astype(bool)), "serializer mask from the packed record"
            gathered.append({k: v.numpy() for k, v in g.items()})
        prev = {k: v.copy() for k, v in out.items()}
    stats = sharding.all_reduce_stats(e.stats())
    np.savez(os.path.join(tmp, f"rank{rank}.npz"), hashes=np.stack(hashes), stats=stats, first=sh.first)
    if rank == 0:
        np.savez(os.path.join(tmp, "gathered.npz"),
                 **{f"{k}_{t}": v for t, g in enumerate(gathered) for k, v in g.items()})
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shards_reproduce_the_unsharded_run(oracle_lib, tmp_path):
    import torch.multiprocessing as mp

    from generalsreinforcementlearning_b200 import _abi
    from helpers import new_engine

    total, T, world = 24, 12, 2
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, total, T, str(tmp_path)), nprocs=world, join=True)

    # the unsharded run
    e = new_engine(oracle_lib, 10, 10, 2, total)
    e.reset_seeded(np.arange(total, dtype=np.int64) + 12345)
    out = e.alloc_outputs_host()
    prev_obs = np.zeros_like(out["obs"])
    e.observe(e.outputs(obs=prev_obs))
    g = np.load(os.path.join(tmp_path, "gathered.npz"))
    r0, r1 = (np.load(os.path.join(tmp_path, f"rank{r}.npz")) for r in range(2))
    assert int(r0["first"]) == 0 and int(r1["first"]) == 12
    for t in range(T):
        e.step_fused(None, e.outputs(**out), _abi.STEP_FLAG_RANDOM_POLICY, 99)
        full = np.concatenate([r0["hashes"][t], r1["hashes"][t]])
        assert np.array_equal(full, e.state_hash()), f"turn {t}: sharded games diverge from the unsharded run"
        # gathered experience == the records of the unsharded run, in (env, player) order
        env, ply = np.nonzero(out["action_index"] >= 0)
        assert np.array_equal(g[f"env_id_{t}"], env.astype(np.int32))
        assert np.array_equal(g[f"player_{t}"], ply.astype(np.int32))
        assert np.array_equal(g[f"action_{t}"], out["action_index"][env, ply])
        assert np.array_equal(g[f"reward_{t}"].view(np.uint32), out["reward"][env, ply].view(np.uint32))
        assert np.array_equal(g[f"done_{t}"], out["done"][env])
        assert np.array_equal(g[f"next_state_{t}"].view(np.uint32), out["obs"][env, ply].view(np.uint32))
        assert np.array_equal(g[f"state_{t}"].view(np.uint32), prev_obs[env, ply].view(np.uint32))
        assert (g[f"turn_{t}"] == t + 1).all()
        prev_obs = out["obs"].copy()
    # statistics are summed over ranks on every rank
    assert np.array_equal(r0["stats"], r1["stats"])
    assert np.array_equal(r0["stats"], e.stats().astype(np.int64))
