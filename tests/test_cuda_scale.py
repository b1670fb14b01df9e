"""BASELINE-size runs: parity through digests (too large to copy every plane back) and
size-independent properties of the turn engine."""
import numpy as np
import pytest

from generalsreinforcementlearning_b200 import _abi
from helpers import new_engine

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("W,H,P,B,T", [(10, 10, 2, 65536, 40), (15, 15, 2, 32768, 30), (20, 20, 2, 16384, 60),
                                        (20, 20, 4, 8192, 40),
                                        # the BASELINE configurations at their FULL batch sizes (configs[2], the headline, configs[3])
                                        (15, 15, 2, 262144, 20), (20, 20, 2, 65536, 20), (20, 20, 4, 65536, 20)])
def test_digest_parity_at_scale(cuda_lib, oracle_lib, W, H, P, B, T):
    """Every env's full-state digest and its reward/done/mask planes match the oracle while
    both play the same counter-based random policy (BASELINE configs 2-4)."""
    import torch

    gc = new_engine(cuda_lib, W, H, P, B, host_threads=0)
    oc = new_engine(oracle_lib, W, H, P, B, host_threads=0)
    seeds = np.arange(B, dtype=np.int64) + 12345
    gc.reset_seeded(seeds)
    oc.reset_seeded(seeds)
    assert np.array_equal(gc.state_hash(), oc.state_hash())
    dev = torch.device("cuda:0")
    obs = torch.empty((B, P, 9, H, W), dtype=torch.float32, device=dev)
    mask = torch.empty((B, P, gc.mask_words), dtype=torch.int32, device=dev)
    reward = torch.empty((B, P), dtype=torch.float32, device=dev)
    done = torch.empty(B, dtype=torch.uint8, device=dev)
    oo = oc.alloc_outputs_host()
    oo_small = {k: oo[k] for k in ("mask_bits", "reward", "done")}
    for t in range(T):
        gc.step_fused(None, gc.outputs(obs=obs, mask_bits=mask, reward=reward, done=done),
                      _abi.STEP_FLAG_RANDOM_POLICY, 2024)
        with_obs = t % 10 == 9 or t == T - 1
        oc.step_fused(None, oc.outputs(**(oo if with_obs else oo_small)), _abi.STEP_FLAG_RANDOM_POLICY, 2024)
        assert np.array_equal(gc.state_hash(), oc.state_hash()), f"turn {t}"
        assert np.array_equal(reward.cpu().numpy().view(np.uint32), oo["reward"].view(np.uint32)), f"turn {t}"
        assert np.array_equal(done.cpu().numpy(), oo["done"]), f"turn {t}"
        assert np.array_equal(mask.cpu().numpy().view(np.uint32), oo["mask_bits"]), f"turn {t}"
        if with_obs:  # observation planes through row digests computed on the device
            rows, words = B * P, 9 * W * H
            assert np.array_equal(gc.buffer_hash(obs, words, rows), oc.buffer_hash(oo["obs"], words, rows)), f"turn {t}"
    assert np.array_equal(gc.stats(), oc.stats())


def test_properties_at_headline_size(cuda_lib):
    """20x20 2p, 65,536 games: invariants that need no oracle."""
    import torch

    W, H, P, B = 20, 20, 2, 65536
    e = new_engine(cuda_lib, W, H, P, B, host_threads=0)
    seeds = np.arange(B, dtype=np.int64) % 4096 + 1  # 16 replicas of 4,096 distinct maps
    e.reset_seeded(seeds)
    dev = torch.device("cuda:0")
    obs = torch.empty((B, P, 9, H, W), dtype=torch.float32, device=dev)
    reward = torch.empty((B, P), dtype=torch.float32, device=dev)
    acts = torch.empty((B, e.A, 8), dtype=torch.uint8, device=dev)
    for t in range(50):
        # replicas must receive identical moves: draw for the first 4,096 envs and tile them
        e.sample_actions(7, acts)
        tiled = acts[:4096].repeat(16, 1, 1).contiguous()
        e.step_fused(tiled, e.outputs(obs=obs, reward=reward))
    e.sync()
    h = e.state_hash().reshape(16, 4096)
    assert (h == h[0]).all(), "identical seeds + identical moves must give identical games"
    o = obs.view(16, 4096, P, 9, H * W)
    assert torch.equal(o[0], o[5]) and torch.equal(o[0], o[15])
    # channel algebra of StateToTensor: visible + fog == 1; own/enemy/neutral/mountain partition the visible tiles
    assert torch.equal(obs[:, :, 7] + obs[:, :, 8], torch.ones_like(obs[:, :, 7]))
    part = obs[:, :, 2] + obs[:, :, 3] + obs[:, :, 4] + obs[:, :, 6]
    assert torch.equal(part, obs[:, :, 7])
    assert float(obs.min()) >= 0.0 and float(obs.max()) <= 1.0
    st = e.get_state(0, 256)
    # armies are conserved up to production: every tile army is non-negative and mountains stay empty/neutral
    assert (st["army"] >= 0).all()
    assert (st["owner"][st["type"] == 3] == -1).all() and (st["army"][st["type"] == 3] == 0).all()
    assert (st["turn"] == 50).all()
    # a step on finished games is idempotent
    s2 = {k: v.copy() for k, v in st.items()}
    s2["game_over"][:] = 1
    e.set_state(s2, 0)
    before = e.state_hash()[:256].copy()
    e.step(None)
    assert np.array_equal(e.state_hash()[:256], before)


@pytest.mark.parametrize("W,H,P,T", [(10, 10, 2, 500), (20, 20, 2, 500)])
def test_10k_games_full_episode_replay(cuda_lib, oracle_lib, W, H, P, T):
    """The north star's parity run: 10,000 seeded games played to the 500-turn episode cap
    (generals_env.py:58) with the random-legal-move policy, then re-seeded with the next
    block of seeds and played on.  Every turn: every game's full-state digest,
    reward bits, done, winner, step_error and packed mask equal the oracle's; observation
    digests every 25th turn.  The recorded actions are REPLAYED into the CUDA engine from the
    oracle's draw, so the two engines see identical action sequences by construction."""
    import torch

    B = 10000
    gc = new_engine(cuda_lib, W, H, P, B, host_threads=0)
    oc = new_engine(oracle_lib, W, H, P, B, host_threads=0)
    dev = torch.device("cuda:0")
    obs = torch.empty((B, P, 9, H, W), dtype=torch.float32, device=dev)
    go = gc.alloc_outputs_host()
    del go["obs"]
    oo = oc.alloc_outputs_host()
    oo_small = {k: v for k, v in oo.items() if k != "obs"}
    finished = errors = 0
    for episode, turns in ((0, T), (1, 40)):
        seeds = np.arange(B, dtype=np.int64) + 12345 + episode * B
        gc.reset_seeded(seeds)
        oc.reset_seeded(seeds)
        assert np.array_equal(gc.state_hash(), oc.state_hash())
        for t in range(turns):
            acts = oc.sample_actions(2024 + episode)
            with_obs = t % 25 == 24
            gc.step_fused(acts, gc.outputs(obs=obs, **go))
            oc.step_fused(acts, oc.outputs(**(oo if with_obs else oo_small)))
            ctx = f"episode {episode} turn {t}"
            assert np.array_equal(gc.state_hash(), oc.state_hash()), ctx
            for k in go:
                a, b = go[k], oo[k]
                if a.dtype == np.float32:
                    a, b = a.view(np.uint32), b.view(np.uint32)
                assert np.array_equal(a, b), f"{ctx}: {k}"
            if with_obs:
                rows, words = B * P, 9 * W * H
                assert np.array_equal(gc.buffer_hash(obs, words, rows), oc.buffer_hash(oo["obs"], words, rows)), ctx
            errors += int((oo["step_error"] != 0).sum())
        finished += int(oo["done"].sum())
    assert np.array_equal(gc.stats(), oc.stats())
    # the run must have exercised the abort path (Q5) — and, on the small board, game endings
    assert errors > 0
    if W == 10:
        assert finished > 0


def test_step_is_cuda_graph_capturable(cuda_lib, oracle_lib):
    """With device buffers a fused step is exactly one kernel launch on the env's stream, so a rollout
    segment can be captured in a CUDA graph (launch-bound small batches).  Eight self-play turns are
    captured once and replayed; the state must equal the oracle's after every replay."""
    import torch

    W, H, P, B, K = 10, 10, 2, 512, 8
    e = new_engine(cuda_lib, W, H, P, B)
    o = new_engine(oracle_lib, W, H, P, B)
    seeds = np.arange(B, dtype=np.int64) + 77
    e.reset_seeded(seeds)
    o.reset_seeded(seeds)
    dev = torch.device("cuda:0")
    obs = torch.zeros((K, B, P, 9, H, W), dtype=torch.float32, device=dev)  # one observation buffer per captured turn
    reward = torch.zeros((K, B, P), dtype=torch.float32, device=dev)
    done = torch.zeros((K, B), dtype=torch.uint8, device=dev)
    side = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(side):
        e.use_torch_stream()
        for k in range(K):  # warm-up outside the capture (kernel attributes are set on first launch)
            e.step_fused(None, e.outputs(obs=obs[k], reward=reward[k], done=done[k]), _abi.STEP_FLAG_RANDOM_POLICY, 5)
    side.synchronize()
    for k in range(K):
        o.step(None, _abi.STEP_FLAG_RANDOM_POLICY, 5)
    assert np.array_equal(e.state_hash(), o.state_hash())
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side):
        for k in range(K):
            e.step_fused(None, e.outputs(obs=obs[k], reward=reward[k], done=done[k]), _abi.STEP_FLAG_RANDOM_POLICY, 5)
    oo = o.alloc_outputs_host()
    for rep in range(3):
        graph.replay()
        torch.cuda.synchronize()
        for k in range(K):
            o.step_fused(None, o.outputs(**oo), _abi.STEP_FLAG_RANDOM_POLICY, 5)
        assert np.array_equal(e.state_hash(), o.state_hash()), f"replay {rep}"
        assert np.array_equal(reward[K - 1].cpu().numpy().view(np.uint32), oo["reward"].view(np.uint32))
        assert np.array_equal(obs[K - 1].cpu().numpy().view(np.uint32), oo["obs"].view(np.uint32))
    assert int(e.get_state(0, 1)["turn"][0]) == K * 4


@pytest.mark.parametrize("W,H,P,B,T", [(10, 10, 2, 65536, 80), (15, 15, 2, 40002, 50), (20, 20, 2, 32768, 50), (20, 20, 4, 8192, 40)])
def test_overlapped_launches_keep_parity(cuda_lib, oracle_lib, W, H, P, B, T, monkeypatch):
    """Turn launches enqueued back to back overlap on the device (programmatic dependent launch + per-warp epoch words,
    grl_turn.cuh): the next launch's warps start while the previous grid drains and wait only for the warp that held
    their games.  T launches without a host synchronisation in between must leave exactly the state and the read-outs
    of T oracle turns — and of the same launches serialised (GRL_LAUNCH_OVERLAP=0)."""
    import torch

    dev = torch.device("cuda:0")
    seeds = np.arange(B, dtype=np.int64) + 777

    def run(overlap):
        monkeypatch.setenv("GRL_LAUNCH_OVERLAP", "1" if overlap else "0")
        e = new_engine(cuda_lib, W, H, P, B, host_threads=0)
        e.reset_seeded(seeds)
        obs = torch.empty((B, P, 9, H, W), dtype=torch.float32, device=dev)
        mask = torch.empty((B, P, e.mask_words), dtype=torch.int32, device=dev)
        reward = torch.empty((B, P), dtype=torch.float32, device=dev)
        done = torch.empty(B, dtype=torch.uint8, device=dev)
        outs = e.outputs(obs=obs, mask_bits=mask, reward=reward, done=done)
        for t in range(T):          # no synchronisation: every launch may overlap the one before
            if t % 7 == 3:
                e.step(None, _abi.STEP_FLAG_RANDOM_POLICY, 4711)          # step-only launches in the same chain
            else:
                e.step_fused(None, outs, _abi.STEP_FLAG_RANDOM_POLICY, 4711)
        torch.cuda.synchronize()
        return e, obs, mask.cpu().numpy().view(np.uint32), reward.cpu().numpy().view(np.uint32), done.cpu().numpy()

    oc = new_engine(oracle_lib, W, H, P, B, host_threads=0)
    oc.reset_seeded(seeds)
    oo = oc.alloc_outputs_host()
    for t in range(T):
        if t % 7 == 3:
            oc.step(None, _abi.STEP_FLAG_RANDOM_POLICY, 4711)
        else:
            oc.step_fused(None, oc.outputs(**({k: oo[k] for k in ("mask_bits", "reward", "done")} if t < T - 1 else oo)),
                          _abi.STEP_FLAG_RANDOM_POLICY, 4711)
    rows, words = B * P, 9 * W * H
    want_obs = oc.buffer_hash(oo["obs"], words, rows)
    for overlap in (True, False):
        e, obs, mask, reward, done = run(overlap)
        assert np.array_equal(e.state_hash(), oc.state_hash()), f"overlap={overlap}"
        assert np.array_equal(reward, oo["reward"].view(np.uint32)) and np.array_equal(done, oo["done"]), f"overlap={overlap}"
        assert np.array_equal(mask, oo["mask_bits"]), f"overlap={overlap}"
        assert np.array_equal(e.buffer_hash(obs, words, rows), want_obs), f"overlap={overlap}"
        assert np.array_equal(e.stats(), oc.stats())
        e.close()


def test_two_envs_interleaved_on_one_stream(cuda_lib, oracle_lib):
    """Two envs stepped alternately on ONE stream without host synchronisation: a launch overlaps only a predecessor of its
    own env with nothing else enqueued in between (csrc/grl_abi.cu, launch overlap bookkeeping), so the interleaved
    launches serialise; the chains of each env alone, enqueued afterwards, overlap again.  Both must match the oracle."""
    import torch

    stream = torch.cuda.Stream()
    W, H, P, B, T = 10, 10, 2, 32768, 30
    envs, outs, refs = [], [], []
    for k in range(2):
        e = new_engine(cuda_lib, W, H, P, B, host_threads=0)
        e.set_stream(stream.cuda_stream)
        seeds = np.arange(B, dtype=np.int64) + 1000 * (k + 1)
        e.reset_seeded(seeds)
        reward = torch.empty((B, P), dtype=torch.float32, device="cuda:0")
        done = torch.empty(B, dtype=torch.uint8, device="cuda:0")
        envs.append(e)
        outs.append((reward, done, e.outputs(reward=reward, done=done)))
        o = new_engine(oracle_lib, W, H, P, B, host_threads=0)
        o.reset_seeded(seeds)
        refs.append(o)
    with torch.cuda.stream(stream):
        for t in range(T):                      # A, B, A, B, ...
            for k in range(2):
                envs[k].step_fused(None, outs[k][2], _abi.STEP_FLAG_RANDOM_POLICY, 5 + k)
        for k in range(2):                      # then a chain of each env alone
            for t in range(T):
                envs[k].step_fused(None, outs[k][2], _abi.STEP_FLAG_RANDOM_POLICY, 5 + k)
    stream.synchronize()
    for k in range(2):
        oo = refs[k].alloc_outputs_host()
        for t in range(2 * T):
            refs[k].step_fused(None, refs[k].outputs(reward=oo["reward"], done=oo["done"]), _abi.STEP_FLAG_RANDOM_POLICY, 5 + k)
        assert np.array_equal(envs[k].state_hash(), refs[k].state_hash()), f"env {k}"
        assert np.array_equal(outs[k][0].cpu().numpy().view(np.uint32), oo["reward"].view(np.uint32)), f"env {k}"
        assert np.array_equal(outs[k][1].cpu().numpy(), oo["done"]), f"env {k}"
        envs[k].close()
