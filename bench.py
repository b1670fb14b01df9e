#!/usr/bin/env python3
"""bench.py — env-steps/sec of the batched Generals.io turn engine on B200.

Metric (BASELINE.json): env-steps/sec, 20x20 board, 2 players, fog on.  One env-step = one
ProcessTurn of one game + that game's P observation tensors, P packed legal masks, P rewards
and the done flag.  A bench "step" = one fused kernel launch over all B games
resident on the GPU.  The games are first played forward, untimed, to --start-turn (default 200 of
the 500-turn episode) with the counter-based random-legal-move policy, so that a short
timed window sees mid-episode territories; the policy's next moves are then recorded once in an
untimed rollout and replayed, so that the timed region is the turn engine itself and its inputs
(that step's actions) are already resident where the arm reads them.

  python bench.py [--gpus N] [--steps K] [--warmup W]        our arm (CUDA, through the C ABI)
  python bench.py --impl reference ...                       CPU arm: the oracle restatement of
                                                             the Go engine on the host cores

`value`  : device-resident run (recorded actions in HBM, outputs stay in HBM).  The K timed launches are enqueued back
           to back, so consecutive launches overlap on the device (a launch's warps wait for the warp that held their
           games in the previous launch, not for the whole grid: csrc/grl_turn.cuh); `roofline.kernel_ms` is the timed
           region / K, `roofline.kernel_ms_serialised_launches` the same K launches with the overlap switched off (N=1).
`e2e`    : the same rollout replayed through the C ABI with HOST buffers every step: that
           step's actions copied host->device from pinned memory, reward/done/winner/error
           planes copied device->host (observations and masks stay in HBM for an on-GPU learner).
`e2e_host_obs_packed`: every result delivered to HOST memory: packed observation records
           (grl_step_outputs.obs_packed, 1,120 B per env-step) + packed masks + the small planes; a host consumer
           expands records into the fp32 tensors with grl_expand_obs (rate reported, not in the timed region).
`e2e_host_obs`: the fp32 observation tensors themselves copied to the host every step (PCIe-bound).
`shapes` : the other BASELINE configurations, device-timed the same way (10x10 x 65,536, 15x15 x 262,144,
           20x20 x 4 players x 65,536 per GPU), each with its own roofline fraction.
`gym_env` (N=1 only, informational): the generals_gym contract through GeneralsVecEnv at 65,536 envs of 15x15 — one
grl_gym_step launch per step, the random agent, the device-side auto-reset — timed with CUDA events.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np

METRIC = "env-steps/sec (20x20, 2p, fog on)"
UNIT = "env-steps/s"
W, H, P = 20, 20, 2
GAMES_PER_GPU = 65536
BASE_SEED = 12345  # internal/game/engine_test.go:16
POLICY_SEED = 2024


def algorithmic_bytes_per_env_step(W, H, P):
    """Bytes that must cross HBM per env-step with the slab layout of csrc/grl_layout.h
    (DESIGN.md 'Algorithmic bytes'): slab read + terrain read + slab write + outputs."""
    N = W * H
    NW = (N + 31) // 32
    NA = (N + 7) & ~7
    hdr = (8 + 5 * P + 3) & ~3
    off_army = (hdr + (3 * P + 2) * NW + 3) & ~3
    slab = ((off_army + NA // 2 + 7) & ~7) * 4
    static = ((3 * NW + 3) & ~3) * 4
    obs = 9 * N * 4 * P
    mask = ((4 * N + 31) // 32) * 4 * P
    small = 4 * P + 1  # reward + done
    acts = 8 * P       # one grl_action per player read per step
    return dict(read=slab + static + acts, write=slab + obs + mask + small,
                total=2 * slab + static + acts + obs + mask + small,
                slab=slab, static=static, obs=obs, mask=mask)


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def recorded_traffic():
    """dram bytes per launch of the headline kernel from the committed ncu --set full capture — only when that capture
    profiled the very sources the loaded library was built from (tools/ncu_summarize.py --traffic stores their hash);
    a stale capture yields None rather than a number that no longer describes the kernel."""
    try:
        from generalsreinforcementlearning_b200 import build as grl_build

        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f)
        return t if t.get("turn_source_hash") == grl_build.turn_source_hash() else None
    except Exception:
        return None


class ClockSampler:
    """SM clock + throttle reasons polled through NVML DURING the timed region (the region
    is a fraction of a second, so nvidia-smi's 100 ms loop would miss it)."""

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.sm, self.reasons, self.power = [], set(), []
        self.stop_flag = False
        self.thread = None
        self.max_mhz = None
        try:
            import pynvml

            self.nv = pynvml
            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index(gpu_index))
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception as exc:  # pragma: no cover
            self.nv = None
            self.err = repr(exc)

    @staticmethod
    def _physical_index(local):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            try:
                return int(vis.split(",")[local])
            except Exception:
                return local
        return local

    def _poll(self):
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or getattr(
            nv, "nvmlDeviceGetCurrentClocksThrottleReasons")
        while not self.stop_flag:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
                r = get_reasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        if self.nv:
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()

    def stop(self):
        if not self.nv:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable: " + getattr(self, "err", "")]}
        self.stop_flag = True
        self.thread.join(timeout=2)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.sm),
                "power_w_max": max(self.power) if self.power else None}


def load_oracle():
    from generalsreinforcementlearning_b200._abi import BoundLibrary

    path = os.path.join(ROOT, "oracle", "libgrloracle.so")
    if not os.path.exists(path):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
    return BoundLibrary(path, "grlo_")


def time_oracle(games, steps, warmup, threads=0, with_readouts=True, start_turn=0, turns_per_step=1):
    """The CPU arm: C restatement of the Go engine (oracle/), one game per job, all host threads.
    Bounded sample of the same workload: `games` 20x20x2p games, seeds BASE_SEED+i, played forward (untimed) to
    `start_turn` with the same counter-based policy, then the policy's recorded moves replayed from host memory;
    a "step" is `turns_per_step` turns of every game."""
    from generalsreinforcementlearning_b200 import _abi
    from generalsreinforcementlearning_b200.engine import BatchedEngine, make_config

    lib = load_oracle()
    cores = os.cpu_count() or 1
    e = BatchedEngine(lib, make_config(lib, num_envs=games, width=W, height=H, num_players=P, host_threads=threads))
    seeds = np.arange(games, dtype=np.int64) + BASE_SEED

    def rewind():
        e.reset_seeded(seeds)
        for _ in range(start_turn):
            e.step(None, _abi.STEP_FLAG_RANDOM_POLICY, POLICY_SEED)

    rewind()
    n_turns = (warmup + steps) * turns_per_step
    rec = []
    for _ in range(n_turns):  # untimed: record the policy's moves
        a = e.sample_actions(POLICY_SEED)
        rec.append(a)
        e.step(a)
    rewind()
    out = e.alloc_outputs_host()
    outs = e.outputs(obs=out["obs"], mask_bits=out["mask_bits"], reward=out["reward"], done=out["done"])
    run = (lambda a: e.step_fused(a, outs)) if with_readouts else (lambda a: e.step(a))
    for t in range(warmup * turns_per_step):
        run(rec[t])
    s0 = int(e.stats()[0])
    t0 = time.perf_counter()
    for t in range(warmup * turns_per_step, n_turns):
        run(rec[t])
    dt = time.perf_counter() - t0
    done_steps = int(e.stats()[0]) - s0
    e.close()
    return done_steps / dt, dt, cores if threads == 0 else threads, done_steps


def clamp_start_turn(start_turn, turns):
    """The fast-forward may not push the timed window past the 500-turn episode cap."""
    return max(0, min(start_turn, 498 - turns))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    games = args.games
    # a step = `tps` turns of every game, sized so that the K timed steps take >= ~2.5 s on this box (a 0.1 s sample
    # moved the figure by +-25 % between boxes); the rate is per env-step, so the step's size does not enter it
    start = min(args.start_turn, 200)
    tps_max = max(1, (498 - start) // max(1, args.steps + args.warmup))
    # calibration: this box, this state; the better of two short runs (a cold first run would undersize the sample)
    cal = max(time_oracle(games, 2, 1, start_turn=start)[0] for _ in range(2))
    tps = int(max(1, min(np.ceil(2.5 * cal / (games * max(1, args.steps))), tps_max)))
    rate, dt, cores, n = time_oracle(games, args.steps, args.warmup, start_turn=start, turns_per_step=tps)
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(1, args.steps), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int64+f32", "data": "synthetic",
        "config": {"workload": "20x20 2p fog-on random-legal-move rollouts (BASELINE headline config; seeds 12345+i; "
                               f"played forward untimed to turn {start}; the counter-based policy's moves recorded once and "
                               "replayed from host memory)",
                   "games_per_step": games, "turns_per_step": tps, "players": P, "board": [W, H], "episode_cap": 500,
                   "start_turn": start,
                   "parallelism": f"one game per job over {cores} host threads (the GPU arm's games; a step is {tps} turn(s) of each)"},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{games} games x {args.steps * tps} turns from turn {start} ({n} env-steps, {dt:.2f} s) with "
                                   "observation, mask and reward generation; C restatement of the Go engine (Go toolchain "
                                   "unavailable)"},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def gym_contract_rate(envs=65536, board=15, steps=60):
    """Extra, informational: the generals_gym contract through GeneralsVecEnv at the reference's default
    board (15x15) — one grl_gym_step launch per step, the reference's random agent as the policy (drawn inside the step's
    launch, and as a grl_gym_sample launch of its own for comparison), episode
    ends spread over time so that every step re-seeds envs on the device (grl_gym_autoreset).  Timed with CUDA events on
    torch's stream (the env's stream); never allowed to break the headline line."""
    try:
        import torch

        from generalsreinforcementlearning_b200.gym_env import GeneralsVecEnv

        env = GeneralsVecEnv(envs, board, board, max_turns=500, seed=12345, auto_reset="device")
        env.reset()
        env._calls.copy_(torch.randint(0, env.max_turns, (envs,), device=env._calls.device, dtype=torch.int32))

        def timed(one_step):
            for _ in range(5):
                one_step()
            torch.cuda.synchronize()
            l0 = env.engine.launch_count()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                one_step()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / steps, (env.engine.launch_count() - l0) / steps

        ms_sampler, launches_sampler = timed(lambda: env.step(env.sample_actions()))   # the agent as a launch of its own
        ms, launches = timed(lambda: env.step(None))                                   # the agent drawn inside the step
        env.close()
        return {"value": envs / (ms * 1e-3), "unit": "env-steps/s", "envs": envs, "board": [board, board], "ms_per_vector_step": ms,
                "launches_per_step": launches,
                "with_sampler_launch": {"value": envs / (ms_sampler * 1e-3), "ms_per_vector_step": ms_sampler,
                                        "launches_per_step": launches_sampler},
                "what": "GeneralsVecEnv.step(None): one grl_gym_step launch that also draws the reference's random agent from the "
                        "direction masks in registers (the same indices grl_gym_sample returns) + device-side auto-reset of the "
                        "~130 envs whose episode ends each step (grl_gym_autoreset); with_sampler_launch = step(sample_actions()), "
                        "the agent as a separate grl_gym_sample launch over the mask bytes; reference gym path: 12 steps/s per env"}
    except Exception as exc:  # noqa: BLE001 - informational figure only
        return {"unavailable": repr(exc)[:200]}


SHAPES = [  # the other BASELINE.json configurations: (W, H, P, games per GPU, which config)
    (10, 10, 2, 65536, "configs[1]: 2-player 10x10 random-action rollouts, 65,536 games"),
    (15, 15, 2, 262144, "configs[2]: 2-player 15x15 fog on, 262,144 games with action masks + observation planes"),
    (20, 20, 4, 65536, "configs[3]: 4-player 20x20 fog on, 20 cities, game-sharded (65,536 games per GPU)"),
]


class Rollout:
    """B games of one shape on this rank's GPU: played forward to a start turn, the policy's next moves recorded, and the
    recorded stretch replayed (device-timed) from the same state as often as needed."""

    def __init__(self, lib, torch, dev, stream, Wb, Hb, Pb, B, rank, start_turn, turns):
        from generalsreinforcementlearning_b200.engine import BatchedEngine, make_config

        self.torch, self.dev, self.stream, self.B, self.turns = torch, dev, stream, B, turns
        self.e = BatchedEngine(lib, make_config(lib, num_envs=B, width=Wb, height=Hb, num_players=Pb, device=dev.index,
                                                env_id_base=rank * B, host_threads=0, max_actions=Pb))
        self.e.set_stream(stream.cuda_stream)
        self.seeds = np.arange(B, dtype=np.int64) + BASE_SEED + rank * B
        self.start = clamp_start_turn(start_turn, turns)
        e = self.e
        self.obs = torch.empty((B, Pb, 9, Hb, Wb), dtype=torch.float32, device=dev)
        self.mask = torch.empty((B, Pb, e.mask_words), dtype=torch.int32, device=dev)
        self.reward = torch.empty((B, Pb), dtype=torch.float32, device=dev)
        self.done = torch.empty(B, dtype=torch.uint8, device=dev)
        self.outs = e.outputs(obs=self.obs, mask_bits=self.mask, reward=self.reward, done=self.done)
        # record the rollout's actions once (untimed)
        self.rewind()
        self.drec = torch.empty((turns, B, e.A, 8), dtype=torch.uint8, device=dev)   # HBM-resident inputs
        for t in range(turns):
            e.sample_actions(POLICY_SEED, self.drec[t])
            e.step_fused(self.drec[t], e.outputs(done=self.done))
        torch.cuda.synchronize()

    def rewind(self):
        """Back to the start turn: the policy is counter-based, so replaying it reproduces the state exactly."""
        from generalsreinforcementlearning_b200 import _abi

        self.e.reset_seeded(self.seeds)
        for _ in range(self.start):
            self.e.step(None, _abi.STEP_FLAG_RANDOM_POLICY, POLICY_SEED)

    def timed(self, warm, K, barrier, before=None):
        """K fused launches after `warm` untimed ones; returns (ms on this rank, env-steps executed, launches)."""
        torch, e = self.torch, self.e
        self.rewind()
        for t in range(warm):
            e.step_fused(self.drec[t], self.outs)
        torch.cuda.synchronize()
        s0, l0 = int(e.stats()[0]), e.launch_count()
        if before:
            before()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record(self.stream)
        for t in range(warm, warm + K):
            e.step_fused(self.drec[t], self.outs)
        ev1.record(self.stream)
        barrier()
        launches = e.launch_count() - l0   # before stats(), which launches a counting kernel of its own
        return ev0.elapsed_time(ev1), int(e.stats()[0]) - s0, launches


def run_cuda(args):
    import torch
    import torch.distributed as dist

    from generalsreinforcementlearning_b200 import load_library

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # stdout carries exactly one JSON line: NCCL's own log lines ("NCCL version ...") go to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)

    lib = load_library()
    B = args.games
    K, Wm = args.steps, max(3, args.warmup)
    if Wm + K + 2 > 498:
        raise SystemExit("warmup+steps must stay below the 500-turn episode cap")
    stream = torch.cuda.Stream(device=dev)  # the kernel's stream: CUDA events are recorded on it
    torch.cuda.set_stream(stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    def sum_over_ranks(v):
        if world > 1:
            t = torch.tensor([v], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            return float(t.item())
        return float(v)

    peak, peak_src = measured_peak_gbs()

    # ---------------- device-resident arm (the headline) -----------------------------------
    ro = Rollout(lib, torch, dev, stream, W, H, P, B, rank, args.start_turn, Wm + K)
    e = ro.e
    sampler = ClockSampler(local)
    ms, env_steps, launches = ro.timed(Wm, K, barrier, before=(sampler.start if rank == 0 else None))
    clocks = sampler.stop() if rank == 0 else None
    ms_max = max_over_ranks(ms)
    total_steps = sum_over_ranks(env_steps)
    value = total_steps / (ms_max * 1e-3)
    if args.quick:
        if rank == 0:
            print(json.dumps({"value": value, "ms_per_step": ms_max / K, "lib": os.environ.get("GRL_LIB_PATH", "default"),
                              "clocks": clocks}))
        e.close()
        return 0

    # ---------------- end-to-end arms: host actions in, host results out, every step ----------
    A = e.A
    rec = torch.empty((Wm + K, B, A, 8), dtype=torch.uint8).pin_memory()    # host copy of the recorded moves
    rec.copy_(ro.drec)
    torch.cuda.synchronize()
    h_reward = torch.empty((B, P), dtype=torch.float32).pin_memory()
    h_done = torch.empty(B, dtype=torch.uint8).pin_memory()
    h_winner = torch.empty(B, dtype=torch.int8).pin_memory()
    h_err = torch.empty(B, dtype=torch.uint8).pin_memory()
    h_mask = torch.empty((B, P, e.mask_words), dtype=torch.int32).pin_memory()
    small = h_reward.numel() * 4 + h_done.numel() + h_winner.numel() + h_err.numel()

    def e2e_run(mode):
        if mode == "fp32":       # every fp32 observation tensor + mask to the host
            h_obs = torch.empty((B, P, 9, H, W), dtype=torch.float32).pin_memory()
            o = e.outputs(obs=h_obs, mask_bits=h_mask, reward=h_reward, done=h_done, winner=h_winner, step_error=h_err)
            d2h, steps_k = h_obs.numel() * 4 + h_mask.numel() * 4, min(K, 8)
        elif mode == "packed":   # packed observation records + mask to the host: everything a host consumer needs
            h_packed = torch.empty((B, e.packed_words), dtype=torch.int32).pin_memory()
            o = e.outputs(obs_packed=h_packed, mask_bits=h_mask, reward=h_reward, done=h_done, winner=h_winner, step_error=h_err)
            d2h, steps_k = h_packed.numel() * 4 + h_mask.numel() * 4, K
        else:                    # observations and masks stay in HBM for an on-GPU learner
            o = e.outputs(obs=ro.obs, mask_bits=ro.mask, reward=h_reward, done=h_done, winner=h_winner, step_error=h_err)
            d2h, steps_k = 0, K
        ro.rewind()
        for t in range(Wm):
            e.step_fused(rec[t], o)
        s0 = int(e.stats()[0])
        barrier()
        t0 = time.perf_counter()
        for t in range(Wm, Wm + steps_k):
            e.step_fused(rec[t], o)  # returns after the D2H copies completed (host buffers are valid)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        n = int(e.stats()[0]) - s0
        dt_max = max_over_ranks(dt * 1e3) * 1e-3
        extra = {}
        if mode == "packed" and rank == 0:   # what a host consumer pays to get fp32 tensors back (not in the timed region)
            n_exp = min(B, 8192)
            out = np.empty((n_exp, P, 9, H, W), np.float32)
            t1 = time.perf_counter()
            e.expand_obs(h_packed.numpy()[:n_exp].view(np.uint32), out)
            extra = {"host_expand_env_steps_per_s": n_exp / (time.perf_counter() - t1), "host_expand_threads": os.cpu_count()}
        return sum_over_ranks(n) / dt_max, B * A * 8, d2h + small, steps_k, extra

    # the same K launches with the launch overlap switched off (GRL_LAUNCH_OVERLAP is read when an env is created): what one
    # launch costs when every launch waits for the whole grid before it, reported beside the headline, not instead of it
    serial_ms = None
    if world == 1:
        prev = os.environ.get("GRL_LAUNCH_OVERLAP")
        os.environ["GRL_LAUNCH_OVERLAP"] = "0"
        try:
            ro_s = Rollout(lib, torch, dev, stream, W, H, P, B, rank, args.start_turn, Wm + K)
            ms_s0, _, _ = ro_s.timed(Wm, K, barrier)
            serial_ms = ms_s0 / K
            ro_s.e.close()
            del ro_s
            torch.cuda.empty_cache()
        finally:
            if prev is None:
                del os.environ["GRL_LAUNCH_OVERLAP"]
            else:
                os.environ["GRL_LAUNCH_OVERLAP"] = prev

    e2e_val, h2d, d2h, _, _ = e2e_run("device")
    e2e_packed, h2d_p, d2h_p, _, exp_info = e2e_run("packed")
    e2e_full, h2d_f, d2h_f, k_full, _ = e2e_run("fp32")
    e.close()
    del ro

    # ---------------- the other BASELINE shapes, device-timed the same way -------------------
    shapes = []
    for (Ws, Hs, Ps, Bs, what) in SHAPES:
        try:
            r = Rollout(lib, torch, dev, stream, Ws, Hs, Ps, Bs, rank, args.start_turn, Wm + K)
            ms_s, n_s, _ = r.timed(Wm, K, barrier)
            r.e.close()
            del r
            torch.cuda.empty_cache()
            ms_s_max, n_tot = max_over_ranks(ms_s), sum_over_ranks(n_s)
            alg_s = algorithmic_bytes_per_env_step(Ws, Hs, Ps)["total"]
            ach = alg_s * (n_s / K) / (ms_s / K * 1e-3) / 1e9
            shapes.append({"board": [Ws, Hs], "players": Ps, "games_per_gpu": Bs, "config": what,
                           "value": n_tot / (ms_s_max * 1e-3), "unit": UNIT, "ms_per_step": ms_s_max / K,
                           "bytes_per_env_step": alg_s,
                           "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak}})
        except Exception as exc:  # noqa: BLE001 - the extra shapes never break the headline line
            shapes.append({"board": [Ws, Hs], "players": Ps, "games_per_gpu": Bs, "unavailable": repr(exc)[:200]})

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    alg = algorithmic_bytes_per_env_step(W, H, P)
    per_launch_bytes = alg["total"] * (env_steps / K)
    kernel_ms = ms / K
    achieved = per_launch_bytes / (kernel_ms * 1e-3) / 1e9
    traffic = recorded_traffic()

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        rate, dt, cores, n = time_oracle(4096, 12, 2)          # probe
        # bounded sample: ~15 s of CPU work on this box, mid-episode turns of the same rollouts
        steps_c = 240
        games = int(min(32768, max(2048, (rate * 15.0 / steps_c) // 1024 * 1024)))
        rate, dt, cores, n = time_oracle(games, steps_c, 2, start_turn=args.start_turn)
        rate_bare, dt_bare, _, n_bare = time_oracle(8192, 100, 2, with_readouts=False, start_turn=args.start_turn)
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
               "value_turn_only": rate_bare,  # ProcessTurn alone, no observation / mask generation
               "sample": f"{games} games x {steps_c} turns from turn {args.start_turn} ({n} env-steps, {dt:.1f} s) incl. "
                         "observation/mask/reward; C restatement of the Go engine (oracle/), Go toolchain unavailable"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm,
        "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u16/u32 bitmask + f32 planes", "data": "synthetic",
        "config": {"workload": "20x20 2p fog-on random-legal-move rollouts (BASELINE headline config; seeds 12345+i; played "
                               f"forward untimed to turn {clamp_start_turn(args.start_turn, Wm + K)}; the counter-based policy's "
                               "moves recorded once and replayed from HBM)",
                   "games_per_gpu": B, "players": P, "board": [W, H], "episode_cap": 500,
                   "start_turn": clamp_start_turn(args.start_turn, Wm + K),
                   "cache": f"working set {(alg['total'] * B) / 1e6:.0f} MB per step > 126 MB L2 (no flush needed)",
                   "parallelism": f"games sharded by env index over {world} GPU(s), no collective on the step path"},
        "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "what": "per step through grl_step_fused with HOST buffers: actions copied host->device from pinned memory, "
                        "reward/done/winner/step_error delivered to pinned host memory (written in place by the kernel); "
                        "observation and mask planes stay in HBM for an on-GPU learner"},
        "e2e_host_obs_packed": {"value": e2e_packed, "unit": UNIT, "h2d_bytes_per_step": h2d_p, "d2h_bytes_per_step": d2h_p,
                                "what": "EVERY result delivered to host memory each step: packed observation records "
                                        "(grl_step_outputs.obs_packed) + packed masks + reward/done/winner/step_error; "
                                        "grl_expand_obs turns records into the fp32 tensors bit for bit on the host", **exp_info},
        "e2e_host_obs": {"value": e2e_full, "unit": UNIT, "h2d_bytes_per_step": h2d_f, "d2h_bytes_per_step": d2h_f,
                         "steps": k_full, "what": "as e2e plus every fp32 observation tensor and mask copied to the host"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "peak_source": peak_src, "bytes_per_env_step": alg["total"],
                     "kernel": "grl_turn_kernel<2,20,20,32,true,true,false>", "kernel_ms": kernel_ms,
                     "traffic": traffic.get("dram_bytes_per_launch") if traffic else None,
                     "traffic_source": traffic.get("source") if traffic else "no ncu capture of this build under profiles/",
                     "frac_of_nominal_8TBs": achieved / 8000.0,
                     "launches": "the K timed launches are enqueued back to back and overlap on the device (csrc/grl_turn.cuh, "
                                 "launch overlap): kernel_ms = timed region / K",
                     "kernel_ms_serialised_launches": serial_ms},
        "clocks": clocks,
        "env_steps_per_launch": env_steps / K,
        "shapes": shapes,
    }
    if cpu:
        line["cpu_baseline"] = cpu
    if world == 1 and not args.no_cpu_baseline:
        gym = gym_contract_rate()
        if gym:
            line["gym_env"] = gym
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--start-turn", type=int, default=200,
                    help="play the games forward (untimed) to this turn before the timed window (500-turn episodes)")
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--games", type=int, default=GAMES_PER_GPU, help="games resident per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--quick", action="store_true", help="device-resident arm only (kernel experiments)")
    args = ap.parse_args()
    if args.start_turn < 0:
        raise SystemExit("--start-turn must be >= 0")
    if args.impl == "reference":
        return run_reference(args)
    return run_cuda(args)


if __name__ == "__main__":
    sys.exit(main())
