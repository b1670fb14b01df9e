#!/usr/bin/env python3
"""bench.py — env-steps/sec of the batched Generals.io turn engine on B200.

Metric (BASELINE.json): env-steps/sec, 20x20 board, 2 players, fog on.  One env-step = one
ProcessTurn of one game + that game's P observation tensors, P packed legal masks, P rewards
and the done flag.  A bench "step" = one fused kernel launch over all B games
resident on the GPU.  The moves are the counter-based random-legal-move policy's,
recorded once in an untimed rollout and replayed, so that the timed region is the turn engine
itself and its inputs (that step's actions) are already resident where the arm reads them.

  python bench.py [--gpus N] [--steps K] [--warmup W]        our arm (CUDA, through the C ABI)
  python bench.py --impl reference ...                       CPU arm: the oracle restatement of
                                                             the Go engine on the host cores

`value`  : device-resident run (recorded actions in HBM, outputs stay in HBM).
`e2e`    : the same rollout replayed through the C ABI with HOST buffers every step: that
           step's actions copied host->device from pinned memory, reward/done/winner/error
           planes copied device->host (observations and masks stay in HBM for an on-GPU learner).
`e2e_host_obs`: additionally copies every observation tensor and mask to the host (PCIe-bound).
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
    """dram bytes per launch from the committed ncu --set full capture, if any."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f)
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


def time_oracle(games, steps, warmup, threads=0, with_readouts=True):
    """The CPU arm: C restatement of the Go engine (oracle/), one game per job, all host threads.
    Bounded sample of the same workload: `games` 20x20x2p games, seeds BASE_SEED+i, the same
    recorded random-legal-move actions replayed from host memory."""
    from generalsreinforcementlearning_b200.engine import BatchedEngine, make_config

    lib = load_oracle()
    cores = os.cpu_count() or 1
    e = BatchedEngine(lib, make_config(lib, num_envs=games, width=W, height=H, num_players=P, host_threads=threads))
    seeds = np.arange(games, dtype=np.int64) + BASE_SEED
    e.reset_seeded(seeds)
    rec = []
    for _ in range(warmup + steps):  # untimed: record the policy's moves
        a = e.sample_actions(POLICY_SEED)
        rec.append(a)
        e.step(a)
    e.reset_seeded(seeds)
    out = e.alloc_outputs_host()
    outs = e.outputs(obs=out["obs"], mask_bits=out["mask_bits"], reward=out["reward"], done=out["done"])
    run = (lambda a: e.step_fused(a, outs)) if with_readouts else (lambda a: e.step(a))
    for t in range(warmup):
        run(rec[t])
    s0 = int(e.stats()[0])
    t0 = time.perf_counter()
    for t in range(warmup, warmup + steps):
        run(rec[t])
    dt = time.perf_counter() - t0
    done_steps = int(e.stats()[0]) - s0
    e.close()
    return done_steps / dt, dt, cores if threads == 0 else threads, done_steps


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    games = 8192
    rate, dt, cores, n = time_oracle(games, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(1, args.steps), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int64+f32", "data": "synthetic",
        "config": {"workload": "20x20 2p fog-on random-legal-move rollouts (BASELINE headline config; seeds 12345+i; "
                               "the counter-based policy's moves recorded once and replayed from host memory)",
                   "games_per_step": games, "players": P, "board": [W, H], "episode_cap": 500,
                   "parallelism": f"one game per job over {cores} host threads (bounded sample of the GPU arm's workload)"},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{games} games x {args.steps} turns ({n} env-steps) with observation, mask and reward "
                                   "generation; C restatement of the Go engine (Go toolchain unavailable)"},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def gym_contract_rate(envs=65536, board=15, steps=60):
    """Extra, informational: the generals_gym contract through GeneralsVecEnv at the reference's default
    board (15x15) — one grl_gym_step launch per step, the reference's random agent (grl_gym_sample) as the policy, episode
    ends spread over time so that every step re-seeds envs on the device (grl_gym_autoreset).  Timed with CUDA events on
    torch's stream (the env's stream); never allowed to break the headline line."""
    try:
        import torch

        from generalsreinforcementlearning_b200.gym_env import GeneralsVecEnv

        env = GeneralsVecEnv(envs, board, board, max_turns=500, seed=12345, auto_reset="device")
        env.reset()
        env._calls.copy_(torch.randint(0, env.max_turns, (envs,), device=env._calls.device, dtype=torch.int32))
        for _ in range(5):
            env.step(env.sample_actions())
        torch.cuda.synchronize()
        l0 = env.engine.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            env.step(env.sample_actions())
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        launches = (env.engine.launch_count() - l0) / steps
        env.close()
        return {"value": envs / (ms * 1e-3), "unit": "env-steps/s", "envs": envs, "board": [board, board], "ms_per_vector_step": ms,
                "launches_per_step": launches,
                "what": "GeneralsVecEnv.step (one grl_gym_step launch) + random agent (grl_gym_sample) + device-side auto-reset of "
                        "the ~130 envs whose episode ends each step (grl_gym_autoreset); reference gym path: 12 steps/s per env"}
    except Exception as exc:  # noqa: BLE001 - informational figure only
        return {"unavailable": repr(exc)[:200]}


def run_cuda(args):
    import torch
    import torch.distributed as dist

    from generalsreinforcementlearning_b200 import _abi, load_library
    from generalsreinforcementlearning_b200.engine import BatchedEngine, make_config

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
    if Wm + K + 2 > 500:
        raise SystemExit("warmup+steps must stay below the 500-turn episode cap")
    e = BatchedEngine(lib, make_config(lib, num_envs=B, width=W, height=H, num_players=P, device=local,
                                       env_id_base=rank * B, host_threads=0))
    stream = torch.cuda.Stream(device=dev)  # the kernel's stream: CUDA events are recorded on it
    torch.cuda.set_stream(stream)
    e.set_stream(stream.cuda_stream)
    seeds = np.arange(B, dtype=np.int64) + BASE_SEED + rank * B

    obs = torch.empty((B, P, 9, H, W), dtype=torch.float32, device=dev)
    mask = torch.empty((B, P, e.mask_words), dtype=torch.int32, device=dev)
    reward = torch.empty((B, P), dtype=torch.float32, device=dev)
    done = torch.empty(B, dtype=torch.uint8, device=dev)
    winner = torch.empty(B, dtype=torch.int8, device=dev)
    serr = torch.empty(B, dtype=torch.uint8, device=dev)
    outs = e.outputs(obs=obs, mask_bits=mask, reward=reward, done=done)

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

    # ---------------- record the rollout's actions once (untimed) ---------------------------
    A = e.A
    e.reset_seeded(seeds)
    drec = torch.empty((Wm + K, B, A, 8), dtype=torch.uint8, device=dev)   # HBM-resident inputs
    rec = torch.empty((Wm + K, B, A, 8), dtype=torch.uint8).pin_memory()    # host copy for the e2e arm
    for t in range(Wm + K):
        e.sample_actions(POLICY_SEED, drec[t])
        e.step_fused(drec[t], e.outputs(done=done))
    rec.copy_(drec)
    torch.cuda.synchronize()

    # ---------------- device-resident arm -------------------------------------------------
    e.reset_seeded(seeds)
    for t in range(Wm):
        e.step_fused(drec[t], outs)
    torch.cuda.synchronize()
    steps_before = int(e.stats()[0])
    launches_before = e.launch_count()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    for t in range(Wm, Wm + K):
        e.step_fused(drec[t], outs)
    ev1.record(stream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = e.launch_count() - launches_before
    clocks = sampler.stop() if rank == 0 else None
    env_steps = int(e.stats()[0]) - steps_before
    ms_max = max_over_ranks(ms)
    total_steps = sum_over_ranks(env_steps)
    value = total_steps / (ms_max * 1e-3)
    del drec
    if args.quick:
        if rank == 0:
            print(json.dumps({"value": value, "ms_per_step": ms_max / K, "lib": os.environ.get("GRL_LIB_PATH", "default"),
                              "clocks": clocks}))
        e.close()
        return 0

    # ---------------- end-to-end arm: host actions in, host results out, every step ----------
    h_reward = torch.empty((B, P), dtype=torch.float32).pin_memory()
    h_done = torch.empty(B, dtype=torch.uint8).pin_memory()
    h_winner = torch.empty(B, dtype=torch.int8).pin_memory()
    h_err = torch.empty(B, dtype=torch.uint8).pin_memory()

    def e2e_run(full_obs):
        e.reset_seeded(seeds)
        if full_obs:
            h_obs = torch.empty((B, P, 9, H, W), dtype=torch.float32).pin_memory()
            h_mask = torch.empty((B, P, e.mask_words), dtype=torch.int32).pin_memory()
            o = e.outputs(obs=h_obs, mask_bits=h_mask, reward=h_reward, done=h_done, winner=h_winner, step_error=h_err)
            d2h = h_obs.numel() * 4 + h_mask.numel() * 4
        else:
            o = e.outputs(obs=obs, mask_bits=mask, reward=h_reward, done=h_done, winner=h_winner, step_error=h_err)
            d2h = 0
        d2h += h_reward.numel() * 4 + h_done.numel() + h_winner.numel() + h_err.numel()
        h2d = B * A * 8
        steps_k = K if not full_obs else min(K, 8)
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
        return sum_over_ranks(n) / dt_max, h2d, d2h, steps_k

    e2e_val, h2d, d2h, _ = e2e_run(False)
    e2e_full, h2d_f, d2h_f, k_full = e2e_run(True)
    # the replay must reproduce the device-resident rollout's work (same trajectories)
    e.close()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    alg = algorithmic_bytes_per_env_step(W, H, P)
    peak, peak_src = measured_peak_gbs()
    per_launch_bytes = alg["total"] * (env_steps / K)
    kernel_ms = ms / K
    achieved = per_launch_bytes / (kernel_ms * 1e-3) / 1e9
    traffic = recorded_traffic()

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        rate, dt, cores, n = time_oracle(4096, 12, 2)          # probe
        # bounded sample: ~15 s of CPU work on this box, the same 400-turn rollouts
        steps_c = 400
        games = int(min(32768, max(2048, (rate * 15.0 / steps_c) // 1024 * 1024)))
        rate, dt, cores, n = time_oracle(games, steps_c, 2)
        rate_bare, dt_bare, _, n_bare = time_oracle(8192, 100, 2, with_readouts=False)
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
               "value_turn_only": rate_bare,  # ProcessTurn alone, no observation / mask generation
               "sample": f"{games} games x {steps_c} turns ({n} env-steps, {dt:.1f} s) incl. observation/mask/reward; "
                         "C restatement of the Go engine (oracle/), Go toolchain unavailable"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm,
        "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u16/u32 bitmask + f32 planes", "data": "synthetic",
        "config": {"workload": "20x20 2p fog-on random-legal-move rollouts (BASELINE headline config; seeds 12345+i; "
                               "the counter-based policy's moves recorded once and replayed from HBM)",
                   "games_per_gpu": B, "players": P, "board": [W, H], "episode_cap": 500,
                   "cache": f"working set {(alg['total'] * B) / 1e6:.0f} MB per step > 126 MB L2 (no flush needed)",
                   "parallelism": f"games sharded by env index over {world} GPU(s), no collective on the step path"},
        "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "what": "per step through grl_step_fused with HOST buffers: actions copied host->device from pinned memory, "
                        "reward/done/winner/step_error delivered to pinned host memory (written in place by the kernel); "
                        "observation and mask planes stay in HBM for an on-GPU learner"},
        "e2e_host_obs": {"value": e2e_full, "unit": UNIT, "h2d_bytes_per_step": h2d_f, "d2h_bytes_per_step": d2h_f,
                         "steps": k_full, "what": "as e2e plus every observation tensor and mask copied to the host"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "peak_source": peak_src, "bytes_per_env_step": alg["total"],
                     "kernel": "grl_turn_kernel<2,20,20,true,true>", "kernel_ms": kernel_ms,
                     "traffic": (traffic or {}).get("dram_bytes_per_launch") if traffic else None,
                     "frac_of_nominal_8TBs": achieved / 8000.0},
        "clocks": clocks,
        "env_steps_per_launch": env_steps / K,
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
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--games", type=int, default=GAMES_PER_GPU, help="games resident per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--quick", action="store_true", help="device-resident arm only (kernel experiments)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_cuda(args)


if __name__ == "__main__":
    sys.exit(main())
