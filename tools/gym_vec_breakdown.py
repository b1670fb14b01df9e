"""Where a GeneralsVecEnv step's time goes: host enqueue time against device time, for the whole step and for each of its
three library calls alone (grl_gym_sample, grl_gym_step, grl_gym_autoreset).  Prints one JSON line per board."""
import json, os, sys, time
sys.path.insert(0, os.getcwd())
import torch
from generalsreinforcementlearning_b200.gym_env import GeneralsVecEnv


def timed(fn, steps):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    return {"host_enqueue_ms": (t1 - t0) * 1e3 / steps, "device_ms": e0.elapsed_time(e1) / steps}


def main():
    B = int(os.environ.get("GRL_B", 65536))
    for W in (15, 20, 10):
        for mode in ("device", "host", "host_reset"):
            env = GeneralsVecEnv(B, W, W, max_turns=500, seed=3, auto_reset=mode)
            env.reset()
            env._calls.copy_(torch.randint(0, env.max_turns, (B,), device=env._calls.device, dtype=torch.int32))
            for _ in range(8):
                env.step(env.sample_actions())
            out = {"board": W, "envs": B, "auto_reset": mode}
            out["vector_step"] = timed(lambda: env.step(env.sample_actions()), 60)
            if mode == "device":
                a = env.sample_actions()
                out["sample_only"] = timed(lambda: env.sample_actions(), 60)
                o = env._out[0]
                def ar():
                    env.engine.gym_autoreset(env.max_turns, env._base_seed, terminated=o["terminated"], truncated=o["truncated"],
                                             episode=env._episode_dev, turns=env._turns, calls=env._calls, obs=env._obs,
                                             mask=env._mask, stats=env._stats, final_obs=env._final_obs)
                out["autoreset_only"] = timed(ar, 60)
                def st():
                    env._opp_draws += 1
                    env.engine.gym_step(env.max_turns, env._opp_draws, action=a, opponent_action=None,
                                        obs=env._obs, mask=env._mask, stats=env._stats, actions=env._actions,
                                        prev_stats=env._prev_stats, turns=env._turns, calls=env._calls, reward=o["reward"],
                                        terminated=o["terminated"], truncated=o["truncated"], valid=o["valid"], done=env._done,
                                        winner=o["winner"], step_error=o["step_error"], n_finished=env._nfin)
                out["step_only"] = timed(st, 60)
            print(json.dumps(out), flush=True)
            env.close()


if __name__ == "__main__":
    main()
