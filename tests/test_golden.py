"""Committed REGRESSION fixtures (tests/golden/rollout_*.npz, made by tests/tools/make_golden.py): fixed-seed maps and
120-turn trajectories recorded from this repository's own oracle.  They are not reference data — the Go engine cannot
run here — and pin nothing against the reference; they freeze behaviour, so that a later edit of the oracle OR of the
kernels that moves any trajectory is noticed.  (Reference pins live elsewhere: tests/kats.py — the reference's own Go
tests —, the seeded mapgen counts, and tests/golden/gym_ref/ — outputs of the reference's own Python client.)
The CPU leg checks the oracle against the fixtures; the GPU leg runs the same rollouts through libgrlcuda.so and must
reproduce every recorded plane bit for bit."""
import glob
import os

import numpy as np
import pytest

from generalsreinforcementlearning_b200 import _abi
from helpers import new_engine

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "rollout_*.npz")))
SEED, POLICY_SEED = 12345, 2024


def _check(lib, path):
    g = np.load(path)
    name = os.path.basename(path)[len("rollout_"):-len(".npz")]
    dims, P = name.rsplit("x", 1)
    W, H = (int(v) for v in dims.split("x"))
    P = int(P[:-1])
    T, B = g["hash"].shape
    e = new_engine(lib, W, H, P, B)
    e.reset_seeded(np.arange(B, dtype=np.int64) + SEED)
    st = e.get_state()
    assert np.array_equal(st["owner"], g["boards_owner"]), "mapgen owner plane"
    assert np.array_equal(st["army"], g["boards_army"]), "mapgen army plane"
    assert np.array_equal(st["type"], g["boards_type"]), "mapgen type plane"
    out = e.alloc_outputs_host()
    for t in range(T):
        e.step_fused(None, e.outputs(**out), _abi.STEP_FLAG_RANDOM_POLICY, POLICY_SEED)
        assert np.array_equal(e.state_hash(), g["hash"][t]), f"{name} turn {t}: state digest"
        assert np.array_equal(out["reward"].view(np.uint32), g["reward"][t]), f"{name} turn {t}: reward bits"
        assert np.array_equal(out["done"], g["done"][t]), f"{name} turn {t}: done"
        assert np.array_equal(out["winner"], g["winner"][t]), f"{name} turn {t}: winner"
        assert np.array_equal(out["step_error"], g["step_error"][t]), f"{name} turn {t}: step_error"
        assert np.array_equal(out["mask_bits"], g["mask"][t]), f"{name} turn {t}: mask"
        assert np.array_equal(e.buffer_hash(out["obs"], 9 * W * H, B * P), g["obs_hash"][t]), f"{name} turn {t}: obs"
    assert np.array_equal(out["obs"].view(np.uint32), g["obs_last"].view(np.uint32))
    e.close()


def test_fixtures_exist():
    assert len(GOLDEN) >= 6


@pytest.mark.parametrize("path", GOLDEN, ids=os.path.basename)
def test_oracle_reproduces_golden(oracle_lib, path):
    _check(oracle_lib, path)


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLDEN, ids=os.path.basename)
def test_cuda_reproduces_golden(cuda_lib, path):
    _check(cuda_lib, path)
