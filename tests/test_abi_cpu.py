"""CPU-side checks of the product library: it loads, exports exactly what include/grlcuda.h
declares, and its host-only entry points (config, map generation) agree with the oracle.
No CUDA compute is attempted here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from generalsreinforcementlearning_b200 import _abi, build as grl_build
from generalsreinforcementlearning_b200._abi import BoundLibrary, Config

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "grlcuda.h")


def header_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(grl_[a-z_0-9]+)\s*\(", text)))


@pytest.fixture(scope="module")
def product_lib():
    path = grl_build.build()
    return BoundLibrary(path, "grl_")


def test_header_and_binding_agree():
    declared = header_functions()
    bound = sorted("grl_" + n for n in _abi.ABI_FUNCTIONS)
    assert declared == bound


def test_product_exports_every_declared_symbol(product_lib):
    for name in header_functions():
        assert hasattr(product_lib.cdll, name), name
    assert product_lib.abi_version() == 3


def test_oracle_exports_the_same_abi(oracle_lib):
    for name in header_functions():
        assert hasattr(oracle_lib.cdll, "grlo_" + name[len("grl_"):]), name


def test_struct_sizes_match_header():
    # grl_action is 8 bytes; grl_config is 16 int32 + 11 floats
    assert _abi.ACTION_DTYPE.itemsize == 8
    assert C.sizeof(Config) == 16 * 4 + 11 * 4
    assert C.sizeof(_abi.StepOutputs) == 8 * 8
    assert C.sizeof(_abi.StatePlanes) == 14 * 8


def test_default_config_matches_reference_and_oracle(product_lib, oracle_lib):
    a, b = Config(), Config()
    assert product_lib.default_config(C.byref(a)) == 0
    assert oracle_lib.default_config(C.byref(b)) == 0
    assert bytes(a) == bytes(b)
    # internal/config/config.go:198-209, experience/rewards.go:23-37
    assert (a.city_ratio, a.city_start_army, a.min_general_spacing) == (20, 40, 5)
    assert (a.production_general, a.production_city, a.production_normal, a.normal_growth_interval) == (1, 1, 1, 25)
    assert a.fog_of_war == 1
    assert abs(a.reward.army_advantage - 0.05) < 1e-9 and a.reward.win_game == 1.0


def test_bad_config_is_rejected_without_touching_cuda(product_lib):
    cfg = Config()
    product_lib.default_config(C.byref(cfg))
    cfg.width = 33
    h = C.c_void_p()
    assert product_lib.create(C.byref(cfg), C.byref(h)) == -1
    assert b"width" in product_lib.last_error()


@pytest.mark.parametrize("W,H,P", [(5, 5, 2), (10, 10, 2), (15, 15, 2), (20, 20, 2), (20, 20, 4), (25, 25, 4),
                                    (32, 32, 8), (8, 8, 2), (7, 13, 3)])
def test_product_mapgen_matches_oracle(product_lib, oracle_lib, W, H, P):
    """Host map generation in the product (grl_mapgen.cpp) and in the oracle are written
    separately; both must give the reference's seeded maps."""
    cfg = Config()
    product_lib.default_config(C.byref(cfg))
    cfg.width, cfg.height, cfg.num_players = W, H, P
    N = W * H
    seeds = [12345, 12346, 0, -7, 1, 42, 2**31 - 1, 2**31, 987654321987] + list(range(100, 140))
    for seed in seeds:
        outs = []
        for lib in (product_lib, oracle_lib):
            o, a, t = (np.zeros(N, np.int32) for _ in range(3))
            st = lib.mapgen(C.byref(cfg), seed, o.ctypes.data, a.ctypes.data, t.ctypes.data)
            outs.append((st, o, a, t))
        assert outs[0][0] == outs[1][0], seed
        if outs[0][0] == 0:
            for x, y in zip(outs[0][1:], outs[1][1:]):
                assert np.array_equal(x, y), seed


@pytest.mark.parametrize("W,H,P,fog", [(20, 20, 2, 1), (15, 15, 2, 1), (10, 10, 2, 1), (20, 20, 4, 1), (7, 13, 3, 1),
                                        (5, 5, 2, 1), (32, 32, 8, 1), (15, 15, 2, 0), (1, 1, 1, 1)])
def test_packed_observation_records_expand_to_state_to_tensor(product_lib, oracle_lib, W, H, P, fog):
    """grl_step_outputs.obs_packed + grl_expand_obs (the host-delivery path): the product's vectorised host expander
    and the oracle's tile-by-tile one both turn the oracle's packed records into exactly the float32 tensors the same
    step wrote (Serializer.StateToTensor, serializer.go:37-109), bit for bit."""
    from generalsreinforcementlearning_b200.engine import BatchedEngine, make_config

    B = 24
    cfg = make_config(oracle_lib, num_envs=B, width=W, height=H, num_players=P, max_actions=max(2, P), host_threads=1,
                      fog_of_war=fog)
    e = BatchedEngine(oracle_lib, cfg)
    if W * H >= 25:
        e.reset_seeded(np.arange(B, dtype=np.int64) + 777)
    else:   # no room for generals: a hand-made board
        e.reset_boards(np.zeros((B, 1), np.int32), np.full((B, 1), 3, np.int32), np.ones((B, 1), np.int32))
    RW = e.packed_words
    N = W * H
    assert RW == product_lib.obs_packed_words(W, H, P) == (((2 * P + 2) * ((N + 31) // 32) + ((N + 7) & ~7) // 2 + 3) & ~3)
    out = e.alloc_outputs_host()
    packed = np.zeros((B, RW), np.uint32)
    for t in range(60):
        e.step_fused(None, e.outputs(obs_packed=packed, **out), _abi.STEP_FLAG_RANDOM_POLICY, 9)
        if t % 6:
            continue
        for lib in (product_lib, oracle_lib):
            got = np.full((B, P, 9, H, W), np.nan, np.float32)
            assert lib.expand_obs(W, H, P, packed.ctypes.data, B, got.ctypes.data, 3 if lib is product_lib else 1) == 0
            assert np.array_equal(got.view(np.uint32), out["obs"].view(np.uint32)), (lib.prefix, t)
    # armies at and above the clip (serializer.go:84-88) and the largest the plane holds
    st = e.get_state()
    st["army"][:, 0] = [999, 1000, 1001, 65535, 0, 1] * (B // 6)
    e.set_state(st)
    e.observe(e.outputs(obs=out["obs"], obs_packed=packed))
    got = product_lib_expand(product_lib, W, H, P, packed)
    assert np.array_equal(got.view(np.uint32), out["obs"].view(np.uint32))
    e.close()


def product_lib_expand(lib, W, H, P, packed):
    got = np.empty((packed.shape[0], P, 9, H, W), np.float32)
    assert lib.expand_obs(W, H, P, packed.ctypes.data, packed.shape[0], got.ctypes.data, 0) == 0
    return got


def test_graft_entry_build_runs():
    """The driver's "does it build" check: __graft_entry__.build() compiles whatever is stale (nothing, normally), loads
    the library and compares its ABI version with the header's."""
    import importlib
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if root not in sys.path:
        sys.path.insert(0, root)
    entry = importlib.import_module("__graft_entry__")
    entry.build()
