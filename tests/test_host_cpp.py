"""The C++ host mirror of the reference's `internal/game` surface (generalsreinforcementlearning_b200/host).

tests/cpp/host_test.cpp transliterates the reference's Go tests for the turn path (engine_test.go,
action_mask_test.go, core/action_test.go, core/movement_test.go, experience/buffer_test.go, experience/collector_test.go) against `grl::game::Engine`.  Without a GPU
the binary binds the CPU oracle's copy of the C ABI, which checks the host layer's own logic (action packing,
error synthesis, plane slicing, the turn barrier, the renderer); on the GPU box it binds libgrlcuda.so — the
product path."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "generalsreinforcementlearning_b200", "host")
BIN = os.path.join(HOST, "host_test")


def _build():
    subprocess.check_call(["make", "-s", "-C", HOST])
    assert os.path.exists(BIN)


def _run(lib, prefix):
    _build()
    proc = subprocess.run([BIN, "--lib", lib, "--prefix", prefix], stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                          text=True, timeout=600)
    assert proc.returncode == 0, proc.stdout[-4000:]
    assert " 0 failures" in proc.stdout, proc.stdout[-2000:]
    return proc.stdout


def test_host_mirror_logic_on_oracle_binding():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
    out = _run(os.path.join(ROOT, "oracle", "libgrloracle.so"), "grlo_")
    assert "ok   TestEngine_PlayerEliminationAndTileTurnover" in out


def test_host_library_exports():
    """libgrlhost.so loads and exports the C++ surface (mangled names of the mirrored methods)."""
    _build()
    syms = subprocess.run(["nm", "-DC", os.path.join(HOST, "libgrlhost.so")], stdout=subprocess.PIPE, text=True).stdout
    for name in ("grl::game::NewEngine(", "grl::game::Engine::Step(", "grl::game::Engine::GameState()",
                 "grl::game::Engine::IsGameOver()", "grl::game::Engine::GetWinner()",
                 "grl::game::Engine::GetLegalActionMask(int)", "grl::game::Engine::ComputePlayerVisibility(int)",
                 "grl::game::Engine::GetChangedTiles()", "grl::game::Engine::GetVisibilityChangedTiles()",
                 "grl::game::Engine::Board[abi:cxx11](int)", "grl::game::EnginePool::StepAll(",
                 "grl::Library::Default()"):
        assert name in syms, name


def test_host_default_binding_fails_loudly_without_the_cuda_library(tmp_path):
    """No CPU engine stands behind the host layer: with no libgrlcuda.so to bind, the default binding is an
    error.  (Run from a copy of the host library so the in-tree CUDA library is not next to it.)"""
    _build()
    import shutil
    shutil.copy(os.path.join(HOST, "libgrlhost.so"), tmp_path / "libgrlhost.so")
    shutil.copy(BIN, tmp_path / "host_test")
    env = dict(os.environ, GRLCUDA_LIB=str(tmp_path / "missing.so"), LD_LIBRARY_PATH=str(tmp_path))
    proc = subprocess.run([str(tmp_path / "host_test")], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                          env=env, cwd=tmp_path)
    assert proc.returncode == 2 and "libgrlcuda.so not found" in proc.stdout, proc.stdout


@pytest.mark.gpu
def test_reference_go_tests_through_cuda_library():
    out = _run(os.path.join(ROOT, "generalsreinforcementlearning_b200", "csrc", "libgrlcuda.so"), "grl_")
    assert "bound" in out and "libgrlcuda.so" in out
