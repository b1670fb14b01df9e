"""The reference's known-answer tests (tests/kats.py) through libgrlcuda.so on the GPU."""
import pytest

import kats

KATS = [getattr(kats, n) for n in sorted(dir(kats)) if n.startswith("kat_")]

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kat", KATS, ids=lambda f: f.__name__)
def test_kat(cuda_lib, kat):
    kat(cuda_lib)
