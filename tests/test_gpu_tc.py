"""Tensor-core building blocks (TMA + tcgen05 + TMEM) against numpy."""
import numpy as np
import pytest

from tests.util import make_model

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def model():
    return make_model(seed=11, bits=9, mode="RAW")[0]


@pytest.mark.parametrize("N", [16, 32, 64])
def test_tc_gemm_self_test(model, N):
    rng = np.random.default_rng(N)
    A = rng.uniform(-1, 1, size=(128, 512)).astype(np.float16)
    W = rng.uniform(-0.05, 0.05, size=(N, 512)).astype(np.float16)
    got = model.debug_tc_gemm(A, W)
    want = A.astype(np.float32) @ W.astype(np.float32).T
    np.testing.assert_allclose(got, want, rtol=0, atol=2e-4)
