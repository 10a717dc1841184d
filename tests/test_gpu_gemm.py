"""tcgen05 GEMM kernel (TMA + TMEM, kind::tf32) against float64 NumPy, through psm_debug_gemm."""
import numpy as np
import pytest

import psm_b200
from psm_b200 import _capi
from helpers import rel_l2

pytestmark = pytest.mark.gpu

SHAPES = [  # M, N, K, splits
    (128, 128, 32, 1),        # one k-block, one stage
    (128, 128, 256, 1),       # wraps the 3-stage ring
    (128, 64, 96, 1),         # BN = 64 path
    (256, 512, 512, 1),       # several M and N tiles (Dense layer shape)
    (128, 128, 2048, 8),      # split-K
    (128, 192, 640, 3),       # BN = 64 with N = 3 tiles, uneven split (20 k-blocks over 3)
    (128, 128, 32768, 128),   # the PCA projection shape of configs[1]
    (256, 16384, 128, 1),     # the PCA inverse shape
]


@pytest.mark.parametrize("M,N,K,splits", SHAPES)
@pytest.mark.parametrize("mode", [_capi.GEMM_TC_TF32, _capi.GEMM_TC_3XTF32, _capi.GEMM_FP32_SIMT])
def test_gemm_matches_numpy(mode, M, N, K, splits):
    if mode == _capi.GEMM_FP32_SIMT and (N % 64 or K % 16):
        pytest.skip("shape not supported by the SIMT kernel")
    rng = np.random.default_rng(M + N + K)
    A = rng.standard_normal((M, K)).astype(np.float32)
    B = rng.standard_normal((N, K)).astype(np.float32)
    C = psm_b200.debug_gemm(A, B, mode=mode, splits=splits).astype(np.float64).sum(axis=0)
    ref = A.astype(np.float64) @ B.astype(np.float64).T
    assert not np.isnan(C).any()
    err = rel_l2(C, ref)
    tol = {_capi.GEMM_TC_TF32: 2e-3, _capi.GEMM_TC_3XTF32: 1.5e-5, _capi.GEMM_FP32_SIMT: 3e-6}[mode]
    assert err < tol, err


def test_3xtf32_is_exact_on_tf32_representable_inputs():
    """Inputs with <= 10 mantissa bits and small integer products: every mode must be exact."""
    rng = np.random.default_rng(0)
    A = rng.integers(-8, 9, size=(128, 256)).astype(np.float32)
    B = rng.integers(-8, 9, size=(128, 256)).astype(np.float32)
    ref = A.astype(np.float64) @ B.astype(np.float64).T
    for mode in (_capi.GEMM_TC_TF32, _capi.GEMM_TC_3XTF32):
        C = psm_b200.debug_gemm(A, B, mode=mode)[0]
        np.testing.assert_array_equal(C.astype(np.float64), ref)
