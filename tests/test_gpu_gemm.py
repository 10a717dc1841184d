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


STACKS = [  # M, dims, clusters cap
    (121, [128, 512, 512, 512, 128], 0),     # MLP_small at configs[1] (UTL:437-439): one 128-row tile
    (124, [45, 512, 512, 512, 48], 0),       # the shipped thesis model (PMP:121-134): widths padded to 128
    (441, [128, 512, 512, 512, 128], 0),     # configs[4]: four row tiles, 32 output tiles per layer over 16 clusters
    (300, [100, 256, 640, 70], 3),           # uneven widths, more tiles than clusters
    (64, [128, 128], 0),                     # a single linear layer
]


@pytest.mark.parametrize("M,dims,clusters", STACKS)
@pytest.mark.parametrize("mode", [_capi.GEMM_TC_3XTF32, _capi.GEMM_TC_TF32])
def test_dense_stack_matches_numpy(mode, M, dims, clusters):
    rng = np.random.default_rng(M + sum(dims))
    x = rng.standard_normal((M, dims[0])).astype(np.float32)
    ks = [(rng.standard_normal((dims[i], dims[i + 1])) / np.sqrt(dims[i])).astype(np.float32) for i in range(len(dims) - 1)]
    bs = [(0.1 * rng.standard_normal(dims[i + 1])).astype(np.float32) for i in range(len(dims) - 1)]
    out = psm_b200.debug_dense_stack(x, ks, bs, mode=mode, clusters=clusters)
    a = x.astype(np.float64)
    for i, (k, b) in enumerate(zip(ks, bs)):
        a = a @ k.astype(np.float64) + b.astype(np.float64)
        if i < len(ks) - 1:
            a = np.maximum(a, 0.0)
    assert not np.isnan(out).any()
    err = rel_l2(out.astype(np.float64), a)
    assert err < (2e-5 if mode == _capi.GEMM_TC_3XTF32 else 5e-3), err


def test_dense_stack_is_deterministic():
    rng = np.random.default_rng(5)
    dims = [128, 512, 512, 128]
    x = rng.standard_normal((200, dims[0])).astype(np.float32)
    ks = [(rng.standard_normal((dims[i], dims[i + 1])) / np.sqrt(dims[i])).astype(np.float32) for i in range(3)]
    bs = [np.zeros(dims[i + 1], np.float32) for i in range(3)]
    a = psm_b200.debug_dense_stack(x, ks, bs)
    b = psm_b200.debug_dense_stack(x, ks, bs, clusters=2)
    np.testing.assert_array_equal(a, b)


def test_tf32_operand_truncation(monkeypatch):
    """The 3xTF32 kernels feed the raw FP32 tile as the `hi` operand and rely on the tensor core dropping the low 13
    mantissa bits (truncation).  If the hardware rounded instead, hi + lo != x: the result must be bit-identical to
    the variant whose converters store the explicitly masked hi tile."""
    rng = np.random.default_rng(11)
    A = rng.standard_normal((128, 512)).astype(np.float32)
    B = rng.standard_normal((128, 512)).astype(np.float32)
    monkeypatch.delenv('PSM_TF32_MASK_HI', raising=False)
    c_raw = psm_b200.debug_gemm(A, B, mode=_capi.GEMM_TC_3XTF32)[0]
    monkeypatch.setenv('PSM_TF32_MASK_HI', '1')
    c_mask = psm_b200.debug_gemm(A, B, mode=_capi.GEMM_TC_3XTF32)[0]
    np.testing.assert_array_equal(c_raw, c_mask)
