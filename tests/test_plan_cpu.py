"""Host-side plan compiler (C++ in libpsm_b200.so, no GPU) against the oracle's literal loops."""
import warnings

import numpy as np
import pytest

import psm_b200
from psm_b200 import _capi
from oracle import assemble as oasm
from oracle.pipeline import DeltasOracle, GradPOracle, ThesisOracle


def oracle_plan(variant, H, W, ov):
    if variant == 'thesis':
        o = ThesisOracle(None)
        assert o.avance == ov
    else:
        o = (DeltasOracle(None, overlap=ov) if variant == 'deltaU_to_deltaP' else GradPOracle(None, avance=ov))
    o.grid_shape_y, o.grid_shape_x = H, W
    return o.block_plan()


def disc_mask(H, W, cy, cx, r):
    yy, xx = np.mgrid[0:H, 0:W]
    return (((yy - cy) ** 2 + (xx - cx) ** 2) > r * r).astype(np.uint8)


def eval_plan(plan, blocks, mask, ref_bc=0.0):
    """NumPy evaluation of the compiled plan: task means -> sequential recurrence -> owner gather."""
    B, F = plan['n_blocks'], plan['n_fields']
    org = plan['origins']
    means = np.zeros(len(plan['tasks']))
    for t, (src, msk, ch, y0, y1, x0, x1, cnt) in enumerate(plan['tasks']):
        if msk >= 0:
            m = mask[org[msk, 0] + y0:org[msk, 0] + y1, org[msk, 1] + x0:org[msk, 1] + x1] != 0
        else:                                                   # plain mean, no mask (PMP:437)
            m = np.ones((y1 - y0, x1 - x0), bool)
        assert m.sum() == cnt
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", category=RuntimeWarning)
            means[t] = np.mean(blocks[src, ch, y0:y1, x0:x1][m])
    c = np.zeros((F, B))
    for f in range(F):
        for k in range(B):
            ta, tb, par, is_nan = plan['rec'][f, k]
            c[f, k] = means[ta] - ((means[tb] - c[f, par]) if tb >= 0 else ref_bc)
            assert bool(is_nan) == bool(np.isnan(c[f, k]))
    H, W = mask.shape
    ow = plan['owner']
    yy, xx = np.mgrid[0:H, 0:W]
    fields = []
    for f in range(F):
        fields.append(blocks[ow, f, yy - org[ow, 0], xx - org[ow, 1]] - c[f][ow])
    return c, fields


SIZES = [(240, 330), (300, 420), (397, 998), (129, 130), (500, 700), (1000, 1000)]


@pytest.mark.parametrize("H,W", SIZES)
@pytest.mark.parametrize("variant,ov", [('deltaU_to_deltaP', 32), ('U_to_gradP', 96), ('deltaU_to_deltaP', 13)])
def test_block_plan_and_owner_map_bit_exact(variant, ov, H, W):
    st = 128 - ov
    if (H - 128) % st == 0:
        pytest.skip("reference undefined (p_i == 0)")
    mask = np.ones((H, W), np.uint8)
    plan = psm_b200.compile_plan(variant, H, W, mask, overlap=ov)
    n_x, n_y, origins, il = oracle_plan(variant, H, W, ov)
    np.testing.assert_array_equal(plan['origins'], np.array(origins, dtype=np.int32))
    np.testing.assert_array_equal(plan['indices_list'], np.array(il, dtype=np.int32))
    ow = oasm.owner_map('deltas' if variant == 'deltaU_to_deltaP' else 'grad', il, n_x, n_y, 128, ov, W, H)
    np.testing.assert_array_equal(plan['owner'], ow)


def test_rejects_geometries_where_the_reference_is_undefined():
    mask = np.ones((224, 330), np.uint8)                   # (H - 128) % 96 == 0  ->  p_i == 0
    with pytest.raises(_capi.PsmError) as e:
        psm_b200.compile_plan('deltaU_to_deltaP', 224, 330, mask)
    assert e.value.code == _capi.PSM_ERR_GEOMETRY
    with pytest.raises(_capi.PsmError):                    # n_x == 0
        psm_b200.compile_plan('deltaU_to_deltaP', 300, 128, np.ones((300, 128), np.uint8))
    with pytest.raises(_capi.PsmError):                    # smaller than a block
        psm_b200.compile_plan('deltaU_to_deltaP', 100, 330, np.ones((100, 330), np.uint8))


CASES = [
    ('deltaU_to_deltaP', 32, 240, 330, None),
    ('deltaU_to_deltaP', 32, 300, 420, (150, 126, 20)),
    ('deltaU_to_deltaP', 32, 240, 330, (112, 170, 70)),     # empty bottom strip -> NaN-fallback branch
    ('deltaU_to_deltaP', 32, 240, 330, (112, 266, 75)),     # right-most column strip empty -> NaN chain
    ('deltaU_to_deltaP', 32, 500, 700, (250, 210, 80)),
    ('U_to_gradP', 96, 240, 340, None),
    ('U_to_gradP', 96, 240, 340, (120, 102, 20)),
    ('U_to_gradP', 96, 300, 421, (140, 200, 75)),
    ('thesis', 12, 260, 380, None),
    ('thesis', 12, 260, 380, (130, 120, 25)),
    ('thesis', 12, 400, 500, (200, 150, 90)),              # empty strips -> NaN BC_ups -> BC_alter branch
    ('thesis', 12, 130, 250, None),                         # n_y == 0: first and last row only
    ('thesis', 12, 500, 700, (250, 520, 60)),
    ('thesis', 12, 360, 128 + 116 * 2, None),               # p_j == 0: the -1 column's last-row write is empty
]


@pytest.mark.parametrize("variant,ov,H,W,disc", CASES)
def test_closed_form_recurrence_equals_sequential_assembly(variant, ov, H, W, disc):
    """c_k = m[a] - (m[b] - c[parent]) over the static plan reproduces SMC:221-350 / GRAD:282-361
    (offsets, NaN pattern and assembled field) on random blocks."""
    rng = np.random.default_rng(H * 1000 + W)
    mask = np.ones((H, W), np.uint8) if disc is None else disc_mask(H, W, *disc)
    plan = psm_b200.compile_plan(variant, H, W, mask, overlap=ov)
    B, F = plan['n_blocks'], plan['n_fields']
    blocks = rng.standard_normal((B, F, 128, 128))
    n_x, n_y, origins, il = oracle_plan(variant, H, W, ov)
    x_array = np.zeros((B, 128, 128, 3))
    for k, (y0, x0) in enumerate(origins):
        x_array[k, :, :, 2] = mask[y0:y0 + 128, x0:x0 + 128] * 0.37
    c, fields = eval_plan(plan, blocks, mask)
    for f in range(F):
        if variant == 'thesis':
            ref, offs, shift = oasm.assemble_thesis(blocks[:, 0], x_array, il, n_x, n_y, 128, ov, W, H, return_offsets=True)
            mine = fields[0]
            sh = np.mean(3 * mine[:, -1] - mine[:, -2]) / 3
        elif variant == 'deltaU_to_deltaP':
            ref, offs, shift = oasm.assemble_deltas(blocks[:, 0], x_array, il, n_x, n_y, 128, ov, W, H, return_offsets=True)
            mine = fields[0]
            sh = np.mean(3 * mine[:, -1] - mine[:, -2]) / 3
        else:
            name = ('dp_dx', 'dp_dy')[f]
            ref, offs, shift = oasm.assemble_gradp(name, blocks[:, f], x_array, il, n_x, n_y, 128, ov, W, H, return_offsets=True)
            ref = ref[0, :, :, 0]
            mine = fields[f]
            sh = (np.mean(3 * mine[:, 0] - mine[:, 1]) / 3) if f == 0 else (np.mean(3 * mine[1, :] - mine[2, :]) / 3)
        assert np.array_equal(np.isnan(offs), np.isnan(c[f]))
        np.testing.assert_allclose(c[f], offs, rtol=0, atol=1e-12, equal_nan=True)
        assert np.array_equal(np.isnan(ref), np.isnan(mine - sh))
        np.testing.assert_allclose(mine - sh, ref, rtol=0, atol=1e-11, equal_nan=True)
        # the shift as the kernels evaluate it: per-block runs of the two lines (psm_plan_shift_lines)
        acc, npx = 0.0, 0
        for (lf, blk, y0, y1, x0, x1, coef, n) in plan['lines']:
            if lf != f:
                continue
            assert (y1 - y0) * (x1 - x0) == n
            acc += coef * (blocks[blk, f, y0:y1, x0:x1].sum() - n * c[f][blk])
            npx += n
        L = H if (variant != 'U_to_gradP' or f == 0) else W
        assert npx == 2 * L
        np.testing.assert_allclose(acc / L / 3, sh, rtol=0, atol=1e-11, equal_nan=True)


def test_library_exports_every_declared_symbol():
    """Every function include/psm_b200.h declares is exported by the built library."""
    import os
    import re
    here = os.path.dirname(os.path.abspath(__file__))
    hdr = open(os.path.join(here, '..', 'include', 'psm_b200.h')).read()
    declared = set(re.findall(r'\b(psm_[a-z_0-9]+)\s*\(', hdr)) - {'psm_handle'}
    assert declared == set(_capi.SYMBOLS), declared ^ set(_capi.SYMBOLS)
    lib = _capi.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.psm_api_version() == 5


@pytest.mark.parametrize("H,W", [(260, 380), (130, 250), (397, 998), (1000, 1000), (360, 360)])
def test_thesis_block_plan_bit_exact(H, W):
    """PMP:303-332: right -> left plus the extra -1 block per row, B = (n_y + 2)(n_x + 2)."""
    plan = psm_b200.compile_plan('thesis', H, W, np.ones((H, W), np.uint8), overlap=12)
    n_x, n_y, origins, il = oracle_plan('thesis', H, W, 12)
    assert plan['n_blocks'] == (n_y + 2) * (n_x + 2) == len(origins)
    np.testing.assert_array_equal(plan['origins'], np.array(origins, dtype=np.int32))
    np.testing.assert_array_equal(plan['indices_list'], np.array(il, dtype=np.int32))


def _random_mask(rng, H, W):
    """Union of a few random discs / bars removed from the flow domain (some large enough to empty whole strips)."""
    mask = np.ones((H, W), np.uint8)
    yy, xx = np.mgrid[0:H, 0:W]
    for _ in range(int(rng.integers(0, 4))):
        cy, cx, r = rng.integers(0, H), rng.integers(0, W), rng.integers(5, max(6, min(H, W) // 3))
        mask[((yy - cy) ** 2 + (xx - cx) ** 2) <= r * r] = 0
    if rng.random() < 0.3:                                   # a bar: empties full-width / full-height strips
        if rng.random() < 0.5:
            y = int(rng.integers(0, H - 40)); mask[y:y + int(rng.integers(10, 40)), :] = 0
        else:
            x = int(rng.integers(0, W - 40)); mask[:, x:x + int(rng.integers(10, 40))] = 0
    return mask


@pytest.mark.parametrize("variant,ov", [('deltaU_to_deltaP', 32), ('U_to_gradP', 96), ('thesis', 12)])
def test_random_geometries_against_the_literal_loops(variant, ov):
    """Seeded sweep over grid sizes and obstacle masks: offsets, NaN pattern, assembled field and shift of the compiled
    plan equal the reference loop restated in oracle/assemble.py -- including masks that empty strips (NaN branches)."""
    rng = np.random.default_rng({'deltaU_to_deltaP': 101, 'U_to_gradP': 202, 'thesis': 303}[variant])
    st = 128 - ov
    done = 0
    while done < 12:
        H, W = int(rng.integers(130, 520)), int(rng.integers(260, 640))
        if variant != 'thesis' and (H - 128) % st == 0:
            continue
        mask = _random_mask(rng, H, W)
        try:
            plan = psm_b200.compile_plan(variant, H, W, mask, overlap=ov)
        except _capi.PsmError as e:                          # geometries the reference itself rejects (asserts, unbound names)
            assert e.code == _capi.PSM_ERR_GEOMETRY
            continue
        B, F = plan['n_blocks'], plan['n_fields']
        blocks = rng.standard_normal((B, F, 128, 128))
        n_x, n_y, origins, il = oracle_plan(variant, H, W, ov)
        x_array = np.zeros((B, 128, 128, 3))
        for k, (y0, x0) in enumerate(origins):
            x_array[k, :, :, 2] = mask[y0:y0 + 128, x0:x0 + 128] * 0.5
        c, fields = eval_plan(plan, blocks, mask)
        for f in range(F):
            with warnings.catch_warnings():
                warnings.simplefilter("ignore", category=RuntimeWarning)
                if variant == 'thesis':
                    ref, offs, _ = oasm.assemble_thesis(blocks[:, 0], x_array, il, n_x, n_y, 128, ov, W, H, return_offsets=True)
                elif variant == 'deltaU_to_deltaP':
                    ref, offs, _ = oasm.assemble_deltas(blocks[:, 0], x_array, il, n_x, n_y, 128, ov, W, H, return_offsets=True)
                else:
                    ref, offs, _ = oasm.assemble_gradp(('dp_dx', 'dp_dy')[f], blocks[:, f], x_array, il, n_x, n_y, 128, ov, W, H,
                                                       return_offsets=True)
                    ref = ref[0, :, :, 0]
                mine = fields[f]
                if variant == 'U_to_gradP':
                    sh = (np.mean(3 * mine[:, 0] - mine[:, 1]) / 3) if f == 0 else (np.mean(3 * mine[1, :] - mine[2, :]) / 3)
                else:
                    sh = np.mean(3 * mine[:, -1] - mine[:, -2]) / 3
            assert np.array_equal(np.isnan(offs), np.isnan(c[f])), (H, W)
            np.testing.assert_allclose(c[f], offs, rtol=0, atol=1e-11, equal_nan=True)
            assert np.array_equal(np.isnan(ref), np.isnan(mine - sh)), (H, W)
            np.testing.assert_allclose(mine - sh, ref, rtol=0, atol=1e-10, equal_nan=True)
        done += 1


@pytest.mark.parametrize("variant,ov,H,W", [('deltaU_to_deltaP', 32, 300, 420), ('U_to_gradP', 96, 240, 340), ('thesis', 12, 260, 380)])
def test_per_block_constants_are_removed_by_the_offset_chain(variant, ov, H, W):
    """Known-answer test: blocks that are pure per-block constants re-assemble to ONE constant field -- every strip
    mean then equals the block's constant, so the chain cancels them exactly, in the literal loop and in the compiled
    plan alike.  (For a non-constant truth the reference is only approximate: e.g. SMC:292 compares a 32-row strip with
    the mean of a (128 - p_i)-row strip; equality of plan and loop on such fields is covered above.)"""
    rng = np.random.default_rng(7)
    mask = disc_mask(H, W, H // 2, W // 3, 18)
    plan = psm_b200.compile_plan(variant, H, W, mask, overlap=ov)
    B, F = plan['n_blocks'], plan['n_fields']
    n_x, n_y, origins, il = oracle_plan(variant, H, W, ov)
    blocks = np.zeros((B, F, 128, 128))
    x_array = np.zeros((B, 128, 128, 3))
    for k, (y0, x0) in enumerate(origins):
        x_array[k, :, :, 2] = mask[y0:y0 + 128, x0:x0 + 128] * 0.5
        blocks[k] += rng.standard_normal((F, 1, 1))
    c, fields = eval_plan(plan, blocks, mask)
    for f in range(F):
        if variant == 'thesis':
            ref = oasm.assemble_thesis(blocks[:, 0], x_array, il, n_x, n_y, 128, ov, W, H)
        elif variant == 'deltaU_to_deltaP':
            ref = oasm.assemble_deltas(blocks[:, 0], x_array, il, n_x, n_y, 128, ov, W, H)
        else:
            ref = oasm.assemble_gradp(('dp_dx', 'dp_dy')[f], blocks[:, f], x_array, il, n_x, n_y, 128, ov, W, H)[0, :, :, 0]
        assert np.ptp(ref) < 1e-12 and np.ptp(fields[f]) < 1e-12
        assert abs(ref.mean()) < 1e-12                         # ... and the global shift pins that constant to zero
