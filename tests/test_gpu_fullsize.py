"""Parity at BASELINE.json's full single-GPU sizes (configs[1] and configs[2]: ~1 M cells, 1000 x 1000 grid).

The oracle (reference arithmetic: float64 around a float32 Dense stack) finishes one step at this size in well under a
second once the tables exist, so the check is direct -- Delta p at the cells and the assembled field within the
north-star bound of 1e-3 relative L2 -- plus size-independent properties: the gather reproduces a linear field exactly
where the table is valid, the 5-column resident-U(t-1) mode equals the 7-column mode bit for bit, and a second step on
the same input reproduces the first bit for bit (graph replay, buffers reused, no stale state).
Tables: SciPy Qhull for cells -> grid (as the reference), closed-form grid -> cell tables (Qhull on 10^6 regular grid
points takes minutes; `tables.regular_grid_back_tables`).
"""
import numpy as np
import pytest

import psm_b200
from psm_b200 import synthetic as syn, tables as ptables
from oracle.pipeline import DeltasOracle, GradPOracle
from helpers import oracle_params, rel_l2

pytestmark = pytest.mark.gpu


def _case(variant):
    deltas = variant == 'deltaU_to_deltaP'
    mesh = syn.make_mesh(seed=0, **syn.CONFIGS['c2'])
    F = syn.make_fields(mesh, seed=0)
    params = syn.make_params(seed=0, pc_in=128, pc_p=128, standardization='std' if deltas else 'max_abs',
                             n_out_channels=1 if deltas else 2,
                             maxs=syn.DEFAULT_MAXS if deltas else (1.0, 0.536, 0.999, 0.8, 0.7))
    t = ptables.build_tables(mesh['cells'], mesh['top'], mesh['obst'], F['p_prev'], variant=variant, back='closed_form')
    o = (DeltasOracle if deltas else GradPOracle)(oracle_params(params))
    o.compute_only_once(mesh['cells'], mesh['top'], mesh['obst'], F['p_prev'],
                        tables=(t['vert'], t['weights'], t['vert_back'], t['weights_back']))
    return mesh, F, params, t, o


def test_deltas_one_million_cells_against_oracle():
    mesh, F, params, t, o = _case('deltaU_to_deltaP')
    n = mesh['cells'].shape[0]
    assert n > 950_000
    r = o.time_step(F['Ux'], F['Uy'], F['dUx'], F['dUy'])
    p_ref, _ = o.to_cells(r['field'], F['p_prev'])
    cells7 = syn.pack_cells(mesh, F, with_delta=True)
    with psm_b200.PressureSurrogate('deltaU_to_deltaP', input_cols=7) as sm:
        sm.load_params(params)
        sm.init_tables(t)
        g = sm.geometry()
        assert (g['n_blocks'], g['n_x'], g['n_y'], g['p_i']) == (121, 10, 9, 8)
        out, rc = sm.predict(cells7)
        assert rc == 0
        for _ in range(6):                                                               # replayed graph, reused buffers:
            out2, _ = sm.predict(cells7)                                                 # every step bit-identical (this caught a
            np.testing.assert_array_equal(out, out2)                                     # generic-read / async-refill race in round 1)
        field = sm.stage('field')[0]
        offsets = sm.stage('offsets')[0]
        # linear field through the gather: U = (a + b x + c y) reproduces itself at valid pixels (barycentric exactness)
        x, y = mesh['cells'][:, 0], mesh['cells'][:, 1]
        lin = cells7.copy()
        lin[:, 5] = 0.3 + 0.2 * x - 0.1 * y
        lin[:, 6] = -0.1 + 0.05 * x + 0.4 * y
        sm.predict(lin)
        grid = sm.stage('grid')
        sc = sm.stage('scalars')
    assert rel_l2(out - F['p_prev'], p_ref - F['p_prev']) < 1e-3
    assert rel_l2(field, r['field']) < 1e-3
    np.testing.assert_allclose(offsets, r['offsets'], rtol=0, atol=1e-3 * np.abs(r['offsets']).max())
    valid = t['sdfunct'] != 0
    valid[0, 0] = False        # raster quirk SMC:161,432: pixel (0,0) also receives the LAST invalid point's value
    X0, Y0 = ptables.uniform_grid(*t['bbox'], t['delta'])
    gx = (0.3 + 0.2 * X0 - 0.1 * Y0).reshape(t['H'], t['W']) / (sc[0] * params['maxs'][0])
    gy = (-0.1 + 0.05 * X0 + 0.4 * Y0).reshape(t['H'], t['W']) / (sc[0] * params['maxs'][1])
    assert np.abs(grid[0][valid] - gx[valid]).max() < 2e-6 * np.abs(gx).max()
    assert np.abs(grid[1][valid] - gy[valid]).max() < 2e-6 * np.abs(gy).max()
    # 5-column rows with U(t-1) resident on the device: the same step, bit for bit
    with psm_b200.PressureSurrogate('deltaU_to_deltaP', input_cols=5) as sm5:
        sm5.load_params(params)
        sm5.init_tables(t)
        prev = syn.pack_cells(mesh, F, with_delta=False)
        prev[:, 0] -= F['dUx']
        prev[:, 1] -= F['dUy']
        _, rc0 = sm5.predict(prev)
        assert rc0 == psm_b200.PSM_SKIPPED                                               # no U(t-1) yet
        out5, rc5 = sm5.predict(syn.pack_cells(mesh, F, with_delta=False))
    assert rc5 == 0
    # dU = U - (U - dU) rounds differently from the dU column: compare at the bound, not bitwise
    assert rel_l2(out5 - F['p_prev'], p_ref - F['p_prev']) < 1e-3


def test_gradp_one_million_cells_against_oracle():
    mesh, F, params, t, o = _case('U_to_gradP')
    r = o.time_step(F['Ux'], F['Uy'])
    ref = np.stack([o.to_cells(r['dp_dx']), o.to_cells(r['dp_dy'])], axis=1)
    with psm_b200.PressureSurrogate('U_to_gradP') as sm:
        sm.load_params(params)
        sm.init_tables(t)
        g = sm.geometry()
        assert (g['n_blocks'], g['n_x'], g['n_y']) == (841, 28, 27)
        out, rc = sm.predict(syn.pack_cells(mesh, F, with_delta=False))
        assert rc == 0
        field = sm.stage('field')
    assert np.array_equal(np.isnan(out), np.isnan(ref))
    ok = ~np.isnan(ref)
    assert rel_l2(out[ok], ref[ok]) < 1e-3
    assert rel_l2(field[0], r['dp_dx']) < 1e-3 and rel_l2(field[1], r['dp_dy']) < 1e-3


def _deltas_case(workload, back, seed=0, standardization='std'):
    mesh = syn.make_mesh(seed=seed, **syn.CONFIGS[workload])
    F = syn.make_fields(mesh, seed=seed)
    params = syn.make_params(seed=seed, pc_in=128, pc_p=128, standardization=standardization)
    t = ptables.build_tables(mesh['cells'], mesh['top'], mesh['obst'], F['p_prev'], variant='deltaU_to_deltaP', back=back)
    o = DeltasOracle(oracle_params(params))
    o.compute_only_once(mesh['cells'], mesh['top'], mesh['obst'], F['p_prev'],
                        tables=(t['vert'], t['weights'], t['vert_back'], t['weights_back']))
    return mesh, F, params, t, o


def test_deltas_c1_reference_domain_against_oracle():
    """BASELINE.json configs[0]: the reference's own 15 m x 2 m test-case domain (`test_case/system/blockMeshDict`,
    PMP:195-203): 400 x 3000 grid, n_x = 30, n_y = 2, 124 blocks, ~49 k coarse cells (~24 pixels per cell).  BOTH table
    directions come from SciPy's Qhull here (1.2 M regular grid points for the grid -> cell table, PMP:211), so this is the
    full-size check of the table path the reference itself takes."""
    mesh, F, params, t, o = _deltas_case('c1', back=True)
    r = o.time_step(F['Ux'], F['Uy'], F['dUx'], F['dUy'])
    p_ref, _ = o.to_cells(r['field'], F['p_prev'])
    cells7 = syn.pack_cells(mesh, F, with_delta=True)
    with psm_b200.PressureSurrogate('deltaU_to_deltaP', input_cols=7) as sm:
        sm.load_params(params)
        sm.init_tables(t)
        g = sm.geometry()
        assert (g['grid_h'], g['grid_w'], g['n_blocks'], g['n_x'], g['n_y']) == (400, 3000, 124, 30, 2)
        out, rc = sm.predict(cells7)
        assert rc == 0
        out2, _ = sm.predict(cells7)
        np.testing.assert_array_equal(out, out2)
        field = sm.stage('field')[0]
        offsets = sm.stage('offsets')[0]
        origins, il = sm.plan()
    np.testing.assert_array_equal(il, np.asarray(r['indices_list']))                       # block layout bit-exact (SMC:464-479)
    kept_ref, kept = p_ref == F['p_prev'], out == F['p_prev']
    assert np.array_equal(kept, kept_ref)                                                  # same cells fall back to p_prev (PMP:496)
    assert rel_l2(out - F['p_prev'], p_ref - F['p_prev']) < 1e-3
    assert rel_l2(field, r['field']) < 1e-3
    np.testing.assert_allclose(offsets, r['offsets'], rtol=0, atol=1e-3 * np.abs(r['offsets']).max())


def test_deltas_c5_four_million_cells_against_oracle():
    """BASELINE.json configs[4] geometry (2000 x 2000 grid, ~4 M cells, 441 blocks): the mesh whose per-step latency
    `bench.py --workload c5 --steps 1000` reports.  One step against the oracle, a replayed step bit for bit, and the
    5-column resident-U(t-1) entry the 1000-step run uses."""
    mesh, F, params, t, o = _deltas_case('c5', back='closed_form')
    assert mesh['cells'].shape[0] > 3_900_000
    r = o.time_step(F['Ux'], F['Uy'], F['dUx'], F['dUy'])
    p_ref, _ = o.to_cells(r['field'], F['p_prev'])
    cells7 = syn.pack_cells(mesh, F, with_delta=True)
    with psm_b200.PressureSurrogate('deltaU_to_deltaP', input_cols=7) as sm:
        sm.load_params(params)
        sm.init_tables(t)
        g = sm.geometry()
        assert (g['n_blocks'], g['n_x'], g['n_y'], g['p_i']) == (441, 20, 19, 48)
        out, rc = sm.predict(cells7)
        assert rc == 0
        for _ in range(3):
            out2, _ = sm.predict(cells7)
            np.testing.assert_array_equal(out, out2)
        field = sm.stage('field')[0]
    assert rel_l2(out - F['p_prev'], p_ref - F['p_prev']) < 1e-3
    assert rel_l2(field, r['field']) < 1e-3
    with psm_b200.PressureSurrogate('deltaU_to_deltaP', input_cols=5) as sm5:
        sm5.load_params(params)
        sm5.init_tables(t)
        prev = syn.pack_cells(mesh, F, with_delta=False)
        prev[:, 0] -= F['dUx']
        prev[:, 1] -= F['dUy']
        _, rc0 = sm5.predict(prev)
        out5, rc5 = sm5.predict(syn.pack_cells(mesh, F, with_delta=False))
    assert (rc0, rc5) == (psm_b200.PSM_SKIPPED, 0)
    assert rel_l2(out5 - F['p_prev'], p_ref - F['p_prev']) < 1e-3
