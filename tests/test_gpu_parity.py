"""GPU parity: the CUDA path (through the C-ABI) against the CPU oracle on the same seeded inputs.

Tolerances: integer tables / plans bit-exact; float32 device storage against the float64 oracle:
grid 2e-6 abs (values are O(1)), PCA coordinates / MLP output 1e-4 relative L2, deltaP (and grad p)
1e-3 relative L2 -- the bound BASELINE.json's north_star states.
"""
import numpy as np
import pytest

import psm_b200
from psm_b200 import synthetic as syn, tables as ptables, _capi
from oracle.pipeline import DeltasOracle, GradPOracle
from helpers import oracle_params, load_golden, rel_l2

pytestmark = pytest.mark.gpu


def folded_forward_table(vert, weights, indices, H, W):
    """What the device table must hold (DESIGN.md 'validity fold'), from the oracle's tables."""
    G = H * W
    src = np.full(G, -1, dtype=np.int64)
    flat = indices[:, 0] * W + indices[:, 1]
    src[flat] = np.arange(G)                      # duplicates: last wins, like NumPy fancy assignment
    v = np.zeros((G, 3), np.int32)
    w = np.zeros((G, 3), np.float32)
    ok = src >= 0
    v[ok] = vert[src[ok]]
    ww = weights[src[ok]].copy()
    ww[np.any(ww < 0, axis=1)] = 0.0
    w[ok] = ww.astype(np.float32)
    return v, w


class Case:
    def __init__(self, variant, mesh_kw, seed=0, pc_in=128, pc_p=128, input_cols=None, **kw):
        self.variant = variant
        self.mesh = syn.make_mesh(seed=seed, **mesh_kw)
        self.F = syn.make_fields(self.mesh, seed=seed)
        deltas = variant == 'deltaU_to_deltaP'
        self.params = syn.make_params(seed=seed, pc_in=pc_in, pc_p=pc_p,
                                      standardization='std' if deltas else 'max_abs',
                                      n_out_channels=1 if deltas else 2,
                                      maxs=syn.DEFAULT_MAXS if deltas else (1.0, 0.536, 0.999, 0.8, 0.7))
        probe = self.F['p_prev']
        self.tables = ptables.build_tables(self.mesh['cells'], self.mesh['top'], self.mesh['obst'], probe, variant=variant)
        P = oracle_params(self.params)
        self.oracle = DeltasOracle(P) if deltas else GradPOracle(P)
        self.oracle.compute_only_once(self.mesh['cells'], self.mesh['top'], self.mesh['obst'], probe,
                                      tables=(self.tables['vert'], self.tables['weights'],
                                              self.tables['vert_back'], self.tables['weights_back']))
        self.sm = psm_b200.PressureSurrogate(variant=variant, input_cols=input_cols, **kw)
        self.sm.load_params(self.params)
        self.sm.init_tables(self.tables)
        self.cells = syn.pack_cells(self.mesh, self.F, with_delta=(self.sm.input_cols == 7))


@pytest.fixture(scope="module")
def deltas_case():
    c = Case('deltaU_to_deltaP', syn.CONFIGS['tiny'], seed=3)
    yield c
    c.sm.close()


@pytest.fixture(scope="module")
def gradp_case():
    c = Case('U_to_gradP', dict(H=240, W=340, nx=130, ny=90, R=0.1), seed=4, pc_in=48, pc_p=40)
    yield c
    c.sm.close()


def test_product_tables_equal_oracle_tables(deltas_case):
    """Shim-side init (psm_b200.tables) against the oracle restatement of SMC:89-180."""
    c = deltas_case
    o2 = DeltasOracle(c.oracle.params)
    o2.compute_only_once(c.mesh['cells'], c.mesh['top'], c.mesh['obst'], c.F['p_prev'])
    np.testing.assert_array_equal(c.tables['vert'], o2.vert)
    np.testing.assert_array_equal(c.tables['weights'], o2.weights)
    np.testing.assert_array_equal(c.tables['vert_back'], o2.vert_back)
    np.testing.assert_array_equal(c.tables['indices'], o2.indices)
    np.testing.assert_allclose(c.tables['sdfunct'], o2.sdfunct[:, :, 0], rtol=0, atol=1e-14)
    assert np.array_equal(c.tables['sdfunct'] != 0, o2.sdfunct[:, :, 0] != 0)


def test_device_tables_and_plan_bit_exact(deltas_case):
    c = deltas_case
    v, w = c.sm.forward_table()
    ev, ew = folded_forward_table(c.oracle.vert, c.oracle.weights, c.oracle.indices, c.sm.H, c.sm.W)
    np.testing.assert_array_equal(v, ev)
    np.testing.assert_array_equal(w.view(np.uint32), ew.view(np.uint32))      # bit-exact after f64->f32
    origins, il = c.sm.plan()
    n_x, n_y, o_orig, o_il = c.oracle.block_plan()
    np.testing.assert_array_equal(origins, np.array(o_orig, np.int32))
    np.testing.assert_array_equal(il, np.array(o_il, np.int32))
    g = c.sm.geometry()
    assert (g['n_x'], g['n_y']) == (n_x, n_y)


def test_deltas_stages_match_oracle(deltas_case):
    c = deltas_case
    out, rc = c.sm.predict(c.cells)
    assert rc == _capi.PSM_OK
    r = c.oracle.time_step(c.F['Ux'], c.F['Uy'], c.F['dUx'], c.F['dUy'])
    sc = c.sm.stage('scalars')
    assert sc[0] == r['U_max_norm']                                             # bit-exact (no FMA)
    grid = c.sm.stage('grid')
    np.testing.assert_allclose(grid[0], r['grid'][:, :, 0], rtol=0, atol=2e-6 * np.abs(r['grid'][:, :, 0]).max())
    np.testing.assert_allclose(grid[1], r['grid'][:, :, 1], rtol=0, atol=2e-6 * np.abs(r['grid'][:, :, 1]).max())
    assert rel_l2(c.sm.stage('x_input'), r['x_input']) < 1e-4
    P = c.oracle.params
    assert rel_l2(c.sm.stage('mlp_out'), r['mlp_out'] * P.std_out + P.mean_out) < 1e-4
    blocks = c.sm.stage('blocks')[:, 0]
    assert rel_l2(blocks, r['blocks'][..., 0]) < 1e-4
    offs = c.sm.stage('offsets')[0]
    np.testing.assert_allclose(offs, r['offsets'], rtol=0, atol=1e-4 * np.abs(r['blocks']).max())
    field = c.sm.stage('field')[0]
    assert np.array_equal(np.isnan(field), np.isnan(r['field']))
    assert rel_l2(field, r['field']) < 1e-3
    p_ref, _ = c.oracle.to_cells(r['field'], c.F['p_prev'], additive=True)
    assert rel_l2(out - c.F['p_prev'], p_ref - c.F['p_prev']) < 1e-3
    # cells outside the grid hull keep the previous pressure exactly (PMP:496)
    _, interp = c.oracle.to_cells(r['field'], c.F['p_prev'])
    nanc = np.isnan(interp)
    assert nanc.any()
    np.testing.assert_array_equal(out[nanc], c.F['p_prev'][nanc])
    assert c.sm.launch_count() >= 10


def test_deltas_matches_reference_golden():
    """End to end against the field the reference itself produced (tests/golden/smc_small.npz)."""
    z, mesh_kw, seed = load_golden('smc_small')
    mesh = syn.make_mesh(seed=seed, **mesh_kw)
    F = syn.make_fields(mesh, seed=seed)
    params = syn.make_params(seed=seed, pc_in=int(z['pc_in']), pc_p=int(z['pc_p']), standardization='std')
    with psm_b200.PressureSurrogate('deltaU_to_deltaP') as sm:
        sm.load_params(params)
        sm.init_from_mesh(mesh['cells'], mesh['top'], mesh['obst'], F['p_prev'])
        sm.predict(syn.pack_cells(mesh, F))
        field = sm.stage('field')[0]
        origins, il = sm.plan()
        np.testing.assert_array_equal(il, z['indices_list'])
        assert rel_l2(field, z['field']) < 1e-3
        blocks = sm.stage('blocks')[:, 0]
        assert rel_l2(blocks[:, ::4, ::4], z['blocks_sub']) < 1e-4
        np.testing.assert_allclose(sm.stage('offsets')[0], z['offsets'], rtol=0, atol=1e-4 * np.abs(z['blocks_sub']).max())


def test_deltas_nan_fallback_branch_matches_reference_golden():
    z, mesh_kw, seed = load_golden('smc_bigobst')
    mesh = syn.make_mesh(seed=seed, **mesh_kw)
    F = syn.make_fields(mesh, seed=seed)
    params = syn.make_params(seed=seed, pc_in=int(z['pc_in']), pc_p=int(z['pc_p']), standardization='std')
    with psm_b200.PressureSurrogate('deltaU_to_deltaP') as sm:
        sm.load_params(params)
        sm.init_from_mesh(mesh['cells'], mesh['top'], mesh['obst'], F['p_prev'])
        sm.predict(syn.pack_cells(mesh, F))
        assert rel_l2(sm.stage('field')[0], z['field']) < 1e-3


def test_gradp_stages_match_oracle_and_golden(gradp_case):
    c = gradp_case
    out, rc = c.sm.predict(c.cells)
    assert rc == _capi.PSM_OK
    r = c.oracle.time_step(c.F['Ux'], c.F['Uy'])
    origins, il = c.sm.plan()
    np.testing.assert_array_equal(origins, np.array(r['origins'], np.int32))
    np.testing.assert_array_equal(il, np.array(r['indices_list'], np.int32))
    assert rel_l2(c.sm.stage('x_input'), r['x_input']) < 1e-4
    blocks = c.sm.stage('blocks')
    assert rel_l2(blocks[:, 0], r['blocks'][..., 0]) < 1e-4
    assert rel_l2(blocks[:, 1], r['blocks'][..., 1]) < 1e-4
    field = c.sm.stage('field')
    assert rel_l2(field[0], r['dp_dx']) < 1e-3
    assert rel_l2(field[1], r['dp_dy']) < 1e-3
    offs = c.sm.stage('offsets')
    np.testing.assert_allclose(offs[0], r['dp_dx_offsets'], rtol=0, atol=1e-4 * np.abs(r['blocks']).max())
    np.testing.assert_allclose(offs[1], r['dp_dy_offsets'], rtol=0, atol=1e-4 * np.abs(r['blocks']).max())
    gx, gy = c.oracle.to_cells(r['dp_dx']), c.oracle.to_cells(r['dp_dy'])
    assert out.shape == (c.cells.shape[0], 2)
    assert np.array_equal(np.isnan(out[:, 0]), np.isnan(gx))
    ok = ~np.isnan(gx)
    assert rel_l2(out[ok, 0], gx[ok]) < 1e-3 and rel_l2(out[ok, 1], gy[ok]) < 1e-3


def test_gradp_matches_reference_golden():
    z, mesh_kw, seed = load_golden('grad_small')
    mesh = syn.make_mesh(seed=seed, **mesh_kw)
    F = syn.make_fields(mesh, seed=seed)
    params = syn.make_params(seed=seed, pc_in=int(z['pc_in']), pc_p=int(z['pc_p']), standardization='max_abs',
                             n_out_channels=2, maxs=(1.0, 0.536, 0.999, 0.8, 0.7))
    rng = np.random.default_rng(seed + 5)
    lab = 0.01 * rng.standard_normal((mesh['cells'].shape[0], 3))
    with psm_b200.PressureSurrogate('U_to_gradP') as sm:
        sm.load_params(params)
        sm.init_from_mesh(mesh['cells'], mesh['top'], mesh['obst'], lab[:, 0])
        sm.predict(syn.pack_cells(mesh, F, with_delta=False))
        field = sm.stage('field')
        assert rel_l2(field[0], z['dp_dx']) < 1e-3
        assert rel_l2(field[1], z['dp_dy']) < 1e-3


def test_resident_previous_velocity_mode_and_skip_rule():
    """5-column input: dU is formed on the device from the resident U(t-1); first call and
    'irrelevant' steps (SMC:410-415) return p_prev with PSM_SKIPPED."""
    mesh = syn.make_mesh(seed=5, **syn.CONFIGS['tiny'])
    F = syn.make_fields(mesh, seed=5)
    params = syn.make_params(seed=5, pc_in=32, pc_p=32)
    t = ptables.build_tables(mesh['cells'], mesh['top'], mesh['obst'], F['p_prev'])
    a = psm_b200.PressureSurrogate('deltaU_to_deltaP', input_cols=5)
    b = psm_b200.PressureSurrogate('deltaU_to_deltaP', input_cols=7)
    for sm in (a, b):
        sm.load_params(params)
        sm.init_tables(t)
    c5 = syn.pack_cells(mesh, F, with_delta=False)
    out0, rc0 = a.predict(c5)
    assert rc0 == _capi.PSM_SKIPPED
    np.testing.assert_array_equal(out0, F['p_prev'])
    c5b = c5.copy()
    c5b[:, 0] += F['dUx']
    c5b[:, 1] += F['dUy']
    out1, rc1 = a.predict(c5b)
    assert rc1 == _capi.PSM_OK
    c7 = np.concatenate([c5b, (c5b[:, 0] - c5[:, 0])[:, None], (c5b[:, 1] - c5[:, 1])[:, None]], axis=1)
    out2, rc2 = b.predict(c7)
    assert rc2 == _capi.PSM_OK
    np.testing.assert_array_equal(out1, out2)
    c7s = c7.copy()
    c7s[:, 5:7] *= 1e-4                                        # |dU|/|U| ~ 1e-6 < 1e-4
    out3, rc3 = b.predict(c7s)
    assert rc3 == _capi.PSM_SKIPPED
    np.testing.assert_array_equal(out3, c7s[:, 4])
    a.close()
    b.close()


@pytest.mark.parametrize("mode,tol", [(1, 2e-3), (2, 1e-4)])
def test_gemm_modes_cross_check(deltas_case, mode, tol):
    """Single-pass TF32 and the CUDA-core FP32 kernel against the default 3xTF32 tensor-core path."""
    c = deltas_case
    ref, _ = c.sm.predict(c.cells)
    f_ref = c.sm.stage('field')[0]
    with psm_b200.PressureSurrogate('deltaU_to_deltaP', gemm_mode=mode) as sm:
        sm.load_params(c.params)
        sm.init_tables(c.tables)
        out, rc = sm.predict(c.cells)
        assert rc == _capi.PSM_OK
        assert rel_l2(sm.stage('field')[0], f_ref) < tol
        assert rel_l2(out - c.F['p_prev'], ref - c.F['p_prev']) < tol


def test_error_codes_not_crashes(deltas_case):
    c = deltas_case
    with pytest.raises(_capi.PsmError) as e:
        c.sm.predict(c.cells[:-1])
    assert e.value.code == _capi.PSM_ERR_INVALID
    sm = psm_b200.PressureSurrogate('deltaU_to_deltaP')
    with pytest.raises(_capi.PsmError) as e:
        sm.init_tables(c.tables)                               # params not loaded yet
    assert e.value.code == _capi.PSM_ERR_STATE
    sm.close()
    sm.close()                                                 # idempotent destroy


def test_linearity_of_assembly_in_block_constants(deltas_case):
    """Size-independent property: predicting twice gives identical results (no hidden state in 7-col mode)."""
    c = deltas_case
    o1, _ = c.sm.predict(c.cells)
    o2, _ = c.sm.predict(c.cells)
    np.testing.assert_array_equal(o1, o2)


def test_cluster_dense_layers_equal_split_k_gemms_on_several_batch_tiles(monkeypatch):
    """The cluster split-K Dense kernel (DSMEM reduction, fused epilogue) against the split-K GEMM + reduce kernels
    and the oracle, on a mesh with more than 128 blocks (two 128-row batch tiles => two clusters)."""
    c = Case('U_to_gradP', dict(H=500, W=420, nx=150, ny=180, R=0.1), seed=9, pc_in=45, pc_p=48)
    try:
        assert c.sm.geometry()['n_blocks'] > 128
        c.sm.predict(c.cells)
        fused = c.sm.stage('mlp_out')
        monkeypatch.setenv('PSM_NO_DENSE_CLUSTER', '1')
        with psm_b200.PressureSurrogate('U_to_gradP') as sm:
            sm.load_params(c.params)
            sm.init_tables(c.tables)
            sm.predict(c.cells)
            layered = sm.stage('mlp_out')
        r = c.oracle.time_step(c.F['Ux'], c.F['Uy'])
        P = c.oracle.params
        want = r['mlp_out'] * P.max_abs_output_PCA
        assert rel_l2(fused, layered) < 2e-5
        assert rel_l2(fused, want) < 1e-4
    finally:
        c.sm.close()


def test_fused_gather_extraction_is_bit_identical_to_the_two_kernel_path(deltas_case, monkeypatch):
    """gather_extract_kernel writes the block operand itself; the separate extract kernel re-reads the grid.
    Same arithmetic, so every downstream stage must agree bit for bit."""
    c = deltas_case
    c.sm.predict(c.cells)
    fused = {k: c.sm.stage(k) for k in ('grid', 'x_input', 'blocks', 'field')}
    monkeypatch.setenv('PSM_NO_FUSED_EXTRACT', '1')
    with psm_b200.PressureSurrogate('deltaU_to_deltaP') as sm:
        sm.load_params(c.params)
        sm.init_tables(c.tables)
        sm.predict(c.cells)
        for k, v in fused.items():
            np.testing.assert_array_equal(sm.stage(k), v, err_msg=k)


@pytest.mark.parametrize("variant", ['deltaU_to_deltaP', 'U_to_gradP'])
def test_gaussian_post_filter(variant):
    """filter_sigma = 10 reproduces the reference's apply_filter=True (SMC:353-356 / GRAD:366-367); the golden
    fixture smc_small holds the reference's own filtered field."""
    deltas = variant == 'deltaU_to_deltaP'
    if deltas:
        z, mesh_kw, seed = load_golden('smc_small')
        c = Case(variant, mesh_kw, seed=seed, pc_in=int(z['pc_in']), pc_p=int(z['pc_p']), filter_sigma=10.0)
    else:
        c = Case(variant, dict(H=240, W=340, nx=130, ny=90, R=0.1), seed=4, pc_in=48, pc_p=40, filter_sigma=10.0)
    try:
        out, rc = c.sm.predict(c.cells)
        assert rc == 0
        field = c.sm.stage('field')
        if deltas:
            r = c.oracle.time_step(c.F['Ux'], c.F['Uy'], c.F['dUx'], c.F['dUy'], apply_filter=True)
            assert rel_l2(field[0], r['field']) < 1e-3
            assert rel_l2(field[0][::2, ::2], z['field_filtered_sub']) < 1e-3        # the reference's own output
            p_ref, _ = c.oracle.to_cells(r['field'], c.F['p_prev'])
            assert rel_l2(out - c.F['p_prev'], p_ref - c.F['p_prev']) < 1e-3
            r0 = c.oracle.time_step(c.F['Ux'], c.F['Uy'], c.F['dUx'], c.F['dUy'])
            assert rel_l2(r['field'], r0['field']) > 1e-2                            # the filter is not a no-op here
        else:
            r = c.oracle.time_step(c.F['Ux'], c.F['Uy'], apply_filter=True)
            assert rel_l2(field[0], r['dp_dx']) < 1e-3 and rel_l2(field[1], r['dp_dy']) < 1e-3
    finally:
        c.sm.close()


def test_min_max_standardisation_matches_oracle():
    """SMC:513-520,535-536: the third standardisation method, (z - min) / (max - min) in, r * (max - min) + min out."""
    mesh = syn.make_mesh(seed=8, **syn.CONFIGS['tiny'])
    F = syn.make_fields(mesh, seed=8)
    params = syn.make_params(seed=8, pc_in=40, pc_p=24, standardization='min_max')
    t = ptables.build_tables(mesh['cells'], mesh['top'], mesh['obst'], F['p_prev'])
    o = DeltasOracle(oracle_params(params))
    o.compute_only_once(mesh['cells'], mesh['top'], mesh['obst'], F['p_prev'],
                        tables=(t['vert'], t['weights'], t['vert_back'], t['weights_back']))
    r = o.time_step(F['Ux'], F['Uy'], F['dUx'], F['dUy'])
    with psm_b200.PressureSurrogate('deltaU_to_deltaP') as sm:
        sm.load_params(params)
        sm.init_tables(t)
        out, rc = sm.predict(syn.pack_cells(mesh, F))
        assert rc == 0
        assert rel_l2(sm.stage('x_input'), r['x_input']) < 1e-4
        assert rel_l2(sm.stage('field')[0], r['field']) < 1e-3
    with pytest.raises(ValueError):
        psm_b200.surrogate._marshal_params(dict(params, standardization='robust'))


def test_native_field_entry_matches_the_row_entry(deltas_case):
    """psm_predict_fields (the solver's own U / p arrays, no row packing) against psm_predict on the packed rows:
    same kernels after the first one, so bit-identical; p = NULL returns the raw prediction; the resident-U(t-1) form
    reproduces the 5-column behaviour (first call PSM_SKIPPED)."""
    import torch
    c = deltas_case
    F, n = c.F, c.mesh['cells'].shape[0]
    ref, rc = c.sm.predict(c.cells)
    assert rc == 0
    for stride in (3, 2):
        U = np.zeros((n, stride)); U[:, 0], U[:, 1] = F['Ux'], F['Uy']
        dU = np.zeros((n, stride)); dU[:, 0], dU[:, 1] = F['dUx'], F['dUy']
        out, rc = c.sm.predict_fields(U, p=F['p_prev'], dU=dU)
        assert rc == 0
        np.testing.assert_array_equal(out, ref)
        raw, rc = c.sm.predict_fields(U, p=None, dU=dU)
        assert rc == 0
        kept = ref == F['p_prev']
        assert np.all(raw[kept] == 0.0)                                       # cells that keep p_prev: delta_p = 0
        np.testing.assert_allclose(F['p_prev'] + raw, ref, rtol=0, atol=1e-12 * np.abs(ref).max())
    # a fresh buffer every call (what a Python caller does): nothing is keyed by the caller's pointers
    for _ in range(3):
        out, _ = c.sm.predict_fields(U.copy(), p=F['p_prev'].copy(), dU=dU.copy())
        np.testing.assert_array_equal(out, ref)
    # device-pointer form: p read in place
    dUt, ddU, dp = torch.from_numpy(U).cuda(), torch.from_numpy(dU).cuda(), torch.from_numpy(F['p_prev']).cuda()
    dout = torch.empty(n, dtype=torch.float64, device='cuda')
    torch.cuda.synchronize()
    for _ in range(3):                                                        # captured once, then replayed
        assert c.sm.predict_fields_device(dUt.data_ptr(), 2, n, dout.data_ptr(), d_p_ptr=dp.data_ptr(), d_dU_ptr=ddU.data_ptr()) == 0
        np.testing.assert_array_equal(dout.cpu().numpy(), ref)
    # resident U(t-1): first call has nothing to difference against
    with psm_b200.PressureSurrogate('deltaU_to_deltaP', input_cols=5) as sm5:
        sm5.load_params(c.params)
        sm5.init_tables(c.tables)
        U3 = np.zeros((n, 3)); U3[:, 0], U3[:, 1] = F['Ux'] - F['dUx'], F['Uy'] - F['dUy']
        out0, rc0 = sm5.predict_fields(U3, p=F['p_prev'])
        assert rc0 == psm_b200.PSM_SKIPPED
        np.testing.assert_array_equal(out0, F['p_prev'])
        U3[:, 0], U3[:, 1] = F['Ux'], F['Uy']
        out1, rc1 = sm5.predict_fields(U3, p=F['p_prev'])
        assert rc1 == 0
        rows5 = syn.pack_cells(c.mesh, F, with_delta=False)
    with psm_b200.PressureSurrogate('deltaU_to_deltaP', input_cols=5) as sm5b:
        sm5b.load_params(c.params)
        sm5b.init_tables(c.tables)
        prev = rows5.copy(); prev[:, 0] -= F['dUx']; prev[:, 1] -= F['dUy']
        sm5b.predict(prev)
        out_rows, _ = sm5b.predict(rows5)
    np.testing.assert_array_equal(out1, out_rows)


def test_gradp_pressure_recovery_matches_reference():
    """f.2: integrate_field on four quadrants + stitch (GRAD:371-416, 585-628) on the GPU.  (a) against the CPU restatement
    applied to the GPU's own gradient fields: FP64 on both sides, only the scan order differs; (b) against the pressure
    field the reference's timeStep itself produced (tests/golden/grad_integrate.npz), end to end within the 1e-3 bound."""
    from oracle import integrate as ointeg
    z, mesh_kw, seed = load_golden('grad_integrate')
    mesh = syn.make_mesh(seed=seed, **mesh_kw)
    F = syn.make_fields(mesh, seed=seed)
    params = syn.make_params(seed=seed, pc_in=int(z['pc_in']), pc_p=int(z['pc_p']), standardization='max_abs',
                             n_out_channels=2, maxs=(1.0, 0.536, 0.999, 0.8, 0.7))
    rng = np.random.default_rng(seed + 5)
    lab = 0.01 * rng.standard_normal((mesh['cells'].shape[0], 3))
    bb = z['bbox']
    with psm_b200.PressureSurrogate('U_to_gradP') as sm:
        sm.load_params(params)
        sm.init_from_mesh(mesh['cells'], mesh['top'], mesh['obst'], lab[:, 0])
        sm.predict(syn.pack_cells(mesh, F, with_delta=False))
        field = sm.stage('field')
        p = sm.integrate_gradp(mesh['top'], float(z['x0_min']))
        with pytest.raises(psm_b200.PsmError):                       # a centre row that misses the obstacle: the reference fails too
            sm.integrate_gradp(mesh['top'], float(z['x0_min']), center_row=20)
    assert [sm.H, sm.W] == list(z['grid_shape'])
    sdf = z['sdfunct'][:, :, None]
    ref_own, cx = ointeg.recover_pressure(field[0].astype(np.float64), field[1].astype(np.float64), sdf, bb[0], bb[1], bb[2], bb[3],
                                          float(z['x0_min']), 5e-3)
    np.testing.assert_allclose(p, ref_own, rtol=0, atol=1e-11 * np.abs(ref_own).max())
    assert rel_l2(field[0], z['dp_dx']) < 1e-3 and rel_l2(field[1], z['dp_dy']) < 1e-3
    assert rel_l2(p, z['p_field']) < 1e-3


def test_strip_sums_from_the_inverse_epilogue_match_the_means_kernel(deltas_case, monkeypatch):
    """PSM_STRIP_FUSE=1: the masked strip means come out of the PCA-inverse epilogue (FP32 partials of 32 pixels, FP64 beyond)
    instead of task_means_kernel (FP64 throughout) -- the path every sharded handle takes.  Same means to ~1e-6, same pressures."""
    c = deltas_case
    ref, _ = c.sm.predict(c.cells)
    means_ref = c.sm.stage('means')
    off_ref = c.sm.stage('offsets')
    monkeypatch.setenv('PSM_STRIP_FUSE', '1')
    with psm_b200.PressureSurrogate('deltaU_to_deltaP') as sm:
        sm.load_params(c.params)
        sm.init_tables(c.tables)
        out, rc = sm.predict(c.cells)
        out2, _ = sm.predict(c.cells)
        means = sm.stage('means')
        off = sm.stage('offsets')
    assert rc == 0
    np.testing.assert_array_equal(out, out2)
    assert np.array_equal(np.isnan(means), np.isnan(means_ref))
    ok = ~np.isnan(means_ref)
    np.testing.assert_allclose(means[ok], means_ref[ok], rtol=0, atol=2e-6 * np.abs(means_ref[ok]).max())
    np.testing.assert_allclose(off, off_ref, rtol=0, atol=1e-5 * np.abs(off_ref).max())
    assert rel_l2(out - c.F['p_prev'], ref - c.F['p_prev']) < 1e-5
