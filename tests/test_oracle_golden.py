"""Pin the oracle against outputs of the reference itself (tests/golden/make_golden.py)."""
import hashlib

import numpy as np
import pytest

from psm_b200 import synthetic as syn
from oracle.pipeline import DeltasOracle, GradPOracle
from oracle import interp as ointerp, domain as odomain
from helpers import oracle_params, load_golden


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize("name", ["smc_small", "smc_bigobst"])
def test_deltas_oracle_matches_reference(name):
    z, mesh_kw, seed = load_golden(name)
    mesh = syn.make_mesh(seed=seed, **mesh_kw)
    F = syn.make_fields(mesh, seed=seed)
    P = oracle_params(syn.make_params(seed=seed, pc_in=int(z['pc_in']), pc_p=int(z['pc_p']), standardization='std'))
    o = DeltasOracle(P, delta=5e-3, shape=128, overlap=32)
    o.compute_only_once(mesh['cells'], mesh['top'], mesh['obst'], F['p_prev'], back_tables=False)
    # init tables: bit-exact integers, weights to round-off (LAPACK kernels may differ per CPU)
    assert [o.grid_shape_y, o.grid_shape_x] == list(z['grid_shape'])
    assert sha(o.vert.astype(np.int32)) == str(z['vert_sha'])
    np.testing.assert_array_equal(o.vert[::17].astype(np.int32), z['vert_sub'])
    np.testing.assert_allclose(o.weights[::17], z['weights_sub'], rtol=0, atol=1e-12)
    assert sha(o.indices.astype(np.int64)) == str(z['indices_sha'])
    assert sha(o.sdfunct != 0) == str(z['sdfunct_nonzero_sha'])
    np.testing.assert_allclose(o.sdfunct[:, :, 0], z['sdfunct'], rtol=1e-6, atol=1e-7)
    # step
    r = o.time_step(F['Ux'], F['Uy'], F['dUx'], F['dUy'])
    assert (r['n_x'], r['n_y']) == (int(z['n_x']), int(z['n_y']))
    np.testing.assert_array_equal(np.array(r['indices_list']), z['indices_list'])
    np.testing.assert_allclose(r['x_array'][:, ::8, ::8, :], z['x_array_sub'], rtol=0, atol=1e-13)
    scale = np.abs(z['blocks_sub']).max()
    np.testing.assert_allclose(r['blocks'][:, ::4, ::4, 0], z['blocks_sub'], rtol=0, atol=2e-6 * scale)
    np.testing.assert_allclose(r['offsets'], z['offsets'], rtol=0, atol=2e-6 * scale)
    assert np.array_equal(np.isnan(r['field']), np.isnan(z['field']))
    np.testing.assert_allclose(r['field'], z['field'], rtol=0, atol=4e-6 * scale)
    # the optional Gaussian post-filter of the reference (apply_filter=True, SMC:353-356)
    rf = o.time_step(F['Ux'], F['Uy'], F['dUx'], F['dUy'], apply_filter=True)
    assert np.array_equal(np.isnan(rf['field'][::2, ::2]), np.isnan(z['field_filtered_sub']))
    np.testing.assert_allclose(rf['field'][::2, ::2], z['field_filtered_sub'], rtol=0, atol=4e-6 * scale)


def test_deltas_raster_loop_equals_vectorised():
    """The vectorised raster is the literal SMC:168-176 loop."""
    z, mesh_kw, seed = load_golden("smc_small")
    mesh = syn.make_mesh(seed=seed, **mesh_kw)
    F = syn.make_fields(mesh, seed=seed)
    P = oracle_params(syn.make_params(seed=seed, pc_in=8, pc_p=8))
    a = DeltasOracle(P)
    a.compute_only_once(mesh['cells'], mesh['top'], mesh['obst'], F['p_prev'], back_tables=False)
    b = DeltasOracle(P)
    b.compute_only_once(mesh['cells'], mesh['top'], mesh['obst'], F['p_prev'], tables=(a.vert, a.weights),
                        literal_raster=True)
    np.testing.assert_array_equal(a.indices, b.indices)
    np.testing.assert_array_equal(a.sdfunct, b.sdfunct)


def test_gradp_oracle_matches_reference():
    z, mesh_kw, seed = load_golden("grad_small")
    mesh = syn.make_mesh(seed=seed, **mesh_kw)
    F = syn.make_fields(mesh, seed=seed)
    P = oracle_params(syn.make_params(seed=seed, pc_in=int(z['pc_in']), pc_p=int(z['pc_p']), standardization='max_abs',
                                      n_out_channels=2, maxs=(1.0, 0.536, 0.999, 0.8, 0.7)))
    o = GradPOracle(P, delta=5e-3, shape=128, avance=96)
    rng = np.random.default_rng(seed + 5)
    lab = 0.01 * rng.standard_normal((mesh['cells'].shape[0], 3))
    o.compute_only_once(mesh['cells'], mesh['top'], mesh['obst'], lab[:, 0], back_tables=False)
    assert [o.grid_shape_y, o.grid_shape_x] == list(z['grid_shape'])
    assert sha(o.vert.astype(np.int32)) == str(z['vert_sha'])
    assert sha(o.indices.astype(np.int64)) == str(z['indices_sha'])
    np.testing.assert_allclose(o.sdfunct[:, :, 0], z['sdfunct'], rtol=1e-6, atol=1e-7)
    r = o.time_step(F['Ux'], F['Uy'])
    assert (r['n_x'], r['n_y']) == (int(z['n_x']), int(z['n_y']))
    np.testing.assert_array_equal(np.array(r['indices_list']), z['indices_list'])
    np.testing.assert_allclose(r['x_array'][:, ::8, ::8, :], z['x_array_sub'], rtol=0, atol=1e-13)
    scale = np.abs(z['blocks_dx_sub']).max()
    np.testing.assert_allclose(r['blocks'][:, ::4, ::4, 0], z['blocks_dx_sub'], rtol=0, atol=2e-6 * scale)
    np.testing.assert_allclose(r['blocks'][:, ::4, ::4, 1], z['blocks_dy_sub'], rtol=0, atol=2e-6 * scale)
    np.testing.assert_allclose(r['dp_dx'], z['dp_dx'], rtol=0, atol=4e-6 * scale)
    np.testing.assert_allclose(r['dp_dy'], z['dp_dy'], rtol=0, atol=4e-6 * scale)


def test_pmp_init_tables_match_reference():
    """Solver-side init (PMP:172-247): both table directions, [::10] distance field, raster."""
    z, mesh_kw, seed = load_golden("pmp_init_small")
    mesh = syn.make_mesh(seed=seed, **mesh_kw)
    F = syn.make_fields(mesh, seed=seed)
    cells = mesh['cells']
    x_min, x_max = round(np.min(cells[:, 0]), 2), round(np.max(cells[:, 0]), 2)
    y_min, y_max = round(np.min(cells[:, 1]), 2), round(np.max(cells[:, 1]), 2)
    X0, Y0 = ointerp.create_uniform_grid(x_min, x_max, y_min, y_max, 5e-3)
    xy0 = np.c_[X0, Y0]
    vf, wf = ointerp.interp_weights(cells, xy0)
    vb, wb = ointerp.interp_weights(xy0, cells)
    H, W = int(round((y_max - y_min) / 5e-3)), int(round((x_max - x_min) / 5e-3))
    assert [H, W] == list(z['grid_shape'])
    assert sha(vf.astype(np.int32)) == str(z['vert_fwd_sha'])
    np.testing.assert_allclose(wf[::17], z['weights_fwd_sub'], rtol=0, atol=1e-12)
    assert sha(vb.astype(np.int32)) == str(z['vert_back_sha'])
    np.testing.assert_array_equal(vb[::5].astype(np.int32), z['vert_back_sub'])
    np.testing.assert_allclose(wb[::5], z['weights_back_sub'], rtol=0, atol=1e-10)
    dom, sdf, _ = odomain.domain_dist(xy0, mesh['top'], mesh['obst'], 'pmp')
    probe = ointerp.interpolate_fill(F['Ux'], vf, wf)
    ind, sdfunct = odomain.index_raster(X0, Y0, 5e-3, H, W, dom, probe, sdf)
    # PMP:225 allocates `indices` with np.empty: only the rows its raster loop writes (PMP:233-243) are defined and hashed
    wr = dom & ~np.isnan(probe)
    assert sha(wr.astype(np.uint8)) == str(z['written_sha'])
    assert sha(ind[wr].astype(np.int64)) == str(z['indices_sha'])
    np.testing.assert_allclose(sdfunct[:, :, 0], z['sdfunct'], rtol=1e-6, atol=1e-7)


def test_thesis_oracle_matches_reference_py_func():
    """ThesisOracle (PMP init_func + py_func restated) against the pressures the reference module itself returned."""
    from oracle.pipeline import ThesisOracle
    z, mesh_kw, seed = load_golden("pmp_step_small")
    mesh = syn.make_mesh(seed=seed, **mesh_kw)
    F = syn.make_fields(mesh, seed=seed)
    P = oracle_params(syn.make_params(seed=seed, pc_in=int(z['pc_in']), pc_p=int(z['pc_p']), standardization='max_abs'))
    o = ThesisOracle(P)
    o.init_func(mesh['cells'], mesh['top'], mesh['obst'], F['Ux'])
    assert [o.grid_shape_y, o.grid_shape_x] == list(z['grid_shape'])
    wr = o.valid_rows                                                            # rows the raster loop writes (PMP:233-243)
    assert sha(wr.astype(np.uint8)) == str(z['written_sha'])
    assert sha(o.indices[wr].astype(np.int64)) == str(z['indices_sha'])
    r = o.py_func(F['Ux'], F['Uy'], F['p_prev'])
    assert len(r['origins']) == (r['n_y'] + 2) * (r['n_x'] + 2)                  # the extra -1 column
    kept_ref = z['p'] == F['p_prev']
    kept = r['p'] == F['p_prev']
    assert np.array_equal(kept, kept_ref)
    scale = np.abs(z['p'] - F['p_prev']).max()
    np.testing.assert_allclose(r['p'], z['p'], rtol=0, atol=5e-6 * scale)


def test_pressure_recovery_oracle_matches_reference_bit_for_bit():
    """oracle/integrate.py against the unmodified reference: Evaluation.integrate_field called directly on a seeded block
    (all four direction pairs) and the pressure field its timeStep produced after the quadrant stitch (GRAD:371-416, 585-628)."""
    from oracle import integrate as ointeg
    z, mesh_kw, seed = load_golden('grad_integrate')
    sdf = z['sdfunct'][:, :, None]
    bb = z['bbox']
    H, W = z['grid_shape']
    xl, yl = np.linspace(bb[0], bb[1], W), np.linspace(bb[2], bb[3], H)
    blk = np.random.default_rng(int(z['kat_block_seed'])).standard_normal((150, 120, 2))
    k = 0
    for dx_ in (1, -1):
        for dy_ in (1, -1):
            np.testing.assert_array_equal(ointeg.integrate_field(blk.copy(), sdf, xl, yl, dx_, dy_), z['kat'][k])
            k += 1
    small = blk[:24, :40].copy()
    np.testing.assert_array_equal(ointeg.integrate_field_literal(small.copy(), sdf, xl, yl, -1, 1), ointeg.integrate_field(small, sdf, xl, yl, -1, 1))
    p, cx = ointeg.recover_pressure(z['dp_dx'], z['dp_dy'], sdf, bb[0], bb[1], bb[2], bb[3], float(z['x0_min']), 5e-3)
    np.testing.assert_array_equal(p, z['p_field'])
