"""Host-side table builders (psm_b200/tables.py) that have no GPU dependency."""
import numpy as np

from psm_b200 import synthetic as syn, tables as ptables
from oracle import interp as ointerp


def test_closed_form_back_tables_interpolate_linear_fields_exactly():
    """`regular_grid_back_tables` (opt-in for huge meshes): barycentric weights on the fixed-diagonal triangulation of
    the regular grid -- non-negative, summing to one, exact for linear fields, and equal to what the Qhull tables of
    PMP:211 give on any linear field (every triangulation interpolates a linear function exactly)."""
    mesh = syn.make_mesh(seed=8, **syn.CONFIGS['tiny'])
    cells = mesh['cells']
    x_min, x_max = round(float(cells[:, 0].min()), 3), round(float(cells[:, 0].max()), 3)
    y_min, y_max = round(float(cells[:, 1].min()), 3), round(float(cells[:, 1].max()), 3)
    X0, Y0 = ptables.uniform_grid(x_min, x_max, y_min, y_max, 5e-3)
    W = int(round((x_max - x_min) / 5e-3))
    vert, wts = ptables.regular_grid_back_tables(cells, X0[:W], Y0[::W], W)
    inside = ~np.any(wts < 0, axis=1)
    assert inside.mean() > 0.95
    np.testing.assert_allclose(wts[inside].sum(axis=1), 1.0, rtol=0, atol=1e-12)
    assert (wts[inside] >= 0).all() and (wts[inside] <= 1 + 1e-12).all()
    lin = 0.7 - 1.3 * X0 + 0.45 * Y0
    got = np.einsum('nj,nj->n', np.take(lin, vert), wts)
    want = 0.7 - 1.3 * cells[:, 0] + 0.45 * cells[:, 1]
    np.testing.assert_allclose(got[inside], want[inside], rtol=0, atol=1e-12)
    # the same cells are inside / outside as with the reference's Qhull tables, and both reproduce the linear field
    xy0 = np.stack([X0, Y0], axis=1)
    vq, wq = ointerp.interp_weights(xy0, cells)
    inside_q = ~np.any(wq < -1e-12, axis=1)
    assert (inside == inside_q).mean() > 0.999          # cells exactly on the hull may flip
    both = inside & inside_q
    np.testing.assert_allclose(np.einsum('nj,nj->n', np.take(lin, vq), wq)[both], got[both], rtol=0, atol=1e-11)


def test_build_tables_variants_share_the_forward_tables_but_not_the_mask_rules():
    """deltaU_to_deltaP rounds the bbox to 3 decimals and sub-samples the walls [::5] (SMC:102-106,138-139); the thesis
    module rounds to 2 decimals and sub-samples [::10] (PMP:197-201,93-94): same mesh, generally different rasters."""
    mesh = syn.make_mesh(seed=9, **syn.CONFIGS['tiny'])
    F = syn.make_fields(mesh, seed=9)
    a = ptables.build_tables(mesh['cells'], mesh['top'], mesh['obst'], F['p_prev'], variant='deltaU_to_deltaP', back=None)
    b = ptables.build_tables(mesh['cells'], mesh['top'], mesh['obst'], F['Ux'], variant='thesis', back=None)
    assert a['vert'].shape[1] == b['vert'].shape[1] == 3
    assert (a['H'], a['W']) == (b['H'], b['W'])         # this mesh's bbox is representable with 2 decimals
    assert a['vert_back'] is None and b['vert_back'] is None
    assert np.array_equal(a['sdfunct'] != 0, b['sdfunct'] != 0) or (a['sdfunct'] != 0).sum() != (b['sdfunct'] != 0).sum()
    assert np.abs(a['sdfunct'] - b['sdfunct']).max() < 0.05     # different wall sub-sampling: close, not identical
