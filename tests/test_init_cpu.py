"""Host-only pieces of psm_init_mesh (csrc/psm_init.cu) against the NumPy shim (psm_b200/tables.py): no GPU needed."""
import numpy as np
import pytest

import psm_b200
from psm_b200 import synthetic as syn, tables as ptables
from psm_b200.surrogate import mesh_hash, mesh_grid, back_tables_closed_form


@pytest.mark.parametrize("variant", ['deltaU_to_deltaP', 'U_to_gradP', 'thesis'])
@pytest.mark.parametrize("mesh_kw", [syn.CONFIGS['tiny'], dict(H=240, W=340, nx=130, ny=90, R=0.1), dict(H=397, W=2998, nx=300, ny=40, R=0.4)])
def test_bbox_rounding_and_grid_shape_match_the_shim(variant, mesh_kw):
    """round(x, 3) / round(x, 2) (SMC:102-106 / PMP:197-201) and int(round((max - min) / delta)) (SMC:148-149) in C."""
    mesh = syn.make_mesh(seed=3, **mesh_kw)
    cells = mesh['cells'] + 1.2345e-4                     # off the lattice: the rounding has something to do
    nd = 3 if variant == 'deltaU_to_deltaP' else 2
    ref = (round(float(cells[:, 0].min()), nd), round(float(cells[:, 0].max()), nd),
           round(float(cells[:, 1].min()), nd), round(float(cells[:, 1].max()), nd))
    bbox, H, W = mesh_grid(variant, 5e-3, cells)
    assert bbox == ref
    assert (H, W) == (int(round((ref[3] - ref[2]) / 5e-3)), int(round((ref[1] - ref[0]) / 5e-3)))


def test_python_round_half_cases():
    """Values that sit on a decimal tie in binary: C must round them like Python's round (correctly rounded decimal)."""
    for v in (0.0005, 0.0015, 0.0025, 2.675, 1.005, -0.0045, 0.125, 0.375, 7.4995):
        cells = np.array([[v, v], [v + 1.0, v + 2.0], [v + 0.5, v + 0.25]])
        for variant, nd in (('deltaU_to_deltaP', 3), ('U_to_gradP', 2)):
            bbox, _, _ = mesh_grid(variant, 5e-3, cells)
            assert bbox[0] == round(float(v), nd) and bbox[1] == round(float(v + 1.0), nd), (v, nd, bbox)


def test_closed_form_back_tables_bit_identical():
    mesh = syn.make_mesh(seed=5, **syn.CONFIGS['tiny'])
    t = ptables.build_tables(mesh['cells'], mesh['top'], mesh['obst'], np.zeros(mesh['cells'].shape[0]), back='closed_form')
    X0, Y0 = ptables.uniform_grid(*t['bbox'], t['delta'])
    W = t['W']
    cells = mesh['cells'].copy()
    cells[:7] += 10.0                                      # some cells outside the grid hull: the (-1, 1, 1) marker
    vb_ref, wb_ref = ptables.regular_grid_back_tables(cells, X0[:W], Y0[::W], W)
    vb, wb = back_tables_closed_form(cells, X0[:W], Y0[::W])
    np.testing.assert_array_equal(vb, vb_ref)
    np.testing.assert_array_equal(wb, wb_ref)


def test_mesh_hash_is_stable_and_sensitive():
    mesh = syn.make_mesh(seed=6, **syn.CONFIGS['tiny'])
    p = np.arange(mesh['cells'].shape[0], dtype=np.float64)
    a = mesh_hash('deltaU_to_deltaP', 5e-3, mesh['cells'], mesh['top'], mesh['obst'], p)
    assert len(a) == 16 and a == mesh_hash('deltaU_to_deltaP', 5e-3, mesh['cells'].copy(), mesh['top'], mesh['obst'], p)
    moved = mesh['cells'].copy()
    moved[11, 1] = np.nextafter(moved[11, 1], 1.0)
    others = {mesh_hash('deltaU_to_deltaP', 5e-3, moved, mesh['top'], mesh['obst'], p),
              mesh_hash('U_to_gradP', 5e-3, mesh['cells'], mesh['top'], mesh['obst'], p),
              mesh_hash('deltaU_to_deltaP', 4e-3, mesh['cells'], mesh['top'], mesh['obst'], p),
              mesh_hash('deltaU_to_deltaP', 5e-3, mesh['cells'], mesh['top'][::-1], mesh['obst'], p),
              mesh_hash('deltaU_to_deltaP', 5e-3, mesh['cells'], mesh['top'], mesh['obst'], np.where(p == 3, np.nan, p))}
    assert a not in others and len(others) == 5
    # a finite probe decides nothing (SMC:165-169): its values stay out of the key
    assert a == mesh_hash('deltaU_to_deltaP', 5e-3, mesh['cells'], mesh['top'], mesh['obst'], p + 1.0)
    assert a == mesh_hash('deltaU_to_deltaP', 5e-3, mesh['cells'], mesh['top'], mesh['obst'], None)
