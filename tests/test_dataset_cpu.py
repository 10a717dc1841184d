"""Dataset wire format reader (psm_b200/dataset.py) against the reference's own slicing (UTL:57-104, SMC:100-125, 382-402)."""
import numpy as np
import pytest

from psm_b200 import dataset as ds, synthetic as syn


def _reference_index(array, item):
    """utils.index, UTL:94-104, literally."""
    for idx, val in np.ndenumerate(array):
        if val == item:
            return idx


@pytest.mark.parametrize("variant,ncol", [('deltaU_to_deltaP', 11), ('U_to_gradP', 8), ('thesis', 6)])
def test_frame_matches_reference_slicing(variant, ncol):
    mesh = syn.make_mesh(seed=2, **syn.CONFIGS['tiny'])
    n = mesh['cells'].shape[0]
    rng = np.random.default_rng(5)
    cols = rng.standard_normal((n, ncol))
    cols[:, 3:5] = mesh['cells']
    sim_data = np.concatenate([ds.write_padded(cols, n + 37)] * 3, axis=1)           # 3 time frames, like [sim, time, row, C]
    sim_data[0, 1, :n] *= 2.0
    top_b = np.concatenate([ds.write_padded(mesh['top'], mesh['top'].shape[0] + 11)] * 3, axis=1)
    obst_b = np.concatenate([ds.write_padded(mesh['obst'], mesh['obst'].shape[0] + 5)] * 3, axis=1)
    fr = ds.read_frame(sim_data, top_b, obst_b, 0, 1, variant)
    # the reference: data = f['sim_data'][sim:sim+1, time:time+1]; indice = index(data[0,0,:,0], -100.0)[0]; data[0,0,:indice,c]
    data = sim_data[0:1, 1:2]
    indice = _reference_index(data[0, 0, :, 0], -100.0)[0]
    assert fr['n_cells'] == indice == n
    np.testing.assert_array_equal(fr['Ux'], data[0, 0, :indice, 0].astype(np.float64))
    np.testing.assert_array_equal(fr['cells_xy'], data[0, 0, :indice, 3:5].astype(np.float64))
    assert fr['top'].shape[0] == _reference_index(top_b[0, 1, :, 0], -100.0)[0] == mesh['top'].shape[0]
    assert fr['obst'].shape[0] == mesh['obst'].shape[0]
    rows = ds.solver_rows(fr, variant)
    assert rows.shape == (n, 7 if variant == 'deltaU_to_deltaP' else 5)
    if variant == 'deltaU_to_deltaP':
        np.testing.assert_array_equal(rows[:, 4], fr['p'] - fr['dp'])                # p_prev inferred as SMC:639-641
        np.testing.assert_array_equal(rows[:, 5], data[0, 0, :n, 5].astype(np.float64))
        ch = ds.deltaU_change(fr)                                                    # SMC:396-398
        delta_U, delta_U_prev = data[0, 0, :n, 5:7].astype(np.float64), data[0, 0, :n, 8:10].astype(np.float64)
        ref = np.abs(delta_U - delta_U_prev).sum(axis=-1)
        np.testing.assert_array_equal(ch, ref / ref.max())
    with pytest.raises(ValueError):
        ds.read_frame(sim_data[..., :4], top_b, obst_b, 0, 0, variant)


def test_unpadded_arrays_are_taken_whole():
    a = np.arange(12, dtype=np.float32).reshape(1, 1, 4, 3)
    assert ds.first_pad(a[0, 0, :, 0]) == 4
