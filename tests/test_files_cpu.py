"""Flat binary containers of INTEGRATION.md route B (psm_save_tables / psm_save_params): host-only round trip."""
import numpy as np

import psm_b200
from psm_b200 import synthetic as syn, tables as ptables


def _read(f, dtype, n):
    a = np.frombuffer(f.read(np.dtype(dtype).itemsize * n), dtype=dtype)
    assert a.size == n
    return a


def test_tables_file_round_trip(tmp_path):
    mesh = syn.make_mesh(seed=2, **syn.CONFIGS['tiny'])
    F = syn.make_fields(mesh, seed=2)
    t = ptables.build_tables(mesh['cells'], mesh['top'], mesh['obst'], F['p_prev'])
    path = tmp_path / 'tables.bin'
    psm_b200.save_tables(t, path)
    G, N = t['H'] * t['W'], t['n_cells']
    with open(path, 'rb') as f:
        assert f.read(8) == b'PSMTBL01'
        assert _read(f, np.int64, 1)[0] == N
        assert tuple(_read(f, np.int32, 4)) == (t['H'], t['W'], 1, 0)
        np.testing.assert_array_equal(_read(f, np.int32, G * 3).reshape(G, 3), t['vert'])
        np.testing.assert_array_equal(_read(f, np.float64, G * 3).reshape(G, 3), t['weights'])
        np.testing.assert_array_equal(_read(f, np.int64, G * 2).reshape(G, 2), t['indices'])
        np.testing.assert_array_equal(_read(f, np.float64, G).reshape(t['H'], t['W']), t['sdfunct'])
        np.testing.assert_array_equal(_read(f, np.int32, N * 3).reshape(N, 3), t['vert_back'])
        np.testing.assert_array_equal(_read(f, np.float64, N * 3).reshape(N, 3), t['weights_back'])
        assert f.read() == b''


def test_params_file_round_trip(tmp_path):
    p = syn.make_params(seed=3, pc_in=20, pc_p=12)
    path = tmp_path / 'params.bin'
    psm_b200.save_params(p, path)
    S2 = 128 * 128
    with open(path, 'rb') as f:
        assert f.read(8) == b'PSMPRM01'
        shape, n_out, pc_in, pc_p, std, n_dense = _read(f, np.int32, 6)
        assert (shape, n_out, pc_in, pc_p, std, n_dense) == (128, 1, 20, 12, 0, len(p['mlp_weights']))
        np.testing.assert_array_equal(_read(f, np.float64, 5)[:4], np.asarray(p['maxs'], np.float64)[:4])
        _read(f, np.float64, 2)
        dims = _read(f, np.int32, n_dense + 1)
        assert list(dims) == [20] + [w.shape[1] for w in p['mlp_weights']]
        np.testing.assert_array_equal(_read(f, np.float64, pc_in * S2 * 3).reshape(pc_in, -1), p['pca_in_components'])
        np.testing.assert_array_equal(_read(f, np.float64, S2 * 3), p['pca_in_mean'])
        np.testing.assert_array_equal(_read(f, np.float64, pc_p * S2).reshape(pc_p, -1), p['pca_out_components'])
        np.testing.assert_array_equal(_read(f, np.float64, S2), p['pca_out_mean'])
        for name, n in (('mean_in', pc_in), ('std_in', pc_in), ('mean_out', pc_p), ('std_out', pc_p)):
            np.testing.assert_array_equal(_read(f, np.float64, n), p[name])
        for w, b in zip(p['mlp_weights'], p['mlp_biases']):
            np.testing.assert_array_equal(_read(f, np.float32, w.size).reshape(w.shape), w)
            np.testing.assert_array_equal(_read(f, np.float32, b.size), b)
        assert f.read() == b''


def test_tables_file_cli(tmp_path):
    from psm_b200 import tables_file
    mesh = syn.make_mesh(seed=4, **syn.CONFIGS['tiny'])
    F = syn.make_fields(mesh, seed=4)
    np.save(tmp_path / 'cells.npy', syn.pack_cells(mesh, F, with_delta=False))
    np.save(tmp_path / 'top.npy', mesh['top'])
    np.save(tmp_path / 'obst.npy', mesh['obst'])
    out = tmp_path / 't.bin'
    assert tables_file.main(['--cells', str(tmp_path / 'cells.npy'), '--top', str(tmp_path / 'top.npy'),
                             '--obst', str(tmp_path / 'obst.npy'), '--out', str(out)]) == 0
    t = ptables.build_tables(mesh['cells'], mesh['top'], mesh['obst'], F['p_prev'])
    ref = tmp_path / 'ref.bin'
    psm_b200.save_tables(t, ref)
    assert out.read_bytes() == ref.read_bytes()
