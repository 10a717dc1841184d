"""The C++ block-row partitioner (csrc/psm_partition.cpp, psm_shard_build) against psm_b200/shard.py partition(): every array of
every rank's shard must be identical -- a C / C++ caller shards a mesh without the Python shim."""
import ctypes as C

import numpy as np
import pytest

import psm_b200
from psm_b200 import synthetic as syn, tables as ptables, shard as pshard, _capi as capi
from psm_b200.surrogate import _marshal_tables, _VARIANT_CODE


def _arr(ptr, n, dtype):
    if n == 0 or not ptr:
        return np.zeros(0, dtype)
    return np.ctypeslib.as_array(ptr, shape=(n,)).astype(dtype, copy=True)


@pytest.mark.parametrize("variant,mesh_kw,world,near", [('deltaU_to_deltaP', dict(H=500, W=420, nx=220, ny=260, R=0.12), 2, 0.0),
                                                        ('deltaU_to_deltaP', dict(H=500, W=420, nx=220, ny=260, R=0.12), 3, 0.05),
                                                        ('U_to_gradP', dict(H=340, W=300, nx=150, ny=170, R=0.1), 2, 0.0)])
def test_cpp_partitioner_equals_python_partition(variant, mesh_kw, world, near):
    lib = capi.load()
    mesh = syn.make_mesh(seed=4, **mesh_kw)
    F = syn.make_fields(mesh, seed=4)
    t = ptables.build_tables(mesh['cells'], mesh['top'], mesh['obst'], F['p_prev'], variant=variant)
    overlap = 32 if variant == 'deltaU_to_deltaP' else 96
    ref = pshard.partition(t, mesh['cells'], world, variant=variant, near_wall_sdf=near)
    T, keep = _marshal_tables(t)
    cells = np.ascontiguousarray(mesh['cells'], dtype=np.float64)
    for rank in range(world):
        S = C.c_void_p()
        rc = lib.psm_shard_build(C.byref(T), cells.ctypes.data_as(C.POINTER(C.c_double)), 2, float(t['bbox'][2]), float(t['delta']),
                                 _VARIANT_CODE[variant], 128, overlap, float(near), rank, world, C.byref(S))
        assert rc == 0
        v = lib.psm_shard_view(S).contents
        r = ref[rank]
        for k in ('rank', 'world', 'row0', 'row1', 'ext_rows', 'send_rows', 'blk_row0', 'blk_row1', 'local_ext_rows', 'n_owned', 'n_ghost',
                  'n_ghost_pix'):
            assert getattr(v, k) == r[k], (k, getattr(v, k), r[k])
        assert (v.grid_h, v.grid_w) == (r['H'], r['W'])
        nq = (r['row1'] - r['row0'] + r['local_ext_rows']) * r['W']
        np.testing.assert_array_equal(_arr(v.vert, nq * 3, np.int32).reshape(nq, 3), r['vert'])
        np.testing.assert_array_equal(_arr(v.weights, nq * 3, np.float64).reshape(nq, 3), r['weights'])
        np.testing.assert_array_equal(_arr(v.sdfunct, r['sdfunct'].size, np.float64).reshape(r['sdfunct'].shape), r['sdfunct'])
        np.testing.assert_array_equal(_arr(v.mask_global, r['H'] * r['W'], np.uint8).reshape(r['H'], r['W']), r['mask_global'])
        np.testing.assert_array_equal(_arr(v.vert_back, r['n_owned'] * 3, np.int32).reshape(-1, 3), r['vert_back'])
        np.testing.assert_array_equal(_arr(v.weights_back, r['n_owned'] * 3, np.float64).reshape(-1, 3), r['weights_back'])
        for k in ('cell_send_ptr', 'cell_recv_ptr', 'pix_send_ptr', 'pix_recv_ptr'):
            np.testing.assert_array_equal(_arr(getattr(v, k), world + 1, np.int64), r[k])
        np.testing.assert_array_equal(_arr(v.cell_send_idx, int(r['cell_send_ptr'][-1]), np.int32), r['cell_send_idx'])
        np.testing.assert_array_equal(_arr(v.pix_send_idx, int(r['pix_send_ptr'][-1]), np.int32), r['pix_send_idx'])
        np.testing.assert_array_equal(_arr(v.ghost_pix, r['n_ghost_pix'], np.int64), r['ghost_pix'])
        own_p, rank_p = capi.c_int64_p(), capi.c_int32_p()
        assert lib.psm_shard_cells(S, C.byref(own_p), C.byref(rank_p)) == 0
        np.testing.assert_array_equal(_arr(own_p, r['n_owned'], np.int64), r['owned_ids'])
        np.testing.assert_array_equal(_arr(rank_p, mesh['cells'].shape[0], np.int32), r['cell_rank'])
        lib.psm_shard_free(S)


def test_cpp_partitioner_rejects_too_many_ranks():
    lib = capi.load()
    mesh = syn.make_mesh(seed=4, **syn.CONFIGS['tiny'])
    t = ptables.build_tables(mesh['cells'], mesh['top'], mesh['obst'], np.zeros(mesh['cells'].shape[0]))
    T, keep = _marshal_tables(t)
    cells = np.ascontiguousarray(mesh['cells'], dtype=np.float64)
    S = C.c_void_p()
    assert lib.psm_shard_build(C.byref(T), cells.ctypes.data_as(C.POINTER(C.c_double)), 2, float(t['bbox'][2]), 5e-3, 0, 128, 32, 0.0, 0, 64,
                               C.byref(S)) == -3
