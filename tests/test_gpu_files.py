"""INTEGRATION.md route B on the GPU: a handle initialised from the flat files, and the plain-C driver
(examples/c_driver/psm_driver.c: C-ABI only, no Python / PyTorch in the process), must reproduce the
array-initialised handle bit for bit."""
import os
import subprocess

import numpy as np
import pytest

import psm_b200
from psm_b200 import synthetic as syn, tables as ptables

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def case(tmp_path_factory):
    d = tmp_path_factory.mktemp('route_b')
    mesh = syn.make_mesh(seed=6, **syn.CONFIGS['tiny'])
    F = syn.make_fields(mesh, seed=6)
    params = syn.make_params(seed=6, pc_in=40, pc_p=24)
    t = ptables.build_tables(mesh['cells'], mesh['top'], mesh['obst'], F['p_prev'])
    psm_b200.save_tables(t, d / 'tables.bin')
    psm_b200.save_params(params, d / 'params.bin')
    cells = syn.pack_cells(mesh, F)                      # 7 columns: {Ux,Uy,Cx,Cy,p,dUx,dUy}
    with psm_b200.PressureSurrogate('deltaU_to_deltaP') as sm:
        sm.load_params(params)
        sm.init_tables(t)
        ref, rc = sm.predict(cells)
        assert rc == 0
    return d, cells, ref


def test_handle_from_files_is_bit_identical(case):
    d, cells, ref = case
    with psm_b200.PressureSurrogate('deltaU_to_deltaP') as sm:
        sm.load_params_file(d / 'params.bin')
        sm.init_from_file(d / 'tables.bin')
        out, rc = sm.predict(cells)
    assert rc == 0
    np.testing.assert_array_equal(out, ref)


def test_plain_c_driver(case):
    d, cells, ref = case
    exe = os.path.join(REPO, 'examples', 'c_driver', 'psm_driver')
    r = subprocess.run(['make', '-C', os.path.dirname(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    cells.astype(np.float64).tofile(d / 'cells.bin')
    r = subprocess.run([exe, str(d / 'params.bin'), str(d / 'tables.bin'), str(d / 'cells.bin'), str(cells.shape[0]),
                        str(cells.shape[1]), '0', '5', str(d / 'p_out.bin')], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert 'status 0' in r.stdout, r.stdout
    out = np.fromfile(d / 'p_out.bin', dtype=np.float64)
    np.testing.assert_array_equal(out, ref)


def test_corrupt_or_foreign_files_are_rejected_with_a_message(case, tmp_path):
    """A file for another block edge, a truncated payload or a header with absurd sizes must come back as PSM_ERR_INVALID with
    psm_last_error set -- never an out-of-bounds read or an exception across the C boundary."""
    d, cells, ref = case
    good = (d / 'params.bin').read_bytes()

    def load(blob, name):
        path = tmp_path / name
        path.write_bytes(blob)
        with psm_b200.PressureSurrogate('deltaU_to_deltaP') as sm:
            with pytest.raises(psm_b200.PsmError) as e:
                sm.load_params_file(path)
        assert e.value.code == -1 and len(str(e.value)) > 30, str(e.value)
        return str(e.value)

    hdr = np.frombuffer(good[8:32], dtype=np.int32).copy()
    other = hdr.copy(); other[0] = 64                                # written for shape 64: the handle uses 128
    assert 'shape' in load(good[:8] + other.tobytes() + good[32:], 'shape64.bin')
    huge = hdr.copy(); huge[2] = 2 ** 30                             # pc_in that would drive a 10^14-byte allocation
    assert 'pc_in' in load(good[:8] + huge.tobytes() + good[32:], 'huge.bin')
    assert 'payload' in load(good[:-100], 'truncated.bin')
    assert 'PSMPRM01' in load(b'NOTAFILE' + good[8:], 'magic.bin')
    tgood = (d / 'tables.bin').read_bytes()
    for blob, word in ((tgood[:-8], 'payload'), (tgood[:8] + np.int64(-5).tobytes() + tgood[16:], 'range')):
        path = tmp_path / 't.bin'
        path.write_bytes(blob)
        with psm_b200.PressureSurrogate('deltaU_to_deltaP') as sm:
            sm.load_params_file(d / 'params.bin')
            with pytest.raises(psm_b200.PsmError) as e:
                sm.init_from_file(path)
        assert word in str(e.value), str(e.value)


def _read_tables_file(path):
    with open(path, 'rb') as f:
        assert f.read(8) == b'PSMTBL01'
        n = int(np.frombuffer(f.read(8), np.int64)[0])
        H, W, hb, _ = np.frombuffer(f.read(16), np.int32)
        G = H * W
        rd = lambda dt, k: np.frombuffer(f.read(np.dtype(dt).itemsize * k), dt)      # noqa: E731
        d = dict(n=n, H=int(H), W=int(W), vert=rd(np.int32, G * 3).reshape(G, 3), weights=rd(np.float64, G * 3).reshape(G, 3),
                 indices=rd(np.int64, G * 2).reshape(G, 2), sdfunct=rd(np.float64, G).reshape(H, W))
        if hb:
            d['vert_back'], d['weights_back'] = rd(np.int32, n * 3).reshape(n, 3), rd(np.float64, n * 3).reshape(n, 3)
        return d


@pytest.mark.parametrize("variant", ['deltaU_to_deltaP', 'U_to_gradP', 'thesis'])
def test_init_mesh_builds_the_shim_tables_and_serves_them_from_the_cache(variant, tmp_path):
    """psm_init_mesh: bbox / grid / GPU mask + distance kernel / raster / closed-form back tables inside the library, against
    psm_b200.tables (the NumPy / SciPy shim): raster and mask bit-exact, distances to 1e-14 (cKDTree vs the kernel's FP64
    sqrt); then a second handle initialised from the table cache alone -- no cells -> grid tables passed, no Qhull -- must
    reproduce the first one bit for bit."""
    import glob
    deltas = variant == 'deltaU_to_deltaP'
    mesh = syn.make_mesh(seed=9, **syn.CONFIGS['tiny'])
    F = syn.make_fields(mesh, seed=9)
    params = syn.make_params(seed=9, pc_in=32, pc_p=24, standardization='std' if deltas else 'max_abs',
                             n_out_channels=2 if variant == 'U_to_gradP' else 1,
                             maxs=syn.DEFAULT_MAXS if variant != 'U_to_gradP' else (1.0, 0.536, 0.999, 0.8, 0.7))
    probe = F['p_prev'] if deltas else F['Ux']
    t = ptables.build_tables(mesh['cells'], mesh['top'], mesh['obst'], probe, variant=variant, back='closed_form')
    cells = syn.pack_cells(mesh, F, with_delta=deltas)
    with psm_b200.PressureSurrogate(variant) as ref:
        ref.load_params(params)
        ref.init_tables(t)
        out_ref, _ = ref.predict(cells)
        owner_ref = ref.owner_map()
    with psm_b200.PressureSurrogate(variant) as a:
        a.load_params(params)
        a.init_mesh(mesh['cells'], mesh['top'], mesh['obst'], probe, back='closed_form', cache_dir=tmp_path)
        out_a, _ = a.predict(cells)
        np.testing.assert_array_equal(a.owner_map(), owner_ref)
    files = glob.glob(str(tmp_path / 'psm_tables_*_cf.bin'))
    assert len(files) == 1
    d = _read_tables_file(files[0])
    assert (d['H'], d['W'], d['n']) == (t['H'], t['W'], t['n_cells'])
    np.testing.assert_array_equal(d['indices'], t['indices'])
    np.testing.assert_array_equal(d['sdfunct'] != 0, t['sdfunct'] != 0)
    np.testing.assert_allclose(d['sdfunct'], t['sdfunct'], rtol=0, atol=1e-14)
    np.testing.assert_array_equal(d['vert_back'], t['vert_back'])
    np.testing.assert_array_equal(d['weights_back'], t['weights_back'])
    ok = ~np.isnan(out_ref)
    assert np.array_equal(np.isnan(out_a), np.isnan(out_ref))
    np.testing.assert_allclose(out_a[ok], out_ref[ok], rtol=0, atol=1e-6 * np.abs(out_ref[ok]).max())
    with psm_b200.PressureSurrogate(variant) as b:                  # cache hit: psm_mesh with vert == NULL
        b.load_params(params)
        M = psm_b200._capi.PsmMesh()
        import ctypes as C
        cxy, top, obst, pr = (np.ascontiguousarray(x, dtype=np.float64) for x in (mesh['cells'], mesh['top'], mesh['obst'], probe))
        M.n_cells, M.cells_xy, M.xy_stride, M.back_closed_form = cxy.shape[0], cxy.ctypes.data_as(C.POINTER(C.c_double)), 2, 1
        M.top, M.n_top, M.obst, M.n_obst = top.ctypes.data_as(C.POINTER(C.c_double)), top.shape[0], obst.ctypes.data_as(C.POINTER(C.c_double)), obst.shape[0]
        M.probe, M.cache_dir = pr.ctypes.data_as(C.POINTER(C.c_double)), str(tmp_path).encode()
        b._check(b.lib.psm_init_mesh(b._h, C.byref(M)))
        b._after_init()
        out_b, _ = b.predict(cells)
    np.testing.assert_array_equal(out_b, out_a)
    with psm_b200.PressureSurrogate(variant) as c:                  # miss and no tables: a clear error, not a crash
        c.load_params(params)
        M.cache_dir = str(tmp_path / 'nowhere').encode()
        rc = c.lib.psm_init_mesh(c._h, C.byref(M))
        assert rc == -4 and b'cache' in c.lib.psm_last_error(c._h)


def test_plain_c_driver_initialises_from_raw_arrays_and_native_fields(case, tmp_path):
    """The C program with no Python in the process: psm_init_mesh from the raw arrays (cell centres, "top" and "obstacle" points)
    with the Delaunay tables served by the table cache, every step through psm_predict_fields on U double[n][3] + p double[n].
    Same pressures as the Python-initialised handle fed the same (closed-form back) tables."""
    d, cells, ref = case
    mesh = syn.make_mesh(seed=6, **syn.CONFIGS['tiny'])
    F = syn.make_fields(mesh, seed=6)
    params = syn.make_params(seed=6, pc_in=40, pc_p=24)
    cache = tmp_path / 'cache'
    cache.mkdir()
    with psm_b200.PressureSurrogate('deltaU_to_deltaP') as sm:            # first run of the case: fills the cache (Qhull once)
        sm.load_params(params)
        sm.init_mesh(mesh['cells'], mesh['top'], mesh['obst'], F['p_prev'], back='closed_form', cache_dir=cache)
        expect, rc = sm.predict(cells)
        assert rc == 0
    exe = os.path.join(REPO, 'examples', 'c_driver', 'psm_driver')
    r = subprocess.run(['make', '-C', os.path.dirname(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    cells.astype(np.float64).tofile(tmp_path / 'cells.bin')
    mesh['top'].astype(np.float64).tofile(tmp_path / 'top.bin')
    mesh['obst'].astype(np.float64).tofile(tmp_path / 'obst.bin')
    env = dict(os.environ, PSM_DRIVER_CACHE=str(cache), PSM_DRIVER_TOP=str(tmp_path / 'top.bin'), PSM_DRIVER_OBST=str(tmp_path / 'obst.bin'),
               PSM_DRIVER_FIELDS='1')
    r = subprocess.run([exe, str(d / 'params.bin'), 'unused', str(tmp_path / 'cells.bin'), str(cells.shape[0]), str(cells.shape[1]), '0', '4',
                        str(tmp_path / 'p_out.bin')], capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stdout + r.stderr
    assert 'status 0' in r.stdout, r.stdout
    np.testing.assert_array_equal(np.fromfile(tmp_path / 'p_out.bin', dtype=np.float64), expect)
