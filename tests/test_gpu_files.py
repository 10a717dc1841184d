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
