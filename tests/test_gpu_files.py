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
