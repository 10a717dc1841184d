"""Drop-in for the solver-side ``python_module`` the OpenFOAM solver imports
(Thesis_Work/Chapter5/parallelized/test_case/python_module.py: ``init_func`` PMP:172-247,
``py_func`` PMP:249-517; called from FOAM/PythonComm_init.H:94 and FOAM/PythonComm.H:24-27).

Same two function names, same argument meaning, same return shape -- but the per-step body is
one call into the CUDA library (``psm_predict``) instead of NumPy/TensorFlow on rank 0.

Artefacts are looked up like the reference does, in the current directory, but from ONE file:
``psm_params.npz`` (written by ``psm_b200.params.save_npz``), or set ``PSM_PARAMS`` to its path.
Environment: ``PSM_VARIANT`` (deltaU_to_deltaP | U_to_gradP | thesis = the arithmetic of the reference's own
module: U -> p, avance 12, near-wall fallback 0.05), ``PSM_DEVICE`` (CUDA ordinal),
``PSM_INPUT_COLS`` (5 | 7).

Differences from the reference module, all on the boundary rather than in the arithmetic:
  * the MPI gather-to-root (PMP:179-185,258,511) is gone: every rank owns a GPU handle;
  * errors raise ``PsmError`` instead of returning 0 and crashing the caller (PMS:440-444);
  * with 5 input columns the deltaU variant keeps U(t-1) on the device and forms dU itself; the
    first call therefore returns p_prev (status PSM_SKIPPED).
"""
import os

import numpy as np

from psm_b200 import PressureSurrogate, tables as _tables
from psm_b200 import params as _params

_state = {'sm': None}


def _surrogate():
    if _state['sm'] is None:
        variant = os.environ.get('PSM_VARIANT', 'deltaU_to_deltaP')
        cols = int(os.environ.get('PSM_INPUT_COLS', '5'))
        sm = PressureSurrogate(variant=variant, device=int(os.environ.get('PSM_DEVICE', '0')), input_cols=cols,
                               near_wall_sdf=float(os.environ.get('PSM_NEAR_WALL_SDF', '0.05' if variant == 'thesis' else '0')))   # PMP:492-494
        sm.load_params(_params.load_npz(os.environ.get('PSM_PARAMS', 'psm_params.npz')))
        _state['sm'] = sm
    return _state['sm']


def set_surrogate(sm):
    """Install an already-configured handle (tests, embedding applications)."""
    _state['sm'] = sm


def init_func(array, top_boundary, obst_boundary, placeholder=None):
    """PMP:172-247.  ``array`` [nCells, >=5] = {Ux, Uy, Cx, Cy, p, ...}; builds the Qhull tables,
    mask, distance field and raster once and uploads them.  Returns 0 like the reference."""
    sm = _surrogate()
    array = np.asarray(array, dtype=np.float64)
    probe = array[:, 4] if sm.variant == 'deltaU_to_deltaP' else array[:, 0]     # SMC:165 p ; PMP:230 ux
    t = _tables.build_tables(array[:, 2:4], np.asarray(top_boundary, dtype=np.float64),
                             np.asarray(obst_boundary, dtype=np.float64), probe, variant=sm.variant, delta=sm.delta)
    sm.init_tables(t)
    return 0


def py_func(array_in, placeholder=None):
    """PMP:249-517.  Returns p [nCells] float64 (deltaU_to_deltaP: p_prev + delta_p) or
    grad p [nCells, 2] (U_to_gradP)."""
    out, _ = _surrogate().predict(array_in)
    return out
