"""B200-native pressure-surrogate hot path (deltaU_to_deltaP / U_to_gradP).

Host-side mirror of the reference interface over the C-ABI library ``libpsm_b200.so``:
``PressureSurrogate`` (handle), ``tables.build_tables`` (once-per-mesh Qhull tables),
``synthetic`` (seeded inputs), and -- one directory up -- ``python_module.py`` with the
reference's ``init_func`` / ``py_func`` names.  No CPU fallback exists.
"""
from .surrogate import PressureSurrogate, compile_plan, debug_gemm, debug_dense_stack, save_params, save_tables, register_host_buffer, unregister_host_buffer   # noqa: F401
from . import synthetic, tables                           # noqa: F401
from ._capi import PsmError, PSM_OK, PSM_SKIPPED          # noqa: F401
