"""Once-per-mesh tables (and, optionally, the model artefacts) as flat binary files for C/C++ callers.

  python -m psm_b200.tables_file --cells cells.npy --top top.npy --obst obst.npy --out psm_tables.bin \
         [--variant deltaU_to_deltaP] [--delta 5e-3] [--back qhull|closed_form|none] [--params psm_params.npz --params-out psm_params.bin]
         [--cache-dir psm_cache]

With ``--cache-dir`` the table file is ALSO stored in the table cache under the hash of the mesh (``psm_mesh_hash``), where
``psm_init_mesh`` of a C/C++ caller finds it: the solver then initialises from its raw arrays (cell centres, "top" and
"obstacle" patch points) with no file name to agree on and no interpreter.

``cells`` is the solver's ``double[nCells][>=4]`` array ``{Ux,Uy,Cx,Cy,...}`` (FOAM/PythonComm_init.H:53-60) or just the
``[nCells][2]`` cell centres; ``top`` / ``obst`` are the boundary-face centres of the patches named "top" and "obstacle"
(init.H:33-37,62-91).  The tables are built exactly like ``init_func`` does (SciPy Qhull, PMP:203-243 / SMC:110-178) and
written with ``psm_save_tables``; a solver then calls ``psm_init_from_file`` and needs neither SciPy nor an interpreter
(INTEGRATION.md route B).
"""
import argparse

import numpy as np

from . import params as _params, tables as _tables
from .surrogate import save_params, save_tables


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split('\n')[0])
    ap.add_argument('--cells', required=True)
    ap.add_argument('--top', required=True)
    ap.add_argument('--obst', required=True)
    ap.add_argument('--out', required=True)
    ap.add_argument('--variant', default='deltaU_to_deltaP', choices=['deltaU_to_deltaP', 'U_to_gradP'])
    ap.add_argument('--delta', type=float, default=5e-3)
    ap.add_argument('--back', default='qhull', choices=['qhull', 'closed_form', 'none'])
    ap.add_argument('--params', default=None, help='npz written by psm_b200.params.save_npz')
    ap.add_argument('--params-out', default=None)
    ap.add_argument('--cache-dir', default=None)
    a = ap.parse_args(argv)
    cells = np.load(a.cells)
    xy = cells[:, 2:4] if cells.shape[1] >= 4 else cells[:, :2]
    probe = cells[:, 4] if cells.shape[1] >= 5 else np.zeros(len(xy))
    t = _tables.build_tables(np.ascontiguousarray(xy), np.load(a.top), np.load(a.obst), probe, variant=a.variant, delta=a.delta,
                             back=None if a.back == 'none' else a.back)
    save_tables(t, a.out)
    print('%s: %d cells, grid %d x %d' % (a.out, t['n_cells'], t['H'], t['W']))
    if a.cache_dir and a.back != 'none':
        import os
        from .surrogate import mesh_hash
        os.makedirs(a.cache_dir, exist_ok=True)
        key = mesh_hash(a.variant, a.delta, np.ascontiguousarray(xy), np.load(a.top), np.load(a.obst), probe)
        path = os.path.join(a.cache_dir, 'psm_tables_%s_%s.bin' % (key, 'cf' if a.back == 'closed_form' else 'qh'))
        save_tables(t, path)
        print(path)
    if a.params:
        save_params(_params.load_npz(a.params), a.params_out or 'psm_params.bin')
        print(a.params_out or 'psm_params.bin')
    return 0


if __name__ == '__main__':
    raise SystemExit(main())
