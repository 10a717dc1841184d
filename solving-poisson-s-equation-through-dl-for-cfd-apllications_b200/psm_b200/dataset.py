"""Dataset wire format of the reference's offline evaluation, read into what the C-ABI takes.

The reference stores every simulation in one HDF5 file (``Thesis_Work/Chapter4/MLP/M_fU/DataGen+Training/data_generation/
data_generation.py:64-74``; read back by ``utils.read_dataset``, UTL:57-71):

  ``sim_data``   float32 [n_sims, n_times, 200000, C]   one row per cell, rows beyond the mesh padded with -100.0 (``padding``,
                                                         data_generation.py:7-12)
  ``top_bound``  float32 [n_sims, n_times, 20000, 2]    x, y of the "top" patch face centres, padded with -100.0
  ``obst_bound`` float32 [n_sims, n_times, 20000, 2]    x, y of the "obstacle" patch

and cuts the padding with ``utils.index(data[0, 0, :, 0], -100.0)[0]`` -- the first row whose first column equals -100.0
(UTL:94-104; SMC:100, 118, 125).  Column meaning per variant:

  deltaU_to_deltaP (SMC:386-402, 102-114): 0 Ux, 1 Uy, 2 p, 3 x, 4 y, 5-6 delta_U, 7 delta_p, 8-9 delta_U_prev, 10 delta_p_prev
  U_to_gradP       (GRAD:432-436, 174-185): 0 Ux, 1 Uy, 2 p, 3 x, 4 y, 6 dP/dx, 7 dP/dy
  thesis / Chapter 4 (data_generation.py:37-59):  0 Ux, 1 Uy, 2 p, 3 x, 4 y, 5 f_U

This module only maps arrays: any object indexable like an h5py dataset works (h5py datasets, NumPy arrays, memmaps), so the
container library stays the caller's choice (h5py is not part of this image; ``open_hdf5`` uses it when it is installed).
"""
import numpy as np

PAD = -100.0                                             # data_generation.py:9

COLUMNS = {
    'deltaU_to_deltaP': dict(Ux=0, Uy=1, p=2, x=3, y=4, dUx=5, dUy=6, dp=7, dUx_prev=8, dUy_prev=9, dp_prev=10),
    'U_to_gradP': dict(Ux=0, Uy=1, p=2, x=3, y=4, dpdx=6, dpdy=7),
    'thesis': dict(Ux=0, Uy=1, p=2, x=3, y=4, f_U=5),
}


def first_pad(column, pad=PAD):
    """``utils.index(array, item)[0]`` (UTL:94-104): index of the first element equal to ``pad``; the full length if none
    (the reference would return None and slice ``[:None]`` -- the whole array)."""
    hit = np.flatnonzero(np.asarray(column) == pad)
    return int(hit[0]) if hit.size else int(np.asarray(column).shape[0])


def read_frame(sim_data, top_bound, obst_bound, sim, time, variant='deltaU_to_deltaP'):
    """One time frame -> dict with ``cells_xy`` [n,2], ``top`` [nt,2], ``obst`` [no,2] (float64) and every named column of
    the variant as float64 [n] (what ``Evaluation.timeStep`` slices out, SMC:382-402 / GRAD:429-436)."""
    data = np.asarray(sim_data[sim:sim + 1, time:time + 1, ...])            # UTL:67-69
    top = np.asarray(top_bound[sim:sim + 1, time:time + 1, ...])
    obst = np.asarray(obst_bound[sim:sim + 1, time:time + 1, ...])
    n = first_pad(data[0, 0, :, 0])                                         # SMC:100
    nt = first_pad(top[0, 0, :, 0])                                         # SMC:118
    no = first_pad(obst[0, 0, :, 0])                                        # SMC:125
    cols = COLUMNS[variant]
    if data.shape[-1] <= max(cols.values()):
        raise ValueError('%s needs %d columns, the dataset has %d' % (variant, max(cols.values()) + 1, data.shape[-1]))
    out = {k: np.ascontiguousarray(data[0, 0, :n, c], dtype=np.float64) for k, c in cols.items() if k not in ('x', 'y')}
    out['cells_xy'] = np.ascontiguousarray(data[0, 0, :n, cols['x']:cols['y'] + 1], dtype=np.float64)
    out['top'] = np.ascontiguousarray(top[0, 0, :nt, :], dtype=np.float64)
    out['obst'] = np.ascontiguousarray(obst[0, 0, :no, :], dtype=np.float64)
    out['n_cells'] = n
    return out


def solver_rows(frame, variant='deltaU_to_deltaP', with_delta=True, p_prev=None):
    """The ``double[n][5]`` (or ``[n][7]``) rows ``psm_predict`` takes, from a frame: {Ux, Uy, Cx, Cy, p[, dUx, dUy]}.
    For deltaU_to_deltaP the pressure column is the PREVIOUS pressure: ``p - delta_p`` (SMC:639-641 infers it the same
    way) unless ``p_prev`` is given."""
    if p_prev is None:
        p_prev = frame['p'] - frame['dp'] if variant == 'deltaU_to_deltaP' else frame['p']
    cols = [frame['Ux'], frame['Uy'], frame['cells_xy'][:, 0], frame['cells_xy'][:, 1], p_prev]
    if variant == 'deltaU_to_deltaP' and with_delta:
        cols += [frame['dUx'], frame['dUy']]
    return np.ascontiguousarray(np.stack(cols, axis=1), dtype=np.float64)


def deltaU_change(frame):
    """SMC:396-398: where delta_U changed in the last time step, scaled to [0, 1] (input of the delta-U change weighting)."""
    d = np.abs(np.stack([frame['dUx'] - frame['dUx_prev'], frame['dUy'] - frame['dUy_prev']], axis=-1)).sum(axis=-1)
    return d / d.max()


def write_padded(arrays, n_rows):
    """The writer side (``padding``, data_generation.py:7-12) for tests and converters: [n, C] -> float32 [1, 1, n_rows, C]."""
    a = np.asarray(arrays, dtype=np.float32)
    out = np.full((1, 1, n_rows, a.shape[1]), PAD, dtype=np.float32)
    out[0, 0, :a.shape[0]] = a
    return out


def open_hdf5(path):
    """(sim_data, top_bound, obst_bound) datasets of an HDF5 file written by data_generation.py; needs h5py."""
    try:
        import h5py
    except ImportError as e:                                  # not in this image; the arrays can come from anywhere else
        raise ImportError('h5py is needed to open %s; read_frame() accepts any array-like instead' % path) from e
    f = h5py.File(path, 'r')
    return f['sim_data'], f['top_bound'], f['obst_bound']
