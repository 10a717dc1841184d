#!/usr/bin/env python
"""Generate golden vectors by running the REFERENCE ITSELF (build container only).

The reference has no tests or golden vectors and cannot be
imported as-is here: tensorflow, h5py, shapely, matplotlib, dask and mpi4py are
neither installed nor in the offline wheelhouse.  This script installs minimal
stand-ins for exactly those third-party names in ``sys.modules``, imports the
reference modules UNMODIFIED from /root/reference, and records what the reference's
own code returns on seeded synthetic inputs:

  * ``pressureSM_deltas.SM_call.Evaluation`` -- ``computeOnlyOnce`` (SMC:89-180),
    ``timeStep`` up to and including ``assemble_prediction`` (SMC:367-575, 182-365)
  * ``Eval_dual_Dense_onlycil.Evaluation``   -- ``computeOnlyOnce`` (GRAD:160-253),
    ``timeStep`` up to the four ``assemble_prediction`` calls (GRAD:418-547, 255-369)
  * ``python_module.init_func``              -- PMP:172-247 (tables both ways, mask, raster)

Stand-ins (the same substitutions the oracle documents):
  tensorflow.keras Dense stack -> float32 NumPy MLP read from an .npz
  h5py dataset                 -> ``read_dataset`` patched to return in-memory arrays
  shapely MultiPoint.convex_hull -> scipy.spatial.ConvexHull ring
  matplotlib Path.contains_points -> convex point-in-polygon
  matplotlib.pyplot, dask      -> inert mocks
  mpi4py                       -> single-rank communicator (gather -> [x], scatter -> x[0])
  PCA pickles                  -> scikit-learn IncrementalPCA objects with the attributes set

Everything arithmetic -- Qhull tables, barycentric weights, mask/raster loops, block
extraction, PCA transform, standardisation, the assembly loop -- is the reference's own
code.  Outputs go to ``tests/golden/*.npz``; nothing at test time reads /root/reference.

Usage:  python tests/golden/make_golden.py          (writes next to this file)
"""
import hashlib
import os
import pickle
import sys
import tempfile
import types
from unittest import mock

import numpy as np

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
PKG = os.path.join(REPO, 'solving-poisson-s-equation-through-dl-for-cfd-apllications_b200')
REF = '/root/reference'
sys.path.insert(0, PKG)
sys.path.insert(0, REPO)

from psm_b200 import synthetic as syn          # noqa: E402  (numpy-only input generator)
from oracle.domain import convex_hull_points, contains_points_convex   # noqa: E402  (stand-in geometry)


# --------------------------------------------------------------------------- stand-ins
class _NumpyMLP:
    """float32 Dense stack: relu on all but the last layer (NNS:24-33 / PMP:121-134)."""

    def __init__(self, weights=None, biases=None):
        self.weights, self.biases = weights, biases

    def load(self, path):
        z = np.load(path)
        n = len([k for k in z.files if k.startswith('W')])
        self.weights = [z['W%d' % i].astype(np.float32) for i in range(n)]
        self.biases = [z['b%d' % i].astype(np.float32) for i in range(n)]
        return self

    def load_weights(self, path):
        self.load(path)

    def summary(self):
        return 'numpy stand-in for the Keras Dense stack'

    def __call__(self, x):
        h = np.asarray(x, dtype=np.float32)
        for li, (w, b) in enumerate(zip(self.weights, self.biases)):
            h = h @ w + b
            if li < len(self.weights) - 1:
                h = np.maximum(h, np.float32(0))
        return h


class _HullRing:
    def __init__(self, pts):
        ring = convex_hull_points(pts)
        ring = np.concatenate([ring, ring[:1]])        # shapely rings are closed
        self.exterior = types.SimpleNamespace(coords=types.SimpleNamespace(xy=(ring[:, 0], ring[:, 1])))


class _MultiPoint:
    def __init__(self, pts):
        self.convex_hull = _HullRing(np.asarray(pts, dtype=np.float64))


class _Path:
    def __init__(self, pts):
        pts = np.asarray(pts, dtype=np.float64)
        if np.allclose(pts[0], pts[-1]):
            pts = pts[:-1]
        self.pts = pts

    def contains_points(self, xy):
        return contains_points_convex(self.pts, xy)


class _Comm:
    def Get_rank(self):
        return 0

    def Get_size(self):
        return 1

    def gather(self, x, root=0):
        return [x]

    def scatter(self, x, root=0):
        return x[0]


def install_stand_ins():
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    # tensorflow / keras
    keras_models = mod('tensorflow.keras.models', load_model=lambda path, **kw: _NumpyMLP().load(path))

    class _Sym:                                            # symbolic tensor of the functional API
        def __init__(self, chain):
            self.chain = chain

    class _Dense:
        def __init__(self, units, activation=None, **kw):
            self.units, self.activation = units, activation

        def __call__(self, s):
            return _Sym(s.chain + [self])

    def _Input(shape, **kw):
        return _Sym([])

    def _Model(inputs, outputs, name=None):
        return _NumpyMLP()

    anything = mock.MagicMock()
    keras_layers = mod('tensorflow.keras.layers', Dense=_Dense, Input=_Input)
    keras_layers.__getattr__ = lambda name: anything        # unused layer names imported by PMP
    keras = mod('tensorflow.keras', models=keras_models, layers=keras_layers, Model=_Model, Input=_Input,
                regularizers=anything)
    for sub in ('utils', 'callbacks', 'optimizers'):
        m = mod('tensorflow.keras.' + sub)
        m.__getattr__ = lambda name: anything
        setattr(keras, sub, m)
    tf_config = types.SimpleNamespace(list_physical_devices=lambda kind: [],
                                      experimental=types.SimpleNamespace(set_memory_growth=lambda *a: None))
    mod('tensorflow', keras=keras, config=tf_config)
    # inert mocks
    for name in ('h5py', 'dask', 'dask.array', 'matplotlib', 'matplotlib.pyplot'):
        sys.modules[name] = mock.MagicMock()
    mod('matplotlib.path', Path=_Path)
    sys.modules['matplotlib'].path = sys.modules['matplotlib.path']
    mod('shapely')
    mod('shapely.geometry', MultiPoint=_MultiPoint)
    mpi = mod('mpi4py', rc=types.SimpleNamespace(initialize=True, finalize=True))
    mpi.MPI = types.SimpleNamespace(COMM_WORLD=_Comm())
    mod('mpi4py.MPI', COMM_WORLD=mpi.MPI.COMM_WORLD)


class _Captured(Exception):
    pass


def make_pca(components, mean, evr):
    from sklearn.decomposition import IncrementalPCA
    pca = IncrementalPCA(n_components=components.shape[0], whiten=False)
    pca.components_ = components
    pca.mean_ = mean
    pca.explained_variance_ratio_ = evr
    pca.explained_variance_ = evr.copy()
    pca.var_ = np.ones_like(mean)
    pca.n_components_ = components.shape[0]
    pca.n_features_in_ = components.shape[1]
    pca.n_samples_seen_ = 1000
    return pca


def evr_for(n_total, n_keep, var):
    """explained_variance_ratio_ whose cumsum first exceeds ``var`` at index n_keep
    (SMC:86-87 takes argmax(cumsum > var))."""
    evr = np.full(n_total, (1.0 - var) / (2.0 * n_total))
    evr[:n_keep] = (var - 1e-6) / n_keep
    evr[n_keep] += 2e-6
    return evr


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def padded(a, n_pad=4):
    """[1,1,N+n_pad,C] float64 frame padded with -100.0 (data_generation.py layout)."""
    out = np.full((1, 1, a.shape[0] + n_pad, a.shape[1]), -100.0)
    out[0, 0, :a.shape[0]] = a
    return out


def write_param_files(tmp, P, extra, var, model_name):
    """maxs, PCA pickles (with ``extra`` unused trailing components), scaler, model weights."""
    rng = np.random.default_rng(99)
    np.savetxt(os.path.join(tmp, 'maxs'), P['maxs'])
    for tag, comp, mean in (('input', P['pca_in_components'], P['pca_in_mean']),
                            ('p', P['pca_out_components'], P['pca_out_mean'])):
        n_keep, K = comp.shape
        more = (rng.standard_normal((extra, K)) / np.sqrt(K)).astype(comp.dtype)
        pca = make_pca(np.concatenate([comp, more]), mean, evr_for(n_keep + extra, n_keep, var))
        for fname in ('ipca_%s.pkl' % tag, 'ipca_%s_more.pkl' % tag):
            with open(os.path.join(tmp, fname), 'wb') as f:
                pickle.dump(pca, f)
    if P['standardization'] == 'std':
        np.savez(os.path.join(tmp, 'mean_std.npz'), mean_in=P['mean_in'], std_in=P['std_in'],
                 mean_out=P['mean_out'], std_out=P['std_out'])
    else:
        np.savetxt(os.path.join(tmp, 'maxs_PCA'), [P['max_abs_input_PCA'], P['max_abs_output_PCA']])
    kw = {}
    for i, (w, b) in enumerate(zip(P['mlp_weights'], P['mlp_biases'])):
        kw['W%d' % i], kw['b%d' % i] = w, b
    np.savez(os.path.join(tmp, model_name), **kw)
    os.rename(os.path.join(tmp, model_name + '.npz'), os.path.join(tmp, model_name))


# --------------------------------------------------------------------------- cases
def golden_smc(name, mesh_kw, seed, pc_in=24, pc_p=20):
    """Run the reference Improved_SM evaluation (SMC) and record its outputs."""
    import pressureSM_deltas.SM_call as SMC
    from pressureSM_deltas import utils as UTL
    mesh = syn.make_mesh(seed=seed, **mesh_kw)
    F = syn.make_fields(mesh, seed=seed)
    P = syn.make_params(seed=seed, pc_in=pc_in, pc_p=pc_p, standardization='std')
    n = mesh['cells'].shape[0]
    rng = np.random.default_rng(seed + 5)
    dp_label = 0.01 * rng.standard_normal(n)
    cols = np.stack([F['Ux'], F['Uy'], F['p_prev'], mesh['cells'][:, 0], mesh['cells'][:, 1], F['dUx'], F['dUy'],
                     dp_label, 0.5 * F['dUx'], 0.5 * F['dUy'], dp_label], axis=1)
    frame = (padded(cols), padded(mesh['top']), padded(mesh['obst']))
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        write_param_files(tmp, P, extra=8, var=0.95, model_name='model.h5')
        os.chdir(tmp)
        try:
            UTL.read_dataset = lambda path, sim, time: tuple(a.copy() for a in frame)
            UTL.plot_random_blocks = lambda *a, **k: None
            ev = SMC.Evaluation(5e-3, 128, 32, 0.95, 0.95, 'unused.hdf5', 'model.h5', 128, 'std')
            assert (ev.pc_in, ev.pc_p) == (pc_in, pc_p), (ev.pc_in, ev.pc_p)
            ev.pred_minus_true_block, ev.pred_minus_true_squared_block = [], []
            ev.computeOnlyOnce(0)
            cap = {}
            orig = ev.assemble_prediction

            def wrapped(array, indices_list, n_x, n_y, *rest):
                cap['blocks'] = np.array(array, copy=True)
                cap['indices_list'] = np.array(indices_list)
                cap['n_x'], cap['n_y'] = n_x, n_y
                a2 = np.array(array, copy=True)          # the reference corrects `array` in place
                cap['field'], _ = orig(array, indices_list, n_x, n_y, *rest)
                cap['blocks_corrected'] = np.array(array, copy=True)
                # same call with apply_filter=True (SMC:353-356: scipy.ndimage.gaussian_filter, sigma (10, 10))
                cap['field_filtered'], _ = orig(a2, indices_list, n_x, n_y, True, *rest[1:])
                raise _Captured()

            ev.assemble_prediction = wrapped
            try:
                ev.timeStep(0, 0, False, False, False, False)
            except _Captured:
                pass
        finally:
            os.chdir(cwd)
    out = dict(
        mesh_kw=np.array(repr(mesh_kw)), seed=seed, pc_in=pc_in, pc_p=pc_p,
        grid_shape=np.array([ev.grid_shape_y, ev.grid_shape_x]),
        vert_sha=np.array(sha(ev.vert.astype(np.int32))), vert_sub=ev.vert[::17].astype(np.int32),
        weights_sub=ev.weights[::17], indices_sha=np.array(sha(ev.indices.astype(np.int64))),
        indices_sub=ev.indices[::17].astype(np.int32),
        sdfunct=ev.sdfunct[:, :, 0].astype(np.float32), sdfunct_nonzero_sha=np.array(sha(ev.sdfunct != 0)),
        x_array_sub=ev.x_array[:, ::8, ::8, :], blocks_sub=cap['blocks'][:, ::4, ::4],
        offsets=(cap['blocks'] - cap['blocks_corrected']).reshape(cap['blocks'].shape[0], -1)[:, 0],
        indices_list=cap['indices_list'], n_x=cap['n_x'], n_y=cap['n_y'], field=cap['field'],
        field_filtered_sub=cap['field_filtered'][::2, ::2])
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **out)
    print(name, 'grid', out['grid_shape'], 'blocks', cap['blocks'].shape, 'nan in field', int(np.isnan(cap['field']).sum()))


def golden_grad(name, mesh_kw, seed, pc_in=24, pc_p=20):
    """Run the reference U_to_gradP evaluation (GRAD) and record its outputs."""
    sys.path.insert(0, os.path.join(REF, 'Improved_SM/U_to_gradP/evaluation'))
    import Eval_dual_Dense_onlycil as GRAD
    mesh = syn.make_mesh(seed=seed, **mesh_kw)
    F = syn.make_fields(mesh, seed=seed)
    P = syn.make_params(seed=seed, pc_in=pc_in, pc_p=pc_p, standardization='max_abs', n_out_channels=2,
                        maxs=(1.0, 0.536, 0.999, 0.8, 0.7))
    n = mesh['cells'].shape[0]
    rng = np.random.default_rng(seed + 5)
    lab = 0.01 * rng.standard_normal((n, 3))
    cols = np.stack([F['Ux'], F['Uy'], F['p_prev'], mesh['cells'][:, 0], mesh['cells'][:, 1], lab[:, 0],
                     lab[:, 1], lab[:, 2]], axis=1)
    frame = (padded(cols), padded(mesh['top']), padded(mesh['obst']))
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        write_param_files(tmp, P, extra=8, var=0.95, model_name='model_1.h5')
        os.chdir(tmp)
        try:
            ev = GRAD.Evaluation(5e-3, 128, 96, 0.95, 0.95, 'unused.hdf5', 'model_1.h5', 512)
            ev.read_dataset = lambda path, sim, time: tuple(a.copy() for a in frame)
            assert (ev.pc_in, ev.pc_p) == (pc_in, pc_p), (ev.pc_in, ev.pc_p)
            ev.computeOnlyOnce(0)
            cap = {'calls': []}
            orig = ev.assemble_prediction

            def wrapped(field, array, indices_list, n_x, n_y, *rest):
                before = np.array(array, copy=True)
                res = orig(field, array, indices_list, n_x, n_y, *rest)
                cap['calls'].append((field, before, np.array(res[0, :, :, 0], copy=True)))
                cap['indices_list'], cap['n_x'], cap['n_y'] = np.array(indices_list), n_x, n_y
                if len(cap['calls']) == 2:
                    raise _Captured()
                return res

            ev.assemble_prediction = wrapped
            try:
                ev.timeStep(0, 0, False, False, False, False)
            except _Captured:
                pass
        finally:
            os.chdir(cwd)
    (f0, b0, r0), (f1, b1, r1) = cap['calls']
    assert (f0, f1) == ('dp_dx', 'dp_dy')
    out = dict(
        mesh_kw=np.array(repr(mesh_kw)), seed=seed, pc_in=pc_in, pc_p=pc_p,
        grid_shape=np.array([ev.grid_shape_y, ev.grid_shape_x]),
        vert_sha=np.array(sha(ev.vert.astype(np.int32))), indices_sha=np.array(sha(ev.indices.astype(np.int64))),
        sdfunct=ev.sdfunct[:, :, 0].astype(np.float32), x_array_sub=ev.x_array[:, ::8, ::8, :],
        blocks_dx_sub=b0[:, ::4, ::4], blocks_dy_sub=b1[:, ::4, ::4], indices_list=cap['indices_list'],
        n_x=cap['n_x'], n_y=cap['n_y'], dp_dx=r0, dp_dy=r1)
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **out)
    print(name, 'grid', out['grid_shape'], 'blocks', b0.shape, 'nan', int(np.isnan(r0).sum()), int(np.isnan(r1).sum()))



def golden_grad_integrate(name, mesh_kw, seed, pc_in=24, pc_p=20):
    """Run the reference U_to_gradP ``timeStep`` to its END: the two assembled gradient fields, then the pressure recovery
    (``integrate_field`` GRAD:371-416 on four quadrants + the stitch GRAD:585-628).  The recovered field is a local of
    ``timeStep``; it is captured where the reference hands it to ``np.ma.array`` for plotting (GRAD:631).  Needs a grid whose
    hard-coded centre row 200 (GRAD:592) crosses the obstacle.  Also records ``integrate_field`` called directly on a seeded
    block (all four direction combinations)."""
    sys.path.insert(0, os.path.join(REF, 'Improved_SM/U_to_gradP/evaluation'))
    import Eval_dual_Dense_onlycil as GRAD
    mesh = syn.make_mesh(seed=seed, **mesh_kw)
    F = syn.make_fields(mesh, seed=seed)
    P = syn.make_params(seed=seed, pc_in=pc_in, pc_p=pc_p, standardization='max_abs', n_out_channels=2,
                        maxs=(1.0, 0.536, 0.999, 0.8, 0.7))
    n = mesh['cells'].shape[0]
    rng = np.random.default_rng(seed + 5)
    lab = 0.01 * rng.standard_normal((n, 3))
    cols = np.stack([F['Ux'], F['Uy'], F['p_prev'], mesh['cells'][:, 0], mesh['cells'][:, 1], lab[:, 0],
                     lab[:, 1], lab[:, 2]], axis=1)
    frame = (padded(cols), padded(mesh['top']), padded(mesh['obst']))
    cwd = os.getcwd()
    cap = {'fields': [], 'ma': []}
    with tempfile.TemporaryDirectory() as tmp:
        write_param_files(tmp, P, extra=8, var=0.95, model_name='model_1.h5')
        os.chdir(tmp)
        try:
            ev = GRAD.Evaluation(5e-3, 128, 96, 0.95, 0.95, 'unused.hdf5', 'model_1.h5', 512)
            ev.read_dataset = lambda path, sim, time: tuple(a.copy() for a in frame)
            ev.computeOnlyOnce(0)
            orig = ev.assemble_prediction

            def wrapped(field, array, indices_list, n_x, n_y, *rest):
                res = orig(field, array, indices_list, n_x, n_y, *rest)
                cap['fields'].append((field, np.array(res[0, :, :, 0], copy=True)))
                return res

            ev.assemble_prediction = wrapped
            GRAD.plt.subplots = lambda *a, **k: (mock.MagicMock(), mock.MagicMock())
            ma_orig = np.ma.array

            def ma_hook(data, *a, **k):
                cap['ma'].append(np.array(data, copy=True))
                return ma_orig(data, *a, **k)

            np.ma.array = ma_hook
            try:
                ev.timeStep(0, 0, False, False, False, False)
            finally:
                np.ma.array = ma_orig
            # integrate_field on its own: seeded block narrower and lower than the grid, every direction combination
            r2 = np.random.default_rng(seed + 77)
            blk = r2.standard_normal((150, 120, 2))
            xl = np.linspace(ev.min_x, ev.max_x, ev.grid_shape_x)
            yl = np.linspace(ev.min_y, ev.max_y, ev.grid_shape_y)
            kat = [ev.integrate_field(blk.copy(), xl, yl, direction_x=dx_, direction_y=dy_) for dx_ in (1, -1) for dy_ in (1, -1)]
        finally:
            os.chdir(cwd)
    (f0, r0), (f1, r1) = cap['fields'][0], cap['fields'][1]
    assert (f0, f1) == ('dp_dx', 'dp_dy')
    p_field = cap['ma'][0]                                     # GRAD:631: np.ma.array(field, mask=...)
    out = dict(mesh_kw=np.array(repr(mesh_kw)), seed=seed, pc_in=pc_in, pc_p=pc_p,
               grid_shape=np.array([ev.grid_shape_y, ev.grid_shape_x]), sdfunct=ev.sdfunct[:, :, 0].astype(np.float64),
               bbox=np.array([ev.min_x, ev.max_x, ev.min_y, ev.max_y]), x0_min=np.array(ev.X0.min()),
               dp_dx=r0, dp_dy=r1, p_field=p_field, kat_block_seed=seed + 77, kat=np.stack(kat))
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **out)
    print(name, 'grid', out['grid_shape'], 'p_field', p_field.shape, 'nan', int(np.isnan(p_field).sum()))


class sentinel_empty:
    """PMP:225 allocates ``indices`` with ``np.empty`` and its raster loop (PMP:233-243) writes only the valid rows, so the
    other rows are uninitialised memory.  While ``init_func`` runs, ``np.empty`` is wrapped to pre-fill integer-free float
    buffers with a sentinel: the rows the loop writes can then be told apart and ONLY those are hashed."""
    VALUE = -7.0

    def __enter__(self):
        self._orig = np.empty

        def filled(shape, *a, **k):
            out = self._orig(shape, *a, **k)
            if out.dtype.kind == 'f':
                out.fill(self.VALUE)
            return out
        np.empty = filled
        return self

    def __exit__(self, *exc):
        np.empty = self._orig


def written_rows(indices):
    return ~np.all(indices == sentinel_empty.VALUE, axis=1)


def golden_pmp_init(name, mesh_kw, seed):
    """Run the reference solver-side ``init_func`` (PMP:172-247): both table directions,
    ``domain_dist`` ([::10] sub-sampling, 2-decimal bbox) and the raster loop."""
    mesh = syn.make_mesh(seed=seed, **mesh_kw)
    F = syn.make_fields(mesh, seed=seed)
    P = syn.make_params(seed=seed, pc_in=24, pc_p=20, standardization='max_abs')
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        write_param_files(tmp, P, extra=8, var=0.95, model_name='weights.h5')
        os.chdir(tmp)
        try:
            sys.path.insert(0, os.path.join(REF, 'Thesis_Work/Chapter5/parallelized/test_case'))
            import python_module as PMP
            arr = syn.pack_cells(mesh, F, with_delta=False)
            with sentinel_empty():
                PMP.init_func(arr, mesh['top'], mesh['obst'], 0)
        finally:
            os.chdir(cwd)
    wr = written_rows(PMP.indices)
    out = dict(
        mesh_kw=np.array(repr(mesh_kw)), seed=seed,
        grid_shape=np.array([PMP.grid_shape_y, PMP.grid_shape_x]),
        vert_fwd_sha=np.array(sha(PMP.vert_OFtoNP.astype(np.int32))), weights_fwd_sub=PMP.weights_OFtoNP[::17],
        vert_back_sha=np.array(sha(PMP.vert_NPtoOF.astype(np.int32))), vert_back_sub=PMP.vert_NPtoOF[::5].astype(np.int32),
        weights_back_sub=PMP.weights_NPtoOF[::5],
        written_sha=np.array(sha(wr.astype(np.uint8))), indices_sha=np.array(sha(PMP.indices[wr].astype(np.int64))),
        sdfunct=PMP.sdfunct[:, :, 0].astype(np.float32))
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **out)
    print(name, 'grid', out['grid_shape'])


def golden_pmp_step(name, mesh_kw, seed, pc_in=24, pc_p=20):
    """Run the reference solver-side module end to end: ``init_func`` then ``py_func`` (PMP:249-517) -- the call the
    PISO loop makes (FOAM/PythonComm.H:24-27).  PC_input is cut at 0.995 and PC_p at 0.95 (PMP:112-113)."""
    mesh = syn.make_mesh(seed=seed, **mesh_kw)
    F = syn.make_fields(mesh, seed=seed)
    P = syn.make_params(seed=seed, pc_in=pc_in, pc_p=pc_p, standardization='max_abs')
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        write_param_files(tmp, P, extra=8, var=0.95, model_name='weights.h5')
        rng = np.random.default_rng(99)
        comp, mean = P['pca_in_components'], P['pca_in_mean']
        more = (rng.standard_normal((8, comp.shape[1])) / np.sqrt(comp.shape[1])).astype(comp.dtype)
        with open(os.path.join(tmp, 'ipca_input_more.pkl'), 'wb') as f:
            pickle.dump(make_pca(np.concatenate([comp, more]), mean, evr_for(pc_in + 8, pc_in, 0.995)), f)
        os.chdir(tmp)
        try:
            sys.path.insert(0, os.path.join(REF, 'Thesis_Work/Chapter5/parallelized/test_case'))
            sys.modules.pop('python_module', None)             # the module reads its artefacts at import time
            import python_module as PMP
            assert (PMP.PC_input, PMP.PC_p) == (pc_in, pc_p), (PMP.PC_input, PMP.PC_p)
            PMP.memory = lambda: ''
            arr = syn.pack_cells(mesh, F, with_delta=False)
            with sentinel_empty():
                PMP.init_func(arr, mesh['top'], mesh['obst'], 0)
            wr = written_rows(PMP.indices)
            PMP.indices[~wr] = 0           # what uninitialised memory holds 'in practice' (SURVEY.md 9.3): py_func reads these rows
            p = np.array(PMP.py_func(arr, 0), dtype=np.float64)
        finally:
            os.chdir(cwd)
    out = dict(mesh_kw=np.array(repr(mesh_kw)), seed=seed, pc_in=pc_in, pc_p=pc_p,
               grid_shape=np.array([PMP.grid_shape_y, PMP.grid_shape_x]), p=p,
               written_sha=np.array(sha(wr.astype(np.uint8))), indices_sha=np.array(sha(PMP.indices[wr].astype(np.int64))))
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **out)
    print(name, 'grid', out['grid_shape'], 'cells', p.shape, 'kept p_prev at', int((p == arr[:, 4]).sum()), 'cells')


GOLDEN_CASES = {
    # name: (kind, mesh kwargs, seed)
    'smc_small': ('smc', dict(H=240, W=330, nx=130, ny=90, R=0.1), 11),
    'smc_bigobst': ('smc', dict(H=240, W=330, nx=130, ny=90, R=0.35, center=(0.3575, -0.0375)), 12),      # empty strips -> NaN chains
    'grad_small': ('grad', dict(H=240, W=340, nx=130, ny=90, R=0.1), 13),
    'grad_integrate': ('grad_integrate', dict(H=400, W=340, nx=130, ny=150, R=0.1), 16),   # row 200 (GRAD:592) crosses the cylinder
    'pmp_init_small': ('pmp_init', dict(H=240, W=340, nx=130, ny=90, R=0.1), 14),
    'pmp_step_small': ('pmp_step', dict(H=260, W=380, nx=150, ny=100, R=0.1), 15),
}


def main():
    only = set(sys.argv[1:])                                   # optional: names of the cases to (re)generate
    install_stand_ins()
    sys.path.insert(0, os.path.join(REF, 'Improved_SM/deltaU_to_deltaP/source'))
    for name, (kind, mesh_kw, seed) in GOLDEN_CASES.items():
        if only and name not in only:
            continue
        if kind == 'smc':
            golden_smc(name, mesh_kw, seed)
        elif kind == 'grad':
            golden_grad(name, mesh_kw, seed)
        elif kind == 'grad_integrate':
            golden_grad_integrate(name, mesh_kw, seed)
        elif kind == 'pmp_step':
            golden_pmp_step(name, mesh_kw, seed)
        else:
            golden_pmp_init(name, mesh_kw, seed)


if __name__ == '__main__':
    main()
