"""Python face of the C-ABI handle: ``PressureSurrogate``.

Thin by design -- argument marshalling only.  All per-step arithmetic happens in the CUDA
library behind ``psm_predict``; see ``python_module.py`` for the reference-named
``init_func`` / ``py_func`` surface built on top of this class.
"""
import ctypes as C

import numpy as np

from . import _capi as capi
from . import tables as _tables

_VARIANT_CODE = {'deltaU_to_deltaP': capi.PSM_DELTAU_TO_DELTAP, 'U_to_gradP': capi.PSM_U_TO_GRADP,
                 'thesis': capi.PSM_THESIS_U_TO_P}     # 'thesis': the solver module the reference ships (PMP), U -> p


def _ptr(a, ctype):
    return a.ctypes.data_as(C.POINTER(ctype))


def compile_plan(variant, H, W, mask, shape=128, overlap=32):
    """Host-only plan compiler (no GPU): returns dict(origins, indices_list, owner, rec, tasks)."""
    lib = capi.load()
    mask = np.ascontiguousarray(mask, dtype=np.uint8)
    assert mask.shape == (H, W)
    nb, nf, nt = C.c_int32(), C.c_int32(), C.c_int32()
    v = _VARIANT_CODE[variant]
    rc = lib.psm_plan_sizes(v, H, W, shape, overlap, _ptr(mask, C.c_uint8), C.byref(nb), C.byref(nf), C.byref(nt))
    if rc < 0:
        raise capi.PsmError(rc, 'plan rejected (geometry on which the reference is undefined)')
    B, F, T = nb.value, nf.value, nt.value
    origins = np.zeros((B, 2), np.int32)
    il = np.zeros((B, 2), np.int32)
    owner = np.zeros((H, W), np.int32)
    rec = np.zeros((F, B, 4), np.int32)
    tasks = np.zeros((T, 8), np.int32)
    rc = lib.psm_plan_compile(v, H, W, shape, overlap, _ptr(mask, C.c_uint8), _ptr(origins, C.c_int32),
                              _ptr(il, C.c_int32), _ptr(owner, C.c_int32), _ptr(rec, C.c_int32), _ptr(tasks, C.c_int32))
    if rc < 0:
        raise capi.PsmError(rc, 'plan rejected')
    nl = C.c_int32()
    lib.psm_plan_shift_lines(v, H, W, shape, overlap, _ptr(mask, C.c_uint8), C.byref(nl), None)
    lines = np.zeros((nl.value, 8), np.int32)
    lib.psm_plan_shift_lines(v, H, W, shape, overlap, _ptr(mask, C.c_uint8), C.byref(nl), _ptr(lines, C.c_int32))
    return dict(origins=origins, indices_list=il, owner=owner, rec=rec, tasks=tasks, lines=lines, n_blocks=B, n_fields=F)


def mesh_hash(variant, delta, cells_xy, top, obst, probe=None):
    """Key of the table cache (``psm_mesh_hash``): 16 hex digits over variant, spacing, cell centres, boundary points, probe."""
    cells_xy = np.ascontiguousarray(cells_xy, dtype=np.float64)
    top = np.ascontiguousarray(top, dtype=np.float64)
    obst = np.ascontiguousarray(obst, dtype=np.float64)
    pr = None if probe is None else np.ascontiguousarray(probe, dtype=np.float64)
    out = C.create_string_buffer(17)
    rc = capi.load().psm_mesh_hash(_VARIANT_CODE[variant], float(delta), _ptr(cells_xy, C.c_double), 2, cells_xy.shape[0],
                                   _ptr(top, C.c_double), top.shape[0], _ptr(obst, C.c_double), obst.shape[0],
                                   None if pr is None else _ptr(pr, C.c_double), out)
    if rc < 0:
        raise capi.PsmError(rc, 'psm_mesh_hash failed')
    return out.value.decode()


def mesh_grid(variant, delta, cells_xy):
    """Rounded bounding box and grid shape exactly as ``psm_init_mesh`` derives them: ((x_min, x_max, y_min, y_max), H, W)."""
    cells_xy = np.ascontiguousarray(cells_xy, dtype=np.float64)
    bbox = (C.c_double * 4)()
    H, W = C.c_int32(), C.c_int32()
    rc = capi.load().psm_mesh_grid(_VARIANT_CODE[variant], float(delta), _ptr(cells_xy, C.c_double), 2, cells_xy.shape[0], bbox,
                                   C.byref(H), C.byref(W))
    if rc < 0:
        raise capi.PsmError(rc, 'psm_mesh_grid failed')
    return tuple(bbox), H.value, W.value


def back_tables_closed_form(cells_xy, X0_row, Y0_col):
    """``psm_back_tables_closed_form`` (the C++ port of ``tables.regular_grid_back_tables``)."""
    cells_xy = np.ascontiguousarray(cells_xy, dtype=np.float64)
    X0_row = np.ascontiguousarray(X0_row, dtype=np.float64)
    Y0_col = np.ascontiguousarray(Y0_col, dtype=np.float64)
    n = cells_xy.shape[0]
    vb, wb = np.empty((n, 3), np.int32), np.empty((n, 3), np.float64)
    rc = capi.load().psm_back_tables_closed_form(_ptr(cells_xy, C.c_double), 2, n, _ptr(X0_row, C.c_double), X0_row.size,
                                                 _ptr(Y0_col, C.c_double), Y0_col.size, _ptr(vb, C.c_int32), _ptr(wb, C.c_double))
    if rc < 0:
        raise capi.PsmError(rc, 'psm_back_tables_closed_form failed')
    return vb, wb


def register_host_buffer(a):
    """Page-lock a NumPy buffer that lives for the whole run (FOAM/PythonComm_init.H:53 allocates the solver's rows once),
    so that psm_predict copies it at full PCIe speed and the step stays one CUDA graph.  Returns True on success."""
    return capi.load().psm_register_host_buffer(C.c_void_p(a.ctypes.data), a.nbytes) == 0


def unregister_host_buffer(a):
    return capi.load().psm_unregister_host_buffer(C.c_void_p(a.ctypes.data)) == 0


def debug_gemm(A, B, mode=0, splits=1, device=0):
    """C[splits, M, N] = A[M, K] @ B[N, K].T on the GPU with one of the library's GEMM kernels (tests)."""
    lib = capi.load()
    A = np.ascontiguousarray(A, dtype=np.float32)
    B = np.ascontiguousarray(B, dtype=np.float32)
    M, K = A.shape
    N = B.shape[0]
    Cc = np.empty((splits, M, N), dtype=np.float32)
    rc = lib.psm_debug_gemm(device, mode, M, N, K, _ptr(A, C.c_float), _ptr(B, C.c_float), _ptr(Cc, C.c_float), splits)
    if rc < 0:
        raise capi.PsmError(rc, 'psm_debug_gemm failed')
    return Cc


def _marshal_params(p):
    """dict -> (PsmParams, arrays that must stay alive while the struct is in use)."""
    f64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)     # noqa: E731
    P = capi.PsmParams()
    maxs = list(np.asarray(p['maxs'], dtype=np.float64)) + [1.0] * 5
    P.maxs = (C.c_double * 5)(*maxs[:5])
    cin, cout = f64(p['pca_in_components']), f64(p['pca_out_components'])
    min_, mout = f64(p['pca_in_mean']), f64(p['pca_out_mean'])
    P.pc_in, P.pc_p = cin.shape[0], cout.shape[0]
    P.n_out_channels = int(p.get('n_out_channels', 1))
    keep = [cin, cout, min_, mout]
    P.pca_in_components, P.pca_in_mean = _ptr(cin, C.c_double), _ptr(min_, C.c_double)
    P.pca_out_components, P.pca_out_mean = _ptr(cout, C.c_double), _ptr(mout, C.c_double)
    method = p.get('standardization', 'std')
    if method == 'std':                                  # SMC:505-512 (mean_std.npz)
        P.standardization = capi.PSM_STD
        for name in ('mean_in', 'std_in', 'mean_out', 'std_out'):
            a = f64(p[name])
            keep.append(a)
            setattr(P, name, _ptr(a, C.c_double))
    elif method == 'min_max':                            # SMC:513-520 (min_max_values.npz): min_* in the mean_* slots, max_* in the std_* slots
        P.standardization = capi.PSM_MIN_MAX
        for name, key in (('mean_in', 'min_in'), ('std_in', 'max_in'), ('mean_out', 'min_out'), ('std_out', 'max_out')):
            a = f64(p[key])
            keep.append(a)
            setattr(P, name, _ptr(a, C.c_double))
    elif method == 'max_abs':                            # SMC:521-523
        P.standardization = capi.PSM_MAX_ABS
        P.max_abs_input_PCA = float(p['max_abs_input_PCA'])
        P.max_abs_output_PCA = float(p['max_abs_output_PCA'])
    else:
        raise ValueError("standardization must be 'std', 'min_max' or 'max_abs' (SMC:505-525), got %r" % (method,))
    ws = [np.ascontiguousarray(w, dtype=np.float32) for w in p['mlp_weights']]
    bs = [np.ascontiguousarray(b, dtype=np.float32) for b in p['mlp_biases']]
    dims = np.ascontiguousarray([ws[0].shape[0]] + [w.shape[1] for w in ws], dtype=np.int32)
    P.n_dense = len(ws)
    P.layer_dims = _ptr(dims, C.c_int32)
    wp = (capi.c_float_p * len(ws))(*[_ptr(w, C.c_float) for w in ws])
    bp = (capi.c_float_p * len(bs))(*[_ptr(b, C.c_float) for b in bs])
    P.dense_kernels, P.dense_biases = wp, bp
    keep += ws + bs + [dims, wp, bp]
    return P, keep


def _marshal_tables(t):
    """dict from ``psm_b200.tables.build_tables`` -> (PsmTables, arrays to keep alive)."""
    T = capi.PsmTables()
    T.n_cells, T.grid_h, T.grid_w = int(t['n_cells']), int(t['H']), int(t['W'])
    vert = np.ascontiguousarray(t['vert'], dtype=np.int32)
    wts = np.ascontiguousarray(t['weights'], dtype=np.float64)
    ind = np.ascontiguousarray(t['indices'], dtype=np.int64)
    sdf = np.ascontiguousarray(t['sdfunct'], dtype=np.float64).reshape(T.grid_h, T.grid_w)
    T.vert, T.weights = _ptr(vert, C.c_int32), _ptr(wts, C.c_double)
    T.indices, T.sdfunct = _ptr(ind, C.c_int64), _ptr(sdf, C.c_double)
    keep = [vert, wts, ind, sdf]
    if t.get('vert_back') is not None:
        vb = np.ascontiguousarray(t['vert_back'], dtype=np.int32)
        wb = np.ascontiguousarray(t['weights_back'], dtype=np.float64)
        T.vert_back, T.weights_back = _ptr(vb, C.c_int32), _ptr(wb, C.c_double)
        keep += [vb, wb]
    return T, keep


def save_params(p, path, shape=128):
    """Write the model artefacts to the flat binary container ``psm_load_params_file`` reads (host only, no GPU)."""
    P, keep = _marshal_params(p)
    rc = capi.load().psm_save_params(C.byref(P), int(shape), str(path).encode())
    if rc < 0:
        raise capi.PsmError(rc, 'psm_save_params failed for %s' % path)


def save_tables(t, path):
    """Write the once-per-mesh tables to the flat binary container ``psm_init_from_file`` reads (host only, no GPU)."""
    T, keep = _marshal_tables(t)
    rc = capi.load().psm_save_tables(C.byref(T), str(path).encode())
    if rc < 0:
        raise capi.PsmError(rc, 'psm_save_tables failed for %s' % path)


def debug_dense_stack(x, kernels, biases, mode=0, clusters=0, device=0):
    """The fused Dense-stack kernel on its own (tests): Keras-layout kernels [in][out], ReLU between layers,
    linear last layer.  Returns float32 [M, out]."""
    lib = capi.load()
    x = np.ascontiguousarray(x, dtype=np.float32)
    ks = [np.ascontiguousarray(k, dtype=np.float32) for k in kernels]
    bs = [np.ascontiguousarray(b, dtype=np.float32) for b in biases]
    dims = np.array([x.shape[1]] + [k.shape[1] for k in ks], dtype=np.int32)
    n = len(ks)
    kp = (capi.c_float_p * n)(*[_ptr(k, C.c_float) for k in ks])
    bp = (capi.c_float_p * n)(*[_ptr(b, C.c_float) for b in bs])
    out = np.empty((x.shape[0], int(dims[-1])), dtype=np.float32)
    rc = lib.psm_debug_dense_stack(device, mode, x.shape[0], n, _ptr(dims, C.c_int32), kp, bp, _ptr(x, C.c_float),
                                   _ptr(out, C.c_float), clusters)
    if rc < 0:
        raise capi.PsmError(rc, 'psm_debug_dense_stack failed')
    return out


class PressureSurrogate:
    """One mesh, one GPU, one stream.  Mirrors the lifetime of the reference's module globals
    (PMP:103-118 params, PMP:193 tables) behind an explicit handle."""

    def __init__(self, variant='deltaU_to_deltaP', device=0, delta=5e-3, shape=128, overlap=None, input_cols=None,
                 additive=True, ref_bc=0.0, skip_threshold=1e-4, near_wall_sdf=0.0, timings=False, gemm_mode=0, filter_sigma=0.0):
        if variant not in _VARIANT_CODE:
            raise ValueError('variant must be one of %s' % list(_VARIANT_CODE))
        self.lib = capi.load()
        self.variant = variant
        if overlap is None:       # EP:90 overlap_ratio 0.25 ; GRAD:708 avance ; PMP:304 avance = int(0.1 * 128)
            overlap = {'deltaU_to_deltaP': 32, 'U_to_gradP': 96, 'thesis': 12}[variant]
        if variant == 'thesis':
            additive = False                                             # py_func returns p itself (PMP:490)
        if input_cols is None:
            input_cols = 7 if variant == 'deltaU_to_deltaP' else 5
        self.input_cols = input_cols
        self.n_fields = 2 if variant == 'U_to_gradP' else 1
        self.delta = delta
        cfg = capi.PsmConfig(variant=_VARIANT_CODE[variant], device=device, delta=delta, shape=shape, overlap=overlap,
                             input_cols=input_cols, additive=int(additive), ref_bc=ref_bc,
                             skip_threshold=skip_threshold, near_wall_sdf=near_wall_sdf,
                             enable_timings=int(timings), gemm_mode=int(gemm_mode), filter_sigma=float(filter_sigma))
        self._h = C.c_void_p()
        rc = self.lib.psm_create(C.byref(self._h), C.byref(cfg))
        if rc < 0:
            raise capi.PsmError(rc, (self.lib.psm_last_error(None) or b'').decode())
        self.n_cells = None
        self._keep = []

    # ------------------------------------------------------------------ lifetime
    def close(self):
        if getattr(self, '_h', None) is not None and self._h:
            self.lib.psm_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc):
        return capi.check(self._h, rc)

    # ------------------------------------------------------------------ init
    def load_params(self, p):
        """``p``: dict as produced by ``psm_b200.synthetic.make_params`` / ``psm_b200.params.load_reference_dir``."""
        P, keep = _marshal_params(p)
        self._check(self.lib.psm_load_params(self._h, C.byref(P)))
        self.pc_in, self.pc_p = P.pc_in, P.pc_p
        return self

    def load_params_file(self, path, pc=None):
        """Artefacts from the flat file written by ``save_params`` (route B of INTEGRATION.md).  ``pc`` = (pc_in, pc_p)
        if the caller wants them mirrored on the Python object (the library reads them from the file)."""
        self._check(self.lib.psm_load_params_file(self._h, str(path).encode()))
        if pc is not None:
            self.pc_in, self.pc_p = pc
        return self

    def init_tables(self, t):
        """``t``: dict from ``psm_b200.tables.build_tables`` (PMP:172-247 equivalent)."""
        T, keep = _marshal_tables(t)
        self._check(self.lib.psm_init_with_tables(self._h, C.byref(T)))
        return self._after_init()

    def init_from_file(self, path):
        """Tables from the flat file written by ``save_tables`` / ``python -m psm_b200.tables_file``."""
        self._check(self.lib.psm_init_from_file(self._h, str(path).encode()))
        return self._after_init()

    def _after_init(self):
        g = self.geometry()
        self.n_cells, self.H, self.W = g['n_cells'], g['grid_h'], g['grid_w']
        self.n_blocks = g['n_blocks']
        return self

    # ------------------------------------------------------------------ multi-GPU
    @staticmethod
    def unique_id():
        """Rank 0: the NCCL unique id (bytes) to ship to the other ranks (MPI_Bcast, torch.distributed ...)."""
        buf = C.create_string_buffer(capi.UNIQUE_ID_BYTES)
        rc = capi.load().psm_comm_get_unique_id(buf)
        if rc < 0:
            raise capi.PsmError(rc, (capi.load().psm_last_error(None) or b'').decode())
        return buf.raw

    def comm_init(self, unique_id, rank, world):
        """Collective: attach this handle to the communicator of ``world`` ranks (before ``init_shard``)."""
        buf = C.create_string_buffer(bytes(unique_id), capi.UNIQUE_ID_BYTES)
        self._check(self.lib.psm_comm_init(self._h, buf, int(rank), int(world)))
        return self

    def init_shard(self, sh):
        """``sh``: one dict of ``psm_b200.shard.partition`` -- this rank's block rows (PMP:179-185 replaced)."""
        T = capi.PsmShard()
        T.rank, T.world, T.grid_h, T.grid_w = int(sh['rank']), int(sh['world']), int(sh['H']), int(sh['W'])
        T.row0, T.row1, T.ext_rows, T.send_rows = int(sh['row0']), int(sh['row1']), int(sh['ext_rows']), int(sh['send_rows'])
        T.blk_row0, T.blk_row1 = int(sh['blk_row0']), int(sh['blk_row1'])
        T.local_ext_rows = int(sh.get('local_ext_rows', 0))
        T.n_owned, T.n_ghost, T.n_ghost_pix = int(sh['n_owned']), int(sh['n_ghost']), int(sh['n_ghost_pix'])
        keep = []

        def arr(name, dtype, ctype):
            a = sh.get(name)
            if a is None:
                return None
            a = np.ascontiguousarray(a, dtype=dtype)
            keep.append(a)
            return _ptr(a, ctype)
        T.mask_global = arr('mask_global', np.uint8, C.c_uint8)
        T.vert, T.weights = arr('vert', np.int32, C.c_int32), arr('weights', np.float64, C.c_double)
        T.sdfunct = arr('sdfunct', np.float64, C.c_double)
        T.vert_back, T.weights_back = arr('vert_back', np.int32, C.c_int32), arr('weights_back', np.float64, C.c_double)
        T.cell_send_ptr, T.cell_recv_ptr = arr('cell_send_ptr', np.int64, C.c_int64), arr('cell_recv_ptr', np.int64, C.c_int64)
        T.pix_send_ptr, T.pix_recv_ptr = arr('pix_send_ptr', np.int64, C.c_int64), arr('pix_recv_ptr', np.int64, C.c_int64)
        T.cell_send_idx, T.pix_send_idx = arr('cell_send_idx', np.int32, C.c_int32), arr('pix_send_idx', np.int32, C.c_int32)
        if sh.get('ghost_pix') is not None and int(sh['n_ghost_pix']) > 0:
            T.ghost_pix = arr('ghost_pix', np.int64, C.c_int64)
        self._check(self.lib.psm_init_sharded(self._h, C.byref(T)))
        self.n_cells, self.W = T.n_owned, T.grid_w
        self.H = T.row1 - T.row0
        self.H_ext = self.H + T.ext_rows + T.local_ext_rows
        self.H_global = T.grid_h
        g = self.geometry()
        self.n_blocks = g['n_blocks']
        return self

    def route_init(self, dest_rank, dest_index):
        """Collective: install the cell route of this rank (``psm_b200.shard.route``): its cells may be any subset of the mesh."""
        dr = np.ascontiguousarray(dest_rank, dtype=np.int32)
        di = np.ascontiguousarray(dest_index, dtype=np.int32)
        R = capi.PsmRoute(n_local=dr.size, dest_rank=_ptr(dr, C.c_int32), dest_index=_ptr(di, C.c_int32))
        self._check(self.lib.psm_route_init(self._h, C.byref(R)))
        self.n_routed = dr.size
        return self

    def predict_routed(self, U, p=None, dU=None, out=None):
        """``predict_fields`` on this rank's own (routed) cells; collective.  Returns (out, status)."""
        U = np.ascontiguousarray(U, dtype=np.float64)
        n = U.shape[0]
        if dU is not None:
            dU = np.ascontiguousarray(dU, dtype=np.float64)
        if p is not None:
            p = np.ascontiguousarray(p, dtype=np.float64)
        if out is None:
            out = np.empty(n if self.n_fields == 1 else (n, 2), dtype=np.float64)
        rc = self._check(self.lib.psm_predict_routed(self._h, U.ctypes.data, U.shape[1], dU.ctypes.data if dU is not None else None,
                                                     p.ctypes.data if p is not None else None, n, out.ctypes.data))
        return out, rc

    def init_from_mesh(self, cells_xy, top, obst, probe_values, back=True):
        t = _tables.build_tables(cells_xy, top, obst, probe_values, variant=self.variant, delta=self.delta, back=back)
        return self.init_tables(t)

    def init_mesh(self, cells_xy, top, obst, probe_values, back=True, cache_dir=None, tables=None):
        """``psm_init_mesh``: bbox / grid / flow mask / distance field / raster inside the library (the distance field on the GPU);
        only the cells -> grid Delaunay comes from here (SciPy's Qhull, like the reference) -- and not even that when the table
        cache (``cache_dir``) already holds this mesh.  ``back``: True = Qhull grid -> cell tables (PMP:211), 'closed_form', or
        False.  ``tables`` = (vert, weights) to reuse tables built elsewhere."""
        cells_xy = np.ascontiguousarray(cells_xy, dtype=np.float64)
        top = np.ascontiguousarray(top, dtype=np.float64)
        obst = np.ascontiguousarray(obst, dtype=np.float64)
        probe = np.ascontiguousarray(probe_values, dtype=np.float64)
        M = capi.PsmMesh()
        M.n_cells, M.cells_xy, M.xy_stride = cells_xy.shape[0], _ptr(cells_xy, C.c_double), 2
        M.top, M.n_top, M.obst, M.n_obst, M.probe = _ptr(top, C.c_double), top.shape[0], _ptr(obst, C.c_double), obst.shape[0], _ptr(probe, C.c_double)
        M.back_closed_form = int(back == 'closed_form')
        keep = [cells_xy, top, obst, probe]
        hit = False
        if cache_dir is not None:
            M.cache_dir = str(cache_dir).encode()
            import os
            hit = os.path.exists(os.path.join(str(cache_dir), 'psm_tables_%s_%s.bin' % (
                mesh_hash(self.variant, self.delta, cells_xy, top, obst, probe), 'cf' if M.back_closed_form else 'qh')))
        if not hit:
            bbox, H, W = mesh_grid(self.variant, self.delta, cells_xy)
            if tables is None:
                X0, Y0 = _tables.uniform_grid(bbox[0], bbox[1], bbox[2], bbox[3], self.delta)
                xy0 = np.stack([X0, Y0], axis=1)
                vert, wts = _tables.barycentric_tables(cells_xy, xy0)                       # UTL:38-44 (Qhull)
                if back is True:
                    vb, wb = _tables.barycentric_tables(xy0, cells_xy)                      # PMP:211
                    keep += [vb, wb]
                    M.vert_back, M.weights_back = _ptr(vb, C.c_int32), _ptr(wb, C.c_double)
            else:
                vert, wts = np.ascontiguousarray(tables[0], dtype=np.int32), np.ascontiguousarray(tables[1], dtype=np.float64)
            keep += [vert, wts]
            M.vert, M.weights = _ptr(vert, C.c_int32), _ptr(wts, C.c_double)
        self._check(self.lib.psm_init_mesh(self._h, C.byref(M)))
        return self._after_init()

    # ------------------------------------------------------------------ per step
    def predict(self, cells, out=None):
        """``py_func`` body: host double[n][input_cols] -> host double[n] (or [n][2]).
        Returns (out, status) with status PSM_OK or PSM_SKIPPED."""
        cells = np.ascontiguousarray(cells, dtype=np.float64)
        if cells.ndim != 2 or cells.shape[1] != self.input_cols:
            raise ValueError('cells must be [n, %d]' % self.input_cols)
        n = cells.shape[0]
        if out is None:
            out = np.empty(n if self.n_fields == 1 else (n, 2), dtype=np.float64)
        rc = self._check(self.lib.psm_predict(self._h, cells.ctypes.data, n, out.ctypes.data))
        return out, rc

    def predict_fields(self, U, p=None, dU=None, out=None):
        """The step on the solver's native field arrays (``psm_predict_fields``): ``U`` host double[n, 3] (OpenFOAM ``vector``)
        or [n, 2]; ``p`` host double[n] or None (then the raw prediction comes back); ``dU`` like ``U`` or None (resident
        U(t-1)).  Returns (out, status)."""
        U = np.ascontiguousarray(U, dtype=np.float64)
        if U.ndim != 2 or U.shape[1] not in (2, 3):
            raise ValueError('U must be [n, 3] or [n, 2]')
        n = U.shape[0]
        if dU is not None:
            dU = np.ascontiguousarray(dU, dtype=np.float64)
            if dU.shape != U.shape:
                raise ValueError('dU must have the shape of U')
        if p is not None:
            p = np.ascontiguousarray(p, dtype=np.float64)
            if p.shape != (n,):
                raise ValueError('p must be [n]')
        if out is None:
            out = np.empty(n if self.n_fields == 1 else (n, 2), dtype=np.float64)
        rc = self._check(self.lib.psm_predict_fields(self._h, U.ctypes.data, U.shape[1], dU.ctypes.data if dU is not None else None,
                                                     p.ctypes.data if p is not None else None, n, out.ctypes.data))
        return out, rc

    def predict_fields_device(self, d_U_ptr, u_stride, n_cells, d_out_ptr, d_p_ptr=0, d_dU_ptr=0, sync=True):
        """Device-pointer form of ``predict_fields`` (ints from e.g. ``torch.Tensor.data_ptr()``; 0 = NULL)."""
        vp = lambda x: C.c_void_p(x) if x else None      # noqa: E731
        return self._check(self.lib.psm_predict_fields_device(self._h, vp(d_U_ptr), int(u_stride), vp(d_dU_ptr), vp(d_p_ptr),
                                                              n_cells, vp(d_out_ptr), int(sync)))

    def predict_device(self, d_cells_ptr, n_cells, d_out_ptr, sync=True):
        """Device-pointer entry (ints from e.g. ``torch.Tensor.data_ptr()``); ``d_out_ptr`` may be 0."""
        return self._check(self.lib.psm_predict_device(self._h, C.c_void_p(d_cells_ptr), n_cells,
                                                       C.c_void_p(d_out_ptr) if d_out_ptr else None, int(sync)))

    def synchronize(self):
        return self._check(self.lib.psm_synchronize(self._h))

    def stream_ptr(self):
        """cudaStream_t of the handle as an int (e.g. for ``torch.cuda.ExternalStream``)."""
        st = C.c_void_p()
        self._check(self.lib.psm_get_stream(self._h, C.byref(st)))
        return st.value or 0

    # ------------------------------------------------------------------ introspection
    def geometry(self):
        g = capi.PsmGeometry()
        self._check(self.lib.psm_get_geometry(self._h, C.byref(g)))
        return {k: getattr(g, k) for k, _ in g._fields_ if k != 'reserved'}

    def plan(self):
        B = self.geometry()['n_blocks']
        o, il = np.zeros((B, 2), np.int32), np.zeros((B, 2), np.int32)
        self._check(self.lib.psm_get_plan(self._h, _ptr(o, C.c_int32), _ptr(il, C.c_int32)))
        return o, il

    def owner_map(self):
        ow = np.zeros((getattr(self, 'H_global', self.H), self.W), np.int32)
        self._check(self.lib.psm_get_owner_map(self._h, _ptr(ow, C.c_int32)))
        return ow

    def forward_table(self):
        v, w = np.zeros((self.H * self.W, 3), np.int32), np.zeros((self.H * self.W, 3), np.float32)
        self._check(self.lib.psm_get_forward_table(self._h, _ptr(v, C.c_int32), _ptr(w, C.c_float)))
        return v, w

    def stage(self, name):
        g = self.geometry()
        Bg, B, F, S, Cn = g['n_blocks'], g['n_local_blocks'], g['n_fields'], g['shape'], self.n_fields
        spec = {
            'grid': (capi.STAGE_GRID, (2, getattr(self, 'H_ext', self.H), self.W), np.float32),
            'x_input': (capi.STAGE_XINPUT, (B, self.pc_in), np.float32),
            'mlp_out': (capi.STAGE_MLPOUT, (B, self.pc_p), np.float32),
            'blocks': (capi.STAGE_BLOCKS, (B, Cn, S, S), np.float32),
            'offsets': (capi.STAGE_OFFSETS, (F, Bg), np.float64),
            'field': (capi.STAGE_FIELD, (F, self.H, self.W), np.float32),
            'scalars': (capi.STAGE_SCALARS, (4,), np.float64),
            'means': (capi.STAGE_MEANS, (g['n_tasks'],), np.float64),
            'xu': (capi.STAGE_XU, (B, 2, S, S), np.float32),
        }[name]
        out = np.empty(spec[1], dtype=spec[2])
        self._check(self.lib.psm_get_stage(self._h, spec[0], out.ctypes.data, out.nbytes))
        return out

    def integrate_gradp(self, top, x0_min, center_row=200):
        """U_to_gradP: pressure field [H, W] recovered from the gradient fields of the last step (GRAD:371-416, 585-628).
        ``top``: the top/bottom boundary points whose bounding box the reference integrates over (GRAD:205);
        ``x0_min``: x of the first grid column."""
        top = np.asarray(top, dtype=np.float64)
        g = capi.PsmIntegrateGeometry(min_x=float(top[:, 0].min()), max_x=float(top[:, 0].max()), min_y=float(top[:, 1].min()),
                                      max_y=float(top[:, 1].max()), x0_min=float(x0_min), center_row=int(center_row))
        out = np.empty((self.H, self.W), dtype=np.float64)
        self._check(self.lib.psm_integrate_gradp(self._h, C.byref(g), out.ctypes.data))
        return out

    def timings(self):
        ms = np.zeros(capi.N_TIMINGS, np.float32)
        self._check(self.lib.psm_get_timings(self._h, _ptr(ms, C.c_float), capi.N_TIMINGS))
        return dict(zip(capi.TIMING_NAMES, ms.tolist()))

    def set_timings(self, on):
        self._check(self.lib.psm_set_timings(self._h, int(on)))

    def launch_count(self):
        return int(self.lib.psm_get_launch_count(self._h))

    def wait_ns(self, reset=True):
        """Multi-GPU wait histogram (psm_get_wait_ns): {'ns': [3], 'count': [3]} per exchange phase since the last reset."""
        ns, cnt = (C.c_uint64 * 3)(), (C.c_uint32 * 3)()
        self._check(self.lib.psm_get_wait_ns(self._h, ns, cnt, int(bool(reset))))
        return {'ns': [int(x) for x in ns], 'count': [int(x) for x in cnt]}
