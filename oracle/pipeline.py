"""Oracle: the surrogate pipelines end to end (test infrastructure).

``DeltasOracle``  -- Improved_SM deltaU_to_deltaP: SMC:89-180 (init), SMC:382-575 minus the
                     label / metric / plot lines (step), plus the solver-side grid->cell
                     back-interpolation of PMP:481-496.
``GradPOracle``   -- U_to_gradP: GRAD:160-253 (init), GRAD:418-547 (step).

Everything is float64 except the MLP, which is float32 like the Keras model
(NNS:24-33); see ``oracle/__init__.py`` for the substitutions.  All intermediates are
returned so that tests can compare stage by stage.
"""
import numpy as np

from . import interp as _interp
from . import domain as _domain
import scipy.ndimage as _ndimage

from . import assemble as _assemble


class SurrogateParams:
    """The artefacts the reference loads at import/initialisation time (SMC:70-87,
    SMC:505-511, PMP:103-118): ``maxs`` scalars, the two PCA maps, the PCA-space scaler
    and the Dense stack.  All arrays are NumPy; MLP kernels are Keras-shaped [in, out]."""

    def __init__(self, maxs, pca_in_components, pca_in_mean, pca_out_components, pca_out_mean,
                 mlp_weights, mlp_biases, standardization='std', mean_in=None, std_in=None,
                 mean_out=None, std_out=None, max_abs_input_PCA=None, max_abs_output_PCA=None,
                 n_out_channels=1, min_in=None, max_in=None, min_out=None, max_out=None):
        self.maxs = np.asarray(maxs, dtype=np.float64)
        self.pca_in_components = np.asarray(pca_in_components)       # [pc_in, S*S*3]
        self.pca_in_mean = np.asarray(pca_in_mean)                   # [S*S*3]
        self.pca_out_components = np.asarray(pca_out_components)     # [pc_p, S*S*C]
        self.pca_out_mean = np.asarray(pca_out_mean)                 # [S*S*C]
        self.mlp_weights = [np.asarray(w, dtype=np.float32) for w in mlp_weights]
        self.mlp_biases = [np.asarray(b, dtype=np.float32) for b in mlp_biases]
        self.standardization = standardization
        self.mean_in, self.std_in = mean_in, std_in
        self.mean_out, self.std_out = mean_out, std_out
        self.max_abs_input_PCA = max_abs_input_PCA
        self.max_abs_output_PCA = max_abs_output_PCA
        self.min_in, self.max_in, self.min_out, self.max_out = min_in, max_in, min_out, max_out     # 'min_max', SMC:513-520
        self.n_out_channels = n_out_channels

    @property
    def pc_in(self):
        return self.pca_in_components.shape[0]

    @property
    def pc_p(self):
        return self.pca_out_components.shape[0]


def select_num_pc(explained_variance_ratio, var, max_num_PC):
    """SMC:86-87 -- argmax(cumsum > var) if 1 < argmax <= max_num_PC else max_num_PC."""
    a = int(np.argmax(np.asarray(explained_variance_ratio).cumsum() > var))
    return a if (a > 1 and a <= max_num_PC) else max_num_PC


def mlp_forward(x, weights, biases):
    """NNS:24-33 -- Dense(relu) x n_layers then a linear Dense; float32 like Keras
    (the float64 input is cast to float32 on entry to the model, SMC:530)."""
    h = np.asarray(x, dtype=np.float32)
    for li, (w, b) in enumerate(zip(weights, biases)):
        h = h @ w + b
        if li < len(weights) - 1:
            h = np.maximum(h, np.float32(0))
    return h


def pca_transform(x_flat, components, mean):
    """``pca.transform(X)[:, :pc]`` for whiten=False, as spelled out at PMP:349."""
    return np.dot(x_flat - mean, components.T)


class DeltasOracle:
    """Improved_SM deltaU_to_deltaP (SMC) with the solver-side back-interpolation (PMP)."""

    def __init__(self, params, delta=5e-3, shape=128, overlap=32):
        self.params = params
        self.delta = delta
        self.shape = shape
        self.overlap = overlap

    # ------------------------------------------------------------------ init
    def compute_only_once(self, cell_xy, top, obst, probe_values, tables=None, back_tables=True,
                          literal_raster=False):
        """SMC:89-180.  ``cell_xy`` [N,2] cell centres, ``top``/``obst`` boundary points,
        ``probe_values`` [N] the field whose interpolation decides pixel validity (``p`` of
        the first frame, SMC:165-169).  ``tables`` = (vert, weights[, vert_back, weights_back])
        to reuse precomputed Qhull tables (e.g. the fast structured builder for huge meshes).
        The back tables are the solver-side addition PMP:211."""
        cell_xy = np.asarray(cell_xy, dtype=np.float64)
        x_min = round(np.min(cell_xy[:, 0]), 3)                 # SMC:102-106
        x_max = round(np.max(cell_xy[:, 0]), 3)
        y_min = round(np.min(cell_xy[:, 1]), 3)
        y_max = round(np.max(cell_xy[:, 1]), 3)
        X0, Y0 = _interp.create_uniform_grid(x_min, x_max, y_min, y_max, self.delta)
        self.X0, self.Y0 = X0, Y0
        xy0 = np.concatenate((np.expand_dims(X0, axis=1), np.expand_dims(Y0, axis=1)), axis=-1)
        self.xy0 = xy0
        if tables is None:
            self.vert, self.weights = _interp.interp_weights(cell_xy, xy0)       # SMC:115
            if back_tables:
                self.vert_back, self.weights_back = _interp.interp_weights(xy0, cell_xy)   # PMP:211
        else:
            self.vert, self.weights = tables[0], tables[1]
            if len(tables) > 2:
                self.vert_back, self.weights_back = tables[2], tables[3]
        domain_bool, sdf, _ = _domain.domain_dist(xy0, np.asarray(top), np.asarray(obst), 'smc',
                                                  x_min, x_max, y_min, y_max)
        self.domain_bool, self.sdf = domain_bool, sdf
        self.grid_shape_y = int(round((y_max - y_min) / self.delta))            # SMC:148-149
        self.grid_shape_x = int(round((x_max - x_min) / self.delta))
        p_interp = _interp.interpolate_fill(np.asarray(probe_values, dtype=np.float64), self.vert, self.weights)
        self.indices, self.sdfunct = _domain.index_raster(
            X0, Y0, self.delta, self.grid_shape_y, self.grid_shape_x, domain_bool, p_interp, sdf,
            literal=literal_raster)
        return 0

    # ------------------------------------------------------------------ plan
    def block_plan(self):
        """SMC:458-479 -- extraction origins, right->left then top->bottom."""
        H, W = self.grid_shape_y, self.grid_shape_x
        shape, overlap = self.shape, self.overlap
        n_x = int(np.ceil((W - shape) / (shape - overlap)))
        n_y = int((H - shape) / (shape - overlap))
        origins, indices_list = [], []
        for i in range(n_y + 2):
            for j in range(n_x + 1):
                x_0 = W - j * shape + j * overlap - shape
                if j == n_x:
                    x_0 = 0
                y_0 = i * shape - i * overlap
                if i == n_y + 1:
                    y_0 = H - shape
                origins.append((y_0, x_0))
                indices_list.append([i, n_x - j])
        return n_x, n_y, origins, indices_list

    # ------------------------------------------------------------------ step
    def time_step(self, Ux, Uy, dUx, dUy, threshold=1e-4, apply_filter=False):
        """SMC:382-575 (inference part).  Returns a dict of intermediates; ``None`` for an
        'irrelevant' time step (SMC:407-415)."""
        P = self.params
        Ux = np.asarray(Ux, dtype=np.float64).reshape(-1, 1)
        Uy = np.asarray(Uy, dtype=np.float64).reshape(-1, 1)
        dUx = np.asarray(dUx, dtype=np.float64).reshape(-1, 1)
        dUy = np.asarray(dUy, dtype=np.float64).reshape(-1, 1)
        U_max_norm = np.max(np.sqrt(np.square(Ux) + np.square(Uy)))            # SMC:404
        deltaU_max_norm = np.max(np.sqrt(np.square(dUx) + np.square(dUy)))
        if (deltaU_max_norm / U_max_norm) < threshold:                        # SMC:410-415
            return None
        dUx_adim = dUx / U_max_norm
        dUy_adim = dUy / U_max_norm
        dUx_interp = _interp.interpolate_fill(dUx_adim, self.vert, self.weights)   # SMC:422-423
        dUy_interp = _interp.interpolate_fill(dUy_adim, self.vert, self.weights)

        H, W = self.grid_shape_y, self.grid_shape_x
        grid = np.zeros(shape=(1, H, W, 3))
        grid[0, :, :, 0:1][tuple(self.indices.T)] = dUx_interp.reshape(-1, 1)     # SMC:432-434
        grid[0, :, :, 1:2][tuple(self.indices.T)] = dUy_interp.reshape(-1, 1)
        grid[0, :, :, 2:3] = self.sdfunct
        grid[np.isnan(grid)] = 0
        max_abs_Ux, max_abs_Uy, max_abs_dist, max_abs_p = P.maxs[0], P.maxs[1], P.maxs[2], P.maxs[3]
        grid[0, :, :, 0:1] = grid[0, :, :, 0:1] / max_abs_Ux                      # SMC:441-443
        grid[0, :, :, 1:2] = grid[0, :, :, 1:2] / max_abs_Uy
        grid[0, :, :, 2:3] = grid[0, :, :, 2:3] / max_abs_dist

        shape, overlap = self.shape, self.overlap
        n_x, n_y, origins, indices_list = self.block_plan()
        x_list = [grid[0:1, y0:y0 + shape, x0:x0 + shape, 0:3] for (y0, x0) in origins]   # SMC:476
        x_array = np.concatenate(x_list)
        N = x_array.shape[0]
        input_flat = x_array.reshape((N, -1))                                     # SMC:491-492
        input_transformed = pca_transform(input_flat, P.pca_in_components, P.pca_in_mean)   # SMC:494

        if P.standardization == 'std':                                            # SMC:505-523
            x_input = (input_transformed - P.mean_in) / P.std_in
        elif P.standardization == 'min_max':
            x_input = (input_transformed - P.min_in) / (P.max_in - P.min_in)
        elif P.standardization == 'max_abs':
            x_input = input_transformed / P.max_abs_input_PCA
        else:
            raise ValueError("Standardization method not valid")
        res_concat = mlp_forward(x_input, P.mlp_weights, P.mlp_biases)            # SMC:530
        mlp_out = res_concat.copy()
        if P.standardization == 'std':                                            # SMC:532-537
            res_concat = (res_concat * P.std_out) + P.mean_out
        elif P.standardization == 'min_max':
            res_concat = res_concat * (P.max_out - P.min_out) + P.min_out
        else:
            res_concat = res_concat * P.max_abs_output_PCA
        res_flat_inv = np.dot(res_concat, P.pca_out_components) + P.pca_out_mean  # SMC:541
        blocks = res_flat_inv.reshape((N, shape, shape, 1))
        blocks = blocks * max_abs_p * pow(U_max_norm, 2.0)                        # SMC:551
        field, offsets, shift = _assemble.assemble_deltas(
            blocks[..., 0], x_array, indices_list, n_x, n_y, shape, overlap, W, H,
            Ref_BC=0.0, return_offsets=True)                                      # SMC:570-575
        if apply_filter:                                                          # SMC:353-356 (the same SciPy call)
            field = _ndimage.gaussian_filter(field, sigma=(10, 10), order=0)
        return dict(U_max_norm=U_max_norm, grid=grid[0], x_array=x_array, z=input_transformed,
                    x_input=x_input, mlp_out=mlp_out, blocks=blocks, offsets=offsets, shift=shift,
                    field=field, n_x=n_x, n_y=n_y, origins=origins, indices_list=indices_list)

    def to_cells(self, field, p_prev=None, additive=True, near_wall_sdf=None):
        """Solver-side grid->cell back-interpolation, PMP:481-496: gather through ``indices``
        (with the (0,0) quirk), ``interpolate_fill`` with the grid->cell tables, optional
        near-wall fallback (``sdf_mesh < near_wall_sdf``, PMP:492-494), NaN -> previous
        pressure.  deltaU_to_deltaP returns ``p_prev + delta_p`` (SMC:644-645: "p_t-1 + delta_p")."""
        unif = field[tuple(self.indices.T)]                                       # PMP:481
        interp = _interp.interpolate_fill(unif, self.vert_back, self.weights_back)    # PMP:485
        nan = np.isnan(interp)
        dp = interp.copy()
        use_prev = nan.copy()
        if near_wall_sdf is not None:
            # PMP:492 -- np.take on the 2-D raster indexes it flat (row-major), no `indices` hop
            sdf_mesh = _interp.interpolate_fill(self.sdfunct[:, :, 0], self.vert_back, self.weights_back)
            use_prev |= (sdf_mesh < near_wall_sdf)
        if p_prev is None:
            p_prev = np.zeros_like(dp)
        p_prev = np.asarray(p_prev, dtype=np.float64)
        if additive:
            dp[use_prev] = 0.0
            return p_prev + dp, interp
        out = dp
        out[use_prev] = p_prev[use_prev]
        return out, interp


class GradPOracle:
    """U_to_gradP (GRAD): two-channel output, left->right plan, scalar max-abs PCA scaling."""

    def __init__(self, params, delta=5e-3, shape=128, avance=96):
        self.params = params
        self.delta = delta
        self.shape = shape
        self.avance = avance

    def compute_only_once(self, cell_xy, top, obst, probe_values, tables=None, back_tables=True,
                          literal_raster=False):
        """GRAD:160-253 (bbox rounded to 2 decimals, GRAD:174-178)."""
        cell_xy = np.asarray(cell_xy, dtype=np.float64)
        x_min = round(np.min(cell_xy[:, 0]), 2)
        x_max = round(np.max(cell_xy[:, 0]), 2)
        y_min = round(np.min(cell_xy[:, 1]), 2)
        y_max = round(np.max(cell_xy[:, 1]), 2)
        X0, Y0 = _interp.create_uniform_grid(x_min, x_max, y_min, y_max, self.delta)
        self.X0, self.Y0 = X0, Y0
        xy0 = np.concatenate((np.expand_dims(X0, axis=1), np.expand_dims(Y0, axis=1)), axis=-1)
        self.xy0 = xy0
        if tables is None:
            self.vert, self.weights = _interp.interp_weights(cell_xy, xy0)       # GRAD:187
            if back_tables:
                self.vert_back, self.weights_back = _interp.interp_weights(xy0, cell_xy)
        else:
            self.vert, self.weights = tables[0], tables[1]
            if len(tables) > 2:
                self.vert_back, self.weights_back = tables[2], tables[3]
        domain_bool, sdf, bb = _domain.domain_dist(xy0, np.asarray(top), np.asarray(obst), 'grad')
        self.min_x, self.max_x, self.min_y, self.max_y = bb
        self.domain_bool, self.sdf = domain_bool, sdf
        self.grid_shape_y = int(round((y_max - y_min) / self.delta))
        self.grid_shape_x = int(round((x_max - x_min) / self.delta))
        probe = _interp.interpolate_fill(np.asarray(probe_values, dtype=np.float64), self.vert, self.weights)
        self.indices, self.sdfunct = _domain.index_raster(
            X0, Y0, self.delta, self.grid_shape_y, self.grid_shape_x, domain_bool, probe, sdf,
            literal=literal_raster)
        return 0

    def block_plan(self):
        """GRAD:476-500 -- extraction origins, left->right then top->bottom."""
        H, W = self.grid_shape_y, self.grid_shape_x
        shape, avance = self.shape, self.avance
        n_x = int(np.ceil((W - shape) / (shape - avance)))
        n_y = int((H - shape) / (shape - avance))
        origins, indices_list = [], []
        for i in range(n_y + 2):
            for j in range(n_x + 1):
                x_0 = j * shape - j * avance
                if j == n_x:
                    x_0 = W - shape
                y_0 = i * shape - i * avance
                if i == n_y + 1:
                    y_0 = H - shape
                origins.append((y_0, x_0))
                indices_list.append([i, j])
        return n_x, n_y, origins, indices_list

    def time_step(self, Ux, Uy, apply_filter=False):
        """GRAD:429-547 (inference part)."""
        P = self.params
        Ux = np.asarray(Ux, dtype=np.float64).reshape(-1, 1)
        Uy = np.asarray(Uy, dtype=np.float64).reshape(-1, 1)
        U_max_norm = np.max(np.sqrt(np.square(Ux) + np.square(Uy)))            # GRAD:439
        Ux_adim = Ux / U_max_norm
        Uy_adim = Uy / U_max_norm
        Ux_interp = _interp.interpolate_fill(Ux_adim, self.vert, self.weights)    # GRAD:450-451
        Uy_interp = _interp.interpolate_fill(Uy_adim, self.vert, self.weights)
        H, W = self.grid_shape_y, self.grid_shape_x
        grid = np.zeros(shape=(1, H, W, 3))
        grid[0, :, :, 0:1][tuple(self.indices.T)] = Ux_interp.reshape(-1, 1)      # GRAD:455-457
        grid[0, :, :, 1:2][tuple(self.indices.T)] = Uy_interp.reshape(-1, 1)
        grid[0, :, :, 2:3] = self.sdfunct
        grid[np.isnan(grid)] = 0
        grid[0, :, :, 0:1] = grid[0, :, :, 0:1] / P.maxs[0]                       # GRAD:464-466
        grid[0, :, :, 1:2] = grid[0, :, :, 1:2] / P.maxs[1]
        grid[0, :, :, 2:3] = grid[0, :, :, 2:3] / P.maxs[2]

        shape, avance = self.shape, self.avance
        n_x, n_y, origins, indices_list = self.block_plan()
        x_array = np.concatenate([grid[0:1, y0:y0 + shape, x0:x0 + shape, 0:3] for (y0, x0) in origins])
        N = x_array.shape[0]
        input_flat = x_array.reshape((N, -1))
        input_transformed = pca_transform(input_flat, P.pca_in_components, P.pca_in_mean)   # GRAD:518
        x_input = input_transformed / P.max_abs_input_PCA                         # GRAD:525
        res_concat = mlp_forward(x_input, P.mlp_weights, P.mlp_biases)            # GRAD:530
        mlp_out = res_concat.copy()
        res_concat = res_concat * P.max_abs_output_PCA                            # GRAD:531
        res_flat_inv = np.dot(res_concat, P.pca_out_components) + P.pca_out_mean  # GRAD:533
        blocks = res_flat_inv.reshape((N, shape, shape, 2))                       # GRAD:534
        out = {}
        for ch, name in enumerate(('dp_dx', 'dp_dy')):                            # GRAD:544-545
            f, offs, sh = _assemble.assemble_gradp(name, blocks[..., ch], x_array, indices_list, n_x, n_y,
                                                   shape, avance, W, H, Ref_BC=0.0, return_offsets=True)
            out[name] = f[0, :, :, 0]
            if apply_filter:                                                      # GRAD:366-367
                out[name] = _ndimage.gaussian_filter(out[name], sigma=(10, 10), order=0)
            out[name + '_offsets'] = offs
            out[name + '_shift'] = sh
        out.update(U_max_norm=U_max_norm, grid=grid[0], x_array=x_array, z=input_transformed,
                   x_input=x_input, mlp_out=mlp_out, blocks=blocks, n_x=n_x, n_y=n_y,
                   origins=origins, indices_list=indices_list)
        return out

    def to_cells(self, field):
        """Grid->cell back-interpolation in the style of PMP:481-485 (NaN kept as NaN: the
        variant has no previous-gradient fallback)."""
        unif = field[tuple(self.indices.T)]
        return _interp.interpolate_fill(unif, self.vert_back, self.weights_back)


class ThesisOracle:
    """Thesis solver module (PMP: ``init_func`` PMP:172-247, ``py_func`` PMP:249-517): U -> p on blocks cut with
    ``avance = int(0.1 * 128) = 12`` shared columns, extra left-most block column, scalar max-abs PCA scaling,
    distance channel fed unscaled (PMP:292), forward interpolation WITHOUT the NaN fill (PMP:64-65,280-281)."""

    def __init__(self, params, delta=5e-3, shape=128):
        self.params = params
        self.delta = delta
        self.shape = shape
        self.avance = int(0.1 * shape)                                            # PMP:304

    def init_func(self, cell_xy, top, obst, ux_probe, tables=None):
        cell_xy = np.asarray(cell_xy, dtype=np.float64)
        x_min = round(np.min(cell_xy[:, 0]), 2)                                   # PMP:197-201
        x_max = round(np.max(cell_xy[:, 0]), 2)
        y_min = round(np.min(cell_xy[:, 1]), 2)
        y_max = round(np.max(cell_xy[:, 1]), 2)
        X0, Y0 = _interp.create_uniform_grid(x_min, x_max, y_min, y_max, self.delta)
        self.X0, self.Y0 = X0, Y0
        xy0 = np.concatenate((np.expand_dims(X0, axis=1), np.expand_dims(Y0, axis=1)), axis=-1)
        if tables is None:
            self.vert, self.weights = _interp.interp_weights(cell_xy, xy0, idw_fallback=False)      # PMP:210
            self.vert_back, self.weights_back = _interp.interp_weights(xy0, cell_xy, idw_fallback=False)   # PMP:211
        else:
            self.vert, self.weights, self.vert_back, self.weights_back = tables
        domain_bool, sdf, _ = _domain.domain_dist(xy0, np.asarray(top), np.asarray(obst), 'pmp')    # PMP:214
        self.grid_shape_y = int(round((y_max - y_min) / self.delta))
        self.grid_shape_x = int(round((x_max - x_min) / self.delta))
        ux_interp = _interp.interpolate_fill(np.asarray(ux_probe, dtype=np.float64), self.vert, self.weights)   # PMP:230-231
        # PMP:225 allocates `indices` with np.empty; the raster leaves invalid points untouched -- zero in practice
        self.indices, self.sdfunct = _domain.index_raster(X0, Y0, self.delta, self.grid_shape_y, self.grid_shape_x,
                                                          domain_bool, ux_interp, sdf)
        self.valid_rows = np.asarray(domain_bool, dtype=bool) & ~np.isnan(ux_interp)      # rows PMP:233-243 writes
        return 0

    def block_plan(self):
        """PMP:303-332: right -> left, and after the last regular block of every row the block at x = 0 tagged -1."""
        H, W, shape, avance = self.grid_shape_y, self.grid_shape_x, self.shape, self.avance
        n_x = int((W - shape) / (shape - avance))
        n_y = int((H - shape) / (shape - avance))
        origins, indices_list = [], []
        for i in range(n_y + 2):
            for j in range(n_x + 1):
                x_0 = (W - shape) - j * shape + j * avance
                y_0 = (H - shape) if i == n_y + 1 else i * shape - i * avance
                origins.append((y_0, x_0))
                indices_list.append([i, n_x - j])
                if j == n_x:
                    origins.append((y_0, 0))
                    indices_list.append([i, -1])
        return n_x, n_y, origins, indices_list

    def py_func(self, Ux, Uy, p_prev, near_wall_sdf=0.05):
        """PMP:249-517 on one rank.  Returns a dict with the cell pressures ``p`` and intermediates."""
        P = self.params
        Ux = np.asarray(Ux, dtype=np.float64).reshape(-1, 1)
        Uy = np.asarray(Uy, dtype=np.float64).reshape(-1, 1)
        p_prev = np.asarray(p_prev, dtype=np.float64)
        U_max_norm = np.max(np.sqrt(np.square(Ux) + np.square(Uy)))              # PMP:270
        Ux_interp = _interp.interpolate(Ux / U_max_norm, self.vert, self.weights)     # PMP:280-281 (no NaN fill)
        Uy_interp = _interp.interpolate(Uy / U_max_norm, self.vert, self.weights)
        H, W = self.grid_shape_y, self.grid_shape_x
        grid = np.zeros(shape=(1, H, W, 3))
        grid[0, :, :, 0:1][tuple(self.indices.T)] = Ux_interp.reshape(-1, 1) / P.maxs[0]      # PMP:290-292
        grid[0, :, :, 1:2][tuple(self.indices.T)] = Uy_interp.reshape(-1, 1) / P.maxs[1]
        grid[0, :, :, 2:3] = self.sdfunct                                         # unscaled
        grid[np.isnan(grid)] = 0
        shape, avance = self.shape, self.avance
        n_x, n_y, origins, indices_list = self.block_plan()
        x_array = np.concatenate([grid[0:1, y0:y0 + shape, x0:x0 + shape, 0:3] for (y0, x0) in origins])
        N = x_array.shape[0]
        input_flat = x_array.reshape((N, -1))
        input_transformed = pca_transform(input_flat, P.pca_in_components, P.pca_in_mean)     # PMP:349
        x_input = input_transformed / P.max_abs_input_PCA                         # PMP:351
        res_concat = mlp_forward(x_input, P.mlp_weights, P.mlp_biases)            # PMP:360
        mlp_out = res_concat.copy()
        res_flat_inv = np.dot(res_concat * P.max_abs_output_PCA, P.pca_out_components) + P.pca_out_mean   # PMP:365
        blocks = res_flat_inv.reshape((N, shape, shape, 1))
        field, offsets, shift = _assemble.assemble_thesis(blocks[..., 0], x_array, indices_list, n_x, n_y, shape, avance,
                                                          W, H, return_offsets=True)
        p_adim_unif = field[tuple(self.indices.T)]                                # PMP:481
        p_interp = _interp.interpolate_fill(p_adim_unif, self.vert_back, self.weights_back)   # PMP:485
        p = p_interp * P.maxs[3] * pow(U_max_norm, 2.0)                           # PMP:490
        if near_wall_sdf is not None:
            sdf_mesh = _interp.interpolate_fill(self.sdfunct[:, :, 0], self.vert_back, self.weights_back)   # PMP:492
            p[sdf_mesh < near_wall_sdf] = p_prev[sdf_mesh < near_wall_sdf]
        p[np.isnan(p_interp)] = p_prev[np.isnan(p_interp)]                        # PMP:496
        return dict(p=p, U_max_norm=U_max_norm, grid=grid[0], x_array=x_array, x_input=x_input, mlp_out=mlp_out,
                    blocks=blocks, field=field, offsets=offsets, shift=shift, n_x=n_x, n_y=n_y, origins=origins,
                    indices_list=indices_list)
