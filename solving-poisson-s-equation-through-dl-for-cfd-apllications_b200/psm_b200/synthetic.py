"""Seeded synthetic meshes, flow fields and surrogate parameters.

The reference ships neither datasets nor the PCA pickles (``.MISSING_LARGE_BLOBS``), so
tests, ``smoke()`` and ``bench.py`` run on synthetic flow-past-cylinder inputs and
random-init parameters of the reference architecture (PCA -> 3x512 ReLU MLP -> PCA^-1,
UTL:437-439, NNS:8-38).  NumPy only; everything is a plain dict of arrays so that the
CUDA path and the CPU oracle are fed identical bytes.
"""
import numpy as np

# shipped ``Thesis_Work/Chapter5/parallelized/test_case/maxs`` (max_abs_Ux, Uy, dist, p)
DEFAULT_MAXS = (1.0, 0.536, 0.999, 0.511)


def make_mesh(H, W, nx, ny, R=0.4, seed=0, delta=5e-3, n_top=3000, n_obst=400, jitter=0.3, center=(0.0, 0.0)):
    """Jittered-lattice cell centres around a cylinder of radius R at the origin.

    The cell-centre bounding box is exactly ``W*delta x H*delta`` (so the uniform grid of
    UTL:111-125 is ``H x W``), x in [-0.3*W*delta, 0.7*W*delta], centred on the obstacle in y.
    Interior centres are jittered by +-``jitter`` lattice spacings (general position =>
    unique Delaunay), the outermost ring is not, so the convex hull equals the bbox.
    ``center`` moves the cylinder (used to swallow whole overlap strips in NaN-chain tests).
    Returns dict(cells[N,2], top[2*n_top,2], obst[n_obst,2], H, W, delta, R).
    """
    rng = np.random.default_rng(seed)
    Lx, Ly = W * delta, H * delta
    x_min, x_max = -0.3 * Lx, 0.7 * Lx
    y_min, y_max = -0.5 * Ly, 0.5 * Ly
    xs = np.linspace(x_min, x_max, nx)
    ys = np.linspace(y_min, y_max, ny)
    XX, YY = np.meshgrid(xs, ys)
    hx, hy = xs[1] - xs[0], ys[1] - ys[0]
    XX[1:-1, 1:-1] += rng.uniform(-jitter, jitter, size=(ny - 2, nx - 2)) * hx
    YY[1:-1, 1:-1] += rng.uniform(-jitter, jitter, size=(ny - 2, nx - 2)) * hy
    cx, cy = center
    keep = ((XX - cx) ** 2 + (YY - cy) ** 2) > R * R
    cells = np.stack([XX[keep], YY[keep]], axis=1)
    tx = np.linspace(x_min, x_max, n_top)
    top = np.concatenate([np.stack([tx, np.full(n_top, y_max)], axis=1),
                          np.stack([tx, np.full(n_top, y_min)], axis=1)])
    th = np.linspace(0.0, 2.0 * np.pi, n_obst, endpoint=False)
    obst = np.stack([cx + R * np.cos(th), cy + R * np.sin(th)], axis=1)
    return dict(cells=np.ascontiguousarray(cells), top=top, obst=obst, H=H, W=W, delta=delta, R=R, center=(cx, cy),
                bbox=(x_min, x_max, y_min, y_max))


def _fourier_field(xy, rng, n_modes, scale):
    f = np.zeros(xy.shape[0])
    for _ in range(n_modes):
        k = rng.uniform(-1.0, 1.0, size=2) * 2.0 * np.pi / scale
        f += rng.normal() * np.sin(xy[:, 0] * k[0] + xy[:, 1] * k[1] + rng.uniform(0, 2 * np.pi))
    return f / np.sqrt(n_modes)


def make_fields(mesh, seed=0, dU_scale=1e-2, stress=False):
    """Smooth seeded fields on the cells: U (max|U| ~ 1), dU = dU_scale * independent smooth
    field, p_prev.  ``stress=True`` gives i.i.d. N(0,1) fields instead."""
    rng = np.random.default_rng(seed + 1000003)
    xy = mesh['cells']
    n = xy.shape[0]
    if stress:
        return dict(Ux=rng.normal(size=n), Uy=rng.normal(size=n), dUx=dU_scale * rng.normal(size=n),
                    dUy=dU_scale * rng.normal(size=n), p_prev=rng.normal(size=n))
    R = mesh['R']
    L = max(mesh['W'], mesh['H']) * mesh['delta']
    x, y = xy[:, 0], xy[:, 1]
    x, y = x - mesh['center'][0], y - mesh['center'][1]
    r2 = np.maximum(x * x + y * y, R * R)
    Ux = 1.0 - R * R * (x * x - y * y) / (r2 * r2) + 0.1 * _fourier_field(xy, rng, 8, L)
    Uy = -2.0 * R * R * x * y / (r2 * r2) + 0.1 * _fourier_field(xy, rng, 8, L)
    m = np.max(np.sqrt(Ux * Ux + Uy * Uy))
    Ux, Uy = Ux / m, Uy / m
    dUx = dU_scale * _fourier_field(xy, rng, 8, L / 4)
    dUy = dU_scale * _fourier_field(xy, rng, 8, L / 4)
    p_prev = 0.5 * _fourier_field(xy, rng, 8, L)
    return dict(Ux=Ux, Uy=Uy, dUx=dUx, dUy=dUy, p_prev=p_prev)


def pack_cells(mesh, fields, with_delta=True):
    """Row-major ``double[n][5]`` = {Ux, Uy, Cx, Cy, p} as FOAM/PythonComm.H:2-9 fills it, or
    ``double[n][7]`` with the two extra deltaU columns."""
    cols = [fields['Ux'], fields['Uy'], mesh['cells'][:, 0], mesh['cells'][:, 1], fields['p_prev']]
    if with_delta:
        cols += [fields['dUx'], fields['dUy']]
    return np.ascontiguousarray(np.stack(cols, axis=1), dtype=np.float64)


def make_params(seed=0, shape=128, n_out_channels=1, pc_in=128, pc_p=128, hidden=(512, 512, 512),
                standardization='std', maxs=DEFAULT_MAXS):
    """Random-init parameters of the reference architecture.

    PCA ``components_`` ~ N(0, 1/K) float32, ``mean_`` ~ N(0, 0.1^2); PCA-space scaler
    mean ~ N(0, 0.1^2), std ~ U(0.5, 1.5); MLP kernels Glorot-uniform [in, out] (Keras
    layout, NNS:24-33), biases ~ N(0, 0.01^2)."""
    rng = np.random.default_rng(seed + 7919)
    K_in = shape * shape * 3
    K_out = shape * shape * n_out_channels
    p = dict(maxs=np.asarray(maxs, dtype=np.float64), shape=shape, n_out_channels=n_out_channels,
             standardization=standardization)
    p['pca_in_components'] = (rng.standard_normal((pc_in, K_in), dtype=np.float32) / np.float32(np.sqrt(K_in)))
    p['pca_in_mean'] = (0.1 * rng.standard_normal(K_in)).astype(np.float32)
    p['pca_out_components'] = (rng.standard_normal((pc_p, K_out), dtype=np.float32) / np.float32(np.sqrt(K_out)))
    p['pca_out_mean'] = (0.1 * rng.standard_normal(K_out)).astype(np.float32)
    if standardization == 'std':
        p['mean_in'] = 0.1 * rng.standard_normal(pc_in)
        p['std_in'] = rng.uniform(0.5, 1.5, size=pc_in)
        p['mean_out'] = 0.1 * rng.standard_normal(pc_p)
        p['std_out'] = rng.uniform(0.5, 1.5, size=pc_p)
    elif standardization == 'min_max':                    # SMC:513-520 (min_max_values.npz)
        p['min_in'] = -1.0 + 0.1 * rng.standard_normal(pc_in)
        p['max_in'] = 1.0 + 0.1 * rng.standard_normal(pc_in)
        p['min_out'] = -1.5 + 0.1 * rng.standard_normal(pc_p)
        p['max_out'] = 1.5 + 0.1 * rng.standard_normal(pc_p)
    else:
        p['max_abs_input_PCA'] = 1.7
        p['max_abs_output_PCA'] = 2.3
    dims = [pc_in] + list(hidden) + [pc_p]
    ws, bs = [], []
    for a, b in zip(dims[:-1], dims[1:]):
        lim = np.sqrt(6.0 / (a + b))
        ws.append(rng.uniform(-lim, lim, size=(a, b)).astype(np.float32))
        bs.append((0.01 * rng.standard_normal(b)).astype(np.float32))
    p['mlp_weights'] = ws
    p['mlp_biases'] = bs
    return p


# BASELINE.json configs: name -> mesh arguments
CONFIGS = {
    'c1': dict(H=400, W=3000, nx=866, ny=58, R=0.4),       # ~49.4 k cells, B=124
    'c2': dict(H=1000, W=1000, nx=1000, ny=1000, R=0.4),   # ~0.98 M cells, B=121 (841 for gradP)
    'c4': dict(H=4000, W=4000, nx=4000, ny=4000, R=0.4),   # ~16 M cells, B=1764
    'c5': dict(H=2000, W=2000, nx=2000, ny=2000, R=0.4),   # ~4 M cells, B=441
    'tiny': dict(H=300, W=420, nx=160, ny=110, R=0.1),     # test fixture, B=4x5
}
