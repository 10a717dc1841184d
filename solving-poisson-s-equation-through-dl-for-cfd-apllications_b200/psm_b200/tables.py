"""Once-per-mesh tables for the C-ABI (``psm_init_with_tables``).

The interpolation tables are built with SciPy's Qhull binding -- the library the reference
itself calls (UTL:38-44, PMP:154-161) -- so they are bit-identical to the reference's; the mask,
distance field and raster are vectorised restatements of SMC:117-178 / GRAD:190-248 / PMP:72-99,
216-243.  Per-step work never comes back here: everything below runs once in ``init_func``.
"""
import numpy as np
from scipy.spatial import ConvexHull, Delaunay, cKDTree

VARIANTS = ('deltaU_to_deltaP', 'U_to_gradP', 'thesis')


def uniform_grid(x_min, x_max, y_min, y_max, delta):
    """Cell-centred uniform grid flattened row-major, y outer (UTL:111-125)."""
    X0 = np.linspace(x_min + delta / 2, x_max - delta / 2, num=int(round((x_max - x_min) / delta)))
    Y0 = np.linspace(y_min + delta / 2, y_max - delta / 2, num=int(round((y_max - y_min) / delta)))
    XX0, YY0 = np.meshgrid(X0, Y0)
    return XX0.flatten(), YY0.flatten()


def barycentric_tables(src_xy, dst_xy):
    """Qhull Delaunay of ``src_xy``; enclosing simplex + barycentric weights of every ``dst_xy``
    (UTL:38-44).  Points outside the hull wrap to the last simplex like ``np.take(..., -1)``
    and get negative weights (-> NaN fill downstream, UTL:89)."""
    tri = Delaunay(src_xy)
    simplex = tri.find_simplex(dst_xy)
    vert = np.take(tri.simplices, simplex, axis=0)
    T = np.take(tri.transform, simplex, axis=0)
    d = dst_xy - T[:, 2]
    bary = np.einsum('njk,nk->nj', T[:, :2, :], d)
    wts = np.hstack((bary, 1 - bary.sum(axis=1, keepdims=True)))
    return np.ascontiguousarray(vert, dtype=np.int32), np.ascontiguousarray(wts, dtype=np.float64)


def regular_grid_back_tables(cells_xy, X0_row, Y0_col, W):
    """Grid -> cell tables in closed form (no Qhull): the source points are the regular grid, so the
    enclosing square is a floor() and the two triangles of a square follow a fixed diagonal
    ((gi,gj)-(gi+1,gj+1)).  Every triangulation of a regular grid is Delaunay (co-circular corners),
    so this is one of the valid answers; it is NOT the particular one Qhull's degenerate-input
    handling picks for PMP:211, hence opt-in (``back='closed_form'``) for meshes where Qhull on
    millions of grid points would take minutes.  Cells outside the grid hull get the same wrapped
    'negative weight' marker the reference produces for them (-> previous-pressure fallback)."""
    x0, y0 = X0_row[0], Y0_col[0]
    dx = (X0_row[-1] - X0_row[0]) / (len(X0_row) - 1)
    dy = (Y0_col[-1] - Y0_col[0]) / (len(Y0_col) - 1)
    H = len(Y0_col)
    fx = (cells_xy[:, 0] - x0) / dx
    fy = (cells_xy[:, 1] - y0) / dy
    outside = (fx < 0) | (fy < 0) | (fx > W - 1) | (fy > H - 1)
    gj = np.clip(np.floor(fx).astype(np.int64), 0, W - 2)
    gi = np.clip(np.floor(fy).astype(np.int64), 0, H - 2)
    u, v = fx - gj, fy - gi
    lower = u >= v                                   # triangle (0,0),(1,0),(1,1)  vs  (0,0),(1,1),(0,1)
    p00, p10, p11, p01 = gi * W + gj, gi * W + gj + 1, (gi + 1) * W + gj + 1, (gi + 1) * W + gj
    vert = np.where(lower[:, None], np.stack([p00, p10, p11], 1), np.stack([p00, p11, p01], 1)).astype(np.int32)
    wl = np.stack([1 - u, u - v, v], 1)
    wu = np.stack([1 - v, u, v - u], 1)
    wts = np.where(lower[:, None], wl, wu)
    wts[outside] = np.array([-1.0, 1.0, 1.0])
    return np.ascontiguousarray(vert), np.ascontiguousarray(wts, dtype=np.float64)


def _inside_convex(hull_pts, pts):
    inside = np.ones(pts.shape[0], dtype=bool)
    n = hull_pts.shape[0]
    for k in range(n):
        a, b = hull_pts[k], hull_pts[(k + 1) % n]
        inside &= ((b[0] - a[0]) * (pts[:, 1] - a[1]) - (b[1] - a[1]) * (pts[:, 0] - a[0])) > 0
    return inside


def _nearest(pts, xy0):
    d, _ = cKDTree(pts).query(xy0, k=1)
    return d


def flow_mask_and_distance(xy0, top, obst, variant, bbox):
    """Inside-bbox and not inside the obstacle's convex hull; distance to the nearest
    (sub-sampled) wall point, zero outside the flow (SMC:117-143, GRAD:190-214, PMP:72-99)."""
    x_min, x_max, y_min, y_max = bbox
    if variant == 'deltaU_to_deltaP':
        max_x, max_y = max(top[:, 0].max(), x_max), min(top[:, 1].max(), y_max)      # SMC:120
        min_x, min_y = max(top[:, 0].min(), x_min), min(top[:, 1].min(), y_min)      # SMC:121
        step = 5
    else:
        max_x, max_y, min_x, min_y = top[:, 0].max(), top[:, 1].max(), top[:, 0].min(), top[:, 1].min()
        step = 2 if variant == 'U_to_gradP' else 10
    inside = (xy0[:, 0] <= max_x) & (xy0[:, 0] >= min_x) & (xy0[:, 1] <= max_y) & (xy0[:, 1] >= min_y)
    hull = np.asarray(obst, dtype=np.float64)[ConvexHull(obst).vertices]
    domain = inside & ~_inside_convex(hull, xy0)
    sdf = np.minimum(_nearest(obst[::step], xy0), _nearest(top[::step], xy0)) * domain
    return domain, sdf


def build_tables(cells_xy, top, obst, probe_values, variant='deltaU_to_deltaP', delta=5e-3,
                 back=True, precomputed=None):
    """Everything ``psm_init_with_tables`` needs, as a dict of contiguous arrays.

    ``probe_values`` [n_cells]: the field whose interpolation decides pixel validity (SMC:165-169
    uses p, PMP:230-231 uses Ux, GRAD:235-236 uses dP/dx).  ``precomputed`` optionally supplies
    (vert, weights[, vert_back, weights_back]) from another builder."""
    assert variant in VARIANTS
    cells_xy = np.ascontiguousarray(cells_xy, dtype=np.float64)
    top = np.asarray(top, dtype=np.float64)
    obst = np.asarray(obst, dtype=np.float64)
    nd = 3 if variant == 'deltaU_to_deltaP' else 2                # SMC:102-106 vs GRAD:174-178 / PMP:197-201
    x_min, x_max = round(float(cells_xy[:, 0].min()), nd), round(float(cells_xy[:, 0].max()), nd)
    y_min, y_max = round(float(cells_xy[:, 1].min()), nd), round(float(cells_xy[:, 1].max()), nd)
    X0, Y0 = uniform_grid(x_min, x_max, y_min, y_max, delta)
    xy0 = np.stack([X0, Y0], axis=1)
    H = int(round((y_max - y_min) / delta))
    W = int(round((x_max - x_min) / delta))
    if precomputed is None:
        vert, weights = barycentric_tables(cells_xy, xy0)
        vert_back = weights_back = None
        if back == 'closed_form':
            vert_back, weights_back = regular_grid_back_tables(cells_xy, X0[:W], Y0[::W], W)
        elif back:
            vert_back, weights_back = barycentric_tables(xy0, cells_xy)          # PMP:211
    else:
        vert, weights = precomputed[0], precomputed[1]
        vert_back, weights_back = (precomputed[2], precomputed[3]) if len(precomputed) > 2 else (None, None)
    domain, sdf = flow_mask_and_distance(xy0, top, obst, variant, (x_min, x_max, y_min, y_max))
    probe = np.einsum('nj,nj->n', np.take(np.asarray(probe_values, dtype=np.float64), vert), weights)
    probe[np.any(weights < 0, axis=1)] = np.nan
    ok = domain & ~np.isnan(probe)
    jj = np.rint((X0 - X0.min()) / delta).astype(np.int64)        # SMC:170-171 int(round(.))
    ii = np.rint((Y0 - Y0.min()) / delta).astype(np.int64)
    indices = np.zeros((X0.shape[0], 2), dtype=np.int64)          # invalid points stay (0,0), SMC:161
    indices[ok, 0] = ii[ok]
    indices[ok, 1] = jj[ok]
    sdfunct = np.zeros((H, W), dtype=np.float64)
    sdfunct[ii[ok], jj[ok]] = sdf[ok]
    return dict(n_cells=cells_xy.shape[0], H=H, W=W, vert=np.ascontiguousarray(vert, dtype=np.int32),
                weights=np.ascontiguousarray(weights, dtype=np.float64),
                vert_back=None if vert_back is None else np.ascontiguousarray(vert_back, dtype=np.int32),
                weights_back=None if weights_back is None else np.ascontiguousarray(weights_back, dtype=np.float64),
                indices=indices, sdfunct=sdfunct, bbox=(x_min, x_max, y_min, y_max), delta=delta)
