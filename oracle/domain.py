"""Oracle: flow-domain mask, distance field and index raster (test infrastructure).

Follows SMC:100-178 (Improved_SM ``computeOnlyOnce``), GRAD:171-248 and PMP:72-99,
216-243 (solver-side ``domain_dist`` + raster loop).
"""
import numpy as np
from scipy.spatial import ConvexHull
from scipy.spatial import distance


def convex_hull_points(obst):
    """SMC:128-133 -- ``MultiPoint(obst).convex_hull.exterior`` restated with Qhull."""
    hull = ConvexHull(np.asarray(obst, dtype=np.float64))
    return np.asarray(obst, dtype=np.float64)[hull.vertices]      # counter-clockwise in 2-D


def contains_points_convex(hull_pts, pts):
    """SMC:135-136 -- ``mpltPath.Path(hull_pts).contains_points(pts)`` for a convex polygon.

    Strict interior test through the sign of the edge cross products (CCW polygon);
    differs from matplotlib only on the polygon boundary (measure zero).
    """
    pts = np.asarray(pts, dtype=np.float64)
    inside = np.ones(pts.shape[0], dtype=bool)
    n = hull_pts.shape[0]
    for k in range(n):
        a = hull_pts[k]
        b = hull_pts[(k + 1) % n]
        cross = (b[0] - a[0]) * (pts[:, 1] - a[1]) - (b[1] - a[1]) * (pts[:, 0] - a[0])
        inside &= cross > 0
    return inside


def min_cdist(xy0, pts, chunk=65536):
    """``distance.cdist(xy0, pts).min(axis=1)`` (SMC:143) evaluated in row chunks
    (same arithmetic, bounded memory)."""
    out = np.empty(xy0.shape[0], dtype=np.float64)
    for s in range(0, xy0.shape[0], chunk):
        out[s:s + chunk] = distance.cdist(xy0[s:s + chunk], pts).min(axis=1)
    return out


def domain_dist(xy0, top, obst, variant, x_min=None, x_max=None, y_min=None, y_max=None):
    """Mask + distance field.

    variant 'smc' : SMC:117-143 -- bbox test with the reference's max/min mix against the
                    rounded cell bbox (SMC:120-121), boundary sub-sampling ``[::5]``.
    variant 'grad': GRAD:190-214 -- bbox of ``top`` only, sub-sampling ``[::2]``.
    variant 'pmp' : PMP:72-99  -- bbox of ``top`` only, sub-sampling ``[::10]``.
    Returns (domain_bool bool[M], sdf float64[M]).
    """
    if variant == 'smc':
        max_x, max_y = np.max([(top[:, 0]).max(), x_max]), np.min([(top[:, 1]).max(), y_max])
        min_x, min_y = np.max([(top[:, 0]).min(), x_min]), np.min([(top[:, 1]).min(), y_min])
        step = 5
    elif variant in ('grad', 'pmp'):
        max_x, max_y, min_x, min_y = np.max(top[:, 0]), np.max(top[:, 1]), np.min(top[:, 0]), np.min(top[:, 1])
        step = 2 if variant == 'grad' else 10
    else:
        raise ValueError(variant)
    is_inside_domain = (xy0[:, 0] <= max_x) * (xy0[:, 0] >= min_x) * (xy0[:, 1] <= max_y) * (xy0[:, 1] >= min_y)
    hull_pts = convex_hull_points(obst)
    is_inside_obst = contains_points_convex(hull_pts, xy0)
    domain_bool = is_inside_domain * ~is_inside_obst
    top_s = top[0:top.shape[0]:step, :]
    obst_s = obst[0:obst.shape[0]:step, :]
    sdf = np.minimum(min_cdist(xy0, obst_s), min_cdist(xy0, top_s)) * domain_bool
    return domain_bool, sdf, (min_x, max_x, min_y, max_y)


def index_raster(X0, Y0, delta, grid_shape_y, grid_shape_x, domain_bool, probe_interp, sdf, literal=False):
    """SMC:156-178 / PMP:220-243 / GRAD:226-248.

    For every grid point that is inside the domain and whose probe interpolation is
    not NaN: ``indices[k] = (ii, jj)`` and ``sdfunct[ii, jj] = sdf[k]``; every other
    point keeps ``indices == (0, 0)`` (``np.zeros`` at SMC:161).
    ``literal=True`` runs the reference's Python loop; the default is the vectorised
    equivalent (``np.rint`` == ``int(round(.))``: both round half to even).
    Returns (indices int64[M,2], sdfunct float64[H,W,1]).
    """
    x0 = np.min(X0)
    y0 = np.min(Y0)
    dx = dy = delta
    indices = np.zeros((X0.shape[0], 2))
    sdfunct = np.zeros((grid_shape_y, grid_shape_x, 1))
    if literal:
        xy0 = np.c_[X0, Y0]
        for (step, x_y) in enumerate(xy0):
            if domain_bool[step] * (~np.isnan(probe_interp[step])):
                jj = int(round((x_y[..., 0] - x0) / dx))
                ii = int(round((x_y[..., 1] - y0) / dy))
                indices[step, 0] = ii
                indices[step, 1] = jj
                sdfunct[ii, jj, :] = sdf[step]
    else:
        ok = domain_bool.astype(bool) & ~np.isnan(probe_interp)
        jj = np.rint((X0 - x0) / dx).astype(np.int64)
        ii = np.rint((Y0 - y0) / dy).astype(np.int64)
        indices[ok, 0] = ii[ok]
        indices[ok, 1] = jj[ok]
        sdfunct[ii[ok], jj[ok], 0] = sdf[ok]
    return indices.astype(int), sdfunct
