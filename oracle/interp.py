"""Oracle: uniform grid + barycentric interpolation tables (test infrastructure).

Follows UTL:22-55 (interp_weights), UTL:75-90 (interpolate_fill),
UTL:111-125 (create_uniform_grid) and the solver-side twins PMP:42-70,154-161.
The triangulation is scipy's Qhull binding -- the same library the reference
calls (``scipy.spatial.qhull.Delaunay`` is an alias of ``scipy.spatial.Delaunay``).
"""
import numpy as np
from scipy.spatial import Delaunay, cKDTree


def create_uniform_grid(x_min, x_max, y_min, y_max, delta):
    """UTL:111-125 / PMP:42-48 -- cell-centred uniform grid, flattened row-major (y outer)."""
    X0 = np.linspace(x_min + delta / 2, x_max - delta / 2, num=int(round((x_max - x_min) / delta)))
    Y0 = np.linspace(y_min + delta / 2, y_max - delta / 2, num=int(round((y_max - y_min) / delta)))
    XX0, YY0 = np.meshgrid(X0, Y0)
    return XX0.flatten(), YY0.flatten()


def interp_weights(xyz, uvw, d=2, idw_fallback=False):
    """UTL:22-55 (idw_fallback=True) / PMP:154-161, GRAD:69-84 (idw_fallback=False).

    Returns (vertices int32[M,3], wts float64[M,3]).  ``simplex == -1`` wraps to the
    LAST simplex through ``np.take`` (negative weights result), exactly as the
    reference.  The IDW branch of UTL:47-53 references the unimported name
    ``sklearn`` (latent NameError in the reference); when ``idw_fallback`` is set it
    is restated with a KD-tree so the branch can be exercised -- k=3 neighbours,
    weights 1/max(d^2, 1e-6), normalised.
    """
    tri = Delaunay(xyz)
    simplex = tri.find_simplex(uvw)
    vertices = np.take(tri.simplices, simplex, axis=0)
    temp = np.take(tri.transform, simplex, axis=0)
    delta = uvw - temp[:, d]
    bary = np.einsum('njk,nk->nj', temp[:, :d, :], delta)
    wts = np.hstack((bary, 1 - bary.sum(axis=1, keepdims=True)))
    if idw_fallback:
        valid = ~(simplex == -1)
        if (~valid).any():
            tree = cKDTree(xyz)
            nndist, nni = tree.query(np.array(uvw)[~valid], k=3)
            invalid = np.flatnonzero(~valid)
            vertices[invalid] = nni
            inv = 1. / np.maximum(nndist ** 2, 1e-6)
            wts[invalid] = inv / inv.sum(axis=-1)[:, None]
    return vertices, wts


def interpolate(values, vtx, wts):
    """PMP:64-65 -- no NaN fill (garbage, not NaN, for outside points)."""
    return np.einsum('nj,nj->n', np.take(values, vtx), wts)


def interpolate_fill(values, vtx, wts, fill_value=np.nan):
    """UTL:75-90 / PMP:67-70 -- NaN where any weight is negative."""
    ret = np.einsum('nj,nj->n', np.take(values, vtx), wts)
    ret[np.any(wts < 0, axis=1)] = fill_value
    return ret
