"""Oracle: pressure recovery from the assembled gradient field (test infrastructure).

Restates ``Evaluation.integrate_field`` (GRAD:371-416) and the four-quadrant stitch of ``timeStep`` (GRAD:585-628) of
``Improved_SM/U_to_gradP/evaluation/Eval_dual_Dense_onlycil.py``, quirks included:

  * the "reset at the obstacle" (GRAD:389-395) indexes with ``sdfunct[i, :, 0].astype(int)`` -- the distance field truncated
    to an integer, i.e. an index array of 0s (1s, 2s where a pixel is >= 1 m, 2 m from every wall), taken from grid row ``i``
    = the row index INSIDE the block (also for the two lower quadrants), over the full grid width whatever the block width;
  * ``j *= direction_x`` (GRAD:408-409) only moves the anchor column / row: every pixel receives
    ``SdPy[i, j0] - SdPy[i0, j0] + SdPx[i, j] - SdPx[i, j0]`` with ``j0`` = first (or last) column, ``i0`` = first (or last) row;
  * the stitch compresses two neighbouring columns with two DIFFERENT row masks and subtracts them element by element
    (GRAD:606, 620); the centre row is hard-coded to 200 (GRAD:592).
Pinned against outputs of the unmodified reference: ``tests/golden/grad_integrate.npz`` (``tests/golden/make_golden.py``).
"""
import numpy as np


def integrate_field_literal(block, sdfunct, xl, yl, direction_x=1, direction_y=1):
    """GRAD:371-416, line by line (the pixel loop GRAD:405-414 included): small inputs only."""
    to_cumsum_dPdx = block[..., 0].copy()
    to_cumsum_dPdy = block[..., 1].copy()
    ll = []
    for i in range(to_cumsum_dPdx.shape[0]):
        aaa = to_cumsum_dPdx[i, :].copy()
        ccc = np.cumsum(aaa)
        nn = sdfunct[i, :, 0].astype(int)
        dd = np.diff(np.concatenate(([0.], ccc[nn])))
        aaa[nn] = -dd
        SdPx = np.cumsum(aaa) * np.diff(xl)[0]
        ll.append(SdPx.reshape((1, SdPx.shape[0])))
    SdPx = np.concatenate(ll, axis=0)
    SdPy = np.cumsum(to_cumsum_dPdy, axis=0) * np.diff(yl)[0]
    Phat = np.zeros(SdPx.shape)
    for i in range(Phat.shape[0]):
        for j in range(Phat.shape[1]):
            j *= direction_x
            i *= direction_y
            initial_j = 0
            initial_i = 0
            if direction_x == -1:
                initial_j = -1
            if direction_y == -1:
                initial_i = -1
            Phat[i, j] += np.sum([SdPy[i, initial_j], -SdPy[initial_i, initial_j], SdPx[i, j], -SdPx[i, initial_j]])
    return Phat


def integrate_field(block, sdfunct, xl, yl, direction_x=1, direction_y=1):
    """GRAD:371-416 with the pixel loop vectorised (same sums in the same order: np.sum of four terms, left to right)."""
    dPdx = block[..., 0].copy()
    dPdy = block[..., 1].copy()
    dx, dy = np.diff(xl)[0], np.diff(yl)[0]
    SdPx = np.empty(dPdx.shape)
    for i in range(dPdx.shape[0]):
        aaa = dPdx[i, :].copy()
        ccc = np.cumsum(aaa)
        nn = sdfunct[i, :, 0].astype(int)
        dd = np.diff(np.concatenate(([0.], ccc[nn])))
        aaa[nn] = -dd
        SdPx[i] = np.cumsum(aaa) * dx
    SdPy = np.cumsum(dPdy, axis=0) * dy
    j0 = -1 if direction_x == -1 else 0
    i0 = -1 if direction_y == -1 else 0
    return ((SdPy[:, j0][:, None] + (-SdPy[i0, j0])) + SdPx) + (-SdPx[:, j0][:, None])


def recover_pressure(dp_dx, dp_dy, sdfunct, min_x, max_x, min_y, max_y, x0_min, delta, center_row=200):
    """GRAD:585-628: the pressure field from the two assembled gradient fields [H, W]; ``sdfunct`` [H, W, 1]."""
    H, W = dp_dx.shape
    gradP = np.stack([dp_dx, dp_dy], axis=-1)[None]
    xl = np.linspace(min_x, max_x, W)
    yl = np.linspace(min_y, max_y, H)
    zero = sdfunct[center_row, :, 0] == 0
    center_p_x = int(((xl[zero].max() + xl[zero].min()) / 2 - x0_min) / delta)
    center_p_y = int(center_row)
    result = np.empty((H, W))
    block_1 = gradP[0, :center_p_y, center_p_x - 1:, :].copy()
    mask1 = sdfunct[:center_p_y, center_p_x - 1, 0] != 0
    pBlock1 = integrate_field(block_1, sdfunct, xl, yl, direction_x=-1)
    result[:center_p_y, center_p_x - 1:] = pBlock1
    block_2 = gradP[0, :center_p_y, :center_p_x, :].copy()
    mask2 = sdfunct[:center_p_y, center_p_x, 0] != 0
    pBlock2 = integrate_field(block_2, sdfunct, xl, yl)
    result[:center_p_y, :center_p_x] = pBlock2 - (pBlock2[:, -1][mask2] - pBlock1[:, 0][mask1]).mean()
    block_3 = gradP[0, center_p_y:, center_p_x - 1:, :].copy()
    mask3 = sdfunct[center_p_y:, center_p_x - 1, 0] != 0
    pBlock3 = integrate_field(block_3, sdfunct, xl, yl, direction_x=-1, direction_y=-1)
    result[center_p_y:, center_p_x - 1:] = pBlock3
    block_4 = gradP[0, center_p_y:, :center_p_x, :].copy()
    mask4 = sdfunct[center_p_y:, center_p_x, 0] != 0
    pBlock4 = integrate_field(block_4, sdfunct, xl, yl, direction_y=-1)
    result[center_p_y:, :center_p_x] = pBlock4 - (pBlock4[:, -1][mask4] - pBlock3[:, 0][mask3]).mean()
    return result, center_p_x
