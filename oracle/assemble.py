"""Oracle: block re-assembly with the sequential mean-offset correction (test infrastructure).

``assemble_deltas`` follows SMC:182-350 (Improved_SM deltaU_to_deltaP),
``assemble_gradp`` follows GRAD:255-361 (U_to_gradP, one call per output field).
Both walk the blocks in extraction order, subtract a scalar correction from each
block, and overwrite the block's rectangle in the result (later blocks win).
The optional Gaussian filter (SMC:352-356) and deltaU-change weighting
(SMC:358-363) are off in the hot path (EP:105, SMC:573) and not restated here.
"""
import warnings
import numpy as np


def _mmean(vals, mask):
    """``np.mean(vals[mask != 0])`` -- NaN for an empty selection, as NumPy does."""
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", category=RuntimeWarning)
        return np.mean(vals[mask != 0])


def _place_deltas(result, pf, idx_i, idx_j, n_x, n_y, shape, overlap, shape_x, shape_y, p_i):
    """Placement step, SMC:332-348 (later blocks overwrite earlier ones)."""
    st = shape - overlap
    if [idx_i, idx_j] == [n_y + 1, 0]:
        result[-p_i:shape_y, 0:shape] = pf[-p_i:]
    elif idx_j == 0:
        result[st * idx_i:st * idx_i + shape, 0:shape] = pf
    elif idx_i == (n_y + 1):
        jj = n_x - idx_j
        result[-p_i:, shape_x - shape - jj * st:shape_x - jj * st] = pf[-p_i:]
    else:
        jj = n_x - idx_j
        result[st * idx_i:st * idx_i + shape, shape_x - shape - jj * st:shape_x - jj * st] = pf


def _place_gradp(result, pf, idx_i, idx_j, n_x, n_y, shape, avance, shape_y, izl):
    """Placement step, GRAD:345-356."""
    st = shape - avance
    if [idx_i, idx_j] == [n_y + 1, n_x]:
        result[0, (shape_y - st):shape_y, -izl:, 0] = pf[avance:shape, -izl:]
    elif idx_j == n_x:
        result[0, idx_i * st:idx_i * st + shape, -izl:, 0] = pf[:, -izl:]
    elif idx_i == (n_y + 1):
        result[0, (shape_y - st):shape_y, idx_j * st:shape + idx_j * st, 0] = pf[avance:shape, :]
    else:
        result[0, idx_i * st:idx_i * st + shape, idx_j * st:shape + idx_j * st, 0] = pf


def owner_map(variant, indices_list, n_x, n_y, shape, overlap, shape_x, shape_y):
    """Which block's value survives at every pixel: the placement step replayed with block ids."""
    if variant == 'deltas':
        result = np.full((shape_y, shape_x), -1.0)
        p_i = shape_y - ((shape - overlap) * n_y + shape)
        for k, (idx_i, idx_j) in enumerate(indices_list):
            _place_deltas(result, np.full((shape, shape), float(k)), idx_i, idx_j, n_x, n_y, shape, overlap,
                          shape_x, shape_y, p_i)
        return result.astype(np.int64)
    result = np.full((1, shape_y, shape_x, 1), -1.0)
    p_j = (shape_x - shape) - n_x * (shape - overlap)
    for k, (idx_i, idx_j) in enumerate(indices_list):
        _place_gradp(result, np.full((shape, shape), float(k)), idx_i, idx_j, n_x, n_y, shape, overlap, shape_y,
                     overlap - p_j)
    return result[0, :, :, 0].astype(np.int64)


def assemble_deltas(array, x_array, indices_list, n_x, n_y, shape, overlap, shape_x, shape_y,
                    Ref_BC=0.0, return_offsets=False):
    """SMC:182-350.  ``array`` [B,S,S] predicted blocks (NOT modified: a copy is corrected),
    ``x_array`` [B,S,S,3] extracted inputs (channel 2 = distance field -> flow mask)."""
    array = np.array(array, dtype=np.float64, copy=True)
    result = np.empty(shape=(shape_y, shape_x))
    BC_ups = np.zeros(n_x + 1)
    p_i = shape_y - ((shape - overlap) * n_y + shape)          # SMC:213
    p_j = shape_x - ((shape - overlap) * n_x + shape)          # SMC:216
    offsets = np.zeros(array.shape[0])
    old = None
    for k in range(x_array.shape[0]):
        idx_i, idx_j = indices_list[k]
        fb = x_array[k, :, :, 2]
        pf = array[k, ...]

        def side(width):
            # SMC:235-236 / 239-240: previous (already corrected) block's LEFT strip under
            # THIS block's left-strip mask, against this block's right strip.
            ant = _mmean(old[:, :width], fb[:, :width])
            return _mmean(pf[:, -width:], fb[:, -width:]) - ant

        if idx_i == 0:                                          # first row, SMC:228-246
            if k == 0:
                c = _mmean(pf[:, -1], fb[:, -1]) - Ref_BC
            else:
                c = side(overlap)
            if idx_j == 0:
                c = side(overlap - p_j)
            pf -= c
            BC_ups[idx_j] = _mmean(pf[-overlap:, :], fb[-overlap:, :])
        elif idx_i != n_y + 1:                                  # middle rows, SMC:249-283
            if np.isnan(BC_ups[idx_j]):
                if idx_j == 0:
                    c = side(overlap - p_j)
                elif idx_j == n_x:
                    c = _mmean(pf[:overlap, :], fb[:overlap, :]) - BC_ups[idx_j]
                else:
                    c = side(overlap)
            else:
                c = _mmean(pf[:overlap, :], fb[:overlap, :]) - BC_ups[idx_j]
            pf -= c
            BC_ups[idx_j] = _mmean(pf[-overlap:, :], fb[-overlap:, :])
            if idx_i == n_y:
                BC_ups[idx_j] = _mmean(pf[-(shape - p_i):, :], fb[-(shape - p_i):, :])
        else:                                                   # last row, SMC:286-328
            if idx_j == n_x:
                c = _mmean(pf[-p_i - overlap:-p_i, :], fb[-p_i - overlap:-p_i, :]) - BC_ups[idx_j]
            else:
                n_up_non_nans = (fb[-p_i - overlap:-p_i, :] != 0).sum()
                if (n_up_non_nans) / 128 ** 2 > 0.9:            # SMC:307 (unreachable for overlap <= 115)
                    c = side(overlap - p_j) if idx_j == 0 else side(overlap)
                else:
                    c = _mmean(pf[:-p_i, :], fb[:-p_i, :]) - BC_ups[idx_j]
            pf -= c
        offsets[k] = c
        old = pf

        _place_deltas(result, pf, idx_i, idx_j, n_x, n_y, shape, overlap, shape_x, shape_y, p_i)

    shift = np.mean(3 * result[:, -1] - result[:, -2]) / 3        # SMC:350
    result -= shift
    if return_offsets:
        return result, offsets, shift
    return result


def assemble_gradp(field, array, x_array, indices_list, n_x, n_y, shape, avance, shape_x, shape_y,
                   Ref_BC=0.0, return_offsets=False):
    """GRAD:255-361.  ``field`` is 'dp_dx' or 'dp_dy'; blocks ordered left->right, top->bottom."""
    array = np.array(array, dtype=np.float64, copy=True)
    result = np.empty(shape=(1, shape_y, shape_x, 1))
    BC_ups = np.zeros(n_x + 1)
    p_i = shape_y - (shape * (n_y + 1) - n_y * avance)          # GRAD:277
    p_j = (shape_x - shape) - n_x * (shape - avance)            # GRAD:278
    offsets = np.zeros(array.shape[0])
    old = None
    intersect_zone_limit = None
    for k in range(x_array.shape[0]):
        idx_i, idx_j = indices_list[k]
        fb = x_array[k, :, :, 2]
        pf = array[k, ...]

        def side(width):
            # GRAD:305-306 / 309-310: previous block's RIGHT strip under THIS block's
            # right-strip mask, against this block's left strip.
            ant = _mmean(old[:, -width:], fb[:, -width:])
            return _mmean(pf[:, :width], fb[:, :width]) - ant

        if idx_i == 0:                                          # GRAD:288-312
            if k == 0:
                if field == 'dp_dx':
                    col = 0
                    while (fb[:, col] != 0).sum() == 0:
                        col += 1
                        assert col < shape
                    c = _mmean(pf[:, col], fb[:, col]) - Ref_BC
                elif field == 'dp_dy':
                    c = _mmean(pf[1, :], fb[1, :]) - Ref_BC
                else:
                    raise ValueError(field)
            else:
                c = side(avance)
            if idx_j == n_x:
                intersect_zone_limit = avance - p_j
                c = side(intersect_zone_limit)
            pf -= c
            BC_ups[idx_j] = _mmean(pf[-avance:, :], fb[-avance:, :])
        elif idx_i != n_y + 1:                                  # GRAD:314-328
            if np.isnan(BC_ups[idx_j]):
                if idx_j == n_x:
                    intersect_zone_limit = avance - p_j
                    c = side(intersect_zone_limit)
                else:
                    c = side(avance)
            else:
                c = _mmean(pf[:avance, :], fb[:avance, :]) - BC_ups[idx_j]
            pf -= c
            BC_ups[idx_j] = _mmean(pf[-avance:, :], fb[-avance:, :])
            if idx_i == n_y:
                BC_ups[idx_j] = _mmean(pf[-(shape - p_i):, :], fb[-(shape - p_i):, :])
        else:                                                   # GRAD:330-341
            if np.isnan(BC_ups[idx_j]):
                if idx_j == n_x:
                    intersect_zone_limit = avance - p_j
                    c = side(intersect_zone_limit)
                else:
                    c = side(avance)
            else:
                c = _mmean(pf[-p_i - avance:-p_i, :], fb[-p_i - avance:-p_i, :]) - BC_ups[idx_j]
            pf -= c
        offsets[k] = c
        old = pf

        _place_gradp(result, pf, idx_i, idx_j, n_x, n_y, shape, avance, shape_y, intersect_zone_limit)

    if field == 'dp_dx':                                        # GRAD:358-361
        shift = np.mean(3 * result[:, :, 0, :] - result[:, :, 1, :]) / 3
    else:
        shift = np.mean(3 * result[:, 1, :, :] - result[:, 2, :, :]) / 3
    result -= shift
    if return_offsets:
        return result, offsets, shift
    return result


def assemble_thesis(array, x_array, indices_list, n_x, n_y, shape, avance, shape_x, shape_y, return_offsets=False):
    """PMP:372-473 (thesis solver module): blocks in extraction order -- right -> left, plus the extra left-most
    column tagged ``-1`` -- each corrected by a scalar and written into the result (later blocks win), then the
    global shift PMP:472.  ``array`` [B,S,S] is corrected on a copy.  Every quirk is kept: the right-strip
    correction chain of row 0 (``BC_ant_0``), the ``-1`` column chained through ``BC_up_`` (whose middle-row
    update is an UNMASKED mean, PMP:437), the ``BC_alter`` fallback when ``BC_ups`` is NaN."""
    array = np.array(array, dtype=np.float64, copy=True)
    result = np.empty(shape=(shape_y, shape_x))
    BC_up = 0
    BC_alter = 0
    BC_ups = np.zeros(n_x + 1)
    offsets = np.zeros(array.shape[0])
    p_j = (shape_x - shape) - n_x * shape + n_x * avance
    p = shape_y - (shape * (n_y + 1) - n_y * avance)
    for i in range(array.shape[0]):
        idx = list(indices_list[i])
        flow_bool = x_array[i, :, :, 2]
        res = array[i]
        if idx[0] == 0:
            if idx[1] == n_x:
                BC_coor = _mmean(res[:, (shape - avance):shape], flow_bool[:, (shape - avance):shape]) - BC_up
                res -= BC_coor
                BC_ups[idx[1]] = _mmean(res[(shape - avance):shape, (shape - avance):shape],
                                        flow_bool[(shape - avance):shape, (shape - avance):shape])
            elif idx[1] == -1:
                BC_coor = _mmean(res[:, p_j:p_j + avance], flow_bool[:, p_j:p_j + avance]) - BC_ant_0
                res -= BC_coor
                BC_up_ = _mmean(res[(shape - avance):shape, p_j:p_j + avance], flow_bool[(shape - avance):shape, p_j:p_j + avance])
            else:
                BC_coor = _mmean(res[:, (shape - avance):shape], flow_bool[:, (shape - avance):shape]) - BC_ant_0
                res -= BC_coor
                BC_ups[idx[1]] = _mmean(res[(shape - avance):shape, :], flow_bool[(shape - avance):shape, :])
            BC_ant_0 = _mmean(res[:, 0:avance], flow_bool[:, 0:avance])
        elif idx[0] == n_y + 1:
            if idx[1] == -1:
                BC_coor = _mmean(res[shape - p - avance:shape - p, p_j:p_j + avance],
                                 flow_bool[shape - p - avance:shape - p, p_j:p_j + avance]) - BC_up_
                res -= BC_coor
            else:
                if np.isnan(BC_ups[idx[1]]):
                    BC_coor = _mmean(res[:, shape - avance:shape], flow_bool[:, shape - avance:shape]) - BC_alter
                else:
                    BC_coor = _mmean(res[shape - p - avance:shape - p, :], flow_bool[shape - p - avance:shape - p, :]) - BC_ups[idx[1]]
                res -= BC_coor
        else:
            if idx[1] == -1:
                BC_coor = _mmean(res[0:avance, p_j:p_j + avance], flow_bool[0:avance, p_j:p_j + avance]) - BC_up_
                res -= BC_coor
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore", category=RuntimeWarning)
                    BC_up_ = np.mean(res[(shape - avance):shape, p_j:p_j + avance])          # PMP:437 -- no mask
            else:
                if np.isnan(BC_ups[idx[1]]):
                    BC_coor = _mmean(res[:, shape - avance:shape], flow_bool[:, shape - avance:shape]) - BC_alter
                else:
                    BC_coor = _mmean(res[0:avance, :], flow_bool[0:avance, :]) - BC_ups[idx[1]]
                res -= BC_coor
                BC_ups[idx[1]] = _mmean(res[(shape - avance):shape, :], flow_bool[(shape - avance):shape, :])
        offsets[i] = BC_coor
        BC_alter = _mmean(res[:, 0:avance], flow_bool[:, 0:avance])                       # PMP:449
        st = shape - avance
        if idx == [n_y + 1, -1]:
            wdt = shape_x - (n_x + 1) * st - avance
            result[(shape_y - st):shape_y, 0:wdt] = res[avance:shape, 0:wdt]
        elif idx[1] == -1:
            result[idx[0] * st:idx[0] * st + shape, 0:shape] = res
        elif idx[0] == n_y + 1:
            j = n_x - idx[1]
            result[(shape_y - st):shape_y, shape_x - shape - j * st:shape_x - j * st] = res[avance:shape, :]
        else:
            j = n_x - idx[1]
            result[idx[0] * st:idx[0] * st + shape, shape_x - shape - j * st:shape_x - j * st] = res
    shift = np.mean(3 * result[:, -1] - result[:, -2]) / 3                             # PMP:472
    result -= shift
    if return_offsets:
        return result, offsets, shift
    return result
