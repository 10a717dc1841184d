"""Block-row partitioner for the multi-GPU path (``psm_init_sharded``).

The reference gathers every MPI rank's cells to rank 0, computes there, and scatters the pressures
(PMP:179-185, 258, 501-511).  Here the grid is cut by BLOCK ROWS of the extraction plan (SMC:461-479 /
GRAD:479-500): rank g gathers and places pixel rows ``[row0, row1)``, evaluates the blocks of its block
rows, and owns the cells whose centre falls into its pixel rows.  What crosses ranks per step:

  1. max|U|^2 (all-reduce) and the *ghost cells* a rank's forward table references but does not own;
  2. the ``overlap`` grid rows below ``row1`` that complete the rank's last block row (halo strip);
  3. the masked strip means / shift-line sums (all-reduce of disjoint slots);
  4. the *ghost pixels* of the assembled field a rank's grid->cell table references (boundary rows,
     and pixel (0,0) for the raster quirk of PMP:481).

Everything here is static per mesh and runs once in ``init``.  Host-side NumPy only; no CUDA.
"""
import numpy as np


def fold_forward_table(vert, weights, indices, H, W):
    """Forward table after the validity fold: what pixel q of the grid receives.

    ``grid[...][tuple(indices.T)] = interp`` (SMC:432-434): for duplicate targets the LAST source
    point wins, so pixel (0,0) takes the entry of the last invalid point; pixels no point maps to
    stay 0; entries with a negative weight are NaN (UTL:89) -> 0 (SMC:438)."""
    G = H * W
    src = np.full(G, -1, dtype=np.int64)
    flat = indices[:, 0] * W + indices[:, 1]
    src[flat] = np.arange(G)
    fv = np.zeros((G, 3), np.int32)
    fw = np.zeros((G, 3), np.float64)
    ok = src >= 0
    fv[ok] = vert[src[ok]]
    ww = weights[src[ok]].copy()
    ww[np.any(ww < 0, axis=1)] = 0.0
    fw[ok] = ww
    return fv, fw


def hop_back_table(vert_back, weights_back, indices, sdfunct, W, near_wall_sdf=0.0):
    """Grid->cell table with the ``indices`` hop folded in (PMP:481) and the keep-p_prev mask:
    NaN rule (negative weight, PMP:496) and near-wall rule (PMP:492-494)."""
    flat = indices[:, 0] * W + indices[:, 1]
    bv = flat[vert_back].astype(np.int64)
    keep = np.any(weights_back < 0, axis=1)
    if near_wall_sdf and near_wall_sdf > 0:
        sdf_mesh = np.einsum('nj,nj->n', np.take(sdfunct.reshape(-1), vert_back), weights_back)
        keep |= (~keep) & (sdf_mesh < near_wall_sdf)
    return bv, keep


def block_row_split(H, shape, overlap, world):
    """Contiguous block-row ranges per rank.  Regular rows 0..n_y are split evenly; the thick last
    row (y0 = H - shape, SMC:471-472) stays with the last rank.  Returns (ranges, pixel_rows)."""
    stride = shape - overlap
    n_y = (H - shape) // stride
    n_reg = n_y + 1
    need = -(-overlap // stride)                 # rows of the next rank an overlap strip can reach into
    if world > n_reg or (world > 1 and n_reg // world < need):
        raise ValueError('%d block rows cannot be split over %d ranks (each needs >= %d)' % (n_reg, world, need))
    cuts = [round(g * n_reg / world) for g in range(world + 1)]
    ranges, rows = [], []
    for g in range(world):
        b0, b1 = cuts[g], cuts[g + 1]
        last = g == world - 1
        ranges.append((b0, b1 + 1 if last else b1))
        rows.append((b0 * stride, H if last else b1 * stride))
    return ranges, rows


def _csr(counts):
    p = np.zeros(len(counts) + 1, np.int64)
    p[1:] = np.cumsum(counts)
    return p


def partition(tables, cells_xy, world, variant='deltaU_to_deltaP', shape=128, overlap=None, near_wall_sdf=0.0):
    """Split GLOBAL tables (``psm_b200.tables.build_tables``) into ``world`` shard dicts.

    Each dict holds what ``psm_shard`` needs plus ``owned_ids`` (global ids of the cells this rank
    passes to / gets from ``psm_predict``, ascending)."""
    if overlap is None:
        overlap = 32 if variant == 'deltaU_to_deltaP' else 96
    H, W = int(tables['H']), int(tables['W'])
    G = H * W
    delta = tables['delta']
    y_min = tables['bbox'][2]
    N = int(tables['n_cells'])
    sdf = np.ascontiguousarray(tables['sdfunct'], dtype=np.float64).reshape(H, W)
    mask = np.ascontiguousarray(sdf != 0.0, dtype=np.uint8)
    fv, fw = fold_forward_table(tables['vert'], tables['weights'], tables['indices'], H, W)
    have_back = tables.get('vert_back') is not None
    if have_back:
        bv, keep = hop_back_table(tables['vert_back'], tables['weights_back'], tables['indices'], sdf, W, near_wall_sdf)
    ranges, rows = block_row_split(H, shape, overlap, world)
    row_lo = np.array([r[0] for r in rows])
    pix_rank = lambda q: np.searchsorted(row_lo, q // W, side='right') - 1          # noqa: E731
    cell_row = np.clip(np.floor((np.asarray(cells_xy)[:, 1] - y_min) / delta).astype(np.int64), 0, H - 1)
    cell_rank = np.searchsorted(row_lo, cell_row, side='right') - 1

    owned = [np.flatnonzero(cell_rank == g) for g in range(world)]
    local_of = np.empty(N, np.int64)                   # global cell id -> local id on its owner
    for g in range(world):
        local_of[owned[g]] = np.arange(owned[g].size)

    shards = []
    for g in range(world):
        r0, r1 = rows[g]
        q0, q1 = r0 * W, r1 * W
        v, w = fv[q0:q1], fw[q0:q1]
        live = np.any(w != 0.0, axis=1)
        need = np.unique(v[live])
        ghost = need[cell_rank[need] != g]
        ghost = ghost[np.lexsort((ghost, cell_rank[ghost]))]            # by owner rank, then global id
        lut = np.zeros(N, np.int64)
        lut[owned[g]] = np.arange(owned[g].size)
        lut[ghost] = owned[g].size + np.arange(ghost.size)
        lv = np.where(live[:, None], lut[v], 0).astype(np.int32)
        sh = dict(rank=g, world=world, H=H, W=W, row0=r0, row1=r1,
                  ext_rows=0 if g == world - 1 else overlap, send_rows=0 if g == 0 else overlap,
                  blk_row0=ranges[g][0], blk_row1=ranges[g][1], mask_global=mask,
                  owned_ids=owned[g], ghost_ids=ghost, n_owned=int(owned[g].size), n_ghost=int(ghost.size),
                  vert=np.ascontiguousarray(lv), weights=np.ascontiguousarray(w),
                  cell_recv_ptr=_csr(np.bincount(cell_rank[ghost], minlength=world)))
        sh['sdfunct'] = np.ascontiguousarray(sdf[r0:r1 + sh['ext_rows']])
        if have_back:
            b = bv[owned[g]]
            k = keep[owned[g]]
            needp = np.unique(b[~k])
            gp = needp[(needp < q0) | (needp >= q1)]
            # pixel ids ascend with the row, hence with the owner rank: `gp` is already grouped by owner
            lb = np.empty_like(b)
            inside = (b >= q0) & (b < q1)
            lb[inside] = b[inside] - q0
            if gp.size:
                lb[~inside] = (q1 - q0) + np.clip(np.searchsorted(gp, b[~inside]), 0, gp.size - 1)
            else:
                lb[~inside] = 0
            lb[k] = 0
            lb[k, 0] = -1
            sh.update(vert_back=np.ascontiguousarray(lb.astype(np.int32)),
                      weights_back=np.ascontiguousarray(tables['weights_back'][owned[g]]),
                      ghost_pix=gp, n_ghost_pix=int(gp.size),
                      pix_recv_ptr=_csr(np.bincount(pix_rank(gp), minlength=world) if gp.size else np.zeros(world, np.int64)))
        else:
            sh.update(vert_back=None, weights_back=None, ghost_pix=np.zeros(0, np.int64), n_ghost_pix=0,
                      pix_recv_ptr=np.zeros(world + 1, np.int64))
        shards.append(sh)

    # send lists: rank p sends to g exactly g's ghosts owned by p, in g's ghost order
    for p in range(world):
        cs, ps, cc, pc = [], [], [], []
        for g in range(world):
            gh = shards[g]['ghost_ids']
            a, b = shards[g]['cell_recv_ptr'][p], shards[g]['cell_recv_ptr'][p + 1]
            ids = gh[a:b] if g != p else gh[:0]
            cs.append(local_of[ids])
            cc.append(ids.size)
            gp = shards[g]['ghost_pix']
            a, b = shards[g]['pix_recv_ptr'][p], shards[g]['pix_recv_ptr'][p + 1]
            pid = gp[a:b] if g != p else gp[:0]
            ps.append(pid - rows[p][0] * W)
            pc.append(pid.size)
        shards[p]['cell_send_ptr'] = _csr(cc)
        shards[p]['cell_send_idx'] = np.ascontiguousarray(np.concatenate(cs).astype(np.int32))
        shards[p]['pix_send_ptr'] = _csr(pc)
        shards[p]['pix_send_idx'] = np.ascontiguousarray(np.concatenate(ps).astype(np.int32))
    return shards
