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


def _fill_ghost_gaps(ghost, cell_rank, max_gap=16, slack=0.25):
    """The ghost cells a rank needs from one owner are (nearly) whole cell rows next to the rank boundary, minus the few cells
    the forward table happens not to reference.  Closing the small holes (<= ``max_gap`` ids, at most ``slack`` more cells in
    total) turns the list into a handful of runs of consecutive cell ids, which the owner's prep kernel pushes with coalesced
    stores as it converts them (``SendRun`` in csrc/psm_kernels.cuh).  Far-away singles (the cells of the (0,0) raster quirk)
    stay singles.  ``ghost`` must be sorted by (owner, id); returns the same ordering."""
    if ghost.size == 0:
        return ghost
    out = []
    owners = cell_rank[ghost]
    for o in np.unique(owners):
        g = ghost[owners == o]
        d = np.diff(g)
        holes = np.flatnonzero((d > 1) & (d <= max_gap + 1))
        if holes.size == 0 or int((d[holes] - 1).sum()) > slack * g.size + 64:
            out.append(g)
            continue
        extra = np.concatenate([np.arange(g[k] + 1, g[k + 1]) for k in holes])
        extra = extra[cell_rank[extra] == o]
        out.append(np.unique(np.concatenate([g, extra])))
    return np.concatenate(out)


def route(cell_rank, my_cells):
    """Where a solver rank's own cells go (``psm_route``).  ``cell_rank`` [N]: owner rank of every GLOBAL cell (the rank whose
    pixel rows contain it -- ``partition`` / ``band_phase1`` compute it); ``my_cells``: global ids of the cells this solver rank
    holds, in its own order.  Returns (dest_rank, dest_index): the owner and the cell's position among the owner's cells (which
    are passed to ``psm_predict`` in ascending global id, ``owned_ids``)."""
    cell_rank = np.asarray(cell_rank)
    my_cells = np.asarray(my_cells, dtype=np.int64)
    pos = np.empty(cell_rank.size, np.int64)
    for g in np.unique(cell_rank):
        sel = np.flatnonzero(cell_rank == g)
        pos[sel] = np.arange(sel.size)
    return cell_rank[my_cells].astype(np.int32), pos[my_cells].astype(np.int32)


def _csr(counts):
    p = np.zeros(len(counts) + 1, np.int64)
    p[1:] = np.cumsum(counts)
    return p


def partition(tables, cells_xy, world, variant='deltaU_to_deltaP', shape=128, overlap=None, near_wall_sdf=0.0,
              halo='cells'):
    """Split GLOBAL tables (``psm_b200.tables.build_tables``) into ``world`` shard dicts.

    Each dict holds what ``psm_shard`` needs plus ``owned_ids`` (global ids of the cells this rank
    passes to / gets from ``psm_predict``, ascending).  ``halo``: how the overlap strip below a rank's
    rows reaches it -- 'cells': as ghost cells, the rank gathers those rows itself (one exchange
    fewer per step); 'grid': as gathered grid rows sent by rank+1."""
    assert halo in ('cells', 'grid')
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
        ext = 0 if g == world - 1 else overlap
        lext = ext if halo == 'cells' else 0
        v, w = fv[q0:q1 + lext * W], fw[q0:q1 + lext * W]
        live = np.any(w != 0.0, axis=1)
        need = np.unique(v[live])
        ghost = need[cell_rank[need] != g]
        ghost = ghost[np.lexsort((ghost, cell_rank[ghost]))]            # by owner rank, then global id
        ghost = _fill_ghost_gaps(ghost, cell_rank)
        lut = np.zeros(N, np.int64)
        lut[owned[g]] = np.arange(owned[g].size)
        lut[ghost] = owned[g].size + np.arange(ghost.size)
        lv = np.where(live[:, None], lut[v], 0).astype(np.int32)
        sh = dict(rank=g, world=world, H=H, W=W, row0=r0, row1=r1, cell_rank=cell_rank,
                  ext_rows=ext - lext, send_rows=0 if (g == 0 or halo == 'cells') else overlap, local_ext_rows=lext,
                  blk_row0=ranges[g][0], blk_row1=ranges[g][1], mask_global=mask,
                  owned_ids=owned[g], ghost_ids=ghost, n_owned=int(owned[g].size), n_ghost=int(ghost.size),
                  vert=np.ascontiguousarray(lv), weights=np.ascontiguousarray(w),
                  cell_recv_ptr=_csr(np.bincount(cell_rank[ghost], minlength=world)))
        sh['sdfunct'] = np.ascontiguousarray(sdf[r0:r1 + ext])
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


# ------------------------------------------------------------------------------------------------
# Band-local construction: the same shard as partition(), without ever triangulating the whole mesh.
# Each rank triangulates only the cells near its own pixel rows (Delaunay is local: a triangle whose
# circumcircle stays `margin_px` pixels away from the band edge is a triangle of the global mesh), so
# init time and memory stay per-GPU constant as ranks are added.  Three phases because two small
# facts are global: the last invalid pixel (the (0,0) raster quirk, SMC:161,432) and who needs which
# ghost cell; the caller moves the phase summaries between ranks (torch.distributed / MPI).
def band_phase1(cells_xy, top, obst, probe_values, rank, world, variant='deltaU_to_deltaP', delta=5e-3, shape=128,
                overlap=None, near_wall_sdf=0.0, margin_px=12, halo='cells', keep_raw=False):
    """``keep_raw``: also keep this rank's rows of the UNFOLDED cells -> grid table (global cell ids, float64 weights) as
    ``L['raw_vert']`` / ``L['raw_weights']`` -- concatenated over the ranks they are the global table of UTL:38-44 (bench.py
    builds a single-GPU handle and the oracle from them to check the sharded run)."""
    from . import tables as T
    assert halo in ('cells', 'grid')
    if overlap is None:
        overlap = 32 if variant == 'deltaU_to_deltaP' else 96
    cells_xy = np.ascontiguousarray(cells_xy, dtype=np.float64)
    nd = 3 if variant == 'deltaU_to_deltaP' else 2
    x_min, x_max = round(float(cells_xy[:, 0].min()), nd), round(float(cells_xy[:, 0].max()), nd)
    y_min, y_max = round(float(cells_xy[:, 1].min()), nd), round(float(cells_xy[:, 1].max()), nd)
    H, W = int(round((y_max - y_min) / delta)), int(round((x_max - x_min) / delta))
    Xr = np.linspace(x_min + delta / 2, x_max - delta / 2, num=W)
    Yc = np.linspace(y_min + delta / 2, y_max - delta / 2, num=H)
    ranges, rows = block_row_split(H, shape, overlap, world)
    r0, r1 = rows[rank]
    ext = 0 if rank == world - 1 else overlap
    ta, tb = max(0, r0 - 2), min(H, r1 + max(ext, 2))
    row_lo = np.array([r[0] for r in rows])
    cell_row = np.clip(np.floor((cells_xy[:, 1] - y_min) / delta).astype(np.int64), 0, H - 1)
    cell_rank = np.searchsorted(row_lo, cell_row, side='right') - 1
    owned = np.flatnonzero(cell_rank == rank)
    band = np.flatnonzero((cells_xy[:, 1] >= Yc[ta] - margin_px * delta) & (cells_xy[:, 1] <= Yc[tb - 1] + margin_px * delta))
    XX, YY = np.meshgrid(Xr, Yc[ta:tb])
    xy0 = np.stack([XX.ravel(), YY.ravel()], axis=1)
    lv, w = T.barycentric_tables(cells_xy[band], xy0)                      # UTL:38-44 on the band
    gv = band[lv]                                                          # global cell ids
    domain, sdf = T.flow_mask_and_distance(xy0, np.asarray(top, dtype=np.float64), np.asarray(obst, dtype=np.float64),
                                           variant, (x_min, x_max, y_min, y_max))
    probe = np.einsum('nj,nj->n', np.take(np.asarray(probe_values, dtype=np.float64), gv), w)
    neg = np.any(w < 0, axis=1)
    ok = domain & ~neg & ~np.isnan(probe)
    sdfunct = np.where(ok, sdf, 0.0).reshape(tb - ta, W)
    ok2 = ok.reshape(tb - ta, W)
    lext = ext if halo == 'cells' else 0
    o0, o1 = (r0 - ta) * W, (r1 - ta) * W                                  # own rows inside the band rows
    og = o1 + lext * W                                                     # + the overlap rows gathered locally
    fv = np.where(ok[o0:og, None], gv[o0:og], 0)
    fw = np.where(ok[o0:og, None], w[o0:og], 0.0)
    bad = np.flatnonzero(~ok[o0:o1])
    last_invalid = None
    if bad.size:
        q = int(bad[-1])
        wq = w[o0 + q].copy()
        if np.any(wq < 0):
            wq[:] = 0.0
        last_invalid = (r0 * W + q, gv[o0 + q].copy(), wq)
    L = dict(rank=rank, world=world, H=H, W=W, rows=rows, ranges=ranges, row0=r0, row1=r1, ext_rows=ext - lext,
             local_ext_rows=lext, send_rows=0 if (rank == 0 or halo == 'cells') else overlap, ta=ta, tb=tb, owned=owned, cell_rank=cell_rank, fv=fv, fw=fw,
             sdfunct=np.ascontiguousarray(sdfunct[r0 - ta:r1 + ext - ta]), ok_rows=ok2, near_wall_sdf=near_wall_sdf,
             bbox=(x_min, x_max, y_min, y_max), delta=delta)
    if keep_raw:
        L['raw_vert'] = np.ascontiguousarray(gv[o0:o1].astype(np.int32))
        L['raw_weights'] = np.ascontiguousarray(w[o0:o1])
    # grid -> cell table of the owned cells in closed form (tables.regular_grid_back_tables), hop and keep mask
    bvert, bw = T.regular_grid_back_tables(cells_xy[owned], Xr, Yc, W)
    brow = bvert // W
    if bvert.size and (brow.min() < ta or brow.max() >= tb):
        raise ValueError('a grid->cell triangle leaves the band rows')
    okv = ok2[brow - ta, bvert % W]
    keep = np.any(bw < 0, axis=1)
    if near_wall_sdf and near_wall_sdf > 0:
        sdf_mesh = np.einsum('nj,nj->n', sdfunct[brow - ta, bvert % W], bw)
        keep |= (~keep) & (sdf_mesh < near_wall_sdf)
    L.update(bv=np.where(okv, bvert, 0).astype(np.int64), bw=bw, keep=keep)
    summary = dict(rank=rank, last_invalid=last_invalid, mask_rows=np.ascontiguousarray(sdfunct[r0 - ta:r1 - ta] != 0.0, dtype=np.uint8))
    return L, summary


def band_phase2(L, summaries):
    """Apply the global facts of phase 1 (mask of the whole grid, the (0,0) quirk) and list this rank's ghosts."""
    summaries = sorted(summaries, key=lambda s: s['rank'])
    L['mask_global'] = np.ascontiguousarray(np.concatenate([s['mask_rows'] for s in summaries], axis=0))
    cands = [s['last_invalid'] for s in summaries if s['last_invalid'] is not None]
    if L['rank'] == 0 and cands:
        q, v, w = max(cands, key=lambda c: c[0])
        L['fv'][0], L['fw'][0] = v, w                                        # pixel (0,0) <- last invalid point
    rank, W = L['rank'], L['W']
    live = np.any(L['fw'] != 0.0, axis=1)
    need = np.unique(L['fv'][live])
    ghost = need[L['cell_rank'][need] != rank]
    ghost = ghost[np.lexsort((ghost, L['cell_rank'][ghost]))]
    ghost = _fill_ghost_gaps(ghost, L['cell_rank'])
    q0, q1 = L['row0'] * W, L['row1'] * W
    needp = np.unique(L['bv'][~L['keep']])
    gp = needp[(needp < q0) | (needp >= q1)]
    L.update(live=live, ghost_ids=ghost, ghost_pix=gp)
    return L, dict(rank=rank, ghost_ids=ghost, ghost_pix=gp)


def band_phase3(L, ghost_lists):
    """Local numbering and the static exchange lists -> the dict ``PressureSurrogate.init_shard`` takes."""
    ghost_lists = sorted(ghost_lists, key=lambda s: s['rank'])
    rank, world, W = L['rank'], L['world'], L['W']
    rows = L['rows']
    row_lo = np.array([r[0] for r in rows])
    owned, ghost, gp = L['owned'], L['ghost_ids'], L['ghost_pix']
    q0, q1 = L['row0'] * W, L['row1'] * W
    ids = np.concatenate([owned, ghost])
    order = np.argsort(ids, kind='stable')
    pos = np.searchsorted(ids[order], L['fv'])
    lv = np.where(L['live'][:, None], order[np.clip(pos, 0, ids.size - 1)], 0).astype(np.int32)
    b = L['bv']
    inside = (b >= q0) & (b < q1)
    lb = np.empty_like(b)
    lb[inside] = b[inside] - q0
    lb[~inside] = (q1 - q0) + (np.clip(np.searchsorted(gp, b[~inside]), 0, max(gp.size - 1, 0)) if gp.size else 0)
    lb[L['keep']] = 0
    lb[L['keep'], 0] = -1
    pix_rank = lambda q: np.searchsorted(row_lo, q // W, side='right') - 1          # noqa: E731
    owned_sorted_pos = lambda gids: np.searchsorted(owned, gids)                    # noqa: E731  (owned is ascending)
    cs, ps, cc, pc = [], [], [], []
    for g in range(world):
        gl = ghost_lists[g]
        if g == rank:
            cs.append(np.zeros(0, np.int64)); ps.append(np.zeros(0, np.int64)); cc.append(0); pc.append(0)
            continue
        mine = gl['ghost_ids'][L['cell_rank'][gl['ghost_ids']] == rank]
        cs.append(owned_sorted_pos(mine)); cc.append(mine.size)
        pm = gl['ghost_pix'][(gl['ghost_pix'] >= q0) & (gl['ghost_pix'] < q1)]
        ps.append(pm - q0); pc.append(pm.size)
    sh = dict(rank=rank, world=world, H=L['H'], W=W, row0=L['row0'], row1=L['row1'], ext_rows=L['ext_rows'], cell_rank=L['cell_rank'],
              send_rows=L['send_rows'], local_ext_rows=L['local_ext_rows'], blk_row0=L['ranges'][rank][0], blk_row1=L['ranges'][rank][1],
              mask_global=L['mask_global'], owned_ids=owned, ghost_ids=ghost, n_owned=int(owned.size), n_ghost=int(ghost.size),
              vert=np.ascontiguousarray(lv), weights=np.ascontiguousarray(L['fw']), sdfunct=L['sdfunct'],
              vert_back=np.ascontiguousarray(lb.astype(np.int32)), weights_back=np.ascontiguousarray(L['bw']),
              ghost_pix=gp, n_ghost_pix=int(gp.size),
              cell_recv_ptr=_csr(np.bincount(L['cell_rank'][ghost], minlength=world) if ghost.size else np.zeros(world, np.int64)),
              pix_recv_ptr=_csr(np.bincount(pix_rank(gp), minlength=world) if gp.size else np.zeros(world, np.int64)),
              cell_send_ptr=_csr(cc), cell_send_idx=np.ascontiguousarray(np.concatenate(cs).astype(np.int32)),
              pix_send_ptr=_csr(pc), pix_send_idx=np.ascontiguousarray(np.concatenate(ps).astype(np.int32)))
    return sh


def build_band_shards_serial(cells_xy, top, obst, probe_values, world, **kw):
    """All ranks' band shards in one process (tests, small meshes): the three phases with the
    exchanges replaced by list passing."""
    p1 = [band_phase1(cells_xy, top, obst, probe_values, r, world, **kw) for r in range(world)]
    p2 = [band_phase2(L, [s for _, s in p1]) for L, _ in p1]
    return [band_phase3(L, [s for _, s in p2]) for L, _ in p2]
