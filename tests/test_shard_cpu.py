"""Host-side logic of the multi-GPU path on the CPU (no GPU): the block-row partitioner and the
four-exchange protocol of DESIGN.md section 5, run as world_size-2/3 ``gloo`` processes.

Each process restates one rank's step in NumPy on ITS shard only (owned cells, own pixel rows, local
blocks), doing the same exchanges the CUDA library does with NCCL -- max|U|, ghost cells, overlap
strips, strip means / shift-line sums, ghost pixels -- through torch.distributed/gloo.  The gathered
result must equal the unsharded oracle (float64 both sides, so the bound is round-off).
"""
import os
import socket
import warnings

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import psm_b200
from psm_b200 import synthetic as syn, tables as ptables, shard as pshard
from oracle.pipeline import DeltasOracle, GradPOracle, mlp_forward, pca_transform
from helpers import oracle_params

MESHES = {'deltaU_to_deltaP': dict(H=500, W=420, nx=160, ny=200, R=0.12),
          'U_to_gradP': dict(H=340, W=300, nx=110, ny=130, R=0.1)}


def make_case(variant, seed=5):
    deltas = variant == 'deltaU_to_deltaP'
    mesh = syn.make_mesh(seed=seed, **MESHES[variant])
    F = syn.make_fields(mesh, seed=seed)
    params = syn.make_params(seed=seed, pc_in=24, pc_p=16, hidden=(32, 32), standardization='std' if deltas else 'max_abs',
                             n_out_channels=1 if deltas else 2,
                             maxs=syn.DEFAULT_MAXS if deltas else (1.0, 0.536, 0.999, 0.8, 0.7))
    tables = ptables.build_tables(mesh['cells'], mesh['top'], mesh['obst'], F['p_prev'], variant=variant)
    return mesh, F, params, tables


# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("H,ov,world", [(500, 32, 2), (1000, 32, 4), (4000, 32, 8), (340, 96, 2), (1000, 96, 8)])
def test_block_row_split_covers_grid(H, ov, world):
    ranges, rows = pshard.block_row_split(H, 128, ov, world)
    st = 128 - ov
    n_y = (H - 128) // st
    assert ranges[0][0] == 0 and ranges[-1][1] == n_y + 2
    assert rows[0][0] == 0 and rows[-1][1] == H
    for g in range(world):
        if g:
            assert ranges[g][0] == ranges[g - 1][1] and rows[g][0] == rows[g - 1][1]
        b0, b1 = ranges[g]
        assert rows[g][0] == b0 * st
        # every block of the rank lies inside its rows + the overlap strip it receives
        last_y0 = (H - 128) if (g == world - 1) else (b1 - 1) * st
        ext = 0 if g == world - 1 else ov
        assert last_y0 + 128 <= rows[g][1] + ext
        if g < world - 1:
            assert rows[g + 1][1] - rows[g + 1][0] >= ov        # the strip comes from ONE neighbour


def test_block_row_split_rejects_too_many_ranks():
    with pytest.raises(ValueError):
        pshard.block_row_split(300, 128, 32, 4)
    with pytest.raises(ValueError):
        pshard.block_row_split(340, 128, 96, 4)


@pytest.mark.parametrize("variant,world", [('deltaU_to_deltaP', 2), ('deltaU_to_deltaP', 3), ('U_to_gradP', 2)])
def test_partition_is_a_relabelling_of_the_global_tables(variant, world):
    mesh, F, params, t = make_case(variant)
    H, W = t['H'], t['W']
    shards = pshard.partition(t, mesh['cells'], world, variant=variant, halo='grid' if world == 3 else 'cells')
    fv, fw = pshard.fold_forward_table(t['vert'], t['weights'], t['indices'], H, W)
    bv, keep = pshard.hop_back_table(t['vert_back'], t['weights_back'], t['indices'], t['sdfunct'], W)
    owned_all = np.concatenate([s['owned_ids'] for s in shards])
    assert np.array_equal(np.sort(owned_all), np.arange(t['n_cells']))          # every cell owned exactly once
    for s in shards:
        q0, q1 = s['row0'] * W, s['row1'] * W
        qg = q1 + s['local_ext_rows'] * W
        l2g = np.concatenate([s['owned_ids'], s['ghost_ids']])
        live = np.any(s['weights'] != 0, axis=1)
        assert np.array_equal(l2g[s['vert']][live], fv[q0:qg][live])
        assert np.array_equal(s['weights'], fw[q0:qg])
        pix_l2g = np.concatenate([np.arange(q0, q1), s['ghost_pix']])
        k = keep[s['owned_ids']]
        assert np.array_equal(s['vert_back'][:, 0] < 0, k)
        assert np.array_equal(pix_l2g[s['vert_back'][~k]], bv[s['owned_ids']][~k])
        # what the peers send is exactly what this rank expects, in its ghost order
        for p in range(world):
            sp = shards[p]
            a, b = sp['cell_send_ptr'][s['rank']], sp['cell_send_ptr'][s['rank'] + 1]
            ra, rb = s['cell_recv_ptr'][p], s['cell_recv_ptr'][p + 1]
            assert np.array_equal(sp['owned_ids'][sp['cell_send_idx'][a:b]], s['ghost_ids'][ra:rb])
            a, b = sp['pix_send_ptr'][s['rank']], sp['pix_send_ptr'][s['rank'] + 1]
            ra, rb = s['pix_recv_ptr'][p], s['pix_recv_ptr'][p + 1]
            assert np.array_equal(sp['pix_send_idx'][a:b] + sp['row0'] * W, s['ghost_pix'][ra:rb])
        assert s['cell_send_ptr'][s['rank'] + 1] == s['cell_send_ptr'][s['rank']]     # nothing to itself


# ------------------------------------------------------------------------------------------------
def _exchange(send_arrays, recv_counts, width, rank, world):
    """Static sparse exchange over gloo: send_arrays[p] -> rank p; returns the concatenated ghosts."""
    reqs, recv = [], []
    for p in range(world):
        if p == rank:
            recv.append(np.zeros((0, width)))
            continue
        buf = torch.zeros((int(recv_counts[p]), width), dtype=torch.float64)
        recv.append(buf)
        if buf.numel():
            reqs.append(dist.irecv(buf, src=p))
        if send_arrays[p].size:
            reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(send_arrays[p], dtype=np.float64)), dst=p))
    for r in reqs:
        r.wait()
    return np.concatenate([np.asarray(b) for b in recv])


def _rank_step(rank, world, port, variant, q, halo='cells'):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        deltas = variant == 'deltaU_to_deltaP'
        mesh, F, params, t = make_case(variant)
        ov = 32 if deltas else 96
        P = oracle_params(params)
        sh = pshard.partition(t, mesh['cells'], world, variant=variant, halo=halo)[rank]
        H, W, S = sh['H'], sh['W'], 128
        own = sh['owned_ids']
        # ---- exchange 1: max|U|^2 and ghost cells -------------------------------------------------
        Ux, Uy = F['Ux'][own], F['Uy'][own]
        fld = np.stack([F['dUx'][own], F['dUy'][own]], 1) if deltas else np.stack([Ux, Uy], 1)
        um = torch.tensor([np.max(np.square(Ux) + np.square(Uy))], dtype=torch.float64)
        dist.all_reduce(um, op=dist.ReduceOp.MAX)
        U_max_norm = float(np.sqrt(um.item()))
        sp = sh['cell_send_ptr']
        ghosts = _exchange([fld[sh['cell_send_idx'][sp[p]:sp[p + 1]]] for p in range(world)],
                           np.diff(sh['cell_recv_ptr']), 2, rank, world)
        uv = np.concatenate([fld, ghosts]) / U_max_norm
        # ---- gather own rows, then exchange 2: the overlap strip --------------------------------------
        r0, r1, ext = sh['row0'], sh['row1'], sh['ext_rows']
        g = np.einsum('qjc,qj->qc', uv[sh['vert']], sh['weights']).reshape(r1 - r0 + sh['local_ext_rows'], W, 2)
        g = np.nan_to_num(g, nan=0.0) / np.array([P.maxs[0], P.maxs[1]])
        reqs = []
        if sh['send_rows']:
            reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(g[:sh['send_rows']])), dst=rank - 1))
        halo = torch.zeros((ext, W, 2), dtype=torch.float64)
        if ext:
            reqs.append(dist.irecv(halo, src=rank + 1))
        for r in reqs:
            r.wait()
        grid = np.concatenate([g, np.asarray(halo)], 0)
        grid = np.concatenate([grid, (sh['sdfunct'] / P.maxs[2])[..., None]], -1)
        # ---- local blocks through PCA / MLP / PCA^-1 --------------------------------------------------
        plan = psm_b200.compile_plan(variant, H, W, sh['mask_global'], overlap=ov)
        org = plan['origins']
        ncolb = org.shape[0] // (plan['indices_list'][:, 0].max() + 1)
        kb0, kb1 = sh['blk_row0'] * ncolb, sh['blk_row1'] * ncolb
        x_array = np.stack([grid[org[k, 0] - r0:org[k, 0] - r0 + S, org[k, 1]:org[k, 1] + S] for k in range(kb0, kb1)])
        z = pca_transform(x_array.reshape(kb1 - kb0, -1), P.pca_in_components, P.pca_in_mean)
        x_in = (z - P.mean_in) / P.std_in if P.standardization == 'std' else z / P.max_abs_input_PCA
        r_out = mlp_forward(x_in, P.mlp_weights, P.mlp_biases)
        r_out = r_out * P.std_out + P.mean_out if P.standardization == 'std' else r_out * P.max_abs_output_PCA
        C = P.n_out_channels
        blocks = (np.dot(r_out, P.pca_out_components) + P.pca_out_mean).reshape(kb1 - kb0, S, S, C)
        if deltas:
            blocks = blocks * P.maxs[3] * U_max_norm ** 2
        blocks = np.moveaxis(blocks, -1, 1)                                        # [B_loc, C, S, S]
        # ---- exchange 3: strip means + shift-line sums (disjoint slots, all-reduce sum) ------------------
        mask = sh['mask_global']
        nt = len(plan['tasks'])
        slots = np.zeros(nt + len(plan['lines']))
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", category=RuntimeWarning)
            for i, (src, msk, ch, y0, y1, x0, x1, cnt) in enumerate(plan['tasks']):
                if kb0 <= src < kb1:
                    m = mask[org[msk, 0] + y0:org[msk, 0] + y1, org[msk, 1] + x0:org[msk, 1] + x1] != 0
                    slots[i] = np.mean(blocks[src - kb0, ch, y0:y1, x0:x1][m])
        for i, (lf, blk, y0, y1, x0, x1, coef, n) in enumerate(plan['lines']):
            if kb0 <= blk < kb1:
                slots[nt + i] = blocks[blk - kb0, lf, y0:y1, x0:x1].sum()
        ts = torch.from_numpy(slots)
        dist.all_reduce(ts)
        slots = np.asarray(ts)
        B, Fn = plan['n_blocks'], plan['n_fields']
        c = np.zeros((Fn, B))
        for f in range(Fn):
            for k in range(B):
                ta, tb, par, _ = plan['rec'][f, k]
                c[f, k] = slots[ta] - ((slots[tb] - c[f, par]) if tb >= 0 else 0.0)
        shift = np.zeros(Fn)
        for i, (lf, blk, y0, y1, x0, x1, coef, n) in enumerate(plan['lines']):
            shift[lf] += coef * (slots[nt + i] - n * c[lf, blk])
        for f in range(Fn):
            shift[f] /= 3.0 * (H if (deltas or f == 0) else W)
        # ---- placement of own rows, exchange 4: ghost pixels, grid->cell gather ----------------------------
        ow = plan['owner'][r0:r1]
        assert ow.min() >= kb0 and ow.max() < kb1
        yy, xx = np.mgrid[r0:r1, 0:W]
        field = np.stack([blocks[ow - kb0, f, yy - org[ow, 0], xx - org[ow, 1]] - c[f][ow] - shift[f] for f in range(Fn)])
        flat = field.reshape(Fn, -1)
        sp = sh['pix_send_ptr']
        gp = _exchange([flat[:, sh['pix_send_idx'][sp[p]:sp[p + 1]]].T for p in range(world)],
                       np.diff(sh['pix_recv_ptr']), Fn, rank, world)
        ext_field = np.concatenate([flat, gp.T], axis=1)
        vb, wb = sh['vert_back'], sh['weights_back']
        keep = vb[:, 0] < 0
        vals = np.einsum('fnj,nj->nf', ext_field[:, np.where(keep[:, None], 0, vb)], wb)
        if deltas:
            dp = vals[:, 0]
            dp[keep | np.isnan(dp)] = 0.0
            out = F['p_prev'][own] + dp
        else:
            out = vals
            out[keep] = np.nan
        gathered = [None] * world
        dist.gather_object((own, out, field, (r0, r1)), gathered if rank == 0 else None, dst=0)
        if rank == 0:
            q.put(gathered)
        dist.barrier()
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("variant,world,halo", [('deltaU_to_deltaP', 2, 'cells'), ('U_to_gradP', 2, 'cells'),
                                                ('deltaU_to_deltaP', 3, 'grid'), ('deltaU_to_deltaP', 4, 'cells')])
def test_sharded_step_over_gloo_equals_unsharded_oracle(variant, world, halo):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_rank_step, args=(r, world, port, variant, q, halo)) for r in range(world)]
    for p in procs:
        p.start()
    gathered = q.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    deltas = variant == 'deltaU_to_deltaP'
    mesh, F, params, t = make_case(variant)
    o = DeltasOracle(oracle_params(params)) if deltas else GradPOracle(oracle_params(params))
    o.compute_only_once(mesh['cells'], mesh['top'], mesh['obst'], F['p_prev'],
                        tables=(t['vert'], t['weights'], t['vert_back'], t['weights_back']))
    n = mesh['cells'].shape[0]
    full = np.full(n if deltas else (n, 2), np.nan)
    fld = np.zeros((1 if deltas else 2, t['H'], t['W']))
    for own, out, field, (r0, r1) in gathered:
        full[own] = out
        fld[:, r0:r1] = field
    if deltas:
        r = o.time_step(F['Ux'], F['Uy'], F['dUx'], F['dUy'])
        ref, _ = o.to_cells(r['field'], F['p_prev'])
        ref_f = r['field'][None]
    else:
        r = o.time_step(F['Ux'], F['Uy'])
        ref = np.stack([o.to_cells(r['dp_dx']), o.to_cells(r['dp_dy'])], axis=1)
        ref_f = np.stack([r['dp_dx'], r['dp_dy']])
    scale = np.nanmax(np.abs(ref_f))
    np.testing.assert_allclose(fld, ref_f, rtol=0, atol=2e-5 * scale)       # float32 Dense stack on both sides
    assert np.array_equal(np.isnan(full), np.isnan(ref))
    np.testing.assert_allclose(full, ref, rtol=0, atol=2e-5 * scale, equal_nan=True)


@pytest.mark.parametrize("variant,world", [('deltaU_to_deltaP', 2), ('deltaU_to_deltaP', 3), ('U_to_gradP', 2)])
def test_band_local_shards_equal_partition_of_global_tables(variant, world):
    """Triangulating only a band of cells per rank gives the same shard as cutting the global tables
    (same simplices; Qhull may list a simplex's vertices in another order, so entries are compared
    after sorting by cell id; weights to round-off)."""
    deltas = variant == 'deltaU_to_deltaP'
    mesh = syn.make_mesh(seed=5, **MESHES[variant])
    F = syn.make_fields(mesh, seed=5)
    t = ptables.build_tables(mesh['cells'], mesh['top'], mesh['obst'], F['p_prev'], variant=variant, back='closed_form')
    ref = pshard.partition(t, mesh['cells'], world, variant=variant, near_wall_sdf=0.05 if deltas else 0.0)
    got = pshard.build_band_shards_serial(mesh['cells'], mesh['top'], mesh['obst'], F['p_prev'], world, variant=variant,
                                          near_wall_sdf=0.05 if deltas else 0.0)

    def canon(v, w):
        o = np.argsort(v, axis=1, kind='stable')
        return np.take_along_axis(v, o, 1), np.take_along_axis(w, o, 1)
    for a, b in zip(ref, got):
        for k in ('rank', 'world', 'H', 'W', 'row0', 'row1', 'ext_rows', 'send_rows', 'blk_row0', 'blk_row1', 'n_owned',
                  'n_ghost', 'n_ghost_pix', 'local_ext_rows'):
            assert a[k] == b[k], k
        for k in ('mask_global', 'owned_ids', 'ghost_ids', 'ghost_pix', 'cell_send_ptr', 'cell_send_idx', 'cell_recv_ptr',
                  'pix_send_ptr', 'pix_send_idx', 'pix_recv_ptr', 'vert_back'):
            assert np.array_equal(a[k], b[k]), k
        np.testing.assert_allclose(a['sdfunct'], b['sdfunct'], rtol=0, atol=1e-14)
        np.testing.assert_allclose(a['weights_back'], b['weights_back'], rtol=0, atol=0)
        va, wa = canon(a['vert'], a['weights'])
        vb, wb = canon(b['vert'], b['weights'])
        live = np.any(wa != 0, axis=1)
        assert np.array_equal(live, np.any(wb != 0, axis=1))
        assert np.array_equal(va[live], vb[live])
        np.testing.assert_allclose(wa, wb, rtol=0, atol=1e-11)


def test_route_sends_every_cell_to_its_block_row_owner():
    """psm_b200.shard.route: for a random (scotch-like) decomposition every cell is routed to the rank whose pixel rows contain
    it, at its position among that rank's owned cells (ascending global id) -- the order psm_predict takes them in."""
    mesh = syn.make_mesh(seed=3, H=500, W=420, nx=220, ny=260, R=0.12)
    F = syn.make_fields(mesh, seed=3)
    t = ptables.build_tables(mesh['cells'], mesh['top'], mesh['obst'], F['p_prev'])
    world = 3
    shards = pshard.partition(t, mesh['cells'], world)
    n = mesh['cells'].shape[0]
    rng = np.random.default_rng(9)
    assign = rng.integers(0, 5, size=n)                      # 5 solver ranks onto 3 GPU ranks
    seen = np.zeros(n, bool)
    for r in range(5):
        mine = rng.permutation(np.flatnonzero(assign == r))
        dr, di = pshard.route(shards[0]['cell_rank'], mine)
        for g in range(world):
            sel = dr == g
            np.testing.assert_array_equal(shards[g]['owned_ids'][di[sel]], mine[sel])
        seen[mine] = True
    assert seen.all()
