#!/usr/bin/env python
"""Multi-GPU parity worker (one process per GPU; launched by tests/test_gpu_multi.py or by hand):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
      tests/mgpu_worker.py --variant deltaU_to_deltaP

Every rank builds the same seeded mesh and global tables, takes its block-row shard
(psm_b200.shard.partition), and runs the collective psm_predict on the rows of the cells it owns.
Rank 0 gathers the pressures and checks them against (a) the CPU oracle (<= 1e-3 rel-L2, the
north-star bound) and (b) a single-GPU handle on the whole mesh (same kernels: tight bound).
"""
import argparse
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
PKG = os.path.join(REPO, 'solving-poisson-s-equation-through-dl-for-cfd-apllications_b200')
for p in (PKG, REPO, HERE):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch                                   # noqa: E402
import torch.distributed as dist               # noqa: E402

import psm_b200                                # noqa: E402
from psm_b200 import synthetic as syn, tables as ptables, shard as pshard   # noqa: E402

MESHES = {'deltaU_to_deltaP': dict(H=500, W=420, nx=220, ny=260, R=0.12),
          'U_to_gradP': dict(H=340, W=300, nx=150, ny=170, R=0.1)}
# enough block rows for 4 and 8 ranks (deltas: stride 96, >= 1 block row per rank; gradP: stride 32, >= 3 per rank)
TALL = {'deltaU_to_deltaP': dict(H=1100, W=420, nx=210, ny=560, R=0.12),
        'U_to_gradP': dict(H=1000, W=300, nx=150, ny=500, R=0.1)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--variant', default='deltaU_to_deltaP')
    ap.add_argument('--near-wall', type=float, default=0.0)
    ap.add_argument('--halo', default='cells', choices=['cells', 'grid'])
    ap.add_argument('--comm', default='p2p', choices=['p2p', 'nccl'])
    ap.add_argument('--builder', default='partition', choices=['partition', 'band'],
                    help="partition: cut GLOBAL Qhull tables; band: every rank triangulates only its own band (what bench.py uses)")
    ap.add_argument('--legacy', action='store_true', help='separate push kernels + assembled field instead of the fused flow')
    args = ap.parse_args()
    if args.legacy:
        os.environ['PSM_MGPU_LEGACY'] = '1'
    if args.comm == 'nccl':
        os.environ['PSM_COMM'] = 'nccl'
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    local = int(os.environ.get('LOCAL_RANK', rank))
    dist.init_process_group('gloo')
    torch.cuda.set_device(local)
    variant = args.variant
    deltas = variant == 'deltaU_to_deltaP'
    mesh = syn.make_mesh(seed=11, **(TALL if world > 2 else MESHES)[variant])
    F = syn.make_fields(mesh, seed=11)
    params = syn.make_params(seed=11, pc_in=64, pc_p=48, standardization='std' if deltas else 'max_abs',
                             n_out_channels=1 if deltas else 2,
                             maxs=syn.DEFAULT_MAXS if deltas else (1.0, 0.536, 0.999, 0.8, 0.7))
    if args.builder == 'band':
        # closed-form grid -> cell tables on both sides (the band builder's), cells -> grid from Qhull
        tables = ptables.build_tables(mesh['cells'], mesh['top'], mesh['obst'], F['p_prev'], variant=variant, back='closed_form')
        Lb, s1 = pshard.band_phase1(mesh['cells'], mesh['top'], mesh['obst'], F['p_prev'], rank, world, variant=variant,
                                    near_wall_sdf=args.near_wall, halo=args.halo)
        all1 = [None] * world
        dist.all_gather_object(all1, s1)
        Lb, s2 = pshard.band_phase2(Lb, all1)
        all2 = [None] * world
        dist.all_gather_object(all2, s2)
        sh = pshard.band_phase3(Lb, all2)
    else:
        tables = ptables.build_tables(mesh['cells'], mesh['top'], mesh['obst'], F['p_prev'], variant=variant)
        shards = pshard.partition(tables, mesh['cells'], world, variant=variant, near_wall_sdf=args.near_wall, halo=args.halo)
        sh = shards[rank]
    cells = syn.pack_cells(mesh, F, with_delta=deltas)

    ids = [psm_b200.PressureSurrogate.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    sm = psm_b200.PressureSurrogate(variant, device=local, near_wall_sdf=args.near_wall)
    sm.load_params(params)
    sm.comm_init(ids[0], rank, world)
    sm.init_shard(sh)
    out = None
    own = sh['owned_ids']
    for _ in range(8):                          # repeated steps: buffers are reused, exchanges must not run ahead of their readers
        out, rc = sm.predict(cells[own])
    # the native-field entry on every rank (its own cells): same kernels after the first one -> bit-identical
    U3 = np.zeros((own.size, 3))
    U3[:, 0], U3[:, 1] = F['Ux'][own], F['Uy'][own]
    dU3 = None
    if deltas:
        dU3 = np.zeros((own.size, 3))
        dU3[:, 0], dU3[:, 1] = F['dUx'][own], F['dUy'][own]
    out_f, rc_f = sm.predict_fields(U3, p=F['p_prev'][own], dU=dU3)
    fields_equal = bool(np.array_equal(out_f, out, equal_nan=True)) and rc_f == rc
    # device-pointer entry: the whole step as one replayed graph
    d_in = torch.from_numpy(np.ascontiguousarray(cells[own])).cuda()
    d_o = torch.empty(out.shape, dtype=torch.float64, device='cuda')
    torch.cuda.synchronize()
    for _ in range(4):
        dist.barrier()
        sm.predict_device(d_in.data_ptr(), own.size, d_o.data_ptr(), sync=True)
    fields_equal = fields_equal and bool(np.array_equal(d_o.cpu().numpy(), out, equal_nan=True))
    # cell routing: a scotch-like decomposition -- every solver rank holds a random subset of the cells in a random order, the
    # library moves the rows to the block-row owners and the pressures back (replaces the gather-to-root of PMP:258, 501-511)
    n_all = mesh['cells'].shape[0]
    rng = np.random.default_rng(123)
    assign = rng.integers(0, world, size=n_all)
    mine = rng.permutation(np.flatnonzero(assign == rank))
    dr, di = pshard.route(sh['cell_rank'], mine)
    sm.route_init(dr, di)
    Ua = np.zeros((mine.size, 3))
    Ua[:, 0], Ua[:, 1] = F['Ux'][mine], F['Uy'][mine]
    dUa = None
    if deltas:
        dUa = np.zeros((mine.size, 3))
        dUa[:, 0], dUa[:, 1] = F['dUx'][mine], F['dUy'][mine]
    for _ in range(2):
        out_routed, rc_routed = sm.predict_routed(Ua, p=F['p_prev'][mine], dU=dUa)
    offsets = sm.stage('offsets')
    field = sm.stage('field')
    geo = sm.geometry()
    gathered = [None] * world
    dist.gather_object((sh['owned_ids'], out, offsets, field, (sh['row0'], sh['row1']), rc, fields_equal, mine, out_routed), gathered if rank == 0 else None, dst=0)
    sm.close()
    ok = True
    if rank == 0:
        from oracle.pipeline import DeltasOracle, GradPOracle, SurrogateParams
        P = SurrogateParams(**{k: v for k, v in params.items() if k != 'shape'})
        o = DeltasOracle(P) if deltas else GradPOracle(P)
        o.compute_only_once(mesh['cells'], mesh['top'], mesh['obst'], F['p_prev'],
                            tables=(tables['vert'], tables['weights'], tables['vert_back'], tables['weights_back']))
        n = mesh['cells'].shape[0]
        nf = 1 if deltas else 2
        full = np.full(n if deltas else (n, 2), np.nan)
        H, W = tables['H'], tables['W']
        fld = np.zeros((nf, H, W), np.float32)
        all_fields_equal = True
        routed = np.full(n if deltas else (n, 2), np.nan)
        for (own, o_out, o_offs, o_field, (r0, r1), o_rc, o_feq, o_mine, o_routed) in gathered:
            all_fields_equal = all_fields_equal and o_feq
            full[own] = o_out
            routed[o_mine] = o_routed
            fld[:, r0:r1] = o_field
            assert np.array_equal(np.isnan(o_offs), np.isnan(gathered[0][2])) and np.allclose(o_offs, gathered[0][2], rtol=0, atol=0, equal_nan=True), \
                'offsets differ between ranks'
        with psm_b200.PressureSurrogate(variant, device=local, near_wall_sdf=args.near_wall) as one:
            one.load_params(params)
            one.init_tables(tables)
            single, _ = one.predict(cells)
            single_field = one.stage('field')
        rel = lambda a, b: float(np.linalg.norm(np.nan_to_num(a - b)) / max(np.linalg.norm(np.nan_to_num(b)), 1e-300))   # noqa: E731
        if deltas:
            r = o.time_step(F['Ux'], F['Uy'], F['dUx'], F['dUy'])
            p_ref, _ = o.to_cells(r['field'], F['p_prev'], near_wall_sdf=args.near_wall if args.near_wall > 0 else None)
            e_or = rel(full - F['p_prev'], p_ref - F['p_prev'])
            e_field = rel(fld[0], r['field'])
        else:
            r = o.time_step(F['Ux'], F['Uy'])
            p_ref = np.stack([o.to_cells(r['dp_dx']), o.to_cells(r['dp_dy'])], axis=1)
            assert np.array_equal(np.isnan(full), np.isnan(p_ref)), 'NaN pattern differs from the oracle'
            e_or = rel(full, p_ref)
            e_field = max(rel(fld[0], r['dp_dx']), rel(fld[1], r['dp_dy']))
        e_single = rel(full, single)
        e_sf = rel(fld, single_field)
        assert geo['peer_memory_exchange'] == int(args.comm == 'p2p' and args.halo == 'cells'), geo
        print('mgpu %s world=%d builder=%s: rel-L2 vs oracle cells %.2e field %.2e | vs single-GPU cells %.2e field %.2e | '
              'ghost cells/pix on rank0 %d/%d | fields/device entry bit-identical: %s' %
              (variant, world, args.builder, e_or, e_field, e_single, e_sf, geo['n_ghost_cells'], geo['n_ghost_pix'], all_fields_equal))
        routed_equal = bool(np.array_equal(routed, full, equal_nan=True))
        print('routed (random decomposition) == block-row owned rows, bit for bit: %s' % routed_equal)
        ok = e_or < 1e-3 and e_field < 1e-3 and e_single < 1e-5 and e_sf < 1e-5 and all_fields_equal and routed_equal
        print('MGPU_PARITY_OK' if ok else 'MGPU_PARITY_FAIL')
    dist.barrier()
    dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == '__main__':
    sys.exit(main())
