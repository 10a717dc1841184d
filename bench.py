#!/usr/bin/env python
"""Benchmark of the pressure-surrogate hot path (BASELINE.json metric: mesh cells/s, ms per step).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference] [--workload c2]

One "step" = one per-timestep surrogate prediction (``py_func`` in the reference) over the whole
synthetic mesh.  N=1 runs BASELINE.json configs[1]: deltaU_to_deltaP at ~1 M cells on a 1000x1000
grid, random-init weights of the reference architecture (pc 128 -> 3x512 -> pc 128).
  value    : cells/s with the solver's double[n][7] rows already resident in HBM (psm_predict_device),
             timed with CUDA events on the handle's stream, L2 flushed between steps.
  e2e      : cells/s through the public host-buffer API (psm_predict: pinned host rows in, host
             pressures out, H2D + D2H inside the timed region).
  roofline : dominant kernel's algorithmic bytes / its CUDA-event time vs MEASURED_PEAKS.json.
  cpu_baseline / --impl reference : the NumPy/SciPy oracle (the reference's own calls; TF Dense stack
             replaced by float32 NumPy because TensorFlow is not installable here) on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(REPO, 'solving-poisson-s-equation-through-dl-for-cfd-apllications_b200')
for p in (PKG, REPO):
    if p not in sys.path:
        sys.path.insert(0, p)

from psm_b200 import synthetic as syn, tables as ptables    # noqa: E402

HBM_FALLBACK_GBS = 6650.0      # /opt/skills/guides/B200_PROFILING.md fallback


# kernel measured by each HBM stage's CUDA events (key of profiles/traffic_*.json, written by profiles/summarize_full.py)
STAGE_KERNEL = {'gather': ('gather_extract_kernel', 'gather_kernel'), 'back_gather': ('back_kernel',), 'place': ('place_kernel',),
                'extract': ('extract_kernel',), 'prep': ('prep_kernel',)}


def measured_traffic(workload, stage):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the stage's kernel from the committed
    `ncu --set full` capture of this workload (profiles/traffic_<workload>_*.json, newest round), else None."""
    import glob
    files = sorted(glob.glob(os.path.join(REPO, 'profiles', 'traffic_%s_*.json' % workload)))
    if not files:
        return None, None
    try:
        d = json.load(open(files[-1]))
        for base in STAGE_KERNEL.get(stage, ()):
            for k in d:                                   # template instances are listed as name<args>
                if k == base or k.startswith(base + '<'):
                    return d[k]['dram_traffic_bytes'], os.path.basename(files[-1])
    except Exception:
        pass
    return None, None


def measured_peaks():
    try:
        with open(os.path.join(REPO, 'MEASURED_PEAKS.json')) as f:
            return json.load(f), 'measured'
    except Exception:
        return {'hbm_gbs': HBM_FALLBACK_GBS}, 'fallback'


def build_case(workload, variant, seed=0, back='closed_form'):
    """Synthetic mesh + fields + parameters + once-per-mesh tables (init is not part of a step)."""
    mesh = syn.make_mesh(seed=seed, **syn.CONFIGS[workload])
    F = syn.make_fields(mesh, seed=seed)
    deltas = variant == 'deltaU_to_deltaP'
    params = syn.make_params(seed=seed, pc_in=128, pc_p=128, standardization='std' if deltas else 'max_abs',
                             n_out_channels=1 if deltas else 2,
                             maxs=syn.DEFAULT_MAXS if deltas else (1.0, 0.536, 0.999, 0.8, 0.7))
    t0 = time.time()
    tables = ptables.build_tables(mesh['cells'], mesh['top'], mesh['obst'], F['p_prev'], variant=variant, back=back)
    return mesh, F, params, tables, time.time() - t0


def sharded_grid_rows(world, variant):
    """Weak scaling: ~1 M cells (1000 grid rows x 1000 columns) per GPU, stacked in y.  Rows are nudged so
    that (H - 128) is not a multiple of the stride -- the reference is undefined there."""
    stride = 96 if variant == 'deltaU_to_deltaP' else 32
    H = 1000 * world
    while (H - 128) % stride == 0:
        H += 8
    return H


def build_sharded_case(world, rank, variant, dist, mesh_kw=None, seed=0):
    """One rank's share of the N-GPU workload, built band-locally (psm_b200.shard.band_phase1..3): every
    rank triangulates only the cells around its own block rows; the two small global facts travel by
    all_gather_object.  Returns (mesh, fields, params, shard, seconds)."""
    from psm_b200 import shard as pshard
    if mesh_kw is None:
        H = sharded_grid_rows(world, variant)
        mesh_kw = dict(H=H, W=1000, nx=1000, ny=H, R=0.4)
    mesh = syn.make_mesh(seed=seed, **mesh_kw)
    F = syn.make_fields(mesh, seed=seed)
    deltas = variant == 'deltaU_to_deltaP'
    params = syn.make_params(seed=seed, pc_in=128, pc_p=128, standardization='std' if deltas else 'max_abs',
                             n_out_channels=1 if deltas else 2,
                             maxs=syn.DEFAULT_MAXS if deltas else (1.0, 0.536, 0.999, 0.8, 0.7))
    t0 = time.time()
    L, s1 = pshard.band_phase1(mesh['cells'], mesh['top'], mesh['obst'], F['p_prev'], rank, world, variant=variant)
    all1 = [None] * world
    dist.all_gather_object(all1, s1)
    L, s2 = pshard.band_phase2(L, all1)
    all2 = [None] * world
    dist.all_gather_object(all2, s2)
    sh = pshard.band_phase3(L, all2)
    return mesh, F, params, sh, time.time() - t0


def stage_bytes(n_cells, G, B, C, pc_in, pc_p, ncol, S=128, fused_extract=True):
    """Algorithmic bytes per step and stage (SURVEY.md section 8d / DESIGN.md): FP32 device storage,
    f64 only at the ABI.  Overlap re-reads and L2-resident intermediates are NOT counted twice.
    With the fused gather+extraction kernel (the default when W % 4 == 0) the block operand is written by the
    gather itself: its bytes move from 'extract' to 'gather'; the grid planes are neither written nor re-read."""
    S2 = S * S
    xu = 4 * B * 2 * S2
    return {
        # read rows, write float2 field + p_prev; 5-column mode also reads and rewrites the resident U(t-1)
        'prep': n_cells * (ncol * 8 + 8 + 8 + (32 if ncol == 5 else 0)),
        # tables + each cell value once + (fused: the block operand | else: the two grid planes)
        'gather': G * (12 + 12) + 8 * n_cells + (xu if fused_extract else 8 * G),
        'extract': 0 if fused_extract else 8 * G + xu,        # grid read once, operand written
        'pca_project': 4 * B * 2 * S2 + 4 * 2 * S2 * pc_in + 4 * B * pc_in,
        'mlp': 4 * (pc_in * 512 + 2 * 512 * 512 + 512 * pc_p) + 8 * B * pc_p,
        'pca_inverse': 4 * pc_p * S2 * C + 4 * B * S2 * C,
        'strip_means': 4 * B * S2 * C,
        'offsets': 0,
        'place': 4 * G * C + 4 * G * C + 2 * G,               # owner pixels read + field write + owner map
        'back_gather': n_cells * (12 + 12 + 12 * C + 8 + 8 * C),
    }


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.lines, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for ln in self.lines:
            f = [x.strip() for x in ln.split(',')]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


def oracle_for(variant, params, mesh, F, tables):
    from oracle.pipeline import DeltasOracle, GradPOracle, SurrogateParams
    P = SurrogateParams(**{k: v for k, v in params.items() if k != 'shape'})
    o = DeltasOracle(P) if variant == 'deltaU_to_deltaP' else GradPOracle(P)
    o.compute_only_once(mesh['cells'], mesh['top'], mesh['obst'], F['p_prev'],
                        tables=(tables['vert'], tables['weights'], tables['vert_back'], tables['weights_back']))
    return o


def oracle_step(o, variant, F):
    if variant == 'deltaU_to_deltaP':
        r = o.time_step(F['Ux'], F['Uy'], F['dUx'], F['dUy'])
        p, _ = o.to_cells(r['field'], F['p_prev'])
        return p
    r = o.time_step(F['Ux'], F['Uy'])
    return np.stack([o.to_cells(r['dp_dx']), o.to_cells(r['dp_dy'])], axis=1)


def time_oracle(o, variant, F, steps, warmup):
    for _ in range(warmup):
        oracle_step(o, variant, F)
    t0 = time.perf_counter()
    for _ in range(steps):
        oracle_step(o, variant, F)
    return (time.perf_counter() - t0) / max(steps, 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='native', choices=['native', 'reference'])
    ap.add_argument('--workload', default=None, help='c1 | c2 | c5 | c4 | tiny (default: c2)')
    ap.add_argument('--variant', default='deltaU_to_deltaP', choices=['deltaU_to_deltaP', 'U_to_gradP'])
    ap.add_argument('--cpu-steps', type=int, default=5, help='oracle steps for the cpu_baseline leg')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--input-cols', type=int, default=5, choices=[5, 7],
                    help='5: rows {Ux,Uy,Cx,Cy,p} exactly as FOAM/PythonComm.H:2-9 fills them, the handle keeps U(t-1) resident '
                         'and forms dU on the device (two alternating velocity fields are fed so that dU != 0 every step); '
                         '7: rows carry dUx,dUy as two extra columns')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    workload = args.workload or 'c2'
    variant = args.variant
    ncol = args.input_cols if variant == 'deltaU_to_deltaP' else 5
    scaling = 'weak'
    sharded_kw = None
    if world > 1 and args.impl == 'native' and args.workload:
        # a FIXED domain (e.g. c4 = BASELINE.json configs[3], 16 M cells) cut over the ranks: strong scaling
        scaling, sharded_kw = 'strong', dict(syn.CONFIGS[workload])
        cfg = {'workload': '%s: %s, synthetic mesh %s sharded by block rows over %d B200, NCCL halo / ghost / strip-mean '
                           'exchanges, pc_in=pc_p=128, MLP 3x512, random-init' % (workload, variant, sharded_kw, world),
               'l2': 'flushed between timed steps (256 MiB write, then 256 MiB read so that no dirty flush lines remain)', 'input_cols': ncol}
    elif world > 1 and args.impl == 'native':
        Hs = sharded_grid_rows(world, variant)
        cfg = {'workload': 'c2 x %d: %s, %d x 1000 grid (~1 M cells per GPU) sharded by block rows over %d B200, NCCL halo / '
                           'ghost / strip-mean exchanges, pc_in=pc_p=128, MLP 3x512, random-init' % (world, variant, Hs, world),
               'l2': 'flushed between timed steps (256 MiB write, then 256 MiB read so that no dirty flush lines remain)', 'input_cols': ncol}
    else:
        cfg = {'workload': '%s: %s, synthetic flow-past-cylinder mesh %s, pc_in=pc_p=128, MLP 3x512, random-init'
                           % (workload, variant, syn.CONFIGS[workload]),
               'l2': 'flushed between timed steps (256 MiB write, then 256 MiB read so that no dirty flush lines remain)', 'input_cols': ncol}

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == 'reference':
        if rank != 0:
            return 0
        import torch
        torch.set_num_threads(os.cpu_count() or 1)
        mesh, F, params, tables, t_init = build_case(workload, variant)
        o = oracle_for(variant, params, mesh, F, tables)
        steps, warm = min(args.steps, 10), min(args.warmup, 1)
        sec = time_oracle(o, variant, F, steps, warm)
        n = mesh['cells'].shape[0]
        v = n / sec
        line = {'impl': 'reference', 'metric': 'surrogate_cells_per_s', 'value': v, 'unit': 'cells/s',
                'n_gpus': args.gpus, 'steps': steps, 'warmup': warm, 'ms_per_step': sec * 1e3,
                'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64 (f32 MLP)',
                'data': 'synthetic', 'config': cfg,
                'cpu_baseline': {'value': v, 'unit': 'cells/s', 'cores': os.cpu_count(), 'kind': 'port',
                                 'sample': '%d full steps of the NumPy/SciPy oracle on the same %d-cell mesh '
                                           '(TF Dense stack replaced by float32 NumPy; init excluded)' % (steps, n)},
                'e2e': {'value': v, 'unit': 'cells/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ native arm (B200)
    import torch
    import psm_b200
    if not torch.cuda.is_available():
        raise SystemExit('bench.py --impl native needs a B200 (no CPU fallback exists)')
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    sm = psm_b200.PressureSurrogate(variant, device=local_rank, input_cols=ncol, timings=False)
    if world > 1:
        # ONE domain sharded by block rows (BASELINE.json configs[3] shape, sized for weak scaling)
        mesh, F, params, sh, t_init = build_sharded_case(world, rank, variant, dist, sharded_kw)
        ids = [psm_b200.PressureSurrogate.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        sm.load_params(params)
        sm.comm_init(ids[0], rank, world)
        sm.init_shard(sh)
        n = sh['n_owned']
        cells_np = np.ascontiguousarray(syn.pack_cells(mesh, F, with_delta=(ncol == 7))[sh['owned_ids']])
        tables = None
    else:
        mesh, F, params, tables, t_init = build_case(workload, variant)
        n = mesh['cells'].shape[0]
        sm.load_params(params)
        sm.init_tables(tables)
        cells_np = syn.pack_cells(mesh, F, with_delta=(ncol == 7))
    geo = sm.geometry()
    if world > 1:       # name the transport the handle actually selected
        cfg['workload'] = cfg['workload'].replace('NCCL halo / ghost / strip-mean exchanges',
                                                  'ghost-cell / strip-mean / ghost-pixel exchanges pushed over NVLink peer memory (cudaIpc)'
                                                  if geo.get('peer_memory_exchange') else 'NCCL ghost-cell / strip-mean / ghost-pixel exchanges')
    # two alternating time levels, U and U + dU: with 5 columns the device forms dU = +-(dU) itself (never zero,
    # so the reference's skip rule SMC:410-415 never short-cuts a timed step); with 7 columns both carry dU
    cells_b = cells_np.copy()
    own = sh['owned_ids'] if world > 1 else slice(None)
    cells_b[:, 0] += F['dUx'][own]
    cells_b[:, 1] += F['dUy'][own]
    # like the solver (FOAM/PythonComm_init.H:53) ONE input buffer lives for the whole run and is refilled in place
    # before every step -- the refill (the solver's forAll loop, FOAM/PythonComm.H:2-9) is outside the timed intervals
    h_src = [torch.from_numpy(cells_np), torch.from_numpy(cells_b)]
    h_in = torch.empty_like(h_src[0]).pin_memory()
    h_out = torch.empty(n if sm.n_fields == 1 else (n, 2), dtype=torch.float64).pin_memory()
    d_src = [x.cuda() for x in h_src]
    d_in = torch.empty_like(d_src[0])
    d_out = torch.empty_like(h_out, device='cuda')
    state = {'i': 0, 'skipped': 0}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
    flush_rd = torch.zeros(32 << 20, dtype=torch.int64, device='cuda')      # 256 MiB
    ext = torch.cuda.ExternalStream(sm.stream_ptr())
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run_device(k, timed):
        evs, tot = [], np.zeros(len(psm_b200._capi.TIMING_NAMES))
        with torch.cuda.stream(ext):
            for _ in range(k):
                state['i'] ^= 1
                d_in.copy_(d_src[state['i']])
                flush.zero_()            # evicts the previous step's lines (write 256 MiB) ...
                flush_rd.sum()           # ... then a 256 MiB read pass leaves L2 full of CLEAN foreign lines, so the
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(ext)           # timed step does not also pay for writing the flush buffer back
                sm.predict_device(d_in.data_ptr(), n, d_out.data_ptr(), sync=False)
                e1.record(ext)
                evs.append((e0, e1))
                if timed:
                    sm.synchronize()
                    tm = sm.timings()
                    tot += np.array([tm[k2] for k2 in psm_b200._capi.TIMING_NAMES])
        torch.cuda.synchronize()
        return sum(a.elapsed_time(b) for a, b in evs), tot

    run_device(max(args.warmup, 3), False)
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    ms_total, _ = run_device(args.steps, False)
    assert sm.synchronize() == 0, 'the last timed step was short-cut by the skip rule'
    barrier()
    launches = sm.launch_count() * args.steps
    sm.set_timings(True)                       # per-stage breakdown in a separate pass (events cost a few us)
    _, stage_ms = run_device(args.steps, True)
    sm.set_timings(False)
    # end to end through the host-buffer API
    def host_step():
        state['i'] ^= 1
        h_in.copy_(h_src[state['i']])
        if world > 1:
            dist.barrier()                       # collective call: all ranks enter together, refills not timed
        t0 = time.perf_counter()
        _, rc = sm.predict(h_in.numpy(), out=h_out.numpy())
        state['skipped'] += int(rc != 0)
        return time.perf_counter() - t0
    for _ in range(3):
        host_step()
    state['skipped'] = 0
    barrier()
    e2e_each = [host_step() for _ in range(args.steps)]
    e2e_s = sum(e2e_each)
    assert state['skipped'] == 0, 'a timed step was short-cut by the skip rule'
    # deltaU_to_deltaP falls back to p_prev (always finite); U_to_gradP keeps NaN where the reference's grid->cell
    # interpolation is NaN (cells outside the grid hull, GRAD has no previous-gradient fallback)
    assert np.isfinite(h_out.numpy()).mean() > (0.999 if variant == 'deltaU_to_deltaP' else 0.95)
    barrier()
    # where the end-to-end time goes (separate short pass with the per-stage events on: eager launches, not timed above)
    sm.set_timings(True)
    e2e_parts = np.zeros(3)
    for _ in range(5):
        host_step()
        tm = sm.timings()
        e2e_parts += np.array([tm['h2d'], sum(tm[k2] for k2 in psm_b200._capi.TIMING_NAMES[1:-1]), tm['d2h']])
    e2e_parts /= 5
    sm.set_timings(False)
    barrier()
    clocks = sampler.stop()
    if world > 1:
        t = torch.tensor([ms_total, e2e_s], dtype=torch.float64, device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total, e2e_s = t.tolist()
        nt = torch.tensor([float(n)], dtype=torch.float64, device='cuda')
        dist.all_reduce(nt)
        n_total = nt.item()
    else:
        n_total = float(n)
    ms_step = ms_total / args.steps
    value = n_total / (ms_step * 1e-3)
    e2e = n_total / (e2e_s / args.steps)

    if rank == 0:
        peaks, peak_src = measured_peaks()
        fused = geo['grid_w'] % 4 == 0 and not os.environ.get('PSM_NO_FUSED_EXTRACT')
        sb = stage_bytes(n, (geo['row1'] - geo['row0']) * geo['grid_w'], geo['n_local_blocks'], sm.n_fields, sm.pc_in, sm.pc_p, ncol,
                         fused_extract=fused)
        stage_avg = dict(zip(psm_b200._capi.TIMING_NAMES, (stage_ms / args.steps).tolist()))
        stages = {k: {'ms': stage_avg[k], 'GBps': (sb[k] / (stage_avg[k] * 1e-3) / 1e9) if stage_avg.get(k, 0) > 0 else None}
                  for k in sb}
        hbm_stages = ['gather', 'back_gather', 'place', 'extract', 'prep']
        dom = max(hbm_stages, key=lambda k: stage_avg[k])
        ach = sb[dom] / (stage_avg[dom] * 1e-3) / 1e9
        traffic, traffic_src = measured_traffic(workload if world == 1 else 'n%d' % world, dom)
        roof = {'kernel': dom, 'bound': 'hbm', 'achieved': ach, 'peak': peaks['hbm_gbs'], 'unit': 'GB/s',
                'frac': ach / peaks['hbm_gbs'], 'traffic': traffic, 'traffic_source': traffic_src, 'peak_source': peak_src,
                'algorithmic_bytes_per_launch': sb[dom]}
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            torch.set_num_threads(os.cpu_count() or 1)
            o = oracle_for(variant, params, mesh, F, tables)
            sec = time_oracle(o, variant, F, args.cpu_steps, 1)
            cpu = {'value': n / sec, 'unit': 'cells/s', 'cores': os.cpu_count(), 'kind': 'port',
                   'sample': '%d full steps of the NumPy/SciPy oracle on the same %d-cell mesh (TF Dense stack '
                             'replaced by float32 NumPy; init excluded)' % (args.cpu_steps, n), 'ms_per_step': sec * 1e3}
        line = {'metric': 'surrogate_cells_per_s', 'value': value, 'unit': 'cells/s', 'n_gpus': world,
                'steps': args.steps, 'warmup': max(args.warmup, 3), 'ms_per_step': ms_step, 'higher_is_better': True,
                'scaling': scaling, 'vs_baseline': None, 'dtype': 'f32 (f64 at the ABI and in the offset chain)',
                'data': 'synthetic', 'config': cfg, 'clocks': clocks,
                'e2e': {'value': e2e, 'unit': 'cells/s', 'h2d_bytes_per_step': int(n * ncol * 8),
                        'd2h_bytes_per_step': int(n * sm.n_fields * 8), 'ms_per_step': e2e_s / args.steps * 1e3,
                        'p50_ms': float(np.percentile(e2e_each, 50) * 1e3), 'p99_ms': float(np.percentile(e2e_each, 99) * 1e3),
                        'device_events_ms': {'h2d': float(e2e_parts[0]), 'kernels': float(e2e_parts[1]), 'd2h': float(e2e_parts[2])}},
                'gpu_launches': launches, 'roofline': roof, 'cpu_baseline': cpu, 'stages': stages,
                'geometry': geo, 'init_tables_s': t_init, 'n_cells_total': n_total}
        print(json.dumps(line))
    sm.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == '__main__':
    sys.exit(main())
