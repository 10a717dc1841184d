#!/usr/bin/env python
"""Benchmark of the pressure-surrogate hot path (BASELINE.json metric: mesh cells/s, ms per step).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference] [--workload c2|c5|c1|c4|c2xN]

One "step" = one per-timestep surrogate prediction (``py_func`` in the reference) over the whole synthetic mesh.
  N = 1  : BASELINE.json configs[1] (c2): deltaU_to_deltaP at ~1 M cells on a 1000 x 1000 grid, random-init weights of the
           reference architecture (pc 128 -> 3 x 512 -> pc 128).
  N > 1  : BASELINE.json configs[3] (c4): ONE 4000 x 4000 domain (~16 M cells) cut by block rows over the N GPUs -- strong
           scaling.  ``--workload c2xN`` keeps round 1's weak-scaling domain (~1 M cells per GPU).
  value    : cells/s with the solver's rows already resident in HBM (psm_predict_device), CUDA events on the handle's
             stream, L2 flushed between steps.
  e2e      : cells/s through the host-buffer entry point a solver adapter calls, psm_predict_fields: the solver's own
             U (double[n][3], OpenFOAM `vector`) and p (double[n]) arrays in pinned host memory in, pressures out, H2D + D2H
             inside the timed region.  `e2e.rows5` is the same through the drop-in double[n][5] row layout (psm_predict).
  roofline : dominant HBM kernel's algorithmic bytes / its CUDA-event time vs MEASURED_PEAKS.json; `roofline_c5` the same
             kernels on the 4 M-cell mesh (working set > L2); `gemm` the achieved TFLOP/s of the three contractions.
  parity   : N > 1 only: every rank's pressures against a single-GPU handle on the same mesh and against the CPU oracle.
  cpu_baseline / --impl reference : the NumPy/SciPy oracle (the reference's own calls; TF Dense stack replaced by float32
             NumPy because TensorFlow is not installable here) on the host cores, on a bounded sample for the large meshes.
  --latency: BASELINE.json configs[4]: per-step latency (mean / p50 / p99) over --steps consecutive steps, a new velocity
             field every step, tables and weights resident.
"""
import argparse
import glob
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
BF16_FALLBACK_TFLOPS = 1590.0
L2_NOTE = 'flushed between timed steps (256 MiB write, then 256 MiB read so that no dirty flush lines remain)'

# kernel measured by each HBM stage's CUDA events (key of profiles/traffic_*.json, written by profiles/summarize_full.py)
STAGE_KERNEL = {'gather': ('gather_extract_kernel', 'gather_kernel'), 'back_gather': ('back4_kernel', 'back_kernel'),
                'place': ('place_kernel',), 'extract': ('extract_kernel',), 'prep': ('prep_bulk_kernel', 'prep_kernel')}
HBM_STAGES = ['gather', 'back_gather', 'place', 'extract', 'prep']


def measured_traffic(workload, stage):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the stage's kernel from THIS round's committed
    `ncu --set full` capture of this workload (profiles/traffic_<workload>_r2*.json, newest), else None: a capture of an
    older kernel is not a measurement of this run."""
    files = sorted(glob.glob(os.path.join(REPO, 'profiles', 'traffic_%s_r2*.json' % workload)))
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
        return {'hbm_gbs': HBM_FALLBACK_GBS, 'bf16_tflops': BF16_FALLBACK_TFLOPS}, 'fallback'


def make_params_for(variant, seed=0):
    deltas = variant == 'deltaU_to_deltaP'
    return syn.make_params(seed=seed, pc_in=128, pc_p=128, standardization='std' if deltas else 'max_abs',
                           n_out_channels=1 if deltas else 2,
                           maxs=syn.DEFAULT_MAXS if deltas else (1.0, 0.536, 0.999, 0.8, 0.7))


def build_case(mesh_kw, variant, seed=0, back='closed_form'):
    """Synthetic mesh + fields + parameters + once-per-mesh tables (init is not part of a step)."""
    mesh = syn.make_mesh(seed=seed, **mesh_kw)
    F = syn.make_fields(mesh, seed=seed)
    params = make_params_for(variant, seed)
    t0 = time.time()
    tables = ptables.build_tables(mesh['cells'], mesh['top'], mesh['obst'], F['p_prev'], variant=variant, back=back)
    return mesh, F, params, tables, time.time() - t0


def weak_mesh_kw(world, variant):
    """Weak scaling (--workload c2xN): ~1 M cells (1000 grid rows x 1000 columns) per GPU, stacked in y.  Rows are nudged so
    that (H - 128) is not a multiple of the stride -- the reference is undefined there."""
    stride = 96 if variant == 'deltaU_to_deltaP' else 32
    H = 1000 * world
    while (H - 128) % stride == 0:
        H += 8
    return dict(H=H, W=1000, nx=1000, ny=H, R=0.4)


def resolve_workload(args, world):
    """-> (name, mesh_kw, scaling).  N = 1: c2.  N > 1: c4 (BASELINE configs[3], strong); c2xN: weak."""
    name = args.workload or ('c2' if world == 1 else 'c4')
    if name in ('c2xN', 'c2xn', 'weak'):
        return 'c2x%d' % world, weak_mesh_kw(world, args.variant), 'weak'
    if name not in syn.CONFIGS:
        raise SystemExit('unknown workload %r (have %s, c2xN)' % (name, sorted(syn.CONFIGS)))
    return name, dict(syn.CONFIGS[name]), ('strong' if world > 1 else 'weak')


def config_dict(name, mesh_kw, variant, world, ncol):
    """The `config` object: IDENTICAL in the native and the reference arm for the same command line."""
    if world > 1:
        wl = ('%s: %s, synthetic flow-past-cylinder mesh %s as ONE domain sharded by block rows over %d B200 (ghost cells / strip '
              'means / ghost pixels exchanged between neighbouring ranks), pc_in=pc_p=128, MLP 3x512, random-init'
              % (name, variant, mesh_kw, world))
    else:
        wl = '%s: %s, synthetic flow-past-cylinder mesh %s, pc_in=pc_p=128, MLP 3x512, random-init' % (name, variant, mesh_kw)
    return {'workload': wl, 'l2': L2_NOTE, 'input_cols': ncol,
            'tables': "cells -> grid: SciPy Qhull (as the reference); grid -> cell: closed form on the regular grid (back='closed_form'), "
                      "the same tables in the GPU path and the CPU oracle"}


def build_sharded_case(world, rank, variant, dist, mesh_kw, seed=0):
    """One rank's share of the N-GPU workload, built band-locally (psm_b200.shard.band_phase1..3): every
    rank triangulates only the cells around its own block rows; the two small global facts travel by
    all_gather_object.  Returns (mesh, fields, params, shard, band, seconds)."""
    from psm_b200 import shard as pshard
    mesh = syn.make_mesh(seed=seed, **mesh_kw)
    F = syn.make_fields(mesh, seed=seed)
    params = make_params_for(variant, seed)
    t0 = time.time()
    L, s1 = pshard.band_phase1(mesh['cells'], mesh['top'], mesh['obst'], F['p_prev'], rank, world, variant=variant, keep_raw=True)
    all1 = [None] * world
    dist.all_gather_object(all1, s1)
    L, s2 = pshard.band_phase2(L, all1)
    all2 = [None] * world
    dist.all_gather_object(all2, s2)
    sh = pshard.band_phase3(L, all2)
    return mesh, F, params, sh, L, time.time() - t0


def stage_bytes(n_cells, G, B, C, pc_in, pc_p, ncol, S=128, fused_extract=True, from_blocks=True, grid_a=False):
    """Algorithmic bytes per step and stage (SURVEY.md section 8d / DESIGN.md): FP32 device storage,
    f64 only at the ABI.  Overlap re-reads and L2-resident intermediates are NOT counted twice.
    With the fused gather+extraction kernel (the default when W % 4 == 0) the block operand is written by the
    gather itself: its bytes move from 'extract' to 'gather'; the grid planes are neither written nor re-read.
    grid_a (single-GPU default): the gather writes the two grid planes and the projection fetches its A tiles from them by TMA --
    no block operand exists; the projection's unique bytes are the planes, not B overlapping copies of them."""
    S2 = S * S
    xu = 4 * B * 2 * S2
    if grid_a:
        fused_extract = False
    return {
        # read rows, write float2 field + p_prev; 5-column mode also reads and rewrites the resident U(t-1)
        'prep': n_cells * (ncol * 8 + 8 + 8 + (32 if ncol == 5 else 0)),
        # tables + each cell value once + (fused: the block operand | else: the two grid planes)
        'gather': G * (12 + 12) + 8 * n_cells + (xu if fused_extract else 8 * G),
        'extract': 0 if (fused_extract or grid_a) else 8 * G + xu,        # grid read once, operand written
        'pca_project': (8 * G if grid_a else 4 * B * 2 * S2) + 4 * 2 * S2 * pc_in + 4 * B * pc_in,
        'mlp': 4 * (pc_in * 512 + 2 * 512 * 512 + 512 * pc_p) + 8 * B * pc_p,
        'pca_inverse': 4 * pc_p * S2 * C + 4 * B * S2 * C,
        'strip_means': 0,                                      # row partials come out of the PCA-inverse epilogue
        'offsets': 0,
        'place': 4 * G * C + 4 * G * C + 2 * G,               # owner pixels read + field write + owner map
        # tables (12 idx + 12 w [+ 6 owner ids]) + 3 gathered pixels per field + p_prev + out
        'back_gather': n_cells * (12 + 12 + (6 if from_blocks else 0) + 12 * C + 8 + 8 * C),
    }


def stage_flops(B, C, pc_in, pc_p, S=128):
    """Algorithmic FLOPs of the three dense contractions: the distance third of the projection is folded at init."""
    S2 = S * S
    return {'pca_project': 2.0 * B * 2 * S2 * pc_in, 'mlp': 2.0 * B * (pc_in * 512 + 2 * 512 * 512 + 512 * pc_p),
            'pca_inverse': 2.0 * B * pc_p * S2 * C}


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


def reference_sample_kw(name, mesh_kw, world):
    """Bounded sample of the workload for the CPU arm: the whole mesh up to ~1.2 M cells; for the larger domains a band of
    512 grid rows over the full width (what one rank of an 8-way split of c4 holds), or c2 for the weak c2xN domain."""
    cells = mesh_kw['nx'] * mesh_kw['ny']
    if cells <= 1_300_000:
        return dict(mesh_kw), 'the whole mesh'
    if name.startswith('c2x'):
        return dict(syn.CONFIGS['c2']), "one rank's share of the weak-scaling domain (the c2 mesh, ~1 M cells)"
    H = 512      # ~2 M cells: table building (Qhull) + a few oracle steps stay within a few minutes on the box's host cores
    kw = dict(mesh_kw, H=H, ny=int(round(mesh_kw['ny'] * H / mesh_kw['H'])))
    return kw, 'a band of %d of the %d grid rows over the full width (%d x %d cell lattice)' % (H, mesh_kw['H'], kw['nx'], kw['ny'])


def run_reference(args, name, mesh_kw, variant, cfg):
    """--impl reference: the oracle port on the host cores, bounded sample, same config / metric / unit."""
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    kw, what = reference_sample_kw(name, mesh_kw, args.gpus)
    mesh, F, params, tables, _ = build_case(kw, variant)
    o = oracle_for(variant, params, mesh, F, tables)
    steps, warm = max(1, min(args.steps, 10 if mesh['cells'].shape[0] < 2_000_000 else 3)), min(args.warmup, 1)
    sec = time_oracle(o, variant, F, steps, warm)
    n = mesh['cells'].shape[0]
    v = n / sec
    line = {'impl': 'reference', 'metric': 'surrogate_cells_per_s', 'value': v, 'unit': 'cells/s',
            'n_gpus': args.gpus, 'steps': steps, 'warmup': warm, 'ms_per_step': sec * 1e3,
            'higher_is_better': True, 'scaling': 'strong' if (args.gpus > 1 and not name.startswith('c2x')) else 'weak',
            'vs_baseline': None, 'dtype': 'f64 (f32 MLP)', 'data': 'synthetic', 'config': cfg,
            'cpu_baseline': {'value': v, 'unit': 'cells/s', 'cores': os.cpu_count(), 'kind': 'port',
                             'sample': '%d full steps of the NumPy/SciPy oracle on %s: %d cells (TF Dense stack replaced by '
                                       'float32 NumPy; init excluded)' % (steps, what, n)},
            'e2e': {'value': v, 'unit': 'cells/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line))
    return 0


class DeviceCase:
    """A handle + its buffers + the timing loops, for one mesh on this rank."""

    def __init__(self, sm, torch, dist, world, n, cells_np, F_own, ncol, variant):
        self.sm, self.torch, self.dist, self.world, self.n, self.ncol, self.variant = sm, torch, dist, world, n, ncol, variant
        # two alternating time levels, U and U + dU: with 5 columns the device forms dU = +-(dU) itself (never zero,
        # so the reference's skip rule SMC:410-415 never short-cuts a timed step); with 7 columns both carry dU
        cells_b = cells_np.copy()
        cells_b[:, 0] += F_own['dUx']
        cells_b[:, 1] += F_own['dUy']
        # like the solver (FOAM/PythonComm_init.H:53) ONE input buffer lives for the whole run and is refilled in place
        # before every step -- the refill (the solver's forAll loop, FOAM/PythonComm.H:2-9) is outside the timed intervals
        self.h_src = [torch.from_numpy(cells_np), torch.from_numpy(cells_b)]
        self.h_in = torch.empty_like(self.h_src[0]).pin_memory()
        nf = sm.n_fields
        self.h_out = torch.empty(n if nf == 1 else (n, 2), dtype=torch.float64).pin_memory()
        # the solver's native arrays: U as double[n][3] (OpenFOAM vector), p as double[n]
        self.hU_src = []
        for c in (cells_np, cells_b):
            U3 = np.zeros((n, 3))
            U3[:, 0], U3[:, 1] = c[:, 0], c[:, 1]
            self.hU_src.append(torch.from_numpy(U3))
        self.hU = torch.empty_like(self.hU_src[0]).pin_memory()
        self.hp = torch.from_numpy(np.ascontiguousarray(cells_np[:, 4])).pin_memory()
        self.hdU = None
        if ncol == 7:
            dU3 = np.zeros((n, 3))
            dU3[:, 0], dU3[:, 1] = cells_np[:, 5], cells_np[:, 6]
            self.hdU = torch.from_numpy(dU3).pin_memory()
        self.d_src = [x.cuda() for x in self.h_src]
        self.d_in = torch.empty_like(self.d_src[0])
        self.d_out = torch.empty_like(self.h_out, device='cuda')
        self.state = {'i': 0, 'skipped': 0}
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
        self.flush_rd = torch.zeros(32 << 20, dtype=torch.int64, device='cuda')      # 256 MiB
        self.ext = torch.cuda.ExternalStream(sm.stream_ptr())
        self.sync_t = torch.zeros(1, device='cuda')
        torch.cuda.synchronize()

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def run_device(self, k, timed, flush=True):
        import psm_b200
        torch, sm, st = self.torch, self.sm, self.state
        names = psm_b200._capi.TIMING_NAMES
        evs, tot = [], np.zeros(len(names))
        with torch.cuda.stream(self.ext):
            for _ in range(k):
                st['i'] ^= 1
                self.d_in.copy_(self.d_src[st['i']])
                if flush:
                    self.flush.zero_()       # evicts the previous step's lines (write 256 MiB) ...
                    self.flush_rd.sum()      # ... then a 256 MiB read pass leaves L2 full of CLEAN foreign lines, so the
                if self.world > 1 and flush:
                    # the ranks' flush kernels end at different times: a one-element all-reduce in stream order lines the ranks up
                    # again before the timed interval opens (the barrier the timing contract brackets a step with), so a step is not
                    # charged for its neighbours' flush
                    self.dist.all_reduce(self.sync_t)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(self.ext)          # timed step does not also pay for writing the flush buffer back
                sm.predict_device(self.d_in.data_ptr(), self.n, self.d_out.data_ptr(), sync=False)
                e1.record(self.ext)
                evs.append((e0, e1))
                if timed:
                    sm.synchronize()
                    tm = sm.timings()
                    tot += np.array([tm[k2] for k2 in names])
        torch.cuda.synchronize()
        each = [a.elapsed_time(b) for a, b in evs]
        return sum(each), tot, each

    def host_step(self, mode):
        """One end-to-end step through a host entry point; the refill of the solver's buffer is outside the timed interval (the
        refilled lines are still dirty in the CPU caches when the DMA reads them, as in the solver: profiles/pcie_probe.py
        measures 0.555 ms for U that way against 0.445 ms from DRAM)."""
        st, sm = self.state, self.sm
        st['i'] ^= 1
        if mode == 'fields':
            self.hU.copy_(self.hU_src[st['i']])
        else:
            self.h_in.copy_(self.h_src[st['i']])
        if self.world > 1:
            self.dist.barrier()                   # collective call: all ranks enter together, refills not timed
        t0 = time.perf_counter()
        if mode == 'fields':
            _, rc = sm.predict_fields(self.hU.numpy(), p=self.hp.numpy(), dU=None if self.hdU is None else self.hdU.numpy(),
                                      out=self.h_out.numpy())
        else:
            _, rc = sm.predict(self.h_in.numpy(), out=self.h_out.numpy())
        dt = time.perf_counter() - t0
        st['skipped'] += int(rc != 0)
        return dt

    def time_host(self, mode, steps):
        for _ in range(3):
            self.host_step(mode)
        self.state['skipped'] = 0
        self.barrier()
        each = [self.host_step(mode) for _ in range(steps)]
        assert self.state['skipped'] == 0, 'a timed step was short-cut by the skip rule'
        return each

    def host_parts(self, mode, reps=5):
        """Where the end-to-end time goes: separate short pass with the per-stage events on (eager, not part of the timing)."""
        import psm_b200
        names = psm_b200._capi.TIMING_NAMES
        self.sm.set_timings(True)
        parts = np.zeros(3)
        for _ in range(reps):
            self.host_step(mode)
            tm = self.sm.timings()
            parts += np.array([tm['h2d'], sum(tm[k2] for k2 in names[1:-1]), tm['d2h']])
        self.sm.set_timings(False)
        return parts / reps


def stage_report(sb, stage_avg):
    """Per-stage times of the eager, event-instrumented pass.  A stage is bracketed by two event records which break the
    programmatic overlap between kernels; the stages that launch nothing (extract / offsets / place on the default path)
    measure exactly that overhead, which is subtracted before a bandwidth is quoted (`ms` stays the raw figure)."""
    empty = [stage_avg[k] for k in ('extract', 'offsets', 'place') if sb.get(k, 0) == 0 and stage_avg.get(k, 0) > 0]
    ovh = min(empty) if empty else 0.0
    out = {}
    for k in sb:
        ms = stage_avg.get(k, 0.0)
        net = max(ms - ovh, 1e-6)
        out[k] = {'ms': ms, 'ms_net': net if ms > 0 else 0.0, 'GBps': (sb[k] / (net * 1e-3) / 1e9) if (ms > 0 and sb[k] > 0) else None}
    return out, ovh


def roofline_block(sb, stages, peaks, peak_src, workload, live_traffic=True):
    dom = max(HBM_STAGES, key=lambda k: stages[k]['ms_net'] if sb[k] > 0 else 0.0)
    ach = stages[dom]['GBps']
    traffic, traffic_src = measured_traffic(workload, dom) if live_traffic else (None, None)
    return {'kernel': dom, 'bound': 'hbm', 'achieved': ach, 'peak': peaks['hbm_gbs'], 'unit': 'GB/s',
            'frac': ach / peaks['hbm_gbs'], 'traffic': traffic, 'traffic_source': traffic_src, 'peak_source': peak_src,
            'algorithmic_bytes_per_launch': sb[dom],
            'hbm_kernels': {k: {'GBps': stages[k]['GBps'], 'frac': stages[k]['GBps'] / peaks['hbm_gbs']}
                            for k in HBM_STAGES if sb[k] > 0 and stages[k]['GBps'] and stages[k]['ms_net'] > 1e-3}}   # launched stages only


def single_gpu_measure(torch, psm_b200, args, name, mesh_kw, variant, ncol, local_rank, cpu=True, latency=False):
    """Everything measured on one mesh with one GPU; returns the pieces of the JSON line."""
    mesh, F, params, tables, t_init = build_case(mesh_kw, variant)
    n = mesh['cells'].shape[0]
    sm = psm_b200.PressureSurrogate(variant, device=local_rank, input_cols=ncol, timings=False)
    sm.load_params(params)
    sm.init_tables(tables)
    geo = sm.geometry()
    cells_np = syn.pack_cells(mesh, F, with_delta=(ncol == 7))
    dc = DeviceCase(sm, torch, None, 1, n, cells_np, F, ncol, variant)
    dc.run_device(max(args.warmup, 3), False)
    dc.barrier()
    ms_total, _, each = dc.run_device(args.steps, False)
    assert sm.synchronize() == 0, 'the last timed step was short-cut by the skip rule'
    launches = sm.launch_count() * args.steps
    sm.set_timings(True)                       # per-stage breakdown in a separate pass
    _, stage_ms, _ = dc.run_device(args.steps, True)
    sm.set_timings(False)
    res = dict(n=n, geo=geo, t_init=t_init, ms_step=ms_total / args.steps, launches=launches, sm=sm, dc=dc,
               mesh=mesh, F=F, params=params, tables=tables)
    fused = geo['grid_w'] % 4 == 0 and not os.environ.get('PSM_NO_FUSED_EXTRACT')
    grid_a = fused and not os.environ.get('PSM_NO_GRID_A') and sm.pc_in <= 128
    sb = stage_bytes(n, geo['grid_h'] * geo['grid_w'], geo['n_blocks'], sm.n_fields, sm.pc_in, sm.pc_p, ncol, fused_extract=fused, grid_a=grid_a)
    stage_avg = dict(zip(psm_b200._capi.TIMING_NAMES, (stage_ms / args.steps).tolist()))
    res['stages'], res['event_overhead_ms'] = stage_report(sb, stage_avg)
    res['sb'] = sb
    fl = stage_flops(geo['n_blocks'], sm.n_fields, sm.pc_in, sm.pc_p)
    res['flops'] = fl
    return res


def gemm_block(res, peaks):
    """Achieved TFLOP/s of the three contractions (algorithmic FLOPs / CUDA-event time) against the TF32 tensor peak, taken
    as half the measured dense bf16 figure (no separate TF32 measurement exists); 3xTF32 issues 3 MMAs per algorithmic one."""
    tf32_peak = 0.5 * peaks.get('bf16_tflops', BF16_FALLBACK_TFLOPS)
    out = {}
    for k, f in res['flops'].items():
        ms = res['stages'][k]['ms_net']
        tf = f / (ms * 1e-3) / 1e12 if ms > 0 else None
        out[k] = {'ms': res['stages'][k]['ms'], 'tflops': tf, 'frac_of_tf32_peak': (tf / tf32_peak) if tf else None,
                  'GBps': res['stages'][k]['GBps']}
    # NOT measured by this run: tensor-pipe activity of the same kernels in the round's ncu capture (cold-cache, serialised replay)
    out['tensor_pipe_pct_ncu'] = {'pca_project': 25.5, 'mlp': [3.4, 6.3, 6.3, 6.4], 'pca_inverse': 20.6, 'pca_project_c3': 50.1,
                                  'pca_inverse_c3': 45.4, 'source': 'profiles/r2m_full_summary.txt, profiles/r2m_full_summary_c3.txt '
                                                                    '(sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active, ncu --set full, c2 / c3 step)'}
    out['note'] = ('algorithmic FLOPs (3xTF32 runs 3 tensor passes per algorithmic one); TF32 peak taken as 0.5 x measured bf16 '
                   '(%.0f TFLOP/s); at 121-441 rows these contractions are latency / HBM bound, tensor-pipe utilisation from ncu is '
                   'in profiles/' % tf32_peak)
    return out


def latency_run(torch, args, res):
    """BASELINE.json configs[4]: --steps consecutive pressure steps on one fixed geometry through the C-ABI, a new velocity
    field every step (8 pre-generated fields cycled: dU != 0 each step), tables and weights resident, no L2 flush."""
    sm, dc, n = res['sm'], res['dc'], res['n']
    mesh = res['mesh']
    K = 8
    Us, dUs = [], []
    for k in range(K):
        Fk = syn.make_fields(mesh, seed=100 + k)
        U3 = np.zeros((n, 3))
        U3[:, 0], U3[:, 1] = Fk['Ux'], Fk['Uy']
        Us.append(torch.from_numpy(U3).pin_memory())
    dU_d = [u.cuda() for u in Us]
    p_d = dc.hp.cuda()
    out_d = torch.empty(n, dtype=torch.float64, device='cuda') if sm.n_fields == 1 else torch.empty((n, 2), dtype=torch.float64, device='cuda')
    host, dev = [], []
    for it in range(args.steps + 5):
        u = Us[it % K]
        t0 = time.perf_counter()
        sm.predict_fields(u.numpy(), p=dc.hp.numpy(), out=dc.h_out.numpy())
        host.append(time.perf_counter() - t0)
    torch.cuda.synchronize()
    for it in range(args.steps + 5):
        t0 = time.perf_counter()
        sm.predict_fields_device(dU_d[it % K].data_ptr(), 3, n, out_d.data_ptr(), d_p_ptr=p_d.data_ptr(), sync=True)
        dev.append(time.perf_counter() - t0)
    host, dev = np.array(host[5:]) * 1e3, np.array(dev[5:]) * 1e3
    q = lambda a: {'mean_ms': float(a.mean()), 'p50_ms': float(np.percentile(a, 50)), 'p99_ms': float(np.percentile(a, 99)),   # noqa: E731
                   'max_ms': float(a.max())}
    return {'steps': args.steps, 'host_fields': q(host), 'device_fields': q(dev),
            'note': 'wall clock around each call incl. the final stream synchronise; host: psm_predict_fields (pinned U[n][3] + p[n] in, '
                    'p out); device: psm_predict_fields_device (sync=1); consecutive steps, no L2 flush, new U every step'}


def multi_parity(torch, psm_b200, dist, rank, world, variant, ncol, local_rank, mesh, F, params, sh, L, out_own, tag):
    """After the timed runs: every rank's pressures against (a) a single-GPU handle on the SAME mesh, built on rank 0 from the
    ranks' own band tables (so also the band-local table builder is checked against one global table set) and (b) the CPU
    oracle run on rank 0.  The other ranks wait on a file flag (no collective: the oracle takes minutes on 16 M cells)."""
    base = '/dev/shm/psm_bench_%s_' % tag
    np.savez(base + 'r%d.npz' % rank, owned=sh['owned_ids'], out=out_own, row0=L['row0'], row1=L['row1'],
             vert=L['raw_vert'], weights=L['raw_weights'])
    dist.barrier()
    res = None
    if rank == 0:
        try:
            n = mesh['cells'].shape[0]
            H, W = sh['H'], sh['W']
            vert = np.zeros((H * W, 3), np.int32)
            wts = np.zeros((H * W, 3), np.float64)
            full = np.full(n if variant == 'deltaU_to_deltaP' else (n, 2), np.nan)
            for r in range(world):
                z = np.load(base + 'r%d.npz' % r)
                vert[int(z['row0']) * W:int(z['row1']) * W] = z['vert']
                wts[int(z['row0']) * W:int(z['row1']) * W] = z['weights']
                full[z['owned']] = z['out']
            t = ptables.build_tables(mesh['cells'], mesh['top'], mesh['obst'], F['p_prev'], variant=variant, back='closed_form',
                                     precomputed=(vert, wts))
            x_min, x_max, y_min, y_max = t['bbox']
            X0, Y0 = ptables.uniform_grid(x_min, x_max, y_min, y_max, t['delta'])
            vb, wb = ptables.regular_grid_back_tables(mesh['cells'], X0[:W], Y0[::W], W)
            t['vert_back'], t['weights_back'] = vb, wb
            cells = syn.pack_cells(mesh, F, with_delta=(ncol == 7))
            with psm_b200.PressureSurrogate(variant, device=local_rank, input_cols=ncol) as one:
                one.load_params(params)
                one.init_tables(t)
                if ncol == 5 and variant == 'deltaU_to_deltaP':          # same two time levels, same order as the last timed host step
                    prev = cells.copy()
                    prev[:, 0] += F['dUx']
                    prev[:, 1] += F['dUy']
                    one.predict(prev)
                single, rc1 = one.predict(cells)
            rel = lambda a, b: float(np.linalg.norm(np.nan_to_num(a - b)) / max(np.linalg.norm(np.nan_to_num(b)), 1e-300))   # noqa: E731
            pp = F['p_prev'] if variant == 'deltaU_to_deltaP' else 0.0
            ppc = pp if variant == 'deltaU_to_deltaP' else 0.0
            res = {'vs_single_gpu': rel(full - ppc, single - ppc), 'single_gpu_status': int(rc1),
                   'nan_pattern_equal': bool(np.array_equal(np.isnan(full), np.isnan(single)))}
            if not os.environ.get('PSM_BENCH_NO_ORACLE'):
                t0 = time.time()
                o = oracle_for(variant, params, mesh, F, t)
                Fo = dict(F)
                if ncol == 5 and variant == 'deltaU_to_deltaP':
                    Fo['dUx'], Fo['dUy'] = -F['dUx'], -F['dUy']          # the last step went from U + dU back to U
                p_ref = oracle_step(o, variant, Fo)
                res['vs_oracle'] = rel(full - ppc, p_ref - ppc)
                res['oracle_s'] = time.time() - t0
            res['bounds'] = {'vs_single_gpu': 1e-5, 'vs_oracle': 1e-3}
            res['ok'] = bool(res['vs_single_gpu'] <= 1e-5 and res.get('vs_oracle', 0.0) <= 1e-3)
        except Exception as e:       # the bench line must still come out; the failure is recorded in it
            import traceback
            res = {'ok': False, 'error': '%s: %s' % (type(e).__name__, e), 'trace': traceback.format_exc()[-1500:]}
        open(base + 'done', 'w').write('1')
    else:
        while not os.path.exists(base + 'done'):
            time.sleep(0.5)
    dist.barrier()
    if rank == 0:
        for f in glob.glob(base + '*'):
            try:
                os.remove(f)
            except OSError:
                pass
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='native', choices=['native', 'reference'])
    ap.add_argument('--workload', default=None, help='c1 | c2 | c5 | c4 | c2xN | tiny (default: c2 on one GPU, c4 sharded on several)')
    ap.add_argument('--variant', default='deltaU_to_deltaP', choices=['deltaU_to_deltaP', 'U_to_gradP'])
    ap.add_argument('--cpu-steps', type=int, default=5, help='oracle steps for the cpu_baseline leg')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-c5', action='store_true', help='skip the roofline_c5 block (same kernels on the 4 M-cell mesh)')
    ap.add_argument('--no-parity', action='store_true', help='N > 1: skip the parity check after the timed runs')
    ap.add_argument('--latency', action='store_true', help='configs[4]: per-step latency over --steps consecutive steps')
    ap.add_argument('--input-cols', type=int, default=5, choices=[5, 7],
                    help='5: rows {Ux,Uy,Cx,Cy,p} exactly as FOAM/PythonComm.H:2-9 fills them, the handle keeps U(t-1) resident '
                         'and forms dU on the device (two alternating velocity fields are fed so that dU != 0 every step); '
                         '7: rows carry dUx,dUy as two extra columns')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    variant = args.variant
    ncol = args.input_cols if variant == 'deltaU_to_deltaP' else 5
    name, mesh_kw, scaling = resolve_workload(args, max(world, args.gpus))
    cfg = config_dict(name, mesh_kw, variant, max(world, args.gpus), ncol)

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == 'reference':
        if rank != 0:
            return 0
        return run_reference(args, name, mesh_kw, variant, cfg)

    # ------------------------------------------------------------------ native arm (B200)
    import torch
    import psm_b200
    if not torch.cuda.is_available():
        raise SystemExit('bench.py --impl native needs a B200 (no CPU fallback exists)')
    torch.cuda.set_device(local_rank)
    peaks, peak_src = measured_peaks()
    sampler = ClockSampler(local_rank)

    if world == 1:
        sampler.start()
        res = single_gpu_measure(torch, psm_b200, args, name, mesh_kw, variant, ncol, local_rank)
        sm, dc, n, geo = res['sm'], res['dc'], res['n'], res['geo']
        e2e_each = dc.time_host('fields', args.steps)
        rows_each = dc.time_host('rows', max(10, args.steps // 2))
        # deltaU_to_deltaP falls back to p_prev (always finite); U_to_gradP keeps NaN where the reference's grid->cell
        # interpolation is NaN (cells outside the grid hull, GRAD has no previous-gradient fallback)
        assert np.isfinite(dc.h_out.numpy()).mean() > (0.999 if variant == 'deltaU_to_deltaP' else 0.95)
        parts_f = dc.host_parts('fields')
        parts_r = dc.host_parts('rows')
        clocks = sampler.stop()
        lat = latency_run(torch, args, res) if args.latency else None
        cpu = None
        if not args.no_cpu_baseline:
            torch.set_num_threads(os.cpu_count() or 1)
            kw, what = reference_sample_kw(name, mesh_kw, 1)
            if kw == mesh_kw:
                mesh_c, F_c, o = res['mesh'], res['F'], oracle_for(variant, res['params'], res['mesh'], res['F'], res['tables'])
            else:
                mesh_c, F_c, params_c, tables_c, _ = build_case(kw, variant)
                o = oracle_for(variant, params_c, mesh_c, F_c, tables_c)
            sec = time_oracle(o, variant, F_c, args.cpu_steps, 1)
            nc = mesh_c['cells'].shape[0]
            cpu = {'value': nc / sec, 'unit': 'cells/s', 'cores': os.cpu_count(), 'kind': 'port',
                   'sample': '%d full steps of the NumPy/SciPy oracle on %s: %d cells (TF Dense stack replaced by float32 NumPy; '
                             'init excluded)' % (args.cpu_steps, what, nc), 'ms_per_step': sec * 1e3}
            del o
        roof = roofline_block(res['sb'], res['stages'], peaks, peak_src, name)
        gemm = gemm_block(res, peaks)
        e2e_s = float(np.mean(e2e_each))
        u_bytes, p_bytes = n * 24 * (2 if ncol == 7 else 1), n * 8
        line = {'metric': 'surrogate_cells_per_s', 'value': n / (res['ms_step'] * 1e-3), 'unit': 'cells/s', 'n_gpus': 1,
                'steps': args.steps, 'warmup': max(args.warmup, 3), 'ms_per_step': res['ms_step'], 'higher_is_better': True,
                'scaling': scaling, 'vs_baseline': None, 'dtype': 'f32 (f64 at the ABI and in the offset chain)',
                'data': 'synthetic', 'config': cfg, 'clocks': clocks,
                'e2e': {'value': n / e2e_s, 'unit': 'cells/s', 'h2d_bytes_per_step': int(u_bytes + p_bytes),
                        'd2h_bytes_per_step': int(n * sm.n_fields * 8), 'ms_per_step': e2e_s * 1e3,
                        'entry_point': 'psm_predict_fields: pinned U double[n][3] (OpenFOAM vector layout) + p double[n] in, p out; '
                                       'p goes up in chunks while the kernels run, the grid->cell gather and the copy of the pressures '
                                       'down follow chunk by chunk (both directions of the link busy at once)',
                        'p50_ms': float(np.percentile(e2e_each, 50) * 1e3), 'p99_ms': float(np.percentile(e2e_each, 99) * 1e3),
                        'device_events_ms': {'h2d_U': float(parts_f[0]), 'kernels_p_copy_and_chunked_d2h': float(parts_f[1] + parts_f[2])},
                        'rows5': {'entry_point': 'psm_predict: pinned double[n][%d] rows (the reference layout, FOAM/PythonComm.H:2-9)' % ncol,
                                  'value': n / float(np.mean(rows_each)), 'ms_per_step': float(np.mean(rows_each) * 1e3),
                                  'h2d_bytes_per_step': int(n * ncol * 8), 'd2h_bytes_per_step': int(n * sm.n_fields * 8),
                                  'device_events_ms': {'h2d': float(parts_r[0]), 'kernels_and_chunked_d2h': float(parts_r[1] + parts_r[2])}}},
                'gpu_launches': res['launches'], 'roofline': roof, 'gemm': gemm, 'cpu_baseline': cpu, 'stages': res['stages'],
                'stage_event_overhead_ms': res['event_overhead_ms'],
                'tables': "cells -> grid: SciPy Qhull (as the reference); grid -> cell: closed form (psm_b200.tables.regular_grid_back_tables, "
                          "back='closed_form'), handed to both the GPU path and the CPU oracle",
                'geometry': geo, 'init_tables_s': res['t_init'], 'n_cells_total': float(n)}
        if lat:
            line['latency'] = lat
        sm.close()
        del dc, res
        torch.cuda.empty_cache()
        if name == 'c2' and not args.no_c5 and not args.latency:
            # the same kernels where the working set no longer fits the 126 MB L2 (north-star: >= 70 % of HBM roofline)
            a5 = argparse.Namespace(**vars(args))
            a5.steps, a5.warmup = min(args.steps, 20), 3
            r5 = single_gpu_measure(torch, psm_b200, a5, 'c5', dict(syn.CONFIGS['c5']), variant, ncol, local_rank)
            rc5 = roofline_block(r5['sb'], r5['stages'], peaks, peak_src, 'c5')
            rc5.update({'workload': 'c5: %s' % syn.CONFIGS['c5'], 'n_cells': r5['n'], 'ms_per_step': r5['ms_step'],
                        'cells_per_s': r5['n'] / (r5['ms_step'] * 1e-3), 'stages_ms': {k: v['ms'] for k, v in r5['stages'].items()},
                        'gemm': gemm_block(r5, peaks)})
            line['roofline_c5'] = rc5
            r5['sm'].close()
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ N > 1: one domain over the ranks
    import torch.distributed as dist
    dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    sm = psm_b200.PressureSurrogate(variant, device=local_rank, input_cols=ncol, timings=False)
    mesh, F, params, sh, L, t_init = build_sharded_case(world, rank, variant, dist, mesh_kw)
    ids = [psm_b200.PressureSurrogate.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    sm.load_params(params)
    sm.comm_init(ids[0], rank, world)
    sm.init_shard(sh)
    n = sh['n_owned']
    own = sh['owned_ids']
    cells_np = np.ascontiguousarray(syn.pack_cells(mesh, F, with_delta=(ncol == 7))[own])
    geo = sm.geometry()
    F_own = {k: v[own] for k, v in F.items()}
    dc = DeviceCase(sm, torch, dist, world, n, cells_np, F_own, ncol, variant)
    dc.run_device(max(args.warmup, 3), False)
    dc.barrier()
    sampler.start()
    sm.wait_ns(reset=True)
    ms_total, _, each_dev = dc.run_device(args.steps, False)
    assert sm.synchronize() == 0, 'the last timed step was short-cut by the skip rule'
    # per-phase wait histogram of the timed steps: what CTA 0 of each consuming kernel spent waiting for its peers' pushes
    wz = sm.wait_ns(reset=True)
    mine = {'rank': rank, 'cells': int(n), 'ms_per_step': ms_total / args.steps, 'p50_ms': float(np.percentile(each_dev, 50)),
            'max_ms': float(np.max(each_dev)),
            'wait_us_per_step': {ph: (wz['ns'][i] / max(wz['count'][i], 1)) * 1e-3
                                 for i, ph in enumerate(('ghost_cells_and_maxima', 'strip_means', 'ghost_pixels'))}}
    # the same steps back to back without the L2 flush between them (what a solver loop does): the flush kernels of the ranks
    # end at slightly different times, which the timed interval above pays for as waiting
    dc.barrier()
    ms_nf, _, _ = dc.run_device(args.steps, False, flush=False)
    wz2 = sm.wait_ns(reset=True)
    mine['ms_per_step_no_flush'] = ms_nf / args.steps
    mine['wait_us_per_step_no_flush'] = {ph: (wz2['ns'][i] / max(wz2['count'][i], 1)) * 1e-3
                                         for i, ph in enumerate(('ghost_cells_and_maxima', 'strip_means', 'ghost_pixels'))}
    per_rank = [None] * world
    dist.all_gather_object(per_rank, mine)
    dc.barrier()
    launches = sm.launch_count() * args.steps
    sm.set_timings(True)
    _, stage_ms, _ = dc.run_device(args.steps, True)
    sm.set_timings(False)
    e2e_each = dc.time_host('fields', args.steps)
    assert np.isfinite(dc.h_out.numpy()).mean() > (0.999 if variant == 'deltaU_to_deltaP' else 0.95)
    dc.barrier()
    parts_f = dc.host_parts('fields')
    dc.barrier()
    clocks = sampler.stop()
    # parity of exactly what was benchmarked: one more collective step on the first time level (U, dU as in the fields)
    parity = None
    if not args.no_parity:
        if ncol == 5 and variant == 'deltaU_to_deltaP':
            prev = cells_np.copy()
            prev[:, 0] += F_own['dUx']
            prev[:, 1] += F_own['dUy']
            sm.predict(prev)
        out_own, _ = sm.predict(cells_np)
        tag = os.environ.get('MASTER_PORT', '0')
        parity = multi_parity(torch, psm_b200, dist, rank, world, variant, ncol, local_rank, mesh, F, params, sh, L, out_own, tag)
    e2e_s = float(np.sum(e2e_each))
    t = torch.tensor([ms_total, e2e_s], dtype=torch.float64, device='cuda')
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, e2e_s = t.tolist()
    nt = torch.tensor([float(n)], dtype=torch.float64, device='cuda')
    dist.all_reduce(nt)
    n_total = nt.item()
    ms_step = ms_total / args.steps
    if rank == 0:
        fused = geo['grid_w'] % 4 == 0 and not os.environ.get('PSM_NO_FUSED_EXTRACT')
        sb = stage_bytes(n, (geo['row1'] - geo['row0']) * geo['grid_w'], geo['n_local_blocks'], sm.n_fields, sm.pc_in, sm.pc_p, ncol,
                         fused_extract=fused, from_blocks=bool(geo.get('back_from_blocks', 0)))
        stage_avg = dict(zip(psm_b200._capi.TIMING_NAMES, (stage_ms / args.steps).tolist()))
        stages, ovh = stage_report(sb, stage_avg)
        roof = roofline_block(sb, stages, peaks, peak_src, 'n%d' % world, live_traffic=False)
        roof['note'] = "rank 0's share; per-stage times of the eager pass include waiting for the neighbouring ranks"
        line = {'metric': 'surrogate_cells_per_s', 'value': n_total / (ms_step * 1e-3), 'unit': 'cells/s', 'n_gpus': world,
                'steps': args.steps, 'warmup': max(args.warmup, 3), 'ms_per_step': ms_step, 'higher_is_better': True,
                'scaling': scaling, 'vs_baseline': None, 'dtype': 'f32 (f64 at the ABI and in the offset chain)',
                'data': 'synthetic', 'config': cfg, 'clocks': clocks,
                'transport': 'exchanges pushed over NVLink peer memory (cudaIpc)' if geo.get('peer_memory_exchange') else 'NCCL send/recv + all-reduce',
                'e2e': {'value': n_total / (e2e_s / args.steps), 'unit': 'cells/s', 'h2d_bytes_per_step': int(n * 32 + (n * 24 if ncol == 7 else 0)),
                        'd2h_bytes_per_step': int(n * sm.n_fields * 8), 'ms_per_step': e2e_s / args.steps * 1e3,
                        'entry_point': 'psm_predict_fields on every rank (its own cells), pinned U double[n][3] + p double[n]; bytes are per rank',
                        'device_events_ms': {'h2d_U': float(parts_f[0]), 'kernels_p_copy_and_chunked_d2h': float(parts_f[1] + parts_f[2])}},
                'gpu_launches': launches, 'roofline': roof, 'cpu_baseline': None, 'stages': stages, 'stage_event_overhead_ms': ovh,
                'parity': parity, 'parity_rel_l2': None if not parity else parity.get('vs_oracle', parity.get('vs_single_gpu')),
                'rank_alignment': 'a one-element NCCL all-reduce in stream order closes each L2 flush, so the ranks open their timed '
                                  'interval together; per_rank[].ms_per_step_no_flush = the same steps back to back, no flush, no alignment',
                'per_rank': per_rank, 'limiting_phase': max(per_rank[0]['wait_us_per_step'],
                                                            key=lambda ph: max(r['wait_us_per_step'][ph] for r in per_rank)),
                'geometry': geo, 'init_tables_s': t_init, 'n_cells_total': n_total}
        print(json.dumps(line))
    sm.close()
    dist.destroy_process_group()
    return 0


if __name__ == '__main__':
    sys.exit(main())
