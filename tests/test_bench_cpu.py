"""CPU tests of bench.py's bookkeeping (no GPU): the workload the contract names per GPU count, one `config` object for both arms,
the algorithmic bytes behind the roofline figures, the bounded reference sample, and the roofline block itself."""
import argparse
import importlib.util
import os

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location('bench', os.path.join(REPO, 'bench.py'))
bench = importlib.util.module_from_spec(spec)
spec.loader.exec_module(bench)


def _args(**kw):
    d = dict(workload=None, variant='deltaU_to_deltaP')
    d.update(kw)
    return argparse.Namespace(**d)


def test_default_workloads_follow_baseline_configs():
    name, kw, scaling = bench.resolve_workload(_args(), 1)
    assert name == 'c2' and kw['H'] == kw['W'] == 1000 and scaling == 'weak'            # configs[1]
    for n in (2, 4, 8):
        name, kw, scaling = bench.resolve_workload(_args(), n)
        assert name == 'c4' and kw['H'] == kw['W'] == 4000 and scaling == 'strong'      # configs[3]: one 16 M-cell domain
    name, kw, scaling = bench.resolve_workload(_args(workload='c2xN'), 4)
    assert name == 'c2x4' and scaling == 'weak' and kw['W'] == 1000 and kw['H'] >= 4000
    with pytest.raises(SystemExit):
        bench.resolve_workload(_args(workload='nope'), 1)


def test_config_is_one_object_for_both_arms():
    name, kw, _ = bench.resolve_workload(_args(), 8)
    a = bench.config_dict(name, kw, 'deltaU_to_deltaP', 8, 5)
    b = bench.config_dict(name, kw, 'deltaU_to_deltaP', 8, 5)
    assert a == b and 'workload' in a and 'l2' in a and 'closed form' in a['tables'] and 'model' not in a


def test_reference_sample_is_bounded_and_on_the_same_domain():
    kw, what = bench.reference_sample_kw('c2', dict(bench.syn.CONFIGS['c2']), 1)
    assert kw == bench.syn.CONFIGS['c2'] and what == 'the whole mesh'
    kw, what = bench.reference_sample_kw('c4', dict(bench.syn.CONFIGS['c4']), 8)
    assert kw['W'] == 4000 and kw['nx'] == 4000 and kw['H'] == 512 and kw['nx'] * kw['ny'] <= 2_200_000      # full width, a band of rows
    kw, _ = bench.reference_sample_kw('c2x8', bench.weak_mesh_kw(8, 'deltaU_to_deltaP'), 8)
    assert kw == bench.syn.CONFIGS['c2']


def test_algorithmic_bytes_per_stage():
    n, G, B = 979_929, 1_000_000, 121
    xu = 4 * B * 2 * 128 * 128
    fused = bench.stage_bytes(n, G, B, 1, 128, 128, 5, fused_extract=True)
    grid = bench.stage_bytes(n, G, B, 1, 128, 128, 5, fused_extract=True, grid_a=True)
    plain = bench.stage_bytes(n, G, B, 1, 128, 128, 5, fused_extract=False)
    assert fused['gather'] == 24 * G + 8 * n + xu and fused['extract'] == 0          # SURVEY 8(d) K1 + the operand write
    assert grid['gather'] == 24 * G + 8 * n + 8 * G and grid['extract'] == 0          # planes written, no operand
    assert plain['gather'] == grid['gather'] and plain['extract'] == 8 * G + xu
    assert grid['pca_project'] < fused['pca_project']                                 # unique bytes: the planes, not B copies
    assert fused['prep'] == n * (40 + 16 + 32) and fused['back_gather'] == n * (24 + 6 + 12 + 8 + 8)
    fl = bench.stage_flops(B, 1, 128, 128)
    assert fl['pca_project'] == 2.0 * B * 2 * 128 * 128 * 128


def test_roofline_block_names_the_longest_launched_hbm_stage():
    sb = bench.stage_bytes(1_000_000, 1_000_000, 121, 1, 128, 128, 5, grid_a=True)
    avg = {'prep': 0.0173, 'gather': 0.0143, 'extract': 0.0026, 'pca_project': 0.025, 'mlp': 0.03, 'pca_inverse': 0.0143,
           'strip_means': 0.02, 'offsets': 0.0025, 'place': 0.0026, 'back_gather': 0.0162}
    stages, ovh = bench.stage_report(sb, avg)
    assert abs(ovh - 0.0025) < 1e-9 and stages['prep']['ms_net'] == pytest.approx(0.0148)      # event overhead of the empty stages
    r = bench.roofline_block(sb, stages, {'hbm_gbs': 6456.5}, 'measured', 'c2', live_traffic=False)
    assert r['kernel'] == 'prep' and r['bound'] == 'hbm' and r['unit'] == 'GB/s' and r['traffic'] is None
    assert r['frac'] == pytest.approx(sb['prep'] / 0.0148e-3 / 1e9 / 6456.5)
    assert set(r['hbm_kernels']) == {'prep', 'gather', 'back_gather'}                            # place / extract launch nothing
