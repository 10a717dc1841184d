"""Artefact handling on the host (psm_b200/params.py): the reference's component-count rules and the one-file container."""
import types

import numpy as np

from psm_b200 import params as pp, synthetic as syn


def _evr(n_total, n_keep, var):
    evr = np.full(n_total, (1.0 - var) / (2.0 * n_total))
    evr[:n_keep] = (var - 1e-6) / n_keep
    evr[n_keep] += 2e-6
    return evr


def test_component_count_rules():
    # SMC:86-87: argmax(cumsum > var), used only if 1 < argmax <= max_num_PC, else max_num_PC
    assert pp.select_num_pc(_evr(40, 24, 0.95), 0.95, 128) == 24
    assert pp.select_num_pc(_evr(200, 150, 0.95), 0.95, 128) == 128          # above the cap
    assert pp.select_num_pc(np.array([0.97, 0.02, 0.01]), 0.95, 128) == 128  # argmax == 0 is not > 1 -> cap (reference quirk)
    assert pp.select_num_pc(np.array([0.5, 0.46, 0.04]), 0.95, 128) == 128   # argmax == 1 likewise
    assert pp.select_num_pc(np.full(10, 0.05), 0.95, 128) == 128             # never exceeded: argmax of all-False is 0
    # PMP:112-113: plain argmax, no cap
    assert pp.select_num_pc_thesis(_evr(40, 24, 0.995), 0.995) == 24
    assert pp.select_num_pc_thesis(np.array([0.97, 0.02, 0.01]), 0.95) == 0


def test_from_reference_objects_and_npz_round_trip(tmp_path):
    P = syn.make_params(seed=1, pc_in=12, pc_p=9, standardization='max_abs')
    rng = np.random.default_rng(0)
    K_in, K_out = P['pca_in_components'].shape[1], P['pca_out_components'].shape[1]
    pca_in = types.SimpleNamespace(components_=np.concatenate([P['pca_in_components'], rng.standard_normal((5, K_in))]),
                                   mean_=P['pca_in_mean'], explained_variance_ratio_=_evr(17, 12, 0.995))
    pca_p = types.SimpleNamespace(components_=np.concatenate([P['pca_out_components'], rng.standard_normal((5, K_out))]),
                                  mean_=P['pca_out_mean'], explained_variance_ratio_=_evr(14, 9, 0.95))
    q = pp.from_reference_objects(P['maxs'], pca_in, pca_p, P['mlp_weights'], P['mlp_biases'],
                                  maxs_PCA=(P['max_abs_input_PCA'], P['max_abs_output_PCA']), thesis=True)
    assert q['pca_in_components'].shape[0] == 12 and q['pca_out_components'].shape[0] == 9
    np.testing.assert_array_equal(q['pca_in_components'], P['pca_in_components'])
    assert q['standardization'] == 'max_abs'
    path = tmp_path / 'psm_params.npz'
    pp.save_npz(path, q)
    r = pp.load_npz(path)
    assert r['standardization'] == 'max_abs' and float(r['max_abs_input_PCA']) == float(P['max_abs_input_PCA'])
    for a, b in zip(r['mlp_weights'], P['mlp_weights']):
        np.testing.assert_array_equal(a, b)
    np.testing.assert_array_equal(r['pca_out_components'], P['pca_out_components'])
    # SMC rule with the 'std' scaler
    Ps = syn.make_params(seed=2, pc_in=12, pc_p=9, standardization='std')
    pca_in.explained_variance_ratio_ = _evr(17, 12, 0.95)
    q2 = pp.from_reference_objects(Ps['maxs'], pca_in, pca_p, Ps['mlp_weights'], Ps['mlp_biases'], var_in=0.95, var_p=0.95,
                                   max_num_PC=128, scaler={k: Ps[k] for k in ('mean_in', 'std_in', 'mean_out', 'std_out')})
    assert q2['standardization'] == 'std' and q2['pca_in_components'].shape[0] == 12
