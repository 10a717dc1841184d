"""Parameter files: one ``.npz`` holding what the reference spreads over ``maxs``, ``maxs_PCA``,
``mean_std.npz``, ``ipca_*.pkl`` and the Keras ``.h5`` (SMC:70-87,505-511; PMP:103-118)."""
import numpy as np


def save_npz(path, p):
    d = {k: v for k, v in p.items() if k not in ('mlp_weights', 'mlp_biases')}
    for i, (w, b) in enumerate(zip(p['mlp_weights'], p['mlp_biases'])):
        d['dense_%d_kernel' % i] = w
        d['dense_%d_bias' % i] = b
    np.savez(path, **d)


def load_npz(path):
    z = np.load(path, allow_pickle=False)
    p = {}
    for k in z.files:
        if not k.startswith('dense_'):
            v = z[k]
            p[k] = v.item() if v.shape == () else v
    n = len([k for k in z.files if k.endswith('_kernel')])
    p['mlp_weights'] = [z['dense_%d_kernel' % i] for i in range(n)]
    p['mlp_biases'] = [z['dense_%d_bias' % i] for i in range(n)]
    if 'standardization' in p:
        p['standardization'] = str(p['standardization'])
    return p


def select_num_pc(explained_variance_ratio, var, max_num_PC):
    """SMC:86-87: argmax(cumsum > var) when 1 < argmax <= max_num_PC, else max_num_PC."""
    a = int(np.argmax(np.asarray(explained_variance_ratio).cumsum() > var))
    return a if 1 < a <= max_num_PC else max_num_PC


def select_num_pc_thesis(explained_variance_ratio, var):
    """PMP:112-113: ``int(argmax(cumsum > var))`` -- no lower bound, no cap (0.95 for p, 0.995 for the input)."""
    return int(np.argmax(np.asarray(explained_variance_ratio).cumsum() > var))


def from_reference_objects(maxs, pca_in, pca_p, dense_kernels, dense_biases, var_in=0.95, var_p=0.95,
                           max_num_PC=128, scaler=None, maxs_PCA=None, n_out_channels=1, thesis=False):
    """Assemble the dict from the reference's own artefact objects (unpickled PCA objects with
    ``components_``/``mean_``/``explained_variance_ratio_``, Keras weight lists).  ``thesis=True`` applies the
    component cuts of the solver module (PMP:112-113: 0.995 for the input, 0.95 for p, uncapped) instead of SMC:86-87."""
    if thesis:
        pc_in = select_num_pc_thesis(pca_in.explained_variance_ratio_, 0.995)
        pc_p = select_num_pc_thesis(pca_p.explained_variance_ratio_, 0.95)
    else:
        pc_in = select_num_pc(pca_in.explained_variance_ratio_, var_in, max_num_PC)
        pc_p = select_num_pc(pca_p.explained_variance_ratio_, var_p, max_num_PC)
    p = dict(maxs=np.asarray(maxs, dtype=np.float64), n_out_channels=n_out_channels,
             pca_in_components=np.asarray(pca_in.components_)[:pc_in], pca_in_mean=np.asarray(pca_in.mean_),
             pca_out_components=np.asarray(pca_p.components_)[:pc_p], pca_out_mean=np.asarray(pca_p.mean_),
             mlp_weights=[np.asarray(w, dtype=np.float32) for w in dense_kernels],
             mlp_biases=[np.asarray(b, dtype=np.float32) for b in dense_biases])
    if scaler is not None and 'min_in' in scaler:       # min_max_values.npz (SMC:513-520)
        p.update(standardization='min_max', min_in=scaler['min_in'], max_in=scaler['max_in'],
                 min_out=scaler['min_out'], max_out=scaler['max_out'])
    elif scaler is not None:
        p.update(standardization='std', mean_in=scaler['mean_in'], std_in=scaler['std_in'],
                 mean_out=scaler['mean_out'], std_out=scaler['std_out'])
    else:
        p.update(standardization='max_abs', max_abs_input_PCA=float(maxs_PCA[0]), max_abs_output_PCA=float(maxs_PCA[1]))
    return p
