"""Repeat the same c2 step and report which stage (if any) differs between calls (graph replay, reused buffers)."""
import hashlib, sys, os
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, 'solving-poisson-s-equation-through-dl-for-cfd-apllications_b200'))
import psm_b200
from psm_b200 import synthetic as syn, tables as ptables
wl = sys.argv[1] if len(sys.argv) > 1 else 'c2'
mesh = syn.make_mesh(seed=0, **syn.CONFIGS[wl]); F = syn.make_fields(mesh, seed=0)
params = syn.make_params(seed=0, pc_in=128, pc_p=128)
t = ptables.build_tables(mesh['cells'], mesh['top'], mesh['obst'], F['p_prev'], back='closed_form')
cells = syn.pack_cells(mesh, F, with_delta=True)
h = lambda a: hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()[:10]
with psm_b200.PressureSurrogate('deltaU_to_deltaP', input_cols=7) as sm:
    sm.load_params(params); sm.init_tables(t)
    ref = None
    for k in range(int(os.environ.get('PROBE_CALLS', '6'))):
        out, rc = sm.predict(cells)
        st = {n: sm.stage(n) for n in ('xu', 'x_input', 'mlp_out', 'blocks', 'means', 'offsets', 'scalars')}
        sig = {n: h(v) for n, v in st.items()}; sig['out'] = h(out)
        if ref is None:
            ref, ref_st, ref_out = sig, st, out
        diff = [n for n in sig if sig[n] != ref[n]]
        msg = ''
        for n in diff:
            a, b = (out, ref_out) if n == 'out' else (st[n], ref_st[n])
            d = np.abs(np.nan_to_num(a.astype(np.float64)) - np.nan_to_num(b.astype(np.float64)))
            msg += ' %s: %d elems, max %.3e;' % (n, int((d > 0).sum()), d.max())
            if n in ('xu', 'x_input'):
                rows = np.unique(np.nonzero(d.reshape(d.shape[0], -1) > 0)[0])
                msg += ' [blocks %s]' % rows[:12].tolist()
        print('call %d rc %d differs from call 0 in: %s' % (k, rc, msg or 'nothing'))
