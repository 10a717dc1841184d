"""Shared test helpers: build oracle objects from the plain parameter dicts."""
import ast
import os

import numpy as np

from oracle.pipeline import SurrogateParams

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def oracle_params(p):
    return SurrogateParams(**{k: v for k, v in p.items() if k != 'shape'})


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + '.npz'))
    mesh_kw = ast.literal_eval(str(z['mesh_kw']))
    return z, mesh_kw, int(z['seed'])


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))
