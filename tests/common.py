"""Shared helpers for the test-suite: golden access, case replay, tolerances."""
import os

import numpy as np

from oracle import cases as C
from oracle import nmf_oracle as O

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')

# north_star tolerances: relative Frobenius difference of W, H and relative difference of recon_err
TOL_FACTOR = {'float32': 1e-4, 'float64': 1e-10}
TOL_ERR = {'float32': 1e-5, 'float64': 1e-10}

_cache = {}


def golden(fname='nmf_cases.npz'):
    if fname not in _cache:
        with np.load(os.path.join(GOLDEN_DIR, fname)) as z:
            _cache[fname] = {k: z[k] for k in z.files}
    return _cache[fname]


def golden_case(name):
    g = golden()
    case = C.CASES_BY_NAME[name]
    P = case['grid'][0] * case['grid'][1]
    out = []
    for r in range(P):
        d = {k.split('/', 2)[2]: v for k, v in g.items() if k.startswith('%s/%d/' % (name, r))}
        out.append(d)
    return out


def oracle_inputs(case):
    """Replay the RNG convention of oracle/cases.py for every virtual rank."""
    p_r, p_c = case['grid']
    P = p_r * p_c
    rngs = [np.random.RandomState(case['seed']) for _ in range(P)]
    A = None
    for r in range(P):
        A = C.draw_global(case, rngs[r])
    blocks = O.split_matrix(A, p_r, p_c)
    factors = None
    if case['given_factors']:
        g = O.VGrid(p_r, p_c)
        sh = O.compute_dims(blocks, g, case['k'])
        factors = []
        for r in range(P):
            ml, nl = (sh.m_loc[r], sh.n_loc[r]) if sh.topo == '2d' else blocks[r].shape
            factors.append(C.draw_given_factors(case, rngs[r], ml, nl))
    return A, blocks, rngs, factors


def run_oracle(case):
    A, blocks, rngs, factors = oracle_inputs(case)
    p_r, p_c = case['grid']
    return O.fit(blocks, p_r, p_c, case['k'], case['norm'], case['method'], case['itr'], rngs=rngs,
                 factors=factors, prune=case['prune'], W_update=case['W_update'])


def rel_fro(X, Y):
    X = np.asarray(X, dtype=np.float64)
    Y = np.asarray(Y, dtype=np.float64)
    d = np.linalg.norm(Y)
    return np.linalg.norm(X - Y) / (d if d > 0 else 1.0)
