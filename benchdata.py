"""Synthetic problem definitions of the benchmark and of the parity tests at the BASELINE sizes
(SURVEY 8d): plain NumPy, shared by both arms of bench.py and by tests/ -- it imports neither the
product (pyseer_b200) nor the oracle."""
import math

import numpy as np

SEED = 20261017


# ----------------------------------------------------------------------------------------
def lmm_problem(n, seed=SEED, clonal=0, n_cov=0):
    """X (N x D, last column ones), y (continuous, heritable), normalised kinship K.
    clonal > 0: `clonal` founder genotypes, every sample a founder plus 0.5 % private mutations --
    a block-structured kinship of numerical rank ~ clonal << N."""
    rng = np.random.RandomState(seed % (2 ** 31))
    m = 2 * n
    af = rng.uniform(0.05, 0.95, m)
    if clonal:
        founders = (rng.uniform(size=(clonal, m)) < af).astype(np.float32)
        G = founders[rng.randint(0, clonal, size=n)]
        flip = rng.uniform(size=(n, m)) < 0.005
        G = np.where(flip, 1.0 - G, G).astype(np.float32)
    else:
        G = (rng.uniform(size=(n, m)) < af).astype(np.float32)
    K = (G @ G.T).astype(np.float64)
    g = G.astype(np.float64) @ rng.normal(size=m)
    g = (g - g.mean()) / g.std()
    y = math.sqrt(0.5) * g + math.sqrt(0.5) * rng.normal(size=n)
    K *= float(n) / np.diag(K).sum()                    # lmm.py:107-112
    X = np.ones((n, 1))
    if n_cov:
        X = np.c_[rng.normal(size=(n, n_cov)), X]        # lmm.py:95-99: covariates, then ones
    return X, y, K


def fixed_problem(n, dims=10, seed=SEED):
    """configs[2]: binary phenotype, population structure carried by `dims` MDS components
    scaled as input.py:135-136."""
    rng = np.random.RandomState(seed % (2 ** 31) + 2)
    m = rng.uniform(-1, 1, size=(n, dims))
    m = m / np.abs(m).max(0)
    lin = m[:, :3].sum(1) + rng.normal(size=n)
    y = (lin > np.median(lin)).astype(float)
    return m, y


def burden_regions(n_regions, seed_offset=0):
    """Member lists of bench.py's burden workload: region r is the union of 1-20 consecutive
    record rows.  Returns (offsets[n_regions + 1], members)."""
    rng = np.random.RandomState(SEED % (2 ** 31) + 17 + seed_offset)
    sizes = rng.randint(1, 21, size=n_regions)
    offs = np.zeros(n_regions + 1, dtype=np.int64)
    offs[1:] = np.cumsum(sizes)
    return offs, np.arange(int(offs[-1]), dtype=np.int32)



def fixed_cont_problem(n, dims=10):
    """Fixed effects with a continuous phenotype (OLS)."""
    rng = np.random.RandomState(SEED % (2 ** 31) + 3)
    m, _ = fixed_problem(n, dims)
    return m, m[:, :3].sum(1) + rng.normal(size=n)
