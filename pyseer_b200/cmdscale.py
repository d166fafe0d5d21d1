"""Classical multidimensional scaling (what pyseer/cmdscale.py:15-54 computes): double-centre
the squared distances, eigendecompose, keep the positive-eigenvalue components."""
import numpy as np


def cmdscale(D):
    D = np.asarray(D, dtype=float)
    n = D.shape[0]
    J = np.eye(n) - np.full((n, n), 1.0 / n)
    B = -0.5 * J.dot(D ** 2).dot(J)
    w, V = np.linalg.eigh(B)
    order = np.argsort(w)[::-1]
    w, V = w[order], V[:, order]
    pos = w > 0
    return V[:, pos] * np.sqrt(w[pos]), w[pos]
