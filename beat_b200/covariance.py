"""
Operand producers for the misfit: the weights ``U = chol(C^-1)^T`` and ``log|C|`` per dataset.

Host-side mirror of ``beat.heart.Covariance`` (reference: beat/heart.py:104-263) and the noise structures of
``beat/covariance.py:24-91``.  These run at set-up and between SMC stages (``update_weights``,
beat/models/seismic.py:1509-1534), never per evaluation, so they are plain numpy/scipy here; their outputs are
uploaded with ``Context.update_weights`` and stay resident in HBM.
"""
from __future__ import annotations

import numpy as np
from scipy import linalg


def exponential_data_covariance(n, dt, tzero):
    """Toeplitz ``exp(-|ti - tj| / T0)`` sub-covariance without variance (beat/covariance.py:24-51)."""
    idx = np.arange(n)
    return np.exp(-np.abs(idx[:, None] - idx[None, :]) * (dt / tzero))


def identity_data_covariance(n, dt=None, tzero=None):
    """beat/covariance.py:54-66."""
    return np.eye(n)


NoiseStructureCatalog = {"variance": identity_data_covariance, "exponential": exponential_data_covariance}


class Covariance:
    """Data + prediction covariances of one dataset; same public surface as the reference class."""

    def __init__(self, data=None, pred_g=None, pred_v=None):
        self.data, self.pred_g, self.pred_v = data, pred_g, pred_v

    @property
    def c_total(self):                                             # heart.py:158-164
        tot = np.array(self.data, dtype=np.float64, copy=True)
        for extra in (self.pred_g, self.pred_v):
            if extra is not None:
                if extra.size != tot.size:
                    if extra.sum() == 0.0:
                        continue
                    raise ValueError("covariances defined but size inconsistent!")
                tot = tot + extra
        return tot

    def inverse(self, factor=1.0):                                 # heart.py:173-181
        Cx = self.c_total * factor
        if Cx.sum() == 0:
            raise ValueError("No covariances given!")
        return np.linalg.inv(Cx)

    def chol(self, factor=1.0):                                    # heart.py:201-209
        Cx = self.c_total * factor
        if Cx.sum() == 0:
            raise ValueError("No covariances given!")
        return linalg.cholesky(Cx, lower=True)

    @property
    def chol_inverse(self):                                        # heart.py:211-237
        try:
            return np.linalg.cholesky(self.inverse()).T
        except np.linalg.LinAlgError:
            inverse_chol = np.linalg.inv(self.chol().T)
            _, chol_ur = np.linalg.qr(inverse_chol.T)
            return chol_ur

    @property
    def log_pdet(self):                                            # heart.py:239-245
        return float(np.log(np.diag(self.chol())).sum() * 2.0)


def log_determinant(A, inverse=False):
    """beat/heart.py:65-89."""
    chol = linalg.cholesky(A, lower=True)
    if inverse:
        chol = np.linalg.inv(chol)
    return float(np.log(np.diag(chol)).sum() * 2.0)


def smoothing_operator_nearest_neighbor(n_patch_strike, n_patch_dip, patch_size_strike, patch_size_dip):
    """Second-order nearest-neighbour Laplacian on the patch grid (beat/models/laplacian.py:172-258), vectorised."""
    n = n_patch_dip * n_patch_strike
    r, c = np.divmod(np.arange(n), n_patch_strike)
    ddip, dstr = 1.0 / patch_size_dip ** 2, 1.0 / patch_size_strike ** 2
    L = np.zeros((n, n))
    up, down, left, right = r > 0, r < n_patch_dip - 1, c > 0, c < n_patch_strike - 1
    i = np.arange(n)
    L[i, i] = -(ddip * (up.astype(float) + down) + dstr * (left.astype(float) + right))
    L[i[up], i[up] - n_patch_strike] = ddip
    L[i[down], i[down] + n_patch_strike] = ddip
    L[i[left], i[left] - 1] = dstr
    L[i[right], i[right] + 1] = dstr
    return L
