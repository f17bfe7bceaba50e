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


# ------------------------------------------------------------------------------------------------------------
# Per-stage covariance update: residuals at the MAP point -> non-Toeplitz data
# covariance -> weights.  Reference: SeismicNoiseAnalyser.do_non_toeplitz (beat/covariance.py:307-325),
# estimators beat/covariance.py:716-771, running rms beat/utility.py:1141-1161, repair beat/utility.py:1034-1056,
# 1111-1138, then Covariance.chol_inverse / log_pdet and `wmap.weights[i].set_value` (beat/models/seismic.py:1509-1534).
# ------------------------------------------------------------------------------------------------------------
def running_window_rms(data, window_size, mode="valid"):
    """beat/utility.py:1141-1161."""
    return np.sqrt(np.convolve(np.power(data, 2), np.ones(window_size) / float(window_size), mode))


def autocovariance(data):
    """beat/covariance.py:716-736: autocov[j] = 1/n sum_k (d[j+k]-mean)(d[k]-mean), as one correlation."""
    x = np.asarray(data, dtype=np.float64) - np.mean(data)
    n = x.size
    return np.correlate(x, x, mode="full")[n - 1:] / n


def toeplitz_covariance(data, window_size):
    """beat/covariance.py:739-751."""
    from scipy.linalg import toeplitz
    stds = running_window_rms(data, window_size=window_size, mode="same")
    return toeplitz(autocovariance(data / stds)), stds


def non_toeplitz_covariance(data, window_size):
    """beat/covariance.py:754-771."""
    toe, stds = toeplitz_covariance(data, window_size)
    return toe * stds[:, None] * stds[None, :]


def repair_covariance(x, epsilon=np.finfo(np.float64).eps):
    """beat/utility.py:1111-1138."""
    eigval, eigvec = np.linalg.eigh(x)
    return eigvec.dot(np.diag(np.maximum(eigval, epsilon))).dot(eigvec.T)


def ensure_cov_psd(cov):
    """beat/utility.py:1034-1056."""
    try:
        np.linalg.cholesky(cov)
    except np.linalg.LinAlgError:
        cov = repair_covariance(cov)
    return cov


def weights_from_residuals_host(residuals, structure=None):
    """numpy path: residuals [nt, ns] -> (U [nt, ns, ns], log_pdet [nt]) for the `non-toeplitz` noise structure
    (get_data_covariances, beat/covariance.py:397-427: cov_d = ensure_cov_psd(scaling * covariance_structure))."""
    residuals = np.asarray(residuals, dtype=np.float64)
    nt, ns = residuals.shape
    U, lp = np.empty((nt, ns, ns)), np.empty(nt)
    ws = ns // 5                                                        # covariance.py:316
    if ws == 0:
        raise ValueError("Length of trace too short! Please widen taper in time domain or frequency bands in spectral domain.")
    for t in range(nt):
        C = non_toeplitz_covariance(residuals[t], ws)
        if structure is not None:
            C = C * structure
        cov = Covariance(data=ensure_cov_psd(C))
        U[t], lp[t] = cov.chol_inverse, cov.log_pdet
    return U, lp


def weights_from_residuals_device(residuals, device=None):
    """Batched GPU version of :func:`weights_from_residuals_host` (torch / cuSOLVER, float64): all datasets of a
    wavemap at once.  residuals: [nt, ns] numpy or torch.  Returns (U, log_pdet, C) as torch tensors on `device`.

    Stage-boundary work only (once per SMC stage), hence library calls: conv1d for the running rms, FFT for the
    autocovariance, batched Cholesky / triangular solves for U = chol(C^-1)^T and log|C|.  Matrices that are not
    positive definite are repaired by eigenvalue clipping exactly like ``ensure_cov_psd``."""
    import torch
    if device is None:
        device = torch.device("cuda", 0)
    r = torch.as_tensor(np.asarray(residuals) if not hasattr(residuals, "device") else residuals, dtype=torch.float64, device=device)
    nt, ns = r.shape
    ws = ns // 5
    if ws == 0:
        raise ValueError("Length of trace too short! Please widen taper in time domain or frequency bands in spectral domain.")
    # running rms, numpy.convolve(..., mode="same") semantics
    kern = torch.full((1, 1, ws), 1.0 / ws, dtype=torch.float64, device=device)
    pad_l = ws // 2 if ws % 2 == 0 else (ws - 1) // 2      # numpy 'same' alignment of the window
    pad_r = ws - 1 - pad_l
    sq = torch.nn.functional.pad((r * r)[:, None, :], (pad_l, pad_r))
    stds = torch.sqrt(torch.nn.functional.conv1d(sq, kern)[:, 0, :])
    x = r / stds
    x = x - x.mean(dim=1, keepdim=True)
    nfft = 2 * ns
    f = torch.fft.rfft(x, n=nfft, dim=1)
    ac = torch.fft.irfft(f * torch.conj(f), n=nfft, dim=1)[:, :ns] / ns                 # autocovariance, lags 0..ns-1
    idx = (torch.arange(ns, device=device)[:, None] - torch.arange(ns, device=device)[None, :]).abs()
    C = ac[:, idx] * stds[:, :, None] * stds[:, None, :]                                 # toeplitz * stds stds^T
    L, info = torch.linalg.cholesky_ex(C)
    bad = torch.nonzero(info).flatten().tolist()
    for t in bad:                                                                         # ensure_cov_psd / repair_covariance
        w, v = torch.linalg.eigh(C[t])
        C[t] = (v * torch.clamp(w, min=float(np.finfo(np.float64).eps))) @ v.T
    if bad:
        L = torch.linalg.cholesky(C)
    log_pdet = 2.0 * torch.log(torch.diagonal(L, dim1=1, dim2=2)).sum(dim=1)
    eye = torch.eye(ns, dtype=torch.float64, device=device).expand(nt, ns, ns)
    Cinv = torch.cholesky_solve(eye, L)
    Cinv = 0.5 * (Cinv + Cinv.transpose(1, 2))
    Lc, info2 = torch.linalg.cholesky_ex(Cinv)
    U = Lc.transpose(1, 2).contiguous()
    for t in torch.nonzero(info2).flatten().tolist():                                    # QR fallback, heart.py:234-237
        inv_chol = torch.linalg.inv(L[t].T)
        _, R = torch.linalg.qr(inv_chol.T)
        U[t] = R
    return U, log_pdet, C
