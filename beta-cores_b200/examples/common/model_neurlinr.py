"""Neural-linear regression (Bayesian linear head on fixed features): device potentials + host
conjugate posterior.  Drop-in for the hot-path functions of examples/common/model_neurlinr.py.

Data rows are z_n = [phi(x_n), y_n]; with u = phi.theta and r2 = y^2 - 2 u y + u^2:
    neurlinr_loglikelihood(z, th, sigsq)          = -1/2 log(2 pi s2) - r2/(2 s2)                      (model_neurlinr.py:90-97)
    neurlinr_beta_likelihood(z, th, beta, sigsq)  = (2 pi s2)^(-beta/2) (-(b+1)/b e^{-b r2/(2 s2)} + 1/sqrt(1+b))   (:102-110)
"""
import numpy as np
import scipy.linalg as sl

from bayesiancoresets.potentials import DevicePotential

neurlinr_loglikelihood = DevicePotential('neurlin', 'loglik', name='neurlinr_loglikelihood')
neurlinr_beta_likelihood = DevicePotential('neurlin', 'betalik', name='neurlinr_beta_likelihood')


def weighted_post(th0, Sig0inv, sigsq, z, w):
    """conjugate weighted posterior of the linear head (model_neurlinr.py:115-122); host, D x D."""
    z = np.atleast_2d(z)
    X, Y = z[:, :-1], z[:, -1]
    LSigpInv = np.linalg.cholesky(Sig0inv + (w[:, np.newaxis]*X).T.dot(X)/sigsq)
    LSigp = sl.solve_triangular(LSigpInv, np.eye(LSigpInv.shape[0]), lower=True, overwrite_b=True, check_finite=False)
    mup = np.dot(LSigp.dot(LSigp.T), np.dot(Sig0inv, th0) + (w[:, np.newaxis]*Y[:, np.newaxis]*X).sum(axis=0)/sigsq)
    return mup, LSigp, LSigpInv


def make_conjugate_sampler(mu0, Sig0inv, sigsq, device=False, prefetch=False):
    """sampler(S, wts, pts) with the conjugate posterior of the linear head (model_neurlinr.py:115-122).  device=True: the
    host factors the D x D posterior precision, the samples are formed on the GPU by a triangular solve against the normals
    and stay there (bayesiancoresets/util/samplers.py); prefetch=True: normals drawn one call ahead (util/rng.py)."""
    D = mu0.shape[0]
    if device and D <= 160:
        from bayesiancoresets.util import rng
        from bayesiancoresets.util.samplers import ConjugateDeviceSampler
        Sig0inv_mu0 = np.dot(Sig0inv, mu0)

        def precision(wts, pts):
            if pts.shape[0] == 0:
                return Sig0inv + 0., Sig0inv_mu0 + 0.
            z = np.atleast_2d(pts)
            X, Y = z[:, :-1], z[:, -1]
            return Sig0inv + (wts[:, np.newaxis]*X).T.dot(X)/sigsq, Sig0inv_mu0 + (wts[:, np.newaxis]*Y[:, np.newaxis]*X).sum(axis=0)/sigsq
        return ConjugateDeviceSampler(D, precision, rng.activate() if prefetch else None,
                                      device_model=(2, Sig0inv, None, Sig0inv_mu0, sigsq))

    def sampler(S, wts, pts):
        if pts.shape[0] == 0:
            wts = np.zeros(1)
            pts = np.zeros((1, D+1))
        mu, L, _ = weighted_post(mu0, Sig0inv, sigsq, pts, wts)
        if device:
            from bayesiancoresets.util.samplers import affine_samples
            return affine_samples(mu, L, np.random.randn(S, D))
        return mu + np.random.randn(S, D).dot(L.T)
    return sampler


def encode_dataset(nl, Z, batch=1 << 20):
    """rows z_n = [phi(x_n), y_n] (float64) for the whole dataset, computed ONCE with a fixed encoder.

    The reference re-encodes the rows inside every likelihood call (examples/zellner_neural_linear/main.py:110-114:
    `deep_encoder(nl, pts)` is part of the lambdas), which keeps the potentials opaque to the projector and, with the
    BatchNorm layers of examples/common/neural.py:126-133 in training mode, makes phi depend on which rows share a call.
    With the encoder frozen (`nl.feature_extractor.eval()` for the torch module, or any object with `encode`), encoding
    the data up front gives the same features for every call, and the coreset classes can take the result as `data` with
    the bound device potentials above -- the fused tensor-core path -- instead of the black-box route.  `nl.encode`
    may take / return numpy arrays or torch tensors; features come back in the encoder's precision (fp32 in the
    reference) and are widened to float64 exactly like `np.hstack` does in the reference's lambda."""
    Z = np.atleast_2d(np.asarray(Z))
    out = []
    for s in range(0, Z.shape[0], batch):
        x = Z[s:s+batch, :-1].astype(np.float32)
        try:
            import torch
            phi = nl.encode(torch.from_numpy(x)) if isinstance(nl, torch.nn.Module) else nl.encode(x)
            if isinstance(phi, torch.Tensor):
                phi = phi.detach().cpu().numpy()
        except ImportError:
            phi = nl.encode(x)
        out.append(np.hstack((np.asarray(phi), Z[s:s+batch, -1][:, np.newaxis].astype(np.float32))).astype(np.float64))
    return np.vstack(out) if out else np.zeros((0, 0))
