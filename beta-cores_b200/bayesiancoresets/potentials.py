"""Device potentials: the likelihood callbacks of the reference, as objects the projectors can
dispatch on.

A `DevicePotential` is callable with the reference function's own signature and returns the same
(n, S) ndarray -- computed by the materialise kernel, through bc_project_materialise -- so it
drops in wherever the reference passes a likelihood function.  When a projector receives one
(fully bound), the coreset classes skip the (n, S) matrix altogether and run the fused kernels.
"""
import math
import numpy as np

from . import _native as nv

MODELS = {'logistic': nv.MODEL_LOGISTIC, 'gaussian': nv.MODEL_GAUSSIAN, 'neurlin': nv.MODEL_NEURLIN}
KINDS = {'loglik': nv.KIND_LOGLIK, 'betalik': nv.KIND_BETALIK, 'betagrad': nv.KIND_BETAGRAD}


class DevicePotential(object):
    """model in {'logistic','gaussian','neurlin'}; kind in {'loglik','betalik','betagrad'}.

    `needs` lists the model constants that must be bound (Gaussian: Siginv, logdetSig;
    neural-linear: sigsq) before the potential can run fused; calling the object with the
    reference's full positional signature binds them on the fly.
    """

    def __init__(self, model, kind, bound=None, name=None):
        self.model = model
        self.kind = kind
        self.model_id = MODELS[model]
        self.kind_id = KINDS[kind]
        self.bound = dict(bound or {})
        self.__name__ = name or '%s_%s' % (model, kind)

    needs_by_model = {'logistic': (), 'gaussian': ('Siginv', 'logdetSig'), 'neurlin': ('sigsq',)}

    @property
    def needs(self):
        return self.needs_by_model[self.model]

    @property
    def takes_beta(self):
        return self.kind in ('betalik', 'betagrad')

    def is_bound(self):
        return all(k in self.bound for k in self.needs)

    def bind(self, **consts):
        """Return a copy with model constants fixed, e.g. gaussian_beta_likelihood.bind(Siginv=..., logdetSig=...).
        Use this where the reference drivers wrap the function in a lambda
        (examples/zellner_gaussian/main.py:58-67): a lambda hides the potential from the projector."""
        b = dict(self.bound)
        for k, v in consts.items():
            if k not in self.needs:
                raise TypeError('%s has no constant %r' % (self.__name__, k))
            b[k] = np.ascontiguousarray(v, dtype=np.float64) if k == 'Siginv' else float(v)
        return DevicePotential(self.model, self.kind, b, self.__name__)

    def feature_dim(self, ncols):
        """contraction length D for data rows with `ncols` columns (neural-linear rows are [phi, y])"""
        return ncols - 1 if self.model == 'neurlin' else ncols

    def params(self, D, beta=None):
        """The 8 scalars the epilogue functor reads (bc_models.cuh), formed with the reference's
        own expressions so constants round identically."""
        p = [0.0] * 8
        if self.takes_beta and beta is None:
            raise TypeError('%s needs beta' % self.__name__)
        if self.model == 'logistic':
            if self.takes_beta:
                p[0] = beta
                p[1] = (beta+1.)/beta                                     # model_lr.py:85
        elif self.model == 'gaussian':
            d = float(D)
            logdet = self.bound['logdetSig']
            p[0] = -D/2*np.log(2*np.pi) - 1./2.*logdet                    # gaussian.py:13
            if self.takes_beta:
                p[1] = 1./beta                                            # gaussian.py:42
                p[2] = -.5*beta
                p[3] = (1+beta)**(-.5*d-1)                                # gaussian.py:43
                p[4] = np.log((2*np.pi)**(-.5*d)*(np.exp(logdet)**(-.5)))  # gaussian.py:54 logcnst
                p[5] = 1./(beta)**2                                       # gaussian.py:59
                p[6] = 1./(2.*beta)                                       # gaussian.py:60
                p[7] = (1+beta)**(-.5*d-1.)*np.log(1.+beta)               # gaussian.py:61
        else:
            s2 = self.bound['sigsq']
            p[0] = -1./2.*np.log(2.*np.pi*s2)                             # model_neurlinr.py:96
            p[1] = 1./(2.*s2)
            if self.takes_beta:
                p[2] = 1./(2*np.pi*s2)**(beta/2.)                         # model_neurlinr.py:108
                p[3] = -(beta+1.)/beta
                p[4] = -beta/(2.*s2)
                p[5] = 1./np.sqrt(1.+beta)
        return [float(v) for v in p]

    # ---- reference-signature call: returns the UN-centred (n, S) matrix as a host array ----
    def __call__(self, pts, samples, *rest):
        rest = list(rest)
        beta = rest.pop(0) if self.takes_beta else None
        pot = self
        missing = [k for k in self.needs if k not in self.bound]
        if rest:
            if len(rest) != len(missing):
                raise TypeError('%s: expected %d model constants %s, got %d' % (self.__name__, len(missing), missing, len(rest)))
            pot = self.bind(**dict(zip(missing, rest)))
        elif missing:
            raise TypeError('%s: unbound model constants %s' % (self.__name__, missing))
        from .coreset.projector import evaluate_potential
        return evaluate_potential(pot, pts, samples, beta, centred=False)


class DeviceGradient(object):
    """d(log-likelihood)/d(point) of a built-in model, as the object `BlackBoxProjector(..., grad_loglikelihood=...)`
    takes (examples/common/model_lr.py:107-114 grad_z_log_likelihood, gaussian.py:17-20 gaussian_grad_x_loglikelihood).

    BatchPSVICoreset recognises it and contracts the gradient with the residual on the device
    (bc_core_pgrad) without forming the (M, S, D) tensor of the reference.  Called directly with the reference's
    signature it returns that tensor (built on the device from the same kernel, one unit residual at a time is not needed:
    the tensor is sigma(m)[:, :, None] * th[None] resp. (th Siginv)[None] - (x Siginv)[:, None], assembled from the
    materialise pass) -- only meant for small inputs."""

    def __init__(self, model, bound=None, name=None):
        if model not in ('logistic', 'gaussian'):
            raise ValueError('no point-gradient for model %r (the reference defines none, model_neurlinr.py:99-100)' % (model,))
        self.model = model
        self.bound = dict(bound or {})
        self.__name__ = name or '%s_grad_point' % model

    @property
    def needs(self):
        return ('Siginv',) if self.model == 'gaussian' else ()

    def is_bound(self):
        return all(k in self.bound for k in self.needs)

    def bind(self, **consts):
        b = dict(self.bound)
        for k, v in consts.items():
            if k not in self.needs:
                raise TypeError('%s has no constant %r' % (self.__name__, k))
            b[k] = np.ascontiguousarray(v, dtype=np.float64)
        return DeviceGradient(self.model, b, self.__name__)

    def __call__(self, pts, samples, *rest):
        raise NotImplementedError('%s is a device gradient descriptor: BatchPSVICoreset contracts it on the device; the (n, S, D) '
                                  'tensor of the reference is never formed' % self.__name__)
