"""Projected ADAM (bayesiancoresets/util/opt.py:36-77) with the update on the device.

`grd(x)` receives the current iterate as a HOST ndarray (the coreset classes' gradient has to hand
it to the user's sampler) and may return either a host ndarray or a CUDA tensor; moments, the ADAM
update and the clamp run in bc_adam_step, bit-compatible with the numpy expression order.
"""
import ctypes
import numpy as np
import torch

from .. import _native as nv
from .._device import Engine, ptr, stream_ptr


def _adam_loop(x0, grd, nn_idcs, opt_itrs, step_sched, b1, b2, eps):
    eng = Engine.get()
    x0 = np.asarray(x0, dtype=np.float64)
    n = x0.shape[0]
    if n == 0:
        for i in range(opt_itrs):
            grd(x0.copy())          # keep callback / RNG call counts identical to the reference
        return x0.copy()
    x = eng.upload(x0)
    m1 = eng.zeros(n)
    m2 = eng.zeros(n)
    mask = None
    if nn_idcs is not None:
        mk = np.zeros(n, dtype=np.uint8)
        mk[np.asarray(nn_idcs, dtype=np.int64)] = 1
        mask = eng.upload(mk, dtype=torch.uint8)
    xh = x0.copy()
    ctx = eng.ctx()
    for i in range(opt_itrs):
        g = grd(xh, x) if getattr(grd, 'wants_device_iterate', False) else grd(xh)
        if not isinstance(g, torch.Tensor):
            g = eng.upload(np.asarray(g, dtype=np.float64))
        c1 = 1.-b1**(i+1)
        c2 = 1.-b2**(i+1)
        nv.call('bc_adam_step', ctx, ptr(g), ptr(x), ptr(m1), ptr(m2), n, float(step_sched(i)), b1, b2, c1, c2, eps,
                ptr(mask), stream_ptr())
        xh = x.cpu().numpy()        # one D2H + sync per step: the next sampler call needs the weights
    return xh


def nn_opt(x0, grd, opt_itrs=1000, step_sched=lambda i: 1./(i+1), b1=0.9, b2=0.999, eps=1e-8, verbose=False):
    return _adam_loop(x0, grd, None, opt_itrs, step_sched, b1, b2, eps)


def partial_nn_opt(x0, grd, nn_idcs, opt_itrs=1000, step_sched=lambda i: 1./(i+1), b1=0.9, b2=0.999, eps=1e-8, verbose=False):
    return _adam_loop(x0, grd, nn_idcs, opt_itrs, step_sched, b1, b2, eps)
