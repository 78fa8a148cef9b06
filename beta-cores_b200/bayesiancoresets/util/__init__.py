from .opt import nn_opt, partial_nn_opt
from .log import set_verbosity
from .errors import NumericalPrecisionError

TOL = 1e-12


def set_tolerance(tol):
    global TOL
    TOL = tol
