"""ctypes binding of libbetacores.so (the C ABI declared in include/betacores.h).

There is NO CPU path: importing this module without the compiled CUDA library, or calling it
without a CUDA device, raises.  PyTorch is used only for device memory, streams and
torch.distributed; every arithmetic step of the hot path is one of the kernels behind these calls.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('BC_LIB_PATH') or os.path.join(os.path.dirname(_HERE), 'lib', 'libbetacores.so')

c_int, c_i64, c_dbl, c_vp = ctypes.c_int, ctypes.c_int64, ctypes.c_double, ctypes.c_void_p

# name -> argument types (all return int); mirrors include/betacores.h one to one
SIGNATURES = {
    'bc_create': [c_int, ctypes.POINTER(c_vp)],
    'bc_destroy': [c_vp],
    'bc_set_potential': [c_vp, c_int, c_int, c_int, ctypes.POINTER(c_dbl), c_vp],
    'bc_set_samples': [c_vp, c_vp, c_int, c_int, c_vp],
    'bc_rowquad': [c_vp, c_vp, c_i64, c_i64, c_vp, c_vp],
    'bc_project_colsum': [c_vp, c_vp, c_i64, c_vp, c_i64, c_vp, c_vp, c_vp],
    'bc_project_score': [c_vp, c_vp, c_i64, c_vp, c_i64, c_vp, c_vp, c_i64, c_vp, c_vp, c_vp],
    'bc_project_materialise': [c_vp, c_vp, c_i64, c_vp, c_i64, c_vp, c_vp, c_i64, c_vp, c_vp, c_int, c_vp],
    'bc_q_image_bytes': [c_i64, ctypes.POINTER(c_i64)],
    'bc_quantise_rows': [c_vp, c_vp, c_i64, c_i64, c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_vp],
    'bc_q_gather_rows': [c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_vp, c_vp, c_vp, c_vp],
    'bc_feature_exponents': [c_vp, c_vp, c_i64, c_i64, c_int, c_vp, c_vp],
    'bc_set_feature_exponents': [c_vp, c_vp, c_int, c_vp],
    'bc_set_contraction_digits': [c_vp, c_int],
    'bc_project_colsum_q': [c_vp, c_vp, c_vp, c_i64, c_vp, c_vp, c_vp],
    'bc_project_score_q': [c_vp, c_vp, c_vp, c_i64, c_vp, c_vp, c_i64, c_vp, c_vp, c_vp],
    'bc_contraction_q': [c_vp, c_vp, c_vp, c_i64, c_vp, c_i64, c_vp],
    'bc_colsum_combine': [c_vp, c_vp, c_int, c_int, c_vp, c_vp],
    'bc_core_resid': [c_vp, c_vp, c_dbl, c_vp, c_int, c_int, c_i64, c_vp, c_vp, c_vp],
    'bc_core_maxcorr': [c_vp, c_vp, c_int, c_int, c_i64, c_vp, c_int, c_vp, c_vp],
    'bc_core_grad': [c_vp, c_vp, c_int, c_int, c_i64, c_vp, c_vp, c_vp],
    'bc_core_pgrad': [c_vp, c_vp, c_int, c_i64, c_vp, c_vp, c_vp, c_i64, c_vp],
    'bc_dense_pgrad': [c_vp, c_vp, c_int, c_int, c_int, c_vp, c_vp, c_int, c_vp, c_i64, c_vp],
    'bc_laplace_logistic': [c_vp, c_vp, c_i64, c_vp, c_int, c_int, c_vp, c_vp, c_int, c_dbl, c_vp, c_vp],
    'bc_laplace_logistic_factor': [c_vp, c_vp, c_i64, c_vp, c_int, c_int, c_vp, c_vp, c_int, c_dbl, c_vp, c_vp],
    'bc_conjugate_factor': [c_vp, c_int, c_vp, c_i64, c_vp, c_int, c_int, c_vp, c_vp, c_vp, c_dbl, c_vp, c_vp, c_vp, c_vp],
    'bc_sample_solve': [c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_vp, c_int, c_vp],
    'bc_set_sample_slot': [c_vp, c_int],
    'bc_sample_solve_hinted': [c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_vp, c_int, c_vp, c_vp],
    'bc_sample_affine': [c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_vp, c_int, c_vp],
    'bc_adam_step': [c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_dbl, c_dbl, c_dbl, c_dbl, c_dbl, c_dbl, c_vp, c_vp],
    'bc_dense_rownorms': [c_vp, c_vp, c_i64, c_int, c_i64, c_vp, c_vp],
    'bc_dense_center': [c_vp, c_vp, c_i64, c_int, c_i64, c_vp],
    'bc_dense_colsum': [c_vp, c_vp, c_i64, c_int, c_i64, c_vp, c_vp],
    'bc_dense_score': [c_vp, c_int, c_vp, c_i64, c_int, c_i64, c_vp, c_vp, c_vp, c_i64, c_vp, c_vp, c_vp],
    'bc_dense_combine': [c_vp, c_vp, c_i64, c_int, c_vp, c_vp, c_int, c_vp, c_vp],
    'bc_dense_gather': [c_vp, c_vp, c_i64, c_int, c_vp, c_i64, c_vp, c_i64, c_vp],
    'bc_transpose': [c_vp, c_vp, c_i64, c_i64, c_i64, c_vp, c_i64, c_vp],
    'bc_vec_step': [c_vp, c_int, c_vp, c_vp, c_vp, c_int, c_dbl, c_vp, c_vp, c_vp],
    'bc_nnls': [c_vp, c_vp, c_int, c_vp, c_int, c_vp, c_vp, c_vp, c_int, c_vp, c_vp],
    'bc_solver_iterations': [c_vp, c_int, c_int, c_vp, c_i64, c_int, c_i64, c_vp, c_vp, c_vp, c_dbl, c_dbl, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp,
                             c_vp, c_vp],
    'bc_fit_pow_poly': [c_dbl, c_int, ctypes.POINTER(c_dbl), ctypes.POINTER(c_dbl)],
    'bc_fit_pow_tab': [c_dbl, ctypes.POINTER(c_dbl), ctypes.POINTER(c_dbl), ctypes.POINTER(c_dbl), ctypes.POINTER(c_dbl)],
    'bc_set_potential_form': [c_vp, c_int],
    'bc_host_project': [c_int, c_int, c_int, c_int, ctypes.POINTER(c_dbl), c_vp, c_vp, c_i64, c_i64, c_vp, c_int, c_vp, c_int],
}
PLAIN = {'bc_potential_form': ([c_vp], c_int), 'bc_nnls_max_columns': ([], c_int), 'bc_sample_slot': ([c_vp], c_int), 'bc_add_launch_count': ([c_i64], c_i64), 'bc_version': ([], c_int), 'bc_contraction_digits': ([c_vp], c_int), 'bc_q_max_features': ([], c_int), 'bc_launch_count': ([], c_i64), 'bc_last_cuda_error': ([], c_int), 'bc_sm_count': ([c_vp], c_int),
         'bc_colsum_ld': ([c_int], c_int), 'bc_error_string': ([c_int], ctypes.c_char_p)}



class StepArgs(ctypes.Structure):
    """bc_step_args of include/betacores.h"""
    _fields_ = [('d_theta', c_vp), ('S', c_int), ('ldt', c_int),
                ('d_image', c_vp), ('d_rowscale', c_vp), ('d_rowaux_q', c_vp),
                ('d_gimage', c_vp), ('d_growscale', c_vp), ('d_growaux', c_vp),
                ('d_X', c_vp), ('ldx', c_i64), ('d_rowaux', c_vp),
                ('d_rows', c_vp), ('n', c_i64), ('scaling', c_dbl),
                ('d_pts', c_vp), ('ldp', c_i64), ('M', c_int), ('d_pts_rowaux', c_vp),
                ('d_Vc', c_vp), ('ldv', c_i64),
                ('d_parts', c_vp), ('d_colsum', c_vp), ('d_resid', c_vp), ('d_grad', c_vp),
                ('d_w', c_vp), ('d_m1', c_vp), ('d_m2', c_vp),
                ('lr', c_dbl), ('b1', c_dbl), ('b2', c_dbl), ('c1', c_dbl), ('c2', c_dbl), ('eps', c_dbl), ('d_nn_mask', c_vp),
                ('ev_pass_begin', c_vp), ('ev_pass_end', c_vp), ('phase', c_int), ('nparts', c_int), ('d_parts_all', c_vp),
                ('d_sched', c_vp), ('d_step_counter', c_vp)]


SIGNATURES['bc_greedy_opt_step'] = [c_vp, ctypes.POINTER(StepArgs), c_vp]


class MTState(ctypes.Structure):
    """bc_mt_state of include/betacores.h: numpy's legacy RandomState in its own representation"""
    _fields_ = [('key', ctypes.c_uint32*624), ('pos', ctypes.c_int32), ('has_gauss', ctypes.c_int32), ('gauss', c_dbl)]


SIGNATURES['bc_mt_randn'] = [ctypes.POINTER(MTState), c_vp, c_i64, c_int]
SIGNATURES['bc_mt_randint'] = [ctypes.POINTER(MTState), c_i64, c_vp, c_i64]

MODEL_LOGISTIC, MODEL_GAUSSIAN, MODEL_NEURLIN = 0, 1, 2
KIND_LOGLIK, KIND_BETALIK, KIND_BETAGRAD = 0, 1, 2
SCORE_FW, SCORE_GIGA, SCORE_CORR, SCORE_OMP = 0, 1, 2, 3
VEC_GIGA_DIR, VEC_GIGA_STEP, VEC_RESID, VEC_FW_STEP = 0, 1, 2, 3


class NativeError(RuntimeError):
    pass


_lib = None


def lib():
    """Load the shared library (once).  Fails loudly if it was not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NativeError('libbetacores.so not found at %s -- build it with `python beta-cores_b200/build.py` '
                              '(there is no CPU fallback)' % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        for name, args in SIGNATURES.items():
            fn = getattr(L, name)
            fn.argtypes = args
            fn.restype = c_int
        for name, (args, res) in PLAIN.items():
            fn = getattr(L, name)
            fn.argtypes = args
            fn.restype = res
        _lib = L
    return _lib


def check(rc, what):
    if rc != 0:
        L = lib()
        msg = L.bc_error_string(rc).decode()
        if rc == -5:
            msg += ' [cudaError %d]' % L.bc_last_cuda_error()
        raise NativeError('%s failed: %s' % (what, msg))


def call(name, *args):
    check(getattr(lib(), name)(*args), name)


def params8(values):
    arr = (c_dbl * 8)()
    for i, v in enumerate(values):
        arr[i] = float(v)
    return arr
