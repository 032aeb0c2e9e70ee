"""ctypes binding of libdfol_b200.so (C ABI declared in include/dfol_b200.h).

The library is the product: there is no CPU or PyTorch fallback. ``lib()`` raises if the shared object has not
been built (``python -c 'import __graft_entry__ as g; g.build()'`` or ``python -m dfol_vqa_b200.build``).
"""

import ctypes
import os
from ctypes import c_float, c_int, c_int64, c_uint64, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# DFOL_LIB_PATH: another build of the same ABI (A/B timing of kernel variants); the default is the in-tree library
LIB_PATH = os.environ.get('DFOL_LIB_PATH') or os.path.join(_HERE, 'libdfol_b200.so')
_lib = None


class K(object):
    """Constants of include/dfol_b200.h."""
    ABI_VERSION = 6
    ACT_NONE, ACT_ELU, ACT_SIGMOID, ACT_LOGSIGMOID = 0, 1, 2, 3
    MUL_NONE, MUL_SIGMOID_GRAD, MUL_ELU_GRAD = 0, 1, 2
    INSTR_WORDS = 12
    OP_SELECT, OP_FILTER, OP_RELATE, OP_PUSH = 1, 2, 3, 4
    OP_EXIST, OP_AND, OP_OR, OP_VERIFY_ATTRS, OP_CHOOSE_ATTR, OP_CHOOSE_REL = 16, 17, 18, 19, 20, 21
    OP_ALL_SAME, OP_TWO_SAME, OP_COMPARE = 22, 23, 24
    F_NEG, F_ROUNDTRIP, F_SUBJECT, F_NAME_NEG, F_NAME_ROUNDTRIP = 1, 2, 4, 8, 16
    F_NORMALISE, F_NEGATE_RESULT, F_IS_LESS, F_HARD = 32, 64, 128, 256
    OPT_NEG = 1 << 30


TERMINAL_NAMES = ('exist', 'end', 'and', 'or', 'verify_attrs', 'verify_rel', 'choose_attr', 'query_attr',
                  'choose_rel', 'all_same', 'all_different', 'two_same', 'two_different', 'compare')

P = c_void_p
_SIGNATURES = {
    'dfol_version': (c_int, []),
    'dfol_last_error': (ctypes.c_char_p, []),
    'dfol_gemm_f32': (c_int, [P, c_int64, c_int64, P, c_int64, c_int64, P, c_int64, P, c_int, c_int, c_int, c_int,
                              c_int, c_int, P, c_int64, c_int, c_int, P, P, P, P, P, c_float, P]),
    'dfol_gemm_bf16_tc': (c_int, [P, c_int64, P, c_int64, P, c_int64, P, c_int, c_int, c_int, c_int, c_int, c_int, P,
                                  P, P, P, P, c_float, P]),
    'dfol_cast_bf16': (c_int, [P, c_int64, P, c_int64, c_int64, c_int, P]),
    'dfol_box_position': (c_int, [P, c_int64, c_int, P, c_int64, c_int, c_int64, P]),
    'dfol_pair_hidden_fwd': (c_int, [P, c_int64, P, c_int64, P, c_int64, P, P, c_int64, c_int, c_int, c_int, P, P,
                                     P, c_int, c_int, P]),
    'dfol_pair_hidden_bwd': (c_int, [P, c_int64, P, c_int64, P, c_int64, P, c_int64, P, c_int64, P, c_int, c_int,
                                     P, P, P, c_int, P]),
    'dfol_colsum': (c_int, [P, c_int64, c_int64, c_int, P, P]),
    'dfol_act_grad_mul': (c_int, [P, c_int64, P, c_int64, c_int64, c_int, c_int, P]),
    'dfol_program_fwd': (c_int, [P, P, P, c_int, P, P, P, P, P, P, P, P, P, P, c_int, P]),
    'dfol_program_bwd': (c_int, [P, P, P, c_int, P, P, P, P, P, P, P, P, P, P, c_int, P, P, P, P]),
    'dfol_program_fwd_fast': (c_int, [P, P, P, c_int, P, P, P, P, P, P, P, P, P, P, P, c_int, P]),
    'dfol_program_bwd_fast': (c_int, [P, P, P, c_int, P, P, P, P, P, P, P, P, P, P, P, c_int, P, P, P, P]),
    'dfol_loss_fwd_bwd': (c_int, [P, P, P, c_int, c_int, c_int, c_float, P, P, P]),
    'dfol_table_layer_bwd': (c_int, [P, P, P, P, P, c_int, P, P, P, P, P, P, c_int64, P, c_int64, c_int, P,
                                     c_int64, P, P, P]),
    'dfol_table_layer_bwd_fused': (c_int, [P, P, P, P, P, c_int, c_int, P, P, P, P, P, P, c_int64, P, c_int64, c_int,
                                           c_int, P, c_int64, c_int, c_int, P, P, P]),
    'dfol_gemm_bf16_tc_dgrad': (c_int, [P, c_int64, P, c_int64, P, c_int64, c_int, c_int, c_int, c_int, P, c_int64,
                                        c_int, c_float, P]),
    'dfol_rel_slots_fwd': (c_int, [P, c_int64, c_int, P, c_int64, P, P, P, c_int, P, P, P, P, P, c_int, c_int,
                                   c_float, P, P, P]),
    'dfol_pair_layer_fwd_tc': (c_int, [P, c_int64, P, c_int64, P, c_int64, c_int, P, c_int, c_int, c_int, c_int, P,
                                       c_int64, P, P, P, c_int, P, P, P, P, P, c_float, P, P]),
    'dfol_pair_layer_dgrad_tc': (c_int, [P, c_int64, P, c_int64, P, c_int64, c_int, c_int, c_int, c_int, P, c_int64,
                                         c_int, P]),
    'dfol_pair_layer_fwd_cluster': (c_int, [P, c_int64, P, c_int64, P, c_int64, c_int, P, c_int, c_int, c_int, c_int,
                                            P]),
    'dfol_pair_layer_dgrad_cluster': (c_int, [P, c_int64, P, c_int64, P, c_int64, c_int, c_int, c_int, c_int, P,
                                              c_int64, c_int, c_float, P]),
    'dfol_pair_layer_dgrad_wgrad_cluster': (c_int, [P, c_int64, P, c_int64, P, c_int64, c_int, c_int, c_int, c_int, P,
                                                    c_int64, c_int, c_float, P, c_int64, c_int, P]),
    'dfol_cast_jobs': (c_int, [P, c_int, c_int64, P]),
    'dfol_pair_hidden_fwd_mma': (c_int, [P, c_int64, P, c_int64, P, c_int64, P, P, c_int64, c_int, P, P, P, P, c_int,
                                         c_int, P, P]),
    'dfol_rel_slots_fwd_tc': (c_int, [P, c_int64, c_int64, c_int, c_int, P, c_int64, P, P, P, c_int, P, P, P, P, P, c_int,
                                      c_int, c_float, P, P, P, P]),
    'dfol_mod_tape_record_size': (c_int, []),
    'dfol_mod_tape_fwd': (c_int, [P, c_int, c_int, P, P, P, P, P, P, P, P, P, c_int, c_int, P, P, P, P, P]),
    'dfol_mod_tape_bwd': (c_int, [P, c_int, c_int, P, P, P, P, c_int, c_int, P, P, P, P, P, P, P, P]),
    'dfol_lstm_cell_fwd': (c_int, [P, c_int64, P, P, c_int, P, P, P, P, P, P, P, P, P, P, P, c_int, P]),
    'dfol_lstm_cell_bwd': (c_int, [P, P, P, c_int, P, P, P, P, c_int64, P, P, P, P, P, P, c_int, P]),
    'dfol_mod_out_fwd': (c_int, [P, P, P, P, P, c_int, c_int, P, P, c_int, P]),
    'dfol_mod_out_bwd': (c_int, [P, P, P, P, c_int, c_int, P, P, P, c_int, P]),
    'dfol_dropout_scale': (c_int, [P, c_int64, c_int64, c_int, c_int, c_uint64, c_int, c_float, P]),
    'dfol_pair_features_bwd': (c_int, [P, c_int64, c_int, c_int, P, c_int64, P, c_int64, P, P, P, P, c_int64, P]),
    'dfol_pair_features_dropout': (c_int, [P, c_int64, c_int, c_int, P, c_int64, c_int, c_int, P, P, P, P, c_int64,
                                           c_uint64, c_int, c_float, P]),
    'dfol_cast_job_size': (c_int, []),
    'dfol_obj_finish': (c_int, [P, c_int64, c_int, P, c_int64, c_int, P, c_int64, c_int64, P]),
    'dfol_pair_hidden_fwd_tc': (c_int, [P, c_int64, P, c_int64, P, c_int64, P, P, c_int64, c_int, P, P, P, P, c_int,
                                        c_int, P]),
    'dfol_pair_hidden_bwd_tc': (c_int, [P, c_int64, P, P, P, c_int64, P, c_int64, P, c_int, P, P, P, c_int, c_int,
                                        P]),
    'dfol_table_layer_bwd_tc': (c_int, [P, P, P, P, P, c_int, c_int, c_int, P, P, P, P, P, P, c_int64, P, c_int64,
                                        c_int, P, c_int64, c_int, P, P, P, c_float, P]),
    'dfol_table_layer_bwd_mma': (c_int, [P, P, P, P, P, c_int, c_int, P, P, P, P, P, P, c_int, c_int64, P, c_int64, P,
                                         c_int64, c_int, P, c_int64, c_int, P, P, P, P, P]),
    'dfol_pair_chain_fwd': (c_int, [P, c_int64, P, c_int64, P, c_int64, P, P, c_int64, P, P, c_int64, c_int, P, c_int64, P,
                                    P, P, P, P, c_int64, c_int, c_int, P]),
    'dfol_split3_bf16': (c_int, [P, c_int64, c_int64, c_int, P, c_int64, c_int, c_int, c_int, P]),
    'dfol_gemm_bf16_tc_exact': (c_int, [P, c_int64, P, c_int64, P, c_int64, c_int, P, c_int, c_int, c_int, c_int, c_int, P,
                                        P, P, P, P, c_float, P]),
    'dfol_gemm_bf16_tc_wgrad': (c_int, [P, c_int64, P, c_int64, P, c_int64, c_int, c_int, c_int64, P]),
    'dfol_gemm_bf16_tc_wgrad_seg': (c_int, [P, c_int64, P, c_int64, P, c_int64, P, c_int64, P, c_int64, c_int, c_int,
                                            c_int, c_int64, P]),
    'dfol_pair_hidden_bwd_bf16': (c_int, [P, c_int64, P, c_int64, P, c_int64, P, c_int64, P, c_int, P, P, P, c_int,
                                          c_int, P]),
    'dfol_colsum_bf16': (c_int, [P, c_int64, c_int64, c_int, P, P]),
    'dfol_table_grad_dense': (c_int, [P, P, P, P, c_int, P, P, P, P, P, P, c_int64, P]),
    'dfol_answers': (c_int, [P, P, c_int, c_int, c_float, P, P, P, P, P]),
    'dfol_sumsq': (c_int, [P, c_int64, P, P]),
    'dfol_l1_regularize': (c_int, [P, P, c_int64, c_float, c_float, P, P]),
    'dfol_adam_step': (c_int, [P, P, P, P, c_int64, P, c_float, c_float, c_float, c_float, c_float, c_float, c_int,
                               P]),
}


def exported_symbols():
    return sorted(_SIGNATURES)


def lib():
    """The loaded library; raises if it is missing or its ABI version does not match."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError('libdfol_b200.so is not built (%s). Build it with `python -m dfol_vqa_b200.build`; '
                               'there is no fallback path.' % LIB_PATH)
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        if handle.dfol_version() != K.ABI_VERSION:
            raise RuntimeError('libdfol_b200.so ABI version %d != %d' % (handle.dfol_version(), K.ABI_VERSION))
        _lib = handle
    return _lib


def check(rc, what=''):
    if rc != 0:
        raise RuntimeError('%s failed (%d): %s' % (what or 'dfol call', rc, lib().dfol_last_error().decode()))


def ptr(t):
    """Device pointer of a tensor (None -> NULL). The tensor must stay alive until the stream has consumed it."""
    if t is None:
        return None
    return t.data_ptr()


def stream_ptr(device=None):
    return torch.cuda.current_stream(device).cuda_stream


# launch counter: the number of OUR kernels launched (bench.py reports it as gpu_launches)
launches = 0
# optional per-launch CUDA-event trace (bench.py): list of (entry point, meta dict, start event, end event).
# ``meta`` is whatever the caller put in ``next_meta`` just before the call (tag, algorithmic flops / bytes).
trace = None
next_meta = None


def call(name, *args):
    global launches, next_meta
    launches += 1
    if trace is None:
        next_meta = None
        check(getattr(lib(), name)(*args), name)
        return
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    check(getattr(lib(), name)(*args), name)
    e1.record()
    trace.append((name, next_meta or {}, e0, e1))
    next_meta = None
