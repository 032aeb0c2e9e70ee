"""The C-ABI library loads and exports every symbol include/dfol_b200.h declares (no compute calls: CPU only)."""

import ctypes
import os
import re

import helpers  # noqa: F401  (path setup)
from dfol_vqa_b200 import capi

HEADER = os.path.join(helpers.REPO, 'include', 'dfol_b200.h')


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(dfol_[a-z0-9_]+)\s*\(', text)))


def test_header_symbols_are_exported():
    if not os.path.exists(capi.LIB_PATH):
        from dfol_vqa_b200 import build
        build.build()
    handle = ctypes.CDLL(capi.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 18
    for name in names:
        assert hasattr(handle, name), name


def test_binding_covers_header():
    assert sorted(capi.exported_symbols()) == declared_symbols()
    assert capi.lib().dfol_version() == capi.K.ABI_VERSION
    assert capi.lib().dfol_last_error() is not None


def test_constants_match_header():
    text = open(HEADER).read()
    defs = dict(re.findall(r'#define\s+DFOL_([A-Z0-9_]+)\s+\(?([0-9<> ]+)\)?', text))
    pairs = {'ACT_ELU': capi.K.ACT_ELU, 'ACT_SIGMOID': capi.K.ACT_SIGMOID, 'ACT_LOGSIGMOID': capi.K.ACT_LOGSIGMOID,
             'INSTR_WORDS': capi.K.INSTR_WORDS, 'OP_SELECT': capi.K.OP_SELECT, 'OP_FILTER': capi.K.OP_FILTER,
             'OP_RELATE': capi.K.OP_RELATE, 'OP_PUSH': capi.K.OP_PUSH, 'OP_EXIST': capi.K.OP_EXIST,
             'OP_AND': capi.K.OP_AND, 'OP_OR': capi.K.OP_OR, 'OP_VERIFY_ATTRS': capi.K.OP_VERIFY_ATTRS,
             'OP_CHOOSE_ATTR': capi.K.OP_CHOOSE_ATTR, 'OP_CHOOSE_REL': capi.K.OP_CHOOSE_REL,
             'OP_ALL_SAME': capi.K.OP_ALL_SAME, 'OP_TWO_SAME': capi.K.OP_TWO_SAME, 'OP_COMPARE': capi.K.OP_COMPARE,
             'F_NEG': capi.K.F_NEG, 'F_ROUNDTRIP': capi.K.F_ROUNDTRIP, 'F_SUBJECT': capi.K.F_SUBJECT,
             'F_NAME_NEG': capi.K.F_NAME_NEG, 'F_NAME_ROUNDTRIP': capi.K.F_NAME_ROUNDTRIP,
             'F_NORMALISE': capi.K.F_NORMALISE, 'F_NEGATE_RESULT': capi.K.F_NEGATE_RESULT,
             'F_IS_LESS': capi.K.F_IS_LESS, 'F_HARD': capi.K.F_HARD, 'ABI_VERSION': capi.K.ABI_VERSION}
    for name, value in pairs.items():
        assert eval(defs[name]) == value, name
