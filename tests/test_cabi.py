"""CPU: the C-ABI library loads and exports every symbol declared in include/unirec_b200.h, and the ctypes
signature table agrees with the header (argument count and pointer/int/int64/float kinds).  No compute calls."""
import os
import re

from unirec_b200 import _cabi

HEADER = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'include', 'unirec_b200.h')


def _declarations():
    text = open(HEADER).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    decls = {}
    for m in re.finditer(r'\bint\s+(ur_\w+)\s*\(([^;]*?)\)\s*;', text, flags=re.S):
        name, args = m.group(1), m.group(2).strip()
        kinds = ''
        if args and args != 'void':
            for a in args.split(','):
                a = a.strip()
                if '*' in a:
                    kinds += 'p'
                elif a.startswith('int64_t'):
                    kinds += 'l'
                elif a.startswith('float'):
                    kinds += 'f'
                elif a.startswith('int'):
                    kinds += 'i'
                else:
                    raise AssertionError('unparsed argument %r in %s' % (a, name))
        decls[name] = kinds
    return decls


def test_header_and_binding_table_agree():
    decls = _declarations()
    assert set(decls) == set(_cabi.SIGNATURES), set(decls) ^ set(_cabi.SIGNATURES)
    for name, kinds in decls.items():
        assert _cabi.SIGNATURES[name] == kinds, (name, _cabi.SIGNATURES[name], kinds)


def test_library_loads_and_exports_every_symbol():
    lib = _cabi.lib()
    for name in _declarations():
        assert hasattr(lib, name), name
    assert lib.ur_version() == 4


def test_every_entry_point_cites_the_reference():
    text = open(HEADER).read()
    assert text.count('replaces:') >= 10
