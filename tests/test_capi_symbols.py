"""CPU-side checks of the drop-in boundary: the C-ABI library builds for sm_100a, loads, and
exports every symbol include/opus_b200.h declares; creating a decoder without a GPU fails loudly."""
import ctypes
import os
import re

import pytest

import util
from opus_b200 import capi


@pytest.fixture(scope='module')
def libpath():
    return capi.build_library()


def test_header_symbols_exported(libpath):
    hdr = open(os.path.join(util.ROOT, 'include', 'opus_b200.h')).read()
    hdr = re.sub(r'/\*.*?\*/', '', hdr, flags=re.S)
    declared = set(re.findall(r'\b(opusb200_\w+)\s*\(', hdr))
    assert declared == set(capi.SYMBOLS), declared ^ set(capi.SYMBOLS)
    L = ctypes.CDLL(libpath)
    for s in declared:
        assert hasattr(L, s), s


def test_kernels_are_sm100a(libpath):
    import subprocess
    out = subprocess.run(['cuobjdump', '-lelf', libpath], capture_output=True, text=True).stdout
    assert 'sm_100a' in out, out


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    import opus_b200
    with pytest.raises(opus_b200.OpusB200Error):
        opus_b200.DecoderBatch(4)


def test_product_does_not_touch_oracle():
    """Nothing under opus_b200/ or include/ may reference oracle/ (the checker)."""
    for d in ('opus_b200', 'include'):
        for root, _, files in os.walk(os.path.join(util.ROOT, d)):
            for f in files:
                if f.endswith(('.py', '.cu', '.cuh', '.h', '.cpp')):
                    txt = open(os.path.join(root, f), errors='ignore').read()
                    assert 'liboracle' not in txt and 'oracle/' not in txt.replace('the reference oracle', ''), os.path.join(root, f)


def test_header_is_plain_c_and_links(libpath, tmp_path):
    """include/opus_b200.h must be usable from C (the reference is C): compile a C89-pedantic translation unit that takes the address of
    every declared entry point, and link it against the library (no CUDA headers, no C++)."""
    import subprocess
    hdr = open(os.path.join(util.ROOT, 'include', 'opus_b200.h')).read()
    names = sorted(set(re.findall(r'\b(opusb200_\w+)\s*\(', re.sub(r'/\*.*?\*/', '', hdr, flags=re.S))))
    src = tmp_path / 'use_header.c'
    src.write_text('#include "opus_b200.h"\n#include <stdio.h>\ntypedef void (*fn)(void);\nint main(void) {\n  fn f[] = {\n' +
                   ''.join('    (fn)%s,\n' % n for n in names) +
                   '  };\n  printf("%d entry points, %s\\n", (int)(sizeof f / sizeof f[0]), opusb200_version());\n  return f[0] == 0;\n}\n')
    exe = tmp_path / 'use_header'
    r = subprocess.run(['gcc', '-std=c99', '-Wall', '-Wextra', '-pedantic', '-Werror', '-Wno-pedantic', '-I', os.path.join(util.ROOT, 'include'),
                        str(src), '-o', str(exe), libpath, '-Wl,-rpath,' + os.path.dirname(libpath)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and 'entry points' in r.stdout, (r.stdout, r.stderr)
