"""Generated CELT tables (tools/gen_tables.py) vs the reference's pre-computed static tables.

The cross-check against /root/reference only runs where the reference is mounted (this
container); everywhere else the committed sha256 of the generated header pins the content."""
import hashlib
import os
import re
import struct
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HDR = os.path.join(ROOT, 'opus_b200', 'csrc', 'celt_tables_data.h')
REF = '/root/reference'


def _parse_arrays(text):
    out = {}
    for m in re.finditer(r'(?:static const|OB_TABLE)\s+([\w ]+?)\s+(\w+)((?:\[\d*\])+)\s*=\s*\{(.*?)\};', text, re.S):
        body = re.sub(r'/\*.*?\*/', '', m.group(4), flags=re.S)
        body = body.replace('{', ' ').replace('}', ' ')
        vals = [v.strip() for v in body.split(',') if v.strip()]
        out[m.group(2)] = vals
    return out


def _f32(s):
    s = s.rstrip('fF')
    return struct.unpack('I', struct.pack('f', float(s)))[0]


def _ints(v):
    return [int(x.rstrip('uU')) for x in v]


def test_generator_is_reproducible(tmp_path):
    out = tmp_path / 't.h'
    subprocess.check_call([sys.executable, os.path.join(ROOT, 'tools', 'gen_tables.py'), str(out)])
    assert out.read_bytes() == open(HDR, 'rb').read()


def test_header_checksum_pinned():
    want = open(os.path.join(ROOT, 'tests', 'golden', 'celt_tables_data.sha256')).read().split()[0]
    assert hashlib.sha256(open(HDR, 'rb').read()).hexdigest() == want


@pytest.mark.skipif(not os.path.exists(REF + '/celt/static_modes_float.h'), reason='reference not mounted')
def test_tables_match_reference_static_mode():
    mine = _parse_arrays(open(HDR).read())
    txt = open(REF + '/celt/static_modes_float.h').read()
    # only the first (48 kHz) mode: cut at the 96 kHz section
    txt = txt[:txt.index('window240')]
    ref = _parse_arrays(txt)
    for a, b in [('ob_window', 'window120'), ('ob_fft_tw', 'fft_twiddles48000_960'), ('ob_mdct_trig', 'mdct_twiddles960')]:
        assert [_f32(x) for x in mine[a]] == [_f32(x) for x in ref[b]], a
    for a, b in [('ob_logN', 'logN400'), ('ob_cache_index', 'cache_index50'), ('ob_cache_bits', 'cache_bits50'),
                 ('ob_cache_caps', 'cache_caps50'), ('ob_fft_bitrev480', 'fft_bitrev480'),
                 ('ob_fft_bitrev240', 'fft_bitrev240'), ('ob_fft_bitrev120', 'fft_bitrev120'),
                 ('ob_fft_bitrev60', 'fft_bitrev60')]:
        assert _ints(mine[a]) == _ints(ref[b]), a
    # FFT factor lists
    facs = re.findall(r'\{([\d, ]+)\},\s*/\* factors \*/', txt)
    got = _ints(mine['ob_fft_factors'])
    for k, f in enumerate(facs[:4]):
        want = [int(x) for x in f.replace(' ', '').strip(',').split(',')][:10]
        assert got[10 * k:10 * k + 10] == want, (k, want)
    # PVQ U(N,K) rows
    cw = open(REF + '/celt/cwrs.c').read()
    body = cw[cw.index('CELT_PVQ_U_DATA[1272]={'):]
    body = body[body.index('{') + 1:body.index('};')]
    keep, skip = [], False
    for l in body.split('\n'):
        if l.startswith('#if defined(CWRS_EXTRA_ROWS)'):
            skip = True
        elif l.startswith('#endif'):
            skip = False
        elif not skip:
            keep.append(l)
    body = '\n'.join(keep)
    body = re.sub(r'/\*.*?\*/', '', body, flags=re.S)
    refu = [int(x.strip().rstrip('uU')) for x in body.replace('\n', ' ').split(',') if x.strip()]
    assert len(refu) == 1272
    assert _ints(mine['ob_pvq_u']) == refu
    m = open(REF + '/celt/modes.c').read()
    assert _ints(mine['ob_ebands']) == _ints(_parse_arrays(m)['eband5ms'])
    assert _ints(mine['ob_band_alloc']) == _ints(_parse_arrays(m)['band_allocation'])
    q = open(REF + '/celt/quant_bands.c').read()
    qa = _parse_arrays(q)
    assert _ints(mine['ob_e_prob']) == _ints(qa['e_prob_model'])
