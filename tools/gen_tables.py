#!/usr/bin/env python3
"""Generate the constant tables of the Opus/CELT 48 kHz mode (frame 960, overlap 120) from
first principles and write them as one plain-C data header.

Nothing here is copied from the reference tree: every table is recomputed from its defining
formula (RFC 6716 / the reference's mode construction code, cited per table), and
tests/test_tables.py cross-checks the result value-by-value against the reference's
pre-computed static tables whenever /root/reference is mounted.

The only literal inputs are normative bitstream constants of RFC 6716 (band edges, the
allocation matrix, the coarse-energy Laplace model, a few iCDFs).

Float tables: the reference ships its float tables printed with "%#0.8g" (celt/dump_modes/
dump_modes.c:42), i.e. 8 significant digits, which is NOT enough to round-trip a float32.  The
values the reference actually computes with are float32(decimal8(f(x))) with f evaluated in
double (verified: 0 mismatches over all 2880 entries); we reproduce that double rounding so
the tables are bit-identical.
"""
import math
import struct
import sys
import os


# ------------------------------------------------------------------------------------------
# Normative constants (RFC 6716 section 4.3; reference: celt/modes.c:42-63, quant_bands.c:53-138,
# celt/celt.h:194-198, celt/rate.c:42-48, celt/celt.c:320-327)
# ------------------------------------------------------------------------------------------
NB_EBANDS = 21
MAX_LM = 3
SHORT_MDCT = 120
OVERLAP = 120
BITRES = 3
MAX_PSEUDO = 40
MAX_FINE_BITS = 8
FINE_OFFSET = 21
QTHETA_OFFSET = 4
QTHETA_OFFSET_TWOPHASE = 16

EBANDS = [0, 1, 2, 3, 4, 5, 6, 7, 8, 10, 12, 14, 16, 20, 24, 28, 34, 40, 48, 60, 78, 100]

BAND_ALLOCATION = [
    0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,
    90, 80, 75, 69, 63, 56, 49, 40, 34, 29, 20, 18, 10, 0, 0, 0, 0, 0, 0, 0, 0,
    110, 100, 90, 84, 78, 71, 65, 58, 51, 45, 39, 32, 26, 20, 12, 0, 0, 0, 0, 0, 0,
    118, 110, 103, 93, 86, 80, 75, 70, 65, 59, 53, 47, 40, 31, 23, 15, 4, 0, 0, 0, 0,
    126, 119, 112, 104, 95, 89, 83, 78, 72, 66, 60, 54, 47, 39, 32, 25, 17, 12, 1, 0, 0,
    134, 127, 120, 114, 103, 97, 91, 85, 78, 72, 66, 60, 54, 47, 41, 35, 29, 23, 16, 10, 1,
    144, 137, 130, 124, 113, 107, 101, 95, 88, 82, 76, 70, 64, 57, 51, 45, 39, 33, 26, 15, 1,
    152, 145, 138, 132, 123, 117, 111, 105, 98, 92, 86, 80, 74, 67, 61, 55, 49, 43, 36, 20, 1,
    162, 155, 148, 142, 133, 127, 121, 115, 108, 102, 96, 90, 84, 77, 71, 65, 59, 53, 46, 30, 1,
    172, 165, 158, 152, 143, 137, 131, 125, 118, 112, 106, 100, 94, 87, 81, 75, 69, 63, 56, 45, 20,
    200, 200, 200, 200, 200, 200, 200, 200, 198, 193, 188, 183, 178, 173, 168, 163, 158, 153, 148, 129, 104,
]

# Coarse energy Laplace model: [LM][intra][2*band] = (P(0) in Q8, decay in Q8)
E_PROB_MODEL = [
    [[72, 127, 65, 129, 66, 128, 65, 128, 64, 128, 62, 128, 64, 128, 64, 128, 92, 78, 92, 79, 92, 78,
      90, 79, 116, 41, 115, 40, 114, 40, 132, 26, 132, 26, 145, 17, 161, 12, 176, 10, 177, 11],
     [24, 179, 48, 138, 54, 135, 54, 132, 53, 134, 56, 133, 55, 132, 55, 132, 61, 114, 70, 96, 74, 88,
      75, 88, 87, 74, 89, 66, 91, 67, 100, 59, 108, 50, 120, 40, 122, 37, 97, 43, 78, 50]],
    [[83, 78, 84, 81, 88, 75, 86, 74, 87, 71, 90, 73, 93, 74, 93, 74, 109, 40, 114, 36, 117, 34,
      117, 34, 143, 17, 145, 18, 146, 19, 162, 12, 165, 10, 178, 7, 189, 6, 190, 8, 177, 9],
     [23, 178, 54, 115, 63, 102, 66, 98, 69, 99, 74, 89, 71, 91, 73, 91, 78, 89, 86, 80, 92, 66,
      93, 64, 102, 59, 103, 60, 104, 60, 117, 52, 123, 44, 138, 35, 133, 31, 97, 38, 77, 45]],
    [[61, 90, 93, 60, 105, 42, 107, 41, 110, 45, 116, 38, 113, 38, 112, 38, 124, 26, 132, 27, 136, 19,
      140, 20, 155, 14, 159, 16, 158, 18, 170, 13, 177, 10, 187, 8, 192, 6, 175, 9, 159, 10],
     [21, 178, 59, 110, 71, 86, 75, 85, 84, 83, 91, 66, 88, 73, 87, 72, 92, 75, 98, 72, 105, 58,
      107, 54, 115, 52, 114, 55, 112, 56, 129, 51, 132, 40, 150, 33, 140, 29, 98, 35, 77, 42]],
    [[42, 121, 96, 66, 108, 43, 111, 40, 117, 44, 123, 32, 120, 36, 119, 33, 127, 33, 134, 34, 139, 21,
      147, 23, 152, 20, 158, 25, 154, 26, 166, 21, 173, 16, 184, 13, 184, 10, 150, 13, 139, 15],
     [22, 178, 63, 114, 74, 82, 84, 83, 92, 82, 103, 62, 96, 72, 96, 67, 101, 73, 107, 72, 113, 55,
      118, 52, 125, 52, 118, 52, 117, 55, 135, 49, 137, 39, 157, 32, 145, 29, 97, 33, 77, 40]],
]

# Mean band energy (Q4 -> float), RFC 6716 4.3.2.1
E_MEANS_Q4 = [103, 100, 92, 85, 81, 77, 72, 70, 78, 75, 73, 71, 78, 74, 69, 72, 70, 74, 76, 71, 60, 60, 60, 60, 60]

LOG2_FRAC_TABLE = [0, 8, 13, 16, 19, 21, 23, 24, 26, 27, 28, 29, 30, 31, 32, 32, 33, 34, 34, 35, 36, 36, 37, 37]
TRIM_ICDF = [126, 124, 119, 109, 87, 41, 19, 9, 4, 2, 0]
SPREAD_ICDF = [25, 23, 2, 0]
TAPSET_ICDF = [2, 1, 0]
SMALL_ENERGY_ICDF = [2, 1, 0]
TF_SELECT = [[0, -1, 0, -1, 0, -1, 0, -1], [0, -1, 0, -2, 1, 0, 1, -1],
             [0, -2, 0, -3, 2, 0, 1, -1], [0, -2, 0, -3, 3, 0, 1, -1]]


def f32(x):
    return struct.unpack('f', struct.pack('f', float(x)))[0]


def f32_via_8g(x):
    """double -> '%#.8g' -> float32 (see module docstring)."""
    return f32(float('%#.8g' % x))


def ilog(v):
    return int(v).bit_length()


def log2_frac(val, frac):
    """Conservative (rounded-up) log2 with `frac` fractional bits; celt/cwrs.c:48-73."""
    l = ilog(val)
    if val & (val - 1):
        if l > 16:
            val = ((val - 1) >> (l - 16)) + 1
        else:
            val <<= 16 - l
        l = (l - 1) << frac
        while True:
            b = val >> 16
            l += b << frac
            val = (val + b) >> b
            val = (val * val + 0x7FFF) >> 15
            if frac <= 0:
                break
            frac -= 1
        return l + (1 if val > 0x8000 else 0)
    return (l - 1) << frac


# ------------------------------------------------------------------------------------------
# PVQ codebook-size recurrence U(N,K) (RFC 6716 4.3.4.2; reference celt/cwrs.c:197-211)
# ------------------------------------------------------------------------------------------
PVQ_KMAX = [176, 176, 176, 176, 176, 176, 96, 54, 37, 28, 24, 19, 18, 16, 14]  # last column per row


def pvq_u_table():
    maxk = 178
    U = [[0] * (maxk + 1) for _ in range(16)]
    U[0][0] = 1
    for n in range(1, 16):
        for k in range(1, maxk + 1):
            U[n][k] = U[n - 1][k] + U[n][k - 1] + U[n - 1][k - 1]
    data, rowoff = [], []
    for n in range(15):
        rowoff.append(len(data) - n)  # so that data[rowoff[n]+k] = U(n,k) for k>=n
        for k in range(n, PVQ_KMAX[n] + 1):
            v = U[n][k] if n <= k else U[k][n]
            assert v < (1 << 32), (n, k)
            data.append(v)
    return U, data, rowoff


def pvq_v(U, n, k):
    a = U[min(n, k)][max(n, k)]
    b = U[min(n, k + 1)][max(n, k + 1)]
    return a + b


def get_pulses(i):
    return i if i < 8 else (8 + (i & 7)) << ((i >> 3) - 1)


def fits_in32(n, k):
    """V(n,k) < 2^32 ; celt/rate.c:52-69 states it with tables, we evaluate it exactly."""
    # exact big-int evaluation of V(n,k) via the closed recurrence on a small DP
    return _V_big(n, k) < (1 << 32)


_vcache = {}


def _V_big(n, k):
    # V(n,k) = sum_{j} 2^j C(n,j) C(k-1,j-1), number of PVQ codewords
    key = (n, k)
    if key not in _vcache:
        if k == 0:
            _vcache[key] = 1
        else:
            _vcache[key] = sum((1 << j) * math.comb(n, j) * math.comb(k - 1, j - 1) for j in range(1, min(n, k) + 1))
    return _vcache[key]


def pulse_cache(logN):
    """Pulse cache (bits needed for K pulses in N dims, per band and LM) and per-band caps.
    celt/rate.c:72-245."""
    LMs = MAX_LM
    cindex = [-1] * ((LMs + 2) * NB_EBANDS)
    entries = []
    curr = 0
    for i in range(LMs + 2):
        for j in range(NB_EBANDS):
            N = ((EBANDS[j + 1] - EBANDS[j]) << i) >> 1
            found = -1
            for k in range(i + 1):
                stop = False
                for n in range(NB_EBANDS):
                    if k == i and n >= j:
                        break
                    if N == ((EBANDS[n + 1] - EBANDS[n]) << k) >> 1:
                        found = cindex[k * NB_EBANDS + n]
                        stop = True
                        break
                if stop:
                    # the reference keeps scanning k after a hit only because of loop structure;
                    # later hits give the same index (same N => same cache entry)
                    pass
            cindex[i * NB_EBANDS + j] = found
            if found == -1 and N != 0:
                K = 0
                while fits_in32(N, get_pulses(K + 1)) and K < MAX_PSEUDO:
                    K += 1
                cindex[i * NB_EBANDS + j] = curr
                entries.append((N, K, curr))
                curr += K + 1
    bits = [0] * curr
    for (N, K, off) in entries:
        bits[off] = K
        for j in range(1, K + 1):
            bits[off + j] = log2_frac(_V_big(N, get_pulses(j)), BITRES) - 1
    caps = []
    for i in range(LMs + 1):
        for C in (1, 2):
            for j in range(NB_EBANDS):
                N0 = EBANDS[j + 1] - EBANDS[j]
                if (N0 << i) == 1:
                    max_bits = (C * (1 + MAX_FINE_BITS)) << BITRES
                else:
                    LM0 = 0
                    if N0 > 2:
                        N0 >>= 1
                        LM0 -= 1
                    elif N0 <= 1:
                        LM0 = min(i, 1)
                        N0 <<= LM0
                    off = cindex[(LM0 + 1) * NB_EBANDS + j]
                    max_bits = bits[off + bits[off]] + 1
                    N = N0
                    for k in range(i - LM0):
                        max_bits <<= 1
                        offset = ((logN[j] + ((LM0 + k) << BITRES)) >> 1) - QTHETA_OFFSET
                        num = 459 * ((2 * N - 1) * offset + max_bits)
                        den = ((2 * N - 1) << 9) - 459
                        qb = min((num + (den >> 1)) // den, 57)
                        max_bits += qb
                        N <<= 1
                    if C == 2:
                        max_bits <<= 1
                        offset = ((logN[j] + (i << BITRES)) >> 1) - (QTHETA_OFFSET_TWOPHASE if N == 2 else QTHETA_OFFSET)
                        ndof = 2 * N - 1 - (1 if N == 2 else 0)
                        num = (512 if N == 2 else 487) * (max_bits + ndof * offset)
                        den = (ndof << 9) - (512 if N == 2 else 487)
                        qb = min((num + (den >> 1)) // den, 64 if N == 2 else 61)
                        max_bits += qb
                    ndof = C * N + (1 if (C == 2 and N > 2) else 0)
                    offset = ((logN[j] + (i << BITRES)) >> 1) - FINE_OFFSET
                    if N == 2:
                        offset += (1 << BITRES) >> 2
                    num = max_bits + ndof * offset
                    den = (ndof - 1) << BITRES
                    qb = min((num + (den >> 1)) // den, MAX_FINE_BITS)
                    max_bits += (C * qb) << BITRES
                max_bits = (4 * max_bits // (C * ((EBANDS[j + 1] - EBANDS[j]) << i))) - 64
                assert 0 <= max_bits < 256
                caps.append(max_bits)
    return cindex, bits, caps


# ------------------------------------------------------------------------------------------
# FFT plan: factorisation, digit-reversal permutation, twiddles (celt/kiss_fft.c:322-430)
# ------------------------------------------------------------------------------------------
def fft_factors(n):
    p = 4
    fac = []
    nn = n
    while True:
        while nn % p:
            p = {4: 2, 2: 3}.get(p, p + 2)
            if p > 32000 or p * p > nn:
                p = nn
        nn //= p
        assert p <= 5
        fac.append(p)
        if p == 2 and len(fac) > 2:  # 'stages > 1' evaluated before stages++
            fac[-1] = 4
            fac[1] = 2
        if nn <= 1:
            break
    fac = fac[::-1]
    out = []
    m = n
    for p in fac:
        m //= p
        out.append((p, m))
    return out


def fft_bitrev(n, factors):
    table = [0] * n

    def rec(fout, f, fstride, facs):
        p, m = facs[0]
        if m == 1:
            for j in range(p):
                table[f] = fout + j
                f += fstride
        else:
            for j in range(p):
                rec(fout, f, fstride * p, facs[1:])
                f += fstride
                fout += m
    rec(0, 0, 1, factors)
    return table


def main(out_path):
    U, pvq_data, pvq_rowoff = pvq_u_table()
    logN = [log2_frac(EBANDS[i + 1] - EBANDS[i], BITRES) for i in range(NB_EBANDS)]
    cindex, cbits, ccaps = pulse_cache(logN)

    # window (celt/modes.c:  sin(pi/2 * sin^2(pi/2 (i+.5)/overlap)) ), Vorbis power-complementary
    window = []
    for i in range(OVERLAP):
        s = math.sin(.5 * math.pi * (i + .5) / OVERLAP)
        window.append(f32_via_8g(math.sin(.5 * math.pi * s * s)))

    # FFT twiddles for nfft=480 (shared by the smaller sizes through a stride)
    nfft = 480
    tw = []
    for i in range(nfft):
        phase = (-2 * 3.14159265358979323846264338327 / nfft) * i
        tw.append((f32_via_8g(math.cos(phase)), f32_via_8g(math.sin(phase))))
    plans = {}
    for sh in range(4):
        n = nfft >> sh
        fac = fft_factors(n)
        plans[n] = (fac, fft_bitrev(n, fac))

    # MDCT twiddles: for N = 1920, 960, 480, 240: cos(2 pi (i+1/8)/N), i < N/2  (celt/mdct.c:56-106)
    PI = 3.1415926535897931
    trig = []
    N = 1920
    for sh in range(4):
        for i in range(N // 2):
            trig.append(f32_via_8g(math.cos(2 * PI * (i + .125) / N)))
        N >>= 1
    assert len(trig) == 1800

    def carr(name, ctype, vals, fmt, per_line=12):
        s = 'OB_TABLE %s %s[%d] = {\n' % (ctype, name, len(vals))
        for i in range(0, len(vals), per_line):
            s += '  ' + ', '.join(fmt(v) for v in vals[i:i + per_line]) + ',\n'
        return s + '};\n\n'

    ff = lambda v: ('%.9g' % v) + ('f' if ('.' in '%.9g' % v or 'e' in '%.9g' % v) else '.f')
    fi = lambda v: '%d' % v
    fu = lambda v: '%uu' % v

    o = []
    o.append('/* GENERATED by tools/gen_tables.py -- do not edit.  Constant tables of the Opus CELT\n'
             ' * 48 kHz mode (N=960, overlap 120, 21 bands), recomputed from their defining formulas.\n'
             ' * The includer defines OB_TABLE (e.g. `static const` or `static __device__ const`). */\n')
    o.append('#ifndef OB_TABLE\n#error "define OB_TABLE before including celt_tables_data.h"\n#endif\n'
             '#ifndef OB_CELT_TABLES_DATA_H\n#define OB_CELT_TABLES_DATA_H\n\n')
    o.append(carr('ob_ebands', 'short', EBANDS, fi, 22))
    o.append(carr('ob_band_alloc', 'unsigned char', BAND_ALLOCATION, fi, 21))
    o.append(carr('ob_logN', 'short', logN, fi, 21))
    o.append(carr('ob_cache_index', 'short', cindex, fi, 15))
    o.append(carr('ob_cache_bits', 'unsigned char', cbits, fi, 20))
    o.append(carr('ob_cache_caps', 'unsigned char', ccaps, fi, 21))
    o.append(carr('ob_pvq_u', 'unsigned int', pvq_data, fu, 8))
    o.append(carr('ob_pvq_rowoff', 'short', pvq_rowoff, fi, 15))
    eprob = [v for lm in E_PROB_MODEL for intra in lm for v in intra]
    o.append(carr('ob_e_prob', 'unsigned char', eprob, fi, 21))
    o.append(carr('ob_emeans', 'float', [q / 16.0 for q in E_MEANS_Q4], ff, 10))
    o.append(carr('ob_log2_frac_tab', 'unsigned char', LOG2_FRAC_TABLE, fi, 24))
    o.append(carr('ob_trim_icdf', 'unsigned char', TRIM_ICDF, fi, 11))
    o.append(carr('ob_spread_icdf', 'unsigned char', SPREAD_ICDF, fi, 4))
    o.append(carr('ob_tapset_icdf', 'unsigned char', TAPSET_ICDF, fi, 4))
    o.append(carr('ob_small_energy_icdf', 'unsigned char', SMALL_ENERGY_ICDF, fi, 4))
    o.append(carr('ob_tf_select', 'signed char', [v for r in TF_SELECT for v in r], fi, 8))
    o.append(carr('ob_window', 'float', window, ff, 6))
    o.append(carr('ob_fft_tw', 'float', [v for p in tw for v in p], ff, 6))
    for n in (480, 240, 120, 60):
        o.append(carr('ob_fft_bitrev%d' % n, 'short', plans[n][1], fi, 20))
    o.append(carr('ob_mdct_trig', 'float', trig, ff, 6))
    # factor lists as (radix, m) pairs, 5 stages max, zero padded
    facs = []
    for n in (480, 240, 120, 60):
        fl = [v for pm in plans[n][0] for v in pm]
        fl += [0] * (10 - len(fl))
        facs += fl
    o.append(carr('ob_fft_factors', 'short', facs, fi, 10))
    o.append('#endif /* OB_CELT_TABLES_DATA_H */\n')
    with open(out_path, 'w') as f:
        f.write(''.join(o))
    print('wrote', out_path, 'pvq', len(pvq_data), 'cache', len(cbits))


if __name__ == '__main__':
    here = os.path.dirname(os.path.abspath(__file__))
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(here, '..', 'opus_b200', 'csrc', 'celt_tables_data.h'))
