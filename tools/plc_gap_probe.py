"""Concealed-stream differences on the GPU, stream by stream, next to the reference's own spread: for the populations of tests/test_gpu_plc.py prints,
per lossy stream, max |gpu - ref_c|, max |ref_simd - ref_c| (the reference's C and SSE/AVX2 builds on the same packets and losses), and the frame
at which the GPU first departs by more than 1e-5 (is it a lost frame?)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import opus_b200, util
from test_gpu_plc import _streams

def ref_decode(kind, Fs, ch, fsd, pk, sz):
    L = util.load_ref(kind)
    T = len(sz)
    p, z = np.ascontiguousarray(pk), np.ascontiguousarray(sz, np.int32)
    ref, rr = np.zeros((T * fsd, ch), np.float32), np.zeros(T, np.uint32)
    assert L.refh_decode_stream_float(Fs, ch, fsd, T, util._ptr(p), p.shape[1], util._ptr(z), util._ptr(ref), util._ptr(rr)) == 0
    return ref

rows = []
for ch, fs48, Fs in [(2, 960, 48000), (1, 960, 48000), (2, 480, 48000), (2, 240, 48000), (1, 120, 48000), (2, 960, 16000)]:
    n, nfr = 40, 30
    rng = np.random.default_rng(ch * 100 + fs48 + Fs // 1000)
    pk, sz = _streams(n, ch, fs48, nfr, seed=fs48 + ch)
    lost = rng.random((n, nfr)) < 0.18
    lost[:4] = False; lost[4, 6:16] = True; lost[5, 0:2] = True; lost[6, 3::2] = True
    sz = np.where(lost, 0, sz).astype(np.int32)
    fsd = fs48 * Fs // 48000
    dec = opus_b200.DecoderBatch(n, channels=ch, Fs=Fs)
    outs = np.zeros((n, nfr * fsd, ch), np.float32)
    for t in range(nfr):
        got, ns = dec.decode(np.ascontiguousarray(pk[:, t]), np.ascontiguousarray(sz[:, t]), frame_size=fsd, dtype=np.float32)
        outs[:, t * fsd:(t + 1) * fsd] = got.reshape(n, fsd, ch)
    dec.close()
    for s in range(n):
        if not lost[s].any():
            continue
        rc = ref_decode('c', Fs, ch, fsd, pk[s], sz[s])
        rs = ref_decode('simd', Fs, ch, fsd, pk[s], sz[s]) if util.ref_available('simd') else rc
        eg = np.abs(outs[s] - rc).reshape(nfr, -1).max(1)
        es = np.abs(rs - rc).reshape(nfr, -1).max(1)
        first = int(np.argmax(eg > 1e-5)) if (eg > 1e-5).any() else -1
        rows.append((eg.max(), es.max(), ch, fs48, Fs, s, first, bool(lost[s][first]) if first >= 0 else None, float(np.abs(rc).max())))
rows.sort(reverse=True)
print('gpu_err   simd_err  ratio   ch fs48 Fs stream first_bad_frame lost? peak')
for r in rows[:25]:
    print('%.2e  %.2e  %6.1f  %d %4d %5d %3d %3d %s %.2f' % (r[0], r[1], r[0] / max(r[1], 1e-9), r[2], r[3], r[4], r[5], r[6], r[7], r[8]))
g = np.array([r[0] for r in rows]); s = np.array([r[1] for r in rows])
print('streams %d  gpu: max %.2e median %.2e   simd: max %.2e median %.2e   gpu <= 8 x simd + 2e-5 on %d of %d' % (len(rows), g.max(), np.median(g), s.max(), np.median(s), int((g <= 8 * s + 2e-5).sum()), len(rows)))
