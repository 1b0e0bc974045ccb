"""Where does the GPU decoder's float PCM depart from the reference?  Decodes the population of tests/test_gpu_decode.py on the GPU, compares with the
oracle (bit-identical to the reference's C build), and saves per-stream / per-frame error statistics plus the worst streams' signals for offline analysis."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import opus_b200, util
from test_gpu_decode import _make_streams

out = {}
for ch, fs in [(2, 960), (1, 960), (2, 480)]:
    n, nfr = 48, 12
    pk, sz = _make_streams(n, ch, fs, nfr, seed=ch * 10 + fs)
    dec = opus_b200.DecoderBatch(n, channels=ch)
    outs = np.zeros((n, nfr * fs, ch), np.float32)
    for t in range(nfr):
        got, ns = dec.decode(pk[:, t], sz[:, t], frame_size=fs, dtype=np.float32)
        outs[:, t * fs:(t + 1) * fs] = got.reshape(n, fs, ch)
    dec.close()
    refs = np.stack([util.oracle_decode_float(pk[s], sz[s], ch, fs)[0].reshape(nfr * fs, ch) for s in range(n)])
    err = np.abs(outs - refs)
    per_frame = err.reshape(n, nfr, fs, ch).max(axis=(2, 3))
    worst = int(err.max(axis=(1, 2)).argmax())
    print(ch, fs, 'max err %.3g  median stream max %.3g  worst stream %d  frames %s' % (err.max(), np.median(err.max(axis=(1, 2))), worst,
          np.array2string(per_frame[worst], precision=2)))
    key = 'c%d_f%d' % (ch, fs)
    out[key + '_per_frame'] = per_frame
    out[key + '_worst_gpu'] = outs[worst]
    out[key + '_worst_ref'] = refs[worst]
    out[key + '_worst_pk'] = pk[worst]
    out[key + '_worst_sz'] = sz[worst]
os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
np.savez_compressed(os.path.join(ROOT, 'gpurun_out', 'float_gap_probe.npz'), **out)
