"""Builds and drives tests/emu/libemu_decode.so: the decoder KERNEL SOURCES compiled for the
host SIMT emulator (test-only; see tests/emu/emu_runtime.h)."""
import ctypes
import os
import subprocess

import numpy as np

import util

EMU_DIR = os.path.join(util.ROOT, 'tests', 'emu')
_lib = None


def build_emu(name='emu_decode'):
    src = os.path.join(EMU_DIR, name + '.cpp')
    out = os.path.join(EMU_DIR, 'lib%s.so' % name)
    deps = [src, os.path.join(EMU_DIR, 'emu_runtime.h')] + [
        os.path.join(util.ROOT, 'opus_b200', 'csrc', f) for f in os.listdir(os.path.join(util.ROOT, 'opus_b200', 'csrc'))]
    def fresh():
        return os.path.exists(out) and all(os.path.getmtime(out) >= os.path.getmtime(d) for d in deps)
    if fresh():
        return out
    import fcntl
    with open(out + '.lock', 'w') as lk:   # pytest-xdist workers: one builds, the others wait and then find it fresh
        fcntl.flock(lk, fcntl.LOCK_EX)
        if not fresh():
            tmp = out + '.tmp.%d' % os.getpid()
            subprocess.check_call(['g++', '-O1', '-g', '-std=c++17', '-fPIC', '-shared', '-ffp-contract=off', '-Wno-unknown-pragmas',
                                   '-I' + os.path.join(util.ROOT, 'opus_b200', 'csrc'), '-I' + EMU_DIR, '-o', tmp, src])
            os.replace(tmp, out)
    return out


_ops_lib = None


def load_emu_ops():
    global _ops_lib
    if _ops_lib is None:
        _ops_lib = ctypes.CDLL(build_emu('emu_ops'))
    return _ops_lib


def load_emu():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build_emu())
        _lib.emu_decoder_create.restype = ctypes.c_void_p
    return _lib


class EmuDecoder:
    def __init__(self, n_streams, channels, Fs=48000):
        self.L = load_emu()
        self.n, self.ch = n_streams, channels
        self.h = ctypes.c_void_p(self.L.emu_decoder_create(n_streams, channels))
        self.L.emu_decoder_set_fs(self.h, Fs)

    def decode(self, packets, lens, int16=False, frame_size=5760):
        """packets [n][stride] uint8, lens [n] (0 = lost: frame_size samples are concealed) -> (pcm [n][5760*ch], samples[n], final_range[n])"""
        self.L.emu_decoder_set_frame_size(self.h, frame_size)
        packets = np.ascontiguousarray(packets, np.uint8)
        lens = np.ascontiguousarray(lens, np.int32)
        stride = 5760 * self.ch
        pcm = np.zeros((self.n, stride), np.int16 if int16 else np.float32)
        ns = np.zeros(self.n, np.int32)
        fr = np.zeros(self.n, np.uint32)
        r = self.L.emu_decoder_decode(self.h, util._ptr(packets), packets.shape[1], util._ptr(lens), util._ptr(pcm), stride,
                                      1 if int16 else 0, util._ptr(ns), util._ptr(fr))
        assert r == 0
        return pcm, ns, fr

    def close(self):
        self.L.emu_decoder_destroy(self.h)


def emu_decode_stream(pkts, sizes, channels, frame_size, Fs=48000):
    """One stream, many packets (like util.oracle_decode_float); frame_size is in samples at the output rate Fs."""
    d = EmuDecoder(1, channels, Fs)
    T = len(sizes)
    out = np.zeros((T * frame_size, channels), np.float32)
    rng = np.zeros(T, np.uint32)
    for t in range(T):
        pcm, ns, fr = d.decode(pkts[t:t + 1], sizes[t:t + 1], frame_size=frame_size)
        assert ns[0] == frame_size, (t, ns[0])
        out[t * frame_size:(t + 1) * frame_size] = pcm[0, :frame_size * channels].reshape(frame_size, channels)
        rng[t] = fr[0]
    d.close()
    return out, rng
