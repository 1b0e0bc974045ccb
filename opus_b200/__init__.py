"""opus_b200 — B200-native batched Opus (CELT) hot path.

Host-side mirror of the reference's per-stream decoder API (include/opus.h:458-591), batched:
one `DecoderBatch` = N independent Opus decoders resident on one GPU.  All compute happens in
the sm_100a kernels of libopus_b200.so (opus_b200/csrc/), reached through the C ABI declared in
include/opus_b200.h.  There is no CPU implementation in this package: importing works without a
GPU, creating a decoder does not.
"""
from .capi import (OpusB200Error, lib, library_path, build_library, OPUS_OK, OPUS_BAD_ARG, OPUS_BUFFER_TOO_SMALL,
                   OPUS_INTERNAL_ERROR, OPUS_INVALID_PACKET, OPUS_UNIMPLEMENTED, OPUS_INVALID_STATE, OPUS_ALLOC_FAIL)
from .decoder import DecoderBatch
from .ops import Ops
from .encoder import EncoderBatch, OpusEncoderBatch

__all__ = ['DecoderBatch', 'EncoderBatch', 'OpusEncoderBatch', 'Ops', 'OpusB200Error', 'lib', 'library_path', 'build_library']
