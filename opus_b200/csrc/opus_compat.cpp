// opus_compat.cpp — libopus_b200_compat.so: the reference's public per-stream C API (include/opus.h, include/opus_multistream.h of
// xiph/opus) as a thin shim over the batched C ABI of libopus_b200.so, so that an application written against libopus links against
// this library unchanged for the CELT-only path.  Every entry point below has the reference's name, signature and
// return conventions; the object behind an OpusDecoder / OpusEncoder is a batch of ONE stream on the GPU, the object behind an
// OpusMSDecoder is one batch of `coupled` stereo streams plus one of `streams - coupled` mono streams (one multistream packet = one
// batched launch per kind instead of the reference's serial loop, src/opus_multistream_decoder.c:238-253).  No codec arithmetic lives
// here: packet framing helpers (src/opus.c:168-400, restated) and the calls into the C ABI.  There is no CPU fallback: without a CUDA
// device the create functions fail with OPUS_INTERNAL_ERROR.
//
// Not provided (the symbols are absent, an application that needs them does not link): SILK / hybrid operation (encoder applications
// VOIP / AUDIO / RESTRICTED_SILK return OPUS_UNIMPLEMENTED from create, SILK / hybrid packets OPUS_UNIMPLEMENTED from decode), FEC,
// DRED, the 24-bit entry points, the repacketizer, the multistream ENCODER, projection.
#include <stdarg.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <new>
#include <vector>

#include "../../include/opus_b200.h"
#include "../../include/opus_b200_compat.h"

namespace {
const int kMaxFrame48 = 5760;   // 120 ms at 48 kHz
int g_device = -1;
int device() {
  if (g_device < 0) {
    const char *e = getenv("OPUSB200_DEVICE");
    g_device = e ? atoi(e) : 0;
  }
  return g_device;
}
bool valid_fs(opus_int32 Fs) { return Fs == 48000 || Fs == 24000 || Fs == 16000 || Fs == 12000 || Fs == 8000; }
}  // namespace

struct OpusDecoder {
  uint32_t magic;
  OpusB200Decoder *d;
  opus_int32 Fs;
  int channels;
  std::vector<unsigned char> *pkt;   // staging: the batched call takes a stride
};
struct OpusEncoder {
  uint32_t magic;
  OpusB200OpusEncoder *e;
  opus_int32 Fs;
  int channels, application;
  std::vector<unsigned char> *pkt;
};
struct OpusMSDecoder {
  uint32_t magic;
  OpusB200Decoder *dc, *dm;   // coupled (stereo) streams, uncoupled (mono) streams
  opus_int32 Fs;
  int nb_channels, nb_streams, nb_coupled;
  unsigned char mapping[256];
  std::vector<unsigned char> *pk;
  std::vector<float> *buf;
};
static const uint32_t kDecMagic = 0x0b200dec, kEncMagic = 0x0b200e2c, kMsMagic = 0x0b20035d;

extern "C" {

// ---- src/celt/celt.c opus_strerror / opus_get_version_string ----
const char *opus_strerror(int error) {
  static const char *const s[8] = {"success", "invalid argument", "buffer too small", "internal error", "corrupted stream", "request not implemented",
                                   "invalid state", "memory allocation failed"};
  return (error > 0 || error < -7) ? "unknown error" : s[-error];
}
const char *opus_get_version_string(void) { return "libopus_b200_compat (B200 batched CELT path behind the libopus API)"; }

// ---- packet helpers (src/opus.c:168-400, src/opus_decoder.c:1217-1290) ----
static int parse_size(const unsigned char *data, opus_int32 len, opus_int16 *size) {
  if (len < 1) { *size = -1; return -1; }
  if (data[0] < 252) { *size = data[0]; return 1; }
  if (len < 2) { *size = -1; return -1; }
  *size = (opus_int16)(4 * data[1] + data[0]);
  return 2;
}
int opus_packet_get_samples_per_frame(const unsigned char *data, opus_int32 Fs) {
  int audiosize;
  if (data[0] & 0x80) {
    audiosize = ((data[0] >> 3) & 0x3);
    audiosize = (Fs << audiosize) / 400;
  } else if ((data[0] & 0x60) == 0x60) {
    audiosize = (data[0] & 0x08) ? Fs / 50 : Fs / 100;
  } else {
    audiosize = ((data[0] >> 3) & 0x3);
    if (audiosize == 3) audiosize = Fs * 60 / 1000;
    else audiosize = (Fs << audiosize) / 100;
  }
  return audiosize;
}
// opus_packet_parse_impl (src/opus.c:222-390)
static int packet_parse_impl(const unsigned char *data, opus_int32 len, int self_delimited, unsigned char *out_toc, const unsigned char *frames[48],
                             opus_int16 size[48], int *payload_offset, opus_int32 *packet_offset) {
  int i, bytes, count, cbr = 0;
  unsigned char ch, toc;
  opus_int32 last_size, pad = 0;
  const unsigned char *data0 = data;
  if (size == NULL || len < 0) return OPUS_BAD_ARG;
  if (len == 0) return OPUS_INVALID_PACKET;
  const int framesize = opus_packet_get_samples_per_frame(data, 48000);
  toc = *data++;
  len--;
  last_size = len;
  switch (toc & 0x3) {
    case 0: count = 1; break;
    case 1:
      count = 2;
      cbr = 1;
      if (!self_delimited) {
        if (len & 0x1) return OPUS_INVALID_PACKET;
        last_size = len / 2;
        size[0] = (opus_int16)last_size;
      }
      break;
    case 2:
      count = 2;
      bytes = parse_size(data, len, size);
      len -= bytes;
      if (size[0] < 0 || size[0] > len) return OPUS_INVALID_PACKET;
      data += bytes;
      last_size = len - size[0];
      break;
    default:
      if (len < 1) return OPUS_INVALID_PACKET;
      ch = *data++;
      count = ch & 0x3F;
      if (count <= 0 || framesize * (opus_int32)count > 5760) return OPUS_INVALID_PACKET;
      len--;
      if (ch & 0x40) {
        int p;
        do {
          if (len <= 0) return OPUS_INVALID_PACKET;
          p = *data++;
          len--;
          const int tmp = p == 255 ? 254 : p;
          len -= tmp;
          pad += tmp;
        } while (p == 255);
      }
      if (len < 0) return OPUS_INVALID_PACKET;
      cbr = !(ch & 0x80);
      if (!cbr) {
        last_size = len;
        for (i = 0; i < count - 1; i++) {
          bytes = parse_size(data, len, size + i);
          len -= bytes;
          if (size[i] < 0 || size[i] > len) return OPUS_INVALID_PACKET;
          data += bytes;
          last_size -= bytes + size[i];
        }
        if (last_size < 0) return OPUS_INVALID_PACKET;
      } else if (!self_delimited) {
        last_size = len / count;
        if (last_size * count != len) return OPUS_INVALID_PACKET;
        for (i = 0; i < count - 1; i++) size[i] = (opus_int16)last_size;
      }
      break;
  }
  if (self_delimited) {
    bytes = parse_size(data, len, size + count - 1);
    len -= bytes;
    if (size[count - 1] < 0 || size[count - 1] > len) return OPUS_INVALID_PACKET;
    data += bytes;
    if (cbr) {
      if (size[count - 1] * count > len) return OPUS_INVALID_PACKET;
      for (i = 0; i < count - 1; i++) size[i] = size[count - 1];
    } else if (bytes + size[count - 1] > last_size) return OPUS_INVALID_PACKET;
  } else {
    if (last_size > 1275) return OPUS_INVALID_PACKET;
    size[count - 1] = (opus_int16)last_size;
  }
  if (payload_offset) *payload_offset = (int)(data - data0);
  for (i = 0; i < count; i++) {
    if (frames) frames[i] = data;
    data += size[i];
  }
  if (packet_offset) *packet_offset = pad + (opus_int32)(data - data0);
  if (out_toc) *out_toc = toc;
  return count;
}
int opus_packet_parse(const unsigned char *data, opus_int32 len, unsigned char *out_toc, const unsigned char *frames[48], opus_int16 size[48],
                      int *payload_offset) {
  return packet_parse_impl(data, len, 0, out_toc, frames, size, payload_offset, NULL);
}
int opus_packet_get_bandwidth(const unsigned char *data) {
  int bandwidth;
  if (data[0] & 0x80) {
    bandwidth = OPUS_BANDWIDTH_MEDIUMBAND + ((data[0] >> 5) & 0x3);
    if (bandwidth == OPUS_BANDWIDTH_MEDIUMBAND) bandwidth = OPUS_BANDWIDTH_NARROWBAND;
  } else if ((data[0] & 0x60) == 0x60) {
    bandwidth = (data[0] & 0x10) ? OPUS_BANDWIDTH_FULLBAND : OPUS_BANDWIDTH_SUPERWIDEBAND;
  } else {
    bandwidth = OPUS_BANDWIDTH_NARROWBAND + ((data[0] >> 5) & 0x3);
  }
  return bandwidth;
}
int opus_packet_get_nb_channels(const unsigned char *data) { return (data[0] & 0x4) ? 2 : 1; }
int opus_packet_get_nb_frames(const unsigned char packet[], opus_int32 len) {
  if (len < 1) return OPUS_BAD_ARG;
  const int count = packet[0] & 0x3;
  if (count == 0) return 1;
  if (count != 3) return 2;
  if (len < 2) return OPUS_INVALID_PACKET;
  return packet[1] & 0x3F;
}
int opus_packet_get_nb_samples(const unsigned char packet[], opus_int32 len, opus_int32 Fs) {
  const int count = opus_packet_get_nb_frames(packet, len);
  if (count < 0) return count;
  const int samples = count * opus_packet_get_samples_per_frame(packet, Fs);
  if (samples * 25 > Fs * 3) return OPUS_INVALID_PACKET;
  return samples;
}
int opus_decoder_get_nb_samples(const OpusDecoder *dec, const unsigned char packet[], opus_int32 len) {
  return opus_packet_get_nb_samples(packet, len, dec->Fs);
}

// ---- decoder (include/opus.h:458-591, src/opus_decoder.c) ----
int opus_decoder_get_size(int channels) { return (channels < 1 || channels > 2) ? 0 : (int)sizeof(OpusDecoder); }
int opus_decoder_init(OpusDecoder *st, opus_int32 Fs, int channels) {
  if (!st || !valid_fs(Fs) || (channels != 1 && channels != 2)) return OPUS_BAD_ARG;
  memset(st, 0, sizeof(*st));
  const int r = opusb200_decoder_create(device(), Fs, channels, 1, &st->d);
  if (r != OPUSB200_OK) return r;
  st->magic = kDecMagic; st->Fs = Fs; st->channels = channels;
  st->pkt = new (std::nothrow) std::vector<unsigned char>();
  return st->pkt ? OPUS_OK : OPUS_ALLOC_FAIL;
}
OpusDecoder *opus_decoder_create(opus_int32 Fs, int channels, int *error) {
  OpusDecoder *st = (OpusDecoder *)malloc(sizeof(OpusDecoder));
  int r = st ? opus_decoder_init(st, Fs, channels) : OPUS_ALLOC_FAIL;
  if (r != OPUS_OK) { free(st); st = NULL; }
  if (error) *error = r;
  return st;
}
void opus_decoder_destroy(OpusDecoder *st) {
  if (!st || st->magic != kDecMagic) return;
  opusb200_decoder_destroy(st->d);
  delete st->pkt;
  st->magic = 0;
  free(st);
}
static int decode_common(OpusDecoder *st, const unsigned char *data, opus_int32 len, void *pcm, int is_float, int frame_size, int decode_fec) {
  if (!st || st->magic != kDecMagic || !pcm || frame_size <= 0 || len < 0) return OPUS_BAD_ARG;
  if (decode_fec) return OPUS_UNIMPLEMENTED;   // in-band FEC lives in SILK frames
  if (len > 0 && !data) return OPUS_BAD_ARG;
  // opus_decode_native limits the frame size to what a packet can hold (src/opus_decoder.c:905-908, 934)
  if (frame_size > st->Fs / 25 * 3) frame_size = st->Fs / 25 * 3;
  if (data == NULL || len == 0) {   // packet loss: frame_size must be a multiple of 2.5 ms (:757-778)
    if (frame_size % (st->Fs / 400) != 0) return OPUS_BAD_ARG;
    len = 0;
  }
  int32_t l = len, ns = 0;
  const unsigned char dummy = 0;
  const int stride = len > 0 ? len : 1;
  int r;
  if (is_float) r = opusb200_decode_float(st->d, len > 0 ? data : &dummy, stride, &l, (float *)pcm, frame_size * st->channels, frame_size, &ns);
  else r = opusb200_decode(st->d, len > 0 ? data : &dummy, stride, &l, (int16_t *)pcm, frame_size * st->channels, frame_size, &ns);
  return r != OPUSB200_OK ? r : ns;
}
int opus_decode(OpusDecoder *st, const unsigned char *data, opus_int32 len, opus_int16 *pcm, int frame_size, int decode_fec) {
  return decode_common(st, data, len, pcm, 0, frame_size, decode_fec);
}
int opus_decode_float(OpusDecoder *st, const unsigned char *data, opus_int32 len, float *pcm, int frame_size, int decode_fec) {
  return decode_common(st, data, len, pcm, 1, frame_size, decode_fec);
}
int opus_decoder_ctl(OpusDecoder *st, int request, ...) {
  if (!st || st->magic != kDecMagic) return OPUS_BAD_ARG;
  va_list ap;
  va_start(ap, request);
  int ret = OPUS_OK;
  switch (request) {
    case OPUS_RESET_STATE: ret = opusb200_decoder_reset(st->d); break;
    case OPUS_GET_FINAL_RANGE_REQUEST: {
      opus_uint32 *v = va_arg(ap, opus_uint32 *);
      if (!v) { ret = OPUS_BAD_ARG; break; }
      uint32_t r = 0;
      ret = opusb200_decoder_get_final_range(st->d, &r);
      *v = r;
    } break;
    case OPUS_SET_GAIN_REQUEST:
    case OPUS_SET_PHASE_INVERSION_DISABLED_REQUEST: ret = opusb200_decoder_ctl(st->d, request, va_arg(ap, opus_int32)); break;
    case OPUS_GET_GAIN_REQUEST:
    case OPUS_GET_PITCH_REQUEST:
    case OPUS_GET_LAST_PACKET_DURATION_REQUEST:
    case OPUS_GET_PHASE_INVERSION_DISABLED_REQUEST:
    case OPUS_GET_SAMPLE_RATE_REQUEST: {
      opus_int32 *v = va_arg(ap, opus_int32 *);
      if (!v) { ret = OPUS_BAD_ARG; break; }
      int32_t r = 0;
      ret = opusb200_decoder_get_ctl(st->d, request, &r);
      *v = r;
    } break;
    default: ret = OPUS_UNIMPLEMENTED;
  }
  va_end(ap);
  return ret;
}

// ---- encoder (include/opus.h:170-370, src/opus_encoder.c) ----
int opus_encoder_get_size(int channels) { return (channels < 1 || channels > 2) ? 0 : (int)sizeof(OpusEncoder); }
int opus_encoder_init(OpusEncoder *st, opus_int32 Fs, int channels, int application) {
  if (!st || !valid_fs(Fs) || (channels != 1 && channels != 2)) return OPUS_BAD_ARG;
  memset(st, 0, sizeof(*st));
  const int r = opusb200_opus_encoder_create(device(), Fs, channels, application, 1, &st->e);
  if (r != OPUSB200_OK) return r;
  st->magic = kEncMagic; st->Fs = Fs; st->channels = channels; st->application = application;
  st->pkt = new (std::nothrow) std::vector<unsigned char>();
  return st->pkt ? OPUS_OK : OPUS_ALLOC_FAIL;
}
OpusEncoder *opus_encoder_create(opus_int32 Fs, int channels, int application, int *error) {
  OpusEncoder *st = (OpusEncoder *)malloc(sizeof(OpusEncoder));
  int r = st ? opus_encoder_init(st, Fs, channels, application) : OPUS_ALLOC_FAIL;
  if (r != OPUS_OK) { free(st); st = NULL; }
  if (error) *error = r;
  return st;
}
void opus_encoder_destroy(OpusEncoder *st) {
  if (!st || st->magic != kEncMagic) return;
  opusb200_opus_encoder_destroy(st->e);
  delete st->pkt;
  st->magic = 0;
  free(st);
}
static opus_int32 encode_common(OpusEncoder *st, const void *pcm, int is16, int frame_size, unsigned char *data, opus_int32 max_data_bytes) {
  if (!st || st->magic != kEncMagic || !pcm || !data || max_data_bytes <= 0) return OPUS_BAD_ARG;
  const int stride = max_data_bytes > 1280 * 6 ? 1280 * 6 : max_data_bytes;   // room for six 1276-byte frames and their code-3 sizes
  int32_t nb = 0;
  const int r = is16 ? opusb200_opus_encode(st->e, (const int16_t *)pcm, frame_size, data, stride, max_data_bytes, &nb)
                     : opusb200_opus_encode_float(st->e, (const float *)pcm, frame_size, data, stride, max_data_bytes, &nb);
  return r != OPUSB200_OK ? r : nb;
}
opus_int32 opus_encode(OpusEncoder *st, const opus_int16 *pcm, int frame_size, unsigned char *data, opus_int32 max_data_bytes) {
  return encode_common(st, pcm, 1, frame_size, data, max_data_bytes);
}
opus_int32 opus_encode_float(OpusEncoder *st, const float *pcm, int frame_size, unsigned char *data, opus_int32 max_data_bytes) {
  return encode_common(st, pcm, 0, frame_size, data, max_data_bytes);
}
int opus_encoder_ctl(OpusEncoder *st, int request, ...) {
  if (!st || st->magic != kEncMagic) return OPUS_BAD_ARG;
  va_list ap;
  va_start(ap, request);
  int ret;
  switch (request) {
    case OPUS_RESET_STATE: ret = opusb200_opus_encoder_ctl(st->e, request, 0); break;
    case OPUS_GET_FINAL_RANGE_REQUEST: {
      opus_uint32 *v = va_arg(ap, opus_uint32 *);
      if (!v) { ret = OPUS_BAD_ARG; break; }
      uint32_t r = 0;
      ret = opusb200_opus_encoder_get_final_range(st->e, &r);
      *v = r;
    } break;
    case OPUS_GET_SAMPLE_RATE_REQUEST: {
      opus_int32 *v = va_arg(ap, opus_int32 *);
      if (!v) { ret = OPUS_BAD_ARG; break; }
      *v = st->Fs;
      ret = OPUS_OK;
    } break;
    case OPUS_GET_LOOKAHEAD_REQUEST: {   // src/opus_encoder.c:3086-3093: Fs/400 for the delay-restricted applications
      opus_int32 *v = va_arg(ap, opus_int32 *);
      if (!v) { ret = OPUS_BAD_ARG; break; }
      *v = st->Fs / 400;
      ret = OPUS_OK;
    } break;
    default: ret = (request & 1) ? OPUS_UNIMPLEMENTED /* other getters */ : opusb200_opus_encoder_ctl(st->e, request, va_arg(ap, opus_int32));
  }
  va_end(ap);
  return ret;
}

// ---- multistream decoder (include/opus_multistream.h:508-726, src/opus_multistream_decoder.c) ----
opus_int32 opus_multistream_decoder_get_size(int nb_streams, int nb_coupled_streams) {
  if (nb_streams < 1 || nb_coupled_streams > nb_streams || nb_coupled_streams < 0) return 0;
  return (opus_int32)sizeof(OpusMSDecoder);
}
int opus_multistream_decoder_init(OpusMSDecoder *st, opus_int32 Fs, int channels, int streams, int coupled_streams, const unsigned char *mapping) {
  if (!st || !mapping || !valid_fs(Fs)) return OPUS_BAD_ARG;
  if (channels > 255 || channels < 1 || coupled_streams > streams || streams < 1 || coupled_streams < 0 || streams > 255 - coupled_streams) return OPUS_BAD_ARG;
  memset(st, 0, sizeof(*st));
  for (int i = 0; i < channels; i++) {   // validate_layout (src/opus_multistream.c:41-54)
    st->mapping[i] = mapping[i];
    if (mapping[i] >= streams + coupled_streams && mapping[i] != 255) return OPUS_BAD_ARG;
  }
  st->Fs = Fs; st->nb_channels = channels; st->nb_streams = streams; st->nb_coupled = coupled_streams;
  int r = OPUSB200_OK;
  if (coupled_streams > 0) r = opusb200_decoder_create(device(), Fs, 2, coupled_streams, &st->dc);
  if (r == OPUSB200_OK && streams > coupled_streams) r = opusb200_decoder_create(device(), Fs, 1, streams - coupled_streams, &st->dm);
  if (r != OPUSB200_OK) { opusb200_decoder_destroy(st->dc); opusb200_decoder_destroy(st->dm); return r; }
  st->pk = new (std::nothrow) std::vector<unsigned char>();
  st->buf = new (std::nothrow) std::vector<float>();
  if (!st->pk || !st->buf) return OPUS_ALLOC_FAIL;
  st->magic = kMsMagic;
  return OPUS_OK;
}
OpusMSDecoder *opus_multistream_decoder_create(opus_int32 Fs, int channels, int streams, int coupled_streams, const unsigned char *mapping, int *error) {
  OpusMSDecoder *st = (OpusMSDecoder *)malloc(sizeof(OpusMSDecoder));
  int r = st ? opus_multistream_decoder_init(st, Fs, channels, streams, coupled_streams, mapping) : OPUS_ALLOC_FAIL;
  if (r != OPUS_OK) { free(st); st = NULL; }
  if (error) *error = r;
  return st;
}
void opus_multistream_decoder_destroy(OpusMSDecoder *st) {
  if (!st || st->magic != kMsMagic) return;
  opusb200_decoder_destroy(st->dc);
  opusb200_decoder_destroy(st->dm);
  delete st->pk;
  delete st->buf;
  st->magic = 0;
  free(st);
}
// one stream's frames as a standard (not self-delimited) packet: code 0 for one frame, else code 3 VBR, which can carry any frame sizes
static int repack_standard(unsigned char toc, const unsigned char *frames[48], const opus_int16 size[48], int count, unsigned char *out) {
  int n = 0;
  if (count == 1) {
    out[n++] = toc & 0xFC;
    memcpy(out + n, frames[0], size[0]);
    return n + size[0];
  }
  out[n++] = (toc & 0xFC) | 3;
  out[n++] = (unsigned char)(0x80 | count);
  for (int i = 0; i < count - 1; i++) {
    if (size[i] < 252) out[n++] = (unsigned char)size[i];
    else { out[n] = (unsigned char)(252 + (size[i] & 3)); out[n + 1] = (unsigned char)((size[i] - out[n]) >> 2); n += 2; }
  }
  for (int i = 0; i < count; i++) { memcpy(out + n, frames[i], size[i]); n += size[i]; }
  return n;
}
static int ms_decode_common(OpusMSDecoder *st, const unsigned char *data, opus_int32 len, void *pcm, int is_float, int frame_size, int decode_fec) {
  if (!st || st->magic != kMsMagic || !pcm || frame_size <= 0 || len < 0) return OPUS_BAD_ARG;
  if (decode_fec) return OPUS_UNIMPLEMENTED;
  const opus_int32 Fs = st->Fs;
  if (frame_size > Fs / 25 * 3) frame_size = Fs / 25 * 3;
  const int S = st->nb_streams, SC = st->nb_coupled, SM = S - SC, do_plc = len == 0 || data == NULL;
  if (!do_plc && len < 2 * S - 1) return OPUS_INVALID_PACKET;
  const int stride = 2 * 48 + 48 * 1275 / 8 + 1500;   // a stream packet of up to 120 ms; CELT frames of one packet sum to well below this
  st->pk->assign((size_t)S * stride, 0);
  std::vector<int32_t> lens(S, 0), ns(S, 0);
  if (!do_plc) {   // opus_multistream_packet_validate + the per-stream split (src/opus_multistream_decoder.c:147-176, 238-257)
    const unsigned char *p = data;
    opus_int32 left = len;
    int samples = 0;
    for (int s = 0; s < S; s++) {
      if (left <= 0) return OPUS_INVALID_PACKET;
      const unsigned char *frames[48];
      opus_int16 size[48];
      unsigned char toc;
      opus_int32 packet_offset = 0;
      const int count = packet_parse_impl(p, left, s != S - 1, &toc, frames, size, NULL, &packet_offset);
      if (count < 0) return count;
      const int tmp = count * opus_packet_get_samples_per_frame(p, Fs);
      if (tmp * 25 > Fs * 3) return OPUS_INVALID_PACKET;
      if (s != 0 && samples != tmp) return OPUS_INVALID_PACKET;
      samples = tmp;
      int tot = 2 + 2 * count;
      for (int i = 0; i < count; i++) tot += size[i];
      if (tot > stride) return OPUS_INVALID_PACKET;
      lens[s] = repack_standard(toc, frames, size, count, st->pk->data() + (size_t)s * stride);
      p += packet_offset;
      left -= packet_offset;
    }
    if (samples > frame_size) return OPUS_BUFFER_TOO_SMALL;
  } else if (frame_size % (Fs / 400) != 0) return OPUS_BAD_ARG;
  // two batched calls: the coupled streams as one stereo batch, the others as one mono batch
  const size_t esz = is_float ? sizeof(float) : sizeof(int16_t);
  st->buf->assign(((size_t)SC * 2 + SM) * frame_size * esz / sizeof(float) + 8, 0.f);
  char *bc = (char *)st->buf->data(), *bm = bc + (size_t)SC * 2 * frame_size * esz;
  int r = OPUSB200_OK;
  if (SC > 0) r = is_float ? opusb200_decode_float(st->dc, st->pk->data(), stride, lens.data(), (float *)bc, 2 * frame_size, frame_size, ns.data())
                           : opusb200_decode(st->dc, st->pk->data(), stride, lens.data(), (int16_t *)bc, 2 * frame_size, frame_size, ns.data());
  if (r == OPUSB200_OK && SM > 0)
    r = is_float ? opusb200_decode_float(st->dm, st->pk->data() + (size_t)SC * stride, stride, lens.data() + SC, (float *)bm, frame_size, frame_size, ns.data() + SC)
                 : opusb200_decode(st->dm, st->pk->data() + (size_t)SC * stride, stride, lens.data() + SC, (int16_t *)bm, frame_size, frame_size, ns.data() + SC);
  if (r != OPUSB200_OK) return r;
  for (int s = 0; s < S; s++) if (ns[s] <= 0) return ns[s];
  const int n_out = ns[S - 1];
  // channel mapping (src/opus_multistream_decoder.c:265-305, get_left/right/mono_channel src/opus_multistream.c:57-91)
  for (int c = 0; c < st->nb_channels; c++) {
    const int m = st->mapping[c];
    const char *src = NULL;
    int sstride = 0;
    if (m != 255) {
      if (m < 2 * SC) { src = bc + ((size_t)(m >> 1) * 2 * frame_size + (m & 1)) * esz; sstride = 2; }
      else { src = bm + (size_t)(m - 2 * SC) * frame_size * esz; sstride = 1; }
    }
    for (int i = 0; i < n_out; i++) {
      if (is_float) ((float *)pcm)[(size_t)i * st->nb_channels + c] = src ? ((const float *)src)[(size_t)i * sstride] : 0.f;
      else ((int16_t *)pcm)[(size_t)i * st->nb_channels + c] = src ? ((const int16_t *)src)[(size_t)i * sstride] : 0;
    }
  }
  return n_out;
}
int opus_multistream_decode(OpusMSDecoder *st, const unsigned char *data, opus_int32 len, opus_int16 *pcm, int frame_size, int decode_fec) {
  return ms_decode_common(st, data, len, pcm, 0, frame_size, decode_fec);
}
int opus_multistream_decode_float(OpusMSDecoder *st, const unsigned char *data, opus_int32 len, float *pcm, int frame_size, int decode_fec) {
  return ms_decode_common(st, data, len, pcm, 1, frame_size, decode_fec);
}
int opus_multistream_decoder_ctl(OpusMSDecoder *st, int request, ...) {
  if (!st || st->magic != kMsMagic) return OPUS_BAD_ARG;
  va_list ap;
  va_start(ap, request);
  int ret = OPUS_OK;
  switch (request) {
    case OPUS_RESET_STATE:
      if (st->dc) ret = opusb200_decoder_reset(st->dc);
      if (ret == OPUS_OK && st->dm) ret = opusb200_decoder_reset(st->dm);
      break;
    case OPUS_GET_SAMPLE_RATE_REQUEST: { opus_int32 *v = va_arg(ap, opus_int32 *); if (v) *v = st->Fs; else ret = OPUS_BAD_ARG; } break;
    case OPUS_GET_FINAL_RANGE_REQUEST: {   // the XOR over the streams (src/opus_multistream_decoder.c:455-475)
      opus_uint32 *v = va_arg(ap, opus_uint32 *);
      if (!v) { ret = OPUS_BAD_ARG; break; }
      std::vector<uint32_t> r(st->nb_streams, 0);
      if (st->dc) ret = opusb200_decoder_get_final_range(st->dc, r.data());
      if (ret == OPUS_OK && st->dm) ret = opusb200_decoder_get_final_range(st->dm, r.data() + st->nb_coupled);
      uint32_t x = 0;
      for (int s = 0; s < st->nb_streams; s++) x ^= r[s];
      *v = x;
    } break;
    case OPUS_SET_GAIN_REQUEST:
    case OPUS_SET_PHASE_INVERSION_DISABLED_REQUEST: {
      const opus_int32 val = va_arg(ap, opus_int32);
      if (st->dc) ret = opusb200_decoder_ctl(st->dc, request, val);
      if (ret == OPUS_OK && st->dm) ret = opusb200_decoder_ctl(st->dm, request, val);
    } break;
    default: ret = OPUS_UNIMPLEMENTED;
  }
  va_end(ap);
  return ret;
}

}  // extern "C"
