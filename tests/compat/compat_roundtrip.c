/* tests/compat/compat_roundtrip.c — a libopus client program, written against the public API only (it also compiles against the
 * reference's own <opus.h>): encodes synthetic stereo audio with opus_encode(), decodes the packets with opus_decode(), checks that
 * encoder and decoder agree on OPUS_GET_FINAL_RANGE for every packet (what the reference's tests/test_opus_encode.c:501 asserts) and
 * prints a digest of the packets so that the caller can compare with the reference's output.  Linked against libopus_b200_compat.so. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#ifdef USE_REFERENCE_HEADERS
#include <opus.h>
#else
#include "opus_b200_compat.h"
#endif

int main(int argc, char **argv) {
  const int Fs = 48000, ch = 2, fs = 960, nframes = argc > 1 ? atoi(argv[1]) : 12;
  int err = 0, f, i;
  OpusEncoder *enc = opus_encoder_create(Fs, ch, OPUS_APPLICATION_RESTRICTED_CELT, &err);
  if (!enc) { printf("encoder_create failed: %s\n", opus_strerror(err)); return 2; }
  OpusDecoder *dec = opus_decoder_create(Fs, ch, &err);
  if (!dec) { printf("decoder_create failed: %s\n", opus_strerror(err)); return 2; }
  opus_encoder_ctl(enc, OPUS_SET_BITRATE(64000));
  opus_encoder_ctl(enc, OPUS_SET_VBR(1));
  opus_encoder_ctl(enc, OPUS_SET_COMPLEXITY(10));
  opus_int16 *pcm = (opus_int16 *)malloc(sizeof(opus_int16) * fs * ch), *out = (opus_int16 *)malloc(sizeof(opus_int16) * fs * ch);
  unsigned char pkt[1276];
  unsigned long digest = 1469598103UL;
  double phase = 0;
  for (f = 0; f < nframes; f++) {
    for (i = 0; i < fs; i++) {
      const double t = (f * fs + i) / 48000.0;
      phase += 2 * 3.141592653589793 * (440 + 200 * sin(2 * 3.141592653589793 * 1.5 * t)) / 48000.0;
      pcm[2 * i] = (opus_int16)(9000 * sin(phase) + 2000 * sin(2 * 3.141592653589793 * 3000 * t));
      pcm[2 * i + 1] = (opus_int16)(7000 * sin(phase + 0.7) + ((i * 7919 + f * 104729) % 601 - 300));
    }
    const int nb = opus_encode(enc, pcm, fs, pkt, sizeof(pkt));
    if (nb < 0) { printf("opus_encode failed at frame %d: %s\n", f, opus_strerror(nb)); return 3; }
    opus_uint32 er = 0, dr = 0;
    opus_encoder_ctl(enc, OPUS_GET_FINAL_RANGE(&er));
    const int ns = opus_decode(dec, pkt, nb, out, fs, 0);
    if (ns != fs) { printf("opus_decode returned %d at frame %d\n", ns, f); return 4; }
    opus_decoder_ctl(dec, OPUS_GET_FINAL_RANGE(&dr));
    if (er != dr) { printf("final range mismatch at frame %d: enc %08x dec %08x\n", f, er, dr); return 5; }
    for (i = 0; i < nb; i++) digest = (digest ^ pkt[i]) * 1099511628211UL;
    printf("frame %d bytes %d toc %02x range %08x nb_samples %d\n", f, nb, pkt[0], er, opus_packet_get_nb_samples(pkt, nb, Fs));
  }
  opus_int32 dur = 0;
  opus_decoder_ctl(dec, OPUS_GET_LAST_PACKET_DURATION(&dur));
  printf("digest %016lx last_duration %d\n", digest, dur);
  /* a lost packet is concealed */
  if (opus_decode(dec, NULL, 0, out, fs, 0) != fs) { printf("PLC failed\n"); return 6; }
  opus_encoder_destroy(enc);
  opus_decoder_destroy(dec);
  free(pcm);
  free(out);
  printf("compat roundtrip ok\n");
  return 0;
}
