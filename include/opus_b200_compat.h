/* opus_b200_compat.h — the subset of the reference's public API (xiph/opus include/opus.h, include/opus_multistream.h,
 * include/opus_defines.h, include/opus_types.h) that libopus_b200_compat.so exports with the reference's names, signatures, request
 * numbers and return codes, so that a program written against libopus compiles against this header (or against the reference's own
 * opus.h) and links against libopus_b200_compat.so unchanged.  Each object is a batch of one stream on the GPU (a multistream decoder:
 * one batch per stream kind); codec behaviour and limits are those of include/opus_b200.h: CELT-only packets on decode, the
 * RESTRICTED_CELT / RESTRICTED_LOWDELAY applications on encode.  SURVEY 8(b) B1.
 *
 *   reference declaration (include/...)                      here
 *   opus.h:460-591   opus_decoder_get_size/create/init/destroy, opus_decode, opus_decode_float, opus_decoder_ctl        same
 *   opus.h:174-367   opus_encoder_get_size/create/init/destroy, opus_encode, opus_encode_float, opus_encoder_ctl        same
 *   opus.h:713-788   opus_packet_parse / _get_bandwidth / _get_samples_per_frame / _get_nb_channels / _get_nb_frames /
 *                    _get_nb_samples, opus_decoder_get_nb_samples                                                       same
 *   opus_multistream.h:508-726  opus_multistream_decoder_get_size/create/init/destroy, opus_multistream_decode[_float], _ctl   same
 *   opus_defines.h:850,860      opus_strerror, opus_get_version_string                                                  same
 * Not exported: opus_decode24 / opus_encode24, FEC / DRED entry points, the repacketizer, the multistream encoder, projection. */
#ifndef OPUS_B200_COMPAT_H
#define OPUS_B200_COMPAT_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int32_t opus_int32;
typedef uint32_t opus_uint32;
typedef int16_t opus_int16;
typedef uint16_t opus_uint16;

/* error codes (include/opus_defines.h:46-60) */
#define OPUS_OK 0
#define OPUS_BAD_ARG (-1)
#define OPUS_BUFFER_TOO_SMALL (-2)
#define OPUS_INTERNAL_ERROR (-3)
#define OPUS_INVALID_PACKET (-4)
#define OPUS_UNIMPLEMENTED (-5)
#define OPUS_INVALID_STATE (-6)
#define OPUS_ALLOC_FAIL (-7)

/* values (include/opus_defines.h:199-241) */
#define OPUS_AUTO (-1000)
#define OPUS_BITRATE_MAX (-1)
#define OPUS_APPLICATION_VOIP 2048
#define OPUS_APPLICATION_AUDIO 2049
#define OPUS_APPLICATION_RESTRICTED_LOWDELAY 2051
#define OPUS_APPLICATION_RESTRICTED_SILK 2052
#define OPUS_APPLICATION_RESTRICTED_CELT 2053
#define OPUS_SIGNAL_VOICE 3001
#define OPUS_SIGNAL_MUSIC 3002
#define OPUS_BANDWIDTH_NARROWBAND 1101
#define OPUS_BANDWIDTH_MEDIUMBAND 1102
#define OPUS_BANDWIDTH_WIDEBAND 1103
#define OPUS_BANDWIDTH_SUPERWIDEBAND 1104
#define OPUS_BANDWIDTH_FULLBAND 1105

/* ctl request numbers (include/opus_defines.h:131-185) */
#define OPUS_SET_BITRATE_REQUEST 4002
#define OPUS_SET_MAX_BANDWIDTH_REQUEST 4004
#define OPUS_SET_VBR_REQUEST 4006
#define OPUS_SET_BANDWIDTH_REQUEST 4008
#define OPUS_SET_COMPLEXITY_REQUEST 4010
#define OPUS_SET_PACKET_LOSS_PERC_REQUEST 4014
#define OPUS_SET_VBR_CONSTRAINT_REQUEST 4020
#define OPUS_SET_FORCE_CHANNELS_REQUEST 4022
#define OPUS_SET_SIGNAL_REQUEST 4024
#define OPUS_GET_LOOKAHEAD_REQUEST 4027
#define OPUS_RESET_STATE 4028
#define OPUS_GET_SAMPLE_RATE_REQUEST 4029
#define OPUS_GET_FINAL_RANGE_REQUEST 4031
#define OPUS_GET_PITCH_REQUEST 4033
#define OPUS_SET_GAIN_REQUEST 4034
#define OPUS_SET_LSB_DEPTH_REQUEST 4036
#define OPUS_GET_LAST_PACKET_DURATION_REQUEST 4039
#define OPUS_SET_PREDICTION_DISABLED_REQUEST 4042
#define OPUS_GET_GAIN_REQUEST 4045
#define OPUS_SET_PHASE_INVERSION_DISABLED_REQUEST 4046
#define OPUS_GET_PHASE_INVERSION_DISABLED_REQUEST 4047

/* the ctl macros applications use (include/opus_defines.h:243-790): (request, value) / (request, pointer) */
#define OPUS_SET_BITRATE(x) OPUS_SET_BITRATE_REQUEST, (opus_int32)(x)
#define OPUS_SET_MAX_BANDWIDTH(x) OPUS_SET_MAX_BANDWIDTH_REQUEST, (opus_int32)(x)
#define OPUS_SET_VBR(x) OPUS_SET_VBR_REQUEST, (opus_int32)(x)
#define OPUS_SET_BANDWIDTH(x) OPUS_SET_BANDWIDTH_REQUEST, (opus_int32)(x)
#define OPUS_SET_COMPLEXITY(x) OPUS_SET_COMPLEXITY_REQUEST, (opus_int32)(x)
#define OPUS_SET_PACKET_LOSS_PERC(x) OPUS_SET_PACKET_LOSS_PERC_REQUEST, (opus_int32)(x)
#define OPUS_SET_VBR_CONSTRAINT(x) OPUS_SET_VBR_CONSTRAINT_REQUEST, (opus_int32)(x)
#define OPUS_SET_FORCE_CHANNELS(x) OPUS_SET_FORCE_CHANNELS_REQUEST, (opus_int32)(x)
#define OPUS_SET_SIGNAL(x) OPUS_SET_SIGNAL_REQUEST, (opus_int32)(x)
#define OPUS_SET_LSB_DEPTH(x) OPUS_SET_LSB_DEPTH_REQUEST, (opus_int32)(x)
#define OPUS_SET_PREDICTION_DISABLED(x) OPUS_SET_PREDICTION_DISABLED_REQUEST, (opus_int32)(x)
#define OPUS_SET_PHASE_INVERSION_DISABLED(x) OPUS_SET_PHASE_INVERSION_DISABLED_REQUEST, (opus_int32)(x)
#define OPUS_SET_GAIN(x) OPUS_SET_GAIN_REQUEST, (opus_int32)(x)
#define OPUS_GET_LOOKAHEAD(x) OPUS_GET_LOOKAHEAD_REQUEST, (opus_int32 *)(x)
#define OPUS_GET_SAMPLE_RATE(x) OPUS_GET_SAMPLE_RATE_REQUEST, (opus_int32 *)(x)
#define OPUS_GET_FINAL_RANGE(x) OPUS_GET_FINAL_RANGE_REQUEST, (opus_uint32 *)(x)
#define OPUS_GET_PITCH(x) OPUS_GET_PITCH_REQUEST, (opus_int32 *)(x)
#define OPUS_GET_GAIN(x) OPUS_GET_GAIN_REQUEST, (opus_int32 *)(x)
#define OPUS_GET_LAST_PACKET_DURATION(x) OPUS_GET_LAST_PACKET_DURATION_REQUEST, (opus_int32 *)(x)
#define OPUS_GET_PHASE_INVERSION_DISABLED(x) OPUS_GET_PHASE_INVERSION_DISABLED_REQUEST, (opus_int32 *)(x)

typedef struct OpusDecoder OpusDecoder;
typedef struct OpusEncoder OpusEncoder;
typedef struct OpusMSDecoder OpusMSDecoder;

const char *opus_strerror(int error);
const char *opus_get_version_string(void);

int opus_decoder_get_size(int channels);
OpusDecoder *opus_decoder_create(opus_int32 Fs, int channels, int *error);
int opus_decoder_init(OpusDecoder *st, opus_int32 Fs, int channels);
int opus_decode(OpusDecoder *st, const unsigned char *data, opus_int32 len, opus_int16 *pcm, int frame_size, int decode_fec);
int opus_decode_float(OpusDecoder *st, const unsigned char *data, opus_int32 len, float *pcm, int frame_size, int decode_fec);
int opus_decoder_ctl(OpusDecoder *st, int request, ...);
void opus_decoder_destroy(OpusDecoder *st);

int opus_encoder_get_size(int channels);
OpusEncoder *opus_encoder_create(opus_int32 Fs, int channels, int application, int *error);
int opus_encoder_init(OpusEncoder *st, opus_int32 Fs, int channels, int application);
opus_int32 opus_encode(OpusEncoder *st, const opus_int16 *pcm, int frame_size, unsigned char *data, opus_int32 max_data_bytes);
opus_int32 opus_encode_float(OpusEncoder *st, const float *pcm, int frame_size, unsigned char *data, opus_int32 max_data_bytes);
int opus_encoder_ctl(OpusEncoder *st, int request, ...);
void opus_encoder_destroy(OpusEncoder *st);

int opus_packet_parse(const unsigned char *data, opus_int32 len, unsigned char *out_toc, const unsigned char *frames[48], opus_int16 size[48],
                      int *payload_offset);
int opus_packet_get_bandwidth(const unsigned char *data);
int opus_packet_get_samples_per_frame(const unsigned char *data, opus_int32 Fs);
int opus_packet_get_nb_channels(const unsigned char *data);
int opus_packet_get_nb_frames(const unsigned char packet[], opus_int32 len);
int opus_packet_get_nb_samples(const unsigned char packet[], opus_int32 len, opus_int32 Fs);
int opus_decoder_get_nb_samples(const OpusDecoder *dec, const unsigned char packet[], opus_int32 len);

opus_int32 opus_multistream_decoder_get_size(int streams, int coupled_streams);
OpusMSDecoder *opus_multistream_decoder_create(opus_int32 Fs, int channels, int streams, int coupled_streams, const unsigned char *mapping, int *error);
int opus_multistream_decoder_init(OpusMSDecoder *st, opus_int32 Fs, int channels, int streams, int coupled_streams, const unsigned char *mapping);
int opus_multistream_decode(OpusMSDecoder *st, const unsigned char *data, opus_int32 len, opus_int16 *pcm, int frame_size, int decode_fec);
int opus_multistream_decode_float(OpusMSDecoder *st, const unsigned char *data, opus_int32 len, float *pcm, int frame_size, int decode_fec);
int opus_multistream_decoder_ctl(OpusMSDecoder *st, int request, ...);
void opus_multistream_decoder_destroy(OpusMSDecoder *st);

#ifdef __cplusplus
}
#endif
#endif
