// tests/emu/emu_decode.cpp — TEST INFRASTRUCTURE ONLY.
// Compiles the decoder KERNEL SOURCES (opus_b200/csrc/*.cuh) for the host SIMT emulator so that
// their logic can be unit-tested and debugged without a GPU.  Never linked into the product.
#define OB_HOST_EMU 1
#include "../../opus_b200/csrc/celt_dec_kernels.cuh"
#include "../../opus_b200/csrc/celt_dec_plc.cuh"

#include <stdio.h>
#include <math.h>

using namespace ob;

struct EmuDecoder {
  int n, channels, downsample, frame_size, disable_inv, decode_gain;
  float gain;
  std::vector<ObDecBandState> bstate;
  std::vector<ObDecSynthState> sstate;
  std::vector<ObFrameSide> side;
  std::vector<ObScript> script;
  std::vector<float> X;
};

static void entropy_entry(void *arg) {  // one thread per stream
  const DecEntropyParams &p = *(const DecEntropyParams *)arg;
  int stream = block_id() * 32 + lane_id();
  if (stream < p.n_streams) dec_entropy_thread(p, stream);
}
static void pvq_entry(void *arg) {  // one warp per stream (one warp per block in the emulator)
  const DecPvqParams &p = *(const DecPvqParams *)arg;
  PvqShared *tab = (PvqShared *)dyn_smem();
  for (int i = lane_id(); i < 1272; i += 32) tab->u[i] = ob_pvq_u[i];
  if (lane_id() < 15) tab->rowoff[lane_id()] = ob_pvq_rowoff[lane_id()];
  bsync();
  dec_pvq_warp(p, block_id(), tab, dyn_smem() + ((sizeof(PvqShared) + 15) & ~(size_t)15));
}
static void synth_entry(void *arg) {
  const DecSynthParams &p = *(const DecSynthParams *)arg;
  dec_synth_block(p, block_id(), dyn_smem());
}

static void clip_entry(void *arg) {
  const DecClipParams &p = *(const DecClipParams *)arg;
  dec_softclip_packet_block(p, block_id());
}
static void plc_entry(void *arg) {
  const DecPlcParams &p = *(const DecPlcParams *)arg;
  dec_plc_block(p, block_id(), dyn_smem());
}

extern "C" {

void *emu_decoder_create(int n_streams, int channels) {
  EmuDecoder *d = new EmuDecoder;
  d->n = n_streams;
  d->channels = channels;
  d->downsample = 1;
  d->frame_size = 5760;
  d->disable_inv = 0; d->decode_gain = 0; d->gain = 1.f;
  d->bstate.resize(n_streams);
  d->sstate.resize(n_streams);
  d->side.resize(n_streams);
  d->script.resize(n_streams);
  d->X.assign((size_t)n_streams * 2 * OB_MAXN, 0.f);
  memset(d->bstate.data(), 0, sizeof(ObDecBandState) * n_streams);
  memset(d->sstate.data(), 0, sizeof(ObDecSynthState) * n_streams);
  for (int s = 0; s < n_streams; s++)
    for (int i = 0; i < 2 * OB_NBANDS; i++) d->bstate[s].oldLogE[i] = d->bstate[s].oldLogE2[i] = -28.f;
  return d;
}
void emu_decoder_set_fs(void *h, int Fs) { ((EmuDecoder *)h)->downsample = 48000 / Fs; }
// OPUS_SET_GAIN (4034) / OPUS_SET_PHASE_INVERSION_DISABLED (4046), as opusb200_decoder_ctl
void emu_decoder_ctl(void *h, int request, int value) {
  EmuDecoder *d = (EmuDecoder *)h;
  if (request == 4034) { d->decode_gain = value; d->gain = (float)exp(0.6931471805599453094 * (6.48814081e-4f * (float)value)); }
  if (request == 4046) d->disable_inv = value;
}
int emu_decoder_get_ctl(void *h, int stream, int request) {
  const ObDecBandState &b = ((EmuDecoder *)h)->bstate[stream];
  if (request == 4033) return (b.plc_flags & OB_PF_HAS_PACKET) ? b.pf_period : 0;
  if (request == 4039) return b.last_status;
  return -1;
}
void emu_decoder_set_frame_size(void *h, int fs) { ((EmuDecoder *)h)->frame_size = fs; }  // opus_decode's frame_size (output-rate samples)
// test hook: overwrite a stream's synthesis history with the oracle's (decode_mem[0 .. 2048+120), de-emphasis memory), so that a
// concealed frame can be compared without the last-bit differences of the frames before it
void emu_decoder_sync_history(void *h, int stream, int c, const float *mem, float preemph) {
  ObDecSynthState &st = ((EmuDecoder *)h)->sstate[stream];
  for (int k = 0; k < OB_HIST; k++) st.hist[c][(st.pos + (uint32_t)k) & (OB_HIST - 1)] = mem[k];
  for (int i = 0; i < OB_OVERLAP; i++) st.tail[c][i] = mem[OB_HIST + i];
  st.deemph_mem[c] = preemph;
}
void emu_decoder_destroy(void *h) { delete (EmuDecoder *)h; }

// packets: [n][stride]; pcm: [n][pcm_stride] (float or int16); out_samples[n]; final_range[n]
int emu_decoder_decode(void *h, const unsigned char *packets, int stride, const int *lens, void *pcm, int pcm_stride,
                       int pcm_is_int16, int *out_samples, unsigned *final_range) {
  EmuDecoder *d = (EmuDecoder *)h;
  // like the CUDA host side: int16 output goes through a float staging area and one packet-wide soft clip (dec_softclip_packet_block)
  // whenever a packet may hold several frames; the emulator always allows that
  const int defer_clip = pcm_is_int16;
  const int stage_stride = d->frame_size * d->channels;
  std::vector<float> stage(defer_clip ? (size_t)d->n * stage_stride : 0, 0.f);
  for (int k = 0; k < 48; k++) {
    DecEntropyParams ep;
    ep.packets = packets; ep.pkt_len = lens; ep.pkt_stride = stride; ep.n_streams = d->n; ep.channels = d->channels;
    ep.frame_index = k; ep.state = d->bstate.data(); ep.side = d->side.data(); ep.script = d->script.data();
    ep.out_samples = out_samples; ep.max_samples = d->frame_size; ep.disable_inv = d->disable_inv; ep.downsample = d->downsample;
    emu_launch(entropy_entry, &ep, (d->n + 31) / 32, 32, 0);
    DecPvqParams pp;
    pp.side = d->side.data(); pp.script = d->script.data(); pp.state = d->bstate.data(); pp.X = d->X.data(); pp.n_streams = d->n;
    emu_launch(pvq_entry, &pp, d->n, 32, ((sizeof(PvqShared) + 15) & ~(size_t)15) + sizeof(PvqWS));
    int any = 0;
    for (int s = 0; s < d->n; s++) {
      int st = d->side[s].status;
      if (st > 0) any = 1;
    }
    if (!any) break;
    // synthesis needs per-stream pcm offsets (streams may have different frame sizes): launch per distinct offset
    // (emulator only; the CUDA host side uses a per-stream offset array)
    for (int s = 0; s < d->n; s++) {
      if (d->side[s].status <= 0) continue;
      DecSynthParams sp;
      sp.side = d->side.data() + s; sp.X = d->X.data() + (size_t)s * 2 * OB_MAXN; sp.state = d->sstate.data() + s;
      sp.pcm = pcm_is_int16 ? (void *)((int16_t *)pcm + (size_t)s * pcm_stride) : (void *)((float *)pcm + (size_t)s * pcm_stride);
      sp.pcm_stride = 0; sp.frame_index = k; sp.pcm_is_int16 = pcm_is_int16; sp.n_streams = 1; sp.channels = d->channels; sp.downsample = d->downsample; sp.apply_gain = d->decode_gain != 0; sp.gain = d->gain;
      if (defer_clip) { sp.pcm = stage.data() + (size_t)s * stage_stride; sp.pcm_is_int16 = 0; sp.defer_clip = 1; }
      const int lost_here = OB_PLC_TYPE(d->side[s].plc) >= OB_PLC_PITCH;   // (emulator only: skip the launch that would return at once)
      if (!lost_here) emu_launch(synth_entry, &sp, 1, 32 * d->channels, sizeof(SynthWS) * d->channels);
      DecPlcParams lp;
      lp.side = sp.side; lp.state = sp.state; lp.pcm = sp.pcm; lp.pcm_stride = 0; lp.pcm_is_int16 = sp.pcm_is_int16; lp.defer_clip = sp.defer_clip; lp.n_streams = 1;
      lp.channels = d->channels; lp.downsample = d->downsample; lp.apply_gain = sp.apply_gain; lp.gain = sp.gain;
      if (lost_here) emu_launch(plc_entry, &lp, 1, 64, sizeof(PlcWS));
    }
  }
  if (defer_clip) {
    DecClipParams cp;
    cp.stage = stage.data(); cp.stage_stride = stage_stride; cp.pcm = (int16_t *)pcm; cp.pcm_stride = pcm_stride; cp.out_samples = out_samples;
    cp.pkt_len = lens; cp.state = d->sstate.data(); cp.n_streams = d->n; cp.channels = d->channels;
    emu_launch(clip_entry, &cp, d->n, 32 * d->channels, 0);
  }
  for (int s = 0; s < d->n; s++) final_range[s] = d->bstate[s].final_range;
  return 0;
}

const float *emu_decoder_X(void *h, int stream) { return ((EmuDecoder *)h)->X.data() + (size_t)stream * 2 * OB_MAXN; }
const ObFrameSide *emu_decoder_side(void *h, int stream) { return &((EmuDecoder *)h)->side[stream]; }

long emu_yield_count(void) { return emu().yields; }
}
