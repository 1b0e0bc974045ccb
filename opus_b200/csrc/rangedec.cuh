// rangedec.cuh — Opus range decoder as warp-uniform device code.
//
// Every lane of the warp that owns a stream holds an identical copy of the decoder state in
// registers and executes the same instruction stream (no divergence, no communication); the
// packet bytes sit in shared memory, so each byte fetch is one broadcast LDS.
// Semantics: RFC 6716 section 4.1; reference celt/entdec.c:91-266, celt/entcode.c:69-88,
// celt/laplace.c:94-134.  All arithmetic is 32-bit unsigned and bit-exact.
//
// Code-size discipline: the band decoder is a large, branchy program and every warp of an SM is
// at a different place in it, so the instruction cache is the scarce resource (first profile:
// 72 % of stall samples were "no instruction").  Each primitive therefore exists ONCE as a
// non-inlined function that takes and returns the whole decoder state BY VALUE — the sm_100a ABI
// passes such a 14-word struct in registers, so there is no local-memory traffic — and the
// member functions below are thin inline wrappers that keep call sites readable.
#pragma once
#include "simt.cuh"

namespace ob {

struct RangeDec {
  const unsigned char *buf;  // shared memory
  uint32_t storage, end_offs, end_window;
  int nend_bits, nbits_total;
  uint32_t offs, rng, val, ext;
  int rem, error;
  int ret;  // result of the last primitive

  OB_MDEV void init(const unsigned char *b, uint32_t len);
  OB_MDEV int tell() const { return nbits_total - ilog32(rng); }
  OB_MDEV uint32_t tell_frac() const;
  OB_MDEV unsigned decode(unsigned ft);
  OB_MDEV void update(unsigned fl, unsigned fh, unsigned ft);
  OB_MDEV int bit_logp(unsigned logp);
  OB_MDEV int icdf(const unsigned char *tab, unsigned ftb);
  OB_MDEV uint32_t bits(unsigned nb);
  OB_MDEV uint32_t uint(uint32_t ft);
  OB_MDEV int laplace(unsigned fs, int decay);
};

// entdec.c:102-117 (inlined into the primitives below: they are the only callers)
OB_DEV void rd_renorm(RangeDec &d) {
#pragma unroll 1
  while (d.rng <= (1u << 23)) {
    d.nbits_total += 8;
    d.rng <<= 8;
    int sym = d.rem;
    d.rem = d.offs < d.storage ? d.buf[d.offs++] : 0;
    sym = (sym << 8 | d.rem) >> 1;
    d.val = ((d.val << 8) + (255 & ~sym)) & 0x7FFFFFFFu;
  }
}
// entdec.c:119-137
OB_DEVNI RangeDec rd_init(const unsigned char *b, uint32_t len) {
  RangeDec d;
  d.buf = b; d.storage = len; d.end_offs = 0; d.end_window = 0; d.nend_bits = 0;
  d.nbits_total = 9; d.offs = 0; d.rng = 128; d.ext = 0; d.ret = 0;
  d.rem = d.offs < d.storage ? d.buf[d.offs++] : 0;
  d.val = d.rng - 1 - (d.rem >> 1);
  d.error = 0;
  rd_renorm(d);
  return d;
}
// entcode.c:69-88
OB_DEVNI uint32_t rd_tell_frac(uint32_t rng, int nbits_total) {
  uint32_t nbits = (uint32_t)nbits_total << 3;
  int l = ilog32(rng);
  uint32_t r = rng >> (l - 16);
  unsigned b = (r >> 12) - 8;
  // thresholds 2^16 * 2^((k+1)/8 - 1), k=0..7 (exact transition points of the 1/8-bit log)
  unsigned thr = b == 0 ? 35733u : b == 1 ? 38967u : b == 2 ? 42495u : b == 3 ? 46340u
               : b == 4 ? 50535u : b == 5 ? 55109u : b == 6 ? 60097u : 65535u;
  b += r > thr;
  l = (l << 3) + b;
  return nbits - l;
}
// entdec.c:139-144 / 146-151
OB_DEVNI RangeDec rd_decode(RangeDec d, unsigned ft) {
  d.ext = d.rng / ft;
  unsigned s = (unsigned)(d.val / d.ext);
  unsigned s1 = s + 1;
  d.ret = ft - (s1 < ft ? s1 : ft);
  return d;
}
// entdec.c:153-160
OB_DEVNI RangeDec rd_update(RangeDec d, unsigned fl, unsigned fh, unsigned ft) {
  uint32_t s = d.ext * (ft - fh);
  d.val -= s;
  d.rng = fl > 0 ? d.ext * (fh - fl) : d.rng - s;
  rd_renorm(d);
  return d;
}
// entdec.c:162-177
OB_DEVNI RangeDec rd_bit_logp(RangeDec d, unsigned logp) {
  uint32_t r = d.rng, v = d.val, s = r >> logp;
  int ret = v < s;
  if (!ret) d.val = v - s;
  d.rng = ret ? s : r - s;
  rd_renorm(d);
  d.ret = ret;
  return d;
}
// entdec.c:179-198; table in global memory
OB_DEVNI RangeDec rd_icdf(RangeDec d, const unsigned char *tab, unsigned ftb) {
  uint32_t s = d.rng, v = d.val, r = s >> ftb, t;
  int ret = -1;
#pragma unroll 1
  do { t = s; s = r * tab[++ret]; } while (v < s);
  d.val = v - s;
  d.rng = t - s;
  rd_renorm(d);
  d.ret = ret;
  return d;
}
// entdec.c:246-266
OB_DEVNI RangeDec rd_bits(RangeDec d, unsigned nb) {
  uint32_t window = d.end_window;
  int avail = d.nend_bits;
  if ((unsigned)avail < nb) {
#pragma unroll 1
    do {
      int byte = d.end_offs < d.storage ? d.buf[d.storage - ++d.end_offs] : 0;
      window |= (uint32_t)byte << avail;
      avail += 8;
    } while (avail <= 24);
  }
  d.ret = (int)(window & (((uint32_t)1 << nb) - 1u));
  window >>= nb; avail -= nb;
  d.end_window = window; d.nend_bits = avail; d.nbits_total += nb;
  return d;
}
// entdec.c:219-244
OB_DEVNI RangeDec rd_uint(RangeDec d, uint32_t ft) {
  ft--;
  int ftb = ilog32(ft);
  if (ftb > 8) {
    ftb -= 8;
    unsigned f = (unsigned)(ft >> ftb) + 1;
    d = rd_decode(d, f);
    unsigned s = d.ret;
    d = rd_update(d, s, s + 1, f);
    d = rd_bits(d, ftb);
    uint32_t t = (uint32_t)s << ftb | (uint32_t)d.ret;
    if (t <= ft) { d.ret = (int)t; return d; }
    d.error = 1;
    d.ret = (int)ft;
    return d;
  } else {
    ft++;
    d = rd_decode(d, (unsigned)ft);
    unsigned s = d.ret;
    d = rd_update(d, s, s + 1, (unsigned)ft);
    d.ret = (int)s;
    return d;
  }
}
// laplace.c:94-134 (LAPLACE_MINP = 1, LAPLACE_NMIN = 16)
OB_DEVNI RangeDec rd_laplace(RangeDec d, unsigned fs, int decay) {
  int v = 0;
  d.ext = d.rng >> 15;  // ec_decode_bin(15)
  unsigned sq = (unsigned)(d.val / d.ext), s1 = sq + 1u;
  unsigned fl = 0, fm = 32768u - (s1 < 32768u ? s1 : 32768u);
  if (fm >= fs) {
    unsigned ft = 32768u - 32u - fs;
    v++;
    fl = fs;
    fs = ((ft * (uint32_t)(16384 - decay)) >> 15) + 1;
#pragma unroll 1
    while (fs > 1 && fm >= fl + 2 * fs) {
      fs *= 2;
      fl += fs;
      fs = ((fs - 2) * (uint32_t)decay) >> 15;
      fs += 1;
      v++;
    }
    if (fs <= 1) {
      int di = (fm - fl) >> 1;
      v += di;
      fl += 2 * di;
    }
    if (fm < fl + fs) v = -v; else fl += fs;
  }
  unsigned fh = fl + fs;
  d = rd_update(d, fl, fh < 32768u ? fh : 32768u, 32768u);
  d.ret = v;
  return d;
}

OB_MDEV void RangeDec::init(const unsigned char *b, uint32_t len) { *this = rd_init(b, len); }
OB_MDEV uint32_t RangeDec::tell_frac() const { return rd_tell_frac(rng, nbits_total); }
OB_MDEV unsigned RangeDec::decode(unsigned ft) { *this = rd_decode(*this, ft); return (unsigned)ret; }
OB_MDEV void RangeDec::update(unsigned fl, unsigned fh, unsigned ft) { *this = rd_update(*this, fl, fh, ft); }
OB_MDEV int RangeDec::bit_logp(unsigned logp) { *this = rd_bit_logp(*this, logp); return ret; }
OB_MDEV int RangeDec::icdf(const unsigned char *tab, unsigned ftb) { *this = rd_icdf(*this, tab, ftb); return ret; }
OB_MDEV uint32_t RangeDec::bits(unsigned nb) { *this = rd_bits(*this, nb); return (uint32_t)ret; }
OB_MDEV uint32_t RangeDec::uint(uint32_t ft) { *this = rd_uint(*this, ft); return (uint32_t)ret; }
OB_MDEV int RangeDec::laplace(unsigned fs, int decay) { *this = rd_laplace(*this, fs, decay); return ret; }

}  // namespace ob
