// rangedec.cuh — Opus range decoder as warp-uniform device code.
//
// Every lane of the warp that owns a stream holds an identical copy of the decoder state in
// registers and executes the same instruction stream (no divergence, no communication); the
// packet bytes sit in shared memory, so each byte fetch is one broadcast LDS.
// Semantics: RFC 6716 section 4.1; reference celt/entdec.c:91-266, celt/entcode.c:69-88,
// celt/laplace.c:94-134.  All arithmetic is 32-bit unsigned and bit-exact.
#pragma once
#include "simt.cuh"

namespace ob {

struct RangeDec {
  const unsigned char *buf;  // shared memory
  uint32_t storage, end_offs, end_window;
  int nend_bits, nbits_total;
  uint32_t offs, rng, val, ext;
  int rem, error;

  OB_MDEV int next_byte() { return offs < storage ? buf[offs++] : 0; }
  OB_MDEV int next_byte_from_end() { return end_offs < storage ? buf[storage - ++end_offs] : 0; }

  // entdec.c:102-117
  OB_MDEV void renorm() {
    while (rng <= (1u << 23)) {
      nbits_total += 8;
      rng <<= 8;
      int sym = rem;
      rem = next_byte();
      sym = (sym << 8 | rem) >> 1;
      val = ((val << 8) + (255 & ~sym)) & 0x7FFFFFFFu;
    }
  }
  // entdec.c:119-137
  OB_MDEV void init(const unsigned char *b, uint32_t len) {
    buf = b; storage = len; end_offs = 0; end_window = 0; nend_bits = 0;
    nbits_total = 9; offs = 0; rng = 128; ext = 0;
    rem = next_byte();
    val = rng - 1 - (rem >> 1);
    error = 0;
    renorm();
  }
  OB_MDEV int tell() const { return nbits_total - ilog32(rng); }
  // entcode.c:69-88
  OB_MDEV uint32_t tell_frac() const {
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
  OB_MDEV unsigned decode(unsigned ft) {
    ext = rng / ft;
    unsigned s = (unsigned)(val / ext);
    unsigned s1 = s + 1;
    return ft - (s1 < ft ? s1 : ft);
  }
  OB_MDEV unsigned decode_bin(unsigned bits) {
    ext = rng >> bits;
    unsigned s = (unsigned)(val / ext);
    unsigned s1 = s + 1u, top = 1u << bits;
    return top - (s1 < top ? s1 : top);
  }
  OB_MDEV void update(unsigned fl, unsigned fh, unsigned ft) {
    uint32_t s = ext * (ft - fh);
    val -= s;
    rng = fl > 0 ? ext * (fh - fl) : rng - s;
    renorm();
  }
  OB_MDEV int bit_logp(unsigned logp) {
    uint32_t r = rng, v = val, s = r >> logp;
    int ret = v < s;
    if (!ret) val = v - s;
    rng = ret ? s : r - s;
    renorm();
    return ret;
  }
  // icdf: table in global/constant memory
  OB_MDEV int icdf(const unsigned char *tab, unsigned ftb) {
    uint32_t s = rng, v = val, r = s >> ftb, t;
    int ret = -1;
    do { t = s; s = r * tab[++ret]; } while (v < s);
    val = v - s;
    rng = t - s;
    renorm();
    return ret;
  }
  OB_MDEV uint32_t bits(unsigned nb) {
    uint32_t window = end_window;
    int avail = nend_bits;
    if ((unsigned)avail < nb) {
      do { window |= (uint32_t)next_byte_from_end() << avail; avail += 8; } while (avail <= 24);
    }
    uint32_t ret = window & (((uint32_t)1 << nb) - 1u);
    window >>= nb; avail -= nb;
    end_window = window; nend_bits = avail; nbits_total += nb;
    return ret;
  }
  OB_MDEV uint32_t uint(uint32_t ft) {
    ft--;
    int ftb = ilog32(ft);
    if (ftb > 8) {
      ftb -= 8;
      unsigned f = (unsigned)(ft >> ftb) + 1;
      unsigned s = decode(f);
      update(s, s + 1, f);
      uint32_t t = (uint32_t)s << ftb | bits(ftb);
      if (t <= ft) return t;
      error = 1;
      return ft;
    } else {
      ft++;
      unsigned s = decode((unsigned)ft);
      update(s, s + 1, (unsigned)ft);
      return s;
    }
  }
  // laplace.c:94-134 (LAPLACE_MINP = 1, LAPLACE_NMIN = 16)
  OB_MDEV int laplace(unsigned fs, int decay) {
    int v = 0;
    unsigned fl = 0, fm = decode_bin(15);
    if (fm >= fs) {
      unsigned ft = 32768u - 32u - fs;
      v++;
      fl = fs;
      fs = ((ft * (uint32_t)(16384 - decay)) >> 15) + 1;
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
    update(fl, fh < 32768u ? fh : 32768u, 32768u);
    return v;
  }
};

}  // namespace ob
