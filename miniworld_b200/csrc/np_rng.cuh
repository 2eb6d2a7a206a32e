// np_rng.cuh -- numpy's Generator(PCG64) stream, draw for draw.
//
// The reference draws every random number through gymnasium's `self.np_random`
// (Generator(PCG64(SeedSequence(seed))), reference miniworld.py:551): placement
// (`choice(n, p)`, `uniform(low[3], high[3])`, `uniform(-pi, pi)`, miniworld.py:865-901),
// texture variants (`integers(0, n)`, opengl.py:133), level choices (pickupobjects.py:64-65)
// and, under domain randomisation, three `uniform` draws per step (miniworld.py:677-680,
// params.py:99-101).  numpy is a third-party dependency of the reference (numpy>=1.22;
// this image: 2.3.5); the algorithms below restate its published PCG64 / Lemire code paths
//  so that resets and domain-rand steps can run on the device and
// still land on the reference's exact stream.  Seeding (SeedSequence) stays on the host:
// the initial 128-bit state / increment are uploaded with mwb_seed().
#pragma once
#include "hd.h"

struct NpRng {
  uint64_t s_hi, s_lo, inc_hi, inc_lo;
  int32_t has32;
  uint32_t cache;
};

// state = state * 0x2360ED051FC65DA44385DF649FCCF645 + inc (mod 2^128), then XSL-RR output
MWB_DEV uint64_t rng_next64(NpRng& r) {
  const uint64_t MH = 0x2360ED051FC65DA4ULL, ML = 0x4385DF649FCCF645ULL;
  uint64_t lo = r.s_lo * ML;
  uint64_t hi = umulhi64(r.s_lo, ML) + r.s_hi * ML + r.s_lo * MH;
  uint64_t lo2 = lo + r.inc_lo;
  hi = hi + r.inc_hi + (lo2 < lo ? 1u : 0u);
  r.s_lo = lo2;
  r.s_hi = hi;
  uint64_t x = hi ^ lo2;
  unsigned rot = (unsigned)(hi >> 58);
  return (x >> rot) | (x << ((64u - rot) & 63u));
}

// low half of a fresh 64-bit draw; the high half is buffered for the next call
MWB_DEV uint32_t rng_next32(NpRng& r) {
  if (r.has32) {
    r.has32 = 0;
    return r.cache;
  }
  uint64_t v = rng_next64(r);
  r.has32 = 1;
  r.cache = (uint32_t)(v >> 32);
  return (uint32_t)v;
}

// Generator.random(): 53-bit mantissa
MWB_DEV double rng_random(NpRng& r) { return (double)(rng_next64(r) >> 11) * (1.0 / 9007199254740992.0); }

// Generator.uniform(lo, hi) given lo and (hi - lo)
MWB_DEV double rng_uniform(NpRng& r, double lo, double range) { return d_add(lo, d_mul(range, rng_random(r))); }

// Generator.integers(0, n) for 1 <= n <= 2^32: Lemire's bounded method on 32-bit draws;
// n == 1 consumes nothing.
MWB_DEV uint32_t rng_integers(NpRng& r, uint32_t n) {
  uint32_t rng = n - 1u;
  if (rng == 0u) return 0u;
  if (rng == 0xFFFFFFFFu) return rng_next32(r);
  uint32_t excl = rng + 1u;
  uint64_t m = (uint64_t)rng_next32(r) * (uint64_t)excl;
  uint32_t left = (uint32_t)m;
  if (left < excl) {
    uint32_t thr = (0xFFFFFFFFu - rng) % excl;
    while (left < thr) {
      m = (uint64_t)rng_next32(r) * (uint64_t)excl;
      left = (uint32_t)m;
    }
  }
  return (uint32_t)(m >> 32);
}
