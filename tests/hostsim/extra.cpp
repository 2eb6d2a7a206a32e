// Test-only exports of the device math helpers (compiled into libmwb_hostsim.so).
#include "../../miniworld_b200/csrc/libm_sincos.cuh"
#include "../../miniworld_b200/csrc/np_rng.cuh"

extern "C" void hs_sincos(const double* x, int n, double* s, double* c) {
  for (int i = 0; i < n; ++i) {
    s[i] = mwb_libm::sin_glibc(x[i]);
    c[i] = mwb_libm::cos_glibc(x[i]);
  }
}

// kind: 0 random(), 1 uniform(lo, lo + rng), 2 integers(0, n) -- one draw per entry
extern "C" void hs_rng_draws(uint64_t* st /*s_hi,s_lo,inc_hi,inc_lo*/, int* has32, uint32_t* cache, const int* kind,
                             const double* a, const double* b, int n, double* out) {
  NpRng r;
  r.s_hi = st[0]; r.s_lo = st[1]; r.inc_hi = st[2]; r.inc_lo = st[3];
  r.has32 = *has32; r.cache = *cache;
  for (int i = 0; i < n; ++i) {
    if (kind[i] == 0) out[i] = rng_random(r);
    else if (kind[i] == 1) out[i] = rng_uniform(r, a[i], b[i]);
    else out[i] = (double)rng_integers(r, (uint32_t)a[i]);
  }
  st[0] = r.s_hi; st[1] = r.s_lo;
  *has32 = r.has32; *cache = r.cache;
}
