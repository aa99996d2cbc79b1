// Counter-based random draws of the stochopy_b200 kernels.
//
// Philox4x32-10 (Salmon, Moraes, Dror, Shaw, SC'11) with
//   key     = (seed & 0xffffffff, seed >> 32)
//   counter = (block, row, generation, purpose)
// A draw depends only on what it is for -- not on the launch shape nor on the
// number of GPUs -- and is reproduced bit for bit by oracle/philox.py in the
// tests.  One block = 4 x u32 = 4 fp32 columns or 2 fp64 columns, i.e. exactly
// one 16-byte vector of a row tile.
#pragma once
#include <stdint.h>

namespace sp {

enum Purpose : uint32_t {
  kLhsJitter = 1,
  kDeCross = 2,
  kDeIndex = 3,
  kDeRepair = 4,
  kPsoR1 = 5,  // r1 AND r2 (16-bit pieces of one call, pso_r12)
  kPsoR2 = 6,  // retired
  kPsoRestart = 7,
  kEsZ = 8,
  kVdInject = 9,
  kEsMean0 = 10,
  kVdV0 = 11,
  kNaWalk = 12,
};

template <int ROUNDS = 10>
__device__ __forceinline__ uint4 philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint64_t seed) {
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < ROUNDS; ++r) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0;
    c1 = lo1;
    c2 = n2;
    c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}

// the stream a purpose tag selects (run-time tag: uniform branch)
__device__ __forceinline__ uint4 philox4x32_for(uint32_t purpose, uint32_t c0, uint32_t c1, uint32_t c2, uint64_t seed);

// Philox4x32-10 with the ten round keys taken from kernel parameters (constant bank
// operands of the LOP3s) instead of being re-derived from the seed for every call.
struct PhiloxKeys {
  uint32_t k[20];
};
inline PhiloxKeys philox_keys(uint64_t seed) {
  PhiloxKeys r;
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  for (int i = 0; i < 10; ++i) {
    r.k[2 * i] = k0;
    r.k[2 * i + 1] = k1;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return r;
}
// ROUNDS: 10 everywhere except the N(0,1) draws of the evolution strategies (purpose kEsZ), which take the
// 7-round variant -- the fewest rounds for which Philox4x32 passes BigCrush (Salmon et al., SC'11, table 2);
// mirrored by oracle/philox.py (ROUNDS_BY_PURPOSE).
constexpr int kEsZRounds = 7;
template <int ROUNDS = 10>
__device__ __forceinline__ uint4 philox4x32_keyed(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const PhiloxKeys& K) {
#pragma unroll
  for (int r = 0; r < ROUNDS; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ K.k[2 * r], n2 = hi0 ^ c3 ^ K.k[2 * r + 1];
    c0 = n0;
    c1 = lo1;
    c2 = n2;
    c3 = lo0;
  }
  return make_uint4(c0, c1, c2, c3);
}

// ---- DE crossover stream: Philox2x32-10 on 16-bit pieces ------------------------------------------
// The binomial crossover needs one Bernoulli(CR) decision per coordinate -- a 16-bit uniform
// (CR resolved to 2^-16 = 1.5e-5) is plenty, and four of them come out of ONE Philox2x32-10 call
// (20 instructions per 4 coordinates instead of ~44 for Philox4x32-10 on 32-bit words).
//   stream id   (key, o0, o1) = words x, y, z of Philox4x32-10(counter = (0, 0, generation, kDeCross), key = seed),
//               derived on the host once per generation (de_cross_keys)
//   call        (x, y) = Philox2x32-10(counter = (row + o0, group ^ o1), key),  group = column / 4
//   column j    piece (j & 3) of (x >> 16, x & 0xffff, y >> 16, y & 0xffff);  u = piece * 2^-16;  take iff u <= CR
// mirrored bit for bit by oracle/philox.py::de_cross_uniform.
struct CrossKeys {
  uint32_t k[10];  // round keys key + r * 0x9E3779B9
  uint32_t o0, o1;
  uint32_t cut_hi;  // piece <= floor(CR * 65536)  <=>  (piece << 16 | low bits) <= cut_hi
  uint32_t pad;
};
inline void philox4x32_host(uint32_t (&c)[4], uint64_t seed) {
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
    c[0] = n0;
    c[1] = (uint32_t)p1;
    c[2] = n2;
    c[3] = (uint32_t)p0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}
inline CrossKeys de_cross_keys(uint64_t seed, int it, double CR_as_T) {
  uint32_t c[4] = {0u, 0u, (uint32_t)it, (uint32_t)kDeCross};
  philox4x32_host(c, seed);
  CrossKeys r;
  for (int i = 0; i < 10; ++i) r.k[i] = c[0] + (uint32_t)i * 0x9E3779B9u;
  r.o0 = c[1];
  r.o1 = c[2];
  const double t = CR_as_T * 65536.0;
  const uint32_t cut = t < 0.0 ? 0u : (t >= 65536.0 ? 65536u : (uint32_t)t);
  r.cut_hi = cut >= 65536u ? 0xFFFFFFFFu : ((cut << 16) | 0xFFFFu);
  r.pad = 0;
  return r;
}
__device__ __forceinline__ uint2 philox2x32_keyed(uint32_t c0, uint32_t c1, const CrossKeys& K) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi = __umulhi(0xD256D193u, c0), lo = 0xD256D193u * c0;
    c0 = hi ^ K.k[r] ^ c1;
    c1 = lo;
  }
  return make_uint2(c0, c1);
}
// the four crossover decisions of column group `group` of `row`
__device__ __forceinline__ void de_cross_take(uint32_t row, uint32_t group, const CrossKeys& K, bool (&take)[4]) {
  const uint2 w = philox2x32_keyed(row + K.o0, group ^ K.o1, K);
  take[0] = w.x <= K.cut_hi;
  take[1] = (w.x << 16) <= K.cut_hi;
  take[2] = w.y <= K.cut_hi;
  take[3] = (w.y << 16) <= K.cut_hi;
}

// ---- PSO velocity coefficients: r1 and r2 of four columns from ONE Philox4x32-10 call -------------
//   (w0..w3) = Philox4x32-10(counter = (column / 4, row, generation, kPsoR1), key = seed)
//   column j = 4 g + p:  r1 = (w_p >> 16) * 2^-16,  r2 = (w_p & 0xffff) * 2^-16   (16-bit uniforms, exact in fp32)
// instead of two calls and eight 24-bit conversions; mirrored by oracle/philox.py::pso_uniforms.
// fp32 conversion without an integer->float instruction: the piece becomes the low mantissa bits of 2^23
// (one PRMT), and f * 2^-16 - 128 is exact (one FFMA).
__device__ __forceinline__ float piece_hi_f32(uint32_t w) {
  return fmaf(__uint_as_float(__byte_perm(w, 0x4B000000u, 0x7632)), 1.52587890625e-05f, -128.0f);
}
__device__ __forceinline__ float piece_lo_f32(uint32_t w) {
  return fmaf(__uint_as_float(__byte_perm(w, 0x4B000000u, 0x7610)), 1.52587890625e-05f, -128.0f);
}
// j0: first column of this lane's vector (multiple of VEC)
__device__ __forceinline__ void pso_r12(uint4 w, int j0, float (&r1)[4], float (&r2)[4]) {
  const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    r1[e] = piece_hi_f32(ww[e]);
    r2[e] = piece_lo_f32(ww[e]);
  }
}
__device__ __forceinline__ void pso_r12(uint4 w, int j0, double (&r1)[2], double (&r2)[2]) {
  const uint32_t a = (j0 & 2) ? w.z : w.x, b = (j0 & 2) ? w.w : w.y;
  r1[0] = (double)(a >> 16) * 1.52587890625e-05;
  r2[0] = (double)(a & 0xffffu) * 1.52587890625e-05;
  r1[1] = (double)(b >> 16) * 1.52587890625e-05;
  r2[1] = (double)(b & 0xffffu) * 1.52587890625e-05;
}

__device__ __forceinline__ uint4 philox4x32_for(uint32_t purpose, uint32_t c0, uint32_t c1, uint32_t c2, uint64_t seed) {
  return purpose == kEsZ ? philox4x32<kEsZRounds>(c0, c1, c2, purpose, seed) : philox4x32<10>(c0, c1, c2, purpose, seed);
}

// U[0,1) for one vector of a row: fp32 -> 4 x 24-bit, fp64 -> 2 x 53-bit
__device__ __forceinline__ void uniform_block(uint4 o, float (&u)[4]) {
  u[0] = __uint2float_rn(o.x >> 8) * 5.9604644775390625e-08f;
  u[1] = __uint2float_rn(o.y >> 8) * 5.9604644775390625e-08f;
  u[2] = __uint2float_rn(o.z >> 8) * 5.9604644775390625e-08f;
  u[3] = __uint2float_rn(o.w >> 8) * 5.9604644775390625e-08f;
}
__device__ __forceinline__ void uniform_block(uint4 o, double (&u)[2]) {
  unsigned long long a = ((unsigned long long)o.x << 21) | (o.y >> 11);
  unsigned long long b = ((unsigned long long)o.z << 21) | (o.w >> 11);
  u[0] = __ull2double_rn(a) * 1.1102230246251565e-16;
  u[1] = __ull2double_rn(b) * 1.1102230246251565e-16;
}

// N(0,1) by Box-Muller on the same block.  fp32 (definition mirrored by oracle/philox.py::normal):
// every word gives a 23-bit fraction f = as_float(0x3f800000 | (w & 0x7fffff)) in [1, 2) with ONE
// logic instruction (no integer->float conversion: those share the 16-lane XU pipe with the MUFUs);
//   radius   u = 2 - f in (0, 1],  x = f - 1 = 1 - u (both exact),  r = sqrt(-2 ln u)
//   angle    t = f' - 1.5 in [-0.5, 0.5),  (z0, z1) = r (cos 2 pi t, sin 2 pi t)
// with the SFU lg2 / sqrt / sin / cos approximations (abs error ~2^-21 on reduced arguments);
// -2 ln u switches to the series of -2 ln(1 - x) for x < 2^-6 so small radii keep their relative
// accuracy.  |z - exact| <= 4e-6 + 2e-6 |z| (tests/test_gpu_es.py::test_normal_draws_match_oracle).
__device__ __forceinline__ float fast_sqrt(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float unit_fraction(uint32_t w) {  // (w & 0x7fffff) | 0x3f800000 as ONE lop3 (both masks in registers)
  uint32_t r;
  asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(r) : "r"(w), "r"(0x007fffffu), "r"(0x3f800000u));
  return __uint_as_float(r);
}
__device__ __forceinline__ float fast_log2(float x) {  // x is never subnormal here: no range fix-up around the MUFU
  float r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float radius_from(uint32_t w) {  // sqrt(-2 ln u), u = 2 - f
  const float f = unit_fraction(w), x = f - 1.0f;
  const float series = x * (2.0f + x * (1.0f + x * (0.66666669f + 0.5f * x)));
  const float viaLog = -1.3862944f * fast_log2(2.0f - f);
  return fast_sqrt(x >= 0.015625f ? viaLog : series);
}
__device__ __forceinline__ void normal_pair(uint32_t wr, uint32_t wa, float* z0, float* z1) {
  const float r = radius_from(wr);
  const float a = fmaf(unit_fraction(wa), 6.2831855f, -9.424778f);  // 2 pi (f - 1.5) in [-pi, pi)
  *z0 = r * __cosf(a);
  *z1 = r * __sinf(a);
}
__device__ __forceinline__ void normal_block(uint4 o, float (&z)[4]) {
  normal_pair(o.x, o.y, &z[0], &z[1]);
  normal_pair(o.z, o.w, &z[2], &z[3]);
}
__device__ __forceinline__ void normal_block(uint4 o, double (&z)[2]) {
  unsigned long long a = ((unsigned long long)o.x << 21) | (o.y >> 11);
  unsigned long long b = ((unsigned long long)o.z << 21) | (o.w >> 11);
  double u1 = (__ull2double_rn(a) + 1.0) * 1.1102230246251565e-16;
  double u2 = __ull2double_rn(b) * 1.1102230246251565e-16;
  double r = sqrt(-2.0 * log(u1));
  double sn, cs;
  sincospi(2.0 * u2, &sn, &cs);
  z[0] = r * cs;
  z[1] = r * sn;
}

// integer in [0, n) from one word (multiply-shift; bias < n / 2^32)
__device__ __forceinline__ uint32_t bounded(uint32_t word, uint32_t n) { return __umulhi(word, n); }

// ---- keyed permutation of [0, P) for the Latin hypercube ----------------------
__device__ __forceinline__ uint32_t mix32(uint32_t h) {
  h ^= h >> 16;
  h *= 0x85EBCA6Bu;
  h ^= h >> 13;
  h *= 0xC2B2AE35u;
  h ^= h >> 16;
  return h;
}
// 4-round Feistel on 2*half bits + cycle walking (bijective on [0,P))
__device__ __forceinline__ uint32_t lhs_permute(uint32_t i, uint32_t P, uint32_t col, uint64_t seed) {
  uint32_t bits = 32 - __clz(P - 1);
  if (bits < 2) bits = 2;
  const uint32_t half = (bits + 1) >> 1, mask = (1u << half) - 1u;
  const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32), colk = col * 0x9E3779B1u;
  uint32_t v = i;
  do {
    uint32_t L = v >> half, R = v & mask;
#pragma unroll
    for (uint32_t rnd = 0; rnd < 4; ++rnd) {
      uint32_t f = mix32(R ^ colk ^ ((rnd & 1u) ? k1 : k0) ^ ((rnd + 1u) * 0x7F4A7C15u));
      uint32_t nr = (L ^ f) & mask;
      L = R;
      R = nr;
    }
    v = (L << half) | R;
  } while (v >= P);
  return v;
}

}  // namespace sp
