// Host-only helpers of libfastfp_b200.so (no device code).
//
// fastfp_hash64: a 64-bit content hash of a host buffer, fast enough (memory-bandwidth class, several
// threads for large buffers) that the Python mirror can fingerprint EVERY byte of the caller's
// (Nvecs, Ts, sigmas/TNTs) lists on each call: the device pack is a cache of those arrays, and the
// reference is a pure function of its arguments (fastfp/fastfp.py:52), so an in-place edit of any entry
// must rebuild the pack. (Round 1 sampled 16 points per array; an edit between the samples went unseen.)
#include <cstring>
#include <thread>
#include <vector>

#include "../../include/fastfp_b200.h"

namespace {

inline uint64_t rotl(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
inline uint64_t mix(uint64_t h) {
  h ^= h >> 33; h *= 0xff51afd7ed558ccdULL; h ^= h >> 33; h *= 0xc4ceb9fe1a85ec53ULL; h ^= h >> 33;
  return h;
}

// four independent multiply-rotate lanes over 32-byte stripes (the xxHash64 round structure), so the
// loop runs at load bandwidth; the tail is folded byte-wise. Position-dependent: permutations change it.
uint64_t hash_block(const unsigned char* p, size_t n, uint64_t seed) {
  const uint64_t P1 = 0x9E3779B185EBCA87ULL, P2 = 0xC2B2AE3D27D4EB4FULL;
  uint64_t v[4] = {seed + P1 + P2, seed + P2, seed, seed - P1};
  size_t i = 0;
  for (; i + 32 <= n; i += 32) {
    uint64_t w[4];
    std::memcpy(w, p + i, 32);
    for (int k = 0; k < 4; ++k) v[k] = rotl(v[k] + w[k] * P2, 31) * P1;
  }
  uint64_t h = rotl(v[0], 1) + rotl(v[1], 7) + rotl(v[2], 12) + rotl(v[3], 18) + (uint64_t)n;
  for (; i < n; ++i) h = rotl(h ^ (p[i] * P1), 11) * P2;
  return mix(h);
}

}  // namespace

extern "C" uint64_t fastfp_hash64(const void* data, int64_t nbytes, uint64_t seed) {
  if (!data || nbytes <= 0) return mix(seed ^ 0x51ed270b7a1c2d4fULL);
  const unsigned char* p = static_cast<const unsigned char*>(data);
  const size_t n = (size_t)nbytes, BLK = (size_t)4 << 20;  // fixed blocks: the value does not depend on the thread count
  const size_t nblk = (n + BLK - 1) / BLK;
  std::vector<uint64_t> part(nblk);
  auto work = [&](size_t b0, size_t b1) {
    for (size_t b = b0; b < b1; ++b) {
      const size_t lo = b * BLK, len = lo + BLK <= n ? BLK : n - lo;
      part[b] = hash_block(p + lo, len, seed + b);
    }
  };
  unsigned nt = std::thread::hardware_concurrency();
  nt = nt > 16 ? 16 : (nt < 1 ? 1 : nt);
  if (nblk < 4 || nt == 1) {
    work(0, nblk);
  } else {
    if (nt > nblk) nt = (unsigned)nblk;
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; ++t) th.emplace_back(work, nblk * t / nt, nblk * (t + 1) / nt);
    for (auto& t : th) t.join();
  }
  uint64_t h = seed ^ (uint64_t)n;
  for (size_t b = 0; b < nblk; ++b) h = mix(h ^ part[b]) + 0x9E3779B97F4A7C15ULL * (b + 1);
  return mix(h);
}
