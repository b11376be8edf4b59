// Host-only helpers of libfastfp_b200.so (no device code).
//
// fastfp_hash64: a 64-bit content hash of a host buffer, fast enough (memory-bandwidth class, several
// threads for large buffers) that the Python mirror can fingerprint EVERY byte of the caller's
// (Nvecs, Ts, sigmas/TNTs) lists on each call: the device pack is a cache of those arrays, and the
// reference is a pure function of its arguments (fastfp/fastfp.py:52), so an in-place edit of any entry
// must rebuild the pack. (Round 1 sampled 16 points per array; an edit between the samples went unseen.)
#include <atomic>
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

constexpr size_t BLK = (size_t)4 << 20;  // fixed blocks: a hash value does not depend on the thread count

inline uint64_t combine(const uint64_t* part, size_t nblk, size_t n, uint64_t seed) {
  uint64_t h = seed ^ (uint64_t)n;
  for (size_t b = 0; b < nblk; ++b) h = mix(h ^ part[b]) + 0x9E3779B97F4A7C15ULL * (b + 1);
  return mix(h);
}

unsigned pool_size(size_t nitems) {
  unsigned nt = std::thread::hardware_concurrency();
  nt = nt > 32 ? 32 : (nt < 1 ? 1 : nt);
  return nt > nitems ? (unsigned)nitems : nt;
}
}  // namespace

// n buffers in one call: the (buffer, 4 MiB block) pieces of ALL of them are dealt to one set of threads, so a list of
// many medium-sized arrays (one T matrix per pulsar: 3-6 MB each) hashes at memory bandwidth too -- per-array calls ran
// those single-threaded, ~18 ms per C2-sized call, which bounded the step time of the short workloads from the host side.
// out[i] equals fastfp_hash64(ptrs[i], nbytes[i], seeds[i]).
extern "C" int fastfp_hash64_many(const void* const* ptrs, const int64_t* nbytes, int32_t n, const uint64_t* seeds,
                                  uint64_t* out) {
  if (n < 0 || (n > 0 && (!ptrs || !nbytes || !seeds || !out))) return FASTFP_ERR_INVALID;
  std::vector<size_t> first((size_t)n + 1, 0);
  for (int i = 0; i < n; ++i) {
    if (nbytes[i] < 0 || (nbytes[i] > 0 && !ptrs[i])) return FASTFP_ERR_INVALID;
    first[(size_t)i + 1] = first[i] + ((size_t)nbytes[i] + BLK - 1) / BLK;
  }
  const size_t nitems = first[(size_t)n];
  std::vector<uint64_t> part(nitems);
  std::atomic<size_t> next{0};
  auto work = [&]() {
    int i = 0;
    for (;;) {
      const size_t it = next.fetch_add(1, std::memory_order_relaxed);
      if (it >= nitems) return;
      while (first[(size_t)i + 1] <= it) ++i;  // items are handed out in increasing order
      const size_t b = it - first[i], nb = (size_t)nbytes[i], lo = b * BLK;
      part[it] = hash_block(static_cast<const unsigned char*>(ptrs[i]) + lo, lo + BLK <= nb ? BLK : nb - lo, seeds[i] + b);
    }
  };
  const unsigned nt = pool_size(nitems);
  if (nt <= 1 || nitems < 4) {
    work();
  } else {
    std::vector<std::thread> th;
    for (unsigned t = 0; t + 1 < nt; ++t) th.emplace_back(work);
    work();
    for (auto& t : th) t.join();
  }
  for (int i = 0; i < n; ++i)
    out[i] = nbytes[i] > 0 ? combine(part.data() + first[i], first[(size_t)i + 1] - first[i], (size_t)nbytes[i], seeds[i])
                           : mix(seeds[i] ^ 0x51ed270b7a1c2d4fULL);
  return 0;
}

extern "C" uint64_t fastfp_hash64(const void* data, int64_t nbytes, uint64_t seed) {
  if (!data || nbytes <= 0) return mix(seed ^ 0x51ed270b7a1c2d4fULL);
  uint64_t out = 0;
  const void* p = data;
  fastfp_hash64_many(&p, &nbytes, 1, &seed, &out);
  return out;
}
