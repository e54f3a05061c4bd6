#include <immintrin.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <thread>
#include <vector>
#include <cstring>
static void pack(const uint8_t* src, uint32_t* dst, size_t n) { // n % 32 == 0
  for (size_t i = 0; i < n; i += 32) {
    __m256i v = _mm256_loadu_si256((const __m256i*)(src + i));
    __m256i z = _mm256_cmpeq_epi8(v, _mm256_setzero_si256());
    dst[i >> 5] = ~(uint32_t)_mm256_movemask_epi8(z);
  }
}
static void unpack(const uint32_t* src, uint8_t* dst, size_t n) {
  const __m256i shuf = _mm256_setr_epi8(0,0,0,0,0,0,0,0,1,1,1,1,1,1,1,1,2,2,2,2,2,2,2,2,3,3,3,3,3,3,3,3);
  const __m256i bit = _mm256_set1_epi64x(0x8040201008040201ULL);
  for (size_t i = 0; i < n; i += 32) {
    __m256i v = _mm256_set1_epi32((int)src[i >> 5]);
    v = _mm256_shuffle_epi8(v, shuf);
    v = _mm256_and_si256(v, bit);
    v = _mm256_cmpeq_epi8(v, bit);
    v = _mm256_and_si256(v, _mm256_set1_epi8(1));
    _mm256_stream_si256((__m256i*)(dst + i), v);
  }
}
int main(int argc, char** argv) {
  int T = argc > 1 ? atoi(argv[1]) : 1;
  size_t n = (size_t)1 << 30;
  uint8_t* a = (uint8_t*)aligned_alloc(64, n); uint8_t* c = (uint8_t*)aligned_alloc(64, n);
  uint32_t* b = (uint32_t*)aligned_alloc(64, n / 8);
  for (size_t i = 0; i < n; ++i) a[i] = (i * 2654435761u >> 13) & 1;
  memset(c, 0, n); memset(b, 0, n / 8);
  for (int rep = 0; rep < 2; ++rep) {
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    for (int t = 0; t < T; ++t) th.emplace_back([&, t] { size_t lo = n / T * t & ~(size_t)31, hi = (t == T - 1) ? n : (n / T * (t + 1) & ~(size_t)31); pack(a + lo, b + lo / 32, hi - lo); });
    for (auto& x : th) x.join();
    auto t1 = std::chrono::steady_clock::now();
    th.clear();
    for (int t = 0; t < T; ++t) th.emplace_back([&, t] { size_t lo = n / T * t & ~(size_t)31, hi = (t == T - 1) ? n : (n / T * (t + 1) & ~(size_t)31); unpack(b + lo / 32, c + lo, hi - lo); });
    for (auto& x : th) x.join();
    auto t2 = std::chrono::steady_clock::now();
    double p = std::chrono::duration<double>(t1 - t0).count(), u = std::chrono::duration<double>(t2 - t1).count();
    printf("threads %d: pack %.1f GB/s, unpack %.1f GB/s, equal %d\n", T, n / p / 1e9, n / u / 1e9, memcmp(a, c, n) == 0);
  }
}
