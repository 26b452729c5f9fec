// brute-force spot check of the claim: (float)((double)x / (double)h) == x / h and (float)((double)x*(double)h) == x*h for floats x, h
#include <stdio.h>
#include <stdint.h>
#include <string.h>
#include <stdlib.h>
int main() {
  uint64_t bad = 0, n = 0;
  uint32_t seed = 12345;
  for (int hi = 0; hi < 2000; ++hi) {
    seed = seed * 1664525u + 1013904223u;
    float h = 0.01f + (seed >> 8) * (10.0f / 16777216.0f);
    for (uint32_t k = 0; k < 2000000; ++k) {
      seed = seed * 1664525u + 1013904223u;
      uint32_t bits = (seed & 0x007fffffu) | ((uint32_t)(100 + (seed >> 27)) << 23) | (seed & 0x80000000u);
      float x; memcpy(&x, &bits, 4);
      volatile float a = (float)((double)x / (double)h), b = x / h;
      volatile float c = (float)((double)x * (double)h), d = x * h;
      if (a != b || c != d) ++bad;
      ++n;
    }
  }
  printf("%llu pairs, %llu mismatches\n", (unsigned long long)n, (unsigned long long)bad);
  return bad != 0;
}
