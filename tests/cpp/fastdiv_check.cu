// Host-side check of make_fastdiv (peclr_b200/csrc/ptx.cuh): the multiply-high + shift the conv kernels use for
// tile -> (w, h, n) must equal integer division for every divisor / dividend a launch can produce.
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#include "../../peclr_b200/csrc/ptx.cuh"

static int host_div(const peclr::FastDiv& f, int x) {  // the arithmetic of fd_div (device: __umulhi)
  return f.d == 1 ? x : (int)((uint32_t)(((uint64_t)(uint32_t)x * f.mul) >> 32) >> f.shr);
}

int main() {
  uint64_t lcg = 12345;
  long checked = 0;
  for (int d = 1; d <= 70000; d = d < 5000 ? d + 1 : d + 97) {
    const peclr::FastDiv f = peclr::make_fastdiv(d);
    const int probes[] = {0, 1, d - 1, d, d + 1, 2 * d - 1, 2 * d, 1000003, 0x7ffffffe, 0x7fffffff};
    for (int x : probes) {
      if (x < 0) continue;
      if (host_div(f, x) != x / d) { printf("FAIL d=%d x=%d got %d want %d\n", d, x, host_div(f, x), x / d); return 1; }
      ++checked;
    }
    for (int i = 0; i < 200; ++i) {
      lcg = lcg * 6364136223846793005ull + 1442695040888963407ull;
      const int x = (int)((lcg >> 33) & 0x7fffffff);
      if (host_div(f, x) != x / d) { printf("FAIL d=%d x=%d got %d want %d\n", d, x, host_div(f, x), x / d); return 1; }
      ++checked;
    }
  }
  printf("ok %ld\n", checked);
  return 0;
}
