// Host check (no GPU needed): the run-invariant partial evaluation of Philox4x32-10 used by the fused kernel
// (PhiloxPairGen) yields the same words as ten plain rounds (PairWords::full == philox4x32_10) for random
// keys, lanes, pairs and chain ids.  Built and run by tests/test_host_logic.py.
#include <cstdio>
#include <cstdlib>
#include <random>
#include "../../rwm_pt_pytorch_b200/csrc/rwmpt_kernel.cuh"
using namespace rwmpt;
int main() {
  std::mt19937_64 g(12345);
  long bad = 0, n = 0;
  for (int it = 0; it < 20000; ++it) {
    uint32_t rk[20];
    const uint32_t k0 = (uint32_t)g(), k1 = (uint32_t)g();
    for (int q = 0; q < 10; ++q) { rk[2 * q] = k0 + q * 0x9E3779B9u; rk[2 * q + 1] = k1 + q * 0xBB67AE85u; }
    const int sub = (int)(g() % 32);
    unsigned long long pair = g() >> (it % 3 == 0 ? 20 : 40);
    const unsigned long long gid = (it % 5 == 0) ? g() : (g() & 0xfffff);
    PhiloxPairGen<3> gen;
    gen.init(rk, sub, pair, gid);
    for (int j = 0; j < 4; ++j, ++pair) {
      if ((uint32_t)(pair >> 32) != gen.hi32) gen.init(rk, sub, pair, gid);
      PairWords<5, RWMPT_P_NORMAL> pw;
      pw.full(rk, sub, pair, gid);
      uint32_t w[12];
      gen.gen(rk, (uint32_t)pair, w);
      for (int k = 0; k < 3; ++k) {
        const uint4 r = philox4x32_10(pair_c0(pair, sub, k), (uint32_t)pair, (uint32_t)gid, (uint32_t)(gid >> 32), k0, k1);
        const uint32_t ref[4] = {r.x, r.y, r.z, r.w};
        for (int q = 0; q < 4; ++q, ++n) bad += (w[4 * k + q] != ref[q]) + (pw.w[4 * k + q] != ref[q]);
      }
    }
  }
  std::printf("checked %ld words, %ld mismatches\n", n, bad);
  return bad != 0;
}
