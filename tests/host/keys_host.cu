// Host-side check of the __host__ __device__ key functions in eprecon_b200/csrc/common.cuh (compiled by
// tests/test_host_keys_cpu.py with nvcc, runs on the CPU):
//   mode "order":  for random in-range coordinates, sorting by the compact Z-order key gives the SAME permutation as sorting by
//                  the 64-bit key (the executor's 24 / 32-bit tiers must not change the row order), and unkey round-trips;
//   mode "hash N": prints ep_sphash of N (x, y, z, b) quadruples read from stdin, one per line.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <random>
#include <vector>

#include "common.cuh"

static int check_order(int cb, int bb, int n, unsigned seed) {
  std::mt19937 rng(seed);
  const int h = 1 << (cb - 1);
  std::uniform_int_distribution<int> dc(-h, h - 1), db(0, (1 << bb) - 1);
  std::vector<int> x(n), y(n), z(n), b(n);
  std::vector<uint64_t> wide(n), compact(n);
  for (int i = 0; i < n; ++i) {
    // cluster around the sign change and the range ends, where a wrong bit level would show
    const int r = (int)(rng() % 4);
    auto pick = [&]() { return r == 0 ? (int)(rng() % 5) - 2 : r == 1 ? h - 1 - (int)(rng() % 3) : r == 2 ? -h + (int)(rng() % 3) : dc(rng); };
    x[i] = pick(); y[i] = pick(); z[i] = pick(); b[i] = db(rng);
    if (!ep_morton_compact_ok(x[i], y[i], z[i], b[i], cb, bb)) return 1;
    wide[i] = ep_morton_key(x[i], y[i], z[i], b[i]);
    compact[i] = ep_morton_key_compact(x[i], y[i], z[i], b[i], cb);
    if (compact[i] >> (3 * cb + bb)) return 2;                       // fits the advertised width
    int ux, uy, uz, ub;
    ep_morton_unkey_compact(compact[i], cb, ux, uy, uz, ub);
    if (ux != x[i] || uy != y[i] || uz != z[i] || ub != b[i]) return 3;
    ep_morton_unkey(wide[i], ux, uy, uz, ub);
    if (ux != x[i] || uy != y[i] || uz != z[i] || ub != b[i]) return 4;
  }
  for (int i = 0; i < n; ++i)
    for (int j = i + 1; j < std::min(n, i + 40); ++j) {
      const bool lw = wide[i] < wide[j], lc = compact[i] < compact[j];
      const bool ew = wide[i] == wide[j], ec = compact[i] == compact[j];
      if (lw != lc || ew != ec) return 5;
    }
  std::vector<int> pw(n), pc(n);
  std::iota(pw.begin(), pw.end(), 0);
  pc = pw;
  std::stable_sort(pw.begin(), pw.end(), [&](int a, int c) { return wide[a] < wide[c]; });
  std::stable_sort(pc.begin(), pc.end(), [&](int a, int c) { return compact[a] < compact[c]; });
  if (pw != pc) return 6;
  // out-of-range points are rejected
  if (ep_morton_compact_ok(h, 0, 0, 0, cb, bb) || ep_morton_compact_ok(0, -h - 1, 0, 0, cb, bb) ||
      ep_morton_compact_ok(0, 0, 0, 1 << bb, cb, bb) || ep_morton_compact_ok(0, 0, 0, -1, cb, bb))
    return 7;
  return 0;
}

int main(int argc, char** argv) {
  if (argc >= 2 && !strcmp(argv[1], "order")) {
    const int tiers[][2] = {{8, 0}, {9, 5}, {7, 3}, {12, 4}, {16, 8}};
    for (auto& t : tiers) {
      const int rc = check_order(t[0], t[1], 20000, 17u + t[0]);
      if (rc) { printf("FAIL cb=%d bb=%d code=%d\n", t[0], t[1], rc); return 1; }
    }
    printf("OK\n");
    return 0;
  }
  if (argc >= 3 && !strcmp(argv[1], "hash")) {
    const int n = atoi(argv[2]);
    for (int i = 0; i < n; ++i) {
      int x, y, z, b;
      if (scanf("%d %d %d %d", &x, &y, &z, &b) != 4) return 2;
      printf("%llu\n", (unsigned long long)ep_sphash(x, y, z, b));
    }
    return 0;
  }
  return 3;
}
