// Host-side enumeration of mlp_pair_kernel's tile list (ttl_mlp::Sched): prints one line per visited tile
// as "cluster position layer m_blk n_blk" for the geometry given on the command line.
//   mlp_schedule_check n_layers n_m group_m n_clusters nn0 nn1 nn2 nn3
#include <cstdio>
#include <cstdlib>

#include "ttl_mlp.cuh"

int main(int argc, char** argv) {
  if (argc < 9) return 2;
  ttl_mlp::Sched base;
  base.n_layers = atoi(argv[1]);
  base.n_m = atoi(argv[2]);
  base.group_m = atoi(argv[3]);
  const int n_clusters = atoi(argv[4]);
  base.nn0 = atoi(argv[5]); base.nn1 = atoi(argv[6]); base.nn2 = atoi(argv[7]); base.nn3 = atoi(argv[8]);
  base.n_groups = (base.n_m + base.group_m - 1) / base.group_m;
  for (int c = 0; c < n_clusters; ++c) {
    ttl_mlp::Sched sc = base;
    int pos = 0;
    for (bool ok = sc.start(c); ok; ok = sc.advance(n_clusters), ++pos) {
      int l, m, n;
      sc.tile(l, m, n);
      printf("%d %d %d %d %d\n", c, pos, l, m, n);
    }
  }
  return 0;
}
