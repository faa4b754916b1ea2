// iq_cut.h -- boundary cut on an overlap slab (host side; restates /root/reference/src/graphcut.jl:5-84).
#pragma once
#include <cstdint>
#include <vector>

namespace iqcut {

// Scratch space reused across cuts by one host thread.
struct Work {
  int sz[3] = {0, 0, 0};
  int dim = -1;
  std::vector<int> nbr;          // [nvox][6] neighbour index or -1; dir = 2*d (+d) / 2*d+1 (-d)
  std::vector<double> cap;       // [nvox][6] residual capacities
  std::vector<uint8_t> tree;     // 0 free, 1 source tree, 2 sink tree
  std::vector<int8_t> par;       // direction towards the parent, or kTerminal / kNone
  std::vector<uint8_t> term;     // 1 = source slice, 2 = sink slice, 0 = interior
  std::vector<int> active, orphans, queue;
  std::vector<uint8_t> inactive_q, reach;
  std::vector<int> stamp;        // origin-check cache
  std::vector<double> A, B;      // slab extraction buffers (used by the driver)
  std::vector<uint8_t> keep, cutmask;
};

// keep[u] = 1 iff voxel u is NOT able to reach the sink slice in the residual graph of a maximum flow
// (labels 0/1 of the reference's Boykov-Kolmogorov call, graphcut.jl:73-81).  A, B: column-major slabs.
void graphcut(const double* A, const double* B, const int sz[3], int dim, uint8_t* keep, Work& w);

}  // namespace iqcut
