// iq_cut.h -- boundary cut on an overlap slab (host side; restates /root/reference/src/graphcut.jl:5-84).
#pragma once
#include <cstdint>
#include <vector>

namespace iqcut {

typedef unsigned __int128 u128;

// One lattice node: residual capacities towards the 6 neighbours (dir = 2*d (+d) / 2*d+1 (-d)) and the
// Boykov-Kolmogorov bookkeeping (FP64 capacities: exactly one cache line).
template <typename T>
struct NodeT {
  T cap[6];
  int stamp;        // origin-check cache (augmentation counter)
  uint8_t tree;     // 0 free, 1 source tree, 2 sink tree
  int8_t par;       // direction towards the parent, or kTerminal / kNone / kOrphan
  uint8_t term;     // 1 = source slice, 2 = sink slice, 0 = interior
  uint8_t valid;    // bit dir set = neighbour in direction dir exists
  uint8_t inq;      // already in the active queue
  uint8_t pad[7];
};
using Node = NodeT<double>;
static_assert(sizeof(Node) == 64, "one node per cache line");

// Scratch space reused across cuts by one host thread.
struct Work {
  std::vector<Node> nodes;
  std::vector<NodeT<u128>> nodes_x;
  std::vector<int> active, orphans, queue;
  std::vector<uint8_t> reach;
  std::vector<double> A, B;      // slab extraction buffers (used by the driver)
  std::vector<uint8_t> keep, cutmask;
};

// keep[u] = 1 iff voxel u is NOT able to reach the sink slice in the residual graph of a maximum flow
// (labels 0/1 of the reference's Boykov-Kolmogorov call, graphcut.jl:73-81).  A, B: column-major slabs.
void graphcut(const double* A, const double* B, const int sz[3], int dim, uint8_t* keep, Work& w);

// The same cut with the FP64 capacities of graphcut.jl:52 taken as exact integers (every double is m * 2^e; scaled by
// 2^-emin they are all integers) and the max-flow run in 128-bit integer arithmetic: no rounding, hence THE "can reach
// the sink" set of those capacities, whatever the max-flow algorithm.  This is what makes the cut well defined on
// integer-valued (categorical) slabs, where (Du+Dv)/eps capacities sit next to O(1) ones, equal-cost cuts abound and
// an FP64 max-flow returns whichever its own rounding favours.  Returns false (keep untouched) when the dynamic range
// of the capacities does not fit 128 bits.
bool graphcut_exact(const double* A, const double* B, const int sz[3], int dim, uint8_t* keep, Work& w);

// true iff every value of both slabs is an integer (the default rule for choosing the exact cut)
bool integer_valued(const double* A, const double* B, int n);

}  // namespace iqcut
