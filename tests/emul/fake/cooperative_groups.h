// TEST INFRASTRUCTURE: single-lane stand-in for cooperative_groups tiles (see fake/cuda_runtime.h).
#pragma once
namespace cooperative_groups {
struct thread_block {};
inline thread_block this_thread_block() { return thread_block(); }
template <int G> struct thread_block_tile {
  static_assert(G == 1, "host emulation runs one lane per world");
  int thread_rank() const { return 0; }
  void sync() const {}
  template <class T> T shfl_xor(T v, int) const { return v; }
  template <class T> T shfl_up(T v, int) const { return v; }
  template <class T> T shfl(T v, int) const { return v; }
  unsigned ballot(bool p) const { return p ? 1u : 0u; }
};
template <int G> thread_block_tile<G> tiled_partition(const thread_block&) { return thread_block_tile<G>(); }
}  // namespace cooperative_groups
