// Shared helpers for the stag_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/stag_b200.h"

namespace stag {

void set_error(const char* fmt, ...);
void count_launch();  // every kernel launch of the library is counted (stag_launch_count)

#define STAG_CHECK_ARG(cond, ...)                 \
  do {                                            \
    if (!(cond)) {                                \
      ::stag::set_error(__VA_ARGS__);             \
      return STAG_EINVAL;                         \
    }                                             \
  } while (0)

#define STAG_CUDA(call)                                                              \
  do {                                                                               \
    cudaError_t err__ = (call);                                                      \
    if (err__ != cudaSuccess) {                                                      \
      ::stag::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call,                 \
                        cudaGetErrorString(err__));                                  \
      return STAG_ECUDA;                                                             \
    }                                                                                \
  } while (0)

#define STAG_LAUNCH_CHECK()                                                          \
  do {                                                                               \
    ::stag::count_launch();                                                          \
    cudaError_t err__ = cudaGetLastError();                                          \
    if (err__ != cudaSuccess) {                                                      \
      ::stag::set_error("%s:%d kernel launch -> %s", __FILE__, __LINE__,             \
                        cudaGetErrorString(err__));                                  \
      return STAG_ECUDA;                                                             \
    }                                                                                \
  } while (0)

constexpr int kHubThreshold = 128;  // rows with more stored edges than this are split
constexpr int kHubSegment = 128;    // edges per hub segment
constexpr int kRangeEdges = 64;     // row ranges (stream items) start a new item every this many stored edges ...
constexpr int kRangeEdgesSmall = 16;        // ... or every 16 on small graphs, where a launch has few items and the kernels are
constexpr int64_t kSmallGraphEdges = 65536; // bound by the length of one item's edge walk rather than by throughput
__host__ __device__ inline int range_edges_for(int64_t num_edges) {
  return num_edges <= kSmallGraphEdges ? kRangeEdgesSmall : kRangeEdges;
}
constexpr int kRangeRows = 256;     // ... and at least every this many rows

int num_sms();  // SM count of the current device (cached per device)

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

}  // namespace stag
