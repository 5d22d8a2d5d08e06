// Device-wide exclusive scan of an int32 array (three launches: tile scan, tile-sum scan, add).
// Shared by the graph builder and the samplers.  Deterministic; in place.
#pragma once

#include "common.cuh"

namespace agnn {

constexpr int kScanThreads = 256;
constexpr int kScanTile = 4096;

// workspace: ceil(n / kScanTile) int32 (tile sums).  data[i] <- sum_{j<i} data[j]; to obtain the total
// scan n + 1 entries with data[n] = 0.
size_t scan_workspace_bytes(int64_t n);
int exclusive_scan_i32(int32_t* data, int64_t n, int32_t* tile_sums, cudaStream_t stream);

}  // namespace agnn
