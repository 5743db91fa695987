// scan.cuh -- small hand-written scan / compaction helpers shared by the labelling and light passes
#pragma once
#include "common.cuh"

namespace tsim {

constexpr int SCAN_TILE = 4096;   // elements per CTA (256 threads x 16)

// bytes of scratch `exclusive_scan_i32` needs for n elements (tile status words + the tile ticket counter)
inline size_t scan_tmp_bytes(long long n) { return (size_t)(n > 0 ? (n + SCAN_TILE - 1) / SCAN_TILE : 0) * 8 + 16; }

// exclusive scan of `n` ints in place, ONE pass over the data (chained scan with decoupled look-back); tmp: scan_tmp_bytes(n),
// 8-byte aligned; total -> *total_out (device).
// n_dev (optional, device): only the first min(n, *n_dev) elements are live (the rest is neither read nor written).
tsim_status exclusive_scan_i32(int32_t *data, long long n, int32_t *tmp, int32_t *total_out, cudaStream_t cs, const int32_t *n_dev = nullptr);

}  // namespace tsim
