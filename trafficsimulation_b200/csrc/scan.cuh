// scan.cuh -- small hand-written scan / compaction helpers shared by the labelling and light passes
#pragma once
#include "common.cuh"

namespace tsim {

constexpr int SCAN_TILE = 2048;   // elements per CTA (256 threads x 8)

// exclusive scan of `n` ints in place (tmp: >= div_up(n, SCAN_TILE) ints); total -> *total_out (device).
// n_dev (optional, device): only the first min(n, *n_dev) elements are live; tiles beyond are skipped.
tsim_status exclusive_scan_i32(int32_t *data, long long n, int32_t *tmp, int32_t *total_out, cudaStream_t cs, const int32_t *n_dev = nullptr);

// single-CTA exclusive scan of tile counts in place, total -> *n_out
__global__ void scan_tiles_kernel(int ntiles, int32_t *tile_count, int32_t *n_out);

}  // namespace tsim
