// k_digest.cu -- position-weighted 64-bit digest of a band of plane rows.
//
// Row-band shards (sharded.py, "lean" mode) do not refresh their halos after every pass: a shard computes its halo rows
// itself, and what it misses beyond the window edge can only spoil rows near that edge.  Whether the spoilt zone stayed
// away from the rows next to a cut is CHECKED: both shards digest the rows on either side of the cut after every pass
// and compare once, at the end of the pipeline.  The digest of a row depends on its GLOBAL position, so the two shards
// (whose windows start at different rows) agree exactly when the bytes agree.
#include "common.cuh"

namespace tsim {

typedef unsigned long long u64;

__device__ __forceinline__ u64 mix64(u64 x) {   // splitmix64 finaliser
    x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull;
    x ^= x >> 27; x *= 0x94d049bb133111ebull;
    return x ^ (x >> 31);
}

// bytes [lo, hi) of a plane whose element 0 sits at global byte offset g0; 8 bytes per thread where alignment allows
__device__ u64 digest_bytes(const uint8_t *p, long long lo, long long hi, long long g0, u64 salt, long long tid, long long nth) {
    u64 h = 0;
    if ((((uintptr_t)(p + lo)) & 7) == 0 && ((g0 + lo) & 7) == 0) {
        const long long nw = (hi - lo) >> 3;
        const u64 *w = reinterpret_cast<const u64 *>(p + lo);
        for (long long i = tid; i < nw; i += nth) h += mix64(w[i] ^ mix64((u64)((g0 + lo) / 8 + i) + salt));
        lo += nw << 3;
    }
    for (long long i = lo + tid; i < hi; i += nth) h += mix64((u64)p[i] ^ mix64((u64)(g0 + i) * 8 + 7 + salt));
    return h;
}

__global__ void __launch_bounds__(256) rows_digest_kernel(int W, long long y0_global, int row_lo, int row_hi, tsim_planes p, int what, u64 *out) {
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
    const long long c0 = (long long)row_lo * W, c1 = (long long)row_hi * W, g = y0_global * W;
    u64 h = 0;
    if (what & 1) h += digest_bytes(p.cell_type, c0, c1, g, 0x1111ull << 48, tid, nth);
    if (what & 2) h += digest_bytes(reinterpret_cast<const uint8_t *>(p.dirs), c0 * 2, c1 * 2, g * 2, 0x2222ull << 48, tid, nth);
    if (what & 4) h += digest_bytes(p.aux, c0, c1, g, 0x3333ull << 48, tid, nth);
    if (what & 8) h += digest_bytes(reinterpret_cast<const uint8_t *>(p.block_id), c0 * 4, c1 * 4, g * 4, 0x4444ull << 48, tid, nth);
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) h += __shfl_xor_sync(0xffffffffu, h, o);
    if ((threadIdx.x & 31) == 0 && h) atomicAdd(out, h);
}

}  // namespace tsim

using namespace tsim;

extern "C" tsim_status tsim_rows_digest(const tsim_cfg *cfg, const tsim_planes *p, int32_t row_lo, int32_t row_hi, int32_t what, uint64_t *out,
                                        void *stream) {
    tsim_status st = check_cfg(cfg);
    if (st != TSIM_OK) return st;
    if (!p || !out || row_lo < 0 || row_hi > cfg->win_rows || row_lo > row_hi || (what & ~15) ||
        ((what & 1) && !p->cell_type) || ((what & 2) && !p->dirs) || ((what & 4) && !p->aux) || ((what & 8) && !p->block_id)) {
        set_error("tsim_rows_digest: bad arguments");
        return TSIM_ERR_CONFIG;
    }
    if (row_lo == row_hi || !what) return TSIM_OK;
    const long long cells = (long long)(row_hi - row_lo) * cfg->width;
    long long blocks = (cells / 8 + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks < 1) blocks = 1;
    rows_digest_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(cfg->width, cfg->win_y0, row_lo, row_hi, *p, what, (u64 *)out);
    TSIM_LAUNCH_CHECK();
    return TSIM_OK;
}

// ---- measurement aid (profiles/write_peak.py): what a WRITE-ONLY kernel reaches on this device with data that is not a constant
// (a constant fill runs above the copy peak: the memory system compresses it).  Every thread stores `streams` x 16 bytes of a
// hash of its index per step, to `streams` planes of n bytes each laid out back to back, like the layout passes that write T, A, D.
namespace tsim {
__global__ void __launch_bounds__(256) write_probe_kernel(uint8_t *base, long long n_bytes, int streams) {
    const long long nvec = n_bytes / 16;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
        const u64 h = mix64((u64)i);
        for (int s = 0; s < streams; s++) {
            const u64 g = h + (u64)s * 0x9e3779b97f4a7c15ull;
            reinterpret_cast<uint4 *>(base + (size_t)s * n_bytes)[i] = make_uint4((uint32_t)g, (uint32_t)(g >> 32), (uint32_t)(g >> 13), (uint32_t)(g >> 45));
        }
    }
}
}  // namespace tsim

extern "C" tsim_status tsim_debug_write_probe(void *base, long long n_bytes, int32_t streams, int32_t blocks, void *stream) {
    if (!base || n_bytes < 16 || streams < 1 || streams > 8 || blocks < 1) { set_error("tsim_debug_write_probe: bad arguments"); return TSIM_ERR_CONFIG; }
    write_probe_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((uint8_t *)base, n_bytes, streams);
    TSIM_LAUNCH_CHECK();
    return TSIM_OK;
}
