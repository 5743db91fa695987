// bitplane.cuh -- helpers for the 64-cells-per-word bit-planes shared by the labelling, entrance and light passes
#pragma once
#include "common.cuh"

namespace tsim {

typedef unsigned long long u64;

// In-register transpose of a 64 x 64 bit block held by a warp as rows `lane` (a0) and `lane + 32` (a1): six
// butterfly stages (Hacker's Delight), the five cross-lane ones with one 64-bit shuffle per word.
__device__ __forceinline__ void t64(u64 &a0, u64 &a1, int lane) {
    { const u64 t = ((a0 >> 32) ^ a1) & 0x00000000ffffffffull; a0 ^= t << 32; a1 ^= t; }
#pragma unroll
    for (int j = 16; j >= 1; j >>= 1) {
        const u64 m = j == 16 ? 0x0000ffff0000ffffull : j == 8 ? 0x00ff00ff00ff00ffull : j == 4 ? 0x0f0f0f0f0f0f0f0full
                    : j == 2 ? 0x3333333333333333ull : 0x5555555555555555ull;
        const u64 p0 = __shfl_xor_sync(0xffffffffu, a0, j), p1 = __shfl_xor_sync(0xffffffffu, a1, j);
        if ((lane & j) == 0) { a0 ^= (((a0 >> j) ^ p0) & m) << j; a1 ^= (((a1 >> j) ^ p1) & m) << j; }
        else { a0 ^= ((p0 >> j) ^ a0) & m; a1 ^= ((p1 >> j) ^ a1) & m; }
    }
}

// rows by*64 .. by*64+63 of word column bx of `src` ([src_rows][src_wp]) become words `by` of rows bx*64 .. bx*64+63
// of `dst` ([dst_rows][dst_wp])
__device__ __forceinline__ void transpose_block(const u64 *__restrict__ src, int src_rows, int src_wp, u64 *__restrict__ dst, int dst_rows, int dst_wp,
                                                int bx, int by, int lane, bool coherent) {
    const int r0 = by * 64 + lane, r1 = r0 + 32;
    const u64 *p0 = src + (size_t)r0 * src_wp + bx, *p1 = src + (size_t)r1 * src_wp + bx;
    u64 a0 = r0 < src_rows ? (coherent ? __ldcg(p0) : *p0) : 0ull, a1 = r1 < src_rows ? (coherent ? __ldcg(p1) : *p1) : 0ull;
    t64(a0, a1, lane);
    const int d0 = bx * 64 + lane, d1 = d0 + 32;
    if (d0 < dst_rows) dst[(size_t)d0 * dst_wp + by] = a0;
    if (d1 < dst_rows) dst[(size_t)d1 * dst_wp + by] = a1;
}


// Transpose of a 256-row x 4-word tile of TWO planes by a CTA of 256 threads through shared memory, so that both
// global sides move whole 32-byte sectors (a lone 64 x 64 block moves 8 useful bytes per sector) and the loads of both
// planes are in flight together.  Tile (tx, ty) = rows ty*256 .. +255, words tx*4 .. +3 of the sources; it lands in
// rows tx*256 .. +255, words ty*4 .. +3 of the destinations.  s_in / s_out: [2][256][5] words.
// `compare`: read the destination first and return (uniformly over the CTA) whether a word changed; otherwise store
// unconditionally and return true.
__device__ __forceinline__ bool transpose_tile2(const u64 *__restrict__ srcA, const u64 *__restrict__ srcB, int src_rows, int src_wp,
                                                u64 *__restrict__ dstA, u64 *__restrict__ dstB, int dst_rows, int dst_wp, int tx, int ty, bool coherent,
                                                bool compare, u64 (*s_in)[256][5], u64 (*s_out)[256][5]) {
    const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
    {
        const int r = ty * 256 + t;
        const size_t o = (size_t)r * src_wp + tx * 4;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const bool ok = r < src_rows && tx * 4 + k < src_wp;
            s_in[0][t][k] = ok ? (coherent ? __ldcg(srcA + o + k) : srcA[o + k]) : 0ull;
            s_in[1][t][k] = ok ? (coherent ? __ldcg(srcB + o + k) : srcB[o + k]) : 0ull;
        }
    }
    __syncthreads();
#pragma unroll
    for (int b = wid; b < 32; b += 8) {
        const int pl = b >> 4, i = (b >> 2) & 3, j = b & 3;
        u64 a0 = s_in[pl][64 * i + lane][j], a1 = s_in[pl][64 * i + 32 + lane][j];
        t64(a0, a1, lane);
        s_out[pl][64 * j + lane][i] = a0;
        s_out[pl][64 * j + 32 + lane][i] = a1;
    }
    __syncthreads();
    if (!compare) {
        const int r = tx * 256 + t;
        const size_t o = (size_t)r * dst_wp + ty * 4;
        if (r < dst_rows) {
#pragma unroll
            for (int k = 0; k < 4; k++)
                if (ty * 4 + k < dst_wp) { dstA[o + k] = s_out[0][t][k]; dstB[o + k] = s_out[1][t][k]; }
        }
        __syncthreads();
        return true;
    }
    bool ch = false;
    {
        const int r = tx * 256 + t;
        const size_t o = (size_t)r * dst_wp + ty * 4;
        if (r < dst_rows) {
#pragma unroll
            for (int k = 0; k < 4; k++)
                if (ty * 4 + k < dst_wp) {
                    const u64 a = s_out[0][t][k], b = s_out[1][t][k];
                    if (__ldcg(dstA + o + k) != a || __ldcg(dstB + o + k) != b) { dstA[o + k] = a; dstB[o + k] = b; ch = true; }
                }
        }
    }
    return __syncthreads_or(ch);
}

// Transpose of a 512-row x 8-word tile of one plane by a CTA of 256 threads through shared memory.  Every row of the
// tile is 64 contiguous bytes on BOTH global sides (the tile lands as 512 rows x 8 words), and four lanes cover a row
// with 16-byte accesses, so a warp instruction moves 8 rows x 64 B: 8x fewer memory requests than word-wise access
// (the 8-byte version was bound by the L2 request rate, profiles/).  Tile (tx, ty) = rows ty*512.., words tx*8.. of
// `src`; it lands in rows tx*512.., words ty*8.. of `dst`.  s_in / s_out: [512][9] words each.
// compare: read the destination first; returns (uniformly) whether a word changed, else stores blindly and returns true.
// The three steps are separate so that a kernel that transposes two planes can have the second plane's loads in flight while
// it works on the first (the profile of the one-plane-after-the-other version had its warps waiting on the loads: each thread's
// 8 loads of 16 bytes went out four at a time, and nothing overlapped them).
struct Tile512Regs { u64 v[16]; };

__device__ __forceinline__ void tile512_load(const u64 *__restrict__ src, int src_rows, int src_wp, int tx, int ty, bool coherent, Tile512Regs &q) {
    const int t = threadIdx.x;
    const bool vec_in = (src_wp & 1) == 0;
#pragma unroll
    for (int it = 0; it < 8; it++) {
        const int idx = it * 256 + t, lr = idx >> 2, part = idx & 3;   // local row, 16-byte part of its 64 bytes
        const int r = ty * 512 + lr, w = tx * 8 + part * 2;
        u64 a = 0ull, b = 0ull;
        if (r < src_rows) {
            const u64 *p = src + (size_t)r * src_wp + w;
            if (vec_in && w + 1 < src_wp) {
                const ulonglong2 v = coherent ? __ldcg(reinterpret_cast<const ulonglong2 *>(p)) : *reinterpret_cast<const ulonglong2 *>(p);
                a = v.x; b = v.y;
            } else {
                if (w < src_wp) a = coherent ? __ldcg(p) : p[0];
                if (w + 1 < src_wp) b = coherent ? __ldcg(p + 1) : p[1];
            }
        }
        q.v[2 * it] = a; q.v[2 * it + 1] = b;
    }
}

__device__ __forceinline__ void tile512_stage(const Tile512Regs &q, u64 (*s_in)[9]) {
    const int t = threadIdx.x;
#pragma unroll
    for (int it = 0; it < 8; it++) {
        const int idx = it * 256 + t, lr = idx >> 2, part = idx & 3;
        s_in[lr][part * 2] = q.v[2 * it]; s_in[lr][part * 2 + 1] = q.v[2 * it + 1];
    }
}

// s_in (staged, after a __syncthreads) -> transposed -> dst.  compare: read the destination first; returns (uniformly)
// whether a word changed, else stores blindly and returns true.
__device__ __forceinline__ bool tile512_finish(u64 *__restrict__ dst, int dst_rows, int dst_wp, int tx, int ty, bool compare, u64 (*s_in)[9],
                                               u64 (*s_out)[9]) {
    const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
    const bool vec_out = (dst_wp & 1) == 0;
#pragma unroll 2
    for (int blk = wid; blk < 64; blk += 8) {
        const int i = blk >> 3, j = blk & 7;   // 64-row block i, word j of the tile
        u64 a0 = s_in[64 * i + lane][j], a1 = s_in[64 * i + 32 + lane][j];
        t64(a0, a1, lane);
        s_out[64 * j + lane][i] = a0;
        s_out[64 * j + 32 + lane][i] = a1;
    }
    __syncthreads();
    bool ch = !compare;
    if (compare) {   // all destination words of this thread in flight together, then the differing ones are rewritten
        u64 old[16];
#pragma unroll
        for (int it = 0; it < 8; it++) {
            const int idx = it * 256 + t, lr = idx >> 2, part = idx & 3;
            const int r = tx * 512 + lr, w = ty * 8 + part * 2;
            const u64 *p = dst + (size_t)r * dst_wp + w;
            old[2 * it] = (r < dst_rows && w < dst_wp) ? __ldcg(p) : 0ull;
            old[2 * it + 1] = (r < dst_rows && w + 1 < dst_wp) ? __ldcg(p + 1) : 0ull;
        }
#pragma unroll
        for (int it = 0; it < 8; it++) {
            const int idx = it * 256 + t, lr = idx >> 2, part = idx & 3;
            const int r = tx * 512 + lr, w = ty * 8 + part * 2;
            if (r >= dst_rows || w >= dst_wp) continue;
            const u64 a = s_out[lr][part * 2], b = s_out[lr][part * 2 + 1];
            const bool two = w + 1 < dst_wp;
            if (old[2 * it] == a && (!two || old[2 * it + 1] == b)) continue;
            ch = true;
            u64 *p = dst + (size_t)r * dst_wp + w;
            if (vec_out && two) *reinterpret_cast<ulonglong2 *>(p) = make_ulonglong2(a, b);
            else { p[0] = a; if (two) p[1] = b; }
        }
    } else {
#pragma unroll 4
        for (int it = 0; it < 8; it++) {
            const int idx = it * 256 + t, lr = idx >> 2, part = idx & 3;
            const int r = tx * 512 + lr, w = ty * 8 + part * 2;
            if (r >= dst_rows || w >= dst_wp) continue;
            const u64 a = s_out[lr][part * 2], b = s_out[lr][part * 2 + 1];
            u64 *p = dst + (size_t)r * dst_wp + w;
            const bool two = w + 1 < dst_wp;
            if (vec_out && two) *reinterpret_cast<ulonglong2 *>(p) = make_ulonglong2(a, b);
            else { p[0] = a; if (two) p[1] = b; }
        }
    }
    return __syncthreads_or(ch);   // also: s_in / s_out may be reused
}

// 16-bit membership mask of a 16-byte strip: bit k = type byte k is in `set` (bit t of `set` = type t)
__device__ __forceinline__ uint32_t strip_set_mask(const uint4 &q, uint32_t set) {
    uint32_t m = 0;
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int k = 0; k < 4; k++)
#pragma unroll
        for (int b = 0; b < 4; b++) m |= ((set >> ((w[k] >> (8 * b)) & 31u)) & 1u) << (4 * k + b);
    return m;
}

// SWAR range test (Bit Twiddling Hacks "hasbetween"): high bit of every byte b of x with lo < b < hi (bytes >= 128 never match)
__device__ __forceinline__ uint32_t bytes_between(uint32_t x, uint32_t lo, uint32_t hi) {
    const uint32_t a = x & 0x7f7f7f7fu;
    return (0x01010101u * (127u + hi) - a) & ~x & (a + 0x01010101u * (127u - lo)) & 0x80808080u;
}
__device__ __forceinline__ uint32_t collapse4(uint32_t hibits) { return ((hibits >> 7) * 0x01020408u) >> 24; }   // byte k's flag -> bit k

// a set of cell types as at most two exclusive byte ranges (lo < t < hi); lo2 == hi2 == 0: one range only
struct TypeRanges { uint32_t lo1, hi1, lo2, hi2; };
__device__ __forceinline__ uint32_t strip_range_mask(const uint32_t (&w)[4], const TypeRanges r) {
    uint32_t m = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        uint32_t h = bytes_between(w[k], r.lo1, r.hi1);
        if (r.hi2) h |= bytes_between(w[k], r.lo2, r.hi2);
        m |= collapse4(h) << (4 * k);
    }
    return m;
}

// four lanes of a quad hold the 16-bit masks of four consecutive strips -> the 64-bit word (valid in every lane of the quad)
__device__ __forceinline__ u64 quad_pack(uint32_t m16, int q) {
    u64 v = (u64)m16 << (16 * q);
    v |= __shfl_xor_sync(0xffffffffu, v, 1);
    v |= __shfl_xor_sync(0xffffffffu, v, 2);
    return v;
}

// bits [start, start + len) of a bit-row of `nwords` words (len <= 64); positions outside the row read as 0
__device__ __forceinline__ u64 extract_bits(const u64 *__restrict__ row, int nwords, int start, int len) {
    int shl = 0;
    if (start < 0) { shl = -start; len += start; start = 0; }
    if (len <= 0 || row == nullptr) return 0ull;
    const int w = start >> 6, b = start & 63;
    const u64 lo = w < nwords ? row[w] : 0ull, hi = w + 1 < nwords ? row[w + 1] : 0ull;
    u64 v = b ? (lo >> b) | (hi << (64 - b)) : lo;
    if (len < 64) v &= (1ull << len) - 1ull;
    return v << shl;
}

}  // namespace tsim
