// hx_kernels.cu — hand-written sm_100a kernels of the HyperEx primer-search hot path.
//
//   hx_tile_build_kernel / hx_bucket_scan_kernel / hx_tile_scatter_kernel
//       records -> tiles (whole short record, or overlapped window of a long one), ordered
//       longest-first (CTA-private counting sort) so the lanes of a warp finish together.
//   hx_classify_kernel / hx_classify_iupac_kernel
//       utils.rs:207-228 sequence_type: SWAR search for U / T over every record (HBM-bound), then the
//       IUPAC-set check for the few records that contain neither.
//   hx_myers_pair_kernel<W, NP, TMA, KWM>
//       utils.rs:301-351: the Myers Pv/Mv/score recurrence of rust-bio Myers<u64>::find_all_lazy for
//       ALL patterns of a group per text byte, fused with the reference's hit-selection rule evaluated
//       as a streaming arg-min (no hit lists are materialised).  TMA selects the text path (vector
//       loads / TMA bulk copies), KWM >= 0 swaps Myers for Wu-Manber automata when -m is 0..2.
//   hx_combine_kernel / hx_combine_long_kernel
//       join the per-window summaries of a multi-window record into hx_result.
//   hx_find_kernel / hx_scan_u64_kernel
//       known-answer mirror of find_all_lazy (count / scan / ordered write).
//
// Integer bit-parallel work: no tensor cores by design (BASELINE.json north_star).
#include "hx_device.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>

// ------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ int hx_alpha_of(uint32_t raw) {
    // utils.rs:211-227: 0 dna, 1 rna, 2 none
    const bool u = raw & HX_RAW_U, t = raw & HX_RAW_T;
    if (u && !t) return 1;
    if (t && !u) return 0;
    if (!u && !t) return (raw & HX_RAW_NONIUPAC) ? 2 : 0;
    return 2;
}

__device__ __forceinline__ uint32_t hx_w32(const uint4 &v, int i) {
    return i == 0 ? v.x : i == 1 ? v.y : i == 2 ? v.z : v.w;
}

// a & b & c as ONE LOP3 whose shape the compiler keeps (plain C is re-associated into long chains)
__device__ __forceinline__ uint32_t hx_and3(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0x80;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

__device__ __forceinline__ uint4 hx_ldg16(const uint8_t *p) {
    return __ldg(reinterpret_cast<const uint4 *>(p));
}

// 16-byte load of record text.  HX_LDPOL selects the cache policy (A/B measured with ncu at the bench size,
// profiles/r01_text_path_ab.md): 0 default (L1 + L2 normal), 1 no L1 allocation, 2 no L1 allocation +
// L2 evict-first (harmful: the second 16-byte half of a sector is re-fetched from DRAM).
#ifndef HX_LDPOL
#define HX_LDPOL 0
#endif
__device__ __forceinline__ uint64_t hx_policy_evict_first() {
    uint64_t pol = 0;
#if HX_LDPOL >= 2 || HX_CHUNK32
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
#endif
    return pol;
}
__device__ __forceinline__ uint4 hx_ldg16_stream(const uint8_t *p, uint64_t pol) {
#if HX_LDPOL == 2
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0, %1, %2, %3}, [%4], %5;"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(p), "l"(pol));
    return v;
#elif HX_LDPOL == 3  // text leaves the L2 first (it is read once), but still allocates in the L1 (two 16-byte loads per sector)
    uint4 v;
    asm volatile("ld.global.nc.L2::cache_hint.v4.u32 {%0, %1, %2, %3}, [%4], %5;"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(p), "l"(pol));
    return v;
#elif HX_LDPOL == 4  // ... and stays in the L1 until its second half has been read
    uint4 v;
    asm volatile("ld.global.nc.L1::evict_last.L2::cache_hint.v4.u32 {%0, %1, %2, %3}, [%4], %5;"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(p), "l"(pol));
    return v;
#elif HX_LDPOL == 1
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(p));
    (void)pol;
    return v;
#else
    (void)pol;
    return hx_ldg16(p);
#endif
}

// One whole 32-byte sector per thread: LDG.E.256, no L1 allocation (every byte of the sector is in registers after
// this one load), first to leave the L2 (the scan reads text exactly once).  p must be 32-byte aligned.
__device__ __forceinline__ void hx_ldg32_stream(const uint8_t *p, uint64_t pol, uint4 &lo, uint4 &hi) {
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v8.u32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8], %9;"
                 : "=r"(lo.x), "=r"(lo.y), "=r"(lo.z), "=r"(lo.w), "=r"(hi.x), "=r"(hi.y), "=r"(hi.z), "=r"(hi.w)
                 : "l"(p), "l"(pol));
}

// ------------------------------------------------------------------------------------
// tile construction
// ------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t hx_bucket_of(uint32_t span, uint32_t gran) {
    uint32_t b = span / gran;
    return b < HX_NBUCKETS ? b : HX_NBUCKETS - 1;
}

__device__ __forceinline__ void hx_emit_tile(const HxBuildArgs &a, uint32_t *sh_hist, uint32_t r, uint32_t len, uint32_t tl,
                                             uint32_t base, uint32_t i) {
    HxTile t;
    t.rec = r;
    t.own_start = i * tl;
    t.own_len = min(tl, len - t.own_start);
    t.warm = min(a.warm, t.own_start);
    a.tiles[base + i] = t;
    atomicAdd(&sh_hist[hx_bucket_of(t.warm + t.own_len, a.bucket_gran)], 1u);  // CTA-private, flushed once
}

// One thread per record decides the tiling and reserves a contiguous tile range with one atomicAdd (no scan).
// Short records are written by their own thread; a record with many windows (a contig) is written by the
// whole warp, and is also appended to the long-record list consumed by hx_combine_long_kernel.
__global__ void __launch_bounds__(256) hx_tile_build_kernel(const HxBuildArgs a) {
    __shared__ uint32_t sh_hist[HX_NBUCKETS];
    for (int i = threadIdx.x; i < HX_NBUCKETS; i += blockDim.x) sh_hist[i] = 0;
    __syncthreads();
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31u;
    uint32_t len = 0, nt = 0, tl = 0, base = 0;
    bool ok = false;
    if (r < a.n_records) {
        const uint64_t len64 = a.offsets[r + 1] - a.offsets[r];
        if (len64 > 0xFFFFFF00ull) {  // record positions are 32-bit (hx_result.f_start / r_end)
            *a.overflow = 2;
        } else {
            len = (uint32_t)len64;
            if (len == 0) {
                nt = 0;
            } else if (len <= a.single_max) {
                nt = 1;
                tl = len;
            } else {
                nt = (len + a.tile_len - 1) / a.tile_len;
                tl = ((len + nt - 1) / nt + 15u) & ~15u;  // equal-sized windows, 16-byte granular
                nt = (len + tl - 1) / tl;
            }
            base = nt ? atomicAdd(a.ntiles, nt) : 0u;
            if (base + nt > a.tile_cap) {
                *a.overflow = 1;
                nt = 0;
            } else {
                ok = true;
            }
        }
        a.tile_first[r] = ok ? base : 0u;
        a.tile_cnt[r] = ok ? nt : 0u;
    }
    if (ok && nt <= 4u)
        for (uint32_t i = 0; i < nt; ++i) hx_emit_tile(a, sh_hist, r, len, tl, base, i);
    uint32_t longmask = __ballot_sync(0xffffffffu, ok && nt > 4u);
    while (longmask) {
        const int src = __ffs((int)longmask) - 1;
        longmask &= longmask - 1;
        const uint32_t R = __shfl_sync(0xffffffffu, r, src), NT = __shfl_sync(0xffffffffu, nt, src);
        const uint32_t TL = __shfl_sync(0xffffffffu, tl, src), LEN = __shfl_sync(0xffffffffu, len, src);
        const uint32_t BASE = __shfl_sync(0xffffffffu, base, src);
        for (uint32_t i = lane; i < NT; i += 32u) hx_emit_tile(a, sh_hist, R, LEN, TL, BASE, i);
        if (lane == 0 && NT > HX_LONG_TILES) a.long_list[atomicAdd(a.n_long, 1u)] = R;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < HX_NBUCKETS; i += blockDim.x)
        if (sh_hist[i]) atomicAdd(&a.hist[i], sh_hist[i]);
}

// cursor[b] = number of tiles in buckets longer than b (longest bucket is processed first)
__global__ void __launch_bounds__(HX_NBUCKETS) hx_bucket_scan_kernel(const uint32_t *hist, uint32_t *cursor) {
    __shared__ uint32_t warp_sum[32];
    const int i = threadIdx.x;
    const int b = HX_NBUCKETS - 1 - i;
    const uint32_t own = hist[b];
    uint32_t x = own;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
        if ((i & 31) >= d) x += y;
    }
    if ((i & 31) == 31) warp_sum[i >> 5] = x;
    __syncthreads();
    if (i < 32) {
        uint32_t w = warp_sum[i];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t y = __shfl_up_sync(0xffffffffu, w, d);
            if (i >= d) w += y;
        }
        warp_sum[i] = w;
    }
    __syncthreads();
    const uint32_t before = (i >> 5) ? warp_sum[(i >> 5) - 1] : 0u;
    cursor[b] = before + x - own;
}

__global__ void __launch_bounds__(256) hx_tile_scatter_kernel(const HxTile *tiles, const uint32_t *ntiles,
                                                              uint32_t tile_cap, uint32_t gran, uint32_t *cursor,
                                                              uint32_t *order) {
    // counting-sort scatter: per CTA and chunk of 256 tiles, count per bucket in shared memory, reserve one
    // global range per non-empty bucket, then place every tile at range start + its CTA-local rank
    __shared__ uint32_t sh_cnt[HX_NBUCKETS];
    __shared__ uint32_t sh_base[HX_NBUCKETS];
    const uint32_t n = min(*ntiles, tile_cap);
    for (uint32_t c0 = blockIdx.x * blockDim.x; c0 < n; c0 += gridDim.x * blockDim.x) {
        for (int i = threadIdx.x; i < HX_NBUCKETS; i += blockDim.x) sh_cnt[i] = 0;
        __syncthreads();
        const uint32_t t = c0 + threadIdx.x;
        uint32_t b = 0, rank = 0;
        if (t < n) {
            const HxTile tl = tiles[t];
            b = hx_bucket_of(tl.warm + tl.own_len, gran);
            rank = atomicAdd(&sh_cnt[b], 1u);
        }
        __syncthreads();
        for (int i = threadIdx.x; i < HX_NBUCKETS; i += blockDim.x)
            if (sh_cnt[i]) sh_base[i] = atomicAdd(&cursor[i], sh_cnt[i]);
        __syncthreads();
        if (t < n) order[sh_base[b] + rank] = t;
        __syncthreads();
    }
}

int hx_launch_tile_build(const HxBuildArgs &a, uint32_t *cursor, uint32_t *order, cudaStream_t st) {
    if (a.n_records == 0) return 0;
    hx_tile_build_kernel<<<(a.n_records + 255) / 256, 256, 0, st>>>(a);  // whole warps: no early exit inside
    hx_bucket_scan_kernel<<<1, HX_NBUCKETS, 0, st>>>(a.hist, cursor);
    uint32_t blocks = (a.tile_cap + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    hx_tile_scatter_kernel<<<blocks, 256, 0, st>>>(a.tiles, a.ntiles, a.tile_cap, a.bucket_gran, cursor, order);
    return 3;
}

// ------------------------------------------------------------------------------------
// utils.rs:207-228 sequence_type.  Pass 1 (every record, HBM-bound): does the upper-cased text contain 'U'
// or 'T'?  One warp per tile, 16-byte coalesced loads, SWAR on 32-bit words: OR 0x20 folds the case
// (utils.rs:210), XOR with the letter and the zero-byte test (x - 0x01010101) & ~x & 0x80808080 detects it.
// Pass 2 (only records with neither 'U' nor 'T', utils.rs:218-224): is every byte one of "ACGTRYSWKMBDHVN"?
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) hx_classify_kernel(const uint8_t *__restrict__ arena,
                                                          const uint64_t *__restrict__ offsets,
                                                          const uint64_t off_base,
                                                          const HxTile *__restrict__ tiles,
                                                          const uint32_t *__restrict__ ntiles_p, uint32_t tile_cap,
                                                          uint32_t *__restrict__ rec_raw) {
    const uint32_t ntiles = min(*ntiles_p, tile_cap);
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    // the (tile, record offset) pair of the warp's NEXT tile is fetched while the current tile streams
    uint32_t t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    HxTile tl_n = {0u, 0u, 0u, 0u};
    uint64_t off_n = 0;
    if (t < ntiles) {
        tl_n = tiles[t];
        off_n = offsets[tl_n.rec];
    }
    for (; t < ntiles; t += warps) {
        const HxTile tl = tl_n;
        const uint64_t s = off_n - off_base + tl.own_start;
        const uint64_t e = s + tl.own_len;
        if (t + warps < ntiles) {
            tl_n = tiles[t + warps];
            off_n = offsets[tl_n.rec];
        }
        uint32_t zu = 0, zt = 0, f = 0;
        // Interior: whole 16-byte chunks, no per-chunk range test (ncu: the masked first/last chunk and the 64-bit
        // range compares were half of the instructions of a kernel that should be HBM-bound).  Head and tail
        // bytes (< 16 each) are checked one byte per lane.  (Explicit load pipelining — two or four chunks per
        // lane in flight — measured slower.)
        const uint64_t a_lo = (s + 15ull) & ~15ull, a_hi = e & ~15ull;
        if (a_lo <= a_hi) {
            for (uint64_t A = a_lo + lane * 16ull; A < a_hi; A += 512ull) {
                const uint4 v = hx_ldg16(arena + A);
#pragma unroll
                for (int wi = 0; wi < 4; ++wi) {
                    const uint32_t lw = hx_w32(v, wi) | 0x20202020u;
                    const uint32_t xu = lw ^ 0x75757575u, xt = lw ^ 0x74747474u;
                    zu |= (xu - 0x01010101u) & ~xu;
                    zt |= (xt - 0x01010101u) & ~xt;
                }
            }
            const uint64_t p = lane < 16u ? s + lane : a_hi + (lane - 16u);
            if (lane < 16u ? p < a_lo : p < e) {
                const uint32_t c = arena[p] | 0x20u;
                f |= (c == 'u' ? HX_RAW_U : 0u) | (c == 't' ? HX_RAW_T : 0u);
            }
        } else if (s + lane < e) {  // the whole tile lies inside one 16-byte chunk
            const uint32_t c = arena[s + lane] | 0x20u;
            f |= (c == 'u' ? HX_RAW_U : 0u) | (c == 't' ? HX_RAW_T : 0u);
        }
        f |= ((zu & 0x80808080u) ? HX_RAW_U : 0u) | ((zt & 0x80808080u) ? HX_RAW_T : 0u);
        f = __reduce_or_sync(0xffffffffu, f);
        if (lane == 0 && f) atomicOr(&rec_raw[tl.rec], f);
    }
}

// counts the records that have text but neither 'U' nor 'T' (the only ones pass 2 has to look at)
__global__ void __launch_bounds__(256) hx_count_plain_kernel(const uint32_t *__restrict__ rec_raw,
                                                             const uint32_t *__restrict__ tile_cnt, uint32_t n_records,
                                                             uint32_t *__restrict__ n_plain) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    const bool plain = r < n_records && tile_cnt[r] != 0u && !(rec_raw[r] & (HX_RAW_U | HX_RAW_T));
    const uint32_t m = __ballot_sync(0xffffffffu, plain);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(n_plain, (uint32_t)__popc(m));
}

__global__ void __launch_bounds__(256) hx_classify_iupac_kernel(const uint8_t *__restrict__ arena,
                                                                const uint64_t *__restrict__ offsets,
                                                                const uint64_t off_base,
                                                                const HxTile *__restrict__ tiles,
                                                                const uint32_t *__restrict__ ntiles_p, uint32_t tile_cap,
                                                                const uint32_t *__restrict__ n_plain,
                                                                uint32_t *__restrict__ rec_raw) {
    if (*n_plain == 0u) return;  // every record has a 'U' or a 'T': nothing to check (the usual case)
    __shared__ uint8_t bad[256];  // 1: byte is not in "ACGTRYSWKMBDHVN" after to_ascii_uppercase
    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
        const int u = (i >= 'a' && i <= 'z') ? i - 32 : i;
        const bool iupac = u == 'A' || u == 'C' || u == 'G' || u == 'T' || u == 'R' || u == 'Y' || u == 'S' ||
                           u == 'W' || u == 'K' || u == 'M' || u == 'B' || u == 'D' || u == 'H' || u == 'V' ||
                           u == 'N';
        bad[i] = iupac ? 0 : 1;
    }
    __syncthreads();
    const uint32_t ntiles = min(*ntiles_p, tile_cap);
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; t < ntiles; t += warps) {
        const HxTile tl = tiles[t];
        // The record's U/T evidence is complete (pass 1 finished).  Bits set by this pass cannot disturb the
        // test: HX_RAW_NONIUPAC is not part of it.
        if (rec_raw[tl.rec] & (HX_RAW_U | HX_RAW_T)) continue;
        const uint64_t s = offsets[tl.rec] - off_base + tl.own_start;
        const uint64_t e = s + tl.own_len;
        uint32_t f = 0;
        for (uint64_t A = (s & ~15ull) + lane * 16ull; A < e; A += 512ull) {
            const uint4 v = hx_ldg16(arena + A);
#pragma unroll
            for (int wi = 0; wi < 4; ++wi) {
                const uint32_t w = hx_w32(v, wi);
#pragma unroll
                for (int bi = 0; bi < 4; ++bi) {
                    const uint64_t pos = A + wi * 4 + bi;
                    if (pos >= s && pos < e) f |= bad[(w >> (8 * bi)) & 0xffu];
                }
            }
        }
        f = __reduce_or_sync(0xffffffffu, f);
        if (lane == 0 && f) atomicOr(&rec_raw[tl.rec], HX_RAW_NONIUPAC);
    }
}

int hx_launch_classify(const uint8_t *arena, const uint64_t *offsets, uint64_t off_base, const HxTile *tiles,
                       const uint32_t *ntiles, uint32_t tile_cap, uint32_t *rec_raw, const uint32_t *tile_cnt,
                       uint32_t n_records, uint32_t *n_plain, cudaStream_t st) {
    if (tile_cap == 0) return 0;
    uint64_t blocks = ((uint64_t)tile_cap * 32 + 255) / 256;
    if (blocks > 148 * 8 * 4) blocks = 148 * 8 * 4;
    hx_classify_kernel<<<(uint32_t)blocks, 256, 0, st>>>(arena, offsets, off_base, tiles, ntiles, tile_cap, rec_raw);
    hx_count_plain_kernel<<<(n_records + 255) / 256, 256, 0, st>>>(rec_raw, tile_cnt, n_records, n_plain);
    hx_classify_iupac_kernel<<<(uint32_t)blocks, 256, 0, st>>>(arena, offsets, off_base, tiles, ntiles, tile_cap, n_plain,
                                                               rec_raw);
    return 3;
}

// packed text: the host packer's per-record evidence (HX_REC_* of include/hyperex_b200.h == HX_RAW_*) becomes rec_raw
__global__ void __launch_bounds__(256) hx_flags_kernel(const uint8_t *__restrict__ flags, uint32_t *__restrict__ rec_raw,
                                                       uint32_t n_records) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n_records) rec_raw[r] = flags[r] & (HX_RAW_U | HX_RAW_T | HX_RAW_NONIUPAC);
}

int hx_launch_flags(const uint8_t *flags, uint32_t *rec_raw, uint32_t n_records, cudaStream_t st) {
    if (n_records == 0) return 0;
    hx_flags_kernel<<<(n_records + 255) / 256, 256, 0, st>>>(flags, rec_raw, n_records);
    return 1;
}

// ------------------------------------------------------------------------------------
// The Myers step (rust-bio 1.6.0 Myers::step).  The pattern is
// TOP-ALIGNED in the word: bit (i + BITS - m) belongs to pattern position i, so `bound` is
// the sign bit and the score update needs no per-pattern mask.  The unused low bits keep
// pv = 1, mv = 0 and eq = 0, which makes them feed ph = mh = 0 into pattern bit 0 exactly
// like the `<< 1` of the reference feeds a 0 (free start in the text).
// The score is kept biased: s = dist - (k + 1), so "dist <= k" is the sign bit of s.
// (Measured on B200, profiles/int_peak_r01.json: moving the score update to IMAD.HI or to a carry chain is
// slower than what ptxas emits for the plain C below, so there is a single step variant.)
// ------------------------------------------------------------------------------------
__device__ __forceinline__ void hx_step(const uint32_t eq, uint32_t &pv, uint32_t &mv, int &s) {
    const uint32_t xv = eq | mv;
    const uint32_t xh = (((eq & pv) + pv) ^ pv) | eq;
    uint32_t ph = mv | ~(xh | pv);
    uint32_t mh = pv & xh;
    s += (int)(ph >> 31) - (int)(mh >> 31);  // ptxas: LEA.HI + LEA.HI.SX32 (one ALU op per sign bit)
    ph <<= 1;                                // ptxas: IMAD.IADD x+x on the FMA pipe
    mh <<= 1;
    pv = mh | ~(xv | ph);
    mv = ph & xv;
}

__device__ __forceinline__ void hx_step(const uint64_t eq, uint64_t &pv, uint64_t &mv, int &s) {
    const uint64_t xv = eq | mv;
    const uint64_t xh = (((eq & pv) + pv) ^ pv) | eq;
    uint64_t ph = mv | ~(xh | pv);
    uint64_t mh = pv & xh;
    s += (int)(ph >> 63) - (int)(mh >> 63);
    ph <<= 1;
    mh <<= 1;
    pv = mh | ~(xv | ph);
    mv = ph & xv;
}

template <typename W, int NP>
struct HxRow {
    static constexpr int WB = sizeof(W);
    static constexpr int NV = (NP * WB + 15) / 16;
    uint4 v[NV];
    __device__ __forceinline__ void load(const uint8_t *row) {
        const uint4 *p = reinterpret_cast<const uint4 *>(row);
#pragma unroll
        for (int i = 0; i < NV; ++i) v[i] = p[i];
    }
    // row = 32-bit shared-space address of the Peq row: explicit ld.shared.  With a generic pointer ptxas rebuilt the CTA's
    // shared window base inside the hot loop of the packed -m 0 kernel (S2R SR_CgaCtaId + LEA per trip) and gave the S2R
    // the scoreboard of the text prefetch, so every trip waited for the chunk in flight (ncu: 20 % of all stall samples).
    // Measured per kernel: +4 % for packed -m 0, -1 % for Myers and -3 % for the Wu-Manber levels, which keep the pointer.
    __device__ __forceinline__ void load(uint32_t row) {
#pragma unroll
        for (int i = 0; i < NV; ++i)
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(v[i].x), "=r"(v[i].y), "=r"(v[i].z), "=r"(v[i].w)
                         : "r"(row + 16u * i));
    }
    __device__ __forceinline__ W eq(int p) const {
        if (WB == 4) return (W)hx_w32(v[p >> 2], p & 3);
        const uint4 &q = v[p >> 1];
        return (p & 1) ? (W)(((uint64_t)q.w << 32) | q.z) : (W)(((uint64_t)q.y << 32) | q.x);
    }
};

// ------------------------------------------------------------------------------------
// hx_myers_pair_kernel<W, NP, TMA, KWM>
//   one thread = one tile; per text byte the thread loads one Peq row (all NP patterns of the
//   group, shared memory, 16-byte LDS) and advances NP independent automata held in registers.
//
//   A hit (sign bit of any biased score) is rare.  The lane that sees it only PUSHES a compact
//   event record {position, the NP biased scores as bytes} into its private queue in local
//   memory (a few PRMT + fire-and-forget STL, no long-latency operation while the other 31
//   lanes wait; with the text staged by TMA the L1 is free to hold this state).  Queues are
//   DRAINED warp-synchronously — all lanes together, whenever any lane's queue is nearly full and
//   at the last word of the warp's scan — so the pairing logic below runs with full lanes:
//     bestR[p]  = lexicographic min (dist, end) over all hits of pattern p in the tile
//     bestF[p]  = same, restricted to hits usable as a forward primer (end >= m_p - 1,
//                 utils.rs:334)
//     pair[q]   = arg-min of (f_dist + r_dist, f_end, r_end) over pairs with r_end > f_end
//                 (utils.rs:338-349), evaluated as a stream: at a reverse hit the only forward
//                 candidate that can win is bestF (DESIGN.md §3.2 proves the equivalence).
//   All loops are warp-uniform (trip count = the warp's longest tile; lanes past their own end
//   scan null bytes, whose all-zero Peq row cannot produce an owned hit).
//   The CTA stages the Peq table of ONE alphabet at a time; a CTA that holds both DNA and RNA
//   records (rare) runs the scan twice.
//   Records that consist of a single tile get their final hx_result written here; windows of
//   longer records write summaries for hx_combine_kernel.
// ------------------------------------------------------------------------------------
// The text path is a template switch so that both variants ship, are parity-tested and can be A/B-timed:
//   TMA = false  per-thread 16-byte LDG.128, one chunk ahead (measured 3 % faster on B200: a bulk
//                copy is a warp-uniform instruction, so 32 independent per-lane streams issue as 32
//                serialised UBLKCPs; profiles/r01_text_path_ab.md)
//   TMA = true   text staged into shared memory by TMA bulk copies (cp.async.bulk + mbarrier),
//                two HX_CH-byte stages per thread; keeps the text out of the L1
#ifndef HX_CPASYNC
#define HX_CPASYNC 3  // text through shared memory with cp.async (see CPA in the scan kernel): 0 off, 1 packed -m 0 kernels, 2 all packed kernels, 3 every kernel but the TMA variant
#endif
#ifndef HX_QCAP
#define HX_QCAP 16  // event records per thread queue (local memory; pushes are fire-and-forget stores)
#endif
#ifndef HX_CH
#define HX_CH 128   // bytes per text stage of one thread (two stages in flight)
#endif

// ---- TMA (cp.async.bulk) + mbarrier primitives; SASS: UBLKCP / SYNCS ----
__device__ __forceinline__ uint32_t hx_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void hx_mbar_init(uint32_t mbar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
}
__device__ __forceinline__ void hx_mbar_expect_tx(uint32_t mbar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void hx_bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(mbar)
                 : "memory");
}
__device__ __forceinline__ void hx_mbar_wait(uint32_t mbar, uint32_t parity) {
    uint32_t ok = 0;
    for (uint32_t spin = 0; !ok; ++spin) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok)
                     : "r"(mbar), "r"(parity)
                     : "memory");
        if (spin > (1u << 26)) __trap();  // a lost copy must fail loudly, never hang the GPU
    }
}
__device__ __forceinline__ void hx_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Packed text (NIB = true, hx_run_batch_packed): nibble i of the arena is base i; codes of include/hyperex_b200.h.
// hx_code_byte gives the text byte a code stands for under the record's alphabet (code 4 is 'T' in a DNA record and 'U'
// in an RNA record; code 0, "none of the 16 letters", is the null byte whose Peq rows are empty).
__device__ __forceinline__ uint32_t hx_code_byte(uint32_t code, int rna) {
    // "\0ACGTRYSWKMBDHVN" as four little-endian words
    const uint32_t w = code < 4u ? 0x47434100u : code < 8u ? 0x53595254u : code < 12u ? 0x424d4b57u : 0x4e564844u;
    const uint32_t b = (w >> (8u * (code & 3u))) & 0xffu;
    return (rna && code == 4u) ? (uint32_t)'U' : b;
}
template <bool NIB>
__device__ __forceinline__ uint32_t hx_text_byte(const uint8_t *__restrict__ text, uint64_t i, int rna) {
    if constexpr (!NIB) return text[i];
    return hx_code_byte((text[i >> 1] >> (4u * (uint32_t)(i & 1ull))) & 0xfu, rna);
}

// Suffix filter (LF = true): a primer longer than 32 symbols is scanned as the 32-bit automaton of its LAST 32
// symbols.  An alignment of the whole primer ending at text position j with <= k edits contains an alignment of
// that suffix ending at j with <= k edits, so {j : dist_suffix(j) <= k} is a superset of the primer's hits and a
// 64-bit automaton never has to be stepped over the whole text.  Each candidate is verified here, in the
// warp-synchronous drain: a fresh 64-bit automaton of the whole primer over the m + k text bytes that end at j
// reports the exact distance whenever it is <= k, and never <= k otherwise (the window-restart property of
// DESIGN.md §3.1; at the record start the scan is the full scan).  Returns the unbiased distance at `pos`.
// VW = uint32_t when the whole primer has <= 32 symbols: the upper half of its top-aligned 64-bit Peq word is then the
// top-aligned 32-bit word, and the automaton costs half the instructions.
template <typename VW>
__device__ __forceinline__ VW hx_vrow(const uint64_t *__restrict__ vt, uint32_t byte) {
    if constexpr (sizeof(VW) == 8) return __ldg(vt + byte);
    else return __ldg(reinterpret_cast<const uint32_t *>(vt) + 2u * byte + 1u);
}
template <bool NIB, typename VW>
__device__ __forceinline__ uint32_t hx_verify_long(const uint8_t *__restrict__ text, uint64_t rec0, uint32_t pos,
                                                   const uint64_t *__restrict__ vt, uint32_t m, uint32_t k, int rna,
                                                   VW &pv, VW &mv) {
    // text + rec0 (bytes, or nibbles when NIB) is the record's first symbol; pv / mv leave with the automaton's state
    // after `pos`, so that the caller can continue it to pos + 1
    const uint32_t span = m + k - 1u;
    pv = ~(VW)0;
    mv = 0;
    int s = (int)m;
    for (uint32_t i = pos >= span ? pos - span : 0u; i <= pos; ++i)
        hx_step(hx_vrow<VW>(vt, hx_text_byte<NIB>(text, rec0 + i, rna)), pv, mv, s);
    return (uint32_t)s;
}

// Pattern packing at -m 0 (PK = true): a primer longer than 16 symbols is scanned as the Shift-Or automaton of its LAST
// 16 symbols, so a candidate end `pos` means text[pos - 15 .. pos] matches that suffix exactly.  The primer occurs at
// pos iff its first m - 16 symbols also match text[pos - m + 1 .. pos - 16]: vt (the top-aligned 64-bit Peq table of the
// whole primer) has bit (64 - m + i) set in row b iff symbol i accepts the text byte b.
template <bool NIB, int SFX>
__device__ __forceinline__ bool hx_verify_exact(const uint8_t *__restrict__ text, uint64_t rec0, uint32_t pos,
                                                const uint64_t *__restrict__ vt, uint32_t m, int rna) {
    // SFX: symbols of the scanned suffix (16, or 8 with four automata per word)
    if (pos + 1u < m) return false;  // the primer does not fit before pos
    const uint64_t t0 = rec0 + (pos + 1u - m);
    uint64_t ok = 1;
    for (uint32_t i = 0; i + (uint32_t)SFX < m; ++i) ok &= __ldg(vt + hx_text_byte<NIB>(text, t0 + i, rna)) >> (64u - m + i);
    return ok & 1ull;
}

#ifndef HX_WI_UNROLL
#define HX_WI_UNROLL ((PK && NR == 1) ? 2 : 1)  // see the loop over the four text words of a chunk
#endif
#ifndef HX_PK_MINB
#define HX_PK_MINB 5    // resident CTAs per SM the packed kernels (8 words of state) are compiled for
#endif
template <typename W, int NP, bool TMA, int KWM, bool LF, bool PK = false, bool NIB = false, int PPW = 2>
__global__ void __launch_bounds__(HX_THREADS, ((sizeof(W) * (PK ? NP / PPW : NP) <= 16) ? 8 : ((sizeof(W) * (PK ? NP / PPW : NP) <= 32) ? (PK ? HX_PK_MINB : 5) : 4)) * (128 / HX_THREADS))
    hx_myers_pair_kernel(const HxScanArgs a) {
    // NIB: the arena holds 4-bit codes (hx_run_batch_packed); a 16-byte chunk is 32 symbols and the shared-memory
    // Peq table has 16 rows, gathered from the 256-row table of the record alphabet
    static_assert(!NIB || !TMA, "packed text: vector-load text path only");
    constexpr int SYM_BITS = NIB ? 4 : 8;            // bits per text symbol
    constexpr int SPW = 32 / SYM_BITS;               // symbols per 32-bit word
    // Text chunk = what a thread fetches per load: 16 bytes (LDG.128, allocating in the L1: the second half of the
    // 32-byte sector is an L1 hit or an L2 hit).  HX_CHUNK32=1 builds the Myers kernels on ASCII text with one 256-bit,
    // non-allocating, L2-evict-first load per 32-byte sector instead (hx_ldg32_stream).  Measured (profiles/
    // r02_traffic_experiments.md): 13.9 vs 13.5 ms, and DRAM reads go UP (3.6 vs 2.3 GB per 1.5 GB of text): DRAM is
    // read 64 bytes at a time and an evict-first sector does not survive in the L2 until its thread comes back for it.
#ifndef HX_CHUNK32
#define HX_CHUNK32 0
#endif
    constexpr int NH = (HX_CHUNK32 && KWM < 0 && !NIB) ? 2 : 1;  // 16-byte halves per chunk
    // CPA: the kernels stage their text through shared memory with cp.async (LDGSTS), one 16-byte slot per
    // thread and stage.  Not for bandwidth: a register-destination prefetch holds one of the six scoreboards of a warp for
    // a whole chunk, the hit / drain paths need scoreboards too, and ptxas made an instruction of the hot loop wait for
    // the chunk in flight (ncu: one IMAD.MOV carried 10 % of all stall samples of the four-per-word kernel).  An
    // asynchronous copy owns no scoreboard until cp.async.wait_group asks for it.  Measured per 0.3 Gbases: -m 0 544 ->
    // 518 us; device step on c2: -m 1 186 -> 215 Gbases/s, -m 2 126 -> 130; the Myers kernels gain 0.5-1 % (13.56 -> 13.48 ms
    // at -m 0, 14.81 -> 14.67 at -m 3) and take the same path: text reaches the automata through shared memory in every
    // default kernel (north_star), by per-thread 16-byte asynchronous copies rather than per-thread TMA bulk copies
    // (TMA = true: the tested option, 3 % slower, see above).
    constexpr bool CPA = HX_CPASYNC && (PK || HX_CPASYNC >= 3) && (KWM == 0 || HX_CPASYNC >= 2) && !TMA && NH == 1;
    constexpr int CHB = 16 * NH;                     // bytes per chunk
    constexpr int CPS = HX_CH / CHB;                 // chunks per TMA stage
    constexpr int CSH = (NIB ? 5 : 4) + (NH - 1);    // log2(symbols per chunk)
    constexpr uint32_t SPC = 1u << CSH;              // symbols per chunk
    constexpr int NROWS = NIB ? 16 : 256;
    constexpr bool WM = KWM >= 0;           // small-k fast path: Wu-Manber / Shift-Or automata instead of Myers
    constexpr int NR = WM ? KWM + 1 : 1;    // error levels 0..k
    constexpr int WB = sizeof(W);
    // PK: NP is the (even) number of pattern halves; NW words hold them, two 16-bit automata each
    // PPW: automata ("parts") per word when PK — 2 of <= 16 symbols, or at -m 0 4 of <= 8 symbols
    static_assert(!PK || (KWM >= 0 && sizeof(W) == 4 && NP % PPW == 0 && !TMA), "packing: Shift-Or / Wu-Manber, 32-bit words, whole words");
    static_assert(PPW == 2 || (PPW == 4 && PK && KWM == 0), "four automata per word: Shift-Or (-m 0) only");
    constexpr int PB = 32 / PPW;                                      // bits per part
    constexpr uint32_t PMASK = PPW == 4 ? 0xffu : 0xffffu;            // one part
    constexpr uint32_t FENCE32 = PPW == 4 ? 0xfefefeffu : 0xfffeffffu;  // clears what a left shift moves across a part boundary
    constexpr uint32_t TOPS = PPW == 4 ? 0x80808080u : 0x80008000u;   // the top bit of every part: 0 = hit
    constexpr int NW = PK ? NP / PPW : NP;    // automaton words a thread steps per text byte
    constexpr int STRIDE = hx_row_stride(NW, WB);
    constexpr int TBYTES = 256 * STRIDE;   // one alphabet's table in global memory
    constexpr int SBYTES = NROWS * STRIDE; // ... and what of it is staged in shared memory
    constexpr int SW = (NP + 3) / 4;  // score words of an event: the NP biased scores as bytes
    // words of a queued event record: position + scores; packed automata queue the level bits of their words instead
    // (cheaper to produce on the divergent hit path; the warp-synchronous drain turns them into scores)
    constexpr int RW = PK ? 1 + NW : 1 + SW;
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ HxGroup g;
    // A word of text can add one event per symbol before the queue is looked at again (4 bytes, or 8 nibbles of packed
    // text): the drain starts while a whole word still fits.
    constexpr int QCAP = NIB ? HX_QCAP + 8 : HX_QCAP;
    uint32_t queue[QCAP][RW];  // event records; indexed dynamically => local memory
    // per-thread text pipeline (TMA variant): two HX_CH-byte stages, each with its own mbarrier
    // (count 1): the thread arms the barrier with the byte count, issues one TMA bulk copy
    // global -> shared and later waits for the phase; no cross-thread synchronisation is involved.
    uint8_t *const txt = smem + SBYTES;
    const uint32_t stage0 = hx_smem_u32(txt + (size_t)threadIdx.x * HX_CH);
    const uint32_t stage1 = stage0 + HX_THREADS * HX_CH;
    const uint32_t mbar0 = hx_smem_u32(txt + 2 * HX_THREADS * HX_CH + (size_t)threadIdx.x * 8);
    const uint32_t mbar1 = mbar0 + HX_THREADS * 8;
    uint32_t uses0 = 0, uses1 = 0;  // completed waits per stage => phase parity
    if constexpr (TMA) {
        hx_mbar_init(mbar0, 1);
        hx_mbar_init(mbar1, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        hx_fence_proxy_async();
    }
    if (threadIdx.x < sizeof(HxGroup) / 4)
        reinterpret_cast<uint32_t *>(&g)[threadIdx.x] = reinterpret_cast<const uint32_t *>(a.group)[threadIdx.x];
    __syncthreads();  // g is visible

    // Persistent CTAs: the grid is at most what the device holds at once (hx_scan_launch_one), and every CTA walks
    // batches of HX_THREADS tiles (longest tiles first).  What this buys: a thread's event queue and pairing state (local
    // memory) are the SAME addresses for every tile it scans, so they stay in the L1/L2 instead of being written back to
    // DRAM and fetched again once per tile (ncu, round 1: 0.75 GB each way per 1.5 GB of text), and the Peq table is
    // staged once per CTA, not once per batch.
    //   a.sched != NULL  the next batch comes from a device-wide counter (CTAs drift apart like hardware-scheduled ones;
    //                    the last CTA to leave zeroes the counters for the next launch on the stream)
    //   a.sched == NULL  round r hands batch r * gridDim.x + j to CTA j, in alternating direction
    const uint32_t ntiles = min(*a.ntiles, a.tile_stride);
    const uint32_t nbatches = (ntiles + HX_THREADS - 1u) / HX_THREADS;
    __shared__ uint32_t s_batch;
    uint32_t round = 0;
    const auto next_batch = [&]() -> uint32_t {
        if (a.sched) {
            __syncthreads();  // everybody has read s_batch
            if (threadIdx.x == 0) s_batch = gridDim.x + atomicAdd(a.sched, 1u);
            __syncthreads();
            return s_batch;
        }
        ++round;
        return round * gridDim.x + ((round & 1u) ? gridDim.x - 1u - blockIdx.x : blockIdx.x);
    };
    int staged = -1;  // alphabet whose table is in shared memory
    for (uint32_t batch = blockIdx.x; batch < nbatches; batch = next_batch()) {
    const uint32_t t = batch * HX_THREADS + threadIdx.x;
    const bool have = t < ntiles;
    const uint32_t tile_id = have ? a.order[t] : 0u;
    HxTile tl = {0u, 0u, 0u, 0u};
    int alpha = 2;
    bool single = false;
    if (have) {
        tl = a.tiles[tile_id];
        alpha = hx_alpha_of(a.rec_raw[tl.rec]);  // 2: utils.rs:276-279, record skipped
        single = a.tile_cnt[tl.rec] == 1u;
    }

    W pv[NW], mv[NW];   // Myers state
    int s[NP];          // Myers biased score; for WM: filled on the rare path only
    W rv[NR][NW];       // WM state: bit = 0 <=> prefix matches with <= d errors (Shift-Or convention)
    uint64_t bestF[NP], bestR[NP];
    uint64_t phi[HX_MAX_GROUP_PAIRS];
    uint32_t plo[HX_MAX_GROUP_PAIRS];
    const int bias = (int)a.k + 1;
    if constexpr (PK) {
        // a half starts with its pattern bits (the top m of the 16) set and the unused low bits clear; a half without
        // a pattern (odd pattern count) stays all-ones: its table entries are all-ones too, it never reports
        // (level d starts with its d lowest pattern bits cleared: a prefix of length <= d matches with <= d deletions)
#pragma unroll
        for (int j = 0; j < NW; ++j) {
#pragma unroll
            for (int d = 0; d < NR; ++d) {
                uint32_t w = 0;
#pragma unroll
                for (int h = 0; h < PPW; ++h) {
                    const int p = PPW * j + h, mp = g.pk_m[p];
                    const uint32_t part = p < g.np ? (d < mp ? (PMASK << (PB - mp + d)) & PMASK : 0u) : PMASK;
                    w |= part << (PB * h);
                }
                rv[d][j] = (W)w;
            }
            pv[j] = 0;
            mv[j] = 0;
        }
#pragma unroll
        for (int p = 0; p < NP; ++p) {
            s[p] = 0;
            bestF[p] = HX_NONE64;
            bestR[p] = HX_NONE64;
        }
    }
#pragma unroll
    for (int p = 0; p < (PK ? 0 : NP); ++p) {
        pv[p] = ~(W)0;
        mv[p] = 0;
        s[p] = (int)g.m[p] - bias;
        if constexpr (WM) {
            // pattern bits are the top m bits; level d starts with its d lowest pattern bits cleared (a
            // prefix of length <= d always matches with <= d deletions); the unused low bits stay 0
            const int m = g.m[p];
#pragma unroll
            for (int d = 0; d < NR; ++d) rv[d][p] = d < m ? (~(W)0 << (sizeof(W) * 8 - m + d)) : (W)0;
        }
        bestF[p] = HX_NONE64;
        bestR[p] = HX_NONE64;
    }
    for (int q = 0; q < HX_MAX_GROUP_PAIRS; ++q) {
        phi[q] = HX_NONE64;
        plo[q] = 0;
    }
    const int dz = (int)a.two - 2;  // run-time zero: keeps the pairing state in local memory, not registers

    // (symbol indices; with packed text a symbol is a nibble and the arena address is index / 2)
    const uint64_t rec0 = have ? a.offsets[tl.rec] - a.off_base : 0ull;
    const uint64_t a0 = rec0 + tl.own_start - tl.warm;
    const uint64_t ab = a0 & ~(uint64_t)(SPC - 1u);
    const uint32_t lead = (uint32_t)(a0 - ab);
    const uint32_t pos0 = tl.own_start - tl.warm - lead;  // record position of chunk 0 symbol 0 (may wrap)
    const uint8_t *src = a.arena + (NIB ? ab >> 1 : ab);
    // UB text bytes are processed per trip of the innermost loop (narrow groups amortise the loop
    // overhead over more bytes; wide groups keep the hot loop small).
#ifndef HX_WM_UB24
// text bytes per trip of the Wu-Manber loop when 16 < automata x levels <= 24; measured: 4 is 7 % faster than 2 for
// 8 automata at -m 2 (5.93 -> 5.54 ms per 0.75 Gbases): the level shifts of a byte are reused by the next one
#define HX_WM_UB24 4
#endif
#ifndef HX_PK0_UB
#define HX_PK0_UB 4  // text symbols per trip of the packed -m 0 loop (one hit test per trip; a trip with a hit is replayed)
#endif
#ifndef HX_WM_UBWIDE
#define HX_WM_UBWIDE 2  // ... and beyond 24 (15 automata at -m 1): 2 is 7 % faster than 1 (2.62 -> 2.44 ms per 0.45 Gbases)
#endif
    constexpr int UB = WM ? ((PK && KWM == 0) ? HX_PK0_UB : (KWM == 0 || NW * (KWM + 1) <= 16) ? 4 : (NW * (KWM + 1) <= 24 ? HX_WM_UB24 : HX_WM_UBWIDE))
                          : (NP * WB <= 24) ? 4 : (NP * WB <= 40 ? 2 : 1);
    static_assert(SPW % UB == 0, "text symbols per trip must divide the word");
    constexpr int WI_UNROLL = HX_WI_UNROLL;
    uint32_t cnt = 0;  // records in this lane's queue
    const uint64_t pol = hx_policy_evict_first();

    for (int pass = 0; pass < 2; ++pass) {
        const bool mine = alpha == pass;
        if (!__syncthreads_or(mine ? 1 : 0)) continue;  // no record of this alphabet in the CTA
        if (staged != pass) {  // (every scan of the table in shared memory ended before the barrier above)
            staged = pass;
            const uint4 *tsrc = reinterpret_cast<const uint4 *>(PK ? a.tables_pk : WM ? a.tables_wm : a.tables) + (size_t)pass * (TBYTES / 16);
            uint4 *dst = reinterpret_cast<uint4 *>(smem);
            if constexpr (NIB) {  // row of code c = row of the byte the code stands for in this alphabet
                constexpr int R16 = STRIDE / 16;
                for (int i = threadIdx.x; i < SBYTES / 16; i += HX_THREADS)
                    dst[i] = __ldg(tsrc + (size_t)hx_code_byte((uint32_t)(i / R16), pass) * R16 + i % R16);
            } else {
                for (int i = threadIdx.x; i < TBYTES / 16; i += HX_THREADS) dst[i] = __ldg(tsrc + i);
            }
            __syncthreads();
        }
        // CTA-relative shared-space address of the table, taken from the symbol: converting the generic pointer instead
        // makes ptxas rebuild the cluster-form address (S2R SR_CgaCtaId + LEA) inside the hot loop of the wider kernels
        uint32_t tab_s;
        asm("mov.u32 %0, smem;" : "=r"(tab_s));
        const uint8_t *const tab_g = smem;
        // packed -m 0: shared-space address; everything else: the generic pointer (see HxRow::load)
        const auto tab = [&] {
            if constexpr (PK && NR == 1) return tab_s;
            else return tab_g;
        }();
        const uint32_t own_len = mine ? tl.own_len : 0u;  // a lane that is not scanning owns nothing
        const uint32_t nchunks = mine ? (lead + tl.warm + tl.own_len + (SPC - 1u)) >> CSH : 0u;
        const uint32_t nchunks_w = __reduce_max_sync(0xffffffffu, nchunks);

        const bool mask_lead = nchunks && tl.own_start == tl.warm && lead;
        const uint32_t nstages = (nchunks + CPS - 1u) / CPS;
        auto issue = [&](uint32_t st) {
            if (st < nstages) {
                const uint32_t bytes = min((uint32_t)HX_CH, (nchunks - CPS * st) * (uint32_t)CHB);
                const uint32_t mb = (st & 1u) ? mbar1 : mbar0;
                hx_mbar_expect_tx(mb, bytes);
                hx_bulk_g2s((st & 1u) ? stage1 : stage0, src + (size_t)st * HX_CH, bytes, mb);
            }
        };
        // one chunk in flight (two measured slower: more registers)
        uint4 nxt[NH];
#pragma unroll
        for (int h = 0; h < NH; ++h) nxt[h] = make_uint4(0, 0, 0, 0);
        const uint32_t cpa_slot = hx_smem_u32(txt + (size_t)threadIdx.x * 16);  // stage 1: + HX_THREADS * 16
        const auto fetch = [&](uint32_t c) {
            if constexpr (CPA) {
                asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n\tcp.async.commit_group;" ::"r"(cpa_slot + (c & 1u) * (HX_THREADS * 16u)),
                             "l"(src + (uint64_t)CHB * c)
                             : "memory");
            } else if constexpr (NH == 2) hx_ldg32_stream(src + (uint64_t)CHB * c, pol, nxt[0], nxt[NH - 1]);
            else nxt[0] = hx_ldg16_stream(src + (uint64_t)CHB * c, pol);
        };
        if constexpr (TMA) {
            issue(0);
            issue(1);
        } else {
            if (nchunks) fetch(0);
        }

        for (uint32_t c = 0; c < nchunks_w; ++c) {
            uint4 chunk[NH];
#pragma unroll
            for (int h = 0; h < NH; ++h) chunk[h] = make_uint4(0, 0, 0, 0);
            const uint32_t st = c / CPS, cc = c % CPS;
            if constexpr (TMA) {
                if (c < nchunks) {
                    if (cc == 0) {
                        if (st & 1u) hx_mbar_wait(mbar1, uses1++ & 1u);
                        else hx_mbar_wait(mbar0, uses0++ & 1u);
                    }
                    const uint32_t sa = ((st & 1u) ? stage1 : stage0) + cc * (uint32_t)CHB;
#pragma unroll
                    for (int h = 0; h < NH; ++h)
                        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                                     : "=r"(chunk[h].x), "=r"(chunk[h].y), "=r"(chunk[h].z), "=r"(chunk[h].w)
                                     : "r"(sa + 16u * h)
                                     : "memory");
                }
            } else if constexpr (CPA) {
                if (c < nchunks) {  // chunk c was the only copy in flight
                    asm volatile("cp.async.wait_group 0;" ::: "memory");
                    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                                 : "=r"(chunk[0].x), "=r"(chunk[0].y), "=r"(chunk[0].z), "=r"(chunk[0].w)
                                 : "r"(cpa_slot + (c & 1u) * (HX_THREADS * 16u))
                                 : "memory");
                    if (c + 1 < nchunks) fetch(c + 1);
                }
            } else {
#pragma unroll
                for (int h = 0; h < NH; ++h) chunk[h] = nxt[h];
                if (c + 1 < nchunks) {
                    fetch(c + 1);
                } else {
#pragma unroll
                    for (int h = 0; h < NH; ++h) nxt[h] = make_uint4(0, 0, 0, 0);
                }
            }
            if (c == 0 && mask_lead) {
                // The scan starts at the record start: bytes before it belong to the previous record.
                // Byte 0 selects the all-zero Peq row, under which a fresh automaton stays fresh.
#pragma unroll
                for (int h = 0; h < NH; ++h) {
                    uint32_t w[4] = {chunk[h].x, chunk[h].y, chunk[h].z, chunk[h].w};
#pragma unroll
                    for (int wi = 0; wi < 4; ++wi) {
                        const int z = (int)lead - SPW * (4 * h + wi);  // symbols of this word to clear
                        if (z >= SPW) w[wi] = 0;
                        else if (z > 0) w[wi] &= 0xffffffffu << (SYM_BITS * z);
                    }
                    chunk[h] = make_uint4(w[0], w[1], w[2], w[3]);
                }
            }
#pragma unroll 1
            for (int h = 0; h < NH; ++h) {
            const uint4 cur = (NH == 2 && h) ? chunk[NH - 1] : chunk[0];
            // The four words of a chunk.  With the register-destination prefetch of round 1 a rolled loop rotated them
            // through cur.x and ptxas made the rotation wait for the chunk in flight, so the packed -m 0 kernel unrolled all
            // four.  Since the chunk comes out of shared memory (CPA) that reason is gone and the copies only cost
            // instruction-cache space (the trip AND its replay path are replicated): c2 device step at -m 0, words
            // unrolled 4 / 1 / 2: 535 / 567 / 578 Gbases/s.  HX_WI_UNROLL = 2.
#pragma unroll(WI_UNROLL)
            for (int wi = 0; wi < 4; ++wi) {
                uint32_t w = hx_w32(cur, wi);
#pragma unroll 1
                for (int b0 = 0; b0 < SPW; b0 += UB) {
                    if constexpr (PK && NR == 1) {
                        // Fast trip: step the UB symbols without looking for hits one by one (4 branches, their
                        // reconvergence points and compares were 12 % of the loop), only OR-ing the hit bits together.
                        // When a lane does see a hit, the trip is replayed symbol by symbol from the kept state.
                        W keep[NW];
#pragma unroll
                        for (int j = 0; j < NW; ++j) keep[j] = rv[0][j];
                        uint32_t wf = w;
                        uint32_t accb[UB];  // per symbol: AND of the NW words, as a tree (a 32-word AND chain is 16 dependent LOP3)
#pragma unroll
                        for (int bi = 0; bi < UB; ++bi) {
                            const uint32_t byte = wf & ((1u << SYM_BITS) - 1u);
                            wf >>= SYM_BITS;
                            HxRow<W, NW> row;
                            row.load(tab + byte * STRIDE);
                            uint32_t t3[(NW + 2) / 3];
#pragma unroll
                            for (int j = 0; j < NW; ++j) rv[0][j] = ((rv[0][j] << 1) & (W)FENCE32) | row.eq(j);
#pragma unroll
                            for (int j = 0; j < (NW + 2) / 3; ++j)
                                t3[j] = hx_and3((uint32_t)rv[0][3 * j], 3 * j + 1 < NW ? (uint32_t)rv[0][3 * j + 1] : ~0u,
                                                3 * j + 2 < NW ? (uint32_t)rv[0][3 * j + 2] : ~0u);
                            accb[bi] = hx_and3(t3[0], (NW + 2) / 3 > 1 ? t3[(NW + 2) / 3 > 1 ? 1 : 0] : ~0u,
                                               (NW + 2) / 3 > 2 ? t3[(NW + 2) / 3 > 2 ? 2 : 0] : ~0u);
                        }
                        static_assert(NW <= 9 && UB <= 4, "AND tree of the fast trip");
                        const uint32_t hit = ~hx_and3(hx_and3(accb[0], accb[UB > 1 ? 1 : 0], accb[UB > 2 ? 2 : 0]), accb[UB > 3 ? 3 : 0], ~0u);
                        if (!(hit & (W)TOPS)) {
                            w = wf;
                            continue;
                        }
#pragma unroll
                        for (int j = 0; j < NW; ++j) rv[0][j] = keep[j];
                    }
#pragma unroll
                    for (int bi = 0; bi < UB; ++bi) {
                        const uint32_t byte = w & ((1u << SYM_BITS) - 1u);
                        w >>= SYM_BITS;
                        HxRow<W, NW> row;
                        row.load(tab + byte * STRIDE);
                        int any = 0;
                        if constexpr (PK) {
                            // two Shift-Or automata per word: R' = ((R << 1) & fence) | neq.  Shift-Or has no carries,
                            // so the only coupling of the halves is the bit the shift moves from the low half's top
                            // into the high half's bottom; the fence clears it (one LOP3 with the OR).
                            // The Wu-Manber levels (see below) use the same fenced shift for R(d-1) << 1; the other
                            // shifted terms are ANDed with it, which clears their stray bit as well.
                            constexpr W FENCE = (W)FENCE32;
                            W acc = ~(W)0;
#pragma unroll
                            for (int j = 0; j < NW; ++j) {
                                const W n = row.eq(j);
                                W prev_old = rv[0][j], prev_old_s = (prev_old << 1) & FENCE;
                                W prev_new = prev_old_s | n;
                                rv[0][j] = prev_new;
#pragma unroll
                                for (int d = 1; d < NR; ++d) {
                                    const W cur_old = rv[d][j];
                                    const W cur_old_s = d + 1 < NR ? (W)((cur_old << 1) & FENCE) : (W)(cur_old << 1);
                                    const W cur_new = ((cur_old_s | n) & prev_old_s) & prev_old & (prev_new << 1);
                                    rv[d][j] = cur_new;
                                    prev_old = cur_old;
                                    prev_old_s = cur_old_s;
                                    prev_new = cur_new;
                                }
                                acc &= prev_new;
                            }
                            any = (~acc & (W)TOPS) ? -1 : 0;
                        } else if constexpr (WM) {
                            // Wu-Manber, Shift-Or form: row.eq(p) is the COMPLEMENTED Peq word
                            //   R0' = (R0 << 1) | neq
                            //   Rd' = ((Rd << 1) | neq) & R(d-1) & (R(d-1) << 1) & (R(d-1)' << 1)
                            // (insertion, substitution, deletion); dist <= k  <=>  top bit of Rk is 0.
                            W acc = ~(W)0;
#pragma unroll
                            for (int p = 0; p < NP; ++p) {
                                const W n = row.eq(p);
                                W prev_old = rv[0][p], prev_old_s = prev_old << 1;
                                W prev_new = prev_old_s | n;
                                rv[0][p] = prev_new;
#pragma unroll
                                for (int d = 1; d < NR; ++d) {
                                    const W cur_old = rv[d][p], cur_old_s = cur_old << 1;
                                    const W cur_new = ((cur_old_s | n) & prev_old_s) & prev_old & (prev_new << 1);
                                    rv[d][p] = cur_new;
                                    prev_old = cur_old;
                                    prev_old_s = cur_old_s;
                                    prev_new = cur_new;
                                }
                                acc &= prev_new;
                            }
                            any = (~acc >> (sizeof(W) * 8 - 1)) ? -1 : 0;
                        } else {
#pragma unroll
                            for (int p = 0; p < NP; ++p) {
                                hx_step(row.eq(p), pv[p], mv[p], s[p]);
                                any |= s[p];
                            }
                        }
                        if (any < 0) {
                            // ---- rare: some pattern has dist <= k here; record the event ----
                            const uint32_t pos = pos0 + (c << CSH) + (4 * h + wi) * SPW + b0 + bi;
                            if (pos - tl.own_start < own_len) {
                                if constexpr (PK) {
                                    // This path runs once per hit position and lane with the other 31 lanes idle (a read
                                    // with 15 primer sites makes a warp take it for a third of its text bytes), so it only
                                    // queues what the drain needs: per word, the top bit of both halves at every level,
                                    // level d at bit 31 - d / 15 - d.
                                    uint32_t *r = queue[cnt];
                                    r[0] = pos;
#pragma unroll
                                    for (int j = 0; j < NW; ++j) {
                                        uint32_t lv = NR == 1 ? (uint32_t)rv[0][j] : (uint32_t)rv[0][j] & TOPS;
#pragma unroll
                                        for (int d = 1; d < NR; ++d) lv |= ((uint32_t)rv[d][j] & TOPS) >> d;
                                        r[1 + j] = lv;
                                    }
                                    ++cnt;
                                } else if constexpr (WM) {
                                    // Exact distance of a hit = the lowest level whose top bit is 0.  "<= d errors"
                                    // implies "<= d + 1 errors", so the top bits are monotone over the levels and that
                                    // level is simply the NUMBER of levels whose top bit is set (k + 1, i.e. a biased
                                    // score of 0 = no hit, when all are).  One add per level instead of a compare +
                                    // select: this path runs once per hit position and lane with 31 lanes idle, and at
                                    // -m 2 a read has ~65 hit positions (ncu/SASS: it was as long as 1.5 text bytes).
#pragma unroll
                                    for (int p = 0; p < NP; ++p) {
                                        int sp = -bias;
#pragma unroll
                                        for (int d = 0; d < NR; ++d) sp += (int)(rv[d][p] >> (sizeof(W) * 8 - 1));
                                        s[p] = sp;
                                    }
                                }
                                if constexpr (!PK) {
                                uint32_t *r = queue[cnt];
                                r[0] = pos;
#pragma unroll
                                for (int j = 0; j < SW; ++j) {
                                    const uint32_t b0_ = (uint32_t)s[4 * j];
                                    const uint32_t b1_ = (4 * j + 1 < NP) ? (uint32_t)s[4 * j + 1] : 0x7fu;
                                    const uint32_t b2_ = (4 * j + 2 < NP) ? (uint32_t)s[4 * j + 2] : 0x7fu;
                                    const uint32_t b3_ = (4 * j + 3 < NP) ? (uint32_t)s[4 * j + 3] : 0x7fu;
                                    r[1 + j] =
                                        __byte_perm(__byte_perm(b0_, b1_, 0x0040), __byte_perm(b2_, b3_, 0x0040), 0x5410);
                                }
                                ++cnt;
                                }
                            }
                        }
                    }
                }
                // ---- warp-synchronous drain: all lanes process their queued events together ----
                const bool last_word = (c + 1 == nchunks_w) && h == NH - 1 && wi == 3;
                if (__any_sync(0xffffffffu, cnt > (uint32_t)(QCAP - SPW) || (last_word && cnt))) {
                    // verification automaton of the last candidate (LF): a true site of a filtered primer yields up
                    // to 2k + 1 candidates at consecutive ends, and the next one costs a single step when the
                    // automaton that verified this one is kept (a longer warm-up leaves the distance exact)
                    uint64_t v_pv = 0, v_mv = 0;
                    int v_s = 0, v_pat = -1;
                    uint32_t v_pos = 0;
                    (void)v_pv; (void)v_mv; (void)v_s; (void)v_pat; (void)v_pos;
                    for (uint32_t i = 0; i < cnt; ++i) {
                        const uint32_t *r = queue[i];
                        const uint32_t pos = r[0];
                        uint32_t sw[SW];
                        uint32_t hm = 0;  // bit p: pattern p hit (its biased score byte is negative)
                        if constexpr (PK) {
                            // distance of a half = the number of levels whose top bit is set (monotone over the levels)
#pragma unroll
                            for (int j = 0; j < SW; ++j) sw[j] = 0;
                            if constexpr (KWM == 0) {
                                // -m 0: a hit is distance 0 (biased score -1), so only the hit mask is needed — the
                                // complemented top bit of every part, gathered word by word
#pragma unroll
                                for (int j = 0; j < NW; ++j) {
                                    const uint32_t hb = ~r[1 + j] & TOPS;
                                    const uint32_t bits = PPW == 4 ? (((hb >> 7) & 1u) | ((hb >> 14) & 2u) | ((hb >> 21) & 4u) | ((hb >> 28) & 8u))
                                                                   : (((hb >> 15) & 1u) | ((hb >> 30) & 2u));
                                    hm |= bits << (PPW * j);
                                }
#pragma unroll
                                for (int j = 0; j < SW; ++j) sw[j] = 0xffffffffu;  // every byte: biased score -1 (only read for hits)
                            } else {
#pragma unroll
                            for (int p = 0; p < NP; ++p) {
                                const uint32_t lv = (r[1 + p / PPW] >> (PB * (p % PPW + 1) - NR)) & ((1u << NR) - 1u);
                                const int sc = __popc(lv) - bias;
                                hm |= (sc < 0 ? 1u : 0u) << p;
                                sw[p >> 2] |= ((uint32_t)sc & 0xffu) << (8 * (p & 3));
                            }
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < SW; ++j) {
                                sw[j] = r[1 + j];
                                const uint32_t neg = sw[j] & 0x80808080u;
                                hm |= (((neg >> 7) & 1u) | ((neg >> 14) & 2u) | ((neg >> 21) & 4u) | ((neg >> 28) & 8u)) << (4 * j);
                            }
                        }
                        if constexpr (LF && PK && KWM == 0) {
                            // events of 16-symbol suffixes are candidates: the rest of the primer must match too
                            uint32_t lm = hm & g.pk_longmask;
                            while (lm) {
                                const int p = __ffs((int)lm) - 1;
                                lm &= lm - 1;
                                if (!hx_verify_exact<NIB, PB>(a.arena, rec0, pos, a.vtables + ((size_t)g.pk_vslot[p] * 2 + pass) * 256,
                                                              g.mfull[p], pass))
                                    hm &= ~(1u << p);
                            }
                        } else if constexpr (LF) {
                            // events of suffix-filtered patterns are candidates: exact distance of the whole primer
                            uint32_t lm = hm & (PK ? g.pk_longmask : g.longmask);  // (only lanes with a tile hold events)
                            while (lm) {
                                const int p = __ffs((int)lm) - 1;
                                lm &= lm - 1;
                                const uint64_t *vt = a.vtables + ((size_t)(PK ? g.pk_vslot[p] : g.vslot[p]) * 2 + pass) * 256;
                                const uint32_t mfull = g.mfull[p];
                                uint32_t d;
                                if (mfull <= 32u) {  // 32-bit verification automaton in the low halves of the kept state
                                    uint32_t pv32 = (uint32_t)v_pv, mv32 = (uint32_t)v_mv;
                                    if (v_pat == p && v_pos + 1u == pos) {
                                        hx_step(hx_vrow<uint32_t>(vt, hx_text_byte<NIB>(a.arena, rec0 + pos, pass)), pv32, mv32, v_s);
                                        d = (uint32_t)v_s;
                                    } else {
                                        d = hx_verify_long<NIB, uint32_t>(a.arena, rec0, pos, vt, mfull, a.k, pass, pv32, mv32);
                                        v_s = (int)d;
                                        v_pat = p;
                                    }
                                    v_pv = pv32;
                                    v_mv = mv32;
                                } else if (v_pat == p && v_pos + 1u == pos) {
                                    hx_step(hx_vrow<uint64_t>(vt, hx_text_byte<NIB>(a.arena, rec0 + pos, pass)), v_pv, v_mv, v_s);
                                    d = (uint32_t)v_s;
                                } else {
                                    d = hx_verify_long<NIB, uint64_t>(a.arena, rec0, pos, vt, mfull, a.k, pass, v_pv, v_mv);
                                    v_s = (int)d;
                                    v_pat = p;
                                }
                                v_pos = pos;
                                const uint32_t sh = 8u * (p & 3);
                                if (d > a.k) hm &= ~(1u << p);
#pragma unroll
                                for (int j = 0; j < SW; ++j)
                                    if ((p >> 2) == j) sw[j] = (sw[j] & ~(0xffu << sh)) | (((d - (uint32_t)bias) & 0xffu) << sh);
                            }
                        }
                        // reverse roles first: forward candidates must end strictly before pos
                        uint32_t m1 = hm;
                        while (m1) {
                            const int p = __ffs((int)m1) - 1;
                            m1 &= m1 - 1;
                            uint32_t word = sw[0];
#pragma unroll
                            for (int j = 1; j < SW; ++j) word = (p >> 2) == j ? sw[j] : word;
                            const uint32_t d = (uint32_t)((int)(int8_t)(word >> (8 * (p & 3))) + bias);
                            uint32_t rm = g.rmask[p];  // pairs whose reverse pattern is p
                            while (rm) {
                                const int q = __ffs((int)rm) - 1;
                                rm &= rm - 1;
                                const uint64_t bf = bestF[g.pair_f[q] + dz];
                                if (bf != HX_NONE64) {
                                    const uint32_t comb = (uint32_t)(bf >> 32) + d;
                                    if (comb < (uint32_t)(phi[q + dz] >> 32)) {
                                        phi[q + dz] = ((uint64_t)comb << 32) | (uint32_t)bf;
                                        plo[q + dz] = pos;
                                    }
                                }
                            }
                        }
                        uint32_t m2 = hm;
                        while (m2) {
                            const int p = __ffs((int)m2) - 1;
                            m2 &= m2 - 1;
                            uint32_t word = sw[0];
#pragma unroll
                            for (int j = 1; j < SW; ++j) word = (p >> 2) == j ? sw[j] : word;
                            const uint32_t d = (uint32_t)((int)(int8_t)(word >> (8 * (p & 3))) + bias);
                            const uint64_t key = ((uint64_t)d << 32) | pos;
                            if (key < bestR[p + dz]) bestR[p + dz] = key;
                            if (pos + 1 >= g.mfull[p] && key < bestF[p + dz]) bestF[p + dz] = key;
                        }
                    }
                    cnt = 0;
                }
            }
            }  // halves of the chunk
            if constexpr (TMA) {
                if (cc == CPS - 1u && c < nchunks) {
                    hx_fence_proxy_async();  // my reads of this stage precede the TMA write that refills it
                    issue(st + 2u);
                }
            }
        }
    }

    if (!have || alpha == 2) continue;
    if (single) {
        // the tile is the whole record: final result of utils.rs:312-351 for every pair of the group
        uint32_t *out = reinterpret_cast<uint32_t *>(a.out) + (size_t)tl.rec * a.n_pairs_total * 3;
        for (int q = 0; q < g.npairs; ++q) {
            const int f = g.pair_f[q], r = g.pair_r[q];
            uint32_t flags = (bestR[f + dz] != HX_NONE64 ? 1u : 0u) | (bestR[r + dz] != HX_NONE64 ? 2u : 0u) |
                             (alpha == 1 ? 16u : 0u);
            uint32_t f_start = 0, r_end = 0, comb = 0;
            const uint64_t hi = phi[q + dz];
            if (hi != HX_NONE64) {
                flags |= 4u;
                f_start = (uint32_t)hi - ((uint32_t)g.mfull[f] - 1u);
                r_end = plo[q + dz];
                comb = (uint32_t)(hi >> 32) & 0xffu;
            }
            uint32_t *o = out + (size_t)g.pair_global[q] * 3;
            o[0] = f_start;
            o[1] = r_end;
            o[2] = flags | (comb << 8);
        }
        continue;
    }
    const size_t ts = a.tile_stride;
    for (int p = 0; p < NP; ++p) {
        if (PK && p >= g.np) break;  // the padding half of an odd pattern count owns no summary slot
        a.sum_f[(size_t)(g.slot_base + p) * ts + tile_id] = bestF[p + dz];
        a.sum_r[(size_t)(g.slot_base + p) * ts + tile_id] = bestR[p + dz];
    }
    for (int q = 0; q < g.npairs; ++q) {
        a.pair_hi[(size_t)g.pair_global[q] * ts + tile_id] = phi[q + dz];
        a.pair_lo[(size_t)g.pair_global[q] * ts + tile_id] = plo[q + dz];
    }
    }  // batches of this CTA
    if (a.sched && threadIdx.x == 0 && atomicAdd(a.sched + 1, 1u) == gridDim.x - 1u) {
        a.sched[0] = 0u;  // every CTA has taken its last batch number: nobody reads the counters any more
        a.sched[1] = 0u;
    }
}

template <typename W, int NP, bool TMA, int KWM, bool LF = false, bool PK = false, bool NIB = false, int PPW = 2>
static int hx_scan_launch_one(const HxScanArgs &a, uint32_t tile_cap, cudaStream_t st) {
    const int smem = hx_table_bytes(PK ? NP / PPW : NP, sizeof(W)) / (NIB ? 16 : 1) + (TMA ? 2 * HX_THREADS * (HX_CH + 8) : 0) +
                     ((HX_CPASYNC && (PK || HX_CPASYNC >= 3) && (KWM == 0 || HX_CPASYNC >= 2) && !TMA) ? 2 * HX_THREADS * 16 : 0);  // + the cp.async text slots (CPA in the kernel)
    // per (instantiation, device) once; contexts on different host threads may race here, so the flags are atomic
    // (setting the attribute twice is harmless, skipping it is not)
    // max_ctas: what the device holds at once (CTAs per SM x SMs) — the kernel's CTAs are persistent and walk the batches
    static std::atomic<int> max_ctas[64];
    int dev = 0;
    cudaGetDevice(&dev);
    int resident = dev < 64 ? max_ctas[dev].load(std::memory_order_acquire) : 0;
    if (resident <= 0) {
        cudaFuncSetAttribute(hx_myers_pair_kernel<W, NP, TMA, KWM, LF, PK, NIB, PPW>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        int per_sm = 0, sms = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, hx_myers_pair_kernel<W, NP, TMA, KWM, LF, PK, NIB, PPW>, HX_THREADS, smem);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        resident = per_sm > 0 && sms > 0 ? per_sm * sms : 1 << 30;
        if (dev < 64) max_ctas[dev].store(resident, std::memory_order_release);
    }
    // HYPEREX_SCAN_GRID=full: one CTA per batch, as before (A/B of the persistent grid; same kernel, one round per CTA)
    static const bool full_grid = [] { const char *e = getenv("HYPEREX_SCAN_GRID"); return e && !strcmp(e, "full"); }();
    const uint32_t nb = (tile_cap + HX_THREADS - 1) / HX_THREADS;
    const uint32_t blocks = full_grid ? nb : std::min<uint32_t>(nb, (uint32_t)resident);
    static const bool static_rounds = [] { const char *e = getenv("HYPEREX_SCAN_GRID"); return e && !strcmp(e, "static"); }();
    if (full_grid || static_rounds) {
        HxScanArgs b = a;
        b.sched = nullptr;
        hx_myers_pair_kernel<W, NP, TMA, KWM, LF, PK, NIB, PPW><<<blocks, HX_THREADS, smem, st>>>(b);
        return 1;
    }
    hx_myers_pair_kernel<W, NP, TMA, KWM, LF, PK, NIB, PPW><<<blocks, HX_THREADS, smem, st>>>(a);
    return 1;
}

template <typename W, int NP>
struct HxDispatch {
    static int run(int np, bool lf, const HxScanArgs &a, uint32_t cap, cudaStream_t st) {
        if (np == NP) {
            if (a.nib) {  // packed text: one variant per algorithm (the suffix-filter code path covers groups without one)
                if constexpr (sizeof(W) == 4) {
                    if (a.wm_k == 0) return hx_scan_launch_one<W, NP, false, 0, true, false, true>(a, cap, st);
                    if (a.wm_k == 1) return hx_scan_launch_one<W, NP, false, 1, true, false, true>(a, cap, st);
                    if (a.wm_k == 2) return hx_scan_launch_one<W, NP, false, 2, true, false, true>(a, cap, st);
                    return hx_scan_launch_one<W, NP, false, -1, true, false, true>(a, cap, st);
                } else {
                    return hx_scan_launch_one<W, NP, false, -1, false, false, true>(a, cap, st);
                }
            }
            if constexpr (sizeof(W) == 4) {
                if (lf) {  // group with suffix-filtered primers: vector-load text path only
                    if (a.wm_k == 0) return hx_scan_launch_one<W, NP, false, 0, true>(a, cap, st);
                    if (a.wm_k == 1) return hx_scan_launch_one<W, NP, false, 1, true>(a, cap, st);
                    if (a.wm_k == 2) return hx_scan_launch_one<W, NP, false, 2, true>(a, cap, st);
                    return hx_scan_launch_one<W, NP, false, -1, true>(a, cap, st);
                }
                if (a.wm_k == 0) return hx_scan_launch_one<W, NP, false, 0>(a, cap, st);
                if (a.wm_k == 1) return hx_scan_launch_one<W, NP, false, 1>(a, cap, st);
                if (a.wm_k == 2) return hx_scan_launch_one<W, NP, false, 2>(a, cap, st);
            }
            return a.variant == 1 ? hx_scan_launch_one<W, NP, true, -1>(a, cap, st)
                                  : hx_scan_launch_one<W, NP, false, -1>(a, cap, st);
        }
        return HxDispatch<W, NP - 1>::run(np, lf, a, cap, st);
    }
};
template <typename W>
struct HxDispatch<W, 0> {
    static int run(int, bool, const HxScanArgs &, uint32_t, cudaStream_t) { return -1; }
};

// packed small-k scan: NW words = PPW * NW automata of <= 32 / PPW symbols
template <int NW, int K, int PPW = 2>
static int hx_packed_dispatch(int nw, bool lf, const HxScanArgs &a, uint32_t cap, cudaStream_t st) {
    if (nw == NW) {
        if (a.nib) return hx_scan_launch_one<uint32_t, PPW * NW, false, K, true, true, true, PPW>(a, cap, st);
        return lf ? hx_scan_launch_one<uint32_t, PPW * NW, false, K, true, true, false, PPW>(a, cap, st)
                  : hx_scan_launch_one<uint32_t, PPW * NW, false, K, false, true, false, PPW>(a, cap, st);
    }
    if constexpr (NW > 1) return hx_packed_dispatch<NW - 1, K, PPW>(nw, lf, a, cap, st);
    return -1;
}

int hx_launch_scan(const HxScanArgs &a, const HxGroup &hg, uint32_t tile_cap, cudaStream_t st) {
    if (tile_cap == 0) return 0;
    if (a.packed && hg.packed && !hg.word64) {
        const bool lf = hg.pk_longmask != 0u;
        if (hg.pk_parts == 4 && a.wm_k == 0) return hx_packed_dispatch<HX_MAX_NP32 / 4, 0, 4>((hg.np + 3) / 4, lf, a, tile_cap, st);
        const int nw = (hg.np + 1) / 2;
        if (a.wm_k == 0) return hx_packed_dispatch<HX_MAX_NP32 / 2, 0>(nw, lf, a, tile_cap, st);
        if (a.wm_k == 1) return hx_packed_dispatch<HX_MAX_NP32 / 2, 1>(nw, lf, a, tile_cap, st);
        if (a.wm_k == 2) return hx_packed_dispatch<HX_MAX_NP32 / 2, 2>(nw, lf, a, tile_cap, st);
    }
    if (hg.word64) return HxDispatch<uint64_t, HX_MAX_NP64>::run(hg.np, false, a, tile_cap, st);
    return HxDispatch<uint32_t, HX_MAX_NP32>::run(hg.np, hg.longmask != 0u, a, tile_cap, st);
}

// ------------------------------------------------------------------------------------
// hx_combine_kernel: one thread per (pair, record).  Walks the record's tiles left to right
// and joins their summaries with the same streaming rule, at tile granularity:
//   cross candidates  (forward hit in an earlier tile, reverse hit in this tile) = bestF so far
//                     + this tile's bestR;  within-tile candidates come from the scan kernel.
// Result = lexicographic min of (combined, f_end, r_end)  ==  utils.rs:327-351.
// ------------------------------------------------------------------------------------
struct HxResultDev {
    uint32_t f_start, r_end, meta;
};

__global__ void __launch_bounds__(256)
    hx_combine_kernel(const uint32_t *__restrict__ tile_first, const uint32_t *__restrict__ tile_cnt,
                      const uint32_t *__restrict__ rec_raw, const HxPairRef *__restrict__ pairs, uint32_t n_pairs,
                      uint32_t n_records, const uint64_t *__restrict__ sum_f, const uint64_t *__restrict__ sum_r,
                      const uint64_t *__restrict__ pair_hi, const uint32_t *__restrict__ pair_lo, uint32_t ts,
                      HxResultDev *__restrict__ out) {
    // grid-stride over (record, pair) with the pair as the fast index: neighbouring threads store neighbouring
    // hx_result entries, and the grid stays far below gridDim.x's limit for any n_pairs x n_records
    const uint64_t total = (uint64_t)n_pairs * n_records;
    for (uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t r = (uint32_t)(idx / n_pairs);
        const uint32_t q = (uint32_t)(idx % n_pairs);
        const int alpha = hx_alpha_of(rec_raw[r]);
        HxResultDev res = {0u, 0u, 0u};
        if (alpha == 2) {
            res.meta = 8u;
            out[(size_t)r * n_pairs + q] = res;
            continue;
        }
        const HxPairRef pr = pairs[q];
        const uint32_t t0 = tile_first[r], nt = tile_cnt[r];
        if (nt == 1u) continue;             // single-tile record: hx_myers_pair_kernel wrote the final result
        if (nt > HX_LONG_TILES) continue;  // contig: hx_combine_long_kernel
        uint64_t bestF = HX_NONE64;          // (dist<<32 | end) of the best valid forward hit so far
        uint64_t bhi = HX_NONE64;            // combined<<32 | f_end
        uint32_t blo = 0;                    // r_end
        bool fany = false, rany = false;
        for (uint32_t i = 0; i < nt; ++i) {
            const size_t t = t0 + i;
            const uint64_t tf = sum_f[(size_t)pr.f_slot * ts + t];
            const uint64_t tfr = sum_r[(size_t)pr.f_slot * ts + t];
            const uint64_t tr = sum_r[(size_t)pr.r_slot * ts + t];
            fany |= tfr != HX_NONE64;
            rany |= tr != HX_NONE64;
            if (bestF != HX_NONE64 && tr != HX_NONE64) {
                const uint64_t chi = (((bestF >> 32) + (tr >> 32)) << 32) | (uint32_t)bestF;
                const uint32_t clo = (uint32_t)tr;
                if (chi < bhi || (chi == bhi && clo < blo)) {
                    bhi = chi;
                    blo = clo;
                }
            }
            const uint64_t whi = pair_hi[(size_t)q * ts + t];
            if (whi != HX_NONE64) {
                const uint32_t wlo = pair_lo[(size_t)q * ts + t];
                if (whi < bhi || (whi == bhi && wlo < blo)) {
                    bhi = whi;
                    blo = wlo;
                }
            }
            if (tf < bestF) bestF = tf;
        }
        uint32_t flags = (fany ? 1u : 0u) | (rany ? 2u : 0u) | (alpha == 1 ? 16u : 0u);
        if (bhi != HX_NONE64) {
            flags |= 4u;
            res.f_start = (uint32_t)bhi - (pr.len_f - 1u);
            res.r_end = blo;
            res.meta = flags | (uint32_t)((bhi >> 32) & 0xffu) << 8;
        } else {
            res.meta = flags;
        }
        out[(size_t)r * n_pairs + q] = res;
    }
}

// Records with many windows (contigs): one warp per (long record, pair).  Every lane joins a contiguous run of
// tiles with the streaming rule, then the 32 ordered partial summaries are merged by a shuffle tree with the
// (associative, order-sensitive) segment join: best pair = lex-min(left, right, left.bestF + right.bestR).
struct HxSeg {
    uint64_t bestF, bestR, phi;
    uint32_t plo, any;
};

__device__ __forceinline__ void hx_seg_join(HxSeg &L, const HxSeg &R) {
    if (L.bestF != HX_NONE64 && R.bestR != HX_NONE64) {
        const uint64_t chi = (((L.bestF >> 32) + (R.bestR >> 32)) << 32) | (uint32_t)L.bestF;
        const uint32_t clo = (uint32_t)R.bestR;
        if (chi < L.phi || (chi == L.phi && clo < L.plo)) {
            L.phi = chi;
            L.plo = clo;
        }
    }
    if (R.phi < L.phi || (R.phi == L.phi && R.phi != HX_NONE64 && R.plo < L.plo)) {
        L.phi = R.phi;
        L.plo = R.plo;
    }
    if (R.bestF < L.bestF) L.bestF = R.bestF;
    if (R.bestR < L.bestR) L.bestR = R.bestR;
    L.any |= R.any;
}

__global__ void __launch_bounds__(256)
    hx_combine_long_kernel(const uint32_t *__restrict__ long_list, const uint32_t *__restrict__ n_long_p,
                           const uint32_t *__restrict__ tile_first, const uint32_t *__restrict__ tile_cnt,
                           const uint32_t *__restrict__ rec_raw, const HxPairRef *__restrict__ pairs, uint32_t n_pairs,
                           const uint64_t *__restrict__ sum_f, const uint64_t *__restrict__ sum_r,
                           const uint64_t *__restrict__ pair_hi, const uint32_t *__restrict__ pair_lo, uint32_t ts,
                           HxResultDev *__restrict__ out) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t items = (uint64_t)(*n_long_p) * n_pairs;
    const uint64_t warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t it = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; it < items; it += warps) {
        const uint32_t r = long_list[it / n_pairs], q = (uint32_t)(it % n_pairs);
        const int alpha = hx_alpha_of(rec_raw[r]);
        if (alpha == 2) continue;  // hx_combine_kernel wrote the "alphabet none" result
        const HxPairRef pr = pairs[q];
        const uint32_t t0 = tile_first[r], nt = tile_cnt[r];
        const uint32_t per = (nt + 31u) / 32u;
        const uint32_t lo = min(nt, lane * per), hi = min(nt, lo + per);
        HxSeg S = {HX_NONE64, HX_NONE64, HX_NONE64, 0u, 0u};
        for (uint32_t i = lo; i < hi; ++i) {
            const size_t t = t0 + i;
            HxSeg T;
            T.bestF = sum_f[(size_t)pr.f_slot * ts + t];
            T.bestR = sum_r[(size_t)pr.r_slot * ts + t];
            T.phi = pair_hi[(size_t)q * ts + t];
            T.plo = pair_lo[(size_t)q * ts + t];
            T.any = (sum_r[(size_t)pr.f_slot * ts + t] != HX_NONE64 ? 1u : 0u) | (T.bestR != HX_NONE64 ? 2u : 0u);
            hx_seg_join(S, T);
        }
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            HxSeg R;
            R.bestF = __shfl_down_sync(0xffffffffu, S.bestF, d);
            R.bestR = __shfl_down_sync(0xffffffffu, S.bestR, d);
            R.phi = __shfl_down_sync(0xffffffffu, S.phi, d);
            R.plo = __shfl_down_sync(0xffffffffu, S.plo, d);
            R.any = __shfl_down_sync(0xffffffffu, S.any, d);
            if ((lane & (2 * d - 1)) == 0) hx_seg_join(S, R);
        }
        if (lane == 0) {
            HxResultDev res = {0u, 0u, 0u};
            uint32_t flags = S.any | (alpha == 1 ? 16u : 0u);
            if (S.phi != HX_NONE64) {
                flags |= 4u;
                res.f_start = (uint32_t)S.phi - (pr.len_f - 1u);
                res.r_end = S.plo;
                flags |= (uint32_t)((S.phi >> 32) & 0xffu) << 8;
            }
            res.meta = flags;
            out[(size_t)r * n_pairs + q] = res;
        }
    }
}

int hx_launch_combine(const uint32_t *tile_first, const uint32_t *tile_cnt, const uint32_t *rec_raw,
                      const HxPairRef *pairs, uint32_t n_pairs, uint32_t n_records, const uint64_t *sum_f,
                      const uint64_t *sum_r, const uint64_t *pair_hi, const uint32_t *pair_lo, uint32_t tile_stride,
                      const uint32_t *long_list, const uint32_t *n_long, void *out, cudaStream_t st) {
    const uint64_t total = (uint64_t)n_pairs * n_records;
    if (total == 0) return 0;
    hx_combine_long_kernel<<<148 * 4, 256, 0, st>>>(long_list, n_long, tile_first, tile_cnt, rec_raw, pairs, n_pairs, sum_f,
                                                   sum_r, pair_hi, pair_lo, tile_stride, (HxResultDev *)out);
    const uint32_t blocks = (uint32_t)std::min<uint64_t>((total + 255) / 256, 148ull * 64);
    hx_combine_kernel<<<blocks, 256, 0, st>>>(tile_first, tile_cnt, rec_raw, pairs, n_pairs,
                                                                      n_records, sum_f, sum_r, pair_hi, pair_lo,
                                                                      tile_stride, (HxResultDev *)out);
    return 2;
}

// ------------------------------------------------------------------------------------
// Known-answer mirror of Myers<u64>::find_all_lazy(text, k).collect() (utils.rs:305-309).
// pass 0 counts the hits of every window, hx_scan_u64_kernel turns the counts into write
// offsets, pass 1 writes (end, dist) in ascending text order.
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
    hx_find_kernel(const uint8_t *__restrict__ text, uint64_t n, const uint64_t *__restrict__ table, uint32_t m,
                   uint32_t k, uint32_t tile_len, uint32_t ntiles, uint64_t *__restrict__ counts,
                   uint64_t *__restrict__ out_end, uint8_t *__restrict__ out_dist, uint64_t cap, int pass) {
    __shared__ uint64_t peq[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) peq[i] = table[i];
    __syncthreads();
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ntiles) return;
    const uint64_t own_start = (uint64_t)t * tile_len;
    const uint64_t own_end = min(own_start + tile_len, n);
    const uint64_t warm = min((uint64_t)(m + k - 1), own_start);
    uint64_t pv = ~0ull, mv = 0;
    int s = (int)m - (int)k - 1;
    uint64_t cnt = 0;
    const uint64_t base = pass ? counts[t] : 0;
    for (uint64_t i = own_start - warm; i < own_end; ++i) {
        hx_step(peq[text[i]], pv, mv, s);
        if (s < 0 && i >= own_start) {
            if (pass) {
                const uint64_t o = base + cnt;
                if (o < cap) {
                    out_end[o] = i;
                    out_dist[o] = (uint8_t)(s + (int)k + 1);
                }
            }
            ++cnt;
        }
    }
    if (!pass) counts[t] = cnt;
}

// exclusive scan of counts[0..n) in place, total written to counts[n]; single CTA
__global__ void __launch_bounds__(1024) hx_scan_u64_kernel(uint64_t *counts, uint32_t n) {
    __shared__ uint64_t warp_sum[32];
    __shared__ uint64_t carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (uint32_t base = 0; base < n; base += 1024) {
        const uint32_t i = base + threadIdx.x;
        const uint64_t own = i < n ? counts[i] : 0;
        uint64_t x = own;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint64_t y = __shfl_up_sync(0xffffffffu, x, d);
            if (lane >= d) x += y;
        }
        if (lane == 31) warp_sum[wid] = x;
        __syncthreads();
        if (wid == 0) {
            uint64_t w = warp_sum[lane];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                uint64_t y = __shfl_up_sync(0xffffffffu, w, d);
                if (lane >= d) w += y;
            }
            warp_sum[lane] = w;
        }
        __syncthreads();
        const uint64_t carry = carry_s;
        const uint64_t before = wid ? warp_sum[wid - 1] : 0;
        if (i < n) counts[i] = carry + before + x - own;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = carry + before + x;
        __syncthreads();
    }
    if (threadIdx.x == 0) counts[n] = carry_s;
}

int hx_launch_find(const uint8_t *text, uint64_t n, const uint8_t *table, uint32_t m, uint32_t k, uint32_t tile_len,
                   uint32_t ntiles, uint64_t *counts, uint64_t *out_end, uint8_t *out_dist, uint64_t cap, int pass,
                   cudaStream_t st) {
    if (ntiles == 0) return 0;
    hx_find_kernel<<<(ntiles + 127) / 128, 128, 0, st>>>(text, n, (const uint64_t *)table, m, k, tile_len, ntiles,
                                                        counts, out_end, out_dist, cap, pass);
    if (pass == 0) {
        hx_scan_u64_kernel<<<1, 1024, 0, st>>>(counts, ntiles);
        return 2;
    }
    return 1;
}

// ------------------------------------------------------------------------------------
// On-device extraction: the FASTA text of utils.rs:354-369 is assembled on the GPU.
//   hx_fasta_len_kernel     bytes of every (record, pair) entry in output order (0 when no region)
//   hx_scan_blocks_kernel / hx_scan_u64_kernel / hx_scan_add_kernel   exclusive scan -> write offsets
//   hx_fasta_write_kernel   one warp per entry copies '>' id tail slice '\n' (original case, utils.rs:368)
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    hx_fasta_len_kernel(const HxResultDev *__restrict__ res, const uint64_t *__restrict__ id_off,
                        const uint32_t *__restrict__ tail_off, uint32_t n_pairs, uint64_t n_entries,
                        uint64_t *__restrict__ lens) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_entries) return;
    const HxResultDev x = res[i];
    uint64_t len = 0;
    if ((x.meta & 7u) == 7u) {
        const uint64_t r = i / n_pairs;
        const uint32_t p = (uint32_t)(i % n_pairs);
        len = 1 + (id_off[r + 1] - id_off[r]) + (tail_off[p + 1] - tail_off[p]) + ((uint64_t)x.r_end + 1 - x.f_start) + 1;
    }
    lens[i] = len;
}

#define HX_SCAN_ITEMS 4096  // elements per block of the multi-block scan

// per block: exclusive scan of its HX_SCAN_ITEMS elements in place, block total to sums[block]
__global__ void __launch_bounds__(1024) hx_scan_blocks_kernel(uint64_t *data, uint64_t n, uint64_t *sums) {
    __shared__ uint64_t warp_sum[32];
    const uint64_t base = (uint64_t)blockIdx.x * HX_SCAN_ITEMS + (uint64_t)threadIdx.x * 4;
    uint64_t v[4], t = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        v[j] = base + j < n ? data[base + j] : 0;
        t += v[j];
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint64_t x = t;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint64_t y = __shfl_up_sync(0xffffffffu, x, d);
        if (lane >= d) x += y;
    }
    if (lane == 31) warp_sum[wid] = x;
    __syncthreads();
    if (wid == 0) {
        uint64_t w = warp_sum[lane];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint64_t y = __shfl_up_sync(0xffffffffu, w, d);
            if (lane >= d) w += y;
        }
        warp_sum[lane] = w;
    }
    __syncthreads();
    uint64_t run = (wid ? warp_sum[wid - 1] : 0) + x - t;  // exclusive prefix of this thread inside the block
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (base + j < n) data[base + j] = run;
        run += v[j];
    }
    if (threadIdx.x == 1023) sums[blockIdx.x] = warp_sum[31];
}

__global__ void __launch_bounds__(1024) hx_scan_add_kernel(uint64_t *data, uint64_t n, const uint64_t *sums) {
    const uint64_t add = sums[blockIdx.x];
    const uint64_t base = (uint64_t)blockIdx.x * HX_SCAN_ITEMS + (uint64_t)threadIdx.x * 4;
#pragma unroll
    for (int j = 0; j < 4; ++j)
        if (base + j < n) data[base + j] += add;
}

__global__ void __launch_bounds__(256)
    hx_fasta_write_kernel(const uint8_t *__restrict__ arena, const uint64_t *__restrict__ offsets, uint64_t off_base,
                          const HxResultDev *__restrict__ res, const uint8_t *__restrict__ ids,
                          const uint64_t *__restrict__ id_off, const uint8_t *__restrict__ tails,
                          const uint32_t *__restrict__ tail_off, uint32_t n_pairs, uint64_t n_entries,
                          const uint64_t *__restrict__ where, uint8_t *__restrict__ out) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t i = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n_entries; i += warps) {
        const HxResultDev x = res[i];
        if ((x.meta & 7u) != 7u) continue;
        const uint64_t r = i / n_pairs;
        const uint32_t p = (uint32_t)(i % n_pairs);
        uint8_t *o = out + where[i];
        if (lane == 0) *o = '>';
        ++o;
        const uint8_t *src = ids + id_off[r];
        uint64_t len = id_off[r + 1] - id_off[r];
        for (uint64_t j = lane; j < len; j += 32) o[j] = src[j];
        o += len;
        src = tails + tail_off[p];
        len = tail_off[p + 1] - tail_off[p];
        for (uint64_t j = lane; j < len; j += 32) o[j] = src[j];
        o += len;
        src = arena + (offsets[r] - off_base) + x.f_start;
        len = (uint64_t)x.r_end + 1 - x.f_start;
        for (uint64_t j = lane; j < len; j += 32) o[j] = src[j];
        if (lane == 0) o[len] = '\n';
    }
}

// lens (n_entries + 1 slots) is replaced by exclusive offsets; lens[n_entries] receives the total.
// sums must hold ceil(n_entries / HX_SCAN_ITEMS) + 1 elements.
int hx_launch_fasta_offsets(const void *res, const uint64_t *id_off, const uint32_t *tail_off, uint32_t n_pairs,
                            uint64_t n_entries, uint64_t *lens, uint64_t *sums, cudaStream_t st) {
    if (n_entries == 0) return 0;
    const uint32_t nblk = (uint32_t)((n_entries + HX_SCAN_ITEMS - 1) / HX_SCAN_ITEMS);
    hx_fasta_len_kernel<<<(uint32_t)((n_entries + 255) / 256), 256, 0, st>>>((const HxResultDev *)res, id_off, tail_off,
                                                                            n_pairs, n_entries, lens);
    hx_scan_blocks_kernel<<<nblk, 1024, 0, st>>>(lens, n_entries, sums);
    hx_scan_u64_kernel<<<1, 1024, 0, st>>>(sums, nblk);  // exclusive scan of the block totals, grand total at sums[nblk]
    hx_scan_add_kernel<<<nblk, 1024, 0, st>>>(lens, n_entries, sums);
    return 4;
}

int hx_launch_fasta_write(const uint8_t *arena, const uint64_t *offsets, uint64_t off_base, const void *res,
                          const uint8_t *ids, const uint64_t *id_off, const uint8_t *tails, const uint32_t *tail_off,
                          uint32_t n_pairs, uint64_t n_entries, const uint64_t *where, uint8_t *out, cudaStream_t st) {
    if (n_entries == 0) return 0;
    uint64_t blocks = (n_entries * 32 + 255) / 256;
    if (blocks > 148 * 32) blocks = 148 * 32;
    hx_fasta_write_kernel<<<(uint32_t)blocks, 256, 0, st>>>(arena, offsets, off_base, (const HxResultDev *)res, ids, id_off,
                                                           tails, tail_off, n_pairs, n_entries, where, out);
    return 1;
}

// ---- GFF3 lines of utils.rs:378-385 on the device: id \t hyperex \t region \t start \t end <tail>, 1-based ----
__device__ __forceinline__ uint32_t hx_digits10(uint32_t v) {
    uint32_t d = 1;
    while (v >= 10u) {
        v /= 10u;
        ++d;
    }
    return d;
}

__global__ void __launch_bounds__(256)
    hx_gff_len_kernel(const HxResultDev *__restrict__ res, const uint64_t *__restrict__ id_off,
                      const uint32_t *__restrict__ tail_off, uint32_t n_pairs, uint64_t n_entries,
                      uint64_t *__restrict__ lens) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_entries) return;
    const HxResultDev x = res[i];
    uint64_t len = 0;
    if ((x.meta & 7u) == 7u) {
        const uint64_t r = i / n_pairs;
        const uint32_t p = (uint32_t)(i % n_pairs);
        len = (id_off[r + 1] - id_off[r]) + 16 /* "\thyperex\tregion\t" */ + hx_digits10(x.f_start + 1u) + 1 +
              hx_digits10(x.r_end + 1u) + (tail_off[p + 1] - tail_off[p]);
    }
    lens[i] = len;
}

__global__ void __launch_bounds__(256)
    hx_gff_write_kernel(const HxResultDev *__restrict__ res, const uint8_t *__restrict__ ids,
                        const uint64_t *__restrict__ id_off, const uint8_t *__restrict__ tails,
                        const uint32_t *__restrict__ tail_off, uint32_t n_pairs, uint64_t n_entries,
                        const uint64_t *__restrict__ where, uint8_t *__restrict__ out) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t i = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n_entries; i += warps) {
        const HxResultDev x = res[i];
        if ((x.meta & 7u) != 7u) continue;
        const uint64_t r = i / n_pairs;
        const uint32_t p = (uint32_t)(i % n_pairs);
        uint8_t *o = out + where[i];
        const uint8_t *src = ids + id_off[r];
        uint64_t len = id_off[r + 1] - id_off[r];
        for (uint64_t j = lane; j < len; j += 32) o[j] = src[j];
        o += len;
        const uint32_t a = x.f_start + 1u, b = x.r_end + 1u;
        const uint32_t da = hx_digits10(a), db = hx_digits10(b);
        if (lane < 16u) o[lane] = (uint8_t)"\thyperex\tregion\t"[lane];
        o += 16;
        if (lane == 0) {
            uint32_t v = a;
            for (int j = (int)da - 1; j >= 0; --j) {
                o[j] = (uint8_t)('0' + v % 10u);
                v /= 10u;
            }
            o[da] = '\t';
            v = b;
            for (int j = (int)db - 1; j >= 0; --j) {
                o[da + 1 + j] = (uint8_t)('0' + v % 10u);
                v /= 10u;
            }
        }
        o += da + 1 + db;
        src = tails + tail_off[p];
        len = tail_off[p + 1] - tail_off[p];
        for (uint64_t j = lane; j < len; j += 32) o[j] = src[j];
    }
}

int hx_launch_gff_offsets(const void *res, const uint64_t *id_off, const uint32_t *tail_off, uint32_t n_pairs,
                          uint64_t n_entries, uint64_t *lens, uint64_t *sums, cudaStream_t st) {
    if (n_entries == 0) return 0;
    const uint32_t nblk = (uint32_t)((n_entries + HX_SCAN_ITEMS - 1) / HX_SCAN_ITEMS);
    hx_gff_len_kernel<<<(uint32_t)((n_entries + 255) / 256), 256, 0, st>>>((const HxResultDev *)res, id_off, tail_off,
                                                                          n_pairs, n_entries, lens);
    hx_scan_blocks_kernel<<<nblk, 1024, 0, st>>>(lens, n_entries, sums);
    hx_scan_u64_kernel<<<1, 1024, 0, st>>>(sums, nblk);
    hx_scan_add_kernel<<<nblk, 1024, 0, st>>>(lens, n_entries, sums);
    return 4;
}

int hx_launch_gff_write(const void *res, const uint8_t *ids, const uint64_t *id_off, const uint8_t *tails,
                        const uint32_t *tail_off, uint32_t n_pairs, uint64_t n_entries, const uint64_t *where, uint8_t *out,
                        cudaStream_t st) {
    if (n_entries == 0) return 0;
    uint64_t blocks = (n_entries * 32 + 255) / 256;
    if (blocks > 148 * 32) blocks = 148 * 32;
    hx_gff_write_kernel<<<(uint32_t)blocks, 256, 0, st>>>((const HxResultDev *)res, ids, id_off, tails, tail_off, n_pairs,
                                                         n_entries, where, out);
    return 1;
}
