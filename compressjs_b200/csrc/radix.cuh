// radix.cuh -- segmented LSD radix sort (8-bit digits), single pass over the data per digit:
//   k_radix_hist : one read of the keys -> per-(segment, pass) digit histograms (only for the
//                  64-bit round keys; the initial 4-byte-prefix sort gets its histograms for free
//                  from the block's byte histogram, see bwt.cu)
//   k_radix_pass : "onesweep" pass = read (key,val) once, rank inside the tile with warp
//                  match + one leader lane per digit group bumping a warp-private shared-memory
//                  counter, chain the per-digit tile offsets with a decoupled look-back, stage the
//                  tile in shared memory in bucket order and write every bucket run out coalesced.
// Segments are the independent bzip2 blocks of a batch (slot = seg << seg_shift | i), or a
// single flat segment for the compacted "still unsorted" suffixes of a doubling round.
// This is the hot kernel of the forward BWT (reference: lib/BWT.js:372-417 does the same
// job with SA-IS on one block at a time).
#pragma once
#include "common.cuh"

#define RADIX_BITS 8
#define RADIX 256
#define RP_THREADS 256
#ifndef RP_ITEMS
#define RP_ITEMS 16
#endif
#define RP_TILE (RP_THREADS * RP_ITEMS)
#define RP_WARPS (RP_THREADS / 32)
#define RH_THREADS 256
#define RH_TILE (RH_THREADS * 32)
// MATCH.ANY runs on the narrow ADU pipe (~2 cycles per distinct value in the warp, ~60 cycles for
// random digits); eight ballots on the bits of the digit give the same mask from the ALU side.
// RP_HW_MATCH of the RP_ITEMS items use the hardware instruction so that both pipes stay busy.
#ifndef RP_HW_MATCH
#define RP_HW_MATCH 4
#endif
template <bool HW>
__device__ __forceinline__ u32 match_digit(u32 d) {
  if (HW) return __match_any_sync(FULL_MASK, d);
  u32 m = FULL_MASK;
#pragma unroll
  for (int b = 0; b < RADIX_BITS; b++) {
    const bool bit = (d >> b) & 1u;
    const u32 bal = __ballot_sync(FULL_MASK, bit);
    m &= bit ? bal : ~bal;
  }
  return m;
}

// status word of the per-digit chained scan: 2 flag bits + 30-bit count
#define RS_AGG 0x40000000u
#define RS_PREFIX 0x80000000u
#define RS_FLAGS 0xC0000000u
#define RS_VALUE 0x3FFFFFFFu

template <typename KeyT>
__global__ void __launch_bounds__(RH_THREADS)
k_radix_hist(const KeyT* __restrict__ keys, const u32* __restrict__ seg_n, u32 tiles_per_seg, u32 seg_shift,
             u32* __restrict__ hist, u32 npass, u32 begin_bit) {
  __shared__ u32 h[8 * RADIX];
  for (u32 i = threadIdx.x; i < npass * RADIX; i += RH_THREADS) h[i] = 0;
  __syncthreads();
  const u32 seg = blockIdx.x / tiles_per_seg, lt = blockIdx.x % tiles_per_seg;
  const u32 n = seg_n[seg];
  const u32 start = lt * RH_TILE;
  if (start >= n) return;
  const u32 count = min((u32)RH_TILE, n - start);
  const KeyT* p = keys + ((size_t)seg << seg_shift) + start;
  for (u32 i = threadIdx.x; i < count; i += RH_THREADS) {
    const KeyT k = p[i];
    for (u32 ps = 0; ps < npass; ps++) atomicAdd(&h[ps * RADIX + ((u32)(k >> (begin_bit + ps * RADIX_BITS)) & (RADIX - 1))], 1u);
  }
  __syncthreads();
  for (u32 i = threadIdx.x; i < npass * RADIX; i += RH_THREADS)
    if (h[i]) atomicAdd(&hist[(size_t)seg * npass * RADIX + i], h[i]);
}

template <typename KeyT, bool HAS_VALS>
struct RadixSmem {
  KeyT key[RP_TILE];
  u32 val[HAS_VALS ? RP_TILE : 1];
  u32 whist[RP_WARPS][RADIX];
  u32 excl[RADIX];   // exclusive prefix of the digit totals inside this tile
  int gbase[RADIX];  // global destination of bucket d's first element minus excl[d]
  u32 ws[RP_WARPS + 1];
  u32 tile;
};

// hist_stride / hist_pass_stride let several passes share one histogram (the 4-byte-prefix keys of
// all rotations have the same digit histogram in every pass: the block's byte histogram).
// iota != 0: values are synthesised as the global slot index (first pass of the initial sort).
// HAS_VALS = false: keys-only sort (the initial BWT sort packs (4-byte prefix << 32 | suffix id) into
// one 64-bit record, so every record is a single 8-byte load, stage and store).
template <typename KeyT, bool HAS_VALS>
__global__ void __launch_bounds__(RP_THREADS)
k_radix_pass(const KeyT* __restrict__ kin, const u32* __restrict__ vin, KeyT* __restrict__ kout,
             u32* __restrict__ vout, const u32* __restrict__ seg_n, u32 tiles_per_seg, u32 seg_shift,
             const u32* __restrict__ hist, u32 hist_seg_stride, u32 hist_off, u32 shift, u32* ticket, u32* status, int iota,
             const u8* __restrict__ pack_L = nullptr, u32* __restrict__ pack_P = nullptr) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  RadixSmem<KeyT, HAS_VALS>& s = *reinterpret_cast<RadixSmem<KeyT, HAS_VALS>*>(smem_raw);
  const u32 tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  if (tid == 0) s.tile = atomicAdd(ticket, 1u);
#pragma unroll
  for (u32 i = 0; i < RP_WARPS; i++) s.whist[i][tid] = 0;
  __syncthreads();
  const u32 tile = s.tile;
  const u32 seg = tile / tiles_per_seg, lt = tile - seg * tiles_per_seg;
  const u32 n = seg_n[seg];
  const u32 start = lt * RP_TILE;
  if (start >= n) return;  // every later tile of this segment is empty too: nobody waits on us
  const u32 count = min((u32)RP_TILE, n - start);
  const size_t base = ((size_t)seg << seg_shift) + start;
  const KeyT* kp = kin + base + w * (32 * RP_ITEMS) + lane;
  const u32 woff = w * (32 * RP_ITEMS) + lane;  // tile offset of this thread's item 0
  u32* wh = s.whist[w];

  KeyT key[RP_ITEMS];
  u32 rnk[RP_ITEMS / 2];  // two 16-bit in-bucket ranks per register (a tile holds 4096 records)
  if (count == RP_TILE) {
    // ---- full tile: no per-item predicates ----
#pragma unroll
    for (int k = 0; k < RP_ITEMS; k++) key[k] = kp[k * 32];
#pragma unroll
    for (int k = 0; k < RP_ITEMS; k++) {
      const u32 d = (u32)(key[k] >> shift) & (RADIX - 1);
      const u32 m = (k % (RP_ITEMS / (RP_HW_MATCH ? RP_HW_MATCH : 1)) == 0 && RP_HW_MATCH) ? match_digit<true>(d) : match_digit<false>(d);
      const u32 leader = 31 - __clz(m);
      const u32 before = __popc(m & lanemask_lt());
      u32 prev = 0;
      if (lane == leader) { prev = wh[d]; wh[d] = prev + before + 1; }  // leader is the highest lane: before+1 == popc(m)
      prev = __shfl_sync(FULL_MASK, prev, leader);
      const u32 r = prev + before;
      if (k & 1) rnk[k >> 1] |= r << 16; else rnk[k >> 1] = r;
    }
  } else {
#pragma unroll
    for (int k = 0; k < RP_ITEMS; k++) key[k] = (woff + k * 32 < count) ? kp[k * 32] : (KeyT)0;
#pragma unroll
    for (int k = 0; k < RP_ITEMS; k++) {
      const bool valid = woff + k * 32 < count;
      const u32 d = valid ? ((u32)(key[k] >> shift) & (RADIX - 1)) : RADIX;
      const u32 m = __match_any_sync(FULL_MASK, d);
      const u32 leader = 31 - __clz(m);
      const u32 before = __popc(m & lanemask_lt());
      u32 prev = 0;
      if (valid && lane == leader) { prev = wh[d]; wh[d] = prev + before + 1; }
      prev = __shfl_sync(FULL_MASK, prev, leader);
      const u32 r = prev + before;
      if (k & 1) rnk[k >> 1] |= r << 16; else rnk[k >> 1] = r;
    }
  }
  __syncthreads();
  // ---- per digit: prefix over warps, tile totals, digit scan, chained scan over tiles ----
  {
    const u32 d = tid;  // RP_THREADS == RADIX
    u32 acc = 0;
#pragma unroll
    for (int ww = 0; ww < RP_WARPS; ww++) {
      const u32 c = s.whist[ww][d];
      s.whist[ww][d] = acc;
      acc += c;
    }
    const u32 total = acc;
    u32 dummy;
    const u32 ex = block_excl_add<RP_THREADS, u32>(total, s.ws, &dummy);
    s.excl[d] = ex;
    // bucket start inside the segment, from the up-front histogram
    const u32 hcount = hist[(size_t)seg * hist_seg_stride + hist_off + d];
    const u32 hbase = block_excl_add<RP_THREADS, u32>(hcount, s.ws, &dummy);
    // decoupled look-back over the earlier tiles of this segment
    u32* st = status + (size_t)tile * RADIX + d;
    u32 excl_tiles = 0;
    if (lt == 0) {
      st_volatile_u32(st, RS_PREFIX | total);
    } else {
      st_volatile_u32(st, RS_AGG | total);
      const u32* look = st - RADIX;
      while (true) {
        u32 v;
        do {
          v = ld_volatile_u32(look);
        } while ((v & RS_FLAGS) == 0);
        excl_tiles += v & RS_VALUE;
        if (v & RS_PREFIX) break;
        look -= RADIX;
      }
      st_volatile_u32(st, RS_PREFIX | (excl_tiles + total));
    }
    s.gbase[d] = (int)(hbase + excl_tiles) - (int)ex;
  }
  __syncthreads();
  // ---- stage the tile in bucket order (values are fetched only now: fewer live registers) ----
  {
    const u32* vp = vin + base + woff;
    const u32 vbase = (u32)base + woff;
#pragma unroll
    for (int k = 0; k < RP_ITEMS; k++) {
      if (count == RP_TILE || woff + k * 32 < count) {
        const u32 d = (u32)(key[k] >> shift) & (RADIX - 1);
        const u32 r = (k & 1) ? (rnk[k >> 1] >> 16) : (rnk[k >> 1] & 0xffffu);
        const u32 p = s.excl[d] + wh[d] + r;
        s.key[p] = key[k];
        if (HAS_VALS) s.val[p] = iota ? (vbase + k * 32) : vp[k * 32];
      }
    }
  }
  __syncthreads();
  // ---- write every bucket run out coalesced ----
  KeyT* ko = kout + ((size_t)seg << seg_shift);
  u32* vo = vout + ((size_t)seg << seg_shift);
#pragma unroll
  for (int k = 0; k < RP_ITEMS; k++) {
    const u32 p = k * RP_THREADS + tid;
    if (count == RP_TILE || p < count) {
      const KeyT kk = s.key[p];
      const u32 d = (u32)(kk >> shift) & (RADIX - 1);
      const u32 dst = (u32)((int)p + s.gbase[d]);
      if (HAS_VALS && pack_P) {
        // inverse-BWT epilogue (decode.cu): the record that sorts to row `dst` is the successor pointer of that row;
        // P[row] = successor << 8 | L[row] (lib/Bzip2.js:370-381) -- the sorted keys and values themselves are not needed
        const size_t row = ((size_t)seg << seg_shift) + dst;
        pack_P[row] = ((s.val[p] & ((1u << seg_shift) - 1u)) << 8) | pack_L[row];
      } else {
        ko[dst] = kk;
        if (HAS_VALS) vo[dst] = s.val[p];
      }
    }
  }
}
