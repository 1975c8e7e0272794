// radix.cuh -- segmented LSD radix sort (8-bit digits), single pass over the data per digit:
//   k_radix_hist : one read of the keys -> per-(segment, pass) digit histograms
//   k_radix_pass : "onesweep" pass = read (key,val) once, rank inside the tile with
//                  warp match/shuffle, chain the per-digit tile offsets with a decoupled
//                  look-back, stage the tile in shared memory in bucket order and write
//                  every bucket run out coalesced.
// Segments are the independent bzip2 blocks of a batch (slot = seg << seg_shift | i), or a
// single flat segment for the compacted "still unsorted" suffixes of a doubling round.
// This is the hot kernel of the forward BWT (reference: lib/BWT.js:372-417 does the same
// job with SA-IS on one block at a time).
#pragma once
#include "common.cuh"

#define RADIX_BITS 8
#define RADIX 256
#define RP_THREADS 256
#define RP_ITEMS 16
#define RP_TILE (RP_THREADS * RP_ITEMS)
#define RP_WARPS (RP_THREADS / 32)
#define RH_THREADS 256
#define RH_TILE (RH_THREADS * 32)

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
  for (u32 i0 = 0; i0 < count; i0 += RH_THREADS) {  // warp-uniform trip count
    const u32 i = i0 + threadIdx.x;
    const bool valid = i < count;
    KeyT k = valid ? p[i] : (KeyT)0;
    for (u32 ps = 0; ps < npass; ps++) {
      u32 d = valid ? ((u32)(k >> (begin_bit + ps * RADIX_BITS)) & (RADIX - 1)) : RADIX;
      // aggregate equal digits inside the warp first: constant digits (high key bits) would
      // otherwise serialise 32-way on one shared-memory counter
      u32 m = __match_any_sync(FULL_MASK, d);
      if (valid && (m & lanemask_lt()) == 0) atomicAdd(&h[ps * RADIX + d], (u32)__popc(m));
    }
  }
  __syncthreads();
  for (u32 i = threadIdx.x; i < npass * RADIX; i += RH_THREADS)
    if (h[i]) atomicAdd(&hist[(size_t)seg * npass * RADIX + i], h[i]);
}

template <typename KeyT>
struct RadixSmem {
  KeyT key[RP_TILE];
  u32 val[RP_TILE];
  u32 whist[RP_WARPS][RADIX];
  u32 excl[RADIX];   // exclusive prefix of the digit totals inside this tile
  int gbase[RADIX];  // global destination of bucket d's first element minus excl[d]
  u32 ws[RP_WARPS + 1];
  u32 tile;
};

// iota != 0: values are synthesised as the global slot index (first pass of the initial sort).
template <typename KeyT>
__global__ void __launch_bounds__(RP_THREADS)
k_radix_pass(const KeyT* __restrict__ kin, const u32* __restrict__ vin, KeyT* __restrict__ kout,
             u32* __restrict__ vout, const u32* __restrict__ seg_n, u32 tiles_per_seg, u32 seg_shift,
             const u32* __restrict__ hist, u32 npass, u32 pass, u32 shift, u32* ticket, u32* status, int iota) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  RadixSmem<KeyT>& s = *reinterpret_cast<RadixSmem<KeyT>*>(smem_raw);
  const u32 tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  if (tid == 0) s.tile = atomicAdd(ticket, 1u);
  for (u32 i = tid; i < RP_WARPS * RADIX; i += RP_THREADS) (&s.whist[0][0])[i] = 0;
  __syncthreads();
  const u32 tile = s.tile;
  const u32 seg = tile / tiles_per_seg, lt = tile % tiles_per_seg;
  const u32 n = seg_n[seg];
  const u32 start = lt * RP_TILE;
  if (start >= n) return;  // every later tile of this segment is empty too: nobody waits on us
  const u32 count = min((u32)RP_TILE, n - start);
  const size_t base = ((size_t)seg << seg_shift) + start;

  // ---- load, warp-striped (coalesced) ----
  KeyT key[RP_ITEMS];
  u32 val[RP_ITEMS];
  u32 rnk[RP_ITEMS];
#pragma unroll
  for (int k = 0; k < RP_ITEMS; k++) {
    u32 off = w * (32 * RP_ITEMS) + k * 32 + lane;
    bool valid = off < count;
    key[k] = valid ? kin[base + off] : (KeyT)0;
    val[k] = valid ? (iota ? (u32)(base + off) : vin[base + off]) : 0u;
  }
  // ---- rank inside the warp: match + popc, warp-private digit counters ----
#pragma unroll
  for (int k = 0; k < RP_ITEMS; k++) {
    u32 off = w * (32 * RP_ITEMS) + k * 32 + lane;
    bool valid = off < count;
    u32 d = valid ? ((u32)(key[k] >> shift) & (RADIX - 1)) : RADIX;
    u32 m = __match_any_sync(FULL_MASK, d);
    u32 before = __popc(m & lanemask_lt());
    u32 prev = 0;
    if (valid) prev = s.whist[w][d];
    __syncwarp();
    if (valid && before == 0) s.whist[w][d] = prev + __popc(m);
    __syncwarp();
    rnk[k] = prev + before;
  }
  __syncthreads();
  // ---- per digit: prefix over warps, tile totals, digit scan, chained scan over tiles ----
  {
    const u32 d = tid;  // RP_THREADS == RADIX
    u32 acc = 0;
#pragma unroll
    for (int ww = 0; ww < RP_WARPS; ww++) {
      u32 c = s.whist[ww][d];
      s.whist[ww][d] = acc;
      acc += c;
    }
    const u32 total = acc;
    u32 dummy;
    u32 ex = block_excl_add<RP_THREADS, u32>(total, s.ws, &dummy);
    s.excl[d] = ex;
    // bucket start inside the segment, from the up-front histogram
    u32 hcount = hist[((size_t)seg * npass + pass) * RADIX + d];
    u32 hbase = block_excl_add<RP_THREADS, u32>(hcount, s.ws, &dummy);
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
  // ---- stage the tile in bucket order ----
#pragma unroll
  for (int k = 0; k < RP_ITEMS; k++) {
    u32 off = w * (32 * RP_ITEMS) + k * 32 + lane;
    if (off < count) {
      u32 d = (u32)(key[k] >> shift) & (RADIX - 1);
      u32 p = s.excl[d] + s.whist[w][d] + rnk[k];
      s.key[p] = key[k];
      s.val[p] = val[k];
    }
  }
  __syncthreads();
  // ---- write every bucket run out coalesced ----
  const size_t segbase = (size_t)seg << seg_shift;
#pragma unroll
  for (int k = 0; k < RP_ITEMS; k++) {
    u32 p = k * RP_THREADS + tid;
    if (p < count) {
      KeyT kk = s.key[p];
      u32 d = (u32)(kk >> shift) & (RADIX - 1);
      size_t dst = segbase + (size_t)((int)p + s.gbase[d]);
      kout[dst] = kk;
      vout[dst] = s.val[p];
    }
  }
}
