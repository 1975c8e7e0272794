// radix_host.cuh -- host driver of the segmented onesweep radix sort (radix.cuh).
#pragma once
#include "ctx.h"
#include "radix.cuh"

// shared_hist != nullptr: a ready [nseg][256] histogram that is valid for EVERY pass (initial BWT sort);
// otherwise the per-pass histograms are computed by k_radix_hist from the keys.
template <typename KeyT, bool HAS_VALS = true>
static void radix_sort(Ctx& c, KeyT*& kin, u32*& vin, KeyT*& kout, u32*& vout, const u32* d_seg_n, u32 nseg, u32 seg_shift,
                       u32 max_seg_n, u32 begin_bit, u32 npass, bool iota_first, u64 total_elems, const u32* shared_hist = nullptr,
                       const u8* pack_L = nullptr, u32* pack_P = nullptr) {
  if (npass == 0 || total_elems == 0) return;
  static bool attr_set = false;  // per translation unit (kernels are instantiated per TU)
  if (!attr_set) {
    CUDA_CHECK(cudaFuncSetAttribute(k_radix_pass<KeyT, HAS_VALS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)sizeof(RadixSmem<KeyT, HAS_VALS>)));
    attr_set = true;
  }
  const u32 tps = (max_seg_n + RP_TILE - 1) / RP_TILE;
  const u32 htps = (max_seg_n + RH_TILE - 1) / RH_TILE;
  const size_t ntiles = (size_t)tps * nseg;
  DBuf<u32> hist, status(c, ntiles * RADIX * npass), ticket(c, npass);
  if (!shared_hist) {
    hist.alloc(c, (size_t)nseg * npass * RADIX);
    CUDA_CHECK(cudaMemsetAsync(hist, 0, (size_t)nseg * npass * RADIX * 4, c.stream));
    k_radix_hist<KeyT><<<htps * nseg, RH_THREADS, 0, c.stream>>>(kin, d_seg_n, htps, seg_shift, hist, npass, begin_bit);
    KLAUNCH(c); KCHECK();
    c.stats.bwt_bytes += total_elems * sizeof(KeyT);
  }
  // one memset for the look-back state of all passes
  CUDA_CHECK(cudaMemsetAsync(status, 0, ntiles * RADIX * npass * 4, c.stream));
  CUDA_CHECK(cudaMemsetAsync(ticket, 0, 4 * npass, c.stream));
  for (u32 p = 0; p < npass; p++) {
    const int iota = (iota_first && p == 0) ? 1 : 0;
    size_t ev = c.begin(ST_RADIX);
    k_radix_pass<KeyT, HAS_VALS><<<(unsigned)ntiles, RP_THREADS, sizeof(RadixSmem<KeyT, HAS_VALS>), c.stream>>>(
        kin, vin, kout, vout, d_seg_n, tps, seg_shift, shared_hist ? shared_hist : hist.p, shared_hist ? RADIX : npass * RADIX,
        shared_hist ? 0 : p * RADIX, begin_bit + p * RADIX_BITS, ticket.p + p, status.p + (size_t)p * ntiles * RADIX, iota,
        (p + 1 == npass) ? pack_L : nullptr, (p + 1 == npass) ? pack_P : nullptr);
    c.end(ev);
    KLAUNCH(c); KCHECK();
    const u64 bytes = total_elems * (2 * sizeof(KeyT) + (HAS_VALS ? (iota ? 4 : 8) : 0));
    c.stats.radix_launches++;
    c.stats.radix_bytes += bytes;
    c.stats.bwt_bytes += bytes;
    std::swap(kin, kout);
    std::swap(vin, vout);
  }
  // result is in (kin, vin) after the swaps
}
