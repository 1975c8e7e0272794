// mtf.cu -- symbol map, move-to-front and zero-run (RUNA/RUNB) coding of the BWT output.
//
// Reference: lib/Bzip2.js:743-815 (compressBlock: used[] map, MTF list M, RLE2 emit) and
// lib/Bzip2.js:53-60 (mtf()).  The reference walks the block byte by byte with a linear
// search in M.  Parallel form:
//   k_used        : 256-bit "byte occurs in block" map per block
//   k_mtf_lastpos : per 4 KiB chunk, last position of every byte value inside the chunk
//   k_mtf_prefix  : per block, running max over the chunks -> last occurrence BEFORE each chunk;
//                   the MTF list at a chunk start is "bytes by most recent occurrence, then the
//                   not-yet-seen used bytes in ascending order" (the initial list M)
//   k_mtf_ranks   : one warp per chunk, 32 bytes per step: every byte value carries a recency key
//                   (255 - rank at the chunk start, or 256 + position of its last occurrence inside
//                   the chunk); the rank of a byte is the number of live keys above its own, counted
//                   through a bucketed live-key bitmap with suffix sums, and the bytes of one step
//                   that precede each other are settled with SWAR pair compares
//                   The same pass summarises the chunk for the zero-run coder: leading / trailing zeros and the number
//                   of symbols its non-zero ranks and interior runs will emit.
//   k_rle2_scan   : one warp per block walks the chunk summaries: output offset and carried-in run length of every chunk
//                   (a run belongs to the chunk that holds the non-zero rank ending it); final run + EOB + m
//   k_rle2        : zero ranks form runs -> bijective base-2 RUNA/RUNB digits (lib/Bzip2.js:783-794); every chunk knows
//                   its offsets, so there is no chain between tiles; symbols u16 + histogram
#include "enc.h"

#define MTF_CHUNK 4096

__global__ void __launch_bounds__(256) k_used(const u8* __restrict__ U, const u32* __restrict__ seg_n, u32 tiles_per_seg, u32* __restrict__ used) {
  __shared__ u32 f[8];
  if (threadIdx.x < 8) f[threadIdx.x] = 0;
  __syncthreads();
  const u32 seg = blockIdx.x / tiles_per_seg, lt = blockIdx.x % tiles_per_seg;
  const u32 n = seg_n[seg];
  const u32 start = lt * (256 * 64);
  if (start >= n) return;
  const u8* p = U + ((size_t)seg << SEG_SHIFT);
  u32 loc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (u32 i = start + threadIdx.x; i < min(n, start + 256 * 64); i += 256) {
    u8 c = p[i];
    loc[c >> 5] |= 1u << (c & 31);
  }
#pragma unroll
  for (int k = 0; k < 8; k++)
    if (loc[k]) atomicOr(&f[k], loc[k]);
  __syncthreads();
  if (threadIdx.x < 8 && f[threadIdx.x]) atomicOr(&used[seg * 8 + threadIdx.x], f[threadIdx.x]);
}

__global__ void __launch_bounds__(256) k_used_from_hist(const u32* __restrict__ hist, u32* __restrict__ used) {
  const u32 bal = __ballot_sync(FULL_MASK, hist[blockIdx.x * 256 + threadIdx.x] != 0);
  if ((threadIdx.x & 31) == 0) used[blockIdx.x * 8 + (threadIdx.x >> 5)] = bal;
}

// lastpos[(seg*cps + chunk)*256 + c] = (last position of byte c inside the chunk) + 1, 0 if none
__global__ void __launch_bounds__(256) k_mtf_lastpos(const u8* __restrict__ U, const u32* __restrict__ seg_n, u32 cps, u32* __restrict__ lastpos) {
  __shared__ u32 last[256];
  last[threadIdx.x] = 0;
  __syncthreads();
  const u32 seg = blockIdx.x / cps, ch = blockIdx.x % cps;
  const u32 n = seg_n[seg];
  const u32 start = ch * MTF_CHUNK;
  if (start < n) {
    const u8* p = U + ((size_t)seg << SEG_SHIFT);
    const u32 end = min(n, start + MTF_CHUNK);
    for (u32 i = start + threadIdx.x; i < end; i += 256) atomicMax(&last[p[i]], i + 1);
  }
  __syncthreads();
  lastpos[(size_t)blockIdx.x * 256 + threadIdx.x] = last[threadIdx.x];
}

// in place: lastpos[chunk] := max over earlier chunks (exclusive)
__global__ void __launch_bounds__(256) k_mtf_prefix(const u32* __restrict__ seg_n, u32 cps, u32* __restrict__ lastpos) {
  const u32 seg = blockIdx.x;
  const u32 n = seg_n[seg];
  const u32 nch = (n + MTF_CHUNK - 1) / MTF_CHUNK;
  u32 run = 0;
  for (u32 ch = 0; ch < nch; ch++) {
    u32* p = lastpos + ((size_t)seg * cps + ch) * 256 + threadIdx.x;
    u32 t = *p;
    *p = run;
    run = max(run, t);
  }
}

#define MR_WARPS 8
#define MB_BUCKETS 36   // recency keys live in [0, 256 + 4096): 34 buckets of 128 values (+ slack)
struct MtfWarp {
  u32 skey[256];              // scratch for the start-of-chunk ranking
  u16 K[256];                 // recency key of every byte value (larger = used more recently)
  u32 bm[MB_BUCKETS][4];      // bitmap of the key values that are currently somebody's key
  u32 S[MB_BUCKETS];          // S[b] = number of live keys in buckets above b
  u32 Pw[16];                 // the window's previous-use times, two 16-bit values per word
};
// MTF rank of a byte = number of byte values used more recently than it.  Every byte value carries
// a 15-bit recency key: 255 - (list position at the chunk start) until it is used inside the chunk,
// 256 + (position inside the chunk) afterwards.  A warp ranks 32 bytes per step:
//   * P_i = time of the previous use of lane i's byte (an earlier lane of the window, or its key)
//   * rank_i = #{keys alive before the window that are > P_i}        (bucket suffix sums + bitmap)
//            + #{k in (prev_i, i) : P_k < P_i}                       (first use after P_i inside the window)
//   * only the last use of a byte inside the window rewrites its key.
__global__ void __launch_bounds__(MR_WARPS * 32)
k_mtf_ranks(const u8* __restrict__ U, const u32* __restrict__ seg_n, u32 cps, const u32* __restrict__ lastpos,
            const u32* __restrict__ used, u8* __restrict__ R, u32 nblk, uint4* __restrict__ rsum) {
  __shared__ MtfWarp sm[MR_WARPS];
  const u32 w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const u32 gchunk = blockIdx.x * MR_WARPS + w;
  const u32 seg = gchunk / cps, ch = gchunk % cps;
  // (every warp of the grid maps to a valid (seg, chunk) pair or exits as a whole)
  if (seg >= nblk) return;
  const u32 n = seg_n[seg];
  const u32 start = ch * MTF_CHUNK;
  if (start >= n) return;
  const u32 count = min((u32)MTF_CHUNK, n - start);
  MtfWarp& s = sm[w];
  // ---- list position of every byte value at the chunk start (rank by counting) ----
  const u32* lp = lastpos + (size_t)gchunk * 256;
  u32 mykey[8];
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const u32 c = lane * 8 + i;
    const u32 l = lp[c];
    const bool isused = (used[seg * 8 + (c >> 5)] >> (c & 31)) & 1;
    u32 key;
    if (l) key = 0x40000000u | l;          // seen: most recent first
    else if (isused) key = 0x200u + (255u - c);  // not seen yet: ascending byte value
    else key = 255u - c;                    // never occurs: behind everything
    mykey[i] = key;
    s.skey[c] = key;
  }
  __syncwarp();
  u32 rk[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (u32 c2 = 0; c2 < 256; c2++) {
    const u32 k2 = s.skey[c2];
#pragma unroll
    for (int i = 0; i < 8; i++) rk[i] += (k2 > mykey[i]) ? 1u : 0u;
  }
#pragma unroll
  for (int i = 0; i < 8; i++) s.K[lane * 8 + i] = (u16)(255u - rk[i]);
  // all 256 start keys 0..255 are alive: buckets 0 and 1 full, the rest empty
  for (u32 i = lane; i < MB_BUCKETS * 4; i += 32) (&s.bm[0][0])[i] = (i < 8) ? 0xffffffffu : 0u;
  __syncwarp();
  const u8* src = U + ((size_t)seg << SEG_SHIFT) + start;
  u8* dst = R + ((size_t)seg << SEG_SHIFT) + start;
  const u32 H = 0x80008000u, ONE = 0x00010001u;
  // zero-run summary of the chunk (warp-uniform state + a per-lane count of emitted symbols)
  u32 z_open = 0, z_lead = 0, z_acc = 0;
  bool z_seen = false;
  for (u32 base = 0; base < count; base += 32) {
    const bool valid = base + lane < count;
    const u32 c = valid ? (u32)src[base + lane] : (256u + lane);
    // ---- bucket suffix sums of the keys alive before this window ----
    if (lane < MB_BUCKETS) {
      // handled below with a full-warp reverse scan (MB_BUCKETS > 32 -> two steps)
    }
    {
      u32 c0 = __popc(s.bm[lane][0]) + __popc(s.bm[lane][1]) + __popc(s.bm[lane][2]) + __popc(s.bm[lane][3]);
      u32 c1 = 0;
      if (lane < MB_BUCKETS - 32) c1 = __popc(s.bm[32 + lane][0]) + __popc(s.bm[32 + lane][1]) + __popc(s.bm[32 + lane][2]) + __popc(s.bm[32 + lane][3]);
      // inclusive suffix sums over lanes (high lanes first)
      u32 hi = c1;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const u32 t = __shfl_down_sync(FULL_MASK, hi, o); if (lane + o < 32) hi += t; }
      const u32 tot_hi = __shfl_sync(FULL_MASK, hi, 0);
      u32 lo = c0;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const u32 t = __shfl_down_sync(FULL_MASK, lo, o); if (lane + o < 32) lo += t; }
      s.S[lane] = lo - c0 + tot_hi;
      if (lane < MB_BUCKETS - 32) s.S[32 + lane] = hi - c1;
    }
    // ---- who used my byte last? ----
    const u32 m = __match_any_sync(FULL_MASK, c);
    const u32 pm = m & lanemask_lt();
    const int prev = pm ? (31 - __clz(pm)) : -1;
    const bool is_last = (m >> lane) <= 1u;  // no higher lane holds the same byte
    const u32 tb = 256u + base;
    const u32 q = valid ? (u32)s.K[c] : 0u;
    const u32 Pi = prev >= 0 ? (tb + (u32)prev) : q;
    // publish P (two lanes per word)
    {
      const u32 other = __shfl_down_sync(FULL_MASK, Pi, 1);
      if (!(lane & 1)) s.Pw[lane >> 1] = Pi | (other << 16);
    }
    __syncwarp();
    // ---- keys alive before the window and more recent than P_i (only when P_i is such a key) ----
    u32 rank = 0;
    if (prev < 0 && valid) {
      const u32 b = q >> 7, off = q & 127u, wi = off >> 5;
      rank = s.S[b];
      const u32 w0 = s.bm[b][0], w1 = s.bm[b][1], w2 = s.bm[b][2], w3 = s.bm[b][3];
      const u32 cur = wi == 0 ? w0 : (wi == 1 ? w1 : (wi == 2 ? w2 : w3));
      rank += __popc(cur & ((0xfffffffeu) << (off & 31u)));
      if (wi < 1) rank += __popc(w1);
      if (wi < 2) rank += __popc(w2);
      if (wi < 3) rank += __popc(w3);
    }
    // ---- first uses after P_i inside the window: k in (prev_i, i) with P_k < P_i ----
    {
      const u32 PiPi = ((Pi * ONE) | H) - ONE;          // (P_i | 0x8000) - 1 in both halves
      const u32 lanes2 = ((lane * ONE) | H) - ONE;        // for k < i
      const u32 pe = (u32)(prev + 1) * ONE;               // for k >= prev + 1
      u32 acc = 0;
#pragma unroll
      for (u32 j = 0; j < 16; j++) {
        const u32 Pk = s.Pw[j];
        const u32 kk = (2 * j) | ((2 * j + 1) << 16);
        const u32 z1 = PiPi - Pk;                         // bit15/31: P_k < P_i
        const u32 z2 = lanes2 - kk;                       // k < i
        const u32 z3 = ((kk + ONE) | H) - ONE - pe;       // k + 1 > prev + 1 - 1 ... k >= prev + 1
        acc |= ((z1 & z2 & z3) & H) >> j;
      }
      rank += __popc(acc);
    }
    if (valid) dst[base + lane] = (u8)rank;
    {
      const u32 nzm = __ballot_sync(FULL_MASK, valid && rank != 0);
      const u32 nvalid = min(32u, count - base);
      if (valid && rank != 0) {
        const u32 below = nzm & lanemask_lt();
        if (below) {
          const u32 L = lane - (31 - __clz(below)) - 1;          // zeros since the previous non-zero of this step
          z_acc += 1 + (L ? 31 - __clz(L + 1) : 0);
        } else if (z_seen) {
          const u32 L = z_open + lane;
          z_acc += 1 + (L ? 31 - __clz(L + 1) : 0);
        } else {
          z_lead = z_open + lane;                                 // the run that reaches back to the chunk start is not ours to count
          z_acc += 1;
        }
      }
      if (nzm) { z_seen = true; z_open = nvalid - 1 - (31 - __clz(nzm)); }
      else z_open += nvalid;
    }
    // ---- the last use of every byte in the window rewrites its key ----
    if (valid && is_last) {
      atomicAnd(&s.bm[q >> 7][(q & 127u) >> 5], ~(1u << (q & 31u)));
      s.K[c] = (u16)(tb + lane);
    }
    const u32 newbits = __ballot_sync(FULL_MASK, valid && is_last);
    __syncwarp();
    if (lane == 0) s.bm[tb >> 7][(tb & 127u) >> 5] = newbits;  // windows are 32-aligned: this word is ours alone
    __syncwarp();
  }
  {
    const u32 inner = warp_reduce_add(z_acc);
    const u32 lead = z_seen ? warp_reduce_max(z_lead) : count;
    if (lane == 0) rsum[gchunk] = make_uint4(lead, z_seen ? z_open : count, inner, z_seen ? 0u : 1u);
  }
}

// One warp per block: walk the chunk summaries in order (lib/Bzip2.js:783-794 flushes a run when the next non-zero
// rank arrives, so the run is charged to the chunk holding that rank).  plan[chunk] = (output offset, zeros carried in).
__global__ void __launch_bounds__(32)
k_rle2_scan(const uint4* __restrict__ rsum, const u32* __restrict__ seg_n, u32 cps, const u32* __restrict__ used, uint2* __restrict__ plan,
            u16* __restrict__ A, u32* __restrict__ m_out, u32* __restrict__ freq) {
  const u32 seg = blockIdx.x, lane = threadIdx.x;
  const u32 n = seg_n[seg];
  if (n == 0) return;
  const u32 nch = (n + MTF_CHUNK - 1) / MTF_CHUNK;
  u32 carry = 0, o = 0;
  for (u32 c0 = 0; c0 < nch; c0 += 32) {
    const u32 k = c0 + lane;
    uint4 v = make_uint4(0, 0, 0, 1);
    if (k < nch) v = rsum[(size_t)seg * cps + k];
    u32 my_o = 0, my_c = 0;
    const u32 lim = min(32u, nch - c0);
    for (u32 j = 0; j < lim; j++) {
      const u32 lead = __shfl_sync(FULL_MASK, v.x, j), trail = __shfl_sync(FULL_MASK, v.y, j);
      const u32 inner = __shfl_sync(FULL_MASK, v.z, j), allz = __shfl_sync(FULL_MASK, v.w, j);
      if (lane == j) { my_o = o; my_c = carry; }
      if (allz) carry += lead;
      else {
        const u32 r = carry + lead;
        o += (r ? 31 - __clz(r + 1) : 0) + inner;
        carry = trail;
      }
    }
    if (k < nch) plan[(size_t)seg * cps + k] = make_uint2(my_o, my_c);
  }
  if (lane == 0) {
    u16* a = A + ((size_t)seg << SEG_SHIFT);
    u32* fq = freq + (size_t)seg * HUFF_MAXSYM;
    u32 L = carry, f0 = 0, f1 = 0;
    while (L) {  // the run still open at the end of the block
      if (L & 1) { a[o++] = 0; f0++; L -= 1; }
      else { a[o++] = 1; f1++; L -= 2; }
      L >>= 1;
    }
    u32 alpha = 0;
    for (int k = 0; k < 8; k++) alpha += __popc(used[seg * 8 + k]);
    a[o] = (u16)(alpha + 1);  // end of block symbol
    if (f0) atomicAdd(&fq[0], f0);
    if (f1) atomicAdd(&fq[1], f1);
    atomicAdd(&fq[alpha + 1], 1u);
    m_out[seg] = o + 1;
  }
}

// ---- RLE2 ---------------------------------------------------------------------------------
#define R2_THREADS 256
#define R2_ITEMS 16
#define R2_TILE (R2_THREADS * R2_ITEMS)   // == MTF_CHUNK: one tile per chunk summary

__global__ void __launch_bounds__(R2_THREADS)
k_rle2(const u8* __restrict__ R, const u32* __restrict__ seg_n, u32 tps, const uint2* __restrict__ plan, u16* __restrict__ A,
       u32* __restrict__ freq) {
  __shared__ u32 hist[HUFF_MAXSYM];
  __shared__ u32 ws[R2_THREADS / 32 + 1];
  // the tile's symbols are staged here at the alignment (mod 8 symbols = 16 bytes) they have in global memory, then
  // copied out in 16-byte pieces: at most one symbol per rank plus the digits of the run that was carried in
  __shared__ __align__(16) u16 stage[R2_TILE + 48];
  const u32 tid = threadIdx.x;
  const u32 seg = blockIdx.x / tps, lt = blockIdx.x % tps;
  const u32 n = seg_n[seg];
  const u32 start = lt * R2_TILE;
  if (start >= n) return;
  for (u32 i = tid; i < HUFF_MAXSYM; i += R2_THREADS) hist[i] = 0;
  const uint2 pl = plan[(size_t)seg * tps + lt];
  const u8* r = R + ((size_t)seg << SEG_SHIFT);
  u16* a = A + ((size_t)seg << SEG_SHIFT);
  const u32 p0 = start + tid * R2_ITEMS;
  u8 v[R2_ITEMS];
  {
    const uint4 x = *reinterpret_cast<const uint4*>(r + p0);  // inside the 1 MiB slot; bytes past n are ignored below
    const u32 xw[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
    for (int j = 0; j < R2_ITEMS; j++) v[j] = (u8)(xw[j >> 2] >> (8 * (j & 3)));
  }
  // last non-zero position (+1) among this thread's items
  u32 lastnz = 0;
#pragma unroll
  for (int j = 0; j < R2_ITEMS; j++)
    if (p0 + j < n && v[j] != 0) lastnz = p0 + j + 1;
  // exclusive max scan over threads
  u32 inc = warp_incl_max(lastnz);
  u32 exw = __shfl_up_sync(FULL_MASK, inc, 1);
  if (lane_id() == 0) exw = 0;
  if (lane_id() == 31) ws[tid >> 5] = inc;
  __syncthreads();
  if (tid < 32) {
    u32 x = (tid < R2_THREADS / 32) ? ws[tid] : 0u;
    u32 xi = warp_incl_max(x);
    u32 xe = __shfl_up_sync(FULL_MASK, xi, 1);
    if (tid == 0) xe = 0;
    if (tid < R2_THREADS / 32) ws[tid] = xe;
  }
  __syncthreads();
  const u32 ex_run = max(exw, ws[tid >> 5]);
  __syncthreads();
  u32 rs = max(start - pl.y, ex_run);  // position after the last non-zero before p (pl.y zeros were carried into the tile)
  // symbols emitted per item: a non-zero rank flushes the run in front of it, then writes itself
  u32 rlen[R2_ITEMS];
  u32 sum = 0;
#pragma unroll
  for (int j = 0; j < R2_ITEMS; j++) {
    const u32 p = p0 + j;
    rlen[j] = 0;
    if (p < n && v[j] != 0) {
      const u32 L = p - rs;
      rlen[j] = L;
      sum += 1 + (L ? 31 - __clz(L + 1) : 0);
      rs = p + 1;
    }
  }
  u32 tot_off;
  const u32 ex_off = block_excl_add<R2_THREADS, u32>(sum, ws, &tot_off);
  const u32 first = pl.x & 7u;
  u32 o = first + ex_off;  // index into `stage`
#pragma unroll
  for (int j = 0; j < R2_ITEMS; j++) {
    const u32 p = p0 + j;
    if (p < n && v[j] != 0) {
      u32 L = rlen[j];
      while (L) {  // lib/Bzip2.js:783-794 emitLastRun
        if (L & 1) { stage[o++] = 0; atomicAdd(&hist[0], 1u); L -= 1; }
        else { stage[o++] = 1; atomicAdd(&hist[1], 1u); L -= 2; }
        L >>= 1;
      }
      const u32 sy = (u32)v[j] + 1;
      stage[o++] = (u16)sy;
      atomicAdd(&hist[sy], 1u);
    }
  }
  __syncthreads();
  {
    const u32 last = first + tot_off;
    u16* ag = a + (pl.x - first);  // 16-byte aligned: the slot base is, and (pl.x - first) is a multiple of 8 symbols
    for (u32 c8 = tid * 8u; c8 < last; c8 += R2_THREADS * 8u) {
      if (c8 >= first && c8 + 8u <= last) {
        *reinterpret_cast<uint4*>(ag + c8) = *reinterpret_cast<const uint4*>(stage + c8);
      } else {
        const u32 e = min(c8 + 8u, last);
        for (u32 x = max(c8, first); x < e; x++) ag[x] = stage[x];
      }
    }
  }
  for (u32 i = tid; i < HUFF_MAXSYM; i += R2_THREADS)
    if (hist[i]) atomicAdd(&freq[(size_t)seg * HUFF_MAXSYM + i], hist[i]);
}

void mtf_rle2_batch(Ctx& c, const u8* d_T, const u8* d_U, const u32* d_n, const u32* h_n, u32 nblk, u16* d_sym, u32* d_m, u32* d_freq,
                    u32* d_used, const u32* d_bytehist) {
  (void)d_T;
  u32 n_max = 0;
  for (u32 b = 0; b < nblk; b++) n_max = h_n[b] > n_max ? h_n[b] : n_max;
  if (n_max == 0) return;
  const u32 utiles = (n_max + 256 * 64 - 1) / (256 * 64);
  CUDA_CHECK(cudaMemsetAsync(d_used, 0, (size_t)nblk * 8 * 4, c.stream));
  CUDA_CHECK(cudaMemsetAsync(d_freq, 0, (size_t)nblk * HUFF_MAXSYM * 4, c.stream));
  CUDA_CHECK(cudaMemsetAsync(d_m, 0, (size_t)nblk * 4, c.stream));
  if (d_bytehist) k_used_from_hist<<<nblk, 256, 0, c.stream>>>(d_bytehist, d_used);  // the BWT column is a permutation of the block
  else k_used<<<utiles * nblk, 256, 0, c.stream>>>(d_U, d_n, utiles, d_used);
  KLAUNCH(c); KCHECK();
  const u32 cps = (n_max + MTF_CHUNK - 1) / MTF_CHUNK;
  DBuf<u32> lastpos(c, (size_t)nblk * cps * 256);
  DBuf<u8> R(c, (size_t)nblk << SEG_SHIFT);
  DBuf<uint4> rsum(c, (size_t)nblk * cps);
  DBuf<uint2> plan(c, (size_t)nblk * cps);
  k_mtf_lastpos<<<cps * nblk, 256, 0, c.stream>>>(d_U, d_n, cps, lastpos);
  KLAUNCH(c); KCHECK();
  k_mtf_prefix<<<nblk, 256, 0, c.stream>>>(d_n, cps, lastpos);
  KLAUNCH(c); KCHECK();
  {
    const u32 chunks = cps * nblk;
    k_mtf_ranks<<<(chunks + MR_WARPS - 1) / MR_WARPS, MR_WARPS * 32, 0, c.stream>>>(d_U, d_n, cps, lastpos, d_used, R, nblk, rsum);
    KLAUNCH(c); KCHECK();
  }
  k_rle2_scan<<<nblk, 32, 0, c.stream>>>(rsum, d_n, cps, d_used, plan, d_sym, d_m, d_freq);
  KLAUNCH(c); KCHECK();
  static_assert(R2_TILE == MTF_CHUNK, "one zero-run tile per MTF chunk");
  k_rle2<<<cps * nblk, R2_THREADS, 0, c.stream>>>(R, d_n, cps, plan, d_sym, d_freq);
  KLAUNCH(c); KCHECK();
}
