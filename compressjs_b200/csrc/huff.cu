// huff.cu -- per-block Huffman table search and code-length assignment on the GPU.
//
// Reference: lib/Bzip2.js:826-843 (table count rule, two seed tables), :671-684
// (assignSelectors), :685-733 (optimizeHuffmanGroups: split the most used table at the median
// group cost -- STABLE sort), :551-579 (StaticHuffman ctor) and
// lib/HuffmanAllocator.js (huffalloc.cuh).
//
// One CTA owns one bzip2 block for the whole search (the refinement rounds are sequential
// inside a block); hundreds of blocks are in flight.  Inside the CTA:
//   * a 50-symbol group is costed under all tables at once: the <=6 code lengths of a symbol
//     are packed 10 bits apart into one 64-bit word, so one shared-memory add per symbol
//     accumulates every table's cost (staged through shared memory, conflict-free stride 25)
//   * the stable median split needs no sort: histogram of costs -> threshold cost, then an
//     ordered prefix count among the groups that sit exactly on the threshold
//   * tables are rebuilt with a rank-by-counting sort of (freq<<9|sym) and one thread per table
//     running the exact in-place allocator
#include "enc.h"
#include "huffalloc.cuh"

#define HF_THREADS 256
#define HF_TILE_GROUPS 256
#define HF_TILE_WORDS (HF_TILE_GROUPS * 25)

struct HuffSmem {
  u32 tile[HF_TILE_WORDS];             // staged symbols: 256 groups x 25 words
  u16 cost[SEL_STRIDE];                // best cost per group
  u8 sel[SEL_STRIDE];                  // selector per group
  unsigned long long pk[HUFF_MAXSYM + 2];  // packed code lengths (10 bits per table)
  u32 freq[HUFF_MAXGROUPS][HUFF_MAXSYM + 2];
  int work[HUFF_MAXGROUPS][HUFF_MAXSYM + 2];   // allocator arrays (sorted frequencies -> lengths)
  u32 skey[HUFF_MAXGROUPS][HUFF_MAXSYM + 2];   // sort keys (freq << 9 | sym)
  u16 order[HUFF_MAXGROUPS][HUFF_MAXSYM + 2];  // sorted position -> symbol
  u8 len[HUFF_MAXGROUPS][HUFF_MAXSYM + 6];
  u32 chist[1024];
  u32 ws[HF_THREADS / 32 + 1];
  u32 gcount[HUFF_MAXGROUPS];
  u32 misc[8];
};

// build the code lengths of `ntab` tables from s.freq (all threads)
__device__ void build_tables(HuffSmem& s, u32 ntab, u32 A) {
  const u32 tid = threadIdx.x;
  for (u32 i = tid; i < ntab * A; i += HF_THREADS) {
    const u32 t = i / A, sym = i % A;
    s.skey[t][sym] = (s.freq[t][sym] << 9) | sym;  // lib/Bzip2.js:566-568
  }
  __syncthreads();
  // rank by counting (keys are unique)
  for (u32 i = tid; i < ntab * A; i += HF_THREADS) {
    const u32 t = i / A, sym = i % A;
    const u32 k = s.skey[t][sym];
    u32 r = 0;
    for (u32 j = 0; j < A; j++) r += (s.skey[t][j] < k) ? 1u : 0u;
    s.order[t][r] = (u16)sym;
    s.work[t][r] = (int)(k >> 9);
  }
  __syncthreads();
  if (tid < ntab) ha_allocate(s.work[tid], (int)A, 20);  // MAX_HUFCODE_BITS lib/Bzip2.js:40
  __syncthreads();
  for (u32 i = tid; i < ntab * A; i += HF_THREADS) {
    const u32 t = i / A, r = i % A;
    s.len[t][s.order[t][r]] = (u8)s.work[t][r];
  }
  __syncthreads();
  for (u32 sym = tid; sym < A; sym += HF_THREADS) {
    unsigned long long p = 0;
    for (u32 t = 0; t < ntab; t++) p |= (unsigned long long)s.len[t][sym] << (10 * t);
    s.pk[sym] = p;
  }
  __syncthreads();
}

// lib/Bzip2.js:671-684: every group goes to the table that codes it in the fewest bits
// (ties -> lowest table index).  Fills s.sel / s.cost.
// Tiles of 256 groups (6400 words) are staged through shared memory; the next tile's global loads are
// issued into registers before the current tile is consumed, so the L2/HBM latency overlaps the math.
struct TileRegs { u32 r[HF_TILE_WORDS / HF_THREADS]; };
__device__ __forceinline__ void tile_fetch(TileRegs& t, const u32* __restrict__ symw, u32 w0, u32 nwords) {
#pragma unroll
  for (int k = 0; k < HF_TILE_WORDS / HF_THREADS; k++) {
    const u32 i = w0 + k * HF_THREADS + threadIdx.x;
    t.r[k] = i < nwords ? symw[i] : 0u;
  }
}
__device__ __forceinline__ void tile_store(const TileRegs& t, u32* tile) {
#pragma unroll
  for (int k = 0; k < HF_TILE_WORDS / HF_THREADS; k++) tile[k * HF_THREADS + threadIdx.x] = t.r[k];
}

__device__ void assign_selectors(HuffSmem& s, const u32* __restrict__ symw, u32 m, u32 nsel, u32 ntab) {
  const u32 tid = threadIdx.x;
  const u32 nwords = (m + 1) >> 1;
  TileRegs tr;
  tile_fetch(tr, symw, 0, nwords);
  for (u32 g0 = 0; g0 < nsel; g0 += HF_TILE_GROUPS) {
    tile_store(tr, s.tile);
    __syncthreads();
    if (g0 + HF_TILE_GROUPS < nsel) tile_fetch(tr, symw, (g0 + HF_TILE_GROUPS) * 25, nwords);
    const u32 g = g0 + tid;
    if (g < nsel) {
      const u32 cnt = min(50u, m - 50u * g);
      unsigned long long acc = 0;
      if (cnt == 50) {
#pragma unroll 5
        for (u32 k = 0; k < 25; k++) {
          const u32 w = s.tile[tid * 25 + k];
          acc += s.pk[w & 0xffffu] + s.pk[w >> 16];
        }
      } else {
        for (u32 k = 0; k < 25; k++) {
          const u32 w = s.tile[tid * 25 + k];
          if (2 * k < cnt) acc += s.pk[w & 0xffffu];
          if (2 * k + 1 < cnt) acc += s.pk[w >> 16];
        }
      }
      u32 best = 0, bc = (u32)(acc & 1023u);
      for (u32 t = 1; t < ntab; t++) {
        const u32 cst = (u32)((acc >> (10 * t)) & 1023u);
        if (cst < bc) { best = t; bc = cst; }
      }
      s.sel[g] = (u8)best;
      s.cost[g] = (u16)bc;
    }
    __syncthreads();
  }
}

__device__ void recount(HuffSmem& s, const u32* __restrict__ symw, u32 m, u32 nsel, u32 ntab, u32 A) {
  const u32 tid = threadIdx.x;
  const u32 nwords = (m + 1) >> 1;
  for (u32 i = tid; i < ntab * (HUFF_MAXSYM + 2); i += HF_THREADS) (&s.freq[0][0])[i] = 0;
  TileRegs tr;
  tile_fetch(tr, symw, 0, nwords);
  __syncthreads();
  for (u32 g0 = 0; g0 < nsel; g0 += HF_TILE_GROUPS) {
    tile_store(tr, s.tile);
    __syncthreads();
    if (g0 + HF_TILE_GROUPS < nsel) tile_fetch(tr, symw, (g0 + HF_TILE_GROUPS) * 25, nwords);
    const u32 g = g0 + tid;
    if (g < nsel) {
      const u32 cnt = min(50u, m - 50u * g);
      u32* f = s.freq[s.sel[g]];
      for (u32 k = 0; k < 25; k++) {
        const u32 w = s.tile[tid * 25 + k];
        if (2 * k < cnt) atomicAdd(&f[w & 0xffffu], 1u);
        if (2 * k + 1 < cnt) atomicAdd(&f[w >> 16], 1u);
      }
    }
    __syncthreads();
  }
  (void)A;
}

// ---- move-to-front over <= 6 table ids, list packed as six nibbles -------------------------------
__device__ __forceinline__ u32 mtf6_find(u32 list, u32 v) {
  u32 j = 0;
#pragma unroll
  for (u32 k = 1; k < HUFF_MAXGROUPS; k++) j = (((list >> (4 * k)) & 15u) == v) ? k : j;
  return j;
}
__device__ __forceinline__ u32 mtf6_front(u32 list, u32 j) {  // move the entry at position j to the front
  const u32 v = (list >> (4 * j)) & 15u;
  const u32 low = list & ((1u << (4 * j)) - 1u);
  const u32 high = list & ~((1u << (4 * (j + 1))) - 1u);
  return high | (low << 4) | v;
}
// The selector MTF is value based ("find table id v"), so a run of selectors is summarised by its
// RECENCY list: the distinct ids it used, most recent first (count in bits 28..31, unused nibbles 0).
// list after the run from any start list Y = R ++ (Y minus R).
__device__ __forceinline__ u32 rec_apply(u32 R, u32 v) {
  u32 cnt = R >> 28, lst = R & 0x00ffffffu, j = cnt;
#pragma unroll
  for (u32 k = 0; k < HUFF_MAXGROUPS; k++) j = (k < cnt && ((lst >> (4 * k)) & 15u) == v) ? k : j;
  if (j == cnt) {
    lst = ((lst << 4) | v) & 0x00ffffffu;
    cnt++;
  } else {
    const u32 low = lst & ((1u << (4 * j)) - 1u);
    const u32 high = lst & ~((1u << (4 * (j + 1))) - 1u);
    lst = high | (low << 4) | v;
  }
  return (cnt << 28) | lst;
}
__device__ __forceinline__ bool rec_has(u32 R, u32 v) {
  const u32 cnt = R >> 28;
  bool f = false;
#pragma unroll
  for (u32 k = 0; k < HUFF_MAXGROUPS; k++) f = f || (k < cnt && ((R >> (4 * k)) & 15u) == v);
  return f;
}
__device__ __forceinline__ u32 rec_op(u32 A, u32 B) {  // A = earlier run, B = later run
  u32 cnt = B >> 28, lst = B & 0x00ffffffu;
  const u32 ca = A >> 28;
#pragma unroll
  for (u32 k = 0; k < HUFF_MAXGROUPS; k++) {
    const u32 v = (A >> (4 * k)) & 15u;
    if (k < ca && !rec_has(B, v)) { lst |= v << (4 * cnt); cnt++; }
  }
  return (cnt << 28) | lst;
}
__device__ __forceinline__ u32 rec_full(u32 R) {  // R ++ (identity minus R): the full 6-entry list
  u32 cnt = R >> 28, lst = R & 0x00ffffffu;
#pragma unroll
  for (u32 v = 0; v < HUFF_MAXGROUPS; v++)
    if (!rec_has(R, v)) { lst |= v << (4 * cnt); cnt++; }
  return lst;
}

__global__ void __launch_bounds__(HF_THREADS)
k_huffman(const u16* __restrict__ sym, const u32* __restrict__ m_arr, const u32* __restrict__ freq0, const u32* __restrict__ used,
          u8* __restrict__ sel_out, u8* __restrict__ selmtf_out, HuffBlk* __restrict__ hb_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  HuffSmem& s = *reinterpret_cast<HuffSmem*>(smem_raw);
  const u32 tid = threadIdx.x;
  const u32 blk = blockIdx.x;
  const u32 m = m_arr[blk];
  HuffBlk* hb = hb_out + blk;
  if (m == 0) {
    if (tid == 0) { hb->ngroups = 0; hb->nsel = 0; hb->alpha = 0; hb->m = 0; hb->body_bits = 0; }
    return;
  }
  u32 alpha = 0;
  for (int k = 0; k < 8; k++) alpha += __popc(used[blk * 8 + k]);
  const u32 A = alpha + 2;                    // RUNA, RUNB, alpha-1 MTF positions, EOB
  const u32 nsel = (m + HUFF_GROUP - 1) / HUFF_GROUP;
  const u32* symw = reinterpret_cast<const u32*>(sym + ((size_t)blk << SEG_SHIFT));
  u32 target;                                 // lib/Bzip2.js:826-830
  if (m >= 2400) target = 6; else if (m >= 1200) target = 5; else if (m >= 600) target = 4; else if (m >= 200) target = 3; else target = 2;
  // seed tables: global frequencies, flat frequencies (lib/Bzip2.js:835-837)
  for (u32 i = tid; i < A; i += HF_THREADS) { s.freq[0][i] = freq0[(size_t)blk * HUFF_MAXSYM + i]; s.freq[1][i] = 1; }
  __syncthreads();
  u32 ng = 2;
  build_tables(s, ng, A);
  while (ng < target) {
    assign_selectors(s, symw, m, nsel, ng);
    // which table is used most? (first maximum, lib/Bzip2.js:699)
    if (tid < HUFF_MAXGROUPS) s.gcount[tid] = 0;
    for (u32 i = tid; i < 1024; i += HF_THREADS) s.chist[i] = 0;
    __syncthreads();
    {
      u32 l0 = 0, l1 = 0, l2 = 0, l3 = 0, l4 = 0, l5 = 0;
      for (u32 g = tid; g < nsel; g += HF_THREADS) {
        const u32 v = s.sel[g];
        l0 += v == 0; l1 += v == 1; l2 += v == 2; l3 += v == 3; l4 += v == 4; l5 += v == 5;
      }
      if (l0) atomicAdd(&s.gcount[0], l0);
      if (l1) atomicAdd(&s.gcount[1], l1);
      if (l2) atomicAdd(&s.gcount[2], l2);
      if (l3) atomicAdd(&s.gcount[3], l3);
      if (l4) atomicAdd(&s.gcount[4], l4);
      if (l5) atomicAdd(&s.gcount[5], l5);
    }
    __syncthreads();
    u32 which = 0;
    for (u32 t = 1; t < ng; t++) if (s.gcount[t] > s.gcount[which]) which = t;
    const u32 cntw = s.gcount[which];
    // histogram of the costs of the groups coded by `which`
    for (u32 g = tid; g < nsel; g += HF_THREADS) if (s.sel[g] == which) atomicAdd(&s.chist[s.cost[g]], 1u);
    __syncthreads();
    // stable sort by cost, upper half [cntw>>1, cntw) moves to the new table (lib/Bzip2.js:710-714):
    // threshold cost cstar: below = #(cost < cstar) <= half < #(cost <= cstar)
    if (tid == 0) {
      const u32 half = cntw >> 1;
      u32 cum = 0, cstar = 0;
      for (u32 cv = 0; cv < 1024; cv++) {
        if (cum + s.chist[cv] > half) { cstar = cv; break; }
        cum += s.chist[cv];
      }
      s.misc[0] = cstar;
      s.misc[1] = half - cum;  // how many of the groups with cost == cstar stay (the first ones in index order)
    }
    __syncthreads();
    const u32 cstar = s.misc[0], keep_eq = s.misc[1];
    {
      // ordered prefix count of (sel == which && cost == cstar) over groups in index order
      const u32 per = (nsel + HF_THREADS - 1) / HF_THREADS;
      const u32 ga = tid * per, gb = min(nsel, ga + per);
      u32 eq = 0;
      for (u32 g = ga; g < gb; g++) eq += (s.sel[g] == which && s.cost[g] == cstar) ? 1u : 0u;
      u32 tot;
      u32 ex = block_excl_add<HF_THREADS, u32>(eq, s.ws, &tot);
      for (u32 g = ga; g < gb; g++) {
        if (s.sel[g] != which) continue;
        const u32 cg = s.cost[g];
        bool move;
        if (cg > cstar) move = true;
        else if (cg < cstar) move = false;
        else { move = ex >= keep_eq; ex++; }
        if (move) s.sel[g] = (u8)ng;
      }
    }
    __syncthreads();
    ng++;
    recount(s, symw, m, nsel, ng, A);
    build_tables(s, ng, A);
  }
  assign_selectors(s, symw, m, nsel, ng);  // lib/Bzip2.js:843
  // ---- results + bit accounting ----
  // sum of the code bits
  unsigned long long bits = 0;
  for (u32 g = tid; g < nsel; g += HF_THREADS) bits += s.cost[g];
  {
    // block reduce (64-bit)
    __shared__ unsigned long long red[HF_THREADS / 32];
    for (int o = 16; o > 0; o >>= 1) bits += __shfl_xor_sync(FULL_MASK, bits, o);
    if ((tid & 31) == 0) red[tid >> 5] = bits;
    __syncthreads();
    bits = 0;
    for (int i = 0; i < HF_THREADS / 32; i++) bits += red[i];
    __syncthreads();
  }
  u8* so = sel_out + (size_t)blk * SEL_STRIDE;
  for (u32 g = tid; g < nsel; g += HF_THREADS) so[g] = s.sel[g];
  for (u32 i = tid; i < ng * A; i += HF_THREADS) hb->len[i / A][i % A] = s.len[i / A][i % A];
  // selectors: MTF over the table ids, unary (lib/Bzip2.js:850-862).  Parallel over threads: every
  // thread composes the permutation of its run of selectors, an exclusive scan of the compositions
  // gives its start list, then it replays its run.
  u32 selbits = 0;
  {
    const u32 per = (nsel + HF_THREADS - 1) / HF_THREADS;
    const u32 ga = min(nsel, tid * per), gb = min(nsel, ga + per);
    u32 P = 0;  // empty recency list
    for (u32 g = ga; g < gb; g++) P = rec_apply(P, s.sel[g]);
    u32* sa = s.chist;          // reuse: 2 x 256 words
    u32* sb = s.chist + HF_THREADS;
    sa[tid] = P;
    __syncthreads();
    u32 *src = sa, *dst = sb;
    for (u32 o = 1; o < HF_THREADS; o <<= 1) {
      u32 x = src[tid];
      if (tid >= o) x = rec_op(src[tid - o], x);
      dst[tid] = x;
      __syncthreads();
      u32* tmp = src; src = dst; dst = tmp;
    }
    u32 L = rec_full(tid ? src[tid - 1] : 0u);
    u8* sm = selmtf_out + (size_t)blk * SEL_STRIDE;
    for (u32 g = ga; g < gb; g++) {
      const u32 j = mtf6_find(L, s.sel[g]);
      L = mtf6_front(L, j);
      sm[g] = (u8)j;
      selbits += j + 1;
    }
    __syncthreads();
  }
  {
    unsigned long long b2 = selbits;
    __shared__ unsigned long long red2[HF_THREADS / 32];
    for (int o = 16; o > 0; o >>= 1) b2 += __shfl_xor_sync(FULL_MASK, b2, o);
    if ((tid & 31) == 0) red2[tid >> 5] = b2;
    __syncthreads();
    b2 = 0;
    for (int i = 0; i < HF_THREADS / 32; i++) b2 += red2[i];
    bits += b2;
  }
  if (tid == 0) {
    // header: 48 magic + 32 crc + 1 + 24 pidx + 16 + 16 per used range + 3 + 15 (lib/Bzip2.js:740-758, 847-849)
    unsigned long long hbits = 48 + 32 + 1 + 24 + 16 + 3 + 15;
    for (u32 r = 0; r < 16; r++) {
      const u32 w = used[blk * 8 + (r >> 1)];
      if ((w >> ((r & 1) * 16)) & 0xffffu) hbits += 16;
    }
    // tables: 5 bits + per symbol 2*|delta| + 1 (lib/Bzip2.js:610-629)
    for (u32 t = 0; t < ng; t++) {
      hbits += 5;
      u32 cur = s.len[t][0];
      for (u32 i = 0; i < A; i++) {
        const u32 l = s.len[t][i];
        hbits += 2 * (l > cur ? l - cur : cur - l) + 1;
        cur = l;
      }
    }
    hb->ngroups = ng; hb->nsel = nsel; hb->alpha = alpha; hb->m = m;
    hb->body_bits = hbits + bits;
  }
}

void huffman_batch(Ctx& c, const u16* d_sym, const u32* d_m, const u32* d_freq, const u32* d_used, u32 nblk, u8* d_sel, u8* d_selmtf,
                   HuffBlk* d_hb) {
  static bool attr = false;
  if (!attr) {
    CUDA_CHECK(cudaFuncSetAttribute(k_huffman, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(HuffSmem)));
    attr = true;
  }
  k_huffman<<<nblk, HF_THREADS, sizeof(HuffSmem), c.stream>>>(d_sym, d_m, d_freq, d_used, d_sel, d_selmtf, d_hb);
  KLAUNCH(c); KCHECK();
}
