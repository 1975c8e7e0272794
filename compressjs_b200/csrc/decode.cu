// decode.cu -- many-block parallel bzip2 decode on the GPU.
//
// Reference: lib/Bzip2.js:90-548 (Bunzip: _start_bunzip, _get_next_block, _read_bunzip, decode,
// decodeBlock, table).  The reference decodes strictly serially, one bit at a time.  Here:
//
//   k_scan_magic   : every bit offset is tested for the 48-bit block / end-of-stream magics
//                    (blocks start at arbitrary bit positions and a .bz2 has no index)
//   k_hdec         : one CTA (4 warps) per candidate block: header parse (symbol map, selectors, code
//                    length tables -- lib/Bzip2.js:137-275), canonical decode tables exactly as the
//                    reference builds them (limit/base/permute) plus a 9-bit LUT derived from them,
//                    then the Huffman symbol stream (lib/Bzip2.js:288-307): per 50-symbol group the
//                    (symbol, length) that would start at EVERY bit offset of a window, jump tables
//                    1/2/4 symbols ahead, and one thread following 13 jumps
//   k_unmtf_a/scan/map: RUNA/RUNB expansion and inverse move-to-front (lib/Bzip2.js:312-361), chunk
//                    parallel: one serial list pass per chunk that records start-list POSITIONS and
//                    the chunk's composite permutation, a per-block scan over the chunks, then a
//                    lane-parallel pass that maps positions to bytes and expands the runs
//   inverse BWT    : T-vector by one onesweep radix pass on the L column (lib/Bzip2.js:370-381),
//                    then the n-step pointer chase (lib/Bzip2.js:418-423) is broken into ~7000
//                    independent walks per block between sampled rows; every walk records the bytes
//                    it passes, a serial pass over the 7000 walk summaries orders them, and the block
//                    is assembled by copies (only what lies behind a walk's 512-byte record is
//                    walked again)
//   bwt_inverse_sentinel: BWT.unbwtransform (lib/BWT.js:352-363) on the same walk kernels
//   k_unrle_*      : RLE1 decode (lib/Bzip2.js:424-436): count bytes are identified from local
//                    synchronisation points (8 bytes per thread, decided in registers), output
//                    offsets from tile sums + one warp scan per block, tiles expanded in shared
//                    memory, CRC32 per block
//   host           : walks the block chain (a block must start exactly where the previous one
//                    ended), folds/validates CRCs and raises the reference's errors in stream order.
#include <algorithm>
#include <vector>
#include "enc.h"
#include "radix.cuh"
#include "radix_host.cuh"

#define WHOLEPI 0x314159265359ull
#define SQRTPI 0x177245385090ull
#define SEL_CAP 32768
#define DEC_OK 0
#define DEC_NOT_BZIP (-2)
#define DEC_DATA_ERROR (-5)
#define DEC_OBSOLETE (-7)

struct Cand {
  u64 pos;     // bit position of the magic
  u32 type;    // 1 = block, 2 = end of stream
  u32 next32;  // the 32 bits that follow the magic (block CRC / stream CRC)
};

struct CandRes {
  int status;      // 0 or a (negative) reference error code
  u32 detail;      // 1 = "initial position out of bounds"
  u32 m;           // decoded symbols incl. EOB
  u32 orig;        // origPointer
  u32 sym_total;   // distinct bytes
  u32 n;           // block length after un-MTF (filled later)
  u32 rawlen;      // bytes after RLE1 decode (filled later)
  u32 pad;
  u64 endbit;      // bit position just behind the EOB code
  u8 sym_to_byte[256];
};

// ---- magic scan ---------------------------------------------------------------------------
// One thread per aligned 32-bit word of the stream = 32 bit offsets: the words w, w+1, w+2 (big endian) hold the 80 bits
// that a candidate starting in word w can span (48-bit magic + the 32 bits behind it need w+3 as well).  The first 32
// bits of the magic are tested with one funnel shift and one compare per offset.
__global__ void k_scan_magic(const u8* __restrict__ in, u64 n, Cand* __restrict__ cands, u32* count, u32 cap) {
  const u64 wi = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  const u64 nwords = (n + 3) / 4;      // the private copy of the input is zero padded (dec_open): reads up to n + 31 are safe
  if (wi >= nwords) return;
  const u32* words = reinterpret_cast<const u32*>(in);
  const u32 w0 = __byte_perm(words[wi], 0, 0x0123), w1 = __byte_perm(words[wi + 1], 0, 0x0123), w2 = __byte_perm(words[wi + 2], 0, 0x0123),
            w3 = __byte_perm(words[wi + 3], 0, 0x0123);
  const u32 M1 = (u32)(WHOLEPI >> 16), M2 = (u32)(SQRTPI >> 16);
#pragma unroll 8
  for (u32 b = 0; b < 32; b++) {
    const u32 h = __funnelshift_l(w1, w0, b);            // stream bits [b, b + 32) of this word pair
    if (h == M1 || h == M2) {
      const u32 mid = __funnelshift_l(w2, w1, b);        // bits [b + 32, b + 64)
      const u64 v = ((u64)h << 16) | (mid >> 16);
      const u64 pos = wi * 32 + b;
      if ((v == WHOLEPI || v == SQRTPI) && pos < n * 8) {
        const u32 idx = atomicAdd(count, 1u);
        if (idx < cap) {
          Cand c;
          c.pos = pos;
          c.type = v == WHOLEPI ? 1u : 2u;
          const u32 lo = __funnelshift_l(w3, w2, b);     // bits [b + 64, b + 96)
          c.next32 = (mid << 16) | (lo >> 16);           // the 32 bits behind the 48-bit magic
          cands[idx] = c;
        }
      }
    }
  }
}

// ---- bit reader (one thread) ----------------------------------------------------------------
struct BitReader {
  const u32* words; u64 nwords; u64 widx; u64 buf; u32 avail;
  __device__ __forceinline__ u32 fetch() {
    u32 w = widx < nwords ? words[widx] : 0u;  // bits past EOF read as zeros (lib/BitStream.js:88-89)
    widx++;
    return __byte_perm(w, 0, 0x0123);
  }
  __device__ __forceinline__ void init(const u8* base, u64 nbytes, u64 bitpos) {
    words = reinterpret_cast<const u32*>(base);
    nwords = (nbytes + 3) / 4;  // the input buffer is zero padded to a multiple of 4 (+16)
    widx = bitpos >> 5;
    const u32 skip = (u32)(bitpos & 31);
    buf = (u64)fetch() << 32;
    buf <<= skip;
    avail = 32 - skip;
  }
  __device__ __forceinline__ void ensure() {
    if (avail <= 32) { buf |= (u64)fetch() << (32 - avail); avail += 32; }
  }
  __device__ __forceinline__ u32 peek(u32 k) { return k ? (u32)(buf >> (64 - k)) : 0u; }
  __device__ __forceinline__ void skip(u32 k) { buf <<= k; avail -= k; }
  __device__ __forceinline__ u32 get(u32 k) { ensure(); u32 v = peek(k); skip(k); return v; }
  __device__ __forceinline__ u64 tell() const { return widx * 32 - avail; }
};

// ---- header + Huffman decode: one CTA per candidate ------------------------------------------
// One CTA per block.  The per-group phases are spread over the CTA's warps, and a block's ~18 000 groups are strictly
// serial, so a launch lasts as long as one block takes: with all SM slots taken (9 x 128 threads: ~1300 blocks, a 1 GiB
// file) four warps per block give the best throughput; with fewer blocks the same thread budget goes to fewer, wider CTAs
// (256 or 512 threads: fewer window offsets per thread, a shorter group).
#define HD_THREADS 128
#define HD_LUT_BITS 9   // 6 tables x 512 entries: keeps a warp's state under 19 KB so that 12 blocks fit per SM
#define HD_WIN 512
#define HD_STAGE 64
struct HdecWarp {
  int limit[HUFF_MAXGROUPS][22];
  int base[HUFF_MAXGROUPS][22];
  u16 permute[HUFF_MAXGROUPS][HUFF_MAXSYM + 2];
  u16 lut[HUFF_MAXGROUPS][1 << HD_LUT_BITS];
  u8 len[HUFF_MAXGROUPS][HUFF_MAXSYM + 2];
  int minlen[HUFF_MAXGROUPS], maxlen[HUFF_MAXGROUPS];
  int status; u32 ngroups, nsel, symcount;
  // speculative window decode of the symbol stream
  u32 win[HD_STAGE + 20]; // staged stream words (big-endian), refilled every ~2048 bits
  u16 wsym[HD_WIN + 32]; // symbol that would start at every bit offset of the window; tail stays 0
  u8 wlen[HD_WIN + 32]; // its code length (0 = no valid code there); tail stays 0
  u16 spos[HUFF_GROUP + 6];
  u16 J1[HD_WIN + 32], J2[HD_WIN + 32], J4[HD_WIN + 32];  // chain jump tables: 1, 2 and 4 symbols ahead
  u16 q4[16];
  u32 c_cnt, c_pos, c_flag;
  u64 P0;
};

// code longer than the LUT covers: the reference's limit search (lib/Bzip2.js:296-306); returns sym | len << 9,
// 0 when no code of the table starts with these 20 bits
__device__ __noinline__ u32 hdec_slow(const HdecWarp& s, u32 g, u32 bits20, int minLen, int maxLen) {
  int i = minLen;
  int j = (int)(bits20 >> (20 - i));
  for (;;) {
    if (i > maxLen) return 0;                                            // :299
    if (j <= s.limit[g][i]) break;
    i++;
    if (i > 20) return 0;
    j = (int)(bits20 >> (20 - i));
  }
  const int jj = j - s.base[g][i];
  if (jj >= 0 && jj < HUFF_MAXSYM) return (u32)s.permute[g][jj] | ((u32)i << 9);  // :306
  return 0;
}

// 128 threads: 9 CTAs per SM (56 registers, 19 KB of shared memory each): a 1 GiB file's ~1200 blocks are resident at once
template <int HD_T>
__global__ void __launch_bounds__(HD_T, 1152 / HD_T)
k_hdec(const u8* __restrict__ in, u64 nbytes, const Cand* __restrict__ cands, u32 first, u32 count, u32 dbuf_size, u8* __restrict__ sel_buf,
       u16* __restrict__ sym_out, CandRes* __restrict__ res) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  HdecWarp& s = *reinterpret_cast<HdecWarp*>(smem_raw);
  const u32 lane = threadIdx.x;  // 0..HD_T-1: one CTA per candidate block
  const u32 ci = blockIdx.x;
  if (ci >= count) return;
  const Cand cd = cands[first + ci];
  CandRes* r = res + ci;
  u8* sel = sel_buf + (size_t)ci * SEL_CAP;
  u16* so = sym_out + ((size_t)ci << SEG_SHIFT);
  BitReader br;
  u32 symTotal = 0;
  if (lane == 0) {
    s.status = 0;
    r->detail = 0; r->m = 0; r->n = 0; r->rawlen = 0; r->endbit = 0;
    br.init(in, nbytes, cd.pos + 48 + 32);
    do {
      if (br.get(1)) { s.status = DEC_OBSOLETE; break; }           // lib/Bzip2.js:143
      const u32 orig = br.get(24);
      r->orig = orig;
      if (orig > dbuf_size) { s.status = DEC_DATA_ERROR; r->detail = 1; break; }  // :146
      const u32 t16 = br.get(16);
      for (int i = 0; i < 16; i++) {
        if (t16 & (1u << (15 - i))) {
          const u32 k = br.get(16);
          for (int j = 0; j < 16; j++)
            if (k & (1u << (15 - j))) r->sym_to_byte[symTotal++] = (u8)(i * 16 + j);
        }
      }
      const u32 gc = br.get(3);
      if (gc < 2 || gc > 6) { s.status = DEC_DATA_ERROR; break; }    // :167
      const u32 ns = br.get(15);
      if (ns == 0) { s.status = DEC_DATA_ERROR; break; }             // :174
      u8 mtf[HUFF_MAXGROUPS + 2];  // the reference's list is a zero-filled 256-entry buffer: slot gc reads as 0
      for (u32 i = 0; i < HUFF_MAXGROUPS + 2; i++) mtf[i] = (u8)(i < gc ? i : 0);
      bool bad = false;
      for (u32 i = 0; i < ns && !bad; i++) {
        u32 j = 0;
        while (br.get(1)) { if (j >= gc) { bad = true; break; } j++; }   // :184-185
        if (bad) break;
        const u8 v = mtf[j];
        for (u32 k = j; k > 0; k--) mtf[k] = mtf[k - 1];
        mtf[0] = v;
        sel[i] = v;
      }
      if (bad) { s.status = DEC_DATA_ERROR; break; }
      const u32 symCount = symTotal + 2;
      for (u32 g = 0; g < gc && !bad; g++) {
        int t = (int)br.get(5);
        for (u32 i = 0; i < symCount; i++) {
          for (;;) {
            if (t < 1 || t > 20) { bad = true; break; }               // :203
            if (!br.get(1)) break;
            if (!br.get(1)) t++; else t--;
          }
          if (bad) break;
          s.len[g][i] = (u8)t;
        }
      }
      if (bad) { s.status = DEC_DATA_ERROR; break; }
      s.ngroups = gc; s.nsel = ns; s.symcount = symCount;
    } while (0);
    r->sym_total = symTotal;
  }
  __syncthreads();
  if (s.status != 0) {
    if (lane == 0) r->status = s.status;
    return;
  }
  const u32 gc = s.ngroups, symCount = s.symcount;
  // ---- limit / base / permute exactly as lib/Bzip2.js:216-274, one lane per table ----
  if (lane < gc) {
    const u32 g = lane;
    int minLen = s.len[g][0], maxLen = s.len[g][0];
    for (u32 i = 1; i < symCount; i++) {
      const int l = s.len[g][i];
      if (l > maxLen) maxLen = l; else if (l < minLen) minLen = l;
    }
    s.minlen[g] = minLen; s.maxlen[g] = maxLen;
    int temp[22];
    for (int i = 0; i < 22; i++) { temp[i] = 0; s.limit[g][i] = 0; s.base[g][i] = 0; }
    for (u32 i = 0; i < HUFF_MAXSYM; i++) s.permute[g][i] = 0;
    int pp = 0;
    for (int i = minLen; i <= maxLen; i++)
      for (u32 t = 0; t < symCount; t++)
        if (s.len[g][t] == i) s.permute[g][pp++] = (u16)t;
    for (u32 i = 0; i < symCount; i++) temp[s.len[g][i]]++;
    pp = 0;
    int t = 0;
    for (int i = minLen; i < maxLen; i++) {
      pp += temp[i];
      s.limit[g][i] = pp - 1;
      pp <<= 1;
      t += temp[i];
      s.base[g][i + 1] = pp - t;
    }
    s.limit[g][maxLen] = pp + temp[maxLen] - 1;
    s.base[g][minLen] = 0;
  }
  __syncthreads();
  // ---- LUT (HD_LUT_BITS bits) derived from the reference's decode loop (lib/Bzip2.js:296-307) ----
  for (u32 e = lane; e < gc << HD_LUT_BITS; e += HD_T) {
    const u32 g = e >> HD_LUT_BITS, p = e & ((1u << HD_LUT_BITS) - 1);
    const int minLen = s.minlen[g], maxLen = s.maxlen[g];
    u16 ent = 0;  // 0 = needs more than HD_LUT_BITS bits (or fails): take the slow path
    int i = minLen;
    if (i <= HD_LUT_BITS) {
      int j = (int)(p >> (HD_LUT_BITS - i));
      for (;;) {
        if (i > maxLen) break;
        if (j <= s.limit[g][i]) {
          const int jj = j - s.base[g][i];
          if (jj >= 0 && jj < HUFF_MAXSYM) ent = (u16)((i << 9) | s.permute[g][jj]);
          break;
        }
        i++;
        if (i > HD_LUT_BITS) break;
        j = (j << 1) | (int)((p >> (HD_LUT_BITS - i)) & 1);
      }
    }
    s.lut[g][p] = ent;
  }
  __syncthreads();
  // ---- symbol stream (lib/Bzip2.js:288-307).  Per 50-symbol group the table is fixed, so all 32
  // lanes decode the (symbol, length) that WOULD start at every bit offset of a 512-bit window, and
  // lane 0 only follows the chain pos += len[pos] through shared memory; the symbols on the chain are
  // then written out by the whole warp. ----
  {
    const u32 ns = s.nsel, eob = s.symcount - 1;  // symTotal + 1
    if (lane == 0) s.P0 = br.tell();
    __syncthreads();
    u64 P = s.P0;
    const u32* words = reinterpret_cast<const u32*>(in);
    const u64 nwords = (nbytes + 3) / 4;
    u32 m = 0, selector = 0;
    int status = 0;
    bool done = false;
    if (lane < 32) { s.wlen[HD_WIN + lane] = 0; s.wsym[HD_WIN + lane] = 0; }
    u64 stage_w0 = ~0ull;  // index of the stream word held in win[0]
    u32 wlim = HD_WIN;     // bit offsets decoded per window: adapts to the size of the previous group
    __syncthreads();
    while (!done) {
      if (selector >= ns) { status = DEC_DATA_ERROR; break; }          // :291
      const u32 g = sel[selector++];
      const u16* lut = s.lut[g];
      const int minLen = s.minlen[g], maxLen = s.maxlen[g];
      u32 remaining = HUFF_GROUP;
      while (remaining && !done) {
        // (re)stage HD_STAGE + 18 words when the 544-bit window would leave the staged range
        const u64 w0 = P >> 5;
        if (stage_w0 == ~0ull || w0 < stage_w0 || w0 + 18 > stage_w0 + HD_STAGE + 18) {
          __syncthreads();
          for (u32 i = lane; i < HD_STAGE + 18; i += HD_T) {
            const u64 wi = w0 + i;
            const u32 wv = wi < nwords ? words[wi] : 0u;
            s.win[i] = __byte_perm(wv, 0, 0x0123);
          }
          stage_w0 = w0;
          __syncthreads();
        }
        const u32 shiftbase = (u32)(P & 31);
        u32 j1r[HD_WIN / HD_T], j2r[HD_WIN / HD_T];  // this thread's jump targets, kept in registers between the passes
        {
          // thread t decodes offsets t, t+128, ...: same bit shift every time
          const u32 sh = (shiftbase + lane) & 31;
          const u32* wp = s.win + (u32)(w0 - stage_w0) + ((shiftbase + lane) >> 5);
#pragma unroll
          for (u32 k = 0; k < HD_WIN / HD_T; k++) {
            const u32 o = lane + HD_T * k;
            j1r[k] = o;
            if (o < wlim) {
              const u32 hiw = wp[(HD_T / 32) * k], low = wp[(HD_T / 32) * k + 1];
              const u32 bits20 = __funnelshift_l(low, hiw, sh) >> 12;
              u32 ent = lut[bits20 >> (20 - HD_LUT_BITS)];
              if (!ent) ent = hdec_slow(s, g, bits20, minLen, maxLen);
              const u32 len = ent >> 9;
              s.wsym[o] = (u16)(ent & 511u);
              s.wlen[o] = (u8)len;
              // J1[o] = o + len[o] (an offset without a code maps to itself, one beyond the window parks there)
              j1r[k] = min(o + len, wlim + 31u);
              s.J1[o] = (u16)j1r[k];
            }
          }
        }
        if (lane < 32) {
          // everything behind the decoded window parks the chain
          const u32 o = wlim + lane;
          s.wlen[o] = 0; s.wsym[o] = 0;
          s.J1[o] = (u16)o; s.J2[o] = (u16)o; s.J4[o] = (u16)o;
        }
        __syncthreads();
#pragma unroll
        for (u32 k = 0; k < HD_WIN / HD_T; k++)
          if (lane + HD_T * k < wlim) j2r[k] = s.J1[j1r[k]];
#pragma unroll
        for (u32 k = 0; k < HD_WIN / HD_T; k++)
          if (lane + HD_T * k < wlim) s.J2[lane + HD_T * k] = (u16)j2r[k];
        __syncthreads();
#pragma unroll
        for (u32 k = 0; k < HD_WIN / HD_T; k++)
          if (lane + HD_T * k < wlim) j1r[k] = s.J2[j2r[k]];
#pragma unroll
        for (u32 k = 0; k < HD_WIN / HD_T; k++)
          if (lane + HD_T * k < wlim) s.J4[lane + HD_T * k] = (u16)j1r[k];
        __syncthreads();
        if (lane == 0) {
          // the only serial part: 13 dependent shared-memory loads cover 52 symbols
          u32 pos = 0;
#pragma unroll
          for (u32 i = 0; i < 13; i++) { s.q4[i] = (u16)pos; pos = s.J4[pos]; }
        }
        __syncthreads();
        if (lane < 13) {
          const u32 p0 = s.q4[lane], p1 = s.J1[p0], p2 = s.J1[p1], p3 = s.J1[p2];
          s.spos[4 * lane] = (u16)p0; s.spos[4 * lane + 1] = (u16)p1; s.spos[4 * lane + 2] = (u16)p2; s.spos[4 * lane + 3] = (u16)p3;
        }
        __syncthreads();
        if (lane < 32) {
          // first warp: where does the group stop inside this window?
          const u32 lim = remaining;  // <= 50
          u32 first_stop = 0xffffffffu, first_eob = 0xffffffffu;
          for (u32 i = lane; i < lim; i += 32) {
            const u32 p = s.spos[i];
            if (s.wlen[p] == 0 && first_stop == 0xffffffffu) first_stop = i;  // no code here, or past the window
            const u32 sy = s.wsym[p];
            if (sy >= eob && sy > 1 && first_eob == 0xffffffffu) first_eob = i;
          }
          first_stop = __reduce_min_sync(FULL_MASK, first_stop);
          first_eob = __reduce_min_sync(FULL_MASK, first_eob);
          u32 cntw = lim, flag = 0;
          if (first_eob < first_stop && first_eob < lim) { cntw = first_eob + 1; flag = 1; }
          else if (first_stop < lim) { cntw = first_stop; flag = (s.spos[first_stop] < wlim) ? 2u : 0u; }  // :299/:306 vs. window exhausted
          if (lane == 0) { s.c_cnt = cntw; s.c_pos = s.spos[cntw]; s.c_flag = flag; }
        }
        __syncthreads();
        const u32 cnt = s.c_cnt, flag = s.c_flag;
        if (m + cnt >= SEG_SIZE) { status = DEC_DATA_ERROR; done = true; break; }
        if (lane < cnt) so[m + lane] = s.wsym[s.spos[lane]];  // cnt <= 50
        m += cnt;
        remaining -= cnt;
        P += s.c_pos;
        {
          // next window: enough for a whole group at the current bits/symbol plus slack
          const u32 est = cnt ? (s.c_pos * HUFF_GROUP) / cnt + 64u : HD_WIN;
          wlim = min((u32)HD_WIN, max(96u, (est + 31u) & ~31u));
        }
        if (flag == 2) { status = DEC_DATA_ERROR; done = true; }
        else if (flag == 1) done = true;
        __syncthreads();
      }
    }
    if (lane == 0) {
      r->status = status;
      r->m = m;
      r->endbit = P;
    }
  }
}

// ---- inverse MTF + run expansion --------------------------------------------------------------
#define UM_CHUNK 4096
#define UM_WARPS 8
struct ChunkSum {
  u64 leadval;   // sum (d_j+1) << j over the leading run digits
  u32 nlead;     // number of leading run digits
  u32 rest;      // bytes produced by everything behind the leading digits
  u32 trail;     // run digits at the end of the chunk
  u32 allrun;    // chunk consists of run digits only
  u32 nsyms;
  u32 bad;
};

// One THREAD per 4 Ki-symbol chunk: the serial list pass of the reference (lib/Bzip2.js:355-360) on a private list that
// starts as the identity, so what comes out are POSITIONS IN THE CHUNK'S START LIST (k_unmtf_map turns them into bytes
// once k_unmtf_scan has composed the chunks' final lists).  The list lives in shared memory as 64-bit words, word k of
// thread t at [k][t]: no bank conflicts, whatever word a lane touches.  Moving list[idx] to the front shifts idx/8
// words by one byte: a short loop whose length follows the rank (small on anything compressible), against the 48
// warp instructions per symbol of a warp-wide register list.
#define UA_THREADS 128
__global__ void __launch_bounds__(UA_THREADS)
k_unmtf_a(const u16* __restrict__ sym, const CandRes* __restrict__ res, u32 ncand, u32 cps, ChunkSum* __restrict__ sums, u8* __restrict__ perms,
          u8* __restrict__ symb) {
  __shared__ u64 W[32][UA_THREADS];
  const u32 t = threadIdx.x;
  const u32 gchunk = blockIdx.x * UA_THREADS + t;
  const u32 ci = gchunk / cps, ch = gchunk % cps;
  if (ci >= ncand) return;
  const CandRes* r = res + ci;
  if (r->status != 0) return;
  const u32 m = r->m;
  const u32 start = ch * UM_CHUNK;
  if (start >= m) return;
  const u32 count = min((u32)UM_CHUNK, m - start);
  const u16* s = sym + ((size_t)ci << SEG_SHIFT) + start;
  const u32 symTotal = r->sym_total;
#pragma unroll 8
  for (u32 k = 0; k < 32; k++) W[k][t] = 0x0706050403020100ull + k * 0x0808080808080808ull;  // identity list
  u64 leadval = 0; u32 nlead = 0, rest = 0, krun = 0, bad = 0;
  bool seen_lit = false;
  u8* sb = symb + ((size_t)ci << SEG_SHIFT) + start;
  u32 front = 0;
  for (u32 base = 0; base < count; base += 8) {
    // eight symbols per step: one 16-byte load, one 8-byte store (the slot is 1 MiB: reading past `count` stays inside it)
    const uint4 q = *reinterpret_cast<const uint4*>(s + base);
    const u32 qw[4] = {q.x, q.y, q.z, q.w};
    u32 keep_lo = 0, keep_hi = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const u32 sy = (qw[j >> 1] >> (16 * (j & 1))) & 0xffffu;
      if (base + j < count) {
        if (sy <= 1) {
          if (!seen_lit) {
            if (nlead < 21) leadval += (u64)(sy + 1) << nlead; else bad = 1;   // >= 2^21 bytes: over any dbufSize
            nlead++;
          } else {
            if (krun < 21) rest = min(rest + ((sy + 1) << krun), 0x3fffffffu); else bad = 1;  // saturate: no wrap-around
            krun++;
          }
        } else {
          seen_lit = true; krun = 0;
          if (sy <= symTotal) {
            const u32 idx = sy - 1, wq = idx >> 3, bp = idx & 7u;
            const u64 x = W[wq][t];
            const u32 b = (u32)(x >> (8 * bp)) & 0xffu;
            u64 carry = b;
            for (u32 k = 0; k < wq; k++) {
              const u64 y = W[k][t];
              W[k][t] = (y << 8) | carry;
              carry = y >> 56;
            }
            const u64 msk = bp == 7 ? ~0ull : ((1ull << (8 * (bp + 1))) - 1ull);
            W[wq][t] = (((x << 8) | carry) & msk) | (x & ~msk);
            front = b;
            rest++;
          }
        }
      }
      if (j < 4) keep_lo |= front << (8 * j); else keep_hi |= front << (8 * (j - 4));
    }
    *reinterpret_cast<uint2*>(sb + base) = make_uint2(keep_lo, keep_hi);
  }
  {
    ChunkSum cs;
    cs.leadval = leadval; cs.nlead = nlead; cs.rest = rest; cs.trail = seen_lit ? krun : nlead; cs.allrun = seen_lit ? 0u : 1u;
    cs.nsyms = count; cs.bad = bad;
    sums[gchunk] = cs;
  }
  u64* p = reinterpret_cast<u64*>(perms + (size_t)gchunk * 256);
#pragma unroll 8
  for (u32 k = 0; k < 32; k++) p[k] = W[k][t];
}

struct ChunkStart {
  u32 k;    // run digits immediately before the chunk
  u32 off;  // output offset of the chunk
};

__global__ void __launch_bounds__(32)
k_unmtf_scan(CandRes* __restrict__ res, u32 ncand, u32 cps, u32 dbuf_size, const ChunkSum* __restrict__ sums, const u8* __restrict__ perms,
             u8* __restrict__ lists, ChunkStart* __restrict__ starts) {
  __shared__ u8 L[256], P[256];
  const u32 ci = blockIdx.x, lane = threadIdx.x;
  CandRes* r = res + ci;
  if (r->status != 0) return;
  const u32 m = r->m;
  const u32 nch = (m + UM_CHUNK - 1) / UM_CHUNK;
  for (u32 i = lane; i < 256; i += 32) L[i] = (u8)i;
  __syncwarp();
  u64 off = 0; u32 k = 0; int status = 0;
  for (u32 ch = 0; ch < nch; ch++) {
    const size_t gc = (size_t)ci * cps + ch;
    const ChunkSum cs = sums[gc];
    for (u32 i = lane; i < 256; i += 32) { lists[gc * 256 + i] = L[i]; P[i] = perms[gc * 256 + i]; }
    if (lane == 0) { ChunkStart st; st.k = k; st.off = (u32)off; starts[gc] = st; }
    if (cs.bad || (cs.nlead && k + cs.nlead > 32)) { status = DEC_DATA_ERROR; break; }
    off += (cs.leadval << k) + cs.rest;
    if (off > dbuf_size) { status = DEC_DATA_ERROR; break; }        // lib/Bzip2.js:338,354
    k = cs.allrun ? k + cs.nsyms : cs.trail;
    __syncwarp();
    u8 nl[8];
    for (int j = 0; j < 8; j++) nl[j] = L[P[lane * 8 + j]];
    __syncwarp();
    for (int j = 0; j < 8; j++) L[lane * 8 + j] = nl[j];
    __syncwarp();
  }
  if (lane == 0) {
    if (!status && r->orig >= (u32)off) status = DEC_DATA_ERROR;      // lib/Bzip2.js:368
    r->status = status;
    r->n = (u32)off;
  }
}

// Second pass, fully lane parallel: byte = sym_to_byte[start list[symb]], run digit i of a run weighs
// (digit+1) << i (lib/Bzip2.js:312-361); output offsets by a warp scan of the weights.
__global__ void __launch_bounds__(UM_WARPS * 32)
k_unmtf_map(const u16* __restrict__ sym, const u8* __restrict__ symb, const CandRes* __restrict__ res, u32 ncand, u32 cps,
            const u8* __restrict__ lists, const ChunkStart* __restrict__ starts, u8* __restrict__ tt) {
  __shared__ u8 Ts[UM_WARPS][256];
  const u32 w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const u32 gchunk = blockIdx.x * UM_WARPS + w;
  const u32 ci = gchunk / cps, ch = gchunk % cps;
  if (ci >= ncand) return;
  const CandRes* r = res + ci;
  if (r->status != 0) return;
  const u32 m = r->m;
  const u32 start = ch * UM_CHUNK;
  if (start >= m) return;
  const u32 count = min((u32)UM_CHUNK, m - start);
  const u16* s = sym + ((size_t)ci << SEG_SHIFT) + start;
  const u8* sb = symb + ((size_t)ci << SEG_SHIFT) + start;
  const u32 symTotal = r->sym_total;
  u8* T = Ts[w];
  {
    const u8* lp = lists + (size_t)gchunk * 256;
    const u8* s2b = r->sym_to_byte;
    for (u32 i = lane; i < 256; i += 32) T[i] = s2b[lp[i]];
  }
  __syncwarp();
  const ChunkStart st = starts[gchunk];
  u32 k = st.k, o = st.off;
  u8* out = tt + ((size_t)ci << SEG_SHIFT);
  for (u32 base = 0; base < count; base += 32) {
    const bool valid = base + lane < count;
    const u32 sy = valid ? s[base + lane] : 0xffffu;
    const u32 b = T[valid ? sb[base + lane] : 0];
    const bool isrun = sy <= 1;
    const u32 runmask = __ballot_sync(FULL_MASK, isrun);
    const u32 nonrun_below = ~runmask & lanemask_lt();
    const u32 kk = nonrun_below ? lane - (31 - __clz(nonrun_below)) - 1 : lane + k;
    const u32 wgt = isrun ? (sy + 1) << kk : (sy <= symTotal ? 1u : 0u);
    const u32 inc = warp_incl_add(wgt);
    const u32 tot = __shfl_sync(FULL_MASK, inc, 31);
    const u32 oo = o + inc - wgt;
    if (wgt == 1) out[oo] = (u8)b;
    // runs longer than one byte: short ones by their own lane, long ones by the whole warp
    const bool shortrun = wgt > 1 && wgt <= 16;
    if (shortrun) for (u32 x = 0; x < wgt; x++) out[oo + x] = (u8)b;
    u32 big = __ballot_sync(FULL_MASK, wgt > 16);
    while (big) {
      const u32 l = __ffs(big) - 1;
      big &= big - 1;
      const u32 bo = __shfl_sync(FULL_MASK, oo, l), bw = __shfl_sync(FULL_MASK, wgt, l), bb = __shfl_sync(FULL_MASK, b, l);
      for (u32 x = lane; x < bw; x += 32) out[bo + x] = (u8)bb;
    }
    o += tot;
    const u32 nonrun = ~runmask;
    k = nonrun ? (u32)__clz(nonrun) : k + 32;  // run digits at the end of this batch
  }
}

// ---- inverse BWT ------------------------------------------------------------------------------
__global__ void k_ibwt_keys(const u8* __restrict__ tt, const u32* __restrict__ seg_n, u32 nslots, u32* __restrict__ key) {
  const u32 g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= nslots) return;
  if ((g & SEG_MASK) < seg_n[g >> SEG_SHIFT]) key[g] = tt[g];
}
#define IB_SHIFT 7
#define IB_STEP (1u << IB_SHIFT)
#define IB_SEGS (SEG_SIZE / IB_STEP + 1)  // sampled rows per block + the start row
#define IB_VCAP 16384

#define IB_CAP 512u   // bytes a walk records on its way (four sampling steps); 1.8 % of the bytes lie behind that and are re-walked

struct Seg { u32 len, next; };
struct Visit { u32 row, off, len; };

// The bytes a walk passes are kept in a slot of IB_CAP bytes per sampled row (slotA: the 8192 multiples of 2^IB_SHIFT of
// every block, 4 MiB per block; slotB: the start row's slot at the head of the block's 4 MiB), so that once the order of
// the walks is known the block is assembled by copying instead of walking it a second time.
__device__ __forceinline__ u8* ib_slot(u8* slotA, u8* slotB, u32 ci, u32 sid) {
  return sid < IB_SEGS - 1 ? slotA + ((size_t)ci << (SEG_SHIFT + 2)) + (size_t)sid * IB_CAP : slotB + ((size_t)ci << (SEG_SHIFT + 2));
}

// walk from every sampled row (multiples of 2^IB_SHIFT and the start row) to the next sampled row
__global__ void k_ibwt_walk1(const u32* __restrict__ P, const CandRes* __restrict__ res, u32 ncand, Seg* __restrict__ segs, u32* __restrict__ capr,
                             u8* __restrict__ slotA, u8* __restrict__ slotB) {
  const u32 gid = blockIdx.x * blockDim.x + threadIdx.x;
  const u32 ci = gid / IB_SEGS, sid = gid % IB_SEGS;
  if (ci >= ncand) return;
  const CandRes* r = res + ci;
  if (r->status != 0) return;
  const u32 n = r->n;
  const u32* p = P + ((size_t)ci << SEG_SHIFT);
  const u32 r0 = p[r->orig] >> 8;  // first row whose byte is output (lib/Bzip2.js:386-391)
  u32 a;
  if (sid == IB_SEGS - 1) { if ((r0 & (IB_STEP - 1)) == 0) return; a = r0; }
  else { a = sid << IB_SHIFT; if (a >= n) return; }
  uint4* slot = reinterpret_cast<uint4*>(ib_slot(slotA, slotB, ci, sid));
  u32 row = a, steps = 0, rowcap = 0;
  // 32 steps at a time: the bytes of one chunk go out as ONE 32-byte sector (a sector written four bytes at a time is
  // evicted half full while 300 000 walks stream through the L2)
  for (bool more = true; more;) {
    u32 w[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (u32 k = 0; k < 32; k++) {
      const u32 e = p[row];
      w[k >> 2] |= (e & 0xffu) << (8u * (k & 3u));
      row = e >> 8;
      steps++;
      if (steps == IB_CAP) rowcap = row;
      if (!((row & (IB_STEP - 1)) != 0 && row != r0 && steps < n)) { more = false; break; }
    }
    const u32 chunk = (steps - 1) >> 5;
    if (chunk < IB_CAP / 32) {
      slot[2 * chunk] = make_uint4(w[0], w[1], w[2], w[3]);
      slot[2 * chunk + 1] = make_uint4(w[4], w[5], w[6], w[7]);
    }
  }
  Seg sg;
  sg.len = steps;
  sg.next = ((row & (IB_STEP - 1)) == 0) ? (row >> IB_SHIFT) : (IB_SEGS - 1);
  segs[(size_t)ci * IB_SEGS + sid] = sg;
  capr[(size_t)ci * IB_SEGS + sid] = rowcap;
}

// order the segments along the chain that starts at the start row.  One CTA per block: the segment
// table (64 KB) is staged in shared memory so that the serial walk costs a shared-memory load per step.
// tails: the visits that are longer than their slot (indices into the block's visit list).
__global__ void __launch_bounds__(128)
k_ibwt_chain(const u32* __restrict__ P, const CandRes* __restrict__ res, u32 ncand, const Seg* __restrict__ segs,
             Visit* __restrict__ visits, u32* __restrict__ nvisits, u32* __restrict__ tails, u32* __restrict__ ntails) {
  extern __shared__ __align__(8) unsigned char chain_smem[];
  Seg* ss = reinterpret_cast<Seg*>(chain_smem);
  const u32 ci = blockIdx.x;
  if (ci >= ncand) return;
  const CandRes* r = res + ci;
  if (r->status != 0) { if (threadIdx.x == 0) { nvisits[ci] = 0; ntails[ci] = 0; } return; }
  const u32 n = r->n;
  const u32 nseg = min((u32)IB_SEGS, (n >> IB_SHIFT) + 2);
  for (u32 i = threadIdx.x; i < nseg; i += blockDim.x) ss[i] = segs[(size_t)ci * IB_SEGS + i];
  if (threadIdx.x == 0) ss[IB_SEGS - 1] = segs[(size_t)ci * IB_SEGS + IB_SEGS - 1];
  __syncthreads();
  if (threadIdx.x != 0) return;
  const u32* p = P + ((size_t)ci << SEG_SHIFT);
  const u32 r0 = p[r->orig] >> 8;
  u32 cur = ((r0 & (IB_STEP - 1)) == 0) ? (r0 >> IB_SHIFT) : (IB_SEGS - 1);
  u32 off = 0, nv = 0, nt = 0;
  Visit* v = visits + (size_t)ci * IB_VCAP;
  u32* tl = tails + (size_t)ci * IB_VCAP;
  while (off < n) {
    const Seg sg = ss[cur];
    const u32 len = min(sg.len, n - off);
    if (nv >= IB_VCAP) { nv = 0xffffffffu; nt = 0; break; }  // degenerate (periodic) block: fall back to one serial walk
    Visit vv;
    vv.row = (cur == IB_SEGS - 1) ? r0 : (cur << IB_SHIFT);
    vv.off = off; vv.len = len;
    if (len > IB_CAP) tl[nt++] = nv;
    v[nv++] = vv;
    off += len;
    cur = sg.next;
  }
  nvisits[ci] = nv;
  ntails[ci] = nt;
}

// Every visit's recorded bytes are copied to their place in the block: one warp per visit, aligned 32-bit words in the
// middle (realigned from the slot with a funnel shift), single bytes at the two ends (the neighbouring visits own the
// rest of those words).  IB_PLACE_CTAS CTAs share the visits of a block.
#define IB_PLACE_CTAS 4
#define IB_PLACE_THREADS 256
__global__ void __launch_bounds__(IB_PLACE_THREADS)
k_ibwt_place(const u32* __restrict__ P, const CandRes* __restrict__ res, u32 ncand, const Visit* __restrict__ visits,
             const u32* __restrict__ nvisits, const u8* __restrict__ slotA, const u8* __restrict__ slotB, u8* __restrict__ out) {
  const u32 ci = blockIdx.x / IB_PLACE_CTAS;
  if (ci >= ncand) return;
  if (res[ci].status != 0) return;
  const u32 nv = nvisits[ci];
  if (nv == 0xffffffffu) return;
  const u32 lane = threadIdx.x & 31u;
  const u32 wstride = IB_PLACE_CTAS * (IB_PLACE_THREADS / 32);
  const u32* p = P + ((size_t)ci << SEG_SHIFT);
  const u32 r0 = p[res[ci].orig] >> 8;
  const bool r0_own = (r0 & (IB_STEP - 1)) != 0;
  u8* ob = out + ((size_t)ci << SEG_SHIFT);
  for (u32 vi = (blockIdx.x % IB_PLACE_CTAS) * (IB_PLACE_THREADS / 32) + (threadIdx.x >> 5); vi < nv; vi += wstride) {
    const Visit v = visits[(size_t)ci * IB_VCAP + vi];
    const u32 sid = (r0_own && v.row == r0) ? (IB_SEGS - 1) : (v.row >> IB_SHIFT);
    const u8* srcb = ib_slot(const_cast<u8*>(slotA), const_cast<u8*>(slotB), ci, sid);
    const u32* src = reinterpret_cast<const u32*>(srcb);
    u8* o = ob + v.off;
    const u32 len = min(v.len, IB_CAP);
    const u32 head = min(len, (4u - ((u32)(size_t)o & 3u)) & 3u);   // bytes in front of the first aligned word
    const u32 nwords = (len - head) >> 2;
    if (lane < head) o[lane] = srcb[lane];
    u32* ow = reinterpret_cast<u32*>(o + head);
    const u32 sh = 8u * head;                                        // word w of the destination = source bytes head + 4w ..
    for (u32 w = lane; w < nwords; w += 32) {
      const u32 lo = src[w], hi = head ? src[w + 1] : 0u;           // src[w + 1] stays inside the slot: head + 4w + 3 < len <= IB_CAP
      ow[w] = __funnelshift_r(lo, hi, sh);
    }
    const u32 done = head + 4u * nwords;
    if (lane < len - done) o[done + lane] = srcb[done + lane];
  }
}

// what lies behind the recorded part of a long walk is walked again; so is a whole degenerate block
__global__ void k_ibwt_tail(const u32* __restrict__ P, const CandRes* __restrict__ res, u32 ncand, const Visit* __restrict__ visits,
                            const u32* __restrict__ nvisits, const u32* __restrict__ tails, const u32* __restrict__ ntails,
                            const u32* __restrict__ capr, u8* __restrict__ out) {
  const u32 gid = blockIdx.x * blockDim.x + threadIdx.x;
  const u32 ci = gid / IB_VCAP, ti = gid % IB_VCAP;
  if (ci >= ncand) return;
  const CandRes* r = res + ci;
  if (r->status != 0) return;
  const u32 nv = nvisits[ci];
  const u32* p = P + ((size_t)ci << SEG_SHIFT);
  u8* o = out + ((size_t)ci << SEG_SHIFT);
  u32 row, off, len;
  if (nv == 0xffffffffu) {
    if (ti != 0) return;
    row = p[r->orig] >> 8; off = 0; len = r->n;
  } else {
    if (ti >= ntails[ci]) return;
    const Visit v = visits[(size_t)ci * IB_VCAP + tails[(size_t)ci * IB_VCAP + ti]];
    const u32 r0 = p[r->orig] >> 8;
    const u32 sid = (v.row == r0 && (r0 & (IB_STEP - 1)) != 0) ? (IB_SEGS - 1) : (v.row >> IB_SHIFT);
    row = capr[(size_t)ci * IB_SEGS + sid]; off = v.off + IB_CAP; len = v.len - IB_CAP;
  }
  for (u32 t = 0; t < len; t++) {
    const u32 e = p[row];
    o[off + t] = (u8)e;
    row = e >> 8;
  }
}

static void dec_attr_once();

// ---- BWT.unbwtransform (lib/BWT.js:352-363): inverse of the sentinel BWT ------------------------------
// Reference walk: t = 0; for i = n-1..0: U[i] = T[t]; t = LF[t] + C[T[t]]; t += (t < pidx).  LF[t] + C[T[t]] is
// the rank x of position t in the stable order by byte, i.e. the inverse of the sorted-position vector that one
// radix pass produces.  P[t] = next(t) << 8 | T[t] feeds the same sampled-row walks as the bzip2 decoder; the
// spare slot 2^20-1 holds the pseudo start entry (orig) whose successor is row 0.
__global__ void k_unbwt_pack(const u8* __restrict__ L, const u32* __restrict__ sorted_pos, u32 n, u32 pidx, u32* __restrict__ P) {
  const u32 x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x == 0) P[SEG_SIZE - 1] = 0;
  if (x >= n) return;
  const u32 t = sorted_pos[x] & SEG_MASK;
  P[t] = ((x + (x < pidx ? 1u : 0u)) << 8) | L[t];
}
__global__ void k_unbwt_setup(CandRes* r, u32 n) {
  r->status = 0; r->detail = 0; r->m = 0; r->orig = SEG_SIZE - 1; r->sym_total = 0; r->n = n; r->rawlen = n; r->pad = 0; r->endbit = 0;
}
__global__ void k_reverse_bytes(const u8* __restrict__ in, u32 n, u8* __restrict__ out) {
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[n - 1 - i] = in[i];
}
static void dec_attr_once() {
  static bool attr = false;
  if (attr) return;
  CUDA_CHECK(cudaFuncSetAttribute(k_hdec<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(HdecWarp)));
  CUDA_CHECK(cudaFuncSetAttribute(k_hdec<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(HdecWarp)));
  CUDA_CHECK(cudaFuncSetAttribute(k_hdec<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(HdecWarp)));
  CUDA_CHECK(cudaFuncSetAttribute(k_ibwt_chain, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(Seg) * IB_SEGS)));
  attr = true;
}
// d_L, d_out: device buffers of n bytes; 2 <= n <= 2^20 - 2; 0 <= pidx <= n
void bwt_inverse_sentinel(Ctx& c, const u8* d_L, u32 n, u32 pidx, u8* d_out) {
  dec_attr_once();
  DBuf<u32> keyA(c, SEG_SIZE), keyB(c, SEG_SIZE), valA(c, SEG_SIZE), valB(c, SEG_SIZE), dn(c, 1), nvis(c, 1);
  DBuf<u8> tmp(c, SEG_SIZE);
  DBuf<CandRes> res(c, 1);
  DBuf<Seg> segs(c, IB_SEGS);
  DBuf<Visit> visits(c, IB_VCAP);
  DBuf<u32> capr(c, IB_SEGS), tails(c, IB_VCAP), ntails(c, 1);
  DBuf<u8> slots(c, (size_t)SEG_SIZE * 4 + IB_CAP);  // slotA; the start row's slot sits behind it
  c.to_device(dn, &n, 4);
  k_ibwt_keys<<<(SEG_SIZE + 255) / 256, 256, 0, c.stream>>>(d_L, dn, SEG_SIZE, keyA);
  KLAUNCH(c); KCHECK();
  u32 *kin = keyA.p, *kout = keyB.p, *vin = valA.p, *vout = valB.p;
  radix_sort<u32>(c, kin, vin, kout, vout, dn.p, 1, SEG_SHIFT, n, 0, 1, true, n);  // swaps the pairs: vin = sorted positions
  u32* P = kout;  // the other key buffer is free now
  CUDA_CHECK(cudaMemsetAsync(P, 0, (size_t)SEG_SIZE * 4, c.stream));  // rows >= n lead back to row 0: the walks stay in bounds
  k_unbwt_pack<<<(n + 255) / 256, 256, 0, c.stream>>>(d_L, vin, n, pidx, P);
  KLAUNCH(c); KCHECK();
  k_unbwt_setup<<<1, 1, 0, c.stream>>>(res, n);
  KLAUNCH(c); KCHECK();
  u8* slotA = slots.p; u8* slotB = slots.p + (size_t)SEG_SIZE * 4;
  k_ibwt_walk1<<<(IB_SEGS + 127) / 128, 128, 0, c.stream>>>(P, res, 1, segs, capr, slotA, slotB);
  KLAUNCH(c); KCHECK();
  k_ibwt_chain<<<1, 128, sizeof(Seg) * IB_SEGS, c.stream>>>(P, res, 1, segs, visits, nvis, tails, ntails);
  KLAUNCH(c); KCHECK();
  k_ibwt_place<<<IB_PLACE_CTAS, IB_PLACE_THREADS, 0, c.stream>>>(P, res, 1, visits, nvis, slotA, slotB, tmp);
  KLAUNCH(c); KCHECK();
  k_ibwt_tail<<<(IB_VCAP + 127) / 128, 128, 0, c.stream>>>(P, res, 1, visits, nvis, tails, ntails, capr, tmp);
  KLAUNCH(c); KCHECK();
  k_reverse_bytes<<<(n + 255) / 256, 256, 0, c.stream>>>(tmp, n, d_out);
  KLAUNCH(c); KCHECK();
  c.sync();
}

// ---- RLE1 decode ------------------------------------------------------------------------------
#define UR_THREADS 256
#define UR_ITEMS 8
#define UR_TILE (UR_THREADS * UR_ITEMS)
#define UE_STAGE 6144u   // bytes of expanded output a tile stages in shared memory

__device__ __forceinline__ bool unrle_sync(const u8* b, u32 i) {
  if (i == 0) return true;
  if (b[i] == b[i - 1]) return false;
  if (i >= 4 && b[i - 1] == b[i - 2] && b[i - 2] == b[i - 3] && b[i - 3] == b[i - 4]) return false;
  return true;
}
// the reference's loop (lib/Bzip2.js:424-436) from a synchronisation point i to the next one: cls[j] = 1 for repeat counts
__device__ __noinline__ void unrle_walk(const u8* __restrict__ b, u8* __restrict__ c, u32 i, u32 n) {
  u32 j = i, run = 0;
  int prev = -1;
  for (;;) {
    const int v = b[j];
    c[j] = 0;
    run = (v == prev) ? run + 1 : 1;
    prev = v;
    j++;
    if (j >= n) break;
    if (run == 4) {
      c[j] = 1;
      j++;
      run = 0; prev = -1;
      if (j >= n) break;
    }
    if (unrle_sync(b, j)) break;
  }
}
// cls[i] = 1 when byte i is a repeat count.  A byte that differs from its predecessor and does not follow four equal
// bytes is certainly a literal that starts a new run (a synchronisation point); the bytes between two such points are
// classified by the serial loop.  One thread per 8 bytes: the 13 bytes that decide its synchronisation points sit in
// registers, a literal followed by another synchronisation point is final at once (all 8 of them: one 8-byte store),
// and only real run starts walk.
__global__ void __launch_bounds__(256) k_unrle_classify(const u8* __restrict__ rle, const CandRes* __restrict__ res, u32 ncand, u8* __restrict__ cls) {
  const u32 g = blockIdx.x * blockDim.x + threadIdx.x;
  const u32 ci = g >> (SEG_SHIFT - 3), i0 = (g & (SEG_MASK >> 3)) << 3;
  if (ci >= ncand) return;
  const CandRes* r = res + ci;
  if (r->status != 0 || i0 >= r->n) return;
  const u32 n = r->n;
  const u8* b = rle + ((size_t)ci << SEG_SHIFT);
  u8* c = cls + ((size_t)ci << SEG_SHIFT);
  const u64 prev = i0 ? *reinterpret_cast<const u64*>(b + i0 - 8) : 0ull;
  const u64 cur = *reinterpret_cast<const u64*>(b + i0);
  const u64 next = (i0 + 8 < SEG_SIZE) ? *reinterpret_cast<const u64*>(b + i0 + 8) : 0ull;
  // E bit (j + 3): byte i0 + j equals its predecessor, j = -3 .. 8 (bytes past n only make a position look like a run: it walks)
  u32 E = 0;
#pragma unroll
  for (int j = -3; j <= 8; j++) {
    const u32 x = j < 0 ? (u32)(prev >> (8 * (j + 8))) : (j < 8 ? (u32)(cur >> (8 * j)) : (u32)(next >> (8 * (j - 8))));
    const u32 y = (j - 1) < 0 ? (u32)(prev >> (8 * (j - 1 + 8))) : ((j - 1) < 8 ? (u32)(cur >> (8 * (j - 1))) : (u32)(next >> (8 * (j - 1 - 8))));
    if (((x ^ y) & 0xffu) == 0 && (int)i0 + j >= 1) E |= 1u << (j + 3);
  }
  // S bit k: position i0 + k is a synchronisation point, k = 0 .. 8
  u32 S = 0;
#pragma unroll
  for (int k = 0; k <= 8; k++) {
    const bool e0 = (E >> (k + 3)) & 1u, run4 = ((E >> k) & 7u) == 7u && i0 + k >= 4;  // E(i-3), E(i-2), E(i-1)
    if (!e0 && !run4) S |= 1u << k;
  }
  if ((S & 0xffu) == 0xffu && i0 + 8 <= n) {
    // eight literals, each followed by a synchronisation point or the end... unless the ninth position continues a run of
    // the eighth: then position 7 starts a walk
    if ((S >> 8) & 1u || i0 + 8 >= n) { *reinterpret_cast<u64*>(c + i0) = 0ull; return; }
    *reinterpret_cast<u32*>(c + i0) = 0u;
    c[i0 + 4] = 0; c[i0 + 5] = 0; c[i0 + 6] = 0;
    unrle_walk(b, c, i0 + 7, n);
    return;
  }
#pragma unroll
  for (int k = 0; k < 8; k++) {
    const u32 i = i0 + k;
    if (i < n && ((S >> k) & 1u)) {
      if (((S >> (k + 1)) & 1u) || i + 1 >= n) c[i] = 0;
      else unrle_walk(b, c, i, n);
    }
  }
}

// expanded size of every tile (no chain between tiles: the per-block scan below is a separate, tiny kernel)
__global__ void __launch_bounds__(UR_THREADS)
k_unrle_tilesum(const u8* __restrict__ rle, const u8* __restrict__ cls, const CandRes* __restrict__ res, u32 tps, u32* __restrict__ tilesum) {
  __shared__ u32 ws[UR_THREADS / 32];
  const u32 tid = threadIdx.x;
  const u32 ci = blockIdx.x / tps, lt = blockIdx.x % tps;
  const CandRes* r = res + ci;
  if (r->status != 0) return;
  const u32 n = r->n;
  const u32 start = lt * UR_TILE;
  if (start >= n) return;
  const u8* b = rle + ((size_t)ci << SEG_SHIFT);
  const u8* c = cls + ((size_t)ci << SEG_SHIFT);
  u32 sum = 0;
  const u32 p0 = start + tid * UR_ITEMS;
  static_assert(UR_ITEMS == 8, "one 8-byte load per array");
  if (p0 < n) {
    const u64 bv = *reinterpret_cast<const u64*>(b + p0), cv = *reinterpret_cast<const u64*>(c + p0);
#pragma unroll
    for (int j = 0; j < UR_ITEMS; j++)
      if (p0 + j < n) sum += ((u32)(cv >> (8 * j)) & 0xffu) ? ((u32)(bv >> (8 * j)) & 0xffu) : 1u;
  }
  sum = warp_reduce_add(sum);
  if ((tid & 31u) == 0) ws[tid >> 5] = sum;
  __syncthreads();
  if (tid < 32) {
    u32 v = tid < UR_THREADS / 32 ? ws[tid] : 0u;
    v = warp_reduce_add(v);
    if (tid == 0) tilesum[(size_t)ci * tps + lt] = v;
  }
}
// one warp per block: output offset of every tile, decoded size of the block
__global__ void __launch_bounds__(32)
k_unrle_tileoff(CandRes* __restrict__ res, u32 tps, const u32* __restrict__ tilesum, u32* __restrict__ tileoff) {
  const u32 ci = blockIdx.x, lane = threadIdx.x;
  CandRes* r = res + ci;
  if (r->status != 0) return;
  const u32 n = r->n;
  const u32 nt = (n + UR_TILE - 1) / UR_TILE;
  u32 run = 0;
  for (u32 t0 = 0; t0 < nt; t0 += 32) {
    const u32 t = t0 + lane;
    const u32 v = t < nt ? tilesum[(size_t)ci * tps + t] : 0u;
    const u32 inc = warp_incl_add(v);
    if (t < nt) tileoff[(size_t)ci * tps + t] = run + inc - v;
    run += __shfl_sync(FULL_MASK, inc, 31);
  }
  if (lane == 0) r->rawlen = run;
}

__global__ void __launch_bounds__(UR_THREADS)
k_unrle_emit(const u8* __restrict__ rle, const u8* __restrict__ cls, const CandRes* __restrict__ res, u32 tps, const u32* __restrict__ tileoff,
             const u64* __restrict__ outbase, u8* __restrict__ out) {
  __shared__ u32 ws[UR_THREADS / 32 + 1];
  __shared__ __align__(16) u8 stg[UE_STAGE + 32];
  const u32 tid = threadIdx.x;
  const u32 ci = blockIdx.x / tps, lt = blockIdx.x % tps;
  const u64 ob = outbase[ci];
  if (ob == ~0ull) return;  // candidate is not part of the stream
  const CandRes* r = res + ci;
  const u32 n = r->n;
  const u32 start = lt * UR_TILE;
  if (start >= n) return;
  const u8* b = rle + ((size_t)ci << SEG_SHIFT);
  const u8* c = cls + ((size_t)ci << SEG_SHIFT);
  u32 len[UR_ITEMS];
  u32 sum = 0;
  const u32 p0 = start + tid * UR_ITEMS;
  u64 bv = 0, cv = 0;
  if (p0 < n) { bv = *reinterpret_cast<const u64*>(b + p0); cv = *reinterpret_cast<const u64*>(c + p0); }
#pragma unroll
  for (int j = 0; j < UR_ITEMS; j++) {
    const u32 p = p0 + j;
    len[j] = (p < n) ? (((u32)(cv >> (8 * j)) & 0xffu) ? ((u32)(bv >> (8 * j)) & 0xffu) : 1u) : 0u;
    sum += len[j];
  }
  u32 total;
  const u32 ex = block_excl_add<UR_THREADS, u32>(sum, ws, &total);
  u8* otile = out + ob + tileoff[(size_t)ci * tps + lt];
  // A tile that expands to at most UE_STAGE bytes (every tile of run-free data, most others) is put together in shared
  // memory, at the alignment (mod 16) it has in the output, and leaves in 16-byte stores; longer ones go out directly.
  const bool staged = total <= UE_STAGE;
  const u32 mis = (u32)(size_t)otile & 15u;
  u8* o = staged ? stg + mis + ex : otile + ex;
#pragma unroll
  for (int j = 0; j < UR_ITEMS; j++) {
    const u32 p = p0 + j;
    if (p < n) {
      if ((u32)(cv >> (8 * j)) & 0xffu) {
        const u8 v = b[p - 1];
        for (u32 x = 0; x < len[j]; x++) o[x] = v;
      } else {
        o[0] = (u8)(bv >> (8 * j));
      }
      o += len[j];
    }
  }
  if (!staged) return;
  __syncthreads();
  {
    const u32 last = mis + total;
    u8* og = otile - mis;  // 16-byte aligned
    for (u32 c16 = tid * 16u; c16 < last; c16 += UR_THREADS * 16u) {
      if (c16 >= mis && c16 + 16u <= last) {
        *reinterpret_cast<uint4*>(og + c16) = *reinterpret_cast<const uint4*>(stg + c16);
      } else {
        const u32 e = min(c16 + 16u, last);
        for (u32 x = max(c16, mis); x < e; x++) og[x] = stg[x];
      }
    }
  }
}

// ---- host ---------------------------------------------------------------------------------------
static std::string hexs(u32 v) {
  char b[16];
  snprintf(b, sizeof b, "%x", v);
  return b;
}

struct Event { int kind; size_t cand; u32 a, b; int code; std::string msg; };  // kind: 0 block, 1 eos, 2 error

// One decode in flight.  open() parses the header, finds every block candidate and decodes the share
// [lo, hi) of them (all of them on one GPU); finish() walks the chain over ALL candidates' results
// (imported from the other ranks when sharded), expands + CRC-checks the blocks of the own share and
// raises the reference's errors in stream order.
struct DecSession {
  Ctx* c = nullptr;
  size_t n = 0;
  u32 dbuf_size = 0;
  DBuf<u8> din;
  std::vector<Cand> cands;         // every magic found, sorted by position
  std::vector<size_t> blk_idx;     // block candidates (index into cands)
  std::vector<Cand> bc;            // the same as Cand records
  std::vector<CandRes> hres;       // per block candidate (valid for [lo,hi) after open, for all after import)
  size_t lo = 0, hi = 0;           // own share of the block candidates
  bool single = false, eos_single = false;
  DBuf<Cand> dcand;
  DBuf<CandRes> dres;              // own share only
  DBuf<u8> rle, cls;               // cls (count-byte classes) is kept for the whole share only while that is cheap (keep_cls)
  bool keep_cls = true;
  DBuf<u32> tileoff;
  int err_event = -1;              // index of the first failing event (sharded mode)
};

static const u32 UR_TPS = SEG_SIZE / UR_TILE;
#define DEC_KEEP_CLS 16384u
// blocks per decode batch ($B2_DEC_BATCH: test hook, small batches exercise the batch seams on small inputs)
static u32 dec_batch_blocks(const Ctx& c) {
  if (const char* e = getenv("B2_DEC_BATCH")) { const int v = atoi(e); if (v >= 1) return (u32)v; }
  return std::max(c.bwt_batch, 2048u);
}
static u32 dec_keep_cls_limit() {
  if (const char* e = getenv("B2_DEC_KEEP_CLS")) { const int v = atoi(e); if (v >= 0) return (u32)v; }  // test hook
  return DEC_KEEP_CLS;
}

static void dec_open(Ctx& c, DecSession& S, const u8* d_in_user, size_t n, bool single_block, u64 bitpos, int rank, int world) {
  S.c = &c; S.n = n; S.single = single_block;
  // padded private copy of the input (aligned word reads past the end must be safe)
  S.din.alloc(c, n + 32);
  CUDA_CHECK(cudaMemsetAsync(S.din.p + (n & ~(size_t)3), 0, (n + 32) - (n & ~(size_t)3), c.stream));
  if (n) CUDA_CHECK(cudaMemcpyAsync(S.din, d_in_user, n, cudaMemcpyDeviceToDevice, c.stream));
  u8 hdr[4] = {0, 0, 0, 0};
  if (n >= 4) CUDA_CHECK(cudaMemcpyAsync(hdr, S.din, 4, cudaMemcpyDeviceToHost, c.stream));
  CUDA_CHECK(cudaStreamSynchronize(c.stream));
  // lib/Bzip2.js:105-124 _start_bunzip
  if (n < 4 || hdr[0] != 'B' || hdr[1] != 'Z' || hdr[2] != 'h') throw B2Error{DEC_NOT_BZIP, "Not bzip data: bad magic"};
  int level = hdr[3] - 0x30;
  if (level < 1 || level > 9) throw B2Error{DEC_NOT_BZIP, "Not bzip data: level out of range"};
  S.dbuf_size = 100000u * (u32)level;
  // The kernels decode every candidate under the largest block size: the members of a multistream file may have
  // different levels (lib/Bzip2.js:105-124 re-reads the level per member), and which member a candidate belongs to is
  // only known when the host walks the chain, where the member's own limit is applied (dec_finish).
  const u32 dbuf_size = 900000u;
  const u8* din = S.din;

  // ---- 1. candidates ----
  std::vector<Cand>& cands = S.cands;
  {
    StageScope ss(c, ST_SCAN);
    // Highly repetitive input compresses to a few dozen bytes per block (and multistream files may hold thousands of
    // tiny members), so the number of magics is not bounded by the usual ~100 KB per block: when the first guess is too
    // small the scan counts them all and runs once more with exactly that capacity.
    u32 cap = (u32)(n / 8000 + 1024);
    DBuf<Cand> dc(c, cap);
    DBuf<u32> dcount(c, 1);
    u32 cnt = 0;
    for (int attempt = 0; attempt < 2; attempt++) {
      CUDA_CHECK(cudaMemsetAsync(dcount, 0, 4, c.stream));
      k_scan_magic<<<(unsigned)(((n + 3) / 4 + 255) / 256), 256, 0, c.stream>>>(din, n, dc, dcount, cap);
      KLAUNCH(c); KCHECK();
      CUDA_CHECK(cudaMemcpyAsync(&cnt, dcount, 4, cudaMemcpyDeviceToHost, c.stream));
      CUDA_CHECK(cudaStreamSynchronize(c.stream));
      if (cnt <= cap) break;
      cap = cnt;
      dc.alloc(c, cap);
    }
    if (cnt > cap) throw B2Error{B2_ERR_CUDA, "magic scan did not settle"};
    cands.resize(cnt);
    if (cnt) CUDA_CHECK(cudaMemcpyAsync(cands.data(), dc, sizeof(Cand) * cnt, cudaMemcpyDeviceToHost, c.stream));
    CUDA_CHECK(cudaStreamSynchronize(c.stream));
    std::sort(cands.begin(), cands.end(), [](const Cand& a, const Cand& b) { return a.pos < b.pos; });
  }
  std::vector<size_t>& blk_idx = S.blk_idx;
  if (single_block) {
    // lib/Bzip2.js:482-503: seekBit(pos) then one _get_next_block
    const Cand* hit = nullptr;
    for (auto& cd : cands) if (cd.pos == bitpos) hit = &cd;
    if (!hit) throw B2Error{DEC_NOT_BZIP, "Not bzip data"};
    if (hit->type == 2) { S.eos_single = true; return; }
    blk_idx.push_back((size_t)(hit - cands.data()));
  } else {
    for (size_t i = 0; i < cands.size(); i++) if (cands[i].type == 1) blk_idx.push_back(i);
  }
  const size_t nb_all = blk_idx.size();
  S.bc.resize(nb_all);
  for (size_t i = 0; i < nb_all; i++) S.bc[i] = cands[blk_idx[i]];
  S.hres.assign(nb_all, CandRes());
  for (auto& r : S.hres) { memset(&r, 0, sizeof r); r.status = DEC_DATA_ERROR; }
  S.lo = (size_t)rank * nb_all / (size_t)world;
  S.hi = (size_t)(rank + 1) * nb_all / (size_t)world;
  const size_t nb = S.hi - S.lo;
  {
    // every block of the own share keeps 2 MiB (L column + count-byte classes; 1 MiB beyond DEC_KEEP_CLS blocks) until the
    // stream is assembled, and a batch of up to 2048 blocks needs ~20 MiB of scratch per block: say so instead of failing
    // inside an allocation
    const size_t need = nb * ((size_t)(nb <= dec_keep_cls_limit() ? 2 : 1) << 20) + std::min<size_t>(nb, dec_batch_blocks(c)) * ((size_t)20 << 20) + n;
    // memory the stream-ordered pool holds but does not use is available too: when that covers the call (every call
    // after the first of a kind) the driver is not asked at all -- cudaMemGetInfo takes milliseconds on a busy context
    uint64_t reserved = 0, used = 0;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, c.device) == cudaSuccess) {
      cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReservedMemCurrent, &reserved);
      cudaMemPoolGetAttribute(pool, cudaMemPoolAttrUsedMemCurrent, &used);
    }
    cudaGetLastError();
    const size_t spare = (size_t)(reserved > used ? reserved - used : 0);
    size_t free_b = 0, total_b = 0;
    if (need > spare) {
      if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) {
        if (need > free_b + spare) {
          char msg[256];
          snprintf(msg, sizeof msg, "stream of %zu blocks needs about %zu MiB of device memory for one call (%zu MiB free): decode it in parts "
                                    "(Bzip2.table + decompressBlock) or over several GPUs (decompress_file_sharded)", nb, need >> 20, free_b >> 20);
          throw B2Error{B2_ERR_CUDA, msg};
        }
      } else cudaGetLastError();
    }
  }
  S.dcand.alloc(c, nb_all ? nb_all : 1);
  S.dres.alloc(c, nb ? nb : 1);
  S.rle.alloc(c, (nb ? nb : 1) << SEG_SHIFT);
  // The count-byte classes of a block are needed twice (length scan here, expansion in dec_finish).  Up to DEC_KEEP_CLS
  // blocks they stay on the device in between; a stream of more blocks (tens of GB of level-1 data) keeps only the L
  // columns (1 MiB per block) and classifies a second time, batch by batch, when it expands them.
  S.keep_cls = nb <= dec_keep_cls_limit();
  const u32 DBc = dec_batch_blocks(c);
  S.cls.alloc(c, (size_t)(S.keep_cls ? (nb ? nb : 1) : std::min<size_t>(nb, DBc)) << SEG_SHIFT);
  S.tileoff.alloc(c, (nb ? nb : 1) * (size_t)UR_TPS);
  if (nb_all) CUDA_CHECK(cudaMemcpyAsync(S.dcand, S.bc.data(), sizeof(Cand) * nb_all, cudaMemcpyHostToDevice, c.stream));
  CUDA_CHECK(cudaStreamSynchronize(c.stream));
  CandRes* hres = S.hres.data() + S.lo;  // own share
  Cand* dcand = S.dcand.p + S.lo;
  DBuf<CandRes>& dres = S.dres;
  DBuf<u8>& rle = S.rle; DBuf<u8>& cls = S.cls; DBuf<u32>& tileoff = S.tileoff;
  const u32 ur_tps = UR_TPS;

  // ---- 2. decode the own share of the candidate blocks, in batches ----
  // the per-block Huffman stage is one CTA per block and latency bound: give it every block at once
  // (about 19 MB of scratch per block; 180 GB of HBM take thousands)
  const u32 DB = dec_batch_blocks(c);
  dec_attr_once();
  if (nb) {
    const u32 nbm = (u32)std::min<size_t>(DB, nb);
    DBuf<u16> sym(c, (size_t)nbm << SEG_SHIFT);
    DBuf<u8> selbuf(c, (size_t)nbm * SEL_CAP), tt(c, (size_t)nbm << SEG_SHIFT), symb(c, (size_t)nbm << SEG_SHIFT);
    const u32 cps = SEG_SIZE / UM_CHUNK;
    DBuf<ChunkSum> sums(c, (size_t)nbm * cps);
    DBuf<u8> perms(c, (size_t)nbm * cps * 256), lists(c, (size_t)nbm * cps * 256);
    DBuf<ChunkStart> starts(c, (size_t)nbm * cps);
    DBuf<u32> keyA(c, (size_t)nbm << SEG_SHIFT), keyB(c, (size_t)nbm << SEG_SHIFT), valA(c, (size_t)nbm << SEG_SHIFT), valB(c, (size_t)nbm << SEG_SHIFT);
    DBuf<u32> dn(c, nbm), nvis(c, nbm), tilesum(c, (size_t)nbm * ur_tps);
    DBuf<Seg> segs(c, (size_t)nbm * IB_SEGS);
    DBuf<Visit> visits(c, (size_t)nbm * IB_VCAP);
    DBuf<u32> capr(c, (size_t)nbm * IB_SEGS), tails(c, (size_t)nbm * IB_VCAP), ntails(c, nbm);
    std::vector<u32> hn(nbm);
    for (size_t k0 = 0; k0 < nb; k0 += DB) {
      const u32 cnt = (u32)std::min<size_t>(DB, nb - k0);
      CandRes* rb = dres.p + k0;
      {
        StageScope ss(c, ST_HDEC);
        static int sms = 0;
        if (!sms) CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c.device));
        if (cnt <= 2u * (u32)sms) k_hdec<512><<<cnt, 512, sizeof(HdecWarp), c.stream>>>(din, n, dcand, (u32)k0, cnt, dbuf_size, selbuf, sym, rb);
        else if (cnt <= 4u * (u32)sms) k_hdec<256><<<cnt, 256, sizeof(HdecWarp), c.stream>>>(din, n, dcand, (u32)k0, cnt, dbuf_size, selbuf, sym, rb);
        else k_hdec<HD_THREADS><<<cnt, HD_THREADS, sizeof(HdecWarp), c.stream>>>(din, n, dcand, (u32)k0, cnt, dbuf_size, selbuf, sym, rb);
        KLAUNCH(c); KCHECK();
      }
      {
        StageScope ss(c, ST_UNMTF);
        const u32 chunks = cnt * cps;
        k_unmtf_a<<<(chunks + UA_THREADS - 1) / UA_THREADS, UA_THREADS, 0, c.stream>>>(sym, rb, cnt, cps, sums, perms, symb);
        KLAUNCH(c); KCHECK();
        k_unmtf_scan<<<cnt, 32, 0, c.stream>>>(rb, cnt, cps, dbuf_size, sums, perms, lists, starts);
        KLAUNCH(c); KCHECK();
        k_unmtf_map<<<(chunks + UM_WARPS - 1) / UM_WARPS, UM_WARPS * 32, 0, c.stream>>>(sym, symb, rb, cnt, cps, lists, starts, tt);
        KLAUNCH(c); KCHECK();
      }
      CUDA_CHECK(cudaMemcpyAsync(hres + k0, rb, sizeof(CandRes) * cnt, cudaMemcpyDeviceToHost, c.stream));
      CUDA_CHECK(cudaStreamSynchronize(c.stream));
      u32 nmax = 0; u64 ntot = 0;
      for (u32 i = 0; i < cnt; i++) { hn[i] = hres[k0 + i].status == 0 ? hres[k0 + i].n : 0; nmax = std::max(nmax, hn[i]); ntot += hn[i]; }
      if (nmax) {
        StageScope ss(c, ST_IBWT);
        CUDA_CHECK(cudaMemcpyAsync(dn, hn.data(), cnt * 4, cudaMemcpyHostToDevice, c.stream));
        // T-vector: one stable counting-sort pass over the L column itself (byte keys, the values are the row numbers);
        // the pass writes P[row] = successor << 8 | L[row] directly (radix.cuh: pack epilogue)
        u8* kin = tt.p; u8* kout = nullptr;
        u32 *vin = valA, *vout = valB;
        u32* Pp = keyB;
        radix_sort<u8>(c, kin, vin, kout, vout, dn, cnt, SEG_SHIFT, nmax, 0, 1, true, ntot, nullptr, tt.p, Pp);
        // all blocks of the batch walk together: the launch lasts as long as its longest segment walk, so fewer,
        // bigger launches win over keeping the packed T-vectors L2 resident (measured: 75 ms -> 36 ms per GiB)
        const u32 ib_sub = cnt;
        for (u32 s0 = 0; s0 < cnt; s0 += ib_sub) {
          const u32 sc = std::min<u32>(ib_sub, cnt - s0);
          const u32* Ps = Pp + ((size_t)s0 << SEG_SHIFT);
          // the walks record into the free key / value buffers of the sort (4 MiB per block each)
          u8* slotA = reinterpret_cast<u8*>(keyA.p) + ((size_t)s0 << (SEG_SHIFT + 2));
          u8* slotB = reinterpret_cast<u8*>(valA.p) + ((size_t)s0 << (SEG_SHIFT + 2));
          k_ibwt_walk1<<<(sc * IB_SEGS + 127) / 128, 128, 0, c.stream>>>(Ps, rb + s0, sc, segs.p + (size_t)s0 * IB_SEGS, capr.p + (size_t)s0 * IB_SEGS, slotA, slotB);
          KLAUNCH(c); KCHECK();
          k_ibwt_chain<<<sc, 128, sizeof(Seg) * IB_SEGS, c.stream>>>(Ps, rb + s0, sc, segs.p + (size_t)s0 * IB_SEGS, visits.p + (size_t)s0 * IB_VCAP, nvis.p + s0,
                                                                   tails.p + (size_t)s0 * IB_VCAP, ntails.p + s0);
          KLAUNCH(c); KCHECK();
          u8* ob = rle.p + ((k0 + s0) << SEG_SHIFT);
          k_ibwt_place<<<sc * IB_PLACE_CTAS, IB_PLACE_THREADS, 0, c.stream>>>(Ps, rb + s0, sc, visits.p + (size_t)s0 * IB_VCAP, nvis.p + s0, slotA, slotB, ob);
          KLAUNCH(c); KCHECK();
          k_ibwt_tail<<<(sc * IB_VCAP + 127) / 128, 128, 0, c.stream>>>(Ps, rb + s0, sc, visits.p + (size_t)s0 * IB_VCAP, nvis.p + s0,
                                                                      tails.p + (size_t)s0 * IB_VCAP, ntails.p + s0, capr.p + (size_t)s0 * IB_SEGS, ob);
          KLAUNCH(c); KCHECK();
        }
      }
      if (nmax) {
        StageScope ss(c, ST_UNRLE);
        const u32 nslots = cnt << SEG_SHIFT;
        u8* clsb = cls.p + (S.keep_cls ? (k0 << SEG_SHIFT) : 0);
        k_unrle_classify<<<(nslots / 8 + 255) / 256, 256, 0, c.stream>>>(rle.p + (k0 << SEG_SHIFT), rb, cnt, clsb);
        KLAUNCH(c); KCHECK();
        k_unrle_tilesum<<<cnt * ur_tps, UR_THREADS, 0, c.stream>>>(rle.p + (k0 << SEG_SHIFT), clsb, rb, ur_tps, tilesum);
        KLAUNCH(c); KCHECK();
        k_unrle_tileoff<<<cnt, 32, 0, c.stream>>>(rb, ur_tps, tilesum, tileoff.p + k0 * ur_tps);
        KLAUNCH(c); KCHECK();
        CUDA_CHECK(cudaMemcpyAsync(hres + k0, rb, sizeof(CandRes) * cnt, cudaMemcpyDeviceToHost, c.stream));
      }
      CUDA_CHECK(cudaStreamSynchronize(c.stream));
      c.stats.blocks += cnt;
    }
  }

}

// fills *first_err (event index, or -1) instead of throwing when `sharded`
static int dec_finish(Ctx& c, DecSession& S, int multistream, u8* d_out, size_t out_cap, size_t* out_n, std::vector<u64>* tab_pos,
                      std::vector<u32>* tab_len, u8** d_out_alloc, bool sharded, u64* shard_info) {
  *out_n = 0;
  if (d_out_alloc) *d_out_alloc = nullptr;
  if (S.eos_single) return 0;
  const size_t n = S.n;
  const size_t nb_all = S.blk_idx.size();
  std::vector<Cand>& cands = S.cands;
  std::vector<Cand>& bc = S.bc;
  std::vector<CandRes>& hres = S.hres;
  u32 cur_dbuf = S.dbuf_size;  // dbufSize of the member the walk is in (lib/Bzip2.js:121)
  const u32 ur_tps = UR_TPS;
  // ---- 3. walk the chain in stream order (lib/Bzip2.js:454-481 / 508-548) ----
  std::vector<Event> events;
  std::vector<u64> outbase(nb_all ? nb_all : 1, ~0ull);
  u64 total_out = 0;
  auto find_cand = [&](u64 pos) -> long {
    size_t lo = 0, hi = cands.size();
    while (lo < hi) { size_t mid = (lo + hi) / 2; if (cands[mid].pos < pos) lo = mid + 1; else hi = mid; }
    return (lo < cands.size() && cands[lo].pos == pos) ? (long)lo : -1;
  };
  std::vector<long> cand_to_blk(cands.size(), -1);
  for (size_t i = 0; i < nb_all; i++) cand_to_blk[S.blk_idx[i]] = (long)i;
  auto block_event = [&](size_t bi) -> bool {  // returns false when the walk must stop (error recorded)
    const CandRes& r = hres[bi];
    // the member's own limits, in the reference's order: randomised bit (:143), origPointer (:146), then the body
    if (r.status != DEC_OBSOLETE && r.orig > cur_dbuf) {
      events.push_back({2, bi, 0, 0, DEC_DATA_ERROR, "Data error: initial position out of bounds"});
      return false;
    }
    if (r.status != 0) {
      std::string msg = r.status == DEC_OBSOLETE ? "Obsolete (pre 0.9.5) bzip format not supported." : "Data error";
      if (r.detail == 1) msg += ": initial position out of bounds";
      events.push_back({2, bi, 0, 0, r.status, msg});
      return false;
    }
    if (r.n > cur_dbuf) {  // dbufCount would have run over dbufSize (lib/Bzip2.js:338,354)
      events.push_back({2, bi, 0, 0, DEC_DATA_ERROR, "Data error"});
      return false;
    }
    outbase[bi] = total_out;
    total_out += r.rawlen;
    events.push_back({0, bi, 0, 0, 0, ""});
    return true;
  };
  if (S.single) {
    block_event(0);
  } else {
    u64 pos = 32;
    u32 stream_crc = 0;
    for (;;) {
      if ((pos + 7) / 8 >= n) break;  // 'eof' in inputStream && inputStream.eof() (lib/Bzip2.js:462)
      const long ci = find_cand(pos);
      if (ci < 0) { events.push_back({2, 0, 0, 0, DEC_NOT_BZIP, "Not bzip data"}); break; }
      if (cands[ci].type == 1) {
        const size_t bi = (size_t)cand_to_blk[ci];
        stream_crc = cands[ci].next32 ^ ((stream_crc << 1) | (stream_crc >> 31));  // lib/Bzip2.js:138-139
        if (!block_event(bi)) break;
        pos = hres[bi].endbit;
      } else {
        events.push_back({1, 0, stream_crc, cands[ci].next32, 0, ""});
        pos += 80;
        const u64 bytepos = (pos + 7) / 8;
        if (multistream && bytepos < n) {
          // _start_bunzip on the byte stream (resyncs to the next byte)
          u8 h2[4] = {0, 0, 0, 0};
          const size_t avail = (size_t)std::min<u64>(4, n - bytepos);
          CUDA_CHECK(cudaMemcpyAsync(h2, S.din.p + bytepos, avail, cudaMemcpyDeviceToHost, c.stream));
          CUDA_CHECK(cudaStreamSynchronize(c.stream));
          if (avail != 4 || h2[0] != 'B' || h2[1] != 'Z' || h2[2] != 'h') { events.push_back({2, 0, 0, 0, DEC_NOT_BZIP, "Not bzip data: bad magic"}); break; }
          const int lv = h2[3] - 0x30;
          if (lv < 1 || lv > 9) { events.push_back({2, 0, 0, 0, DEC_NOT_BZIP, "Not bzip data: level out of range"}); break; }
          cur_dbuf = (u32)lv * 100000u;
          stream_crc = 0;
          pos = (bytepos + 4) * 8;
        } else break;
      }
    }
  }

  // ---- 4. expand the chain blocks of the own share, CRC them ----
  // own output window: [my_off, my_off + my_len) of the decoded stream
  u64 my_off = 0, my_len = 0;
  {
    bool first = true;
    for (size_t i = S.lo; i < S.hi; i++)
      if (outbase[i] != ~0ull) {
        if (first) { my_off = outbase[i]; first = false; }
        my_len = outbase[i] + hres[i].rawlen - my_off;
      }
  }
  const size_t nb = S.hi - S.lo;
  u8* dout = d_out;
  DBuf<u8> own;
  if (!d_out) {
    own.alloc(c, my_len ? my_len : 1);
    dout = own.p;
  } else if (my_len > out_cap) {
    *out_n = (size_t)my_len;
    throw B2Error{B2_ERR_BAD_ARG, "output buffer too small"};
  }
  std::vector<u32> got_crc(nb ? nb : 1, 0);
  if (nb) {
    StageScope ss(c, ST_UNRLE);
    std::vector<u64> ob(nb);
    for (size_t i = 0; i < nb; i++) ob[i] = outbase[S.lo + i] == ~0ull ? ~0ull : outbase[S.lo + i] - my_off;
    DBuf<u64> dob(c, nb);
    CUDA_CHECK(cudaMemcpyAsync(dob, ob.data(), 8 * nb, cudaMemcpyHostToDevice, c.stream));
    if (S.keep_cls) {
      k_unrle_emit<<<(unsigned)(nb * ur_tps), UR_THREADS, 0, c.stream>>>(S.rle, S.cls, S.dres, ur_tps, S.tileoff, dob, dout);
      KLAUNCH(c); KCHECK();
    } else {
      const size_t DBc = dec_batch_blocks(c);
      for (size_t k0 = 0; k0 < nb; k0 += DBc) {
        const u32 cnt = (u32)std::min<size_t>(DBc, nb - k0);
        k_unrle_classify<<<(unsigned)((((size_t)cnt << SEG_SHIFT) / 8 + 255) / 256), 256, 0, c.stream>>>(S.rle.p + (k0 << SEG_SHIFT), S.dres.p + k0, cnt, S.cls);
        KLAUNCH(c); KCHECK();
        k_unrle_emit<<<(unsigned)(cnt * ur_tps), UR_THREADS, 0, c.stream>>>(S.rle.p + (k0 << SEG_SHIFT), S.cls, S.dres.p + k0, ur_tps, S.tileoff.p + k0 * ur_tps,
                                                                            dob.p + k0, dout);
        KLAUNCH(c); KCHECK();
      }
    }
    std::vector<BlkInfo> ranges(nb);
    for (size_t i = 0; i < nb; i++) {
      memset(&ranges[i], 0, sizeof(BlkInfo));
      if (ob[i] != ~0ull) { ranges[i].s = ob[i]; ranges[i].e = ob[i] + hres[S.lo + i].rawlen; }
    }
    DBuf<BlkInfo> dr(c, nb);
    DBuf<u32> dcrc(c, nb);
    CUDA_CHECK(cudaMemcpyAsync(dr, ranges.data(), sizeof(BlkInfo) * nb, cudaMemcpyHostToDevice, c.stream));
    crc_ranges(c, dout, dr, ranges, dcrc);
    CUDA_CHECK(cudaMemcpyAsync(got_crc.data(), dcrc, 4 * nb, cudaMemcpyDeviceToHost, c.stream));
    CUDA_CHECK(cudaStreamSynchronize(c.stream));
  }
  // ---- 5. replay the events: first failure in stream order wins ----
  S.err_event = -1;
  int err_code = 0;
  std::string err_msg;
  for (size_t ei = 0; ei < events.size() && S.err_event < 0; ei++) {
    const Event& ev = events[ei];
    if (ev.kind == 0) {
      if (ev.cand >= S.lo && ev.cand < S.hi) {  // CRCs of foreign blocks are checked by their owners
        const u32 want = bc[ev.cand].next32, got = got_crc[ev.cand - S.lo];
        if (want != got) { S.err_event = (int)ei; err_code = DEC_DATA_ERROR; err_msg = "Data error: Bad block CRC (got " + hexs(got) + " expected " + hexs(want) + ")"; }
      }
      if (S.err_event < 0 && tab_pos) { tab_pos->push_back(bc[ev.cand].pos); tab_len->push_back(hres[ev.cand].rawlen); }
    } else if (ev.kind == 1) {
      if (!tab_pos && ev.a != ev.b) { S.err_event = (int)ei; err_code = DEC_DATA_ERROR; err_msg = "Data error: Bad stream CRC (got " + hexs(ev.a) + " expected " + hexs(ev.b) + ")"; }
    } else {
      S.err_event = (int)ei; err_code = ev.code; err_msg = ev.msg;
    }
  }
  if (shard_info) { shard_info[0] = my_off; shard_info[1] = my_len; shard_info[2] = total_out; shard_info[3] = (u64)(long long)S.err_event; shard_info[4] = (u64)(long long)err_code; }
  if (S.err_event >= 0) {
    if (!sharded) throw B2Error{err_code, err_msg};
    throw B2Error{err_code, err_msg};  // the caller (sharded) compares shard_info[3] across ranks and keeps the earliest
  }
  *out_n = (size_t)my_len;
  if (!d_out && d_out_alloc) { *d_out_alloc = own.p; own.p = nullptr; }
  return 0;
}

int bzip2_decompress_device(Ctx& c, const u8* d_in_user, size_t n, int multistream, u8* d_out, size_t out_cap, size_t* out_n, bool single_block,
                            u64 bitpos, std::vector<u64>* tab_pos, std::vector<u32>* tab_len, u8** d_out_alloc) {
  *out_n = 0;
  if (d_out_alloc) *d_out_alloc = nullptr;
  DecSession S;
  dec_open(c, S, d_in_user, n, single_block, bitpos, 0, 1);
  return dec_finish(c, S, multistream, d_out, out_cap, out_n, tab_pos, tab_len, d_out_alloc, false, nullptr);
}

// ---- sharded decode: open on every rank, exchange results, finish ------------
static DecSession* g_shard = nullptr;
void dec_shard_open(Ctx& c, const u8* d_in, size_t n, int rank, int world, u64* info) {
  delete g_shard;
  g_shard = new DecSession();
  dec_open(c, *g_shard, d_in, n, false, 0, rank, world);
  info[0] = g_shard->blk_idx.size(); info[1] = g_shard->lo; info[2] = g_shard->hi;
}
void dec_shard_export(u64* buf) {
  if (!g_shard) throw B2Error{B2_ERR_BAD_ARG, "no sharded decode in flight"};
  for (size_t i = g_shard->lo; i < g_shard->hi; i++) {
    const CandRes& r = g_shard->hres[i];
    u64* o = buf + (i - g_shard->lo) * 6;
    o[0] = (u64)(long long)r.status; o[1] = r.detail; o[2] = r.endbit; o[3] = r.n; o[4] = r.rawlen; o[5] = r.orig;
  }
}
int dec_shard_finish(Ctx& c, const u64* all, int multistream, u8* d_out, size_t out_cap, u64* res) {
  if (!g_shard) throw B2Error{B2_ERR_BAD_ARG, "no sharded decode in flight"};
  DecSession& S = *g_shard;
  for (size_t i = 0; i < S.hres.size(); i++) {
    if (i >= S.lo && i < S.hi) continue;
    const u64* o = all + i * 6;
    CandRes& r = S.hres[i];
    r.status = (int)(long long)o[0]; r.detail = (u32)o[1]; r.endbit = o[2]; r.n = (u32)o[3]; r.rawlen = (u32)o[4]; r.orig = (u32)o[5];
  }
  size_t out_n = 0;
  struct Closer { ~Closer() { delete g_shard; g_shard = nullptr; } } closer;
  return dec_finish(c, S, multistream, d_out, out_cap, &out_n, nullptr, nullptr, nullptr, true, res);
}
