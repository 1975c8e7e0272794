// bwtc_core.cuh -- the serial half of compressjs' BWTC container (lib/BWTC.js), written once for device and
// host: adaptive models that turn a block's symbols into frequency triples, and the range coder that turns
// triples into bytes (and back).  Both are single dependency chains: on the GPU the model runs as one thread
// per block (blocks in parallel), the coder as one thread per file.  The same code compiles with gcc
// (tests/host/bwtc_host.c) so that the CPU test-suite checks it against the oracle without a GPU.
//
// Reference: lib/RangeCoder.js:27-232, lib/FenwickModel.js:15-165, lib/DefSumModel.js:11-131,
// lib/LogDistanceModel.js:8-49 over lib/NoModel.js:8-30, lib/BWTC.js:12-231, lib/Util.js:105-220.
#pragma once
#include <stdint.h>
#ifdef __CUDACC__
#define BC_FN __host__ __device__ inline
#else
#define BC_FN static inline
#endif

typedef uint8_t bc_u8;
typedef uint16_t bc_u16;
typedef uint32_t bc_u32;
typedef uint64_t bc_u64;

BC_FN int bc_fls(bc_u32 v) { int r = 0; while (v) { r++; v >>= 1; } return r; }  // Util.js:301-316

// ---- frequency triples ---------------------------------------------------------------------------------------
// encodeFreq(sy_f, lt_f, tot_f) (RangeCoder.js:81-91); encodeShift(sy, lt, shift) is the same arithmetic with
// tot_f = 1 << shift (:92-102: r = range >> shift, "last symbol" iff lt + sy >= 2^shift), so one record serves both.
BC_FN bc_u64 bc_triple(bc_u32 sy, bc_u32 lt, bc_u32 tot) { return (bc_u64)sy | ((bc_u64)lt << 21) | ((bc_u64)tot << 42); }
typedef struct { bc_u64* t; bc_u32 n, cap; } bc_emit;
BC_FN void bc_put(bc_emit* e, bc_u32 sy, bc_u32 lt, bc_u32 tot) {
  if (e->n < e->cap) e->t[e->n] = bc_triple(sy, lt, tot);
  e->n++;  // n > cap afterwards = overflow, checked by the caller
}
BC_FN void bc_put_bit(bc_emit* e, bc_u32 b) { bc_put(e, 1, b ? 1u : 0u, 2); }                 // RangeCoder.js:104-106
BC_FN void bc_put_raw(bc_emit* e, int bits, bc_u32 symbol) {                                   // NoModel.js:15-21
  for (int i = bits - 1; i >= 0; i--) bc_put_bit(e, (symbol >> i) & 1);
}
// LogDistanceModel(blockSize, 0, NoModel, NoModel).encode (LogDistanceModel.js:26-39)
BC_FN int bc_lgbits(bc_u32 blockSize) { return bc_fls((bc_u32)bc_fls(blockSize - 1)); }
BC_FN void bc_put_distance(bc_emit* e, int lgbits, bc_u32 distance) {
  if (distance < 2) { bc_put_raw(e, lgbits, distance); return; }
  const int lg = bc_fls(distance);
  bc_put_raw(e, lgbits, (bc_u32)lg);
  bc_put_raw(e, lg - 1, distance & ((1u << (lg - 1)) - 1));
}

// ---- FenwickModel (levels 6..9) ----------------------------------------------------------------------------------
#define BC_ESC_MASK 0x0000FFFFu
#define BC_SYM_MASK 0xFFFF0000u
#define BC_SCALE_MASK 0xFFFEFFFEu
#define BC_F_PROB_MAX 0xFF00u   // BWTC.js:7
#define BC_F_PROB_INCR 0x0100u  // BWTC.js:8
typedef struct { bc_u32 numSyms; bc_u32 tree[2 * 260]; } bc_fen;

BC_FN void bc_fen_sum(bc_fen* f) {  // FenwickModel.js:155-161
  for (bc_u32 i = f->numSyms - 1; i > 0; i--) f->tree[i] = f->tree[2 * i] + f->tree[2 * i + 1];
}
BC_FN void bc_fen_init(bc_fen* f, bc_u32 size) {  // :15-33
  f->numSyms = size + 1;
  for (bc_u32 i = 0; i < 2 * 260; i++) f->tree[i] = 0;
  bc_u32 i;
  for (i = 0; i < size; i++) f->tree[f->numSyms + i] = 1;       // escape probability 1, symbol probability 0
  f->tree[f->numSyms + i] = BC_F_PROB_INCR << 16;                // the escape symbol
  bc_fen_sum(f);
}
BC_FN void bc_fen_rescale(bc_fen* f) {  // :125-154
  bc_u32 i, prob; int noEscape = 1;
  for (i = 0; i < f->numSyms - 1; i++) {
    prob = f->tree[f->numSyms + i];
    if (prob & BC_ESC_MASK) { noEscape = 0; continue; }
    prob = (prob & BC_SCALE_MASK) >> 1;
    if (prob == 0) { prob = 1; noEscape = 0; }
    f->tree[f->numSyms + i] = prob;
  }
  prob = (f->tree[f->numSyms + i] & BC_SCALE_MASK) >> 1;
  if (noEscape) prob = 0; else if (prob == 0) prob = 1u << 16;
  f->tree[f->numSyms + i] = prob;
  bc_fen_sum(f);
}
// one trip up the tree (:60-86).  sy_leaf = the leaf value read BEFORE any escape was coded (:49); esc = code the
// symbol in the escape (not-yet-seen) distribution
BC_FN void bc_fen_step(bc_fen* f, bc_emit* e, bc_u32 symbol, bc_u32 sy_leaf, int esc) {
  bc_u32 i = f->numSyms + symbol;
  bc_u32 mask = BC_SYM_MASK, shift = 16, update = BC_F_PROB_INCR << 16;
  if (esc) { mask = BC_ESC_MASK; update -= 1; shift = 0; }
  else if (symbol == f->numSyms - 1 && (f->tree[1] & BC_ESC_MASK) == 1) update = 0u - f->tree[i];  // the last escape
  bc_u32 lt_f = 0;
  while (i > 1) {
    const bc_u32 parent = i >> 1;
    if (i & 1) lt_f += f->tree[2 * parent];
    f->tree[i] += update;
    i = parent;
  }
  const bc_u32 tot_f = f->tree[1];
  f->tree[1] += update;
  bc_put(e, (sy_leaf & mask) >> shift, (lt_f & mask) >> shift, (tot_f & mask) >> shift);
  if (((f->tree[1] & BC_SYM_MASK) >> 16) >= BC_F_PROB_MAX) bc_fen_rescale(f);
}
BC_FN void bc_fen_encode(bc_fen* f, bc_emit* e, bc_u32 symbol) {  // :47-87 without the recursion
  const bc_u32 sy_leaf = f->tree[f->numSyms + symbol];
  if ((sy_leaf & BC_SYM_MASK) == 0) {
    const bc_u32 escSym = f->numSyms - 1;
    bc_fen_step(f, e, escSym, f->tree[f->numSyms + escSym], 0);
    bc_fen_step(f, e, symbol, sy_leaf, 1);
  } else {
    bc_fen_step(f, e, symbol, sy_leaf, 0);
  }
}

// ---- DefSumModel (levels 1..5) -----------------------------------------------------------------------------------
#define BC_DS_TOTAL 256u
typedef struct { bc_u32 numSyms, updateCount, updateThresh; bc_u16 prob[262], escape[262], update[262]; } bc_dsm;
BC_FN void bc_dsm_init(bc_dsm* m, bc_u32 size) {  // DefSumModel.js:11-35
  for (bc_u32 i = 0; i < 262; i++) { m->prob[i] = 0; m->escape[i] = 0; m->update[i] = 0; }
  m->numSyms = size;
  m->prob[size + 1] = BC_DS_TOTAL;
  for (bc_u32 i = 0; i <= size; i++) m->escape[i] = (bc_u16)i;
  m->updateCount = 0;
  m->updateThresh = BC_DS_TOTAL - BC_DS_TOTAL / 2;
}
BC_FN void bc_dsm_update(bc_dsm* m, bc_u32 symbol) {  // :39-93
  if (symbol == m->numSyms) {
    if (m->update[symbol] >= 40) return;
    if (m->updateCount >= m->updateThresh - 1) return;
  }
  m->update[symbol]++;
  m->updateCount++;
  if (m->updateCount < m->updateThresh) return;
  bc_u32 cumProb = 0, cumEscProb = 0, odd = 0, i;
  m->escape[0] = 0; m->prob[0] = 0;
  for (i = 0; i < m->numSyms + 1; i++) {
    const bc_u32 newProb = ((bc_u32)(m->prob[i + 1] - m->prob[i]) >> 1) + m->update[i];
    m->prob[i] = (bc_u16)cumProb;
    m->escape[i] = (bc_u16)cumEscProb;
    if (newProb) { cumProb += newProb; if (newProb & 1) odd++; }
    else cumEscProb++;
  }
  m->prob[i] = (bc_u16)cumProb;
  m->updateThresh = BC_DS_TOTAL - (cumProb - odd) / 2;
  for (i = 0; i < m->numSyms + 1; i++) m->update[i] = 0;
  m->update[m->numSyms] = 1;
  m->updateCount = 1;
}
BC_FN void bc_dsm_encode(bc_dsm* m, bc_emit* e, bc_u32 symbol) {  // :94-111 without the recursion
  bc_u32 lt_f = m->prob[symbol], sy_f = m->prob[symbol + 1] - lt_f;
  if (sy_f) { bc_put(e, sy_f, lt_f, BC_DS_TOTAL); bc_dsm_update(m, symbol); return; }
  const bc_u32 esc = m->numSyms;
  bc_put(e, m->prob[esc + 1] - m->prob[esc], m->prob[esc], BC_DS_TOTAL);
  bc_dsm_update(m, esc);
  lt_f = m->escape[symbol];
  sy_f = m->escape[symbol + 1] - lt_f;
  bc_put(e, sy_f, lt_f, m->escape[m->numSyms]);
  bc_dsm_update(m, symbol);
}

// ---- one block -> triples (lib/BWTC.js:44-137) ------------------------------------------------------------------
// used: 256-bit map of the bytes present; sym: the block's MTF/zero-run symbols exactly as the bzip2 stage
// produces them (0 = RUNA, 1 = RUNB, rank+1; lib/BWTC.js:112-137 is the same coding without an end-of-block
// symbol), nsym of them.  scratch holds the model.
typedef union { bc_fen fen; bc_dsm dsm; } bc_model;
// block header (lib/BWTC.js:50-83): size flag / length, primary index, the tree of used bytes; returns the alphabet size
BC_FN bc_u32 bc_block_header(bc_emit* e, bc_u32 blockSize, bc_u32 length, bc_u32 pidx1, const bc_u32* used) {
  const int lgbits = bc_lgbits(blockSize);
  if (length == blockSize) bc_put(e, 1, 0, 3);                                    // :50-52 "full size block"
  else { bc_put(e, 1, 1, 3); bc_put_distance(e, lgbits, length); }                // :54-55
  bc_put_distance(e, lgbits, pidx1);                                              // :59
  bc_u16 useTree[512];                                                            // :61-83
  for (int i = 0; i < 256; i++) useTree[256 + i] = (bc_u16)((used[i >> 5] >> (i & 31)) & 1u);
  for (int i = 255; i > 0; i--) useTree[i] = (bc_u16)(useTree[2 * i] + useTree[2 * i + 1]);
  useTree[0] = 1;
  for (bc_u32 i = 1; i < 512; i++) {
    const bc_u32 parent = i >> 1, full = 1u << (9 - bc_fls(i));
    if (useTree[parent] == 0 || useTree[parent] == full * 2) continue;
    if (i >= 256) bc_put_bit(e, useTree[i]);
    else bc_put(e, 1, useTree[i] == 0 ? 0u : (useTree[i] == full ? 2u : 1u), 3);
  }
  return useTree[1];
}
BC_FN void bc_block_triples(bc_emit* e, bc_model* scratch, bc_u32 blockSize, bc_u32 length, bc_u32 pidx1, const bc_u32* used,
                            const bc_u16* sym, bc_u32 nsym, int fast) {
  const bc_u32 alphabetSize = bc_block_header(e, blockSize, length, pidx1, used);
  if (fast) {                                                                     // :109-111
    bc_dsm_init(&scratch->dsm, alphabetSize + 1);
    for (bc_u32 k = 0; k < nsym; k++) bc_dsm_encode(&scratch->dsm, e, sym[k]);
  } else {
    bc_fen_init(&scratch->fen, alphabetSize + 1);
    for (bc_u32 k = 0; k < nsym; k++) bc_fen_encode(&scratch->fen, e, sym[k]);
  }
}

// ---- range coder, encode side ---------------------------------------------------------------------------------------
#define BC_TOP 0x80000000u
#define BC_SHIFT_BITS 23
#define BC_EXTRA_BITS 7
#define BC_BOTTOM (BC_TOP >> 8)
typedef struct { bc_u32 low, range, buffer, bytecount; bc_u64 help; bc_u8* out; bc_u64 n, cap; } bc_enc;
BC_FN void bc_out(bc_enc* rc, bc_u32 b) { if (rc->n < rc->cap) rc->out[rc->n] = (bc_u8)b; rc->n++; }
BC_FN void bc_enc_start(bc_enc* rc, bc_u8* out, bc_u64 cap, bc_u32 c) {  // RangeCoder.js:66-72, initlength 1
  rc->low = 0; rc->range = BC_TOP; rc->buffer = c; rc->help = 0; rc->bytecount = 1; rc->out = out; rc->n = 0; rc->cap = cap;
}
BC_FN void bc_enc_normalize(bc_enc* rc) {  // :40-61
  while (rc->range <= BC_BOTTOM) {
    if (rc->low < (0xFFu << BC_SHIFT_BITS)) {
      bc_out(rc, rc->buffer);
      for (; rc->help; rc->help--) bc_out(rc, 0xFF);
      rc->buffer = (rc->low >> BC_SHIFT_BITS) & 0xFF;
    } else if (rc->low & BC_TOP) {
      bc_out(rc, rc->buffer + 1);
      for (; rc->help; rc->help--) bc_out(rc, 0x00);
      rc->buffer = (rc->low >> BC_SHIFT_BITS) & 0xFF;
    } else {
      rc->help++;
    }
    rc->range <<= 8;
    rc->low = (rc->low << 8) & (BC_TOP - 1);
    rc->bytecount++;
  }
}
BC_FN void bc_enc_code(bc_enc* rc, bc_u64 triple) {  // :81-91
  const bc_u32 sy_f = (bc_u32)(triple & 0x1FFFFF), lt_f = (bc_u32)((triple >> 21) & 0x1FFFFF), tot_f = (bc_u32)(triple >> 42);
  bc_enc_normalize(rc);
  const bc_u32 r = rc->range / tot_f;
  const bc_u32 tmp = r * lt_f;
  rc->low += tmp;
  if (lt_f + sy_f < tot_f) rc->range = r * sy_f; else rc->range -= tmp;
}
BC_FN void bc_enc_finish(bc_enc* rc) {  // :118-144
  bc_enc_normalize(rc);
  rc->bytecount += 5;
  bc_u32 tmp = rc->low >> BC_SHIFT_BITS;
  if ((rc->low & (BC_BOTTOM - 1)) >= ((rc->bytecount & 0xFFFFFF) >> 1)) tmp++;
  if (tmp > 0xFF) { bc_out(rc, rc->buffer + 1); for (; rc->help; rc->help--) bc_out(rc, 0x00); }
  else { bc_out(rc, rc->buffer); for (; rc->help; rc->help--) bc_out(rc, 0xFF); }
  bc_out(rc, tmp & 0xFF);
  bc_out(rc, (rc->bytecount >> 16) & 0xFF);
  bc_out(rc, (rc->bytecount >> 8) & 0xFF);
  bc_out(rc, rc->bytecount & 0xFF);
}
// file header (Util.js:105-141 with suppressFinalByte): "bwtc", the size + 1 in big-endian 7-bit groups; the last
// group (flagged 0x80) is not written but becomes the coder's first buffered byte.  Returns the bytes written.
BC_FN bc_u32 bc_file_header(bc_u8* out, bc_u64 fileSize, bc_u32* finalByte) {
  bc_u8 grp[12]; int ng = 0;
  bc_u64 v = fileSize + 1;
  do { grp[ng++] = (bc_u8)(v & 0x7F); v >>= 7; } while (v);
  grp[0] |= 0x80;
  bc_u32 k = 0;
  out[k++] = 'b'; out[k++] = 'w'; out[k++] = 't'; out[k++] = 'c';
  for (int i = ng - 1; i >= 1; i--) out[k++] = grp[i];
  *finalByte = grp[0];
  return k;
}

// ---- decode side (model and coder cannot be separated: every cumulative frequency depends on the model state) ----
typedef struct { bc_u32 low, range, buffer, help; const bc_u8* in; bc_u64 n, pos; } bc_dec;
BC_FN bc_u32 bc_get(bc_dec* rc) { return rc->pos < rc->n ? (bc_u32)rc->in[rc->pos++] : 0xFFFFFFFFu; }  // EOF = -1
BC_FN void bc_dec_start(bc_dec* rc, const bc_u8* in, bc_u64 n, bc_u64 pos) {  // :150-159, first byte already consumed
  rc->in = in; rc->n = n; rc->pos = pos; rc->help = 0;
  rc->buffer = bc_get(rc);
  rc->low = rc->buffer >> (8 - BC_EXTRA_BITS);
  rc->range = 1u << BC_EXTRA_BITS;
}
BC_FN void bc_dec_normalize(bc_dec* rc) {  // :161-170
  while (rc->range <= BC_BOTTOM) {
    rc->low = (rc->low << 8) | ((rc->buffer << BC_EXTRA_BITS) & 0xFF);
    rc->buffer = bc_get(rc);
    rc->low |= rc->buffer >> (8 - BC_EXTRA_BITS);
    rc->range <<= 8;
  }
}
BC_FN bc_u32 bc_dec_cul(bc_dec* rc, bc_u32 tot_f) {  // :177-189 (decodeCulShift is the same with tot = 1 << shift)
  bc_dec_normalize(rc);
  rc->help = rc->range / tot_f;
  const bc_u32 tmp = rc->low / rc->help;
  return tmp >= tot_f ? tot_f - 1 : tmp;
}
BC_FN void bc_dec_update(bc_dec* rc, bc_u32 sy_f, bc_u32 lt_f, bc_u32 tot_f) {  // :197-205
  const bc_u32 tmp = rc->help * lt_f;
  rc->low -= tmp;
  if (lt_f + sy_f < tot_f) rc->range = rc->help * sy_f; else rc->range -= tmp;
}
BC_FN bc_u32 bc_dec_bit(bc_dec* rc) { const bc_u32 t = bc_dec_cul(rc, 2); bc_dec_update(rc, 1, t, 2); return t; }
BC_FN bc_u32 bc_dec_raw(bc_dec* rc, int bits) { bc_u32 r = 0; for (int i = bits - 1; i >= 0; i--) r = (r << 1) | bc_dec_bit(rc); return r; }
BC_FN bc_u32 bc_dec_distance(bc_dec* rc, int lgbits) {  // LogDistanceModel.js:40-47
  const bc_u32 lg = bc_dec_raw(rc, lgbits);
  if (lg < 2) return lg;
  if (lg > 31) return 0xFFFFFFFFu;
  return (1u << (lg - 1)) + bc_dec_raw(rc, (int)lg - 1);
}
BC_FN bc_u32 bc_fen_decode1(bc_fen* f, bc_dec* rc, int esc) {  // FenwickModel.js:88-123
  bc_u32 mask = BC_SYM_MASK, shift = 16, update = BC_F_PROB_INCR << 16;
  if (esc) { mask = BC_ESC_MASK; update -= 1; shift = 0; }
  const bc_u32 tot_f = (f->tree[1] & mask) >> shift;
  if (tot_f == 0) return 0xFFFFFFFFu;
  const bc_u32 prob = bc_dec_cul(rc, tot_f);
  bc_u32 i = 1, lt_f = 0;
  while (i < f->numSyms) {
    f->tree[i] += update;
    const bc_u32 leftProb = (f->tree[2 * i] & mask) >> shift;
    i *= 2;
    if (prob - lt_f >= leftProb) { lt_f += leftProb; i++; }
  }
  const bc_u32 symbol = i - f->numSyms;
  const bc_u32 sy_f = (f->tree[i] & mask) >> shift;
  f->tree[i] += update;
  bc_dec_update(rc, sy_f, lt_f, tot_f);
  if (symbol == f->numSyms - 1 && (f->tree[1] & BC_ESC_MASK) == 1) {
    update = 0u - f->tree[i];
    while (i >= 1) { f->tree[i] += update; i >>= 1; }
  }
  if (((f->tree[1] & BC_SYM_MASK) >> 16) >= BC_F_PROB_MAX) bc_fen_rescale(f);
  return symbol;
}
BC_FN bc_u32 bc_fen_decode(bc_fen* f, bc_dec* rc) {  // :115-122
  bc_u32 s = bc_fen_decode1(f, rc, 0);
  if (s == f->numSyms - 1) s = bc_fen_decode1(f, rc, 1);
  return s;
}
BC_FN bc_u32 bc_dsm_decode(bc_dsm* m, bc_dec* rc) {  // DefSumModel.js:112-131 (tables searched, not cached)
  bc_u32 prob = bc_dec_cul(rc, BC_DS_TOTAL), symbol = 0;
  while (symbol < m->numSyms && !(m->prob[symbol] <= prob && prob < m->prob[symbol + 1])) symbol++;
  bc_u32 lt_f = m->prob[symbol], sy_f = m->prob[symbol + 1] - lt_f;
  bc_dec_update(rc, sy_f, lt_f, BC_DS_TOTAL);
  bc_dsm_update(m, symbol);
  if (symbol != m->numSyms) return symbol;
  const bc_u32 tot_f = m->escape[m->numSyms];
  if (tot_f == 0) return 0xFFFFFFFFu;
  prob = bc_dec_cul(rc, tot_f);
  symbol = 0;
  while (symbol + 1 < m->numSyms && !(m->escape[symbol] <= prob && prob < m->escape[symbol + 1])) symbol++;
  lt_f = m->escape[symbol];
  sy_f = m->escape[symbol + 1] - lt_f;
  bc_dec_update(rc, sy_f, lt_f, tot_f);
  bc_dsm_update(m, symbol);
  return symbol;
}
// One block (lib/BWTC.js:156-222): returns 0 = block decoded (L column of *length bytes in `L`, *pidx1 = start index
// for unbwtransform), 1 = "no more blocks", negative = the stream is nonsense (the reference has no checks there).
BC_FN int bc_decode_block(bc_dec* rc, bc_model* scratch, bc_u32 blockSize, int fast, bc_u8* L, bc_u32* length, bc_u32* pidx1) {
  const int lgbits = bc_lgbits(blockSize);
  const bc_u32 ind = bc_dec_cul(rc, 3);                                           // :158-159
  bc_dec_update(rc, 1, ind, 3);
  bc_u32 len;
  if (ind == 0) len = blockSize;
  else if (ind == 1) len = bc_dec_distance(rc, lgbits);
  else return 1;
  if (len == 0 || len > blockSize) return -5;
  const bc_u32 p1 = bc_dec_distance(rc, lgbits);                                  // :170
  if (p1 > len) return -5;
  bc_u16 useTree[512];                                                            // :172-187
  useTree[0] = 1;
  for (bc_u32 i = 1; i < 512; i++) {
    const bc_u32 parent = i >> 1, full = 1u << (9 - bc_fls(i));
    if (useTree[parent] == 0 || useTree[parent] == full * 2) useTree[i] = (bc_u16)(useTree[parent] >> 1);
    else if (i >= 256) useTree[i] = (bc_u16)bc_dec_bit(rc);
    else { const bc_u32 v = bc_dec_cul(rc, 3); bc_dec_update(rc, 1, v, 3); useTree[i] = (bc_u16)(v == 2 ? full : v); }
  }
  bc_u8 M[256];
  bc_u32 alphabetSize = 0;
  for (bc_u32 i = 0; i < 256; i++) if (useTree[256 + i]) M[alphabetSize++] = (bc_u8)i;
  if (alphabetSize == 0) return -5;
  if (fast) bc_dsm_init(&scratch->dsm, alphabetSize + 1); else bc_fen_init(&scratch->fen, alphabetSize + 1);
  bc_u64 val = 1;
  bc_u32 i = 0;
  while (i < len) {                                                               // :200-222, inverse MTF folded in
    const bc_u32 c = fast ? bc_dsm_decode(&scratch->dsm, rc) : bc_fen_decode(&scratch->fen, rc);
    if (c == 0xFFFFFFFFu || c > alphabetSize) return -5;
    if (c <= 1) {
      const bc_u64 cnt = val * (c + 1);
      if (cnt > len - i) return -5;
      for (bc_u64 k = 0; k < cnt; k++) L[i++] = M[0];                              // rank 0 = the front of the list
      val *= 2;
    } else {
      val = 1;
      bc_u32 j = c - 1;
      const bc_u8 b = M[j];
      for (; j > 0; j--) M[j] = M[j - 1];
      M[0] = b;
      L[i++] = b;
    }
  }
  *length = len; *pidx1 = p1;
  return 0;
}
