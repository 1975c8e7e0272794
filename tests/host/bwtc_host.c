/* Host build of compressjs_b200/csrc/bwtc_core.cuh for the CPU test-suite (no GPU needed): the serial model and
 * range-coder code that the CUDA kernels of bwtc.cu run, driven here with block data prepared by the test
 * (sentinel BWT from the oracle, MTF + zero-run symbols computed below the way the bzip2 GPU stage emits them). */
#include <stdlib.h>
#include <string.h>
#include "../../compressjs_b200/csrc/bwtc_core.cuh"
#define EXPORT __attribute__((visibility("default")))

/* U = concatenated sentinel-BWT outputs of the blocks; returns the stream size (or 0 on overflow) */
EXPORT size_t host_bwtc_encode(int level, uint32_t nblocks, const uint32_t* lengths, const uint32_t* pidx1, const uint8_t* U,
                               uint64_t file_size, uint8_t* out, size_t cap) {
  const uint32_t blockSize = (uint32_t)level * 100000u;
  const int fast = level <= 5;
  uint32_t finalByte;
  const uint32_t hdr = bc_file_header(out, file_size, &finalByte);
  bc_enc rc;
  bc_enc_start(&rc, out + hdr, cap - hdr, finalByte);
  bc_enc_code(&rc, bc_triple(1, (uint32_t)level, 256));                       /* encoder.encodeByte(blockSize), BWTC.js:21 */
  const uint32_t tcap = 2 * (blockSize + 1) + 1024;
  uint64_t* tr = (uint64_t*)malloc((size_t)tcap * 8);
  uint16_t* sym = (uint16_t*)malloc((size_t)(blockSize + 2) * 2);
  bc_model* model = (bc_model*)malloc(sizeof(bc_model));
  size_t off = 0;
  int bad = 0;
  for (uint32_t b = 0; b < nblocks; b++) {
    const uint8_t* u = U + off;
    const uint32_t len = lengths[b];
    off += len;
    uint32_t used[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (uint32_t i = 0; i < len; i++) used[u[i] >> 5] |= 1u << (u[i] & 31);
    uint8_t M[256]; uint32_t a = 0;
    for (uint32_t i = 0; i < 256; i++) if ((used[i >> 5] >> (i & 31)) & 1) M[a++] = (uint8_t)i;
    uint32_t ns = 0, run = 0;
    for (uint32_t i = 0; i <= len; i++) {
      uint32_t j = 0;
      if (i < len) {
        while (M[j] != u[i]) j++;
        const uint8_t c = M[j];
        for (uint32_t k = j; k > 0; k--) M[k] = M[k - 1];
        M[0] = c;
        if (j == 0) { run++; continue; }
      }
      while (run) { if (run & 1) { sym[ns++] = 0; run -= 1; } else { sym[ns++] = 1; run -= 2; } run >>= 1; }
      if (i < len) sym[ns++] = (uint16_t)(j + 1);
    }
    bc_emit e = {tr, 0, tcap};
    bc_block_triples(&e, model, blockSize, len, pidx1[b], used, sym, ns, fast);
    if (e.n > e.cap) { bad = 1; break; }
    for (uint32_t k = 0; k < e.n; k++) bc_enc_code(&rc, tr[k]);
  }
  bc_enc_code(&rc, bc_triple(1, 2, 3));                                           /* "no more blocks", BWTC.js:141 */
  bc_enc_finish(&rc);
  free(tr); free(sym); free(model);
  if (bad || rc.n > rc.cap) return 0;
  return hdr + (size_t)rc.n;
}

/* Decodes the container down to the L columns (inverse MTF done): Lout = concatenated blocks; returns the number of
 * blocks or a negative code */
EXPORT int host_bwtc_decode(const uint8_t* in, size_t n, uint8_t* Lout, size_t Lcap, uint32_t* lengths, uint32_t* pidx1, uint32_t maxblocks,
                            uint64_t* file_size_plus1) {
  if (n < 5 || memcmp(in, "bwtc", 4)) return -2;
  size_t pos = 4; uint64_t fs = 0;
  for (;;) {
    if (pos >= n) return -5;
    const uint32_t c = in[pos++];
    if (c & 0x80) { fs += c & 0x7F; break; }
    fs = (fs + c) * 128;
  }
  *file_size_plus1 = fs;
  bc_dec rc;
  bc_dec_start(&rc, in, n, pos);
  const uint32_t level = bc_dec_cul(&rc, 256);
  bc_dec_update(&rc, 1, level, 256);
  if (level < 1 || level > 9) return -5;
  bc_model* model = (bc_model*)malloc(sizeof(bc_model));
  uint32_t nb = 0; size_t off = 0; int r = 0;
  for (;;) {
    if ((size_t)level * 100000u > Lcap - off && 0) { r = -6; break; }
    uint8_t* L = (uint8_t*)malloc((size_t)level * 100000u);
    uint32_t len = 0, p1 = 0;
    r = bc_decode_block(&rc, model, level * 100000u, level <= 5, L, &len, &p1);
    if (r == 0) {
      if (nb >= maxblocks || off + len > Lcap) { free(L); r = -6; break; }
      memcpy(Lout + off, L, len); off += len; lengths[nb] = len; pidx1[nb] = p1; nb++;
    }
    free(L);
    if (r) break;
  }
  free(model);
  return r == 1 ? (int)nb : r;
}
