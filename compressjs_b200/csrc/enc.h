// enc.h -- shared declarations of the encode stages.
#pragma once
#include "ctx.h"

// One bzip2 block as cut by the RLE1 stage (lib/Bzip2.js:636-667 readBlock).
struct BlkInfo {
  u64 s;    // first raw byte
  u64 e;    // one past the last raw byte consumed
  u64 b;    // end of the (re-phased) run the block starts in; == s when the block starts on a run start
  u64 Wb;   // RLE1 output bytes produced by raw[0,b) under maximal-run phases
  u32 ofs;  // RLE1 bytes produced by raw[s,b) with a fresh run state at s
  u32 n;    // post-RLE1 length of the block (<= blockSize)
};

struct Rle1Plan {
  DBuf<u32> tile_carry;   // per raw tile: length (mod 255) of the run entering the tile
  DBuf<u64> tile_prefix;  // per raw tile: W(tile start)
  DBuf<u8> tile_plain;    // per raw tile: 1 = no four equal bytes in a row in or into the tile (RLE1 copies it byte for byte)
  DBuf<BlkInfo> blocks;   // device block table
  std::vector<BlkInfo> h_blocks;
  size_t nblocks = 0;      // entries of h_blocks
  size_t first_index = 0;  // global block index of h_blocks[0] (range plans)
  size_t total_guess = 0;  // ceil(W(N) / blockSize)
  u64 w_total = 0;         // W at the end of the buffer (RLE1 bytes of the whole input up to there)
  u64 ntiles = 0;
};

#define RLE_TILE 4096

void rle1_plan(Ctx& c, const u8* d_in, size_t n, int level, Rle1Plan& plan);
// st0 / W0: run state and RLE1 output in front of the buffer when it is a share of a larger input (0, 0 for a whole file);
// agg_state (host, 2 x u64, optional): receives the aggregate run state of the buffer and the length of its leading run.
void rle1_plan_ex(Ctx& c, const u8* d_in, size_t n, int level, Rle1Plan& plan, long long spec_first, size_t spec_count, bool tiles_only,
                  u64 st0 = 0, u64 W0 = 0, u64* agg_state = nullptr);
// materialise blocks [first, first+count) of the plan into the slot layout at d_T (u8[count<<20]);
// d_n receives their lengths, d_crc their CRCs.
void rle1_materialize(Ctx& c, const u8* d_in, size_t n, const Rle1Plan& plan, size_t first, size_t count, u8* d_T, u32* d_n, u32* d_crc);

// MTF + RLE2 (lib/Bzip2.js:743-815): U (slot layout) -> symbols u16 (slot layout), m, freq, used map
void mtf_rle2_batch(Ctx& c, const u8* d_T, const u8* d_U, const u32* d_n, const u32* h_n, u32 nblk, u16* d_sym, u32* d_m, u32* d_freq /*[nblk][258]*/,
                    u32* d_used /*[nblk][8]*/, const u32* d_bytehist = nullptr /*[nblk][256] byte histograms of the blocks, if known*/);

#define HUFF_MAXSYM 258
#define HUFF_MAXGROUPS 6
#define HUFF_GROUP 50
// per-block result of the Huffman stage
struct HuffBlk {
  u32 ngroups, nsel, alpha, m;
  u64 body_bits;  // bits of the block from the 48-bit magic through the last Huffman code
  u8 len[HUFF_MAXGROUPS][HUFF_MAXSYM + 6];
};
// Table optimisation (lib/Bzip2.js:671-733, 826-843 + HuffmanAllocator.js): selectors u8 (slot layout >> 0, stride SEL_STRIDE)
#define SEL_STRIDE 18432
void huffman_batch(Ctx& c, const u16* d_sym, const u32* d_m, const u32* d_freq, const u32* d_used, u32 nblk, u8* d_sel, u8* d_selmtf,
                   HuffBlk* d_hb);
// bit packing of blocks at their final bit offsets (lib/Bzip2.js:740-741,749-758,847-874)
void pack_batch(Ctx& c, const u16* d_sym, const u8* d_sel, const u8* d_selmtf, const HuffBlk* d_hb, const u32* d_used, const u32* d_pidx,
                const u32* d_crc, const u64* d_bitoff, const u32* d_flag, u32 nblk, u32 max_m, u32* d_out_words);

#include <vector>
void crc_ranges(Ctx& c, const u8* d_data, const BlkInfo* d_ranges, const std::vector<BlkInfo>& h_ranges, u32* d_crc_out);
