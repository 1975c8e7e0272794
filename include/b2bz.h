/*
 * b2bz.h -- C ABI of libb2bz.so, the B200-native bzip2 / BWT block pipeline.
 *
 * This is the drop-in boundary for compressjs' bzip2 hot path.  Every entry point
 * below replaces one JavaScript function of the reference (file:line under
 * /root/reference); a Node N-API addon (compressjs_b200/napi/addon.cc), or any other
 * FFI, binds exactly these symbols.  Plain pointers and sizes only.
 *
 * Conventions
 *   - return value 0 = OK; negative = the reference's Bunzip.Err code
 *     (lib/Bzip2.js:62-72: -2 NOT_BZIP_DATA, -5 DATA_ERROR, -7 OBSOLETE_INPUT) or
 *     B2_ERR_* below.  b2_last_error() returns the reference's message text.
 *   - inputs are borrowed for the duration of the call; outputs are allocated by the
 *     library in pinned host memory and released with b2_free().
 *   - all work runs on the GPU selected by b2_init(); there is NO CPU fallback: if no
 *     CUDA device is usable every call fails with B2_ERR_CUDA.
 *   - calls are synchronous and serialised by an internal mutex.
 */
#ifndef B2BZ_H
#define B2BZ_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2_OK 0
#define B2_ERR_NOT_BZIP_DATA (-2) /* lib/Bzip2.js:66 */
#define B2_ERR_DATA_ERROR (-5)    /* lib/Bzip2.js:69 */
#define B2_ERR_OBSOLETE_INPUT (-7) /* lib/Bzip2.js:71 */
#define B2_ERR_BAD_LEVEL (-100)   /* lib/Bzip2.js:888-890 "Invalid block size multiplier" */
#define B2_ERR_BAD_ARG (-101)
#define B2_ERR_BAD_MAGIC (-102)   /* lib/Util.js:151-153 Error("Bad magic") of the BWTC container */
#define B2_ERR_CUDA (-200)        /* CUDA runtime failure or no device: never falls back to CPU */

/* Select the CUDA device (ordinal) used by this process and create the context.
 * Called implicitly with device 0 (or $B2_DEVICE / $LOCAL_RANK) by the first call. */
int b2_init(int device);
void b2_shutdown(void);
const char* b2_last_error(void);
void b2_free(void* p);

/* ---- compressjs.Bzip2 (lib/Bzip2.js) -------------------------------------------- */
/* Bzip2.compressFile(input, output, level)            lib/Bzip2.js:879-929 */
int b2_bzip2_compress(const uint8_t* in, size_t n, int level, uint8_t** out, size_t* out_n);
/* Bzip2.decompressFile(input, output, multistream)    lib/Bzip2.js:454-481
 * Members of a multistream file may have different levels.  On an error nothing is returned (the reference hands the
 * blocks in front of the error to its output stream first); one call keeps ~2 MiB per block of the file on the device. */
int b2_bzip2_decompress(const uint8_t* in, size_t n, int multistream, uint8_t** out, size_t* out_n);
/* Bzip2.decompressBlock(input, bitPos, output)        lib/Bzip2.js:482-503 */
int b2_bzip2_decompress_block(const uint8_t* in, size_t n, uint64_t bitpos, uint8_t** out, size_t* out_n);
/* Bzip2.table(input, callback, multistream)           lib/Bzip2.js:508-548
 * (the callback is replayed by the host shim from the two arrays) */
int b2_bzip2_table(const uint8_t* in, size_t n, int multistream, uint64_t** bitpos, uint32_t** sizes, size_t* count);

/* ---- compressjs.BWT (lib/BWT.js) ------------------------------------------------- */
/* BWT.bwtransform2(T, U, n, 256) -> pidx  (cyclic)    lib/BWT.js:372-417
 * n is limited to 900 000 (the largest bzip2 block; the reference has no limit): longer strings return B2_ERR_BAD_ARG. */
int32_t b2_bwt_cyclic(const uint8_t* T, uint8_t* U, int32_t n);
/* many independent blocks at once (what compressFile does per block, batched):
 * block k is T + offs[k], length lens[k] (each <= 900000); U gets the same layout;
 * pidx[k] receives each block's primary index. */
int b2_bwt_cyclic_batch(const uint8_t* T, uint8_t* U, const uint64_t* offs, const int32_t* lens, int32_t* pidx, size_t nblocks);
/* The sentinel ("string is terminated by EOF") family used by the reference's BWTC container.  One string of
 * at most 2^20 - 2 = 1 048 574 bytes per call (the slot size of the block pipeline; BWTC blocks are <= 900 000).
 * BWT.suffixsort(T, SA, n)                            lib/BWT.js:305-321 ; returns 0 */
int b2_suffixsort(const uint8_t* T, int32_t* SA, int32_t n);
/* BWT.bwtransform(T, U, A, n) -> pidx + 1             lib/BWT.js:328-350 (A is scratch in the reference) */
int32_t b2_bwt_sentinel(const uint8_t* T, uint8_t* U, int32_t n);
/* BWT.unbwtransform(T, U, LF, n, pidx)                lib/BWT.js:352-363 (L = the transformed string, pidx as
 * returned by bwtransform; LF is scratch in the reference) */
int b2_bwt_inverse(const uint8_t* L, uint8_t* out, int32_t n, int32_t pidx);
/* ---- compressjs.BWTC (lib/BWTC.js) -- new, bound by one serial coder thread: see compressjs_b200/csrc/bwtc.cu -----
 * BWTC.compressFile(input, output, level)             lib/BWTC.js:12-139 (level outside 1..9 means 9, as there) */
int b2_bwtc_compress(const uint8_t* in, size_t n, int level, uint8_t** out, size_t* out_n);
/* BWTC.decompressFile(input, output)                  lib/BWTC.js:141-231 (streams that carry their size) */
int b2_bwtc_decompress(const uint8_t* in, size_t n, uint8_t** out, size_t* out_n);
/* CRC32 helper object of lib/CRC32.js:72-103 (bzip2 polynomial, MSB first) */
uint32_t b2_crc32_bzip2(const uint8_t* p, size_t n);

/* ---- device-resident entry points (buffers already in HBM) ----------------------- */
/* Same semantics as b2_bzip2_compress, but `d_in` / `d_out` are device pointers on the
 * b2_init() device (e.g. torch tensors' data_ptr()).  out_cap must be >= b2_bzip2_bound(n).
 * Used by bench.py for the HBM-resident `value` and by the multi-GPU host layer. */
size_t b2_bzip2_bound(size_t n);
int b2_bzip2_compress_dev(const void* d_in, size_t n, int level, void* d_out, size_t out_cap, size_t* out_n);
/* Decode with device buffers; *out_n receives the decoded size; fails with
 * B2_ERR_BAD_ARG (and the needed size in *out_n) if out_cap is too small. */
int b2_bzip2_decompress_dev(const void* d_in, size_t n, int multistream, void* d_out, size_t out_cap, size_t* out_n);

/* ---- block-range encode for multi-GPU sharding ------------- */
/* Encodes blocks [first, first+count) of the stream that b2_bzip2_compress would produce
 * for (d_in, n, level) WITHOUT file header/trailer: the fragment starts at bit offset
 * (*bit_phase in 0..7, chosen by the caller = global bit offset mod 8) inside d_out and is
 * *out_bits long.  block_crcs (host, cap entries) receives the per-block CRCs so the
 * caller can fold the stream CRC.  total_blocks receives the number of blocks in the file. */
int b2_bzip2_plan(const void* d_in, size_t n, int level, size_t* total_blocks);
/* Speculative range plan for rank `rank` of `world`: cuts only this rank's share of the blocks, starting from
 * the boundary implied by W-space arithmetic (exact unless a run-phase slip happened earlier in the file).
 * info[0..5] = raw start, raw end, first block, planned count, blocks actually cut, total block guess.
 * The ranks must verify end(r) == start(r+1), start(0) == 0, end(last) == n and cut == planned on every rank
 * (compressjs_b200/sharded.py does); otherwise fall back to b2_bzip2_plan.  The plan is cached for the next
 * b2_bzip2_encode_range_dev on the same buffer. */
int b2_bzip2_plan_spec(const void* d_in, size_t n, int level, int rank, int world, uint64_t* info);
/* Sharded input: every rank holds only a contiguous share of the input (followed by a halo: the first bytes of the
 * next share, so that a block that starts in the share can be finished).
 * summary[0..3] = {aggregate RLE1 run state of the share (packed, opaque), length of its leading run, RLE1 bytes of the
 * share when no run enters it, share length}; the host layer combines the summaries of all ranks in order
 * (compressjs_b200/sharded.py: share_plan_inputs) into the run state / RLE1 output in front of every share. */
int b2_bzip2_share_summary(const void* d_share, size_t n, uint64_t* summary);
/* Cuts blocks [first, first+count) of the whole input inside the buffer d_buf[0, n) = share + halo, given the run state
 * and RLE1 output in front of it.  Speculative like b2_bzip2_plan_spec (same checks by the caller); info[0..5] = raw
 * start, raw end (offsets inside d_buf), first, planned, blocks cut, RLE1 output up to the end of the buffer.  The plan
 * is cached for the next b2_bzip2_encode_range_dev(d_buf, n, level, first, count, ...). */
int b2_bzip2_plan_share(const void* d_buf, size_t n, int level, uint64_t state_in, uint64_t w_in, size_t first, size_t count, uint64_t* info);
/* dst := the first nbits of src moved to start at bit `phase` (0..7, MSB first), zero outside; dst must hold
 * ceil((phase+nbits)/32)*4 bytes and may not overlap src.  Used to align a fragment to its global bit offset. */
int b2_bitshift_dev(const void* d_src, uint64_t nbits, int phase, void* d_dst);
int b2_bzip2_encode_range_dev(const void* d_in, size_t n, int level, size_t first, size_t count, int bit_phase,
                              void* d_out, size_t out_cap, uint64_t* out_bits, uint32_t* block_crcs);

/* ---- sharded decode (BASELINE config 5) ----------------------------------- */
/* Every rank holds the compressed stream.  open: scan the magics and decode this rank's share of the
 * candidate blocks; info = {block candidates in the file, first, one-past-last of the own share}.
 * export: 6 x uint64 per own candidate (status, detail, end bit, block length, decoded length, 0).
 * finish: `all` = the exported rows of ALL candidates in order (all-gathered by the caller); walks the
 * block chain, expands and CRC-checks the own blocks into d_out; res = {offset of the own output in the
 * decoded stream, its length, total decoded length, index of the first failing event or -1, its code}. */
int b2_dec_shard_open(const void* d_in, size_t n, int rank, int world, uint64_t* info);
int b2_dec_shard_export(uint64_t* buf);
int b2_dec_shard_finish(const uint64_t* all, int multistream, void* d_out, size_t out_cap, uint64_t* res);

/* ---- instrumentation -------------------------------------------------------------- */
typedef struct b2_stats {
  /* GPU milliseconds of the last call, from CUDA events on the library's stream */
  float ms_total, ms_h2d, ms_d2h;
  float ms_rle1, ms_bwt, ms_mtf, ms_huff, ms_pack;          /* encode stages */
  float ms_scan, ms_hdec, ms_unmtf, ms_ibwt, ms_unrle;      /* decode stages */
  float ms_radix;            /* time inside the radix-sort pass kernel (dominant BWT kernel) */
  uint64_t radix_launches;   /* pass-kernel launches in the last call */
  uint64_t radix_bytes;      /* algorithmic bytes moved by those launches (read+written) */
  uint64_t bwt_bytes;        /* algorithmic bytes of the whole BWT stage (all its kernels) */
  uint64_t bwt_rounds;       /* prefix-doubling rounds executed (max over batches) */
  uint64_t kernel_launches;  /* all kernels launched by the last call */
  uint64_t blocks;           /* bzip2 blocks processed */
  uint64_t raw_bytes, comp_bytes;
  /* MSD path of the forward BWT (bwt_msd.cu): one scatter pass + one shared-memory bucket sort per batch */
  uint64_t msd_launches;       /* batches that took the path (one launch of each of the two kernels) */
  uint64_t msd_scatter_bytes;  /* algorithmic bytes of k_msd_scatter (text in, records out) */
  uint64_t msd_bucket_bytes;   /* algorithmic bytes of k_msd_bucket (records in, column out) */
  float ms_msd_scatter, ms_msd_bucket;
} b2_stats;
void b2_get_stats(b2_stats* s);

/* Per-block trace of the last compress call (for stage-by-stage parity tests). */
typedef struct b2_block_trace {
  int32_t n, pidx, m, alpha, ngroups, nsel;
  uint32_t crc, pad;
  uint64_t raw_start, raw_len, bit_start, bit_len;
} b2_block_trace;
size_t b2_last_trace(b2_block_trace* out, size_t cap);

#ifdef __cplusplus
}
#endif
#endif /* B2BZ_H */
