#include "ctx.h"
#include <vector>
u32 crc32_device(Ctx& c, const u8* d_p, size_t n) { throw B2Error{-200,"stub"}; }
void bzip2_compress_device(Ctx& c, const u8* d_in, size_t n, int level, u8* d_out, size_t out_cap, size_t* out_n,
                           size_t first_block, size_t block_count, int bit_phase, bool whole_file, u64* out_bits,
                           std::vector<u32>* crcs_out, size_t* total_blocks) { throw B2Error{-200,"stub"}; }
int bzip2_decompress_device(Ctx& c, const u8* d_in, size_t n, int multistream, u8* d_out, size_t out_cap, size_t* out_n,
                            bool single_block, u64 bitpos, std::vector<u64>* tab_pos, std::vector<u32>* tab_len,
                            u8** d_out_alloc) { throw B2Error{-200,"stub"}; }
