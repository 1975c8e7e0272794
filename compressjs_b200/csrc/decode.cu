#include "ctx.h"
#include <vector>
int bzip2_decompress_device(Ctx& c, const u8* d_in, size_t n, int multistream, u8* d_out, size_t out_cap, size_t* out_n,
                            bool single_block, u64 bitpos, std::vector<u64>* tab_pos, std::vector<u32>* tab_len,
                            u8** d_out_alloc) { throw B2Error{-200,"decode not built yet"}; }
