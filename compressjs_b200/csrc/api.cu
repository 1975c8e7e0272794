// api.cu -- the C ABI declared in include/b2bz.h.  Thin: argument checks, H2D/D2H staging,
// error translation.  All compute is in the CUDA stages (rle1.cu, bwt.cu, mtf.cu, huff.cu,
// decode.cu).  There is no CPU fallback anywhere: without a CUDA device every call fails.
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>
#include "ctx.h"

// stages implemented in the other translation units
void bwt_forward_batch(Ctx& c, const u8* d_T, u8* d_U, const u32* d_n, const u32* h_n, u32 nblk, u32* d_pidx, bool sentinel = false,
                       u32* d_sa_out = nullptr, u32* d_hist_out = nullptr);
u32 crc32_device(Ctx& c, const u8* d_p, size_t n);
void bzip2_compress_device(Ctx& c, const u8* d_in, size_t n, int level, u8* d_out, size_t out_cap, size_t* out_n,
                           size_t first_block, size_t block_count, int bit_phase, bool whole_file, u64* out_bits,
                           std::vector<u32>* crcs_out, size_t* total_blocks, long long spec_first = -2, size_t spec_count = 0,
                           u64* spec_range = nullptr);
void bitshift_device(Ctx& c, const void* src, u64 nbits, int phase, void* dst);
void bzip2_share_summary(Ctx& c, const u8* d_in, size_t n, u64* out);
void bzip2_plan_share(Ctx& c, const u8* d_buf, size_t n, int level, u64 st0, u64 W0, size_t first, size_t count, u64* info);
void bwt_inverse_sentinel(Ctx& c, const u8* d_L, u32 n, u32 pidx, u8* d_out);
size_t bwtc_bound(size_t n);
void bwtc_compress_device(Ctx& c, const u8* d_in, size_t n, int level, u8* d_out, size_t out_cap, size_t* out_n);
void bwtc_decompress_device(Ctx& c, const u8* d_in, size_t n, const u8* h_head, size_t head_n, u8* d_out, size_t out_cap, size_t* out_n);
void bzip2_compress_host(Ctx& c, const u8* h_in, size_t n, int level, u8* d_in, size_t win, u8* d_out, size_t out_cap, u8* h_out,
                         size_t h_out_cap, size_t* out_n, bool pinned_in);
void dec_shard_open(Ctx& c, const u8* d_in, size_t n, int rank, int world, u64* info);
void dec_shard_export(u64* buf);
int dec_shard_finish(Ctx& c, const u64* all, int multistream, u8* d_out, size_t out_cap, u64* res);
int bzip2_decompress_device(Ctx& c, const u8* d_in, size_t n, int multistream, u8* d_out, size_t out_cap, size_t* out_n,
                            bool single_block, u64 bitpos, std::vector<u64>* tab_pos, std::vector<u32>* tab_len,
                            u8** d_out_alloc);

static std::mutex g_mu;
static Ctx* g_ctx = nullptr;
static thread_local std::string g_err;
static std::map<void*, size_t> g_pinned_live;                  // pointers handed to the caller
static std::multimap<size_t, void*> g_pinned_free;             // cached pinned buffers by capacity

void Ctx::collect() {
  float acc[ST_COUNT];
  for (int i = 0; i < ST_COUNT; i++) acc[i] = 0.f;
  for (auto& t : ev_tags) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, ev_pool[t.second].a, ev_pool[t.second].b) == cudaSuccess) acc[t.first] += ms;
  }
  stats.ms_total = acc[ST_TOTAL]; stats.ms_h2d = acc[ST_H2D]; stats.ms_d2h = acc[ST_D2H];
  stats.ms_rle1 = acc[ST_RLE1]; stats.ms_bwt = acc[ST_BWT]; stats.ms_mtf = acc[ST_MTF];
  stats.ms_huff = acc[ST_HUFF]; stats.ms_pack = acc[ST_PACK]; stats.ms_scan = acc[ST_SCAN];
  stats.ms_hdec = acc[ST_HDEC]; stats.ms_unmtf = acc[ST_UNMTF]; stats.ms_ibwt = acc[ST_IBWT];
  stats.ms_unrle = acc[ST_UNRLE]; stats.ms_radix = acc[ST_RADIX];
  stats.ms_msd_scatter = acc[ST_MSD_SCATTER]; stats.ms_msd_bucket = acc[ST_MSD_BUCKET];
  // overlapped copies of the pipelined host path: span from the first to the last copy on their stream
  for (int w = 0; w < 2; w++) {
    float ms = 0.f;
    if (copy_used[w] && cudaEventElapsedTime(&ms, copy_ev[w][0], copy_ev[w][1]) == cudaSuccess) (w ? stats.ms_d2h : stats.ms_h2d) += ms;
  }
}

// ---- small control transfers through mapped pinned memory ----------------------------------
// During the pipelined host encode the copy engines are busy with 64 MiB uploads and per-batch downloads; an
// 8-byte cudaMemcpyAsync would wait behind them for milliseconds.  A tiny kernel moves control data through a
// mapped pinned staging area instead (SM loads/stores over PCIe, no copy engine).
__global__ void k_stage_copy(void* dst, const void* src, size_t bytes) {
  const size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x, step = (size_t)gridDim.x * blockDim.x;
  if ((((size_t)dst | (size_t)src | bytes) & 3) == 0) {
    u32* d = (u32*)dst; const u32* s = (const u32*)src;
    for (size_t i = i0; i < bytes / 4; i += step) d[i] = s[i];
  } else {
    u8* d = (u8*)dst; const u8* s = (const u8*)src;
    for (size_t i = i0; i < bytes; i += step) d[i] = s[i];
  }
}
size_t Ctx::stage_take(size_t bytes) {
  if (!stage_h) {
    CUDA_CHECK(cudaHostAlloc((void**)&stage_h, stage_cap, cudaHostAllocMapped));
    CUDA_CHECK(cudaHostGetDevicePointer((void**)&stage_d, stage_h, 0));
  }
  const size_t need = (bytes + 15) & ~(size_t)15;
  if (stage_used + need > stage_cap) sync();
  const size_t off = stage_used;
  stage_used += need;
  return off;
}
void Ctx::to_device(void* ddst, const void* hsrc, size_t bytes) {
  if (!bytes) return;
  if (bytes > stage_cap / 2) { CUDA_CHECK(cudaMemcpyAsync(ddst, hsrc, bytes, cudaMemcpyHostToDevice, stream)); CUDA_CHECK(cudaStreamSynchronize(stream)); return; }
  const size_t off = stage_take(bytes);
  memcpy(stage_h + off, hsrc, bytes);
  k_stage_copy<<<(unsigned)std::min<size_t>((bytes / 4 + 255) / 256 + 1, 64), 256, 0, stream>>>(ddst, stage_d + off, bytes);
  CUDA_CHECK(cudaGetLastError());
}
void Ctx::to_host(void* hdst, const void* dsrc, size_t bytes) {
  if (!bytes) return;
  if (bytes > stage_cap / 2) { CUDA_CHECK(cudaMemcpyAsync(hdst, dsrc, bytes, cudaMemcpyDeviceToHost, stream)); return; }
  const size_t off = stage_take(bytes);
  k_stage_copy<<<(unsigned)std::min<size_t>((bytes / 4 + 255) / 256 + 1, 64), 256, 0, stream>>>(stage_d + off, dsrc, bytes);
  CUDA_CHECK(cudaGetLastError());
  fetches.push_back({hdst, off, bytes});
}
void Ctx::sync() {
  CUDA_CHECK(cudaStreamSynchronize(stream));
  for (auto& f : fetches) memcpy(f.dst, stage_h + f.off, f.bytes);
  fetches.clear();
  stage_used = 0;
}

static int pick_device() {
  const char* e = getenv("B2_DEVICE");
  if (e && *e) return atoi(e);
  e = getenv("LOCAL_RANK");
  if (e && *e) {
    int cnt = 0;
    if (cudaGetDeviceCount(&cnt) == cudaSuccess && cnt > 0) return atoi(e) % cnt;
  }
  return 0;
}

static Ctx& ctx_locked() {
  if (!g_ctx) {
    int dev = pick_device();
    int cnt = 0;
    cudaError_t e = cudaGetDeviceCount(&cnt);
    if (e != cudaSuccess || cnt == 0)
      throw B2Error{B2_ERR_CUDA, std::string("no CUDA device available (libb2bz has no CPU fallback): ") + cudaGetErrorString(e)};
    if (dev >= cnt) throw B2Error{B2_ERR_CUDA, "requested CUDA device does not exist"};
    CUDA_CHECK(cudaSetDevice(dev));
    Ctx* c = new Ctx();
    c->device = dev;
    memset(&c->stats, 0, sizeof c->stats);
    CUDA_CHECK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CUDA_CHECK(cudaStreamCreateWithFlags(&c->h2d_stream, cudaStreamNonBlocking));
    CUDA_CHECK(cudaStreamCreateWithFlags(&c->d2h_stream, cudaStreamNonBlocking));
    for (int w = 0; w < 2; w++) for (int k = 0; k < 2; k++) CUDA_CHECK(cudaEventCreate(&c->copy_ev[w][k]));
    cudaMemPool_t pool;
    CUDA_CHECK(cudaDeviceGetDefaultMemPool(&pool, dev));
    uint64_t thr = UINT64_MAX;
    CUDA_CHECK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
    const char* b = getenv("B2_BWT_BATCH");
    if (b && atoi(b) > 0) c->bwt_batch = (u32)atoi(b);
    const char* msd = getenv("B2_BWT_MSD");
    if (msd && *msd) c->bwt_msd = atoi(msd) != 0;
    const char* w8 = getenv("B2_BWT_PREFIX8");
    if (w8 && *w8) { c->bwt_wide_forced = true; c->bwt_wide = atoi(w8) != 0; }
    g_ctx = c;
  } else {
    CUDA_CHECK(cudaSetDevice(g_ctx->device));
  }
  return *g_ctx;
}

static void* pinned_alloc(size_t bytes) {
  if (bytes == 0) bytes = 1;
  auto it = g_pinned_free.lower_bound(bytes);
  if (it != g_pinned_free.end() && it->first <= bytes * 2 + 4096) {
    void* p = it->second; size_t cap = it->first;
    g_pinned_free.erase(it);
    g_pinned_live[p] = cap;
    return p;
  }
  void* p = nullptr;
  size_t cap = (bytes + 4095) & ~(size_t)4095;
  CUDA_CHECK(cudaHostAlloc(&p, cap, cudaHostAllocDefault));
  g_pinned_live[p] = cap;
  return p;
}

// give a live pinned buffer back to the cache (caller holds g_mu)
static void pinned_release(void* p) {
  auto it = g_pinned_live.find(p);
  if (it == g_pinned_live.end()) return;
  size_t cap = it->second;
  g_pinned_live.erase(it);
  size_t cached = 0;
  for (auto& kv : g_pinned_free) cached += kv.first;
  if (cached + cap > ((size_t)8 << 30)) cudaFreeHost(p);
  else g_pinned_free.insert({cap, p});
}

template <typename F>
static int guarded(F f) {
  std::lock_guard<std::mutex> lk(g_mu);
  try {
    g_err.clear();
    return f();
  } catch (const B2Error& e) {
    g_err = e.msg;
    if (g_ctx && e.code == B2_ERR_CUDA) { cudaStreamSynchronize(g_ctx->stream); cudaGetLastError(); }
    return e.code;
  } catch (const std::exception& e) {
    g_err = e.what();
    return B2_ERR_CUDA;
  }
}

extern "C" {

int b2_init(int device) {
  return guarded([&]() {
    if (g_ctx && g_ctx->device != device) throw B2Error{B2_ERR_BAD_ARG, "b2_init: context already bound to another device"};
    if (!g_ctx) {
      char buf[16]; snprintf(buf, sizeof buf, "%d", device);
      setenv("B2_DEVICE", buf, 1);
    }
    ctx_locked();
    return 0;
  });
}

void b2_shutdown(void) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (!g_ctx) return;
  cudaSetDevice(g_ctx->device);
  cudaStreamSynchronize(g_ctx->stream);
  for (auto& kv : g_pinned_free) cudaFreeHost(kv.second);
  g_pinned_free.clear();
  for (auto& e : g_ctx->ev_pool) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
  if (g_ctx->stage_h) cudaFreeHost(g_ctx->stage_h);
  cudaStreamDestroy(g_ctx->stream);
  cudaStreamDestroy(g_ctx->h2d_stream);
  cudaStreamDestroy(g_ctx->d2h_stream);
  for (int w = 0; w < 2; w++) for (int k = 0; k < 2; k++) cudaEventDestroy(g_ctx->copy_ev[w][k]);
  delete g_ctx;
  g_ctx = nullptr;
}

const char* b2_last_error(void) { return g_err.c_str(); }

void b2_free(void* p) {
  if (!p) return;
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_pinned_live.find(p) == g_pinned_live.end()) { free(p); return; }
  pinned_release(p);
}

void b2_get_stats(b2_stats* s) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_ctx) *s = g_ctx->stats; else memset(s, 0, sizeof *s);
}

size_t b2_last_trace(b2_block_trace* out, size_t cap) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (!g_ctx) return 0;
  size_t n = g_ctx->trace.size();
  if (out) for (size_t i = 0; i < n && i < cap; i++) out[i] = g_ctx->trace[i];
  return n;
}

// ---- BWT ---------------------------------------------------------------------------------
int b2_bwt_cyclic_batch(const uint8_t* T, uint8_t* U, const uint64_t* offs, const int32_t* lens, int32_t* pidx, size_t nblocks) {
  return guarded([&]() {
    Ctx& c = ctx_locked();
    c.reset_call();
    for (size_t k = 0; k < nblocks; k++)
      if (lens[k] < 0 || lens[k] > 900000) throw B2Error{B2_ERR_BAD_ARG, "block length out of range (0..900000)"};
    {
      StageScope tot(c, ST_TOTAL);
      for (size_t k0 = 0; k0 < nblocks; k0 += c.bwt_batch) {
        const u32 nb = (u32)std::min<size_t>(c.bwt_batch, nblocks - k0);
        DBuf<u8> dT(c, (size_t)nb << SEG_SHIFT), dU(c, (size_t)nb << SEG_SHIFT);
        DBuf<u32> dn(c, nb), dp(c, nb);
        std::vector<u32> hn(nb);
        for (u32 b = 0; b < nb; b++) {
          hn[b] = (u32)lens[k0 + b];
          if (hn[b]) CUDA_CHECK(cudaMemcpyAsync(dT.p + ((size_t)b << SEG_SHIFT), T + offs[k0 + b], hn[b], cudaMemcpyHostToDevice, c.stream));
        }
        CUDA_CHECK(cudaMemcpyAsync(dn, hn.data(), nb * 4, cudaMemcpyHostToDevice, c.stream));
        CUDA_CHECK(cudaMemsetAsync(dp, 0, nb * 4, c.stream));
        {
          StageScope s(c, ST_BWT);
          bwt_forward_batch(c, dT, dU, dn, hn.data(), nb, dp);
        }
        std::vector<u32> hp(nb);
        CUDA_CHECK(cudaMemcpyAsync(hp.data(), dp, nb * 4, cudaMemcpyDeviceToHost, c.stream));
        for (u32 b = 0; b < nb; b++)
          if (hn[b]) CUDA_CHECK(cudaMemcpyAsync(U + offs[k0 + b], dU.p + ((size_t)b << SEG_SHIFT), hn[b], cudaMemcpyDeviceToHost, c.stream));
        c.sync();
        for (u32 b = 0; b < nb; b++) pidx[k0 + b] = (int32_t)hp[b];
        c.stats.blocks += nb;
      }
    }
    c.sync();
    c.collect();
    return 0;
  });
}

int32_t b2_bwt_cyclic(const uint8_t* T, uint8_t* U, int32_t n) {
  if (n <= 1) {  // lib/BWT.js:376-379
    if (n == 1) U[0] = T[0];
    return 0;
  }
  uint64_t off = 0;
  int32_t pidx = 0;
  int rc = b2_bwt_cyclic_batch(T, U, &off, &n, &pidx, 1);
  return rc < 0 ? rc : pidx;
}

// suffix array / sentinel BWT of one string (<= 2^20 - 2 bytes, the slot size of the block pipeline)
static int sentinel_sort(const uint8_t* T, int32_t n, int32_t* SA, uint8_t* U, int32_t* pidx1) {
  return guarded([&]() {
    if (n < 2 || (uint32_t)n > SEG_SIZE - 2) throw B2Error{B2_ERR_BAD_ARG, "length out of range (2..1048574)"};
    Ctx& c = ctx_locked();
    c.reset_call();
    {
      StageScope tot(c, ST_TOTAL);
      DBuf<u8> dT(c, SEG_SIZE), dU(c, SEG_SIZE);
      DBuf<u32> dn(c, 1), dp(c, 1), dsa(c, SEG_SIZE);
      u32 hn = (u32)n, hp = 0;
      CUDA_CHECK(cudaMemcpyAsync(dT.p, T, n, cudaMemcpyHostToDevice, c.stream));
      c.to_device(dn, &hn, 4);
      CUDA_CHECK(cudaMemsetAsync(dp, 0, 4, c.stream));
      bwt_forward_batch(c, dT, U ? dU.p : nullptr, dn, &hn, 1, dp, true, SA ? dsa.p : nullptr);
      c.to_host(&hp, dp, 4);
      if (SA) CUDA_CHECK(cudaMemcpyAsync(SA, dsa.p, (size_t)n * 4, cudaMemcpyDeviceToHost, c.stream));
      if (U) CUDA_CHECK(cudaMemcpyAsync(U, dU.p, n, cudaMemcpyDeviceToHost, c.stream));
      c.sync();
      if (pidx1) *pidx1 = (int32_t)hp;
    }
    c.sync();
    c.collect();
    return 0;
  });
}
int b2_suffixsort(const uint8_t* T, int32_t* SA, int32_t n) {
  if (n <= 1) {  // lib/BWT.js:307-310
    if (n == 1) SA[0] = 0;
    return 0;
  }
  return sentinel_sort(T, n, SA, nullptr, nullptr);
}
int32_t b2_bwt_sentinel(const uint8_t* T, uint8_t* U, int32_t n) {
  if (n <= 1) {  // lib/BWT.js:332-335
    if (n == 1) U[0] = T[0];
    return n < 0 ? 0 : n;
  }
  int32_t p = 0;
  int rc = sentinel_sort(T, n, nullptr, U, &p);
  return rc < 0 ? rc : p;
}
int b2_bwt_inverse(const uint8_t* L, uint8_t* out, int32_t n, int32_t pidx) {
  if (n <= 0) return 0;
  if (n == 1) { out[0] = L[0]; return 0; }
  return guarded([&]() {
    if ((uint32_t)n > SEG_SIZE - 2) throw B2Error{B2_ERR_BAD_ARG, "length out of range (0..1048574)"};
    if (pidx < 0 || pidx > n) throw B2Error{B2_ERR_BAD_ARG, "primary index out of range"};
    Ctx& c = ctx_locked();
    c.reset_call();
    {
      StageScope tot(c, ST_TOTAL);
      DBuf<u8> dL(c, n), dO(c, n);
      CUDA_CHECK(cudaMemcpyAsync(dL.p, L, n, cudaMemcpyHostToDevice, c.stream));
      bwt_inverse_sentinel(c, dL, (u32)n, (u32)pidx, dO);
      CUDA_CHECK(cudaMemcpyAsync(out, dO.p, n, cudaMemcpyDeviceToHost, c.stream));
    }
    c.sync();
    c.collect();
    return 0;
  });
}

// ---- BWTC container (experimental, see bwtc.cu) ----------------------------------------------
int b2_bwtc_compress(const uint8_t* in, size_t n, int level, uint8_t** out, size_t* out_n) {
  return guarded([&]() {
    Ctx& c = ctx_locked();
    c.reset_call();
    const size_t cap = bwtc_bound(n);
    size_t produced = 0;
    void* host = nullptr;
    {
      StageScope tot(c, ST_TOTAL);
      DBuf<u8> din(c, n ? n : 1), dout(c, cap);
      if (n) CUDA_CHECK(cudaMemcpyAsync(din.p, in, n, cudaMemcpyHostToDevice, c.stream));
      bwtc_compress_device(c, din, n, level, dout, cap, &produced);
      host = pinned_alloc(produced);
      CUDA_CHECK(cudaMemcpyAsync(host, dout.p, produced, cudaMemcpyDeviceToHost, c.stream));
    }
    c.sync();
    c.collect();
    c.stats.raw_bytes = n; c.stats.comp_bytes = produced;
    *out = (uint8_t*)host; *out_n = produced;
    return 0;
  });
}
int b2_bwtc_decompress(const uint8_t* in, size_t n, uint8_t** out, size_t* out_n) {
  return guarded([&]() {
    Ctx& c = ctx_locked();
    c.reset_call();
    size_t produced = 0;
    void* host = nullptr;
    {
      StageScope tot(c, ST_TOTAL);
      // the decoded size is in the header: parse it before any device work
      size_t pos = 4; uint64_t fs = 0;
      if (n < 5 || memcmp(in, "bwtc", 4)) throw B2Error{B2_ERR_BAD_MAGIC, "Bad magic"};
      for (;;) {
        if (pos >= n || pos > 4 + 9) throw B2Error{B2_ERR_DATA_ERROR, "truncated or oversized BWTC header"};
        const uint32_t ch = in[pos++];
        if (ch & 0x80) { fs += ch & 0x7F; break; }
        fs = (fs + ch) * 128;
      }
      if (fs == 0) throw B2Error{B2_ERR_BAD_ARG, "BWTC streams of unknown size are not supported"};
      const size_t size = (size_t)(fs - 1);
      DBuf<u8> din(c, n), dout(c, size ? size : 1);
      CUDA_CHECK(cudaMemcpyAsync(din.p, in, n, cudaMemcpyHostToDevice, c.stream));
      bwtc_decompress_device(c, din, n, in, std::min<size_t>(n, 16), dout, size, &produced);
      host = pinned_alloc(produced);
      if (produced) CUDA_CHECK(cudaMemcpyAsync(host, dout.p, produced, cudaMemcpyDeviceToHost, c.stream));
    }
    c.sync();
    c.collect();
    c.stats.raw_bytes = produced; c.stats.comp_bytes = n;
    *out = (uint8_t*)host; *out_n = produced;
    return 0;
  });
}

uint32_t b2_crc32_bzip2(const uint8_t* p, size_t n) {
  uint32_t crc = 0;
  int rc = guarded([&]() {
    Ctx& c = ctx_locked();
    c.reset_call();
    DBuf<u8> d(c, n ? n : 1);
    if (n) CUDA_CHECK(cudaMemcpyAsync(d, p, n, cudaMemcpyHostToDevice, c.stream));
    crc = crc32_device(c, d, n);
    return 0;
  });
  (void)rc;
  return crc;
}

// ---- bzip2 -------------------------------------------------------------------------------
size_t b2_bzip2_bound(size_t n) {
  // worst case per block: Huffman codes up to 20 bits for <= n+1 symbols would be 2.5x, but the
  // flat 2nd table bounds every 50-group at 50*ceil(log2(258)) = 450 bits = 9 bits/symbol;
  // RLE1 can expand the block stream by 5/4.  Use a comfortable 1.5x + per-block headers.
  size_t blocks = n / 99981 + 2;
  return n + n / 2 + blocks * 4096 + 64;
}

int b2_bzip2_compress_dev(const void* d_in, size_t n, int level, void* d_out, size_t out_cap, size_t* out_n) {
  return guarded([&]() {
    if (level < 1 || level > 9) throw B2Error{B2_ERR_BAD_LEVEL, "Invalid block size multiplier"};
    Ctx& c = ctx_locked();
    c.reset_call();
    {
      StageScope tot(c, ST_TOTAL);
      bzip2_compress_device(c, (const u8*)d_in, n, level, (u8*)d_out, out_cap, out_n, 0, (size_t)-1, 0, true, nullptr, nullptr, nullptr);
    }
    c.sync();
    c.collect();
    c.stats.raw_bytes = n; c.stats.comp_bytes = *out_n;
    return 0;
  });
}

int b2_bzip2_compress(const uint8_t* in, size_t n, int level, uint8_t** out, size_t* out_n) {
  return guarded([&]() {
    if (level < 1 || level > 9) throw B2Error{B2_ERR_BAD_LEVEL, "Invalid block size multiplier"};
    Ctx& c = ctx_locked();
    c.reset_call();
    size_t cap = b2_bzip2_bound(n), produced = 0;
    void* host = pinned_alloc(cap);
    try {
      StageScope tot(c, ST_TOTAL);
      // Inputs above the streaming window pass through the device in windows (bounded device memory: files larger than
      // HBM); $B2_STREAM_WINDOW sets the window in bytes (default 8 GiB, at least 64 MiB -- a test hook allows less).
      size_t win = (size_t)8 << 30;
      if (const char* e = getenv("B2_STREAM_WINDOW")) { const long long v = atoll(e); if (v >= (1 << 20)) win = (size_t)v; }
      if (win > n) win = n;
      const size_t dcap = win < n ? b2_bzip2_bound(win) + 64 : cap;
      DBuf<u8> din(c, win ? win : 1), dout(c, dcap);
      c.sync();  // the buffers are used from the copy streams as well
      cudaPointerAttributes pa;
      const bool pinned_in = n && cudaPointerGetAttributes(&pa, in) == cudaSuccess && pa.type == cudaMemoryTypeHost;
      cudaGetLastError();
      bzip2_compress_host(c, in, n, level, din, win, dout, dcap, (u8*)host, cap, &produced, pinned_in);
    } catch (...) {
      pinned_release(host);
      throw;
    }
    c.sync();
    c.collect();
    c.stats.raw_bytes = n; c.stats.comp_bytes = produced;
    *out = (uint8_t*)host; *out_n = produced;
    return 0;
  });
}

int b2_bzip2_plan(const void* d_in, size_t n, int level, size_t* total_blocks) {
  return guarded([&]() {
    if (level < 1 || level > 9) throw B2Error{B2_ERR_BAD_LEVEL, "Invalid block size multiplier"};
    Ctx& c = ctx_locked();
    c.reset_call();
    size_t dummy = 0;
    bzip2_compress_device(c, (const u8*)d_in, n, level, nullptr, 0, &dummy, 0, 0, 0, false, nullptr, nullptr, total_blocks);
    c.sync();
    return 0;
  });
}

int b2_dec_shard_open(const void* d_in, size_t n, int rank, int world, uint64_t* info) {
  return guarded([&]() {
    if (world < 1 || rank < 0 || rank >= world) throw B2Error{B2_ERR_BAD_ARG, "bad rank/world"};
    Ctx& c = ctx_locked();
    c.reset_call();
    {
      StageScope tot(c, ST_TOTAL);
      dec_shard_open(c, (const u8*)d_in, n, rank, world, info);
    }
    c.sync();
    c.collect();
    return 0;
  });
}
int b2_dec_shard_export(uint64_t* buf) {
  return guarded([&]() { dec_shard_export(buf); return 0; });
}
int b2_dec_shard_finish(const uint64_t* all, int multistream, void* d_out, size_t out_cap, uint64_t* res) {
  return guarded([&]() {
    Ctx& c = ctx_locked();
    return dec_shard_finish(c, all, multistream, (u8*)d_out, out_cap, res);
  });
}

int b2_bitshift_dev(const void* d_src, uint64_t nbits, int phase, void* d_dst) {
  return guarded([&]() {
    if (phase < 0 || phase > 7 || (((size_t)d_src | (size_t)d_dst) & 3)) throw B2Error{B2_ERR_BAD_ARG, "bad phase or unaligned buffers"};
    Ctx& c = ctx_locked();
    bitshift_device(c, d_src, nbits, phase, d_dst);
    return 0;
  });
}

int b2_bzip2_plan_spec(const void* d_in, size_t n, int level, int rank, int world, uint64_t* info) {
  return guarded([&]() {
    if (level < 1 || level > 9) throw B2Error{B2_ERR_BAD_LEVEL, "Invalid block size multiplier"};
    if (world < 1 || rank < 0 || rank >= world) throw B2Error{B2_ERR_BAD_ARG, "bad rank/world"};
    Ctx& c = ctx_locked();
    c.reset_call();
    size_t dummy = 0;
    bzip2_compress_device(c, (const u8*)d_in, n, level, nullptr, 0, &dummy, 0, 0, 0, false, nullptr, nullptr, nullptr, -1,
                          ((size_t)rank << 32) | (size_t)world, info);
    c.sync();
    return 0;
  });
}

int b2_bzip2_share_summary(const void* d_share, size_t n, uint64_t* summary) {
  return guarded([&]() {
    Ctx& c = ctx_locked();
    c.reset_call();
    bzip2_share_summary(c, (const u8*)d_share, n, summary);
    c.sync();
    return 0;
  });
}

int b2_bzip2_plan_share(const void* d_buf, size_t n, int level, uint64_t state_in, uint64_t w_in, size_t first, size_t count, uint64_t* info) {
  return guarded([&]() {
    if (level < 1 || level > 9) throw B2Error{B2_ERR_BAD_LEVEL, "Invalid block size multiplier"};
    Ctx& c = ctx_locked();
    c.reset_call();
    bzip2_plan_share(c, (const u8*)d_buf, n, level, state_in, w_in, first, count, info);
    c.sync();
    return 0;
  });
}

int b2_bzip2_encode_range_dev(const void* d_in, size_t n, int level, size_t first, size_t count, int bit_phase, void* d_out,
                              size_t out_cap, uint64_t* out_bits, uint32_t* block_crcs) {
  return guarded([&]() {
    if (level < 1 || level > 9) throw B2Error{B2_ERR_BAD_LEVEL, "Invalid block size multiplier"};
    if (bit_phase < 0 || bit_phase > 7) throw B2Error{B2_ERR_BAD_ARG, "bit_phase must be 0..7"};
    Ctx& c = ctx_locked();
    c.reset_call();
    size_t bytes = 0;
    std::vector<u32> crcs;
    {
      StageScope tot(c, ST_TOTAL);
      bzip2_compress_device(c, (const u8*)d_in, n, level, (u8*)d_out, out_cap, &bytes, first, count, bit_phase, false, out_bits, &crcs, nullptr);
    }
    c.sync();
    c.collect();
    if (block_crcs) for (size_t i = 0; i < crcs.size(); i++) block_crcs[i] = crcs[i];
    return 0;
  });
}

static int decode_common(const uint8_t* in, size_t n, int multistream, bool single, u64 bitpos, uint8_t** out, size_t* out_n,
                         std::vector<u64>* tp, std::vector<u32>* tl) {
  Ctx& c = ctx_locked();
  c.reset_call();
  size_t produced = 0;
  void* host = nullptr;
  int rc = 0;
  {
    StageScope tot(c, ST_TOTAL);
    DBuf<u8> din(c, n + 16);
    {
      StageScope s(c, ST_H2D);
      CUDA_CHECK(cudaMemsetAsync(din.p + n, 0, 16, c.stream));
      if (n) CUDA_CHECK(cudaMemcpyAsync(din, in, n, cudaMemcpyHostToDevice, c.stream));
    }
    u8* dres = nullptr;
    struct DresGuard { Ctx& c; u8*& p; ~DresGuard() { if (p) { c.dfree(p); p = nullptr; } } } dres_guard{c, dres};  // also on exceptions
    rc = bzip2_decompress_device(c, din, n, multistream, nullptr, 0, &produced, single, bitpos, tp, tl, &dres);
    if (rc == 0 && out) {
      host = pinned_alloc(produced);
      try {
        StageScope s(c, ST_D2H);
        if (produced) CUDA_CHECK(cudaMemcpyAsync(host, dres, produced, cudaMemcpyDeviceToHost, c.stream));
        c.sync();
      } catch (...) {
        pinned_release(host);
        throw;
      }
    }
    c.sync();
  }
  c.sync();
  c.collect();
  c.stats.raw_bytes = produced; c.stats.comp_bytes = n;
  if (rc) return rc;
  if (out) { *out = (uint8_t*)host; *out_n = produced; }
  return 0;
}

int b2_bzip2_decompress(const uint8_t* in, size_t n, int multistream, uint8_t** out, size_t* out_n) {
  return guarded([&]() { return decode_common(in, n, multistream, false, 0, out, out_n, nullptr, nullptr); });
}

int b2_bzip2_decompress_block(const uint8_t* in, size_t n, uint64_t bitpos, uint8_t** out, size_t* out_n) {
  return guarded([&]() { return decode_common(in, n, 0, true, bitpos, out, out_n, nullptr, nullptr); });
}

int b2_bzip2_table(const uint8_t* in, size_t n, int multistream, uint64_t** bitpos, uint32_t** sizes, size_t* count) {
  return guarded([&]() {
    std::vector<u64> tp; std::vector<u32> tl;
    int rc = decode_common(in, n, multistream, false, 0, nullptr, nullptr, &tp, &tl);
    if (rc) return rc;
    *count = tp.size();
    *bitpos = (uint64_t*)malloc(sizeof(uint64_t) * (tp.size() + 1));
    *sizes = (uint32_t*)malloc(sizeof(uint32_t) * (tp.size() + 1));
    for (size_t i = 0; i < tp.size(); i++) { (*bitpos)[i] = tp[i]; (*sizes)[i] = tl[i]; }
    return 0;
  });
}

int b2_bzip2_decompress_dev(const void* d_in, size_t n, int multistream, void* d_out, size_t out_cap, size_t* out_n) {
  return guarded([&]() {
    Ctx& c = ctx_locked();
    c.reset_call();
    int rc;
    {
      StageScope tot(c, ST_TOTAL);
      rc = bzip2_decompress_device(c, (const u8*)d_in, n, multistream, (u8*)d_out, out_cap, out_n, false, 0, nullptr, nullptr, nullptr);
    }
    c.sync();
    c.collect();
    c.stats.raw_bytes = *out_n; c.stats.comp_bytes = n;
    return rc;
  });
}

}  // extern "C"
