// bwtc.cu -- compressjs' BWTC container (lib/BWTC.js:12-231) on the GPU.
//
// STATUS: new in round 1.  The serial model / range-coder code (bwtc_core.cuh) is verified on its host build against
// the oracle (tests/test_host_api.py::test_bwtc_core_matches_oracle); the kernels and the orchestration below ran on a
// B200 once with the last seconds of the round's GPU budget (tools/bwtc_try.py -> profiles/r1e_bwtc_try.txt: 12 cases,
// levels 1 and 9, one to three blocks, encode bit exact and decode correct).  Not yet measured at BASELINE config 4.
//
// Per block the container needs: sentinel BWT (lib/BWT.js:328-350), MTF over the used bytes, zero runs as
// RUNA/RUNB digits, an adaptive model (Fenwick tree, or the deferred-sum model below level 6) that turns every
// symbol into a (sy_f, lt_f, tot_f) triple, and ONE range coder that runs over the whole file.  The first three are
// the block-parallel kernels of the bzip2 path (bwt.cu in sentinel mode, mtf.cu: the symbol values are identical,
// bzip2 merely appends an end-of-block symbol).  The model is a serial chain per block (k_bwtc_model: one thread per
// block, blocks in parallel), the coder a serial chain per file (k_bwtc_code: one thread; its `range` recurrence
// needs an integer division per symbol and nothing shortens it).  Decode cannot even split model and coder:
// k_bwtc_decode is one thread for the whole stream, followed by the inverse sentinel BWT per block (decode.cu).
#include <algorithm>
#include <vector>
#include "enc.h"
#include "bwtc_core.cuh"

void bwt_forward_batch(Ctx& c, const u8* d_T, u8* d_U, const u32* d_n, const u32* h_n, u32 nblk, u32* d_pidx, bool sentinel, u32* d_sa_out, u32* d_hist_out = nullptr);
void bwt_inverse_sentinel(Ctx& c, const u8* d_L, u32 n, u32 pidx, u8* d_out);

struct BwtcState {
  bc_enc rc;
  u32 overflow;  // a block's triples did not fit its buffer
  u32 pad;
};

__global__ void k_bwtc_start(BwtcState* st, u8* out, u64 cap, u32 finalByte, u32 level) {
  if (threadIdx.x || blockIdx.x) return;
  bc_enc_start(&st->rc, out, cap, finalByte);                 // lib/BWTC.js:13-14
  bc_enc_code(&st->rc, bc_triple(1, level, 256));             // encoder.encodeByte(blockSize), :21
  st->overflow = 0;
}

// one thread per block (lane 0 of its warp): header + model -> triples
__global__ void __launch_bounds__(32)
k_bwtc_model(const u16* __restrict__ sym, const u32* __restrict__ d_m, const u32* __restrict__ d_n, const u32* __restrict__ d_pidx1,
             const u32* __restrict__ d_used, u32 blockSize, int fast, u64* __restrict__ triples, u32 tcap, u32* __restrict__ tcount) {
  if (threadIdx.x) return;
  const u32 b = blockIdx.x;
  bc_model model;
  bc_emit e;
  e.t = triples + (size_t)b * tcap; e.n = 0; e.cap = tcap;
  const u32 m = d_m[b];
  bc_block_triples(&e, &model, blockSize, d_n[b], d_pidx1[b], d_used + (size_t)b * 8, sym + ((size_t)b << SEG_SHIFT), m ? m - 1 : 0, fast);
  tcount[b] = e.n;
}

// Fenwick model (levels 6..9) with one WARP per block: the tree lives in shared memory, the nodes on the path from a
// leaf to the root sit in different lanes (lane l owns node (numSyms + symbol) >> l), so the cumulative frequency is one
// warp reduction over the left siblings and the update one add per lane; the rescale (every ~128 symbols,
// lib/FenwickModel.js:125-161) runs over the leaves in parallel and re-sums the tree level by level.
// Same triples as bc_fen_encode (bwtc_core.cuh), which the host tests check against the oracle.
struct FenWarp { u32 tree[2 * 260]; };
__device__ __forceinline__ void fenw_sum(u32* tree, u32 numSyms, u32 lane) {
  for (int k = 8; k >= 0; k--) {               // nodes [2^k, 2^(k+1)) only depend on deeper ones
    const u32 lo = 1u << k, hi = min(2u << k, numSyms);
    for (u32 i = lo + lane; i < hi; i += 32) tree[i] = tree[2 * i] + tree[2 * i + 1];
    __syncwarp();
  }
}
__device__ __forceinline__ void fenw_rescale(u32* tree, u32 numSyms, u32 lane) {
  u32 noEscape = 1;
  for (u32 i = lane; i + 1 < numSyms; i += 32) {
    u32 prob = tree[numSyms + i];
    if (prob & BC_ESC_MASK) { noEscape = 0; continue; }
    prob = (prob & BC_SCALE_MASK) >> 1;
    if (prob == 0) { prob = 1; noEscape = 0; }
    tree[numSyms + i] = prob;
  }
  noEscape = __all_sync(FULL_MASK, noEscape);
  if (lane == 0) {
    u32 prob = (tree[2 * numSyms - 1] & BC_SCALE_MASK) >> 1;
    if (noEscape) prob = 0; else if (prob == 0) prob = 1u << 16;
    tree[2 * numSyms - 1] = prob;
  }
  __syncwarp();
  fenw_sum(tree, numSyms, lane);
}
// one trip up the tree (bc_fen_step); every lane returns the triple
__device__ __forceinline__ u64 fenw_step(u32* tree, u32 numSyms, u32 symbol, u32 sy_leaf, int esc, u32 lane) {
  const u32 i = numSyms + symbol;
  u32 mask = BC_SYM_MASK, shift = 16, update = BC_F_PROB_INCR << 16;
  const u32 root = tree[1];
  if (esc) { mask = BC_ESC_MASK; update -= 1; shift = 0; }
  else if (symbol == numSyms - 1 && (root & BC_ESC_MASK) == 1) update = 0u - tree[i];  // the last escape
  const u32 node = lane < 31 ? (i >> lane) : 0u;
  const bool on_path = node > 1;
  u32 contrib = (on_path && (node & 1u)) ? tree[node - 1] : 0u;   // left sibling
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) contrib += __shfl_xor_sync(FULL_MASK, contrib, o);
  __syncwarp();
  if (on_path) tree[node] += update;
  if (lane == 31) tree[1] = root + update;
  __syncwarp();
  const u64 tr = bc_triple((sy_leaf & mask) >> shift, (contrib & mask) >> shift, (root & mask) >> shift);
  if ((((root + update) & BC_SYM_MASK) >> 16) >= BC_F_PROB_MAX) fenw_rescale(tree, numSyms, lane);
  return tr;
}
__global__ void __launch_bounds__(32)
k_bwtc_model_fenwick(const u16* __restrict__ sym, const u32* __restrict__ d_m, const u32* __restrict__ d_n, const u32* __restrict__ d_pidx1,
                     const u32* __restrict__ d_used, u32 blockSize, u64* __restrict__ triples, u32 tcap, u32* __restrict__ tcount) {
  __shared__ FenWarp fw;
  __shared__ u32 s_hdr[2];
  const u32 b = blockIdx.x, lane = threadIdx.x;
  u64* out = triples + (size_t)b * tcap;
  const u32 m = d_m[b], nsym = m ? m - 1 : 0;
  if (lane == 0) {
    bc_emit e;
    e.t = out; e.n = 0; e.cap = tcap;
    s_hdr[1] = bc_block_header(&e, blockSize, d_n[b], d_pidx1[b], d_used + (size_t)b * 8);
    s_hdr[0] = e.n;
  }
  __syncwarp();
  u32 n_out = s_hdr[0];
  const u32 size = s_hdr[1] + 1, numSyms = size + 1;      // bc_fen_init(alphabetSize + 1)
  u32* tree = fw.tree;
  for (u32 i = lane; i < 2 * 260; i += 32) tree[i] = 0;
  __syncwarp();
  for (u32 i = lane; i < size; i += 32) tree[numSyms + i] = 1;
  if (lane == 0) tree[numSyms + size] = BC_F_PROB_INCR << 16;
  __syncwarp();
  fenw_sum(tree, numSyms, lane);
  const u16* sp = sym + ((size_t)b << SEG_SHIFT);
  for (u32 k0 = 0; k0 < nsym; k0 += 32) {
    const u32 mine = k0 + lane < nsym ? sp[k0 + lane] : 0u;
    const u32 cnt = min(32u, nsym - k0);
    for (u32 j = 0; j < cnt; j++) {
      const u32 symbol = __shfl_sync(FULL_MASK, mine, j);
      const u32 sy_leaf = tree[numSyms + symbol];
      u64 t0, t1 = 0;
      u32 produced = 1;
      if ((sy_leaf & BC_SYM_MASK) == 0) {
        const u32 escSym = numSyms - 1;
        t0 = fenw_step(tree, numSyms, escSym, tree[numSyms + escSym], 0, lane);
        t1 = fenw_step(tree, numSyms, symbol, sy_leaf, 1, lane);
        produced = 2;
      } else {
        t0 = fenw_step(tree, numSyms, symbol, sy_leaf, 0, lane);
      }
      if (lane == 0) {
        if (n_out < tcap) out[n_out] = t0;
        if (produced == 2 && n_out + 1 < tcap) out[n_out + 1] = t1;
      }
      n_out += produced;
    }
  }
  if (lane == 0) tcount[b] = n_out;
}

// One warp: the blocks of a batch through the file's range coder.  The recurrence on (low, range) is serial and runs in
// lane 0; what can be taken off its critical path is done by the whole warp, 32 symbols at a time: the coalesced load
// of the triples, their unpacking, and a reciprocal of every total so that the serial step replaces the integer
// division range / tot_f (RangeCoder.js:83) by a multiply-high and an exact correction.
__device__ __forceinline__ void bc_enc_code_fast(bc_enc* rc, u32 sy_f, u32 lt_f, u32 tot_f, u32 magic) {
  bc_enc_normalize(rc);
  u32 r = __umulhi(rc->range, magic);          // floor(range * floor((2^32 - 1) / tot) / 2^32) <= range / tot
  u32 rem = rc->range - r * tot_f;
  while (rem >= tot_f) { r++; rem -= tot_f; }  // at most two steps
  const u32 tmp = r * lt_f;
  rc->low += tmp;
  if (lt_f + sy_f < tot_f) rc->range = r * sy_f; else rc->range -= tmp;
}
__global__ void __launch_bounds__(32) k_bwtc_code(BwtcState* st, const u64* __restrict__ triples, const u32* __restrict__ tcount, u32 nblk, u32 tcap) {
  __shared__ uint4 s_t[2][32];   // (sy, lt, tot, reciprocal) of a batch of 32 symbols, double buffered
  const u32 lane = threadIdx.x;
  bc_enc rc = st->rc;
  u32 overflow = st->overflow;
  for (u32 b = 0; b < nblk && !overflow; b++) {
    const u32 n = tcount[b];
    if (n > tcap) { overflow = 1; break; }
    const u64* t = triples + (size_t)b * tcap;
    u64 nxt = lane < n ? t[lane] : 0ull;        // one batch ahead: the load latency hides behind the serial steps
    for (u32 k0 = 0, it = 0; k0 < n; k0 += 32, it++) {
      const u64 tr = nxt;
      if (k0 + 32 + lane < n) nxt = t[k0 + 32 + lane];
      const u32 tot = (u32)(tr >> 42);
      s_t[it & 1][lane] = make_uint4((u32)(tr & 0x1FFFFF), (u32)((tr >> 21) & 0x1FFFFF), tot, tot ? 0xFFFFFFFFu / tot : 0u);
      __syncwarp();
      if (lane == 0) {
        const u32 cnt = min(32u, n - k0);
        const uint4* q = s_t[it & 1];
        uint4 cur = q[0];
        for (u32 j = 0; j < cnt; j++) {
          const uint4 nx = q[(j + 1) & 31];     // fetched one step ahead of its use
          bc_enc_code_fast(&rc, cur.x, cur.y, cur.z, cur.w);
          cur = nx;
        }
      }
      // (no second barrier: the next batch goes to the other buffer, and lane 0 is past this one by the time the
      // buffer is written again two batches later -- all lanes wait for lane 0 at the next __syncwarp)
    }
  }
  if (lane == 0) {
    st->rc = rc;
    st->overflow = overflow;
  }
}

__global__ void k_bwtc_finish(BwtcState* st) {
  if (threadIdx.x || blockIdx.x) return;
  bc_enc_code(&st->rc, bc_triple(1, 2, 3));                   // "no more blocks", lib/BWTC.js:141
  bc_enc_finish(&st->rc);
}

size_t bwtc_bound(size_t n) { return n + n / 8 + (n / 100000 + 2) * 1024 + 64; }

// BWTC.compressFile on device buffers; the size of the input is written into the header (Util.js:118-134: a buffer
// input has a known size).  out_cap >= bwtc_bound(n).
void bwtc_compress_device(Ctx& c, const u8* d_in, size_t n, int level, u8* d_out, size_t out_cap, size_t* out_n) {
  if (level < 1 || level > 9) level = 9;                      // lib/BWTC.js:16-19
  const u32 blockSize = (u32)level * 100000u;
  const int fast = level <= 5;                                // :22
  if (out_cap < 32) throw B2Error{B2_ERR_BAD_ARG, "output buffer too small"};
  u8 hdr[24];
  u32 finalByte = 0;
  const u32 hlen = bc_file_header(hdr, n, &finalByte);
  c.to_device(d_out, hdr, hlen);
  DBuf<BwtcState> st(c, 1);
  k_bwtc_start<<<1, 1, 0, c.stream>>>(st, d_out + hlen, out_cap - hlen, finalByte, (u32)level);
  KLAUNCH(c); KCHECK();
  const size_t nblocks = (n + blockSize - 1) / blockSize;
  if (nblocks) {
    const u32 B = (u32)std::min<size_t>(c.bwt_batch, nblocks);
    const u32 tcap = 2 * (blockSize + 1) + 1024;              // every symbol can cost an escape and a literal
    DBuf<u8> T(c, (size_t)B << SEG_SHIFT), U(c, (size_t)B << SEG_SHIFT);
    DBuf<u16> sym(c, (size_t)B << SEG_SHIFT);
    DBuf<u32> dn(c, B), dpidx(c, B), dm(c, B), dfreq(c, (size_t)B * HUFF_MAXSYM), dused(c, (size_t)B * 8), tcount(c, B);
    DBuf<u64> triples(c, (size_t)B * tcap);
    std::vector<u32> hn(B);
    for (size_t k0 = 0; k0 < nblocks; k0 += B) {
      const u32 nb = (u32)std::min<size_t>(B, nblocks - k0);
      for (u32 b = 0; b < nb; b++) hn[b] = (u32)std::min<size_t>(blockSize, n - (k0 + b) * blockSize);
      const u32 full = hn[nb - 1] == blockSize ? nb : nb - 1;   // only the last block of the file can be short
      if (full)
        CUDA_CHECK(cudaMemcpy2DAsync(T.p, SEG_SIZE, d_in + k0 * blockSize, blockSize, blockSize, full, cudaMemcpyDeviceToDevice, c.stream));
      if (full < nb)
        CUDA_CHECK(cudaMemcpyAsync(T.p + ((size_t)full << SEG_SHIFT), d_in + (k0 + full) * blockSize, hn[full], cudaMemcpyDeviceToDevice, c.stream));
      c.to_device(dn, hn.data(), 4 * nb);
      CUDA_CHECK(cudaMemsetAsync(dpidx, 0, 4 * nb, c.stream));
      {
        StageScope s(c, ST_BWT);
        bwt_forward_batch(c, T, U, dn, hn.data(), nb, dpidx, true, nullptr);   // U = sentinel BWT, dpidx = pidx + 1
      }
      {
        StageScope s(c, ST_MTF);
        mtf_rle2_batch(c, T, U, dn, hn.data(), nb, sym, dm, dfreq, dused);
      }
      {
        StageScope s(c, ST_HUFF);   // statistics: the model takes the slot of the Huffman stage (ms_huff) ...
        if (fast) k_bwtc_model<<<nb, 32, 0, c.stream>>>(sym, dm, dn, dpidx, dused, blockSize, fast, triples, tcap, tcount);
        else k_bwtc_model_fenwick<<<nb, 32, 0, c.stream>>>(sym, dm, dn, dpidx, dused, blockSize, triples, tcap, tcount);
        KLAUNCH(c); KCHECK();
      }
      {
        StageScope s(c, ST_PACK);   // ... and the serial range coder the slot of the bit packer (ms_pack)
        k_bwtc_code<<<1, 32, 0, c.stream>>>(st, triples, tcount, nb, tcap);
        KLAUNCH(c); KCHECK();
      }
      c.stats.blocks += nb;
    }
  }
  k_bwtc_finish<<<1, 1, 0, c.stream>>>(st);
  KLAUNCH(c); KCHECK();
  BwtcState h;
  c.to_host(&h, st, sizeof h);
  c.sync();
  if (h.overflow) throw B2Error{B2_ERR_CUDA, "internal error: BWTC triple buffer overflow"};
  if (h.rc.n > h.rc.cap) throw B2Error{B2_ERR_BAD_ARG, "output buffer too small for the compressed stream"};
  *out_n = hlen + (size_t)h.rc.n;
}

// ---- decode ---------------------------------------------------------------------------------------------------
struct BwtcDecResult {
  int status;       // 0 ok, negative: the stream is nonsense
  u32 nblocks;
  u32 level;
  u32 pad;
};

// one thread: the whole stream down to the L columns (inverse MTF folded in), block b at slot b << 20
__global__ void k_bwtc_decode(const u8* __restrict__ in, u64 n, u64 pos, u32 maxblocks, u8* __restrict__ L, u32* __restrict__ lengths,
                              u32* __restrict__ pidx1, BwtcDecResult* res) {
  if (threadIdx.x || blockIdx.x) return;
  bc_dec rc;
  bc_dec_start(&rc, in, n, pos);                                // lib/BWTC.js:142-143
  const u32 level = bc_dec_cul(&rc, 256);                       // decoder.decodeByte(), :144
  bc_dec_update(&rc, 1, level, 256);
  res->level = level;
  res->nblocks = 0;
  if (level < 1 || level > 9) { res->status = B2_ERR_DATA_ERROR; return; }
  bc_model model;
  u32 nb = 0;
  int r = 0;
  for (;;) {
    if (nb >= maxblocks) {
      // only "no more blocks" may follow
      const u32 ind = bc_dec_cul(&rc, 3);
      r = ind == 2 ? 1 : B2_ERR_DATA_ERROR;
      break;
    }
    u32 len = 0, p1 = 0;
    r = bc_decode_block(&rc, &model, level * 100000u, level <= 5, L + ((size_t)nb << SEG_SHIFT), &len, &p1);
    if (r) break;
    lengths[nb] = len; pidx1[nb] = p1;
    nb++;
  }
  res->nblocks = nb;
  res->status = r == 1 ? 0 : B2_ERR_DATA_ERROR;
}

// BWTC.decompressFile on device buffers.  h_head = the first bytes of the stream on the host (>= 16 or all of it).
void bwtc_decompress_device(Ctx& c, const u8* d_in, size_t n, const u8* h_head, size_t head_n, u8* d_out, size_t out_cap, size_t* out_n) {
  if (head_n < 5 || h_head[0] != 'b' || h_head[1] != 'w' || h_head[2] != 't' || h_head[3] != 'c')
    throw B2Error{B2_ERR_BAD_MAGIC, "Bad magic"};             // lib/Util.js:151-153
  size_t pos = 4;
  u64 fs = 0;
  for (;;) {                                                   // lib/Util.js:211-220 readUnsignedNumber
    // nine 7-bit groups hold any size below 2^63; a longer number cannot be the size of a real file and would overflow
    if (pos >= head_n || pos > 4 + 9) throw B2Error{B2_ERR_DATA_ERROR, "truncated or oversized BWTC header"};
    const u32 ch = h_head[pos++];
    if (ch & 0x80) { fs += ch & 0x7F; break; }
    fs = (fs + ch) * 128;
  }
  if (fs == 0) throw B2Error{B2_ERR_BAD_ARG, "BWTC streams of unknown size are not supported"};
  const u64 size = fs - 1;
  *out_n = (size_t)size;
  if (size > out_cap) throw B2Error{B2_ERR_BAD_ARG, "output buffer too small"};
  // the level is inside the coded stream: size the buffers for the smallest block size
  // (the header is not trusted: a block costs at least a few coded bytes, so n compressed bytes cannot hold more than
  // n / 2 blocks, and the decoded size must be something this GPU can hold)
  if (size > ((u64)1 << 40)) throw B2Error{B2_ERR_DATA_ERROR, "Data error: implausible BWTC size field"};
  const u64 mb64 = std::min<u64>(size / 100000u + 2, (u64)n / 2 + 2);
  const u32 maxblocks = (u32)mb64;
  DBuf<u8> L(c, (size_t)maxblocks << SEG_SHIFT);
  DBuf<u32> lengths(c, maxblocks), pidx1(c, maxblocks);
  DBuf<BwtcDecResult> res(c, 1);
  {
    StageScope s(c, ST_HDEC);
    k_bwtc_decode<<<1, 1, 0, c.stream>>>(d_in, n, pos, maxblocks, L, lengths, pidx1, res);
    KLAUNCH(c); KCHECK();
  }
  BwtcDecResult h;
  c.to_host(&h, res, sizeof h);
  c.sync();
  if (h.status) throw B2Error{B2_ERR_DATA_ERROR, "Data error: BWTC stream is corrupt"};
  std::vector<u32> hl(h.nblocks), hp(h.nblocks);
  if (h.nblocks) {
    c.to_host(hl.data(), lengths, 4 * h.nblocks);
    c.to_host(hp.data(), pidx1, 4 * h.nblocks);
    c.sync();
  }
  u64 total = 0;
  for (u32 b = 0; b < h.nblocks; b++) total += hl[b];
  if (total != size) throw B2Error{B2_ERR_DATA_ERROR, "outputsize does not match decoded input"};   // lib/Util.js:69-71
  u64 off = 0;
  StageScope s(c, ST_IBWT);
  for (u32 b = 0; b < h.nblocks; b++) {                         // BWT.unbwtransform per block, lib/BWTC.js:224
    const u8* Lb = L.p + ((size_t)b << SEG_SHIFT);
    if (hl[b] == 1) CUDA_CHECK(cudaMemcpyAsync(d_out + off, Lb, 1, cudaMemcpyDeviceToDevice, c.stream));
    else bwt_inverse_sentinel(c, Lb, hl[b], hp[b], d_out + off);
    off += hl[b];
  }
  c.stats.blocks += h.nblocks;
}
