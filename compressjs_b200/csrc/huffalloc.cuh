// huffalloc.cuh -- length-limited canonical Huffman code length allocation, in place on a
// sorted (ascending) frequency array.  Runs on one GPU thread per table (the tables of many
// bzip2 blocks are built concurrently); also compilable for the host so that the unit tests can
// check it against the reference's known-answer vectors without a GPU.
//
// Reference: lib/HuffmanAllocator.js:52-222 (jbzip2's Moffat-Katajainen in-place allocator with
// Milidiu/Pessoa/Laber node relocation).  Lengths must be bit-identical to it -- any other
// optimal length assignment would change the .bz2 bytes.
#pragma once
#ifdef __CUDACC__
#define HA_FN __host__ __device__ __forceinline__
#else
#define HA_FN static inline
#endif

// Every entry first() looks at is an extended parent pointer in [0, 2*len) (pass 1 has converted all
// of a[0..len-3]), so `x % len` is one conditional subtract -- no integer division on the GPU.
HA_FN int ha_mod(int x, int len) { return x >= len ? x - len : x; }

// HuffmanAllocator.js:52-75
HA_FN int ha_first(const int* a, int len, int i, int nodesToMove) {
  const int limit = i;
  int k = len - 2;
  while (i >= nodesToMove && ha_mod(a[i], len) > limit) {
    k = i;
    i -= (limit - i + 1);
  }
  if (nodesToMove - 1 > i) i = nodesToMove - 1;
  while (k > i + 1) {
    const int t = (i + k) >> 1;
    if (ha_mod(a[t], len) > limit) k = t; else i = t;
  }
  return k;
}

// HuffmanAllocator.js:199-222 (with :79-105, :114-124, :131-148, :157-188 inlined below)
HA_FN void ha_allocate(int* a, int len, int maxLen) {
  if (len == 2) { a[1] = 1; a[0] = 1; return; }
  if (len == 1) { a[0] = 1; return; }
  if (len <= 0) return;
  // pass 1: extended parent pointers (:79-105)
  a[0] += a[1];
  {
    int head = 0, top = 2, temp;
    for (int tail = 1; tail < len - 1; tail++) {
      if (top >= len || a[head] < a[top]) { temp = a[head]; a[head++] = tail; }
      else temp = a[top++];
      if (top >= len || (head < tail && a[head] < a[top])) { temp += a[head]; a[head++] = tail + len; }
      else temp += a[top++];
      a[tail] = temp;
    }
  }
  // pass 2: nodes to relocate (:114-124)
  int nodesToRelocate = len - 2;
  for (int depth = 1; depth < maxLen - 1 && nodesToRelocate > 1; depth++)
    nodesToRelocate = ha_first(a, len, nodesToRelocate - 1, 0);
  // pass 3
  if (ha_mod(a[0], len) >= nodesToRelocate) {
    // :131-148
    int firstNode = len - 2, nextNode = len - 1;
    for (int depth = 1, avail = 2; avail > 0; depth++) {
      const int lastNode = firstNode;
      firstNode = ha_first(a, len, lastNode - 1, 0);
      for (int i = avail - (lastNode - firstNode); i > 0; i--) a[nextNode--] = depth;
      avail = (lastNode - firstNode) << 1;
    }
  } else {
    // :213-221 + :157-188
    int fl = 0;
    for (unsigned v = (unsigned)(nodesToRelocate - 1); v; v >>= 1) fl++;
    const int insertDepth = maxLen - fl;
    const int nodesToMove = nodesToRelocate;
    int firstNode = len - 2, nextNode = len - 1;
    int depth = (insertDepth == 1) ? 2 : 1;
    int left = (insertDepth == 1) ? nodesToMove - 2 : nodesToMove;
    for (int avail = depth << 1; avail > 0; depth++) {
      const int lastNode = firstNode;
      firstNode = (firstNode <= nodesToMove) ? firstNode : ha_first(a, len, lastNode - 1, nodesToMove);
      int offset = 0;
      if (depth >= insertDepth) {
        const int cap = 1 << (depth - insertDepth);
        offset = left < cap ? left : cap;
      } else if (depth == insertDepth - 1) {
        offset = 1;
        if (a[firstNode] == lastNode) firstNode++;
      }
      for (int i = avail - (lastNode - firstNode + offset); i > 0; i--) a[nextNode--] = depth;
      left -= offset;
      avail = (lastNode - firstNode + offset) << 1;
    }
  }
}
