// addon.cc -- Node N-API veneer over include/b2bz.h (cannot be built in the graft image: no node
// headers).  It only marshals Buffers; every byte of work happens behind the C ABI.
// Build (where node-gyp exists): node-gyp configure build  with libraries: ["-lb2bz"].
#include <node_api.h>
#include <stdint.h>
#include "../../include/b2bz.h"

static void fin(napi_env, void* data, void*) { b2_free(data); }

static napi_value fail(napi_env env, int rc) {
  napi_value msg, err, code;
  napi_create_string_utf8(env, b2_last_error(), NAPI_AUTO_LENGTH, &msg);
  if (rc == B2_ERR_BAD_LEVEL) napi_create_error(env, nullptr, msg, &err);        // `new Error(...)` lib/Bzip2.js:888-890
  else napi_create_type_error(env, nullptr, msg, &err);                            // `new TypeError(...)` lib/Bzip2.js:82-88
  napi_create_int32(env, rc, &code);
  napi_set_named_property(env, err, "errorCode", code);
  napi_throw(env, err);
  return nullptr;
}

static bool buf_arg(napi_env env, napi_value v, const uint8_t** p, size_t* n) {
  void* d = nullptr;
  if (napi_get_buffer_info(env, v, &d, n) != napi_ok) return false;
  *p = (const uint8_t*)d;
  return true;
}

// compressFile(buffer, level) -> Buffer            (Bzip2.compressFile, lib/Bzip2.js:879)
static napi_value CompressFile(napi_env env, napi_callback_info info) {
  size_t argc = 2; napi_value argv[2];
  napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr);
  const uint8_t* in; size_t n; int32_t level = 9;
  if (!buf_arg(env, argv[0], &in, &n)) return fail(env, B2_ERR_BAD_ARG);
  napi_get_value_int32(env, argv[1], &level);
  uint8_t* out; size_t out_n;
  int rc = b2_bzip2_compress(in, n, level, &out, &out_n);
  if (rc) return fail(env, rc);
  napi_value buf; napi_create_external_buffer(env, out_n, out, fin, nullptr, &buf);
  return buf;
}
// decompressFile(buffer, multistream) -> Buffer    (Bunzip.decode, lib/Bzip2.js:454)
static napi_value DecompressFile(napi_env env, napi_callback_info info) {
  size_t argc = 2; napi_value argv[2];
  napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr);
  const uint8_t* in; size_t n; bool ms = false;
  if (!buf_arg(env, argv[0], &in, &n)) return fail(env, B2_ERR_BAD_ARG);
  if (argc > 1) napi_get_value_bool(env, argv[1], &ms);
  uint8_t* out; size_t out_n;
  int rc = b2_bzip2_decompress(in, n, ms ? 1 : 0, &out, &out_n);
  if (rc) return fail(env, rc);
  napi_value buf; napi_create_external_buffer(env, out_n, out, fin, nullptr, &buf);
  return buf;
}
// decompressBlock(buffer, bitpos) -> Buffer        (Bunzip.decodeBlock, lib/Bzip2.js:482)
static napi_value DecompressBlock(napi_env env, napi_callback_info info) {
  size_t argc = 2; napi_value argv[2];
  napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr);
  const uint8_t* in; size_t n; double pos = 0;
  if (!buf_arg(env, argv[0], &in, &n)) return fail(env, B2_ERR_BAD_ARG);
  napi_get_value_double(env, argv[1], &pos);
  uint8_t* out; size_t out_n;
  int rc = b2_bzip2_decompress_block(in, n, (uint64_t)pos, &out, &out_n);
  if (rc) return fail(env, rc);
  napi_value buf; napi_create_external_buffer(env, out_n, out, fin, nullptr, &buf);
  return buf;
}
// integer argument that must be a JS number; false (and B2_ERR_BAD_ARG thrown by the caller) otherwise
static bool int_arg(napi_env env, napi_value v, int32_t* out) {
  napi_valuetype ty;
  if (napi_typeof(env, v, &ty) != napi_ok || ty != napi_number) return false;
  return napi_get_value_int32(env, v, out) == napi_ok;
}
// n bytes must exist in both the source and the destination view
static bool len_ok(int32_t n, size_t a, size_t b) { return n >= 0 && (size_t)n <= a && (size_t)n <= b; }

// table(buffer, multistream) -> [[bitpos, size], ...]   (Bunzip.table, lib/Bzip2.js:508)
static napi_value Table(napi_env env, napi_callback_info info) {
  size_t argc = 2; napi_value argv[2];
  napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr);
  const uint8_t* in; size_t n; bool ms = false;
  if (!buf_arg(env, argv[0], &in, &n)) return fail(env, B2_ERR_BAD_ARG);
  if (argc > 1) napi_get_value_bool(env, argv[1], &ms);
  uint64_t* bp; uint32_t* sz; size_t cnt;
  int rc = b2_bzip2_table(in, n, ms ? 1 : 0, &bp, &sz, &cnt);
  if (rc) return fail(env, rc);
  napi_value arr; napi_create_array_with_length(env, cnt, &arr);
  for (size_t i = 0; i < cnt; i++) {
    napi_value row, a, b;
    napi_create_array_with_length(env, 2, &row);
    napi_create_double(env, (double)bp[i], &a); napi_create_uint32(env, sz[i], &b);
    napi_set_element(env, row, 0, a); napi_set_element(env, row, 1, b);
    napi_set_element(env, arr, (uint32_t)i, row);
  }
  b2_free(bp); b2_free(sz);
  return arr;
}
// bwtransform2(T, U, n) -> pidx                     (BWT.bwtransform2, lib/BWT.js:372)
static napi_value Bwtransform2(napi_env env, napi_callback_info info) {
  size_t argc = 3; napi_value argv[3];
  napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr);
  const uint8_t *T, *U; size_t tn, un; int32_t n = 0;
  if (!buf_arg(env, argv[0], &T, &tn) || !buf_arg(env, argv[1], &U, &un)) return fail(env, B2_ERR_BAD_ARG);
  if (argc < 3 || !int_arg(env, argv[2], &n) || !len_ok(n, tn, un)) return fail(env, B2_ERR_BAD_ARG);
  int32_t p = b2_bwt_cyclic(T, (uint8_t*)U, n);
  if (p < 0) return fail(env, p);
  napi_value r; napi_create_int32(env, p, &r);
  return r;
}

// bwtcCompressFile(buffer, level) -> Buffer        (BWTC.compressFile, lib/BWTC.js:12)
static napi_value BwtcCompressFile(napi_env env, napi_callback_info info) {
  size_t argc = 2; napi_value argv[2];
  napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr);
  const uint8_t* in; size_t n; int32_t level = 9;
  if (!buf_arg(env, argv[0], &in, &n)) return fail(env, B2_ERR_BAD_ARG);
  if (argc > 1) napi_get_value_int32(env, argv[1], &level);
  uint8_t* out; size_t out_n;
  int rc = b2_bwtc_compress(in, n, level, &out, &out_n);
  if (rc) return fail(env, rc);
  napi_value buf; napi_create_external_buffer(env, out_n, out, fin, nullptr, &buf);
  return buf;
}
// bwtcDecompressFile(buffer) -> Buffer              (BWTC.decompressFile, lib/BWTC.js:141)
static napi_value BwtcDecompressFile(napi_env env, napi_callback_info info) {
  size_t argc = 1; napi_value argv[1];
  napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr);
  const uint8_t* in; size_t n;
  if (!buf_arg(env, argv[0], &in, &n)) return fail(env, B2_ERR_BAD_ARG);
  uint8_t* out; size_t out_n;
  int rc = b2_bwtc_decompress(in, n, &out, &out_n);
  if (rc == B2_ERR_BAD_MAGIC) { napi_throw_error(env, nullptr, "Bad magic"); return nullptr; }   // lib/Util.js:151-153
  if (rc) return fail(env, rc);
  napi_value buf; napi_create_external_buffer(env, out_n, out, fin, nullptr, &buf);
  return buf;
}
// suffixsort(T, SA /* Int32Array */, n) -> 0        (BWT.suffixsort, lib/BWT.js:305)
static napi_value Suffixsort(napi_env env, napi_callback_info info) {
  size_t argc = 3; napi_value argv[3];
  napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr);
  const uint8_t* T; size_t tn; int32_t n = 0;
  napi_typedarray_type ty; size_t len; void* sa; napi_value ab; size_t off;
  if (!buf_arg(env, argv[0], &T, &tn)) return fail(env, B2_ERR_BAD_ARG);
  if (napi_get_typedarray_info(env, argv[1], &ty, &len, &sa, &ab, &off) != napi_ok || ty != napi_int32_array) return fail(env, B2_ERR_BAD_ARG);
  if (argc < 3 || !int_arg(env, argv[2], &n) || !len_ok(n, tn, len)) return fail(env, B2_ERR_BAD_ARG);
  int rc = b2_suffixsort(T, (int32_t*)sa, n);
  if (rc < 0) return fail(env, rc);
  napi_value r; napi_create_int32(env, 0, &r);
  return r;
}
// bwtransform(T, U, n) -> pidx + 1                   (BWT.bwtransform, lib/BWT.js:328; A is scratch, dropped)
static napi_value Bwtransform(napi_env env, napi_callback_info info) {
  size_t argc = 3; napi_value argv[3];
  napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr);
  const uint8_t *T, *U; size_t tn, un; int32_t n = 0;
  if (!buf_arg(env, argv[0], &T, &tn) || !buf_arg(env, argv[1], &U, &un)) return fail(env, B2_ERR_BAD_ARG);
  if (argc < 3 || !int_arg(env, argv[2], &n) || !len_ok(n, tn, un)) return fail(env, B2_ERR_BAD_ARG);
  int32_t p = b2_bwt_sentinel(T, (uint8_t*)U, n);
  if (p < 0) return fail(env, p);
  napi_value r; napi_create_int32(env, p, &r);
  return r;
}
// unbwtransform(T, U, n, pidx)                       (BWT.unbwtransform, lib/BWT.js:352; LF is scratch, dropped)
static napi_value Unbwtransform(napi_env env, napi_callback_info info) {
  size_t argc = 4; napi_value argv[4];
  napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr);
  const uint8_t *T, *U; size_t tn, un; int32_t n = 0, pidx = 0;
  if (!buf_arg(env, argv[0], &T, &tn) || !buf_arg(env, argv[1], &U, &un)) return fail(env, B2_ERR_BAD_ARG);
  if (argc < 4 || !int_arg(env, argv[2], &n) || !int_arg(env, argv[3], &pidx) || !len_ok(n, tn, un) || pidx < 0 || pidx > n)
    return fail(env, B2_ERR_BAD_ARG);
  int rc = b2_bwt_inverse(T, (uint8_t*)U, n, pidx);
  if (rc < 0) return fail(env, rc);
  return nullptr;
}

static napi_value Init(napi_env env, napi_value exports) {
  napi_property_descriptor d[] = {
      {"compressFile", nullptr, CompressFile, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"decompressFile", nullptr, DecompressFile, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"decompressBlock", nullptr, DecompressBlock, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"table", nullptr, Table, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"bwtransform2", nullptr, Bwtransform2, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"suffixsort", nullptr, Suffixsort, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"bwtransform", nullptr, Bwtransform, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"unbwtransform", nullptr, Unbwtransform, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"bwtcCompressFile", nullptr, BwtcCompressFile, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"bwtcDecompressFile", nullptr, BwtcDecompressFile, nullptr, nullptr, nullptr, napi_default, nullptr},
  };
  napi_define_properties(env, exports, sizeof d / sizeof d[0], d);
  return exports;
}
NAPI_MODULE(NODE_GYP_MODULE_NAME, Init)
