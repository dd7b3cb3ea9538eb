/*
 * N-API shim over the C ABI of libntru_b200.so (include/ntru_b200.h).
 *
 * NOT COMPILED OR RUN IN THIS REPOSITORY'S IMAGE: node, npm and node_api.h are absent.
 * It is the binding a maintainer of numtel/ntru-circom would add next to index.js; build with
 *     gcc -shared -fPIC -I<node>/include/node -I../../include ntru_napi.c -L../../ntru-circom_b200 -lntru_b200 -o ntru_b200.node
 *
 * Surface (all synchronous, like the reference):
 *   create(N, p, q, device)                          -> external handle
 *   setPublicKey(h: Uint16Array)                     / setPrivateKey(f: Int8Array, fp: Uint8Array)
 *   encryptBatch(B, r: Uint8Array, m: Uint8Array)    -> {value, quotientE, remainderE}   (Uint16Array, fixed length)
 *   decryptBatch(B, e: Uint16Array)                  -> {value, quotient1, remainder1, quotient2, remainder2}
 *   sum(B, e: Uint16Array)                           -> Uint16Array(N)
 * Trimming / expandArray / {value, inputs, params} assembly stay in JavaScript (bindings/node/index.mjs).
 */
#include <node_api.h>
#include <stdint.h>
#include <stdlib.h>

#include "ntru_b200.h"

#define CHECK(env, rc, ctx)                                                         \
  do {                                                                              \
    if ((rc) != NTRU_OK) {                                                          \
      napi_throw_error((env), NULL, (ctx) ? ntru_last_error(ctx) : ntru_strerror(rc)); \
      return NULL;                                                                  \
    }                                                                               \
  } while (0)

static void finalize_ctx(napi_env env, void *data, void *hint) { (void)env; (void)hint; ntru_destroy((ntru_ctx *)data); }

static ntru_ctx *get_ctx(napi_env env, napi_value v) {
  void *p = NULL;
  napi_get_value_external(env, v, &p);
  return (ntru_ctx *)p;
}

static void *typed_data(napi_env env, napi_value v, size_t *len) {
  napi_typedarray_type t; void *data; napi_value ab; size_t off;
  napi_get_typedarray_info(env, v, &t, len, &data, &ab, &off);
  return data;
}

static napi_value make_typed(napi_env env, napi_typedarray_type t, size_t count, size_t elem, void **data) {
  napi_value ab, ta;
  napi_create_arraybuffer(env, count * elem, data, &ab);
  napi_create_typedarray(env, t, count, ab, 0, &ta);
  return ta;
}

static napi_value Create(napi_env env, napi_callback_info info) {
  size_t argc = 4; napi_value argv[4]; int32_t a[4];
  napi_get_cb_info(env, info, &argc, argv, NULL, NULL);
  for (int i = 0; i < 4; ++i) napi_get_value_int32(env, argv[i], &a[i]);
  ntru_ctx *ctx = NULL;
  int rc = ntru_create(&ctx, a[0], a[1], a[2], a[3]);
  CHECK(env, rc, NULL);
  napi_value ext;
  napi_create_external(env, ctx, finalize_ctx, NULL, &ext);
  return ext;
}

static napi_value SetPublicKey(napi_env env, napi_callback_info info) {
  size_t argc = 2, len; napi_value argv[2];
  napi_get_cb_info(env, info, &argc, argv, NULL, NULL);
  ntru_ctx *ctx = get_ctx(env, argv[0]);
  CHECK(env, ntru_set_public_key(ctx, (const uint16_t *)typed_data(env, argv[1], &len)), ctx);
  return NULL;
}

static napi_value SetPrivateKey(napi_env env, napi_callback_info info) {
  size_t argc = 3, len; napi_value argv[3];
  napi_get_cb_info(env, info, &argc, argv, NULL, NULL);
  ntru_ctx *ctx = get_ctx(env, argv[0]);
  CHECK(env, ntru_set_private_key(ctx, (const int8_t *)typed_data(env, argv[1], &len),
                                  (const uint8_t *)typed_data(env, argv[2], &len)), ctx);
  return NULL;
}

static napi_value EncryptBatch(napi_env env, napi_callback_info info) {
  size_t argc = 5, len; napi_value argv[5]; uint32_t B, N;
  napi_get_cb_info(env, info, &argc, argv, NULL, NULL);
  ntru_ctx *ctx = get_ctx(env, argv[0]);
  napi_get_value_uint32(env, argv[1], &B);
  napi_get_value_uint32(env, argv[2], &N);
  void *value, *quo, *rem;
  napi_value out, v0 = make_typed(env, napi_uint16_array, (size_t)B * N, 2, &value),
                  v1 = make_typed(env, napi_uint16_array, (size_t)B * (N + 1), 2, &quo),
                  v2 = make_typed(env, napi_uint16_array, (size_t)B * (N + 1), 2, &rem);
  CHECK(env, ntru_encrypt_batch(ctx, B, (const uint8_t *)typed_data(env, argv[3], &len),
                                (const uint8_t *)typed_data(env, argv[4], &len), value, quo, rem), ctx);
  napi_create_object(env, &out);
  napi_set_named_property(env, out, "value", v0);
  napi_set_named_property(env, out, "quotientE", v1);
  napi_set_named_property(env, out, "remainderE", v2);
  return out;
}

static napi_value DecryptBatch(napi_env env, napi_callback_info info) {
  size_t argc = 4, len; napi_value argv[4]; uint32_t B, N;
  napi_get_cb_info(env, info, &argc, argv, NULL, NULL);
  ntru_ctx *ctx = get_ctx(env, argv[0]);
  napi_get_value_uint32(env, argv[1], &B);
  napi_get_value_uint32(env, argv[2], &N);
  void *value, *q1, *r1, *q2, *r2;
  napi_value out, v0 = make_typed(env, napi_uint8_array, (size_t)B * N, 1, &value),
                  v1 = make_typed(env, napi_uint16_array, (size_t)B * (N + 1), 2, &q1),
                  v2 = make_typed(env, napi_uint16_array, (size_t)B * (N + 1), 2, &r1),
                  v3 = make_typed(env, napi_uint8_array, (size_t)B * (N + 1), 1, &q2),
                  v4 = make_typed(env, napi_uint8_array, (size_t)B * (N + 1), 1, &r2);
  CHECK(env, ntru_decrypt_batch(ctx, B, (const uint16_t *)typed_data(env, argv[3], &len), value, q1, r1, q2, r2), ctx);
  napi_create_object(env, &out);
  napi_set_named_property(env, out, "value", v0);
  napi_set_named_property(env, out, "quotient1", v1);
  napi_set_named_property(env, out, "remainder1", v2);
  napi_set_named_property(env, out, "quotient2", v3);
  napi_set_named_property(env, out, "remainder2", v4);
  return out;
}

static napi_value Sum(napi_env env, napi_callback_info info) {
  size_t argc = 4, len; napi_value argv[4]; uint32_t B, N;
  napi_get_cb_info(env, info, &argc, argv, NULL, NULL);
  ntru_ctx *ctx = get_ctx(env, argv[0]);
  napi_get_value_uint32(env, argv[1], &B);
  napi_get_value_uint32(env, argv[2], &N);
  void *out_data;
  napi_value out = make_typed(env, napi_uint16_array, N, 2, &out_data);
  CHECK(env, ntru_sum(ctx, B, (const uint16_t *)typed_data(env, argv[3], &len), out_data), ctx);
  return out;
}

/* verifyKeysBatch(ctx, B, N, f:Int8Array, fq:Uint16Array, fp:Uint8Array, g:Int8Array) -- index.js:141-197 for B keys */
static napi_value VerifyKeysBatch(napi_env env, napi_callback_info info) {
  size_t argc = 7, len; napi_value argv[7]; uint32_t B, N;
  napi_get_cb_info(env, info, &argc, argv, NULL, NULL);
  ntru_ctx *ctx = get_ctx(env, argv[0]);
  napi_get_value_uint32(env, argv[1], &B);
  napi_get_value_uint32(env, argv[2], &N);
  void *o[6];
  static const char *names[6] = {"quotientFq", "remainderFq", "quotientFp", "remainderFp", "quotientH", "remainderH"};
  napi_value out, v[6];
  for (int i = 0; i < 6; ++i) {
    const int bytes = (i == 2 || i == 3) ? 1 : 2;
    v[i] = make_typed(env, bytes == 1 ? napi_uint8_array : napi_uint16_array, (size_t)B * (N + 1), bytes, &o[i]);
  }
  CHECK(env, ntru_verify_keys_batch(ctx, B, (const int8_t *)typed_data(env, argv[3], &len), (const uint16_t *)typed_data(env, argv[4], &len),
                                    (const uint8_t *)typed_data(env, argv[5], &len), (const int8_t *)typed_data(env, argv[6], &len),
                                    o[0], o[1], o[2], o[3], o[4], o[5]), ctx);
  napi_create_object(env, &out);
  for (int i = 0; i < 6; ++i) napi_set_named_property(env, out, names[i], v[i]);
  return out;
}

/* keygenBatch(ctx, B, N, f:Int8Array, g:Int8Array) -> {fq, fp, h, valid} -- index.js:30-79 for B keys */
static napi_value KeygenBatch(napi_env env, napi_callback_info info) {
  size_t argc = 5, len; napi_value argv[5]; uint32_t B, N;
  napi_get_cb_info(env, info, &argc, argv, NULL, NULL);
  ntru_ctx *ctx = get_ctx(env, argv[0]);
  napi_get_value_uint32(env, argv[1], &B);
  napi_get_value_uint32(env, argv[2], &N);
  void *fq, *fp, *h, *valid;
  napi_value out, v0 = make_typed(env, napi_uint16_array, (size_t)B * N, 2, &fq), v1 = make_typed(env, napi_uint8_array, (size_t)B * N, 1, &fp),
                  v2 = make_typed(env, napi_uint16_array, (size_t)B * N, 2, &h), v3 = make_typed(env, napi_uint8_array, B, 1, &valid);
  CHECK(env, ntru_keygen_batch(ctx, B, (const int8_t *)typed_data(env, argv[3], &len), (const int8_t *)typed_data(env, argv[4], &len),
                               fq, fp, h, valid), ctx);
  napi_create_object(env, &out);
  napi_set_named_property(env, out, "fq", v0);
  napi_set_named_property(env, out, "fp", v1);
  napi_set_named_property(env, out, "h", v2);
  napi_set_named_property(env, out, "valid", v3);
  return out;
}

static napi_value Init(napi_env env, napi_value exports) {
  napi_property_descriptor d[] = {
      {"create", 0, Create, 0, 0, 0, napi_default, 0},           {"setPublicKey", 0, SetPublicKey, 0, 0, 0, napi_default, 0},
      {"setPrivateKey", 0, SetPrivateKey, 0, 0, 0, napi_default, 0}, {"encryptBatch", 0, EncryptBatch, 0, 0, 0, napi_default, 0},
      {"decryptBatch", 0, DecryptBatch, 0, 0, 0, napi_default, 0},   {"sum", 0, Sum, 0, 0, 0, napi_default, 0},
      {"verifyKeysBatch", 0, VerifyKeysBatch, 0, 0, 0, napi_default, 0}, {"keygenBatch", 0, KeygenBatch, 0, 0, 0, napi_default, 0},
  };
  napi_define_properties(env, exports, sizeof d / sizeof d[0], d);
  return exports;
}

NAPI_MODULE(NODE_GYP_MODULE_NAME, Init)
