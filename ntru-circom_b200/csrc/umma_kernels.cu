// placeholder: tcgen05 schedule not built yet
#include "ntru_internal.cuh"
namespace ntru {
int umma_init(ntru_ctx *ctx) { ctx->tensor_ok = false; return NTRU_OK; }
int umma_prepare_public(ntru_ctx *) { return NTRU_OK; }
int umma_prepare_private(ntru_ctx *) { return NTRU_OK; }
int umma_encrypt(ntru_ctx *ctx, size_t, const uint8_t *, const uint8_t *, uint16_t *, uint16_t *, uint16_t *) {
  return fail(ctx, NTRU_E_UNSUPPORTED, "tensor schedule not built");
}
int umma_decrypt(ntru_ctx *ctx, size_t, const uint16_t *, uint8_t *, uint16_t *, uint16_t *, uint8_t *, uint8_t *) {
  return fail(ctx, NTRU_E_UNSUPPORTED, "tensor schedule not built");
}
}  // namespace ntru
