// tc_path.cu — bf16 tensor-core path (placeholder until the tcgen05 kernels land).
#include "ctx.h"
int tc_create(lstm_ctx* ctx) { return lstm_fail(ctx, LSTM_ERR_UNSUPPORTED, "LSTM_BF16 path not built yet"); }
void tc_destroy(lstm_ctx*) {}
int tc_params_changed(lstm_ctx*) { return LSTM_OK; }
int tc_forward(lstm_ctx* ctx) { return lstm_fail(ctx, LSTM_ERR_UNSUPPORTED, "LSTM_BF16 path not built yet"); }
int tc_backward(lstm_ctx* ctx) { return lstm_fail(ctx, LSTM_ERR_UNSUPPORTED, "LSTM_BF16 path not built yet"); }
int tc_get_activation(lstm_ctx* ctx, int, int, float*, size_t) { return lstm_fail(ctx, LSTM_ERR_UNSUPPORTED, "LSTM_BF16 path not built yet"); }
