// Fused log-mel + CMN front-end -- placeholder until the fused kernel lands.
#include "common.cuh"

namespace sdb {
int fbank_launch(sd_ctx* ctx, const float*, int, int, const float*, const sd_fbank_params*, float*) {
    return ctx->fail(SD_ERR_UNSUPPORTED, "sd_fbank: fused fbank kernel not built yet");
}
}  // namespace sdb
