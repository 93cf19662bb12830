// Tiled pair kernel (placeholder until the shared-memory variant lands): forwards to the direct kernel.
#include "sphgpu_internal.h"

namespace sph {

int launchPairTiled(sphgpu_ctx* ctx) {
    const int saved = ctx->variant;
    ctx->variant = 1;
    const int rc = launchPair(ctx);
    ctx->variant = saved;
    return rc;
}

} // namespace sph
