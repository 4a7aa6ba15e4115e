// placeholder until the wavefront pipeline lands
#include "context.h"
namespace vg {
struct RenderState {};
int render_run(vg_ctx* ctx, int, int, float*) { return ctx->fail(VG_ERR_UNSUPPORTED, "vg_render: not built yet"); }
int render_clear(vg_ctx*) { return VG_OK; }
int render_fb_device(vg_ctx* ctx, float**) { return ctx->fail(VG_ERR_UNSUPPORTED, "not built yet"); }
void render_invalidate(vg_ctx*) {}
void render_destroy(vg_ctx*) {}
}
