// placeholder until the render loop lands
#include "kdpt_internal.h"
using namespace kdpt;
extern "C" {
int kdpt_render(kdpt_ctx*, int, int, int, const kdpt_params*, kdpt_render_stats*) { setError("kdpt_render: not built"); return KDPT_ERR_STATE; }
int kdpt_clear_image(kdpt_ctx*) { setError("not built"); return KDPT_ERR_STATE; }
int kdpt_get_image(kdpt_ctx*, float*) { setError("not built"); return KDPT_ERR_STATE; }
int kdpt_image_device(kdpt_ctx*, float**, int*) { setError("not built"); return KDPT_ERR_STATE; }
}
