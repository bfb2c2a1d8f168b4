// obj_loader.cpp — host-side OBJ/MTL loader (Object::load_from_file, object.rs:106-454).
// Placeholder until the loader row is built: the entry points exist so the
// ABI is complete, and fail loudly.
#include <cstddef>
#include <cstdint>

#include "../../include/draw_b200.h"

namespace drawb200 {
int loader_fail(int code, const char *msg);
}

extern "C" {
int draw_object_load_obj(const char *, draw_image_loader, void *, draw_object **out) {
    if (out) *out = nullptr;
    return drawb200::loader_fail(DRAW_ERR_INTERNAL, "draw_object_load_obj: loader not built yet");
}
void draw_object_free(draw_object *) {}
int draw_object_desc_of(const draw_object *, draw_object_desc *) {
    return drawb200::loader_fail(DRAW_ERR_INTERNAL, "draw_object_desc_of: loader not built yet");
}
}
