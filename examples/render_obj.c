/* The call sequence of the reference's Application::run (mororo18/draw src/app/mod.rs:120-210) through the C ABI:
 *   load an OBJ/MTL model, add it to a scene, render a few frames with a moving camera, read the frame back and
 *   export it.   cc -I include examples/render_obj.c -L draw_b200 -ldraw_b200 -Wl,-rpath,$PWD/draw_b200 -o render_obj
 *   ./render_obj models/lemur/lemur.obj out.png [width height]                                                     */
#include <stdio.h>
#include <stdlib.h>

#include "draw_b200.h"

#define CHECK(call)                                                                   \
    do {                                                                              \
        if ((call) != DRAW_OK) {                                                      \
            fprintf(stderr, "%s failed: %s\n", #call, draw_last_error());             \
            return 1;                                                                 \
        }                                                                             \
    } while (0)

int main(int argc, char **argv) {
    if (argc < 3) {
        fprintf(stderr, "usage: %s model.obj out.png [width height]\n", argv[0]);
        return 2;
    }
    const size_t width = argc > 4 ? (size_t)atoi(argv[3]) : 800, height = argc > 4 ? (size_t)atoi(argv[4]) : 600;

    draw_object *object = NULL; /* Object::load_from_file, object.rs:106; textures decoded by the library (PNG, JPEG) */
    CHECK(draw_object_load_obj(argv[1], draw_image_loader_builtin, NULL, &object));
    draw_object_desc desc;
    CHECK(draw_object_desc_of(object, &desc));

    draw_scene *scene = NULL;   /* Scene::new, scene/mod.rs:760 */
    draw_canvas *canvas = NULL; /* Canvas::new, canvas.rs:366 */
    CHECK(draw_scene_create(width, height, &scene));
    CHECK(draw_canvas_create(width, height, &canvas));
    CHECK(draw_canvas_init_depth(canvas, 100000.0f)); /* src/app/mod.rs:126 */
    CHECK(draw_scene_add_object(scene, &desc, NULL));  /* add_obj, scene/mod.rs:788: the geometry is copied to the device */
    draw_object_free(object);
    CHECK(draw_canvas_enable_host_mirror(canvas, 1));  /* the frame follows each render to host memory */

    const uint8_t *frame = NULL;
    size_t frame_bytes = 0;
    for (int k = 0; k < 8; k++) { /* the event loop's body, src/app/mod.rs:196-202 */
        const float pos[3] = {40.0f * (float)k - 140.0f, 20.0f, 260.0f}, dir[3] = {-pos[0], -pos[1], -pos[2]};
        CHECK(draw_scene_set_camera(scene, pos, dir));            /* scene.camera = Camera::new(..) */
        CHECK(draw_scene_render(scene, canvas));                  /* Scene::render: enqueues and returns */
        CHECK(draw_canvas_map_host(canvas, &frame, &frame_bytes)); /* Canvas::as_bytes_slice: B,G,R,pad, row 0 = top */
    }
    draw_frame_stats stats;
    CHECK(draw_canvas_last_frame_stats(canvas, &stats));
    printf("%zu x %zu, %u triangles in, %u records, %u tile references, %u KiB to the host for the last frame; first pixel %u %u %u\n", width,
           height, stats.input_triangles, stats.setup_records, stats.tile_refs, stats.mirror_kbytes, frame[2], frame[1], frame[0]);
    CHECK(draw_canvas_export_png(canvas, argv[2])); /* Application::export_frame_as(Png), src/app/mod.rs:316 */
    draw_canvas_destroy(canvas);
    draw_scene_destroy(scene);
    return 0;
}
