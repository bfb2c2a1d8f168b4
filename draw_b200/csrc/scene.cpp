// scene.cpp — host side of libdraw_b200.so: the C ABI of include/draw_b200.h.
//
// Mirrors the reference's Scene / Canvas / Camera verbs (mororo18/draw src/renderer/scene/mod.rs,
// canvas.rs) and owns the device memory: scene geometry as SoA arrays uploaded once per
// add_object, per-frame work buffers, and the canvas' colour + depth buffers with a pinned
// host mirror.  Per-frame uniforms are computed here in the reference's float32 operation
// order (host_math.hpp) and handed to the kernels by value.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <new>
#include <set>
#include <string>
#include <vector>

#include "../../include/draw_b200.h"
#include "device_types.h"
#include "host_math.hpp"

namespace drawb200 {
extern thread_local unsigned g_front_ctas;  // k_front.cu
extern thread_local unsigned g_raster_ctas; // k_raster.cu
extern thread_local unsigned g_tile_ctas;   // k_tile.cu
// k_front.cu / k_raster.cu / k_tile.cu / k_sort.cu
cudaError_t launch_front(const FrameUniforms *dU, const SceneDev &S, const FrameDev &W, cudaStream_t stream);
int front_max_ctas_per_sm();
uint32_t tile_grid_items(const FrameUniforms &U);
void launch_raster(const FrameUniforms *dU, const FrameDev &W, cudaStream_t stream);
cudaError_t launch_fill_u64(unsigned long long *dst, size_t n, unsigned long long value, cudaStream_t stream);
void launch_tile(const FrameUniforms &U, const FrameUniforms *dU, const SceneDev &S, const FrameDev &W, cudaStream_t stream);
void launch_sort_transparent(const FrameUniforms *dU, const SceneDev &S, const void *ranges, uint32_t n_ranges, uint32_t *keys0,
                             uint32_t *keys1, uint32_t *perm0, uint32_t *perm1, uint32_t *scratch, cudaStream_t stream); // k_sort.cu
cudaError_t launch_clear(uint8_t *color, float *depth, size_t n_pixels, float depth_max, cudaStream_t stream,
                         uint64_t *launches);
cudaError_t launch_fill_u32(uint32_t *dst, size_t n, uint32_t value, cudaStream_t stream, uint64_t *launches);
cudaError_t launch_overlay(const OverlayParams &P, cudaStream_t stream, uint64_t *launches);                        // k_overlay.cu
cudaError_t launch_mirror(const uint8_t *color, uint8_t *host_color, const uint8_t *tile_state, uint8_t *mirror_state, int W_, int H_,
                          uint32_t *counters, uint32_t *status_word, cudaStream_t stream, uint64_t *launches);      // k_mirror.cu
cudaError_t launch_flag_signal(uint32_t *flag, uint32_t value, cudaStream_t stream);                               // k_sync.cu
cudaError_t launch_flags_wait(const uint32_t *flags, uint32_t n, uint32_t value, uint32_t *error_word, cudaStream_t stream);
} // namespace drawb200

using namespace drawb200;

// ------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------
namespace {

thread_local std::string g_error = "";
int g_device = -1; // device for *_create; -1 = whatever is current

int fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_error = buf;
    return code;
}

#define CU(expr)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (expr);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            const int code_ = (e_ == cudaErrorMemoryAllocation) ? DRAW_ERR_OUT_OF_MEMORY           \
                              : (e_ == cudaErrorNoDevice || e_ == cudaErrorInsufficientDriver)     \
                                  ? DRAW_ERR_NO_DEVICE                                             \
                                  : DRAW_ERR_CUDA;                                                 \
            return fail(code_, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
        }                                                                                          \
    } while (0)

#define TRY(expr)             \
    do {                      \
        int rc_ = (expr);     \
        if (rc_ != DRAW_OK) return rc_; \
    } while (0)

// Growable device array.
template <typename T> struct DevBuf {
    T *ptr = nullptr;
    size_t cap = 0;
    ~DevBuf() { release(); }
    void release() {
        if (ptr) cudaFree(ptr);
        ptr = nullptr;
        cap = 0;
    }
    int reserve(size_t n) {
        if (n <= cap && ptr) return DRAW_OK;
        release();
        if (n == 0) n = 1;
        CU(cudaMalloc(&ptr, n * sizeof(T)));
        cap = n;
        return DRAW_OK;
    }
    int upload(const std::vector<T> &host, cudaStream_t st = nullptr) {
        TRY(reserve(host.size()));
        if (!host.empty()) CU(cudaMemcpyAsync(ptr, host.data(), host.size() * sizeof(T), cudaMemcpyHostToDevice, st));
        return DRAW_OK;
    }
};

struct HostMesh {
    std::vector<uint32_t> tris; // 9 per triangle, object-local indices
    uint32_t material;          // object-local
};
struct HostObject {
    std::string name;
    std::vector<float> pos, nrm, uv; // 3 floats each
    std::vector<HostMesh> opaque, transparent; // Object::new split, object.rs:45-53
    std::vector<MaterialDev> materials;        // offsets relative to this object's texel block
    std::vector<uint8_t> texels;
};

std::mutex g_registry_mutex;
std::set<draw_scene *> g_live_scenes;

} // namespace

// ------------------------------------------------------------------------------------------
// handles
// ------------------------------------------------------------------------------------------
// What a captured frame graph depends on besides the uniforms (which it reads from device memory).
struct GraphKey {
    SceneDev scene;
    FrameDev work;
    uint32_t n_coarse, tiles_x, tile_y_begin, tile_y_end, row_step, row_phase, grids;
};

struct draw_scene {
    int device = 0;
    size_t width = 0, height = 0;
    CameraState camera;
    f3 light{0.0f, 300.0f, 300.0f};
    std::vector<HostObject> objects;
    bool geometry_dirty = true;
    uint64_t launches = 0;

    // device geometry
    DevBuf<float> d_pos[3];
    DevBuf<float4> d_pos4, d_nrm4;
    DevBuf<float2> d_uv2;
    DevBuf<uint32_t> d_idx[9], d_mat, d_tslot;
    DevBuf<MaterialDev> d_materials;
    DevBuf<uint8_t> d_texels;
    SceneDev dev{};
    // host copy of the draw-order index streams of transparent meshes (re-sorted every frame)
    struct TransparentRange {
        size_t object, mesh, first_tri; // first_tri = position in the global draw order
    };
    std::vector<TransparentRange> transparent_ranges;
    // device painter sort (k_sort.cu): one (first, n, scratch base) per transparent mesh, key / permutation ping-pong, scratch
    DevBuf<uint4> d_sort_ranges;
    DevBuf<uint32_t> d_sort_keys[2], d_sort_perm[2], d_sort_tmp;

    // Per-frame work buffers, n_sets deep, each with its own stream: a frame runs entirely on its work set's
    // stream (the canvas' stream only brackets it), so frames that use different sets overlap — the front
    // kernel of the next frames runs beside k_tile of frame k.
    struct WorkSet {
        DevBuf<float4> vA, vLH;
        DevBuf<uint32_t> l_count, t_count, l_offset, t_offset, ms_weight, list_refs, t_refs, counters, tile_cycles,
            tile_order, empty_tiles;
        DevBuf<uint32_t> block_loc;
        DevBuf<RasterRec> rrec, trrec;
        DevBuf<PrepRec> prep, tprep;
        DevBuf<ShadeRec> srec, tsrec;
        DevBuf<uint2> l_pairs, t_pairs, m_refs, s_refs;
        DevBuf<uint4> huge_jobs;
        DevBuf<unsigned long long> key_pages;
        size_t pages_clean = 0; // pages [0, pages_clean) of key_pages.ptr are known to be empty
        FrameDev work{};
        cudaStream_t stream = nullptr;   // the frame using this set runs here
        DevBuf<FrameUniforms> d_uniforms;
        FrameUniforms *h_uniforms = nullptr; // pinned staging of d_uniforms
        cudaGraphExec_t graph_exec = nullptr;
        GraphKey graph_key{};
        uint32_t bar_base = 0;           // value of the set's grid-barrier counter before its next k_front (device_types.h)
        uint32_t front_grid = 0;         // CTAs of the k_front launch (captured in the graph)
        cudaEvent_t geo_done = nullptr;  // k_front of the frame using this set has finished (it reads the index streams)
        cudaEvent_t canvas_ready = nullptr; // canvas stream: the canvas of the frame using this set may be written
        cudaEvent_t frame_done = nullptr; // the whole frame using this set has finished
        bool frame_pending = false;
    };
    static constexpr int MAX_WORK_SETS = 8;
    WorkSet sets[MAX_WORK_SETS];
    int n_sets = 8;
    int next_set = 0, last_set = 0;
    bool debug_tile_cycles = false;
    DevBuf<uint4> d_trace;        // debug timeline (draw_scene_debug_trace), shared by the work sets
    DevBuf<uint32_t> d_trace_count;
    bool debug_trace = false;
    size_t rec_cap = 0, refs_cap = 0;
    bool needs_prepare = true; // the next render sets up every work set (streams, buffers, key pages, frame graph), not only its own
    bool graphs_ok = true; // cleared if the frame cannot be captured as a CUDA graph (direct launches then)
    // optional per-kernel timing (draw_scene_set_kernel_timing): events before k_sort_transparent, k_front, k_raster,
    // k_tile and after k_tile
    bool kernel_timing = false;
    cudaEvent_t kev[N_FRAME_KERNELS + 1] = {};
    bool kev_recorded = false;
};

struct draw_canvas {
    int device = 0;
    size_t width = 0, height = 0;
    int off_x = 0, off_y = 0;
    float depth_max = 0.0f;
    bool has_depth = false;
    bool depth_update = false;
    DevBuf<uint8_t> d_color;
    DevBuf<float> d_depth;
    uint8_t *ext_color = nullptr;
    float *ext_depth = nullptr;
    uint8_t *h_color = nullptr; // pinned mirror
    size_t h_color_cap = 0;
    bool host_dirty = true;
    bool host_mirror = false;   // every render also refreshes the pinned mirror (draw_canvas_enable_host_mirror)
    // Incremental read-back (k_mirror.cu).  tile_state: per tile, the mask of its 64x8 strips that hold something else than
    // the clear colour after the last render (written by k_tile); mirror_state: the same for the frame the host mirror holds.
    uint8_t *h_color_dev = nullptr;     // the mirror's device-side address
    DevBuf<uint8_t> tile_state, mirror_state;
    DevBuf<uint32_t> mirror_counters;
    size_t state_tiles = 0;
    bool frame_clean = false;           // the colour buffer is exactly what a whole-canvas render left: tile_state describes it
    bool mirror_valid = false;          // mirror_state describes the mirror
    cudaStream_t own_stream = nullptr, stream = nullptr;
    size_t stripe_y0 = 0, stripe_y1 = 0; // rows; y1 == 0 means whole canvas
    uint32_t row_step = 1, row_phase = 0; // tile rows ty % row_step == row_phase only (draw_canvas_set_tile_rows)
    bool empty_tile_color = true;         // draw_canvas_set_empty_tile_color
    // Pinned, device-mapped status blocks (N_STATUS_WORDS words each): k_tile posts a frame's counters into the block
    // the frame was given, so several frames may be enqueued on a canvas before the host looks at any of them.
    static constexpr int STATUS_SLOTS = 8;
    uint32_t *h_status = nullptr;
    int next_status_slot = 0;
    cudaEvent_t slot_event[STATUS_SLOTS] = {}; // recorded on the canvas' stream behind the frame that posts into the slot
    cudaEvent_t join_event = nullptr;    // draw_canvas_stream_wait
    // What a frame enqueued on this canvas and not yet looked at by the host was rendered with: an overflowed frame
    // — and every frame enqueued after it — is rendered again from exactly these inputs, in order, whatever the
    // scene's camera / light and the canvas' offset / partition have become since.
    struct FrameInputs {
        draw_scene *scene = nullptr;
        int status_slot = 0;
        int mirror_path = 0; // how the frame went to the host mirror: 0 not at all, 1 whole-frame copy, 2 changed tiles (k_mirror)
        CameraState camera;
        f3 light;
        int off_x = 0, off_y = 0;
        size_t stripe_y0 = 0, stripe_y1 = 0;
        uint32_t row_step = 1, row_phase = 0;
        float depth_max = 0.0f;
    };
    std::vector<FrameInputs> pending;
    // draw_canvas_draw_triangles scratch: the batch's vertices, its records, the bin masks and the per-bin flags
    DevBuf<uint8_t> ov_verts, ov_recs, ov_cmds;
    DevBuf<uint32_t> ov_masks, ov_any;
    draw_frame_stats stats{};
    uint64_t launches = 0;

    uint8_t *color() const { return ext_color ? ext_color : d_color.ptr; }
    float *depth() const { return ext_depth ? ext_depth : d_depth.ptr; }
};

// Texture (scene/mod.rs:206-216) as Canvas::draw_triangle uses it: map_kd, RGBA8, resident on the device
struct draw_texture {
    int device = 0;
    uint32_t width = 0, height = 0;
    DevBuf<uint8_t> texels;
};

namespace {

int env_int(const char *name, int fallback) {
    const char *v = std::getenv(name);
    return v && *v ? std::atoi(v) : fallback;
}

// Tuning knobs, read from the environment once (defaults are what bench.py measures).
struct Config {
    int graphs = env_int("DRAW_B200_GRAPH", 1);   // replay each frame as a CUDA graph
    int prio = env_int("DRAW_B200_PRIO", 0);      // work-set streams at the highest priority
    int sets = std::min(std::max(env_int("DRAW_B200_SETS", 8), 1), 8); // frames in flight per scene
    int tile_ctas = std::max(1, env_int("DRAW_B200_TILE_CTAS", 148 * 3)); // persistent CTAs of k_tile: three per SM leave room for the other frames' kernels
    int split_min_cost = std::max(1, env_int("DRAW_B200_SPLIT_MIN_COST", TILE_SPLIT_MIN_COST));
    int split_div = std::min(std::max(1, env_int("DRAW_B200_SPLIT_DIV", TILE_SPLIT_DIV)), (int)TILE_EXTRA_ITEMS);
    int split_max = std::min(std::max(1, env_int("DRAW_B200_SPLIT_MAX", TILE_MAX_SPLIT)), (int)TILE_MAX_SPLIT);
    // grids; 0 = by scene size (set_launch_grids): with several frames in flight a kernel costs the pipeline its CTAs' residency
    int front_cps = std::max(0, env_int("DRAW_B200_FRONT_CPS", 0)); // k_front CTAs per SM
    int raster_ctas = std::max(0, env_int("DRAW_B200_RASTER_CTAS", 0));
    int mirror_tiles = env_int("DRAW_B200_MIRROR_TILES", 1); // host mirror: copy only the tiles that changed when most of the frame is clear colour (k_mirror.cu)
    int sort_large = std::max(0, env_int("DRAW_B200_SORT_LARGE", 16)); // k_tile: lists of at least this many large references are tested front to back (0: never)
    int clear_first = env_int("DRAW_B200_CLEAR_FIRST", 0); // k_tile: empty-tile stores before (1) or after (0) a CTA's raster item
    int rec_cap = std::max(0, env_int("DRAW_B200_REC_CAP", 0));   // initial record / reference capacities (tests force overflows)
    int refs_cap = std::max(0, env_int("DRAW_B200_REFS_CAP", 0));
};
const Config g_cfg;

int ensure_device(int device) {
    int cur = -1;
    CU(cudaGetDevice(&cur));
    if (cur != device) CU(cudaSetDevice(device));
    return DRAW_OK;
}

int pick_device(int *out) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        cudaGetLastError();
        return fail(DRAW_ERR_NO_DEVICE, "no usable CUDA device (%s); this library has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    }
    if (g_device >= 0) {
        if (g_device >= n) return fail(DRAW_ERR_INVALID_ARGUMENT, "device %d out of range (%d devices)", g_device, n);
        *out = g_device;
        CU(cudaSetDevice(g_device));
    } else {
        CU(cudaGetDevice(out));
    }
    return DRAW_OK;
}

// Rebuilds the device geometry from the host objects: SoA, all objects concatenated, triangles
// in the reference's draw order (per object: opaque meshes, then transparent meshes).
int fetch_transparent_order(draw_scene *s);
int upload_geometry(draw_scene *s) {
    TRY(fetch_transparent_order(s));
    std::vector<float> pos[3];
    std::vector<float4> pos4, nrm4;
    std::vector<float2> uv2;
    std::vector<uint32_t> idx[9], mat, tslot;
    std::vector<MaterialDev> materials;
    std::vector<uint8_t> texels;
    s->transparent_ranges.clear();
    uint32_t n_transparent = 0;
    bool any_transparent = false;
    for (const HostObject &o : s->objects)
        if (!o.transparent.empty()) any_transparent = true;

    for (size_t oi = 0; oi < s->objects.size(); oi++) {
        const HostObject &o = s->objects[oi];
        const uint32_t vbase = (uint32_t)pos4.size(), nbase = (uint32_t)nrm4.size(), tbase = (uint32_t)uv2.size();
        const uint32_t mbase = (uint32_t)materials.size(), xbase = (uint32_t)texels.size();
        for (size_t i = 0; i < o.pos.size() / 3; i++) {
            for (int c = 0; c < 3; c++) pos[c].push_back(o.pos[3 * i + c]);
            pos4.push_back(make_float4(o.pos[3 * i], o.pos[3 * i + 1], o.pos[3 * i + 2], 0.0f));
        }
        for (size_t i = 0; i < o.nrm.size() / 3; i++) nrm4.push_back(make_float4(o.nrm[3 * i], o.nrm[3 * i + 1], o.nrm[3 * i + 2], 0.0f));
        for (size_t i = 0; i < o.uv.size() / 3; i++) uv2.push_back(make_float2(o.uv[3 * i], o.uv[3 * i + 1]));
        for (MaterialDev m : o.materials) {
            m.ka_off += xbase;
            m.kd_off += xbase;
            materials.push_back(m);
        }
        texels.insert(texels.end(), o.texels.begin(), o.texels.end());
        auto push_mesh = [&](const HostMesh &m, bool transparent) {
            const size_t n = m.tris.size() / 9;
            for (size_t t = 0; t < n; t++) {
                const uint32_t *p = &m.tris[9 * t];
                for (int c = 0; c < 3; c++) {
                    idx[c].push_back(p[c] + vbase);
                    idx[3 + c].push_back(p[3 + c] + tbase);
                    idx[6 + c].push_back(p[6 + c] + nbase);
                }
                mat.push_back((m.material + mbase) | (transparent ? 0x80000000u : 0u));
                if (any_transparent) tslot.push_back(transparent ? n_transparent++ : 0u);
            }
        };
        for (const HostMesh &m : o.opaque) push_mesh(m, false);
        for (size_t mi = 0; mi < o.transparent.size(); mi++) {
            s->transparent_ranges.push_back({oi, mi, mat.size()});
            push_mesh(o.transparent[mi], true);
        }
    }
    if (mat.size() >= (1u << 29)) return fail(DRAW_ERR_INVALID_ARGUMENT, "too many triangles (%zu)", mat.size());

    for (int c = 0; c < 3; c++) {
        while (pos[c].size() % 4) pos[c].push_back(0.0f); // k_front reads the streams four vertices at a time
        TRY(s->d_pos[c].upload(pos[c]));
    }
    TRY(s->d_pos4.upload(pos4));
    TRY(s->d_nrm4.upload(nrm4));
    TRY(s->d_uv2.upload(uv2));
    for (int c = 0; c < 9; c++) TRY(s->d_idx[c].upload(idx[c]));
    TRY(s->d_mat.upload(mat));
    TRY(s->d_tslot.upload(tslot));
    TRY(s->d_materials.upload(materials));
    TRY(s->d_texels.upload(texels));
    CU(cudaDeviceSynchronize()); // the host vectors above are about to go away

    SceneDev &d = s->dev;
    d.px = s->d_pos[0].ptr; d.py = s->d_pos[1].ptr; d.pz = s->d_pos[2].ptr;
    d.pos4 = s->d_pos4.ptr; d.nrm4 = s->d_nrm4.ptr; d.uv2 = s->d_uv2.ptr;
    for (int c = 0; c < 9; c++) d.idx[c] = s->d_idx[c].ptr;
    d.tri_mat = s->d_mat.ptr;
    d.tri_tslot = any_transparent ? s->d_tslot.ptr : nullptr;
    d.materials = s->d_materials.ptr;
    d.texels = s->d_texels.ptr;
    d.n_vertices = (uint32_t)pos4.size();
    d.n_triangles = (uint32_t)mat.size();
    d.n_transparent = n_transparent;
    d.n_materials = (uint32_t)materials.size();
    if (n_transparent) {
        std::vector<uint4> ranges;
        uint32_t base = 0;
        for (const draw_scene::TransparentRange &tr : s->transparent_ranges) {
            const uint32_t n = (uint32_t)(s->objects[tr.object].transparent[tr.mesh].tris.size() / 9);
            ranges.push_back(make_uint4((uint32_t)tr.first_tri, n, base, 0u));
            base += n;
        }
        TRY(s->d_sort_ranges.upload(ranges));
        for (int i = 0; i < 2; i++) {
            TRY(s->d_sort_keys[i].reserve(base));
            TRY(s->d_sort_perm[i].reserve(base));
        }
        TRY(s->d_sort_tmp.reserve(base));
        CU(cudaDeviceSynchronize());
    }
    s->geometry_dirty = false;
    return DRAW_OK;
}

int ensure_work_buffers(draw_scene *s, draw_scene::WorkSet &ws, size_t n_tiles) {
    const SceneDev &d = s->dev;
    TRY(ws.vA.reserve(d.n_vertices));
    TRY(ws.vLH.reserve(2 * (size_t)d.n_vertices));
    if (s->rec_cap == 0) s->rec_cap = g_cfg.rec_cap ? (size_t)g_cfg.rec_cap : 2 * (size_t)d.n_triangles + 4096;
    if (s->refs_cap == 0) s->refs_cap = g_cfg.refs_cap ? (size_t)g_cfg.refs_cap : std::max<size_t>((size_t)1 << 22, 4 * (size_t)d.n_triangles);
    const size_t t_refs_cap = d.n_transparent ? s->refs_cap : 1;
    TRY(ws.rrec.reserve(s->rec_cap));
    TRY(ws.srec.reserve(s->rec_cap));
    TRY(ws.prep.reserve(s->rec_cap));
    TRY(ws.trrec.reserve(4 * (size_t)d.n_transparent));
    TRY(ws.tsrec.reserve(4 * (size_t)d.n_transparent));
    TRY(ws.tprep.reserve(4 * (size_t)d.n_transparent));
    TRY(ws.l_count.reserve(n_tiles));
    TRY(ws.t_count.reserve(n_tiles));
    TRY(ws.l_offset.reserve(n_tiles));
    TRY(ws.t_offset.reserve(n_tiles));
    TRY(ws.ms_weight.reserve(n_tiles));
    TRY(ws.l_pairs.reserve(s->refs_cap));
    TRY(ws.list_refs.reserve(s->refs_cap));
    TRY(ws.t_pairs.reserve(t_refs_cap));
    TRY(ws.t_refs.reserve(t_refs_cap));
    const size_t huge_cap = std::min<size_t>(s->rec_cap, (size_t)1 << 20);
    TRY(ws.huge_jobs.reserve(huge_cap));
    TRY(ws.m_refs.reserve(s->refs_cap));
    TRY(ws.s_refs.reserve(s->refs_cap));
    {
        // key pages: one per tile, all empty between frames (k_tile leaves the pages it consumed empty)
        const unsigned long long *before = ws.key_pages.ptr;
        TRY(ws.key_pages.reserve(n_tiles * TILE_W * TILE_H));
        if (ws.key_pages.ptr != before) ws.pages_clean = 0;
        if (n_tiles > ws.pages_clean) {
            cudaStream_t st = ws.stream ? ws.stream : 0;
            CU(launch_fill_u64(ws.key_pages.ptr + ws.pages_clean * TILE_W * TILE_H, (n_tiles - ws.pages_clean) * TILE_W * TILE_H, KEY_EMPTY, st));
            if (!ws.stream) CU(cudaStreamSynchronize(0));
            ws.pages_clean = n_tiles;
        }
    }
    if (!ws.counters.ptr) {
        TRY(ws.counters.reserve(N_COUNTERS));
        CU(cudaMemset(ws.counters.ptr, 0, N_COUNTERS * sizeof(uint32_t))); // the grid-barrier word starts at bar_base = 0
        ws.bar_base = 0;
    }
    TRY(ws.tile_order.reserve((size_t)COST_BUCKETS * (n_tiles + TILE_EXTRA_ITEMS)));
    TRY(ws.empty_tiles.reserve(n_tiles));
    TRY(ws.block_loc.reserve((size_t)d.n_triangles / 128 + 2)); // one entry per 128-triangle block (k_front.cu: FRONT_THREADS)
    FrameDev &w = ws.work;
    w.vA = ws.vA.ptr; w.vLH = ws.vLH.ptr;
    w.rrec = ws.rrec.ptr; w.srec = ws.srec.ptr; w.prep = ws.prep.ptr;
    w.t_rrec = ws.trrec.ptr; w.t_srec = ws.tsrec.ptr; w.t_prep = ws.tprep.ptr;
    w.l_count = ws.l_count.ptr; w.t_count = ws.t_count.ptr; w.l_offset = ws.l_offset.ptr; w.t_offset = ws.t_offset.ptr;
    w.ms_weight = ws.ms_weight.ptr;
    w.l_pairs = ws.l_pairs.ptr; w.t_pairs = ws.t_pairs.ptr; w.list_refs = ws.list_refs.ptr; w.t_refs = ws.t_refs.ptr;
    w.m_refs = ws.m_refs.ptr; w.s_refs = ws.s_refs.ptr;
    w.huge_jobs = ws.huge_jobs.ptr; w.huge_cap = (uint32_t)huge_cap;
    w.key_pages = ws.key_pages.ptr;
    w.counters = ws.counters.ptr;
    w.tile_order = ws.tile_order.ptr;
    w.bucket_cap = (uint32_t)(n_tiles + TILE_EXTRA_ITEMS);
    w.empty_tiles = ws.empty_tiles.ptr;
    w.block_loc = ws.block_loc.ptr;
    w.rec_cap = (uint32_t)s->rec_cap;
    w.refs_cap = (uint32_t)std::min(s->refs_cap, t_refs_cap == 1 ? s->refs_cap : t_refs_cap);
    w.tile_cycles = nullptr;
    if (s->debug_tile_cycles) {
        TRY(ws.tile_cycles.reserve(4 * n_tiles));
        w.tile_cycles = ws.tile_cycles.ptr;
    }
    w.trace = nullptr;
    w.trace_count = nullptr;
    w.trace_cap = 0;
    w.trace_tag = (uint32_t)(&ws - s->sets);
    if (s->debug_trace) {
        w.trace = s->d_trace.ptr;
        w.trace_count = s->d_trace_count.ptr;
        w.trace_cap = (uint32_t)s->d_trace.cap;
    }
    if (!ws.geo_done) CU(cudaEventCreateWithFlags(&ws.geo_done, cudaEventDisableTiming));
    if (!ws.canvas_ready) CU(cudaEventCreateWithFlags(&ws.canvas_ready, cudaEventDisableTiming));
    if (!ws.frame_done) CU(cudaEventCreateWithFlags(&ws.frame_done, cudaEventDisableTiming));
    return DRAW_OK;
}

// Painter sort of every transparent mesh (scene/mod.rs:1100-1115): stable, far to near by
// f32::total_cmp of the centroid distance, persistent across frames; re-uploads the index
// streams of meshes whose order changed.
// The device keeps the transparent meshes' triangle lists in the order the painter sort left them
// (k_sort.cu; the reference's lists persist the same way, scene/mod.rs:1100-1115).  Before the geometry
// is rebuilt from the host objects (an object was added) that order is read back into them.
int fetch_transparent_order(draw_scene *s) {
    if (s->transparent_ranges.empty() || !s->d_idx[0].ptr) return DRAW_OK;
    CU(cudaDeviceSynchronize());
    for (const draw_scene::TransparentRange &tr : s->transparent_ranges) {
        HostObject &o = s->objects[tr.object];
        HostMesh &m = o.transparent[tr.mesh];
        const size_t n = m.tris.size() / 9;
        uint32_t vbase = 0, nbase = 0, tbase = 0; // global index bases of this object
        for (size_t oi = 0; oi < tr.object; oi++) {
            vbase += (uint32_t)(s->objects[oi].pos.size() / 3);
            nbase += (uint32_t)(s->objects[oi].nrm.size() / 3);
            tbase += (uint32_t)(s->objects[oi].uv.size() / 3);
        }
        std::vector<uint32_t> stream_host(n);
        for (int c = 0; c < 9; c++) {
            const uint32_t base = c < 3 ? vbase : (c < 6 ? tbase : nbase);
            CU(cudaMemcpy(stream_host.data(), s->d_idx[c].ptr + tr.first_tri, n * sizeof(uint32_t), cudaMemcpyDeviceToHost));
            for (size_t t = 0; t < n; t++) m.tris[9 * t + c] = stream_host[t] - base;
        }
    }
    return DRAW_OK;
}

// The launches of one frame on the work set's stream (directly, or under stream capture): the uniforms, the
// painter sort if the scene has transparent meshes, k_front, k_raster, k_tile.
int launch_frame(draw_scene *s, draw_scene::WorkSet &ws, const FrameUniforms &U, cudaStream_t side, cudaEvent_t *ev,
                 bool capturing) {
    const FrameUniforms *dU = ws.d_uniforms.ptr;
    // ws.canvas_ready is recorded on the canvas' stream outside the graph: an external event of the capture
    const unsigned ext = capturing ? cudaEventWaitExternal : 0u;
    CU(cudaMemcpyAsync(ws.d_uniforms.ptr, ws.h_uniforms, sizeof(FrameUniforms), cudaMemcpyHostToDevice, side));
    if (ev) cudaEventRecord(ev[0], side);
    // painter sort of the transparent meshes, in place in the shared index streams (enqueue_frame has ordered
    // this stream after the previous frame's k_front, which reads them)
    if (s->dev.n_transparent)
        launch_sort_transparent(dU, s->dev, s->d_sort_ranges.ptr, (uint32_t)s->transparent_ranges.size(), s->d_sort_keys[0].ptr,
                                s->d_sort_keys[1].ptr, s->d_sort_perm[0].ptr, s->d_sort_perm[1].ptr, s->d_sort_tmp.ptr, side);
    if (ev) cudaEventRecord(ev[1], side);
    ws.front_grid = g_front_ctas;
    CU(launch_front(dU, s->dev, ws.work, side));
    // a real event in both paths (an event-record node under capture): the next frame's painter sort waits for it
    CU(cudaEventRecordWithFlags(ws.geo_done, side, capturing ? cudaEventRecordExternal : 0u));
    if (ev) cudaEventRecord(ev[2], side);
    launch_raster(dU, ws.work, side);
    if (ev) cudaEventRecord(ev[3], side);
    // what the canvas' stream still does with the canvas comes before the kernel that writes it
    CU(cudaStreamWaitEvent(side, ws.canvas_ready, ext));
    launch_tile(U, dU, s->dev, ws.work, side);
    if (ev) cudaEventRecord(ev[4], side);
    return DRAW_OK;
}

int ensure_host_mirror(draw_canvas *c, size_t bytes) {
    if (c->h_color_cap >= bytes) return DRAW_OK;
    if (c->h_color) cudaFreeHost(c->h_color);
    c->h_color = nullptr;
    c->h_color_cap = 0;
    CU(cudaMallocHost(&c->h_color, bytes));
    c->h_color_cap = bytes;
    c->host_dirty = true;
    c->mirror_valid = false;
    void *dev = nullptr;
    CU(cudaHostGetDevicePointer(&dev, c->h_color, 0));
    c->h_color_dev = static_cast<uint8_t *>(dev);
    return DRAW_OK;
}

// Per-tile state of the canvas' frame and of its host mirror (k_mirror.cu), sized for the current canvas.
int ensure_tile_state(draw_canvas *c) {
    const size_t n = ((c->width + TILE_W - 1) / TILE_W) * ((c->height + TILE_H - 1) / TILE_H);
    if (c->state_tiles == n && c->tile_state.ptr) return DRAW_OK;
    TRY(c->tile_state.reserve(n));
    TRY(c->mirror_state.reserve(n));
    TRY(c->mirror_counters.reserve(2));
    CU(cudaMemsetAsync(c->tile_state.ptr, (int)TILE_STRIPS_ALL, n, c->stream)); // unknown content: every strip
    CU(cudaMemsetAsync(c->mirror_state.ptr, (int)TILE_STRIPS_ALL, n, c->stream));
    CU(cudaMemsetAsync(c->mirror_counters.ptr, 0, 2 * sizeof(uint32_t), c->stream));
    c->state_tiles = n;
    c->frame_clean = false;
    c->mirror_valid = false;
    return DRAW_OK;
}

// The canvas' colour buffer was written by something other than a whole-canvas render.
void touch_frame(draw_canvas *c) {
    c->host_dirty = true;
    c->frame_clean = false;
}

// Brings the pinned host mirror up to date with the device frame, on stream st.  When the frame is what a whole-canvas
// render left, the mirror's state is known and most of the last frame was clear colour, only the tiles that differ cross
// the bus (k_mirror, *path = 2, tile count posted to status_word); otherwise the whole frame is copied (*path = 1).
int refresh_mirror(draw_canvas *c, cudaStream_t st, uint32_t *status_word, int *path) {
    const size_t bytes = c->width * c->height * 4;
    TRY(ensure_host_mirror(c, bytes));
    TRY(ensure_tile_state(c));
    const bool sparse = c->stats.empty_tiles * 2u > c->state_tiles; // the last frame looked at on this canvas
    if (g_cfg.mirror_tiles && c->frame_clean && c->mirror_valid && sparse) {
        CU(launch_mirror(c->color(), c->h_color_dev, c->tile_state.ptr, c->mirror_state.ptr, (int)c->width, (int)c->height,
                         c->mirror_counters.ptr, status_word, st, &c->launches));
        if (path) *path = 2;
    } else {
        CU(cudaMemcpyAsync(c->h_color, c->color(), bytes, cudaMemcpyDeviceToHost, st));
        if (c->frame_clean) { // the mirror now holds this frame: its state is the frame's
            CU(cudaMemcpyAsync(c->mirror_state.ptr, c->tile_state.ptr, c->state_tiles, cudaMemcpyDeviceToDevice, st));
            c->mirror_valid = true;
        } else {
            c->mirror_valid = false;
        }
        if (path) *path = 1;
    }
    c->host_dirty = false;
    return DRAW_OK;
}

// Grids of the frame's kernels, by scene size (a kernel costs the pipeline its CTAs' residency).  k_front: one CTA of
// 128 threads per SM for small scenes — it fits beside three k_tile CTAs of another frame, and with several frames in
// flight that is worth more than the few microseconds a second CTA per SM saves a lone frame (measured: C3 26.3 k
// frames/s against 23.3 k) — and as many as fit (eight) for large ones, whose phases need the latency hiding.
void set_launch_grids(const draw_scene *s) {
    static const int front_cap = front_max_ctas_per_sm();
    static const int n_sm = [] {
        int dev = 0, n = 148;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) n = 148;
        return n;
    }();
    const bool small_scene = s->dev.n_triangles <= 200000u;
    const int cps = std::min(front_cap, g_cfg.front_cps ? g_cfg.front_cps : (small_scene ? 1 : 8)); // CTAs of 128 threads
    g_front_ctas = (unsigned)(n_sm * std::max(1, cps));
    // k_raster (128-thread CTAs): four per SM for small scenes (measured on C3: 296 / 444 / 592 / 740 CTAs give 28.1 / 27.8 / 27.6 / 27.5 k
    // frames/s back to back; fewer CTAs cost a lone frame its raster warps: C4's k_raster alone 50 -> 71 us at 444), sixteen for large ones
    g_raster_ctas = g_cfg.raster_ctas ? (unsigned)g_cfg.raster_ctas : (small_scene ? 592u : 2368u);
    g_tile_ctas = (unsigned)g_cfg.tile_ctas;
}

// Frame constants that depend on the scene, the canvas and the tuning knobs (everything but the work set).
void fill_uniforms(draw_scene *s, const draw_canvas *c, FrameUniforms &U) {
    const uint32_t tiles_x = (uint32_t)((c->width + TILE_W - 1) / TILE_W), tiles_y = (uint32_t)((c->height + TILE_H - 1) / TILE_H);
    const m4 m = transformation_matrix(s->camera, s->width, s->height); // :904
    std::memcpy(U.m, m.v, sizeof U.m);
    plane4 planes[6];
    s->camera.view_planes(planes); // :908
    for (int i = 0; i < 6; i++) {
        U.planes[i][0] = planes[i].nx; U.planes[i][1] = planes[i].ny;
        U.planes[i][2] = planes[i].nz; U.planes[i][3] = planes[i].k;
    }
    U.cam[0] = s->camera.position.x; U.cam[1] = s->camera.position.y; U.cam[2] = s->camera.position.z;
    U.light[0] = s->light.x; U.light[1] = s->light.y; U.light[2] = s->light.z;
    U.off_x = (float)c->off_x;
    U.off_y = (float)c->off_y;
    U.depth_max = c->depth_max;
    U.canvas_w = (uint32_t)c->width;
    U.canvas_h = (uint32_t)c->height;
    U.tiles_x = tiles_x;
    U.tiles_y = tiles_y;
    U.n_coarse = tiles_x * tiles_y;
    U.split_min_cost = (uint32_t)g_cfg.split_min_cost;
    U.split_div = (uint32_t)g_cfg.split_div;
    U.split_max = (uint32_t)g_cfg.split_max;
    const size_t y0 = c->stripe_y1 ? c->stripe_y0 : 0, y1 = c->stripe_y1 ? c->stripe_y1 : c->height;
    U.tile_y_begin = (uint32_t)(y0 / TILE_H);
    U.tile_y_end = (uint32_t)((y1 + TILE_H - 1) / TILE_H);
    U.row_step = c->row_step ? c->row_step : 1u;
    U.row_phase = c->row_phase % U.row_step;
    U.bar_base = 0; // set per work set by enqueue_frame
    U.clear_first = (uint32_t)g_cfg.clear_first;
    U.sort_large = (uint32_t)g_cfg.sort_large;
    U.empty_tile_color = c->empty_tile_color ? 1u : 0u;
    U.status_host = c->h_status + (size_t)c->next_status_slot * N_STATUS_WORDS; // pinned, mapped: valid on the device (unified addressing)
    U.tile_state = c->tile_state.ptr; // enqueue_frame has sized it
    U.color = c->color();
    U.depth = c->depth();
}

// Makes one work set ready for frames of this scene / canvas geometry: its stream and pinned uniforms, its
// buffers (key pages filled), and — unless a measurement or debug tap needs the direct path — the frame's
// launches captured once as a CUDA graph (they do not change from frame to frame: the uniforms are read
// from device memory).
int ensure_set_ready(draw_scene *s, draw_scene::WorkSet &ws, const FrameUniforms &U, bool want_graph) {
    if (!ws.stream) {
        int prio_low = 0, prio_high = 0;
        CU(cudaDeviceGetStreamPriorityRange(&prio_low, &prio_high));
        CU(cudaStreamCreateWithPriority(&ws.stream, cudaStreamNonBlocking, g_cfg.prio ? prio_high : prio_low));
        CU(cudaMallocHost(&ws.h_uniforms, sizeof(FrameUniforms)));
        TRY(ws.d_uniforms.reserve(1));
    }
    TRY(ensure_work_buffers(s, ws, U.n_coarse));
    if (!want_graph || !s->graphs_ok) return DRAW_OK;
    GraphKey key{};
    key.scene = s->dev;
    key.work = ws.work;
    key.n_coarse = U.n_coarse; key.tiles_x = U.tiles_x;
    key.tile_y_begin = U.tile_y_begin; key.tile_y_end = U.tile_y_end; key.row_step = U.row_step; key.row_phase = U.row_phase;
    key.grids = g_front_ctas + 4099u * g_tile_ctas + 1000003u * g_raster_ctas;
    if (ws.graph_exec && std::memcmp(&key, &ws.graph_key, sizeof key) == 0) return DRAW_OK;
    if (ws.graph_exec) CU(cudaGraphExecDestroy(ws.graph_exec));
    ws.graph_exec = nullptr;
    CU(cudaStreamBeginCapture(ws.stream, cudaStreamCaptureModeThreadLocal));
    const int rc = launch_frame(s, ws, U, ws.stream, nullptr, true);
    cudaGraph_t graph = nullptr;
    cudaError_t e = cudaStreamEndCapture(ws.stream, &graph);
    if (rc == DRAW_OK && e == cudaSuccess) {
        e = cudaGraphInstantiate(&ws.graph_exec, graph, 0);
        if (e == cudaSuccess) e = cudaGraphUpload(ws.graph_exec, ws.stream); // the first replay does not pay the upload
    }
    if (graph) cudaGraphDestroy(graph);
    if (rc != DRAW_OK || e != cudaSuccess) {
        // e.g. a driver that cannot capture a cooperative launch: frames are launched directly from now on
        cudaGetLastError();
        if (ws.graph_exec) cudaGraphExecDestroy(ws.graph_exec);
        ws.graph_exec = nullptr;
        s->graphs_ok = false;
        return DRAW_OK;
    }
    ws.graph_key = key;
    return DRAW_OK;
}

int enqueue_frame(draw_scene *s, draw_canvas *c) {
    TRY(ensure_device(s->device));
    if (s->geometry_dirty) TRY(upload_geometry(s));
    TRY(ensure_tile_state(c));
    FrameUniforms U{};
    fill_uniforms(s, c, U);
    if (U.tiles_x >= MAX_TILES_X || U.tiles_y >= MAX_TILES_Y)
        return fail(DRAW_ERR_INVALID_ARGUMENT, "canvas %zux%zu is too large for the tile work list", c->width, c->height);
    draw_scene::WorkSet &ws = s->sets[s->next_set];
    const draw_scene::WorkSet &prev_ws = s->sets[s->last_set];
    s->last_set = s->next_set;
    s->next_set = (s->next_set + 1) % s->n_sets;

    set_launch_grids(s);
    // A frame's launches are replayed as a CUDA graph: one launch call instead of a copy, three or four kernels and
    // two events.  Measurement taps and the debug taps use the direct path.
    const bool timing = s->kernel_timing;
    const bool want_graph = g_cfg.graphs && !timing && !s->debug_tile_cycles;
    if (s->needs_prepare) {
        // First frame of this scene (or of its new geometry / capacities): every work set is set up now — stream,
        // ~30 buffers, the fill of its key pages, graph capture and instantiation — so that no later frame pays a
        // few milliseconds of set-up in the middle of a steady stream of frames (draw_scene_prepare does the same).
        for (int i = 0; i < s->n_sets; i++) {
            if (s->sets[i].frame_pending) CU(cudaEventSynchronize(s->sets[i].frame_done));
            TRY(ensure_set_ready(s, s->sets[i], U, want_graph));
        }
        s->needs_prepare = false;
    }
    if (ws.frame_pending) CU(cudaEventSynchronize(ws.frame_done)); // the pinned uniforms of the set are about to be rewritten
    TRY(ensure_set_ready(s, ws, U, want_graph));
    const bool use_graph = want_graph && s->graphs_ok && ws.graph_exec;

    // ---- enqueue ----------------------------------------------------------------------------------
    // Everything of the frame runs on the work set's own stream.  The canvas' stream only brackets the frame:
    // k_tile first waits for whatever the canvas stream still does with the canvas, and the canvas stream then
    // waits for the frame.  Frames that use different work sets are therefore independent and overlap; a set's
    // next frame follows its previous one in stream order.
    cudaStream_t side = ws.stream, st = c->stream;
    U.bar_base = ws.bar_base;
    *ws.h_uniforms = U;
    // (k_tile waits for this event; k_front and k_raster do not touch the canvas and do not wait)
    CU(cudaEventRecord(ws.canvas_ready, st));
    if (s->dev.n_transparent) {
        // the painter sort rewrites the shared index streams: order it after the previous frame's k_front
        // (geo_done is a real event in both paths: launch_frame records it with cudaEventRecordExternal under capture)
        if (&prev_ws != &ws && prev_ws.frame_pending) CU(cudaStreamWaitEvent(side, prev_ws.geo_done, 0));
    }
    if (ws.work.tile_cycles) CU(cudaMemsetAsync(ws.work.tile_cycles, 0, 4 * (size_t)U.n_coarse * sizeof(uint32_t), side)); // debug taps are atomicMax'd

    cudaEvent_t *ev = nullptr;
    if (timing) {
        for (int i = 0; i < N_FRAME_KERNELS + 1; i++)
            if (!s->kev[i]) CU(cudaEventCreate(&s->kev[i]));
        ev = s->kev;
        s->kev_recorded = true;
    }
    if (use_graph) CU(cudaGraphLaunch(ws.graph_exec, side));
    else TRY(launch_frame(s, ws, U, side, ev, false));
    ws.bar_base += 3u * ws.front_grid; // k_front's three grid barriers (k_front.cu; the third is only waited at if the frame has huge records)
    CU(cudaEventRecord(ws.frame_done, side));
    ws.frame_pending = true;
    CU(cudaStreamWaitEvent(st, ws.frame_done, 0));
    s->launches += 2 + (s->dev.n_transparent ? 1 : 0) + (tile_grid_items(U) ? 1 : 0);
    CU(cudaGetLastError());
    c->host_dirty = true;
    // whole canvas, empty tiles written: the colour buffer is this render's and nothing else's
    c->frame_clean = c->stripe_y1 == 0 && U.row_step == 1u && c->empty_tile_color && !c->ext_color;
    int mirror_path = 0;
    if (c->host_mirror && !c->ext_color) // the frame follows its render to the host without waiting for the host to ask (map_host then only waits)
        TRY(refresh_mirror(c, st, U.status_host + CNT_MIRROR_STRIPS, &mirror_path));
    {
        draw_canvas::FrameInputs in;
        in.scene = s;
        in.status_slot = c->next_status_slot;
        in.mirror_path = mirror_path;
        in.camera = s->camera;
        in.light = s->light;
        in.off_x = c->off_x; in.off_y = c->off_y;
        in.stripe_y0 = c->stripe_y0; in.stripe_y1 = c->stripe_y1;
        in.row_step = c->row_step; in.row_phase = c->row_phase;
        in.depth_max = c->depth_max;
        if (!c->slot_event[in.status_slot]) CU(cudaEventCreateWithFlags(&c->slot_event[in.status_slot], cudaEventDisableTiming));
        CU(cudaEventRecord(c->slot_event[in.status_slot], st)); // behind the frame and its way to the host
        c->pending.push_back(in);
        c->next_status_slot = (c->next_status_slot + 1) % draw_canvas::STATUS_SLOTS;
    }
    return DRAW_OK;
}

// Statistics of a completed frame, from the status block k_tile posted them to.
void record_stats(draw_canvas *c, const draw_canvas::FrameInputs &frame) {
    const uint32_t *st = c->h_status + (size_t)frame.status_slot * N_STATUS_WORDS;
    c->stats.setup_records = st[CNT_RECORDS];
    c->stats.large_refs = st[CNT_L_PAIRS];
    c->stats.medium_refs = st[CNT_MEDIUM];
    c->stats.small_refs = st[CNT_SMALL];
    c->stats.transparent_refs = st[CNT_T_PAIRS];
    c->stats.tile_refs = c->stats.large_refs + c->stats.medium_refs + c->stats.small_refs + c->stats.transparent_refs;
    c->stats.empty_tiles = st[CNT_EMPTY];
    c->stats.work_items = st[CNT_ITEMS];
    // KiB that went to the host mirror: strips of TILE_W x 8 pixels x 4 bytes, or the whole frame
    c->stats.mirror_kbytes = frame.mirror_path == 2 ? st[CNT_MIRROR_STRIPS] * (uint32_t)(TILE_W * TILE_STRIP_H * 4 / 1024)
                             : frame.mirror_path == 1 ? (uint32_t)((c->width * c->height * 4 + 1023) / 1024) : 0u;
    for (int i = 0; i < 5; i++) c->stats.front_phase_ns[i] = st[CNT_PHASE_NS + i + 1] - st[CNT_PHASE_NS + i];
    c->stats.front_phase_ns[5] = st[CNT_PHASE_NS + 6] - st[CNT_PHASE_NS + 2];                              // triangle phase, slowest CTA
    c->stats.front_phase_ns[6] = st[CNT_PHASE_NS + 7] ? st[CNT_PHASE_NS + 7] - st[CNT_PHASE_NS + 6] : 0u; // huge-record phase (after the barrier)
    for (int i = 0; i < 5; i++) c->stats.front_block_ns[i] = st[CNT_PHASE_NS + 8 + i];
    bool alive;
    {
        std::lock_guard<std::mutex> lk(g_registry_mutex);
        alive = g_live_scenes.count(frame.scene) != 0;
    }
    if (alive) c->stats.input_triangles = frame.scene->dev.n_triangles;
}

// Waits for the canvas' stream and settles the frames enqueued on it since the last call: statistics of the last
// one; if any of them overflowed a work buffer, the buffers are grown and that frame and every later one are
// rendered again, in order, each from the inputs it was enqueued with (so what the host reads is always complete).
int finish_frame(draw_canvas *c) {
    TRY(ensure_device(c->device));
    CU(cudaStreamSynchronize(c->stream));
    for (int guard = 0; !c->pending.empty(); guard++) {
        const std::vector<draw_canvas::FrameInputs> frames = c->pending;
        c->pending.clear();
        size_t first_bad = frames.size();
        uint32_t overflow = 0, n_rec = 0, n_refs = 0;
        for (size_t i = 0; i < frames.size(); i++) {
            const uint32_t *st = c->h_status + (size_t)frames[i].status_slot * N_STATUS_WORDS;
            if (st[CNT_OVERFLOW] && first_bad == frames.size()) first_bad = i;
            if (i >= first_bad) {
                overflow |= st[CNT_OVERFLOW];
                n_rec = std::max(n_rec, st[CNT_RECORDS]);
                n_refs = std::max(n_refs, st[CNT_REFS_NEEDED]);
            }
        }
        record_stats(c, frames.back());
        if (first_bad == frames.size()) break;
        c->stats.overflow |= overflow;
        if (overflow & OVERFLOW_STALL) return fail(DRAW_ERR_INTERNAL, "a grid barrier of k_front timed out (the launch was not co-resident)");
        if (overflow & OVERFLOW_HUGE) return fail(DRAW_ERR_INTERNAL, "more than 2^20 records each cover hundreds of tiles");
        if (guard >= 4) return fail(DRAW_ERR_INTERNAL, "work buffers keep overflowing");
        CU(cudaDeviceSynchronize()); // the work sets are about to be reallocated
        for (size_t i = first_bad; i < frames.size(); i++) {
            const draw_canvas::FrameInputs &in = frames[i];
            draw_scene *fs = in.scene;
            {
                std::lock_guard<std::mutex> lk(g_registry_mutex);
                if (!g_live_scenes.count(fs)) return fail(DRAW_ERR_INTERNAL, "frame overflowed a work buffer and its scene is gone");
            }
            if (i == first_bad || fs != frames[i - 1].scene) {
                if (overflow & OVERFLOW_RECORDS)
                    fs->rec_cap = std::max(fs->rec_cap, std::max<size_t>((size_t)n_rec + n_rec / 4 + 1024, 4 * (size_t)fs->dev.n_triangles + 1024));
                if (overflow & OVERFLOW_REFS) fs->refs_cap = std::max(fs->refs_cap, (size_t)n_refs + n_refs / 4 + 4096);
                else if (overflow & OVERFLOW_RECORDS) fs->refs_cap = std::max(fs->refs_cap, 4 * fs->rec_cap);
                fs->needs_prepare = true;
            }
            // render the frame again from the inputs it was rendered with, then put the current state back
            const CameraState cam_now = fs->camera;
            const f3 light_now = fs->light;
            const int off_x_now = c->off_x, off_y_now = c->off_y;
            const size_t sy0_now = c->stripe_y0, sy1_now = c->stripe_y1;
            const uint32_t rstep_now = c->row_step, rphase_now = c->row_phase;
            const float dmax_now = c->depth_max;
            fs->camera = in.camera; fs->light = in.light;
            c->off_x = in.off_x; c->off_y = in.off_y;
            c->stripe_y0 = in.stripe_y0; c->stripe_y1 = in.stripe_y1;
            c->row_step = in.row_step; c->row_phase = in.row_phase;
            c->depth_max = in.depth_max;
            const int rc = enqueue_frame(fs, c);
            fs->camera = cam_now; fs->light = light_now;
            c->off_x = off_x_now; c->off_y = off_y_now;
            c->stripe_y0 = sy0_now; c->stripe_y1 = sy1_now;
            c->row_step = rstep_now; c->row_phase = rphase_now;
            c->depth_max = dmax_now;
            TRY(rc);
        }
        CU(cudaStreamSynchronize(c->stream));
    }
    return DRAW_OK;
}

int fill_color_black(draw_canvas *c, size_t first, size_t count) {
    if (!count) return DRAW_OK;
    // Pixel::black(): b=0 g=0 r=0 pad=255 (canvas.rs:63-70,128-130)
    CU(launch_fill_u32(reinterpret_cast<uint32_t *>(c->d_color.ptr) + first, count, 0xFF000000u, c->stream, &c->launches));
    return DRAW_OK;
}

} // namespace

// ------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------
#define GUARD_BEGIN try {
#define GUARD_END                                                                  \
    }                                                                              \
    catch (const std::bad_alloc &) { return fail(DRAW_ERR_OUT_OF_MEMORY, "host allocation failed"); } \
    catch (const std::exception &e) { return fail(DRAW_ERR_INTERNAL, "internal error: %s", e.what()); } \
    catch (...) { return fail(DRAW_ERR_INTERNAL, "internal error"); }

extern "C" {

int draw_version(void) { return DRAW_B200_VERSION; }
const char *draw_last_error(void) { return g_error.c_str(); }

int draw_device_count(int *out_count) {
    if (!out_count) return fail(DRAW_ERR_INVALID_ARGUMENT, "out_count is NULL");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        cudaGetLastError();
        n = 0;
    }
    *out_count = n;
    return DRAW_OK;
}

int draw_set_device(int device) {
    if (device < 0) return fail(DRAW_ERR_INVALID_ARGUMENT, "negative device index");
    g_device = device;
    return DRAW_OK;
}

int draw_tile_size(void) { return TILE_H; }

// ---- Scene ---------------------------------------------------------------------------------

int draw_scene_create(size_t width, size_t height, draw_scene **out) {
    GUARD_BEGIN
    if (!out) return fail(DRAW_ERR_INVALID_ARGUMENT, "out is NULL");
    *out = nullptr;
    if (width == 0 || height == 0) return fail(DRAW_ERR_INVALID_ARGUMENT, "scene size must be non-zero");
    int dev = 0;
    TRY(pick_device(&dev));
    draw_scene *s = new draw_scene();
    s->device = dev;
    s->width = width;
    s->height = height;
    s->n_sets = g_cfg.sets;
    // Scene::new, scene/mod.rs:760-786
    const f3 pos{0.0f, 0.0f, 150.0f};
    const f3 dir = scale(pos, -1.0f);
    const float ratio = (float)width / (float)height;
    s->camera = CameraState::make(pos, dir, ratio);
    {
        std::lock_guard<std::mutex> lk(g_registry_mutex);
        g_live_scenes.insert(s);
    }
    *out = s;
    return DRAW_OK;
    GUARD_END
}

void draw_scene_destroy(draw_scene *scene) {
    if (!scene) return;
    {
        std::lock_guard<std::mutex> lk(g_registry_mutex);
        g_live_scenes.erase(scene);
    }
    int cur = -1;
    if (cudaGetDevice(&cur) == cudaSuccess) {
        if (cur != scene->device) cudaSetDevice(scene->device);
        cudaDeviceSynchronize();
        for (int i = 0; i < N_FRAME_KERNELS + 1; i++)
            if (scene->kev[i]) cudaEventDestroy(scene->kev[i]);
        for (draw_scene::WorkSet &ws : scene->sets) {
            if (ws.graph_exec) cudaGraphExecDestroy(ws.graph_exec);
            if (ws.stream) cudaStreamDestroy(ws.stream);
            if (ws.h_uniforms) cudaFreeHost(ws.h_uniforms);
            if (ws.geo_done) cudaEventDestroy(ws.geo_done);
            if (ws.canvas_ready) cudaEventDestroy(ws.canvas_ready);
            if (ws.frame_done) cudaEventDestroy(ws.frame_done);
        }
    }
    delete scene;
}

int draw_scene_add_object(draw_scene *scene, const draw_object_desc *desc, uint32_t *out_id) {
    GUARD_BEGIN
    if (!scene || !desc) return fail(DRAW_ERR_INVALID_ARGUMENT, "scene or desc is NULL");
    if ((desc->n_positions && !desc->positions) || (desc->n_normals && !desc->normals) || (desc->n_uvs && !desc->uvs) ||
        (desc->n_meshes && !desc->meshes) || (desc->n_materials && !desc->materials))
        return fail(DRAW_ERR_INVALID_ARGUMENT, "NULL array with non-zero count");
    HostObject o;
    o.name = desc->name ? desc->name : "";
    o.pos.assign(desc->positions, desc->positions + 3 * desc->n_positions);
    o.nrm.assign(desc->normals, desc->normals + 3 * desc->n_normals);
    o.uv.assign(desc->uvs, desc->uvs + 3 * desc->n_uvs);
    // texel block of this object; offset 0 holds the 1x1 white default map (scene/mod.rs:128-135)
    o.texels = {255, 255, 255, 255};
    // maps are stored with 4 bytes per texel (r, g, b, x): 3-component images are padded
    auto add_map = [&](const draw_texture_map &m, uint32_t &off, uint32_t &w, uint32_t &h) -> int {
        if (!m.pixels) {
            off = 0; w = 1; h = 1;
            return DRAW_OK;
        }
        if (m.width == 0 || m.height == 0 || (m.components != 3 && m.components != 4))
            return fail(DRAW_ERR_INVALID_ARGUMENT, "texture map must be non-empty with 3 or 4 components");
        const size_t n = (size_t)m.width * m.height;
        off = (uint32_t)o.texels.size();
        o.texels.resize(o.texels.size() + 4 * n);
        uint8_t *dst = o.texels.data() + off;
        if (m.components == 4) std::memcpy(dst, m.pixels, 4 * n);
        else
            for (size_t i = 0; i < n; i++) {
                dst[4 * i] = m.pixels[3 * i]; dst[4 * i + 1] = m.pixels[3 * i + 1];
                dst[4 * i + 2] = m.pixels[3 * i + 2]; dst[4 * i + 3] = 255;
            }
        w = m.width; h = m.height;
        return DRAW_OK;
    };
    struct SeenMap { const uint8_t *pixels; uint32_t w, h, comp, off; };
    std::vector<SeenMap> seen; // identical images (map_Ka == map_Kd is common) share storage
    for (size_t i = 0; i < desc->n_materials; i++) {
        const draw_material &m = desc->materials[i];
        MaterialDev d{};
        for (int c = 0; c < 3; c++) { d.ka[c] = m.ka[c]; d.kd[c] = m.kd[c]; d.ks[c] = m.ks[c]; }
        d.alpha = m.alpha;
        auto add_dedup = [&](const draw_texture_map &tm, uint32_t &off, uint32_t &w, uint32_t &h) -> int {
            for (const SeenMap &sm : seen) // same pointer AND same layout: a sub-image or another component view is a different map
                if (tm.pixels && sm.pixels == tm.pixels && sm.w == tm.width && sm.h == tm.height && sm.comp == tm.components) {
                    off = sm.off; w = tm.width; h = tm.height;
                    return DRAW_OK;
                }
            TRY(add_map(tm, off, w, h));
            if (tm.pixels) seen.push_back({tm.pixels, tm.width, tm.height, tm.components, off});
            return DRAW_OK;
        };
        TRY(add_dedup(m.map_ka, d.ka_off, d.ka_w, d.ka_h));
        TRY(add_dedup(m.map_kd, d.kd_off, d.kd_w, d.kd_h));
        o.materials.push_back(d);
    }
    for (size_t i = 0; i < desc->n_meshes; i++) {
        const draw_mesh &dm = desc->meshes[i];
        if (dm.n_triangles && !dm.triangles) return fail(DRAW_ERR_INVALID_ARGUMENT, "mesh %zu: NULL triangles", i);
        // Object::new indexes textures[texture_idx] (object.rs:46-48) and would panic
        if (dm.material_idx >= desc->n_materials)
            return fail(DRAW_ERR_INVALID_ARGUMENT, "mesh %zu: material index %u out of range (%zu materials)", i,
                        dm.material_idx, desc->n_materials);
        HostMesh hm;
        hm.material = dm.material_idx;
        hm.tris.assign(dm.triangles, dm.triangles + 9 * dm.n_triangles);
        for (size_t t = 0; t < dm.n_triangles; t++) {
            const uint32_t *p = &hm.tris[9 * t];
            for (int c = 0; c < 3; c++)
                if (p[c] >= desc->n_positions || p[3 + c] >= desc->n_uvs || p[6 + c] >= desc->n_normals)
                    return fail(DRAW_ERR_INVALID_ARGUMENT, "mesh %zu triangle %zu: index out of range", i, t); // mesh.rs:46-48
        }
        if (desc->materials[dm.material_idx].alpha < 1.0f) o.transparent.push_back(std::move(hm)); // object.rs:48-52
        else o.opaque.push_back(std::move(hm));
    }
    scene->objects.push_back(std::move(o));
    scene->geometry_dirty = true;
    scene->rec_cap = scene->refs_cap = 0; // re-derive capacities for the new triangle count
    scene->needs_prepare = true;
    if (out_id) *out_id = (uint32_t)scene->objects.size() - 1;
    return DRAW_OK;
    GUARD_END
}

int draw_scene_set_camera(draw_scene *scene, const float pos[3], const float dir[3]) {
    if (!scene || !pos || !dir) return fail(DRAW_ERR_INVALID_ARGUMENT, "NULL argument");
    const float ratio = (float)scene->width / (float)scene->height; // scene/mod.rs:768
    scene->camera = CameraState::make(f3{pos[0], pos[1], pos[2]}, f3{dir[0], dir[1], dir[2]}, ratio);
    return DRAW_OK;
}

int draw_scene_get_camera(const draw_scene *scene, float pos[3], float dir[3]) {
    if (!scene || !pos || !dir) return fail(DRAW_ERR_INVALID_ARGUMENT, "NULL argument");
    pos[0] = scene->camera.position.x; pos[1] = scene->camera.position.y; pos[2] = scene->camera.position.z;
    dir[0] = scene->camera.direction.x; dir[1] = scene->camera.direction.y; dir[2] = scene->camera.direction.z;
    return DRAW_OK;
}

int draw_scene_set_camera_pos(draw_scene *scene, const float pos[3]) {
    if (!scene || !pos) return fail(DRAW_ERR_INVALID_ARGUMENT, "NULL argument");
    scene->camera.position = f3{pos[0], pos[1], pos[2]};
    return DRAW_OK;
}

int draw_scene_camera_move(draw_scene *scene, draw_camera_dir dir, float dist) {
    if (!scene) return fail(DRAW_ERR_INVALID_ARGUMENT, "scene is NULL");
    CameraState &c = scene->camera;
    switch (dir) { // scene/mod.rs:381-405
    case DRAW_CAMERA_UP: c.position = add(c.position, scale(c.up, dist)); break;
    case DRAW_CAMERA_DOWN: c.position = add(c.position, scale(c.up, -dist)); break;
    case DRAW_CAMERA_LEFT: c.position = add(c.position, scale(c.u, -dist)); break;
    case DRAW_CAMERA_RIGHT: c.position = add(c.position, scale(c.u, dist)); break;
    case DRAW_CAMERA_FOWARD: c.position = add(c.position, scale(unit(cross3(c.up, c.u)), dist)); break;
    case DRAW_CAMERA_BACKWARD: c.position = add(c.position, scale(unit(cross3(c.u, c.up)), dist)); break;
    default: return fail(DRAW_ERR_INVALID_ARGUMENT, "unknown camera direction %d", (int)dir);
    }
    return DRAW_OK;
}

int draw_scene_move_camera_direction(draw_scene *scene, int dx, int dy) {
    if (!scene) return fail(DRAW_ERR_INVALID_ARGUMENT, "scene is NULL");
    // the reference asserts dx < width and dy < height (scene/mod.rs:804-805)
    if (dx >= (long long)scene->width || dy >= (long long)scene->height)
        return fail(DRAW_ERR_INVALID_ARGUMENT, "dx/dy must be smaller than the scene size");
    CameraState &c = scene->camera;
    const float fx = (float)dx / (float)scene->width, fy = (float)dy / (float)scene->height;
    c.direction = unit(add(add(c.direction, scale(c.u, fx)), scale(c.v, fy))); // :430-432
    c.update_basis();
    return DRAW_OK;
}

int draw_scene_set_light(draw_scene *scene, const float pos[3]) {
    if (!scene || !pos) return fail(DRAW_ERR_INVALID_ARGUMENT, "NULL argument");
    scene->light = f3{pos[0], pos[1], pos[2]};
    return DRAW_OK;
}

int draw_scene_render(draw_scene *scene, draw_canvas *canvas) {
    GUARD_BEGIN
    if (!scene || !canvas) return fail(DRAW_ERR_INVALID_ARGUMENT, "scene or canvas is NULL");
    if (scene->device != canvas->device)
        return fail(DRAW_ERR_INVALID_ARGUMENT, "scene (device %d) and canvas (device %d) live on different devices",
                    scene->device, canvas->device);
    if (!canvas->has_depth) return fail(DRAW_ERR_INVALID_ARGUMENT, "Depth not initialized"); // canvas.rs:914
    if (!canvas->pending.empty()) {
        // Settle the earlier frames of this canvas that have completed (no wait): a frame that overflowed a work buffer
        // is re-rendered, with everything enqueued after it, before the new frame goes in.  A canvas has STATUS_SLOTS
        // status blocks: with that many frames in flight on it, the oldest is waited for.
        TRY(ensure_device(canvas->device));
        while (!canvas->pending.empty()) {
            const draw_canvas::FrameInputs &front = canvas->pending.front();
            if ((int)canvas->pending.size() >= draw_canvas::STATUS_SLOTS - 1) CU(cudaEventSynchronize(canvas->slot_event[front.status_slot]));
            else if (cudaEventQuery(canvas->slot_event[front.status_slot]) != cudaSuccess) break;
            if (canvas->h_status[(size_t)front.status_slot * N_STATUS_WORDS + CNT_OVERFLOW]) {
                TRY(finish_frame(canvas));
                break;
            }
            record_stats(canvas, front);
            canvas->pending.erase(canvas->pending.begin());
        }
        cudaGetLastError(); // cudaErrorNotReady is not sticky, but keep the error state clean
    } else {
        canvas->stats.overflow = 0; // of the frames enqueued from here on
    }
    return enqueue_frame(scene, canvas);
    GUARD_END
}

int draw_scene_prepare(draw_scene *scene, draw_canvas *canvas) {
    GUARD_BEGIN
    if (!scene || !canvas) return fail(DRAW_ERR_INVALID_ARGUMENT, "scene or canvas is NULL");
    if (scene->device != canvas->device)
        return fail(DRAW_ERR_INVALID_ARGUMENT, "scene (device %d) and canvas (device %d) live on different devices",
                    scene->device, canvas->device);
    TRY(ensure_device(scene->device));
    if (scene->geometry_dirty) TRY(upload_geometry(scene));
    FrameUniforms U{};
    fill_uniforms(scene, canvas, U);
    if (U.tiles_x >= MAX_TILES_X || U.tiles_y >= MAX_TILES_Y)
        return fail(DRAW_ERR_INVALID_ARGUMENT, "canvas %zux%zu is too large for the tile work list", canvas->width, canvas->height);
    set_launch_grids(scene);
    const bool want_graph = g_cfg.graphs && !scene->kernel_timing && !scene->debug_tile_cycles;
    for (int i = 0; i < scene->n_sets; i++) {
        if (scene->sets[i].frame_pending) CU(cudaEventSynchronize(scene->sets[i].frame_done));
        TRY(ensure_set_ready(scene, scene->sets[i], U, want_graph));
    }
    scene->needs_prepare = false;
    CU(cudaDeviceSynchronize()); // key pages are filled, graphs uploaded: the next render only enqueues
    return DRAW_OK;
    GUARD_END
}

int draw_scene_get_uniforms(draw_scene *scene, float matrix[16], float planes[24]) {
    if (!scene || !matrix || !planes) return fail(DRAW_ERR_INVALID_ARGUMENT, "NULL argument");
    const m4 m = transformation_matrix(scene->camera, scene->width, scene->height);
    std::memcpy(matrix, m.v, 16 * sizeof(float));
    plane4 p[6];
    scene->camera.view_planes(p);
    for (int i = 0; i < 6; i++) {
        planes[4 * i] = p[i].nx; planes[4 * i + 1] = p[i].ny; planes[4 * i + 2] = p[i].nz; planes[4 * i + 3] = p[i].k;
    }
    return DRAW_OK;
}

int draw_scene_read_vertex_visual(draw_scene *scene, draw_canvas *canvas, size_t first, size_t count, float *out) {
    GUARD_BEGIN
    if (!scene || !canvas || !out) return fail(DRAW_ERR_INVALID_ARGUMENT, "NULL argument");
    TRY(finish_frame(canvas));
    if (first + count > scene->dev.n_vertices) return fail(DRAW_ERR_INVALID_ARGUMENT, "vertex range out of bounds");
    const draw_scene::WorkSet &ws = scene->sets[scene->last_set];
    std::vector<float4> a(count), lh(2 * count);
    if (count) {
        CU(cudaMemcpy(a.data(), ws.vA.ptr + first, count * sizeof(float4), cudaMemcpyDeviceToHost));
        CU(cudaMemcpy(lh.data(), ws.vLH.ptr + 2 * first, 2 * count * sizeof(float4), cudaMemcpyDeviceToHost));
    }
    for (size_t i = 0; i < count; i++) {
        float *o = out + 7 * i;
        o[0] = lh[2 * i].x; o[1] = lh[2 * i].y; o[2] = lh[2 * i].z;         // light
        o[3] = lh[2 * i].w; o[4] = lh[2 * i + 1].x; o[5] = lh[2 * i + 1].y; // halfway
        o[6] = a[i].z;                                                       // depth
    }
    return DRAW_OK;
    GUARD_END
}

int draw_scene_counts(const draw_scene *scene, size_t *n_objects, size_t *n_triangles, size_t *n_vertices) {
    if (!scene) return fail(DRAW_ERR_INVALID_ARGUMENT, "scene is NULL");
    size_t t = 0, v = 0;
    for (const HostObject &o : scene->objects) {
        v += o.pos.size() / 3;
        for (const HostMesh &m : o.opaque) t += m.tris.size() / 9;
        for (const HostMesh &m : o.transparent) t += m.tris.size() / 9;
    }
    if (n_objects) *n_objects = scene->objects.size();
    if (n_triangles) *n_triangles = t;
    if (n_vertices) *n_vertices = v;
    return DRAW_OK;
}

int draw_scene_launch_count(const draw_scene *scene, uint64_t *out) {
    if (!scene || !out) return fail(DRAW_ERR_INVALID_ARGUMENT, "NULL argument");
    *out = scene->launches;
    return DRAW_OK;
}

int draw_scene_set_kernel_timing(draw_scene *scene, int enabled) {
    if (!scene) return fail(DRAW_ERR_INVALID_ARGUMENT, "scene is NULL");
    scene->kernel_timing = enabled != 0;
    if (!enabled) scene->kev_recorded = false;
    return DRAW_OK;
}

int draw_scene_last_kernel_times(draw_scene *scene, draw_canvas *canvas, float ms[4]) {
    GUARD_BEGIN
    if (!scene || !canvas || !ms) return fail(DRAW_ERR_INVALID_ARGUMENT, "NULL argument");
    if (!scene->kev_recorded) return fail(DRAW_ERR_INVALID_ARGUMENT, "kernel timing was not enabled for the last frame");
    TRY(finish_frame(canvas));
    for (int i = 0; i < N_FRAME_KERNELS; i++) CU(cudaEventElapsedTime(&ms[i], scene->kev[i], scene->kev[i + 1])); // k_sort_transparent k_front k_raster k_tile
    return DRAW_OK;
    GUARD_END
}

int draw_scene_debug_tile_cycles(draw_scene *scene, draw_canvas *canvas, int enable, uint32_t *out, size_t n) {
    GUARD_BEGIN
    if (!scene) return fail(DRAW_ERR_INVALID_ARGUMENT, "NULL argument");
    scene->debug_tile_cycles = enable != 0;
    if (!out) return DRAW_OK;
    if (!canvas) return fail(DRAW_ERR_INVALID_ARGUMENT, "NULL argument");
    TRY(finish_frame(canvas));
    DevBuf<uint32_t> &cyc = scene->sets[scene->last_set].tile_cycles;
    if (!cyc.ptr || n > cyc.cap) return fail(DRAW_ERR_INVALID_ARGUMENT, "no tile cycles recorded");
    // layout: [tile] whole item, [n_tiles + tile] end of phase A, [2 n_tiles + tile] end of phase C, [3 n_tiles + tile] end of phase D
    CU(cudaMemcpy(out, cyc.ptr, n * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    return DRAW_OK;
    GUARD_END
}

int draw_scene_debug_trace(draw_scene *scene, int enable, uint32_t *out, size_t cap_records, size_t *n_records) {
    GUARD_BEGIN
    if (!scene) return fail(DRAW_ERR_INVALID_ARGUMENT, "NULL argument");
    TRY(ensure_device(scene->device));
    CU(cudaDeviceSynchronize());
    if (out && n_records) {
        *n_records = 0;
        if (scene->d_trace_count.ptr) {
            uint32_t n = 0;
            CU(cudaMemcpy(&n, scene->d_trace_count.ptr, sizeof n, cudaMemcpyDeviceToHost));
            const size_t m = std::min<size_t>({(size_t)n, scene->d_trace.cap, cap_records});
            if (m) CU(cudaMemcpy(out, scene->d_trace.ptr, m * sizeof(uint4), cudaMemcpyDeviceToHost));
            *n_records = m;
        }
    }
    scene->debug_trace = enable != 0;
    if (enable) {
        TRY(scene->d_trace.reserve(1u << 18));
        TRY(scene->d_trace_count.reserve(1));
        CU(cudaMemset(scene->d_trace_count.ptr, 0, sizeof(uint32_t)));
    }
    return DRAW_OK;
    GUARD_END
}

int draw_scene_debug_list_counts(draw_scene *scene, draw_canvas *canvas, uint32_t *out, size_t n, size_t *n_coarse) {
    GUARD_BEGIN
    if (!scene || !canvas) return fail(DRAW_ERR_INVALID_ARGUMENT, "NULL argument");
    TRY(finish_frame(canvas));
    const size_t tiles_x = (canvas->width + TILE_W - 1) / TILE_W, tiles_y = (canvas->height + TILE_H - 1) / TILE_H;
    const size_t coarse = tiles_x * tiles_y;
    if (n_coarse) *n_coarse = coarse;
    if (!out) return DRAW_OK;
    if (n != 3 * coarse) return fail(DRAW_ERR_INVALID_ARGUMENT, "n must be %zu", 3 * coarse);
    const draw_scene::WorkSet &ws = scene->sets[scene->last_set];
    CU(cudaMemcpy(out, ws.l_count.ptr, coarse * sizeof(uint32_t), cudaMemcpyDeviceToHost));              // large references
    CU(cudaMemcpy(out + coarse, ws.ms_weight.ptr, coarse * sizeof(uint32_t), cudaMemcpyDeviceToHost));   // medium / small weight
    CU(cudaMemcpy(out + 2 * coarse, ws.t_count.ptr, coarse * sizeof(uint32_t), cudaMemcpyDeviceToHost)); // transparent references
    return DRAW_OK;
    GUARD_END
}

// ---- Canvas --------------------------------------------------------------------------------

int draw_canvas_create(size_t width, size_t height, draw_canvas **out) {
    GUARD_BEGIN
    if (!out) return fail(DRAW_ERR_INVALID_ARGUMENT, "out is NULL");
    *out = nullptr;
    if (width == 0 || height == 0 || width > 65535 || height > 65535)
        return fail(DRAW_ERR_INVALID_ARGUMENT, "canvas size must be in 1..65535");
    int dev = 0;
    TRY(pick_device(&dev));
    draw_canvas *c = new draw_canvas();
    c->device = dev;
    c->width = width;
    c->height = height;
    auto cleanup = [&](int rc) {
        draw_canvas_destroy(c);
        return rc;
    };
    if (cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking) != cudaSuccess)
        return cleanup(fail(DRAW_ERR_CUDA, "cudaStreamCreate failed"));
    c->stream = c->own_stream;
    if (cudaMallocHost(&c->h_status, draw_canvas::STATUS_SLOTS * N_STATUS_WORDS * sizeof(uint32_t)) != cudaSuccess)
        return cleanup(fail(DRAW_ERR_OUT_OF_MEMORY, "cudaMallocHost failed"));
    std::memset(c->h_status, 0, draw_canvas::STATUS_SLOTS * N_STATUS_WORDS * sizeof(uint32_t));
    int rc = c->d_color.reserve(width * height * 4);
    if (rc) return cleanup(rc);
    rc = fill_color_black(c, 0, width * height); // vec![Pixel::black(); len], canvas.rs:368
    if (rc) return cleanup(rc);
    *out = c;
    return DRAW_OK;
    GUARD_END
}

void draw_canvas_destroy(draw_canvas *canvas) {
    if (!canvas) return;
    int cur = -1;
    if (cudaGetDevice(&cur) == cudaSuccess) {
        if (cur != canvas->device) cudaSetDevice(canvas->device);
        if (canvas->stream) cudaStreamSynchronize(canvas->stream);
        if (canvas->join_event) cudaEventDestroy(canvas->join_event);
        for (cudaEvent_t e : canvas->slot_event)
            if (e) cudaEventDestroy(e);
        if (canvas->own_stream) cudaStreamDestroy(canvas->own_stream);
        if (canvas->h_color) cudaFreeHost(canvas->h_color);
        if (canvas->h_status) cudaFreeHost(canvas->h_status);
    }
    delete canvas;
}

int draw_canvas_init_depth(draw_canvas *canvas, float depth) {
    GUARD_BEGIN
    if (!canvas) return fail(DRAW_ERR_INVALID_ARGUMENT, "canvas is NULL");
    TRY(ensure_device(canvas->device));
    canvas->depth_max = depth;
    TRY(canvas->d_depth.reserve(canvas->width * canvas->height));
    canvas->has_depth = true;
    uint32_t bits;
    std::memcpy(&bits, &depth, 4);
    CU(launch_fill_u32(reinterpret_cast<uint32_t *>(canvas->depth()), canvas->width * canvas->height, bits,
                       canvas->stream, &canvas->launches));
    return DRAW_OK;
    GUARD_END
}

int draw_canvas_apply_offset(draw_canvas *canvas, int x, int y) {
    if (!canvas) return fail(DRAW_ERR_INVALID_ARGUMENT, "canvas is NULL");
    canvas->off_x = x;
    canvas->off_y = y;
    return DRAW_OK;
}

int draw_canvas_resize(draw_canvas *canvas, size_t width, size_t height) {
    GUARD_BEGIN
    if (!canvas) return fail(DRAW_ERR_INVALID_ARGUMENT, "canvas is NULL");
    if (width == 0 || height == 0 || width > 65535 || height > 65535)
        return fail(DRAW_ERR_INVALID_ARGUMENT, "canvas size must be in 1..65535");
    if (canvas->ext_color || canvas->ext_depth)
        return fail(DRAW_ERR_INVALID_ARGUMENT, "cannot resize a canvas bound to external memory");
    TRY(finish_frame(canvas));
    // Vec::resize on the flat pixel vector (canvas.rs:390): keep the first min(old,new) pixels
    const size_t old_n = canvas->width * canvas->height, new_n = width * height;
    DevBuf<uint8_t> fresh;
    TRY(fresh.reserve(new_n * 4));
    const size_t keep = std::min(old_n, new_n);
    CU(cudaMemcpyAsync(fresh.ptr, canvas->d_color.ptr, keep * 4, cudaMemcpyDeviceToDevice, canvas->stream));
    CU(cudaStreamSynchronize(canvas->stream));
    std::swap(fresh.ptr, canvas->d_color.ptr);
    std::swap(fresh.cap, canvas->d_color.cap);
    canvas->width = width;
    canvas->height = height;
    TRY(fill_color_black(canvas, keep, new_n - keep));
    canvas->stripe_y0 = canvas->stripe_y1 = 0;
    canvas->row_step = 1;
    canvas->row_phase = 0;
    touch_frame(canvas);
    // self.init_depth(self.depth_max), canvas.rs:392 — allocates the depth buffer even if none existed
    return draw_canvas_init_depth(canvas, canvas->depth_max);
    GUARD_END
}

int draw_canvas_clear(draw_canvas *canvas) {
    GUARD_BEGIN
    if (!canvas) return fail(DRAW_ERR_INVALID_ARGUMENT, "canvas is NULL");
    TRY(ensure_device(canvas->device));
    CU(launch_clear(canvas->color(), canvas->has_depth ? canvas->depth() : nullptr, canvas->width * canvas->height,
                    canvas->depth_max, canvas->stream, &canvas->launches));
    touch_frame(canvas);
    return DRAW_OK;
    GUARD_END
}

int draw_canvas_enable_depth_update(draw_canvas *canvas) {
    if (!canvas) return fail(DRAW_ERR_INVALID_ARGUMENT, "canvas is NULL");
    canvas->depth_update = true;
    return DRAW_OK;
}
int draw_canvas_disable_depth_update(draw_canvas *canvas) {
    if (!canvas) return fail(DRAW_ERR_INVALID_ARGUMENT, "canvas is NULL");
    canvas->depth_update = false;
    return DRAW_OK;
}

// ------------------------------------------------------------------------------------------
// Canvas::draw_triangle (canvas.rs:435-575)
// ------------------------------------------------------------------------------------------
int draw_texture_create(const draw_texture_map *map_kd, draw_texture **out) {
    GUARD_BEGIN
    if (!map_kd || !out) return fail(DRAW_ERR_INVALID_ARGUMENT, "NULL argument");
    // get_rgba_slice (scene/mod.rs:137-152) requires four components
    if (!map_kd->pixels || map_kd->components != 4 || map_kd->width == 0 || map_kd->height == 0)
        return fail(DRAW_ERR_INVALID_ARGUMENT, "draw_texture_create needs an RGBA map (components == 4) with pixels");
    if (g_device >= 0) TRY(ensure_device(g_device));
    std::unique_ptr<draw_texture> t(new draw_texture);
    CU(cudaGetDevice(&t->device));
    t->width = map_kd->width;
    t->height = map_kd->height;
    const size_t bytes = (size_t)t->width * t->height * 4;
    TRY(t->texels.reserve(bytes));
    CU(cudaMemcpy(t->texels.ptr, map_kd->pixels, bytes, cudaMemcpyHostToDevice));
    *out = t.release();
    return DRAW_OK;
    GUARD_END
}

void draw_texture_destroy(draw_texture *texture) {
    if (!texture) return;
    cudaSetDevice(texture->device);
    delete texture;
}

int draw_canvas_draw_commands(draw_canvas *canvas, const draw_vertex2d *vertices, size_t n_triangles, const draw_command2d *commands,
                              size_t n_commands, const draw_texture *texture) {
    GUARD_BEGIN
    if (!canvas || !texture || (!vertices && n_triangles) || (!commands && n_commands)) return fail(DRAW_ERR_INVALID_ARGUMENT, "NULL argument");
    if (!canvas->has_depth) return fail(DRAW_ERR_INVALID_ARGUMENT, "Depth not initialized"); // get_pixel_depth indexes an empty Vec
    if (texture->device != canvas->device) return fail(DRAW_ERR_INVALID_ARGUMENT, "texture and canvas live on different devices");
    // the commands own consecutive triangle ranges, in submission order
    std::vector<unsigned long long> first(n_commands + 1, 0);
    std::vector<OverlayClip> clips(std::max<size_t>(n_commands, 1));
    for (size_t k = 0; k < n_commands; k++) {
        first[k + 1] = first[k] + commands[k].n_triangles;
        clips[k].has = commands[k].has_clip ? 1u : 0u;
        clips[k].c[0] = commands[k].clip.x0; clips[k].c[1] = commands[k].clip.y0;
        clips[k].c[2] = commands[k].clip.x1; clips[k].c[3] = commands[k].clip.y1;
    }
    if (first[n_commands] != n_triangles) return fail(DRAW_ERR_INVALID_ARGUMENT, "the commands' triangle counts must add up to n_triangles");
    if (n_triangles == 0) return DRAW_OK;
    TRY(ensure_device(canvas->device));
    // A frame still in flight may have to be rendered again (overflow); the triangles go on top of the settled frame.
    TRY(finish_frame(canvas));
    OverlayParams P{};
    P.width = (uint32_t)canvas->width;
    P.height = (uint32_t)canvas->height;
    P.bins_x = (P.width + OVERLAY_BIN - 1) / OVERLAY_BIN;
    P.bins_y = (P.height + OVERLAY_BIN - 1) / OVERLAY_BIN;
    P.texels = texture->texels.ptr;
    P.tex_w = texture->width;
    P.tex_h = texture->height;
    P.color = reinterpret_cast<uint32_t *>(canvas->color());
    P.depth = canvas->depth();
    P.depth_update = canvas->depth_update ? 1u : 0u;
    // the command table: [n_commands + 1] first-triangle numbers, then the clipping rectangles (pageable sources: the copies
    // have left these vectors when cudaMemcpyAsync returns)
    const size_t first_bytes = (n_commands + 1) * sizeof(unsigned long long);
    TRY(canvas->ov_cmds.reserve(first_bytes + n_commands * sizeof(OverlayClip)));
    CU(cudaMemcpyAsync(canvas->ov_cmds.ptr, first.data(), first_bytes, cudaMemcpyHostToDevice, canvas->stream));
    CU(cudaMemcpyAsync(canvas->ov_cmds.ptr + first_bytes, clips.data(), n_commands * sizeof(OverlayClip), cudaMemcpyHostToDevice, canvas->stream));
    P.cmd_first = reinterpret_cast<const unsigned long long *>(canvas->ov_cmds.ptr);
    P.cmd_clip = reinterpret_cast<const OverlayClip *>(canvas->ov_cmds.ptr + first_bytes);
    P.n_cmds = (uint32_t)n_commands;
    const size_t n_bins = (size_t)P.bins_x * P.bins_y;
    for (size_t at = 0; at < n_triangles; at += OVERLAY_MAX_BATCH) { // batches keep submission order on the stream
        const uint32_t n = (uint32_t)std::min<size_t>(OVERLAY_MAX_BATCH, n_triangles - at);
        P.n = n;
        P.first = (uint32_t)at;
        P.words = (n + 31) / 32;
        TRY(canvas->ov_verts.reserve((size_t)n * 3 * sizeof(draw_vertex2d)));
        TRY(canvas->ov_recs.reserve((size_t)n * OVERLAY_REC_BYTES));
        TRY(canvas->ov_masks.reserve(n_bins * P.words));
        TRY(canvas->ov_any.reserve(n_bins));
        CU(cudaMemcpyAsync(canvas->ov_verts.ptr, vertices + at * 3, (size_t)n * 3 * sizeof(draw_vertex2d), cudaMemcpyHostToDevice,
                           canvas->stream));
        P.verts = canvas->ov_verts.ptr;
        P.recs = canvas->ov_recs.ptr;
        P.masks = canvas->ov_masks.ptr;
        P.bin_any = canvas->ov_any.ptr;
        CU(launch_overlay(P, canvas->stream, &canvas->launches));
    }
    touch_frame(canvas);
    return DRAW_OK;
    GUARD_END
}

int draw_canvas_draw_triangles(draw_canvas *canvas, const draw_vertex2d *vertices, size_t n_triangles, const draw_texture *texture,
                               const draw_rect *clipping_rect) {
    draw_command2d cmd{};
    cmd.n_triangles = n_triangles;
    cmd.has_clip = clipping_rect ? 1 : 0;
    if (clipping_rect) cmd.clip = *clipping_rect;
    return draw_canvas_draw_commands(canvas, vertices, n_triangles, &cmd, 1, texture);
}

int draw_canvas_size(const draw_canvas *canvas, size_t *width, size_t *height) {
    if (!canvas) return fail(DRAW_ERR_INVALID_ARGUMENT, "canvas is NULL");
    if (width) *width = canvas->width;
    if (height) *height = canvas->height;
    return DRAW_OK;
}

int draw_canvas_map_host(draw_canvas *canvas, const uint8_t **out_bytes, size_t *out_len) {
    GUARD_BEGIN
    if (!canvas || !out_bytes) return fail(DRAW_ERR_INVALID_ARGUMENT, "NULL argument");
    TRY(finish_frame(canvas));
    const size_t bytes = canvas->width * canvas->height * 4;
    TRY(ensure_host_mirror(canvas, bytes));
    if (canvas->host_dirty) {
        TRY(refresh_mirror(canvas, canvas->stream, nullptr, nullptr));
        CU(cudaStreamSynchronize(canvas->stream));
    }
    *out_bytes = canvas->h_color;
    if (out_len) *out_len = bytes;
    return DRAW_OK;
    GUARD_END
}

int draw_canvas_enable_host_mirror(draw_canvas *canvas, int enabled) {
    GUARD_BEGIN
    if (!canvas) return fail(DRAW_ERR_INVALID_ARGUMENT, "NULL argument");
    TRY(ensure_device(canvas->device));
    canvas->host_mirror = enabled != 0;
    if (canvas->host_mirror) TRY(ensure_host_mirror(canvas, canvas->width * canvas->height * 4));
    return DRAW_OK;
    GUARD_END
}

int draw_canvas_read_depth(draw_canvas *canvas, float *dst, size_t n_floats) {
    GUARD_BEGIN
    if (!canvas || !dst) return fail(DRAW_ERR_INVALID_ARGUMENT, "NULL argument");
    if (!canvas->has_depth) return fail(DRAW_ERR_INVALID_ARGUMENT, "Depth not initialized");
    if (n_floats != canvas->width * canvas->height)
        return fail(DRAW_ERR_INVALID_ARGUMENT, "n_floats must be width*height");
    TRY(finish_frame(canvas));
    CU(cudaMemcpyAsync(dst, canvas->depth(), n_floats * sizeof(float), cudaMemcpyDeviceToHost, canvas->stream));
    CU(cudaStreamSynchronize(canvas->stream));
    return DRAW_OK;
    GUARD_END
}

int draw_canvas_sync(draw_canvas *canvas) {
    GUARD_BEGIN
    if (!canvas) return fail(DRAW_ERR_INVALID_ARGUMENT, "canvas is NULL");
    return finish_frame(canvas);
    GUARD_END
}

int draw_canvas_last_frame_stats(draw_canvas *canvas, draw_frame_stats *out) {
    GUARD_BEGIN
    if (!canvas || !out) return fail(DRAW_ERR_INVALID_ARGUMENT, "NULL argument");
    TRY(finish_frame(canvas));
    *out = canvas->stats;
    return DRAW_OK;
    GUARD_END
}

int draw_canvas_device_ptrs(draw_canvas *canvas, void **out_color, void **out_depth) {
    if (!canvas) return fail(DRAW_ERR_INVALID_ARGUMENT, "canvas is NULL");
    if (out_color) *out_color = canvas->color();
    if (out_depth) *out_depth = canvas->has_depth ? canvas->depth() : nullptr;
    return DRAW_OK;
}

int draw_canvas_bind_external(draw_canvas *canvas, void *color_dev, void *depth_dev) {
    GUARD_BEGIN
    if (!canvas) return fail(DRAW_ERR_INVALID_ARGUMENT, "canvas is NULL");
    TRY(finish_frame(canvas));
    canvas->ext_color = static_cast<uint8_t *>(color_dev);
    canvas->ext_depth = static_cast<float *>(depth_dev);
    if (depth_dev) canvas->has_depth = true;
    touch_frame(canvas);
    return DRAW_OK;
    GUARD_END
}

int draw_canvas_ipc_export(draw_canvas *canvas, uint8_t handle[64]) {
    GUARD_BEGIN
    if (!canvas || !handle) return fail(DRAW_ERR_INVALID_ARGUMENT, "NULL argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
    TRY(ensure_device(canvas->device));
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, canvas->d_color.ptr));
    std::memcpy(handle, &h, 64);
    return DRAW_OK;
    GUARD_END
}

int draw_ipc_open(const uint8_t handle[64], void **out_dev_ptr) {
    GUARD_BEGIN
    if (!handle || !out_dev_ptr) return fail(DRAW_ERR_INVALID_ARGUMENT, "NULL argument");
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle, 64);
    CU(cudaIpcOpenMemHandle(out_dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return DRAW_OK;
    GUARD_END
}

int draw_ipc_close(void *dev_ptr) {
    GUARD_BEGIN
    if (!dev_ptr) return DRAW_OK;
    CU(cudaIpcCloseMemHandle(dev_ptr));
    return DRAW_OK;
    GUARD_END
}

int draw_device_alloc(size_t bytes, void **out_dev_ptr) {
    GUARD_BEGIN
    if (!out_dev_ptr || bytes == 0) return fail(DRAW_ERR_INVALID_ARGUMENT, "NULL argument or zero size");
    int dev = 0;
    TRY(pick_device(&dev));
    CU(cudaMalloc(out_dev_ptr, bytes));
    CU(cudaMemset(*out_dev_ptr, 0, bytes));
    CU(cudaDeviceSynchronize());
    return DRAW_OK;
    GUARD_END
}

int draw_device_free(void *dev_ptr) {
    GUARD_BEGIN
    if (dev_ptr) CU(cudaFree(dev_ptr));
    return DRAW_OK;
    GUARD_END
}

int draw_ipc_export(void *dev_ptr, uint8_t handle[64]) {
    GUARD_BEGIN
    if (!dev_ptr || !handle) return fail(DRAW_ERR_INVALID_ARGUMENT, "NULL argument");
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, dev_ptr));
    std::memcpy(handle, &h, 64);
    return DRAW_OK;
    GUARD_END
}

int draw_flag_signal(void *flag_dev, uint32_t value, draw_canvas *canvas) {
    GUARD_BEGIN
    if (!flag_dev || !canvas) return fail(DRAW_ERR_INVALID_ARGUMENT, "NULL argument");
    TRY(ensure_device(canvas->device));
    CU(launch_flag_signal(static_cast<uint32_t *>(flag_dev), value, canvas->stream));
    return DRAW_OK;
    GUARD_END
}

int draw_flags_wait(const void *flags_dev, uint32_t n_flags, uint32_t value, void *error_word_dev, draw_canvas *canvas) {
    GUARD_BEGIN
    if (!flags_dev || !canvas) return fail(DRAW_ERR_INVALID_ARGUMENT, "NULL argument");
    if (n_flags > 64) return fail(DRAW_ERR_INVALID_ARGUMENT, "at most 64 flags");
    TRY(ensure_device(canvas->device));
    CU(launch_flags_wait(static_cast<const uint32_t *>(flags_dev), n_flags, value, static_cast<uint32_t *>(error_word_dev), canvas->stream));
    return DRAW_OK;
    GUARD_END
}

int draw_canvas_set_stream(draw_canvas *canvas, void *cuda_stream) {
    GUARD_BEGIN
    if (!canvas) return fail(DRAW_ERR_INVALID_ARGUMENT, "canvas is NULL");
    TRY(finish_frame(canvas));
    canvas->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : canvas->own_stream;
    return DRAW_OK;
    GUARD_END
}

int draw_canvas_export_png(draw_canvas *canvas, const char *path) {
    GUARD_BEGIN
    if (!canvas || !path) return fail(DRAW_ERR_INVALID_ARGUMENT, "NULL argument");
    const uint8_t *bgra = nullptr;
    size_t len = 0;
    TRY(draw_canvas_map_host(canvas, &bgra, &len));
    // app/mod.rs:347-352: "reverse the RGB order" — swap bytes 0 and 2 of every pixel, keep the pad byte as alpha
    std::vector<uint8_t> rgba(len);
    for (size_t i = 0; i + 3 < len; i += 4) {
        rgba[i] = bgra[i + 2]; rgba[i + 1] = bgra[i + 1]; rgba[i + 2] = bgra[i]; rgba[i + 3] = bgra[i + 3];
    }
    return draw_image_write_png(path, rgba.data(), (uint32_t)canvas->width, (uint32_t)canvas->height, 4);
    GUARD_END
}

int draw_canvas_export_jpeg(draw_canvas *canvas, const char *path) {
    GUARD_BEGIN
    if (!canvas || !path) return fail(DRAW_ERR_INVALID_ARGUMENT, "NULL argument");
    const uint8_t *bgra = nullptr;
    size_t len = 0;
    TRY(draw_canvas_map_host(canvas, &bgra, &len));
    std::vector<uint8_t> rgba(len); // app/mod.rs:347-352, as in draw_canvas_export_png
    for (size_t i = 0; i + 3 < len; i += 4) {
        rgba[i] = bgra[i + 2]; rgba[i + 1] = bgra[i + 1]; rgba[i + 2] = bgra[i]; rgba[i + 3] = bgra[i + 3];
    }
    // write_img hands stbi_write_jpg `width * PIXEL_BYTES` where it takes the quality (app/mod.rs:371-377)
    const size_t quality = canvas->width * 4;
    return draw_image_write_jpg(path, rgba.data(), (uint32_t)canvas->width, (uint32_t)canvas->height, 4, (int)std::min<size_t>(quality, 1 << 20));
    GUARD_END
}

int draw_canvas_stream_wait(draw_canvas *canvas, void *cuda_stream) {
    GUARD_BEGIN
    if (!canvas) return fail(DRAW_ERR_INVALID_ARGUMENT, "canvas is NULL");
    TRY(ensure_device(canvas->device));
    cudaStream_t waiter = static_cast<cudaStream_t>(cuda_stream);
    if (waiter == canvas->stream) return DRAW_OK; // already ordered
    if (!canvas->join_event) CU(cudaEventCreateWithFlags(&canvas->join_event, cudaEventDisableTiming));
    CU(cudaEventRecord(canvas->join_event, canvas->stream));
    CU(cudaStreamWaitEvent(waiter, canvas->join_event, 0));
    return DRAW_OK;
    GUARD_END
}

int draw_canvas_set_stripe(draw_canvas *canvas, size_t y0, size_t y1) {
    if (!canvas) return fail(DRAW_ERR_INVALID_ARGUMENT, "canvas is NULL");
    if (y0 >= y1 || y1 > canvas->height) return fail(DRAW_ERR_INVALID_ARGUMENT, "stripe must satisfy y0 < y1 <= height");
    if (y0 % TILE_H || (y1 % TILE_H && y1 != canvas->height))
        return fail(DRAW_ERR_INVALID_ARGUMENT, "stripe bounds must be multiples of the tile height %d (or the canvas height)", TILE_H);
    if (y0 == 0 && y1 == canvas->height) canvas->stripe_y0 = canvas->stripe_y1 = 0;
    else {
        canvas->stripe_y0 = y0;
        canvas->stripe_y1 = y1;
    }
    return DRAW_OK;
}

int draw_canvas_set_empty_tile_color(draw_canvas *canvas, int enabled) {
    if (!canvas) return fail(DRAW_ERR_INVALID_ARGUMENT, "canvas is NULL");
    canvas->empty_tile_color = enabled != 0;
    return DRAW_OK;
}

int draw_canvas_set_tile_rows(draw_canvas *canvas, uint32_t phase, uint32_t step) {
    if (!canvas) return fail(DRAW_ERR_INVALID_ARGUMENT, "canvas is NULL");
    if (step == 0 || phase >= step) return fail(DRAW_ERR_INVALID_ARGUMENT, "tile rows must satisfy phase < step, step >= 1");
    canvas->row_step = step;
    canvas->row_phase = phase;
    return DRAW_OK;
}

} // extern "C"

namespace drawb200 {
int loader_fail(int code, const char *msg) { return fail(code, "%s", msg); }
} // namespace drawb200
