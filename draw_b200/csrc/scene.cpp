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
#include <mutex>
#include <new>
#include <set>
#include <string>
#include <vector>

#include "../../include/draw_b200.h"
#include "device_types.h"
#include "host_math.hpp"

namespace drawb200 {
thread_local int g_pdl_enabled = 1;
thread_local int g_kernel_priority_set = 0, g_kernel_priority = 0;
extern thread_local unsigned g_clip_ctas; // k_geometry.cu
extern thread_local unsigned g_bin_ctas;  // k_binning.cu
extern thread_local unsigned g_raster_ctas; // k_raster.cu
extern thread_local unsigned g_clear_ctas, g_tile_ctas;
// k_geometry.cu / k_binning.cu / k_tile.cu
void launch_vertex(const FrameUniforms &U, const FrameUniforms *dU, const SceneDev &S, const FrameDev &W, cudaStream_t stream);
void launch_setup(const FrameUniforms &U, const FrameUniforms *dU, const SceneDev &S, const FrameDev &W, cudaStream_t stream);
void launch_clip(const FrameUniforms &U, const FrameUniforms *dU, const SceneDev &S, const FrameDev &W, cudaStream_t stream);
void launch_bin_count(const FrameUniforms &U, const FrameUniforms *dU, const FrameDev &W, cudaStream_t stream);
void launch_alloc(const FrameUniforms &U, const FrameUniforms *dU, const FrameDev &W, cudaStream_t stream);
void launch_bin_fill(const FrameUniforms &U, const FrameUniforms *dU, const FrameDev &W, cudaStream_t stream);
void launch_raster(const FrameUniforms &U, const FrameUniforms *dU, const FrameDev &W, cudaStream_t stream);
cudaError_t launch_fill_u64(unsigned long long *dst, size_t n, unsigned long long value, cudaStream_t stream);
void launch_tile(const FrameUniforms &U, const FrameUniforms *dU, const SceneDev &S, const FrameDev &W, cudaStream_t stream);
void launch_clear_empty(const FrameUniforms &U, const FrameUniforms *dU, const FrameDev &W, cudaStream_t stream);
void launch_shade(const FrameUniforms &U, const FrameUniforms *dU, const SceneDev &S, const FrameDev &W, cudaStream_t stream);
void launch_sort_transparent(const FrameUniforms *dU, const SceneDev &S, const void *ranges, uint32_t n_ranges, uint32_t *keys0,
                             uint32_t *keys1, uint32_t *perm0, uint32_t *perm1, uint32_t *scratch, cudaStream_t stream); // k_sort.cu
cudaError_t launch_clear(uint8_t *color, float *depth, size_t n_pixels, float depth_max, cudaStream_t stream,
                         uint64_t *launches);
cudaError_t launch_fill_u32(uint32_t *dst, size_t n, uint32_t value, cudaStream_t stream, uint64_t *launches);
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
    uint32_t n_coarse, n_lists, tiles_x, tile_y_begin, tile_y_end, clear_ctas;
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
    DevBuf<float> d_pos[3], d_nrm[3], d_uv[2];
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

    // Per-frame work buffers, N_WORK_SETS deep, each with its own side stream: the vertex / setup /
    // binning kernels of the next frames (a chain of six small, latency-bound launches) run on side
    // streams while k_tile of frame k still runs on its canvas' stream.
    struct WorkSet {
        DevBuf<float> vert[9];
        DevBuf<uint32_t> flags, list_count, list_offset, refs, counters, tile_cycles, tile_cost, tile_order, empty_tiles, shade_tiles;
        DevBuf<unsigned long long> scan_desc;
        DevBuf<uint2> clip_queue;
        DevBuf<RasterRec> rrec, trrec;
        DevBuf<PrepRec> prep;
        DevBuf<uint2> m_refs, s_refs;
        DevBuf<uint32_t> tile_page;
        DevBuf<unsigned long long> key_pages;
        size_t pages_clean = 0; // pages [0, pages_clean) of key_pages.ptr are known to be empty
        DevBuf<ShadeRec> srec, tsrec;
        FrameDev work{};
        cudaStream_t stream = nullptr;   // the frame using this set runs here ...
        cudaStream_t aux_stream = nullptr; // ... and its k_clear_empty here, beside k_bin<fill> / k_raster / k_tile
        DevBuf<FrameUniforms> d_uniforms;
        FrameUniforms *h_uniforms = nullptr; // pinned staging of d_uniforms
        cudaGraphExec_t graph_exec = nullptr;
        GraphKey graph_key{};
        cudaEvent_t alloc_done = nullptr; // side stream: k_alloc has listed the frame's empty tiles and work items
        cudaEvent_t geo_done = nullptr;  // side stream: binning of the frame using this set has finished
        cudaEvent_t canvas_ready = nullptr; // canvas stream: the canvas of the frame using this set may be written
        cudaEvent_t clear_done = nullptr; // aux stream: k_clear_empty has finished
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
    // optional per-kernel timing (draw_scene_set_kernel_timing): 0..6 around the six side-stream kernels,
    // 7 / 8 around k_tile on the canvas stream
    bool kernel_timing = false;
    cudaEvent_t kev[N_FRAME_KERNELS + 4] = {};
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
    cudaStream_t own_stream = nullptr, stream = nullptr;
    size_t stripe_y0 = 0, stripe_y1 = 0; // rows; y1 == 0 means whole canvas
    uint32_t *h_status = nullptr;        // pinned: counters of the last frame
    cudaEvent_t join_event = nullptr;    // draw_canvas_stream_wait
    bool frame_pending = false;
    draw_scene *last_scene = nullptr;
    // what the pending frame was rendered with: an overflowed frame is rendered again from exactly these
    // inputs, whatever the scene's camera / light and the canvas' offset / stripe have become since
    struct FrameInputs {
        CameraState camera;
        f3 light;
        int off_x = 0, off_y = 0;
        size_t stripe_y0 = 0, stripe_y1 = 0;
        float depth_max = 0.0f;
    } pending_inputs;
    draw_frame_stats stats{};
    uint64_t launches = 0;

    uint8_t *color() const { return ext_color ? ext_color : d_color.ptr; }
    float *depth() const { return ext_depth ? ext_depth : d_depth.ptr; }
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
    int pdl = env_int("DRAW_B200_PDL", 3);        // programmatic dependent launch: 0 off, 1 early trigger, 2 late, 3 early for a lone frame
    int sets = std::min(std::max(env_int("DRAW_B200_SETS", 8), 1), 8); // frames in flight per scene
    int pages = std::max(0, env_int("DRAW_B200_PAGES", 16384));        // key pages per work set (0: k_raster off)
    int clear_ctas = std::max(1, env_int("DRAW_B200_CLEAR_CTAS", 148 * 4));
    int tile_ctas = std::max(1, env_int("DRAW_B200_TILE_CTAS", 148 * 3)); // persistent CTAs of k_tile: three per SM leave room for the other frames' kernels
    int split_min_cost = std::max(1, env_int("DRAW_B200_SPLIT_MIN_COST", TILE_SPLIT_MIN_COST));
    int split_div = std::min(std::max(1, env_int("DRAW_B200_SPLIT_DIV", TILE_SPLIT_DIV)), (int)TILE_EXTRA_ITEMS);
    int defer_max = std::max(0, env_int("DRAW_B200_DEFER_MAX", 0));  // k_shade takes tiles with fewer large references (0: off)
    int split_max = std::min(std::max(1, env_int("DRAW_B200_SPLIT_MAX", TILE_MAX_SPLIT)), (int)TILE_MAX_SPLIT);
    int bin_rpw = std::max(0, env_int("DRAW_B200_BIN_RPW", 8));   // k_bin: warp-per-record up to this many records per warp of the grid
    // grids of the geometry / binning kernels; 0 = by scene size (enqueue_frame): with several frames in flight a
    // kernel costs the pipeline its CTAs' residency, so small scenes get small grids
    int clip_ctas = std::max(0, env_int("DRAW_B200_CLIP_CTAS", 0));
    int bin_ctas = std::max(0, env_int("DRAW_B200_BIN_CTAS", 0));
    int raster_ctas = std::max(0, env_int("DRAW_B200_RASTER_CTAS", 0));
    int cost_shade = std::max(0, env_int("DRAW_B200_COST_SHADE", 0)); // k_alloc: cost of shading a covered tile (0: not counted)
    int kprio = env_int("DRAW_B200_KPRIO", 0); // 1: geometry / binning / k_raster launches get a higher priority than k_tile and k_clear_empty; 2: the reverse
    int skip = env_int("DRAW_B200_SKIP", 0); // timing experiments only: bit i set = kernel i of the frame is not launched (frames are then wrong)
    int clear_in_tile = env_int("DRAW_B200_CLEAR_IN_TILE", 2); // who writes the empty tiles: 0 k_clear_empty on its own stream; k_tile's CTAs 1 before / 2 after each raster item, 3 alternating
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
    std::vector<float> pos[3], nrm[3], uv[2];
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
        const uint32_t vbase = (uint32_t)pos[0].size(), nbase = (uint32_t)nrm[0].size(), tbase = (uint32_t)uv[0].size();
        const uint32_t mbase = (uint32_t)materials.size(), xbase = (uint32_t)texels.size();
        for (size_t i = 0; i < o.pos.size() / 3; i++)
            for (int c = 0; c < 3; c++) pos[c].push_back(o.pos[3 * i + c]);
        for (size_t i = 0; i < o.nrm.size() / 3; i++)
            for (int c = 0; c < 3; c++) nrm[c].push_back(o.nrm[3 * i + c]);
        for (size_t i = 0; i < o.uv.size() / 3; i++)
            for (int c = 0; c < 2; c++) uv[c].push_back(o.uv[3 * i + c]);
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
    if (mat.size() >= (1u << 30)) return fail(DRAW_ERR_INVALID_ARGUMENT, "too many triangles (%zu)", mat.size());

    for (int c = 0; c < 3; c++) TRY(s->d_pos[c].upload(pos[c]));
    for (int c = 0; c < 3; c++) TRY(s->d_nrm[c].upload(nrm[c]));
    for (int c = 0; c < 2; c++) TRY(s->d_uv[c].upload(uv[c]));
    for (int c = 0; c < 9; c++) TRY(s->d_idx[c].upload(idx[c]));
    TRY(s->d_mat.upload(mat));
    TRY(s->d_tslot.upload(tslot));
    TRY(s->d_materials.upload(materials));
    TRY(s->d_texels.upload(texels));
    CU(cudaDeviceSynchronize()); // the host vectors above are about to go away

    SceneDev &d = s->dev;
    d.px = s->d_pos[0].ptr; d.py = s->d_pos[1].ptr; d.pz = s->d_pos[2].ptr;
    d.nx = s->d_nrm[0].ptr; d.ny = s->d_nrm[1].ptr; d.nz = s->d_nrm[2].ptr;
    d.tu = s->d_uv[0].ptr; d.tv = s->d_uv[1].ptr;
    for (int c = 0; c < 9; c++) d.idx[c] = s->d_idx[c].ptr;
    d.tri_mat = s->d_mat.ptr;
    d.tri_tslot = any_transparent ? s->d_tslot.ptr : nullptr;
    d.materials = s->d_materials.ptr;
    d.texels = s->d_texels.ptr;
    d.n_vertices = (uint32_t)pos[0].size();
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

int ensure_work_buffers(draw_scene *s, draw_scene::WorkSet &ws, size_t n_lists) {
    const SceneDev &d = s->dev;
    for (int i = 0; i < 9; i++) TRY(ws.vert[i].reserve(d.n_vertices));
    TRY(ws.flags.reserve(d.n_vertices));
    if (s->rec_cap == 0) s->rec_cap = 2 * (size_t)d.n_triangles + 4096;
    if (s->refs_cap == 0) s->refs_cap = std::max<size_t>((size_t)1 << 22, 4 * (size_t)d.n_triangles);
    TRY(ws.rrec.reserve(s->rec_cap));
    TRY(ws.srec.reserve(s->rec_cap));
    TRY(ws.prep.reserve(s->rec_cap));
    TRY(ws.trrec.reserve(4 * (size_t)d.n_transparent));
    TRY(ws.tsrec.reserve(4 * (size_t)d.n_transparent));
    TRY(ws.list_count.reserve(n_lists));
    TRY(ws.list_offset.reserve(n_lists + 1));
    TRY(ws.refs.reserve(s->refs_cap));
    TRY(ws.m_refs.reserve(s->refs_cap));
    TRY(ws.s_refs.reserve(s->refs_cap));
    TRY(ws.tile_page.reserve(n_lists / LISTS_PER_TILE));
    {
        // key pages for k_raster: one per tile at most (DRAW_B200_PAGES caps the pool; 0 disables k_raster)
        const size_t want = std::min<size_t>(n_lists / LISTS_PER_TILE, (size_t)g_cfg.pages);
        const unsigned long long *before = ws.key_pages.ptr;
        TRY(ws.key_pages.reserve(want * TILE_W * TILE_H));
        if (ws.key_pages.ptr != before) ws.pages_clean = 0;
        if (want > ws.pages_clean) {
            cudaStream_t st = ws.stream ? ws.stream : 0;
            CU(launch_fill_u64(ws.key_pages.ptr + ws.pages_clean * TILE_W * TILE_H, (want - ws.pages_clean) * TILE_W * TILE_H, KEY_EMPTY, st));
            if (!ws.stream) CU(cudaStreamSynchronize(0));
            ws.pages_clean = want;
        }
        ws.work.page_cap = (uint32_t)want;
    }
    TRY(ws.counters.reserve(N_COUNTERS));
    TRY(ws.tile_cost.reserve(n_lists));
    TRY(ws.tile_order.reserve(n_lists + TILE_EXTRA_ITEMS));
    TRY(ws.empty_tiles.reserve(n_lists / LISTS_PER_TILE));
    TRY(ws.shade_tiles.reserve(n_lists / LISTS_PER_TILE));
    TRY(ws.scan_desc.reserve((size_t)d.n_triangles / 256 + 2));
    TRY(ws.clip_queue.reserve(d.n_triangles));
    FrameDev &w = ws.work;
    w.v_lx = ws.vert[0].ptr; w.v_ly = ws.vert[1].ptr; w.v_lz = ws.vert[2].ptr;
    w.v_hx = ws.vert[3].ptr; w.v_hy = ws.vert[4].ptr; w.v_hz = ws.vert[5].ptr;
    w.v_depth = ws.vert[6].ptr; w.v_sx = ws.vert[7].ptr; w.v_sy = ws.vert[8].ptr;
    w.v_flags = ws.flags.ptr;
    w.rrec = ws.rrec.ptr; w.srec = ws.srec.ptr; w.prep = ws.prep.ptr;
    w.t_rrec = ws.trrec.ptr; w.t_srec = ws.tsrec.ptr;
    w.list_count = ws.list_count.ptr; w.list_offset = ws.list_offset.ptr; w.list_refs = ws.refs.ptr;
    w.m_refs = ws.m_refs.ptr; w.s_refs = ws.s_refs.ptr;
    w.tile_page = ws.tile_page.ptr; w.key_pages = ws.key_pages.ptr;
    w.counters = ws.counters.ptr;
    w.tile_cost = ws.tile_cost.ptr;
    w.tile_order = ws.tile_order.ptr;
    w.empty_tiles = ws.empty_tiles.ptr;
    w.shade_tiles = ws.shade_tiles.ptr;
    w.scan_desc = ws.scan_desc.ptr;
    w.clip_queue = ws.clip_queue.ptr;
    w.rec_cap = (uint32_t)s->rec_cap;
    w.refs_cap = (uint32_t)s->refs_cap;
    w.tile_cycles = nullptr;
    if (s->debug_tile_cycles) {
        TRY(ws.tile_cycles.reserve(n_lists));
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
    if (!ws.alloc_done) CU(cudaEventCreateWithFlags(&ws.alloc_done, cudaEventDisableTiming));
    if (!ws.geo_done) CU(cudaEventCreateWithFlags(&ws.geo_done, cudaEventDisableTiming));
    if (!ws.canvas_ready) CU(cudaEventCreateWithFlags(&ws.canvas_ready, cudaEventDisableTiming));
    if (!ws.clear_done) CU(cudaEventCreateWithFlags(&ws.clear_done, cudaEventDisableTiming));
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

// The launches of one frame on the work set's streams (directly, or under stream capture).
int launch_frame(draw_scene *s, draw_scene::WorkSet &ws, const FrameUniforms &U, cudaStream_t side, cudaEvent_t *ev,
                 bool capturing) {
    const FrameUniforms *dU = ws.d_uniforms.ptr;
    // ws.canvas_ready is recorded on the canvas' stream outside the graph: an external event of the capture
    const unsigned ext = capturing ? cudaEventWaitExternal : 0u;
    CU(cudaMemcpyAsync(ws.d_uniforms.ptr, ws.h_uniforms, sizeof(FrameUniforms), cudaMemcpyHostToDevice, side));
    // painter sort of the transparent meshes, in place in the shared index streams (enqueue_frame has ordered
    // this stream after the previous frame's geometry, which reads them)
    if (s->dev.n_transparent)
        launch_sort_transparent(dU, s->dev, s->d_sort_ranges.ptr, (uint32_t)s->transparent_ranges.size(), s->d_sort_keys[0].ptr,
                                s->d_sort_keys[1].ptr, s->d_sort_perm[0].ptr, s->d_sort_perm[1].ptr, s->d_sort_tmp.ptr, side);
    int prio_least = 0, prio_greatest = 0;
    if (g_cfg.kprio) cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest);
    const int prio_chain = g_cfg.kprio == 1 ? prio_greatest : prio_least, prio_tile = g_cfg.kprio == 1 ? prio_least : prio_greatest;
    g_kernel_priority_set = g_cfg.kprio != 0;
    g_kernel_priority = prio_chain;
    if (ev) cudaEventRecord(ev[0], side);
    if (!(g_cfg.skip & 1)) launch_vertex(U, dU, s->dev, ws.work, side);
    if (ev) cudaEventRecord(ev[1], side);
    if (!(g_cfg.skip & 2)) launch_setup(U, dU, s->dev, ws.work, side);
    if (ev) cudaEventRecord(ev[2], side);
    if (!(g_cfg.skip & 4)) launch_clip(U, dU, s->dev, ws.work, side);
    if (ev) cudaEventRecord(ev[3], side);
    if (!(g_cfg.skip & 8)) launch_bin_count(U, dU, ws.work, side);
    if (ev) cudaEventRecord(ev[4], side);
    if (!(g_cfg.skip & 16)) launch_alloc(U, dU, ws.work, side);
    CU(cudaEventRecord(ws.alloc_done, side));
    if (ev) cudaEventRecord(ev[5], side);
    // the empty tiles are cleared as soon as k_alloc has listed them: the stores stream to HBM under the
    // rest of the chain and under k_tile's dense tiles (disjoint pixels)
    // (clear-in-tile mode: k_tile's CTAs write them between their raster items and there is no such launch)
    if (!U.clear_in_tile) {
        CU(cudaStreamWaitEvent(ws.aux_stream, ws.alloc_done, 0));
        CU(cudaStreamWaitEvent(ws.aux_stream, ws.canvas_ready, ext));
        if (ev) cudaEventRecord(ev[N_FRAME_KERNELS - 1], ws.aux_stream);
        g_kernel_priority = prio_tile;
        if (!(g_cfg.skip & 128)) launch_clear_empty(U, dU, ws.work, ws.aux_stream);
        if (ev) cudaEventRecord(ev[N_FRAME_KERNELS], ws.aux_stream);
        CU(cudaEventRecord(ws.clear_done, ws.aux_stream));
    } else if (ev) {
        cudaEventRecord(ev[N_FRAME_KERNELS - 1], side);
        cudaEventRecord(ev[N_FRAME_KERNELS], side);
    }
    g_kernel_priority = prio_chain;
    if (!(g_cfg.skip & 32)) launch_bin_fill(U, dU, ws.work, side);
    if (ev) cudaEventRecord(ev[6], side);
    if (!(g_cfg.skip & 64)) launch_raster(U, dU, ws.work, side);
    if (ev) cudaEventRecord(ev[7], side);
    // a real event in both paths (an event-record node under capture): the next frame's painter sort waits for it
    CU(cudaEventRecordWithFlags(ws.geo_done, side, capturing ? cudaEventRecordExternal : 0u));
    CU(cudaStreamWaitEvent(side, ws.canvas_ready, ext));
    if (ev) cudaEventRecord(ev[N_FRAME_KERNELS + 1], side);
    g_kernel_priority = prio_tile;
    if (!(g_cfg.skip & 256)) launch_tile(U, dU, s->dev, ws.work, side);
    if (ev) cudaEventRecord(ev[N_FRAME_KERNELS + 2], side);
    if (!(g_cfg.skip & 512)) launch_shade(U, dU, s->dev, ws.work, side);
    if (ev) cudaEventRecord(ev[N_FRAME_KERNELS + 3], side);
    g_kernel_priority_set = 0;
    if (!U.clear_in_tile) CU(cudaStreamWaitEvent(side, ws.clear_done, 0));
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
    return DRAW_OK;
}

// Grids of the frame's kernels, by scene size (a kernel costs the pipeline its CTAs' residency).
void set_launch_grids(const draw_scene *s) {
    const bool small_scene = s->dev.n_triangles <= 200000u;
    g_clip_ctas = g_cfg.clip_ctas ? (unsigned)g_cfg.clip_ctas : (small_scene ? 74u : 296u);
    g_bin_ctas = g_cfg.bin_ctas ? (unsigned)g_cfg.bin_ctas : (small_scene ? 296u : 592u);
    g_raster_ctas = g_cfg.raster_ctas ? (unsigned)g_cfg.raster_ctas : (small_scene ? 592u : 1184u);
    g_clear_ctas = (unsigned)g_cfg.clear_ctas;
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
    U.n_lists = LISTS_PER_TILE * U.n_coarse;
    U.has_transparent = s->dev.n_transparent != 0;
    U.split_min_cost = (uint32_t)g_cfg.split_min_cost;
    U.split_div = (uint32_t)g_cfg.split_div;
    U.split_max = (uint32_t)g_cfg.split_max;
    U.defer_max = (uint32_t)g_cfg.defer_max;
    U.clear_in_tile = (uint32_t)g_cfg.clear_in_tile;
    U.bin_records_per_warp = (uint32_t)g_cfg.bin_rpw;
    U.cost_shade = (uint32_t)g_cfg.cost_shade;
    const size_t y0 = c->stripe_y1 ? c->stripe_y0 : 0, y1 = c->stripe_y1 ? c->stripe_y1 : c->height;
    U.tile_y_begin = (uint32_t)(y0 / TILE_H);
    U.tile_y_end = (uint32_t)((y1 + TILE_H - 1) / TILE_H);
    U.status_host = c->h_status; // pinned, mapped: the pointer is valid on the device (unified addressing)
    U.color = c->color();
    U.depth = c->depth();
}

// Makes one work set ready for frames of this scene / canvas geometry: its streams and pinned uniforms, its
// buffers (key pages filled), and — unless a measurement or debug tap needs the direct path — the frame's
// launches captured once as a CUDA graph (they do not change from frame to frame: the uniforms are read
// from device memory).
int ensure_set_ready(draw_scene *s, draw_scene::WorkSet &ws, const FrameUniforms &U, bool want_graph) {
    if (!ws.stream) {
        int prio_low = 0, prio_high = 0;
        CU(cudaDeviceGetStreamPriorityRange(&prio_low, &prio_high));
        CU(cudaStreamCreateWithPriority(&ws.stream, cudaStreamNonBlocking, g_cfg.prio ? prio_high : prio_low));
        CU(cudaStreamCreateWithFlags(&ws.aux_stream, cudaStreamNonBlocking));
        CU(cudaMallocHost(&ws.h_uniforms, sizeof(FrameUniforms)));
        TRY(ws.d_uniforms.reserve(1));
    }
    TRY(ensure_work_buffers(s, ws, U.n_lists));
    if (!want_graph) return DRAW_OK;
    GraphKey key{};
    key.scene = s->dev;
    key.work = ws.work;
    key.n_coarse = U.n_coarse; key.n_lists = U.n_lists; key.tiles_x = U.tiles_x;
    key.tile_y_begin = U.tile_y_begin; key.tile_y_end = U.tile_y_end;
    key.clear_ctas = g_clear_ctas + 65536u * g_tile_ctas + 7u * g_clip_ctas + 1000003u * g_bin_ctas + 15485863u * g_raster_ctas;
    if (ws.graph_exec && std::memcmp(&key, &ws.graph_key, sizeof key) == 0) return DRAW_OK;
    if (ws.graph_exec) CU(cudaGraphExecDestroy(ws.graph_exec));
    ws.graph_exec = nullptr;
    const int pdl_saved = g_pdl_enabled;
    g_pdl_enabled = 0; // plain kernel nodes: the graph already removes the launch gaps
    CU(cudaStreamBeginCapture(ws.stream, cudaStreamCaptureModeThreadLocal));
    const int rc = launch_frame(s, ws, U, ws.stream, nullptr, true);
    cudaGraph_t graph = nullptr;
    const cudaError_t e = cudaStreamEndCapture(ws.stream, &graph);
    g_pdl_enabled = pdl_saved;
    if (rc != DRAW_OK) {
        if (graph) cudaGraphDestroy(graph);
        return rc;
    }
    CU(e);
    const cudaError_t ei = cudaGraphInstantiate(&ws.graph_exec, graph, 0);
    cudaGraphDestroy(graph);
    CU(ei);
    CU(cudaGraphUpload(ws.graph_exec, ws.stream)); // the first replay does not pay the upload
    ws.graph_key = key;
    return DRAW_OK;
}

int enqueue_frame(draw_scene *s, draw_canvas *c) {
    TRY(ensure_device(s->device));
    if (s->geometry_dirty) TRY(upload_geometry(s));
    FrameUniforms U{};
    fill_uniforms(s, c, U);
    if (U.tiles_x >= MAX_TILES_X || U.tiles_y >= MAX_TILES_Y)
        return fail(DRAW_ERR_INVALID_ARGUMENT, "canvas %zux%zu is too large for the tile work list", c->width, c->height);
    draw_scene::WorkSet &ws = s->sets[s->next_set];
    const draw_scene::WorkSet &prev_ws = s->sets[s->last_set];
    s->last_set = s->next_set;
    s->next_set = (s->next_set + 1) % s->n_sets;

    set_launch_grids(s);
    // A frame's launches are replayed as a CUDA graph: one launch call instead of nine kernels, six events and
    // a copy.  Measurement taps and the debug taps use the direct path.
    const bool timing = s->kernel_timing;
    const bool use_graph = g_cfg.graphs && !timing && !s->debug_tile_cycles;
    if (s->needs_prepare) {
        // First frame of this scene (or of its new geometry / capacities): every work set is set up now — streams,
        // ~30 buffers, the fill of its key pages, graph capture and instantiation — so that no later frame pays a
        // few milliseconds of set-up in the middle of a steady stream of frames (draw_scene_prepare does the same).
        for (int i = 0; i < s->n_sets; i++) {
            if (s->sets[i].frame_pending) CU(cudaEventSynchronize(s->sets[i].frame_done));
            TRY(ensure_set_ready(s, s->sets[i], U, use_graph));
        }
        s->needs_prepare = false;
    }
    if (ws.frame_pending) CU(cudaEventSynchronize(ws.frame_done)); // the pinned uniforms of the set are about to be rewritten
    TRY(ensure_set_ready(s, ws, U, use_graph));

    // early trigger only for a lone frame: with other frames in flight the idle dependents would hold SM slots
    bool others_in_flight = false;
    for (int i = 0; i < s->n_sets; i++)
        if (&s->sets[i] != &ws && s->sets[i].frame_pending && cudaEventQuery(s->sets[i].frame_done) == cudaErrorNotReady) others_in_flight = true;
    cudaGetLastError(); // cudaErrorNotReady is not sticky, but keep the error state clean
    const int pdl_mode = g_cfg.pdl; // 0 off, 1 always early, 2 never early, 3 early for a lone frame
    g_pdl_enabled = pdl_mode != 0;
    U.pdl_early = pdl_mode == 1 || (pdl_mode == 3 && !others_in_flight);

    // ---- enqueue ----------------------------------------------------------------------------------
    // Everything of the frame runs on the work set's own streams: the geometry chain and k_tile on
    // ws.stream, k_clear_empty beside them on ws.aux_stream (forked after k_alloc, joined after k_tile).
    // The canvas' stream only brackets the frame: the set's stream first waits for whatever the canvas
    // stream still does with the canvas, and the canvas stream then waits for the frame.  Frames that
    // use different work sets are therefore independent and overlap; a set's next frame follows its
    // previous one in stream order.
    cudaStream_t side = ws.stream, st = c->stream;
    *ws.h_uniforms = U;
    // what the canvas stream still does with the canvas comes before the two kernels that write it
    // (k_clear_empty, k_tile wait for this event; the geometry chain does not touch the canvas and does not wait)
    CU(cudaEventRecord(ws.canvas_ready, st));
    if (s->dev.n_transparent) {
        // the painter sort rewrites the shared index streams: order it after the previous frame's geometry
        // (geo_done is a real event in both paths: launch_frame records it with cudaEventRecordExternal under capture)
        if (&prev_ws != &ws && prev_ws.frame_pending) CU(cudaStreamWaitEvent(side, prev_ws.geo_done, 0));
    }
    if (ws.work.tile_cycles) CU(cudaMemsetAsync(ws.work.tile_cycles, 0, U.n_lists * sizeof(uint32_t), side)); // debug taps are atomicMax'd

    cudaEvent_t *ev = nullptr;
    if (timing) {
        for (int i = 0; i < N_FRAME_KERNELS + 4; i++)
            if (!s->kev[i]) CU(cudaEventCreate(&s->kev[i]));
        ev = s->kev;
        s->kev_recorded = true;
    }
    if (use_graph) CU(cudaGraphLaunch(ws.graph_exec, side));
    else TRY(launch_frame(s, ws, U, side, ev, false));
    CU(cudaEventRecord(ws.frame_done, side));
    ws.frame_pending = true;
    CU(cudaStreamWaitEvent(st, ws.frame_done, 0));
    s->launches += 5 + (s->dev.n_triangles ? 2 : 0) + (s->dev.n_transparent ? 1 : 0) + (U.tile_y_end > U.tile_y_begin ? (U.defer_max ? 3 : 2) - (U.clear_in_tile ? 1 : 0) : 0);
    CU(cudaGetLastError());
    c->frame_pending = true;
    c->host_dirty = true;
    c->last_scene = s;
    c->pending_inputs.camera = s->camera;
    c->pending_inputs.light = s->light;
    c->pending_inputs.off_x = c->off_x; c->pending_inputs.off_y = c->off_y;
    c->pending_inputs.stripe_y0 = c->stripe_y0; c->pending_inputs.stripe_y1 = c->stripe_y1;
    c->pending_inputs.depth_max = c->depth_max;
    if (c->host_mirror && !c->ext_color) {
        // the frame follows its render to the host without waiting for the host to ask (map_host then only waits)
        const size_t bytes = c->width * c->height * 4;
        TRY(ensure_host_mirror(c, bytes));
        CU(cudaMemcpyAsync(c->h_color, c->color(), bytes, cudaMemcpyDeviceToHost, st));
        c->host_dirty = false;
    }
    return DRAW_OK;
}

// Waits for the canvas' stream; if the last frame overflowed a work buffer, grows it and
// renders the frame again (so what the host reads is always a complete frame).
int finish_frame(draw_canvas *c) {
    TRY(ensure_device(c->device));
    CU(cudaStreamSynchronize(c->stream));
    int guard = 0;
    while (c->frame_pending) {
        c->frame_pending = false;
        draw_scene *s = c->last_scene;
        bool alive;
        {
            std::lock_guard<std::mutex> lk(g_registry_mutex);
            alive = g_live_scenes.count(s) != 0;
        }
        const uint32_t n_rec = c->h_status[0], n_refs = c->h_status[1], overflow = c->h_status[2];
        c->stats.setup_records = n_rec;
        c->stats.tile_refs = n_refs;
        c->stats.empty_tiles = c->h_status[13];
        c->stats.key_pages = c->h_status[11];
        c->stats.clear_in_tile = (uint32_t)g_cfg.clear_in_tile;
        if (alive) {
            c->stats.input_triangles = s->dev.n_triangles;
            c->stats.transparent_slots = s->dev.n_transparent * 4;
        }
        if (!overflow) break;
        c->stats.overflow = overflow;
        if (!alive) return fail(DRAW_ERR_INTERNAL, "frame overflowed a work buffer and its scene is gone");
        if (++guard > 4) return fail(DRAW_ERR_INTERNAL, "work buffers keep overflowing");
        if (overflow & OVERFLOW_RECORDS)
            s->rec_cap = std::max<size_t>((size_t)n_rec + n_rec / 4 + 1024, 4 * (size_t)s->dev.n_triangles + 1024);
        if (overflow & OVERFLOW_REFS) s->refs_cap = (size_t)n_refs + n_refs / 4 + 4096;
        else if (overflow & OVERFLOW_RECORDS) s->refs_cap = std::max(s->refs_cap, 4 * s->rec_cap);
        CU(cudaDeviceSynchronize()); // the work sets are about to be reallocated
        s->needs_prepare = true;
        // Render the frame again from the inputs it was rendered with (the scene's camera / light and the
        // canvas' offset / stripe may have moved on since), then put the current state back.
        const draw_canvas::FrameInputs in = c->pending_inputs;
        const CameraState cam_now = s->camera;
        const f3 light_now = s->light;
        const int off_x_now = c->off_x, off_y_now = c->off_y;
        const size_t sy0_now = c->stripe_y0, sy1_now = c->stripe_y1;
        const float dmax_now = c->depth_max;
        s->camera = in.camera; s->light = in.light;
        c->off_x = in.off_x; c->off_y = in.off_y;
        c->stripe_y0 = in.stripe_y0; c->stripe_y1 = in.stripe_y1;
        c->depth_max = in.depth_max;
        const int rc = enqueue_frame(s, c);
        s->camera = cam_now; s->light = light_now;
        c->off_x = off_x_now; c->off_y = off_y_now;
        c->stripe_y0 = sy0_now; c->stripe_y1 = sy1_now;
        c->depth_max = dmax_now;
        TRY(rc);
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
        for (draw_scene::WorkSet &ws : scene->sets) {
            if (ws.alloc_done) cudaEventDestroy(ws.alloc_done);
            if (ws.geo_done) cudaEventDestroy(ws.geo_done);
        }
        for (int i = 0; i < N_FRAME_KERNELS + 4; i++)
            if (scene->kev[i]) cudaEventDestroy(scene->kev[i]);
        for (draw_scene::WorkSet &ws : scene->sets) {
            if (ws.graph_exec) cudaGraphExecDestroy(ws.graph_exec);
            if (ws.stream) cudaStreamDestroy(ws.stream);
            if (ws.aux_stream) cudaStreamDestroy(ws.aux_stream);
            if (ws.h_uniforms) cudaFreeHost(ws.h_uniforms);
            if (ws.canvas_ready) cudaEventDestroy(ws.canvas_ready);
            if (ws.clear_done) cudaEventDestroy(ws.clear_done);
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
    if (canvas->frame_pending) {
        // settle the previous frame's status first (grows buffers if it overflowed)
        TRY(ensure_device(canvas->device));
        if (cudaStreamQuery(canvas->stream) == cudaSuccess) TRY(finish_frame(canvas));
    }
    canvas->stats.overflow = 0;
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
    const bool use_graph = g_cfg.graphs && !scene->kernel_timing && !scene->debug_tile_cycles;
    for (int i = 0; i < scene->n_sets; i++) {
        if (scene->sets[i].frame_pending) CU(cudaEventSynchronize(scene->sets[i].frame_done));
        TRY(ensure_set_ready(scene, scene->sets[i], U, use_graph));
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
    std::vector<float> tmp(count);
    const int order[7] = {0, 1, 2, 3, 4, 5, 6}; // light xyz, halfway xyz, depth
    for (int k = 0; k < 7; k++) {
        CU(cudaMemcpy(tmp.data(), scene->sets[scene->last_set].vert[order[k]].ptr + first, count * sizeof(float), cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < count; i++) out[7 * i + k] = tmp[i];
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

int draw_scene_last_kernel_times(draw_scene *scene, draw_canvas *canvas, float ms[10]) {
    GUARD_BEGIN
    if (!scene || !canvas || !ms) return fail(DRAW_ERR_INVALID_ARGUMENT, "NULL argument");
    if (!scene->kev_recorded) return fail(DRAW_ERR_INVALID_ARGUMENT, "kernel timing was not enabled for the last frame");
    TRY(finish_frame(canvas));
    for (int i = 0; i < 7; i++) CU(cudaEventElapsedTime(&ms[i], scene->kev[i], scene->kev[i + 1])); // k_vertex .. k_raster
    CU(cudaEventElapsedTime(&ms[7], scene->kev[N_FRAME_KERNELS - 1], scene->kev[N_FRAME_KERNELS]));     // k_clear_empty (aux stream)
    CU(cudaEventElapsedTime(&ms[8], scene->kev[N_FRAME_KERNELS + 1], scene->kev[N_FRAME_KERNELS + 2])); // k_tile
    CU(cudaEventElapsedTime(&ms[9], scene->kev[N_FRAME_KERNELS + 2], scene->kev[N_FRAME_KERNELS + 3])); // k_shade
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
    // layout: [tile] whole CTA, [n_tiles + tile] end of phase A, [2 n_tiles + tile] end of phase B
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
    const size_t coarse = tiles_x * tiles_y, lists = coarse * LISTS_PER_TILE;
    if (n_coarse) *n_coarse = coarse;
    if (!out) return DRAW_OK;
    if (n != lists) return fail(DRAW_ERR_INVALID_ARGUMENT, "n must be %zu", lists);
    CU(cudaMemcpy(out, scene->sets[scene->last_set].list_count.ptr, lists * sizeof(uint32_t), cudaMemcpyDeviceToHost));
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
    if (cudaMallocHost(&c->h_status, 16 * sizeof(uint32_t)) != cudaSuccess)
        return cleanup(fail(DRAW_ERR_OUT_OF_MEMORY, "cudaMallocHost failed"));
    std::memset(c->h_status, 0, 16 * sizeof(uint32_t));
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
    canvas->host_dirty = true;
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
    canvas->host_dirty = true;
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
        CU(cudaMemcpyAsync(canvas->h_color, canvas->color(), bytes, cudaMemcpyDeviceToHost, canvas->stream));
        CU(cudaStreamSynchronize(canvas->stream));
        canvas->host_dirty = false;
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
    canvas->host_dirty = true;
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

} // extern "C"

namespace drawb200 {
int loader_fail(int code, const char *msg) { return fail(code, "%s", msg); }
} // namespace drawb200
