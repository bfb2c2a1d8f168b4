// k_sort.cu — painter sort of the transparent meshes on the device (mororo18/draw scene/mod.rs:1100-1115).
//
// The reference sorts every transparent mesh's triangle list IN PLACE, once per frame, by the distance
// of the triangle's centroid to the camera, farthest first:
//     triangles.sort_by(|a, b| dist(center(a), cam).total_cmp(&dist(center(b), cam)).reverse())
// `sort_by` is stable and the list keeps its order from frame to frame, so equal keys stay in whatever
// order earlier frames left them.  Here the mesh's triangles are the range [first, first + n) of the nine
// draw-order index streams (SceneDev::idx); the kernel sorts that range in place with a stable LSD radix
// sort on the key, so the streams always hold exactly the order the reference's list would have.
//
//   one CTA per transparent mesh (ranges are independent, meshes are small next to the opaque scene)
//   key   = ~order(dist): `order` maps f32 to u32 so that unsigned compare == f32::total_cmp; the
//           complement turns "descending, stable" into "ascending, stable"
//   sort  = 4 passes of 8 bits over (key, position) pairs; a pass whose digit is the same for all keys
//           is skipped (distances share their exponent byte); a range already in order is left alone
//           (a still camera: the usual case)
//   rank  = within a chunk of 1024 elements in list order: __match_any_sync gives the rank among the
//           warp's equal digits, a 32 x 256 table of warp counts scanned per digit the rank among warps
//   apply = each of the nine streams is gathered through the permutation into scratch and copied back
#include "device_math.cuh"

namespace drawb200 {

constexpr int SORT_THREADS = 1024, SORT_WARPS = SORT_THREADS / 32;

struct SortRange {
    uint32_t first, n, base, pad; // first triangle (draw order), triangles, offset into the scratch arrays
};

__device__ __forceinline__ uint32_t painter_key(const SceneDev &S, const v3 cam, uint32_t tri) {
    const uint32_t i0 = S.idx[0][tri], i1 = S.idx[1][tri], i2 = S.idx[2][tri];
    const float4 pa = __ldg(S.pos4 + i0), pb = __ldg(S.pos4 + i1), pc = __ldg(S.pos4 + i2);
    const v3 a{pa.x, pa.y, pa.z}, b{pb.x, pb.y, pb.z}, c{pc.x, pc.y, pc.z};
    const v3 center = v_div(v_add(v_add(a, b), c), 3.0f); // scene/mod.rs:1108
    const float d = v_norm(v_sub(center, cam));            // Vec3::dist, linalg.rs:177-180
    const uint32_t bits = __float_as_uint(d);
    const uint32_t order = bits ^ ((bits >> 31) ? 0xFFFFFFFFu : 0x80000000u); // unsigned compare == total_cmp
    return ~order;                                                              // .reverse()
}

__global__ void __launch_bounds__(SORT_THREADS) k_sort_transparent(const FrameUniforms *__restrict__ Up, const SceneDev S,
                                                                   const SortRange *__restrict__ ranges, uint32_t *keys0,
                                                                   uint32_t *keys1, uint32_t *perm0, uint32_t *perm1,
                                                                   uint32_t *scratch) {
    __shared__ uint32_t hist[256], digit_base[256], warp_total[8];
    __shared__ uint32_t warp_count[SORT_WARPS][256];
    const SortRange r = ranges[blockIdx.x];
    const uint32_t n = r.n, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (n < 2) return;
    const v3 cam{Up->cam[0], Up->cam[1], Up->cam[2]};
    uint32_t *kin = keys0 + r.base, *kout = keys1 + r.base, *pin = perm0 + r.base, *pout = perm1 + r.base;

    // ---- keys, identity permutation ---------------------------------------------------------------
    for (uint32_t i = tid; i < n; i += SORT_THREADS) {
        kin[i] = painter_key(S, cam, r.first + i);
        pin[i] = i;
    }
    __syncthreads();
    bool disorder = false;
    for (uint32_t i = tid + 1; i < n; i += SORT_THREADS) disorder = disorder || kin[i - 1] > kin[i];
    if (!__syncthreads_or(disorder)) return; // already in painter order (stable: nothing moves)

    // ---- stable LSD radix sort --------------------------------------------------------------------
#pragma unroll 1
    for (int pass = 0; pass < 4; pass++) {
        const int shift = 8 * pass;
        if (tid < 256) hist[tid] = 0;
        __syncthreads();
        for (uint32_t i = tid; i < n; i += SORT_THREADS) atomicAdd(&hist[(kin[i] >> shift) & 255u], 1u);
        __syncthreads();
        if (__syncthreads_or(tid < 256 && hist[tid] == n)) continue; // one digit for all: the pass is the identity
        if (tid < 256) { // exclusive scan of the histogram
            const uint32_t v = hist[tid];
            uint32_t inc = v;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t u = __shfl_up_sync(0xFFFFFFFFu, inc, d);
                if (lane >= (uint32_t)d) inc += u;
            }
            if (lane == 31) warp_total[warp] = inc;
            digit_base[tid] = inc - v;
        }
        __syncthreads();
        if (tid < 256) {
            uint32_t before = 0;
            for (uint32_t w = 0; w < warp; w++) before += warp_total[w];
            digit_base[tid] += before;
        }
        __syncthreads();
#pragma unroll 1
        for (uint32_t c0 = 0; c0 < n; c0 += SORT_THREADS) { // chunks in list order
            for (uint32_t j = tid; j < SORT_WARPS * 256; j += SORT_THREADS) (&warp_count[0][0])[j] = 0;
            __syncthreads();
            const uint32_t i = c0 + tid;
            const bool valid = i < n;
            const uint32_t key = valid ? kin[i] : 0u, from = valid ? pin[i] : 0u;
            const uint32_t digit = valid ? (key >> shift) & 255u : 256u + lane; // idle lanes match nobody
            const uint32_t peers = __match_any_sync(0xFFFFFFFFu, digit);
            const uint32_t rank = (uint32_t)__popc(peers & ((1u << lane) - 1u));
            if (valid && rank == 0) warp_count[warp][digit] = (uint32_t)__popc(peers);
            __syncthreads();
            if (tid < 256) { // per digit: counts of the warps -> their first output position
                uint32_t run = digit_base[tid];
#pragma unroll 8
                for (int w = 0; w < SORT_WARPS; w++) {
                    const uint32_t t = warp_count[w][tid];
                    warp_count[w][tid] = run;
                    run += t;
                }
                digit_base[tid] = run; // the next chunk continues here
            }
            __syncthreads();
            if (valid) {
                const uint32_t pos = warp_count[warp][digit] + rank;
                kout[pos] = key;
                pout[pos] = from;
            }
            __syncthreads();
        }
        uint32_t *t = kin; kin = kout; kout = t;
        t = pin; pin = pout; pout = t;
    }

    // ---- apply: the nine index streams follow the permutation --------------------------------------
    // (the streams are const for every other kernel; this one owns them between frames, scene.cpp)
    uint32_t *tmp = scratch + r.base;
#pragma unroll 1
    for (int c = 0; c < 9; c++) {
        uint32_t *stream = const_cast<uint32_t *>(S.idx[c]) + r.first;
        for (uint32_t i = tid; i < n; i += SORT_THREADS) tmp[i] = stream[pin[i]];
        __syncthreads();
        for (uint32_t i = tid; i < n; i += SORT_THREADS) stream[i] = tmp[i];
        __syncthreads();
    }
}

void launch_sort_transparent(const FrameUniforms *dU, const SceneDev &S, const void *ranges, uint32_t n_ranges,
                             uint32_t *keys0, uint32_t *keys1, uint32_t *perm0, uint32_t *perm1, uint32_t *scratch,
                             cudaStream_t stream) {
    if (!n_ranges) return;
    k_sort_transparent<<<n_ranges, SORT_THREADS, 0, stream>>>(dU, S, static_cast<const SortRange *>(ranges), keys0, keys1,
                                                              perm0, perm1, scratch);
}

} // namespace drawb200
