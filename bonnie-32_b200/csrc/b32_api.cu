// b32_api.cu — the C ABI of include/b32_raster.h: context, device memory, streams, and the
// orchestration of one render_mesh_15 call (render.rs:2302-2638) as a sequence of kernels on the
// context's stream.  There is no CPU fallback: without a CUDA device b32_ctx_create fails.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "b32_launch.h"
#include "b32_raster.h"

using namespace b32;

namespace {

template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t cap = 0;     // elements
    // Growing discards the contents.  Enqueued work of any stream may still read the old buffer: wait for the whole
    // device before it is freed (growth is rare; steady state never gets here).
    cudaError_t reserve(size_t n) {
        if (n <= cap) return cudaSuccess;
        size_t ncap = std::max(n, cap + cap / 2);
        T* q = nullptr;
        cudaError_t e = cudaMalloc(&q, ncap * sizeof(T));
        if (e != cudaSuccess) return e;
        if (p) { cudaDeviceSynchronize(); cudaFree(p); }
        p = q; cap = ncap;
        return cudaSuccess;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

}  // namespace

struct b32_mesh {
    b32_vertex* verts = nullptr;
    b32_face* faces = nullptr;
    uint32_t nv = 0, nf = 0;
    bool has_nonopaque = false;     // some face has blend != Opaque or editor_alpha < 255
};

// What fixes the topology (kernels, grids) of an enqueued frame; everything else is a kernel argument.
struct FrameKey {
    const void* verts = nullptr; const void* faces = nullptr;
    uint32_t nv = 0, nf = 0, width = 0, height = 0;
    uint8_t rgb888 = 0, pass1 = 0, clear = 0, valid = 0, ordered = 0, spot = 0, prefix = 0;
    bool operator==(const FrameKey& o) const {
        return verts == o.verts && faces == o.faces && nv == o.nv && nf == o.nf && width == o.width && height == o.height &&
               rgb888 == o.rgb888 && pass1 == o.pass1 && clear == o.clear && valid == o.valid && ordered == o.ordered && spot == o.spot && prefix == o.prefix;
    }
};
struct FrameGraph {
    cudaGraph_t graph = nullptr;
    GraphPatch patch;
    FrameKey key, seen;
    void destroy() {
        if (patch.exec) cudaGraphExecDestroy(patch.exec);
        if (graph) cudaGraphDestroy(graph);
        patch = GraphPatch{}; graph = nullptr; key = FrameKey{};
    }
};

struct b32_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaDeviceProp prop{};
    std::string err;
    int deferred = B32_OK;
    uint64_t launches = 0;

    // framebuffer (render.rs:10-15)
    uint32_t width = 0, height = 0;
    DevBuf<uint32_t> fb_rgba;
    DevBuf<float> fb_z;

    // textures
    DevBuf<uint16_t> texels;
    DevBuf<uint32_t> texmask;          // 1 bit per texel of the pool: the texel writes on a black-keyed surface
    uint32_t texmask_words = 0;        // multiple of 4 (16-byte bulk copies)
    DevBuf<TexDev> texdesc;
    std::vector<TexDev> texdesc_h;
    uint32_t ntex = 0;

    // textures of the RGB888 sibling (render_mesh): one Color (r, g, b, blend) = u32 per texel
    DevBuf<uint32_t> texels8;
    DevBuf<uint32_t> tex8mask;         // 1 bit per texel: tag != Erase
    uint32_t tex8mask_words = 0;
    DevBuf<TexDev> tex8desc;           // TexDev.blend = 1 iff the texture holds texels with a PS1 blend tag
    uint32_t ntex8 = 0;

    // staging geometry for host-buffer calls
    DevBuf<b32_vertex> verts;
    DevBuf<b32_face> faces;

    // per-call work buffers
    DevBuf<TVert> tv;
    DevBuf<SurfRec> recs;
    DevBuf<uint64_t> keys;
    // CallState + pass-1 tile counters live together in one of two sets used alternately: k_setup of call i
    // zeroes the set of call i+1, so no memset sits in front of a call on the stream
    DevBuf<uint32_t> state_ring;
    uint32_t state_stride = 0;         // words per set
    int state_cur = 0;
    uint32_t* tile_count = nullptr;    // current set's counters: ordered-pass entries per mask tile
    DevBuf<uint4> masks;               // tile masks [mask tile][face group] (binning without bins, b32_device.cuh)
    DevBuf<BinHead> heads, obins;      // per-face heads; obins: scratch slices of the ordered pass's crowded tiles
    DevBuf<BinHead> crowd;             // pass 1: scratch for tiles with more candidates than one window (allocated for large meshes only)
    uint32_t obin_cap = 0;             // entries per tile slice of obins (a power of two; 0 = none allocated)
    uint32_t obin_tiles = 0;           // ... allocated for this many tiles
    DevBuf<WireTri> wire;
    DevBuf<uint32_t> wire_table;       // open-addressing table of the wireframe phase's edge de-duplication
    DevBuf<b32_line> lines;            // overlay lines (b32_draw_lines): the list, then per pixel owner + next, then round flags
    DevBuf<uint32_t> line_scratch;
    std::vector<LightDev> lights_h;
    bool async_pending = false;
    // One of the last PREFIX_HINT_CALLS calls on this context drew large surfaces with stepped edge values in fixed-point
    // mode (k_setup stored its call number in the mapped word behind hstat): the next calls run the shared-prefix fill
    // (k_fill_opaque<.., PRE>).  Blocking and enqueue-only callers alike: nobody waits for the word.
    static constexpr uint32_t PREFIX_HINT_CALLS = 8;
    bool prefix_hint = false;
    uint32_t call_seq = 0;
    uint32_t* stepped_seq = nullptr;   // host view of the mapped word
    uint32_t* stepped_seq_dev = nullptr;
    bool next_prefix_hint() const {
        const uint32_t seen = __atomic_load_n(stepped_seq, __ATOMIC_RELAXED);
        return seen != 0 && (call_seq + 1 - seen) <= PREFIX_HINT_CALLS;
    }
    CallParams last_params{};
    DevBuf<LightDev> lights;
    DevBuf<float> dbg;
    CallState* state = nullptr;        // device: current set's CallState
    uint32_t* sticky = nullptr;        // device: error bits of enqueue-only calls
    CallState* state_h = nullptr;      // pinned host
    HostStatus* hstat = nullptr;       // pinned + mapped: what a blocking call's kernels publish (b32_device.cuh)
    HostStatus* hstat_dev = nullptr;   // ... its device address
    // opt-in (b32_ctx_frame_timings): enqueued frames report the same way, into a ring of status blocks
    static constexpr uint32_t N_FRAME_STATUS = 8;
    HostStatus* fstat = nullptr; HostStatus* fstat_dev = nullptr;
    struct FrameSlot { uint32_t seq = 0; uint8_t last = 0, pass1 = 0, ordered = 0; } fslot[N_FRAME_STATUS];
    uint32_t fslot_next = 0;
    bool frame_timings = false;
    uint32_t host_seq = 0;
    uint8_t* pinned = nullptr;         // pinned staging ring for pageable host buffers
    size_t pinned_bytes = 0;
    uint32_t last_nf = 0;

    cudaEvent_t ev[8] = {};
    cudaEvent_t h2d_ev[2] = {};        // slot-free events of the pinned staging ring
    bool h2d_pending[2] = {false, false};
    int h2d_next = 0;
    float kernel_ms[7] = {};
    float emit_ms = 0.0f;
    float wire_ms = 0.0f;

    // enqueued frames of one mesh replay as a CUDA graph (see render_device); a few topologies are kept (round robin)
    static constexpr int N_FRAME_GRAPHS = 32;
    FrameGraph fgs[N_FRAME_GRAPHS];
    FrameKey fg_seen[N_FRAME_GRAPHS];  // topologies seen once, not captured yet
    int fg_next = 0, fg_seen_next = 0;
    uint64_t graph_launches = 0, graph_captures = 0;

    // measurement aid (b32_debug_timing_ring): enqueued frames are launched plainly with events around their kernel groups
    std::vector<cudaEvent_t> tring;    // 3 events per slot: before k_setup, before the fill(s), after
    uint32_t tring_n = 0, tring_count = 0;

    LaunchCtx L() { return LaunchCtx{stream, (uint32_t)prop.multiProcessorCount, &launches, nullptr}; }
};

namespace {

int fail(b32_ctx* c, int code, const std::string& msg) { if (c) c->err = msg; return code; }
int cuda_fail(b32_ctx* c, cudaError_t e, const char* what) {
    return fail(c, B32_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}
#define CK(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) return cuda_fail(ctx, _e, #call); } while (0)
// A host thread may hold contexts on several GPUs: every entry point first selects its context's device.
#define USE_DEVICE(ctx) do { if ((ctx) && cudaSetDevice((ctx)->device) != cudaSuccess) return cuda_fail((ctx), cudaGetLastError(), "cudaSetDevice"); } while (0)

constexpr uint32_t STATE_WORDS = 16;    // CallState padded to 64 bytes; the tile counters follow
static_assert(sizeof(CallState) <= STATE_WORDS * 4, "CallState must fit its slot");

int32_t host_fx_from_f32(float f) {                 // Fixed32::from_f32, fixed.rs:125-127 (saturating `as i32`)
    float s = f * 4096.0f;
    if (s != s) return 0;
    if (s >= 2147483648.0f) return INT32_MAX;
    if (s <= -2147483648.0f) return INT32_MIN;
    return (int32_t)s;
}

int fill_params(b32_ctx* ctx, CallParams& p, const b32_camera* cam, const b32_settings* s, const b32_fog* fog,
                uint32_t nv, uint32_t nf, std::vector<LightDev>& lights, bool rgb888 = false) {
    std::memset(&p, 0, sizeof(p));
    for (int i = 0; i < 3; ++i) {
        p.cam_pos[i] = cam->position[i]; p.bx[i] = cam->basis_x[i]; p.by[i] = cam->basis_y[i]; p.bz[i] = cam->basis_z[i];
        p.fcam_pos[i] = host_fx_from_f32(cam->position[i]);
        p.fbx[i] = host_fx_from_f32(cam->basis_x[i]); p.fby[i] = host_fx_from_f32(cam->basis_y[i]); p.fbz[i] = host_fx_from_f32(cam->basis_z[i]);
    }
    p.width = ctx->width; p.height = ctx->height;
    p.tiles_x = (ctx->width + TILE_W - 1) / TILE_W; p.tiles_y = (ctx->height + TILE_H - 1) / TILE_H;
    p.viewport_scale = host_fx_from_f32(((float)std::min(ctx->width, ctx->height) / 2.0f) * 0.75f);   // fixed.rs:398
    p.half_w = (int32_t)(((uint32_t)((int32_t)ctx->width / 2)) << 12);                                // fixed.rs:399-400
    p.half_h = (int32_t)(((uint32_t)((int32_t)ctx->height / 2)) << 12);
    p.nv = nv; p.nf = nf; p.ntex = rgb888 ? ctx->ntex8 : ctx->ntex;
    p.rgb888 = rgb888 ? 1 : 0;
    p.vwords = 9;
    // tile-mask geometry: coarser mask tiles for frames / meshes whose table would be too large (b32_device.cuh)
    p.n_groups = (nf + SETUP_GROUP - 1) / SETUP_GROUP;
    for (p.mshift = 0;; ++p.mshift) {
        p.mtiles_x = (p.tiles_x + (1u << p.mshift) - 1) >> p.mshift; p.mtiles_y = (p.tiles_y + (1u << p.mshift) - 1) >> p.mshift;
        const uint64_t n_mt = (uint64_t)p.mtiles_x * p.mtiles_y;
        if ((n_mt <= MASK_TILES_MAX && n_mt * std::max<uint32_t>(p.n_groups, 1) * sizeof(uint4) <= MASK_BYTES_MAX) || n_mt <= 1) break;
    }
    const uint32_t mask_words = rgb888 ? ctx->tex8mask_words : ctx->texmask_words;
    p.mask_smem_words = mask_words <= (uint32_t)OP_MASK_SMEM_WORDS ? mask_words : 0;
    p.affine_textures = s->affine_textures != 0; p.use_zbuffer = s->use_zbuffer != 0; p.shading = s->shading;
    p.backface_cull = s->backface_cull != 0; p.dithering = s->dithering != 0; p.use_fixed_point = s->use_fixed_point != 0;
    p.xray_mode = s->xray_mode != 0; p.ortho = s->ortho_enabled != 0;
    p.wire_back = (s->backface_cull && s->backface_wireframe) ? 1 : 0;      // render.rs:2576
    p.wire_front = s->wireframe_overlay ? 1 : 0;                            // render.rs:2606, :2550
    p.ambient = s->ambient; p.ortho_zoom = s->ortho_zoom; p.ortho_cx = s->ortho_center_x; p.ortho_cy = s->ortho_center_y;
    if (fog) {
        p.fog_enabled = 1; p.fog_start = fog->start; p.fog_falloff = fog->falloff; p.fog_cull = fog->cull_distance;
        p.fog_r = fog->r; p.fog_g = fog->g; p.fog_b = fog->b; p.fog_blend = fog->blend;
    }
    if (s->shading > B32_SHADE_GOURAUD) return fail(ctx, B32_ERR_INVALID, "settings.shading out of range");
    lights.clear();
    if (s->n_lights && !s->lights) return fail(ctx, B32_ERR_INVALID, "settings.lights is NULL");
    for (uint32_t i = 0; i < s->n_lights; ++i) {
        const b32_light& l = s->lights[i];
        if (l.type > B32_LIGHT_SPOT) return fail(ctx, B32_ERR_INVALID, "light type out of range");
        LightDev d{};
        d.type = l.type; d.px = l.position[0]; d.py = l.position[1]; d.pz = l.position[2];
        d.dx = l.direction[0]; d.dy = l.direction[1]; d.dz = l.direction[2];
        d.radius = l.radius; d.angle = l.angle; d.intensity = l.intensity;
        d.cr = (float)l.r / 255.0f; d.cg = (float)l.g / 255.0f; d.cb = (float)l.b / 255.0f;   // render.rs:1062-1064
        d.enabled = l.enabled;
        if (l.type == B32_LIGHT_SPOT && l.enabled && s->shading != B32_SHADE_NONE) p.has_spot = 1;
        lights.push_back(d);
    }
    p.n_lights = (uint32_t)lights.size();
    ctx->prefix_hint = ctx->next_prefix_hint();
    p.call_seq = ++ctx->call_seq; p.stepped_seq_host = ctx->stepped_seq_dev;
    p.prefer_prefix = ctx->prefix_hint ? 1 : 0;
    return B32_OK;
}

// copy a host buffer to the device on the context stream; pageable memory goes through the pinned ring
// consume = true: `src` may be reused by the caller as soon as this returns, so even pinned sources go through the staging ring
int h2d(b32_ctx* ctx, void* dst, const void* src, size_t bytes, bool consume = false) {
    if (!bytes) return B32_OK;
    cudaPointerAttributes a{};
    bool pinned_src = !consume && cudaPointerGetAttributes(&a, src) == cudaSuccess && (a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged);
    cudaGetLastError();
    if (pinned_src) { CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream)); return B32_OK; }
    // two-slot ring: memcpy into slot k overlaps the DMA of slot k^1.  `src` is consumed by the memcpy, so the call does
    // not wait for the DMA: a slot's event is only waited for when the slot is needed again (by this or a later copy).
    const size_t slot = ctx->pinned_bytes / 2;
    size_t off = 0;
    while (off < bytes) {
        const int k = ctx->h2d_next;
        size_t n = std::min(slot, bytes - off);
        if (ctx->h2d_pending[k]) { CK(cudaEventSynchronize(ctx->h2d_ev[k])); ctx->h2d_pending[k] = false; }
        std::memcpy(ctx->pinned + k * slot, (const uint8_t*)src + off, n);
        CK(cudaMemcpyAsync((uint8_t*)dst + off, ctx->pinned + k * slot, n, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaEventRecord(ctx->h2d_ev[k], ctx->stream)); ctx->h2d_pending[k] = true;
        off += n; ctx->h2d_next = k ^ 1;
    }
    return B32_OK;
}

// Experiment switch: the vertex transform as its own kernel (one vertex per thread) in front of k_setup.
static bool split_transform() { static const bool on = std::getenv("B32_SPLIT_TRANSFORM") != nullptr; return on; }

int ensure_work(b32_ctx* ctx, const CallParams& p) {
    uint32_t m = std::max<uint32_t>(p.nf, 1);
    CK(ctx->recs.reserve(m));
    CK(ctx->keys.reserve(m));
    CK(ctx->heads.reserve(m));
    if (split_transform()) CK(ctx->tv.reserve(std::max<uint32_t>(p.nv, 1)));
    CK(ctx->masks.reserve(std::max<size_t>((size_t)p.mtiles_x * p.mtiles_y * p.n_groups, 1)));
    // A tile is crowded when more than OP_SORT_MAX_ENTRIES of the mesh's faces touch it, so only meshes well beyond that can
    // have any: they get 5 head-sized slots of scratch per face (a face's bounding box touches ~4 tiles; a slot = one head, or four face indices), capped at 256 MB.  A tile that
    // finds the scratch exhausted walks its windows in list order instead (same result, later early-out).
    if (p.nf > 4u * OP_SORT_MAX_ENTRIES) CK(ctx->crowd.reserve(std::min<size_t>((size_t)p.nf * 5, (size_t)16 << 20)));
    uint32_t ntiles = p.tiles_x * p.tiles_y;
    uint32_t need = STATE_WORDS + ((std::max<uint32_t>(ntiles, 1) + 3u) & ~3u);
    if (need > ctx->state_stride) {
        CK(cudaStreamSynchronize(ctx->stream));
        ctx->state_ring.release();
        CK(ctx->state_ring.reserve((size_t)need * 2));
        CK(cudaMemsetAsync(ctx->state_ring.p, 0, (size_t)need * 2 * sizeof(uint32_t), ctx->stream));
        ctx->state_stride = need;
        ctx->state_cur = 0;
        ctx->state = reinterpret_cast<CallState*>(ctx->state_ring.p);
        ctx->tile_count = ctx->state_ring.p + STATE_WORDS;
    }
    return B32_OK;
}

// the scratch slices of the ordered pass: `cap` entries (a power of two) for each of the frame's tiles
uint32_t ordered_scratch_cap(const b32_ctx* ctx, uint32_t ntiles) {
    return (ctx->obin_cap && ctx->obin_tiles >= ntiles) ? ctx->obin_cap : 0u;
}

int upload_lights(b32_ctx* ctx, const std::vector<LightDev>& lights) {
    if (lights.empty()) return B32_OK;
    // lights change rarely: skip the copy (and its sync) when the bytes are the ones already on the device
    size_t bytes = lights.size() * sizeof(LightDev);
    if (ctx->lights_h.size() == lights.size() && std::memcmp(ctx->lights_h.data(), lights.data(), bytes) == 0) return B32_OK;
    CK(ctx->lights.reserve(lights.size()));
    CK(cudaStreamSynchronize(ctx->stream));       // earlier calls may still read the old lights
    CK(cudaMemcpyAsync(ctx->lights.p, lights.data(), bytes, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));       // `lights` is a host temporary
    ctx->lights_h = lights;
    return B32_OK;
}

// Pass 2 (semi-transparent surfaces, back to front) and x-ray mode: strict draw-order replay (k_fill_ordered).
// obin_max = the largest number of ordered entries k_setup counted in one mask tile: tiles with more than fit shared
// memory sort in a slice of a global scratch, sized here (so the pass never has to be redone).
int render_ordered(b32_ctx* ctx, const CallParams& p, uint32_t obin_max) {
    LaunchCtx L = ctx->L();
    const uint32_t ntiles = p.tiles_x * p.tiles_y;
    if (obin_max > (uint32_t)ORD_SORT_MAX && obin_max > ordered_scratch_cap(ctx, ntiles)) {
        uint32_t cap2 = 2; while (cap2 < obin_max) cap2 <<= 1;               // power of two: the tile sort pads in place
        CK(ctx->obins.reserve((size_t)ntiles * cap2));
        ctx->obin_cap = cap2; ctx->obin_tiles = ntiles;
    }
    launch_fill_ordered(L, ctx->recs.p, ctx->masks.p, ctx->obins.p, ctx->keys.p, p.rgb888 ? ctx->tex8desc.p : ctx->texdesc.p,
                        p.rgb888 ? reinterpret_cast<const uint16_t*>(ctx->texels8.p) : ctx->texels.p,
                        ctx->fb_rgba.p, ctx->fb_z.p, ctx->state, ctx->sticky, p, ordered_scratch_cap(ctx, ntiles));
    CK(cudaGetLastError());
    return B32_OK;
}

// Spin until the kernel's last CTA has published `seq` in the mapped status block (b32_device.cuh, HostStatus).
// Every few thousand polls the stream is queried: a failed launch / kernel must not hang the caller.
int wait_stamp(b32_ctx* ctx, int k, uint32_t seq) {
    const uint32_t* flag = &ctx->hstat->stamp[k].seq;
    for (uint32_t spins = 1;; ++spins) {
        if (__atomic_load_n(flag, __ATOMIC_ACQUIRE) == seq) return B32_OK;
        if ((spins & 0x3FFF) == 0) {
            cudaError_t e = cudaStreamQuery(ctx->stream);
            if (e == cudaSuccess) {                               // the stream drained: the flag is there, or the kernel never ran
                if (__atomic_load_n(flag, __ATOMIC_ACQUIRE) == seq) return B32_OK;
                return fail(ctx, B32_ERR_CUDA, "a frame kernel finished without reporting (launch failed?)");
            }
            if (e != cudaErrorNotReady) { cudaGetLastError(); return fail(ctx, B32_ERR_CUDA, cudaGetErrorString(e)); }
        }
#if defined(__x86_64__)
        __builtin_ia32_pause();
#endif
    }
}
static float stamp_ms(const b32_ctx* ctx, int k) {
    const KernelStamp& t = ctx->hstat->stamp[k];
    return t.t1 > t.t0 ? (float)((double)(t.t1 - t.t0) * 1e-6) : 0.0f;
}

// Turn the frame just captured on the context's stream into an executable graph and remember its kernel nodes.
bool finish_capture(b32_ctx* ctx, FrameGraph& fg, const FrameKey& key) {
    cudaGraph_t graph = nullptr;
    if (cudaStreamEndCapture(ctx->stream, &graph) != cudaSuccess || !graph) { cudaGetLastError(); return false; }
    fg.destroy();
    fg.graph = graph;
    if (cudaGraphInstantiate(&fg.patch.exec, graph, 0) != cudaSuccess) { cudaGetLastError(); fg.destroy(); return false; }
    cudaGraphNode_t nodes[16]; size_t n = 16;
    if (cudaGraphGetNodes(graph, nodes, &n) != cudaSuccess || n > 8) { cudaGetLastError(); fg.destroy(); return false; }
    fg.patch.n = 0;
    for (size_t i = 0; i < n; ++i) {
        cudaGraphNodeType t;
        cudaKernelNodeParams np{};
        if (cudaGraphNodeGetType(nodes[i], &t) != cudaSuccess || t != cudaGraphNodeTypeKernel ||
            cudaGraphKernelNodeGetParams(nodes[i], &np) != cudaSuccess) { cudaGetLastError(); fg.destroy(); return false; }
        fg.patch.node[fg.patch.n] = nodes[i]; fg.patch.func[fg.patch.n] = np.func; ++fg.patch.n;
    }
    fg.key = key;
    return true;
}

// Everything the kernels of one frame need besides the context's own buffers.
struct FrameArgs {
    const b32_vertex* verts; const b32_face* faces;
    const TexDev* texdesc; const uint16_t* texels; const uint32_t* texmask;
    bool clear; uint32_t clear_color;
    uint32_t* zero_next;             // the CallState + tile-counter set this frame's k_setup zeroes for the next call
    CallParams p;
};

// The kernels of one frame on L: Framebuffer::clear (optional), TRANSFORM + CULL + setup + binning, DRAW pass 1.
// L decides how: direct launches, launches into a capturing stream, or patches of an instantiated graph's nodes.
// ev_setup / ev_fill (nullable) are recorded in front of the two kernel groups (synchronous calls time them).
int launch_frame(b32_ctx* ctx, const LaunchCtx& L, const FrameArgs& a, cudaEvent_t ev_setup, cudaEvent_t ev_fill, cudaEvent_t ev_end = nullptr) {
    const CallParams& p = a.p;
    if (ev_setup) CK(cudaEventRecord(ev_setup, L.stream));
    const bool wire_on = p.wire_back || p.wire_front;
    const TVert* tv = nullptr;
    if (split_transform() && p.nf) { launch_transform(L, a.verts, ctx->tv.p, nullptr, p); tv = ctx->tv.p; }
    launch_setup(L, a.verts, a.faces, tv, a.texdesc, ctx->lights.p, ctx->recs.p, ctx->keys.p, ctx->heads.p, ctx->masks.p,
                 ctx->tile_count, wire_on ? ctx->wire.p : nullptr, ctx->state, a.zero_next, ctx->state_stride,
                 ctx->fb_rgba.p, ctx->fb_z.p, a.clear ? ctx->width * ctx->height : 0u, a.clear_color, p);    // the frame's clear rides in k_setup
    if (ev_fill) CK(cudaEventRecord(ev_fill, L.stream));
    if (!p.wire_front) {
        const bool pass1 = !(p.xray_mode && !p.rgb888);     // x-ray: every surface goes through the ordered replay
        if (pass1)
            launch_fill_opaque(L, ctx->recs.p, ctx->masks.p, ctx->heads.p, a.texdesc, a.texels, a.texmask,
                               ctx->fb_rgba.p, ctx->fb_z.p, ctx->state, ctx->sticky, ctx->crowd.p, (uint32_t)std::min<size_t>(ctx->crowd.cap, 0xFFFFFFFFu), p);
        if (p.enq_ordered)                                  // enqueue-only frames that may hold pass-2 surfaces: no host round trip
            launch_fill_ordered(L, ctx->recs.p, ctx->masks.p, ctx->obins.p, ctx->keys.p, a.texdesc, a.texels,
                                ctx->fb_rgba.p, ctx->fb_z.p, ctx->state, ctx->sticky, p, ordered_scratch_cap(ctx, p.tiles_x * p.tiles_y));
    }
    if (ev_end) CK(cudaEventRecord(ev_end, L.stream));
    return B32_OK;
}

// Enqueue-only frames: the second frame of one topology is captured into a CUDA graph, later ones re-parameterise
// its kernel nodes in place and launch it — one driver call per frame instead of one per kernel.
enum class Submit { PLAIN, CAPTURE, PATCH };

Submit pick_submit_mode(b32_ctx* ctx, const FrameKey& key, FrameGraph*& fg) {
    static const bool no_graph = std::getenv("B32_NO_GRAPH") != nullptr;      // experiments: plain launches only
    fg = nullptr;
    if (no_graph) return Submit::PLAIN;
    for (FrameGraph& g : ctx->fgs) if (g.patch.exec && g.key == key) { fg = &g; return Submit::PATCH; }
    bool seen = false;
    for (const FrameKey& k : ctx->fg_seen) seen = seen || k == key;
    // a caller that cycles through more topologies than the cache holds would capture forever: stop building
    // graphs once captures stop paying for themselves (plain launches are always correct)
    const bool thrashing = ctx->graph_captures > 2 * b32_ctx::N_FRAME_GRAPHS && ctx->graph_launches < 4 * ctx->graph_captures;
    if (seen && !thrashing) {
        fg = &ctx->fgs[ctx->fg_next];
        ctx->fg_next = (ctx->fg_next + 1) % b32_ctx::N_FRAME_GRAPHS;
        ++ctx->graph_captures;
        return Submit::CAPTURE;
    }
    ctx->fg_seen[ctx->fg_seen_next] = key;
    ctx->fg_seen_next = (ctx->fg_seen_next + 1) % b32_ctx::N_FRAME_GRAPHS;
    return Submit::PLAIN;
}

int enqueue_frame(b32_ctx* ctx, const FrameArgs& a, const FrameKey& key) {
    cudaStream_t st = ctx->stream;
    LaunchCtx L = ctx->L();
    FrameGraph* fg = nullptr;
    Submit mode = pick_submit_mode(ctx, key, fg);
    ctx->async_pending = true;
    if (ctx->tring_n) {                                    // measurement: plain launches, timed per kernel group
        cudaEvent_t* e = &ctx->tring[(size_t)(ctx->tring_count % ctx->tring_n) * 3];
        ++ctx->tring_count;
        return launch_frame(ctx, L, a, e[0], e[1], e[2]);
    }
    if (mode == Submit::CAPTURE) {
        if (cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed) == cudaSuccess) {
            launch_frame(ctx, L, a, nullptr, nullptr);
            if (finish_capture(ctx, *fg, key)) { CK(cudaGraphLaunch(fg->patch.exec, st)); ++ctx->graph_launches; return B32_OK; }
        }
        cudaGetLastError();                                // could not build the graph: launch this frame the ordinary way
    } else if (mode == Submit::PATCH) {
        LaunchCtx P = L;
        P.patch = &fg->patch; fg->patch.err = cudaSuccess;
        for (bool& u : fg->patch.used) u = false;
        launch_frame(ctx, P, a, nullptr, nullptr);
        if (fg->patch.err == cudaSuccess) { CK(cudaGraphLaunch(fg->patch.exec, st)); ++ctx->graph_launches; return B32_OK; }
        fg->destroy(); cudaGetLastError();                 // topology changed under the key: fall back to plain launches
    }
    return launch_frame(ctx, L, a, nullptr, nullptr);
}

// One render_mesh_15 / render_mesh (render.rs:2302-2572 / :1971-2259) on device-resident geometry.
//   wait=true : returns when the frame is in the framebuffer; fills *tm.
//   wait=false: only enqueues (no host round trip).  all_opaque = the caller promises there is no pass 2: the ordered
//               replay is not even launched; otherwise it is enqueued behind pass 1 and exits at once when it has nothing to do.
// clear_rgba (nullable): Framebuffer::clear first, as part of the same frame.
int render_device(b32_ctx* ctx, const b32_vertex* d_verts, uint32_t nv, const b32_face* d_faces, uint32_t nf,
                  const b32_camera* cam, const b32_settings* s, const b32_fog* fog, b32_timings* tm, bool wait, bool rgb888 = false,
                  const uint8_t* clear_rgba = nullptr, bool all_opaque = false, uint32_t vwords = 9, uint32_t faces_implicit = 0, uint32_t uniform_flags = 0) {
    if (!cam || !s) return fail(ctx, B32_ERR_INVALID, "camera/settings is NULL");
    if (ctx->width == 0 || ctx->height == 0) return fail(ctx, B32_ERR_INVALID, "framebuffer has zero size (call b32_fb_resize)");
    FrameArgs a{};
    CallParams& p = a.p;
    std::vector<LightDev> lights;
    int rc = fill_params(ctx, p, cam, s, fog, nv, nf, lights, rgb888);
    if (rc != B32_OK) return rc;
    p.vwords = (uint8_t)vwords; p.faces_implicit = (uint8_t)faces_implicit; p.uniform_flags = uniform_flags;
    if (tm) std::memset(tm, 0, sizeof(*tm));
    a.verts = d_verts; a.faces = d_faces;
    a.texdesc = rgb888 ? ctx->tex8desc.p : ctx->texdesc.p;
    a.texels = rgb888 ? reinterpret_cast<const uint16_t*>(ctx->texels8.p) : ctx->texels.p;
    a.texmask = rgb888 ? ctx->tex8mask.p : ctx->texmask.p;
    a.clear = clear_rgba != nullptr;
    a.clear_color = clear_rgba ? ((uint32_t)clear_rgba[0] | ((uint32_t)clear_rgba[1] << 8) | ((uint32_t)clear_rgba[2] << 16) |
                                  ((uint32_t)clear_rgba[3] << 24)) : 0u;
    for (float& k : ctx->kernel_ms) k = 0.0f;
    ctx->last_nf = nf;
    ctx->last_params = p;
    if (nf == 0) {
        if (a.clear) launch_fb_clear(ctx->L(), ctx->fb_rgba.p, ctx->fb_z.p, ctx->width * ctx->height, a.clear_color);
        return B32_OK;
    }

    rc = ensure_work(ctx, p); if (rc) return rc;
    rc = upload_lights(ctx, lights); if (rc) return rc;
    LaunchCtx L = ctx->L();
    cudaStream_t st = ctx->stream;
    CallState hs{};
    p.async_call = wait ? 0 : 1;
    p.enq_ordered = (!wait && !all_opaque) ? 1 : 0;
    ctx->last_params = p;
    if (p.wire_back || p.wire_front) CK(ctx->wire.reserve(nf));
    // take the set the previous call's k_setup zeroed; this call's k_setup zeroes the other one
    // (nothing between here and the launch can fail, so the two sets never get out of step)
    ctx->state_cur ^= 1;
    ctx->state = reinterpret_cast<CallState*>(ctx->state_ring.p + (size_t)ctx->state_cur * ctx->state_stride);
    ctx->tile_count = ctx->state_ring.p + (size_t)ctx->state_cur * ctx->state_stride + STATE_WORDS;
    a.zero_next = ctx->state_ring.p + (size_t)(ctx->state_cur ^ 1) * ctx->state_stride;
    if (!wait) {
        FrameKey key;
        key.verts = d_verts; key.faces = d_faces; key.nv = nv; key.nf = nf; key.width = ctx->width; key.height = ctx->height;
        key.rgb888 = rgb888; key.pass1 = !((p.xray_mode && !rgb888) || p.wire_front); key.clear = a.clear; key.valid = 1;
        key.ordered = p.enq_ordered; key.spot = p.has_spot; key.prefix = fill_uses_edge_prefix(p) || p.prefer_prefix;
        if (ctx->frame_timings) {                          // the frame's kernels publish counters + times (b32_frame_timings reads them)
            const uint32_t slot = ctx->fslot_next++ % b32_ctx::N_FRAME_STATUS;
            b32_ctx::FrameSlot& fs = ctx->fslot[slot];
            fs.seq = ++ctx->host_seq; fs.pass1 = key.pass1; fs.ordered = p.enq_ordered && !p.wire_front;
            fs.last = fs.ordered ? HS_ORDERED : fs.pass1 ? HS_FILL : HS_SETUP;
            p.host = ctx->fstat_dev + slot; p.host_seq = fs.seq;
            ctx->last_params = p;
        }
        return enqueue_frame(ctx, a, key);
    }
    // Blocking call: no events, no copy back.  The kernels report in host-mapped memory (HostStatus): k_setup's last CTA
    // the counters — the host reads them while pass 1 runs and launches the ordered pass behind it if there is one — and
    // every kernel its start / end times.
    const uint32_t seq = ++ctx->host_seq;
    p.host = ctx->hstat_dev; p.host_seq = seq;
    ctx->last_params = p;
    for (KernelStamp& t : ctx->hstat->stamp) { t.t0 = 0; t.t1 = 0; }
    rc = launch_frame(ctx, L, a, nullptr, nullptr); if (rc) return rc;
    CK(cudaGetLastError());
    rc = wait_stamp(ctx, HS_SETUP, seq); if (rc) return rc;
    hs = ctx->hstat->state;
    if (hs.oob == 2) return fail(ctx, B32_ERR_INVALID, "face blend mode out of range (not a BlendMode)");
    if (hs.oob) return fail(ctx, B32_ERR_OOB_INDEX, "face vertex index out of range (reference: slice index panic)");
    {   // the reference panics on a NaN key in a sorted slice of length >= 2 (render.rs:2531; RGB888: one list, :2161)
        bool nan_abort = rgb888 ? (!p.use_zbuffer && hs.nan_opaque && hs.n_opaque + hs.n_transp >= 2)
                                : (hs.nan_transp && hs.n_transp >= 2) || (!p.use_zbuffer && hs.nan_opaque && hs.n_opaque >= 2);
        if (nan_abort) return fail(ctx, B32_ERR_NAN_DEPTH, "NaN depth key in a sorted pass (reference: partial_cmp().unwrap() panic, render.rs:2531)");
    }
    // RGB888: one surface that may read the framebuffer sends the whole list through the ordered replay (pass 1 was skipped)
    const bool all_ordered = rgb888 ? hs.n_transp > 0 : p.xray_mode != 0;
    const bool pass1 = !p.wire_front && !(p.xray_mode && !rgb888);
    const bool need_ordered = !p.wire_front && (all_ordered ? (hs.n_opaque + hs.n_transp) > 0 : (!rgb888 && hs.n_transp > 0));
    if (need_ordered) { rc = render_ordered(ctx, p, hs.obin_max); if (rc) return rc; }
    if (need_ordered || pass1) { rc = wait_stamp(ctx, need_ordered ? HS_ORDERED : HS_FILL, seq); if (rc) return rc; }
    ctx->kernel_ms[0] = stamp_ms(ctx, HS_SETUP);
    ctx->kernel_ms[1] = pass1 ? stamp_ms(ctx, HS_FILL) : 0.0f;
    ctx->kernel_ms[3] = need_ordered ? stamp_ms(ctx, HS_ORDERED) : 0.0f;
    bool wire_too_long = false;
    if (p.wire_back || p.wire_front) {          // WIREFRAME phase, render.rs:2574-2635
        uint32_t tsize = 64;
        while (tsize < 6ull * nf && tsize < 0x80000000u) tsize <<= 1;             // load factor <= 1/2 with 3 edges per face
        CK(ctx->wire_table.reserve(tsize));
        CK(cudaEventRecord(ctx->ev[0], st));
        if (p.wire_back) launch_wire(L, ctx->wire.p, 1, 80u | (80u << 8) | (100u << 16) | 0xFF000000u, true, ctx->wire_table.p, tsize,
                                     ctx->fb_rgba.p, ctx->fb_z.p, ctx->state, p);
        if (p.wire_front) launch_wire(L, ctx->wire.p, 2, 200u | (200u << 8) | (220u << 16) | 0xFF000000u, false, ctx->wire_table.p, tsize,
                                      ctx->fb_rgba.p, ctx->fb_z.p, ctx->state, p);
        CK(cudaEventRecord(ctx->ev[1], st));
        CK(cudaMemcpyAsync(ctx->state_h, ctx->state, sizeof(CallState), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        CK(cudaGetLastError());
        cudaEventElapsedTime(&ctx->wire_ms, ctx->ev[0], ctx->ev[1]);
        wire_too_long = ctx->state_h->wire_too_long != 0;
    } else ctx->wire_ms = 0.0f;
    if (tm) {
        const float* k = ctx->kernel_ms;
        tm->transform_ms = 0.0f;                    // the transform is fused into the cull/setup kernel
        tm->cull_ms = k[0];                         // ... as is fog (fog_ms stays 0)
        tm->sort_ms = 0.0f;                         // pass 1 needs no sort; pass 2 sorts per tile inside its fill kernel
        tm->draw_ms = k[1] + k[2] + k[3];
        tm->wireframe_ms = ctx->wire_ms;
        tm->triangles_drawn = hs.n_opaque + hs.n_transp;                            // render.rs:2545
    }
    if (wire_too_long)
        return fail(ctx, B32_ERR_UNSUPPORTED, "wireframe phase: an edge longer than 2^24 pixels was not drawn (non-finite or absurd screen coordinates)");
    return B32_OK;
}

}  // namespace

// =================================================================================================
extern "C" {

int b32_ctx_create(int device, b32_ctx** out) {
    if (!out) return B32_ERR_INVALID;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n) { cudaGetLastError(); return B32_ERR_NO_DEVICE; }
    b32_ctx* ctx = new b32_ctx();
    ctx->device = device;
    auto bail = [&](cudaError_t e, const char* what) { fprintf(stderr, "b32_ctx_create: %s: %s\n", what, cudaGetErrorString(e)); delete ctx; return B32_ERR_CUDA; };
    cudaError_t e;
    if ((e = cudaSetDevice(device)) != cudaSuccess) return bail(e, "cudaSetDevice");
    if ((e = cudaGetDeviceProperties(&ctx->prop, device)) != cudaSuccess) return bail(e, "cudaGetDeviceProperties");
    if ((e = (cudaError_t)init_kernel_attributes()) != cudaSuccess) return bail(e, "cudaFuncSetAttribute");
    if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) return bail(e, "cudaStreamCreate");
    for (auto& ev : ctx->ev) if ((e = cudaEventCreate(&ev)) != cudaSuccess) return bail(e, "cudaEventCreate");
    for (auto& ev : ctx->h2d_ev) if ((e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)) != cudaSuccess) return bail(e, "cudaEventCreate");
    if ((e = cudaMalloc(&ctx->sticky, sizeof(uint32_t))) != cudaSuccess) return bail(e, "cudaMalloc");
    if ((e = cudaMemset(ctx->sticky, 0, sizeof(uint32_t))) != cudaSuccess) return bail(e, "cudaMemset");
    if ((e = cudaMallocHost(&ctx->state_h, sizeof(CallState))) != cudaSuccess) return bail(e, "cudaMallocHost");
    if ((e = cudaHostAlloc(&ctx->hstat, sizeof(HostStatus) + 64, cudaHostAllocMapped)) != cudaSuccess) return bail(e, "cudaHostAlloc");
    std::memset(ctx->hstat, 0, sizeof(HostStatus) + 64);
    if ((e = cudaHostGetDevicePointer(&ctx->hstat_dev, ctx->hstat, 0)) != cudaSuccess) return bail(e, "cudaHostGetDevicePointer");
    ctx->stepped_seq = reinterpret_cast<uint32_t*>(ctx->hstat + 1);               // the word behind the status block
    ctx->stepped_seq_dev = reinterpret_cast<uint32_t*>(ctx->hstat_dev + 1);
    ctx->pinned_bytes = 8u << 20;
    if ((e = cudaMallocHost(&ctx->pinned, ctx->pinned_bytes)) != cudaSuccess) return bail(e, "cudaMallocHost");
    ctx->texdesc.reserve(1); ctx->texels.reserve(1); ctx->texmask.reserve(4); ctx->lights.reserve(1);
    ctx->tex8desc.reserve(1); ctx->texels8.reserve(1); ctx->tex8mask.reserve(4);
    *out = ctx;
    return B32_OK;
}

void b32_ctx_destroy(b32_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    ctx->fb_rgba.release(); ctx->fb_z.release(); ctx->texels.release(); ctx->texmask.release(); ctx->texdesc.release();
    ctx->texels8.release(); ctx->tex8mask.release(); ctx->tex8desc.release(); ctx->verts.release(); ctx->faces.release();
    ctx->tv.release(); ctx->recs.release(); ctx->keys.release(); ctx->state_ring.release(); ctx->masks.release();
    ctx->heads.release(); ctx->obins.release(); ctx->crowd.release(); ctx->wire.release(); ctx->wire_table.release();
    ctx->lights.release(); ctx->dbg.release(); ctx->lines.release(); ctx->line_scratch.release();
    for (FrameGraph& g : ctx->fgs) g.destroy();
    for (cudaEvent_t e : ctx->tring) cudaEventDestroy(e);
    if (ctx->sticky) cudaFree(ctx->sticky);
    if (ctx->state_h) cudaFreeHost(ctx->state_h);
    if (ctx->hstat) cudaFreeHost(ctx->hstat);
    if (ctx->fstat) cudaFreeHost(ctx->fstat);
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    for (auto& ev : ctx->ev) if (ev) cudaEventDestroy(ev);
    for (auto& ev : ctx->h2d_ev) if (ev) cudaEventDestroy(ev);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char* b32_last_error(const b32_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
void* b32_ctx_stream(b32_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
uint64_t b32_kernel_launches(const b32_ctx* ctx) { return ctx ? ctx->launches : 0; }

// errors of calls that were only enqueued surface here (the device kept a sticky error word)
static int collect_async(b32_ctx* ctx) {
    if (!ctx->async_pending) return B32_OK;
    ctx->async_pending = false;
    uint32_t sticky = 0;
    CK(cudaMemcpyAsync(ctx->state_h, ctx->sticky, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    sticky = *reinterpret_cast<uint32_t*>(ctx->state_h);
    if (!sticky) return B32_OK;
    CK(cudaMemsetAsync(ctx->sticky, 0, sizeof(uint32_t), ctx->stream));
    if (sticky & 32u) return fail(ctx, B32_ERR_INVALID, "an enqueued call had a face whose blend mode is not a BlendMode");
    if (sticky & 1u) return fail(ctx, B32_ERR_OOB_INDEX, "an enqueued call had a face vertex index out of range");
    if (sticky & 2u) return fail(ctx, B32_ERR_NAN_DEPTH, "an enqueued call had a NaN depth key in a sorted pass");
    if (sticky & 8u) return fail(ctx, B32_ERR_INVALID, "an enqueued call had semi-transparent surfaces (pass 2 was not drawn): B32_RENDER_ALL_OPAQUE was wrong");
    return fail(ctx, B32_ERR_UNSUPPORTED, "an enqueued call had a tile with more semi-transparent surfaces than fit shared memory (pass 2 was not drawn): "
                                          "make one blocking call with this mesh first (it sizes the scratch)");
}

int b32_sync(b32_ctx* ctx) {
    USE_DEVICE(ctx);
    if (!ctx) return B32_ERR_INVALID;
    int rc = collect_async(ctx);
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaGetLastError());
    return rc;
}

int b32_fb_resize(b32_ctx* ctx, uint32_t width, uint32_t height) {
    USE_DEVICE(ctx);
    if (!ctx) return B32_ERR_INVALID;
    if (width == ctx->width && height == ctx->height) return B32_OK;          // render.rs:28
    if (width > 65535 || height > 65535) return fail(ctx, B32_ERR_UNSUPPORTED, "framebuffer dimension > 65535");
    CK(cudaStreamSynchronize(ctx->stream));
    size_t n = (size_t)width * height;
    CK(ctx->fb_rgba.reserve(std::max<size_t>(n, 1)));
    CK(ctx->fb_z.reserve(std::max<size_t>(n, 1)));
    ctx->width = width; ctx->height = height;
    launch_fb_clear(ctx->L(), ctx->fb_rgba.p, ctx->fb_z.p, (uint32_t)n, 0u);   // pixels 0, zbuffer f32::MAX (render.rs:31-32)
    return B32_OK;
}

int b32_fb_clear(b32_ctx* ctx, uint8_t r, uint8_t g, uint8_t b, uint8_t a) {
    USE_DEVICE(ctx);
    if (!ctx) return B32_ERR_INVALID;
    uint32_t c = (uint32_t)r | ((uint32_t)g << 8) | ((uint32_t)b << 16) | ((uint32_t)a << 24);
    launch_fb_clear(ctx->L(), ctx->fb_rgba.p, ctx->fb_z.p, ctx->width * ctx->height, c);
    CK(cudaGetLastError());
    return B32_OK;
}

int b32_fb_clear_gradient(b32_ctx* ctx, uint8_t tr, uint8_t tg, uint8_t tb, uint8_t br, uint8_t bg, uint8_t bb, uint8_t a) {
    USE_DEVICE(ctx);
    if (!ctx) return B32_ERR_INVALID;
    uint32_t top = (uint32_t)tr | ((uint32_t)tg << 8) | ((uint32_t)tb << 16) | ((uint32_t)a << 24);
    uint32_t bottom = (uint32_t)br | ((uint32_t)bg << 8) | ((uint32_t)bb << 16);
    launch_fb_clear_gradient(ctx->L(), ctx->fb_rgba.p, ctx->fb_z.p, ctx->width, ctx->height, top, bottom);
    CK(cudaGetLastError());
    return B32_OK;
}

int b32_fb_upload(b32_ctx* ctx, const uint8_t* rgba, const float* z) {
    USE_DEVICE(ctx);
    if (!ctx || !rgba) return B32_ERR_INVALID;
    size_t n = (size_t)ctx->width * ctx->height;
    int rc = h2d(ctx, ctx->fb_rgba.p, rgba, n * 4); if (rc) return rc;
    if (z) { rc = h2d(ctx, ctx->fb_z.p, z, n * 4); if (rc) return rc; }
    CK(cudaStreamSynchronize(ctx->stream));
    return B32_OK;
}

int b32_fb_download(b32_ctx* ctx, uint8_t* rgba, float* z) {
    USE_DEVICE(ctx);
    if (!ctx) return B32_ERR_INVALID;
    size_t n = (size_t)ctx->width * ctx->height;
    if (rgba) CK(cudaMemcpyAsync(rgba, ctx->fb_rgba.p, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (z) CK(cudaMemcpyAsync(z, ctx->fb_z.p, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return collect_async(ctx);
}

int b32_fb_size(const b32_ctx* ctx, uint32_t* width, uint32_t* height) {
    if (!ctx) return B32_ERR_INVALID;
    if (width) *width = ctx->width;
    if (height) *height = ctx->height;
    return B32_OK;
}

int b32_textures_set(b32_ctx* ctx, const b32_tex_desc* descs, uint32_t n) {
    USE_DEVICE(ctx);
    if (!ctx || (n && !descs)) return B32_ERR_INVALID;
    if (n > 0xFFFF) return fail(ctx, B32_ERR_UNSUPPORTED, "more than 65535 textures (face.flags carries a 16-bit texture id)");
    CK(cudaStreamSynchronize(ctx->stream));
    size_t total = 0, max_idx_bytes = 0;
    for (uint32_t i = 0; i < n; ++i) {
        const b32_tex_desc& d = descs[i];
        if (d.format > B32_TEX_IDX4 || d.blend_mode > B32_BLEND_ERASE) return fail(ctx, B32_ERR_INVALID, "texture format/blend out of range");
        size_t px = (size_t)d.width * d.height;
        if (px && !d.pixels) return fail(ctx, B32_ERR_INVALID, "texture pixels is NULL");
        if (d.format != B32_TEX_RGB555 && d.clut_len && !d.clut) return fail(ctx, B32_ERR_INVALID, "texture clut is NULL");
        total += px;
        if (d.format == B32_TEX_IDX8) max_idx_bytes = std::max(max_idx_bytes, px);
        if (d.format == B32_TEX_IDX4) max_idx_bytes = std::max(max_idx_bytes, (px + 1) / 2);
    }
    if (total > 0xFFFFFFFFull) return fail(ctx, B32_ERR_UNSUPPORTED, "texel pool exceeds 2^32 texels");
    CK(ctx->texels.reserve(std::max<size_t>(total, 1)));
    CK(ctx->texdesc.reserve(std::max<uint32_t>(n, 1)));
    ctx->texdesc_h.resize(n);
    DevBuf<uint8_t> idx; DevBuf<uint16_t> clut;
    if (max_idx_bytes) { CK(idx.reserve(max_idx_bytes)); CK(clut.reserve(256)); }
    size_t off = 0;
    int rc = B32_OK;
    for (uint32_t i = 0; i < n && rc == B32_OK; ++i) {
        const b32_tex_desc& d = descs[i];
        size_t px = (size_t)d.width * d.height;
        ctx->texdesc_h[i] = TexDev{(uint32_t)off, d.width, d.height, d.blend_mode};
        if (px) {
            if (d.format == B32_TEX_RGB555) {
                rc = h2d(ctx, ctx->texels.p + off, d.pixels, px * 2);
            } else {
                size_t bytes = d.format == B32_TEX_IDX8 ? px : (px + 1) / 2;
                uint32_t clen = std::min<uint32_t>(d.clut_len, 256);
                rc = h2d(ctx, idx.p, d.pixels, bytes);
                if (rc == B32_OK && clen) rc = h2d(ctx, clut.p, d.clut, clen * 2);
                if (rc == B32_OK) launch_tex_expand(ctx->L(), idx.p, clut.p, clen, d.format, (uint32_t)px, ctx->texels.p + off);
                cudaStreamSynchronize(ctx->stream);      // idx/clut staging is reused by the next texture
            }
        }
        off += px;
    }
    idx.release(); clut.release();
    if (rc) return rc;
    {   // the visibility walk of k_fill_opaque asks "does this texel write?" of a 1-bit mask, not of the texel
        uint32_t words = (uint32_t)(((total + 31) / 32 + 3) & ~(size_t)3);
        CK(ctx->texmask.reserve(std::max<uint32_t>(words, 4)));
        launch_tex_mask(ctx->L(), ctx->texels.p, (uint32_t)total, words, ctx->texmask.p);
        ctx->texmask_words = words;
    }
    if (n) CK(cudaMemcpyAsync(ctx->texdesc.p, ctx->texdesc_h.data(), n * sizeof(TexDev), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->ntex = n;
    return B32_OK;
}

int b32_textures_set_rgb888(b32_ctx* ctx, const b32_tex8_desc* descs, uint32_t n) {
    USE_DEVICE(ctx);
    if (!ctx || (n && !descs)) return B32_ERR_INVALID;
    if (n > 0xFFFF) return fail(ctx, B32_ERR_UNSUPPORTED, "more than 65535 textures (face.flags carries a 16-bit texture id)");
    CK(cudaStreamSynchronize(ctx->stream));
    size_t total = 0, max_px = 0;
    for (uint32_t i = 0; i < n; ++i) {
        size_t px = (size_t)descs[i].width * descs[i].height;
        if (descs[i].blend_mode > B32_BLEND_ERASE) return fail(ctx, B32_ERR_INVALID, "texture blend out of range");
        if (px && !descs[i].pixels) return fail(ctx, B32_ERR_INVALID, "texture pixels is NULL");
        total += px; max_px = std::max(max_px, px);
    }
    if (total > 0xFFFFFFFFull) return fail(ctx, B32_ERR_UNSUPPORTED, "texel pool exceeds 2^32 texels");
    CK(ctx->texels8.reserve(std::max<size_t>(total, 1)));
    CK(ctx->tex8desc.reserve(std::max<uint32_t>(n, 1)));
    std::vector<TexDev> desc_h(n);
    size_t off = 0;
    for (uint32_t i = 0; i < n; ++i) {
        size_t px = (size_t)descs[i].width * descs[i].height;
        desc_h[i] = TexDev{(uint32_t)off, descs[i].width, descs[i].height, 0u};       // .blend is computed on the device below
        int rc = h2d(ctx, ctx->texels8.p + off, descs[i].pixels, px * 4); if (rc) return rc;
        off += px;
    }
    if (n) CK(cudaMemcpyAsync(ctx->tex8desc.p, desc_h.data(), n * sizeof(TexDev), cudaMemcpyHostToDevice, ctx->stream));
    uint32_t words = (uint32_t)(((total + 31) / 32 + 3) & ~(size_t)3);
    CK(ctx->tex8mask.reserve(std::max<uint32_t>(words, 4)));
    launch_tex8_scan(ctx->L(), ctx->texels8.p, (uint32_t)total, words, ctx->tex8mask.p, ctx->tex8desc.p, n, (uint32_t)std::min<size_t>(max_px, 0xFFFFFFFFu));
    CK(cudaStreamSynchronize(ctx->stream));      // desc_h is a host temporary
    CK(cudaGetLastError());
    ctx->tex8mask_words = words;
    ctx->ntex8 = n;
    return B32_OK;
}

int b32_render_mesh(b32_ctx* ctx, const b32_vertex* vertices, uint32_t nv, const b32_face* faces, uint32_t nf,
                    const b32_camera* camera, const b32_settings* settings, b32_timings* timings) {
    USE_DEVICE(ctx);
    if (!ctx) return B32_ERR_INVALID;
    if ((nv && !vertices) || (nf && !faces)) return fail(ctx, B32_ERR_INVALID, "vertices/faces is NULL");
    CK(ctx->verts.reserve((size_t)std::max<uint32_t>(nv, 1) + 1));      // + 1: k_setup's staged window may overrun by < 16 bytes
    CK(ctx->faces.reserve(std::max<uint32_t>(nf, 1)));
    int rc = h2d(ctx, ctx->verts.p, vertices, (size_t)nv * sizeof(b32_vertex)); if (rc) return rc;
    rc = h2d(ctx, ctx->faces.p, faces, (size_t)nf * sizeof(b32_face)); if (rc) return rc;
    return render_device(ctx, ctx->verts.p, nv, ctx->faces.p, nf, camera, settings, nullptr, timings, true, true);
}

int b32_render_mesh_resident(b32_ctx* ctx, const b32_mesh* mesh, const b32_camera* camera, const b32_settings* settings, b32_timings* timings) {
    USE_DEVICE(ctx);
    if (!ctx || !mesh) return B32_ERR_INVALID;
    return render_device(ctx, mesh->verts, mesh->nv, mesh->faces, mesh->nf, camera, settings, nullptr, timings, true, true);
}

int b32_render_mesh_15(b32_ctx* ctx, const b32_vertex* vertices, uint32_t nv, const b32_face* faces, uint32_t nf,
                       const b32_camera* camera, const b32_settings* settings, const b32_fog* fog, b32_timings* timings) {
    USE_DEVICE(ctx);
    if (!ctx) return B32_ERR_INVALID;
    if ((nv && !vertices) || (nf && !faces)) return fail(ctx, B32_ERR_INVALID, "vertices/faces is NULL");
    CK(ctx->verts.reserve((size_t)std::max<uint32_t>(nv, 1) + 1));      // + 1: k_setup's staged window may overrun by < 16 bytes
    CK(ctx->faces.reserve(std::max<uint32_t>(nf, 1)));
    int rc = h2d(ctx, ctx->verts.p, vertices, (size_t)nv * sizeof(b32_vertex)); if (rc) return rc;
    rc = h2d(ctx, ctx->faces.p, faces, (size_t)nf * sizeof(b32_face)); if (rc) return rc;
    return render_device(ctx, ctx->verts.p, nv, ctx->faces.p, nf, camera, settings, fog, timings, true);
}

int b32_render_mesh_15_ex(b32_ctx* ctx, const void* vertices, uint32_t nv, const void* faces, uint32_t nf,
                          const b32_camera* camera, const b32_settings* settings, const b32_fog* fog, uint32_t flags, b32_timings* timings) {
    USE_DEVICE(ctx);
    if (!ctx) return B32_ERR_INVALID;
    if ((nv && !vertices) || (nf && !faces) || !settings) return fail(ctx, B32_ERR_INVALID, "vertices/faces/settings is NULL");
    const bool async = (flags & B32_RENDER_ASYNC) != 0, no_normal = (flags & B32_VTX_NO_NORMAL) != 0;
    const bool uniform = (flags & B32_FACES_UNIFORM) != 0, implicit = uniform || (flags & B32_FACES_IMPLICIT) != 0;
    if (no_normal && settings->shading != B32_SHADE_NONE) return fail(ctx, B32_ERR_INVALID, "B32_VTX_NO_NORMAL needs settings.shading == None");
    if (implicit && (uint64_t)nv < 3ull * nf) return fail(ctx, B32_ERR_OOB_INDEX, "B32_FACES_IMPLICIT / B32_FACES_UNIFORM: fewer than 3 * nf vertices");
    const bool wire = (settings->backface_cull && settings->backface_wireframe) || settings->wireframe_overlay;
    if (async && wire) return fail(ctx, B32_ERR_INVALID, "the wireframe phase cannot be enqueued");
    bool all_opaque = async && (flags & B32_RENDER_ALL_OPAQUE) != 0;
    if (all_opaque) {          // the promise is checked where that is free: settings and bound textures
        if (settings->xray_mode) return fail(ctx, B32_ERR_INVALID, "B32_RENDER_ALL_OPAQUE with x-ray mode");
        for (const TexDev& t : ctx->texdesc_h) if (t.blend != B32_BLEND_OPAQUE) return fail(ctx, B32_ERR_INVALID, "B32_RENDER_ALL_OPAQUE with a bound texture that has a blend mode");
    }
    const size_t vbytes = no_normal ? sizeof(b32_vertex_nn) : sizeof(b32_vertex), fbytes = implicit ? sizeof(uint32_t) : sizeof(b32_face);
    // staging buffers are sized in b32_vertex / b32_face units; + 1 vertex: k_setup's staged window may overrun by < 16 bytes
    CK(ctx->verts.reserve(((size_t)nv * vbytes + sizeof(b32_vertex) - 1) / sizeof(b32_vertex) + 1));
    int rc = h2d(ctx, ctx->verts.p, vertices, (size_t)nv * vbytes); if (rc) return rc;
    uint32_t uniform_flags = 0;
    if (uniform) {             // the one flags word travels with the kernel parameters: no face buffer, no copy
        uniform_flags = *static_cast<const uint32_t*>(faces);
    } else {
        CK(ctx->faces.reserve(std::max<size_t>(((size_t)nf * fbytes + sizeof(b32_face) - 1) / sizeof(b32_face), 1)));
        rc = h2d(ctx, ctx->faces.p, faces, (size_t)nf * fbytes); if (rc) return rc;
    }
    return render_device(ctx, ctx->verts.p, nv, uniform ? nullptr : ctx->faces.p, nf, camera, settings, fog, async ? nullptr : timings, !async, false, nullptr, all_opaque,
                         no_normal ? 6 : 9, uniform ? 2u : implicit ? 1u : 0u, uniform_flags);
}

int b32_frame_15_enqueue(b32_ctx* ctx, const uint8_t* clear_rgba, const b32_mesh* mesh, const b32_camera* camera,
                         const b32_settings* settings, const b32_fog* fog) {
    USE_DEVICE(ctx);
    if (!ctx || !mesh || !settings) return B32_ERR_INVALID;
    // only the wireframe phase needs the host (status read-back between its kernels): such a frame is rendered synchronously
    bool wire = (settings->backface_cull && settings->backface_wireframe) || settings->wireframe_overlay;
    bool may_blend = mesh->has_nonopaque || settings->xray_mode;
    for (const TexDev& t : ctx->texdesc_h) may_blend = may_blend || t.blend != B32_BLEND_OPAQUE;
    return render_device(ctx, mesh->verts, mesh->nv, mesh->faces, mesh->nf, camera, settings, fog, nullptr, wire, false, clear_rgba, !may_blend);
}

uint64_t b32_graph_launches(const b32_ctx* ctx) { return ctx ? ctx->graph_launches : 0; }

int b32_ctx_frame_timings(b32_ctx* ctx, int enable) {
    USE_DEVICE(ctx);
    if (!ctx) return B32_ERR_INVALID;
    if (enable && !ctx->fstat) {
        CK(cudaHostAlloc(&ctx->fstat, sizeof(HostStatus) * b32_ctx::N_FRAME_STATUS, cudaHostAllocMapped));
        std::memset(ctx->fstat, 0, sizeof(HostStatus) * b32_ctx::N_FRAME_STATUS);
        CK(cudaHostGetDevicePointer(&ctx->fstat_dev, ctx->fstat, 0));
    }
    ctx->frame_timings = enable != 0;
    return B32_OK;
}

int b32_frame_timings(b32_ctx* ctx, b32_timings* out) {
    if (!ctx || !out) return B32_ERR_INVALID;
    std::memset(out, 0, sizeof(*out));
    if (!ctx->fstat) return fail(ctx, B32_ERR_INVALID, "b32_ctx_frame_timings was not enabled");
    // newest first: the most recent enqueued frame whose last kernel has reported
    for (uint32_t back = 1; back <= b32_ctx::N_FRAME_STATUS && back <= ctx->fslot_next; ++back) {
        const uint32_t slot = (ctx->fslot_next - back) % b32_ctx::N_FRAME_STATUS;
        const b32_ctx::FrameSlot& fs = ctx->fslot[slot];
        const HostStatus& h = ctx->fstat[slot];
        if (!fs.seq || __atomic_load_n(&h.stamp[fs.last].seq, __ATOMIC_ACQUIRE) != fs.seq) continue;
        auto ms = [&](int k) { const KernelStamp& t = h.stamp[k]; return (t.seq == fs.seq && t.t1 > t.t0) ? (float)((double)(t.t1 - t.t0) * 1e-6) : 0.0f; };
        out->cull_ms = ms(HS_SETUP);                        // transform + cull + setup are one kernel (see b32_timings)
        out->draw_ms = (fs.pass1 ? ms(HS_FILL) : 0.0f) + (fs.ordered ? ms(HS_ORDERED) : 0.0f);
        out->triangles_drawn = h.state.n_opaque + h.state.n_transp;
        return B32_OK;
    }
    return B32_OK;                                          // nothing has finished yet: zeros
}

int b32_fb_download_async(b32_ctx* ctx, uint8_t* rgba, float* z) {
    USE_DEVICE(ctx);
    if (!ctx) return B32_ERR_INVALID;
    size_t n = (size_t)ctx->width * ctx->height;
    if (rgba) CK(cudaMemcpyAsync(rgba, ctx->fb_rgba.p, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (z) CK(cudaMemcpyAsync(z, ctx->fb_z.p, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    return B32_OK;
}

int b32_mesh_upload(b32_ctx* ctx, const b32_vertex* vertices, uint32_t nv, const b32_face* faces, uint32_t nf, b32_mesh** out) {
    USE_DEVICE(ctx);
    if (!ctx || !out) return B32_ERR_INVALID;
    if ((nv && !vertices) || (nf && !faces)) return fail(ctx, B32_ERR_INVALID, "vertices/faces is NULL");
    b32_mesh* m = new b32_mesh();
    m->nv = nv; m->nf = nf;
    for (uint32_t i = 0; i < nf; ++i) {
        uint32_t fl = faces[i].flags;
        if (((fl >> 16) & 7u) != B32_BLEND_OPAQUE || (fl >> 24) != 255u) { m->has_nonopaque = true; break; }
    }
    cudaError_t e = cudaMalloc(&m->verts, (size_t)nv * sizeof(b32_vertex) + 16);      // + 16: k_setup's staged window may overrun by < 16 bytes
    if (e == cudaSuccess) e = cudaMalloc(&m->faces, std::max<size_t>((size_t)nf * sizeof(b32_face), 16));
    if (e != cudaSuccess) { if (m->verts) cudaFree(m->verts); delete m; return cuda_fail(ctx, e, "cudaMalloc(mesh)"); }
    int rc = h2d(ctx, m->verts, vertices, (size_t)nv * sizeof(b32_vertex));
    if (rc == B32_OK) rc = h2d(ctx, m->faces, faces, (size_t)nf * sizeof(b32_face));
    cudaStreamSynchronize(ctx->stream);
    if (rc) { cudaFree(m->verts); cudaFree(m->faces); delete m; return rc; }
    *out = m;
    return B32_OK;
}

void b32_mesh_free(b32_ctx* ctx, b32_mesh* mesh) {
    if (!mesh) return;
    if (ctx) cudaSetDevice(ctx->device);
    if (ctx) cudaStreamSynchronize(ctx->stream);
    cudaFree(mesh->verts); cudaFree(mesh->faces);
    delete mesh;
}

int b32_render_mesh_15_resident(b32_ctx* ctx, const b32_mesh* mesh, const b32_camera* camera, const b32_settings* settings,
                                const b32_fog* fog, b32_timings* timings) {
    USE_DEVICE(ctx);
    if (!ctx || !mesh) return B32_ERR_INVALID;
    return render_device(ctx, mesh->verts, mesh->nv, mesh->faces, mesh->nf, camera, settings, fog, timings, true);
}

int b32_render_mesh_15_enqueue(b32_ctx* ctx, const b32_mesh* mesh, const b32_camera* camera, const b32_settings* settings, const b32_fog* fog) {
    USE_DEVICE(ctx);
    if (!ctx || !mesh || !settings) return B32_ERR_INVALID;
    // Neither pass needs a host round trip: a mesh / texture set that can produce pass-2 surfaces (or x-ray mode) has the
    // ordered replay enqueued behind pass 1.  Only the wireframe phase is rendered synchronously.
    bool wire = (settings->backface_cull && settings->backface_wireframe) || settings->wireframe_overlay;
    bool may_blend = mesh->has_nonopaque || settings->xray_mode;
    for (const TexDev& t : ctx->texdesc_h) may_blend = may_blend || t.blend != B32_BLEND_OPAQUE;
    return render_device(ctx, mesh->verts, mesh->nv, mesh->faces, mesh->nf, camera, settings, fog, nullptr, wire, false, nullptr, !may_blend);
}

int b32_render_mesh_placed(b32_ctx* ctx, const b32_mesh* mesh, const b32_placement* pl, const b32_camera* camera, const b32_settings* settings,
                           const b32_fog* fog, int rgb888, uint32_t flags, b32_timings* timings) {
    USE_DEVICE(ctx);
    if (!ctx || !mesh || !pl || !settings) return B32_ERR_INVALID;
    // scene.rs:123: has_transform
    bool has_transform = std::fabs(pl->facing) > 0.0001f || std::fabs(pl->world_pos[0]) > 0.0001f ||
                         std::fabs(pl->world_pos[1]) > 0.0001f || std::fabs(pl->world_pos[2]) > 0.0001f;
    const b32_vertex* verts = mesh->verts;
    if (has_transform && mesh->nv) {
        // the placed copy lives in the host-call staging buffer: stream order keeps it alive until this call's kernels ran
        CK(ctx->verts.reserve((size_t)mesh->nv + 1));
        launch_place(ctx->L(), mesh->verts, ctx->verts.p, mesh->nv, pl->cos_f, pl->sin_f, pl->world_pos);
        verts = ctx->verts.p;
    }
    bool wait = !(flags & B32_RENDER_ASYNC);
    bool all_opaque = false;
    if (!wait) {          // same rule as b32_render_mesh_15_enqueue
        wait = (settings->backface_cull && settings->backface_wireframe) || settings->wireframe_overlay;
        bool may_blend = rgb888 || mesh->has_nonopaque || settings->xray_mode;
        for (const TexDev& t : ctx->texdesc_h) may_blend = may_blend || t.blend != B32_BLEND_OPAQUE;
        all_opaque = !may_blend;
    }
    return render_device(ctx, verts, mesh->nv, mesh->faces, mesh->nf, camera, settings, rgb888 ? nullptr : fog,
                         (flags & B32_RENDER_ASYNC) ? nullptr : timings, wait, rgb888 != 0, nullptr, all_opaque);
}

int b32_render_skybox_mesh(b32_ctx* ctx, const b32_sky_vertex* vertices, uint32_t nv, const uint32_t* faces, uint32_t nf, const b32_camera* camera) {
    USE_DEVICE(ctx);
    if (!ctx || !camera) return B32_ERR_INVALID;
    if ((nv && !vertices) || (nf && !faces)) return fail(ctx, B32_ERR_INVALID, "vertices/faces is NULL");
    if (ctx->width == 0 || ctx->height == 0) return fail(ctx, B32_ERR_INVALID, "framebuffer has zero size (call b32_fb_resize)");
    if (nf == 0) return B32_OK;
    b32_settings s{};                     // project(): the float path, no ortho (render.rs:89, :107)
    CallParams p; std::vector<LightDev> lights;
    int rc = fill_params(ctx, p, camera, &s, nullptr, nv, nf, lights); if (rc) return rc;
    rc = ensure_work(ctx, p); if (rc) return rc;
    // the sky mesh reuses the mesh staging and record buffers (a SkyRec is smaller than a SurfRec)
    static_assert(sizeof(SkyRec) <= sizeof(SurfRec) && sizeof(b32_sky_vertex) <= sizeof(b32_vertex), "staging reuse");
    CK(ctx->verts.reserve((size_t)std::max<uint32_t>(nv, 1) + 1));      // + 1: k_setup's staged window may overrun by < 16 bytes
    CK(ctx->faces.reserve(std::max<uint32_t>(nf, 1)));      // 3 x u32 per face fits the 16-byte b32_face slots
    cudaStream_t st = ctx->stream;
    // The usual sky (a few thousand faces): indices are checked here, so nothing can fail on the device and the pass is
    // only enqueued — no wait, no status read-back.  Larger ones are checked on the device and waited for.
    const bool enqueue_only = nf <= 65536;
    if (enqueue_only) {
        for (uint32_t i = 0; i < nf * 3; ++i)
            if (faces[i] >= nv) return fail(ctx, B32_ERR_OOB_INDEX, "skybox face vertex index out of range (reference: slice index panic)");
    }
    rc = h2d(ctx, ctx->verts.p, vertices, (size_t)nv * sizeof(b32_sky_vertex), enqueue_only); if (rc) return rc;
    rc = h2d(ctx, ctx->faces.p, faces, (size_t)nf * 12, enqueue_only); if (rc) return rc;
    ctx->state_cur ^= 1;
    ctx->state = reinterpret_cast<CallState*>(ctx->state_ring.p + (size_t)ctx->state_cur * ctx->state_stride);
    ctx->tile_count = ctx->state_ring.p + (size_t)ctx->state_cur * ctx->state_stride + STATE_WORDS;
    uint32_t* zero_next = ctx->state_ring.p + (size_t)(ctx->state_cur ^ 1) * ctx->state_stride;
    launch_sky(ctx->L(), reinterpret_cast<const b32_sky_vertex*>(ctx->verts.p), reinterpret_cast<const uint32_t*>(ctx->faces.p),
               reinterpret_cast<SkyRec*>(ctx->recs.p), ctx->heads.p, ctx->masks.p, ctx->fb_rgba.p, ctx->state,
               zero_next, ctx->state_stride, p);
    CK(cudaGetLastError());
    if (enqueue_only) return B32_OK;
    CK(cudaMemcpyAsync(ctx->state_h, ctx->state, sizeof(CallState), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    if (ctx->state_h->oob) return fail(ctx, B32_ERR_OOB_INDEX, "skybox face vertex index out of range (reference: slice index panic)");
    return B32_OK;
}

int b32_render_stars(b32_ctx* ctx, const b32_star* stars, uint32_t n, const b32_camera* camera, float size) {
    USE_DEVICE(ctx);
    if (!ctx || !camera) return B32_ERR_INVALID;
    if (n && !stars) return fail(ctx, B32_ERR_INVALID, "stars is NULL");
    if (n == 0 || ctx->width == 0 || ctx->height == 0) return B32_OK;
    b32_settings s{};                     // project(): the float path (render.rs:181)
    b32_camera cam = *camera;
    cam.position[0] = cam.position[1] = cam.position[2] = 0.0f;      // directions, not positions (:178)
    CallParams p; std::vector<LightDev> lights;
    int rc = fill_params(ctx, p, &cam, &s, nullptr, n, 0, lights); if (rc) return rc;
    static_assert(sizeof(b32_star) * 2 == sizeof(b32_line), "stars are staged in the line list buffer");
    CK(ctx->lines.reserve((n + 1) / 2));
    CK(ctx->line_scratch.reserve((size_t)ctx->width * ctx->height));
    rc = h2d(ctx, ctx->lines.p, stars, (size_t)n * sizeof(b32_star), true); if (rc) return rc;    // enqueue-only: the list is consumed here
    float sz = size != size || size < 1.0f ? 1.0f : size;            // size.max(1.0) as i32 (:202)
    int32_t isz = sz >= 3.0f ? 3 : (int32_t)sz;
    launch_stars(ctx->L(), reinterpret_cast<const b32_star*>(ctx->lines.p), n, isz, ctx->line_scratch.p, ctx->fb_rgba.p, p);
    CK(cudaGetLastError());
    return B32_OK;
}

int b32_draw_lines(b32_ctx* ctx, const b32_line* lines, uint32_t n) {
    USE_DEVICE(ctx);
    if (!ctx) return B32_ERR_INVALID;
    if (n && !lines) return fail(ctx, B32_ERR_INVALID, "lines is NULL");
    if (n == 0 || ctx->width == 0 || ctx->height == 0) return B32_OK;
    bool any_blended = false;
    for (uint32_t i = 0; i < n; ++i) {
        const b32_line& l = lines[i];
        if (l.kind > B32_LINE_THICK || (l.kind == B32_LINE_2D && l.mode > B32_BLEND_ERASE) || (l.kind == B32_LINE_THICK && !(l.z0 == std::trunc(l.z0) && std::fabs(l.z0) < 16777216.0f)))
            return fail(ctx, B32_ERR_INVALID, "line " + std::to_string(i) + ": unknown kind or blend mode");
        const int32_t m = B32_LINE_MAX_COORD;
        const bool circle = l.kind == B32_LINE_CIRCLE || l.kind == B32_LINE_CIRCLE_ALPHA;
        if (l.x0 < -m || l.x0 > m || l.y0 < -m || l.y0 > m || (!circle && (l.x1 < -m || l.x1 > m || l.y1 < -m || l.y1 > m)) ||
            (circle && (l.x1 < -32767 || l.x1 > 32767)))            // radius: dx * dx + dy * dy must stay inside i32 as in the reference
            return fail(ctx, B32_ERR_UNSUPPORTED, "line " + std::to_string(i) + ": coordinate beyond B32_LINE_MAX_COORD (radius beyond 32767)");
        bool overwrites = l.kind == B32_LINE_2D ? (l.mode == B32_BLEND_OPAQUE || l.mode == B32_BLEND_ERASE)
                                                : (l.kind != B32_LINE_2D_ALPHA && l.kind != B32_LINE_3D_ALPHA && l.kind != B32_LINE_CIRCLE_ALPHA);
        any_blended |= !overwrites;
    }
    constexpr uint32_t N_FLAGS = 32;
    const size_t px = (size_t)ctx->width * ctx->height;
    CK(ctx->lines.reserve(n));
    CK(ctx->line_scratch.reserve(3 * px + n + N_FLAGS));
    uint32_t* owner = ctx->line_scratch.p;
    uint32_t* next[2] = {owner + px, owner + 2 * px};
    uint32_t* wait = owner + 3 * px;
    uint32_t* flags = wait + n;
    int rc = h2d(ctx, ctx->lines.p, lines, (size_t)n * sizeof(b32_line), true); if (rc) return rc;   // the list is consumed here
    cudaStream_t st = ctx->stream;
    LaunchCtx L = ctx->L();
    launch_lines_begin(L, ctx->lines.p, n, owner, next[0], wait, flags, N_FLAGS, ctx->fb_rgba.p, ctx->fb_z.p, ctx->width, ctx->height, any_blended);
    if (any_blended) {
        // one blended operation per pixel per round, in list order; batches of rounds run until a batch ends with
        // nothing waiting (rounds after the last useful one return at once)
        uint32_t plane = 0;
        auto round = [&](uint32_t r) {
            launch_lines_round(L, ctx->lines.p, n, owner, next[plane], next[plane ^ 1], wait, flags, r, ctx->fb_rgba.p, ctx->fb_z.p, ctx->width, ctx->height);
            plane ^= 1;
        };
        round(0);
        for (uint32_t batch = 3;; batch = std::min(2 * batch + 1, N_FLAGS - 1)) {
            for (uint32_t r = 1; r <= batch; ++r) round(r);
            CK(cudaMemcpyAsync(ctx->state_h, flags + batch, 4, cudaMemcpyDeviceToHost, st));    // pinned scratch word
            CK(cudaStreamSynchronize(st));
            if (*reinterpret_cast<uint32_t*>(ctx->state_h) == 0) break;
            CK(cudaMemsetAsync(flags, 0, N_FLAGS * 4, st));
            CK(cudaMemsetAsync(flags, 1, 4, st));                   // the next batch's round 1 sees "waiting"
        }
    }
    CK(cudaGetLastError());
    return B32_OK;
}

void* b32_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
void b32_host_free(void* p) { if (p) cudaFreeHost(p); }

int b32_debug_transform(b32_ctx* ctx, const b32_vertex* vertices, uint32_t nv, const b32_camera* camera, const b32_settings* settings,
                        float* out_screen, float* out_cam) {
    USE_DEVICE(ctx);
    if (!ctx || !camera || !settings || (nv && (!vertices || !out_screen || !out_cam))) return B32_ERR_INVALID;
    if (nv == 0) return B32_OK;
    CallParams p; std::vector<LightDev> lights;
    int rc = fill_params(ctx, p, camera, settings, nullptr, nv, 0, lights); if (rc) return rc;
    CK(ctx->verts.reserve(nv)); CK(ctx->tv.reserve(nv)); CK(ctx->dbg.reserve((size_t)nv * 3));
    rc = h2d(ctx, ctx->verts.p, vertices, (size_t)nv * sizeof(b32_vertex)); if (rc) return rc;
    launch_transform(ctx->L(), ctx->verts.p, ctx->tv.p, ctx->dbg.p, p);
    std::vector<float4> tv(nv);
    CK(cudaMemcpyAsync(tv.data(), ctx->tv.p, (size_t)nv * sizeof(float4), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(out_cam, ctx->dbg.p, (size_t)nv * 12, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    for (uint32_t i = 0; i < nv; ++i) { out_screen[i * 3] = tv[i].x; out_screen[i * 3 + 1] = tv[i].y; out_screen[i * 3 + 2] = tv[i].z; }
    return B32_OK;
}

// Measurement aid: with n > 0 the next enqueued frames of this context are launched plainly (no CUDA graph) with events in
// front of k_setup, in front of the fill kernel(s) and behind them, n frames deep; n = 0 switches it off.
int b32_debug_timing_ring(b32_ctx* ctx, uint32_t n) {
    USE_DEVICE(ctx);
    if (!ctx) return B32_ERR_INVALID;
    CK(cudaStreamSynchronize(ctx->stream));
    for (cudaEvent_t e : ctx->tring) cudaEventDestroy(e);
    ctx->tring.clear(); ctx->tring_n = 0; ctx->tring_count = 0;
    for (uint32_t i = 0; i < 3 * n; ++i) { cudaEvent_t e; CK(cudaEventCreate(&e)); ctx->tring.push_back(e); }
    ctx->tring_n = n;
    return B32_OK;
}
int b32_debug_prefix_hint(b32_ctx* ctx) { return ctx ? (ctx->next_prefix_hint() ? 1 : 0) : -1; }
// Device times of the timed frames (after a sync): setup_ms[i], fill_ms[i] for the last min(frames, n) frames. Returns their number.
int b32_debug_timing_read(b32_ctx* ctx, float* setup_ms, float* fill_ms, uint32_t cap) {
    if (!ctx || !ctx->tring_n) return 0;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    uint32_t n = std::min<uint32_t>(std::min<uint32_t>(ctx->tring_count, ctx->tring_n), cap);
    for (uint32_t i = 0; i < n; ++i) {
        cudaEvent_t* e = &ctx->tring[(size_t)i * 3];
        if (cudaEventElapsedTime(&setup_ms[i], e[0], e[1]) != cudaSuccess || cudaEventElapsedTime(&fill_ms[i], e[1], e[2]) != cudaSuccess) { cudaGetLastError(); return (int)i; }
    }
    return (int)n;
}

int b32_debug_kernel_times(b32_ctx* ctx, float* out_ms, uint32_t cap) {
    if (!ctx || !out_ms) return 0;
    uint32_t n = std::min<uint32_t>(cap, 7);
    for (uint32_t i = 0; i < n; ++i) out_ms[i] = ctx->kernel_ms[i];
    return (int)n;
}

int b32_debug_draw_order(b32_ctx* ctx, uint32_t* out_face_idx, uint32_t cap, uint32_t* n) {
    USE_DEVICE(ctx);
    // Test hook: the draw order of the last call = stable sort of (pass, depth key) over the drawn faces
    // (render.rs:2522-2542).  Pass 1 is order-free on the device and pass 2 sorts per tile, so the global
    // order only exists here, computed on the host from the device's keys.
    if (!ctx || !n) return B32_ERR_INVALID;
    CK(cudaStreamSynchronize(ctx->stream));
    std::vector<uint64_t> keys(ctx->last_nf);
    if (ctx->last_nf) CK(cudaMemcpy(keys.data(), ctx->keys.p, (size_t)ctx->last_nf * 8, cudaMemcpyDeviceToHost));
    std::vector<uint32_t> idx;
    for (uint32_t i = 0; i < ctx->last_nf; ++i) if ((keys[i] >> 32) < 2) idx.push_back(i);
    std::stable_sort(idx.begin(), idx.end(), [&](uint32_t a, uint32_t b) { return keys[a] < keys[b]; });
    *n = (uint32_t)idx.size();
    for (uint32_t i = 0; i < idx.size() && i < cap && out_face_idx; ++i) out_face_idx[i] = idx[i];
    return B32_OK;
}

}  // extern "C"
