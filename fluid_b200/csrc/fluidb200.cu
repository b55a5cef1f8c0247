// fluidb200.cu -- host side of libfluidb200.so: handle lifecycle, phase driver and
// the C ABI declared in include/fluidb200.h.  The phase order and the buffer
// semantics follow (*Fluid).Simulate, pkg/fluid/fluid.go:79-109 of the reference.
#include "kernels.cuh"
#include "advect_fused.cuh"
#include "rb_fused.cuh"
#include "rbq_fused.cuh"
#include "rbq_stream.cuh"
#include "advect_tile.cuh"
#include "multigrid.cuh"

#include <cudaTypedefs.h>      // PFN_cuTensorMapEncodeTiled (resolved through the runtime: no -lcuda)
#include <cmath>
#include <cstdio>
#include <map>
#include <cstdlib>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

#define FB_VERSION 100

enum { SCR_CURL = 0, SCR_A, SCR_B, SCR_C, SCR_D, SCR_E, SCR_F, SCR_BKU, SCR_BKV, SCR_NOISEU, SCR_NOISEV, SCR_VIEW, SCR_SNAP, SCR_N };

struct fb_handle {
    fb_config cfg;
    Grid g;
    int device;
    cudaStream_t stream;
    size_t plane_floats;          // floats per device plane
    float *f[FB_NFIELDS];         // plane currently playing each role (roles swap, planes do not move)
    float *scr[SCR_N];
    std::vector<float *> pool;    // free scratch planes of the fast path
    unsigned char *mask;          // neighbour mask of S (advect_fused.cuh), rebuilt when S changes
    bool mask_dirty;
    bool literal;                 // FB_FLAG_LITERAL: reference-shaped kernels with physical copies
    bool exact_shadow;            // FB_FLAG_EXACT_SHADOW: keep newU/newV/newM complete (white-box mode)
    bool p_zero;                  // pressure plane known to be all zero
    bool rb_attr_set, rbq_attr_set, adv_tile_attr_set;
    unsigned char *tile_flags;                          // per AT_TI x AT_TJ tile: all cells active (advect_tile.cuh)
    int tile_ntx, tile_nty;
    std::map<std::pair<const void *, int>, CUtensorMap> tmaps;   // (plane, box lines) -> 2-D tensor map of the tile loads
    bool want_stats;
    bool fuse_turb;               // apply addTurbulence in the fused solve's write-out (else its own pass)
    int nsm;
    bool noise_ready;
    float *mirror[FB_NFIELDS];    // pinned host mirrors (lazy)
    unsigned *d_red;              // 64 reduction slots
    unsigned *h_red;              // pinned copy
    int *d_bad;                   // halo-violation flag
    // exact-solver scheduling state
    int2 *d_order; int ntiles, nTa, nTb, order_T;
    int *d_tile_counter; int *d_done; int epoch;
    // multigrid V-cycle: coarse grid (fluid.go:632-633) and its three arrays, allocated on first use
    CoarseGrid cg; float *mg_rhs, *mg_cS, *mg_cP;
    fb_edit_cmd *d_cmds; size_t d_cmds_cap;
    std::vector<fb_edit_cmd> staged_cmds;   // host copy of what d_cmds holds (per-step lists repeat)
    cudaEvent_t ev0, ev1;
    cudaStream_t copy_stream;     // device -> host leg of asynchronous views
    cudaEvent_t ev_snap, ev_view; // snapshot taken / view landed in host memory
    bool view_in_flight;
    // peer-memory halo exchange (fb_halo_export / connect / exchange)
    float *halo_send;             // [2 buffers][2 sides][3 fields][lines][pitch] + flags, IPC-exported
    size_t halo_buf_floats;       // floats in one of the two buffers
    int halo_lines;
    unsigned halo_epoch;
    const float *halo_peer[2];    // neighbours' send buffers (side 0 = lower i)
    void *halo_peer_ipc[2];       // what cudaIpcOpenMemHandle returned (nullptr for same-process peers)
    fb_particle_dev *d_particles; int *d_alive; size_t particles_cap;   // fb_advect_particles scratch
    std::vector<int> h_alive;
    fb_handle *halo_peer_local[2];// same-process neighbours: their post is awaited with an event, not by spinning
    cudaEvent_t ev_halo;          // recorded after every post
    // overlapped exchange inside fb_step_local (FB_OPT_HALO_OVERLAP): second stream, its events, its own epochs
    bool halo_overlap;
    cudaStream_t halo_stream;
    cudaEvent_t ev_ovl_uv, ev_ovl_m, ev_ovl_done;
    unsigned ovl_epoch_uv, ovl_epoch_m;
    bool prof;
    std::vector<cudaEvent_t> prof_pool;                 // recycled events
    std::vector<std::pair<int, std::pair<cudaEvent_t, cudaEvent_t>>> prof_pairs;
    uint64_t launches;
    fb_solve_stats stats;
    std::string err;
};

// ---- error plumbing -----------------------------------------------------------
static int fail(fb_handle *h, int code, const char *what, cudaError_t e = cudaSuccess)
{
    if (h) {
        h->err = what;
        if (e != cudaSuccess) { h->err += ": "; h->err += cudaGetErrorString(e); }
    }
    return code;
}
#define CK(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) return fail(h, FB_ERR_CUDA, #call, _e); } while (0)
#define CKL(what) do { h->launches++; cudaError_t _e = cudaGetLastError(); if (_e != cudaSuccess) return fail(h, FB_ERR_CUDA, what, _e); } while (0)
#define TRY(expr) do { int _s = (expr); if (_s != FB_OK) return _s; } while (0)

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// ---- per-phase event timing (fb_profile_*) ----------------------------------------
// the large kernels are timed one launch at a time as well (slots FB_PROF_K_*), inside their phase's pair
#define PROF_SLOT_k_advect_velocity_full FB_PROF_K_ADVECT_VELOCITY
#define PROF_SLOT_k_bfecc_velocity_correct FB_PROF_K_BFECC_VELOCITY
#define PROF_SLOT_k_advect_smoke_full FB_PROF_K_ADVECT_SMOKE
#define PROF_SLOT_k_bfecc_smoke_correct FB_PROF_K_BFECC_SMOKE
struct ProfScope {
    fb_handle *h; int phase; cudaEvent_t a, b; bool on;
    static cudaEvent_t get(fb_handle *h) {
        cudaEvent_t e = nullptr;
        if (!h->prof_pool.empty()) { e = h->prof_pool.back(); h->prof_pool.pop_back(); }
        else if (cudaEventCreate(&e) != cudaSuccess) e = nullptr;
        return e;
    }
    ProfScope(fb_handle *h_, int phase_) : h(h_), phase(phase_), a(nullptr), b(nullptr), on(h_->prof) {
        if (!on) return;
        a = get(h); b = get(h);
        if (!a || !b) { on = false; return; }
        cudaEventRecord(a, h->stream);
    }
    ~ProfScope() {
        if (!on) return;
        cudaEventRecord(b, h->stream);
        h->prof_pairs.push_back({phase, {a, b}});
    }
};

static int scratch(fb_handle *h, int which, float **out)
{
    if (!h->scr[which]) {
        CK(cudaMalloc(&h->scr[which], h->plane_floats * sizeof(float)));
        CK(cudaMemsetAsync(h->scr[which], 0, h->plane_floats * sizeof(float), h->stream));
    }
    *out = h->scr[which];
    return FB_OK;
}

// scratch planes of the fast path: taken for an output, the plane they replace is given back
static int take_plane(fb_handle *h, float **out)
{
    if (!h->pool.empty()) { *out = h->pool.back(); h->pool.pop_back(); return FB_OK; }
    CK(cudaMalloc(out, h->plane_floats * sizeof(float)));
    CK(cudaMemsetAsync(*out, 0, h->plane_floats * sizeof(float), h->stream));
    return FB_OK;
}
static inline void give_plane(fb_handle *h, float *p) { h->pool.push_back(p); }

// rows x columns launch geometry for full-plane kernels: x along j (unit stride)
static inline void plane_launch(const Grid &g, int ib, int ie, dim3 &grid, dim3 &block, int cols = -1)
{
    block = dim3(128, 2, 1);
    if (cols < 0) cols = g.NY;
    grid = dim3(cdiv(cols, block.x), cdiv(ie - ib, block.y), 1);
}

// ---- lifecycle ------------------------------------------------------------------
extern "C" int fb_version(void) { return FB_VERSION; }

extern "C" int fb_default_params(fb_params *p)
{
    if (!p) return FB_ERR_INVALID;
    p->relaxation = 1.9f;             // fluid.go:8
    p->confinement = 0.0f;            // fluid.go:59
    p->viscosity_diffusion = 0.0f;    // fluid.go:60
    p->pressure_damping = 1.0f;       // fluid.go:61
    p->turbulence_strength = 0.02f;   // fluid.go:62
    p->smoke_advection = 1.0f;        // fluid.go:63
    p->use_multigrid = 0;             // fluid.go:64
    p->multigrid_levels = 2;          // fluid.go:65
    p->use_bfecc = 0;                 // fluid.go:66
    p->solver = FB_SOLVER_EXACT;
    p->iters = 8;                     // fluid.go:81
    return FB_OK;
}

extern "C" int fb_create(const fb_config *cfg, fb_handle **out)
{
    if (!cfg || !out) return FB_ERR_INVALID;
    *out = nullptr;
    if (cfg->width < 1 || cfg->height < 1) return FB_ERR_INVALID;
    const int nranks = cfg->nranks < 1 ? 1 : cfg->nranks;
    if (cfg->rank < 0 || cfg->rank >= nranks) return FB_ERR_INVALID;
    fb_handle *h = new (std::nothrow) fb_handle();
    if (!h) return FB_ERR_NOMEM;
    h->cfg = *cfg;
    h->cfg.nranks = nranks;
    h->device = cfg->device;
    cudaError_t e = cudaSetDevice(h->device);
    if (e != cudaSuccess) { fprintf(stderr, "fluidb200: cudaSetDevice(%d): %s\n", h->device, cudaGetErrorString(e)); delete h; return FB_ERR_CUDA; }

    Grid &g = h->g;
    g.NX = cfg->width + 2;
    g.NY = cfg->height + 2;
    g.pitch = cdiv(g.NY, 32) * 32;
    // slab decomposition of the interior lines 1..NX-2 over ranks; the ring lines
    // i = 0 and i = NX-1 go to the first / last rank.
    {
        const long long W = cfg->width;
        const long long lo = 1 + W * cfg->rank / nranks, hi = 1 + W * (cfg->rank + 1) / nranks;
        g.i_lo = (int)lo; g.i_hi = (int)hi;
        if (cfg->rank == 0) g.i_lo = 0;
        if (cfg->rank == nranks - 1) g.i_hi = g.NX;
    }
    h->literal = (cfg->flags & FB_FLAG_LITERAL) != 0 || getenv("FLUIDB200_LITERAL") != nullptr;
    h->exact_shadow = (cfg->flags & FB_FLAG_EXACT_SHADOW) != 0;
    h->mask_dirty = true;
    h->want_stats = true;
    // measured on B200: the solve's 4 writer warps become its critical path when they also do the
    // turbulence (sqrt + 2 noise loads): 0.50 ms fused vs 0.35 + 0.08 ms as a separate pass
    h->fuse_turb = getenv("FLUIDB200_FUSE_TURB") != nullptr;
    h->nsm = 148;
    cudaDeviceGetAttribute(&h->nsm, cudaDevAttrMultiProcessorCount, h->device);
    int ghost = nranks > 1 ? (cfg->ghost > 0 ? cfg->ghost : 32) : 0;
    h->cfg.ghost = ghost;
    g.i_alloc0 = g.i_lo - ghost < 0 ? 0 : g.i_lo - ghost;
    const int alloc_end = g.i_hi + ghost > g.NX ? g.NX : g.i_hi + ghost;
    g.lines_alloc = alloc_end - g.i_alloc0;
    h->plane_floats = (size_t)g.lines_alloc * (size_t)g.pitch;

#define CKC(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) { fprintf(stderr, "fluidb200: %s: %s\n", #call, cudaGetErrorString(_e)); fb_destroy(h); return _e == cudaErrorMemoryAllocation ? FB_ERR_NOMEM : FB_ERR_CUDA; } } while (0)
    CKC(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    for (int k = 0; k < FB_NFIELDS; k++) {
        CKC(cudaMalloc(&h->f[k], h->plane_floats * sizeof(float)));
        CKC(cudaMemsetAsync(h->f[k], 0, h->plane_floats * sizeof(float), h->stream));   // New(): all zero => all solid
    }
    CKC(cudaMalloc(&h->mask, h->plane_floats));
    CKC(cudaMemsetAsync(h->mask, 0, h->plane_floats, h->stream));
    CKC(cudaMalloc(&h->d_red, 64 * sizeof(unsigned)));
    CKC(cudaMemsetAsync(h->d_red, 0, 64 * sizeof(unsigned), h->stream));
    CKC(cudaMallocHost(&h->h_red, 64 * sizeof(unsigned)));
    CKC(cudaMalloc(&h->d_bad, sizeof(int)));
    CKC(cudaMemsetAsync(h->d_bad, 0, sizeof(int), h->stream));
    CKC(cudaMalloc(&h->d_tile_counter, sizeof(int)));
    CKC(cudaEventCreate(&h->ev0));
    CKC(cudaEventCreate(&h->ev1));
    CKC(cudaStreamSynchronize(h->stream));
#undef CKC
    *out = h;
    return FB_OK;
}

extern "C" int fb_destroy(fb_handle *h)
{
    if (!h) return FB_OK;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    for (int k = 0; k < FB_NFIELDS; k++) { if (h->f[k]) cudaFree(h->f[k]); if (h->mirror[k]) cudaFreeHost(h->mirror[k]); }
    for (int k = 0; k < SCR_N; k++) if (h->scr[k]) cudaFree(h->scr[k]);
    for (float *p : h->pool) cudaFree(p);
    if (h->mask) cudaFree(h->mask);
    if (h->tile_flags) cudaFree(h->tile_flags);
    if (h->d_red) cudaFree(h->d_red);
    if (h->h_red) cudaFreeHost(h->h_red);
    if (h->d_bad) cudaFree(h->d_bad);
    if (h->d_order) cudaFree(h->d_order);
    if (h->d_tile_counter) cudaFree(h->d_tile_counter);
    if (h->d_done) cudaFree(h->d_done);
    if (h->d_cmds) cudaFree(h->d_cmds);
    if (h->mg_rhs) cudaFree(h->mg_rhs);
    if (h->mg_cS) cudaFree(h->mg_cS);
    if (h->mg_cP) cudaFree(h->mg_cP);
    for (auto &pr : h->prof_pairs) { cudaEventDestroy(pr.second.first); cudaEventDestroy(pr.second.second); }
    for (auto e : h->prof_pool) cudaEventDestroy(e);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->copy_stream) { cudaStreamSynchronize(h->copy_stream); cudaStreamDestroy(h->copy_stream); }
    for (int sd = 0; sd < 2; sd++) if (h->halo_peer_ipc[sd]) cudaIpcCloseMemHandle(h->halo_peer_ipc[sd]);
    if (h->halo_send) cudaFree(h->halo_send);
    if (h->d_particles) cudaFree(h->d_particles);
    if (h->d_alive) cudaFree(h->d_alive);
    if (h->ev_halo) cudaEventDestroy(h->ev_halo);
    if (h->halo_stream) {
        cudaStreamSynchronize(h->halo_stream); cudaStreamDestroy(h->halo_stream);
        cudaEventDestroy(h->ev_ovl_uv); cudaEventDestroy(h->ev_ovl_m); cudaEventDestroy(h->ev_ovl_done);
    }
    if (h->ev_snap) cudaEventDestroy(h->ev_snap);
    if (h->ev_view) cudaEventDestroy(h->ev_view);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return FB_OK;
}

extern "C" const char *fb_last_error(const fb_handle *h) { return h ? h->err.c_str() : "null handle"; }

extern "C" int fb_dims(const fb_handle *h, int64_t *nx, int64_t *ny, int64_t *i_lo, int64_t *i_hi)
{
    if (!h) return FB_ERR_INVALID;
    if (nx) *nx = h->g.NX;
    if (ny) *ny = h->g.NY;
    if (i_lo) *i_lo = h->g.i_lo;
    if (i_hi) *i_hi = h->g.i_hi;
    return FB_OK;
}

// ---- halo exchange through peer memory -----------------------------------------------------
// Four send buffers + 64 flag words: buffers 0 / 1 alternate for fb_halo_exchange (flag word 0 of the buffer), buffers 2 / 3
// for the exchange overlapped with fb_step_local (word 1 = U, V posted, word 2 = M posted): either protocol's two-buffer
// argument (kernels.cuh) holds on its own pair however the two are interleaved.
#define HALO_NBUF 4
static inline size_t halo_total_bytes(const fb_handle *h) { return HALO_NBUF * h->halo_buf_floats * sizeof(float) + 256; }
static inline unsigned *halo_flags(const fb_handle *h, const float *base, int buf)
{
    return reinterpret_cast<unsigned *>(const_cast<float *>(base) + HALO_NBUF * h->halo_buf_floats) + 16 * buf;
}

extern "C" int fb_halo_export(fb_handle *h, int32_t lines, void *ipc_handle64, uint64_t *device_ptr, size_t *bytes)
{
    if (!h) return FB_ERR_INVALID;
    CK(cudaSetDevice(h->device));
    const Grid &g = h->g;
    if (lines < 1 || lines > h->cfg.ghost) return fail(h, FB_ERR_INVALID, "halo wider than the ghost zone");
    if (lines > g.i_hi - g.i_lo) return fail(h, FB_ERR_INVALID, "halo wider than the slab");
    if (h->halo_send && h->halo_lines != lines) return fail(h, FB_ERR_INVALID, "halo width cannot change after export");
    if (!h->halo_send) {
        h->halo_lines = lines;
        h->halo_buf_floats = (size_t)6 * lines * g.pitch;
        CK(cudaMalloc(&h->halo_send, halo_total_bytes(h)));
        CK(cudaEventCreateWithFlags(&h->ev_halo, cudaEventDisableTiming));
        CK(cudaMemsetAsync(h->halo_send, 0, halo_total_bytes(h), h->stream));
        CK(cudaStreamSynchronize(h->stream));
    }
    if (ipc_handle64) {
        cudaIpcMemHandle_t mh;
        static_assert(sizeof(mh) == 64, "cudaIpcMemHandle_t is 64 bytes");
        CK(cudaIpcGetMemHandle(&mh, h->halo_send));
        memcpy(ipc_handle64, &mh, 64);
    }
    if (device_ptr) *device_ptr = (uint64_t)(uintptr_t)h->halo_send;
    if (bytes) *bytes = halo_total_bytes(h);
    return FB_OK;
}

extern "C" int fb_halo_connect(fb_handle *h, int32_t side, const void *ipc_handle64, uint64_t device_ptr)
{
    if (!h || side < 0 || side > 1) return FB_ERR_INVALID;
    CK(cudaSetDevice(h->device));
    if (!h->halo_send) return fail(h, FB_ERR_INVALID, "fb_halo_connect before fb_halo_export");
    const Grid &g = h->g;
    const int recv_i = side == 0 ? g.i_lo - h->halo_lines : g.i_hi;
    if (recv_i < g.i_alloc0 || recv_i + h->halo_lines > g.i_alloc0 + g.lines_alloc)
        return fail(h, FB_ERR_INVALID, "no neighbour on that side");
    if (h->halo_peer_ipc[side]) { cudaIpcCloseMemHandle(h->halo_peer_ipc[side]); h->halo_peer_ipc[side] = nullptr; }
    if (ipc_handle64) {
        cudaIpcMemHandle_t mh;
        memcpy(&mh, ipc_handle64, 64);
        void *p = nullptr;
        CK(cudaIpcOpenMemHandle(&p, mh, cudaIpcMemLazyEnablePeerAccess));
        h->halo_peer_ipc[side] = p;
        h->halo_peer[side] = static_cast<const float *>(p);
    } else {
        if (!device_ptr) return fail(h, FB_ERR_INVALID, "fb_halo_connect: neither an IPC handle nor a device pointer");
        h->halo_peer[side] = reinterpret_cast<const float *>((uintptr_t)device_ptr);   // same process
    }
    return FB_OK;
}

// Same-process neighbour (several handles in one process, possibly on one device): attach by handle.
// Its post is then awaited with a stream-ordered event before the pull is launched, so that the
// pull kernel never spins on a device that still has to run the post it waits for.
extern "C" int fb_halo_connect_local(fb_handle *h, int32_t side, fb_handle *peer)
{
    if (!h || !peer || side < 0 || side > 1) return FB_ERR_INVALID;
    if (!peer->halo_send) return fail(h, FB_ERR_INVALID, "fb_halo_connect_local: the peer has not exported");
    TRY(fb_halo_connect(h, side, nullptr, (uint64_t)(uintptr_t)peer->halo_send));
    h->halo_peer_local[side] = peer;
    return FB_OK;
}

// pack fields [field0, field0 + nf) of the boundary lines into send buffer `buf` and publish `epoch` in flag word `word`
static int halo_post_fields(fb_handle *h, cudaStream_t st, int buf, int field0, int nf, int word, unsigned epoch)
{
    const Grid &g = h->g;
    HaloPack a;
    a.src[0] = h->f[FB_U]; a.src[1] = h->f[FB_V]; a.src[2] = h->f[FB_M];
    a.dst = h->halo_send + (size_t)buf * h->halo_buf_floats;
    a.i_lo = g.i_lo; a.i_hi = g.i_hi; a.lines = h->halo_lines; a.pitch = g.pitch; a.i_alloc0 = g.i_alloc0;
    a.field0 = field0; a.nf = nf;
    k_halo_pack<<<dim3(cdiv(g.pitch / 4, 256), h->halo_lines, 2 * nf), 256, 0, st>>>(a);
    CKL("k_halo_pack");
    k_halo_publish<<<1, 1, 0, st>>>(halo_flags(h, h->halo_send, buf) + word, epoch);
    CKL("k_halo_publish");
    return FB_OK;
}
// pull the same fields out of the connected neighbours' buffer `buf` once their flag word reached `epoch`
static int halo_pull_fields(fb_handle *h, cudaStream_t st, int buf, int field0, int nf, int word, unsigned epoch)
{
    const Grid &g = h->g;
    HaloPull a;
    a.dst[0] = h->f[FB_U]; a.dst[1] = h->f[FB_V]; a.dst[2] = h->f[FB_M];
    for (int sd = 0; sd < 2; sd++) {
        a.peer[sd] = h->halo_peer[sd] ? h->halo_peer[sd] + (size_t)buf * h->halo_buf_floats : nullptr;
        a.peer_flag[sd] = h->halo_peer[sd] ? halo_flags(h, h->halo_peer[sd], buf) + word : nullptr;
    }
    a.recv_i[0] = g.i_lo - h->halo_lines; a.recv_i[1] = g.i_hi;
    a.lines = h->halo_lines; a.pitch = g.pitch; a.i_alloc0 = g.i_alloc0; a.epoch = epoch; a.bad = h->d_bad;
    a.field0 = field0; a.nf = nf;
    k_halo_pull<<<dim3(cdiv(g.pitch / 4, 256), h->halo_lines, 2 * nf), 256, 0, st>>>(a);
    CKL("k_halo_pull");
    return FB_OK;
}

// phase 1 of an exchange: pack the boundary lines and publish the epoch (never waits)
extern "C" int fb_halo_post(fb_handle *h)
{
    if (!h) return FB_ERR_INVALID;
    CK(cudaSetDevice(h->device));
    if (!h->halo_send) return fail(h, FB_ERR_INVALID, "fb_halo_post before fb_halo_export");
    h->halo_epoch++;
    TRY(halo_post_fields(h, h->stream, (int)(h->halo_epoch & 1u), 0, 3, 0, h->halo_epoch));
    CK(cudaEventRecord(h->ev_halo, h->stream));
    return FB_OK;
}

// phase 2: pull the ghost lines out of the connected neighbours' send buffers (waits on their flags)
extern "C" int fb_halo_pull(fb_handle *h)
{
    if (!h) return FB_ERR_INVALID;
    CK(cudaSetDevice(h->device));
    if (!h->halo_send) return fail(h, FB_ERR_INVALID, "fb_halo_pull before fb_halo_export");
    if (!h->halo_peer[0] && !h->halo_peer[1]) return FB_OK;
    for (int sd = 0; sd < 2; sd++)
        if (h->halo_peer_local[sd]) {
            if (h->halo_peer_local[sd]->halo_epoch != h->halo_epoch)
                return fail(h, FB_ERR_INVALID, "fb_halo_pull: a same-process neighbour has not posted this exchange yet");
            CK(cudaStreamWaitEvent(h->stream, h->halo_peer_local[sd]->ev_halo, 0));
        }
    return halo_pull_fields(h, h->stream, (int)(h->halo_epoch & 1u), 0, 3, 0, h->halo_epoch);
}

// The exchange for the NEXT step, overlapped with the rest of this one (FB_OPT_HALO_OVERLAP; SURVEY.md 8e "boundary tiles
// first").  U and V are final once the velocity advection is done: they are packed, published and the neighbours' lines pulled on
// a second stream while the smoke passes run.  The last smoke pass is launched for the boundary strips first, then M takes the
// same route while the interior lines are computed.  The pulls write ghost lines only.  Those of M are not read by anything still
// running.  Those of U, V are read by the smoke passes inside the zone this rank recomputed itself (the reach), where the pulled
// values are bit-identical to the local ones -- that identity is what makes N slabs equal one GPU -- and not outside it.
static bool halo_overlap_on(const fb_handle *h)
{
    return h->halo_overlap && h->cfg.nranks > 1 && h->halo_send && (h->halo_peer[0] || h->halo_peer[1]) &&
           !h->halo_peer_local[0] && !h->halo_peer_local[1];
}
static int halo_overlap_setup(fb_handle *h)
{
    if (h->halo_stream) return FB_OK;
    // highest priority: its few small blocks go in front of the smoke pass's as soon as an SM has room
    int prio_lo = 0, prio_hi = 0;
    CK(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    CK(cudaStreamCreateWithPriority(&h->halo_stream, cudaStreamNonBlocking, prio_hi));
    CK(cudaEventCreateWithFlags(&h->ev_ovl_uv, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&h->ev_ovl_m, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&h->ev_ovl_done, cudaEventDisableTiming));
    return FB_OK;
}
// fields [field0, field0 + nf) are final in the second stream's own order: exchange them there
static int halo_exchange_on_second_stream(fb_handle *h, int field0, int nf, int word, unsigned epoch)
{
    const int buf = 2 + (int)(epoch & 1u);
    TRY(halo_post_fields(h, h->halo_stream, buf, field0, nf, word, epoch));
    return halo_pull_fields(h, h->halo_stream, buf, field0, nf, word, epoch);
}
// fields [field0, field0 + nf) are final on h->stream: exchange them on the second stream
static int halo_overlap_fields(fb_handle *h, cudaEvent_t ev, int field0, int nf, int word, unsigned epoch)
{
    CK(cudaEventRecord(ev, h->stream));
    CK(cudaStreamWaitEvent(h->halo_stream, ev, 0));
    const int buf = 2 + (int)(epoch & 1u);
    TRY(halo_post_fields(h, h->halo_stream, buf, field0, nf, word, epoch));
    return halo_pull_fields(h, h->halo_stream, buf, field0, nf, word, epoch);
}

extern "C" int fb_halo_exchange(fb_handle *h)
{
    if (!h) return FB_ERR_INVALID;
    ProfScope ps(h, FB_PROF_HALO);
    TRY(fb_halo_post(h));
    return fb_halo_pull(h);
}

extern "C" int fb_ghost_lines(const fb_handle *h, int32_t *ghost)
{
    if (!h || !ghost) return FB_ERR_INVALID;
    *ghost = h->cfg.ghost;
    return FB_OK;
}

extern "C" int fb_stream(fb_handle *h, void **s) { if (!h || !s) return FB_ERR_INVALID; *s = (void *)h->stream; return FB_OK; }

// The fused solver's warp pipeline gives up (instead of hanging the GPU) when a hand-off
// never arrives; the flag it leaves behind becomes an error here.
static int check_pipeline(fb_handle *h)
{
    int dbg[6] = {0, 0, 0, 0, 0, 0};
    CK(cudaMemcpyAsync(dbg, h->d_red + 48, sizeof(dbg), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (dbg[0]) {
        char msg[200];
        snprintf(msg, sizeof(msg), "fused solver pipeline timed out: waited-for role %d line %d, thread %d, block (%d,%d), parity %d",
                 dbg[1] >> 20, dbg[1] & 0xfffff, dbg[2], dbg[3], dbg[4], dbg[5]);
        CK(cudaMemsetAsync(h->d_red + 48, 0, sizeof(dbg), h->stream));
        return fail(h, FB_ERR_CUDA, msg);
    }
    return FB_OK;
}

extern "C" int fb_synchronize(fb_handle *h)
{
    if (!h) return FB_ERR_INVALID;
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->stream));
    return check_pipeline(h);
}

extern "C" int fb_timer_start(fb_handle *h)
{
    if (!h) return FB_ERR_INVALID;
    CK(cudaSetDevice(h->device));
    CK(cudaEventRecord(h->ev0, h->stream));
    return FB_OK;
}

extern "C" int fb_timer_stop(fb_handle *h, float *ms)
{
    if (!h || !ms) return FB_ERR_INVALID;
    CK(cudaSetDevice(h->device));
    CK(cudaEventRecord(h->ev1, h->stream));
    CK(cudaEventSynchronize(h->ev1));
    CK(cudaEventElapsedTime(ms, h->ev0, h->ev1));
    return FB_OK;
}

extern "C" int fb_set_option(fb_handle *h, int32_t option, int32_t value)
{
    if (!h) return FB_ERR_INVALID;
    if (option == FB_OPT_SOLVE_STATS) { h->want_stats = value != 0; return FB_OK; }
    if (option == FB_OPT_HALO_OVERLAP) { h->halo_overlap = value != 0; return FB_OK; }
    return fail(h, FB_ERR_INVALID, "unknown option");
}

extern "C" int fb_profile_enable(fb_handle *h, int32_t on)
{
    if (!h) return FB_ERR_INVALID;
    h->prof = on != 0;
    return FB_OK;
}

extern "C" int fb_profile_read(fb_handle *h, float *ms, int32_t *calls)
{
    if (!h || !ms || !calls) return FB_ERR_INVALID;
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->stream));
    for (int k = 0; k < FB_PROF_NPHASES; k++) { ms[k] = 0.0f; calls[k] = 0; }
    for (auto &pr : h->prof_pairs) {
        float t = 0.0f;
        if (cudaEventElapsedTime(&t, pr.second.first, pr.second.second) == cudaSuccess) { ms[pr.first] += t; calls[pr.first]++; }
        h->prof_pool.push_back(pr.second.first);
        h->prof_pool.push_back(pr.second.second);
    }
    h->prof_pairs.clear();
    return FB_OK;
}

// Diagnostics: div_fast / sqrt_fast of k_confine_fast against the IEEE instructions on n random operand sets.
extern "C" int fb_selftest_fastmath(int32_t device, uint64_t n, uint32_t seed, int32_t mode, uint64_t *counts6)
{
    if (!counts6 || mode < 0 || mode > 3) return FB_ERR_INVALID;
    if (cudaSetDevice(device) != cudaSuccess) return FB_ERR_CUDA;
    unsigned long long *d = nullptr;
    if (cudaMalloc(&d, 6 * sizeof(unsigned long long)) != cudaSuccess) return FB_ERR_CUDA;
    cudaMemset(d, 0, 6 * sizeof(unsigned long long));
    k_selftest_fastmath<<<148 * 8, 256>>>(n, seed, mode, d);
    unsigned long long hcnt[6];
    const cudaError_t e = cudaMemcpy(hcnt, d, sizeof(hcnt), cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) return FB_ERR_CUDA;
    for (int k = 0; k < 6; k++) counts6[k] = hcnt[k];
    return FB_OK;
}

extern "C" int fb_launch_count(const fb_handle *h, uint64_t *count)
{
    if (!h || !count) return FB_ERR_INVALID;
    *count = h->launches;
    return FB_OK;
}

// ---- small building blocks --------------------------------------------------------
// Compute range of this rank for a phase: owned lines widened by `extra` ghost
// lines (clipped to what is allocated and to the domain).
static inline void range(const fb_handle *h, int extra, int &ib, int &ie)
{
    const Grid &g = h->g;
    ib = g.i_lo - extra; ie = g.i_hi + extra;
    if (ib < g.i_alloc0) ib = g.i_alloc0;
    if (ie > g.i_alloc0 + g.lines_alloc) ie = g.i_alloc0 + g.lines_alloc;
}

static int copy_plane(fb_handle *h, float *dst, const float *src)
{
    CK(cudaMemcpyAsync(dst, src, h->plane_floats * sizeof(float), cudaMemcpyDeviceToDevice, h->stream));
    return FB_OK;
}

static int copy_border(fb_handle *h, float *dst, const float *src)
{
    int ib, ie; range(h, h->cfg.ghost, ib, ie);
    const int n = (ie - ib) + 2 * h->g.NY;
    k_copy_border<<<cdiv(n, 256), 256, 0, h->stream>>>(h->g, dst, src, ib, ie);
    CKL("k_copy_border");
    return FB_OK;
}

static int check_bad(fb_handle *h)
{
    if (h->cfg.nranks <= 1) return FB_OK;   // single GPU: every tap is resident
    int bad = 0;
    CK(cudaMemcpyAsync(&bad, h->d_bad, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (bad) {
        CK(cudaMemsetAsync(h->d_bad, 0, sizeof(int), h->stream));
        if (bad == 2) return fail(h, FB_ERR_HALO, "halo exchange: a neighbour never published its boundary lines");
        return fail(h, FB_ERR_HALO, "semi-Lagrangian trace left the ghost zone; create the handle with more ghost lines");
    }
    return FB_OK;
}

// ---- projection -------------------------------------------------------------------
static void omega_schedule(const fb_params *p, unsigned iters, float *omega)
{
    // fluid.go:162-170 (float32 arithmetic; this TU is compiled with --fmad=false)
    const float initial = p->relaxation;
    const float minRelax = 1.2f;
    for (unsigned it = 0; it < iters && it < 32; it++) {
        volatile float prog = (float)it / (float)iters;
        volatile float t = (initial - minRelax) * prog;
        omega[it] = initial - t;
    }
}

// Red-black schedule, one omega per HALF sweep (red, black, red, ...): the reference's
// omega(iter) for every iteration but the last, which closes with a plain Gauss-Seidel
// red half sweep (1.0) and a half-relaxed black one (0.5).  Red-black otherwise leaves
// the entire residual on one colour; the damped close spreads it over both, which is
// what brings max|div| down to the lexicographic solver's.
static void omega_schedule_redblack(const fb_params *p, unsigned iters, float *omega)
{
    float per_iter[32];
    omega_schedule(p, iters, per_iter);
    for (unsigned it = 0; it < iters && it < 32; it++) omega[2 * it] = omega[2 * it + 1] = per_iter[it];
    if (iters > 0) { omega[2 * iters - 2] = 1.0f; omega[2 * iters - 1] = 0.5f; }
}

static int ensure_order(fb_handle *h)
{
    const Grid &g = h->g;
    if (h->d_order) return FB_OK;
    const int T = WF_TMAX;
    h->nTa = cdiv(g.NX - 2 + T - 1, WF_TI);
    h->nTb = cdiv(g.NY - 2 + T - 1, WF_TJ);
    h->ntiles = h->nTa * h->nTb;
    std::vector<int2> order;
    order.reserve(h->ntiles);
    for (int d = 0; d <= h->nTa + h->nTb - 2; d++)
        for (int a = 0; a < h->nTa; a++) {
            const int b = d - a;
            if (b < 0 || b >= h->nTb) continue;
            order.push_back(make_int2(a, b));
        }
    CK(cudaMalloc(&h->d_order, sizeof(int2) * (size_t)h->ntiles));
    CK(cudaMemcpyAsync(h->d_order, order.data(), sizeof(int2) * (size_t)h->ntiles, cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaMalloc(&h->d_done, sizeof(int) * (size_t)h->ntiles));
    CK(cudaMemsetAsync(h->d_done, 0, sizeof(int) * (size_t)h->ntiles, h->stream));
    h->epoch = 0;
    return FB_OK;
}

// Run sweeps [first, first+count) of an `iters`-sweep solve, fused WF_TMAX at a time.
static int exact_sweeps(fb_handle *h, const SolveParams &base, int first, int count)
{
    TRY(ensure_order(h));
    int nsm = 148;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, h->device);
    int done = 0;
    while (done < count) {
        const int T = (count - done) < WF_TMAX ? (count - done) : WF_TMAX;
        SolveParams sp = base;
        sp.sweeps = T;
        sp.sweep0 = first + done;
        CK(cudaMemsetAsync(h->d_tile_counter, 0, sizeof(int), h->stream));
        h->epoch++;
        int grid = h->ntiles < nsm * 4 ? h->ntiles : nsm * 4;
        k_gs_wavefront<<<grid, dim3(WF_TI, T, 1), 0, h->stream>>>(h->g, h->f[FB_U], h->f[FB_V], h->f[FB_S], h->f[FB_P], sp,
                                                                 h->d_order, h->ntiles, h->nTb, h->d_tile_counter,
                                                                 h->d_done, h->epoch, h->d_red);
        CKL("k_gs_wavefront");
        done += T;
    }
    return FB_OK;
}

static int read_stats(fb_handle *h, unsigned iters)
{
    CK(cudaMemcpyAsync(h->h_red, h->d_red, 32 * sizeof(unsigned), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    for (unsigned k = 0; k < 32; k++) {
        float v; unsigned b = h->h_red[k];
        memcpy(&v, &b, 4);
        h->stats.max_div[k] = k < iters ? v : 0.0f;
    }
    return FB_OK;
}

static int project_redblack_fused(fb_handle *h, const fb_params *p, float dt, unsigned iters, bool fuse_turbulence, int ext = 0,
                                  const float *omega_half_sweeps = nullptr);

// solveMultigridVCycle (fluid.go:560-599): per cycle 3 smoothing sweeps at 1.5, residual ->
// restriction -> 40 coarse sweeps at 1.6 -> prolongation -> correction, 3 smoothing sweeps at 1.2.
// FB_SOLVER_EXACT runs every sweep (fine and coarse) in the reference's lexicographic order
// through wavefronts and is bit-identical to the reference; the red-black solvers run the same
// cycle with red-black sweeps on both levels.  stats: max_div[k] = max|div| of the third
// pre-smoothing sweep of cycle k (the value fluid.go:575 tests), sweeps_run = cycles entered.
static int solve_multigrid_vcycle(fb_handle *h, const fb_params *p, float dt, unsigned iters)
{
    if (h->cfg.nranks > 1) return fail(h, FB_ERR_UNSUPPORTED, "the multigrid V-cycle is single-GPU only");
    const Grid &g = h->g;
    CoarseGrid &c = h->cg;
    if (!h->mg_rhs) {
        c.NX = (g.NX + 1) / 2; c.NY = (g.NY + 1) / 2;      // fluid.go:632-633
        c.pitch = cdiv(c.NY, 32) * 32;
        const size_t bytes = (size_t)c.NX * (size_t)c.pitch * sizeof(float);
        CK(cudaMalloc(&h->mg_rhs, bytes)); CK(cudaMalloc(&h->mg_cS, bytes)); CK(cudaMalloc(&h->mg_cP, bytes));
    }
    const bool exact = p->solver == FB_SOLVER_EXACT;
    SolveParams sp;
    memset(&sp, 0, sizeof(sp));
    for (int k = 0; k < 3; k++) { sp.omega[k] = 1.5f; sp.omega[3 + k] = 1.2f; }   // fluid.go:574, 596
    sp.damping = p->pressure_damping;
    { volatile float dh = h->cfg.density * h->cfg.h; sp.cp = dh / dt; }           // fluid.go:561
    const float tolerance = 1e-5f;
    const int coarse_iters = 40;              // fluid.go:689
    const float coarse_relaxation = 1.6f;     // fluid.go:690
    h->stats.sweeps_run = 0;
    h->stats.rolled_back = 0;
    for (unsigned k = 0; k < 32; k++) h->stats.max_div[k] = 0.0f;
    if (iters == 0) return FB_OK;
    h->p_zero = false;

    int ib, ie; range(h, 0, ib, ie);
    dim3 rb_grid, rb_block;
    plane_launch(g, ib, ie, rb_grid, rb_block, g.NY / 2 + 1);
    // FB_SOLVER_REDBLACK_PRESSURE smooths through the fused pressure-form solver, 3 sweeps in one pass over HBM
    // (2.50 -> 1.85 ms per cycle at 4098^2).  FB_SOLVER_REDBLACK keeps the unfused half sweeps: its fused face-form
    // kernel is issue-bound and three sweeps of it cost what six k_redblack_half launches do (measured 2.54 vs 2.50 ms).
    const bool fused_q = p->solver == FB_SOLVER_REDBLACK_PRESSURE;
    const bool want_stats0 = h->want_stats;
    h->want_stats = true;                      // the early exit of fluid.go:575 needs max|div| of the third sweep
    struct Restore { fb_handle *h; bool v; ~Restore() { h->want_stats = v; } } restore{h, want_stats0};
    // three fine-grid sweeps with omega[first .. first+3); the pre-smoothing's max|div| per sweep lands in d_red[0..2]
    auto smooth = [&](int first) -> int {
        if (exact) return exact_sweeps(h, sp, first, 3);
        if (fused_q) {
            float om[6];
            for (int q = 0; q < 6; q++) om[q] = sp.omega[first];
            return project_redblack_fused(h, p, dt, 3, false, 0, om);
        }
        for (int s = 0; s < 3; s++)
            for (int colour = 0; colour < 2; colour++) {
                k_redblack_half<<<rb_grid, rb_block, 0, h->stream>>>(g, h->f[FB_U], h->f[FB_V], h->f[FB_S], h->f[FB_P], colour,
                                                                     sp.omega[first + s], sp.damping, sp.cp, h->d_red + first + s, ib, ie);
                CKL("k_redblack_half");
            }
        return FB_OK;
    };
    const dim3 blk(128, 2, 1);
    for (unsigned iter = 0; iter < iters; iter++) {
        CK(cudaMemsetAsync(h->d_red, 0, 32 * sizeof(unsigned), h->stream));
        TRY(smooth(0));
        CK(cudaMemcpyAsync(h->h_red, h->d_red, 8 * sizeof(unsigned), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        float maxDiv; memcpy(&maxDiv, &h->h_red[2], 4);
        h->stats.max_div[iter] = maxDiv;
        h->stats.sweeps_run = (int)iter + 1;
        if (maxDiv < tolerance) break;                                            // fluid.go:575-577

        k_mg_restrict<<<dim3(cdiv(c.NY, blk.x), cdiv(c.NX, blk.y)), blk, 0, h->stream>>>(g, c, h->f[FB_U], h->f[FB_V], h->f[FB_S], h->f[FB_P],
                                                                                        h->mg_rhs, h->mg_cS, h->mg_cP);
        CKL("k_mg_restrict");
        if (exact) {
            const int tau_last = (c.NX - 2) + (c.NY - 2) + 2 * (coarse_iters - 1);
            const dim3 dgrid(cdiv(c.NX - 2 > 0 ? c.NX - 2 : 1, 128), coarse_iters);
            if (c.NX > 2 && c.NY > 2)
                for (int tau = 2; tau <= tau_last; tau++) {
                    k_mg_coarse_diag<<<dgrid, 128, 0, h->stream>>>(c, h->mg_cP, h->mg_cS, h->mg_rhs, tau, coarse_iters, coarse_relaxation);
                    CKL("k_mg_coarse_diag");
                }
        } else {
            const dim3 cgrid(cdiv(c.NY / 2 + 1, blk.x), cdiv(c.NX - 2 > 0 ? c.NX - 2 : 1, blk.y));
            for (int it = 0; it < coarse_iters; it++)
                for (int colour = 0; colour < 2; colour++) {
                    k_mg_coarse_redblack<<<cgrid, blk, 0, h->stream>>>(c, h->mg_cP, h->mg_cS, h->mg_rhs, colour, coarse_relaxation);
                    CKL("k_mg_coarse_redblack");
                }
        }
        k_mg_apply<<<dim3(cdiv(g.NY, blk.x), cdiv(g.NX, blk.y)), blk, 0, h->stream>>>(g, c, h->f[FB_U], h->f[FB_V], h->f[FB_S], h->f[FB_P],
                                                                                     h->mg_cP, sp.cp);
        CKL("k_mg_apply");
        TRY(smooth(3));
    }
    // fb_get_solve_stats reads the device slots: park the per-cycle values there (the smoothing sweeps used them as scratch)
    unsigned bits[32];
    memcpy(bits, h->stats.max_div, sizeof(bits));
    CK(cudaMemcpyAsync(h->d_red, bits, sizeof(bits), cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return FB_OK;
}

// makeIncompressible (fluid.go:144-155) + solveSingleGrid (fluid.go:157-186)
static int make_incompressible(fb_handle *h, const fb_params *p, float dt, unsigned iters)
{
    if (iters > 32) return fail(h, FB_ERR_INVALID, "at most 32 sweeps per solve");
    TRY(copy_border(h, h->f[FB_NEWU], h->f[FB_U]));
    TRY(copy_border(h, h->f[FB_NEWV], h->f[FB_V]));
    if (p->use_multigrid && p->multigrid_levels > 1) return solve_multigrid_vcycle(h, p, dt, iters);   // fluid.go:148-150
    SolveParams sp;
    memset(&sp, 0, sizeof(sp));
    if (p->solver == FB_SOLVER_EXACT) omega_schedule(p, iters, sp.omega);
    else omega_schedule_redblack(p, iters, sp.omega);
    sp.damping = p->pressure_damping;
    {
        volatile float dh = h->cfg.density * h->cfg.h;
        sp.cp = dh / dt;                                   // fluid.go:158
    }
    CK(cudaMemsetAsync(h->d_red, 0, 32 * sizeof(unsigned), h->stream));
    h->stats.sweeps_run = (int)iters;
    h->stats.rolled_back = 0;
    if (iters == 0) return FB_OK;

    if (p->solver == FB_SOLVER_EXACT) {
        if (h->cfg.nranks > 1) return fail(h, FB_ERR_UNSUPPORTED, "exact (lexicographic) solver is single-GPU only");
        h->p_zero = false;
        // The early exit of fluid.go:175 depends on a full sweep's max|div|, which a
        // fused wavefront only knows afterwards: run optimistically from a backup
        // and, if some sweep k < iters-1 met the tolerance, replay exactly k+1 sweeps.
        float *bkU, *bkV, *bkP;
        TRY(scratch(h, SCR_BKU, &bkU)); TRY(scratch(h, SCR_BKV, &bkV)); TRY(scratch(h, SCR_VIEW, &bkP));
        TRY(copy_plane(h, bkU, h->f[FB_U])); TRY(copy_plane(h, bkV, h->f[FB_V])); TRY(copy_plane(h, bkP, h->f[FB_P]));
        TRY(exact_sweeps(h, sp, 0, (int)iters));
        TRY(read_stats(h, iters));
        const float tolerance = 1e-5f;
        int stop = -1;
        for (unsigned k = 0; k + 1 < iters; k++) if (h->stats.max_div[k] < tolerance) { stop = (int)k; break; }
        if (stop >= 0) {
            TRY(copy_plane(h, h->f[FB_U], bkU)); TRY(copy_plane(h, h->f[FB_V], bkV)); TRY(copy_plane(h, h->f[FB_P], bkP));
            CK(cudaMemsetAsync(h->d_red, 0, 32 * sizeof(unsigned), h->stream));
            TRY(exact_sweeps(h, sp, 0, stop + 1));
            TRY(read_stats(h, (unsigned)stop + 1));
            h->stats.sweeps_run = stop + 1;
            h->stats.rolled_back = 1;
        }
        return FB_OK;
    }

    if (!h->literal || p->solver == FB_SOLVER_REDBLACK_PRESSURE) return project_redblack_fused(h, p, dt, iters, false);

    // red-black, unfused reference path: one launch per half sweep
    h->p_zero = false;
    int ib, ie; range(h, h->cfg.ghost, ib, ie);
    dim3 grid, block;
    plane_launch(h->g, ib, ie, grid, block, h->g.NY / 2 + 1);
    for (unsigned it = 0; it < iters; it++)
        for (int colour = 0; colour < 2; colour++) {
            k_redblack_half<<<grid, block, 0, h->stream>>>(h->g, h->f[FB_U], h->f[FB_V], h->f[FB_S], h->f[FB_P], colour,
                                                           sp.omega[2 * it + colour], sp.damping, sp.cp, h->d_red + it, ib, ie);
            CKL("k_redblack_half");
        }
    return FB_OK;
}

// ---- the other phases ---------------------------------------------------------------
static int clear_pressure(fb_handle *h)   // fluid.go:83
{
    CK(cudaMemsetAsync(h->f[FB_P], 0, h->plane_floats * sizeof(float), h->stream));
    h->p_zero = true;
    return FB_OK;
}

static int apply_viscosity(fb_handle *h, const fb_params *p, float dt)   // fluid.go:112-142
{
    if (!(p->viscosity_diffusion > 0.0f)) return FB_OK;
    volatile float visc = p->viscosity_diffusion * dt;
    TRY(copy_plane(h, h->f[FB_NEWU], h->f[FB_U]));
    TRY(copy_plane(h, h->f[FB_NEWV], h->f[FB_V]));
    int ib, ie; range(h, 0, ib, ie);
    dim3 grid, block; plane_launch(h->g, ib, ie, grid, block);
    k_viscosity<<<grid, block, 0, h->stream>>>(h->g, h->f[FB_U], h->f[FB_V], h->f[FB_S], h->f[FB_NEWU], h->f[FB_NEWV], visc, ib, ie);
    CKL("k_viscosity");
    TRY(copy_plane(h, h->f[FB_U], h->f[FB_NEWU]));
    TRY(copy_plane(h, h->f[FB_V], h->f[FB_NEWV]));
    return FB_OK;
}

static int confinement(fb_handle *h, const fb_params *p, float dt)   // fluid.go:449-493
{
    float *curl;
    TRY(scratch(h, SCR_CURL, &curl));
    int ib, ie;
    range(h, 1, ib, ie);
    dim3 grid, block; plane_launch(h->g, ib, ie, grid, block);
    k_curl<<<grid, block, 0, h->stream>>>(h->g, h->f[FB_U], h->f[FB_V], h->f[FB_S], curl, h->cfg.h, ib, ie);
    CKL("k_curl");
    range(h, 0, ib, ie);
    plane_launch(h->g, ib, ie, grid, block);
    k_confine<<<grid, block, 0, h->stream>>>(h->g, h->f[FB_U], h->f[FB_V], h->f[FB_S], curl, h->cfg.h, dt, p->confinement, ib, ie);
    CKL("k_confine");
    return FB_OK;
}

static int turbulence(fb_handle *h, const fb_params *p, float dt)   // fluid.go:496-526
{
    if (!(p->turbulence_strength > 0.0f)) return FB_OK;
    float *nU, *nV;
    TRY(scratch(h, SCR_NOISEU, &nU)); TRY(scratch(h, SCR_NOISEV, &nV));
    int ib, ie; range(h, 0, ib, ie);
    dim3 grid, block; plane_launch(h->g, ib, ie, grid, block);
    if (!h->noise_ready) {
        k_noise_init<<<grid, block, 0, h->stream>>>(h->g, nU, nV, ib, ie);
        CKL("k_noise_init");
        h->noise_ready = true;
    }
    volatile float ts = p->turbulence_strength * dt;
    k_turbulence<<<grid, block, 0, h->stream>>>(h->g, h->f[FB_U], h->f[FB_V], h->f[FB_S], nU, nV, ts, ib, ie);
    CKL("k_turbulence");
    return FB_OK;
}

static int handle_borders(fb_handle *h, int ext = 0)   // fluid.go:236-289
{
    int ib, ie; range(h, ext, ib, ie);
    const int n = (ie - ib) + 2 * h->g.NY;
    k_handle_borders<<<cdiv(n, 256), 256, 0, h->stream>>>(h->g, h->f[FB_U], h->f[FB_V], h->f[FB_S], ib, ie);
    CKL("k_handle_borders");
    return FB_OK;
}

static int advect_velocity(fb_handle *h, float dt)   // fluid.go:291-333
{
    TRY(copy_border(h, h->f[FB_NEWU], h->f[FB_U]));
    TRY(copy_border(h, h->f[FB_NEWV], h->f[FB_V]));
    int ib, ie; range(h, 0, ib, ie);
    dim3 grid, block; plane_launch(h->g, ib, ie, grid, block);
    k_trace_velocity<false><<<grid, block, 0, h->stream>>>(h->g, h->f[FB_U], h->f[FB_V], h->f[FB_S], h->f[FB_U], h->f[FB_V],
                                                           h->f[FB_NEWU], h->f[FB_NEWV], dt, h->cfg.h, ib, ie, h->d_bad);
    CKL("k_trace_velocity");
    TRY(copy_plane(h, h->f[FB_U], h->f[FB_NEWU]));   // copy(f.U, f.newU), fluid.go:331 (Q-6)
    TRY(copy_plane(h, h->f[FB_V], h->f[FB_NEWV]));
    return FB_OK;
}

static int advect_smoke(fb_handle *h, const fb_params *p, float dt)   // fluid.go:400-434
{
    TRY(copy_border(h, h->f[FB_NEWM], h->f[FB_M]));
    int ib, ie; range(h, 0, ib, ie);
    dim3 grid, block; plane_launch(h->g, ib, ie, grid, block);
    k_trace_smoke<false><<<grid, block, 0, h->stream>>>(h->g, h->f[FB_U], h->f[FB_V], h->f[FB_S], h->f[FB_M], h->f[FB_M],
                                                        h->f[FB_NEWM], dt, h->cfg.h, p->smoke_advection,
                                                        p->viscosity_diffusion, ib, ie, h->d_bad);
    CKL("k_trace_smoke");
    TRY(copy_plane(h, h->f[FB_M], h->f[FB_NEWM]));
    return FB_OK;
}

static int advect_velocity_bfecc(fb_handle *h, float dt)   // fluid.go:911-994 (Q-10)
{
    if (h->cfg.nranks > 1) return fail(h, FB_ERR_UNSUPPORTED, "BFECC needs halo exchanges between its passes; drive it per pass from the host layer");
    float *origU, *origV, *fwdU, *fwdV, *bwdU, *bwdV;
    TRY(scratch(h, SCR_A, &origU)); TRY(scratch(h, SCR_B, &origV));
    TRY(scratch(h, SCR_C, &fwdU)); TRY(scratch(h, SCR_D, &fwdV));
    TRY(scratch(h, SCR_E, &bwdU)); TRY(scratch(h, SCR_F, &bwdV));
    TRY(copy_plane(h, origU, h->f[FB_U])); TRY(copy_plane(h, origV, h->f[FB_V]));
    TRY(advect_velocity(h, dt));
    TRY(copy_plane(h, fwdU, h->f[FB_U])); TRY(copy_plane(h, fwdV, h->f[FB_V]));
    TRY(copy_plane(h, h->f[FB_U], origU)); TRY(copy_plane(h, h->f[FB_V], origV));
    CK(cudaMemsetAsync(bwdU, 0, h->plane_floats * sizeof(float), h->stream));
    CK(cudaMemsetAsync(bwdV, 0, h->plane_floats * sizeof(float), h->stream));
    TRY(copy_border(h, bwdU, fwdU)); TRY(copy_border(h, bwdV, fwdV));
    int ib, ie; range(h, 0, ib, ie);
    dim3 grid, block; plane_launch(h->g, ib, ie, grid, block);
    k_trace_velocity<true><<<grid, block, 0, h->stream>>>(h->g, h->f[FB_U], h->f[FB_V], h->f[FB_S], fwdU, fwdV, bwdU, bwdV,
                                                          dt, h->cfg.h, ib, ie, h->d_bad);
    CKL("k_trace_velocity<back>");
    // corrected field goes straight into U,V (copy(f.U, corrU), fluid.go:991)
    k_bfecc_correct<false><<<grid, block, 0, h->stream>>>(h->g, origU, bwdU, h->f[FB_U], ib, ie);
    CKL("k_bfecc_correct");
    k_bfecc_correct<false><<<grid, block, 0, h->stream>>>(h->g, origV, bwdV, h->f[FB_V], ib, ie);
    CKL("k_bfecc_correct");
    TRY(advect_velocity(h, dt));
    return FB_OK;
}

static int advect_smoke_bfecc(fb_handle *h, const fb_params *p, float dt)   // fluid.go:997-1051 (Q-11)
{
    if (h->cfg.nranks > 1) return fail(h, FB_ERR_UNSUPPORTED, "BFECC needs halo exchanges between its passes; drive it per pass from the host layer");
    float *origM, *fwdM, *bwdM;
    TRY(scratch(h, SCR_A, &origM)); TRY(scratch(h, SCR_C, &fwdM)); TRY(scratch(h, SCR_E, &bwdM));
    TRY(copy_plane(h, origM, h->f[FB_M]));
    TRY(advect_smoke(h, p, dt));
    TRY(copy_plane(h, fwdM, h->f[FB_M]));
    CK(cudaMemsetAsync(bwdM, 0, h->plane_floats * sizeof(float), h->stream));
    TRY(copy_border(h, bwdM, fwdM));
    int ib, ie; range(h, 0, ib, ie);
    dim3 grid, block; plane_launch(h->g, ib, ie, grid, block);
    k_trace_smoke<true><<<grid, block, 0, h->stream>>>(h->g, h->f[FB_U], h->f[FB_V], h->f[FB_S], h->f[FB_M], fwdM, bwdM, dt,
                                                       h->cfg.h, p->smoke_advection, p->viscosity_diffusion, ib, ie, h->d_bad);
    CKL("k_trace_smoke<back>");
    k_bfecc_correct<true><<<grid, block, 0, h->stream>>>(h->g, origM, bwdM, h->f[FB_M], ib, ie);
    CKL("k_bfecc_correct<nonneg>");
    TRY(advect_smoke(h, p, dt));
    return FB_OK;
}

// ======================= fast path: fused kernels, pointer swaps =======================
static int ensure_mask(fb_handle *h)
{
    if (!h->mask_dirty) return FB_OK;
    const Grid &g = h->g;
    const int ib = g.i_alloc0, ie = g.i_alloc0 + g.lines_alloc;
    dim3 grid, block; plane_launch(g, ib, ie, grid, block);
    k_build_mask<<<grid, block, 0, h->stream>>>(g, h->f[FB_S], h->mask, ib, ie);
    CKL("k_build_mask");
    if (!h->tile_flags) {
        h->tile_ntx = cdiv(g.NY, AT_TJ); h->tile_nty = cdiv(g.NX, AT_TI);
        CK(cudaMalloc(&h->tile_flags, (size_t)h->tile_ntx * h->tile_nty));
    }
    k_tile_flags<<<dim3(h->tile_ntx, h->tile_nty, 1), 256, 0, h->stream>>>(g, h->mask, h->tile_flags, h->tile_ntx);
    CKL("k_tile_flags");
    h->mask_dirty = false;
    return FB_OK;
}

static AdvCtx adv_ctx(const fb_handle *h)
{
    const Grid &g = h->g;
    AdvCtx c;
    c.NX = g.NX; c.NY = g.NY; c.pitch = g.pitch; c.i_alloc0 = g.i_alloc0; c.lines_alloc = g.lines_alloc;
    c.xr_hi = (g.i_alloc0 + g.lines_alloc == g.NX) ? g.lines_alloc - 1 : g.lines_alloc - 2;
    volatile float hh = h->cfg.h;
    volatile float h1 = 1.0f / hh, h2 = hh / 2.0f;
    volatile float xmax = (float)g.NX * hh, ymax = (float)g.NY * hh;
    c.h = hh; c.h1 = h1; c.h2 = h2; c.xmax = xmax; c.ymax = ymax;
    c.nx1f = (float)(g.NX - 1); c.ny1f = (float)(g.NY - 1);
    return c;
}

// launch helpers: CHECK (ghost-zone guard) only when the grid is split over ranks
#define ADV_LAUNCH(kern, ...) do { \
        ProfScope _ks(h, PROF_SLOT_##kern); \
        dim3 _grid, _block; adv_launch(h->g.NY, ib, ie, _grid, _block); \
        if (h->cfg.nranks > 1) kern<true><<<_grid, _block, 0, h->stream>>>(__VA_ARGS__); \
        else kern<false><<<_grid, _block, 0, h->stream>>>(__VA_ARGS__); \
        CKL(#kern); } while (0)

// advectVelocity / the BFECC correct pass on shared-memory tiles (advect_tile.cuh); FLUIDB200_ADV_FULL=1 selects the
// global-gather kernels of round 1 (A/B).  Arguments: (c, srcU, srcV, mask, inU, inV, outU, outV, dt, ib, ie, bad) as for
// k_advect_velocity_full / k_bfecc_velocity_correct.
// 2-D tensor map of a plane for the tile kernels' TMA loads: [lines_alloc][pitch] floats, box = box_lines x AT_PW, elements
// outside the plane read as zero.  Planes live as long as the handle, so the maps are made once per (plane, box).
static int tile_tmap(fb_handle *h, const float *plane, int box_lines, CUtensorMap *out)
{
    const auto key = std::make_pair((const void *)plane, box_lines);
    auto it = h->tmaps.find(key);
    if (it != h->tmaps.end()) { *out = it->second; return FB_OK; }
    static PFN_cuTensorMapEncodeTiled encode = nullptr;
    if (!encode) {
        cudaDriverEntryPointQueryResult q;
        void *fn = nullptr;
        CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
        if (!fn || q != cudaDriverEntryPointSuccess) return fail(h, FB_ERR_CUDA, "cuTensorMapEncodeTiled is not available in this driver");
        encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled>(fn);
    }
    const Grid &g = h->g;
    const cuuint64_t dims[2] = { (cuuint64_t)g.pitch, (cuuint64_t)g.lines_alloc };
    const cuuint64_t strides[1] = { (cuuint64_t)g.pitch * sizeof(float) };
    const cuuint32_t box[2] = { (cuuint32_t)AT_PW, (cuuint32_t)box_lines };
    const cuuint32_t estr[2] = { 1, 1 };
    CUtensorMap tm;
    const CUresult r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(plane), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(h, FB_ERR_CUDA, "cuTensorMapEncodeTiled failed");
    h->tmaps[key] = tm;
    *out = tm;
    return FB_OK;
}

static int adv_tile_attrs(fb_handle *h)
{
    if (h->adv_tile_attr_set) return FB_OK;
    CK(cudaFuncSetAttribute(k_advect_velocity_tile<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM));
    CK(cudaFuncSetAttribute(k_advect_velocity_tile<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM));
    CK(cudaFuncSetAttribute(k_bfecc_velocity_tile<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_BSMEM));
    CK(cudaFuncSetAttribute(k_bfecc_velocity_tile<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_BSMEM));
    CK(cudaFuncSetAttribute(k_confine_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_CSMEM));
    CK(cudaFuncSetAttribute(k_advect_smoke_tile<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SSMEM));
    CK(cudaFuncSetAttribute(k_advect_smoke_tile<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SSMEM));
    CK(cudaFuncSetAttribute(k_bfecc_smoke_tile<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SBSMEM));
    CK(cudaFuncSetAttribute(k_bfecc_smoke_tile<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SBSMEM));
    h->adv_tile_attr_set = true;
    return FB_OK;
}
static const bool adv_full = getenv("FLUIDB200_ADV_FULL") != nullptr;
#define ADV_VELOCITY_LAUNCH(c, sU, sV, mk, aU, aV, oU, oV, dt, ib, ie, bad) do { \
        if (adv_full) { ADV_LAUNCH(k_advect_velocity_full, c, sU, sV, mk, aU, aV, oU, oV, dt, ib, ie, bad); break; } \
        TRY(adv_tile_attrs(h)); \
        CUtensorMap _tU, _tV; TRY(tile_tmap(h, sU, AT_TL, &_tU)); TRY(tile_tmap(h, sV, AT_TL, &_tV)); \
        ProfScope _ks(h, FB_PROF_K_ADVECT_VELOCITY); \
        const dim3 _grid(h->tile_ntx, cdiv(ie, AT_TI) - (ib) / AT_TI, 1); \
        if (h->cfg.nranks > 1) k_advect_velocity_tile<true><<<_grid, AT_THREADS, AT_SMEM, h->stream>>>(c, _tU, _tV, sU, sV, mk, h->tile_flags, h->tile_ntx, aU, aV, oU, oV, dt, ib, ie, bad); \
        else k_advect_velocity_tile<false><<<_grid, AT_THREADS, AT_SMEM, h->stream>>>(c, _tU, _tV, sU, sV, mk, h->tile_flags, h->tile_ntx, aU, aV, oU, oV, dt, ib, ie, bad); \
        CKL("k_advect_velocity_tile"); } while (0)
// advectSmoke / the BFECC smoke correct pass on tiles; the diffusion term (viscosityDiffusion > 0) stays with the round-1 kernel
#define ADV_SMOKE_LAUNCH(c, sU, sV, mk, sM, shM, oM, dt, sa, visc, ib, ie, bad) do { \
        if (adv_full || (visc) > 0.0f) { ADV_LAUNCH(k_advect_smoke_full, c, sU, sV, mk, sM, shM, oM, dt, sa, visc, ib, ie, bad); break; } \
        TRY(adv_tile_attrs(h)); \
        CUtensorMap _tU, _tV, _tM; TRY(tile_tmap(h, sU, AT_SVL, &_tU)); TRY(tile_tmap(h, sV, AT_SVL, &_tV)); TRY(tile_tmap(h, sM, AT_STL, &_tM)); \
        ProfScope _ks(h, FB_PROF_K_ADVECT_SMOKE); \
        const dim3 _grid(h->tile_ntx, cdiv(ie, AT_STI) - (ib) / AT_STI, 1); \
        if (h->cfg.nranks > 1) k_advect_smoke_tile<true><<<_grid, AT_THREADS, AT_SSMEM, h->stream>>>(c, _tU, _tV, _tM, sM, mk, h->tile_flags, h->tile_ntx, shM, oM, dt, sa, ib, ie, bad); \
        else k_advect_smoke_tile<false><<<_grid, AT_THREADS, AT_SSMEM, h->stream>>>(c, _tU, _tV, _tM, sM, mk, h->tile_flags, h->tile_ntx, shM, oM, dt, sa, ib, ie, bad); \
        CKL("k_advect_smoke_tile"); } while (0)
#define ADV_BFECC_SMOKE_LAUNCH(c, sU, sV, mk, oM, fM, cM, dt, sa, ib, ie, bad) do { \
        if (adv_full) { ADV_LAUNCH(k_bfecc_smoke_correct, c, sU, sV, mk, oM, fM, cM, dt, sa, ib, ie, bad); break; } \
        TRY(adv_tile_attrs(h)); \
        CUtensorMap _tU, _tV, _tO, _tF; \
        TRY(tile_tmap(h, sU, AT_SVL, &_tU)); TRY(tile_tmap(h, sV, AT_SVL, &_tV)); TRY(tile_tmap(h, oM, AT_SVL, &_tO)); TRY(tile_tmap(h, fM, AT_STL, &_tF)); \
        ProfScope _ks(h, FB_PROF_K_BFECC_SMOKE); \
        const dim3 _grid(h->tile_ntx, cdiv(ie, AT_STI) - (ib) / AT_STI, 1); \
        if (h->cfg.nranks > 1) k_bfecc_smoke_tile<true><<<_grid, AT_THREADS, AT_SBSMEM, h->stream>>>(c, _tU, _tV, _tO, _tF, fM, mk, h->tile_flags, h->tile_ntx, cM, dt, sa, ib, ie, bad); \
        else k_bfecc_smoke_tile<false><<<_grid, AT_THREADS, AT_SBSMEM, h->stream>>>(c, _tU, _tV, _tO, _tF, fM, mk, h->tile_flags, h->tile_ntx, cM, dt, sa, ib, ie, bad); \
        CKL("k_bfecc_smoke_tile"); } while (0)
#define ADV_BFECC_VELOCITY_LAUNCH(c, sU, sV, mk, fU, fV, oU, oV, dt, ib, ie, bad) do { \
        if (adv_full) { ADV_LAUNCH(k_bfecc_velocity_correct, c, sU, sV, mk, fU, fV, oU, oV, dt, ib, ie, bad); break; } \
        TRY(adv_tile_attrs(h)); \
        CUtensorMap _tU, _tV, _tFU, _tFV; \
        TRY(tile_tmap(h, sU, AT_BVL, &_tU)); TRY(tile_tmap(h, sV, AT_BVL, &_tV)); TRY(tile_tmap(h, fU, AT_BTL, &_tFU)); TRY(tile_tmap(h, fV, AT_BTL, &_tFV)); \
        ProfScope _ks(h, FB_PROF_K_BFECC_VELOCITY); \
        const dim3 _grid(h->tile_ntx, cdiv(ie, AT_BTI) - (ib) / AT_BTI, 1); \
        if (h->cfg.nranks > 1) k_bfecc_velocity_tile<true><<<_grid, AT_THREADS, AT_BSMEM, h->stream>>>(c, _tU, _tV, _tFU, _tFV, sU, sV, mk, h->tile_flags, h->tile_ntx, fU, fV, oU, oV, dt, ib, ie, bad); \
        else k_bfecc_velocity_tile<false><<<_grid, AT_THREADS, AT_BSMEM, h->stream>>>(c, _tU, _tV, _tFU, _tFV, sU, sV, mk, h->tile_flags, h->tile_ntx, fU, fV, oU, oV, dt, ib, ie, bad); \
        CKL("k_bfecc_velocity_tile"); } while (0)

// newX := X as the reference's copy() leaves them, only when the caller asked for it
static int sync_shadow(fb_handle *h, int live, int shadow)
{
    if (!h->exact_shadow) return FB_OK;
    return copy_plane(h, h->f[shadow], h->f[live]);
}

// Geometry of the fused red-black pass for this grid.
static void rb_geometry(const fb_handle *h, int ib, int ie, int &TJ, int &WL, int &nstrips, int &chunk, int &nchunks)
{
    const Grid &g = h->g;
    nstrips = cdiv(g.NY, RB_TJ_MAX);
    TJ = cdiv(cdiv(g.NY, nstrips), 4) * 4;
    nstrips = cdiv(g.NY, TJ);
    WL = TJ + 2 * RB_H + 4;
    const int lines = ie - ib;
    nchunks = h->nsm / nstrips;
    if (nchunks < 1) nchunks = 1;
    const int max_chunks = cdiv(lines, 48);           // keep the halo overhead (32 lines) bounded
    if (nchunks > max_chunks) nchunks = max_chunks;
    if (nchunks < 1) nchunks = 1;
    chunk = cdiv(lines, nchunks);
    nchunks = cdiv(lines, chunk);
}

// makeIncompressible with the red-black ordering, every iteration fused in one pass
// over HBM (two passes when iters > 8).  Optionally applies addTurbulence to the lines
// it writes.  Outputs go to fresh planes; roles are swapped afterwards.
// `omega_half_sweeps` (2 * iters values: red, black, red, ...) replaces the schedule of fluid.go:169-170 --
// the V-cycle's smoothing sweeps run at a constant relaxation (fluid.go:574, 596).
static int project_redblack_fused(fb_handle *h, const fb_params *p, float dt, unsigned iters, bool fuse_turbulence, int ext,
                                  const float *omega_half_sweeps)
{
    TRY(ensure_mask(h));
    SolveParams sp;
    memset(&sp, 0, sizeof(sp));
    if (omega_half_sweeps) for (unsigned q = 0; q < 2 * iters && q < 64; q++) sp.omega[q] = omega_half_sweeps[q];
    else omega_schedule_redblack(p, iters, sp.omega);
    { volatile float dh = h->cfg.density * h->cfg.h; sp.cp = dh / dt; }
    // with slabs `ext` ghost lines are recomputed too, so that the next phase finds them current
    int ib, ie; range(h, ext, ib, ie);
    int TJ, WL, nstrips, chunk, nchunks;
    rb_geometry(h, ib, ie, TJ, WL, nstrips, chunk, nchunks);
    const size_t smem = (size_t)RB_NL * WL * 13;
    if (!h->rb_attr_set) {
        CK(cudaFuncSetAttribute(k_rb_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        h->rb_attr_set = true;
    }
    float *nU = nullptr, *nV = nullptr;
    if (fuse_turbulence) {
        TRY(scratch(h, SCR_NOISEU, &nU)); TRY(scratch(h, SCR_NOISEV, &nV));
        if (!h->noise_ready) {
            const Grid &g = h->g;
            dim3 grid, block; plane_launch(g, g.i_alloc0, g.i_alloc0 + g.lines_alloc, grid, block);
            k_noise_init<<<grid, block, 0, h->stream>>>(g, nU, nV, g.i_alloc0, g.i_alloc0 + g.lines_alloc);
            CKL("k_noise_init");
            h->noise_ready = true;
        }
    }
    const bool pressure_form = p->solver == FB_SOLVER_REDBLACK_PRESSURE;
    // k_rbq_fused (shared-memory ring, 24 warps per SM) is the default: the register pipelines of rbq_stream.cuh measured slower
    // at 4098^2 (0.25 - 0.32 ms against 0.18); FLUIDB200_RBQ_STREAM=1 selects them (A/B, profiles/r02_rbq_stream.md)
    static const bool rbq_ring = getenv("FLUIDB200_RBQ_STREAM") == nullptr;
    if (pressure_form && !rbq_ring) {
        // k_rbq_stream: one warp (= CTA) per strip x chunk; all warps resident in ONE wave, least halo recomputation
        const int lines = ie - ib;
        double best = -1.0;
        const int s_min = cdiv(h->g.NY, RS_TJ_MAX);
        for (int ns = s_min; ns <= s_min + 8; ns++) {
            const int tj = cdiv(cdiv(h->g.NY, ns), 16) * 16;
            if (tj > RS_TJ_MAX || tj < 16) continue;
            const int nstr = cdiv(h->g.NY, tj);
            const int slots = h->nsm * RS_CPS;
            int nch = slots / nstr; if (nch < 1) nch = 1;
            const int max_chunks = cdiv(lines, 32);
            if (nch > max_chunks) nch = max_chunks;
            if (nch < 1) nch = 1;
            const int ch = cdiv(lines, nch);
            nch = cdiv(lines, ch);
            const int ctas = nstr * nch;
            const int waves = cdiv(ctas, slots);
            const double util = (double)ctas / (waves * slots);
            const double overhead = ((double)RS_W / tj) * ((double)(ch + 2 * RS_H) / ch);
            const double score = util / overhead / waves;
            if (score > best) { best = score; TJ = tj; nstrips = nstr; chunk = ch; nchunks = nch; }
        }
        WL = RS_W;
        if (!h->rbq_attr_set) {
            CK(cudaFuncSetAttribute(k_rbq_stream<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, RS_SMEM));
            CK(cudaFuncSetAttribute(k_rbq_stream<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, RS_SMEM));
            h->rbq_attr_set = true;
        }
    } else if (pressure_form) {
        // pick strips x chunks: as many of the SMs as possible in ONE wave, least halo recomputation
        const int lines = ie - ib;
        double best = -1.0;
        const int s_min = cdiv(h->g.NY, RQ_TJ_MAX);
        for (int ns = s_min; ns <= s_min + 12; ns++) {
            const int tj = cdiv(cdiv(h->g.NY, ns), 16) * 16;
            if (tj > RQ_TJ_MAX || tj < 16) continue;
            const int nstr = cdiv(h->g.NY, tj);
            const int slots = h->nsm * RQ_MINB;                          // CTAs resident at once
            int nch = slots / nstr; if (nch < 1) nch = 1;
            const int max_chunks = cdiv(lines, 48);
            if (nch > max_chunks) nch = max_chunks;
            if (nch < 1) nch = 1;
            const int ch = cdiv(lines, nch);
            nch = cdiv(lines, ch);
            const int ctas = nstr * nch;
            const int waves = cdiv(ctas, slots);
            const double util = (double)ctas / (waves * slots);
            const double overhead = ((double)(tj + 2 * RQ_H + 16) / tj) * ((double)(ch + 2 * RQ_H) / ch);
            const double score = util / overhead / waves;
            if (score > best) { best = score; TJ = tj; nstrips = nstr; chunk = ch; nchunks = nch; }
        }
        WL = RQ_WL;
        if (!h->rbq_attr_set) {
            // static shared memory (the per-CTA debug pointer) counts against the 227 KB too
            CK(cudaFuncSetAttribute(k_rbq_fused<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
            CK(cudaFuncSetAttribute(k_rbq_fused<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
            h->rbq_attr_set = true;
        }
    }
    unsigned done = 0;
    while (pressure_form && done < iters) {
        const unsigned k = iters - done < 8 ? iters - done : 8;
        float *Uo, *Vo, *Po;
        TRY(take_plane(h, &Uo)); TRY(take_plane(h, &Vo)); TRY(take_plane(h, &Po));
        RBQ a;
        memset(&a, 0, sizeof(a));
        a.g = h->g;
        a.U = h->f[FB_U]; a.V = h->f[FB_V];
        a.Pin = h->p_zero ? nullptr : h->f[FB_P];
        a.mask = h->mask;
        a.Uo = Uo; a.Vo = Vo; a.Po = Po;
        for (unsigned q = 0; q < 2 * k; q++) {
            volatile float w = sp.omega[2 * done + q] * p->pressure_damping;
            volatile float w4 = w * 0.25f;
            a.wd[q] = w; a.nwd[q] = -w; a.c4[q] = w4;
        }
        a.cp = sp.cp;
        a.nstages = (int)(2 * k); a.stage0 = (int)(2 * done);
        a.TJ = TJ; a.WL = WL; a.chunk = chunk; a.ib = ib; a.ie = ie;
        a.stats = h->d_red;
        a.debug = reinterpret_cast<int *>(h->d_red + 48);
        { const char *x = getenv("FLUIDB200_RBQ_X"); a.xflags = x ? atoi(x) : 0; }
#ifdef RQ_TRACE
        static long long *d_trace = nullptr;
        if (!d_trace) CK(cudaMalloc(&d_trace, 10 * 512 * 4 * sizeof(long long)));
        CK(cudaMemsetAsync(d_trace, 0, 10 * 512 * 4 * sizeof(long long), h->stream));
        a.trace = d_trace;
#endif
        const bool last = done + k == iters;
        if (fuse_turbulence && last) {
            volatile float ts = p->turbulence_strength * dt;
            a.noiseU = nU; a.noiseV = nV; a.turb = ts;
        }
        const size_t smem_q = rq_smem_bytes(WL, TJ);   // planes + TMA staging ring + mbarriers + progress counters
        ProfScope _ks(h, FB_PROF_K_PRESSURE_SOLVE);
        if (!rbq_ring) {
            // fewer than 8 iterations: the trailing stages run with wd = 0 (memset above), which leaves q as it is
            if (h->want_stats) k_rbq_stream<true><<<dim3(nstrips, nchunks, 1), RS_THREADS, RS_SMEM, h->stream>>>(a);
            else k_rbq_stream<false><<<dim3(nstrips, nchunks, 1), RS_THREADS, RS_SMEM, h->stream>>>(a);
            CKL("k_rbq_stream");
        } else {
            if (h->want_stats) k_rbq_fused<true><<<dim3(nstrips, nchunks, 1), RQ_THREADS, smem_q, h->stream>>>(a);
            else k_rbq_fused<false><<<dim3(nstrips, nchunks, 1), RQ_THREADS, smem_q, h->stream>>>(a);
            CKL("k_rbq_fused");
#ifdef RQ_TRACE
            if (const char *path = getenv("FLUIDB200_RBQ_TRACE")) {
                std::vector<long long> host(10 * 512 * 4);
                CK(cudaMemcpyAsync(host.data(), a.trace, host.size() * sizeof(long long), cudaMemcpyDeviceToHost, h->stream));
                CK(cudaStreamSynchronize(h->stream));
                if (FILE *fp = fopen(path, "wb")) { fwrite(host.data(), sizeof(long long), host.size(), fp); fclose(fp); }
            }
#endif
        }
        give_plane(h, h->f[FB_U]); give_plane(h, h->f[FB_V]); give_plane(h, h->f[FB_P]);
        h->f[FB_U] = Uo; h->f[FB_V] = Vo; h->f[FB_P] = Po;
        h->p_zero = false;
        done += k;
    }
    while (done < iters) {
        const unsigned k = iters - done < 8 ? iters - done : 8;
        float *Uo, *Vo, *Po;
        TRY(take_plane(h, &Uo)); TRY(take_plane(h, &Vo)); TRY(take_plane(h, &Po));
        RBFused a;
        memset(&a, 0, sizeof(a));
        a.g = h->g;
        a.U = h->f[FB_U]; a.V = h->f[FB_V];
        a.Pin = h->p_zero ? nullptr : h->f[FB_P];
        a.mask = h->mask;
        a.Uo = Uo; a.Vo = Vo; a.Po = Po;
        for (unsigned q = 0; q < 2 * k; q++) a.omega[q] = sp.omega[2 * done + q];
        a.damping = p->pressure_damping; a.cp = sp.cp;
        a.nstages = (int)(2 * k); a.stage0 = (int)(2 * done);
        a.TJ = TJ; a.WL = WL; a.chunk = chunk; a.ib = ib; a.ie = ie;
        a.stats = h->d_red;
        const bool last = done + k == iters;
        if (fuse_turbulence && last) {
            volatile float ts = p->turbulence_strength * dt;
            a.noiseU = nU; a.noiseV = nV; a.turb = ts;
        }
        k_rb_fused<<<dim3(nstrips, nchunks, 1), RB_THREADS, smem, h->stream>>>(a);
        CKL("k_rb_fused");
        give_plane(h, h->f[FB_U]); give_plane(h, h->f[FB_V]); give_plane(h, h->f[FB_P]);
        h->f[FB_U] = Uo; h->f[FB_V] = Vo; h->f[FB_P] = Po;
        h->p_zero = false;
        done += k;
    }
    return FB_OK;
}

static int confine_turbulence_fast(fb_handle *h, const fb_params *p, float dt, bool do_confine, bool do_turb, int ext = 0)
{
    if (!do_confine && !do_turb) return FB_OK;
    TRY(ensure_mask(h));
    float *nU = nullptr, *nV = nullptr;
    const Grid &g = h->g;
    if (do_turb) {
        TRY(scratch(h, SCR_NOISEU, &nU)); TRY(scratch(h, SCR_NOISEV, &nV));
        if (!h->noise_ready) {
            dim3 grid, block; plane_launch(g, g.i_alloc0, g.i_alloc0 + g.lines_alloc, grid, block);
            k_noise_init<<<grid, block, 0, h->stream>>>(g, nU, nV, g.i_alloc0, g.i_alloc0 + g.lines_alloc);
            CKL("k_noise_init");
            h->noise_ready = true;
        }
    }
    float *dU, *dV;
    TRY(take_plane(h, &dU)); TRY(take_plane(h, &dV));
    int ib, ie; range(h, ext, ib, ie);
    dim3 grid(cdiv(g.NY, CT_J), cdiv(ie - ib, CT_I), 1);
    volatile float ts = do_turb ? p->turbulence_strength * dt : 0.0f;
    // k_confine_tile (TMA-staged tiles, lane per cell) measured SLOWER than this kernel (0.158 against 0.148 ms with
    // confinement, 0.065 against ~0.05 turbulence only): the pass is division / square-root bound and four cells per thread
    // give it the instruction-level parallelism a lane per cell lacks.  FLUIDB200_CONFINE_TILE=1 selects the tile form (A/B).
    static const bool confine_tile = getenv("FLUIDB200_CONFINE_TILE") != nullptr;
    ProfScope _ks(h, FB_PROF_K_CONFINE_TURBULENCE);
    // k_confine_fast: the same pass with the divisions / square roots as straight-line code (advect_fused.cuh); needs h inside
    // the divisor range its sequences are exact for.  FLUIDB200_CONFINE_IEEE=1 keeps the IEEE-instruction kernel (A/B).
    static const bool confine_ieee = getenv("FLUIDB200_CONFINE_IEEE") != nullptr;
    if (!confine_ieee && !confine_tile && h->cfg.h >= FD_BLO && h->cfg.h < FD_HHI) {
        k_confine_fast<<<grid, CT_J * CT_I / 4, 0, h->stream>>>(g, h->f[FB_U], h->f[FB_V], h->mask, nU, nV, dU, dV, h->cfg.h, dt,
                                                              do_confine ? p->confinement : 0.0f, ts, ib, ie);
        CKL("k_confine_fast");
    } else if (adv_full || !confine_tile) {
        k_confine_turbulence<<<grid, CT_J * CT_I / 4, 0, h->stream>>>(g, h->f[FB_U], h->f[FB_V], h->mask, nU, nV, dU, dV, h->cfg.h, dt,
                                                                    do_confine ? p->confinement : 0.0f, ts, ib, ie);
        CKL("k_confine_turbulence");
    } else {
        TRY(adv_tile_attrs(h));
        CUtensorMap tU, tV;
        TRY(tile_tmap(h, h->f[FB_U], AT_CVL, &tU)); TRY(tile_tmap(h, h->f[FB_V], AT_CVL, &tV));
        const AdvCtx c = adv_ctx(h);
        const dim3 tgrid(h->tile_ntx, cdiv(ie, AT_CTI) - ib / AT_CTI, 1);
        k_confine_tile<<<tgrid, AT_THREADS, AT_CSMEM, h->stream>>>(c, tU, tV, h->mask, nU, nV, dU, dV, h->cfg.h, dt,
                                                                  do_confine ? p->confinement : 0.0f, ts, ib, ie, h->d_bad);
        CKL("k_confine_tile");
    }
    give_plane(h, h->f[FB_U]); give_plane(h, h->f[FB_V]);
    h->f[FB_U] = dU; h->f[FB_V] = dV;
    return FB_OK;
}

// advectVelocity: one kernel, roles swapped instead of copy(f.U, f.newU)
static int advect_velocity_fast(fb_handle *h, float dt, int ext = 0)
{
    TRY(ensure_mask(h));
    float *dU, *dV;
    TRY(take_plane(h, &dU)); TRY(take_plane(h, &dV));
    int ib, ie; range(h, ext, ib, ie);
    const AdvCtx c = adv_ctx(h);
    ADV_VELOCITY_LAUNCH(c, h->f[FB_U], h->f[FB_V], h->mask, h->f[FB_NEWU], h->f[FB_NEWV], dU, dV, dt, ib, ie, h->d_bad);
    give_plane(h, h->f[FB_U]); give_plane(h, h->f[FB_V]);
    h->f[FB_U] = dU; h->f[FB_V] = dV;
    TRY(sync_shadow(h, FB_U, FB_NEWU)); TRY(sync_shadow(h, FB_V, FB_NEWV));
    return FB_OK;
}

// The pass that writes the step's final smoke field `outM` over the owned lines.  With `overlap` (fb_step_local,
// FB_OPT_HALO_OVERLAP) the boundary strips -- the lines the neighbours will pull -- go first, the exchange of M starts on the
// second stream, and the interior follows.
static int halo_overlap_fields(fb_handle *h, cudaEvent_t ev, int field0, int nf, int word, unsigned epoch);
static int halo_exchange_on_second_stream(fb_handle *h, int field0, int nf, int word, unsigned epoch);
static int smoke_final_pass(fb_handle *h, const fb_params *p, const AdvCtx &c, const float *srcM, float *outM, float dt, int ib, int ie,
                            bool overlap)
{
    if (!overlap) {
        ADV_SMOKE_LAUNCH(c, h->f[FB_U], h->f[FB_V], h->mask, srcM, h->f[FB_NEWM], outM, dt,
                         p->smoke_advection, p->viscosity_diffusion, ib, ie, h->d_bad);
        return FB_OK;
    }
    // the strips run on the second (high-priority) stream, next to the interior launch on the handle's stream: two launches
    // of ~130 CTAs in front of the interior cost a wave tail each
    const Grid &g = h->g;
    const int L = h->halo_lines;
    int lo = ib, hi = ie;                      // what is left for the interior launch
    CK(cudaEventRecord(h->ev_ovl_m, h->stream));
    CK(cudaStreamWaitEvent(h->halo_stream, h->ev_ovl_m, 0));
    struct Restore {                           // the launch macros return on error: put the handle back whatever happens
        fb_handle *h; cudaStream_t stream; float *M;
        ~Restore() { h->stream = stream; h->f[FB_M] = M; }
    } restore{h, h->stream, h->f[FB_M]};
    h->stream = h->halo_stream;
    if (h->halo_peer[0] && hi - lo > L) {
        const int e = g.i_lo + L;
        ADV_SMOKE_LAUNCH(c, h->f[FB_U], h->f[FB_V], h->mask, srcM, h->f[FB_NEWM], outM, dt,
                         p->smoke_advection, p->viscosity_diffusion, lo, e, h->d_bad);
        lo = e;
    }
    if (h->halo_peer[1] && hi - lo > L) {
        const int b = g.i_hi - L;
        ADV_SMOKE_LAUNCH(c, h->f[FB_U], h->f[FB_V], h->mask, srcM, h->f[FB_NEWM], outM, dt,
                         p->smoke_advection, p->viscosity_diffusion, b, hi, h->d_bad);
        hi = b;
    }
    h->stream = restore.stream;
    const bool strips_done = (!h->halo_peer[0] || lo > ib) && (!h->halo_peer[1] || hi < ie);
    h->f[FB_M] = outM;                        // what the pack reads and the pull writes
    if (strips_done) TRY(halo_exchange_on_second_stream(h, 2, 1, 2, ++h->ovl_epoch_m));
    if (hi > lo)
        ADV_SMOKE_LAUNCH(c, h->f[FB_U], h->f[FB_V], h->mask, srcM, h->f[FB_NEWM], outM, dt,
                         p->smoke_advection, p->viscosity_diffusion, lo, hi, h->d_bad);
    if (!strips_done)                         // slab thinner than two strips: the whole pass first
        TRY(halo_overlap_fields(h, h->ev_ovl_m, 2, 1, 2, ++h->ovl_epoch_m));
    return FB_OK;
}

static int advect_smoke_fast(fb_handle *h, const fb_params *p, float dt, int ext = 0, bool overlap = false)
{
    TRY(ensure_mask(h));
    float *dM;
    TRY(take_plane(h, &dM));
    int ib, ie; range(h, ext, ib, ie);
    const AdvCtx c = adv_ctx(h);
    TRY(smoke_final_pass(h, p, c, h->f[FB_M], dM, dt, ib, ie, overlap));
    give_plane(h, h->f[FB_M]);
    h->f[FB_M] = dM;
    TRY(sync_shadow(h, FB_M, FB_NEWM));
    return FB_OK;
}

// advectVelocityBFECC in three passes over complete planes (fluid.go:911-994)
static int advect_velocity_bfecc_fast(fb_handle *h, float dt, int ext_fwd = 0, int ext_corr = 0, int ext_final = 0)
{
    TRY(ensure_mask(h));
    float *fU, *fV, *cU, *cV;
    TRY(take_plane(h, &fU)); TRY(take_plane(h, &fV)); TRY(take_plane(h, &cU)); TRY(take_plane(h, &cV));
    int ib, ie; range(h, ext_fwd, ib, ie);
    const AdvCtx c = adv_ctx(h);
    // pass 1: forward advection of the original field
    ADV_VELOCITY_LAUNCH(c, h->f[FB_U], h->f[FB_V], h->mask, h->f[FB_NEWU], h->f[FB_NEWV], fU, fV, dt, ib, ie, h->d_bad);
    // pass 2: back-trace through the original velocities + compensation + clamp
    range(h, ext_corr, ib, ie);
    ADV_BFECC_VELOCITY_LAUNCH(c, h->f[FB_U], h->f[FB_V], h->mask, fU, fV, cU, cV, dt, ib, ie, h->d_bad);
    // pass 3: the corrected field advects itself; skipped faces keep the stale scratch
    // value, which after pass 1 is the forward result there == the old scratch value
    float *oU = h->f[FB_U], *oV = h->f[FB_V];
    range(h, ext_final, ib, ie);
    ADV_VELOCITY_LAUNCH(c, cU, cV, h->mask, h->f[FB_NEWU], h->f[FB_NEWV], oU, oV, dt, ib, ie, h->d_bad);
    give_plane(h, fU); give_plane(h, fV); give_plane(h, cU); give_plane(h, cV);
    TRY(sync_shadow(h, FB_U, FB_NEWU)); TRY(sync_shadow(h, FB_V, FB_NEWV));
    return FB_OK;
}

static int advect_smoke_bfecc_fast(fb_handle *h, const fb_params *p, float dt, int ext_fwd = 0, int ext_corr = 0, bool overlap = false)
{
    TRY(ensure_mask(h));
    float *fM, *cM;
    TRY(take_plane(h, &fM)); TRY(take_plane(h, &cM));
    int ib, ie; range(h, ext_fwd, ib, ie);
    const AdvCtx c = adv_ctx(h);
    ADV_SMOKE_LAUNCH(c, h->f[FB_U], h->f[FB_V], h->mask, h->f[FB_M], h->f[FB_NEWM], fM, dt,
                     p->smoke_advection, p->viscosity_diffusion, ib, ie, h->d_bad);
    range(h, ext_corr, ib, ie);
    ADV_BFECC_SMOKE_LAUNCH(c, h->f[FB_U], h->f[FB_V], h->mask, h->f[FB_M], fM, cM, dt, p->smoke_advection, ib, ie, h->d_bad);
    float *oM = h->f[FB_M];
    range(h, 0, ib, ie);
    TRY(smoke_final_pass(h, p, c, cM, oM, dt, ib, ie, overlap));
    give_plane(h, fM); give_plane(h, cM);
    TRY(sync_shadow(h, FB_M, FB_NEWM));
    return FB_OK;
}

// ---- edits ----------------------------------------------------------------------------
static EditFields edit_fields(fb_handle *h)
{
    EditFields f;
    f.U = h->f[FB_U]; f.V = h->f[FB_V]; f.nU = h->f[FB_NEWU]; f.nV = h->f[FB_NEWV];
    f.P = h->f[FB_P]; f.S = h->f[FB_S]; f.M = h->f[FB_M]; f.nM = h->f[FB_NEWM];
    f.latch_smoke = (!h->literal && !h->exact_shadow) ? 1 : 0;
    return f;
}

static long long cmd_cells(const Grid &g, const fb_edit_cmd &c)
{
    if (c.op == FB_EDIT_CIRCLE_OBSTACLE) return (long long)(2 * c.i1 + 1) * (2 * c.i1 + 1);
    if (c.op == FB_EDIT_RESET) return (long long)g.NX * g.NY;
    long long ni = (long long)c.i1 - c.i0, nj = (long long)c.j1 - c.j0;
    return ni > 0 && nj > 0 ? ni * nj : 0;
}

static int validate_cmd(fb_handle *h, const fb_edit_cmd &c)
{
    const Grid &g = h->g;
    switch (c.op) {
    case FB_EDIT_SET_SOLID: case FB_EDIT_SET_VELOCITY: case FB_EDIT_ADD_SMOKE: case FB_EDIT_SET_SMOKE:
    case FB_EDIT_SET_VELOCITY_IF_FLUID: case FB_EDIT_ADD_SMOKE_IF_FLUID:
        // the reference panics on out-of-range indices (walls.go:6-11, 63-68, 75-80)
        if (c.i0 < 0 || c.j0 < 0 || c.i1 > g.NX || c.j1 > g.NY || c.i1 < c.i0 || c.j1 < c.j0)
            return fail(h, FB_ERR_INVALID, "edit rectangle out of range");
        return FB_OK;
    case FB_EDIT_APPLY_FORCE:      // silently ignores ring / out of range (fluid.go:762-764)
    case FB_EDIT_RESET:
        return FB_OK;
    case FB_EDIT_CIRCLE_OBSTACLE:  // clips to the domain (fluid.go:897-899)
        if (c.i1 < 0) return fail(h, FB_ERR_INVALID, "negative radius");
        return FB_OK;
    default:
        return fail(h, FB_ERR_INVALID, "unknown edit op");
    }
}

// `staged`: the list already sits in h->d_cmds (fb_step stages its per-step list once)
static int run_edits(fb_handle *h, const fb_edit_cmd *cmds, size_t n, bool validate, bool staged = false)
{
    if (n == 0) return FB_OK;
    if (!cmds) return FB_ERR_INVALID;
    if (validate) for (size_t q = 0; q < n; q++) TRY(validate_cmd(h, cmds[q]));
    if (!staged && h->staged_cmds.size() == n && memcmp(h->staged_cmds.data(), cmds, n * sizeof(fb_edit_cmd)) == 0)
        staged = true;                        // the same list is already on the device
    if (!staged) {
        h->staged_cmds.assign(cmds, cmds + n);
        if (n > h->d_cmds_cap) {
            if (h->d_cmds) { CK(cudaStreamSynchronize(h->stream)); CK(cudaFree(h->d_cmds)); h->d_cmds = nullptr; }
            size_t cap = n < 1024 ? 1024 : n * 2;
            CK(cudaMalloc(&h->d_cmds, cap * sizeof(fb_edit_cmd)));
            h->d_cmds_cap = cap;
        }
        // the staging copy must finish before the caller's buffer may change
        CK(cudaMemcpyAsync(h->d_cmds, cmds, n * sizeof(fb_edit_cmd), cudaMemcpyHostToDevice, h->stream));
        CK(cudaStreamSynchronize(h->stream));
    }
    const EditFields f = edit_fields(h);
    for (size_t q = 0; q < n; q++) {
        if (cmds[q].op == FB_EDIT_SET_SOLID || cmds[q].op == FB_EDIT_CIRCLE_OBSTACLE) h->mask_dirty = true;
        if (cmds[q].op == FB_EDIT_RESET) h->p_zero = false;
    }
    const long long SMALL = 1 << 14;
    size_t q = 0;
    while (q < n) {
        if (cmd_cells(h->g, cmds[q]) > SMALL) {
            const fb_edit_cmd &c = cmds[q];
            int nj = c.op == FB_EDIT_CIRCLE_OBSTACLE ? 2 * c.i1 + 1 : (c.op == FB_EDIT_RESET ? h->g.NY : c.j1 - c.j0);
            int ni = c.op == FB_EDIT_CIRCLE_OBSTACLE ? 2 * c.i1 + 1 : (c.op == FB_EDIT_RESET ? h->g.NX : c.i1 - c.i0);
            dim3 grid(cdiv(nj, 256), ni < 4096 ? ni : 4096, 1);
            k_edit_one<<<grid, 256, 0, h->stream>>>(h->g, f, c);
            CKL("k_edit_one");
            q++;
        } else {
            size_t e = q;
            while (e < n && cmd_cells(h->g, cmds[e]) <= SMALL) e++;
            k_edits_seq<<<1, 1024, 0, h->stream>>>(h->g, f, h->d_cmds + q, (int)(e - q));
            CKL("k_edits_seq");
            q = e;
        }
    }
    return FB_OK;
}

extern "C" int fb_edit(fb_handle *h, const fb_edit_cmd *cmds, size_t n)
{
    if (!h) return FB_ERR_INVALID;
    CK(cudaSetDevice(h->device));
    return run_edits(h, cmds, n, true);
}

extern "C" int fb_apply_force_radius(fb_handle *h, int32_t cx, int32_t cy, float fx, float fy, int32_t radius)
{
    if (!h) return FB_ERR_INVALID;
    CK(cudaSetDevice(h->device));
    std::vector<fb_edit_cmd> cmds;
    auto push = [&](int i, int j, float ax, float ay) {
        fb_edit_cmd c; c.op = FB_EDIT_APPLY_FORCE; c.i0 = i; c.j0 = j; c.i1 = i + 1; c.j1 = j + 1; c.a = ax; c.b = ay;
        cmds.push_back(c);
    };
    if (radius <= 0) {
        push(cx, cy, fx, fy);                              // fluid.go:775-778
    } else {
        const float r2 = (float)((long long)radius * radius);
        for (int i = cx - radius; i <= cx + radius; i++)
            for (int j = cy - radius; j <= cy + radius; j++) {
                if (i < 1 || i >= h->g.NX - 1 || j < 1 || j >= h->g.NY - 1) continue;
                volatile float dx = (float)(i - cx), dy = (float)(j - cy);
                volatile float dx2 = dx * dx, dy2 = dy * dy;
                volatile float dist2 = dx2 + dy2;
                if (dist2 > r2) continue;
                volatile float num = -3.0f * dist2;
                volatile float arg = num / r2;
                const float weight = (float)exp((double)arg);   // fluid.go:792
                volatile float wx = fx * weight, wy = fy * weight;
                push(i, j, wx, wy);
            }
    }
    return run_edits(h, cmds.data(), cmds.size(), false);
}

// ---- the step -------------------------------------------------------------------------
static int check_params(fb_handle *h, const fb_params *p)
{
    if (!p) return fail(h, FB_ERR_INVALID, "null params");
    if (p->solver != FB_SOLVER_EXACT && p->solver != FB_SOLVER_REDBLACK && p->solver != FB_SOLVER_REDBLACK_PRESSURE)
        return fail(h, FB_ERR_INVALID, "unknown solver");
    if (p->iters < 0 || p->iters > 32) return fail(h, FB_ERR_INVALID, "iters out of range");
    return FB_OK;
}

extern "C" int fb_step(fb_handle *h, const fb_params *p, float dt, int32_t nsteps,
                       const fb_edit_cmd *per_step, size_t n_per_step)
{
    if (!h) return FB_ERR_INVALID;
    CK(cudaSetDevice(h->device));
    TRY(check_params(h, p));
    if (h->cfg.nranks > 1) return fail(h, FB_ERR_UNSUPPORTED, "fb_step is single-rank; with nranks > 1 the host layer sequences fb_phase calls and halo exchanges");
    if (n_per_step) for (size_t q = 0; q < n_per_step; q++) TRY(validate_cmd(h, per_step[q]));
    const unsigned iters = p->iters > 0 ? (unsigned)p->iters : 8u;
    for (int s = 0; s < nsteps; s++) {
        { ProfScope ps(h, FB_PROF_EDITS); TRY(run_edits(h, per_step, n_per_step, false, s > 0)); }
        if (h->literal) {
            { ProfScope ps(h, FB_PROF_CLEAR_PRESSURE); TRY(clear_pressure(h)); }               // fluid.go:83
            if (p->viscosity_diffusion > 0.0f) { ProfScope ps(h, FB_PROF_VISCOSITY); TRY(apply_viscosity(h, p, dt)); }   // fluid.go:86-88
            { ProfScope ps(h, FB_PROF_PROJECT); TRY(make_incompressible(h, p, dt, iters)); }   // fluid.go:90
            if (p->confinement != 0.0f) { ProfScope ps(h, FB_PROF_CONFINEMENT); TRY(confinement(h, p, dt)); }   // fluid.go:92-94
            if (p->turbulence_strength > 0.0f) { ProfScope ps(h, FB_PROF_TURBULENCE); TRY(turbulence(h, p, dt)); }   // fluid.go:97-99
            { ProfScope ps(h, FB_PROF_BORDERS); TRY(handle_borders(h)); }                      // fluid.go:101
            if (p->use_bfecc) {                                                                // fluid.go:102-108
                { ProfScope ps(h, FB_PROF_ADVECT_VELOCITY); TRY(advect_velocity_bfecc(h, dt)); }
                { ProfScope ps(h, FB_PROF_ADVECT_SMOKE); TRY(advect_smoke_bfecc(h, p, dt)); }
            } else {
                { ProfScope ps(h, FB_PROF_ADVECT_VELOCITY); TRY(advect_velocity(h, dt)); }
                { ProfScope ps(h, FB_PROF_ADVECT_SMOKE); TRY(advect_smoke(h, p, dt)); }
            }
            continue;
        }
        // fused path: same phase order, same arithmetic, fewer passes over HBM
        const bool mg = p->use_multigrid && p->multigrid_levels > 1;   // V-cycle: unfused sweeps that read p
        const bool rb = p->solver != FB_SOLVER_EXACT && !mg;
        const bool conf = p->confinement != 0.0f, turb = p->turbulence_strength > 0.0f;
        if (rb) h->p_zero = true;                       // fill(p, 0) is folded into the fused solve
        else { ProfScope ps(h, FB_PROF_CLEAR_PRESSURE); TRY(clear_pressure(h)); }
        if (p->viscosity_diffusion > 0.0f) { ProfScope ps(h, FB_PROF_VISCOSITY); TRY(apply_viscosity(h, p, dt)); }
        bool turb_done = false;
        {
            ProfScope ps(h, FB_PROF_PROJECT);
            if (rb && iters > 0) {
                TRY(copy_border(h, h->f[FB_NEWU], h->f[FB_U]));
                TRY(copy_border(h, h->f[FB_NEWV], h->f[FB_V]));
                CK(cudaMemsetAsync(h->d_red, 0, 32 * sizeof(unsigned), h->stream));
                h->stats.sweeps_run = (int)iters; h->stats.rolled_back = 0;
                turb_done = turb && !conf && h->fuse_turb;   // turbulence rides on the solve's write-out
                TRY(project_redblack_fused(h, p, dt, iters, turb_done));
            } else {
                if (rb) TRY(clear_pressure(h));
                TRY(make_incompressible(h, p, dt, iters));
            }
        }
        if (conf || (turb && !turb_done)) {
            ProfScope ps(h, conf ? FB_PROF_CONFINEMENT : FB_PROF_TURBULENCE);
            TRY(confine_turbulence_fast(h, p, dt, conf, turb));
        }
        { ProfScope ps(h, FB_PROF_BORDERS); TRY(handle_borders(h)); }
        if (p->use_bfecc) {
            { ProfScope ps(h, FB_PROF_ADVECT_VELOCITY); TRY(advect_velocity_bfecc_fast(h, dt)); }
            { ProfScope ps(h, FB_PROF_ADVECT_SMOKE); TRY(advect_smoke_bfecc_fast(h, p, dt)); }
        } else {
            { ProfScope ps(h, FB_PROF_ADVECT_VELOCITY); TRY(advect_velocity_fast(h, dt)); }
            { ProfScope ps(h, FB_PROF_ADVECT_SMOKE); TRY(advect_smoke_fast(h, p, dt)); }
        }
    }
    return FB_OK;
}

// One Simulate on this rank's slab of a grid split over ranks.  The caller has just
// exchanged halos: the `ghost` lines of U, V and M on each side hold the neighbours'
// current values.  No further communication happens inside the step: every phase is
// recomputed redundantly on as many ghost lines as the phases after it will read
// (`reach` = how many lines a semi-Lagrangian trace may travel, bilinear tap included).
extern "C" int fb_step_local(fb_handle *h, const fb_params *p, float dt, int32_t reach,
                             const fb_edit_cmd *per_step, size_t n_per_step)
{
    if (!h) return FB_ERR_INVALID;
    CK(cudaSetDevice(h->device));
    TRY(check_params(h, p));
    if (p->solver == FB_SOLVER_EXACT && h->cfg.nranks > 1)
        return fail(h, FB_ERR_UNSUPPORTED, "the lexicographic solver does not decompose into slabs; use a red-black solver");
    if (h->literal) return fail(h, FB_ERR_UNSUPPORTED, "slab steps use the fused path");
    if (p->viscosity_diffusion > 0.0f) return fail(h, FB_ERR_UNSUPPORTED, "viscosity is not available in slab steps");
    if (p->use_multigrid && p->multigrid_levels > 1) return fail(h, FB_ERR_UNSUPPORTED, "the multigrid V-cycle is not available in slab steps");
    if (n_per_step) for (size_t q = 0; q < n_per_step; q++) TRY(validate_cmd(h, per_step[q]));
    const unsigned iters = p->iters > 0 ? (unsigned)p->iters : 8u;
    if (iters > 8) return fail(h, FB_ERR_UNSUPPORTED, "slab steps fuse at most 8 iterations (one pass)");
    const bool multi = h->cfg.nranks > 1;
    const int G = h->cfg.ghost, w = reach < 1 ? 1 : reach;
    const bool conf = p->confinement != 0.0f, turb = p->turbulence_strength > 0.0f;
    // extents (ghost lines recomputed per phase), from the end of the step backwards
    int e_smoke_fwd = 0, e_smoke_corr = 0, e_final = 1, e_corr = 0, e_fwd = 0, e_ct;
    if (p->use_bfecc) {
        e_smoke_corr = w; e_smoke_fwd = 2 * w;
        e_final = 2 * w + 1; e_corr = e_final + w; e_fwd = e_corr + w;
        e_ct = e_fwd + w;
    } else {
        e_ct = e_final + w;
    }
    const int e_proj = e_ct + (conf ? 2 : 0);
    if (multi && (e_proj + RQ_H > G || 3 * w > G))
        return fail(h, FB_ERR_HALO, "ghost zone too narrow for this reach; create the handle with more ghost lines");
    if (!multi) { e_smoke_fwd = e_smoke_corr = e_final = e_corr = e_fwd = e_ct = 0; }
    const int ep = multi ? e_proj : 0;

    { ProfScope ps(h, FB_PROF_EDITS); TRY(run_edits(h, per_step, n_per_step, false, false)); }
    h->p_zero = true;
    bool turb_done = false;
    {
        ProfScope ps(h, FB_PROF_PROJECT);
        CK(cudaMemsetAsync(h->d_red, 0, 32 * sizeof(unsigned), h->stream));
        h->stats.sweeps_run = (int)iters; h->stats.rolled_back = 0;
        turb_done = turb && !conf && h->fuse_turb;
        TRY(project_redblack_fused(h, p, dt, iters, turb_done, ep));
    }
    if (conf || (turb && !turb_done)) {
        ProfScope ps(h, conf ? FB_PROF_CONFINEMENT : FB_PROF_TURBULENCE);
        TRY(confine_turbulence_fast(h, p, dt, conf, turb, e_ct));
    }
    { ProfScope ps(h, FB_PROF_BORDERS); TRY(handle_borders(h, e_ct)); }
    const bool ovl = multi && halo_overlap_on(h);
    if (ovl) TRY(halo_overlap_setup(h));
    if (p->use_bfecc) {
        { ProfScope ps(h, FB_PROF_ADVECT_VELOCITY); TRY(advect_velocity_bfecc_fast(h, dt, e_fwd, e_corr, e_final)); }
        if (ovl) TRY(halo_overlap_fields(h, h->ev_ovl_uv, 0, 2, 1, ++h->ovl_epoch_uv));
        { ProfScope ps(h, FB_PROF_ADVECT_SMOKE); TRY(advect_smoke_bfecc_fast(h, p, dt, e_smoke_fwd, e_smoke_corr, ovl)); }
    } else {
        { ProfScope ps(h, FB_PROF_ADVECT_VELOCITY); TRY(advect_velocity_fast(h, dt, e_final)); }
        if (ovl) TRY(halo_overlap_fields(h, h->ev_ovl_uv, 0, 2, 1, ++h->ovl_epoch_uv));
        { ProfScope ps(h, FB_PROF_ADVECT_SMOKE); TRY(advect_smoke_fast(h, p, dt, 0, ovl)); }
    }
    if (ovl) {
        // whatever is queued on the handle's stream next sees the refreshed ghost lines: the next fb_step_local needs no
        // fb_halo_exchange (unless the host edits fields in between)
        ProfScope ps(h, FB_PROF_HALO);
        CK(cudaEventRecord(h->ev_ovl_done, h->halo_stream));
        CK(cudaStreamWaitEvent(h->stream, h->ev_ovl_done, 0));
    }
    return FB_OK;   // ghost-zone violations are latched on the device: fb_check_halo
}

extern "C" int fb_check_halo(fb_handle *h)
{
    if (!h) return FB_ERR_INVALID;
    CK(cudaSetDevice(h->device));
    return check_bad(h);
}

extern "C" int fb_phase(fb_handle *h, int32_t phase, const fb_params *p, float dt, uint32_t iters)
{
    if (!h) return FB_ERR_INVALID;
    CK(cudaSetDevice(h->device));
    TRY(check_params(h, p));
    switch (phase) {
    case FB_PHASE_MAKE_INCOMPRESSIBLE: return make_incompressible(h, p, dt, iters);
    case FB_PHASE_ADVECT_VELOCITY: TRY(h->literal ? advect_velocity(h, dt) : advect_velocity_fast(h, dt)); return check_bad(h);
    case FB_PHASE_ADVECT_SMOKE: TRY(h->literal ? advect_smoke(h, p, dt) : advect_smoke_fast(h, p, dt)); return check_bad(h);
    case FB_PHASE_HANDLE_BORDERS: return handle_borders(h);
    case FB_PHASE_CONFINEMENT: return h->literal ? confinement(h, p, dt) : confine_turbulence_fast(h, p, dt, true, false);
    case FB_PHASE_TURBULENCE: return h->literal ? turbulence(h, p, dt) : confine_turbulence_fast(h, p, dt, false, p->turbulence_strength > 0.0f);
    case FB_PHASE_ADVECT_VELOCITY_BFECC: return h->literal ? advect_velocity_bfecc(h, dt) : advect_velocity_bfecc_fast(h, dt);
    case FB_PHASE_ADVECT_SMOKE_BFECC: return h->literal ? advect_smoke_bfecc(h, p, dt) : advect_smoke_bfecc_fast(h, p, dt);
    case FB_PHASE_VISCOSITY: return apply_viscosity(h, p, dt);
    case FB_PHASE_CLEAR_PRESSURE: return clear_pressure(h);
    case FB_PHASE_PROJECT: {
        ProfScope ps(h, FB_PROF_PROJECT);
        if (h->literal || p->solver == FB_SOLVER_EXACT || iters == 0 || (p->use_multigrid && p->multigrid_levels > 1)) TRY(clear_pressure(h));
        else h->p_zero = true;                         // the fused red-black solvers never read a zero pressure
        return make_incompressible(h, p, dt, iters);
    }
    default: return fail(h, FB_ERR_INVALID, "unknown phase");
    }
}

extern "C" int fb_get_solve_stats(fb_handle *h, fb_solve_stats *out)
{
    if (!h || !out) return FB_ERR_INVALID;
    CK(cudaSetDevice(h->device));
    if (h->stats.rolled_back == 0) TRY(read_stats(h, (unsigned)h->stats.sweeps_run));
    TRY(check_pipeline(h));
    *out = h->stats;
    return FB_OK;
}

// ---- transfers ------------------------------------------------------------------------
static float *field_ptr(fb_handle *h, int field) { return (field >= 0 && field < FB_NFIELDS) ? h->f[field] : nullptr; }

extern "C" int fb_upload(fb_handle *h, int32_t field, const float *host)
{
    if (!h || !host) return FB_ERR_INVALID;
    float *d = field_ptr(h, field);
    if (!d) return fail(h, FB_ERR_INVALID, "bad field");
    CK(cudaSetDevice(h->device));
    const Grid &g = h->g;
    // all lines this rank holds (owned + ghosts) are taken from the global array
    const float *src = host + (size_t)g.i_alloc0 * g.NY;
    CK(cudaMemcpy2DAsync(d, (size_t)g.pitch * 4, src, (size_t)g.NY * 4, (size_t)g.NY * 4, (size_t)g.lines_alloc,
                         cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (field == FB_S) h->mask_dirty = true;
    if (field == FB_P) h->p_zero = false;
    return FB_OK;
}

extern "C" int fb_download(fb_handle *h, int32_t field, float *host)
{
    if (!h || !host) return FB_ERR_INVALID;
    float *d = field_ptr(h, field);
    if (!d) return fail(h, FB_ERR_INVALID, "bad field");
    CK(cudaSetDevice(h->device));
    const Grid &g = h->g;
    float *dst = host + (size_t)g.i_lo * g.NY;
    CK(cudaMemcpy2DAsync(dst, (size_t)g.NY * 4, d + g.at(g.i_lo, 0), (size_t)g.pitch * 4, (size_t)g.NY * 4,
                         (size_t)(g.i_hi - g.i_lo), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return FB_OK;
}

extern "C" int fb_host_mirror(fb_handle *h, int32_t field, float **ptr, size_t *count)
{
    if (!h || !ptr) return FB_ERR_INVALID;
    if (field < 0 || field >= FB_NFIELDS) return fail(h, FB_ERR_INVALID, "bad field");
    CK(cudaSetDevice(h->device));
    const size_t n = (size_t)h->g.NX * h->g.NY;
    if (!h->mirror[field]) {
        CK(cudaMallocHost(&h->mirror[field], n * sizeof(float)));
        memset(h->mirror[field], 0, n * sizeof(float));
    }
    *ptr = h->mirror[field];
    if (count) *count = n;
    return FB_OK;
}

// ---- views and reductions ---------------------------------------------------------------
static float key2f(unsigned k)
{
    unsigned b = (k & 0x80000000u) ? (k ^ 0x80000000u) : ~k;
    float f; memcpy(&f, &b, 4);
    return f;
}

static unsigned f2key_host(float f)
{
    unsigned b; memcpy(&b, &f, 4);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

extern "C" int fb_view(fb_handle *h, int32_t kind, float *out, float *min_value, float *max_value)
{
    if (!h) return FB_ERR_INVALID;
    CK(cudaSetDevice(h->device));
    const Grid &g = h->g;
    // the pipelined forms own the reduction slots and the view scratch until fb_view_end / fb_render_end
    if (h->view_in_flight) return fail(h, FB_ERR_INVALID, "fb_view: a pipelined view is in flight (fb_view_end / fb_render_end first)");
    int ib, ie; range(h, 0, ib, ie);
    dim3 grid, block; plane_launch(g, ib, ie, grid, block);
    // sentinels of pressure.go:6-7 / fluid.go:810-811
    unsigned init[2] = { f2key_host(3.402823466e+38f), f2key_host(-3.402823466e+38f) };
    CK(cudaMemcpyAsync(h->d_red + 32, init, sizeof(init), cudaMemcpyHostToDevice, h->stream));
    const float *src = nullptr;
    switch (kind) {
    case FB_VIEW_SMOKE: src = h->f[FB_M]; break;
    case FB_VIEW_PRESSURE: src = h->f[FB_P]; break;
    case FB_VIEW_VELOCITY_MAGNITUDE: case FB_VIEW_VORTICITY: {
        float *view;
        TRY(scratch(h, SCR_VIEW, &view));
        if (kind == FB_VIEW_VORTICITY)
            k_view<FB_VIEW_VORTICITY><<<grid, block, 0, h->stream>>>(g, h->f[FB_U], h->f[FB_V], h->f[FB_S], view, h->cfg.h, h->d_red + 32, ib, ie);
        else
            k_view<FB_VIEW_VELOCITY_MAGNITUDE><<<grid, block, 0, h->stream>>>(g, h->f[FB_U], h->f[FB_V], h->f[FB_S], view, h->cfg.h, h->d_red + 32, ib, ie);
        CKL("k_view");
        src = view;
        break;
    }
    default: return fail(h, FB_ERR_INVALID, "unknown view");
    }
    if (kind == FB_VIEW_SMOKE || kind == FB_VIEW_PRESSURE) {
        k_minmax_all<<<minmax_blocks(ie - ib), 256, 0, h->stream>>>(g, src, h->d_red + 32, ib, ie);
        CKL("k_minmax_all");
    }
    CK(cudaMemcpyAsync(h->h_red + 32, h->d_red + 32, 2 * sizeof(unsigned), cudaMemcpyDeviceToHost, h->stream));
    if (out) {
        float *dst = out + (size_t)g.i_lo * g.NY;
        CK(cudaMemcpy2DAsync(dst, (size_t)g.NY * 4, src + g.at(g.i_lo, 0), (size_t)g.pitch * 4, (size_t)g.NY * 4,
                             (size_t)(g.i_hi - g.i_lo), cudaMemcpyDeviceToHost, h->stream));
    }
    CK(cudaStreamSynchronize(h->stream));
    if (min_value) *min_value = key2f(h->h_red[32]);
    if (max_value) *max_value = key2f(h->h_red[33]);
    return FB_OK;
}

// Pipelined view: the snapshot and the reduction are queued behind the work already on the
// handle's stream, the transfer runs on a second stream, and the call returns at once.  A frame
// loop calls fb_view_begin(k), queues Simulate k+1, then fb_view_end(k): the PCIe transfer of
// frame k overlaps the computation of frame k+1 (main/main.go draws frame k meanwhile).
extern "C" int fb_view_begin(fb_handle *h, int32_t kind, float *out)
{
    if (!h || !out) return FB_ERR_INVALID;
    if (h->view_in_flight) return fail(h, FB_ERR_INVALID, "fb_view_begin: a view is already in flight");
    CK(cudaSetDevice(h->device));
    const Grid &g = h->g;
    int ib, ie; range(h, 0, ib, ie);
    dim3 grid, block; plane_launch(g, ib, ie, grid, block);
    if (!h->copy_stream) {
        CK(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&h->ev_snap, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&h->ev_view, cudaEventDisableTiming));
    }
    float *snap;
    TRY(scratch(h, SCR_SNAP, &snap));
    k_minmax_init<<<1, 1, 0, h->stream>>>(h->d_red + 32);
    CKL("k_minmax_init");
    const float *src = nullptr;
    int reduce = 1;
    switch (kind) {
    case FB_VIEW_SMOKE: src = h->f[FB_M]; break;
    case FB_VIEW_PRESSURE: src = h->f[FB_P]; break;
    case FB_VIEW_VELOCITY_MAGNITUDE: case FB_VIEW_VORTICITY: {
        float *view;
        TRY(scratch(h, SCR_VIEW, &view));
        if (kind == FB_VIEW_VORTICITY)
            k_view<FB_VIEW_VORTICITY><<<grid, block, 0, h->stream>>>(g, h->f[FB_U], h->f[FB_V], h->f[FB_S], view, h->cfg.h, h->d_red + 32, ib, ie);
        else
            k_view<FB_VIEW_VELOCITY_MAGNITUDE><<<grid, block, 0, h->stream>>>(g, h->f[FB_U], h->f[FB_V], h->f[FB_S], view, h->cfg.h, h->d_red + 32, ib, ie);
        CKL("k_view");
        src = view;
        reduce = 0;          // k_view already reduced over the fluid interior (Q-14)
        break;
    }
    default: return fail(h, FB_ERR_INVALID, "unknown view");
    }
    k_snapshot_minmax<<<grid, block, 0, h->stream>>>(g, src, snap, h->d_red + 32, ib, ie, reduce);
    CKL("k_snapshot_minmax");
    CK(cudaEventRecord(h->ev_snap, h->stream));
    CK(cudaStreamWaitEvent(h->copy_stream, h->ev_snap, 0));
    CK(cudaMemcpyAsync(out + (size_t)ib * g.NY, snap, (size_t)(ie - ib) * g.NY * sizeof(float), cudaMemcpyDeviceToHost,
                       h->copy_stream));
    CK(cudaMemcpyAsync(h->h_red + 32, h->d_red + 32, 2 * sizeof(unsigned), cudaMemcpyDeviceToHost, h->copy_stream));
    CK(cudaEventRecord(h->ev_view, h->copy_stream));
    // the next kernel that touches d_red[32..33] or the snapshot is the next fb_view_begin, after fb_view_end
    h->view_in_flight = true;
    return FB_OK;
}

// Decimated 8-bit pipelined view: the full-field min / max reduction of `kind`, then every stride-th cell of every
// stride-th line quantised against it on the device (k_quantize_u8); what travels is ceil(lines/stride) x ceil(NumY/stride)
// BYTES instead of lines x NumY floats.  Same begin / end protocol and in-flight slot as fb_view_begin.
extern "C" int fb_view_u8_begin(fb_handle *h, int32_t kind, int32_t stride, uint8_t *out, int32_t *out_lines, int32_t *out_cols)
{
    if (!h || !out || stride < 1) return FB_ERR_INVALID;
    if (h->view_in_flight) return fail(h, FB_ERR_INVALID, "fb_view_u8_begin: a view is already in flight");
    CK(cudaSetDevice(h->device));
    const Grid &g = h->g;
    int ib, ie; range(h, 0, ib, ie);
    dim3 grid, block; plane_launch(g, ib, ie, grid, block);
    if (!h->copy_stream) {
        CK(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&h->ev_snap, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&h->ev_view, cudaEventDisableTiming));
    }
    float *snap;
    TRY(scratch(h, SCR_SNAP, &snap));
    k_minmax_init<<<1, 1, 0, h->stream>>>(h->d_red + 32);
    CKL("k_minmax_init");
    const float *src = nullptr;
    switch (kind) {
    case FB_VIEW_SMOKE: src = h->f[FB_M]; break;
    case FB_VIEW_PRESSURE: src = h->f[FB_P]; break;
    case FB_VIEW_VELOCITY_MAGNITUDE: case FB_VIEW_VORTICITY: {
        float *view;
        TRY(scratch(h, SCR_VIEW, &view));
        if (kind == FB_VIEW_VORTICITY)
            k_view<FB_VIEW_VORTICITY><<<grid, block, 0, h->stream>>>(g, h->f[FB_U], h->f[FB_V], h->f[FB_S], view, h->cfg.h, h->d_red + 32, ib, ie);
        else
            k_view<FB_VIEW_VELOCITY_MAGNITUDE><<<grid, block, 0, h->stream>>>(g, h->f[FB_U], h->f[FB_V], h->f[FB_S], view, h->cfg.h, h->d_red + 32, ib, ie);
        CKL("k_view");
        src = view;
        break;
    }
    default: return fail(h, FB_ERR_INVALID, "unknown view");
    }
    if (kind == FB_VIEW_SMOKE || kind == FB_VIEW_PRESSURE) {
        k_minmax_all<<<minmax_blocks(ie - ib), 256, 0, h->stream>>>(g, src, h->d_red + 32, ib, ie);
        CKL("k_minmax_all");
    }
    // owned lines whose global index is a multiple of stride: oi0 .. oi0 + ni - 1 (in units of stride)
    const int oi0 = cdiv(ib, stride), ni = cdiv(ie, stride) - oi0, nj = cdiv(g.NY, stride);
    if (ni > 0) {
        unsigned char *q = reinterpret_cast<unsigned char *>(snap);
        const dim3 qb(32, 8, 1), qg(cdiv(nj, 32), cdiv(ni, 8), 1);
        k_quantize_u8<<<qg, qb, 0, h->stream>>>(g, src, q, h->d_red + 32, stride, oi0, ni, nj);
        CKL("k_quantize_u8");
        CK(cudaEventRecord(h->ev_snap, h->stream));
        CK(cudaStreamWaitEvent(h->copy_stream, h->ev_snap, 0));
        CK(cudaMemcpyAsync(out, q, (size_t)ni * nj, cudaMemcpyDeviceToHost, h->copy_stream));
    } else {
        CK(cudaEventRecord(h->ev_snap, h->stream));
        CK(cudaStreamWaitEvent(h->copy_stream, h->ev_snap, 0));
    }
    CK(cudaMemcpyAsync(h->h_red + 32, h->d_red + 32, 2 * sizeof(unsigned), cudaMemcpyDeviceToHost, h->copy_stream));
    CK(cudaEventRecord(h->ev_view, h->copy_stream));
    if (out_lines) *out_lines = ni > 0 ? ni : 0;
    if (out_cols) *out_cols = nj;
    h->view_in_flight = true;
    return FB_OK;
}

extern "C" int fb_view_end(fb_handle *h, float *min_value, float *max_value)
{
    if (!h) return FB_ERR_INVALID;
    if (!h->view_in_flight) return fail(h, FB_ERR_INVALID, "fb_view_end: no view in flight");
    CK(cudaSetDevice(h->device));
    CK(cudaEventSynchronize(h->ev_view));
    h->view_in_flight = false;
    if (min_value) *min_value = key2f(h->h_red[32]);
    if (max_value) *max_value = key2f(h->h_red[33]);
    return FB_OK;
}

// ---- Draw's pixel pass and advectParticles on the device -----------------
// fb_render_begin / fb_render_end: like fb_view_begin / fb_view_end, but what travels is the RGBA image
// main/main.go:550-574 builds from the view (colormap of main/colors.go, solid cells black), in the
// image layout of fluidToImageIndex (main.go:795): [NumY rows][NumX pixels][4 bytes], row jj showing
// fluid column NumY-1-jj.  `range` = {min, max} to colour with (a multi-GPU host passes the
// all-reduced pair); NULL = this handle's own min / max of the view (Q-14 semantics).
extern "C" int fb_render_begin(fb_handle *h, int32_t kind, uint8_t *rgba_out, const float *color_range)
{
    if (!h || !rgba_out) return FB_ERR_INVALID;
    if (h->view_in_flight) return fail(h, FB_ERR_INVALID, "fb_render_begin: a view is already in flight");
    CK(cudaSetDevice(h->device));
    const Grid &g = h->g;
    int ib, ie; range(h, 0, ib, ie);
    dim3 grid, block; plane_launch(g, ib, ie, grid, block);
    if (!h->copy_stream) {
        CK(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&h->ev_snap, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&h->ev_view, cudaEventDisableTiming));
    }
    float *snap;
    TRY(scratch(h, SCR_SNAP, &snap));          // the image: (ie - ib) * NumY pixels of 4 bytes, a plane holds them
    k_minmax_init<<<1, 1, 0, h->stream>>>(h->d_red + 32);
    CKL("k_minmax_init");
    const float *src = nullptr;
    switch (kind) {
    case FB_VIEW_SMOKE: src = h->f[FB_M]; break;
    case FB_VIEW_PRESSURE: src = h->f[FB_P]; break;
    case FB_VIEW_VELOCITY_MAGNITUDE: case FB_VIEW_VORTICITY: {
        float *view;
        TRY(scratch(h, SCR_VIEW, &view));
        if (kind == FB_VIEW_VORTICITY)
            k_view<FB_VIEW_VORTICITY><<<grid, block, 0, h->stream>>>(g, h->f[FB_U], h->f[FB_V], h->f[FB_S], view, h->cfg.h, h->d_red + 32, ib, ie);
        else
            k_view<FB_VIEW_VELOCITY_MAGNITUDE><<<grid, block, 0, h->stream>>>(g, h->f[FB_U], h->f[FB_V], h->f[FB_S], view, h->cfg.h, h->d_red + 32, ib, ie);
        CKL("k_view");
        src = view;
        break;
    }
    default: return fail(h, FB_ERR_INVALID, "unknown view");
    }
    if (kind == FB_VIEW_SMOKE || kind == FB_VIEW_PRESSURE) {
        k_minmax_all<<<minmax_blocks(ie - ib), 256, 0, h->stream>>>(g, src, h->d_red + 32, ib, ie);
        CKL("k_minmax_all");
    }
    if (color_range) {
        k_minmax_set<<<1, 1, 0, h->stream>>>(h->d_red + 32, color_range[0], color_range[1]);
        CKL("k_minmax_set");
    }
    const dim3 rgrid(cdiv(g.NY, 32), cdiv(ie - ib, 32), 1), rblock(32, 8, 1);
    unsigned *img = reinterpret_cast<unsigned *>(snap);
    if (kind == FB_VIEW_VORTICITY) k_render<1><<<rgrid, rblock, 0, h->stream>>>(g, src, h->f[FB_S], h->d_red + 32, img, ib, ie);
    else k_render<0><<<rgrid, rblock, 0, h->stream>>>(g, src, h->f[FB_S], h->d_red + 32, img, ib, ie);
    CKL("k_render");
    CK(cudaEventRecord(h->ev_snap, h->stream));
    CK(cudaStreamWaitEvent(h->copy_stream, h->ev_snap, 0));
    // this rank's pixel columns [ib, ie) of every image row
    if (ie - ib == g.NX)      // whole image: one contiguous transfer (the 2-D form runs at ~85 % of it)
        CK(cudaMemcpyAsync(rgba_out, img, (size_t)g.NX * g.NY * 4, cudaMemcpyDeviceToHost, h->copy_stream));
    else
        CK(cudaMemcpy2DAsync(rgba_out + (size_t)ib * 4, (size_t)g.NX * 4, img, (size_t)(ie - ib) * 4, (size_t)(ie - ib) * 4,
                             (size_t)g.NY, cudaMemcpyDeviceToHost, h->copy_stream));
    CK(cudaMemcpyAsync(h->h_red + 32, h->d_red + 32, 2 * sizeof(unsigned), cudaMemcpyDeviceToHost, h->copy_stream));
    CK(cudaEventRecord(h->ev_view, h->copy_stream));
    h->view_in_flight = true;
    return FB_OK;
}

extern "C" int fb_render_end(fb_handle *h, float *min_value, float *max_value) { return fb_view_end(h, min_value, max_value); }

extern "C" int fb_render(fb_handle *h, int32_t kind, uint8_t *rgba_out, const float *color_range, float *min_value, float *max_value)
{
    TRY(fb_render_begin(h, kind, rgba_out, color_range));
    return fb_render_end(h, min_value, max_value);
}

// advectParticles (main/main.go:512-546) for n particles in host memory, in place; survivors keep their
// order (the reference's `alive = append(alive, *p)`), *n_alive says how many.  The four bilinear samples
// per particle run on the device; the stable filter is a host pass over the flags that came back.
extern "C" int fb_advect_particles(fb_handle *h, fb_particle *particles, size_t n, float dt, size_t *n_alive)
{
    if (!h || (n && !particles) || !n_alive) return FB_ERR_INVALID;
    static_assert(sizeof(fb_particle) == sizeof(fb_particle_dev), "particle layouts must match");
    *n_alive = 0;
    if (n == 0) return FB_OK;
    CK(cudaSetDevice(h->device));
    if (h->cfg.nranks > 1) return fail(h, FB_ERR_UNSUPPORTED, "fb_advect_particles: particles roam the whole grid; single-GPU handles only");
    if (n > h->particles_cap) {
        if (h->d_particles) cudaFree(h->d_particles);
        if (h->d_alive) cudaFree(h->d_alive);
        h->d_particles = nullptr; h->d_alive = nullptr; h->particles_cap = 0;
        const size_t cap = n + n / 2 + 1024;
        CK(cudaMalloc(&h->d_particles, cap * sizeof(fb_particle_dev)));
        CK(cudaMalloc(&h->d_alive, cap * sizeof(int)));
        h->particles_cap = cap;
    }
    h->h_alive.resize(n);
    CK(cudaMemcpyAsync(h->d_particles, particles, n * sizeof(fb_particle), cudaMemcpyHostToDevice, h->stream));
    k_advect_particles<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(h->g, h->f[FB_U], h->f[FB_V], h->f[FB_S], h->d_particles,
                                                                          h->d_alive, n, dt, h->cfg.h, h->d_bad);
    CKL("k_advect_particles");
    CK(cudaMemcpyAsync(particles, h->d_particles, n * sizeof(fb_particle), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(h->h_alive.data(), h->d_alive, n * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    size_t a = 0;
    for (size_t k = 0; k < n; k++)
        if (h->h_alive[k]) { if (a != k) particles[a] = particles[k]; a++; }
    *n_alive = a;
    return FB_OK;
}

extern "C" int fb_reduce(fb_handle *h, int32_t kind, float *out)
{
    if (!h || !out) return FB_ERR_INVALID;
    CK(cudaSetDevice(h->device));
    int ib, ie; range(h, 0, ib, ie);
    dim3 grid, block; plane_launch(h->g, ib, ie, grid, block);
    CK(cudaMemsetAsync(h->d_red + 40, 0, sizeof(unsigned), h->stream));
    if (kind == FB_REDUCE_MAX_DIVERGENCE)
        k_reduce_max<FB_REDUCE_MAX_DIVERGENCE><<<grid, block, 0, h->stream>>>(h->g, h->f[FB_U], h->f[FB_V], h->f[FB_S], h->d_red + 40, ib, ie);
    else if (kind == FB_REDUCE_MAX_ABS_VELOCITY)
        k_reduce_max<FB_REDUCE_MAX_ABS_VELOCITY><<<grid, block, 0, h->stream>>>(h->g, h->f[FB_U], h->f[FB_V], h->f[FB_S], h->d_red + 40, ib, ie);
    else return fail(h, FB_ERR_INVALID, "unknown reduction");
    CKL("k_reduce_max");
    CK(cudaMemcpyAsync(h->h_red + 40, h->d_red + 40, sizeof(unsigned), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    memcpy(out, h->h_red + 40, 4);
    return FB_OK;
}

extern "C" int fb_sample_velocity(fb_handle *h, size_t n, const float *xy, float *uv)
{
    if (!h || (n && (!xy || !uv))) return FB_ERR_INVALID;
    if (n == 0) return FB_OK;
    CK(cudaSetDevice(h->device));
    float *dxy = nullptr, *duv = nullptr;
    CK(cudaMalloc(&dxy, n * 2 * sizeof(float)));
    cudaError_t e = cudaMalloc(&duv, n * 2 * sizeof(float));
    if (e != cudaSuccess) { cudaFree(dxy); return fail(h, FB_ERR_CUDA, "cudaMalloc", e); }
    int rc = FB_OK;
    do {
        if ((e = cudaMemcpyAsync(dxy, xy, n * 2 * sizeof(float), cudaMemcpyHostToDevice, h->stream)) != cudaSuccess) break;
        k_sample_velocity<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(h->g, h->f[FB_U], h->f[FB_V], dxy, duv, n, h->cfg.h, h->d_bad);
        h->launches++;
        if ((e = cudaGetLastError()) != cudaSuccess) break;
        if ((e = cudaMemcpyAsync(uv, duv, n * 2 * sizeof(float), cudaMemcpyDeviceToHost, h->stream)) != cudaSuccess) break;
        e = cudaStreamSynchronize(h->stream);
    } while (0);
    if (e != cudaSuccess) rc = fail(h, FB_ERR_CUDA, "fb_sample_velocity", e);
    cudaFree(dxy); cudaFree(duv);
    return rc;
}

// ---- halo regions -----------------------------------------------------------------------
extern "C" int fb_halo_region(fb_handle *h, int32_t field, int32_t side, int32_t lines,
                              void **send_ptr, void **recv_ptr, size_t *bytes)
{
    if (!h) return FB_ERR_INVALID;
    float *d = field_ptr(h, field);
    if (!d) return fail(h, FB_ERR_INVALID, "bad field");
    const Grid &g = h->g;
    if (lines < 1 || lines > h->cfg.ghost) return fail(h, FB_ERR_INVALID, "halo wider than the ghost zone");
    if (lines > g.i_hi - g.i_lo) return fail(h, FB_ERR_INVALID, "halo wider than the slab");
    int send_i, recv_i;
    if (side == 0) { send_i = g.i_lo; recv_i = g.i_lo - lines; }
    else if (side == 1) { send_i = g.i_hi - lines; recv_i = g.i_hi; }
    else return fail(h, FB_ERR_INVALID, "side must be 0 or 1");
    if (recv_i < g.i_alloc0 || recv_i + lines > g.i_alloc0 + g.lines_alloc)
        return fail(h, FB_ERR_INVALID, "no neighbour on that side");
    if (send_ptr) *send_ptr = d + g.at(send_i, 0);
    if (recv_ptr) *recv_ptr = d + g.at(recv_i, 0);
    if (bytes) *bytes = (size_t)lines * g.pitch * sizeof(float);
    return FB_OK;
}
