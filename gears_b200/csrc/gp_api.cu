// gp_api.cu — host side of libgears_physics_cuda.so: the C ABI declared in include/gears_physics_cuda.h.
// Owns all device memory, enqueues the kernels of gp_kernels.cuh on one non-blocking stream per handle.
// No CPU fallback: every entry point fails with GP_ERR_NO_DEVICE / GP_ERR_CUDA if the device path is unusable.
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/gears_physics_cuda.h"
#include "gp_kernels.cuh"
#include "gp_multi.cuh"
#include "radix_sort.cuh"
#include "scan.cuh"

using namespace gp;

static std::string g_create_error;

struct gp_world {
    gp_config cfg{};
    int device = 0;
    int sm_count = 148;
    cudaStream_t st = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool timed = false;

    uint32_t n = 0, cap = 0;
    uint64_t pair_cap = 0;
    bool uploaded = false;
    bool dead = false;  // a slab neighbour timed out: no further steps until gp_upload (GP_ERR_PEER_TIMEOUT)

    // body state: two sets of five float4 arrays; every step gathers set `cur` into set `cur^1` in cell order
    float4 *B[2] = {nullptr, nullptr};  // 64 B records: P, V, Lo, Hi (see struct Bodies)
    float4 *A[2] = {nullptr, nullptr};  // acc | flags
    int cur = 0;
    float* ext = nullptr;
    uint32_t* slot_of = nullptr;  // body id -> current slot (rebuilt on demand: slots change every step)
    float4* rawbox = nullptr;        // per body id: the un-rotated collision box (ray cast, obstacle grid)
    uint32_t* entity_dev = nullptr;  // caller's entity ids as uploaded (become the body ids of a slab world)
    bool have_entity = false;
    bool slot_valid = false;
    bool track_slots = false;  // K3 keeps slot_of current (set by the first per-index update: the host tracks a dirty set)
    // broad phase
    float4* sbox = nullptr;       // world AABBs of the current step in cell order (lo|tag, hi|rank)
    uint32_t *key = nullptr, *lrank = nullptr;  // per input slot: cell key, arrival rank within the cell
    uint32_t* perm = nullptr;                   // per output slot: the input slot it came from
    uint32_t* LB = nullptr;                     // cell histogram, scanned in place into cell start offsets
    uint32_t lb_cap = 0;
    uint32_t* lb_tiles = nullptr;               // scan scratch
    uint32_t max_cells = 0;
    int key_bits = 1;
    float thr = 0.f;
    // chains + sweep
    uint32_t *chain_cnt = nullptr, *offs = nullptr, *tile_sums = nullptr;
    uint4* pairs = nullptr;
    unsigned long long* adj = nullptr;
    uint32_t* frontier[2] = {nullptr, nullptr};
    uint32_t* heavy = nullptr;
    uint32_t heavy_cap = 0;
    float* dmax = nullptr;        // per body: largest displacement during the current sweep
    float* reach = nullptr;       // per body: largest displacement over the earlier sweeps of this step
    uint32_t *moved_bits = nullptr, *esc_bits = nullptr;  // per body: moved at all / left the outer margin
    uint2* watch = nullptr;       // lazy candidates: pairs between the inner and the outer margin
    // component-local repair scratch
    uint32_t *dirty_bits = nullptr, *dlist = nullptr, *dpairs = nullptr, *offs2 = nullptr, *deg2 = nullptr;
    unsigned long long* adj2 = nullptr;
    uint32_t* dheavy = nullptr;
    uint32_t *xhead = nullptr, *xnext = nullptr;  // per-body lists of the pairs appended after K4 (tiny repair)
    bool tiny_repair = false;  // per-component warp repair in front of the cooperative one (set at upload)
    uint32_t* escapees = nullptr;  // bodies that left their margin
    int sweep_grid = 0, fp_grid = 0;
    // control
    Ctl *ctl = nullptr, *ctl_snap = nullptr;
    Ctl* h_ctl = nullptr;  // pinned
    uint32_t *d_small = nullptr;  // scratch: ext histogram (EXT_BUCKETS) + bounds (6)
    // staging for host-layout arrays
    unsigned char* stage = nullptr;
    size_t stage_bytes = 0;
    // pipelined frames (gp_frame_submit): copy streams, double-buffered device staging, body constants in id order
    cudaStream_t st_h2d = nullptr, st_d2h = nullptr;
    float4* home = nullptr;
    bool home_valid = false;
    float* fin[2] = {nullptr, nullptr};      // 3 x (n x 3) floats: pos, vel, acc
    float4* fout8 = nullptr;                 // n x 8 floats: whole-sector rows the scattered read-back writes
    float* fout[2] = {nullptr, nullptr};     // n x 6 floats: what goes to the host
    uint32_t frame_cap = 0;
    cudaEvent_t ev_h2d[2] = {nullptr, nullptr}, ev_in_free[2] = {nullptr, nullptr}, ev_unpacked[2] = {nullptr, nullptr},
                ev_d2h[2] = {nullptr, nullptr};
    uint64_t frames_submitted = 0;

    // profiling (GP_FLAG_PROFILE): events between kernel groups
    static constexpr int PROF_STEPS = 512;
    std::vector<cudaEvent_t> prof_ev;  // (GP_NUM_KERNEL_GROUPS + 1) per step
    uint32_t prof_n = 0;               // steps recorded since the last read
    uint32_t prof_launch[GP_NUM_KERNEL_GROUPS] = {};
    uint64_t prof_mark_launches = 0;
    int prof_group = -1;

    // GP_PROF_FINE=1 (diagnostics): an event after every launch of a step; mean time per kernel printed at gp_destroy
    bool fine = false;
    std::vector<std::pair<const char*, cudaEvent_t>> fine_ev;
    size_t fine_used = 0;
    std::vector<std::pair<std::string, std::pair<double, uint64_t>>> fine_acc;
    uint64_t fine_steps = 0;
    bool fine_reset_done = false;

    gp_step_stats last{};
    uint64_t steps_total = 0;
    uint32_t last_np = 0;
    std::string err;
    uint64_t launches = 0;
    gp_multi* multi = nullptr;
};

// ---- fine-grained timing (GP_PROF_FINE=1): everything enqueued since the previous mark is charged to `name` ----
static void fine_mark(gp_world* w, const char* name) {
    if (!w->fine || w->fine_used >= (1u << 18)) return;
    if (w->fine_used == w->fine_ev.size()) {
        cudaEvent_t e;
        cudaEventCreate(&e);
        w->fine_ev.push_back({name, e});
    }
    w->fine_ev[w->fine_used].first = name;
    cudaEventRecord(w->fine_ev[w->fine_used].second, w->st);
    w->fine_used++;
}
static void fine_collect(gp_world* w) {  // the stream is idle
    for (size_t i = 1; i < w->fine_used; ++i) {
        const char* nm = w->fine_ev[i].first;
        if (!strcmp(nm, "begin")) {
            w->fine_steps++;
            continue;
        }
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, w->fine_ev[i - 1].second, w->fine_ev[i].second) != cudaSuccess) continue;
        auto it = std::find_if(w->fine_acc.begin(), w->fine_acc.end(), [&](const auto& a) { return a.first == nm; });
        if (it == w->fine_acc.end()) w->fine_acc.push_back({nm, {0.0, 0}}), it = w->fine_acc.end() - 1;
        it->second.first += ms, it->second.second++;
    }
    if (w->fine_used && !strcmp(w->fine_ev[0].first, "begin")) w->fine_steps++;
    w->fine_used = 0;
}
static void fine_print(gp_world* w) {
    if (!w->fine || !w->fine_steps) return;
    double tot = 0;
    for (auto& a : w->fine_acc) tot += a.second.first;
    fprintf(stderr, "[gp fine] %llu steps, %.1f us/step\n", (unsigned long long)w->fine_steps, 1e3 * tot / w->fine_steps);
    for (auto& a : w->fine_acc)
        fprintf(stderr, "[gp fine] %-22s %8.1f us/step  (%.2f launches/step)\n", a.first.c_str(),
                1e3 * a.second.first / w->fine_steps, (double)a.second.second / w->fine_steps);
    // phases inside k_contacts, as seen by thread 0 of the grid (globaltimer)
    Ctl c;
    if (cudaMemcpy(&c, w->ctl, sizeof c, cudaMemcpyDeviceToHost) != cudaSuccess) return;
    static const char* names[12] = {"sweep: first frontier", "sweep: grid levels", "sweep: tail levels (CTA 0)",
                                    "sweep: stats",          "verify: requery",    "verify: watch list",
                                    "repair: link+walk",     "repair: exec",       "repair: list sweep",
                                    "repair: component-local", "repair: re-verify", "slab proof"};
    for (int i = 0; i < 12; ++i)
        if (c.dbg_phase_ns[i])
            fprintf(stderr, "[gp fine]   k_contacts / %-26s %8.1f us/step\n", names[i],
                    1e-3 * (double)c.dbg_phase_ns[i] / w->fine_steps);
    fprintf(stderr, "[gp fine]   k_contacts / levels run by CTA 0 alone: %.2f per step\n",
            (double)c.dbg_tail_rounds / w->fine_steps);
    for (int i = 0; i < 16; ++i)
        if (c.dbg_level_ns[i])
            fprintf(stderr, "[gp fine]   first sweep, slot %2d: %8.1f us/step, %10.1f pairs/step%s\n", i,
                    1e-3 * (double)c.dbg_level_ns[i] / w->fine_steps, (double)c.dbg_level_n[i] / w->fine_steps,
                    i == 0 ? "  (the scan that lists level 0)" : "");
    fprintf(stderr, "[gp fine]   per step: %.1f escapees, %.1f requeried pairs, %.1f activated watch pairs, %.2f repair rounds; "
            "per-component repair: %u rounds ok, %u fell back, largest component %u bodies / %u pairs\n",
            (double)c.dbg_counts[0] / w->fine_steps, (double)c.dbg_counts[1] / w->fine_steps,
            (double)c.dbg_counts[2] / w->fine_steps, (double)c.dbg_counts[3] / w->fine_steps, c.dbg_tiny[0], c.dbg_tiny[1], c.dbg_tiny[4], c.dbg_tiny[5]);
}

static gp_status fail(gp_world* w, gp_status s, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (w)
        w->err = buf;
    else
        g_create_error = buf;
    return s;
}

#define CK(call)                                                                                            \
    do {                                                                                                    \
        cudaError_t e_ = (call);                                                                            \
        if (e_ != cudaSuccess)                                                                              \
            return fail(w, GP_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

template <typename T>
static cudaError_t dmalloc(T** p, size_t count) {
    if (*p) {
        cudaFree(*p);
        *p = nullptr;
    }
    if (count == 0) count = 1;
    return cudaMalloc((void**)p, count * sizeof(T));
}

static inline uint32_t cdiv(uint64_t a, uint32_t b) { return (uint32_t)((a + b - 1) / b); }

static void free_frames(gp_world* w) {
    for (int b = 0; b < 2; ++b) {
        if (w->fin[b]) cudaFree(w->fin[b]), w->fin[b] = nullptr;
        if (w->fout[b]) cudaFree(w->fout[b]), w->fout[b] = nullptr;
        if (b == 0 && w->fout8) cudaFree(w->fout8), w->fout8 = nullptr;
        for (cudaEvent_t* e : {&w->ev_h2d[b], &w->ev_in_free[b], &w->ev_unpacked[b], &w->ev_d2h[b]})
            if (*e) cudaEventDestroy(*e), *e = nullptr;
    }
    if (w->home) cudaFree(w->home), w->home = nullptr;
    if (w->st_h2d) cudaStreamDestroy(w->st_h2d), w->st_h2d = nullptr;
    if (w->st_d2h) cudaStreamDestroy(w->st_d2h), w->st_d2h = nullptr;
    w->frame_cap = 0, w->home_valid = false;
}

static void free_all(gp_world* w) {
    free_frames(w);
    void* ptrs[] = {w->B[0],   w->B[1],  w->A[0],  w->A[1], w->ext,        w->slot_of, w->entity_dev, w->rawbox,
                    w->sbox,  w->key,     w->lrank,     w->perm, w->lb_tiles,   w->LB,
                    w->chain_cnt, w->offs,   w->tile_sums,  w->pairs, w->adj,
                    w->frontier[0], w->frontier[1], w->heavy, w->dmax, w->reach, w->escapees, w->moved_bits, w->esc_bits, w->watch, w->dirty_bits, w->dlist, w->dpairs, w->offs2, w->deg2, w->adj2, w->dheavy, w->xhead, w->xnext, w->ctl, w->ctl_snap, w->d_small, w->stage};
    for (void* p : ptrs)
        if (p) cudaFree(p);
    if (w->h_ctl) cudaFreeHost(w->h_ctl);
}

extern "C" int32_t gp_abi_version(void) { return GP_ABI_VERSION; }

extern "C" const char* gp_last_error(const gp_world* w) { return w ? w->err.c_str() : g_create_error.c_str(); }

extern "C" void gp_damping_factors(float dt, float* out3) {
    // platform libm expf == what Rust's f32::exp lowers to on linux-gnu (physics.rs:437)
    out3[0] = expf(-1.6f * dt);
    out3[1] = expf(-1.8f * dt);
    out3[2] = expf(-3.5f * dt);
}

extern "C" void* gp_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) return nullptr;
    return p;
}
extern "C" void gp_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

extern "C" uint64_t gp_kernel_launches(const gp_world* w) { return w ? w->launches : 0; }

extern "C" gp_status gp_create(const gp_config* cfg, gp_world** out) {
    if (!out) return fail(nullptr, GP_ERR_BAD_ARG, "gp_create: out is NULL");
    *out = nullptr;
    gp_config c{};
    c.struct_size = sizeof(gp_config);
    if (cfg) {
        if (cfg->struct_size == 0 || cfg->struct_size > sizeof(gp_config))
            return fail(nullptr, GP_ERR_BAD_ARG, "gp_create: bad gp_config.struct_size %u", cfg->struct_size);
        memcpy(&c, cfg, cfg->struct_size);
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(nullptr, GP_ERR_NO_DEVICE, "gp_create: no CUDA device (this library has no CPU path)");
    if (c.device < 0 || c.device >= ndev) return fail(nullptr, GP_ERR_BAD_ARG, "gp_create: device %d of %d", c.device, ndev);
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, c.device) != cudaSuccess || prop.major != 10)
        return fail(nullptr, GP_ERR_NO_DEVICE, "gp_create: device %d is sm_%d%d; kernels are built for sm_100a only",
                    c.device, prop.major, prop.minor);
    gp_world* w = new gp_world();
    w->cfg = c;
    w->device = c.device;
    w->sm_count = prop.multiProcessorCount;
    gp_status rs = GP_OK;
    do {
        if (cudaSetDevice(w->device) != cudaSuccess) { rs = GP_ERR_CUDA; break; }
        if (cudaStreamCreateWithFlags(&w->st, cudaStreamNonBlocking) != cudaSuccess) { rs = GP_ERR_CUDA; break; }
        if (cudaEventCreate(&w->ev0) != cudaSuccess || cudaEventCreate(&w->ev1) != cudaSuccess) { rs = GP_ERR_CUDA; break; }
            if (dmalloc(&w->ctl, 1) || dmalloc(&w->ctl_snap, 1) || dmalloc(&w->d_small, 2 * EXT_BUCKETS + 8)) { rs = GP_ERR_CUDA; break; }
        if (cudaMallocHost((void**)&w->h_ctl, sizeof(Ctl)) != cudaSuccess) { rs = GP_ERR_CUDA; break; }
        if (cudaMemset(w->ctl, 0, sizeof(Ctl)) != cudaSuccess) { rs = GP_ERR_CUDA; break; }
        int occ = 0;
        if (cudaFuncSetAttribute(k_contacts, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ContactSmem)) !=
                cudaSuccess ||
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_contacts, COOP_THREADS, sizeof(ContactSmem)) !=
                cudaSuccess ||
            occ < 1) {
            rs = GP_ERR_NO_DEVICE;  // no kernel image for this device
            break;
        }
        w->sweep_grid = w->sm_count;  // one CTA per SM: a grid barrier costs per CTA
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_find_pairs, WP_THREADS, 0) != cudaSuccess || occ < 1) {
            rs = GP_ERR_NO_DEVICE;
            break;
        }
        w->fp_grid = occ * w->sm_count;  // persistent warps: every SM slot filled once
        {
            const char* e = getenv("GP_PROF_FINE");
            w->fine = e && atoi(e) != 0;
        }
    } while (0);
    if (rs != GP_OK) {
        fail(nullptr, rs, "gp_create: CUDA setup failed: %s", cudaGetErrorString(cudaGetLastError()));
        free_all(w);
        delete w;
        return rs;
    }
    *out = w;
    return GP_OK;
}

extern "C" void gp_destroy(gp_world* w) {
    if (!w) return;
    cudaSetDevice(w->device);
    if (w->st) cudaStreamSynchronize(w->st);
    if (w->fine) fine_collect(w), fine_print(w);
    for (auto& e : w->fine_ev) cudaEventDestroy(e.second);
    if (w->multi) gp_multi_destroy(w->multi);
    free_all(w);
    for (auto& e : w->prof_ev) cudaEventDestroy(e);
    if (w->ev0) cudaEventDestroy(w->ev0);
    if (w->ev1) cudaEventDestroy(w->ev1);
    if (w->st) cudaStreamDestroy(w->st);
    delete w;
}

static gp_status ensure_stage(gp_world* w, size_t bytes) {
    if (bytes <= w->stage_bytes) return GP_OK;
    CK(dmalloc(&w->stage, bytes));
    w->stage_bytes = bytes;
    return GP_OK;
}

static gp_status alloc_bodies(gp_world* w, uint32_t cap) {
    for (int i = 0; i < 2; ++i) {
        CK(dmalloc(&w->B[i], 4 * (size_t)cap));
        CK(dmalloc(&w->A[i], cap));
    }
    CK(dmalloc(&w->ext, cap));
    CK(dmalloc(&w->slot_of, cap));
    CK(dmalloc(&w->entity_dev, cap));
    CK(dmalloc(&w->rawbox, 2 * (size_t)cap));
    CK(dmalloc(&w->sbox, 2 * (size_t)cap));
    CK(dmalloc(&w->key, cap));
    CK(dmalloc(&w->lrank, cap));
    CK(dmalloc(&w->perm, cap));
    CK(dmalloc(&w->chain_cnt, (size_t)cap + 1));
    CK(dmalloc(&w->offs, (size_t)cap + 2));
    CK(dmalloc(&w->tile_sums, (size_t)cdiv(cap, SC_TILE) + 1));
    CK(dmalloc(&w->dmax, cap));
    CK(dmalloc(&w->reach, cap));
    CK(cudaMemset(w->reach, 0, sizeof(float) * (size_t)cap));
    CK(dmalloc(&w->escapees, cap));
    CK(dmalloc(&w->moved_bits, (size_t)cap / 32 + 1));
    CK(dmalloc(&w->esc_bits, (size_t)cap / 32 + 1));
    CK(dmalloc(&w->dirty_bits, (size_t)cap / 32 + 1));
    CK(cudaMemset(w->dirty_bits, 0, sizeof(uint32_t) * ((size_t)cap / 32 + 1)));
    CK(dmalloc(&w->dlist, cap));
    CK(dmalloc(&w->offs2, cap));
    CK(dmalloc(&w->deg2, cap));
    CK(dmalloc(&w->xhead, cap));
    CK(cudaMemset(w->dmax, 0, sizeof(float) * (size_t)cap));
    w->cap = cap;
    return GP_OK;
}

static gp_status alloc_pairs(gp_world* w, uint64_t pair_cap) {
    if (pair_cap > 0x7fff0000ull) pair_cap = 0x7fff0000ull;
    CK(dmalloc(&w->pairs, 2 * pair_cap));  // 32 B records
    CK(dmalloc(&w->adj, 2 * pair_cap));
    CK(dmalloc(&w->watch, pair_cap));
    CK(dmalloc(&w->dpairs, pair_cap));
    CK(dmalloc(&w->xnext, 2 * pair_cap));
    CK(dmalloc(&w->adj2, 2 * pair_cap));
    CK(dmalloc(&w->frontier[0], pair_cap));
    CK(dmalloc(&w->frontier[1], pair_cap));
    w->heavy_cap = (uint32_t)(2 * pair_cap / CHAIN_INLINE + 16);  // a heavy chain has > CHAIN_INLINE of the 2*pair_cap entries
    CK(dmalloc(&w->heavy, w->heavy_cap));
    CK(dmalloc(&w->dheavy, w->heavy_cap));
    w->pair_cap = pair_cap;
    return GP_OK;
}

static bool user_bounds_degenerate(const float* lo, const float* hi) {
    for (int k = 0; k < 3; ++k)
        if (!(hi[k] > lo[k])) return true;
    return false;
}

// Choose the extent threshold / cell size and the grid from what was uploaded.
static gp_status setup_grid(gp_world* w) {
    const uint32_t n = w->n;
    uint32_t* d_hist = w->d_small;
    uint32_t* d_bounds = w->d_small + 2 * EXT_BUCKETS;
    uint32_t h_small[2 * EXT_BUCKETS + 8];
    CK(cudaMemsetAsync(d_hist, 0, 2 * EXT_BUCKETS * sizeof(uint32_t), w->st));
    if (n) {
        k_ext_hist<<<std::min<uint32_t>(cdiv(n, 256), 1024), 256, 0, w->st>>>(n, w->ext, w->A[w->cur], d_hist);
        w->launches++;
    }
    CK(cudaMemcpyAsync(h_small, d_hist, 2 * EXT_BUCKETS * sizeof(uint32_t), cudaMemcpyDeviceToHost, w->st));
    CK(cudaStreamSynchronize(w->st));
    const uint32_t max_big = w->cfg.max_big ? w->cfg.max_big : 64u;
    // smallest bucket edge that leaves at most max_big bodies above it
    uint64_t above = 0;
    int b = EXT_BUCKETS - 1;
    for (; b > 0; --b) {
        if (above + h_small[b] > max_big) break;
        above += h_small[b];
    }
    if (w->cfg.flags & GP_FLAG_BIG_STATIC_ONLY) {
        // slab worlds replicate big bodies instead of exchanging them, so only static bodies may be big: the grid
        // must host every dynamic body, whatever that costs in cell size
        int b_dyn = 0;
        for (int k = EXT_BUCKETS - 1; k > 0; --k)
            if (h_small[EXT_BUCKETS + k]) {
                b_dyn = k;
                break;
            }
        if (b < b_dyn) b = b_dyn;
    }
    // b = highest bucket that must stay in the grid; its upper edge bounds every grid-class extent
    float thr = ext_bucket_edge(b);
    if (b == EXT_BUCKETS - 1) thr = INFINITY;
    w->thr = thr;
    // classify + bounds of grid-class centres
    uint32_t init_bounds[6] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0u, 0u, 0u};
    CK(cudaMemcpyAsync(d_bounds, init_bounds, sizeof init_bounds, cudaMemcpyHostToDevice, w->st));
    if (n) {
        k_classify<<<cdiv(n, 256), 256, 0, w->st>>>(n, thr, Bodies{w->B[w->cur]}, w->A[w->cur], d_bounds);
        w->launches++;
    }
    uint32_t hb[6];
    CK(cudaMemcpyAsync(hb, d_bounds, sizeof hb, cudaMemcpyDeviceToHost, w->st));
    CK(cudaStreamSynchronize(w->st));
    float lo[3], hi[3];
    for (int k = 0; k < 3; ++k) {
        if (hb[k] == 0xffffffffu || hb[3 + k] == 0u) {
            lo[k] = 0.f, hi[k] = 1.f;  // no grid-class body
        } else {
            lo[k] = ord2f(hb[k]), hi[k] = ord2f(hb[3 + k]);
        }
    }
    const gp_config& c = w->cfg;
    const bool user_bounds = !(c.grid_min[0] == 0 && c.grid_min[1] == 0 && c.grid_min[2] == 0 && c.grid_max[0] == 0 &&
                               c.grid_max[1] == 0 && c.grid_max[2] == 0);
    if (user_bounds)
        for (int k = 0; k < 3; ++k) lo[k] = c.grid_min[k], hi[k] = c.grid_max[k];
    // margins are relative to each body's own extent (a contact moves a body by at most 0.4 of the overlap,
    // which is bounded by the body's extent)
    float m_max = c.margin_max > 0 ? c.margin_max : 2.0f;
    float m_init = c.margin_init > 0 ? c.margin_init : 0.05f;
    if (m_init > m_max) m_init = m_max;
    if (!isfinite(thr)) return fail(w, GP_ERR_BAD_ARG, "gp_upload: more than max_big bodies have a non-finite extent");
    const float ext_max = thr;  // upper bound of every grid-class extent
    // Sparse scenes: cells no finer than ~GP_TARGET_OCC bodies per cell keep the cell table small enough to
    // stay L2-resident (10 table lookups per body in K4) at the price of a few more AABB tests per body.
    float cell_min = c.cell_size;
    {
        uint64_t n_grid = 0;
        for (int k = 0; k <= b; ++k) n_grid += h_small[k];
        const double vol = std::max(1e-30, ((double)hi[0] - lo[0]) * ((double)hi[1] - lo[1]) * ((double)hi[2] - lo[2]));
        const char* env = getenv("GP_TARGET_OCC");
        const double occ = env ? atof(env) : 0.3;  // measured on uniform_16M: pair find gets cheaper with finer cells,
                                                   // the table memset + scan dearer; ~0.3 body per cell is the sweet spot
        if (n_grid > 0 && occ > 0 && !user_bounds_degenerate(lo, hi)) {
            const float auto_cell = (float)cbrt(occ * vol / (double)n_grid);
            if (auto_cell > cell_min) cell_min = auto_cell;
        }
    }
    // the cell table is sized for the finest grid (margin = m_init); coarser grids reuse it
    const uint32_t MAXC = 1u << 28;
    GridParams g0 = make_grid(lo, hi, ext_max, m_init, cell_min, MAXC);
    int bits = 1;
    while ((1ull << bits) <= (uint64_t)g0.ncells + 1) ++bits;  // keys: cells, ncells = big list, ncells+1 = dead
    w->key_bits = bits;
    if (g0.ncells + 3 > w->lb_cap) {
        CK(dmalloc(&w->LB, (size_t)g0.ncells + 4));
        w->lb_cap = g0.ncells + 3;
        CK(dmalloc(&w->lb_tiles, (size_t)cdiv(w->lb_cap, SC_TILE) + 1));
    }
    // control block
    Ctl hc{};
    hc.margin = m_init;
    hc.margin_min = m_init;
    hc.margin_max = m_max;
    {
        // lazy candidates need repair sweeps; with GP_FLAG_NO_RETRY every candidate enters the sweep at once
        const bool lazy = !(c.flags & GP_FLAG_NO_RETRY);
        const char* e1 = getenv("GP_INNER_FRAC");
        const char* e2 = getenv("GP_REPAIR_TOL");
        hc.lazy = lazy ? 1u : 0u;
        hc.margin_in = lazy ? m_init * (e1 ? (float)atof(e1) : 0.25f) : m_init;
        hc.repair_tol = e2 ? (uint32_t)atoi(e2) : 1u;
        // Measured: sweep group 0.87 -> 0.73 ms at 16M bodies, 0.47 -> 0.42 ms at 2M.
        // GP_TINY_REPAIR=0 leaves every repair to the cooperative path.
        const char* e3 = getenv("GP_TINY_REPAIR");
        w->tiny_repair = e3 ? atoi(e3) != 0 : true;
    }
    for (int k = 0; k < 3; ++k) hc.lo[k] = lo[k], hc.hi[k] = hi[k];
    hc.ext_max = ext_max;
    hc.cell_cfg = cell_min;
    hc.max_cells = g0.ncells;
    hc.grid = g0;
    hc.n_in = n, hc.n = n;
    w->max_cells = g0.ncells;
    CK(cudaMemcpyAsync(w->ctl, &hc, sizeof hc, cudaMemcpyHostToDevice, w->st));
    CK(cudaStreamSynchronize(w->st));
    w->steps_total = 0;
    return GP_OK;
}

extern "C" gp_status gp_upload(gp_world* w, uint32_t n, const float* pos3, const float* rot4, const float* scale3,
                               const float* vel3, const float* acc3, const float* box_min3, const float* box_max3,
                               const float* mass, const float* restitution, const uint8_t* is_static,
                               const uint32_t* entity, const uint32_t* rank) {
    if (!w) return GP_ERR_BAD_ARG;
    if (n && (!pos3 || !vel3 || !box_min3 || !box_max3 || !mass || !restitution || !is_static))
        return fail(w, GP_ERR_BAD_ARG, "gp_upload: a required array is NULL");
    if (n >= 0x7fffffffu) return fail(w, GP_ERR_BAD_ARG, "gp_upload: n too large");
    CK(cudaSetDevice(w->device));
    if (w->cfg.max_bodies && n > w->cfg.max_bodies)
        return fail(w, GP_ERR_CAPACITY, "gp_upload: %u bodies > max_bodies %u", n, w->cfg.max_bodies);
    const uint32_t want = std::max<uint32_t>(w->cfg.max_bodies, std::max<uint32_t>(n, 1u));
    if (want > w->cap) {
        gp_status s = alloc_bodies(w, want);
        if (s != GP_OK) return s;
    }
    uint64_t want_pairs = w->cfg.max_pairs ? w->cfg.max_pairs : std::max<uint64_t>(16ull * w->cap, 1ull << 20);
    if (want_pairs > w->pair_cap) {
        gp_status s = alloc_pairs(w, want_pairs);
        if (s != GP_OK) return s;
    }
    if (w->multi) {  // a fresh upload starts a fresh (single-GPU) world; call gp_comm_init again for slabs
        gp_multi_destroy(w->multi);
        w->multi = nullptr;
    }
    w->n = n;
    w->cur = 0;
    // stage host-layout arrays on the device, then pack
    const size_t f = sizeof(float);
    size_t off = 0;
    auto take = [&](size_t bytes) {
        size_t o = off;
        off += (bytes + 255) & ~(size_t)255;
        return o;
    };
    const size_t o_pos = take(3 * f * n), o_rot = take(4 * f * n), o_scale = take(3 * f * n), o_vel = take(3 * f * n),
                 o_acc = take(3 * f * n), o_bmin = take(3 * f * n), o_bmax = take(3 * f * n), o_mass = take(f * n),
                 o_rest = take(f * n), o_stat = take(n), o_ent = take(4 * (size_t)n), o_rank = take(4 * (size_t)n);
    gp_status s = ensure_stage(w, off + 256);
    if (s != GP_OK) return s;
    unsigned char* sg = w->stage;
    auto h2d = [&](size_t o, const void* src, size_t bytes) -> cudaError_t {
        if (!src || !bytes) return cudaSuccess;
        return cudaMemcpyAsync(sg + o, src, bytes, cudaMemcpyHostToDevice, w->st);
    };
    CK(h2d(o_pos, pos3, 3 * f * n));
    CK(h2d(o_rot, rot4, 4 * f * n));
    CK(h2d(o_scale, scale3, 3 * f * n));
    CK(h2d(o_vel, vel3, 3 * f * n));
    CK(h2d(o_acc, acc3, 3 * f * n));
    CK(h2d(o_bmin, box_min3, 3 * f * n));
    CK(h2d(o_bmax, box_max3, 3 * f * n));
    CK(h2d(o_mass, mass, f * n));
    CK(h2d(o_rest, restitution, f * n));
    CK(h2d(o_stat, is_static, n));
    CK(h2d(o_ent, entity, 4 * (size_t)n));
    CK(h2d(o_rank, rank, 4 * (size_t)n));
    if (n) {
        k_pack<<<cdiv(n, 256), 256, 0, w->st>>>(
            n, nullptr, (const float*)(sg + o_pos), rot4 ? (const float*)(sg + o_rot) : nullptr,
            scale3 ? (const float*)(sg + o_scale) : nullptr, (const float*)(sg + o_vel),
            acc3 ? (const float*)(sg + o_acc) : nullptr, (const float*)(sg + o_bmin), (const float*)(sg + o_bmax),
            (const float*)(sg + o_mass), (const float*)(sg + o_rest), (const uint8_t*)(sg + o_stat),
            entity ? (const uint32_t*)(sg + o_ent) : nullptr, rank ? (const uint32_t*)(sg + o_rank) : nullptr,
            Bodies{w->B[0]}, w->A[0], w->ext, 0, nullptr, w->rawbox);
        w->launches++;
        CK(cudaGetLastError());
        w->have_entity = entity != nullptr;
        if (entity) CK(cudaMemcpyAsync(w->entity_dev, sg + o_ent, 4 * (size_t)n, cudaMemcpyDeviceToDevice, w->st));
    }
    CK(cudaMemsetAsync(w->dmax, 0, sizeof(float) * (size_t)w->cap, w->st));
    CK(cudaMemsetAsync(w->reach, 0, sizeof(float) * (size_t)w->cap, w->st));
    s = setup_grid(w);
    if (s != GP_OK) return s;
    w->uploaded = true;
    w->home_valid = false;
    w->dead = false;
    w->slot_valid = false;
    w->last = gp_step_stats{};
    w->last_np = 0;
    return GP_OK;
}

// body id -> slot map of the current set (slots change with every step)
static gp_status ensure_slot_map(gp_world* w) {
    if (w->slot_valid || w->n == 0) return GP_OK;
    k_slot_map<<<cdiv(w->n, 256), 256, 0, w->st>>>(w->n, Bodies{w->B[w->cur]}, w->slot_of);
    w->launches++;
    CK(cudaGetLastError());
    w->slot_valid = true;
    return GP_OK;
}

static gp_status stage_state(gp_world* w, uint32_t k, const uint32_t* idx, const float* pos3, const float* vel3,
                             const float* acc3) {
    if (!w || !w->uploaded) return w ? fail(w, GP_ERR_STATE, "no bodies uploaded") : GP_ERR_BAD_ARG;
    if (w->multi) return fail(w, GP_ERR_STATE, "per-index state updates are not available on a slab world");
    CK(cudaSetDevice(w->device));
    if (k == 0) return GP_OK;
    if (idx) {
        gp_status sm = ensure_slot_map(w);
        if (sm != GP_OK) return sm;
    }
    const size_t b3 = 3 * sizeof(float) * (size_t)k;
    const size_t a3 = (b3 + 255) & ~(size_t)255;
    gp_status s = ensure_stage(w, 3 * a3 + (((size_t)k * 4 + 255) & ~(size_t)255) + 256);
    if (s != GP_OK) return s;
    unsigned char* sg = w->stage;
    if (pos3) CK(cudaMemcpyAsync(sg, pos3, b3, cudaMemcpyHostToDevice, w->st));
    if (vel3) CK(cudaMemcpyAsync(sg + a3, vel3, b3, cudaMemcpyHostToDevice, w->st));
    if (acc3) CK(cudaMemcpyAsync(sg + 2 * a3, acc3, b3, cudaMemcpyHostToDevice, w->st));
    if (idx) CK(cudaMemcpyAsync(sg + 3 * a3, idx, (size_t)k * 4, cudaMemcpyHostToDevice, w->st));
    k_set_state<<<cdiv(k, 256), 256, 0, w->st>>>(k, idx ? (const uint32_t*)(sg + 3 * a3) : nullptr, w->slot_of,
                                                 pos3 ? (const float*)sg : nullptr,
                                                 vel3 ? (const float*)(sg + a3) : nullptr,
                                                 acc3 ? (const float*)(sg + 2 * a3) : nullptr,
                                                 Bodies{w->B[w->cur]}, w->A[w->cur]);
    w->launches++;
    CK(cudaGetLastError());
    return GP_OK;
}

extern "C" gp_status gp_set_state(gp_world* w, const float* pos3, const float* vel3, const float* acc3) {
    if (!w) return GP_ERR_BAD_ARG;
    return stage_state(w, w->n, nullptr, pos3, vel3, acc3);
}

extern "C" gp_status gp_update_dirty(gp_world* w, uint32_t k, const uint32_t* idx, const float* pos3,
                                     const float* vel3, const float* acc3) {
    if (!w) return GP_ERR_BAD_ARG;
    if (k && !idx) return fail(w, GP_ERR_BAD_ARG, "gp_update_dirty: idx is NULL");
    for (uint32_t i = 0; i < k; ++i)
        if (idx[i] >= w->n) return fail(w, GP_ERR_BAD_ARG, "gp_update_dirty: idx[%u]=%u out of range", i, idx[i]);
    gp_status s = stage_state(w, k, idx, pos3, vel3, acc3);
    if (s != GP_OK) return s;
    w->track_slots = true;  // a host that sends dirty sets does so every frame: K3 keeps the id -> slot map current
    // Pinned buffers (gp_host_alloc) are copied asynchronously -- they must then stay untouched until the next call
    // that synchronises (gp_step, gp_sync, gp_download); pageable ones may be short-lived: wait for the copies.
    bool pinned = true;
    for (const void* p : {(const void*)idx, (const void*)pos3, (const void*)vel3, (const void*)acc3}) {
        if (!p) continue;
        cudaPointerAttributes at{};
        if (cudaPointerGetAttributes(&at, p) != cudaSuccess || at.type != cudaMemoryTypeHost) {
            cudaGetLastError();
            pinned = false;
        }
    }
    if (!pinned) CK(cudaStreamSynchronize(w->st));
    return GP_OK;
}

extern "C" gp_status gp_update_shape(gp_world* w, uint32_t k, const uint32_t* idx, const float* box_min3,
                                     const float* box_max3, const float* rot4, const float* scale3, const float* mass,
                                     const float* restitution, const uint8_t* is_static) {
    if (!w || !w->uploaded) return w ? fail(w, GP_ERR_STATE, "no bodies uploaded") : GP_ERR_BAD_ARG;
    if (k == 0) return GP_OK;
    if (!idx || !box_min3 || !box_max3 || !mass || !restitution || !is_static)
        return fail(w, GP_ERR_BAD_ARG, "gp_update_shape: a required array is NULL");
    for (uint32_t i = 0; i < k; ++i)
        if (idx[i] >= w->n) return fail(w, GP_ERR_BAD_ARG, "gp_update_shape: idx[%u]=%u out of range", i, idx[i]);
    if (w->multi) return fail(w, GP_ERR_STATE, "per-index shape updates are not available on a slab world");
    CK(cudaSetDevice(w->device));
    {
        gp_status sm = ensure_slot_map(w);
        if (sm != GP_OK) return sm;
    }
    const size_t f = sizeof(float);
    size_t off = 0;
    auto take = [&](size_t bytes) {
        size_t o = off;
        off += (bytes + 255) & ~(size_t)255;
        return o;
    };
    const size_t o_idx = take(4 * (size_t)k), o_rot = take(4 * f * k), o_scale = take(3 * f * k),
                 o_bmin = take(3 * f * k), o_bmax = take(3 * f * k), o_mass = take(f * k), o_rest = take(f * k),
                 o_stat = take(k);
    gp_status s = ensure_stage(w, off + 256);
    if (s != GP_OK) return s;
    unsigned char* sg = w->stage;
    CK(cudaMemcpyAsync(sg + o_idx, idx, 4 * (size_t)k, cudaMemcpyHostToDevice, w->st));
    if (rot4) CK(cudaMemcpyAsync(sg + o_rot, rot4, 4 * f * k, cudaMemcpyHostToDevice, w->st));
    if (scale3) CK(cudaMemcpyAsync(sg + o_scale, scale3, 3 * f * k, cudaMemcpyHostToDevice, w->st));
    CK(cudaMemcpyAsync(sg + o_bmin, box_min3, 3 * f * k, cudaMemcpyHostToDevice, w->st));
    CK(cudaMemcpyAsync(sg + o_bmax, box_max3, 3 * f * k, cudaMemcpyHostToDevice, w->st));
    CK(cudaMemcpyAsync(sg + o_mass, mass, f * k, cudaMemcpyHostToDevice, w->st));
    CK(cudaMemcpyAsync(sg + o_rest, restitution, f * k, cudaMemcpyHostToDevice, w->st));
    CK(cudaMemcpyAsync(sg + o_stat, is_static, k, cudaMemcpyHostToDevice, w->st));
    k_pack<<<cdiv(k, 256), 256, 0, w->st>>>(k, (const uint32_t*)(sg + o_idx), nullptr,
                                            rot4 ? (const float*)(sg + o_rot) : nullptr,
                                            scale3 ? (const float*)(sg + o_scale) : nullptr, nullptr, nullptr,
                                            (const float*)(sg + o_bmin), (const float*)(sg + o_bmax),
                                            (const float*)(sg + o_mass), (const float*)(sg + o_rest),
                                            (const uint8_t*)(sg + o_stat), nullptr, nullptr, Bodies{w->B[w->cur]},
                                            w->A[w->cur], nullptr, 1, w->slot_of, w->rawbox);
    w->launches++;
    CK(cudaGetLastError());
    // a body that outgrew the cell joins the big list; the grid itself is kept
    k_classify<<<cdiv(w->n, 256), 256, 0, w->st>>>(w->n, w->thr, Bodies{w->B[w->cur]}, w->A[w->cur],
                                                   w->d_small + 2 * EXT_BUCKETS);
    w->launches++;
    w->home_valid = false;
    CK(cudaStreamSynchronize(w->st));
    return GP_OK;
}

// ---- profiling marks: event e(step, g) is recorded BEFORE group g; e(step, NUM) after the last ----
static inline bool prof_on(const gp_world* w) {
    return (w->cfg.flags & GP_FLAG_PROFILE) && w->prof_n < (uint32_t)gp_world::PROF_STEPS;
}
static void prof_begin_step(gp_world* w) {
    if (!prof_on(w)) return;
    if (w->prof_ev.empty()) {
        w->prof_ev.resize((size_t)gp_world::PROF_STEPS * (GP_NUM_KERNEL_GROUPS + 1));
        for (auto& e : w->prof_ev) cudaEventCreate(&e);
    }
    w->prof_group = -1;
}
static void prof_mark(gp_world* w, int group) {  // everything enqueued from here on belongs to `group`
    if (!prof_on(w)) return;
    cudaEvent_t* ev = &w->prof_ev[(size_t)w->prof_n * (GP_NUM_KERNEL_GROUPS + 1)];
    if (w->prof_group >= 0) w->prof_launch[w->prof_group] += (uint32_t)(w->launches - w->prof_mark_launches);
    // groups may be skipped: give skipped groups a zero-length interval
    for (int g = w->prof_group + 1; g <= group; ++g) cudaEventRecord(ev[g], w->st);
    w->prof_group = group;
    w->prof_mark_launches = w->launches;
}
static void prof_end_step(gp_world* w) {
    if (!prof_on(w)) return;
    prof_mark(w, GP_NUM_KERNEL_GROUPS);
    w->prof_n++;
}

extern "C" gp_status gp_kernel_times_read(gp_world* w, gp_kernel_times* out) {
    if (!w || !out) return GP_ERR_BAD_ARG;
    if (!(w->cfg.flags & GP_FLAG_PROFILE)) return fail(w, GP_ERR_STATE, "gp_kernel_times_read needs GP_FLAG_PROFILE");
    CK(cudaSetDevice(w->device));
    CK(cudaStreamSynchronize(w->st));
    memset(out, 0, sizeof *out);
    out->steps = w->prof_n;
    for (uint32_t s = 0; s < w->prof_n; ++s) {
        cudaEvent_t* ev = &w->prof_ev[(size_t)s * (GP_NUM_KERNEL_GROUPS + 1)];
        for (int g = 0; g < GP_NUM_KERNEL_GROUPS; ++g) {
            float ms = 0.f;
            CK(cudaEventElapsedTime(&ms, ev[g], ev[g + 1]));
            out->ms[g] += ms;
        }
    }
    for (int g = 0; g < GP_NUM_KERNEL_GROUPS; ++g) out->launches[g] = w->prof_launch[g], w->prof_launch[g] = 0;
    w->prof_n = 0;
    if (w->fine && !w->fine_reset_done) {  // the fine-grained accumulators restart ONCE (bench.py drops its warm-up this way)
        w->fine_reset_done = true;
        fine_collect(w);
        w->fine_acc.clear();
        w->fine_steps = 0;
        Ctl* dc = w->ctl;
        CK(cudaMemsetAsync(dc->dbg_phase_ns, 0, sizeof dc->dbg_phase_ns, w->st));
        CK(cudaMemsetAsync(dc->dbg_level_ns, 0, sizeof dc->dbg_level_ns + sizeof dc->dbg_level_n, w->st));
        CK(cudaMemsetAsync(dc->dbg_counts, 0, sizeof dc->dbg_counts, w->st));
        CK(cudaMemsetAsync(&dc->dbg_tail_rounds, 0, sizeof(uint32_t), w->st));
        CK(cudaStreamSynchronize(w->st));
    }
    return GP_OK;
}

// Upper bound of the live body count that sizes the grids (the exact count is device-resident: ctl->n).
static inline uint32_t n_bound(const gp_world* w) { return w->multi ? w->cap : w->n; }

// Slab worlds: where the per-step exactness proof keeps its marks (gp_multi_impl.cuh)
static uint32_t* gp_multi_taint_bits(gp_multi* m);
static uint32_t* gp_multi_seed_bits(gp_multi* m);
static uint32_t* gp_multi_uf_parent(gp_multi* m);

// K6: the whole contact phase -- ordered sweep, verification of the speculative pair set, repair rounds while they
// are needed, the slab exactness proof, end of step -- is ONE cooperative launch (gp_contact_kernel.cuh).
// start_phase 1: the host continues the repair loop of a synchronous step (gp_step).
static const uint32_t REPAIR_ROUNDS_IN_KERNEL = 6;
static gp_status enqueue_contacts(gp_world* w, float dt, const float* d, uint32_t start_phase, int latch,
                                  uint32_t max_rounds) {
    const int c = w->cur, o = c ^ 1;
    ContactArgs a{};
    a.B1 = Bodies{w->B[o]}, a.B0 = Bodies{w->B[c]}, a.A0 = w->A[c], a.A1 = w->A[o];
    a.sbox = w->sbox, a.LB = w->LB;
    a.pairs = w->pairs, a.pair_cap = (uint32_t)w->pair_cap, a.watch = w->watch;
    a.frontier0 = w->frontier[0], a.frontier1 = w->frontier[1];
    a.ctl = w->ctl, a.ctl_snap = w->ctl_snap;
    a.dmax = w->dmax, a.reach = w->reach;
    a.moved_bits = w->moved_bits, a.esc_bits = w->esc_bits, a.escapees = w->escapees, a.esc_cap = w->cap;
    a.offs = w->offs, a.adj = w->adj, a.chain_cnt = w->chain_cnt, a.perm = w->perm;
    a.xhead = w->xhead, a.xnext = w->xnext;
    a.ra = RepairArrays{w->dirty_bits, w->dlist, w->dpairs, w->offs2, w->deg2, w->adj2, w->dheavy, w->heavy_cap};
    a.recs = (uint32_t*)w->adj2;  // scratch of the general repair, which runs (if at all) behind the per-component one
    a.rec_cap = (uint32_t)std::min<size_t>(w->pair_cap * 4 / TINY_REC, 1u << 20);
    a.taint_bits = w->multi ? gp_multi_taint_bits(w->multi) : nullptr;
    a.seed_bits = w->multi ? gp_multi_seed_bits(w->multi) : nullptr;
    a.uf_parent = w->multi ? gp_multi_uf_parent(w->multi) : nullptr;
    a.dt = dt, a.d16 = d[0], a.d18 = d[1], a.d35 = d[2];
    a.start_phase = start_phase, a.max_rounds = max_rounds, a.tiny = w->tiny_repair ? 1u : 0u, a.latch = latch;
    a.profile = w->fine ? 1u : 0u;
    void* args[] = {&a};
    CK(cudaLaunchCooperativeKernel((void*)k_contacts, dim3(w->sweep_grid), dim3(COOP_THREADS), args, sizeof(ContactSmem),
                                   w->st));
    w->launches++;
    fine_mark(w, start_phase ? "k_contacts(repair)" : "k_contacts");
    return GP_OK;
}

// K5: chain offsets, fill, sort + link.
static void enqueue_chains(gp_world* w) {
    const uint32_t nb = n_bound(w);
    cudaStream_t st = w->st;
    exclusive_scan_u32(w->chain_cnt, nb, w->offs, w->tile_sums, st, &w->launches, nullptr);
    fine_mark(w, "chain scan");
    k_fill_chains<<<w->sm_count * 8, 256, 0, st>>>(w->pairs, (uint32_t)w->pair_cap, w->offs, w->chain_cnt, w->adj,
                                                   w->ctl, nullptr);
    fine_mark(w, "k_fill_chains");
    k_sort_chains<<<cdiv(nb, 256), 256, 0, st>>>(w->offs, w->adj, w->pairs, w->heavy, w->heavy_cap, w->ctl, nullptr);
    fine_mark(w, "k_sort_chains");
    k_sort_heavy<<<w->sm_count * 8, 256, 0, st>>>(w->offs, w->adj, w->pairs, w->heavy, w->heavy_cap, w->ctl, nullptr);
    fine_mark(w, "k_sort_heavy");
    w->launches += 3;
}

// One step = phase A (cell keys + histogram; on a slab also the packing of what the neighbours need), the slab
// exchange (multi-GPU only, gp_multi_impl.cuh), phase B (everything else).  Reads set `cur`, leaves the result (in
// cell order) in set `cur^1`.
static void step_damping(float dt, const float* damp3, float* d) {
    if (damp3)
        d[0] = damp3[0], d[1] = damp3[1], d[2] = damp3[2];
    else
        gp_damping_factors(dt, d);
}

static gp_status enqueue_phase_a(gp_world* w, float dt, const float* damp3) {
    cudaStream_t st = w->st;
    float d[3];
    step_damping(dt, damp3, d);
    const int c = w->cur;
    const uint32_t nb = n_bound(w);  // grids are sized for this; kernels stop at the device-resident counts
    prof_begin_step(w);
    fine_mark(w, "begin");
    w->slot_valid = false;  // (the callers set it again once a step is accepted, if K3 keeps the map current)
    if (nb) {
        prof_mark(w, 0);
        // K1 + K2: one-pass radix (bucket) sort by cell key: histogram with the keys, scan, scatter
        const uint32_t ncnt = w->max_cells + 2;  // counters: every cell of the finest grid + big list + dead
        CK(cudaMemsetAsync(w->LB, 0, sizeof(uint32_t) * ((size_t)ncnt + 1), st));
        fine_mark(w, "memset LB");
        if (w->multi) {
            gp_multi_begin_step(w->multi);  // (peer-to-peer transport: wait for buffer credit)
            k_keys<true><<<cdiv(nb, 256 * K1_TILES), 256, 0, st>>>(Bodies{w->B[c]}, w->A[c], w->key, w->lrank, w->LB, w->ctl,
                                                        gp_multi_send(w->multi, 0), gp_multi_send(w->multi, 1), dt, d[0],
                                                        d[1], d[2]);
            w->launches += 1;
            fine_mark(w, "k_keys<slab>");
            gp_multi_seal(w->multi);        // counts into the headers (+ ready flags on the peer-to-peer transport)
            fine_mark(w, "seal");
        } else {
            k_keys<false><<<cdiv(nb, 256), 256, 0, st>>>(Bodies{w->B[c]}, w->A[c], w->key, w->lrank, w->LB, w->ctl,
                                                         nullptr, nullptr, dt, d[0], d[1], d[2]);
            w->launches++;
            fine_mark(w, "k_keys");
        }
    }
    CK(cudaGetLastError());
    return GP_OK;
}

static gp_status enqueue_phase_b(gp_world* w, float dt, const float* damp3, int latch, int repair_rounds) {
    cudaStream_t st = w->st;
    float d[3];
    step_damping(dt, damp3, d);
    const int c = w->cur, o = c ^ 1;
    const uint32_t nb = n_bound(w);
    if (nb) {
        if (w->multi) {
            prof_mark(w, 1);
            gp_multi_before_append(w->multi);  // (peer-to-peer transport: wait for the neighbours' ready flags)
            fine_mark(w, "wait neighbours");
            k_append<<<w->sm_count * 2, 256, 0, st>>>(gp_multi_recv(w->multi, 0), gp_multi_recv(w->multi, 1),
                                                      Bodies{w->B[c]}, w->A[c], w->key, w->lrank, w->LB, w->ctl);
            w->launches++;
            gp_multi_after_append(w->multi);   // (peer-to-peer transport: hand the buffer credit back)
            fine_mark(w, "k_append+credit");
        }
        prof_mark(w, 2);
        const uint32_t ncnt = w->max_cells + 2;
        exclusive_scan_u32(w->LB, ncnt, w->LB, w->lb_tiles, st, &w->launches);
        fine_mark(w, "cell scan");
        prof_mark(w, 3);
        k_scatter_cells<<<cdiv(nb, 256), 256, 0, st>>>(w->key, w->lrank, Bodies{w->B[c]}, w->A[c], Bodies{w->B[o]}, w->A[o],
                                                       w->sbox, w->chain_cnt, w->perm, w->dmax, w->reach, w->moved_bits,
                                                       w->esc_bits, w->xhead, w->LB, w->ctl,
                                                       (w->track_slots && !w->multi) ? w->slot_of : nullptr,
                                                       w->multi ? gp_multi_seed_bits(w->multi) : nullptr,
                                                       w->multi ? gp_multi_uf_parent(w->multi) : nullptr, dt, d[0], d[1], d[2]);
        w->launches += 1;
        fine_mark(w, "k_scatter_cells");
        prof_mark(w, 4);
        k_find_pairs<<<std::min<uint32_t>(w->fp_grid, cdiv(nb, WP_THREADS)), WP_THREADS, 0, st>>>(
            w->sbox, w->LB, w->pairs, w->watch, w->chain_cnt, (uint32_t)w->pair_cap, w->ctl);
        w->launches += 1;
        fine_mark(w, "k_find_pairs");
        prof_mark(w, 5);
        enqueue_chains(w);
        prof_mark(w, 6);
        // sweep + verification of the speculative broad phase + repair rounds + (slabs) exactness proof + end of step
        gp_status s6 = enqueue_contacts(w, dt, d, 0, latch, (uint32_t)repair_rounds);
        if (s6 != GP_OK) return s6;
        prof_mark(w, 7);
    } else {
        prof_mark(w, 7);
        k_finish_step<<<1, 1, 0, st>>>(w->ctl, w->ctl_snap, latch);
        w->launches++;
    }
    prof_end_step(w);
    CK(cudaGetLastError());
    return GP_OK;
}

static gp_status enqueue_step(gp_world* w, float dt, const float* damp3, int latch, int repair_rounds) {
    gp_status s = enqueue_phase_a(w, dt, damp3);
    if (s != GP_OK) return s;
    if (w->multi) {
        s = gp_multi_enqueue_exchange(w->multi);
        if (s != GP_OK) return s;
    }
    return enqueue_phase_b(w, dt, damp3, latch, repair_rounds);
}

static void fill_stats(gp_world* w, const Ctl& c, uint32_t retries) {
    gp_step_stats& s = w->last;
    s.n_bodies = c.n;  // live bodies of the step (device-resident count; on a slab: owned + halo copies)
    s.n_big = c.n - std::min(c.n_small, c.n);
    s.n_candidates = (uint64_t)c.n_pairs + c.n_watch - c.n_activated;
    s.n_snapshot = c.n_snapshot;
    s.n_contacts = c.n_contacts;
    s.n_rounds = c.n_rounds;
    s.n_retries = retries + c.n_repairs;
    s.margin = c.margin;
    uint32_t mb = c.max_disp_bits;
    memcpy(&s.max_displacement, &mb, 4);
    s.cell_size = 1.0f / c.grid.inv_cell;
    s.grid_dim[0] = c.grid.dx, s.grid_dim[1] = c.grid.dy, s.grid_dim[2] = c.grid.dz;
    s.n_heavy = c.n_heavy;
    s.steps_total = w->steps_total;
    s.steps_inexact = c.steps_inexact;
    s.n_migrated_out = c.n_migrated_out, s.n_migrated_in = c.n_migrated_in;
    s.n_halo_out = c.n_sent[0] + c.n_sent[1] - c.n_migrated_out;
    s.n_halo_in = c.n_recv - c.n_migrated_in;
    s.n_unverified = c.n_unverified;
    s.unverified_total = c.unverified_total;
    s.n_watch = c.n_watch;
    s.n_activated = c.n_activated;
    w->last_np = (uint32_t)std::min<uint64_t>(c.n_pairs, w->pair_cap);
}

static gp_status read_ctl(gp_world* w) {
    CK(cudaMemcpyAsync(w->h_ctl, w->ctl_snap, sizeof(Ctl), cudaMemcpyDeviceToHost, w->st));
    CK(cudaStreamSynchronize(w->st));
    if (w->fine) fine_collect(w);
    return GP_OK;
}

static const int REPAIR_ROUNDS_ENQUEUED = (int)REPAIR_ROUNDS_IN_KERNEL;  // repair rounds a step may take on the device

extern "C" gp_status gp_step(gp_world* w, float dt, const float* damp3, gp_step_stats* stats) {
    if (!w) return GP_ERR_BAD_ARG;
    if (!w->uploaded) return fail(w, GP_ERR_STATE, "gp_step before gp_upload");
    CK(cudaSetDevice(w->device));
    if (w->multi) {  // slab worlds step asynchronously (the exchange is collective); failures latch
        gp_status a = gp_step_async(w, dt, damp3, 1);
        if (a != GP_OK) return a;
        return gp_sync(w, stats);
    }
    gp_status result = GP_OK;
    uint32_t retries = 0;
    float d[3];
    if (damp3)
        d[0] = damp3[0], d[1] = damp3[1], d[2] = damp3[2];
    else
        gp_damping_factors(dt, d);
    const bool no_retry = (w->cfg.flags & GP_FLAG_NO_RETRY) != 0;
    CK(cudaEventRecord(w->ev0, w->st));
    for (;;) {
        gp_status s = enqueue_step(w, dt, damp3, 0, no_retry ? 0 : REPAIR_ROUNDS_ENQUEUED);
        if (s != GP_OK) return s;
        s = read_ctl(w);
        if (s != GP_OK) return s;
        // the candidate set was still growing: keep repairing until k_requery finds nothing (exact result)
        int extra = 0;
        while (w->h_ctl->step_flags == SF_MARGIN && !no_retry && extra < 64) {
            s = enqueue_contacts(w, dt, d, 1, 0, REPAIR_ROUNDS_IN_KERNEL);
            if (s != GP_OK) return s;
            s = read_ctl(w);
            if (s != GP_OK) return s;
            ++extra;
        }
        const Ctl& c = *w->h_ctl;
        const uint32_t f = c.step_flags;
        if (f & SF_INTERNAL) return fail(w, GP_ERR_CUDA, "a device-side bounded wait expired (flags 0x%x)", f);
        if (f & (SF_HEAVY_OVERFLOW | SF_ESCAPEE_OVERFLOW))
            return fail(w, GP_ERR_CAPACITY, "internal work list overflow (flags 0x%x)", f);
        if (f & SF_PAIR_OVERFLOW) {
            if (no_retry || retries >= 8)
                return fail(w, GP_ERR_PAIR_OVERFLOW, "%u candidate pairs > max_pairs %llu", c.n_pairs,
                            (unsigned long long)w->pair_cap);
            const uint64_t need = std::max<uint64_t>(c.n_pairs, c.n_watch);
            gp_status a = alloc_pairs(w, std::max<uint64_t>(2 * w->pair_cap, need + need / 4));
            if (a != GP_OK) return a;
            ++retries;
            result = GP_WARN_MARGIN_RETRIED;
            continue;  // k_finish_step reset the step; re-run it from the untouched input buffers
        }
        if (f & SF_MARGIN) {
            fill_stats(w, c, retries);
            if (stats) *stats = w->last;
            const uint32_t nrep = c.n_repairs;
            const float mrg = c.margin;
            // give up on this step: drop the per-step counters (the input set is untouched)
            k_finish_step<<<1, 1, 0, w->st>>>(w->ctl, w->ctl_snap, 2);
            w->launches += 1;
            CK(cudaStreamSynchronize(w->st));
            return fail(w, GP_ERR_MARGIN,
                        "the candidate pair set was still growing after %u repair sweeps (margin %.4g): step not applied",
                        nrep, mrg);
        }
        // repair sweeps for activated watch pairs are routine; bodies out-running the OUTER margin are worth a note
        if (c.n_requeried) result = GP_WARN_MARGIN_RETRIED;
        w->cur ^= 1;
        w->steps_total++;
        w->slot_valid = w->track_slots && !w->multi;  // K3 rewrote the id -> slot map while it scattered
        fill_stats(w, c, retries);
        break;
    }
    CK(cudaEventRecord(w->ev1, w->st));
    w->timed = true;
    if (stats) *stats = w->last;
    return result;
}

extern "C" gp_status gp_step_async(gp_world* w, float dt, const float* damp3, uint32_t nsteps) {
    if (!w) return GP_ERR_BAD_ARG;
    if (!w->uploaded) return fail(w, GP_ERR_STATE, "gp_step_async before gp_upload");
    CK(cudaSetDevice(w->device));
    if (w->multi && gp_multi_is_local(w->multi))
        return fail(w, GP_ERR_STATE, "in-process slabs advance together: use gp_step_group");
    if (w->dead)
        return fail(w, GP_ERR_PEER_TIMEOUT, "a slab neighbour timed out earlier: this world is incomplete, upload again");
    CK(cudaEventRecord(w->ev0, w->st));
    for (uint32_t s = 0; s < nsteps; ++s) {
        gp_status r = enqueue_step(w, dt, damp3, 1, REPAIR_ROUNDS_ENQUEUED);
        if (r != GP_OK) return r;
        w->cur ^= 1;
        w->steps_total++;
        w->slot_valid = w->track_slots && !w->multi;
    }
    CK(cudaEventRecord(w->ev1, w->st));
    w->timed = true;
    return GP_OK;
}

extern "C" gp_status gp_sync(gp_world* w, gp_step_stats* stats) {
    if (!w) return GP_ERR_BAD_ARG;
    CK(cudaSetDevice(w->device));
    if (!w->uploaded) {
        CK(cudaStreamSynchronize(w->st));
        return GP_OK;
    }
    gp_status s = read_ctl(w);
    if (s != GP_OK) return s;
    const Ctl& c = *w->h_ctl;
    if (w->steps_total) fill_stats(w, c, 0);
    if (stats) *stats = w->last;
    if (c.unverified_total) {  // "since the last gp_sync": start a fresh sum
        CK(cudaMemsetAsync(&w->ctl->unverified_total, 0, sizeof(uint32_t), w->st));
        CK(cudaMemsetAsync(&w->ctl_snap->unverified_total, 0, sizeof(uint32_t), w->st));
    }
    if (getenv("GP_DEBUG_TINY"))
        fprintf(stderr, "[gp] tiny repair: rounds ok %u, fell back %u, components %u, too large %u, largest %u bodies / %u pairs\n",
                c.dbg_tiny[0], c.dbg_tiny[1], c.dbg_tiny[2], c.dbg_tiny[3], c.dbg_tiny[4], c.dbg_tiny[5]);
    if (c.sticky_flags) {
        const uint32_t f = c.sticky_flags;
        // clear the latch so the caller can continue after reacting
        Ctl* dc = w->ctl;
        CK(cudaMemsetAsync(&dc->sticky_flags, 0, sizeof(uint32_t), w->st));
        CK(cudaMemsetAsync(&dc->steps_inexact, 0, sizeof(uint32_t), w->st));  // reported below: start a fresh count
        CK(cudaMemsetAsync(&w->ctl_snap->sticky_flags, 0, sizeof(uint32_t), w->st));
        CK(cudaMemsetAsync(&w->ctl_snap->steps_inexact, 0, sizeof(uint32_t), w->st));
        CK(cudaStreamSynchronize(w->st));
        if (f & SF_PEER_TIMEOUT) {
            w->dead = true;
            return fail(w, GP_ERR_PEER_TIMEOUT, "a slab neighbour did not arrive at the exchange within the time-out "
                        "(GP_PEER_TIMEOUT_MS): bodies in flight are lost, %u step(s) invalid; upload again", c.steps_inexact);
        }
        if (f & SF_EXCHANGE_OVERFLOW)
            return fail(w, GP_ERR_CAPACITY, "%u async step(s) were inexact: a slab exchange buffer (max_exchange) or "
                        "max_bodies was too small", c.steps_inexact);
        if (f & SF_PAIR_OVERFLOW)
            return fail(w, GP_ERR_PAIR_OVERFLOW, "%u async step(s) were inexact: candidate pairs exceeded max_pairs",
                        c.steps_inexact);
        if (f & SF_MARGIN)
            return fail(w, GP_ERR_MARGIN,
                        "%u async step(s) were inexact: the candidate pair set was still growing when the "
                        "pre-enqueued repair sweeps ran out", c.steps_inexact);
        return fail(w, GP_ERR_CAPACITY, "%u async step(s) hit an internal work list overflow (0x%x)", c.steps_inexact, f);
    }
    return GP_OK;
}

extern "C" gp_status gp_last_step_ms(gp_world* w, float* ms) {
    if (!w || !ms) return GP_ERR_BAD_ARG;
    if (!w->timed) return fail(w, GP_ERR_STATE, "no step timed yet");
    CK(cudaSetDevice(w->device));
    CK(cudaEventSynchronize(w->ev1));
    CK(cudaEventElapsedTime(ms, w->ev0, w->ev1));
    return GP_OK;
}

extern "C" gp_status gp_download(gp_world* w, float* pos3, float* vel3) {
    if (!w || !w->uploaded) return w ? fail(w, GP_ERR_STATE, "no bodies uploaded") : GP_ERR_BAD_ARG;
    if (w->multi) return fail(w, GP_ERR_STATE, "gp_download addresses bodies by upload index; on a slab world use gp_download_owned");
    CK(cudaSetDevice(w->device));
    const uint32_t n = w->n;
    if (n == 0) return gp_sync(w, nullptr);
    const size_t b3 = 3 * sizeof(float) * (size_t)n;
    const size_t a3 = (b3 + 255) & ~(size_t)255;
    gp_status s = ensure_stage(w, 2 * a3 + 256);
    if (s != GP_OK) return s;
    float* dp = (float*)w->stage;
    float* dv = (float*)(w->stage + a3);
    k_unpack<<<cdiv(n, 256), 256, 0, w->st>>>(n, Bodies{w->B[w->cur]}, pos3 ? dp : nullptr,
                                              vel3 ? dv : nullptr, nullptr, 1);
    w->launches++;
    if (pos3) CK(cudaMemcpyAsync(pos3, dp, b3, cudaMemcpyDeviceToHost, w->st));
    if (vel3) CK(cudaMemcpyAsync(vel3, dv, b3, cudaMemcpyDeviceToHost, w->st));
    CK(cudaStreamSynchronize(w->st));
    return gp_sync(w, nullptr);
}

static gp_status next_rows_ready(gp_world* w, const char* what);

extern "C" gp_status gp_cap_velocity(gp_world* w, uint32_t k, const uint32_t* idx) {
    if (!w) return GP_ERR_BAD_ARG;
    gp_status s = next_rows_ready(w, "gp_cap_velocity");
    if (s != GP_OK) return s;
    if (!idx) k = w->n;
    if (k == 0) return GP_OK;
    for (uint32_t i = 0; idx && i < k; ++i)
        if (idx[i] >= w->n) return fail(w, GP_ERR_BAD_ARG, "gp_cap_velocity: idx[%u]=%u out of range", i, idx[i]);
    uint32_t* d_idx = nullptr;
    if (idx) {
        s = ensure_stage(w, (size_t)k * 4 + 256);
        if (s != GP_OK) return s;
        d_idx = (uint32_t*)w->stage;
        CK(cudaMemcpyAsync(d_idx, idx, (size_t)k * 4, cudaMemcpyHostToDevice, w->st));
    }
    k_cap_velocity<<<cdiv(k, 256), 256, 0, w->st>>>(k, d_idx, w->slot_of, Bodies{w->B[w->cur]});
    w->launches++;
    CK(cudaGetLastError());
    if (idx) CK(cudaStreamSynchronize(w->st));  // idx may be pageable and short-lived; the staging buffer is shared
    return GP_OK;
}

// ---- pipelined frames ----
static gp_status ensure_frames(gp_world* w) {
    if (!w->st_h2d) {
        CK(cudaStreamCreateWithFlags(&w->st_h2d, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&w->st_d2h, cudaStreamNonBlocking));
        for (int b = 0; b < 2; ++b)
            for (cudaEvent_t* e : {&w->ev_h2d[b], &w->ev_in_free[b], &w->ev_unpacked[b], &w->ev_d2h[b]})
                CK(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
    }
    if (w->frame_cap < w->n) {
        // (buffers of frames in flight: none -- the first frame after a (re)allocation waits for everything)
        CK(cudaStreamSynchronize(w->st));
        CK(cudaStreamSynchronize(w->st_h2d));
        CK(cudaStreamSynchronize(w->st_d2h));
        for (int b = 0; b < 2; ++b) {
            CK(dmalloc(&w->fin[b], 9 * (size_t)w->cap));
            CK(dmalloc(&w->fout[b], 6 * (size_t)w->cap));
        }
        CK(dmalloc(&w->fout8, 2 * (size_t)w->cap));
        CK(dmalloc(&w->home, 3 * (size_t)w->cap));
        w->frame_cap = w->cap;
        w->home_valid = false;
        w->frames_submitted = 0;
    }
    if (!w->home_valid) {
        k_build_home<<<cdiv(w->n, 256), 256, 0, w->st>>>(w->n, Bodies{w->B[w->cur]}, w->A[w->cur], w->home);
        w->launches++;
        CK(cudaGetLastError());
        w->home_valid = true;
    }
    return GP_OK;
}

extern "C" gp_status gp_frame_submit(gp_world* w, const float* pos3, const float* vel3, const float* acc3, float dt,
                                     const float* damp3, float* state6_out) {
    if (!w) return GP_ERR_BAD_ARG;
    if (!w->uploaded) return fail(w, GP_ERR_STATE, "gp_frame_submit before gp_upload");
    if (w->multi) return fail(w, GP_ERR_STATE, "gp_frame_submit addresses bodies by upload index: not available on a slab world");
    if (!pos3 || !vel3 || !acc3 || !state6_out) return fail(w, GP_ERR_BAD_ARG, "gp_frame_submit: a buffer is NULL");
    CK(cudaSetDevice(w->device));
    const uint32_t n = w->n;
    if (n == 0) return GP_OK;
    gp_status s = ensure_frames(w);
    if (s != GP_OK) return s;
    const int b = (int)(w->frames_submitted & 1u);
    const bool reuse = w->frames_submitted >= 2;  // the buffers of parity b were last used by frame t-2
    const size_t row = 3 * sizeof(float) * (size_t)n, stride = 3 * (size_t)w->frame_cap;
    float* in = w->fin[b];
    // copy stream: inputs of this frame (after the refresh of frame t-2 has consumed the buffer)
    if (reuse) CK(cudaStreamWaitEvent(w->st_h2d, w->ev_in_free[b], 0));
    CK(cudaMemcpyAsync(in, pos3, row, cudaMemcpyHostToDevice, w->st_h2d));
    CK(cudaMemcpyAsync(in + stride, vel3, row, cudaMemcpyHostToDevice, w->st_h2d));
    CK(cudaMemcpyAsync(in + 2 * stride, acc3, row, cudaMemcpyHostToDevice, w->st_h2d));
    CK(cudaEventRecord(w->ev_h2d[b], w->st_h2d));
    // step stream: refresh (bodies land in id order; K1..K3 sort them as in every step), step, read-back
    CK(cudaStreamWaitEvent(w->st, w->ev_h2d[b], 0));
    k_set_state_all<<<cdiv(n, 256), 256, 0, w->st>>>(n, w->home, in, in + stride, in + 2 * stride, Bodies{w->B[w->cur]},
                                                      w->A[w->cur]);
    w->launches++;
    fine_mark(w, "k_set_state_all");
    CK(cudaEventRecord(w->ev_in_free[b], w->st));
    s = enqueue_step(w, dt, damp3, 1, REPAIR_ROUNDS_ENQUEUED);
    if (s != GP_OK) return s;
    w->cur ^= 1;
    w->steps_total++;
    if (reuse) CK(cudaStreamWaitEvent(w->st, w->ev_d2h[b], 0));  // the download of frame t-2 has left the buffer
    k_unpack8<<<cdiv(n, 256), 256, 0, w->st>>>(n, Bodies{w->B[w->cur]}, w->fout8);
    k_rows8_to_6<<<cdiv(n, 256), 256, 0, w->st>>>(n, w->fout8, w->fout[b]);
    w->launches += 2;
    fine_mark(w, "k_unpack8+rows6");
    CK(cudaEventRecord(w->ev_unpacked[b], w->st));
    // download stream
    CK(cudaStreamWaitEvent(w->st_d2h, w->ev_unpacked[b], 0));
    CK(cudaMemcpyAsync(state6_out, w->fout[b], 6 * sizeof(float) * (size_t)n, cudaMemcpyDeviceToHost, w->st_d2h));
    CK(cudaEventRecord(w->ev_d2h[b], w->st_d2h));
    w->frames_submitted++;
    w->slot_valid = w->track_slots;
    return GP_OK;
}

extern "C" gp_status gp_frame_wait(gp_world* w, uint32_t max_in_flight, gp_step_stats* stats) {
    if (!w) return GP_ERR_BAD_ARG;
    CK(cudaSetDevice(w->device));
    if (w->frames_submitted == 0 || !w->st_d2h) return max_in_flight ? GP_OK : gp_sync(w, stats);
    if (max_in_flight == 0) {
        CK(cudaStreamSynchronize(w->st_d2h));
        return gp_sync(w, stats);
    }
    if (max_in_flight >= 2) return GP_OK;  // the events bound the queue at two frames
    // one frame may stay in flight: wait for the download of the frame before the last one
    if (w->frames_submitted >= 2) CK(cudaEventSynchronize(w->ev_d2h[(int)((w->frames_submitted - 2) & 1u)]));
    return GP_OK;
}

extern "C" gp_status gp_pairs(gp_world* w, int32_t which, uint32_t* pairs2, uint64_t cap, uint64_t* n_out) {
    if (!w || !n_out) return GP_ERR_BAD_ARG;
    if (!(w->cfg.flags & GP_FLAG_RECORD_PAIRS)) return fail(w, GP_ERR_STATE, "gp_pairs needs GP_FLAG_RECORD_PAIRS");
    if (which != 0 && which != 1) return fail(w, GP_ERR_BAD_ARG, "gp_pairs: which must be 0 or 1");
    CK(cudaSetDevice(w->device));
    *n_out = 0;
    const uint32_t np = w->last_np;
    if (np == 0) return GP_OK;
    unsigned long long *k0 = nullptr, *k1 = nullptr;
    uint32_t *v0 = nullptr, *v1 = nullptr, *cnt = nullptr, *hist = nullptr;
    uint2* outp = nullptr;
    gp_status rs = GP_OK;
    do {
        if (dmalloc(&k0, np) || dmalloc(&k1, np) || dmalloc(&v0, np) || dmalloc(&v1, np) || dmalloc(&cnt, 1) ||
            dmalloc(&hist, (size_t)RS_RADIX * w->sm_count * 4) || dmalloc(&outp, np)) {
            rs = fail(w, GP_ERR_CUDA, "gp_pairs: out of device memory");
            break;
        }
        cudaMemsetAsync(cnt, 0, 4, w->st);
        k_export_pairs<<<cdiv(np, 256), 256, 0, w->st>>>(w->pairs, np, which, k0, v0, cnt);
        w->launches++;
        uint32_t m = 0;
        cudaMemcpyAsync(&m, cnt, 4, cudaMemcpyDeviceToHost, w->st);
        if (cudaStreamSynchronize(w->st) != cudaSuccess) {
            rs = fail(w, GP_ERR_CUDA, "gp_pairs: %s", cudaGetErrorString(cudaGetLastError()));
            break;
        }
        *n_out = m;
        if (m == 0) break;
        const int r = rs_sort<unsigned long long>(k0, v0, k1, v1, m, 63, false, hist, w->sm_count, w->st, &w->launches);
        k_gather_pairs<<<cdiv(m, 256), 256, 0, w->st>>>(w->pairs, r ? v1 : v0, m, Bodies{w->B[w->cur]}, outp);
        w->launches++;
        const uint64_t ncopy = std::min<uint64_t>(m, cap);
        if (pairs2 && ncopy) cudaMemcpyAsync(pairs2, outp, ncopy * sizeof(uint2), cudaMemcpyDeviceToHost, w->st);
        if (cudaStreamSynchronize(w->st) != cudaSuccess) {
            rs = fail(w, GP_ERR_CUDA, "gp_pairs: %s", cudaGetErrorString(cudaGetLastError()));
            break;
        }
    } while (0);
    cudaFree(k0), cudaFree(k1), cudaFree(v0), cudaFree(v1), cudaFree(cnt), cudaFree(hist), cudaFree(outp);
    return rs;
}

// ---- "next" rows ----
static gp_status next_rows_ready(gp_world* w, const char* what) {
    if (!w->uploaded) return fail(w, GP_ERR_STATE, "%s before gp_upload", what);
    if (w->multi) return fail(w, GP_ERR_STATE, "%s addresses bodies by upload index: not available on a slab world", what);
    CK(cudaSetDevice(w->device));
    return ensure_slot_map(w);
}

extern "C" gp_status gp_instances(gp_world* w, const float* rot4, const float* scale3, float* out25) {
    if (!w || !out25) return GP_ERR_BAD_ARG;
    gp_status s = next_rows_ready(w, "gp_instances");
    if (s != GP_OK) return s;
    const uint32_t n = w->n;
    if (n == 0) return GP_OK;
    const size_t b_rot = ((size_t)n * 16 + 255) & ~(size_t)255, b_sc = ((size_t)n * 12 + 255) & ~(size_t)255;
    const size_t b_out = (size_t)n * 100;
    s = ensure_stage(w, b_rot + b_sc + b_out + 256);
    if (s != GP_OK) return s;
    float* d_rot = (float*)w->stage;
    float* d_sc = (float*)(w->stage + b_rot);
    float* d_out = (float*)(w->stage + b_rot + b_sc);
    if (rot4) CK(cudaMemcpyAsync(d_rot, rot4, (size_t)n * 16, cudaMemcpyHostToDevice, w->st));
    if (scale3) CK(cudaMemcpyAsync(d_sc, scale3, (size_t)n * 12, cudaMemcpyHostToDevice, w->st));
    k_instances<<<cdiv(n, 128), 128, 0, w->st>>>(n, Bodies{w->B[w->cur]}, rot4 ? d_rot : nullptr, scale3 ? d_sc : nullptr,
                                                d_out);
    w->launches++;
    CK(cudaMemcpyAsync(out25, d_out, b_out, cudaMemcpyDeviceToHost, w->st));
    CK(cudaStreamSynchronize(w->st));
    return GP_OK;
}

extern "C" gp_status gp_raycast(gp_world* w, uint32_t nrays, const float* origin3, const float* dir3, uint32_t ntargets,
                                const uint32_t* targets, uint8_t* hits) {
    if (!w || !hits || (nrays && (!origin3 || !dir3))) return GP_ERR_BAD_ARG;
    gp_status s = next_rows_ready(w, "gp_raycast");
    if (s != GP_OK) return s;
    if (!targets) ntargets = w->n;
    for (uint32_t i = 0; targets && i < ntargets; ++i)
        if (targets[i] >= w->n) return fail(w, GP_ERR_BAD_ARG, "gp_raycast: targets[%u]=%u out of range", i, targets[i]);
    const size_t total = (size_t)nrays * ntargets;
    if (total == 0) return GP_OK;
    const size_t b_ray = ((size_t)nrays * 12 + 255) & ~(size_t)255, b_t = ((size_t)ntargets * 4 + 255) & ~(size_t)255;
    s = ensure_stage(w, 2 * b_ray + b_t + total + 256);
    if (s != GP_OK) return s;
    float* d_o = (float*)w->stage;
    float* d_d = (float*)(w->stage + b_ray);
    uint32_t* d_t = (uint32_t*)(w->stage + 2 * b_ray);
    uint8_t* d_h = (uint8_t*)(w->stage + 2 * b_ray + b_t);
    CK(cudaMemcpyAsync(d_o, origin3, (size_t)nrays * 12, cudaMemcpyHostToDevice, w->st));
    CK(cudaMemcpyAsync(d_d, dir3, (size_t)nrays * 12, cudaMemcpyHostToDevice, w->st));
    if (targets) CK(cudaMemcpyAsync(d_t, targets, (size_t)ntargets * 4, cudaMemcpyHostToDevice, w->st));
    k_raycast<<<cdiv(total, 256), 256, 0, w->st>>>(nrays, d_o, d_d, ntargets, targets ? d_t : nullptr, w->slot_of,
                                                  Bodies{w->B[w->cur]}, w->rawbox, d_h);
    w->launches++;
    CK(cudaMemcpyAsync(hits, d_h, total, cudaMemcpyDeviceToHost, w->st));
    CK(cudaStreamSynchronize(w->st));
    return GP_OK;
}

extern "C" gp_status gp_obstacle_grid(gp_world* w, uint32_t k, const uint32_t* idx, float cell_size,
                                      const int32_t origin[3], const uint32_t dims[3], uint32_t* bitmap,
                                      int32_t bounds_out[6]) {
    if (!w || !origin || !dims || !bitmap) return GP_ERR_BAD_ARG;
    gp_status s = next_rows_ready(w, "gp_obstacle_grid");
    if (s != GP_OK) return s;
    if (!idx) k = w->n;
    for (uint32_t i = 0; idx && i < k; ++i)
        if (idx[i] >= w->n) return fail(w, GP_ERR_BAD_ARG, "gp_obstacle_grid: idx[%u]=%u out of range", i, idx[i]);
    const unsigned long long cells = (unsigned long long)dims[0] * dims[1] * dims[2];
    if (cells == 0 || cells > (1ull << 36)) return fail(w, GP_ERR_BAD_ARG, "gp_obstacle_grid: bad dims");
    const size_t words = (size_t)((cells + 31) / 32);
    const size_t b_idx = ((size_t)k * 4 + 255) & ~(size_t)255;
    s = ensure_stage(w, b_idx + words * 4 + 256 + 64);
    if (s != GP_OK) return s;
    uint32_t* d_idx = (uint32_t*)w->stage;
    int* d_bounds = (int*)(w->stage + b_idx);
    uint32_t* d_bits = (uint32_t*)(w->stage + b_idx + 64);
    const int init[6] = {2147483647, 2147483647, 2147483647, -2147483647 - 1, -2147483647 - 1, -2147483647 - 1};
    CK(cudaMemcpyAsync(d_bounds, init, sizeof init, cudaMemcpyHostToDevice, w->st));
    CK(cudaMemsetAsync(d_bits, 0, words * 4, w->st));
    if (idx && k) CK(cudaMemcpyAsync(d_idx, idx, (size_t)k * 4, cudaMemcpyHostToDevice, w->st));
    if (k) {
        k_obstacle_grid<<<cdiv((uint64_t)k * 32, 256), 256, 0, w->st>>>(k, idx ? d_idx : nullptr, w->slot_of,
                                                                        Bodies{w->B[w->cur]}, w->rawbox, cell_size, origin[0],
                                                                        origin[1], origin[2], dims[0], dims[1], dims[2], d_bits,
                                                                        d_bounds);
        w->launches++;
    }
    CK(cudaMemcpyAsync(bitmap, d_bits, words * 4, cudaMemcpyDeviceToHost, w->st));
    if (bounds_out) CK(cudaMemcpyAsync(bounds_out, d_bounds, sizeof init, cudaMemcpyDeviceToHost, w->st));
    CK(cudaStreamSynchronize(w->st));
    return GP_OK;
}

// ---- test hooks ----
template <typename KeyT>
static gp_status test_sort(int32_t device, uint32_t n, uint32_t key_bits, KeyT* keys, uint32_t* vals) {
    gp_world* w = nullptr;
    if (cudaSetDevice(device) != cudaSuccess) return GP_ERR_NO_DEVICE;
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    KeyT *k0 = nullptr, *k1 = nullptr;
    uint32_t *v0 = nullptr, *v1 = nullptr, *hist = nullptr;
    CK(dmalloc(&k0, n));
    CK(dmalloc(&k1, n));
    CK(dmalloc(&v0, n));
    CK(dmalloc(&v1, n));
    CK(dmalloc(&hist, (size_t)RS_RADIX * prop.multiProcessorCount * 4));
    CK(cudaMemcpy(k0, keys, sizeof(KeyT) * (size_t)n, cudaMemcpyHostToDevice));
    if (vals) CK(cudaMemcpy(v0, vals, 4 * (size_t)n, cudaMemcpyHostToDevice));
    const int r = rs_sort<KeyT>(k0, v0, k1, v1, n, (int)key_bits, vals == nullptr, hist, prop.multiProcessorCount, 0, nullptr);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(keys, r ? k1 : k0, sizeof(KeyT) * (size_t)n, cudaMemcpyDeviceToHost));
    if (vals) CK(cudaMemcpy(vals, r ? v1 : v0, 4 * (size_t)n, cudaMemcpyDeviceToHost));
    cudaFree(k0), cudaFree(k1), cudaFree(v0), cudaFree(v1), cudaFree(hist);
    return GP_OK;
}

extern "C" gp_status gp_test_radix_sort(int32_t device, uint32_t n, uint32_t key_bits, uint32_t* keys_inout,
                                        uint32_t* vals_inout) {
    return test_sort<uint32_t>(device, n, key_bits, keys_inout, vals_inout);
}
extern "C" gp_status gp_test_radix_sort64(int32_t device, uint32_t n, uint32_t key_bits, uint64_t* keys_inout,
                                          uint32_t* vals_inout) {
    return test_sort<unsigned long long>(device, n, key_bits, (unsigned long long*)keys_inout, vals_inout);
}
extern "C" gp_status gp_test_exclusive_scan(int32_t device, uint32_t n, uint32_t* data_inout, uint32_t* total) {
    gp_world* w = nullptr;
    if (cudaSetDevice(device) != cudaSuccess) return GP_ERR_NO_DEVICE;
    uint32_t *in = nullptr, *out = nullptr, *ts = nullptr;
    CK(dmalloc(&in, n));
    CK(dmalloc(&out, (size_t)n + 1));
    CK(dmalloc(&ts, (size_t)cdiv(n, SC_TILE) + 1));
    CK(cudaMemcpy(in, data_inout, 4 * (size_t)n, cudaMemcpyHostToDevice));
    exclusive_scan_u32(in, n, out, ts, 0, nullptr);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(data_inout, out, 4 * (size_t)n, cudaMemcpyDeviceToHost));
    if (total) CK(cudaMemcpy(total, out + n, 4, cudaMemcpyDeviceToHost));
    cudaFree(in), cudaFree(out), cudaFree(ts);
    return GP_OK;
}

#include "gp_multi_impl.cuh"
