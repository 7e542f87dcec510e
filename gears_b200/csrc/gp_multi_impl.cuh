// gp_multi_impl.cuh — multi-GPU slabs along x (included at the end of gp_api.cu; sees gp_world).
//
// One gp_world per GPU.  A rank owns the bodies whose AABB centre x lies in its slab [x_lo, x_hi).  Per step, on the
// step stream and without any host synchronisation:
//   K1 (k_keys<true>)  integrates, packs (a) bodies whose centre left the slab -> ownership moves to the neighbour,
//                      (b) bodies whose inflated AABB comes within H of a face -> halo copies, into one fixed-size
//                      buffer per face (count in the header);
//   exchange           ncclSend/ncclRecv of the two buffers inside one ncclGroup (NVLink 5 / NVSwitch: both
//                      neighbours at full bandwidth), or peer copies when all slabs live in one process;
//   k_append           received records join the input set; the normal step follows (sort, pairs, sweep, repair);
//   verification       taint kernels decide which owned bodies are provably identical to the single-GPU result.
// The reference has no distributed runtime at all; this decomposition is new.
#pragma once
#include <dlfcn.h>

// NCCL is bound at run time (dlopen) so that single-GPU users of the library need neither libnccl nor nccl.h: the few
// types of its C API the slab transport touches are declared here (stable across NCCL 2.x).
extern "C" {
typedef struct ncclComm* ncclComm_t;
typedef struct {
    char internal[128];
} ncclUniqueId;
typedef enum { ncclSuccess = 0 } ncclResult_t;
typedef enum { ncclChar = 0, ncclInt32 = 2, ncclUint32 = 3 } ncclDataType_t;
typedef enum { ncclSum = 0, ncclProd = 1, ncclMax = 2, ncclMin = 3 } ncclRedOp_t;
}

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*CommAbort)(ncclComm_t) = nullptr;  // optional: tear-down without the peers (a dead slab world)
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi g_nccl;
static std::mutex g_nccl_mu;

static bool nccl_load(std::string* why) {
    std::lock_guard<std::mutex> lk(g_nccl_mu);
    if (g_nccl.handle) return true;
    // prefer a copy that is already in the process (e.g. the one torch.distributed loaded)
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) {
        if (why) *why = std::string("libnccl.so.2 not found: ") + dlerror();
        return false;
    }
    NcclApi a;
    a.handle = h;
#define GP_SYM(field, name)                                                      \
    *(void**)(&a.field) = dlsym(h, name);                                        \
    if (!a.field) {                                                              \
        if (why) *why = std::string("NCCL symbol missing: ") + name;             \
        return false;                                                            \
    }
    GP_SYM(GetUniqueId, "ncclGetUniqueId")
    GP_SYM(CommInitRank, "ncclCommInitRank")
    GP_SYM(CommDestroy, "ncclCommDestroy")
    *(void**)(&a.CommAbort) = dlsym(h, "ncclCommAbort");
    GP_SYM(GroupStart, "ncclGroupStart")
    GP_SYM(GroupEnd, "ncclGroupEnd")
    GP_SYM(Send, "ncclSend")
    GP_SYM(Recv, "ncclRecv")
    GP_SYM(AllReduce, "ncclAllReduce")
    GP_SYM(GetErrorString, "ncclGetErrorString")
#undef GP_SYM
    g_nccl = a;
    return true;
}

struct gp_multi {
    gp_world* w = nullptr;
    gp_slab_config cfg{};
    bool local = false;            // all slabs in this process: peer copies instead of NCCL
    // peer-to-peer transport (one process per GPU): K1 writes into the neighbour's receive buffer over NVLink
    bool p2p = false;
    unsigned char* arena = nullptr;          // [256 B flags][recv L parity 0][recv L 1][recv R 0][recv R 1]
    size_t arena_bytes = 0;
    size_t off_recv[2][2] = {{0, 0}, {0, 0}};
    unsigned char* peer_arena[2] = {nullptr, nullptr};     // the neighbours' arenas, opened through CUDA IPC
    size_t peer_off_recv[2][2] = {{0, 0}, {0, 0}};         // where MY records go in neighbour s's arena, per parity
    uint32_t step_id = 0;
    ncclComm_t comm = nullptr;
    gp_world* peer[2] = {nullptr, nullptr};  // local transport: left / right neighbour world
    float4* send[2] = {nullptr, nullptr};
    float4* recv[2] = {nullptr, nullptr};
    size_t buf_bytes[2] = {0, 0};  // per face: 16 B header + xcap[side] records (both sides of a face agree)
    uint32_t proposal = 0;         // records per face buffer this rank asks for
    uint32_t* taint_bits = nullptr;
    uint32_t* seed_bits = nullptr;  // marked by K3, consumed (and cleared) by the proof
    uint32_t* uf_parent = nullptr;  // the proof's union-find forest (reset by K3)
    uint32_t* scratch = nullptr;   // [0] taint gate, [1..3] prepare counts, [4] compact counter
    cudaEvent_t ev_packed = nullptr;   // local transport: this world's send buffers are ready
    cudaEvent_t ev_copied = nullptr;   // local transport: this world has copied what it needs from its neighbours
    bool verify = true;
    unsigned long long timeout_ns = 30ull * 1000000000ull;  // peer-to-peer waits (GP_PEER_TIMEOUT_MS)
    uint32_t* rb_buf = nullptr;          // re-balancing scratch: [0..1] min/max, [2..2+64) faces, [128..128+RB_BINS) histogram
    uint32_t owned_count = 0xffffffffu;  // rows of the last gp_download_owned ...
    uint64_t owned_gen = 0;              // ... and the step count it was taken at (gp_set_state_owned checks both)
};

static void gp_multi_destroy(gp_multi* m) {
    if (!m) return;
    for (int s = 0; s < 2; ++s)
        if (m->peer_arena[s]) cudaIpcCloseMemHandle(m->peer_arena[s]);
    if (m->arena) cudaFree(m->arena);
    // (ncclCommDestroy expects every rank to come along; a world whose neighbour timed out is on its own)
    if (m->comm && m->w && m->w->dead && g_nccl.CommAbort)
        g_nccl.CommAbort(m->comm);
    else if (m->comm && g_nccl.CommDestroy)
        g_nccl.CommDestroy(m->comm);
    for (int s = 0; s < 2; ++s) {
        if (m->send[s]) cudaFree(m->send[s]);
        if (m->recv[s]) cudaFree(m->recv[s]);
    }
    if (m->taint_bits) cudaFree(m->taint_bits);
    if (m->seed_bits) cudaFree(m->seed_bits);
    if (m->uf_parent) cudaFree(m->uf_parent);
    if (m->rb_buf) cudaFree(m->rb_buf);
    if (m->scratch) cudaFree(m->scratch);
    if (m->ev_packed) cudaEventDestroy(m->ev_packed);
    if (m->ev_copied) cudaEventDestroy(m->ev_copied);
    delete m;
}
static bool gp_multi_is_local(gp_multi* m) { return m->local; }
// flag words in an arena: ready[side][parity] = words side*2+parity, credit[side] = words 4+side
static inline uint32_t* arena_flag(unsigned char* arena, int word) { return arena ? (uint32_t*)arena + word : nullptr; }
static float4* gp_multi_send(gp_multi* m, int side) {
    if (!m->p2p) return m->send[side];
    return m->peer_arena[side] ? (float4*)(m->peer_arena[side] + m->peer_off_recv[side][m->step_id & 1u]) : nullptr;
}
static float4* gp_multi_recv(gp_multi* m, int side) {
    if (!m->p2p) return m->recv[side];
    return m->peer_arena[side] ? (float4*)(m->arena + m->off_recv[side][m->step_id & 1u]) : nullptr;
}
static void gp_multi_begin_step(gp_multi* m) {
    gp_world* w = m->w;
    m->step_id++;
    if (!m->p2p || m->step_id <= 2) return;
    // the buffer of parity step_id&1 was last filled for step_id-2: the neighbours must have consumed that
    k_wait_flags<<<1, 2, 0, w->st>>>(m->peer_arena[0] ? arena_flag(m->arena, 4) : nullptr,
                                     m->peer_arena[1] ? arena_flag(m->arena, 5) : nullptr, m->step_id - 2, w->ctl,
                                     m->timeout_ns);
    w->launches++;
}
static void gp_multi_seal(gp_multi* m) {
    gp_world* w = m->w;
    if (m->p2p) {
        const uint32_t par = m->step_id & 1u;
        // I am my left neighbour's RIGHT side (1) and my right neighbour's LEFT side (0)
        k_seal_p2p<<<1, 1, 0, w->st>>>(gp_multi_send(m, 0), gp_multi_send(m, 1), arena_flag(m->peer_arena[0], 1 * 2 + par),
                                       arena_flag(m->peer_arena[1], 0 * 2 + par), m->step_id, w->ctl);
    } else {
        k_seal_sends<<<1, 1, 0, w->st>>>(m->send[0], m->send[1], w->ctl);
    }
    w->launches++;
}
static void gp_multi_before_append(gp_multi* m) {
    gp_world* w = m->w;
    if (!m->p2p) return;
    const uint32_t par = m->step_id & 1u;
    k_wait_flags<<<1, 2, 0, w->st>>>(m->peer_arena[0] ? arena_flag(m->arena, 0 * 2 + par) : nullptr,
                                     m->peer_arena[1] ? arena_flag(m->arena, 1 * 2 + par) : nullptr, m->step_id, w->ctl,
                                     m->timeout_ns);
    w->launches++;
}
static void gp_multi_after_append(gp_multi* m) {
    gp_world* w = m->w;
    if (!m->p2p) return;
    k_post_flags<<<1, 1, 0, w->st>>>(arena_flag(m->peer_arena[0], 4 + 1), arena_flag(m->peer_arena[1], 4 + 0), m->step_id);
    w->launches++;
}

#define CKN(call)                                                                                             \
    do {                                                                                                      \
        ncclResult_t r_ = (call);                                                                             \
        if (r_ != ncclSuccess)                                                                                \
            return fail(w, GP_ERR_NCCL, "%s failed: %s (%s:%d)", #call, g_nccl.GetErrorString(r_), __FILE__, __LINE__); \
    } while (0)

static gp_status gp_multi_enqueue_exchange(gp_multi* m) {
    gp_world* w = m->w;
    if (m->local || m->p2p) return GP_OK;  // gp_step_group drives the copies / K1 already wrote into the neighbour
    const int rank = m->cfg.rank;
    CKN(g_nccl.GroupStart());
    // fixed-size messages (the count travels in the header): no host round trip to size them
    if (m->send[0]) CKN(g_nccl.Send(m->send[0], m->buf_bytes[0], ncclChar, rank - 1, m->comm, w->st));
    if (m->send[1]) CKN(g_nccl.Send(m->send[1], m->buf_bytes[1], ncclChar, rank + 1, m->comm, w->st));
    if (m->recv[0]) CKN(g_nccl.Recv(m->recv[0], m->buf_bytes[0], ncclChar, rank - 1, m->comm, w->st));
    if (m->recv[1]) CKN(g_nccl.Recv(m->recv[1], m->buf_bytes[1], ncclChar, rank + 1, m->comm, w->st));
    CKN(g_nccl.GroupEnd());
    return GP_OK;
}

static uint32_t* gp_multi_taint_bits(gp_multi* m) { return m->verify ? m->taint_bits : nullptr; }
static uint32_t* gp_multi_seed_bits(gp_multi* m) { return m->verify ? m->seed_bits : nullptr; }
static uint32_t* gp_multi_uf_parent(gp_multi* m) { return m->verify ? m->uf_parent : nullptr; }

// common part of gp_comm_init / gp_comm_init_local
static gp_status slab_setup(gp_world* w, const gp_slab_config* cfg, bool local) {
    if (!w->uploaded) return fail(w, GP_ERR_STATE, "gp_comm_init before gp_upload");
    if (w->multi) return fail(w, GP_ERR_STATE, "this world already belongs to a slab decomposition");
    if (!cfg || cfg->struct_size == 0 || cfg->struct_size > sizeof(gp_slab_config))
        return fail(w, GP_ERR_BAD_ARG, "bad gp_slab_config");
    gp_slab_config c{};
    memcpy(&c, cfg, cfg->struct_size);
    if (c.nranks < 1 || c.rank < 0 || c.rank >= c.nranks) return fail(w, GP_ERR_BAD_ARG, "bad rank %d of %d", c.rank, c.nranks);
    if (!(c.x_lo < c.x_hi)) return fail(w, GP_ERR_BAD_ARG, "empty slab [%g, %g)", c.x_lo, c.x_hi);
    CK(cudaSetDevice(w->device));
    CK(cudaStreamSynchronize(w->st));
    CK(cudaMemcpy(w->h_ctl, w->ctl, sizeof(Ctl), cudaMemcpyDeviceToHost));
    Ctl hc = *w->h_ctl;
    // coverage: four (largest grid-class extents) unless the caller says otherwise: about three layers of contact
    // neighbours at the usual margins.  Too small a halo is not silent: the verification counts what it cannot prove.
    if (!(c.halo > 0)) c.halo = 4.0f * hc.ext_max;
    if (c.nranks > 1 && c.x_hi - c.x_lo < 2.0f * c.halo && c.rank > 0 && c.rank + 1 < c.nranks)
        return fail(w, GP_ERR_BAD_ARG, "slab width %g is below twice the halo %g: a body could reach a rank two slabs away",
                    c.x_hi - c.x_lo, c.halo);
    gp_multi* m = new gp_multi();
    m->w = w;
    m->cfg = c;
    m->local = local;
    m->verify = !(c.flags & GP_SLAB_NO_VERIFY);
    if (const char* e = getenv("GP_PEER_TIMEOUT_MS")) m->timeout_ns = (unsigned long long)std::max(1.0, atof(e)) * 1000000ull;
    hc.x_lo = c.x_lo, hc.x_hi = c.x_hi, hc.halo = c.halo;
    hc.has_left = c.rank > 0, hc.has_right = c.rank + 1 < c.nranks;
    hc.cap = w->cap;
    hc.xcap[0] = hc.xcap[1] = 0;
    gp_status rs = GP_OK;
    do {
        if (dmalloc(&m->scratch, 16) || dmalloc(&m->taint_bits, (size_t)w->cap / 32 + 1) ||
            dmalloc(&m->seed_bits, (size_t)w->cap / 32 + 1) || dmalloc(&m->uf_parent, (size_t)w->cap) ||
            cudaMemsetAsync(m->seed_bits, 0, sizeof(uint32_t) * ((size_t)w->cap / 32 + 1), w->st) != cudaSuccess ||
            cudaEventCreateWithFlags(&m->ev_packed, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&m->ev_copied, cudaEventDisableTiming) != cudaSuccess) {
            rs = fail(w, GP_ERR_CUDA, "slab setup: out of device memory");
            break;
        }
        cudaMemsetAsync(m->scratch, 0, 16 * sizeof(uint32_t), w->st);
        cudaMemcpyAsync(w->ctl, &hc, sizeof hc, cudaMemcpyHostToDevice, w->st);
        // body ids -> global entity ids; how many records would the faces send today?
        const uint32_t* d_ent = w->have_entity ? (const uint32_t*)w->entity_dev : nullptr;
        if (w->n) k_slab_prepare<<<cdiv(w->n, 256), 256, 0, w->st>>>(w->n, Bodies{w->B[w->cur]}, w->A[w->cur], d_ent, m->scratch + 1, w->ctl);
        w->launches++;
        uint32_t counts[3] = {0, 0, 0};
        if (cudaMemcpyAsync(counts, m->scratch + 1, sizeof counts, cudaMemcpyDeviceToHost, w->st) != cudaSuccess ||
            cudaStreamSynchronize(w->st) != cudaSuccess) {
            rs = fail(w, GP_ERR_CUDA, "slab setup: %s", cudaGetErrorString(cudaGetLastError()));
            break;
        }
        if (counts[2]) {
            rs = fail(w, GP_ERR_BAD_ARG, "%u dynamic bodies are larger than the grid cell: big bodies must be static "
                      "(they are replicated on every rank)", counts[2]);
            break;
        }
        m->proposal = c.max_exchange ? c.max_exchange : 2u * std::max(counts[0], counts[1]) + 8192u;
        if (cudaStreamSynchronize(w->st) != cudaSuccess) rs = fail(w, GP_ERR_CUDA, "slab setup failed");
    } while (0);
    if (rs != GP_OK) {
        gp_multi_destroy(m);
        return rs;
    }
    w->multi = m;
    return GP_OK;
}

// both sides of a face have agreed on cap[side] records: allocate the buffers, publish the capacities
static gp_status slab_alloc_buffers(gp_world* w, const uint32_t cap[2]) {
    gp_multi* m = w->multi;
    CK(cudaSetDevice(w->device));
    const bool has[2] = {m->cfg.rank > 0, m->cfg.rank + 1 < m->cfg.nranks};
    uint32_t xc[2] = {0, 0};
    for (int s = 0; s < 2; ++s) {
        if (!has[s]) continue;
        xc[s] = cap[s];
        m->buf_bytes[s] = sizeof(float4) * (1 + (size_t)cap[s] * XREC);
        CK(dmalloc(&m->send[s], m->buf_bytes[s] / sizeof(float4)));
        CK(dmalloc(&m->recv[s], m->buf_bytes[s] / sizeof(float4)));
        CK(cudaMemsetAsync(m->send[s], 0, sizeof(float4), w->st));
        CK(cudaMemsetAsync(m->recv[s], 0, sizeof(float4), w->st));
    }
    Ctl* dc = w->ctl;
    CK(cudaMemcpyAsync(&dc->xcap[0], xc, sizeof xc, cudaMemcpyHostToDevice, w->st));
    CK(cudaStreamSynchronize(w->st));
    return GP_OK;
}

// Peer-to-peer transport: one arena per world (flags + double-buffered receive buffers), shared with the two
// neighbours through CUDA IPC; the handles travel over the NCCL communicator once, at set-up.
struct ArenaInfo {
    cudaIpcMemHandle_t handle;
    unsigned long long off_recv[2][2];
    unsigned long long valid;
};
static gp_status slab_setup_p2p(gp_world* w) {
    gp_multi* m = w->multi;
    const int rank = m->cfg.rank, nr = m->cfg.nranks;
    size_t off = 256;
    for (int s = 0; s < 2; ++s)
        for (int par = 0; par < 2; ++par) {
            m->off_recv[s][par] = off;
            off += (m->buf_bytes[s] + 255) & ~(size_t)255;
        }
    m->arena_bytes = off;
    CK(cudaMalloc((void**)&m->arena, m->arena_bytes));
    CK(cudaMemset(m->arena, 0, m->arena_bytes));
    ArenaInfo mine{}, theirs[2]{};
    CK(cudaIpcGetMemHandle(&mine.handle, m->arena));
    for (int s = 0; s < 2; ++s)
        for (int par = 0; par < 2; ++par) mine.off_recv[s][par] = m->off_recv[s][par];
    mine.valid = 1;
    ArenaInfo* d = nullptr;  // [0] mine, [1] left's, [2] right's
    CK(cudaMalloc((void**)&d, 3 * sizeof(ArenaInfo)));
    CK(cudaMemset(d, 0, 3 * sizeof(ArenaInfo)));
    CK(cudaMemcpyAsync(d, &mine, sizeof mine, cudaMemcpyHostToDevice, w->st));
    CKN(g_nccl.GroupStart());
    if (rank > 0) CKN(g_nccl.Send(d, sizeof(ArenaInfo), ncclChar, rank - 1, m->comm, w->st));
    if (rank + 1 < nr) CKN(g_nccl.Send(d, sizeof(ArenaInfo), ncclChar, rank + 1, m->comm, w->st));
    if (rank > 0) CKN(g_nccl.Recv(d + 1, sizeof(ArenaInfo), ncclChar, rank - 1, m->comm, w->st));
    if (rank + 1 < nr) CKN(g_nccl.Recv(d + 2, sizeof(ArenaInfo), ncclChar, rank + 1, m->comm, w->st));
    CKN(g_nccl.GroupEnd());
    CK(cudaMemcpyAsync(theirs, d + 1, 2 * sizeof(ArenaInfo), cudaMemcpyDeviceToHost, w->st));
    CK(cudaStreamSynchronize(w->st));
    cudaFree(d);
    bool ok = true;
    for (int s = 0; s < 2; ++s) {
        if (!theirs[s].valid) continue;  // no neighbour on that side
        void* p = nullptr;
        if (cudaIpcOpenMemHandle(&p, theirs[s].handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
            cudaGetLastError();
            ok = false;
            break;
        }
        m->peer_arena[s] = (unsigned char*)p;
        // on neighbour s I am its side 1-s
        for (int par = 0; par < 2; ++par) m->peer_off_recv[s][par] = (size_t)theirs[s].off_recv[1 - s][par];
    }
    // every rank must take the same decision (a face cannot mix transports)
    {
        int* dflag = nullptr;
        CK(cudaMalloc((void**)&dflag, sizeof(int)));
        const int mine_ok = ok ? 1 : 0;
        int all_ok = 0;
        CK(cudaMemcpyAsync(dflag, &mine_ok, sizeof(int), cudaMemcpyHostToDevice, w->st));
        CKN(g_nccl.AllReduce(dflag, dflag, 1, ncclInt32, ncclMin, m->comm, w->st));
        CK(cudaMemcpyAsync(&all_ok, dflag, sizeof(int), cudaMemcpyDeviceToHost, w->st));
        CK(cudaStreamSynchronize(w->st));
        cudaFree(dflag);
        ok = all_ok != 0;
    }
    if (!ok) {  // no peer access between some devices: stay on the NCCL exchange
        for (int s = 0; s < 2; ++s)
            if (m->peer_arena[s]) cudaIpcCloseMemHandle(m->peer_arena[s]), m->peer_arena[s] = nullptr;
        cudaFree(m->arena);
        m->arena = nullptr;
        return GP_OK;
    }
    m->p2p = true;
    return GP_OK;
}

extern "C" gp_status gp_comm_unique_id(uint8_t id[GP_UNIQUE_ID_BYTES]) {
    static_assert(sizeof(ncclUniqueId) <= GP_UNIQUE_ID_BYTES, "unique id does not fit");
    memset(id, 0, GP_UNIQUE_ID_BYTES);
    std::string why;
    if (!nccl_load(&why)) {
        fail(nullptr, GP_ERR_NCCL, "%s", why.c_str());
        return GP_ERR_NCCL;
    }
    ncclUniqueId uid;
    if (g_nccl.GetUniqueId(&uid) != ncclSuccess) return GP_ERR_NCCL;
    memcpy(id, &uid, sizeof uid);
    return GP_OK;
}

extern "C" gp_status gp_comm_init(gp_world* w, const uint8_t id[GP_UNIQUE_ID_BYTES], const gp_slab_config* cfg) {
    if (!w || !id) return GP_ERR_BAD_ARG;
    std::string why;
    if (!nccl_load(&why)) return fail(w, GP_ERR_NCCL, "%s", why.c_str());
    gp_status s = slab_setup(w, cfg, false);
    if (s != GP_OK) return s;
    ncclUniqueId uid;
    memcpy(&uid, id, sizeof uid);
    CK(cudaSetDevice(w->device));
    ncclResult_t r = g_nccl.CommInitRank(&w->multi->comm, w->multi->cfg.nranks, uid, w->multi->cfg.rank);
    if (r != ncclSuccess) {
        gp_multi_destroy(w->multi);
        w->multi = nullptr;
        return fail(w, GP_ERR_NCCL, "ncclCommInitRank failed: %s", g_nccl.GetErrorString(r));
    }
    // agree with each neighbour on the message size of the shared face: the larger of the two proposals
    gp_multi* m = w->multi;
    const int rank = m->cfg.rank, nr = m->cfg.nranks;
    uint32_t* sc = m->scratch;  // [8] mine, [9] left's, [10] right's
    uint32_t mine[3] = {m->proposal, 0, 0};
    CK(cudaMemcpyAsync(sc + 8, mine, sizeof mine, cudaMemcpyHostToDevice, w->st));
    CKN(g_nccl.GroupStart());
    if (rank > 0) CKN(g_nccl.Send(sc + 8, 4, ncclChar, rank - 1, m->comm, w->st));
    if (rank + 1 < nr) CKN(g_nccl.Send(sc + 8, 4, ncclChar, rank + 1, m->comm, w->st));
    if (rank > 0) CKN(g_nccl.Recv(sc + 9, 4, ncclChar, rank - 1, m->comm, w->st));
    if (rank + 1 < nr) CKN(g_nccl.Recv(sc + 10, 4, ncclChar, rank + 1, m->comm, w->st));
    CKN(g_nccl.GroupEnd());
    uint32_t got[3];
    CK(cudaMemcpyAsync(got, sc + 8, sizeof got, cudaMemcpyDeviceToHost, w->st));
    CK(cudaStreamSynchronize(w->st));
    const uint32_t cap[2] = {std::max(got[0], got[1]), std::max(got[0], got[2])};
    gp_status sb = slab_alloc_buffers(w, cap);
    if (sb != GP_OK) return sb;
    const char* tr = getenv("GP_SLAB_TRANSPORT");
    if (tr && !strcmp(tr, "nccl")) return GP_OK;  // keep ncclSend/ncclRecv for the per-step exchange
    return slab_setup_p2p(w);
}

extern "C" gp_status gp_comm_init_local(gp_world* const* worlds, int32_t n, const gp_slab_config* cfgs) {
    if (!worlds || !cfgs || n < 1) return GP_ERR_BAD_ARG;
    for (int i = 0; i < n; ++i) {
        if (!worlds[i]) return GP_ERR_BAD_ARG;
        if (cfgs[i].rank != i || cfgs[i].nranks != n)
            return fail(worlds[i], GP_ERR_BAD_ARG, "gp_comm_init_local: cfgs[%d] must have rank %d of %d", i, i, n);
    }
    for (int i = 0; i < n; ++i) {
        gp_status s = slab_setup(worlds[i], &cfgs[i], true);
        if (s != GP_OK) {
            for (int j = 0; j < i; ++j) gp_multi_destroy(worlds[j]->multi), worlds[j]->multi = nullptr;
            return s;
        }
    }
    for (int i = 0; i < n; ++i) {
        gp_multi* m = worlds[i]->multi;
        m->peer[0] = i > 0 ? worlds[i - 1] : nullptr;
        m->peer[1] = i + 1 < n ? worlds[i + 1] : nullptr;
        // message size of a face = the larger of the two sides' proposals
        const uint32_t cap[2] = {m->peer[0] ? std::max(m->proposal, m->peer[0]->multi->proposal) : 0u,
                                 m->peer[1] ? std::max(m->proposal, m->peer[1]->multi->proposal) : 0u};
        gp_status s = slab_alloc_buffers(worlds[i], cap);
        if (s != GP_OK) return s;
    }
    return GP_OK;
}

// In-process slabs: all worlds advance in lock step; the exchange is a peer copy ordered by events.
extern "C" gp_status gp_step_group(gp_world* const* worlds, int32_t n, float dt, const float* damp3, uint32_t nsteps) {
    if (!worlds || n < 1) return GP_ERR_BAD_ARG;
    for (int i = 0; i < n; ++i)
        if (!worlds[i] || !worlds[i]->multi || !worlds[i]->multi->local)
            return worlds[i] ? fail(worlds[i], GP_ERR_STATE, "gp_step_group needs worlds set up by gp_comm_init_local") : GP_ERR_BAD_ARG;
    for (uint32_t s = 0; s < nsteps; ++s) {
        for (int i = 0; i < n; ++i) {
            gp_world* w = worlds[i];
            gp_multi* m = w->multi;
            CK(cudaSetDevice(w->device));
            if (s == 0) CK(cudaEventRecord(w->ev0, w->st));
            // the neighbours must have finished copying out of my send buffers (previous step) before I refill them
            for (int side = 0; side < 2; ++side)
                if (m->peer[side]) CK(cudaStreamWaitEvent(w->st, m->peer[side]->multi->ev_copied, 0));
            gp_status r = enqueue_phase_a(w, dt, damp3);
            if (r != GP_OK) return r;
            CK(cudaEventRecord(m->ev_packed, w->st));
        }
        for (int i = 0; i < n; ++i) {
            gp_world* w = worlds[i];
            gp_multi* m = w->multi;
            CK(cudaSetDevice(w->device));
            for (int side = 0; side < 2; ++side) {
                gp_world* p = m->peer[side];
                if (!p) continue;
                CK(cudaStreamWaitEvent(w->st, p->multi->ev_packed, 0));
                // my left neighbour's RIGHT send buffer is my left receive buffer, and vice versa
                CK(cudaMemcpyPeerAsync(m->recv[side], w->device, p->multi->send[side ^ 1], p->device, m->buf_bytes[side], w->st));
            }
            CK(cudaEventRecord(m->ev_copied, w->st));
        }
        for (int i = 0; i < n; ++i) {
            gp_world* w = worlds[i];
            CK(cudaSetDevice(w->device));
            gp_status r = enqueue_phase_b(w, dt, damp3, 1, REPAIR_ROUNDS_ENQUEUED);
            if (r != GP_OK) return r;
            w->cur ^= 1;
            w->steps_total++;
            if (s + 1 == nsteps) {
                CK(cudaEventRecord(w->ev1, w->st));
                w->timed = true;
            }
        }
    }
    return GP_OK;
}

// flags + scan of the owned bodies of a slab world (chain_cnt / offs / tile_sums are free between steps); *n = count
static gp_status scan_owned(gp_world* w, uint32_t* n) {
    gp_multi* m = w->multi;
    const uint32_t cap = w->cap;
    k_owned_flags<<<cdiv(cap, 256), 256, 0, w->st>>>(cap, w->ctl, w->A[w->cur], m->cfg.rank == 0, w->chain_cnt);
    w->launches++;
    exclusive_scan_u32(w->chain_cnt, cap, w->offs, w->tile_sums, w->st, &w->launches);
    CK(cudaMemcpyAsync(n, w->offs + cap, sizeof(uint32_t), cudaMemcpyDeviceToHost, w->st));
    CK(cudaStreamSynchronize(w->st));
    return GP_OK;
}

extern "C" gp_status gp_owned_count(gp_world* w, uint32_t* n) {
    if (!w || !n) return GP_ERR_BAD_ARG;
    if (!w->uploaded) return fail(w, GP_ERR_STATE, "no bodies uploaded");
    if (!w->multi) {
        *n = w->n;
        return GP_OK;
    }
    CK(cudaSetDevice(w->device));
    return scan_owned(w, n);
}

// entity ids + state of the bodies this rank owns now (after migration), in slot order
extern "C" gp_status gp_download_owned(gp_world* w, uint32_t* entity, float* pos3, float* vel3, uint8_t* unverified) {
    if (!w || !entity || !pos3 || !vel3) return GP_ERR_BAD_ARG;
    if (!w->uploaded) return fail(w, GP_ERR_STATE, "no bodies uploaded");
    CK(cudaSetDevice(w->device));
    uint32_t n = 0;
    if (w->multi) {
        gp_status s = scan_owned(w, &n);
        if (s != GP_OK) return s;
    } else {
        n = w->n;
    }
    if (n == 0) return gp_sync(w, nullptr);
    const size_t b3 = 3 * sizeof(float) * (size_t)n, a3 = (b3 + 255) & ~(size_t)255, bi = ((size_t)n * 4 + 255) & ~(size_t)255;
    gp_status s = ensure_stage(w, 2 * a3 + 2 * bi + 256);
    if (s != GP_OK) return s;
    float* dp = (float*)w->stage;
    float* dv = (float*)(w->stage + a3);
    uint32_t* di = (uint32_t*)(w->stage + 2 * a3);
    uint8_t* du = (uint8_t*)(w->stage + 2 * a3 + bi);
    CK(cudaMemsetAsync(du, 0, n, w->st));
    if (w->multi) {
        k_owned_gather<<<cdiv(w->cap, 256), 256, 0, w->st>>>(
            w->cap, w->ctl, Bodies{w->B[w->cur]}, w->chain_cnt, w->offs, di, dp, dv,
            w->multi->verify && w->steps_total ? w->multi->taint_bits : nullptr, du, w->dlist);
        w->multi->owned_count = n;
        w->multi->owned_gen = w->steps_total;
    } else {
        k_unpack<<<cdiv(n, 256), 256, 0, w->st>>>(n, Bodies{w->B[w->cur]}, dp, dv, di, 0);
    }
    w->launches++;
    CK(cudaMemcpyAsync(pos3, dp, b3, cudaMemcpyDeviceToHost, w->st));
    CK(cudaMemcpyAsync(vel3, dv, b3, cudaMemcpyDeviceToHost, w->st));
    CK(cudaMemcpyAsync(entity, di, (size_t)n * 4, cudaMemcpyDeviceToHost, w->st));
    if (unverified) CK(cudaMemcpyAsync(unverified, du, n, cudaMemcpyDeviceToHost, w->st));
    CK(cudaStreamSynchronize(w->st));
    return gp_sync(w, nullptr);
}

// The rows gp_download_owned returned last, written back (external systems changed them): row t = the t-th owned body
// of that download.  Valid only while no step has run since (the slots change with every step).
extern "C" gp_status gp_set_state_owned(gp_world* w, uint32_t k, const float* pos3, const float* vel3, const float* acc3) {
    if (!w) return GP_ERR_BAD_ARG;
    if (!w->uploaded) return fail(w, GP_ERR_STATE, "no bodies uploaded");
    if (!w->multi) return fail(w, GP_ERR_STATE, "gp_set_state_owned is for slab worlds; use gp_set_state");
    gp_multi* m = w->multi;
    if (m->owned_gen != w->steps_total || m->owned_count == 0xffffffffu)
        return fail(w, GP_ERR_STATE, "gp_set_state_owned: call gp_download_owned first (and no step in between)");
    if (k != m->owned_count) return fail(w, GP_ERR_BAD_ARG, "gp_set_state_owned: %u rows, the last download had %u", k, m->owned_count);
    CK(cudaSetDevice(w->device));
    if (k == 0) return GP_OK;
    const size_t b3 = 3 * sizeof(float) * (size_t)k, a3 = (b3 + 255) & ~(size_t)255;
    gp_status s = ensure_stage(w, 3 * a3 + 256);
    if (s != GP_OK) return s;
    unsigned char* sg = w->stage;
    if (pos3) CK(cudaMemcpyAsync(sg, pos3, b3, cudaMemcpyHostToDevice, w->st));
    if (vel3) CK(cudaMemcpyAsync(sg + a3, vel3, b3, cudaMemcpyHostToDevice, w->st));
    if (acc3) CK(cudaMemcpyAsync(sg + 2 * a3, acc3, b3, cudaMemcpyHostToDevice, w->st));
    k_set_state_slots<<<cdiv(k, 256), 256, 0, w->st>>>(k, w->dlist, pos3 ? (const float*)sg : nullptr,
                                                        vel3 ? (const float*)(sg + a3) : nullptr,
                                                        acc3 ? (const float*)(sg + 2 * a3) : nullptr, Bodies{w->B[w->cur]},
                                                        w->A[w->cur]);
    w->launches++;
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(w->st));  // the host arrays may be pageable and short-lived
    return GP_OK;
}

// ---------------------------------------------------------------------------------------------
// Re-balancing of the slab faces ("re-balanced every K steps").  Between two steps: histogram of
// the owned bodies' centre x over the global x range, summed over the ranks; new faces = equal-count quantiles; a
// face moves by at most max_shift per call, so that what changes owner in the next step is no more than a halo's
// worth of bodies (they travel through the ordinary migration path of K1, which hands a body to the ADJACENT rank).
// Every rank computes every face from the same summed histogram, so both sides of a face get the same bits.
// ---------------------------------------------------------------------------------------------
static const int RB_MAX_RANKS = 64;

static gp_status rb_alloc(gp_world* w) {
    gp_multi* m = w->multi;
    if (!m->rb_buf) CK(dmalloc(&m->rb_buf, 128 + RB_BINS));
    return GP_OK;
}
static void rb_enqueue_minmax(gp_world* w) {
    gp_multi* m = w->multi;
    const uint32_t init[2] = {0xffffffffu, 0u};
    cudaMemcpyAsync(m->rb_buf, init, sizeof init, cudaMemcpyHostToDevice, w->st);
    k_owned_minmax<<<cdiv(w->cap, 256), 256, 0, w->st>>>(w->cap, w->ctl, Bodies{w->B[w->cur]}, w->A[w->cur], m->rb_buf);
    w->launches++;
}
static void rb_enqueue_hist(gp_world* w, float x0, float inv_bin) {
    gp_multi* m = w->multi;
    cudaMemsetAsync(m->rb_buf + 128, 0, RB_BINS * sizeof(uint32_t), w->st);
    k_owned_xhist<<<w->sm_count * 2, 256, 0, w->st>>>(w->cap, w->ctl, Bodies{w->B[w->cur]}, w->A[w->cur], x0, inv_bin,
                                                       m->rb_buf + 128);
    w->launches++;
}
// faces[0..nr]: faces[0] = -inf, faces[nr] = +inf.  Returns false (faces untouched) when a slab would get too narrow.
static bool rb_new_faces(const uint32_t* hist, float x0, float binw, int nr, float* faces, float max_shift, float min_width) {
    unsigned long long total = 0;
    for (uint32_t b = 0; b < RB_BINS; ++b) total += hist[b];
    if (total == 0) return false;
    std::vector<float> out(faces, faces + nr + 1);
    unsigned long long cum = 0;
    uint32_t b = 0;
    for (int r = 1; r < nr; ++r) {
        const double target = (double)total * r / nr;
        while (b < RB_BINS && (double)(cum + hist[b]) < target) cum += hist[b++];
        const double frac = (b < RB_BINS && hist[b]) ? (target - (double)cum) / (double)hist[b] : 0.0;
        float f = x0 + (float)(((double)b + frac) * (double)binw);
        f = std::min(std::max(f, faces[r] - max_shift), faces[r] + max_shift);
        out[r] = f;
    }
    for (int r = 1; r < nr; ++r) {
        const float left = r == 1 ? -INFINITY : out[r - 1];
        if (!(out[r] - left >= min_width)) return false;
    }
    for (int r = 1; r < nr; ++r) faces[r] = out[r];
    return true;
}
static gp_status rb_apply(gp_world* w, float x_lo, float x_hi) {
    gp_multi* m = w->multi;
    m->cfg.x_lo = x_lo, m->cfg.x_hi = x_hi;
    const float v[2] = {x_lo, x_hi};
    CK(cudaMemcpyAsync(&w->ctl->x_lo, v, sizeof v, cudaMemcpyHostToDevice, w->st));  // (x_lo, x_hi are adjacent in Ctl)
    CK(cudaStreamSynchronize(w->st));
    return GP_OK;
}

extern "C" gp_status gp_slab_rebalance(gp_world* w, float max_shift, uint32_t* moved_out) {
    if (!w) return GP_ERR_BAD_ARG;
    if (moved_out) *moved_out = 0;
    if (!w->multi || w->multi->local) return fail(w, GP_ERR_STATE, "gp_slab_rebalance needs a world attached with gp_comm_init");
    gp_multi* m = w->multi;
    const int nr = m->cfg.nranks, rank = m->cfg.rank;
    if (nr < 2) return GP_OK;
    if (nr > RB_MAX_RANKS) return fail(w, GP_ERR_BAD_ARG, "gp_slab_rebalance: more than %d ranks", RB_MAX_RANKS);
    CK(cudaSetDevice(w->device));
    gp_status s = rb_alloc(w);
    if (s != GP_OK) return s;
    if (!(max_shift > 0)) max_shift = 0.5f * m->cfg.halo;
    uint32_t* d = m->rb_buf;
    // global range of the owned centres + everybody's left face (each rank fills its own slot, the sum is exact)
    rb_enqueue_minmax(w);
    uint32_t slots[RB_MAX_RANKS] = {};
    memcpy(&slots[rank], &m->cfg.x_lo, 4);
    CK(cudaMemcpyAsync(d + 2, slots, sizeof slots, cudaMemcpyHostToDevice, w->st));
    CKN(g_nccl.AllReduce(d, d, 1, ncclUint32, ncclMin, m->comm, w->st));
    CKN(g_nccl.AllReduce(d + 1, d + 1, 1, ncclUint32, ncclMax, m->comm, w->st));
    CKN(g_nccl.AllReduce(d + 2, d + 2, RB_MAX_RANKS, ncclUint32, ncclSum, m->comm, w->st));
    uint32_t h[2 + RB_MAX_RANKS];
    CK(cudaMemcpyAsync(h, d, sizeof h, cudaMemcpyDeviceToHost, w->st));
    CK(cudaStreamSynchronize(w->st));
    if (h[0] == 0xffffffffu || h[1] == 0u) return GP_OK;  // nobody owns anything
    const float gx0 = ord2f(h[0]), gx1 = ord2f(h[1]);
    if (!(gx1 > gx0)) return GP_OK;
    const float binw = (gx1 - gx0) / (float)RB_BINS * 1.0001f;
    rb_enqueue_hist(w, gx0, 1.0f / binw);
    CKN(g_nccl.AllReduce(d + 128, d + 128, RB_BINS, ncclUint32, ncclSum, m->comm, w->st));
    std::vector<uint32_t> hist(RB_BINS);
    CK(cudaMemcpyAsync(hist.data(), d + 128, RB_BINS * sizeof(uint32_t), cudaMemcpyDeviceToHost, w->st));
    CK(cudaStreamSynchronize(w->st));
    float faces[RB_MAX_RANKS + 1];
    for (int r = 0; r < nr; ++r) memcpy(&faces[r], &h[2 + r], 4);
    faces[0] = -INFINITY, faces[nr] = INFINITY;
    const float old_lo = faces[rank], old_hi = faces[rank + 1];
    if (!rb_new_faces(hist.data(), gx0, binw, nr, faces, max_shift, 2.0f * m->cfg.halo)) return GP_OK;
    if (moved_out) *moved_out = (faces[rank] != old_lo) || (faces[rank + 1] != old_hi);
    return rb_apply(w, faces[rank], faces[rank + 1]);
}

extern "C" gp_status gp_slab_rebalance_local(gp_world* const* worlds, int32_t n, float max_shift, uint32_t* moved_out) {
    if (!worlds || n < 1) return GP_ERR_BAD_ARG;
    if (moved_out) *moved_out = 0;
    if (n > RB_MAX_RANKS) return GP_ERR_BAD_ARG;
    for (int i = 0; i < n; ++i)
        if (!worlds[i] || !worlds[i]->multi || !worlds[i]->multi->local)
            return worlds[i] ? fail(worlds[i], GP_ERR_STATE, "gp_slab_rebalance_local needs worlds set up by gp_comm_init_local") : GP_ERR_BAD_ARG;
    if (n < 2) return GP_OK;
    uint32_t lo = 0xffffffffu, hi = 0u;
    for (int i = 0; i < n; ++i) {
        gp_world* w = worlds[i];
        CK(cudaSetDevice(w->device));
        gp_status s = rb_alloc(w);
        if (s != GP_OK) return s;
        rb_enqueue_minmax(w);
        uint32_t mm[2];
        CK(cudaMemcpyAsync(mm, w->multi->rb_buf, sizeof mm, cudaMemcpyDeviceToHost, w->st));
        CK(cudaStreamSynchronize(w->st));
        lo = std::min(lo, mm[0]), hi = std::max(hi, mm[1]);
    }
    if (lo == 0xffffffffu || hi == 0u) return GP_OK;
    const float gx0 = ord2f(lo), gx1 = ord2f(hi);
    if (!(gx1 > gx0)) return GP_OK;
    const float binw = (gx1 - gx0) / (float)RB_BINS * 1.0001f;
    std::vector<uint32_t> hist(RB_BINS, 0u), part(RB_BINS);
    for (int i = 0; i < n; ++i) {
        gp_world* w = worlds[i];
        CK(cudaSetDevice(w->device));
        rb_enqueue_hist(w, gx0, 1.0f / binw);
        CK(cudaMemcpyAsync(part.data(), w->multi->rb_buf + 128, RB_BINS * sizeof(uint32_t), cudaMemcpyDeviceToHost, w->st));
        CK(cudaStreamSynchronize(w->st));
        for (uint32_t b = 0; b < RB_BINS; ++b) hist[b] += part[b];
    }
    float faces[RB_MAX_RANKS + 1];
    for (int r = 0; r < n; ++r) faces[r] = worlds[r]->multi->cfg.x_lo;
    faces[0] = -INFINITY, faces[n] = INFINITY;
    float max_halo = 0.f;
    for (int i = 0; i < n; ++i) max_halo = std::max(max_halo, worlds[i]->multi->cfg.halo);
    if (!(max_shift > 0)) max_shift = 0.5f * max_halo;
    std::vector<float> old(faces, faces + n + 1);
    if (!rb_new_faces(hist.data(), gx0, binw, n, faces, max_shift, 2.0f * max_halo)) return GP_OK;
    for (int i = 0; i < n; ++i) {
        gp_world* w = worlds[i];
        if (faces[i] == old[i] && faces[i + 1] == old[i + 1]) continue;
        if (moved_out) *moved_out = 1;
        CK(cudaSetDevice(w->device));
        gp_status s = rb_apply(w, faces[i], faces[i + 1]);
        if (s != GP_OK) return s;
    }
    return GP_OK;
}
