// gears_oracle.cpp — CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE)
//
// A plain C++ restatement of the Gears physics step, used ONLY as the checker for
// the CUDA path (tests/, __graft_entry__.smoke(), bench.py cpu_baseline / --impl reference).
// Nothing under gears_b200/ may link, import or call this file.
//
// PARITY PIN STATUS: "parity unpinned by the reference".  The reference (Rust) cannot be
// compiled here (no cargo/rustc) and its own tests hold no numeric golden vectors
// (reference crates/gears-ecs/src/components/physics.rs:558-997 assert booleans and
// inequalities only).  This oracle is pinned against (a) every boolean/equality assertion
// of those tests, ported to tests/test_oracle_reference_tests.py, and (b) the derived
// float32 known-answer vectors in tests/golden/known_answers.json.
//
// Reference lines followed (all paths relative to /root/reference):
//   constants            crates/gears-ecs/src/components/physics.rs:7-20
//   world AABB           crates/gears-ecs/src/components/physics.rs:84-135
//   intersects           crates/gears-ecs/src/components/physics.rs:148-165
//   resolve              crates/gears-ecs/src/components/physics.rs:176-299
//   update_pos           crates/gears-ecs/src/components/physics.rs:405-462
//   cap_velocity         crates/gears-ecs/src/components/physics.rs:506-520
//   step (pass 1, 2)     crates/gears-app/src/systems/update_systems.rs:368-487
// Third-party arithmetic restated from its published source (not vendored in the reference):
//   cgmath 0.18.0 (Cargo.lock:393-399): Vector3 ops are component-wise; dot = (x*x'+y*y')+z*z';
//   magnitude = sqrt(dot); Quaternion*Vector3: t = cross(q.v,v) + v*q.s; r = cross(q.v,t)*2 + v.
//
// Build: g++ -O2 -ffp-contract=off -fopenmp -shared -fPIC   (NO fast-math, NO FMA contraction:
// Rust never contracts a*b+c, so neither may we.)

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#if defined(__FAST_MATH__)
#error "the oracle must be built without -ffast-math"
#endif

namespace {

struct V3 {
    float x, y, z;
};
static inline V3 v3(float x, float y, float z) { return V3{x, y, z}; }
static inline V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline V3 operator*(V3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
static inline V3 operator/(V3 a, float s) { return v3(a.x / s, a.y / s, a.z / s); }
// cgmath 0.18 InnerSpace::dot for Vector3: element-wise product, then left-fold sum.
static inline float dot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
static inline float magnitude(V3 a) { return sqrtf(dot(a, a)); }
// cgmath 0.18 Vector3::cross
static inline V3 cross(V3 a, V3 b) {
    return v3((a.y * b.z) - (a.z * b.y), (a.z * b.x) - (a.x * b.z), (a.x * b.y) - (a.y * b.x));
}

// physics.rs:7-20
constexpr float MAX_HORIZONTAL_VELOCITY = 20.0f;
constexpr float MAX_VERTICAL_VELOCITY = 40.0f;
constexpr float POSITION_CORRECTION_FACTOR = 0.4f;
constexpr float POSITION_CORRECTION_SLOP = 0.01f;
constexpr float FRICTION_COEFFICIENT = 0.8f;
constexpr float VELOCITY_THRESHOLD = 0.05f;
constexpr float GRAVITY_ACCELERATION = 28.0f;
constexpr float FALL_MULTIPLIER = 1.75f;
constexpr float APEX_MULTIPLIER = 0.6f;

struct Quat {  // cgmath Quaternion { v: Vector3, s }  — memory order (v.x, v.y, v.z, s)
    V3 v;
    float s;
};

struct Body {
    V3 pos;
    Quat rot;
    V3 scale;  // (1,1,1) when the entity has no Scale component (physics.rs:89-90)
    V3 vel, acc;
    V3 bmin, bmax;
    float mass, restitution;
    bool is_static;
};

struct Aabb {
    V3 lo, hi;
};

// physics.rs:84-135  get_rotated_aabb
static Aabb world_aabb(const Body& b) {
    const V3 c[8] = {
        v3(b.bmin.x, b.bmin.y, b.bmin.z), v3(b.bmax.x, b.bmin.y, b.bmin.z),
        v3(b.bmin.x, b.bmax.y, b.bmin.z), v3(b.bmax.x, b.bmax.y, b.bmin.z),
        v3(b.bmin.x, b.bmin.y, b.bmax.z), v3(b.bmax.x, b.bmin.y, b.bmax.z),
        v3(b.bmin.x, b.bmax.y, b.bmax.z), v3(b.bmax.x, b.bmax.y, b.bmax.z),
    };
    Aabb r{};
    for (int i = 0; i < 8; ++i) {
        V3 scaled = v3(c[i].x * b.scale.x, c[i].y * b.scale.y, c[i].z * b.scale.z);
        // cgmath: Quaternion * Vector3
        V3 tmp = cross(b.rot.v, scaled) + (scaled * b.rot.s);
        V3 rotated = (cross(b.rot.v, tmp) * 2.0f) + scaled;
        V3 w = rotated + b.pos;
        if (i == 0) {
            r.lo = w;
            r.hi = w;
        } else {
            r.lo.x = fminf(r.lo.x, w.x);
            r.lo.y = fminf(r.lo.y, w.y);
            r.lo.z = fminf(r.lo.z, w.z);
            r.hi.x = fmaxf(r.hi.x, w.x);
            r.hi.y = fmaxf(r.hi.y, w.y);
            r.hi.z = fmaxf(r.hi.z, w.z);
        }
    }
    return r;
}

// physics.rs:159-164 — strict inequalities: touching boxes do not intersect
static inline bool aabb_intersects(const Aabb& a, const Aabb& b) {
    return a.lo.x < b.hi.x && a.hi.x > b.lo.x && a.lo.y < b.hi.y && a.hi.y > b.lo.y &&
           a.lo.z < b.hi.z && a.hi.z > b.lo.z;
}

// physics.rs:405-462
static void update_pos(Body& b, float dt) {
    if (b.is_static) return;
    const float acceleration_threshold = 0.01f;
    const bool falling = b.vel.y < 0.0f;
    const bool near_apex = fabsf(b.vel.y) < 2.0f && b.vel.y > 0.0f;
    const float gravity_multiplier = falling ? FALL_MULTIPLIER : (near_apex ? APEX_MULTIPLIER : 1.0f);
    b.vel.y -= GRAVITY_ACCELERATION * gravity_multiplier * dt;
    const bool is_accelerating = magnitude(b.acc) > acceleration_threshold;
    const float damping_coefficient = is_accelerating ? 1.6f : (falling ? 1.8f : 3.5f);
    const float damping_factor = expf(-damping_coefficient * dt);
    const float min_velocity = 0.005f;
    if (falling) {
        b.vel.x *= damping_factor;
        b.vel.z *= damping_factor;
    } else {
        b.vel = b.vel * damping_factor;
    }
    b.vel.x += b.acc.x * dt;
    b.vel.z += b.acc.z * dt;
    if (magnitude(b.vel) < min_velocity) b.vel = v3(0.0f, 0.0f, 0.0f);
    b.pos = b.pos + b.vel * dt;
}

// physics.rs:176-299.  Precondition in the reference: intersects() returned true.
static void resolve(Body& a, Body& b) {
    const Aabb A = world_aabb(a), B = world_aabb(b);
    const float overlap_x = fminf(A.hi.x, B.hi.x) - fmaxf(A.lo.x, B.lo.x);
    const float overlap_y = fminf(A.hi.y, B.hi.y) - fmaxf(A.lo.y, B.lo.y);
    const float overlap_z = fminf(A.hi.z, B.hi.z) - fmaxf(A.lo.z, B.lo.z);
    const V3 a_center = (A.lo + A.hi) * 0.5f;
    const V3 b_center = (B.lo + B.hi) * 0.5f;
    const V3 center_diff = b_center - a_center;

    float min_overlap;
    V3 normal;
    if (overlap_x <= overlap_y && overlap_x <= overlap_z) {
        min_overlap = overlap_x;
        normal = v3(center_diff.x > 0.0f ? 1.0f : -1.0f, 0.0f, 0.0f);
    } else if (overlap_y <= overlap_z) {
        min_overlap = overlap_y;
        normal = v3(0.0f, center_diff.y > 0.0f ? 1.0f : -1.0f, 0.0f);
    } else {
        min_overlap = overlap_z;
        normal = v3(0.0f, 0.0f, center_diff.z > 0.0f ? 1.0f : -1.0f);
    }

    const float inv_mass_a = a.is_static ? 0.0f : 1.0f / a.mass;
    const float inv_mass_b = b.is_static ? 0.0f : 1.0f / b.mass;
    const float total_inv_mass = inv_mass_a + inv_mass_b;
    if (total_inv_mass <= 0.0f) return;

    const V3 relative_velocity = b.vel - a.vel;
    const float vel_along_normal = dot(relative_velocity, normal);
    if (vel_along_normal > 0.0f) return;

    if (min_overlap > POSITION_CORRECTION_SLOP) {
        const float correction_magnitude =
            (min_overlap - POSITION_CORRECTION_SLOP) * POSITION_CORRECTION_FACTOR / total_inv_mass;
        const V3 correction = normal * correction_magnitude;
        if (!a.is_static) a.pos = a.pos - correction * inv_mass_a;
        if (!b.is_static) b.pos = b.pos + correction * inv_mass_b;
    }

    const float restitution =
        fabsf(vel_along_normal) < VELOCITY_THRESHOLD ? 0.0f : (a.restitution + b.restitution) * 0.5f;
    const float impulse_scalar = -(1.0f + restitution) * vel_along_normal / total_inv_mass;
    const V3 impulse = normal * impulse_scalar;
    if (!a.is_static) a.vel = a.vel - impulse * inv_mass_a;
    if (!b.is_static) b.vel = b.vel + impulse * inv_mass_b;

    const V3 tangent_velocity = relative_velocity - normal * vel_along_normal;
    const float tangent_speed = magnitude(tangent_velocity);
    if (tangent_speed > 0.001f) {
        const V3 tangent = tangent_velocity / tangent_speed;
        const float friction_impulse = FRICTION_COEFFICIENT * fabsf(impulse_scalar);
        const float max_friction = tangent_speed;
        const float effective = fminf(friction_impulse, max_friction * total_inv_mass);
        if (!a.is_static) a.vel = a.vel - tangent * effective * inv_mass_a;
        if (!b.is_static) b.vel = b.vel + tangent * effective * inv_mass_b;
    }

    if (!a.is_static && magnitude(a.vel) < VELOCITY_THRESHOLD * 0.5f) a.vel = v3(0, 0, 0);
    if (!b.is_static && magnitude(b.vel) < VELOCITY_THRESHOLD * 0.5f) b.vel = v3(0, 0, 0);
}

// physics.rs:506-520
static void cap_velocity(V3& v) {
    V3 h = v3(v.x, 0.0f, v.z);
    float m = magnitude(h);
    if (m > MAX_HORIZONTAL_VELOCITY) {
        V3 n = h * (1.0f / m);  // cgmath normalize(): self * (1 / magnitude)
        v.x = n.x * MAX_HORIZONTAL_VELOCITY;
        v.z = n.z * MAX_HORIZONTAL_VELOCITY;
    }
    // f32::clamp
    if (v.y < -MAX_VERTICAL_VELOCITY) v.y = -MAX_VERTICAL_VELOCITY;
    if (v.y > MAX_VERTICAL_VELOCITY) v.y = MAX_VERTICAL_VELOCITY;
}

// ---------------------------------------------------------------------------------------
// Scene container in ITERATION ORDER: slot k holds the body the reference's
// `physics_entities[k]` names (update_systems.rs:370-371, order from ecs lib.rs:422-433).
// ---------------------------------------------------------------------------------------
struct Scene {
    uint32_t n = 0;
    std::vector<Body> b;        // in iteration order
    std::vector<uint32_t> src;  // src[k] = caller's index of the body at slot k
};

struct SoA {
    uint32_t n;
    float* pos3;
    const float* rot4;    // nullable → identity
    const float* scale3;  // nullable → (1,1,1)
    float* vel3;
    const float* acc3;
    const float* box_min3;
    const float* box_max3;
    const float* mass;
    const float* restitution;
    const uint8_t* is_static;
    const uint32_t* rank;  // nullable → identity; rank[i] = slot of caller body i
};

static int load_scene(const SoA& s, Scene& sc) {
    sc.n = s.n;
    sc.b.resize(s.n);
    sc.src.assign(s.n, 0xFFFFFFFFu);
    for (uint32_t i = 0; i < s.n; ++i) {
        uint32_t k = s.rank ? s.rank[i] : i;
        if (k >= s.n || sc.src[k] != 0xFFFFFFFFu) return -1;  // rank must be a permutation
        sc.src[k] = i;
        Body& b = sc.b[k];
        b.pos = v3(s.pos3[3 * i], s.pos3[3 * i + 1], s.pos3[3 * i + 2]);
        if (s.rot4)
            b.rot = Quat{v3(s.rot4[4 * i], s.rot4[4 * i + 1], s.rot4[4 * i + 2]), s.rot4[4 * i + 3]};
        else
            b.rot = Quat{v3(0, 0, 0), 1.0f};
        b.scale = s.scale3 ? v3(s.scale3[3 * i], s.scale3[3 * i + 1], s.scale3[3 * i + 2]) : v3(1, 1, 1);
        b.vel = v3(s.vel3[3 * i], s.vel3[3 * i + 1], s.vel3[3 * i + 2]);
        b.acc = s.acc3 ? v3(s.acc3[3 * i], s.acc3[3 * i + 1], s.acc3[3 * i + 2]) : v3(0, 0, 0);
        b.bmin = v3(s.box_min3[3 * i], s.box_min3[3 * i + 1], s.box_min3[3 * i + 2]);
        b.bmax = v3(s.box_max3[3 * i], s.box_max3[3 * i + 1], s.box_max3[3 * i + 2]);
        b.mass = s.mass[i];
        b.restitution = s.restitution[i];
        b.is_static = s.is_static[i] != 0;
    }
    return 0;
}

static void store_scene(const Scene& sc, const SoA& s) {
    for (uint32_t k = 0; k < sc.n; ++k) {
        uint32_t i = sc.src[k];
        const Body& b = sc.b[k];
        s.pos3[3 * i] = b.pos.x, s.pos3[3 * i + 1] = b.pos.y, s.pos3[3 * i + 2] = b.pos.z;
        s.vel3[3 * i] = b.vel.x, s.vel3[3 * i + 1] = b.vel.y, s.vel3[3 * i + 2] = b.vel.z;
    }
}

struct PairSink {
    uint32_t* out;  // 2 u32 per pair: caller indices (first = earlier in iteration order)
    uint64_t cap, n;
    void push(uint32_t a, uint32_t b) {
        if (out && n < cap) {
            out[2 * n] = a;
            out[2 * n + 1] = b;
        }
        ++n;
    }
};

// Pass 1, update_systems.rs:378-405 (rayon par_iter → OpenMP; bodies are independent).
static void pass1(Scene& sc, float dt, int threads) {
    const int64_t n = sc.n;
#pragma omp parallel for schedule(static) num_threads(threads) if (threads > 1)
    for (int64_t k = 0; k < n; ++k) update_pos(sc.b[k], dt);
}

// Cached world AABBs in slot order as SoA, so the O(N^2) inner loop streams.  The cache is
// refreshed with the literal 8-corner transform whenever a body's position changes, so every
// test sees exactly what physics.rs:156-157 would recompute.
struct AabbCache {
    std::vector<float> lx, ly, lz, hx, hy, hz;
    void resize(size_t n) {
        lx.resize(n), ly.resize(n), lz.resize(n), hx.resize(n), hy.resize(n), hz.resize(n);
    }
    void set(size_t k, const Aabb& a) {
        lx[k] = a.lo.x, ly[k] = a.lo.y, lz[k] = a.lo.z, hx[k] = a.hi.x, hy[k] = a.hi.y, hz[k] = a.hi.z;
    }
    Aabb get(size_t k) const { return Aabb{v3(lx[k], ly[k], lz[k]), v3(hx[k], hy[k], hz[k])}; }
};

// Index of the first slot in [j, n) whose cached AABB intersects (lo, hi), or n.  The
// predicate is physics.rs:159-164 evaluated without short-circuit in blocks so the compiler
// can vectorise the scan; the result is identical to the scalar left-to-right search.
static inline uint32_t first_hit(const AabbCache& c, uint32_t j, uint32_t n, const Aabb& a) {
    const float *lx = c.lx.data(), *ly = c.ly.data(), *lz = c.lz.data();
    const float *hx = c.hx.data(), *hy = c.hy.data(), *hz = c.hz.data();
    const float alx = a.lo.x, aly = a.lo.y, alz = a.lo.z, ahx = a.hi.x, ahy = a.hi.y, ahz = a.hi.z;
    while (j < n) {
        const uint32_t end = std::min<uint32_t>(j + 256u, n);
        int any = 0;
        for (uint32_t k = j; k < end; ++k)
            any |= int(alx < hx[k]) & int(ahx > lx[k]) & int(aly < hy[k]) & int(ahy > ly[k]) & int(alz < hz[k]) &
                   int(ahz > lz[k]);
        if (any) {
            for (uint32_t k = j; k < end; ++k)
                if (alx < hx[k] && ahx > lx[k] && aly < hy[k] && ahy > ly[k] && alz < hz[k] && ahz > lz[k]) return k;
        }
        j = end;
    }
    return n;
}

// S_snapshot: pairs intersecting on the post-integration state, both-static
// pairs listed separately.  Brute force; for tests only.
static void snapshot_pairs_brute(const Scene& sc, const AabbCache& c, PairSink& snap, PairSink& stat) {
    const uint32_t n = sc.n;
    for (uint32_t i = 0; i < n; ++i) {
        const Aabb a = c.get(i);
        for (uint32_t j = first_hit(c, i + 1, n, a); j < n; j = first_hit(c, j + 1, n, a)) {
            if (sc.b[i].is_static && sc.b[j].is_static)
                stat.push(sc.src[i], sc.src[j]);
            else
                snap.push(sc.src[i], sc.src[j]);
        }
    }
}

// Pass 2, update_systems.rs:408-487: the sequential i<j sweep, test at the pair's turn
// with CURRENT positions, resolve in place.  S_sweep = pairs whose test was true
// (both-static pairs are no-ops via physics.rs:225 and are left out of S_sweep).
static void pass2_brute(Scene& sc, AabbCache& c, PairSink& sweep) {
    const uint32_t n = sc.n;
    for (uint32_t i = 0; i < n; ++i) {
        Aabb a = c.get(i);
        for (uint32_t j = first_hit(c, i + 1, n, a); j < n; j = first_hit(c, j + 1, n, a)) {
            Body &A = sc.b[i], &B = sc.b[j];
            if (!(A.is_static && B.is_static)) sweep.push(sc.src[i], sc.src[j]);
            resolve(A, B);
            a = world_aabb(A);
            c.set(i, a);
            c.set(j, world_aabb(B));
        }
    }
}

// ---------------------------------------------------------------------------------------
// Grid-accelerated sweep ("B-grid"): NOT the reference's algorithm, but constructed to give
// bit-identical results to pass2_brute (asserted in tests at N where brute force finishes).
//
// Invariant used: while the outer index i is fixed, no body other than i changes position
// except body j at its own pair (i,j) — and j is never visited again for this i.  So the
// set of j>i that can intersect i during its loop lies within i's AABB swept over the loop.
// We query a spatial hash with i's AABB inflated by R, walk the candidates in slot order,
// and re-query (with doubled R, only for slots beyond the current j) if i drifts past R.
// Small bodies sit in the hash keyed by the cell of their centre at insertion time and are
// re-inserted when they drift more than `tol`; bodies larger than the cell go to a short
// list that is always tested.
// ---------------------------------------------------------------------------------------
// Dense uniform grid with intrusive per-cell lists (head / next / prev): O(1) insert and remove, one array access
// per cell lookup.  Cell indices are clamped to the grid, so a body that drifts outside the bounds seen at
// construction lands in an edge cell and a query that pokes outside is clamped the same (monotone) way: the cells a
// query visits always contain every body whose centre lies in the query box.
struct DenseGrid {
    float cs = 1.0f, inv = 1.0f;
    float lo[3] = {0, 0, 0};
    int64_t dim[3] = {1, 1, 1};
    std::vector<int32_t> head, next, prev;
    std::vector<uint32_t> cell_of;
    void init(const float blo[3], const float bhi[3], float cell, size_t n) {
        cs = cell;
        for (;;) {
            double cells = 1;
            for (int k = 0; k < 3; ++k) {
                lo[k] = blo[k];
                double d = std::floor(((double)bhi[k] - (double)blo[k]) / (double)cs) + 1.0;
                if (!(d >= 1.0)) d = 1.0;
                dim[k] = (int64_t)std::min(d, 1.0e6);
                cells *= (double)dim[k];
            }
            if (cells <= 2.0e8) break;
            cs *= 1.26f;  // the table must fit: coarser cells only add candidates
        }
        inv = 1.0f / cs;
        head.assign((size_t)(dim[0] * dim[1] * dim[2]), -1);
        next.assign(n, -1), prev.assign(n, -1), cell_of.assign(n, 0);
    }
    int64_t cell(float v, int k) const {
        const float f = floorf((v - lo[k]) * inv);
        if (!(f > 0.0f)) return 0;  // (NaN lands in cell 0; a NaN box never intersects anything)
        return std::min<int64_t>((int64_t)std::min(f, 2.0e6f), dim[k] - 1);
    }
    size_t index(int64_t x, int64_t y, int64_t z) const { return (size_t)(x + dim[0] * (y + dim[1] * z)); }
    void insert(uint32_t k, V3 c) {
        const size_t ci = index(cell(c.x, 0), cell(c.y, 1), cell(c.z, 2));
        cell_of[k] = (uint32_t)ci;
        prev[k] = -1, next[k] = head[ci];
        if (head[ci] >= 0) prev[head[ci]] = (int32_t)k;
        head[ci] = (int32_t)k;
    }
    void remove(uint32_t k) {
        const size_t ci = cell_of[k];
        if (prev[k] >= 0) next[prev[k]] = next[k]; else head[ci] = next[k];
        if (next[k] >= 0) prev[next[k]] = prev[k];
    }
};

static inline V3 center_of(const Aabb& a) { return (a.lo + a.hi) * 0.5f; }
static inline float max_extent(const Aabb& a) {
    return fmaxf(a.hi.x - a.lo.x, fmaxf(a.hi.y - a.lo.y, a.hi.z - a.lo.z));
}

struct SweepGrid {
    Scene& sc;
    AabbCache& c;
    DenseGrid g;
    uint32_t n_in_grid = 0;
    float med = 1.0f, thr = 1.0f, tol = 0.0f;
    std::vector<uint32_t> big;
    std::vector<V3> ins_c;      // centre at insertion
    std::vector<Aabb> ins_box;  // AABB at insertion
    std::vector<uint8_t> in_grid;

    SweepGrid(Scene& s, AabbCache& cc) : sc(s), c(cc) {
        const uint32_t n = sc.n;
        std::vector<float> ext(n);
        for (uint32_t k = 0; k < n; ++k) ext[k] = max_extent(c.get(k));
        std::vector<float> tmp(ext);
        std::sort(tmp.begin(), tmp.end());
        med = std::max(tmp[n / 2], 1e-6f);
        // cell size = the largest extent that leaves at most 64 bodies above it (they go to the big list, which is
        // always tested), but never below the median extent: the tightest grid that still hosts all but 64 bodies
        // (a pile of unit boxes gets unit cells instead of 4x cells holding ~80 bodies each)
        thr = n > 64 ? std::max(med, tmp[n - 65]) : 4.0f * med;
        tol = 0.125f * med;
        ins_c.resize(n);
        ins_box.resize(n);
        in_grid.assign(n, 0);
        float blo[3] = {INFINITY, INFINITY, INFINITY}, bhi[3] = {-INFINITY, -INFINITY, -INFINITY};
        for (uint32_t k = 0; k < n; ++k) {
            const Aabb a = c.get(k);
            const V3 ce = center_of(a);
            if (!(ext[k] <= thr) || !std::isfinite(ext[k]) || !std::isfinite(ce.x) || !std::isfinite(ce.y) ||
                !std::isfinite(ce.z)) {
                big.push_back(k);
            } else {
                in_grid[k] = 1;
                ++n_in_grid;
                ins_c[k] = ce;
                ins_box[k] = a;
                blo[0] = std::min(blo[0], ce.x), blo[1] = std::min(blo[1], ce.y), blo[2] = std::min(blo[2], ce.z);
                bhi[0] = std::max(bhi[0], ce.x), bhi[1] = std::max(bhi[1], ce.y), bhi[2] = std::max(bhi[2], ce.z);
            }
        }
        if (!n_in_grid) blo[0] = blo[1] = blo[2] = 0.0f, bhi[0] = bhi[1] = bhi[2] = 1.0f;
        g.init(blo, bhi, thr + 2.0f * tol, n);
        for (uint32_t k = 0; k < n; ++k)
            if (in_grid[k]) g.insert(k, ins_c[k]);
    }
    // body k moved: refresh its cached AABB, re-insert if it drifted more than tol
    void refresh(uint32_t k) {
        Aabb a = world_aabb(sc.b[k]);
        c.set(k, a);
        if (!in_grid[k]) return;
        const Aabb& o = ins_box[k];
        if (a.lo.x < o.lo.x - tol || a.lo.y < o.lo.y - tol || a.lo.z < o.lo.z - tol || a.hi.x > o.hi.x + tol ||
            a.hi.y > o.hi.y + tol || a.hi.z > o.hi.z + tol) {
            g.remove(k);
            ins_c[k] = center_of(a);
            ins_box[k] = a;
            g.insert(k, ins_c[k]);
        }
    }
    // all slots k > after whose current AABB can overlap q inflated by R (superset), ascending
    void query(const Aabb& q, float R, uint32_t after, std::vector<uint32_t>& cand) {
        cand.clear();
        // the stored centre of any small partner lies within [q.lo - reach, q.hi + reach]
        const float reach = R + tol + 0.5f * thr;
        const int64_t x0 = g.cell(q.lo.x - reach, 0), x1 = g.cell(q.hi.x + reach, 0);
        const int64_t y0 = g.cell(q.lo.y - reach, 1), y1 = g.cell(q.hi.y + reach, 1);
        const int64_t z0 = g.cell(q.lo.z - reach, 2), z1 = g.cell(q.hi.z + reach, 2);
        const double ncell = double(x1 - x0 + 1) * double(y1 - y0 + 1) * double(z1 - z0 + 1);
        if (ncell > 4.0 * double(n_in_grid) + 64.0) {
            // huge query box (e.g. a floor): every body in the grid is a candidate
            for (uint32_t k = after + 1; k < sc.n; ++k)
                if (in_grid[k]) cand.push_back(k);
        } else {
            for (int64_t z = z0; z <= z1; ++z)
                for (int64_t y = y0; y <= y1; ++y)
                    for (int64_t x = x0; x <= x1; ++x)
                        for (int32_t k = g.head[g.index(x, y, z)]; k >= 0; k = g.next[k])
                            if ((uint32_t)k > after) cand.push_back((uint32_t)k);
        }
        for (uint32_t k : big)
            if (k > after) cand.push_back(k);
        std::sort(cand.begin(), cand.end());
        cand.erase(std::unique(cand.begin(), cand.end()), cand.end());
    }
};

static void pass2_grid(Scene& sc, AabbCache& c, PairSink& sweep) {
    const uint32_t n = sc.n;
    if (n == 0) return;
    SweepGrid sg(sc, c);
    std::vector<uint32_t> cand;
    for (uint32_t i = 0; i < n; ++i) {
        float R = 0.125f * sg.med;
        uint32_t after = i;  // only slots > after are still to be visited for this i
        Aabb q = c.get(i);   // i's AABB when the candidate list was built
        sg.query(q, R, after, cand);
        size_t pos = 0;
        while (pos < cand.size()) {
            const uint32_t j = cand[pos++];
            after = j;
            const Aabb a = c.get(i), b = c.get(j);
            if (!aabb_intersects(a, b)) continue;
            Body &A = sc.b[i], &B = sc.b[j];
            if (!(A.is_static && B.is_static)) sweep.push(sc.src[i], sc.src[j]);
            resolve(A, B);
            sg.refresh(i);
            sg.refresh(j);
            const Aabb na = c.get(i);
            if (na.lo.x < q.lo.x - R || na.lo.y < q.lo.y - R || na.lo.z < q.lo.z - R || na.hi.x > q.hi.x + R ||
                na.hi.y > q.hi.y + R || na.hi.z > q.hi.z + R) {
                R *= 2.0f;
                q = na;
                sg.query(q, R, after, cand);
                pos = 0;
            }
        }
    }
}

// Snapshot pairs through the same grid (exact: nothing moves while it runs).
static void snapshot_pairs_grid(Scene& sc, AabbCache& c, PairSink& snap, PairSink& stat) {
    const uint32_t n = sc.n;
    if (n == 0) return;
    SweepGrid sg(sc, c);
    std::vector<uint32_t> cand;
    for (uint32_t i = 0; i < n; ++i) {
        const Aabb a = c.get(i);
        sg.query(a, 0.0f, i, cand);
        for (uint32_t j : cand) {
            if (!aabb_intersects(a, c.get(j))) continue;
            if (sc.b[i].is_static && sc.b[j].is_static)
                stat.push(sc.src[i], sc.src[j]);
            else
                snap.push(sc.src[i], sc.src[j]);
        }
    }
}

}  // namespace

// =======================================================================================
// C ABI for the test harness (ctypes).  All pair outputs are (caller index of the earlier
// body, caller index of the later body) in the reference's iteration order.
// =======================================================================================
extern "C" {

struct go_step_out {
    uint64_t n_snapshot, n_sweep, n_static_static;
};

// mode: 0 = brute force (the reference's algorithm), 1 = grid-accelerated (bit-identical)
// want_snapshot: compute S_snapshot / static-static lists (costly in brute mode).
int go_step(uint32_t n, float* pos3, const float* rot4, const float* scale3, float* vel3, const float* acc3,
            const float* box_min3, const float* box_max3, const float* mass, const float* restitution,
            const uint8_t* is_static, const uint32_t* rank, float dt, int steps, int mode, int threads,
            int want_snapshot, uint32_t* snapshot_pairs, uint64_t snapshot_cap, uint32_t* sweep_pairs,
            uint64_t sweep_cap, uint32_t* static_pairs, uint64_t static_cap, go_step_out* out) {
    SoA s{n, pos3, rot4, scale3, vel3, acc3, box_min3, box_max3, mass, restitution, is_static, rank};
    Scene sc;
    if (load_scene(s, sc) != 0) return -1;
    AabbCache c;
    c.resize(n);
    PairSink snap{snapshot_pairs, snapshot_cap, 0}, sweep{sweep_pairs, sweep_cap, 0},
        stat{static_pairs, static_cap, 0};
    for (int st = 0; st < steps; ++st) {
        const bool last = (st == steps - 1);
        pass1(sc, dt, threads);
        for (uint32_t k = 0; k < n; ++k) c.set(k, world_aabb(sc.b[k]));
        PairSink nul{nullptr, 0, 0};
        if (want_snapshot && last) {
            if (mode == 0)
                snapshot_pairs_brute(sc, c, snap, stat);
            else
                snapshot_pairs_grid(sc, c, snap, stat);
        }
        if (mode == 0)
            pass2_brute(sc, c, last ? sweep : nul);
        else
            pass2_grid(sc, c, last ? sweep : nul);
    }
    store_scene(sc, s);
    if (out) {
        out->n_snapshot = snap.n;
        out->n_sweep = sweep.n;
        out->n_static_static = stat.n;
    }
    return 0;
}

// ---- single-function entry points for known-answer tests ----
void go_update_pos(float* pos3, float* vel3, const float* acc3, int is_static, float dt) {
    Body b{};
    b.pos = v3(pos3[0], pos3[1], pos3[2]);
    b.vel = v3(vel3[0], vel3[1], vel3[2]);
    b.acc = v3(acc3[0], acc3[1], acc3[2]);
    b.is_static = is_static != 0;
    update_pos(b, dt);
    pos3[0] = b.pos.x, pos3[1] = b.pos.y, pos3[2] = b.pos.z;
    vel3[0] = b.vel.x, vel3[1] = b.vel.y, vel3[2] = b.vel.z;
}

static Body make_body(const float* pos3, const float* rot4, const float* scale3, const float* vel3,
                      const float* bmin, const float* bmax, float mass, float rest, int is_static) {
    Body b{};
    b.pos = v3(pos3[0], pos3[1], pos3[2]);
    b.rot = rot4 ? Quat{v3(rot4[0], rot4[1], rot4[2]), rot4[3]} : Quat{v3(0, 0, 0), 1.0f};
    b.scale = scale3 ? v3(scale3[0], scale3[1], scale3[2]) : v3(1, 1, 1);
    b.vel = vel3 ? v3(vel3[0], vel3[1], vel3[2]) : v3(0, 0, 0);
    b.acc = v3(0, 0, 0);
    b.bmin = v3(bmin[0], bmin[1], bmin[2]);
    b.bmax = v3(bmax[0], bmax[1], bmax[2]);
    b.mass = mass;
    b.restitution = rest;
    b.is_static = is_static != 0;
    return b;
}

void go_world_aabb(const float* bmin, const float* bmax, const float* pos3, const float* rot4,
                   const float* scale3, float* out_min3, float* out_max3) {
    Body b = make_body(pos3, rot4, scale3, nullptr, bmin, bmax, 1.0f, 0.2f, 0);
    Aabb a = world_aabb(b);
    out_min3[0] = a.lo.x, out_min3[1] = a.lo.y, out_min3[2] = a.lo.z;
    out_max3[0] = a.hi.x, out_max3[1] = a.hi.y, out_max3[2] = a.hi.z;
}

int go_intersects(const float* a_min, const float* a_max, const float* a_pos, const float* a_rot,
                  const float* a_scale, const float* b_min, const float* b_max, const float* b_pos,
                  const float* b_rot, const float* b_scale) {
    Body a = make_body(a_pos, a_rot, a_scale, nullptr, a_min, a_max, 1, 0.2f, 0);
    Body b = make_body(b_pos, b_rot, b_scale, nullptr, b_min, b_max, 1, 0.2f, 0);
    return aabb_intersects(world_aabb(a), world_aabb(b)) ? 1 : 0;
}

// check_and_resolve_collision, physics.rs:474-499.  Returns 1 if the boxes intersected.
int go_check_and_resolve(float* a_pos, float* a_vel, const float* a_rot, const float* a_scale,
                         const float* a_min, const float* a_max, float a_mass, float a_rest, int a_static,
                         float* b_pos, float* b_vel, const float* b_rot, const float* b_scale,
                         const float* b_min, const float* b_max, float b_mass, float b_rest, int b_static) {
    Body a = make_body(a_pos, a_rot, a_scale, a_vel, a_min, a_max, a_mass, a_rest, a_static);
    Body b = make_body(b_pos, b_rot, b_scale, b_vel, b_min, b_max, b_mass, b_rest, b_static);
    const bool hit = aabb_intersects(world_aabb(a), world_aabb(b));
    if (hit) resolve(a, b);
    a_pos[0] = a.pos.x, a_pos[1] = a.pos.y, a_pos[2] = a.pos.z;
    a_vel[0] = a.vel.x, a_vel[1] = a.vel.y, a_vel[2] = a.vel.z;
    b_pos[0] = b.pos.x, b_pos[1] = b.pos.y, b_pos[2] = b.pos.z;
    b_vel[0] = b.vel.x, b_vel[1] = b.vel.y, b_vel[2] = b.vel.z;
    return hit ? 1 : 0;
}

void go_cap_velocity(float* vel3) {
    V3 v = v3(vel3[0], vel3[1], vel3[2]);
    cap_velocity(v);
    vel3[0] = v.x, vel3[1] = v.y, vel3[2] = v.z;
}

// ---- "next" rows ----------------------------------------------------------
// N1  Instance::to_raw, crates/gears-renderer/src/instance.rs:22-31:
//     model = Matrix4::from_translation(p) * Matrix4::from(q) * Matrix4::from_nonuniform_scale(s); normal = Matrix3::from(q).
//     cgmath 0.18.0 semantics restated (not vendored): Matrix3/4::from(Quaternion) builds x2 = x+x ..., xx2 = x2*x ...,
//     columns (1-yy2-zz2, xy2+sz2, xz2-sy2), (xy2-sz2, 1-xx2-zz2, yz2+sx2), (xz2+sy2, yz2-sx2, 1-xx2-yy2);
//     Matrix4*Matrix4: column j = ((a*r[j][0] + b*r[j][1]) + c*r[j][2]) + d*r[j][3] with a..d the columns of the left
//     operand.  Output: 16 floats column-major + 9 floats column-major (what `.into()` yields).
static void mat4_mul(const float* A, const float* R, float* O) {  // column-major 4x4
    for (int j = 0; j < 4; ++j)
        for (int i = 0; i < 4; ++i) {
            float acc = A[0 * 4 + i] * R[j * 4 + 0];
            acc = acc + A[1 * 4 + i] * R[j * 4 + 1];
            acc = acc + A[2 * 4 + i] * R[j * 4 + 2];
            acc = acc + A[3 * 4 + i] * R[j * 4 + 3];
            O[j * 4 + i] = acc;
        }
}
void go_instance_raw(const float* pos3, const float* rot4 /*v.xyz, s*/, const float* scale3, float* out25) {
    const float x = rot4[0], y = rot4[1], z = rot4[2], w = rot4[3];
    const float x2 = x + x, y2 = y + y, z2 = z + z;
    const float xx2 = x2 * x, xy2 = x2 * y, xz2 = x2 * z, yy2 = y2 * y, yz2 = y2 * z, zz2 = z2 * z;
    const float sy2 = y2 * w, sz2 = z2 * w, sx2 = x2 * w;
    const float r3[9] = {1.0f - yy2 - zz2, xy2 + sz2, xz2 - sy2, xy2 - sz2, 1.0f - xx2 - zz2,
                         yz2 + sx2,        xz2 + sy2, yz2 - sx2, 1.0f - xx2 - yy2};
    const float T[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, pos3[0], pos3[1], pos3[2], 1};
    const float R[16] = {r3[0], r3[1], r3[2], 0, r3[3], r3[4], r3[5], 0, r3[6], r3[7], r3[8], 0, 0, 0, 0, 1};
    const float S[16] = {scale3[0], 0, 0, 0, 0, scale3[1], 0, 0, 0, 0, scale3[2], 0, 0, 0, 0, 1};
    float TR[16];
    mat4_mul(T, R, TR);
    mat4_mul(TR, S, out25);
    for (int k = 0; k < 9; ++k) out25[16 + k] = r3[k];
}

// N3  Weapon::ray_intersects_aabb, crates/gears-ecs/src/components/interactive.rs:94-158 (slab method; uses
//     pos + collision_box.min/max: rotation and scale are ignored by the reference)
int go_ray_intersects_aabb(const float* origin3, const float* dir3, const float* pos3, const float* bmin3,
                           const float* bmax3) {
    float tmin = 0.0f, tmax = 3.40282347e+38f;  // f32::MAX
    for (int i = 0; i < 3; ++i) {
        const float o = origin3[i], d = dir3[i];
        const float mn = pos3[i] + bmin3[i], mx = pos3[i] + bmax3[i];
        if (fabsf(d) < 1e-8f) {
            if (o < mn || o > mx) return 0;
        } else {
            const float inv = 1.0f / d;
            float t1 = (mn - o) * inv, t2 = (mx - o) * inv;
            if (t1 > t2) std::swap(t1, t2);
            tmin = fmaxf(tmin, t1);  // f32::max / f32::min = IEEE maxNum / minNum
            tmax = fminf(tmax, t2);
            if (tmin > tmax) return 0;
        }
    }
    return (tmax >= 0.0f && tmin <= tmax) ? 1 : 0;
}

// N4  PathfindingGrid::add_obstacle, crates/gears-ecs/src/components/pathfinding.rs:296-338 with
//     GridPos::from_world_pos (:57-63): cell = round(coord / cell_size) (f32::round = half away from zero, `as i32`
//     saturates, NaN -> 0).  Writes the inclusive cell range [lo, hi] of the obstacle.
static int32_t f2i_sat(float f) {
    if (f != f) return 0;
    if (f >= 2147483648.0f) return 2147483647;
    if (f <= -2147483648.0f) return (-2147483647 - 1);
    return (int32_t)f;
}
void go_obstacle_cells(const float* pos3, const float* bmin3, const float* bmax3, float cell_size, int32_t* lo3,
                       int32_t* hi3) {
    for (int i = 0; i < 3; ++i) {
        lo3[i] = f2i_sat(roundf((pos3[i] + bmin3[i]) / cell_size));
        hi3[i] = f2i_sat(roundf((pos3[i] + bmax3[i]) / cell_size));
    }
}

// exp(-c*dt) for c in {1.6, 1.8, 3.5} with the platform libm expf — what Rust's f32::exp calls.
void go_damping_factors(float dt, float* out3) {
    out3[0] = expf(-1.6f * dt);
    out3[1] = expf(-1.8f * dt);
    out3[2] = expf(-3.5f * dt);
}

int go_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

}  // extern "C"
