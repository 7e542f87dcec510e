// gp_mirror_kernels.cuh — component mirror: upload / refresh / download kernels and RigidBody::update_pos.
// Part of the physics step; see gp_kernels.cuh for the map of kernels to reference lines.
#pragma once
#include "gp_common.cuh"

namespace gp {

// ---------------------------------------------------------------------------------------------
// Upload: derive per-body AABB offsets.  off_min/off_max = component-wise min/max over the 8
// corners of rot*(corner*scale) (physics.rs:93-119 without the translation).  Because rounding is
// monotone, min_i fl(r_i + p) == fl(min_i r_i + p): adding pos afterwards gives the reference's AABB.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void corner_offsets(V3 bmin, V3 bmax, float4 q /*v.xyz, s*/, V3 sc, V3& omin, V3& omax) {
    const V3 qv = mk(q.x, q.y, q.z);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        V3 c = mk((i & 1) ? bmax.x : bmin.x, (i & 2) ? bmax.y : bmin.y, (i & 4) ? bmax.z : bmin.z);
        V3 s = mk(c.x * sc.x, c.y * sc.y, c.z * sc.z);
        V3 t = cross3(qv, s) + (s * q.w);
        V3 r = (cross3(qv, t) * 2.0f) + s;
        if (i == 0) {
            omin = r;
            omax = r;
        } else {
            omin = mk(fminf(omin.x, r.x), fminf(omin.y, r.y), fminf(omin.z, r.z));
            omax = mk(fmaxf(omax.x, r.x), fmaxf(omax.y, r.y), fmaxf(omax.z, r.z));
        }
    }
}

// Packs host-layout staging arrays into the device SoA.  idx == nullptr: body i = thread i.
__global__ void k_pack(uint32_t k, const uint32_t* __restrict__ idx, const float* __restrict__ pos3,
                       const float* __restrict__ rot4, const float* __restrict__ scale3, const float* __restrict__ vel3,
                       const float* __restrict__ acc3, const float* __restrict__ bmin3, const float* __restrict__ bmax3,
                       const float* __restrict__ mass, const float* __restrict__ rest, const uint8_t* __restrict__ stat,
                       const uint32_t* __restrict__ entity, const uint32_t* __restrict__ rank, Bodies B,
                       float4* A, float* ext, int shape_only,
                       const uint32_t* __restrict__ slot_of, float4* __restrict__ rawbox /* by body id */) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= k) return;
    const uint32_t i = idx ? slot_of[idx[t]] : t;  // bodies are re-sorted every step: caller index -> current slot
    const V3 bmin = mk(bmin3[3 * t], bmin3[3 * t + 1], bmin3[3 * t + 2]);
    const V3 bmax = mk(bmax3[3 * t], bmax3[3 * t + 1], bmax3[3 * t + 2]);
    const float4 q = rot4 ? make_float4(rot4[4 * t], rot4[4 * t + 1], rot4[4 * t + 2], rot4[4 * t + 3])
                          : make_float4(0.f, 0.f, 0.f, 1.f);
    const V3 sc = scale3 ? mk(scale3[3 * t], scale3[3 * t + 1], scale3[3 * t + 2]) : mk(1.f, 1.f, 1.f);
    V3 omin, omax;
    corner_offsets(bmin, bmax, q, sc, omin, omax);
    const uint32_t fl_static = stat[t] ? F_STATIC : 0u;
    if (shape_only) {
        float4 p = B.P(i), v = B.V(i), a = A[i], lo = B.Lo(i), hi = B.Hi(i);
        p.w = mass[t];
        v.w = rest[t];
        a.w = __uint_as_float((__float_as_uint(a.w) & ~F_STATIC) | fl_static);
        lo = make_float4(omin.x, omin.y, omin.z, lo.w);
        hi = make_float4(omax.x, omax.y, omax.z, hi.w);
        B.P(i) = p, B.V(i) = v, A[i] = a, B.Lo(i) = lo, B.Hi(i) = hi;
    } else {
        B.P(i) = make_float4(pos3[3 * t], pos3[3 * t + 1], pos3[3 * t + 2], mass[t]);
        B.V(i) = make_float4(vel3[3 * t], vel3[3 * t + 1], vel3[3 * t + 2], rest[t]);
        A[i] = make_float4(acc3 ? acc3[3 * t] : 0.f, acc3 ? acc3[3 * t + 1] : 0.f, acc3 ? acc3[3 * t + 2] : 0.f,
                           __uint_as_float(fl_static));
        B.Lo(i) = make_float4(omin.x, omin.y, omin.z, __uint_as_float(i));  // .w = body id (caller index)
        (void)entity;
        B.Hi(i) = make_float4(omax.x, omax.y, omax.z, __uint_as_float(rank ? rank[t] : i));
    }
    if (ext) ext[i] = fmaxf(omax.x - omin.x, fmaxf(omax.y - omin.y, omax.z - omin.z));
    // the un-rotated, un-scaled collision box: what the ray cast and the obstacle grid of the reference use
    const uint32_t id = idx ? idx[t] : t;
    rawbox[2 * (size_t)id] = make_float4(bmin.x, bmin.y, bmin.z, 0.f);
    rawbox[2 * (size_t)id + 1] = make_float4(bmax.x, bmax.y, bmax.z, 0.f);
}

// histogram of body extents in quarter-octave buckets (bucket b: ext <= 2^((b-64)/4)), plus bounds of centres
constexpr int EXT_BUCKETS = 160;
__device__ __forceinline__ int ext_bucket(float e) {
    if (!(e > 0.f)) return 0;
    if (isinf(e) || isnan(e)) return EXT_BUCKETS - 1;
    int b = (int)ceilf(log2f(e) * 4.0f) + 64;
    // make the bucket edge an upper bound despite log2f rounding
    if (b >= 0 && b < EXT_BUCKETS && exp2f((float)(b - 64) * 0.25f) < e) ++b;
    return min(max(b, 0), EXT_BUCKETS - 1);
}
__host__ __device__ inline float ext_bucket_edge(int b) { return exp2f((float)(b - 64) * 0.25f); }

// hist[0 .. EXT_BUCKETS) counts every body, hist[EXT_BUCKETS .. 2 EXT_BUCKETS) the dynamic ones only
__global__ void k_ext_hist(uint32_t n, const float* __restrict__ ext, const float4* __restrict__ A,
                           uint32_t* __restrict__ hist) {
    __shared__ uint32_t sh[2 * EXT_BUCKETS];
    for (int i = threadIdx.x; i < 2 * EXT_BUCKETS; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int b = ext_bucket(ext[i]);
        atomicAdd(&sh[b], 1u);
        if (!(__float_as_uint(A[i].w) & F_STATIC)) atomicAdd(&sh[EXT_BUCKETS + b], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * EXT_BUCKETS; i += blockDim.x)
        if (sh[i]) atomicAdd(&hist[i], sh[i]);
}

// order-preserving float <-> uint for atomicMin/Max
__device__ __forceinline__ uint32_t f2ord(float f) {
    uint32_t u = __float_as_uint(f);
    return (u & TOPBIT) ? ~u : (u | TOPBIT);
}
__host__ __device__ inline float ord2f(uint32_t u) {
    uint32_t v = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
#ifdef __CUDA_ARCH__
    return __uint_as_float(v);
#else
    float f;
    memcpy(&f, &v, 4);
    return f;
#endif
}

// classify by extent threshold and reduce the bounds of grid-class body centres
__global__ void k_classify(uint32_t n, float thr, const Bodies B, float4* A,
                           uint32_t* __restrict__ bounds /*6: min xyz, max xyz (ordered uints)*/) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    float lo[3] = {3.0e38f, 3.0e38f, 3.0e38f}, hi[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
    if (i < n) {
        float4 a = A[i];
        uint32_t f = __float_as_uint(a.w) & ~F_BIG;
        const float4 l4 = B.Lo(i), h4 = B.Hi(i);
        const bool big = !(fmaxf(h4.x - l4.x, fmaxf(h4.y - l4.y, h4.z - l4.z)) <= thr);
        if (big) f |= F_BIG;
        a.w = __uint_as_float(f);
        A[i] = a;
        if (!big) {
            const float4 p = B.P(i), l = l4, h = h4;
            const float c[3] = {((l.x + p.x) + (h.x + p.x)) * 0.5f, ((l.y + p.y) + (h.y + p.y)) * 0.5f,
                                ((l.z + p.z) + (h.z + p.z)) * 0.5f};
#pragma unroll
            for (int k = 0; k < 3; ++k)
                if (isfinite(c[k])) lo[k] = hi[k] = c[k];
        }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o));
            hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o));
        }
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            if (lo[k] < 3.0e38f) atomicMin(&bounds[k], f2ord(lo[k]));
            if (hi[k] > -3.0e38f) atomicMax(&bounds[3 + k], f2ord(hi[k]));
        }
    }
}

// The body arrays are physically re-sorted into cell order by every step (K3), so the slot of a body changes.
// Lo[slot].w keeps the body id (the caller's index at upload; the global entity id in multi-GPU mode).
__global__ void k_slot_map(uint32_t n, const Bodies B, uint32_t* __restrict__ slot_of) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) slot_of[__float_as_uint(B.Lo(i).w)] = i;
}

// bulk refresh of mutable state from host-layout arrays (gp_set_state: idx == nullptr, every slot gathers the
// row of its body id;  gp_update_dirty: row t belongs to body idx[t])
__global__ void k_set_state(uint32_t k, const uint32_t* __restrict__ idx, const uint32_t* __restrict__ slot_of,
                            const float* __restrict__ pos3, const float* __restrict__ vel3,
                            const float* __restrict__ acc3, Bodies B, float4* A) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= k) return;
    uint32_t i, r;
    if (idx) {
        i = slot_of[idx[t]], r = t;
    } else {
        i = t, r = __float_as_uint(B.Lo(t).w);
    }
    if (pos3) {
        float4 p = B.P(i);
        p.x = pos3[3 * (size_t)r], p.y = pos3[3 * (size_t)r + 1], p.z = pos3[3 * (size_t)r + 2];
        B.P(i) = p;
    }
    if (vel3) {
        float4 v = B.V(i);
        v.x = vel3[3 * (size_t)r], v.y = vel3[3 * (size_t)r + 1], v.z = vel3[3 * (size_t)r + 2];
        B.V(i) = v;
    }
    if (acc3) {
        float4 a = A[i];
        a.x = acc3[3 * (size_t)r], a.y = acc3[3 * (size_t)r + 1], a.z = acc3[3 * (size_t)r + 2];
        A[i] = a;
    }
}

// device -> host layout.  by_id: row = body id (gp_download); else row = slot (gp_download_owned, + ids)
__global__ void k_unpack(uint32_t n, const Bodies B, float* pos3, float* vel3, uint32_t* ids, int by_id) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t id = __float_as_uint(B.Lo(i).w);
    const size_t r = by_id ? id : i;
    if (pos3) {
        const float4 p = B.P(i);
        pos3[3 * r] = p.x, pos3[3 * r + 1] = p.y, pos3[3 * r + 2] = p.z;
    }
    if (vel3) {
        const float4 v = B.V(i);
        vel3[3 * r] = v.x, vel3[3 * r + 1] = v.y, vel3[3 * r + 2] = v.z;
    }
    if (ids) ids[i] = id;
}

// RigidBody::cap_velocity (physics.rs:506-520; called by MovementController::update_pos, controllers.rs:307, after it
// wrote the player's velocity): horizontal speed limited to 20 by cgmath's normalize() = v * (1 / |v|), vertical
// velocity f32::clamp'ed to +-40.  Row t = body idx[t] (idx == nullptr: every body).
__global__ void k_cap_velocity(uint32_t k, const uint32_t* __restrict__ idx, const uint32_t* __restrict__ slot_of, Bodies B) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= k) return;
    const uint32_t i = idx ? slot_of[idx[t]] : t;
    float4 v = B.V(i);
    const float m = len3(mk(v.x, 0.0f, v.z));
    if (m > MAX_HORIZONTAL_VELOCITY) {
        const float inv = 1.0f / m;
        v.x = (v.x * inv) * MAX_HORIZONTAL_VELOCITY;
        v.z = (v.z * inv) * MAX_HORIZONTAL_VELOCITY;
    }
    if (v.y < -MAX_VERTICAL_VELOCITY) v.y = -MAX_VERTICAL_VELOCITY;
    if (v.y > MAX_VERTICAL_VELOCITY) v.y = MAX_VERTICAL_VELOCITY;
    B.V(i) = v;
}

// ---- pipelined frames (gp_frame_submit): whole-world state refresh and read-back without random sector traffic ----
// "Home" = the per-body constants in BODY-ID order (48 B: min offset | mass, max offset | restitution, flags | rank),
// built once per upload / shape change.  A frame's state arrives in body-id order, so k_set_state_all streams
// staging + home in and writes the records out IN ID ORDER (slot = id), whole 64 B lines, no read-modify-write; the
// step's own K1/K3 then sort them into cell order as they do every step.  (The per-index path, k_set_state, gathers
// 12-byte rows by id and read-modify-writes records at random: 7.6 GB of DRAM reads to ingest 0.6 GB at 16 M bodies.)
__global__ void k_build_home(uint32_t n, const Bodies B, const float4* __restrict__ A, float4* __restrict__ home) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 p = B.P(i), v = B.V(i), lo = B.Lo(i), hi = B.Hi(i), a = A[i];
    const size_t id = __float_as_uint(lo.w);
    home[3 * id] = make_float4(lo.x, lo.y, lo.z, p.w);
    home[3 * id + 1] = make_float4(hi.x, hi.y, hi.z, v.w);
    home[3 * id + 2] = make_float4(__uint_as_float(__float_as_uint(a.w) & (F_STATIC | F_BIG)), hi.w, 0.f, 0.f);
}
__global__ void __launch_bounds__(256)
    k_set_state_all(uint32_t n, const float4* __restrict__ home, const float* __restrict__ pos3,
                    const float* __restrict__ vel3, const float* __restrict__ acc3, Bodies B, float4* __restrict__ A) {
    const uint32_t id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= n) return;
    const float4 h0 = home[3 * (size_t)id], h1 = home[3 * (size_t)id + 1], h2 = home[3 * (size_t)id + 2];
    const size_t r = 3 * (size_t)id;
    st256(&B.P(id), make_float4(pos3[r], pos3[r + 1], pos3[r + 2], h0.w), make_float4(vel3[r], vel3[r + 1], vel3[r + 2], h1.w));
    st256(&B.Lo(id), make_float4(h0.x, h0.y, h0.z, __uint_as_float(id)), make_float4(h1.x, h1.y, h1.z, h2.y));
    A[id] = make_float4(acc3[r], acc3[r + 1], acc3[r + 2], h2.x);
}
// state8[id] = (pos.xyz, vel.xyz, 0, 0): ONE whole 32 B sector per body, so the scattered store needs no
// read-modify-write (two 12-byte rows per body, as gp_download writes them, are partial sectors)
__global__ void __launch_bounds__(256) k_unpack8(uint32_t n, const Bodies B, float4* __restrict__ state8) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 p, v, lo, hi;
    ld256(&B.P(i), p, v);
    ld256(&B.Lo(i), lo, hi);
    const size_t id = __float_as_uint(lo.w);
    st256(state8 + 2 * id, make_float4(p.x, p.y, p.z, v.x), make_float4(v.y, v.z, 0.f, 0.f));
}

// 32 B rows -> 24 B rows (pos.xyz, vel.xyz), streaming: the download then moves 24 B per body instead of 32
__global__ void __launch_bounds__(256) k_rows8_to_6(uint32_t n, const float4* __restrict__ state8, float* __restrict__ state6) {
    const uint32_t id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= n) return;
    float4 a, b;
    ld256(state8 + 2 * (size_t)id, a, b);
    float* o = state6 + 6 * (size_t)id;
    o[0] = a.x, o[1] = a.y, o[2] = a.z, o[3] = a.w, o[4] = b.x, o[5] = b.y;
}

// RigidBody::update_pos, physics.rs:405-462 (a.w carries the flags; static bodies are untouched)
__device__ __forceinline__ void integrate_body(float4& p, float4& v, const float4 a, float dt, float d16, float d18,
                                               float d35) {
    if (__float_as_uint(a.w) & F_STATIC) return;
    const bool falling = v.y < 0.0f;
    const bool near_apex = fabsf(v.y) < 2.0f && v.y > 0.0f;
    const float gm = falling ? FALL_MULTIPLIER : (near_apex ? APEX_MULTIPLIER : 1.0f);
    v.y = v.y - ((GRAVITY_ACCELERATION * gm) * dt);
    const bool accelerating = len3(mk(a.x, a.y, a.z)) > 0.01f;
    const float d = accelerating ? d16 : (falling ? d18 : d35);
    if (falling) {
        v.x = v.x * d;
        v.z = v.z * d;
    } else {
        v.x = v.x * d;
        v.y = v.y * d;
        v.z = v.z * d;
    }
    v.x = v.x + a.x * dt;
    v.z = v.z + a.z * dt;
    if (len3(mk(v.x, v.y, v.z)) < 0.005f) {
        v.x = 0.0f, v.y = 0.0f, v.z = 0.0f;
    }
    p.x = p.x + v.x * dt;
    p.y = p.y + v.y * dt;
    p.z = p.z + v.z * dt;
}


}  // namespace gp
