// gp_next_kernels.cuh — next rows (instances, ray cast, obstacle grid) and pair-set export.
// Part of the physics step; see gp_kernels.cuh for the map of kernels to reference lines.
#pragma once
#include "gp_common.cuh"

namespace gp {

// ---------------------------------------------------------------------------------------------
// "Next" rows: consumers of the device-resident body state right after the step.
// ---------------------------------------------------------------------------------------------
// N1  pass 3 of update_physics (update_systems.rs:489-527): Instance::to_raw (gears-renderer instance.rs:22-31) for
// every body in one launch instead of one queue.write_buffer per entity.  cgmath 0.18 semantics, no FMA:
// model = T(pos) * M4(q) * S(scale), normal = M3(q); column j of A*R = ((a*R[j][0] + b*R[j][1]) + c*R[j][2]) + d*R[j][3].
__device__ __forceinline__ void mat4_mul(const float* A, const float* R, float* O) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float acc = A[0 * 4 + i] * R[j * 4 + 0];
            acc = acc + A[1 * 4 + i] * R[j * 4 + 1];
            acc = acc + A[2 * 4 + i] * R[j * 4 + 2];
            acc = acc + A[3 * 4 + i] * R[j * 4 + 3];
            O[j * 4 + i] = acc;
        }
}
__global__ void __launch_bounds__(128)
    k_instances(uint32_t n, const Bodies B, const float* __restrict__ rot4, const float* __restrict__ scale3,
                float* __restrict__ out25) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t id = __float_as_uint(B.Lo(i).w);
    const float4 p = B.P(i);
    const float x = rot4 ? rot4[4 * (size_t)id] : 0.f, y = rot4 ? rot4[4 * (size_t)id + 1] : 0.f;
    const float z = rot4 ? rot4[4 * (size_t)id + 2] : 0.f, w = rot4 ? rot4[4 * (size_t)id + 3] : 1.f;
    const float sx = scale3 ? scale3[3 * (size_t)id] : 1.f, sy = scale3 ? scale3[3 * (size_t)id + 1] : 1.f;
    const float sz = scale3 ? scale3[3 * (size_t)id + 2] : 1.f;
    const float x2 = x + x, y2 = y + y, z2 = z + z;
    const float xx2 = x2 * x, xy2 = x2 * y, xz2 = x2 * z, yy2 = y2 * y, yz2 = y2 * z, zz2 = z2 * z;
    const float sy2 = y2 * w, sz2 = z2 * w, sx2 = x2 * w;
    const float r3[9] = {1.0f - yy2 - zz2, xy2 + sz2, xz2 - sy2, xy2 - sz2, 1.0f - xx2 - zz2,
                         yz2 + sx2,        xz2 + sy2, yz2 - sx2, 1.0f - xx2 - yy2};
    const float T[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, p.x, p.y, p.z, 1};
    const float R[16] = {r3[0], r3[1], r3[2], 0, r3[3], r3[4], r3[5], 0, r3[6], r3[7], r3[8], 0, 0, 0, 0, 1};
    const float S[16] = {sx, 0, 0, 0, 0, sy, 0, 0, 0, 0, sz, 0, 0, 0, 0, 1};
    float TR[16], M[16];
    mat4_mul(T, R, TR);
    mat4_mul(TR, S, M);
    float* o = out25 + 25 * (size_t)id;
#pragma unroll
    for (int k = 0; k < 16; ++k) o[k] = M[k];
#pragma unroll
    for (int k = 0; k < 9; ++k) o[16 + k] = r3[k];
}

// N3  Weapon::ray_intersects_aabb (gears-ecs interactive.rs:94-158), batched: thread = (ray, target)
__global__ void __launch_bounds__(256)
    k_raycast(uint32_t nrays, const float* __restrict__ origin3, const float* __restrict__ dir3, uint32_t ntargets,
              const uint32_t* __restrict__ targets, const uint32_t* __restrict__ slot_of, const Bodies B,
              const float4* __restrict__ rawbox, uint8_t* __restrict__ hits) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)nrays * ntargets) return;
    const uint32_t r = (uint32_t)(t / ntargets), k = (uint32_t)(t % ntargets);
    const uint32_t id = targets ? targets[k] : k;
    const float4 p = B.P(slot_of[id]);
    const float4 bmin = rawbox[2 * (size_t)id], bmax = rawbox[2 * (size_t)id + 1];
    const float o[3] = {origin3[3 * r], origin3[3 * r + 1], origin3[3 * r + 2]};
    const float d[3] = {dir3[3 * r], dir3[3 * r + 1], dir3[3 * r + 2]};
    const float mn[3] = {p.x + bmin.x, p.y + bmin.y, p.z + bmin.z};
    const float mx[3] = {p.x + bmax.x, p.y + bmax.y, p.z + bmax.z};
    float tmin = 0.0f, tmax = 3.40282347e+38f;
    bool hit = true;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        if (!hit) break;
        if (fabsf(d[i]) < 1e-8f) {
            if (o[i] < mn[i] || o[i] > mx[i]) hit = false;
        } else {
            const float inv = 1.0f / d[i];
            float t1 = (mn[i] - o[i]) * inv, t2 = (mx[i] - o[i]) * inv;
            if (t1 > t2) {
                const float tt = t1;
                t1 = t2, t2 = tt;
            }
            tmin = fmaxf(tmin, t1);
            tmax = fminf(tmax, t2);
            if (tmin > tmax) hit = false;
        }
    }
    hits[t] = (hit && tmax >= 0.0f && tmin <= tmax) ? 1 : 0;
}

// N4  PathfindingGrid::add_obstacle (gears-ecs pathfinding.rs:296-338, GridPos::from_world_pos :57-63): rasterise the
// collision boxes of k bodies into a dense bitmap; one warp per obstacle, lanes stride over its cells.
__device__ __forceinline__ int f2i_sat(float f) { return __float2int_rz(f); }  // saturating, NaN -> 0 (= Rust `as i32`)
__global__ void __launch_bounds__(256)
    k_obstacle_grid(uint32_t k, const uint32_t* __restrict__ idx, const uint32_t* __restrict__ slot_of, const Bodies B,
                    const float4* __restrict__ rawbox, float cell, int ox, int oy, int oz, uint32_t dx, uint32_t dy,
                    uint32_t dz, uint32_t* __restrict__ bitmap, int* __restrict__ bounds /*6, pre-set to +-INT*/) {
    const uint32_t wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (wid >= k) return;
    const uint32_t id = idx ? idx[wid] : wid;
    const float4 p = B.P(slot_of[id]);
    const float4 bmin = rawbox[2 * (size_t)id], bmax = rawbox[2 * (size_t)id + 1];
    const int lo[3] = {f2i_sat(roundf((p.x + bmin.x) / cell)), f2i_sat(roundf((p.y + bmin.y) / cell)),
                       f2i_sat(roundf((p.z + bmin.z) / cell))};
    const int hi[3] = {f2i_sat(roundf((p.x + bmax.x) / cell)), f2i_sat(roundf((p.y + bmax.y) / cell)),
                       f2i_sat(roundf((p.z + bmax.z) / cell))};
    if (lane == 0) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {  // the reference folds the corner cells into its bounds unconditionally
            atomicMin(&bounds[a], lo[a]);
            atomicMax(&bounds[3 + a], hi[a]);
        }
    }
    // clip to the bitmap window
    const long long x0 = max((long long)lo[0], (long long)ox), x1 = min((long long)hi[0], (long long)ox + dx - 1);
    const long long y0 = max((long long)lo[1], (long long)oy), y1 = min((long long)hi[1], (long long)oy + dy - 1);
    const long long z0 = max((long long)lo[2], (long long)oz), z1 = min((long long)hi[2], (long long)oz + dz - 1);
    if (x0 > x1 || y0 > y1 || z0 > z1) return;
    const unsigned long long nx = x1 - x0 + 1, ny = y1 - y0 + 1, nz = z1 - z0 + 1, tot = nx * ny * nz;
    for (unsigned long long c = lane; c < tot; c += 32) {
        const unsigned long long cx = c % nx, cy = (c / nx) % ny, cz = c / (nx * ny);
        const unsigned long long bit = (unsigned long long)(x0 - ox + cx) +
                                       (unsigned long long)dx * ((y0 - oy + cy) + (unsigned long long)dy * (z0 - oz + cz));
        atomicOr(&bitmap[bit >> 5], 1u << (bit & 31u));
    }
}

// ---- pair-set export (gp_pairs): compact flagged pairs into (key = rank_a<<32|rank_b, value = pair id) ----
__global__ void k_export_pairs(const uint4* __restrict__ pairs, uint32_t np, int which,
                               unsigned long long* __restrict__ keys, uint32_t* __restrict__ vals, uint32_t* counter) {
    const uint32_t pid = blockIdx.x * blockDim.x + threadIdx.x;
    if (pid >= np) return;
    const uint4 r = pairs[2 * (size_t)pid];
    const bool in = which == 0 ? (r.z & TOPBIT) != 0 : (r.w & TOPBIT) != 0;
    if (!in) return;
    cg::coalesced_group g = cg::coalesced_threads();
    uint32_t base = 0;
    if (g.thread_rank() == 0) base = atomicAdd(counter, g.size());
    base = g.shfl(base, 0);
    const uint32_t s = base + g.thread_rank();
    keys[s] = ((unsigned long long)(r.z & ~TOPBIT) << 32) | (r.w & ~TOPBIT);
    vals[s] = pid;
}

__global__ void k_gather_pairs(const uint4* __restrict__ pairs, const uint32_t* __restrict__ order, uint32_t m,
                               const Bodies B, uint2* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const uint4 r = pairs[2 * (size_t)order[i]];
    // pair records hold slots of the current (cell-sorted) set; report body ids
    out[i] = make_uint2(__float_as_uint(B.Lo(r.x & ~TOPBIT).w), __float_as_uint(B.Lo(r.y & ~TOPBIT).w));
}


}  // namespace gp
