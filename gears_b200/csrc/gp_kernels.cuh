// gp_kernels.cuh — the physics step as sm_100a CUDA kernels.
//
// Reference behaviour reproduced (paths relative to /root/reference):
//   K1  k_keys            RigidBody::update_pos (in registers)   crates/gears-ecs/src/components/physics.rs:405-462
//                         + world AABB centre -> cell key + cell histogram; slabs: migration / halo packing
//   K2  (scan.cuh)        (new) in-place scan of the histogram = cell start table (one-pass bucket sort by cell)
//   K3  k_scatter_cells   update_pos again on the fly + world AABB (get_rotated_aabb, physics.rs:84-135, via per-body
//                         min/max offsets); bodies are physically re-sorted into cell order
//   K4  k_find_pairs      CollisionBox::intersects                physics.rs:148-165  (margin-inflated superset; a warp
//                         per tile of 32 slots, 256-bit record loads, flattened tests, per-warp staged bursts; exact
//                         snapshot test for S_snapshot)
//   K5  k_fill_chains / k_sort_chains / k_sort_heavy   (new) per-body contact chains in the reference's pair order
//   K6  k_contacts        ONE cooperative launch (gp_contact_kernel.cuh), phases separated by grid barriers:
//       sweep_phase       check_and_resolve_collision             physics.rs:474-499 -> intersects + resolve (physics.rs:176-299)
//                         in the order of update_systems.rs:408-487
//       requery_phase / verify_watch_phase                        (new) verification of the speculative pair set
//       tiny_link / tiny_walk / tiny_exec phases                  (new) repair, small components: one warp re-executes a
//                                                                  component's pairs in (rank_a, rank_b) order on chip
//       repair_local_phase + sweep_phase (list mode)              (new) repair, large components
//       taint_phase       (new) multi-GPU: which owned bodies are provably independent of unseen bodies (union-find)
//       finish_step       (new) margin adaptation, counters, stats snapshot
//
// Ordering argument: the reference visits pairs (i,j), i<j in iteration order,
// lexicographically.  Only pairs that share a DYNAMIC body interact (static bodies are never written,
// physics.rs:245-270).  For a body x the pairs involving x, in reference order, are exactly its
// partners sorted by partner rank.  So executing a pair when it is at the head of both of its
// dynamic endpoints' chains (Kahn peeling, one round per dependency level) reproduces the sequential
// sweep bit for bit; pairs within a round touch disjoint dynamic bodies.
//
// All arithmetic: IEEE binary32, compiled with --fmad=false (Rust never contracts), IEEE div/sqrt.
//
// Files: gp_common (constants, control block, helpers), gp_mirror_kernels (upload / refresh / download, update_pos),
// gp_cellsort_kernels (K1, K3, slab exchange helpers), gp_pairfind_kernels (K4), gp_chain_kernels (K5),
// gp_sweep_kernels (K6), gp_repair_kernels (verification, repair, end of step), gp_slab_kernels (multi-GPU proof),
// gp_next_kernels (instances, ray cast, obstacle grid, pair export).
#pragma once
#include "gp_common.cuh"
#include "gp_mirror_kernels.cuh"
#include "gp_cellsort_kernels.cuh"
#include "gp_pairfind_kernels.cuh"
#include "gp_chain_kernels.cuh"
#include "gp_sweep_kernels.cuh"
#include "gp_repair_kernels.cuh"
#include "gp_slab_kernels.cuh"
#include "gp_contact_kernel.cuh"
#include "gp_next_kernels.cuh"
