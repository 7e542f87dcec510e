// gp_multi.cuh — multi-GPU (spatial slab) extension: forward declarations used by gp_api.cu.
#pragma once
#include "../../include/gears_physics_cuda.h"

struct gp_multi;
static void gp_multi_destroy(gp_multi* m);
static gp_status gp_multi_step(gp_multi* m, float dt, const float* damp3, gp_step_stats* stats);
static gp_status gp_multi_step_async(gp_multi* m, float dt, const float* damp3, uint32_t nsteps);
static gp_status gp_multi_sync(gp_multi* m, gp_step_stats* stats);
