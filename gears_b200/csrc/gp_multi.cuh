// gp_multi.cuh — multi-GPU (spatial slab) extension: what gp_api.cu needs to know about it.
#pragma once
#include <cuda_runtime.h>

#include "../../include/gears_physics_cuda.h"

struct gp_multi;
static void gp_multi_destroy(gp_multi* m);
static float4* gp_multi_send(gp_multi* m, int side);  // side 0 = left neighbour, 1 = right; nullptr if none
static float4* gp_multi_recv(gp_multi* m, int side);
static gp_status gp_multi_enqueue_exchange(gp_multi* m);     // NCCL transport: send/recv on the step stream
static bool gp_multi_is_local(gp_multi* m);
static void gp_multi_begin_step(gp_multi* m);
static void gp_multi_seal(gp_multi* m);
static void gp_multi_before_append(gp_multi* m);
static void gp_multi_after_append(gp_multi* m);
