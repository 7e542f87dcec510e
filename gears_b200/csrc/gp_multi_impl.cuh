// gp_multi_impl.cuh — multi-GPU slabs over NCCL send/recv (included at the end of gp_api.cu).
#pragma once

struct gp_multi {
    gp_world* w;
};
static void gp_multi_destroy(gp_multi* m) { delete m; }
static gp_status gp_multi_step(gp_multi* m, float, const float*, gp_step_stats*) {
    return fail(m->w, GP_ERR_STATE, "multi-GPU stepping is not available in this build");
}
static gp_status gp_multi_step_async(gp_multi* m, float, const float*, uint32_t) {
    return fail(m->w, GP_ERR_STATE, "multi-GPU stepping is not available in this build");
}
static gp_status gp_multi_sync(gp_multi* m, gp_step_stats*) {
    return fail(m->w, GP_ERR_STATE, "multi-GPU stepping is not available in this build");
}

extern "C" gp_status gp_comm_unique_id(uint8_t id[GP_UNIQUE_ID_BYTES]) {
    memset(id, 0, GP_UNIQUE_ID_BYTES);
    return GP_ERR_NCCL;
}
extern "C" gp_status gp_comm_init(gp_world* w, const uint8_t*, int32_t, int32_t, float, float) {
    return fail(w, GP_ERR_NCCL, "multi-GPU slabs are not available in this build");
}
extern "C" gp_status gp_owned_count(gp_world* w, uint32_t* n) {
    if (!w || !n) return GP_ERR_BAD_ARG;
    *n = w->n;
    return GP_OK;
}
extern "C" gp_status gp_download_owned(gp_world* w, uint32_t*, float*, float*) {
    return fail(w, GP_ERR_STATE, "multi-GPU slabs are not available in this build");
}
