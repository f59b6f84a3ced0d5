// The Level-3 inner scan of subcycleGOMELT as one C-ABI call (gomelt_l3_substeps_f32) and the batched
// source tables it uses.  Host-side orchestration only: every arithmetic step is one of the kernels of
// k_level_step.cu / k_aux.cu / k_transfer.cu, launched back to back on the caller's stream with no
// host synchronisation (the reference runs this loop as a jax.lax.scan, cF:3408 / 3586).
#include <math.h>

#include "common.cuh"

using namespace gomelt;

extern "C" int gomelt_l3_substeps_f32(const gomelt_props_t* props, const gomelt_substeps_args_t* a, void* stream) {
    if (!props || !a || !a->x || !a->y || !a->z || !a->rows || !a->T_in || !a->T_a || !a->T_b || !a->S1 || !a->tables) {
        set_error("gomelt_l3_substeps_f32: NULL argument");
        return GOMELT_E_NULL;
    }
    if (a->n < 1 || a->n > GOMELT_MAX_SUBSTEPS) {
        set_error("gomelt_l3_substeps_f32: n = %d outside 1..%d", a->n, GOMELT_MAX_SUBSTEPS);
        return GOMELT_E_SIZE;
    }
    if (a->T_a == a->T_in || a->T_a == a->T_b) {
        set_error("gomelt_l3_substeps_f32: T_a must differ from T_in and T_b");
        return GOMELT_E_FLAGS;
    }
    if (a->faces && !(a->faces_n > 0.f)) {
        set_error("gomelt_l3_substeps_f32: faces given but faces_n = %g", (double)a->faces_n);
        return GOMELT_E_FLAGS;
    }
    const gomelt_grid_t& g = a->grid;
    float coef[GOMELT_MAX_SUBSTEPS];
    int rc = gomelt_source_tables_batch_f32(props, &g, a->x, a->y, a->z, a->rows, a->n, a->tables, coef, stream);
    if (rc) return rc;
    const long long nface = gomelt_faces_count(g.nx, g.ny, g.nz);
    const bool compact_faces = a->faces && a->faces_scratch;
    if (compact_faces) {  // both parent interpolants at the face nodes, once for the whole block
        if (!a->faces->u2) {
            set_error("gomelt_l3_substeps_f32: faces_scratch needs faces->u2 (parent_old)");
            return GOMELT_E_NULL;
        }
        rc = gomelt_faces_gather_f32(a->faces, a->faces_scratch, a->faces_scratch + nface, stream);
        if (rc) return rc;
    }
    const size_t stride = (size_t)g.nx + g.ny + g.nz;
    const float* Tin = a->T_in;
    for (int i = 0; i < a->n; ++i) {
        float* Tout = (i & 1) ? a->T_b : a->T_a;
        const float* row = a->rows + 7 * (size_t)i;
        gomelt_step_args_t s = {};
        s.grid = g;
        s.T0 = Tin;
        s.S1 = (i == 0 && a->S1_in) ? a->S1_in : a->S1;
        const float* tb = a->tables + stride * i;
        s.src_x = tb;
        s.src_y = tb + g.nx;
        s.src_z = tb + g.nx + g.ny;
        s.src_coef = coef[i];
        s.dt = row[5];
        s.nz_active = g.nz;
        s.n_substrate = a->n_substrate;
        s.flags = (a->flags & (GOMELT_STEP_CLAMP | GOMELT_STEP_SKIP_FACES | GOMELT_STEP_WRITE_S2 | GOMELT_STEP_ACCUM)) |
                  GOMELT_STEP_WRITE_S1 | GOMELT_STEP_FUSED_FLUX;
        s.T_out = Tout;
        s.S1_out = a->S1;
        s.S2_out = a->S2;
        s.S2_prev = a->S2;
        s.accum = a->accum;
        s.max_accum = a->max_accum;
        s.bk_queue = a->bk_queue;
        s.bk_queue_words = a->bk_queue_words;
        s.bk_queue_keep = i > 0 ? 1 : 0;   // (the first substep zeroes the header, every sweep leaves it zeroed)
        if (a->step_events) cudaEventRecord((cudaEvent_t)a->step_events[2 * i], (cudaStream_t)stream);
        rc = gomelt_level_step_f32(props, &s, stream);
        if (a->step_events) cudaEventRecord((cudaEvent_t)a->step_events[2 * i + 1], (cudaStream_t)stream);
        if (rc) return rc;
        if (compact_faces) {
            const float alpha = (float)(i + 1) / a->faces_n;  // cF:3386-3387, float32 like the traced scalars
            rc = gomelt_faces_blend_f32(a->faces_scratch, a->faces_scratch + nface, g.nx, g.ny, g.nz, alpha, 1.0f - alpha,
                                        a->faces->has_clamp, a->faces->clamp_min, Tout, stream);
            if (rc) return rc;
        } else if (a->faces) {
            gomelt_interp_args_t f = *a->faces;
            f.alpha = (float)(i + 1) / a->faces_n;  // cF:3386-3387, float32 like the traced scalars
            f.beta = 1.0f - f.alpha;
            f.faces_only = 1;
            f.mode = GOMELT_INTERP_SET;
            f.out = Tout;
            rc = gomelt_interp_f32(&f, stream);
            if (rc) return rc;
        }
        Tin = Tout;
    }
    if (a->T_last) *a->T_last = (float*)Tin;
    return 0;
}
