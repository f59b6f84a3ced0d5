/*
 * gomelt_xla_ffi.h - the XLA FFI custom-call layer over the C ABI of gomelt_abi.h (BASELINE.json north_star: "hand-written
 * sm_100a CUDA kernels registered through a thin C-ABI XLA FFI custom-call layer").
 *
 * One handler per entry point, `XLA_FFI_Error* Gomelt<Name>Ffi(XLA_FFI_CallFrame*)`, written against the XLA FFI *C* API
 * only (xla/ffi/api/c_api.h).  Register with
 *     jax.ffi.register_ffi_target("gomelt_<name>", jax.ffi.pycapsule(lib.Gomelt<Name>Ffi), platform="CUDA")
 * and call with jax.ffi.ffi_call (gomelt_b200/xla_ffi.py builds the attributes).
 *
 * Calling convention (the same for every handler, so that the layer stays a forwarding layer and the ABI structs remain
 * the single source of truth):
 *   - attribute "args" (u8 array) = the bytes of the entry point's argument struct - the gomelt_*_args_t of gomelt_abi.h,
 *     or, for the entry points that take plain parameters, the gomelt_ffi_*_t below - in which every DEVICE POINTER field
 *     holds a *slot code* instead of an address (as an int64 in the pointer's 8 bytes):
 *         0 = NULL,  k > 0 = operand buffer k-1,  k < 0 = result buffer -k-1;
 *     in-place fields name a result that the JAX side aliases to the corresponding operand (input_output_aliases);
 *   - attribute "props" (u8 array) = gomelt_props_t, where the entry point takes one;
 *   - host-side arrays of the C ABI travel as attributes: "rows" (f32 array: toolpath rows [n][7]), "laser" (f32[3]);
 *     they are compile-time constants of the XLA computation - the toolpath of a block is known when the block is
 *     traced, exactly as the reference traces it (static args of jax.jit);
 *   - peer-memory addresses (gomelt_step_args_t.peer_*, halo_sync*) are not XLA buffers: they are passed through
 *     unchanged;
 *   - host out-parameters (coefficients a caller can compute itself, T_last, l1_in_spare) are not returned: the
 *     handlers document where the result lands instead.
 * Stages other than EXECUTE are no-ops; the metadata extension is answered with API version XLA_FFI_API_MAJOR.MINOR.
 */
#ifndef GOMELT_XLA_FFI_H
#define GOMELT_XLA_FFI_H

#include "gomelt_abi.h"

#ifdef __cplusplus
extern "C" {
#endif

/* argument structs of the entry points that take plain parameters in gomelt_abi.h (pointer fields = slot codes) */
typedef struct { const float *T, *S1; int64_t nn, n_substrate; float *S1_out; uint8_t *S2_out; float *k_out, *rhocp_out; } gomelt_ffi_state_props_t;
typedef struct { gomelt_grid_t grid; const float *T0; int32_t nz_active, add; float *flux; } gomelt_ffi_surface_flux_t;
typedef struct { gomelt_grid_t grid; const float *x, *y, *z; float laserP; float *tx, *ty, *tz; } gomelt_ffi_source_tables_t;
typedef struct { gomelt_grid_t grid; const float *x, *y, *z; int32_t n; float *tables; } gomelt_ffi_source_tables_batch_t;
typedef struct { const void *src; void *dst; int32_t elem_size; const int32_t *ix, *iy, *iz; int32_t nx, ny, nz, big_nx, big_ny, scatter; } gomelt_ffi_box_copy_t;
typedef struct { const float *src; int32_t sdims[3], slo[3]; float *dst; int32_t ddims[3], dlo[3], n[3]; } gomelt_ffi_patch_copy_t;
typedef struct { float *F; const float *tx, *ty, *tz; int32_t nx, ny, nz; float coef; int32_t accumulate; } gomelt_ffi_rank1_t;
typedef struct { gomelt_axis_t fine[3], parent[3]; float laserP; float *tx, *ty, *tz; } gomelt_ffi_coarse_source_tables_t;
typedef struct { gomelt_axis_t fine[3], parent[3]; float wq_fine; int32_t n; float *tables, *F; int32_t accumulate; } gomelt_ffi_projected_source_t;
typedef struct { const float *face_a, *face_b; int32_t ntx, nty, ntz; float alpha, beta; int32_t has_clamp; float clamp_min; float *out; } gomelt_ffi_faces_blend_t;
typedef struct { gomelt_interp_args_t interp; float *face_a, *face_b; } gomelt_ffi_faces_gather_t;
typedef struct { const float *x; int64_t n; float *out3; } gomelt_ffi_minmax_t;
typedef struct { float *x; int64_t n; float lo; } gomelt_ffi_clamp_min_t;
typedef struct { const float *T3; const uint8_t *resetmask; float dt, T_liquidus; float *accum0, *max_accum0; const int32_t *ix, *iy, *iz;
                 int32_t nx, ny, nz, big_nx, big_ny; } gomelt_ffi_accum_single_step_t;
typedef struct { gomelt_substeps_args_t substeps; gomelt_interp_args_t faces; int32_t has_faces; } gomelt_ffi_l3_substeps_t;
typedef struct { gomelt_hier_t hier; int32_t N2, N3; float *max_accum, *accum; } gomelt_ffi_subcycle_t;   /* + attribute "rows" */
typedef struct { gomelt_hier_t hier; uint8_t *resetmask; } gomelt_ffi_step_t;                              /* + attribute "rows" (one row) */
typedef struct { gomelt_hier_t hier; float dt; } gomelt_ffi_dwell_step_t;

/* The handlers (XLA_FFI_Handler signature; declared with void* so that this header does not need c_api.h). */
#define GOMELT_FFI_HANDLERS(X) \
    X(GomeltLevelStepFfi) X(GomeltStatePropsFfi) X(GomeltSurfaceFluxFfi) X(GomeltSourceTablesFfi) X(GomeltSourceTablesBatchFfi) \
    X(GomeltInterpFfi) X(GomeltFacesGatherFfi) X(GomeltFacesBlendFfi) X(GomeltBoxCopyFfi) X(GomeltRank1Ffi) \
    X(GomeltCoarseSourceTablesFfi) X(GomeltProjectedSourceFfi) X(GomeltProjectFfi) X(GomeltShiftWindowFfi) X(GomeltClampMinFfi) \
    X(GomeltMinMaxFfi) X(GomeltAccumSingleStepFfi) X(GomeltL3SubstepsFfi) X(GomeltSubcycleFfi) X(GomeltStepFfi) X(GomeltDwellStepFfi) \
    X(GomeltPatchCopyFfi)
#define GOMELT_FFI_DECLARE(name) void* name(void* call_frame);
GOMELT_FFI_HANDLERS(GOMELT_FFI_DECLARE)
#undef GOMELT_FFI_DECLARE

/* number of handlers above / their names (for registration loops) */
int gomelt_xla_ffi_handler_count(void);
const char* gomelt_xla_ffi_handler_name(int i);

#ifdef __cplusplus
}
#endif
#endif /* GOMELT_XLA_FFI_H */
