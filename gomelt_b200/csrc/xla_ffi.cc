// XLA FFI handlers over the C ABI (include/gomelt_xla_ffi.h has the calling convention; include/gomelt_abi.h the
// entry points).  Written against the XLA FFI C API only (xla/ffi/api/c_api.h): the real header when a jaxlib is
// installed (build.py puts its include directory first), else the subset in third_party/xla_ffi_min.  A forwarding
// layer: every handler copies the argument struct out of the "args" attribute, replaces the slot codes in its device
// pointer fields by the addresses of the call frame's operand / result buffers, takes the stream from the execution
// context and calls the C-ABI entry point.  No arithmetic, no allocation, no synchronisation.
#include <stddef.h>
#include <string.h>

#include <string>

#include "gomelt_xla_ffi.h"
#include "xla/ffi/api/c_api.h"

namespace {

struct Frame {
    XLA_FFI_CallFrame* f;
    std::string err;

    XLA_FFI_Error* fail(XLA_FFI_Error_Code code, const std::string& msg) const {
        XLA_FFI_Error_Create_Args a;
        a.struct_size = XLA_FFI_STRUCT_SIZE(XLA_FFI_Error_Create_Args, errc);
        a.extension_start = nullptr;
        a.message = msg.c_str();
        a.errc = code;
        return f->api->XLA_FFI_Error_Create(&a);
    }
    XLA_FFI_Error* from_rc(int rc, const char* what) const {
        if (rc == 0) return nullptr;
        const char* m = gomelt_last_error();
        return fail(rc < 0 ? XLA_FFI_Error_Code_INVALID_ARGUMENT : XLA_FFI_Error_Code_INTERNAL,
                    std::string(what) + ": " + (m ? m : ""));
    }
    // attribute by name: ARRAY -> (data, element count, dtype); SCALAR -> (value, 1, dtype)
    bool attr(const char* name, const void** data, size_t* count, XLA_FFI_DataType* dtype) const {
        const size_t n = strlen(name);
        for (int64_t i = 0; i < f->attrs.size; ++i) {
            const XLA_FFI_ByteSpan* s = f->attrs.names[i];
            if (s->len != n || memcmp(s->ptr, name, n) != 0) continue;
            if (f->attrs.types[i] == XLA_FFI_AttrType_ARRAY) {
                const XLA_FFI_Array* a = static_cast<const XLA_FFI_Array*>(f->attrs.attrs[i]);
                *data = a->data; *count = a->size; *dtype = a->dtype;
                return true;
            }
            if (f->attrs.types[i] == XLA_FFI_AttrType_SCALAR) {
                const XLA_FFI_Scalar* a = static_cast<const XLA_FFI_Scalar*>(f->attrs.attrs[i]);
                *data = a->value; *count = 1; *dtype = a->dtype;
                return true;
            }
            return false;
        }
        return false;
    }
    // copy a POD struct out of a u8-array attribute
    bool blob(const char* name, void* dst, size_t size) {
        const void* d = nullptr;
        size_t n = 0;
        XLA_FFI_DataType t = XLA_FFI_DataType_INVALID;
        if (!attr(name, &d, &n, &t) || (t != XLA_FFI_DataType_U8 && t != XLA_FFI_DataType_S8) || n != size) {
            err = std::string("attribute \"") + name + "\": expected a u8 array of " + std::to_string(size) +
                  " bytes (ABI version " + std::to_string(GOMELT_ABI_VERSION) + "), got " + std::to_string(n);
            return false;
        }
        memcpy(dst, d, size);
        return true;
    }
    const float* f32_array(const char* name, size_t count) {
        const void* d = nullptr;
        size_t n = 0;
        XLA_FFI_DataType t = XLA_FFI_DataType_INVALID;
        if (!attr(name, &d, &n, &t) || t != XLA_FFI_DataType_F32 || n != count) {
            err = std::string("attribute \"") + name + "\": expected an f32 array of " + std::to_string(count) + " elements";
            return nullptr;
        }
        return static_cast<const float*>(d);
    }
    // slot code -> device address
    bool resolve(void** field) {
        int64_t code;
        memcpy(&code, field, sizeof code);
        if (code == 0) {
            *field = nullptr;
            return true;
        }
        if (code > 0) {
            if (code > f->args.size) { err = "slot code " + std::to_string(code) + " beyond the operands"; return false; }
            *field = static_cast<XLA_FFI_Buffer*>(f->args.args[code - 1])->data;
            return true;
        }
        if (-code > f->rets.size) { err = "slot code " + std::to_string(code) + " beyond the results"; return false; }
        *field = static_cast<XLA_FFI_Buffer*>(f->rets.rets[-code - 1])->data;
        return true;
    }
    bool patch(void* base, const size_t* offs, size_t n) {
        for (size_t i = 0; i < n; ++i)
            if (!resolve(reinterpret_cast<void**>(static_cast<char*>(base) + offs[i]))) return false;
        return true;
    }
    XLA_FFI_Error* stream(void** st) const {
        XLA_FFI_Stream_Get_Args a;
        a.struct_size = XLA_FFI_STRUCT_SIZE(XLA_FFI_Stream_Get_Args, stream);
        a.extension_start = nullptr;
        a.ctx = f->ctx;
        a.stream = nullptr;
        XLA_FFI_Error* e = f->api->XLA_FFI_Stream_Get(&a);
        *st = a.stream;
        return e;
    }
};

// registration-time query (metadata extension) and non-execute stages; returns true when the call is finished
bool prologue(XLA_FFI_CallFrame* f) {
    for (XLA_FFI_Extension_Base* e = f->extension_start; e; e = e->next)
        if (e->type == XLA_FFI_Extension_Metadata) {
            XLA_FFI_Metadata* m = reinterpret_cast<XLA_FFI_Metadata_Extension*>(e)->metadata;
            m->api_version.major_version = XLA_FFI_API_MAJOR;
            m->api_version.minor_version = XLA_FFI_API_MINOR;
            m->traits = 0;  // host-side attribute decoding per call: not claimed command-buffer compatible
            return true;
        }
    return f->stage != XLA_FFI_ExecutionStage_EXECUTE;
}

#define OFF(S, m) offsetof(S, m)
#define AXIS3(S, m) OFF(S, m[0].coords), OFF(S, m[1].coords), OFF(S, m[2].coords)
#define NOFF(a) (sizeof(a) / sizeof((a)[0]))

const size_t kStep[] = {OFF(gomelt_step_args_t, T0), OFF(gomelt_step_args_t, S1), OFF(gomelt_step_args_t, rhs),
                        OFF(gomelt_step_args_t, src_x), OFF(gomelt_step_args_t, src_y), OFF(gomelt_step_args_t, src_z),
                        OFF(gomelt_step_args_t, topflux), OFF(gomelt_step_args_t, T_out), OFF(gomelt_step_args_t, S1_out),
                        OFF(gomelt_step_args_t, S2_out), OFF(gomelt_step_args_t, S2_prev), OFF(gomelt_step_args_t, accum),
                        OFF(gomelt_step_args_t, max_accum), OFF(gomelt_step_args_t, bk_queue)};
const size_t kInterp[] = {AXIS3(gomelt_interp_args_t, src), OFF(gomelt_interp_args_t, u), OFF(gomelt_interp_args_t, u2),
                          OFF(gomelt_interp_args_t, tx), OFF(gomelt_interp_args_t, ty), OFF(gomelt_interp_args_t, tz),
                          OFF(gomelt_interp_args_t, map_x), OFF(gomelt_interp_args_t, map_y), OFF(gomelt_interp_args_t, map_z),
                          OFF(gomelt_interp_args_t, base), OFF(gomelt_interp_args_t, out)};
const size_t kProject[] = {AXIS3(gomelt_project_args_t, fine), AXIS3(gomelt_project_args_t, parent), OFF(gomelt_project_args_t, A),
                           OFF(gomelt_project_args_t, A2), OFF(gomelt_project_args_t, coef), OFF(gomelt_project_args_t, first_x),
                           OFF(gomelt_project_args_t, first_y), OFF(gomelt_project_args_t, first_z),
                           OFF(gomelt_project_args_t, cellsum), OFF(gomelt_project_args_t, V), OFF(gomelt_project_args_t, coef_T),
                           OFF(gomelt_project_args_t, coef_S1), OFF(gomelt_project_args_t, wtab_x),
                           OFF(gomelt_project_args_t, wtab_y), OFF(gomelt_project_args_t, wtab_z)};
const size_t kShift[] = {AXIS3(gomelt_shift_args_t, L1), OFF(gomelt_shift_args_t, T1), AXIS3(gomelt_shift_args_t, mid),
                         OFF(gomelt_shift_args_t, Tp_mid), AXIS3(gomelt_shift_args_t, old), OFF(gomelt_shift_args_t, Tp_old),
                         OFF(gomelt_shift_args_t, tx), OFF(gomelt_shift_args_t, ty), OFF(gomelt_shift_args_t, tz),
                         OFF(gomelt_shift_args_t, Tp_new), OFF(gomelt_shift_args_t, T_new)};
const size_t kSubsteps[] = {OFF(gomelt_substeps_args_t, x), OFF(gomelt_substeps_args_t, y), OFF(gomelt_substeps_args_t, z),
                            OFF(gomelt_substeps_args_t, T_in), OFF(gomelt_substeps_args_t, T_a), OFF(gomelt_substeps_args_t, T_b),
                            OFF(gomelt_substeps_args_t, S1_in), OFF(gomelt_substeps_args_t, S1), OFF(gomelt_substeps_args_t, tables),
                            OFF(gomelt_substeps_args_t, S2), OFF(gomelt_substeps_args_t, accum),
                            OFF(gomelt_substeps_args_t, max_accum), OFF(gomelt_substeps_args_t, faces_scratch),
                            OFF(gomelt_substeps_args_t, bk_queue)};
#define LEVEL_PTRS(L) OFF(gomelt_hier_t, L.x), OFF(gomelt_hier_t, L.y), OFF(gomelt_hier_t, L.z), OFF(gomelt_hier_t, L.T0), \
                      OFF(gomelt_hier_t, L.S1), OFF(gomelt_hier_t, L.Tprime0), OFF(gomelt_hier_t, L.S2)
#define PAIR_PTRS(P) OFF(gomelt_hier_t, P.first_x), OFF(gomelt_hier_t, P.first_y), OFF(gomelt_hier_t, P.first_z), \
                     OFF(gomelt_hier_t, P.wtab_x), OFF(gomelt_hier_t, P.wtab_y), OFF(gomelt_hier_t, P.wtab_z)
#define OV_PTRS(O) OFF(gomelt_hier_t, O.ix), OFF(gomelt_hier_t, O.iy), OFF(gomelt_hier_t, O.iz), OFF(gomelt_hier_t, O.cx), \
                   OFF(gomelt_hier_t, O.cy), OFF(gomelt_hier_t, O.cz)
const size_t kHier[] = {LEVEL_PTRS(L1), LEVEL_PTRS(L2), LEVEL_PTRS(L3), PAIR_PTRS(L2L1), PAIR_PTRS(L3L1), PAIR_PTRS(L3L2),
                        OV_PTRS(ov2), OV_PTRS(ov3), OFF(gomelt_hier_t, L0_S1), OFF(gomelt_hier_t, L0_S2), OFF(gomelt_hier_t, l0_ix),
                        OFF(gomelt_hier_t, l0_iy), OFF(gomelt_hier_t, l0_iz), OFF(gomelt_hier_t, l0p_ix), OFF(gomelt_hier_t, l0p_iy),
                        OFF(gomelt_hier_t, l0p_iz), OFF(gomelt_hier_t, L1_spare), OFF(gomelt_hier_t, work)};
const size_t kPatchCopy[] = {OFF(gomelt_ffi_patch_copy_t, src), OFF(gomelt_ffi_patch_copy_t, dst)};

// generic body: A = argument struct type; OFFS = its device-pointer offsets; CALL(frame, args, props, stream) -> rc
#define GOMELT_HANDLER(NAME, A, OFFS, NEEDS_PROPS, WHAT, CALL)                                              \
    extern "C" void* NAME(void* call_frame) {                                                               \
        XLA_FFI_CallFrame* cf = static_cast<XLA_FFI_CallFrame*>(call_frame);                                \
        if (prologue(cf)) return nullptr;                                                                   \
        Frame fr{cf, {}};                                                                                   \
        A a;                                                                                                \
        gomelt_props_t props;                                                                               \
        memset(&props, 0, sizeof props);                                                                    \
        if (!fr.blob("args", &a, sizeof a) || (NEEDS_PROPS && !fr.blob("props", &props, sizeof props)) ||   \
            !fr.patch(&a, OFFS, NOFF(OFFS)))                                                                \
            return fr.fail(XLA_FFI_Error_Code_INVALID_ARGUMENT, std::string(WHAT) + ": " + fr.err);          \
        void* st = nullptr;                                                                                 \
        if (XLA_FFI_Error* e = fr.stream(&st)) return e;                                                    \
        int rc;                                                                                             \
        CALL;                                                                                               \
        return fr.from_rc(rc, WHAT);                                                                        \
    }

const size_t kStateProps[] = {OFF(gomelt_ffi_state_props_t, T), OFF(gomelt_ffi_state_props_t, S1), OFF(gomelt_ffi_state_props_t, S1_out),
                              OFF(gomelt_ffi_state_props_t, S2_out), OFF(gomelt_ffi_state_props_t, k_out),
                              OFF(gomelt_ffi_state_props_t, rhocp_out)};
const size_t kSurfaceFlux[] = {OFF(gomelt_ffi_surface_flux_t, T0), OFF(gomelt_ffi_surface_flux_t, flux)};
const size_t kSourceTables[] = {OFF(gomelt_ffi_source_tables_t, x), OFF(gomelt_ffi_source_tables_t, y), OFF(gomelt_ffi_source_tables_t, z),
                                OFF(gomelt_ffi_source_tables_t, tx), OFF(gomelt_ffi_source_tables_t, ty),
                                OFF(gomelt_ffi_source_tables_t, tz)};
const size_t kSourceTablesBatch[] = {OFF(gomelt_ffi_source_tables_batch_t, x), OFF(gomelt_ffi_source_tables_batch_t, y),
                                     OFF(gomelt_ffi_source_tables_batch_t, z), OFF(gomelt_ffi_source_tables_batch_t, tables)};
const size_t kBoxCopy[] = {OFF(gomelt_ffi_box_copy_t, src), OFF(gomelt_ffi_box_copy_t, dst), OFF(gomelt_ffi_box_copy_t, ix),
                           OFF(gomelt_ffi_box_copy_t, iy), OFF(gomelt_ffi_box_copy_t, iz)};
const size_t kRank1[] = {OFF(gomelt_ffi_rank1_t, F), OFF(gomelt_ffi_rank1_t, tx), OFF(gomelt_ffi_rank1_t, ty), OFF(gomelt_ffi_rank1_t, tz)};
const size_t kCoarseTables[] = {AXIS3(gomelt_ffi_coarse_source_tables_t, fine), AXIS3(gomelt_ffi_coarse_source_tables_t, parent),
                                OFF(gomelt_ffi_coarse_source_tables_t, tx), OFF(gomelt_ffi_coarse_source_tables_t, ty),
                                OFF(gomelt_ffi_coarse_source_tables_t, tz)};
const size_t kProjSource[] = {AXIS3(gomelt_ffi_projected_source_t, fine), AXIS3(gomelt_ffi_projected_source_t, parent),
                              OFF(gomelt_ffi_projected_source_t, tables), OFF(gomelt_ffi_projected_source_t, F)};
const size_t kFacesBlend[] = {OFF(gomelt_ffi_faces_blend_t, face_a), OFF(gomelt_ffi_faces_blend_t, face_b), OFF(gomelt_ffi_faces_blend_t, out)};
const size_t kFacesGather[] = {AXIS3(gomelt_ffi_faces_gather_t, interp.src), OFF(gomelt_ffi_faces_gather_t, interp.u),
                               OFF(gomelt_ffi_faces_gather_t, interp.u2), OFF(gomelt_ffi_faces_gather_t, interp.tx),
                               OFF(gomelt_ffi_faces_gather_t, interp.ty), OFF(gomelt_ffi_faces_gather_t, interp.tz),
                               OFF(gomelt_ffi_faces_gather_t, face_a), OFF(gomelt_ffi_faces_gather_t, face_b)};
const size_t kMinMax[] = {OFF(gomelt_ffi_minmax_t, x), OFF(gomelt_ffi_minmax_t, out3)};
const size_t kClamp[] = {OFF(gomelt_ffi_clamp_min_t, x)};
const size_t kAccum[] = {OFF(gomelt_ffi_accum_single_step_t, T3), OFF(gomelt_ffi_accum_single_step_t, resetmask),
                         OFF(gomelt_ffi_accum_single_step_t, accum0), OFF(gomelt_ffi_accum_single_step_t, max_accum0),
                         OFF(gomelt_ffi_accum_single_step_t, ix), OFF(gomelt_ffi_accum_single_step_t, iy),
                         OFF(gomelt_ffi_accum_single_step_t, iz)};

}  // namespace

GOMELT_HANDLER(GomeltLevelStepFfi, gomelt_step_args_t, kStep, true, "gomelt_level_step_f32",
               rc = gomelt_level_step_f32(&props, &a, st))
GOMELT_HANDLER(GomeltStatePropsFfi, gomelt_ffi_state_props_t, kStateProps, true, "gomelt_state_props_f32",
               rc = gomelt_state_props_f32(&props, a.T, a.S1, a.nn, a.n_substrate, a.S1_out, a.S2_out, a.k_out, a.rhocp_out, st))
GOMELT_HANDLER(GomeltSurfaceFluxFfi, gomelt_ffi_surface_flux_t, kSurfaceFlux, true, "gomelt_surface_flux_f32",
               rc = gomelt_surface_flux_f32(&props, &a.grid, a.T0, a.nz_active, a.flux, a.add, st))
// (the coefficient 6 sqrt3 P eta wq that the C entry point returns on the host is a closed form of the attributes)
GOMELT_HANDLER(GomeltSourceTablesFfi, gomelt_ffi_source_tables_t, kSourceTables, true, "gomelt_source_tables_f32",
               const float* laser = fr.f32_array("laser", 3);
               if (!laser) return fr.fail(XLA_FFI_Error_Code_INVALID_ARGUMENT, fr.err);
               float coef = 0.f;
               rc = gomelt_source_tables_f32(&props, &a.grid, a.x, a.y, a.z, laser, a.laserP, a.tx, a.ty, a.tz, &coef, st))
GOMELT_HANDLER(GomeltSourceTablesBatchFfi, gomelt_ffi_source_tables_batch_t, kSourceTablesBatch, true, "gomelt_source_tables_batch_f32",
               if (a.n < 1 || a.n > GOMELT_MAX_SUBSTEPS) return fr.fail(XLA_FFI_Error_Code_INVALID_ARGUMENT, "n out of range");
               const float* rows = fr.f32_array("rows", 7 * (size_t)a.n);
               if (!rows) return fr.fail(XLA_FFI_Error_Code_INVALID_ARGUMENT, fr.err);
               float coef[GOMELT_MAX_SUBSTEPS];
               rc = gomelt_source_tables_batch_f32(&props, &a.grid, a.x, a.y, a.z, rows, a.n, a.tables, coef, st))
GOMELT_HANDLER(GomeltInterpFfi, gomelt_interp_args_t, kInterp, false, "gomelt_interp_f32", rc = gomelt_interp_f32(&a, st))
GOMELT_HANDLER(GomeltFacesGatherFfi, gomelt_ffi_faces_gather_t, kFacesGather, false, "gomelt_faces_gather_f32",
               rc = gomelt_faces_gather_f32(&a.interp, a.face_a, a.face_b, st))
GOMELT_HANDLER(GomeltFacesBlendFfi, gomelt_ffi_faces_blend_t, kFacesBlend, false, "gomelt_faces_blend_f32",
               rc = gomelt_faces_blend_f32(a.face_a, a.face_b, a.ntx, a.nty, a.ntz, a.alpha, a.beta, a.has_clamp, a.clamp_min, a.out, st))
GOMELT_HANDLER(GomeltBoxCopyFfi, gomelt_ffi_box_copy_t, kBoxCopy, false, "gomelt_box_copy",
               rc = gomelt_box_copy(a.src, a.dst, a.elem_size, a.ix, a.iy, a.iz, a.nx, a.ny, a.nz, a.big_nx, a.big_ny, a.scatter, st))
GOMELT_HANDLER(GomeltPatchCopyFfi, gomelt_ffi_patch_copy_t, kPatchCopy, false, "gomelt_patch_copy_f32",
               rc = gomelt_patch_copy_f32(a.src, a.sdims, a.slo, a.dst, a.ddims, a.dlo, a.n, st))
GOMELT_HANDLER(GomeltRank1Ffi, gomelt_ffi_rank1_t, kRank1, false, "gomelt_rank1_f32",
               rc = gomelt_rank1_f32(a.F, a.tx, a.ty, a.tz, a.nx, a.ny, a.nz, a.coef, a.accumulate, st))
GOMELT_HANDLER(GomeltCoarseSourceTablesFfi, gomelt_ffi_coarse_source_tables_t, kCoarseTables, true, "gomelt_coarse_source_tables_f32",
               const float* laser = fr.f32_array("laser", 3);
               if (!laser) return fr.fail(XLA_FFI_Error_Code_INVALID_ARGUMENT, fr.err);
               float coef = 0.f;
               rc = gomelt_coarse_source_tables_f32(&props, a.fine, a.parent, laser, a.laserP, a.tx, a.ty, a.tz, &coef, st))
GOMELT_HANDLER(GomeltProjectedSourceFfi, gomelt_ffi_projected_source_t, kProjSource, true, "gomelt_projected_source_f32",
               if (a.n < 1 || a.n > GOMELT_MAX_SUBSTEPS) return fr.fail(XLA_FFI_Error_Code_INVALID_ARGUMENT, "n out of range");
               const float* rows = fr.f32_array("rows", 7 * (size_t)a.n);
               if (!rows) return fr.fail(XLA_FFI_Error_Code_INVALID_ARGUMENT, fr.err);
               rc = gomelt_projected_source_f32(&props, a.fine, a.parent, a.wq_fine, rows, a.n, a.tables, a.F, a.accumulate, st))
GOMELT_HANDLER(GomeltProjectFfi, gomelt_project_args_t, kProject, false, "gomelt_project_f32",
               // coef_props: non-zero in the blob = "use the props attribute"
               if (a.coef_props) {
                   if (!fr.blob("props", &props, sizeof props)) return fr.fail(XLA_FFI_Error_Code_INVALID_ARGUMENT, fr.err);
                   a.coef_props = &props;
               }
               rc = gomelt_project_f32(&a, st))
GOMELT_HANDLER(GomeltShiftWindowFfi, gomelt_shift_args_t, kShift, false, "gomelt_shift_window_f32", rc = gomelt_shift_window_f32(&a, st))
GOMELT_HANDLER(GomeltClampMinFfi, gomelt_ffi_clamp_min_t, kClamp, false, "gomelt_clamp_min_f32", rc = gomelt_clamp_min_f32(a.x, a.n, a.lo, st))
GOMELT_HANDLER(GomeltMinMaxFfi, gomelt_ffi_minmax_t, kMinMax, false, "gomelt_minmax_f32", rc = gomelt_minmax_f32(a.x, a.n, a.out3, st))
GOMELT_HANDLER(GomeltAccumSingleStepFfi, gomelt_ffi_accum_single_step_t, kAccum, false, "gomelt_accum_single_step_f32",
               rc = gomelt_accum_single_step_f32(a.T3, a.resetmask, a.dt, a.T_liquidus, a.accum0, a.max_accum0, a.ix, a.iy, a.iz, a.nx,
                                                 a.ny, a.nz, a.big_nx, a.big_ny, st))

// The Level-3 inner scan: substep i writes T_a (i even) / T_b (i odd), so the newest field is in T_a when n is odd, in T_b
// when n is even - the host out-parameter T_last of the C entry point is not needed.
extern "C" void* GomeltL3SubstepsFfi(void* call_frame) {
    XLA_FFI_CallFrame* cf = static_cast<XLA_FFI_CallFrame*>(call_frame);
    if (prologue(cf)) return nullptr;
    Frame fr{cf, {}};
    gomelt_ffi_l3_substeps_t a;
    gomelt_props_t props;
    static const size_t kFaces[] = {AXIS3(gomelt_interp_args_t, src), OFF(gomelt_interp_args_t, u), OFF(gomelt_interp_args_t, u2),
                                    OFF(gomelt_interp_args_t, tx), OFF(gomelt_interp_args_t, ty), OFF(gomelt_interp_args_t, tz)};
    if (!fr.blob("args", &a, sizeof a) || !fr.blob("props", &props, sizeof props) || !fr.patch(&a.substeps, kSubsteps, NOFF(kSubsteps)) ||
        (a.has_faces && !fr.patch(&a.faces, kFaces, NOFF(kFaces))))
        return fr.fail(XLA_FFI_Error_Code_INVALID_ARGUMENT, "gomelt_l3_substeps_f32: " + fr.err);
    if (a.substeps.n < 1 || a.substeps.n > GOMELT_MAX_SUBSTEPS) return fr.fail(XLA_FFI_Error_Code_INVALID_ARGUMENT, "n out of range");
    const float* rows = fr.f32_array("rows", 7 * (size_t)a.substeps.n);
    if (!rows) return fr.fail(XLA_FFI_Error_Code_INVALID_ARGUMENT, fr.err);
    a.substeps.rows = rows;
    a.substeps.step_events = nullptr;   // host-side measurement hook: not reachable from an XLA computation
    a.substeps.faces = a.has_faces ? &a.faces : nullptr;
    a.substeps.T_last = nullptr;
    void* st = nullptr;
    if (XLA_FFI_Error* e = fr.stream(&st)) return e;
    return fr.from_rc(gomelt_l3_substeps_f32(&props, &a.substeps, st), "gomelt_l3_substeps_f32");
}

// The steppers (stepGOMELT cF:2304, subcycleGOMELT cF:3224, stepGOMELTDwellTime cF:2617 - the functions the reference jits).
// State is updated in place: the JAX side aliases the state operands to results.  With L1_spare given the new Level-1
// field is left there (hand both Level-1 buffers in as operands / results and swap them on the JAX side).
extern "C" void* GomeltSubcycleFfi(void* call_frame) {
    XLA_FFI_CallFrame* cf = static_cast<XLA_FFI_CallFrame*>(call_frame);
    if (prologue(cf)) return nullptr;
    Frame fr{cf, {}};
    gomelt_ffi_subcycle_t a;
    gomelt_props_t props;
    static const size_t kExtra[] = {OFF(gomelt_ffi_subcycle_t, max_accum), OFF(gomelt_ffi_subcycle_t, accum)};
    if (!fr.blob("args", &a, sizeof a) || !fr.blob("props", &props, sizeof props) || !fr.patch(&a.hier, kHier, NOFF(kHier)) ||
        !fr.patch(&a, kExtra, NOFF(kExtra)))
        return fr.fail(XLA_FFI_Error_Code_INVALID_ARGUMENT, "gomelt_subcycle_f32: " + fr.err);
    if (a.N2 < 1 || a.N3 < 1 || (long long)a.N2 * a.N3 > GOMELT_MAX_SUBSTEPS) return fr.fail(XLA_FFI_Error_Code_INVALID_ARGUMENT, "N2 * N3 out of range");
    const float* rows = fr.f32_array("rows", 7 * (size_t)a.N2 * a.N3);
    if (!rows) return fr.fail(XLA_FFI_Error_Code_INVALID_ARGUMENT, fr.err);
    void* st = nullptr;
    if (XLA_FFI_Error* e = fr.stream(&st)) return e;
    int32_t in_spare = 0;
    a.hier.l1_solve = nullptr;   // (a host callback: the slab-decomposed drop-in drives the C ABI directly)
    a.hier.l1_user = nullptr;
    return fr.from_rc(gomelt_subcycle_f32(&props, &a.hier, rows, a.N2, a.N3, a.max_accum, a.accum, &in_spare, st), "gomelt_subcycle_f32");
}
extern "C" void* GomeltStepFfi(void* call_frame) {
    XLA_FFI_CallFrame* cf = static_cast<XLA_FFI_CallFrame*>(call_frame);
    if (prologue(cf)) return nullptr;
    Frame fr{cf, {}};
    gomelt_ffi_step_t a;
    gomelt_props_t props;
    static const size_t kExtra[] = {OFF(gomelt_ffi_step_t, resetmask)};
    if (!fr.blob("args", &a, sizeof a) || !fr.blob("props", &props, sizeof props) || !fr.patch(&a.hier, kHier, NOFF(kHier)) ||
        !fr.patch(&a, kExtra, NOFF(kExtra)))
        return fr.fail(XLA_FFI_Error_Code_INVALID_ARGUMENT, "gomelt_step_f32: " + fr.err);
    const float* row = fr.f32_array("rows", 7);
    if (!row) return fr.fail(XLA_FFI_Error_Code_INVALID_ARGUMENT, fr.err);
    void* st = nullptr;
    if (XLA_FFI_Error* e = fr.stream(&st)) return e;
    int32_t in_spare = 0;
    a.hier.l1_solve = nullptr;   // (a host callback: the slab-decomposed drop-in drives the C ABI directly)
    a.hier.l1_user = nullptr;
    return fr.from_rc(gomelt_step_f32(&props, &a.hier, row, a.resetmask, &in_spare, st), "gomelt_step_f32");
}
extern "C" void* GomeltDwellStepFfi(void* call_frame) {
    XLA_FFI_CallFrame* cf = static_cast<XLA_FFI_CallFrame*>(call_frame);
    if (prologue(cf)) return nullptr;
    Frame fr{cf, {}};
    gomelt_ffi_dwell_step_t a;
    gomelt_props_t props;
    if (!fr.blob("args", &a, sizeof a) || !fr.blob("props", &props, sizeof props) || !fr.patch(&a.hier, kHier, NOFF(kHier)))
        return fr.fail(XLA_FFI_Error_Code_INVALID_ARGUMENT, "gomelt_dwell_step_f32: " + fr.err);
    void* st = nullptr;
    if (XLA_FFI_Error* e = fr.stream(&st)) return e;
    int32_t in_spare = 0;
    a.hier.l1_solve = nullptr;   // (a host callback: the slab-decomposed drop-in drives the C ABI directly)
    a.hier.l1_user = nullptr;
    return fr.from_rc(gomelt_dwell_step_f32(&props, &a.hier, a.dt, &in_spare, st), "gomelt_dwell_step_f32");
}

#define GOMELT_FFI_NAME(name) #name,
static const char* const kHandlerNames[] = {GOMELT_FFI_HANDLERS(GOMELT_FFI_NAME)};
extern "C" int gomelt_xla_ffi_handler_count(void) { return (int)(sizeof(kHandlerNames) / sizeof(kHandlerNames[0])); }
extern "C" const char* gomelt_xla_ffi_handler_name(int i) {
    return (i >= 0 && i < gomelt_xla_ffi_handler_count()) ? kHandlerNames[i] : nullptr;
}
// 1: the XLA FFI handler symbols are part of this library (always, since the layer builds against the C API subset of
// third_party/xla_ffi_min when no jaxlib is installed)
extern "C" int gomelt_xla_ffi_available(void) { return 1; }
