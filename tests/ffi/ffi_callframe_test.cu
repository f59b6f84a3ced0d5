// Builds XLA_FFI_CallFrames by hand - the way the XLA runtime does - and calls the handlers of libgomelt_sm100.so:
//   1. registration-time metadata query, non-EXECUTE stages, a malformed frame (error object through the API table);
//   2. (GPU) GomeltLevelStepFfi, GomeltInterpFfi, GomeltPatchCopyFfi and GomeltDwellStepFfi against the direct C-ABI calls on the same
//      inputs, bit for bit, on the stream the fake runtime hands out.
// usage: ffi_callframe_test [--no-gpu]      exit code 0 = all checks passed
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "gomelt_xla_ffi.h"
#include "xla/ffi/api/c_api.h"

static int g_fail = 0;
#define CHECK(cond, ...)                         \
    do {                                         \
        if (!(cond)) {                           \
            printf("FAIL %s:%d: ", __FILE__, __LINE__); \
            printf(__VA_ARGS__);                 \
            printf("\n");                        \
            ++g_fail;                            \
        }                                        \
    } while (0)

// ---- a fake runtime: the two API entries the handlers use -----------------------------------------------------
struct XLA_FFI_Error {
    std::string msg;
    XLA_FFI_Error_Code code;
};
struct XLA_FFI_ExecutionContext {
    cudaStream_t stream;
};
static XLA_FFI_Error* api_error_create(XLA_FFI_Error_Create_Args* a) { return new XLA_FFI_Error{a->message, a->errc}; }
static XLA_FFI_Error* api_stream_get(XLA_FFI_Stream_Get_Args* a) {
    a->stream = a->ctx->stream;
    return nullptr;
}
static XLA_FFI_Api make_api() {
    XLA_FFI_Api api;
    memset(&api, 0, sizeof api);
    api.struct_size = sizeof api;
    api.api_version.major_version = XLA_FFI_API_MAJOR;
    api.api_version.minor_version = XLA_FFI_API_MINOR;
    api.XLA_FFI_Error_Create = api_error_create;
    api.XLA_FFI_Stream_Get = api_stream_get;
    return api;
}

// ---- call-frame builder ----------------------------------------------------------------------------------------
struct FrameBuilder {
    std::vector<XLA_FFI_Buffer> arg_bufs, ret_bufs;
    std::vector<std::vector<int64_t>> dims;
    std::vector<void*> args, rets, attrs;
    std::vector<XLA_FFI_ArgType> arg_types;
    std::vector<XLA_FFI_RetType> ret_types;
    std::vector<XLA_FFI_AttrType> attr_types;
    std::vector<XLA_FFI_ByteSpan> names;
    std::vector<XLA_FFI_ByteSpan*> name_ptrs;
    std::vector<XLA_FFI_Array> arrays;
    std::vector<std::string> name_store;
    XLA_FFI_CallFrame frame;

    FrameBuilder() {
        arg_bufs.reserve(64); ret_bufs.reserve(64); dims.reserve(128); arrays.reserve(16); names.reserve(16); name_store.reserve(16);
    }
    XLA_FFI_Buffer make(void* p, XLA_FFI_DataType t, int64_t n) {
        dims.push_back({n});
        XLA_FFI_Buffer b;
        b.struct_size = sizeof b; b.extension_start = nullptr; b.dtype = t; b.data = p; b.dims = dims.back().data(); b.rank = 1;
        return b;
    }
    int arg(void* p, XLA_FFI_DataType t, int64_t n) { arg_bufs.push_back(make(p, t, n)); return (int)arg_bufs.size(); }      // slot code
    int ret(void* p, XLA_FFI_DataType t, int64_t n) { ret_bufs.push_back(make(p, t, n)); return -(int)ret_bufs.size(); }    // slot code
    void attr_array(const char* name, XLA_FFI_DataType t, const void* data, size_t count) {  // (names must be added in sorted order)
        name_store.push_back(name);
        names.push_back({name_store.back().c_str(), name_store.back().size()});
        arrays.push_back({t, count, const_cast<void*>(data)});
        attr_types.push_back(XLA_FFI_AttrType_ARRAY);
    }
    XLA_FFI_CallFrame* build(const XLA_FFI_Api* api, XLA_FFI_ExecutionContext* ctx, XLA_FFI_ExecutionStage stage) {
        args.clear(); rets.clear(); attrs.clear(); arg_types.clear(); ret_types.clear(); name_ptrs.clear();
        for (auto& b : arg_bufs) { args.push_back(&b); arg_types.push_back(XLA_FFI_ArgType_BUFFER); }
        for (auto& b : ret_bufs) { rets.push_back(&b); ret_types.push_back(XLA_FFI_RetType_BUFFER); }
        for (size_t i = 0; i < arrays.size(); ++i) { attrs.push_back(&arrays[i]); name_ptrs.push_back(&names[i]); }
        memset(&frame, 0, sizeof frame);
        frame.struct_size = sizeof frame;
        frame.api = api; frame.ctx = ctx; frame.stage = stage;
        frame.args = {sizeof(XLA_FFI_Args), nullptr, (int64_t)args.size(), arg_types.data(), args.data()};
        frame.rets = {sizeof(XLA_FFI_Rets), nullptr, (int64_t)rets.size(), ret_types.data(), rets.data()};
        frame.attrs = {sizeof(XLA_FFI_Attrs), nullptr, (int64_t)attrs.size(), attr_types.data(), name_ptrs.data(), attrs.data()};
        return &frame;
    }
};
template <typename T>
static void set_slot(T*& field, int code) {
    int64_t c = code;
    memcpy(&field, &c, sizeof c);
}
template <typename T>
static void set_slot(const T*& field, int code) {
    int64_t c = code;
    memcpy(&field, &c, sizeof c);
}

static gomelt_props_t example_props() {
    gomelt_props_t p;
    memset(&p, 0, sizeof p);
    p.k_powder = 0.4f; p.k_bulk_a0 = 4.23f; p.k_bulk_a1 = 0.016f; p.k_fluid = 29.f;
    p.cp_solid_a0 = 383.1f; p.cp_solid_a1 = 0.174f; p.cp_mushy = 3235.f; p.cp_fluid = 769.f; p.rho = 8e-6f;
    p.T_amb = 298.15f; p.T_solidus = 1533.f; p.T_liquidus = 1609.f; p.T_boiling = 3038.f;
    p.h_conv = 15.f; p.sigma_sb = 5.67e-8f; p.vareps = 0.3f; p.evc = 0.82f; p.Lev = 6457000.f;
    p.CM_coeff = 1.1f; p.CT_coeff = 45000.f; p.CP_coeff = 54.f;
    p.laser_radius = 0.1f; p.laser_depth = 0.1f; p.laser_eta = 0.45f;
    return p;
}

static void host_only_checks(const XLA_FFI_Api& api) {
    // 1. metadata query: every handler reports the API version and does nothing else
    for (int i = 0; i < gomelt_xla_ffi_handler_count(); ++i) CHECK(gomelt_xla_ffi_handler_name(i) != nullptr, "handler name %d", i);
    CHECK(gomelt_xla_ffi_handler_count() == 22, "handler count %d", gomelt_xla_ffi_handler_count());
    XLA_FFI_Metadata md;
    memset(&md, 0xff, sizeof md);
    XLA_FFI_Metadata_Extension ext;
    ext.extension_base = {sizeof ext, XLA_FFI_Extension_Metadata, nullptr};
    ext.metadata = &md;
    FrameBuilder fb;
    XLA_FFI_CallFrame* f = fb.build(&api, nullptr, XLA_FFI_ExecutionStage_EXECUTE);
    f->extension_start = &ext.extension_base;
    CHECK(GomeltLevelStepFfi(f) == nullptr && GomeltSubcycleFfi(f) == nullptr, "metadata call returned an error");
    CHECK(md.api_version.major_version == XLA_FFI_API_MAJOR && md.api_version.minor_version == XLA_FFI_API_MINOR, "metadata version");
    // 2. stages other than EXECUTE are no-ops even on an empty frame
    f = fb.build(&api, nullptr, XLA_FFI_ExecutionStage_PREPARE);
    CHECK(GomeltInterpFfi(f) == nullptr && GomeltProjectFfi(f) == nullptr, "PREPARE stage must be a no-op");
    // 3. EXECUTE without attributes: an INVALID_ARGUMENT error object made by the runtime's own factory
    f = fb.build(&api, nullptr, XLA_FFI_ExecutionStage_EXECUTE);
    XLA_FFI_Error* e = static_cast<XLA_FFI_Error*>(GomeltLevelStepFfi(f));
    CHECK(e && e->code == XLA_FFI_Error_Code_INVALID_ARGUMENT && e->msg.find("args") != std::string::npos, "missing-attribute error");
    delete e;
    // 4. a blob of the wrong size (ABI mismatch) is refused
    char junk[8] = {0};
    FrameBuilder fb2;
    fb2.attr_array("args", XLA_FFI_DataType_U8, junk, sizeof junk);
    e = static_cast<XLA_FFI_Error*>(GomeltInterpFfi(fb2.build(&api, nullptr, XLA_FFI_ExecutionStage_EXECUTE)));
    CHECK(e && e->code == XLA_FFI_Error_Code_INVALID_ARGUMENT, "wrong blob size must be refused");
    delete e;
}

static std::vector<float> d2h(const float* d, size_t n) {
    std::vector<float> h(n);
    cudaMemcpy(h.data(), d, n * 4, cudaMemcpyDeviceToHost);
    return h;
}
static float* h2d(const std::vector<float>& h) {
    float* d;
    cudaMalloc(&d, h.size() * 4);
    cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    return d;
}

static void gpu_checks(const XLA_FFI_Api& api) {
    cudaStream_t st;
    cudaStreamCreate(&st);
    XLA_FFI_ExecutionContext ctx{st};
    const gomelt_props_t props = example_props();
    // ---- level step: a Level-1 dwell sweep (fast kernel shape) -----------------------------------------------
    const int nx = 70, ny = 23, nz = 9;
    const size_t nn = (size_t)nx * ny * nz;
    std::vector<float> T0(nn), S1(nn);
    srand(7);
    for (size_t i = 0; i < nn; ++i) { T0[i] = 400.f + 900.f * (rand() / (float)RAND_MAX); S1[i] = (rand() & 1) ? 1.f : 0.f; }
    float *dT0 = h2d(T0), *dS1 = h2d(S1), *dA, *dB;
    cudaMalloc(&dA, nn * 4); cudaMalloc(&dB, nn * 4);
    cudaMemset(dA, 0, nn * 4); cudaMemset(dB, 0, nn * 4);
    gomelt_step_args_t a;
    memset(&a, 0, sizeof a);
    a.grid = {nx, ny, nz, 0.2f, 0.2f, 0.2f};
    a.dt = 2e-3f; a.nz_active = nz - 1; a.n_substrate = 2 * nx * ny;
    a.flags = GOMELT_STEP_BC_CONST | GOMELT_STEP_FUSED_FLUX;
    for (int q = 0; q < 5; ++q) a.bc5[q] = 300.f + q;
    gomelt_step_args_t direct = a;
    direct.T0 = dT0; direct.S1 = dS1; direct.T_out = dA;
    CHECK(gomelt_level_step_f32(&props, &direct, st) == 0, "direct level step: %s", gomelt_last_error());
    {
        FrameBuilder fb;
        gomelt_step_args_t blob = a;
        set_slot(blob.T0, fb.arg(dT0, XLA_FFI_DataType_F32, nn));
        set_slot(blob.S1, fb.arg(dS1, XLA_FFI_DataType_F32, nn));
        set_slot(blob.T_out, fb.ret(dB, XLA_FFI_DataType_F32, nn));
        fb.attr_array("args", XLA_FFI_DataType_U8, &blob, sizeof blob);
        fb.attr_array("props", XLA_FFI_DataType_U8, &props, sizeof props);
        XLA_FFI_Error* e = static_cast<XLA_FFI_Error*>(GomeltLevelStepFfi(fb.build(&api, &ctx, XLA_FFI_ExecutionStage_EXECUTE)));
        CHECK(e == nullptr, "GomeltLevelStepFfi: %s", e ? e->msg.c_str() : "");
    }
    cudaStreamSynchronize(st);
    {
        auto x = d2h(dA, nn), y = d2h(dB, nn);
        double moved = 0;
        for (size_t i = 0; i < nn; ++i) moved += fabs(x[i] - T0[i]);
        CHECK(memcmp(x.data(), y.data(), nn * 4) == 0 && moved > 1.0, "level step through the FFI differs from the direct call");
    }
    // ---- interpolation with a time blend and a clamp -----------------------------------------------------------
    const int mx = 31, my = 17, mz = 11;
    std::vector<float> cx(nx), cy(ny), cz(nz), tx(mx), ty(my), tz(mz);
    for (int i = 0; i < nx; ++i) cx[i] = 0.2f * i;
    for (int i = 0; i < ny; ++i) cy[i] = 0.2f * i;
    for (int i = 0; i < nz; ++i) cz[i] = 0.2f * i - 1.6f;
    for (int i = 0; i < mx; ++i) tx[i] = 1.0f + 0.04f * i;
    for (int i = 0; i < my; ++i) ty[i] = 0.6f + 0.04f * i;
    for (int i = 0; i < mz; ++i) tz[i] = -0.4f + 0.04f * i;
    float *dcx = h2d(cx), *dcy = h2d(cy), *dcz = h2d(cz), *dtx = h2d(tx), *dty = h2d(ty), *dtz = h2d(tz), *dO1, *dO2;
    const size_t mn = (size_t)mx * my * mz;
    cudaMalloc(&dO1, mn * 4); cudaMalloc(&dO2, mn * 4);
    gomelt_interp_args_t ia;
    memset(&ia, 0, sizeof ia);
    ia.src[0].n = nx; ia.src[1].n = ny; ia.src[2].n = nz;
    ia.alpha = 0.6f; ia.beta = 0.4f; ia.ntx = mx; ia.nty = my; ia.ntz = mz; ia.mode = GOMELT_INTERP_SET; ia.has_clamp = 1; ia.clamp_min = 500.f;
    gomelt_interp_args_t id = ia;
    id.src[0].coords = dcx; id.src[1].coords = dcy; id.src[2].coords = dcz; id.u = dT0; id.u2 = dA; id.tx = dtx; id.ty = dty; id.tz = dtz; id.out = dO1;
    CHECK(gomelt_interp_f32(&id, st) == 0, "direct interp: %s", gomelt_last_error());
    {
        FrameBuilder fb;
        gomelt_interp_args_t blob = ia;
        set_slot(blob.src[0].coords, fb.arg(dcx, XLA_FFI_DataType_F32, nx));
        set_slot(blob.src[1].coords, fb.arg(dcy, XLA_FFI_DataType_F32, ny));
        set_slot(blob.src[2].coords, fb.arg(dcz, XLA_FFI_DataType_F32, nz));
        set_slot(blob.u, fb.arg(dT0, XLA_FFI_DataType_F32, nn));
        set_slot(blob.u2, fb.arg(dA, XLA_FFI_DataType_F32, nn));
        set_slot(blob.tx, fb.arg(dtx, XLA_FFI_DataType_F32, mx));
        set_slot(blob.ty, fb.arg(dty, XLA_FFI_DataType_F32, my));
        set_slot(blob.tz, fb.arg(dtz, XLA_FFI_DataType_F32, mz));
        set_slot(blob.out, fb.ret(dO2, XLA_FFI_DataType_F32, mn));
        fb.attr_array("args", XLA_FFI_DataType_U8, &blob, sizeof blob);
        XLA_FFI_Error* e = static_cast<XLA_FFI_Error*>(GomeltInterpFfi(fb.build(&api, &ctx, XLA_FFI_ExecutionStage_EXECUTE)));
        CHECK(e == nullptr, "GomeltInterpFfi: %s", e ? e->msg.c_str() : "");
    }
    cudaStreamSynchronize(st);
    {
        auto x = d2h(dO1, mn), y = d2h(dO2, mn);
        CHECK(memcmp(x.data(), y.data(), mn * 4) == 0 && x[mn / 2] >= 500.f, "interp through the FFI differs from the direct call");
    }
    // ---- gomelt_patch_copy_f32 (box transfers of the slab-decomposed drop-in) through its handler ----
    {
        const int32_t sd[3] = {nx, ny, nz}, lo[3] = {2, 1, 1}, n[3] = {nx - 5, ny - 3, nz - 2};
        const size_t bn = (size_t)n[0] * n[1] * n[2];
        float *dP1, *dP2;
        cudaMalloc(&dP1, bn * 4); cudaMalloc(&dP2, bn * 4);
        const int32_t zero[3] = {0, 0, 0};
        CHECK(gomelt_patch_copy_f32(dT0, sd, lo, dP1, n, zero, n, st) == 0, "direct patch copy: %s", gomelt_last_error());
        FrameBuilder fb;
        gomelt_ffi_patch_copy_t blob;
        memset(&blob, 0, sizeof blob);
        for (int d = 0; d < 3; ++d) { blob.sdims[d] = sd[d]; blob.slo[d] = lo[d]; blob.ddims[d] = n[d]; blob.n[d] = n[d]; }
        set_slot(blob.src, fb.arg(dT0, XLA_FFI_DataType_F32, nn));
        set_slot(blob.dst, fb.ret(dP2, XLA_FFI_DataType_F32, bn));
        fb.attr_array("args", XLA_FFI_DataType_U8, &blob, sizeof blob);
        XLA_FFI_Error* e = static_cast<XLA_FFI_Error*>(GomeltPatchCopyFfi(fb.build(&api, &ctx, XLA_FFI_ExecutionStage_EXECUTE)));
        CHECK(e == nullptr, "GomeltPatchCopyFfi: %s", e ? e->msg.c_str() : "");
        cudaStreamSynchronize(st);
        auto x = d2h(dP1, bn), y = d2h(dP2, bn);
        CHECK(memcmp(x.data(), y.data(), bn * 4) == 0 && x[bn / 2] > 100.f, "patch copy through the FFI differs from the direct call");
        cudaFree(dP1); cudaFree(dP2);
    }
    // ---- stepGOMELTDwellTime as one FFI call (the hierarchy struct, Level 1 only), result left in L1_spare ----
    gomelt_hier_t h;
    memset(&h, 0, sizeof h);
    h.L1.grid = a.grid; h.L1.n_substrate = a.n_substrate; h.nz_active_L1 = a.nz_active;
    for (int q = 0; q < 5; ++q) h.bc5[q] = a.bc5[q];
    float *dSp, *dWork;
    cudaMalloc(&dSp, nn * 4); cudaMalloc(&dWork, 4096);
    cudaMemset(dSp, 0, nn * 4);
    h.work_floats = 1024;
    {
        FrameBuilder fb;
        gomelt_ffi_dwell_step_t blob;
        memset(&blob, 0, sizeof blob);
        blob.hier = h; blob.dt = a.dt;
        set_slot(blob.hier.L1.x, fb.arg(dcx, XLA_FFI_DataType_F32, nx));
        set_slot(blob.hier.L1.y, fb.arg(dcy, XLA_FFI_DataType_F32, ny));
        set_slot(blob.hier.L1.z, fb.arg(dcz, XLA_FFI_DataType_F32, nz));
        set_slot(blob.hier.L1.T0, fb.arg(dT0, XLA_FFI_DataType_F32, nn));
        set_slot(blob.hier.L1.S1, fb.arg(dS1, XLA_FFI_DataType_F32, nn));
        set_slot(blob.hier.work, fb.arg(dWork, XLA_FFI_DataType_F32, 1024));
        set_slot(blob.hier.L1_spare, fb.ret(dSp, XLA_FFI_DataType_F32, nn));
        fb.attr_array("args", XLA_FFI_DataType_U8, &blob, sizeof blob);
        fb.attr_array("props", XLA_FFI_DataType_U8, &props, sizeof props);
        XLA_FFI_Error* e = static_cast<XLA_FFI_Error*>(GomeltDwellStepFfi(fb.build(&api, &ctx, XLA_FFI_ExecutionStage_EXECUTE)));
        CHECK(e == nullptr, "GomeltDwellStepFfi: %s", e ? e->msg.c_str() : "");
    }
    cudaStreamSynchronize(st);
    {
        auto x = d2h(dA, nn), y = d2h(dSp, nn);
        CHECK(memcmp(x.data(), y.data(), nn * 4) == 0, "dwell step through the FFI differs from the level step it is");
    }
    CHECK(cudaGetLastError() == cudaSuccess, "CUDA error at the end");
}

int main(int argc, char** argv) {
    const bool no_gpu = argc > 1 && strcmp(argv[1], "--no-gpu") == 0;
    CHECK(gomelt_xla_ffi_available() == 1, "gomelt_xla_ffi_available");
    const XLA_FFI_Api api = make_api();
    host_only_checks(api);
    if (!no_gpu) gpu_checks(api);
    printf(g_fail ? "ffi_callframe_test: %d check(s) FAILED\n" : "ffi_callframe_test: ok%.0d\n", g_fail);
    return g_fail ? 1 : 0;
}
