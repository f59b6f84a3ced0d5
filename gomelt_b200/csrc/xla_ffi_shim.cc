// XLA FFI handlers over the C ABI (include/gomelt_abi.h) - the "thin C-ABI XLA FFI custom-call layer"
// of BASELINE.json's north_star.  COMPILE-GUARDED: jaxlib (which ships xla/ffi/api/ffi.h) is not in
// this image nor on the GPU box, so build.py compiles this file only when the header is found
// (GOMELT_XLA_INCLUDE=<jaxlib>/include, or `python -c "import jaxlib"` succeeding).  Everything that
// is testable here goes through the same extern "C" symbols by ctypes (gomelt_b200/_lib.py).
//
// Registration on the JAX side:
//     jax.ffi.register_ffi_target("gomelt_level_step_f32", jax.ffi.pycapsule(lib.GomeltLevelStepFfi),
//                                 platform="CUDA")
//     jax.ffi.ffi_call("gomelt_level_step_f32", (ShapeDtypeStruct(T0.shape, f32), ...))(T0, S1, ...)
// The typed-FFI names below follow the header-only C++ API of xla/ffi/api/ffi.h as documented by
// upstream JAX ("Foreign function interface" guide); they could not be compiled offline.
#if defined(__has_include)
#if __has_include("xla/ffi/api/ffi.h")
#define GOMELT_HAVE_XLA_FFI 1
#endif
#endif

#ifdef GOMELT_HAVE_XLA_FFI
#include <cuda_runtime_api.h>

#include "gomelt_abi.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

namespace {

ffi::Error to_error(int rc, const char* what) {
    if (rc == 0) return ffi::Error::Success();
    return ffi::Error(rc < 0 ? ffi::ErrorCode::kInvalidArgument : ffi::ErrorCode::kInternal,
                      std::string(what) + ": " + gomelt_last_error());
}

// One explicit sweep of one level (K1).  Operands: T0, S1, rhs (may be size 0), src_x/y/z (may be
// size 0), topflux (may be size 0).  Results: T_out, S1_out.  Attributes: the POD fields of
// gomelt_props_t / gomelt_grid_t / gomelt_step_args_t passed as one byte blob each, so that the
// binding stays a plain forwarding layer and the ABI structs remain the single source of truth.
ffi::Error LevelStepImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> T0, ffi::Buffer<ffi::F32> S1,
                         ffi::Buffer<ffi::F32> rhs, ffi::Buffer<ffi::F32> src_x, ffi::Buffer<ffi::F32> src_y,
                         ffi::Buffer<ffi::F32> src_z, ffi::Buffer<ffi::F32> topflux,
                         ffi::ResultBuffer<ffi::F32> T_out, ffi::ResultBuffer<ffi::F32> S1_out,
                         ffi::Span<const uint8_t> props_blob, ffi::Span<const uint8_t> args_blob) {
    if (props_blob.size() != sizeof(gomelt_props_t) || args_blob.size() != sizeof(gomelt_step_args_t))
        return ffi::Error(ffi::ErrorCode::kInvalidArgument, "gomelt: props/args blob size mismatch (ABI version?)");
    gomelt_props_t props;
    gomelt_step_args_t a;
    memcpy(&props, props_blob.begin(), sizeof(props));
    memcpy(&a, args_blob.begin(), sizeof(a));
    a.T0 = T0.typed_data();
    a.S1 = S1.typed_data();
    a.rhs = rhs.element_count() ? rhs.typed_data() : nullptr;
    const bool has_src = src_x.element_count() != 0;
    a.src_x = has_src ? src_x.typed_data() : nullptr;
    a.src_y = has_src ? src_y.typed_data() : nullptr;
    a.src_z = has_src ? src_z.typed_data() : nullptr;
    a.topflux = topflux.element_count() ? topflux.typed_data() : nullptr;
    a.T_out = T_out->typed_data();
    a.S1_out = (a.flags & GOMELT_STEP_WRITE_S1) ? S1_out->typed_data() : nullptr;
    return to_error(gomelt_level_step_f32(&props, &a, stream), "gomelt_level_step_f32");
}

}  // namespace

XLA_FFI_DEFINE_HANDLER_SYMBOL(GomeltLevelStepFfi, LevelStepImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()   // T0
                                  .Arg<ffi::Buffer<ffi::F32>>()   // S1
                                  .Arg<ffi::Buffer<ffi::F32>>()   // rhs
                                  .Arg<ffi::Buffer<ffi::F32>>()   // src_x
                                  .Arg<ffi::Buffer<ffi::F32>>()   // src_y
                                  .Arg<ffi::Buffer<ffi::F32>>()   // src_z
                                  .Arg<ffi::Buffer<ffi::F32>>()   // topflux
                                  .Ret<ffi::Buffer<ffi::F32>>()   // T_out
                                  .Ret<ffi::Buffer<ffi::F32>>()   // S1_out
                                  .Attr<ffi::Span<const uint8_t>>("props")
                                  .Attr<ffi::Span<const uint8_t>>("args"));
#else
// No jaxlib headers: nothing to compile.  gomelt_xla_ffi_available() lets the Python side say so.
#endif

extern "C" int gomelt_xla_ffi_available(void) {
#ifdef GOMELT_HAVE_XLA_FFI
    return 1;
#else
    return 0;
#endif
}
