// K1 — fused level step (sm_100a).
//
// One launch = one explicit sweep of one level:
//   computeStateProperties cF:2567-2614 + solveMatrixFreeFE cF:582-642 + "+Fc+Corr" cF:642
//   + substitute_Tbar cF:1848-1865 + assignBCs cF:1568-1595 + max(T_amb,.) + melt-time
//   bookkeeping cF:3568-3578.
//
// Formulation (DESIGN.md "K1"): on a box hex8 with 2x2x2 Gauss,
//   Ke = kbar * sum_d (V/h_d^2) D (x) M (x) M,  D = [[1,-1],[-1,1]],  M = [[1/3,1/6],[1/6,1/3]],
// which is diagonal in the +-1 (Haar) basis of the 8 corners.  The update is evaluated as
//   T_new = T + (F + Corr - sum_e kbar_e (Ke0 T_e)[a]) * (64 dt / V) / sum_e sum_8 rhocp
// i.e. the reference's (sum_e (diag(Me) - Ke) T_e + F + Corr) / sum_e Me without the
// M*T - K*T cancellation.  No (ne,8,8) or (ne,8) array is ever materialised.
//
// Mapping: k_level_step_v2.cuh (one warp per 60 x RY column patch, packed f32x2, no shared memory in the plane
// loop).  The first version (CTA tile, shared-memory y exchange, one barrier per plane, scalar math) ran 135 us
// at 10.26 M nodes against 61 us and was removed; its ncu summary is kept in profiles/r01_k1_v1_ncu_full.txt.
#include <stdarg.h>

#include <stdlib.h>

#include "common.cuh"
#include "k_level_step_v2.cuh"
#include "k_level_step_v3.cuh"

static_assert(GOMELT_HALO_SYNC_HEAD >= 2, "two arrival counters precede the strip counters");

namespace gomelt {


// Development switches (A/B of kernel variants) exist only in a -DGOMELT_DEBUG build (GOMELT_NVCC_FLAGS=-DGOMELT_DEBUG
// python gomelt_b200/build.py --force); the shipped library reads no environment variable on the step path.
#ifdef GOMELT_DEBUG
static int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return (v && *v) ? atoi(v) : dflt;
}
#define GM_DEV_SWITCH(name, dflt) ([] { static const int v = env_int(name, dflt); return v; }())
#else
#define GM_DEV_SWITCH(name, dflt) (dflt)
#endif

int sm_count() {  // SMs of the current device (148 on B200), asked once per device
    static thread_local int cached_dev = -1, cached_n = 0;
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev != cached_dev) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = GOMELT_SM_COUNT;
        cached_dev = dev;
        cached_n = n;
    }
    return cached_n;
}

template <int RY, int WPB, int FEAT, int MINB = 1, bool STAGE = false>
static void launch_v2(const StepParams& sp, int nch, cudaStream_t st) {
    dim3 block(32, WPB);
    const int strips = (sp.ny + RY - 1) / RY;
    dim3 grid((sp.nx + 2 * K1_TX - 1) / (2 * K1_TX), (strips + WPB - 1) / WPB, nch);
    level_step_v2<RY, WPB, FEAT, MINB, STAGE><<<grid, block, 0, st>>>(sp), count_launch();
}

// Exact-feature instances of the general kernel for the call shapes the steppers issue on grids too small for the
// fast kernel; anything else runs the generic instance.  The substrate override is compiled in (its test is two
// integer compares per row and n_substrate = 0 switches it off), so a shape is looked up with K1F_NSUB set.
constexpr int F_L3_SUB = K1F_SRC | K1F_FLUX | K1F_S1OUT | K1F_CLAMP | K1F_SKIP | K1F_NSUB;  // subcycleL3_Part1 cF:3367-3412
constexpr int F_L3_SUB2 = F_L3_SUB | K1F_S2OUT | K1F_ACCUM;                    // subcycleL3_Part2 cF:3530-3590
constexpr int F_L3_STEP = K1F_SRC | K1F_FLUX | K1F_CLAMP | K1F_SKIP | K1F_NSUB;  // stepGOMELT Level 3
constexpr int F_L2 = K1F_RHS | K1F_FLUX | K1F_CLAMP | K1F_SKIP | K1F_NSUB;     // Level 2 (step / subcycle)
constexpr int F_L1 = K1F_RHS | K1F_FLUX | K1F_CLAMP | K1F_BCCONST | K1F_NSUB;  // Level 1 (step / subcycle)
constexpr int F_L1_DWELL_SUB = K1F_FLUX | K1F_BCCONST | K1F_NSUB;              // stepGOMELTDwellTime
constexpr int F_L1_DWELL_SUB_PEER = F_L1_DWELL_SUB | K1F_PEER;                 // ... with in-kernel peer stores (legacy)

// Tensor maps of a field for the TMA ring of level_step_v3 (K1F_TMA): nn floats as a 1-D tensor with K3_BOX-element
// boxes, and as the overlapping 2-D view.  cuTensorMapEncodeTiled is taken from the driver through the runtime (no
// link-time dependency on libcuda).  The steppers ping-pong a handful of buffers, so the encoded maps are kept in a
// small per-thread cache keyed by (pointer, nn, nx): no driver call on the hot path after the first sweeps.
struct TmapPair {
    const void* base;
    unsigned long long nn;
    int nx, ok2;
    CUtensorMap m1, m2;
};
static bool field_tmaps(const void* base, unsigned long long nn, int nx, CUtensorMap& m1, CUtensorMap& m2, int& ok2) {
    typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static encode_fn enc = [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            f = nullptr;
        return (encode_fn)f;
    }();
    if (!enc || ((uintptr_t)base & 15u) != 0) return false;
    constexpr int NC = 16;
    static thread_local TmapPair cache[NC];
    static thread_local int next = 0;
    for (int i = 0; i < NC; ++i)
        if (cache[i].base == base && cache[i].nn == nn && cache[i].nx == nx) {
            m1 = cache[i].m1; m2 = cache[i].m2; ok2 = cache[i].ok2;
            return true;
        }
    TmapPair e;
    e.base = base; e.nn = nn; e.nx = nx; e.ok2 = 1;
    const cuuint64_t dim1[1] = {nn}, stride0[1] = {0};  // stride unused for rank 1
    const cuuint32_t box1[1] = {K3_BOX}, estr[2] = {1, 1};
    const cuuint64_t dim2[2] = {nn, 6}, stride2[1] = {(cuuint64_t)(nx & ~3) * 4u};
    const cuuint32_t box2[2] = {K3_BOX, 6};
    void* b = const_cast<void*>(base);
    if (enc(&e.m1, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 1, b, dim1, stride0, box1, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return false;
    if (enc(&e.m2, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, b, dim2, stride2, box2, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        e.ok2 = 0;  // the six 1-D boxes per field and plane serve every tile then
    cache[next] = e;
    next = (next + 1) % NC;
    m1 = e.m1; m2 = e.m2; ok2 = e.ok2;
    return true;
}
static bool make_field_tmaps(StepParams& sp) {
    const unsigned long long nn = (unsigned long long)sp.nx * sp.ny * sp.nz;
    int okT = 0, okS = 0;
    if (!field_tmaps(sp.T0, nn, sp.nx, sp.tm[0], sp.tm2[0], okT)) return false;
    if (!field_tmaps(sp.S1, nn, sp.nx, sp.tm[1], sp.tm2[1], okS)) return false;
    sp.tm2_ok = (okT && okS && !GM_DEV_SWITCH("GOMELT_K1_TMA_1D", 0)) ? 1 : 0;
    return true;
}

template <int RY, int FEAT, int MINB = 1>
static void launch_v3(const StepParams& sp, int nch, cudaStream_t st) {
    dim3 grid((sp.nx - 2 + 2 * K1_TX - 1) / (2 * K1_TX), (sp.ny - 2 + RY - 1) / RY, nch);
    if constexpr ((FEAT & K1F_TMA) != 0) {  // 8 resident warps x 20 KB: ask for the large shared-memory carve-out once
        static const cudaError_t carve = cudaFuncSetAttribute(level_step_v3<RY, FEAT, MINB>,
                                                               cudaFuncAttributePreferredSharedMemoryCarveout,
                                                               (int)cudaSharedmemCarveoutMaxShared);
        (void)carve;
    }
    level_step_v3<RY, FEAT, MINB><<<grid, 32, 0, st>>>(sp), count_launch();
}

// v3 (k_level_step_v3.cuh) serves the Dirichlet-side-face shapes of the steppers on grids that hold a full tile.
// Returns false when the call is not one of them (v2 takes it).
static bool gridDim_fits_u16(const StepParams& sp) {  // queue entries pack the tile indices into 16 bits each
    return (sp.nx - 2 + 2 * K1_TX - 1) / (2 * K1_TX) < 65536 && (sp.ny - 2 + 4 - 1) / 4 < 65536;
}

static bool try_launch_v3(StepParams& sp, int nch, cudaStream_t st) {
    constexpr int RY = 4;
    const int f = sp.feat;
    if (GM_DEV_SWITCH("GOMELT_K1_V2", 0) || (sp.flags & GOMELT_STEP_GENERAL_KERNEL) || !(f & (K1F_SKIP | K1F_BCCONST)) ||
        (f & K1F_TOP))
        return false;
    if (sp.nx < 2 * K1_TX + 2 || sp.ny < RY + 2 || sp.nzl < 2 || sp.nsub_rem != 0) return false;
    // plane 0 is never finalised: its row stores are redirected to plane 1, which the same warp must overwrite
    if (sp.zbeg == 0 && (sp.zchunk < 2 || sp.zend < 2)) return false;
    constexpr int V3_L3_SUB = K1F_SRC | K1F_FLUX | K1F_S1OUT | K1F_CLAMP | K1F_NSUB;  // subcycleL3_Part1
    constexpr int V3_L3_SUB2 = V3_L3_SUB | K1F_S2OUT | K1F_ACCUM;                       // subcycleL3_Part2 (melt-time bookkeeping)
    constexpr int V3_L3_STEP = K1F_SRC | K1F_FLUX | K1F_CLAMP | K1F_NSUB;             // stepGOMELT Level 3
    constexpr int V3_RHS = K1F_RHS | K1F_FLUX | K1F_CLAMP | K1F_NSUB;                 // Level 2 and Level 1 (step / subcycle)
    constexpr int V3_RHS_NC = K1F_RHS | K1F_FLUX | K1F_NSUB;                          // ... without the clamp (stepGOMELT parents)
    constexpr int V3_DWELL = K1F_FLUX | K1F_NSUB;                                     // stepGOMELTDwellTime (no clamp)
    // The TMA plane ring (+ cold-plane path) needs 16-byte aligned field pointers and the driver's tensor-map encoder;
    // without them the general kernel takes the call.
    if (!make_field_tmaps(sp)) return false;
    // z-slab ranks: the boundary planes go to the neighbours after the step - halo_exchange_kernel (launch_step) when
    // halo_sync is given, else halo_push_kernel here and the caller orders the sweeps (round-1 form, kept for A/B and for
    // callers that bring their own barrier).
    const bool do_push = (f & K1F_PEER) && !sp.hsync && (sp.zend - sp.zbeg) >= 1;
    const int fsw = (f & ~(K1F_SKIP | K1F_BCCONST | K1F_PEER)) | K1F_NSUB;
#define GM_V3(FEATS) launch_v3<RY, (FEATS) | K1F_TMA>(sp, nch, st)
    switch (fsw) {
        case V3_L3_SUB:
            // in place (the steppers): a node's state is stored only when it changed
            if (sp.S1out == sp.S1) GM_V3(V3_L3_SUB | K1F_S1INPLACE);
            else GM_V3(V3_L3_SUB);
            break;
        case V3_L3_SUB2: {
            // (only where the launch fills the GPU: on a small window the second launch costs more than the tail it removes)
            const long long tiles = (long long)((sp.nx - 2 + 2 * K1_TX - 1) / (2 * K1_TX)) * ((sp.ny - 2 + RY - 1) / RY) * nch;
            const bool queue = sp.bkq != nullptr && sp.bkq_cap > 0 && gridDim_fits_u16(sp) && (tiles >= 4LL * sm_count() || sp.bkq_force);
            if (!queue) sp.bkq = nullptr;
            else if (sp.bkq_reset && cudaMemsetAsync(sp.bkq, 0, 8, st) != cudaSuccess) return false;
            if (sp.S1out == sp.S1) GM_V3(V3_L3_SUB2 | K1F_S1INPLACE);
            else GM_V3(V3_L3_SUB2);
            if (queue)  // the hot planes the step queued: one warp per (tile, plane), spread over the GPU
                bookkeep_queue_kernel<RY, true, true><<<sm_count(), 128, 0, st>>>(sp), count_launch();
            break;
        }
        case V3_L3_STEP: GM_V3(V3_L3_STEP); break;
        case V3_RHS: GM_V3(V3_RHS); break;  // (the rhs rows are fetched a plane ahead into registers there)
        case V3_RHS_NC: GM_V3(V3_RHS_NC); break;
        case V3_RHS | K1F_S1OUT:  // Level 2 inside subcycleGOMELT: the state advances in place with the solve
            if (sp.S1out != sp.S1) return false;
            GM_V3(V3_RHS | K1F_S1OUT | K1F_S1INPLACE);
            break;
        case V3_DWELL: GM_V3(V3_DWELL); break;
        default: return false;
    }
#undef GM_V3
    // Level-1 Dirichlet constants (the step never stores a face node): with halo_sync the exchange kernel writes them
    // (launch_step), else face_const_kernel - which, on the round-1 z-slab path, also fills the neighbours' ghost planes
    if ((f & K1F_BCCONST) && !sp.hsync) {
        {
            const long long n = (long long)(2 * sp.nx + 2 * sp.ny) * (sp.zend - sp.zbeg) + (sp.zbeg == 0 ? (long long)sp.nx * sp.ny : 0);
            const int cap = 4 * sm_count();
            const int blocks = (int)((n + 255) / 256 < cap ? (n + 255) / 256 : cap);
            face_const_kernel<<<blocks, 256, 0, st>>>(sp.Tout, sp.nx, sp.ny, sp.nz, sp.zbeg, sp.zend, sp.bc[0], sp.bc[1], sp.bc[2],
                                                      sp.bc[3], sp.bc[4], do_push ? sp.peer_lo : nullptr,
                                                      do_push ? sp.peer_hi : nullptr), count_launch();
        }
    }
    if (do_push) {
        const int plane = sp.nx * sp.ny;
        dim3 grid((plane / 4 + 255) / 256, 2);
        halo_push_kernel<<<grid, 256, 0, st>>>(sp.Tout, plane, sp.zbeg, sp.zend - 1, sp.peer_lo, sp.peer_hi), count_launch();
    }
    return true;
}

static int launch_kernels(StepParams& sp, cudaStream_t st);

static int launch_step(StepParams& sp, cudaStream_t st) {
    if (!sp.hsync) return launch_kernels(sp, st);
    // z-slab rank under the halo protocol: the step, then the exchange of the two boundary planes of T_out
    float* const plo = sp.peer_lo;
    float* const phi = sp.peer_hi;
    sp.peer_lo = sp.peer_hi = nullptr;  // the step itself stores nothing remotely
    sp.feat &= ~K1F_PEER;
    const int rc = launch_kernels(sp, st);
    if (rc) return rc;
    // faces: by the exchange kernel when the fast kernel ran (the general kernel writes its own Dirichlet faces)
    const bool faces_here = (sp.feat & K1F_BCCONST) && sp.ran_v3;
    const int blocks = sm_count();  // per plane: two co-resident blocks of the exchange per SM
    halo_exchange_kernel<<<dim3(blocks, 2), HALO_THREADS, 0, st>>>(sp.Tout, sp.nx, sp.ny, sp.zbeg, sp.zend - 1, plo, phi, sp.bc[0],
                                                                  sp.bc[1], sp.bc[2], sp.bc[3], sp.hsync, sp.hsync_lo, sp.hsync_hi,
                                                                  (unsigned)blocks * (sp.hseq + 1u), faces_here ? sp.Tout : nullptr, sp.nz,
                                                                  sp.zbeg, sp.zend, sp.bc[4]), count_launch();
    return check_launch("gomelt_level_step_f32 (halo exchange)");
}

static int launch_kernels(StepParams& sp, cudaStream_t st) {
    const int nch = (sp.zend - sp.zbeg + sp.zchunk - 1) / sp.zchunk;
    sp.ran_v3 = 0;
    if (try_launch_v3(sp, nch, st)) {
        sp.ran_v3 = 1;
        return check_launch("gomelt_level_step_f32");
    }
    constexpr int RY = 4, WPB = 1;  // one warp per CTA: warps are independent, finest SM balance
    const int f = sp.feat;
    if (GM_DEV_SWITCH("GOMELT_K1_GENERIC", 0)) {
        launch_v2<RY, WPB, K1F_ALL | K1F_GENERIC>(sp, nch, st);
    } else {
        switch (f | K1F_NSUB) {
            case F_L3_SUB: launch_v2<RY, WPB, F_L3_SUB>(sp, nch, st); break;
            case F_L3_SUB2: launch_v2<RY, WPB, F_L3_SUB2>(sp, nch, st); break;
            case F_L3_STEP: launch_v2<RY, WPB, F_L3_STEP>(sp, nch, st); break;
            case F_L2: launch_v2<RY, WPB, F_L2>(sp, nch, st); break;
            case F_L1: launch_v2<RY, WPB, F_L1>(sp, nch, st); break;
            case F_L1_DWELL_SUB: launch_v2<RY, WPB, F_L1_DWELL_SUB>(sp, nch, st); break;
            case F_L1_DWELL_SUB_PEER: launch_v2<RY, WPB, F_L1_DWELL_SUB_PEER>(sp, nch, st); break;
            default: launch_v2<RY, WPB, K1F_ALL | K1F_GENERIC>(sp, nch, st); break;
        }
    }
    return check_launch("gomelt_level_step_f32");
}

}  // namespace gomelt

using namespace gomelt;

extern "C" long long gomelt_halo_sync_words(void) { return GOMELT_HALO_SYNC_HEAD; }

extern "C" int gomelt_level_step_f32(const gomelt_props_t* props, const gomelt_step_args_t* a, void* stream) {
    if (!props || !a || !a->T0 || !a->S1 || !a->T_out) {
        set_error("gomelt_level_step_f32: NULL props/args/T0/S1/T_out");
        return GOMELT_E_NULL;
    }
    const gomelt_grid_t& g = a->grid;
    const int zbeg = (a->z_begin == 0 && a->z_end == 0) ? 0 : a->z_begin;
    const int zend = (a->z_begin == 0 && a->z_end == 0) ? g.nz : a->z_end;
    if (g.nx < 2 || g.ny < 2 || g.nz < 2 || a->nz_active < 0 || a->nz_active > g.nz || zbeg < 0 || zbeg >= zend ||
        zend > g.nz ||
        (long long)g.nx * g.ny * g.nz > 2000000000LL || !(a->dt > 0.f)) {
        set_error("gomelt_level_step_f32: bad grid %d x %d x %d (nz_active %d, dt %g)", g.nx, g.ny, g.nz,
                  a->nz_active, (double)a->dt);
        return GOMELT_E_SIZE;
    }
    if (a->T_out == a->T0) {
        set_error("gomelt_level_step_f32: T_out must not alias T0");
        return GOMELT_E_FLAGS;
    }
    if (((a->flags & GOMELT_STEP_WRITE_S1) && !a->S1_out) || ((a->flags & GOMELT_STEP_WRITE_S2) && !a->S2_out) ||
        ((a->flags & GOMELT_STEP_ACCUM) && (!a->S2_prev || !a->accum || !a->max_accum)) ||
        ((a->flags & GOMELT_STEP_BC_CONST) && (a->flags & GOMELT_STEP_SKIP_FACES))) {
        set_error("gomelt_level_step_f32: flags 0x%x inconsistent with the pointers given", a->flags);
        return GOMELT_E_FLAGS;
    }
    const bool any_src = a->src_x || a->src_y || a->src_z;
    if (any_src && !(a->src_x && a->src_y && a->src_z)) {
        set_error("gomelt_level_step_f32: src_x/src_y/src_z must be all set or all NULL");
        return GOMELT_E_NULL;
    }
    StepParams sp;
    sp.nx = g.nx; sp.ny = g.ny; sp.nz = g.nz; sp.nzl = a->nz_active;
    sp.nsub = a->n_substrate;
    const double hx = g.hx, hy = g.hy, hz = g.hz, V = hx * hy * hz;
    const double c[3] = {V / (hx * hx), V / (hy * hy), V / (hz * hz)};
    const double muD[2] = {0.0, 2.0}, muM[2] = {0.5, 1.0 / 6.0};
    double lam_d[8];
    for (int s = 0; s < 8; ++s) {
        const int sd[3] = {s & 1, (s >> 1) & 1, (s >> 2) & 1};
        double lam = 0.0;
        for (int d = 0; d < 3; ++d) {
            double t = c[d] * muD[sd[d]];
            for (int e = 0; e < 3; ++e)
                if (e != d) t *= muM[sd[e]];
            lam += t;
        }
        sp.lam[s] = (float)(lam / 64.0);  // 1/8 (Haar inverse) * 1/8 (kbar = k8/8)
        lam_d[s] = lam / 64.0;
    }
    // v3: mode pairs (sx,sz) = (0,1), (1,0), (1,1); l_s = lambda'[sy = 0], l_d = lambda'[sy = 1]; everything
    // relative to s = lambda'[2] (see StepParams)
    const int pair_s[3] = {4, 1, 5}, pair_d[3] = {6, 3, 7};
    const double s_norm = lam_d[2];
    for (int q = 0; q < 3; ++q) {
        sp.lamA[q] = (float)((lam_d[pair_s[q]] + lam_d[pair_d[q]]) / s_norm);
        sp.lamB[q] = (float)((lam_d[pair_s[q]] - lam_d[pair_d[q]]) / s_norm);
    }
    sp.cdt = (float)(64.0 * (double)a->dt / V);
    sp.dt = a->dt;
    sp.pk = fold_props(*props);
    sp.fk = fold_flux(*props, g);
    {
        const double mscale = 1.0 / ((64.0 * (double)a->dt / V) * s_norm);
        sp.n_ca0 = (float)((double)props->rho * (double)props->cp_solid_a0 * mscale);
        sp.n_ca1 = (float)((double)props->rho * (double)props->cp_solid_a1 * mscale);
        sp.n_cmushy = (float)((double)props->rho * (double)props->cp_mushy * mscale);
        sp.n_cfluid = (float)((double)props->rho * (double)props->cp_fluid * mscale);
        sp.n_inv_s = (float)(1.0 / s_norm);
        sp.n_wq = (float)((double)sp.fk.wq / s_norm);
    }
    sp.T0 = a->T0; sp.S1 = a->S1; sp.rhs = a->rhs;
    sp.srcx = a->src_x; sp.srcy = a->src_y; sp.srcz = a->src_z; sp.scoef = a->src_coef;
    sp.topflux = a->topflux;
    sp.Tout = a->T_out; sp.S1out = a->S1_out; sp.S2out = a->S2_out;
    sp.S2prev = a->S2_prev; sp.accum = a->accum; sp.maxacc = a->max_accum;
    for (int q = 0; q < 5; ++q) sp.bc[q] = a->bc5[q];
    sp.flags = a->flags;
    sp.zbeg = zbeg; sp.zend = zend;
    sp.peer_lo = a->peer_lo; sp.peer_hi = a->peer_hi;
    sp.exp = GM_DEV_SWITCH("GOMELT_K1_EXP", 0);
    sp.hsync = a->halo_sync; sp.hsync_lo = a->halo_sync_lo; sp.hsync_hi = a->halo_sync_hi;
    sp.bkq = a->bk_queue;
    sp.bkq_cap = (a->bk_queue && a->bk_queue_words > 2) ? (unsigned)((a->bk_queue_words - 2) / 2) : 0u;
    sp.bkq_reset = (a->bk_queue_keep & 1) ? 0 : 1;
    sp.bkq_force = (a->bk_queue_keep & 2) ? 1 : 0;
    if (sp.hsync) {
        if ((a->peer_lo != nullptr) != (a->halo_sync_lo != nullptr) || (a->peer_hi != nullptr) != (a->halo_sync_hi != nullptr) ||
            !(a->flags & GOMELT_STEP_BC_CONST)) {
            set_error("gomelt_level_step_f32: halo_sync needs GOMELT_STEP_BC_CONST and a peer plane + a peer counter block per neighbour");
            return GOMELT_E_FLAGS;
        }
        sp.hseq = a->halo_seq;
    } else {
        sp.hseq = 0;
    }
    {
        const long long Pn = (long long)g.nx * g.ny;
        const long long ns = a->n_substrate < 0 ? 0 : a->n_substrate;
        sp.nsub_planes = (int)(ns / Pn < g.nz ? ns / Pn : g.nz);
        sp.nsub_rem = sp.nsub_planes < g.nz ? (int)(ns - (long long)sp.nsub_planes * Pn) : 0;
        int f = 0;
        if (a->rhs) f |= K1F_RHS;
        if (any_src) f |= K1F_SRC;
        if (a->topflux) f |= K1F_TOP;
        if (a->flags & GOMELT_STEP_FUSED_FLUX) f |= K1F_FLUX;
        if (a->flags & GOMELT_STEP_WRITE_S1) f |= K1F_S1OUT;
        if (a->flags & GOMELT_STEP_WRITE_S2) f |= K1F_S2OUT;
        if (a->flags & GOMELT_STEP_ACCUM) f |= K1F_ACCUM;
        if (a->flags & GOMELT_STEP_BC_CONST) f |= K1F_BCCONST;
        if (a->flags & GOMELT_STEP_SKIP_FACES) f |= K1F_SKIP;
        if (a->flags & GOMELT_STEP_CLAMP) f |= K1F_CLAMP;
        if (ns > 0) f |= K1F_NSUB;
        if (a->peer_lo || a->peer_hi) f |= K1F_PEER;
        sp.feat = f;
    }
    sp.zchunk = a->z_chunk > 0 ? a->z_chunk : (zend - zbeg);
    if (a->z_chunk <= 0) {
        // Small grids (the example's levels: 6 .. 50 tiles) are a handful of warps that each march the whole column, one
        // dependent plane after the other: cut the march into z-chunks (bit-identical results; a chunk re-reads one plane)
        // until there are about two warps per SM, at least 4 planes per chunk.
        const long long tiles = (long long)((g.nx - 2 + 2 * K1_TX - 1) / (2 * K1_TX)) * ((g.ny - 2 + 3) / 4);
        const int planes = zend - zbeg;
        const long long want = 2LL * sm_count();
        if (tiles > 0 && tiles < want && planes >= 8) {
            long long nch = (want + tiles - 1) / tiles;
            if (nch > planes / 4) nch = planes / 4;
            if (nch > 1) sp.zchunk = (int)((planes + nch - 1) / nch);
        }
    }
    if (any_src && sp.zchunk > K1_SRCZ_MAX) sp.zchunk = K1_SRCZ_MAX;  // the chunk's z-factors live in shared memory
    if ((sp.feat & (K1F_S2OUT | K1F_ACCUM)) && sp.zchunk > 62) sp.zchunk = 62;  // v3 keeps a 64-bit mask of hot planes per chunk
    return launch_step(sp, (cudaStream_t)stream);
}
