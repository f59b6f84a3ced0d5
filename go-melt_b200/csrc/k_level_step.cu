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
// Mapping: a CTA owns a 30 x (BY-2) patch of node columns (+1 halo) and marches in z over a
// z-chunk.  Thread (tx,ty) owns node column (i,j) and the element column whose low corner it
// is.  Per plane: node state -> x pair sums/differences by warp shuffle -> y stage through
// shared memory -> z stage in registers (forward Haar); scale by lambda*k8; backward Haar
// z (registers) -> y (shared memory) -> x (shuffle); one coalesced store.  One barrier per
// plane (forward and backward exchanges are double-buffered and skewed by one plane).
#include <stdarg.h>

#include <stdlib.h>

#include "common.cuh"
#include "k_level_step_v2.cuh"

namespace gomelt {


template <int BY>
__global__ void __launch_bounds__(32 * BY) level_step_kernel(const __grid_constant__ StepParams p) {
    __shared__ float4 s_fwd[2][BY][32];
    __shared__ float4 s_bwd[2][BY][32];

    const int tx = threadIdx.x, ty = threadIdx.y;
    const int nx = p.nx, ny = p.ny, nz = p.nz, nzl = p.nzl;
    const int i = blockIdx.x * 30 + tx - 1;
    const int j = blockIdx.y * (BY - 2) + ty - 1;
    const bool in_dom = (i >= 0) && (i < nx) && (j >= 0) && (j < ny);
    const bool elem_xy = (i >= 0) && (j >= 0) && (i + 1 < nx) && (j + 1 < ny) && (tx < 31) && (ty < BY - 1);
    const bool out_xy = in_dom && (tx >= 1) && (tx <= 30) && (ty >= 1) && (ty <= BY - 2);
    const int za = p.zbeg + blockIdx.z * p.zchunk;
    const int zb = min(p.zend, za + p.zchunk);  // this CTA finalises node planes [za, zb)
    const int l0 = max(za - 1, 0);
    const int lload_max = min(min(zb, nz - 1), nzl - 1);  // last plane that carries data
    const long long P = (long long)nx * ny;
    const long long col = in_dom ? ((long long)j * nx + i) : 0;
    const int flags = p.flags;
    const bool on_face_xy = (i == 0) || (i == nx - 1) || (j == 0) || (j == ny - 1);

    const float sfx = (p.srcx && in_dom) ? p.srcx[i] * p.scoef : 0.f;
    const float sfy = (p.srcy && in_dom) ? p.srcy[j] : 0.f;

    // pipeline registers
    float F00 = 0.f, F01 = 0.f, F10 = 0.f, F11 = 0.f, kfp = 0.f, mfp = 0.f;  // face(l-1)
    float tk00 = 0.f, tk01 = 0.f, tk10 = 0.f, tk11 = 0.f, m8p = 0.f;        // top of layer l-2
    float low0 = 0.f, low1 = 0.f, high0 = 0.f, high1 = 0.f, mz = 0.f;         // plane l-2 (post E of iter l-1)
    float Tm1 = 0.f, Tm2 = 0.f;

    // prefetch plane l0
    float Tld = 0.f, Sld = 0.f;
    if (in_dom && l0 <= lload_max) {
        Tld = __ldg(p.T0 + l0 * P + col);
        Sld = __ldg(p.S1 + l0 * P + col);
    }

    for (int l = l0; l <= zb + 1; ++l) {
        const int buf = l & 1;
        // ---- A: node state of plane l -------------------------------------------------
        float Tn = 0.f, kn = 0.f, mn = 0.f;
        const bool nodal = in_dom && (l <= lload_max);
        if (nodal) {
            const long long n = l * P + col;
            bool s1, s2;
            Tn = Tld;
            node_props(p.pk, Tn, Sld, n < p.nsub, kn, mn, s1, s2);
            if (out_xy && l >= za && l < zb) {
                if (flags & GOMELT_STEP_WRITE_S1) p.S1out[n] = s1 ? 1.f : 0.f;
                if (flags & GOMELT_STEP_ACCUM) {
                    // cF:3568-3578
                    const bool prev = p.S2prev[n] != 0;
                    float ac = p.accum[n];
                    const float reset = (!prev && s2) ? ac : 0.f;
                    p.maxacc[n] = fmaxf(reset, p.maxacc[n]);
                    p.accum[n] = ac + (s2 ? p.dt : 0.f) - reset;
                }
                if (flags & GOMELT_STEP_WRITE_S2) p.S2out[n] = s2 ? 1 : 0;
            }
        }
        // prefetch plane l+1
        if (in_dom && (l + 1 <= lload_max)) {
            Tld = __ldg(p.T0 + (l + 1) * P + col);
            Sld = __ldg(p.S1 + (l + 1) * P + col);
        }
        // ---- B: x stage (pair sums / differences along x), publish -------------------
        {
            const float Tr = __shfl_down_sync(0xffffffffu, Tn, 1);
            const float kr = __shfl_down_sync(0xffffffffu, kn, 1);
            const float mr = __shfl_down_sync(0xffffffffu, mn, 1);
            s_fwd[buf][ty][tx] = make_float4(Tn + Tr, Tr - Tn, kn + kr, mn + mr);
            s_bwd[buf][ty][tx] = make_float4(high0, high1, mz, 0.f);
        }
        __syncthreads();
        // ---- H: finalise node plane l-2 ------------------------------------------------
        {
            const float4 b = s_bwd[buf][ty > 0 ? ty - 1 : 0][tx];
            const float E0 = low0 + b.x;
            const float E1 = low1 + b.y;
            const float my = mz + b.z;
            const float left = E0 - E1;
            const float right = E0 + E1;
            const float KT = left + __shfl_up_sync(0xffffffffu, right, 1);
            const float mnode = my + __shfl_up_sync(0xffffffffu, my, 1);
            const int f = l - 2;
            if (out_xy && f >= za && f < zb) {
                const long long n = f * P + col;
                float Tnew;
                if (f < nzl) {
                    float r = p.rhs ? __ldg(p.rhs + n) : 0.f;
                    if (p.srcz) r = fmaf(sfx * sfy, __ldg(p.srcz + f), r);
                    if (p.topflux && f == nzl - 1) r += __ldg(p.topflux + col);
                    Tnew = fmaf(r - KT, __fdividef(p.cdt, mnode), Tm2);
                } else {
                    Tnew = p.pk.T_amb;  // substitute_Tbar cF:2183
                }
                bool skip = false;
                if (flags & GOMELT_STEP_BC_CONST) {  // assignBCs order: y-, y+, x-, x+, z-
                    if (j == 0) Tnew = p.bc[0];
                    if (j == ny - 1) Tnew = p.bc[1];
                    if (i == 0) Tnew = p.bc[2];
                    if (i == nx - 1) Tnew = p.bc[3];
                    if (f == 0) Tnew = p.bc[4];
                } else if (flags & GOMELT_STEP_SKIP_FACES) {
                    skip = on_face_xy || (f == 0);
                }
                if (flags & GOMELT_STEP_CLAMP) Tnew = fmaxf(p.pk.T_amb, Tnew);
                if (!skip) p.Tout[n] = Tnew;
            }
        }
        // ---- D: y stage -> face(l) ------------------------------------------------------
        const float4 a = s_fwd[buf][ty][tx];
        const float4 c = s_fwd[buf][ty < BY - 1 ? ty + 1 : ty][tx];
        const float G00 = a.x + c.x, G01 = c.x - a.x, G10 = a.y + c.y, G11 = c.y - a.y;
        const float kf = a.z + c.z, mf = a.w + c.w;
        // ---- E: element layer e = l-1 (between planes l-1 and l) -------------------------
        {
            const int e = l - 1;
            const bool lay = elem_xy && (e >= l0) && (e + 1 <= lload_max) && (e + 1 <= zb);
            const float k8 = lay ? (kfp + kf) : 0.f;
            const float m8 = lay ? (mfp + mf) : 0.f;
            // forward z: H[.,.,0] = prev + cur ; H[.,.,1] = cur - prev
            const float H000 = F00 + G00, H001 = G00 - F00;  // H000 has lambda = 0
            const float H010 = F01 + G01, H011 = G01 - F01;
            const float H100 = F10 + G10, H101 = G10 - F10;
            const float H110 = F11 + G11, H111 = G11 - F11;
            (void)H000;
            // scaled, backward z:  bottom' = l0*H0 - l1*H1,  top' = l0*H0 + l1*H1
            const float t00 = p.lam[4] * H001;  // (sx,sy)=(0,0): lambda[0]=0
            const float g01 = p.lam[2] * H010, g10 = p.lam[1] * H100, g11 = p.lam[3] * H110;
            const float b01 = fmaf(-p.lam[6], H011, g01), u01 = fmaf(p.lam[6], H011, g01);
            const float b10 = fmaf(-p.lam[5], H101, g10), u10 = fmaf(p.lam[5], H101, g10);
            const float b11 = fmaf(-p.lam[7], H111, g11), u11 = fmaf(p.lam[7], H111, g11);
            // z reduction onto node plane l-1: W = top(layer l-2) + bottom(layer l-1)
            const float W00 = fmaf(-k8, t00, tk00);
            const float W01 = fmaf(k8, b01, tk01);
            const float W10 = fmaf(k8, b10, tk10);
            const float W11 = fmaf(k8, b11, tk11);
            tk00 = k8 * t00;
            tk01 = k8 * u01;
            tk10 = k8 * u10;
            tk11 = k8 * u11;
            mz = m8p + m8;
            m8p = m8;
            // backward y
            low0 = W00 - W01;
            high0 = W00 + W01;
            low1 = W10 - W11;
            high1 = W10 + W11;
        }
        F00 = G00; F01 = G01; F10 = G10; F11 = G11; kfp = kf; mfp = mf;
        Tm2 = Tm1;
        Tm1 = Tn;
    }
}

static int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return (v && *v) ? atoi(v) : dflt;
}

template <int RY, int WPB, int FEAT, int MINB = 1>
static void launch_v2(const StepParams& sp, int nch, cudaStream_t st) {
    dim3 block(32, WPB);
    const int strips = (sp.ny + RY - 1) / RY;
    dim3 grid((sp.nx + 2 * K1_TX - 1) / (2 * K1_TX), (strips + WPB - 1) / WPB, nch);
    level_step_v2<RY, WPB, FEAT, MINB><<<grid, block, 0, st>>>(sp);
}

// exact-feature instances for the hot call shapes; anything else runs the generic instance
constexpr int F_L3_BENCH = K1F_SRC | K1F_FLUX | K1F_S1OUT | K1F_CLAMP;
constexpr int F_L3_SUB = F_L3_BENCH | K1F_SKIP;
constexpr int F_L3_SUB2 = F_L3_SUB | K1F_S2OUT | K1F_ACCUM;
constexpr int F_L1_DWELL = K1F_FLUX | K1F_BCCONST;
constexpr int F_L1_DWELL_SUB = F_L1_DWELL | K1F_NSUB;

static int launch_step(const StepParams& sp, cudaStream_t st) {
    const int nch = (sp.zend - sp.zbeg + sp.zchunk - 1) / sp.zchunk;
    static const int variant = env_int("GOMELT_K1_VARIANT", 2);  // 1 = v1 (smem, 1 column / thread); dev A/B only
    static const int generic_only = env_int("GOMELT_K1_GENERIC", 0);
    constexpr int RY = 4, WPB = 1;  // one warp per CTA: warps are independent, finest SM balance
    if (variant == 1) {
        constexpr int BY = 8;
        dim3 block(32, BY);
        dim3 grid((sp.nx + 29) / 30, (sp.ny + BY - 3) / (BY - 2), nch);
        level_step_kernel<BY><<<grid, block, 0, st>>>(sp);
    } else if (generic_only) {
        launch_v2<RY, WPB, K1F_ALL | K1F_GENERIC>(sp, nch, st);
    } else {
        switch (sp.feat) {
            case F_L3_BENCH: {
                static const int wpb = env_int("GOMELT_K1_WPB", WPB);  // dev tuning knob
                static const int exp = env_int("GOMELT_K1_EXP", 0);    // dev A/B of tile shape / register cap
                if (exp == 1) launch_v2<4, 1, F_L3_BENCH, 12>(sp, nch, st);
                else if (exp == 2) launch_v2<3, 1, F_L3_BENCH, 1>(sp, nch, st);
                else if (exp == 3) launch_v2<3, 1, F_L3_BENCH, 12>(sp, nch, st);
                else if (exp == 4) launch_v2<2, 1, F_L3_BENCH, 1>(sp, nch, st);
                else if (exp == 5) launch_v2<2, 1, F_L3_BENCH, 16>(sp, nch, st);
                else if (exp == 6) launch_v2<4, 1, F_L3_BENCH, 10>(sp, nch, st);
                else if (wpb == 2) launch_v2<RY, 2, F_L3_BENCH>(sp, nch, st);
                else if (wpb == 4) launch_v2<RY, 4, F_L3_BENCH>(sp, nch, st);
                else launch_v2<RY, WPB, F_L3_BENCH>(sp, nch, st);
                break;
            }
            case F_L3_SUB: launch_v2<RY, WPB, F_L3_SUB>(sp, nch, st); break;
            case F_L3_SUB2: launch_v2<RY, WPB, F_L3_SUB2>(sp, nch, st); break;
            case F_L1_DWELL: launch_v2<RY, WPB, F_L1_DWELL>(sp, nch, st); break;
            case F_L1_DWELL_SUB: launch_v2<RY, WPB, F_L1_DWELL_SUB>(sp, nch, st); break;
            default: launch_v2<RY, WPB, K1F_ALL | K1F_GENERIC>(sp, nch, st); break;
        }
    }
    return check_launch("gomelt_level_step_f32");
}

}  // namespace gomelt

using namespace gomelt;

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
    }
    sp.cdt = (float)(64.0 * (double)a->dt / V);
    sp.dt = a->dt;
    sp.pk = fold_props(*props);
    sp.fk = fold_flux(*props, g);
    sp.T0 = a->T0; sp.S1 = a->S1; sp.rhs = a->rhs;
    sp.srcx = a->src_x; sp.srcy = a->src_y; sp.srcz = a->src_z; sp.scoef = a->src_coef;
    sp.topflux = a->topflux;
    sp.Tout = a->T_out; sp.S1out = a->S1_out; sp.S2out = a->S2_out;
    sp.S2prev = a->S2_prev; sp.accum = a->accum; sp.maxacc = a->max_accum;
    for (int q = 0; q < 5; ++q) sp.bc[q] = a->bc5[q];
    sp.flags = a->flags;
    sp.zbeg = zbeg; sp.zend = zend;
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
        sp.feat = f;
    }
    sp.zchunk = a->z_chunk > 0 ? a->z_chunk : (zend - zbeg);
    return launch_step(sp, (cudaStream_t)stream);
}
