// K1 v2 — fused level step, register-tiled and packed (sm_100a).  See DESIGN.md "K1".
//
// Same mathematics as v1 (k_level_step.cu header): Haar-diagonalised hex8 stiffness, lumped mass,
// forward Euler, no (ne,8)/(ne,8,8) arrays.  What changed is the mapping:
//
//  * one WARP owns a 60 x RY patch of node columns (+1 halo ring) and marches in z; no shared
//    memory, no barriers: the halo ring is recomputed from (L1/L2-resident) redundant loads;
//  * every thread carries TWO node columns 30 apart in x as the two halves of a float2 and all
//    arithmetic is issued as packed FADD2 / FMUL2 / FFMA2 (Blackwell f32x2): the kernel is
//    issue-bound, not FP32-throughput-bound, and packing halves the FP issue slots;
//  * RY rows per thread: the y stage of the analysis / synthesis runs in registers; only the
//    x stage crosses lanes (3 + 2 shuffles per node);
//  * stage order  x -> z -> y  (analysis) and  y -> z -> x  (synthesis) keeps the state carried
//    from plane to plane at 4 + 4 values per column; the plane loop is unrolled by two so the
//    carried state ping-pongs between two register sets instead of being moved;
//  * loads are unpredicated: out-of-domain columns / rows read a clamped (valid) address and are
//    neutralised downstream - element columns by the mask xm, element rows by a warp-uniform test;
//  * analysis, element layer and explicit update are fused row by row, so the only arrays that
//    live across the row loop are the carried state and the prefetched next plane.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace gomelt {

struct StepParams {
    int nx, ny, nz, nzl;   // nzl = active planes
    long long nsub;
    float lam[8];          // lambda'[sx + 2 sy + 4 sz] (already /64: two 1/8 factors folded)
    float lamA[3], lamB[3];  // v3: l_s + l_d and l_s - l_d of the (sx,sz) = (0,1), (1,0), (1,1) mode pairs
    float cdt;             // 64 dt / V
    float dt;
    PropK pk;
    const float* T0;
    const float* S1;
    const float* rhs;
    const float* srcx;
    const float* srcy;
    const float* srcz;
    float scoef;
    const float* topflux;
    float* Tout;
    float* S1out;
    uint8_t* S2out;
    const uint8_t* S2prev;
    float* accum;
    float* maxacc;
    float bc[5];
    int flags;
    int zchunk;
    int zbeg, zend;        // planes to finalise
    int feat;              // K1F_* bits of this call (v2)
    int nsub_planes, nsub_rem;  // n_substrate = nsub_planes * nx*ny + nsub_rem
    FluxK fk;              // top-surface flux constants (K1F_FLUX)
    float* peer_lo;        // neighbour ghost planes in peer memory (K1F_PEER), or nullptr
    float* peer_hi;
    int exp;               // dev experiments (GOMELT_K1_EXP), 0 in production
    // halo exchange with release / acquire counters (halo_exchange_kernel in k_level_step_v3.cuh): this rank's counter
    // block, the neighbours' (peer-mapped), sweep sequence number
    unsigned* hsync;
    unsigned* hsync_lo;
    unsigned* hsync_hi;
    unsigned hseq;
    int ran_v3;            // set by the launcher: the fast kernel took the call (its Dirichlet faces come from a face pass)
    unsigned* bkq;         // hot-plane work queue of the melt-time bookkeeping (gomelt_step_args_t.bk_queue) or nullptr
    unsigned bkq_cap;      // its capacity in entries
    int bkq_reset;         // zero its header before the step (0: the previous sweep left it zeroed)
    int bkq_force;         // use the queue on small grids too (tests)
    // v3 normalisation: stiffness modes divided by s = lambda'[2], masses by cdt * s, loads by s (so that
    // T_new = T + (rr/s - KT/s) / (mnode/(cdt s)) needs neither the lambda'[2] nor the cdt multiply)
    float n_ca0, n_ca1, n_cmushy, n_cfluid, n_inv_s, n_wq;
    alignas(64) CUtensorMap tm[2];   // K1F_TMA: T0 and S1 as 1-D tensors of nn floats, one box = one ring row
    alignas(64) CUtensorMap tm2[2];  // ... and as overlapping 2-D views (x + y (nx & ~3)), one box = six ring rows
    int tm2_ok;
};

#define GM_DI __device__ __forceinline__

struct f2 {
    float2 v;
};
GM_DI f2 mk2(float a, float b) { return f2{make_float2(a, b)}; }
GM_DI f2 splat(float a) { return f2{make_float2(a, a)}; }
GM_DI f2 operator+(f2 a, f2 b) { return f2{__fadd2_rn(a.v, b.v)}; }
GM_DI f2 operator-(f2 a, f2 b) { return f2{__ffma2_rn(b.v, make_float2(-1.f, -1.f), a.v)}; }
GM_DI f2 operator*(f2 a, f2 b) { return f2{__fmul2_rn(a.v, b.v)}; }
GM_DI f2 fma2(f2 a, f2 b, f2 c) { return f2{__ffma2_rn(a.v, b.v, c.v)}; }
GM_DI f2 neg(f2 a) { return mk2(-a.v.x, -a.v.y); }
GM_DI f2 shdn(f2 a) {
    return mk2(__shfl_down_sync(0xffffffffu, a.v.x, 1), __shfl_down_sync(0xffffffffu, a.v.y, 1));
}
GM_DI f2 shup(f2 a) {
    return mk2(__shfl_up_sync(0xffffffffu, a.v.x, 1), __shfl_up_sync(0xffffffffu, a.v.y, 1));
}

GM_DI float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// computeStateProperties cF:2567-2614 for one node as straight-line selects (the compiler turns the
// equivalent nested ternaries into divergent branches).  kb / cs = bulk conductivity / solid rho*cp
// already evaluated (packed) by the caller; sub = 1 forces S1 (substrate).  Returns S1' (0/1) and S2.
GM_DI void props_sel(const PropK& q, float T, float S1in, int sub, float kb, float cs, float& k, float& m,
                     float& s1f, int& s2) {
    asm("{\n\t.reg .pred p1, p2, p3;\n\t"
        "setp.ge.f32 p2, %4, %6;\n\t"          // S2 = T >= T_liquidus
        "setp.gt.f32 p3, %4, %7;\n\t"          // T > T_solidus (mushy when also !S2)
        "setp.gt.f32 p1, %5, 0f3EFF7CEE;\n\t"  // S1 > 0.499
        "setp.ne.or.s32 p1, %8, 0, p1;\n\t"    // | substrate
        "selp.f32 %0, %9, %10, p1;\n\t"        // bulk : powder
        "selp.f32 %0, %11, %0, p2;\n\t"        // fluid
        "selp.f32 %1, %12, %13, p3;\n\t"       // mushy : solid
        "selp.f32 %1, %14, %1, p2;\n\t"        // fluid
        "or.pred p1, p1, p2;\n\t"
        "selp.f32 %2, 0f3F800000, 0f00000000, p1;\n\t"
        "selp.s32 %3, 1, 0, p2;\n\t}"
        : "=&f"(k), "=&f"(m), "=f"(s1f), "=r"(s2)
        : "f"(T), "f"(S1in), "f"(q.T_liq), "f"(q.T_sol), "r"(sub), "f"(kb), "f"(q.k_powder), "f"(q.k_fluid),
          "f"(q.c_mushy), "f"(cs), "f"(q.c_fluid));
}

// Same without the state outputs (halo rows, calls that do not write S1 / S2).
GM_DI void props_sel_km(const PropK& q, float T, float S1in, int sub, float kb, float cs, float& k, float& m) {
    asm("{\n\t.reg .pred p1, p2, p3;\n\t"
        "setp.ge.f32 p2, %2, %4;\n\t"
        "setp.gt.f32 p3, %2, %5;\n\t"
        "setp.gt.f32 p1, %3, 0f3EFF7CEE;\n\t"
        "setp.ne.or.s32 p1, %6, 0, p1;\n\t"
        "selp.f32 %0, %7, %8, p1;\n\t"
        "selp.f32 %0, %9, %0, p2;\n\t"
        "selp.f32 %1, %10, %11, p3;\n\t"
        "selp.f32 %1, %12, %1, p2;\n\t}"
        : "=&f"(k), "=&f"(m)
        : "f"(T), "f"(S1in), "f"(q.T_liq), "f"(q.T_sol), "r"(sub), "f"(kb), "f"(q.k_powder), "f"(q.k_fluid),
          "f"(q.c_mushy), "f"(cs), "f"(q.c_fluid));
}

// Plane base pointers are made opaque so that every access is base + 4 * (32-bit offset) =
// one IMAD.WIDE, instead of a re-associated 64-bit index per access.
template <typename T>
GM_DI T* opaque(T* q) {
    asm volatile("" : "+l"(q));
    return q;
}

constexpr int K1_TX = 30;  // owned columns per half-warp tile (32 lanes - 2 halo)
constexpr int K1_SRCZ_MAX = 254;  // planes per z-chunk when the call carries source tables

// Both halves of a packed pair through ONE address: q points at the column of half .y (always >= 0), half
// .x is K1_TX columns (120 bytes) below it as an immediate offset, so a pair costs one 64-bit add instead of
// two widened index computations.  fa / fb guard the two accesses (columns outside the grid are never
// touched; a guarded-off load keeps the old - finite - register value, which is neutralised downstream).
GM_DI void ld2(const char* q, int fa, int fb, f2& v) {
    asm("{\n\t.reg .pred pa, pb;\n\t"
        "setp.ne.s32 pa, %3, 0;\n\t"
        "setp.ne.s32 pb, %4, 0;\n\t"
        "@pa ld.global.nc.f32 %0, [%2+-120];\n\t"
        "@pb ld.global.nc.f32 %1, [%2];\n\t}"
        : "+f"(v.v.x), "+f"(v.v.y)
        : "l"(q), "r"(fa), "r"(fb));
}
GM_DI void st2(char* q, int fa, int fb, f2 v) {
    asm volatile("{\n\t.reg .pred pa, pb;\n\t"
                 "setp.ne.s32 pa, %3, 0;\n\t"
                 "setp.ne.s32 pb, %4, 0;\n\t"
                 "@pa st.global.f32 [%0+-120], %1;\n\t"
                 "@pb st.global.f32 [%0], %2;\n\t}"
                 ::"l"(q), "f"(v.v.x), "f"(v.v.y), "r"(fa), "r"(fb)
                 : "memory");
}
// A load pinned in program order (volatile asm is not moved across the volatile stores): used for the
// one-plane-ahead fetch of the source z-factor, which the scheduler otherwise sinks next to its use.
GM_DI float ldg_pinned(const float* q) {
    float v;
    asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(q));
    return v;
}
static_assert(K1_TX * 4 == 120, "ld2 / st2 hard-code the half distance");

// Compile-time feature mask: a kernel instance carries code only for the features in FEAT.
// K1F_GENERIC additionally tests the run-time flags / pointers, so one instance serves any call.
enum : int {
    K1F_RHS = 1 << 0,      // rhs array
    K1F_SRC = 1 << 1,      // rank-1 source tables
    K1F_TOP = 1 << 2,      // top-plane flux load
    K1F_S1OUT = 1 << 3,
    K1F_S2OUT = 1 << 4,
    K1F_ACCUM = 1 << 5,
    K1F_BCCONST = 1 << 6,
    K1F_SKIP = 1 << 7,
    K1F_CLAMP = 1 << 8,
    K1F_NSUB = 1 << 9,     // substrate override present (n_substrate > 0)
    K1F_FLUX = 1 << 10,    // top-surface flux (computeConvRadBC) evaluated in the step from T0's top plane
    K1F_PEER = 1 << 11,    // boundary planes of T_out are also stored to the z-neighbours' ghost planes (NVLink)
    K1F_S1INPLACE = 1 << 13,  // v3: S1_out is S1: a node's state is stored only when it changed (it rarely does)
    K1F_TMA = 1 << 14,     // v3: raw planes through a TMA ring in shared memory (1-D tensor maps StepParams::tm)
    K1F_PF = 1 << 12,      // v3: L2 prefetch of the plane three ahead (latency-bound many-wave shapes)
    K1F_ALL = (1 << 12) - 1,
    K1F_GENERIC = 1 << 30,
};

// computeConvRadBC cF:2270-2293 with the hardware approximations (rcp / sqrt / ex2, <= 2 ulp each): used
// by the fused top-plane epilogue, where 8 (RY + 1) flux evaluations per thread sit on every warp's tail.
// The surface load is a small part of a node's update, so T stays well inside the 1e-5 parity tolerance
// (tests/test_level_step_gpu.py checks it against the exact-libm stand-alone kernel and the oracle).
GM_DI float flux_fast(const FluxK& f, float Tq) {
    Tq = fminf(Tq, f.T_cap);
    const float invT = rcp_approx(Tq);
    const float E_pv = fmaf(f.cp_fluid, Tq - f.T_amb, f.Lev);
    float root, ex;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(root) : "f"(f.CM * invT));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"((-1.4426950408889634f * f.CT) * (invT - f.invT_b)));
    const float S = f.evcCP * ex * root * E_pv;
    const float T2 = Tq * Tq;
    const float q = f.h_conv * (f.T_amb - Tq) + f.sig_eps * (f.Tamb4 - T2 * T2) - S;
    return q * 1e-6f;
}

template <int RY>
struct K1State {           // x-staged fields of one plane (loaded rows) + T of the owned rows
    f2 Xs[RY + 2], Xd[RY + 2], kx[RY + 2], mx[RY + 2], T[RY];
};
struct FinalPtrs {         // plane base pointers (bytes) of the plane being finalised
    char* out;
    const char* rhs;
};
template <int RY, bool STAGE>
struct K1Raw;              // prefetched raw plane (T0 and S1 of the RY + 2 loaded rows, both halves)
template <int RY>
struct K1Raw<RY, false> {  // held in registers
    f2 Tr[RY + 2], Sr[RY + 2];
    GM_DI f2 T(int r) const { return Tr[r]; }
    GM_DI f2 S(int r) const { return Sr[r]; }
};
template <int RY>
struct K1Raw<RY, true> {   // staged in shared memory by cp.async: [field][row][lane] float2, b = this lane's slot
    const float2* b;
    GM_DI f2 T(int r) const { return f2{b[r * 32]}; }
    GM_DI f2 S(int r) const { return f2{b[(RY + 2 + r) * 32]}; }
};

GM_DI unsigned smem_u32(const void* q) { return (unsigned)__cvta_generic_to_shared(q); }
// 4-byte cp.async of both halves of a pair (half .x 120 bytes below the address of half .y); a guarded-off
// half is zero-filled (src-size 0: nothing is read).
GM_DI void cp_async_pair(unsigned dst, const char* q, int fa, int fb) {
    asm volatile("cp.async.ca.shared.global [%0], [%1+-120], 4, %2;\n\t"
                 "cp.async.ca.shared.global [%0+4], [%1], 4, %3;" ::"r"(dst), "l"(q), "r"(fa * 4), "r"(fb * 4)
                 : "memory");
}
GM_DI void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
GM_DI void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int RY, int WPB, int FEAT, int MINB = 1, bool STAGE = false>
__global__ void __launch_bounds__(32 * WPB, MINB) level_step_v2(const __grid_constant__ StepParams p) {
    constexpr int NR = RY + 2;  // loaded rows
    using Raw = K1Raw<RY, STAGE>;
    __shared__ float2 s_stage[STAGE ? WPB : 1][2][2][STAGE ? NR : 1][STAGE ? 32 : 1];  // [warp][buffer][T|S][row][lane]
    constexpr bool GEN = (FEAT & K1F_GENERIC) != 0;
    const int rtf = p.feat;     // run-time feature bits, set by the launcher
#define K1_HAS(bit) (((FEAT & (bit)) != 0) && (!GEN || (rtf & (bit)) != 0))
    const int lane = threadIdx.x;
    const int nx = p.nx, ny = p.ny, nz = p.nz, nzl = p.nzl;
    const int j0 = (blockIdx.y * WPB + threadIdx.y) * RY;  // first owned row
    if (j0 >= ny) return;                                   // whole warp; no barriers in this kernel
    const int ia = blockIdx.x * (2 * K1_TX) + lane - 1;     // column of half .x ; half .y is 30 further
    const int ib = ia + K1_TX;
    const bool lown = (lane >= 1) && (lane <= K1_TX);
    const bool owna = lown && (ia < nx), ownb = lown && (ib < nx);
    const int la = (ia >= 0 && ia < nx) ? 1 : 0, lb = (ib < nx) ? 1 : 0;  // load guards
    const int oa = owna ? 1 : 0, ob = ownb ? 1 : 0;                      // store guards
    const f2 xm = mk2((ia >= 0 && ia + 1 < nx) ? 1.f : 0.f, (ib >= 0 && ib + 1 < nx) ? 1.f : 0.f);
    const int P = nx * ny;  // nn < 2^31 is validated by the launcher
    const int za = p.zbeg + blockIdx.z * p.zchunk;
    const int zb = min(p.zend, za + p.zchunk);  // this warp finalises node planes [za, zb)
    const int lfirst = max(za - 1, 0);
    const int llast = min(min(zb, nz - 1), nzl - 1);  // last plane that carries data
    // rows are clamped (warp-uniform), columns are guarded: in-plane BYTE offset of half .y per loaded row
    const int cola = min(max(ia, 0), nx - 1), colb = min(max(ib, 0), nx - 1);
    unsigned offb[NR];
#pragma unroll
    for (int r = 0; r < NR; ++r) offb[r] = 4u * (unsigned)(min(max(j0 - 1 + r, 0), ny - 1) * nx + ib);
    const bool edge_lo = (j0 == 0), edge_hi = (j0 + RY >= ny);
    const int nsub_planes = p.nsub_planes, nsub_rem = p.nsub_rem;  // substrate: whole planes (cF:575-578) + rest

    // A value loaded outside the plane loop is passed through one real ALU op right away (xor with an
    // opaque zero), so that its scoreboard wait happens here and not at its first use inside the loop,
    // where the same scoreboard also tracks the next plane's prefetch (measured: 1.4 stall cycles / issue).
    const int opaque_zero = nx >> 31;
    auto settle = [&](float x) { return __int_as_float(__float_as_int(x) ^ opaque_zero); };
    f2 sfx = splat(0.f);
    float sfy[RY];
#pragma unroll
    for (int r = 0; r < RY; ++r) sfy[r] = 0.f;
    if (K1_HAS(K1F_SRC)) {
        sfx = mk2(__ldg(p.srcx + cola) * p.scoef, __ldg(p.srcx + colb) * p.scoef);
#pragma unroll
        for (int r = 0; r < RY; ++r) sfy[r] = settle(__ldg(p.srcy + min(j0 + r, ny - 1)));
    }

    f2 Tt0[RY], Tt1[RY], myp[RY];  // top contributions (stiffness sx = 0 / 1, mass) of the previous layer
#pragma unroll
    for (int r = 0; r < RY; ++r) Tt0[r] = Tt1[r] = myp[r] = splat(0.f);

    auto load_plane = [&](int l, Raw& raw) {
        const char* Tl = (const char*)(p.T0 + (size_t)l * P);
        const char* Sl = (const char*)(p.S1 + (size_t)l * P);
        if constexpr (STAGE) {
            const unsigned d = smem_u32(raw.b);
#pragma unroll
            for (int r = 0; r < NR; ++r) {
                cp_async_pair(d + (unsigned)(r * 32 * sizeof(float2)), Tl + offb[r], la, lb);
                cp_async_pair(d + (unsigned)((NR + r) * 32 * sizeof(float2)), Sl + offb[r], la, lb);
            }
            cp_async_commit();
        } else {
#pragma unroll
            for (int r = 0; r < NR; ++r) {
                ld2(Tl + offb[r], la, lb, raw.Tr[r]);
                ld2(Sl + offb[r], la, lb, raw.Sr[r]);
            }
        }
    };
    // plane data of the older of (at most) two groups in flight has landed
    auto landed = [&](bool newer_in_flight) {
        if constexpr (STAGE) {
            if (newer_in_flight) cp_async_wait<1>();
            else cp_async_wait<0>();
        }
    };

    // ---- node state of loaded row r of plane l (+ S1 / S2 / melt-time outputs) and its x stage -----
    auto row_a = [&](size_t pl, char* so, int r, bool wr, int subrow, f2 T, f2 S, f2& xs, f2& xd, f2& kxr, f2& mxr) {
        const f2 kb = fma2(splat(p.pk.k_a1), T, splat(p.pk.k_a0)), cs = fma2(splat(p.pk.c_a1), T, splat(p.pk.c_a0));
        int suba = 0, subb = 0;
        const int eb = (int)(offb[r] >> 2), ea = eb - K1_TX;  // in-plane node ids of the two halves
        if (K1_HAS(K1F_NSUB)) {
            suba = ea < subrow;
            subb = eb < subrow;
        }
        f2 kn, mn, s1f = splat(0.f);
        int s2a = 0, s2b = 0;
        const bool outs = (r >= 1 && r <= RY) && (K1_HAS(K1F_S1OUT) || K1_HAS(K1F_S2OUT) || K1_HAS(K1F_ACCUM));
        if (outs) {
            props_sel(p.pk, T.v.x, S.v.x, suba, kb.v.x, cs.v.x, kn.v.x, mn.v.x, s1f.v.x, s2a);
            props_sel(p.pk, T.v.y, S.v.y, subb, kb.v.y, cs.v.y, kn.v.y, mn.v.y, s1f.v.y, s2b);
        } else {
            props_sel_km(p.pk, T.v.x, S.v.x, suba, kb.v.x, cs.v.x, kn.v.x, mn.v.x);
            props_sel_km(p.pk, T.v.y, S.v.y, subb, kb.v.y, cs.v.y, kn.v.y, mn.v.y);
        }
        if (outs) {
            // warp-uniform conditions are folded into the store guards, not branched on: branches split the
            // unrolled plane body into per-row basic blocks and stop the scheduler from interleaving rows
            // S1' is a node-local function of (T0, S1): the halo planes of a chunk (wr false) may be written
            // too - their owner writes the same value - so only the row guard remains (plane-invariant)
            const int gw = (j0 + r - 1 < ny) ? 1 : 0;
            if (K1_HAS(K1F_S1OUT)) st2(so + offb[r], oa & gw, ob & gw, s1f);
            if ((K1_HAS(K1F_ACCUM) || K1_HAS(K1F_S2OUT)) && gw && wr) {
                if (K1_HAS(K1F_ACCUM)) {  // cF:3568-3578
                    if (owna) {
                        const size_t n = pl + ea;
                        const bool prev = p.S2prev[n] != 0;
                        const float ac = p.accum[n];
                        const float reset = (!prev && s2a) ? ac : 0.f;
                        p.maxacc[n] = fmaxf(reset, p.maxacc[n]);
                        p.accum[n] = ac + (s2a ? p.dt : 0.f) - reset;
                    }
                    if (ownb) {
                        const size_t n = pl + eb;
                        const bool prev = p.S2prev[n] != 0;
                        const float ac = p.accum[n];
                        const float reset = (!prev && s2b) ? ac : 0.f;
                        p.maxacc[n] = fmaxf(reset, p.maxacc[n]);
                        p.accum[n] = ac + (s2b ? p.dt : 0.f) - reset;
                    }
                }
                if (K1_HAS(K1F_S2OUT)) {
                    if (owna) p.S2out[pl + ea] = (uint8_t)s2a;
                    if (ownb) p.S2out[pl + eb] = (uint8_t)s2b;
                }
            }
        }
        const f2 Tr = shdn(T), kr = shdn(kn), mr = shdn(mn);
        xs = Tr + T;
        xd = Tr - T;
        kxr = (kr + kn) * xm;
        mxr = (mr + mn) * xm;
    };

    // ---- first plane of a chunk: x stage only -----------------------------------------------------
    auto first_plane = [&](int l, const Raw& raw, K1State<RY>& st) {
        const bool wr = (l >= za) && (l < zb);
        const int subrow = (l < nsub_planes) ? (1 << 30) : ((l == nsub_planes) ? nsub_rem : 0);
        const size_t pl = (size_t)l * P;
        char* so = K1_HAS(K1F_S1OUT) ? (char*)(p.S1out + pl) : nullptr;
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            const f2 Tr_ = raw.T(r);
            row_a(pl, so, r, wr, subrow, Tr_, raw.S(r), st.Xs[r], st.Xd[r], st.kx[r], st.mx[r]);
            if (r >= 1 && r <= RY) st.T[r - 1] = Tr_;
        }
    };

    // ---- write one finalised owned row of plane f ---------------------------------------------------
    auto store_row = [&](int f, const FinalPtrs& fp, int r, f2 Tn, int g = 1) {
        const int j = j0 + r;
        bool ska = false, skb = false;
        if (K1_HAS(K1F_BCCONST)) {  // assignBCs order: y-, y+, x-, x+, z-
            if (j == 0) Tn = splat(p.bc[0]);
            if (j == ny - 1) Tn = splat(p.bc[1]);
            if (ia == 0) Tn.v.x = p.bc[2];
            if (ib == 0) Tn.v.y = p.bc[2];
            if (ia == nx - 1) Tn.v.x = p.bc[3];
            if (ib == nx - 1) Tn.v.y = p.bc[3];
            if (f == 0) Tn = splat(p.bc[4]);
        } else if (K1_HAS(K1F_SKIP)) {
            const bool fy = (j == 0) || (j == ny - 1) || (f == 0);
            ska = fy || (ia == 0) || (ia == nx - 1);
            skb = fy || (ib == 0) || (ib == nx - 1);
        }
        if (K1_HAS(K1F_CLAMP)) Tn = mk2(fmaxf(p.pk.T_amb, Tn.v.x), fmaxf(p.pk.T_amb, Tn.v.y));
        const int ga = (owna && !ska) ? g : 0, gb = (ownb && !skb) ? g : 0;
        st2(fp.out + offb[r + 1], ga, gb, Tn);
        if (K1_HAS(K1F_PEER)) {  // halo exchange fused into the step: plain stores to peer-mapped memory
            if (f == p.zbeg && p.peer_lo) st2((char*)p.peer_lo + offb[r + 1], ga, gb, Tn);
            if (f == p.zend - 1 && p.peer_hi) st2((char*)p.peer_hi + offb[r + 1], ga, gb, Tn);
        }
    };

    // ---- explicit update of owned row r of plane f from its assembled action (z0, z1, mz) ---------
    auto final_row = [&](int f, const FinalPtrs& fp, int r, f2 sz, bool topf, f2 Tf, f2 z0, f2 z1, f2 mz,
                         f2 fl = f2{make_float2(0.f, 0.f)}) {
        // x stage of the synthesis (shuffles are executed by the whole warp)
        const f2 KT = (z0 - z1) + shup(z0 + z1);
        const f2 mnode = mz + shup(mz);
        {   // rows past the grid are computed and guarded off (no branch)
            const int g = (j0 + r < ny) ? 1 : 0;
            f2 rr = splat(0.f);
            if (K1_HAS(K1F_RHS)) ld2(fp.rhs + offb[r + 1], la, lb, rr);
            if (K1_HAS(K1F_SRC)) rr = fma2(sz, splat(sfy[r]), rr);
            if (K1_HAS(K1F_FLUX)) rr = rr + fl;
            if (topf) {
                f2 tf = splat(0.f);
                ld2((const char*)p.topflux + offb[r + 1], la, lb, tf);
                rr = rr + tf;
            }
            const f2 w = splat(p.cdt) * mk2(rcp_approx(mnode.v.x), rcp_approx(mnode.v.y));
            store_row(f, fp, r, fma2(rr - KT, w, Tf), g);
        }
    };

    // ---- plane l: node state, x stage, then the element layer (l-1, l) row by row; finalises
    //      plane l-1 when do_final.  pv = state of plane l-1, cu <- state of plane l.
    auto step_plane = [&](int l, const Raw& raw, const K1State<RY>& pv, K1State<RY>& cu, bool do_final, f2 sz) {
        const bool wr = (l >= za) && (l < zb);
        const int subrow = (l < nsub_planes) ? (1 << 30) : ((l == nsub_planes) ? nsub_rem : 0);
        const int f = l - 1;
        const size_t pl = (size_t)l * P;
        char* so = K1_HAS(K1F_S1OUT) ? (char*)(p.S1out + pl) : nullptr;
        // A chunk's lower halo plane (f < za, once per chunk) is not finalised here: its rows are computed like
        // any other and stored to plane za instead, where the next step overwrites them (same thread, same
        // addresses, program order) - no per-plane guard in the loop.
        FinalPtrs fp;
        fp.out = (char*)(p.Tout + (do_final ? pl - P : pl));
        fp.rhs = K1_HAS(K1F_RHS) ? (const char*)(p.rhs + (pl - P)) : nullptr;
        // the top-flux plane nzl-1 is always the last data plane of its chunk: finalised by last_plane
        constexpr bool topf = false;
        const f2 l1 = splat(p.lam[1]), l2 = splat(p.lam[2]), l3 = splat(p.lam[3]), l4 = splat(p.lam[4]),
                 l5 = splat(p.lam[5]), l6 = splat(p.lam[6]), l7 = splat(p.lam[7]);
        f2 c00, c01, c10, c11, cm;            // carries of the previous element row
        f2 zl00, zl01, zl10, zl11, kzl, mzl;  // z-staged fields of the lower loaded row
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            f2 xs, xd, kxr, mxr;
            const f2 Tr_ = raw.T(r);
            row_a(pl, so, r, wr, subrow, Tr_, raw.S(r), xs, xd, kxr, mxr);
            cu.Xs[r] = xs; cu.Xd[r] = xd; cu.kx[r] = kxr; cu.mx[r] = mxr;
            if (r >= 1 && r <= RY) cu.T[r - 1] = Tr_;
            // z stage of the analysis
            const f2 zu00 = xs + pv.Xs[r], zu01 = xs - pv.Xs[r];
            const f2 zu10 = xd + pv.Xd[r], zu11 = xd - pv.Xd[r];
            const f2 kzu = kxr + pv.kx[r], mzu = mxr + pv.mx[r];
            if (r >= 1) {
                const int e = r - 1;  // element row between loaded rows e and e+1
                f2 k8 = kzu + kzl;
                f2 m8 = mzu + mzl;
                // element rows that touch an out-of-domain node row do not exist
                if ((e == 0 && edge_lo) || (edge_hi && j0 + e >= ny)) k8 = m8 = splat(0.f);
                // y stage: sums (sy = 0) and differences (sy = 1); H000 is never needed (lambda = 0)
                const f2 Hd00 = zu00 - zl00;
                const f2 Hs01 = zu01 + zl01, Hd01 = zu01 - zl01;
                const f2 Hs10 = zu10 + zl10, Hd10 = zu10 - zl10;
                const f2 Hs11 = zu11 + zl11, Hd11 = zu11 - zl11;
                // scaled by lambda'[sx + 2 sy + 4 sz]
                const f2 q00 = (l2 * Hd00) * k8;  // (sx,sz) = (0,0): only the sy = 1 mode acts
                const f2 g01 = l4 * Hs01, g10 = l1 * Hs10, g11 = l5 * Hs11;
                const f2 d01 = l6 * Hd01, d10 = l3 * Hd10, d11 = l7 * Hd11;
                if (e >= 1) {  // lower node row of this element row = loaded row e = owned row e-1
                    const f2 R00 = c00 - q00;
                    const f2 R01 = fma2(k8, g01 - d01, c01);
                    const f2 R10 = fma2(k8, g10 - d10, c10);
                    const f2 R11 = fma2(k8, g11 - d11, c11);
                    // z stage of the synthesis: bottom (plane l-1) and top (plane l) parts
                    const f2 z0 = Tt0[e - 1] + (R00 - R01);
                    const f2 z1 = Tt1[e - 1] + (R10 - R11);
                    Tt0[e - 1] = R00 + R01;
                    Tt1[e - 1] = R10 + R11;
                    const f2 my = cm + m8;
                    const f2 mz = myp[e - 1] + my;
                    myp[e - 1] = my;
                    final_row(do_final ? f : -1, fp, e - 1, sz, topf, pv.T[e - 1], z0, z1, mz);
                }
                if (e < RY) {
                    c00 = q00;
                    c01 = (g01 + d01) * k8;
                    c10 = (g10 + d10) * k8;
                    c11 = (g11 + d11) * k8;
                    cm = m8;
                }
            }
            zl00 = zu00; zl01 = zu01; zl10 = zu10; zl11 = zu11; kzl = kzu; mzl = mzu;
        }
    };

    // ---- source z-factor of plane f: the chunk's factors are copied to shared memory once (a per-plane
    //      uniform global load - even a predicated-off refill - makes the consumer wait on the scoreboard it
    //      shares with the next plane's prefetch: measured 0.9 stall cycles / issue).  The launcher keeps
    //      chunks of source-carrying calls within K1_SRCZ_MAX planes.
    __shared__ float s_srcz[(FEAT & K1F_SRC) ? WPB : 1][(FEAT & K1F_SRC) ? K1_SRCZ_MAX + 2 : 1];
    if (K1_HAS(K1F_SRC)) {
        for (int i = lane; i <= llast - lfirst; i += 32) s_srcz[threadIdx.y][i] = __ldg(p.srcz + lfirst + i);
        __syncwarp();
    }
    auto srcz_at = [&](int f) -> float { return K1_HAS(K1F_SRC) ? s_srcz[threadIdx.y][f - lfirst] : 0.f; };

    // ---- computeConvRadBC cF:2207-2301 fused: load of the top face of element layer nzl-2 on the owned
    //      nodes of plane nzl-1, from that plane's T0 (still in the prefetch registers).  This lane's
    //      element column spans its node column and the right neighbour's; per element 2x2 Gauss points,
    //      N2^T (q wq); a node sums its <= 4 elements in increasing element id like the reference's
    //      scatter-add: (i-1,j-1) a=2, (i,j-1) a=3, (i-1,j) a=1, (i,j) a=0.
    auto fused_top_flux = [&](const Raw& raw, f2* fl) {
        // The plane loop's registers are dead here; the element loop is kept rolled (one code copy: an
        // unrolled version evicts the plane loop from the instruction cache) by staging through shared memory.
        __shared__ float2 sT[WPB][NR][32];
        __shared__ float2 sA[WPB][RY + 1][4][32];
        const int wy = threadIdx.y;
#pragma unroll
        for (int r = 0; r < NR; ++r) sT[wy][r][lane] = raw.T(r).v;
        __syncwarp();
        const float* sTf = reinterpret_cast<const float*>(&sT[wy][0][0]);
        float* sAf = reinterpret_cast<float*>(&sA[wy][0][0][0]);
        const int lane1 = min(lane + 1, 31);
        const float g = 0.57735026918962576f;
        const float NA = 0.25f * (1.f + g) * (1.f + g), NB = 0.25f * (1.f + g) * (1.f - g),
                    NC = 0.25f * (1.f - g) * (1.f - g);
        const float wq = p.fk.wq;
#pragma unroll 1
        for (int it = 0; it < 2 * (RY + 1); ++it) {
            const int e = it >> 1, h = it & 1;  // element row between loaded rows e, e + 1; half of the pair
            const float T0n = sTf[(e * 32 + lane) * 2 + h], T1n = sTf[(e * 32 + lane1) * 2 + h];
            const float T2n = sTf[((e + 1) * 32 + lane1) * 2 + h], T3n = sTf[((e + 1) * 32 + lane) * 2 + h];
            const int ic = h ? ib : ia;
            const bool exists = (ic >= 0) && (ic + 1 < nx) && !((e == 0 && edge_lo) || (edge_hi && j0 + e >= ny));
            const float q0 = flux_fast(p.fk, fmaf(NB, T3n, fmaf(NC, T2n, fmaf(NB, T1n, NA * T0n)))) * wq;
            const float q1 = flux_fast(p.fk, fmaf(NC, T3n, fmaf(NB, T2n, fmaf(NA, T1n, NB * T0n)))) * wq;
            const float q2 = flux_fast(p.fk, fmaf(NB, T3n, fmaf(NA, T2n, fmaf(NB, T1n, NC * T0n)))) * wq;
            const float q3 = flux_fast(p.fk, fmaf(NA, T3n, fmaf(NB, T2n, fmaf(NC, T1n, NB * T0n)))) * wq;
            sAf[((e * 4 + 0) * 32 + lane) * 2 + h] = exists ? fmaf(NB, q3, fmaf(NC, q2, fmaf(NB, q1, NA * q0))) : 0.f;
            sAf[((e * 4 + 1) * 32 + lane) * 2 + h] = exists ? fmaf(NC, q3, fmaf(NB, q2, fmaf(NA, q1, NB * q0))) : 0.f;
            sAf[((e * 4 + 2) * 32 + lane) * 2 + h] = exists ? fmaf(NB, q3, fmaf(NA, q2, fmaf(NB, q1, NC * q0))) : 0.f;
            sAf[((e * 4 + 3) * 32 + lane) * 2 + h] = exists ? fmaf(NA, q3, fmaf(NB, q2, fmaf(NC, q1, NB * q0))) : 0.f;
        }
        __syncwarp();
        const int lm = max(lane - 1, 0);
#pragma unroll
        for (int r = 0; r < RY; ++r)  // owned row r = loaded row r + 1: upper node of element row r, lower of r + 1
            fl[r] = ((f2{sA[wy][r][2][lm]} + f2{sA[wy][r][3][lane]}) + f2{sA[wy][r + 1][1][lm]}) +
                    f2{sA[wy][r + 1][0][lane]};
    };

    // ---- last data plane of a chunk top: no layer above, its action is (Tt0, Tt1, myp) --------------
    auto last_plane = [&](int f, const f2* Tf, const Raw& raw) {
        f2 sz = splat(0.f);
        if (K1_HAS(K1F_SRC)) sz = sfx * splat(srcz_at(f));
        const bool topf = K1_HAS(K1F_TOP) && (f == nzl - 1);
        f2 fl[RY];
#pragma unroll
        for (int r = 0; r < RY; ++r) fl[r] = splat(0.f);
        if (K1_HAS(K1F_FLUX) && f == nzl - 1 && nzl >= 2) fused_top_flux(raw, fl);  // warp-uniform
        FinalPtrs fp;
        fp.out = (char*)(p.Tout + (size_t)f * P);
        fp.rhs = K1_HAS(K1F_RHS) ? (const char*)(p.rhs + (size_t)f * P) : nullptr;
#pragma unroll
        for (int r = 0; r < RY; ++r) final_row(f, fp, r, sz, topf, Tf[r], Tt0[r], Tt1[r], myp[r], fl[r]);
    };

    // ---- inactive planes (substitute_Tbar cF:2183) with the Dirichlet faces applied -------------
    auto fill_inactive = [&](int f) {
        FinalPtrs fp;
        fp.out = (char*)(p.Tout + (size_t)f * P);
        fp.rhs = nullptr;
#pragma unroll
        for (int r = 0; r < RY; ++r)
            if (j0 + r < ny) store_row(f, fp, r, splat(p.pk.T_amb));
    };

    int fdone = za;  // planes [za, fdone) are finalised
    if (lfirst <= llast) {
        Raw rawA, rawB;
        if constexpr (STAGE) {
            rawA.b = &s_stage[threadIdx.y][0][0][0][lane];
            rawB.b = &s_stage[threadIdx.y][1][0][0][lane];
        } else {  // zero-initialised: guarded-off loads keep these (finite) values
#pragma unroll
            for (int r = 0; r < NR; ++r) rawA.Tr[r] = rawA.Sr[r] = rawB.Tr[r] = rawB.Sr[r] = splat(0.f);
        }
        K1State<RY> stA, stB;
        load_plane(lfirst, rawA);
        if (lfirst + 1 <= llast) load_plane(lfirst + 1, rawB);
        landed(lfirst + 1 <= llast);
        first_plane(lfirst, rawA, stA);
        // main loop, unrolled by two so the carried state ping-pongs: (stA, rawB) -> stB, (stB, rawA) -> stA
        // source z-factor of the plane finalised next, fetched one plane ahead
        for (int l = lfirst + 1; l <= llast; l += 2) {
            if (l + 1 <= llast) load_plane(l + 1, rawA);
            landed(l + 1 <= llast);
            step_plane(l, rawB, stA, stB, l - 1 >= za, sfx * splat(srcz_at(l - 1)));
            if (l + 1 > llast) break;
            if (l + 2 <= llast) load_plane(l + 2, rawB);
            landed(l + 2 <= llast);
            step_plane(l + 1, rawA, stB, stA, l >= za, sfx * splat(srcz_at(l)));
        }
        if (llast >= za && llast < zb) {
            if (((llast - lfirst) & 1) != 0) last_plane(llast, stB.T, rawB);  // parity of the plane held in stB
            else last_plane(llast, stA.T, rawA);
        }
        fdone = max(za, llast + 1);
    }
    for (int f = fdone; f < zb; ++f) fill_inactive(f);  // planes >= nz_active
#undef K1_HAS
}

}  // namespace gomelt
