// K1 v3 — fused level step for the call shapes whose four side faces (x-, x+, y-, y+) and bottom face are
// Dirichlet (GOMELT_STEP_SKIP_FACES: child levels, faces come from the parent, assignBCsFine cF:1598-1620;
// GOMELT_STEP_BC_CONST: Level 1, assignBCs cF:1568-1595).  That is every call the steppers issue
// (stepGOMELT, subcycleGOMELT, stepGOMELTDwellTime); the natural-boundary form stays on v2.
//
// Same mathematics, mapping and plane pipeline as v2 (k_level_step_v2.cuh: one warp per 60 x RY column patch
// marching in z, packed f32x2, no shared memory in the plane loop).  What v3 removes is everything that
// exists only because a v2 tile may hang over the grid edge:
//
//  * tiles cover the OWNED region [1, nx-2] x [1, ny-2] only (face nodes are never stored by the step) and
//    the last tile / strip is shifted inwards so that all 64 x (RY+2) loaded columns / rows are inside the
//    grid: the overlap is computed twice with bit-identical results (same operations on the same data), so
//    the duplicate stores are benign.  No load guards, no element-column mask, no edge-row selects, one
//    store predicate (the 30 owned lanes);
//  * the substrate override (cF:2589-2590) is whole planes in every reference call (getSubstrateNodes
//    cF:562-579), so it becomes a per-plane threshold (S1 > -1) instead of a per-node compare;
//  * the y stage of the analysis and the lambda scaling are folded:  with A = l_s + l_d, B = l_s - l_d,
//      l_s (zu + zl) - l_d (zu - zl) = B zu + A zl,     l_s (zu + zl) + l_d (zu - zl) = A zu + B zl,
//    and B z is formed once per loaded row: 42 instead of 84 packed instructions per warp-plane.
//
//  * TMA plane ring + cold-plane path (K1F_TMA, the default): the raw planes of T0 and S1 arrive three planes ahead
//    through a 4-stage ring in shared memory filled by the TMA unit (an overlapping 2-D tensor-map view of the flat
//    field, see the ring in the kernel), and one vote per plane picks a plane body with 2 instead of 7 property
//    selects per node where nothing is above the solidus (bit-identical values).  DESIGN.md section 3.
//
// Measured on B200 the packed-FP instruction stream is bound by register-file read bandwidth (FADD2 / FFMA2
// with register operands occupy the scheduler for 2 / 3 cycles and nothing co-issues in their shadow:
// bench_tools/ubench_pipes.cu), i.e. by the instruction count itself - hence this variant.
#pragma once
#include "k_level_step_v2.cuh"

namespace gomelt {

// computeStateProperties cF:2567-2614 as selects; thr = 0.499, or -1 on substrate planes (forces S1).
GM_DI void props3(const PropK& q, float c_mushy, float c_fluid, float T, float S1in, float thr, float kb, float cs, float& k,
                  float& m, float& s1f) {
    asm("{\n\t.reg .pred p1, p2, p3;\n\t"
        "setp.ge.f32 p2, %3, %6;\n\t"
        "setp.gt.f32 p3, %3, %7;\n\t"
        "setp.gt.f32 p1, %4, %5;\n\t"
        "selp.f32 %0, %8, %9, p1;\n\t"
        "selp.f32 %0, %10, %0, p2;\n\t"
        "selp.f32 %1, %11, %12, p3;\n\t"
        "selp.f32 %1, %13, %1, p2;\n\t"
        "or.pred p1, p1, p2;\n\t"
        "selp.f32 %2, 0f3F800000, 0f00000000, p1;\n\t}"
        : "=&f"(k), "=&f"(m), "=f"(s1f)
        : "f"(T), "f"(S1in), "f"(thr), "f"(q.T_liq), "f"(q.T_sol), "f"(kb), "f"(q.k_powder), "f"(q.k_fluid),
          "f"(c_mushy), "f"(cs), "f"(c_fluid));
}
GM_DI void props3_km(const PropK& q, float c_mushy, float c_fluid, float T, float S1in, float thr, float kb, float cs,
                     float& k, float& m) {
    asm("{\n\t.reg .pred p1, p2, p3;\n\t"
        "setp.ge.f32 p2, %2, %5;\n\t"
        "setp.gt.f32 p3, %2, %6;\n\t"
        "setp.gt.f32 p1, %3, %4;\n\t"
        "selp.f32 %0, %7, %8, p1;\n\t"
        "selp.f32 %0, %9, %0, p2;\n\t"
        "selp.f32 %1, %10, %11, p3;\n\t"
        "selp.f32 %1, %12, %1, p2;\n\t}"
        : "=&f"(k), "=&f"(m)
        : "f"(T), "f"(S1in), "f"(thr), "f"(q.T_liq), "f"(q.T_sol), "f"(kb), "f"(q.k_powder), "f"(q.k_fluid),
          "f"(c_mushy), "f"(cs), "f"(c_fluid));
}

// Three-input maximum (one FMNMX3 on sm_100a); NaN operands are ignored like fmaxf does.
GM_DI float fmax3(float a, float b, float c) {
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}
struct ColdTag { static constexpr bool value = true; };
struct HotTag { static constexpr bool value = false; };

// ---- TMA plane ring (K1F_TMA) ------------------------------------------------------------------------------
// A field is described to the TMA unit as a ONE-dimensional tensor of nn floats (StepParams::tm): a 1-D box of 64
// elements needs no row stride; StepParams::tm2 is an overlapping 2-D view of the same array (see the ring in the
// kernel for why and how the two are used).
constexpr int K3_BOX = 96;  // elements per box row: 62 columns + up to 3 + 5 * 3 of alignment slack, 384-byte pitch
constexpr int K3_NS = 4;  // ring depth: planes l .. l+3 are resident / in flight while plane l is processed
GM_DI void mbar_init(unsigned bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
GM_DI void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
GM_DI void mbar_wait(unsigned bar, unsigned parity) {
    asm volatile("{\n\t.reg .pred P1;\n\t"
                 "WAIT_%=:\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
                 "@!P1 bra WAIT_%=;\n\t}" ::"r"(bar), "r"(parity)
                 : "memory");
}
template <unsigned OFF>
GM_DI float lds_f32(unsigned addr) {  // ordered with the mbarrier wait before it (both volatile)
    float v;
    asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"(addr), "n"(OFF));
    return v;
}
GM_DI bool elect_one() {  // one lane of the (converged) warp; lets the compiler keep the TMA operands in uniform registers
    unsigned pred;
    asm volatile("{\n\t.reg .pred P1;\n\t"
                 "elect.sync _|P1, 0xffffffff;\n\t"
                 "selp.u32 %0, 1, 0, P1;\n\t}"
                 : "=r"(pred));
    return pred != 0;
}
GM_DI void tma_load_row(unsigned dst, const void* tmap, int elem, unsigned bar) {
    asm volatile("cp.async.bulk.tensor.1d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2}], [%3];" ::"r"(dst),
                 "l"(tmap), "r"(elem), "r"(bar)
                 : "memory");
}

GM_DI void tma_load_box2(unsigned dst, const void* tmap, int x, unsigned bar) {  // rows y = 0 .. of the 2-D view at column x
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
                 "l"(tmap), "r"(x), "r"(0), "r"(bar)
                 : "memory");
}

// Guarded byte-pair store (S2): half .y at q, half .x 30 nodes below it.
GM_DI void st2_u8(uint8_t* q, int fa, int fb, int a, int b) {
    asm volatile("{\n\t.reg .pred pa, pb;\n\t"
                 "setp.ne.s32 pa, %3, 0;\n\t"
                 "setp.ne.s32 pb, %4, 0;\n\t"
                 "@pa st.global.u8 [%0+-30], %1;\n\t"
                 "@pb st.global.u8 [%0], %2;\n\t}"
                 ::"l"(q), "r"(a), "r"(b), "r"(fa), "r"(fb)
                 : "memory");
}
static_assert(K1_TX == 30, "st2_u8 hard-codes the half distance");

// Unguarded pair load: half .y at q, half .x 30 columns (120 bytes) below it.
GM_DI f2 ld2u(const char* q) {
    const float* f = reinterpret_cast<const float*>(q);
    return mk2(__ldg(f - K1_TX), __ldg(f));
}

// Melt-time bookkeeping (cF:3568-3578) of a plane that holds molten nodes.  It runs only for the few warp-planes
// that intersect the melt pool (the plane loop only tracks the maximum temperature it loaded and votes once per
// plane), and only AFTER the warp's z march, when the carried state is dead and its registers are free.  S2 = (T0 >= T_liquidus) is recomputed from T0 for the nodes this warp owns (rowmask:
// bit r = loaded row r is owned here; qa / qb: this lane's halves are); accum / max_accum change only where the
// node is molten NOW (s2 = 0: reset = 0, max(0, max_accum) = max_accum for the non-negative times it holds,
// accum + 0 - 0 = accum), so S2_prev, accum and max_accum are touched only there.  o0..o5 = row byte offsets.
template <bool F_S2, bool F_ACC, int NP>
GM_DI void bookkeep_planes(const StepParams& p, const size_t* pl, const unsigned* rowmask, int qa, int qb, unsigned o0,
                           unsigned o1, unsigned o2, unsigned o3, unsigned o4, unsigned o5) {
    // The few warps over the melt pool are the tail of a single-wave kernel, so this path is written for latency:
    // NP planes at a time, and every load of a stage - of all NP planes - is issued before the first use (T0 from
    // cache, then S2_prev / accum / max_accum): two dependent memory round trips per NP planes.  rowmask[q] = 0
    // switches plane q off (an odd plane out).
    const unsigned o[6] = {o0, o1, o2, o3, o4, o5};
    bool da[NP][6], db[NP][6];
    float ta[NP][6], tb[NP][6];
#pragma unroll
    for (int q = 0; q < NP; ++q)
#pragma unroll
        for (int r = 0; r < 6; ++r) {
            const bool mine = (rowmask[q] >> r) & 1u;
            const size_t nb = pl[q] + (o[r] >> 2);
            ta[q][r] = (mine && qa) ? __ldg(p.T0 + nb - K1_TX) : 0.f;
            tb[q][r] = (mine && qb) ? __ldg(p.T0 + nb) : 0.f;
        }
#pragma unroll
    for (int q = 0; q < NP; ++q)
#pragma unroll
        for (int r = 0; r < 6; ++r) {
            da[q][r] = ta[q][r] >= p.pk.T_liq;  // T_liquidus > 0, so a node this warp does not own is never "molten"
            db[q][r] = tb[q][r] >= p.pk.T_liq;
        }
    if (F_ACC) {
        uint8_t pa[NP][6], pb[NP][6];
        float aca[NP][6], acb[NP][6], mxa[NP][6], mxb[NP][6];
#pragma unroll
        for (int q = 0; q < NP; ++q)
#pragma unroll
            for (int r = 0; r < 6; ++r) {
                const size_t nb = pl[q] + (o[r] >> 2), na = nb - K1_TX;
                pa[q][r] = da[q][r] ? p.S2prev[na] : (uint8_t)1;   // S2_prev is read before S2 is written (in place)
                pb[q][r] = db[q][r] ? p.S2prev[nb] : (uint8_t)1;
                aca[q][r] = da[q][r] ? p.accum[na] : 0.f;
                acb[q][r] = db[q][r] ? p.accum[nb] : 0.f;
                mxa[q][r] = da[q][r] ? p.maxacc[na] : 0.f;
                mxb[q][r] = db[q][r] ? p.maxacc[nb] : 0.f;
            }
#pragma unroll
        for (int q = 0; q < NP; ++q)
#pragma unroll
            for (int r = 0; r < 6; ++r) {
                const size_t nb = pl[q] + (o[r] >> 2), na = nb - K1_TX;
                // reset = accum when the node has just melted
                const float ra = pa[q][r] ? 0.f : aca[q][r], rb = pb[q][r] ? 0.f : acb[q][r];
                if (da[q][r]) {
                    p.maxacc[na] = fmaxf(ra, mxa[q][r]);
                    p.accum[na] = aca[q][r] + p.dt - ra;
                }
                if (db[q][r]) {
                    p.maxacc[nb] = fmaxf(rb, mxb[q][r]);
                    p.accum[nb] = acb[q][r] + p.dt - rb;
                }
            }
    }
    if (F_S2) {
#pragma unroll
        for (int q = 0; q < NP; ++q)
#pragma unroll
            for (int r = 0; r < 6; ++r) {
                const bool mine = (rowmask[q] >> r) & 1u;
                const size_t nb = pl[q] + (o[r] >> 2);
                if (mine && qa) p.S2out[nb - K1_TX] = da[q][r] ? 1 : 0;
                if (mine && qb) p.S2out[nb] = db[q][r] ? 1 : 0;
            }
    }
}


template <int RY>
struct K3State {  // x-staged fields of one plane (loaded rows) + T of the owned rows
    f2 Xs[RY + 2], Xd[RY + 2], kx[RY + 2], mx[RY + 2], T[RY];
};
template <int RY>
struct K3Raw {    // prefetched raw plane
    f2 Tr[RY + 2], Sr[RY + 2];
};

// FEAT: K1F_RHS, K1F_SRC, K1F_FLUX, K1F_S1OUT, K1F_CLAMP, K1F_NSUB (whole planes), K1F_BCCONST, K1F_PEER;
// K1F_SKIP is implied when K1F_BCCONST is absent.  Requires nx >= 62, ny >= RY + 2, nz_active >= 2.
template <int RY, int FEAT, int MINB = 1>
__global__ void __launch_bounds__(32, MINB) level_step_v3(const __grid_constant__ StepParams p) {
    constexpr int NR = RY + 2;
    constexpr bool F_RHS = (FEAT & K1F_RHS) != 0, F_SRC = (FEAT & K1F_SRC) != 0, F_FLUX = (FEAT & K1F_FLUX) != 0;
    constexpr bool F_S1 = (FEAT & K1F_S1OUT) != 0, F_CLAMP = (FEAT & K1F_CLAMP) != 0, F_NSUB = (FEAT & K1F_NSUB) != 0;
    constexpr bool F_S2 = (FEAT & K1F_S2OUT) != 0, F_ACC = (FEAT & K1F_ACCUM) != 0;
    constexpr bool F_PEER = (FEAT & K1F_PEER) != 0, F_PF = (FEAT & K1F_PF) != 0, F_INPLACE = (FEAT & K1F_S1INPLACE) != 0;
    constexpr bool F_TMA = (FEAT & K1F_TMA) != 0;
    // the cold-plane path needs the whole plane before its first row, i.e. the TMA ring (see plane_is_cold); in the
    // corrector substeps a plane that took it also needs no vote on the liquidus for its melt-time bookkeeping
    constexpr bool USE_COLD = F_TMA;
    constexpr int NS = K3_NS;
    const int lane = threadIdx.x;
    const int nx = p.nx, ny = p.ny, nz = p.nz, nzl = p.nzl;
    const int c0 = min((int)blockIdx.x * (2 * K1_TX), nx - (2 * K1_TX + 2));  // column of lane 0, half .x
    const int j0 = min(1 + (int)blockIdx.y * RY, ny - 1 - RY);                  // first owned row
    const int ia = c0 + lane, ib = ia + K1_TX;
    const int own = (lane >= 1 && lane <= K1_TX) ? 1 : 0;
    // the same flag through two expressions ptxas does not identify: with one predicate on both stores of a pair
    // it turns the predication into a divergent branch around the row's tail (BSSY / BSYNC per row)
    int owna = ((unsigned)(lane - 1) < (unsigned)K1_TX) ? 1 : 0, ownb = (int)((0x7FFFFFFEu >> lane) & 1u);
#ifdef GOMELT_K1_ABLATE  // timing-only ablations: GOMELT_K1_EXP bit 16 = no T stores, 32 = no S1 stores,
    if (p.exp & 16) owna = ownb = 0;  // 64 = every load hits two cached planes
#endif
    int sa = own | (ia == 0 ? 1 : 0), sb = own | (ib == nx - 1 ? 1 : 0);  // lanes that store S1 (faces too)
#ifdef GOMELT_K1_ABLATE
    if (p.exp & 32) sa = sb = 0;
#endif
    // S2 / melt-time bookkeeping (cF:3568-3578) is a read-modify-write of every node, faces included, so it has
    // exactly ONE owner per node: the owned lanes / rows that are not the overlap of a shifted tile / strip with
    // its predecessor, plus the face column / row of the outermost tiles
    const int x_new = (int)blockIdx.x * (2 * K1_TX) + 1, y_new = (int)blockIdx.y * RY + 1;  // first non-duplicate column / row
    const int qa = ((own && ia >= x_new) || ia == 0) ? 1 : 0, qb = ((own && ib >= x_new) || ib == nx - 1) ? 1 : 0;
    const int P = nx * ny;
    const int za = p.zbeg + blockIdx.z * p.zchunk;
    const int zb = min(p.zend, za + p.zchunk);  // this warp finalises node planes [za, zb)
    const int lfirst = max(za - 1, 0);
    const int llast = min(min(zb, nz - 1), nzl - 1);  // last plane that carries data
    unsigned off[NR];                                  // in-plane byte offset of half .y per loaded row
#pragma unroll
    for (int r = 0; r < NR; ++r) off[r] = 4u * (unsigned)((j0 - 1 + r) * nx + ib);
    const int nsub_planes = p.nsub_planes;

    const int opaque_zero = nx >> 31;
    auto settle = [&](float x) { return __int_as_float(__float_as_int(x) ^ opaque_zero); };
    f2 sfx = splat(0.f);
    float sfy[RY];
#pragma unroll
    for (int r = 0; r < RY; ++r) sfy[r] = 0.f;
    if (F_SRC) {
        sfx = mk2(__ldg(p.srcx + ia) * (p.scoef * p.n_inv_s), __ldg(p.srcx + ib) * (p.scoef * p.n_inv_s));
#pragma unroll
        for (int r = 0; r < RY; ++r) sfy[r] = settle(__ldg(p.srcy + j0 + r));
    }
    // Level-1 Dirichlet constants on the side faces (assignBCs order y-, y+, x-, x+: later wins on edges):
    // the face nodes are the halo lanes / rows of the outermost tiles; they are written once per plane.
    f2 Tt0[RY], Tt1[RY], myp[RY];
#pragma unroll
    for (int r = 0; r < RY; ++r) Tt0[r] = Tt1[r] = myp[r] = splat(0.f);

    auto load_plane = [&](int l, K3Raw<RY>& raw) {
#ifdef GOMELT_K1_ABLATE
        if (p.exp & 64) l = l & 1;
#endif
        const char* Tl = (const char*)(p.T0 + (size_t)l * P);
        const char* Sl = (const char*)(p.S1 + (size_t)l * P);
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            raw.Tr[r] = ld2u(Tl + off[r]);
            raw.Sr[r] = ld2u(Sl + off[r]);
        }
    };

    // K1F_TMA: the raw planes arrive through a ring of NS stages in shared memory, [stage][T|S][row][96 floats],
    // filled by the TMA unit three planes ahead of their use (one elected lane issues the plane's boxes against the
    // stage's mbarrier), instead of one plane ahead into 48 registers.  A box must start on a 16-byte boundary of the
    // field (the TMA unit faults otherwise) and a 2-D map wants a row stride that is a multiple of 16 bytes, which
    // the rows of an odd nx do not give.  So the field is described as rows of S = nx & ~3 elements that START
    // anywhere (an overlapping 2-D view of the flat array: element (x, y) is x + y S): box row r begins at
    // a = (first element of the tile's row 0, rounded down to a multiple of four) + r S, and the tile's row r begins
    // sh0 + r (nx & 3) elements into it (sh0 = the rounding remainder, warp-uniform per plane).  One 96 x 6 box per
    // field and plane.  Where that box would reach past the end of the array (the last rows of the last plane) the
    // same rows are fetched by six 1-D boxes per field, which the 1-D map clips at nn.
    constexpr int RW = K3_BOX;  // floats per ring row (384-byte pitch: every 1-D box destination is 128-byte aligned)
    __shared__ alignas(128) float s_ring[F_TMA ? NS : 1][2][F_TMA ? NR : 1][F_TMA ? RW : 1];
    __shared__ alignas(8) unsigned long long s_bar[F_TMA ? NS : 1];
    const int row0 = (j0 - 1) * nx + c0;  // in-plane element of the tile's first loaded row, column c0
    const int Sq = nx & ~3, dq = nx & 3;
    const int nn_all = P * nz;
    static_assert((NS & (NS - 1)) == 0, "stage = plane & (NS - 1)");
    // shared-window addresses and the running plane element are formed once (every cvta costs an S2UR + ULEA pair)
    // (laundered through an opaque move: the compiler otherwise rematerialises the conversion at every use)
    unsigned ring0 = F_TMA ? smem_u32(&s_ring[0][0][0][0]) : 0u, bar0 = F_TMA ? smem_u32(&s_bar[0]) : 0u;
    if (F_TMA) asm volatile("mov.b32 %0, %0;\n\tmov.b32 %1, %1;" : "+r"(ring0), "+r"(bar0));
    constexpr unsigned STAGE_B = 2u * NR * RW * 4u, FIELD_B = NR * RW * 4u;
    const int a_max = nn_all - (NR - 1) * Sq - RW;  // last box start whose six rows stay inside the array
    const bool use2d = p.tm2_ok != 0;
    const unsigned row_pitch = 4u * (unsigned)(RW + dq);  // bytes from a tile row to the next inside a stage
    // descriptor addresses (generic addresses of the __grid_constant__ parameter), formed once as well
    unsigned long long d2T = (unsigned long long)&p.tm2[0], d2S = (unsigned long long)&p.tm2[1];
    if (F_TMA) asm volatile("mov.b64 %0, %0;\n\tmov.b64 %1, %1;" : "+l"(d2T), "+l"(d2S));
    auto ring_issue = [&](int l) {
        if (l > llast) return;
        const unsigned s = (unsigned)(l - lfirst) & (NS - 1);
        // (no __syncwarp: the stage was last read a whole plane ago, and the votes since then are warp-wide)
        if (elect_one()) {
            const unsigned bar = bar0 + 8u * s, dst = ring0 + STAGE_B * s;
            mbar_expect_tx(bar, STAGE_B);
            const int a = (l * P + row0) & ~3;
            if (use2d && a <= a_max) {
                tma_load_box2(dst, (const void*)d2T, a, bar);
                tma_load_box2(dst + FIELD_B, (const void*)d2S, a, bar);
            } else {
#pragma unroll 1
                for (int r = 0; r < NR; ++r) {
                    tma_load_row(dst + RW * 4u * r, &p.tm[0], a + r * Sq, bar);
                    tma_load_row(dst + FIELD_B + RW * 4u * r, &p.tm[1], a + r * Sq, bar);
                }
            }
        }
    };
    auto ring_fetch = [&](int l, K3Raw<RY>& raw) {
        const unsigned k = (unsigned)(l - lfirst), s = k & (NS - 1);
        mbar_wait(bar0 + 8u * s, (k / NS) & 1u);
        const unsigned b = ring0 + STAGE_B * s + 4u * (unsigned)(((l * P + row0) & 3) + lane);
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            const unsigned q = b + (unsigned)r * row_pitch;
            raw.Tr[r] = mk2(lds_f32<0>(q), lds_f32<4 * K1_TX>(q));
            raw.Sr[r] = mk2(lds_f32<FIELD_B>(q), lds_f32<FIELD_B + 4 * K1_TX>(q));
        }
    };

    // L2 prefetch of plane l: one instruction per loaded row and field, lane i touching 8 bytes further than lane
    // i-1, so that the 32 lanes cover the 256 bytes both halves of the row segment span
    const unsigned pf_lane = 8u * (unsigned)lane - 4u * (unsigned)K1_TX;
    auto l2_prefetch = [&](int l) {
        const char* Tl = (const char*)(p.T0 + (size_t)l * P);
        const char* Sl = (const char*)(p.S1 + (size_t)l * P);
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            asm volatile("prefetch.global.L2 [%0];" ::"l"(Tl + off[r] - 4u * (unsigned)lane + pf_lane));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(Sl + off[r] - 4u * (unsigned)lane + pf_lane));
        }
    };

    // ---- S2 / melt-time bookkeeping of plane l, after its rows (see bookkeep_plane) -------------------------------
    f2 tmax = splat(0.f);  // running maximum of the temperatures loaded for the current plane
    unsigned rows_mine = 0;  // loaded rows whose nodes this strip owns for the bookkeeping (plane-invariant)
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        const int j = j0 - 1 + r;
        if ((r >= 1 && r <= RY) ? (j >= y_new) : (r == 0 ? j == 0 : j == ny - 1)) rows_mine |= 1u << r;
    }
    // After the rows of plane l (branch-free): one vote; a cold plane (nothing molten among the loaded nodes) gets
    // S2 = 0 on its owned nodes by guarded byte stores, a hot one is remembered in `hotmask` (bit l - lfirst; the
    // launcher keeps bookkeeping chunks within 64 planes) and handled by flush_bookkeeping() after the march: the
    // melt-pool warps are the tail of a single-wave kernel, and there the registers of the carried state are free,
    // so every load of a hot plane can be in flight at once.
    unsigned long long hotmask = 0;
    // known_cold: the plane went through the cold body (plane_is_cold: nothing above the solidus, so nothing molten)
    auto bookkeep = [&](int l, bool known_cold) {
        if (!(F_S2 || F_ACC)) return;
        const int mine = (l >= za && l < zb) ? 1 : 0;  // a chunk's halo plane belongs to the neighbouring chunk
        int cold = mine;
        if (!known_cold) {
            const float hot = fmaxf(tmax.v.x, tmax.v.y);
            tmax = splat(0.f);
            cold = __any_sync(0xffffffffu, hot >= p.pk.T_liq) ? 0 : mine;
        }
        hotmask |= (unsigned long long)(mine & ~cold & 1) << (l - lfirst);
        if (F_S2) {
            const size_t pl = (size_t)l * P;
#pragma unroll
            for (int r = 0; r < NR; ++r) {
                const int g = (int)((rows_mine >> r) & 1u) & cold;
                st2_u8(p.S2out + pl + (off[r] >> 2), qa & g, qb & g, 0, 0);
            }
        }
    };
    auto flush_bookkeeping = [&]() {
        if (!(F_S2 || F_ACC)) return;
        static_assert(NR == 6, "bookkeep_planes takes the six row offsets of RY = 4");
        unsigned long long m = hotmask;  // warp-uniform
        hotmask = 0;
        if (p.bkq && m) {
            // With a work queue the hot planes are handed to bookkeep_queue_kernel, which follows on the stream: the few
            // warps over the melt pool are the tail of a single-wave launch, and their bookkeeping - two dependent memory
            // round trips per pair of planes, serial in this warp - is spread over the whole GPU instead.
            const unsigned n = (unsigned)__popcll(m);
            unsigned base = 0;
            if (lane == 0) base = atomicAdd(p.bkq, n);
            base = __shfl_sync(0xffffffffu, base, 0);
            if (base + n <= p.bkq_cap) {
                if (lane == 0) {
                    unsigned* q = p.bkq + 2 + 2 * (size_t)base;
                    const unsigned tile = (unsigned)blockIdx.x | ((unsigned)blockIdx.y << 16);
                    while (m) {
                        const int b = __ffsll((long long)m) - 1;
                        m &= m - 1;
                        q[0] = tile;
                        q[1] = (unsigned)(lfirst + b);
                        q += 2;
                    }
                }
                return;
            }
            // queue full: this warp keeps its planes; the slots it reserved below the capacity are marked empty
            if (lane == 0)
                for (unsigned e = base; e < p.bkq_cap; ++e) p.bkq[3 + 2 * (size_t)e] = 0xffffffffu;
        }
#pragma unroll 1
        while (m) {  // two hot planes per pass
            size_t pl[2];
            unsigned rm[2];
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int b = m ? __ffsll((long long)m) - 1 : 0;
                rm[q] = m ? rows_mine : 0u;
                m &= m - 1;  // (0 stays 0)
                pl[q] = (size_t)(lfirst + b) * P;
            }
            bookkeep_planes<F_S2, F_ACC, 2>(p, pl, rm, qa, qb, off[0], off[1], off[2], off[3], off[4], off[5]);
        }
    };

    // ---- node state of loaded row r (+ S1 output) and its x stage ------------------------------------
    // `cold` (ColdTag): the warp has voted that no node it loaded of this plane is above the solidus / at the liquidus
    // (plane_is_cold below), so S2 = S3 = 0 there and computeStateProperties cF:2567-2614 collapses to
    // k = S1' ? k_solid(T) : k_powder, rho cp = rho cp_solid(T), S1' = S1 > thr: two selects per node instead of
    // seven.  The values are the ones the general selects produce, bit for bit.
    // In place (F_INPLACE) a node's state is stored only when it changed, which it rarely does: on a cold plane the
    // rows only OR the bit difference of S1' and S1 into `chg`, and the plane's tail (flush_state) stores the changed
    // nodes after one vote - no compare / guard / predicated-off store per node in the plane body.
    unsigned chg = 0;
    auto row_a = [&](auto cold, char* so, int r, float thr, f2 T, f2 S, f2& xs, f2& xd, f2& kxr, f2& mxr) {
        constexpr bool COLD = decltype(cold)::value;
        // masses carry the 1 / (cdt * lambda'[2]) normalisation (StepParams)
        const f2 kb = fma2(splat(p.pk.k_a1), T, splat(p.pk.k_a0)), cs = fma2(splat(p.n_ca1), T, splat(p.n_ca0));
        f2 kn, mn, s1f = splat(0.f);
        if (COLD) {
            const bool pa = S.v.x > thr, pb = S.v.y > thr;
            kn = mk2(pa ? kb.v.x : p.pk.k_powder, pb ? kb.v.y : p.pk.k_powder);
            mn = cs;
            if (F_S1) s1f = mk2(pa ? 1.f : 0.f, pb ? 1.f : 0.f);
        } else if (F_S1) {
            // S1' is node-local, so every loaded node (halo rows, halo planes of a chunk, face lanes) may be
            // written: its owner writes the same value.  This is what covers the face nodes without a guard.
            props3(p.pk, p.n_cmushy, p.n_cfluid, T.v.x, S.v.x, thr, kb.v.x, cs.v.x, kn.v.x, mn.v.x, s1f.v.x);
            props3(p.pk, p.n_cmushy, p.n_cfluid, T.v.y, S.v.y, thr, kb.v.y, cs.v.y, kn.v.y, mn.v.y, s1f.v.y);
        } else {
            props3_km(p.pk, p.n_cmushy, p.n_cfluid, T.v.x, S.v.x, thr, kb.v.x, cs.v.x, kn.v.x, mn.v.x);
            props3_km(p.pk, p.n_cmushy, p.n_cfluid, T.v.y, S.v.y, thr, kb.v.y, cs.v.y, kn.v.y, mn.v.y);
        }
        if (F_S1) {
            if (F_INPLACE && COLD)  // cold planes: deferred (flush_state)
                chg |= (__float_as_uint(s1f.v.x) ^ __float_as_uint(S.v.x)) | (__float_as_uint(s1f.v.y) ^ __float_as_uint(S.v.y));
            else if (F_INPLACE)     // in place (the steppers): S1' == S1 for all but the nodes that melt in this substep
                st2(so + off[r], (s1f.v.x != S.v.x) ? sa : 0, (s1f.v.y != S.v.y) ? sb : 0, s1f);
            else
                st2(so + off[r], sa, sb, s1f);
        }
        // the corrector substeps (subcycleL3_Part2) track the plane's maximum: bookkeep() votes on it after the rows
        if ((F_S2 || F_ACC) && !COLD) tmax = mk2(fmaxf(tmax.v.x, T.v.x), fmaxf(tmax.v.y, T.v.y));
        const f2 Tr = shdn(T), kr = shdn(kn), mr = shdn(mn);
        xs = Tr + T;
        xd = Tr - T;
        kxr = kr + kn;
        mxr = mr + mn;
    };
    // In-place state of a cold plane l: nothing to do unless some loaded node's S1' differs from its S1 (a powder node
    // of a substrate plane, a fractional Level-1 state).  The rare path re-reads the plane (cache hits; another warp
    // may already have stored the same S1' for a shared halo node, which S1' maps to itself) and stores the changed
    // nodes with the guards of the direct form.
    auto flush_state = [&](int l) {
        if (!(F_S1 && F_INPLACE && USE_COLD)) return;
        const bool any = __any_sync(0xffffffffu, chg != 0);
        chg = 0;
        if (!any) return;
        const float thr = (F_NSUB && l < nsub_planes) ? -1.0f : 0.499f;
        const char* Tl = (const char*)(p.T0 + (size_t)l * P);
        const char* Sl = (const char*)(p.S1 + (size_t)l * P);
        char* so = (char*)(p.S1out + (size_t)l * P);
#pragma unroll 1
        for (int r = 0; r < NR; ++r) {
            const unsigned o = 4u * (unsigned)((j0 - 1 + r) * nx + ib);
            const float* tf = reinterpret_cast<const float*>(Tl + o);
            const volatile float* sf = reinterpret_cast<const volatile float*>(Sl + o);
            const f2 T = mk2(__ldg(tf - K1_TX), __ldg(tf)), S = mk2(sf[-K1_TX], sf[0]);
            f2 s1f;
            s1f.v.x = (S.v.x > thr || T.v.x >= p.pk.T_liq) ? 1.f : 0.f;
            s1f.v.y = (S.v.y > thr || T.v.y >= p.pk.T_liq) ? 1.f : 0.f;
            st2(so + o, (s1f.v.x != S.v.x) ? sa : 0, (s1f.v.y != S.v.y) ? sb : 0, s1f);
        }
    };
    // One vote per plane on the temperatures the warp loaded for it (6 three-input maxima + 1).
    auto plane_is_cold = [&](const K3Raw<RY>& raw) -> bool {
        static_assert(NR == 6, "six loaded rows");
        const float m0 = fmax3(raw.Tr[0].v.x, raw.Tr[0].v.y, raw.Tr[1].v.x), m1 = fmax3(raw.Tr[1].v.y, raw.Tr[2].v.x, raw.Tr[2].v.y);
        const float m2 = fmax3(raw.Tr[3].v.x, raw.Tr[3].v.y, raw.Tr[4].v.x), m3 = fmax3(raw.Tr[4].v.y, raw.Tr[5].v.x, raw.Tr[5].v.y);
        const float m = fmaxf(fmax3(m0, m1, m2), m3);
        return !__any_sync(0xffffffffu, (m > p.pk.T_sol) || (m >= p.pk.T_liq));
    };
    auto plane_thr = [&](int l) -> float { return (F_NSUB && l < nsub_planes) ? -1.0f : 0.499f; };

    auto first_plane = [&](int l, const K3Raw<RY>& raw, K3State<RY>& st) {
        char* so = F_S1 ? (char*)(p.S1out + (size_t)l * P) : nullptr;
        const float thr = plane_thr(l);
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            row_a(HotTag{}, so, r, thr, raw.Tr[r], raw.Sr[r], st.Xs[r], st.Xd[r], st.kx[r], st.mx[r]);
            if (r >= 1 && r <= RY) st.T[r - 1] = raw.Tr[r];
        }
        bookkeep(l, 0);
    };

    // ---- write one finalised owned row of plane f (side faces are not owned; see the header) ---------
    auto store_row = [&](int f, char* out, int r, f2 Tn) {
        if (F_CLAMP) Tn = mk2(fmaxf(p.pk.T_amb, Tn.v.x), fmaxf(p.pk.T_amb, Tn.v.y));
        st2(out + off[r + 1], owna, ownb, Tn);
        if (F_PEER) {  // halo exchange fused into the step: plain stores to peer-mapped memory
            if (f == p.zbeg && p.peer_lo) st2((char*)p.peer_lo + off[r + 1], owna, ownb, Tn);
            if (f == p.zend - 1 && p.peer_hi) st2((char*)p.peer_hi + off[r + 1], owna, ownb, Tn);
        }
    };

    // use_fl is a literal at every call site (the lambdas are inlined): only the top plane carries a flux
    // rq: the plane's rhs rows prefetched a plane ahead (TMA variant), or nullptr: loaded here
    auto final_row = [&](int f, char* out, const char* rhs, const f2* rq, int r, f2 sz, f2 Tf, f2 z0, f2 z1, f2 mz,
                         bool use_fl, f2 fl) {
        const f2 nKT = (z1 - z0) - shup(z0 + z1);  // minus the stiffness action on this node
        const f2 mnode = mz + shup(mz);
        const f2 rc = mk2(rcp_approx(mnode.v.x), rcp_approx(mnode.v.y));
        // normalised form: T_new = T + (rr / s - KT / s) / (mnode / (cdt s)); rc is the reciprocal of the latter
        if (F_RHS || F_SRC || (F_FLUX && use_fl)) {
            f2 rr = nKT;
            if (F_RHS) rr = fma2(rq ? rq[r] : ld2u(rhs + off[r + 1]), splat(p.n_inv_s), rr);
            if (F_SRC) rr = fma2(sz, splat(sfy[r]), rr);
            if (F_FLUX && use_fl) rr = rr + fl;
            store_row(f, out, r, fma2(rr, rc, Tf));
        } else {
            store_row(f, out, r, fma2(nKT, rc, Tf));
        }
    };

    const f2 A01 = splat(p.lamA[0]), A10 = splat(p.lamA[1]), A11 = splat(p.lamA[2]);
    const f2 B01 = splat(p.lamB[0]), B10 = splat(p.lamB[1]), B11 = splat(p.lamB[2]);

    // ---- plane l: node state, x stage, then the element layer (l-1, l) row by row; finalises plane l-1
    //      when do_final.  pv = state of plane l-1, cu <- state of plane l.
    auto step_plane = [&](auto cold, int l, const K3Raw<RY>& raw, const K3State<RY>& pv, K3State<RY>& cu, bool do_final,
                          f2 sz, const f2* rq) {
        const int f = l - 1;
        const size_t pl = (size_t)l * P;
        char* so = F_S1 ? (char*)(p.S1out + pl) : nullptr;
        const float thr = plane_thr(l);
        // A plane that must not be finalised here (a chunk's lower halo plane; the Dirichlet bottom plane 0) is
        // computed like any other and stored to plane l instead, where the next step overwrites it (same
        // thread, same addresses, program order) - no per-plane guard in the loop.
        char* out = (char*)(p.Tout + (do_final ? pl - P : pl));
        const char* rhs = F_RHS ? (const char*)(p.rhs + (pl - P)) : nullptr;
        f2 c00, c01, c10, c11, cm;                            // carries of the previous element row
        f2 zl00, zl01, zl10, zl11, kzl, mzl, bl01, bl10, bl11;  // z-staged fields of the lower loaded row
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            f2 xs, xd, kxr, mxr;
            row_a(cold, so, r, thr, raw.Tr[r], raw.Sr[r], xs, xd, kxr, mxr);
            cu.Xs[r] = xs; cu.Xd[r] = xd; cu.kx[r] = kxr; cu.mx[r] = mxr;
            if (r >= 1 && r <= RY) cu.T[r - 1] = raw.Tr[r];
            // z stage of the analysis
            const f2 zu00 = xs + pv.Xs[r], zu01 = xs - pv.Xs[r];
            const f2 zu10 = xd + pv.Xd[r], zu11 = xd - pv.Xd[r];
            const f2 kzu = kxr + pv.kx[r], mzu = mxr + pv.mx[r];
            const f2 bu01 = B01 * zu01, bu10 = B10 * zu10, bu11 = B11 * zu11;
            if (r >= 1) {
                const int e = r - 1;  // element row between loaded rows e and e+1
                const f2 k8 = kzu + kzl;
                const f2 m8 = mzu + mzl;
                const f2 q00 = (zu00 - zl00) * k8;  // (sx,sz) = (0,0): only the sy = 1 mode acts (lambda = 1 after normalisation)
                if (e >= 1) {  // lower node row of this element row = loaded row e = owned row e-1
                    const f2 R00 = c00 - q00;
                    const f2 R01 = fma2(k8, fma2(A01, zl01, bu01), c01);
                    const f2 R10 = fma2(k8, fma2(A10, zl10, bu10), c10);
                    const f2 R11 = fma2(k8, fma2(A11, zl11, bu11), c11);
                    // z stage of the synthesis: bottom (plane l-1) and top (plane l) parts
                    const f2 z0 = Tt0[e - 1] + (R00 - R01);
                    const f2 z1 = Tt1[e - 1] + (R10 - R11);
                    Tt0[e - 1] = R00 + R01;
                    Tt1[e - 1] = R10 + R11;
                    const f2 my = cm + m8;
                    const f2 mz = myp[e - 1] + my;
                    myp[e - 1] = my;
                    final_row(f, out, rhs, rq, e - 1, sz, pv.T[e - 1], z0, z1, mz, false, splat(0.f));
                }
                if (e < RY) {
                    c00 = q00;
                    c01 = fma2(A01, zu01, bl01) * k8;
                    c10 = fma2(A10, zu10, bl10) * k8;
                    c11 = fma2(A11, zu11, bl11) * k8;
                    cm = m8;
                }
            }
            zl00 = zu00; zl01 = zu01; zl10 = zu10; zl11 = zu11; kzl = kzu; mzl = mzu;
            bl01 = bu01; bl10 = bu10; bl11 = bu11;
        }
        if (decltype(cold)::value) flush_state(l);
        bookkeep(l, decltype(cold)::value);
    };
    const bool allow_cold = !(p.flags & GOMELT_STEP_NO_COLD_PLANES);
    auto run_plane = [&](int l, const K3Raw<RY>& raw, const K3State<RY>& pv, K3State<RY>& cu, bool do_final, f2 sz,
                         const f2* rq) {
        if constexpr (USE_COLD) {
            if (allow_cold && plane_is_cold(raw)) {
                step_plane(ColdTag{}, l, raw, pv, cu, do_final, sz, rq);
                return;
            }
        }
        step_plane(HotTag{}, l, raw, pv, cu, do_final, sz, rq);
    };
    // rhs rows of the owned nodes of plane f, fetched a plane before final_row needs them (TMA variant: T and S1 no
    // longer come through the LSU, so these are the only long-latency loads left in the plane body)
    auto load_rhs = [&](int f, f2* rq) {
        if (!F_RHS) return;
        const char* q = (const char*)(p.rhs + (size_t)f * P);
#pragma unroll
        for (int r = 0; r < RY; ++r) rq[r] = ld2u(q + off[r + 1]);
    };

    // ---- the chunk's source z-factors live in shared memory (see v2) ---------------------------------
    __shared__ float s_srcz[F_SRC ? K1_SRCZ_MAX + 2 : 1];
    if (F_SRC) {
        for (int i = lane; i <= llast - lfirst; i += 32) s_srcz[i] = __ldg(p.srcz + lfirst + i);
        __syncwarp();
    }
    auto srcz_at = [&](int f) -> float { return F_SRC ? s_srcz[f - lfirst] : 0.f; };

    // ---- computeConvRadBC cF:2207-2301 fused (see v2): every element of the tile exists here ----------
    auto fused_top_flux = [&](const K3Raw<RY>& raw, f2* fl) {
        // staging: 6.5 KB of its own, or - with the TMA ring - the ring itself: the top plane is the last one, every
        // stage has been read by then and nothing is in flight
        __shared__ float2 sT_own[F_TMA ? 1 : NR][F_TMA ? 1 : 32];
        __shared__ float2 sA_own[F_TMA ? 1 : RY + 1][F_TMA ? 1 : 4][F_TMA ? 1 : 32];
        static_assert(!F_TMA || sizeof(float2) * 32 * (NR + 4 * (RY + 1)) <= sizeof(float) * 2 * 2 * NR * K3_BOX, "ring too small");
        float2(*sT)[32] = F_TMA ? reinterpret_cast<float2(*)[32]>(&s_ring[0][0][0][0]) : reinterpret_cast<float2(*)[32]>(&sT_own[0][0]);
        float2(*sA)[4][32] = F_TMA ? reinterpret_cast<float2(*)[4][32]>(&s_ring[0][0][0][0] + 2 * 32 * NR)
                                   : reinterpret_cast<float2(*)[4][32]>(&sA_own[0][0][0]);
        __syncwarp();
#pragma unroll
        for (int r = 0; r < NR; ++r) sT[r][lane] = raw.Tr[r].v;
        __syncwarp();
        const float* sTf = reinterpret_cast<const float*>(&sT[0][0]);
        float* sAf = reinterpret_cast<float*>(&sA[0][0][0]);
        const int lane1 = min(lane + 1, 31);
        const float g = 0.57735026918962576f;
        const float NA = 0.25f * (1.f + g) * (1.f + g), NB = 0.25f * (1.f + g) * (1.f - g),
                    NC = 0.25f * (1.f - g) * (1.f - g);
        const float wq = p.n_wq;  // hx hy / 4 with the load normalisation
#pragma unroll 1
        for (int it = 0; it < 2 * (RY + 1); ++it) {
            const int e = it >> 1, h = it & 1;  // element row between loaded rows e, e + 1; half of the pair
            const float T0n = sTf[(e * 32 + lane) * 2 + h], T1n = sTf[(e * 32 + lane1) * 2 + h];
            const float T2n = sTf[((e + 1) * 32 + lane1) * 2 + h], T3n = sTf[((e + 1) * 32 + lane) * 2 + h];
            const float q0 = flux_fast(p.fk, fmaf(NB, T3n, fmaf(NC, T2n, fmaf(NB, T1n, NA * T0n)))) * wq;
            const float q1 = flux_fast(p.fk, fmaf(NC, T3n, fmaf(NB, T2n, fmaf(NA, T1n, NB * T0n)))) * wq;
            const float q2 = flux_fast(p.fk, fmaf(NB, T3n, fmaf(NA, T2n, fmaf(NB, T1n, NC * T0n)))) * wq;
            const float q3 = flux_fast(p.fk, fmaf(NA, T3n, fmaf(NB, T2n, fmaf(NC, T1n, NB * T0n)))) * wq;
            sAf[((e * 4 + 0) * 32 + lane) * 2 + h] = fmaf(NB, q3, fmaf(NC, q2, fmaf(NB, q1, NA * q0)));
            sAf[((e * 4 + 1) * 32 + lane) * 2 + h] = fmaf(NC, q3, fmaf(NB, q2, fmaf(NA, q1, NB * q0)));
            sAf[((e * 4 + 2) * 32 + lane) * 2 + h] = fmaf(NB, q3, fmaf(NA, q2, fmaf(NB, q1, NC * q0)));
            sAf[((e * 4 + 3) * 32 + lane) * 2 + h] = fmaf(NA, q3, fmaf(NB, q2, fmaf(NC, q1, NB * q0)));
        }
        __syncwarp();
        const int lm = max(lane - 1, 0);
#pragma unroll
        for (int r = 0; r < RY; ++r)  // owned row r = loaded row r + 1: upper node of element row r, lower of r + 1
            fl[r] = ((f2{sA[r][2][lm]} + f2{sA[r][3][lane]}) + f2{sA[r + 1][1][lm]}) + f2{sA[r + 1][0][lane]};
    };

    // ---- last data plane of a chunk top: no layer above, its action is (Tt0, Tt1, myp) --------------
    auto last_plane = [&](int f, const f2* Tf, const K3Raw<RY>& raw, const f2* rq = nullptr) {
        f2 sz = splat(0.f);
        if (F_SRC) sz = sfx * splat(srcz_at(f));
        f2 fl[RY];
#pragma unroll
        for (int r = 0; r < RY; ++r) fl[r] = splat(0.f);
        if (F_FLUX && f == nzl - 1) fused_top_flux(raw, fl);  // warp-uniform
        char* out = (char*)(p.Tout + (size_t)f * P);
        const char* rhs = F_RHS ? (const char*)(p.rhs + (size_t)f * P) : nullptr;
#pragma unroll
        for (int r = 0; r < RY; ++r) final_row(f, out, rhs, rq, r, sz, Tf[r], Tt0[r], Tt1[r], myp[r], true, fl[r]);
    };

    int fdone = za;  // planes [za, fdone) are finalised
    if (lfirst <= llast) {
        K3State<RY> stA, stB;
        if constexpr (F_TMA) {
            if (lane == 0) {
#pragma unroll
                for (int s = 0; s < NS; ++s) mbar_init(bar0 + 8u * s, 1);
                asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            }
            __syncwarp();
            K3Raw<RY> raw;
#pragma unroll
            for (int q = 0; q < NS; ++q) ring_issue(lfirst + q);
            f2 rqA[RY], rqB[RY];  // rhs rows of the plane finalised by the next step (ping-pong like the state)
#pragma unroll
            for (int r = 0; r < RY; ++r) rqA[r] = rqB[r] = splat(0.f);
            ring_fetch(lfirst, raw);
            load_rhs(lfirst, rqA);
            first_plane(lfirst, raw, stA);
            for (int l = lfirst + 1; l <= llast; l += 2) {
                ring_issue(l + NS - 1);  // into the stage plane l - 1 has just left
                ring_fetch(l, raw);
                load_rhs(l, rqB);
                run_plane(l, raw, stA, stB, l - 1 >= max(za, 1), sfx * splat(srcz_at(l - 1)), F_RHS ? rqA : nullptr);
                if (l + 1 > llast) break;
                ring_issue(l + NS);
                ring_fetch(l + 1, raw);
                load_rhs(l + 1, rqA);
                run_plane(l + 1, raw, stB, stA, l >= max(za, 1), sfx * splat(srcz_at(l)), F_RHS ? rqB : nullptr);
            }
            if (llast >= za && llast < zb) {  // raw holds plane llast
                if (((llast - lfirst) & 1) != 0) last_plane(llast, stB.T, raw, F_RHS ? rqB : nullptr);
                else last_plane(llast, stA.T, raw, F_RHS ? rqA : nullptr);
            }
        } else {
            K3Raw<RY> rawA, rawB;
            load_plane(lfirst, rawA);
            if (lfirst + 1 <= llast) load_plane(lfirst + 1, rawB);
            first_plane(lfirst, rawA, stA);
            // main loop, unrolled by two so the carried state ping-pongs: (stA, rawB) -> stB, (stB, rawA) -> stA
            // plane 0 is the Dirichlet bottom face: never finalised (l - 1 >= 1)
            for (int l = lfirst + 1; l <= llast; l += 2) {
                if (l + 1 <= llast) load_plane(l + 1, rawA);
                if (F_PF && l + 3 <= llast) l2_prefetch(l + 3);
                run_plane(l, rawB, stA, stB, l - 1 >= max(za, 1), sfx * splat(srcz_at(l - 1)), nullptr);
                if (l + 1 > llast) break;
                if (l + 2 <= llast) load_plane(l + 2, rawB);
                if (F_PF && l + 4 <= llast) l2_prefetch(l + 4);
                run_plane(l + 1, rawA, stB, stA, l >= max(za, 1), sfx * splat(srcz_at(l)), nullptr);
            }
            if (llast >= za && llast < zb) {
                if (((llast - lfirst) & 1) != 0) last_plane(llast, stB.T, rawB);  // parity of the plane held in stB
                else last_plane(llast, stA.T, rawA);
            }
        }
        fdone = max(za, llast + 1);
        flush_bookkeeping();
    }
    // planes >= nz_active (substitute_Tbar cF:2183); the side faces belong to the face pass
    for (int f = max(fdone, 1); f < zb; ++f) {
        char* out = (char*)(p.Tout + (size_t)f * P);
#pragma unroll
        for (int r = 0; r < RY; ++r) st2(out + off[r + 1], owna, ownb, splat(p.pk.T_amb));
    }
}

// Melt-time bookkeeping of the hot planes that level_step_v3 queued (StepParams::bkq: word 0 = number of entries, word 1
// = blocks of this kernel that have finished, then (tile, plane) pairs): one warp per entry, with the tile geometry formed exactly as the step kernel forms it
// (one owner per node: the owned lanes / rows that are not the overlap of a shifted tile, plus the face column / row of the
// outermost tiles).  Entries beyond the capacity were handled by the step kernel itself.
template <int RY, bool F_S2, bool F_ACC>
__global__ void __launch_bounds__(128) bookkeep_queue_kernel(const __grid_constant__ StepParams p) {
    static_assert(RY == 4, "bookkeep_planes takes the six row offsets of RY = 4");
    constexpr int NR = RY + 2;
    const unsigned n = min(p.bkq[0], p.bkq_cap);
    const int lane = threadIdx.x & 31;
    const int nx = p.nx, ny = p.ny;
    const size_t P = (size_t)nx * ny;
    for (unsigned e = blockIdx.x * 4u + (threadIdx.x >> 5); e < n; e += gridDim.x * 4u) {
        const unsigned tile = p.bkq[2 + 2 * (size_t)e], l = p.bkq[3 + 2 * (size_t)e];
        if (l == 0xffffffffu) continue;  // reserved by a warp that found the queue full and kept its planes
        const int bx = (int)(tile & 0xffffu), by = (int)(tile >> 16);
        const int c0 = min(bx * (2 * K1_TX), nx - (2 * K1_TX + 2));
        const int j0 = min(1 + by * RY, ny - 1 - RY);
        const int ia = c0 + lane, ib = ia + K1_TX;
        const int own = (lane >= 1 && lane <= K1_TX) ? 1 : 0;
        const int x_new = bx * (2 * K1_TX) + 1, y_new = by * RY + 1;
        const int qa = ((own && ia >= x_new) || ia == 0) ? 1 : 0, qb = ((own && ib >= x_new) || ib == nx - 1) ? 1 : 0;
        unsigned off[NR], rows_mine = 0;
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            const int j = j0 - 1 + r;
            off[r] = 4u * (unsigned)(j * nx + ib);
            if ((r >= 1 && r <= RY) ? (j >= y_new) : (r == 0 ? j == 0 : j == ny - 1)) rows_mine |= 1u << r;
        }
        const size_t pl[1] = {(size_t)l * P};
        const unsigned rm[1] = {rows_mine};
        bookkeep_planes<F_S2, F_ACC, 1>(p, pl, rm, qa, qb, off[0], off[1], off[2], off[3], off[4], off[5]);
    }
    // the last block to finish leaves the header zeroed for the next sweep (every block has read the count by then)
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(p.bkq + 1, 1u) == gridDim.x - 1) {
            p.bkq[0] = 0u;
            p.bkq[1] = 0u;
        }
    }
}

// Level-1 Dirichlet constants on the five faces of T_out (assignBCs cF:1568-1595, order y-, y+, x-, x+, z-:
// the later face wins on shared edges), planes [zbeg, zend), and - for a z-slab rank - on the same nodes of the
// neighbours' ghost planes (peer_lo <- plane zbeg, peer_hi <- plane zend-1).  One thread per face node:
// ~2 (nx + ny) nz + nx ny nodes, a fraction of a percent of a sweep.
__global__ void face_const_kernel(float* __restrict__ T, int nx, int ny, int nz, int zbeg, int zend, float b0, float b1,
                                  float b2, float b3, float b4, float* __restrict__ peer_lo, float* __restrict__ peer_hi) {
    const int per_plane = 2 * nx + 2 * ny;  // y- row, y+ row, x- column, x+ column
    const long long nside = (long long)per_plane * (zend - zbeg);
    const long long nbot = (zbeg == 0) ? (long long)nx * ny : 0;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < nside + nbot;
         t += (long long)gridDim.x * blockDim.x) {
        if (t < nside) {
            const int z = zbeg + (int)(t / per_plane), q = (int)(t % per_plane);
            if (z == 0) continue;  // the bottom plane is written whole below
            int i, j;
            float v;
            if (q < nx) { i = q; j = 0; v = (i == 0) ? b2 : (i == nx - 1) ? b3 : b0; }
            else if (q < 2 * nx) { i = q - nx; j = ny - 1; v = (i == 0) ? b2 : (i == nx - 1) ? b3 : b1; }
            else if (q < 2 * nx + ny) { i = 0; j = q - 2 * nx; v = b2; }
            else { i = nx - 1; j = q - 2 * nx - ny; v = b3; }
            const size_t in_plane = (size_t)j * nx + i;
            T[(size_t)z * nx * ny + in_plane] = v;
            if (z == zbeg && peer_lo) peer_lo[in_plane] = v;
            if (z == zend - 1 && peer_hi) peer_hi[in_plane] = v;
        } else {
            T[t - nside] = b4;
            if (zend == 1 && peer_hi) peer_hi[t - nside] = b4;
        }
    }
}

// Fused-halo push for a z-slab rank: the first / last owned plane of T_out (faces included) copied into the lower /
// upper neighbour's ghost plane in peer-mapped memory.  Every thread stores one 16-byte quad that is aligned on the
// DESTINATION side (the source is read with four scalar loads, it is L2-resident and arbitrarily misaligned relative
// to it), so a warp puts 512 contiguous, aligned bytes on NVLink per store; the ragged head and tail of the plane go
// as scalars.  The 120-byte row segments stored from inside the stencil kernel (K1F_PEER) cost 13-20 us per sweep
// once the TMA ring no longer leaves load stalls to hide them in.
__global__ void halo_push_kernel(const float* __restrict__ T, int plane, int zlo, int zhi, float* __restrict__ peer_lo,
                                 float* __restrict__ peer_hi) {
    const float* src = blockIdx.y == 0 ? T + (size_t)zlo * plane : T + (size_t)zhi * plane;
    float* dst = blockIdx.y == 0 ? peer_lo : peer_hi;
    if (!dst) return;
    const int head = (int)((4u - (((uintptr_t)dst >> 2) & 3u)) & 3u);  // scalars before the first aligned quad
    const int nquad = (plane - head) / 4;
    const long long t0 = blockIdx.x * (long long)blockDim.x + threadIdx.x, stride = (long long)gridDim.x * blockDim.x;
    for (long long q = t0; q < nquad; q += stride) {
        const long long i = head + 4 * q;
        const float4 v = make_float4(__ldcg(src + i), __ldcg(src + i + 1), __ldcg(src + i + 2), __ldcg(src + i + 3));
        *reinterpret_cast<float4*>(dst + i) = v;
    }
    if (t0 < head) dst[t0] = __ldcg(src + t0);
    const int tail0 = head + 4 * nquad;
    if (t0 < plane - tail0) dst[tail0 + t0] = __ldcg(src + tail0 + t0);
}

// ---- halo exchange of a z-slab rank with release / acquire counters (gomelt_step_args_t.halo_sync) -------------------
// One launch right after the step: blockIdx.y = 0 copies the first owned plane of T_out into the lower neighbour's upper
// ghost plane, blockIdx.y = 1 the last owned plane into the upper neighbour's lower ghost plane - destination-aligned
// 16-byte stores over NVLink, NB quads per thread in flight, the Dirichlet constants of the face rows / columns
// substituted on the way (the step never stores a face node).  Every block then publishes its part with a system-scope
// release add on the neighbour's arrival counter, and waits - acquire - until this rank's own counters show that both
// neighbours' planes of the same sweep have arrived, so that the NEXT step on this stream may read its ghost planes:
// no barrier launch, no NCCL call, no host involvement, and two temperature buffers suffice (a neighbour can only push
// sweep q + 1 after it has seen this rank's signal of sweep q, i.e. after this rank has finished reading sweep q's input).
// Counter block: [0] blocks of the LOWER neighbour that have delivered into this rank's lower ghost plane, [1] UPPER.
//
// An in-kernel variant (the stencil warp that finishes a strip of a boundary plane last pushes it and signals; ghost planes
// acquired by the warps that read them) was measured and dropped: every latency-bound operation at the end of a warp's march
// (fence, counter update, copy) holds the CTA slot while it waits, and a sweep is ~3.6 waves of CTAs: +49 us per 150 us
// sweep in loop-back on one GPU (fence 7, copy 16, counters + last-arriver fences 25), against ~12 us for this kernel.
GM_DI unsigned ld_acquire_sys(const unsigned* q) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(q) : "memory");
    return v;
}
GM_DI void red_release_sys_add(unsigned* q, unsigned v) {
    asm volatile("red.release.sys.global.add.u32 [%0], %1;" ::"l"(q), "r"(v) : "memory");
}
constexpr int HALO_THREADS = 256;
constexpr int HALO_NB = 4;  // quads per thread and pass
__global__ void __launch_bounds__(HALO_THREADS) halo_exchange_kernel(const float* __restrict__ T, int nx, int ny, int zlo, int zhi,
                                                                      float* __restrict__ peer_lo, float* __restrict__ peer_hi,
                                                                      float by0, float by1, float bx0, float bx1,
                                                                      unsigned* sync_mine, unsigned* sync_lo, unsigned* sync_hi,
                                                                      unsigned need, float* __restrict__ Tw, int nz, int zbeg,
                                                                      int zend, float bz0) {
    // assignBCs cF:1568-1595 on the owned planes of T_out (the step never stores a face node): y-, y+, x-, x+, z-, the
    // later face wins on shared edges.  Folded into this launch - a never-taken "am I a face CTA" branch inside the
    // stencil kernel costs it 5 % (measured), a separate launch costs a launch.
    if (Tw) {
        const int per_plane = 2 * nx + 2 * ny;
        const long long nside = (long long)per_plane * (zend - zbeg);
        const long long nbot = (zbeg == 0) ? (long long)nx * ny : 0;
        const long long stride = (long long)gridDim.x * gridDim.y * HALO_THREADS;
        for (long long t = ((long long)blockIdx.y * gridDim.x + blockIdx.x) * HALO_THREADS + threadIdx.x; t < nside + nbot; t += stride) {
            if (t < nside) {
                const int z = zbeg + (int)(t / per_plane), q = (int)(t % per_plane);
                if (z == 0) continue;
                int i, j;
                float v;
                if (q < nx) { i = q; j = 0; v = (i == 0) ? bx0 : (i == nx - 1) ? bx1 : by0; }
                else if (q < 2 * nx) { i = q - nx; j = ny - 1; v = (i == 0) ? bx0 : (i == nx - 1) ? bx1 : by1; }
                else if (q < 2 * nx + ny) { i = 0; j = q - 2 * nx; v = bx0; }
                else { i = nx - 1; j = q - 2 * nx - ny; v = bx1; }
                Tw[(size_t)z * nx * ny + (size_t)j * nx + i] = v;
            } else {
                Tw[t - nside] = bz0;
            }
        }
    }
    const int which = blockIdx.y;
    float* __restrict__ dst = which == 0 ? peer_lo : peer_hi;
    unsigned* peer_sync = which == 0 ? sync_lo : sync_hi;
    const int plane = nx * ny;
    if (dst) {
        const float* __restrict__ src = T + (size_t)(which == 0 ? zlo : zhi) * plane;
        const int elast = (ny - 1) * nx;
        auto fix = [&](float v, int e, int i) -> float {
            v = e < nx ? by0 : (e >= elast ? by1 : v);
            return i == 0 ? bx0 : (i == nx - 1 ? bx1 : v);
        };
        const int head = (int)((4u - ((unsigned)((uintptr_t)dst >> 2) & 3u)) & 3u);  // scalars before the first aligned quad
        const int nquad = (plane - head) >> 2;
        const int t = blockIdx.x * HALO_THREADS + threadIdx.x, nthreads = gridDim.x * HALO_THREADS;
        if (t < head) dst[t] = fix(__ldcg(src + t), t, t % nx);
        for (int base = 0; base < nquad; base += nthreads * HALO_NB) {
            float4 v[HALO_NB];
#pragma unroll
            for (int u = 0; u < HALO_NB; ++u) {
                const int q = base + u * nthreads + t;
                if (q < nquad) {
                    const float* s4 = src + head + 4 * q;
                    v[u] = make_float4(__ldcg(s4), __ldcg(s4 + 1), __ldcg(s4 + 2), __ldcg(s4 + 3));
                }
            }
#pragma unroll
            for (int u = 0; u < HALO_NB; ++u) {
                const int q = base + u * nthreads + t;
                if (q < nquad) {
                    const int e = head + 4 * q;
                    int i = e % nx;  // column of the quad's first element; the quad may wrap into the next row
                    float4 w;
                    w.x = fix(v[u].x, e, i);
                    i = (i + 1 == nx) ? 0 : i + 1;
                    w.y = fix(v[u].y, e + 1, i);
                    i = (i + 1 == nx) ? 0 : i + 1;
                    w.z = fix(v[u].z, e + 2, i);
                    i = (i + 1 == nx) ? 0 : i + 1;
                    w.w = fix(v[u].w, e + 3, i);
                    *reinterpret_cast<float4*>(dst + e) = w;
                }
            }
        }
        const int t0 = head + 4 * nquad;
        if (t < plane - t0) dst[t0 + t] = fix(__ldcg(src + t0 + t), t0 + t, (t0 + t) % nx);
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) red_release_sys_add(peer_sync + (which == 0 ? 1 : 0), 1u);
    }
    // arrivals: the lower neighbour's top plane ([0]) and the upper neighbour's bottom plane ([1]) of this sweep
    if (threadIdx.x == 0) {
        if (peer_lo) while ((int)(ld_acquire_sys(sync_mine + 0) - need) < 0) __nanosleep(100);
        if (peer_hi) while ((int)(ld_acquire_sys(sync_mine + 1) - need) < 0) __nanosleep(100);
    }
}

}  // namespace gomelt
