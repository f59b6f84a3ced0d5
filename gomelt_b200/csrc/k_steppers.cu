// The reference's jitted steppers as single native calls: stepGOMELT cF:2304-2397, subcycleGOMELT cF:3224-3632,
// stepGOMELTDwellTime cF:2617-2664.  Host-side orchestration only: every arithmetic step is one of the kernels of
// k_level_step.cu / k_aux.cu / k_transfer.cu, issued back to back on the caller's stream with no host synchronisation,
// no allocation and no Python in between (the reference runs each entry point as one jax.jit; at the example's sizes a
// subcycle block is ~150 launches of a few microseconds each, so the issue rate of the host IS the run time).
//
// What the sequence no longer contains, compared with a call-by-call transcription of the reference:
//   * no k / rho*cp arrays: the correction projections evaluate computeStateProperties at the nodes they read;
//   * no F + V additions: the projected source is written into the load vector and the corrections accumulate into it;
//   * no per-row passes over a parent level: computeLevelSource is two launches for any number of laser rows;
//   * no copies of the state: Level-2 / Level-3 state is advanced in place where the reference's functional update
//     allows it (the corrector pass), on scratch copies where it does not (the predictor pass), and the last substep of
//     a level writes straight into the caller's buffer.
#include <string.h>

#include "common.cuh"

using namespace gomelt;

namespace {

struct Carver {
    float* p;
    long long left;
    bool ok;
    float* take(long long n) {
        n = (n + 31) & ~31LL;  // 128-byte granules: every carved field can feed the TMA ring of K1
        if (n > left) {
            ok = false;
            left = 0;
            return nullptr;
        }
        float* r = p;
        p += n;
        left -= n;
        return r;
    }
};

inline long long nn_of(const gomelt_grid_t& g) { return (long long)g.nx * g.ny * g.nz; }
inline long long nsum_of(const gomelt_grid_t& g) { return (long long)g.nx + g.ny + g.nz; }
inline long long cells_of(const gomelt_pair_t& q) { return (long long)q.ncell[0] * q.ncell[1] * q.ncell[2] * 8; }
inline void axes_of(const gomelt_level_t& L, gomelt_axis_t ax[3]) {
    ax[0].coords = L.x; ax[0].n = L.grid.nx;
    ax[1].coords = L.y; ax[1].n = L.grid.ny;
    ax[2].coords = L.z; ax[2].n = L.grid.nz;
}
inline float wq_of(const gomelt_grid_t& g) { return ((g.hx * g.hy) * g.hz) * 0.125f; }

struct Work {
    float *rhs1, *T1a;
    float *rhs2, *T2a, *T2b, *S12p, *Tp2n;
    float *T3a, *T3b, *S13p, *Tp3c, *Tp3h, *tables3, *faces3;
    float *ptab, *cs21, *cs31, *cs32;
    uint8_t* preS2;
    uint32_t* bkq;
    long long bkq_words;
};

long long carve(const gomelt_hier_t& h, int N2, int N3, float* base, long long have, Work* w) {
    Carver c{base, have, true};
    const long long n1 = nn_of(h.L1.grid), n2 = nn_of(h.L2.grid), n3 = nn_of(h.L3.grid);
    Work t;
    t.rhs1 = c.take(n1); t.T1a = c.take(n1);
    t.rhs2 = c.take(n2); t.T2a = c.take(n2); t.T2b = c.take(n2); t.S12p = c.take(n2); t.Tp2n = c.take(n2);
    t.T3a = c.take(n3); t.T3b = c.take(n3); t.S13p = c.take(n3); t.Tp3c = c.take(n3);
    t.Tp3h = c.take((long long)(N2 > 0 ? N2 : 1) * ((n3 + 31) & ~31LL));
    t.tables3 = c.take((long long)(N3 > 0 ? N3 : 1) * nsum_of(h.L3.grid));
    t.faces3 = c.take(2 * gomelt_faces_count(h.L3.grid.nx, h.L3.grid.ny, h.L3.grid.nz));
    const long long ps = nsum_of(h.L1.grid) > nsum_of(h.L2.grid) ? nsum_of(h.L1.grid) : nsum_of(h.L2.grid);
    t.ptab = c.take((long long)GOMELT_MAX_SUBSTEPS * ps);
    t.cs21 = c.take(cells_of(h.L2L1)); t.cs31 = c.take(cells_of(h.L3L1)); t.cs32 = c.take(cells_of(h.L3L2));
    t.preS2 = reinterpret_cast<uint8_t*>(c.take((n3 + 3) / 4));
    t.bkq_words = 2 + 2 * (n3 / 120 + 1024);   // hot-plane queue of the corrector substeps (gomelt_step_args_t.bk_queue)
    t.bkq = reinterpret_cast<uint32_t*>(c.take(t.bkq_words));
    if (w) *w = t;
    return c.ok ? have - c.left : -1;
}

__global__ void fill_f32_kernel(float* __restrict__ x, long long n, float v) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) x[i] = v;
}
__global__ void reset_mask_kernel(const uint8_t* __restrict__ pre, const uint8_t* __restrict__ now, uint8_t* __restrict__ out,
                                  long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = (!pre[i] && now[i]) ? 1 : 0;  // ((1 - 2 preS2) * S2) == 1, cF:2394
}
// blockIdx.z = window plane, blockIdx.y = group of rpb window rows: the Level-0 row / plane offsets are formed once per row,
// no index division per node (the flat grid-stride form with its 64-bit divisions ran 81 us per 10.3 M-node window)
__global__ void accum_single_step_kernel(const float* __restrict__ T3, const uint8_t* __restrict__ mask, float dt, float Tliq,
                                         float* __restrict__ acc, float* __restrict__ mx, const int* __restrict__ ix,
                                         const int* __restrict__ iy, const int* __restrict__ iz, int nx, int ny, int nz, int bnx,
                                         int bny, int rpb) {
    for (int k = blockIdx.z; k < nz; k += gridDim.z) {
        const long long gk = (long long)iz[k] * bnx * bny;
        const int j1 = min(ny, (int)(blockIdx.y + 1) * rpb);
        for (int j = blockIdx.y * rpb; j < j1; ++j) {
            const long long grow = gk + (long long)iy[j] * bnx, trow = ((long long)k * ny + j) * nx;
            for (int i = threadIdx.x; i < nx; i += blockDim.x) {
                const long long g = grow + ix[i], t = trow + i;
                const float a = acc[g];
                const float reset = mask[t] ? a : a * 0.0f;        // accum * (all_reset > 0)
                mx[g] = fmaxf(reset, mx[g]);
                float an = __fadd_rn(a, -reset);
                an = __fadd_rn(an, (T3[t] > Tliq) ? dt : 0.0f);   // melting_temp: accum[idx] += (T > T_melt) * dt
                acc[g] = an;
            }
        }
    }
}
// dst[ix[i] + iy[j] * bnx + iz[k] * bnx * bny] = 0 over a tensor-product index set (one block row per (j, k) line)
__global__ void box_zero_u8_kernel(uint8_t* __restrict__ dst, const int* __restrict__ ix, const int* __restrict__ iy,
                                   const int* __restrict__ iz, int nx, int ny, int nz, int bnx, int bny) {
    for (int line = blockIdx.x; line < ny * nz; line += gridDim.x) {
        const int j = line % ny, k = line / ny;
        uint8_t* d = dst + (long long)iy[j] * bnx + (long long)iz[k] * bnx * bny;
        for (int i = threadIdx.x; i < nx; i += blockDim.x) d[ix[i]] = 0;
    }
}
// x = max(x, lo) on the node box [b0, b0 + bn) of an (nx, ny, nz) array
__global__ void box_clamp_min_kernel(float* __restrict__ x, int nx, int ny, int b0x, int b0y, int b0z, int bnx, int bny, int bnz,
                                     float lo) {
    for (int line = blockIdx.x; line < bny * bnz; line += gridDim.x) {
        const int j = line % bny, k = line / bny;
        float* d = x + ((long long)(b0z + k) * ny + (b0y + j)) * nx + b0x;
        for (int i = threadIdx.x; i < bnx; i += blockDim.x) d[i] = fmaxf(d[i], lo);
    }
}
inline int blocks_for(long long n) {
    long long b = (n + 255) / 256;
    const long long cap = 8LL * sm_count();
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

#define GM_TRY(expr)          \
    do {                      \
        const int rc_ = (expr); \
        if (rc_) return rc_;  \
    } while (0)

struct Ctx {
    const gomelt_props_t* props;
    const gomelt_hier_t* h;
    Work w;
    cudaStream_t st;
    void* stream;
    gomelt_axis_t a1[3], a2[3], a3[3];
    float T_amb;
};

// fused level step of one level (gomelt_level_step_f32)
int k1(const Ctx& c, const gomelt_level_t& L, const float* T0, const float* S1, float* Tout, float dt, const float* rhs,
       const float* tables, float coef, int flags, float* S1_out = nullptr) {
    if (&L == &c.h->L1 && c.h->l1_solve) {  // slab-decomposed Level 1: the hook sweeps the slabs of every rank
        const int rc = c.h->l1_solve(c.h->l1_user, T0, S1, Tout, dt, rhs, flags);
        if (rc) set_error("Level-1 solve hook returned %d", rc);
        return rc;
    }
    gomelt_step_args_t s;
    memset(&s, 0, sizeof s);
    s.grid = L.grid;
    s.T0 = T0; s.S1 = S1; s.rhs = rhs;
    if (tables) {
        s.src_x = tables; s.src_y = tables + L.grid.nx; s.src_z = tables + L.grid.nx + L.grid.ny;
        s.src_coef = coef;
    }
    s.dt = dt;
    s.nz_active = (&L == &c.h->L1) ? c.h->nz_active_L1 : L.grid.nz;
    s.n_substrate = L.n_substrate;
    s.flags = flags | (S1_out ? GOMELT_STEP_WRITE_S1 : 0);
    for (int q = 0; q < 5; ++q) s.bc5[q] = c.h->bc5[q];
    s.T_out = Tout;
    s.S1_out = S1_out;
    return gomelt_level_step_f32(c.props, &s, c.stream);
}
const int F_L1 = GOMELT_STEP_BC_CONST | GOMELT_STEP_FUSED_FLUX;
const int F_CHILD = GOMELT_STEP_SKIP_FACES | GOMELT_STEP_FUSED_FLUX;

// assignBCsFine cF:1598-1620 (+ time blend cF:3349 / 3389, + max(T_amb, .) when clamp): the five faces of Tchild
int faces(const Ctx& c, const gomelt_axis_t parent[3], const float* u, const float* u2, float alpha, float beta,
          const gomelt_level_t& child, float* Tchild, bool clamp) {
    gomelt_interp_args_t a;
    memset(&a, 0, sizeof a);
    for (int d = 0; d < 3; ++d) a.src[d] = parent[d];
    a.u = u; a.u2 = u2; a.alpha = alpha; a.beta = beta;
    a.tx = child.x; a.ty = child.y; a.tz = child.z;
    a.ntx = child.grid.nx; a.nty = child.grid.ny; a.ntz = child.grid.nz;
    a.mode = GOMELT_INTERP_SET; a.faces_only = 1;
    a.has_clamp = clamp ? 1 : 0; a.clamp_min = c.T_amb;
    a.out = Tchild;
    return gomelt_interp_f32(&a, c.stream);
}

// getNewTprime cF:2060-2099: parent[overlap] <- I_fine(fineT) in place; Tp <- fineT - I_parent(parentT)
int new_tprime(const Ctx& c, const gomelt_level_t& fine, const gomelt_axis_t fa[3], const float* fineT,
               const gomelt_level_t& parent, const gomelt_axis_t pa[3], float* parentT, const gomelt_overlap_t& ov, float* Tp) {
    gomelt_interp_args_t a;
    memset(&a, 0, sizeof a);
    for (int d = 0; d < 3; ++d) a.src[d] = fa[d];
    a.u = fineT; a.alpha = 1.f;
    a.tx = ov.cx; a.ty = ov.cy; a.tz = ov.cz;
    a.ntx = ov.n[0]; a.nty = ov.n[1]; a.ntz = ov.n[2];
    a.mode = GOMELT_INTERP_SET;
    a.map_x = ov.ix; a.map_y = ov.iy; a.map_z = ov.iz; a.map_nx = parent.grid.nx; a.map_ny = parent.grid.ny;
    a.out = parentT;
    GM_TRY(gomelt_interp_f32(&a, c.stream));
    memset(&a, 0, sizeof a);
    for (int d = 0; d < 3; ++d) a.src[d] = pa[d];
    a.u = parentT; a.alpha = 1.f;
    a.tx = fine.x; a.ty = fine.y; a.tz = fine.z;
    a.ntx = fine.grid.nx; a.nty = fine.grid.ny; a.ntz = fine.grid.nz;
    a.mode = GOMELT_INTERP_RSUB; a.base = fineT; a.out = Tp;
    return gomelt_interp_f32(&a, c.stream);
}

// V (+)= projected correction term; the coefficient (k or rho*cp) is evaluated from (cT, cS1) inside the kernel
int project(const Ctx& c, const gomelt_pair_t& pr, float* cellsum, const gomelt_axis_t fa[3], const gomelt_axis_t pa[3],
            const float* A, const float* A2, const float* cT, const float* cS1, long long cnsub, int mode, float scale, float* V) {
    gomelt_project_args_t a;
    memset(&a, 0, sizeof a);
    for (int d = 0; d < 3; ++d) { a.fine[d] = fa[d]; a.parent[d] = pa[d]; }
    a.A = A; a.A2 = A2; a.coef = nullptr;
    a.coef_T = cT; a.coef_S1 = cS1; a.coef_n_substrate = cnsub; a.coef_props = c.props;
    a.mode = mode; a.scale = scale;
    for (int d = 0; d < 3; ++d) { a.cell0[d] = pr.cell0[d]; a.ncell[d] = pr.ncell[d]; }
    a.first_x = pr.first_x; a.first_y = pr.first_y; a.first_z = pr.first_z;
    a.elems_per_cell_hint = pr.elems_per_cell_hint;
    a.wtab_x = pr.wtab_x; a.wtab_y = pr.wtab_y; a.wtab_z = pr.wtab_z;
    for (int d = 0; d < 3; ++d) { a.rmax[d] = pr.rmax[d]; a.uniform_off[d] = pr.uniform_off[d]; }
    const gomelt_grid_t& gf = (fa == c.a3) ? c.h->L3.grid : c.h->L2.grid;
    const gomelt_grid_t& gp = (pa == c.a1) ? c.h->L1.grid : c.h->L2.grid;
    a.hf[0] = gf.hx; a.hf[1] = gf.hy; a.hf[2] = gf.hz;
    a.hc[0] = gp.hx; a.hc[1] = gp.hy; a.hc[2] = gp.hz;
    a.cellsum = cellsum; a.V = V; a.accumulate = 1;
    return gomelt_project_f32(&a, c.stream);
}

// updateStateProperties cF:2546-2556 / 3272-3278: Level-2 S1 -> Level-1 overlap nodes, substrate planes -> 1
int push_S1_to_L1(const Ctx& c) {
    const gomelt_hier_t& h = *c.h;
    gomelt_interp_args_t a;
    memset(&a, 0, sizeof a);
    for (int d = 0; d < 3; ++d) a.src[d] = c.a2[d];
    a.u = h.L2.S1; a.alpha = 1.f;
    a.tx = h.ov2.cx; a.ty = h.ov2.cy; a.tz = h.ov2.cz;
    a.ntx = h.ov2.n[0]; a.nty = h.ov2.n[1]; a.ntz = h.ov2.n[2];
    a.mode = GOMELT_INTERP_SET;
    a.map_x = h.ov2.ix; a.map_y = h.ov2.iy; a.map_z = h.ov2.iz; a.map_nx = h.L1.grid.nx; a.map_ny = h.L1.grid.ny;
    a.out = h.L1.S1;
    GM_TRY(gomelt_interp_f32(&a, c.stream));
    if (h.L1.n_substrate > 0) {
        fill_f32_kernel<<<blocks_for(h.L1.n_substrate), 256, 0, c.st>>>(h.L1.S1, h.L1.n_substrate, 1.0f), count_launch();
        GM_TRY(check_launch("push_S1_to_L1"));
    }
    return 0;
}

// cF:2390-2392 / 3628-3630: Level-3 state back into Level 0
int scatter_L0(const Ctx& c) {
    const gomelt_hier_t& h = *c.h;
    const gomelt_grid_t& g = h.L3.grid;
    GM_TRY(gomelt_box_copy(h.L3.S1, h.L0_S1, 4, h.l0_ix, h.l0_iy, h.l0_iz, g.nx, g.ny, g.nz, h.L0_nx, h.L0_ny, 1, c.stream));
    // cF:2391: Level-0 S2 = 0 everywhere, then the window.  The only nodes that can be non-zero are the ones the previous
    // call scattered: with their index set given, clear those instead of the whole state grid (672 MB at config-4 size)
    if (h.l0p_ix && h.l0p_iy && h.l0p_iz && h.l0p_n[0] > 0 && h.l0p_n[1] > 0 && h.l0p_n[2] > 0) {
        const long long lines = (long long)h.l0p_n[1] * h.l0p_n[2];
        const long long cap = 16LL * sm_count();
        box_zero_u8_kernel<<<(int)(lines < cap ? lines : cap), 256, 0, c.st>>>(h.L0_S2, h.l0p_ix, h.l0p_iy, h.l0p_iz, h.l0p_n[0],
                                                                              h.l0p_n[1], h.l0p_n[2], h.L0_nx, h.L0_ny), count_launch();
        GM_TRY(check_launch("scatter_L0"));
    } else if (cudaMemsetAsync(h.L0_S2, 0, (size_t)h.L0_nx * h.L0_ny * h.L0_nz, c.st) != cudaSuccess) {
        return check_launch("scatter_L0");
    }
    return gomelt_box_copy(h.L3.S2, h.L0_S2, 1, h.l0_ix, h.l0_iy, h.l0_iz, g.nx, g.ny, g.nz, h.L0_nx, h.L0_ny, 1, c.stream);
}

// max(T_amb, .) on the Level-1 nodes that Level 2 can see (the parent cells under the window, one node wider): what the
// predictor pass needs of the clamp of cF:2360 - its Level-1 field is only prolonged / injected there and then discarded
int clamp_l1_box(const Ctx& c, float* T1) {
    const gomelt_hier_t& h = *c.h;
    const gomelt_grid_t& g = h.L1.grid;
    const int dims[3] = {g.nx, g.ny, g.nz};
    int lo[3], n[3];
    for (int d = 0; d < 3; ++d) {
        int a = h.L2L1.cell0[d] - 1, b = h.L2L1.cell0[d] + h.L2L1.ncell[d] + 1;
        a = a < 0 ? 0 : a;
        b = b > dims[d] - 1 ? dims[d] - 1 : b;
        lo[d] = a;
        n[d] = b - a + 1;
        if (n[d] < 1) return 0;
    }
    const long long lines = (long long)n[1] * n[2];
    const long long cap = 16LL * sm_count();
    box_clamp_min_kernel<<<(int)(lines < cap ? lines : cap), 128, 0, c.st>>>(T1, g.nx, g.ny, lo[0], lo[1], lo[2], n[0], n[1], n[2],
                                                                              c.T_amb), count_launch();
    return check_launch("clamp_l1_box");
}

int copy_f32(const Ctx& c, float* dst, const float* src, long long n) {
    if (dst == src) return 0;
    if (cudaMemcpyAsync(dst, src, (size_t)n * 4, cudaMemcpyDeviceToDevice, c.st) != cudaSuccess) return check_launch("copy");
    return 0;
}

int check_hier(const gomelt_props_t* props, const gomelt_hier_t* h, const char* who, bool need_windows) {
    if (!props || !h || !h->L1.T0 || !h->L1.S1 || !h->L1.x || !h->L1.y || !h->L1.z || !h->work) {
        set_error("%s: NULL props / hierarchy / Level-1 field / work", who);
        return GOMELT_E_NULL;
    }
    if (need_windows) {
        const gomelt_level_t* Ls[2] = {&h->L2, &h->L3};
        for (const gomelt_level_t* L : Ls)
            if (!L->T0 || !L->S1 || !L->Tprime0 || !L->x || !L->y || !L->z) {
                set_error("%s: NULL Level-2 / Level-3 field", who);
                return GOMELT_E_NULL;
            }
        if (!h->L3.S2 || !h->L0_S1 || !h->L0_S2 || !h->l0_ix || !h->l0_iy || !h->l0_iz || !h->ov2.ix || !h->ov3.ix ||
            !h->ov2.cx || !h->ov3.cx || !h->L2L1.first_x || !h->L3L1.first_x || !h->L3L2.first_x) {
            set_error("%s: NULL Level-3 S2 / Level-0 state / overlap set / pair grouping", who);
            return GOMELT_E_NULL;
        }
    }
    return 0;
}

int make_ctx(Ctx& c, const gomelt_props_t* props, const gomelt_hier_t* h, int N2, int N3, void* stream, const char* who) {
    c.props = props; c.h = h; c.stream = stream; c.st = (cudaStream_t)stream;
    axes_of(h->L1, c.a1); axes_of(h->L2, c.a2); axes_of(h->L3, c.a3);
    c.T_amb = props->T_amb;
    if (carve(*h, N2, N3, h->work, h->work_floats, &c.w) < 0) {
        set_error("%s: work_floats = %lld is smaller than gomelt_hier_work_floats() = %lld", who, (long long)h->work_floats,
                  gomelt_hier_work_floats(h, N2, N3));
        return GOMELT_E_SIZE;
    }
    return 0;
}

}  // namespace

extern "C" long long gomelt_hier_work_floats(const gomelt_hier_t* h, int32_t N2, int32_t N3) {
    if (!h) return -1;
    return carve(*h, N2, N3, nullptr, (1LL << 60), nullptr);
}

extern "C" int gomelt_dwell_step_f32(const gomelt_props_t* props, const gomelt_hier_t* h, float dt, int32_t* l1_in_spare,
                                     void* stream) {
    GM_TRY(check_hier(props, h, "gomelt_dwell_step_f32", false));
    Ctx c;
    c.props = props; c.h = h; c.stream = stream; c.st = (cudaStream_t)stream;
    c.T_amb = props->T_amb;
    const long long n1 = nn_of(h->L1.grid);
    float* dst = h->L1_spare;
    if (!dst) {
        if (h->work_floats < n1) {
            set_error("gomelt_dwell_step_f32: work_floats < Level-1 nodes and no L1_spare");
            return GOMELT_E_SIZE;
        }
        dst = h->work;
    }
    // stepGOMELTDwellTime cF:2617-2664: flux -> properties -> solve (Corr = 0) -> inactive fill -> BCs; no clamp
    GM_TRY(k1(c, h->L1, h->L1.T0, h->L1.S1, dst, dt, nullptr, nullptr, 0.f, F_L1));
    if (h->L1_spare) {
        if (l1_in_spare) *l1_in_spare = 1;
        return 0;
    }
    if (l1_in_spare) *l1_in_spare = 0;
    return copy_f32(c, h->L1.T0, dst, n1);
}

extern "C" int gomelt_step_f32(const gomelt_props_t* props, const gomelt_hier_t* h, const float* row, uint8_t* resetmask,
                               int32_t* l1_in_spare, void* stream) {
    GM_TRY(check_hier(props, h, "gomelt_step_f32", true));
    if (!row) {
        set_error("gomelt_step_f32: NULL toolpath row");
        return GOMELT_E_NULL;
    }
    Ctx c;
    GM_TRY(make_ctx(c, props, h, 1, 1, stream, "gomelt_step_f32"));
    const gomelt_level_t &L1 = h->L1, &L2 = h->L2, &L3 = h->L3;
    const Work& w = c.w;
    const long long n1 = nn_of(L1.grid), n2 = nn_of(L2.grid), n3 = nn_of(L3.grid);
    const float dt = row[5];
    // updateStateProperties cF:2513-2564 (k and rho*cp are evaluated where they are consumed)
    if (cudaMemcpyAsync(w.preS2, L3.S2, (size_t)n3, cudaMemcpyDeviceToDevice, c.st) != cudaSuccess) return check_launch("gomelt_step_f32");
    GM_TRY(gomelt_state_props_f32(props, L3.T0, L3.S1, n3, L3.n_substrate, L3.S1, L3.S2, nullptr, nullptr, stream));
    GM_TRY(gomelt_state_props_f32(props, L2.T0, L2.S1, n2, L2.n_substrate, L2.S1, nullptr, nullptr, nullptr, stream));
    GM_TRY(push_S1_to_L1(c));
    // loads: laser source on Level 3 (rank-1 tables) and its projections on Levels 1 and 2 (computeSources cF:928-988);
    // the surface fluxes (computeConvRadBC cF:2348-2350) are evaluated inside each level step
    float coef3 = 0.f;
    GM_TRY(gomelt_source_tables_batch_f32(props, &L3.grid, L3.x, L3.y, L3.z, row, 1, w.tables3, &coef3, stream));
    GM_TRY(gomelt_projected_source_f32(props, c.a3, c.a1, wq_of(L3.grid), row, 1, w.ptab, w.rhs1, 0, stream));
    GM_TRY(gomelt_projected_source_f32(props, c.a3, c.a2, wq_of(L3.grid), row, 1, w.ptab, w.rhs2, 0, stream));
    // computeCoarseTprimeTerm_jax cF:1477-1565
    GM_TRY(project(c, h->L3L1, w.cs31, c.a3, c.a1, L3.Tprime0, nullptr, L3.T0, L3.S1, L3.n_substrate, 0, 1.f, w.rhs1));
    GM_TRY(project(c, h->L2L1, w.cs21, c.a2, c.a1, L2.Tprime0, nullptr, L2.T0, L2.S1, L2.n_substrate, 0, 1.f, w.rhs1));
    GM_TRY(project(c, h->L3L2, w.cs32, c.a3, c.a2, L3.Tprime0, nullptr, L3.T0, L3.S1, L3.n_substrate, 0, 1.f, w.rhs2));
    float* T1n = h->L1_spare ? h->L1_spare : w.T1a;
    for (int pass = 0; pass < 2; ++pass) {
        // computeSolutions cF:2135-2204: the parents are prolonged UNCLAMPED onto the child faces, the
        // jnp.maximum(T_amb, .) of cF:2360-2362 / 2380-2382 follows
        GM_TRY(k1(c, L1, L1.T0, L1.S1, T1n, dt, w.rhs1, nullptr, 0.f, F_L1 | (h->l1_solve ? GOMELT_L1_CLAMP_AFTER : 0)));
        GM_TRY(k1(c, L2, L2.T0, L2.S1, w.T2a, dt, w.rhs2, nullptr, 0.f, F_CHILD));
        GM_TRY(faces(c, c.a1, T1n, nullptr, 1.f, 0.f, L2, w.T2a, false));
        // Level 3: same T0, F, k, rho*cp and Corr = 0 in both passes (cF:2199-2201): the corrector re-uses the
        // predictor interior and only the faces change
        if (pass == 0) GM_TRY(k1(c, L3, L3.T0, L3.S1, w.T3a, dt, nullptr, w.tables3, coef3, F_CHILD | GOMELT_STEP_CLAMP));
        GM_TRY(faces(c, c.a2, w.T2a, nullptr, 1.f, 0.f, L3, w.T3a, true));
        // (predictor: only the box under Level 2 is looked at again; with a slab-decomposed Level 1 the field is a mirror of
        // that box in both passes and the slabs clamp themselves)
        if (pass == 0 || h->l1_solve) GM_TRY(clamp_l1_box(c, T1n));
        else GM_TRY(gomelt_clamp_min_f32(T1n, n1, c.T_amb, stream));
        GM_TRY(gomelt_clamp_min_f32(w.T2a, n2, c.T_amb, stream));
        if (pass == 0) {
            // getBothNewTprimes cF:2102-2132, then computeCoarseTprimeMassTerm_jax cF:1396-1474
            GM_TRY(new_tprime(c, L3, c.a3, w.T3a, L2, c.a2, w.T2a, h->ov3, w.Tp3c));
            GM_TRY(new_tprime(c, L2, c.a2, w.T2a, L1, c.a1, T1n, h->ov2, w.Tp2n));
            const float inv_dt = 1.0f / dt;
            GM_TRY(project(c, h->L3L1, w.cs31, c.a3, c.a1, w.Tp3c, L3.Tprime0, L3.T0, L3.S1, L3.n_substrate, 1, inv_dt, w.rhs1));
            GM_TRY(project(c, h->L2L1, w.cs21, c.a2, c.a1, w.Tp2n, L2.Tprime0, L2.T0, L2.S1, L2.n_substrate, 1, inv_dt, w.rhs1));
            GM_TRY(project(c, h->L3L2, w.cs32, c.a3, c.a2, w.Tp3c, L3.Tprime0, L3.T0, L3.S1, L3.n_substrate, 1, inv_dt, w.rhs2));
        }
    }
    // final T' update for the next step (cF:2385-2387)
    GM_TRY(copy_f32(c, L3.T0, w.T3a, n3));
    GM_TRY(new_tprime(c, L3, c.a3, L3.T0, L2, c.a2, w.T2a, h->ov3, L3.Tprime0));
    GM_TRY(new_tprime(c, L2, c.a2, w.T2a, L1, c.a1, T1n, h->ov2, L2.Tprime0));
    GM_TRY(copy_f32(c, L2.T0, w.T2a, n2));
    if (h->L1_spare) {
        if (l1_in_spare) *l1_in_spare = 1;
    } else {
        if (l1_in_spare) *l1_in_spare = 0;
        GM_TRY(copy_f32(c, L1.T0, T1n, n1));
    }
    GM_TRY(scatter_L0(c));
    if (resetmask) {
        reset_mask_kernel<<<blocks_for(n3), 256, 0, c.st>>>(w.preS2, L3.S2, resetmask, n3), count_launch();
        GM_TRY(check_launch("gomelt_step_f32"));
    }
    return 0;
}

extern "C" int gomelt_subcycle_f32(const gomelt_props_t* props, const gomelt_hier_t* h, const float* rows, int32_t N2, int32_t N3,
                                   float* max_accum, float* accum, int32_t* l1_in_spare, void* stream) {
    GM_TRY(check_hier(props, h, "gomelt_subcycle_f32", true));
    if (!rows || !max_accum || !accum) {
        set_error("gomelt_subcycle_f32: NULL rows / accum / max_accum");
        return GOMELT_E_NULL;
    }
    if (N2 < 1 || N3 < 1 || N3 > GOMELT_MAX_SUBSTEPS || (long long)N2 * N3 > GOMELT_MAX_SUBSTEPS) {
        set_error("gomelt_subcycle_f32: N2 = %d, N3 = %d (N2 * N3 <= %d)", N2, N3, GOMELT_MAX_SUBSTEPS);
        return GOMELT_E_SIZE;
    }
    Ctx c;
    GM_TRY(make_ctx(c, props, h, N2, N3, stream, "gomelt_subcycle_f32"));
    const gomelt_level_t &L1 = h->L1, &L2 = h->L2, &L3 = h->L3;
    const Work& w = c.w;
    const long long n1 = nn_of(L1.grid), n2 = nn_of(L2.grid), n3 = nn_of(L3.grid);
    const long long n3p = (n3 + 31) & ~31LL;
    const float fN2 = (float)N2, fN3 = (float)N3;
    const float wq3 = wq_of(L3.grid);
    float dt_all = 0.f;
    for (int r = 0; r < N2 * N3; ++r) dt_all += rows[7 * r + 5];

    // ---- Level 1, predictor (cF:3262-3306): state push, load = projected source + gradient corrections ----
    GM_TRY(push_S1_to_L1(c));
    GM_TRY(gomelt_projected_source_f32(props, c.a3, c.a1, wq3, rows, N2 * N3, w.ptab, w.rhs1, 0, stream));
    GM_TRY(project(c, h->L3L1, w.cs31, c.a3, c.a1, L3.Tprime0, nullptr, L3.T0, L3.S1, L3.n_substrate, 0, 1.f, w.rhs1));
    GM_TRY(project(c, h->L2L1, w.cs21, c.a2, c.a1, L2.Tprime0, nullptr, L2.T0, L2.S1, L2.n_substrate, 0, 1.f, w.rhs1));
    GM_TRY(k1(c, L1, L1.T0, L1.S1, w.T1a, dt_all, w.rhs1, nullptr, 0.f, F_L1 | GOMELT_STEP_CLAMP));
    const float* L1new = w.T1a;

    // The inner scan (subcycleL3_Part1 / _Part2, cF:3367-3412 / 3530-3590) is gomelt_l3_substeps_f32: all N3 source
    // tables in one launch, then N3 x (fused level step + face prolongation from Level 2).
    gomelt_interp_args_t fa;
    auto l3_block = [&](const float* T3in, const float* S1in, float* S1io, float* Ta, float* Tb, const float* rows_i,
                        const float* L2new, const float* L2old, bool bookkeeping, float** Tlast) -> int {
        memset(&fa, 0, sizeof fa);
        for (int d = 0; d < 3; ++d) fa.src[d] = c.a2[d];
        fa.u = L2new; fa.u2 = L2old;
        fa.tx = L3.x; fa.ty = L3.y; fa.tz = L3.z;
        fa.ntx = L3.grid.nx; fa.nty = L3.grid.ny; fa.ntz = L3.grid.nz;
        fa.faces_only = 1; fa.has_clamp = 1; fa.clamp_min = c.T_amb;
        gomelt_substeps_args_t s;
        memset(&s, 0, sizeof s);
        s.grid = L3.grid;
        s.x = L3.x; s.y = L3.y; s.z = L3.z;
        s.n = N3; s.rows = rows_i;
        s.T_in = T3in; s.T_a = Ta; s.T_b = Tb;
        s.S1_in = S1in; s.S1 = S1io;
        s.n_substrate = L3.n_substrate;
        s.flags = GOMELT_STEP_SKIP_FACES | GOMELT_STEP_CLAMP | (bookkeeping ? (GOMELT_STEP_WRITE_S2 | GOMELT_STEP_ACCUM) : 0);
        s.tables = w.tables3;
        if (bookkeeping) { s.S2 = L3.S2; s.accum = accum; s.max_accum = max_accum; s.bk_queue = w.bkq; s.bk_queue_words = w.bkq_words; }
        s.faces = &fa; s.faces_n = fN3;
        s.T_last = Tlast;
        s.faces_scratch = N3 > 1 ? w.faces3 : nullptr;
        return gomelt_l3_substeps_f32(props, &s, stream);
    };

    for (int pass = 0; pass < 2; ++pass) {
        const bool corr = pass == 1;
        if (corr) {
            // ---- Level-1 corrector (cF:3432-3456): mass-term corrections from the predictor's T' ----
            const float* T2p = (N2 - 1) % 2 == 0 ? w.T2a : w.T2b;           // Level-2 field the predictor ended on
            GM_TRY(new_tprime(c, L2, c.a2, T2p, L1, c.a1, w.T1a, h->ov2, w.Tp2n));
            const float inv = 1.0f / dt_all;
            GM_TRY(project(c, h->L3L1, w.cs31, c.a3, c.a1, w.Tp3h + (long long)(N2 - 1) * n3p, L3.Tprime0, L3.T0, L3.S1, L3.n_substrate,
                           1, inv, w.rhs1));
            GM_TRY(project(c, h->L2L1, w.cs21, c.a2, c.a1, w.Tp2n, L2.Tprime0, L2.T0, L2.S1, L2.n_substrate, 1, inv, w.rhs1));
            float* dst = h->L1_spare ? h->L1_spare : w.T1a;
            GM_TRY(k1(c, L1, L1.T0, L1.S1, dst, dt_all, w.rhs1, nullptr, 0.f, F_L1 | GOMELT_STEP_CLAMP));
            L1new = dst;
        }
        // ---- Level-2 / Level-3 passes: predictor cF:3308-3430 on scratch state, corrector cF:3458-3622 in place ----
        const float* T2 = L2.T0;
        float* S12 = corr ? L2.S1 : w.S12p;
        if (!corr) GM_TRY(copy_f32(c, w.S12p, L2.S1, n2));
        const float* T3 = L3.T0;
        const float* Tp3 = L3.Tprime0;
        float* S13 = corr ? L3.S1 : w.S13p;
        for (int i2 = 0; i2 < N2; ++i2) {
            const bool last = corr && i2 == N2 - 1 && N2 >= 2;  // the caller's buffers are dead by then: write into them
            const float a2 = (float)(i2 + 1) / fN2, b2 = 1.0f - a2;
            const float* rows_i = rows + 7 * (size_t)N3 * i2;
            float dt2 = 0.f;
            for (int r = 0; r < N3; ++r) dt2 += rows_i[7 * r + 5];
            const float* S13cur = (!corr && i2 == 0) ? L3.S1 : S13;
            // computeLevelSource cF:2667-2730 + computeL2TprimeTerms_Part1 cF:2857-2914 (+ _Part2 cF:3169-3221)
            GM_TRY(gomelt_projected_source_f32(props, c.a3, c.a2, wq3, rows_i, N3, w.ptab, w.rhs2, 0, stream));
            GM_TRY(project(c, h->L3L2, w.cs32, c.a3, c.a2, Tp3, nullptr, T3, S13cur, L3.n_substrate, 0, 1.f, w.rhs2));
            if (corr)
                GM_TRY(project(c, h->L3L2, w.cs32, c.a3, c.a2, w.Tp3h + (long long)i2 * n3p, Tp3, T3, S13cur, L3.n_substrate, 1,
                               1.0f / dt2, w.rhs2));
            // computeL2Temperature cF:2917-2957: solve, faces <- a2 * L1_new + b2 * L1_old, max(T_amb, .)
            float* T2n = last ? L2.T0 : (i2 % 2 == 0 ? w.T2a : w.T2b);
            GM_TRY(k1(c, L2, T2, S12, T2n, dt2, w.rhs2, nullptr, 0.f, F_CHILD | GOMELT_STEP_CLAMP, S12));
            GM_TRY(faces(c, c.a1, L1new, L1.T0, a2, b2, L2, T2n, true));
            // Level-3 substeps; the block's last field lands in the caller's buffer on the final block
            float *Ta, *Tb;
            if (last) {
                float* other = (T3 == w.T3a) ? w.T3b : w.T3a;
                Ta = (N3 % 2 == 1) ? L3.T0 : other;
                Tb = (N3 % 2 == 1) ? other : L3.T0;
            } else {
                Ta = (T3 == w.T3a) ? w.T3b : w.T3a;   // T_a != T_in; T_in may be T_b
                Tb = (T3 == w.T3a) ? w.T3a : w.T3b;
            }
            float* T3n = nullptr;
            GM_TRY(l3_block(T3, (!corr && i2 == 0) ? L3.S1 : nullptr, S13, Ta, Tb, rows_i, T2n, T2, corr, &T3n));
            // getNewTprime cF:2060-2099 (Level 3 -> Level 2)
            float* Tpn = corr ? (i2 == N2 - 1 ? L3.Tprime0 : w.Tp3c) : w.Tp3h + (long long)i2 * n3p;
            GM_TRY(new_tprime(c, L3, c.a3, T3n, L2, c.a2, T2n, h->ov3, Tpn));
            T2 = T2n; T3 = T3n; Tp3 = Tpn;
        }
        if (corr) {
            GM_TRY(copy_f32(c, L2.T0, T2, n2));   // (no-ops when the last block wrote in place)
            GM_TRY(copy_f32(c, L3.T0, T3, n3));
        }
    }
    // ---- final T' of Level 2 and the injected Level-1 field (cF:3624-3626), Level-3 state into Level 0 ----
    float* L1fin = h->L1_spare ? h->L1_spare : w.T1a;
    GM_TRY(new_tprime(c, L2, c.a2, L2.T0, L1, c.a1, L1fin, h->ov2, L2.Tprime0));
    if (h->L1_spare) {
        if (l1_in_spare) *l1_in_spare = 1;
    } else {
        if (l1_in_spare) *l1_in_spare = 0;
        GM_TRY(copy_f32(c, L1.T0, L1fin, n1));
    }
    return scatter_L0(c);
}

__global__ void patch_copy_kernel(const float* __restrict__ src, int snx, int sny, int sx, int sy, int sz, float* __restrict__ dst,
                                  int dnx, int dny, int dx, int dy, int dz, int nx, int ny, int nz) {
    // one block row per (y, z) line of the box, threads along x
    for (int line = blockIdx.x; line < ny * nz; line += gridDim.x) {
        const int j = line % ny, k = line / ny;
        const float* s = src + ((long long)(sz + k) * sny + (sy + j)) * snx + sx;
        float* d = dst + ((long long)(dz + k) * dny + (dy + j)) * dnx + dx;
        for (int i = threadIdx.x; i < nx; i += blockDim.x) d[i] = s[i];
    }
}

extern "C" int gomelt_patch_copy_f32(const float* src, const int32_t sdims[3], const int32_t slo[3], float* dst,
                                     const int32_t ddims[3], const int32_t dlo[3], const int32_t n[3], void* stream) {
    if (!src || !dst || !sdims || !slo || !ddims || !dlo || !n) {
        set_error("gomelt_patch_copy_f32: NULL argument");
        return GOMELT_E_NULL;
    }
    for (int d = 0; d < 3; ++d)
        if (n[d] < 0 || slo[d] < 0 || dlo[d] < 0 || slo[d] + n[d] > sdims[d] || dlo[d] + n[d] > ddims[d]) {
            set_error("gomelt_patch_copy_f32: box [%d, %d) of axis %d outside the source (%d) or the destination (%d at %d)",
                      slo[d], slo[d] + n[d], d, sdims[d], ddims[d], dlo[d]);
            return GOMELT_E_SIZE;
        }
    if (n[0] == 0 || n[1] == 0 || n[2] == 0) return 0;
    const long long lines = (long long)n[1] * n[2];
    const int threads = n[0] >= 256 ? 256 : (n[0] >= 128 ? 128 : (n[0] >= 64 ? 64 : 32));
    const long long cap = 16LL * sm_count();
    patch_copy_kernel<<<(int)(lines < cap ? lines : cap), threads, 0, (cudaStream_t)stream>>>(
        src, sdims[0], sdims[1], slo[0], slo[1], slo[2], dst, ddims[0], ddims[1], dlo[0], dlo[1], dlo[2], n[0], n[1], n[2]),
        count_launch();
    return check_launch("gomelt_patch_copy_f32");
}

extern "C" int gomelt_accum_single_step_f32(const float* T3, const uint8_t* resetmask, float dt, float T_liquidus, float* accum0,
                                            float* max_accum0, const int32_t* ix, const int32_t* iy, const int32_t* iz, int32_t nx,
                                            int32_t ny, int32_t nz, int32_t big_nx, int32_t big_ny, void* stream) {
    if (!T3 || !resetmask || !accum0 || !max_accum0 || !ix || !iy || !iz) {
        set_error("gomelt_accum_single_step_f32: NULL argument");
        return GOMELT_E_NULL;
    }
    if (nx < 1 || ny < 1 || nz < 1) {
        set_error("gomelt_accum_single_step_f32: empty window");
        return GOMELT_E_SIZE;
    }
    const int rpb = (long long)ny * nz >= 16LL * sm_count() ? 4 : 1;   // rows per block
    const int threads = nx >= 256 ? 256 : (nx >= 128 ? 128 : 64);
    accum_single_step_kernel<<<dim3(1, (ny + rpb - 1) / rpb, nz < 65535 ? nz : 65535), threads, 0, (cudaStream_t)stream>>>(
        T3, resetmask, dt, T_liquidus, accum0, max_accum0, ix, iy, iz, nx, ny, nz, big_nx, big_ny, rpb), count_launch();
    return check_launch("gomelt_accum_single_step_f32");
}
