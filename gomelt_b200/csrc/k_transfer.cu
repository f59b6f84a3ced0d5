// Inter-level transfer kernels (sm_100a): K2 inject / T', K3 correction-vector projection,
// K4 face prolongation, K5 window shift, K6 projected source tables, box gather / scatter.
//
// The reference materialises every transfer operator - (n,8) weight + index matrices
// (interpolatePointsMatrix cF:1028-1107) and four (n_fine_elem, 8, 8) arrays + a BCOO matrix per level
// pair (computeCoarseFineShapeFunctions cF:1213-1358) - and rebuilds them on every window move.  On
// nested uniform grids each weight is a function of 1-D quantities, so nothing is materialised here:
// weights are evaluated in registers from the levels' 1-D node-coordinate arrays with the reference's
// own float32 formulas (floor((x-x0)/h) cell search, compute3DN cF:1361-1393 products, the +-1e-2
// validity window), so that cell decisions are identical and values agree to rounding.
#include <math.h>

#include "common.cuh"

namespace gomelt {

struct AxisView {
    const float* c;  // node coordinates
    int n;           // nodes
};

struct InterpParams {
    AxisView sx, sy, sz;       // source level
    const float* u;            // source field
    const float* u2;           // optional second source field: value = alpha*u + beta*u2
    float alpha, beta;
    const float* tx;           // target tensor grid
    const float* ty;
    const float* tz;
    int ntx, nty, ntz;
    int mode;                  // GOMELT_INTERP_*
    int faces_only;
    float clamp_min;
    int has_clamp;
    const int* mx;             // optional scatter map (tensor product of index vectors)
    const int* my;
    const int* mz;
    int map_nx, map_ny;
    const float* base;         // RSUB: out = base - I
    float* out;
};

__device__ __forceinline__ int cell_of(float x, float x0, float h, int ne) {
    // clip(floor((x - x0) / h), 0, ne - 1), IEEE division like jnp (cF:1069-1077)
    const float q = floorf(__fdiv_rn(__fsub_rn(x, x0), h));
    int e = (int)q;
    e = e < 0 ? 0 : e;
    return e > ne - 1 ? ne - 1 : e;
}

constexpr int ROWS3 = 8;  // target rows per block of the 3-D launch forms (a block per row is too little work per block)

// ---- lean form of the interpolant for full tensor-product target grids ---------------------------------------------------
// The per-axis part of a target's weights (cell by the floor rule, the two distances to the cell's nodes) depends on one
// target coordinate only: a thread keeps the x part of its column for all the rows it handles, the y part is formed
// once per row and the z part once per block.  The 8 weights are then the reference's own products
// ((ax * ay) * az) * inv_vol with its +-1e-2 validity window and [0, 1] clip (cF:1375-1391, 1101-1104): same values as
// interp_kernel, a quarter of the instructions (no divisions, no coordinate loads, 32-bit indices in the inner part).
struct Ax1 {
    int e;
    float a0, a1;  // x1 - x, x - x0 of cell e
};
__device__ __forceinline__ Ax1 ax1_of(const AxisView& ax, float h, float x) {
    Ax1 r;
    r.e = cell_of(x, ax.c[0], h, ax.n - 1);
    r.a0 = __fsub_rn(ax.c[r.e + 1], x);
    r.a1 = __fsub_rn(x, ax.c[r.e]);
    return r;
}
struct SrcGeom {  // per source level, formed once per thread
    float hx, hy, hz, inv_vol;
    int nnx, nnxy, nny;
};
__device__ __forceinline__ SrcGeom geom_of(const AxisView& sx, const AxisView& sy, const AxisView& sz) {
    SrcGeom g;
    g.hx = __fsub_rn(sx.c[1], sx.c[0]);
    g.hy = __fsub_rn(sy.c[1], sy.c[0]);
    g.hz = __fsub_rn(sz.c[1], sz.c[0]);
    g.inv_vol = __fdiv_rn(1.0f, __fmul_rn(__fmul_rn(g.hx, g.hy), g.hz));
    g.nnx = sx.n;
    g.nnxy = sx.n * sy.n;
    g.nny = sy.n;
    return g;
}
template <bool BLEND>
__device__ __forceinline__ float tri_eval(const float* __restrict__ u, const float* __restrict__ u2, float alpha, float beta,
                                          const SrcGeom& g, const Ax1& X, const Ax1& Y, const Ax1& Z) {
    // the 8 parent values are requested first and unconditionally (the cell is clamped into the parent, so the addresses
    // are valid for every target), the validity window is a select at the end: no branch stands between the loads of
    // one interpolation and the next, so a caller with several parents (the window shift) has all of them in flight at once
    const int b = X.e + Y.e * g.nnx + Z.e * g.nnxy;
    const int nd[8] = {b, b + 1, b + 1 + g.nnx, b + g.nnx, b + g.nnxy, b + 1 + g.nnxy, b + 1 + g.nnx + g.nnxy, b + g.nnx + g.nnxy};
    float v[8];
#pragma unroll
    for (int a = 0; a < 8; ++a) {
        v[a] = __ldg(u + nd[a]);
        if (BLEND) v[a] = __fadd_rn(__fmul_rn(alpha, v[a]), __fmul_rn(beta, __ldg(u2 + nd[a])));
    }
    const float xy00 = __fmul_rn(X.a0, Y.a0), xy10 = __fmul_rn(X.a1, Y.a0), xy11 = __fmul_rn(X.a1, Y.a1), xy01 = __fmul_rn(X.a0, Y.a1);
    float N[8];  // hex8 local order
    N[0] = __fmul_rn(__fmul_rn(xy00, Z.a0), g.inv_vol);
    N[1] = __fmul_rn(__fmul_rn(xy10, Z.a0), g.inv_vol);
    N[2] = __fmul_rn(__fmul_rn(xy11, Z.a0), g.inv_vol);
    N[3] = __fmul_rn(__fmul_rn(xy01, Z.a0), g.inv_vol);
    N[4] = __fmul_rn(__fmul_rn(xy00, Z.a1), g.inv_vol);
    N[5] = __fmul_rn(__fmul_rn(xy10, Z.a1), g.inv_vol);
    N[6] = __fmul_rn(__fmul_rn(xy11, Z.a1), g.inv_vol);
    N[7] = __fmul_rn(__fmul_rn(xy01, Z.a1), g.inv_vol);
    const float lo = fminf(fminf(fminf(N[0], N[1]), fminf(N[2], N[3])), fminf(fminf(N[4], N[5]), fminf(N[6], N[7])));
    const float hi = fmaxf(fmaxf(fmaxf(N[0], N[1]), fmaxf(N[2], N[3])), fmaxf(fmaxf(N[4], N[5]), fmaxf(N[6], N[7])));
    float acc = 0.f;
#pragma unroll
    for (int a = 0; a < 8; ++a) acc = __fadd_rn(acc, __fmul_rn(fminf(fmaxf(N[a], 0.f), 1.f), v[a]));
    return (lo >= -1e-2f && hi <= 1.0f + 1e-2f) ? acc : 0.f;
}

// ---- marching form ---------------------------------------------------------------------------------------------------
// Full target grid, no index map (interp3_march_kernel): a thread owns one target column and walks ROWSM rows of one
// target plane.  Three things make it cheaper than evaluating every target on its own (interp_kernel, which the
// mapped / faces-only calls still use, and which the tests reach with an identity index map) while producing the SAME bits:
//  * the y part of a row (cell, distances), the z part of the plane and 1 / (hx hy hz) are formed once per block into
//    shared memory (SrcShared), the x part once per thread; the field the result is combined with (ADD / RSUB) is
//    requested for all rows before anything else;
//  * the 8 parent values stay in registers while consecutive rows lie in the same parent cell (r_y - 1 rows out of r_y),
//    and when the row moves up one cell the upper four become the lower four: 8 / r_y .. 4 / r_y loads per target, not 8;
//  * for a target inside the parent (every distance >= -1e-4 h) the validity window needs one product: rounding is
//    monotonic, so max_a N_a = ((max ax * max ay) * max az) * inv_vol bit for bit, and min_a N_a >= -1e-4 holds by
//    construction.  Distances clamped at 0 beforehand give exactly the zeros the reference's clip(N, 0, 1) gives for
//    a target a hair outside its cell (x = c[e] - 1 ulp with floor() landing in e: one factor negative -> N <= 0 -> 0;
//    two such axes at once give N <= 1e-8 instead of 0: seen as 1-ulp differences at ~1e-4 of the nodes when parent and
//    target have the SAME spacing, never for a coarser parent - tests/test_transfer_gpu.py compares the two kernels bit for bit).
// Targets outside the parent (overhanging windows) take tri_eval as before.  At C2 size (10.3 M targets): T' = child -
// I(parent) 67.6 -> 57.4 us (ratio 2), 66.4 -> 51.2 us (ratio 5), the blended form 74.6 -> 56.2 us against the per-target 3-D
// kernel this replaced (profiles/r02_quick_interp.json); bound by instruction issue (71-78 % issue-active), not by DRAM.
struct AxC {
    int e;
    float c0, c1, mx;  // max(a0, 0), max(a1, 0), max(a0, a1)
    bool ok;           // inside the parent along this axis
};
__device__ __forceinline__ AxC axc_of(const AxisView& ax, float h, float x) {
    const Ax1 a = ax1_of(ax, h, x);
    AxC r;
    r.e = a.e;
    r.c0 = fmaxf(a.a0, 0.f);
    r.c1 = fmaxf(a.a1, 0.f);
    r.mx = fmaxf(a.a0, a.a1);
    r.ok = fminf(a.a0, a.a1) >= -1e-4f * h;
    return r;
}
__device__ __forceinline__ float4 row_pack(const AxC& Y) { return make_float4(__int_as_float(Y.ok ? Y.e : ~Y.e), Y.c0, Y.c1, Y.mx); }
// Parent values of one parent row at a column's x pair and z pair: (x0,z0), (x1,z0), (x0,z1), (x1,z1).
template <bool BLEND>
__device__ __forceinline__ void row4_load(float v[4], const float* __restrict__ u, const float* __restrict__ u2, float alpha, float beta,
                                          const SrcGeom& g, int bxz, int row) {
    const int b = bxz + row * g.nnx;
    v[0] = __ldg(u + b); v[1] = __ldg(u + b + 1); v[2] = __ldg(u + b + g.nnxy); v[3] = __ldg(u + b + 1 + g.nnxy);
    if (BLEND) {
        const float w0 = __ldg(u2 + b), w1 = __ldg(u2 + b + 1), w2 = __ldg(u2 + b + g.nnxy), w3 = __ldg(u2 + b + 1 + g.nnxy);
        v[0] = __fadd_rn(__fmul_rn(alpha, v[0]), __fmul_rn(beta, w0));
        v[1] = __fadd_rn(__fmul_rn(alpha, v[1]), __fmul_rn(beta, w1));
        v[2] = __fadd_rn(__fmul_rn(alpha, v[2]), __fmul_rn(beta, w2));
        v[3] = __fadd_rn(__fmul_rn(alpha, v[3]), __fmul_rn(beta, w3));
    }
}
struct Rows2 {   // the two parent rows of the current cell
    int E;
    float lo[4], hi[4];
};
template <bool BLEND>
__device__ __forceinline__ void rows_enter(Rows2& c, const float* __restrict__ u, const float* __restrict__ u2, float alpha, float beta,
                                           const SrcGeom& g, int bxz, int e) {
    if (e == c.E + 1) {
#pragma unroll
        for (int q = 0; q < 4; ++q) c.lo[q] = c.hi[q];
    } else {
        row4_load<BLEND>(c.lo, u, u2, alpha, beta, g, bxz, e);
    }
    row4_load<BLEND>(c.hi, u, u2, alpha, beta, g, bxz, e + 1);
    c.E = e;
}
__device__ __forceinline__ float tri_fast(const float lo[4], const float hi[4], float inv_vol, const AxC& X, float yc0, float yc1, float ymx,
                                          const AxC& Z) {
    const float xy00 = __fmul_rn(X.c0, yc0), xy10 = __fmul_rn(X.c1, yc0), xy11 = __fmul_rn(X.c1, yc1), xy01 = __fmul_rn(X.c0, yc1);
    const float top = __fmul_rn(__fmul_rn(__fmul_rn(X.mx, ymx), Z.mx), inv_vol);
    float acc = 0.f;   // hex8 local order, as tri_eval
    acc = __fadd_rn(acc, __fmul_rn(fminf(__fmul_rn(__fmul_rn(xy00, Z.c0), inv_vol), 1.f), lo[0]));
    acc = __fadd_rn(acc, __fmul_rn(fminf(__fmul_rn(__fmul_rn(xy10, Z.c0), inv_vol), 1.f), lo[1]));
    acc = __fadd_rn(acc, __fmul_rn(fminf(__fmul_rn(__fmul_rn(xy11, Z.c0), inv_vol), 1.f), hi[1]));
    acc = __fadd_rn(acc, __fmul_rn(fminf(__fmul_rn(__fmul_rn(xy01, Z.c0), inv_vol), 1.f), hi[0]));
    acc = __fadd_rn(acc, __fmul_rn(fminf(__fmul_rn(__fmul_rn(xy00, Z.c1), inv_vol), 1.f), lo[2]));
    acc = __fadd_rn(acc, __fmul_rn(fminf(__fmul_rn(__fmul_rn(xy10, Z.c1), inv_vol), 1.f), lo[3]));
    acc = __fadd_rn(acc, __fmul_rn(fminf(__fmul_rn(__fmul_rn(xy11, Z.c1), inv_vol), 1.f), hi[3]));
    acc = __fadd_rn(acc, __fmul_rn(fminf(__fmul_rn(__fmul_rn(xy01, Z.c1), inv_vol), 1.f), hi[2]));
    return top <= 1.0f + 1e-2f ? acc : 0.f;
}
// a target outside the parent: the plain evaluation, out of line (rare; keeps the unrolled row bodies small)
template <bool BLEND>
__device__ __noinline__ float tri_plain(const float* __restrict__ u, const float* __restrict__ u2, float alpha, float beta, AxisView sx,
                                        AxisView sy, AxisView sz, float x, float y, float z) {
    const SrcGeom g = geom_of(sx, sy, sz);
    return tri_eval<BLEND>(u, u2, alpha, beta, g, ax1_of(sx, g.hx, x), ax1_of(sy, g.hy, y), ax1_of(sz, g.hz, z));
}

constexpr int ROWSM = 8;   // target rows per thread of the marching forms

// What a block shares about one parent: the y part of its R rows, the z part of its plane, 1 / (hx hy hz) - formed by
// R + 2 threads of the block, read by all after one barrier (none of it depends on the column).
struct SrcShared {
    float4 row[ROWSM];
    float4 z;
    float inv_vol;
    float pad[3];
};
__device__ __forceinline__ void src_to_shared(SrcShared& s, int t, const AxisView& sx, const AxisView& sy, const AxisView& sz,
                                              const float* __restrict__ ty, int j0, int nrows, float z) {
    if (t < 0 || t > ROWSM + 1) return;
    if (t < ROWSM) {
        if (t < nrows) s.row[t] = row_pack(axc_of(sy, __fsub_rn(sy.c[1], sy.c[0]), ty[j0 + t]));
    } else if (t == ROWSM) {
        s.z = row_pack(axc_of(sz, __fsub_rn(sz.c[1], sz.c[0]), z));
    } else {
        s.inv_vol = geom_of(sx, sy, sz).inv_vol;
    }
}
__device__ __forceinline__ AxC row_unpack(const float4& r) {
    AxC a;
    const int e = __float_as_int(r.x);
    a.ok = e >= 0;
    a.e = a.ok ? e : ~e;
    a.c0 = r.y; a.c1 = r.z; a.mx = r.w;
    return a;
}

template <bool BLEND, int MODE>
__global__ void __launch_bounds__(256, 4) interp3_march_kernel(const InterpParams p, const int rpb) {   // rpb: rows per block (R, or 1 on small grids)
    constexpr int R = ROWSM;
    __shared__ SrcShared sh;
    const int j0 = blockIdx.y * rpb, j1 = min(p.nty, j0 + rpb);
    const int i = min((int)(blockIdx.x * blockDim.x + threadIdx.x), p.ntx - 1);   // the last block's spare threads shadow the last column
    const bool owner = (int)(blockIdx.x * blockDim.x + threadIdx.x) < p.ntx;
    const int k = blockIdx.z;
    const size_t o0 = ((size_t)k * p.nty + j0) * p.ntx + i;
    // everything that waits for memory is requested before the barrier: the field the result is combined with (all
    // rows), this column's x part, the block's shared parts
    float prev[R];
    if (MODE != GOMELT_INTERP_SET) {
        const float* __restrict__ b = MODE == GOMELT_INTERP_ADD ? p.out : p.base;
#pragma unroll
        for (int q = 0; q < R; ++q) prev[q] = j0 + q < j1 ? b[o0 + (size_t)q * p.ntx] : 0.f;
    }
    SrcGeom g;
    g.nnx = p.sx.n; g.nnxy = p.sx.n * p.sy.n; g.nny = p.sy.n;
    const AxC X = axc_of(p.sx, __fsub_rn(p.sx.c[1], p.sx.c[0]), p.tx[i]);
    src_to_shared(sh, threadIdx.x, p.sx, p.sy, p.sz, p.ty, j0, j1 - j0, p.tz[k]);
    __syncthreads();
    if (!owner) return;
    const AxC Z = row_unpack(sh.z);
    g.inv_vol = sh.inv_vol;
    const bool xz = X.ok && Z.ok;
    const int bxz = X.e + Z.e * g.nnxy;
    Rows2 c;
    c.E = -(1 << 30);
#pragma unroll
    for (int q = 0; q < R; ++q) {
        if (j0 + q >= j1) break;
        const float4 r = sh.row[q];
        const int e = __float_as_int(r.x);
        float acc;
        if (xz && e >= 0) {
            if (e != c.E) rows_enter<BLEND>(c, p.u, p.u2, p.alpha, p.beta, g, bxz, e);
            acc = tri_fast(c.lo, c.hi, g.inv_vol, X, r.y, r.z, r.w, Z);
        } else {
            acc = tri_plain<BLEND>(p.u, p.u2, p.alpha, p.beta, p.sx, p.sy, p.sz, p.tx[i], p.ty[j0 + q], p.tz[k]);
        }
        float res;
        if (MODE == GOMELT_INTERP_SET) res = acc;
        else if (MODE == GOMELT_INTERP_ADD) res = __fadd_rn(prev[q], acc);
        else res = __fsub_rn(prev[q], acc);
        if (p.has_clamp) res = fmaxf(res, p.clamp_min);
        p.out[o0 + (size_t)q * p.ntx] = res;
    }
}

__device__ __forceinline__ void face_node(int w, int ntx, int nty, int ntz, int& i, int& j, int& k);

// One thread per target node (x fastest => coalesced output; source reads hit L1/L2).
__global__ void interp_kernel(const InterpParams p) {
    const long long total = (long long)p.ntx * p.nty * p.ntz;
    const float hx = __fsub_rn(p.sx.c[1], p.sx.c[0]), hy = __fsub_rn(p.sy.c[1], p.sy.c[0]),
                hz = __fsub_rn(p.sz.c[1], p.sz.c[0]);
    const float inv_vol = __fdiv_rn(1.0f, __fmul_rn(__fmul_rn(hx, hy), hz));
    const int nnx = p.sx.n, nnxy = p.sx.n * p.sy.n;
    // faces_only: enumerate the nodes of the 5 Dirichlet faces directly (z- plane, then the y and x faces of
    // the planes above it) instead of visiting every node of the level
    const long long fA = (long long)p.ntx * p.nty, fB = 2LL * p.ntx * (p.ntz - 1),
                    fC = 2LL * (p.nty - 2) * (p.ntz - 1);
    const long long count = p.faces_only ? fA + fB + fC : total;
    // full-grid passes are launched with blockIdx.y / .z = target row / plane (grid3 = 1): no index divisions
    const bool grid3 = gridDim.y > 1 || gridDim.z > 1;
    const long long w0 = grid3 ? (long long)blockIdx.x * blockDim.x + threadIdx.x
                               : blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long wend = grid3 ? p.ntx : count;
    const long long wstep = grid3 ? (long long)gridDim.x * blockDim.x : (long long)gridDim.x * blockDim.x;
    for (long long w = w0; w < wend; w += wstep) {
        int i, j, k;
        if (grid3) {
            i = (int)w;
            j = blockIdx.y;
            k = blockIdx.z;
        } else if (p.faces_only && count <= 0x7fffffffLL) {
            face_node((int)w, p.ntx, p.nty, p.ntz, i, j, k);   // 32-bit index arithmetic
        } else if (!p.faces_only) {
            i = (int)(w % p.ntx);
            j = (int)((w / p.ntx) % p.nty);
            k = (int)(w / ((long long)p.ntx * p.nty));
        } else if (w < fA) {
            k = 0;
            j = (int)(w / p.ntx);
            i = (int)(w % p.ntx);
        } else if (w < fA + fB) {
            const long long u = w - fA;
            const int r = (int)(u % (2 * p.ntx));
            k = 1 + (int)(u / (2 * p.ntx));
            j = r < p.ntx ? 0 : p.nty - 1;
            i = r % p.ntx;
        } else {
            const long long u = w - fA - fB;
            const int r = (int)(u % (2 * (p.nty - 2)));
            k = 1 + (int)(u / (2 * (p.nty - 2)));
            i = (r & 1) ? p.ntx - 1 : 0;
            j = 1 + (r >> 1);
        }
        const long long t = i + (long long)j * p.ntx + (long long)k * p.ntx * p.nty;
        const float x = p.tx[i], y = p.ty[j], z = p.tz[k];
        const int ex = cell_of(x, p.sx.c[0], hx, p.sx.n - 1);
        const int ey = cell_of(y, p.sy.c[0], hy, p.sy.n - 1);
        const int ez = cell_of(z, p.sz.c[0], hz, p.sz.n - 1);
        const float x0 = p.sx.c[ex], x1 = p.sx.c[ex + 1];
        const float y0 = p.sy.c[ey], y1 = p.sy.c[ey + 1];
        const float z0 = p.sz.c[ez], z1 = p.sz.c[ez + 1];
        const float ax0 = __fsub_rn(x1, x), ax1 = __fsub_rn(x, x0);
        const float ay0 = __fsub_rn(y1, y), ay1 = __fsub_rn(y, y0);
        const float az0 = __fsub_rn(z1, z), az1 = __fsub_rn(z, z0);
        // the 8 parent values first, unconditionally (the cell is clamped into the parent): no branch before the loads
        const long long b = ex + (long long)ey * nnx + (long long)ez * nnxy;
        const long long nd[8] = {b, b + 1, b + 1 + nnx, b + nnx, b + nnxy, b + 1 + nnxy, b + 1 + nnx + nnxy,
                                 b + nnx + nnxy};
        float pv[8];
#pragma unroll
        for (int a = 0; a < 8; ++a) {
            pv[a] = p.u[nd[a]];
            if (p.u2) pv[a] = __fadd_rn(__fmul_rn(p.alpha, pv[a]), __fmul_rn(p.beta, p.u2[nd[a]]));
        }
        // compute3DN cF:1375-1391, hex8 local order
        float N[8];
        N[0] = __fmul_rn(__fmul_rn(__fmul_rn(ax0, ay0), az0), inv_vol);
        N[1] = __fmul_rn(__fmul_rn(__fmul_rn(ax1, ay0), az0), inv_vol);
        N[2] = __fmul_rn(__fmul_rn(__fmul_rn(ax1, ay1), az0), inv_vol);
        N[3] = __fmul_rn(__fmul_rn(__fmul_rn(ax0, ay1), az0), inv_vol);
        N[4] = __fmul_rn(__fmul_rn(__fmul_rn(ax0, ay0), az1), inv_vol);
        N[5] = __fmul_rn(__fmul_rn(__fmul_rn(ax1, ay0), az1), inv_vol);
        N[6] = __fmul_rn(__fmul_rn(__fmul_rn(ax1, ay1), az1), inv_vol);
        N[7] = __fmul_rn(__fmul_rn(__fmul_rn(ax0, ay1), az1), inv_vol);
        bool valid = true;
#pragma unroll
        for (int a = 0; a < 8; ++a) valid = valid && (N[a] >= -1e-2f) && (N[a] <= 1.0f + 1e-2f);
        float acc = 0.f;
#pragma unroll
        for (int a = 0; a < 8; ++a) acc = __fadd_rn(acc, __fmul_rn(fminf(fmaxf(N[a], 0.f), 1.f), pv[a]));
        if (!valid) acc = 0.f;
        long long o = t;
        if (p.mx) o = p.mx[i] + (long long)p.my[j] * p.map_nx + (long long)p.mz[k] * p.map_nx * p.map_ny;
        float r;
        if (p.mode == GOMELT_INTERP_SET) r = acc;
        else if (p.mode == GOMELT_INTERP_ADD) r = __fadd_rn(p.out[o], acc);
        else r = __fsub_rn(p.base[o], acc);
        if (p.has_clamp) r = fmaxf(r, p.clamp_min);
        p.out[o] = r;
    }
}

// ---- face prolongation split for the Level-3 inner scan --------------------------------------------------
// The N3 substeps of a subcycle block prolong the SAME two parent fields with different blend factors
// (alpha_i = (i+1)/N3, cF:3386-3389).  I(alpha u + beta u2) = alpha I(u) + beta I(u2), so both interpolants
// are gathered once per block into compact face arrays (the 16 scattered parent reads per face node happen
// once instead of N3 times) and every substep only blends and scatters them.
__device__ __forceinline__ void face_node(int w, int ntx, int nty, int ntz, int& i, int& j, int& k) {
    // same enumeration as interp_kernel's faces_only: z- plane, then the y faces and the x faces of the planes above
    const int fA = ntx * nty, fB = 2 * ntx * (ntz - 1);
    if (w < fA) {
        k = 0;
        j = w / ntx;
        i = w - j * ntx;
    } else if (w < fA + fB) {
        const int u = w - fA;
        const int q = u / (2 * ntx), r = u - q * (2 * ntx);
        k = 1 + q;
        j = r < ntx ? 0 : nty - 1;
        i = r < ntx ? r : r - ntx;
    } else {
        const int u = w - fA - fB;
        const int q = u / (2 * (nty - 2)), r = u - q * (2 * (nty - 2));
        k = 1 + q;
        i = (r & 1) ? ntx - 1 : 0;
        j = 1 + (r >> 1);
    }
}

__global__ void faces_gather_kernel(const InterpParams p, float* __restrict__ fa, float* __restrict__ fb, int nface) {
    const float hx = __fsub_rn(p.sx.c[1], p.sx.c[0]), hy = __fsub_rn(p.sy.c[1], p.sy.c[0]),
                hz = __fsub_rn(p.sz.c[1], p.sz.c[0]);
    const float inv_vol = __fdiv_rn(1.0f, __fmul_rn(__fmul_rn(hx, hy), hz));
    const int nnx = p.sx.n, nnxy = p.sx.n * p.sy.n;
    for (int w = blockIdx.x * blockDim.x + threadIdx.x; w < nface; w += gridDim.x * blockDim.x) {
        int i, j, k;
        face_node(w, p.ntx, p.nty, p.ntz, i, j, k);
        const float x = p.tx[i], y = p.ty[j], z = p.tz[k];
        const int ex = cell_of(x, p.sx.c[0], hx, p.sx.n - 1);
        const int ey = cell_of(y, p.sy.c[0], hy, p.sy.n - 1);
        const int ez = cell_of(z, p.sz.c[0], hz, p.sz.n - 1);
        const float ax0 = __fsub_rn(p.sx.c[ex + 1], x), ax1 = __fsub_rn(x, p.sx.c[ex]);
        const float ay0 = __fsub_rn(p.sy.c[ey + 1], y), ay1 = __fsub_rn(y, p.sy.c[ey]);
        const float az0 = __fsub_rn(p.sz.c[ez + 1], z), az1 = __fsub_rn(z, p.sz.c[ez]);
        float N[8];  // compute3DN cF:1375-1391, hex8 local order
        N[0] = __fmul_rn(__fmul_rn(__fmul_rn(ax0, ay0), az0), inv_vol);
        N[1] = __fmul_rn(__fmul_rn(__fmul_rn(ax1, ay0), az0), inv_vol);
        N[2] = __fmul_rn(__fmul_rn(__fmul_rn(ax1, ay1), az0), inv_vol);
        N[3] = __fmul_rn(__fmul_rn(__fmul_rn(ax0, ay1), az0), inv_vol);
        N[4] = __fmul_rn(__fmul_rn(__fmul_rn(ax0, ay0), az1), inv_vol);
        N[5] = __fmul_rn(__fmul_rn(__fmul_rn(ax1, ay0), az1), inv_vol);
        N[6] = __fmul_rn(__fmul_rn(__fmul_rn(ax1, ay1), az1), inv_vol);
        N[7] = __fmul_rn(__fmul_rn(__fmul_rn(ax0, ay1), az1), inv_vol);
        bool valid = true;
#pragma unroll
        for (int a = 0; a < 8; ++a) valid = valid && (N[a] >= -1e-2f) && (N[a] <= 1.0f + 1e-2f);
        const long long b = ex + (long long)ey * nnx + (long long)ez * nnxy;
        const long long nd[8] = {b, b + 1, b + 1 + nnx, b + nnx, b + nnxy, b + 1 + nnxy, b + 1 + nnx + nnxy,
                                 b + nnx + nnxy};
        float accA = 0.f, accB = 0.f;
        if (valid) {
#pragma unroll
            for (int a = 0; a < 8; ++a) {
                const float wt = fminf(fmaxf(N[a], 0.f), 1.f);
                accA = __fadd_rn(accA, __fmul_rn(wt, p.u[nd[a]]));
                accB = __fadd_rn(accB, __fmul_rn(wt, p.u2[nd[a]]));
            }
        }
        fa[w] = accA;
        fb[w] = accB;
    }
}

__global__ void faces_blend_kernel(const float* __restrict__ fa, const float* __restrict__ fb, int nface, int ntx, int nty,
                                   int ntz, float alpha, float beta, int has_clamp, float clamp_min,
                                   float* __restrict__ out) {
    for (int w = blockIdx.x * blockDim.x + threadIdx.x; w < nface; w += gridDim.x * blockDim.x) {
        int i, j, k;
        face_node(w, ntx, nty, ntz, i, j, k);
        float r = __fadd_rn(__fmul_rn(alpha, fa[w]), __fmul_rn(beta, fb[w]));
        if (has_clamp) r = fmaxf(r, clamp_min);
        out[(size_t)i + (size_t)j * ntx + (size_t)k * ntx * nty] = r;
    }
}

// ---- box gather / scatter between a window and a larger grid ----------------------------------
template <typename T>
__global__ void box_copy_kernel(const T* __restrict__ src, T* __restrict__ dst, const int* __restrict__ ix,
                                const int* __restrict__ iy, const int* __restrict__ iz, int nx, int ny, int nz,
                                int big_nx, int big_ny, int scatter) {
    // blockIdx.y / .z = window row / plane (launched that way when ny, nz <= 65535), else a flat grid-stride loop
    if (gridDim.y > 1 || gridDim.z > 1) {
        const int k = blockIdx.z;
        const int rpb = (ny + (int)gridDim.y - 1) / (int)gridDim.y;
        const int i = blockIdx.x * blockDim.x + threadIdx.x;   // (the 3-D launch covers x with blockIdx.x)
        if (i >= nx) return;
        const int xi = ix[i];
        const long long gz = (long long)iz[k] * big_nx * big_ny;
        if (rpb == ROWS3) {
            // a group of ROWS3 rows: all loads of the group in flight before the first store (the copy is latency-bound
            // otherwise: one element per thread and row)
            const int j0 = blockIdx.y * ROWS3;
            T v[ROWS3];
            long long go[ROWS3];
#pragma unroll
            for (int q = 0; q < ROWS3; ++q) {
                const int j = min(j0 + q, ny - 1);
                go[q] = (long long)iy[j] * big_nx + gz + xi;
                v[q] = scatter ? src[((long long)k * ny + j) * nx + i] : src[go[q]];
            }
#pragma unroll
            for (int q = 0; q < ROWS3; ++q) {
                if (j0 + q >= ny) break;
                if (scatter) dst[go[q]] = v[q];
                else dst[((long long)k * ny + j0 + q) * nx + i] = v[q];
            }
            return;
        }
        for (int j = blockIdx.y * rpb; j < min(ny, (int)(blockIdx.y + 1) * rpb); ++j) {
            const long long gb = (long long)iy[j] * big_nx + gz;
            const long long tb = ((long long)k * ny + j) * nx;
            if (scatter) dst[gb + xi] = src[tb + i];
            else dst[tb + i] = src[gb + xi];
        }
        return;
    }
    const long long total = (long long)nx * ny * nz;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(t % nx);
        const int j = (int)((t / nx) % ny);
        const int k = (int)(t / ((long long)nx * ny));
        const long long g = ix[i] + (long long)iy[j] * big_nx + (long long)iz[k] * big_nx * big_ny;
        if (scatter) dst[g] = src[t];
        else dst[t] = src[g];
    }
}

// ---- rank-1 accumulate: F[n] (+)= coef * tx[ix] * ty[iy] * tz[iz] ------------------------------------
__global__ void rank1_kernel(float* __restrict__ F, const float* __restrict__ tx, const float* __restrict__ ty,
                             const float* __restrict__ tz, int nx, int ny, int nz, float coef, int accumulate) {
    const long long total = (long long)nx * ny * nz;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(t % nx);
        const int j = (int)((t / nx) % ny);
        const int k = (int)(t / ((long long)nx * ny));
        const float v = coef * ((tx[i] * ty[j]) * tz[k]);
        F[t] = accumulate ? F[t] + v : v;
    }
}

// ---- K6 for parent levels: 1-D table of the projected Gaussian ------------------------------------------
// t[ic] = sum over fine Gauss points xq inside the parent cells adjacent to parent node ic of
//         hat_ic(xq) * c * exp(-3 (xq - v)^2 / s^2),   hat = the parent's 1-D shape function evaluated
// like compute3DN's factors ((x1 - xq)/h or (xq - x0)/h), parent cell by the floor rule.
// (computeSources cF:928-988 / computeLevelSource cF:2667-2730 are separable: Nc and Q are tensor products.)
__global__ void coarse_source_table_kernel(const float* __restrict__ xf, int nf, const float* __restrict__ xc, int nc,
                                           float v, float inv_s2, float c, float* __restrict__ t) {
    const int ic = blockIdx.x * blockDim.x + threadIdx.x;
    if (ic >= nc) return;
    const float g = 0.57735026918962576f;
    const float Nlo = 0.5f * (1.f + g), Nhi = 0.5f * (1.f - g);
    const float hc = xc[1] - xc[0];
    const float inv_hc = 1.0f / hc;
    float acc = 0.f;
    for (int e = 0; e < nf - 1; ++e) {
        const float x0 = xf[e], x1 = xf[e + 1];
        const float xq0 = Nlo * x0 + Nhi * x1, xq1 = Nhi * x0 + Nlo * x1;
        // parent cell of the element = cell of its first Gauss point (cF:1351)
        const int ec = cell_of(xq0, xc[0], hc, nc - 1);
        if (ec != ic && ec + 1 != ic) continue;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const float xq = q == 0 ? xq0 : xq1;
            // shape values of the cell that holds this Gauss point, scattered to the element's cell (see project_cells_kernel)
            const int eq = q == 0 ? ec : cell_of(xq1, xc[0], hc, nc - 1);
            const float xc0 = xc[eq], xc1 = xc[eq + 1];
            const float w = (ec == ic) ? (xc1 - xq) * inv_hc : (xq - xc0) * inv_hc;
            const float d = xq - v;
            acc += w * (c * expf(-3.f * d * d * inv_s2));
        }
    }
    t[ic] = acc;
}

// ---- K3: correction vectors -----------------------------------------------------------------------------
// Per parent cell: sum over its fine elements and their 8 Gauss points of
//   GRAD:  - wq * sum_d dNc_d[q,c] * kbar * (dN_d[q,:] . A)          (cF:1477-1565, A = T'0)
//   MASS:  - wq/dt * Nc[q,c] * (N[q,:] . A) * cbar                    (cF:1396-1474, A = T'new - T'old)
// for the 8 parent nodes c of the cell -> cellsum[cell][8]; a second kernel sums, per parent node,
// the <= 8 adjacent cells in a fixed order (deterministic, no float atomics).
struct ProjParams {
    AxisView fx, fy, fz;   // fine level
    AxisView cx, cy, cz;   // parent level
    const float* A;        // fine field
    const float* A2;       // optional: field = A - A2
    const float* coef;     // fine nodal k (GRAD) or rho*cp (MASS), or nullptr: evaluated from (cT, cS1)
    const float* cT;       // fine temperature / state the coefficient is evaluated from (computeStateProperties)
    const float* cS1;
    long long cnsub;
    PropK pk;
    int mode;              // 0 = GRAD, 1 = MASS
    float scale;           // MASS: 1/dt
    // parent-cell box that contains fine elements, and the fine-element range of each parent cell
    int c0x, c0y, c0z, ncx, ncy, ncz;
    const int* fsx;        // [ncx+1] first fine element (x) of parent cell c0x + i
    const int* fsy;
    const int* fsz;
    float* cellsum;        // [ncx*ncy*ncz][8]
};

__device__ __forceinline__ void hat_factors(float xq, const float* xc, int ec, float& w0, float& w1) {
    w0 = __fsub_rn(xc[ec + 1], xq);  // (x1 - xq)  -> parent node ec
    w1 = __fsub_rn(xq, xc[ec]);      // (xq - x0)  -> parent node ec + 1
}

template <int G>  // lanes per parent cell (8 or 32)
__global__ void project_cells_kernel(const ProjParams p) {
    const int lane = threadIdx.x % G;
    const long long cell = (blockIdx.x * (long long)blockDim.x + threadIdx.x) / G;
    const long long ncell = (long long)p.ncx * p.ncy * p.ncz;
    const bool active = cell < ncell;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (active) {
        const int ci = (int)(cell % p.ncx), cj = (int)((cell / p.ncx) % p.ncy), ck = (int)(cell / ((long long)p.ncx * p.ncy));
        const int ex0 = p.fsx[ci], ex1 = p.fsx[ci + 1];
        const int ey0 = p.fsy[cj], ey1 = p.fsy[cj + 1];
        const int ez0 = p.fsz[ck], ez1 = p.fsz[ck + 1];
        const int nex = ex1 - ex0, ney = ey1 - ey0, nez = ez1 - ez0;
        const int nel = nex * ney * nez;
        const float hfx = __fsub_rn(p.fx.c[1], p.fx.c[0]), hfy = __fsub_rn(p.fy.c[1], p.fy.c[0]),
                    hfz = __fsub_rn(p.fz.c[1], p.fz.c[0]);
        const float hcx = __fsub_rn(p.cx.c[1], p.cx.c[0]), hcy = __fsub_rn(p.cy.c[1], p.cy.c[0]),
                    hcz = __fsub_rn(p.cz.c[1], p.cz.c[0]);
        const float inv_cvol = __fdiv_rn(1.0f, __fmul_rn(__fmul_rn(hcx, hcy), hcz));
        const float wq = (hfx * hfy * hfz) * 0.125f;
        const float g = 0.57735026918962576f;
        const float Nlo = 0.5f * (1.f + g), Nhi = 0.5f * (1.f - g);
        const float gdx = 2.0f / hfx, gdy = 2.0f / hfy, gdz = 2.0f / hfz;  // d(xi)/dx
        const int fnx = p.fx.n, fnxy = p.fx.n * p.fy.n;
        for (int t = lane; t < nel; t += G) {
            const int ex = ex0 + t % nex, ey = ey0 + (t / nex) % ney, ez = ez0 + t / (nex * ney);
            const long long b = ex + (long long)ey * fnx + (long long)ez * fnxy;
            const long long nd[8] = {b, b + 1, b + 1 + fnx, b + fnx, b + fnxy, b + 1 + fnxy, b + 1 + fnx + fnxy,
                                     b + fnx + fnxy};
            float a[8], cbar = 0.f;
#pragma unroll
            for (int n = 0; n < 8; ++n) {
                a[n] = p.A[nd[n]];
                if (p.A2) a[n] -= p.A2[nd[n]];
                if (p.coef) {
                    cbar += p.coef[nd[n]];
                } else {
                    float kk, rr;
                    bool b1, b2;
                    node_props(p.pk, p.cT[nd[n]], p.cS1[nd[n]], nd[n] < p.cnsub, kk, rr, b1, b2);
                    cbar += p.mode == 1 ? rr : kk;
                }
            }
            cbar *= 0.125f;
            // corner values in (x,y,z)-bit order for separable evaluation: v[bx][by][bz]
            const float v000 = a[0], v100 = a[1], v110 = a[2], v010 = a[3], v001 = a[4], v101 = a[5], v111 = a[6],
                        v011 = a[7];
            const float xf0 = p.fx.c[ex], xf1 = p.fx.c[ex + 1];
            const float yf0 = p.fy.c[ey], yf1 = p.fy.c[ey + 1];
            const float zf0 = p.fz.c[ez], zf1 = p.fz.c[ez + 1];
#pragma unroll
            for (int qz = 0; qz < 2; ++qz) {
                const float zq = qz == 0 ? Nlo * zf0 + Nhi * zf1 : Nhi * zf0 + Nlo * zf1;
                const float sz0 = qz == 0 ? Nlo : Nhi, sz1 = qz == 0 ? Nhi : Nlo;  // fine 1-D shape values
                // the parent's shape functions are those of the cell that holds THIS Gauss point (cF:1283-1335); their
                // values are scattered to the nodes of the element's cell (Gauss point 0, cF:1351).  The two cells are the
                // same whenever the fine grid nests in the parent (every GO-MELT configuration: integer element ratios).
                float cz0, cz1;
                hat_factors(zq, p.cz.c, cell_of(zq, p.cz.c[0], hcz, p.cz.n - 1), cz0, cz1);
#pragma unroll
                for (int qy = 0; qy < 2; ++qy) {
                    const float yq = qy == 0 ? Nlo * yf0 + Nhi * yf1 : Nhi * yf0 + Nlo * yf1;
                    const float sy0 = qy == 0 ? Nlo : Nhi, sy1 = qy == 0 ? Nhi : Nlo;
                    float cy0, cy1;
                    hat_factors(yq, p.cy.c, cell_of(yq, p.cy.c[0], hcy, p.cy.n - 1), cy0, cy1);
#pragma unroll
                    for (int qx = 0; qx < 2; ++qx) {
                        const float xq = qx == 0 ? Nlo * xf0 + Nhi * xf1 : Nhi * xf0 + Nlo * xf1;
                        const float sx0 = qx == 0 ? Nlo : Nhi, sx1 = qx == 0 ? Nhi : Nlo;
                        float cx0, cx1;
                        hat_factors(xq, p.cx.c, cell_of(xq, p.cx.c[0], hcx, p.cx.n - 1), cx0, cx1);
                        // fine-side interpolants at this Gauss point
                        const float e00 = sx0 * v000 + sx1 * v100, e10 = sx0 * v010 + sx1 * v110;
                        const float e01 = sx0 * v001 + sx1 * v101, e11 = sx0 * v011 + sx1 * v111;
                        const float d00 = v100 - v000, d10 = v110 - v010, d01 = v101 - v001, d11 = v111 - v011;
                        if (p.mode == 1) {
                            const float val = sz0 * (sy0 * e00 + sy1 * e10) + sz1 * (sy0 * e01 + sy1 * e11);
                            const float s = -(p.scale * wq) * (val * cbar);
                            // Nc[q,c] = ((fx * fy) * fz) * inv_vol, hex8 local order
                            acc[0] += s * (((cx0 * cy0) * cz0) * inv_cvol);
                            acc[1] += s * (((cx1 * cy0) * cz0) * inv_cvol);
                            acc[2] += s * (((cx1 * cy1) * cz0) * inv_cvol);
                            acc[3] += s * (((cx0 * cy1) * cz0) * inv_cvol);
                            acc[4] += s * (((cx0 * cy0) * cz1) * inv_cvol);
                            acc[5] += s * (((cx1 * cy0) * cz1) * inv_cvol);
                            acc[6] += s * (((cx1 * cy1) * cz1) * inv_cvol);
                            acc[7] += s * (((cx0 * cy1) * cz1) * inv_cvol);
                        } else {
                            // grad A at the Gauss point (fine shape-function derivatives, diagonal Jacobian)
                            const float gx = 0.5f * gdx * (sz0 * (sy0 * d00 + sy1 * d10) + sz1 * (sy0 * d01 + sy1 * d11));
                            const float gy = 0.5f * gdy * (sz0 * (e10 - e00) + sz1 * (e11 - e01));
                            const float gz = 0.5f * gdz * ((sy0 * e01 + sy1 * e11) - (sy0 * e00 + sy1 * e10));
                            const float fxk = -wq * cbar * gx * inv_cvol, fyk = -wq * cbar * gy * inv_cvol,
                                        fzk = -wq * cbar * gz * inv_cvol;
                            // dNc/dx = -+ (fy * fz) / vol etc. (cF:1290-1335)
                            acc[0] += fxk * (-(cy0 * cz0)) + fyk * (-(cx0 * cz0)) + fzk * (-(cx0 * cy0));
                            acc[1] += fxk * (cy0 * cz0) + fyk * (-(cx1 * cz0)) + fzk * (-(cx1 * cy0));
                            acc[2] += fxk * (cy1 * cz0) + fyk * (cx1 * cz0) + fzk * (-(cx1 * cy1));
                            acc[3] += fxk * (-(cy1 * cz0)) + fyk * (cx0 * cz0) + fzk * (-(cx0 * cy1));
                            acc[4] += fxk * (-(cy0 * cz1)) + fyk * (-(cx0 * cz1)) + fzk * (cx0 * cy0);
                            acc[5] += fxk * (cy0 * cz1) + fyk * (-(cx1 * cz1)) + fzk * (cx1 * cy0);
                            acc[6] += fxk * (cy1 * cz1) + fyk * (cx1 * cz1) + fzk * (cx1 * cy1);
                            acc[7] += fxk * (-(cy1 * cz1)) + fyk * (cx0 * cz1) + fzk * (cx0 * cy1);
                        }
                    }
                }
            }
        }
    }
    // segmented (G-lane) butterfly reduction: fixed order, deterministic
#pragma unroll
    for (int off = G / 2; off >= 1; off >>= 1) {
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], off, G);
    }
    if (active && lane == 0) {
#pragma unroll
        for (int c = 0; c < 8; ++c) p.cellsum[cell * 8 + c] = acc[c];
    }
}


// ---- K3, tiled form (project_tile_kernel): the fast path of gomelt_project_f32 -------------------------------------------
// Same sums as project_cells_kernel, organised for the 10 M-node windows: a CTA owns a box of parent cells, stages the
// fine nodes under it ONCE in shared memory - field value (A - A2) and nodal coefficient, the latter evaluated there from
// (T, S1) with computeStateProperties when no coefficient array is given - and then walks the fine elements out of
// shared memory.  The parent's shape-function factors at the two Gauss points of every fine element come from three
// 1-D tables built once per window position (wtab_[xyz][e] = (x1 - xq0, xq0 - x0, x1 - xq1, xq1 - x0) of the parent cell
// that holds each Gauss point), and both terms are evaluated by sum factorisation (the hex8 shape functions and their
// parent counterparts are tensor products; only the element-mean coefficient couples the axes):
//   MASS  val[q] = N[q,:] . A_e by three 2 x 2 stages, times cbar, then back through the parent factors by three stages;
//   GRAD  dA/dx at a Gauss point does not depend on its x index (trilinear), so each direction is a 2-D transform of the
//         four edge differences and the sum over the third Gauss index is a factor 2.
// ~100 FP operations per fine element instead of ~450, every fine node read once per CTA instead of eight times per
// element.  Partial sums of the threads that share a parent cell are added in a fixed order: deterministic.
struct TileParams {
    int fnx, fny, fnz;           // fine nodes
    const float* A;
    const float* A2;
    const float* coef;
    const float* cT;
    const float* cS1;
    long long cnsub;
    PropK pk;
    int mode;
    float sw;                    // MASS: scale * wq ; GRAD: wq
    float inv_hfx, inv_hfy, inv_hfz, inv_cvol;
    const float4* wx;            // [fine elements per axis]
    const float4* wy;
    const float4* wz;
    const int* fsx;              // first fine element of parent cell c0 + i
    const int* fsy;
    const int* fsz;
    int ncx, ncy, ncz;           // parent cells with fine elements
    int pcx, pcy, pcz;           // parent cells per CTA
    int G;                       // threads per parent cell (power of two)
    float* cellsum;
};
// floor(n / d) for the small non-negative integers of a tile (n < 4096, d <= 256) without an integer division:
// n * (1/d) carries a relative error of 1e-7, the half-step offset keeps the truncation on the right side
__device__ __forceinline__ int fastdiv(int n, float inv_d, float half_step) { return (int)(__int2float_rn(n) * inv_d + half_step); }
constexpr int TILE_NODECAP = 2304;
constexpr int TILE_THREADS = 256;

__global__ void __launch_bounds__(TILE_THREADS) project_tile_kernel(const TileParams p) {
    __shared__ float2 s_node[TILE_NODECAP];          // (A - A2, coefficient)
    __shared__ float s_part[TILE_THREADS][9];        // per-thread partial sums (padded)
    const int tid = threadIdx.x;
    const int ci0 = blockIdx.x * p.pcx, cj0 = blockIdx.y * p.pcy, ck0 = blockIdx.z * p.pcz;
    const int ci1 = min(ci0 + p.pcx, p.ncx), cj1 = min(cj0 + p.pcy, p.ncy), ck1 = min(ck0 + p.pcz, p.ncz);
    const int ex0 = p.fsx[ci0], ey0 = p.fsy[cj0], ez0 = p.fsz[ck0];
    const int bnx = p.fsx[ci1] - ex0 + 1, bny = p.fsy[cj1] - ey0 + 1, bnz = p.fsz[ck1] - ez0 + 1;  // node box
    const int bn = bnx * bny * bnz;
    const long long fnxy = (long long)p.fnx * p.fny;
    const int bxy = bnx * bny;
    const float inv_bxy = 1.0f / (float)bxy, hs_bxy = 0.5f * inv_bxy, inv_bnx = 1.0f / (float)bnx, hs_bnx = 0.5f * inv_bnx;
    for (int n = tid; n < bn; n += TILE_THREADS) {
        const int k = fastdiv(n, inv_bxy, hs_bxy), r = n - k * bxy, j = fastdiv(r, inv_bnx, hs_bnx), i = r - j * bnx;
        const long long g = (ex0 + i) + (long long)(ey0 + j) * p.fnx + (long long)(ez0 + k) * fnxy;
        float a = __ldg(p.A + g);
        if (p.A2) a -= __ldg(p.A2 + g);
        float c;
        if (p.coef) {
            c = __ldg(p.coef + g);
        } else {
            float kk, rr;
            bool b1, b2;
            node_props(p.pk, __ldg(p.cT + g), __ldg(p.cS1 + g), g < p.cnsub, kk, rr, b1, b2);
            c = p.mode == 1 ? rr : kk;
        }
        s_node[n] = make_float2(a, c);
    }
    __syncthreads();
    const int tcx = ci1 - ci0, tcy = cj1 - cj0, tcz = ck1 - ck0;
    const int ncell_t = tcx * tcy * tcz;
    const int cl = tid / p.G, gl = tid - cl * p.G;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (cl < ncell_t) {
        const int lk = cl / (tcx * tcy), lr = cl - lk * (tcx * tcy), lj = lr / tcx, li = lr - lj * tcx;
        const int fx0 = p.fsx[ci0 + li], fx1 = p.fsx[ci0 + li + 1];
        const int fy0 = p.fsy[cj0 + lj], fy1 = p.fsy[cj0 + lj + 1];
        const int fz0 = p.fsz[ck0 + lk], fz1 = p.fsz[ck0 + lk + 1];
        const int nex = fx1 - fx0, ney = fy1 - fy0, nel = nex * ney * (fz1 - fz0);
        const int nexy = nex * ney;
        const float inv_nexy = 1.0f / (float)nexy, hs_nexy = 0.5f * inv_nexy, inv_nex = 1.0f / (float)nex, hs_nex = 0.5f * inv_nex;
        const float g3 = 0.57735026918962576f;
        const float Nlo = 0.5f * (1.f + g3), Nhi = 0.5f * (1.f - g3);
        for (int t = gl; t < nel; t += p.G) {
            const int ez = fastdiv(t, inv_nexy, hs_nexy), tr = t - ez * nexy, ey = fastdiv(tr, inv_nex, hs_nex), ex = tr - ey * nex;
            const int gx_ = fx0 + ex, gy_ = fy0 + ey, gz_ = fz0 + ez;  // global fine element
            const int b = (gx_ - ex0) + (gy_ - ey0) * bnx + (gz_ - ez0) * bnx * bny;
            // corner (bx, by, bz) at b + bx + by * bnx + bz * bnx * bny
            float a[2][2][2];
            float csum = 0.f;
#pragma unroll
            for (int bz = 0; bz < 2; ++bz)
#pragma unroll
                for (int by = 0; by < 2; ++by)
#pragma unroll
                    for (int bx = 0; bx < 2; ++bx) {
                        const float2 v = s_node[b + bx + by * bnx + bz * bnx * bny];
                        a[bx][by][bz] = v.x;
                        csum += v.y;
                    }
            const float cbar = csum * 0.125f;
            const float4 tx = __ldg(p.wx + gx_), ty = __ldg(p.wy + gy_), tz = __ldg(p.wz + gz_);
            const float wX[2][2] = {{tx.x, tx.y}, {tx.z, tx.w}};  // [Gauss point][parent corner]
            const float wY[2][2] = {{ty.x, ty.y}, {ty.z, ty.w}};
            const float wZ[2][2] = {{tz.x, tz.y}, {tz.z, tz.w}};
            const float Nq[2][2] = {{Nlo, Nhi}, {Nhi, Nlo}};       // fine 1-D shape values [Gauss point][corner]
            float out[2][2][2] = {{{0.f, 0.f}, {0.f, 0.f}}, {{0.f, 0.f}, {0.f, 0.f}}};  // [cx][cy][cz]
            if (p.mode == 1) {
                float fx[2][2][2], fy[2][2][2], s[2][2][2];
#pragma unroll
                for (int q = 0; q < 2; ++q)
#pragma unroll
                    for (int by = 0; by < 2; ++by)
#pragma unroll
                        for (int bz = 0; bz < 2; ++bz) fx[q][by][bz] = Nq[q][0] * a[0][by][bz] + Nq[q][1] * a[1][by][bz];
#pragma unroll
                for (int qx = 0; qx < 2; ++qx)
#pragma unroll
                    for (int q = 0; q < 2; ++q)
#pragma unroll
                        for (int bz = 0; bz < 2; ++bz) fy[qx][q][bz] = Nq[q][0] * fx[qx][0][bz] + Nq[q][1] * fx[qx][1][bz];
                const float f = -(p.sw * cbar);
#pragma unroll
                for (int qx = 0; qx < 2; ++qx)
#pragma unroll
                    for (int qy = 0; qy < 2; ++qy)
#pragma unroll
                        for (int q = 0; q < 2; ++q) s[qx][qy][q] = f * (Nq[q][0] * fy[qx][qy][0] + Nq[q][1] * fy[qx][qy][1]);
                float u[2][2][2], v[2][2][2];
#pragma unroll
                for (int cx = 0; cx < 2; ++cx)
#pragma unroll
                    for (int qy = 0; qy < 2; ++qy)
#pragma unroll
                        for (int qz = 0; qz < 2; ++qz) u[cx][qy][qz] = wX[0][cx] * s[0][qy][qz] + wX[1][cx] * s[1][qy][qz];
#pragma unroll
                for (int cx = 0; cx < 2; ++cx)
#pragma unroll
                    for (int cy = 0; cy < 2; ++cy)
#pragma unroll
                        for (int qz = 0; qz < 2; ++qz) v[cx][cy][qz] = wY[0][cy] * u[cx][0][qz] + wY[1][cy] * u[cx][1][qz];
#pragma unroll
                for (int cx = 0; cx < 2; ++cx)
#pragma unroll
                    for (int cy = 0; cy < 2; ++cy)
#pragma unroll
                        for (int cz = 0; cz < 2; ++cz) out[cx][cy][cz] = (wZ[0][cz] * v[cx][cy][0] + wZ[1][cz] * v[cx][cy][1]) * p.inv_cvol;
            } else {
                const float f = -(p.sw * cbar) * p.inv_cvol * 2.0f;  // the Gauss index along the derivative sums to a factor 2
                // x: edge differences d[by][bz], interpolated to (qy, qz), back through wY, wZ, sign by cx
                {
                    float d[2][2], gq[2][2], t1[2][2];
#pragma unroll
                    for (int by = 0; by < 2; ++by)
#pragma unroll
                        for (int bz = 0; bz < 2; ++bz) d[by][bz] = a[1][by][bz] - a[0][by][bz];
#pragma unroll
                    for (int qy = 0; qy < 2; ++qy)
#pragma unroll
                        for (int qz = 0; qz < 2; ++qz)
                            gq[qy][qz] = (f * p.inv_hfx) * (Nq[qz][0] * (Nq[qy][0] * d[0][0] + Nq[qy][1] * d[1][0]) +
                                                            Nq[qz][1] * (Nq[qy][0] * d[0][1] + Nq[qy][1] * d[1][1]));
#pragma unroll
                    for (int cy = 0; cy < 2; ++cy)
#pragma unroll
                        for (int qz = 0; qz < 2; ++qz) t1[cy][qz] = wY[0][cy] * gq[0][qz] + wY[1][cy] * gq[1][qz];
#pragma unroll
                    for (int cy = 0; cy < 2; ++cy)
#pragma unroll
                        for (int cz = 0; cz < 2; ++cz) {
                            const float o = wZ[0][cz] * t1[cy][0] + wZ[1][cz] * t1[cy][1];
                            out[0][cy][cz] -= o;
                            out[1][cy][cz] += o;
                        }
                }
                {  // y: d[bx][bz], interpolated to (qx, qz), back through wX, wZ
                    float d[2][2], gq[2][2], t1[2][2];
#pragma unroll
                    for (int bx = 0; bx < 2; ++bx)
#pragma unroll
                        for (int bz = 0; bz < 2; ++bz) d[bx][bz] = a[bx][1][bz] - a[bx][0][bz];
#pragma unroll
                    for (int qx = 0; qx < 2; ++qx)
#pragma unroll
                        for (int qz = 0; qz < 2; ++qz)
                            gq[qx][qz] = (f * p.inv_hfy) * (Nq[qz][0] * (Nq[qx][0] * d[0][0] + Nq[qx][1] * d[1][0]) +
                                                            Nq[qz][1] * (Nq[qx][0] * d[0][1] + Nq[qx][1] * d[1][1]));
#pragma unroll
                    for (int cx = 0; cx < 2; ++cx)
#pragma unroll
                        for (int qz = 0; qz < 2; ++qz) t1[cx][qz] = wX[0][cx] * gq[0][qz] + wX[1][cx] * gq[1][qz];
#pragma unroll
                    for (int cx = 0; cx < 2; ++cx)
#pragma unroll
                        for (int cz = 0; cz < 2; ++cz) {
                            const float o = wZ[0][cz] * t1[cx][0] + wZ[1][cz] * t1[cx][1];
                            out[cx][0][cz] -= o;
                            out[cx][1][cz] += o;
                        }
                }
                {  // z: d[bx][by], interpolated to (qx, qy), back through wX, wY
                    float d[2][2], gq[2][2], t1[2][2];
#pragma unroll
                    for (int bx = 0; bx < 2; ++bx)
#pragma unroll
                        for (int by = 0; by < 2; ++by) d[bx][by] = a[bx][by][1] - a[bx][by][0];
#pragma unroll
                    for (int qx = 0; qx < 2; ++qx)
#pragma unroll
                        for (int qy = 0; qy < 2; ++qy)
                            gq[qx][qy] = (f * p.inv_hfz) * (Nq[qy][0] * (Nq[qx][0] * d[0][0] + Nq[qx][1] * d[1][0]) +
                                                            Nq[qy][1] * (Nq[qx][0] * d[0][1] + Nq[qx][1] * d[1][1]));
#pragma unroll
                    for (int cx = 0; cx < 2; ++cx)
#pragma unroll
                        for (int qy = 0; qy < 2; ++qy) t1[cx][qy] = wX[0][cx] * gq[0][qy] + wX[1][cx] * gq[1][qy];
#pragma unroll
                    for (int cx = 0; cx < 2; ++cx)
#pragma unroll
                        for (int cy = 0; cy < 2; ++cy) {
                            const float o = wY[0][cy] * t1[cx][0] + wY[1][cy] * t1[cx][1];
                            out[cx][cy][0] -= o;
                            out[cx][cy][1] += o;
                        }
                }
            }
            // hex8 local order: 0 (000) 1 (100) 2 (110) 3 (010) 4 (001) 5 (101) 6 (111) 7 (011), bits = (x, y, z)
            acc[0] += out[0][0][0]; acc[1] += out[1][0][0]; acc[2] += out[1][1][0]; acc[3] += out[0][1][0];
            acc[4] += out[0][0][1]; acc[5] += out[1][0][1]; acc[6] += out[1][1][1]; acc[7] += out[0][1][1];
        }
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) s_part[tid][c] = acc[c];
    __syncthreads();
    // one thread per (parent cell of the tile, corner): the G partial sums in thread order
    for (int o = tid; o < ncell_t * 8; o += TILE_THREADS) {
        const int cell = o >> 3, c = o & 7;
        float s = 0.f;
        for (int g = 0; g < p.G; ++g) s += s_part[cell * p.G + g][c];
        const int lk = fastdiv(cell, 1.0f / (float)(tcx * tcy), 0.5f / (float)(tcx * tcy)), lr = cell - lk * (tcx * tcy);
        const int lj = fastdiv(lr, 1.0f / (float)tcx, 0.5f / (float)tcx), li = lr - lj * tcx;
        const long long gc = (ci0 + li) + (long long)(cj0 + lj) * p.ncx + (long long)(ck0 + lk) * p.ncx * p.ncy;
        p.cellsum[gc * 8 + c] = s;
    }
}

// ---- K3, marching form (project_march_kernel): the fast path for nested windows ------------------------------------------
// Every window pair of a GO-MELT run nests with an integer element ratio in x and y (r_x, r_y fine elements per parent
// cell; in z the counts may differ from cell to cell: whole powder layers shift the windows, so z keeps the fsz table).
// There the tile kernel's ~400 instructions per fine node (staging, index arithmetic, 8 shared-memory reads and ~150
// FP operations per element, a serial in-CTA reduction) are the whole cost of a projection - it runs at 8 % of the HBM
// roofline.  This form has no staging and no synchronisation:
//   * a warp owns epw = floor(31 / r_x) r_x element columns x RY element rows (RY a multiple of r_y) and marches in z
//     through a chunk of parent cells; lane l loads node column x0 + l of the RY + 1 node rows of the next plane
//     (coalesced), evaluates the nodal coefficient there once, and gets column l + 1 by one shuffle per row and field;
//   * with M_d[c][b] = sum_q W_d[q][c] N[q][b] (the parent's hat factor times the fine shape value, summed over the two
//     Gauss points of the axis: a 2 x 2 matrix per fine element and axis, formed from the wtab entries) both terms are
//     tensor products of 1-D operators on the element's 2 x 2 x 2 corner values,
//         MASS  out[c] = f cbar (Mx (x) My (x) Mz) a,       GRAD  out[c] = f kbar (Dx (x) My (x) Mz + Mx (x) Dy (x) Mz + Mx (x) My (x) Dz) a
//     (D = corner difference with the sign of the parent corner; the Gauss sum along a derivative axis is the factor 2),
//     so everything that involves one node plane only (the x / y stages) is computed once per plane and shared by the
//     element below and the element above, and the per-element work is the z stage, the element-mean coefficient and
//     12 (GRAD: per term, before the signs) or 8 (MASS) accumulations - ~50 FP operations per element instead of ~150;
//   * the sums of a parent cell stay in registers until the march leaves the cell (fsz), then the r_x lanes of the cell
//     are added by a fixed shuffle tree and one lane stores the 8 corner sums: deterministic, no atomics.
struct MarchParams {
    int fnx, fny, fnz;
    const float* A;
    const float* A2;
    const float* coef;
    const float* cT;
    const float* cS1;
    long long cnsub;
    PropK pk;
    float sw;                    // MASS: scale * wq ; GRAD: wq
    float inv_hfx, inv_hfy, inv_hfz, inv_cvol;
    const float4* wx;
    const float4* wy;
    const float4* wz;
    const int* fsz;              // first fine element (z) of parent cell c0z + i
    int ncx, ncy, ncz;
    int rx, ry;                  // fine elements per parent cell in x, y (uniform)
    int offx, offy;              // fine elements missing from the first parent cell (a window that starts inside a cell)
    int epw;                     // element columns per warp: (31 / rx) * rx
    int czn;                     // parent z-cells per chunk (blockIdx.z)
    float* cellsum;
};
constexpr int MARCH_WARPS = 4;

struct M22 { float m00, m01, m10, m11; };  // M[c][b]
__device__ __forceinline__ M22 m_of(const float4 w) {  // w = (W[q0][c0], W[q0][c1], W[q1][c0], W[q1][c1]); N[q0] = (lo, hi), N[q1] = (hi, lo)
    const float g3 = 0.57735026918962576f;
    const float lo = 0.5f * (1.f + g3), hi = 0.5f * (1.f - g3);
    M22 m;
    m.m00 = w.x * lo + w.z * hi; m.m01 = w.x * hi + w.z * lo;
    m.m10 = w.y * lo + w.w * hi; m.m11 = w.y * hi + w.w * lo;
    return m;
}

// FAST: the steppers' call shape (coefficient evaluated from (T, S1); A2 given exactly in MASS mode) without the run-time
// selects of the general shape.
// SPLIT = 2: a band is HALF a parent cell row (r_y = 2 RY element rows per cell; MC = 1): the tall ratio (10) otherwise
// needs ~255 registers for its 10 carried rows and has none left for the plane prefetch.  The two halves of a cell add
// their corner sums into a zeroed cellsum with atomicAdd - two terms onto 0, so the result does not depend on their order.
template <int MODE, int RY, int MC, bool FAST, int SPLIT = 1>   // RY element rows per thread = MC parent y-cells of r_y = RY / MC rows each
__global__ void __launch_bounds__(32 * MARCH_WARPS) project_march_kernel(const MarchParams p) {
    constexpr int NR = RY + 1;
    constexpr int ry = RY / MC;
    static_assert(SPLIT == 1 || MC == 1, "half bands hold part of one cell");
    constexpr int NACC = MODE == 1 ? 8 : 12;
    static_assert(RY % MC == 0, "a band holds whole parent cells");
    const int lane = threadIdx.x & 31;
    const int wchunk = blockIdx.x * MARCH_WARPS + (threadIdx.x >> 5);
    const int cpw = p.epw / p.rx;                 // parent x-cells per warp
    if (wchunk * cpw >= p.ncx) return;            // (whole warp)
    const int cj0 = SPLIT > 1 ? (int)blockIdx.y / SPLIT : (int)blockIdx.y * MC;
    const int row0 = (int)blockIdx.y * RY;        // first element row of the band (before the window's offset)
    const int ck0 = blockIdx.z * p.czn, ck1 = min(ck0 + p.czn, p.ncz);
    const int ex = wchunk * p.epw + lane - p.offx;    // element column = node column of this lane
    const bool colv = lane < p.epw && ex >= 0 && ex < p.fnx - 1;
    const int ic = min(max(ex, 0), p.fnx - 1);
    const int seg = lane % p.rx;                  // lane within its parent cell
    const int ci = wchunk * cpw + lane / p.rx;
    const M22 Mx = m_of(__ldg(p.wx + min(max(ex, 0), p.fnx - 2)));
    M22 My[RY];
    float rowv[RY];                               // 1 where the element row exists (first / last parent cell may be partial)
    int idx[NR];                                  // flat node index of this lane's node in each row of the current plane
    const int Pi = p.fnx * p.fny;
    const int kz0 = __ldg(p.fsz + ck0), kz1 = __ldg(p.fsz + ck1);   // node planes kz0 .. kz1
#pragma unroll
    for (int r = 0; r < NR; ++r) idx[r] = kz0 * Pi + min(max(row0 + r - p.offy, 0), p.fny - 1) * p.fnx + ic;
#pragma unroll
    for (int r = 0; r < RY; ++r) {
        const int j = row0 + r - p.offy;
        My[r] = m_of(__ldg(p.wy + min(max(j, 0), p.fny - 2)));
        rowv[r] = (j >= 0 && j < p.fny - 1) ? 0.125f : 0.f;   // (times the 1/8 of the element mean)
    }
    // carried from plane to plane: the node values (GRAD) and the in-plane stage results of every element row
    struct Carry {
        float ap[NR], anp[NR];                                // GRAD: node values (own column, next column)
        float q0[RY], q1[RY], q2[RY], q3[RY], csp[RY];        // GRAD: Px[0], Px[1], Py[0], Py[1]; MASS: P[cx][cy] = q[cx * 2 + cy]
    };
    // raw node values of one plane: all loads of a plane are independent (one memory round trip per plane), and the next
    // plane's are issued before the current plane is worked on
    struct Raw { float a[NR], b[NR], t[NR], s[NR]; };
    constexpr bool PREFETCH = RY <= 6;
    float S[MC][NACC];                // running sums per parent y-cell of the band
#pragma unroll
    for (int t = 0; t < MC; ++t)
#pragma unroll
        for (int c = 0; c < NACC; ++c) S[t][c] = 0.f;
    int ck = ck0;
    int knext = __ldg(p.fsz + ck0 + 1);                              // plane at which cell ck is complete
    const bool hasB = FAST ? MODE == 1 : p.A2 != nullptr, hasC = FAST ? false : p.coef != nullptr;
    auto load_raw = [&](Raw& w) {   // the plane idx[] points at; then idx[] moves on
#pragma unroll
        for (int r = 0; r < NR; ++r) w.a[r] = __ldg(p.A + idx[r]);
        if (hasB) {
#pragma unroll
            for (int r = 0; r < NR; ++r) w.b[r] = __ldg(p.A2 + idx[r]);
        }
        if (hasC) {
#pragma unroll
            for (int r = 0; r < NR; ++r) w.t[r] = __ldg(p.coef + idx[r]);
        } else {
#pragma unroll
            for (int r = 0; r < NR; ++r) {
                w.t[r] = __ldg(p.cT + idx[r]);
                w.s[r] = __ldg(p.cS1 + idx[r]);
            }
        }
#pragma unroll
        for (int r = 0; r < NR; ++r) idx[r] += Pi;
    };
    auto flush = [&]() {
        // parent cell layer ck is complete: corner sums, x reduction over the cell's r_x lanes, store
#pragma unroll
        for (int t = 0; t < MC; ++t) {
            float* s = S[t];
            float o[8];    // hex8 local order: 0 (000) 1 (100) 2 (110) 3 (010) 4 (001) 5 (101) 6 (111) 7 (011), bits = (x, y, z)
            if (MODE == 1) {
                const float f = colv ? -(p.sw * p.inv_cvol) : 0.f;
                // s[(cx * 2 + cy) * 2 + cz]
                o[0] = f * s[0]; o[1] = f * s[4]; o[2] = f * s[6]; o[3] = f * s[2];
                o[4] = f * s[1]; o[5] = f * s[5]; o[6] = f * s[7]; o[7] = f * s[3];
            } else {
                const float f = colv ? -(p.sw * p.inv_cvol) * 2.0f : 0.f;
                const float gx = f * p.inv_hfx, gy = f * p.inv_hfy, gz = f * p.inv_hfz;
                // out[cx][cy][cz] = sgn(cx) gx AX[cy][cz] + sgn(cy) gy AY[cx][cz] + sgn(cz) gz AZ[cx][cy]
                float ax[4], ay[4], az[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) { ax[c] = gx * s[c]; ay[c] = gy * s[4 + c]; az[c] = gz * s[8 + c]; }
#define GM_OUT(cx, cy, cz) ((cx ? ax[cy * 2 + cz] : -ax[cy * 2 + cz]) + (cy ? ay[cx * 2 + cz] : -ay[cx * 2 + cz]) + (cz ? az[cx * 2 + cy] : -az[cx * 2 + cy]))
                o[0] = GM_OUT(0, 0, 0); o[1] = GM_OUT(1, 0, 0); o[2] = GM_OUT(1, 1, 0); o[3] = GM_OUT(0, 1, 0);
                o[4] = GM_OUT(0, 0, 1); o[5] = GM_OUT(1, 0, 1); o[6] = GM_OUT(1, 1, 1); o[7] = GM_OUT(0, 1, 1);
#undef GM_OUT
            }
            for (int d = 1; d < p.rx; d <<= 1) {
                const bool take = seg + d < p.rx;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const float v = __shfl_down_sync(0xffffffffu, o[c], d);
                    o[c] += take ? v : 0.f;
                }
            }
            const int cj = cj0 + t;
            if (seg == 0 && lane < p.epw && ci < p.ncx && cj < p.ncy) {
                float* cs = p.cellsum + (((long long)ck * p.ncy + cj) * p.ncx + ci) * 8;
                if (SPLIT > 1) {
#pragma unroll
                    for (int c = 0; c < 8; ++c) atomicAdd(cs + c, o[c]);
                } else {
                    float4* dst = reinterpret_cast<float4*>(cs);
                    dst[0] = make_float4(o[0], o[1], o[2], o[3]);
                    dst[1] = make_float4(o[4], o[5], o[6], o[7]);
                }
            }
#pragma unroll
            for (int c = 0; c < NACC; ++c) s[c] = 0.f;
        }
    };
    // one node plane: cur = its raw values, cp = what the previous plane left, cn = what this plane leaves (the same
    // object: every entry is read before it is overwritten)
    auto plane = [&](int k, const Raw& cur, const Carry& cp, Carry& cn) {
        // substrate nodes: flat id < cnsub (cF:2589-2590), as a threshold on idx[], which already points `ahead` planes
        // past this one (its own load and, unless this is the last plane, the prefetch)
        const int ahead = (PREFETCH && k < kz1) ? 2 : 1;
        long long left = p.cnsub - (long long)k * Pi;
        left = left < 0 ? 0 : (left > Pi ? Pi : left);
        const int sub_thr = (int)left + (k + ahead) * Pi;
        float a[NR], an[NR], cs2[NR];
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            a[r] = hasB ? cur.a[r] - cur.b[r] : cur.a[r];
            if (hasC) {
                cs2[r] = cur.t[r];
            } else {
                float kk, rr;
                bool b1, b2;
                node_props(p.pk, cur.t[r], cur.s[r], idx[r] < sub_thr, kk, rr, b1, b2);
                cs2[r] = MODE == 1 ? rr : kk;
            }
        }
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            an[r] = __shfl_down_sync(0xffffffffu, a[r], 1);
            cs2[r] += __shfl_down_sync(0xffffffffu, cs2[r], 1);
        }
        const bool have_prev = k > kz0;
        M22 Mz;
        if (have_prev) Mz = m_of(__ldg(p.wz + (k - 1)));
        if (MODE == 1) {
            float tn0[NR], tn1[NR];   // x stage per node row: tn[cx] = Mx[cx][0] a + Mx[cx][1] an
#pragma unroll
            for (int r = 0; r < NR; ++r) {
                tn0[r] = Mx.m00 * a[r] + Mx.m01 * an[r];
                tn1[r] = Mx.m10 * a[r] + Mx.m11 * an[r];
            }
#pragma unroll
            for (int t = 0; t < MC; ++t) {
                float L[4] = {0.f, 0.f, 0.f, 0.f}, U[4] = {0.f, 0.f, 0.f, 0.f};   // sum over the cell's rows of cbar P, lower / upper plane
#pragma unroll
                for (int rr = 0; rr < ry; ++rr) {
                    const int r = t * ry + rr;
                    const float cs = cs2[r] + cs2[r + 1];
                    // P[cx][cy] = My[cy][0] tn[r][cx] + My[cy][1] tn[r + 1][cx]
                    const float p00 = My[r].m00 * tn0[r] + My[r].m01 * tn0[r + 1], p01 = My[r].m10 * tn0[r] + My[r].m11 * tn0[r + 1];
                    const float p10 = My[r].m00 * tn1[r] + My[r].m01 * tn1[r + 1], p11 = My[r].m10 * tn1[r] + My[r].m11 * tn1[r + 1];
                    if (have_prev) {
                        const float cb = (cs + cp.csp[r]) * rowv[r];
                        L[0] += cb * cp.q0[r]; L[1] += cb * cp.q1[r]; L[2] += cb * cp.q2[r]; L[3] += cb * cp.q3[r];
                        U[0] += cb * p00; U[1] += cb * p01; U[2] += cb * p10; U[3] += cb * p11;
                    }
                    cn.q0[r] = p00; cn.q1[r] = p01; cn.q2[r] = p10; cn.q3[r] = p11; cn.csp[r] = cs;
                }
                if (have_prev) {
                    float* s = S[t];   // s[(cx * 2 + cy) * 2 + cz] += Mz[cz][0] L + Mz[cz][1] U
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        s[2 * c] += Mz.m00 * L[c] + Mz.m01 * U[c];
                        s[2 * c + 1] += Mz.m10 * L[c] + Mz.m11 * U[c];
                    }
                }
            }
        } else {
            float tz0[NR], tz1[NR];   // z-difference, x stage per node row
            if (have_prev) {
#pragma unroll
                for (int r = 0; r < NR; ++r) {
                    const float d = a[r] - cp.ap[r], dn = an[r] - cp.anp[r];
                    tz0[r] = Mx.m00 * d + Mx.m01 * dn;
                    tz1[r] = Mx.m10 * d + Mx.m11 * dn;
                }
            }
#pragma unroll
            for (int t = 0; t < MC; ++t) {
                // sums over the cell's rows of kbar times the in-plane stage results: x term [cy], y term [cx], lower / upper plane
                float LX[2] = {0.f, 0.f}, UX[2] = {0.f, 0.f}, LY[2] = {0.f, 0.f}, UY[2] = {0.f, 0.f};
                float* s = S[t];
#pragma unroll
                for (int rr = 0; rr < ry; ++rr) {
                    const int r = t * ry + rr;
                    const float cs = cs2[r] + cs2[r + 1];
                    const float dx0 = an[r] - a[r], dx1 = an[r + 1] - a[r + 1];
                    const float dy0 = a[r + 1] - a[r], dy1 = an[r + 1] - an[r];
                    const float px0 = My[r].m00 * dx0 + My[r].m01 * dx1, px1 = My[r].m10 * dx0 + My[r].m11 * dx1;   // Px[cy]
                    const float py0 = Mx.m00 * dy0 + Mx.m01 * dy1, py1 = Mx.m10 * dy0 + Mx.m11 * dy1;               // Py[cx]
                    if (have_prev) {
                        const float kb = (cs + cp.csp[r]) * rowv[r];
                        LX[0] += kb * cp.q0[r]; LX[1] += kb * cp.q1[r]; UX[0] += kb * px0; UX[1] += kb * px1;
                        LY[0] += kb * cp.q2[r]; LY[1] += kb * cp.q3[r]; UY[0] += kb * py0; UY[1] += kb * py1;
                        // AZ[cx][cy] = s[8 + cx * 2 + cy] += kbar (My[cy][0] tz[r][cx] + My[cy][1] tz[r + 1][cx])
                        const float k00 = kb * My[r].m00, k01 = kb * My[r].m01, k10 = kb * My[r].m10, k11 = kb * My[r].m11;
                        s[8] += k00 * tz0[r] + k01 * tz0[r + 1]; s[9] += k10 * tz0[r] + k11 * tz0[r + 1];
                        s[10] += k00 * tz1[r] + k01 * tz1[r + 1]; s[11] += k10 * tz1[r] + k11 * tz1[r + 1];
                    }
                    cn.q0[r] = px0; cn.q1[r] = px1; cn.q2[r] = py0; cn.q3[r] = py1; cn.csp[r] = cs;
                }
                if (have_prev) {
                    // AX[cy][cz] = s[cy * 2 + cz], AY[cx][cz] = s[4 + cx * 2 + cz]: z stage once per cell and plane
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        s[2 * c] += Mz.m00 * LX[c] + Mz.m01 * UX[c];
                        s[2 * c + 1] += Mz.m10 * LX[c] + Mz.m11 * UX[c];
                        s[4 + 2 * c] += Mz.m00 * LY[c] + Mz.m01 * UY[c];
                        s[4 + 2 * c + 1] += Mz.m10 * LY[c] + Mz.m11 * UY[c];
                    }
                }
            }
#pragma unroll
            for (int r = 0; r < NR; ++r) { cn.ap[r] = a[r]; cn.anp[r] = an[r]; }
        }
        while (ck < ck1 && k == knext) {
            flush();
            ++ck;
            if (ck < ck1) knext = __ldg(p.fsz + ck + 1);
        }
    };
    Raw cur, nxt;
    Carry cy;
    load_raw(cur);
    for (int k = kz0; k <= kz1; ++k) {
        if (PREFETCH) {
            if (k < kz1) load_raw(nxt);
        } else if (k > kz0) {   // (the tall bands have no registers to spare for a second raw plane)
            load_raw(cur);
        }
        plane(k, cur, cy, cy);
        if (PREFETCH) cur = nxt;
    }
}

// z-chunks of whole parent cells: enough warps for ~2.5 resident sets of 8 per SM, no more - a finer cut (chosen from the
// kernel's occupancy by a rounds x planes model) measured 0-20 % slower: every chunk re-reads a plane and re-forms its
// tables, and short marches lose the plane prefetch's head start.
static int march_pick_chunk(MarchParams& mp, int nwx, int nby) {
    const long long target = 20LL * sm_count();
    long long nzc = (target + (long long)nwx * nby - 1) / ((long long)nwx * nby);
    if (nzc > mp.ncz) nzc = mp.ncz;
    if (nzc < 1) nzc = 1;
    return (int)((mp.ncz + nzc - 1) / nzc);
}

template <int MODE, int RY, int MC, bool FAST, int SPLIT = 1>
static void launch_march_i(MarchParams& mp, int nwx, int nby, int fine_layers, cudaStream_t st) {
    (void)fine_layers;
    nby *= SPLIT;
    if (SPLIT > 1) cudaMemsetAsync(mp.cellsum, 0, (size_t)mp.ncx * mp.ncy * mp.ncz * 8 * sizeof(float), st);   // the halves add
    mp.czn = march_pick_chunk(mp, nwx, nby);
    const dim3 grid((nwx + MARCH_WARPS - 1) / MARCH_WARPS, nby, (mp.ncz + mp.czn - 1) / mp.czn);
    project_march_kernel<MODE, RY, MC, FAST, SPLIT><<<grid, 32 * MARCH_WARPS, 0, st>>>(mp);
}

template <int MODE, bool FAST>
static bool launch_march_f(MarchParams& mp, int nwx, int nby, int fine_layers, cudaStream_t st) {
    switch (mp.ry) {
        case 1: launch_march_i<MODE, 4, 4, FAST>(mp, nwx, nby, fine_layers, st); break;
        case 2: launch_march_i<MODE, 4, 2, FAST>(mp, nwx, nby, fine_layers, st); break;
        case 3: launch_march_i<MODE, 3, 1, FAST>(mp, nwx, nby, fine_layers, st); break;
        case 4: launch_march_i<MODE, 4, 1, FAST>(mp, nwx, nby, fine_layers, st); break;
        case 5: launch_march_i<MODE, 5, 1, FAST>(mp, nwx, nby, fine_layers, st); break;
        case 6: launch_march_i<MODE, 6, 1, FAST>(mp, nwx, nby, fine_layers, st); break;
        case 8: launch_march_i<MODE, 8, 1, FAST>(mp, nwx, nby, fine_layers, st); break;
        case 10: launch_march_i<MODE, 5, 1, FAST, 2>(mp, nwx, nby, fine_layers, st); break;   // two half-cell bands
        default: return false;
    }
    count_launch();
    return true;
}
template <int MODE>
static bool launch_march(MarchParams& mp, int nwx, int nby, int fine_layers, cudaStream_t st) {
    const bool fast = !mp.coef && ((MODE == 1) == (mp.A2 != nullptr));
    return fast ? launch_march_f<MODE, true>(mp, nwx, nby, fine_layers, st) : launch_march_f<MODE, false>(mp, nwx, nby, fine_layers, st);
}
static inline int march_cells_per_band(int ry) { return ry == 1 ? 4 : (ry == 2 ? 2 : 1); }
static inline bool march_ry_ok(int ry) { return ry == 1 || ry == 2 || ry == 3 || ry == 4 || ry == 5 || ry == 6 || ry == 8 || ry == 10; }

// parent node (gi,gj,gk) of the box [c0, c0 + nc] (nodes) <- its <= 8 adjacent cells, increasing cell id
// (a 3-D launch without the index divisions was measured: 263 us against 248 us per C2 block for the 19 launches - the
// kernel waits for its 32-byte-strided gathers, not for the divisions)
__global__ void project_nodes_kernel(const float* __restrict__ cellsum, int c0x, int c0y, int c0z, int ncx, int ncy,
                                     int ncz, int pnx, int pny, float* __restrict__ V, int accumulate) {
    const long long total = (long long)(ncx + 1) * (ncy + 1) * (ncz + 1);
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(t % (ncx + 1)), j = (int)((t / (ncx + 1)) % (ncy + 1)),
                  k = (int)(t / ((long long)(ncx + 1) * (ncy + 1)));
        float s = 0.f;
        for (int dk = 1; dk >= 0; --dk)
            for (int dj = 1; dj >= 0; --dj)
                for (int di = 1; di >= 0; --di) {  // cell (i-di, j-dj, k-dk): this node is its corner (di,dj,dk)
                    const int ci = i - di, cj = j - dj, ck = k - dk;
                    if (ci < 0 || cj < 0 || ck < 0 || ci >= ncx || cj >= ncy || ck >= ncz) continue;
                    const int corner = dk * 4 + (dj == 0 ? (di == 0 ? 0 : 1) : (di == 0 ? 3 : 2));
                    s += cellsum[(((long long)ck * ncy + cj) * ncx + ci) * 8 + corner];
                }
        const long long n = (c0x + i) + (long long)(c0y + j) * pnx + (long long)(c0z + k) * pnx * pny;
        V[n] = accumulate ? V[n] + s : s;
    }
}


// ---- batched projected source: tables of all rows in one launch, then one pass over the parent ------------------
struct SrcBatch {
    float v[GOMELT_MAX_SUBSTEPS][3];
    float c[GOMELT_MAX_SUBSTEPS];
    int n;
};
__global__ void coarse_source_table_batch_kernel(const float* __restrict__ fx, const float* __restrict__ fy,
                                                 const float* __restrict__ fz, int nfx, int nfy, int nfz,
                                                 const float* __restrict__ cx, const float* __restrict__ cy,
                                                 const float* __restrict__ cz, int ncx, int ncy, int ncz,
                                                 const __grid_constant__ SrcBatch sb, float inv_r2, float inv_d2, float rc, float dc,
                                                 float* __restrict__ tables) {
    const int axis = blockIdx.y, r = blockIdx.z;
    const float* xf = axis == 0 ? fx : (axis == 1 ? fy : fz);
    const float* xc = axis == 0 ? cx : (axis == 1 ? cy : cz);
    const int nf = axis == 0 ? nfx : (axis == 1 ? nfy : nfz), nc = axis == 0 ? ncx : (axis == 1 ? ncy : ncz);
    const int ic = blockIdx.x * blockDim.x + threadIdx.x;
    if (ic >= nc) return;
    const float v = sb.v[r][axis], inv_s2 = axis == 2 ? inv_d2 : inv_r2, c = axis == 2 ? dc : rc;
    const float g = 0.57735026918962576f;
    const float Nlo = 0.5f * (1.f + g), Nhi = 0.5f * (1.f - g);
    const float hc = xc[1] - xc[0];
    const float inv_hc = 1.0f / hc;
    float acc = 0.f;
    // only fine elements whose first Gauss point lies in parent cell ic - 1 or ic contribute: bracket them (two elements
    // of slack on either side; the exact cell test below decides)
    const float hf = xf[1] - xf[0];
    int elo = (int)floorf((xc[max(ic - 1, 0)] - xf[0]) / hf) - 2, ehi = (int)floorf((xc[min(ic + 1, nc - 1)] - xf[0]) / hf) + 2;
    if (ic <= 1) elo = 0;            // (clipped cells: everything below / above the parent grid belongs to the end cells)
    if (ic >= nc - 2) ehi = nf - 2;
    elo = max(elo, 0);
    ehi = min(ehi, nf - 2);
    for (int e = elo; e <= ehi; ++e) {  // same body as coarse_source_table_kernel
        const float x0 = xf[e], x1 = xf[e + 1];
        const float xq0 = Nlo * x0 + Nhi * x1, xq1 = Nhi * x0 + Nlo * x1;
        const int ec = cell_of(xq0, xc[0], hc, nc - 1);
        if (ec != ic && ec + 1 != ic) continue;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const float xq = q == 0 ? xq0 : xq1;
            const int eq = q == 0 ? ec : cell_of(xq1, xc[0], hc, nc - 1);
            const float xc0 = xc[eq], xc1 = xc[eq + 1];
            const float w = (ec == ic) ? (xc1 - xq) * inv_hc : (xq - xc0) * inv_hc;
            const float d = xq - v;
            acc += w * (c * expf(-3.f * d * d * inv_s2));
        }
    }
    tables[(size_t)r * (ncx + ncy + ncz) + (axis == 0 ? 0 : (axis == 1 ? ncx : ncx + ncy)) + ic] = acc;
}
__global__ void rank_n_kernel(float* __restrict__ F, const float* __restrict__ tables, int nx, int ny, int nz,
                              const __grid_constant__ SrcBatch sb, int accumulate) {
    // blockIdx.y = group of ROWS3 rows, blockIdx.z = plane k; per laser row the x factor is read once per thread and
    // the z factor once per block
    const int stride = nx + ny + nz;
    const int k = blockIdx.z;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nx) return;
    const int j0 = blockIdx.y * ROWS3, j1 = min(ny, j0 + ROWS3);
    float f[ROWS3];
#pragma unroll
    for (int q = 0; q < ROWS3; ++q) f[q] = (accumulate && j0 + q < j1) ? F[((size_t)k * ny + j0 + q) * nx + i] : 0.f;
    for (int r = 0; r < sb.n; ++r) {  // laser rows in order, like the row-by-row accumulation of gomelt_rank1_f32
        const float* tb = tables + (size_t)r * stride;
        const float txv = tb[i], tzv = tb[nx + ny + k], c = sb.c[r];
        // the tables are Gaussians that underflow to exactly 0 some 6 laser radii away (all but ~8 % of the columns and most
        // planes of a C2-size parent): there the row adds c * ((0 * ty) * tz) = 0, i.e. nothing
        if (txv == 0.f || tzv == 0.f) continue;
#pragma unroll
        for (int q = 0; q < ROWS3; ++q)
            if (j0 + q < j1) f[q] = f[q] + c * ((txv * tb[nx + j0 + q]) * tzv);
    }
#pragma unroll
    for (int q = 0; q < ROWS3; ++q)
        if (j0 + q < j1) F[((size_t)k * ny + j0 + q) * nx + i] = f[q];
}

// ---- K5: fused window shift -----------------------------------------------------------------------------------------
struct ShiftParams {
    AxisView ax, ay, az;  const float* T1;      // Level 1
    AxisView mx, my, mz;  const float* Tpm;     // Level-2 window at its old position (or nullptr)
    AxisView ox, oy, oz;  const float* Tpo;     // this window at its old position
    const float *tx, *ty, *tz;
    int ntx, nty, ntz;
    float *Tp_new, *T_new;
};
// The y part of the block's rows and the z part of its plane do not depend on the column: one warp per parent forms them
// (lane q < rows: row q, lane ROWS3: the plane) and the block reads them after one barrier, instead of every thread
// repeating the cell searches (IEEE divisions) of its rows: 144 -> 130 us (Level 3), 104 -> 97 us (Level 2) at C2 size;
// with tri_eval requesting its 8 parent values before anything else and deciding validity by a select, the loads of
// all three interpolations of a target are in flight together: 119 / 87 us.
// Tried on top of it and dropped (bench_tools/quick_interp.py, one box): the marching form of interp3_march_kernel with
// a register cache per parent - the window's own old field has the SAME spacing, every row enters a new cell and its
// loads come from DRAM: 190 / 143 us even with the next parent row requested a row ahead; the short form of the weights
// (tri_fast) with per-target loads: 138 / 106 us at 64 registers, 137 / 104 with a 48-register cap.
__device__ __forceinline__ float4 ax1_pack(const Ax1& a) { return make_float4(__int_as_float(a.e), a.a0, a.a1, 0.f); }
__device__ __forceinline__ Ax1 ax1_unpack(const float4& r) {
    Ax1 a;
    a.e = __float_as_int(r.x); a.a0 = r.y; a.a1 = r.z;
    return a;
}
__global__ void __launch_bounds__(256) shift_window_kernel(const ShiftParams p, const int rpb) {
    __shared__ float4 s_ax[3][ROWS3 + 1];
    const int i = min((int)(blockIdx.x * blockDim.x + threadIdx.x), p.ntx - 1);  // blockIdx.y = group of rpb target rows, blockIdx.z = target plane
    const bool owner = (int)(blockIdx.x * blockDim.x + threadIdx.x) < p.ntx;
    const int k = blockIdx.z;
    const int j0 = blockIdx.y * rpb, j1 = min(p.nty, j0 + rpb);
    const float x = p.tx[i];
    const SrcGeom g1 = geom_of(p.ax, p.ay, p.az), go = geom_of(p.ox, p.oy, p.oz);
    const Ax1 X1 = ax1_of(p.ax, g1.hx, x), Xo = ax1_of(p.ox, go.hx, x);
    SrcGeom gm = go;
    Ax1 Xm = Xo;
    if (p.Tpm) {
        gm = geom_of(p.mx, p.my, p.mz);
        Xm = ax1_of(p.mx, gm.hx, x);
    }
    {
        const int which = threadIdx.x >> 5, t = threadIdx.x & 31;
        if (which < (p.Tpm ? 3 : 2) && (t < j1 - j0 || t == ROWS3)) {
            const AxisView& ay = which == 0 ? p.oy : which == 1 ? p.ay : p.my;
            const AxisView& az = which == 0 ? p.oz : which == 1 ? p.az : p.mz;
            const SrcGeom& g = which == 0 ? go : which == 1 ? g1 : gm;
            s_ax[which][t] = t == ROWS3 ? ax1_pack(ax1_of(az, g.hz, p.tz[k])) : ax1_pack(ax1_of(ay, g.hy, p.ty[j0 + t]));
        }
    }
    __syncthreads();
    if (!owner) return;
    const Ax1 Zo = ax1_unpack(s_ax[0][ROWS3]), Z1 = ax1_unpack(s_ax[1][ROWS3]);
    const Ax1 Zm = p.Tpm ? ax1_unpack(s_ax[2][ROWS3]) : Zo;
    for (int j = j0; j < j1; ++j) {
        const size_t w = ((size_t)k * p.nty + j) * p.ntx + i;
        const float tp = tri_eval<false>(p.Tpo, nullptr, 1.f, 0.f, go, Xo, ax1_unpack(s_ax[0][j - j0]), Zo);
        const float t1 = tri_eval<false>(p.T1, nullptr, 1.f, 0.f, g1, X1, ax1_unpack(s_ax[1][j - j0]), Z1);
        float rest = tp;
        if (p.Tpm) rest = __fadd_rn(tri_eval<false>(p.Tpm, nullptr, 1.f, 0.f, gm, Xm, ax1_unpack(s_ax[2][j - j0]), Zm), tp);  // T1on3 + (Tp2on3 + Tp3)
        p.Tp_new[w] = tp;
        p.T_new[w] = __fadd_rn(t1, rest);
    }
}

// Row groups of the 3-D launch forms (blockIdx.y): ROWS3 rows per block on large grids; one row per block where that
// would leave the GPU with fewer than ~4 blocks per SM (the example's 101 x 101 x 11-node windows: the rows of a block run
// one after the other, each with its chain of dependent loads).
static inline int row_groups(int nx_blocks, int ny, int nz) {
    const long long blocks = (long long)nx_blocks * ((ny + ROWS3 - 1) / ROWS3) * nz;
    return blocks >= 4LL * sm_count() ? (ny + ROWS3 - 1) / ROWS3 : ny;
}

static inline int row_groups_march(int nx_blocks, int ny, int nz) {
    const long long blocks = (long long)nx_blocks * ((ny + ROWSM - 1) / ROWSM) * nz;
    return blocks >= 4LL * sm_count() ? (ny + ROWSM - 1) / ROWSM : ny;
}
static inline int grid_for(long long n, int threads) {
    long long b = (n + threads - 1) / threads;
    const long long cap = (long long)sm_count() * 16;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace gomelt

using namespace gomelt;

static bool axis_ok(const gomelt_axis_t& a) { return a.coords && a.n >= 2; }

extern "C" int gomelt_interp_f32(const gomelt_interp_args_t* a, void* stream) {
    if (!a || !a->u || !a->out || !a->tx || !a->ty || !a->tz || !axis_ok(a->src[0]) || !axis_ok(a->src[1]) ||
        !axis_ok(a->src[2])) {
        set_error("gomelt_interp_f32: NULL argument / source axis with < 2 nodes");
        return GOMELT_E_NULL;
    }
    if (a->ntx < 1 || a->nty < 1 || a->ntz < 1) {
        set_error("gomelt_interp_f32: empty target grid");
        return GOMELT_E_SIZE;
    }
    if ((a->mode == GOMELT_INTERP_RSUB && !a->base) || a->mode < 0 || a->mode > 2 ||
        ((a->map_x || a->map_y || a->map_z) && !(a->map_x && a->map_y && a->map_z))) {
        set_error("gomelt_interp_f32: bad mode / base / map");
        return GOMELT_E_FLAGS;
    }
    InterpParams p;
    p.sx = {a->src[0].coords, a->src[0].n};
    p.sy = {a->src[1].coords, a->src[1].n};
    p.sz = {a->src[2].coords, a->src[2].n};
    p.u = a->u; p.u2 = a->u2; p.alpha = a->alpha; p.beta = a->beta;
    p.tx = a->tx; p.ty = a->ty; p.tz = a->tz; p.ntx = a->ntx; p.nty = a->nty; p.ntz = a->ntz;
    p.mode = a->mode; p.faces_only = a->faces_only; p.clamp_min = a->clamp_min; p.has_clamp = a->has_clamp;
    p.mx = a->map_x; p.my = a->map_y; p.mz = a->map_z; p.map_nx = a->map_nx; p.map_ny = a->map_ny;
    p.base = a->base; p.out = a->out;
    long long total = (long long)a->ntx * a->nty * a->ntz;
    if (a->faces_only)
        total = (long long)a->ntx * a->nty + 2LL * a->ntx * (a->ntz - 1) + 2LL * (a->nty - 2) * (a->ntz - 1);
    const long long src_nn = (long long)a->src[0].n * a->src[1].n * a->src[2].n;
    if (!a->faces_only && !a->map_x && a->ntz <= 65535 && a->nty <= 65535 && src_nn < 2000000000LL) {
        const int threads = a->ntx >= 256 ? 256 : (a->ntx >= 128 ? 128 : 64);
        const int nxb = (a->ntx + threads - 1) / threads;
        const int groups = row_groups_march(nxb, a->nty, a->ntz);
        const int rpb = (a->nty + groups - 1) / groups;
        const dim3 grid(nxb, groups, a->ntz);
        cudaStream_t st = (cudaStream_t)stream;
#define GOMELT_I3M(B, M) interp3_march_kernel<B, M><<<grid, threads, 0, st>>>(p, rpb)
        if (a->u2) {
            if (a->mode == GOMELT_INTERP_SET) GOMELT_I3M(true, GOMELT_INTERP_SET);
            else if (a->mode == GOMELT_INTERP_ADD) GOMELT_I3M(true, GOMELT_INTERP_ADD);
            else GOMELT_I3M(true, GOMELT_INTERP_RSUB);
        } else {
            if (a->mode == GOMELT_INTERP_SET) GOMELT_I3M(false, GOMELT_INTERP_SET);
            else if (a->mode == GOMELT_INTERP_ADD) GOMELT_I3M(false, GOMELT_INTERP_ADD);
            else GOMELT_I3M(false, GOMELT_INTERP_RSUB);
        }
#undef GOMELT_I3M
        count_launch();
    } else if (!a->faces_only && a->nty <= 65535 && a->ntz <= 65535) {
        // index-mapped target sets (the injection of getNewTprime): blockIdx.y / .z = target row / plane, no index divisions
        const int threads = a->ntx >= 256 ? 256 : (a->ntx >= 128 ? 128 : 64);
        interp_kernel<<<dim3((a->ntx + threads - 1) / threads, a->nty, a->ntz), threads, 0, (cudaStream_t)stream>>>(p), count_launch();
    } else {
        interp_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(p), count_launch();
    }
    return check_launch("gomelt_interp_f32");
}

extern "C" long long gomelt_faces_count(int32_t ntx, int32_t nty, int32_t ntz) {
    return (long long)ntx * nty + 2LL * ntx * (ntz - 1) + 2LL * (nty - 2) * (ntz - 1);
}

extern "C" int gomelt_faces_gather_f32(const gomelt_interp_args_t* a, float* face_a, float* face_b, void* stream) {
    if (!a || !a->u || !a->u2 || !face_a || !face_b || !a->tx || !a->ty || !a->tz || !axis_ok(a->src[0]) ||
        !axis_ok(a->src[1]) || !axis_ok(a->src[2])) {
        set_error("gomelt_faces_gather_f32: NULL argument / source axis with < 2 nodes (u and u2 are both required)");
        return GOMELT_E_NULL;
    }
    const long long nface = gomelt_faces_count(a->ntx, a->nty, a->ntz);
    if (a->ntx < 2 || a->nty < 2 || a->ntz < 2 || nface > 2000000000LL) {
        set_error("gomelt_faces_gather_f32: bad target grid");
        return GOMELT_E_SIZE;
    }
    InterpParams p = {};
    p.sx = {a->src[0].coords, a->src[0].n};
    p.sy = {a->src[1].coords, a->src[1].n};
    p.sz = {a->src[2].coords, a->src[2].n};
    p.u = a->u; p.u2 = a->u2;
    p.tx = a->tx; p.ty = a->ty; p.tz = a->tz; p.ntx = a->ntx; p.nty = a->nty; p.ntz = a->ntz;
    faces_gather_kernel<<<grid_for(nface, 128), 128, 0, (cudaStream_t)stream>>>(p, face_a, face_b, (int)nface), count_launch();
    return check_launch("gomelt_faces_gather_f32");
}

extern "C" int gomelt_faces_blend_f32(const float* face_a, const float* face_b, int32_t ntx, int32_t nty, int32_t ntz,
                                      float alpha, float beta, int32_t has_clamp, float clamp_min, float* out,
                                      void* stream) {
    if (!face_a || !face_b || !out) {
        set_error("gomelt_faces_blend_f32: NULL argument");
        return GOMELT_E_NULL;
    }
    const long long nface = gomelt_faces_count(ntx, nty, ntz);
    if (ntx < 2 || nty < 2 || ntz < 2 || nface > 2000000000LL) {
        set_error("gomelt_faces_blend_f32: bad target grid");
        return GOMELT_E_SIZE;
    }
    faces_blend_kernel<<<grid_for(nface, 256), 256, 0, (cudaStream_t)stream>>>(face_a, face_b, (int)nface, ntx, nty, ntz,
                                                                              alpha, beta, has_clamp, clamp_min, out), count_launch();
    return check_launch("gomelt_faces_blend_f32");
}

extern "C" int gomelt_box_copy(const void* src, void* dst, int32_t elem_size, const int32_t* ix, const int32_t* iy,
                               const int32_t* iz, int32_t nx, int32_t ny, int32_t nz, int32_t big_nx, int32_t big_ny,
                               int32_t scatter, void* stream) {
    if (!src || !dst || !ix || !iy || !iz) {
        set_error("gomelt_box_copy: NULL argument");
        return GOMELT_E_NULL;
    }
    if (nx < 1 || ny < 1 || nz < 1 || (elem_size != 1 && elem_size != 4)) {
        set_error("gomelt_box_copy: bad size (elem_size must be 1 or 4)");
        return GOMELT_E_SIZE;
    }
    const long long total = (long long)nx * ny * nz;
    cudaStream_t st = (cudaStream_t)stream;
    const bool g3 = nz <= 65535 && (ny > ROWS3 || nz > 1);
    const dim3 grid = g3 ? dim3((nx + 255) / 256, row_groups((nx + 255) / 256, ny, nz), nz) : dim3(grid_for(total, 256));
    if (elem_size == 4)
        box_copy_kernel<float><<<grid, 256, 0, st>>>((const float*)src, (float*)dst, ix, iy, iz, nx, ny,
                                                     nz, big_nx, big_ny, scatter), count_launch();
    else
        box_copy_kernel<uint8_t><<<grid, 256, 0, st>>>((const uint8_t*)src, (uint8_t*)dst, ix, iy, iz,
                                                       nx, ny, nz, big_nx, big_ny, scatter), count_launch();
    return check_launch("gomelt_box_copy");
}

extern "C" int gomelt_rank1_f32(float* F, const float* tx, const float* ty, const float* tz, int32_t nx, int32_t ny,
                                int32_t nz, float coef, int32_t accumulate, void* stream) {
    if (!F || !tx || !ty || !tz) {
        set_error("gomelt_rank1_f32: NULL argument");
        return GOMELT_E_NULL;
    }
    const long long total = (long long)nx * ny * nz;
    rank1_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(F, tx, ty, tz, nx, ny, nz, coef, accumulate), count_launch();
    return check_launch("gomelt_rank1_f32");
}

extern "C" int gomelt_coarse_source_tables_f32(const gomelt_props_t* p, const gomelt_axis_t fine[3],
                                               const gomelt_axis_t parent[3], const float laser_xyz[3], float laserP,
                                               float* tx, float* ty, float* tz, float* coef, void* stream) {
    if (!p || !fine || !parent || !laser_xyz || !tx || !ty || !tz || !coef) {
        set_error("gomelt_coarse_source_tables_f32: NULL argument");
        return GOMELT_E_NULL;
    }
    for (int d = 0; d < 3; ++d)
        if (!axis_ok(fine[d]) || !axis_ok(parent[d])) {
            set_error("gomelt_coarse_source_tables_f32: axis %d needs >= 2 nodes", d);
            return GOMELT_E_SIZE;
        }
    const float pcoeff = 6.f * sqrtf(3.f) * laserP * p->laser_eta;
    const float rcoeff = 1.f / (p->laser_radius * sqrtf((float)M_PI));
    const float dcoeff = 1.f / (p->laser_depth * sqrtf((float)M_PI));
    const float rsq = p->laser_radius * p->laser_radius, dsq = p->laser_depth * p->laser_depth;
    // wq of the FINE element: h from the device coordinates is not readable here; the caller passes it
    // folded into laserP?  No: coef = pcoeff only; the fine wq is applied by the caller (it knows h).
    *coef = pcoeff;
    cudaStream_t st = (cudaStream_t)stream;
    float* out[3] = {tx, ty, tz};
    const float cc[3] = {rcoeff, rcoeff, dcoeff};
    const float is2[3] = {1.f / rsq, 1.f / rsq, 1.f / dsq};
    for (int d = 0; d < 3; ++d)
        coarse_source_table_kernel<<<(parent[d].n + 63) / 64, 64, 0, st>>>(fine[d].coords, fine[d].n, parent[d].coords,
                                                                           parent[d].n, laser_xyz[d], is2[d], cc[d],
                                                                           out[d]), count_launch();
    return check_launch("gomelt_coarse_source_tables_f32");
}

extern "C" int gomelt_project_f32(const gomelt_project_args_t* a, void* stream) {
    if (!a || !a->A || !a->V || !a->cellsum || !a->first_x || !a->first_y || !a->first_z ||
        (!a->coef && !(a->coef_T && a->coef_S1 && a->coef_props))) {
        set_error("gomelt_project_f32: NULL argument (coef, or coef_T + coef_S1 + coef_props)");
        return GOMELT_E_NULL;
    }
    for (int d = 0; d < 3; ++d)
        if (!axis_ok(a->fine[d]) || !axis_ok(a->parent[d])) {
            set_error("gomelt_project_f32: axis %d needs >= 2 nodes", d);
            return GOMELT_E_SIZE;
        }
    if (a->ncell[0] < 1 || a->ncell[1] < 1 || a->ncell[2] < 1 || a->mode < 0 || a->mode > 1) {
        set_error("gomelt_project_f32: bad cell box / mode");
        return GOMELT_E_SIZE;
    }
    ProjParams p;
    p.fx = {a->fine[0].coords, a->fine[0].n}; p.fy = {a->fine[1].coords, a->fine[1].n};
    p.fz = {a->fine[2].coords, a->fine[2].n};
    p.cx = {a->parent[0].coords, a->parent[0].n}; p.cy = {a->parent[1].coords, a->parent[1].n};
    p.cz = {a->parent[2].coords, a->parent[2].n};
    p.A = a->A; p.A2 = a->A2; p.coef = a->coef; p.mode = a->mode; p.scale = a->scale;
    p.cT = a->coef_T; p.cS1 = a->coef_S1; p.cnsub = a->coef_n_substrate;
    if (!a->coef) p.pk = fold_props(*a->coef_props);
    p.c0x = a->cell0[0]; p.c0y = a->cell0[1]; p.c0z = a->cell0[2];
    p.ncx = a->ncell[0]; p.ncy = a->ncell[1]; p.ncz = a->ncell[2];
    p.fsx = a->first_x; p.fsy = a->first_y; p.fsz = a->first_z;
    p.cellsum = a->cellsum;
    const long long ncell = (long long)p.ncx * p.ncy * p.ncz;
    cudaStream_t st = (cudaStream_t)stream;
    bool tiled = a->wtab_x && a->wtab_y && a->wtab_z && a->rmax[0] >= 1 && a->rmax[1] >= 1 && a->rmax[2] >= 1;
    // marching form: uniform integer ratio in x and y (every cell of the box holds exactly rmax fine elements: the counts
    // sum to the number of fine elements), any grouping in z; elems_per_cell_hint < 0 asks for the tile kernel (A/B)
    bool marched = false;
    const int offx = a->uniform_off[0], offy = a->uniform_off[1];
    const bool nest_x = offx > 0 ? offx < a->rmax[0] : (offx == 0 && (long long)a->rmax[0] * p.ncx == a->fine[0].n - 1);
    const bool nest_y = offy > 0 ? offy < a->rmax[1] : (offy == 0 && (long long)a->rmax[1] * p.ncy == a->fine[1].n - 1);
    const long long fnn = (long long)a->fine[0].n * a->fine[1].n * a->fine[2].n;
    if (tiled && a->elems_per_cell_hint >= 0 && nest_x && nest_y && a->rmax[0] <= 31 && march_ry_ok(a->rmax[1]) &&
        fnn + 2LL * a->fine[0].n * a->fine[1].n < 2147483647LL) {
        MarchParams mp;
        mp.fnx = a->fine[0].n; mp.fny = a->fine[1].n; mp.fnz = a->fine[2].n;
        mp.A = a->A; mp.A2 = a->A2; mp.coef = a->coef; mp.cT = a->coef_T; mp.cS1 = a->coef_S1; mp.cnsub = a->coef_n_substrate;
        if (!a->coef) mp.pk = fold_props(*a->coef_props);
        mp.wx = reinterpret_cast<const float4*>(a->wtab_x); mp.wy = reinterpret_cast<const float4*>(a->wtab_y);
        mp.wz = reinterpret_cast<const float4*>(a->wtab_z);
        mp.fsz = a->first_z;
        mp.ncx = p.ncx; mp.ncy = p.ncy; mp.ncz = p.ncz;
        mp.rx = a->rmax[0]; mp.ry = a->rmax[1];
        mp.offx = offx; mp.offy = offy;
        mp.epw = (31 / mp.rx) * mp.rx;
        const float wq = (a->hf[0] * a->hf[1] * a->hf[2]) * 0.125f;
        mp.sw = a->mode == 1 ? a->scale * wq : wq;
        mp.inv_hfx = 1.0f / a->hf[0]; mp.inv_hfy = 1.0f / a->hf[1]; mp.inv_hfz = 1.0f / a->hf[2];
        mp.inv_cvol = 1.0f / ((a->hc[0] * a->hc[1]) * a->hc[2]);
        mp.cellsum = a->cellsum;
        const int cpw = mp.epw / mp.rx;
        const int nwx = (p.ncx + cpw - 1) / cpw;
        const int mc = march_cells_per_band(mp.ry);
        const int nby = (p.ncy + mc - 1) / mc;
        if (nby <= 65535 && p.ncz <= 65535)
            marched = a->mode == 1 ? launch_march<1>(mp, nwx, nby, a->fine[2].n - 1, st) : launch_march<0>(mp, nwx, nby, a->fine[2].n - 1, st);
    }
    TileParams tp;
    if (marched) {
        tiled = true;
    } else if (tiled) {
        // parent cells per CTA: about 32 x 8 x 4 fine elements, the node box within the shared-memory capacity
        int pc[3] = {32 / a->rmax[0], 8 / a->rmax[1], 4 / a->rmax[2]};
        for (int d = 0; d < 3; ++d) pc[d] = pc[d] < 1 ? 1 : pc[d];
        auto nodes = [&] { return (long long)(pc[0] * a->rmax[0] + 1) * (pc[1] * a->rmax[1] + 1) * (pc[2] * a->rmax[2] + 1); };
        while ((nodes() > TILE_NODECAP || pc[0] * pc[1] * pc[2] > TILE_THREADS) && (pc[0] > 1 || pc[1] > 1 || pc[2] > 1)) {
            int d = pc[0] >= pc[1] && pc[0] >= pc[2] ? 0 : (pc[1] >= pc[2] ? 1 : 2);
            if (pc[d] == 1) d = pc[0] > 1 ? 0 : (pc[1] > 1 ? 1 : 2);
            pc[d] = (pc[d] + 1) / 2;
        }
        tiled = nodes() <= TILE_NODECAP && pc[0] * pc[1] * pc[2] <= TILE_THREADS;
        if (tiled) {
            tp.fnx = a->fine[0].n; tp.fny = a->fine[1].n; tp.fnz = a->fine[2].n;
            tp.A = a->A; tp.A2 = a->A2; tp.coef = a->coef; tp.cT = a->coef_T; tp.cS1 = a->coef_S1; tp.cnsub = a->coef_n_substrate;
            if (!a->coef) tp.pk = fold_props(*a->coef_props);
            tp.mode = a->mode;
            tp.wx = reinterpret_cast<const float4*>(a->wtab_x); tp.wy = reinterpret_cast<const float4*>(a->wtab_y);
            tp.wz = reinterpret_cast<const float4*>(a->wtab_z);
            tp.fsx = a->first_x; tp.fsy = a->first_y; tp.fsz = a->first_z;
            tp.ncx = p.ncx; tp.ncy = p.ncy; tp.ncz = p.ncz;
            tp.pcx = pc[0]; tp.pcy = pc[1]; tp.pcz = pc[2];
            int G = 1;
            while (G * 2 * pc[0] * pc[1] * pc[2] <= TILE_THREADS) G *= 2;
            tp.G = G;
            tp.cellsum = a->cellsum;
            // element sizes as the kernels derive them from the coordinate arrays are passed by the caller (hf, hc)
            const float hfx = a->hf[0], hfy = a->hf[1], hfz = a->hf[2];
            const float wq = (hfx * hfy * hfz) * 0.125f;
            tp.sw = a->mode == 1 ? a->scale * wq : wq;
            tp.inv_hfx = 1.0f / hfx; tp.inv_hfy = 1.0f / hfy; tp.inv_hfz = 1.0f / hfz;
            tp.inv_cvol = 1.0f / ((a->hc[0] * a->hc[1]) * a->hc[2]);
            dim3 grid((p.ncx + pc[0] - 1) / pc[0], (p.ncy + pc[1] - 1) / pc[1], (p.ncz + pc[2] - 1) / pc[2]);
            project_tile_kernel<<<grid, TILE_THREADS, 0, st>>>(tp), count_launch();
        }
    }
    if (tiled) {
        // (cellsum is complete: fall through to the node pass)
    } else if ((a->elems_per_cell_hint < 0 ? -a->elems_per_cell_hint : a->elems_per_cell_hint) <= 16) {
        const long long threads = ncell * 8;
        project_cells_kernel<8><<<(int)((threads + 127) / 128), 128, 0, st>>>(p), count_launch();
    } else {
        const long long threads = ncell * 32;
        project_cells_kernel<32><<<(int)((threads + 127) / 128), 128, 0, st>>>(p), count_launch();
    }
    int rc = check_launch("gomelt_project_f32 (cells)");
    if (rc) return rc;
    const long long nnode = (long long)(p.ncx + 1) * (p.ncy + 1) * (p.ncz + 1);
    project_nodes_kernel<<<grid_for(nnode, 256), 256, 0, st>>>(a->cellsum, p.c0x, p.c0y, p.c0z, p.ncx, p.ncy, p.ncz,
                                                               a->parent[0].n, a->parent[1].n, a->V, a->accumulate), count_launch();
    return check_launch("gomelt_project_f32 (nodes)");
}

extern "C" int gomelt_projected_source_f32(const gomelt_props_t* p, const gomelt_axis_t fine[3], const gomelt_axis_t parent[3],
                                           float wq_fine, const float* rows, int32_t n, float* tables, float* F,
                                           int32_t accumulate, void* stream) {
    if (!p || !fine || !parent || !rows || !tables || !F) {
        set_error("gomelt_projected_source_f32: NULL argument");
        return GOMELT_E_NULL;
    }
    if (n < 1 || n > GOMELT_MAX_SUBSTEPS) {
        set_error("gomelt_projected_source_f32: n = %d outside 1..%d", n, GOMELT_MAX_SUBSTEPS);
        return GOMELT_E_SIZE;
    }
    for (int d = 0; d < 3; ++d)
        if (!axis_ok(fine[d]) || !axis_ok(parent[d])) {
            set_error("gomelt_projected_source_f32: axis %d needs >= 2 nodes", d);
            return GOMELT_E_SIZE;
        }
    const float rcoeff = 1.f / (p->laser_radius * sqrtf((float)M_PI));
    const float dcoeff = 1.f / (p->laser_depth * sqrtf((float)M_PI));
    const float rsq = p->laser_radius * p->laser_radius, dsq = p->laser_depth * p->laser_depth;
    SrcBatch sb;
    sb.n = n;
    for (int r = 0; r < n; ++r) {
        const float* row = rows + 7 * (size_t)r;
        sb.v[r][0] = row[0]; sb.v[r][1] = row[1]; sb.v[r][2] = row[2];
        const float pcoeff = 6.f * sqrtf(3.f) * row[6] * p->laser_eta;   // computeSourceFunction_jax cF:1014
        sb.c[r] = (float)((double)(pcoeff * wq_fine) / (double)n);        // mean over the rows (cF:2726)
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int ncx = parent[0].n, ncy = parent[1].n, ncz = parent[2].n;
    const int nmax = ncx > ncy ? (ncx > ncz ? ncx : ncz) : (ncy > ncz ? ncy : ncz);
    coarse_source_table_batch_kernel<<<dim3((nmax + 63) / 64, 3, n), 64, 0, st>>>(
        fine[0].coords, fine[1].coords, fine[2].coords, fine[0].n, fine[1].n, fine[2].n, parent[0].coords, parent[1].coords,
        parent[2].coords, ncx, ncy, ncz, sb, 1.f / rsq, 1.f / dsq, rcoeff, dcoeff, tables), count_launch();
    if (ncz > 65535) {
        set_error("gomelt_projected_source_f32: parent grid too large (nz <= 65535)");
        return GOMELT_E_SIZE;
    }
    rank_n_kernel<<<dim3((ncx + 255) / 256, (ncy + ROWS3 - 1) / ROWS3, ncz), 256, 0, st>>>(F, tables, ncx, ncy, ncz, sb,
                                                                                       accumulate), count_launch();
    return check_launch("gomelt_projected_source_f32");
}

extern "C" int gomelt_shift_window_f32(const gomelt_shift_args_t* a, void* stream) {
    if (!a || !a->T1 || !a->Tp_old || !a->tx || !a->ty || !a->tz || !a->Tp_new || !a->T_new) {
        set_error("gomelt_shift_window_f32: NULL argument");
        return GOMELT_E_NULL;
    }
    for (int d = 0; d < 3; ++d)
        if (!axis_ok(a->L1[d]) || !axis_ok(a->old[d]) || (a->Tp_mid && !axis_ok(a->mid[d]))) {
            set_error("gomelt_shift_window_f32: axis %d needs >= 2 nodes", d);
            return GOMELT_E_SIZE;
        }
    if (a->ntx < 1 || a->nty < 1 || a->ntz < 1) {
        set_error("gomelt_shift_window_f32: empty target grid");
        return GOMELT_E_SIZE;
    }
    if (a->Tp_new == a->Tp_old || a->T_new == a->T1 || a->Tp_new == a->Tp_mid || a->T_new == a->Tp_old || a->T_new == a->Tp_mid) {
        set_error("gomelt_shift_window_f32: outputs must not alias the inputs");
        return GOMELT_E_FLAGS;
    }
    ShiftParams p;
    p.ax = {a->L1[0].coords, a->L1[0].n}; p.ay = {a->L1[1].coords, a->L1[1].n}; p.az = {a->L1[2].coords, a->L1[2].n};
    p.T1 = a->T1;
    p.mx = {a->mid[0].coords, a->mid[0].n}; p.my = {a->mid[1].coords, a->mid[1].n}; p.mz = {a->mid[2].coords, a->mid[2].n};
    p.Tpm = a->Tp_mid;
    p.ox = {a->old[0].coords, a->old[0].n}; p.oy = {a->old[1].coords, a->old[1].n}; p.oz = {a->old[2].coords, a->old[2].n};
    p.Tpo = a->Tp_old;
    p.tx = a->tx; p.ty = a->ty; p.tz = a->tz; p.ntx = a->ntx; p.nty = a->nty; p.ntz = a->ntz;
    p.Tp_new = a->Tp_new; p.T_new = a->T_new;
    if (a->ntz > 65535) {
        set_error("gomelt_shift_window_f32: target grid too large (ntz <= 65535)");
        return GOMELT_E_SIZE;
    }
    const int nxb = (a->ntx + 255) / 256, groups = row_groups(nxb, a->nty, a->ntz);
    shift_window_kernel<<<dim3(nxb, groups, a->ntz), 256, 0, (cudaStream_t)stream>>>(p, (a->nty + groups - 1) / groups), count_launch();
    return check_launch("gomelt_shift_window_f32");
}
