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
    for (long long w = blockIdx.x * (long long)blockDim.x + threadIdx.x; w < count;
         w += (long long)gridDim.x * blockDim.x) {
        int i, j, k;
        if (!p.faces_only) {
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
        const long long b = ex + (long long)ey * nnx + (long long)ez * nnxy;
        const long long nd[8] = {b, b + 1, b + 1 + nnx, b + nnx, b + nnxy, b + 1 + nnxy, b + 1 + nnx + nnxy,
                                 b + nnx + nnxy};
        float acc = 0.f;
        if (valid) {
#pragma unroll
            for (int a = 0; a < 8; ++a) {
                const float w = fminf(fmaxf(N[a], 0.f), 1.f);
                float v = p.u[nd[a]];
                if (p.u2) v = __fadd_rn(__fmul_rn(p.alpha, v), __fmul_rn(p.beta, p.u2[nd[a]]));
                acc = __fadd_rn(acc, __fmul_rn(w, v));
            }
        }
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

// parent node (gi,gj,gk) of the box [c0, c0 + nc] (nodes) <- its <= 8 adjacent cells, increasing cell id
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
    for (int e = 0; e < nf - 1; ++e) {  // same loop as coarse_source_table_kernel
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
    const long long total = (long long)nx * ny * nz;
    const int stride = nx + ny + nz;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(t % nx);
        const int j = (int)((t / nx) % ny);
        const int k = (int)(t / ((long long)nx * ny));
        float f = accumulate ? F[t] : 0.f;
        for (int r = 0; r < sb.n; ++r) {  // rows in order, like the row-by-row accumulation of gomelt_rank1_f32
            const float* tb = tables + (size_t)r * stride;
            f = f + sb.c[r] * ((tb[i] * tb[nx + j]) * tb[nx + ny + k]);
        }
        F[t] = f;
    }
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
// the interpolant of interp_kernel (same operations in the same order) at one point
__device__ __forceinline__ float trilinear_at(const AxisView& sx, const AxisView& sy, const AxisView& sz, const float* __restrict__ u,
                                              float x, float y, float z) {
    const float hx = __fsub_rn(sx.c[1], sx.c[0]), hy = __fsub_rn(sy.c[1], sy.c[0]), hz = __fsub_rn(sz.c[1], sz.c[0]);
    const float inv_vol = __fdiv_rn(1.0f, __fmul_rn(__fmul_rn(hx, hy), hz));
    const int nnx = sx.n, nnxy = sx.n * sy.n;
    const int ex = cell_of(x, sx.c[0], hx, sx.n - 1), ey = cell_of(y, sy.c[0], hy, sy.n - 1), ez = cell_of(z, sz.c[0], hz, sz.n - 1);
    const float ax0 = __fsub_rn(sx.c[ex + 1], x), ax1 = __fsub_rn(x, sx.c[ex]);
    const float ay0 = __fsub_rn(sy.c[ey + 1], y), ay1 = __fsub_rn(y, sy.c[ey]);
    const float az0 = __fsub_rn(sz.c[ez + 1], z), az1 = __fsub_rn(z, sz.c[ez]);
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
    if (!valid) return 0.f;
    const long long b = ex + (long long)ey * nnx + (long long)ez * nnxy;
    const long long nd[8] = {b, b + 1, b + 1 + nnx, b + nnx, b + nnxy, b + 1 + nnxy, b + 1 + nnx + nnxy, b + nnx + nnxy};
    float acc = 0.f;
#pragma unroll
    for (int a = 0; a < 8; ++a) acc = __fadd_rn(acc, __fmul_rn(fminf(fmaxf(N[a], 0.f), 1.f), u[nd[a]]));
    return acc;
}
__global__ void shift_window_kernel(const ShiftParams p) {
    const long long total = (long long)p.ntx * p.nty * p.ntz;
    for (long long w = blockIdx.x * (long long)blockDim.x + threadIdx.x; w < total; w += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(w % p.ntx), j = (int)((w / p.ntx) % p.nty), k = (int)(w / ((long long)p.ntx * p.nty));
        const float x = p.tx[i], y = p.ty[j], z = p.tz[k];
        const float tp = trilinear_at(p.ox, p.oy, p.oz, p.Tpo, x, y, z);
        const float t1 = trilinear_at(p.ax, p.ay, p.az, p.T1, x, y, z);
        float rest = tp;
        if (p.Tpm) rest = __fadd_rn(trilinear_at(p.mx, p.my, p.mz, p.Tpm, x, y, z), tp);  // T1on3 + (Tp2on3 + Tp3)
        p.Tp_new[w] = tp;
        p.T_new[w] = __fadd_rn(t1, rest);
    }
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
    interp_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(p), count_launch();
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
    if (elem_size == 4)
        box_copy_kernel<float><<<grid_for(total, 256), 256, 0, st>>>((const float*)src, (float*)dst, ix, iy, iz, nx, ny,
                                                                     nz, big_nx, big_ny, scatter), count_launch();
    else
        box_copy_kernel<uint8_t><<<grid_for(total, 256), 256, 0, st>>>((const uint8_t*)src, (uint8_t*)dst, ix, iy, iz,
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
    if (a->elems_per_cell_hint <= 16) {
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
    const long long total = (long long)ncx * ncy * ncz;
    rank_n_kernel<<<grid_for(total, 256), 256, 0, st>>>(F, tables, ncx, ncy, ncz, sb, accumulate), count_launch();
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
    const long long total = (long long)a->ntx * a->nty * a->ntz;
    shift_window_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(p), count_launch();
    return check_launch("gomelt_shift_window_f32");
}
