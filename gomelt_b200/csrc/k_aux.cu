// Small per-level kernels: stand-alone state/properties (a5), top-surface flux (a11),
// separable source tables (K6 / a7-a9), error plumbing and an FP32 issue-rate diagnostic.
#include <math.h>
#include <stdarg.h>
#include <atomic>
#include <string.h>

#include "common.cuh"

namespace gomelt {

static thread_local char g_err[512] = "";

static std::atomic<long long> g_launches{0};
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
long long launches_so_far() { return g_launches.load(std::memory_order_relaxed); }

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// ---- computeStateProperties cF:2567-2614 ------------------------------------------------
__global__ void state_props_kernel(PropK pk, const float* __restrict__ T, const float* __restrict__ S1,
                                   long long nn, long long nsub, float* S1o, uint8_t* S2o, float* ko,
                                   float* ro) {
    for (long long n = blockIdx.x * (long long)blockDim.x + threadIdx.x; n < nn;
         n += (long long)gridDim.x * blockDim.x) {
        float k, r;
        bool s1, s2;
        node_props(pk, T[n], S1[n], n < nsub, k, r, s1, s2);
        if (S1o) S1o[n] = s1 ? 1.f : 0.f;
        if (S2o) S2o[n] = s2 ? 1 : 0;
        if (ko) ko[n] = k;
        if (ro) ro[n] = r;
    }
}

__global__ void clamp_min_kernel(float* __restrict__ x, long long n, float lo) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        x[i] = fmaxf(x[i], lo);
}

// ---- computeConvRadBC cF:2207-2301 -------------------------------------------------------
// One thread per top-plane node; gathers its <= 4 adjacent top elements in increasing element
// id (the order the reference's scatter-add applies them), recomputing each element's 4 Gauss
// fluxes: deterministic, no atomics; the top plane is 1/nz of the level.
__global__ void surface_flux_kernel(FluxK fk, int nx, int ny, const float* __restrict__ Tp /* plane */,
                                    float* __restrict__ flux, int add) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= nx || j >= ny) return;
    const float g = 0.57735026918962576f;  // 1/sqrt(3)
    // N[q][a] = 1/4 (1 + xi_q xi_a)(1 + eta_q eta_a), local order (-,-),(+,-),(+,+),(-,+)
    const float sx[4] = {-1.f, 1.f, 1.f, -1.f}, sy[4] = {-1.f, -1.f, 1.f, 1.f};
    float acc = 0.f;
    for (int ey = j - 1; ey <= j; ++ey) {
        for (int ex = i - 1; ex <= i; ++ex) {
            if (ex < 0 || ey < 0 || ex >= nx - 1 || ey >= ny - 1) continue;
            const float Ta[4] = {Tp[ey * nx + ex], Tp[ey * nx + ex + 1], Tp[(ey + 1) * nx + ex + 1],
                                 Tp[(ey + 1) * nx + ex]};
            const int ax = i - ex, ay = j - ey;          // local corner of this node
            const int a = ay == 0 ? (ax == 0 ? 0 : 1) : (ax == 0 ? 3 : 2);
            float contrib = 0.f;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float Tq = 0.f;
#pragma unroll
                for (int b = 0; b < 4; ++b)
                    Tq += 0.25f * (1.f + g * sx[q] * sx[b]) * (1.f + g * sy[q] * sy[b]) * Ta[b];
                const float Nqa = 0.25f * (1.f + g * sx[q] * sx[a]) * (1.f + g * sy[q] * sy[a]);
                contrib += Nqa * (flux_at(fk, Tq) * fk.wq);
            }
            acc += contrib;
        }
    }
    const int n = j * nx + i;
    flux[n] = add ? flux[n] + acc : acc;
}

// ---- K6: separable source tables ----------------------------------------------------------
// t[i] = sum_{e in {i-1,i}} sum_{q in {0,1}} N1[q][a] * c * exp(-3 (x_q - v)^2 / s^2),
// x_q = N1[q][0] x_e + N1[q][1] x_{e+1}  (getQuadratureCoords cF:3151-3166), N1[q][a] = (1 +- 1/sqrt3)/2.
__device__ __forceinline__ float source_table_value(const float* __restrict__ x, int n, int i, float v,
                                                    float inv_s2_3, float c) {
    const float g = 0.57735026918962576f;
    const float Nlo = 0.5f * (1.f + g), Nhi = 0.5f * (1.f - g);  // weight of the near / far node
    float acc = 0.f;
    for (int e = i - 1; e <= i; ++e) {
        if (e < 0 || e >= n - 1) continue;
        const float x0 = x[e], x1 = x[e + 1];
        const float xq0 = Nlo * x0 + Nhi * x1;  // Gauss point near x0
        const float xq1 = Nhi * x0 + Nlo * x1;  // Gauss point near x1
        const float Q0 = c * expf(-3.f * (xq0 - v) * (xq0 - v) * inv_s2_3);
        const float Q1 = c * expf(-3.f * (xq1 - v) * (xq1 - v) * inv_s2_3);
        // node i is local corner 1 of element i-1, corner 0 of element i
        acc += (e == i) ? (Nlo * Q0 + Nhi * Q1) : (Nhi * Q0 + Nlo * Q1);
    }
    return acc;
}

// All three axes of up to GOMELT_MAX_SUBSTEPS laser positions in ONE launch: blockIdx.y = axis,
// blockIdx.z = laser row; tables of row s live at tables + s * (nx + ny + nz) as [tx | ty | tz].
struct TableBatch {
    float v[GOMELT_MAX_SUBSTEPS][3];
};
__global__ void source_table_batch_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                          const float* __restrict__ z, int nx, int ny, int nz,
                                          const __grid_constant__ TableBatch tb, float inv_r2, float inv_d2, float rc,
                                          float dc, float* __restrict__ tables) {
    const int axis = blockIdx.y, s = blockIdx.z;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const float* c = axis == 0 ? x : (axis == 1 ? y : z);
    const int n = axis == 0 ? nx : (axis == 1 ? ny : nz);
    if (i >= n) return;
    float* out = tables + (size_t)s * (nx + ny + nz) + (axis == 0 ? 0 : (axis == 1 ? nx : nx + ny));
    out[i] = source_table_value(c, n, i, tb.v[s][axis], axis == 2 ? inv_d2 : inv_r2, axis == 2 ? dc : rc);
}

// ---- FP32 issue-rate diagnostic -----------------------------------------------------------
template <int KIND>
__global__ void fp32_rate_kernel(int iters, float* sink, float bin, float cin) {
    float a0 = threadIdx.x * 1e-3f, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f;
    float a4 = a0 + 4.f, a5 = a0 + 5.f, a6 = a0 + 6.f, a7 = a0 + 7.f;
    // b, c live in registers (thread-dependent) so that KIND 0 measures the 3-register FFMA form
    const float b = bin + (float)threadIdx.x * 1e-30f, c = cin + (float)threadIdx.x * 1e-30f;
    for (int it = 0; it < iters; ++it) {
        if (KIND == 0) {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                a0 = fmaf(a0, b, c); a1 = fmaf(a1, b, c); a2 = fmaf(a2, b, c); a3 = fmaf(a3, b, c);
                a4 = fmaf(a4, b, c); a5 = fmaf(a5, b, c); a6 = fmaf(a6, b, c); a7 = fmaf(a7, b, c);
            }
        } else if (KIND == 1) {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                a0 += c; a1 += c; a2 += c; a3 += c; a4 += c; a5 += c; a6 += c; a7 += c;
            }
        } else {
            float2 p0 = make_float2(a0, a1), p1 = make_float2(a2, a3), p2 = make_float2(a4, a5),
                   p3 = make_float2(a6, a7);
            const float2 bb = make_float2(b, b), cc = make_float2(c, c);
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                p0 = __ffma2_rn(p0, bb, cc); p1 = __ffma2_rn(p1, bb, cc);
                p2 = __ffma2_rn(p2, bb, cc); p3 = __ffma2_rn(p3, bb, cc);
            }
            a0 = p0.x; a1 = p0.y; a2 = p1.x; a3 = p1.y; a4 = p2.x; a5 = p2.y; a6 = p3.x; a7 = p3.y;
        }
    }
    const float s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (s == 12345.678f) sink[0] = s;
}

// ---- min / max monitor (printLevelMaxMin cF:3635-3665 as one reduction; N3 of SURVEY.md 8(f)) -----------
__device__ __forceinline__ void atomic_min_f32(float* a, float v) {
    if (v >= 0.f) atomicMin(reinterpret_cast<int*>(a), __float_as_int(v));
    else atomicMax(reinterpret_cast<unsigned*>(a), __float_as_uint(v));
}
__device__ __forceinline__ void atomic_max_f32(float* a, float v) {
    if (v >= 0.f) atomicMax(reinterpret_cast<int*>(a), __float_as_int(v));
    else atomicMin(reinterpret_cast<unsigned*>(a), __float_as_uint(v));
}
__global__ void minmax_init_kernel(float* out) {
    out[0] = __int_as_float(0x7f800000);  // +inf
    out[1] = __int_as_float(0xff800000);  // -inf
    out[2] = 0.f;
}
__global__ void minmax_kernel(const float* __restrict__ x, long long n, float* out) {
    float lo = __int_as_float(0x7f800000), hi = __int_as_float(0xff800000);
    int bad = 0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float v = x[i];
        if (isfinite(v)) {
            lo = fminf(lo, v);
            hi = fmaxf(hi, v);
        } else {
            ++bad;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        bad += __shfl_xor_sync(0xffffffffu, bad, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (lo <= hi) {  // the warp saw at least one finite value
            atomic_min_f32(out, lo);
            atomic_max_f32(out + 1, hi);
        }
        if (bad) atomicAdd(out + 2, (float)bad);
    }
}

}  // namespace gomelt

using namespace gomelt;

extern "C" const char* gomelt_last_error(void) { return g_err; }
extern "C" int gomelt_minmax_f32(const float* x, int64_t n, float* out3, void* stream) {
    if (!x || !out3 || n < 1) {
        set_error("gomelt_minmax_f32: NULL argument / empty field");
        return GOMELT_E_NULL;
    }
    cudaStream_t st = (cudaStream_t)stream;
    minmax_init_kernel<<<1, 1, 0, st>>>(out3), count_launch();
    long long blocks = (n + 4 * 256 - 1) / (4 * 256);
    if (blocks > 8 * sm_count()) blocks = 8 * sm_count();
    minmax_kernel<<<(int)blocks, 256, 0, st>>>(x, n, out3), count_launch();
    return check_launch("gomelt_minmax_f32");
}

extern "C" int gomelt_clamp_min_f32(float* x, int64_t n, float lo, void* stream) {
    if (!x || n < 1) {
        set_error("gomelt_clamp_min_f32: NULL argument / empty field");
        return GOMELT_E_NULL;
    }
    long long blocks = (n + 255) / 256;
    if (blocks > 8 * sm_count()) blocks = 8 * sm_count();
    clamp_min_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(x, n, lo), count_launch();
    return check_launch("gomelt_clamp_min_f32");
}

extern "C" int gomelt_abi_version(void) { return GOMELT_ABI_VERSION; }
namespace gomelt { long long launches_so_far(); }
extern "C" long long gomelt_launch_count(void) { return gomelt::launches_so_far(); }

extern "C" int gomelt_state_props_f32(const gomelt_props_t* props, const float* T, const float* S1, int64_t nn,
                                      int64_t n_substrate, float* S1_out, uint8_t* S2_out, float* k_out,
                                      float* rhocp_out, void* stream) {
    if (!props || !T || !S1) {
        set_error("gomelt_state_props_f32: NULL props/T/S1");
        return GOMELT_E_NULL;
    }
    if (nn <= 0) {
        set_error("gomelt_state_props_f32: nn = %lld", (long long)nn);
        return GOMELT_E_SIZE;
    }
    const int threads = 256;
    const long long want = (nn + threads - 1) / threads;
    const int blocks = (int)(want < sm_count() * 8 ? want : sm_count() * 8);
    state_props_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(fold_props(*props), T, S1, nn, n_substrate,
                                                                     S1_out, S2_out, k_out, rhocp_out), count_launch();
    return check_launch("gomelt_state_props_f32");
}

extern "C" int gomelt_surface_flux_f32(const gomelt_props_t* p, const gomelt_grid_t* g, const float* T0,
                                       int32_t nz_active, float* flux, int32_t add, void* stream) {
    if (!p || !g || !T0 || !flux) {
        set_error("gomelt_surface_flux_f32: NULL argument");
        return GOMELT_E_NULL;
    }
    if (g->nx < 2 || g->ny < 2 || nz_active < 2 || nz_active > g->nz) {
        set_error("gomelt_surface_flux_f32: bad grid / nz_active %d", nz_active);
        return GOMELT_E_SIZE;
    }
    const FluxK fk = fold_flux(*p, *g);
    dim3 block(32, 8), grid((g->nx + 31) / 32, (g->ny + 7) / 8);
    const float* plane = T0 + (long long)(nz_active - 1) * g->nx * g->ny;
    surface_flux_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(fk, g->nx, g->ny, plane, flux, add), count_launch();
    return check_launch("gomelt_surface_flux_f32");
}

extern "C" int gomelt_source_tables_f32(const gomelt_props_t* p, const gomelt_grid_t* g, const float* x,
                                        const float* y, const float* z, const float laser_xyz[3], float laserP,
                                        float* tx, float* ty, float* tz, float* coef, void* stream) {
    if (!p || !g || !x || !y || !z || !laser_xyz || !tx || !ty || !tz || !coef) {
        set_error("gomelt_source_tables_f32: NULL argument");
        return GOMELT_E_NULL;
    }
    // computeSourceFunction_jax cF:1014-1018, float32 like the traced scalars
    const float pcoeff = 6.f * sqrtf(3.f) * laserP * p->laser_eta;
    const float rcoeff = 1.f / (p->laser_radius * sqrtf((float)M_PI));
    const float dcoeff = 1.f / (p->laser_depth * sqrtf((float)M_PI));
    const float rsq = p->laser_radius * p->laser_radius, dsq = p->laser_depth * p->laser_depth;
    const float wq = (g->hx * g->hy * g->hz) * 0.125f;
    *coef = pcoeff * wq;
    // one launch: the three tables are one "batch" of a single laser row writing to tx / ty / tz
    TableBatch tb;
    tb.v[0][0] = laser_xyz[0]; tb.v[0][1] = laser_xyz[1]; tb.v[0][2] = laser_xyz[2];
    const int nmax = g->nx > g->ny ? (g->nx > g->nz ? g->nx : g->nz) : (g->ny > g->nz ? g->ny : g->nz);
    if (ty == tx + g->nx && tz == ty + g->ny) {
        source_table_batch_kernel<<<dim3((nmax + 127) / 128, 3, 1), 128, 0, (cudaStream_t)stream>>>(
            x, y, z, g->nx, g->ny, g->nz, tb, 1.f / rsq, 1.f / dsq, rcoeff, dcoeff, tx), count_launch();
    } else {  // separately allocated tables: one axis per launch
        const float* cs[3] = {x, y, z};
        float* ts[3] = {tx, ty, tz};
        const int ns[3] = {g->nx, g->ny, g->nz};
        for (int d = 0; d < 3; ++d) {
            TableBatch t1;
            t1.v[0][0] = t1.v[0][1] = t1.v[0][2] = laser_xyz[d];
            // a 1-axis launch: present axis d as "x" of a grid with ny = nz = 0 blocks
            source_table_batch_kernel<<<dim3((ns[d] + 127) / 128, 1, 1), 128, 0, (cudaStream_t)stream>>>(
                cs[d], cs[d], cs[d], ns[d], 0, 0, t1, d == 2 ? 1.f / dsq : 1.f / rsq, 1.f / dsq,
                d == 2 ? dcoeff : rcoeff, dcoeff, ts[d]), count_launch();
        }
    }
    return check_launch("gomelt_source_tables_f32");
}

extern "C" int gomelt_source_tables_batch_f32(const gomelt_props_t* p, const gomelt_grid_t* g, const float* x,
                                              const float* y, const float* z, const float* rows, int32_t n,
                                              float* tables, float* coef, void* stream) {
    if (!p || !g || !x || !y || !z || !rows || !tables || !coef) {
        set_error("gomelt_source_tables_batch_f32: NULL argument");
        return GOMELT_E_NULL;
    }
    if (n < 1 || n > GOMELT_MAX_SUBSTEPS) {
        set_error("gomelt_source_tables_batch_f32: n = %d outside 1..%d", n, GOMELT_MAX_SUBSTEPS);
        return GOMELT_E_SIZE;
    }
    const float rcoeff = 1.f / (p->laser_radius * sqrtf((float)M_PI));
    const float dcoeff = 1.f / (p->laser_depth * sqrtf((float)M_PI));
    const float rsq = p->laser_radius * p->laser_radius, dsq = p->laser_depth * p->laser_depth;
    const float wq = (g->hx * g->hy * g->hz) * 0.125f;
    TableBatch tb;
    for (int s = 0; s < n; ++s) {
        const float* row = rows + 7 * (size_t)s;
        tb.v[s][0] = row[0]; tb.v[s][1] = row[1]; tb.v[s][2] = row[2];
        coef[s] = (6.f * sqrtf(3.f) * row[6] * p->laser_eta) * wq;  // computeSourceFunction_jax cF:1014
    }
    const int nmax = g->nx > g->ny ? (g->nx > g->nz ? g->nx : g->nz) : (g->ny > g->nz ? g->ny : g->nz);
    source_table_batch_kernel<<<dim3((nmax + 127) / 128, 3, n), 128, 0, (cudaStream_t)stream>>>(
        x, y, z, g->nx, g->ny, g->nz, tb, 1.f / rsq, 1.f / dsq, rcoeff, dcoeff, tables), count_launch();
    return check_launch("gomelt_source_tables_batch_f32");
}

extern "C" int gomelt_diag_fp32_rate(int32_t kind, int32_t iters, int32_t blocks, int32_t threads, float* sink,
                                     double* ops, void* stream) {
    if (!sink || !ops) {
        set_error("gomelt_diag_fp32_rate: NULL argument");
        return GOMELT_E_NULL;
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (kind == 0) fp32_rate_kernel<0><<<blocks, threads, 0, st>>>(iters, sink, 1.0000001f, 1e-7f), count_launch();
    else if (kind == 1) fp32_rate_kernel<1><<<blocks, threads, 0, st>>>(iters, sink, 1.0000001f, 1e-7f), count_launch();
    else fp32_rate_kernel<2><<<blocks, threads, 0, st>>>(iters, sink, 1.0000001f, 1e-7f), count_launch();
    *ops = (double)iters * 64.0 * (double)blocks * (double)threads;
    return check_launch("gomelt_diag_fp32_rate");
}
