// Shared device/host helpers for libgomelt_sm100 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "gomelt_abi.h"

#ifndef GOMELT_SM_COUNT
#define GOMELT_SM_COUNT 148  // fallback when the device attribute cannot be read (B200: 2 dies x 74 SMs)
#endif

namespace gomelt {

void set_error(const char* fmt, ...);
int sm_count();  // multiprocessors of the current device (cudaDevAttrMultiProcessorCount, cached)
void count_launch(int n = 1);  // process-wide count of kernels launched by the library (gomelt_launch_count)

inline int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return (int)e;
    }
    return 0;
}

// Pre-folded constants of computeStateProperties (cF:2567-2614): the /1000 of k and the
// rho* of rhocp are folded into the coefficients on the host (double -> float).
struct PropK {
    float T_liq, T_sol, T_amb;
    float k_powder, k_a0, k_a1, k_fluid;  // already / 1000
    float c_a0, c_a1, c_mushy, c_fluid;   // already * rho
};

inline PropK fold_props(const gomelt_props_t& p) {
    PropK q;
    q.T_liq = p.T_liquidus;
    q.T_sol = p.T_solidus;
    q.T_amb = p.T_amb;
    q.k_powder = (float)((double)p.k_powder / 1000.0);
    q.k_a0 = (float)((double)p.k_bulk_a0 / 1000.0);
    q.k_a1 = (float)((double)p.k_bulk_a1 / 1000.0);
    q.k_fluid = (float)((double)p.k_fluid / 1000.0);
    q.c_a0 = (float)((double)p.rho * (double)p.cp_solid_a0);
    q.c_a1 = (float)((double)p.rho * (double)p.cp_solid_a1);
    q.c_mushy = (float)((double)p.rho * (double)p.cp_mushy);
    q.c_fluid = (float)((double)p.rho * (double)p.cp_fluid);
    return q;
}

// computeConvRadBC cF:2207-2301: constants of the top-surface flux, folded once on the host.
struct FluxK {
    float T_amb, T_cap, invT_b, Lev, cp_fluid, CM, CT, evcCP, h_conv, sig_eps, Tamb4;
    float wq;  // hx*hy/4
};

inline FluxK fold_flux(const gomelt_props_t& p, const gomelt_grid_t& g) {
    FluxK fk;
    fk.T_amb = p.T_amb;
    fk.T_cap = p.T_boiling + 1000.f;
    fk.invT_b = 1.0f / p.T_boiling;
    fk.Lev = p.Lev;
    fk.cp_fluid = p.cp_fluid;
    fk.CM = p.CM_coeff;
    fk.CT = p.CT_coeff;
    fk.evcCP = p.evc * p.CP_coeff;
    fk.h_conv = p.h_conv;
    fk.sig_eps = p.sigma_sb * p.vareps;
    const float Ta2 = p.T_amb * p.T_amb;
    fk.Tamb4 = Ta2 * Ta2;
    fk.wq = (g.hx * g.hy) * 0.25f;
    return fk;
}

// convection + radiation + evaporation flux at one surface Gauss point (cF:2270-2293)
__device__ __forceinline__ float flux_at(const FluxK& f, float Tq) {
    Tq = fminf(Tq, f.T_cap);
    const float invT = 1.0f / Tq;
    const float E_pv = f.Lev + f.cp_fluid * (Tq - f.T_amb);
    const float MolMot = sqrtf(f.CM * invT);
    const float S = f.evcCP * expf(-f.CT * (invT - f.invT_b)) * MolMot * E_pv;
    const float T2 = Tq * Tq;
    float q = f.h_conv * (f.T_amb - Tq) + f.sig_eps * (f.Tamb4 - T2 * T2) - S;
    return q * 1e-6f;
}

// State + properties of one node.  S1 in: float (thresholded at 0.499), forced to 1 on substrate.
__device__ __forceinline__ void node_props(const PropK& q, float T, float S1in, bool substrate,
                                           float& k, float& rhocp, bool& s1, bool& s2) {
    s2 = (T >= q.T_liq);
    const bool s3 = (T > q.T_sol) && (T < q.T_liq);
    s1 = (S1in > 0.499f) || s2 || substrate;
    const float kb = fmaf(q.k_a1, T, q.k_a0);
    k = s2 ? q.k_fluid : (s1 ? kb : q.k_powder);
    const float cs = fmaf(q.c_a1, T, q.c_a0);
    rhocp = s2 ? q.c_fluid : (s3 ? q.c_mushy : cs);
}

}  // namespace gomelt
