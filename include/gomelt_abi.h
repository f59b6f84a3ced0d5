/*
 * gomelt_abi.h — C ABI of libgomelt_sm100.so, the B200 (sm_100a) implementation of
 * GO-MELT's multilevel explicit FE thermal time-step.
 *
 * The reference (JLnorthwestern/GO-MELT) has no FFI: its "operator API" for this path is the
 * set of computeFunctions.py entry points the driver calls.  Each entry
 * point below names the reference function(s) it replaces (cF = go_melt/computeFunctions.py,
 * gm = go_melt/go_melt.py).  The same symbols are what an XLA-FFI shim or a ctypes/cffi
 * binding would bind.
 *
 * Conventions
 *   - plain C: POD structs, raw device pointers, sizes; no torch / C++ types.
 *   - every launcher is stream-ordered (`stream` is a cudaStream_t passed as void*),
 *     never allocates, never synchronises, and is re-entrant.
 *   - return value: 0 = ok, <0 = bad argument (GOMELT_E_*), >0 = cudaError_t of the launch.
 *     gomelt_last_error() returns a thread-local message for the last non-zero return.
 *   - all fields are float32 / int32 like the reference (jax_enable_x64 is never set, cF:15).
 *   - node numbering is x-fastest: n = ix + iy*nx + iz*nx*ny (cF:646-689).
 */
#ifndef GOMELT_ABI_H
#define GOMELT_ABI_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GOMELT_ABI_VERSION 2

#define GOMELT_E_NULL   (-1) /* required pointer is NULL            */
#define GOMELT_E_SIZE   (-2) /* grid / count out of range           */
#define GOMELT_E_FLAGS  (-3) /* inconsistent flag combination       */
#define GOMELT_E_ALIGN  (-4) /* pointer not aligned as required     */

/* Uniform structured hex8 grid of one level (cF:115-157: elements, nodes, h). */
typedef struct gomelt_grid {
    int32_t nx, ny, nz;   /* nodes per axis (= elements + 1)                    */
    float   hx, hy, hz;   /* element size, float32 as Level["h"] holds it        */
} gomelt_grid_t;

/* Material / process constants, float32 exactly as they are inside jax.jit
 * (SetupProperties cF:267-345; derived CM/CT/CP cF:339-343; h_conv already *1e6 cF:319). */
typedef struct gomelt_props {
    float k_powder, k_bulk_a0, k_bulk_a1, k_fluid;          /* W/m K            */
    float cp_solid_a0, cp_solid_a1, cp_mushy, cp_fluid;     /* J/kg K           */
    float rho;                                              /* kg/mm^3          */
    float T_amb, T_solidus, T_liquidus, T_boiling;          /* K                */
    float h_conv, sigma_sb, vareps, evc, Lev;               /* surface terms    */
    float CM_coeff, CT_coeff, CP_coeff;
    float laser_radius, laser_depth, laser_eta;             /* mm, mm, -        */
} gomelt_props_t;

/* ---- K1: fused level step -------------------------------------------------------------
 * Replaces, for one level and one sweep:
 *   computeStateProperties cF:2567-2614  (S1/S2/k/rhocp from T0,S1 — evaluated per staged node)
 *   solveMatrixFreeFE      cF:582-642    (element-mean k, rhocp; lumped mass; forward Euler)
 *   substitute_Tbar        cF:1848-1865  (planes >= nz_active <- T_amb)
 *   assignBCs              cF:1568-1595  (GOMELT_STEP_BC_CONST: 5 faces <- bc5[y-,y+,x-,x+,z-])
 *   the "+ Fc + Corr" of cF:642 (rhs array, rank-1 source tables, top-plane flux load)
 *   jnp.maximum(T_amb, .)  cF:2360-2362 etc. (GOMELT_STEP_CLAMP)
 *   melt-time bookkeeping  cF:3568-3578  (GOMELT_STEP_ACCUM)
 * Child levels: GOMELT_STEP_SKIP_FACES leaves the 5 Dirichlet faces of T_out untouched;
 * gomelt_face_bc_f32 (assignBCsFine cF:1598-1620) writes them.
 */
#define GOMELT_STEP_CLAMP      0x01  /* T_out = max(T_amb, T_out)                              */
#define GOMELT_STEP_WRITE_S1   0x02  /* write thresholded S1 (f32 0/1) to S1_out               */
#define GOMELT_STEP_WRITE_S2   0x04  /* write S2 = (T0 >= T_liquidus) as uint8 to S2_out       */
#define GOMELT_STEP_BC_CONST   0x08  /* Level-1 Dirichlet constants on 5 faces (all planes)    */
#define GOMELT_STEP_SKIP_FACES 0x10  /* do not write the 5 Dirichlet faces (child levels)      */
#define GOMELT_STEP_ACCUM      0x20  /* update accum / max_accum with S2_prev -> S2 transition */
#define GOMELT_STEP_FUSED_FLUX 0x40  /* computeConvRadBC cF:2207-2301 evaluated inside the step from T0 on plane
                                        nz_active-1 and added to that plane's load (instead of `topflux`)   */
#define GOMELT_STEP_GENERAL_KERNEL 0x80 /* run the general (natural-boundary capable) kernel even when the call
                                        qualifies for the Dirichlet-side-face fast kernel; results agree to f32
                                        rounding (tests A/B the two kernels through this bit)               */
#define GOMELT_STEP_NO_COLD_PLANES 0x100 /* fast kernel: evaluate the general property selects on every plane,
                                        also where the warp voted "nothing above the solidus" (bit-identical
                                        results; tests and timings A/B the cold-plane path through this bit) */

typedef struct gomelt_step_args {
    gomelt_grid_t grid;
    const float  *T0;          /* [nn] temperature at t                                         */
    const float  *S1;          /* [nn] powder(0)/bulk(1) state, float as in the reference       */
    const float  *rhs;         /* [nn] or NULL: nodal load (projected source + T' corrections)  */
    const float  *src_x, *src_y, *src_z; /* rank-1 source tables [nx],[ny],[nz] or all NULL     */
    float         src_coef;    /* F[n] += src_coef * src_x[ix]*src_y[iy]*src_z[iz]              */
    const float  *topflux;     /* [nx*ny] or NULL: surface load added on plane nz_active-1      */
    float         dt;
    int32_t       nz_active;   /* planes [0,nz_active) are active (tmp_ne_nn, cF:495-517); 0..nz */
    int64_t       n_substrate; /* S1 := 1 for node ids < n_substrate (cF:2592)                  */
    int32_t       flags;       /* GOMELT_STEP_*                                                 */
    float         bc5[5];      /* y-, y+, x-, x+, z- values for GOMELT_STEP_BC_CONST            */
    float        *T_out;       /* [nn] temperature at t+dt (must not alias T0)                  */
    float        *S1_out;      /* [nn] or NULL (may alias S1)                                   */
    uint8_t      *S2_out;      /* [nn] or NULL                                                  */
    const uint8_t *S2_prev;    /* [nn] previous melt flag (ACCUM); may alias S2_out             */
    float        *accum;       /* [nn] accumulated melt time, updated in place (ACCUM)          */
    float        *max_accum;   /* [nn] max accumulated melt time, updated in place (ACCUM)      */
    int32_t       z_chunk;     /* planes per z-chunk, 0 = library default                       */
    int32_t       z_begin, z_end; /* finalise planes [z_begin, z_end) only; 0,0 = all planes.  A
                                * z-slab rank passes its local array (ghost planes included) and
                                * the owned range: planes z_begin-1 and z_end are read as ghosts;
                                * nz_active / n_substrate are then in local planes / node ids.   */
    float        *peer_lo;     /* NULL, or [nx*ny] in a z-neighbour GPU's memory (peer-mapped, NVLink): the
                                * finalised plane z_begin of T_out is ALSO stored there - the lower
                                * neighbour's upper ghost plane - by the same kernel (halo exchange fused
                                * into the step; no separate copy / collective).                         */
    float        *peer_hi;     /* likewise plane z_end-1 -> the upper neighbour's lower ghost plane      */
    /* Halo exchange with release / acquire counters (ABI 2; Level-1 z-slab ranks, GOMELT_STEP_BC_CONST).  With
     * halo_sync set the call is the step AND the exchange of its two boundary planes, entirely on the device: after the
     * step one more kernel copies the first / last owned plane of T_out (Dirichlet face constants included) into the
     * neighbours' ghost planes (peer_lo / peer_hi, NVLink peer stores), bumps the neighbours' arrival counters with a
     * system-scope release and waits - acquire - until this rank's counters show that both neighbours' planes of the
     * same sweep have arrived.  When the call has completed on the stream, the ghost planes of T_out are current: the
     * next sweep can be issued at once; no barrier launch, no NCCL call, no host involvement.  It is a neighbour
     * collective: every rank of the slab decomposition must issue the same sequence of calls.  Contract: two
     * temperature buffers that alternate as T0 / T_out, every rank's counter block zeroed once (then a real barrier),
     * the ghost planes of the first T0 filled by the caller, halo_seq = 0, 1, 2, ... in step on all ranks.  S1 ghosts
     * are the caller's business (they do not change in a sweep).  NULL = off. */
    uint32_t     *halo_sync;    /* [gomelt_halo_sync_words()] this rank's counter block, peer-visible memory            */
    uint32_t     *halo_sync_lo; /* the lower neighbour's counter block (peer-mapped), NULL on the lowest rank           */
    uint32_t     *halo_sync_hi; /* the upper neighbour's, NULL on the highest rank                                       */
    uint32_t      halo_seq;     /* number of sweeps this slab has done under the protocol before this one               */
    /* Optional scratch of the melt-time bookkeeping (GOMELT_STEP_ACCUM on the fast kernel): with it the step kernel only
     * QUEUES the planes of a tile that hold molten nodes and a second launch does their accum / max_accum / S2 update with
     * one warp per (tile, plane) - the warps over the melt pool are otherwise the tail of a one-wave launch (72.6 -> ~55 us
     * per 10 M-node corrector sweep with one melt pool).  bk_queue_words >= 2 + 2 * entries uint32 words; one entry per
     * (60 x 4-node tile, plane) that is hot: nn / 120 + 1024 entries always suffice.  NULL: everything inside the step. */
    uint32_t     *bk_queue;
    int64_t       bk_queue_words;
    int32_t       bk_queue_keep;  /* bit 0 clear: the call zeroes the queue's two header words first; set: the caller
                                   * guarantees they are zero - every call that uses the queue leaves them zeroed (saves a
                                   * memset node per sweep inside a block of substeps).  Bit 1: use the queue whatever the
                                   * grid size (by default only launches that fill the GPU use it: on a small window the
                                   * second launch costs more than the tail it removes) */
} gomelt_step_args_t;

int gomelt_level_step_f32(const gomelt_props_t *props, const gomelt_step_args_t *args, void *stream);

/* words of a halo counter block (two arrival counters, padded) */
#define GOMELT_HALO_SYNC_HEAD 32
long long gomelt_halo_sync_words(void);

/* computeStateProperties cF:2567-2614 as a stand-alone op (outputs may be NULL). */
int gomelt_state_props_f32(const gomelt_props_t *props, const float *T, const float *S1, int64_t nn,
                           int64_t n_substrate, float *S1_out, uint8_t *S2_out, float *k_out,
                           float *rhocp_out, void *stream);

/* computeConvRadBC cF:2207-2301: convection + radiation + evaporation load of the top face of
 * element layer (nz_active-2) on the nodes of plane nz_active-1.  flux[nx*ny] is overwritten
 * (add = 0) or accumulated into (add = 1). */
int gomelt_surface_flux_f32(const gomelt_props_t *props, const gomelt_grid_t *grid, const float *T0,
                            int32_t nz_active, float *flux, int32_t add, void *stream);

/* K6 — computeSourcesL3 cF:2960-3012 / computeSourceFunction_jax cF:991-1025 as three 1-D
 * tables (the Gaussian is separable and N is a tensor product):
 *   F[n] = coef * tx[ix]*ty[iy]*tz[iz],  coef returned in *coef (host float).
 * x,y,z are the level's node-coordinate arrays on the device. */
int gomelt_source_tables_f32(const gomelt_props_t *props, const gomelt_grid_t *grid, const float *x,
                             const float *y, const float *z, const float laser_xyz[3], float laserP,
                             float *tx, float *ty, float *tz, float *coef, void *stream);

/* Batched K6: the tables of n laser rows in ONE launch.  rows = HOST array [n][7] of toolpath rows
 * (x, y, z, Ljump, Ldwell, dt, P - cP:71-74); tables = device [n][nx + ny + nz] laid out [tx | ty | tz] per
 * row; coef = HOST [n] out.  n <= GOMELT_MAX_SUBSTEPS. */
#define GOMELT_MAX_SUBSTEPS 64
int gomelt_source_tables_batch_f32(const gomelt_props_t *props, const gomelt_grid_t *grid, const float *x,
                                   const float *y, const float *z, const float *rows, int32_t n, float *tables,
                                   float *coef, void *stream);

/* ---- inter-level transfers (no materialised operators; see DESIGN.md "Transfers") ---------------------
 * A level is seen through its three 1-D node-coordinate arrays (Level["node_coords"], cF:32-64). */
typedef struct gomelt_axis {
    const float *coords;   /* [n] device array of node coordinates along the axis */
    int32_t      n;        /* nodes (= elements + 1)                              */
} gomelt_axis_t;

#define GOMELT_INTERP_SET  0   /* out[o] = I                 */
#define GOMELT_INTERP_ADD  1   /* out[o] += I                */
#define GOMELT_INTERP_RSUB 2   /* out[o] = base[o] - I       */

/* K2 / K4 / K5 - trilinear interpolation of a source-level field at the nodes of a tensor-product
 * target grid.  Replaces interpolatePoints cF:1131-1210, interpolatePointsMatrix cF:1028-1107 +
 * interpolate_w_matrix cF:1110-1128 (weights are recomputed, never stored), the face-only use in
 * assignBCsFine cF:1598-1620 (faces_only: targets on the y-,y+,x-,x+,z- faces), the time blend
 * alpha*new + beta*old of cF:3349/3389 (u2), the scatter into the parent's overlap nodes of
 * getNewTprime cF:2086-2090 (map_*), T' = T - I(parent) cF:2094-2097 (RSUB) and the
 * max(I(T0), T_amb) of the layer change gm:215-218 (clamp).  Points outside the source grid by
 * more than the reference's 1e-2 weight window contribute 0 (cF:1101-1102). */
typedef struct gomelt_interp_args {
    gomelt_axis_t src[3];
    const float  *u;              /* source field [src nn]                                   */
    const float  *u2;             /* NULL, or second source field: value = alpha*u + beta*u2 */
    float         alpha, beta;
    const float  *tx, *ty, *tz;   /* target coordinates                                      */
    int32_t       ntx, nty, ntz;
    int32_t       mode;           /* GOMELT_INTERP_*                                         */
    int32_t       faces_only;
    int32_t       has_clamp;
    float         clamp_min;      /* result = max(result, clamp_min) when has_clamp          */
    const int32_t *map_x, *map_y, *map_z; /* NULL, or index vectors: o = mx[i] + my[j]*map_nx + mz[k]*map_nx*map_ny */
    int32_t       map_nx, map_ny;
    const float  *base;           /* RSUB                                                    */
    float        *out;
} gomelt_interp_args_t;

int gomelt_interp_f32(const gomelt_interp_args_t *args, void *stream);

/* Face prolongation of a subcycle block in two parts (assignBCsFine cF:1598-1620 with the time blend of
 * cF:3386-3389): the block's N3 substeps prolong the same two parent fields with different blend factors, and
 * I(alpha u + beta u2) = alpha I(u) + beta I(u2).  gomelt_faces_gather_f32 interpolates u and u2 (both required)
 * at the nodes of the five Dirichlet faces of the target grid once, into compact arrays of
 * gomelt_faces_count(ntx, nty, ntz) floats each; gomelt_faces_blend_f32 writes
 * out[face node] = max(clamp_min, alpha*face_a + beta*face_b) (the clamp when has_clamp). */
long long gomelt_faces_count(int32_t ntx, int32_t nty, int32_t ntz);
int gomelt_faces_gather_f32(const gomelt_interp_args_t *args, float *face_a, float *face_b, void *stream);
int gomelt_faces_blend_f32(const float *face_a, const float *face_b, int32_t ntx, int32_t nty, int32_t ntz,
                           float alpha, float beta, int32_t has_clamp, float clamp_min, float *out, void *stream);

/* Window <-> big-grid copies through a tensor-product index set (getOverlapRegion cF:1642-1669):
 * scatter = 0: dst[t] = src[idx(t)]   (S1/S2 regather from Level 0, cF:2500-2502)
 * scatter = 1: dst[idx(t)] = src[t]   (Level-3 state back to Level 0, cF:2390-2392)
 * elem_size 4 (float) or 1 (uint8). */
int gomelt_box_copy(const void *src, void *dst, int32_t elem_size, const int32_t *ix, const int32_t *iy,
                    const int32_t *iz, int32_t nx, int32_t ny, int32_t nz, int32_t big_nx, int32_t big_ny,
                    int32_t scatter, void *stream);

/* F[n] (+)= coef * tx[ix]*ty[iy]*tz[iz]: a projected source term added to a parent load vector. */
int gomelt_rank1_f32(float *F, const float *tx, const float *ty, const float *tz, int32_t nx, int32_t ny,
                     int32_t nz, float coef, int32_t accumulate, void *stream);

/* K6 for the parents: computeSources cF:928-988 / computeLevelSource cF:2667-2730.  The laser source
 * integrated at the FINE level's Gauss points and projected with the PARENT's shape functions is
 * rank-1: Fc[n] = coef * wq_fine * tx[ix]*ty[iy]*tz[iz] (*coef returns 6 sqrt3 P eta). */
int gomelt_coarse_source_tables_f32(const gomelt_props_t *props, const gomelt_axis_t fine[3],
                                    const gomelt_axis_t parent[3], const float laser_xyz[3], float laserP,
                                    float *tx, float *ty, float *tz, float *coef, void *stream);

/* computeLevelSource cF:2667-2730 / computeSources cF:928-988 in two launches for ANY number of laser rows: the tables
 * of all rows (one launch), then F[node] (+)= sum_r c_r tx_r[ix] ty_r[iy] tz_r[iz] with c_r = 6 sqrt3 P_r eta wq_fine / n
 * (wq_fine = hx hy hz / 8 of the FINE level), rows summed in order.  rows = HOST [n][7] toolpath rows (P in column 6); tables = device scratch
 * [n * (parent nx + ny + nz)]; n <= GOMELT_MAX_SUBSTEPS. */
int gomelt_projected_source_f32(const gomelt_props_t *props, const gomelt_axis_t fine[3], const gomelt_axis_t parent[3],
                                float wq_fine, const float *rows, int32_t n, float *tables, float *F, int32_t accumulate,
                                void *stream);

/* K5 - window shift of moveEverything cF:2400-2510 for one window level in ONE launch: at the nodes of the window's new
 * position (tx, ty, tz)   Tp_new = I_old(Tp_old),   T_new = I_L1(T1) + (I_mid(Tp_mid) + Tp_new)   (Level 3: mid = the
 * Level-2 window at its old position; Level 2: Tp_mid = NULL and T_new = I_L1(T1) + Tp_new, cF:2439-2443, 2460-2464).
 * Interpolants as in gomelt_interp_f32. */
typedef struct gomelt_shift_args {
    gomelt_axis_t L1[3];  const float *T1;
    gomelt_axis_t mid[3]; const float *Tp_mid;
    gomelt_axis_t old[3]; const float *Tp_old;
    const float  *tx, *ty, *tz;
    int32_t       ntx, nty, ntz;
    float        *Tp_new, *T_new;   /* must not alias the inputs */
} gomelt_shift_args_t;
int gomelt_shift_window_f32(const gomelt_shift_args_t *args, void *stream);

/* x[i] = max(x[i], lo) in place (the jnp.maximum(T_amb, .) of stepGOMELT cF:2360-2362 where it has to follow the face
 * prolongation from the unclamped parent). */
int gomelt_clamp_min_f32(float *x, int64_t n, float lo, void *stream);

/* K3 - fine -> parent correction vectors, integrated at the fine Gauss points:
 *   mode 0: V[c] (+)= - sum wq * grad Nc . (kbar grad A)          computeCoarseTprimeTerm_jax cF:1477-1565,
 *                                                                  computeL1/L2TprimeTerms_Part1 cF:2733-2914
 *   mode 1: V[c] (+)= - scale * sum wq * Nc * (rcbar * A)          computeCoarseTprimeMassTerm_jax cF:1396-1474,
 *                                                                  ..._Part2 cF:3057-3221 (A = A - A2, scale = 1/dt)
 * The fine elements are grouped by parent cell: cell0 / ncell = box of parent cells that contain fine
 * elements, first_d[i] = first fine element (axis d) of parent cell cell0[d] + i (length ncell[d] + 1).
 * cellsum is scratch [ncell_x*ncell_y*ncell_z*8].  Deterministic (no float atomics). */
typedef struct gomelt_project_args {
    gomelt_axis_t fine[3], parent[3];
    const float  *A, *A2, *coef;
    int32_t       mode;
    float         scale;
    int32_t       cell0[3], ncell[3];
    const int32_t *first_x, *first_y, *first_z;
    int32_t       elems_per_cell_hint;
    float        *cellsum;
    float        *V;
    int32_t       accumulate;
    /* coef == NULL: the nodal coefficient is evaluated inside the kernel from the fine level's state with
     * computeStateProperties cF:2567-2614 (mode 0: k, mode 1: rho*cp) - no k / rho*cp array is materialised */
    const float  *coef_T, *coef_S1;
    int64_t       coef_n_substrate;
    const gomelt_props_t *coef_props;
    /* Fast (tiled) form, optional: per fine element e and axis, the parent's shape-function factors at its two Gauss
     * points, wtab_d[4 e .. 4 e + 3] = (x1 - xq0, xq0 - x0, x1 - xq1, xq1 - x0) with [x0, x1] the parent cell that holds
     * each Gauss point (device arrays, 16-byte aligned, built once per window position); rmax = the largest number of
     * fine elements per parent cell along each axis; hf / hc = element sizes of the fine / parent level.  All NULL / 0:
     * the general kernel derives everything from the coordinate arrays.  With the tables, windows that nest in x and y
     * (every parent cell of the box holds exactly rmax[0] x rmax[1] element columns; any grouping in z) take the marching
     * kernel (no staging, plane-shared sum factorisation: csrc/k_transfer.cu project_march_kernel), the others the
     * shared-memory tile kernel; elems_per_cell_hint < 0 (magnitude = the hint) asks for the tile kernel (A/B). */
    const float  *wtab_x, *wtab_y, *wtab_z;
    int32_t       rmax[3];
    float         hf[3], hc[3];
    /* Nested grouping along x and y, stated by the caller (the first_d tables live on the device): parent cell i of the
     * box holds the fine elements [i * rmax[d] - uniform_off[d], (i + 1) * rmax[d] - uniform_off[d]) clipped to the
     * fine grid, i.e. uniform_off[d] (0 <= off < rmax[d]) elements are missing from the first cell - a window that
     * starts inside a parent cell, as the Level-3 window does relative to Level 1.  0: the library itself checks
     * ncell[d] * rmax[d] == number of fine elements; -1: not nested (tile / general kernel).  [2] is ignored (z follows
     * first_z). */
    int32_t       uniform_off[3];
} gomelt_project_args_t;

int gomelt_project_f32(const gomelt_project_args_t *args, void *stream);

/* ---- the Level-3 inner scan as ONE call ------------------------------------------------------------------
 * subcycleL3_Part1 / subcycleL3_Part2 of subcycleGOMELT (the inner jax.lax.scan, cF:3367-3412 / 3530-3590): n
 * substeps of   computeSourcesL3 cF:2960 -> computeConvRadBC cF:2207 -> computeSolutions_L3 cF:3015 (explicit
 * solve, Dirichlet faces <- alpha*parent_new + beta*parent_old with alpha = (i+1)/faces_n (cF:3386-3389),
 * max(T_amb, .)), state update, and - with GOMELT_STEP_ACCUM - the melt-time bookkeeping cF:3568-3578.
 * Launches: 1 (all source tables) + n (fused level steps) + n (face prolongation, when `faces` is given; see
 * faces_scratch).
 * Substep i writes W_i = T_a (i even) / T_b (i odd) and reads W_(i-1), substep 0 reads T_in, which is never
 * written unless it is T_b (so a caller can keep its initial field, or ping-pong by passing T_in = T_b).  The
 * buffer holding the newest field is returned in *T_last.  S1_in likewise is read by substep 0 only (NULL = S1).
 * flags: GOMELT_STEP_CLAMP | GOMELT_STEP_SKIP_FACES | GOMELT_STEP_WRITE_S2 | GOMELT_STEP_ACCUM are honoured;
 * WRITE_S1 (in place) and FUSED_FLUX are always on. */
typedef struct gomelt_substeps_args {
    gomelt_grid_t grid;
    const float  *x, *y, *z;      /* device node-coordinate arrays of the level                               */
    int32_t       n;              /* substeps, 1..GOMELT_MAX_SUBSTEPS                                         */
    const float  *rows;           /* HOST [n][7] toolpath rows: x, y, z, Ljump, Ldwell, dt, P                 */
    const float  *T_in;           /* [nn] temperature at the start of the block                               */
    float        *T_a, *T_b;      /* ping-pong temperature buffers [nn]; T_a != T_in                          */
    const float  *S1_in;          /* [nn] state at the start of the block, or NULL = S1                       */
    float        *S1;             /* [nn] written by every substep, read from substep 1 on                    */
    int64_t       n_substrate;
    int32_t       flags;
    float        *tables;         /* device scratch [n * (nx + ny + nz)]                                      */
    uint8_t      *S2;             /* [nn] in place (WRITE_S2 / ACCUM) or NULL                                 */
    float        *accum, *max_accum; /* [nn] in place (ACCUM) or NULL                                         */
    const gomelt_interp_args_t *faces; /* NULL, or the face prolongation from the parent: u = parent_new, u2 =
                                     parent_old, tx/ty/tz = this level's coordinates; out / alpha / beta /
                                     faces_only are set per substep by the call                                */
    float         faces_n;        /* fN3 (float, as the reference divides: alpha = (i+1)/fN3, beta = 1-alpha) */
    float       **T_last;         /* HOST out: buffer holding the newest temperature (may be NULL)            */
    uint32_t     *bk_queue;       /* NULL, or the scratch of gomelt_step_args_t.bk_queue (used by the ACCUM substeps)   */
    int64_t       bk_queue_words;
    void * const *step_events;    /* NULL, or HOST array of 2 n cudaEvent_t: recorded on `stream` right before / after the
                                     fused level step of every substep (measurement: the kernel's duration as it runs
                                     inside the block, without a caller's issue path between the records)             */
    float        *faces_scratch;  /* NULL, or device scratch [2 * gomelt_faces_count(nx, ny, nz)]: the two parent
                                     fields are interpolated at the face nodes ONCE per call (gomelt_faces_gather_f32)
                                     and every substep blends them (gomelt_faces_blend_f32) instead of
                                     re-interpolating; launches 1 + 1 + 2n; agrees with the NULL path to rounding */
} gomelt_substeps_args_t;

int gomelt_l3_substeps_f32(const gomelt_props_t *props, const gomelt_substeps_args_t *args, void *stream);

/* ---- the reference's jitted steppers as single native calls --------------------------------------------------------
 * stepGOMELT cF:2304-2397, subcycleGOMELT cF:3224-3632, stepGOMELTDwellTime cF:2617-2664 and the device part of
 * moveEverything cF:2400-2510: the whole kernel sequence of one call is issued from C++ on the caller's stream with no
 * host synchronisation and no allocation (the reference runs each as one jax.jit; these are the XLA-FFI targets of
 * INTEGRATION.md).  State is updated IN PLACE in the caller's buffers; intermediates live in `work`. */
typedef struct gomelt_level {
    gomelt_grid_t grid;
    const float  *x, *y, *z;     /* device node-coordinate arrays                                   */
    float        *T0;            /* [nn] in / out                                                   */
    float        *S1;            /* [nn] in / out                                                   */
    float        *Tprime0;       /* [nn] in / out (Levels 2 and 3; NULL on Level 1)                 */
    uint8_t      *S2;            /* [nn] in / out (Level 3; NULL otherwise)                         */
    int64_t       n_substrate;   /* getSubstrateNodes cF:562-579                                    */
} gomelt_level_t;

typedef struct gomelt_pair {     /* fine -> parent element grouping of one level pair (the role of Shapes[.])  */
    int32_t       cell0[3], ncell[3];
    const int32_t *first_x, *first_y, *first_z;
    int32_t       elems_per_cell_hint;
    const float  *wtab_x, *wtab_y, *wtab_z;   /* see gomelt_project_args_t (may be NULL) */
    int32_t       rmax[3];
    int32_t       uniform_off[3];             /* see gomelt_project_args_t */
} gomelt_pair_t;

typedef struct gomelt_overlap {  /* parent nodes under a window: index vectors and their coordinates            */
    const int32_t *ix, *iy, *iz;
    const float  *cx, *cy, *cz;
    int32_t       n[3];
} gomelt_overlap_t;

/* see gomelt_hier_t.l1_solve: flags = the GOMELT_STEP_* bits of the Level-1 solve it replaces, rhs may be NULL.
 * GOMELT_L1_CLAMP_AFTER: the stepper prolongs the UNCLAMPED field onto the child faces and applies max(T_amb, .)
 * afterwards (computeSolutions cF:2184-2196, then cF:2360-2362): the slabs do the same once the box has been shipped. */
#define GOMELT_L1_CLAMP_AFTER 0x10000
typedef int (*gomelt_l1_solve_fn)(void *user, const float *T0, const float *S1, float *T_out, float dt, const float *rhs,
                                  int32_t flags);

typedef struct gomelt_hier {
    gomelt_level_t   L1, L2, L3;
    gomelt_pair_t    L2L1, L3L1, L3L2;
    gomelt_overlap_t ov2, ov3;             /* Level 2 in Level 1, Level 3 in Level 2 (overlapNodes / overlapCoords)   */
    float           *L0_S1;                /* Level-0 state grid (cF:206-246)                                          */
    uint8_t         *L0_S2;
    int32_t          L0_nx, L0_ny, L0_nz;
    const int32_t   *l0_ix, *l0_iy, *l0_iz; /* the Level-3 window's nodes in Level 0 (lengths = Level-3 nx, ny, nz)    */
    const int32_t   *l0p_ix, *l0p_iy, *l0p_iz; /* NULL, or the index set the PREVIOUS stepper call scattered Level-3 S2 to:
                                              the only Level-0 S2 entries that can be non-zero (cF:2391 zeroes the grid
                                              before every scatter) - cleared instead of the whole state grid           */
    int32_t          l0p_n[3];
    float            bc5[5];               /* Level-1 Dirichlet values y-, y+, x-, x+, z-                              */
    int32_t          nz_active_L1;         /* active Level-1 planes (tmp_ne_nn, cF:495-517)                            */
    float           *L1_spare;             /* [nn1] second Level-1 temperature buffer: the new Level-1 field is left
                                              THERE and *l1_in_spare is set to 1 (the caller swaps its handles: no
                                              copy of the part-scale field); NULL = copy back into L1.T0             */
    float           *work;                 /* device scratch                                                           */
    int64_t          work_floats;          /* >= gomelt_hier_work_floats(...)                                          */
    /* Slab-decomposed Level 1 (gomelt_b200/dist.py): when set, every Level-1 solve of the steppers
     * (computeSolutions cF:2172-2185, computeL1Temperature cF:2813-2854, stepGOMELTDwellTime cF:2617-2664) is handed to
     * this function instead of gomelt_level_step_f32 on the arrays of L1.  On the rank that owns the laser, L1.T0 / L1.S1
     * / the load vector are full-size MIRRORS that are valid on the Level-1 box under the Level-2 window only; the hook
     * ships that box of S1 and of the load to the slab owners (gomelt_patch_copy + the transport of the host side),
     * all ranks sweep their slabs, and the box of the new temperature comes back into T_out.  It runs on the calling
     * host thread, between kernels issued on `stream`; a non-zero return aborts the stepper with that code. */
    gomelt_l1_solve_fn l1_solve;
    void            *l1_user;
} gomelt_hier_t;

long long gomelt_hier_work_floats(const gomelt_hier_t *h, int32_t N2, int32_t N3);

/* rows = HOST [N2*N3][7] toolpath rows with the laser power in column 6; accum / max_accum = device [nn3] melt-time
 * windows, in / out (cF:3568-3578). */
int gomelt_subcycle_f32(const gomelt_props_t *props, const gomelt_hier_t *h, const float *rows, int32_t N2, int32_t N3,
                        float *max_accum, float *accum, int32_t *l1_in_spare, void *stream);
/* row = HOST [7]; resetmask = device [nn3] out: nodes that have just melted ((1 - 2 preS2) * S2 == 1, cF:2394). */
int gomelt_step_f32(const gomelt_props_t *props, const gomelt_hier_t *h, const float *row, uint8_t *resetmask,
                    int32_t *l1_in_spare, void *stream);
int gomelt_dwell_step_f32(const gomelt_props_t *props, const gomelt_hier_t *h, float dt, int32_t *l1_in_spare, void *stream);

/* Melt-time bookkeeping of a single-step row (gm:339-357 + melting_temp cF:3696-3712) in one launch, on the Level-0
 * arrays through the Level-3 window's index vectors: reset = accum * resetmask; max_accum = max(reset, max_accum);
 * accum += -reset; accum += (T3 > T_liquidus) * dt. */
int gomelt_accum_single_step_f32(const float *T3, const uint8_t *resetmask, float dt, float T_liquidus, float *accum0,
                                 float *max_accum0, const int32_t *ix, const int32_t *iy, const int32_t *iz, int32_t nx,
                                 int32_t ny, int32_t nz, int32_t big_nx, int32_t big_ny, void *stream);

/* Patch exchange, device side (SURVEY.md 8b gomelt_patch_exchange; 8e "fine <-> coarse across ranks"): copy the box
 * [lo, lo + n) of a 3-D x-fastest float32 array into the box of the same size at dlo of another one.  Either side may be
 * a contiguous staging buffer (dims = n, lo = 0: pack / unpack around a send / recv) or a peer-mapped array of another
 * rank (symmetric memory: the copy then IS the exchange, over NVLink).  Used for the Level-1 box under the Level-2
 * window: T old / new down and up, injected T, S1 and the load vector (getNewTprime cF:2060-2099,
 * computeCoarseTprimeTerm_jax cF:1477-1565, updateStateProperties cF:2546-2556). */
int gomelt_patch_copy_f32(const float *src, const int32_t sdims[3], const int32_t slo[3], float *dst, const int32_t ddims[3],
                          const int32_t dlo[3], const int32_t n[3], void *stream);

/* Monitor (printLevelMaxMin cF:3635-3665): out3 = {min, max over the finite values, number of non-finite values}
 * of x[0..n) in one reduction (device memory, 3 floats; read it back when convenient). */
int gomelt_minmax_f32(const float *x, int64_t n, float *out3, void *stream);

const char *gomelt_last_error(void);
int gomelt_abi_version(void);
/* Kernels launched by this library in this process so far (every <<< >>> site counts itself). */
long long gomelt_launch_count(void);

/* 1 when the library was built with jaxlib's xla/ffi/api/ffi.h and exports the XLA FFI handler symbols
 * (GomeltLevelStepFfi, ...: csrc/xla_ffi_shim.cc), 0 otherwise (this image: no jaxlib). */
int gomelt_xla_ffi_available(void);

/* Diagnostics: FP32 issue-rate micro-benchmark.  kind: 0 = FFMA (3-reg),
 * 1 = FADD, 2 = packed FFMA2 (fma.rn.f32x2).  Returns lane-ops executed in *ops. */
int gomelt_diag_fp32_rate(int32_t kind, int32_t iters, int32_t blocks, int32_t threads, float *sink,
                          double *ops, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* GOMELT_ABI_H */
