// Device-side views of the shared (per-pattern) index data and of the per-instance workspace.
//
// HBM layout ("tiled structure of arrays", batch innermost):
//   an instance belongs to tile t = b / TILE and lane = b % TILE.  Every per-instance array is a run
//   of ROWS; row r of tile t is the TILE consecutive doubles  ws[(t*rows_total + r)*TILE + lane].
//   A warp that walks rows therefore issues one fully coalesced 256-byte access per row whatever the
//   row index is, and all rows of one tile are contiguous (column slices of L are contiguous runs).
//   Index data (patterns, maps, schedules) is shared by the whole batch and read with warp-uniform
//   loads.
#pragma once

#include <cstddef>

namespace eicos
{

// per-instance status while the interior-point loop is running
enum : int
{
    ST_ACTIVE = -1000,
    EXIT_OPTIMAL = 0,
    EXIT_PINF = 1,
    EXIT_DINF = 2,
    EXIT_MAXIT = -1,
    EXIT_NUMERICS = -2,
    EXIT_OUTCONE = -3,
    EXIT_FATAL = -7,
    EXIT_INACC = 10,
    EXIT_NOT_CONVERGED = -87
}; // include/eicos.hpp:8-21 of the reference

// scalar rows (doubles), one row each
enum ScalarRow : int
{
    S_KAP, S_TAU, S_CX, S_BY, S_HZ, // Work scalars (include/eicos.hpp:107-111)
    S_PCOST, S_DCOST, S_PRES, S_DRES, S_PINFRES, S_DINFRES, S_GAP, S_RELGAP,
    S_SIGMA, S_MU, S_STEP, S_STEP_AFF, S_KAPOVERT, // Information doubles (:51-65)
    S_WORK_END,
    S_BEST = S_WORK_END,                 // same 18 rows again for w_best
    S_BEST_END = S_BEST + S_WORK_END,
    S_RT = S_BEST_END, S_NX, S_NY, S_NZ, S_NS, S_HRESX, S_HRESY, S_HRESZ,
    S_RESX0, S_RESY0, S_RESZ0, S_PRES_PREV,
    S_DTAU_DENOM, S_DTAUAFF, S_DKAPAFF,
    S_COUNT
};

// integer rows
enum IntRow : int
{
    J_ITER, J_PINF, J_DINF, J_HAS_PINFRES, J_HAS_DINFRES, J_HAS_RELGAP, J_NIT1, J_NIT2, J_NIT3, // Information ints
    J_WORK_END,
    J_BEST = J_WORK_END,
    J_BEST_END = J_BEST + J_WORK_END,
    J_STATUS = J_BEST_END, // ST_ACTIVE or the exit code
    J_SCALEFAIL,           // first cone whose scaling update failed this iteration (or nc)
    J_COUNT
};

// Row offsets of every per-instance array inside a tile's workspace block.
struct Layout
{
    int c, h, b;                    // equilibrated problem vectors
    int x, y, z, s, lam;            // current iterate (Work)
    int bx, by, bz, bs, blam;       // best iterate (w_best)
    int rx, ry, rz;                 // residuals
    int lpv, lpw;                   // LP cone scalings
    int cpar, cq;                   // SOC scalings: 8 rows per cone (CP_*), q vectors
    int V;                          // scaling block values of the KKT matrix (cacheIndices order)
    int Lx, LTx, D, Dinv;           // factor: columns, rows (copy), pivots, reciprocals
    int rhs1, rhs2, sol1, sol2;     // KKT-space vectors (length N)
    int xw, dxr, e;                 // triangular-solve work vector, refinement step, residual
    int dsw, wdz, dsaff, ds1;       // dsaff_by_W, W_times_dzaff, dsaff, scratch (length m)
    int sc;                         // S_COUNT scalar rows
    int rows_total;
    int irows_total;                // integer rows (J_COUNT)
};

enum ConeParam : int
{
    CP_ETA, CP_ETA2, CP_A, CP_D1, CP_U0, CP_U1, CP_V1, CP_W, CP_COUNT
};

struct PhaseDev
{
    int begin, end, parallel;
};

// Shared index data on the device (all pointers are device pointers).
struct DevPattern
{
    int n, p, m, l, nc, N, mt, qtot, nnzL, nnzV, nphases, maxcol;
    const int *cone_dim, *cone_z, *cone_k, *cone_q, *zk;
    const int *Gp, *Gi, *Grp, *Grj, *Grv;
    const int *Ap, *Ai, *Arp, *Arj, *Arv;
    const double *Gx, *Ax, *xeq, *Aeq, *Geq;
    const int *pinv, *Lp, *Li, *Lio, *Lcsr, *Lrp, *Lrj;
    const int *KLp, *KLvidx, *KLpos;
    const double *KLval;
    const int *upd_tail, *upd_rel_p, *upd_rel;
    const int *tasks;
    const PhaseDev *phases;
    const int *Vkind; // per V entry: what resetKKTScalings writes (0 -> -1, 1 -> 0, 2 -> +1)
};

} // namespace eicos
