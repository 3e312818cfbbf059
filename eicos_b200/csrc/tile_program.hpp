// The per-tile interior-point program: every step of EiCOS's Solver::solve (reference
// src/eicos.cpp:848-1262) for TILE instances at a time, lane = instance.
//
// One CTA owns one tile.  Its warps ("workers") split rows / cones / elimination-tree tasks among
// themselves and meet at CTA barriers; per-instance reductions over a vector run down the rows in
// each worker and are combined through shared memory in a fixed order, so every warp of the CTA
// holds bit-identical per-lane scalars and all control flow is uniform across the CTA.
// Sparse structure is never looked up through CSR/CSC arrays on the device: each worker decodes
// its own instruction stream (streams.hpp) with coalesced chunk loads + warp shuffles.
//
// The same source compiles two ways:
//   * nvcc (default): TILE = 32, workers = warps of the CTA  -> the product.
//   * -DEICOS_EMU (tests/emu only): TILE = 1, one worker, plain C++ -> lets the CPU-only test
//     tier execute the kernel logic against the oracle.  It is never linked into the product.
#pragma once

#include "layout.hpp"
#include "streams.hpp"
#include "symbolic.hpp"

#include <cfloat>
#include <cmath>

#ifdef EICOS_EMU
#include <barrier>
#define EI_DEV inline
#define EI_LDG(p) (*(p))
#define EI_PREFETCH(p) ((void)0)
#else
#include <cuda_runtime.h>
#define EI_DEV __device__ __forceinline__
#define EI_LDG(p) __ldg(p)
#define EI_PREFETCH(p) asm volatile("prefetch.global.L2 [%0];" ::"l"(p))
#endif

namespace eicos
{

#ifdef EICOS_EMU
constexpr int TILE = 1;
#else
constexpr int TILE = 32;
#endif
constexpr int KRED = 17;    // widest block reduction
constexpr int PF_ROWS = 24; // how many value rows ahead the sweeps prefetch into L2

struct KArgs
{
    DevPattern P;
    Layout L;
    double *ws; // [tiles][rows_total][TILE]
    int *iws;   // [tiles][irows_total][TILE]
    double *acc_global; // factor accumulators when they do not fit shared memory, else null
    int batch;  // instances handled by this launch's chunk
    int first;  // global index (within the device's batch) of the chunk's first instance
    // instance-major device buffers for load/store (any may be null)
    const double *in_c, *in_h, *in_b;
    const double *base_c, *base_h, *base_b; // raw vectors shared by the batch (used when in_* is null)
    double *out_x, *out_y, *out_z, *out_s;
    int *out_exit, *out_iter;
    double *out_info; // [batch][S_WORK_END]
    int *out_iinfo;   // [batch][J_WORK_END]
    // solveKKT parameters
    int rhs, sol, initialize, nitrow;
    int keep_sticky;
    int pre_equilibrated; // inputs are already divided by the equilibration vectors
    unsigned int *active_count; // device counter: instances still iterating after the head step
    unsigned long long *ir_rounds; // device counter: solve rounds executed (tile-rounds)
};

struct Team
{
    int lane, wk, nwk;
    double *red;   // [nwk][KRED][TILE]
    double *acc;   // [nwk][maxcol][TILE]   (factor kernel only)
    double *stage; // this worker's staging slots + lane: slot s lives at stage[s * TILE]
#ifdef EICOS_EMU
    std::barrier<> *bar; // workers of a tile are real threads in the emulator
    void sync() const
    {
        if (bar)
            bar->arrive_and_wait();
    }
    bool all(bool v) const { return v; }
    bool any(bool v) const { return v; }
#else
    __device__ __forceinline__ void sync() const { __syncthreads(); }
    __device__ __forceinline__ bool all(bool v) const { return __all_sync(0xffffffffu, v); }
    __device__ __forceinline__ bool any(bool v) const { return __any_sync(0xffffffffu, v); }
#endif
};

// ------------------------------------------------------------------ instruction stream readers
// All lanes of the warp call get() together (never under a per-lane branch).
struct IStream
{
#ifdef EICOS_EMU
    const int *p;
    EI_DEV void open(const int *base, int) { p = base; }
    EI_DEV int get() { return *p++; }
#else
    const int *p;
    int cur, nxt, pos, lane;
    EI_DEV void open(const int *base, int lane_)
    {
        lane = lane_;
        cur = __ldg(base + lane);
        nxt = __ldg(base + STREAM_CHUNK + lane);
        p = base + 2 * STREAM_CHUNK;
        pos = 0;
    }
    EI_DEV int get()
    {
        if (pos == STREAM_CHUNK)
        {
            cur = nxt;
            nxt = __ldg(p + lane);
            p += STREAM_CHUNK;
            pos = 0;
        }
        const int v = __shfl_sync(0xffffffffu, cur, pos);
        ++pos;
        return v;
    }
#endif
};

struct DStream
{
#ifdef EICOS_EMU
    const double *p;
    EI_DEV void open(const double *base, int) { p = base; }
    EI_DEV double get() { return *p++; }
#else
    const double *p;
    double cur, nxt;
    int pos, lane;
    EI_DEV void open(const double *base, int lane_)
    {
        lane = lane_;
        cur = __ldg(base + lane);
        nxt = __ldg(base + STREAM_CHUNK + lane);
        p = base + 2 * STREAM_CHUNK;
        pos = 0;
    }
    EI_DEV double get()
    {
        if (pos == STREAM_CHUNK)
        {
            cur = nxt;
            nxt = __ldg(p + lane);
            p += STREAM_CHUNK;
            pos = 0;
        }
        const double v = __shfl_sync(0xffffffffu, cur, pos);
        ++pos;
        return v;
    }
#endif
};

// ------------------------------------------------------------------ small helpers
#define ROWD(base, r) (base)[(size_t)(r) * TILE + tm.lane]

EI_DEV bool ei_isnan(double v) { return v != v; }
EI_DEV double dmax(double a, double b) { return a > b ? a : b; } // std::max(a,b) semantics (returns a on ties / NaN in b)
EI_DEV double dmin(double a, double b) { return b < a ? b : a; } // std::min(a,b)

template <int K>
EI_DEV void team_sum(const Team &tm, double (&v)[K])
{
    if (tm.nwk == 1)
        return;
    tm.sync();
    for (int i = 0; i < K; i++)
        tm.red[(size_t)(tm.wk * K + i) * TILE + tm.lane] = v[i];
    tm.sync();
    for (int i = 0; i < K; i++)
    {
        double s = 0.0;
        for (int w = 0; w < tm.nwk; w++)
            s += tm.red[(size_t)(w * K + i) * TILE + tm.lane];
        v[i] = s;
    }
}

template <int K>
EI_DEV void team_max(const Team &tm, double (&v)[K])
{
    if (tm.nwk == 1)
        return;
    tm.sync();
    for (int i = 0; i < K; i++)
        tm.red[(size_t)(tm.wk * K + i) * TILE + tm.lane] = v[i];
    tm.sync();
    for (int i = 0; i < K; i++)
    {
        double s = tm.red[(size_t)i * TILE + tm.lane];
        for (int w = 1; w < tm.nwk; w++)
            s = dmax(s, tm.red[(size_t)(w * K + i) * TILE + tm.lane]);
        v[i] = s;
    }
}

template <int K>
EI_DEV void team_min(const Team &tm, double (&v)[K])
{
    if (tm.nwk == 1)
        return;
    tm.sync();
    for (int i = 0; i < K; i++)
        tm.red[(size_t)(tm.wk * K + i) * TILE + tm.lane] = v[i];
    tm.sync();
    for (int i = 0; i < K; i++)
    {
        double s = tm.red[(size_t)i * TILE + tm.lane];
        for (int w = 1; w < tm.nwk; w++)
            s = dmin(s, tm.red[(size_t)(w * K + i) * TILE + tm.lane]);
        v[i] = s;
    }
}

struct TileMem
{
    double *T;
    int *I;
};

EI_DEV TileMem tile_mem(const KArgs &a, int tile)
{
    TileMem t;
    t.T = a.ws + (size_t)tile * a.L.rows_total * TILE;
    t.I = a.iws + (size_t)tile * a.L.irows_total * TILE;
    return t;
}

EI_DEV bool lane_active(const Team &tm, const TileMem &t) { return ROWD(t.I, J_STATUS) == ST_ACTIVE; }

// ------------------------------------------------------------------ asynchronous staging
// A warp issues in order, so a load that is consumed right away limits it to ~2 loads in flight.
// Independent work is therefore done block-wise: pass 1 issues every load of the block into this
// worker's shared-memory slots with cp.async (no register dependency, up to STAGE_SLOTS rows =
// 8 KB in flight per warp), pass 2 re-walks the same stream words and computes from shared memory.
// Each lane only ever reads the slot words it wrote itself, so cp.async.wait_all is the only fence.
EI_DEV void stage_issue(double *slot_lane, const double *src_lane)
{
#ifdef EICOS_EMU
    *slot_lane = *src_lane;
#else
    const unsigned sa = (unsigned)__cvta_generic_to_shared(slot_lane);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(src_lane) : "memory");
#endif
}
EI_DEV void stage_wait()
{
#ifndef EICOS_EMU
    asm volatile("cp.async.wait_all;" ::: "memory");
#endif
}

EI_DEV const double *rowp(const Team &tm, const double *T, int row) { return T + (size_t)row * TILE + tm.lane; }

// sum_k val_k * vec[idx_k] folded into `v` with sign: one mat-vec row straight from global memory
EI_DEV double row_accumulate(const Team &tm, IStream &is, DStream &ds, const double *T, int vec, double v, double sign)
{
    const int cnt = is.get();
    for (int k = 0; k < cnt; k++)
    {
        const int idx = is.get();
        const double val = ds.get();
        v += (sign * val) * ROWD(T, vec + idx);
    }
    return v;
}

// Walks one worker's share of a mat-vec row set (rows first, first+nwk, ...) block by block.
//   extra(row, k)  -> row offset of the k-th extra operand of `row` (k < NEX), staged with the gathers
//   finish(row, ex, v) receives the extras and v = init(ex) + sum sign*val*vec[idx]
template <int NEX, class Extra, class Init, class Finish>
EI_DEV void rowset_run(const Team &tm, const double *T, const int *stream, const double *vals, const int *seg,
                       int first, int vec, double sign, Extra extra, Init init, Finish finish)
{
    IStream is;
    DStream ds;
    is.open(stream + EI_LDG(seg + tm.wk * 3), tm.lane);
    ds.open(vals + EI_LDG(seg + tm.wk * 3 + 1), tm.lane);
    const int nblocks = EI_LDG(seg + tm.wk * 3 + 2);
    int row = first + tm.wk;
    for (int b = 0; b < nblocks; b++)
    {
        const int nr = is.get();
        if (nr < 0)
        { // oversize row: no staging
            double ex[NEX > 0 ? NEX : 1];
            for (int k = 0; k < NEX; k++)
                ex[k] = *rowp(tm, T, extra(row, k));
            const double v = row_accumulate(tm, is, ds, T, vec, init(ex), sign);
            finish(row, ex, v);
            row += tm.nwk;
            continue;
        }
        const IStream mark = is;
        double *sp = tm.stage;
        int rr = row;
        for (int t = 0; t < nr; t++, rr += tm.nwk)
        {
            for (int k = 0; k < NEX; k++, sp += TILE)
                stage_issue(sp, rowp(tm, T, extra(rr, k)));
            const int cnt = is.get();
            for (int k = 0; k < cnt; k++, sp += TILE)
                stage_issue(sp, rowp(tm, T, vec + is.get()));
        }
        stage_wait();
        is = mark;
        sp = tm.stage;
        for (int t = 0; t < nr; t++, row += tm.nwk)
        {
            double ex[NEX > 0 ? NEX : 1];
            for (int k = 0; k < NEX; k++, sp += TILE)
                ex[k] = *sp;
            double v = init(ex);
            const int cnt = is.get();
            for (int k = 0; k < cnt; k++, sp += TILE)
            {
                (void)is.get();
                v += (sign * ds.get()) * *sp;
            }
            finish(row, ex, v);
        }
    }
}

// ------------------------------------------------------------------ W products (src/eicos.cpp:485-507)
// out = W * in for the lanes' current scalings; in/out are z-shaped (expanded) row offsets.
EI_DEV void cone_scale(const Team &tm, const KArgs &a, double *T, int in, int out, bool write)
{
    const DevPattern &P = a.P;
    const Layout &L = a.L;
    for (int k = tm.wk; k < P.l; k += tm.nwk)
    {
        const double v = ROWD(T, L.lpw + k) * ROWD(T, in + k);
        if (write)
            ROWD(T, out + k) = v;
    }
    for (int c = tm.wk; c < P.nc; c += tm.nwk)
    {
        const int d = EI_LDG(P.cone_dim + c), ks = EI_LDG(P.cone_k + c), qo = EI_LDG(P.cone_q + c);
        const double *cp = T + (size_t)(L.cpar + c * CP_COUNT) * TILE + tm.lane;
        const double eta = cp[CP_ETA * TILE], ca = cp[CP_A * TILE];
        double zeta = 0.0;
        for (int k = 1; k < d; k++)
            zeta += ROWD(T, L.cq + qo + k - 1) * ROWD(T, in + ks + k);
        const double z0 = ROWD(T, in + ks);
        const double factor = z0 + zeta / (1. + ca);
        if (write)
        {
            ROWD(T, out + ks) = eta * (ca * z0 + zeta);
            for (int k = 1; k < d; k++)
                ROWD(T, out + ks + k) = eta * (ROWD(T, in + ks + k) + factor * ROWD(T, L.cq + qo + k - 1));
        }
    }
}

// ------------------------------------------------------------------ line search (src/eicos.cpp:1380-1469)
// lambda, ds, dz are z-shaped row offsets.  Returns the clamped step for every lane.
// TODO(parity): the reference's `continue` on lknorm2<=0 skips the cone offset advance; here later
// cones keep their own offsets (differs only after lambda has already left the cone).
EI_DEV double line_search(const Team &tm, const KArgs &a, double *T, int lam, int ds, int dz,
                          double tau, double dtau, double kap, double dkap)
{
    const DevPattern &P = a.P;
    double alpha;
    double mn[3] = {DBL_MAX, DBL_MAX, DBL_MAX}; // rhomin, sigmamin, min over cones of 1/conic_step
    for (int k = tm.wk; k < P.l; k += tm.nwk)
    {
        const double lk = ROWD(T, lam + k);
        mn[0] = dmin(mn[0], ROWD(T, ds + k) / lk);
        mn[1] = dmin(mn[1], ROWD(T, dz + k) / lk);
    }
    for (int c = tm.wk; c < P.nc; c += tm.nwk)
    {
        const int d = EI_LDG(P.cone_dim + c), zs = EI_LDG(P.cone_k + c);
        const double l0 = ROWD(T, lam + zs);
        double sq = 0.0;
        for (int k = 1; k < d; k++)
        {
            const double v = ROWD(T, lam + zs + k);
            sq += v * v;
        }
        const double lknorm2 = l0 * l0 - sq;
        if (lknorm2 <= 0.)
            continue;
        const double lknorm = sqrt(lknorm2);
        const double lknorminv = 1. / lknorm;
        const double lk0 = l0 / lknorm;
        double dsdot = 0.0, dzdot = 0.0;
        for (int k = 1; k < d; k++)
        {
            const double lkb = ROWD(T, lam + zs + k) / lknorm;
            dsdot += lkb * ROWD(T, ds + zs + k);
            dzdot += lkb * ROWD(T, dz + zs + k);
        }
        const double ds0 = ROWD(T, ds + zs), dz0 = ROWD(T, dz + zs);
        const double lds = lk0 * ds0 - dsdot, ldz = lk0 * dz0 - dzdot;
        const double rho0 = lknorminv * lds, sig0 = lknorminv * ldz;
        const double frho = (lds + ds0) / (lk0 + 1.), fsig = (ldz + dz0) / (lk0 + 1.);
        double ar = 0.0, as = 0.0;
        for (int k = 1; k < d; k++)
        {
            const double lkb = ROWD(T, lam + zs + k) / lknorm;
            const double r = lknorminv * (ROWD(T, ds + zs + k) - frho * lkb);
            const double s = lknorminv * (ROWD(T, dz + zs + k) - fsig * lkb);
            ar += r * r;
            as += s * s;
        }
        const double rhonorm = sqrt(ar) - rho0, signorm = sqrt(as) - sig0;
        const double conic_step = dmax(0., dmax(signorm, rhonorm));
        if (conic_step != 0.)
            mn[2] = dmin(mn[2], 1. / conic_step);
    }
    team_min<3>(tm, mn);
    if (P.l > 0)
    {
        const double rhomin = mn[0], sigmamin = mn[1];
        const double eps = 1e-13;
        if (-sigmamin > -rhomin)
            alpha = sigmamin < 0. ? 1. / (-sigmamin) : 1. / eps;
        else
            alpha = rhomin < 0. ? 1. / (-rhomin) : 1. / eps;
    }
    else
        alpha = 10.;
    const double mt = -tau / dtau, mk = -kap / dkap;
    if (mt > 0. && mt < alpha)
        alpha = mt;
    if (mk > 0. && mk < alpha)
        alpha = mk;
    alpha = dmin(mn[2], alpha);
    // std::clamp(alpha, stepmin, stepmax)
    if (alpha < Settings::stepmin)
        alpha = Settings::stepmin;
    else if (Settings::stepmax < alpha)
        alpha = Settings::stepmax;
    return alpha;
}

// ------------------------------------------------------------------ numeric LDL' (Eigen factorize, src/eicos.cpp:900,1164)
// Left-looking by column on the fixed pattern, walking the level schedule of the elimination
// tree.  Column j: gather its KKT entries, subtract the contribution of every earlier column k
// with L(j,k) != 0 (row j of L is a contiguous run of the row-ordered copy; the tail of column k
// below row j is a contiguous run of the column-ordered copy), divide by the pivot and store the
// column in both orders.  Everything structural comes from the worker's instruction stream.
EI_DEV void tile_factor(const Team &tm, const KArgs &a, int tile)
{
    const TileMem t = tile_mem(a, tile);
    const bool act = lane_active(tm, t);
    if (!tm.any(act))
        return;
    const DevPattern &P = a.P;
    const Layout &L = a.L;
    double *T = t.T;
    double *acc = tm.acc + (size_t)tm.wk * P.maxcol * TILE + tm.lane;
    bool zero_pivot = false;
    for (int ph = 0; ph < P.nphases; ph++)
    {
        const int *seg = P.fa_seg + ((size_t)ph * tm.nwk + tm.wk) * 3;
        const int nt = EI_LDG(seg + 1);
        if (nt > 0)
        {
            IStream is;
            DStream ds;
            is.open(P.fa + EI_LDG(seg), tm.lane);
            ds.open(P.fa_val + EI_LDG(seg + 2), tm.lane);
            for (int q = 0; q < nt; q++)
            {
                const int j = is.get(), cnt = is.get(), nK = is.get(), nR = is.get();
                for (int c = 0; c < cnt; c++)
                    acc[(size_t)c * TILE] = 0.0;
                double d = 0.0;
                for (int e = 0; e < nK; e++)
                {
                    const int vi = is.get(), pos = is.get();
                    const double val = vi >= 0 ? ROWD(T, L.V + vi) : ds.get();
                    if (pos < 0)
                        d = val;
                    else
                        acc[(size_t)pos * TILE] = val;
                }
                for (int r = 0; r < nR; r++)
                {
                    const int k = is.get(), fp = is.get(), tl = is.get();
                    const double ljk = ROWD(T, L.LTx + fp);
                    const double w = ljk * ROWD(T, L.D + k);
                    d -= ljk * w;
                    for (int u = 0; u < tl; u++)
                    {
                        const int rel = is.get(), bp = is.get();
                        acc[(size_t)rel * TILE] -= ROWD(T, L.Lx + bp) * w;
                    }
                }
                ROWD(T, L.D + j) = d;
                ROWD(T, L.Dinv + j) = 1.0 / d;
                zero_pivot = zero_pivot || (d == 0.0); // Eigen reports NumericalIssue only on an exactly zero pivot
                for (int c = 0; c < cnt; c++)
                {
                    const int bp = is.get(), fp = is.get();
                    const double lv = acc[(size_t)c * TILE] / d;
                    ROWD(T, L.Lx + bp) = lv;
                    ROWD(T, L.LTx + fp) = lv;
                }
            }
        }
        tm.sync();
    }
    if (zero_pivot && act)
        ROWD(t.I, J_STATUS) = EXIT_FATAL;
}

// ------------------------------------------------------------------ triangular solves (Eigen solve, src/eicos.cpp:1477,1599)
// forward:  xw = L^-1 P rhs       (rows of L, dot form; the permutation is folded into the gather)
// backward: out = P' L^-T D^-1 xw (columns of L, dot form; results land in KKT order directly)
// Phases come from the sweep's own stream (streams.hpp): block phases hold independent tasks that
// are staged asynchronously; serial phases hold the chains of the elimination tree, reduced to the
// recurrence between their own members (results of the last three tasks are forwarded in registers).
// Serial stream layout: hdr(0) hdr(1) ent(0) hdr(2) ent(1) ..., so the start value of the next task
// is already being loaded while the current one is computed.
EI_DEV void ldl_forward(const Team &tm, const KArgs &a, double *T, int rhs)
{
    const DevPattern &P = a.P;
    const Layout &L = a.L;
    for (int ph = 0; ph < P.nph_fw; ph++)
    {
        const int *seg = P.fw_seg + ((size_t)ph * tm.nwk + tm.wk) * 4;
        const int count = EI_LDG(seg + 1);
        if (count > 0)
        {
            IStream is;
            is.open(P.fw + EI_LDG(seg), tm.lane);
            const double *lv = T + (size_t)(L.LTx + EI_LDG(seg + 2)) * TILE + tm.lane;
            if (EI_LDG(seg + 3) == SEG_BLOCKS)
            {
                for (int b = 0; b < count; b++)
                {
                    const int nt = is.get();
                    if (nt < 0)
                    { // oversize row: straight from global memory
                        const int i = is.get(), r = is.get(), cnt = is.get();
                        double v = r >= 0 ? ROWD(T, rhs + r) : ROWD(T, L.xw + i);
                        for (int k = 0; k < cnt; k++, lv += TILE)
                            v -= *lv * ROWD(T, L.xw + is.get());
                        ROWD(T, L.xw + i) = v;
                        continue;
                    }
                    const IStream mark = is;
                    double *sp = tm.stage;
                    for (int t = 0; t < nt; t++)
                    {
                        const int i = is.get(), r = is.get(), cnt = is.get();
                        stage_issue(sp, r >= 0 ? rowp(tm, T, rhs + r) : rowp(tm, T, L.xw + i));
                        sp += TILE;
                        for (int k = 0; k < cnt; k++, lv += TILE, sp += 2 * TILE)
                        {
                            stage_issue(sp, lv);
                            stage_issue(sp + TILE, rowp(tm, T, L.xw + is.get()));
                        }
                    }
                    stage_wait();
                    is = mark;
                    sp = tm.stage;
                    for (int t = 0; t < nt; t++)
                    {
                        const int i = is.get();
                        (void)is.get();
                        const int cnt = is.get();
                        double v = *sp;
                        sp += TILE;
                        for (int k = 0; k < cnt; k++, sp += 2 * TILE)
                        {
                            (void)is.get();
                            v -= sp[0] * sp[TILE];
                        }
                        ROWD(T, L.xw + i) = v;
                    }
                }
            }
            else
            {
                double p1 = 0.0, p2 = 0.0, p3 = 0.0;
                int i = is.get(), r = is.get(), cnt = is.get();
                double v = r >= 0 ? ROWD(T, rhs + r) : ROWD(T, L.xw + i);
                for (int q = 0; q < count; q++)
                {
                    int ni = 0, ncnt = 0;
                    double nv = 0.0;
                    if (q + 1 < count)
                    {
                        ni = is.get();
                        r = is.get();
                        ncnt = is.get();
                        nv = r >= 0 ? ROWD(T, rhs + r) : ROWD(T, L.xw + ni);
                    }
                    for (int k = 0; k < cnt; k++, lv += TILE)
                    {
                        const int c = is.get();
                        EI_PREFETCH(lv + (size_t)PF_ROWS * TILE);
                        const double xv = c >= 0 ? ROWD(T, L.xw + c) : (c == FWD_PREV1 ? p1 : (c == FWD_PREV2 ? p2 : p3));
                        v -= *lv * xv;
                    }
                    ROWD(T, L.xw + i) = v;
                    p3 = p2;
                    p2 = p1;
                    p1 = v;
                    i = ni;
                    cnt = ncnt;
                    v = nv;
                }
            }
        }
        tm.sync();
    }
}

// out = solution (KKT order).  If x >= 0: additionally x += solution for the lanes with `cont`.
// Task header: [xw/Dinv row j | INIT_PARTIAL, out row o | ~o for the external part of a chain task].
EI_DEV void ldl_backward(const Team &tm, const KArgs &a, double *T, int out, int x, bool cont)
{
    const DevPattern &P = a.P;
    const Layout &L = a.L;
    const bool accumulate = x >= 0 && cont;
    for (int ph = 0; ph < P.nph_bw; ph++)
    {
        const int *seg = P.bw_seg + ((size_t)ph * tm.nwk + tm.wk) * 4;
        const int count = EI_LDG(seg + 1);
        if (count > 0)
        {
            IStream is;
            is.open(P.bw + EI_LDG(seg), tm.lane);
            const double *lv = T + (size_t)(L.Lx + EI_LDG(seg + 2)) * TILE + tm.lane;
            if (EI_LDG(seg + 3) == SEG_BLOCKS)
            {
                for (int b = 0; b < count; b++)
                {
                    const int nt = is.get();
                    if (nt < 0)
                    {
                        const int j = is.get(), oe = is.get(), cnt = is.get();
                        const int o = oe >= 0 ? oe : ~oe;
                        double v = j >= 0 ? ROWD(T, L.Dinv + j) * ROWD(T, L.xw + j) : ROWD(T, out + o);
                        for (int k = 0; k < cnt; k++, lv += TILE)
                            v -= *lv * ROWD(T, out + is.get());
                        ROWD(T, out + o) = v;
                        if (accumulate && oe >= 0)
                            ROWD(T, x + o) += v;
                        continue;
                    }
                    const IStream mark = is;
                    double *sp = tm.stage;
                    for (int t = 0; t < nt; t++)
                    {
                        const int j = is.get(), oe = is.get(), cnt = is.get();
                        if (j >= 0)
                        {
                            stage_issue(sp, rowp(tm, T, L.Dinv + j));
                            stage_issue(sp + TILE, rowp(tm, T, L.xw + j));
                        }
                        else
                            stage_issue(sp, rowp(tm, T, out + (oe >= 0 ? oe : ~oe)));
                        sp += 2 * TILE;
                        for (int k = 0; k < cnt; k++, lv += TILE, sp += 2 * TILE)
                        {
                            stage_issue(sp, lv);
                            stage_issue(sp + TILE, rowp(tm, T, out + is.get()));
                        }
                    }
                    stage_wait();
                    is = mark;
                    sp = tm.stage;
                    for (int t = 0; t < nt; t++)
                    {
                        const int j = is.get(), oe = is.get(), cnt = is.get();
                        const int o = oe >= 0 ? oe : ~oe;
                        double v = j >= 0 ? sp[0] * sp[TILE] : sp[0];
                        sp += 2 * TILE;
                        for (int k = 0; k < cnt; k++, sp += 2 * TILE)
                        {
                            (void)is.get();
                            v -= sp[0] * sp[TILE];
                        }
                        ROWD(T, out + o) = v;
                        if (accumulate && oe >= 0)
                            ROWD(T, x + o) += v;
                    }
                }
            }
            else
            {
                double p1 = 0.0, p2 = 0.0, p3 = 0.0;
                int j = is.get(), o = is.get(), cnt = is.get();
                double v = j >= 0 ? ROWD(T, L.Dinv + j) * ROWD(T, L.xw + j) : ROWD(T, out + o);
                for (int q = 0; q < count; q++)
                {
                    int no = 0, ncnt = 0;
                    double nv = 0.0;
                    if (q + 1 < count)
                    {
                        j = is.get();
                        no = is.get();
                        ncnt = is.get();
                        nv = j >= 0 ? ROWD(T, L.Dinv + j) * ROWD(T, L.xw + j) : ROWD(T, out + no);
                    }
                    for (int k = 0; k < cnt; k++, lv += TILE)
                    {
                        const int c = is.get();
                        EI_PREFETCH(lv + (size_t)PF_ROWS * TILE);
                        const double xv = c >= 0 ? ROWD(T, out + c) : (c == FWD_PREV1 ? p1 : (c == FWD_PREV2 ? p2 : p3));
                        v -= *lv * xv;
                    }
                    ROWD(T, out + o) = v;
                    if (accumulate)
                        ROWD(T, x + o) += v;
                    p3 = p2;
                    p2 = p1;
                    p1 = v;
                    o = no;
                    cnt = ncnt;
                    v = nv;
                }
            }
        }
        tm.sync();
    }
}

// ------------------------------------------------------------------ KKT residual for iterative refinement (src/eicos.cpp:1511-1576)
// e = rhs - Ktrue * x with the un-regularised scaling block (identity while initialising),
// returns ||e||_inf per lane.
EI_DEV double kkt_residual(const Team &tm, const KArgs &a, double *T, int rhs, int x, bool initialize)
{
    const DevPattern &P = a.P;
    const Layout &L = a.L;
    const double delta = Settings::deltastat;
    const int n = P.n, p = P.p, zb = P.n + P.p;
    double nerr = 0.0;
    if (n > 0)
        rowset_run<2>(
            tm, T, P.rx, P.rx_val, P.rx_seg, 0, x, -1.0,
            [&](int j, int k) { return (k == 0 ? rhs : x) + j; },
            [&](const double *ex) { return ex[0]; },
            [&](int j, const double *ex, double v) {
                v -= delta * ex[1];
                ROWD(T, L.e + j) = v;
                nerr = dmax(nerr, fabs(v));
            });
    if (p > 0)
        rowset_run<2>(
            tm, T, P.ry, P.ry_val, P.ry_seg, 0, x, -1.0,
            [&](int i, int k) { return (k == 0 ? rhs : x) + n + i; },
            [&](const double *ex) { return ex[0]; },
            [&](int i, const double *ex, double v) {
                v += delta * ex[1];
                ROWD(T, L.e + n + i) = v;
                nerr = dmax(nerr, fabs(v));
            });
    if (P.l > 0)
        rowset_run<3>(
            tm, T, P.rz, P.rz_val, P.rz_seg, 0, x, -1.0,
            [&](int i, int k) { return k == 0 ? rhs + zb + i : (k == 1 ? x + zb + i : L.lpv + i); },
            [&](const double *ex) { return ex[0]; },
            [&](int i, const double *ex, double v) {
                const double dz = ex[1];
                v += delta * dz;
                v += initialize ? dz : ex[2] * dz;
                ROWD(T, L.e + zb + i) = v;
                nerr = dmax(nerr, fabs(v));
            });
    if (P.nc > 0)
    {
        IStream is;
        DStream ds;
        is.open(P.rc + EI_LDG(P.rc_seg + tm.wk * 2), tm.lane);
        ds.open(P.rc_val + EI_LDG(P.rc_seg + tm.wk * 2 + 1), tm.lane);
        for (int c = tm.wk; c < P.nc; c += tm.nwk)
        {
            const int d = is.get(), ks = is.get(), qo = is.get();
            const int kb = zb + ks; // KKT row of the cone's first entry
            const double *cp = T + (size_t)(L.cpar + c * CP_COUNT) * TILE + tm.lane;
            const double eta2 = cp[CP_ETA2 * TILE], d1 = cp[CP_D1 * TILE], u0 = cp[CP_U0 * TILE];
            const double u1 = cp[CP_U1 * TILE], v1 = cp[CP_V1 * TILE];
            const double x1 = ROWD(T, x + kb), x3 = ROWD(T, x + kb + d), x4 = ROWD(T, x + kb + d + 1);
            double qtx2 = 0.0;
            for (int k = 1; k < d; k++)
                qtx2 += ROWD(T, L.cq + qo + k - 1) * ROWD(T, x + kb + k);
            const double vu = v1 * x3 + u1 * x4;
            for (int k = 0; k < d; k++)
            {
                const double xk = ROWD(T, x + kb + k);
                double v = row_accumulate(tm, is, ds, T, x, ROWD(T, rhs + kb + k), -1.0);
                if (k < d - 1)
                    v += delta * xk;
                else
                    v -= delta * xk;
                if (initialize)
                    v += xk;
                else if (k == 0)
                    v += eta2 * (d1 * x1 + u0 * x4);
                else
                    v += eta2 * (xk + vu * ROWD(T, L.cq + qo + k - 1));
                ROWD(T, L.e + kb + k) = v;
                nerr = dmax(nerr, fabs(v));
            }
            const double e3 = initialize ? x3 : eta2 * (v1 * qtx2 + x3);
            const double e4 = initialize ? x4 : eta2 * (u0 * x1 + u1 * qtx2 - x4);
            ROWD(T, L.e + kb + d) = e3;
            ROWD(T, L.e + kb + d + 1) = e4;
            nerr = dmax(nerr, dmax(fabs(e3), fabs(e4)));
        }
    }
    double red[1] = {nerr};
    team_max<1>(tm, red);
    return red[0];
}

// ------------------------------------------------------------------ solveKKT (src/eicos.cpp:1471-1620)
// sol = K^-1 rhs followed by up to nitref refinement rounds; every lane stops on its own
// criterion, the tile loops until all of its lanes have stopped.
EI_DEV void tile_solve_kkt(const Team &tm, const KArgs &a, int tile)
{
    const TileMem t = tile_mem(a, tile);
    const bool act = lane_active(tm, t);
    if (!tm.any(act))
        return;
    const DevPattern &P = a.P;
    const Layout &L = a.L;
    double *T = t.T;
    const int rhs = a.rhs, sol = a.sol;
    const bool init = a.initialize != 0;

    double mx[1] = {0.0};
    for (int r = tm.wk; r < P.N; r += tm.nwk)
        mx[0] = dmax(mx[0], fabs(ROWD(T, rhs + r)));
    team_max<1>(tm, mx);
    const double threshold = (1. + mx[0]) * Settings::linsysacc;

    ldl_forward(tm, a, T, rhs);
    ldl_backward(tm, a, T, sol, -1, false);

    double nerr_prev = DBL_MAX;
    int kref = 0;
    bool done = !act;
    unsigned rounds = 0;
    for (;;)
    {
        const double nerr = kkt_residual(tm, a, T, rhs, sol, init);
        bool rollback = false;
        if (!done)
        {
            if (kref > 0 && nerr > nerr_prev)
            {
                rollback = true;
                kref--;
                done = true;
            }
            else if (kref == Settings::nitref || nerr < threshold || (kref > 0 && nerr_prev < Settings::irerrfact * nerr))
                done = true;
            else
                nerr_prev = nerr;
        }
        if (tm.any(rollback))
        { // x -= dx_ref for the lanes whose last refinement made things worse
            for (int r = tm.wk; r < P.N; r += tm.nwk)
                if (rollback)
                    ROWD(T, sol + r) -= ROWD(T, L.dxr + r);
        }
        if (tm.all(done))
            break;
        tm.sync(); // e complete before the forward sweep gathers it
        ldl_forward(tm, a, T, L.e);
        ldl_backward(tm, a, T, L.dxr, sol, !done);
        if (!done)
            kref++;
        rounds++;
    }
    tm.sync();
    if (tm.wk == 0)
    {
        if (act && a.nitrow >= 0)
            ROWD(t.I, a.nitrow) = kref;
#ifndef EICOS_EMU
        if (tm.lane == 0 && a.ir_rounds)
            atomicAdd(a.ir_rounds, (unsigned long long)(rounds + 1));
#endif
    }
}

// ------------------------------------------------------------------ start of a solve (src/eicos.cpp:855-894)
EI_DEV void tile_init(const Team &tm, const KArgs &a, int tile)
{
    const TileMem t = tile_mem(a, tile);
    const DevPattern &P = a.P;
    const Layout &L = a.L;
    double *T = t.T;
    const int n = P.n, zb = P.n + P.p;
    const bool valid = tile * TILE + tm.lane < a.batch;
    if (tm.wk == 0)
    {
        ROWD(t.I, J_STATUS) = valid ? (int)ST_ACTIVE : (int)EXIT_FATAL;
        if (!a.keep_sticky)
        {
            ROWD(t.I, J_HAS_PINFRES) = 0;
            ROWD(t.I, J_HAS_DINFRES) = 0;
            ROWD(t.I, J_HAS_RELGAP) = 0;
            ROWD(T, L.sc + S_PINFRES) = 0.0;
            ROWD(T, L.sc + S_DINFRES) = 0.0;
            ROWD(T, L.sc + S_RELGAP) = 0.0;
        }
    }
    for (int k = tm.wk; k < P.nnzV; k += tm.nwk)
    { // resetKKTScalings :807-846
        const int kind = EI_LDG(P.Vkind + k);
        ROWD(T, L.V + k) = kind == 0 ? -1.0 : (kind == 1 ? 0.0 : 1.0);
    }
    // rhs1 = [0; b; h], rhs2 = [-c; 0; 0]; resx0.. = max(1, ||c||), ... (:865-894)
    double nr[3] = {0.0, 0.0, 0.0};
    for (int r = tm.wk; r < n; r += tm.nwk)
    {
        const double v = ROWD(T, L.chb + r);
        ROWD(T, L.rhs1 + r) = 0.0;
        ROWD(T, L.rhs2 + r) = -v;
        nr[0] += v * v;
    }
    for (int r = n + tm.wk; r < zb; r += tm.nwk)
    {
        const double v = ROWD(T, L.chb + r);
        ROWD(T, L.rhs1 + r) = v;
        ROWD(T, L.rhs2 + r) = 0.0;
        nr[1] += v * v;
    }
    for (int r = zb + tm.wk; r < P.N; r += tm.nwk)
    {
        const double v = ROWD(T, L.chb + r);
        ROWD(T, L.rhs1 + r) = v;
        ROWD(T, L.rhs2 + r) = 0.0;
        nr[2] += v * v;
    }
    team_sum<3>(tm, nr);
    if (tm.wk == 0)
    {
        ROWD(T, L.sc + S_RESX0) = dmax(1., sqrt(nr[0]));
        ROWD(T, L.sc + S_RESY0) = dmax(1., sqrt(nr[1]));
        ROWD(T, L.sc + S_RESZ0) = dmax(1., sqrt(nr[2]));
    }
}

// bringToCone (src/eicos.cpp:761-805): dst = sign*src + (1+alpha) e ; src, dst z-shaped row offsets
EI_DEV void bring_to_cone(const Team &tm, const KArgs &a, double *T, int src, double sign, int dst, bool write)
{
    const DevPattern &P = a.P;
    double al[1] = {-Settings::gamma};
    for (int k = tm.wk; k < P.l; k += tm.nwk)
    {
        const double r = sign * ROWD(T, src + k);
        if (r <= 0 && -r > al[0])
            al[0] = -r;
    }
    for (int c = tm.wk; c < P.nc; c += tm.nwk)
    {
        const int d = EI_LDG(P.cone_dim + c), kb = EI_LDG(P.cone_k + c);
        double sq = 0.0;
        for (int k = 1; k < d; k++)
        {
            const double v = sign * ROWD(T, src + kb + k);
            sq += v * v;
        }
        const double cres = sign * ROWD(T, src + kb) - sqrt(sq);
        if (cres <= 0 && -cres > al[0])
            al[0] = -cres;
    }
    team_max<1>(tm, al);
    const double alpha = al[0] + 1.;
    if (!write)
        return;
    for (int k = tm.wk; k < P.l; k += tm.nwk)
        ROWD(T, dst + k) = sign * ROWD(T, src + k) + alpha;
    for (int c = tm.wk; c < P.nc; c += tm.nwk)
    {
        const int d = EI_LDG(P.cone_dim + c), kb = EI_LDG(P.cone_k + c);
        ROWD(T, dst + kb) = sign * ROWD(T, src + kb) + alpha;
        for (int k = 1; k < d; k++)
            ROWD(T, dst + kb + k) = sign * ROWD(T, src + kb + k);
    }
}

// initial point (src/eicos.cpp:933-992), after the two initial KKT solves
EI_DEV void tile_init_point(const Team &tm, const KArgs &a, int tile)
{
    const TileMem t = tile_mem(a, tile);
    const bool act = lane_active(tm, t);
    if (!tm.any(act))
        return;
    const DevPattern &P = a.P;
    const Layout &L = a.L;
    double *T = t.T;
    const int n = P.n, zb = P.n + P.p;
    for (int j = tm.wk; j < n; j += tm.nwk)
    {
        if (act)
            ROWD(T, L.w + j) = ROWD(T, L.sol1 + j);
        ROWD(T, L.rhs1 + j) = -ROWD(T, L.chb + j);
    }
    for (int i = tm.wk; i < P.p; i += tm.nwk)
        if (act)
            ROWD(T, L.w + n + i) = ROWD(T, L.sol2 + n + i);
    bring_to_cone(tm, a, T, L.sol1 + zb, -1.0, L.s, act);
    bring_to_cone(tm, a, T, L.sol2 + zb, 1.0, L.w + zb, act);
    if (tm.wk == 0 && act)
    {
        ROWD(T, L.sc + S_KAP) = 1.;
        ROWD(T, L.sc + S_TAU) = 1.;
        ROWD(T, L.sc + S_STEP) = 0.;
        ROWD(T, L.sc + S_STEP_AFF) = 0.;
        ROWD(T, L.sc + S_PRES_PREV) = DBL_MAX;
        ROWD(t.I, J_PINF) = 0;
        ROWD(t.I, J_DINF) = 0;
        ROWD(t.I, J_ITER) = 0;
    }
}

// ------------------------------------------------------------------ per-lane scalar state of `Work` + `Information`
struct WState
{
    double d[S_WORK_END];
    int i[J_WORK_END];
};

EI_DEV void ws_load(const Team &tm, const double *T, const int *I, int sc, int soff, int ioff, WState &w)
{
    for (int k = 0; k < S_WORK_END; k++)
        w.d[k] = ROWD(T, sc + soff + k);
    for (int k = 0; k < J_WORK_END; k++)
        w.i[k] = ROWD(I, ioff + k);
}
EI_DEV void ws_store(const Team &tm, double *T, int *I, int sc, int soff, int ioff, const WState &w)
{
    for (int k = 0; k < S_WORK_END; k++)
        ROWD(T, sc + soff + k) = w.d[k];
    for (int k = 0; k < J_WORK_END; k++)
        ROWD(I, ioff + k) = w.i[k];
}

// Information::isBetterThan (src/eicos.cpp:23-68)
EI_DEV bool ws_better(const WState &a, const WState &o)
{
    const bool gapmu = (a.d[S_GAP] > 0. && o.d[S_GAP] > 0. && a.d[S_GAP] < o.d[S_GAP]) &&
                       (a.d[S_MU] > 0. && a.d[S_MU] < o.d[S_MU]);
    if (a.i[J_HAS_PINFRES] && a.d[S_KAPOVERT] > 1.)
    {
        if (o.i[J_HAS_PINFRES])
            return gapmu && (a.d[S_PINFRES] > 0. && a.d[S_PINFRES] < o.d[S_PRES]);
        return gapmu;
    }
    return gapmu && (a.d[S_PRES] > 0. && a.d[S_PRES] < o.d[S_PRES]) &&
           (a.d[S_DRES] > 0. && a.d[S_DRES] < o.d[S_DRES]) &&
           (a.d[S_KAPOVERT] > 0. && a.d[S_KAPOVERT] < o.d[S_KAPOVERT]);
}

// checkExitConditions (src/eicos.cpp:526-641).  An empty std::optional compares "less than"
// any double there, which is what the (!has || v < tol) terms reproduce.
EI_DEV int ws_check_exit(WState &w, bool reduced)
{
    const double feastol = reduced ? Settings::feastol_inacc : Settings::feastol;
    const double abstol = reduced ? Settings::abstol_inacc : Settings::abstol;
    const double reltol = reduced ? Settings::reltol_inacc : Settings::reltol;
    const double tau = w.d[S_TAU], kap = w.d[S_KAP];
    if ((-w.d[S_CX] > 0. || -w.d[S_BY] - w.d[S_HZ] >= -abstol) &&
        (w.d[S_PRES] < feastol && w.d[S_DRES] < feastol) &&
        (w.d[S_GAP] < abstol || !w.i[J_HAS_RELGAP] || w.d[S_RELGAP] < reltol))
    {
        w.i[J_PINF] = 0;
        w.i[J_DINF] = 0;
        return reduced ? EXIT_OPTIMAL + EXIT_INACC : EXIT_OPTIMAL;
    }
    else if (w.i[J_HAS_DINFRES] && w.d[S_DINFRES] < feastol && tau < kap)
    {
        w.i[J_PINF] = 0;
        w.i[J_DINF] = 1;
        return reduced ? EXIT_DINF + EXIT_INACC : EXIT_DINF;
    }
    else if ((w.i[J_HAS_PINFRES] && w.d[S_PINFRES] < feastol && tau < kap) ||
             (tau < feastol && kap < feastol && (!w.i[J_HAS_PINFRES] || w.d[S_PINFRES] < feastol)))
    {
        w.i[J_PINF] = 1;
        w.i[J_DINF] = 0;
        return reduced ? EXIT_PINF + EXIT_INACC : EXIT_PINF;
    }
    return EXIT_NOT_CONVERGED;
}

// ------------------------------------------------------------------ NT scaling of one cone (src/eicos.cpp:419-474)
struct ConeScaling
{
    int stage; // 0 ok, 1 failed the residual test (nothing assigned), 2 failed c^2/u0^2 - d (eta, q assigned)
    double eta, eta2, a, d1, u0, u1, v1, w, snorm, znorm, gamma;
};

EI_DEV ConeScaling cone_scaling(const Team &tm, const double *T, int srow, int zrow, int d)
{
    ConeScaling r;
    const double s0 = ROWD(T, srow), z0 = ROWD(T, zrow);
    double ss = 0.0, zz = 0.0;
    for (int k = 1; k < d; k++)
    {
        const double sv = ROWD(T, srow + k), zv = ROWD(T, zrow + k);
        ss += sv * sv;
        zz += zv * zv;
    }
    const double sres = s0 * s0 - ss, zres = z0 * z0 - zz;
    r.stage = 0;
    if (sres <= 0 || zres <= 0)
    {
        r.stage = 1;
        return r;
    }
    r.snorm = sqrt(sres);
    r.znorm = sqrt(zres);
    r.eta2 = r.snorm / r.znorm;
    r.eta = sqrt(r.eta2);
    double g = 0.0;
    for (int k = 0; k < d; k++)
        g += (ROWD(T, srow + k) / r.snorm) * (ROWD(T, zrow + k) / r.znorm);
    g = sqrt(0.5 * (1. + g));
    r.gamma = g;
    const double av = (0.5 / g) * (s0 / r.snorm + z0 / r.znorm);
    double w = 0.0;
    for (int k = 1; k < d; k++)
    {
        const double qk = (0.5 / g) * (ROWD(T, srow + k) / r.snorm - ROWD(T, zrow + k) / r.znorm);
        w += qk * qk;
    }
    const double cc = (1. + av) + w / (1. + av);
    const double dd = 1. + 2. / (1. + av) + w / ((1. + av) * (1. + av));
    const double d1 = dmax(0., 0.5 * (av * av + w * (1. - (cc * cc) / (1. + w * dd))));
    const double u0sq = av * av + w - d1;
    const double c2byu02 = (cc * cc) / u0sq;
    if (c2byu02 - dd <= 0)
    {
        r.stage = 2;
        return r;
    }
    r.d1 = d1;
    r.u0 = sqrt(u0sq);
    r.u1 = sqrt(c2byu02);
    r.v1 = sqrt(c2byu02 - dd);
    r.a = av;
    r.w = w;
    return r;
}

// ------------------------------------------------------------------ head of an iteration
// computeResiduals + updateStatistics + safeguards / exit tests + best-iterate bookkeeping
// (src/eicos.cpp:997-1158), then for the lanes that keep iterating: updateScalings,
// updateKKTScalings and RHSaffine (:1160-1162, :1176).  Lanes that stop are back-scaled in place.
EI_DEV void tile_head(const Team &tm, const KArgs &a, int tile)
{
    const TileMem t = tile_mem(a, tile);
    const bool act = lane_active(tm, t);
    if (!tm.any(act))
        return;
    const DevPattern &P = a.P;
    const Layout &L = a.L;
    double *T = t.T;
    int *I = t.I;
    const int n = P.n, p = P.p, zb = P.n + P.p;
    const double tau = ROWD(T, L.sc + S_TAU), kap = ROWD(T, L.sc + S_KAP);

    enum { HX2, RX2, CX, NX2, HY2, RY2, BY, NY2, HZ2, RZ2, HZ, NZ2, NS2, GAP, NRED };
    double r[NRED];
    for (int k = 0; k < NRED; k++)
        r[k] = 0.0;
    if (n > 0)
        rowset_run<2>(
            tm, T, P.rx, P.rx_val, P.rx_seg, 0, L.w, -1.0,
            [&](int j, int k) { return (k == 0 ? L.chb : L.w) + j; },
            [&](const double *) { return 0.0; },
            [&](int j, const double *ex, double v) {
                const double cj = ex[0], xj = ex[1];
                r[HX2] += v * v;
                v -= tau * cj;
                ROWD(T, L.r + j) = v;
                r[RX2] += v * v;
                r[CX] += cj * xj;
                r[NX2] += xj * xj;
            });
    if (p > 0)
        rowset_run<2>(
            tm, T, P.ry, P.ry_val, P.ry_seg, 0, L.w, 1.0,
            [&](int i, int k) { return (k == 0 ? L.chb : L.w) + n + i; },
            [&](const double *) { return 0.0; },
            [&](int i, const double *ex, double v) {
                const double bi = ex[0], yi = ex[1];
                r[HY2] += v * v;
                v -= tau * bi;
                ROWD(T, L.r + n + i) = v;
                r[RY2] += v * v;
                r[BY] += bi * yi;
                r[NY2] += yi * yi;
            });
    if (P.l > 0)
        rowset_run<3>(
            tm, T, P.rz, P.rz_val, P.rz_seg, 0, L.w, 1.0,
            [&](int i, int k) { return k == 0 ? L.s + i : (k == 1 ? L.w + zb + i : L.chb + zb + i); },
            [&](const double *ex) { return ex[0]; },
            [&](int i, const double *ex, double v) {
                const double si = ex[0], zi = ex[1], hi = ex[2];
                r[HZ2] += v * v;
                v -= tau * hi;
                ROWD(T, L.r + zb + i) = v;
                r[RZ2] += v * v;
                r[HZ] += hi * zi;
                r[NZ2] += zi * zi;
                r[NS2] += si * si;
                r[GAP] += si * zi;
            });
    if (P.nc > 0)
    {
        IStream is;
        DStream ds;
        is.open(P.rc + EI_LDG(P.rc_seg + tm.wk * 2), tm.lane);
        ds.open(P.rc_val + EI_LDG(P.rc_seg + tm.wk * 2 + 1), tm.lane);
        for (int c = tm.wk; c < P.nc; c += tm.nwk)
        {
            const int d = is.get(), ks = is.get();
            (void)is.get();
            for (int k = 0; k < d; k++)
            {
                const int e = ks + k;
                const double si = ROWD(T, L.s + e), zi = ROWD(T, L.w + zb + e), hi = ROWD(T, L.chb + zb + e);
                double v = row_accumulate(tm, is, ds, T, L.w, si, 1.0);
                r[HZ2] += v * v;
                v -= tau * hi;
                ROWD(T, L.r + zb + e) = v;
                r[RZ2] += v * v;
                r[HZ] += hi * zi;
                r[NZ2] += zi * zi;
                r[NS2] += si * si;
                r[GAP] += si * zi;
            }
        }
    }
    team_sum<NRED>(tm, r);

    // ---- updateStatistics (:691-728) on registers; every warp computes the same values
    WState w, best;
    ws_load(tm, T, I, L.sc, 0, 0, w);
    ws_load(tm, T, I, L.sc, S_BEST, J_BEST, best);
    const double hresx = sqrt(r[HX2]), hresy = sqrt(r[HY2]), hresz = sqrt(r[HZ2]);
    const double nx = sqrt(r[NX2]), ny = sqrt(r[NY2]), nz = sqrt(r[NZ2]), ns = sqrt(r[NS2]);
    const double cx = r[CX], by = p > 0 ? r[BY] : 0., hz = r[HZ];
    const double rt = kap + cx + by + hz;
    w.d[S_CX] = cx;
    w.d[S_BY] = by;
    w.d[S_HZ] = hz;
    w.d[S_GAP] = r[GAP];
    w.d[S_MU] = (r[GAP] + kap * tau) / ((P.l + P.nc) + 1);
    w.d[S_KAPOVERT] = kap / tau;
    w.d[S_PCOST] = cx / tau;
    w.d[S_DCOST] = -(hz + by) / tau;
    if (w.d[S_PCOST] < 0.)
    {
        w.i[J_HAS_RELGAP] = 1;
        w.d[S_RELGAP] = w.d[S_GAP] / (-w.d[S_PCOST]);
    }
    else if (w.d[S_DCOST] > 0.)
    {
        w.i[J_HAS_RELGAP] = 1;
        w.d[S_RELGAP] = w.d[S_GAP] / w.d[S_DCOST];
    }
    else
        w.i[J_HAS_RELGAP] = 0;
    const double resx0 = ROWD(T, L.sc + S_RESX0), resy0 = ROWD(T, L.sc + S_RESY0), resz0 = ROWD(T, L.sc + S_RESZ0);
    const double nry = p > 0 ? sqrt(r[RY2]) / dmax(resy0 + nx, 1.) : 0.;
    const double nrz = sqrt(r[RZ2]) / dmax(resz0 + nx + ns, 1.);
    w.d[S_PRES] = dmax(nry, nrz) / tau;
    w.d[S_DRES] = sqrt(r[RX2]) / dmax(resx0 + ny + nz, 1.) / tau;
    if ((hz + by) / dmax(ny + nz, 1.) < -Settings::reltol)
    {
        w.i[J_HAS_PINFRES] = 1;
        w.d[S_PINFRES] = hresx / dmax(ny + nz, 1.);
    }
    if (cx / dmax(nx, 1.) < -Settings::reltol)
    {
        w.i[J_HAS_DINFRES] = 1;
        w.d[S_DINFRES] = dmax(hresy / dmax(nx, 1.), hresz / dmax(nx + ns, 1.));
    }

    // ---- safeguards, exit tests, best iterate (:1010-1158)
    const int iter = w.i[J_ITER];
    double pres_prev = ROWD(T, L.sc + S_PRES_PREV);
    int code = ST_ACTIVE;
    bool restore = false, save = false;
    if (act)
    {
        if (iter > 0 && (w.d[S_PRES] > Settings::safeguard * pres_prev || w.d[S_GAP] < 0.))
        {
            restore = true;
            w = best;
            code = ws_check_exit(w, true);
            if (code == EXIT_NOT_CONVERGED)
                code = EXIT_NUMERICS;
        }
        else
        {
            pres_prev = w.d[S_PRES];
            code = ws_check_exit(w, false);
            if (code == EXIT_NOT_CONVERGED)
            {
                if (iter > 0 && w.d[S_STEP] == Settings::stepmin * Settings::gamma)
                {
                    restore = true;
                    w = best;
                    code = ws_check_exit(w, true);
                    if (code == EXIT_NOT_CONVERGED)
                        code = EXIT_NUMERICS;
                }
                else if (iter == Settings::iter_max)
                {
                    if (!ws_better(w, best))
                    {
                        restore = true;
                        w = best;
                    }
                    code = ws_check_exit(w, true);
                    if (code == EXIT_NOT_CONVERGED)
                        code = EXIT_MAXIT;
                }
                else if (ei_isnan(w.d[S_PCOST]))
                {
                    if (!(iter == 0 || ws_better(w, best)))
                    {
                        restore = true;
                        w = best;
                        code = ws_check_exit(w, true);
                        if (code == EXIT_NOT_CONVERGED)
                            code = EXIT_NUMERICS;
                    } // else: the reference leaves not_converged_yet (-87) in place (:1117-1121)
                }
                else
                    code = ST_ACTIVE;
            }
        }
        if (code == ST_ACTIVE && (iter == 0 || ws_better(w, best)))
        {
            save = true;
            best = w;
        }
    }
    const bool fin = act && code != ST_ACTIVE;
    const bool cont = act && !fin;
    tm.sync(); // every warp has read the old state rows before warp 0 overwrites them
    if (tm.wk == 0 && act)
    {
        ws_store(tm, T, I, L.sc, 0, 0, w);
        if (save)
            ws_store(tm, T, I, L.sc, S_BEST, J_BEST, best);
        ROWD(T, L.sc + S_RT) = rt;
        ROWD(T, L.sc + S_PRES_PREV) = pres_prev;
        ROWD(I, J_STATUS) = code;
    }

    // ---- vector part of `w = w_best` / `w_best = w`, and backscale (:1271-1277) for finished lanes
    const double ftau = w.d[S_TAU];
    if (tm.any(restore || save || fin))
    {
        for (int q = tm.wk; q < P.N; q += tm.nwk)
        {
            double v = ROWD(T, L.w + q);
            if (restore)
                v = ROWD(T, L.wb + q);
            if (save)
                ROWD(T, L.wb + q) = v;
            if (fin)
            {
                const double eq = q < n ? EI_LDG(P.xeq + q) : (q < zb ? EI_LDG(P.Aeq + q - n) : EI_LDG(P.GeqE + q - zb));
                v = v / (eq * ftau);
            }
            if (restore || fin)
                ROWD(T, L.w + q) = v;
        }
        for (int e = tm.wk; e < P.mt; e += tm.nwk)
        {
            double sv = ROWD(T, L.s + e), lv = ROWD(T, L.lam + e);
            if (restore)
            {
                sv = ROWD(T, L.bs + e);
                lv = ROWD(T, L.blam + e);
            }
            if (save)
            {
                ROWD(T, L.bs + e) = sv;
                ROWD(T, L.blam + e) = lv;
            }
            if (fin)
                sv = sv * (EI_LDG(P.GeqE + e) / ftau);
            if (restore || fin)
            {
                ROWD(T, L.s + e) = sv;
                ROWD(T, L.lam + e) = lv;
            }
        }
    }
    if (!tm.any(cont))
        return;
#ifndef EICOS_EMU
    if (tm.wk == 0 && a.active_count)
    {
        const unsigned bal = __ballot_sync(0xffffffffu, cont);
        if (tm.lane == 0)
            atomicAdd(a.active_count, (unsigned)__popc(bal));
    }
#else
    if (tm.wk == 0 && a.active_count && cont)
        *a.active_count += 1;
#endif

    // ---- updateScalings (:411-479); its return value is ignored by the caller (:1160), so after a
    // failure at cone c the LP part and cones < c are new, cone c is partly new and lambda is stale.
    const int sz = L.w + zb; // z rows of the iterate
    for (int k = tm.wk; k < P.l; k += tm.nwk)
    {
        const double v = ROWD(T, L.s + k) / ROWD(T, sz + k);
        if (cont)
        {
            ROWD(T, L.lpv + k) = v;
            ROWD(T, L.lpw + k) = sqrt(v);
        }
    }
    int first_fail = P.nc;
    if (P.nc > 0)
    {
        double ff[1] = {(double)P.nc};
        for (int c = tm.wk; c < P.nc; c += tm.nwk)
        {
            const int ks = EI_LDG(P.cone_k + c);
            const ConeScaling cs = cone_scaling(tm, T, L.s + ks, sz + ks, EI_LDG(P.cone_dim + c));
            if (cs.stage != 0)
                ff[0] = dmin(ff[0], (double)c);
        }
        team_min<1>(tm, ff);
        first_fail = (int)ff[0];
        for (int c = tm.wk; c < P.nc; c += tm.nwk)
        {
            if (!cont || c > first_fail)
                continue; // (divergent per lane, cone-local work only)
            const int d = EI_LDG(P.cone_dim + c), ks = EI_LDG(P.cone_k + c), qo = EI_LDG(P.cone_q + c);
            const ConeScaling cs = cone_scaling(tm, T, L.s + ks, sz + ks, d);
            if (cs.stage == 1)
                continue;
            double *cp = T + (size_t)(L.cpar + c * CP_COUNT) * TILE + tm.lane;
            cp[CP_ETA2 * TILE] = cs.eta2;
            cp[CP_ETA * TILE] = cs.eta;
            for (int k = 1; k < d; k++)
                ROWD(T, L.cq + qo + k - 1) = (0.5 / cs.gamma) * (ROWD(T, L.s + ks + k) / cs.snorm - ROWD(T, sz + ks + k) / cs.znorm);
            if (cs.stage == 2)
                continue;
            cp[CP_D1 * TILE] = cs.d1;
            cp[CP_U0 * TILE] = cs.u0;
            cp[CP_U1 * TILE] = cs.u1;
            cp[CP_V1 * TILE] = cs.v1;
            cp[CP_A * TILE] = cs.a;
            cp[CP_W * TILE] = cs.w;
        }
        tm.sync();
    }
    cone_scale(tm, a, T, sz, L.lam, cont && first_fail == P.nc);

    // ---- updateKKTScalings (:1691-1732) into the V rows, RHSaffine (:1670-1689) into rhs2
    const double delta = Settings::deltastat;
    for (int k = tm.wk; k < P.l; k += tm.nwk)
        ROWD(T, L.V + k) = -ROWD(T, L.lpv + k) - delta;
    for (int c = tm.wk; c < P.nc; c += tm.nwk)
    {
        const int d = EI_LDG(P.cone_dim + c), ks = EI_LDG(P.cone_k + c), qo = EI_LDG(P.cone_q + c);
        const double *cp = T + (size_t)(L.cpar + c * CP_COUNT) * TILE + tm.lane;
        const double eta2 = cp[CP_ETA2 * TILE], d1 = cp[CP_D1 * TILE], u0 = cp[CP_U0 * TILE];
        const double u1 = cp[CP_U1 * TILE], v1 = cp[CP_V1 * TILE];
        // cones before c contributed sum(3 dim + 1) V entries; ks - 2c - l = sum of their dims
        int vb = L.V + P.l + 3 * (ks - 2 * c - P.l) + c;
        ROWD(T, vb++) = -eta2 * d1 - delta;
        for (int k = 1; k < d; k++)
            ROWD(T, vb++) = -eta2 - delta;
        ROWD(T, vb++) = -eta2;
        for (int k = 1; k < d; k++)
            ROWD(T, vb++) = -eta2 * v1 * ROWD(T, L.cq + qo + k - 1);
        ROWD(T, vb++) = eta2 + delta;
        ROWD(T, vb++) = -eta2 * u0;
        for (int k = 1; k < d; k++)
            ROWD(T, vb++) = -eta2 * u1 * ROWD(T, L.cq + qo + k - 1);
    }
    for (int q = tm.wk; q < P.N; q += tm.nwk)
    { // [rx; -ry; s - rz], the slot rows of s and rz are zero
        const double rv = ROWD(T, L.r + q);
        ROWD(T, L.rhs2 + q) = q < n ? rv : (q < zb ? -rv : ROWD(T, L.s + q - zb) - rv);
    }
}

// ------------------------------------------------------------------ affine step -> centering -> combined RHS
// (src/eicos.cpp:1181-1210 with RHScombined :1282-1325, conicProduct :1357, conicDivision :1330)
EI_DEV void tile_mid(const Team &tm, const KArgs &a, int tile)
{
    const TileMem t = tile_mem(a, tile);
    const bool act = lane_active(tm, t);
    if (!tm.any(act))
        return;
    const DevPattern &P = a.P;
    const Layout &L = a.L;
    double *T = t.T;
    const int n = P.n, p = P.p, zb = P.n + P.p;
    const double tau = ROWD(T, L.sc + S_TAU), kap = ROWD(T, L.sc + S_KAP), rt = ROWD(T, L.sc + S_RT);
    const double mu = ROWD(T, L.sc + S_MU);

    // c'dx, b'dy, h'dz for both solutions; the slot rows of h are zero but are skipped anyway
    double dt[6] = {0, 0, 0, 0, 0, 0};
    for (int q = tm.wk; q < n; q += tm.nwk)
    {
        const double cv = ROWD(T, L.chb + q);
        dt[0] += cv * ROWD(T, L.sol1 + q);
        dt[3] += cv * ROWD(T, L.sol2 + q);
    }
    for (int q = n + tm.wk; q < zb; q += tm.nwk)
    {
        const double cv = ROWD(T, L.chb + q);
        dt[1] += cv * ROWD(T, L.sol1 + q);
        dt[4] += cv * ROWD(T, L.sol2 + q);
    }
    for (int q = zb + tm.wk; q < zb + P.l; q += tm.nwk)
    {
        const double cv = ROWD(T, L.chb + q);
        dt[2] += cv * ROWD(T, L.sol1 + q);
        dt[5] += cv * ROWD(T, L.sol2 + q);
    }
    for (int c = tm.wk; c < P.nc; c += tm.nwk)
    {
        const int d = EI_LDG(P.cone_dim + c), kb = zb + EI_LDG(P.cone_k + c);
        for (int k = 0; k < d; k++)
        {
            const double hv = ROWD(T, L.chb + kb + k);
            dt[2] += hv * ROWD(T, L.sol1 + kb + k);
            dt[5] += hv * ROWD(T, L.sol2 + kb + k);
        }
    }
    team_sum<6>(tm, dt);
    const double dtau_denom = kap / tau - dt[0] - dt[1] - dt[2];
    const double dtauaff = (rt - kap + dt[3] + dt[4] + dt[5]) / dtau_denom;
    for (int e = tm.wk; e < P.mt; e += tm.nwk) // dz2 += dtauaff * dz1 (slot rows are never read again)
        ROWD(T, L.sol2 + zb + e) += dtauaff * ROWD(T, L.sol1 + zb + e);
    tm.sync();
    cone_scale(tm, a, T, L.sol2 + zb, L.wdz, true);
    tm.sync();
    for (int e = tm.wk; e < P.mt; e += tm.nwk)
        ROWD(T, L.dsw + e) = -ROWD(T, L.wdz + e) - ROWD(T, L.lam + e);
    tm.sync();
    const double dkapaff = -kap - kap / tau * dtauaff;
    const double step_aff = line_search(tm, a, T, L.lam, L.dsw, L.wdz, tau, dtauaff, kap, dkapaff);
    const double om = 1. - step_aff;
    double sigma = om * om * om; // std::pow(x, 3)
    if (sigma < Settings::sigmamin)
        sigma = Settings::sigmamin;
    else if (Settings::sigmamax < sigma)
        sigma = Settings::sigmamax;
    if (tm.wk == 0 && act)
    {
        ROWD(T, L.sc + S_DTAU_DENOM) = dtau_denom;
        ROWD(T, L.sc + S_DTAUAFF) = dtauaff;
        ROWD(T, L.sc + S_DKAPAFF) = dkapaff;
        ROWD(T, L.sc + S_STEP_AFF) = step_aff;
        ROWD(T, L.sc + S_SIGMA) = sigma;
    }

    // RHScombined
    const double sigmamu = sigma * mu, oms = 1. - sigma;
    for (int r = tm.wk; r < n + p; r += tm.nwk)
        ROWD(T, L.rhs2 + r) *= oms;
    for (int k = tm.wk; k < P.l; k += tm.nwk)
    {
        const double lk = ROWD(T, L.lam + k);
        double d1 = lk * lk;
        d1 += ROWD(T, L.dsw + k) * ROWD(T, L.wdz + k);
        d1 -= sigmamu;
        const double q = d1 / lk; // conicDivision, LP part
        ROWD(T, L.dsw + k) = q;
        ROWD(T, L.rhs2 + zb + k) = -oms * ROWD(T, L.r + zb + k) + ROWD(T, L.lpw + k) * q;
    }
    for (int c = tm.wk; c < P.nc; c += tm.nwk)
    {
        const int d = EI_LDG(P.cone_dim + c), zs = EI_LDG(P.cone_k + c), qo = EI_LDG(P.cone_q + c);
        const int kb = zb + zs;
        const double *cp = T + (size_t)(L.cpar + c * CP_COUNT) * TILE + tm.lane;
        const double eta = cp[CP_ETA * TILE], ca = cp[CP_A * TILE];
        // ds1 = lambda o lambda ; ds2 = (W\ds_aff) o (W dz_aff)
        const double l0 = ROWD(T, L.lam + zs), u0 = ROWD(T, L.dsw + zs), v0 = ROWD(T, L.wdz + zs);
        double ll = 0.0, uv = 0.0;
        for (int k = 0; k < d; k++)
        {
            const double lk = ROWD(T, L.lam + zs + k);
            ll += lk * lk;
            uv += ROWD(T, L.dsw + zs + k) * ROWD(T, L.wdz + zs + k);
        }
        double w0 = ll - sigmamu;
        w0 += uv;
        ROWD(T, L.ds1 + zs) = w0;
        for (int k = 1; k < d; k++)
        {
            const double lk = ROWD(T, L.lam + zs + k);
            double v = l0 * lk + l0 * lk;
            v += u0 * ROWD(T, L.wdz + zs + k) + v0 * ROWD(T, L.dsw + zs + k);
            ROWD(T, L.ds1 + zs + k) = v;
        }
        // dsw = lambda \ ds1
        double rho = 0.0, zeta = 0.0;
        for (int k = 1; k < d; k++)
        {
            const double lk = ROWD(T, L.lam + zs + k);
            rho += lk * lk;
            zeta += lk * ROWD(T, L.ds1 + zs + k);
        }
        rho = l0 * l0 - rho;
        const double factor = (zeta / l0 - w0) / rho;
        const double q0 = (l0 * w0 - zeta) / rho;
        ROWD(T, L.dsw + zs) = q0;
        for (int k = 1; k < d; k++)
            ROWD(T, L.dsw + zs + k) = factor * ROWD(T, L.lam + zs + k) + ROWD(T, L.ds1 + zs + k) / l0;
        // ds1 = W * dsw, then the cone rows of rhs2
        double zt = 0.0;
        for (int k = 1; k < d; k++)
            zt += ROWD(T, L.cq + qo + k - 1) * ROWD(T, L.dsw + zs + k);
        const double fz = q0 + zt / (1. + ca);
        ROWD(T, L.rhs2 + kb) = -oms * ROWD(T, L.r + kb) + eta * (ca * q0 + zt);
        for (int k = 1; k < d; k++)
            ROWD(T, L.rhs2 + kb + k) = -oms * ROWD(T, L.r + kb + k) +
                                       eta * (ROWD(T, L.dsw + zs + k) + fz * ROWD(T, L.cq + qo + k - 1));
        ROWD(T, L.rhs2 + kb + d) = 0.0;
        ROWD(T, L.rhs2 + kb + d + 1) = 0.0;
    }
}

// ------------------------------------------------------------------ combined step and iterate update (src/eicos.cpp:1214-1252)
EI_DEV void tile_tail(const Team &tm, const KArgs &a, int tile)
{
    const TileMem t = tile_mem(a, tile);
    const bool act = lane_active(tm, t);
    if (!tm.any(act))
        return;
    const DevPattern &P = a.P;
    const Layout &L = a.L;
    double *T = t.T;
    const int n = P.n, zb = P.n + P.p;
    const double tau = ROWD(T, L.sc + S_TAU), kap = ROWD(T, L.sc + S_KAP), rt = ROWD(T, L.sc + S_RT);
    const double mu = ROWD(T, L.sc + S_MU), sigma = ROWD(T, L.sc + S_SIGMA);
    const double dtau_denom = ROWD(T, L.sc + S_DTAU_DENOM), dtauaff = ROWD(T, L.sc + S_DTAUAFF);
    const double dkapaff = ROWD(T, L.sc + S_DKAPAFF);

    double dt[3] = {0, 0, 0};
    for (int q = tm.wk; q < n; q += tm.nwk)
        dt[0] += ROWD(T, L.chb + q) * ROWD(T, L.sol2 + q);
    for (int q = n + tm.wk; q < zb; q += tm.nwk)
        dt[1] += ROWD(T, L.chb + q) * ROWD(T, L.sol2 + q);
    for (int q = zb + tm.wk; q < zb + P.l; q += tm.nwk)
        dt[2] += ROWD(T, L.chb + q) * ROWD(T, L.sol2 + q);
    for (int c = tm.wk; c < P.nc; c += tm.nwk)
    {
        const int d = EI_LDG(P.cone_dim + c), kb = zb + EI_LDG(P.cone_k + c);
        for (int k = 0; k < d; k++)
            dt[2] += ROWD(T, L.chb + kb + k) * ROWD(T, L.sol2 + kb + k);
    }
    team_sum<3>(tm, dt);
    const double bkap = kap * tau + dkapaff * dtauaff - sigma * mu;
    const double dtau = ((1. - sigma) * rt - bkap / tau + dt[0] + dt[1] + dt[2]) / dtau_denom;
    for (int r = tm.wk; r < P.N; r += tm.nwk)
        ROWD(T, L.sol2 + r) += dtau * ROWD(T, L.sol1 + r);
    tm.sync();
    cone_scale(tm, a, T, L.sol2 + zb, L.wdz, true);
    tm.sync();
    for (int e = tm.wk; e < P.mt; e += tm.nwk)
        ROWD(T, L.dsw + e) = -(ROWD(T, L.dsw + e) + ROWD(T, L.wdz + e));
    tm.sync();
    const double dkap = -(bkap + kap * dtau) / tau;
    const double step = Settings::gamma * line_search(tm, a, T, L.lam, L.dsw, L.wdz, tau, dtau, kap, dkap);
    cone_scale(tm, a, T, L.dsw, L.dsaff, true);
    tm.sync();
    if (!act)
        return; // no barriers below
    for (int q = tm.wk; q < zb + P.l; q += tm.nwk)
        ROWD(T, L.w + q) += step * ROWD(T, L.sol2 + q);
    for (int c = tm.wk; c < P.nc; c += tm.nwk)
    { // cone rows of z; the expansion slots of the iterate stay zero
        const int d = EI_LDG(P.cone_dim + c), kb = zb + EI_LDG(P.cone_k + c);
        for (int k = 0; k < d; k++)
            ROWD(T, L.w + kb + k) += step * ROWD(T, L.sol2 + kb + k);
    }
    for (int e = tm.wk; e < P.mt; e += tm.nwk)
        ROWD(T, L.s + e) += step * ROWD(T, L.dsaff + e);
    if (tm.wk == 0)
    {
        ROWD(T, L.sc + S_KAP) = kap + step * dkap;
        ROWD(T, L.sc + S_TAU) = tau + step * dtau;
        ROWD(T, L.sc + S_STEP) = step;
        ROWD(t.I, J_ITER) += 1;
    }
}

// ------------------------------------------------------------------ data in / results out
// Inputs are instance-major (what the C ABI receives); they are equilibrated on the way in
// (c / x_equil, h / G_equil, b / A_equil : src/eicos.cpp:364-371) and land KKT-shaped in `chb`.
EI_DEV void tile_load(const Team &tm, const KArgs &a, int tile)
{
    const TileMem t = tile_mem(a, tile);
    const DevPattern &P = a.P;
    const Layout &L = a.L;
    const int n = P.n, zb = P.n + P.p;
    int inst = tile * TILE + tm.lane;
    if (inst >= a.batch)
        inst = a.batch - 1; // padding lanes replay the last instance; their results are never stored
    const size_t g = (size_t)a.first + inst;
    for (int j = tm.wk; j < n; j += tm.nwk)
    {
        const double v = a.in_c ? a.in_c[g * n + j] : EI_LDG(a.base_c + j);
        ROWD(t.T, L.chb + j) = a.pre_equilibrated ? v : v / EI_LDG(P.xeq + j);
    }
    for (int i = tm.wk; i < P.p; i += tm.nwk)
    {
        const double v = a.in_b ? a.in_b[g * P.p + i] : EI_LDG(a.base_b + i);
        ROWD(t.T, L.chb + n + i) = a.pre_equilibrated ? v : v / EI_LDG(P.Aeq + i);
    }
    for (int i = tm.wk; i < P.m; i += tm.nwk)
    {
        const double v = a.in_h ? a.in_h[g * P.m + i] : EI_LDG(a.base_h + i);
        const int e = EI_LDG(P.zk + i);
        ROWD(t.T, L.chb + zb + e) = a.pre_equilibrated ? v : v / EI_LDG(P.GeqE + e);
    }
}

EI_DEV void tile_store(const Team &tm, const KArgs &a, int tile)
{
    const TileMem t = tile_mem(a, tile);
    const DevPattern &P = a.P;
    const Layout &L = a.L;
    const int n = P.n, zb = P.n + P.p;
    const int inst = tile * TILE + tm.lane;
    if (inst >= a.batch)
        return;
    const size_t g = (size_t)a.first + inst;
    if (a.out_x)
        for (int j = tm.wk; j < n; j += tm.nwk)
            a.out_x[g * n + j] = ROWD(t.T, L.w + j);
    if (a.out_y)
        for (int i = tm.wk; i < P.p; i += tm.nwk)
            a.out_y[g * P.p + i] = ROWD(t.T, L.w + n + i);
    if (a.out_z)
        for (int i = tm.wk; i < P.m; i += tm.nwk)
            a.out_z[g * P.m + i] = ROWD(t.T, L.w + zb + EI_LDG(P.zk + i));
    if (a.out_s)
        for (int i = tm.wk; i < P.m; i += tm.nwk)
            a.out_s[g * P.m + i] = ROWD(t.T, L.s + EI_LDG(P.zk + i));
    if (tm.wk == 0)
    {
        if (a.out_exit)
            a.out_exit[g] = ROWD(t.I, J_STATUS);
        if (a.out_iter)
            a.out_iter[g] = ROWD(t.I, J_ITER);
        if (a.out_info)
            for (int k = 0; k < S_WORK_END; k++)
                a.out_info[g * S_WORK_END + k] = ROWD(t.T, L.sc + k);
        if (a.out_iinfo)
            for (int k = 0; k < J_WORK_END; k++)
                a.out_iinfo[g * J_WORK_END + k] = ROWD(t.I, k);
    }
}

} // namespace eicos
