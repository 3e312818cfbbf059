// Builds the per-worker instruction streams (see streams.hpp).  Pure host code.
#include "streams.hpp"

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>

namespace eicos
{
namespace
{
// ------------------------------------------------------------------ machine programs (machine.hpp)

// ---- forward sweep  xw = L^-1 P rhs, rows of L in dot form in ascending column order (the summation
// order of Eigen's column-oriented forward substitution).  L is stored once, column-major; the row-order
// walk is just the order of the load list.  Row i:  x_i = rhs[pinv i] - sum_k L(i,k) x_k, one operation per
// entry, chained through the partial sum.  Selectors: 1 = right-hand side (KKT order), 3 = xw (out vector).
void build_forward(const Symbolic &S, const Layout &L, MProgram &P)
{
    ivec xval(S.N, -1);
    for (int i = 0; i < S.N; i++)
    {
        const int t0 = S.Lr.p[i], cnt = S.Lr.p[i + 1] - t0;
        if (cnt == 0)
        { // a row of L without entries: xw_i IS the right-hand side entry (rhs - 0 * 0 leaves every bit, -0 included).
          // No operation and no copy: whoever reads xw_i - later rows here, the backward sweep's first product -
          // reads the right-hand side instead (MPC02: 3 497 of 5 991 rows, 22 % of the sweep's operations).
            xval[i] = P.new_value(1, S.pinv[i]);
            continue;
        }
        int prev = -1;
        for (int q = 0; q < cnt; q++)
        {
            MOp op;
            op.c = q == 0 ? MSrc::load(1, S.pinv[i]) : MSrc::value(prev);
            op.a = MSrc::load(0, L.Lx + S.Lr.v[t0 + q]);
            op.b = MSrc::value(xval[S.Lr.j[t0 + q]]);
            op.dst = prev = P.new_value(3, i);
            if (q == cnt - 1)
            {
                op.flags |= MF_OUT;
                op.out_row = i;
            }
            P.ops.push_back(op);
        }
        xval[i] = prev;
    }
}

// ---- backward sweep  out = P' L^-T D^-1 xw, columns in reverse elimination order (dot form, Eigen's
// order); results land in KKT order.  Column k:  v = (1/d_k) xw_k;  v -= L(i,k) x_i for the rows of the
// column; out[pinv k] = v.  accumulate: the last operation of a column also hands the row of the accumulated
// solution to the finish functor (x += v for the instances that continue refining).
// Selectors: 1 = output vector, 2 = accumulated solution, 3 = xw, 4 = the right-hand side of the forward sweep.
void build_backward(const Symbolic &S, const Layout &L, MProgram &P, bool accumulate)
{
    ivec xval(S.N, -1);
    for (int k = S.N - 1; k >= 0; k--)
    {
        const int o = S.pinv[k], u0 = S.Lp[k], cnt = S.Lp[k + 1] - u0;
        int prev = -1;
        for (int q = 0; q <= cnt; q++)
        {
            MOp op;
            if (q == 0)
            { // Eigen: diag.inverse() * x  (a product: 1/d * xw + (-0))
                op.c = MSrc::negzero();
                op.a = MSrc::load(0, L.Dinv + k);
                // (a row of L without entries has no xw: the forward sweep left its right-hand side entry in place)
                op.b = S.Lr.p[k + 1] == S.Lr.p[k] ? MSrc::load(4, S.pinv[k]) : MSrc::load(3, k);
                op.flags |= MF_POS;
            }
            else
            {
                op.c = MSrc::value(prev);
                op.a = MSrc::load(0, L.Lx + u0 + q - 1);
                op.b = MSrc::value(xval[S.Li[u0 + q - 1]]);
            }
            op.dst = prev = P.new_value(1, o);
            if (q == cnt)
            {
                op.flags |= MF_OUT;
                op.out_row = o;
                if (accumulate)
                {
                    op.flags |= MF_FIN | (FIN_ACC << MF_KIND_SHIFT);
                    op.x3 = MSrc::load(2, o);
                }
            }
            P.ops.push_back(op);
        }
        xval[k] = prev;
    }
}

// ---- the rows of the KKT mat-vecs: x, y and z rows (the two expansion slots of every second-order cone
// excepted) in elimination order; the entries of a row keep the order of the CSC / CSR data (G entries
// before A entries in an x row).
struct MatvecRows
{
    ivec order;
    std::vector<std::vector<std::pair<int, double>>> ent; // per K row: (K column, shared coefficient)
    std::vector<ivec> crow;                                // ... and the workspace row of the per-instance coefficient (pim)
};
void matvec_rows(const Symbolic &S, const Layout &L, MatvecRows &R)
{
    const int n = S.n, p = S.p, zb = S.n + S.p;
    for (int r = 0; r < zb; r++)
        R.order.push_back(r);
    for (int i = 0; i < S.m; i++)
        R.order.push_back(zb + S.zk[i]); // K-space row of z entry i (expanded index)
    std::sort(R.order.begin(), R.order.end(), [&](int a, int b) { return S.P[a] < S.P[b]; });
    R.ent.assign(S.N, {});
    R.crow.assign(S.N, {});
    for (int j = 0; j < n; j++)
    {
        for (int k = S.G.p[j]; k < S.G.p[j + 1]; k++)
        {
            R.ent[j].push_back({zb + S.zk[S.G.i[k]], S.G.x[k]});
            R.crow[j].push_back(L.Gx + k);
        }
        for (int k = S.A.p[j]; k < S.A.p[j + 1]; k++)
        {
            R.ent[j].push_back({n + S.A.i[k], S.A.x[k]});
            R.crow[j].push_back(L.Ax + k);
        }
    }
    for (int i = 0; i < p; i++)
        for (int t = S.Ar.p[i]; t < S.Ar.p[i + 1]; t++)
        {
            R.ent[n + i].push_back({S.Ar.j[t], S.A.x[S.Ar.v[t]]});
            R.crow[n + i].push_back(L.Ax + S.Ar.v[t]);
        }
    for (int i = 0; i < S.m; i++)
        for (int t = S.Gr.p[i]; t < S.Gr.p[i + 1]; t++)
        {
            R.ent[zb + S.zk[i]].push_back({S.Gr.j[t], S.G.x[S.Gr.v[t]]});
            R.crow[zb + S.zk[i]].push_back(L.Gx + S.Gr.v[t]);
        }
}
// rows [first, last) of part `part` of `parts`: contiguous stretches of the elimination order with about the same
// number of entries (the mat-vec rows are independent of each other, so parts can run on different warps)
void matvec_part(const MatvecRows &R, int part, int parts, size_t &first, size_t &last)
{
    std::vector<size_t> weight(R.order.size() + 1, 0);
    for (size_t k = 0; k < R.order.size(); k++)
        weight[k + 1] = weight[k] + 2 + R.ent[R.order[k]].size();
    const auto cut = [&](int q) {
        const size_t target = weight.back() * (size_t)q / (size_t)parts;
        return (size_t)(std::lower_bound(weight.begin(), weight.end(), target) - weight.begin());
    };
    first = part == 0 ? 0 : std::min(cut(part), R.order.size());
    last = part + 1 == parts ? R.order.size() : std::min(cut(part + 1), R.order.size());
}

int mv_kind(const Symbolic &S, int r)
{
    const int zb = S.n + S.p;
    return r < S.n ? MV_X : (r < zb ? MV_Y : (r < zb + S.l ? MV_Z : MV_ZC));
}

// ---- residual of the iterative refinement  e = rhs - Ktrue x  (src/eicos.cpp:1511-1576), LP part:
//   row r:  v = rhs_r - sum_k coefficient_k x_k;  x rows: v -= delta x_r;  y rows: v += delta x_r;
//   LP z rows: v += delta x_r, v -= (-w_r^2) x_r (the scalings row holds -w^2, -1 while the scalings are the identity);
//   rows of second-order cones stop after the sum (the cone block is applied cone by cone afterwards).
// The vector x is an external value per row: gathered once, parked in a slot while it has further uses.
// Selectors: 1 = rhs, 2 = x, 3 = LP scalings, 4 = e (out vector).
void build_matvec(const Symbolic &S, const Layout &L, MProgram &P, int &mv_rows, bool pim, int part = 0, int parts = 1)
{
    MatvecRows R;
    matvec_rows(S, L, R);
    size_t first, last;
    matvec_part(R, part, parts, first, last);
    const int zb = S.n + S.p;
    const double delta = Settings::deltastat;
    P.keep_loads = true;
    ivec xv(S.N);
    for (int c = 0; c < S.N; c++)
        xv[c] = P.new_value(2, c);
    // A row is one dependent chain (Eigen's summation order).  A row much longer than the others (a dense column of
    // G' or A': MPC02 has one of 1 497 entries) would run alone at the end of the program, one operation per bundle;
    // its operations are spread over the whole program instead, so that the scheduler's look-ahead window always has
    // the chain's next link next to short rows.  The order inside every row - and so every result - is unchanged.
    std::vector<std::vector<MOp>> chains; // long rows
    std::vector<MOp> shorts;              // operations of all other rows, row after row
    size_t total_ops = 0;
    for (size_t at = first; at < last; at++)
        total_ops += R.ent[R.order[at]].size() + 2;
    const size_t long_row = std::max<size_t>(128, total_ops / 64);
    for (size_t at = first; at < last; at++)
    {
        const int r = R.order[at];
        const int kind = mv_kind(S, r);
        std::vector<MOp> row;
        for (size_t q = 0; q < R.ent[r].size(); q++)
        {
            MOp op;
            op.a = pim ? MSrc::load(0, R.crow[r][q]) : MSrc::constant(R.ent[r][q].second);
            op.b = MSrc::value(xv[R.ent[r][q].first]);
            row.push_back(op);
        }
        if (kind != MV_ZC)
        {
            MOp op;
            op.a = MSrc::constant(kind == MV_X ? delta : -delta);
            op.b = MSrc::value(xv[r]);
            row.push_back(op);
        }
        if (kind == MV_Z)
        {
            MOp op;
            op.a = MSrc::load(3, r - zb); // (the row holds -w^2, and -1 during the initial solves: no MF_POS, no MF_AONE)
            op.b = MSrc::value(xv[r]);
            row.push_back(op);
        }
        if (row.empty())
            row.push_back(MOp());
        int prev = -1;
        for (size_t q = 0; q < row.size(); q++)
        {
            MOp &op = row[q];
            op.c = q == 0 ? MSrc::load(1, r) : MSrc::value(prev);
            op.dst = prev = P.new_value(4, r);
            if (q + 1 == row.size())
            {
                op.flags |= MF_OUT;
                op.out_row = r;
                if (kind != MV_ZC)
                    op.flags |= MF_FIN | (FIN_ABSMAX << MF_KIND_SHIFT);
            }
        }
        if (row.size() >= long_row)
            chains.push_back(row);
        else
            shorts.insert(shorts.end(), row.begin(), row.end());
    }
    {
        size_t nlong = 0;
        for (const auto &c : chains)
            nlong += c.size();
        // one link of a long row after every `gap` operations of short rows (the long rows one after the other)
        const size_t gap = nlong ? std::max<size_t>(1, shorts.size() / nlong) : 0;
        size_t ci = 0, cq = 0, since = 0;
        const auto link = [&]() {
            if (ci == chains.size())
                return false;
            P.ops.push_back(chains[ci][cq]);
            if (++cq == chains[ci].size())
                ci++, cq = 0;
            return true;
        };
        for (const MOp &op : shorts)
        {
            P.ops.push_back(op);
            if (gap && ++since >= gap)
            {
                since = 0;
                link();
            }
        }
        while (link())
        {
        }
    }
    mv_rows = (int)R.order.size();
}

// ---- computeResiduals (src/eicos.cpp:643-689): rx = -G'z - A'y - tau c, ry = A x - tau b, rz = s + G x - tau h
// and the sums of updateStatistics, which the finish functor (tile_program.hpp: ResidFin) collects:
//   x row:  v = 0 - sum;  PRE (hresx += v^2);  v -= tau c_j  -> out;  FIN (c_j, x_j)
//   y row:  v = 0 + sum;  PRE;  v -= tau b_i -> out;  FIN (b_i, y_i)
//   z row:  v = s_i (FIRST: s_i, z_i);  v += sum;  PRE;  v -= tau h_i -> out;  FIN (h_i, z_i)
// Selectors: 1 = [c | b | h], 2 = [x | y | z], 3 = s, 4 = r (out vector), 5 = scalar rows.
void build_resid(const Symbolic &S, const Layout &L, MProgram &P, bool pim, int part, int parts)
{
    MatvecRows R;
    matvec_rows(S, L, R);
    size_t first, last;
    matvec_part(R, part, parts, first, last);
    const int zb = S.n + S.p;
    P.keep_loads = true;
    ivec xv(S.N);
    for (int c = 0; c < S.N; c++)
        xv[c] = P.new_value(2, c);
    // tau: copied into a slot once (an A operand of every row)
    const int tau_home = P.new_value(5, S_TAU);
    const int tau = P.new_value(-1, 0);
    {
        MOp op;
        op.c = MSrc::value(tau_home);
        op.dst = tau;
        P.ops.push_back(op);
    }
    for (size_t at = first; at < last; at++)
    {
        const int r = R.order[at];
        const int kind = mv_kind(S, r);
        const bool zrow = kind >= MV_Z;
        const int pre = kind == MV_X ? RS_PRE_X : (kind == MV_Y ? RS_PRE_Y : RS_PRE_Z);
        const int fin = kind == MV_X ? RS_FIN_X : (kind == MV_Y ? RS_FIN_Y : RS_FIN_Z);
        const size_t cnt = R.ent[r].size();
        int prev = -1;
        if (zrow)
        { // v = s_i
            MOp op;
            op.c = MSrc::load(3, r - zb);
            op.x3 = MSrc::value(xv[r]);
            op.flags |= MF_FIN | ((cnt == 0 ? RS_FIRST_PRE_Z : RS_FIRST_Z) << MF_KIND_SHIFT);
            op.dst = prev = P.new_value(4, r);
            P.ops.push_back(op);
        }
        for (size_t q = 0; q < cnt; q++)
        {
            MOp op;
            op.c = prev >= 0 ? MSrc::value(prev) : MSrc::zero();
            op.a = pim ? MSrc::load(0, R.crow[r][q]) : MSrc::constant(R.ent[r][q].second);
            op.b = MSrc::value(xv[R.ent[r][q].first]);
            if (kind != MV_X)
                op.flags |= MF_POS;
            if (q + 1 == cnt)
                op.flags |= MF_FIN | (pre << MF_KIND_SHIFT);
            op.dst = prev = P.new_value(4, r);
            P.ops.push_back(op);
        }
        MOp op; // the tau term
        op.c = prev >= 0 ? MSrc::value(prev) : MSrc::zero();
        op.a = MSrc::value(tau);
        op.b = MSrc::load(1, r);
        op.x3 = MSrc::value(xv[r]);
        op.flags |= MF_OUT | MF_FIN | (fin << MF_KIND_SHIFT);
        op.out_row = r;
        op.dst = P.new_value(4, r);
        P.ops.push_back(op);
    }
}

// per K slot: row of the workspace holding the per-instance A / G value (per-instance-matrices mode), or -1
ivec ag_rows(const Symbolic &S, const Layout &L)
{
    ivec row(S.Ki.size(), -1);
    for (size_t k = 0; k < S.AGslot.size(); k++)
        row[S.AGslot[k]] = S.AGsrc[k] >= 0 ? L.Ax + S.AGsrc[k] : L.Gx + (-S.AGsrc[k] - 1);
    return row;
}

// ---- numeric factorisation (Eigen's ldlt.factorize, src/eicos.cpp:900,1164), right-looking in elimination
// order, as a machine program.  Every entry (i,j) of L and every pivot owns an accumulator that starts from
// its KKT value (shared constant, per-instance row of the scaling block / of the instance's matrices, or 0
// for fill) on its first touch.  Step k:
//     rd = 1 / acc(k,k)                              -> row Dinv + k   (the sweeps multiply by it, like Eigen's
//                                                       diag.inverse() * x; the finish functor flags rd = inf)
//     l_i = acc(i,k) * rd  for the rows i of column k -> row Lx + u    (column-major = the order the sweeps stream it in)
//     acc(i1,i2) -= l_i2 * acc(i1,k)  for every pair i1 >= i2 of the column   (the products Eigen's up-looking
//                                                       kernel forms, accumulated in ascending k)
// Accumulators are values of the program: they sit in slots between their updates; when the slots run out
// they wait in home rows of their own (Layout::acc, allocated only for patterns that need them; the D row
// for a pivot).  Untouched per-instance entries are
// external values (gathered when first used, kept while they have further uses); shared constants ride in
// the records.  HBM traffic: the scaling block in, L and 1/D out.  Selector 0 only (absolute rows).
void build_factor(const Symbolic &S, const Layout &L, MProgram &P, bool pim)
{
    struct Init
    {
        int kind = 0; // 0 zero (fill), 1 constant, 2 row
        int vrow = -1;
        double c = 0.0;
    };
    const ivec agrow = pim ? ag_rows(S, L) : ivec(S.Ki.size(), -1);
    std::vector<Init> di(S.N), ei(S.nnzL);
    for (int j = 0; j < S.N; j++)
        for (int e = S.KLp[j]; e < S.KLp[j + 1]; e++)
        {
            const int slot = S.KLslot[e], vi = S.Kvidx[slot], pos = S.KLpos[e];
            Init &t = pos < 0 ? di[j] : ei[S.Lp[j] + pos];
            if (vi >= 0 || agrow[slot] >= 0)
            {
                t.kind = 2;
                t.vrow = vi >= 0 ? L.V + vi : agrow[slot];
            }
            else
            {
                t.kind = 1;
                t.c = S.Kshared[slot];
            }
        }
    P.keep_loads = true;
    ivec dcur(S.N, -1), ecur(S.nnzL, -1); // current value of a touched accumulator
    const auto start = [&](const Init &t) { // first-touch value of an accumulator, as a C operand
        return t.kind == 2 ? MSrc::load(0, t.vrow) : (t.kind == 1 ? MSrc::constant(t.c) : MSrc::zero());
    };
    std::vector<MSrc> av, lv;
    for (int k = 0; k < S.N; k++)
    {
        const int u0 = S.Lp[k], cnt = S.Lp[k + 1] - u0;
        int rd;
        {
            MOp op;
            op.c = dcur[k] >= 0 ? MSrc::value(dcur[k]) : start(di[k]);
            op.flags |= MF_RECIP | MF_OUT | MF_FIN | (FIN_PIVOT << MF_KIND_SHIFT);
            op.out_row = L.Dinv + k;
            op.dst = rd = P.new_value(0, L.Dinv + k);
            P.ops.push_back(op);
        }
        if (cnt == 0)
            continue;
        // the finished accumulators of the column: a value, a shared constant, or an untouched per-instance row
        av.assign(cnt, MSrc());
        lv.assign(cnt, MSrc());
        for (int e = 0; e < cnt; e++)
        {
            const Init &t = ei[u0 + e];
            if (ecur[u0 + e] >= 0)
                av[e] = MSrc::value(ecur[u0 + e]);
            else if (t.kind == 2)
                av[e] = MSrc::value(P.new_value(0, t.vrow)); // external: lives in its row
            else
                av[e] = MSrc::constant(t.kind == 1 ? t.c : 0.0);
            MOp op; // l = a * rd  (a product: + (-0))
            op.c = MSrc::negzero();
            op.flags |= MF_POS | MF_OUT;
            op.out_row = L.Lx + u0 + e;
            if (av[e].kind == MS_CONST)
            {
                op.a = av[e];
                op.b = MSrc::value(rd);
            }
            else
            {
                op.a = MSrc::value(rd);
                op.b = av[e];
            }
            op.dst = P.new_value(0, L.Lx + u0 + e);
            lv[e] = MSrc::value(op.dst);
            P.ops.push_back(op);
        }
        for (int e1 = 0; e1 < cnt; e1++)
        {
            const int i1 = S.Li[u0 + e1];
            for (int e2 = 0; e2 <= e1; e2++)
            {
                int *cur;
                const Init *init;
                int home;
                if (e2 < e1)
                {
                    const int i2 = S.Li[u0 + e2];
                    const int *b = S.Li.data() + S.Lp[i2], *e = S.Li.data() + S.Lp[i2 + 1];
                    const int *f = std::lower_bound(b, e, i1);
                    if (f == e || *f != i1)
                        throw std::logic_error("factor program: update outside the pattern of L");
                    const int ut = (int)(f - S.Li.data());
                    cur = &ecur[ut];
                    init = &ei[ut];
                    // (not the row of the L entry: the finished accumulator is still read after l has been written there)
                    home = L.acc >= 0 ? L.acc + ut : -1;
                }
                else
                {
                    cur = &dcur[i1];
                    init = &di[i1];
                    home = L.D + i1;
                }
                MOp op; // acc -= l(i2) * a(i1)
                op.c = *cur >= 0 ? MSrc::value(*cur) : start(*init);
                if (op.c.kind == MS_CONST && av[e1].kind == MS_CONST)
                { // a record carries one constant: the accumulator starts as a value of its own
                    MOp mv;
                    mv.c = op.c;
                    mv.dst = P.new_value(home >= 0 ? 0 : -1, std::max(home, 0));
                    P.ops.push_back(mv);
                    op.c = MSrc::value(mv.dst);
                }
                if (av[e1].kind == MS_CONST)
                {
                    op.a = av[e1];
                    op.b = lv[e2];
                }
                else
                {
                    op.a = lv[e2];
                    op.b = av[e1];
                }
                op.dst = *cur = P.new_value(home >= 0 ? 0 : -1, std::max(home, 0));
                P.ops.push_back(op);
            }
        }
    }
}
} // namespace

void build_streams(const Symbolic &S, const Layout &L, int W, int max_sw_slots, int max_fa_slots, HostStreams &H, bool pim)
{
    constexpr int MACHINE_TUNE_SLOTS = 16, MACHINE_TUNE_FA_SLOTS = 24; // = the engine's default budgets (engine.cu: MAX_SW_SLOTS, MAX_FA_SLOTS)
    H = HostStreams();
    H.workers = W;
    H.sw_budget = max_sw_slots;
    H.fa_budget = max_fa_slots;
    for (int k = 0; k < S.N; k++)
        for (int u = S.Lp[k]; u + 1 < S.Lp[k + 1]; u++)
            if (S.Li[u] >= S.Li[u + 1])
                throw std::logic_error("columns of L must have ascending rows");
    // The factor program has one word per Schur update.  Patterns that fill into large dense fronts
    // (hundreds of millions of updates) belong to a supernodal / dense-front path, which this engine
    // does not have yet: refuse instead of compiling a multi-gigabyte program.
    if (S.fma_count > MAX_FACTOR_UPDATES)
        throw std::runtime_error("pattern fills too much for the row-program factorisation (" + std::to_string(S.fma_count) +
                                 " Schur updates per factorisation; limit " + std::to_string(MAX_FACTOR_UPDATES) + ")");
    {
        const int slots = std::max(2, max_sw_slots);
        MProgram pf, pb, pbp, pm;
        build_forward(S, L, pf);
        build_backward(S, L, pb, true);
        build_backward(S, L, pbp, false);
        build_matvec(S, L, pm, H.mv_rows, pim);
        // deep = slots of the deeper-ring variants (k >= 1): those run few tiles per SM, where shared memory is free
        // (MPC02 with 32 instead of 16 slots: the mat-vec copies 18.0 k instead of 25.7 k rows per pass)
        const auto compile = [&](const char *name, const MProgram &p, MachineCode (&c)[M_VARIANTS], int budget, int tune, int deep = 0) {
            // (diagnostics) EICOS_SCHED_WINDOW_<name> pins the scheduler window of one program
            const std::string key = std::string("EICOS_SCHED_WINDOW_") + name;
            const char *v = std::getenv(key.c_str());
            machine_compile(p, budget, c[0], tune, variant_groups(0), v ? std::max(M_U, std::atoi(v)) : 0);
            for (int k = 1; k < M_VARIANTS; k++) // same order of operations, deeper ring
                machine_compile(p, deep > 0 ? deep : budget, c[k], tune, variant_groups(k), c[0].window);
            if (std::getenv("EICOS_DBG_PROGRAMS"))
                for (int k = 0; k < M_VARIANTS; k++)
                    std::fprintf(stderr, "machine %s[%d]: ops %lld nop %lld bundles %d loads %d far %lld pads %lld spills %lld slots %d window %d\n",
                                 name, k, c[k].nops, c[k].nnop, c[k].nbundles, c[k].nld, c[k].far, c[k].pads, c[k].spills, c[k].slot_rows,
                                 c[k].window);
        };
        // (a budget pinned from the environment - the tests starve the slots on purpose - holds for every variant)
        const int deep_slots = std::getenv("EICOS_MAX_SW_SLOTS") ? slots : 2 * slots;
        compile("fw", pf, H.fw, slots, MACHINE_TUNE_SLOTS, deep_slots);
        compile("bw", pb, H.bw, slots, MACHINE_TUNE_SLOTS, deep_slots);
        compile("bwp", pbp, H.bwp, slots, MACHINE_TUNE_SLOTS, deep_slots);
        compile("mv", pm, H.mv, slots, MACHINE_TUNE_SLOTS, deep_slots);
        // The rows of a mat-vec do not depend on each other: computeResiduals is compiled as M_MV_PARTS programs over
        // contiguous stretches of the rows (its fourteen sums are then combined in a fixed order, whichever way the
        // parts are run), and the refinement residual a second time in that form - for launches with few tiles, where
        // the parts run on different warps of the CTA, each with a machine of its own (shallow ring).
        for (int k = 0; k < M_MV_PARTS; k++)
        {
            MProgram pr, pw;
            int rows_unused = 0;
            build_resid(S, L, pr, pim, k, M_MV_PARTS);
            build_matvec(S, L, pw, rows_unused, pim, k, M_MV_PARTS);
            machine_compile(pr, slots, H.rs[k], MACHINE_TUNE_SLOTS, M_PART_GROUPS, 0);
            machine_compile(pw, slots, H.mvw[k], MACHINE_TUNE_SLOTS, M_PART_GROUPS, 0);
            if (std::getenv("EICOS_DBG_PROGRAMS"))
                for (auto nq : {std::make_pair("rs", &H.rs[k]), std::make_pair("mvw", &H.mvw[k])})
                    std::fprintf(stderr, "machine %s part %d: ops %lld nop %lld bundles %d loads %d far %lld pads %lld spills %lld slots %d window %d\n",
                                 nq.first, k, nq.second->nops, nq.second->nnop, nq.second->nbundles, nq.second->nld, nq.second->far,
                                 nq.second->pads, nq.second->spills, nq.second->slot_rows, nq.second->window);
        }
        // two-job forms: the same operations, every vector operand two rows wide
        {
            const int pslots = std::max(2, std::min(14, max_sw_slots * 14 / 16)); // (14 two-row slots: seven tiles per SM)
            H.pair_budget = pslots;
            const auto pair = [&](const char *name, MProgram &p, MachineCode &c, int window) {
                p.nr = 2; // (same order of the operations as the one-job form)
                machine_compile(p, pslots, c, 0, M_PAIR_GROUPS, window);
                if (std::getenv("EICOS_DBG_PROGRAMS"))
                    std::fprintf(stderr, "machine %s: ops %lld nop %lld bundles %d loads %d far %lld pads %lld spills %lld slots %d window %d\n", name,
                                 c.nops, c.nnop, c.nbundles, c.nld, c.far, c.pads, c.spills, c.slot_rows, c.window);
            };
            pair("fw2", pf, H.fw2, H.fw[0].window);
            pair("bw2", pb, H.bw2, H.bw[0].window);
            pair("bwp2", pbp, H.bwp2, H.bwp[0].window);
            pair("mv2", pm, H.mv2, H.mv[0].window);
        }
        MProgram pa;
        build_factor(S, L, pa, pim);
        compile("fa", pa, H.fa, std::max(2, max_fa_slots), MACHINE_TUNE_FA_SLOTS);
        for (int k = 0; k < M_VARIANTS; k++)
            for (const MachineCode *q : {&H.fw[k], &H.bw[k], &H.bwp[k], &H.mv[k]})
                H.sw_slots = std::max(H.sw_slots, q->slot_rows);
        for (int k = 0; k < M_MV_PARTS; k++)
            H.sw_slots = std::max(H.sw_slots, std::max(H.rs[k].slot_rows, H.mvw[k].slot_rows));
        H.sw_far = H.fw[0].far + H.bwp[0].far;
        H.fa_slots = std::max(H.fa[0].slot_rows, H.fa[1].slot_rows);
        H.fa_home = H.fa[0].spills + H.fa[0].far;
    }
}

void refresh_stream_values(const Symbolic &S, const Layout &L, HostStreams &H, bool pim)
{
    HostStreams fresh;
    build_streams(S, L, H.workers, std::max(H.sw_budget, 2), std::max(H.fa_budget, 2), fresh, pim);
    std::vector<std::pair<MachineCode *, MachineCode *>> pq = {{&H.mv2, &fresh.mv2}};
    for (int k = 0; k < M_VARIANTS; k++)
        for (auto q : {std::make_pair(&H.mv[k], &fresh.mv[k]), std::make_pair(&H.fa[k], &fresh.fa[k])})
            pq.push_back(q);
    for (int k = 0; k < M_MV_PARTS; k++)
        for (auto q : {std::make_pair(&H.rs[k], &fresh.rs[k]), std::make_pair(&H.mvw[k], &fresh.mvw[k])})
            pq.push_back(q);
    for (auto &q : pq)
    { // the mat-vec and factor programs carry the shared coefficients inline
        if (q.second->ops.size() != q.first->ops.size())
            throw std::logic_error("machine program changed shape on a value refresh");
        q.first->ops.swap(q.second->ops);
    }
}

} // namespace eicos
