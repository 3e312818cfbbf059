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
void align_chunk(ivec &s)
{
    while (s.size() % STREAM_CHUNK)
        s.push_back(0);
}
void align_chunk(dvec &s)
{
    while (s.size() % STREAM_CHUNK)
        s.push_back(0.0);
}
void pad_tail(ivec &s)
{
    align_chunk(s);
    s.insert(s.end(), STREAM_PAD, 0);
}
void pad_tail(dvec &s)
{
    align_chunk(s);
    s.insert(s.end(), STREAM_PAD, 0.0);
}

// Shared-memory slots for the live ranges of a slot program (linear scan: a value gets a slot when
// it is first touched and gives it back after its last use; when none is free it lives at home).
struct SlotPool
{
    ivec free_;
    int top = 0; // slots ever used
    long long home = 0;
    explicit SlotPool(int n)
    {
        for (int s = n - 1; s >= 0; s--)
            free_.push_back(s);
    }
    int take(int home_row)
    {
        if (free_.empty())
        {
            home++;
            return SLOT_HOME + home_row;
        }
        const int s = free_.back();
        free_.pop_back();
        top = std::max(top, s + 1);
        return s;
    }
    void give(int code)
    {
        if (code >= 0 && code < SLOT_HOME)
            free_.push_back(code);
    }
};

// Host model of the device FIFO: every pop appends the row to the load list and returns the ring row
// the device will find it in.
struct FifoSim
{
    ivec &ld;
    int npop = 0;
    explicit FifoSim(ivec &l) : ld(l) {}
    int pop(int base, int row)
    {
        if (row < 0 || row > LD_ROW_MASK)
            throw std::logic_error("load list: row out of range");
        ld.push_back((base << LD_BASE_SHIFT) | row);
        return npop++ % FIFO_ROWS;
    }
    // does a check point that makes `pops` pops, the first one being pop number `first`, enter a new group?
    static bool crosses(int first, int pops)
    {
        if (pops <= 0)
            return false;
        const int before = first == 0 ? 0 : (first - 1) / FIFO_GROUP + 1;
        return (first + pops - 1) / FIFO_GROUP + 1 > before;
    }
    // A value written to global memory when `prod` pops had been made may be loaded through the FIFO as
    // pop number j only if the group of j is issued afterwards (sync points come up to FIFO_GROUP - 1
    // pops early, the first FIFO_AHEAD groups are issued before the program starts).
    static bool far_safe(int prod, int j)
    {
        const int g = j / FIFO_GROUP;
        return g >= FIFO_AHEAD && prod <= (g - FIFO_AHEAD) * FIFO_GROUP - FIFO_GROUP;
    }
};

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
        int prev = -1;
        for (int q = 0; q < std::max(cnt, 1); q++)
        {
            MOp op;
            op.c = q == 0 ? MSrc::load(1, S.pinv[i]) : MSrc::value(prev);
            if (cnt > 0)
            {
                op.a = MSrc::load(0, L.Lx + S.Lr.v[t0 + q]);
                op.b = MSrc::value(xval[S.Lr.j[t0 + q]]);
            }
            op.dst = prev = P.new_value(3, i);
            if (q == std::max(cnt, 1) - 1)
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
// Selectors: 1 = output vector, 2 = accumulated solution, 3 = xw.
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
                op.b = MSrc::load(3, k);
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
int mv_kind(const Symbolic &S, int r)
{
    const int zb = S.n + S.p;
    return r < S.n ? MV_X : (r < zb ? MV_Y : (r < zb + S.l ? MV_Z : MV_ZC));
}

// ---- residual of the iterative refinement  e = rhs - Ktrue x  (src/eicos.cpp:1511-1576), LP part:
//   row r:  v = rhs_r - sum_k coefficient_k x_k;  x rows: v -= delta x_r;  y rows: v += delta x_r;
//   LP z rows: v += delta x_r, v += w_r^2 x_r (x_r alone while the scalings are the identity: MF_AONE);
//   rows of second-order cones stop after the sum (the cone block is applied cone by cone afterwards).
// The vector x is an external value per row: gathered once, parked in a slot while it has further uses.
// Selectors: 1 = rhs, 2 = x, 3 = LP scalings, 4 = e (out vector).
void build_matvec(const Symbolic &S, const Layout &L, MProgram &P, int &mv_rows, bool pim)
{
    MatvecRows R;
    matvec_rows(S, L, R);
    const int zb = S.n + S.p;
    const double delta = Settings::deltastat;
    P.keep_loads = true;
    ivec xv(S.N);
    for (int c = 0; c < S.N; c++)
        xv[c] = P.new_value(2, c);
    for (int r : R.order)
    {
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
            op.a = MSrc::load(3, r - zb);
            op.b = MSrc::value(xv[r]);
            op.flags |= MF_POS | MF_AONE;
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
            P.ops.push_back(op);
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
void build_resid(const Symbolic &S, const Layout &L, MProgram &P, bool pim)
{
    MatvecRows R;
    matvec_rows(S, L, R);
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
    for (int r : R.order)
    {
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

// ---- numeric factorisation, right-looking in elimination order.  Every entry (i,j) of L and every
// pivot owns an accumulator that starts from the KKT value (shared constant, per-instance scaling
// value, or 0 for fill) on its first touch.  Step k: d = acc(k,k); for the rows i of column k
// a_i = acc(i,k), l_i = a_i / d (stored, column-major = the order both sweeps stream it in); then
// for every pair i1 >= i2 of the column  acc(i1,i2) -= l_i2 * a_i1  (the products Eigen's
// up-looking kernel forms, accumulated in ascending k).
//   ops:   [src_d, cnt, cnt x src, cnt(cnt+1)/2 x target]   constants in fa_val, V rows in the load list
void build_factor(const Symbolic &S, const Layout &L, int max_slots, HostStreams &H, bool pim)
{
    struct Init
    {
        int kind = OPK_ZERO, vrow = -1;
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
                t.kind = OPK_FIFO;
                t.vrow = vi >= 0 ? L.V + vi : agrow[slot];
            }
            else
            {
                t.kind = OPK_CONST;
                t.c = S.Kshared[slot];
            }
        }
    SlotPool pool(max_slots);
    ivec dcode(S.N, -1), ecode(S.nnzL, -1);
    const auto source = [&](int code, const Init &t) {
        if (code >= 0)
            H.fa.push_back(code);
        else if (t.kind == OPK_FIFO)
        {
            H.fa.push_back(SRC_FIFO);
            H.fa_ld.push_back(t.vrow);
        }
        else if (t.kind == OPK_CONST)
        {
            H.fa.push_back(SRC_CONST);
            H.fa_val.push_back(t.c);
        }
        else
            H.fa.push_back(SRC_ZERO);
    };
    const auto target = [&](int &code, const Init &t, int home_row) {
        if (code >= 0)
        {
            H.fa.push_back(code);
            return;
        }
        code = pool.take(home_row);
        H.fa.push_back(code | (t.kind << OPK_SHIFT));
        if (t.kind == OPK_FIFO)
            H.fa_ld.push_back(t.vrow);
        else if (t.kind == OPK_CONST)
            H.fa_val.push_back(t.c);
    };
    for (int k = 0; k < S.N; k++)
    {
        const int u0 = S.Lp[k], cnt = S.Lp[k + 1] - u0;
        source(dcode[k], di[k]);
        H.fa.push_back(cnt);
        for (int e = 0; e < cnt; e++)
            source(ecode[u0 + e], ei[u0 + e]);
        for (int e1 = 0; e1 < cnt; e1++)
        {
            const int i1 = S.Li[u0 + e1];
            for (int e2 = 0; e2 < e1; e2++)
            {
                const int i2 = S.Li[u0 + e2];
                const int *b = S.Li.data() + S.Lp[i2], *e = S.Li.data() + S.Lp[i2 + 1];
                const int *f = std::lower_bound(b, e, i1);
                if (f == e || *f != i1)
                    throw std::logic_error("factor program: update outside the pattern of L");
                const int ut = (int)(f - S.Li.data());
                target(ecode[ut], ei[ut], L.Lx + ut);
            }
            target(dcode[i1], di[i1], L.D + i1);
        }
        pool.give(dcode[k]);
        for (int e = 0; e < cnt; e++)
            pool.give(ecode[u0 + e]);
    }
    H.fa_nld = (int)H.fa_ld.size();
    H.fa_slots = pool.top;
    H.fa_home = pool.home;
    pad_tail(H.fa);
    pad_tail(H.fa_ld);
    pad_tail(H.fa_val);
}
// ---- numeric factorisation in record form (streams.hpp); same right-looking algorithm and the
// same arithmetic as build_factor, but every operand is a shared-memory row known to the host.
// Returns false (leaving H untouched) when the pattern does not qualify.
bool build_factor_fast(const Symbolic &S, const Layout &L, int max_slots, HostStreams &H, bool pim)
{
    if (S.maxcol > FA_FAST_COL)
        return false;
    struct Init
    {
        int kind = FA_ZERO, vrow = -1;
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
                t.kind = FA_ROW;
                t.vrow = vi >= 0 ? L.V + vi : agrow[slot];
            }
            else
            {
                t.kind = FA_CONST;
                t.c = S.Kshared[slot];
            }
        }
    ivec ops, ld;
    dvec val;
    SlotPool pool(max_slots);
    FifoSim F(ld);
    ivec dcode(S.N, -1), ecode(S.nnzL, -1); // slot of a touched accumulator
    const int slot0 = FIFO_ROWS;
    const auto source = [&](int code, const Init &t) {
        if (code >= 0)
            return slot0 + code;
        if (t.kind == FA_ROW)
            return F.pop(0, t.vrow);
        if (t.kind == FA_CONST)
        {
            val.push_back(t.c);
            return FA_CONST << FA_KIND_SHIFT;
        }
        return FA_ZERO << FA_KIND_SHIFT;
    };
    bool ok = true;
    const auto target = [&](int &code, const Init &t, int home_row) {
        if (code >= 0)
            return (slot0 + code) | ((slot0 + code) << 8);
        code = pool.take(home_row);
        if (code >= SLOT_HOME)
        {
            ok = false;
            return 0;
        }
        const int row = slot0 + code;
        if (t.kind == FA_ROW)
            return row | (F.pop(0, t.vrow) << 8);
        if (t.kind == FA_CONST)
        {
            val.push_back(t.c);
            return row | (FA_CONST << FA_KIND_SHIFT);
        }
        return row | (FA_ZERO << FA_KIND_SHIFT);
    };
    // a record = 4 words whose pops start at `first`
    const auto close_record = [&](size_t at, int first) {
        if (FifoSim::crosses(first, F.npop - first))
            ops[at] |= FA_SYNC;
    };
    for (int k = 0; k < S.N && ok; k++)
    {
        const int u0 = S.Lp[k], cnt = S.Lp[k + 1] - u0;
        int first = F.npop;
        size_t at = ops.size();
        ops.push_back(source(dcode[k], di[k]));
        ops.push_back(cnt);
        for (int e = 0; e < 2; e++)
            ops.push_back(e < cnt ? source(ecode[u0 + e], ei[u0 + e]) : 0);
        close_record(at, first);
        if (cnt > 2)
        {
            first = F.npop;
            at = ops.size();
            for (int e = 2; e < 4; e++)
                ops.push_back(e < cnt ? source(ecode[u0 + e], ei[u0 + e]) : 0);
            ops.push_back(0);
            ops.push_back(0);
            close_record(at, first);
        }
        int inrec = 0;
        for (int e1 = 0; e1 < cnt && ok; e1++)
        {
            const int i1 = S.Li[u0 + e1];
            for (int e2 = 0; e2 <= e1 && ok; e2++)
            {
                if (inrec == 0)
                {
                    first = F.npop;
                    at = ops.size();
                }
                if (e2 < e1)
                {
                    const int i2 = S.Li[u0 + e2];
                    const int *b = S.Li.data() + S.Lp[i2], *e = S.Li.data() + S.Lp[i2 + 1];
                    const int *f = std::lower_bound(b, e, i1);
                    if (f == e || *f != i1)
                        throw std::logic_error("factor program: update outside the pattern of L");
                    const int ut = (int)(f - S.Li.data());
                    ops.push_back(target(ecode[ut], ei[ut], L.Lx + ut));
                }
                else
                    ops.push_back(target(dcode[i1], di[i1], L.D + i1));
                if (++inrec == 4)
                {
                    close_record(at, first);
                    inrec = 0;
                }
            }
        }
        if (inrec > 0)
        {
            while (inrec++ < 4)
                ops.push_back(0);
            close_record(at, first);
        }
        pool.give(dcode[k]);
        for (int e = 0; e < cnt; e++)
            pool.give(ecode[u0 + e]);
    }
    if (!ok || slot0 + pool.top > 255)
        return false;
    H.fa.swap(ops);
    H.fa_ld.swap(ld);
    H.fa_val.swap(val);
    H.fa_nld = (int)H.fa_ld.size();
    H.fa_slots = pool.top;
    H.fa_home = 0;
    H.fa_fast = 1;
    pad_tail(H.fa);
    pad_tail(H.fa_ld);
    pad_tail(H.fa_val);
    return true;
}
} // namespace

void build_streams(const Symbolic &S, const Layout &L, int W, int max_sw_slots, int max_fa_slots, HostStreams &H, bool pim)
{
    constexpr int MACHINE_TUNE_SLOTS = 24; // = the engine's default budget (engine.cu: MAX_SW_SLOTS)
    H = HostStreams();
    H.workers = W;
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
        MProgram pf, pb, pbp, pm, pr;
        build_forward(S, L, pf);
        build_backward(S, L, pb, true);
        build_backward(S, L, pbp, false);
        build_matvec(S, L, pm, H.mv_rows, pim);
        build_resid(S, L, pr, pim);
        const auto compile = [&](const char *name, const MProgram &p, MachineCode &c) {
            // (diagnostics) EICOS_SCHED_WINDOW_<name> pins the scheduler window of one program
            const std::string key = std::string("EICOS_SCHED_WINDOW_") + name;
            if (const char *v = std::getenv(key.c_str()))
            {
                setenv("EICOS_SCHED_WINDOW", v, 1);
                machine_compile(p, slots, c, MACHINE_TUNE_SLOTS);
                unsetenv("EICOS_SCHED_WINDOW");
            }
            else
                machine_compile(p, slots, c, MACHINE_TUNE_SLOTS);
        };
        compile("fw", pf, H.fw);
        compile("bw", pb, H.bw);
        compile("bwp", pbp, H.bwp);
        compile("mv", pm, H.mv);
        compile("rs", pr, H.rs);
        for (const MachineCode *q : {&H.fw, &H.bw, &H.bwp, &H.mv, &H.rs})
            H.sw_slots = std::max(H.sw_slots, q->slot_rows);
        H.sw_far = H.fw.far + H.bwp.far;
        if (std::getenv("EICOS_DBG_PROGRAMS"))
            for (auto nq : {std::make_pair("fw", &H.fw), std::make_pair("bw", &H.bw), std::make_pair("bwp", &H.bwp),
                            std::make_pair("mv", &H.mv), std::make_pair("rs", &H.rs)})
                std::fprintf(stderr, "machine %s: ops %lld nop %lld bundles %d loads %d far %lld pads %lld spills %lld slots %d\n", nq.first,
                             nq.second->nops, nq.second->nnop, nq.second->nbundles, nq.second->nld, nq.second->far, nq.second->pads,
                             nq.second->spills, nq.second->slot_rows, nq.second->window);
    }
    if (!build_factor_fast(S, L, max_fa_slots, H, pim))
        build_factor(S, L, max_fa_slots, H, pim);
}

void refresh_stream_values(const Symbolic &S, const Layout &L, HostStreams &H, bool pim)
{
    HostStreams fresh;
    build_streams(S, L, H.workers, std::max(H.sw_slots, 1), std::max(H.fa_slots, 1), fresh, pim);
    H.fa_val.swap(fresh.fa_val);
    for (auto pq : {std::make_pair(&H.mv, &fresh.mv), std::make_pair(&H.rs, &fresh.rs)})
    { // the mat-vec programs carry the shared coefficients inline
        if (pq.second->ops.size() != pq.first->ops.size())
            throw std::logic_error("mat-vec program changed shape on a value refresh");
        pq.first->ops.swap(pq.second->ops);
    }
}

} // namespace eicos
