// Builds the per-worker instruction streams (see streams.hpp).  Pure host code.
#include "streams.hpp"

#include <algorithm>
#include <cstdint>
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

// Shared-memory slots for the values of a sweep, Belady style: a value gets a slot when it is
// produced; when none is free the live value whose next use is furthest away loses its slot (it is
// still at home in global memory).  Use times are step numbers of the sweep.
struct SlotCache
{
    const std::vector<ivec> &uses;
    ivec holder, slot_of, ptr;
    int top = 0;
    SlotCache(int slots, const std::vector<ivec> &u) : uses(u), holder(slots, -1), slot_of(u.size(), -1), ptr(u.size(), 0) {}
    int next_use(int v) const { return ptr[v] < (int)uses[v].size() ? uses[v][ptr[v]] : INT32_MAX; }
    void used(int v)
    {
        ptr[v]++;
        if (ptr[v] == (int)uses[v].size() && slot_of[v] >= 0)
        {
            holder[slot_of[v]] = -1;
            slot_of[v] = -1;
        }
    }
    int alloc(int v)
    {
        if (next_use(v) == INT32_MAX)
            return -1;
        int s = -1;
        for (int q = 0; q < (int)holder.size() && s < 0; q++)
            if (holder[q] < 0)
                s = q;
        if (s < 0)
        {
            int far = -1;
            for (int q = 0; q < (int)holder.size(); q++)
                if (far < 0 || next_use(holder[q]) > next_use(holder[far]))
                    far = q;
            if (far < 0 || next_use(holder[far]) <= next_use(v))
                return -1;
            slot_of[holder[far]] = -1;
            s = far;
        }
        holder[s] = v;
        slot_of[v] = s;
        top = std::max(top, s + 1);
        return s;
    }
};

// Emits the pair words of one row behind its `hdr` header words (already pushed, the first at
// w0) and sets the sync flags: one check point per 16-byte record.  pair(q) appends nothing itself;
// it pops the L value, resolves the gathered value and returns the pair word without flags.
template <class Pair>
void emit_pairs(ivec &ops, FifoSim &F, size_t w0, int hdr, int first_pop, int cnt, Pair pair)
{
    int q = 0;
    for (int r = hdr; r < 4; r++)
        ops.push_back(q < cnt ? pair(q++) : SW_PAD_PAIR);
    if (FifoSim::crosses(first_pop, F.npop - first_pop))
        ops[w0] |= SW_SYNC_HDR;
    while (q < cnt)
    {
        const int first = F.npop;
        const size_t at = ops.size();
        for (int r = 0; r < 4; r++)
            ops.push_back(q < cnt ? pair(q++) : SW_PAD_PAIR);
        if (FifoSim::crosses(first, F.npop - first))
            ops[at] |= SW_SYNC_PAIR;
    }
}

int pair_word(int lrow, int opnd)
{
    if (opnd < 0 || opnd >= (1 << (31 - SW_OPND_SHIFT)))
        throw std::logic_error("sweep program: operand out of range");
    return lrow | (opnd << SW_OPND_SHIFT);
}

// ---- forward sweep  xw = L^-1 P rhs, rows of L in elimination order, dot form in ascending column
// order (the summation order of Eigen's column-oriented forward substitution).  L is stored once,
// column-major; the row-order walk is just the order of the load list.
//   row i: [cnt | sync, keep | rhs ring row << 8] cnt x pair
void build_forward(const Symbolic &S, const Layout &L, int max_slots, HostStreams &H)
{
    std::vector<ivec> uses(S.N);
    for (int k = 0; k < S.N; k++)
        uses[k].assign(S.Li.begin() + S.Lp[k], S.Li.begin() + S.Lp[k + 1]);
    SlotCache cache(max_slots, uses);
    FifoSim F(H.fw_ld);
    ivec prod(S.N, 0);
    for (int i = 0; i < S.N; i++)
    {
        const int cnt = S.Lr.p[i + 1] - S.Lr.p[i];
        const int first = F.npop;
        const int rhs_row = F.pop(1, S.pinv[i]);
        const size_t w1 = H.fw.size() + 1;
        H.fw.push_back(cnt);
        H.fw.push_back(0);
        emit_pairs(H.fw, F, w1 - 1, 2, first, cnt, [&](int q) {
            const int t = S.Lr.p[i] + q, k = S.Lr.j[t];
            const int lrow = F.pop(0, L.Lx + S.Lr.v[t]);
            int opnd;
            if (cache.slot_of[k] >= 0)
                opnd = SW_SLOT0 + cache.slot_of[k];
            else if (FifoSim::far_safe(prod[k], F.npop))
            {
                opnd = F.pop(3, k); // the work vector xw of the running solve
                H.sw_far++;
            }
            else
            {
                opnd = SW_DIRECT + k; // relative to xw
                H.sw_direct++;
            }
            cache.used(k);
            return pair_word(lrow, opnd);
        });
        const int s = cache.alloc(i);
        H.fw[w1] = (s >= 0 ? SW_SLOT0 + s : SW_NO_KEEP) | (rhs_row << 8);
        prod[i] = F.npop;
    }
    H.fw_nld = (int)H.fw_ld.size();
    H.sw_slots = std::max(H.sw_slots, cache.top);
    pad_tail(H.fw);
    pad_tail(H.fw_ld);
}

// ---- backward sweep  out = P' L^-T D^-1 xw, columns in reverse elimination order (dot form, Eigen's
// order); results land in KKT order.  The home of a finished entry is its output row.
//   column k: [cnt | sync, keep | 1/d ring row << 8 | xw ring row << 16 | accumulated-solution ring row << 24, out row] cnt x pair
// accumulate: the program also loads the row of the solution it adds its result to (refinement
// rounds); the plain program of the first solve leaves those loads out.
void build_backward(const Symbolic &S, const Layout &L, int max_slots, HostStreams &H, bool accumulate)
{
    ivec &ops = accumulate ? H.bw : H.bwp, &ld = accumulate ? H.bw_ld : H.bwp_ld;
    std::vector<ivec> uses(S.N); // value i is used by the columns of row i, latest column first
    for (int i = 0; i < S.N; i++)
        for (int t = S.Lr.p[i + 1] - 1; t >= S.Lr.p[i]; t--)
            uses[i].push_back(S.N - 1 - S.Lr.j[t]);
    for (const ivec &u : uses)
        if (!std::is_sorted(u.begin(), u.end()))
            throw std::logic_error("rows of L must have ascending columns");
    SlotCache cache(max_slots, uses);
    FifoSim F(ld);
    ivec prod(S.N, 0);
    for (int k = S.N - 1; k >= 0; k--)
    {
        const int o = S.pinv[k], cnt = S.Lp[k + 1] - S.Lp[k];
        const int first = F.npop;
        const int drow = F.pop(0, L.Dinv + k), xrow = F.pop(3, k), arow = accumulate ? F.pop(2, o) : SW_ZERO_ROW;
        const size_t w1 = ops.size() + 1;
        ops.push_back(cnt);
        ops.push_back(0);
        ops.push_back(o);
        emit_pairs(ops, F, w1 - 1, 3, first, cnt, [&](int q) {
            const int u = S.Lp[k] + q, i = S.Li[u];
            const int lrow = F.pop(0, L.Lx + u);
            int opnd;
            if (cache.slot_of[i] >= 0)
                opnd = SW_SLOT0 + cache.slot_of[i];
            else if (FifoSim::far_safe(prod[i], F.npop))
            {
                opnd = F.pop(1, S.pinv[i]);
                H.sw_far++;
            }
            else
            {
                opnd = SW_DIRECT + S.pinv[i];
                H.sw_direct++;
            }
            cache.used(i);
            return pair_word(lrow, opnd);
        });
        const int s = cache.alloc(k);
        ops[w1] = (s >= 0 ? SW_SLOT0 + s : SW_NO_KEEP) | (drow << 8) | (xrow << 16) | (arow << 24);
        prod[k] = F.npop;
    }
    (accumulate ? H.bw_nld : H.bwp_nld) = (int)ld.size();
    H.sw_slots = std::max(H.sw_slots, cache.top);
    pad_tail(ops);
    pad_tail(ld);
}

// ---- KKT mat-vec program (streams.hpp).  Rows = x, y and z rows (the two expansion slots of every
// second-order cone excepted) in elimination order; the pairs of a row keep the order of the
// CSC / CSR data (G entries before A entries in an x row).
void build_matvec(const Symbolic &S, const Layout &L, int max_slots, HostStreams &H, bool pim)
{
    const int n = S.n, p = S.p, zb = S.n + S.p;
    ivec order;
    for (int r = 0; r < zb; r++)
        order.push_back(r);
    for (int i = 0; i < S.m; i++)
        order.push_back(zb + S.zk[i]); // K-space row of z entry i (expanded index)
    const int nrows = (int)order.size();
    std::sort(order.begin(), order.end(), [&](int a, int b) { return S.P[a] < S.P[b]; });
    std::vector<std::vector<std::pair<int, double>>> ent(S.N);
    std::vector<ivec> crow(S.N); // per entry: workspace row of the per-instance coefficient (pim)
    for (int j = 0; j < n; j++)
    {
        for (int k = S.G.p[j]; k < S.G.p[j + 1]; k++)
        {
            ent[j].push_back({zb + S.zk[S.G.i[k]], S.G.x[k]});
            crow[j].push_back(L.Gx + k);
        }
        for (int k = S.A.p[j]; k < S.A.p[j + 1]; k++)
        {
            ent[j].push_back({n + S.A.i[k], S.A.x[k]});
            crow[j].push_back(L.Ax + k);
        }
    }
    for (int i = 0; i < p; i++)
        for (int t = S.Ar.p[i]; t < S.Ar.p[i + 1]; t++)
        {
            ent[n + i].push_back({S.Ar.j[t], S.A.x[S.Ar.v[t]]});
            crow[n + i].push_back(L.Ax + S.Ar.v[t]);
        }
    for (int i = 0; i < S.m; i++)
        for (int t = S.Gr.p[i]; t < S.Gr.p[i + 1]; t++)
        {
            ent[zb + S.zk[i]].push_back({S.Gr.j[t], S.G.x[S.Gr.v[t]]});
            crow[zb + S.zk[i]].push_back(L.Gx + S.Gr.v[t]);
        }
    std::vector<ivec> uses(S.N);
    for (int t = 0; t < nrows; t++)
    {
        const int r = order[t];
        uses[r].push_back(t);
        for (auto &e : ent[r])
            uses[e.first].push_back(t);
    }
    for (ivec &u : uses)
        std::sort(u.begin(), u.end());
    SlotCache cache(max_slots, uses);
    FifoSim F(H.mv_ld);
    // resolves one operand: (operand row, keep row)
    const auto operand = [&](int c) {
        std::pair<int, int> r;
        if (cache.slot_of[c] >= 0)
        {
            r = {SW_SLOT0 + cache.slot_of[c], SW_NO_KEEP};
            cache.used(c);
        }
        else
        {
            r.first = F.pop(2, c);
            cache.used(c);
            const int s = cache.alloc(c);
            r.second = s >= 0 ? SW_SLOT0 + s : SW_NO_KEEP;
        }
        return r;
    };
    for (int t = 0; t < nrows; t++)
    {
        const int r = order[t], cnt = (int)ent[r].size();
        const int kind = r < n ? MV_X : (r < zb ? MV_Y : (r < zb + S.l ? MV_Z : MV_ZC));
        if (cnt > MV_CNT_MASK)
            throw std::logic_error("mat-vec program: row too long");
        const int first = F.npop;
        const int ex0 = F.pop(1, r);
        const auto own = operand(r);
        const int ex1 = kind >= MV_Z ? F.pop(3, r - zb) : SW_ZERO_ROW;
        const size_t w0 = H.mv.size();
        H.mv.push_back(cnt | (kind << MV_KIND_SHIFT));
        H.mv.push_back(ex0 | (own.first << 8) | (own.second << 16) | (ex1 << 24));
        H.mv.push_back(r);
        if (pim)
        { // coefficients are rows too: [coefficient ring row | operand | keep], four pairs per further record
            H.mv.push_back(0);
            if (FifoSim::crosses(first, F.npop - first))
                H.mv[w0] |= SW_SYNC_HDR;
            int q = 0, ntail = 0;
            while (q < cnt)
            {
                const int f = F.npop;
                const size_t at = H.mv.size();
                for (int k = 0; k < 4; k++)
                {
                    if (q >= cnt)
                    {
                        H.mv.push_back(MVP_PAD_PAIR);
                        continue;
                    }
                    const int cr = F.pop(0, crow[r][q]);
                    const auto o = operand(ent[r][q].first);
                    H.mv.push_back(cr | (o.first << 8) | (o.second << 16));
                    q++;
                }
                if (FifoSim::crosses(f, F.npop - f))
                    H.mv[at] |= MVP_SYNC;
                ntail++;
            }
            H.mv[w0] = (H.mv[w0] & ~MV_CNT_MASK) | ntail;
            continue;
        }
        int q = 0;
        const auto pair = [&]() {
            if (q >= cnt)
            {
                H.mv_val.push_back(0.0);
                return MV_PAD_PAIR;
            }
            const auto o = operand(ent[r][q].first);
            H.mv_val.push_back(ent[r][q].second);
            q++;
            return o.first | (o.second << 8);
        };
        H.mv.push_back(pair());
        H.mv_val.push_back(0.0); // first record: {c0, 0}
        if (FifoSim::crosses(first, F.npop - first))
            H.mv[w0] |= SW_SYNC_HDR;
        while (q < cnt)
        {
            const int f = F.npop;
            const size_t at = H.mv.size();
            for (int k = 0; k < 4; k++)
                H.mv.push_back(pair());
            if (FifoSim::crosses(f, F.npop - f))
                H.mv[at] |= MV_SYNC_PAIR;
        }
    }
    H.mv_rows = nrows;
    H.mv_nld = (int)H.mv_ld.size();
    H.sw_slots = std::max(H.sw_slots, cache.top);
    pad_tail(H.mv);
    pad_tail(H.mv_ld);
    pad_tail(H.mv_val);
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
    build_forward(S, L, max_sw_slots, H);
    build_backward(S, L, max_sw_slots, H, true);
    build_backward(S, L, max_sw_slots, H, false);
    if (!build_factor_fast(S, L, max_fa_slots, H, pim))
        build_factor(S, L, max_fa_slots, H, pim);
    build_matvec(S, L, max_sw_slots, H, pim);

}

void refresh_stream_values(const Symbolic &S, const Layout &L, HostStreams &H, bool pim)
{
    HostStreams fresh;
    build_streams(S, L, H.workers, std::max(H.sw_slots, 1), std::max(H.fa_slots, 1), fresh, pim);
    H.fa_val.swap(fresh.fa_val);
    H.mv_val.swap(fresh.mv_val);
}

} // namespace eicos
