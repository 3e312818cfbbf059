// Builds the per-worker instruction streams (see streams.hpp).  Pure host code.
#include "streams.hpp"

#include <algorithm>
#include <cstdint>
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

// ------------------------------------------------------------------ pipe-form row programs (streams.hpp)
int field(int row)
{
    if (row < 0 || row >= PR_MAX_ROWS)
        throw std::logic_error("row program: shared-memory row out of range");
    return row << PR_FIELD_SHIFT;
}

// Host model of the data ring and writer of the record stream.  A record is opened with begin(),
// makes its pops, and is closed with end(), which derives the acquire / release counts from the
// pops made in between.
struct Emitter
{
    Program &P;
    const int NR;
    int npop = 0, released = 0;
    int rec_first = 0;
    struct RelAt
    {
        size_t word;
        int fence_bit;
    };
    std::vector<RelAt> rel_at; // release number -> record word holding its flags
    Emitter(Program &p, int nr) : P(p), NR(nr) {}

    int pop(int sel, int row)
    {
        if (row < 0 || row > LD_ROW_MASK2)
            throw std::logic_error("load list: row out of range");
        P.ld.push_back((sel << LD_SEL_SHIFT) | row);
        return npop++ % RING_ROWS;
    }
    void pad()
    {
        P.ld.push_back(LD_NONE);
        npop++;
        P.pads++;
    }
    bool aligned() const { return NR == 1 || npop % 2 == 0; }
    // vector pop: one row per job, adjacent ring rows (the caller may first hoist a single pop to align)
    int vpop(int sel, int row)
    {
        if (!aligned())
            pad();
        const int r = pop(sel, row);
        if (NR == 2)
            pop(sel + LD_JOB_B, row);
        return r;
    }
    void begin() { rec_first = npop; }
    // May a value whose producing record ended when `prod_after` pops had been made be loaded through the
    // ring as the next pop?  Its group is refilled when group G - RING_GROUPS is released, which happens
    // at the end of the record that makes pop 8 (G - RING_GROUPS + 1) - 1; the producer must be an
    // earlier record.  (The first RING_GROUPS groups are issued before the program starts.)
    bool far_safe(int prod_after) const
    {
        const int at = aligned() ? npop : npop + 1;
        const int G = at / RING_GROUP;
        return G >= RING_GROUPS && prod_after < RING_GROUP * (G - RING_GROUPS + 1);
    }
    // the refill of the group the next pop lands in must be preceded by a proxy fence
    void fence_next_pop()
    {
        const int at = aligned() ? npop : npop + 1;
        const int G = at / RING_GROUP;
        const RelAt &r = rel_at.at(G - RING_GROUPS);
        P.ops[r.word] |= r.fence_bit;
    }
    void end(int w0, int w1, int w2, int w3, bool header)
    {
        const int nacq = (npop + RING_GROUP - 1) / RING_GROUP - (rec_first + RING_GROUP - 1) / RING_GROUP;
        const int nrel = npop / RING_GROUP - released;
        if (nacq > 3 || nrel > 3 || nacq < 0 || nrel < 0)
            throw std::logic_error("row program: a record spans too many ring groups");
        const size_t at = P.ops.size();
        if (header)
            w0 |= (nacq << PH_NACQ_SHIFT) | (nrel << PH_NREL_SHIFT);
        else
        {
            if (w0 & ((1 << PR_FIELD_SHIFT) - 1))
                throw std::logic_error("row program: flag bits of a tail record are in use");
            w0 |= (nacq << PT_NACQ_SHIFT) | (nrel << PT_NREL_SHIFT);
        }
        for (int k = 0; k < nrel; k++)
            rel_at.push_back({at, header ? PH_FENCE : PT_FENCE});
        released += nrel;
        P.ops.push_back(w0);
        P.ops.push_back(w1);
        P.ops.push_back(w2);
        P.ops.push_back(w3);
    }
    void finish(int slot_rows)
    {
        begin();
        end(PH_END, 0, 0, 0, true);
        P.nchunks = (int)((P.ops.size() + OPS_CHUNK_WORDS - 1) / OPS_CHUNK_WORDS);
        while (P.ops.size() % OPS_CHUNK_WORDS)
            P.ops.push_back(0);
        P.ops.insert(P.ops.end(), OPS_CHUNK_WORDS, 0); // the record look-ahead may read one record past END
        P.nld = (int)P.ld.size();
        while (P.ld.size() % RING_ROWS)
            P.ld.push_back(LD_NONE);
        P.ld.insert(P.ld.end(), 2 * RING_ROWS, LD_NONE); // the lanes read their next word two rounds ahead
        P.slot_rows = slot_rows;
    }
};

// one multiply-add of a dot-form row: the L value (single pop) and the gathered vector value
struct PairSrc
{
    int lrow;    // workspace row of the L value
    int value;   // index of the gathered value (slot cache / producer bookkeeping)
    int far_sel; // load-list selector of the vector the value lives in
    int far_row; // its row inside that vector
};

// Emits one dot-form row: header record (with up to `ninl` pairs inline) + tail records.
//   pre()      makes the header's own pops (right-hand side, pivot, ...); called once the header is open
//   header(rp) closes the header record: rp.w0 = tail counts / form flags for word 0, rp.inl = inline pair words
struct RowPairs
{
    int inl[2] = {PR_PAD_PAIR, PR_PAD_PAIR};
    int w0 = 0;
};

template <class Pre, class Header>
void emit_row(Emitter &E, SlotCache &cache, const ivec &prod, const std::vector<PairSrc> &pairs, int ninl, Program &P,
              Pre pre, Header header)
{
    const int cnt = (int)pairs.size(), NR = E.NR;
    // kind of every gathered value: 0 slot, 1 far (through the ring), 2 direct (straight from global memory).
    // far_safe() only gets easier as the pop position advances, so testing at the current position is safe.
    ivec kind(cnt);
    bool slow = false;
    for (int q = 0; q < cnt; q++)
    {
        kind[q] = cache.slot_of[pairs[q].value] >= 0 ? 0 : (E.far_safe(prod[pairs[q].value]) ? 1 : 2);
        slow = slow || kind[q] == 2;
    }
    const auto operand = [&](int q) -> int { // field of pair q's gathered value
        const PairSrc &pr = pairs[q];
        int f;
        if (kind[q] == 0)
            f = field(PR_SLOT0 + NR * cache.slot_of[pr.value]);
        else
        {
            if (!E.aligned())
                E.pad();
            if (!E.far_safe(prod[pr.value]))
                throw std::logic_error("row program: far operand not safe at emission");
            E.fence_next_pop();
            f = field(E.vpop(pr.far_sel, pr.far_row));
            P.far++;
        }
        cache.used(pr.value);
        return f;
    };
    E.begin();
    RowPairs rp;
    if (slow)
    { // one record per pair: [L field | flags, 0 = field in word 2 / 1 = global row in word 2, value, -]
        if (cnt > PH_NTAIL_MASK)
            throw std::logic_error("row program: row too long for the slow form");
        pre();
        rp.w0 = PH_SLOW | cnt;
        header(rp);
        for (int q = 0; q < cnt; q++)
        {
            const PairSrc &pr = pairs[q];
            E.begin();
            const int lf = field(E.pop(0, pr.lrow));
            if (kind[q] == 2)
            {
                cache.used(pr.value);
                P.direct++;
                E.end(lf, 1, pr.far_row, 0, false);
            }
            else
                E.end(lf, 0, operand(q), 0, false);
        }
        return;
    }
    int q = 0;
    const auto pair_word = [&](int qq, int lfield = -1) {
        const int lf = lfield >= 0 ? lfield : field(E.pop(0, pairs[qq].lrow));
        const int of = operand(qq);
        return lf | (of << 16);
    };
    const int rest = cnt > ninl ? cnt - ninl : 0;
    const int n4 = rest / 4, r4 = rest % 4;
    // remainder 1..2 -> one 2-pair record; remainder 3 -> one more 4-pair record
    const int ntail4 = n4 + (r4 == 3 ? 1 : 0), has2 = (r4 == 1 || r4 == 2) ? 1 : 0;
    if (ntail4 > PH_NTAIL_MASK)
        throw std::logic_error("row program: row too long");
    rp.w0 = ntail4 | (has2 ? PH_HAS2 : 0) | (cnt > 0 && ninl > 0 ? PH_INLINE : 0);
    // NR = 2: vector pops want an even ring position; an L pop made first realigns it without padding
    int hoisted = -1;
    if (!E.aligned() && cnt > 0 && ninl > 0)
        hoisted = field(E.pop(0, pairs[0].lrow));
    pre();
    for (int k = 0; k < ninl && q < cnt; k++, q++)
        rp.inl[k] = pair_word(q, k == 0 ? hoisted : -1);
    header(rp);
    for (int t = 0; t < ntail4; t++)
    {
        E.begin();
        int w[4];
        for (int k = 0; k < 4; k++)
            w[k] = q < cnt ? pair_word(q++) : PR_PAD_PAIR;
        E.end(w[0], w[1], w[2], w[3], false);
    }
    if (has2)
    {
        E.begin();
        int w[2];
        for (int k = 0; k < 2; k++)
            w[k] = q < cnt ? pair_word(q++) : PR_PAD_PAIR;
        E.end(w[0], w[1], 0, 0, false);
    }
}

// ---- forward sweep  xw = L^-1 P rhs, rows of L in elimination order, dot form in ascending column
// order (the summation order of Eigen's column-oriented forward substitution).  L is stored once,
// column-major; the row-order walk is just the order of the load list.
// Load-list selectors: 1 = right-hand side (KKT order), 3 = the work vector xw (far gathers).
void build_forward(const Symbolic &S, const Layout &L, int NR, int max_values, Program &P)
{
    std::vector<ivec> uses(S.N);
    for (int k = 0; k < S.N; k++)
        uses[k].assign(S.Li.begin() + S.Lp[k], S.Li.begin() + S.Lp[k + 1]);
    SlotCache cache(max_values, uses);
    Emitter E(P, NR);
    ivec prod(S.N, 0);
    std::vector<PairSrc> pairs;
    for (int i = 0; i < S.N; i++)
    {
        pairs.clear();
        for (int t = S.Lr.p[i]; t < S.Lr.p[i + 1]; t++)
            pairs.push_back({L.Lx + S.Lr.v[t], S.Lr.j[t], 3, S.Lr.j[t]});
        int rhs_field = 0;
        size_t hdr_at = 0;
        emit_row(
            E, cache, prod, pairs, 2, P, [&]() { rhs_field = field(E.vpop(1, S.pinv[i])); },
            [&](const RowPairs &rp) {
                hdr_at = P.ops.size();
                E.end(rp.w0, rhs_field, rp.inl[0], rp.inl[1], true);
            });
        const int s = cache.alloc(i);
        P.ops[hdr_at + 1] |= field(s >= 0 ? PR_SLOT0 + NR * s : PR_TRASH) << 16;
        prod[i] = E.npop;
    }
    E.finish(NR * cache.top);
}

// ---- backward sweep  out = P' L^-T D^-1 xw, columns in reverse elimination order (dot form, Eigen's
// order); results land in KKT order.  The home of a finished entry is its output row.
// accumulate: the program also loads the row of the solution it adds its result to (refinement
// rounds); the plain program of the first solve leaves those loads out.
// Load-list selectors: 1 = output vector (far gathers), 2 = accumulated solution, 3 = xw.
void build_backward(const Symbolic &S, const Layout &L, int NR, int max_values, Program &P, bool accumulate)
{
    std::vector<ivec> uses(S.N); // value i is used by the columns of row i, latest column first
    for (int i = 0; i < S.N; i++)
        for (int t = S.Lr.p[i + 1] - 1; t >= S.Lr.p[i]; t--)
            uses[i].push_back(S.N - 1 - S.Lr.j[t]);
    for (const ivec &u : uses)
        if (!std::is_sorted(u.begin(), u.end()))
            throw std::logic_error("rows of L must have ascending columns");
    SlotCache cache(max_values, uses);
    Emitter E(P, NR);
    ivec prod(S.N, 0);
    std::vector<PairSrc> pairs;
    for (int k = S.N - 1; k >= 0; k--)
    {
        const int o = S.pinv[k];
        pairs.clear();
        for (int u = S.Lp[k]; u < S.Lp[k + 1]; u++)
            pairs.push_back({L.Lx + u, S.Li[u], 1, S.pinv[S.Li[u]]});
        int dfield = 0, xfield = 0, afield = field(PR_ZERO);
        bool dpopped = false;
        size_t hdr_at = 0;
        emit_row(
            E, cache, prod, pairs, 0, P,
            [&]() {
                if (!E.aligned()) // the single pop first realigns the ring for the vector pops
                {
                    dfield = field(E.pop(0, L.Dinv + k));
                    dpopped = true;
                }
                xfield = field(E.vpop(3, k));
                if (accumulate)
                    afield = field(E.vpop(2, o));
                if (!dpopped)
                    dfield = field(E.pop(0, L.Dinv + k));
            },
            [&](const RowPairs &rp) {
                hdr_at = P.ops.size();
                E.end(rp.w0, dfield | (xfield << 16), afield << 16, o, true);
            });
        const int s = cache.alloc(k);
        P.ops[hdr_at + 2] |= field(s >= 0 ? PR_SLOT0 + NR * s : PR_TRASH);
        prod[k] = E.npop;
    }
    E.finish(NR * cache.top);
}

// ---- KKT mat-vec program (streams.hpp).  Rows = x, y and z rows (the two expansion slots of every
// second-order cone excepted) in elimination order; the pairs of a row keep the order of the
// CSC / CSR data (G entries before A entries in an x row).
// Load-list selectors: 1 = vector of extra 0 (rhs | c,b,h), 2 = operand vector, 3 = vector of extra 1
// (LP scalings | s; one row per instance even in a pair program).
void push_double(ivec &ops, double v)
{
    int32_t w[2];
    std::memcpy(w, &v, sizeof(v));
    ops.push_back(w[0]);
    ops.push_back(w[1]);
}

void build_matvec(const Symbolic &S, const Layout &L, int NR, int max_values, Program &P, int &mv_rows, bool pim)
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
    SlotCache cache(max_values, uses);
    Emitter E(P, NR);
    // resolves one operand: (operand field, keep field)
    const auto operand = [&](int c) {
        std::pair<int, int> r;
        if (cache.slot_of[c] >= 0)
        {
            r = {field(PR_SLOT0 + NR * cache.slot_of[c]), field(PR_TRASH)};
            cache.used(c);
        }
        else
        {
            r.first = field(E.vpop(2, c));
            cache.used(c);
            const int s = cache.alloc(c);
            r.second = field(s >= 0 ? PR_SLOT0 + NR * s : PR_TRASH);
        }
        return r;
    };
    const int pad_pair = field(PR_ZERO) | (field(PR_TRASH) << 16);
    for (int t = 0; t < nrows; t++)
    {
        const int r = order[t], cnt = (int)ent[r].size();
        const int kind = r < n ? MV_X : (r < zb ? MV_Y : (r < zb + S.l ? MV_Z : MV_ZC));
        E.begin();
        int ex1 = field(PR_ZERO);
        bool ex1_popped = false;
        if (kind >= MV_Z && !E.aligned())
        {
            ex1 = field(E.pop(3, r - zb));
            ex1_popped = true;
        }
        const int ex0 = field(E.vpop(1, r));
        const auto own = operand(r);
        if (kind >= MV_Z && !ex1_popped)
            ex1 = field(E.pop(3, r - zb));
        const int per = pim ? 2 : 4; // pairs per full group
        const int ng = cnt / per, rem = cnt % per;
        // shared coefficients: remainder 1..2 -> a 2-pair group, 3 -> one more full group; pim: remainder 1 -> one more record
        const int ngroups = pim ? ng + (rem ? 1 : 0) : ng + (rem == 3 ? 1 : 0);
        const int has2 = !pim && (rem == 1 || rem == 2);
        if (ngroups > PH_NTAIL_MASK)
            throw std::logic_error("mat-vec program: row too long");
        E.end(ngroups | (has2 ? PH_HAS2 : 0) | (kind << PH_KIND_SHIFT), ex0 | (own.first << 16), own.second | (ex1 << 16), r, true);
        int q = 0;
        if (pim)
        { // [coefficient field | operand field << 16, keep field] x 2 per record
            for (int g = 0; g < ngroups; g++)
            {
                E.begin();
                int w[4];
                for (int k = 0; k < 2; k++)
                {
                    if (q >= cnt)
                    {
                        w[2 * k] = PR_PAD_PAIR;
                        w[2 * k + 1] = field(PR_TRASH);
                        continue;
                    }
                    const int cf = field(E.pop(0, crow[r][q]));
                    const auto o = operand(ent[r][q].first);
                    w[2 * k] = cf | (o.first << 16);
                    w[2 * k + 1] = o.second;
                    q++;
                }
                E.end(w[0], w[1], w[2], w[3], false);
            }
            continue;
        }
        const auto group = [&](int np) { // np pairs + their coefficients
            E.begin();
            int w[4] = {pad_pair, pad_pair, 0, 0};
            double c[4] = {0.0, 0.0, 0.0, 0.0};
            for (int k = 0; k < np; k++)
            {
                w[k] = pad_pair;
                if (q < cnt)
                {
                    const auto o = operand(ent[r][q].first);
                    w[k] = o.first | (o.second << 16);
                    c[k] = ent[r][q].second;
                    q++;
                }
            }
            E.end(w[0], w[1], w[2], w[3], false);
            for (int k = 0; k < np; k++)
                push_double(P.ops, c[k]);
        };
        for (int g = 0; g < ngroups; g++)
            group(4);
        if (has2)
            group(2);
    }
    mv_rows = nrows;
    E.finish(NR * cache.top);
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
    for (int NR = 1; NR <= 2; NR++)
    {
        const int maxv = std::max(1, std::min(max_sw_slots, (PR_MAX_ROWS - PR_SLOT0) / NR));
        const char *dbg = std::getenv("EICOS_DBG_STARVE");
        const int dm = dbg ? std::atoi(dbg) : 0;
        build_forward(S, L, NR, dm & 1 ? 2 : maxv, H.fw[NR - 1]);
        build_backward(S, L, NR, dm & 2 ? 2 : maxv, H.bw[NR - 1], true);
        build_backward(S, L, NR, dm & 4 ? 2 : maxv, H.bwp[NR - 1], false);
        build_matvec(S, L, NR, dm & 8 ? 2 : maxv, H.mv[NR - 1], H.mv_rows, pim);
        for (const Program *q : {&H.fw[NR - 1], &H.bw[NR - 1], &H.bwp[NR - 1], &H.mv[NR - 1]})
            H.sw_slots = std::max(H.sw_slots, q->slot_rows / NR);
    }
    H.sw_far = H.fw[0].far + H.bw[0].far;
    H.sw_direct = H.fw[0].direct + H.bw[0].direct;
    if (!build_factor_fast(S, L, max_fa_slots, H, pim))
        build_factor(S, L, max_fa_slots, H, pim);
}

void refresh_stream_values(const Symbolic &S, const Layout &L, HostStreams &H, bool pim)
{
    HostStreams fresh;
    build_streams(S, L, H.workers, std::max(H.sw_slots, 1), std::max(H.fa_slots, 1), fresh, pim);
    H.fa_val.swap(fresh.fa_val);
    for (int k = 0; k < 2; k++)
    {
        if (fresh.mv[k].ops.size() != H.mv[k].ops.size())
            throw std::logic_error("mat-vec program changed shape on a value refresh");
        H.mv[k].ops.swap(fresh.mv[k].ops);
    }
}

} // namespace eicos
