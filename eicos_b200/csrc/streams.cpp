// Builds the per-worker instruction streams (see streams.hpp).  Pure host code.
#include "streams.hpp"

#include <algorithm>
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

// mat-vec row set for one worker: blocks of rows whose gathers fit the staging slots
template <class Emit>
void rowset(int rows, int W, ivec &s, dvec &v, ivec &seg, Emit &&emit)
{
    seg.assign((size_t)W * 3, 0);
    for (int w = 0; w < W; w++)
    {
        align_chunk(s);
        align_chunk(v);
        seg[w * 3] = (int)s.size();
        seg[w * 3 + 1] = (int)v.size();
        int nblocks = 0;
        std::vector<ivec> rowwords;
        for (int r = w; r < rows; r += W)
        {
            ivec words;
            emit(r, words, v);
            rowwords.push_back(std::move(words));
        }
        size_t a = 0;
        while (a < rowwords.size())
        {
            int slots = (int)rowwords[a].size() - 1 + ROW_EXTRA_SLOTS;
            size_t b = a + 1;
            if (slots > STAGE_SLOTS)
                s.push_back(-1);
            else
            {
                while (b < rowwords.size() && slots + (int)rowwords[b].size() - 1 + ROW_EXTRA_SLOTS <= STAGE_SLOTS)
                {
                    slots += (int)rowwords[b].size() - 1 + ROW_EXTRA_SLOTS;
                    b++;
                }
                s.push_back((int)(b - a));
            }
            for (size_t q = a; q < b; q++)
                s.insert(s.end(), rowwords[q].begin(), rowwords[q].end());
            nblocks++;
            a = b;
        }
        seg[w * 3 + 2] = nblocks;
    }
    pad_tail(s);
    pad_tail(v);
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

// ---- forward sweep  xw = L^-1 P rhs, column by column in elimination order (scatter form, the
// order Eigen's own forward substitution uses): step k consumes the accumulator of row k and
// subtracts L(i,k) x_k from the accumulators of the rows i of column k.  An accumulator starts from
// its right-hand-side row on the first touch.
//   ops:   [src, cnt, cnt x target]        load list: [rhs_k if untouched] { L(i,k) [rhs_i on a first touch] }
void build_forward(const Symbolic &S, const Layout &L, int max_slots, HostStreams &H)
{
    SlotPool pool(max_slots);
    ivec code(S.N, -1);
    for (int k = 0; k < S.N; k++)
    {
        if (code[k] < 0)
        {
            H.fw.push_back(SRC_FIFO);
            H.fw_ld.push_back(~S.pinv[k]);
        }
        else
            H.fw.push_back(code[k]);
        H.fw.push_back(S.Lp[k + 1] - S.Lp[k]);
        for (int u = S.Lp[k]; u < S.Lp[k + 1]; u++)
        {
            const int i = S.Li[u];
            H.fw_ld.push_back(L.Lx + u);
            if (code[i] < 0)
            {
                code[i] = pool.take(L.xw + i);
                H.fw.push_back(code[i] | (OPK_FIFO << OPK_SHIFT));
                H.fw_ld.push_back(~S.pinv[i]);
            }
            else
                H.fw.push_back(code[i]);
        }
        pool.give(code[k]);
    }
    H.fw_nld = (int)H.fw_ld.size();
    H.sw_slots = std::max(H.sw_slots, pool.top);
    H.sw_home += pool.home;
    pad_tail(H.fw);
    pad_tail(H.fw_ld);
}

// ---- backward sweep  out = P' L^-T D^-1 xw, columns in reverse elimination order (dot form):
// step k gathers the finished entries i of column k.  A finished entry is kept in a slot until the
// first column of its row has used it; its home is its own output row (relative to `out`).
//   ops:   [out row, keep code | -1, cnt, cnt x gather]
//   load list: D_k, xw_k, { L(i,k) }, ~out row (accumulated solution; skipped on a plain solve)
void build_backward(const Symbolic &S, const Layout &L, int max_slots, HostStreams &H)
{
    SlotPool pool(max_slots);
    ivec code(S.N, -1), minrow(S.N, -1);
    for (int i = 0; i < S.N; i++)
        if (S.Lr.p[i + 1] > S.Lr.p[i])
        {
            int mn = S.N;
            for (int t = S.Lr.p[i]; t < S.Lr.p[i + 1]; t++)
                mn = std::min(mn, S.Lr.j[t]);
            minrow[i] = mn;
        }
    for (int k = S.N - 1; k >= 0; k--)
    {
        const int o = S.pinv[k];
        H.bw_ld.push_back(L.D + k);
        H.bw_ld.push_back(L.xw + k);
        H.bw.push_back(o);
        const size_t keep_at = H.bw.size();
        H.bw.push_back(-1);
        H.bw.push_back(S.Lp[k + 1] - S.Lp[k]);
        for (int u = S.Lp[k]; u < S.Lp[k + 1]; u++)
        {
            H.bw_ld.push_back(L.Lx + u);
            H.bw.push_back(code[S.Li[u]]);
        }
        for (int u = S.Lp[k]; u < S.Lp[k + 1]; u++)
            if (minrow[S.Li[u]] == k)
                pool.give(code[S.Li[u]]);
        if (minrow[k] >= 0)
        {
            code[k] = pool.take(o);
            H.bw[keep_at] = code[k] < SLOT_HOME ? code[k] : -1;
        }
        H.bw_ld.push_back(~o);
    }
    H.bw_nld = (int)H.bw_ld.size();
    H.sw_slots = std::max(H.sw_slots, pool.top);
    H.sw_home += pool.home;
    pad_tail(H.bw);
    pad_tail(H.bw_ld);
}

// ---- numeric factorisation, right-looking in elimination order.  Every entry (i,j) of L and every
// pivot owns an accumulator that starts from the KKT value (shared constant, per-instance scaling
// value, or 0 for fill) on its first touch.  Step k: d = acc(k,k); for the rows i of column k
// a_i = acc(i,k), l_i = a_i / d (stored, column-major = the order both sweeps stream it in); then
// for every pair i1 >= i2 of the column  acc(i1,i2) -= l_i2 * a_i1  (the products Eigen's
// up-looking kernel forms, accumulated in ascending k).
//   ops:   [src_d, cnt, cnt x src, cnt(cnt+1)/2 x target]   constants in fa_val, V rows in the load list
void build_factor(const Symbolic &S, const Layout &L, int max_slots, HostStreams &H)
{
    struct Init
    {
        int kind = OPK_ZERO, vrow = -1;
        double c = 0.0;
    };
    std::vector<Init> di(S.N), ei(S.nnzL);
    for (int j = 0; j < S.N; j++)
        for (int e = S.KLp[j]; e < S.KLp[j + 1]; e++)
        {
            const int slot = S.KLslot[e], vi = S.Kvidx[slot], pos = S.KLpos[e];
            Init &t = pos < 0 ? di[j] : ei[S.Lp[j] + pos];
            if (vi >= 0)
            {
                t.kind = OPK_FIFO;
                t.vrow = L.V + vi;
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
} // namespace

void build_streams(const Symbolic &S, const Layout &L, int W, int max_sw_slots, int max_fa_slots, HostStreams &H)
{
    H = HostStreams();
    H.workers = W;
    for (int k = 0; k < S.N; k++)
        for (int u = S.Lp[k]; u + 1 < S.Lp[k + 1]; u++)
            if (S.Li[u] >= S.Li[u + 1])
                throw std::logic_error("columns of L must have ascending rows");
    build_forward(S, L, max_sw_slots, H);
    build_backward(S, L, max_sw_slots, H);
    build_factor(S, L, max_fa_slots, H);

    // ---- mat-vec row sets (K-space gather indices); a row = [cnt, idx...]
    const int n = S.n, p = S.p, zb = S.n + S.p;
    rowset(n, W, H.rx, H.rx_val, H.rx_seg, [&](int j, ivec &s, dvec &v) {
        s.push_back((S.G.p[j + 1] - S.G.p[j]) + (S.A.p[j + 1] - S.A.p[j]));
        for (int k = S.G.p[j]; k < S.G.p[j + 1]; k++)
        {
            s.push_back(zb + S.zk[S.G.i[k]]);
            v.push_back(S.G.x[k]);
        }
        for (int k = S.A.p[j]; k < S.A.p[j + 1]; k++)
        {
            s.push_back(n + S.A.i[k]);
            v.push_back(S.A.x[k]);
        }
    });
    rowset(p, W, H.ry, H.ry_val, H.ry_seg, [&](int i, ivec &s, dvec &v) {
        s.push_back(S.Ar.p[i + 1] - S.Ar.p[i]);
        for (int t = S.Ar.p[i]; t < S.Ar.p[i + 1]; t++)
        {
            s.push_back(S.Ar.j[t]);
            v.push_back(S.A.x[S.Ar.v[t]]);
        }
    });
    auto grow = [&](int i, ivec &s, dvec &v) {
        s.push_back(S.Gr.p[i + 1] - S.Gr.p[i]);
        for (int t = S.Gr.p[i]; t < S.Gr.p[i + 1]; t++)
        {
            s.push_back(S.Gr.j[t]);
            v.push_back(S.G.x[S.Gr.v[t]]);
        }
    };
    rowset(S.l, W, H.rz, H.rz_val, H.rz_seg, grow);
    // cones: [dim, first expanded index, first q row] then one row per cone entry (no staging)
    H.rc_seg.assign((size_t)W * 2, 0);
    for (int w = 0; w < W; w++)
    {
        align_chunk(H.rc);
        align_chunk(H.rc_val);
        H.rc_seg[w * 2] = (int)H.rc.size();
        H.rc_seg[w * 2 + 1] = (int)H.rc_val.size();
        for (int c = w; c < S.nc; c += W)
        {
            H.rc.push_back(S.q[c]);
            H.rc.push_back(S.cone_k[c]);
            H.rc.push_back(S.cone_q[c]);
            for (int k = 0; k < S.q[c]; k++)
                grow(S.cone_z[c] + k, H.rc, H.rc_val);
        }
    }
    pad_tail(H.rc);
    pad_tail(H.rc_val);
}

void refresh_stream_values(const Symbolic &S, const Layout &L, HostStreams &H)
{
    HostStreams fresh;
    build_streams(S, L, H.workers, std::max(H.sw_slots, 1), std::max(H.fa_slots, 1), fresh);
    H.fa_val.swap(fresh.fa_val);
    H.rx_val.swap(fresh.rx_val);
    H.ry_val.swap(fresh.ry_val);
    H.rz_val.swap(fresh.rz_val);
    H.rc_val.swap(fresh.rc_val);
}

} // namespace eicos
