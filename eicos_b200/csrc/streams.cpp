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

// tasks of worker w in a phase of the elimination-tree schedule, in the order the device walks them
ivec worker_tasks(const Symbolic &S, const Phase &f, int w, int W, bool backward)
{
    ivec t;
    if (f.parallel)
    {
        if (!backward)
            for (int q = f.begin + w; q < f.end; q += W)
                t.push_back(S.tasks[q]);
        else
            for (int q = f.end - 1 - w; q >= f.begin; q -= W)
                t.push_back(S.tasks[q]);
    }
    else if (w == 0)
    {
        if (!backward)
            for (int q = f.begin; q < f.end; q++)
                t.push_back(S.tasks[q]);
        else
            for (int q = f.end - 1; q >= f.begin; q--)
                t.push_back(S.tasks[q]);
    }
    return t;
}

// One task of a triangular sweep: out = init - sum_k L[src_k] * vec[gather_k]
struct SweepTask
{
    int h0, h1;   // header words (meaning depends on the direction, see tile_program.hpp)
    ivec gather;  // row to gather from, or a FWD_PREV* code
    ivec src;     // natural index of the L value (CSR index t for forward, CSC index u for backward)
};

struct SweepBuilder
{
    int W;
    int init_slots; // staging slots the start value of a task needs (1 forward, 2 backward)
    ivec &stream, &seg, &pos;
    int nphases = 0, next_pos = 0;

    SweepBuilder(int W_, int init_slots_, ivec &stream_, ivec &seg_, ivec &pos_)
        : W(W_), init_slots(init_slots_), stream(stream_), seg(seg_), pos(pos_) {}

    int assign(const SweepTask &t)
    {
        const int first = next_pos;
        for (int s : t.src)
            pos[s] = next_pos++;
        return first;
    }

    // independent tasks, already distributed over the workers
    void parallel_phase(const std::vector<std::vector<SweepTask>> &per_worker)
    {
        for (int w = 0; w < W; w++)
        {
            align_chunk(stream);
            const std::vector<SweepTask> &tk = per_worker[w];
            const int ioff = (int)stream.size(), voff = next_pos;
            int nblocks = 0;
            size_t a = 0;
            while (a < tk.size())
            {
                int slots = init_slots + 2 * (int)tk[a].gather.size();
                size_t b = a + 1;
                if (slots > STAGE_SLOTS)
                    stream.push_back(-1); // oversize task: processed without staging
                else
                {
                    while (b < tk.size() && slots + init_slots + 2 * (int)tk[b].gather.size() <= STAGE_SLOTS)
                    {
                        slots += init_slots + 2 * (int)tk[b].gather.size();
                        b++;
                    }
                    stream.push_back((int)(b - a));
                }
                for (size_t q = a; q < b; q++)
                {
                    assign(tk[q]);
                    stream.push_back(tk[q].h0);
                    stream.push_back(tk[q].h1);
                    stream.push_back((int)tk[q].gather.size());
                    stream.insert(stream.end(), tk[q].gather.begin(), tk[q].gather.end());
                }
                nblocks++;
                a = b;
            }
            seg.insert(seg.end(), {ioff, nblocks, voff, (int)SEG_BLOCKS});
        }
        nphases++;
    }

    // a dependent chain, walked by worker 0
    void serial_phase(const std::vector<SweepTask> &tk)
    {
        for (int w = 0; w < W; w++)
        {
            align_chunk(stream);
            const int ioff = (int)stream.size(), voff = next_pos;
            if (w != 0)
            {
                seg.insert(seg.end(), {ioff, 0, voff, (int)SEG_SERIAL});
                continue;
            }
            auto header = [&](size_t a) {
                stream.push_back(tk[a].h0);
                stream.push_back(tk[a].h1);
                stream.push_back((int)tk[a].gather.size());
            };
            if (!tk.empty())
                header(0);
            for (size_t a = 0; a < tk.size(); a++)
            { // the header of task a+1 precedes the entries of task a (software pipelining on the device)
                if (a + 1 < tk.size())
                    header(a + 1);
                assign(tk[a]);
                stream.insert(stream.end(), tk[a].gather.begin(), tk[a].gather.end());
            }
            seg.insert(seg.end(), {ioff, (int)tk.size(), voff, (int)SEG_SERIAL});
        }
        nphases++;
    }
};

// Schedules one sweep.  `entries(j)` lists (other index, natural value index) of task j;
// `header` fills the two header words of a task (kind: 0 whole task, 1 external part of a chain task,
// 2 chain task that continues a stored partial result, 3 chain task without external part).
template <class Entries, class Header, class GatherRow>
void schedule_sweep(const Symbolic &S, int W, bool backward, SweepBuilder &B, Entries entries, Header header, GatherRow gather_row)
{
    const int nph = (int)S.phases.size();
    ivec where(S.N, -1); // position inside the current serial phase
    for (int step = 0; step < nph; step++)
    {
        const int ph = backward ? nph - 1 - step : step;
        const Phase &f = S.phases[ph];
        if (f.parallel)
        {
            std::vector<std::vector<SweepTask>> pw(W);
            for (int w = 0; w < W; w++)
                for (int j : worker_tasks(S, f, w, W, backward))
                {
                    SweepTask t;
                    header(j, 0, t);
                    for (auto &e : entries(j))
                    {
                        t.gather.push_back(gather_row(e.first));
                        t.src.push_back(e.second);
                    }
                    pw[w].push_back(std::move(t));
                }
            B.parallel_phase(pw);
            continue;
        }
        const ivec tk = worker_tasks(S, f, 0, W, backward);
        for (size_t a = 0; a < tk.size(); a++)
            where[tk[a]] = (int)a;
        // external part: terms that come from phases already processed -> one parallel phase
        std::vector<std::vector<SweepTask>> pw(W);
        std::vector<char> has_ext(tk.size(), 0);
        int rr = 0;
        for (size_t a = 0; a < tk.size(); a++)
        {
            SweepTask t;
            header(tk[a], 1, t);
            for (auto &e : entries(tk[a]))
                if (where[e.first] < 0)
                {
                    t.gather.push_back(gather_row(e.first));
                    t.src.push_back(e.second);
                }
            if (!t.gather.empty())
            {
                has_ext[a] = 1;
                pw[rr++ % W].push_back(std::move(t));
            }
        }
        if (rr > 0)
            B.parallel_phase(pw);
        // internal part: the recurrence along the chain
        std::vector<SweepTask> chain;
        for (size_t a = 0; a < tk.size(); a++)
        {
            SweepTask t;
            header(tk[a], has_ext[a] ? 2 : 3, t);
            for (auto &e : entries(tk[a]))
            {
                const int wpos = where[e.first];
                if (wpos < 0)
                    continue;
                const int dist = (int)a - wpos;
                if (dist <= 0)
                    throw std::logic_error("serial phase: dependency on a later task");
                t.gather.push_back(dist == 1 ? FWD_PREV1 : dist == 2 ? FWD_PREV2 : dist == 3 ? FWD_PREV3 : gather_row(e.first));
                t.src.push_back(e.second);
            }
            chain.push_back(std::move(t));
        }
        B.serial_phase(chain);
        for (int j : tk)
            where[j] = -1;
    }
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
} // namespace

void build_streams(const Symbolic &S, int W, HostStreams &H)
{
    H = HostStreams();
    H.workers = W;
    const int nph = (int)S.phases.size();
    H.fw_pos.assign(S.nnzL, -1);
    H.bw_pos.assign(S.nnzL, -1);

    // ---- forward sweep: rows of L.  header = [xw row i, rhs row pinv[i] | INIT_PARTIAL]
    {
        SweepBuilder B(W, 1, H.fw, H.fw_seg, H.fw_pos);
        schedule_sweep(
            S, W, false, B,
            [&](int i) {
                std::vector<std::pair<int, int>> e;
                for (int t = S.Lr.p[i]; t < S.Lr.p[i + 1]; t++)
                    e.push_back({S.Lr.j[t], t});
                return e;
            },
            [&](int i, int kind, SweepTask &t) { // kind: 0 whole row, 1 external part, 2 chain after an external part, 3 chain
                t.h0 = i;
                t.h1 = kind == 2 ? INIT_PARTIAL : S.pinv[i];
            },
            [&](int c) { return c; });
        H.nph_fw = B.nphases;
        if (B.next_pos != S.nnzL)
            throw std::logic_error("forward stream does not cover L");
        pad_tail(H.fw);
    }
    // ---- backward sweep: columns of L, results land in KKT order.
    //      header = [xw/Dinv row j | INIT_PARTIAL, out row pinv[j]]
    {
        SweepBuilder B(W, 2, H.bw, H.bw_seg, H.bw_pos);
        schedule_sweep(
            S, W, true, B,
            [&](int j) {
                std::vector<std::pair<int, int>> e;
                for (int u = S.Lp[j]; u < S.Lp[j + 1]; u++)
                    e.push_back({S.Li[u], u});
                return e;
            },
            [&](int j, int kind, SweepTask &t) { // an external part stores a partial result: out row encoded as ~row
                t.h0 = kind == 2 ? INIT_PARTIAL : j;
                t.h1 = kind == 1 ? ~S.pinv[j] : S.pinv[j];
            },
            [&](int r) { return S.pinv[r]; });
        H.nph_bw = B.nphases;
        if (B.next_pos != S.nnzL)
            throw std::logic_error("backward stream does not cover L");
        pad_tail(H.bw);
    }

    // ---- numeric factorisation, left-looking by column (every position explicit).
    // Task = [j, kind, cnt, nK, nR] {kind 2: cnt x bwpos of the stored partial column}
    //        nK x [vidx, pos]  groups of <= FA_GROUP row entries: headers [k, fwpos, tail len] then
    //        their tails (rel, bwpos) ...  cnt x [bwpos, fwpos].
    // kind 0: whole column.  Columns of a serial phase (a chain of the elimination tree) are split:
    // kind 1 = contributions of columns OUTSIDE the chain, done in a parallel phase, partial column
    // stored un-normalised in D / Lx; kind 2 = chain task continuing such a partial; kind 3 = chain
    // task without external contributions.
    ivec Lcsr(S.nnzL);
    for (int t = 0; t < S.nnzL; t++)
        Lcsr[S.Lr.v[t]] = t;
    H.fa_seg.clear();
    H.nph_fa = 0;
    {
        ivec where(S.N, -1);
        auto emit_task = [&](int j, int kind, const ivec &rows /* CSR indices t */) {
            const int cnt = S.Lp[j + 1] - S.Lp[j];
            const bool with_k = kind != 2;
            H.fa.push_back(j);
            H.fa.push_back(kind);
            H.fa.push_back(cnt);
            H.fa.push_back(with_k ? S.KLp[j + 1] - S.KLp[j] : 0);
            H.fa.push_back((int)rows.size());
            if (kind == 2)
                for (int q = 0; q < cnt; q++)
                    H.fa.push_back(H.bw_pos[S.Lp[j] + q]);
            if (with_k)
                for (int e = S.KLp[j]; e < S.KLp[j + 1]; e++)
                {
                    const int slot = S.KLslot[e], vi = S.Kvidx[slot];
                    H.fa.push_back(vi);
                    H.fa.push_back(S.KLpos[e]);
                    if (vi < 0)
                        H.fa_val.push_back(S.Kshared[slot]);
                }
            for (size_t g = 0; g < rows.size(); g += FA_GROUP)
            {
                const size_t ge = std::min(rows.size(), g + FA_GROUP);
                for (size_t r = g; r < ge; r++)
                {
                    const int t = rows[r], k = S.Lr.j[t];
                    H.fa.push_back(k);
                    H.fa.push_back(H.fw_pos[t]);
                    H.fa.push_back(S.Lp[k + 1] - S.upd_tail[t]);
                }
                for (size_t r = g; r < ge; r++)
                {
                    const int t = rows[r], k = S.Lr.j[t];
                    const int u0 = S.upd_tail[t], len = S.Lp[k + 1] - u0;
                    for (int q = 0; q < len; q++)
                    {
                        H.fa.push_back(S.upd_rel[S.upd_rel_p[t] + q]);
                        H.fa.push_back(H.bw_pos[u0 + q]);
                    }
                }
            }
            for (int q = 0; q < cnt; q++)
            {
                const int u = S.Lp[j] + q;
                H.fa.push_back(H.bw_pos[u]);
                H.fa.push_back(H.fw_pos[Lcsr[u]]);
            }
        };
        auto emit_phase = [&](const std::vector<std::vector<std::pair<int, std::pair<int, ivec>>>> &pw) {
            for (int w = 0; w < W; w++)
            {
                align_chunk(H.fa);
                align_chunk(H.fa_val);
                H.fa_seg.insert(H.fa_seg.end(), {(int)H.fa.size(), (int)pw[w].size(), (int)H.fa_val.size()});
                for (const auto &tk : pw[w])
                    emit_task(tk.first, tk.second.first, tk.second.second);
            }
            H.nph_fa++;
        };
        auto all_rows = [&](int j) {
            ivec r;
            for (int t = S.Lr.p[j]; t < S.Lr.p[j + 1]; t++)
                r.push_back(t);
            return r;
        };
        for (int ph = 0; ph < nph; ph++)
        {
            const Phase &f = S.phases[ph];
            std::vector<std::vector<std::pair<int, std::pair<int, ivec>>>> pw(W);
            if (f.parallel)
            {
                for (int w = 0; w < W; w++)
                    for (int j : worker_tasks(S, f, w, W, false))
                        pw[w].push_back({j, {0, all_rows(j)}});
                emit_phase(pw);
                continue;
            }
            const ivec tk = worker_tasks(S, f, 0, W, false);
            for (size_t a = 0; a < tk.size(); a++)
                where[tk[a]] = (int)a;
            std::vector<char> has_ext(tk.size(), 0);
            int rr = 0;
            for (size_t a = 0; a < tk.size(); a++)
            {
                ivec ext;
                for (int t = S.Lr.p[tk[a]]; t < S.Lr.p[tk[a] + 1]; t++)
                    if (where[S.Lr.j[t]] < 0)
                        ext.push_back(t);
                if (!ext.empty())
                {
                    has_ext[a] = 1;
                    pw[rr++ % W].push_back({tk[a], {1, ext}});
                }
            }
            if (rr > 0)
                emit_phase(pw);
            std::vector<std::vector<std::pair<int, std::pair<int, ivec>>>> cw(W);
            for (size_t a = 0; a < tk.size(); a++)
            {
                ivec in;
                for (int t = S.Lr.p[tk[a]]; t < S.Lr.p[tk[a] + 1]; t++)
                    if (where[S.Lr.j[t]] >= 0)
                        in.push_back(t);
                cw[0].push_back({tk[a], {has_ext[a] ? 2 : 3, in}});
            }
            emit_phase(cw);
            for (int j : tk)
                where[j] = -1;
        }
    }
    pad_tail(H.fa);
    pad_tail(H.fa_val);

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

void refresh_stream_values(const Symbolic &S, HostStreams &H)
{
    HostStreams fresh;
    build_streams(S, H.workers, fresh);
    H.fa_val.swap(fresh.fa_val);
    H.rx_val.swap(fresh.rx_val);
    H.ry_val.swap(fresh.ry_val);
    H.rz_val.swap(fresh.rz_val);
    H.rc_val.swap(fresh.rc_val);
}

} // namespace eicos
