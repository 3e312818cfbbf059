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

// tasks of worker w in a phase, in the order the device walks them
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
} // namespace

void build_streams(const Symbolic &S, int W, HostStreams &H)
{
    H = HostStreams();
    H.workers = W;
    const int N = S.N, nph = (int)S.phases.size();

    // ---- storage order of the factor values = order of use
    H.fw_base.assign(N, 0);
    H.bw_base.assign(N, 0);
    int pos = 0;
    for (int ph = 0; ph < nph; ph++)
        for (int w = 0; w < W; w++)
            for (int i : worker_tasks(S, S.phases[ph], w, W, false))
            {
                H.fw_base[i] = pos;
                pos += S.Lr.p[i + 1] - S.Lr.p[i];
            }
    if (pos != S.nnzL)
        throw std::logic_error("forward stream does not cover L");
    pos = 0;
    for (int ph = nph - 1; ph >= 0; ph--)
        for (int w = 0; w < W; w++)
            for (int j : worker_tasks(S, S.phases[ph], w, W, true))
            {
                H.bw_base[j] = pos;
                pos += S.Lp[j + 1] - S.Lp[j];
            }
    if (pos != S.nnzL)
        throw std::logic_error("backward stream does not cover L");

    // ---- forward sweep: rows of L
    H.fw_seg.assign((size_t)nph * W * 3, 0);
    for (int ph = 0; ph < nph; ph++)
        for (int w = 0; w < W; w++)
        {
            align_chunk(H.fw);
            const ivec tk = worker_tasks(S, S.phases[ph], w, W, false);
            int *seg = &H.fw_seg[((size_t)ph * W + w) * 3];
            seg[0] = (int)H.fw.size();
            seg[1] = (int)tk.size();
            seg[2] = tk.empty() ? 0 : H.fw_base[tk[0]];
            const bool serial = !S.phases[ph].parallel;
            auto header = [&](size_t a) {
                const int i = tk[a];
                H.fw.push_back(i);
                H.fw.push_back(S.pinv[i]);
                H.fw.push_back(S.Lr.p[i + 1] - S.Lr.p[i]);
            };
            if (!tk.empty())
                header(0);
            for (size_t a = 0; a < tk.size(); a++)
            { // the header of task a+1 precedes the entries of task a (software pipelining on the device)
                if (a + 1 < tk.size())
                    header(a + 1);
                const int i = tk[a];
                for (int t = S.Lr.p[i]; t < S.Lr.p[i + 1]; t++)
                {
                    const int c = S.Lr.j[t];
                    if (serial && a >= 1 && c == tk[a - 1])
                        H.fw.push_back(FWD_PREV1);
                    else if (serial && a >= 2 && c == tk[a - 2])
                        H.fw.push_back(FWD_PREV2);
                    else
                        H.fw.push_back(c);
                }
            }
        }
    pad_tail(H.fw);

    // ---- backward sweep: columns of L, results land in KKT order (row pinv[j])
    H.bw_seg.assign((size_t)nph * W * 3, 0);
    for (int ph = nph - 1; ph >= 0; ph--)
        for (int w = 0; w < W; w++)
        {
            align_chunk(H.bw);
            const ivec tk = worker_tasks(S, S.phases[ph], w, W, true);
            int *seg = &H.bw_seg[((size_t)ph * W + w) * 3];
            seg[0] = (int)H.bw.size();
            seg[1] = (int)tk.size();
            seg[2] = tk.empty() ? 0 : H.bw_base[tk[0]];
            const bool serial = !S.phases[ph].parallel;
            auto header = [&](size_t a) {
                const int j = tk[a];
                H.bw.push_back(j);
                H.bw.push_back(S.pinv[j]);
                H.bw.push_back(S.Lp[j + 1] - S.Lp[j]);
            };
            if (!tk.empty())
                header(0);
            for (size_t a = 0; a < tk.size(); a++)
            {
                if (a + 1 < tk.size())
                    header(a + 1);
                const int j = tk[a];
                for (int u = S.Lp[j]; u < S.Lp[j + 1]; u++)
                {
                    const int r = S.Li[u];
                    if (serial && a >= 1 && r == tk[a - 1])
                        H.bw.push_back(FWD_PREV1);
                    else if (serial && a >= 2 && r == tk[a - 2])
                        H.bw.push_back(FWD_PREV2);
                    else
                        H.bw.push_back(S.pinv[r]);
                }
            }
        }
    pad_tail(H.bw);

    // ---- numeric factorisation, left-looking by column
    ivec Lcsr(S.nnzL);
    for (int t = 0; t < S.nnzL; t++)
        Lcsr[S.Lr.v[t]] = t;
    H.fa_seg.assign((size_t)nph * W * 3, 0);
    for (int ph = 0; ph < nph; ph++)
        for (int w = 0; w < W; w++)
        {
            align_chunk(H.fa);
            align_chunk(H.fa_val);
            const ivec tk = worker_tasks(S, S.phases[ph], w, W, false);
            int *seg = &H.fa_seg[((size_t)ph * W + w) * 3];
            seg[0] = (int)H.fa.size();
            seg[1] = (int)tk.size();
            seg[2] = (int)H.fa_val.size();
            for (int j : tk)
            {
                const int cnt = S.Lp[j + 1] - S.Lp[j];
                H.fa.push_back(j);
                H.fa.push_back(cnt);
                H.fa.push_back(S.KLp[j + 1] - S.KLp[j]);
                H.fa.push_back(S.Lr.p[j + 1] - S.Lr.p[j]);
                H.fa.push_back(H.bw_base[j]);
                H.fa.push_back(H.fw_base[j]);
                for (int e = S.KLp[j]; e < S.KLp[j + 1]; e++)
                {
                    const int slot = S.KLslot[e], vi = S.Kvidx[slot];
                    H.fa.push_back(vi);
                    H.fa.push_back(S.KLpos[e]);
                    if (vi < 0)
                        H.fa_val.push_back(S.Kshared[slot]);
                }
                for (int t = S.Lr.p[j]; t < S.Lr.p[j + 1]; t++)
                {
                    const int k = S.Lr.j[t];
                    const int u0 = S.upd_tail[t], len = S.Lp[k + 1] - u0;
                    H.fa.push_back(k);
                    H.fa.push_back(H.bw_base[k] + (u0 - S.Lp[k]));
                    H.fa.push_back(len);
                    for (int r = 0; r < len; r++)
                        H.fa.push_back(S.upd_rel[S.upd_rel_p[t] + r]);
                }
                for (int q = 0; q < cnt; q++)
                {
                    const int u = S.Lp[j] + q, t = Lcsr[u], row = S.Li[u];
                    H.fa.push_back(H.fw_base[row] + (t - S.Lr.p[row]));
                }
            }
        }
    pad_tail(H.fa);
    pad_tail(H.fa_val);

    // ---- mat-vec row sets (K-space gather indices)
    const int n = S.n, p = S.p, zb = S.n + S.p;
    auto rowset = [&](int rows, ivec &s, dvec &v, ivec &seg, auto &&emit) {
        seg.assign((size_t)W * 2, 0);
        for (int w = 0; w < W; w++)
        {
            align_chunk(s);
            align_chunk(v);
            seg[w * 2] = (int)s.size();
            seg[w * 2 + 1] = (int)v.size();
            for (int r = w; r < rows; r += W)
                emit(r, s, v);
        }
        pad_tail(s);
        pad_tail(v);
    };
    rowset(n, H.rx, H.rx_val, H.rx_seg, [&](int j, ivec &s, dvec &v) {
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
    rowset(p, H.ry, H.ry_val, H.ry_seg, [&](int i, ivec &s, dvec &v) {
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
    rowset(S.l, H.rz, H.rz_val, H.rz_seg, grow);
    rowset(S.nc, H.rc, H.rc_val, H.rc_seg, [&](int c, ivec &s, dvec &v) {
        s.push_back(S.q[c]);
        s.push_back(S.cone_k[c]);
        s.push_back(S.cone_q[c]);
        for (int k = 0; k < S.q[c]; k++)
            grow(S.cone_z[c] + k, s, v);
    });
    (void)p;
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
