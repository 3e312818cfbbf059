// Host-side symbolic analysis; see symbolic.hpp for the reference sites each part replaces.
#include "symbolic.hpp"

#include <algorithm>
#include <cmath>
#include <numeric>
#include <stdexcept>

namespace eicos
{

static CsrView make_csr(const Csc &a)
{
    CsrView r;
    r.p.assign(a.rows + 1, 0);
    r.j.resize(a.nnz());
    r.v.resize(a.nnz());
    for (int k = 0; k < a.nnz(); k++)
        r.p[a.i[k] + 1]++;
    std::partial_sum(r.p.begin(), r.p.end(), r.p.begin());
    ivec fill(r.p.begin(), r.p.end() - 1);
    for (int c = 0; c < a.cols; c++)
        for (int k = a.p[c]; k < a.p[c + 1]; k++)
        {
            const int t = fill[a.i[k]]++;
            r.j[t] = c;
            r.v[t] = k;
        }
    return r;
}

// src/eicos.cpp:302-362.  Three rounds; every round scales rows of A and G by 1/sqrt(row inf-norm)
// (all rows of one cone share the SUM of their norms) and columns by 1/sqrt(col inf-norm over both
// matrices); norms below 1e-6 leave the row/column alone.
void equilibrate(Csc &G, Csc &A, int l, const ivec &q, dvec &xeq, dvec &Aeq, dvec &Geq)
{
    const int n = std::max(G.cols, A.cols), p = A.rows, m = G.rows;
    xeq.assign(n, 1.0);
    Aeq.assign(p, 1.0);
    Geq.assign(m, 1.0);
    dvec cs(n), ra(p), rg(m);
    const auto root = [](double v) { return std::fabs(v) < 1e-6 ? 1.0 : std::sqrt(v); };
    for (int round = 0; round < Settings::equil_iters; round++)
    {
        std::fill(cs.begin(), cs.end(), 0.0);
        std::fill(ra.begin(), ra.end(), 0.0);
        std::fill(rg.begin(), rg.end(), 0.0);
        for (int j = 0; j < A.cols; j++)
            for (int k = A.p[j]; k < A.p[j + 1]; k++)
            {
                const double a = std::fabs(A.x[k]);
                cs[j] = std::max(a, cs[j]);
                ra[A.i[k]] = std::max(a, ra[A.i[k]]);
            }
        for (int j = 0; j < G.cols; j++)
            for (int k = G.p[j]; k < G.p[j + 1]; k++)
            {
                const double a = std::fabs(G.x[k]);
                cs[j] = std::max(a, cs[j]);
                rg[G.i[k]] = std::max(a, rg[G.i[k]]);
            }
        int at = l;
        for (int d : q)
        {
            double tot = 0.0;
            for (int k = 0; k < d; k++)
                tot += rg[at + k];
            std::fill(rg.begin() + at, rg.begin() + at + d, tot);
            at += d;
        }
        for (double &v : cs)
            v = root(v);
        for (double &v : ra)
            v = root(v);
        for (double &v : rg)
            v = root(v);
        // rows first, then columns: two separate divisions per entry, as the reference does
        for (int k = 0; k < A.nnz(); k++)
            A.x[k] /= ra[A.i[k]];
        for (int k = 0; k < G.nnz(); k++)
            G.x[k] /= rg[G.i[k]];
        for (int j = 0; j < A.cols; j++)
            for (int k = A.p[j]; k < A.p[j + 1]; k++)
                A.x[k] /= cs[j];
        for (int j = 0; j < G.cols; j++)
            for (int k = G.p[j]; k < G.p[j + 1]; k++)
                G.x[k] /= cs[j];
        for (int k = 0; k < n; k++)
            xeq[k] *= cs[k];
        for (int k = 0; k < p; k++)
            Aeq[k] *= ra[k];
        for (int k = 0; k < m; k++)
            Geq[k] *= rg[k];
    }
}

// KKT pattern [dI A' G'; . -dI 0; . . -V] (upper), with the sparse SOC expansion.
static void build_kkt(Symbolic &S)
{
    const int n = S.n, p = S.p, l = S.l, N = S.N;
    // Column by column, rows ascending: x-diagonal | y columns (row of A, then diag) | z columns
    // (row of G, then the scaling block).  This is what setFromTriplets yields for the triplets
    // pushed at src/eicos.cpp:1765-1876.
    S.Kp.assign(1, 0);
    S.Ki.clear();
    S.Kvidx.clear();
    S.AGsrc.clear();
    ivec ag_slot_A(S.A.nnz()), ag_slot_G(S.G.nnz());
    std::vector<int> kind; // -1 shared const; else V index (filled below through Vslot)
    struct Entry
    {
        int row, src; // src: 0 = +delta, 1 = -delta, 2 = A value, 3 = G value, 4 = scaling (V)
        int idx;
    };
    std::vector<std::vector<Entry>> cols(N);
    for (int k = 0; k < n; k++)
        cols[k].push_back({k, 0, 0});
    for (int r = 0; r < p; r++)
    {
        for (int t = S.Ar.p[r]; t < S.Ar.p[r + 1]; t++)
            cols[n + r].push_back({S.Ar.j[t], 2, S.Ar.v[t]});
        cols[n + r].push_back({n + r, 1, 0});
    }
    for (int i = 0; i < S.m; i++)
    {
        const int c = n + p + S.zk[i];
        for (int t = S.Gr.p[i]; t < S.Gr.p[i + 1]; t++)
            cols[c].push_back({S.Gr.j[t], 3, S.Gr.v[t]});
    }
    // scaling block: remember (row, col) of every V entry in cacheIndices order
    std::vector<std::pair<int, int>> vpos;
    int d = n + p;
    for (int k = 0; k < l; k++, d++)
        vpos.push_back({d, d});
    for (int dim : S.q)
    {
        const int s = d;
        for (int k = 0; k < dim; k++)
            vpos.push_back({s + k, s + k});
        vpos.push_back({s + dim, s + dim});
        for (int k = 1; k < dim; k++)
            vpos.push_back({s + k, s + dim});
        vpos.push_back({s + dim + 1, s + dim + 1});
        for (int k = 0; k < dim; k++)
            vpos.push_back({s + k, s + dim + 1});
        d = s + dim + 2;
    }
    for (size_t v = 0; v < vpos.size(); v++)
        cols[vpos[v].second].push_back({vpos[v].first, 4, (int)v});
    S.Vslot.assign(vpos.size(), -1);
    for (int c = 0; c < N; c++)
    {
        std::stable_sort(cols[c].begin(), cols[c].end(), [](const Entry &a, const Entry &b) { return a.row < b.row; });
        for (const Entry &e : cols[c])
        {
            const int slot = (int)S.Ki.size();
            S.Ki.push_back(e.row);
            S.Kvidx.push_back(e.src == 4 ? e.idx : -1);
            if (e.src == 2)
                ag_slot_A[e.idx] = slot;
            else if (e.src == 3)
                ag_slot_G[e.idx] = slot;
            else if (e.src == 4)
                S.Vslot[e.idx] = slot;
            kind.push_back(e.src);
        }
        S.Kp.push_back((int)S.Ki.size());
    }
    // AGslot in the reference's traversal order: columns of At (= rows of A), then columns of Gt.
    S.AGslot.clear();
    for (int r = 0; r < p; r++)
        for (int t = S.Ar.p[r]; t < S.Ar.p[r + 1]; t++)
        {
            S.AGslot.push_back(ag_slot_A[S.Ar.v[t]]);
            S.AGsrc.push_back(S.Ar.v[t]);
        }
    for (int i = 0; i < S.m; i++)
        for (int t = S.Gr.p[i]; t < S.Gr.p[i + 1]; t++)
        {
            S.AGslot.push_back(ag_slot_G[S.Gr.v[t]]);
            S.AGsrc.push_back(-S.Gr.v[t] - 1);
        }
    S.Kshared.assign(S.Ki.size(), 0.0);
    for (size_t s = 0; s < kind.size(); s++)
        if (kind[s] == 0)
            S.Kshared[s] = Settings::deltastat;
        else if (kind[s] == 1)
            S.Kshared[s] = -Settings::deltastat;
}

static void fill_shared_values(Symbolic &S)
{
    for (size_t t = 0; t < S.AGslot.size(); t++)
    {
        const int src = S.AGsrc[t];
        S.Kshared[S.AGslot[t]] = src >= 0 ? S.A.x[src] : S.G.x[-src - 1];
    }
}

// Eigen analyzePattern: symmetric pattern -> AMD -> permuted upper pattern -> etree + counts,
// then the column structure of L (rows ascending, which is how the up-looking factorisation
// of the reference fills them).
static void order_and_factor_pattern(Symbolic &S)
{
    const int N = S.N;
    ivec sp(N + 1, 0);
    for (int j = 0; j < N; j++)
        for (int k = S.Kp[j]; k < S.Kp[j + 1]; k++)
        {
            sp[j + 1]++;
            if (S.Ki[k] != j)
                sp[S.Ki[k] + 1]++;
        }
    std::partial_sum(sp.begin(), sp.end(), sp.begin());
    ivec si(sp[N]), fill(sp.begin(), sp.end() - 1);
    for (int j = 0; j < N; j++)
        for (int k = S.Kp[j]; k < S.Kp[j + 1]; k++)
        {
            const int i = S.Ki[k];
            si[fill[j]++] = i;
            if (i != j)
                si[fill[i]++] = j;
        }
    S.pinv = amd_ordering(N, sp, si);
    S.P.assign(N, 0);
    for (int k = 0; k < N; k++)
        S.P[S.pinv[k]] = k;

    // lower-triangular permuted KKT by columns: entry (max(ip,jp), min(ip,jp)) lives in column min
    S.KLp.assign(N + 1, 0);
    for (int j = 0; j < N; j++)
        for (int k = S.Kp[j]; k < S.Kp[j + 1]; k++)
            S.KLp[std::min(S.P[S.Ki[k]], S.P[j]) + 1]++;
    std::partial_sum(S.KLp.begin(), S.KLp.end(), S.KLp.begin());
    S.KLslot.assign(S.Ki.size(), 0);
    ivec klrow(S.Ki.size(), 0);
    fill.assign(S.KLp.begin(), S.KLp.end() - 1);
    for (int j = 0; j < N; j++)
        for (int k = S.Kp[j]; k < S.Kp[j + 1]; k++)
        {
            const int a = S.P[S.Ki[k]], b = S.P[j];
            const int t = fill[std::min(a, b)]++;
            S.KLslot[t] = k;
            klrow[t] = std::max(a, b);
        }

    // elimination tree (Liu) from the rows of the lower pattern == columns of the upper one.
    // Upper column k holds rows i<k  <=>  lower entries (k, i): walk i -> root, stop at flag k.
    ivec up(N + 1, 0), ui;
    for (size_t t = 0; t < klrow.size(); t++)
        up[klrow[t] + 1]++;
    std::partial_sum(up.begin(), up.end(), up.begin());
    ui.assign(klrow.size(), 0);
    fill.assign(up.begin(), up.end() - 1);
    for (int c = 0; c < N; c++)
        for (int t = S.KLp[c]; t < S.KLp[c + 1]; t++)
            ui[fill[klrow[t]]++] = c;
    S.parent.assign(N, -1);
    ivec flag(N, -1), count(N, 0);
    std::vector<ivec> rows_of_col(N);
    for (int k = 0; k < N; k++)
    {
        flag[k] = k;
        for (int t = up[k]; t < up[k + 1]; t++)
            for (int i = ui[t]; i < k && flag[i] != k; i = S.parent[i])
            {
                if (S.parent[i] == -1)
                    S.parent[i] = k;
                count[i]++;
                rows_of_col[i].push_back(k); // k ascending over the outer loop => sorted
                flag[i] = k;
            }
    }
    S.Lp.assign(N + 1, 0);
    for (int k = 0; k < N; k++)
        S.Lp[k + 1] = S.Lp[k] + count[k];
    S.nnzL = S.Lp[N];
    S.Li.resize(S.nnzL);
    for (int k = 0; k < N; k++)
        std::copy(rows_of_col[k].begin(), rows_of_col[k].end(), S.Li.begin() + S.Lp[k]);

    // rows of L
    S.Lr.p.assign(N + 1, 0);
    for (int t = 0; t < S.nnzL; t++)
        S.Lr.p[S.Li[t] + 1]++;
    std::partial_sum(S.Lr.p.begin(), S.Lr.p.end(), S.Lr.p.begin());
    S.Lr.j.assign(S.nnzL, 0);
    S.Lr.v.assign(S.nnzL, 0);
    fill.assign(S.Lr.p.begin(), S.Lr.p.end() - 1);
    for (int c = 0; c < N; c++)
        for (int t = S.Lp[c]; t < S.Lp[c + 1]; t++)
        {
            const int u = fill[S.Li[t]]++;
            S.Lr.j[u] = c;
            S.Lr.v[u] = t;
        }

    // position of every lower-K entry inside its L column
    S.KLpos.assign(S.KLslot.size(), -1);
    for (int c = 0; c < N; c++)
        for (int t = S.KLp[c]; t < S.KLp[c + 1]; t++)
        {
            if (klrow[t] == c)
                continue;
            const int *b = S.Li.data() + S.Lp[c], *e = S.Li.data() + S.Lp[c + 1];
            const int *f = std::lower_bound(b, e, klrow[t]);
            if (f == e || *f != klrow[t])
                throw std::logic_error("KKT entry outside the pattern of L");
            S.KLpos[t] = (int)(f - b);
        }
}

// elimination-tree height, widest column of L, multiply-adds of one numeric factorisation
static void build_shape(Symbolic &S)
{
    const int N = S.N;
    S.level.assign(N, 0);
    for (int j = 0; j < N; j++)
        if (S.parent[j] >= 0)
            S.level[S.parent[j]] = std::max(S.level[S.parent[j]], S.level[j] + 1);
    S.height = N ? *std::max_element(S.level.begin(), S.level.end()) + 1 : 0;
    S.maxcol = 0;
    S.fma_count = 0;
    for (int j = 0; j < N; j++)
    {
        const long long c = S.Lp[j + 1] - S.Lp[j];
        S.maxcol = std::max(S.maxcol, (int)c);
        S.fma_count += c * (c + 1) / 2; // Schur updates of column j (right-looking)
    }
}

void analyze(Symbolic &S, int n, int m, int p, int ncones, const int *q,
             const double *Gpr, const int *Gjc, const int *Gir,
             const double *Apr, const int *Ajc, const int *Air)
{
    // a NULL triple means "matrix absent" (src/eicos.cpp:103-117); then the matching dimension is 0
    const bool hasG = Gpr && Gjc && Gir, hasA = Apr && Ajc && Air;
    S.n = n;
    S.m = hasG ? m : 0;
    S.p = hasA ? p : 0;
    if (n < 0 || m < 0 || p < 0 || ncones < 0)
        throw std::invalid_argument("negative dimension");
    if (hasG && ncones > 0 && !q)
        throw std::invalid_argument("cone dimensions missing");
    // column pointers come from the caller: they must start at 0 and never decrease before they are walked
    for (const int *jc : {hasG ? Gjc : nullptr, hasA ? Ajc : nullptr})
        if (jc)
        {
            if (jc[0] != 0)
                throw std::invalid_argument("CSC column pointers must start at 0");
            for (int j = 0; j < n; j++)
                if (jc[j] > jc[j + 1])
                    throw std::invalid_argument("CSC column pointers must be non-decreasing");
        }
    S.q.assign(q, q + (hasG ? ncones : 0));
    S.nc = (int)S.q.size();
    int sumq = 0;
    for (int d : S.q)
    {
        if (d < 1)
            throw std::invalid_argument("second-order cone of dimension < 1");
        sumq += d;
    }
    S.l = S.m - sumq; // the caller's `l` is ignored, like the reference (src/eicos.cpp:91,155)
    if (S.l < 0)
        throw std::invalid_argument("cone dimensions exceed the number of inequality rows");
    S.N = S.n + S.p + S.m + 2 * S.nc;
    S.mt = S.m + 2 * S.nc;
    S.G = Csc();
    S.A = Csc();
    S.G.rows = S.m;
    S.G.cols = n;
    S.G.p.assign(n + 1, 0);
    S.A.rows = S.p;
    S.A.cols = n;
    S.A.p.assign(n + 1, 0);
    if (hasG)
    {
        S.G.p.assign(Gjc, Gjc + n + 1);
        S.G.i.assign(Gir, Gir + Gjc[n]);
        S.G.x.assign(Gpr, Gpr + Gjc[n]);
    }
    if (hasA)
    {
        S.A.p.assign(Ajc, Ajc + n + 1);
        S.A.i.assign(Air, Air + Ajc[n]);
        S.A.x.assign(Apr, Apr + Ajc[n]);
    }
    for (const Csc *M : {&S.G, &S.A})
        for (int j = 0; j < M->cols; j++)
            for (int k = M->p[j]; k < M->p[j + 1]; k++)
                if (M->i[k] < 0 || M->i[k] >= M->rows || (k > M->p[j] && M->i[k] <= M->i[k - 1]))
                    throw std::invalid_argument("CSC row indices must be in range and strictly ascending per column");

    S.cone_z.clear();
    S.cone_k.clear();
    S.cone_q.clear();
    S.zk.resize(S.m);
    for (int i = 0; i < S.l; i++)
        S.zk[i] = i;
    int z = S.l, k = S.l, qo = 0;
    for (int d : S.q)
    {
        S.cone_z.push_back(z);
        S.cone_k.push_back(k);
        S.cone_q.push_back(qo);
        for (int t = 0; t < d; t++)
            S.zk[z + t] = k + t;
        z += d;
        k += d + 2;
        qo += d - 1;
    }
    S.qtot = qo;

    equilibrate(S.G, S.A, S.l, S.q, S.xeq, S.Aeq, S.Geq);
    S.Gr = make_csr(S.G);
    S.Ar = make_csr(S.A);
    build_kkt(S);
    fill_shared_values(S);
    order_and_factor_pattern(S);
    build_shape(S);
}

void unequilibrate(Symbolic &S)
{
    for (int j = 0; j < S.A.cols; j++)
        for (int k = S.A.p[j]; k < S.A.p[j + 1]; k++)
            S.A.x[k] *= S.Aeq[S.A.i[k]] * S.xeq[j];
    for (int j = 0; j < S.G.cols; j++)
        for (int k = S.G.p[j]; k < S.G.p[j + 1]; k++)
            S.G.x[k] *= S.Geq[S.G.i[k]] * S.xeq[j];
}

void refresh_values(Symbolic &S, const double *Gpr, const double *Apr)
{
    if (Gpr)
        std::copy(Gpr, Gpr + S.G.nnz(), S.G.x.begin());
    if (Apr)
        std::copy(Apr, Apr + S.A.nnz(), S.A.x.begin());
    equilibrate(S.G, S.A, S.l, S.q, S.xeq, S.Aeq, S.Geq);
    fill_shared_values(S);
}

} // namespace eicos
