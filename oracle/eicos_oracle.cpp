/*
 * eicos_oracle.cpp - CPU oracle for the EiCOS Solver::solve hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Loaded by tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / `--impl reference` legs, never by the product.
 *
 * What it restates (file:line into /root/reference unless noted):
 *   - EiCOS itself: src/eicos.cpp (build :132-187, equilibration :256-404, NT scalings
 *     :411-507, exit tests :526-641, residuals/statistics :643-754, bringToCone :761-805,
 *     KKT scalings :807-846/:1691-1732, main loop :848-1262, RHS/cone algebra :1282-1378,
 *     line search :1380-1469, solveKKT :1471-1620, scale2add :1629-1662, RHSaffine
 *     :1670-1689, setupKKT/cacheIndices :1734-1988, updateKKTAG :1990-2030, updateData
 *     :2032-2082) and include/eicos.hpp (types :8-114).
 *   - Eigen (third party, ">= 3.3", NOT in /root/reference; restated from the published
 *     algorithms it implements): AMDOrdering = T. Davis' CSparse cs_amd with Eigen's
 *     "keep the diagonal" variant; SimplicialLDLT = Davis' LDL (ldl_symbolic/ldl_numeric)
 *     with the symmetric permutation applied by twistedBy(); the triangular solves.
 *     Call sites in the reference: src/eicos.cpp:897,900,1164,1477,1599.
 *
 * PARITY: "parity unpinned" against Eigen's own output (it cannot be built here); pinned
 * against the reference tests' exit flags, HiGHS objectives and closed-form SOCPs
 * (tests/test_oracle.py).
 */
#include "eicos_oracle.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>
#include <thread>
#include <vector>

/* -DORA_LONG_DOUBLE builds the same algorithm in 80-bit extended precision (tests/test_oracle.py uses it to
 * see which way a knife-edge decision of the double-precision run would go with 11 more bits); the C API
 * below stays double and converts at the boundary. */
typedef double ora_api_real; /* what crosses the C API, whatever the arithmetic type is */
#ifdef ORA_LONG_DOUBLE
#define double long double
#endif

namespace ora
{
/* std::max over mixed literal / variable types (the long double build) */
template <class A, class B>
inline auto rmax(A a, B b) -> decltype(a + b)
{
    typedef decltype(a + b) T;
    return std::max<T>(a, b);
}

using std::size_t;
typedef std::vector<double> vec;
typedef std::vector<int> ivec;

/* ------------------------------------------------------------------ settings */
/* include/eicos.hpp:23-47 (all const except verbose) */
static const double GAMMA = 0.99, DELTASTAT = 7e-8;
static const double FEASTOL = 1e-8, ABSTOL = 1e-8, RELTOL = 1e-8;
static const double FEASTOL_INACC = 1e-4, ABSTOL_INACC = 5e-5, RELTOL_INACC = 5e-5;
static const int NITREF = 9, EQUIL_ITERS = 3, ITER_MAX = 100;
// test hook (ora_debug_set_iter_max): a cap on the iterations of every solve in this process, 0 = ITER_MAX.  The
// reference's iter_max is a compile-time constant (include/eicos.hpp:45); capping it exposes the iterate after k
// iterations (src/eicos.cpp:1082-1106) to the per-iteration parity tests.
static int g_debug_iter_max = 0;
static const double LINSYSACC = 1e-14, IRERRFACT = 6, STEPMIN = 1e-6, STEPMAX = 0.999;
static const double SIGMAMIN = 1e-4, SIGMAMAX = 1.0, SAFEGUARD = 500;

enum
{
    OPTIMAL = 0,
    PINF = 1,
    DINF = 2,
    MAXIT = -1,
    NUMERICS = -2,
    OUTCONE = -3,
    FATAL = -7,
    INACC = 10,
    NOT_CONVERGED = -87
}; /* include/eicos.hpp:8-21 */

/* ------------------------------------------------------------------ sparse */
struct Csc
{
    int rows = 0, cols = 0;
    ivec p{0}, i;
    vec x;
    int nnz() const { return (int)i.size(); }
};

/* Eigen's `B = A.transpose()` for column-major A: counting transpose, inner indices sorted. */
static Csc transpose(const Csc &a)
{
    Csc t;
    t.rows = a.cols;
    t.cols = a.rows;
    t.p.assign(a.rows + 1, 0);
    t.i.resize(a.nnz());
    t.x.resize(a.nnz());
    for (int k = 0; k < a.nnz(); k++)
        t.p[a.i[k] + 1]++;
    for (int r = 0; r < a.rows; r++)
        t.p[r + 1] += t.p[r];
    ivec next(t.p.begin(), t.p.end() - 1);
    for (int j = 0; j < a.cols; j++)
        for (int k = a.p[j]; k < a.p[j + 1]; k++)
        {
            int q = next[a.i[k]]++;
            t.i[q] = j;
            t.x[q] = a.x[k];
        }
    return t;
}

/* y (+)= alpha * A * x, column-major accumulation like Eigen's sparse*dense kernel */
static void spmv_add(const Csc &a, const double *x, double *y, double alpha)
{
    for (int j = 0; j < a.cols; j++)
    {
        const double xj = alpha * x[j];
        for (int k = a.p[j]; k < a.p[j + 1]; k++)
            y[a.i[k]] += a.x[k] * xj;
    }
}

static double norm2(const double *v, size_t n)
{
    double s = 0;
    for (size_t k = 0; k < n; k++)
        s += v[k] * v[k];
    return std::sqrt(s);
}
static double sqnorm(const double *v, size_t n)
{
    double s = 0;
    for (size_t k = 0; k < n; k++)
        s += v[k] * v[k];
    return s;
}
static double dot(const double *a, const double *b, size_t n)
{
    double s = 0;
    for (size_t k = 0; k < n; k++)
        s += a[k] * b[k];
    return s;
}
static double norminf(const double *v, size_t n)
{
    double s = 0; /* Eigen lpNorm<Infinity> returns 0 for an empty vector */
    for (size_t k = 0; k < n; k++)
        s = rmax(s, std::fabs(v[k]));
    return s;
}

/* ------------------------------------------------------------------ AMD */
/*
 * Approximate minimum degree ordering as Eigen's AMDOrdering runs it on
 * C = K + K' (full symmetric pattern, diagonal KEPT, inner indices ascending).
 * This is CSparse's cs_amd (Davis, "Direct Methods for Sparse Linear Systems",
 * ch. 7) with Eigen's changes: dense threshold max(16,10*sqrt(n)) clipped to
 * n-2; diagonal entries stay in the graph, so a node is "empty" when its
 * degree is 1 and it has a diagonal, and nodes without a structural diagonal
 * are treated as dense.  Output: perm[k] = original index of the k-th pivot.
 */
static inline int flip(int i) { return -i - 2; }

static int wclear(int mark, int lemax, int *w, int n)
{
    if (mark < 2 || mark + lemax < 0)
    {
        for (int k = 0; k < n; k++)
            if (w[k] != 0)
                w[k] = 1;
        mark = 2;
    }
    return mark;
}

static int tdfs(int j, int k, int *head, const int *next, int *post, int *stack)
{
    int top = 0;
    stack[0] = j;
    while (top >= 0)
    {
        const int p = stack[top];
        const int i = head[p];
        if (i == -1)
        {
            top--;
            post[k++] = p;
        }
        else
        {
            head[p] = next[i];
            stack[++top] = i;
        }
    }
    return k;
}

static ivec amd_order(int n, const ivec &Ap, const ivec &Ai)
{
    ivec perm(n + 1, 0);
    if (n == 0)
    {
        perm.resize(0);
        return perm;
    }
    int dense = rmax(16, (int)(10 * std::sqrt((double)n)));
    dense = std::min(n - 2, dense);

    int cnz = Ap[n];
    const int nzmax = cnz + cnz / 5 + 2 * n;
    ivec Cp(Ap.begin(), Ap.end()); /* n+1 entries */
    ivec Ci(nzmax, 0);
    std::copy(Ai.begin(), Ai.begin() + cnz, Ci.begin());

    ivec W(8 * (size_t)(n + 1), 0);
    int *len = &W[0], *nv = len + (n + 1), *next = nv + (n + 1), *head = next + (n + 1);
    int *elen = head + (n + 1), *degree = elen + (n + 1), *w = degree + (n + 1), *hhead = w + (n + 1);
    int *last = perm.data();

    for (int k = 0; k < n; k++)
        len[k] = Cp[k + 1] - Cp[k];
    len[n] = 0;
    for (int i = 0; i <= n; i++)
    {
        head[i] = last[i] = next[i] = hhead[i] = -1;
        nv[i] = 1;
        w[i] = 1;
        elen[i] = 0;
        degree[i] = len[i];
    }
    int mark = wclear(0, 0, w, n);
    int nel = 0, mindeg = 0, lemax = 0;

    /* initial classification */
    for (int i = 0; i < n; i++)
    {
        bool has_diag = false;
        for (int p = Cp[i]; p < Cp[i + 1]; ++p)
            if (Ci[p] == i)
            {
                has_diag = true;
                break;
            }
        const int d = degree[i];
        if (d == 1 && has_diag)
        {
            elen[i] = -2;
            nel++;
            Cp[i] = -1;
            w[i] = 0;
        }
        else if (d > dense || !has_diag)
        {
            nv[i] = 0;
            elen[i] = -1;
            nel++;
            Cp[i] = flip(n);
            nv[n]++;
        }
        else
        {
            if (head[d] != -1)
                last[head[d]] = i;
            next[i] = head[d];
            head[d] = i;
        }
    }
    elen[n] = -2;
    Cp[n] = -1;
    w[n] = 0;

    while (nel < n)
    {
        /* select node of minimum approximate degree */
        int k = -1;
        for (; mindeg < n && (k = head[mindeg]) == -1; mindeg++)
        {
        }
        if (next[k] != -1)
            last[next[k]] = -1;
        head[mindeg] = next[k];
        const int elenk = elen[k];
        int nvk = nv[k];
        nel += nvk;

        /* garbage collection */
        if (elenk > 0 && cnz + mindeg >= nzmax)
        {
            for (int j = 0; j < n; j++)
            {
                int p = Cp[j];
                if (p >= 0)
                {
                    Cp[j] = Ci[p];
                    Ci[p] = flip(j);
                }
            }
            int q = 0;
            for (int p = 0; p < cnz;)
            {
                const int j = flip(Ci[p++]);
                if (j >= 0)
                {
                    Ci[q] = Cp[j];
                    Cp[j] = q++;
                    for (int k3 = 0; k3 < len[j] - 1; k3++)
                        Ci[q++] = Ci[p++];
                }
            }
            cnz = q;
        }

        /* construct new element */
        int dk = 0;
        nv[k] = -nvk;
        int p = Cp[k];
        const int pk1 = (elenk == 0) ? p : cnz;
        int pk2 = pk1;
        for (int k1 = 1; k1 <= elenk + 1; k1++)
        {
            int e, pj, ln;
            if (k1 > elenk)
            {
                e = k;
                pj = p;
                ln = len[k] - elenk;
            }
            else
            {
                e = Ci[p++];
                pj = Cp[e];
                ln = len[e];
            }
            for (int k2 = 1; k2 <= ln; k2++)
            {
                const int i = Ci[pj++];
                const int nvi = nv[i];
                if (nvi <= 0)
                    continue;
                dk += nvi;
                nv[i] = -nvi;
                Ci[pk2++] = i;
                if (next[i] != -1)
                    last[next[i]] = last[i];
                if (last[i] != -1)
                    next[last[i]] = next[i];
                else
                    head[degree[i]] = next[i];
            }
            if (e != k)
            {
                Cp[e] = flip(k);
                w[e] = 0;
            }
        }
        if (elenk != 0)
            cnz = pk2;
        degree[k] = dk;
        Cp[k] = pk1;
        len[k] = pk2 - pk1;
        elen[k] = -2;

        /* find set differences */
        mark = wclear(mark, lemax, w, n);
        for (int pk = pk1; pk < pk2; pk++)
        {
            const int i = Ci[pk];
            const int eln = elen[i];
            if (eln <= 0)
                continue;
            const int nvi = -nv[i];
            const int wnvi = mark - nvi;
            for (int pp = Cp[i]; pp <= Cp[i] + eln - 1; pp++)
            {
                const int e = Ci[pp];
                if (w[e] >= mark)
                    w[e] -= nvi;
                else if (w[e] != 0)
                    w[e] = degree[e] + wnvi;
            }
        }

        /* degree update */
        for (int pk = pk1; pk < pk2; pk++)
        {
            const int i = Ci[pk];
            const int p1 = Cp[i];
            const int p2 = p1 + elen[i] - 1;
            int pn = p1;
            int h = 0, d = 0;
            for (int pp = p1; pp <= p2; pp++)
            {
                const int e = Ci[pp];
                if (w[e] != 0)
                {
                    const int dext = w[e] - mark;
                    if (dext > 0)
                    {
                        d += dext;
                        Ci[pn++] = e;
                        h += e;
                    }
                    else
                    {
                        Cp[e] = flip(k); /* aggressive absorption */
                        w[e] = 0;
                    }
                }
            }
            elen[i] = pn - p1 + 1;
            const int p3 = pn;
            const int p4 = p1 + len[i];
            for (int pp = p2 + 1; pp < p4; pp++)
            {
                const int j = Ci[pp];
                const int nvj = nv[j];
                if (nvj <= 0)
                    continue;
                d += nvj;
                Ci[pn++] = j;
                h += j;
            }
            if (d == 0)
            {
                Cp[i] = flip(k); /* mass elimination */
                const int nvi = -nv[i];
                dk -= nvi;
                nvk += nvi;
                nel += nvi;
                nv[i] = 0;
                elen[i] = -1;
            }
            else
            {
                degree[i] = std::min(degree[i], d);
                Ci[pn] = Ci[p3];
                Ci[p3] = Ci[p1];
                Ci[p1] = k;
                len[i] = pn - p1 + 1;
                h %= n;
                next[i] = hhead[h];
                hhead[h] = i;
                last[i] = h;
            }
        }
        degree[k] = dk;
        lemax = rmax(lemax, dk);
        mark = wclear(mark + lemax, lemax, w, n);

        /* supernode detection */
        for (int pk = pk1; pk < pk2; pk++)
        {
            int i = Ci[pk];
            if (nv[i] >= 0)
                continue;
            const int h = last[i];
            i = hhead[h];
            hhead[h] = -1;
            for (; i != -1 && next[i] != -1; i = next[i], mark++)
            {
                const int ln = len[i];
                const int eln = elen[i];
                for (int pp = Cp[i] + 1; pp <= Cp[i] + ln - 1; pp++)
                    w[Ci[pp]] = mark;
                int jlast = i;
                for (int j = next[i]; j != -1;)
                {
                    bool ok = (len[j] == ln) && (elen[j] == eln);
                    for (int pp = Cp[j] + 1; ok && pp <= Cp[j] + ln - 1; pp++)
                        if (w[Ci[pp]] != mark)
                            ok = false;
                    if (ok)
                    {
                        Cp[j] = flip(i);
                        nv[i] += nv[j];
                        nv[j] = 0;
                        elen[j] = -1;
                        j = next[j];
                        next[jlast] = j;
                    }
                    else
                    {
                        jlast = j;
                        j = next[j];
                    }
                }
            }
        }

        /* finalize new element */
        int pf = pk1;
        for (int pk = pk1; pk < pk2; pk++)
        {
            const int i = Ci[pk];
            const int nvi = -nv[i];
            if (nvi <= 0)
                continue;
            nv[i] = nvi;
            int d = degree[i] + dk - nvi;
            d = std::min(d, n - nel - nvi);
            if (head[d] != -1)
                last[head[d]] = i;
            next[i] = head[d];
            last[i] = -1;
            head[d] = i;
            mindeg = std::min(mindeg, d);
            degree[i] = d;
            Ci[pf++] = i;
        }
        nv[k] = nvk;
        if ((len[k] = pf - pk1) == 0)
        {
            Cp[k] = -1;
            w[k] = 0;
        }
        if (elenk != 0)
            cnz = pf;
    }

    /* postordering of the assembly tree */
    for (int i = 0; i < n; i++)
        Cp[i] = flip(Cp[i]);
    for (int j = 0; j <= n; j++)
        head[j] = -1;
    for (int j = n; j >= 0; j--)
    {
        if (nv[j] > 0)
            continue;
        next[j] = head[Cp[j]];
        head[Cp[j]] = j;
    }
    for (int e = n; e >= 0; e--)
    {
        if (nv[e] <= 0)
            continue;
        if (Cp[e] != -1)
        {
            next[e] = head[Cp[e]];
            head[Cp[e]] = e;
        }
    }
    for (int k = 0, i = 0; i <= n; i++)
        if (Cp[i] == -1)
            k = tdfs(i, k, head, next, perm.data(), w);
    perm.resize(n);
    return perm;
}

/* ------------------------------------------------------------------ LDLT */
/*
 * Eigen::SimplicialLDLT<SparseMatrix<double>, Upper> as used at src/eicos.cpp:897-901,
 * 1164-1166, 1477, 1599.  analyze(): symmetric pattern -> AMD -> Pinv; P = Pinv^-1;
 * permuted upper matrix via twistedBy (entries of a destination column arrive in source
 * traversal order, i.e. generally unsorted); Liu's elimination tree + column counts.
 * factorize(): twistedBy again (values) + up-looking LDL' (Davis' ldl_numeric);
 * fails iff a pivot is exactly zero.  solve(): P b, unit-lower, D^-1 (as a multiply by
 * the reciprocal), unit-upper, P^-1.
 */
struct Ldlt
{
    int n = 0;
    ivec pinv; /* AMD output: pinv[k] = old index of k-th pivot  (Eigen m_Pinv.indices()) */
    ivec P;    /* P[old] = new                                   (Eigen m_P.indices())    */
    ivec Up, Ui, Umap; /* permuted upper pattern; Umap[q] = slot in K's value array */
    ivec parent, Lp, Li, nzcol;
    vec Lx, D;
    bool ok = false;

    void analyze(const Csc &K)
    {
        n = K.cols;
        /* full symmetric pattern, columns ascending (permute_symm_to_fullsymm with no perm) */
        ivec cnt(n + 1, 0);
        for (int j = 0; j < n; j++)
            for (int k = K.p[j]; k < K.p[j + 1]; k++)
            {
                const int i = K.i[k];
                cnt[j + 1]++;
                if (i != j)
                    cnt[i + 1]++;
            }
        for (int j = 0; j < n; j++)
            cnt[j + 1] += cnt[j];
        ivec Sp(cnt), Si(cnt[n]), nxt(cnt.begin(), cnt.end() - 1);
        for (int j = 0; j < n; j++)
            for (int k = K.p[j]; k < K.p[j + 1]; k++)
            {
                const int i = K.i[k];
                Si[nxt[j]++] = i;
                if (i != j)
                    Si[nxt[i]++] = j;
            }
        pinv = amd_order(n, Sp, Si);
        if (const char *o = std::getenv("ORA_ORDERING")) /* experiments only: natural / reverse / rotN */
        {
            const std::string os(o);
            for (int k = 0; k < n; k++)
                pinv[k] = os == "natural" ? k : os == "reverse" ? n - 1 - k : (k + std::atoi(o + 3)) % n;
        }
        P.assign(n, 0);
        for (int k = 0; k < n; k++)
            P[pinv[k]] = k;

        /* twistedBy(P): upper -> upper */
        ivec c2(n + 1, 0);
        for (int j = 0; j < n; j++)
            for (int k = K.p[j]; k < K.p[j + 1]; k++)
                c2[rmax(P[K.i[k]], P[j]) + 1]++;
        for (int j = 0; j < n; j++)
            c2[j + 1] += c2[j];
        Up = c2;
        Ui.assign(K.nnz(), 0);
        Umap.assign(K.nnz(), 0);
        ivec fill(c2.begin(), c2.end() - 1);
        for (int j = 0; j < n; j++)
            for (int k = K.p[j]; k < K.p[j + 1]; k++)
            {
                const int ip = P[K.i[k]], jp = P[j];
                const int q = fill[rmax(ip, jp)]++;
                Ui[q] = std::min(ip, jp);
                Umap[q] = k;
            }

        /* elimination tree and column counts (ldl_symbolic) */
        parent.assign(n, -1);
        nzcol.assign(n, 0);
        ivec tags(n, 0);
        for (int k = 0; k < n; k++)
        {
            parent[k] = -1;
            tags[k] = k;
            nzcol[k] = 0;
            for (int q = Up[k]; q < Up[k + 1]; q++)
            {
                int i = Ui[q];
                if (i < k)
                    for (; tags[i] != k; i = parent[i])
                    {
                        if (parent[i] == -1)
                            parent[i] = k;
                        nzcol[i]++;
                        tags[i] = k;
                    }
            }
        }
        Lp.assign(n + 1, 0);
        for (int k = 0; k < n; k++)
            Lp[k + 1] = Lp[k] + nzcol[k];
        Li.assign(Lp[n], 0);
        Lx.assign(Lp[n], 0.0);
        D.assign(n, 0.0);
        ok = false;
    }

    void factorize(const Csc &K)
    {
        vec y(n, 0.0);
        ivec pattern(n, 0), tags(n, 0);
        ok = true;
        for (int k = 0; k < n; k++)
        {
            y[k] = 0.0;
            int top = n;
            tags[k] = k;
            nzcol[k] = 0;
            for (int q = Up[k]; q < Up[k + 1]; q++)
            {
                int i = Ui[q];
                if (i <= k)
                {
                    y[i] += K.x[Umap[q]];
                    int len;
                    for (len = 0; tags[i] != k; i = parent[i])
                    {
                        pattern[len++] = i;
                        tags[i] = k;
                    }
                    while (len > 0)
                        pattern[--top] = pattern[--len];
                }
            }
            double d = y[k];
            y[k] = 0.0;
            for (; top < n; ++top)
            {
                const int i = pattern[top];
                const double yi = y[i];
                y[i] = 0.0;
                const double l_ki = yi / D[i];
                const int p2 = Lp[i] + nzcol[i];
                int p;
                for (p = Lp[i]; p < p2; ++p)
                    y[Li[p]] -= Lx[p] * yi;
                d -= l_ki * yi;
                Li[p] = k;
                Lx[p] = l_ki;
                ++nzcol[i];
            }
            D[k] = d;
            if (d == 0.0)
            {
                ok = false;
                break;
            }
        }
    }

    void solve(const double *b, double *x) const
    {
        vec t(n);
        for (int i = 0; i < n; i++)
            t[P[i]] = b[i];
        if (!Li.empty())
            for (int j = 0; j < n; j++)
            {
                const double tj = t[j];
                if (tj != 0.0)
                    for (int p = Lp[j]; p < Lp[j + 1]; p++)
                        t[Li[p]] -= tj * Lx[p];
            }
        for (int j = 0; j < n; j++)
            t[j] = (1.0 / D[j]) * t[j];
        if (!Li.empty())
            for (int j = n - 1; j >= 0; j--)
            {
                double tj = t[j];
                for (int p = Lp[j]; p < Lp[j + 1]; p++)
                    tj -= Lx[p] * t[Li[p]];
                t[j] = tj;
            }
        for (int i = 0; i < n; i++)
            x[i] = t[P[i]];
    }
};

/* ------------------------------------------------------------------ solver state */
struct Info /* include/eicos.hpp:49-73; std::optional -> (has_*, value) */
{
    double pcost = 0, dcost = 0, pres = 0, dres = 0;
    bool pinf = false, dinf = false;
    bool has_pinfres = false, has_dinfres = false, has_relgap = false;
    double pinfres = 0, dinfres = 0, gap = 0, relgap = 0;
    double sigma = 0, mu = 0, step = 0, step_aff = 0, kapovert = 0;
    int iter = 0, iter_max = 0, nitref1 = 0, nitref2 = 0, nitref3 = 0;
};

/* optional<double> < double : an empty optional compares less than any value */
static inline bool opt_lt(bool has, double v, double rhs) { return has ? v < rhs : true; }

/* Information::isBetterThan src/eicos.cpp:23-68 */
static bool is_better(const Info &a, const Info &o)
{
    if (a.has_pinfres && a.kapovert > 1.)
    {
        if (o.has_pinfres)
            return (a.gap > 0. && o.gap > 0. && a.gap < o.gap) &&
                   (a.pinfres > 0. && a.pinfres < o.pres) &&
                   (a.mu > 0. && a.mu < o.mu);
        return (a.gap > 0. && o.gap > 0. && a.gap < o.gap) && (a.mu > 0. && a.mu < o.mu);
    }
    return (a.gap > 0. && o.gap > 0. && a.gap < o.gap) &&
           (a.pres > 0. && a.pres < o.pres) &&
           (a.dres > 0. && a.dres < o.dres) &&
           (a.kapovert > 0. && a.kapovert < o.kapovert) &&
           (a.mu > 0. && a.mu < o.mu);
}

struct SOCone /* include/eicos.hpp:81-95 */
{
    int dim = 0;
    double a = 0, d1 = 0, w = 0, eta = 0, eta_square = 0, u0 = 0, u1 = 0, v1 = 0;
    vec q, skbar, zkbar;
};

struct Work /* include/eicos.hpp:97-114 */
{
    vec x, y, z, s, lambda;
    double kap = 0, tau = 0, cx = 0, by = 0, hz = 0;
    Info i;
};

struct Solver
{
    int n = 0, p = 0, m = 0, l = 0, nc = 0, N = 0; /* n_var n_eq n_ineq n_lc n_sc dim_K */
    long long misaligned_cones = 0; /* test hook: line searches that skipped the offset advance (src/eicos.cpp:1423-1424) with cones still to come */
    Work w, wbest;
    vec lpv, lpw;
    std::vector<SOCone> cones;
    Csc G, A, Gt, At;
    vec c, h, b;
    vec rx, ry, rz;
    double hresx = 0, hresy = 0, hresz = 0, rt = 0, nx = 0, ny = 0, nz = 0, ns = 0;
    vec xeq, Aeq, Geq;
    bool equilibrated = false;
    double resx0 = 0, resy0 = 0, resz0 = 0;
    vec dsaff_by_W, W_times_dzaff, dsaff;
    vec rhs1, rhs2;
    Csc K;
    ivec Vslot, AGslot; /* replaces KKT_V_ptr / KKT_AG_ptr: offsets into K.x */
    Ldlt ldlt;
    bool verbose = std::getenv("ORA_VERBOSE") != nullptr;

    /* ---- construction: src/eicos.cpp:91-120 + build :132-187 + allocate :209-249 */
    void build(int n_, int m_, int p_, int ncones, const int *q,
               const ora_api_real *Gpr, const int *Gjc, const int *Gir,
               const ora_api_real *Apr, const int *Ajc, const int *Air,
               const ora_api_real *c_, const ora_api_real *h_, const ora_api_real *b_)
    {
        /* pointer ctor :103-117: a NULL triple leaves the matrix 0x0 and its vector empty */
        int ncz = 0;
        if (Gpr && Gjc && Gir)
        {
            G.rows = m_;
            G.cols = n_;
            G.p.assign(Gjc, Gjc + n_ + 1);
            G.i.assign(Gir, Gir + Gjc[n_]);
            G.x.assign(Gpr, Gpr + Gjc[n_]);
            h.assign(h_, h_ + m_);
            ncz = ncones;
        }
        if (Apr && Ajc && Air)
        {
            A.rows = p_;
            A.cols = n_;
            A.p.assign(Ajc, Ajc + n_ + 1);
            A.i.assign(Air, Air + Ajc[n_]);
            A.x.assign(Apr, Apr + Ajc[n_]);
            b.assign(b_, b_ + p_);
        }
        if (c_)
            c.assign(c_, c_ + n_);
        n = (int)c.size();
        p = A.rows;
        m = G.rows;
        int sumq = 0;
        cones.resize(ncz);
        for (int k = 0; k < ncz; k++)
        {
            cones[k].dim = q[k];
            sumq += q[k];
        }
        l = m - sumq; /* the ctor's `l` argument is ignored (:91,:155) */
        nc = ncz;
        N = n + p + m + 2 * nc;

        w.x.assign(n, 0);
        w.y.assign(p, 0);
        w.z.assign(m, 0);
        w.s.assign(m, 0);
        w.lambda.assign(m, 0);
        lpv.assign(l, 0);
        lpw.assign(l, 0);
        for (SOCone &sc : cones)
        {
            sc.q.assign(sc.dim - 1, 0);
            sc.skbar.assign(sc.dim, 0);
            sc.zkbar.assign(sc.dim, 0);
        }
        W_times_dzaff.assign(m, 0);
        dsaff_by_W.assign(m, 0);
        dsaff.assign(m, 0);
        rx.assign(n, 0);
        ry.assign(p, 0);
        rz.assign(m, 0);
        rhs1.assign(N, 0);
        rhs2.assign(N, 0);

        set_equilibration();
        Gt = transpose(G);
        At = transpose(A);
        setup_kkt();
    }

    /* ---- equilibration: src/eicos.cpp:256-374 */
    void set_equilibration()
    {
        xeq.assign(n, 1.0);
        Aeq.assign(p, 1.0);
        Geq.assign(m, 1.0);
        vec xt(n), At_(p), Gt_(m);
        for (int it = 0; it < EQUIL_ITERS; it++)
        {
            std::fill(xt.begin(), xt.end(), 0.0);
            std::fill(At_.begin(), At_.end(), 0.0);
            std::fill(Gt_.begin(), Gt_.end(), 0.0);
            for (const Csc *M : {&A, &G}) /* maxCols :267 */
                for (int j = 0; j < M->cols; j++)
                    for (int k = M->p[j]; k < M->p[j + 1]; k++)
                        xt[j] = rmax(std::fabs(M->x[k]), xt[j]);
            for (int k = 0; k < A.nnz(); k++) /* maxRows :256 */
                At_[A.i[k]] = rmax(std::fabs(A.x[k]), At_[A.i[k]]);
            for (int k = 0; k < G.nnz(); k++)
                Gt_[G.i[k]] = rmax(std::fabs(G.x[k]), Gt_[G.i[k]]);
            int ind = l; /* collapse each cone to the sum over its rows :338-344 */
            for (const SOCone &sc : cones)
            {
                double total = 0;
                for (int k = 0; k < sc.dim; k++)
                    total += Gt_[ind + k];
                for (int k = 0; k < sc.dim; k++)
                    Gt_[ind + k] = total;
                ind += sc.dim;
            }
            auto sq = [](double a) { return std::fabs(a) < 1e-6 ? 1. : std::sqrt(a); };
            for (double &v : xt)
                v = sq(v);
            for (double &v : At_)
                v = sq(v);
            for (double &v : Gt_)
                v = sq(v);
            for (int k = 0; k < A.nnz(); k++) /* rows first, then columns :353-356 */
                A.x[k] /= At_[A.i[k]];
            for (int k = 0; k < G.nnz(); k++)
                G.x[k] /= Gt_[G.i[k]];
            for (Csc *M : {&A, &G})
                for (int j = 0; j < M->cols; j++)
                    for (int k = M->p[j]; k < M->p[j + 1]; k++)
                        M->x[k] /= xt[j];
            for (int k = 0; k < n; k++)
                xeq[k] *= xt[k];
            for (int k = 0; k < p; k++)
                Aeq[k] *= At_[k];
            for (int k = 0; k < m; k++)
                Geq[k] *= Gt_[k];
        }
        for (int k = 0; k < n; k++)
            c[k] /= xeq[k];
        for (int k = 0; k < p; k++)
            b[k] /= Aeq[k];
        for (int k = 0; k < m; k++)
            h[k] /= Geq[k];
        equilibrated = true;
    }

    /* src/eicos.cpp:376-404 */
    void unset_equilibration()
    {
        for (int j = 0; j < A.cols; j++)
            for (int k = A.p[j]; k < A.p[j + 1]; k++)
                A.x[k] *= Aeq[A.i[k]] * xeq[j];
        for (int j = 0; j < G.cols; j++)
            for (int k = G.p[j]; k < G.p[j + 1]; k++)
                G.x[k] *= Geq[G.i[k]] * xeq[j];
        for (int k = 0; k < n; k++)
            c[k] *= xeq[k];
        for (int k = 0; k < p; k++)
            b[k] *= Aeq[k];
        for (int k = 0; k < m; k++)
            h[k] *= Geq[k];
        equilibrated = false;
    }

    /* ---- KKT assembly: src/eicos.cpp:1734-1890 (setFromTriplets => sorted CSC) and
     *      the slot tables of cacheIndices :1895-1988 */
    void setup_kkt()
    {
        struct T
        {
            int r, c;
            double v;
        };
        std::vector<T> tr;
        for (int k = 0; k < n; k++)
            tr.push_back({k, k, DELTASTAT});
        for (int k = n; k < n + p; k++)
            tr.push_back({k, k, -DELTASTAT});
        for (int col = 0; col < At.cols; col++) /* A' block; column offset A.cols() :1783 */
            for (int k = At.p[col]; k < At.p[col + 1]; k++)
                tr.push_back({At.i[k], A.cols + col, At.x[k]});
        int colK = n + p, colGt = 0;
        for (int k = 0; k < l; k++, colGt++, colK++)
            for (int q = Gt.p[colGt]; q < Gt.p[colGt + 1]; q++)
                tr.push_back({Gt.i[q], colK, Gt.x[q]});
        for (const SOCone &sc : cones)
        {
            for (int k = 0; k < sc.dim; k++, colGt++, colK++)
                for (int q = Gt.p[colGt]; q < Gt.p[colGt + 1]; q++)
                    tr.push_back({Gt.i[q], colK, Gt.x[q]});
            colK += 2;
        }
        int d = n + p;
        for (int k = 0; k < l; k++, d++)
            tr.push_back({d, d, -1.});
        for (const SOCone &sc : cones)
        {
            for (int k = 0; k < sc.dim; k++, d++)
                tr.push_back({d, d, -1.});
            tr.push_back({d, d, -1.});
            for (int k = 1; k < sc.dim; k++)
                tr.push_back({d - sc.dim + k, d, 0.});
            d++;
            tr.push_back({d, d, 1.});
            for (int k = 0; k < sc.dim; k++)
                tr.push_back({d - sc.dim - 1 + k, d, 0.});
            d++;
        }
        std::stable_sort(tr.begin(), tr.end(), [](const T &a, const T &b) { return a.c != b.c ? a.c < b.c : a.r < b.r; });
        K.rows = K.cols = N;
        K.p.assign(N + 1, 0);
        K.i.resize(tr.size());
        K.x.resize(tr.size());
        for (size_t k = 0; k < tr.size(); k++)
        {
            K.p[tr[k].c + 1]++;
            K.i[k] = tr[k].r;
            K.x[k] = tr[k].v;
        }
        for (int j = 0; j < N; j++)
            K.p[j + 1] += K.p[j];

        auto slot = [&](int r, int cidx) {
            for (int k = K.p[cidx]; k < K.p[cidx + 1]; k++)
                if (K.i[k] == r)
                    return k;
            return -1;
        };
        AGslot.clear();
        Vslot.clear();
        colK = n;
        for (int col = 0; col < At.cols; col++, colK++)
            for (int k = At.p[col]; k < At.p[col + 1]; k++)
                AGslot.push_back(slot(At.i[k], colK));
        colGt = 0;
        for (int k = 0; k < l; k++, colGt++, colK++)
            for (int q = Gt.p[colGt]; q < Gt.p[colGt + 1]; q++)
                AGslot.push_back(slot(Gt.i[q], colK));
        for (const SOCone &sc : cones)
        {
            for (int k = 0; k < sc.dim; k++, colGt++, colK++)
                for (int q = Gt.p[colGt]; q < Gt.p[colGt + 1]; q++)
                    AGslot.push_back(slot(Gt.i[q], colK));
            colK += 2;
        }
        d = n + p;
        for (int k = 0; k < l; k++, d++)
            Vslot.push_back(slot(d, d));
        for (const SOCone &sc : cones)
        {
            for (int k = 0; k < sc.dim; k++, d++)
                Vslot.push_back(slot(d, d));
            Vslot.push_back(slot(d, d));
            for (int k = 1; k < sc.dim; k++)
                Vslot.push_back(slot(d - sc.dim + k, d));
            d++;
            Vslot.push_back(slot(d, d));
            for (int k = 0; k < sc.dim; k++)
                Vslot.push_back(slot(d - sc.dim - 1 + k, d));
            d++;
        }
    }

    /* src/eicos.cpp:1990-2030 */
    void update_kkt_ag()
    {
        size_t s = 0;
        for (int k = 0; k < At.nnz(); k++)
            K.x[AGslot[s++]] = At.x[k];
        for (int k = 0; k < Gt.nnz(); k++)
            K.x[AGslot[s++]] = Gt.x[k];
    }

    /* src/eicos.cpp:807-846 */
    void reset_kkt_scalings()
    {
        size_t s = 0;
        for (int k = 0; k < l; k++)
            K.x[Vslot[s++]] = -1.;
        for (const SOCone &sc : cones)
        {
            for (int k = 0; k < sc.dim; k++)
                K.x[Vslot[s++]] = -1.;
            K.x[Vslot[s++]] = -1.;
            for (int k = 1; k < sc.dim; k++)
                K.x[Vslot[s++]] = 0.;
            K.x[Vslot[s++]] = 1.;
            for (int k = 0; k < sc.dim; k++)
                K.x[Vslot[s++]] = 0.;
        }
    }

    /* src/eicos.cpp:1691-1732 */
    void update_kkt_scalings()
    {
        size_t s = 0;
        for (int k = 0; k < l; k++)
            K.x[Vslot[s++]] = -lpv[k] - DELTASTAT;
        for (const SOCone &sc : cones)
        {
            K.x[Vslot[s++]] = -sc.eta_square * sc.d1 - DELTASTAT;
            for (int k = 1; k < sc.dim; k++)
                K.x[Vslot[s++]] = -sc.eta_square - DELTASTAT;
            K.x[Vslot[s++]] = -sc.eta_square;
            for (int k = 1; k < sc.dim; k++)
                K.x[Vslot[s++]] = -sc.eta_square * sc.v1 * sc.q[k - 1];
            K.x[Vslot[s++]] = sc.eta_square + DELTASTAT;
            K.x[Vslot[s++]] = -sc.eta_square * sc.u0;
            for (int k = 1; k < sc.dim; k++)
                K.x[Vslot[s++]] = -sc.eta_square * sc.u1 * sc.q[k - 1];
        }
    }

    /* ---- NT scalings: src/eicos.cpp:411-479 */
    bool update_scalings(const vec &s, const vec &z, vec &lambda)
    {
        for (int k = 0; k < l; k++)
        {
            lpv[k] = s[k] / z[k];
            lpw[k] = std::sqrt(lpv[k]);
        }
        int cs = l;
        for (SOCone &sc : cones)
        {
            const int d = sc.dim;
            const double sres = s[cs] * s[cs] - sqnorm(&s[cs] + 1, d - 1);
            const double zres = z[cs] * z[cs] - sqnorm(&z[cs] + 1, d - 1);
            if (sres <= 0 || zres <= 0)
                return false;
            const double snorm = std::sqrt(sres), znorm = std::sqrt(zres);
            for (int k = 0; k < d; k++)
            {
                sc.skbar[k] = s[cs + k] / snorm;
                sc.zkbar[k] = z[cs + k] / znorm;
            }
            sc.eta_square = snorm / znorm;
            sc.eta = std::sqrt(sc.eta_square);
            double gamma = 1. + dot(sc.skbar.data(), sc.zkbar.data(), d);
            gamma = std::sqrt(0.5 * gamma);
            const double a = (0.5 / gamma) * (sc.skbar[0] + sc.zkbar[0]);
            for (int k = 1; k < d; k++)
                sc.q[k - 1] = (0.5 / gamma) * (sc.skbar[k] - sc.zkbar[k]);
            const double ww = sqnorm(sc.q.data(), d - 1);
            const double cc = (1. + a) + ww / (1. + a);
            const double dd = 1. + 2. / (1. + a) + ww / ((1. + a) * (1. + a));
            const double d1 = rmax(0., 0.5 * (a * a + ww * (1. - (cc * cc) / (1. + ww * dd))));
            const double u0_square = a * a + ww - d1;
            const double c2byu02 = (cc * cc) / u0_square;
            if (c2byu02 - dd <= 0)
                return false;
            sc.d1 = d1;
            sc.u0 = std::sqrt(u0_square);
            sc.u1 = std::sqrt(c2byu02);
            sc.v1 = std::sqrt(c2byu02 - dd);
            sc.a = a;
            sc.w = ww;
            cs += d;
        }
        scale(z, lambda);
        return true;
    }

    /* lambda = W z : src/eicos.cpp:485-507 */
    void scale(const vec &z, vec &lambda)
    {
        for (int k = 0; k < l; k++)
            lambda[k] = lpw[k] * z[k];
        int cs = l;
        for (const SOCone &sc : cones)
        {
            const int d = sc.dim;
            const double zeta = dot(sc.q.data(), &z[cs] + 1, d - 1);
            const double factor = z[cs] + zeta / (1. + sc.a);
            const double z0 = z[cs]; /* z and lambda may alias only at distinct call sites; keep z0 */
            lambda[cs] = sc.eta * (sc.a * z0 + zeta);
            for (int k = 1; k < d; k++)
                lambda[cs + k] = sc.eta * (z[cs + k] + factor * sc.q[k - 1]);
            cs += d;
        }
    }

    /* y += W^2 x on the expanded layout : src/eicos.cpp:1629-1662 (last slot is ASSIGNED) */
    void scale2add(const double *x, double *y)
    {
        for (int k = 0; k < l; k++)
            y[k] += lpv[k] * x[k];
        int cs = l;
        for (const SOCone &sc : cones)
        {
            const int d = sc.dim;
            const int i1 = cs, i2 = i1 + 1, i3 = i2 + d - 1, i4 = i3 + 1;
            y[i1] += sc.eta_square * (sc.d1 * x[i1] + sc.u0 * x[i4]);
            const double v1x3_plus_u1x4 = sc.v1 * x[i3] + sc.u1 * x[i4];
            for (int k = 0; k < d - 1; k++)
                y[i2 + k] += sc.eta_square * (x[i2 + k] + v1x3_plus_u1x4 * sc.q[k]);
            const double qtx2 = dot(sc.q.data(), x + i2, d - 1);
            y[i3] += sc.eta_square * (sc.v1 * qtx2 + x[i3]);
            y[i4] = sc.eta_square * (sc.u0 * x[i1] + sc.u1 * qtx2 - x[i4]);
            cs += d + 2;
        }
    }

    /* src/eicos.cpp:761-805 */
    void bring_to_cone(const vec &r, vec &s)
    {
        double alpha = -GAMMA;
        for (int k = 0; k < l; k++)
            if (r[k] <= 0 && -r[k] > alpha)
                alpha = -r[k];
        int cs = l;
        for (const SOCone &sc : cones)
        {
            const double cres = r[cs] - norm2(&r[cs] + 1, sc.dim - 1);
            cs += sc.dim;
            if (cres <= 0 && -cres > alpha)
                alpha = -cres;
        }
        alpha += 1.;
        s = r;
        for (int k = 0; k < l; k++)
            s[k] += alpha;
        cs = l;
        for (const SOCone &sc : cones)
        {
            s[cs] += alpha;
            cs += sc.dim;
        }
    }

    /* src/eicos.cpp:643-689.  Eigen evaluates `d = a + S*v` as d = a; d += S*v, i.e. the sparse
     * product accumulates INTO the destination (scaleAndAddTo), column by column. */
    void compute_residuals()
    {
        std::fill(rx.begin(), rx.end(), 0.0);
        spmv_add(Gt, w.z.data(), rx.data(), -1.0);
        if (p > 0)
            spmv_add(At, w.y.data(), rx.data(), -1.0);
        hresx = norm2(rx.data(), n);
        for (int k = 0; k < n; k++)
            rx[k] -= w.tau * c[k];
        if (p > 0)
        {
            std::fill(ry.begin(), ry.end(), 0.0);
            spmv_add(A, w.x.data(), ry.data(), 1.0);
            hresy = norm2(ry.data(), p);
            for (int k = 0; k < p; k++)
                ry[k] -= w.tau * b[k];
        }
        else
            hresy = 0.;
        for (int k = 0; k < m; k++)
            rz[k] = w.s[k];
        spmv_add(G, w.x.data(), rz.data(), 1.0);
        hresz = norm2(rz.data(), m);
        for (int k = 0; k < m; k++)
            rz[k] -= w.tau * h[k];
        w.cx = dot(c.data(), w.x.data(), n);
        w.by = p > 0 ? dot(b.data(), w.y.data(), p) : 0.;
        w.hz = dot(h.data(), w.z.data(), m);
        rt = w.kap + w.cx + w.by + w.hz;
        nx = norm2(w.x.data(), n);
        ny = norm2(w.y.data(), p);
        nz = norm2(w.z.data(), m);
        ns = norm2(w.s.data(), m);
    }

    /* src/eicos.cpp:691-754 (pinfres/dinfres are sticky: only ever set) */
    void update_statistics()
    {
        Info &i = w.i;
        i.gap = dot(w.s.data(), w.z.data(), m);
        i.mu = (i.gap + w.kap * w.tau) / ((l + nc) + 1);
        i.kapovert = w.kap / w.tau;
        i.pcost = w.cx / w.tau;
        i.dcost = -(w.hz + w.by) / w.tau;
        if (i.pcost < 0.)
        {
            i.has_relgap = true;
            i.relgap = i.gap / (-i.pcost);
        }
        else if (i.dcost > 0.)
        {
            i.has_relgap = true;
            i.relgap = i.gap / i.dcost;
        }
        else
            i.has_relgap = false;
        const double nry = p > 0 ? norm2(ry.data(), p) / rmax(resy0 + nx, 1.) : 0.;
        const double nrz = norm2(rz.data(), m) / rmax(resz0 + nx + ns, 1.);
        i.pres = rmax(nry, nrz) / w.tau;
        i.dres = norm2(rx.data(), n) / rmax(resx0 + ny + nz, 1.) / w.tau;
        if ((w.hz + w.by) / rmax(ny + nz, 1.) < -RELTOL)
        {
            i.has_pinfres = true;
            i.pinfres = hresx / rmax(ny + nz, 1.);
        }
        if (w.cx / rmax(nx, 1.) < -RELTOL)
        {
            i.has_dinfres = true;
            i.dinfres = rmax(hresy / rmax(nx, 1.), hresz / rmax(nx + ns, 1.));
        }
    }

    /* src/eicos.cpp:526-641.  The reference evaluates pinfres.value() for a print even when it
     * is empty (would throw); the oracle follows the comparison semantics and returns PINF. */
    int check_exit(bool reduced)
    {
        const double feastol = reduced ? FEASTOL_INACC : FEASTOL;
        const double abstol = reduced ? ABSTOL_INACC : ABSTOL;
        const double reltol = reduced ? RELTOL_INACC : RELTOL;
        Info &i = w.i;
        if ((-w.cx > 0. || -w.by - w.hz >= -abstol) &&
            (i.pres < feastol && i.dres < feastol) &&
            (i.gap < abstol || opt_lt(i.has_relgap, i.relgap, reltol)))
        {
            i.pinf = false;
            i.dinf = false;
            return reduced ? OPTIMAL + INACC : OPTIMAL;
        }
        else if (i.has_dinfres && i.dinfres < feastol && w.tau < w.kap)
        {
            i.pinf = false;
            i.dinf = true;
            return reduced ? DINF + INACC : DINF;
        }
        else if ((i.has_pinfres && i.pinfres < feastol && w.tau < w.kap) ||
                 (w.tau < feastol && w.kap < feastol && opt_lt(i.has_pinfres, i.pinfres, feastol)))
        {
            i.pinf = true;
            i.dinf = false;
            return reduced ? PINF + INACC : PINF;
        }
        return NOT_CONVERGED;
    }

    /* src/eicos.cpp:1357-1378 */
    double conic_product(const vec &u, const vec &v, vec &o)
    {
        double mu = 0;
        for (int k = 0; k < l; k++)
        {
            o[k] = u[k] * v[k];
            mu += std::fabs(o[k]);
        }
        int cs = l;
        for (const SOCone &sc : cones)
        {
            const int d = sc.dim;
            const double u0 = u[cs], v0 = v[cs];
            const double o0 = dot(&u[cs], &v[cs], d);
            mu += std::fabs(o0);
            for (int k = 1; k < d; k++)
                o[cs + k] = u0 * v[cs + k] + v0 * u[cs + k];
            o[cs] = o0;
            cs += d;
        }
        return mu;
    }

    /* v = u \ w : src/eicos.cpp:1330-1351 (v may alias neither u nor w at its call site) */
    void conic_division(const vec &u, const vec &ww, vec &v)
    {
        for (int k = 0; k < l; k++)
            v[k] = ww[k] / u[k];
        int cs = l;
        for (const SOCone &sc : cones)
        {
            const int d = sc.dim;
            const double u0 = u[cs], w0 = ww[cs];
            const double rho = u0 * u0 - sqnorm(&u[cs] + 1, d - 1);
            const double zeta = dot(&u[cs] + 1, &ww[cs] + 1, d - 1);
            const double factor = (zeta / u0 - w0) / rho;
            v[cs] = (u0 * w0 - zeta) / rho;
            for (int k = 1; k < d; k++)
                v[cs + k] = factor * u[cs + k] + ww[cs + k] / u0;
            cs += d;
        }
    }

    /* src/eicos.cpp:1380-1469 (the `continue` deliberately skips the cone_start advance) */
    double line_search(const vec &lambda, const vec &ds, const vec &dz, double tau, double dtau, double kap, double dkap)
    {
        double alpha;
        if (l > 0)
        {
            double rhomin = ds[0] / lambda[0], sigmamin = dz[0] / lambda[0];
            for (int k = 1; k < l; k++)
            {
                rhomin = std::min(rhomin, ds[k] / lambda[k]);
                sigmamin = std::min(sigmamin, dz[k] / lambda[k]);
            }
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
        int cs = l;
        for (const SOCone &sc : cones)
        {
            const int d = sc.dim;
            const double lknorm2 = lambda[cs] * lambda[cs] - sqnorm(&lambda[cs] + 1, d - 1);
            if (lknorm2 <= 0.)
            {
                if (&sc != &cones.back())
                    misaligned_cones++;
                continue;
            }
            const double lknorm = std::sqrt(lknorm2);
            const double lknorminv = 1. / lknorm;
            const double lk0 = lambda[cs] / lknorm;
            double dsdot = 0, dzdot = 0;
            for (int k = 1; k < d; k++)
            {
                const double lkb = lambda[cs + k] / lknorm;
                dsdot += lkb * ds[cs + k];
                dzdot += lkb * dz[cs + k];
            }
            const double lkbar_times_dsk = lk0 * ds[cs] - dsdot;
            const double lkbar_times_dzk = lk0 * dz[cs] - dzdot;
            const double rho0 = lknorminv * lkbar_times_dsk;
            double factor = (lkbar_times_dsk + ds[cs]) / (lk0 + 1.);
            double acc = 0;
            for (int k = 1; k < d; k++)
            {
                const double r = lknorminv * (ds[cs + k] - factor * (lambda[cs + k] / lknorm));
                acc += r * r;
            }
            const double rhonorm = std::sqrt(acc) - rho0;
            const double sigma0 = lknorminv * lkbar_times_dzk;
            factor = (lkbar_times_dzk + dz[cs]) / (lk0 + 1.);
            acc = 0;
            for (int k = 1; k < d; k++)
            {
                const double r = lknorminv * (dz[cs + k] - factor * (lambda[cs + k] / lknorm));
                acc += r * r;
            }
            const double sigmanorm = std::sqrt(acc) - sigma0;
            const double conic_step = rmax(0., rmax(sigmanorm, rhonorm));
            if (conic_step != 0.)
                alpha = std::min(1. / conic_step, alpha);
            cs += d;
        }
        return std::min(rmax(alpha, STEPMIN), STEPMAX);
    }

    /* src/eicos.cpp:1471-1620 */
    int solve_kkt(const vec &rhs, vec &dx, vec &dy, vec &dz, bool initialize)
    {
        vec x(N);
        ldlt.solve(rhs.data(), x.data());
        const double error_threshold = (1. + norminf(rhs.data(), N)) * LINSYSACC;
        double nerr_prev = std::numeric_limits<double>::max();
        vec dx_ref(N, 0.0);
        const int mt = m + 2 * nc;
        const double *bx = rhs.data(), *by = rhs.data() + n, *bz = rhs.data() + n + p;
        vec ex(n), ey(p), ez(mt), Gdx(m), e(N);
        int k_ref;
        for (k_ref = 0; k_ref <= NITREF; k_ref++)
        {
            const double *xdx = x.data(), *xdy = x.data() + n;
            for (int k = 0; k < l; k++)
                dz[k] = x[n + p + k];
            int dzi = l, xi = n + p + l;
            for (const SOCone &sc : cones)
            {
                for (int k = 0; k < sc.dim; k++)
                    dz[dzi + k] = x[xi + k];
                dzi += sc.dim;
                xi += sc.dim + 2;
            }
            /* ex = bx - G'dz - A'dy - delta dx  (products accumulate into ex with alpha=-1) */
            for (int k = 0; k < n; k++)
                ex[k] = bx[k];
            spmv_add(Gt, dz.data(), ex.data(), -1.0);
            if (p > 0)
                spmv_add(At, xdy, ex.data(), -1.0);
            for (int k = 0; k < n; k++)
                ex[k] -= DELTASTAT * xdx[k];
            const double nex = norminf(ex.data(), n);
            /* ey = by - A dx + delta dy */
            for (int k = 0; k < p; k++)
                ey[k] = by[k];
            if (p > 0)
                spmv_add(A, xdx, ey.data(), -1.0);
            for (int k = 0; k < p; k++)
                ey[k] += DELTASTAT * xdy[k];
            const double ney = norminf(ey.data(), p);
            /* ez = bz - G dx + (static reg. terms) + V dz_true */
            std::fill(Gdx.begin(), Gdx.end(), 0.0);
            spmv_add(G, xdx, Gdx.data(), 1.0);
            for (int k = 0; k < l; k++)
                ez[k] = bz[k] - Gdx[k] + DELTASTAT * dz[k];
            int ezi = l;
            dzi = l;
            for (const SOCone &sc : cones)
            {
                const int d = sc.dim;
                for (int k = 0; k < d; k++)
                    ez[ezi + k] = bz[ezi + k] - Gdx[dzi + k];
                for (int k = 0; k < d - 1; k++)
                    ez[ezi + k] += DELTASTAT * dz[dzi + k];
                dzi += d;
                ezi += d;
                ez[ezi - 1] -= DELTASTAT * dz[dzi - 1];
                ez[ezi++] = 0.;
                ez[ezi++] = 0.;
            }
            const double *dz_true = x.data() + n + p;
            if (initialize)
                for (int k = 0; k < mt; k++)
                    ez[k] += dz_true[k];
            else
                scale2add(dz_true, ez.data());
            const double nez = norminf(ez.data(), mt);
            double nerr = rmax(nex, nez);
            if (p > 0)
                nerr = rmax(nerr, ney);
            if (k_ref > 0 && nerr > nerr_prev)
            {
                for (int k = 0; k < N; k++)
                    x[k] -= dx_ref[k];
                k_ref--;
                break;
            }
            if (k_ref == NITREF || nerr < error_threshold || (k_ref > 0 && nerr_prev < IRERRFACT * nerr))
                break;
            nerr_prev = nerr;
            std::copy(ex.begin(), ex.end(), e.begin());
            std::copy(ey.begin(), ey.end(), e.begin() + n);
            std::copy(ez.begin(), ez.end(), e.begin() + n + p);
            ldlt.solve(e.data(), dx_ref.data());
            for (int k = 0; k < N; k++)
                x[k] += dx_ref[k];
        }
        for (int k = 0; k < n; k++)
            dx[k] = x[k];
        for (int k = 0; k < p; k++)
            dy[k] = x[n + k];
        for (int k = 0; k < l; k++)
            dz[k] = x[n + p + k];
        int dzi = l, xi = n + p + l;
        for (const SOCone &sc : cones)
        {
            for (int k = 0; k < sc.dim; k++)
                dz[dzi + k] = x[xi + k];
            dzi += sc.dim;
            xi += sc.dim + 2;
        }
        return k_ref;
    }

    /* src/eicos.cpp:1670-1689 */
    void rhs_affine()
    {
        for (int k = 0; k < n; k++)
            rhs2[k] = rx[k];
        for (int k = 0; k < p; k++)
            rhs2[n + k] = -ry[k];
        for (int k = 0; k < l; k++)
            rhs2[n + p + k] = w.s[k] - rz[k];
        int ri = n + p + l, zi = l;
        for (const SOCone &sc : cones)
        {
            for (int k = 0; k < sc.dim; k++)
                rhs2[ri + k] = w.s[zi + k] - rz[zi + k];
            zi += sc.dim;
            ri += sc.dim;
            rhs2[ri++] = 0.;
            rhs2[ri++] = 0.;
        }
    }

    /* src/eicos.cpp:1282-1325 */
    void rhs_combined()
    {
        vec ds1(m), ds2(m);
        conic_product(w.lambda, w.lambda, ds1);
        conic_product(dsaff_by_W, W_times_dzaff, ds2);
        const double sigmamu = w.i.sigma * w.i.mu;
        for (int k = 0; k < l; k++)
        {
            ds1[k] += ds2[k];
            ds1[k] -= sigmamu;
        }
        int k0 = l;
        for (const SOCone &sc : cones)
        {
            ds1[k0] -= sigmamu;
            for (int k = 0; k < sc.dim; k++)
                ds1[k0 + k] += ds2[k0 + k];
            k0 += sc.dim;
        }
        conic_division(w.lambda, ds1, dsaff_by_W);
        scale(dsaff_by_W, ds1);
        const double oms = 1. - w.i.sigma;
        for (int k = 0; k < n + p; k++)
            rhs2[k] *= oms;
        for (int k = 0; k < l; k++)
            rhs2[n + p + k] = -oms * rz[k] + ds1[k];
        int ri = n + p + l;
        k0 = l;
        for (const SOCone &sc : cones)
        {
            for (int k = 0; k < sc.dim; k++)
                rhs2[ri + k] = -oms * rz[k0 + k] + ds1[k0 + k];
            k0 += sc.dim;
            ri += sc.dim;
            rhs2[ri++] = 0.;
            rhs2[ri++] = 0.;
        }
    }

    /* src/eicos.cpp:1271-1277 */
    void backscale()
    {
        for (int k = 0; k < n; k++)
            w.x[k] = w.x[k] / (xeq[k] * w.tau);
        for (int k = 0; k < p; k++)
            w.y[k] = w.y[k] / (Aeq[k] * w.tau);
        for (int k = 0; k < m; k++)
            w.z[k] = w.z[k] / (Geq[k] * w.tau);
        for (int k = 0; k < m; k++)
            w.s[k] = w.s[k] * (Geq[k] / w.tau);
    }

    void init_rhs_and_factor_prep()
    {
        reset_kkt_scalings();
        std::fill(rhs1.begin(), rhs1.end(), 0.0);
        for (int k = 0; k < p; k++)
            rhs1[n + k] = b[k];
        for (int k = 0; k < l; k++)
            rhs1[n + p + k] = h[k];
        int hi = l, ri = n + p + l;
        for (const SOCone &sc : cones)
        {
            for (int k = 0; k < sc.dim; k++)
                rhs1[ri + k] = h[hi + k];
            hi += sc.dim;
            ri += sc.dim + 2;
        }
        std::fill(rhs2.begin(), rhs2.end(), 0.0);
        for (int k = 0; k < n; k++)
            rhs2[k] = -c[k];
        resx0 = rmax(1., norm2(c.data(), n));
        resy0 = rmax(1., norm2(b.data(), p));
        resz0 = rmax(1., norm2(h.data(), m));
    }

    /* src/eicos.cpp:848-1262 */
    int solve()
    {
        int code = FATAL;
        init_rhs_and_factor_prep();
        ldlt.analyze(K);
        ldlt.factorize(K);
        if (!ldlt.ok)
            return FATAL;

        vec dx1(n), dy1(p), dz1(m), dx2(n), dy2(p), dz2(m);
        w.i.nitref1 = solve_kkt(rhs1, dx1, dy1, dz1, true);
        w.x = dx1;
        {
            vec neg(m);
            for (int k = 0; k < m; k++)
                neg[k] = -dz1[k];
            bring_to_cone(neg, w.s);
        }
        w.i.nitref2 = solve_kkt(rhs2, dx2, dy2, dz2, true);
        w.y = dy2;
        bring_to_cone(dz2, w.z);
        for (int k = 0; k < n; k++)
            rhs1[k] = -c[k];
        w.kap = 1.;
        w.tau = 1.;
        w.i.step = 0.;
        w.i.step_aff = 0.;
        w.i.pinf = false;
        w.i.dinf = false;
        w.i.iter_max = g_debug_iter_max > 0 ? g_debug_iter_max : ITER_MAX;
        double pres_prev = std::numeric_limits<double>::max();

        for (w.i.iter = 0; w.i.iter <= w.i.iter_max; w.i.iter++)
        {
            compute_residuals();
            update_statistics();
            if (verbose)
                std::printf("%2d  %+5.3e  %+5.3e  %+2.0e  %2.0e  %2.0e  %2.0e  %2.0e  step=%6.4f sigma=%2.0e IR %d/%d/%d tau=%g kap=%g pinfres=%.17g(%d) dinfres=%.17g(%d) pres=%.17g pres_prev=%.17g\n",
                            w.i.iter, (ora_api_real)w.i.pcost, (ora_api_real)w.i.dcost, (ora_api_real)w.i.gap, (ora_api_real)w.i.pres,
                            (ora_api_real)w.i.dres, (ora_api_real)w.i.kapovert, (ora_api_real)w.i.mu, (ora_api_real)w.i.step,
                            (ora_api_real)w.i.sigma, w.i.nitref1, w.i.nitref2, w.i.nitref3, (ora_api_real)w.tau, (ora_api_real)w.kap,
                            (ora_api_real)w.i.pinfres, (int)w.i.has_pinfres, (ora_api_real)w.i.dinfres, (int)w.i.has_dinfres,
                            (ora_api_real)w.i.pres, (ora_api_real)pres_prev);

            if (w.i.iter > 0 && (w.i.pres > SAFEGUARD * pres_prev || w.i.gap < 0.))
            {
                w = wbest;
                code = check_exit(true);
                if (code == NOT_CONVERGED)
                    code = NUMERICS;
                break;
            }
            pres_prev = w.i.pres;
            code = check_exit(false);
            if (code == NOT_CONVERGED)
            {
                if (w.i.iter > 0 && w.i.step == STEPMIN * GAMMA)
                {
                    w = wbest;
                    code = check_exit(true);
                    if (code == NOT_CONVERGED)
                        code = NUMERICS;
                    break;
                }
                else if (w.i.iter == w.i.iter_max)
                {
                    if (!is_better(w.i, wbest.i))
                        w = wbest;
                    code = check_exit(true);
                    if (code == NOT_CONVERGED)
                        code = MAXIT;
                    break;
                }
                else if (std::isnan(w.i.pcost))
                {
                    if (!(w.i.iter == 0 || is_better(w.i, wbest.i)))
                    {
                        w = wbest;
                        code = check_exit(true);
                        if (code == NOT_CONVERGED)
                            code = NUMERICS;
                    }
                    break; /* note: `code` stays NOT_CONVERGED(-87) on the first branch, as in the reference (:1117-1121) */
                }
            }
            else
                break;

            if (w.i.iter == 0 || is_better(w.i, wbest.i))
                wbest = w;

            update_scalings(w.s, w.z, w.lambda); /* return value ignored (:1160) */
            update_kkt_scalings();
            ldlt.factorize(K);
            if (!ldlt.ok)
                return FATAL; /* skips backscale (:1166-1170) */

            solve_kkt(rhs1, dx1, dy1, dz1, false);
            rhs_affine();
            solve_kkt(rhs2, dx2, dy2, dz2, false);

            const double dtau_denom = w.kap / w.tau - dot(c.data(), dx1.data(), n) - dot(b.data(), dy1.data(), p) - dot(h.data(), dz1.data(), m);
            const double dtauaff = (rt - w.kap + dot(c.data(), dx2.data(), n) + dot(b.data(), dy2.data(), p) + dot(h.data(), dz2.data(), m)) / dtau_denom;
            for (int k = 0; k < m; k++)
                dz2[k] += dtauaff * dz1[k];
            scale(dz2, W_times_dzaff);
            for (int k = 0; k < m; k++)
                dsaff_by_W[k] = -W_times_dzaff[k] - w.lambda[k];
            const double dkapaff = -w.kap - w.kap / w.tau * dtauaff;
            w.i.step_aff = line_search(w.lambda, dsaff_by_W, W_times_dzaff, w.tau, dtauaff, w.kap, dkapaff);
            const double t = 1. - w.i.step_aff;
            const double sigma = std::min(rmax(t * t * t, SIGMAMIN), SIGMAMAX);
            w.i.sigma = sigma;

            rhs_combined();
            w.i.nitref3 = solve_kkt(rhs2, dx2, dy2, dz2, false);
            const double bkap = w.kap * w.tau + dkapaff * dtauaff - sigma * w.i.mu;
            const double dtau = ((1. - sigma) * rt - bkap / w.tau + dot(c.data(), dx2.data(), n) + dot(b.data(), dy2.data(), p) + dot(h.data(), dz2.data(), m)) / dtau_denom;
            for (int k = 0; k < n; k++)
                dx2[k] += dtau * dx1[k];
            for (int k = 0; k < p; k++)
                dy2[k] += dtau * dy1[k];
            for (int k = 0; k < m; k++)
                dz2[k] += dtau * dz1[k];
            scale(dz2, W_times_dzaff);
            for (int k = 0; k < m; k++)
                dsaff_by_W[k] = -(dsaff_by_W[k] + W_times_dzaff[k]);
            const double dkap = -(bkap + w.kap * dtau) / w.tau;
            w.i.step = GAMMA * line_search(w.lambda, dsaff_by_W, W_times_dzaff, w.tau, dtau, w.kap, dkap);
            scale(dsaff_by_W, dsaff);
            for (int k = 0; k < n; k++)
                w.x[k] += w.i.step * dx2[k];
            for (int k = 0; k < p; k++)
                w.y[k] += w.i.step * dy2[k];
            for (int k = 0; k < m; k++)
                w.z[k] += w.i.step * dz2[k];
            for (int k = 0; k < m; k++)
                w.s[k] += w.i.step * dsaff[k];
            w.kap += w.i.step * dkap;
            w.tau += w.i.step * dtau;
        }
        backscale();
        return code;
    }

    /* src/eicos.cpp:2053-2082 : h only follows Gpr, b only follows Apr */
    void update_data_ptr(const ora_api_real *Gpr, const ora_api_real *Apr, const ora_api_real *c_, const ora_api_real *h_, const ora_api_real *b_)
    {
        if (equilibrated)
            unset_equilibration();
        if (Gpr)
        {
            std::copy(Gpr, Gpr + G.nnz(), G.x.begin());
            h.assign(h_, h_ + m);
        }
        if (Apr)
        {
            std::copy(Apr, Apr + A.nnz(), A.x.begin());
            b.assign(b_, b_ + p);
        }
        if (c_)
            c.assign(c_, c_ + n);
        set_equilibration();
        Gt = transpose(G);
        At = transpose(A);
        update_kkt_ag();
    }

    /* src/eicos.cpp:2032-2051 */
    void update_data_full(const ora_api_real *Gpr, const ora_api_real *Apr, const ora_api_real *c_, const ora_api_real *h_, const ora_api_real *b_)
    {
        std::copy(Gpr, Gpr + G.nnz(), G.x.begin());
        std::copy(Apr, Apr + A.nnz(), A.x.begin());
        c.assign(c_, c_ + n);
        h.assign(h_, h_ + m);
        b.assign(b_, b_ + p);
        set_equilibration();
        Gt = transpose(G);
        At = transpose(A);
        update_kkt_ag();
    }
};

} // namespace ora

/* ------------------------------------------------------------------ C ABI */
using ora::Solver;

#ifdef ORA_LONG_DOUBLE
#undef double
#endif

extern "C"
{

void *ora_setup(int n, int m, int p, int /*l*/, int ncones, const int *q,
                const double *Gpr, const int *Gjc, const int *Gir,
                const double *Apr, const int *Ajc, const int *Air,
                const double *c, const double *h, const double *b)
{
    Solver *s = new Solver();
    s->build(n, m, p, ncones, q, Gpr, Gjc, Gir, Apr, Ajc, Air, c, h, b);
    return s;
}

void ora_update_data(void *sv, const double *Gpr, const double *Apr, const double *c, const double *h, const double *b)
{
    ((Solver *)sv)->update_data_ptr(Gpr, Apr, c, h, b);
}

void ora_update_data_full(void *sv, const double *Gpr, const double *Apr, const double *c, const double *h, const double *b)
{
    ((Solver *)sv)->update_data_full(Gpr, Apr, c, h, b);
}

int ora_solve(void *sv) { return ((Solver *)sv)->solve(); }

/* test hook: lineSearch on caller data (lambda, ds, dz of length m) */
double ora_debug_line_search(void *sv, const double *lambda, const double *ds, const double *dz, double tau, double dtau, double kap,
                             double dkap)
{
    Solver *s = (Solver *)sv;
    return s->line_search(ora::vec(lambda, lambda + s->m), ora::vec(ds, ds + s->m), ora::vec(dz, dz + s->m), tau, dtau, kap, dkap);
}

/* test hook: how often lineSearch left `cone_start` behind (the reference's `continue` quirk) in this solver's life */
long long ora_misaligned_cones(void *sv) { return ((Solver *)sv)->misaligned_cones; }

void ora_get_solution(void *sv, double *x, double *y, double *z, double *s)
{
    Solver *S = (Solver *)sv;
    if (x)
        std::copy(S->w.x.begin(), S->w.x.end(), x);
    if (y)
        std::copy(S->w.y.begin(), S->w.y.end(), y);
    if (z)
        std::copy(S->w.z.begin(), S->w.z.end(), z);
    if (s)
        std::copy(S->w.s.begin(), S->w.s.end(), s);
}

void ora_get_info(void *sv, ora_info *o)
{
    const ora::Info &i = ((Solver *)sv)->w.i;
    o->pcost = i.pcost;
    o->dcost = i.dcost;
    o->pres = i.pres;
    o->dres = i.dres;
    o->pinfres = i.pinfres;
    o->dinfres = i.dinfres;
    o->gap = i.gap;
    o->relgap = i.relgap;
    o->sigma = i.sigma;
    o->mu = i.mu;
    o->step = i.step;
    o->step_aff = i.step_aff;
    o->kapovert = i.kapovert;
    o->pinf = i.pinf;
    o->dinf = i.dinf;
    o->has_pinfres = i.has_pinfres;
    o->has_dinfres = i.has_dinfres;
    o->has_relgap = i.has_relgap;
    o->iter = i.iter;
    o->iter_max = i.iter_max;
    o->nitref1 = i.nitref1;
    o->nitref2 = i.nitref2;
    o->nitref3 = i.nitref3;
}

void ora_cleanup(void *sv) { delete (Solver *)sv; }

void ora_dims(void *sv, int *dim_K, int *nnzK, int *nnzL)
{
    Solver *S = (Solver *)sv;
    if (dim_K)
        *dim_K = S->N;
    if (nnzK)
        *nnzK = S->K.nnz();
    if (nnzL)
        *nnzL = (int)S->ldlt.Li.size();
}

void ora_get_symbolic(void *sv, int *pinv, int *parent, int *Lp, int *Li, int *Kp, int *Ki)
{
    Solver *S = (Solver *)sv;
    const ora::Ldlt &f = S->ldlt;
    if (pinv)
        std::copy(f.pinv.begin(), f.pinv.end(), pinv);
    if (parent)
        std::copy(f.parent.begin(), f.parent.end(), parent);
    if (Lp)
        std::copy(f.Lp.begin(), f.Lp.end(), Lp);
    if (Li)
        std::copy(f.Li.begin(), f.Li.end(), Li);
    if (Kp)
        std::copy(S->K.p.begin(), S->K.p.end(), Kp);
    if (Ki)
        std::copy(S->K.i.begin(), S->K.i.end(), Ki);
}

int ora_debug_factor_init(void *sv)
{
    Solver *S = (Solver *)sv;
    S->init_rhs_and_factor_prep();
    S->ldlt.analyze(S->K);
    S->ldlt.factorize(S->K);
    return S->ldlt.ok ? 0 : -1;
}

void ora_debug_get_factor(void *sv, double *Lx, double *D)
{
    Solver *S = (Solver *)sv;
    if (Lx)
        std::copy(S->ldlt.Lx.begin(), S->ldlt.Lx.end(), Lx);
    if (D)
        std::copy(S->ldlt.D.begin(), S->ldlt.D.end(), D);
}

void ora_debug_get_K(void *sv, double *Kx)
{
    Solver *S = (Solver *)sv;
    std::copy(S->K.x.begin(), S->K.x.end(), Kx);
}

void ora_debug_ldl_solve(void *sv, const double *rhs, double *x)
{
    Solver *S = (Solver *)sv;
    ora::vec r(rhs, rhs + S->N), o(S->N);
    S->ldlt.solve(r.data(), o.data());
    std::copy(o.begin(), o.end(), x);
}

int ora_debug_solve_kkt(void *sv, const double *rhs, double *dx, double *dy, double *dz, int initialize)
{
    Solver *S = (Solver *)sv;
    ora::vec r(rhs, rhs + S->N), x(S->n), y(S->p), z(S->m);
    const int k = S->solve_kkt(r, x, y, z, initialize != 0);
    std::copy(x.begin(), x.end(), dx);
    std::copy(y.begin(), y.end(), dy);
    std::copy(z.begin(), z.end(), dz);
    return k;
}

void ora_debug_get_equil(void *sv, double *xe, double *Ae, double *Ge)
{
    Solver *S = (Solver *)sv;
    if (xe)
        std::copy(S->xeq.begin(), S->xeq.end(), xe);
    if (Ae)
        std::copy(S->Aeq.begin(), S->Aeq.end(), Ae);
    if (Ge)
        std::copy(S->Geq.begin(), S->Geq.end(), Ge);
}

void ora_debug_get_data(void *sv, double *Gpr, double *Apr, double *c, double *h, double *b)
{
    Solver *S = (Solver *)sv;
    if (Gpr)
        std::copy(S->G.x.begin(), S->G.x.end(), Gpr);
    if (Apr)
        std::copy(S->A.x.begin(), S->A.x.end(), Apr);
    if (c)
        std::copy(S->c.begin(), S->c.end(), c);
    if (h)
        std::copy(S->h.begin(), S->h.end(), h);
    if (b)
        std::copy(S->b.begin(), S->b.end(), b);
}

void ora_debug_set_iter_max(int iter_max) { ora::g_debug_iter_max = iter_max > 0 ? (iter_max < ora::ITER_MAX ? iter_max : ora::ITER_MAX) : 0; }

double ora_batch_run(int n, int m, int p, int l, int ncones, const int *q,
                     const double *Gpr, const int *Gjc, const int *Gir,
                     const double *Apr, const int *Ajc, const int *Air,
                     const double *c, const double *h, const double *b,
                     int batch,
                     const double *Gs, const double *As,
                     const double *cs, const double *hs, const double *bs,
                     int nthreads, int reset_sticky,
                     int *exitflags, int *iters, double *xs, double *ys, double *zs, double *ss,
                     double *pcosts)
{
    if (nthreads < 1)
        nthreads = 1;
    const size_t nnzG = (Gpr && Gjc) ? (size_t)Gjc[n] : 0, nnzA = (Apr && Ajc) ? (size_t)Ajc[n] : 0;
    std::vector<Solver *> solvers(nthreads);
    for (int t = 0; t < nthreads; t++)
    {
        solvers[t] = new Solver();
        solvers[t]->build(n, m, p, ncones, q, Gpr, Gjc, Gir, Apr, Ajc, Air, c, h, b);
    }
    (void)l;
    auto worker = [&](int t) {
        Solver *S = solvers[t];
        for (int i = t; i < batch; i += nthreads)
        {
            const double *Gi = Gs ? Gs + (size_t)i * nnzG : Gpr;
            const double *Ai = As ? As + (size_t)i * nnzA : Apr;
            const double *ci = cs ? cs + (size_t)i * n : c;
            const double *hi = hs ? hs + (size_t)i * m : h;
            const double *bi = bs ? bs + (size_t)i * p : b;
            S->update_data_ptr(Gi, Ai, ci, hi, bi);
            if (reset_sticky)
            { /* a fresh Solver has empty pinfres/dinfres; the reference never clears them (:720-728) */
                S->w.i.has_pinfres = S->w.i.has_dinfres = false;
                S->w.i.pinfres = S->w.i.dinfres = 0.0;
            }
            const int code = S->solve();
            if (exitflags)
                exitflags[i] = code;
            if (iters)
                iters[i] = S->w.i.iter;
            if (pcosts)
                pcosts[i] = S->w.i.pcost;
            if (xs)
                std::copy(S->w.x.begin(), S->w.x.end(), xs + (size_t)i * n);
            if (ys)
                std::copy(S->w.y.begin(), S->w.y.end(), ys + (size_t)i * p);
            if (zs)
                std::copy(S->w.z.begin(), S->w.z.end(), zs + (size_t)i * m);
            if (ss)
                std::copy(S->w.s.begin(), S->w.s.end(), ss + (size_t)i * m);
        }
    };
    const auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    for (int t = 1; t < nthreads; t++)
        th.emplace_back(worker, t);
    worker(0);
    for (auto &x : th)
        x.join();
    const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    for (Solver *s : solvers)
        delete s;
    return secs;
}

} /* extern "C" */
