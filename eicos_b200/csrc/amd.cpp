// Approximate-minimum-degree ordering, producing the permutation Eigen's AMDOrdering gives for
// a symmetric pattern that still contains its diagonal (what SimplicialLDLT::analyzePattern runs
// at src/eicos.cpp:897 of the reference).  The algorithm is Amestoy/Davis/Duff AMD as published in
// Davis, "Direct Methods for Sparse Linear Systems" (quotient graph, approximate external degrees,
// aggressive absorption, hashed supervariable detection, assembly-tree postorder) with the two
// deviations Eigen makes: diagonal entries stay in the adjacency lists (so "empty" means degree 1
// with a diagonal, and nodes lacking a diagonal count as dense), and the dense threshold is
// min(n-2, max(16, 10 sqrt(n))).  Tie-breaking in the degree lists is LIFO, as there.
//
// Structured as a small state machine (QuotientGraph) so that each stage can be read on its own;
// tests/test_symbolic.py checks it against the oracle's independent restatement on every fixture
// and on random patterns.
#include "symbolic.hpp"

#include <algorithm>
#include <cmath>

namespace eicos
{
namespace
{

inline int enc(int i) { return -i - 2; } // self-inverse marker for "absorbed into i"

struct QuotientGraph
{
    const int n;
    int dense_limit, free_at, capacity;
    int mark = 0, max_elem_size = 0, eliminated = 0, min_degree = 0;
    ivec ptr, adj;                                   // object j's list is adj[ptr[j] .. ptr[j]+len[j])
    ivec len, weight, link, bucket, nelem, deg, stamp, hbucket, back;
    // state of the element under construction
    int piv = -1, piv_elems = 0, piv_weight = 0, piv_deg = 0, e_begin = 0, e_end = 0;

    QuotientGraph(int n_, const ivec &Ap, const ivec &Ai)
        : n(n_), ptr(Ap.begin(), Ap.begin() + n_ + 1), len(n_ + 1), weight(n_ + 1, 1), link(n_ + 1, -1),
          bucket(n_ + 1, -1), nelem(n_ + 1, 0), deg(n_ + 1), stamp(n_ + 1, 1), hbucket(n_ + 1, -1), back(n_ + 1, -1)
    {
        dense_limit = std::min(n - 2, std::max(16, (int)(10 * std::sqrt((double)n))));
        free_at = Ap[n];
        capacity = free_at + free_at / 5 + 2 * n;
        adj.assign(capacity, 0);
        std::copy(Ai.begin(), Ai.begin() + free_at, adj.begin());
        for (int j = 0; j < n; j++)
            len[j] = ptr[j + 1] - ptr[j];
        len[n] = 0;
        for (int j = 0; j <= n; j++)
            deg[j] = len[j];
        refresh_stamps(0);
    }

    void refresh_stamps(int proposed)
    {
        if (proposed < 2 || proposed + max_elem_size < 0)
        {
            for (int j = 0; j < n; j++)
                if (stamp[j] != 0)
                    stamp[j] = 1;
            proposed = 2;
        }
        mark = proposed;
    }

    void push_degree(int i, int d)
    {
        if (bucket[d] != -1)
            back[bucket[d]] = i;
        link[i] = bucket[d];
        bucket[d] = i;
    }

    void classify_initial()
    {
        for (int i = 0; i < n; i++)
        {
            bool diag = false;
            for (int t = ptr[i]; t < ptr[i + 1] && !diag; ++t)
                diag = (adj[t] == i);
            const int d = deg[i];
            if (d == 1 && diag)
            { // isolated: a root of the assembly tree, eliminated at once
                nelem[i] = -2;
                eliminated++;
                ptr[i] = -1;
                stamp[i] = 0;
            }
            else if (d > dense_limit || !diag)
            { // dense: folded into the dummy element n, ordered last
                weight[i] = 0;
                nelem[i] = -1;
                eliminated++;
                ptr[i] = enc(n);
                weight[n]++;
            }
            else
                push_degree(i, d);
        }
        nelem[n] = -2;
        ptr[n] = -1;
        stamp[n] = 0;
    }

    void pop_min()
    {
        piv = -1;
        while (min_degree < n && (piv = bucket[min_degree]) == -1)
            min_degree++;
        if (link[piv] != -1)
            back[link[piv]] = -1;
        bucket[min_degree] = link[piv];
        piv_elems = nelem[piv];
        piv_weight = weight[piv];
        eliminated += piv_weight;
    }

    void compact_if_needed()
    {
        if (!(piv_elems > 0 && free_at + min_degree >= capacity))
            return;
        for (int j = 0; j < n; j++)
        {
            const int t = ptr[j];
            if (t >= 0)
            {
                ptr[j] = adj[t];
                adj[t] = enc(j);
            }
        }
        int dst = 0;
        for (int src = 0; src < free_at;)
        {
            const int j = enc(adj[src++]);
            if (j >= 0)
            {
                adj[dst] = ptr[j];
                ptr[j] = dst++;
                for (int c = 0; c < len[j] - 1; c++)
                    adj[dst++] = adj[src++];
            }
        }
        free_at = dst;
    }

    void unlink_degree(int i)
    {
        if (link[i] != -1)
            back[link[i]] = back[i];
        if (back[i] != -1)
            link[back[i]] = link[i];
        else
            bucket[deg[i]] = link[i];
    }

    void form_element()
    {
        piv_deg = 0;
        weight[piv] = -piv_weight;
        int cursor = ptr[piv];
        e_begin = (piv_elems == 0) ? cursor : free_at;
        e_end = e_begin;
        for (int pass = 1; pass <= piv_elems + 1; pass++)
        {
            int owner, from, count;
            if (pass > piv_elems)
            {
                owner = piv;
                from = cursor;
                count = len[piv] - piv_elems;
            }
            else
            {
                owner = adj[cursor++];
                from = ptr[owner];
                count = len[owner];
            }
            for (int c = 0; c < count; c++)
            {
                const int i = adj[from++];
                const int wi = weight[i];
                if (wi <= 0)
                    continue;
                piv_deg += wi;
                weight[i] = -wi;
                adj[e_end++] = i;
                unlink_degree(i);
            }
            if (owner != piv)
            {
                ptr[owner] = enc(piv);
                stamp[owner] = 0;
            }
        }
        if (piv_elems != 0)
            free_at = e_end;
        deg[piv] = piv_deg;
        ptr[piv] = e_begin;
        len[piv] = e_end - e_begin;
        nelem[piv] = -2;
    }

    void set_differences()
    {
        refresh_stamps(mark);
        for (int t = e_begin; t < e_end; t++)
        {
            const int i = adj[t];
            const int ne = nelem[i];
            if (ne <= 0)
                continue;
            const int wi = -weight[i];
            const int base = mark - wi;
            for (int u = ptr[i]; u <= ptr[i] + ne - 1; u++)
            {
                const int e = adj[u];
                if (stamp[e] >= mark)
                    stamp[e] -= wi;
                else if (stamp[e] != 0)
                    stamp[e] = deg[e] + base;
            }
        }
    }

    void update_degrees()
    {
        for (int t = e_begin; t < e_end; t++)
        {
            const int i = adj[t];
            const int first = ptr[i];
            const int last_elem = first + nelem[i] - 1;
            int out = first, hash = 0, d = 0;
            for (int u = first; u <= last_elem; u++)
            {
                const int e = adj[u];
                if (stamp[e] == 0)
                    continue;
                const int ext = stamp[e] - mark;
                if (ext > 0)
                {
                    d += ext;
                    adj[out++] = e;
                    hash += e;
                }
                else
                { // element e is a subset of the new one: absorb it
                    ptr[e] = enc(piv);
                    stamp[e] = 0;
                }
            }
            nelem[i] = out - first + 1;
            const int vars_from = out;
            const int list_end = first + len[i];
            for (int u = last_elem + 1; u < list_end; u++)
            {
                const int j = adj[u];
                const int wj = weight[j];
                if (wj <= 0)
                    continue;
                d += wj;
                adj[out++] = j;
                hash += j;
            }
            if (d == 0)
            { // indistinguishable from the pivot: eliminate together
                ptr[i] = enc(piv);
                const int wi = -weight[i];
                piv_deg -= wi;
                piv_weight += wi;
                eliminated += wi;
                weight[i] = 0;
                nelem[i] = -1;
            }
            else
            {
                deg[i] = std::min(deg[i], d);
                adj[out] = adj[vars_from];
                adj[vars_from] = adj[first];
                adj[first] = piv;
                len[i] = out - first + 1;
                hash %= n;
                link[i] = hbucket[hash];
                hbucket[hash] = i;
                back[i] = hash;
            }
        }
        deg[piv] = piv_deg;
        max_elem_size = std::max(max_elem_size, piv_deg);
        refresh_stamps(mark + max_elem_size);
    }

    void merge_supervariables()
    {
        for (int t = e_begin; t < e_end; t++)
        {
            int i = adj[t];
            if (weight[i] >= 0)
                continue;
            const int hash = back[i];
            i = hbucket[hash];
            hbucket[hash] = -1;
            for (; i != -1 && link[i] != -1; i = link[i], mark++)
            {
                const int li = len[i], ei = nelem[i];
                for (int u = ptr[i] + 1; u <= ptr[i] + li - 1; u++)
                    stamp[adj[u]] = mark;
                int prev = i;
                for (int j = link[i]; j != -1;)
                {
                    bool same = (len[j] == li) && (nelem[j] == ei);
                    for (int u = ptr[j] + 1; same && u <= ptr[j] + li - 1; u++)
                        same = (stamp[adj[u]] == mark);
                    if (same)
                    {
                        ptr[j] = enc(i);
                        weight[i] += weight[j];
                        weight[j] = 0;
                        nelem[j] = -1;
                        j = link[j];
                        link[prev] = j;
                    }
                    else
                    {
                        prev = j;
                        j = link[j];
                    }
                }
            }
        }
    }

    void close_element()
    {
        int out = e_begin;
        for (int t = e_begin; t < e_end; t++)
        {
            const int i = adj[t];
            const int wi = -weight[i];
            if (wi <= 0)
                continue;
            weight[i] = wi;
            int d = std::min(deg[i] + piv_deg - wi, n - eliminated - wi);
            push_degree(i, d);
            back[i] = -1;
            min_degree = std::min(min_degree, d);
            deg[i] = d;
            adj[out++] = i;
        }
        weight[piv] = piv_weight;
        len[piv] = out - e_begin;
        if (len[piv] == 0)
        {
            ptr[piv] = -1;
            stamp[piv] = 0;
        }
        if (piv_elems != 0)
            free_at = out;
    }

    ivec postorder()
    {
        ivec order(n + 1, 0);
        for (int i = 0; i < n; i++)
            ptr[i] = enc(ptr[i]);
        std::fill(bucket.begin(), bucket.end(), -1);
        for (int j = n; j >= 0; j--)
            if (weight[j] <= 0)
            { // absorbed variables hang under their representative
                link[j] = bucket[ptr[j]];
                bucket[ptr[j]] = j;
            }
        for (int e = n; e >= 0; e--)
            if (weight[e] > 0 && ptr[e] != -1)
            {
                link[e] = bucket[ptr[e]];
                bucket[ptr[e]] = e;
            }
        int placed = 0;
        ivec &stack = stamp;
        for (int root = 0; root <= n; root++)
        {
            if (ptr[root] != -1)
                continue;
            int top = 0;
            stack[0] = root;
            while (top >= 0)
            {
                const int node = stack[top];
                const int child = bucket[node];
                if (child == -1)
                {
                    top--;
                    order[placed++] = node;
                }
                else
                {
                    bucket[node] = link[child];
                    stack[++top] = child;
                }
            }
        }
        order.resize(n);
        return order;
    }
};

} // namespace

ivec amd_ordering(int n, const ivec &Ap, const ivec &Ai)
{
    if (n == 0)
        return ivec();
    QuotientGraph g(n, Ap, Ai);
    g.classify_initial();
    while (g.eliminated < n)
    {
        g.pop_min();
        g.compact_if_needed();
        g.form_element();
        g.set_differences();
        g.update_degrees();
        g.merge_supervariables();
        g.close_element();
    }
    return g.postorder();
}

} // namespace eicos
