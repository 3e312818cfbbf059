// Compiler of the FMA machine (machine.hpp): list scheduling into bundles, then ring / slot allocation.
// Pure host code.
#include "machine.hpp"

#include <algorithm>
#include <climits>
#include <string>
#include <cstdint>
#include <cstring>
#include <cstdlib>
#include <queue>
#include <set>
#include <stdexcept>

namespace eicos
{
namespace
{
inline int field(int row) { return row << M_FIELD_SHIFT; }

// ---- pass 1: bundles.  An operation may run once every value it reads was produced in an EARLIER bundle;
// among the ready ones the longest remaining dependency path goes first (ties: program order, which keeps
// the elimination-order locality of the values).
void schedule(const MProgram &P, int window, std::vector<ivec> &bundles, ivec &bundle_of)
{
    const int n = (int)P.ops.size(), nv = (int)P.vals.size();
    ivec producer(nv, -1);
    for (int t = 0; t < n; t++)
        if (P.ops[t].dst >= 0)
        {
            if (producer[P.ops[t].dst] >= 0)
                throw std::logic_error("machine: value defined twice");
            producer[P.ops[t].dst] = t;
        }
    std::vector<ivec> succ(n);
    ivec npred(n, 0);
    for (int t = 0; t < n; t++)
    {
        const MSrc *src[4] = {&P.ops[t].a, &P.ops[t].b, &P.ops[t].c, &P.ops[t].x3};
        int seen[4], ns = 0;
        for (const MSrc *s : src)
        {
            if (s->kind != MS_VAL)
                continue;
            const int q = producer[s->val];
            if (q < 0)
                continue; // external value
            if (q >= t)
                throw std::logic_error("machine: value used before it is defined");
            bool dup = false;
            for (int k = 0; k < ns; k++)
                dup = dup || seen[k] == q;
            if (dup)
                continue;
            seen[ns++] = q;
            succ[q].push_back(t);
            npred[t]++;
        }
    }
    ivec prio(n, 1);
    for (int t = n - 1; t >= 0; t--)
        for (int c : succ[t])
            prio[t] = std::max(prio[t], 1 + prio[c]);
    // values without a home row must find a slot: their operations go first, while slots are free
    for (int t = 0; t < n; t++)
        if (P.ops[t].dst >= 0 && P.vals[P.ops[t].dst].home_sel < 0 && npred[t] == 0)
            prio[t] = INT_MAX / 2;
    // Only operations close to the oldest unscheduled one (program order = elimination order, whose values
    // are short-lived) are candidates: running far ahead would fill the slots with values nobody reads soon.
    bundle_of.assign(n, -1);
    bundles.clear();
    ivec fresh;
    if (window >= n)
    { // no window: a heap by (priority, program order)
        typedef std::pair<int, int> key;
        std::priority_queue<key> heap;
        for (int t = 0; t < n; t++)
            if (npred[t] == 0)
                heap.push({prio[t], -t});
        int done = 0;
        while (done < n)
        {
            if (heap.empty())
                throw std::logic_error("machine: dependency cycle");
            ivec cur;
            while (!heap.empty() && (int)cur.size() < M_U)
            {
                cur.push_back(-heap.top().second);
                heap.pop();
            }
            std::sort(cur.begin(), cur.end());
            const int b = (int)bundles.size();
            fresh.clear();
            for (int t : cur)
            {
                bundle_of[t] = b;
                done++;
                for (int c : succ[t])
                    if (--npred[c] == 0)
                        fresh.push_back(c);
            }
            for (int c : fresh)
                heap.push({prio[c], -c});
            bundles.push_back(cur);
        }
        return;
    }
    std::set<int> ready; // by index
    for (int t = 0; t < n; t++)
        if (npred[t] == 0)
            ready.insert(t);
    int done = 0, oldest = 0;
    std::vector<std::pair<int, int>> cand; // (-priority, index)
    while (done < n)
    {
        if (ready.empty())
            throw std::logic_error("machine: dependency cycle");
        while (oldest < n && bundle_of[oldest] >= 0)
            oldest++;
        cand.clear();
        for (auto it = ready.begin(); it != ready.end() && (*it <= oldest + window || cand.empty()); ++it)
            cand.push_back({-prio[*it], *it});
        const size_t take = std::min<size_t>(M_U, cand.size());
        std::partial_sort(cand.begin(), cand.begin() + take, cand.end());
        ivec cur;
        for (size_t k = 0; k < take; k++)
            cur.push_back(cand[k].second);
        std::sort(cur.begin(), cur.end());
        const int b = (int)bundles.size();
        fresh.clear();
        for (int t : cur)
        {
            ready.erase(t);
            bundle_of[t] = b;
            done++;
            for (int c : succ[t])
                if (--npred[c] == 0)
                    fresh.push_back(c);
        }
        for (int c : fresh)
            ready.insert(c);
        bundles.push_back(cur);
    }
}

void verify(const MProgram &P, const std::vector<ivec> &bundles, const MachineCode &mc);

// ---- pass 2: ring rows, slots, control words.
// NR = P.nr right-hand sides per pass: a VECTOR operand (B, C, x3, the destination) is NR adjacent rows (one per
// job), an A operand one row.  Vector pops take NR consecutive ring rows starting at a multiple of NR.
void compile_with_window(const MProgram &P, int max_slots, int window, int RG, MachineCode &out)
{
    if (RG < 2 || RG > M_MAX_RING_GROUPS)
        throw std::logic_error("machine: ring depth out of range");
    const int NR = P.nr;
    if (NR != 1 && NR != 2)
        throw std::logic_error("machine: one or two right-hand sides per pass");
    const int RING_ROWS = RG * M_RING_GROUP, ring0 = m_row_slot0(NR) + NR * std::max(max_slots, 0);
    out = MachineCode();
    out.window = window;
    out.ring_groups = RG;
    out.slot_budget = std::max(max_slots, 0);
    out.nr = NR;
    const int n = (int)P.ops.size(), nv = (int)P.vals.size();
    std::vector<ivec> bundles;
    ivec bundle_of;
    schedule(P, window, bundles, bundle_of);

    // uses of every value: the operations that read it, in schedule order; their times are bundle keys
    // (16 x the scheduled bundle number, + 1 for every time an operation is pushed into an inserted bundle)
    std::vector<long long> optime(n, 0);
    std::vector<ivec> uses(nv);
    for (int b = 0; b < (int)bundles.size(); b++)
        for (int t : bundles[b])
        {
            optime[t] = 16LL * b;
            const MSrc *src[4] = {&P.ops[t].a, &P.ops[t].b, &P.ops[t].c, &P.ops[t].x3};
            for (const MSrc *s : src)
                if (s->kind == MS_VAL)
                    uses[s->val].push_back(t);
        }
    ivec uptr(nv, 0);
    const long long NEVER = LLONG_MAX;
    const auto next_use = [&](int v) { return uptr[v] < (int)uses[v].size() ? optime[uses[v][uptr[v]]] : NEVER; };

    // slots (one slot = NR rows)
    ivec slot_of(nv, -1), holder(std::max(max_slots, 0), -1), stamp(std::max(max_slots, 0), -1), free_slots;
    for (int s = (int)holder.size() - 1; s >= 0; s--)
        free_slots.push_back(s);
    int slot_top = 0;
    const auto slot_row = [&](int s) { return m_row_slot0(NR) + NR * s; };
    // where a value can be re-read from: bundle whose store phase wrote its home row (-1: before the program), -2: nowhere
    ivec written_at(nv, -2);
    {
        ivec producer(nv, -1);
        for (int t = 0; t < n; t++)
            if (P.ops[t].dst >= 0)
                producer[P.ops[t].dst] = t;
        for (int v = 0; v < nv; v++)
            if (producer[v] < 0)
            {
                if (P.vals[v].home_sel < 0)
                    throw std::logic_error("machine: external value without a home");
                written_at[v] = -1;
            }
    }

    // ring bookkeeping
    long long npop = 0;
    int released = 0, waited_upto = -1;
    std::vector<long long> first_pop;
    const auto pad_to = [&](long long target) {
        while (npop < target)
        {
            out.ld.push_back(M_LD_NONE);
            npop++;
            out.pads++;
        }
    };
    const auto push_single = [&](int word) {
        out.ld.push_back(word);
        return ring0 + (int)(npop++ % RING_ROWS);
    };
    const auto push_vector = [&](int word) { // NR rows, one per job, at a multiple of NR
        if (npop % NR)
            pad_to(npop + 1);
        const int row = ring0 + (int)(npop % RING_ROWS);
        for (int j = 0; j < NR; j++)
        {
            out.ld.push_back(word | (j ? M_LD_JOB_B : 0));
            npop++;
        }
        return row;
    };
    // control word of the previous bundle is closed once the next bundle knows how much padding it needs
    long long prev_ctrl_at = -1;
    int prev_wait = 0, prev_newg = 0;
    std::vector<long long> released_at; // control word of the bundle that released ring group g (its refill copies group g + RG)
    const auto close_prev = [&]() {
        if (prev_ctrl_at < 0)
            return;
        const int nrel = (int)(npop / M_RING_GROUP) - released;
        if (nrel < 0 || nrel > 31)
            throw std::logic_error("machine: bundle releases too many ring groups");
        released += nrel;
        released_at.resize((size_t)released, prev_ctrl_at);
        out.ops[(size_t)prev_ctrl_at] |= (prev_wait << MF_WAIT_SHIFT) | (nrel << MF_NREL_SHIFT) | (prev_newg << MF_NEWG_SHIFT);
    };
    // rows an operation pops, given what is in the slots now: (rows of one-use loads, rows of re-reads); the
    // one-use loads of a bundle need at most one alignment pad together, every re-read of a vector one of its own
    const auto pops_of = [&](const MOp &op, int &regular, int &reread) {
        const MSrc *src[4] = {&op.a, &op.b, &op.c, &op.x3};
        for (int k = 0; k < 4; k++)
        {
            const bool vec = k != 0;
            if (src[k]->kind == MS_LOAD)
                regular += vec ? NR : 1;
            else if (src[k]->kind == MS_VAL && slot_of[src[k]->val] < 0)
                reread += vec ? NR + (NR - 1) : 1;
        }
    };

    struct Pending
    {
        int val;
        size_t k_at; // word index of the K field
        size_t f_at; // word index of the flags
    };
    for (int b = 0; b < (int)bundles.size(); b++)
    {
        // ---- a bundle's pops must fit the rows in flight: 16 rows are kept for its re-reads (see below), the rest of
        // the window for its one-use loads; operations that do not fit move into a bundle of their own behind it
        {
            ivec &cur = bundles[b];
            int keep = 0, regular = NR - 1, reread = 0;
            for (; keep < (int)cur.size(); keep++)
            {
                int r1 = regular, r2 = reread;
                pops_of(P.ops[cur[keep]], r1, r2);
                if (keep > 0 && (r2 > 16 || r1 + r2 > RING_ROWS - 8))
                    break;
                regular = r1, reread = r2;
            }
            if (regular + reread > RING_ROWS - 8 || reread > 16)
                throw std::logic_error("machine: one operation pops more rows than the ring holds");
            if (keep < (int)cur.size())
            {
                ivec rest(cur.begin() + keep, cur.end());
                cur.resize(keep);
                for (int t : rest)
                    optime[t]++;
                bundles.insert(bundles.begin() + b + 1, rest);
            }
        }
        const ivec &cur = bundles[b];
        // ---- how far back were the values written that this bundle re-reads from their home rows?
        int pbmax = -1;
        for (int t : cur)
        {
            const MSrc *src[4] = {&P.ops[t].a, &P.ops[t].b, &P.ops[t].c, &P.ops[t].x3};
            for (const MSrc *s : src)
                if (s->kind == MS_VAL && slot_of[s->val] < 0)
                {
                    if (written_at[s->val] == -2)
                        throw std::logic_error("machine: value lost (neither in a slot nor at home)");
                    pbmax = std::max(pbmax, written_at[s->val]);
                }
        }
        if (pbmax >= 0) // two group boundaries between the writer's first pop and ours (see the safety rule below)
            pad_to((first_pop[pbmax] / M_RING_GROUP + 2) * M_RING_GROUP);
        close_prev();
        first_pop.push_back(npop);
        const long long window_end = (first_pop[b] / M_RING_GROUP + RG) * M_RING_GROUP; // rows issued so far

        // ---- records; pops of single-use rows first, re-reads of written values last (they need late ring positions)
        const size_t rec0 = out.ops.size();
        out.ops.resize(rec0 + M_BUNDLE_WORDS, 0);
        struct Reread
        {
            size_t at; // word to receive the field
            int val;
            bool vec, keep_b;
            size_t f_at, w5_at;
        };
        struct Load
        {
            size_t at;
            int word;
            bool vec;
        };
        std::vector<Load> loads; // one-use loads of the bundle, laid out once all are known
        std::vector<Reread> rereads;
        std::vector<Pending> dsts;
        ivec last_used; // values whose last use is in this bundle
        for (int u = 0; u < M_U; u++)
        {
            const size_t r = rec0 + (size_t)u * M_REC_WORDS;
            int *w = out.ops.data() + r;
            if (u >= (int)cur.size())
            { // padding operation: scratch = 0 - 0 * 0
                w[0] = w[1] = w[2] = field(M_ROW_ZERO);
                w[3] = field(m_row_trash(NR));
                out.nnop++;
                continue;
            }
            const MOp &op = P.ops[cur[u]];
            out.nops++;
            int flags = op.flags;
            bool has_const = false;
            double cval = 0.0;
            const auto resolve = [&](const MSrc &s, int which) { // which: 0 A, 1 B, 2 C, 3 x3
                const size_t at = r + (which == 3 ? 6 : which);
                const bool vec = which != 0;
                switch (s.kind)
                {
                case MS_ZERO:
                    out.ops[at] = field(M_ROW_ZERO);
                    break;
                case MS_NEGZERO:
                    out.ops[at] = field(m_row_negzero(NR));
                    break;
                case MS_LOAD:
                {
                    if (s.row < 0 || s.row > M_LD_ROW_MASK || s.sel < 0 || s.sel > 7)
                        throw std::logic_error("machine: load row out of range");
                    loads.push_back({at, (s.sel << M_LD_SEL_SHIFT) | s.row, vec && NR > 1});
                    break;
                }
                case MS_CONST:
                    if (which != 0 && which != 2)
                        throw std::logic_error("machine: constant in a B / x3 operand");
                    if (has_const)
                        throw std::logic_error("machine: two constants in one operation");
                    has_const = true;
                    cval = s.c;
                    flags |= which == 0 ? MF_ACONST : MF_CCONST;
                    out.ops[at] = field(M_ROW_ZERO);
                    break;
                case MS_VAL:
                {
                    const int v = s.val;
                    if (!vec && NR > 1)
                        throw std::logic_error("machine: a value as the A operand of a two-job program");
                    if (slot_of[v] >= 0)
                        out.ops[at] = field(slot_row(slot_of[v]));
                    else
                        rereads.push_back({at, v, vec, which == 1, r + 4, r + 5});
                    uptr[v]++;
                    if (uptr[v] == (int)uses[v].size())
                        last_used.push_back(v);
                    break;
                }
                }
            };
            if (op.x3.kind != MS_ZERO && (op.a.kind == MS_CONST || op.c.kind == MS_CONST))
                throw std::logic_error("machine: a functor operand and a constant in one operation");
            resolve(op.a, 0);
            resolve(op.b, 1);
            resolve(op.c, 2);
            if (op.x3.kind != MS_ZERO)
            {
                resolve(op.x3, 3);
                flags |= MF_X3;
            }
            if (has_const)
            {
                int32_t cw[2];
                std::memcpy(cw, &cval, sizeof(cval));
                w = out.ops.data() + r;
                w[6] = cw[0];
                w[7] = cw[1];
            }
            w = out.ops.data() + r;
            w[3] = field(m_row_trash(NR));
            w[4] = flags;
            w[5] = (flags & MF_OUT) ? op.out_row : 0;
            if (op.dst >= 0)
                dsts.push_back({op.dst, r + 3, r + 4});
        }
        // ---- one-use loads: the vector pops want a position that is a multiple of NR, so a one-row pop goes first
        // when the ring position is odd, then all vectors, then the remaining one-row pops
        {
            size_t next_single = 0;
            const auto place_single = [&]() {
                while (next_single < loads.size() && loads[next_single].vec)
                    next_single++;
                if (next_single == loads.size())
                    return false;
                out.ops[loads[next_single].at] = field(push_single(loads[next_single].word));
                next_single++;
                return true;
            };
            if (NR > 1 && npop % NR)
                place_single();
            for (const Load &l : loads)
                if (l.vec)
                    out.ops[l.at] = field(push_vector(l.word));
            while (place_single())
            {
            }
        }
        // ---- re-reads.  The home row of value v was written in the store phase of bundle pb; ring group g is
        // refilled at the end of the first bundle t with  pops(<= t) >= 8 (g - GROUPS + 1),  so the copy of pop
        // p sees the written row iff  p >= 8 (first_pop[pb] / 8 + GROUPS)   (and p < window_end: issued by now).
        std::stable_sort(rereads.begin(), rereads.end(),
                         [&](const Reread &x, const Reread &y) { return written_at[x.val] < written_at[y.val]; });
        for (const Reread &rr : rereads)
        {
            const int v = rr.val, pb = written_at[v];
            if (pb >= 0)
                pad_to((first_pop[pb] / M_RING_GROUP + RG) * M_RING_GROUP);
            const MVal &mv = P.vals[v];
            const int word = (mv.home_sel << M_LD_SEL_SHIFT) | mv.home_row;
            if (pb >= 0)
            { // the row was written by the program: the refill that copies it (issued by the bundle that released the
              // group RG places earlier - a bundle behind pb, by the rule above) needs the proxy fence
                if (rr.vec && npop % NR)
                    pad_to(npop + 1);
                for (long long pp = npop; pp < npop + (rr.vec ? NR : 1); pp++)
                {
                    const long long g = pp / M_RING_GROUP - RG;
                    if (g < 0 || g >= (long long)released_at.size())
                        throw std::logic_error("machine: re-read of a written row in a group nobody refilled");
                    out.ops[(size_t)released_at[(size_t)g]] |= MF_FENCE;
                }
            }
            out.ops[rr.at] = field(rr.vec ? push_vector(word) : push_single(word));
            out.far++;
        }
        if (npop > window_end)
            throw std::logic_error("machine: a bundle pops past the rows in flight");
        // ---- end of the load phase: slots of dead values are free again
        for (int v : last_used)
            if (slot_of[v] >= 0)
            {
                holder[slot_of[v]] = -1;
                free_slots.push_back(slot_of[v]);
                slot_of[v] = -1;
            }
        // ---- a slot for a value: free one, or the one whose holder is needed furthest in the future (if that
        // is later than our own next use and the holder can be re-read from its home row)
        const auto take_slot = [&](int v) -> int {
            if (next_use(v) == NEVER)
                return -1;
            int s = -1;
            if (!free_slots.empty())
            {
                s = free_slots.back();
                free_slots.pop_back();
            }
            else
            {
                int far = -1;
                for (int q = 0; q < (int)holder.size(); q++)
                    if (holder[q] >= 0 && stamp[q] != b && written_at[holder[q]] != -2 &&
                        (far < 0 || next_use(holder[q]) > next_use(holder[far])))
                        far = q;
                if (far < 0 || next_use(holder[far]) <= next_use(v))
                    return -1;
                slot_of[holder[far]] = -1;
                s = far;
            }
            holder[s] = v;
            stamp[s] = b; // not to be evicted again in this bundle: its store is already planned
            slot_of[v] = s;
            slot_top = std::max(slot_top, s + 1);
            return s;
        };
        // ---- destinations
        for (const Pending &d : dsts)
        {
            const int v = d.val;
            const bool has_out = (out.ops[d.f_at] & MF_OUT) != 0;
            if (has_out)
                written_at[v] = b;
            const int s = take_slot(v);
            if (s >= 0)
                out.ops[d.k_at] = field(slot_row(s));
            else if (next_use(v) != NEVER && !has_out)
            { // a partial sum without a slot waits in its home row
                const MVal &mv = P.vals[v];
                if (mv.home_sel < 0)
                    throw MachineOutOfSlots("machine: out of slots for a value without a home row");
                out.ops[d.f_at] |= MF_OUT;
                out.ops[d.f_at + 1] = mv.home_row;
                written_at[v] = b;
                out.spills++;
            }
        }
        // gathered values that are used again may be parked in a slot (MF_BKEEP; not together with MF_OUT: w5 is taken)
        if (P.keep_loads)
            for (const Reread &rr : rereads)
                if (rr.keep_b && slot_of[rr.val] < 0 && !(out.ops[rr.f_at] & (MF_OUT | MF_BKEEP)))
                {
                    const int s = take_slot(rr.val);
                    if (s >= 0)
                    {
                        out.ops[rr.f_at] |= MF_BKEEP;
                        out.ops[rr.w5_at] = field(slot_row(s));
                    }
                }
        // ---- control: wait for the newest group this bundle reads
        prev_wait = 0;
        prev_newg = 0;
        if (npop > first_pop[b])
        {
            const int g_need = (int)((npop - 1) / M_RING_GROUP);
            if (g_need > waited_upto)
            {
                prev_newg = g_need - waited_upto;
                if (prev_newg > 31)
                    throw std::logic_error("machine: a bundle opens too many ring groups");
                const int allowed = RG + released - 1 - g_need;
                if (allowed < 0 || allowed >= RG)
                    throw std::logic_error("machine: wait depth out of range");
                prev_wait = 1;
                while (prev_wait + 1 < M_WAIT_CODES && M_WAIT_N[prev_wait + 1] <= allowed)
                    prev_wait++;
                waited_upto = g_need;
            }
        }
        prev_ctrl_at = (long long)rec0 + 4;
    }
    const int nb = (int)bundles.size();
    if (nb == 0)
    { // an empty program still needs its END bundle
        const size_t rec0 = out.ops.size();
        out.ops.resize(rec0 + M_BUNDLE_WORDS, 0);
        for (int u = 0; u < M_U; u++)
        {
            int *w = out.ops.data() + rec0 + (size_t)u * M_REC_WORDS;
            w[0] = w[1] = w[2] = field(M_ROW_ZERO);
            w[3] = field(m_row_trash(NR));
        }
        prev_ctrl_at = (long long)rec0 + 4;
        prev_wait = 0;
        prev_newg = 0;
    }
    close_prev();
    out.ops[(size_t)prev_ctrl_at] |= MF_END;
    out.nbundles = std::max(nb, 1);
    out.nchunks = (out.nbundles + M_CHUNK_BUNDLES - 1) / M_CHUNK_BUNDLES;
    out.ops.resize((size_t)(out.nchunks + 1) * M_CHUNK_WORDS, 0); // + one chunk the record look-ahead may touch
    out.nld = (int)out.ld.size();
    out.ld.insert(out.ld.end(), (size_t)RING_ROWS, M_LD_NONE); // refills run RG groups ahead of the consumer
    while (out.ld.size() % M_LD_CHUNK_WORDS)
        out.ld.push_back(M_LD_NONE);
    out.nld_chunks = (int)(out.ld.size() / M_LD_CHUNK_WORDS);
    out.slot_rows = NR * slot_top;
    if (P.ops.size() <= 400000 || std::getenv("EICOS_VERIFY_PROGRAMS"))
        verify(P, bundles, out);
}

// Symbolic execution of compiled code against the program it came from: every operand field must hold the
// value the operation means to read at the time the device reads it (ring refills happen where the device
// does them, copies see the home rows as written so far).  Catches scheduling / ring / slot hazards on the
// host, independent of any numerics.  Runs on every compile of a modest program.
void verify(const MProgram &P, const std::vector<ivec> &bundles, const MachineCode &mc)
{
    struct Tag
    {
        int kind = 0; // 0 unknown, 1 ring copy (word, version), 2 value (of job `job`)
        int word = 0, ver = -3, val = -1, job = 0;
    };
    const int NR = mc.nr;
    const int RING_ROWS = mc.ring_groups * M_RING_GROUP, ring0 = m_row_slot0(NR) + NR * mc.slot_budget;
    const int nrows = ring0 + RING_ROWS;
    std::vector<Tag> rows((size_t)nrows);
    // home words (job A) -> value currently stored there (-1: what was there before the program); jobs move together
    std::vector<int> key;
    for (const MVal &v : P.vals)
        if (v.home_sel >= 0)
            key.push_back((v.home_sel << M_LD_SEL_SHIFT) | v.home_row);
    std::sort(key.begin(), key.end());
    key.erase(std::unique(key.begin(), key.end()), key.end());
    std::vector<int> cur(key.size(), -1);
    const auto slot_of_word = [&](int w) -> int {
        w &= ~M_LD_JOB_B;
        auto it = std::lower_bound(key.begin(), key.end(), w);
        return it != key.end() && *it == w ? (int)(it - key.begin()) : -1;
    };
    ivec producer(P.vals.size(), -1);
    for (size_t t = 0; t < P.ops.size(); t++)
        if (P.ops[t].dst >= 0)
            producer[P.ops[t].dst] = (int)t;
    long long issued = 0;
    const auto refill = [&]() {
        for (int k = 0; k < M_RING_GROUP; k++)
        {
            const long long idx = issued * M_RING_GROUP + k;
            if (idx >= (long long)mc.ld.size())
                throw std::logic_error("machine verify: refill past the load list");
            const int w = mc.ld[(size_t)idx];
            if (w == M_LD_NONE)
                continue;
            Tag &t = rows[(size_t)(ring0 + idx % RING_ROWS)];
            t.kind = 1;
            t.word = w;
            const int s = slot_of_word(w);
            t.ver = s >= 0 ? cur[s] : -1;
        }
        issued++;
    };
    for (int g = 0; g < mc.ring_groups; g++)
        refill();
    const auto fail = [&](int b, int u, const std::string &what) {
        throw std::logic_error("machine verify: bundle " + std::to_string(b) + " op " + std::to_string(u) + ": " + what);
    };
    for (int b = 0; b < (int)bundles.size(); b++)
    {
        const int *rec = mc.ops.data() + (size_t)b * M_BUNDLE_WORDS;
        // load phase
        for (int u = 0; u < (int)bundles[b].size(); u++)
        {
            const MOp &op = P.ops[bundles[b][u]];
            const int *w = rec + u * M_REC_WORDS;
            const auto check = [&](const MSrc &s, int f, bool vec, const char *name) {
                const int row0 = f >> M_FIELD_SHIFT;
                for (int j = 0; j < (vec ? NR : 1); j++)
                {
                    const int row = row0 + j;
                    if (row < 0 || row >= nrows)
                        fail(b, u, "field out of range");
                    const Tag &t = rows[(size_t)row];
                    switch (s.kind)
                    {
                    case MS_ZERO:
                        if (row0 != M_ROW_ZERO)
                            fail(b, u, name);
                        break;
                    case MS_NEGZERO:
                        if (row0 != m_row_negzero(NR))
                            fail(b, u, name);
                        break;
                    case MS_CONST:
                        break;
                    case MS_LOAD:
                        if (t.kind != 1 || t.word != (((s.sel << M_LD_SEL_SHIFT) | s.row) | (j ? M_LD_JOB_B : 0)))
                            fail(b, u, name);
                        break;
                    case MS_VAL:
                    {
                        if (t.kind == 2 && t.val == s.val && t.job == j)
                            break;
                        const MVal &mv = P.vals[s.val];
                        const int want = producer[s.val] < 0 ? -1 : s.val;
                        const int hw = mv.home_sel >= 0 ? (((mv.home_sel << M_LD_SEL_SHIFT) | mv.home_row) | (j ? M_LD_JOB_B : 0)) : -1;
                        if (t.kind == 1 && mv.home_sel >= 0 && t.word == hw && t.ver == want)
                            break;
                        fail(b, u, std::string(name) + ": wants value " + std::to_string(s.val) + " job " + std::to_string(j) + " (home word " +
                                       std::to_string(hw) + ", producer " + std::to_string(producer[s.val]) + "), row " + std::to_string(row) +
                                       " holds kind " + std::to_string(t.kind) + " word " + std::to_string(t.word) + " ver " +
                                       std::to_string(t.ver) + " val " + std::to_string(t.val));
                    }
                    }
                }
            };
            check(op.a, w[0], false, "operand A");
            check(op.b, w[1], true, "operand B");
            check(op.c, w[2], true, "operand C");
            if (op.x3.kind != MS_ZERO)
                check(op.x3, w[6], true, "operand x3");
        }
        // store phase
        for (int u = 0; u < (int)bundles[b].size(); u++)
        {
            const MOp &op = P.ops[bundles[b][u]];
            const int *w = rec + u * M_REC_WORDS;
            const int krow = w[3] >> M_FIELD_SHIFT;
            if (krow != m_row_trash(NR))
            {
                if (krow < m_row_slot0(NR) || krow + NR > ring0)
                    fail(b, u, "destination is not a slot");
                for (int j = 0; j < NR; j++)
                {
                    rows[(size_t)krow + j].kind = 2;
                    rows[(size_t)krow + j].val = op.dst;
                    rows[(size_t)krow + j].job = j;
                }
            }
            if (w[4] & MF_OUT)
            {
                if (op.dst < 0 || P.vals[op.dst].home_sel < 0 || P.vals[op.dst].home_row != w[5])
                    fail(b, u, "out row is not the value's home");
                const int s = slot_of_word((P.vals[op.dst].home_sel << M_LD_SEL_SHIFT) | w[5]);
                cur[(size_t)s] = op.dst;
            }
            if (w[4] & MF_BKEEP)
            {
                const int row = w[5] >> M_FIELD_SHIFT;
                if (op.b.kind != MS_VAL || row < m_row_slot0(NR) || row + NR > ring0)
                    fail(b, u, "bad keep");
                for (int j = 0; j < NR; j++)
                {
                    rows[(size_t)row + j].kind = 2;
                    rows[(size_t)row + j].val = op.b.val;
                    rows[(size_t)row + j].job = j;
                }
            }
        }
        for (int k = (rec[4] >> MF_NREL_SHIFT) & 31; k > 0; k--)
            refill();
    }
}

} // namespace

// The look-ahead window of the scheduler trades bundles (latency of one tile) against re-reads (HBM traffic of
// the batch); which one wins depends on the shape of the dependency graph, so a few windows are compiled and
// the cheapest program is kept: cost = 8 x bundles + rows copied from global memory (a bundle costs the warp about
// a hundred instructions whether it is full or not, a copied row three or four and 512 bytes of traffic).
void machine_compile(const MProgram &P, int max_slots, MachineCode &out, int tune_slots, int ring_groups, int window)
{
    if (tune_slots < max_slots)
        tune_slots = max_slots;
    if (const char *v = std::getenv("EICOS_SCHED_WINDOW"))
        window = std::max(M_U, std::atoi(v));
    if (window <= 0)
    {
        const int windows[] = {16, 32, 64, 128, 512, INT_MAX / 2};
        double best = 0;
        for (int w : windows)
        {
            if ((long long)P.ops.size() > 300000)
            { // huge programs (dense fronts): no tuning
                window = 128;
                break;
            }
            MachineCode c;
            try
            {
                compile_with_window(P, tune_slots, w, 3, c);
            }
            catch (const MachineOutOfSlots &)
            { // this order of operations keeps more values without a home row alive than there are slots
                continue;
            }
            const double cost = 8.0 * (double)c.nbundles + (double)c.nld;
            if (window <= 0 || cost < best)
            {
                best = cost;
                window = w;
            }
        }
        if (window <= 0)
            throw MachineOutOfSlots("machine: no order of the operations fits the slots");
    }
    compile_with_window(P, max_slots, window, ring_groups, out);
}

} // namespace eicos
