// The FMA machine: how the numeric kernels of a tile (triangular sweeps, KKT mat-vecs, numeric LDL')
// are executed by ONE warp without waiting for its own latencies.
//
// Every numeric kernel on the hot path is, per instance, a long list of operations
//     k = c - a * b          (src/eicos.cpp:1477,1599: the sweeps of ldlt.solve; :1511-1576: the KKT
//                             mat-vec of the refinement residual; :643-689: computeResiduals; :900,1164:
//                             the Schur updates of ldlt.factorize)
// on rows (one row = the TILE values of one per-instance quantity).  The sparsity pattern is shared by
// the batch and fixed for the life of a handle, so the whole list, its dependencies and its memory
// traffic are known to the host.  The host therefore compiles each kernel into a statically scheduled
// program for a small 3-address machine:
//   * operands live in shared-memory rows: a RING that global rows are gathered into by cp.async a few
//     groups ahead of use (the program's load list, in consumption order; the ring depth is what bounds
//     the bytes one tile has in flight, so every program is compiled for a shallow ring - full machine -
//     and a deep one - few tiles per SM), SLOTS that hold computed
//     values for their live range (Belady allocation, evicted values are re-read from their home row),
//     and three constant rows (0, -0, scratch);
//   * operations are 32-byte records {A, B, C, K, flags, out, constant}; M_U of them form a BUNDLE of
//     mutually independent operations: the warp issues all operand loads of a bundle, then all
//     multiply-adds, then all stores, so M_U dependent chains overlap inside one in-order warp;
//   * bundles are formed by list scheduling of the dependency graph (longest path first), one bundle
//     of latency between producer and consumer; the summation order inside every dot product stays
//     the sequential one (Eigen's), only independent rows are interleaved;
//   * the record stream itself arrives in shared memory by TMA bulk copies (cp.async.bulk + mbarrier).
// The device side is Machine in tile_program.hpp; the emulator build runs the same records.
#pragma once

#include "symbolic.hpp"

#include <stdexcept>
#include <vector>

namespace eicos
{

#ifndef EICOS_M_U
#define EICOS_M_U 4
#endif
constexpr int M_U = EICOS_M_U; // operations per bundle
constexpr int M_REC_WORDS = 8;  // 32-byte records
constexpr int M_BUNDLE_WORDS = M_U * M_REC_WORDS;
constexpr int M_CHUNK_BUNDLES = 8; // bundles per TMA chunk of the record stream (1 KB)
constexpr int M_CHUNK_WORDS = M_CHUNK_BUNDLES * M_BUNDLE_WORDS;
constexpr int M_CHUNKS = 2;        // chunks in the shared-memory ops ring
constexpr int M_RING_GROUP = 8;    // rows per ring group (= one cp.async commit group)
constexpr int M_MAX_RING_GROUPS = 16;
// shared-memory rows: 0, -0, scratch, the slots (the program's slot budget), then the ring
// (the three constant operands are vectors like any B / C operand: nr rows each)
constexpr int M_ROW_ZERO = 0;
constexpr int m_row_negzero(int nr) { return nr; }
constexpr int m_row_trash(int nr) { return 2 * nr; }
constexpr int m_row_slot0(int nr) { return 3 * nr; }
// the load list arrives in shared memory like the records: chunks of M_LD_CHUNK_WORDS words by TMA
constexpr int M_LD_CHUNK_WORDS = 64, M_LD_CHUNK_GROUPS = M_LD_CHUNK_WORDS / M_RING_GROUP, M_LD_CHUNKS = 2;
// WAIT codes of the bundle control: cp.async.wait_group takes an immediate, deep rings get the nearest one below
constexpr int M_WAIT_CODES = 8;
constexpr int M_WAIT_N[M_WAIT_CODES] = {-1, 0, 1, 2, 3, 5, 8, 12};
constexpr int M_FIELD_SHIFT = 9;   // a field is row << 9: the byte offset of a 512-byte row on the device
constexpr int M_LD_NONE = -1;      // load-list word: no copy (padding)
constexpr int M_LD_SEL_SHIFT = 27; // load-list word before materialisation: job B << 30 | selector << 27 | row
constexpr int M_LD_ROW_MASK = (1 << M_LD_SEL_SHIFT) - 1;
constexpr int M_LD_JOB_B = 1 << 30; // the row of the second job's vector (two-job programs)

// record = [A, B, C, K, flags, w5, w6, w7]
//   result = C - A * B   (MF_POS: C + A * B;  MF_RECIP: 1 / C)   -> row K, and
//   MF_OUT:   -> global row  out base + w5   (MF_OUT2: the run's second out base)
//   MF_ACONST / MF_CCONST: A / C is the double in (w6, w7) instead of a row
//   MF_BKEEP: the B operand is also copied to row w5 (a gathered vector entry that is used again)
//   MF_FIN:   the kernel's finish functor sees (kind, w5, result, B, the row named by field w6)
//   MF_AONE:  A = 1.0 when the run's `a_one` switch is set (the LP scaling term of the refinement residual
//             while the scalings are the identity, src/eicos.cpp:1557-1559)
// bundle control, in the flags of a bundle's first record:
//   WAIT  code > 0: cp.async.wait_group(M_WAIT_N[code]) before the operand loads (cp.async form of the data ring)
//   NEWG  ring groups this bundle is the first to read: their mbarriers are waited for (TMA form of the data ring)
//   FENCE proxy fence in front of this bundle's refills (TMA form)
//   NREL  ring groups consumed by the end of this bundle: each is refilled with the next group of the load list
//   END   last bundle
enum : int
{
    MF_OUT = 1 << 0,
    MF_ACONST = 1 << 1,
    MF_CCONST = 1 << 2,
    MF_BKEEP = 1 << 3,
    MF_RECIP = 1 << 4,
    MF_FIN = 1 << 5,
    MF_OUT2 = 1 << 6,
    MF_AONE = 1 << 7,
    MF_KIND_SHIFT = 8, // 4 bits
    MF_X3 = 1 << 12,   // field w6 names a fourth operand row (set by the compiler)
    // bundle control for the TMA data ring (tile_program.hpp: the rows of a ring group arrive by eight bulk copies
    // counted on the group's mbarrier):
    MF_FENCE = 1 << 13,     // the refills at the end of this bundle copy a home row the program itself has written:
                            // a generic -> async proxy fence goes in front of them
    MF_NEWG_SHIFT = 19,     // 5 bits: ring groups this bundle is the first to read (it waits for their mbarriers)
    MF_WAIT_SHIFT = 16, // 3 bits
    MF_NREL_SHIFT = 25, // 5 bits
    MF_END = 1 << 24,
    MF_POS = (int)0x80000000u
};

// ---- what a kernel hands to the compiler
enum MSrcKind : int
{
    MS_ZERO,    // +0.0
    MS_NEGZERO, // -0.0
    MS_VAL,     // a value (computed earlier in the program, or an external one that lives in global memory)
    MS_LOAD,    // a global row read once: selector (which run-time vector) and row
    MS_CONST    // a double shared by the batch (A and C only)
};
struct MSrc
{
    int kind = MS_ZERO;
    int val = -1;
    int sel = 0, row = 0;
    double c = 0.0;
    static MSrc zero() { return MSrc(); }
    static MSrc negzero()
    {
        MSrc s;
        s.kind = MS_NEGZERO;
        return s;
    }
    static MSrc value(int v)
    {
        MSrc s;
        s.kind = MS_VAL;
        s.val = v;
        return s;
    }
    static MSrc load(int sel, int row)
    {
        MSrc s;
        s.kind = MS_LOAD;
        s.sel = sel;
        s.row = row;
        return s;
    }
    static MSrc constant(double c)
    {
        MSrc s;
        s.kind = MS_CONST;
        s.c = c;
        return s;
    }
};
struct MOp
{
    MSrc a, b, c, x3; // x3: a further operand of the finish functor (its field travels in w6; no constants then)
    int dst = -1;     // value this operation defines
    int flags = 0;    // MF_OUT | MF_OUT2 | MF_POS | MF_RECIP | MF_FIN | kind << MF_KIND_SHIFT | MF_AONE
    int out_row = 0;  // MF_OUT: row relative to the run's out base
};
struct MVal
{
    int home_sel = -1; // load-list selector of the vector its home row belongs to (-1: no home)
    int home_row = 0;  // row inside that vector (= out_row of the operation that writes it)
};
struct MProgram
{
    std::vector<MOp> ops;   // in a valid sequential order (every value is defined before it is used)
    std::vector<MVal> vals;
    bool keep_loads = false; // gathered values with further uses may be parked in a slot (MF_BKEEP)
    int nr = 1;              // right-hand sides per pass: B, C, x3 and the destination are vectors of nr adjacent rows
                             // (one per job), A operands (matrix values, constants) are shared by the jobs
    int new_value(int home_sel = -1, int home_row = 0)
    {
        MVal v;
        v.home_sel = home_sel;
        v.home_row = home_row;
        vals.push_back(v);
        return (int)vals.size() - 1;
    }
};

// thrown when a value without a home row finds no slot (the caller may retry with home rows for those values)
struct MachineOutOfSlots : std::runtime_error
{
    using std::runtime_error::runtime_error;
};

// ---- what the compiler produces
struct MachineCode
{
    ivec ops;  // records, bundle after bundle, padded to whole chunks (+ one chunk of look-ahead)
    ivec ld;   // load list: selector << M_LD_SEL_SHIFT | row, or M_LD_NONE; padded
    int nbundles = 0, nchunks = 0, nld = 0, nld_chunks = 0;
    int nr = 1;                  // right-hand sides per pass (MProgram::nr)
    int ring_groups = 4;         // ring rows / M_RING_GROUP the code was compiled for
    int slot_budget = 0;         // slots (of nr rows) the code may use: the ring starts at row M_ROW_SLOT0 + nr * slot_budget
    int window = 0;              // look-ahead window of the scheduler that produced it
    int slot_rows = 0;           // rows used behind M_ROW_SLOT0
    long long nops = 0, nnop = 0; // real operations / padding operations
    long long far = 0, pads = 0, spills = 0; // values re-read from their home row / padding pops / partial sums sent home
};

// list-schedules the program into bundles and allocates ring rows (ring_groups groups of M_RING_GROUP) and
// slots (at most max_slots).  window: look-ahead of the scheduler, 0 = choose (for a 4-group ring and
// tune_slots >= max_slots slots, so that the order of the operations depends on the pattern alone)
void machine_compile(const MProgram &P, int max_slots, MachineCode &out, int tune_slots = 0, int ring_groups = 4, int window = 0);

} // namespace eicos
