// Instruction streams for the device kernels.
//
// Problems they solve (profiles/r01a .. r01d):
//  * walking CSR/CSC index arrays with warp-uniform loads costs one full memory latency per INDEX.
//    A warp instead reads contiguous int32 streams holding, in execution order, everything it will
//    need: 32 words per coalesced access, one chunk ahead of use, broadcast with shuffles.  Values
//    shared by the batch (equilibrated A/G entries, +-delta) travel in a parallel double stream;
//    per-instance values (L, D, vectors) are addressed by ROW (layout.hpp).
//  * a warp issues in order, so a load that is consumed immediately gives a memory-level
//    parallelism of ~2.  Every global READ of the factorisation and of the triangular sweeps is
//    therefore known to the host in consumption order ("load list") and runs through a FIFO of
//    shared-memory rows filled by cp.async FIFO_ROWS ahead of the consumer (tile_program.hpp: Fifo).
//  * gathers through the solution vector / the partially factorised matrix re-read rows from HBM.
//    AMD orderings are local: in elimination order almost every value is consumed within a few
//    dozen steps of being produced.  The host therefore compiles the three numeric kernels into
//    "slot programs": every intermediate value (accumulator of a forward-sweep row, finished entry
//    of the backward sweep, Schur accumulator of an L entry) is given a shared-memory slot for its
//    live range by a linear-scan allocator; values that do not get a slot fall back to their home
//    row in HBM.  What is left as HBM traffic is the algorithmic minimum: L, D and V once, the
//    right-hand side in, the solution out.
//
// The triangular sweeps and the factorisation are run by ONE warp per tile (elimination order,
// no barriers); the mat-vec row sets are split over the workers of the CTA.
#pragma once

#include "layout.hpp"
#include "machine.hpp"
#include "symbolic.hpp"

#include <string>
#include <vector>

namespace eicos
{

constexpr long long MAX_FACTOR_UPDATES = 100LL * 1000 * 1000; // Schur updates per factorisation a factor program may hold
constexpr int STREAM_CHUNK = 32;  // streams are padded to multiples of this many words
constexpr int STREAM_PAD = 640;   // readable words after the last used one (the stream readers fetch up to four 128-word chunks ahead)
constexpr int STAGE_SLOTS = 16;   // rows per staging buffer (one slot = TILE doubles); a worker owns two buffers

// FIFO of asynchronously loaded rows: a ring of FIFO_SLOTS cp.async groups of FIFO_GROUP rows, of which
// FIFO_AHEAD are in flight ahead of the group being consumed and one is slack (the host-placed sync
// points may come up to FIFO_GROUP - 1 rows early).  It lives in the two staging buffers of worker 0.
// The host simulates the FIFO while it builds a program, so operands name their ring row directly
// and the device never counts pops.
constexpr int FIFO_GROUP = 8;
constexpr int FIFO_SLOTS = 4;
constexpr int FIFO_AHEAD = 2;
constexpr int FIFO_ROWS = FIFO_GROUP * FIFO_SLOTS;
static_assert(FIFO_ROWS == 2 * STAGE_SLOTS, "the FIFO ring aliases worker 0's staging buffers");
static_assert(FIFO_SLOTS >= FIFO_AHEAD + 2, "ring = consumed group + slack group + groups in flight");

// ---- load-list words (all programs): bits 30..31 select the base, the rest is the row
//      0 = tile base, 1 / 2 = run-time vectors of the sweep (rhs | out, accumulated solution),
//      3 = the work vector xw of the sweeps / the second extra vector of the mat-vec program
constexpr int LD_BASE_SHIFT = 30;
constexpr int LD_ROW_MASK = (1 << 30) - 1;

// ---- programs of the solveKKT / computeResiduals kernels: FMA-machine code (machine.hpp).
// Load-list selectors (bits 28.. of a load-list word; materialised into absolute tile rows per use):
//   0 = absolute row of the tile (L, 1/D, per-instance matrix values, scalar rows)
//   forward:   1 = right-hand side (KKT order), 3 = work vector xw (the program's out vector)
//   backward:  1 = output vector (the program's out vector), 2 = accumulated solution, 3 = xw
//   mat-vec:   1 = vector of the row's start value (rhs), 2 = operand vector, 3 = LP scalings, 4 = out vector e
//   residuals: 1 = [c | b | h], 2 = iterate [x | y | z], 3 = s, 4 = out vector r, 5 = scalar rows
enum MvKind : int
{
    MV_X = 0, // row of the x block: -(G' z + A' y)
    MV_Y = 1, // row of the y block: A x
    MV_Z = 2, // LP row of the z block: G x
    MV_ZC = 3 // row of a second-order cone: G x (the cone block itself is applied cone by cone afterwards)
};
// kinds the finish functors of tile_program.hpp see
enum : int
{
    FIN_ACC = 1,    // backward sweep of a refinement round: x[row] += result where the instance continues
    FIN_ABSMAX = 2, // residual row: nerr = max(nerr, |result|)
    RS_PRE_X = 3, RS_PRE_Y, RS_PRE_Z, RS_FIN_X, RS_FIN_Y, RS_FIN_Z, RS_FIRST_Z, RS_FIRST_PRE_Z // computeResiduals
};

// ---- factorisation, record form (patterns whose columns of L have at most FA_FAST_COL entries and
// whose live accumulators all fit the slots - e.g. MPC problems): 16-byte records, shared-memory rows
// named directly (ring row of a scaling value, slot of an accumulator), Schur updates unrolled on
// the device with the column in registers.
//   step k: [source of d, number of entries, source 0, source 1] [source 2, source 3, -, -] (if > 2 entries)
//           then the targets of the pairs (e1, e2 <= e1) in order, four per record
//   source word: row | kind << FA_KIND_SHIFT | FA_SYNC     kind: FA_ROW (read the row), FA_ZERO, FA_CONST (next coefficient)
//   target word: accumulator row | start row << 8 | kind << FA_KIND_SHIFT | FA_SYNC
//                (start row = the accumulator itself, or the ring row of the scaling value it starts from)
//   FA_SYNC on the first word of a record covers the pops of that record.
constexpr int FA_FAST_COL = 4;
constexpr int FA_KIND_SHIFT = 16;
constexpr int FA_SYNC = 1 << 30;
enum FaKind : int
{
    FA_ROW = 0,
    FA_ZERO = 1,
    FA_CONST = 2
};

// ---- factorisation, general form: operand codes.  code < SLOT_HOME is a shared-memory slot, anything else the
// home row (code - SLOT_HOME, relative to the tile base).  Bits 28..29 of a TARGET word say how the
// accumulator starts on its first touch.
constexpr int SLOT_HOME = 1 << 16;
constexpr int OP_CODE_MASK = (1 << 28) - 1;
constexpr int OPK_SHIFT = 28;
enum OpKind : int
{
    OPK_RMW = 0,   // accumulator already holds a value
    OPK_ZERO = 1,  // first touch, starts from 0 (fill entry)
    OPK_CONST = 2, // first touch, starts from the next word of the double stream
    OPK_FIFO = 3   // first touch, starts from the next row of the FIFO
};
// Source words (a value that is consumed): an operand code, or one of
constexpr int SRC_FIFO = -1, SRC_CONST = -2, SRC_ZERO = -3;

struct HostStreams
{
    int workers = 1;
    // machine programs: forward sweep, backward sweep (accumulating / plain), refinement residual, computeResiduals
    MachineCode fw, bw, bwp, mv, rs;
    int mv_rows = 0;
    // factor program (FIFO form): ops, load list (+ its length in words), shared-memory slots used
    ivec fa, fa_ld;
    int fa_nld = 0;
    int sw_slots = 0, fa_slots = 0; // sw_slots: rows the sweeps / mat-vecs keep in slots
    int fa_fast = 0; // the factor program is in record form
    long long sw_far = 0, sw_direct = 0, fa_home = 0; // values re-read from their home rows (sweeps) / - / home-row accumulators (factor)
    dvec fa_val;
};

// K-space / expanded indexing used by the row sets: x rows [0,n), y rows [n,n+p), z rows
// n+p+e with e the expanded cone index (2 unused slots after every second-order cone).
// max_sw_slots / max_fa_slots: shared-memory slots the sweeps / the factorisation may use.
// pim: per-instance-matrices mode - the A / G entries of the KKT matrix are rows of the workspace
// (Layout::Ax, Gx) instead of shared coefficients.
void build_streams(const Symbolic &S, const Layout &L, int workers, int max_sw_slots, int max_fa_slots, HostStreams &H,
                   bool pim = false);

// shared values change with updateData: rebuild only the double streams
void refresh_stream_values(const Symbolic &S, const Layout &L, HostStreams &H, bool pim = false);

} // namespace eicos
