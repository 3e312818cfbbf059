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

// ---- triangular sweeps: dot-form row programs made of 16-byte records that the warp reads with
// one uniform 128-bit load each.  Shared-memory rows are numbered ring first ([0, FIFO_ROWS)), then
// one row of zeros (SW_ZERO_ROW), then the slots.
//   first record of a row: [number of pairs | SW_SYNC_HDR,
//                           keep row (0xFF none) | ring rows of the start operands << 8, 16, 24,
//                           (backward only) output row, pairs...]
//   further records: 4 pairs each, SW_SYNC_PAIR on the first one   (sync = issue the next FIFO group,
//                           wait for the current one; it covers every pop of its record)
//   pair word: ring row of the L value | SW_SYNC_PAIR | operand << SW_OPND_SHIFT
//              operand < SW_DIRECT: shared-memory row; else global row (operand - SW_DIRECT) of the home vector
//   the last record of a row is padded with pairs that multiply the zero row by itself.
constexpr int SW_SYNC_HDR = 1 << 31;
constexpr int SW_CNT_MASK = 0x7fffffff;
constexpr int SW_SYNC_PAIR = 1 << 8;
constexpr int SW_OPND_SHIFT = 9;
constexpr int SW_DIRECT = 256;
constexpr int SW_NO_KEEP = 0xFF;
constexpr int SW_ZERO_ROW = FIFO_ROWS;
constexpr int SW_SLOT0 = FIFO_ROWS + 1;
constexpr int SW_PAD_PAIR = SW_ZERO_ROW | (SW_ZERO_ROW << SW_OPND_SHIFT);

// ---- KKT mat-vec program (refinement residual, computeResiduals): one row program over the x, y
// and z rows of [0 A' G'; A 0 0; G 0 0] in elimination order, where the rows that share operands
// are neighbours.  Coefficients are shared by the batch and travel in a parallel stream of 16-byte
// double records; every operand row is loaded once through the FIFO and kept in a slot while it
// has further uses (Belady), so the traffic is the vectors themselves.
//   first record: [number of pairs | kind << MV_KIND_SHIFT | SW_SYNC_HDR,
//                  ring row of extra 0 | own operand << 8 | own keep << 16 | ring row of extra 1 << 24,
//                  K-space row, pair 0]            coefficients: {c0, 0}
//   further records: 4 pairs, MV_SYNC_PAIR on the first   coefficients: {c, c} {c, c}
//   pair word: operand row | keep row << 8 (0xFF none) | MV_SYNC_PAIR
//   load list bases: 1 = vector of extra 0 (rhs | c,b,h), 2 = operand vector, 3 = vector of extra 1 (LP scalings | s)
constexpr int MV_KIND_SHIFT = 24;
constexpr int MV_CNT_MASK = (1 << MV_KIND_SHIFT) - 1;
constexpr int MV_SYNC_PAIR = 1 << 16;
constexpr int MV_PAD_PAIR = SW_ZERO_ROW | (SW_NO_KEEP << 8);
// per-instance-matrices form: no coefficient stream; the first record holds no pair and its word 0
// counts the further records; pair word = ring row of the coefficient | operand << 8 | keep << 16 | MVP_SYNC
constexpr int MVP_SYNC = 1 << 30;
constexpr int MVP_PAD_PAIR = SW_ZERO_ROW | (SW_ZERO_ROW << 8) | (SW_NO_KEEP << 16);
enum MvKind : int
{
    MV_X = 0, // row of the x block: -(G' z + A' y)
    MV_Y = 1, // row of the y block: A x
    MV_Z = 2, // LP row of the z block: G x
    MV_ZC = 3 // row of a second-order cone: G x (the cone block itself is applied cone by cone afterwards)
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
    // slot programs (one warp): ops, load list (+ its length in words), shared-memory slots used
    ivec fw, fw_ld, bw, bw_ld, bwp, bwp_ld, fa, fa_ld, mv, mv_ld; // bwp: backward sweep of a plain solve (no accumulation)
    int fw_nld = 0, bw_nld = 0, bwp_nld = 0, fa_nld = 0, mv_nld = 0, mv_rows = 0;
    dvec mv_val;
    int sw_slots = 0, fa_slots = 0;
    int fa_fast = 0; // the factor program is in record form
    long long sw_far = 0, sw_direct = 0, fa_home = 0; // operands served by far gathers / direct global loads / home rows (statistics)
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
