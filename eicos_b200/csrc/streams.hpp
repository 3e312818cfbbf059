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

// ---- row programs of the solveKKT kernels (triangular sweeps, KKT mat-vec), "pipe" form.
// One warp walks a stream of 16-byte records.  Two shared-memory pipes feed it, both filled by TMA
// bulk copies (cp.async.bulk, completion on mbarriers) that the warp issues for itself:
//   * the ops ring: OPS_CHUNKS chunks of 512 bytes of the record stream, refilled a chunk at a time;
//   * the data ring: RING_ROWS rows (one row = the TILE doubles of one workspace row).  Lane j owns
//     ring row j: whenever the consumer has finished with group g of RING_GROUP rows, the lanes of
//     that group copy their next rows (load list word RING_ROWS further on) into place - one 512-byte
//     bulk copy per lane, all eight in one warp instruction.  RING_ROWS - RING_GROUP rows are in flight
//     ahead of the consumer.  The host simulates the ring while it compiles a program, so every
//     operand field names its shared-memory row and the group acquire / release points are flag
//     bits of the records.
// Shared-memory rows: [0, RING_ROWS) ring, PR_ZERO (+1) rows of zeros, PR_TRASH (+1) where values
// without a slot are dumped, PR_SLOT0.. the slots.  A FIELD is 16 bits: row << 9 (the byte offset of
// the row for 512-byte rows); the low 9 bits of the first field of a record may carry flags.
// NR = 2 ("pair programs"): the same program solves two right-hand sides in one pass over L - every
// vector value (right-hand side, solution entry, slot) occupies two adjacent rows (job A, job B),
// vector pops are aligned to even ring rows, L / D values stay single.
//
// Load-list words: selector << LD_SEL_SHIFT | row; selector 0 = absolute row of the tile, 1..3 = run-time
// vectors of job A, 5..7 = the same vectors of job B (materialised per use, layout.hpp: LdVariant);
// LD_NONE = no copy (alignment padding).
constexpr int RING_ROWS = 32, RING_GROUP = 8, RING_GROUPS = RING_ROWS / RING_GROUP;
constexpr int OPS_CHUNK_WORDS = 128; // 512 bytes = 32 records
constexpr int OPS_CHUNKS = 4;
constexpr int PR_ZERO = RING_ROWS, PR_TRASH = RING_ROWS + 2, PR_SLOT0 = RING_ROWS + 4, PR_MAX_ROWS = 128;
constexpr int PR_FIELD_SHIFT = 9;
constexpr int LD_SEL_SHIFT = 29;
constexpr int LD_ROW_MASK2 = (1 << LD_SEL_SHIFT) - 1;
constexpr int LD_NONE = -1;
constexpr int LD_JOB_B = 4; // added to a selector: the vector of job B
// header word 0 (every program): counts in the low bits, then
constexpr int PH_HAS2 = 1 << 12;      // one 2-pair record follows the 4-pair records
constexpr int PH_INLINE = 1 << 13;    // the header's own pair words are in use
constexpr int PH_SLOW = 1 << 14;      // the row's pairs come one per record (operands read straight from global memory)
constexpr int PH_KIND_SHIFT = 16;     // mat-vec: MvKind (2 bits)
constexpr int PH_NACQ_SHIFT = 24;     // ring groups to acquire before the record's operands are read (2 bits)
constexpr int PH_NREL_SHIFT = 26;     // ring groups to release (= refill) after the record (2 bits)
constexpr int PH_FENCE = 1 << 28;     // the refill reads rows written earlier in this sweep: proxy fence first
constexpr int PH_END = 1 << 31;
constexpr int PH_NTAIL_MASK = 0xfff;
// tail records: the same flags in the low 9 bits of word 0
constexpr int PT_NACQ_SHIFT = 0, PT_NREL_SHIFT = 2, PT_FENCE = 1 << 4;
constexpr int PR_FIELD_MASK = 0xfe00;
//   forward row:   [ntail4 | flags, rhs field | keep field << 16, pair, pair]      pair = L field | operand field << 16
//   backward col:  [ntail4 | flags, 1/d field | xw field << 16, keep field | accumulated-solution field << 16, out row]
//   tails:         4 pairs per record (PH_HAS2: a last record with 2), padded with PR_PAD_PAIR
//   slow rows:     one record per pair [L field | flags, 0 = operand field in word 2 / 1 = global row in word 2, value, -]
//   mat-vec row:   [ngroups4 | kind | flags, extra-0 field | own field << 16, own keep field | extra-1 field << 16, K row]
//   mat-vec pairs: groups of 4 = 3 records [p, p, p, p] [c, c] [c, c]; PH_HAS2: a last group [p, p, -, -] [c, c]
//                  p = operand field | keep field << 16, c = coefficient (double)
//   (per-instance matrices: groups of 2 = 1 record [coefficient field | operand field << 16, keep field, ...] x 2)
constexpr int PR_PAD_PAIR = (PR_ZERO << PR_FIELD_SHIFT) | (PR_ZERO << (PR_FIELD_SHIFT + 16));
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

// One compiled row program: record stream (padded to whole chunks), load list (padded), and what it needs.
struct Program
{
    ivec ops, ld;
    int nchunks = 0; // 512-byte chunks of the record stream up to and including the END record
    int nld = 0;     // load-list words in use (pops, alignment padding included)
    int slot_rows = 0; // shared-memory rows it uses behind PR_SLOT0
    long long far = 0, direct = 0, pads = 0; // statistics: operands gathered through the ring / read straight from global memory / alignment padding pops
};

struct HostStreams
{
    int workers = 1;
    // pipe-form row programs, index = NR - 1 (one / two right-hand sides per pass)
    Program fw[2], bw[2], bwp[2], mv[2]; // bwp: backward sweep of a plain solve (no accumulation into the solution)
    int mv_rows = 0;
    // factor program (FIFO form): ops, load list (+ its length in words), shared-memory slots used
    ivec fa, fa_ld;
    int fa_nld = 0;
    int sw_slots = 0, fa_slots = 0; // sw_slots: values (not rows) the sweeps / mat-vec keep in slots
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
