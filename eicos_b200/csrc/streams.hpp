// Per-worker instruction streams for the device kernels.
//
// Problems they solve (profiles/r01a, r01b):
//  * walking CSR/CSC index arrays with warp-uniform loads costs one full memory latency per INDEX
//    (Lrp[i] -> Lrj[t] -> x[...]).  Every worker (warp) of a CTA instead gets its own contiguous
//    int32 stream holding, in execution order, everything it will need; the warp loads 32 stream
//    words with ONE coalesced access one chunk ahead of use and broadcasts them with shuffles.
//    Values shared by the batch (equilibrated A/G entries, +-delta) travel in a parallel double
//    stream; per-instance values (L, D, vectors) are addressed by ROW (layout.hpp).
//  * warps issue in order, so a load that is consumed immediately gives a memory-level parallelism
//    of ~2 per warp.  Independent work is therefore cut into BLOCKS: the device first issues every
//    load of a block asynchronously into shared memory (cp.async), then computes.  The host decides
//    the block boundaries (slot budget) so the device never has to look ahead.
//  * rows of a serial phase (a chain of the elimination tree) mostly gather from EARLIER phases;
//    those terms are split off into a parallel "external" phase, leaving a short recurrence.
//
// The layout of a stream depends on the number of workers per CTA, so streams are built by the
// engine (not by analyze()).
#pragma once

#include "layout.hpp"
#include "symbolic.hpp"

#include <vector>

namespace eicos
{

constexpr int STREAM_CHUNK = 32;  // words per cooperative load
constexpr int STREAM_PAD = 96;    // readable words after the last used one (two chunks of lookahead)
constexpr int STAGE_SLOTS = 12;   // rows per staging buffer (one slot = TILE doubles); a worker owns two buffers
constexpr int FWD_PREV1 = -1;     // gather code: result of the previous task of this worker
constexpr int FWD_PREV2 = -2;     // ... of the task before that
constexpr int FWD_PREV3 = -3;
constexpr int INIT_PARTIAL = -1;  // task header: the start value is the partial result already stored in the output row
constexpr int FA_GROUP = 4;        // row entries of a factor task whose operands are loaded together
constexpr int ROW_EXTRA_SLOTS = 4; // staging slots a mat-vec row may use besides its gathers

enum SegKind : int
{
    SEG_BLOCKS = 0, // count = number of blocks; block = [ntasks | -1 (one oversize task)] tasks...
    SEG_SERIAL = 1  // count = number of tasks; layout hdr(0) hdr(1) ent(0) hdr(2) ent(1) ...
};

struct HostStreams
{
    int workers = 1;
    // triangular sweeps: phases in PROCESSING order; seg = [phase][worker]{int offset, count, first value row, kind}
    int nph_fw = 0, nph_bw = 0, nph_fa = 0;
    ivec fw, fw_seg, bw, bw_seg;
    ivec fw_pos, bw_pos; // storage row (inside LTx / Lx) of CSR entry t / CSC entry u
    // factorisation: seg = [phase][worker]{int offset, tasks, double offset}
    ivec fa, fa_seg;
    dvec fa_val;
    // mat-vec row sets: seg = [worker]{int offset, double offset, blocks}
    ivec rx, rx_seg, ry, ry_seg, rz, rz_seg, rc, rc_seg;
    dvec rx_val, ry_val, rz_val, rc_val;
};

// K-space / expanded indexing used by the row sets: x rows [0,n), y rows [n,n+p), z rows
// n+p+e with e the expanded cone index (2 unused slots after every second-order cone).
void build_streams(const Symbolic &S, int workers, HostStreams &H);

// shared values change with updateData: rebuild only the double streams
void refresh_stream_values(const Symbolic &S, HostStreams &H);

} // namespace eicos
