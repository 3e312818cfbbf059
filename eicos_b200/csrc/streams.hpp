// Per-worker instruction streams for the device kernels.
//
// Problem they solve (profiles/r01a: issue-active 2.4 %, long-scoreboard stalls): walking CSR/CSC
// index arrays with warp-uniform loads costs one full memory latency per *index* (Lrp[i] -> Lrj[t]
// -> x[...]), and the chain cannot be overlapped.  Instead every worker (warp) of a CTA gets its own
// contiguous int32 stream that already contains, in execution order, everything it will need
// (task headers, gather rows, positions); the warp loads 32 stream words with ONE coalesced access
// (lane l holds word l) one chunk ahead of use and broadcasts them with shuffles.  Values that are
// shared by the batch (equilibrated A/G entries, +-delta) travel in a parallel double stream.
// Per-instance values (L, D, vectors) are addressed by ROW as before (layout.hpp).
//
// The layout of a stream depends on the number of workers per CTA, so streams are built by the
// engine (not by analyze()).
#pragma once

#include "layout.hpp"
#include "symbolic.hpp"

#include <vector>

namespace eicos
{

constexpr int STREAM_CHUNK = 32;   // words per cooperative load
constexpr int STREAM_PAD = 96;     // readable words after the last used one (two chunks of lookahead)
constexpr int FWD_PREV1 = -1;      // gather code: result of the previous task of this worker
constexpr int FWD_PREV2 = -2;      // ... of the task before that

struct HostStreams
{
    int workers = 1;
    // triangular sweeps: seg = [phase][worker]{int offset, tasks, first value row}
    ivec fw, fw_seg, bw, bw_seg;
    ivec fw_base, bw_base; // storage position (row inside LTx / Lx) of the first entry of row i / column j
    // factorisation
    ivec fa, fa_seg;       // seg = [phase][worker]{int offset, tasks, double offset}
    dvec fa_val;
    // mat-vec row sets: seg = [worker]{int offset, double offset}
    ivec rx, rx_seg, ry, ry_seg, rz, rz_seg, rc, rc_seg;
    dvec rx_val, ry_val, rz_val, rc_val;
};

// K-space / expanded indexing used by the row sets: x rows [0,n), y rows [n,n+p), z rows
// n+p+e with e the expanded cone index (2 unused slots after every second-order cone).
void build_streams(const Symbolic &S, int workers, HostStreams &H);

// shared values change with updateData: rebuild only the double streams
void refresh_stream_values(const Symbolic &S, HostStreams &H);

} // namespace eicos
