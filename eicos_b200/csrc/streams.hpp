// The programs of the numeric kernels (machine.hpp: statically scheduled code of the FMA machine), built
// once per sparsity pattern from the symbolic analysis.
//
// What they replace (profiles/r01a .. r01i): walking CSR / CSC index arrays on the device costs a memory
// latency per index; a warp that consumes a load right after issuing it has a memory-level parallelism
// of ~2; and even with every structural decision compiled into a program, ONE in-order warp walking a
// 6 000-step dependency chain waits for its own shared-memory round trips (r01h: issue slots 24 % busy,
// short_scoreboard 0.73 per issue).  The machine programs keep M_U independent chains in flight per warp.
#pragma once

#include "layout.hpp"
#include "machine.hpp"
#include "symbolic.hpp"

#include <algorithm>
#include <cstdlib>
#include <string>
#include <vector>

namespace eicos
{

// Every program is compiled for two depths of the data ring: the shallow one when the machine is full (several
// tiles per SM share the memory latency), the deep one when there are few tiles per SM and the bytes ONE
// tile has in flight are what bounds its speed.
constexpr int M_VARIANTS = 2;
constexpr int M_VARIANT_GROUPS[M_VARIANTS] = {3, 5}; // (measured on B200: 5, 6, 8 and 10 groups are equally good for few tiles, 16 is not)
// (diagnostics) EICOS_SHALLOW_GROUPS overrides the depth of the shallow ring
inline int variant_groups(int v)
{
    if (v == 0)
        if (const char *e = std::getenv("EICOS_SHALLOW_GROUPS"))
            return std::max(2, std::min(M_MAX_RING_GROUPS, std::atoi(e)));
    return M_VARIANT_GROUPS[v];
}
constexpr int M_PAIR_GROUPS = 3; // ring depth of the two-job programs
constexpr int M_MV_PARTS = 4;    // the mat-vec programs exist split into this many independent parts (one warp each when tiles are few)
constexpr int M_PART_GROUPS = 4; // ring depth of those parts

constexpr long long MAX_FACTOR_UPDATES = 20LL * 1000 * 1000; // Schur updates per factorisation a factor program may hold (32 bytes each)

// ---- programs of the solveKKT / computeResiduals kernels: FMA-machine code (machine.hpp).
// Load-list selectors (bits 28.. of a load-list word; materialised into absolute tile rows per use):
//   0 = absolute row of the tile (L, 1/D, per-instance matrix values, scalar rows)
//   forward:   1 = right-hand side (KKT order), 3 = work vector xw (the program's out vector)
//   backward:  1 = output vector (the program's out vector), 2 = accumulated solution, 3 = xw, 4 = the forward sweep's
//              right-hand side (rows of L without entries have no xw: build_forward)
//   mat-vec:   1 = vector of the row's start value (rhs), 2 = operand vector, 3 = LP scalings, 4 = out vector e
//   residuals: 1 = [c | b | h], 2 = iterate [x | y | z], 3 = s, 4 = out vector r, 5 = scalar rows
//   factor:    0 only (the out base is the tile base)
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
    RS_PRE_X = 3, RS_PRE_Y, RS_PRE_Z, RS_FIN_X, RS_FIN_Y, RS_FIN_Z, RS_FIRST_Z, RS_FIRST_PRE_Z, // computeResiduals
    FIN_PIVOT = 11  // factorisation: 1 / pivot (infinite: the pivot was zero)
};

struct HostStreams
{
    int workers = 1;
    int sw_budget = 0, fa_budget = 0; // slot budgets the programs were compiled with
    // machine programs: forward sweep, backward sweep (accumulating / plain), refinement residual, computeResiduals
    // (index = ring variant, streams.hpp: M_VARIANT_GROUPS)
    MachineCode fw[M_VARIANTS], bw[M_VARIANTS], bwp[M_VARIANTS], mv[M_VARIANTS];
    MachineCode rs[M_MV_PARTS], mvw[M_MV_PARTS]; // computeResiduals / the refinement residual in M_MV_PARTS independent parts
    int mv_rows = 0;
    MachineCode fa[M_VARIANTS]; // numeric factorisation
    // two-job programs: the sweeps and the refinement residual for both job sets in one pass (shallow ring)
    MachineCode fw2, bw2, bwp2, mv2;
    int pair_budget = 0; // slots (of two rows) they were compiled with
    int sw_slots = 0, fa_slots = 0; // slot rows the sweeps / mat-vecs and the factorisation use
    long long sw_far = 0, fa_home = 0; // values re-read from their home rows (forward + plain backward sweep) / accumulators that wait in their home rows (factor)
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
