// ECOS-name shim bound straight to the C ABI (include/eicos_b200.h): what a project without the
// reference's test/ecos.h (which goes through EiCOS::Solver, i.e. include/eicos.hpp here) uses to
// compile fixture headers written in the reference's test format.  Same names and argument order as
// the reference's shim (test/ecos.h:7-34); setup failure returns NULL, as the fixtures expect.
#pragma once
#include <cstddef>

#include "eicos_b200.h"

using idxint = int;
using pfloat = double;
using pwork = eicos_solver;

inline pwork *ECOS_setup(idxint n, idxint m, idxint p, idxint l, idxint ncones, idxint *q, idxint /*nexc*/,
                         pfloat *Gpr, idxint *Gjc, idxint *Gir, pfloat *Apr, idxint *Ajc, idxint *Air,
                         pfloat *c, pfloat *h, pfloat *b)
{
    return eicos_setup(n, m, p, l, ncones, q, Gpr, Gjc, Gir, Apr, Ajc, Air, c, h, b, /*device=*/0);
}
inline idxint ECOS_solve(pwork *w) { return eicos_solve(w); }
inline void ECOS_updateData(pwork *w, pfloat *Gpr, pfloat *Apr, pfloat *c, pfloat *h, pfloat *b)
{
    eicos_update_data(w, Gpr, Apr, c, h, b);
}
inline void ECOS_cleanup(pwork *w, idxint /*keepvars*/) { eicos_cleanup(w); }

enum
{
    ECOS_OPTIMAL = 0,
    ECOS_PINF = 1,
    ECOS_DINF = 2,
    ECOS_INACC_OFFSET = 10,
    ECOS_MAXIT = -1,
    ECOS_NUMERICS = -2,
    ECOS_OUTCONE = -3,
    ECOS_FATAL = -7
};
