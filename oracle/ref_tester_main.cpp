// Runs the reference's OWN test fixtures - test/**/*.h of EmbersArc/EiCOS, compiled unchanged from
// where they lie under /root/reference, together with the reference's own test/ecos.h shim and
// test/minunit.h - against THIS repo's drop-in: `#include "eicos.hpp"` in test/ecos.h resolves to
// include/eicos.hpp (the C++ facade over the C ABI), so every ECOS_setup / ECOS_solve /
// ECOS_updateData / ECOS_cleanup of the fixtures runs on the engine.  This file is the only part
// that is ours: the reference's tester main (test/ecostester.cpp) cannot be compiled because it
// includes MPC/MPC01.h, which is missing from the checkout, and its mu_assert aborts
// the process on the first failure - so this main runs ONE named test per process.  The 18 tests
// are the ones test/ecostester.cpp runs, minus MPC01 (infeasible2.h is dead code in the reference:
// it is not included by the tester and does not compile against test/ecos.h).
// TEST INFRASTRUCTURE: built by oracle/Makefile into oracle/_ref/ (never shipped, never loaded by the
// product).  No reference source is copied into the repository.
#include <cstdio>
#include <cstring>

#include "minunit.h"
#include "ecos.h"

#include "MPC/MPC02.h"
#include "updateData/update_data.h"
#include "cvxpyProblems/githubIssue98.h"
#include "feasibilityProblems/feas.h"
#include "unboundedProblems/unboundedLP1.h"
#include "infeasibleProblems/infeasible1.h"
#include "unboundedProblems/unboundedMaxSqrt.h"
#include "emptyProblem/emptyProblem.h"
#include "LPnetlib/lp_25fv47.h"
#include "LPnetlib/lp_adlittle.h"
#include "LPnetlib/lp_afiro.h"
#include "LPnetlib/lp_agg.h"
#include "LPnetlib/lp_agg2.h"
#include "LPnetlib/lp_agg3.h"
#include "LPnetlib/lp_bandm.h"
#include "LPnetlib/lp_beaconfd.h"
#include "LPnetlib/lp_blend.h"
#include "LPnetlib/lp_bnl1.h"

int tests_run = 0;

namespace
{
struct Entry
{
    const char *name;
    char *(*fn)();
};
const Entry TESTS[] = {
    {"MPC02", test_MPC02}, {"update_data", test_update_data}, {"unboundedLP1", test_unboundedLP1},
    {"unboundedMaxSqrt", test_unboundedMaxSqrt}, {"feas", test_feas}, {"infeasible1", test_infeasible1},
    {"lp_25fv47", test_lp_25fv47}, {"lp_adlittle", test_lp_adlittle},
    {"lp_afiro", test_lp_afiro}, {"lp_agg", test_lp_agg}, {"lp_agg2", test_lp_agg2}, {"lp_agg3", test_lp_agg3},
    {"lp_bandm", test_lp_bandm}, {"lp_beaconfd", test_lp_beaconfd}, {"lp_blend", test_lp_blend},
    {"lp_bnl1", test_lp_bnl1}, {"emptyProblem", test_emptyProblem}, {"issue98", test_issue98},
};
} // namespace

int main(int argc, char **argv)
{
    if (argc < 2)
    {
        for (const Entry &e : TESTS)
            printf("%s\n", e.name);
        return 0;
    }
    for (const Entry &e : TESTS)
        if (!strcmp(e.name, argv[1]))
        {
            try
            {
                char *msg = mu_run_test(e.fn); // mu_assert aborts the process when the expectation fails
                if (msg)
                {
                    printf("FAIL %s: %s\n", e.name, msg);
                    return 1;
                }
            }
            catch (const std::exception &ex)
            {
                printf("ERROR %s: %s\n", e.name, ex.what());
                return 2;
            }
            printf("PASS %s\n", e.name);
            return 0;
        }
    printf("unknown test %s\n", argv[1]);
    return 3;
}
