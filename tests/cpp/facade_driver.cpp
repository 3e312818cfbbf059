// Test driver for include/eicos.hpp (the C++ facade over the C ABI).  Reads a problem and a script of
// updateData / batched steps from a binary file written by tests/test_cpp_facade.py, runs them through
// EiCOS::Solver / EiCOS::BatchSolver exactly as the reference's own tests drive the solver
// (test/ecostester.cpp: setup -> solve -> updateData -> solve -> cleanup), and prints the results as
// text (%.17g) for the Python side to compare with the oracle.  Linked against the CUDA library for the
// gpu tier and against the CPU emulator of the kernels (tests/emu) for the CPU tier.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "eicos.hpp"

namespace
{
struct Reader
{
    FILE *f;
    int i32()
    {
        int32_t v = 0;
        if (fread(&v, 4, 1, f) != 1)
            fail();
        return v;
    }
    std::vector<int> ints(size_t k)
    {
        std::vector<int> v(k);
        if (k && fread(v.data(), 4, k, f) != k)
            fail();
        return v;
    }
    std::vector<double> dbls(size_t k)
    {
        std::vector<double> v(k);
        if (k && fread(v.data(), 8, k, f) != k)
            fail();
        return v;
    }
    [[noreturn]] static void fail()
    {
        fprintf(stderr, "facade_driver: truncated input\n");
        exit(3);
    }
};

void print_vec(const char *tag, const double *v, size_t k)
{
    printf("%s", tag);
    for (size_t i = 0; i < k; i++)
        printf(" %.17g", v[i]);
    printf("\n");
}

void report(EiCOS::Solver &S, EiCOS::exitcode code)
{
    const EiCOS::Information &I = S.getInfo();
    printf("solve %d %zu %.17g %.17g %d %d\n", (int)code, I.iter, I.pcost, I.dcost, (int)I.pinf, (int)I.dinf);
    const EiCOS::VectorView x = S.solution();
    print_vec("x", x.data(), x.size());
    std::vector<double> y, z, s;
    S.duals(y, z, s);
    print_vec("y", y.data(), y.size());
    print_vec("z", z.data(), z.size());
    print_vec("s", s.data(), s.size());
}
} // namespace

int main(int argc, char **argv)
{
    if (argc < 2)
        return 2;
    Reader R{fopen(argv[1], "rb")};
    if (!R.f)
        return 2;
    try
    {
        const int n = R.i32(), m = R.i32(), p = R.i32(), l = R.i32(), ncones = R.i32(), nnzG = R.i32(), nnzA = R.i32();
        std::vector<int> q = R.ints(ncones), Gjc = R.ints(n + 1), Gir = R.ints(nnzG), Ajc = R.ints(n + 1), Air = R.ints(nnzA);
        std::vector<double> Gpr = R.dbls(nnzG), Apr = R.dbls(nnzA), c = R.dbls(n), h = R.dbls(m), b = R.dbls(p);
        const bool hasA = p > 0;
        {
            EiCOS::Solver solver(n, m, p, l, ncones, q.data(), Gpr.data(), Gjc.data(), Gir.data(),
                                 hasA ? Apr.data() : nullptr, hasA ? Ajc.data() : nullptr, hasA ? Air.data() : nullptr,
                                 c.data(), h.data(), b.data());
            report(solver, solver.solve());
            const int nupd = R.i32();
            for (int u = 0; u < nupd; u++)
            { // bit 0..4: Gpr, Apr, c, h, b present; bit 5: the all-five overload
                const int mask = R.i32();
                std::vector<double> g2 = R.dbls(mask & 1 ? nnzG : 0), a2 = R.dbls(mask & 2 ? nnzA : 0),
                                    c2 = R.dbls(mask & 4 ? n : 0), h2 = R.dbls(mask & 8 ? m : 0), b2 = R.dbls(mask & 16 ? p : 0);
                if (mask & 32)
                    solver.updateDataFull(g2.data(), a2.data(), c2.data(), h2.data(), b2.data());
                else
                    solver.updateData(mask & 1 ? g2.data() : nullptr, mask & 2 ? a2.data() : nullptr, mask & 4 ? c2.data() : nullptr,
                                      mask & 8 ? h2.data() : nullptr, mask & 16 ? b2.data() : nullptr);
                report(solver, solver.solve());
            }
        }
        const int batch = R.i32();
        if (batch > 0)
        { // bit 0..4: Gs, As, cs, hs, bs stacks present
            const int mask = R.i32();
            const size_t B = (size_t)batch;
            std::vector<double> Gs = R.dbls(mask & 1 ? B * nnzG : 0), As = R.dbls(mask & 2 ? B * nnzA : 0),
                                cs = R.dbls(mask & 4 ? B * n : 0), hs = R.dbls(mask & 8 ? B * m : 0), bs = R.dbls(mask & 16 ? B * p : 0);
            EiCOS::BatchSolver batched(n, m, p, l, ncones, q.data(), Gpr.data(), Gjc.data(), Gir.data(),
                                       hasA ? Apr.data() : nullptr, hasA ? Ajc.data() : nullptr, hasA ? Air.data() : nullptr,
                                       c.data(), h.data(), b.data(), /*instance_matrices=*/(mask & 3) != 0);
            const EiCOS::BatchSolver::Result r =
                batched.solve(batch, mask & 4 ? cs.data() : nullptr, mask & 8 ? hs.data() : nullptr, mask & 16 ? bs.data() : nullptr,
                              mask & 1 ? Gs.data() : nullptr, mask & 2 ? As.data() : nullptr);
            for (int k = 0; k < batch; k++)
            {
                printf("batch %d %d %.17g\n", (int)r.code(k), r.info[(size_t)k].iter, r.info[(size_t)k].pcost);
                print_vec("x", r.solution(k).data(), (size_t)n);
                print_vec("y", r.y.data() + (size_t)k * p, (size_t)p);
                print_vec("z", r.z.data() + (size_t)k * m, (size_t)m);
                print_vec("s", r.s.data() + (size_t)k * m, (size_t)m);
            }
            // the same batch through the multi-device constructor (two slices; both on device 0 when that is all
            // there is): must reproduce the single-handle result bit for bit
            EiCOS::BatchSolver many(n, m, p, l, ncones, q.data(), Gpr.data(), Gjc.data(), Gir.data(),
                                    hasA ? Apr.data() : nullptr, hasA ? Ajc.data() : nullptr, hasA ? Air.data() : nullptr,
                                    c.data(), h.data(), b.data(), std::vector<int>{0, 0}, /*instance_matrices=*/(mask & 3) != 0);
            const EiCOS::BatchSolver::Result r2 =
                many.solve(batch, mask & 4 ? cs.data() : nullptr, mask & 8 ? hs.data() : nullptr, mask & 16 ? bs.data() : nullptr,
                           mask & 1 ? Gs.data() : nullptr, mask & 2 ? As.data() : nullptr);
            const bool same = many.devices() == 2 && r2.x == r.x && r2.y == r.y && r2.z == r.z && r2.s == r.s && r2.exitflag == r.exitflag;
            printf("multi-device %s\n", same ? "ok" : "MISMATCH");
        }
        // error behaviour: a malformed pattern must throw, not crash or fall back
        try
        {
            std::vector<int> bad = Gir;
            if (nnzG > 0)
            {
                bad[0] = m + 3; // row index out of range
                EiCOS::Solver broken(n, m, p, l, ncones, q.data(), Gpr.data(), Gjc.data(), bad.data(),
                                     hasA ? Apr.data() : nullptr, hasA ? Ajc.data() : nullptr, hasA ? Air.data() : nullptr,
                                     c.data(), h.data(), b.data());
                printf("error-check missing\n");
            }
            else
                printf("error-check skipped\n");
        }
        catch (const std::exception &e)
        {
            printf("error-check ok\n");
        }
    }
    catch (const std::exception &e)
    {
        fprintf(stderr, "facade_driver: %s\n", e.what());
        return 1;
    }
    return 0;
}
