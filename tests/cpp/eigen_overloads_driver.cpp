// Compiles and runs the Eigen-typed overloads of include/eicos.hpp (reference include/eicos.hpp:138-148) against the
// mock <Eigen/Sparse> of tests/cpp/mock_eigen: construct from SparseMatrix / VectorXd / VectorXi, solve, updateData
// with the Eigen signature, solution() as an Eigen::Map.  Prints "x <values>" per solve.
#include "eicos.hpp"

#include <cstdio>

#ifndef EICOS_B200_WITH_EIGEN
#error "the mock Eigen headers were not found: build with -Imock_eigen"
#endif

int main()
{
    // minimise x0 + x1  s.t.  x0 >= 1 (LP row),  ||(x1 - 0)|| <= ... a cone row block: (t; u) in Q^2 with t = x1, u = 2  ->  x1 >= 2
    // G x + s = h:  row 0: -x0 + s0 = -1;  cone rows: (-x1; 0) + (s1; s2) = (0; 2)  ->  s = (x1, 2) in Q^2  <=>  x1 >= 2
    Eigen::SparseMatrix<double> G(3, 2), A(0, 2);
    G.setCsc({-1.0, -1.0}, {0, 1, 2}, {0, 1});
    Eigen::VectorXd c{1.0, 1.0}, h{-1.0, 0.0, 2.0}, b;
    Eigen::VectorXi soc{2};
    try
    {
        EiCOS::Solver solver(G, A, c, h, b, soc);
        EiCOS::exitcode code = solver.solve();
        Eigen::Map<const Eigen::VectorXd> x = solver.solution();
        std::printf("solve %d x %.12g %.12g\n", (int)code, x(0), x(1));
        Eigen::VectorXd h2{-3.0, 0.0, 5.0};
        solver.updateData(G, A, c, h2, b);
        code = solver.solve();
        Eigen::Map<const Eigen::VectorXd> x2 = solver.solution();
        std::printf("solve %d x %.12g %.12g\n", (int)code, x2(0), x2(1));
    }
    catch (const std::exception &e)
    {
        std::fprintf(stderr, "eigen_overloads_driver: %s\n", e.what());
        return 1;
    }
    return 0;
}
