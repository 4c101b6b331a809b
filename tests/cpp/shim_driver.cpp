// shim_driver.cpp -- TEST: drives LinSolverB200 through the LinSolverBase interface the way
// NavierStokesSolver does (navierstokes.cpp:151-164 createLinSolver/setMatrix, :566-580 solvePoisson,
// :780-788 getIters/getResidual), on a problem written by tests/test_gpu_shim.py, and writes the results
// back for comparison with the oracle.
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <memory>
#include <string>
#include <vector>

#include "linsolverb200.h"

template <typename T>
static std::vector<T> rd(const std::string &path)
{
    std::ifstream f(path, std::ios::binary | std::ios::ate);
    if (!f) { std::fprintf(stderr, "cannot open %s\n", path.c_str()); std::exit(3); }
    const size_t nb = (size_t)f.tellg();
    f.seekg(0);
    std::vector<T> v(nb / sizeof(T));
    f.read(reinterpret_cast<char *>(v.data()), (std::streamsize)nb);
    return v;
}
template <typename T>
static void wr(const std::string &path, const std::vector<T> &v)
{
    std::ofstream f(path, std::ios::binary);
    f.write(reinterpret_cast<const char *>(v.data()), (std::streamsize)(v.size() * sizeof(T)));
}

int main(int argc, char **argv)
{
    if (argc < 3) return 2;
    const std::string dir = argv[1], cfg = argv[2];
    const std::string solverName = argc > 3 ? argv[3] : "poisson";   // "velocity" / "forces": other systems of the applications
    // meta: dim, nx, ny, nz, perx, pery, perz, has_const, with_grid, nsolves
    const auto meta = rd<int>(dir + "/meta.bin");
    const auto dtv = rd<double>(dir + "/dt.bin");
    _p_Mat A;
    A.rowptr = rd<int>(dir + "/rowptr.bin");
    A.col = rd<int>(dir + "/col.bin");
    A.val = rd<double>(dir + "/val.bin");
    A.n = (int)A.rowptr.size() - 1;
    _p_MatNullSpace nsp;
    nsp.has_const = meta[7] == 1 ? PETSC_TRUE : PETSC_FALSE;
    _p_Vec nullvec;
    if (meta[7] == 2)   // one explicit null-space vector (ibpm.cpp:251-267)
    {
        nullvec.a = rd<double>(dir + "/nullvec.bin");
        nsp.vecs.push_back(&nullvec);
    }
    if (meta[7] != 0) A.nsp = &nsp;
    Mat Am = &A;

    std::shared_ptr<petibm::linsolver::LinSolverBase> solver;   // type::LinSolver
    {
        auto s = std::make_shared<petibm::linsolver::LinSolverB200>(solverName, cfg);
        if (meta[8])
        {
            const PetscInt n[3] = {meta[1], meta[2], meta[3]};
            const PetscBool per[3] = {meta[4] ? PETSC_TRUE : PETSC_FALSE, meta[5] ? PETSC_TRUE : PETSC_FALSE,
                                      meta[6] ? PETSC_TRUE : PETSC_FALSE};
            std::vector<std::vector<PetscReal>> dL = {rd<double>(dir + "/dx.bin"), rd<double>(dir + "/dy.bin")};
            if (meta[0] == 3) dL.push_back(rd<double>(dir + "/dz.bin"));
            if (s->setGridInfo(meta[0], n, per, dL, dtv[0])) return 4;
        }
        solver = s;
    }
    std::string type;
    solver->getType(type);
    if (solver->setMatrix(Am)) return 5;
    _p_Vec b, x;
    b.a = rd<double>(dir + "/b.bin");
    x.a.assign(b.a.size(), 123.0);   // initial content is ignored (zero initial guess)
    Vec bv = &b, xv = &x;
    PetscInt its = -1;
    PetscReal res = -1.0;
    PetscErrorCode ierr = 0;
    for (int q = 0; q < meta[9]; ++q) ierr = solver->solve(xv, bv);
    solver->getIters(its);
    solver->getResidual(res);
    wr(dir + "/x.bin", x.a);
    const std::string kind = static_cast<petibm::linsolver::LinSolverB200 *>(solver.get())->getOperatorKind();
    const double kindCode = kind == "stencil" ? 1.0 : kind == "hybrid" ? 2.0 : kind == "staggered" ? 3.0 : 0.0;
    wr(dir + "/out.bin", std::vector<double>{(double)its, res, (double)ierr, kindCode});
    std::printf("type=%s its=%d res=%.6e ierr=%d\n", type.c_str(), its, res, ierr);
    solver->destroy();
    return 0;
}
