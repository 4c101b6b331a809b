/**
 * \file linsolverb200.h
 * \brief LinSolverB200: the B200-native backend of petibm::linsolver::LinSolverBase.
 *
 * Third backend next to LinSolverKSP (src/linsolver/linsolverksp.cpp) and LinSolverAmgX
 * (src/linsolver/linsolveramgx.cpp) of barbagroup/PetIBM.  This file and linsolverb200.cpp are the only
 * code that sees PETSc types; everything else lives behind the C ABI of include/b200ls.h (libb200ls.so).
 * Drop the two files into PetIBM's include/petibm/ and src/linsolver/, add the factory branch shown in
 * INTEGRATION.md, build with -DPETIBM_USE_B200=ON.
 */
#pragma once

#include <string>
#include <vector>

#ifdef B200_SHIM_STUB
#include "petibm_stub.h"  // tests/cpp: minimal stand-ins for <petscksp.h> and <petibm/linsolver.h>
#else
#include <petscksp.h>

#include <petibm/linsolver.h>
#endif

#include "b200ls.h"

namespace petibm
{
namespace linsolver
{
class LinSolverB200 : public LinSolverBase
{
public:
    /** Same constructor as LinSolverKSP (linsolverksp.cpp:16-20): name + options file. */
    LinSolverB200(const std::string &solverName, const std::string &file);

    ~LinSolverB200() override;

    PetscErrorCode destroy() override;

    /**
     * Grid description the factory hands over (createLinSolver has the whole YAML node: mesh,
     * boundary conditions, dt -- linsolver.cpp:57-58).  dL are the pressure-cell widths per axis exactly
     * as parser::parseMesh returns them.  Optional: without it every matrix takes the general CSR path.
     */
    PetscErrorCode setGridInfo(const PetscInt &dim, const PetscInt n[3], const PetscBool periodic[3],
                               const std::vector<std::vector<PetscReal>> &dL, const PetscReal &dt);

    PetscErrorCode setMatrix(const Mat &A) override;
    PetscErrorCode solve(Vec &x, Vec &b) override;
    PetscErrorCode getIters(PetscInt &iters) override;
    PetscErrorCode getResidual(PetscReal &res) override;

    /** "stencil" (matrix-free separable pressure operator verified against A), "hybrid" (that operator followed by
     *  IBPM's coupling rows), "staggered" (line-coefficient form read out of A: velocity system) or "csr". */
    const std::string &getOperatorKind() const { return opKind; }

protected:
    PetscErrorCode init() override;

    b200ls_solver *handle = nullptr;
    PetscErrorCode initStatus = 0;  // a constructor cannot return the error of init(); every call re-raises it
    bool haveGrid = false;
    PetscInt gdim = 0;
    int64_t gn[3] = {1, 1, 1};
    int gper[3] = {0, 0, 0};
    std::vector<double> gdL[3];
    double gdt = 0.0;
    PetscMPIInt rank = 0, nranks = 1;
    std::string opKind = "none";
    // several ranks: exchange between the DMDA boxes the Vecs arrive in and the solver's slabs
    b200ls_repart *plan = nullptr;
    bool planIdentity = true;               // 1 x 1 x P process grid: the boxes are the slabs
    std::vector<PetscMPIInt> xcounts[4];    // box counts / displs, slab counts / displs (MPI_Alltoallv)
    std::vector<double> xbuf, bslab, xslab; // slab-side exchange buffer and slab-ordered b / x
    // several ranks and a matrix that is not the pressure stencil (velocity system, IBPM's modified Poisson system,
    // forces system): every rank holds a replica of the whole system on its GPU (single-rank handle); solve()
    // all-gathers b and keeps the caller's rows of x
    bool replicated = false;
    std::vector<PetscMPIInt> repCounts, repDispls;  // rows per rank / first row of every rank (MPI_Allgatherv)
    std::vector<double> repB, repX;
    // several ranks and -pc_type mg: the multigrid preconditioner runs on one GPU, so every rank sets up the WHOLE grid and
    // solves a replica; solve() moves b boxes -> slabs (plan), all-gathers the slabs (rank order = natural ordering) and
    // keeps its own slab of x
    bool mgReplica = false;
    b200ls_options savedOpts;                        // to re-create the handle (replica <-> distributed)
    PetscErrorCode newHandle(bool withComm);
};  // LinSolverB200

}  // end of namespace linsolver
}  // end of namespace petibm
