/**
 * \file linsolverb200.cpp
 * \brief Implementation of LinSolverB200 (see linsolverb200.h).
 *
 * Call protocol is the one the applications use (navierstokes.cpp:151-164, 566-580, 780-788):
 * createLinSolver -> setMatrix once (re-called every step by rigidkinematics.cpp:135) -> solve per time
 * step with a zero initial guess -> getIters / getResidual.  Errors are PetscErrorCodes; a diverged solve
 * raises PETSC_ERR_CONV_FAILED exactly like LinSolverKSP::solve (linsolverksp.cpp:96-104).
 */
#include "linsolverb200.h"

#include <cstring>
#include <fstream>
#include <sstream>

namespace petibm
{
namespace linsolver
{
namespace
{
// b200ls status -> PETSc error
#define B200CHK(h, call)                                                                                       \
    do                                                                                                         \
    {                                                                                                          \
        int rc__ = (call);                                                                                     \
        if (rc__ != B200LS_OK)                                                                                 \
            SETERRQ3(PETSC_COMM_WORLD, rc__ == B200LS_ERR_UNSUPPORTED ? PETSC_ERR_SUP : PETSC_ERR_LIB,         \
                     "B200 linear solver: %s failed: %s (%s)", #call, b200ls_error_string(rc__),               \
                     b200ls_last_error(h));                                                                    \
    } while (0)
}  // namespace

LinSolverB200::LinSolverB200(const std::string &solverName, const std::string &file)
    : LinSolverBase(solverName, file)
{
    initStatus = init();
}

LinSolverB200::~LinSolverB200()
{
    PetscBool finalized;
    PetscErrorCode ierr = PetscFinalized(&finalized);
    if (ierr || finalized) return;  // same guard as linsolverksp.cpp:30-31
    if (plan) b200ls_repart_destroy(plan);
    plan = nullptr;
    if (handle) b200ls_destroy(handle);
    handle = nullptr;
}

PetscErrorCode LinSolverB200::destroy()
{
    PetscErrorCode ierr;
    PetscFunctionBeginUser;
    if (plan) b200ls_repart_destroy(plan);
    plan = nullptr;
    if (handle) b200ls_destroy(handle);
    handle = nullptr;
    ierr = LinSolverBase::destroy(); CHKERRQ(ierr);
    PetscFunctionReturn(0);
}

// LinSolverKSP::init (linsolverksp.cpp:48-69): the options file is read with the prefix "<name>_".
PetscErrorCode LinSolverB200::init()
{
    PetscErrorCode ierr;
    PetscFunctionBeginUser;

    // The applications dispatch their null-space handling on this string and abort on anything else
    // (navierstokes.cpp:401-426, ibpm.cpp:248-280).  Reporting the KSP type makes them attach the
    // MatNullSpace to DBNG, which setMatrix reads back -- the behaviour this backend reproduces.
    type = "PETSc KSP";

    ierr = MPI_Comm_rank(PETSC_COMM_WORLD, &rank); CHKERRQ(ierr);
    ierr = MPI_Comm_size(PETSC_COMM_WORLD, &nranks); CHKERRQ(ierr);

    int ndev = 0;
    if (b200ls_device_count(&ndev) != B200LS_OK || ndev <= 0)
        SETERRQ(PETSC_COMM_WORLD, PETSC_ERR_SUP, "B200 linear solver: no CUDA device (there is no CPU fallback).");
    int rc = b200ls_create(&handle, rank % ndev);
    if (rc != B200LS_OK)
        SETERRQ1(PETSC_COMM_WORLD, PETSC_ERR_LIB, "B200 linear solver: b200ls_create failed: %s", b200ls_error_string(rc));

    b200ls_options opts;
    b200ls_default_options(&opts);
    if (config != "None")
    {
        std::ifstream f(config);
        if (!f) SETERRQ1(PETSC_COMM_WORLD, PETSC_ERR_FILE_OPEN, "Could not open the solver configuration file %s", config.c_str());
        std::stringstream ss;
        ss << f.rdbuf();
        char err[512] = {0};
        rc = b200ls_parse_options(ss.str().c_str(), (name + "_").c_str(), &opts, err, sizeof err);
        if (rc != B200LS_OK)
            SETERRQ2(PETSC_COMM_WORLD, rc == B200LS_ERR_UNSUPPORTED ? PETSC_ERR_SUP : PETSC_ERR_ARG_WRONG,
                     "B200 linear solver \"%s\": %s", name.c_str(), err);
    }
    B200CHK(handle, b200ls_set_options(handle, &opts));
    savedOpts = opts;
    if (nranks > 1) B200CHK(handle, b200ls_comm_init(handle, rank, nranks, B200LS_REDUCE_P2P, B200LS_HALO_STORE));
    PetscFunctionReturn(0);
}

// A fresh handle with the options of init(): single-rank (a replica of a general system) or wired to the communicator
// again.  Collective: every rank first drops its mappings of the peers' exchange arenas.
PetscErrorCode LinSolverB200::newHandle(bool withComm)
{
    PetscErrorCode ierr;
    PetscFunctionBeginUser;
    B200CHK(handle, b200ls_comm_disconnect(handle));
    ierr = MPI_Barrier(PETSC_COMM_WORLD); CHKERRQ(ierr);
    b200ls_destroy(handle);
    handle = nullptr;
    int ndev = 0;
    if (b200ls_device_count(&ndev) != B200LS_OK || ndev <= 0)
        SETERRQ(PETSC_COMM_WORLD, PETSC_ERR_SUP, "B200 linear solver: no CUDA device (there is no CPU fallback).");
    const int rc = b200ls_create(&handle, rank % ndev);
    if (rc != B200LS_OK)
        SETERRQ1(PETSC_COMM_WORLD, PETSC_ERR_LIB, "B200 linear solver: b200ls_create failed: %s", b200ls_error_string(rc));
    B200CHK(handle, b200ls_set_options(handle, &savedOpts));
    if (withComm && nranks > 1) B200CHK(handle, b200ls_comm_init(handle, rank, nranks, B200LS_REDUCE_P2P, B200LS_HALO_STORE));
    PetscFunctionReturn(0);
}

PetscErrorCode LinSolverB200::setGridInfo(const PetscInt &dim, const PetscInt n[3], const PetscBool periodic[3],
                                          const std::vector<std::vector<PetscReal>> &dL, const PetscReal &dt)
{
    PetscFunctionBeginUser;
    gdim = dim;
    for (int d = 0; d < 3; ++d)
    {
        gn[d] = (d < dim) ? n[d] : 1;
        gper[d] = (d < dim && periodic[d]) ? 1 : 0;
        gdL[d].assign(d < dim ? dL[d].begin() : dL[0].begin(), d < dim ? dL[d].end() : dL[0].begin());
    }
    gdt = dt;
    haveGrid = true;
    PetscFunctionReturn(0);
}

// LinSolverKSP::setMatrix (linsolverksp.cpp:72-82).  The matrix is read once (like AmgXSolver::setA,
// linsolveramgx.cpp:84); the caller keeps ownership.
PetscErrorCode LinSolverB200::setMatrix(const Mat &A)
{
    PetscErrorCode ierr;
    PetscFunctionBeginUser;
    if (initStatus) SETERRQ1(PETSC_COMM_WORLD, initStatus, "B200 linear solver \"%s\" was not initialised.", name.c_str());

    PetscInt rbeg, rend;
    ierr = MatGetOwnershipRange(A, &rbeg, &rend); CHKERRQ(ierr);
    const PetscInt nloc = rend - rbeg;

    // local rows -> CSR with global column indices
    std::vector<int64_t> rowptr((size_t)nloc + 1, 0);
    std::vector<int32_t> col;
    std::vector<double> val;
    col.reserve((size_t)nloc * 7);
    val.reserve((size_t)nloc * 7);
    for (PetscInt r = rbeg; r < rend; ++r)
    {
        PetscInt nc;
        const PetscInt *cols;
        const PetscScalar *vals;
        ierr = MatGetRow(A, r, &nc, &cols, &vals); CHKERRQ(ierr);
        for (PetscInt q = 0; q < nc; ++q)
        {
            col.push_back((int32_t)cols[q]);
            val.push_back((double)vals[q]);
        }
        rowptr[(size_t)(r - rbeg) + 1] = (int64_t)col.size();
        ierr = MatRestoreRow(A, r, &nc, &cols, &vals); CHKERRQ(ierr);
    }

    // 1. try the matrix-free separable operator, verified entry by entry against A
    bool recognised = false;
    if (plan) b200ls_repart_destroy(plan);
    plan = nullptr;
    if (replicated)
    {
        // the previous matrix was solved as replicas on a single-rank handle: back to a distributed one
        ierr = newHandle(true); CHKERRQ(ierr);
        replicated = false;
        mgReplica = false;
    }
    if (haveGrid && nranks == 1)
    {
        const int64_t nslow = (gdim == 3) ? gn[2] : 1;   // one GPU owns the whole grid (2-D: a single "plane")
        if ((int64_t)nloc == gn[0] * gn[1] * nslow)
        {
            B200CHK(handle, b200ls_set_poisson_stencil(handle, (int)gdim, gn, gper, gdL[0].data(), gdL[1].data(),
                                                       gdim == 3 ? gdL[2].data() : nullptr, gdt, 0, nslow));
            double diff = 0.0;
            const int rc = b200ls_verify_csr(handle, nloc, rowptr.data(), col.data(), val.data(), &diff);
            if (rc == B200LS_OK) recognised = true;
            else if (rc != B200LS_ERR_MISMATCH) B200CHK(handle, rc);
        }
    }
    else if (haveGrid)
    {
        // Several ranks.  The rows are this rank's DMDA box and the columns PETSc global indices
        // (cartesianmesh.cpp:500-538, 709-721); the device solver works on slabs along the slowest axis (z in 3-D,
        // y in 2-D).  The Mat does not carry its DMDA, so every process grid that reproduces the ranks' local sizes
        // is tried against the matrix entries; the one that verifies fixes the box <-> slab exchange of solve().
        std::vector<int64_t> sizes((size_t)nranks, 0);
        const int64_t mine = (int64_t)nloc;
        ierr = MPI_Allgather(&mine, (int)sizeof(int64_t), MPI_BYTE, sizes.data(), (int)sizeof(int64_t), MPI_BYTE,
                             PETSC_COMM_WORLD); CHKERRQ(ierr);
        std::vector<int> cands(3 * 64, 1);
        int found = 0;
        B200CHK(handle, b200ls_repart_candidates((int)gdim, gn, nranks, sizes.data(), cands.data(), 64, &found));
        if (found > 64) found = 64;
        const int64_t nslow = (gdim == 3) ? gn[2] : gn[1];
        mgReplica = false;
        if (found > 0 && nslow >= nranks)
        {
            const int64_t base = nslow / nranks, rem = nslow % nranks;
            const int64_t lo = rank * base + (rank < rem ? rank : rem), hi = lo + base + (rank < rem ? 1 : 0);
            if (savedOpts.pc_type == B200LS_PC_MG)
            {
                // the multigrid preconditioner runs on one GPU: a single-rank handle with the whole grid on every rank
                ierr = newHandle(false); CHKERRQ(ierr);
                replicated = true;   // (the handle is single-rank: a later setMatrix re-creates the distributed one)
                mgReplica = true;
                B200CHK(handle, b200ls_set_poisson_stencil(handle, (int)gdim, gn, gper, gdL[0].data(), gdL[1].data(),
                                                           gdim == 3 ? gdL[2].data() : nullptr, gdt, 0, gdim == 3 ? gn[2] : 1));
            }
            else
            {
                // a second setMatrix replaces the exchange arena: drop the peer mappings everywhere first
                B200CHK(handle, b200ls_comm_disconnect(handle));
                ierr = MPI_Barrier(PETSC_COMM_WORLD); CHKERRQ(ierr);
                B200CHK(handle, b200ls_set_poisson_stencil(handle, (int)gdim, gn, gper, gdL[0].data(), gdL[1].data(),
                                                           gdim == 3 ? gdL[2].data() : nullptr, gdt, lo, hi));
            }
            std::vector<int64_t> natRows((size_t)nloc);
            std::vector<int32_t> natCols(col.size());
            for (int c = 0; c < found && !recognised; ++c)
            {
                b200ls_repart *cand = nullptr;
                B200CHK(handle, b200ls_repart_create(&cand, (int)gdim, gn, &cands[(size_t)3 * c], rank));
                int rc = b200ls_repart_box_rows(cand, natRows.data());
                if (rc == B200LS_OK) rc = b200ls_repart_petsc_to_natural(cand, (int64_t)col.size(), col.data(), natCols.data());
                double diff = 0.0;
                // diagonal within 4 ulp: PETSc's parallel MatMatMult adds the off-process terms of a diagonal entry
                // after the local ones, so rows on a partition boundary round differently per partition (b200ls.h)
                if (rc == B200LS_OK)
                    rc = b200ls_verify_csr_rows(handle, nloc, natRows.data(), rowptr.data(), natCols.data(), val.data(), 4, &diff);
                // every rank must agree, otherwise the exchange plans would be mismatched
                PetscMPIInt ok = (rc == B200LS_OK) ? 1 : 0, all = 0;
                ierr = MPI_Allreduce(&ok, &all, 1, MPI_INT, MPI_MIN, PETSC_COMM_WORLD); CHKERRQ(ierr);
                if (all == 1)
                {
                    plan = cand;
                    recognised = true;
                }
                else
                    b200ls_repart_destroy(cand);
            }
        }
    }
    if (recognised)
    {
        opKind = "stencil";
        if (nranks > 1)
        {
            if (!mgReplica)
            {
                // exchange the CUDA IPC handles of the exchange arenas: MPI is the host transport only
                std::vector<char> mine(64), all((size_t)64 * nranks);
                B200CHK(handle, b200ls_comm_export(handle, mine.data()));
                ierr = MPI_Allgather(mine.data(), 64, MPI_BYTE, all.data(), 64, MPI_BYTE, PETSC_COMM_WORLD); CHKERRQ(ierr);
                B200CHK(handle, b200ls_comm_connect(handle, all.data(), nranks));
            }
            // box <-> slab exchange of solve(): counts/displacements for MPI_Alltoallv and the slab-side buffers
            int identity = 1;
            int64_t nslab = 0;
            B200CHK(handle, b200ls_repart_info(plan, nullptr, nullptr, nullptr, nullptr, nullptr, &nslab, &identity));
            planIdentity = identity != 0;
            std::vector<int64_t> c64[4];
            for (auto &v : c64) v.assign((size_t)nranks, 0);
            B200CHK(handle, b200ls_repart_counts(plan, c64[0].data(), c64[1].data(), c64[2].data(), c64[3].data()));
            for (int q = 0; q < 4; ++q)
            {
                xcounts[q].assign((size_t)nranks, 0);
                for (int r = 0; r < nranks; ++r)
                {
                    if (c64[q][(size_t)r] > 2147483647LL)
                        SETERRQ(PETSC_COMM_WORLD, PETSC_ERR_SUP, "B200 linear solver: exchange block exceeds the MPI count range.");
                    xcounts[q][(size_t)r] = (PetscMPIInt)c64[q][(size_t)r];
                }
            }
            if (!planIdentity || mgReplica)
            {
                xbuf.assign((size_t)nslab, 0.0);
                bslab.assign((size_t)nslab, 0.0);
                xslab.assign((size_t)nslab, 0.0);
            }
            if (mgReplica)
            {
                // slab sizes of all ranks (rank order = natural ordering) for the all-gather of b in solve()
                const long long mineSlab = (long long)nslab;
                std::vector<long long> allSlab((size_t)nranks);
                ierr = MPI_Allgather(&mineSlab, (int)sizeof(long long), MPI_BYTE, allSlab.data(), (int)sizeof(long long), MPI_BYTE,
                                     PETSC_COMM_WORLD); CHKERRQ(ierr);
                repCounts.assign((size_t)nranks, 0);
                repDispls.assign((size_t)nranks, 0);
                long long tot = 0;
                for (int r = 0; r < nranks; ++r)
                {
                    if (allSlab[(size_t)r] > 2147483647LL || tot > 2147483647LL)
                        SETERRQ(PETSC_COMM_WORLD, PETSC_ERR_SUP, "B200 linear solver: replicated grid exceeds the MPI count range.");
                    repCounts[(size_t)r] = (PetscMPIInt)allSlab[(size_t)r];
                    repDispls[(size_t)r] = (PetscMPIInt)tot;
                    tot += allSlab[(size_t)r];
                }
                repB.assign((size_t)tot, 0.0);
                repX.assign((size_t)tot, 0.0);
            }
        }
    }
    else
    {
        int64_t nlocRows = (int64_t)nloc;
        if (nranks > 1)
        {
            // Any other system on several ranks (LinSolverKSP solves the velocity, modified-Poisson and forces systems on
            // any rank count, linsolverksp.cpp:72-105): gather the rows (global column indices, so the rank-ordered
            // concatenation IS the global matrix in PETSc ordering) and solve a replica on every GPU.
            std::vector<PetscMPIInt> nnzCounts((size_t)nranks), nnzDispls((size_t)nranks);
            repCounts.assign((size_t)nranks, 0);
            repDispls.assign((size_t)nranks, 0);
            const long long mine[2] = {(long long)nloc, (long long)col.size()};
            std::vector<long long> all((size_t)2 * nranks);
            ierr = MPI_Allgather(mine, 2 * (int)sizeof(long long), MPI_BYTE, all.data(), 2 * (int)sizeof(long long), MPI_BYTE,
                                 PETSC_COMM_WORLD); CHKERRQ(ierr);
            long long rowsTotal = 0, nnzTotal = 0;
            for (int r = 0; r < nranks; ++r)
            {
                if (all[2 * r] > 2147483647LL || all[2 * r + 1] > 2147483647LL || nnzTotal + all[2 * r + 1] > 2147483647LL)
                    SETERRQ(PETSC_COMM_WORLD, PETSC_ERR_SUP, "B200 linear solver: replicated system exceeds the MPI count range.");
                repCounts[(size_t)r] = (PetscMPIInt)all[2 * r];
                repDispls[(size_t)r] = (PetscMPIInt)rowsTotal;
                nnzCounts[(size_t)r] = (PetscMPIInt)all[2 * r + 1];
                nnzDispls[(size_t)r] = (PetscMPIInt)nnzTotal;
                rowsTotal += all[2 * r];
                nnzTotal += all[2 * r + 1];
            }
            std::vector<int32_t> gcol((size_t)nnzTotal);
            std::vector<double> gval((size_t)nnzTotal);
            std::vector<long long> rowLen((size_t)nloc), gLen((size_t)rowsTotal);
            for (PetscInt r = 0; r < nloc; ++r) rowLen[(size_t)r] = rowptr[(size_t)r + 1] - rowptr[(size_t)r];
            ierr = MPI_Allgatherv(col.data(), (PetscMPIInt)col.size(), MPI_INT, gcol.data(), nnzCounts.data(), nnzDispls.data(),
                                  MPI_INT, PETSC_COMM_WORLD); CHKERRQ(ierr);
            ierr = MPI_Allgatherv(val.data(), (PetscMPIInt)val.size(), MPI_DOUBLE, gval.data(), nnzCounts.data(),
                                  nnzDispls.data(), MPI_DOUBLE, PETSC_COMM_WORLD); CHKERRQ(ierr);
            ierr = MPI_Allgatherv(rowLen.data(), (PetscMPIInt)nloc, MPI_LONG_LONG, gLen.data(), repCounts.data(),
                                  repDispls.data(), MPI_LONG_LONG, PETSC_COMM_WORLD); CHKERRQ(ierr);
            rowptr.assign((size_t)rowsTotal + 1, 0);
            for (long long r = 0; r < rowsTotal; ++r) rowptr[(size_t)r + 1] = rowptr[(size_t)r] + gLen[(size_t)r];
            col.swap(gcol);
            val.swap(gval);
            nlocRows = rowsTotal;
            ierr = newHandle(false); CHKERRQ(ierr);
            replicated = true;
            mgReplica = false;
            repB.assign((size_t)rowsTotal, 0.0);
            repX.assign((size_t)rowsTotal, 0.0);
        }
        const int64_t nloc = nlocRows;  // from here on: the rows this handle holds (all of them for a replica)
        // 2. a staggered-grid matrix with one-dimensional coefficients: the packed velocity system [u | v | w]
        //    (cartesianmesh.cpp:251-273: one point fewer than cells along the field's own direction unless periodic)
        //    or the pressure block followed by IBPM's Lagrangian force rows (ibpm.cpp:164-194).  The structure is read
        //    out of A and verified against every entry by b200ls_set_staggered.
        bool staggered = false, hybrid = false;
        if (haveGrid && !replicated)
        {
            const int64_t n3[3] = {gn[0], gn[1], gdim == 3 ? gn[2] : 1};
            int64_t vel[9], velTotal = 0;
            bool velOk = true;
            for (int f = 0; f < (int)gdim; ++f)
            {
                int64_t sz = 1;
                for (int d = 0; d < 3; ++d)
                {
                    vel[3 * f + d] = n3[d] - ((d == f && !gper[d]) ? 1 : 0);
                    velOk = velOk && vel[3 * f + d] >= 1;
                    sz *= vel[3 * f + d];
                }
                velTotal += sz;
            }
            const int64_t pN = n3[0] * n3[1] * n3[2];
            int rc = B200LS_ERR_MISMATCH;
            if (velOk && (int64_t)nloc == velTotal)
                rc = b200ls_set_staggered(handle, (int)gdim, vel, gper, nloc, rowptr.data(), col.data(), val.data());
            if (rc == B200LS_ERR_MISMATCH && (int64_t)nloc > pN)
            {
                // IBPM: the pressure operator of the mesh (stretched grid: coefficients with face areas) + coupling rows
                rc = b200ls_set_poisson_hybrid(handle, (int)gdim, gn, gper, gdL[0].data(), gdL[1].data(),
                                               gdim == 3 ? gdL[2].data() : nullptr, gdt, nloc, rowptr.data(), col.data(), val.data());
                if (rc == B200LS_OK) hybrid = true;
            }
            if (rc == B200LS_ERR_MISMATCH && (int64_t)nloc > pN)
                rc = b200ls_set_staggered(handle, 1, n3, gper, nloc, rowptr.data(), col.data(), val.data());
            if (rc == B200LS_OK) staggered = true;
            else if (rc != B200LS_ERR_MISMATCH) B200CHK(handle, rc);
        }
        if (staggered)
            opKind = hybrid ? "hybrid" : "staggered";
        else
        {
            // 3. verified fallback: the assembled operator itself, still on the GPU (BN order > 1, anything else)
            B200CHK(handle, b200ls_set_csr(handle, nloc, rowptr.data(), col.data(), val.data()));
            opKind = "csr";
        }
    }

    // 4. the null space the application attached with MatSetNullSpace (navierstokes.cpp:404-413, ibpm.cpp:251-267)
    MatNullSpace nsp = nullptr;
    ierr = MatGetNullSpace(A, &nsp); CHKERRQ(ierr);
    if (nsp)
    {
        PetscBool hasConst;
        PetscInt nv;
        const Vec *vecs;
        ierr = MatNullSpaceGetVecs(nsp, &hasConst, &nv, &vecs); CHKERRQ(ierr);
        std::vector<double> flat;
        for (PetscInt q = 0; q < nv; ++q)
        {
            const PetscScalar *arr;
            ierr = VecGetArrayRead(vecs[q], &arr); CHKERRQ(ierr);
            if (replicated)
            {
                const size_t at = flat.size();
                flat.resize(at + repB.size());
                ierr = MPI_Allgatherv(arr, (PetscMPIInt)nloc, MPI_DOUBLE, flat.data() + at, repCounts.data(), repDispls.data(),
                                      MPI_DOUBLE, PETSC_COMM_WORLD); CHKERRQ(ierr);
            }
            else
                flat.insert(flat.end(), arr, arr + nloc);
            ierr = VecRestoreArrayRead(vecs[q], &arr); CHKERRQ(ierr);
        }
        B200CHK(handle, b200ls_set_nullspace(handle, hasConst ? 1 : 0, (int)nv, nv ? flat.data() : nullptr));
    }
    else
        B200CHK(handle, b200ls_set_nullspace(handle, 0, 0, nullptr));
    PetscFunctionReturn(0);
}

// LinSolverKSP::solve (linsolverksp.cpp:85-105): KSPSolve with a zero initial guess, fatal if reason < 0.
PetscErrorCode LinSolverB200::solve(Vec &x, Vec &b)
{
    PetscErrorCode ierr;
    PetscFunctionBeginUser;
    if (initStatus) SETERRQ1(PETSC_COMM_WORLD, initStatus, "B200 linear solver \"%s\" was not initialised.", name.c_str());
    const PetscScalar *barr;
    PetscScalar *xarr;
    ierr = VecGetArrayRead(b, &barr); CHKERRQ(ierr);
    ierr = VecGetArray(x, &xarr); CHKERRQ(ierr);
    int rc;
    if (mgReplica)
    {
        // DMDA boxes -> slabs (if needed), all-gather the slabs, solve the whole grid on this GPU, own slab -> boxes
        const double *bs = barr;
        if (plan && !planIdentity)
        {
            ierr = MPI_Alltoallv(barr, xcounts[0].data(), xcounts[1].data(), MPI_DOUBLE, xbuf.data(), xcounts[2].data(),
                                 xcounts[3].data(), MPI_DOUBLE, PETSC_COMM_WORLD); CHKERRQ(ierr);
            B200CHK(handle, b200ls_repart_unpack_slab(plan, xbuf.data(), bslab.data()));
            bs = bslab.data();
        }
        ierr = MPI_Allgatherv(bs, repCounts[(size_t)rank], MPI_DOUBLE, repB.data(), repCounts.data(), repDispls.data(),
                              MPI_DOUBLE, PETSC_COMM_WORLD); CHKERRQ(ierr);
        rc = b200ls_solve(handle, repB.data(), repX.data());
        if (rc == B200LS_OK || rc == B200LS_ERR_DIVERGED)
        {
            const double *xs = repX.data() + repDispls[(size_t)rank];
            if (plan && !planIdentity)
            {
                B200CHK(handle, b200ls_repart_pack_slab(plan, xs, xbuf.data()));
                ierr = MPI_Alltoallv(xbuf.data(), xcounts[2].data(), xcounts[3].data(), MPI_DOUBLE, xarr, xcounts[0].data(),
                                     xcounts[1].data(), MPI_DOUBLE, PETSC_COMM_WORLD); CHKERRQ(ierr);
            }
            else
                std::memcpy(xarr, xs, sizeof(double) * (size_t)repCounts[(size_t)rank]);
        }
    }
    else if (replicated)
    {
        // every rank solves the whole system on its GPU: all-gather b, keep the own rows of x
        ierr = MPI_Allgatherv(barr, repCounts[(size_t)rank], MPI_DOUBLE, repB.data(), repCounts.data(), repDispls.data(),
                              MPI_DOUBLE, PETSC_COMM_WORLD); CHKERRQ(ierr);
        rc = b200ls_solve(handle, repB.data(), repX.data());
        if (rc == B200LS_OK || rc == B200LS_ERR_DIVERGED)
            std::memcpy(xarr, repX.data() + repDispls[(size_t)rank], sizeof(double) * (size_t)repCounts[(size_t)rank]);
    }
    else if (plan && !planIdentity)
    {
        // DMDA boxes -> slabs, solve, slabs -> boxes: one MPI_Alltoallv per direction (what VecScatter does inside
        // PETSc's MatMult); the box side needs no packing (b200ls.h, b200ls_repart_*)
        ierr = MPI_Alltoallv(barr, xcounts[0].data(), xcounts[1].data(), MPI_DOUBLE, xbuf.data(), xcounts[2].data(),
                             xcounts[3].data(), MPI_DOUBLE, PETSC_COMM_WORLD); CHKERRQ(ierr);
        B200CHK(handle, b200ls_repart_unpack_slab(plan, xbuf.data(), bslab.data()));
        rc = b200ls_solve(handle, bslab.data(), xslab.data());
        if (rc == B200LS_OK || rc == B200LS_ERR_DIVERGED)
        {
            B200CHK(handle, b200ls_repart_pack_slab(plan, xslab.data(), xbuf.data()));
            ierr = MPI_Alltoallv(xbuf.data(), xcounts[2].data(), xcounts[3].data(), MPI_DOUBLE, xarr, xcounts[0].data(),
                                 xcounts[1].data(), MPI_DOUBLE, PETSC_COMM_WORLD); CHKERRQ(ierr);
        }
    }
    else
        rc = b200ls_solve(handle, barr, xarr);
    ierr = VecRestoreArray(x, &xarr); CHKERRQ(ierr);
    ierr = VecRestoreArrayRead(b, &barr); CHKERRQ(ierr);
    if (rc == B200LS_ERR_DIVERGED)
    {
        int reason = 0;
        b200ls_get_reason(handle, &reason);
        SETERRQ2(PETSC_COMM_WORLD, PETSC_ERR_CONV_FAILED,
                 "PetIBM exited due to B200 solver %s diverged with reason %d.", name.c_str(), reason);
    }
    B200CHK(handle, rc);
    PetscFunctionReturn(0);
}

PetscErrorCode LinSolverB200::getIters(PetscInt &iters)
{
    PetscFunctionBeginUser;
    int its = 0;
    B200CHK(handle, b200ls_get_iters(handle, &its));
    iters = its;
    PetscFunctionReturn(0);
}

PetscErrorCode LinSolverB200::getResidual(PetscReal &res)
{
    PetscFunctionBeginUser;
    double r = 0.0;
    B200CHK(handle, b200ls_get_residual(handle, &r));
    res = r;
    PetscFunctionReturn(0);
}

}  // end of namespace linsolver
}  // end of namespace petibm
