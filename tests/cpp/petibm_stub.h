// petibm_stub.h -- TEST SCAFFOLDING: the handful of PETSc / PetIBM declarations that
// petibm_b200/csrc/petibm_shim/linsolverb200.{h,cpp} touch, as single-process stand-ins, so that the shim
// can be compiled and driven here where PETSc, MPI and yaml-cpp are not installed.  Signatures follow
// PETSc 3.16's public headers and include/petibm/linsolver.h:59-147; nothing here is shipped.
#pragma once
#include <cstdio>
#include <string>
#include <vector>

typedef int PetscErrorCode;
typedef int PetscInt;
typedef int PetscMPIInt;
typedef double PetscReal;
typedef double PetscScalar;
typedef enum { PETSC_FALSE, PETSC_TRUE } PetscBool;
typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
#define PETSC_COMM_WORLD 0
#define MPI_INT 1
#define MPI_BYTE 2
#define MPI_DOUBLE 3
#define MPI_LONG_LONG 4
#define MPI_MIN 1
#define PETSC_ERR_SUP 56
#define PETSC_ERR_ARG_WRONG 62
#define PETSC_ERR_FILE_OPEN 65
#define PETSC_ERR_LIB 76
#define PETSC_ERR_CONV_FAILED 82

#define PetscFunctionBeginUser
#define PetscFunctionReturn(x) return (x)
#define CHKERRQ(ierr) do { if (ierr) return (ierr); } while (0)
#define SETERRQ(comm, code, msg) do { std::fprintf(stderr, "[stub PETSc error %d] %s\n", (int)(code), msg); return (code); } while (0)
#define SETERRQ1(comm, code, fmt, a) do { std::fprintf(stderr, "[stub PETSc error %d] ", (int)(code)); std::fprintf(stderr, fmt, a); std::fprintf(stderr, "\n"); return (code); } while (0)
#define SETERRQ2(comm, code, fmt, a, b) do { std::fprintf(stderr, "[stub PETSc error %d] ", (int)(code)); std::fprintf(stderr, fmt, a, b); std::fprintf(stderr, "\n"); return (code); } while (0)
#define SETERRQ3(comm, code, fmt, a, b, c) do { std::fprintf(stderr, "[stub PETSc error %d] ", (int)(code)); std::fprintf(stderr, fmt, a, b, c); std::fprintf(stderr, "\n"); return (code); } while (0)

inline PetscErrorCode PetscFinalized(PetscBool *f) { *f = PETSC_FALSE; return 0; }
inline int MPI_Comm_rank(MPI_Comm, int *r) { *r = 0; return 0; }
inline int MPI_Comm_size(MPI_Comm, int *s) { *s = 1; return 0; }
inline int MPI_Barrier(MPI_Comm) { return 0; }
inline int MPI_Allreduce(const void *s, void *r, int n, MPI_Datatype, MPI_Op, MPI_Comm) { for (int i = 0; i < n; ++i) ((int *)r)[i] = ((const int *)s)[i]; return 0; }
inline int MPI_Alltoallv(const void *s, const int *sc, const int *sd, MPI_Datatype, void *r, const int *, const int *rd, MPI_Datatype, MPI_Comm) { for (int i = 0; i < sc[0]; ++i) ((double *)r)[rd[0] + i] = ((const double *)s)[sd[0] + i]; return 0; }
inline int MPI_Allgather(const void *s, int n, MPI_Datatype, void *r, int, MPI_Datatype, MPI_Comm) { for (int i = 0; i < n; ++i) ((char *)r)[i] = ((const char *)s)[i]; return 0; }

inline int stub_mpi_size(MPI_Datatype t) { return t == MPI_BYTE ? 1 : t == MPI_INT ? 4 : 8; }
inline int MPI_Allgatherv(const void *s, int n, MPI_Datatype t, void *r, const int *, const int *rd, MPI_Datatype, MPI_Comm) { const int b = stub_mpi_size(t); for (long i = 0; i < (long)n * b; ++i) ((char *)r)[(long)rd[0] * b + i] = ((const char *)s)[i]; return 0; }

struct _p_Vec { std::vector<double> a; };
typedef _p_Vec *Vec;
inline PetscErrorCode VecGetArray(Vec v, PetscScalar **a) { *a = v->a.data(); return 0; }
inline PetscErrorCode VecRestoreArray(Vec, PetscScalar **) { return 0; }
inline PetscErrorCode VecGetArrayRead(Vec v, const PetscScalar **a) { *a = v->a.data(); return 0; }
inline PetscErrorCode VecRestoreArrayRead(Vec, const PetscScalar **) { return 0; }

struct _p_MatNullSpace { PetscBool has_const = PETSC_FALSE; std::vector<Vec> vecs; };
typedef _p_MatNullSpace *MatNullSpace;
struct _p_Mat
{
    PetscInt n = 0;
    std::vector<PetscInt> rowptr, col;
    std::vector<PetscScalar> val;
    MatNullSpace nsp = nullptr;
};
typedef _p_Mat *Mat;
inline PetscErrorCode MatGetOwnershipRange(Mat A, PetscInt *lo, PetscInt *hi) { *lo = 0; *hi = A->n; return 0; }
inline PetscErrorCode MatGetRow(Mat A, PetscInt r, PetscInt *nc, const PetscInt **cols, const PetscScalar **vals)
{
    *nc = A->rowptr[r + 1] - A->rowptr[r];
    *cols = A->col.data() + A->rowptr[r];
    *vals = A->val.data() + A->rowptr[r];
    return 0;
}
inline PetscErrorCode MatRestoreRow(Mat, PetscInt, PetscInt *, const PetscInt **, const PetscScalar **) { return 0; }
inline PetscErrorCode MatGetNullSpace(Mat A, MatNullSpace *n) { *n = A->nsp; return 0; }
inline PetscErrorCode MatNullSpaceGetVecs(MatNullSpace n, PetscBool *hc, PetscInt *nv, const Vec **vecs)
{
    *hc = n->has_const;
    *nv = (PetscInt)n->vecs.size();
    *vecs = n->vecs.data();
    return 0;
}

namespace petibm
{
namespace linsolver
{
// the abstract interface of include/petibm/linsolver.h:59-147 (declarations only)
class LinSolverBase
{
public:
    LinSolverBase() = default;
    LinSolverBase(const std::string &solverName, const std::string &file) : name(solverName), config(file) {}
    virtual ~LinSolverBase() = default;
    virtual PetscErrorCode destroy() { name = config = type = ""; return 0; }
    PetscErrorCode getType(std::string &_type) const { _type = type; return 0; }
    virtual PetscErrorCode setMatrix(const Mat &A) = 0;
    virtual PetscErrorCode solve(Vec &x, Vec &b) = 0;
    virtual PetscErrorCode getIters(PetscInt &iters) = 0;
    virtual PetscErrorCode getResidual(PetscReal &res) = 0;

protected:
    std::string name, config, type;
    virtual PetscErrorCode init() = 0;
};
}  // namespace linsolver
}  // namespace petibm
