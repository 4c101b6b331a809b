/* TEST INFRASTRUCTURE (checker, never linked into the product).
 *
 * The reference's own path on a box that has a real PETSc: exactly the calls LinSolverKSP makes
 * (src/linsolver/linsolverksp.cpp:62-66 KSPCreate/SetType(KSPCG)/SetReusePreconditioner/SetFromOptions,
 * :78-79 KSPReset/KSPSetOperators, :92 KSPSolve, :94 KSPGetConvergedReason, :116 KSPGetIterationNumber,
 * :128 KSPGetResidualNorm) plus the constant null space the applications attach (navierstokes.cpp:404-413),
 * on a matrix and right-hand side written by bench.py / the tests (oracle.Csr arrays, raw little-endian).
 *
 * NOT BUILT OR RUN in the image this repository was developed in (no PETSc, no MPI; bench.py's probe_petsc()
 * reports what it finds).  Build: oracle/build_petsc_driver.sh -> oracle/_ref/petsc_ksp_driver.
 *
 * input file:  int64 n, int64 nnz, int32 rowptr[n+1], int32 col[nnz], double val[nnz], double b[n]
 * output file: int32 reason, int32 its, double rnorm, double seconds, int32 nhist, double hist[nhist], double x[n]
 * usage: petsc_ksp_driver in.bin out.bin [-poisson_ksp_type cg -poisson_pc_type none -poisson_ksp_rtol 0 ...]
 */
#include <petscksp.h>
#include <stdio.h>
#include <stdlib.h>

int main(int argc, char **argv)
{
    PetscErrorCode ierr;
    Mat A; Vec x, b; KSP ksp; MatNullSpace nsp;
    ierr = PetscInitialize(&argc, &argv, NULL, NULL); if (ierr) return ierr;
    if (argc < 3) { PetscPrintf(PETSC_COMM_WORLD, "usage: %s in.bin out.bin [options]\n", argv[0]); return 1; }
    FILE *f = fopen(argv[1], "rb");
    if (!f) SETERRQ(PETSC_COMM_SELF, PETSC_ERR_FILE_OPEN, "cannot open input");
    long long n, nnz;
    if (fread(&n, 8, 1, f) != 1 || fread(&nnz, 8, 1, f) != 1) return 2;
    int *rp = malloc((n + 1) * sizeof(int)), *col = malloc(nnz * sizeof(int));
    double *val = malloc(nnz * sizeof(double)), *bb = malloc(n * sizeof(double));
    if (fread(rp, 4, n + 1, f) != (size_t)(n + 1) || fread(col, 4, nnz, f) != (size_t)nnz ||
        fread(val, 8, nnz, f) != (size_t)nnz || fread(bb, 8, n, f) != (size_t)n) return 2;
    fclose(f);

    ierr = MatCreateSeqAIJ(PETSC_COMM_SELF, n, n, 7, NULL, &A); CHKERRQ(ierr);
    for (PetscInt i = 0; i < n; ++i) {
        PetscInt nc = rp[i + 1] - rp[i];
        PetscInt *cc; PetscMalloc1(nc, &cc);
        for (PetscInt k = 0; k < nc; ++k) cc[k] = col[rp[i] + k];
        ierr = MatSetValues(A, 1, &i, nc, cc, val + rp[i], INSERT_VALUES); CHKERRQ(ierr);
        PetscFree(cc);
    }
    ierr = MatAssemblyBegin(A, MAT_FINAL_ASSEMBLY); CHKERRQ(ierr);
    ierr = MatAssemblyEnd(A, MAT_FINAL_ASSEMBLY); CHKERRQ(ierr);
    /* navierstokes.cpp:404-413 */
    ierr = MatNullSpaceCreate(PETSC_COMM_SELF, PETSC_TRUE, 0, NULL, &nsp); CHKERRQ(ierr);
    ierr = MatSetNullSpace(A, nsp); CHKERRQ(ierr);
    ierr = MatNullSpaceDestroy(&nsp); CHKERRQ(ierr);
    ierr = MatCreateVecs(A, &x, &b); CHKERRQ(ierr);
    { PetscScalar *p; VecGetArray(b, &p); for (PetscInt i = 0; i < n; ++i) p[i] = bb[i]; VecRestoreArray(b, &p); }

    /* linsolverksp.cpp:62-66 */
    ierr = KSPCreate(PETSC_COMM_SELF, &ksp); CHKERRQ(ierr);
    ierr = KSPSetOptionsPrefix(ksp, "poisson_"); CHKERRQ(ierr);
    ierr = KSPSetType(ksp, KSPCG); CHKERRQ(ierr);
    ierr = KSPSetReusePreconditioner(ksp, PETSC_TRUE); CHKERRQ(ierr);
    ierr = KSPSetFromOptions(ksp); CHKERRQ(ierr);
    /* :78-79 */
    ierr = KSPReset(ksp); CHKERRQ(ierr);
    ierr = KSPSetOperators(ksp, A, A); CHKERRQ(ierr);
    PetscInt maxit; KSPGetTolerances(ksp, NULL, NULL, NULL, &maxit);
    PetscReal *hist; PetscMalloc1(maxit + 2, &hist);
    ierr = KSPSetResidualHistory(ksp, hist, maxit + 2, PETSC_TRUE); CHKERRQ(ierr);
    /* :92 (timed; a first solve warms the PC set-up like the first time step does) */
    ierr = KSPSolve(ksp, b, x); CHKERRQ(ierr);
    double t0 = MPI_Wtime();
    ierr = KSPSolve(ksp, b, x); CHKERRQ(ierr);
    double secs = MPI_Wtime() - t0;
    KSPConvergedReason reason; PetscInt its, nh; PetscReal rnorm; const PetscReal *h;
    KSPGetConvergedReason(ksp, &reason); KSPGetIterationNumber(ksp, &its); KSPGetResidualNorm(ksp, &rnorm);
    KSPGetResidualHistory(ksp, &h, &nh);

    f = fopen(argv[2], "wb");
    int r32 = (int)reason, i32 = (int)its, n32 = (int)nh;
    fwrite(&r32, 4, 1, f); fwrite(&i32, 4, 1, f); fwrite(&rnorm, 8, 1, f); fwrite(&secs, 8, 1, f); fwrite(&n32, 4, 1, f);
    fwrite(h, 8, nh, f);
    { const PetscScalar *p; VecGetArrayRead(x, &p); fwrite(p, 8, n, f); VecRestoreArrayRead(x, &p); }
    fclose(f);
    PetscPrintf(PETSC_COMM_SELF, "reason %d its %d rnorm %.17g seconds %.6f\n", r32, i32, (double)rnorm, secs);
    KSPDestroy(&ksp); VecDestroy(&x); VecDestroy(&b); MatDestroy(&A);
    free(rp); free(col); free(val); free(bb);
    return PetscFinalize();
}
