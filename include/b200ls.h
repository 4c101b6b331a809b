/*
 * b200ls.h -- C ABI of libb200ls.so, the B200-native (sm_100a) linear-solver core
 * that sits behind petibm::linsolver::LinSolverBase as a third backend next to
 * LinSolverKSP / LinSolverAmgX.
 *
 * Plain C: opaque handle, plain pointers and sizes, caller-owned buffers, no
 * exceptions, no globals.  Every function returns B200LS_OK (0) or a negative
 * B200LS_ERR_* code; b200ls_last_error(h) gives the text.  There is no CPU
 * fallback anywhere: without a CUDA device every compute entry point fails with
 * B200LS_ERR_CUDA.
 *
 * Reference interface each entry point replaces (paths into barbagroup/PetIBM):
 *
 *   b200ls_create / b200ls_set_options*   LinSolverKSP::LinSolverKSP + ::init
 *                                         src/linsolver/linsolverksp.cpp:16-20,48-69
 *                                         (KSPCreate, KSPSetOptionsPrefix, KSPSetType(KSPCG),
 *                                          PetscOptionsInsertFile, KSPSetFromOptions)
 *   b200ls_set_poisson_stencil            LinSolverKSP::setMatrix            linsolverksp.cpp:72-82
 *   b200ls_set_csr                        (KSPSetOperators on the MPIAIJ Mat assembled by
 *                                          createDivergence/createGradient/createBnHead +
 *                                          MatMatMult, navierstokes.cpp:347-356)
 *   b200ls_set_staggered                  the same call for the velocity system A = I/dt - c nu L
 *                                         (navierstokes.cpp:342-344, createlaplacian.cpp:134-159) and IBPM's
 *                                         modified Poisson system (ibpm.cpp:100-203)
 *   b200ls_set_poisson_hybrid             the same call for IBPM's [D;E] BN [G,-H] on a stretched grid
 *                                         applications/ibpm/ibpm.cpp:164-194 (pressure block = DBNG of the mesh)
 *   options ksp_type preonly + pc_type lu KSPPREONLY + PCLU of the decoupled-IBPM forces system
 *                                         examples/decoupledibpm/(case)/config/forces_solver.info,
 *                                         applications/decoupledibpm/decoupledibpm.cpp:271-285
 *   option pc_type mg                     stands in for -poisson_pc_type gamg / hypre / AmgX AMG of the shipped
 *                                         examples/(app)/(case)/config/poisson_solver.info (extension: PETSc PCMG option names)
 *   b200ls_set_nullspace                  MatSetNullSpace on DBNG            navierstokes.cpp:404-413
 *   b200ls_solve                          LinSolverKSP::solve -> KSPSolve    linsolverksp.cpp:85-105
 *   b200ls_get_iters / _get_residual      LinSolverKSP::getIters/getResidual linsolverksp.cpp:110-132
 *   b200ls_get_reason                     KSPGetConvergedReason              linsolverksp.cpp:94
 *   b200ls_get_history                    KSPGetResidualHistory (parity tests)
 *   b200ls_destroy                        LinSolverKSP::destroy / dtor       linsolverksp.cpp:23-45
 *   b200ls_axis_from_subdomains           parser::parseSubDomains/parseOneSubDomain
 *                                         src/parser/parser.cpp:297-356, misc::stretchGrid
 *                                         include/petibm/misc.h:148-163
 *   b200ls_divergence / _gradient / _project   MatMult(D, ...), MatMult(G / BNG, ...), VecAXPY around the solve
 *                                         navierstokes.cpp:442,540-551,583-615; createdivergence.cpp:140-223,
 *                                         creategradient.cpp:70-128
 *   b200ls_convection                         MatMult(N, ...) of the convection MatShell, createconvection.cpp:39-332
 *   b200ls_repart_*                       the DMDA ownership the Vecs/Mat rows arrive in
 *                                         src/mesh/cartesianmesh.cpp:500-538,709-721 (DMDACreate3d,
 *                                         AOApplicationToPetsc) -> slab partition of the solver
 *   b200ls_comm_*                         PETSC_COMM_WORLD (MPI) inside KSP: VecScatter halos
 *                                         of MatMult_MPIAIJ and MPI_Allreduce of VecDot/VecNorm
 */
#ifndef B200LS_H
#define B200LS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200LS_VERSION 100 /* 0.1.0 */

/* ---- error codes ---- */
#define B200LS_OK 0
#define B200LS_ERR_ARG (-1)         /* bad argument / call order */
#define B200LS_ERR_CUDA (-2)        /* CUDA runtime error or no device */
#define B200LS_ERR_UNSUPPORTED (-3) /* option / operator outside the supported set */
#define B200LS_ERR_NCCL (-4)
#define B200LS_ERR_DIVERGED (-5)    /* KSP reason < 0 (what LinSolverKSP::solve turns into SETERRQ) */
#define B200LS_ERR_MISMATCH (-6)    /* operator verification against a CSR matrix failed */
#define B200LS_ERR_PARSE (-7)       /* options text could not be parsed */

/* ---- enums (values follow PETSc's option strings) ---- */
#define B200LS_KSP_CG 0
#define B200LS_KSP_BCGS 1
#define B200LS_KSP_PREONLY 2 /* with B200LS_PC_LU: direct solve of a small assembled system (decoupled IBPM forces system) */
#define B200LS_PC_NONE 0
#define B200LS_PC_JACOBI 1
#define B200LS_PC_LU 3 /* dense LU with partial pivoting of a matrix with at most 16384 rows; only with B200LS_KSP_PREONLY */
#define B200LS_PC_MG 2 /* geometric multigrid on the separable pressure operator (extension, single GPU) */
#define B200LS_NORM_NONE 0
#define B200LS_NORM_PRECONDITIONED 1
#define B200LS_NORM_UNPRECONDITIONED 2
#define B200LS_NORM_NATURAL 3

/* KSPConvergedReason values reported by b200ls_get_reason (same numbers as PETSc) */
#define B200LS_CONVERGED_RTOL 2
#define B200LS_CONVERGED_ATOL 3
#define B200LS_CONVERGED_ITS 4
#define B200LS_DIVERGED_ITS (-3)
#define B200LS_DIVERGED_DTOL (-4)
#define B200LS_DIVERGED_BREAKDOWN (-5)
#define B200LS_DIVERGED_INDEFINITE_PC (-8)
#define B200LS_DIVERGED_NANORINF (-9)
#define B200LS_DIVERGED_INDEFINITE_MAT (-10)
#define B200LS_DIVERGED_PC_FAILED (-11)

/* halo / reduction transports for the multi-GPU path */
#define B200LS_REDUCE_P2P 0  /* in-kernel all-reduce through peer-mapped mailboxes (NVLink) */
#define B200LS_REDUCE_NCCL 1 /* ncclAllReduce of the scalar partial sums */
#define B200LS_HALO_STORE 0  /* boundary planes stored to the peer's ghost plane by the update kernel */
#define B200LS_HALO_MEMCPY 1 /* cudaMemcpyPeerAsync of the boundary planes */

typedef struct b200ls_solver b200ls_solver;

typedef struct b200ls_options
{
    int ksp_type;   /* default B200LS_KSP_CG (linsolverksp.cpp:64) */
    int pc_type;    /* default B200LS_PC_NONE */
    int norm_type;  /* default B200LS_NORM_PRECONDITIONED */
    int max_it;     /* default 10000 */
    double rtol;    /* default 1e-5 */
    double atol;    /* default 1e-50 */
    double divtol;  /* default 1e4 */
    int check_every; /* host polls the device convergence flag every this many iterations (default 32) */
    int variant;     /* kernel variant selector for experiments; 0 = default */
    /* pc_type mg (PETSc's PCMG option names): -<name>_pc_mg_levels (0 = as many as the grid allows),
     * -<name>_mg_levels_ksp_max_it (Chebyshev/Jacobi smoothing steps before and after the coarse correction, default 2),
     * -<name>_mg_coarse_ksp_max_it (Chebyshev steps on the coarsest level, default 16) */
    int mg_levels;
    int mg_smooth_its;
    int mg_coarse_its;
} b200ls_options;

/* ---- library-level ---- */
int b200ls_version(void);
const char *b200ls_error_string(int code);
int b200ls_device_count(int *n);

/* Grid helper: cell widths of one axis from PetIBM's YAML sub-domain list. dL_out holds
 * sum(cells) doubles on return; *n_out receives that count. cap is the capacity of dL_out. */
int b200ls_axis_from_subdomains(double start, int nsub, const double *ends, const int *cells,
                                const double *ratios, double *dL_out, int cap, int *n_out);

/* Parse PETSc options-file text ("-poisson_ksp_type cg\n-poisson_ksp_atol 1e-6 ...") for the
 * given prefix ("poisson_") into opts (which must hold defaults on entry, see
 * b200ls_default_options). Any option under the prefix that this backend does not implement
 * (gamg, hypre, lu, ...) is an error: B200LS_ERR_UNSUPPORTED, message in errbuf. */
void b200ls_default_options(b200ls_options *opts);
int b200ls_parse_options(const char *text, const char *prefix, b200ls_options *opts, char *errbuf,
                         size_t errlen);

/* ---- solver object ---- */
int b200ls_create(b200ls_solver **out, int device);
int b200ls_destroy(b200ls_solver *h);
const char *b200ls_last_error(const b200ls_solver *h);
int b200ls_set_options(b200ls_solver *h, const b200ls_options *opts);
int b200ls_get_options(const b200ls_solver *h, b200ls_options *opts);
/* Launch-configuration knobs for experiments: "kz_chunk" (z planes per CTA of the fused SpMV
 * kernel, 0 = auto), "upd_blocks" (grid of the update kernel, 0 = auto), "tile" (fused SpMV variant: -1 = auto -- the TMA
 * kernel on non-periodic grids without Jacobi, the cp.async kernel otherwise; 10..18 cp.async tiles, 40..49 TMA variants,
 * 50..56 staging-free variants), "upd_variant", "upd_reverse", "use_graph" / "use_pdl" (CUDA-graph batches and programmatic
 * dependent launch of the CG loop, default 1), "csr_graph" (the same for the assembled-operator paths, default 0),
 * "mg_graph", "mg_tail" (coarse levels of the multigrid cycle as one launch), "mg_fuse" (residual update and reduction of
 * the preconditioned CG loop folded into the first / last fine-level step of the cycle): default 1 since round 2;
 * "sep_tile" (line-coefficient / hybrid operator: 2 = tiled plane-marching kernels, 0 = row-per-thread kernels, -1 = auto:
 * tiled for the BiCGStab solve of non-periodic 3-D systems of >= 2^20 rows, the timed case), "sep_zchunk"
 * (planes per z chunk of the tiled kernels, 0 = as many chunks as fit one resident wave), "sep_stages" (planes in their
 * per-thread cp.async queue: 3 or 4).
 * On several GPUs the keys must be set to the same values on every rank. */
int b200ls_set_tuning(b200ls_solver *h, const char *key, int value);

/* ---- multi-GPU communicator (one process per GPU; z-slab partition of the DMDA grid) ----
 * Call order on every rank:  b200ls_comm_init -> b200ls_set_poisson_stencil (allocates the
 * exchange arena) -> b200ls_comm_export (64-byte CUDA IPC handle of the arena) -> all-gather the
 * handles with the host transport (MPI_Allgather in PetIBM, torch.distributed in the harness)
 * -> b200ls_comm_connect(all handles, rank-major).  For B200LS_REDUCE_NCCL additionally:
 * rank 0 b200ls_nccl_unique_id -> broadcast 128 bytes -> every rank b200ls_nccl_init. */
int b200ls_comm_init(b200ls_solver *h, int rank, int nranks, int reduce_mode, int halo_mode);
int b200ls_comm_export(b200ls_solver *h, void *handle64);
int b200ls_comm_connect(b200ls_solver *h, const void *handles, int nranks);
/* Collective: closes this rank's mappings of the peers' arenas. Call on every rank and barrier on the host
 * transport BEFORE any rank replaces its operator (a second setMatrix), so that no arena is freed while a
 * peer still maps it. No-op when not connected. */
int b200ls_comm_disconnect(b200ls_solver *h);
int b200ls_nccl_unique_id(void *id128);
int b200ls_nccl_init(b200ls_solver *h, const void *id128);

/* ---- DMDA box partition <-> slab partition (host only, no device needed) ----
 * Inside PetIBM the pressure grid is distributed the way DMDACreate{2,3}d(PETSC_DECIDE) chose
 * (src/mesh/cartesianmesh.cpp:500-538): an m x n x p process grid, rank = px + m*(py + n*pz), one box per rank
 * numbered contiguously i-fastest (cartesianmesh.cpp:709-721), first (M mod m) ranks one cell longer per axis.
 * The device solver works on slabs along the slowest axis.  A b200ls_repart plans the exchange between the two
 * (the VecScatter PETSc would do); the caller moves the bytes with ONE all-to-all per direction on its host
 * transport (MPI_Alltoallv in the PetIBM shim).  box -> slab: send box_counts[s] doubles from offset box_displs[s]
 * of the local box-ordered vector to rank s (no packing: whole z-planes of the box), receive slab_counts[q] from
 * rank q at slab_displs[q] of an exchange buffer, then b200ls_repart_unpack_slab.  slab -> box: b200ls_repart_pack_slab,
 * send slab_counts/displs, receive straight into the box-ordered vector at box_counts/displs. */
typedef struct b200ls_repart b200ls_repart;
int b200ls_dmda_split(int64_t M, int m, int64_t *starts /* m + 1 */);
/* n: global cells per axis (2-D: n[2] ignored); procs: (m, n, p) of the DMDA (2-D: procs[2] ignored). */
int b200ls_repart_create(b200ls_repart **out, int dim, const int64_t n[3], const int procs[3], int rank);
int b200ls_repart_destroy(b200ls_repart *p);
/* box_lo/box_hi are reported on three axes with a 2-D grid normalised to (nx, 1, ny); identity != 0 when the box
 * partition already is the slab partition (1 x 1 x P process grid: no exchange needed). Any pointer may be null. */
int b200ls_repart_info(const b200ls_repart *p, int64_t box_lo[3], int64_t box_hi[3], int64_t *nbox, int64_t *slab_lo,
                       int64_t *slab_hi, int64_t *nslab, int *identity);
int b200ls_repart_counts(const b200ls_repart *p, int64_t *box_counts, int64_t *box_displs, int64_t *slab_counts,
                         int64_t *slab_displs);
int b200ls_repart_unpack_slab(const b200ls_repart *p, const double *recvbuf, double *slab);
int b200ls_repart_pack_slab(const b200ls_repart *p, const double *slab, double *sendbuf);
/* PETSc global indices (columns of the assembled Mat) -> natural indices; natural indices of the own rows. */
int b200ls_repart_petsc_to_natural(const b200ls_repart *p, int64_t count, const int32_t *petsc_idx, int32_t *natural_idx);
int b200ls_repart_box_rows(const b200ls_repart *p, int64_t *natural_rows);
/* All process grids of nranks ranks whose boxes have exactly the local sizes the ranks report (sizes[q] = rows
 * owned by rank q, e.g. from MatGetOwnershipRanges).  The Mat handed to setMatrix does not carry its DMDA, so the
 * shim tries these candidates against the matrix entries (b200ls_verify_csr_rows) and keeps the one that matches.
 * Writes min(cap, *found) triples to procs_out. */
int b200ls_repart_candidates(int dim, const int64_t n[3], int nranks, const int64_t *sizes, int *procs_out, int cap,
                             int *found);

/* ---- operator ----
 * Matrix-free separable pressure-Poisson operator DBNG = D (dt I) G of a stretched Cartesian
 * staggered grid:  dx,dy,dz are the GLOBAL pressure-cell widths
 * (dz ignored when dim == 2), periodic[d] the periodicity flags, [slab_lo, slab_hi) the planes this
 * rank owns along the slowest axis (z in 3-D, y in 2-D; ignored for a single GPU in 2-D, where the
 * whole grid is owned).  Vectors passed to b200ls_solve are the rank-local part in
 * PETSc DMDA ordering (i fastest), length nx*ny*(zhi-zlo) (3-D) or nx*(yhi-ylo) (2-D). */
int b200ls_set_poisson_stencil(b200ls_solver *h, int dim, const int64_t n[3], const int periodic[3],
                               const double *dx, const double *dy, const double *dz, double dt,
                               int64_t slab_lo, int64_t slab_hi);

/* Verify the matrix-free operator bit-for-bit against rows of an assembled CSR matrix (what the
 * PetIBM shim extracts with MatGetRow from the Mat given to setMatrix).  Rows are the rank-local
 * rows, columns are global natural indices.  Returns B200LS_ERR_MISMATCH on any difference. */
int b200ls_verify_csr(b200ls_solver *h, int64_t nrows, const int64_t *rowptr, const int32_t *col,
                      const double *val, double *max_abs_diff);

/* The same for rows given in ANY order (rows of a DMDA box, see b200ls_repart_*): natural_rows[r] is the natural
 * index i + nx*(j + ny*k) of local row r, columns are natural indices.  Off-diagonal entries must match bitwise.
 * diag_ulps > 0 lets the diagonal differ by that many units in the last place: on several ranks PETSc's
 * MatMatMult (navierstokes.cpp:351-356) accumulates the off-process terms of a diagonal entry after the local
 * ones, so the rounding of rows on a partition boundary depends on the partition -- PETSc's own result is not
 * bitwise reproducible across process counts there.  0 = bitwise (what b200ls_verify_csr does). */
int b200ls_verify_csr_rows(b200ls_solver *h, int64_t nrows, const int64_t *natural_rows, const int64_t *rowptr,
                           const int32_t *col, const double *val, int diag_ulps, double *max_abs_diff);

/* General assembled operator (any square CSR, single GPU): used for operators the separable
 * stencil cannot express (IBPM modified Poisson, BN order > 1, the velocity system A). */
int b200ls_set_csr(b200ls_solver *h, int64_t nrows, const int64_t *rowptr, const int32_t *col,
                   const double *val);

/* Line-coefficient form of an assembled staggered-grid operator (single GPU).  For the two matrices PetIBM hands
 * to setMatrix that are stencils with one-dimensional coefficients but not the separable pressure operator:
 * the velocity system A = I/dt - c nu L on [u | v | w] (navierstokes.cpp:342-344, createlaplacian.cpp:134-159,
 * solved with bcgs + jacobi) and IBPM's modified Poisson system [D;E] BN [G,-H] (ibpm.cpp:100-203: pressure block +
 * Lagrangian coupling).  nfields (1..3) stencil blocks stored one after the other, dims[3*f + d] points of field f
 * along axis d (1 for the missing axis in 2-D), periodic[d] per axis; rows behind the blocks and columns >= the
 * blocks' total size form a CSR remainder.  The structure (two 1-D arrays per field and axis, the diagonal as a
 * vector, the remainder) is READ OUT OF THE MATRIX and every entry is checked bitwise against it; if anything does
 * not fit, B200LS_ERR_MISMATCH is returned, the solver is left untouched and the caller falls back to
 * b200ls_set_csr.  The device operator is bit-identical to MatMult on the assembled matrix (terms added in
 * ascending column order, no FMA) at 24 B/row instead of ~100. */
int b200ls_set_staggered(b200ls_solver *h, int nfields, const int64_t *dims, const int *periodic, int64_t nrows,
                         const int64_t *rowptr, const int32_t *col, const double *val);
/* The analysis itself (host only, no device needed).  coef: b200ls_staggered_coef_size(nfields, dims) doubles, per
 * field and axis cm[n] then cp[n]; diag: one double per stencil row; rem_rowptr: nrows + 1; rem_col / rem_val:
 * capacity nnz (may be null to only count).  errbuf receives the first mismatch as text. */
int64_t b200ls_staggered_coef_size(int nfields, const int64_t *dims);
int b200ls_staggered_analyze(int nfields, const int64_t *dims, const int *periodic, int64_t nrows, const int64_t *rowptr,
                             const int32_t *col, const double *val, double *coef, double *diag, int64_t *rem_rowptr,
                             int32_t *rem_col, double *rem_val, char *errbuf, size_t errlen);

/* IBPM's modified Poisson system on a STRETCHED grid: the pressure block is D (dt I) G of the mesh -- its coefficients
 * carry face areas, so they are not one-dimensional -- followed by the Lagrangian coupling rows/columns.  The block is
 * checked bitwise against the closed form of b200ls_set_poisson_stencil ((w_a w_b) * (dt * (1/h)), appendix A.1 of
 * SURVEY.md), the diagonal is taken from the matrix, everything else becomes the CSR remainder; same kernels and
 * same bit-identical row sums as b200ls_set_staggered.  With -<name>_pc_type mg the pressure block is preconditioned by
 * the geometric multigrid and the remaining rows by their diagonal.  B200LS_ERR_MISMATCH: the solver is untouched. */
int b200ls_set_poisson_hybrid(b200ls_solver *h, int dim, const int64_t *n, const int *periodic, const double *dx,
                              const double *dy, const double *dz, double dt, int64_t nrows, const int64_t *rowptr,
                              const int32_t *col, const double *val);
/* g: (nx+1) + (ny+1) + (nz+1) doubles (face arrays), diag: nx*ny*nz; the rest as in b200ls_staggered_analyze. */
int b200ls_hybrid_analyze(int dim, const int64_t *n, const int *periodic, const double *dx, const double *dy, const double *dz,
                          double dt, int64_t nrows, const int64_t *rowptr, const int32_t *col, const double *val, double *g,
                          double *diag, int64_t *rem_rowptr, int32_t *rem_col, double *rem_val, char *errbuf, size_t errlen);

/* Null space attached to the operator: has_const != 0 -> the constant vector
 * (MatNullSpaceCreate(comm, PETSC_TRUE, 0, ...), navierstokes.cpp:404-413); nvecs explicit
 * orthonormal vectors of local length nrows (ibpm.cpp:251-267). */
int b200ls_set_nullspace(b200ls_solver *h, int has_const, int nvecs, const double *vecs);

/* y = A x on the device operator, host buffers (tests / verification). */
int b200ls_apply(b200ls_solver *h, const double *x_host, double *y_host);

/* ---- solve ----
 * KSPSolve semantics with zero initial guess: x is overwritten. Host buffers; the H2D copy of b
 * and D2H copy of x happen inside. Returns B200LS_OK if reason > 0, B200LS_ERR_DIVERGED if
 * reason < 0 (x is still written), other codes on failure. */
int b200ls_solve(b200ls_solver *h, const double *b_host, double *x_host);
/* Same with device-resident buffers (length as above, compact i-fastest). */
int b200ls_solve_device(b200ls_solver *h, const double *b_dev, double *x_dev);

int b200ls_get_iters(const b200ls_solver *h, int *its);
int b200ls_get_residual(const b200ls_solver *h, double *rnorm);
int b200ls_get_reason(const b200ls_solver *h, int *reason);
/* Copies min(cap, n) entries of the residual-norm history (entry 0 = initial) and sets *n. */
int b200ls_get_history(const b200ls_solver *h, double *buf, int cap, int *n);

/* ---- the operators on either side of the pressure solve, matrix-free (single GPU) ----
 * With these, b and x of b200ls_solve_device never leave the device: rhs2 = D u* (MatMult(D, UGlobal, rhs2),
 * navierstokes.cpp:540-551) -> solve -> u -= (BN G) dP, p += dP (navierstokes.cpp:583-615) is a chain of launches on the
 * solver's stream.  D (createdivergence.cpp:140-223) and G (creategradient.cpp:70-128) of the mesh given to
 * b200ls_set_poisson_stencil; velocities are PetIBM's packed vector [u | v | w] (cartesianmesh.cpp:251-273: one point
 * fewer than cells along the field's own direction unless periodic), pressures the cell vector of the solve.  Bit-identical
 * to MatMult on the assembled matrices (ascending column order, no FMA).  The boundary terms (DCorrection, bc1) are the
 * application's MatShells and stay with the caller.  *_device: device pointers, asynchronous on b200ls_stream(h). */
int b200ls_velocity_size(b200ls_solver *h, int64_t *nvel, int64_t *npressure);
int b200ls_divergence_device(b200ls_solver *h, const double *u_dev, double *out_dev);               /* out = D u */
int b200ls_gradient_device(b200ls_solver *h, const double *p_dev, double *out_dev, int with_bn);   /* out = G p or (BN G) p */
int b200ls_project_device(b200ls_solver *h, double *u_dev, double *p_dev, const double *dp_dev);   /* u -= BNG dp; p += dp (p may be null) */
/* Convection term N(q) (createconvection.cpp:39-332, the MatShell H of navierstokes.cpp:446-470) on the GHOSTED local
 * arrays of u, v, w: one ghost layer on every side, i fastest -- what DMCompositeScatterArray +
 * Boundary::copyValues2LocalVecs produce in PetIBM (createconvection.cpp:217-221).  Output: packed [u|v|w].
 * b200ls_ghosted_sizes gives the lengths of the three arrays; b200ls_ghosted_from_packed_device fills their interior
 * and the wrap layers of periodic axes from a packed vector (DMGlobalToLocal); the ghost layers of non-periodic axes
 * are the boundary conditions' values and stay with the caller.  qw may be null in 2-D. */
int b200ls_ghosted_sizes(b200ls_solver *h, int64_t sizes[3]);
int b200ls_convection_device(b200ls_solver *h, const double *qu_dev, const double *qv_dev, const double *qw_dev, double *out_dev);
int b200ls_ghosted_from_packed_device(b200ls_solver *h, const double *packed_dev, double *qu_dev, double *qv_dev, double *qw_dev);
int b200ls_convection(b200ls_solver *h, const double *qu_host, const double *qv_host, const double *qw_host, double *out_host);
int b200ls_divergence(b200ls_solver *h, const double *u_host, double *out_host);
int b200ls_gradient(b200ls_solver *h, const double *p_host, double *out_host, int with_bn);
int b200ls_project(b200ls_solver *h, double *u_host, double *p_host, const double *dp_host);

/* ---- measurement hooks (bench.py) ----
 * Device time of the last solve (solve_ms: scatter of b .. gather of x; loop_ms: the iteration loop
 * only) measured with CUDA events on the solver's own stream, number of kernels launched in it,
 * and per-kernel-class accumulated event time when
 * profiling was enabled with b200ls_set_profile(h, 1) (serialises kernels; never used for the
 * headline number). class 0 = fused SpMV kernel, 1 = fused update/reduction kernel. */
int b200ls_get_timing(const b200ls_solver *h, double *solve_ms, double *loop_ms, int64_t *launches);
/* Device time of the last b200ls_solve (host buffers) from before the H2D copy of b to after the D2H
 * copy of x, CUDA events on the solver stream (ms). */
int b200ls_get_e2e_ms(const b200ls_solver *h, double *e2e_ms);
int b200ls_set_profile(b200ls_solver *h, int enable);
int b200ls_get_profile(const b200ls_solver *h, int kclass, double *total_ms, int64_t *count);
/* Run `reps` launches of one kernel class on the current solver state and return the average
 * launch duration in ms (CUDA events on the solver stream); flush_l2 != 0 writes a >L2 buffer
 * between launches. kclass as above; 2 = plain SpMV (b200ls_apply path). */
int b200ls_time_kernel(b200ls_solver *h, int kclass, int reps, int flush_l2, double *avg_ms);
/* Device-side timeline of the reduction-carrying kernels (diagnostics): capacity entries of 5 uint64
 * {kernel start, local grid reduction done, cross-GPU all-reduce done, scalar logic done, kind} in
 * %globaltimer nanoseconds; b200ls_get_trace copies and clears. capacity 0 disables. */
int b200ls_set_trace(b200ls_solver *h, int capacity);
int b200ls_get_trace(b200ls_solver *h, unsigned long long *buf, int capacity, int *n);
void *b200ls_stream(b200ls_solver *h); /* cudaStream_t of the solver, for external event timing */

#ifdef __cplusplus
}
#endif
#endif /* B200LS_H */
