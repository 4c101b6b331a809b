/*
 * petibm_oracle.c -- CPU restatement of PetIBM's pressure-Poisson hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (petibm_b200/, libb200ls.so)
 * may import, link or call this file.  It is used by tests/, by
 * __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs
 * as the checker and as the timed CPU baseline -- never as the thing shipped.
 *
 * PARITY STATUS: the grid / operator part is PINNED against the reference's own
 * golden vectors (tests/mesh/cartesianmesh2d_dirichlet.cpp:171-284,
 * tests/mesh/cartesianmesh2d_yperiodic.cpp:180-290) by tests/test_oracle_mesh.py, and
 * createBnHead (any order) against the reference's row-sum known-answer test
 * (tests/operators/createbnhead_test.cpp:17-61) by tests/test_oracle_operator.py.
 * The Krylov part (KSPSolve_CG / KSPSolve_BCGS / KSPConvergedDefault /
 * MatNullSpaceRemove / PCApply_Jacobi) lives in PETSc 3.16 (pinned by
 * /root/reference/CMakeLists.txt:78-83, ">=3.16,<3.17"), which is NOT vendored
 * in /root/reference and not installed here; it is restated from PETSc's
 * published algorithm (src/ksp/ksp/impls/cg/cg.c, impls/bcgs/bcgs.c,
 * interface/iterativ.c, src/mat/interface/matnull.c) and anchored on the
 * reference's call sites (src/linsolver/linsolverksp.cpp:62-66,78-79,92-104,
 * applications/navierstokes/navierstokes.cpp:347-356,404-413).  The reference
 * holds no test or golden vector for that part: **parity unpinned** for KSP.
 *
 * Floating point: compile with -ffp-contract=off (see oracle/Makefile) so every
 * product and sum is a separately rounded IEEE-754 binary64 operation, in the
 * operand order of the cited reference lines.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------- */
/* 1. grid: parser.cpp:334-356, misc.h:148-163, cartesianmesh.cpp:136-355     */
/* ------------------------------------------------------------------------- */

/* misc.h:148-163  stretchGrid */
ORC_API void orc_stretch_grid(double bg, double ed, int n, double r, double *dL)
{
    dL[0] = (ed - bg) * (r - 1.0) / (pow(r, n) - 1.0);
    for (int i = 1; i < n; ++i) dL[i] = dL[i - 1] * r;
}

/* parser.cpp:334-356 parseOneSubDomain + :297-331 parseSubDomains.
 * Returns the total number of cells; dL must hold sum(cells) doubles. */
ORC_API int orc_axis_from_subdomains(double start, int nsub, const double *ends,
                                     const int *cells, const double *ratios,
                                     double *dL)
{
    int ntot = 0;
    double bg = start;
    for (int s = 0; s < nsub; ++s)
    {
        const int n = cells[s];
        const double ed = ends[s], r = ratios[s];
        if (fabs(r - 1.0) <= 1e-12)
            for (int i = 0; i < n; ++i) dL[ntot + i] = (ed - bg) / n;
        else
            orc_stretch_grid(bg, ed, n, r, dL + ntot);
        ntot += n;
        bg = ed;
    }
    return ntot;
}

/* cartesianmesh.cpp:136-176 createPressureMesh: coord of pressure points. */
ORC_API void orc_pressure_coord(int n, double min, const double *dLp, double *coord)
{
    double acc = 0.0;
    for (int i = 0; i < n; ++i)
    {
        /* std::partial_sum then a + min - 0.5*b */
        acc = (i == 0) ? dLp[0] : acc + dLp[i];
        coord[i] = acc + min - 0.5 * dLp[i];
    }
}

/* cartesianmesh.cpp:179-209 createVertexMesh: n+1 vertex coordinates. */
ORC_API void orc_vertex_coord(int n, double min, const double *dLp, double *coord)
{
    double acc = 0.0;
    coord[0] = 0.0;
    for (int i = 0; i < n; ++i)
    {
        acc = (i == 0) ? dLp[0] : acc + dLp[i];
        coord[i + 1] = acc;
    }
    for (int i = 0; i <= n; ++i) coord[i] += min;
}

/* cartesianmesh.cpp:212-355 createVelocityMesh, one (component, direction) pair.
 * same_dir != 0  <=>  dir == comp.  Writes the "True" arrays INCLUDING ghosts
 * (index 0 is the ghost at -1).  Returns the number of valid points n[comp][dir];
 * *len_out is the length of dLTrue/coordTrue written.
 * dLTrue/coordTrue need room for n+3 doubles. */
ORC_API int orc_velocity_axis(int np, double min, double max, const double *dLp,
                              int same_dir, int periodic, double *dLTrue,
                              double *coordTrue, int *len_out)
{
    int n, len;
    if (same_dir)
    {
        n = np - 1;
        len = n + 2; /* == np + 1 == number of vertices */
        /* coordTrue = vertex coordinates (:233-234) */
        orc_vertex_coord(np, min, dLp, coordTrue);
        /* std::adjacent_difference with f = 0.5*(x+y); first element copied (:243-247) */
        dLTrue[0] = dLp[0];
        for (int i = 1; i < np; ++i) dLTrue[i] = 0.5 * (dLp[i] + dLp[i - 1]);
        if (periodic)
        {
            n += 1;
            dLTrue[len - 1] = 0.5 * (dLp[0] + dLp[np - 1]); /* :257-258 */
            dLTrue[0] = dLTrue[len - 1];                    /* :262 */
            dLTrue[len] = dLTrue[1];                        /* :266 push_back */
            coordTrue[len] = max + dLp[0];                  /* :270-271 */
            len += 1;
        }
        else
            dLTrue[len - 1] = dLp[np - 1]; /* :278 */
    }
    else
    {
        n = np;
        len = n + 2;
        double *c = (double *)malloc(sizeof(double) * np);
        orc_pressure_coord(np, min, dLp, c);
        memcpy(coordTrue + 1, c, sizeof(double) * np);
        memcpy(dLTrue + 1, dLp, sizeof(double) * np);
        free(c);
        if (periodic)
        {
            coordTrue[0] = min - dLp[np - 1] / 2.0; /* :305-308 */
            coordTrue[len - 1] = max + dLp[0] / 2.0;
            dLTrue[0] = dLp[np - 1];
            dLTrue[len - 1] = dLp[0];
        }
        else
        {
            coordTrue[0] = min - dLp[0] / 2.0; /* :318-325 */
            coordTrue[len - 1] = max + dLp[np - 1] / 2.0;
            dLTrue[0] = dLp[0];
            dLTrue[len - 1] = dLp[np - 1];
        }
    }
    *len_out = len;
    return n;
}

/* ------------------------------------------------------------------------- */
/* 2. CSR container                                                           */
/* ------------------------------------------------------------------------- */

typedef struct
{
    int64_t nrows, ncols, nnz;
    int64_t *rowptr; /* nrows+1 */
    int32_t *col;    /* ascending inside a row, like an assembled PETSc AIJ */
    double *val;
} orc_csr;

static orc_csr *csr_alloc(int64_t nrows, int64_t ncols, int64_t nnz)
{
    orc_csr *m = (orc_csr *)calloc(1, sizeof(orc_csr));
    m->nrows = nrows;
    m->ncols = ncols;
    m->nnz = nnz;
    m->rowptr = (int64_t *)calloc((size_t)nrows + 1, sizeof(int64_t));
    m->col = (int32_t *)malloc(sizeof(int32_t) * (size_t)(nnz > 0 ? nnz : 1));
    m->val = (double *)malloc(sizeof(double) * (size_t)(nnz > 0 ? nnz : 1));
    return m;
}

ORC_API void orc_csr_free(orc_csr *m)
{
    if (!m) return;
    free(m->rowptr);
    free(m->col);
    free(m->val);
    free(m);
}
ORC_API int64_t orc_csr_nrows(const orc_csr *m) { return m->nrows; }
ORC_API int64_t orc_csr_ncols(const orc_csr *m) { return m->ncols; }
ORC_API int64_t orc_csr_nnz(const orc_csr *m) { return m->nnz; }
ORC_API void orc_csr_export(const orc_csr *m, int64_t *rowptr, int32_t *col, double *val)
{
    memcpy(rowptr, m->rowptr, sizeof(int64_t) * (size_t)(m->nrows + 1));
    memcpy(col, m->col, sizeof(int32_t) * (size_t)m->nnz);
    memcpy(val, m->val, sizeof(double) * (size_t)m->nnz);
}
ORC_API orc_csr *orc_csr_import(int64_t nrows, int64_t ncols, const int64_t *rowptr,
                                const int32_t *col, const double *val)
{
    orc_csr *m = csr_alloc(nrows, ncols, rowptr[nrows]);
    memcpy(m->rowptr, rowptr, sizeof(int64_t) * (size_t)(nrows + 1));
    memcpy(m->col, col, sizeof(int32_t) * (size_t)m->nnz);
    memcpy(m->val, val, sizeof(double) * (size_t)m->nnz);
    return m;
}

/* One row under construction: MatSetValue semantics (INSERT replaces, ADD adds,
 * negative column ignored), then sorted by column at "assembly". */
typedef struct
{
    int n;
    int32_t c[16];
    double v[16];
} rowbuf;

static void row_set(rowbuf *r, int64_t col, double v, int add)
{
    if (col < 0) return; /* PETSc ignores negative indices */
    for (int k = 0; k < r->n; ++k)
        if (r->c[k] == col)
        {
            r->v[k] = add ? r->v[k] + v : v;
            return;
        }
    r->c[r->n] = (int32_t)col;
    r->v[r->n] = v;
    r->n++;
}
static void row_sort(rowbuf *r)
{
    for (int a = 1; a < r->n; ++a)
    {
        int32_t c = r->c[a];
        double v = r->v[a];
        int b = a - 1;
        while (b >= 0 && r->c[b] > c)
        {
            r->c[b + 1] = r->c[b];
            r->v[b + 1] = r->v[b];
            --b;
        }
        r->c[b + 1] = c;
        r->v[b + 1] = v;
    }
}

/* ------------------------------------------------------------------------- */
/* 3. staggered-grid description (single rank: PETSc ordering == natural)     */
/* ------------------------------------------------------------------------- */

typedef struct
{
    int dim;
    int np[3];       /* pressure cells */
    int per[3];      /* periodic flags */
    int nv[3][3];    /* n[f][dir] for velocity fields */
    int64_t voff[4]; /* packed offsets of u, v, w; voff[dim] = UN */
    int64_t pN;
    const double *dLp[3]; /* pressure cell widths (dL[3][dir]) */
    double *hface[3];     /* dL[f][f][*], index i = face between cell i and i+1 */
} orc_grid;

static void grid_init(orc_grid *g, int dim, const int *np, const int *per,
                      const double *dx, const double *dy, const double *dz)
{
    static const double one = 1.0;
    memset(g, 0, sizeof(*g));
    g->dim = dim;
    for (int d = 0; d < 3; ++d)
    {
        g->np[d] = (d < dim) ? np[d] : 1;
        g->per[d] = (d < dim) ? per[d] : 0;
    }
    g->dLp[0] = dx;
    g->dLp[1] = dy;
    g->dLp[2] = (dim == 3) ? dz : &one; /* cartesianmesh.cpp:91-99 defaults */
    g->pN = (int64_t)g->np[0] * g->np[1] * g->np[2];
    int64_t off = 0;
    for (int f = 0; f < dim; ++f)
    {
        for (int d = 0; d < 3; ++d)
            g->nv[f][d] = (d == f) ? (g->np[d] - 1 + (g->per[d] ? 1 : 0)) : g->np[d];
        g->voff[f] = off;
        off += (int64_t)g->nv[f][0] * g->nv[f][1] * g->nv[f][2];
        /* dL[f][f][i], i in [0, nv): 0.5*(dLp[i] + dLp[i+1]) via adjacent_difference
         * f(x=cur, y=prev) = 0.5*(x+y)  (cartesianmesh.cpp:237-247); the periodic
         * wrap face is f(dLp[0], dLp.back()) (:257-258). */
        const int n = g->nv[f][f];
        g->hface[f] = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
        for (int i = 0; i < n; ++i)
        {
            if (i + 1 < g->np[f])
                g->hface[f][i] = 0.5 * (g->dLp[f][i + 1] + g->dLp[f][i]);
            else
                g->hface[f][i] = 0.5 * (g->dLp[f][0] + g->dLp[f][g->np[f] - 1]);
        }
    }
    g->voff[dim] = off;
}
static void grid_free(orc_grid *g)
{
    for (int f = 0; f < 3; ++f) free(g->hface[f]);
}

/* cartesianmesh.cpp:592-669 getNaturalIndex for field f (3 = pressure).
 * NOTE the reference tests periodic[0][dir] for every field (:608,:620,...). */
static int64_t nat_index(const orc_grid *g, int f, int i, int j, int k)
{
    const int *n = (f == 3) ? g->np : g->nv[f];
    if (i == -1) i = g->per[0] ? n[0] - 1 : -2;
    else if (i == n[0]) i = g->per[0] ? 0 : -2;
    if (i == -2) return -1;
    if (j == -1) j = g->per[1] ? n[1] - 1 : -2;
    else if (j == n[1]) j = g->per[1] ? 0 : -2;
    if (j == -2) return -1;
    if (k == -1) k = g->per[2] ? n[2] - 1 : -2;
    else if (k == n[2]) k = g->per[2] ? 0 : -2;
    if (k == -2) return -1;
    return (int64_t)i + (int64_t)n[0] * ((int64_t)j + (int64_t)n[1] * k);
}
static int64_t packed_index(const orc_grid *g, int f, int i, int j, int k)
{
    int64_t idx = nat_index(g, f, i, j, k);
    if (f == 3 || idx < 0) return idx;
    return g->voff[f] + idx;
}

/* ------------------------------------------------------------------------- */
/* 4. operators                                                               */
/* ------------------------------------------------------------------------- */

/* createdivergence.cpp:103-262 with normalize = PETSC_FALSE.
 * a0[f][side] is the ghost coefficient of the wall-normal velocity on the
 * minus/plus side of direction f (0 for Dirichlet / Convective,
 * singleboundarydirichlet.cpp:33-37; 1 for Neumann, singleboundaryneumann.cpp:27),
 * folded into the interior target column as at :231-242. May be NULL (all 0). */
ORC_API orc_csr *orc_assemble_divergence(int dim, const int *np, const int *per,
                                         const double *dx, const double *dy,
                                         const double *dz, const double *a0)
{
    orc_grid g;
    grid_init(&g, dim, np, per, dx, dy, dz);
    rowbuf *rows = (rowbuf *)calloc((size_t)g.pN, sizeof(rowbuf));
    for (int k = 0; k < g.np[2]; ++k)
        for (int j = 0; j < g.np[1]; ++j)
            for (int i = 0; i < g.np[0]; ++i)
            {
                const int64_t self = nat_index(&g, 3, i, j, k);
                rowbuf *r = &rows[self];
                for (int f = 0; f < dim; ++f)
                {
                    /* kernel[f] at :140-151 */
                    double value;
                    if (f == 0) value = g.dLp[1][j] * g.dLp[2][k];
                    else if (f == 1) value = g.dLp[0][i] * g.dLp[2][k];
                    else value = g.dLp[0][i] * g.dLp[1][j];
                    const int ii = i - (f == 0), jj = j - (f == 1), kk = k - (f == 2);
                    const int64_t cp = packed_index(&g, f, i, j, k);
                    const int64_t cm = packed_index(&g, f, ii, jj, kk);
                    row_set(r, cp, value, 0);
                    row_set(r, cm, -value, 0);
                    if (a0)
                    {
                        /* ghost on the plus side: target is the last interior face
                         * (i-1); ghost on the minus side: target is face i. */
                        if (cp < 0 && a0[2 * f + 1] != 0.0)
                            row_set(r, cm, value * a0[2 * f + 1], 1);
                        if (cm < 0 && a0[2 * f + 0] != 0.0)
                            row_set(r, cp, -value * a0[2 * f + 0], 1);
                    }
                }
            }
    int64_t nnz = 0;
    for (int64_t p = 0; p < g.pN; ++p) nnz += rows[p].n;
    orc_csr *D = csr_alloc(g.pN, g.voff[dim], nnz);
    nnz = 0;
    for (int64_t p = 0; p < g.pN; ++p)
    {
        row_sort(&rows[p]);
        D->rowptr[p] = nnz;
        for (int q = 0; q < rows[p].n; ++q)
        {
            D->col[nnz] = rows[p].c[q];
            D->val[nnz] = rows[p].v[q];
            ++nnz;
        }
    }
    D->rowptr[g.pN] = nnz;
    free(rows);
    grid_free(&g);
    return D;
}

/* creategradient.cpp:36-135 with normalize = PETSC_FALSE. */
ORC_API orc_csr *orc_assemble_gradient(int dim, const int *np, const int *per,
                                       const double *dx, const double *dy,
                                       const double *dz)
{
    orc_grid g;
    grid_init(&g, dim, np, per, dx, dy, dz);
    const int64_t UN = g.voff[dim];
    orc_csr *G = csr_alloc(UN, g.pN, 2 * UN);
    int64_t nnz = 0;
    for (int f = 0; f < dim; ++f)
        for (int k = 0; k < g.nv[f][2]; ++k)
            for (int j = 0; j < g.nv[f][1]; ++j)
                for (int i = 0; i < g.nv[f][0]; ++i)
                {
                    const int s = (f == 0) ? i : (f == 1) ? j : k;
                    const double v = 1.0 / g.hface[f][s]; /* :72,:78,:84 */
                    const int64_t rId = packed_index(&g, f, i, j, k);
                    rowbuf r;
                    r.n = 0;
                    row_set(&r, packed_index(&g, 3, i, j, k), -v, 0);
                    row_set(&r, packed_index(&g, 3, i + (f == 0), j + (f == 1), k + (f == 2)), v, 0);
                    row_sort(&r);
                    G->rowptr[rId] = nnz;
                    for (int q = 0; q < r.n; ++q)
                    {
                        G->col[nnz] = r.c[q];
                        G->val[nnz] = r.v[q];
                        ++nnz;
                    }
                }
    G->rowptr[UN] = nnz;
    G->nnz = nnz;
    grid_free(&g);
    return G;
}

/* createbn.cpp:19-53 for N = 1: BnHead = dt * I (MatShift on an empty pattern). */
ORC_API orc_csr *orc_bnhead_order1(int64_t n, double dt)
{
    orc_csr *B = csr_alloc(n, n, n);
    for (int64_t i = 0; i < n; ++i)
    {
        B->rowptr[i] = i;
        B->col[i] = (int32_t)i;
        B->val[i] = dt;
    }
    B->rowptr[n] = n;
    return B;
}

/* MatMatMult (PETSc SeqAIJ numeric phase): for each row i of A, for each nonzero
 * a_ik in ascending k, for each nonzero b_kj: c_ij += a_ik * b_kj; result columns
 * sorted.  Called as BN*G then D*(BN*G), navierstokes.cpp:351-356. */
ORC_API orc_csr *orc_matmatmult(const orc_csr *A, const orc_csr *B)
{
    const int64_t n = A->nrows, m = B->ncols;
    int64_t *mark = (int64_t *)malloc(sizeof(int64_t) * (size_t)m);
    double *acc = (double *)malloc(sizeof(double) * (size_t)m);
    for (int64_t j = 0; j < m; ++j) mark[j] = -1;
    /* symbolic */
    int64_t nnz = 0;
    for (int64_t i = 0; i < n; ++i)
        for (int64_t p = A->rowptr[i]; p < A->rowptr[i + 1]; ++p)
        {
            const int64_t k = A->col[p];
            for (int64_t q = B->rowptr[k]; q < B->rowptr[k + 1]; ++q)
                if (mark[B->col[q]] != i)
                {
                    mark[B->col[q]] = i;
                    ++nnz;
                }
        }
    orc_csr *C = csr_alloc(n, m, nnz);
    for (int64_t j = 0; j < m; ++j) mark[j] = -1;
    nnz = 0;
    for (int64_t i = 0; i < n; ++i)
    {
        const int64_t start = nnz;
        C->rowptr[i] = start;
        for (int64_t p = A->rowptr[i]; p < A->rowptr[i + 1]; ++p)
        {
            const int64_t k = A->col[p];
            const double a = A->val[p];
            for (int64_t q = B->rowptr[k]; q < B->rowptr[k + 1]; ++q)
            {
                const int32_t j = B->col[q];
                if (mark[j] < start)
                {
                    mark[j] = nnz;
                    C->col[nnz] = j;
                    acc[j] = a * B->val[q];
                    ++nnz;
                }
                else
                    acc[j] += a * B->val[q];
            }
        }
        /* sort the row's columns, then gather values */
        for (int64_t a1 = start + 1; a1 < nnz; ++a1)
        {
            int32_t c = C->col[a1];
            int64_t b1 = a1 - 1;
            while (b1 >= start && C->col[b1] > c)
            {
                C->col[b1 + 1] = C->col[b1];
                --b1;
            }
            C->col[b1 + 1] = c;
        }
        for (int64_t q = start; q < nnz; ++q) C->val[q] = acc[C->col[q]];
        for (int64_t q = start; q < nnz; ++q) mark[C->col[q]] = -1;
    }
    C->rowptr[n] = nnz;
    free(mark);
    free(acc);
    return C;
}

/* MatAXPY(Y, a, X, DIFFERENT_NONZERO_PATTERN): Y <- Y + a*X on the union pattern, columns sorted. */
static orc_csr *csr_axpy(const orc_csr *Y, double a, const orc_csr *X)
{
    const int64_t n = Y->nrows;
    int64_t nnz = 0;
    for (int64_t i = 0; i < n; ++i)
    {
        int64_t p = Y->rowptr[i], q = X->rowptr[i];
        while (p < Y->rowptr[i + 1] || q < X->rowptr[i + 1])
        {
            const int32_t cy = (p < Y->rowptr[i + 1]) ? Y->col[p] : INT32_MAX;
            const int32_t cx = (q < X->rowptr[i + 1]) ? X->col[q] : INT32_MAX;
            if (cy <= cx) ++p;
            if (cx <= cy) ++q;
            ++nnz;
        }
    }
    orc_csr *Z = csr_alloc(n, Y->ncols, nnz);
    nnz = 0;
    for (int64_t i = 0; i < n; ++i)
    {
        Z->rowptr[i] = nnz;
        int64_t p = Y->rowptr[i], q = X->rowptr[i];
        while (p < Y->rowptr[i + 1] || q < X->rowptr[i + 1])
        {
            const int32_t cy = (p < Y->rowptr[i + 1]) ? Y->col[p] : INT32_MAX;
            const int32_t cx = (q < X->rowptr[i + 1]) ? X->col[q] : INT32_MAX;
            double v = 0.0;
            int32_t c = (cy <= cx) ? cy : cx;
            if (cy <= cx) v = Y->val[p++];
            if (cx <= cy) v += a * X->val[q++];
            Z->col[nnz] = c;
            Z->val[nnz] = v;
            ++nnz;
        }
    }
    Z->rowptr[n] = nnz;
    return Z;
}

static orc_csr *csr_copy(const orc_csr *A)
{
    orc_csr *B = csr_alloc(A->nrows, A->ncols, A->nnz);
    memcpy(B->rowptr, A->rowptr, sizeof(int64_t) * (size_t)(A->nrows + 1));
    memcpy(B->col, A->col, sizeof(int32_t) * (size_t)A->nnz);
    memcpy(B->val, A->val, sizeof(double) * (size_t)A->nnz);
    return B;
}

/* createBnHead for any order N (createbn.cpp:19-95):
 *   BnHead = dt*I + sum_{term=2..N} dt^term * coeff^(term-1) * Op^(term-1)
 * with the reference's evaluation order: rightMat = Op, then (term-2) products Op*rightMat,
 * a = pow(dt,term)*pow(coeff,term-1), MatAXPY(BnHead, a, rightMat). */
ORC_API orc_csr *orc_bnhead(const orc_csr *Op, double dt, double coeff, int N)
{
    if (N < 1 || Op->nrows != Op->ncols) return NULL;
    orc_csr *B = orc_bnhead_order1(Op->nrows, dt);
    for (int term = 2; term <= N; ++term)
    {
        orc_csr *right = csr_copy(Op);
        for (int c = 2; c < term; ++c)
        {
            orc_csr *tmp = orc_matmatmult(Op, right);
            orc_csr_free(right);
            right = tmp;
        }
        const double a = pow(dt, term) * pow(coeff, term - 1);
        orc_csr *Z = csr_axpy(B, a, right);
        orc_csr_free(B);
        orc_csr_free(right);
        B = Z;
    }
    return B;
}

/* The literal pipeline of NavierStokesSolver::createOperators for BN order 1
 * (navierstokes.cpp:347-356): D, G, BN = dt I, BNG = BN*G, DBNG = D*BNG. */
ORC_API orc_csr *orc_assemble_dbng_literal(int dim, const int *np, const int *per,
                                           const double *dx, const double *dy,
                                           const double *dz, double dt, const double *a0)
{
    orc_csr *D = orc_assemble_divergence(dim, np, per, dx, dy, dz, a0);
    orc_csr *G = orc_assemble_gradient(dim, np, per, dx, dy, dz);
    orc_csr *BN = orc_bnhead_order1(G->nrows, dt);
    orc_csr *BNG = orc_matmatmult(BN, G);
    orc_csr *DBNG = orc_matmatmult(D, BNG);
    orc_csr_free(D);
    orc_csr_free(G);
    orc_csr_free(BN);
    orc_csr_free(BNG);
    return DBNG;
}

/* Closed form of the same matrix, row by row, with the
 * reference's product grouping: c = (area) * (dt * (1.0/h)); the diagonal is the
 * MatMatMult accumulation over D-row columns in ascending packed index
 * u(i-1),u(i),v(j-1),v(j),w(k-1),w(k).  Only valid for a0 == 0 (no Neumann wall
 * on a wall-normal velocity).  tests/test_oracle_operator.py proves it bit-equal
 * to orc_assemble_dbng_literal; it exists so that 256^3 can be assembled in
 * seconds for the CPU baseline. */
ORC_API orc_csr *orc_assemble_dbng_closed(int dim, const int *np, const int *per,
                                          const double *dx, const double *dy,
                                          const double *dz, double dt)
{
    for (int f = 0; f < dim; ++f)
        if (per[f] && np[f] < 2) return NULL; /* degenerate: a cell that is its own neighbour */
    orc_grid g;
    grid_init(&g, dim, np, per, dx, dy, dz);
    double *gf[3] = {0, 0, 0};
    for (int f = 0; f < dim; ++f)
    {
        gf[f] = (double *)malloc(sizeof(double) * (size_t)(g.nv[f][f] > 0 ? g.nv[f][f] : 1));
        for (int i = 0; i < g.nv[f][f]; ++i) gf[f][i] = dt * (1.0 / g.hface[f][i]);
    }
    int64_t *rp = (int64_t *)calloc((size_t)g.pN + 1, sizeof(int64_t));
    /* count */
#pragma omp parallel for collapse(2)
    for (int k = 0; k < g.np[2]; ++k)
        for (int j = 0; j < g.np[1]; ++j)
            for (int i = 0; i < g.np[0]; ++i)
            {
                const int idx3[3] = {i, j, k};
                int cnt = 1;
                for (int f = 0; f < dim; ++f)
                {
                    const int s = idx3[f], n = g.np[f];
                    const int hasm = (s > 0) || g.per[f];
                    const int hasp = (s < n - 1) || g.per[f];
                    if (n == 2 && g.per[f]) cnt += 1; /* both neighbours are the same cell */
                    else cnt += hasm + hasp;
                }
                rp[1 + (int64_t)i + (int64_t)g.np[0] * (j + (int64_t)g.np[1] * k)] = cnt;
            }
    for (int64_t p = 0; p < g.pN; ++p) rp[p + 1] += rp[p];
    orc_csr *A = csr_alloc(g.pN, g.pN, rp[g.pN]);
    memcpy(A->rowptr, rp, sizeof(int64_t) * (size_t)(g.pN + 1));
    free(rp);
#pragma omp parallel for collapse(2)
    for (int k = 0; k < g.np[2]; ++k)
        for (int j = 0; j < g.np[1]; ++j)
            for (int i = 0; i < g.np[0]; ++i)
            {
                const int idx3[3] = {i, j, k};
                const int64_t self = (int64_t)i + (int64_t)g.np[0] * (j + (int64_t)g.np[1] * k);
                rowbuf r;
                r.n = 0;
                double diag = 0.0;
                row_set(&r, self, 0.0, 0);
                for (int f = 0; f < dim; ++f)
                {
                    const int s = idx3[f], n = g.np[f];
                    double area;
                    if (f == 0) area = g.dLp[1][j] * g.dLp[2][k];
                    else if (f == 1) area = g.dLp[0][i] * g.dLp[2][k];
                    else area = g.dLp[0][i] * g.dLp[1][j];
                    /* faces touching this cell: minus face = velocity index s-1 (wraps to
                     * nv-1 if periodic), plus face = velocity index s (absent at a wall).
                     * MatMatMult walks the D-row in ascending packed column, i.e. in
                     * ascending face index, which differs from (minus, plus) at a
                     * periodic wrap. */
                    int fm = s - 1, fp = s;
                    if (fm < 0) fm = g.per[f] ? g.nv[f][f] - 1 : -1;
                    if (fp >= g.nv[f][f]) fp = -1;
                    int order[2] = {0, 1}; /* 0 = minus face, 1 = plus face */
                    if (fm >= 0 && fp >= 0 && fm > fp) { order[0] = 1; order[1] = 0; }
                    for (int q = 0; q < 2; ++q)
                    {
                        int nb[3] = {i, j, k};
                        if (order[q] == 0 && fm >= 0)
                        {
                            /* D[p,u(s-1)] = -area ; BNG[u(s-1),p(s)] = dt*(+1/h) ; BNG[u(s-1),p(s-1)] = dt*(-1/h) */
                            const double cdiag = (-area) * gf[f][fm];
                            const double coff = (-area) * (-gf[f][fm]);
                            diag += cdiag;
                            nb[f] = (s - 1 + n) % n;
                            row_set(&r, nat_index(&g, 3, nb[0], nb[1], nb[2]), coff, 1);
                        }
                        if (order[q] == 1 && fp >= 0)
                        {
                            /* D[p,u(s)] = +area ; BNG[u(s),p(s)] = dt*(-1/h) ; BNG[u(s),p(s+1)] = dt*(+1/h) */
                            const double cdiag = area * (-gf[f][fp]);
                            const double coff = area * gf[f][fp];
                            diag += cdiag;
                            nb[f] = (s + 1) % n;
                            row_set(&r, nat_index(&g, 3, nb[0], nb[1], nb[2]), coff, 1);
                        }
                    }
                }
                row_set(&r, self, diag, 1);
                row_sort(&r);
                int64_t q0 = A->rowptr[self];
                for (int q = 0; q < r.n; ++q)
                {
                    A->col[q0 + q] = r.c[q];
                    A->val[q0 + q] = r.v[q];
                }
            }
    for (int f = 0; f < dim; ++f) free(gf[f]);
    grid_free(&g);
    return A;
}

/* ------------------------------------------------------------------------- */
/* 5. Vec / Mat kernels in the style of PETSc's Seq implementations           */
/* ------------------------------------------------------------------------- */

static int g_fast = 0; /* 0: strict serial order (parity); 1: OpenMP+SIMD (timing) */

ORC_API void orc_set_fast(int fast, int nthreads)
{
    g_fast = fast;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#else
    (void)nthreads;
#endif
}
ORC_API int orc_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* MatMult_SeqAIJ: sum = 0; for each stored entry in column order sum += a*x */
ORC_API void orc_spmv(const orc_csr *A, const double *x, double *y)
{
    const int64_t n = A->nrows;
    if (g_fast)
    {
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < n; ++i)
        {
            double s = 0.0;
            for (int64_t p = A->rowptr[i]; p < A->rowptr[i + 1]; ++p) s += A->val[p] * x[A->col[p]];
            y[i] = s;
        }
    }
    else
        for (int64_t i = 0; i < n; ++i)
        {
            double s = 0.0;
            for (int64_t p = A->rowptr[i]; p < A->rowptr[i + 1]; ++p) s += A->val[p] * x[A->col[p]];
            y[i] = s;
        }
}

static double vdot(int64_t n, const double *x, const double *y)
{
    double s = 0.0;
    if (g_fast)
    {
#pragma omp parallel for simd reduction(+ : s) schedule(static)
        for (int64_t i = 0; i < n; ++i) s += x[i] * y[i];
    }
    else
        for (int64_t i = 0; i < n; ++i) s += x[i] * y[i];
    return s;
}
static double vsum(int64_t n, const double *x)
{
    double s = 0.0;
    if (g_fast)
    {
#pragma omp parallel for simd reduction(+ : s) schedule(static)
        for (int64_t i = 0; i < n; ++i) s += x[i];
    }
    else
        for (int64_t i = 0; i < n; ++i) s += x[i];
    return s;
}
/* VecNorm_Seq(NORM_2) = sqrt(BLASdot(x,x)) in PETSc 3.16 */
static double vnorm2(int64_t n, const double *x) { return sqrt(vdot(n, x, x)); }
static void vcopy(int64_t n, const double *x, double *y)
{
    if (g_fast)
    {
#pragma omp parallel for simd schedule(static)
        for (int64_t i = 0; i < n; ++i) y[i] = x[i];
    }
    else
        memcpy(y, x, sizeof(double) * (size_t)n);
}
/* VecAXPY: y += a x */
static void vaxpy(int64_t n, double a, const double *x, double *y)
{
#pragma omp parallel for simd schedule(static) if (g_fast)
    for (int64_t i = 0; i < n; ++i) y[i] += a * x[i];
}
/* VecAYPX: y = x + a y */
static void vaypx(int64_t n, double a, const double *x, double *y)
{
#pragma omp parallel for simd schedule(static) if (g_fast)
    for (int64_t i = 0; i < n; ++i) y[i] = x[i] + a * y[i];
}
static void vshift(int64_t n, double s, double *x)
{
#pragma omp parallel for simd schedule(static) if (g_fast)
    for (int64_t i = 0; i < n; ++i) x[i] += s;
}
static void vpointwise_mult(int64_t n, const double *x, const double *d, double *y)
{
#pragma omp parallel for simd schedule(static) if (g_fast)
    for (int64_t i = 0; i < n; ++i) y[i] = x[i] * d[i];
}

typedef struct
{
    int ksp_type;  /* 0 = cg (linsolverksp.cpp:64 default), 1 = bcgs */
    int pc_type;   /* 0 = none, 1 = jacobi */
    int norm_type; /* 0 none, 1 preconditioned (default), 2 unpreconditioned, 3 natural */
    int max_it;    /* default 10000 */
    double rtol;   /* default 1e-5 */
    double atol;   /* default 1e-50 */
    double divtol; /* default 1e4 */
    int has_const_nullspace; /* navierstokes.cpp:404-413 */
    int n_nullvecs;          /* ibpm.cpp:251-267: 1 explicit vector, no constant */
} orc_ksp_opts;

typedef struct
{
    const orc_csr *A;
    orc_ksp_opts o;
    int64_t n;
    const double *nullvecs; /* n_nullvecs x n, orthonormal */
    double *dinv;           /* PCJacobi: 1/diag (1.0 where diag == 0) */
    double rnorm0, ttol;
} ksp_ctx;

/* MatNullSpaceRemove (matnull.c): constant: sum = VecSum/(-1.0*N); VecShift;
 * vectors: alpha_i = vec . v_i ; vec -= alpha_i v_i */
static void nullspace_remove(const ksp_ctx *c, double *v)
{
    const int64_t n = c->n;
    if (c->o.has_const_nullspace && n > 0)
    {
        double sum = vsum(n, v);
        sum = sum / (-1.0 * (double)n);
        vshift(n, sum, v);
    }
    if (c->o.n_nullvecs > 0)
    {
        double alpha[8];
        for (int q = 0; q < c->o.n_nullvecs; ++q) alpha[q] = -vdot(n, v, c->nullvecs + (size_t)q * n);
        for (int q = 0; q < c->o.n_nullvecs; ++q) vaxpy(n, alpha[q], c->nullvecs + (size_t)q * n, v);
    }
}

/* KSP_PCApply = PCApply + KSP_RemoveNullSpace (left preconditioning) */
static void pc_apply(const ksp_ctx *c, const double *r, double *z)
{
    if (c->o.pc_type == 1) vpointwise_mult(c->n, r, c->dinv, z); /* PCApply_Jacobi: VecPointwiseMult */
    else vcopy(c->n, r, z);                                      /* PCNONE: VecCopy */
    nullspace_remove(c, z);
}

/* KSPConvergedDefault (iterativ.c).  Returns reason (0 = keep iterating). */
static int converged_default(ksp_ctx *c, int it, double rnorm)
{
    if (it == 0)
    {
        c->rnorm0 = rnorm;
        c->ttol = fmax(c->o.rtol * c->rnorm0, c->o.atol);
    }
    if (isnan(rnorm) || isinf(rnorm)) return -9; /* KSP_DIVERGED_NANORINF */
    if (rnorm <= c->ttol) return (rnorm < c->o.atol) ? 3 /* ATOL */ : 2 /* RTOL */;
    if (rnorm >= c->o.divtol * c->rnorm0) return -4; /* KSP_DIVERGED_DTOL */
    return 0;
}

static double sgn(double a) { return (a >= 0.0) ? 1.0 : -1.0; } /* PetscSign */

/* KSPSolve with zero initial guess (x is zeroed, itfunc.c) + KSPSolve_CG (cg.c).
 * W shares storage with Z exactly as in cg.c (W = Z). */
static int solve_cg(ksp_ctx *c, const double *b, double *x, int *its_out, double *rnorm_out,
                    double *hist, int hist_cap, int *hist_len)
{
    const int64_t n = c->n;
    double *R = (double *)malloc(sizeof(double) * (size_t)n);
    double *Z = (double *)malloc(sizeof(double) * (size_t)n);
    double *P = (double *)malloc(sizeof(double) * (size_t)n);
    double *W = Z;
    double dpi = 0.0, a, beta = 0.0, betaold = 1.0, bb, dpiold, dp = 0.0;
    int reason = 0, its = 0, nh = 0, i;
    const int nt = c->o.norm_type;

    memset(x, 0, sizeof(double) * (size_t)n);
    vcopy(n, b, R); /* r <- b (x is 0) */

    switch (nt)
    {
    case 1:
        pc_apply(c, R, Z);
        dp = vnorm2(n, Z);
        break;
    case 2:
        dp = vnorm2(n, R);
        break;
    case 3:
        pc_apply(c, R, Z);
        beta = vdot(n, Z, R);
        dp = sqrt(fabs(beta));
        break;
    default:
        dp = 0.0;
    }
    if (nh < hist_cap) hist[nh] = dp;
    nh++;
    reason = converged_default(c, 0, dp);
    if (reason) goto done;

    if (nt != 1 && nt != 3) pc_apply(c, R, Z);
    if (nt != 3) beta = vdot(n, Z, R);

    i = 0;
    do
    {
        its = i + 1;
        if (beta == 0.0)
        {
            reason = 3; /* KSP_CONVERGED_ATOL: "converged due to beta = 0" */
            break;
        }
        else if (i > 0 && beta * betaold < 0.0)
        {
            reason = -8; /* KSP_DIVERGED_INDEFINITE_PC */
            break;
        }
        if (!i)
        {
            vcopy(n, Z, P);
            bb = 0.0;
        }
        else
        {
            bb = beta / betaold;
            vaypx(n, bb, Z, P); /* p <- z + b p */
        }
        dpiold = dpi;
        orc_spmv(c->A, P, W); /* w <- A p */
        dpi = vdot(n, P, W);
        betaold = beta;
        if (dpi == 0.0 || (i > 0 && sgn(dpi) * sgn(dpiold) < 0.0))
        {
            reason = -10; /* KSP_DIVERGED_INDEFINITE_MAT */
            break;
        }
        a = beta / dpi;
        vaxpy(n, a, P, x);  /* x <- x + a p */
        vaxpy(n, -a, W, R); /* r <- r - a w */
        if (nt == 1)
        {
            pc_apply(c, R, Z);
            dp = vnorm2(n, Z);
        }
        else if (nt == 2)
            dp = vnorm2(n, R);
        else if (nt == 3)
        {
            pc_apply(c, R, Z);
            beta = vdot(n, Z, R);
            dp = sqrt(fabs(beta));
        }
        else
            dp = 0.0;
        if (nh < hist_cap) hist[nh] = dp;
        nh++;
        reason = converged_default(c, i + 1, dp);
        if (reason) break;
        if (nt != 1 && nt != 3) pc_apply(c, R, Z);
        if (nt != 3) beta = vdot(n, Z, R);
        i++;
    } while (i < c->o.max_it);
    if (!reason && i >= c->o.max_it) reason = -3; /* KSP_DIVERGED_ITS */
done:
    *its_out = its;
    *rnorm_out = dp;
    *hist_len = nh;
    free(R);
    free(Z);
    free(P);
    return reason;
}

/* y = B (A x)  -- KSP_PCApplyBAorAB with left preconditioning */
static void pc_apply_BA(const ksp_ctx *c, const double *x, double *y, double *work)
{
    orc_spmv(c->A, x, work);
    pc_apply(c, work, y);
}

/* KSPSolve_BCGS (bcgs.c), left preconditioning, zero initial guess. */
static int solve_bcgs(ksp_ctx *c, const double *b, double *x, int *its_out, double *rnorm_out,
                      double *hist, int hist_cap, int *hist_len)
{
    const int64_t n = c->n;
    double *R = (double *)malloc(sizeof(double) * (size_t)n);
    double *RP = (double *)malloc(sizeof(double) * (size_t)n);
    double *V = (double *)calloc((size_t)n, sizeof(double));
    double *T = (double *)malloc(sizeof(double) * (size_t)n);
    double *S = (double *)malloc(sizeof(double) * (size_t)n);
    double *P = (double *)calloc((size_t)n, sizeof(double));
    double *wk = (double *)malloc(sizeof(double) * (size_t)n);
    double rho, rhoold = 1.0, alpha = 1.0, beta, omega, omegaold = 1.0, d1, d2, dp = 0.0;
    int reason = 0, its = 0, nh = 0, i;

    memset(x, 0, sizeof(double) * (size_t)n);
    /* KSPInitialResidual, zero guess, left PC: r = B b */
    pc_apply(c, b, R);
    dp = (c->o.norm_type != 0) ? vnorm2(n, R) : 0.0;
    if (nh < hist_cap) hist[nh] = dp;
    nh++;
    reason = converged_default(c, 0, dp);
    if (reason) goto done;
    vcopy(n, R, RP);
    i = 0;
    do
    {
        rho = vdot(n, R, RP);
        beta = (rho / rhoold) * (alpha / omegaold);
        /* VecAXPBYPCZ(P,1.0,-omegaold*beta,beta,R,V): p = 1*r + (-omegaold*beta)*v + beta*p */
        {
            const double bq = -omegaold * beta;
            for (int64_t q = 0; q < n; ++q) P[q] = R[q] + bq * V[q] + beta * P[q];
        }
        pc_apply_BA(c, P, V, wk);
        d1 = vdot(n, V, RP);
        if (d1 == 0.0)
        {
            reason = -5; /* KSP_DIVERGED_BREAKDOWN */
            break;
        }
        alpha = rho / d1;
        for (int64_t q = 0; q < n; ++q) S[q] = -alpha * V[q] + R[q]; /* VecWAXPY(S,-alpha,V,R) */
        pc_apply_BA(c, S, T, wk);
        d1 = vdot(n, S, T); /* VecDotNorm2(S,T,&d1,&d2) */
        d2 = vdot(n, T, T);
        if (d2 == 0.0)
        {
            d1 = vdot(n, S, S);
            if (d1 != 0.0)
            {
                reason = -5;
                break;
            }
            vaxpy(n, alpha, P, x);
            its++;
            dp = 0.0;
            reason = 2; /* KSP_CONVERGED_RTOL */
            if (nh < hist_cap) hist[nh] = dp;
            nh++;
            break;
        }
        omega = d1 / d2;
        for (int64_t q = 0; q < n; ++q) x[q] = alpha * P[q] + omega * S[q] + x[q]; /* VecAXPBYPCZ(X,alpha,omega,1.0,P,S) */
        for (int64_t q = 0; q < n; ++q) R[q] = -omega * T[q] + S[q];               /* VecWAXPY(R,-omega,T,S) */
        if (c->o.norm_type != 0) dp = vnorm2(n, R);
        rhoold = rho;
        omegaold = omega;
        its++;
        if (nh < hist_cap) hist[nh] = dp;
        nh++;
        reason = converged_default(c, i + 1, dp);
        if (reason) break;
        if (rho == 0.0)
        {
            reason = -5;
            break;
        }
        i++;
    } while (i < c->o.max_it);
    if (!reason && i >= c->o.max_it) reason = -3;
done:
    *its_out = its;
    *rnorm_out = dp;
    *hist_len = nh;
    free(R);
    free(RP);
    free(V);
    free(T);
    free(S);
    free(P);
    free(wk);
    return reason;
}

/* What LinSolverKSP::solve does (linsolverksp.cpp:85-105): KSPSolve(ksp, b, x) on
 * the operator given to setMatrix, then KSPGetConvergedReason.  Returns the
 * KSPConvergedReason; its/rnorm are KSPGetIterationNumber / KSPGetResidualNorm
 * (:110-132); hist is what KSPSetResidualHistory would log. */
ORC_API int orc_ksp_solve(const orc_csr *A, const orc_ksp_opts *opts, const double *nullvecs,
                          const double *b, double *x, int *its, double *rnorm, double *hist,
                          int hist_cap, int *hist_len)
{
    ksp_ctx c;
    memset(&c, 0, sizeof(c));
    c.A = A;
    c.o = *opts;
    c.n = A->nrows;
    c.nullvecs = nullvecs;
    if (opts->pc_type == 1)
    {
        /* PCSetUp_Jacobi: MatGetDiagonal, VecReciprocal, zero diagonal -> 1.0 */
        c.dinv = (double *)malloc(sizeof(double) * (size_t)c.n);
        for (int64_t i = 0; i < c.n; ++i)
        {
            double d = 0.0;
            for (int64_t p = A->rowptr[i]; p < A->rowptr[i + 1]; ++p)
                if (A->col[p] == i) d = A->val[p];
            c.dinv[i] = (d != 0.0) ? 1.0 / d : 1.0;
        }
    }
    int reason;
    if (opts->ksp_type == 1) reason = solve_bcgs(&c, b, x, its, rnorm, hist, hist_cap, hist_len);
    else reason = solve_cg(&c, b, x, its, rnorm, hist, hist_cap, hist_len);
    free(c.dinv);
    return reason;
}

/* ------------------------------------------------------------------------------------------------
 * Convection operator N(q) of createconvection.cpp (a MatShell: ConvectionMult2D :205-262 /
 * ConvectionMult3D :266-332 with the point kernels kernelU/kernelV/kernelW :39-200).
 *
 * Input: the LOCAL (ghosted) arrays of the velocity fields, what DMCompositeScatterArray +
 * Boundary::copyValues2LocalVecs leave in ctx->qLocal (:217-221, :285-289): field f has nf[3f+0..2]
 * points per direction, stored with one ghost layer on every side, i fastest:
 *     q[f][(i+1) + (nx_f+2)*((j+1) + (ny_f+2)*(k+1))]     (2-D: no k)
 * dL[3f+d] = mesh->dL[f][d] (interior part: index 0 = first real point; cartesianmesh.cpp:212-355,
 * produced by orc_velocity_axis).  Output: the three fields, interior points only, i fastest
 * (the blocks of the packed vector [u | v | w]).
 * The expressions are the reference's, term by term and in its order; this file is compiled with
 * -ffp-contract=off, which is what an x86-64 build of PetIBM without -march flags executes.
 * No golden vector: the reference has no test for this operator ("parity unpinned" for it). */
#define QF(f, i, j, k) \
    q[f][((i) + 1) + (size_t)(nf[3 * (f)] + 2) * (((j) + 1) + (size_t)(nf[3 * (f) + 1] + 2) * (dim == 3 ? (k) + 1 : 0))]
ORC_API void orc_convection(int dim, const int *nf, const double *const *dL, const double *const *q,
                            double *const *out)
{
    for (int f = 0; f < dim; ++f)
    {
        const int nx = nf[3 * f], ny = nf[3 * f + 1], nz = dim == 3 ? nf[3 * f + 2] : 1;
        const double *dLx = dL[3 * f], *dLy = dL[3 * f + 1], *dLz = dL[3 * f + 2];
        for (int k = 0; k < nz; ++k)
            for (int j = 0; j < ny; ++j)
                for (int i = 0; i < nx; ++i)
                {
                    double r;
                    if (f == 0)
                    { /* kernelU :39-64 (2-D), :94-126 (3-D) */
                        const double uSelf = QF(0, i, j, k);
                        const double uW = (uSelf + QF(0, i - 1, j, k)) / 2.0;
                        const double uE = (uSelf + QF(0, i + 1, j, k)) / 2.0;
                        const double uS = (uSelf + QF(0, i, j - 1, k)) / 2.0;
                        const double uN = (uSelf + QF(0, i, j + 1, k)) / 2.0;
                        const double vS = (QF(1, i, j - 1, k) + QF(1, i + 1, j - 1, k)) / 2.0;
                        const double vN = (QF(1, i, j, k) + QF(1, i + 1, j, k)) / 2.0;
                        r = (uE * uE - uW * uW) / dLx[i] + (vN * uN - vS * uS) / dLy[j];
                        if (dim == 3)
                        {
                            const double uB = (uSelf + QF(0, i, j, k - 1)) / 2.0;
                            const double uF = (uSelf + QF(0, i, j, k + 1)) / 2.0;
                            const double wB = (QF(2, i, j, k - 1) + QF(2, i + 1, j, k - 1)) / 2.0;
                            const double wF = (QF(2, i, j, k) + QF(2, i + 1, j, k)) / 2.0;
                            r = r + (wF * uF - wB * uB) / dLz[k];
                        }
                    }
                    else if (f == 1)
                    { /* kernelV :67-91, :129-162 */
                        const double vSelf = QF(1, i, j, k);
                        const double uW = (QF(0, i - 1, j, k) + QF(0, i - 1, j + 1, k)) / 2.0;
                        const double uE = (QF(0, i, j, k) + QF(0, i, j + 1, k)) / 2.0;
                        const double vW = (vSelf + QF(1, i - 1, j, k)) / 2.0;
                        const double vE = (vSelf + QF(1, i + 1, j, k)) / 2.0;
                        const double vS = (vSelf + QF(1, i, j - 1, k)) / 2.0;
                        const double vN = (vSelf + QF(1, i, j + 1, k)) / 2.0;
                        r = (uE * vE - uW * vW) / dLx[i] + (vN * vN - vS * vS) / dLy[j];
                        if (dim == 3)
                        {
                            const double vB = (vSelf + QF(1, i, j, k - 1)) / 2.0;
                            const double vF = (vSelf + QF(1, i, j, k + 1)) / 2.0;
                            const double wB = (QF(2, i, j, k - 1) + QF(2, i, j + 1, k - 1)) / 2.0;
                            const double wF = (QF(2, i, j, k) + QF(2, i, j + 1, k)) / 2.0;
                            r = r + (wF * vF - wB * vB) / dLz[k];
                        }
                    }
                    else
                    { /* kernelW :165-200 */
                        const double wSelf = QF(2, i, j, k);
                        const double uW = (QF(0, i - 1, j, k) + QF(0, i - 1, j, k + 1)) / 2.0;
                        const double uE = (QF(0, i, j, k) + QF(0, i, j, k + 1)) / 2.0;
                        const double vS = (QF(1, i, j - 1, k) + QF(1, i, j - 1, k + 1)) / 2.0;
                        const double vN = (QF(1, i, j, k) + QF(1, i, j, k + 1)) / 2.0;
                        const double wW = (wSelf + QF(2, i - 1, j, k)) / 2.0;
                        const double wE = (wSelf + QF(2, i + 1, j, k)) / 2.0;
                        const double wS = (wSelf + QF(2, i, j - 1, k)) / 2.0;
                        const double wN = (wSelf + QF(2, i, j + 1, k)) / 2.0;
                        const double wB = (wSelf + QF(2, i, j, k - 1)) / 2.0;
                        const double wF = (wSelf + QF(2, i, j, k + 1)) / 2.0;
                        r = (uE * wE - uW * wW) / dLx[i] + (vN * wN - vS * wS) / dLy[j] +
                            (wF * wF - wB * wB) / dLz[k];
                    }
                    out[f][i + (size_t)nx * (j + (size_t)ny * k)] = r;
                }
    }
}
#undef QF
