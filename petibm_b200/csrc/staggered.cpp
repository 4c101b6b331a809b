// staggered.cpp -- host-only analysis of an assembled staggered-grid operator (part of libb200ls.so).
//
// Two matrices PetIBM hands to LinSolver::setMatrix are NOT the separable pressure operator of
// b200ls_set_poisson_stencil, but are still "7-point stencils with one-dimensional coefficients":
//
//   * the implicit velocity system A = I/dt - c nu L (navierstokes.cpp:342-344) on the packed vector
//     [u | v | w]: per field a 5-/7-point Laplacian whose off-diagonal in direction d depends only on the
//     index along d (createlaplacian.cpp:134-151: 1/(dLNeg*dLSelf), 1/(dLPos*dLSelf)), while the diagonal
//     also carries the ghost-point coefficients folded in with the boundary's a0 (:157-159, 232-243);
//   * IBPM's modified Poisson system [D;E] BN [G,-H] (ibpm.cpp:100-203): the pressure block is the 5-/7-point
//     operator, plus sparse coupling rows/columns for the Lagrangian forces behind it (appendix A.4).
//
// b200ls_staggered_analyze reads that structure OUT OF THE MATRIX (nothing is recomputed from the mesh or the
// boundary conditions, so there is no second source of truth): per field and direction two 1-D arrays
// (minus/plus neighbour), the diagonal as a vector, and whatever does not belong to the stencil blocks as a
// CSR "remainder".  Every row is then checked entry by entry, bitwise, against that description; any entry that
// does not fit makes the analysis fail (B200LS_ERR_MISMATCH) and the caller keeps the matrix as plain CSR.
// The device kernels (sep_kernels.cuh) add the terms of a row in ascending column order exactly like
// MatMult_SeqAIJ, so the structured operator is bit-identical to the assembled one at 24 B/row instead of ~100.
#include "../../include/b200ls.h"

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <vector>

namespace {

int mismatch(char *errbuf, size_t errlen, const char *fmt, ...)
{
    if (errbuf && errlen)
    {
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(errbuf, errlen, fmt, ap);
        va_end(ap);
    }
    return B200LS_ERR_MISMATCH;
}

struct Field
{
    int64_t n[3], off, size;
    int64_t coef_off[3][2];  // offsets of cm / cp of direction d inside the packed coefficient array
};

// entry (row, column) of a CSR matrix with sorted rows, or nullptr
const double *find_entry(const int64_t *rowptr, const int32_t *col, const double *val, int64_t row, int64_t c)
{
    int64_t lo = rowptr[row], hi = rowptr[row + 1];
    while (lo < hi)
    {
        const int64_t mid = (lo + hi) / 2;
        if (col[mid] < c) lo = mid + 1;
        else hi = mid;
    }
    return (lo < rowptr[row + 1] && col[lo] == c) ? val + lo : nullptr;
}

// Core of both analyses.  area != nullptr: the stencil block is the pressure operator D (dt I) G itself, whose
// coefficient in direction d is (product of the two other cell widths) * coef -- area[a] are the three width arrays and
// coef the face arrays dt/h (given, not discovered).
int check_rows(int nfields, const Field *F, const int *periodic, const double *const *area, int64_t nsep, int64_t nrows,
               const int64_t *rowptr, const int32_t *col, const double *val, const double *coef, double *diag,
               int64_t *rem_rowptr, int32_t *rem_col, double *rem_val, char *errbuf, size_t errlen)
{
    int64_t rem = 0;
    for (int f = 0; f < nfields; ++f)
    {
        const int64_t n0 = F[f].n[0], n1 = F[f].n[1];
        const int64_t stride[3] = {1, n0, n0 * n1};
        for (int64_t l = 0; l < F[f].size; ++l)
        {
            const int64_t row = F[f].off + l;
            const int64_t idx[3] = {l % n0, (l / n0) % n1, l / (n0 * n1)};
            // expected neighbours: (column, coefficient)
            int64_t ecol[6];
            double eval[6];
            bool seen[6];
            int ne = 0;
            for (int d = 0; d < 3; ++d)
            {
                const int64_t n = F[f].n[d], s = idx[d];
                const bool per = periodic[d] && n >= 3;
                double cm = coef[F[f].coef_off[d][0] + s], cp = coef[F[f].coef_off[d][1] + s];
                if (area)
                {
                    // same grouping as createdivergence.cpp:142,146,150 and the kernels: (w_a * w_b) * g
                    const int a = d == 0 ? 1 : 0, b = d == 2 ? 1 : 2;
                    const volatile double ar = area[a][idx[a]] * area[b][idx[b]];
                    const volatile double pm = ar * cm, pp = ar * cp;
                    cm = pm;
                    cp = pp;
                }
                if (cm != 0.0 && (s > 0 || per))
                {
                    ecol[ne] = row + ((s > 0) ? -stride[d] : (n - 1) * stride[d]);
                    eval[ne] = cm;
                    seen[ne++] = false;
                }
                if (cp != 0.0 && (s < n - 1 || per))
                {
                    ecol[ne] = row + ((s < n - 1) ? stride[d] : -(n - 1) * stride[d]);
                    eval[ne] = cp;
                    seen[ne++] = false;
                }
            }
            bool have_diag = false;
            rem_rowptr[row] = rem;
            for (int64_t q = rowptr[row]; q < rowptr[row + 1]; ++q)
            {
                const int64_t c = col[q];
                if (c == row)
                {
                    diag[row] = val[q];
                    have_diag = true;
                }
                else if (c < nsep)
                {
                    int e = -1;
                    for (int t = 0; t < ne; ++t)
                        if (ecol[t] == c) e = t;
                    if (e < 0)
                    {
                        if (val[q] == 0.0) continue;  // an explicit zero outside the pattern contributes nothing
                        return mismatch(errbuf, errlen, "row %lld: entry in column %lld is outside the 7-point pattern", (long long)row, (long long)c);
                    }
                    if (std::memcmp(&val[q], &eval[e], sizeof(double)) != 0)
                        return mismatch(errbuf, errlen, "row %lld, column %lld: %.17g is not the %s %.17g", (long long)row, (long long)c, val[q],
                                        area ? "closed-form coefficient (face area * dt/h)" : "line coefficient", eval[e]);
                    seen[e] = true;
                }
                else
                {
                    if (rem_col) rem_col[rem] = (int32_t)c;
                    if (rem_val) rem_val[rem] = val[q];
                    ++rem;
                }
            }
            if (!have_diag) return mismatch(errbuf, errlen, "row %lld has no diagonal entry", (long long)row);
            for (int t = 0; t < ne; ++t)
                if (!seen[t]) return mismatch(errbuf, errlen, "row %lld: the entry in column %lld is missing", (long long)row, (long long)ecol[t]);
        }
    }
    for (int64_t row = nsep; row < nrows; ++row)
    {
        rem_rowptr[row] = rem;
        for (int64_t q = rowptr[row]; q < rowptr[row + 1]; ++q)
        {
            if (rem_col) rem_col[rem] = col[q];
            if (rem_val) rem_val[rem] = val[q];
            ++rem;
        }
    }
    rem_rowptr[nrows] = rem;
    return B200LS_OK;
}

// rows must be sorted by column (PETSc AIJ rows are): the kernels add the terms in ascending column order
int check_sorted(int64_t nrows, const int64_t *rowptr, const int32_t *col, char *errbuf, size_t errlen)
{
    for (int64_t r = 0; r < nrows; ++r)
        for (int64_t q = rowptr[r]; q < rowptr[r + 1]; ++q)
        {
            if (col[q] < 0 || col[q] >= nrows) return B200LS_ERR_ARG;
            if (q > rowptr[r] && col[q] <= col[q - 1]) return mismatch(errbuf, errlen, "row %lld is not sorted by column", (long long)r);
        }
    return B200LS_OK;
}

}  // namespace

extern "C" {

int64_t b200ls_staggered_coef_size(int nfields, const int64_t *dims)
{
    if (nfields < 1 || nfields > 3 || !dims) return -1;
    int64_t s = 0;
    for (int f = 0; f < nfields; ++f)
        for (int d = 0; d < 3; ++d) s += 2 * dims[3 * f + d];
    return s;
}

int b200ls_staggered_analyze(int nfields, const int64_t *dims, const int *periodic, int64_t nrows, const int64_t *rowptr,
                             const int32_t *col, const double *val, double *coef, double *diag, int64_t *rem_rowptr,
                             int32_t *rem_col, double *rem_val, char *errbuf, size_t errlen)
{
    if (nfields < 1 || nfields > 3 || !dims || !periodic || nrows < 1 || !rowptr || !col || !val || !coef || !diag || !rem_rowptr)
        return B200LS_ERR_ARG;
    if (errbuf && errlen) errbuf[0] = 0;
    Field F[3];
    int64_t nsep = 0, cpos = 0;
    for (int f = 0; f < nfields; ++f)
    {
        F[f].off = nsep;
        F[f].size = 1;
        for (int d = 0; d < 3; ++d)
        {
            F[f].n[d] = dims[3 * f + d];
            if (F[f].n[d] < 1) return B200LS_ERR_ARG;
            // a periodic axis with two cells folds both neighbours into one column: not separable
            if (periodic[d] && F[f].n[d] == 2) return mismatch(errbuf, errlen, "periodic axis %d with two cells", d);
            F[f].size *= F[f].n[d];
            F[f].coef_off[d][0] = cpos;
            F[f].coef_off[d][1] = cpos + F[f].n[d];
            cpos += 2 * F[f].n[d];
        }
        nsep += F[f].size;
    }
    if (nsep > nrows) return mismatch(errbuf, errlen, "the fields hold %lld points but the matrix has %lld rows", (long long)nsep, (long long)nrows);
    {
        const int rc = check_sorted(nrows, rowptr, col, errbuf, errlen);
        if (rc != B200LS_OK) return rc;
    }

    // ---- 1. read the 1-D coefficient arrays from one representative line per field and direction
    for (int f = 0; f < nfields; ++f)
    {
        const int64_t stride[3] = {1, F[f].n[0], F[f].n[0] * F[f].n[1]};
        for (int d = 0; d < 3; ++d)
        {
            const int64_t n = F[f].n[d];
            const bool per = periodic[d] && n >= 3;
            for (int64_t s = 0; s < n; ++s)
            {
                const int64_t row = F[f].off + s * stride[d];
                double cm = 0.0, cp = 0.0;
                if (s > 0 || per)
                {
                    const int64_t nb = (s > 0) ? s - 1 : n - 1;
                    const double *e = find_entry(rowptr, col, val, row, F[f].off + nb * stride[d]);
                    if (e) cm = *e;
                }
                if (s < n - 1 || per)
                {
                    const int64_t nb = (s < n - 1) ? s + 1 : 0;
                    const double *e = find_entry(rowptr, col, val, row, F[f].off + nb * stride[d]);
                    if (e) cp = *e;
                }
                coef[F[f].coef_off[d][0] + s] = cm;
                coef[F[f].coef_off[d][1] + s] = cp;
            }
        }
    }

    // ---- 2. check every row against the description; split off the remainder
    return check_rows(nfields, F, periodic, nullptr, nsep, nrows, rowptr, col, val, coef, diag, rem_rowptr, rem_col, rem_val, errbuf,
                      errlen);
}

// The pressure operator D (dt I) G of the mesh followed by rows/columns that are not part of it (IBPM's modified Poisson
// system [D;E] BN [G,-H], ibpm.cpp:100-203, on a STRETCHED grid, where the face areas make the coefficients vary along
// all axes).  The stencil block is checked bitwise against the closed form of b200ls_set_poisson_stencil (appendix A.1 of
// (w_a w_b) * (dt * (1/h))); g receives the face arrays gx (nx+1) | gy (ny+1) | gz (nz+1), diag the stored
// diagonal, the rest goes to the CSR remainder.
int b200ls_hybrid_analyze(int dim, const int64_t *n, const int *periodic, const double *dx, const double *dy, const double *dz,
                          double dt, int64_t nrows, const int64_t *rowptr, const int32_t *col, const double *val, double *g,
                          double *diag, int64_t *rem_rowptr, int32_t *rem_col, double *rem_val, char *errbuf, size_t errlen)
{
    if ((dim != 2 && dim != 3) || !n || !periodic || !dx || !dy || (dim == 3 && !dz) || nrows < 1 || !rowptr || !col || !val || !g ||
        !diag || !rem_rowptr)
        return B200LS_ERR_ARG;
    if (errbuf && errlen) errbuf[0] = 0;
    const double one = 1.0;
    const double *w[3] = {dx, dy, dim == 3 ? dz : &one};
    Field F;
    F.off = 0;
    F.size = 1;
    int per[3];
    int64_t gpos = 0;
    for (int d = 0; d < 3; ++d)
    {
        F.n[d] = d < dim ? n[d] : 1;
        if (F.n[d] < 1) return B200LS_ERR_ARG;
        per[d] = (d < dim && periodic[d]) ? 1 : 0;
        if (per[d] && F.n[d] < 3) return mismatch(errbuf, errlen, "periodic axis %d needs at least three cells", d);
        F.size *= F.n[d];
        // face arrays exactly as b200ls_set_poisson_stencil builds them: g[s] = dt * (1 / (0.5 (w[s-1] + w[s])))
        double *gd = g + gpos;
        const int64_t m = F.n[d];
        for (int64_t s = 0; s <= m; ++s) gd[s] = 0.0;
        if (d < dim)
        {
            for (int64_t s = 1; s < m; ++s)
            {
                const volatile double hh = 0.5 * (w[d][s] + w[d][s - 1]);
                const volatile double inv = 1.0 / hh;
                gd[s] = dt * inv;
            }
            if (per[d])
            {
                const volatile double hh = 0.5 * (w[d][0] + w[d][m - 1]);
                const volatile double inv = 1.0 / hh;
                gd[0] = gd[m] = dt * inv;
            }
        }
        F.coef_off[d][0] = gpos;      // minus face of cell s
        F.coef_off[d][1] = gpos + 1;  // plus face
        gpos += m + 1;
    }
    if (F.size > nrows) return mismatch(errbuf, errlen, "the grid holds %lld cells but the matrix has %lld rows", (long long)F.size, (long long)nrows);
    {
        const int rc = check_sorted(nrows, rowptr, col, errbuf, errlen);
        if (rc != B200LS_OK) return rc;
    }
    return check_rows(1, &F, per, w, F.size, nrows, rowptr, col, val, g, diag, rem_rowptr, rem_col, rem_val, errbuf, errlen);
}

}  // extern "C"
