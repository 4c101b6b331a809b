// repart.cpp -- host-only part of libb200ls.so: the mapping between PETSc's DMDA ordering of the
// pressure grid and the slab partition the device solver works in.
//
// Inside PetIBM the pressure vectors and the rows of DBNG are distributed the way
// DMDACreate{2,3}d(..., PETSC_DECIDE, ...) chose (src/mesh/cartesianmesh.cpp:500-538): an m x n x p
// process grid, rank r = px + m*(py + n*pz) owning ONE box [xs,xe) x [ys,ye) x [zs,ze) whose points are
// numbered contiguously, i fastest (cartesianmesh.cpp:709-721, AOApplicationToPetsc).  Along every
// axis the first (M mod m) ranks own one more cell (PETSc's default ownership).  The device solver
// cuts the same grid into slabs along the slowest axis (z in 3-D, y in 2-D), one slab per rank, so
// that only face halos of width one cross GPUs.  This file plans the exchange between the two
// distributions (what VecScatter does inside PETSc); the exchange itself is one MPI_Alltoallv per
// direction in the PetIBM shim (torch.distributed.all_to_all_single in the Python harness):
//
//   box -> slab : the part of a box that falls into one slab is a contiguous run of whole z-planes of
//                 the box, so the SEND side needs no packing (counts/displacements into the local
//                 vector); the receive side scatters sub-boxes into full planes (unpack_slab).
//   slab -> box : the reverse (pack_slab gathers the sub-boxes; the receive side is contiguous).
//
// No CUDA in this file and no arithmetic on the data: indices only.
#include "../../include/b200ls.h"

#include <algorithm>
#include <cstring>
#include <vector>

namespace {

// PETSc default ownership along one axis: starts[q] of rank q, starts[m] = M
void split_axis(int64_t M, int m, std::vector<int64_t> &starts)
{
    starts.resize((size_t)m + 1);
    const int64_t base = M / m, rem = M % m;
    for (int q = 0; q <= m; ++q) starts[(size_t)q] = q * base + std::min<int64_t>(q, rem);
}

}  // namespace

struct b200ls_repart
{
    int dim = 3;
    int64_t n[3] = {1, 1, 1};     // normalised to three axes: a 2-D grid is (nx, 1, ny)
    int procs[3] = {1, 1, 1};     // normalised the same way: (m, 1, n)
    int rank = 0, nranks = 1;
    std::vector<int64_t> sx, sy, sz;   // box ownership starts per axis
    std::vector<int64_t> slab;         // slab starts along the slow axis (nranks + 1)
    std::vector<int64_t> offset;       // first PETSc global index of every rank's box (nranks + 1)
    bool identity = false;             // the box partition IS the slab partition (1 x 1 x P)

    void box_of(int q, int64_t lo[3], int64_t hi[3]) const
    {
        const int px = q % procs[0], py = (q / procs[0]) % procs[1], pz = q / (procs[0] * procs[1]);
        lo[0] = sx[(size_t)px];
        hi[0] = sx[(size_t)px + 1];
        lo[1] = sy[(size_t)py];
        hi[1] = sy[(size_t)py + 1];
        lo[2] = sz[(size_t)pz];
        hi[2] = sz[(size_t)pz + 1];
    }
    // planes of box q that fall into slab s
    void overlap(int q, int s, int64_t &k0, int64_t &k1) const
    {
        int64_t lo[3], hi[3];
        box_of(q, lo, hi);
        k0 = std::max(lo[2], slab[(size_t)s]);
        k1 = std::min(hi[2], slab[(size_t)s + 1]);
        if (k1 < k0) k1 = k0;
    }
};

namespace {
// copy between the exchange buffer on the slab side (sub-boxes ordered by peer rank) and the slab-ordered vector
template <bool TO_SLAB>
void slab_copy(const b200ls_repart *p, double *xbuf, double *slabv)
{
    const int64_t nx = p->n[0], ny = p->n[1];
    const int64_t z0 = p->slab[(size_t)p->rank];
    int64_t pos = 0;
    for (int q = 0; q < p->nranks; ++q)
    {
        int64_t lo[3], hi[3], k0, k1;
        p->box_of(q, lo, hi);
        p->overlap(q, p->rank, k0, k1);
        const int64_t xm = hi[0] - lo[0];
        for (int64_t k = k0; k < k1; ++k)
            for (int64_t j = lo[1]; j < hi[1]; ++j)
            {
                double *row = slabv + lo[0] + nx * (j + ny * (k - z0));
                if (TO_SLAB) std::memcpy(row, xbuf + pos, sizeof(double) * (size_t)xm);
                else std::memcpy(xbuf + pos, row, sizeof(double) * (size_t)xm);
                pos += xm;
            }
    }
}
}  // namespace

extern "C" {

int b200ls_dmda_split(int64_t M, int m, int64_t *starts)
{
    if (M < 0 || m <= 0 || !starts) return B200LS_ERR_ARG;
    std::vector<int64_t> s;
    split_axis(M, m, s);
    std::memcpy(starts, s.data(), sizeof(int64_t) * s.size());
    return B200LS_OK;
}

int b200ls_repart_create(b200ls_repart **out, int dim, const int64_t n[3], const int procs[3], int rank)
{
    if (!out || !n || !procs || (dim != 2 && dim != 3)) return B200LS_ERR_ARG;
    b200ls_repart *p = new b200ls_repart();
    p->dim = dim;
    if (dim == 3)
        for (int d = 0; d < 3; ++d)
        {
            p->n[d] = n[d];
            p->procs[d] = procs[d];
        }
    else
    {
        p->n[0] = n[0];
        p->n[1] = 1;
        p->n[2] = n[1];
        p->procs[0] = procs[0];
        p->procs[1] = 1;
        p->procs[2] = procs[1];
    }
    bool ok = true;
    for (int d = 0; d < 3; ++d) ok = ok && p->n[d] >= 1 && p->procs[d] >= 1 && (int64_t)p->procs[d] <= p->n[d];
    p->nranks = p->procs[0] * p->procs[1] * p->procs[2];
    // the slab partition gives every rank at least one plane of the slow axis
    ok = ok && rank >= 0 && rank < p->nranks && (int64_t)p->nranks <= p->n[2];
    if (!ok)
    {
        delete p;
        return B200LS_ERR_ARG;
    }
    p->rank = rank;
    split_axis(p->n[0], p->procs[0], p->sx);
    split_axis(p->n[1], p->procs[1], p->sy);
    split_axis(p->n[2], p->procs[2], p->sz);
    split_axis(p->n[2], p->nranks, p->slab);
    p->offset.assign((size_t)p->nranks + 1, 0);
    for (int q = 0; q < p->nranks; ++q)
    {
        int64_t lo[3], hi[3];
        p->box_of(q, lo, hi);
        p->offset[(size_t)q + 1] = p->offset[(size_t)q] + (hi[0] - lo[0]) * (hi[1] - lo[1]) * (hi[2] - lo[2]);
    }
    p->identity = (p->procs[0] == 1 && p->procs[1] == 1);
    *out = p;
    return B200LS_OK;
}

int b200ls_repart_destroy(b200ls_repart *p)
{
    delete p;
    return B200LS_OK;
}

int b200ls_repart_info(const b200ls_repart *p, int64_t box_lo[3], int64_t box_hi[3], int64_t *nbox, int64_t *slab_lo,
                       int64_t *slab_hi, int64_t *nslab, int *identity)
{
    if (!p) return B200LS_ERR_ARG;
    int64_t lo[3], hi[3];
    p->box_of(p->rank, lo, hi);
    for (int d = 0; d < 3; ++d)
    {
        if (box_lo) box_lo[d] = lo[d];
        if (box_hi) box_hi[d] = hi[d];
    }
    if (nbox) *nbox = p->offset[(size_t)p->rank + 1] - p->offset[(size_t)p->rank];
    if (slab_lo) *slab_lo = p->slab[(size_t)p->rank];
    if (slab_hi) *slab_hi = p->slab[(size_t)p->rank + 1];
    if (nslab) *nslab = p->n[0] * p->n[1] * (p->slab[(size_t)p->rank + 1] - p->slab[(size_t)p->rank]);
    if (identity) *identity = p->identity ? 1 : 0;
    return B200LS_OK;
}

// box -> slab direction.  box_counts/box_displs: what this rank sends out of (slab -> box: receives into) its
// box-ordered local vector, per peer, in doubles; slab_counts/slab_displs: what it receives into (sends out of)
// the exchange buffer on the slab side, ordered by peer rank.  Arrays of nranks entries each.
int b200ls_repart_counts(const b200ls_repart *p, int64_t *box_counts, int64_t *box_displs, int64_t *slab_counts,
                         int64_t *slab_displs)
{
    if (!p) return B200LS_ERR_ARG;
    int64_t lo[3], hi[3];
    p->box_of(p->rank, lo, hi);
    const int64_t area = (hi[0] - lo[0]) * (hi[1] - lo[1]);
    int64_t run = 0;
    for (int s = 0; s < p->nranks; ++s)
    {
        int64_t k0, k1;
        p->overlap(p->rank, s, k0, k1);
        if (box_counts) box_counts[s] = area * (k1 - k0);
        if (box_displs) box_displs[s] = (k1 > k0) ? area * (k0 - lo[2]) : 0;
    }
    for (int q = 0; q < p->nranks; ++q)
    {
        int64_t qlo[3], qhi[3], k0, k1;
        p->box_of(q, qlo, qhi);
        p->overlap(q, p->rank, k0, k1);
        const int64_t c = (qhi[0] - qlo[0]) * (qhi[1] - qlo[1]) * (k1 - k0);
        if (slab_counts) slab_counts[q] = c;
        if (slab_displs) slab_displs[q] = run;
        run += c;
    }
    return B200LS_OK;
}

int b200ls_repart_unpack_slab(const b200ls_repart *p, const double *recvbuf, double *slab)
{
    if (!p || !recvbuf || !slab) return B200LS_ERR_ARG;
    slab_copy<true>(p, const_cast<double *>(recvbuf), slab);
    return B200LS_OK;
}

int b200ls_repart_pack_slab(const b200ls_repart *p, const double *slab, double *sendbuf)
{
    if (!p || !slab || !sendbuf) return B200LS_ERR_ARG;
    slab_copy<false>(p, sendbuf, const_cast<double *>(slab));
    return B200LS_OK;
}

// natural index i + nx*(j + ny*k) of PETSc global indices (columns of the assembled matrix)
int b200ls_repart_petsc_to_natural(const b200ls_repart *p, int64_t count, const int32_t *petsc_idx, int32_t *natural_idx)
{
    if (!p || count < 0 || (count > 0 && (!petsc_idx || !natural_idx))) return B200LS_ERR_ARG;
    const int64_t nx = p->n[0], ny = p->n[1];
    const int64_t total = p->offset[(size_t)p->nranks];
    int q = p->rank;  // columns are mostly local: start the search at the own box
    int64_t lo[3], hi[3];
    p->box_of(q, lo, hi);
    for (int64_t t = 0; t < count; ++t)
    {
        const int64_t g = petsc_idx[t];
        if (g < 0 || g >= total) return B200LS_ERR_ARG;
        if (g < p->offset[(size_t)q] || g >= p->offset[(size_t)q + 1])
        {
            q = (int)(std::upper_bound(p->offset.begin(), p->offset.end(), g) - p->offset.begin()) - 1;
            p->box_of(q, lo, hi);
        }
        const int64_t l = g - p->offset[(size_t)q];
        const int64_t xm = hi[0] - lo[0], ym = hi[1] - lo[1];
        const int64_t i = lo[0] + l % xm, j = lo[1] + (l / xm) % ym, k = lo[2] + l / (xm * ym);
        natural_idx[t] = (int32_t)(i + nx * (j + ny * k));
    }
    return B200LS_OK;
}

// natural indices of this rank's own rows, in box order
int b200ls_repart_box_rows(const b200ls_repart *p, int64_t *natural_rows)
{
    if (!p || !natural_rows) return B200LS_ERR_ARG;
    int64_t lo[3], hi[3];
    p->box_of(p->rank, lo, hi);
    const int64_t nx = p->n[0], ny = p->n[1];
    int64_t t = 0;
    for (int64_t k = lo[2]; k < hi[2]; ++k)
        for (int64_t j = lo[1]; j < hi[1]; ++j)
            for (int64_t i = lo[0]; i < hi[0]; ++i) natural_rows[t++] = i + nx * (j + ny * k);
    return B200LS_OK;
}

// Process grids (m, n, p) of `nranks` ranks whose boxes have exactly the local sizes the ranks report
// (sizes[q] = rows rank q owns).  Writes up to cap triples, most PETSc-like first (fewest cuts of the fastest
// axis last is irrelevant: every candidate is verified against the matrix by the caller); *found = how many exist.
int b200ls_repart_candidates(int dim, const int64_t n[3], int nranks, const int64_t *sizes, int *procs_out, int cap, int *found)
{
    if (!n || !sizes || !found || nranks <= 0 || (dim != 2 && dim != 3) || (cap > 0 && !procs_out)) return B200LS_ERR_ARG;
    int cnt = 0;
    std::vector<int64_t> sx, sy, sz;
    for (int m = 1; m <= nranks; ++m)
    {
        if (nranks % m || (int64_t)m > n[0]) continue;
        const int rest = nranks / m;
        for (int nn = 1; nn <= rest; ++nn)
        {
            if (rest % nn || (int64_t)nn > n[1]) continue;
            const int pp = rest / nn;
            if (dim == 2 && pp != 1) continue;
            if (dim == 3 && (int64_t)pp > n[2]) continue;
            split_axis(n[0], m, sx);
            split_axis(n[1], nn, sy);
            split_axis(dim == 3 ? n[2] : 1, pp, sz);
            bool ok = true;
            for (int q = 0; q < nranks && ok; ++q)
            {
                const int px = q % m, py = (q / m) % nn, pz = q / (m * nn);
                const int64_t sz_q = (sx[(size_t)px + 1] - sx[(size_t)px]) * (sy[(size_t)py + 1] - sy[(size_t)py]) *
                                     (sz[(size_t)pz + 1] - sz[(size_t)pz]);
                ok = (sz_q == sizes[q]);
            }
            if (!ok) continue;
            if (cnt < cap)
            {
                procs_out[3 * cnt + 0] = m;
                procs_out[3 * cnt + 1] = nn;
                procs_out[3 * cnt + 2] = pp;
            }
            ++cnt;
        }
    }
    *found = cnt;
    return B200LS_OK;
}

}  // extern "C"
