/* poisson_cabi.c -- the C ABI of libb200ls.so from plain C: the calls a PetIBM-side binding makes for the pressure system
 * (INTEGRATION.md): options in PETSc's options-file syntax, the mesh's cell widths, solve with host buffers, iterations and
 * residual back.  Build:  gcc -std=c99 -I include examples/poisson_cabi.c -L petibm_b200 -lb200ls -Wl,-rpath,$PWD/petibm_b200 -lm
 * Needs a CUDA device at run time (there is no CPU path). */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "b200ls.h"

#define CHECK(call)                                                                                     \
    do                                                                                                  \
    {                                                                                                   \
        int rc_ = (call);                                                                               \
        if (rc_ != B200LS_OK)                                                                           \
        {                                                                                               \
            fprintf(stderr, "%s -> %s (%s)\n", #call, b200ls_error_string(rc_), b200ls_last_error(h)); \
            return 1;                                                                                   \
        }                                                                                               \
    } while (0)

int main(int argc, char **argv)
{
    const int n1 = argc > 1 ? atoi(argv[1]) : 64;
    b200ls_solver *h = NULL;
    b200ls_options opts;
    char err[256];
    const char *options_text = "-poisson_ksp_type cg\n-poisson_pc_type mg\n-poisson_ksp_rtol 1e-8\n-poisson_ksp_atol 1e-50\n";
    int64_t n[3] = {n1, n1, n1}, i, nvel = 0, np = 0;
    int periodic[3] = {0, 0, 0}, its = 0, reason = 0;
    double *w, *u, *rhs, *dp, *pr, rnorm = 0.0, div0 = 0.0, div1 = 0.0;

    /* axis widths from PetIBM's sub-domain description: one uniform sub-domain [0, 1] */
    double end = 1.0, ratio = 1.0;
    int cells = n1, nw = 0;
    w = (double *)malloc(sizeof(double) * (size_t)n1);
    if (b200ls_axis_from_subdomains(0.0, 1, &end, &cells, &ratio, w, n1, &nw) != B200LS_OK || nw != n1) return 1;

    if (b200ls_create(&h, 0) != B200LS_OK)
    {
        fprintf(stderr, "no CUDA device: this backend has no CPU path\n");
        return 2;
    }
    b200ls_default_options(&opts);
    if (b200ls_parse_options(options_text, "poisson_", &opts, err, sizeof err) != B200LS_OK)
    {
        fprintf(stderr, "%s\n", err);
        return 1;
    }
    CHECK(b200ls_set_options(h, &opts));
    CHECK(b200ls_set_poisson_stencil(h, 3, n, periodic, w, w, w, 0.01, 0, n[2]));
    CHECK(b200ls_set_nullspace(h, 1, 0, NULL));   /* MatNullSpaceCreate(comm, PETSC_TRUE, 0, ...) on DBNG */
    CHECK(b200ls_velocity_size(h, &nvel, &np));

    /* one projection step: rhs = D u*, solve for dp, u = u* - BNG dp */
    u = (double *)malloc(sizeof(double) * (size_t)nvel);
    rhs = (double *)malloc(sizeof(double) * (size_t)np);
    dp = (double *)malloc(sizeof(double) * (size_t)np);
    pr = (double *)calloc((size_t)np, sizeof(double));
    srand(1);
    for (i = 0; i < nvel; ++i) u[i] = (double)rand() / RAND_MAX - 0.5;
    CHECK(b200ls_divergence(h, u, rhs));
    for (i = 0; i < np; ++i) div0 += rhs[i] * rhs[i];
    CHECK(b200ls_solve(h, rhs, dp));
    CHECK(b200ls_get_iters(h, &its));
    CHECK(b200ls_get_residual(h, &rnorm));
    CHECK(b200ls_get_reason(h, &reason));
    CHECK(b200ls_project(h, u, pr, dp));   /* u -= BNG dp; p += dp */
    CHECK(b200ls_divergence(h, u, rhs));
    for (i = 0; i < np; ++i) div1 += rhs[i] * rhs[i];
    printf("%d^3: %d CG iterations (reason %d, residual %.3e); |D u| %.3e -> %.3e\n", n1, its, reason, rnorm, sqrt(div0), sqrt(div1));
    b200ls_destroy(h);
    free(w); free(u); free(rhs); free(dp); free(pr);
    return 0;
}
