// csr_kernels.cuh -- general assembled-operator path (CSR SpMV + fused Krylov vector kernels).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200 {

struct CsrDev
{
    int64_t nrows;
    const int64_t *rowptr;
    const int32_t *col;
    const double *val;
};

}  // namespace b200
