"""Runs the BODIES of the `-m gpu` tests that were written without a GPU at hand (tests/test_zzz_gpu_*.py) on the CPU
emulation of the kernel sources: tests/emu_backend.py stands in for the `petibm_b200` module.  What this proves is that
the tests themselves are executable and that their expectations hold for the kernel logic; the device run is still the
`-m gpu` suite.  Device-only cases (torch CUDA tensors, 128^3, direct C-ABI handles) are left out."""
import pytest

from tests import emu_backend
from tests import test_zzz_gpu_2_staggered as T2
from tests import test_zzz_gpu_3_ops as T3
from tests import test_zzz_gpu_4_multigrid as T4
from tests import test_zzz_gpu_3_direct as T5
from tests.test_emulated_kernels import emu  # noqa: F401  (fixture)


@pytest.fixture(scope="module")
def pbe(emu):  # noqa: F811
    return emu_backend.make_module(emu)


@pytest.mark.parametrize("shape,per", [((11, 9), (0, 0)), ((9, 8, 7), (1, 0, 1)), ((8, 7, 6), (0, 0, 0))])
def test_staggered_velocity_body(pbe, shape, per):
    T2.test_velocity_system_spmv_bit_exact_and_bcgs(pbe, shape, per)


@pytest.mark.parametrize("pc", ["none", "jacobi"])
def test_staggered_ibpm_like_body(pbe, pc):
    T2.test_ibpm_modified_poisson_stencil_block_plus_remainder(pbe, pc)


@pytest.mark.parametrize("shape,per,zchunk,stages", [((70, 19, 12), (0, 0, 0), 5, 3), ((9, 8, 7), (1, 0, 1), 0, 4), ((66, 17), (0, 1), 0, 3)])
def test_tiled_velocity_body(pbe, shape, per, zchunk, stages):
    T2.test_tiled_kernels_velocity_system(pbe, shape, per, zchunk, stages)


@pytest.mark.parametrize("pc,stages", [("none", 3), ("jacobi", 4)])
def test_tiled_ibpm_like_body(pbe, pc, stages):
    T2.test_tiled_kernels_ibpm_like_system_with_remainder(pbe, pc, stages)


def test_tiled_hybrid_body(pbe):
    T2.test_tiled_kernels_hybrid_operator_on_a_stretched_ibpm_system(pbe, 3)


def test_default_kernel_choice_body(pbe):
    T2.test_default_kernel_choice_of_the_line_coefficient_operator(pbe)


def test_staggered_fallback_body(pbe):
    T2.test_a_matrix_without_the_structure_keeps_the_csr_operator(pbe)


@pytest.mark.parametrize("dim", [2, 3])
def test_hybrid_body(pbe, dim):
    T2.test_hybrid_operator_on_a_stretched_ibpm_system(pbe, dim)


@pytest.mark.parametrize("shape,per", [((16, 12), (1, 1)), ((12, 9, 11), (1, 0, 1)), ((14, 10, 8), (0, 0, 0))])
def test_ops_body(pbe, shape, per):
    T3.test_operators_are_the_assembled_products(pbe, shape, per)


@pytest.mark.parametrize("shape,per", [((33, 40), (1, 1)), ((31, 17, 13), (0, 0, 0))])
def test_multigrid_pcg_body(pbe, shape, per):
    T4.test_pcg_with_multigrid_matches_the_restatement(pbe, shape, per)


def test_multigrid_needs_the_stencil_body(pbe):
    T4.test_multigrid_needs_the_separable_operator(pbe)


@pytest.mark.parametrize("dim", [2])
def test_block_multigrid_body(pbe, dim):
    T4.test_block_multigrid_on_a_stretched_ibpm_system(pbe, dim)


def test_direct_solve_body(pbe, tmp_path):
    T5.test_forces_system_direct_solve(pbe, tmp_path)
