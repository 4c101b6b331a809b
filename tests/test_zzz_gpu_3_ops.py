"""GPU checks of the matrix-free operators on either side of the pressure solve (b200ls_divergence / _gradient /
_project; ops_kernels.cuh): bit-identical to MatMult on the oracle's assembled D, G and MatMatMult(BN, G), and the
device-resident chain rhs2 = D u* -> solve -> u -= BNG dP leaves a divergence-free field.  The same kernel sources
run on the CPU emulation in tests/test_emulated_kernels.py."""
import numpy as np
import pytest

from oracle import oracle as orc
from tests import helpers as H

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]


@pytest.fixture(scope="module")
def pb():
    import petibm_b200

    return petibm_b200


@pytest.mark.parametrize("shape,per", [((33, 20), (0, 0)), ((16, 12), (1, 1)), ((20, 14, 10), (0, 0, 0)), ((12, 9, 11), (1, 0, 1)),
                                       ((70, 5, 4), (0, 1, 0))])
def test_operators_are_the_assembled_products(pb, shape, per):
    widths = H.make_widths(shape)
    per3 = list(per) + [0] * (3 - len(shape))
    D = orc.assemble_divergence(widths, per3)
    G = orc.assemble_gradient(widths, per3)
    BNG = orc.matmatmult(orc.bnhead_order1(G.shape[0], 0.01), G)
    s = pb.LinSolverB200("poisson", "None")
    s.setStencil(H.grid_of(widths, per))
    assert s.velocitySize() == G.shape
    rng = np.random.default_rng(5)
    u, p, dp = rng.standard_normal(G.shape[0]), rng.standard_normal(G.shape[1]), rng.standard_normal(G.shape[1])
    assert np.array_equal(s.divergence(u), D.spmv(u))
    assert np.array_equal(s.gradient(p), G.spmv(p))
    assert np.array_equal(s.gradient(p, with_bn=True), BNG.spmv(p))
    u2, p2 = u.copy(), p.copy()
    s.project(u2, p2, dp)
    assert np.array_equal(u2, u + (-1.0) * BNG.spmv(dp)) and np.array_equal(p2, p + 1.0 * dp)
    s.destroy()


@pytest.mark.parametrize("shape,per", [((33, 20), (0, 0)), ((16, 12), (1, 1)), ((20, 14, 10), (0, 0, 0)), ((12, 9, 11), (1, 0, 1)),
                                       ((70, 5, 4), (0, 1, 0))])
def test_convection_is_the_reference_stencil(pb, shape, per):
    """b200ls_convection (k_convection) against the oracle's restatement of createconvection.cpp:39-332: bit-identical,
    host buffers and device-resident (ghosted arrays filled on the device from the packed vector + caller's ghost values)."""
    import torch

    widths = H.make_widths(shape)
    s = pb.LinSolverB200("poisson", "None")
    s.setStencil(H.grid_of(widths, per))
    q, packed, nf = H.ghosted_fields(shape, per)
    ref = np.concatenate([o.ravel() for o in orc.convection(widths, per, q)])
    assert s.ghostedSizes()[: len(shape)] == [a.size for a in q]
    assert np.array_equal(s.convection(q), ref)
    dev = torch.device("cuda", 0)
    qd = [torch.from_numpy(a.ravel().copy()).to(dev) for a in q]
    assert np.array_equal(s.convection(qd).cpu().numpy(), ref)
    # interior + periodic wrap layers refreshed on the device from a new packed vector, ghost values kept
    rng = np.random.default_rng(8)
    packed2 = rng.standard_normal(packed.size)
    s.ghostedFromPacked(torch.from_numpy(packed2).to(dev), qd)
    q2, off = [], 0
    dim = len(shape)
    for f in range(dim):
        a = q[f].copy()
        m = int(np.prod(nf[f]))
        a[tuple([slice(1, -1)] * dim)] = packed2[off: off + m].reshape(tuple(reversed(nf[f])))
        off += m
        for d in range(dim):
            if per[d]:
                ax = dim - 1 - d
                lo = [slice(None)] * dim; hi = [slice(None)] * dim; first = [slice(None)] * dim; last = [slice(None)] * dim
                lo[ax], hi[ax], first[ax], last[ax] = 0, -1, 1, -2
                a[tuple(lo)] = a[tuple(last)]
                a[tuple(hi)] = a[tuple(first)]
        q2.append(a)
    ref2 = np.concatenate([o.ravel() for o in orc.convection(widths, per, q2)])
    assert np.array_equal(s.convection(qd).cpu().numpy(), ref2)
    s.destroy()


def test_device_resident_projection_step(pb):
    """One fractional-step projection with everything on the device: rhs2 = D u*, CG for dP, u = u* - BNG dP.  The
    projected field is divergence-free to the tolerance of the solve."""
    import torch

    shape, per = (48, 40, 32), (0, 0, 0)
    widths = H.make_widths(shape)
    s = pb.LinSolverB200("poisson", "None")
    s.setOptions(pc_type="jacobi", rtol=1e-10, atol=1e-50, max_it=5000)
    s.setStencil(H.grid_of(widths, per))
    s.setNullSpace(True)
    nv, npr = s.velocitySize()
    rng = np.random.default_rng(2)
    dev = torch.device("cuda", 0)
    u = torch.from_numpy(rng.standard_normal(nv)).to(dev)
    p = torch.zeros(npr, dtype=torch.float64, device=dev)
    rhs = s.divergence(u)
    d0 = float(torch.linalg.norm(rhs))
    dp = torch.empty_like(rhs)
    s.solve(dp, rhs)
    assert s.getReason() == 2
    s.project(u, p, dp)
    torch.cuda.synchronize()
    d1 = float(torch.linalg.norm(s.divergence(u)))
    assert d1 <= 1e-7 * d0, (d0, d1)
    assert torch.equal(p, dp)
    s.destroy()
