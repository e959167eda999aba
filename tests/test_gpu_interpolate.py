"""interpolate_vertex! / interpolate_solver! (src/interpolate.jl:62-213) on the device vs the oracle: refining, coarsening,
odd / even meshes, smaller and larger frequency boxes."""
import numpy as np
import pytest

from test_gpu_parity import rel, TOL

pytestmark = pytest.mark.gpu


def _solver(fd, T, U, nK1, nK2, nK3, L, LG, nG, seedF=None, mu=0.2):
    hp = {"t1": 1.0, "t2": -0.3}
    Gb = fd.hubbard_bare_Green(T, nG, LG, μ=mu, **hp)
    G0 = fd.hubbard_bare_Green(T, nG, LG, μ=mu - 0.1, **hp)
    S = fd.NL2_ParquetSolver(nK1, nK2, nK3, L, Gb, G0, np.zeros_like(G0), fd.RefVertex(T, U), T=T)
    if seedF is not None:
        fd.randomize_vertex(S.F, seedF, 0.3)
        S.push("F")
    return S


@pytest.mark.parametrize("Li,Lo,boxes_i,boxes_o", [(3, 4, (6, (2, 2), (2, 2)), (6, (2, 2), (2, 2))),
                                                  (4, 6, (5, (2, 3), (2, 2)), (7, (3, 2), (2, 1))),
                                                  (6, 4, (7, (3, 3), (2, 2)), (6, (2, 3), (2, 3)))])
def test_interpolate_vertex_all_classes(orc, Li, Lo, boxes_i, boxes_o):
    import fddgasolver_jl_b200 as fd
    T, U = 0.5, 2.0
    Fi = fd.NL2_Vertex(fd.RefVertex(T, U), T, boxes_i[0], boxes_i[1], boxes_i[2], Li)
    fd.randomize_vertex(Fi, 3, 0.5)
    So = _solver(fd, T, U, boxes_o[0], boxes_o[1], boxes_o[2], Lo, 2 * Lo, boxes_o[0] + 2, seedF=4)
    fd.interpolate_vertex(So, Fi)
    So.pull("F")
    Fo = fd.NL2_Vertex(fd.RefVertex(T, U), T, boxes_o[0], boxes_o[1], boxes_o[2], Lo)
    orc.interpolate_vertex(Fo, Fi)
    for ch in range(3):
        for cls in ("K1", "K2", "K3"):
            a, b = getattr(So.F.channel(ch), cls), getattr(Fo.channel(ch), cls)
            assert rel(a, b) < TOL, (ch, cls, rel(a, b))
    So.close()


def test_interpolate_solver(orc):
    """interpolate_solver!(So, Si; occ_target, hubbard_params): Σ with edge clamping, μ search, Dyson, bubbles, vertex, symmetrisation"""
    import fddgasolver_jl_b200 as fd
    T, U = 0.5, 2.0
    hp = {"t1": 1.0, "t2": -0.3}
    Si = _solver(fd, T, U, 6, (2, 2), (2, 2), 3, 6, 6, seedF=7)
    Si.init_sym_grp()
    fd.symmetrize_solver(Si)
    fd.SDE(Si, "scPA")
    Si.pull("F", "Σ")
    So = _solver(fd, T, U, 8, (3, 2), (2, 2), 4, 8, 9)
    So.init_sym_grp()
    Ro = orc.OracleSolver(8, (3, 2), (2, 2), 4, So.Gbare, So.G0, So.Σ0, fd.RefVertex(T, U), T=T)
    Ro.init_sym_grp()

    class RiView:
        pass
    Ri = RiView()
    Ri.Σ, Ri.F, Ri.LG, Ri.nG = Si.Σ, Si.F, Si.LG, Si.nG
    fd.interpolate_solver(So, Si, occ_target=0.46, hubbard_params=hp)
    orc.interpolate_solver(Ro, Ri, occ_target=0.46, hubbard_params=hp)
    So.pull("F", "Σ", "G", "Gbare", "Π")
    assert rel(So.Σ, Ro.Σ) < TOL and rel(So.Gbare, Ro.Gbare) < 1e-9 and rel(So.G, Ro.G) < 1e-9
    assert rel(So.Πpp, Ro.Πpp) < 1e-9 and rel(So.Πph, Ro.Πph) < 1e-9
    for a, b in zip(So.F.channels(), Ro.F.channels()):
        for x, y in zip(a.arrays(), b.arrays()):
            assert rel(x, y) < TOL
    assert abs(fd.compute_occupation(So) - 0.46) < 1e-10
    Si.close(); So.close()
