"""Product containers (fddgasolver.jl_b200/types.py) against the oracle's independent restatement of the reference's data
model (oracle/otypes.py): mesh lengths, array shapes, flatten order [γp; γt; γa] x [K1; K2; K3] in column-major element order
(src/channel.jl:155-176, src/vertex.jl:153-167).  CPU only."""
import numpy as np

import fddgasolver_jl_b200 as fd
import otypes as ot


def test_mesh_lengths_and_shapes():
    for N in (1, 2, 5, 16):
        assert fd.types.nB(N) == ot.n_boson(N) == 2 * N - 1
        assert fd.types.nF(N) == ot.n_fermion(N) == 2 * N
    a = fd.NL2_Vertex(fd.RefVertex(0.3, 1.0), 0.3, 5, (3, 2), (2, 1), 3)
    b = ot.ONL2_Vertex(ot.ORefVertex(0.3, 1.0), 0.3, 5, (3, 2), (2, 1), 3)
    for ga, gb in zip(a.channels(), b.channels()):
        for x, y in zip(ga.arrays(), gb.arrays()):
            assert x.shape == y.shape
    assert len(a) == len(b)


def test_flatten_order_matches_the_independent_restatement():
    T = 0.3
    loc = fd.Vertex(fd.RefVertex(T, 1.0), T, 6, (4, 3), (2, 2))
    a = fd.NL2_Vertex(loc, T, 5, (3, 2), (2, 1), 3)
    fd.randomize_vertex(a, 7, 1.0)
    fd.randomize_vertex(loc, 8, 1.0)
    b = ot.adopt(a)
    assert isinstance(b, ot.ONL2_Vertex) and isinstance(b.F0, ot.OVertex) and isinstance(b.F0.F0, ot.ORefVertex)
    xa, xb = a.flatten(), b.flatten()
    assert np.array_equal(xa, xb)
    assert np.array_equal(loc.flatten(), b.F0.flatten())
    # element by element: position of K2[iW, iv, iP, ik] of the t channel in the flattened vector
    g = b.γt
    nK1, nK2 = g.K1.size, g.K2.size
    off_t = len(b.γp)
    iW, iv, iP, ik = 2, 1, 4, 7
    pos = off_t + nK1 + iW + g.K2.shape[0] * (iv + g.K2.shape[1] * (iP + g.K2.shape[2] * ik))
    assert xa[pos] == a.γt.K2[iW, iv, iP, ik]
    # unflatten round trip through the other implementation
    rng = np.random.default_rng(0)
    y = rng.standard_normal(xa.size) + 1j * rng.standard_normal(xa.size)
    a.unflatten(y); b.unflatten(y)
    for ga, gb in zip(a.channels(), b.channels()):
        for u, v in zip(ga.arrays(), gb.arrays()):
            assert np.array_equal(u, v)
    assert np.array_equal(b.flatten(), y)


def test_swave_vertex_shapes_and_flatten_order():
    """NL_Vertex (src/nonlocal/vertex.jl, channel.jl:3-51): K2[Ω, ν, P]; same flatten order; test/test_nonlocal_solver.jl:39-45"""
    T = 0.3
    a = fd.NL_Vertex(fd.Vertex(fd.RefVertex(T, 1.0), T, 6, (4, 3), (2, 2)), T, 5, (3, 2), (2, 1), 3)
    assert a.γp.K1.shape == (9, 9) and a.γp.K2.shape == (5, 4, 9) and a.γp.K3.shape == (3, 2, 2, 9)
    fd.randomize_vertex(a, 11, 1.0)
    b = ot.adopt(a)
    assert isinstance(b, ot.ONL_Vertex) and isinstance(b.F0, ot.OVertex)
    x = a.flatten()
    assert np.array_equal(x, b.flatten()) and len(a) == len(b) == x.size
    g = a.γa
    iW, iv, iP = 3, 2, 5
    pos = 2 * len(a.γp) + g.K1.size + iW + g.K2.shape[0] * (iv + g.K2.shape[1] * iP)
    assert x[pos] == g.K2[iW, iv, iP]
    c = fd.NL_Vertex(a.F0, T, 5, (3, 2), (2, 1), 3)
    c.unflatten(x)
    assert np.array_equal(c.γa.K3, a.γa.K3) and np.array_equal(c.flatten(), x)
