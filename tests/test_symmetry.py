"""Symmetry-class construction (integer, bit-exact): the library's host builder (csrc/fdga_symmetry.cpp) against the
oracle's independent builder, plus structural properties of SymmetryGroup(symmetries, f)."""
import numpy as np
import pytest

CASES = [  # (which, n0, n1, nq)
    (0, 5, 0, 4), (0, 4, 0, 3), (1, 6, 0, 4), (1, 6, 0, 3),
    (2, 3, 2, 3), (3, 3, 2, 3), (2, 2, 3, 4), (3, 2, 3, 4),
    (4, 2, 2, 3), (5, 2, 2, 3), (6, 2, 2, 4), (7, 2, 2, 4), (4, 3, 2, 4), (5, 3, 2, 4),
]


def length(which, n0, n1, nq):
    NP = nq * nq
    if which == 0:
        return 2 * n0 * NP
    if which == 1:
        return (2 * n0 - 1) * NP
    if which in (2, 3):
        return (2 * n0 - 1) * 2 * n1 * NP * NP
    return (2 * n0 - 1) * (2 * n1) ** 2 * NP


@pytest.mark.parametrize("which,n0,n1,nq", CASES)
def test_builder_matches_oracle_bit_exact(orc, which, n0, n1, nq):
    import fddgasolver_jl_b200 as fd
    n = length(which, n0, n1, nq)
    a = fd._lib.build_symmetry_group(which, n0, n1, nq, n)
    b = orc.build_symmetry_group(which, n0, n1, nq, n)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


@pytest.mark.parametrize("which,n0,n1,nq", CASES)
def test_classes_partition_and_representative(orc, which, n0, n1, nq):
    import fddgasolver_jl_b200 as fd
    n = length(which, n0, n1, nq)
    offsets, index, ops = fd._lib.build_symmetry_group(which, n0, n1, nq, n)
    assert offsets[0] == 0 and offsets[-1] == n
    assert np.array_equal(np.sort(index), np.arange(n))          # every element in exactly one class
    reps = index[offsets[:-1]]
    assert np.all(np.diff(reps) > 0)                             # classes discovered in ascending order
    for c in range(len(offsets) - 1):
        mem = index[offsets[c]:offsets[c + 1]]
        assert mem[0] == mem.min()                               # representative = lowest linear index
        assert ops[offsets[c]] == 0                              # identity on the representative
    assert len(offsets) - 1 < n                                  # the group is not trivial


def test_bare_green_is_symmetric(orc):
    """test/test_nonlocal_symmetry.jl:22-47: G of the Hubbard model obeys SGΣ exactly; class fill reproduces G"""
    import ctypes as C
    T, nG, LG = 0.5, 5, 4
    G = orc.hubbard_bare_Green(T, nG, LG, μ=0.3, t1=1.0, t2=0.2, t3=-0.4)
    tbl = orc.build_symmetry_group(0, nG, 0, LG, G.size)
    Gs = G.copy(order="F")
    orc.lib().orc_symmetrize(orc._p(Gs), C.byref(orc.sg_struct(tbl)))
    assert np.max(np.abs(Gs - G)) < 1e-14
    Gfill = np.zeros_like(G)
    flat, src = Gfill.reshape(-1, order="F"), G.reshape(-1, order="F")
    reps = tbl[1][tbl[0][:-1]]
    flat[reps] = src[reps]
    Gfill = np.asfortranarray(flat.reshape(G.shape, order="F"))
    orc.lib().orc_symmetrize(orc._p(Gfill), C.byref(orc.sg_struct(tbl)))
    assert np.max(np.abs(Gfill - G)) < 1e-14
