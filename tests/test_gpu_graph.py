"""CUDA-graph record / replay of library call sequences (include/fdga.h: fdga_graph_*): a replay leaves exactly the state the eager
calls leave, and the guards refuse recordings that are not steady-state cycles, host-synchronising calls and stale graphs."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _solver(nl_method):
    import fddgasolver_jl_b200 as fd
    S = fd.wu_point_solver(nmax=2, nq=4, LG=8, small_reference=True, F0_scale=0.03, F_scale=0.2, nl_method=nl_method)
    x = S.F.flatten()
    S.unflatten_F(x); S.stash_F()
    return S


@pytest.mark.parametrize("nl_method,strategy", [(2, "fdPA"), (2, "scPA"), (1, "fdPA"), (2, "fdPA_new")])
def test_replay_is_bit_identical_to_eager_issue(nl_method, strategy):
    import fddgasolver_jl_b200 as fd
    S = _solver(nl_method)

    def step():
        S.unstash_F()
        fd.iterate_solver(S, strategy, update_Σ=False)
        fd.SDE(S, "scPA")
    step(); step(); S.sync()
    ya = S.flatten_F().copy(); S.pull("Σ"); sa = S.Σ.copy()
    n0 = S.total_launches()
    gid = S.record(step)
    assert S.total_launches() == n0                      # recording executes nothing
    S.Σ[...] = 0; S.push("Σ")
    for _ in range(3):
        S.replay(gid)
    S.sync()
    per_step = (S.total_launches() - n0) // 3
    assert per_step > 20
    S.pull("Σ")
    assert np.array_equal(S.flatten_F(), ya) and np.array_equal(S.Σ, sa)
    S.drop_graph(gid)
    with pytest.raises(fd.FdgaError, match="bad graph id"):
        S.replay(gid)
    S.close()


def test_full_iteration_with_self_energy_update_replays():
    """iterate_solver!(fdPA) with Dyson / bubbles / SDE inside the graph: a replayed fixed-point iteration continues the eager one"""
    import fddgasolver_jl_b200 as fd
    A, B = _solver(2), _solver(2)
    it = lambda S: fd.iterate_solver(S, "fdPA")
    for S in (A, B):
        it(S); it(S)
    gid = B.record(lambda: it(B))
    for _ in range(3):
        it(A); B.replay(gid)
    A.sync(); B.sync()
    A.pull("Σ", "G"); B.pull("Σ", "G")
    assert np.array_equal(A.flatten_F(), B.flatten_F()) and np.array_equal(A.Σ, B.Σ) and np.array_equal(A.G, B.G)
    A.close(); B.close()


def test_guards():
    import fddgasolver_jl_b200 as fd
    S = _solver(2)
    step = lambda: (S.unstash_F(), fd.iterate_solver(S, "fdPA", update_Σ=False), fd.SDE(S, "scPA"))
    step(); step()
    # a host synchronisation inside a recording is refused and the recording is abandoned
    with pytest.raises(fd.FdgaError, match="not inside a graph recording"):
        S.record(lambda: (step(), S.sync()))
    step(); S.sync()
    # not a cycle: unstash alone leaves the derived tables of S.F dirty while the recording started with them current
    fd.build_K3_cache(S)
    with pytest.raises(fd.FdgaError, match="steady-state cycle"):
        S.record(S.unstash_F)
    step(); step(); S.sync()
    gid = S.record(step)
    S.replay(gid); S.sync()
    fd.build_K3_cache(S)                                # refreshes the s-wave tables: not the recorded start state any more
    with pytest.raises(fd.FdgaError, match="record it again"):
        S.replay(gid)
    step(); S.replay(gid); S.sync()                     # back in the cycle
    S.init_sym_grp()                                    # rebuilds device tables: the graph is stale for good
    step()
    with pytest.raises(fd.FdgaError, match="stale graph"):
        S.replay(gid)
    with pytest.raises(fd.FdgaError, match="profiling"):
        S.profile(True); S.record(step)
    S.profile(False)
    S.close()


@pytest.mark.parametrize("nl_method", [2, 1])
def test_automatic_graphs_behind_the_plain_entry_points(nl_method):
    """fdga_iterate_solver / fdga_sde / the mfRG matvec record themselves after two calls from the same lazy state and replay from
    then on; a solver that is forced to issue eagerly (profiling on) walks through bit-identical states"""
    import fddgasolver_jl_b200 as fd
    A, B = _solver(nl_method), _solver(nl_method)
    B.profile(True)                                     # eager launches only
    x = A.F.flatten()
    mapA, mapB = fd.mfRGLinearMap(A), fd.mfRGLinearMap(B)
    for it in range(6):
        for S in (A, B):
            fd.iterate_solver(S, "fdPA")
        if it % 2:
            for S in (A, B):
                fd.SDE(S, "scPA")                       # an extra call now and then: off-cycle states fall back to eager issue
    ya, yb = A.flatten_F().copy(), B.flatten_F().copy()
    A.pull("Σ", "G"); B.pull("Σ", "G")
    assert np.array_equal(ya, yb) and np.array_equal(A.Σ, B.Σ) and np.array_equal(A.G, B.G)
    for _ in range(4):
        xa, xb = mapA.matvec(x), mapB.matvec(x)
        assert np.array_equal(xa, xb)
        x = xa * 0.5
    # after a rebuild of device tables the recorded graphs are dropped, not replayed
    for S in (A, B):
        S.init_sym_grp()
        fd.iterate_solver(S, "fdPA"); fd.iterate_solver(S, "fdPA"); fd.iterate_solver(S, "fdPA")
    assert np.array_equal(A.flatten_F(), B.flatten_F())
    A.close(); B.close()


def test_automatic_graphs_follow_a_change_of_lazy_state():
    """after the bubbles switch from the product form to explicit arrays (solve_using_mfRG's mixing) the entry points are called from
    a different lazy state: a second graph is recorded for it, and the results still equal eager issue"""
    import fddgasolver_jl_b200 as fd
    A, B = _solver(2), _solver(2)
    B.profile(True)
    for S in (A, B):
        for _ in range(3):
            fd.iterate_solver(S, "fdPA", update_Σ=False)
        fd.mix_bubbles(S, 0.7)
        for _ in range(4):
            fd.iterate_solver(S, "fdPA", update_Σ=False)
            fd.SDE(S, "fdPA")
    A.pull("Σ"); B.pull("Σ")
    assert np.array_equal(A.flatten_F(), B.flatten_F()) and np.array_equal(A.Σ, B.Σ)
    A.close(); B.close()
