"""fddgasolver.jl_b200 -- B200-native Bethe-Salpeter / K3-cache / bubble / SDE hot path of fdDGAsolver.jl.

The package holds the CUDA sources + C-ABI (csrc/, built into libfdga.so) and a thin host-side mirror of the
reference's NL2_ParquetSolver interface (solver.py).  Import it as ``fddgasolver_jl_b200`` (the directory
name contains a dot; the top-level shim ``fddgasolver_jl_b200.py`` registers it under that name).
"""
from ._lib import FdgaError, LIB_PATH, EXPORTS  # noqa: F401
from .types import (pCh, tCh, aCh, pSp, xSp, dSp, Channel, NL2_Channel, NL_Channel, RefVertex, Vertex, NL2_Vertex, NL_Vertex, MBEVertex, NL2_MBEVertex,  # noqa: F401
                    vertex_chain, nB, nF)
from .models import hubbard_bare_Green, hubbard_band, siam_bare_Green  # noqa: F401
from .solver import (NL2_ParquetSolver, NL_ParquetSolver, ParquetSolver, init_sym_grp, Dyson, compute_occupation, bubbles, bubbles_real_space,  # noqa: F401
                     bubbles_momentum_space, build_K3_cache, build_K3_cache_mfRG, BSE_L_K2, BSE_L_K3, BSE_K1,
                     BSE_K2, BSE_K3, BSE_K1_new, BSE_K2_new, BSE_K1_1loop, BSE_K2_1loop, BSE_K3_1loop, SDE, SDE_channel_L, iterate_solver, iterate_solver_stepwise, fixed_point, solve, mfRGLinearMap,
                     dqgmres, symmetrize_solver, fixed_point_preconditioned,
                     set_hubbard_bare_Green, compute_hubbard_chemical_potential, mix_bubbles, update_reference, solve_using_mfRG,
                     interpolate_vertex, interpolate_solver, save_solver, load_solver, real_array)
from . import h5min, io, synthetic, types  # noqa: F401,E402
from .io import load_triqs_data  # noqa: F401,E402
from .flow import bare_Green_Ω_flow  # noqa: F401,E402
from .mbe import asymptotic_to_mbe, mbe_to_asymptotic  # noqa: F401,E402
from .synthetic import (parquet_solver_hubbard_parquet_approximation_NL2, parquet_solver_hubbard_parquet_approximation, parquet_solver_siam_parquet_approximation, synthetic_local_vertex,  # noqa: F401
                        wu_point_solver, wu_point_inputs, randomize_vertex)
