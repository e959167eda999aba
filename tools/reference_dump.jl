# tools/reference_dump.jl -- run WITH the reference installed (Julia + fdDGAsolver.jl + MatsubaraFunctions.jl); NOT runnable in the
# build container (no Julia), delivered untested like the ccall shim of INTEGRATION.md.
#
#   julia --project=/path/to/fdDGAsolver.jl tools/reference_dump.jl tests/golden/julia
#
# Writes the inputs and the reference's outputs of one NL2 fdPA iteration (+ SDE, + one mfRG matvec) as raw little-endian
# ComplexF64 / Int64 / UInt8 files plus a manifest (name, eltype, dims per line).  tests/test_external_reference.py replays the
# inputs through libfdga and the CPU oracle and compares: this pins everything DESIGN.md lists as "parity unpinned"
# (class order of SymmetryGroup, bubbles_real_space! for L < LG, the mfRG branches).
using fdDGAsolver, MatsubaraFunctions, StaticArrays, Random
import fdDGAsolver: pCh, tCh, aCh

outdir = length(ARGS) >= 1 ? ARGS[1] : "tests/golden/julia"
mkpath(outdir)
manifest = open(joinpath(outdir, "manifest.txt"), "w")
function dump(name, a::AbstractArray)
    open(joinpath(outdir, name * ".bin"), "w") do io; write(io, Array(a)); end
    println(manifest, name, " ", eltype(a), " ", join(size(a), " "))
end
dump_vertex(prefix, F) = for (c, γ) in (("p", F.γp), ("t", F.γt), ("a", F.γa)), (k, f) in (("K1", γ.K1), ("K2", γ.K2), ("K3", γ.K3))
    dump("$(prefix)_$(c)_$(k)", f.data)
end
function dump_group(name, SG)
    offs = Int64[0]; idx = Int64[]; ops = UInt8[]
    for cl in SG.classes
        for (i, op) in cl; push!(idx, i - 1); push!(ops, UInt8(op.sgn) | (UInt8(op.con) << 1)); end
        push!(offs, length(idx))
    end
    dump(name * "_offsets", offs); dump(name * "_index", idx); dump(name * "_ops", ops)
end

T, U, μ, t1, t2 = 0.5, 2.0, 0.3, 1.0, -0.2
nmax = 2; nG = 4nmax; nK1 = 4nmax; nK2 = (nmax, nmax); nK3 = (nmax, nmax)
k1 = 2pi * SVector(1., 0.); k2 = 2pi * SVector(0., 1.)
mK_G = BrillouinZoneMesh(BrillouinZone(6, k1, k2)); mK_Γ = BrillouinZoneMesh(BrillouinZone(3, k1, k2))
S = parquet_solver_hubbard_parquet_approximation_NL2(nG, nK1, nK2, nK3, mK_G, mK_Γ; T, U, μ, t1, t2)
fdDGAsolver.init_sym_grp!(S)
Random.seed!(1)
unflatten!(S.F, 0.3 .* (rand(ComplexF64, length(flatten(S.F))) .- (0.5 + 0.5im)))
fdDGAsolver.symmetrize_solver!(S)

println(manifest, "# scalars T U nG nK1 nK2 nK3 L LG = ", join((T, U, nG, nK1, nK2..., nK3..., 3, 6), " "))
dump("in_Gbare", S.Gbare.data); dump("in_G0", S.G0.data); dump("in_Sigma0", S.Σ0.data)
dump_vertex("in_F", S.F)
for (n, SG) in (("SGsigma", S.SGΣ), ("SGK1", S.SGpp[1]), ("SGpp2", S.SGpp[2]), ("SGph2", S.SGph[2]), ("SGpp3", S.SGpp[3]),
                ("SGph3", S.SGph[3]), ("SGppL3", S.SGppL[3]), ("SGphL3", S.SGphL[3]))
    dump_group(n, SG)
end
dump("out_Pi0pp", S.Π0pp.data); dump("out_Pipp", S.Πpp.data); dump("out_Piph", S.Πph.data); dump("out_G", S.G.data)

iterate_solver!(S; strategy = :fdPA, update_Σ = true)
dump_vertex("out_F", S.F); dump_vertex("out_FL", S.FL); dump("out_Sigma", S.Σ.data)

A = fdDGAsolver.mfRGLinearMap(S, :fdPA)
x = flatten(S.F) .* 3
dump("mfrg_x", x); dump("mfrg_y", A * x)
close(manifest)
println("wrote ", outdir)
