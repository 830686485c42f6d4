# dump_golden.jl -- write reference-held golden vectors for the ELBO hot path.
#
# Run by someone who HAS the reference toolchain (Julia 0.6 + Celeste.jl at commit 41c4897 with its REQUIRE
# packages and the SDSS field 3900/6/269 that test/SampleData.jl downloads):
#
#     cd Celeste.jl/test && julia ../../celeste.jl_b200/tools/julia/dump_golden.jl OUTDIR
#
# For each SampleData fixture it evaluates DeterministicVI.elbo_likelihood exactly as test/test_elbo.jl:18
# does (Sa = all sources, include_kl irrelevant: the likelihood only) in value / gradient / Hessian mode and
# writes ONE self-describing little-endian file  OUTDIR/<case>.celgold  holding the ElboArgs inputs flattened
# in the layout of include/celeste_cuda.h (the same arrays tests/golden_io.py stores) plus the reference's outputs.
# Commit the files under tests/golden/julia/: tests/test_golden.py::test_oracle_matches_julia_dumps then pins
# the oracle (and through it every CUDA parity test) to numbers the REFERENCE computed.
#
# File format: repeated records  [int32 name_len][name bytes][int32 dtype: 0=f64 1=f32 2=i64 3=u8][int32 ndim]
# [int64 dims...][raw column-major data].  Reader: tests/golden_io.load_julia_dump.
#
# Nothing here is executed by this repository (no Julia in the build image); it is the reference-side half of
# the parity pin, kept next to the Python half so the two cannot drift.

using Celeste: Model, DeterministicVI, SensitiveFloats
import Celeste.Model: ids, Image, ImagePatch
include(joinpath(Pkg.dir("Celeste"), "test", "SampleData.jl"))
using SampleData

const DT = Dict(Float64 => Int32(0), Float32 => Int32(1), Int64 => Int32(2), UInt8 => Int32(3))

function rec(io::IO, name::String, a::Array{T}) where T
    write(io, Int32(length(name))); write(io, name)
    write(io, DT[T]); write(io, Int32(ndims(a)))
    for d in size(a); write(io, Int64(d)); end
    write(io, a)
end
rec(io::IO, name::String, x::Real) = rec(io, name, [Float64(x)])

function dump_case(path::String, ea, vp)
    open(path, "w") do io
        N, S = ea.N, ea.S
        rec(io, "N", [Int64(N)]); rec(io, "S", [Int64(S)])
        rec(io, "active_sources", Int64.(ea.active_sources))
        for n in 1:N
            img = ea.images[n]
            rec(io, "img$(n)_meta", Int64[img.H, img.W, img.b])
            rec(io, "img$(n)_pixels", Array{Float32}(img.pixels))
            # sky may be a lazy SkyIntensity (sky_small x calibration): materialise what elbo_objective.jl:374 reads
            rec(io, "img$(n)_sky", Float32[img.sky[h, w] for h in 1:img.H, w in 1:img.W])
            rec(io, "img$(n)_iota", Array{Float32}(img.nelec_per_nmgy))
        end
        for s in 1:S, n in 1:N
            p = ea.patches[s, n]
            i = "p$(s)_$(n)"
            rec(io, i * "_offset", Int64[p.bitmap_offset[1], p.bitmap_offset[2]])
            rec(io, i * "_bitmap", UInt8.(p.active_pixel_bitmap))
            rec(io, i * "_wcs_jacobian", Array{Float64}(p.wcs_jacobian))
            rec(io, i * "_world_center", Array{Float64}(p.world_center))
            rec(io, i * "_pixel_center", Array{Float64}(p.pixel_center))
            psf = zeros(7, length(p.psf))          # alphaBar, xiBar[2], tauBar col-major (4): celeste_patch.psf
            for (k, pc) in enumerate(p.psf)
                psf[:, k] = [pc.alphaBar, pc.xiBar[1], pc.xiBar[2], pc.tauBar[1, 1], pc.tauBar[2, 1],
                             pc.tauBar[1, 2], pc.tauBar[2, 2]]
            end
            rec(io, i * "_psf", psf)
            rec(io, i * "_itp_coefs", Array{Float64}(p.itp_psf.coefs))    # padded coefficient array, fsm_util.jl:236
        end
        rec(io, "vp", hcat(vp...))                                         # 44 x S
        for (mode, grad, hess) in ((0, false, false), (1, true, false), (2, true, true))
            ev = DeterministicVI.ElboIntermediateVariables(Float64, ea.Sa, grad, hess)
            sf = DeterministicVI.elbo_likelihood(ea, vp, ev)               # test/test_elbo.jl:18
            rec(io, "out$(mode)_v", sf.v[])
            grad && rec(io, "out$(mode)_d", Array{Float64}(sf.d))
            hess && rec(io, "out$(mode)_h", Array{Float64}(sf.h))
            rec(io, "out$(mode)_counters", Int64[ev.active_pixel_counter, ev.inactive_pixel_counter])
        end
        # the full objective, KL included (elbo_objective.jl:481-497), pins rows f.3 / a12
        ea_kl = DeterministicVI.ElboArgs(ea.images, ea.patches, ea.active_sources; include_kl=true)
        sf = DeterministicVI.elbo(ea_kl, vp)
        rec(io, "elbo_kl_v", sf.v[]); rec(io, "elbo_kl_d", Array{Float64}(sf.d)); rec(io, "elbo_kl_h", Array{Float64}(sf.h))
    end
    println("wrote ", path)
end

outdir = length(ARGS) >= 1 ? ARGS[1] : "celgold"
mkpath(outdir)
for (name, gen) in (("star", SampleData.gen_sample_star_dataset), ("galaxy", SampleData.gen_sample_galaxy_dataset),
                    ("two_body", SampleData.gen_two_body_dataset), ("three_body", SampleData.gen_three_body_dataset))
    ea, vp, catalog = gen()
    dump_case(joinpath(outdir, name * ".celgold"), ea, vp)
    # the production shape: one active source, the others as neighbours (ParallelRun.jl:236-253)
    if ea.S > 1
        ea1 = DeterministicVI.ElboArgs(ea.images, ea.patches, [1]; include_kl=false)
        dump_case(joinpath(outdir, name * "_active1.celgold"), ea1, vp)
    end
end
