# TensorQECCUDA.jl -- reference-side binding of libtqec_cuda.so (C ABI: include/tqec.h).
#
# NOT EXECUTED in this repository's CI: the build image and the GPU boxes have no Julia toolchain (SURVEY F4).  The
# same ABI is exercised call for call from Python (tensorqec.jl_b200/_cabi.py, tests/test_gpu_parity.py).  This file
# shows what a TensorQEC.jl maintainer adds: new `CompiledDecoder` subtypes + `compile` / `decode` methods that keep
# `TNMAP` / `TNMMAP`, `compile(decoder, problem)` and `decode(compiled, syndrome)` unchanged
# (src/decoding/interfaces.jl:67-79, 96-119; src/decoding/tndecoder.jl:9-11, 77-80).
module TensorQECCUDA

using TensorQEC
using TensorQEC: Mod2, SimpleTannerGraph, CSSTannerGraph, SimpleSyndrome, CSSSyndrome, CSSErrorPattern,
                 GeneralDecodingProblem, IndependentDepolarizingDecodingProblem, DecodingResult, CompiledDecoder,
                 TNMAP, TNMMAP, reduce2general, nq, ns
import TensorQEC: compile, decode

const LIB = get(ENV, "TQEC_CUDA_LIB", "libtqec_cuda.so")

struct TqecError <: Exception
    code::Cint
    msg::String
end
check(rc::Cint) = rc == 0 ? nothing : throw(TqecError(rc, unsafe_string(ccall((:tqec_last_error, LIB), Cstring, ()))))

# mirror of `tqec_sweep_desc` (include/tqec.h): optional second lowering of a max-plus plan (in-place patch sweep,
# tensorqec.jl_b200/sweep.py:lower_sweep); pass C_NULL in PlanDesc.sweep to run the general kernels
struct SweepDesc
    W::Int32; sg::Int32; n_ss::Int32; n_head_bits::Int32; bp_words::Int32; n_tvals::Int32
    rec::Ptr{Int32}; tb::Ptr{Int32}; lanetab::Ptr{UInt32}; tvals::Ptr{Float64}
    head_bits::Ptr{Int32}; head_state::Ptr{Float64}; head_cfg::Ptr{UInt64}; out_index::Ptr{Int32}
end

# mirror of `tqec_plan_desc` (include/tqec.h)
struct PlanDesc
    semiring::Int32; n_vars::Int32; n_checks::Int32; n_obs::Int32; n_steps::Int32; w_max::Int32
    hdr::Ptr{Int32}; ints::Ptr{Int32}; n_ints::Int64
    tables::Ptr{Float64}; n_tables::Int64
    obs_slot::Ptr{Int32}; device::Int32
    sweep::Ptr{SweepDesc}
    table_bits::Int32          # plans with n_checks <= table_bits are fully tabulated at creation (0 = default 16)
end

mutable struct Plan
    h::Ptr{Cvoid}
    nsw::Int; ncw::Int; n_obs::Int
    function Plan(sch, device::Integer)          # `sch`: the lowered schedule (see `lower` below)
        href = Ref{Ptr{Cvoid}}(C_NULL)
        GC.@preserve sch begin
            sw = sch.sweep                         # `nothing`, or the tables of `lower_sweep` (all plain Vectors)
            swref = sw === nothing ? nothing :
                Ref(SweepDesc(sw.W, sw.sg, sw.n_ss, length(sw.head_bits), sw.bp_words, length(sw.tvals),
                              pointer(sw.rec), pointer(sw.tb), pointer(sw.lanetab), pointer(sw.tvals),
                              pointer(sw.head_bits), pointer(sw.head_state), pointer(sw.head_cfg), pointer(sw.out_index)))
            GC.@preserve sw swref begin
                swp = swref === nothing ? Ptr{SweepDesc}(C_NULL) : Base.unsafe_convert(Ptr{SweepDesc}, swref)
                d = PlanDesc(sch.semiring, sch.n_vars, sch.n_checks, sch.n_obs, sch.n_steps, sch.w_max,
                             pointer(sch.hdr), pointer(sch.ints), length(sch.ints),
                             pointer(sch.tables), length(sch.tables), pointer(sch.obs_slot), device, swp, Int32(0))
                check(ccall((:tqec_plan_create, LIB), Cint, (Ref{PlanDesc}, Ref{Ptr{Cvoid}}), d, href))
            end
        end
        p = new(href[], cld(max(sch.n_checks, 1), 64), cld(max(sch.n_vars, 1), 64), sch.n_obs)
        finalizer(x -> ccall((:tqec_plan_destroy, LIB), Cint, (Ptr{Cvoid},), x.h), p)
    end
end

# Vector{Mod2} columns -> shot-major UInt64 words: this IS `compresscol` (src/codes/mod2.jl:58-71) applied to the
# (bits x shots) matrix, so the reference's own helper produces the ABI layout.
pack(bits::AbstractMatrix{Mod2}) = TensorQEC.compresscol(bits)            # (ceil(nbits/64), shots)
function unpack(words::Matrix{UInt64}, nbits::Int)
    [Mod2((words[(i - 1) >> 6 + 1, s] >> ((i - 1) & 63)) & 1 == 1) for i in 1:nbits, s in axes(words, 2)]
end

# ---- lowering -----------------------------------------------------------------------------------------------------
# The Julia side lowers the factor graph it already builds (tndecoder.jl:33-40, 97-146, 186-219) to the flat schedule of
# include/tqec.h.  The absorption order of the prior factors is the leaf order of the OMEinsum tree chosen by
# `optimize_code` (TreeSA / GreedyMethod): a depth-first walk of the NestedEinsum that lists prior tensors in the order
# they are first contracted.  `TensorQECCUDA.lower` is a line-for-line port of tensorqec.jl_b200/schedule.py:lower
# (merge overlapping priors -> simulate the frontier -> emit per-step tables); it is omitted here for brevity and
# because it cannot be tested without Julia -- the Python implementation is the normative one.
function lower end

# ---- TNMAP ----------------------------------------------------------------------------------------------------------
struct CompiledTNMAPCUDA <: CompiledDecoder
    plan::Plan
    qubit_num::Int
end

function compile(decoder::TNMAP, problem::GeneralDecodingProblem; device::Integer = 0)
    factors = [(ix, vec(t)) for (ix, t) in zip(problem.ptn.code.ixs, problem.ptn.tensors)]   # column-major = first label fastest
    checks = [(problem.tanner.s2q[s], :syn, s) for s in 1:problem.tanner.ns]
    sch = lower(factors, checks, 0, problem.tanner.nq, problem.tanner.ns, 0; optimizer = decoder.optimizer)
    return CompiledTNMAPCUDA(Plan(sch, device), problem.tanner.nq)
end

# single shot: the reference signature (tndecoder.jl:53-57)
decode(ct::CompiledTNMAPCUDA, syn::SimpleSyndrome) =
    DecodingResult(true, vec(decode(ct, reshape(syn.s, :, 1)).error_pattern))

# batch: one column per shot
function decode(ct::CompiledTNMAPCUDA, syndromes::AbstractMatrix{Mod2})
    B = size(syndromes, 2)
    synd = pack(syndromes)
    corr = Matrix{UInt64}(undef, ct.plan.ncw, B)
    logp = Vector{Float64}(undef, B)
    GC.@preserve synd corr logp check(ccall((:tqec_decode_map, LIB), Cint,
        (Ptr{Cvoid}, Ptr{UInt64}, Int64, Ptr{UInt64}, Ptr{Float64}), ct.plan.h, synd, B, corr, logp))
    return DecodingResult(all(isfinite, logp), unpack(corr, ct.qubit_num))
end

# ---- TNMMAP (CSS) ---------------------------------------------------------------------------------------------------
struct CompiledTNMMAPCUDA <: CompiledDecoder
    tanner::CSSTannerGraph
    lx::Matrix{Mod2}
    lz::Matrix{Mod2}
    plan::Plan
end

function marginals(ct::CompiledTNMMAPCUDA, sx::AbstractMatrix{Mod2}, sz::AbstractMatrix{Mod2})
    B = size(sx, 2)
    synd = pack(vcat(sx, sz))
    mar = Array{Float64}(undef, 1 << ct.plan.n_obs, B)          # column-major `mar` of tndecoder.jl:162, one column per shot
    arg = Vector{Int32}(undef, B)
    GC.@preserve synd mar arg check(ccall((:tqec_decode_marginal, LIB), Cint,
        (Ptr{Cvoid}, Ptr{UInt64}, Int64, Ptr{Float64}, Ptr{Int32}), ct.plan.h, synd, B, mar, arg))
    return mar, arg .+ 1                                         # 1-based linear index = findmax(mar)[2]
end

end # module
