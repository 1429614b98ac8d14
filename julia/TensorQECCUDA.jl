# TensorQECCUDA.jl -- reference-side binding of libtqec_cuda.so (C ABI: include/tqec.h).
#
# NOT EXECUTED in this repository's CI: the build image and the GPU boxes have no Julia toolchain (SURVEY F4).  The
# same ABI is exercised call for call from Python (tensorqec.jl_b200/_cabi.py, tests/test_gpu_parity.py,
# tests/test_lower_cpp.py).  This file is what a TensorQEC.jl maintainer adds: new `CompiledDecoder` subtypes and
# `compile` / `decode` methods that keep `TNMAP` / `TNMMAP`, `compile(decoder, problem)` and
# `decode(compiled, syndrome)` unchanged (src/decoding/interfaces.jl:67-79, 96-119; src/decoding/tndecoder.jl:9-11, 77-80).
#
# There is NO lowering code on the Julia side: the host lists the prior tensors and the parity rows of the network it
# already builds (tndecoder.jl:33-40, 97-146, 186-219), optionally adds the leaf order of the contraction tree found by
# OMEinsum's optimiser, and calls `tqec_plan_compile`; the library lowers the graph (frontier schedule, in-place patch
# sweep with its tabulated head, or global-memory passes) and creates the plan.
module TensorQECCUDA

using TensorQEC
using TensorQEC: Mod2, SimpleTannerGraph, CSSTannerGraph, SimpleSyndrome, CSSSyndrome, CSSErrorPattern,
                 GeneralDecodingProblem, IndependentDepolarizingDecodingProblem, DetectorErrorModel, DecodingResult,
                 CompiledDecoder, TNMAP, TNMMAP, NoOptimizer, reduce2general, logical_operator, dem2tanner, nq, ns
import TensorQEC: compile, decode
using OMEinsum: NestedEinsum, DynamicEinCode, optimize_code, uniformsize

const LIB = get(ENV, "TQEC_CUDA_LIB", "libtqec_cuda.so")
const MAXPLUS, SUMPROD = Int32(0), Int32(1)

struct TqecError <: Exception
    code::Cint
    msg::String
end
check(rc::Cint) = rc == 0 ? nothing : throw(TqecError(rc, unsafe_string(ccall((:tqec_last_error, LIB), Cstring, ()))))

# ---- tqec_problem_desc (include/tqec.h) ------------------------------------------------------------------------------
struct ProblemDesc
    semiring::Int32; n_vars::Int32; n_checks::Int32; n_obs::Int32
    n_factors::Int32
    factor_ptr::Ptr{Int32}; factor_vars::Ptr{Int32}; factor_tables::Ptr{Float64}
    n_rows::Int32
    row_ptr::Ptr{Int32}; row_vars::Ptr{Int32}; row_kind::Ptr{Int32}; row_index::Ptr{Int32}
    order::Ptr{Int32}
    head_bits::Int32; table_bits::Int32; device::Int32; flags::Int32; wide_t_max::Int32
end

mutable struct Plan
    h::Ptr{Cvoid}
    nsw::Int; ncw::Int; n_obs::Int
end

"""
    compile_plan(semiring, n_vars, n_checks, n_obs, factors, rows; order, head_bits, table_bits, device)

`factors`: vector of `(labels::Vector{Int}, tensor::Array{Float64})` with 1-based variable labels, tensor axes in
label order (Julia arrays are column-major = "first variable fastest", the layout the ABI asks for).
`rows`: vector of `(vars::Vector{Int}, kind::Symbol, index::Int)`, kind `:syn` (clamped by syndrome bit `index`,
1-based) or `:obs` (open output axis `index`, 1-based).  `order`: `nothing` or a permutation of `1:length(factors)`.
"""
function compile_plan(semiring, n_vars, n_checks, n_obs, factors, rows; order = nothing, head_bits = 0, table_bits = 0,
                      device = 0)
    fptr = Int32[0]; fvars = Int32[]; ftab = Float64[]
    for (ix, t) in factors
        append!(fvars, Int32.(ix .- 1)); append!(ftab, vec(Float64.(t))); push!(fptr, length(fvars))
    end
    rptr = Int32[0]; rvars = Int32[]; rkind = Int32[]; rindex = Int32[]
    for (vs, kind, idx) in rows
        append!(rvars, Int32.(vs .- 1)); push!(rptr, length(rvars))
        push!(rkind, kind === :syn ? 0 : 1); push!(rindex, idx - 1)
    end
    ord = order === nothing ? Int32[] : Int32.(order .- 1)
    href = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve fptr fvars ftab rptr rvars rkind rindex ord begin
        d = ProblemDesc(semiring, n_vars, n_checks, n_obs, length(factors), pointer(fptr), pointer(fvars), pointer(ftab),
                        length(rows), pointer(rptr), pointer(rvars), pointer(rkind), pointer(rindex),
                        order === nothing ? Ptr{Int32}(C_NULL) : pointer(ord), head_bits, table_bits, device, 0, 0)
        check(ccall((:tqec_plan_compile, LIB), Cint, (Ref{ProblemDesc}, Ref{Ptr{Cvoid}}), d, href))
    end
    p = Plan(href[], cld(max(n_checks, 1), 64), cld(max(n_vars, 1), 64), n_obs)
    finalizer(x -> ccall((:tqec_plan_destroy, LIB), Cint, (Ptr{Cvoid},), x.h), p)
    return p
end

"""
    load_plan(path, n_vars, n_checks, n_obs; device = 0)

Plan from a lowered plan that `tqec_lowered_save` wrote (a host lowers once with `tqec_lower`, stores the result and every
rank or later session creates its plan from the file: no lowering, only the upload).
"""
function load_plan(path::AbstractString, n_vars, n_checks, n_obs; device = 0)
    lref = Ref{Ptr{Cvoid}}(C_NULL); href = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:tqec_lowered_load, LIB), Cint, (Cstring, Ref{Ptr{Cvoid}}), path, lref))
    rc = ccall((:tqec_plan_from_lowered, LIB), Cint, (Ptr{Cvoid}, Int32, Ref{Ptr{Cvoid}}), lref[], device, href)
    ccall((:tqec_lowered_destroy, LIB), Cint, (Ptr{Cvoid},), lref[])
    check(rc)
    p = Plan(href[], cld(max(n_checks, 1), 64), cld(max(n_vars, 1), 64), n_obs)
    finalizer(x -> ccall((:tqec_plan_destroy, LIB), Cint, (Ptr{Cvoid},), x.h), p)
    return p
end

# ---- contraction tree -> absorption order ------------------------------------------------------------------------------
# The reference lets OMEinsum choose a contraction tree (`optimize_code`, tndecoder.jl:48, 140-144, 213-217).  The
# frontier schedule absorbs the prior tensors one at a time, so it takes the tree's LEAF ORDER: a depth-first walk
# listing the prior leaves in the order the tree first reaches them.  `prior_leaf(i)` maps the tree's tensor index to the
# prior's position in `factors`, or `nothing` for the other leaves (parity tensors, unity / syndrome vectors).
function leaf_order(code::NestedEinsum, prior_leaf)
    out = Int[]
    walk(c) = c.tensorindex != -1 ? (k = prior_leaf(c.tensorindex); k === nothing || push!(out, k)) : foreach(walk, c.args)
    walk(code)
    return out
end

function optimiser_order(ixs, iy, prior_range, optimizer)
    (optimizer === nothing || optimizer isa NoOptimizer) && return nothing     # library's own sweep / natural order
    code = optimize_code(DynamicEinCode(ixs, iy), uniformsize(DynamicEinCode(ixs, iy), 2), optimizer)
    ord = leaf_order(code, i -> i in prior_range ? i - first(prior_range) + 1 : nothing)
    return isperm(ord) && length(ord) == length(prior_range) ? ord : nothing
end

# ---- bit packing: Vector{Mod2} columns <-> shot-major UInt64 words ---------------------------------------------------------
# This IS `compresscol` (src/codes/mod2.jl:58-71) applied to the (bits x shots) matrix.
pack(bits::AbstractMatrix{Mod2}) = TensorQEC.compresscol(bits)            # (ceil(nbits/64), shots)
unpack(words::Matrix{UInt64}, nbits::Int) =
    [Mod2((words[(i - 1) >> 6 + 1, s] >> ((i - 1) & 63)) & 1 == 1) for i in 1:nbits, s in axes(words, 2)]

# ---- TNMAP (tndecoder.jl:16-57) ----------------------------------------------------------------------------------------
struct CompiledTNMAPCUDA <: CompiledDecoder
    plan::Plan
    qubit_num::Int
end

function compile(decoder::TNMAP, problem::GeneralDecodingProblem; device::Integer = 0, head_bits::Integer = 0)
    t = problem.tanner
    factors = [(collect(ix), tn) for (ix, tn) in zip(problem.ptn.code.ixs, problem.ptn.tensors)]
    rows = [(t.s2q[s], :syn, s) for s in 1:t.ns]
    # the reference's network (stg2uaimodel, tndecoder.jl:33-40): parity factors over (s2q[s]..., nq + s), then the priors
    ixs = vcat([vcat(t.s2q[s], t.nq + s) for s in 1:t.ns], [collect(ix) for ix in problem.ptn.code.ixs])
    order = optimiser_order(ixs, Int[], (t.ns + 1):(t.ns + length(factors)), decoder.optimizer)
    plan = compile_plan(MAXPLUS, t.nq, t.ns, 0, factors, rows; order, head_bits, device)
    return CompiledTNMAPCUDA(plan, t.nq)
end

# single shot: the reference signature (tndecoder.jl:53-57)
function decode(ct::CompiledTNMAPCUDA, syn::SimpleSyndrome)
    res = decode(ct, reshape(syn.s, :, 1))
    return DecodingResult(res.success_tag, vec(res.error_pattern))
end

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

# ---- GF(2) helpers for error_pattern (replace the per-shot SCIP program, ipdecoder.jl:150-169) ----------------------------
mutable struct GF2
    h::Ptr{Cvoid}
    rows::Int; cols::Int
end
function GF2(M::AbstractMatrix{Bool}, device::Integer)
    packed = TensorQEC.compresscol(Mod2.(permutedims(M)))                 # row r of M -> column r: ceil(cols/64) words per row
    href = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve packed check(ccall((:tqec_gf2_create, LIB), Cint, (Int32, Int32, Ptr{UInt64}, Int32, Ref{Ptr{Cvoid}}),
                                    size(M, 1), size(M, 2), packed, device, href))
    g = GF2(href[], size(M, 1), size(M, 2))
    finalizer(x -> ccall((:tqec_gf2_destroy, LIB), Cint, (Ptr{Cvoid},), x.h), g)
    return g
end

"R (n x m) with H (R s) = s for every s in the column space of H (free variables 0)."
function gf2_right_inverse(H::AbstractMatrix{Bool})
    m, n = size(H)
    A = hcat(Matrix{Bool}(H), Matrix{Bool}(LinearAlgebra.I, m, m))
    piv = Int[]; r = 1
    for c in 1:n
        r > m && break
        p = findfirst(@view A[r:m, c]); p === nothing && continue
        p += r - 1
        A[[r, p], :] = A[[p, r], :]
        for k in 1:m
            (k != r && A[k, c]) && (A[k, :] .⊻= A[r, :])
        end
        push!(piv, c); r += 1
    end
    R = falses(n, m)
    for (i, c) in enumerate(piv)
        R[c, :] = A[i, n+1:end]
    end
    return Matrix{Bool}(R)
end

"Undetectable patterns f (H f = 0) whose sector flips L f are a reduced-echelon basis of all reachable flips."
function gf2_sector_fixes(H::AbstractMatrix{Bool}, L::AbstractMatrix{Bool})
    m, n = size(H); k = size(L, 1)
    A = hcat(Matrix{Bool}(permutedims(H)), Matrix{Bool}(LinearAlgebra.I, n, n))
    r = 1
    for c in 1:m
        r > n && break
        p = findfirst(@view A[r:n, c]); p === nothing && continue
        p += r - 1
        A[[r, p], :] = A[[p, r], :]
        for q in 1:n
            (q != r && A[q, c]) && (A[q, :] .⊻= A[r, :])
        end
        r += 1
    end
    ker = A[r:end, m+1:end]                                                # rows: a basis of ker H
    D = isodd.(Int.(ker) * Int.(permutedims(L)))
    M = hcat(D, ker); r = 1
    for j in 1:k
        r > size(M, 1) && break
        p = findfirst(@view M[r:end, j]); p === nothing && continue
        p += r - 1
        M[[r, p], :] = M[[p, r], :]
        for q in 1:size(M, 1)
            (q != r && M[q, j]) && (M[q, :] .⊻= M[r, :])
        end
        r += 1
    end
    return Matrix{Bool}(M[1:r-1, k+1:end])
end

function coset_rep(R::GF2, L::GF2, FIX::GF2, synd::Matrix{UInt64}, sector::Vector{Int32})
    B = size(synd, 2)
    err = Matrix{UInt64}(undef, cld(R.rows, 64), B)
    ok = Vector{UInt8}(undef, B)
    GC.@preserve synd sector err ok check(ccall((:tqec_coset_rep, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{UInt64}, Ptr{Int32}, Int64, Ptr{UInt64}, Ptr{UInt8}),
        R.h, L.h, FIX.h, synd, sector, B, err, ok))
    return err, ok .== 1
end

function marginals(plan::Plan, synd::Matrix{UInt64})
    B = size(synd, 2)
    mar = Array{Float64}(undef, 1 << plan.n_obs, B)             # column-major `mar` of tndecoder.jl:162, one column per shot
    arg = Vector{Int32}(undef, B)
    GC.@preserve synd mar arg check(ccall((:tqec_decode_marginal, LIB), Cint,
        (Ptr{Cvoid}, Ptr{UInt64}, Int64, Ptr{Float64}, Ptr{Int32}), plan.h, synd, B, mar, arg))
    return mar, arg                                              # arg: 0-based linear index = findmax(mar)[2] - 1
end

# ---- TNMMAP, CSS (tndecoder.jl:85-174) ------------------------------------------------------------------------------------
struct CompiledTNMMAPCUDA <: CompiledDecoder
    tanner::CSSTannerGraph
    lx::Matrix{Mod2}
    lz::Matrix{Mod2}
    plan::Plan
    R::GF2; L::GF2; FIX::GF2
end

function compile(decoder::TNMMAP, problem::IndependentDepolarizingDecodingProblem; device::Integer = 0, head_bits::Integer = 0)
    tanner = problem.tanner
    n = nq(tanner); nsx = ns(tanner.stgx); nsz = ns(tanner.stgz)
    lx, lz = logical_operator(tanner); k = size(lx, 1)
    p = problem.pvec
    factors = [([i, i + n], [1 - p.px[i] - p.py[i] - p.pz[i] p.pz[i]; p.px[i] p.py[i]]) for i in 1:n]   # general_decoding.jl:5
    rows = vcat([(tanner.stgx.s2q[i] .+ n, :syn, i) for i in 1:nsx], [(tanner.stgz.s2q[i], :syn, nsx + i) for i in 1:nsz],
                [(findall(x -> x.x, lx[i, :]) .+ n, :obs, i) for i in 1:k],      # iy order of tndecoder.jl:134
                [(findall(x -> x.x, lz[i, :]), :obs, k + i) for i in 1:k])
    plan = compile_plan(SUMPROD, 2n, nsx + nsz, 2k, factors, rows; head_bits, device)   # head_bits = 0: the library's default (14)
    Hx = [a.x for a in tanner.stgx.H]; Hz = [a.x for a in tanner.stgz.H]
    R = falses(2n, nsx + nsz); R[1:n, nsx+1:end] = gf2_right_inverse(Hz); R[n+1:end, 1:nsx] = gf2_right_inverse(Hx)
    L = falses(2k, 2n); FIX = falses(2k, 2n)
    L[1:k, n+1:end] = [a.x for a in lx]; FIX[1:k, n+1:end] = [a.x for a in lz]
    L[k+1:end, 1:n] = [a.x for a in lz]; FIX[k+1:end, 1:n] = [a.x for a in lx]
    return CompiledTNMMAPCUDA(tanner, lx, lz, plan, GF2(Matrix{Bool}(R), device), GF2(Matrix{Bool}(L), device), GF2(Matrix{Bool}(FIX), device))
end

function decode(ct::CompiledTNMMAPCUDA, syn::CSSSyndrome)
    n = nq(ct.tanner)
    synd = pack(reshape(vcat(syn.sx, syn.sz), :, 1))
    mar, arg = marginals(ct.plan, synd)
    err, ok = coset_rep(ct.R, ct.L, ct.FIX, synd, arg)
    e = vec(unpack(err, 2n))
    return DecodingResult(ok[1] && maximum(mar) > 0, CSSErrorPattern(e[1:n], e[n+1:2n]))
end

# ---- TNMMAP, detector error model (tndecoder.jl:176-271) --------------------------------------------------------------------
struct CompiledDEMTNMMAPCUDA <: CompiledDecoder
    tanner::SimpleTannerGraph
    plan::Plan
    R::GF2; L::GF2; FIX::GF2
end

function compile(decoder::TNMMAP, dem::DetectorErrorModel; device::Integer = 0)
    tanner = dem2tanner(dem)
    ne = tanner.nq; nd = tanner.ns
    l2q = [findall(fd -> l in fd, dem.flipped_detectors) for l in dem.logical_list]
    factors = [([e], [1 - dem.error_rates[e], dem.error_rates[e]]) for e in 1:ne]           # tndecoder.jl:202-205
    rows = vcat([(tanner.s2q[d], :syn, d) for d in 1:nd], [(l2q[l], :obs, l) for l in eachindex(l2q)])
    plan = compile_plan(SUMPROD, ne, nd, length(l2q), factors, rows; device)
    H = [a.x for a in tanner.H]
    L = falses(length(l2q), ne)
    for (l, c) in enumerate(l2q); L[l, c] .= true; end
    Lm = Matrix{Bool}(L)
    return CompiledDEMTNMMAPCUDA(tanner, plan, GF2(gf2_right_inverse(H), device), GF2(Lm, device), GF2(gf2_sector_fixes(H, Lm), device))
end

function decode(ct::CompiledDEMTNMMAPCUDA, syn::SimpleSyndrome)
    synd = pack(reshape(syn.s, :, 1))
    mar, arg = marginals(ct.plan, synd)
    err, ok = coset_rep(ct.R, ct.L, ct.FIX, synd, arg)
    return DecodingResult(ok[1] && maximum(mar) > 0, vec(unpack(err, ct.tanner.nq)))
end

# ---- multi-GPU: one process per GPU (Distributed.jl workers), shots sharded, counters all-reduced in the library ------------
# Replaces SimpleMultiprocessing.multiprocess_run (src/multiprocessing.jl:41-52).
mutable struct Comm
    h::Ptr{Cvoid}
end
function unique_id()
    id = Vector{UInt8}(undef, 128)
    GC.@preserve id check(ccall((:tqec_comm_unique_id, LIB), Cint, (Ptr{UInt8},), id))
    return id                                                    # ship to the other workers, e.g. `@everywhere id = \$id`
end
function Comm(nranks::Integer, rank::Integer, id::Vector{UInt8}, device::Integer)
    href = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve id check(ccall((:tqec_comm_init, LIB), Cint, (Int32, Int32, Ptr{UInt8}, Int32, Ref{Ptr{Cvoid}}),
                                nranks, rank, id, device, href))
    c = Comm(href[])
    finalizer(x -> ccall((:tqec_comm_destroy, LIB), Cint, (Ptr{Cvoid},), x.h), c)
    return c
end

struct McDesc
    plan::Ptr{Cvoid}; H::Ptr{Cvoid}; L::Ptr{Cvoid}; row_class::Ptr{Int32}
    model::Int32; n_sites::Int32
    p0::Ptr{Float64}; p1::Ptr{Float64}; p2::Ptr{Float64}
    chunk::Int64
    comm::Ptr{Cvoid}
end

"""
    mc_run(ct, H, L, row_class, px, py, pz; seed, shot_offset, shots, comm)

The fused sample -> syndrome -> decode -> check pipeline of `multi_round_qec` (src/decoding/threshold.jl:1-19) on this
rank's shot range; returns `(logical_x, logical_z, logical_any, shots)`, summed over all ranks when `comm` is given.
"""
function mc_run(ct::CompiledTNMAPCUDA, H::GF2, L::GF2, row_class::Vector{Int32}, px, py, pz; seed = 0, shot_offset = 0,
                shots, comm::Union{Comm,Nothing} = nothing)
    counts = zeros(Int64, 4); ms = Ref{Cfloat}(0)
    p0, p1, p2 = Float64.(px), Float64.(py), Float64.(pz)
    GC.@preserve row_class p0 p1 p2 counts begin
        d = McDesc(ct.plan.h, H.h, L.h, pointer(row_class), 1, length(p0), pointer(p0), pointer(p1), pointer(p2), 0,
                   comm === nothing ? C_NULL : comm.h)
        check(ccall((:tqec_mc_run, LIB), Cint, (Ref{McDesc}, UInt64, Int64, Int64, Ptr{Int64}, Ref{Cfloat}),
                    d, seed, shot_offset, shots, counts, ms))
    end
    return counts, ms[]
end

import LinearAlgebra

end # module
