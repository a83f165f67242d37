# CompressedSensingB200.jl -- thin Julia shim over libcsb200.so (B200 / sm_100a greedy pursuit).
#
# Keeps the call surface of CompressedSensing.jl's greedy-pursuit path
# (/root/reference/src/matchingpursuit.jl): `mp`, `omp`, `gomp` with the same positional and
# keyword methods, returning the same `SparseVector{Float64,Int64}`.  Nothing is exported, as in
# the reference (users write `using CompressedSensingB200: omp, gomp, mp`).  No CUDA.jl, no
# kernels here: every call is one `ccall` into the C ABI declared in include/csb200.h.
#
# NOTE: Julia is not installed in the build image, so this file has been checked by reading only;
# compressedsensing.jl_b200/__init__.py is the same shim over the same symbols in Python/ctypes
# and is what the GPU tests exercise.  See INTEGRATION.md.
module CompressedSensingB200

using SparseArrays
using Libdl

const libcsb200 = get(ENV, "CSB200_LIB", joinpath(@__DIR__, "..", "libcsb200.so"))

const CSB200_F64 = Cint(0)
const CSB200_F32 = Cint(1)

dtype_code(::Type{Float64}) = CSB200_F64
dtype_code(::Type{Float32}) = CSB200_F32

struct CSB200Error <: Exception
    status::Cint
    msg::String
end
Base.showerror(io::IO, e::CSB200Error) = print(io, "csb200 status $(e.status): $(e.msg)")

function check(rc::Cint, ε = nothing)
    rc == 0 && return
    # the reference throws a String here (src/matchingpursuit.jl:74,127)
    rc == -2 && throw("ε = $ε has to be non-negative")
    rc == -9 && throw(DimensionMismatch(unsafe_string(ccall((:csb200_strerror, libcsb200), Cstring, (Cint,), rc))))
    msg = unsafe_string(ccall((:csb200_strerror, libcsb200), Cstring, (Cint,), rc))
    detail = unsafe_string(ccall((:csb200_last_error, libcsb200), Cstring, ()))
    throw(CSB200Error(rc, isempty(detail) ? msg : "$msg: $detail"))
end

# ---------------------------------------------------------------------------------------------
# Dictionary handle: A stays resident in HBM; reuse it across calls.
# `devices = 0:7` replicates it onto several GPUs (csb200_dict_create_multi): `omp(D, B, k)` etc. then split the
# columns of B over those GPUs inside the library (internal host threads, no communication, bit-identical results).
mutable struct Dictionary{T<:Union{Float32,Float64}}
    handle::Ptr{Cvoid}
    M::Int
    N::Int
    function Dictionary(A::StridedMatrix{T}; device::Integer = 0, devices = nothing) where {T<:Union{Float32,Float64}}
        stride(A, 1) == 1 || (A = Matrix(A))
        M, N = size(A)
        h = Ref{Ptr{Cvoid}}(C_NULL)
        if devices === nothing
            GC.@preserve A check(ccall((:csb200_dict_create, libcsb200), Cint,
                (Ptr{Cvoid}, Int64, Int64, Int64, Cint, Cint, Ref{Ptr{Cvoid}}),
                pointer(A), M, N, max(stride(A, 2), M), dtype_code(T), device, h))
        else
            devs = Cint.(collect(devices))
            GC.@preserve A devs check(ccall((:csb200_dict_create_multi, libcsb200), Cint,
                (Ptr{Cvoid}, Int64, Int64, Int64, Cint, Ptr{Cint}, Cint, Ref{Ptr{Cvoid}}),
                pointer(A), M, N, max(stride(A, 2), M), dtype_code(T), devs, length(devs), h))
        end
        d = new{T}(h[], M, N)
        finalizer(d) do x
            x.handle == C_NULL || ccall((:csb200_dict_destroy, libcsb200), Cint, (Ptr{Cvoid},), x.handle)
            x.handle = C_NULL
        end
        return d
    end
end
Dictionary(A::AbstractMatrix; kw...) = Dictionary(Matrix{Float64}(A); kw...)   # anything else is copied
Base.size(D::Dictionary) = (D.M, D.N)
Base.size(D::Dictionary, i::Int) = size(D)[i]
Base.eltype(::Dictionary{T}) where {T} = T

const MatOrDict = Union{AbstractMatrix,Dictionary}
as_dictionary(A::Dictionary; kw...) = A
as_dictionary(A::AbstractMatrix; kw...) = Dictionary(A; kw...)
# A raw matrix is uploaded for this one call: its device copy is released when the call returns, not whenever Julia's
# GC -- which does not see device memory -- gets to the finalizer (a loop of omp(A, b, k) calls would fill the GPU).
# `devices` (e.g. 0:7) replicates a raw matrix onto several GPUs for the call: omp(A, B, k; devices = 0:7).
function with_dictionary(f, A::MatOrDict; devices = nothing)
    D = as_dictionary(A; devices = devices)
    try
        return f(D)
    finally
        A isa Dictionary || finalize(D)
    end
end
signal_eltype(A::Dictionary{T}) where {T} = T
signals(D::Dictionary{T}, b::AbstractVecOrMat) where {T} = begin
    size(b, 1) == D.M || throw(DimensionMismatch("A has $(D.M) rows, b has length $(size(b, 1))"))
    B = Matrix{T}(reshape(b, size(b, 1), :))          # contiguous copy in the dictionary's element type
    B
end

# (index, coefficient) pairs come back in SELECTION order, 0-based; the reference keeps x sorted by index
# (`x[i] = NaN` sorted insert, src/util.jl:120-122)
function to_sparse(N::Int, sel::AbstractVector{Int64}, coef::AbstractVector{Float64}, nnz::Integer)
    idx = sel[1:nnz] .+ 1
    p = sortperm(idx)
    return SparseVector(N, idx[p], coef[1:nnz][p])
end

function run_omp_like(fn::Symbol, D::Dictionary, b, l::Int, ε::Real, k::Int; csc::Bool = false)
    B = signals(D, b)
    nsig = size(B, 2)
    stride = max(k, 1)                 # the C call writes k slots per signal
    sel = Matrix{Int64}(undef, stride, nsig); coef = Matrix{Float64}(undef, stride, nsig)
    nnz = Vector{Int64}(undef, nsig); res = Vector{Float64}(undef, nsig); its = Vector{Int64}(undef, nsig)
    GC.@preserve B sel coef nnz res its begin
        rc = if fn === :omp
            ccall((:csb200_omp, libcsb200), Cint,
                (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Int64, Int64, Cdouble, Ptr{Int64}, Ptr{Cdouble}, Ptr{Int64}, Ptr{Cdouble}, Ptr{Int64}),
                D.handle, pointer(B), D.M, nsig, k, Float64(ε), sel, coef, nnz, res, its)
        else
            ccall((:csb200_gomp, libcsb200), Cint,
                (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Int64, Int64, Int64, Cdouble, Ptr{Int64}, Ptr{Cdouble}, Ptr{Int64}, Ptr{Cdouble}, Ptr{Int64}),
                D.handle, pointer(B), D.M, nsig, l, k, Float64(ε), sel, coef, nnz, res, its)
        end
        check(rc, ε)
    end
    csc && return assemble_csc(D.N, sel, coef, nnz)
    xs = [to_sparse(D.N, view(sel, :, s), view(coef, :, s), nnz[s]) for s in 1:nsig]
    return b isa AbstractVector ? xs[1] : xs
end

# batched result format: the N x nsig coefficient matrix as a SparseMatrixCSC (assembled by the library)
function assemble_csc(N::Int, sel::Matrix{Int64}, coef::Matrix{Float64}, nnz::Vector{Int64})
    stride, nsig = size(sel)
    colptr = Vector{Int64}(undef, nsig + 1); total = sum(nnz)
    rowval = Vector{Int64}(undef, total); nzval = Vector{Float64}(undef, total)
    check(ccall((:csb200_assemble_csc, libcsb200), Cint,
        (Int64, Int64, Ptr{Int64}, Ptr{Cdouble}, Ptr{Int64}, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Cdouble}),
        nsig, stride, sel, coef, nnz, 1, colptr, rowval, nzval))
    return SparseMatrixCSC(N, nsig, colptr, rowval, nzval)
end

# ---------------------------------------------------------------------------------------------
# omp  (src/matchingpursuit.jl:73-91)
function omp(A::MatOrDict, b::AbstractVecOrMat, ε::Real, k::Int = size(A, 1); csc::Bool = false, devices = nothing)
    ε ≥ 0 || throw("ε = $ε has to be non-negative")
    # csc = true: N x nsig SparseMatrixCSC (additive); devices = 0:7: split the columns of b over these GPUs (additive)
    with_dictionary(D -> run_omp_like(:omp, D, b, 1, ε, k; csc = csc), A; devices = devices)
end
omp(A::MatOrDict, b::AbstractVecOrMat, k::Int; kw...) = omp(A, b, eps(eltype(A)), k; kw...)
omp(A::MatOrDict, b::AbstractVecOrMat; max_residual = eps(eltype(A)), sparsity = min(size(A)...), kw...) =
    omp(A, b, max_residual, sparsity; kw...)

# gomp  (src/matchingpursuit.jl:126-148)
function gomp(A::MatOrDict, b::AbstractVecOrMat, l::Int, ε::Real, k::Int = size(A, 1); csc::Bool = false, devices = nothing)
    ε ≥ 0 || throw("ε = $ε has to be non-negative")
    with_dictionary(D -> run_omp_like(:gomp, D, b, l, ε, k; csc = csc), A; devices = devices)
end
gomp(A::MatOrDict, b::AbstractVecOrMat, l::Int, k::Int; kw...) = gomp(A, b, l, eps(eltype(A)), k; kw...)
gomp(A::MatOrDict, b::AbstractVecOrMat, l::Int; max_residual = eps(eltype(A)), sparsity = size(A, 2), kw...) =
    gomp(A, b, l, max_residual, sparsity; kw...)

# fr == ols == oomp == ormp  (src/forward.jl:33-54): forward regression; FP64 dictionaries
fr(A::MatOrDict, b::AbstractVecOrMat, max_ε::Real, min_δ::Real, k::Int = size(A, 1); csc::Bool = false) =
    with_dictionary(D -> fr_on(D, b, max_ε, min_δ, k; csc = csc), A)
function fr_on(D::Dictionary, b::AbstractVecOrMat, max_ε::Real, min_δ::Real, k::Int; csc::Bool = false)
    B = signals(D, b)
    nsig = size(B, 2)
    k = min(k, size(D)...)
    stride = max(k, 1)
    sel = Matrix{Int64}(undef, stride, nsig); coef = Matrix{Float64}(undef, stride, nsig)
    nnz = Vector{Int64}(undef, nsig); res = Vector{Float64}(undef, nsig); its = Vector{Int64}(undef, nsig)
    GC.@preserve B sel coef nnz res its check(ccall((:csb200_fr, libcsb200), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Int64, Int64, Cdouble, Cdouble, Ptr{Int64}, Ptr{Cdouble}, Ptr{Int64}, Ptr{Cdouble}, Ptr{Int64}),
        D.handle, pointer(B), D.M, nsig, k, Float64(max_ε), Float64(min_δ), sel, coef, nnz, res, its))
    csc && return assemble_csc(D.N, sel, coef, nnz)
    xs = [to_sparse(D.N, view(sel, :, s), view(coef, :, s), nnz[s]) for s in 1:nsig]
    return b isa AbstractVector ? xs[1] : xs
end
fr(A::MatOrDict, b::AbstractVecOrMat; max_residual::Real = 0., min_decrease::Real = 0., sparsity::Int = size(A, 2)) =
    fr(A, b, max_residual, min_decrease, sparsity)
const ols = fr
const oomp = fr
const ormp = fr

# sp  (src/twostage.jl:105-117) and oblivious  (src/oblivious.jl:3-8)
sp(A::MatOrDict, b::AbstractVecOrMat, k::Int, δ::Real = 1e-12; maxiter = 16k) =
    with_dictionary(D -> sp_on(D, b, k, δ; maxiter = maxiter), A)
function sp_on(D::Dictionary, b::AbstractVecOrMat, k::Int, δ::Real; maxiter = 16k)
    2k > D.M && error("2k = $(2k) > $(D.M) = length(b) is invalid for Subspace Pursuit")
    B = signals(D, b); nsig = size(B, 2); stride = max(k, 1)
    sel = Matrix{Int64}(undef, stride, nsig); coef = Matrix{Float64}(undef, stride, nsig)
    nnz = Vector{Int64}(undef, nsig); res = Vector{Float64}(undef, nsig); its = Vector{Int64}(undef, nsig)
    GC.@preserve B sel coef nnz res its check(ccall((:csb200_sp, libcsb200), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Int64, Int64, Cdouble, Int64, Ptr{Int64}, Ptr{Cdouble}, Ptr{Int64}, Ptr{Cdouble}, Ptr{Int64}),
        D.handle, pointer(B), D.M, nsig, k, Float64(δ), Int64(maxiter), sel, coef, nnz, res, its))
    xs = [to_sparse(D.N, view(sel, :, s), view(coef, :, s), nnz[s]) for s in 1:nsig]
    return b isa AbstractVector ? xs[1] : xs
end
oblivious(A::MatOrDict, b::AbstractVecOrMat, k::Int) = with_dictionary(D -> oblivious_on(D, b, k), A)
function oblivious_on(D::Dictionary, b::AbstractVecOrMat, k::Int)
    B = signals(D, b); nsig = size(B, 2); stride = max(k, 1)
    sel = Matrix{Int64}(undef, stride, nsig); coef = Matrix{Float64}(undef, stride, nsig)
    nnz = Vector{Int64}(undef, nsig); res = Vector{Float64}(undef, nsig)
    GC.@preserve B sel coef nnz res check(ccall((:csb200_oblivious, libcsb200), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Int64, Int64, Ptr{Int64}, Ptr{Cdouble}, Ptr{Int64}, Ptr{Cdouble}),
        D.handle, pointer(B), D.M, nsig, k, sel, coef, nnz, res))
    xs = [to_sparse(D.N, view(sel, :, s), view(coef, :, s), nnz[s]) for s in 1:nsig]   # length N (the reference: length M)
    return b isa AbstractVector ? xs[1] : xs
end

# mp  (src/matchingpursuit.jl:34-40); x is an optional warm start (single-signal form)
mp(A::MatOrDict, b::AbstractVector, k::Int, x::SparseVector = spzeros(size(A, 2))) =
    with_dictionary(D -> mp_on(D, b, k, x), A)
function mp_on(D::Dictionary, b::AbstractVector, k::Int, x::SparseVector)
    B = signals(D, b)
    stride = max(k, 1)
    sel = Vector{Int64}(undef, stride); coef = Vector{Float64}(undef, stride); res = Vector{Float64}(undef, 1)
    x0i = Int64.(x.nzind .- 1); x0v = Float64.(x.nzval); x0n = Int64[length(x0i)]
    GC.@preserve B sel coef res x0i x0v x0n check(ccall((:csb200_mp, libcsb200), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Int64, Int64, Ptr{Int64}, Ptr{Cdouble}, Ptr{Int64}, Int64, Ptr{Int64}, Ptr{Cdouble}, Ptr{Cdouble}),
        D.handle, pointer(B), D.M, 1, k, isempty(x0i) ? C_NULL : pointer(x0i), isempty(x0i) ? C_NULL : pointer(x0v),
        isempty(x0i) ? C_NULL : pointer(x0n), max(length(x0i), 1), sel, coef, res))
    for j in 1:k                      # x[i] += dot(view(A,:,i), r), in iteration order (:29)
        sel[j] ≥ 0 && (x[sel[j] + 1] += coef[j])
    end
    return x
end

# ---------------------------------------------------------------------------------------------
# Dictionary analysis (src/util.jl:2, 59-61, 96-117) -- csb200_dict_colnorms / csb200_dict_cumbabel
function colnorms(A::MatOrDict)
    with_dictionary(A) do D
        out = Vector{Float64}(undef, D.N)
        check(ccall((:csb200_dict_colnorms, libcsb200), Cint, (Ptr{Cvoid}, Ptr{Cdouble}), D.handle, out))
        out
    end
end
normalize!(A::StridedMatrix) = (A ./= eltype(A).(colnorms(A))'; A)            # src/util.jl:59-61
function cumbabel(A::MatOrDict, k::Int)                                          # src/util.jl:106-117
    with_dictionary(A) do D
        mu = Vector{Float64}(undef, k)
        check(ccall((:csb200_dict_cumbabel, libcsb200), Cint, (Ptr{Cvoid}, Int64, Ptr{Cdouble}), D.handle, k, mu))
        eltype(D).(mu)                                                           # `zeros(eltype(A), k)` (:107)
    end
end
babel(A::MatOrDict, k::Int) = cumbabel(A, k)[k]                                  # src/util.jl:101
coherence(A::MatOrDict) = babel(A, 1)                                            # src/util.jl:98

# ---------------------------------------------------------------------------------------------
# Device-resident batch (csb200_batch_*): upload a signal matrix once, solve repeatedly with no host<->device
# traffic, download when needed.  The stateful counterpart of the reference's `P = OMP(A, b, k)` objects
# (src/matchingpursuit.jl:44-60) for many right-hand sides.
mutable struct Batch{T}
    handle::Ptr{Cvoid}
    dict::Dictionary{T}
    nsig::Int
    max_sparsity::Int
    function Batch(D::Dictionary{T}, max_signals::Integer, max_sparsity::Integer) where {T}
        h = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:csb200_batch_create, libcsb200), Cint, (Ptr{Cvoid}, Int64, Int64, Ref{Ptr{Cvoid}}),
                    D.handle, max_signals, max_sparsity, h))
        b = new{T}(h[], D, 0, max_sparsity)
        finalizer(b) do x
            x.handle == C_NULL || ccall((:csb200_batch_destroy, libcsb200), Cint, (Ptr{Cvoid},), x.handle)
            x.handle = C_NULL
        end
        return b
    end
end
function upload!(P::Batch, b::AbstractVecOrMat)
    B = signals(P.dict, b)
    GC.@preserve B check(ccall((:csb200_batch_upload, libcsb200), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Int64),
                               P.handle, pointer(B), P.dict.M, size(B, 2)))
    P.nsig = size(B, 2)
    return P
end
omp!(P::Batch, k::Int, ε::Real = eps(eltype(P.dict))) =
    (check(ccall((:csb200_batch_omp, libcsb200), Cint, (Ptr{Cvoid}, Int64, Cdouble), P.handle, k, Float64(ε)), ε); P)
gomp!(P::Batch, l::Int, k::Int, ε::Real = eps(eltype(P.dict))) =
    (check(ccall((:csb200_batch_gomp, libcsb200), Cint, (Ptr{Cvoid}, Int64, Int64, Cdouble), P.handle, l, k, Float64(ε)), ε); P)
mp!(P::Batch, iters::Int) =
    (check(ccall((:csb200_batch_mp, libcsb200), Cint, (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Cdouble}, Ptr{Int64}, Int64),
                 P.handle, iters, C_NULL, C_NULL, C_NULL, 0)); P)
# results of the last omp! / gomp! as a vector of SparseVectors (or an N x nsig SparseMatrixCSC with csc = true)
function download(P::Batch; csc::Bool = false)
    stride, nsig = max(P.max_sparsity, 1), P.nsig
    sel = Matrix{Int64}(undef, stride, nsig); coef = Matrix{Float64}(undef, stride, nsig)
    nnz = Vector{Int64}(undef, nsig); res = Vector{Float64}(undef, nsig); its = Vector{Int64}(undef, nsig)
    check(ccall((:csb200_batch_download, libcsb200), Cint,
        (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Cdouble}, Ptr{Int64}, Ptr{Cdouble}, Ptr{Int64}),
        P.handle, stride, sel, coef, nnz, res, its))
    csc && return assemble_csc(P.dict.N, sel, coef, nnz)
    return [to_sparse(P.dict.N, view(sel, :, s), view(coef, :, s), nnz[s]) for s in 1:nsig], res, its
end
function last_solve_ms(P::Batch)
    ms = Ref{Cdouble}(0)
    check(ccall((:csb200_batch_last_solve_ms, libcsb200), Cint, (Ptr{Cvoid}, Ref{Cdouble}), P.handle, ms))
    return ms[]
end

# ---------------------------------------------------------------------------------------------
# Column-sharded single-dictionary mode (one Julia process per GPU, e.g. under MPI.jl): each rank holds the columns
# [n_offset+1, n_offset+size(A_local, 2)] of a dictionary with n_total atoms; the 128-byte communicator id is created
# on rank 0 (`comm_unique_id()`) and broadcast by the host program.
function shard_dictionary(A::StridedMatrix{T}, n_offset::Integer, n_total::Integer; device::Integer = 0) where {T<:Union{Float32,Float64}}
    stride(A, 1) == 1 || (A = Matrix(A))
    M, N = size(A)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve A check(ccall((:csb200_dict_create_shard, libcsb200), Cint,
        (Ptr{Cvoid}, Int64, Int64, Int64, Cint, Cint, Int64, Int64, Ref{Ptr{Cvoid}}),
        pointer(A), M, N, max(stride(A, 2), M), dtype_code(T), device, n_offset, n_total, h))
    return (handle = h[], M = M, n_total = Int(n_total), T = T)      # release with csb200_dict_destroy
end
function comm_unique_id()
    id = Vector{UInt8}(undef, 128)
    check(ccall((:csb200_comm_unique_id, libcsb200), Cint, (Ptr{UInt8},), id))
    return id
end
function comm_create(id::Vector{UInt8}, rank::Integer, nranks::Integer, device::Integer)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:csb200_comm_create, libcsb200), Cint, (Ptr{UInt8}, Cint, Cint, Cint, Ref{Ptr{Cvoid}}), id, rank, nranks, device, h))
    return h[]                                                            # release with csb200_comm_destroy
end
# `omp(A, b, k)` (src/matchingpursuit.jl:73-86) for ONE signal on the column-sharded dictionary; every rank calls it
# with the same b and gets the same SparseVector
function omp_sharded(shard, comm::Ptr{Cvoid}, b::AbstractVector, k::Int, ε::Real = eps(shard.T))
    ε ≥ 0 || throw("ε = $ε has to be non-negative")
    length(b) == shard.M || throw(DimensionMismatch("A has $(shard.M) rows, b has length $(length(b))"))
    bb = Vector{shard.T}(b); stride = max(k, 1)
    sel = Vector{Int64}(undef, stride); coef = Vector{Float64}(undef, stride)
    nnz = Ref{Int64}(0); res = Ref{Cdouble}(0); its = Ref{Int64}(0); ms = Ref{Cdouble}(0)
    GC.@preserve bb check(ccall((:csb200_omp_sharded, libcsb200), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Cdouble, Ptr{Int64}, Ptr{Cdouble}, Ref{Int64}, Ref{Cdouble}, Ref{Int64}, Ref{Cdouble}),
        shard.handle, comm, pointer(bb), k, Float64(ε), sel, coef, nnz, res, its, ms), ε)
    return to_sparse(shard.n_total, sel, coef, nnz[])
end

end # module
