# ToeplitzMatricesB200Ext -- package extension that dispatches the FFT hot path of
# ToeplitzMatrices.jl on device-resident vectors to libtoeplitz_b200.so (hand-written sm_100a
# CUDA behind a C ABI, see include/toeplitz_b200.h).
#
# It follows the structure of the reference's own extension (ext/ToeplitzMatricesStatsBaseExt.jl):
# a weak dependency + an `[extensions]` entry in Project.toml (see INTEGRATION.md) and methods added
# to existing generics.  The weak dependency is any array package whose device vectors are
# `DenseVector`s exposing a raw device pointer (CUDA.jl's `CuArray` is used here ONLY as the
# container/allocator: no CUDA.jl kernels, no CUFFT -- every flop happens in libtoeplitz_b200.so).
#
# NOTE: there is no `julia` in the build image, so this file is a reviewed-only shim.  The tested
# boundary is the Python ctypes binding (toeplitzmatrices.jl_b200/_lib.py + api.py), which marshals
# arguments exactly like the `ccall`s below.
module ToeplitzMatricesB200Ext

using ToeplitzMatrices
using ToeplitzMatrices: AbstractToeplitz, SymmetricToeplitz, Circulant, LowerTriangularToeplitz,
    UpperTriangularToeplitz, TriangularToeplitz, Toeplitz, Hankel, IterativeLinearSolvers
using LinearAlgebra
import LinearAlgebra: mul!, ldiv!, factorize, eigvals, det, inv, pinv, adjoint
import Base: size, sqrt, *
using CUDA: CuArray, CuVector, CuMatrix, CuPtr, stream, context   # container + stream only

const libtb = "libtoeplitz_b200"          # resolved through LD_LIBRARY_PATH / artifacts
const DV{T} = CuVector{T}
const DM{T} = CuMatrix{T}
const BlasT = Union{Float32,Float64,ComplexF32,ComplexF64}

# ---- status -> exception mapping (SURVEY 8b "Errors") -----------------------------------------
const TB_OK, TB_DIM, TB_ARG, TB_UNSUP, TB_CUDA, TB_OOM = 0, 1, 2, 3, 4, 5
lasterr() = unsafe_string(ccall((:tb200_last_error, libtb), Cstring, ()))
function check(st::Cint)
    st == TB_OK && return nothing
    msg = lasterr()
    st == TB_DIM && throw(DimensionMismatch(msg))
    st == TB_ARG && throw(ArgumentError(msg))
    st == TB_OOM && throw(OutOfMemoryError())
    error(msg)                                     # ErrorException (unsupported / CUDA failure)
end

dtcode(::Type{Float32}) = Cint(0); dtcode(::Type{Float64}) = Cint(1)
dtcode(::Type{ComplexF32}) = Cint(2); dtcode(::Type{ComplexF64}) = Cint(3)
kindcode(::Toeplitz) = Cint(0); kindcode(::SymmetricToeplitz) = Cint(1); kindcode(::Circulant) = Cint(2)
kindcode(::LowerTriangularToeplitz) = Cint(3); kindcode(::UpperTriangularToeplitz) = Cint(4); kindcode(::Hankel) = Cint(5)
kindcode(::Type{<:Toeplitz}) = Cint(0); kindcode(::Type{<:SymmetricToeplitz}) = Cint(1); kindcode(::Type{<:Circulant}) = Cint(2)
kindcode(::Type{<:LowerTriangularToeplitz}) = Cint(3); kindcode(::Type{<:UpperTriangularToeplitz}) = Cint(4); kindcode(::Type{<:Hankel}) = Cint(5)
devptr(x) = Ptr{Cvoid}(UInt(pointer(x)))
curstream() = Ptr{Cvoid}(UInt(stream().handle))
ab(z::Number) = (c = ComplexF64(z); Cdouble[real(c), imag(c)])

# ---- the factorization record (replaces ToeplitzFactorization{T,A,S,P}, src/ToeplitzMatrices.jl:97-101)
mutable struct B200ToeplitzFactorization{T,A<:AbstractMatrix{T}} <: Factorization{T}
    handle::Ptr{Cvoid}
    dims::Tuple{Int,Int}
    function B200ToeplitzFactorization{T,A}(h, dims) where {T,A}
        F = new{T,A}(h, dims)
        finalizer(f -> ccall((:tb200_destroy, libtb), Cint, (Ptr{Cvoid},), f.handle), F)
        return F
    end
end
const B200CirculantFactorization{T} = B200ToeplitzFactorization{T,<:Circulant{T}}
size(F::B200ToeplitzFactorization) = F.dims                      # src/special.jl:262-263
size(F::B200ToeplitzFactorization, i::Integer) = F.dims[i]

function _factorize(A, vc, vr, m, n)
    T = eltype(A)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:tb200_factorize, libtb), Cint,
                (Cint, Cint, Int64, Int64, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Ptr{Cvoid}}),
                kindcode(A), dtcode(T), m, n, devptr(vc), vr === nothing ? C_NULL : devptr(vr), curstream(), h))
    return B200ToeplitzFactorization{T,typeof(A)}(h[], (m, n))
end
# src/linearalgebra.jl:122-131, 140-150, 184-192, 415-435; src/hankel.jl:112-121
factorize(A::Toeplitz{T,<:DV,<:DV}) where {T<:BlasT} = _factorize(A, A.vc, A.vr, size(A)...)
factorize(A::SymmetricToeplitz{T,<:DV}) where {T<:BlasT} = _factorize(A, A.v, nothing, size(A)...)
factorize(A::Circulant{T,<:DV}) where {T<:BlasT} = _factorize(A, A.v, nothing, size(A)...)
factorize(A::LowerTriangularToeplitz{T,<:DV}) where {T<:BlasT} = _factorize(A, A.v, nothing, size(A)...)
factorize(A::UpperTriangularToeplitz{T,<:DV}) where {T<:BlasT} = _factorize(A, A.v, nothing, size(A)...)
function factorize(A::Hankel{T,<:DV}) where {T<:BlasT}
    m, n = size(A)
    length(A.v) == m + n - 1 || throw(DimensionMismatch("first dimension of A, $(length(A.v) - n + 1), does not match length of y, $m"))
    _factorize(A, A.v, nothing, m, n)
end

# ---- mul! (src/linearalgebra.jl:4-99, src/hankel.jl:133-137) -----------------------------------
function _mul!(y, F::B200ToeplitzFactorization, x, α, β, ncols, ldy, ldx)
    check(ccall((:tb200_mul, libtb), Cint,
                (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Cint, Ptr{Cvoid}, Int64, Cint, Int64, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cvoid}),
                F.handle, devptr(y), ldy, dtcode(eltype(y)), devptr(x), ldx, dtcode(eltype(x)), ncols, ab(α), ab(β), curstream()))
    return y
end
function mul!(y::DV, F::B200ToeplitzFactorization, x::DV, α::Number, β::Number)
    (length(y) == F.dims[1] && length(x) == F.dims[2]) ||
        throw(DimensionMismatch("Toeplitz factorization does not match size of input and output vector"))
    _mul!(y, F, x, α, β, 1, length(y), length(x))
end
function mul!(C::DM, F::B200ToeplitzFactorization, B::DM, α::Number, β::Number)
    size(C, 2) == size(B, 2) || throw(DimensionMismatch("input and output matrices must have same number of columns"))
    _mul!(C, F, B, α, β, size(B, 2), stride(C, 2), stride(B, 2))          # the column loop is the batch axis
end
const DevStructured{T} = Union{Toeplitz{T,<:DV,<:DV},SymmetricToeplitz{T,<:DV},Circulant{T,<:DV},
                               LowerTriangularToeplitz{T,<:DV},UpperTriangularToeplitz{T,<:DV},Hankel{T,<:DV}}
function mul!(y::DV, A::DevStructured, x::DV, α::Number, β::Number)
    m, n = size(A)
    length(y) == m || throw(DimensionMismatch("first dimension of A, $m, does not match length of y, $(length(y))"))
    length(x) == n || throw(DimensionMismatch("second dimension of A, $n, does not match length of x, $(length(x))"))
    if m + n - 1 < 512            # src/linearalgebra.jl:19-34: direct O(mn) product, no factorization
        vc, vr = A isa Toeplitz ? (A.vc, A.vr) : (A.v, nothing)
        check(ccall((:tb200_mul_direct, libtb), Cint,
                    (Cint, Cint, Int64, Int64, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Cint, Ptr{Cvoid}, Int64, Cint, Int64,
                     Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cvoid}),
                    kindcode(A), dtcode(eltype(A)), m, n, devptr(vc), vr === nothing ? C_NULL : devptr(vr),
                    devptr(y), length(y), dtcode(eltype(y)), devptr(x), length(x), dtcode(eltype(x)), 1, ab(α), ab(β), curstream()))
        return y
    end
    mul!(y, factorize(A), x, α, β)
end
mul!(C::DM, A::DevStructured, B::DM, α::Number, β::Number) = mul!(C, factorize(A), B, α, β)

# ---- Circulant spectral family (src/linearalgebra.jl:183-315) -------------------------------------
function ldiv!(F::B200CirculantFactorization, b::Union{DV,DM})
    size(b, 1) == F.dims[1] || throw(DimensionMismatch("size of Toeplitz factorization does not match the length of the output vector"))
    check(ccall((:tb200_circ_ldiv, libtb), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Int64, Cint, Int64, Ptr{Cvoid}),
                F.handle, devptr(b), stride(b, 2), devptr(b), stride(b, 2), dtcode(eltype(b)), size(b, 2), curstream()))
    return b
end
ldiv!(C::Circulant{T,<:DV}, b::Union{DV,DM}) where {T} = ldiv!(factorize(C), b)
function eigvals(F::B200CirculantFactorization{T}) where {T}
    out = CuVector{complex(float(T))}(undef, F.dims[1])
    check(ccall((:tb200_circ_spectrum, libtb), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}), F.handle, devptr(out), curstream()))
    out
end
eigvals(C::Circulant{T,<:DV}) where {T} = eigvals(factorize(C))
function _map(F::B200ToeplitzFactorization{T,A}, op, tol = 0.0) where {T,A}
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:tb200_circ_map, libtb), Cint, (Ptr{Cvoid}, Cint, Cdouble, Ptr{Cvoid}, Ptr{Ptr{Cvoid}}), F.handle, op, tol, curstream(), h))
    B200ToeplitzFactorization{T,A}(h[], F.dims)
end
function _tocirculant(F::B200ToeplitzFactorization, ::Type{T}) where {T}
    vc = CuVector{T}(undef, F.dims[1])
    check(ccall((:tb200_circ_to_vc, libtb), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Ptr{Cvoid}), F.handle, devptr(vc), dtcode(T), curstream()))
    Circulant(vc)
end
inv(F::B200CirculantFactorization) = _map(F, 0)                                    # :250-253
adjoint(F::B200CirculantFactorization) = _map(F, 3)                                # :214-215
inv(C::Circulant{T,<:DV}) where {T} = _tocirculant(_map(factorize(C), 0), T)       # :244-249
pinv(C::Circulant{T,<:DV}, tol::Real = eps(real(float(one(T))))) where {T} = _tocirculant(_map(factorize(C), 1, tol), T)
sqrt(C::Circulant{T,<:DV}) where {T} = _tocirculant(_map(factorize(C), 2), T)      # :311-315
function *(A::B200CirculantFactorization{TA}, B::B200CirculantFactorization{TB}) where {TA,TB}   # :197-211
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:tb200_circ_mul, libtb), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Ptr{Cvoid}}), A.handle, B.handle, curstream(), h))
    T = promote_type(TA, TB)
    _tocirculant(B200ToeplitzFactorization{T,Circulant{T,CuVector{T}}}(h[], A.dims), T)
end
*(A::Circulant{<:Any,<:DV}, B::Circulant{<:Any,<:DV}) = factorize(A) * factorize(B)

function _precond(which, A)                                                         # strang :255-266 / chan :267-274
    n = size(A, 1); T = eltype(A)
    v = CuVector{T}(undef, n)
    vc, vr = A isa Toeplitz ? (A.vc, A.vr) : (A.v, nothing)
    check(ccall((:tb200_precond_vector, libtb), Cint, (Cint, Cint, Cint, Int64, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}),
                which, kindcode(A), dtcode(T), n, devptr(vc), vr === nothing ? C_NULL : devptr(vr), devptr(v), curstream()))
    Circulant(v)
end
ToeplitzMatrices.strang(A::DevStructured) = _precond(0, A)
ToeplitzMatrices.chan(A::DevStructured) = _precond(1, A)

# ---- iterative solvers (src/iterativeLinearSolvers.jl:9-215) ------------------------------------------
function _krylov(sym, A, x::DV{T}, b::DV{T}, M, max_it, tol) where {T}
    FA = A isa B200ToeplitzFactorization ? A : factorize(A)
    FM = M === nothing ? nothing : (M isa B200ToeplitzFactorization ? M : factorize(M))
    err = Ref{Cdouble}(0); it = Ref{Cint}(0); flag = Ref{Cint}(0)
    check(ccall((sym, libtb), Cint,
                (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cint, Int64, Cint, Cdouble, Ptr{Cdouble}, Ptr{Cint}, Ptr{Cint}, Ptr{Cvoid}),
                FA.handle, FM === nothing ? C_NULL : FM.handle, devptr(x), devptr(b), dtcode(T), length(b), max_it, tol, err, it, flag, curstream()))
    return x, err[], Int(it[]), Int(flag[])
end
IterativeLinearSolvers.cgs(A::Union{DevStructured,B200ToeplitzFactorization}, x::DV{T}, b::DV{T}, M, max_it::Integer, tol::Real) where {T} =
    _krylov(:tb200_cgs, A, x, b, M, max_it, tol)
function IterativeLinearSolvers.cg(A::Union{DevStructured,B200ToeplitzFactorization}, x::DV{T}, b::DV{T}, M, max_it::Integer, tol::Real) where {T<:LinearAlgebra.BlasReal}
    r = _krylov(:tb200_cg, A, x, b, M, max_it, tol)
    r[4] == 2 ? nothing : r                     # the reference's value-less early return (:57-59)
end
# drivers: src/linearalgebra.jl:133-136, 152, 399-402 (same constants: x0 = 0, 1000 iterations, 100eps())
ldiv!(A::Toeplitz{T,<:DV,<:DV}, b::DV) where {T} =
    copyto!(b, IterativeLinearSolvers.cgs(factorize(A), fill!(similar(b), 0), b, factorize(ToeplitzMatrices.strang(A)), 1000, 100eps())[1])
ldiv!(A::SymmetricToeplitz{T,<:DV}, b::DV) where {T} =
    copyto!(b, IterativeLinearSolvers.cg(factorize(A), fill!(similar(b, Float64), 0), b, factorize(ToeplitzMatrices.strang(A)), 1000, 100eps())[1])
ldiv!(A::TriangularToeplitz{T,<:DV}, b::DV) where {T} =
    copyto!(b, IterativeLinearSolvers.cgs(factorize(A), fill!(similar(b), 0), b, factorize(ToeplitzMatrices.chan(A)), 1000, 100eps())[1])

# matrix right-hand sides (src/linearalgebra.jl:102-110 is a serial column loop): all columns iterate together
function _krylov_batched(sym, A, X::DM{T}, B::DM{T}, M, max_it, tol) where {T}
    FA = A isa B200ToeplitzFactorization ? A : factorize(A)
    FM = M === nothing ? nothing : (M isa B200ToeplitzFactorization ? M : factorize(M))
    l = size(B, 2)
    err = zeros(Cdouble, l); it = zeros(Cint, l); flag = zeros(Cint, l)
    check(ccall((sym, libtb), Cint,
                (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Int64, Cint, Int64, Int64, Cint, Cdouble,
                 Ptr{Cdouble}, Ptr{Cint}, Ptr{Cint}, Ptr{Cvoid}),
                FA.handle, FM === nothing ? C_NULL : FM.handle, devptr(X), stride(X, 2), devptr(B), stride(B, 2), dtcode(T),
                size(B, 1), l, max_it, tol, err, it, flag, curstream()))
    return X, err, it, flag
end
function ldiv!(A::Toeplitz{T,<:DV,<:DV}, B::DM) where {T}
    size(A, 1) == size(A, 2) || error("Division: Rectangular case is not supported.")
    copyto!(B, _krylov_batched(:tb200_cgs_batched, A, fill!(similar(B), 0), B, ToeplitzMatrices.strang(A), 1000, 100eps())[1])
end
function ldiv!(A::SymmetricToeplitz{T,<:DV}, B::DM{Float64}) where {T}
    X, _, _, flag = _krylov_batched(:tb200_cg_batched, A, fill!(similar(B), 0), B, ToeplitzMatrices.strang(A), 1000, 100eps())
    any(==(2), flag) && throw(MethodError(copyto!, (B, nothing)))     # the reference's cg returned nothing for that column
    copyto!(B, X)
end
ldiv!(A::TriangularToeplitz{T,<:DV}, B::DM) where {T} =
    copyto!(B, _krylov_batched(:tb200_cgs_batched, A, fill!(similar(B), 0), B, ToeplitzMatrices.chan(A), 1000, 100eps())[1])

# ---- batched Levinson / Durbin (src/directLinearSolvers.jl:8-28, 93-134) ------------------------------
function ToeplitzMatrices.durbin!(y::Union{DV{T},DM{T}}, r::Union{DV{T},DM{T}}) where {T<:Union{Float32,Float64}}
    size(y, 1) == size(r, 1) || throw(DimensionMismatch("length(y) = $(size(y, 1)) ≠ $(size(r, 1)) = length(r)"))
    check(ccall((:tb200_durbin_batched, libtb), Cint, (Cint, Int64, Int64, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Int64, Ptr{Cvoid}),
                dtcode(T), size(r, 1), size(r, 2), devptr(r), stride(r, 2), devptr(y), stride(y, 2), curstream()))
    y
end
function ToeplitzMatrices.levinson!(x::Union{DV{T},DM{T}}, r::Union{DV{T},DM{T}}, b::Union{DV{T},DM{T}}, diag = nothing) where {T<:Union{Float32,Float64}}
    n = size(r, 1) + 1
    size(x, 1) == n || throw(DimensionMismatch("length(x) = $(size(x, 1)) ≠ $n = length(r) + 1"))
    size(b, 1) == n || throw(DimensionMismatch("length(b) = $(size(b, 1)) ≠ $n = length(r) + 1"))
    check(ccall((:tb200_levinson_batched, libtb), Cint,
                (Cint, Int64, Int64, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Ptr{Cvoid}),
                dtcode(T), n, size(b, 2), devptr(r), stride(r, 2), devptr(b), stride(b, 2), devptr(x), stride(x, 2),
                diag === nothing ? C_NULL : devptr(diag), curstream()))
    x
end

# many right-hand sides, one matrix (the column loop of ext/ToeplitzMatricesStatsBaseExt.jl:15-24): the y-recursion runs once
function ToeplitzMatrices.levinson(A::SymmetricToeplitz{T,<:DV}, B::DM{T}) where {T<:Union{Float32,Float64}}
    n = size(A, 1)
    size(B, 1) == n || throw(DimensionMismatch("first dimension of B, $(size(B, 1)), does not match size of A, $n"))
    X = similar(B)
    if 2 <= n <= 512
        check(ccall((:tb200_levinson_multirhs, libtb), Cint,
                    (Cint, Int64, Int64, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Int64, Ptr{Cvoid}),
                    dtcode(T), n, size(B, 2), devptr(A.vc), devptr(B), stride(B, 2), devptr(X), stride(X, 2), curstream()))
        return X
    end
    r0 = A.vc[1:1]
    r = repeat(A.vc[2:end] ./ r0, 1, size(B, 2))
    ToeplitzMatrices.levinson!(X, r, B, repeat(r0, size(B, 2)))
end

# ---- trench (src/directLinearSolvers.jl:36-85) and triangular inverse (src/linearalgebra.jl:337-396) ----
function ToeplitzMatrices.trench(T::SymmetricToeplitz{S,<:DV}) where {S<:Union{Float32,Float64}}
    n = size(T, 1)
    r0 = Array(T.vc[1:1])[1]
    r = isone(r0) ? T.vc[2:end] : T.vc[2:end] ./ r0
    B = CuMatrix{S}(undef, n, n)
    check(ccall((:tb200_trench, libtb), Cint, (Cint, Int64, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Cdouble, Ptr{Cvoid}),
                dtcode(S), n, devptr(r), devptr(B), n, Float64(r0), curstream()))
    Symmetric(B)
end
function _tri_inv_lower(v::DV{T}) where {T}
    n = length(v)
    if n <= 64
        out = similar(v)
        check(ccall((:tb200_tri_smallinv, libtb), Cint, (Cint, Int64, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}),
                    dtcode(T), n, devptr(v), devptr(out), curstream()))
        return out
    end
    np2 = nextpow(2, n)
    n != np2 && return _tri_inv_lower([v; fill!(similar(v, np2 - n), 0)])[1:n]
    nd2 = div(n, 2)
    a1 = _tri_inv_lower(v[1:nd2])
    [a1; -(LowerTriangularToeplitz(a1) * (Toeplitz(v[nd2+1:end], v[nd2+1:-1:2]) * a1))]
end
inv(A::LowerTriangularToeplitz{T,<:DV}) where {T} = LowerTriangularToeplitz(_tri_inv_lower(A.v))
inv(A::UpperTriangularToeplitz{T,<:DV}) where {T} = UpperTriangularToeplitz(_tri_inv_lower(A.v))

# ---- multi-GPU: one Julia process driving several devices ------------------------------------------------------
# Batches shard by columns / systems with no data-path collective (each device multiplies its own column block with
# its own handle).  The only exchange is the spectrum, once per factorization: factorize on the root device, create
# empty handles of the same shape on the others, tb200_bcast_spectrum (ncclBroadcast over NVLink, bound inside the
# library with dlopen -- no NCCL.jl needed).  A handle may be used from several tasks / streams at once: workspaces
# are keyed by stream (the reference's factorization shares one `tmp`, src/ToeplitzMatrices.jl:97-101).
mutable struct B200Comm
    handle::Ptr{Cvoid}
    ndev::Int
    function B200Comm(devs::Vector{Cint})
        h = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:tb200_comm_init, libtb), Cint, (Cint, Ptr{Cint}, Ptr{Ptr{Cvoid}}), length(devs), devs, h))
        c = new(h[], length(devs))
        finalizer(c -> ccall((:tb200_comm_destroy, libtb), Cint, (Ptr{Cvoid},), c.handle), c)
        return c
    end
end

"Empty factorization of the same shape on the CURRENT device (filled by `bcast_spectrum!`)."
function similar_factorization(F::B200ToeplitzFactorization{T,A}) where {T,A}
    h = Ref{Ptr{Cvoid}}(C_NULL)
    m, n = size(F)
    check(ccall((:tb200_factorize_empty, libtb), Cint, (Cint, Cint, Int64, Int64, Ptr{Cvoid}, Ptr{Ptr{Cvoid}}),
                kindcode(A), dtcode(T), m, n, curstream(), h))
    return B200ToeplitzFactorization{T,A}(h[], (m, n))
end

"`Fs[i]` lives on device `i` of the communicator; `root` (0-based) owns the spectrum."
function bcast_spectrum!(comm::B200Comm, Fs::Vector{<:B200ToeplitzFactorization}, root::Integer, streams::Vector{Ptr{Cvoid}})
    hs = Ptr{Cvoid}[F.handle for F in Fs]
    check(ccall((:tb200_bcast_spectrum, libtb), Cint, (Ptr{Cvoid}, Ptr{Ptr{Cvoid}}, Cint, Cint, Ptr{Ptr{Cvoid}}),
                comm.handle, hs, length(hs), root, streams))
    return Fs
end

end # module
