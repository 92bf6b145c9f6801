# B200Arnoldi.jl - the reference-side binding of libb200arnoldi.so.
#
# NOT EXECUTED IN THIS REPOSITORY'S CI: the build image has no Julia.  This is the file a
# maintainer of ArnoldiMethod.jl (or a wrapper package) adds; it leaves the reference package
# untouched and plugs the device path in by multiple dispatch on the array / operator types,
# exactly at the seams SURVEY.md 8(b) lists:
#
#   * operator contract        mul!(y, A, x)                    src/run.jl:21-25, src/expansion.jl:121
#   * array-type contract      ArnoldiWorkspace(V, H; V_tmp, Q) src/ArnoldiMethod.jl:81-92
#   * dispatch seam            iterate_arnoldi!, reinitialize!  src/expansion.jl:12,116
#                              the V*Q lines of _partialschur   src/run.jl:363-365, 382-383
#
# Two levels are offered:
#   b200_partialschur(A; ...)  - whole restart loop inside the library (b2a_partialschur)
#   methods on ArnoldiWorkspace{T,<:B200Matrix} - the reference's own _partialschur drives the
#       device through b2a_iterate_arnoldi / b2a_reinitialize / b2a_rotate_basis.
module B200Arnoldi

using LinearAlgebra, SparseArrays
import ArnoldiMethod
import ArnoldiMethod: ArnoldiWorkspace, PartialSchur, History, partialschur, partialeigen

const LIB = get(ENV, "B200ARNOLDI_LIB", "libb200arnoldi.so")

# ---- enums of include/b200arnoldi.h ---------------------------------------------------------
const B2A_F64, B2A_C64 = Cint(0), Cint(1)
const B2A_INIT_NONE, B2A_INIT_RAND, B2A_INIT_KEEP = Cint(0), Cint(1), Cint(2)
dtype_code(::Type{Float64}) = B2A_F64
dtype_code(::Type{ComplexF64}) = B2A_C64
which_code(w::Symbol) = Cint(findfirst(==(w), (:LM, :LR, :SR, :LI, :SI)) - 1)
which_code(::ArnoldiMethod.LM) = Cint(0); which_code(::ArnoldiMethod.LR) = Cint(1)
which_code(::ArnoldiMethod.SR) = Cint(2); which_code(::ArnoldiMethod.LI) = Cint(3)
which_code(::ArnoldiMethod.SI) = Cint(4)

struct B2AStats
    matvecs::Int64; passes::Int64; second_passes::Int64; breakdowns::Int64; launches::Int64
    bytes::Float64
end
struct B2AParams
    nev::Int32; which::Int32; tol::Float64; mindim::Int32; maxdim::Int32; restarts::Int32
    start_from::Int32; initialize::Int32; seed::UInt64
end
struct B2AHistory
    mvproducts::Int64; nconverged::Int32; converged::Int32; nev::Int32; restarts::Int32
    stats::B2AStats; ms_expand::Float64; ms_rotate::Float64; ms_small::Float64
end

function check(status::Cint)
    status == 0 && return nothing
    msg = unsafe_string(ccall((:b2a_last_error, LIB), Cstring, ()))
    status == -1 && throw(ArgumentError(msg))          # B2A_ERR_ARGUMENT
    status == -2 && throw(DimensionMismatch(msg))      # B2A_ERR_DIMENSION
    status == -6 && throw(msg)                         # schurfact.jl:406 throws a String
    error("libb200arnoldi status $status: $msg")
end

# ---- context ----------------------------------------------------------------------------------
mutable struct Context
    h::Ptr{Cvoid}
    function Context(device::Integer = 0)
        r = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:b2a_ctx_create, LIB), Cint, (Cint, Ref{Ptr{Cvoid}}), device, r))
        finalizer(c -> ccall((:b2a_ctx_destroy, LIB), Cint, (Ptr{Cvoid},), c.h), new(r[]))
    end
end
const DEFAULT_CTX = Ref{Union{Nothing,Context}}(nothing)
default_context() = something(DEFAULT_CTX[], (DEFAULT_CTX[] = Context(0)))

# ---- operator: Julia's SparseMatrixCSC goes in as it is (1-based Int64 colptr/rowval) ----------
mutable struct B200Operator{T}
    h::Ptr{Cvoid}
    n::Int
    ctx::Context
end
Base.eltype(::B200Operator{T}) where {T} = T
Base.size(A::B200Operator) = (A.n, A.n)
Base.size(A::B200Operator, i::Integer) = i <= 2 ? A.n : 1

function B200Operator(A::SparseMatrixCSC{T,Int64}; ctx = default_context(), mode = 0) where {T<:Union{Float64,ComplexF64}}
    n = LinearAlgebra.checksquare(A)
    r = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve A check(ccall((:b2a_csc_create, LIB), Cint,
        (Ptr{Cvoid}, Cint, Int64, Int64, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cint, Cint, Cint, Ref{Ptr{Cvoid}}),
        ctx.h, dtype_code(T), n, nnz(A), A.colptr, A.rowval, A.nzval, 64, 1, mode, r))
    finalizer(o -> ccall((:b2a_op_destroy, LIB), Cint, (Ptr{Cvoid},), o.h), B200Operator{T}(r[], n, ctx))
end

# matrix-free `mul!` contract: f(user, x_dev, y_dev, n, stream)::Cint enqueues y <- A x on `stream`
function B200Operator(::Type{T}, n::Integer, f::Base.CFunction; ctx = default_context()) where {T}
    r = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:b2a_op_from_callback, LIB), Cint,
        (Ptr{Cvoid}, Cint, Int64, Int64, Ptr{Cvoid}, Ptr{Cvoid}, Ref{Ptr{Cvoid}}),
        ctx.h, dtype_code(T), n, n, f, C_NULL, r))
    B200Operator{T}(r[], n, ctx)
end

# Shift-and-invert map x -> (A - sigma I) \ x on the device (the role of `construct_linear_map` in the reference's
# docs/src/index.md:234-262, which wraps `factorize(A)` in a LinearMap): Jacobi-CG on the CSR mat-vec of `A`, for a
# Hermitian positive definite A - sigma I.  `partialschur(shift_invert(Ad), which = :LM)`, then lambda = sigma + 1 / theta.
function shift_invert(A::B200Operator{T}; sigma = zero(T), rtol = 1e-13, maxit = 10_000) where {T}
    r = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:b2a_op_shift_invert, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Cdouble, Cint, Cdouble, Cint, Ref{Ptr{Cvoid}}),
        A.ctx.h, A.h, real(sigma), imag(sigma), 0, rtol, maxit, r))
    S = B200Operator{T}(r[], A.n, A.ctx)
    finalizer(o -> (ccall((:b2a_op_destroy, LIB), Cint, (Ptr{Cvoid},), o.h); A), S)  # keeps A alive as long as S
end

function solve_stats(S::B200Operator)
    solves, iters, worst = Ref{Int64}(0), Ref{Int64}(0), Ref{Cdouble}(0)
    check(ccall((:b2a_op_solve_stats, LIB), Cint, (Ptr{Cvoid}, Ref{Int64}, Ref{Int64}, Ref{Cdouble}), S.h, solves, iters, worst))
    (solves = solves[], iterations = iters[], worst_relres = worst[])
end

# ---- the device-resident basis: an AbstractMatrix whose storage is the library's workspace ----
mutable struct B200Matrix{T} <: AbstractMatrix{T}
    ws::Ptr{Cvoid}          # b2a_ws*
    n::Int
    cols::Int               # maxdim + 1
    ctx::Context
end
Base.size(V::B200Matrix) = (V.n, V.cols)
Base.getindex(V::B200Matrix{T}, i::Int, j::Int) where {T} = Array(view(V, :, j:j))[i]   # debugging only
function Base.Array(v::SubArray{T,2,<:B200Matrix{T}}) where {T}
    V = parent(v); cols = v.indices[2]
    out = Matrix{T}(undef, V.n, length(cols))
    check(ccall((:b2a_ws_get_cols, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Cvoid}, Int64),
        V.ws, first(cols), length(cols), out, V.n))
    out
end

"""
    B200Workspace(T, n, maxdim; ctx) -> ArnoldiWorkspace{T,<:B200Matrix}

The reference's own `ArnoldiWorkspace(V, H; V_tmp, Q)` constructor (src/ArnoldiMethod.jl:81-92) with a
device-resident `V`.  `H` and `Q` are Julia `Matrix` views of the host arrays inside the library handle,
so the reference's m x m code (local_schurfact!, partition_schur_three_way!, restore_arnoldi!, ...)
indexes them as usual.  `V_tmp` is a 0-column stand-in: the rotation is done in place.
"""
function B200Workspace(::Type{T}, n::Integer, maxdim::Integer; ctx = default_context()) where {T}
    r = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:b2a_ws_create, LIB), Cint, (Ptr{Cvoid}, Cint, Int64, Int64, Int64, Cint, Ref{Ptr{Cvoid}}),
        ctx.h, dtype_code(T), n, n, 0, maxdim, r))
    V = B200Matrix{T}(r[], n, maxdim + 1, ctx)
    finalizer(v -> ccall((:b2a_ws_destroy, LIB), Cint, (Ptr{Cvoid},), v.ws), V)
    Hp, Qp, ldh, ldq = Ref{Ptr{Cvoid}}(), Ref{Ptr{Cvoid}}(), Ref{Cint}(), Ref{Cint}()
    check(ccall((:b2a_ws_host_arrays, LIB), Cint, (Ptr{Cvoid}, Ref{Ptr{Cvoid}}, Ref{Cint}, Ref{Ptr{Cvoid}}, Ref{Cint}),
        V.ws, Hp, ldh, Qp, ldq))
    H = unsafe_wrap(Array, Ptr{T}(Hp[]), (maxdim + 1, maxdim))
    Q = unsafe_wrap(Array, Ptr{T}(Qp[]), (maxdim, maxdim))
    ArnoldiWorkspace(V, H; V_tmp = B200Matrix{T}(C_NULL, n, 0, ctx), Q = Q)
end

# ---- dispatch seam 1: expansion ---------------------------------------------------------------
function ArnoldiMethod.iterate_arnoldi!(A::B200Operator{T}, arnoldi::ArnoldiWorkspace{T,<:B200Matrix{T}},
                                        range::UnitRange{Int}) where {T}
    isempty(range) && return arnoldi
    # H of the workspace *is* the library's host H: the new columns appear in place
    check(ccall((:b2a_iterate_arnoldi, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Cint, UInt64, Ptr{Cvoid}, Cint, Ptr{Cvoid}),
        arnoldi.V.ws, A.h, first(range), last(range), rand(UInt64), C_NULL, 0, C_NULL))
    arnoldi
end

function ArnoldiMethod.reinitialize!(arnoldi::ArnoldiWorkspace{T,<:B200Matrix{T}}, j::Int = 0,
                                     populate! = nothing) where {T}
    ok = Ref{Cint}(0)
    mode = populate! === nothing ? B2A_INIT_RAND : B2A_INIT_KEEP
    if populate! !== nothing      # e.g. v -> copyto!(v, v1) (src/run.jl:126): stage on the host, upload
        v = Vector{T}(undef, arnoldi.V.n); populate!(v)
        check(ccall((:b2a_ws_set_col, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Cvoid}), arnoldi.V.ws, j + 1, v))
    end
    check(ccall((:b2a_reinitialize, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, UInt64, Ref{Cint}),
        arnoldi.V.ws, j, mode, rand(UInt64), ok))
    ok[] != 0
end

# ---- dispatch seam 2: the change of basis of the restart (src/run.jl:363-365, 382-383) ----------
# `mul!(view(V_tmp,:,purge:k), view(V,:,purge:maxdim), view(Q,purge:maxdim,purge:k))` lands here; the
# product is written straight back over V (in place by row tiles) and the two copyto! become no-ops.
function LinearAlgebra.mul!(C::SubArray{T,2,<:B200Matrix{T}}, Vv::SubArray{T,2,<:B200Matrix{T}},
                            Qv::SubArray{T,2,<:Matrix{T}}) where {T}
    V = parent(Vv); Q = parent(Qv)
    purge, maxdim = first(Vv.indices[2]), last(Vv.indices[2])
    k = purge + size(Qv, 2) - 1
    if purge == 1 && maxdim == k            # final rotation, run.jl:382
        check(ccall((:b2a_rotate_final, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Cvoid}, Cint, Ptr{Cvoid}),
            V.ws, k, Q, size(Q, 1), C_NULL))
    else                                     # restart rotation + V[:,k+1] <- V[:,maxdim+1], run.jl:363-365
        check(ccall((:b2a_rotate_basis, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Cint, Ptr{Cvoid}, Cint, Ptr{Cvoid}),
            V.ws, purge, k, maxdim, Q, size(Q, 1), C_NULL))
    end
    C
end
Base.copyto!(dst::SubArray{T,<:Any,<:B200Matrix{T}}, src::SubArray{T,<:Any,<:B200Matrix{T}}) where {T} = dst

# ---- whole loop in the library ------------------------------------------------------------------
"""
    b200_partialschur(A::B200Operator; nev, which, tol, mindim, maxdim, restarts, v1) -> PartialSchur, History

Same keywords, defaults and errors as `partialschur` (src/run.jl:100-129); runs `b2a_partialschur`.
"""
function b200_partialschur(A::B200Operator{T}; v1 = nothing, nev::Int = min(6, size(A, 1)),
        which = ArnoldiMethod.LM(), tol::Real = sqrt(eps(real(T))),
        mindim::Int = min(max(10, nev), size(A, 1)), maxdim::Int = min(max(20, 2nev), size(A, 1)),
        restarts::Int = 200) where {T}
    n = size(A, 1)
    nev < 1 && throw(ArgumentError("nev cannot be less than 1"))
    nev <= mindim <= maxdim <= n || throw(ArgumentError("nev ≤ mindim ≤ maxdim ≤ size(A, 1) does not hold, got $nev ≤ $mindim ≤ $maxdim ≤ $n"))
    arnoldi = B200Workspace(T, n, maxdim; ctx = A.ctx)
    init = B2A_INIT_RAND
    if v1 !== nothing
        length(v1) == n || throw(ArgumentError("v1 should have the same dimension as A"))
        v = convert(Vector{T}, v1)
        check(ccall((:b2a_ws_set_col, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Cvoid}), arnoldi.V.ws, 1, v))
        init = B2A_INIT_KEEP
    end
    p = Ref(B2AParams(nev, which_code(which), tol, mindim, maxdim, restarts, 1, init, rand(UInt64)))
    h = Ref{B2AHistory}()
    λ = Vector{ComplexF64}(undef, maxdim)
    check(ccall((:b2a_partialschur, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ref{B2AParams}, Ref{B2AHistory}, Ptr{Cvoid}),
        arnoldi.V.ws, A.h, p, h, λ))
    nc = Int(h[].nconverged)
    Q = view(arnoldi.V, :, 1:nc); R = view(arnoldi.H, 1:nc, 1:nc)      # views, like run.jl:375-376
    PartialSchur(Q, R, λ[1:nc]), History(Int(h[].mvproducts), nc, h[].converged != 0, nev)
end

end # module
