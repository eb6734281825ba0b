# CovarianceFunctionsB200.jl -- companion module that makes the B200 library a drop-in for the lazy-Gramian `mul!` / `\`
# path of CovarianceFunctions.jl v0.3.5.
#
# STATUS: reviewed against the reference sources, NOT executed -- Julia is not installed in the build image nor on the GPU
# box (profiles/r2_julia_probe.txt).  Every `ccall` sequence this file can issue is replayed, call for call, by the C
# harness tests/abi/shim_sequence.c (built against include/covfn_b200.h and run on the GPU by tests/test_gpu_abi_shim.py);
# the comment `# [S<n>]` next to a ccall names the harness scenario that replays it.
#
# It adds MORE SPECIFIC methods than the reference's
#   mul!(y::AbstractVector, G::Gramian, x::AbstractVector, α::Real, β::Real)          src/gramian.jl:78
#   mul!(Y::AbstractMatrix, G::Gramian, X::AbstractMatrix, α::Real, β::Real)          src/gramian.jl:89
#   blockmul!(y::AbstractVecOfVecOrMat, G::Gramian, x::AbstractVecOfVecOrMat, α, β)   src/gramian.jl:241
#   ldiv!(x::AbstractVector, A::LazyFactorization, b::AbstractVector; kwargs...)      src/lazy_linear_algebra.jl:142
#   ldiv!(x::AbstractVector, B::BlockGramian, b::AbstractVector; kwargs...)           src/gramian.jl:236
# for kernels that lower to the device; everything else keeps dispatching to the reference methods (`@invoke`).
module CovarianceFunctionsB200

using LinearAlgebra
using CovarianceFunctions
using CovarianceFunctions: Gramian, EQ, Exp, RQ, MaternP, Dot, Constant, Sum, Product, Power, Lengthscale, Normed,
                           GradientKernel, ValueGradientKernel, IsotropicInput, DotProductInput, input_trait,
                           LazyMatrixSum, LazyFactorization
import BlockFactorizations
using BlockFactorizations: BlockFactorization

const libcovfn = get(ENV, "COVFN_B200_LIB", "libcovfn_b200.so")

# cf_knode_t (include/covfn_b200.h)
struct KNode
    op::Int32
    iparam::Int32
    fparam::Float64
end
const OP_EQ, OP_EXP, OP_RQ, OP_MATERNP, OP_DOT, OP_CONST, OP_SUM, OP_PROD, OP_POW, OP_LENGTHSCALE, OP_ARDSCALE, OP_ARD =
    Int32.(1:12)

# ---- lowering of kernel trees to postfix programs (walks the fields the reference defines) -------------------------
struct NotLowerable <: Exception end
program!(p, ::EQ) = push!(p, KNode(OP_EQ, 0, 0.0))                                   # src/stationary.jl:37-42
program!(p, ::Exp) = push!(p, KNode(OP_EXP, 0, 0.0))                                 # src/stationary.jl:56-60
program!(p, k::RQ) = push!(p, KNode(OP_RQ, k.α isa Integer ? 1 : 0, Float64(k.α)))   # src/stationary.jl:45-53
program!(p, k::MaternP) = push!(p, KNode(OP_MATERNP, k.p, 0.0))                      # src/stationary.jl:117-121
program!(p, ::Dot) = push!(p, KNode(OP_DOT, 0, 0.0))                                 # src/mercer.jl:6-9
program!(p, k::Constant) = k.c isa Real ? push!(p, KNode(OP_CONST, k.c isa Integer ? 1 : 0, Float64(k.c))) : throw(NotLowerable())
function program!(p, k::Sum)                                                          # src/algebra.jl:28-31
    foreach(a -> program!(p, a), k.args); push!(p, KNode(OP_SUM, length(k.args), 0.0))
end
function program!(p, k::Product)                                                      # src/algebra.jl:5-8
    foreach(a -> program!(p, a), k.args); push!(p, KNode(OP_PROD, length(k.args), 0.0))
end
program!(p, k::Power) = (program!(p, k.k); push!(p, KNode(OP_POW, k.p, 0.0)))          # src/algebra.jl:50-54
program!(p, k::Lengthscale) = (program!(p, k.k); push!(p, KNode(OP_LENGTHSCALE, 0, Float64(k.l))))  # src/transformation.jl:6-19
# ARD(k, l) = Normed(k, x -> enorm2(Diagonal(inv.(l)), x))  (src/transformation.jl:42-46): the closure `f` captures `l`, which
# Julia stores as the closure's field `l` (boxed if reassigned; it is not here).  Any other Normed is left to the reference.
function program!(p, k::Normed)
    f = k.n²
    (hasfield(typeof(f), :l) && getfield(f, :l) isa AbstractVector{<:Real}) || throw(NotLowerable())
    l = getfield(f, :l)
    program!(p, k.k)
    foreach(lc -> push!(p, KNode(OP_ARDSCALE, 0, Float64(lc))), l)     # one node per dimension ...
    push!(p, KNode(OP_ARD, length(l), 0.0))                            # ... popped by the ARD node: r² = Σ_c τ_c² / l_c
end
program!(p, k) = throw(NotLowerable())
function program(k)
    p = KNode[]
    try
        program!(p, k)
    catch e
        e isa NotLowerable && return nothing
        rethrow()
    end
    return p
end

# ---- errors -----------------------------------------------------------------------------------------------------------
function check(status::Integer)
    status == 0 && return
    msg = unsafe_string(ccall((:cf_last_error, libcovfn), Cstring, ()))
    status == -2 && throw(DimensionMismatch(msg))
    (status == -4 || status == -7) && throw(DomainError(status, msg))
    (status == -1 || status == -3) && throw(ArgumentError(msg))
    error(msg)
end

# ---- device handles -----------------------------------------------------------------------------------------------------
# One handle per (points x, points y, kernel program, eltype).  `Gramian` is immutable and cheap to rebuild (an optimisation
# loop creates a new one per hyper-parameter value on the SAME x), so the cache is keyed on the point vector x -- weakly: when x
# is collected its entry disappears and the handles' finalizers release the device memory -- and, below it, on everything else
# that defines the device state: objectid(y), the program (by value: kernel structure AND hyper-parameters), the eltype, and a
# fingerprint of the coordinates (points mutated in place must not leave a stale device copy).
mutable struct Handle
    ptr::Ptr{Cvoid}
    function Handle(ptr)
        h = new(ptr)
        finalizer(h) do hh
            hh.ptr == C_NULL || ccall((:cf_gramian_destroy, libcovfn), Cint, (Ptr{Cvoid},), hh.ptr)   # [S3]
            hh.ptr = C_NULL
        end
    end
end
const HandleKey = Tuple{UInt, Vector{KNode}, DataType, UInt}      # (objectid(y), program, T, fingerprint of x and y)
const HANDLES = WeakKeyDict{Any, Dict{HandleKey, Handle}}()
const HANDLES_LOCK = ReentrantLock()
const CHECK_MUTATION = Ref(true)       # hash the coordinates on every call (O(n d), no allocation); false: trust the caller
const MAX_HANDLES_PER_X = Ref(8)       # an optimisation loop leaves one handle per visited hyper-parameter: keep the newest few

"Release every cached device handle now (finalizers run immediately)."
function release!()
    lock(HANDLES_LOCK) do
        for d in values(HANDLES), h in values(d)
            finalize(h)
        end
        empty!(HANDLES)
    end
end

fingerprint(x) = CHECK_MUTATION[] ? foldl((h, xi) -> hash(xi, h), x; init = UInt(length(x))) : UInt(0)

# points must be a contiguous d x n column-major buffer OF THE GRAMIAN'S ELTYPE T (src/gramian.jl:2,30-33,154): a Float32 data
# set under a kernel with a Float64 parameter (0.5 * RQ(2)) is a Float64 Gramian, so the points are converted while packing.
function pack(::Type{T}, x::AbstractVector{<:AbstractVector}) where {T}
    d = length(first(x))
    X = Matrix{T}(undef, d, length(x))
    for (i, xi) in enumerate(x)
        length(xi) == d || throw(DimensionMismatch("inputs have to have the same length: $(d), $(length(xi))"))  # src/util.jl:41
        copyto!(view(X, :, i), xi)
    end
    return X
end
pack(::Type{T}, x::AbstractVector{<:Real}) where {T} = reshape(convert(Vector{T}, collect(x)), 1, :)

function handle(G::Gramian, ::Type{T}, prog::Vector{KNode}) where {T<:Union{Float32, Float64}}
    sym = G.x === G.y
    key = (objectid(G.y), prog, T, hash(sym ? UInt(0) : fingerprint(G.y), fingerprint(G.x)))
    lock(HANDLES_LOCK) do
        perx = get!(() -> Dict{HandleKey, Handle}(), HANDLES, G.x)
        get!(perx, key) do
            if length(perx) >= MAX_HANDLES_PER_X[]       # bounded: drop (and free) the older handles of this x
                foreach(finalize, values(perx)); empty!(perx)
            end
            X = pack(T, G.x)
            Y = sym ? X : pack(T, G.y)
            size(X, 1) == size(Y, 1) || throw(DimensionMismatch("inputs have to have the same length: $(size(X, 1)), $(size(Y, 1))"))
            out = Ref{Ptr{Cvoid}}(C_NULL)
            GC.@preserve X Y prog begin
                check(ccall((:cf_gramian_create, libcovfn), Cint,                                      # [S1] [S2]
                            (Ref{Ptr{Cvoid}}, Ptr{KNode}, Cint, Cint, Cint, Int64, Ptr{Cvoid}, Int64, Int64, Ptr{Cvoid}, Int64),
                            out, prog, length(prog), T === Float64 ? 1 : 0, size(X, 1), size(X, 2), X, size(X, 1),
                            size(Y, 2), sym ? C_NULL : pointer(Y), size(Y, 1)))
            end
            Handle(out[])
        end
    end
end

const Lowerable = Union{EQ, Exp, RQ, MaternP, Dot, Constant, Sum, Product, Power, Lengthscale, Normed}
const F = Union{Float32, Float64}

# ---- mul!(y, G, x, α, β): vector and matrix (src/gramian.jl:78-99) --------------------------------------------------------
# T is the GRAMIAN's eltype; y and x must already be Vector/Matrix{T} (otherwise the reference method runs).
function LinearAlgebra.mul!(y::StridedVector{T}, G::Gramian{T, <:Lowerable}, x::StridedVector{T},
                            α::Real, β::Real) where {T<:F}
    prog = program(G.k)
    (prog === nothing || stride(y, 1) != 1 || stride(x, 1) != 1) &&
        return @invoke mul!(y::AbstractVector, G::Gramian, x::AbstractVector, α::Real, β::Real)
    length(y) == size(G, 1) && length(x) == size(G, 2) ||
        throw(DimensionMismatch("mul!: y $(size(y)), G $(size(G)), x $(size(x))"))
    h = handle(G, T, prog)
    GC.@preserve y x h begin
        check(ccall((:cf_gramian_mul, libcovfn), Cint,                                                 # [S1]
                    (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Int64, Int64, Cdouble, Cdouble),
                    h.ptr, y, length(y), x, length(x), 1, α, β))
    end
    return y
end
function LinearAlgebra.mul!(Y::StridedMatrix{T}, G::Gramian{T, <:Lowerable}, X::StridedMatrix{T},
                            α::Real, β::Real) where {T<:F}
    prog = program(G.k)
    (prog === nothing || stride(Y, 1) != 1 || stride(X, 1) != 1) &&
        return @invoke mul!(Y::AbstractMatrix, G::Gramian, X::AbstractMatrix, α::Real, β::Real)
    size(Y, 1) == size(G, 1) && size(X, 1) == size(G, 2) && size(Y, 2) == size(X, 2) ||
        throw(DimensionMismatch("mul!: Y $(size(Y)), G $(size(G)), X $(size(X))"))
    h = handle(G, T, prog)
    GC.@preserve Y X h begin
        check(ccall((:cf_gramian_mul, libcovfn), Cint,                                                 # [S4]
                    (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Int64, Int64, Cdouble, Cdouble),
                    h.ptr, Y, stride(Y, 2), X, stride(X, 2), size(Y, 2), α, β))
    end
    return Y
end
LinearAlgebra.mul!(y::StridedVector{T}, G::Gramian{T, <:Lowerable}, x::StridedVector{T}) where {T<:F} = mul!(y, G, x, true, false)
LinearAlgebra.mul!(Y::StridedMatrix{T}, G::Gramian{T, <:Lowerable}, X::StridedMatrix{T}) where {T<:F} = mul!(Y, G, X, true, false)

# ---- GradientKernel / ValueGradientKernel: flat (n d) vectors, entry (i-1)d + c -----------------------------------------------
# BlockFactorization(G, isstrided = true) (src/gramian.jl:120-123) hands blockmul! the flat vectors as vectors of contiguous
# views (BlockFactorizations 1.2.2 [upstream]); the flat parent is recovered and checked, anything else runs the reference loop.
const DerivKernel{K} = Union{GradientKernel{<:Any, K, IsotropicInput}, GradientKernel{<:Any, K, DotProductInput},
                             ValueGradientKernel{<:Any, K, IsotropicInput}, ValueGradientKernel{<:Any, K, DotProductInput}}

# the flat Vector{T} behind a vector of contiguous equal-length views that tile it exactly, or `nothing`
function flat_parent(v::AbstractVector{<:AbstractVector{T}}, blk::Int) where {T<:F}
    isempty(v) && return nothing
    p = parent(first(v))
    (p isa Vector{T} && length(p) == blk * length(v)) || return nothing
    for (i, vi) in enumerate(v)
        (vi isa SubArray && parent(vi) === p && length(vi) == blk && first(parentindices(vi)[1]) == (i - 1) * blk + 1 &&
         stride(vi, 1) == 1) || return nothing
    end
    return p
end

# T is the element type of the flat vectors; the device path is taken when the points have the same element type (otherwise the reference's
# promotion rules apply and its method runs).  The derivative operators compute in Float64 inside the library for either T [S10].
points_eltype(G::Gramian) = isempty(G.x) ? Nothing : eltype(first(G.x))
function BlockFactorizations.blockmul!(y::AbstractVector{<:AbstractVector{T}},
                                       G::Gramian{<:Any, <:DerivKernel{<:Lowerable}},
                                       x::AbstractVector{<:AbstractVector{T}}, α::Real = 1, β::Real = 0) where {T<:F}
    prog = points_eltype(G) === T ? program(G.k.k) : nothing
    vg = G.k isa ValueGradientKernel
    blk = length(first(G.x)) + (vg ? 1 : 0)
    yf = prog === nothing ? nothing : flat_parent(y, blk)
    xf = yf === nothing ? nothing : flat_parent(x, blk)
    (yf === nothing || xf === nothing || length(y) != length(G.x) || length(x) != length(G.y)) &&
        return @invoke BlockFactorizations.blockmul!(y::AbstractVector{<:AbstractVecOrMat}, G::Gramian,
                                                      x::AbstractVector{<:AbstractVecOrMat}, α::Real, β::Real)
    h = handle(G, T, prog)
    GC.@preserve yf xf h begin
        if vg
            check(ccall((:cf_value_gradient_mul, libcovfn), Cint,                                      # [S5]
                        (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Int64, Int64, Cdouble, Cdouble),
                        h.ptr, yf, length(yf), xf, length(xf), 1, α, β))
        else
            check(ccall((:cf_gradient_mul, libcovfn), Cint,                                            # [S5]
                        (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Int64, Int64, Cdouble, Cdouble),
                        h.ptr, yf, length(yf), xf, length(xf), 1, α, β))
        end
    end
    return y
end

# ---- A \ b and ldiv!(x, A, b) on the device -----------------------------------------------------------------------------------
# The reference solves LazyMatrixSum(Diagonal, Gramian) and BlockGramian systems with IterativeSolvers.cg!
# (src/gramian.jl:55-60, src/lazy_linear_algebra.jl:135-144, src/gramian.jl:229-238), one host mul! per iteration;
# cf_cg_solve keeps the iterates on the device(s).  `\` needs no method of its own: the reference's `\` allocates x and calls ldiv!.
# Supported cg! keywords: reltol, maxiter (and abstol = 0, the default); anything else (Pl, log = true, ...) runs the reference.
cg_kwargs_ok(kw) = all(k -> k in (:reltol, :maxiter) || (k == :abstol && iszero(kw[k])) || (k == :log && kw[k] == false) ||
                             (k == :initially_zero), keys(kw))

# (reltol = 0 selects the library's default: sqrt(eps(T)), as cg! does for vectors of element type T)
function device_cg!(x::StridedVector{T}, G::Gramian, prog, σ²::Float64, b::StridedVector{T}, deriv::Int, kw) where {T<:F}
    length(x) == length(b) || throw(DimensionMismatch("ldiv!: x $(length(x)), b $(length(b))"))
    get(kw, :initially_zero, false) && fill!(x, 0)       # cg! semantics: x is the initial guess unless initially_zero
    h = handle(G, T, prog)
    iters = Ref{Cint}(0); res = Ref{Cdouble}(0)
    GC.@preserve x b h begin
        check(ccall((:cf_cg_solve, libcovfn), Cint,                                                    # [S6] [S7]
                    (Ptr{Cvoid}, Cdouble, Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Cint, Cint, Ref{Cint}, Ref{Cdouble}),
                    h.ptr, σ², x, b, Float64(get(kw, :reltol, 0.0)), Cint(get(kw, :maxiter, 0)), deriv, iters, res))
    end
    return x
end

# σ²I + K (either order).  Diagonals that are not a multiple of the identity are left to the reference.
const DiagGram{T} = Union{Tuple{<:Diagonal, <:Gramian{T, <:Lowerable}}, Tuple{<:Gramian{T, <:Lowerable}, <:Diagonal}}
function LinearAlgebra.ldiv!(x::StridedVector{T}, A::LazyMatrixSum{<:Any, <:DiagGram{T}}, b::StridedVector{T}; kwargs...) where {T<:F}
    D, G = A.args[1] isa Diagonal ? (A.args[1], A.args[2]) : (A.args[2], A.args[1])
    prog = program(G.k)
    d = D.diag
    uniform = !isempty(d) && all(==(first(d)), d) && first(d) isa Real
    (prog === nothing || !uniform || !cg_kwargs_ok(kwargs) || stride(x, 1) != 1 || stride(b, 1) != 1) &&
        return @invoke ldiv!(x::AbstractVector, A::LazyFactorization, b::AbstractVector; kwargs...)
    device_cg!(x, G, prog, Float64(first(d)), b, 0, kwargs)                                            # [S6]
end

# BlockGramian of a (value-)gradient kernel: G \ b (src/gramian.jl:229-238)
function LinearAlgebra.ldiv!(x::StridedVector{T},
                             B::BlockFactorization{<:Any, <:Gramian{<:Any, <:DerivKernel{<:Lowerable}}},
                             b::StridedVector{T}; kwargs...) where {T<:F}
    G = B.A                                  # the wrapped Gramian (field `A` of BlockFactorization [upstream 1.2.2])
    prog = points_eltype(G) === T ? program(G.k.k) : nothing
    (prog === nothing || !cg_kwargs_ok(kwargs) || stride(x, 1) != 1 || stride(b, 1) != 1) &&
        return @invoke ldiv!(x::AbstractVector, B::BlockFactorization{<:Any, <:Gramian}, b::AbstractVector; kwargs...)
    device_cg!(x, G, prog, 0.0, b, G.k isa ValueGradientKernel ? 2 : 1, kwargs)                        # [S7]
end

# explicit form with the iteration count and the final residual norm
function solve(G::Gramian{Float64, <:Lowerable}, σ²::Real, b::Vector{Float64}; reltol = 0.0, maxiter = 0, x0 = zeros(length(b)))
    prog = program(G.k)
    prog === nothing && throw(ArgumentError("kernel is not lowerable"))
    h = handle(G, Float64, prog)
    x = copy(x0); iters = Ref{Cint}(0); res = Ref{Cdouble}(0)
    GC.@preserve x b h check(ccall((:cf_cg_solve, libcovfn), Cint,
        (Ptr{Cvoid}, Cdouble, Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Cint, Cint, Ref{Cint}, Ref{Cdouble}),
        h.ptr, Float64(σ²), x, b, Float64(reltol), Cint(maxiter), 0, iters, res))
    return x, Int(iters[]), res[]
end

# several GPUs of one box: rows are sharded inside the library (handles created afterwards use all of them)
init(devices::Vector{<:Integer}) = check(ccall((:cf_init, libcovfn), Cint, (Cint, Ptr{Cint}), length(devices), Cint.(devices)))  # [S8]

end # module
