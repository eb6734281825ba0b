# CovarianceFunctionsB200.jl -- companion module that makes the B200 library a drop-in for the lazy-Gramian `mul!`
# path of CovarianceFunctions.jl v0.3.5.
#
# STATUS: syntax-reviewed, NOT executed -- Julia is not installed in the build image nor on the GPU box
# (DESIGN.md section 1).  The executable binding of the same C symbols is the Python mirror
# (covariancefunctions.jl_b200/_lib.py); this file shows what a maintainer adds on the reference side.
#
# It adds MORE SPECIFIC methods of LinearAlgebra.mul! / BlockFactorizations.blockmul! than the reference's
#   mul!(y::AbstractVector, G::Gramian, x::AbstractVector, α::Real, β::Real)          src/gramian.jl:78
#   mul!(Y::AbstractMatrix, G::Gramian, X::AbstractMatrix, α::Real, β::Real)          src/gramian.jl:89
#   blockmul!(y::AbstractVecOfVecOrMat, G::Gramian, x::AbstractVecOfVecOrMat, α, β)   src/gramian.jl:241
# for kernels that lower to the device; everything else keeps dispatching to the reference methods.
module CovarianceFunctionsB200

using LinearAlgebra
using CovarianceFunctions
using CovarianceFunctions: Gramian, EQ, Exp, RQ, MaternP, Dot, Constant, Sum, Product, Power, Lengthscale,
                           GradientKernel, IsotropicInput, input_trait
import BlockFactorizations

const libcovfn = get(ENV, "COVFN_B200_LIB", "libcovfn_b200.so")

# cf_knode_t (include/covfn_b200.h)
struct KNode
    op::Int32
    iparam::Int32
    fparam::Float64
end
const OP_EQ, OP_EXP, OP_RQ, OP_MATERNP, OP_DOT, OP_CONST, OP_SUM, OP_PROD, OP_POW, OP_LENGTHSCALE = Int32.(1:10)

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
function check(status::Cint)
    status == 0 && return
    msg = unsafe_string(ccall((:cf_last_error, libcovfn), Cstring, ()))
    status == -2 && throw(DimensionMismatch(msg))
    status == -4 || status == -7 ? throw(DomainError(msg)) : (status == -1 || status == -3 ? throw(ArgumentError(msg)) : error(msg))
end

# ---- device handles, cached per (kernel, x, y) because Gramian is immutable (SURVEY.md section 3.1) ---------------------
mutable struct Handle
    ptr::Ptr{Cvoid}
    function Handle(ptr)
        h = new(ptr)
        finalizer(h -> ccall((:cf_gramian_destroy, libcovfn), Cint, (Ptr{Cvoid},), h.ptr), h)
    end
end
const HANDLES = IdDict{Any, Handle}()   # keyed on the Gramian's x vector (objectid); cleared by `release!`
release!() = empty!(HANDLES)

# points must be a contiguous d x n column-major buffer: pack Vector{Vector{T}} (src/gramian.jl:2,154)
pack(x::AbstractVector{<:AbstractVector{T}}) where {T} = reduce(hcat, x)::Matrix{T}
pack(x::AbstractVector{T}) where {T<:Real} = reshape(collect(x), 1, :)

function handle(G::Gramian{T}, prog::Vector{KNode}) where {T<:Union{Float32, Float64}}
    get!(HANDLES, G.x) do
        X = pack(G.x)
        Y = G.x === G.y ? X : pack(G.y)
        size(X, 1) == size(Y, 1) || throw(DimensionMismatch("inputs have to have the same length"))
        out = Ref{Ptr{Cvoid}}(C_NULL)
        GC.@preserve X Y prog begin
            check(ccall((:cf_gramian_create, libcovfn), Cint,
                        (Ref{Ptr{Cvoid}}, Ptr{KNode}, Cint, Cint, Cint, Int64, Ptr{Cvoid}, Int64, Int64, Ptr{Cvoid}, Int64),
                        out, prog, length(prog), T === Float64 ? 1 : 0, size(X, 1), size(X, 2), X, size(X, 1),
                        size(Y, 2), G.x === G.y ? C_NULL : pointer(Y), size(Y, 1)))
        end
        Handle(out[])
    end
end

const Lowerable = Union{EQ, Exp, RQ, MaternP, Dot, Constant, Sum, Product, Power, Lengthscale}

# ---- mul!(y, G, x, α, β): vector and matrix (src/gramian.jl:78-99) --------------------------------------------------------
function LinearAlgebra.mul!(y::StridedVecOrMat{T}, G::Gramian{T, <:Lowerable}, x::StridedVecOrMat{T},
                            α::Real = 1, β::Real = 0) where {T<:Union{Float32, Float64}}
    prog = program(G.k)
    prog === nothing && return invoke(mul!, Tuple{typeof(y).name.wrapper, Gramian, typeof(x).name.wrapper, Real, Real}, y, G, x, α, β)
    size(y, 1) == size(G, 1) && size(x, 1) == size(G, 2) && size(y, 2) == size(x, 2) ||
        throw(DimensionMismatch("mul!: y $(size(y)), G $(size(G)), x $(size(x))"))
    h = handle(G, prog)
    GC.@preserve y x begin
        check(ccall((:cf_gramian_mul, libcovfn), Cint,
                    (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Int64, Int64, Cdouble, Cdouble),
                    h.ptr, y, stride(y, 2), x, stride(x, 2), size(y, 2), α, β))
    end
    return y
end

# ---- GradientKernel: flat (n d) vectors, entry (i-1)d + c (BlockFactorization isstrided, src/gramian.jl:120-123) -----------
function BlockFactorizations.blockmul!(y::AbstractVector{<:AbstractVector{T}},
                                       G::Gramian{<:Any, <:GradientKernel{<:Any, <:Lowerable, IsotropicInput}},
                                       x::AbstractVector{<:AbstractVector{T}}, α::Real = 1, β::Real = 0) where {T<:Float64}
    prog = program(G.k.k)
    prog === nothing && return invoke(BlockFactorizations.blockmul!, Tuple{Any, Gramian, Any, Real, Real}, y, G, x, α, β)
    yf, xf = parent(first(y)), parent(first(x))   # BlockFactorization passes views of one flat vector when isstrided
    h = handle(G, prog)
    GC.@preserve yf xf begin
        check(ccall((:cf_gradient_mul, libcovfn), Cint,
                    (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Int64, Int64, Cdouble, Cdouble),
                    h.ptr, yf, length(yf), xf, length(xf), 1, α, β))
    end
    return y
end

# DotProductInput kernels use the same entry point (the library dispatches on the lowered program's trait), and
# ValueGradientKernel blocks have d + 1 entries, entry 1 the value observation (src/gradient.jl:217-239, 400-474):
function BlockFactorizations.blockmul!(y::AbstractVector{<:AbstractVector{T}},
                                       G::Gramian{<:Any, <:GradientKernel{<:Any, <:Lowerable, CovarianceFunctions.DotProductInput}},
                                       x::AbstractVector{<:AbstractVector{T}}, α::Real = 1, β::Real = 0) where {T<:Float64}
    prog = program(G.k.k)
    prog === nothing && return invoke(BlockFactorizations.blockmul!, Tuple{Any, Gramian, Any, Real, Real}, y, G, x, α, β)
    yf, xf = parent(first(y)), parent(first(x))
    h = handle(G, prog)
    GC.@preserve yf xf check(ccall((:cf_gradient_mul, libcovfn), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Int64, Int64, Cdouble, Cdouble), h.ptr, yf, length(yf), xf, length(xf), 1, α, β))
    return y
end
function BlockFactorizations.blockmul!(y::AbstractVector{<:AbstractVector{T}},
                                       G::Gramian{<:Any, <:CovarianceFunctions.ValueGradientKernel{<:Any, <:Lowerable}},
                                       x::AbstractVector{<:AbstractVector{T}}, α::Real = 1, β::Real = 0) where {T<:Float64}
    prog = program(G.k.k)
    (prog === nothing || !(input_trait(G.k) isa Union{IsotropicInput, CovarianceFunctions.DotProductInput})) &&
        return invoke(BlockFactorizations.blockmul!, Tuple{Any, Gramian, Any, Real, Real}, y, G, x, α, β)
    yf, xf = parent(first(y)), parent(first(x))
    h = handle(G, prog)
    GC.@preserve yf xf check(ccall((:cf_value_gradient_mul, libcovfn), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Int64, Int64, Cdouble, Cdouble), h.ptr, yf, length(yf), xf, length(xf), 1, α, β))
    return y
end

# ---- (σ²I + K) \ b on the device (src/gramian.jl:55-60, src/lazy_linear_algebra.jl:135-144) ------------------------------------
function solve(G::Gramian{Float64, <:Lowerable}, σ²::Real, b::Vector{Float64}; reltol = 0.0, maxiter = 0, x0 = zeros(length(b)))
    prog = program(G.k)
    prog === nothing && throw(ArgumentError("kernel is not lowerable"))
    h = handle(G, prog)
    x = copy(x0); iters = Ref{Cint}(0); res = Ref{Cdouble}(0)
    GC.@preserve x b check(ccall((:cf_cg_solve, libcovfn), Cint,
        (Ptr{Cvoid}, Cdouble, Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Cint, Cint, Ref{Cint}, Ref{Cdouble}),
        h.ptr, σ², x, b, reltol, maxiter, 0, iters, res))
    return x, Int(iters[]), res[]
end

# several GPUs of one box: rows are sharded inside the library
init(devices::Vector{<:Integer}) = check(ccall((:cf_init, libcovfn), Cint, (Cint, Ptr{Cint}), length(devices), Cint.(devices)))

end # module
