# make_golden.jl -- golden vectors from the UNMODIFIED reference, so that parity can be pinned on reference outputs.
#
#   julia --project=/path/to/CovarianceFunctions.jl -t auto julia/make_golden.jl [outdir = tests/golden/julia_v1]
#
# STATUS: cannot be run in the build image or on the GPU box (no Julia: profiles/r2_julia_probe.txt).  It needs only the
# reference package and its own dependencies (LinearAlgebra, Random, BlockFactorizations); no third-party I/O package: arrays are
# written as raw little-endian binaries and the manifest as hand-written JSON.  tests/test_julia_golden.py consumes the output
# when it is present and compares BOTH the CPU oracle and the GPU library with it (1e-12 Float64 / 1e-5 Float32); until then that
# test reports "skipped: no Julia golden vectors".
#
# Cases: every BASELINE.json configuration at a small n (inputs are generated HERE with a fixed seed and written next to the
# outputs, so no RNG has to agree between languages), the MaternP control ladder of test/stationary.jl:53-69
# (r^2 = 10^(1:16) eps and 0, p = 0..3), Float32 / promoted-eltype cases, Lengthscale and ARD, 5-argument mul!, and the
# derivative kernels including the coincident-point edge of Exp / MaternP(0) / MaternP(1) (reference: NaN / Inf).
using LinearAlgebra
using Random
using CovarianceFunctions
using CovarianceFunctions: EQ, Exp, RQ, MaternP, Dot, Lengthscale, ARD, GradientKernel, ValueGradientKernel, gramian

outdir = length(ARGS) >= 1 ? ARGS[1] : joinpath(@__DIR__, "..", "tests", "golden", "julia_v1")
mkpath(outdir)
rng = MersenneTwister(20240521)

entries = String[]
function put(name::String, A::AbstractArray)
    B = collect(A)
    open(joinpath(outdir, name * ".bin"), "w") do io
        write(io, htol.(reinterpret(eltype(B) == Float32 ? UInt32 : UInt64, vec(B))))
    end
    return "\"$(name)\": {\"file\": \"$(name).bin\", \"dtype\": \"$(eltype(B) == Float32 ? "f32" : "f64")\", \"shape\": [$(join(size(B), ", "))]}"
end
function case(name; kernel::String, op::String, params = "", arrays...)
    files = [put(name * "_" * String(k), v) for (k, v) in arrays]
    push!(entries, "    {\"name\": \"$(name)\", \"kernel\": \"$(kernel)\", \"op\": \"$(op)\"$(params == "" ? "" : ", " * params), \"arrays\": {" * join(files, ", ") * "}}")
end

vecs(X) = [X[:, i] for i in 1:size(X, 2)]   # Vector{Vector{T}}: the README's data layout

kernels = Dict(
    "EQ" => EQ(), "Exp" => Exp(), "RQ(2)" => RQ(2), "RQ(1.5)" => RQ(1.5), "MaternP(0)" => MaternP(0), "MaternP(1)" => MaternP(1),
    "MaternP(2)" => MaternP(2), "MaternP(3)" => MaternP(3), "Dot^3" => Dot()^3, "1/2*RQ(2)+Dot()^2" => 1 / 2 * RQ(2) + Dot()^2,
    "Lengthscale(EQ,0.7)" => Lengthscale(EQ(), 0.7), "2.5*Lengthscale(MaternP(2),1.3)" => 2.5 * Lengthscale(MaternP(2), 1.3),
    "EQ+1/2*RQ(2)" => EQ() + 1 / 2 * RQ(2), "1/2*EQ+MaternP(2)*RQ(2)" => 1 / 2 * EQ() + MaternP(2) * RQ(2))

# ---- config 1 / 2 and friends: mul!(b, K, a), 3- and 5-argument forms, rectangular, dense --------------------------------------------
for (tag, kname, d, n, m) in (("c1", "MaternP(2)", 3, 512, 512), ("c2", "EQ", 3, 512, 512), ("exp", "Exp", 2, 300, 211), ("rq", "RQ(1.5)", 5, 257, 130),
                              ("ls", "Lengthscale(EQ,0.7)", 3, 400, 400), ("lsm", "2.5*Lengthscale(MaternP(2),1.3)", 4, 300, 300),
                              ("comp", "1/2*EQ+MaternP(2)*RQ(2)", 3, 300, 300), ("c5op", "MaternP(2)", 8, 384, 384))
    k = kernels[kname]
    X = randn(rng, d, n) / (d > 3 ? sqrt(d) : 1)
    Y = n == m ? X : randn(rng, d, m)
    K = n == m ? gramian(k, vecs(X)) : gramian(k, vecs(X), vecs(Y))
    a = randn(rng, m); y0 = randn(rng, n)
    b = zeros(n); mul!(b, K, a)
    y = copy(y0); mul!(y, K, a, -0.7, 1.9)
    case(tag * "_mul"; kernel = kname, op = "mul_vec", params = "\"alpha\": -0.7, \"beta\": 1.9", X = X, Y = Y, a = a, y0 = y0, b = b, y = y,
         M = Matrix(K)[1:min(n, 64), 1:min(m, 64)])
end

# ---- config 3: multi-RHS mul!(B, K, A) with the sum kernel -----------------------------------------------------------------------------
let d = 32, n = 256, p = 8, kname = "1/2*RQ(2)+Dot()^2"
    X = randn(rng, d, n) / sqrt(d); A = randn(rng, n, p)
    K = gramian(kernels[kname], vecs(X))
    B = zeros(n, p); mul!(B, K, A)
    case("c3_mulmat"; kernel = kname, op = "mul_mat", X = X, A = A, B = B)
end

# ---- Float32 data, and Float32 data under a kernel with a Float64 parameter (Gramian eltype promotes to Float64) -------------------------
let d = 3, n = 400
    X = randn(rng, Float32, d, n); a32 = randn(rng, Float32, n); a64 = randn(rng, n)
    K32 = gramian(EQ(), vecs(X)); b32 = zeros(Float32, n); mul!(b32, K32, a32)
    case("f32_eq"; kernel = "EQ", op = "mul_vec", params = "\"eltype\": \"$(eltype(K32))\"", X = X, a = a32, b = b32)
    Km = gramian(MaternP(2), vecs(X)); bm = zeros(eltype(Km), n); mul!(bm, Km, eltype(Km).(a32))
    case("f32_matern2"; kernel = "MaternP(2)", op = "mul_vec", params = "\"eltype\": \"$(eltype(Km))\"", X = X, a = a32, b = bm)
    Kp = gramian(0.5 * RQ(2), vecs(X)); bp = zeros(eltype(Kp), n); mul!(bp, Kp, a64)
    case("f32pts_halfrq"; kernel = "0.5*RQ(2)", op = "mul_vec", params = "\"eltype\": \"$(eltype(Kp))\"", X = X, a = a64, b = bp)
end

# ---- ARD (src/transformation.jl:42-46, test/stationary.jl:132-154) ----------------------------------------------------------------------
let d = 3, n = 300
    X = randn(rng, d, n); a = randn(rng, n); l = exp.(randn(rng, d))
    for kname in ("EQ", "MaternP(2)", "RQ(2)")
        K = gramian(ARD(kernels[kname], l), vecs(X)); b = zeros(n); mul!(b, K, a)
        case("ard_" * replace(kname, r"[()]" => ""); kernel = "ARD(" * kname * ")", op = "mul_vec", X = X, a = a, l = l, b = b)
    end
end

# ---- MaternP control ladder (test/stationary.jl:53-69): values at r^2 = 10^(1:16) eps and 0; first and second derivative in r^2 ----------
let r2 = vcat(0.0, 10.0 .^ (1:16) * eps())
    for p in 0:3
        k = MaternP(p)
        vals = k.(r2)
        d1 = similar(r2); d2 = similar(r2)
        for (i, r) in enumerate(r2)
            k1, k2 = CovarianceFunctions.derivative_laplacian(k, r)      # src/gradient.jl:589-592
            d1[i] = k1; d2[i] = k2
        end
        case("ladder_maternp$(p)"; kernel = "MaternP($(p))", op = "ladder", r2 = r2, k = vals, k1 = d1, k2 = d2)
    end
end

# ---- config 4 and the derivative kernels: BlockFactorization mul!, 5-argument form, ValueGradientKernel --------------------------------------
for (tag, kname, d, n) in (("c4", "EQ", 16, 64), ("gm2", "MaternP(2)", 5, 80), ("gm3", "MaternP(3)", 3, 90), ("gdot", "Dot^3", 4, 70), ("gcomp", "EQ+1/2*RQ(2)", 6, 60))
    k = kernels[kname]
    X = randn(rng, d, n) / sqrt(d)
    G = gramian(GradientKernel(k), vecs(X))
    a = randn(rng, d * n); y0 = randn(rng, d * n)
    b = G * a
    y = copy(y0); mul!(y, G, a, 1.5, -1.0)
    case(tag * "_grad"; kernel = kname, op = "gradient_mul", params = "\"alpha\": 1.5, \"beta\": -1.0", X = X, a = a, y0 = y0, b = b, y = y)
    V = gramian(ValueGradientKernel(k), vecs(X))
    av = randn(rng, (d + 1) * n)
    case(tag * "_valgrad"; kernel = kname, op = "value_gradient_mul", X = X, a = av, b = V * av)
end
# coincident points under kernels that are not differentiable at r = 0: the reference's own answer (NaN / Inf) is what is recorded
for (tag, kname) in (("gexp", "Exp"), ("gm0", "MaternP(0)"), ("gm1", "MaternP(1)"))
    d, n = 3, 24
    X = randn(rng, d, n) / sqrt(d)
    G = gramian(GradientKernel(kernels[kname]), vecs(X))
    a = randn(rng, d * n)
    case(tag * "_grad_coincident"; kernel = kname, op = "gradient_mul", X = X, a = a, b = G * a)
end

# ---- config 5: (sigma^2 I + K) \ y through LazyMatrixSum -> cg! (src/gramian.jl:55-60, src/lazy_linear_algebra.jl:135-144) -------------------
let d = 8, n = 256, s2 = 1e-2
    X = randn(rng, d, n) / sqrt(d); y = randn(rng, n)
    K = gramian(MaternP(2), vecs(X))
    A = Diagonal(fill(s2, n)) + K
    x = A \ y
    case("c5_solve"; kernel = "MaternP(2)", op = "solve", params = "\"sigma2\": $(s2)", X = X, y = y, x = x, r = y - A * x)
    Xg = randn(rng, 4, 48) / 2
    G = gramian(GradientKernel(MaternP(2)), vecs(Xg)); ag = randn(rng, 4 * 48); rhs = G * ag
    xs = G \ rhs
    case("grad_solve"; kernel = "MaternP(2)", op = "gradient_solve", X = Xg, rhs = rhs, x = xs)
end

open(joinpath(outdir, "manifest.json"), "w") do io
    println(io, "{\n  \"version\": 1,\n  \"julia\": \"$(VERSION)\",\n  \"threads\": $(Threads.nthreads()),\n  \"reference\": \"CovarianceFunctions.jl $(pkgversion(CovarianceFunctions))\",")
    println(io, "  \"cases\": [\n" * join(entries, ",\n") * "\n  ]\n}")
end
println("wrote $(length(entries)) cases to $(outdir)")
