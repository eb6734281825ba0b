"""covariancefunctions.jl_b200 -- B200-native lazy-Gramian multiply behind CovarianceFunctions.jl's API.

The directory name is not a valid Python identifier; import it as ``covfn_b200`` (the loader module at the
repository root registers this package under that name).

Public surface (mirrors the reference, /root/reference/src/CovarianceFunctions.jl:3-5):
    gramian, Gramian, mul_ (mul!), kernels EQ / Exp / RQ / MaternP / Dot / Line / Poly / Constant and their algebra,
    Lengthscale, GradientKernel, LazyMatrixSum via ``sigma2 * I(n) + G``.
"""
from ._lib import (CovFnError, CudaError, DimensionMismatch, DomainError, UnsupportedKernel, check, device_count, init, lib,
                   LIB_PATH, SYMBOLS)
from .kernels import (ARD, EQ, RQ, AbstractKernel, Constant, Dot, DotProductInput, Exp, Exponential, ExponentiatedQuadratic,
                      GenericInput, GradientKernel, IsotropicInput, IsotropicKernel, Lengthscale, Line, Matern, MaternP,
                      Poly, Polynomial, Power, Product, RationalQuadratic, Sum, ValueGradientKernel, input_trait)
from .gramian import Diagonal, Gramian, I, LazyMatrixSum, gramian, jit_check, jit_stats, mul_, mul_collective_device, peak_probe

__all__ = [n for n in dir() if not n.startswith("_")]
