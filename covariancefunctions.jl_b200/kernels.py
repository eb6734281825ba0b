"""Host-side mirror of the reference's kernel types for the lazy-Gramian path.

Same names, argument meaning and error behaviour as CovarianceFunctions.jl v0.3.5 (file:line citations are into
/root/reference/src).  Kernel objects are small immutable descriptions; `k.program()` flattens the tree into the postfix
``cf_knode_t`` program the C ABI consumes (include/covfn_b200.h).  `k(x, y)` evaluates a single pair on the host in plain
numpy with the reference's formulas -- it exists so that user code and tests can spot-check entries the way the
reference's tests do (`G[i, j] ~ k(x[i], y[j])`, test/gramian.jl:75-80); it is never used by the multiply path.
"""
from __future__ import annotations

import math
from typing import Sequence

import numpy as np

from ._lib import DomainError, UnsupportedKernel

# cf_op numbering (include/covfn_b200.h)
OP_EQ, OP_EXP, OP_RQ, OP_MATERNP, OP_DOT, OP_CONST, OP_SUM, OP_PROD, OP_POW, OP_LENGTHSCALE, OP_ARDSCALE, OP_ARD = range(1, 13)


# ---- input traits (properties.jl:31-37) ---------------------------------------------------------------------------
class InputTrait:
    def __eq__(self, other):
        return type(self) is type(other)

    def __hash__(self):
        return hash(type(self).__name__)

    def __repr__(self):
        return type(self).__name__ + "()"


class GenericInput(InputTrait):
    pass


class IsotropicInput(InputTrait):
    pass


class DotProductInput(InputTrait):
    pass


def _is_int(v) -> bool:
    return isinstance(v, (int, np.integer)) and not isinstance(v, bool)


class AbstractKernel:
    """AbstractKernel{T} (CovarianceFunctions.jl:32-35).  `eltype` is None for Union{} (parameter-free kernels)."""

    eltype = None

    # -- algebra (algebra.jl:23-25, 45-47, 63) --
    def __mul__(self, other):
        if isinstance(other, AbstractKernel):
            return Product((self, other))
        if isinstance(other, (int, float, np.integer, np.floating)):
            return Product((Constant(other), self))  # c*k and k*c are both Constant(c) * k (algebra.jl:24-25)
        return NotImplemented

    def __rmul__(self, other):
        if isinstance(other, (int, float, np.integer, np.floating)):
            return Product((Constant(other), self))
        return NotImplemented

    def __add__(self, other):
        if isinstance(other, AbstractKernel):
            return Sum((self, other))
        if isinstance(other, (int, float, np.integer, np.floating)):
            return Sum((self, Constant(other)))  # k + c and c + k are both k + Constant(c) (algebra.jl:46-47)
        return NotImplemented

    def __radd__(self, other):
        if isinstance(other, (int, float, np.integer, np.floating)):
            return Sum((self, Constant(other)))
        return NotImplemented

    def __pow__(self, p):
        if not _is_int(p):
            raise TypeError("kernel powers must be Int (Power.p::Int, algebra.jl:52)")
        return Power(self, int(p))

    # -- evaluation --
    def __call__(self, x, y=None):
        if y is None:
            return self._of_scalar(x)
        x = np.atleast_1d(np.asarray(x, dtype=np.float64))
        y = np.atleast_1d(np.asarray(y, dtype=np.float64))
        if x.shape != y.shape:
            from ._lib import DimensionMismatch

            raise DimensionMismatch(f"inputs have to have the same length: {x.size}, {y.size}")  # util.jl:41
        return self._of_pair(x, y)

    def _of_scalar(self, t):
        raise TypeError(f"{type(self).__name__} has no one-argument form")

    def _of_pair(self, x, y):
        raise NotImplementedError

    def program(self):
        """postfix list of (op, iparam, fparam)"""
        raise NotImplementedError

    def input_trait(self) -> InputTrait:  # properties.jl:39
        return GenericInput()


def input_trait(k) -> InputTrait:
    return k.input_trait() if isinstance(k, AbstractKernel) else GenericInput()


class IsotropicKernel(AbstractKernel):
    """k(x, y) = k(euclidean2(x, y)) (stationary.jl:9); one-argument form takes r2 (stationary.jl:6-10)."""

    def _of_pair(self, x, y):
        d = x - y
        return self._of_scalar(float(np.dot(d, d)))

    def input_trait(self):
        return IsotropicInput()  # properties.jl:41


class Constant(IsotropicKernel):
    def __init__(self, c, check: bool = True):  # stationary.jl:15-23
        if check and not (c >= 0):
            raise DomainError(f"Constant is not positive semi-definite: {c}")
        self.c = c
        self.eltype = type(c)

    def _of_scalar(self, r2):
        return self.c

    def _of_pair(self, x, y):
        return self.c

    def program(self):
        return [(OP_CONST, 1 if _is_int(self.c) else 0, float(self.c))]

    def __repr__(self):
        return f"Constant({self.c})"


class ExponentiatedQuadratic(IsotropicKernel):
    def _of_scalar(self, r2):  # stationary.jl:42
        return math.exp(-r2 / 2)

    def program(self):
        return [(OP_EQ, 0, 0.0)]

    def __repr__(self):
        return "EQ()"


EQ = ExponentiatedQuadratic


class RationalQuadratic(IsotropicKernel):
    def __init__(self, alpha):  # stationary.jl:45-51
        if not (0 < alpha):
            raise DomainError("α not positive")
        self.alpha = alpha
        self.eltype = type(alpha)

    def _of_scalar(self, r2):  # stationary.jl:53
        return (1 + r2 / (2 * self.alpha)) ** (-self.alpha)

    def program(self):
        return [(OP_RQ, 1 if _is_int(self.alpha) else 0, float(self.alpha))]

    def __repr__(self):
        return f"RQ({self.alpha})"


RQ = RationalQuadratic


class Exponential(IsotropicKernel):
    def _of_scalar(self, r2):  # stationary.jl:60
        return math.exp(-math.sqrt(r2))

    def program(self):
        return [(OP_EXP, 0, 0.0)]

    def __repr__(self):
        return "Exp()"


Exp = Exponential


def MaternP_coefficients(p: int):
    """stationary.jl:184-191"""
    c = [math.comb(p, i) * (math.factorial(p + i) // math.factorial(p)) for i in range(1, p + 1)]
    return [float(v) for v in reversed(c)]


def MaternP_derivatives_at_zero(p: int):
    """stationary.jl:172-182 (SymEngine there); closed form d_i = (nu/2)^i / prod_{m<=i}(m - nu), nu = p + 1/2,
    checked against symbolic differentiation in tests/test_oracle_relations.py."""
    from fractions import Fraction

    nu = Fraction(2 * p + 1, 2)
    out, num, den = [], Fraction(1), Fraction(1)
    for i in range(1, p + 1):
        num *= nu / 2
        den *= i - nu
        out.append(float(num / den))
    return out


class MaternP(IsotropicKernel):
    def __init__(self, p):  # stationary.jl:123-131
        if isinstance(p, Matern):
            p = int(math.floor(p.nu))
        if not _is_int(p):
            raise TypeError("MaternP(p::Int)")
        if 0 > p:
            raise DomainError(f"p = {p} is negative")
        self.p = int(p)
        self.derivatives = MaternP_derivatives_at_zero(self.p)
        self.coefficients = MaternP_coefficients(self.p)

    def _of_scalar(self, r2):  # stationary.jl:134-158
        p = self.p
        r2 = float(r2)
        taylor_bound = 0.0 if p == 0 else np.finfo(np.float64).eps ** (1 / p)
        if r2 < taylor_bound:
            y, r2i = 1.0, r2
            for i in range(1, p + 1):
                y += self.derivatives[i - 1] * r2i / math.factorial(i)
                r2i *= r2
            return y
        y = 0.0
        r = math.sqrt((2 * p + 1) * r2)
        ri = 1.0
        for i in range(1, p + 1):
            y += self.coefficients[i - 1] * ri
            ri *= 2 * r
        y += ri
        return y * (math.exp(-r) / (math.factorial(2 * p) // math.factorial(p)))

    def program(self):
        return [(OP_MATERNP, self.p, 0.0)]

    def __repr__(self):
        return f"MaternP({self.p})"


class Matern(IsotropicKernel):
    """Matern(nu) exists in the reference (stationary.jl:87-114, needs Bessel K); here it only projects to MaternP."""

    def __init__(self, nu):
        if not (0 < nu):
            raise DomainError(f"ν = {nu} is negative")
        self.nu = nu

    def program(self):
        raise UnsupportedKernel("Matern(ν) with real ν needs besselk; use MaternP(p) (ν = p + 1/2)")


class Lengthscale(IsotropicKernel):
    def __init__(self, k, l):  # transformation.jl:6-16
        if not isinstance(k, IsotropicKernel):
            raise TypeError("Lengthscale(k::IsotropicKernel, l)")
        if np.ndim(l) != 0:
            from ._lib import DimensionMismatch

            raise DimensionMismatch("lengthscale l has to has length 1")
        if not (l > 0):
            raise DomainError(f"l = {l} is non-positive")
        self.k, self.l = k, float(l)

    def _of_scalar(self, r2):  # transformation.jl:19
        return self.k._of_scalar(r2 / self.l**2)

    def program(self):
        return self.k.program() + [(OP_LENGTHSCALE, 0, self.l)]

    def __repr__(self):
        return f"Lengthscale({self.k!r}, {self.l})"


class DotProductKernel(AbstractKernel):
    """k(x, y) = k(dot(x, y)) (mercer.jl:2-3)"""

    def _of_pair(self, x, y):
        return self._of_scalar(float(np.dot(x, y)))


class Dot(DotProductKernel):
    def _of_scalar(self, d):  # mercer.jl:9
        return d

    def program(self):
        return [(OP_DOT, 0, 0.0)]

    def input_trait(self):
        return DotProductInput()  # properties.jl:42

    def __repr__(self):
        return "Dot()"


def Line(sigma=0.0):  # mercer.jl:12
    return Dot() + sigma


def Polynomial(d: int, sigma=0.0):  # mercer.jl:13
    return Line(sigma) ** d


Poly = Polynomial


def sum_and_product_input_trait(args: Sequence[AbstractKernel]) -> InputTrait:
    """properties.jl:47-63: constants are ignored; mixed non-constant traits -> GenericInput"""
    nonconst = [k for k in args if not isinstance(k, Constant)]
    if not nonconst:
        return IsotropicInput()
    trait = input_trait(nonconst[0])
    for k in nonconst[1:]:
        if input_trait(k) != trait:
            return GenericInput()
    return trait


class _NAry(AbstractKernel):
    OP = None

    def __init__(self, args):
        args = tuple(args)
        if not args or not all(isinstance(a, AbstractKernel) for a in args):
            raise TypeError("arguments must be kernels")
        self.args = args
        self._trait = sum_and_product_input_trait(args)

    def input_trait(self):
        return self._trait  # properties.jl:45

    def program(self):
        out = []
        for a in self.args:
            out += a.program()
        return out + [(self.OP, len(self.args), 0.0)]


class Sum(_NAry):
    OP = OP_SUM

    def _of_pair(self, x, y):  # algebra.jl:40
        return sum(k._of_pair(x, y) for k in self.args)

    def _of_scalar(self, t):  # algebra.jl:39
        return sum(k._of_scalar(t) for k in self.args)

    def __add__(self, other):  # +(k::AbstractKernel...) = Sum(k) is n-ary in Julia; left-assoc nesting gives the same value order
        return super().__add__(other)

    def __repr__(self):
        return "(" + " + ".join(map(repr, self.args)) + ")"


class Product(_NAry):
    OP = OP_PROD

    def _of_pair(self, x, y):  # algebra.jl:17
        return math.prod(k._of_pair(x, y) for k in self.args)

    def _of_scalar(self, t):  # algebra.jl:16
        return math.prod(k._of_scalar(t) for k in self.args)

    def __repr__(self):
        return "(" + " * ".join(map(repr, self.args)) + ")"


class Power(AbstractKernel):
    def __init__(self, k, p: int):  # algebra.jl:50-60
        if not isinstance(k, AbstractKernel) or not _is_int(p):
            raise TypeError("Power(k, p::Int)")
        self.k, self.p = k, int(p)

    def _of_pair(self, x, y):  # algebra.jl:62
        return self.k._of_pair(x, y) ** self.p

    def _of_scalar(self, t):  # algebra.jl:61
        return self.k._of_scalar(t) ** self.p

    def input_trait(self):
        return input_trait(self.k)  # properties.jl:43

    def program(self):
        return self.k.program() + [(OP_POW, self.p, 0.0)]

    def __repr__(self):
        return f"{self.k!r}^{self.p}"


class GradientKernel:
    """GradientKernel(k) (gradient.jl:7-24): the d x d matrix-valued kernel of gradient observations.  Captures
    input_trait(k) at construction; the IsotropicInput and DotProductInput elements are lowered to the device
    (gradient.jl:83-92, 107-115); GenericInput kernels are not (dense ForwardDiff fallback in the reference)."""

    block_extra = 0  # block size is d + block_extra

    def __init__(self, k):
        if not isinstance(k, AbstractKernel):
            raise TypeError(f"{type(self).__name__}(k::AbstractKernel)")
        self.k = k
        self._trait = input_trait(k)

    def input_trait(self):
        return self._trait  # gradient.jl:16

    def program(self):
        return self.k.program()

    def __repr__(self):
        return f"GradientKernel({self.k!r})"


class ValueGradientKernel(GradientKernel):
    """ValueGradientKernel(k) (gradient.jl:400-474): (d+1) x (d+1) blocks covering the value and its gradient;
    entry 0 of every block is the value observation (DerivativeKernelElement, gradient.jl:217-239)."""

    block_extra = 1

    def __repr__(self):
        return f"ValueGradientKernel({self.k!r})"


class ARD(AbstractKernel):
    """ARD(k, l::AbstractVector) (transformation.jl:42-45): Normed(k, tau -> enorm2(Diagonal(inv.(l)), tau)), i.e.
    k(sum_c tau_c^2 / l_c), a StationaryKernel (transformation.jl:25).  Lowered to the ABI's ARD node
    (<k> ARDSCALE(l_1) ... ARDSCALE(l_d) ARD(d)); the library applies the metric to the points on the device.
    ARD(k, l::Real) is Lengthscale(k, l) (transformation.jl:46)."""

    def __new__(cls, k, l):
        if np.ndim(l) == 0:
            return Lengthscale(k, l)
        return super().__new__(cls)

    def __init__(self, k, l):
        if not isinstance(k, AbstractKernel):
            raise TypeError("ARD(k, l): k must be a kernel")
        l = np.asarray(l, dtype=np.float64)
        if not np.all(l > 0):
            raise DomainError(f"l = {l} is non-positive")
        self.k, self.l = k, l
        self.eltype = k.eltype

    def _of_pair(self, x, y):
        tau = x - y
        return self.k._of_scalar(float(np.sum(tau * (1.0 / self.l) * tau)))

    def _of_scalar(self, t):  # (m::Normed)(tau) = m.k(m.n2(tau)) (transformation.jl:38)
        tau = np.atleast_1d(np.asarray(t, dtype=np.float64))
        return self.k._of_scalar(float(np.sum(tau * (1.0 / self.l) * tau)))

    def program(self):
        return self.k.program() + [(OP_ARDSCALE, 0, float(lc)) for lc in self.l] + [(OP_ARD, int(self.l.size), 0.0)]

    def __repr__(self):
        return f"ARD({self.k!r}, {self.l.tolist()})"
