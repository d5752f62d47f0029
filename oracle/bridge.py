"""Test-side glue between the product's set objects and the oracle's cone classes.

TEST INFRASTRUCTURE (like the rest of oracle/): imported by tests/, __graft_entry__.smoke() and the CPU-baseline /
reference legs of bench.py only.  The product package (cosmo.jl_b200/) never imports anything from here."""
from __future__ import annotations

from . import cosmo_oracle as O


def to_oracle_cones(sets):
    """Translate cosmo_b200 set objects (cosmo.jl_b200/model.py) into the oracle's cone classes."""
    import cosmo_b200
    M = cosmo_b200.model
    out = []
    for S in sets:
        if isinstance(S, M.ZeroSet):
            out.append(O.ZeroSet(S.dim))
        elif isinstance(S, M.Nonnegatives):
            out.append(O.Nonnegatives(S.dim))
        elif isinstance(S, M.Box):
            out.append(O.Box(S.l, S.u))
        elif isinstance(S, M.SecondOrderCone):
            out.append(O.SecondOrderCone(S.dim))
        elif isinstance(S, M.PsdCone):
            out.append(O.PsdCone(S.dim))
        elif isinstance(S, M.PsdConeTriangle):
            out.append(O.PsdConeTriangle(S.dim))
        elif isinstance(S, M.ComplexPsdConeTriangle):
            out.append(O.ComplexPsdConeTriangle(S.dim))
        elif isinstance(S, M.DualExponentialCone):
            out.append(O.DualExponentialCone(3, S.MAX_ITER, S.TOL))
        elif isinstance(S, M.ExponentialCone):
            out.append(O.ExponentialCone(3, S.MAX_ITER, S.TOL))
        elif isinstance(S, M.DualPowerCone):
            out.append(O.DualPowerCone(S.alpha, S.MAX_ITER, S.TOL))
        elif isinstance(S, M.PowerCone):
            out.append(O.PowerCone(S.alpha, S.MAX_ITER, S.TOL))
        else:
            raise TypeError(S)
    return out
