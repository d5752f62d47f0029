"""A stand-in for cosmo_b200.engine.Engine that answers with the CPU oracle (tests only).

It exists to exercise, without a GPU, the Python side of everything that normally talks to the CUDA engine: the host
glue of Model (scaling, decomposition, warm starts) and the bodies of the GPU tests themselves (so a typo in a GPU
test is found by the CPU run, not at the next GPU run).  It implements the part of the call surface those users need:
ctor, update_settings, warm_start, update_qb, project, solve, w, rho_vec, scaling, close."""
import numpy as np
import scipy.sparse as sp

from cosmo_b200 import engine as E
from oracle import cosmo_oracle as O

_PLAIN = {E.ZERO: O.ZeroSet, E.NONNEG: O.Nonnegatives, E.SOC: O.SecondOrderCone, E.PSD_SQUARE: O.PsdCone,
          E.PSD_TRIANGLE: O.PsdConeTriangle, E.PSD_TRIANGLE_COMPLEX: O.ComplexPsdConeTriangle}


def cones_from_tuples(sets):
    out = []
    for t in sets:
        typ, dim, l, u = t[:4]
        extra = t[4] if len(t) > 4 and t[4] else {}
        if typ == E.BOX:
            out.append(O.Box(l, u))
        elif typ == E.EXP:
            out.append(O.ExponentialCone())
        elif typ == E.DUAL_EXP:
            out.append(O.DualExponentialCone())
        elif typ == E.POW:
            out.append(O.PowerCone(extra["alpha"]))
        elif typ == E.DUAL_POW:
            out.append(O.DualPowerCone(extra["alpha"]))
        else:
            out.append(_PLAIN[typ](dim))
    return out


class OracleEngine:
    instances = []

    def __init__(self, P, q, A, b, sets, settings=None, D=None, E=None, c=1.0, dtype=np.float64, device=0, julia_indexing=True,
                 equilibrate=False):
        self.P, self.q = sp.csc_matrix(P), np.array(q, dtype=float)
        self.A, self.b = sp.csc_matrix(A), np.array(b, dtype=float)
        self.m, self.n = self.A.shape
        self.cones = cones_from_tuples(sets)
        self.st = settings
        self.scaled = D is not None
        self._scal = (np.ones(self.n), np.ones(self.m), 1.0) if D is None else (np.array(D), np.array(E), float(c))
        if equilibrate and D is None and settings is not None and settings.scaling != 0:
            # like the engine: unscaled data + scaling requested -> equilibrate here (scale_ruiz!)
            ost = O.Settings(scaling=int(settings.scaling), MIN_SCALING=settings.MIN_SCALING)
            Ps, qs, As, bs, cones, sm = O.scale_ruiz(self.P, self.q, self.A, self.b, self.cones, ost)
            self.P, self.q, self.A, self.b, self.cones = sp.csc_matrix(Ps), qs, sp.csc_matrix(As), bs, cones
            self._scal = (sm.D, sm.E, sm.c)
            self.scaled = True
        self._w = self._rho = self._warm = None
        OracleEngine.instances.append(self)

    def scaling(self):
        return self._scal

    def psd_stats(self):
        return {"tc_projections": 0, "tc_fallbacks": 0}

    def update_settings(self, st):
        self.st = st

    def warm_start(self, x, s, mu):
        self._warm = (np.array(x), np.array(s), np.array(mu))

    def update_qb(self, q, b):
        if q is not None:
            self.q = np.array(q, dtype=float)
        if b is not None:
            self.b = np.array(b, dtype=float)

    def project(self, ws):
        out = np.array(ws, dtype=float).copy()
        O.project(out, self.cones)
        return out

    def solve(self):
        st = self.st
        ost = O.Settings(scaling=0, kkt_solver="cg", eps_abs=st.eps_abs, eps_rel=st.eps_rel, max_iter=st.max_iter, rho=st.rho,
                         check_termination=st.check_termination, check_infeasibility=st.check_infeasibility,
                         adaptive_rho=bool(st.adaptive_rho), adaptive_rho_interval=st.adaptive_rho_interval,
                         adaptive_rho_max_adaptions=st.adaptive_rho_max_adaptions,
                         accelerator="anderson" if st.accelerator == E.ACC_ANDERSON else "empty",
                         accelerator_mem=st.accelerator_mem, safeguard=bool(st.safeguard), safeguard_tol=st.safeguard_tol)
        r = O.solve(self.P, self.q, self.A, self.b, self.cones, ost)
        self._w, self._rho = r.w, r.rho_vec
        out = E.SolveOutput()
        out.x, out.s, out.mu = r.x, r.s, -r.y
        out.obj_val, out.iter, out.safeguarding_iter, out.status = r.obj_val, r.iter, r.safeguarding_iter, r.status
        out.r_prim, out.r_dual, out.max_norm_prim, out.max_norm_dual = r.info.r_prim, r.info.r_dual, 0.0, 0.0
        out.rho, out.rho_updates, out.times = 0.1, list(r.info.rho_updates), {"iter_time_device": 0.0}
        out.kkt_inner_iterations = out.kkt_multiplications = out.kernel_launches = 0
        return out

    def w(self):
        return self._w

    def rho_vec(self):
        return self._rho

    def close(self):
        pass
