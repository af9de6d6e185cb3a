"""TEST INFRASTRUCTURE — NOT PRODUCT CODE.

ctypes front-end of oracle/liboracle.so (oracle.c) plus small classes that assemble the oracle's
operators from oracle/setup_np.py objects, mirroring the reference's class names
(StiffnessMatrix, MassMatrix, FaceMassMatrix, Helmholtz, DDH). Used only by tests/, smoke()
and bench.py's cpu_baseline / --impl reference legs."""
import ctypes as C
import os
import subprocess
import numpy as np
from . import setup_np as S

_here = os.path.dirname(os.path.abspath(__file__))
_lib = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(_here, "liboracle.so")
        if not os.path.exists(path):
            subprocess.check_call(["make", "-C", _here, "oracle"])
        _lib = C.CDLL(path)
        _lib.orc_max_threads.restype = C.c_int
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f64(a):
    return np.ascontiguousarray(a, np.float64)


def _i32(a):
    return np.ascontiguousarray(a, np.int32)


def set_threads(n):
    lib().orc_set_threads(C.c_int(n))


def max_threads():
    return lib().orc_max_threads()


class H1:
    """mesh + basis + global-to-local map (reference: H1Space)."""

    def __init__(self, mesh, nb):
        self.mesh = mesh
        self.basis = S.Basis(nb)
        self.nb = nb
        I3, self.ndof, self.xy = S.h1space(mesh, self.basis)
        self.I = _i32(I3.reshape(-1))
        self.nel = mesh.n_elem
        self.corners = _f64(mesh.corners.reshape(-1))  # (2,4,nel) column-major

    @classmethod
    def from_arrays(cls, nb, I, xy, corners, boundary_faces_I=None):
        """oracle space over externally supplied index data (bench.py's CPU baseline at sizes where the pure-Python
        setup above would take minutes): I (nel*nb*nb) int32, xy (ndof,2), corners (nel,4,2)."""
        self = cls.__new__(cls)
        self.mesh = None
        self.basis = S.Basis(nb)
        self.nb = nb
        self.I = _i32(np.asarray(I).reshape(-1))
        self.xy = np.asarray(xy, float)
        self.ndof = len(self.xy)
        self.corners = _f64(np.asarray(corners).reshape(-1))
        self.nel = len(self.corners) // 8
        return self

    def measures(self, xq):
        nq = len(xq)
        out = np.zeros(nq * nq * self.nel)
        lib().orc_element_measures(C.c_int64(self.nel), C.c_int(nq), _p(_f64(xq)), _p(self.corners), _p(out))
        return out

    def jacobians(self, xq):
        nq = len(xq)
        out = np.zeros(4 * nq * nq * self.nel)
        lib().orc_element_jacobians(C.c_int64(self.nel), C.c_int(nq), _p(_f64(xq)), _p(self.corners), _p(out))
        return out

    def coordinates(self, xq):
        nq = len(xq)
        out = np.zeros(2 * nq * nq * self.nel)
        lib().orc_element_coordinates(C.c_int64(self.nel), C.c_int(nq), _p(_f64(xq)), _p(self.corners), _p(out))
        return out


class StiffnessMatrix:
    """source/StiffnessMatrix.cpp:40-81 (ctor), :186-211 (action)."""

    def __init__(self, fem, nq=None):
        self.fem = fem
        nb = fem.nb
        self.nq = nq = (nb + 1) if nq is None else nq
        x, w = S.gauss_legendre(nq)
        self.P = _f64(fem.basis.eval(x).T.reshape(-1))  # column-major (nq, nb)
        self.D = _f64(fem.basis.deriv(x).T.reshape(-1))
        J = fem.jacobians(x)
        self.G = np.zeros(3 * nq * nq * fem.nel)
        lib().orc_stiffness_setup(C.c_int64(fem.nel), C.c_int(nq), _p(_f64(w)), _p(J), _p(self.G))

    def action(self, x, y=None, c=1.0):
        """y <- (y if given else 0) + c*S*x ; returns y."""
        fem = self.fem
        y = np.zeros(fem.ndof) if y is None else y
        lib().orc_stiffness_action(C.c_int64(fem.nel), C.c_int(self.nq), C.c_int(fem.nb), _p(self.P), _p(self.D),
                                   _p(self.G), _p(fem.I), C.c_double(c), _p(_f64(x)), _p(y))
        return y


class MassMatrix:
    """source/MassMatrix.cpp:69-135 (ctors: nq = nb+1 unweighted, 1+3nb/2+1 weighted), :213-239."""

    def __init__(self, fem, a=None, nq=None):
        self.fem = fem
        nb = fem.nb
        if nq is None:
            nq = nb + 1 if a is None else 1 + (3 * nb) // 2 + 1
        self.nq = nq
        x, w = S.gauss_legendre(nq)
        self.P = _f64(fem.basis.eval(x).T.reshape(-1))
        detJ = fem.measures(x)
        self.a = np.zeros(nq * nq * fem.nel)
        lib().orc_mass_setup(C.c_int64(fem.nel), C.c_int(nb), C.c_int(nq), _p(None if a is None else _f64(a)),
                             _p(detJ), _p(_f64(w)), _p(fem.I), _p(self.P), _p(self.a))

    def action(self, x, y=None, c=1.0):
        fem = self.fem
        y = np.zeros(fem.ndof) if y is None else y
        lib().orc_mass_action(C.c_int64(fem.nel), C.c_int(self.nq), C.c_int(fem.nb), _p(fem.I), _p(self.P),
                              _p(self.a), C.c_double(c), _p(_f64(x)), _p(y))
        return y


class DiagInvMassMatrix:
    """source/MassMatrix.cpp:241-334."""

    def __init__(self, fem, a=None):
        self.fem = fem
        detJ = fem.measures(fem.basis.x)
        self.p = np.zeros(fem.ndof)
        lib().orc_diag_inv_mass(C.c_int64(fem.ndof), C.c_int64(fem.nel), C.c_int(fem.nb),
                                _p(None if a is None else _f64(a)), _p(detJ), _p(_f64(fem.basis.w)), _p(fem.I), _p(self.p))

    def action(self, x, y=None, c=1.0):
        if y is None:
            return self.p * x
        y += c * self.p * x
        return y


class FaceSpace:
    """source/H1Space.cpp:129-219."""

    def __init__(self, fem, faces):
        self.fem = fem
        self.faces = np.asarray(faces, np.int32)
        fI, proj = S.facespace(fem.mesh, fem.I.reshape(fem.nel, fem.nb, fem.nb), fem.nb, self.faces)
        self.I = _i32(fI.reshape(-1))
        self.proj = _i32(proj)
        self.fdof = len(proj)
        self.nf = len(self.faces)

    @classmethod
    def from_arrays(cls, fem, fI, proj, face_meas):
        """oracle face space over externally supplied index data (see H1.from_arrays)."""
        self = cls.__new__(cls)
        self.fem = fem
        self.faces = None
        self.I = _i32(np.asarray(fI).reshape(-1))
        self.proj = _i32(proj)
        self.fdof = len(self.proj)
        self.nf = len(self.I) // fem.nb
        self._meas = np.asarray(face_meas, float)
        return self

    def measures(self, nq):
        # StraightEdge::measure is constant along the edge (include/Edge.hpp:134-137); (nq, nf) column-major
        if self.faces is None:
            return _f64(np.repeat(self._meas, nq))
        return _f64(np.repeat(self.fem.mesh.edge_meas[self.faces], nq))

    def restrict(self, x):
        return _f64(x)[self.proj].copy()

    def prolong(self, xf, y):
        y[self.proj] += xf
        return y

    def orth(self, x):
        x[self.proj] = 0.0
        return x


class FaceMassMatrix:
    """source/FaceMassMatrix.cpp:51-139 (ctors: nq = nb+1, weighted 1+3nb/2+1), :195-223."""

    def __init__(self, fs, a=None, nq=None):
        self.fs = fs
        nb = fs.fem.nb
        if nq is None:
            nq = nb + 1 if a is None else 1 + (3 * nb) // 2 + 1
        self.nq = nq
        x, w = S.gauss_legendre(nq)
        self.P = _f64(fs.fem.basis.eval(x).T.reshape(-1))
        self.a = np.zeros(nq * fs.nf)
        lib().orc_facemass_setup(C.c_int(fs.nf), C.c_int(nb), C.c_int(nq), _p(_f64(w)), _p(self.P), _p(fs.measures(nq)),
                                 _p(None if a is None else _f64(a)), _p(fs.I), _p(self.a))

    def action(self, x, y=None, c=1.0):
        fs = self.fs
        y = np.zeros(fs.fdof) if y is None else y
        lib().orc_facemass_action(C.c_int(fs.nf), C.c_int(fs.fem.nb), C.c_int(self.nq), _p(self.P), _p(self.a), _p(fs.I),
                                  C.c_double(c), _p(_f64(x)), _p(y))
        return y


class DiagInvFaceMassMatrix:
    """source/FaceMassMatrix.cpp:226-322."""

    def __init__(self, fs, a=None):
        nb = fs.fem.nb
        self.p = np.zeros(fs.fdof)
        lib().orc_diag_inv_facemass(C.c_int(fs.fdof), C.c_int(fs.nf), C.c_int(nb), _p(_f64(fs.fem.basis.w)),
                                    _p(fs.measures(nb)), _p(None if a is None else _f64(a)), _p(fs.I), _p(self.p))

    def action(self, x):
        return self.p * x


class Helmholtz:
    """examples/Helmholtz.hpp:10-80: [Au;Av] = [S u - w^2 M u - w H v ; -(S v - w^2 M v + w H u)]."""

    def __init__(self, omega, a2, af, fem, fs):
        self.omega, self.fem, self.fs = omega, fem, fs
        self.S = StiffnessMatrix(fem)
        self.M = MassMatrix(fem, a2)
        self.H = FaceMassMatrix(fs, af)

    def action(self, x):
        fem, fs = self.fem, self.fs
        y = np.zeros(2 * fem.ndof)
        xf = np.zeros(fs.fdof)
        yf = np.zeros(fs.fdof)
        lib().orc_helmholtz_action(
            C.c_int64(fem.ndof), C.c_int(fs.fdof), C.c_int64(fem.nel), C.c_int(fs.nf), C.c_int(fem.nb),
            C.c_int(self.S.nq), C.c_int(self.M.nq), C.c_int(self.H.nq),
            _p(fem.I), _p(self.S.P), _p(self.S.D), _p(self.S.G), _p(self.M.P), _p(self.M.a),
            _p(fs.I), _p(fs.proj), _p(self.H.P), _p(self.H.a), C.c_double(self.omega), _p(_f64(x)), _p(y), _p(xf), _p(yf))
        return y


class DDH:
    """source/DDH.cpp:323-695 on a uniform_rect(nx, ny) mesh; FP32 substructured operator."""

    def __init__(self, omega, h_a, fem, nx, ny, block=16):
        self.fem = fem
        self.omega = omega
        nb = fem.nb
        d = S.ddh_setup(omega, h_a, fem.mesh, fem.basis, fem.I.reshape(fem.nel, nb, nb), fem.ndof, nx, ny, block)
        self.d = d
        self.nb = nb
        self.mx = block * block
        en = d.en
        J = fem.jacobians(fem.basis.x)
        self.g = np.zeros(3 * self.mx * d.n_domains, np.float32)
        lib().orc_ddh_geom(C.c_int(d.n_domains), C.c_int(d.mx_elem_per_dom), C.c_int(nb), _p(_i32(en.s_elems)),
                           _p(_i32(en.elems.reshape(-1))), _p(_f64(fem.basis.w)), _p(J), _p(self.g))
        self.n_lambda = d.n_lambda
        self.size = 2 * d.n_lambda
        self._keep = dict(s_dof=_i32(en.s_dof), s_fdof=_i32(en.s_fdof), B=_i32(d.B.reshape(-1)), gI=_i32(d.gI.reshape(-1)),
                          sI=_i32(d.sI.reshape(-1)), D=np.ascontiguousarray(d.D.reshape(-1), np.float32),
                          m=np.ascontiguousarray(d.m.reshape(-1)), gmi=np.ascontiguousarray(d.gmi.reshape(-1)),
                          a=np.ascontiguousarray(d.a.reshape(-1)), H=np.ascontiguousarray(d.H.reshape(-1)))

    def _run(self, x, y, lam, upd):
        d, k = self.d, self._keep
        lib().orc_ddh_action(
            C.c_int(d.n_domains), C.c_int(self.nb), C.c_int(self.mx), C.c_int(d.mx_dof), C.c_int(d.mx_fdof),
            C.c_int64(self.fem.ndof), C.c_int64(d.n_lambda), C.c_int(d.nt), C.c_float(self.omega), C.c_float(d.dt),
            _p(k["s_dof"]), _p(k["s_fdof"]), _p(k["B"]), _p(k["gI"]), _p(k["sI"]), _p(k["D"]), _p(self.g), _p(k["m"]),
            _p(k["gmi"]), _p(k["a"]), _p(k["H"]), _p(d.wh_filter), _p(d.cs), _p(d.sn), _p(x), _p(y), _p(lam), _p(upd))

    def action(self, lam):
        """source/DDH.cpp:611-639: out = lambda - T(lambda). Slots nobody writes stay 0 (fresh buffer)."""
        lam = np.ascontiguousarray(lam, np.float32)
        upd = np.zeros(self.size, np.float32)
        self._run(None, None, lam, upd)
        return (np.float32(1.0) * lam + np.float32(-1.0) * upd).astype(np.float32)

    def rhs(self, f):
        """:641-667."""
        b = np.zeros(self.size, np.float32)
        self._run(_f64(f), None, None, b)
        return b

    def postprocess(self, lam, f):
        """:669-695."""
        y = np.zeros(2 * self.fem.ndof)
        self._run(_f64(f), y, np.ascontiguousarray(lam, np.float32), None)
        return y


def gmres(n, x, A, b, m, maxit, tol, P=None, dtype=np.float64):
    """source/gmres.cpp:91-251 (t_gmres + the left-preconditioned overload). A, P: callables v -> A v.
    Returns dict(success, num_iter, num_matvec, res_norm). x is updated in place."""
    sc = dtype
    if P is not None:  # :242-251
        A0 = A
        A = lambda v: P(A0(v))
        b = P(b)
    b = np.asarray(b, sc)
    nrm = lambda v: sc(np.sqrt(np.dot(v, v)))
    bnrm = nrm(b)
    V = np.zeros((m + 1, n), sc)
    H = np.zeros((m + 1, m), sc)
    sn = np.zeros(m, sc)
    cs = np.zeros(m, sc)
    out = dict(success=False, num_iter=0, num_matvec=0, res_norm=[])
    r = np.asarray(A(x), sc)
    out["num_matvec"] += 1
    r = (sc(1) * b + sc(-1) * r).astype(sc)
    r_nrm = nrm(r)
    out["res_norm"].append(float(r_nrm))
    tol = sc(tol)
    if r_nrm < tol * bnrm:
        out["success"] = True
        return out
    it = 1
    while it < maxit:
        V[0] = (sc(1) / r_nrm) * r
        eta = np.zeros(m + 1, sc)
        eta[0] = r_nrm
        k1 = 0
        for k in range(m):
            k1 = k + 1
            w = np.asarray(A(V[k]), sc).copy()
            out["num_matvec"] += 1
            for j in range(k1):
                H[j, k] = sc(np.dot(w, V[j]))
                w = (-H[j, k] * V[j] + sc(1) * w).astype(sc)
            H[k1, k] = nrm(w)
            if H[k1, k] == 0.0:
                V[k1] = w
                break
            V[k1] = (sc(1) / H[k1, k]) * w
            h = H[:, k]
            for i in range(k):  # givens_rotations :7-23
                h1, h2 = h[i], h[i + 1]
                h[i] = cs[i] * h1 + sn[i] * h2
                h[i + 1] = -sn[i] * h1 + cs[i] * h2
            t = sc(np.hypot(h[k], h[k + 1]))
            cs[k] = h[k] / t
            sn[k] = h[k + 1] / t
            h[k] = cs[k] * h[k] + sn[k] * h[k + 1]
            h[k + 1] = 0.0
            eta[k1] = -sn[k] * eta[k]
            eta[k] = cs[k] * eta[k]
            if abs(eta[k1]) < tol * bnrm:
                break
        yv = eta[:k1].copy()  # dtrsv_/strsv_ upper, no-trans, non-unit (:26-44): back substitution
        for i in range(k1 - 1, -1, -1):
            yv[i] = yv[i] / H[i, i]
            for j in range(i):
                yv[j] = yv[j] - yv[i] * H[j, i]
        for k in range(k1):
            x[:] = (yv[k] * V[k] + sc(1) * x).astype(sc)
        r = np.asarray(A(x), sc)
        out["num_matvec"] += 1
        r = (sc(1) * b + sc(-1) * r).astype(sc)
        r_nrm = nrm(r)
        out["res_norm"].append(float(r_nrm))
        if r_nrm < tol * bnrm:
            out["success"] = True
            break
        it += 1
    out["num_iter"] = it
    return out
