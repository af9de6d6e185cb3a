"""TEST INFRASTRUCTURE — NOT PRODUCT CODE.

numpy / pure-Python restatement of the reference's HOST-side setup for the solve path: 1-D tables
(QuadratureRule, Basis), mesh topology (Mesh2D::from_vertices / uniform_rect), the H1Space
global-to-local map, FaceSpace maps, EnsembleSpace maps and the index/coefficient part of DDH::DDH.
Every function cites the reference file:line it follows (relative to /root/reference).
Pure-Python loops: meant for the small parity cases (<= ~64x64 elements).

Parity status: PINNED — scripts/make_golden.py compares every array produced here with the
reference's own host objects (oracle/_ref/ref_driver tables|h1|ensemble) bit for bit, and
tests/test_oracle_golden.py re-checks against the committed fixtures in tests/golden/.
"""
import math
import numpy as np

GL, GLL = 0, 1

# ------------------------------------------------------------------------------------------------
# source/QuadratureRule.cpp
# ------------------------------------------------------------------------------------------------
_GL_NODES = {  # :72-83
    1: [0.0],
    2: [-0.577350269189625764509149, 0.577350269189625764509149],
    3: [-0.774596669241483377035853, 0.0, 0.774596669241483377035853],
    4: [-0.861136311594052575223946, -0.339981043584856264802666, 0.339981043584856264802666, 0.861136311594052575223946],
    5: [-0.906179845938663992797627, -0.538469310105683091036314, 0.0, 0.538469310105683091036314, 0.906179845938663992797627],
    6: [-0.932469514203152027812302, -0.661209386466264513661400, -0.238619186083196908630502, 0.238619186083196908630502, 0.661209386466264513661400, 0.932469514203152027812302],
    7: [-0.949107912342758524526190, -0.741531185599394439863865, -0.405845151377397166906606, 0.0, 0.405845151377397166906606, 0.741531185599394439863865, 0.949107912342758524526190],
    8: [-0.960289856497536231683561, -0.796666477413626739591554, -0.525532409916328985817739, -0.183434642495649804939476, 0.183434642495649804939476, 0.525532409916328985817739, 0.796666477413626739591554, 0.960289856497536231683561],
    9: [-0.968160239507626089835576, -0.836031107326635794299430, -0.613371432700590397308702, -0.324253423403808929038538, 0.0, 0.324253423403808929038538, 0.613371432700590397308702, 0.836031107326635794299430, 0.968160239507626089835576],
    10: [-0.973906528517171720077964, -0.865063366688984510732097, -0.679409568299024406234327, -0.433395394129247190799266, -0.148874338981631210884826, 0.148874338981631210884826, 0.433395394129247190799266, 0.679409568299024406234327, 0.865063366688984510732097, 0.973906528517171720077964],
}
_GLL_NODES = {  # :142-151 (15-digit literals, note the 16-digit one in n=8)
    2: [-1, 1],
    3: [-1, 0, 1],
    4: [-1, -0.447213595499958, 0.447213595499958, 1],
    5: [-1, -0.654653670707977, 0, 0.654653670707977, 1],
    6: [-1, -0.765055323929465, -0.285231516480645, 0.285231516480645, 0.765055323929465, 1],
    7: [-1, -0.830223896278567, -0.468848793470714, 0.0, 0.468848793470714, 0.830223896278567, 1],
    8: [-1, -0.871740148509607, -0.591700181433142, -0.209299217902479, 0.2092992179024789, 0.591700181433142, 0.871740148509607, 1],
    9: [-1, -0.899757995411460, -0.677186279510738, -0.363117463826178, 0, 0.363117463826178, 0.677186279510738, 0.899757995411460, 1],
}


def _jacobiP_next(m, a, b, x, y1, y2):  # :21-27
    yp1 = (2 * m + a + b - 1) * ((2 * m + a + b) * (2 * m + a + b - 2) * x + a * a - b * b) * y1 \
        - 2 * (m + a - 1) * (m + b - 1) * (2 * m + a + b) * y2
    yp1 /= 2 * m * (m + a + b) * (2 * m + a + b - 2)
    return yp1


def _jacobiP(n, a, b, x):  # :29-46
    ym1 = 1.0
    if n == 0:
        return ym1
    y = (a + 1) + 0.5 * (a + b + 2) * (x - 1)
    for m in range(2, n + 1):
        yp1 = _jacobiP_next(m, a, b, x, y, ym1)
        ym1 = y
        y = yp1
    return y


def _jacobiP_derivative(k, n, a, b, x):  # :48-57
    if k > n:
        return 0.0
    s = _lgamma(n + a + b + 1 + k) - _lgamma(n + a + b + 1) - k * math.log(2)
    return math.exp(s) * _jacobiP(n - k, a + k, b + k, x)


def _dsteqr_eigs(d, e):
    """Eigenvalues of the symmetric tridiagonal companion matrix: LAPACK dsteqr_(COMPZ='N'), the
    un-vendored third-party routine the reference links (CMakeLists.txt:47-53). Called through
    ctypes from the OpenBLAS 0.3.15 bundled with this image's opencv wheel (the same library
    oracle/_ref links), falling back to scipy's dsterf (same eigenvalues up to rounding; the three
    Newton steps that follow absorb the difference)."""
    import ctypes, glob, site
    d = np.array(d, float)
    e = np.array(e, float)
    libs = []
    for sp in site.getsitepackages():
        libs += glob.glob(sp + "/opencv_python_headless.libs/libopenblasp-*.so")
    if libs:
        ldir = libs[0].rsplit("/", 1)[0]
        for dep in ("libquadmath-*.so*", "libgfortran-*.so*"):
            for g in glob.glob(ldir + "/" + dep):
                ctypes.CDLL(g, mode=ctypes.RTLD_GLOBAL)
        lib = ctypes.CDLL(libs[0])
        n = ctypes.c_int(len(d))
        ldz = ctypes.c_int(1)
        info = ctypes.c_int(0)
        lib.dsteqr_(ctypes.c_char_p(b"N"), ctypes.byref(n), d.ctypes.data_as(ctypes.c_void_p),
                    e.ctypes.data_as(ctypes.c_void_p), None, ctypes.byref(ldz), None, ctypes.byref(info))
        assert info.value == 0
        return list(d)
    from scipy.linalg import lapack
    d2, info = lapack.dsterf(d, e)
    assert info == 0
    return list(d2)


def _c_lgamma():
    # CPython's math.lgamma is its own implementation; the reference calls glibc's (std::lgamma).
    import ctypes, ctypes.util
    libm = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
    libm.lgamma.restype = ctypes.c_double
    libm.lgamma.argtypes = [ctypes.c_double]
    return libm.lgamma


_lgamma = _c_lgamma()


def gauss_legendre(n):  # :64-132
    a = b = 0.0
    if n in _GL_NODES:
        x = [float(v) for v in _GL_NODES[n]]
    else:
        E = []
        for i in range(n - 1):
            k = float(i + 1)
            E.append(k * math.sqrt(1.0 / (4.0 * k * k - 1.0)))
        x = _dsteqr_eigs([0.0] * n, E)
        for i in range(n // 2):
            for _ in range(3):
                P = _jacobiP(n, a, b, x[i])
                dP = _jacobiP_derivative(1, n, a, b, x[i])
                x[i] -= P / dP
            x[n - 1 - i] = -x[i]
        if n & 1:
            x[n // 2] = 0.0
    w = []
    for i in range(n):
        dP = _jacobiP_derivative(1, n, a, b, x[i])
        w.append(2.0 / (1.0 - x[i] * x[i]) / (dP * dP))
    return np.array(x), np.array(w)


def gauss_lobatto(n):  # :134-202
    if n in _GLL_NODES:
        x = [float(v) for v in _GLL_NODES[n]]
    else:
        E = []
        for i in range(n - 3):
            ii = float(i + 1)
            E.append(math.sqrt(ii * (ii + 2.0) / ((2.0 * ii + 3.0) * (2.0 * ii + 1.0))))
        x = [-1.0] + _dsteqr_eigs([0.0] * (n - 2), E) + [1.0]
        for i in range(1, n // 2):
            for _ in range(3):
                P = _jacobiP(n - 2, 1.0, 1.0, x[i])
                dP = _jacobiP_derivative(1, n - 2, 1.0, 1.0, x[i])
                x[i] -= P / dP
            x[n - 1 - i] = -x[i]
        if n & 1:
            x[n // 2] = 0.0
    w = []
    for i in range(n):
        p = _jacobiP(n - 1, 0.0, 0.0, x[i])
        w.append(2.0 / (n * (n - 1) * (p * p)))
    return np.array(x), np.array(w)


def quadrature(n, kind):
    return gauss_legendre(n) if kind == GL else gauss_lobatto(n)


# ------------------------------------------------------------------------------------------------
# source/Basis.cpp
# ------------------------------------------------------------------------------------------------
class Basis:
    def __init__(self, n):  # :109-140
        self.n = n
        self.x, self.w = gauss_lobatto(n)
        x = self.x
        wb = []
        for i in range(n):  # barycentric_weights :3-24
            t = 1.0
            for j in range(n):
                if i == j:
                    continue
                t *= x[i] - x[j]
            wb.append(1.0 / t)
        diff = max(wb) - min(wb)
        self.wb = [v / diff for v in wb]

    def _interp(self, x0, y):  # :32-52
        eps = np.finfo(float).eps
        A = 0.0
        B = 0.0
        for i in range(self.n):
            xdiff = x0 - self.x[i]
            if x0 == self.x[i] or abs(xdiff) <= eps:
                return y[i]
            C = self.wb[i] / xdiff
            A += C * y[i]
            B += C
        return A / B

    def _deriv(self, x0, y):  # :60-105
        eps = np.finfo(float).eps
        n = self.n
        A = 0.0
        B = 0.0
        p = self._interp(x0, y)
        atnode = False
        inode = -1
        for j in range(n):
            if x0 == self.x[j] or abs(x0 - self.x[j]) <= eps:
                atnode = True
                B = -self.wb[j]
                inode = j
        if atnode:
            for j in range(n):
                if j == inode:
                    continue
                A += self.wb[j] * (p - y[j]) / (x0 - self.x[j])
        else:
            for j in range(n):
                t = self.wb[j] / (x0 - self.x[j])
                A += t * (p - y[j]) / (x0 - self.x[j])
                B += t
        return A / B

    def eval(self, xq):  # :142-155 -> P (m, n) with P[j, i] = phi_i(x_j); stored column-major by callers
        m = len(xq)
        P = np.zeros((m, self.n))
        for i in range(self.n):
            y = [0.0] * self.n
            y[i] = 1.0
            for j in range(m):
                P[j, i] = self._interp(xq[j], y)
        return P

    def deriv(self, xq):  # :157-170
        m = len(xq)
        D = np.zeros((m, self.n))
        for i in range(self.n):
            y = [0.0] * self.n
            y[i] = 1.0
            for j in range(m):
                D[j, i] = self._deriv(xq[j], y)
        return D


# ------------------------------------------------------------------------------------------------
# source/Mesh2D.cpp
# ------------------------------------------------------------------------------------------------
class Mesh:
    pass


def mesh_from_vertices(xy, elems):
    """source/Mesh2D.cpp:11-136. xy (nv,2), elems (nel,4). Edge table columns:
    nodes0,nodes1,el0,el1,side0,side1,delta,is_boundary (el1/side1 = -1 on the boundary).
    The reference's 32-bit edge key (min + nv*max, :64-67) is replaced by a tuple key (R7)."""
    xy = np.asarray(xy, float)
    elems = np.asarray(elems, np.int64)
    nv, nel = len(xy), len(elems)
    emap1 = (0, 1, 3, 0)
    emap2 = (1, 2, 2, 3)
    node_elems = [[] for _ in range(nv)]  # connected_elements: (corner i, element id) in element order (:39-53)
    node_interior = np.zeros(nv, bool)
    for el in range(nel):
        for i in range(4):
            node_elems[elems[el, i]].append((i, el))
    edge_map = {}
    edges = []
    for el in range(nel):
        for s in range(4):
            C0 = int(elems[el, emap1[s]])
            C1 = int(elems[el, emap2[s]])
            k = (min(C0, C1), max(C0, C1))
            if k not in edge_map:
                edge_map[k] = len(edges)
                edges.append([C0, C1, el, -1, s, -1, 1, 1])
            else:
                e = edges[edge_map[k]]
                e0, s0 = e[2], e[4]
                n1 = int(elems[e0, emap1[s0]])
                e[3] = el
                e[5] = s
                e[7] = 0
                e[6] = 1 if C0 == n1 else -1
                node_interior[C0] = True
                node_interior[C1] = True
    m = Mesh()
    m.xy = xy
    m.elems = elems
    m.n_elem = nel
    m.n_nodes = nv
    m.edges = np.array(edges, np.int32).reshape(-1, 8)
    m.boundary_edges = np.array([i for i, e in enumerate(edges) if e[7] == 1], np.int32)
    m.interior_edges = np.array([i for i, e in enumerate(edges) if e[7] == 0], np.int32)
    m.node_elems = node_elems
    d = xy[m.edges[:, 1]] - xy[m.edges[:, 0]]
    # StraightEdge: meas = hypot(dx,dy)/2, length = 2*meas (include/Edge.hpp:95-157)
    m.edge_meas = np.array([math.hypot(a, b) / 2 for a, b in d])
    m.min_h = float(np.min(2.0 * m.edge_meas))
    m.max_h = float(np.max(2.0 * m.edge_meas))
    m.corners = xy[elems]  # (nel, 4, 2)
    return m


def uniform_rect(nx, ax, bx, ny, ay, by):
    """source/Mesh2D.cpp:138-171."""
    dx = (bx - ax) / nx
    dy = (by - ay) / ny
    xy = np.zeros(((nx + 1) * (ny + 1), 2))
    for j in range(ny + 1):
        y = ay + dy * j
        for i in range(nx + 1):
            xy[i + (nx + 1) * j, 0] = ax + dx * i
            xy[i + (nx + 1) * j, 1] = y
    elems = np.zeros((nx * ny, 4), np.int64)
    for j in range(ny):
        for i in range(nx):
            l = lambda a, b: a + (nx + 1) * b
            elems[i + nx * j] = (l(i, j), l(i + 1, j), l(i + 1, j + 1), l(i, j + 1))
    return mesh_from_vertices(xy, elems)


def _E2V(nc, i, f, el):  # source/H1Space.cpp:28-34
    m = i if f in (0, 2) else (nc - 1 if f == 1 else 0)
    n = i if f in (1, 3) else (nc - 1 if f == 2 else 0)
    return m + nc * (n + nc * el)


def _N2V(nc, c, el):  # :37-43
    m = 0 if c in (0, 3) else nc - 1
    n = 0 if c in (0, 1) else nc - 1
    return m + nc * (n + nc * el)


def h1space(mesh, basis):
    """source/H1Space.cpp:11-127. Returns I (nel, nb, nb) with I[el, j, i] = I(i,j,el), ndof, xy (ndof,2)."""
    nb = basis.n
    nel = mesh.n_elem
    mask = {}
    if nb > 2:
        for e in mesh.interior_edges:
            n0, n1, el0, el1, s0, s1, delta, _ = (int(v) for v in mesh.edges[e])
            for i in range(1, nb - 1):
                j = nb - 1 - i if delta < 0 else i
                mask[_E2V(nb, j, s1, el1)] = _E2V(nb, i, s0, el0)
    for k in range(mesh.n_nodes):
        ce = mesh.node_elems[k]
        if not ce:
            continue
        c0, el0 = ce[0]
        v0 = _N2V(nb, c0, el0)
        for c, el in ce[1:]:
            mask[_N2V(nb, c, el)] = v0
    N = nel * nb * nb
    I = np.zeros(N, np.int32)
    l = 0
    for i in range(N):
        if i not in mask:
            I[i] = l
            l += 1
    for v1, v0 in mask.items():
        I[v1] = I[v0]
    ndof = N - len(mask)
    I3 = I.reshape(nel, nb, nb)
    xy = np.zeros((ndof, 2))
    for el in range(nel):
        x = mesh.corners[el]
        for j in range(nb):
            for i in range(nb):
                xi0, xi1 = basis.x[i], basis.x[j]
                b = (0.25 * (1.0 - xi0) * (1.0 - xi1), 0.25 * (1.0 + xi0) * (1.0 - xi1),
                     0.25 * (1.0 + xi0) * (1.0 + xi1), 0.25 * (1.0 - xi0) * (1.0 + xi1))
                x0 = 0.0
                x1 = 0.0
                for c in range(4):
                    x0 += x[c, 0] * b[c]
                    x1 += x[c, 1] * b[c]
                xy[I3[el, j, i]] = (x0, x1)
    return I3, ndof, xy


def facespace(mesh, I3, nb, faces):
    """source/H1Space.cpp:129-187. Returns fI (nf, nb) subspace indices, proj (fdof,) global indices."""
    K = I3.reshape(-1)
    mask = {}
    P = []
    fI = np.zeros((len(faces), nb), np.int32)
    for f, e in enumerate(faces):
        el, s = int(mesh.edges[e, 2]), int(mesh.edges[e, 4])
        for i in range(nb):
            idx = int(K[_E2V(nb, i, s, el)])
            if idx not in mask:
                mask[idx] = len(P)
                P.append(idx)
            fI[f, i] = mask[idx]
    return fI, np.array(P, np.int32)


def uniform_rect_closed_form(nx, ax, bx, ny, ay, by, basis):
    """Vectorised numpy restatement of uniform_rect -> from_vertices -> H1Space -> boundary FaceSpace for structured meshes
    (source/Mesh2D.cpp:138-171, :11-136; source/H1Space.cpp:11-127, :129-187), for sizes where the loops above take minutes
    (bench.py's reference arm at 1024^2). No hash maps: on uniform_rect the first-touch rule has a closed form -
      * element (ex, ey) introduces the nodes with (i > 0 or ex == 0) and (j > 0 or ey == 0), numbered in (j, i) scan order after
        all nodes of the elements before it; every other node copies the element to its left / below;
      * nodal coordinates are written by every element that touches a node, so the LAST element (largest id) wins;
      * boundary edges in edge-id order = boundary (element, side) pairs in (element, side) order, sides 0..3 = bottom, right,
        top, left; face DOFs are numbered first-touch over (face, i).
    Checked entry by entry against the loop versions in tests/test_oracle_golden.py. Returns a dict of arrays."""
    nb = basis.n
    p = nb - 1
    dx = (bx - ax) / nx
    dy = (by - ay) / ny
    vx = ax + dx * np.arange(nx + 1)      # :147-157 (one rounding per product and per sum, as numpy does)
    vy = ay + dy * np.arange(ny + 1)
    ex = np.arange(nx)
    ey = np.arange(ny)
    ni = np.where(ex == 0, nb, p)          # new nodes per element row / column
    nj = np.where(ey == 0, nb, p)
    cnt = (nj[:, None] * ni[None, :]).ravel()             # element id = ex + nx * ey
    base = np.concatenate([[0], np.cumsum(cnt)[:-1]]).reshape(ny, nx)
    ndof = int(cnt.sum())
    # owner (element + local node) of grid node g along one axis: g == 0 -> (0, 0), else ((g-1) // p, g - p * that)
    def owner(n_el):
        g = np.arange(n_el * p + 1)
        e = np.where(g == 0, 0, (g - 1) // p)
        return e, g - p * e
    ox, oi = owner(nx)
    oy, oj = owner(ny)
    gid = (base[oy[:, None], ox[None, :]] + (oj - (oy > 0))[:, None] * ni[ox][None, :] + (oi - (ox > 0))[None, :])  # (gy, gx)
    gx = (ex[:, None] * p + np.arange(nb)[None, :])       # (nx, nb)
    gy = (ey[:, None] * p + np.arange(nb)[None, :])       # (ny, nb)
    # I[el, j, i] with el = ex + nx * ey
    I = gid[gy[:, None, :, None], gx[None, :, None, :]].reshape(nx * ny, nb, nb).astype(np.int32)
    # coordinates: last element touching grid node g: e = min(g // p, n_el - 1), local index g - p e
    def last(n_el):
        g = np.arange(n_el * p + 1)
        e = np.minimum(g // p, n_el - 1)
        return e, g - p * e
    lx, li = last(nx)
    ly, lj = last(ny)
    xi0 = basis.x[li][None, :]
    xi1 = basis.x[lj][:, None]
    b = (0.25 * (1.0 - xi0) * (1.0 - xi1), 0.25 * (1.0 + xi0) * (1.0 - xi1), 0.25 * (1.0 + xi0) * (1.0 + xi1),
         0.25 * (1.0 - xi0) * (1.0 + xi1))
    X0, X1 = vx[lx][None, :], vx[lx + 1][None, :]
    Y0, Y1 = vy[ly][:, None], vy[ly + 1][:, None]
    zero = np.zeros((ny * p + 1, nx * p + 1))
    cx = (X0 + zero, X1 + zero, X1 + zero, X0 + zero)      # corner order 0..3 = (i,j), (i+1,j), (i+1,j+1), (i,j+1)
    cy = (Y0 + zero, Y0 + zero, Y1 + zero, Y1 + zero)
    px = 0.0
    py = 0.0
    for c in range(4):                                      # source/Element.cpp:5-19: x += corner * b, c = 0..3
        px = px + cx[c] * b[c]
        py = py + cy[c] * b[c]
    xy = np.zeros((ndof, 2))
    xy[gid.ravel(), 0] = px.ravel()
    xy[gid.ravel(), 1] = py.ravel()
    corners = np.zeros((ny, nx, 4, 2))
    corners[:, :, 0, 0] = corners[:, :, 3, 0] = vx[:-1][None, :]
    corners[:, :, 1, 0] = corners[:, :, 2, 0] = vx[1:][None, :]
    corners[:, :, 0, 1] = corners[:, :, 1, 1] = vy[:-1][:, None]
    corners[:, :, 2, 1] = corners[:, :, 3, 1] = vy[1:][:, None]
    corners = corners.reshape(nx * ny, 4, 2)
    # boundary faces: (element, side) pairs on the domain boundary, ordered by (element, side)
    EX, EY = np.meshgrid(ex, ey, indexing="xy")
    el_id = (EX + nx * EY).ravel()
    pairs = []
    for side, on in ((0, EY == 0), (1, EX == nx - 1), (2, EY == ny - 1), (3, EX == 0)):
        e = el_id[on.ravel()]
        pairs.append(np.stack([e, np.full_like(e, side)], 1))
    pairs = np.concatenate(pairs)
    pairs = pairs[np.lexsort((pairs[:, 1], pairs[:, 0]))]
    f_el, f_s = pairs[:, 0], pairs[:, 1]
    k = np.arange(nb)[None, :]
    m = np.where(np.isin(f_s, (0, 2))[:, None], k, np.where(f_s[:, None] == 1, nb - 1, 0))
    n = np.where(np.isin(f_s, (1, 3))[:, None], k, np.where(f_s[:, None] == 2, nb - 1, 0))
    idx = I[f_el[:, None], n, m]                            # (nf, nb) global DOFs in (face, i) order
    flat = idx.ravel()
    uniq, first = np.unique(flat, return_index=True)
    order = np.argsort(first)                               # first-touch order
    proj = uniq[order].astype(np.int32)
    rank = np.empty(len(uniq), np.int64)
    rank[order] = np.arange(len(uniq))
    fI = rank[np.searchsorted(uniq, flat)].reshape(idx.shape).astype(np.int32)
    # StraightEdge::meas = length / 2 between the side's two vertices (include/Edge.hpp:95-157)
    side_nodes = np.array([[0, 1], [1, 2], [3, 2], [0, 3]])
    d = corners[f_el, side_nodes[f_s, 1]] - corners[f_el, side_nodes[f_s, 0]]
    meas = np.hypot(d[:, 0], d[:, 1]) / 2
    return dict(I=I, ndof=ndof, xy=xy, corners=corners, face_el=f_el, face_side=f_s, face_I=fI, face_proj=proj, face_meas=meas)


# ------------------------------------------------------------------------------------------------
# source/EnsembleSpace.cpp:11-287
# ------------------------------------------------------------------------------------------------
class Ensemble:
    pass


def ensemble(mesh, I3, nb, n_spaces, labels):
    nel = mesh.n_elem
    E = [[] for _ in range(n_spaces)]
    el2s = np.zeros(nel, np.int64)
    for el in range(nel):
        p = int(labels[el])
        E[p].append(el)
        el2s[el] = len(E[p]) - 1
    s_elems = np.array([len(v) for v in E], np.int32)
    mx_elems = int(s_elems.max())
    elems = -np.ones((n_spaces, mx_elems), np.int32)
    for p in range(n_spaces):
        elems[p, :len(E[p])] = E[p]

    F = [[] for _ in range(n_spaces)]
    shared_faces = []
    for e in range(len(mesh.edges)):
        n0, n1, el0, el1, s0, s1, delta, isb = (int(v) for v in mesh.edges[e])
        S0 = int(labels[el0])
        if isb:
            F[S0].append((e, 0))
        else:
            S1 = int(labels[el1])
            if S0 != S1:
                F[S0].append((e, 0))
                F[S1].append((e, 1))
                shared_faces.append((S0, S1, len(F[S0]) - 1, len(F[S1]) - 1))
    s_faces = np.array([len(v) for v in F], np.int32)
    mx_faces = int(s_faces.max())
    faces = -np.ones((n_spaces, mx_faces), np.int32)
    face_side = -np.ones((n_spaces, mx_faces), np.int32)
    for p in range(n_spaces):
        for i, (f, side) in enumerate(F[p]):
            faces[p, i] = f
            face_side[p, i] = side

    sI = -np.ones((n_spaces, mx_elems, nb, nb), np.int32)  # sI[p, el, j, i] = sI(i,j,el,p)
    s2g = []
    for p in range(n_spaces):
        uniq = {}
        lst = []
        for el in range(s_elems[p]):
            g_el = elems[p, el]
            for j in range(nb):
                for i in range(nb):
                    g = int(I3[g_el, j, i])
                    if g not in uniq:
                        uniq[g] = len(lst)
                        lst.append(g)
                    sI[p, el, j, i] = uniq[g]
        s2g.append(lst)
    s_dof = np.array([len(v) for v in s2g], np.int32)
    mx_ndof = int(s_dof.max())
    gI = -np.ones((n_spaces, mx_ndof), np.int32)
    for p in range(n_spaces):
        gI[p, :s_dof[p]] = s2g[p]

    fI = -np.ones((n_spaces, mx_faces, nb), np.int32)  # fI[p, f, i]
    f2s = []
    for p in range(n_spaces):
        uniq = {}
        lst = []
        for f in range(s_faces[p]):
            e = faces[p, f]
            side = face_side[p, f]
            g_el = int(mesh.edges[e, 2 + side])
            s = int(mesh.edges[e, 4 + side])
            reversed_ = (side == 1 and mesh.edges[e, 6] < 0)
            for i in range(nb):
                j = nb - 1 - i if reversed_ else i
                m = j if s in (0, 2) else (nb - 1 if s == 1 else 0)
                n = j if s in (1, 3) else (nb - 1 if s == 2 else 0)
                el = el2s[g_el]
                idx = int(sI[p, el, n, m])
                if idx not in uniq:
                    uniq[idx] = len(lst)
                    lst.append(idx)
                fI[p, f, i] = uniq[idx]
        f2s.append(lst)
    s_fdof = np.array([len(v) for v in f2s], np.int32)
    mx_fdof = int(s_fdof.max())
    pI = -np.ones((n_spaces, mx_fdof), np.int32)
    for p in range(n_spaces):
        pI[p, :s_fdof[p]] = f2s[p]

    shared_dofs = []
    unique_shared = {}
    for (S0, S1, f0, f1) in shared_faces:
        key = (S0, S1) if S0 < S1 else (S1, S0)
        unq = unique_shared.setdefault(key, set())
        for i in range(nb):
            j0 = int(fI[S0, f0, i])
            j1 = int(fI[S1, f1, i])
            lkey = j0 if S0 < S1 else j1
            if lkey not in unq:
                shared_dofs.append((S0, S1, j0, j1))
                unq.add(lkey)
    en = Ensemble()
    en.n_spaces = n_spaces
    en.s_elems, en.elems, en.mx_elems = s_elems, elems, mx_elems
    en.s_faces, en.faces, en.mx_faces = s_faces, faces, mx_faces
    en.sI, en.gI, en.s_dof, en.mx_ndof = sI, gI, s_dof, mx_ndof
    en.fI, en.pI, en.s_fdof, en.mx_fdof = fI, pI, s_fdof, mx_fdof
    en.cmap = np.array(shared_dofs, np.int32).reshape(-1, 4)
    return en


def ddh_labels(nx, ny, nb, block=16):
    """source/DDH.cpp:336-356."""
    epd = block // nb
    ndx, ndy = nx // epd, ny // epd
    lab = np.zeros(nx * ny, np.int32)
    for j in range(ny):
        for i in range(nx):
            lab[i + nx * j] = (i // epd) + ndx * (j // epd)
    return lab, ndx * ndy


class DDHSetup:
    pass


def ddh_setup(omega, h_a, mesh, basis, I3, g_ndof, nx, ny, block=16):
    """source/DDH.cpp:323-609 (everything except the device call init_geom_factors, which is
    oracle.c:orc_ddh_geom). Arrays are returned in the reference's column-major flat layouts."""
    nb = basis.n
    lab, n_domains = ddh_labels(nx, ny, nb, block)
    en = ensemble(mesh, I3, nb, n_domains, lab)
    d = DDHSetup()
    d.en = en
    d.n_domains = n_domains
    T = (2 * math.pi) / omega
    h = mesh.min_h
    dt = 0.2 * 0.5 * h / (nb * nb)
    nt = int(math.ceil(T / dt))
    dt = T / nt
    d.nt, d.dt = nt, dt
    whf = np.array([dt * (omega / math.pi) * (math.cos(omega * k * dt) - 0.25) for k in range(nt + 1)])
    whf[0] *= 0.5
    whf[nt] *= 0.5
    d.wh_filter = whf.astype(np.float32)
    d.cs = np.array([-math.cos(omega * (0.5 * k * dt)) for k in range(2 * nt + 1)]).astype(np.float32)
    d.sn = np.array([math.sin(omega * (0.5 * k * dt)) for k in range(2 * nt + 1)]).astype(np.float32)

    mx_dof = int(en.s_dof.max())
    mx_fdof = int(en.s_fdof.max())
    mx_el = int(en.s_elems.max())
    d.mx_dof, d.mx_fdof, d.mx_elem_per_dom = mx_dof, mx_fdof, mx_el
    n_shared = len(en.cmap)
    d.n_shared = n_shared
    d.n_lambda = 2 * n_shared

    B = -np.ones((n_domains, 2, mx_fdof), np.int32)  # B[dom, c, j] = B(j, c, dom)
    for k in range(n_shared):
        S0, S1, j0, j1 = (int(v) for v in en.cmap[k])
        B[S0, 0, j0] = k
        B[S0, 1, j0] = n_shared + k
        B[S1, 0, j1] = n_shared + k
        B[S1, 1, j1] = k
    d.B = B

    perm = -np.ones((n_domains, mx_dof), np.int64)
    inv_perm = -np.ones((n_domains, mx_dof), np.int64)
    for s in range(n_domains):
        ndof, fdof = int(en.s_dof[s]), int(en.s_fdof[s])
        pp = set()
        l = 0
        while l < fdof:
            j = int(en.pI[s, l])
            pp.add(j)
            perm[s, l] = j
            l += 1
        for i in range(ndof):
            if i in pp:
                continue
            perm[s, l] = i
            l += 1
        for i in range(ndof):
            inv_perm[s, perm[s, i]] = i

    # the reference leaves never-written entries of its freshly zero-initialised arrays at 0
    gI = np.zeros((n_domains, mx_dof), np.int32)
    sI = np.zeros((n_domains, mx_el, nb, nb), np.int32)
    for s in range(n_domains):
        ndof = int(en.s_dof[s])
        for i in range(ndof):
            gI[s, i] = en.gI[s, perm[s, i]]
        for el in range(int(en.s_elems[s])):
            for l in range(nb):
                for k in range(nb):
                    sI[s, el, l, k] = inv_perm[s, en.sI[s, el, l, k]]
    d.gI, d.sI = gI, sI

    q_x, q_w = basis.x, basis.w
    d.D = basis.deriv(q_x).T.copy().astype(np.float32)  # flat = column-major (nb,nb): D(k,l) at [k + nb*l]
    # detJ at GLL nodes (nb,nb,g_elem)
    g_elem = mesh.n_elem
    detJ = np.zeros((g_elem, nb, nb))
    for el in range(g_elem):
        x = mesh.corners[el]
        for j in range(nb):
            for i in range(nb):
                xi0, xi1 = q_x[i], q_x[j]
                J0 = 0.25 * ((1.0 - xi1) * (x[1, 0] - x[0, 0]) + (1.0 + xi1) * (x[2, 0] - x[3, 0]))
                J1 = 0.25 * ((1.0 - xi1) * (x[1, 1] - x[0, 1]) + (1.0 + xi1) * (x[2, 1] - x[3, 1]))
                J2 = 0.25 * ((1.0 - xi0) * (x[3, 0] - x[0, 0]) + (1.0 + xi0) * (x[2, 0] - x[1, 0]))
                J3 = 0.25 * ((1.0 - xi0) * (x[3, 1] - x[0, 1]) + (1.0 + xi0) * (x[2, 1] - x[1, 1]))
                detJ[el, j, i] = J0 * J3 - J1 * J2
    mi = np.zeros(g_ndof)
    for el in range(g_elem):
        for j in range(nb):
            for i in range(nb):
                mi[I3[el, j, i]] += q_w[i] * q_w[j] * detJ[el, j, i]
    mi = 1.0 / mi

    m = np.zeros((n_domains, mx_dof), np.float32)
    H = np.zeros((n_domains, mx_fdof), np.float32)
    A = np.zeros((n_domains, mx_dof), np.float32)
    gmi = np.zeros((n_domains, mx_dof), np.float32)
    for s in range(n_domains):
        for el in range(int(en.s_elems[s])):
            g_el = en.elems[s, el]
            for j in range(nb):
                for i in range(nb):
                    l = sI[s, el, j, i]
                    # float accumulator += double product: done in double then rounded (C usual conversions)
                    m[s, l] = np.float32(float(m[s, l]) + q_w[i] * q_w[j] * detJ[g_el, j, i])
        for i in range(int(en.s_dof[s])):
            A[s, i] = h_a[gI[s, i]]
            gmi[s, i] = mi[gI[s, i]]
        for f in range(int(en.s_faces[s])):
            ds = mesh.edge_meas[en.faces[s, f]]
            for i in range(nb):
                l = en.fI[s, f, i]
                H[s, l] = np.float32(float(H[s, l]) + ds * q_w[i])
    d.m, d.H, d.a, d.gmi = m, H, A, gmi
    return d
