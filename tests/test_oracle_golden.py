"""CPU: the oracle restatement (oracle/setup_np.py, oracle/oracle.c) against the fixtures generated from the
reference's own host objects (scripts/make_golden.py) and against the reference tests' analytic KATs."""
import numpy as np
import pytest

from conftest import load_mesh_file
from oracle import ops as O
from oracle import setup_np as S
from oracle.rdmp import fnv1a64

TABLE_PAIRS = [(2, 3), (3, 4), (3, 5), (3, 6), (4, 5), (4, 6), (4, 8), (5, 6), (5, 7), (5, 9), (6, 7), (6, 8), (6, 11), (7, 8),
               (7, 9), (7, 12), (8, 9), (8, 10), (8, 14), (9, 10), (9, 15)]


@pytest.mark.parametrize("nb,nq", TABLE_PAIRS)
def test_tables_bit_exact(gold, nb, nq):
    b = S.Basis(nb)
    gx, gw = S.gauss_legendre(nq)
    lx, lw = S.gauss_lobatto(nq)
    mine = {"gll_x": b.x, "gll_w": b.w, "Dnodes": b.deriv(b.x).T.ravel(), "gl_x": gx, "gl_w": gw, "P": b.eval(gx).T.ravel(),
            "D": b.deriv(gx).T.ravel(), "gll_nq_x": lx, "gll_nq_w": lw}
    for k, v in mine.items():
        assert np.array_equal(v, gold.tables["tables_%d_%d_%s" % (nb, nq, k)]), k


def _mesh(tag):
    if tag.startswith("rect"):
        nx = int(tag[4:])
        return S.uniform_rect(nx, -1.0, 1.0, nx, -1.0, 1.0)
    xy, el = load_mesh_file()
    return S.mesh_from_vertices(xy, el)


@pytest.mark.parametrize("tag,nb", [("rect2", 3), ("rect3", 2), ("rect10", 4), ("rect10", 5), ("rect6", 8), ("unstr", 2), ("unstr", 4),
                                    ("unstr", 5), ("unstr", 9)])
def test_h1space_bit_exact(gold, tag, nb):
    mesh = _mesh(tag)
    g = lambda k: gold.h1["h1_%s_%d_%s" % (tag, nb, k)]
    I3, ndof, xy = S.h1space(mesh, S.Basis(nb))
    fI, proj = S.facespace(mesh, I3, nb, mesh.boundary_edges)
    assert ndof == g("ndof")[0]
    assert np.array_equal(mesh.edges.ravel(), g("edges"))
    assert np.array_equal(mesh.boundary_edges, g("boundary_edges"))
    assert np.array_equal(I3.ravel(), g("I"))
    assert np.array_equal(xy.ravel(), g("xy"))
    assert np.array_equal(fI.ravel(), g("face_I"))
    assert np.array_equal(proj, g("face_proj"))
    assert mesh.min_h == g("min_h")[0] and mesh.max_h == g("max_h")[0]


def test_h1space_known_table():
    # uniform_rect(2), nb=3 numbering listed explicitly
    mesh = S.uniform_rect(2, -1.0, 1.0, 2, -1.0, 1.0)
    I3, ndof, _ = S.h1space(mesh, S.Basis(3))
    assert ndof == 25
    assert I3[1].tolist() == [[2, 9, 10], [5, 11, 12], [8, 13, 14]]
    assert I3[2].tolist() == [[6, 7, 8], [15, 16, 17], [18, 19, 20]]
    assert I3[3].tolist() == [[8, 13, 14], [17, 21, 22], [20, 23, 24]]


def test_h1space_hash_rect64(gold):
    mesh = S.uniform_rect(64, -1.0, 1.0, 64, -1.0, 1.0)
    I3, ndof, xy = S.h1space(mesh, S.Basis(5))
    h = gold.hashes["h1_rect64_5"]
    assert ndof == h["ndof"] == (64 * 4 + 1) ** 2
    assert fnv1a64(I3) == h["I"] and fnv1a64(xy) == h["xy"] and fnv1a64(mesh.edges) == h["edges"]


@pytest.mark.parametrize("nx,nb", [(8, 4), (16, 4), (8, 8)])
def test_ensemble_bit_exact(gold, nx, nb):
    mesh = S.uniform_rect(nx, -1.0, 1.0, nx, -1.0, 1.0)
    I3, ndof, _ = S.h1space(mesh, S.Basis(nb))
    lab, nd = S.ddh_labels(nx, nx, nb)
    en = S.ensemble(mesh, I3, nb, nd, lab)
    g = lambda k: gold.ens["ens_%d_%d_%s" % (nx, nb, k)]
    assert nd == g("n_domains")[0]
    for mine, k in [(en.s_dof, "sizes"), (en.s_fdof, "fsizes"), (en.gI, "gI"), (en.elems, "elements"), (en.faces, "faces"),
                    (en.sI, "sI"), (en.fI, "fI"), (en.pI, "pI"), (en.cmap, "cmap")]:
        assert np.array_equal(np.asarray(mine).ravel(), g(k)), k


def test_ddh_lambda_map_quirks():
    # cross-point overwrites of B; nb=4: nx=8 -> 4 overwrites, nx=16 -> 36
    for nx, expect in [(8, 4), (16, 36)]:
        mesh = S.uniform_rect(nx, -1.0, 1.0, nx, -1.0, 1.0)
        b = S.Basis(4)
        I3, ndof, _ = S.h1space(mesh, b)
        d = S.ddh_setup(10.0, np.ones(ndof), mesh, b, I3, ndof, nx, nx)
        # each overwritten (subdomain, face DOF) pair loses both of its B entries (read slot and write slot)
        lost_entries = 4 * d.n_shared - int((d.B >= 0).sum())
        assert lost_entries == 2 * expect


# analytic KATs of the reference tests (tests/mass.cpp:13-111 tol 1e-8, tests/stiffness.cpp:28-98 tol 1e-6), and the
# norms SURVEY §8(c) recorded for them
KAT = {4: (2.818274178244191, 0.3506878750099755), 5: (2.154376929860792, 0.2673370236469549),
       6: (1.745708174993621, 0.2163967248842767), 8: (1.267901291598308, 0.1570508892046682)}
KAT_U = {4: (2.636515923128228, 0.3244734168028175), 5: (2.015751653748351, 0.2473989012502660),
         6: (1.633439412048724, 0.2002688305352837), 8: (1.186378555139718, 0.1453502673916969)}


def _lf(fem, nq, f):
    """(f, phi) with the nb+2 rule: reference include/LinearFunctional.hpp:47-112 (lf_action)."""
    x, w = S.gauss_legendre(nq)
    P = fem.basis.eval(x)  # (nq, nb)
    X = fem.coordinates(x).reshape(fem.nel, nq, nq, 2)
    dJ = fem.measures(x).reshape(fem.nel, nq, nq)
    g = dJ * w[None, None, :] * w[None, :, None] * f(X[..., 0], X[..., 1])  # [el, ty, tx]
    loc = np.einsum("ja,ejk,kb->eab", P, g, P)  # [el, ty', tx']
    out = np.zeros(fem.ndof)
    np.add.at(out, fem.I, loc.reshape(-1))
    return out


@pytest.mark.parametrize("kind", ["rect", "unstr"])
@pytest.mark.parametrize("nb", [4, 5, 6, 8])
def test_operator_kats(kind, nb):
    mesh = _mesh("rect10" if kind == "rect" else "unstr")
    fem = O.H1(mesh, nb)
    x, y = fem.xy[:, 0], fem.xy[:, 1]
    fs_ = (x ** 5 - 5 * x) * (y ** 3 - 3 * y)
    fm = 3 * x * x - 2 * x * y + y + 1
    Sf = O.StiffnessMatrix(fem, nb + 2).action(fs_)
    Mf = O.MassMatrix(fem).action(fm)
    ks, km = (KAT if kind == "rect" else KAT_U)[nb]
    assert abs(np.linalg.norm(Sf) - ks) < 1e-12 * ks
    assert abs(np.linalg.norm(Mf) - km) < 1e-12 * km
    b = _lf(fem, nb + 2, lambda X, Y: 3 * X * X - 2 * X * Y + Y + 1)
    assert np.linalg.norm(Mf - b) / np.linalg.norm(b) < 1e-8
    if nb >= 6:
        L = _lf(fem, nb + 2, lambda X, Y: -6.0 * Y * (X ** 5 - 5 * X) - 20.0 * X ** 3 * (Y ** 3 - 3.0 * Y))
        assert np.linalg.norm(Sf - L) / np.linalg.norm(L) < 1e-6


def test_gmres_toeplitz_kat():
    # tests/gmres.cpp:41-76: GMRES(5) on the nonsymmetric tridiagonal Toeplitz [1,-3,1.5], n=1024, tol 1e-10
    n = 1 << 10
    rng = np.random.default_rng(0)
    xt = rng.random(n)

    def A(v):
        y = -3.0 * v
        y[1:] += 1.0 * v[:-1]
        y[:-1] += 1.5 * v[1:]
        return y

    x = np.zeros(n)
    out = O.gmres(n, x, A, A(xt), 5, 100, 1e-10)
    assert out["success"]
    assert np.linalg.norm(x - xt) / np.linalg.norm(xt) < 1e-8


@pytest.mark.parametrize("nx,ny,nb,box", [(2, 2, 3, (-1, 1, -1, 1)), (5, 3, 4, (-1, 2, 0, 1)), (10, 10, 5, (-1, 1, -1, 1)), (7, 9, 2, (0, 1, 0, 3)),
                                          (6, 6, 8, (-1, 1, -1, 1))])
def test_closed_form_uniform_rect_equals_first_touch(nx, ny, nb, box):
    """the vectorised closed form used by bench.py's reference arm at 1024^2 (oracle/setup_np.py:uniform_rect_closed_form)
    against the loop restatement of the reference's constructors, entry by entry and bit for bit"""
    ax, bx, ay, by = box
    m = S.uniform_rect(nx, ax, bx, ny, ay, by)
    B = S.Basis(nb)
    I3, ndof, xy = S.h1space(m, B)
    fI, proj = S.facespace(m, I3, nb, m.boundary_edges)
    c = S.uniform_rect_closed_form(nx, ax, bx, ny, ay, by, B)
    assert np.array_equal(I3, c["I"]) and ndof == c["ndof"] and np.array_equal(xy, c["xy"])
    assert np.array_equal(m.corners, c["corners"])
    assert np.array_equal(m.edges[m.boundary_edges, 2], c["face_el"]) and np.array_equal(m.edges[m.boundary_edges, 4], c["face_side"])
    assert np.array_equal(fI, c["face_I"]) and np.array_equal(proj, c["face_proj"])
    assert np.array_equal(m.edge_meas[m.boundary_edges], c["face_meas"])


def test_closed_form_matches_reference_at_1024(gold):
    """BASELINE config 2 mesh: global map, nodal coordinates and boundary face space of the closed form hash to what the
    UNMODIFIED reference's H1Space / FaceSpace produce at uniform_rect(1024), n_basis 5 (scripts/make_golden.py)"""
    h = gold.hashes["h1_rect1024_5"]
    c = S.uniform_rect_closed_form(1024, -1.0, 1.0, 1024, -1.0, 1.0, S.Basis(5))
    assert c["ndof"] == h["ndof"]
    assert fnv1a64(c["I"]) == h["I"] and fnv1a64(c["xy"]) == h["xy"]
    assert fnv1a64(c["face_I"]) == h["face_I"] and fnv1a64(c["face_proj"]) == h["face_proj"]
