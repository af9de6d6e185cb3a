"""CPU: the product's host-side setup (libcuddh_b200.so through the C ABI, no kernel launches) against the
golden fixtures from the reference and against the oracle; and the ABI itself (every symbol the header
declares is exported)."""
import ctypes
import os
import re

import numpy as np
import pytest

import cuddhelmholtz_b200 as cb
from conftest import load_mesh_file
from cuddhelmholtz_b200 import capi
from oracle import setup_np as S
from oracle.rdmp import fnv1a64

from test_oracle_golden import TABLE_PAIRS


def test_abi_exports_every_declared_symbol():
    lib = capi.load()
    text = open(capi.HEADER_PATH).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = sorted(set(re.findall(r"\b(cuddh_b200_\w+)\s*\(", text)))
    assert len(names) > 60
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    # and the ctypes prototypes cover them all
    assert sorted(capi.SIGNATURES) == names
    assert lib.cuddh_b200_version() == 100


def test_error_path_returns_status_not_abort():
    lib = capi.load()
    h = ctypes.c_void_p()
    assert lib.cuddh_b200_basis_create(1, ctypes.byref(h)) != 0
    assert b"Basis" in lib.cuddh_b200_last_error()
    with pytest.raises(cb.CuddhError):
        cb.Mesh2D.uniform_rect(0, -1.0, 1.0, 4, -1.0, 1.0)
    mesh = cb.Mesh2D.uniform_rect(6, -1.0, 1.0, 6, -1.0, 1.0)
    fem = cb.H1Space(mesh, cb.Basis(5))
    with pytest.raises(cb.CuddhError, match="n_basis==4"):  # source/DDH.cpp:333-334
        cb.DDH(10.0, np.ones(fem.size()), fem, 6, 6)
    fem4 = cb.H1Space(mesh, cb.Basis(4))
    with pytest.raises(cb.CuddhError, match="multiples"):   # source/DDH.cpp:338-339
        cb.DDH(10.0, np.ones(fem4.size()), fem4, 6, 6)


@pytest.mark.parametrize("nb,nq", TABLE_PAIRS)
def test_tables(gold, nb, nq):
    b = cb.Basis(nb)
    bx, bw = b.quadrature()
    gl = cb.QuadratureRule(nq, cb.GaussLegendre)
    gll = cb.QuadratureRule(nq, cb.GaussLobatto)
    mine = {"gll_x": bx, "gll_w": bw, "Dnodes": b.deriv(bx).T.ravel(), "gl_x": gl.x(), "gl_w": gl.w(),
            "P": b.eval(gl.x()).T.ravel(), "D": b.deriv(gl.x()).T.ravel(), "gll_nq_x": gll.x(), "gll_nq_w": gll.w()}
    for k, v in mine.items():
        ref = gold.tables["tables_%d_%d_%s" % (nb, nq, k)]
        lit = (nq <= 10) if k in ("gl_x", "gl_w", "P", "D") else (nq <= 9) if k.startswith("gll_nq") else True
        if lit:   # literal node tables: bit-exact
            assert np.array_equal(v, ref), k
        else:     # Golub-Welsch sizes: own eigen-solver instead of LAPACK dsteqr_, nodes agree to 1 ulp
            assert np.max(np.abs(v - ref)) <= 4e-16 * max(1.0, np.max(np.abs(ref))) * (1 if "x" in k or "w" in k else 200), k


def _meshes(tag):
    if tag.startswith("rect"):
        nx = int(tag[4:])
        return cb.Mesh2D.uniform_rect(nx, -1.0, 1.0, nx, -1.0, 1.0)
    xy, el = load_mesh_file()
    return cb.Mesh2D.from_vertices(xy, el)


@pytest.mark.parametrize("tag,nb", [("rect2", 3), ("rect3", 2), ("rect10", 4), ("rect10", 5), ("rect6", 8), ("unstr", 2), ("unstr", 4),
                                    ("unstr", 5), ("unstr", 9)])
def test_mesh_h1_facespace_bit_exact(gold, tag, nb):
    mesh = _meshes(tag)
    g = lambda k: gold.h1["h1_%s_%d_%s" % (tag, nb, k)]
    fem = cb.H1Space(mesh, cb.Basis(nb))
    fs = cb.FaceSpace(fem, mesh.boundary_edges())
    assert fem.size() == g("ndof")[0]
    assert np.array_equal(mesh.edges().ravel(), g("edges"))
    assert np.array_equal(mesh.boundary_edges(), g("boundary_edges"))
    assert np.array_equal(fem.global_indices().ravel(), g("I"))
    assert np.array_equal(fem.physical_coordinates().ravel(), g("xy"))
    assert fs.size() == g("fdof")[0]
    assert np.array_equal(fs.subspace_indices().ravel(), g("face_I"))
    assert np.array_equal(fs.global_indices(), g("face_proj"))
    assert mesh.min_h() == g("min_h")[0] and mesh.max_h() == g("max_h")[0]


@pytest.mark.parametrize("key", ["h1_rect64_5", "h1_rect128_4", "h1_rect256_5", "h1_rect64_9", "h1_rect512_4", "h1_rect1024_5"])
def test_h1_hashes_larger(gold, key):
    _, tag, nb = key.split("_")
    nx, nb = int(tag[4:]), int(nb)
    h = gold.hashes[key]
    mesh = cb.Mesh2D.uniform_rect(nx, -1.0, 1.0, nx, -1.0, 1.0)
    fem = cb.H1Space(mesh, cb.Basis(nb))
    fs = cb.FaceSpace(fem, mesh.boundary_edges())
    assert fem.size() == h["ndof"] == (nx * (nb - 1) + 1) ** 2
    assert mesh.n_edges() == h["n_edges"]
    assert fnv1a64(mesh.edges()) == h["edges"]
    assert fnv1a64(fem.global_indices()) == h["I"]
    assert fnv1a64(fem.physical_coordinates()) == h["xy"]
    assert fnv1a64(fs.subspace_indices()) == h["face_I"] and fnv1a64(fs.global_indices()) == h["face_proj"]


@pytest.mark.parametrize("nx,ny,nb", [(1, 1, 2), (1, 5, 3), (7, 1, 4), (13, 9, 5), (5, 11, 9), (33, 17, 2)])
def test_uniform_rect_closed_form_equals_generic_path(monkeypatch, nx, ny, nb):
    # Mesh2D::uniform_rect / H1Space take closed forms of the edge table and of the first-touch numbering on structured meshes
    # (csrc/mesh.cpp, csrc/h1space.cpp); CUDDH_B200_CLOSED_FORM=0 forces the generic from_vertices / first-touch scan
    # (source/Mesh2D.cpp:59-171, source/H1Space.cpp:11-127). Every array must be bit-identical, also on non-square, non-symmetric boxes.
    def build():
        mesh = cb.Mesh2D.uniform_rect(nx, -0.3, 1.7, ny, 0.1, 0.9)
        fem = cb.H1Space(mesh, cb.Basis(nb))
        fs = cb.FaceSpace(fem, mesh.boundary_edges())
        return [mesh.edges().copy(), mesh.boundary_edges().copy(), np.array([mesh.min_h(), mesh.max_h(), mesh.n_edges(), fem.size()]),
                fem.global_indices().copy(), fem.physical_coordinates().copy(), fs.subspace_indices().copy(), fs.global_indices().copy()]
    monkeypatch.setenv("CUDDH_B200_CLOSED_FORM", "0")
    generic = build()
    monkeypatch.delenv("CUDDH_B200_CLOSED_FORM")
    closed = build()
    for a, b in zip(generic, closed):
        assert a.shape == b.shape and np.array_equal(a, b)


def test_uniform_rect_2048_is_not_corrupt():
    # the reference's 32-bit edge key corrupts uniform_rect(2048) (6 293 504 edges, 73 413 630 DOFs).
    # Closed forms: edges = 2 nx (nx+1); ndof = (nx (nb-1) + 1)^2. Checked at 2048 x 64 to keep the CPU suite short,
    # where the same key formula (min + nv*max with nv = 2049*65) already overflows int32.
    nx, ny, nb = 2048, 64, 5
    mesh = cb.Mesh2D.uniform_rect(nx, -1.0, 1.0, ny, -1.0, 1.0)
    assert mesh.n_edges() == nx * (ny + 1) + ny * (nx + 1)
    assert mesh.n_edges("boundary") == 2 * (nx + ny)
    fem = cb.H1Space(mesh, cb.Basis(nb))
    assert fem.size() == (nx * (nb - 1) + 1) * (ny * (nb - 1) + 1)
    I = fem.global_indices()
    assert I.min() == 0 and I.max() == fem.size() - 1
    # first-touch rule: the ids first seen in element e are consecutive and start where element e-1 stopped
    flat = I.reshape(len(I), -1)
    order = np.unique(flat.ravel(), return_index=True)[1]
    assert np.all(np.diff(order) > 0)  # id k first appears before id k+1 in the volume-index scan


@pytest.mark.parametrize("nx,nb,block", [(8, 4, 16), (16, 4, 16), (8, 8, 16), (16, 4, 32), (8, 8, 32)])
def test_ddh_setup_matches_oracle(gold, nx, nb, block):
    om = S.uniform_rect(nx, -1.0, 1.0, nx, -1.0, 1.0)
    ob = S.Basis(nb)
    I3, ndof, _ = S.h1space(om, ob)
    ha = 0.5 + np.random.default_rng(1).random(ndof)
    od = S.ddh_setup(10.0, ha, om, ob, I3, ndof, nx, nx, block)
    mesh = cb.Mesh2D.uniform_rect(nx, -1.0, 1.0, nx, -1.0, 1.0)
    fem = cb.H1Space(mesh, cb.Basis(nb))
    d = cb.DDH(10.0, ha, fem, nx, nx, block)
    info = d.info()
    assert (info["n_domains"], info["n_shared"], info["nt"], info["dt"]) == (od.n_domains, od.n_shared, od.nt, od.dt)
    assert d.size() == 4 * od.n_shared  # DDH::size() = 2 * n_lambda, n_lambda = 2 * n_shared
    for name, ref in [("B", od.B), ("gI", od.gI), ("sI", od.sI), ("cmap", od.en.cmap), ("m", od.m), ("H", od.H), ("a", od.a),
                      ("gmi", od.gmi), ("wh_filter", od.wh_filter), ("cs", od.cs), ("sn", od.sn), ("D", od.D),
                      ("ens_gI", od.en.gI), ("ens_sI", od.en.sI), ("ens_fI", od.en.fI), ("ens_pI", od.en.pI)]:
        assert np.array_equal(d.array(name), np.asarray(ref).ravel()), name
    if block == 16 and (nx, nb) in [(8, 4), (16, 4), (8, 8)]:  # straight against the reference's EnsembleSpace
        g = lambda k: gold.ens["ens_%d_%d_%s" % (nx, nb, k)]
        assert np.array_equal(d.array("cmap"), g("cmap"))
        assert np.array_equal(d.array("ens_gI"), g("gI")) and np.array_equal(d.array("ens_sI"), g("sI"))
        assert np.array_equal(d.array("ens_fI"), g("fI")) and np.array_equal(d.array("ens_pI"), g("pI"))


@pytest.mark.parametrize("key", ["ens_32_4", "ens_64_4", "ens_32_8"])
def test_ddh_ensemble_hashes(gold, key):
    _, nx, nb = key.split("_")
    nx, nb = int(nx), int(nb)
    h = gold.hashes[key]
    mesh = cb.Mesh2D.uniform_rect(nx, -1.0, 1.0, nx, -1.0, 1.0)
    fem = cb.H1Space(mesh, cb.Basis(nb))
    d = cb.DDH(10.0, np.ones(fem.size()), fem, nx, nx, 16)
    info = d.info()
    assert info["n_domains"] == h["n_domains"] and info["n_shared"] == h["n_shared"]
    assert fnv1a64(d.array("cmap")) == h["cmap"]
    assert fnv1a64(d.array("ens_gI")) == h["gI"] and fnv1a64(d.array("ens_sI")) == h["sI"]
    assert fnv1a64(d.array("ens_fI")) == h["fI"] and fnv1a64(d.array("ens_pI")) == h["pI"]


def test_vectors_must_be_device_pointers():
    import torch
    mesh = cb.Mesh2D.uniform_rect(4, -1.0, 1.0, 4, -1.0, 1.0)
    fem = cb.H1Space(mesh, cb.Basis(4))
    fs = cb.FaceSpace(fem, mesh.boundary_edges())
    x = torch.zeros(fem.size(), dtype=torch.float64)
    with pytest.raises(cb.CuddhError, match="DEVICE"):
        fs.restrict(x, x)


def test_ensemble_c_abi_matches_reference(gold):
    # EnsembleSpace through its own C-ABI entry points (what cuddhelmholtz_b200/cxx/include/EnsembleSpace.hpp calls)
    import ctypes as C
    lib = capi.load()
    nx, nb = 8, 4
    mesh = cb.Mesh2D.uniform_rect(nx, -1.0, 1.0, nx, -1.0, 1.0)
    fem = cb.H1Space(mesh, cb.Basis(nb))
    lab, nd = S.ddh_labels(nx, nx, nb)
    lab = np.ascontiguousarray(lab, np.int32)
    h = C.c_void_p()
    capi.check(lib.cuddh_b200_ensemble_create(fem._h, nd, lab.ctypes.data_as(C.c_void_p), C.byref(h)))
    try:
        info = np.zeros(6, np.int64)
        capi.check(lib.cuddh_b200_ensemble_info(h, info.ctypes.data_as(C.c_void_p)))
        assert info[0] == nd and info[5] == 52
        for name, key in [("gI", "gI"), ("sizes", "sizes"), ("elements", "elements"), ("faces", "faces"), ("sI", "sI"), ("fI", "fI"),
                          ("pI", "pI"), ("fsizes", "fsizes"), ("cmap", "cmap"), ("n_elems", "n_elems"), ("n_faces", "n_faces")]:
            cnt = C.c_int64()
            ptr = lib.cuddh_b200_ensemble_array(h, name.encode(), C.byref(cnt))
            arr = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_int)), shape=(cnt.value,)).copy()
            assert np.array_equal(arr, gold.ens["ens_%d_%d_%s" % (nx, nb, key)]), name
        assert lib.cuddh_b200_ensemble_array(h, b"nope", None) is None
        # a label outside [0, n_spaces) is the reference's "illogically labeled" error
        bad = lab.copy()
        bad[0] = nd
        h2 = C.c_void_p()
        assert lib.cuddh_b200_ensemble_create(fem._h, nd, bad.ctypes.data_as(C.c_void_p), C.byref(h2)) != 0
        assert b"illogically" in lib.cuddh_b200_last_error()
    finally:
        lib.cuddh_b200_ensemble_destroy(h)


@pytest.mark.parametrize("tag,nb", [("rect2", 3), ("rect10", 2), ("rect10", 4), ("rect10", 5), ("rect37", 5), ("rect37", 4), ("rect6", 8),
                                    ("unstr", 2), ("unstr", 3), ("unstr", 4), ("unstr", 5), ("unstr", 9)])
@pytest.mark.parametrize("node_major", [0, 1])
def test_assembly_plan_self_check(tag, nb, node_major):
    """the deterministic assembly plans behind Operator::action (both layouts) reproduce the plain sum over global_indices:
    every element node is either written once by its element (interior nodes, node-major plan) or listed exactly once per
    patch with all its contributions; shared DOFs get one partial slot per touching patch. Host only (no GPU)."""
    mesh = cb.Mesh2D.uniform_rect(37, -1.0, 1.0, 29, 0.0, 2.0) if tag == "rect37" else _meshes(tag)
    fem = cb.H1Space(mesh, cb.Basis(nb))
    st = fem.check_plan(node_major)
    assert st["mismatches"] == 0, st
    nel = mesh.n_elem()
    assert st["n_patches"] * st["patch_elems"] >= nel
    if node_major:
        assert st["patch_elems"] % 32 == 0
        # only element-boundary DOFs are listed: strictly fewer than all DOFs as soon as elements have interior nodes
        if nb > 2:
            assert st["listed_dofs"] - st["shared_dofs"] < fem.size()
    if tag == "unstr":
        assert st["dofs_over_four"] >= 0  # the unstructured mesh has valence-5 vertices: the overflow path is exercised when > 0


@pytest.mark.parametrize("nx,nb,world", [(16, 4, 2), (32, 4, 4), (32, 4, 3), (8, 8, 2)])
def test_neighbour_ddh_slot_partition(nx, nb, world):
    """host logic of parallel.NeighbourDDH (no GPU): from the library's B table every lambda slot gets one owning rank, a
    rank's subdomains read owned slots only, write owned or adjacent-slab slots only, and the send / recv lists of
    neighbouring ranks mirror each other."""
    import torch
    from cuddhelmholtz_b200.parallel import NeighbourDDH, subdomain_range
    mesh = cb.Mesh2D.uniform_rect(nx, -1.0, 1.0, nx, -1.0, 1.0)
    fem = cb.H1Space(mesh, cb.Basis(nb))
    ha = np.ones(fem.size())
    d = cb.DDH(10.0, ha, fem, nx, nx, 16)
    nd, n = d.info()["n_domains"], d.size()
    nl = n // 2
    B = d.array("B").reshape(nd, 2, -1)
    R = [NeighbourDDH(d, r, world, device="cpu") for r in range(world)]
    assert torch.equal(sum(r.mask for r in R), torch.ones(n))
    total_sent = 0
    for r in R:
        a, b = subdomain_range(nd, r.rank, world)
        reads = B[a:b, 0, :]
        writes = B[a:b, 1, :]
        assert np.all(r.owner[reads[reads >= 0]] == r.rank)
        assert np.all(np.abs(r.owner[writes[writes >= 0]] - r.rank) <= 1)
        for q, idx in r.send_idx.items():
            assert abs(q - r.rank) == 1 and torch.equal(idx, R[q].recv_idx[r.rank])
            assert np.all(r.owner[idx.numpy()[: idx.numel() // 2]] == q)  # first half: lambda block, second: + n_lambda
            assert torch.equal(idx[idx.numel() // 2:], idx[: idx.numel() // 2] + nl)
            total_sent += idx.numel()
        assert set(r.recv_idx) <= {r.rank - 1, r.rank + 1}
    # only slab-boundary traces travel: a small fraction of the vector
    assert 0 < total_sent < n // 2


@pytest.mark.parametrize("nx,nb,world", [(16, 4, 2), (32, 4, 4), (32, 4, 3), (8, 8, 2), (32, 4, 1)])
def test_library_ddh_partition_matches_python(nx, nb, world):
    """host tables of the library's distributed DDH (csrc/ddh_setup.cpp: DdhDist, no GPU, no NCCL) against the numpy construction
    of parallel.NeighbourDDH: owner of every slot, subdomain ranges, send / recv lists per neighbour (mirrored, ascending)."""
    from cuddhelmholtz_b200.parallel import NeighbourDDH, subdomain_range
    mesh = cb.Mesh2D.uniform_rect(nx, -1.0, 1.0, nx, -1.0, 1.0)
    fem = cb.H1Space(mesh, cb.Basis(nb))
    d = cb.DDH(10.0, np.ones(fem.size()), fem, nx, nx, 16)
    nd, n = d.info()["n_domains"], d.size()
    L = [cb.DDHDist(d, None, r, world) for r in range(world)]
    R = [NeighbourDDH(d, r, world, device="cpu") for r in range(world)]
    owned = 0
    for r in range(world):
        info = L[r].info()
        assert (info["dom_begin"], info["dom_end"]) == subdomain_range(nd, r, world) and info["size"] == n
        assert np.array_equal(L[r].array("owner"), R[r].owner)
        assert info["n_owned"] == int((R[r].owner == r).sum())
        owned += info["n_owned"]
        seg = L[r].array("segments").reshape(-1, 5)
        snd, rcv = L[r].array("send_idx"), L[r].array("recv_idx")
        assert info["n_send"] == len(snd) and info["n_recv"] == len(rcv) and info["bytes_per_action"] == 8 * len(snd)
        peers = sorted(set(R[r].send_idx) | set(R[r].recv_idx))
        assert list(seg[:, 0]) == peers
        for q, so, sc, ro, rc in seg:
            want_s = R[r].send_idx[q].numpy()[: R[r].send_idx[q].numel() // 2] if q in R[r].send_idx else np.zeros(0, np.int64)
            want_r = R[r].recv_idx[q].numpy()[: R[r].recv_idx[q].numel() // 2] if q in R[r].recv_idx else np.zeros(0, np.int64)
            assert np.array_equal(snd[so:so + sc], want_s) and np.array_equal(rcv[ro:ro + rc], want_r)
            # what r packs for q is what q expects from r, in the same order
            segq = L[q].array("segments").reshape(-1, 5)
            row = segq[list(segq[:, 0]).index(r)]
            assert np.array_equal(L[q].array("recv_idx")[row[3]:row[3] + row[4]], snd[so:so + sc])
    assert owned == n // 2


def test_gmres_options_and_knobs():
    """option plumbing of the C ABI (no GPU): defaults, set / get, unknown names, NCCL discovery answers without a device"""
    assert cb.get_option("gmres_orth") == cb.MGS and cb.get_option("max_ctas") == 0
    cb.set_option("gmres_orth", cb.CGS2)
    cb.set_option("max_ctas", 8)
    assert cb.get_option("gmres_orth") == cb.CGS2 and cb.get_option("max_ctas") == 8
    cb.set_option("gmres_orth", cb.MGS)
    cb.set_option("max_ctas", 0)
    assert cb.get_option("nccl_available") in (0, 1) and cb.get_option("nope") == -1
    with pytest.raises(cb.CuddhError):
        cb.set_option("nope", 1)


def test_assembly_plan_hashes_pinned():
    """every array of both assembly plans, bit for bit, against tests/golden/plan_hashes.json (scripts/make_plan_hashes.py):
    the kernels consume these arrays verbatim, so an unintended change of the plan builder shows up here without a GPU"""
    import json
    import os
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "plan_hashes.json")))
    xy, el = load_mesh_file()
    mk = {"unstr": lambda: cb.Mesh2D.from_vertices(xy, el), "r37": lambda: cb.Mesh2D.uniform_rect(37, -1.0, 1.0, 29, 0.0, 2.0),
          "r10": lambda: cb.Mesh2D.uniform_rect(10, -1.0, 1.0, 10, -1.0, 1.0), "r3": lambda: cb.Mesh2D.uniform_rect(3, -1.0, 1.0, 2, -1.0, 1.0)}
    checked = 0
    for key, h in gold.items():
        tag, nb, k = key.split("_")
        if tag not in mk:
            continue  # the 256^2 cases are generator-only (seconds each)
        fem = cb.H1Space(mk[tag](), cb.Basis(int(nb)))
        assert fem.check_plan(int(k))["hash"] == h, key
        checked += 1
    assert checked >= 40


def test_thread_pair_algebra_matches_direct_formulas():
    """host replay of the mirrored-frame thread-pair split of volume_action_pair (csrc/volume_pair.cuh) against the direct
    sum-factorised formulas of source/StiffnessMatrix.cpp:132-182 / source/MassMatrix.cpp:170-205, every (n_basis, n_quad) rule"""
    import importlib.util
    spec = importlib.util.spec_from_file_location("check_pair_algebra", os.path.join(os.path.dirname(__file__), "..", "scripts", "check_pair_algebra.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert mod.main() < 1e-13


def test_sass_table_operands_stay_uniform():
    """SASS of the built library: in the compute branch (after USETMAXREG.TRY_ALLOC) of every warp-specialised volume kernel the 1-D
    table operands must come through the uniform datapath (LDCU), not through per-thread constant loads (LDC): a change that costs
    ptxas the uniformity of the quadrature-row loop (seen once with a run-time choice between two aligned cluster-barrier forms)
    turns ~200 LDCU.64 into LDC.64, spills, and slows the kernel by 30-60 % without changing a single result."""
    import shutil, subprocess
    so = os.path.join(os.path.dirname(cb.__file__), "lib", "libcuddh_b200.so")
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe) or not os.path.exists(so):
        pytest.skip("cuobjdump or the built library is not available")
    text = subprocess.run([exe, "-sass", so], capture_output=True, text=True, timeout=600).stdout
    # the default instances behind the bench numbers: fused Helmholtz (affine / stored metric) n_basis 5 and 4, stand-alone
    # stiffness / weighted mass n_basis 5, thread-pair kernels n_basis 8
    wanted = ["volume_action_wsILi5ELi6ELb1ELi9ELin5ELb1E", "volume_action_wsILi4ELi5ELb1ELi8ELin4ELb1E", "volume_action_wsILi5ELi6ELb1ELi9ELi5ELb0E",
              "volume_action_wsILi4ELi5ELb1ELi8ELin5ELb0E", "volume_action_wsILi5ELi6ELb1ELi9ELin5ELb0E", "volume_action_wsILi5ELi6ELb1ELi0ELi0ELb1E", "volume_action_wsILi5ELi6ELb1ELi0ELi5ELb0E",
              "volume_action_wsILi5ELi9ELb0ELi0ELi5ELb0E", "volume_action_pairILi8ELi9ELb1ELb1E", "volume_action_pairILi8ELi9ELb1ELb0E",
              "volume_action_pairILi8ELi14ELb0ELb0E"]
    checked = set()
    for block in text.split("Function : ")[1:]:
        name = block.split("\n", 1)[0]
        key = next((w for w in wanted if w in name), None)
        if key is None:
            continue
        assert "USETMAXREG.TRY_ALLOC" in block, name
        compute = block.split("USETMAXREG.TRY_ALLOC", 1)[1]
        n_dfma = len(re.findall(r"\bDFMA\b", compute))
        n_ldc = len(re.findall(r"\bLDC\.64\b", compute))
        n_ldcu = len(re.findall(r"\bLDCU\.(?:64|128)\b", compute))
        n_spill = len(re.findall(r"\bSTL\b", compute))
        assert n_dfma >= 140 and n_ldc <= 32 and n_ldcu >= n_dfma // 4 and n_spill <= 4, (name, n_dfma, n_ldc, n_ldcu, n_spill)
        checked.add(key)
    assert checked == set(wanted), sorted(set(wanted) - checked)
