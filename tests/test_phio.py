"""phastaIO POSIX fixture format (phasta_b200/phio.py): the reference's own phIO tests restated
(phSolver/common/test/phIOwrite.cc, phIOwriteReadZeroSz.cc, phIOposixMultiTopo.cc, phIOreadIlwork.cc),
geombc/restart round trips of synthetic parts, genBC1 against the reference's genbc1.f executed by f77np
(tests/golden/f77_genbc1.npz), and the solver on a part that went through the files."""
import os
import sys

import numpy as np
import pytest

from common import make_case, make_oracle, rel_l2
from phasta_b200 import phio

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_cscompare_is_the_references_prefix_match():
    assert phio.cscompare("number of nodes", "number of nodes ")
    assert phio.cscompare("Number Of Nodes", "numberofnodes")
    assert phio.cscompare("number of nodes", "number of nodes with Dirichlet BCs")   # prefix: file order matters
    assert not phio.cscompare("number of nodes with Dirichlet BCs", "number of nodes")
    assert phio.cscompare("connectivity interior?", "connectivity interior linear tetrahedron ")
    assert not phio.cscompare("connectivity boundary?", "connectivity interior linear tetrahedron ")
    assert phio.cscompare("nbc codes?", "nbc codes linear tetrahedron")


def test_write_then_read_number_of_fishes(tmp_path):
    """phIOwrite.cc: one header with one int, one double of data."""
    p = str(tmp_path / "water.dat.1")
    with phio.PhioFile(p, "w") as f:
        f.writeheader("number of fishes", [2], 1, "double")
        f.writedatablock("number of fishes", np.array([1.23]), "double")
    raw = open(p, "rb").read()
    assert b"number of fishes : < 9 > 2 \n" in raw          # 8 data bytes + newline (phastaIO.cc:1636-1643)
    assert b"byteorder magic number : < 5 > 1 \n" in raw
    with phio.PhioFile(p, "r") as f:
        assert f.readheader("number of fishes", 1, "double") == [2]
        assert f.readdatablock("number of fishes", 1, "double")[0] == 1.23


def test_zero_size_block(tmp_path):
    """phIOwriteReadZeroSz.cc: header with zero data items, data block call with zero items."""
    p = str(tmp_path / "water.dat.1")
    with phio.PhioFile(p, "w") as f:
        f.writeheader("number of fishes", [0], 0, "double")
        f.writedatablock("number of fishes", np.zeros(0), "double")
    assert b"number of fishes : < 0 > 0 \n" in open(p, "rb").read()
    with phio.PhioFile(p, "r") as f:
        assert f.readheader("number of fishes", 1, "double") == [0]
        assert f.readdatablock("number of fishes", 0, "double").size == 0


def test_headers_found_out_of_order_and_missing_key(tmp_path):
    """readHeader searches forward, skips other blocks by their byte count and wraps once (phastaIO.cc:239-296)."""
    p = str(tmp_path / "a.dat.1")
    with phio.PhioFile(p, "w") as f:
        f.writeheader("alpha", [3], 3, "integer")
        f.writedatablock("alpha", np.array([1, 2, 3]), "integer")
        f.writeheader("beta", [2, 7], 2, "double")
        f.writedatablock("beta", np.array([0.5, -0.25]), "double")
        f.writeheader("gamma", [5], 0, "integer")
    with phio.PhioFile(p, "r") as f:
        assert f.readheader("gamma", 1) == [5]
        assert f.readheader("beta", 2, "double") == [2, 7]        # behind the cursor: found after the rewind
        assert np.array_equal(f.readdatablock("beta", 2, "double"), [0.5, -0.25])
        assert f.readheader("alpha", 1) == [3]
        assert np.array_equal(f.readdatablock("alpha", 3, "integer"), [1, 2, 3])
        assert f.readheader("no such field", 1) is None
        with pytest.raises(IOError):
            f.readdatablock("alpha", 3, "integer")                # data block without its header


def test_write_sequence_errors(tmp_path):
    with phio.PhioFile(str(tmp_path / "b.dat.1"), "w") as f:
        f.writeheader("alpha", [3], 3, "integer")
        with pytest.raises(IOError):
            f.writedatablock("beta", np.array([1, 2, 3]), "integer")
        f.writeheader("alpha", [3], 3, "integer")
        with pytest.raises(IOError):
            f.writedatablock("alpha", np.array([1, 2]), "integer")


@pytest.mark.parametrize("bc,nparts,boundary,topo", [("channel", 1, True, "tet"), ("mixed", 2, False, "tet"),
                                                     ("channel", 2, True, "tet"), ("channel", 1, False, "mixed"),
                                                     ("none", 1, False, "hex"), ("channel", 1, True, "hex"),
                                                     ("channel", 2, True, "mixed"), ("none", 1, True, "wedge")])
def test_geombc_restart_round_trip(tmp_path, bc, nparts, boundary, topo):
    case = make_case(8, 6, 5, nparts=nparts, bc=bc, boundary=boundary, natural="all" if boundary else "none",
                     topo=topo, periodic_z=(bc != "none"))
    _, _, parts, states = case
    d = str(tmp_path)
    for p, (y, ac) in zip(parts, states):
        path = phio.write_geombc(p, d)
        assert path.endswith("%d-procs_case/geombc.dat.%d" % (nparts, p.rank + 1))
        q = phio.read_geombc(d, p.rank, p.numpe, 64)
        assert (q.nshg, q.numnp, q.numel) == (p.nshg, p.numnp, p.numel)
        assert np.array_equal(q.x, p.x)
        assert np.array_equal(q.lcblk, p.lcblk)                       # bit-exact INT targets
        assert len(q.mien) == len(p.mien) and all(np.array_equal(a, b) for a, b in zip(q.mien, p.mien))
        assert np.array_equal(q.iBC, p.iBC) and np.array_equal(q.iper, p.iper)
        assert np.array_equal(q.ilwork, p.ilwork)
        act = phio.active_bc_mask(p.iBC)
        tol = 0.0 if bc != "mixed" else 4e-16
        assert np.all(np.abs(q.BC - p.BC)[act] <= tol * np.abs(p.BC)[act])
        assert not q.BC[~act].any()                                   # genbc.f:52 `BC = zero` outside the codes
        if p.nelblb:
            assert np.array_equal(q.lcblkb, p.lcblkb)
            for a, b in zip(q.mienb + q.miBCB + q.mBCB, p.mienb + p.miBCB + p.mBCB):
                assert np.array_equal(a, b)
        phio.write_restart(d, p.rank, p.numpe, 120, y, ac)
        y2, ac2, lstep = phio.read_restart(d, p.rank, p.numpe, p.nshg)
        assert lstep == 120 and np.array_equal(y, y2) and np.array_equal(ac, ac2)
    assert open(os.path.join(phio.case_dir(d, nparts), "numstart.dat")).read().split() == ["120"]


def test_multi_topology_blocks_read_in_file_order(tmp_path):
    """phIOposixMultiTopo.cc: two 'connectivity interior' blocks are found by repeating the same wildcard read."""
    _, _, parts, _ = make_case(6, 6, 4, topo="mixed")
    path = phio.write_geombc(parts[0], str(tmp_path))
    seen = []
    with phio.PhioFile(path, "r") as f:
        for _ in range(2):
            h = f.readheader("connectivity interior?", 7)
            assert h[0] > 0 and h[3] > 0
            f.readdatablock("connectivity interior?", h[0] * h[3], "integer")
            seen.append((h[6], h[3]))
    assert seen == [(1, 4), (3, 6)]


def test_ilwork_in_the_file_is_one_based(tmp_path):
    """phIOreadIlwork.cc reads the raw array; ctypes.f:47 makes `iother` 0-based afterwards."""
    _, _, parts, _ = make_case(8, 4, 4, nparts=4)
    p = parts[1]
    path = phio.write_geombc(p, str(tmp_path))
    with phio.PhioFile(path, "r") as f:
        n = f.readheader("size of ilwork array", 1)[0]
        f.readheader("ilwork", 1)
        il = f.readdatablock("ilwork", n, "integer")
    assert il[0] == 2 and il[3] == p.ilwork[3] + 1
    assert np.array_equal(phio.read_geombc(str(tmp_path), 1, 4, 64).ilwork, p.ilwork)


def test_genbc1_matches_the_reference_fortran():
    g = np.load(os.path.join(GOLD, "f77_genbc1.npz"))
    assert np.array_equal(phio.genBC1(g["BCtmp"], g["iBC"]), g["BC"])


def test_bcinp_inverse_round_off():
    g = np.load(os.path.join(GOLD, "f77_genbc1.npz"))
    iBC, BC = g["iBC"], g["BC"]
    BC2 = phio.genBC1(phio.bcinp_from_BC(iBC, BC), iBC)
    act = phio.active_bc_mask(iBC)
    assert np.all(np.abs(BC2 - BC)[act] <= 1e-15 * np.abs(BC)[act])


def test_oracle_on_a_part_read_from_files_is_identical(tmp_path):
    """configs[0] route: synthetic geombc/restart files -> reader -> one ElmGMRe; equal to the in-memory part."""
    case = make_case(7, 5, 4, bc="channel", boundary=True, natural="all")
    params, tables, parts, states = case
    phio.write_geombc(parts[0], str(tmp_path))
    phio.write_restart(str(tmp_path), 0, 1, 0, *states[0])
    q = phio.read_geombc(str(tmp_path), 0, 1, params.ibksiz)
    y, ac, _ = phio.read_restart(str(tmp_path), 0, 1, q.nshg)
    o1 = make_oracle(case)
    o2 = make_oracle((params, tables, [q], [(y, ac)]))
    o1.ElmGMRe()
    o2.ElmGMRe()
    assert np.array_equal(o1.parts[0].res, o2.parts[0].res)
    assert np.array_equal(o1.parts[0].BDiag, o2.parts[0].BDiag)


@pytest.mark.gpu
def test_gpu_solve_from_files_c1_cube(tmp_path):
    """BASELINE.json configs[0]: compressible linear-tet cube, 20x20x21x6 = 50 400 elements, written as
    phastaIO geombc/restart, read back, one implicit step (assembly + EBE GMRES) on the GPU vs the oracle."""
    from phasta_b200.solver import PhastaGPU
    case = make_case(20, 20, 21, bc="channel", etol=1e-6)
    params, tables, parts, states = case
    phio.write_geombc(parts[0], str(tmp_path))
    phio.write_restart(str(tmp_path), 0, 1, 0, *states[0])
    q = phio.read_geombc(str(tmp_path), 0, 1, params.ibksiz)
    assert q.numel == 50400 and np.array_equal(q.lcblk, parts[0].lcblk)
    y, ac, _ = phio.read_restart(str(tmp_path), 0, 1, q.nshg)
    g = PhastaGPU(q, params, tables, device=0)
    res, Dy = g.SolGMRe(y, ac)
    o = make_oracle(case)
    iKs, _ = o.SolGMRe()
    assert g.iKs == iKs
    assert rel_l2(g.rmes, o.parts[0].rmes) < 1e-10
    assert rel_l2(Dy, o.parts[0].Dy) < 1e-8
    g.close()


# ---- a28: genblk / gensav / lcblk / mien against the reference's own genblkPosix.f + gensav.f -----------------------
def _genblk_cases():
    sys.path.insert(0, GOLD)
    from make_golden_genblk import CASES
    return list(CASES)


@pytest.mark.parametrize("name", _genblk_cases())
def test_blocks_match_the_executed_genblkposix(tmp_path, name):
    """tests/golden/f77_genblk.npz holds lcblk and every mien(iblk)%p as the UNMODIFIED common/genblkPosix.f +
    gensav.f (run by f77np over this repo's geombc file) leave them; the mesh generator's own blocking
    (mesh._blocks) and the file reader (phio.read_geombc) must give the same integers, block for block."""
    from make_golden_genblk import build_case
    z = np.load(os.path.join(GOLD, "f77_genblk.npz"))
    (params, tables, parts, states), ibksz = build_case(name)
    p = parts[0]
    q = phio.read_geombc(os.path.dirname(os.path.dirname(phio.write_geombc(p, str(tmp_path)))), 0, 1, ibksz)
    nelblk = int(z["%s_nelblk" % name])
    for mp in (p, q):
        assert mp.nelblk == nelblk
        assert np.array_equal(mp.lcblk, z["%s_lcblk" % name])
        for i in range(nelblk):
            assert np.array_equal(mp.mien[i], z["%s_mien_%d" % (name, i)])
    if name == "mixed":      # tets first, then wedges: lcsyst 1 ... 3, a ragged last block per topology
        assert list(np.unique(z["mixed_lcblk"][2, :-1])) == [1, 3]


def _genbkb_cases():
    sys.path.insert(0, GOLD)
    from make_golden_genblk import BCASES
    return list(BCASES)


@pytest.mark.parametrize("name", _genbkb_cases())
def test_boundary_blocks_match_the_executed_genbkbposix(tmp_path, name):
    """the boundary-element variant: common/genbkbPosix.f + gensvb.f (connectivity boundary / nbc codes / nbc values,
    the zeroing of unset flux values, blocking by IBKSZ) executed by f77np over this repo's geombc file -> lcblkb and
    every mienb / miBCB / mBCB block; mesh._boundary_elements and phio.read_geombc must give the same"""
    from make_golden_genblk import build_case
    z = np.load(os.path.join(GOLD, "f77_genblk.npz"))
    (params, tables, parts, states), ibksz = build_case(name)
    p = parts[0]
    q = phio.read_geombc(os.path.dirname(os.path.dirname(phio.write_geombc(p, str(tmp_path)))), 0, 1, ibksz)
    n = int(z["%s_nelblb" % name])
    assert n > 0
    for mp in (p, q):
        assert mp.nelblb == n
        assert np.array_equal(mp.lcblkb, z["%s_lcblkb" % name])
        for i in range(n):
            assert np.array_equal(mp.mienb[i], z["%s_mienb_%d" % (name, i)])
            assert np.array_equal(mp.miBCB[i], z["%s_mibcb_%d" % (name, i)])
            assert np.array_equal(mp.mBCB[i], z["%s_mbcb_%d" % (name, i)])


def test_genblk_fixture_reproduces_from_the_reference():
    if not os.path.isdir("/root/reference/phSolver/common"):
        pytest.skip("reference sources not present")
    from make_golden_genblk import generate
    z = np.load(os.path.join(GOLD, "f77_genblk.npz"))
    new = generate()
    assert set(new) == set(z.files)
    assert all(np.array_equal(new[k], z[k]) for k in new)


# ---- f-2: the file format against the reference's own reader, readnblk.f (executed by f77np over this repo's files) --
def _readnblk_cases():
    sys.path.insert(0, GOLD)
    from make_golden_readnblk import CASES
    return list(CASES)


@pytest.mark.parametrize("name", _readnblk_cases())
def test_files_read_by_the_executed_readnblk_give_what_the_repo_reader_gives(tmp_path, name):
    """tests/golden/f77_readnblk.npz: what the UNMODIFIED common/readnblk.f (+ genblkPosix / gensav / genbkbPosix /
    gensvb) holds after reading the geombc and restart files phio.write_geombc / write_restart produce -- the
    /conpar/ scalars and derived constants, x, the raw BC arrays, iper, every interior and boundary block, qold, acold,
    lstep.  phio.read_geombc / read_restart must return the same from the same files (the BC arrays after geniBC /
    genBC1, which tests/golden/f77_genbc1.npz pins separately)."""
    from make_golden_readnblk import build_case
    z = np.load(os.path.join(GOLD, "f77_readnblk.npz"))
    Z = lambda k: z["%s_%s" % (name, k)]  # noqa: E731
    (params, tables, parts, states), ibksz, lstep, rank = build_case(name)
    p = parts[rank]
    y, ac = states[rank]
    numpe = len(parts)
    d = str(tmp_path)
    phio.write_geombc(p, d)
    for r in sorted({0, rank}):                            # (rank 0 writes numstart.dat)
        phio.write_restart(d, r, numpe, lstep, *states[r])
    q = phio.read_geombc(d, rank, numpe, ibksz)
    sc = dict(zip(("numnp", "nshg", "numel", "numelb", "nen", "nelblk", "nelblb", "numpbc", "nflow", "ndof", "ndofBC", "ndiBCB",
                   "ndBCB", "nsymdf", "nenb", "lstep", "nlwork", "nshg0"), (int(v) for v in Z("scalars"))))
    assert (sc["numnp"], sc["nshg"], sc["numel"]) == (q.numnp, q.nshg, q.numel)
    assert sc["numelb"] == sum(b.shape[0] for b in q.mienb) and sc["nen"] == max(b.shape[1] for b in q.mien)
    assert (sc["nelblk"], sc["nelblb"]) == (q.nelblk, q.nelblb)            # after genblk / genbkb: blocks, not topologies
    assert (sc["nflow"], sc["ndof"], sc["ndofBC"], sc["ndiBCB"], sc["ndBCB"], sc["nsymdf"]) == (5, 5, 6, 2, 6, 15)
    if numpe == 1:
        assert (sc["lstep"], sc["nlwork"], sc["nshg0"]) == (lstep, 1, q.nshg)
    else:       # 'size of ilwork array' / 'ilwork', then ctypes: iother 0-based, segments as written
        assert (sc["lstep"], sc["nlwork"]) == (lstep, q.nlwork) and np.array_equal(Z("ilwork"), q.ilwork)
        assert int(q.ilwork[0]) == 2                      # the middle part talks to both neighbours
    assert np.array_equal(Z("x"), q.x)
    assert np.array_equal(Z("lcblk"), q.lcblk) and np.array_equal(Z("lcblkb"), q.lcblkb)
    for i in range(q.nelblk):
        assert np.array_equal(Z("mien_%d" % i), q.mien[i])
    for i in range(q.nelblb):
        assert np.array_equal(Z("mienb_%d" % i), q.mienb[i]) and np.array_equal(Z("mibcb_%d" % i), q.miBCB[i])
        assert np.array_equal(Z("mbcb_%d" % i), q.mBCB[i])
    # periodic masters: 0 in the file = the node itself
    raw = Z("iper")
    assert np.array_equal(np.where(raw == 0, np.arange(1, q.nshg + 1), raw), q.iper)
    # essential BCs: geniBC (genibc.f:17-19) and genBC (genbc.f:20-25) + genBC1 on the arrays the Fortran read
    nBC, iBCtmp, BCinp = Z("nBC"), Z("iBCtmp"), Z("BCinp")
    assert sc["numpbc"] == iBCtmp.size == BCinp.shape[0] and BCinp.shape[1] == 12
    iBC = np.zeros(q.nshg, dtype=np.int32)
    sel = nBC != 0
    iBC[sel] = iBCtmp[nBC[sel] - 1]
    assert np.array_equal(iBC, q.iBC)
    BCtmp = np.zeros((q.nshg, 12), order="F")
    BCtmp[sel] = BCinp[nBC[sel] - 1]
    assert np.array_equal(phio.genBC1(BCtmp, iBC), q.BC)
    # restart: the reference keeps the file's column order {p,u,v,w,T}; restar('in') permutes to {u,v,w,p,T}
    y2, ac2, lstep2 = phio.read_restart(d, rank, numpe, q.nshg)
    inv = [1, 2, 3, 0, 4]
    assert lstep2 == lstep and np.array_equal(Z("qold")[:, inv], y2) and np.array_equal(Z("acold")[:, inv], ac2)
    assert np.array_equal(y2, y) and not Z("uold").any()


def test_readnblk_fixture_reproduces_from_the_reference():
    if not os.path.isdir("/root/reference/phSolver/common"):
        pytest.skip("reference sources not present")
    from make_golden_readnblk import generate
    z = np.load(os.path.join(GOLD, "f77_readnblk.npz"))
    new = generate()
    assert set(new) == set(z.files)
    assert all(np.array_equal(new[k], z[k]) for k in new)
