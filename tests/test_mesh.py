"""Integer data of the synthetic fixtures (bit-exact targets of SURVEY 8(a)
a27-a29): element blocking, connectivity orientation, ilwork pairing."""
import numpy as np

from phasta_b200 import make_box


def test_blocks_follow_genblk_rules():
    mp = make_box(5, 4, 3, ibksiz=64, bc="none", periodic_z=False)[0]
    lc = mp.lcblk
    numel = 5 * 4 * 3 * 6
    assert lc[0, 0] == 1 and lc[0, -1] == numel + 1
    npro = np.diff(lc[0])
    assert (npro[:-1] == 64).all() and 0 < npro[-1] <= 64      # genblkPosix.f:52-54
    assert (lc[2, :-1] == 1).all() and (lc[9, :-1] == 4).all() and (lc[4, :-1] == 4).all()
    assert sum(b.shape[0] for b in mp.mien) == numel
    assert all(b.flags.f_contiguous and b.dtype == np.int32 for b in mp.mien)


def test_tets_have_positive_jacobian_and_fill_the_box():
    mp = make_box(4, 3, 3, bc="none", periodic_z=False, perturb=0.15)[0]
    ien = mp.ien_all() - 1
    x = mp.x
    e = np.stack([x[ien[:, k]] - x[ien[:, 3]] for k in range(3)], axis=1)
    vol = np.linalg.det(e) / 6.0
    assert (vol > 0).all()
    assert abs(vol.sum() - 1.0 * 0.5 * 0.5) < 1e-12


def test_ilwork_tasks_pair_up():
    parts = make_box(8, 3, 3, nparts=4, bc="channel", max_seg=5)
    tasks = {}
    for mp in parts:
        il = mp.ilwork
        itk = 1
        for _ in range(il[0]):
            tag, iacc, other, nseg = il[itk:itk + 4]
            segs = il[itk + 4: itk + 4 + 2 * nseg].reshape(nseg, 2)
            nodes = np.concatenate([np.arange(b - 1, b - 1 + ln) for b, ln in segs])
            tasks[(mp.rank, other, tag, iacc)] = mp.gnode[nodes]
            itk += 4 + 2 * nseg
    assert tasks
    for (r, o, tag, iacc), g in tasks.items():
        partner = tasks[(o, r, tag, 1 - iacc)]
        assert np.array_equal(g, partner)            # same global nodes in the same order
        assert (iacc == 1) == (r < o)                 # lower rank is master


def test_periodic_masters_are_not_slaves():
    mp = make_box(3, 3, 3, bc="channel")[0]
    sl = np.nonzero(mp.iBC & (1 << 10))[0]
    assert sl.size == 4 * 4
    m = mp.iper[sl] - 1
    assert ((mp.iBC[m] & (1 << 10)) == 0).all()
    assert np.allclose(mp.x[sl, :2], mp.x[m, :2])
