"""PSF / DCD readers and the selection subset, on files written by the test itself (the reference's
fixture files live under /root/reference, which does not exist on the GPU box)."""
import struct

import numpy as np
import pytest

from pycontact_b200 import synth
from pycontact_b200.io.dcd import DCDReader
from pycontact_b200.io.psf import read_psf
from pycontact_b200.io.selection import SelectionError, select_atoms


def _write_psf(path, topo):
    with open(path, "w") as fh:
        fh.write("PSF\n\n       1 !NTITLE\n REMARKS synthetic\n\n")
        fh.write("%8d !NATOM\n" % topo.n_atoms)
        for i in range(topo.n_atoms):
            fh.write("%8d %-4s %-4d %-4s %-4s %-4s %10.6f %13.4f %11d\n" % (
                i + 1, topo.segids[i], topo.resids[i], topo.resnames[i], topo.names[i], topo.types[i], 0.0, 1.0, 0))
        fh.write("\n%8d !NBOND: bonds\n" % len(topo.bonds))
        for k in range(0, len(topo.bonds), 4):
            fh.write("".join("%8d%8d" % (a + 1, b + 1) for a, b in topo.bonds[k:k + 4]) + "\n")
        fh.write("\n")


def _write_dcd(path, coords, with_cell=True):
    nf, n, _ = coords.shape
    with open(path, "wb") as fh:
        icntrl = np.zeros(20, "<i4")
        icntrl[0] = nf; icntrl[1] = 0; icntrl[2] = 1; icntrl[10] = 1 if with_cell else 0; icntrl[19] = 24
        fh.write(struct.pack("<i", 84) + b"CORD" + icntrl.tobytes() + struct.pack("<i", 84))
        title = b"REMARKS synthetic".ljust(80)
        fh.write(struct.pack("<i", 84) + struct.pack("<i", 1) + title + struct.pack("<i", 84))
        fh.write(struct.pack("<iii", 4, n, 4))
        for f in range(nf):
            if with_cell:
                fh.write(struct.pack("<i", 48) + np.zeros(6, "<f8").tobytes() + struct.pack("<i", 48))
            for ax in range(3):
                fh.write(struct.pack("<i", 4 * n) + coords[f, :, ax].astype("<f4").tobytes() + struct.pack("<i", 4 * n))


@pytest.mark.parametrize("with_cell", [True, False])
def test_psf_dcd_round_trip(tmp_path, with_cell):
    s = synth.make_system("tiny")
    coords = synth.frames_host(s, range(4))
    psf, dcd = str(tmp_path / "t.psf"), str(tmp_path / "t.dcd")
    _write_psf(psf, s.topology)
    _write_dcd(dcd, coords, with_cell)
    t = read_psf(psf)
    assert t.n_atoms == s.topology.n_atoms and t.names == s.topology.names and t.types == s.topology.types
    assert t.segids == s.topology.segids and np.array_equal(t.resids, s.topology.resids)
    assert np.array_equal(t.bonds, s.topology.bonds)
    r = DCDReader(dcd)
    assert len(r) == 4 and r.n_atoms == t.n_atoms
    assert np.array_equal(r.frame(2), coords[2])
    sel = select_atoms(t, s.sel1text)
    assert np.array_equal(sel, s.sel1)
    buf = np.empty((2, len(sel), 3), np.float32)
    r.read_into(buf, 1, 2, sel)
    assert np.array_equal(buf, coords[1:3][:, sel])
    # batch reads: all atoms, many short runs of indices (one gather per run), scattered indices (one fancy gather)
    allbuf = np.empty((4, t.n_atoms, 3), np.float32)
    assert np.array_equal(r.read_into(allbuf, 0, 4), coords)
    rng = np.random.default_rng(3)
    for pick in (np.sort(rng.choice(t.n_atoms, 40, replace=False)), np.sort(rng.choice(t.n_atoms, 400, replace=False)),
                 np.r_[5:9, 100:180, 300:301]):
        out = np.empty((3, len(pick), 3), np.float32)
        assert np.array_equal(r.read_into(out, 1, 3, pick), coords[1:4][:, pick])
    with pytest.raises(IndexError):
        r.read_into(buf, 3, 2, sel)


def test_selection_language(tmp_path):
    s = synth.make_system("tiny")
    t = s.topology
    assert np.array_equal(select_atoms(t, "segid PA* or segid PB*"), np.concatenate([s.sel1, s.sel2]))
    assert len(select_atoms(t, "name OH2")) == 120
    assert np.array_equal(select_atoms(t, "protein and backbone"), s.backbone)
    assert len(select_atoms(t, "not protein")) == 360
    assert len(select_atoms(t, "segid PA0 and resid 1:3 and name CA")) == 3
    assert np.array_equal(select_atoms(t, "index 0 5 7"), [0, 5, 7])
    assert np.array_equal(select_atoms(t, "bynum 1"), [0])
    with pytest.raises(SelectionError):
        select_atoms(t, "segid")
    with pytest.raises(NotImplementedError):
        select_atoms(t, "around 5 segid PA0")


# ---- XTC (csrc/pc_xtc.h through the C ABI; host code, no GPU needed) ------------------------------------------------
XTC_REF = "/root/reference/PyContact/exampleData/md_noPBC.xtc"          # build container only


class _Bits:
    """MSB-first bit writer (the order xdrfile's receivebits reads)."""

    def __init__(self):
        self.acc, self.n, self.out = 0, 0, bytearray()

    def put(self, value, nbits):
        self.acc = (self.acc << nbits) | (value & ((1 << nbits) - 1))
        self.n += nbits
        while self.n >= 8:
            self.n -= 8
            self.out.append((self.acc >> self.n) & 0xff)
        self.acc &= (1 << self.n) - 1

    def bytes(self):
        return bytes(self.out) + (bytes([(self.acc << (8 - self.n)) & 0xff]) if self.n else b"")


def _write_xtc(path, frames_nm, precision=1000.0, cut=0):
    """XTC writer for tests: every atom coded with the frame's full-range bits and flag bit 0 (no runs) -- a valid
    stream for any xdrfile-compatible reader; <= 9 atoms are stored as plain floats like xdrfile does."""
    with open(path, "wb") as fh:
        for f, x in enumerate(np.asarray(frames_nm, np.float32)):
            n = x.shape[0]
            fh.write(struct.pack(">iiif", 1995, n, 10 * f, 2.5 * f) + np.diag([5.0, 6.0, 7.0]).astype(">f4").tobytes())
            fh.write(struct.pack(">i", n))
            if n <= 9:
                fh.write(x.astype(">f4").tobytes())
                continue
            q = np.rint(x.astype(np.float64) * precision).astype(np.int64)
            lo, hi = q.min(0), q.max(0)
            size = (hi - lo + 1).tolist()
            nbits = (size[0] * size[1] * size[2]).bit_length()                 # == xdrfile's sizeofints for these sizes
            bw = _Bits()
            for a in (q - lo).tolist():
                total = (a[0] * size[1] + a[1]) * size[2] + a[2]
                left = nbits
                while left > 8:                                                # low bytes first, whole bytes
                    bw.put(total & 0xff, 8)
                    total >>= 8
                    left -= 8
                bw.put(total, left)
                bw.put(0, 1)                                                   # flag: no run follows
            data = bw.bytes()
            fh.write(struct.pack(">f3i3iii", precision, *lo.tolist(), *hi.tolist(), 9, len(data)))
            fh.write(data + b"\0" * (-len(data) % 4))
        if cut:
            fh.truncate(fh.tell() - cut)


def test_xtc_reader_on_written_files(tmp_path):
    from pycontact_b200.io.xtc import XTCReader, open_trajectory
    rng = np.random.default_rng(5)
    x = rng.uniform(-2.0, 9.0, (4, 300, 3)).astype(np.float32)                 # nm
    p = str(tmp_path / "t.xtc")
    _write_xtc(p, x)
    rd = open_trajectory(p)
    assert isinstance(rd, XTCReader) and rd.n_atoms == 300 and len(rd) == 4
    want = (np.rint(x.astype(np.float64) * 1000.0).astype(np.int32).astype(np.float32) * np.float32(1.0 / np.float32(1000.0))) * np.float32(10.0)
    got = np.empty((4, 300, 3), np.float32)
    rd.read_into(got, 0, 4)
    assert np.array_equal(got, want)                                           # bit-exact: int * (1 / precision) * 10 in float32
    idx = np.array([7, 0, 299, 13], np.int32)
    assert np.array_equal(rd.frame(2, idx), want[2, idx])
    step, time, box = rd.header(3)
    assert step == 30 and time == 7.5 and np.allclose(np.diag(box), [50.0, 60.0, 70.0])
    with pytest.raises(IndexError):
        rd.frame(4)
    # <= 9 atoms: plain floats
    p2 = str(tmp_path / "s.xtc")
    _write_xtc(p2, x[:2, :5])
    rd2 = XTCReader(p2)
    assert rd2.n_atoms == 5 and len(rd2) == 2 and np.array_equal(rd2.frame(1), x[1, :5] * np.float32(10.0))
    # a file cut inside its last frame (interrupted write) ends one frame earlier; garbage is refused
    p3 = str(tmp_path / "c.xtc")
    _write_xtc(p3, x, cut=40)
    assert len(XTCReader(p3)) == 3
    bad = tmp_path / "bad.xtc"
    bad.write_bytes(b"not an xtc file" * 10)
    from pycontact_b200._lib import PcError
    with pytest.raises(PcError):
        XTCReader(str(bad))


@pytest.mark.skipif(not __import__("os").path.exists(XTC_REF), reason="reference fixture only exists in the build container")
def test_xtc_reader_on_the_reference_fixture():
    """exampleData/md_noPBC.xtc (tests/test_basic.py:28-30 opens it): written by GROMACS with runs and the water swap.
    No MDAnalysis here to compare with, so the decode is pinned on physics: the file's 20 437 rigid SPC waters must come
    back with O-H = 1.000 A and H-H = 1.633 A in every frame, to the file's 0.01 A quantisation."""
    from pycontact_b200.io.xtc import XTCReader
    rd = XTCReader(XTC_REF)
    assert (rd.n_atoms, len(rd)) == (65765, 10)
    steps = [rd.header(f)[0] for f in range(10)]
    assert steps == [500000 * f for f in range(10)] and rd.header(9)[1] == 9000.0
    assert abs(rd.header(0)[2][0, 0] - 86.9428) < 1e-3
    w = np.arange(4330, 65641, 3)
    for f in (0, 4, 9):
        x = rd.frame(f).astype(np.float64)
        oh = np.concatenate([np.linalg.norm(x[w + 1] - x[w], axis=1), np.linalg.norm(x[w + 2] - x[w], axis=1)])
        hh = np.linalg.norm(x[w + 2] - x[w + 1], axis=1)
        assert np.abs(oh - 1.0).max() < 0.02 and np.abs(hh - 1.633).max() < 0.02
        assert abs(oh.mean() - 1.0) < 1e-3 and abs(hh.mean() - 1.633) < 1e-3
        prot = np.linalg.norm(np.diff(x[:4330], axis=0), axis=1)
        assert (prot > 10.0).sum() <= 2 and 1.0 < np.median(prot) < 2.6        # two folded chains, atom after atom
    idx = np.arange(0, rd.n_atoms, 11, dtype=np.int32)
    assert np.array_equal(rd.read_into(np.empty((2, idx.size, 3), np.float32), 3, 2, idx)[1], rd.frame(4)[idx])
