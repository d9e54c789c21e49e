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
