"""Golden SASA / find_within vectors from the UNMODIFIED reference (oracle/_ref/libref_gridsearch.so, compiled
from /root/reference/PyContact/cy_modules/src/gridsearch.C by oracle/Makefile) on the rpn11_ubq fixture.

    python tests/golden/make_sasa_golden.py      # run in the build container (needs /root/reference once)

The calls are the ones the reference's SASA widget makes (gui/SasaWidgets.py:156-215): probe radius 1.4,
50 random points (pointstyle 1), pairdist 2 * (2.0 + 1.4), radii vdwRadius(name[0]); selection `segid UBQ`,
once unrestricted and once restricted to the UBQ atoms whose residue has an atom within 5 A of RN11 in frame 0
(the widget's example restriction `segid UBQ and same residue as around 5.0 (segid RN11)`).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import sasa_oracle as so   # noqa: E402

VDW = {"H": 1.0, "C": 1.5, "N": 1.399999976158142, "O": 1.2999999523162842, "F": 1.47, "P": 1.8, "S": 1.899999976158142}


def main():
    z = np.load(os.path.join(HERE, "rpn11_ubq_input.npz"))
    coords, names, segids, resids = z["coords"], z["names"], z["segids"], z["resids"]
    ubq = np.nonzero(segids == "UBQ")[0]
    rn11 = np.nonzero(segids == "RN11")[0]
    rad = np.array([VDW.get(str(n)[0], 1.5) for n in names[ubq]], np.float32)
    pairdist = 2.0 * (2.0 + 1.4)
    # restriction from frame 0 through the reference's own find_within
    n = len(names)
    flg = np.zeros(n, np.int32); flg[ubq] = 1
    oth = np.zeros(n, np.int32); oth[rn11] = 1
    within0 = so.ref_find_within(coords[0], flg, oth, 5.0)
    res_hit = set(resids[np.nonzero(within0)[0]].tolist())
    rl = np.array([1 if int(resids[i]) in res_hit else 0 for i in ubq], np.int32)
    within_all = np.stack([so.ref_find_within(coords[f], flg, oth, 5.0) for f in range(len(coords))])
    free = np.array([so.ref_sasa(coords[f][ubq], rad, pairdist, 50, 1.4, 1, 0, np.array([0], np.int32))
                     for f in range(len(coords))], np.float32)
    restr = np.array([so.ref_sasa(coords[f][ubq], rad, pairdist, 50, 1.4, 1, 1, rl) for f in range(len(coords))], np.float32)
    spiral = np.array([so.ref_sasa(coords[f][ubq], rad, pairdist, 64, 1.4, 0, 0, np.array([0], np.int32))
                       for f in range(0, len(coords), 10)], np.float32)
    out = os.path.join(HERE, "rpn11_ubq_sasa_golden.npz")
    np.savez_compressed(out, sasa_free=free, sasa_restricted=restr, sasa_spiral64_every10=spiral, restricted_list=rl,
                        within5_ubq_of_rn11=np.packbits(within_all.astype(np.uint8), axis=1),
                        within5_counts=within_all.sum(axis=1).astype(np.int32))
    print(out, free[:3], restr[:3], spiral[:2], int(rl.sum()), within_all.sum(axis=1)[:5])


if __name__ == "__main__":
    main()
