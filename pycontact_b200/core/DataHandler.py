"""Session export / import, file-compatible with the reference's ``core/DataHandler.py:4-28``.

A session is one pickle of ``{"contacts", "analyzer", "trajectory", "maps"}``:

* ``contacts``   -- ``analysis.finalAccumulatedContacts`` (list of AccumulatedContact)
* ``analyzer``   -- ``[psf, dcd, cutoff, hbondcutoff, hbondcutangle, sel1text, sel2text, contactResults]``
* ``trajectory`` -- ``analysis.getTrajectoryData()``
* ``maps``       -- ``[lastMap1, lastMap2]``

Sessions written by the reference name its classes (``PyContact.core.Biochemistry.AtomContact`` ...)
and are read here through a class map; sessions written here can name either package's classes, so
that an unmodified PyContact GUI opens them (``writeSessionToFile(..., reference_layout=True)``, the
default).  The array-backed result containers of the GPU path are expanded to the reference's
list-of-lists on export -- a session file is the one place where every contact becomes an object.
"""
from __future__ import annotations

import io
import pickle

from . import Biochemistry as _bio

# classes a session file may mention: reference module path -> local class
_REFERENCE_MODULES = ("PyContact.core.Biochemistry", "PyContact.core.ContactAnalyzer", "PyContact.core.multi_trajectory")
_SESSION_CLASSES = ("AtomContact", "HydrogenBond", "AccumulatedContact", "TempContactAccumulate")


class _SessionUnpickler(pickle.Unpickler):
    """Resolves the reference's class names to this package's attribute-compatible classes."""

    def find_class(self, module, name):
        if module in _REFERENCE_MODULES and hasattr(_bio, name):
            return getattr(_bio, name)
        return super().find_class(module, name)


def _dump_reference_layout(export, fh):
    """pickle.dump with the result classes renamed to the reference's module.

    Implemented as a plain dump followed by a rewrite of the GLOBAL opcodes: protocol 2 names
    classes as ``c<module>\\n<name>\\n``, and no other byte sequence of that shape can occur inside
    a session (strings are length-prefixed).  The object state written by ``__getstate__`` is already
    the reference's attribute dictionary."""
    raw = pickle.dumps(export, protocol=2)
    old = ("c%s\n" % _bio.__name__).encode()
    new = b"cPyContact.core.Biochemistry\n"
    out, pos = io.BytesIO(), 0
    while True:
        hit = raw.find(old, pos)
        if hit < 0:
            out.write(raw[pos:])
            break
        end = raw.index(b"\n", hit + len(old))
        name = raw[hit + len(old):end].decode()
        out.write(raw[pos:hit])
        out.write((new if name in _SESSION_CLASSES else old) + name.encode() + b"\n")
        pos = end + 1
    fh.write(out.getvalue())


def _expand(contactResults):
    """list[frame] of list[AtomContact]; the lazy containers of the GPU path build the objects here."""
    return [list(frame) for frame in contactResults]


class DataHandler:
    """Handles the import or export of a session (same static methods as the reference)."""

    @staticmethod
    def importSessionFromFile(fileName):
        """-> [contacts, arguments, trajArgs, maps, contactResults] (core/DataHandler.py:8-17)."""
        with open(fileName, "rb") as fh:
            importDict = _SessionUnpickler(fh).load()
        contacts = importDict["contacts"]
        arguments = importDict["analyzer"][0:-1]
        trajArgs = importDict["trajectory"]
        maps = importDict["maps"]
        contactResults = importDict["analyzer"][-1]
        return [contacts, arguments, trajArgs, maps, contactResults]

    @staticmethod
    def writeSessionToFile(fileName, analysis, reference_layout=True):
        """Saves the current session (core/DataHandler.py:20-28).  ``reference_layout`` names the
        reference's classes in the file so that PyContact itself can open it."""
        analyzerArgs = [analysis.psf, analysis.dcd, analysis.cutoff, analysis.hbondcutoff,
                        analysis.hbondcutangle, analysis.sel1text, analysis.sel2text,
                        _expand(analysis.contactResults)]
        trajArgs = analysis.getTrajectoryData()
        exportDict = {"contacts": list(analysis.finalAccumulatedContacts), "analyzer": analyzerArgs,
                      "trajectory": trajArgs, "maps": [analysis.lastMap1, analysis.lastMap2]}
        with open(fileName, "wb") as fh:
            if reference_layout:
                _dump_reference_layout(exportDict, fh)
            else:
                pickle.dump(exportDict, fh, protocol=2)
