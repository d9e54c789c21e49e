"""Scripting front end, same surface as the reference's ``core/Scripting.py:1-41``::

    job = PyContactJob(psf, dcd, "title", JobConfig(5.0, 2.5, 120, [0,0,1,1,0], [0,0,1,1,0], "segid RN11", "segid UBQ"))
    job.runJob(1)                   # ncores = number of GPUs here
    job.writeSessionToFile()
"""
from .ContactAnalyzer import *          # noqa: F401,F403  (the reference re-exports the analyzer module)
from .ContactAnalyzer import Analyzer
from .DataHandler import DataHandler


class JobConfig:
    """Configuration/Settings for a contact analysis job (core/Scripting.py:6-15)."""

    def __init__(self, cutoff, hbondcutoff, hbondcutangle, map1, map2, sel1, sel2):
        self.cutoff = cutoff
        self.hbondcutoff = hbondcutoff
        self.hbondcutangle = hbondcutangle
        self.map1 = map1
        self.map2 = map2
        self.sel1 = sel1
        self.sel2 = sel2


class PyContactJob:
    """Runs one analysis from input files and a JobConfig (core/Scripting.py:18-41)."""

    def __init__(self, topo, traj, name, configuration):
        self.topo = topo
        self.traj = traj
        self.name = name
        self.configuration = configuration
        self.analyzer = None

    def runJob(self, ncores=1):
        """Frame scan + accumulation; ``ncores`` is passed on as the number of GPUs to use."""
        print("Running job: " + self.name + " on " + str(ncores) + " cores.")
        c = self.configuration
        self.analyzer = Analyzer(self.topo, self.traj, c.cutoff, c.hbondcutoff, c.hbondcutangle, c.sel1, c.sel2)
        self.analyzer.runFrameScan(ncores)
        self.analyzer.runContactAnalysis(c.map1, c.map2, ncores)

    def writeSessionToFile(self, fname=""):
        """Writes the session to ``fname`` or ``<name>.session`` (a file PyContact's GUI can open)."""
        DataHandler.writeSessionToFile(fname if fname != "" else self.name + ".session", self.analyzer)
        print("Wrote session to file")
