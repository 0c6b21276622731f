"""Target names and task names (``deeprank2/domain/targetstorage.py``)."""
VALUES = "target_values"
BINARY = "binary"
CAPRI = "capri_class"
IRMSD = "irmsd"
LRMSD = "lrmsd"
FNAT = "fnat"
DOCKQ = "dockq"
REGRESS = "regress"
CLASSIF = "classif"

# default task of the built-in targets (dataset.py:153-163)
DEFAULT_TASK = {IRMSD: REGRESS, LRMSD: REGRESS, FNAT: REGRESS, DOCKQ: REGRESS, BINARY: CLASSIF, CAPRI: CLASSIF}
