"""CPU oracle for the RUV-III path (TEST INFRASTRUCTURE ONLY; see ruv_oracle.py header).

PARITY UNPINNED against a run of the R reference (R is not installable here); pinned only on the
invariants the reference's own tests and docs state.  Never imported by `scmerge_b200/`."""
from . import ruv_oracle  # noqa: F401
