"""The reference's own host objects for the binding tests: see oracle/ref_hosts.py (test infrastructure)."""
from ref_hosts import *  # noqa: F401,F403
from ref_hosts import _quiet, _trainer_fields  # noqa: F401
