"""B200-native drop-in for `offpolicy` (marlbenchmark/off-policy): put `off-policy_b200/` first on PYTHONPATH.
Unlike the reference's package __init__ (offpolicy/__init__.py:1) nothing heavy is imported here."""
from offpolicy._b200.refpath import extend as _extend

_extend(__path__)
