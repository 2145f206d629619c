"""Drop-in `VDNPolicy` (reference: offpolicy/algorithms/vdn/algorithm/VDNPolicy.py): identical to QMixPolicy."""
from offpolicy.algorithms.qmix.algorithm.QMixPolicy import QMixPolicy


class VDNPolicy(QMixPolicy):
    pass
