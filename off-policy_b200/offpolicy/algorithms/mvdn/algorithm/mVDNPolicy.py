"""Drop-in `M_VDNPolicy` (reference: offpolicy/algorithms/mvdn/algorithm/mVDNPolicy.py): identical to M_QMixPolicy."""
from offpolicy.algorithms.mqmix.algorithm.mQMixPolicy import M_QMixPolicy


class M_VDNPolicy(M_QMixPolicy):
    pass
