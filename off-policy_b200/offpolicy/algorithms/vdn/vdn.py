"""Drop-in `VDN` trainer (reference: offpolicy/algorithms/vdn/vdn.py): the QMIX learner with the parameter-free
sum mixer (`k_vdn_mix`), i.e. Q_tot = sum over agents of the taken-action Q values."""
from offpolicy.algorithms.qmix.qmix import QMix


class VDN(QMix):
    mixer_kind = "sum"

    def __init__(self, args, num_agents, policies, policy_mapping_fn, **kwargs):
        kwargs["vdn"] = True
        QMix.__init__(self, args, num_agents, policies, policy_mapping_fn, **kwargs)
