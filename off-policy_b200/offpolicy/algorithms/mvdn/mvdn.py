"""Drop-in `M_VDN` trainer (reference: offpolicy/algorithms/mvdn/mvdn.py): M_QMix with the parameter-free sum mixer.  (The
reference's own M_VDNMixer.forward takes one argument but is called with two; the intent is built.)"""
from offpolicy.algorithms.mqmix.mqmix import M_QMix


class M_VDN(M_QMix):
    def __init__(self, args, num_agents, policies, policy_mapping_fn, device=None):
        M_QMix.__init__(self, args, num_agents, policies, policy_mapping_fn, device=device, vdn=True)
