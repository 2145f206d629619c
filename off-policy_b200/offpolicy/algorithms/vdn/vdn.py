"""Drop-in `VDN` trainer (reference: offpolicy/algorithms/vdn/vdn.py:4-8): QMix with the sum mixer."""
from offpolicy.algorithms.qmix.qmix import QMix


class VDN(QMix):
    def __init__(self, args, num_agents, policies, policy_mapping_fn, device=None, episode_length=None):
        super(VDN, self).__init__(args, num_agents, policies, policy_mapping_fn, device=device, episode_length=episode_length, vdn=True)
