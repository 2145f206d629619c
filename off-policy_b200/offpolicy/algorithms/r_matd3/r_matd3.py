"""Drop-in `R_MATD3` (reference: offpolicy/algorithms/r_matd3/r_matd3.py): R-MADDPG with the actor updated every 2nd critic update."""
from offpolicy.algorithms.r_maddpg.r_maddpg import R_MADDPG


class R_MATD3(R_MADDPG):
    actor_every = 2

    def __init__(self, args, num_agents, policies, policy_mapping_fn, **kwargs):
        kwargs["actor_update_interval"] = self.actor_every
        R_MADDPG.__init__(self, args, num_agents, policies, policy_mapping_fn, **kwargs)
