"""Drop-in `R_MATD3Policy` (reference: offpolicy/algorithms/r_matd3/algorithm/rMATD3Policy.py): twin Q heads + target smoothing noise."""
from offpolicy.algorithms.r_maddpg.algorithm.rMADDPGPolicy import R_MADDPGPolicy


class R_MATD3Policy(R_MADDPGPolicy):
    def __init__(self, config, policy_config, train=True):
        noise = config["args"].target_action_noise_std
        R_MADDPGPolicy.__init__(self, config, policy_config, target_noise=noise, td3=True, train=train)
