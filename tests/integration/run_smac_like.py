"""Run the UNMODIFIED reference SMAC runner (offpolicy/runner/rnn/smac_runner.py, what scripts/train_smac_{qmix,vdn}.sh start) on a
synthetic environment with the 3m interface -- StarCraft II itself (pysc2 / the game binary) is not installable here.  The env has
SMAC's API and shapes (3 agents, obs 30, state 48, 9 actions with availability masks that change every step, episode limit 60,
early termination with a 'won' flag, per-agent done flags, one shared team reward) and seeded pseudo-dynamics that depend on the
chosen actions, so a wrong action anywhere changes the rest of the episode.  Same two engines as run_mpe.py; prints one JSON line.
Test infrastructure only.
"""
import json
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from run_mpe import install_shims, ROOT  # noqa: E402


def make_env_class():
    import numpy as np

    class Discrete(object):
        def __init__(self, n):
            self.n = n

    Discrete.__name__ = "Discrete"

    class SyntheticSMAC3m(object):
        """reset() -> (obs (N,O), share_obs (N,S), avail (N,A)); step(one-hot actions (N,A)) -> (obs, share_obs, rewards (N,1), dones (N,1),
        infos [per agent dict], avail) -- the StarCraft2Env contract used through ShareDummyVecEnv (envs/env_wrappers.py:442-468)."""
        N, O, S, A, LIMIT = 3, 30, 48, 9, 60

        def __init__(self, seed):
            from gym.spaces import Discrete as GymDiscrete
            self.num_agents = self.N
            self.observation_space = [[self.O] for _ in range(self.N)]
            self.share_observation_space = [[self.S] for _ in range(self.N)]
            self.action_space = [GymDiscrete(self.A) for _ in range(self.N)]
            self.rs = np.random.RandomState(seed)
            self.W = self.rs.randn(self.A, self.O).astype(np.float32) * 0.3      # how actions push the observation
            self.t = 0

        def seed(self, s):
            self.rs = np.random.RandomState(s)

        def _avail(self):
            av = (self.rs.rand(self.N, self.A) < 0.6).astype(np.float32)
            av[:, 0] = 1.0                                                      # the no-op is always available (like SMAC)
            return av

        def _pack(self):
            state = np.tanh(self.x.mean(0))
            share = np.concatenate([state, np.tanh(self.x[:, :6].reshape(-1))])[:self.S].astype(np.float32)
            return self.x.copy(), np.repeat(share[None], self.N, 0), self.avail.copy()

        def reset(self):
            self.t = 0
            self.x = self.rs.randn(self.N, self.O).astype(np.float32)
            self.avail = self._avail()
            self.health = 1.0
            return self._pack()

        def step(self, actions):
            a = np.argmax(np.asarray(actions), axis=-1)
            assert all(self.avail[i, a[i]] == 1.0 for i in range(self.N)), "an unavailable action was chosen"
            self.t += 1
            self.x = (0.9 * self.x + self.W[a] + 0.1 * self.rs.randn(self.N, self.O)).astype(np.float32)
            r = float(np.tanh(self.x[np.arange(self.N), a].sum()) + 0.1 * (a == 1).sum())
            self.health -= 0.02 * (1 + (a == 0).sum())
            dead = self.health <= 0.0
            done = dead or self.t >= self.LIMIT
            self.avail = self._avail()
            obs, share, avail = self._pack()
            info = {"won": bool(not dead and self.rs.rand() < 0.5)} if done else {}
            return obs, share, np.full((self.N, 1), r, np.float32), np.full((self.N, 1), done), [dict(info) for _ in range(self.N)], avail

        def close(self):
            pass

    return SyntheticSMAC3m


def main():
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--engine", default="b200", choices=["b200", "b200-gpu", "reference"])
    ap.add_argument("--algo", default="qmix")
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--seed", type=int, default=1)
    a, extra = ap.parse_known_args()
    rh = install_shims()
    import numpy as np
    import torch
    torch.set_num_threads(1)
    if a.engine == "reference":
        rh.import_reference()
        device = torch.device("cpu")
    else:
        sys.path.insert(0, os.path.join(ROOT, "off-policy_b200"))
        from offpolicy._b200 import capi
        if a.engine == "b200":
            sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
            from build_emu import build
            capi._install_for_tests(build())
        else:
            capi.lib()
        device = capi.device()
    from offpolicy.config import get_config
    from offpolicy.utils.util import get_cent_act_dim, get_dim_from_space
    from offpolicy.envs.env_wrappers import ShareDummyVecEnv
    from offpolicy.runner.rnn.smac_runner import SMACRunner
    import offpolicy.utils.rec_buffer as rb
    parser = get_config()
    parser.add_argument('--map_name', type=str, default='3m')                                    # train_smac.py:52-59
    parser.add_argument('--use_available_actions', action='store_false', default=True)
    parser.add_argument('--use_same_share_obs', action='store_false', default=True)
    parser.add_argument('--use_global_all_local_state', action='store_true', default=False)
    argv = ["--env_name", "StarCraft2", "--algorithm_name", a.algo, "--experiment_name", "b200", "--map_name", "3m", "--seed", str(a.seed),
            "--buffer_size", "64", "--lr", "5e-4", "--batch_size", "4", "--num_env_steps", str(a.steps), "--num_random_episodes", "2",
            "--log_interval", "100000", "--eval_interval", "10000000", "--save_interval", "10000000", "--gain", "1"] + list(extra)
    all_args = parser.parse_known_args(argv)[0]
    all_args.use_wandb = False
    torch.manual_seed(all_args.seed)
    np.random.seed(all_args.seed)
    Env = make_env_class()
    env = ShareDummyVecEnv([lambda: Env(all_args.seed)])
    eval_env = ShareDummyVecEnv([lambda: Env(all_args.seed * 50000)])
    policy_info = {'policy_0': {"cent_obs_dim": get_dim_from_space(env.share_observation_space[0]), "cent_act_dim": get_cent_act_dim(env.action_space),
                                "obs_space": env.observation_space[0], "share_obs_space": env.share_observation_space[0],
                                "act_space": env.action_space[0]}}
    from pathlib import Path
    config = {"args": all_args, "policy_info": policy_info, "policy_mapping_fn": lambda i: 'policy_0', "env": env, "eval_env": eval_env,
              "num_agents": 3, "device": device, "run_dir": Path(tempfile.mkdtemp()), "buffer_length": Env.LIMIT,
              "use_same_share_obs": all_args.use_same_share_obs, "use_available_actions": all_args.use_available_actions}
    stdout = sys.stdout
    sys.stdout = sys.stderr
    runner = SMACRunner(config=config)
    rewards, wins, infos = [], [], []
    collect = runner.collecter

    def recording_collect(*args, **kw):
        info = collect(*args, **kw)
        rewards.append(float(info["average_episode_rewards"]))
        wins.append(int(info.get("win_rate", -1)))
        return info
    runner.collecter = recording_collect
    name = "train_policy_on_batch"
    train = getattr(runner.trainer, name)

    def recording_train(*args, **kw):
        out = train(*args, **kw)
        infos.append({k: float(v) for k, v in out[0].items() if k != "update_actor"})
        return out
    setattr(runner.trainer, name, recording_train)
    total = 0
    while total < all_args.num_env_steps:
        total = runner.run()
    sys.stdout = stdout
    print(json.dumps(dict(engine=a.engine, algo=a.algo, buffer=rb.__file__, trainer=type(runner.trainer).__module__, env_steps=int(total),
                          train_steps=int(runner.total_train_steps), rewards=rewards, wins=wins, train=infos)))


if __name__ == "__main__":
    main()
