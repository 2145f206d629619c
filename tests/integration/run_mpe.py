"""Run the UNMODIFIED reference runner (offpolicy/runner/rnn/mpe_runner.py, MPE simple_spread = scripts/train_mpe_<algo>.sh) for a
few episodes, either on the reference's own classes (`--engine reference`) or with this repository's drop-in package shadowing
`offpolicy.utils.rec_buffer` / `offpolicy.algorithms.*` (`--engine b200`, CPU fiber-emulated kernels: a `-m "not gpu"` test; or
`--engine b200-gpu` on a B200).  Prints one JSON line: per-episode rewards and per-update train_info.

Test infrastructure only.  Needs the reference checkout (skipped without it).  Packages the build container lacks and the hot path
never touches (gym, wandb, tensorboardX, seaborn, imp) are replaced by inert shims; no reference file is modified.
"""
import argparse
import json
import os
import sys
import tempfile
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def install_shims():
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import ref_harness as rh
    rh._install_gym_shim()
    import gym

    class Env(object):
        pass
    gym.Env = Env
    reg = types.ModuleType("gym.envs.registration")
    reg.EnvSpec = type("EnvSpec", (object,), {"__init__": lambda self, *a, **k: None})
    envs = types.ModuleType("gym.envs")
    envs.registration = reg
    sys.modules["gym.envs"], sys.modules["gym.envs.registration"], gym.envs = envs, reg, envs
    wandb = types.ModuleType("wandb")
    wandb.log = lambda *a, **k: None
    sys.modules["wandb"] = wandb
    tbx = types.ModuleType("tensorboardX")
    tbx.SummaryWriter = type("SummaryWriter", (object,), {"__init__": lambda self, *a, **k: None, "add_scalars": lambda self, *a, **k: None,
                                                          "export_scalars_to_json": lambda self, *a, **k: None, "close": lambda self: None})
    sys.modules["tensorboardX"] = tbx
    sns = types.ModuleType("seaborn")
    sns.color_palette = lambda *a, **k: [(0.3, 0.3, 0.3)] * 16
    sys.modules["seaborn"] = sns
    imp = types.ModuleType("imp")

    def load_source(name, pathname):
        import importlib.util
        spec = importlib.util.spec_from_file_location(name or "scenario", pathname)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod
    imp.load_source = load_source
    sys.modules["imp"] = imp
    return rh


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--engine", default="b200", choices=["b200", "b200-gpu", "reference"])
    ap.add_argument("--algo", default="qmix")
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--runner", default="rnn", choices=["rnn", "mlp"], help="runner/rnn/mpe_runner.py (recurrent algorithms) or runner/mlp/mpe_runner.py")
    ap.add_argument("--scenario", default="simple_spread")
    ap.add_argument("--agents", type=int, default=3)
    a, extra = ap.parse_known_args()          # unknown flags go to the reference's own parser (config.py)
    a.extra = extra
    rh = install_shims()
    import numpy as np
    import torch
    torch.set_num_threads(1)
    if a.engine == "reference":
        rh.import_reference()                              # `offpolicy` = the reference tree only
        device = torch.device("cpu")
    else:
        sys.path.insert(0, os.path.join(ROOT, "off-policy_b200"))      # the drop-in package shadows the hot-path modules
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
    from offpolicy.envs.mpe.MPE_Env import MPEEnv
    from offpolicy.envs.env_wrappers import DummyVecEnv
    if a.runner == "mlp":
        from offpolicy.runner.mlp.mpe_runner import MPERunner
        import offpolicy.utils.mlp_buffer as rb
    else:
        from offpolicy.runner.rnn.mpe_runner import MPERunner
        import offpolicy.utils.rec_buffer as rb
    parser = get_config()
    parser.add_argument('--scenario_name', type=str, default='simple_spread')          # train_mpe.py:49-57
    parser.add_argument("--num_landmarks", type=int, default=3)
    parser.add_argument('--num_agents', type=int, default=3)
    parser.add_argument('--use_same_share_obs', action='store_false', default=True)
    argv = ["--env_name", "MPE", "--algorithm_name", a.algo, "--experiment_name", "b200", "--scenario_name", a.scenario, "--num_agents", str(a.agents),
            "--num_landmarks", "3", "--seed", str(a.seed), "--episode_length", "25", "--tau", "0.005", "--lr", "7e-4",
            "--num_env_steps", str(a.steps), "--batch_size", "4" if a.runner == "rnn" else "16", "--buffer_size", "64" if a.runner == "rnn" else "512",
            "--num_random_episodes", "2", "--train_interval", "25",
            "--log_interval", "100000", "--eval_interval", "10000000", "--save_interval", "10000000"] + list(a.extra)
    all_args = parser.parse_known_args(argv)[0]
    all_args.use_wandb = False
    torch.manual_seed(all_args.seed)
    np.random.seed(all_args.seed)

    def init_env():
        env = MPEEnv(all_args)
        env.seed(all_args.seed)
        return env
    env = DummyVecEnv([init_env])
    if all_args.share_policy:
        policy_info = {'policy_0': {"cent_obs_dim": get_dim_from_space(env.share_observation_space[0]), "cent_act_dim": get_cent_act_dim(env.action_space),
                                    "obs_space": env.observation_space[0], "share_obs_space": env.share_observation_space[0],
                                    "act_space": env.action_space[0]}}
        mapping = lambda i: 'policy_0'
    else:       # `--share_policy` is a store_false flag (config.py:61): one policy per agent, train/train_mpe.py:139-150
        policy_info = {'policy_' + str(i): {"cent_obs_dim": get_dim_from_space(env.share_observation_space[i]), "cent_act_dim": get_cent_act_dim(env.action_space),
                                            "obs_space": env.observation_space[i], "share_obs_space": env.share_observation_space[i],
                                            "act_space": env.action_space[i]} for i in range(a.agents)}
        mapping = lambda i: 'policy_' + str(i)
    from pathlib import Path
    config = {"args": all_args, "policy_info": policy_info, "policy_mapping_fn": mapping, "env": env, "eval_env": env,
              "num_agents": a.agents, "device": device, "use_same_share_obs": all_args.use_same_share_obs, "run_dir": Path(tempfile.mkdtemp())}
    stdout = sys.stdout
    sys.stdout = sys.stderr                                # the reference prints progress
    runner = MPERunner(config=config)
    rewards, infos = [], []
    collect = runner.collecter

    def recording_collect(*args, **kw):
        info = collect(*args, **kw)
        rewards.append(float(info["average_episode_rewards"]))
        return info
    runner.collecter = recording_collect
    q_learning = a.algo in ("qmix", "vdn", "mqmix", "mvdn")
    name = "train_policy_on_batch" if q_learning else "shared_train_policy_on_batch"
    train = getattr(runner.trainer, name)

    def recording_train(*args, **kw):
        out = train(*args, **kw)
        infos.append({k: float(v) for k, v in out[0].items() if k != "update_actor"})
        return out
    setattr(runner.trainer, name, recording_train)
    if not q_learning:
        runner.train = runner.batch_train if hasattr(runner, "batch_train") and runner.train.__name__ == "batch_train" else runner.train
    total = 0
    while total < all_args.num_env_steps:
        total = runner.run()
    sys.stdout = stdout
    print(json.dumps(dict(engine=a.engine, algo=a.algo, buffer=rb.__file__, trainer=type(runner.trainer).__module__, env_steps=int(total),
                          train_steps=int(runner.total_train_steps), rewards=rewards, train=infos)))


if __name__ == "__main__":
    main()
