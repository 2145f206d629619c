"""Worker for test_dp_gloo.py: one data-parallel rank (gloo, CPU emulated kernels)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "off-policy_b200"), os.path.join(ROOT, "tests", "emu")):
    sys.path.insert(0, p)


def main():
    out_dir = sys.argv[1]
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from build_emu import LIB
    from offpolicy._b200 import capi
    capi._install_for_tests(LIB)
    import qmix_checks as qc
    from helpers import load_golden, oracle_from_golden, golden_batch, sub
    if len(sys.argv) > 2 and sys.argv[2] == "mlp":
        return main_mlp(out_dir, rank, world)
    g = load_golden("qmix_small")
    L, cfg, B, T, steps = oracle_from_golden(g)
    Bl = B // world
    args, pol, tr = qc.build_trainer(cfg, Bl, T, debug=False)      # product configuration (k_mid), like bench.py --gpus N
    assert tr.world_size == world
    qc.load_state(pol, tr, sub(g, "init.agent."), sub(g, "init.mixer."), sub(g, "init.tgt_agent."), sub(g, "init.tgt_mixer."))
    full = golden_batch(g, 0)
    sl = slice(rank * Bl, (rank + 1) * Bl)
    shard = tuple(x[..., sl, :] if x.ndim == 4 else x[:, sl] for x in full[:7]) + (None, None)
    info, _, _ = tr.train_policy_on_batch(qc.ref_tuple(shard))
    tr.soft_target_updates()
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), loss=float(info["loss"]), grad_norm=float(info["grad_norm"]),
             Q_tot=float(info["Q_tot"]), theta=tr.theta.numpy(), theta_tgt=tr.theta_tgt.numpy())
    dist.destroy_process_group()


def main_mlp(out_dir, rank, world):
    """M_QMix (transition-level path): each rank trains on its slice of the golden transition batch."""
    import mqmix_checks as mc
    from helpers import load_golden, golden_cfg, sub
    g = load_golden("mqmix_small")
    cfg, B, T, steps = golden_cfg(g)
    Bl = B // world
    args, pol, tr = mc.build(cfg, Bl, debug=False)
    assert tr.world_size == world
    pol.q_network.load_state_dict(sub(g, "init.agent."))
    tr.target_q_network.load_state_dict(sub(g, "init.tgt_agent."))
    tr.mixer.load_state_dict(sub(g, "init.mixer."))
    tr.target_mixer.load_state_dict(sub(g, "init.tgt_mixer."))
    full = mc.golden_transitions(g, 0)
    sl = slice(rank * Bl, (rank + 1) * Bl)
    cut = lambda d: {"policy_0": None if d["policy_0"] is None else (d["policy_0"][:, sl] if d["policy_0"].ndim == 3 else d["policy_0"][sl])}
    shard = tuple(cut(d) for d in full[:11]) + (None, None)
    info, _, _ = tr.train_policy_on_batch(shard, True)
    tr.soft_target_updates()
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), loss=float(info["loss"]), grad_norm=float(info["grad_norm"]),
             Q_tot=float(info["Q_tot"]), theta=tr.theta.numpy(), theta_tgt=tr.theta_tgt.numpy())
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
