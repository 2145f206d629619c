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


if __name__ == "__main__":
    main()
