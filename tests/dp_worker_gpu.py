"""Worker for test_gpu_dp.py: one data-parallel rank on its own GPU (NCCL for the rendezvous; gradient exchange over peer memory
unless MARL_B200_P2P=0 selects the NCCL all-reduce)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "off-policy_b200")):
    sys.path.insert(0, p)


def main():
    out_dir = sys.argv[1]
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from offpolicy._b200 import capi
    capi.lib()
    import qmix_checks as qc
    from helpers import load_golden, oracle_from_golden, golden_batch, sub
    g = load_golden("qmix_small")
    L, cfg, B, T, steps = oracle_from_golden(g)
    Bl = B // world
    args, pol, tr = qc.build_trainer(cfg, Bl, T, debug=False)      # product configuration (k_mid), like bench.py --gpus N
    assert tr.world_size == world
    qc.load_state(pol, tr, sub(g, "init.agent."), sub(g, "init.mixer."), sub(g, "init.tgt_agent."), sub(g, "init.tgt_mixer."))
    res = dict(p2p=int(tr._p2p))
    for s in range(steps):
        full = golden_batch(g, s)
        sl = slice(rank * Bl, (rank + 1) * Bl)
        shard = tuple(x[..., sl, :] if x.ndim == 4 else x[:, sl] for x in full[:7]) + (None, None)
        info, _, _ = tr.train_policy_on_batch(qc.ref_tuple(shard))
        tr.soft_target_updates()
        for k in ("loss", "grad_norm", "Q_tot"):
            res["s%d.%s" % (s, k)] = float(info[k])
        res["s%d.timeout" % s] = float(tr._info[7])
    torch.cuda.synchronize()
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), theta=tr.theta.cpu().numpy(), theta_tgt=tr.theta_tgt.cpu().numpy(), **res)
    dist.barrier()
    os._exit(0)


if __name__ == "__main__":
    main()
