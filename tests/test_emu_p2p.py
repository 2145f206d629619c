"""One-shot all-reduce over peer memory (csrc/p2p.cu), kernel logic on the CPU emulator: two data-parallel "ranks" driven from
one process, their symmetric blocks being plain host tensors.  Each rank trains on half of the golden batch; publish + reduce must
give both ranks the same sums, and the step must reproduce the reference's single-batch step (like tests/test_dp_gloo.py does
with an NCCL-style all-reduce)."""
import ctypes as C

import numpy as np
import torch

import qmix_checks as qc
from helpers import load_golden, oracle_from_golden, golden_batch, sub


import pytest


@pytest.mark.parametrize("debug", [True, False])      # False = the product configuration (k_mid), what bench.py --gpus N runs
def test_two_rank_peer_memory_exchange_equals_single_batch_reference(emu_engine, debug):
    capi = emu_engine
    lib = capi.lib()
    g = load_golden("qmix_small")
    L, cfg, B, T, steps = oracle_from_golden(g)
    world, Bl = 2, B // 2
    ranks = []
    for r in range(world):
        args, pol, tr = qc.build_trainer(cfg, Bl, T, debug=debug, dp_world_size=world)
        assert tr.world_size == world and not tr._p2p
        qc.load_state(pol, tr, sub(g, "init.agent."), sub(g, "init.mixer."), sub(g, "init.tgt_agent."), sub(g, "init.tgt_mixer."))
        ranks.append((pol, tr))
    n = int(lib.mx_qmix_p2p_block_bytes(ranks[0][1].handle)) // 4
    blocks = [torch.zeros(n, dtype=torch.float32) for _ in range(world)]
    for r, (pol, tr) in enumerate(ranks):
        tr.attach_peer_blocks(r, [b.data_ptr() for b in blocks], keep=blocks)
    for s in range(2):                      # two steps: both slot parities, flags advance 1 -> 2
        full = golden_batch(g, s)
        structs = []
        for r, (pol, tr) in enumerate(ranks):
            sl = slice(r * Bl, (r + 1) * Bl)
            shard = tuple(x[..., sl, :] if x.ndim == 4 else x[:, sl] for x in full[:7]) + (None, None)
            b = tr._device_batch(qc.ref_tuple(shard))
            capi.check(lib.mx_qmix_backward_only(tr.handle, C.byref(b), None))
        for pol, tr in ranks:               # every rank publishes before any rank reduces (the emulator cannot spin-wait)
            capi.check(lib.mx_qmix_p2p_publish(tr.handle, None))
        for pol, tr in ranks:
            capi.check(lib.mx_qmix_p2p_reduce(tr.handle, None))
            capi.check(lib.mx_qmix_apply(tr.handle, None))
            tr.soft_target_updates()
        i0, i1 = ranks[0][1]._info, ranks[1][1]._info
        assert float(i0[7]) == 0.0 and float(i1[7]) == 0.0          # no time-out / missing peer flag
        for k, key in enumerate(("loss", "grad_norm", "Q_tot")):
            assert float(i0[k]) == float(i1[k])
            want = float(g["s%d.%s" % (s, key)])
            assert abs(float(i0[k]) - want) <= 1e-4 * abs(want), (s, key, float(i0[k]), want)
        assert torch.equal(ranks[0][1].theta, ranks[1][1].theta)       # replicas stay bit-identical
        assert torch.equal(ranks[0][1].theta_tgt, ranks[1][1].theta_tgt)
        for k, v in ranks[0][0].q_network.state_dict().items():
            want = g["s%d.agent.%s" % (s, k)]
            assert np.abs(v.numpy() - want).max() <= 2.0 * cfg.lr, k     # same bound family as qmix_checks (Adam step ~ lr)
    # block layout (csrc/p2p.cu mx_qmix_p2p_block_bytes): slots of both parities [2][world][slot] | 64 flag words | the fused kernel's line arrays
    slot = -(-(ranks[0][1].P + 8) // 64) * 64
    assert n == 2 * 2 * slot + 64 + 2 * 2 * 2 * slot
    flags = blocks[0][2 * 2 * slot:2 * 2 * slot + 64].view(torch.int32)
    assert int(flags[0]) == 2 and int(flags[1]) == 2


def test_set_peers_argument_errors(emu_engine):
    capi = emu_engine
    lib = capi.lib()
    g = load_golden("qmix_small")
    L, cfg, B, T, steps = oracle_from_golden(g)
    args, pol, tr = qc.build_trainer(cfg, B, T)                   # world_size 1
    ptrs = (C.c_void_p * 2)(1, 2)
    cnt = torch.zeros(4, dtype=torch.int32)
    assert lib.mx_qmix_set_peers(tr.handle, 0, 2, ptrs, capi.ptr(cnt)) != 0     # world does not match cfg.world_size
    assert lib.mx_qmix_p2p_publish(tr.handle, None) != 0
