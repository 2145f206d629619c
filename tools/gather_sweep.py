"""HBM roofline of the replay gather ("demonstrate >= 60 % of HBM peak on C4 or a batch sweep (B >= 256)").

Times k_gather alone at the 8m PER shapes (BASELINE configs[3]: 8 agents, obs 80, state 168, 14 actions, T = 120) for batch
sizes 64 .. 1024 out of a 2 000-episode replay (1 GB: every launch reads episodes that are not in L2), CUDA events around K
back-to-back gathers with K different device-resident index sets; algorithmic bytes = 2 x the sampled episodes' bytes (read + write).
Prints one JSON line per batch size.  Peak = MEASURED_PEAKS.json hbm_gbs.
"""
import ctypes as C, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "off-policy_b200")):
    sys.path.insert(0, p)
import numpy as np, torch
from offpolicy._b200 import capi, factory as rc

N, O, A, S, T, E = 8, 80, 14, 168, 120, int(os.environ.get("SWEEP_E", "2000"))
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
capi.lib()
buf = rc.make_rec_buffers(N, O, A, S, T, E, rng="device", max_batch=1024)
pb = buf.policy_buffers["policy_0"]
rs = np.random.default_rng(0)
for c in range(0, E, 100):
    n = min(100, E - c)
    f = [rs.standard_normal((T + 1, n, N, O), dtype=np.float32), np.repeat(rs.standard_normal((T + 1, n, 1, S), dtype=np.float32), N, 2),
         np.eye(A, dtype=np.float32)[rs.integers(0, A, (T, n, N))], np.zeros((T, n, N, 1), np.float32), np.zeros((T, n, N, 1), np.float32),
         np.zeros((T, n, 1), np.float32), np.ones((T + 1, n, N, A), np.float32)]
    buf.insert(n, *[rc.pd(x) for x in f])
L = pb.L
ep_bytes = 4 * (L.ep_obs + L.ep_share + L.ep_acts + L.ep_avail + L.ep_rew + L.ep_dones + L.ep_dones_env + L.ep_actidx)
lib, K = capi.lib(), 20
stream = torch.cuda.current_stream()
for tma in (2, 0):      # TMA tile copies (forced at every size) vs 16-byte vector loads
    lib.mx_set_option(b"gather_tma", tma)
    for B in (64, 128, 256, 512, 1024):
        sets = [torch.from_numpy(rs.permutation(E)[:B].astype(np.int64)).cuda() for _ in range(K)]     # distinct episodes: no reuse inside a batch
        for k in range(3):
            capi.check(lib.mx_replay_gather(pb.handle, capi.ptr(sets[k]), B, capi.stream_ptr()))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for k in range(K):
            capi.check(lib.mx_replay_gather(pb.handle, capi.ptr(sets[k]), B, capi.stream_ptr()))
        e1.record(stream)
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / K * 1e3
        gbs = 2.0 * B * ep_bytes / (us * 1e-6) / 1e9
        print(json.dumps(dict(kernel="k_gather_tma" if tma else "k_gather", workload="qmix_8m shapes", batch=B, episode_bytes=int(ep_bytes), us_per_gather=round(us, 2),
                              achieved_gbs=round(gbs, 1), peak_gbs=peak, frac=round(gbs / peak, 3),
                              note="includes one 8*B-byte D2D index copy per gather; replay %d episodes = %.2f GB" % (E, E * ep_bytes / 1e9))))

lib.mx_set_option(b"gather_tma", 1)
