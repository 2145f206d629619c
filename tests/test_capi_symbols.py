"""The nvcc-built C-ABI library loads (no GPU needed) and exports every function include/marl_b200.h declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "marl_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mx_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_a_real_surface():
    names = declared_functions()
    assert len(names) >= 30 and "mx_qmix_step" in names and "mx_replay_sample_uniform" in names


def test_cuda_library_exports_every_declared_symbol():
    import sys
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge
    lib_path = ge.LIB if os.path.exists(ge.LIB) else ge.build()
    lib = ctypes.CDLL(lib_path)
    missing = [n for n in declared_functions() if not hasattr(lib, n)]
    assert not missing, missing
    lib.mx_is_cuda_build.restype = ctypes.c_int
    assert lib.mx_is_cuda_build() == 1
    lib.mx_abi_version.restype = ctypes.c_int
    from offpolicy._b200 import capi
    header = open(os.path.join(ROOT, "include", "marl_b200.h")).read()
    want = int(re.search(r"#define\s+MX_ABI_VERSION\s+(\d+)", header).group(1))
    assert lib.mx_abi_version() == want == capi.ABI_VERSION      # header, library and ctypes mirrors agree


def test_product_loader_refuses_to_run_without_gpu():
    """No CPU fallback: the product loader must raise when there is no CUDA device."""
    import pytest
    import torch
    from offpolicy._b200 import capi
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    capi._uninstall_for_tests()
    with pytest.raises(capi.MxError):
        capi.lib()


def test_layout_queries_need_no_gpu():
    from offpolicy._b200 import capi
    lib = capi.load_symbols_only()
    cfg = capi.ReplayCfg(5000, 60, 3, 30, 48, 9, 1, 1, 0, 64, 0.6)
    L = capi.ReplayLayout()
    assert lib.mx_replay_layout_query(ctypes.byref(cfg), ctypes.byref(L)) == 0
    assert L.obs_ld == 32 and L.act_ld == 12 and L.tree_cap == 8192
    assert L.ep_obs == 61 * 3 * 32 and L.total_bytes > 5000 * L.ep_obs * 4
    q = capi.QmixCfg(n_agents=3, obs_dim=30, act_dim=9, state_dim=48, hidden=64, mixer_hidden=32, hyper_hidden=64, hyper_layers=2,
                     episode_len=60, max_batch=32)
    total = ctypes.c_int64()
    n = lib.mx_qmix_param_layout(ctypes.byref(q), None, 0, ctypes.byref(total))
    assert n == 36          # 22 agent tensors + 14 mixer tensors
    arr = (capi.ParamEntry * n)()
    lib.mx_qmix_param_layout(ctypes.byref(q), arr, n, ctypes.byref(total))
    logical = sum(e.rows * (e.cols if e.cols else 1) for e in arr)
    assert logical == 55782  # reference parameter count at 3m shapes
    assert total.value >= logical and all(e.offset % 4 == 0 for e in arr)
    q.hidden = 128
    assert lib.mx_qmix_param_layout(ctypes.byref(q), None, 0, ctypes.byref(total)) < 0
    assert b"hidden_size" in lib.mx_last_error()


def test_struct_mirrors_have_the_library_sizes():
    """Every ctypes mirror in offpolicy/_b200/capi.py and the stub printed in INTEGRATION.md section 2 (what a maintainer copies) must
    have sizeof() equal to the C struct's: a short struct makes the kernels read garbage strides."""
    from offpolicy._b200 import capi
    lib = capi.load_symbols_only()
    for name, t in (("mx_batch", capi.Batch), ("mx_replay_cfg", capi.ReplayCfg), ("mx_replay_layout", capi.ReplayLayout), ("mx_qmix_cfg", capi.QmixCfg),
                    ("mx_maddpg_cfg", capi.MaddpgCfg), ("mx_param_entry", capi.ParamEntry), ("mx_policy_step_args", capi.PolicyStepArgs),
                    ("mx_episodes", capi.Episodes)):
        assert int(lib.mx_sizeof(name.encode())) == ctypes.sizeof(t), name
    assert int(lib.mx_sizeof(b"no_such_struct")) == -1
    # the INTEGRATION.md stub: run its struct definitions and compare
    md = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    block = md[md.index("# offpolicy/utils/b200.py"):]
    block = block[:block.index("```")]
    defs = block[block.index("class ReplayCfg"):block.index("assert lib.mx_abi_version()")]
    ns = {"C": ctypes}
    exec(defs, ns)
    for name, cls in (("mx_replay_cfg", "ReplayCfg"), ("mx_batch", "Batch"), ("mx_replay_layout", "ReplayLayout")):
        assert int(lib.mx_sizeof(name.encode())) == ctypes.sizeof(ns[cls]), "INTEGRATION.md stub of %s is out of date" % name
