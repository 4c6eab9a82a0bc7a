"""The C-ABI shared library: it loads, exports every symbol include/dqn_b200.h declares, validates
configurations on the host, and fails loudly (no CPU fallback) when there is no sm_100a device."""
import ctypes as C
import os
import re

import pytest

import dqn_b200

_lib = dqn_b200.pkg._lib


def header_functions():
    text = open(_lib.HEADER_PATH).read()
    return sorted(set(re.findall(r"DQN_API\s+(?:const\s+char\*|int)\s+(dqn_\w+)\s*\(", text)))


def make_cfg(**kw):
    cfg = _lib.DqnConfig()
    cfg.struct_size = C.sizeof(_lib.DqnConfig)
    vals = dict(device=0, n_agents=1, obs_dim=9, num_actions=4, hidden1=32, hidden2=64, batch_size=64,
                buffer_size=1000, gamma=0.99, opt_kind=1, lr=2e-4, b1=0.9, b2=0.999, eps=1e-8, eps_root=0.0,
                weight_decay=1e-4, seed=1)
    vals.update(kw)
    for k, v in vals.items():
        setattr(cfg, k, v)
    return cfg


def test_library_is_built_in_tree():
    assert os.path.exists(_lib.LIB_PATH), "run __graft_entry__.build() first"
    assert os.path.dirname(_lib.LIB_PATH).endswith(os.path.join("deep-q-learning_b200", "csrc"))


def test_every_declared_symbol_is_exported_and_bound():
    lib = _lib.load()
    names = header_functions()
    assert len(names) >= 25
    assert set(names) == set(_lib.PROTOTYPES), set(names) ^ set(_lib.PROTOTYPES)
    for n in names:
        assert getattr(lib, n) is not None
    assert lib.dqn_abi_version() == 1


def test_config_struct_matches_header_layout():
    lib = _lib.load()
    n = C.c_uint64(0)
    assert lib.dqn_arena_bytes(C.byref(make_cfg()), C.byref(n)) == 0      # struct_size accepted by the C side
    # params(4 x 2760 f32) + ring (1000 x 96 B) + 8 MiB staging + taps ...
    assert n.value > 1000 * 96 + 8 * 2**20
    bad = make_cfg()
    bad.struct_size -= 8
    assert lib.dqn_arena_bytes(C.byref(bad), C.byref(n)) == -1
    assert b"struct_size" in lib.dqn_last_error()


@pytest.mark.parametrize("kw,needle", [
    (dict(obs_dim=0), b"obs_dim"), (dict(obs_dim=17), b"obs_dim"), (dict(num_actions=1), b"num_actions"),
    (dict(num_actions=8), b"num_actions"), (dict(hidden1=64), b"(32, 64)"), (dict(batch_size=0), b"batch_size"),
    (dict(batch_size=2048), b"batch_size"), (dict(buffer_size=0), b"buffer_size"), (dict(n_agents=0), b"n_agents"),
    (dict(opt_kind=5), b"opt_kind"),
])
def test_invalid_configs_are_rejected_with_a_message(kw, needle):
    lib = _lib.load()
    n = C.c_uint64(0)
    assert lib.dqn_arena_bytes(C.byref(make_cfg(**kw)), C.byref(n)) == -1
    assert needle in lib.dqn_last_error()


def test_arena_scales_with_ring_and_population():
    lib = _lib.load()
    a, b, c = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
    lib.dqn_arena_bytes(C.byref(make_cfg(buffer_size=1000)), C.byref(a))
    lib.dqn_arena_bytes(C.byref(make_cfg(buffer_size=2000)), C.byref(b))
    lib.dqn_arena_bytes(C.byref(make_cfg(buffer_size=1000, n_agents=3)), C.byref(c))
    assert 96000 <= b.value - a.value <= 96000 + 256                       # one 96-byte record per slot at D=9
    assert c.value - a.value >= 2 * (96000 + 4 * 2760 * 4)
    lib.dqn_arena_bytes(C.byref(make_cfg(obs_dim=16, buffer_size=1000)), C.byref(b))
    assert b.value - a.value >= 1000 * (160 - 96)                           # D=16 -> 160-byte records


def test_null_handle_calls_fail_cleanly():
    lib = _lib.load()
    assert lib.dqn_synchronize(None) == -1
    assert lib.dqn_sync_target(None, 0, 1) == -1
    assert lib.dqn_train_step(None, 0, 1, 1, None, None) == -1
    assert lib.dqn_destroy(None) == 0


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    lib = _lib.load()
    h = C.c_void_p()
    rc = lib.dqn_create(C.byref(make_cfg()), C.byref(h))
    assert rc == -3 and not h.value                                          # DQN_E_ARCH
    assert b"no CPU fallback" in lib.dqn_last_error()
    with pytest.raises(dqn_b200.DqnError):
        dqn_b200.DqnEngine(9, 4, 100, 8, 0.99, dqn_b200.adam(1e-4))
    with pytest.raises(dqn_b200.DqnError):
        dqn_b200.ReplayBuffer(100, (100, 9), (100,))
