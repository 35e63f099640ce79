"""CPU: the C-ABI library loads and exports every symbol include/mpg_b200.h declares; the product
path fails loudly without a GPU (no CPU fallback)."""
import ctypes
import os
import re
import subprocess

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols(debug=False):
    """Entry points include/mpg_b200.h declares: the product ABI, or (debug=True) the ones inside #ifdef MPG_DEBUG_PROBES."""
    src = open(os.path.join(ROOT, 'include', 'mpg_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    dbg = ''.join(re.findall(r'#ifdef MPG_DEBUG_PROBES(.*?)#endif', src, flags=re.S))
    if debug:
        src = dbg
    else:
        src = re.sub(r'#ifdef MPG_DEBUG_PROBES.*?#endif', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(mpg_[a-z_0-9]+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    from mpg_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), f'{name} declared in include/mpg_b200.h but not exported'
    assert sorted(_lib.SYMBOLS) == declared, 'python binding table out of sync with the header'
    _lib.load()  # sets argtypes for every symbol
    # the development probes live in the debug build only: the product library exports the boundary and nothing else
    probes = _declared_symbols(debug=True)
    assert sorted(_lib.DEBUG_SYMBOLS) == probes and probes
    for name in probes:
        assert not hasattr(lib, name), f'{name} is a development probe and must not be exported by the product library'
    dbg = ctypes.CDLL(_lib.DEBUG_LIB_PATH)
    for name in declared + probes:
        assert hasattr(dbg, name), f'{name} missing from the debug build'
    exported = subprocess.run(['nm', '-D', '--defined-only', _lib.LIB_PATH], capture_output=True, text=True).stdout
    extra = sorted(set(re.findall(r' T (mpg_[a-z_0-9]+)', exported)) - set(declared))
    assert not extra, f'exported but not declared in include/mpg_b200.h: {extra}'


def test_struct_layouts_match_header():
    from mpg_b200._lib import MpgConfig, RolloutParams
    # mpg_config: 6 int32 + float + 16 floats + 3 floats + 2 int32 ; mpg_rollout_params has 3 x 64-bit fields
    assert ctypes.sizeof(MpgConfig) == 4 * (6 + 1 + 16 + 3 + 2)
    assert ctypes.sizeof(RolloutParams) == 4 * 4 + 4 * 8 + 4 * 8 + 4 * 3 + 4 + 8 * 3 + 4 * 2  # incl. padding before int64


@pytest.mark.skipif(torch.cuda.is_available(), reason='checks the no-GPU behaviour')
def test_no_cpu_fallback():
    from mpg_b200.engine import Engine
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        Engine(env_id='PathTracking-v0', obs_dim=6, act_dim=2, obs_scale=[1.] * 6, rew_scale=0.01, rew_shift=0.,
               gamma=0.98)
    from mpg_b200 import _lib
    lib = _lib.load()
    cfg = _lib.MpgConfig()
    cfg.env, cfg.obs_dim, cfg.act_dim, cfg.hidden, cfg.max_rows, cfg.max_horizon = 0, 6, 2, 256, 64, 1
    h = ctypes.c_void_p()
    assert lib.mpg_create(ctypes.byref(cfg), ctypes.byref(h)) != 0
    assert b'no CPU fallback' in lib.mpg_last_error(None)


def test_host_rule_based_weights_match_reference():
    import numpy as np
    from mpg_b200.learners.base import rule_based_weights
    from tests.util import load_golden
    for name in ('rule_weights', 'rule_weights3'):
        case, gold = load_golden(name)
        for i, ite in enumerate(case['iterations']):
            w = rule_based_weights(ite, 9000, 0.1, case['rollout_list'])
            assert np.allclose(w, gold['ws__f32'][i], rtol=2e-5, atol=1e-9), (name, ite, w, gold['ws__f32'][i])


@pytest.mark.skipif(torch.cuda.is_available(), reason='checks the no-GPU behaviour')
def test_trainer_stack_imports_and_fails_loudly_without_gpu():
    """The callers either side of the path (worker / optimizer / evaluator / trainer / buffer) import on a CPU-only
    box and refuse to run there instead of silently falling back."""
    import mpg_b200.buffer  # noqa: F401
    import mpg_b200.evaluator  # noqa: F401
    import mpg_b200.optimizer  # noqa: F401
    import mpg_b200.worker  # noqa: F401
    from mpg_b200.config import default_args
    from mpg_b200.trainer import Trainer
    args = default_args('MPG-v2', 'PathTracking-v0', batch_size=64, num_agent=8, explore_sigma=0.1, max_buffer_size=1000,
                        replay_starts=64, buffer_log_interval=10 ** 9, num_eval_agent=8, num_eval_episode=1,
                        fixed_steps=5, eval_interval=10 ** 9, log_interval=10 ** 9, max_iter=1, log_dir=None)
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        Trainer(args)
