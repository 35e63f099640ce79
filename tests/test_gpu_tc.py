"""GPU: the tcgen05 building blocks (operand images, descriptors, bulk-copy ring, TMEM read-back)
against fp64 matmuls. Split-bf16 (3 products) must reach ~2^-16 relative accuracy."""
import numpy as np
import pytest
import torch

from mpg_b200.config import default_args

pytestmark = pytest.mark.gpu


def _engine():
    from mpg_b200.engine import Engine
    return Engine(debug_lib=True, **vars(default_args('NADP', 'PathTracking-v0')))   # libmpg_b200_dbg.so: the probes are not in the product library


@pytest.mark.parametrize('kind,repeats', [(0, 1), (0, 3), (1, 1), (1, 2), (2, 1), (2, 2)])
def test_tc_gemm_kinds(kind, repeats):
    e = _engine()
    g = torch.Generator(device='cpu').manual_seed(kind * 10 + repeats)
    if kind == 0:
        X = torch.randn(128, 256, generator=g); W = torch.randn(256, 256, generator=g) / 16
        ref = X.double() @ W.double().T
    elif kind == 1:
        X = torch.randn(128, 16, generator=g); W = torch.randn(16, 256, generator=g)
        ref = X.double() @ W.double()
    else:
        X = torch.randn(128, 256, generator=g); W = torch.randn(16, 256, generator=g) / 16
        ref = X.double() @ W.double().T
    Z = e.tc_selftest(kind, e.dev(X), e.dev(W), repeats)
    torch.cuda.synchronize()
    err = (Z.cpu().double() - ref).norm() / ref.norm()
    worst = (Z.cpu().double() - ref).abs().max() / ref.abs().max()
    print(kind, repeats, 'rel-L2', float(err), 'scaled Linf', float(worst))
    assert err < 2e-5 and worst < 1e-4


def test_cta_pair_probe_matches_fp64():
    """cta_group::2 building block (tc_pair_probe.cuh): two CTAs of a cluster, each with 128 rows of A and half of every
    weight stage, one M = 256 UMMA stream issued by the leader; checked like the single-CTA GEMM."""
    e = _engine()
    g = torch.Generator(device='cpu').manual_seed(5)
    X = torch.randn(256, 256, generator=g); W = torch.randn(256, 256, generator=g) / 16
    Z = e.tc_selftest(5, e.dev(X), e.dev(W), 1)
    torch.cuda.synchronize()
    ref = X.double() @ W.double().T
    err = (Z.cpu().double() - ref).norm() / ref.norm()
    assert err < 2e-5
