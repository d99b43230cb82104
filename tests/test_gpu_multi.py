"""Multi-GPU driver (-m gpu, needs >= 2 devices): targets sharded across GPUs, positions
allgathered over NCCL each step.  Single-process mode (ngpus=2, the C
driver's mode); the one-process-per-GPU mode is exercised by bench.py under torchrun."""
import os

import numpy as np
import pytest

import oracle_lib as O
from nbodysim_b200 import Simulation, capi, ic
from nbodysim_b200.bodies import acc3, pos3, vel3

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


def _ngpu():
    try:
        import torch

        return torch.cuda.device_count()
    except Exception:
        return 0


needs2 = pytest.mark.skipif(_ngpu() < 2, reason="needs 2 GPUs")


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


@needs2
@pytest.mark.parametrize("exchange", [0, 1])
def test_two_gpu_refcompat_trajectory_bitexact_vs_golden(exchange):
    """sharding targets does not change any per-target sum: still bit-exact vs the reference -- with the
    peer-to-peer integrate-and-push exchange (0 = auto) and with ncclAllGather (1)"""
    g = np.load(os.path.join(G, "disc1024.npz"))
    with Simulation(g["bodies"], dt=float(g["dt"]), eps=float(g["eps"]), dims=2, rsqrt_mode=capi.RSQRT_REFCOMPAT,
                    ngpus=2, device_ids=[0, 1], exchange=exchange) as s:
        assert s.info()["world"] == 2
        assert s.info()["p2p_exchange"] == (1 if exchange == 0 else 0)
        s.step(int(g["nsteps"]))
        out = s.bodies
    for f in ("pos", "vel", "acc"):
        assert np.array_equal(bits(out[f]), bits(g["end_" + f])), f


@needs2
@pytest.mark.parametrize("exchange", [0, 1])
def test_two_gpu_fast_matches_one_gpu_and_exact(exchange):
    b = ic.plummer(20000, seed=3, dims=3)
    with Simulation(b, dt=1e-3, eps=0.01, dims=3, device_ids=[0]) as s1, \
         Simulation(b, dt=1e-3, eps=0.01, dims=3, ngpus=2, device_ids=[0, 1], exchange=exchange) as s2:
        s2.attract()
        a2 = acc3(s2.download()).astype(np.float64)
        ex = O.orc_exact_acc(b, float(np.float32(0.01)), dims=3)
        r = np.linalg.norm(a2 - ex, axis=1) / np.linalg.norm(ex, axis=1)
        assert np.percentile(r, 99) <= 1e-5
        s1.step(20)
        s2.step(20)
        o1, o2 = s1.bodies.copy(), s2.bodies.copy()
        k1, w1, p1 = s1.energy()
        k2, w2, p2 = s2.energy()
    np.testing.assert_allclose(pos3(o2), pos3(o1), rtol=0, atol=5e-6)
    np.testing.assert_allclose(vel3(o2), vel3(o1), rtol=0, atol=5e-5)
    assert abs(k1 - k2) <= 1e-6 * abs(k1) and abs(w1 - w2) <= 1e-6 * abs(w1)


@needs2
def test_two_gpu_f64_matches_exact():
    b = ic.plummer(5000, seed=8, dims=3)
    with Simulation(b, eps=0.01, dims=3, precision=capi.PRECISION_F64, ngpus=2, device_ids=[0, 1]) as s:
        s.attract()
        _, _, a = s.download_f64()
    ex = O.orc_exact_acc(b, float(np.float32(0.01)), dims=3)
    assert (np.linalg.norm(a - ex, axis=1) / np.linalg.norm(ex, axis=1)).max() <= 1e-12


@needs2
def test_two_gpu_barnes_hut_bitexact_vs_golden():
    """every GPU builds the (replicated) tree and walks its own shard of targets: still bit-exact"""
    g = np.load(os.path.join(G, "bh2000.npz"))
    with Simulation(g["bodies"], dt=float(g["dt"]), dims=2, theta=1.0, eps=1.0, force_algo=capi.FORCE_BARNES_HUT,
                    rsqrt_mode=capi.RSQRT_REFCOMPAT, integ_flags=capi.INTEG_CLAMP | capi.INTEG_BOUNDARY,
                    ngpus=2, device_ids=[0, 1]) as s:
        s.step(int(g["nsteps"]))
        out = s.bodies
    for f in ("pos", "vel", "acc"):
        assert np.array_equal(bits(out[f]), bits(g["end_" + f])), f


@needs2
def test_two_gpu_exchanges_agree_bitwise_over_many_steps():
    """P2P push and NCCL allgather move the same bytes: 60 steps, identical state"""
    b = ic.plummer(30000, seed=11, dims=3)
    outs = []
    for ex in (0, 1):
        with Simulation(b, dt=1e-3, eps=0.01, dims=3, ngpus=2, device_ids=[0, 1], exchange=ex) as s:
            s.step(60)
            outs.append(s.bodies.copy())
    for f in ("pos", "pos_z", "vel", "vel_z", "acc", "acc_z"):
        assert np.array_equal(bits(outs[0][f]), bits(outs[1][f])), f


@needs2
@pytest.mark.parametrize("bh_walk", [1, 2])
def test_two_gpu_octree_walks_equal_one_gpu(bh_walk):
    """each GPU walks the compacted list of its own targets (per-thread and warp-cooperative walk): the accelerations
    of the 3-D octree path equal the one-GPU run bit for bit, and so does the state after a few steps"""
    b = ic.plummer(20000, seed=21, dims=3)
    outs = []
    for ng in (1, 2):
        with Simulation(b, dt=1e-3, eps=0.01, dims=3, theta=0.5, force_algo=capi.FORCE_BARNES_HUT, bh_fix_near_leaves=1,
                        rsqrt_mode=capi.RSQRT_REFCOMPAT, bh_walk=bh_walk, ngpus=ng, device_ids=[0, 1][:ng]) as s:
            s.step(4)
            outs.append(s.bodies.copy())
    for f in ("pos", "pos_z", "vel", "vel_z", "acc", "acc_z"):
        assert np.array_equal(bits(outs[0][f]), bits(outs[1][f])), f
