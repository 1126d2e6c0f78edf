"""Multi-GPU host logic on CPU: range sharding (dspb200.sharding) and the Welch all-reduce protocol, exercised with
world_size = 2 over the gloo backend.  The per-rank compute is done by the oracle here (no GPU in this container);
on the GPU box the same ranges feed dspb200_*_exec_range_dev and NCCL (bench.py --gpus N)."""
import os

import numpy as np
import pytest

from conftest import relerr

import dspb200 as dsp
from dspb200 import sharding as sh
from oracle import dspbase as od
from oracle import filters as of
from oracle import periodograms as op
from oracle import windows as ow


def test_split_range_tiles():
    for total in (0, 1, 7, 100, 2 ** 26 + 4096):
        for world in (1, 2, 3, 8):
            parts = [sh.split_range(total, world, r) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == total
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in parts]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_conv_shards_reassemble(world):
    rng = np.random.default_rng(3)
    u = rng.standard_normal(5000) + 1j * rng.standard_normal(5000)
    v = rng.standard_normal(97) + 1j * rng.standard_normal(97)
    full = od.conv_exact(u, v)
    out = np.zeros_like(full)
    for r in range(world):
        s = sh.conv_shard(u.size, v.size, full.size, world, r)
        local = u[s.in_begin:s.in_end]                                       # what the rank holds
        y = od.conv_exact(local, v)                                          # y[i] <-> global output s.in_begin + i
        out[s.out_begin:s.out_begin + s.out_count] = y[s.out_begin - s.in_begin: s.out_begin - s.in_begin + s.out_count]
    assert relerr(out, full) < 1e-14


@pytest.mark.parametrize("world", [1, 2, 3, 8])
@pytest.mark.parametrize("n,nov", [(256, 128), (1024, 768), (100, 0), (64, 63)])
def test_welch_shards_cover_every_segment_once(world, n, nov):
    length = 20000
    shards = [sh.welch_stream_shard(length, n, nov, world, r) for r in range(world)]
    k = shards[0].k_total
    assert k == (length - n) // (n - nov) + 1
    assert shards[0].seg_begin == 0 and shards[-1].seg_end == k
    assert all(shards[i].seg_end == shards[i + 1].seg_begin for i in range(world - 1))
    rng = np.random.default_rng(5)
    x = rng.standard_normal(length)
    truth, _ = op.welch_pgram(x, n, nov, window=ow.hanning)
    hop = n - nov
    total = np.zeros_like(truth)
    for s in shards:
        if s.seg_end > s.seg_begin:
            assert s.sample_begin == s.seg_begin * hop and s.sample_end <= length
            loc = x[s.sample_begin:s.sample_end]
            p, _ = op.welch_pgram(loc, n, nov, window=ow.hanning)          # mean over local segments
            total += p * (s.seg_end - s.seg_begin) / k                        # == sum with the GLOBAL 1/(k r)
    assert relerr(total, truth) < 1e-13


@pytest.mark.parametrize("world", [1, 2, 5])
def test_resample_shards_reassemble(world):
    from fractions import Fraction
    rng = np.random.default_rng(9)
    x = rng.standard_normal(3001)
    for rate in (Fraction(3, 2), Fraction(5, 9), Fraction(7, 1)):
        h = of.resample_filter(rate)
        full = of.resample(x, rate, h)
        n0, phi0 = dsp.resample_phase(h.size, rate)
        tpp = -(-h.size // rate.numerator)
        got = np.zeros_like(full)
        for r in range(world):
            s = sh.resample_shard(x.size, full.size, rate.numerator, rate.denominator, n0, phi0, tpp, world, r)
            xl = np.zeros_like(x)
            xl[s.in_begin:s.in_end] = x[s.in_begin:s.in_end]                  # everything the rank does not hold is zero
            got[s.j_begin:s.j_begin + s.out_count] = of.resample(xl, rate, h)[s.j_begin:s.j_begin + s.out_count]
        assert np.array_equal(got, full)


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(11)
        length, n, nov = 1 << 15, 512, 256
        x = rng.standard_normal(length).astype(np.float32)                   # same stream on every rank (same seed)
        s = sh.welch_stream_shard(length, n, nov, world, rank)
        norm2 = float(np.sum(ow.hanning(n) ** 2))
        segs = op.arraysplit(x[s.sample_begin:s.sample_end], n, nov, n, ow.hanning(n), f64=True)
        import scipy.fft as sfft
        part = np.zeros(n // 2 + 1)
        op.fft2pow_acc(part, sfft.rfft(segs, axis=1), n, s.k_total * norm2, True)     # scaled by the GLOBAL k*r
        t = torch.from_numpy(part)
        dist.all_reduce(t)                                                    # the one collective of the Welch path
        truth, _ = op.welch_pgram(x, n, nov, window=ow.hanning, f64=True)
        cb, ce = sh.channel_shard(64, world, rank)                            # channel sharding: no collective
        q.put((rank, float(relerr(t.numpy(), truth)), (cb, ce)))
    finally:
        dist.destroy_process_group()


def test_welch_allreduce_world2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort()
    assert all(err < 1e-13 for _, err, _ in res)
    assert res[0][2] == (0, 32) and res[1][2] == (32, 64)
