"""Range sharding of the hot path across GPUs: pure index arithmetic, no device code.

One process per GPU.  A long stream is cut into `world` contiguous sample ranges; every rank
  * convolves / resamples its own OUTPUT range (needs an nv-1 / tapsPerPhase-1 sample halo on the left) -- no collective;
  * accumulates the Welch power of the segments that START in its range (needs up to n - hop samples past its right
    edge), scaled by the GLOBAL 1/(k r); the only exchange is one sum all-reduce of the power vector;
  * or, for multi-channel input, simply owns a contiguous block of channels -- no collective.
"""
from dataclasses import dataclass


def split_range(total, world, rank):
    """Balanced contiguous split of range(total) into `world` parts; returns (begin, end) of part `rank`."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    base, rem = divmod(total, world)
    b = rank * base + min(rank, rem)
    return b, b + base + (1 if rank < rem else 0)


@dataclass(frozen=True)
class ConvShard:
    out_begin: int      # first output sample this rank produces
    out_count: int
    in_begin: int       # first input sample it must hold (out_begin - (nv-1), clipped at 0)
    in_end: int         # one past the last input sample it must hold (clipped at nu)


def conv_shard(nu, nv, nout, world, rank, align=1):
    """Outputs [out_begin, out_begin+out_count) of out[m] = sum_j u[j] v[m-j], m < nout, and the input samples they read.
    `align`: shard boundaries are multiples of `align`.  With align = the overlap-save block length L = nfft - nv + 1 every
    rank cuts its range into the same blocks as a single-GPU run, so the sharded result is bit-identical to it (an output's
    rounding depends on its position inside its block); with align = 1 the results agree to the FFT's rounding error."""
    nblk = -(-nout // align)
    bb, be = split_range(nblk, world, rank)
    b, e = min(nout, bb * align), min(nout, be * align)
    return ConvShard(b, e - b, max(0, b - (nv - 1)), max(0, min(nu, e)))


@dataclass(frozen=True)
class WelchShard:
    k_total: int        # segments of the whole signal (src/periodograms.jl:49-50)
    seg_begin: int      # segments [seg_begin, seg_end) start inside this rank's sample range
    seg_end: int
    sample_begin: int   # samples the rank must hold to transform its segments
    sample_end: int


def welch_stream_shard(length, n, noverlap, world, rank):
    """Segments whose first sample lies in the rank's contiguous sample range of a `length`-sample stream."""
    hop = n - noverlap
    k = (length - n) // hop + 1 if length >= n else 0
    lo, hi = split_range(length, world, rank)
    sb = min(k, -(-lo // hop))          # ceil(lo / hop)
    se = min(k, -(-hi // hop))
    if se <= sb:
        return WelchShard(k, sb, sb, lo, lo)
    return WelchShard(k, sb, se, sb * hop, (se - 1) * hop + n)


@dataclass(frozen=True)
class ResampleShard:
    j_begin: int
    out_count: int
    in_begin: int
    in_end: int


def resample_shard(nx, nout, interp, decim, n0, phi0, taps_per_phase, world, rank):
    """Outputs [j_begin, j_begin+out_count) of the polyphase resampler and the input samples they read."""
    b, e = split_range(nout, world, rank)
    if e <= b:
        return ResampleShard(b, 0, 0, 0)
    n_first = n0 + (phi0 + b * decim) // interp
    n_last = n0 + (phi0 + (e - 1) * decim) // interp
    return ResampleShard(b, e - b, max(0, n_first - (taps_per_phase - 1)), max(0, min(nx, n_last + 1)))


def channel_shard(nchan, world, rank):
    """Contiguous block of channels (columns) owned by `rank`."""
    return split_range(nchan, world, rank)
