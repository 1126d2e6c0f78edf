"""What one rank of the strong-scaling leg costs on its own: rank 3 of 8 of the 2^26-sample conv -> Welch step (2^23 samples
per GPU), no process group.  Prints the time of each stage in a launch loop (CUDA events) for shard sizes 2^20 .. 2^26;
run under `ncu --metrics gpu__time_duration.sum` the launch list gives the pure kernel durations beside it."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench


class FakeDist:
    def __init__(self, world, rank):
        self.world, self.rank, self.pg = world, rank, None
        self.dev = torch.device("cuda", 0)
        self.torch = torch

    def sync_all(self):
        torch.cuda.synchronize()

    def max_over_ranks(self, v):
        return [float(x) for x in v]


def loop_ms(fn, reps):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    from dspb200 import _lib
    _lib.check(_lib.lib.dspb200_set_device(0))
    reps = int(os.environ.get("REPS", "50"))
    for world in (64, 32, 16, 8, 4, 2, 1):
        n = (1 << 26) // world
        d = FakeDist(world, min(3, world - 1))
        cw = bench.ConvWelch(d, n, int(os.environ.get("NFFT", "0")))
        conv = loop_ms(cw.conv, reps)
        welch = loop_ms(cw.welch, reps)
        both = loop_ms(lambda: (cw.conv(), cw.welch()), reps)
        print(json.dumps({"world": world, "samples_per_gpu": n, "conv_ms": round(conv, 4), "welch_ms": round(welch, 4),
                          "step_ms": round(both, 4), "ideal_conv_ms": None, "blocks": (cw.out_cnt + 12287) // 12288,
                          "segments": cw.seg_end - cw.seg_begin}), flush=True)
        del cw
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
