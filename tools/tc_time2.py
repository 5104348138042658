"""Is a chain of tcgen05.mma into ONE accumulator latency-bound?  One thread issues round-robin into 1..8 independent
accumulators (kind::tf32, M=128, K=8, A in tensor memory): cycles per MMA."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scann_b200._abi import lib, check, require_gpu
require_gpu()
out = torch.zeros(4, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for ncols in (32, 64, 128):
    for nacc in (1, 2, 3, 4, 6, 8):
        if nacc * ncols > 256:
            continue
        r = []
        for nmma in (48 * nacc, 480 * nacc):
            check(lib.scann_tc_time(out.data_ptr(), 16 | 8 | 1 | (nacc << 8), nmma, ncols, st)); torch.cuda.synchronize()
            r.append(out[0].item())
        print(f"N={ncols:3d} accumulators={nacc}: {r[0]:.1f} cycles/MMA over {48*nacc}, {r[1]:.1f} over {480*nacc} (math floor {128 * ncols * 8 / 2048:.0f})")
