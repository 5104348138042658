"""Cycles per tcgen05.mma (kind::tf32, M=128, K=8) for N = 32..256, A from shared memory (SS) or tensor memory (TS)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scann_b200._abi import lib, check, require_gpu
require_gpu()
out = torch.zeros(4, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for ncols in (128, 64, 32, 16):
    for mode in (5, 13, 12):
        r = []
        for nmma in (48, 480):
            check(lib.scann_tc_time(out.data_ptr(), mode, nmma, ncols, st)); torch.cuda.synchronize()
            r.append(out[0].item())
        print(f"N={ncols:3d} {'TS' if mode & 1 else 'SS'} {'elect.sync' if mode & 8 else 'tid==0    '}: {r[0]:.1f} cycles/MMA over 48, {r[1]:.1f} over 480"
              f"  (math floor {128 * ncols * 8 / 2048:.0f} at 2048 tf32 FMA/clk/SM)")
