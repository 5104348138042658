import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scann_b200._abi import lib, check, require_gpu
require_gpu()
out = torch.zeros(4, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for ncols in (128, 64):
    for mode in (1, 4, 5, 7):
        for nmma in (48, 480):
            check(lib.scann_tc_time(out.data_ptr(), mode, nmma, ncols, st)); torch.cuda.synchronize()
            print(f"N={ncols} mode={mode} ({'TS' if mode&1 else 'SS'}, cg stride {128 if mode&2 else 144}, {'fast' if mode&4 else 'slow'} issue) nmma={nmma}: {out[0].item():.1f} cycles/MMA")
