"""Checks scann_tc_probe (tcgen05 tile product) against fp64 on the GPU box."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scann_b200 import _abi
from scann_b200._abi import lib, check

def trunc_tf32(x):
    return (x.view(np.uint32) & np.uint32(0xffffe000)).view(np.float32)

def rna_tf32(x):
    u = x.view(np.uint32).astype(np.uint64) + 0x1000
    return (u & 0xffffe000).astype(np.uint32).view(np.float32)

def main():
    _abi.require_gpu()
    rng = np.random.default_rng(0)
    A = rng.standard_normal((128, 128)).astype(np.float32)
    W = (rng.standard_normal((128, 128)) * 0.1).astype(np.float32)
    Ad, Wd = torch.from_numpy(A).cuda(), torch.from_numpy(W).cuda()
    st = torch.cuda.current_stream().cuda_stream
    refs = {0: A.astype(np.float64) @ W.astype(np.float64), 1: A.astype(np.float64).T @ W.astype(np.float64),
            2: A.astype(np.float64) @ W.astype(np.float64).T}
    refs[3] = refs[0]
    ok = True
    for layout in (0, 1, 2, 3):
        for nprod in (1, 3, 4, 5):
            D = torch.full((128, 128), float("nan"), device="cuda")
            check(lib.scann_tc_probe(Ad.data_ptr(), Wd.data_ptr(), D.data_ptr(), layout, nprod, st), "tc_probe")
            torch.cuda.synchronize()
            d = D.cpu().numpy().astype(np.float64)
            ref = refs[layout]
            err = np.abs(d - ref).max() / np.abs(ref).mean()
            rms = np.sqrt(((d - ref) ** 2).mean()) / np.abs(ref).mean()
            f32 = (torch.from_numpy(A if layout != 1 else np.ascontiguousarray(A.T)) @
                   torch.from_numpy(W if layout != 2 else np.ascontiguousarray(W.T))).numpy().astype(np.float64)
            e32 = np.abs(f32 - ref).max() / np.abs(ref).mean()
            msg = f"layout {layout} nprod {nprod}: max err / mean|ref| = {err:.3e} rms {rms:.3e} (fp32 CPU matmul: {e32:.3e})"
            if nprod == 1:
                def prod(fa, fw):
                    a, w = fa(A).astype(np.float64), fw(W).astype(np.float64)
                    return {0: a @ w, 1: a.T @ w, 2: a @ w.T, 3: a @ w}[layout]
                e_tr = np.abs(d - prod(trunc_tf32, trunc_tf32)).max() / np.abs(ref).mean()
                e_rn = np.abs(d - prod(rna_tf32, rna_tf32)).max() / np.abs(ref).mean()
                msg += f"  | vs truncated operands {e_tr:.3e}, vs round-to-nearest operands {e_rn:.3e}"
            print(msg)
            lim = 5e-3 if nprod == 1 else (2e-5 if nprod in (3, 5) else 5e-6)
            ok &= bool(err < lim)
    print("PROBE", "OK" if ok else "FAILED")

if __name__ == "__main__":
    main()
