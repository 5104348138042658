"""Torch-free sanity check of the built library on a GPU box (a few seconds: no `import torch`): the library loads, sees
the device, and one of its sm_100a kernels (scann_transpose_blocks) runs and returns the right bytes.
usage: python tools/so_sanity.py"""
import ctypes, os, sys, time
import numpy as np
t0 = time.time()
here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = ctypes.CDLL(os.path.join(here, "scann_b200", "libscann_b200.so"))
lib.scann_last_error.restype = ctypes.c_char_p
print("version", lib.scann_version(), "SMs", lib.scann_device_sm_count(), "cc", lib.scann_device_cc(), flush=True)
rt = ctypes.CDLL("libcudart.so.12")
def malloc(n):
    p = ctypes.c_void_p()
    assert rt.cudaMalloc(ctypes.byref(p), ctypes.c_size_t(n)) == 0
    return p
n = 128 * 128
src_h = np.arange(2 * n, dtype=np.float32)
off_h = np.array([0, n], np.int32)
src, dst, off = malloc(src_h.nbytes), malloc(src_h.nbytes), malloc(off_h.nbytes)
assert rt.cudaMemcpy(src, src_h.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(src_h.nbytes), 1) == 0
assert rt.cudaMemcpy(off, off_h.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(off_h.nbytes), 1) == 0
rc = lib.scann_transpose_blocks(src, dst, off, 2, None)
print("scann_transpose_blocks rc", rc, lib.scann_last_error() if rc else "", flush=True)
assert rt.cudaDeviceSynchronize() == 0
out = np.empty_like(src_h)
assert rt.cudaMemcpy(out.ctypes.data_as(ctypes.c_void_p), dst, ctypes.c_size_t(out.nbytes), 2) == 0
want = np.concatenate([src_h[:n].reshape(128, 128).T.ravel(), src_h[n:].reshape(128, 128).T.ravel()])
print("transpose correct:", bool(np.array_equal(out, want)), "in %.1f s" % (time.time() - t0))
sys.exit(0 if np.array_equal(out, want) and rc == 0 else 1)
