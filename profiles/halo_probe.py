"""Probe: 3x3 conv A operand served from one halo tile in smem via row-offset UMMA descriptors (see kd_experiments.cu)."""
import ctypes, math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
import subprocess

# test-only library: the probe kernel is NOT part of the product libkidney_b200.so
HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libkd_experiments.so")
if not os.path.exists(SO):
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-shared",
                           "-I", os.path.join(HERE, "..", "kidney_diffusion_b200", "csrc"), "-o", SO, os.path.join(HERE, "csrc", "kd_experiments.cu"),
                           os.path.join(HERE, "..", "kidney_diffusion_b200", "csrc", "kd_abi.cu"), "-lcuda"])
lib = ctypes.CDLL(SO)
lib.kd_exp_halo_probe.restype = ctypes.c_int
lib.kd_exp_halo_probe.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                  ctypes.c_int, ctypes.c_void_p]
g = torch.Generator().manual_seed(0)
H = W = 48
x = torch.randn(1, 64, H, W, generator=g).half().float()
w = (torch.randn(128, 64, 3, 3, generator=g) / math.sqrt(576)).half().float()
ref = F.conv2d(x, w, padding=1)[0]  # [128, H, W]
xd = x.permute(0, 2, 3, 1).contiguous().half().cuda()
wd = w.permute(0, 2, 3, 1).reshape(128, 576).contiguous().half().cuda()
for (h0, w0) in [(8, 16), (0, 0), (32, 40)]:
    tile = ref[:, h0:h0 + 16, w0:w0 + 8].permute(1, 2, 0).reshape(128, 128)
    for variant in (0, 1, 2):
        out = torch.zeros(128, 128, device="cuda")
        rc = lib.kd_exp_halo_probe(xd.data_ptr(), H, W, wd.data_ptr(), out.data_ptr(), h0, w0, variant, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        err = float((out.cpu() - tile).norm() / tile.norm())
        print(f"tile ({h0},{w0}) variant {variant}: rc={rc} rel_l2={err:.3e}")
