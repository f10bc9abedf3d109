"""Scratch: what is cudaLimitMaxL2FetchGranularity inside a torch process?"""
import ctypes, torch
torch.zeros(1, device="cuda")
rt = ctypes.CDLL("libcudart.so.12")
v = ctypes.c_size_t(0)
print("rc", rt.cudaDeviceGetLimit(ctypes.byref(v), 5), "cudaLimitMaxL2FetchGranularity =", v.value)
