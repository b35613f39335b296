import os, sys, json, ctypes
sys.path.insert(0, "/root/repo")
import torch
import bench
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
torch.zeros(1, device=dev)
rt = ctypes.CDLL("libcudart.so.12")
for gran in (64, 32, 128):
    rc = rt.cudaDeviceSetLimit(5, ctypes.c_size_t(gran))   # cudaLimitMaxL2FetchGranularity = 0x05
    val = ctypes.c_size_t()
    rt.cudaDeviceGetLimit(ctypes.byref(val), 5)
    fam = bench.family_numbers(dev, 6539.2)
    print("granularity", gran, "rc", rc, "now", val.value, {k: v["us_per_step"] for k, v in fam.items() if isinstance(v, dict) and ("262144" in k or "1048576" in k or "task" in k)})
