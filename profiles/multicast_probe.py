import ctypes
cu = ctypes.CDLL("libcuda.so.1")
cu.cuInit(0)
n = ctypes.c_int(); cu.cuDeviceGetCount(ctypes.byref(n)); print("devices", n.value)
for d in range(n.value):
    v = ctypes.c_int(-1)
    # CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED = 132, HANDLE_TYPE_FABRIC_SUPPORTED = 128, POSIX_FD = 103, VMM = 102
    out = {}
    for name, a in (("vmm", 102), ("posix_fd", 103), ("fabric", 128), ("multicast", 132)):
        r = cu.cuDeviceGetAttribute(ctypes.byref(v), a, d); out[name] = (r, v.value)
    print(d, out)
