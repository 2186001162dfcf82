"""Device attributes that bound the persisting-L2 window (printed for the record)."""
import ctypes, json
rt = ctypes.CDLL("libcudart.so")
def attr(a):
    v = ctypes.c_int(0)
    rt.cudaDeviceGetAttribute(ctypes.byref(v), a, 0)
    return v.value
print(json.dumps({"l2CacheSize": attr(38), "maxPersistingL2CacheSize": attr(108), "maxAccessPolicyWindowSize": attr(109)}))
