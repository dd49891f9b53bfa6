import torch, os
from torch._C._distributed_c10d import _SymmetricMemory as S
from torch._C._autograd import DeviceType
print("has_multicast_support", S.has_multicast_support(DeviceType.CUDA, 0))
try:
    from cuda.bindings import driver as cu
except Exception:
    from cuda import cuda as cu
cu.cuInit(0)
err, dev = cu.cuDeviceGet(0)
for name in ("CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED", "CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_FABRIC_SUPPORTED", "CU_DEVICE_ATTRIBUTE_VIRTUAL_MEMORY_MANAGEMENT_SUPPORTED", "CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR_SUPPORTED"):
    a = getattr(cu.CUdevice_attribute, name)
    print(name, cu.cuDeviceGetAttribute(a, dev))
print("can_access_peer 0->1", torch.cuda.can_device_access_peer(0, 1) if torch.cuda.device_count() > 1 else None)
os.system("nvidia-smi topo -m | head -12; nvidia-smi -q | grep -i -A3 fabric | head -12")
