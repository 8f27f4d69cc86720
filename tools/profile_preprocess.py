"""One launch of the Atari preprocessing kernel (256 streams) between cudaProfilerStart/Stop, for
`ncu --profile-from-start off --set full ... python tools/profile_preprocess.py`."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dqn_zoo_b200 import _lib, processors

n, H, W = 256, 210, 160
dev = torch.device('cuda:0')
pre = processors.BatchedAtariPreprocessor(num_streams=n, device=dev, device_observations=True)
pre._allocate((H, W, 3))
pre._raw.copy_(torch.randint(0, 256, pre._raw.shape, dtype=torch.uint8, device=dev))
a = torch.tensor([pre._raw[e, 0].data_ptr() for e in range(n)], dtype=torch.int64, device=dev)
b = torch.tensor([pre._raw[e, 1].data_ptr() for e in range(n)], dtype=torch.int64, device=dev)
s = torch.tensor([pre._stacks[e].data_ptr() for e in range(n)], dtype=torch.int64, device=dev)
counts = torch.full((n,), 4, dtype=torch.int32, device=dev)


def launch():
  _lib.call('dz_atari_preprocess', a.data_ptr(), b.data_ptr(), n, C.byref(pre._axis_h.c), C.byref(pre._axis_v.c),
            s.data_ptr(), counts.data_ptr(), 4, C.cast(pre._luma, C.c_void_p), pre._max_band_rows,
            torch.cuda.current_stream().cuda_stream)


for _ in range(3):
  launch()
torch.cuda.synchronize()
torch.cuda.profiler.start()
launch()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
