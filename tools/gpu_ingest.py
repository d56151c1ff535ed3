"""Dev script: the ingest kernels alone (for `ncu --set full -k regex:ingest`) and their CUDA-event times."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cross-section_b200"))
import numpy as np, torch
import xs3d
from xs3d import synthetic as syn, _abi
vol, verts, tans = syn.make_neurite(512)
L = _abi.lib()
V = xs3d.Volume(shape=vol.shape)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
dv = torch.from_numpy(np.ascontiguousarray(vol.view(np.uint8).ravel(order="F"))).cuda()
cv = np.ascontiguousarray(vol)
for name, fn in (("bits, Fortran order (ingest_bits_f32_kernel)", lambda: V.load(vol)),
                 ("bits, C order (ingest_bits_c_kernel)", lambda: V.load(cv)),
                 ("uint8 in HBM, Fortran order (ingest_f16_kernel)", lambda: _abi.check(L.xs3d_volume_load(V._h, dv.data_ptr(), _abi.DEVICE, 0))),
                 ("uint8 in HBM, C order (ingest_c16_kernel)", lambda: _abi.check(L.xs3d_volume_load(V._h, dv.data_ptr(), _abi.DEVICE, 1)))):
  best = 1e9
  for i in range(3):
    flush.fill_(i); torch.cuda.synchronize()
    fn()
    best = min(best, V.timing()["ingest_ms"])
  print("%-55s %.4f ms" % (name, best))
# awkward sizes, device-resident uint8 image: bytes -> bits on the device, then the any-size bit ingest
rng = np.random.default_rng(0)
for shape in ((500, 500, 500), (501, 499, 503)):
  vol2 = rng.random(shape) < 0.3
  with xs3d.Volume(shape=shape) as V2:
    for name, arr, c in (("F", np.asfortranarray(vol2), 0), ("C", np.ascontiguousarray(vol2), 1)):
      dv2 = torch.from_numpy(arr.view(np.uint8).ravel(order="K").copy()).cuda()
      best = 1e9
      for i in range(3):
        flush.fill_(i); torch.cuda.synchronize()
        _abi.check(L.xs3d_volume_load(V2._h, dv2.data_ptr(), _abi.DEVICE, c))
        best = min(best, V2.timing()["ingest_ms"])
      print("uint8 in HBM %s, %s order (pack_bits_device + bit ingest): %.4f ms  (%.0f GB/s algorithmic)" % (shape, name, best, vol2.size * 1.125 / best / 1e6))
