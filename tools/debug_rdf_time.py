import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from neuralmelting_b200 import engine as nm
n, sb, ns = 4000, 64, 512
rng = np.random.default_rng(5)
box = rng.uniform(15.2, 20.5, ns).astype(np.float32)
pos = (rng.uniform(0, 1, (ns, n, 3)) * box[:, None, None]).astype(np.float32)
r = np.linspace(1e-16, 1 / 2, sb) * np.float32(15.2)
d_pos = torch.from_numpy(pos).cuda(); d_box = torch.from_numpy(box).cuda()
d_cnt = torch.zeros((ns, sb), dtype=torch.int32, device="cuda")
for rep in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    nm.rdf_counts_device(d_pos.data_ptr(), d_box.data_ptr(), n, ns, r, d_cnt.data_ptr())
    torch.cuda.synchronize(); t1 = time.perf_counter()
    got = nm.rdf_counts(pos, box, r)
    t2 = time.perf_counter()
    print("device path %.1f ms   host path %.1f ms   equal %s" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, np.array_equal(got, d_cnt.cpu().numpy().astype(np.uint32))))
# same with box sorted (does sample order / box size matter?)
print("mean box", box.mean())
