import sys, os, numpy as np, torch
sys.path.insert(0, os.getcwd())
import mosdepth_b200 as m
from mosdepth_b200 import hostio
import bench
bam, n_reads, n_bytes = bench.make_bam("c3", 0.03, 20261022, 16)
hdr = hostio.read_bam_header(bam)
raw = np.fromfile(bam, dtype=np.uint8)
host = torch.empty(raw.size, dtype=torch.uint8, pin_memory=True); host.numpy()[:] = raw
ctx = m.Context(0); ctx.open_mem(host, hdr); ctx.set_filters(mode=m.MODE_FAST)
for i in range(3):
    if i == 2: os.environ["MSD_TRACE"] = "1"
    info = ctx.run()
print(info.as_dict())
