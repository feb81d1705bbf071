import sys
sys.argv=['x']
sys.path.insert(0,'/root/repo/tools'); sys.path.insert(0,'/root/repo')
import bench_configs as bc, torch
torch.cuda.set_device(0)
for c in [("A", 1024, 1, False, 65536), ("B-sym", 4096, 3, False, 65536), ("D", 16384, 6, False, 16384)]:
    r=bc.full_path(*c); print(r['config'], round(r['ms_per_step'],2), {k: round(v,2) for k,v in r['kernels_ms'].items()})
