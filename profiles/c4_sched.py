"""Step time of a configuration for several (line contexts, frames per line call, frames per step) choices:
python profiles/c4_sched.py c4 NL:SUB:B[:what] ...   (what = both | line | orb)"""
import os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench, spl_slam_b200 as S
cfgname = sys.argv[1]
for spec in sys.argv[2:]:
    f = spec.split(":")
    NL, SUB, B = int(f[0]), int(f[1]), int(f[2]); what = f[3] if len(f) > 3 else "both"
    args = types.SimpleNamespace(orb_contexts=bench.ORB_CONTEXTS, line_contexts=NL, line_sub=SUB, orb_sub=0)
    ex = bench.Extraction(S, torch, bench.CONFIGS[cfgname], B, 0, 1, 0, args)
    ex.run_device(2, what)
    t = ex.run_device(4, what) / 4
    free, tot = torch.cuda.mem_get_info()
    print("%s NL=%d SUB=%d B=%d %s: %.2f ms per step, %.0f frames/s, %.1f GB in use" % (cfgname, ex.NL, ex.SUB, B, what, t, B / t * 1e3, (tot - free) / 1e9), flush=True)
    ex.close(); del ex
    torch.cuda.empty_cache()
