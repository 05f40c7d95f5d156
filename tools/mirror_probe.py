import sys, json, torch
sys.path.insert(0, '.')
import bench
from allsteps_isaaclab_b200.mdp import AllstepsMDP
dev = torch.device('cuda:0')
m = AllstepsMDP(4096, device=dev, seed=3)
for rows, it in ((32*4096, 200), (32*65536, 40)):
    print(json.dumps(bench.time_mirror(torch, m, m.cfg, rows, it, 6453.1)))
