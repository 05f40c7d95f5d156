"""Times as_apply_action (ENV:257-274, four calls per env step) against its HBM roofline: 168 bytes per env."""
import sys, json, torch
sys.path.insert(0, '.')
from allsteps_isaaclab_b200.mdp import AllstepsMDP
dev = torch.device('cuda:0')
for N in (4096, 65536, 1048576):
    m = AllstepsMDP(N, device=dev, seed=3)
    sets = [(-1.5 + 3.0 * torch.rand(N, 21, device=dev), torch.empty(N, 21, device=dev)) for _ in range(6)]
    for a, e in sets:
        m.apply_action(a, e)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    it = 120
    e0.record()
    for i in range(it):
        a, e = sets[i % len(sets)]
        m.apply_action(a, e)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / it * 1e3
    print(json.dumps({"envs": N, "us": us, "GBps": N * 168 / (us * 1e-6) / 1e9, "frac_of_6453": N * 168 / (us * 1e-6) / 1e9 / 6453.1}))
