"""Where the end-to-end time of one solveBatch goes (host buffers -> host result), per phase, for a workload.
Exploration tool: prints milliseconds per phase (each phase bracketed by a device synchronise)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import icnn_b200
from icnn_b200 import bundle_entropy, workloads

name = sys.argv[1] if len(sys.argv) > 1 else "C4"
cfg = workloads.CONFIGS[name]
p, x, y0 = workloads.make_inputs(name)
net = icnn_b200.PICNN.from_params(p)
x_pin = torch.from_numpy(x.astype(np.float32)).pin_memory()
y0_pin = torch.from_numpy(y0).pin_memory()
work = torch.empty_like(y0_pin).pin_memory()


def sync():
    torch.cuda.synchronize()
    return time.perf_counter()


st = None
for rep in range(4):
    work.copy_(y0_pin)
    t0 = sync()
    fg = net.bind(x_pin, affine=cfg["affine"])
    t1 = sync()
    out = bundle_entropy.solveBatch(fg, work, nIter=cfg["nIter"], variant=cfg["variant"], return_state=True, state=st)
    t2 = sync()
    st = out[-1]
    if rep == 3:
        print("%s: bind (H2D x + gates) %.2f ms | solveBatch %.2f ms | total %.2f ms" % (name, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t2 - t0) * 1e3))
import cProfile, pstats
work.copy_(y0_pin)
pr = cProfile.Profile()
pr.enable()
fg = net.bind(x_pin, affine=cfg["affine"])
out = bundle_entropy.solveBatch(fg, work, nIter=cfg["nIter"], variant=cfg["variant"], return_state=True, state=st)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
